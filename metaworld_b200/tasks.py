"""Host-side task registry: everything about a Meta-World V3 task that is not physics.

One `TaskSpec` per reference env class (metaworld/envs/sawyer_*_v3.py): the MJCF file, the static body
its ``reset_model`` moves, the named frames its observation / reward code reads (-> device frame slots
F_TASK0..), the constants from its ``__init__`` (hand_init_pos, hand/goal boxes, random reset space), its
rand_vec rejection rule, and ``task_id`` which selects the device obs / reward / reset code
(csrc/mw_tasks_gen.cuh).
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np


@dataclass
class TaskSpec:
    name: str
    task_id: int
    xml: str
    movable: str | None          # static body whose model.body(name).pos the task's reset_model rewrites
    frames: list                 # [(kind, name)] -> device frame slots F_TASK0 + i
    hand_init_pos: tuple
    hand_low: tuple              # mocap_low/high default to hand_low/high (sawyer_xyz_env.py:196-201)
    hand_high: tuple
    rand_lo: tuple               # _random_reset_space
    rand_hi: tuple
    goal_low: tuple              # goal_space (obs clipping)
    goal_high: tuple
    reject: tuple | None = None  # ((a0,a1),(b0,b1) or fixed point,thr): redraw while |v[a0:a1]-other| < thr
    main_geom: str | None = "objGeom"
    params: tuple = ()

    @property
    def rand_low(self):
        return np.asarray(self.rand_lo, dtype=np.float64)

    @property
    def rand_high(self):
        return np.asarray(self.rand_hi, dtype=np.float64)


_OBJ = [("body", "obj"), ("geom", "objGeom")]
_XY = ((0, 2), (3, 5))   # obj xy vs goal xy


def _cat(a, b):
    return tuple(a) + tuple(b)


_HL, _HH = (-0.5, 0.40, 0.05), (0.5, 1, 0.5)        # hand_low / hand_high shared by every V3 task
_H0 = (0.0, 0.6, 0.2)

_SPECS = [
    # ---- reach / push / pick-place family (sawyer_reach_v3.py, sawyer_push_v3.py, sawyer_pick_place_v3.py)
    TaskSpec("reach-v3", 0, "sawyer_reach_v3", None, _OBJ, _H0, _HL, _HH,
             _cat((-0.1, 0.6, 0.02), (-0.1, 0.8, 0.05)), _cat((0.1, 0.7, 0.02), (0.1, 0.9, 0.3)),
             (-0.1, 0.8, 0.05), (0.1, 0.9, 0.3), reject=(_XY[0], _XY[1], 0.15)),
    TaskSpec("push-v3", 1, "sawyer_push_v3", None, _OBJ, _H0, _HL, _HH,
             _cat((-0.1, 0.6, 0.02), (-0.1, 0.8, 0.01)), _cat((0.1, 0.7, 0.02), (0.1, 0.9, 0.02)),
             (-0.1, 0.8, 0.01), (0.1, 0.9, 0.02), reject=(_XY[0], _XY[1], 0.15)),
    TaskSpec("pick-place-v3", 2, "sawyer_pick_place_v3", None, _OBJ, _H0, _HL, _HH,
             _cat((-0.1, 0.6, 0.02), (-0.1, 0.8, 0.05)), _cat((0.1, 0.7, 0.02), (0.1, 0.9, 0.3)),
             (-0.1, 0.8, 0.05), (0.1, 0.9, 0.3), reject=(_XY[0], _XY[1], 0.15)),
    # ---- sawyer_door_v3.py (door-open)
    TaskSpec("door-open-v3", 3, "sawyer_door_pull", "door", [("geom", "handle")], _H0, _HL, _HH,
             (0.0, 0.85, 0.15), (0.1, 0.95, 0.15), (-0.3, 0.4, 0.1499), (-0.2, 0.5, 0.1501), main_geom=None),
    # ---- sawyer_drawer_open_v3.py / sawyer_drawer_close_v3.py  (goal_space = hand box)
    TaskSpec("drawer-open-v3", 4, "sawyer_drawer", "drawer", [("body", "drawer_link")], _H0, _HL, _HH,
             (-0.1, 0.9, 0.0), (0.1, 0.9, 0.0), _HL, _HH),
    TaskSpec("drawer-close-v3", 5, "sawyer_drawer", "drawer", [("body", "drawer_link")], _H0, _HL, _HH,
             (-0.1, 0.9, 0.0), (0.1, 0.9, 0.0), _HL, _HH),
    # ---- sawyer_button_press_topdown_v3.py
    TaskSpec("button-press-topdown-v3", 6, "sawyer_button_press_topdown", "box",
             [("body", "button"), ("site", "hole"), ("site", "buttonStart")], (0, 0.4, 0.2), _HL, _HH,
             (-0.1, 0.8, 0.115), (0.1, 0.9, 0.115), _HL, _HH, main_geom="btnGeom"),
    # ---- sawyer_peg_insertion_side_v3.py
    TaskSpec("peg-insert-side-v3", 7, "sawyer_peg_insertion_side", "box",
             [("site", "pegGrasp"), ("site", "pegHead"), ("body", "peg"),
              ("site", "bottom_right_corner_collision_box_1"), ("site", "top_left_corner_collision_box_1"),
              ("site", "bottom_right_corner_collision_box_2"), ("site", "top_left_corner_collision_box_2")],
             _H0, _HL, _HH, _cat((0.0, 0.5, 0.02), (-0.35, 0.4, -0.001)), _cat((0.2, 0.7, 0.02), (-0.25, 0.7, 0.001)),
             (-0.32, 0.4, 0.129), (-0.22, 0.7, 0.131), reject=(_XY[0], _XY[1], 0.1), main_geom=None),
    # ---- sawyer_window_open_v3.py / sawyer_window_close_v3.py
    TaskSpec("window-open-v3", 8, "sawyer_window_horizontal", "window", [("site", "handleOpenStart")], (0, 0.4, 0.2), _HL, _HH,
             (-0.1, 0.7, 0.16), (0.1, 0.9, 0.16), _HL, _HH, main_geom=None),
    TaskSpec("window-close-v3", 9, "sawyer_window_horizontal", "window", [("site", "handleCloseStart")], (0, 0.4, 0.2), _HL, _HH,
             (0.0, 0.75, 0.2), (0.0, 0.9, 0.2), _HL, _HH, main_geom=None),
    # ---- wall variants (sawyer_reach_wall_v3.py, sawyer_push_wall_v3.py, sawyer_pick_place_wall_v3.py)
    TaskSpec("reach-wall-v3", 10, "sawyer_reach_wall_v3", None, _OBJ, _H0, _HL, _HH,
             _cat((-0.05, 0.6, 0.015), (-0.05, 0.85, 0.05)), _cat((0.05, 0.65, 0.015), (0.05, 0.9, 0.3)),
             (-0.05, 0.85, 0.05), (0.05, 0.9, 0.3), reject=(_XY[0], _XY[1], 0.15)),
    TaskSpec("push-wall-v3", 11, "sawyer_push_wall_v3", None, _OBJ, _H0, _HL, _HH,
             _cat((-0.05, 0.6, 0.015), (-0.05, 0.85, 0.01)), _cat((0.05, 0.65, 0.015), (0.05, 0.9, 0.02)),
             (-0.05, 0.85, 0.01), (0.05, 0.9, 0.02), reject=(_XY[0], _XY[1], 0.15)),
    TaskSpec("pick-place-wall-v3", 12, "sawyer_pick_place_wall_v3", None, _OBJ, _H0, _HL, _HH,
             _cat((-0.05, 0.6, 0.015), (-0.05, 0.85, 0.05)), _cat((0.05, 0.65, 0.015), (0.05, 0.9, 0.3)),
             (-0.05, 0.85, 0.05), (0.05, 0.9, 0.3), reject=(_XY[0], _XY[1], 0.15)),
    # ---- sawyer_push_back_v3.py, sawyer_sweep_v3.py, sawyer_sweep_into_goal_v3.py
    TaskSpec("push-back-v3", 13, "sawyer_push_back_v3", None, _OBJ, _H0, _HL, _HH,
             _cat((-0.1, 0.8, 0.02), (-0.1, 0.6, 0.0199)), _cat((0.1, 0.85, 0.02), (0.1, 0.7, 0.0201)),
             (-0.1, 0.6, 0.0199), (0.1, 0.7, 0.0201), reject=(_XY[0], _XY[1], 0.15)),
    TaskSpec("sweep-v3", 14, "sawyer_sweep_v3", None, _OBJ, _H0, _HL, _HH,
             (-0.1, 0.6, 0.02), (0.1, 0.7, 0.02), (0.49, 0.6, 0.00), (0.51, 0.7, 0.02)),
    TaskSpec("sweep-into-v3", 15, "sawyer_table_with_hole", None, _OBJ, _H0, _HL, _HH,
             _cat((-0.1, 0.6, 0.02), (-0.001, 0.8399, 0.0199)), _cat((0.1, 0.7, 0.02), (0.001, 0.8401, 0.0201)),
             (-0.001, 0.8399, 0.0199), (0.001, 0.8401, 0.0201), reject=((0, 2), (0.0, 0.84), 0.15)),
    # ---- sawyer_hand_insert_v3.py, sawyer_pick_out_of_hole_v3.py
    TaskSpec("hand-insert-v3", 16, "sawyer_table_with_hole", None, _OBJ, _H0, (-0.5, 0.40, -0.15), _HH,
             _cat((-0.1, 0.6, 0.05), (-0.04, 0.8, -0.0201)), _cat((0.1, 0.7, 0.05), (0.04, 0.88, -0.0199)),
             (-0.04, 0.8, -0.0201), (0.04, 0.88, -0.0199), reject=(_XY[0], _XY[1], 0.15)),
    TaskSpec("pick-out-of-hole-v3", 17, "sawyer_pick_out_of_hole", None, _OBJ, _H0, (-0.5, 0.40, -0.05), _HH,
             _cat((0, 0.75, 0.02), (-0.1, 0.5, 0.15)), _cat((0, 0.75, 0.02), (0.1, 0.6, 0.3)),
             (-0.1, 0.5, 0.15), (0.1, 0.6, 0.3), reject=(_XY[0], _XY[1], 0.15)),
]

TASKS = {t.name: t for t in _SPECS}
TASK_IDS = {t.name: t.task_id for t in _SPECS}
assert sorted(TASK_IDS.values()) == list(range(len(_SPECS)))


def enum_name(name: str) -> str:
    return "T_" + name[:-3].upper().replace("-", "_")


def emit_task_enum() -> str:
    """C enum of task ids for csrc/mw_tasks_gen.cuh (written by build.write_header)."""
    lines = ["/* GENERATED by metaworld_b200/tasks.py:emit_task_enum -- do not edit. */", "#pragma once", "enum {"]
    for t in sorted(_SPECS, key=lambda t: t.task_id):
        lines.append(f"  {enum_name(t.name)} = {t.task_id},")
    lines += ["  T_NTASK", "};"]
    return "\n".join(lines) + "\n"
