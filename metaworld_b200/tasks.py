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
    reject: tuple | None = None  # ((a0,a1),(b0,b1),thr): redraw while |v[a0:a1]-v[b0:b1]| < thr
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


_SPECS = [
    # ---- metaworld/envs/sawyer_reach_v3.py:40-75,125-129
    TaskSpec("reach-v3", 0, "sawyer_reach_v3", None, _OBJ, (0.0, 0.6, 0.2), (-0.5, 0.40, 0.05), (0.5, 1, 0.5),
             _cat((-0.1, 0.6, 0.02), (-0.1, 0.8, 0.05)), _cat((0.1, 0.7, 0.02), (0.1, 0.9, 0.3)),
             (-0.1, 0.8, 0.05), (0.1, 0.9, 0.3), reject=(_XY[0], _XY[1], 0.15)),
]

TASKS = {t.name: t for t in _SPECS}
TASK_IDS = {t.name: t.task_id for t in _SPECS}
