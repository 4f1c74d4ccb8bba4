"""Host-side task registry: everything about a Meta-World V3 task that is not physics.

One `TaskSpec` per reference env class (metaworld/envs/sawyer_*_v3.py): the MJCF file, the static body
its ``reset_model`` moves, the named frames its observation / reward code reads (-> device frame slots
F_TASK0..), the constants from its ``__init__`` (hand_init_pos, hand/obj/goal boxes) and the size of its
``rand_vec``.  ``task_id`` selects the device obs / reward / reset code (csrc/mw_tasks_gen.cuh).
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np


@dataclass
class TaskSpec:
    name: str
    task_id: int
    xml: str
    movable: str | None
    frames: list
    hand_init_pos: tuple
    hand_low: tuple
    hand_high: tuple
    obj_low: tuple
    obj_high: tuple
    goal_low: tuple
    goal_high: tuple
    main_geom: str | None = "objGeom"
    params: tuple = ()
    implemented: bool = True

    @property
    def rand_low(self):
        return np.hstack((self.obj_low, self.goal_low)).astype(np.float64)

    @property
    def rand_high(self):
        return np.hstack((self.obj_high, self.goal_high)).astype(np.float64)


def _t(name, tid, xml, movable, frames, hand_init, hand_lo, hand_hi, obj_lo, obj_hi, goal_lo, goal_hi, **kw):
    return TaskSpec(name, tid, xml, movable, frames, hand_init, hand_lo, hand_hi, obj_lo, obj_hi, goal_lo, goal_hi, **kw)


_OBJ = [("body", "obj"), ("geom", "objGeom")]

TASKS = {t.name: t for t in [
    # metaworld/envs/sawyer_reach_v3.py:40-75
    _t("reach-v3", 0, "sawyer_reach_v3", None, _OBJ, (0.0, 0.6, 0.2), (-0.5, 0.40, 0.05), (0.5, 1, 0.5),
       (-0.1, 0.6, 0.02), (0.1, 0.7, 0.02), (-0.1, 0.8, 0.05), (0.1, 0.9, 0.3)),
]}

TASK_IDS = {t.name: t.task_id for t in TASKS.values()}
