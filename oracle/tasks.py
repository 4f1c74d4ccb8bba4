"""CPU oracle: per-task restatements (reset_model / obs getters / evaluate_state /
compute_reward) of the reference's ``metaworld/envs/sawyer_*_v3.py``.

TEST INFRASTRUCTURE -- PARITY UNPINNED (see oracle/mjphys.h).  Each class cites
the reference file it follows; only the v2 reward (the default,
metaworld/__init__.py:410) is restated.
"""
from __future__ import annotations

import numpy as np

from oracle.sawyer_env import SawyerXYZEnv, hamacher_product, mat2quat_xyzw, rect_prism_tolerance, tolerance

A = np.array
norm = np.linalg.norm


class _FreeObjMixin:
    """Tasks whose object is body 'obj' with geom 'objGeom' (reach/push/pick-place family)."""

    def _get_pos_objects(self):
        return self.get_body_com("obj")

    def _get_quat_objects(self):
        return mat2quat_xyzw(self.data.geom("objGeom").xmat)

    def fix_extreme_obj_pos(self, orig_init_pos):
        diff = self.get_body_com("obj")[:2] - self.get_body_com("obj")[:2]
        adjusted = orig_init_pos[:2] + diff
        return A([adjusted[0], adjusted[1], self.get_body_com("obj")[-1]])


class Reach(_FreeObjMixin, SawyerXYZEnv):
    """metaworld/envs/sawyer_reach_v3.py"""
    xml = "sawyer_reach_v3"
    hand_low, hand_high = (-0.5, 0.40, 0.05), (0.5, 1, 0.5)
    obj_low, obj_high = (-0.1, 0.6, 0.02), (0.1, 0.7, 0.02)
    goal_low, goal_high = (-0.1, 0.8, 0.05), (0.1, 0.9, 0.3)

    def setup(self):
        self.init_config = dict(obj_init_angle=0.3, obj_init_pos=A([0.0, 0.6, 0.02]), hand_init_pos=A([0.0, 0.6, 0.2]))
        self.goal = A([-0.1, 0.8, 0.2])
        self.obj_init_pos = self.init_config["obj_init_pos"]
        self.hand_init_pos = self.init_config["hand_init_pos"]

    def random_reset_space(self):
        return np.hstack((self.obj_low, self.goal_low)), np.hstack((self.obj_high, self.goal_high))

    def reset_model(self):
        self._reset_hand()
        self._target_pos = self.goal.copy()
        self.obj_init_pos = self.fix_extreme_obj_pos(self.init_config["obj_init_pos"])
        goal_pos = self._get_state_rand_vec()
        self._target_pos = goal_pos[3:]
        while norm(goal_pos[:2] - self._target_pos[:2]) < 0.15:
            goal_pos = self._get_state_rand_vec()
            self._target_pos = goal_pos[3:]
        self._target_pos = goal_pos[-3:]
        self.obj_init_pos = goal_pos[:3]
        self._set_obj_xyz(self.obj_init_pos)
        self.model.site("goal").pos = self._target_pos
        return self._get_obs()

    def evaluate_state(self, obs, action):
        reward, reach_dist, in_place = self.compute_reward(action, obs)
        return reward, dict(success=float(reach_dist <= 0.05), near_object=reach_dist, grasp_success=1.0,
                            grasp_reward=reach_dist, in_place_reward=in_place, obj_to_target=reach_dist,
                            unscaled_reward=reward)

    def compute_reward(self, actions, obs):
        tcp = self.tcp_center
        target = self._target_pos
        tcp_to_target = float(norm(tcp - target))
        in_place_margin = float(norm(self.hand_init_pos - target))
        in_place = tolerance(tcp_to_target, bounds=(0, 0.05), margin=in_place_margin, sigmoid="long_tail")
        return 10 * in_place, tcp_to_target, in_place


TASKS = {"reach-v3": Reach}
