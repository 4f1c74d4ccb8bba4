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




class Push(_FreeObjMixin, SawyerXYZEnv):
    """metaworld/envs/sawyer_push_v3.py"""
    xml = "sawyer_push_v3"
    hand_low, hand_high = (-0.5, 0.40, 0.05), (0.5, 1, 0.5)
    obj_low, obj_high = (-0.1, 0.6, 0.02), (0.1, 0.7, 0.02)
    goal_low, goal_high = (-0.1, 0.8, 0.01), (0.1, 0.9, 0.02)
    TARGET_RADIUS = 0.05

    def setup(self):
        self.init_config = dict(obj_init_pos=A([0.0, 0.6, 0.02]), hand_init_pos=A([0.0, 0.6, 0.2]))
        self.goal = A([0.1, 0.8, 0.02])
        self.obj_init_pos = self.init_config["obj_init_pos"]
        self.hand_init_pos = self.init_config["hand_init_pos"]

    def random_reset_space(self):
        return np.hstack((self.obj_low, self.goal_low)), np.hstack((self.obj_high, self.goal_high))

    def reset_model(self):
        self._reset_hand()
        self._target_pos = self.goal.copy()
        self.obj_init_pos = A(self.fix_extreme_obj_pos(self.init_config["obj_init_pos"]))
        goal_pos = self._get_state_rand_vec()
        self._target_pos = goal_pos[3:]
        while norm(goal_pos[:2] - self._target_pos[:2]) < 0.15:
            goal_pos = self._get_state_rand_vec()
            self._target_pos = goal_pos[3:]
        self._target_pos = np.concatenate([goal_pos[-3:-1], [self.obj_init_pos[-1]]])
        self.obj_init_pos = np.concatenate([goal_pos[:2], [self.obj_init_pos[-1]]])
        self._set_obj_xyz(self.obj_init_pos)
        self.model.site("goal").pos = self._target_pos
        return self._get_obs()

    def evaluate_state(self, obs, action):
        obj = obs[4:7]
        reward, tcp_to_obj, tcp_opened, target_to_obj, object_grasped, in_place = self.compute_reward(action, obs)
        info = dict(success=float(target_to_obj <= self.TARGET_RADIUS), near_object=float(tcp_to_obj <= 0.03),
                    grasp_success=float(self.touching_main_object and (tcp_opened > 0) and (obj[2] - 0.02 > self.obj_init_pos[2])),
                    grasp_reward=object_grasped, in_place_reward=in_place, obj_to_target=target_to_obj, unscaled_reward=reward)
        return reward, info

    def compute_reward(self, action, obs):
        obj, tcp_opened = obs[4:7], obs[3]
        tcp_to_obj = float(norm(obj - self.tcp_center))
        target_to_obj = float(norm(obj - self._target_pos))
        target_to_obj_init = float(norm(self.obj_init_pos - self._target_pos))
        in_place = tolerance(target_to_obj, bounds=(0, self.TARGET_RADIUS), margin=target_to_obj_init, sigmoid="long_tail")
        object_grasped = self._gripper_caging_reward(action, obj, object_reach_radius=0.01, obj_radius=0.015,
                                                     pad_success_thresh=0.05, xz_thresh=0.005, high_density=True)
        reward = 2 * object_grasped
        if tcp_to_obj < 0.02 and tcp_opened > 0:
            reward += 1.0 + reward + 5.0 * in_place
        if target_to_obj < self.TARGET_RADIUS:
            reward = 10.0
        return reward, tcp_to_obj, tcp_opened, target_to_obj, object_grasped, in_place


class PickPlace(_FreeObjMixin, SawyerXYZEnv):
    """metaworld/envs/sawyer_pick_place_v3.py"""
    xml = "sawyer_pick_place_v3"
    hand_low, hand_high = (-0.5, 0.40, 0.05), (0.5, 1, 0.5)
    obj_low, obj_high = (-0.1, 0.6, 0.02), (0.1, 0.7, 0.02)
    goal_low, goal_high = (-0.1, 0.8, 0.05), (0.1, 0.9, 0.3)

    def setup(self):
        self.init_config = dict(obj_init_pos=A([0, 0.6, 0.02]), hand_init_pos=A([0, 0.6, 0.2]))
        self.goal = A([0.1, 0.8, 0.2])
        self.hand_init_pos = self.init_config["hand_init_pos"]
        self.obj_init_pos = None

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
        self.init_tcp = self.tcp_center
        self.init_left_pad = self.get_body_com("leftpad")      # live views (reference aliasing)
        self.init_right_pad = self.get_body_com("rightpad")
        self._set_obj_xyz(self.obj_init_pos)
        self.model.site("goal").pos = self._target_pos
        return self._get_obs()

    def evaluate_state(self, obs, action):
        obj = obs[4:7]
        reward, tcp_to_obj, tcp_open, obj_to_target, grasp_reward, in_place_reward = self.compute_reward(action, obs)
        grasp_success = float(self.touching_main_object and (tcp_open > 0) and (obj[2] - 0.02 > self.obj_init_pos[2]))
        info = dict(success=float(obj_to_target <= 0.07), near_object=float(tcp_to_obj <= 0.03), grasp_success=grasp_success,
                    grasp_reward=grasp_reward, in_place_reward=in_place_reward, obj_to_target=obj_to_target, unscaled_reward=reward)
        return reward, info

    def _gripper_caging_reward(self, action, obj_pos, **unused):
        pad_success_margin, x_z_success_margin, obj_radius = 0.05, 0.005, 0.015
        tcp = self.tcp_center
        left_pad, right_pad = self.get_body_com("leftpad"), self.get_body_com("rightpad")
        delta_l = left_pad[1] - obj_pos[1]
        delta_r = obj_pos[1] - right_pad[1]
        right_margin = abs(abs(obj_pos[1] - self.init_right_pad[1]) - pad_success_margin)
        left_margin = abs(abs(obj_pos[1] - self.init_left_pad[1]) - pad_success_margin)
        right_caging = tolerance(delta_r, bounds=(obj_radius, pad_success_margin), margin=right_margin, sigmoid="long_tail")
        left_caging = tolerance(delta_l, bounds=(obj_radius, pad_success_margin), margin=left_margin, sigmoid="long_tail")
        y_caging = hamacher_product(left_caging, right_caging)
        xz = [0, 2]
        tcp_obj_xz = float(norm(tcp[xz] - obj_pos[xz]))
        margin = norm(self.obj_init_pos[xz] - self.init_tcp[xz]) - x_z_success_margin
        x_z_caging = tolerance(tcp_obj_xz, bounds=(0, x_z_success_margin), margin=margin, sigmoid="long_tail")
        gripper_closed = min(max(0, action[-1]), 1)
        caging = hamacher_product(y_caging, x_z_caging)
        gripping = gripper_closed if caging > 0.97 else 0.0
        return (hamacher_product(caging, gripping) + caging) / 2

    def compute_reward(self, action, obs):
        tcp, obj, tcp_opened, target = self.tcp_center, obs[4:7], obs[3], self._target_pos
        obj_to_target = float(norm(obj - target))
        tcp_to_obj = float(norm(obj - tcp))
        in_place = tolerance(obj_to_target, bounds=(0, 0.05), margin=norm(self.obj_init_pos - target), sigmoid="long_tail")
        object_grasped = self._gripper_caging_reward(action, obj)
        reward = hamacher_product(object_grasped, in_place)
        if tcp_to_obj < 0.02 and (tcp_opened > 0) and (obj[2] - 0.01 > self.obj_init_pos[2]):
            reward += 1.0 + 5.0 * in_place
        if obj_to_target < 0.05:
            reward = 10.0
        return reward, tcp_to_obj, tcp_opened, obj_to_target, object_grasped, in_place


class DoorOpen(SawyerXYZEnv):
    """metaworld/envs/sawyer_door_v3.py"""
    xml = "sawyer_door_pull"
    hand_low, hand_high = (-0.5, 0.40, 0.05), (0.5, 1, 0.5)
    obj_low, obj_high = (0.0, 0.85, 0.15), (0.1, 0.95, 0.15)
    goal_low, goal_high = (-0.3, 0.4, 0.1499), (-0.2, 0.5, 0.1501)

    def setup(self):
        self.obj_init_pos = A([0.1, 0.95, 0.15])
        self.hand_init_pos = A([0, 0.6, 0.2])
        self.goal = A([-0.2, 0.7, 0.15])
        j = self.model.joint("doorjoint")
        self.door_qpos_adr = int(self.model.src.arrays["jnt_qposadr"][j])
        self.door_qvel_adr = int(self.model.src.arrays["jnt_dofadr"][j])

    def random_reset_space(self):
        return A(self.obj_low), A(self.obj_high)

    _target_site_config = []

    def _get_pos_objects(self):
        return self.data.geom("handle").xpos.copy()

    def _get_quat_objects(self):
        return mat2quat_xyzw(self.data.geom("handle").xmat)

    def _set_obj_xyz(self, pos):
        qpos, qvel = self.data.qpos.copy(), self.data.qvel.copy()
        qpos[self.door_qpos_adr] = pos
        qvel[self.door_qvel_adr] = 0
        self.set_state(qpos.flatten(), qvel.flatten())

    def reset_model(self):
        self._reset_hand()
        self.obj_init_pos = self._get_state_rand_vec()
        self._target_pos = self.obj_init_pos + A([-0.3, -0.45, 0.0])
        self.model.body("door").pos = self.obj_init_pos
        self.model.site("goal").pos = self._target_pos
        self._set_obj_xyz(A(0))
        return self._get_obs()

    def evaluate_state(self, obs, action):
        reward, reward_grab, reward_ready, reward_success = self.compute_reward(action, obs)
        info = dict(success=float(abs(obs[4] - self._target_pos[0]) <= 0.08), near_object=reward_ready,
                    grasp_success=reward_grab >= 0.5, grasp_reward=reward_grab, in_place_reward=reward_success,
                    obj_to_target=0, unscaled_reward=reward)
        return reward, info

    @staticmethod
    def _reward_pos(obs, theta):
        hand = obs[:3]
        door = obs[4:7] + A([-0.05, 0, 0])
        threshold = 0.12
        radius = norm(hand[:2] - door[:2])
        floor = 0.0 if radius <= threshold else 0.04 * np.log(radius - threshold) + 0.4
        above_floor = 1.0 if hand[2] >= floor else tolerance(floor - hand[2], bounds=(0.0, 0.01), margin=floor / 2.0, sigmoid="long_tail")
        in_place = tolerance(float(norm(hand - door - A([0.05, 0.03, -0.01]))), bounds=(0, threshold / 2.0), margin=0.5, sigmoid="long_tail")
        ready_to_open = hamacher_product(above_floor, in_place)
        door_angle = -theta
        opened = 0.2 * float(theta < -np.pi / 90.0) + 0.8 * tolerance(np.pi / 2.0 + np.pi / 6 - door_angle, bounds=(0, 0.5), margin=np.pi / 3.0, sigmoid="long_tail")
        return ready_to_open, opened

    def compute_reward(self, actions, obs):
        theta = float(self.data.joint("doorjoint").qpos.item())
        reward_grab = float((np.clip(actions[3], -1, 1) + 1.0) / 2.0)
        ready, opened = self._reward_pos(obs, theta)
        reward = 2.0 * hamacher_product(ready, reward_grab) + 8.0 * opened
        if abs(obs[4] - self._target_pos[0]) <= 0.08:
            reward = 10.0
        return reward, reward_grab, ready, opened


class DrawerOpen(SawyerXYZEnv):
    """metaworld/envs/sawyer_drawer_open_v3.py"""
    xml = "sawyer_drawer"
    hand_low, hand_high = (-0.5, 0.40, 0.05), (0.5, 1, 0.5)
    obj_low, obj_high = (-0.1, 0.9, 0.0), (0.1, 0.9, 0.0)
    goal_low, goal_high = hand_low, hand_high
    maxDist = 0.2

    def setup(self):
        self.obj_init_pos = A([0.0, 0.9, 0.0], dtype=np.float32)
        self.hand_init_pos = A([0, 0.6, 0.2], dtype=np.float32)

    def random_reset_space(self):
        return A(self.obj_low), A(self.obj_high)

    def _get_pos_objects(self):
        return self.get_body_com("drawer_link") + A([0.0, -0.16, 0.0])

    def _get_quat_objects(self):
        return self.data.body("drawer_link").xquat

    def reset_model(self):
        self._reset_hand()
        self.obj_init_pos = self._get_state_rand_vec()
        self.model.body("drawer").pos = self.obj_init_pos
        self._target_pos = self.obj_init_pos + A([0.0, -0.16 - self.maxDist, 0.09])
        self.model.site("goal").pos = self._target_pos
        return self._get_obs()

    def evaluate_state(self, obs, action):
        reward, gripper_error, gripped, handle_error, caging_reward, opening_reward = self.compute_reward(action, obs)
        info = dict(success=float(handle_error <= 0.03), near_object=float(gripper_error <= 0.03), grasp_success=float(gripped > 0),
                    grasp_reward=caging_reward, in_place_reward=opening_reward, obj_to_target=handle_error, unscaled_reward=reward)
        return reward, info

    def compute_reward(self, action, obs):
        gripper, handle = obs[:3], obs[4:7]
        handle_error = float(norm(handle - self._target_pos))
        reward_for_opening = tolerance(handle_error, bounds=(0, 0.02), margin=self.maxDist, sigmoid="long_tail")
        handle_pos_init = self._target_pos + A([0.0, self.maxDist, 0.0])
        scale = A([3.0, 3.0, 1.0])
        gripper_error = (handle - gripper) * scale
        gripper_error_init = (handle_pos_init - self.init_tcp) * scale
        reward_for_caging = tolerance(float(norm(gripper_error)), bounds=(0, 0.01), margin=norm(gripper_error_init), sigmoid="long_tail")
        reward = 5.0 * (reward_for_caging + reward_for_opening)
        return reward, float(norm(handle - gripper)), obs[3], handle_error, reward_for_caging, reward_for_opening


class DrawerClose(SawyerXYZEnv):
    """metaworld/envs/sawyer_drawer_close_v3.py"""
    xml = "sawyer_drawer"
    hand_low, hand_high = (-0.5, 0.40, 0.05), (0.5, 1, 0.5)
    obj_low, obj_high = (-0.1, 0.9, 0.0), (0.1, 0.9, 0.0)
    goal_low, goal_high = hand_low, hand_high
    maxDist = 0.15

    def setup(self):
        self.obj_init_pos = A([0.0, 0.9, 0.0], dtype=np.float32)
        self.hand_init_pos = A([0, 0.6, 0.2], dtype=np.float32)

    def random_reset_space(self):
        return A(self.obj_low), A(self.obj_high)

    def _get_pos_objects(self):
        return self.get_body_com("drawer_link") + A([0.0, -0.16, 0.05])

    def _get_quat_objects(self):
        return np.zeros(4)

    def _set_obj_xyz(self, pos):
        qpos, qvel = self.data.qpos.flat.copy(), self.data.qvel.flat.copy()
        qpos[9] = pos
        self.set_state(qpos, qvel)

    def reset_model(self):
        self._reset_hand()
        self.obj_init_pos = self._get_state_rand_vec()
        self.model.body("drawer").pos = self.obj_init_pos
        self._target_pos = self.obj_init_pos + A([0.0, -0.16, 0.09])
        self._set_obj_xyz(A(-self.maxDist))
        self.obj_init_pos = self._get_pos_objects()
        self.model.site("goal").pos = self._target_pos
        return self._get_obs()

    def evaluate_state(self, obs, action):
        reward, tcp_to_obj, _, target_to_obj, object_grasped, in_place = self.compute_reward(action, obs)
        info = dict(success=float(target_to_obj <= self.TARGET_RADIUS + 0.015), near_object=float(tcp_to_obj <= 0.01), grasp_success=1.0,
                    grasp_reward=object_grasped, in_place_reward=in_place, obj_to_target=target_to_obj, unscaled_reward=reward)
        return reward, info

    def compute_reward(self, action, obs):
        obj, tcp, target = obs[4:7], self.tcp_center, self._target_pos.copy()
        target_to_obj = norm(obj - target)
        target_to_obj_init = norm(self.obj_init_pos - target)
        in_place = tolerance(target_to_obj, bounds=(0, self.TARGET_RADIUS), margin=abs(target_to_obj_init - self.TARGET_RADIUS), sigmoid="long_tail")
        handle_reach_radius = 0.005
        tcp_to_obj = float(norm(obj - tcp))
        tcp_to_obj_init = norm(self.obj_init_pos - self.init_tcp)
        reach = tolerance(tcp_to_obj, bounds=(0, handle_reach_radius), margin=abs(tcp_to_obj_init - handle_reach_radius), sigmoid="gaussian")
        gripper_closed = min(max(0, action[-1]), 1)
        reach = hamacher_product(reach, gripper_closed)
        reward = hamacher_product(reach, in_place)
        if target_to_obj <= self.TARGET_RADIUS + 0.015:
            reward = 1.0
        reward *= 10
        return reward, tcp_to_obj, 0, target_to_obj, reach, in_place


class ButtonPressTopdown(SawyerXYZEnv):
    """metaworld/envs/sawyer_button_press_topdown_v3.py"""
    xml = "sawyer_button_press_topdown"
    hand_low, hand_high = (-0.5, 0.40, 0.05), (0.5, 1, 0.5)
    obj_low, obj_high = (-0.1, 0.8, 0.115), (0.1, 0.9, 0.115)
    goal_low, goal_high = hand_low, hand_high
    _target_site_config = []

    def setup(self):
        self.obj_init_pos = A([0, 0.8, 0.115], dtype=np.float32)
        self.hand_init_pos = A([0, 0.4, 0.2], dtype=np.float32)
        self.goal = A([0, 0.88, 0.1])

    def random_reset_space(self):
        return A(self.obj_low), A(self.obj_high)

    def _get_pos_objects(self):
        return self.get_body_com("button") + A([0.0, 0.0, 0.193])

    def _get_quat_objects(self):
        return self.data.body("button").xquat

    def reset_model(self):
        from oracle import mjphys as P
        self._reset_hand()
        goal_pos = self._get_state_rand_vec()
        self.obj_init_pos = goal_pos
        self.model.body("box").pos = self.obj_init_pos
        P.mj_forward(self.model, self.data)
        self._target_pos = self._get_site_pos("hole")
        self._obj_to_target_init = abs(self._target_pos[2] - self._get_site_pos("buttonStart")[2])
        return self._get_obs()

    def evaluate_state(self, obs, action):
        reward, tcp_to_obj, tcp_open, obj_to_target, near_button, button_pressed = self.compute_reward(action, obs)
        info = dict(success=float(obj_to_target <= 0.024), near_object=float(tcp_to_obj <= 0.05), grasp_success=float(tcp_open > 0),
                    grasp_reward=near_button, in_place_reward=button_pressed, obj_to_target=obj_to_target, unscaled_reward=reward)
        return reward, info

    def compute_reward(self, action, obs):
        obj, tcp = obs[4:7], self.tcp_center
        tcp_to_obj = float(norm(obj - tcp))
        tcp_to_obj_init = float(norm(obj - self.init_tcp))
        obj_to_target = abs(self._target_pos[2] - obj[2])
        tcp_closed = 1 - obs[3]
        near_button = tolerance(tcp_to_obj, bounds=(0, 0.01), margin=tcp_to_obj_init, sigmoid="long_tail")
        button_pressed = tolerance(obj_to_target, bounds=(0, 0.005), margin=self._obj_to_target_init, sigmoid="long_tail")
        reward = 5 * hamacher_product(tcp_closed, near_button)
        if tcp_to_obj <= 0.03:
            reward += 5 * button_pressed
        return reward, tcp_to_obj, obs[3], obj_to_target, near_button, button_pressed


class PegInsertSide(SawyerXYZEnv):
    """metaworld/envs/sawyer_peg_insertion_side_v3.py"""
    xml = "sawyer_peg_insertion_side"
    hand_low, hand_high = (-0.5, 0.40, 0.05), (0.5, 1, 0.5)
    obj_low, obj_high = (0.0, 0.5, 0.02), (0.2, 0.7, 0.02)
    rgoal_low, rgoal_high = (-0.35, 0.4, -0.001), (-0.25, 0.7, 0.001)
    goal_low, goal_high = (-0.32, 0.4, 0.129), (-0.22, 0.7, 0.131)
    TARGET_RADIUS = 0.07

    def setup(self):
        self.obj_init_pos = A([0, 0.6, 0.02])
        self.hand_init_pos = A([0, 0.6, 0.2])
        self.goal = A([-0.3, 0.6, 0.0])

    def random_reset_space(self):
        return np.hstack((self.obj_low, self.rgoal_low)), np.hstack((self.obj_high, self.rgoal_high))

    def _get_pos_objects(self):
        return self._get_site_pos("pegGrasp")

    def _get_quat_objects(self):
        return mat2quat_xyzw(self.data.site("pegGrasp").xmat)

    def reset_model(self):
        self._reset_hand()
        pos_peg, pos_box = np.split(self._get_state_rand_vec(), 2)
        while norm(pos_peg[:2] - pos_box[:2]) < 0.1:
            pos_peg, pos_box = np.split(self._get_state_rand_vec(), 2)
        self.obj_init_pos = pos_peg
        self.peg_head_pos_init = self._get_site_pos("pegHead")
        self._set_obj_xyz(self.obj_init_pos)
        self.model.body("box").pos = pos_box
        self._target_pos = pos_box + A([0.03, 0.0, 0.13])
        self.model.site("goal").pos = self._target_pos
        return self._get_obs()

    def evaluate_state(self, obs, action):
        obj = obs[4:7]
        reward, tcp_to_obj, tcp_open, obj_to_target, grasp_reward, in_place_reward, _, _ = self.compute_reward(action, obs)
        grasp_success = float(tcp_to_obj < 0.02 and (tcp_open > 0) and (obj[2] - 0.01 > self.obj_init_pos[2]))
        info = dict(success=float(obj_to_target <= 0.07), near_object=float(tcp_to_obj <= 0.03), grasp_success=grasp_success,
                    grasp_reward=grasp_reward, in_place_reward=in_place_reward, obj_to_target=obj_to_target, unscaled_reward=reward)
        return reward, info

    def compute_reward(self, action, obs):
        tcp, obj, tcp_opened, target = self.tcp_center, obs[4:7], obs[3], self._target_pos
        obj_head = self._get_site_pos("pegHead")
        tcp_to_obj = float(norm(obj - tcp))
        scale = A([1.0, 2.0, 2.0])
        obj_to_target = float(norm((obj_head - target) * scale))
        in_place_margin = float(norm((self.peg_head_pos_init - target) * scale))
        in_place = tolerance(obj_to_target, bounds=(0, self.TARGET_RADIUS), margin=in_place_margin, sigmoid="long_tail")
        ip_orig = in_place
        brc1, tlc1 = self._get_site_pos("bottom_right_corner_collision_box_1"), self._get_site_pos("top_left_corner_collision_box_1")
        brc2, tlc2 = self._get_site_pos("bottom_right_corner_collision_box_2"), self._get_site_pos("top_left_corner_collision_box_2")
        cb1 = rect_prism_tolerance(curr=obj_head, one=tlc1, zero=brc1)
        cb2 = rect_prism_tolerance(curr=obj_head, one=tlc2, zero=brc2)
        collision_boxes = hamacher_product(cb2, cb1)
        in_place = hamacher_product(in_place, collision_boxes)
        object_grasped = self._gripper_caging_reward(action, obj, object_reach_radius=0.01, obj_radius=0.0075, pad_success_thresh=0.03,
                                                     xz_thresh=0.005, high_density=True)
        grasped = tcp_to_obj < 0.08 and (tcp_opened > 0) and (obj[2] - 0.01 > self.obj_init_pos[2])
        if grasped:
            object_grasped = 1.0
        reward = hamacher_product(object_grasped, in_place)
        if grasped:
            reward += 1.0 + 5 * in_place
        if obj_to_target <= 0.07:
            reward = 10.0
        return reward, tcp_to_obj, tcp_opened, obj_to_target, object_grasped, in_place, collision_boxes, ip_orig


class _Window(SawyerXYZEnv):
    xml = "sawyer_window_horizontal"
    hand_low, hand_high = (-0.5, 0.40, 0.05), (0.5, 1, 0.5)
    goal_low, goal_high = hand_low, hand_high
    TARGET_RADIUS = 0.05
    handle_site = None
    reach_sigmoid = "long_tail"

    def random_reset_space(self):
        return A(self.obj_low), A(self.obj_high)

    def _get_pos_objects(self):
        return self._get_site_pos(self.handle_site)

    def _get_quat_objects(self):
        return np.zeros(4)

    def evaluate_state(self, obs, action):
        reward, tcp_to_obj, _, target_to_obj, object_grasped, in_place = self.compute_reward(action, obs)
        info = dict(success=float(target_to_obj <= self.TARGET_RADIUS), near_object=float(tcp_to_obj <= 0.05), grasp_success=1.0,
                    grasp_reward=object_grasped, in_place_reward=in_place, obj_to_target=target_to_obj, unscaled_reward=reward)
        return reward, info

    def compute_reward(self, actions, obs):
        obj, tcp, target = self._get_pos_objects(), self.tcp_center, self._target_pos.copy()
        target_to_obj = float(abs(obj[0] - target[0]))
        target_to_obj_init = float(abs(self.init_x() - target[0]))
        in_place = tolerance(target_to_obj, bounds=(0, self.TARGET_RADIUS), margin=abs(target_to_obj_init - self.TARGET_RADIUS), sigmoid="long_tail")
        handle_radius = 0.02
        tcp_to_obj = float(norm(obj - tcp))
        tcp_to_obj_init = float(norm(self.window_handle_pos_init - self.init_tcp))
        reach = tolerance(tcp_to_obj, bounds=(0, handle_radius), margin=abs(tcp_to_obj_init - handle_radius), sigmoid=self.reach_sigmoid)
        reward = 10 * hamacher_product(reach, in_place)
        return reward, tcp_to_obj, 0.0, target_to_obj, reach, in_place


class WindowOpen(_Window):
    """metaworld/envs/sawyer_window_open_v3.py"""
    obj_low, obj_high = (-0.1, 0.7, 0.16), (0.1, 0.9, 0.16)
    handle_site = "handleOpenStart"

    def setup(self):
        self.obj_init_pos = A([-0.1, 0.785, 0.16], dtype=np.float32)
        self.hand_init_pos = A([0, 0.4, 0.2], dtype=np.float32)

    def init_x(self):
        return self.obj_init_pos[0]

    def reset_model(self):
        self._reset_hand()
        self.obj_init_pos = self._get_state_rand_vec()
        self._target_pos = self.obj_init_pos + A([0.2, 0.0, 0.0])
        self.model.body("window").pos = self.obj_init_pos
        self.window_handle_pos_init = self._get_pos_objects()
        self.data.joint("window_slide").qpos = 0.0
        self.model.site("goal").pos = self._target_pos
        return self._get_obs()


class WindowClose(_Window):
    """metaworld/envs/sawyer_window_close_v3.py"""
    obj_low, obj_high = (0.0, 0.75, 0.2), (0.0, 0.9, 0.2)
    handle_site = "handleCloseStart"
    reach_sigmoid = "gaussian"

    def setup(self):
        self.obj_init_pos = A([0.1, 0.785, 0.16], dtype=np.float32)
        self.hand_init_pos = A([0, 0.4, 0.2], dtype=np.float32)

    def init_x(self):
        return self.window_handle_pos_init[0]

    def reset_model(self):
        self._reset_hand()
        self.obj_init_pos = self._get_state_rand_vec()
        self._target_pos = self.obj_init_pos.copy()
        self.model.body("window").pos = self.obj_init_pos
        self.window_handle_pos_init = self._get_pos_objects() + A([0.2, 0.0, 0.0])
        self.data.joint("window_slide").qpos = 0.2
        self.model.site("goal").pos = self._target_pos
        return self._get_obs()



class _GeomObjMixin:
    """Object observed through geom 'objGeom' (push-wall / pick-place-wall / push-back)."""

    def _get_pos_objects(self):
        return self.data.geom("objGeom").xpos

    def _get_quat_objects(self):
        return mat2quat_xyzw(self.data.geom("objGeom").xmat)

    def adjust_initObjPos(self, orig):
        diff = self.get_body_com("obj")[:2] - self.data.geom("objGeom").xpos[:2]
        adj = orig[:2] + diff
        return A([adj[0], adj[1], self.data.geom("objGeom").xpos[-1]])


def _grip_caging(env, action, obj_pos, obj_radius, grip_add, xz_margin_c):
    """Task-local caging override shared by push-back / sweep / sweep-into / soccer
    (e.g. sawyer_push_back_v3.py:160-254): y-caging and y-gripping from the live pad positions."""
    pad_success_margin = 0.05
    grip_success_margin = obj_radius + grip_add
    tcp = env.tcp_center
    left_pad, right_pad = env.get_body_com("leftpad"), env.get_body_com("rightpad")
    dl, dr = left_pad[1] - obj_pos[1], obj_pos[1] - right_pad[1]
    rm = abs(abs(obj_pos[1] - env.init_right_pad[1]) - pad_success_margin)
    lm = abs(abs(obj_pos[1] - env.init_left_pad[1]) - pad_success_margin)
    rc = tolerance(dr, bounds=(obj_radius, pad_success_margin), margin=rm, sigmoid="long_tail")
    lc = tolerance(dl, bounds=(obj_radius, pad_success_margin), margin=lm, sigmoid="long_tail")
    rg = tolerance(dr, bounds=(obj_radius, grip_success_margin), margin=rm, sigmoid="long_tail")
    lg = tolerance(dl, bounds=(obj_radius, grip_success_margin), margin=lm, sigmoid="long_tail")
    y_caging = hamacher_product(rc, lc)
    y_gripping = hamacher_product(rg, lg)
    xz = [0, 2]
    margin = norm(A(env.obj_init_pos)[xz] - env.init_tcp[xz]) - xz_margin_c
    x_z_caging = tolerance(float(norm(tcp[xz] - obj_pos[xz])), bounds=(0, xz_margin_c), margin=margin, sigmoid="long_tail")
    caging = hamacher_product(y_caging, x_z_caging)
    gripping = y_gripping if caging > 0.95 else 0.0
    return (caging + gripping) / 2


class ReachWall(Reach):
    """metaworld/envs/sawyer_reach_wall_v3.py"""
    xml = "sawyer_reach_wall_v3"
    obj_low, obj_high = (-0.05, 0.6, 0.015), (0.05, 0.65, 0.015)
    goal_low, goal_high = (-0.05, 0.85, 0.05), (0.05, 0.9, 0.3)

    def setup(self):
        self.init_config = dict(obj_init_pos=A([0, 0.6, 0.02]), hand_init_pos=A([0, 0.6, 0.2]))
        self.goal = A([-0.05, 0.8, 0.2])
        self.obj_init_pos = self.init_config["obj_init_pos"]
        self.hand_init_pos = self.init_config["hand_init_pos"]

    def reset_model(self):
        self._reset_hand()
        goal_pos = self._get_state_rand_vec()
        self._target_pos = goal_pos[-3:]
        self.obj_init_pos = goal_pos[:3]
        self._set_obj_xyz(self.obj_init_pos)
        self.model.site("goal").pos = self._target_pos
        return self._get_obs()

    def evaluate_state(self, obs, action):
        reward, tcp_to_object, in_place = self.compute_reward(action, obs)
        return reward, dict(success=float(tcp_to_object <= 0.05), near_object=0.0, grasp_success=0.0, grasp_reward=0.0,
                            in_place_reward=in_place, obj_to_target=tcp_to_object, unscaled_reward=reward)


class PushWall(_GeomObjMixin, SawyerXYZEnv):
    """metaworld/envs/sawyer_push_wall_v3.py"""
    xml = "sawyer_push_wall_v3"
    hand_low, hand_high = (-0.5, 0.40, 0.05), (0.5, 1, 0.5)
    obj_low, obj_high = (-0.05, 0.6, 0.015), (0.05, 0.65, 0.015)
    goal_low, goal_high = (-0.05, 0.85, 0.01), (0.05, 0.9, 0.02)
    midpoint_scale = (3.0, 1.0, 1.0)

    def setup(self):
        self.init_config = dict(obj_init_pos=A([0, 0.6, 0.02]), hand_init_pos=A([0, 0.6, 0.2]))
        self.obj_init_pos = self.init_config["obj_init_pos"]
        self.hand_init_pos = self.init_config["hand_init_pos"]

    def random_reset_space(self):
        return np.hstack((self.obj_low, self.goal_low)), np.hstack((self.obj_high, self.goal_high))

    def reset_model(self):
        self._reset_hand()
        self.obj_init_pos = self.adjust_initObjPos(self.init_config["obj_init_pos"])
        goal_pos = self._get_state_rand_vec()
        self._target_pos = np.concatenate([goal_pos[-3:-1], [self.obj_init_pos[-1]]])
        self.obj_init_pos = np.concatenate([goal_pos[:2], [self.obj_init_pos[-1]]])
        self.model.site("goal").pos = self._target_pos
        self._set_obj_xyz(self.obj_init_pos)
        return self._get_obs()

    def evaluate_state(self, obs, action):
        obj = obs[4:7]
        reward, tcp_to_obj, tcp_open, obj_to_target, grasp_reward, in_place_reward = self.compute_reward(action, obs)
        grasp_success = float(self.touching_main_object and (tcp_open > 0) and (obj[2] - 0.02 > self.obj_init_pos[2]))
        return reward, dict(success=float(obj_to_target <= 0.07), near_object=float(tcp_to_obj <= 0.03), grasp_success=grasp_success,
                            grasp_reward=grasp_reward, in_place_reward=in_place_reward, obj_to_target=obj_to_target, unscaled_reward=reward)

    def compute_reward(self, action, obs):
        tcp, obj, tcp_opened, target = self.tcp_center, obs[4:7], obs[3], self._target_pos
        midpoint = A([-0.05, 0.77, obj[2]])
        tcp_to_obj = float(norm(obj - tcp))
        sc = A(self.midpoint_scale)
        o2m = float(norm((obj - midpoint) * sc))
        o2m_init = float(norm((self.obj_init_pos - midpoint) * sc))
        o2t = float(norm(obj - target))
        o2t_init = float(norm(self.obj_init_pos - target))
        p1 = tolerance(o2m, bounds=(0, 0.05), margin=o2m_init, sigmoid="long_tail")
        p2 = tolerance(o2t, bounds=(0, 0.05), margin=o2t_init, sigmoid="long_tail")
        g = self._gripper_caging_reward(action, obj, object_reach_radius=0.01, obj_radius=0.015, pad_success_thresh=0.05,
                                        xz_thresh=0.005, high_density=True)
        reward = 2 * g
        if tcp_to_obj < 0.02 and tcp_opened > 0:
            reward = 2.0 * g + 1.0 + 4.0 * p1
            if obj[1] > 0.75:
                reward = 2 * g + 1.0 + 4.0 + 3.0 * p2
        if o2t < 0.05:
            reward = 10.0
        return reward, tcp_to_obj, tcp_opened, o2t, g, p2


class PickPlaceWall(_GeomObjMixin, SawyerXYZEnv):
    """metaworld/envs/sawyer_pick_place_wall_v3.py"""
    xml = "sawyer_pick_place_wall_v3"
    hand_low, hand_high = (-0.5, 0.40, 0.05), (0.5, 1, 0.5)
    obj_low, obj_high = (-0.05, 0.6, 0.015), (0.05, 0.65, 0.015)
    goal_low, goal_high = (-0.05, 0.85, 0.05), (0.05, 0.9, 0.3)

    def setup(self):
        self.init_config = dict(obj_init_pos=A([0, 0.6, 0.02]), hand_init_pos=A([0, 0.6, 0.2]))
        self.obj_init_pos = self.init_config["obj_init_pos"]
        self.hand_init_pos = self.init_config["hand_init_pos"]

    def random_reset_space(self):
        return np.hstack((self.obj_low, self.goal_low)), np.hstack((self.obj_high, self.goal_high))

    def reset_model(self):
        self._reset_hand()
        goal_pos = self._get_state_rand_vec()
        self._target_pos = goal_pos[-3:]
        self.obj_init_pos = goal_pos[:3]
        self._set_obj_xyz(self.obj_init_pos)
        self.model.site("goal").pos = self._target_pos
        return self._get_obs()

    evaluate_state = PushWall.evaluate_state

    def compute_reward(self, action, obs):
        tcp, obj, tcp_opened, target = self.tcp_center, obs[4:7], obs[3], self._target_pos
        midpoint = A([self._target_pos[0], 0.77, 0.25])
        tcp_to_obj = float(norm(obj - tcp))
        sc = A([1.0, 1.0, 3.0])
        o2m = float(norm((obj - midpoint) * sc))
        o2m_init = float(norm((self.obj_init_pos - midpoint) * sc))
        o2t = float(norm(obj - target))
        o2t_init = float(norm(self.obj_init_pos - target))
        p1 = tolerance(o2m, bounds=(0, 0.05), margin=o2m_init, sigmoid="long_tail")
        p2 = tolerance(o2t, bounds=(0, 0.05), margin=o2t_init, sigmoid="long_tail")
        g = self._gripper_caging_reward(action=action, obj_pos=obj, obj_radius=0.015, pad_success_thresh=0.05, object_reach_radius=0.01,
                                        xz_thresh=0.005, high_density=False)
        ipg = hamacher_product(g, p1)
        reward = ipg
        if tcp_to_obj < 0.02 and (tcp_opened > 0) and (obj[2] - 0.015 > self.obj_init_pos[2]):
            reward = ipg + 1.0 + 4.0 * p1
            if obj[1] > 0.75:
                reward = ipg + 1.0 + 4.0 + 3.0 * p2
        if o2t < 0.05:
            reward = 10.0
        return reward, tcp_to_obj, tcp_opened, o2t, g, p2


class PushBack(_GeomObjMixin, SawyerXYZEnv):
    """metaworld/envs/sawyer_push_back_v3.py"""
    xml = "sawyer_push_back_v3"
    hand_low, hand_high = (-0.5, 0.40, 0.05), (0.5, 1, 0.5)
    obj_low, obj_high = (-0.1, 0.8, 0.02), (0.1, 0.85, 0.02)
    goal_low, goal_high = (-0.1, 0.6, 0.0199), (0.1, 0.7, 0.0201)
    OBJ_RADIUS, TARGET_RADIUS = 0.007, 0.05

    def setup(self):
        self.init_config = dict(obj_init_pos=A([0, 0.8, 0.02]), hand_init_pos=A([0, 0.6, 0.2], dtype=np.float32))
        self.obj_init_pos = self.init_config["obj_init_pos"]
        self.hand_init_pos = self.init_config["hand_init_pos"]

    def random_reset_space(self):
        return np.hstack((self.obj_low, self.goal_low)), np.hstack((self.obj_high, self.goal_high))

    def reset_model(self):
        self._reset_hand()
        self.obj_init_pos = self.adjust_initObjPos(self.init_config["obj_init_pos"])
        goal_pos = self._get_state_rand_vec()
        self._target_pos = np.concatenate([goal_pos[-3:-1], [self.obj_init_pos[-1]]])
        self.obj_init_pos = np.concatenate([goal_pos[:2], [self.obj_init_pos[-1]]])
        self._set_obj_xyz(self.obj_init_pos)
        self.model.site("goal").pos = self._target_pos
        return self._get_obs()

    def evaluate_state(self, obs, action):
        obj = obs[4:7]
        reward, tcp_to_obj, tcp_opened, target_to_obj, object_grasped, in_place = self.compute_reward(action, obs)
        grasp_success = float(self.touching_main_object and (tcp_opened > 0) and (obj[2] - 0.02 > self.obj_init_pos[2]))
        return reward, dict(success=float(target_to_obj <= 0.07), near_object=float(tcp_to_obj <= 0.03), grasp_success=grasp_success,
                            grasp_reward=object_grasped, in_place_reward=in_place, obj_to_target=target_to_obj, unscaled_reward=reward)

    def compute_reward(self, action, obs):
        obj, tcp_opened = obs[4:7], obs[3]
        tcp_to_obj = float(norm(obj - self.tcp_center))
        t2o = float(norm(obj - self._target_pos))
        t2o_init = float(norm(self.obj_init_pos - self._target_pos))
        in_place = tolerance(t2o, bounds=(0, self.TARGET_RADIUS), margin=t2o_init, sigmoid="long_tail")
        g = _grip_caging(self, action, obj, self.OBJ_RADIUS, 0.003, 0.01)
        reward = hamacher_product(g, in_place)
        if (tcp_to_obj < 0.01) and (0 < tcp_opened < 0.55) and (t2o_init - t2o > 0.01):
            reward += 1.0 + 5.0 * in_place
        if t2o < self.TARGET_RADIUS:
            reward = 10.0
        return reward, tcp_to_obj, tcp_opened, t2o, g, in_place


class Sweep(SawyerXYZEnv):
    """metaworld/envs/sawyer_sweep_v3.py"""
    xml = "sawyer_sweep_v3"
    hand_low, hand_high = (-0.5, 0.40, 0.05), (0.5, 1.0, 0.5)
    obj_low, obj_high = (-0.1, 0.6, 0.02), (0.1, 0.7, 0.02)
    goal_low, goal_high = (0.49, 0.6, 0.00), (0.51, 0.7, 0.02)
    OBJ_RADIUS = 0.02
    grip_add, xz_c = 0.01, 0.005

    def setup(self):
        self.init_config = dict(obj_init_pos=A([0.0, 0.6, 0.02]), hand_init_pos=A([0.0, 0.6, 0.2]))
        self.goal = A([0.5, 0.65, 0.01])
        self.obj_init_pos = self.init_config["obj_init_pos"]
        self.hand_init_pos = self.init_config["hand_init_pos"]

    def random_reset_space(self):
        return A(self.obj_low), A(self.obj_high)

    def _get_pos_objects(self):
        return self.data.body("obj").xpos

    def _get_quat_objects(self):
        return self.data.body("obj").xquat

    def reset_model(self):
        self._reset_hand()
        self._target_pos = self.goal.copy()
        self.obj_init_pos = self.init_config["obj_init_pos"]
        obj_pos = self._get_state_rand_vec()
        self.obj_init_pos = np.concatenate([obj_pos[:2], [self.obj_init_pos[-1]]])
        self._target_pos[1] = obj_pos.copy()[1]
        self._set_obj_xyz(self.obj_init_pos)
        self.model.site("goal").pos = self._target_pos
        return self._get_obs()

    def evaluate_state(self, obs, action):
        reward, tcp_to_obj, tcp_opened, target_to_obj, object_grasped, in_place = self.compute_reward(action, obs)
        grasp_success = float(self.touching_main_object and (tcp_opened > 0))
        return reward, dict(success=float(target_to_obj <= 0.05), near_object=float(tcp_to_obj <= 0.03), grasp_reward=object_grasped,
                            grasp_success=grasp_success, in_place_reward=in_place, obj_to_target=target_to_obj, unscaled_reward=reward)

    def target_for(self, obj):
        return self._target_pos

    def compute_reward(self, action, obs):
        tcp, obj, tcp_opened = self.tcp_center, obs[4:7], obs[3]
        target = self.target_for(obj)
        o2t = float(norm(obj - target))
        tcp_to_obj = float(norm(obj - tcp))
        in_place = tolerance(o2t, bounds=(0, 0.05), margin=norm(self.obj_init_pos - target), sigmoid="long_tail")
        g = _grip_caging(self, action, obj, self.OBJ_RADIUS, self.grip_add, self.xz_c)
        reward = (2 * g) + (6 * hamacher_product(g, in_place))
        if o2t < 0.05:
            reward = 10.0
        return reward, tcp_to_obj, tcp_opened, o2t, g, in_place


class SweepInto(_FreeObjMixin, Sweep):
    """metaworld/envs/sawyer_sweep_into_goal_v3.py"""
    xml = "sawyer_table_with_hole"
    hand_low, hand_high = (-0.5, 0.40, 0.05), (0.5, 1, 0.5)
    obj_low, obj_high = (-0.1, 0.6, 0.02), (0.1, 0.7, 0.02)
    goal_low, goal_high = (-0.001, 0.8399, 0.0199), (0.001, 0.8401, 0.0201)
    grip_add, xz_c = 0.005, 0.01

    def setup(self):
        self.goal = A([0.0, 0.84, 0.02])
        self.obj_init_pos = [0.0, 0.6, 0.02]
        self.hand_init_pos = A([0.0, 0.6, 0.2])

    def random_reset_space(self):
        return np.hstack((self.obj_low, self.goal_low)), np.hstack((self.obj_high, self.goal_high))

    def reset_model(self):
        self._reset_hand()
        self._target_pos = self.goal.copy()
        self.obj_init_pos = self.get_body_com("obj")
        goal_pos = self._get_state_rand_vec()
        while norm(goal_pos[:2] - self._target_pos[:2]) < 0.15:
            goal_pos = self._get_state_rand_vec()
        self.obj_init_pos = np.concatenate([goal_pos[:2], [self.obj_init_pos[-1]]])
        self._set_obj_xyz(self.obj_init_pos)
        self.model.site("goal").pos = self._target_pos
        return self._get_obs()

    def target_for(self, obj):
        return A([self._target_pos[0], self._target_pos[1], obj[2]])


class HandInsert(SawyerXYZEnv):
    """metaworld/envs/sawyer_hand_insert_v3.py"""
    xml = "sawyer_table_with_hole"
    hand_low, hand_high = (-0.5, 0.40, -0.15), (0.5, 1, 0.5)
    obj_low, obj_high = (-0.1, 0.6, 0.05), (0.1, 0.7, 0.05)
    goal_low, goal_high = (-0.04, 0.8, -0.0201), (0.04, 0.88, -0.0199)

    def setup(self):
        self.obj_init_pos = A([0, 0.6, 0.05])
        self.hand_init_pos = A([0, 0.6, 0.2], dtype=np.float32)

    def random_reset_space(self):
        return np.hstack((self.obj_low, self.goal_low)), np.hstack((self.obj_high, self.goal_high))

    def _get_pos_objects(self):
        return self.get_body_com("obj")

    def _get_quat_objects(self):
        return self.data.body("obj").xquat

    def reset_model(self):
        self._reset_hand()
        goal_pos = self._get_state_rand_vec()
        while norm(goal_pos[:2] - goal_pos[-3:-1]) < 0.15:
            goal_pos = self._get_state_rand_vec()
        self.obj_init_pos = np.concatenate([goal_pos[:2], [self.obj_init_pos[-1]]])
        self._target_pos = goal_pos[-3:]
        self._set_obj_xyz(self.obj_init_pos)
        self.model.site("goal").pos = self._target_pos
        return self._get_obs()

    def evaluate_state(self, obs, action):
        obj = obs[4:7]
        reward, tcp_to_obj, tcp_open, obj_to_target, grasp_reward, in_place_reward = self.compute_reward(action, obs)
        gs = float(self.touching_main_object and (tcp_open > 0) and (obj[2] - 0.02 > self.obj_init_pos[2]))
        return reward, dict(success=float(obj_to_target <= 0.05), near_object=float(tcp_to_obj <= 0.03), grasp_success=gs,
                            grasp_reward=grasp_reward, in_place_reward=in_place_reward, obj_to_target=obj_to_target, unscaled_reward=reward)

    def compute_reward(self, action, obs):
        obj = obs[4:7]
        t2o = float(norm(obj - self._target_pos))
        t2o_init = float(norm(self.obj_init_pos - self._target_pos))
        in_place = tolerance(t2o, bounds=(0, self.TARGET_RADIUS), margin=t2o_init, sigmoid="long_tail")
        g = self._gripper_caging_reward(action, obj, object_reach_radius=0.01, obj_radius=0.015, pad_success_thresh=0.05, xz_thresh=0.005,
                                        high_density=True)
        reward = hamacher_product(g, in_place)
        tcp_opened = obs[3]
        tcp_to_obj = float(norm(obj - self.tcp_center))
        if tcp_to_obj < 0.02 and tcp_opened > 0:
            reward += 1.0 + 7.0 * in_place
        if t2o < self.TARGET_RADIUS:
            reward = 10.0
        return reward, tcp_to_obj, tcp_opened, t2o, g, in_place


class PickOutOfHole(SawyerXYZEnv):
    """metaworld/envs/sawyer_pick_out_of_hole_v3.py"""
    xml = "sawyer_pick_out_of_hole"
    hand_low, hand_high = (-0.5, 0.40, -0.05), (0.5, 1, 0.5)
    obj_low, obj_high = (0, 0.75, 0.02), (0, 0.75, 0.02)
    goal_low, goal_high = (-0.1, 0.5, 0.15), (0.1, 0.6, 0.3)

    def setup(self):
        self.obj_init_pos = None
        self.hand_init_pos = A([0.0, 0.6, 0.2])

    def random_reset_space(self):
        return np.hstack((self.obj_low, self.goal_low)), np.hstack((self.obj_high, self.goal_high))

    @property
    def _target_site_config(self):
        return [("goal", self.obj_init_pos if self.obj_init_pos is not None else self.init_right_pad)]

    def _get_pos_objects(self):
        return self.get_body_com("obj")

    def _get_quat_objects(self):
        return self.data.body("obj").xquat

    def reset_model(self):
        self._reset_hand()
        pos_obj, pos_goal = np.split(self._get_state_rand_vec(), 2)
        self.obj_init_pos = pos_obj
        self._set_obj_xyz(self.obj_init_pos)
        self._target_pos = pos_goal
        self.model.site("goal").pos = self._target_pos
        return self._get_obs()

    def evaluate_state(self, obs, action):
        reward, tcp_to_obj, grasp_success, obj_to_target, grasp_reward, in_place_reward = self.compute_reward(action, obs)
        return reward, dict(success=float(obj_to_target <= 0.07), near_object=float(tcp_to_obj <= 0.03), grasp_success=float(grasp_success),
                            grasp_reward=grasp_reward, in_place_reward=in_place_reward, obj_to_target=obj_to_target, unscaled_reward=reward)

    def compute_reward(self, action, obs):
        obj, gripper = obs[4:7], self.tcp_center
        o2t = float(norm(obj - self._target_pos))
        tcp_to_obj = float(norm(obj - gripper))
        in_place_margin = float(norm(self.obj_init_pos - self._target_pos))
        threshold = 0.03
        radius = float(norm(gripper[:2] - self.obj_init_pos[:2]))
        floor = 0.0 if radius <= threshold else 0.015 * np.log(radius - threshold) + 0.15
        above_floor = 1.0 if gripper[2] >= floor else tolerance(max(floor - gripper[2], 0.0), bounds=(0.0, 0.01), margin=0.02, sigmoid="long_tail")
        g = self._gripper_caging_reward(action, obj, object_reach_radius=0.01, obj_radius=0.015, pad_success_thresh=0.02, xz_thresh=0.03,
                                        desired_gripper_effort=0.1, high_density=True)
        in_place = tolerance(o2t, bounds=(0, 0.02), margin=in_place_margin, sigmoid="long_tail")
        reward = hamacher_product(g, in_place)
        grasp_success = (tcp_to_obj < 0.04) and (obj[2] - 0.02 > self.obj_init_pos[2]) and not (obs[3] < 0.33)
        if grasp_success:
            reward += 1.0 + 5.0 * hamacher_product(in_place, above_floor)
        if o2t < self.TARGET_RADIUS:
            reward = 10.0
        return reward, tcp_to_obj, grasp_success, o2t, g, in_place



class _Button(SawyerXYZEnv):
    """Shared parts of the four button tasks (sawyer_button_press*_v3.py)."""
    hand_low, hand_high = (-0.5, 0.40, 0.05), (0.5, 1, 0.5)
    goal_low, goal_high = hand_low, hand_high
    _target_site_config = []
    axis = 1                       # coordinate along which the button travels
    offset = (0.0, -0.193, 0.0)
    success_thr = 0.02

    def setup(self):
        self.obj_init_pos = A([0.0, 0.9, 0.115], dtype=np.float32)
        self.hand_init_pos = A([0, 0.4, 0.2], dtype=np.float32)

    def random_reset_space(self):
        return A(self.obj_low), A(self.obj_high)

    def _get_pos_objects(self):
        return self.get_body_com("button") + A(self.offset)

    def _get_quat_objects(self):
        return self.data.body("button").xquat

    def _set_obj_xyz(self, pos):
        qpos, qvel = self.data.qpos.flat.copy(), self.data.qvel.flat.copy()
        qpos[9] = pos
        qvel[9] = 0
        self.set_state(qpos, qvel)

    def reset_model(self):
        self._reset_hand()
        self.obj_init_pos = self._get_state_rand_vec()
        self.model.body("box").pos = self.obj_init_pos
        self._set_obj_xyz(A(0))
        self._target_pos = self._get_site_pos("hole")
        self._obj_to_target_init = abs(self._target_pos[self.axis] - self._get_site_pos("buttonStart")[self.axis])
        return self._get_obs()

    def evaluate_state(self, obs, action):
        reward, tcp_to_obj, tcp_open, obj_to_target, near_button, button_pressed = self.compute_reward(action, obs)
        return reward, dict(success=float(obj_to_target <= self.success_thr), near_object=float(tcp_to_obj <= 0.05), grasp_success=float(tcp_open > 0),
                            grasp_reward=near_button, in_place_reward=button_pressed, obj_to_target=obj_to_target, unscaled_reward=reward)


class ButtonPress(_Button):
    """metaworld/envs/sawyer_button_press_v3.py"""
    xml = "sawyer_button_press"
    obj_low, obj_high = (-0.1, 0.85, 0.115), (0.1, 0.9, 0.115)
    near_bound, pressed_margin = 0.05, None

    def compute_reward(self, action, obs):
        obj, tcp = obs[4:7], self.tcp_center
        tcp_to_obj = float(norm(obj - tcp))
        tcp_to_obj_init = float(norm(obj - self.init_tcp))
        o2t = abs(self._target_pos[1] - obj[1])
        tcp_closed = max(obs[3], 0.0)
        near = tolerance(tcp_to_obj, bounds=(0, 0.05), margin=tcp_to_obj_init, sigmoid="long_tail")
        pressed = tolerance(o2t, bounds=(0, 0.005), margin=self.pressed_margin_value(), sigmoid="long_tail")
        reward = 2 * hamacher_product(tcp_closed, near)
        if tcp_to_obj <= 0.05:
            reward += 8 * pressed
        return reward, tcp_to_obj, obs[3], o2t, near, pressed

    def pressed_margin_value(self):
        return self._obj_to_target_init


class ButtonPressWall(_Button):
    """metaworld/envs/sawyer_button_press_wall_v3.py"""
    xml = "sawyer_button_press_wall"
    obj_low, obj_high = (-0.05, 0.85, 0.1149), (0.05, 0.9, 0.1151)
    success_thr = 0.03

    def compute_reward(self, action, obs):
        obj, tcp = obs[4:7], self.tcp_center
        tcp_to_obj = float(norm(obj - tcp))
        tcp_to_obj_init = float(norm(obj - self.init_tcp))
        o2t = abs(self._target_pos[1] - obj[1])
        near = tolerance(tcp_to_obj, bounds=(0, 0.01), margin=tcp_to_obj_init, sigmoid="long_tail")
        pressed = tolerance(o2t, bounds=(0, 0.005), margin=self._obj_to_target_init, sigmoid="long_tail")
        if tcp_to_obj > 0.07:
            reward = 2 * hamacher_product((1 - obs[3]) / 2.0, near)
        else:
            reward = 2 + 2 * (1 + obs[3]) + 4 * pressed ** 2
        return reward, tcp_to_obj, obs[3], o2t, near, pressed


class ButtonPressTopdownWall(ButtonPressTopdown):
    """metaworld/envs/sawyer_button_press_topdown_wall_v3.py"""
    xml = "sawyer_button_press_topdown_wall"

    def compute_reward(self, action, obs):
        obj, tcp = obs[4:7], self.tcp_center
        tcp_to_obj = float(norm(obj - tcp))
        tcp_to_obj_init = float(norm(obj - self.init_tcp))
        o2t = abs(self._target_pos[2] - obj[2])
        tcp_closed = max(obs[3], 0.0)
        near = tolerance(tcp_to_obj, bounds=(0, 0.01), margin=tcp_to_obj_init, sigmoid="long_tail")
        pressed = tolerance(o2t, bounds=(0, 0.005), margin=self._obj_to_target_init, sigmoid="long_tail")
        reward = 5 * hamacher_product(tcp_closed, near)
        if tcp_to_obj <= 0.03:
            reward += 5 * pressed
        return reward, tcp_to_obj, obs[3], o2t, near, pressed


class _Coffee(SawyerXYZEnv):
    xml = "sawyer_coffee"
    hand_low, hand_high = (-0.5, 0.4, 0.05), (0.5, 1.0, 0.5)

    def _set_obj_xyz(self, pos):
        """mug free joint comes first in this model; the reference still zeroes qvel[9:15] (sawyer_coffee_pull_v3.py:110-115)"""
        qpos, qvel = self.data.qpos.flatten(), self.data.qvel.flatten()
        qpos[0:3] = pos.copy()
        qvel[9:15] = 0
        self.set_state(qpos, qvel)


class CoffeeButton(_Coffee):
    """metaworld/envs/sawyer_coffee_button_v3.py"""
    obj_low, obj_high = (-0.1, 0.8, -0.001), (0.1, 0.9, 0.001)
    goal_low, goal_high = (-0.101, 0.61, 0.298), (0.101, 0.71, 0.302)
    max_dist = 0.03

    def setup(self):
        self.obj_init_pos = A([0, 0.9, 0.28])
        self.hand_init_pos = A([0.0, 0.4, 0.2])

    def random_reset_space(self):
        return A(self.obj_low), A(self.obj_high)

    @property
    def _target_site_config(self):
        return [("coffee_goal", self._target_pos)]

    def _get_pos_objects(self):
        return self._get_site_pos("buttonStart")

    def _get_quat_objects(self):
        return A([1.0, 0.0, 0.0, 0.0])

    def reset_model(self):
        self._reset_hand()
        self.obj_init_pos = self._get_state_rand_vec()
        self.model.body("coffee_machine").pos = self.obj_init_pos
        self._set_obj_xyz(self.obj_init_pos + A([0.0, -0.22, 0.0]))
        pos_button = self.obj_init_pos + A([0.0, -0.22, 0.3])
        self._target_pos = pos_button + A([0.0, self.max_dist, 0.0])
        return self._get_obs()

    evaluate_state = _Button.evaluate_state
    success_thr = 0.02

    def compute_reward(self, action, obs):
        obj, tcp = obs[4:7], self.tcp_center
        tcp_to_obj = float(norm(obj - tcp))
        tcp_to_obj_init = float(norm(obj - self.init_tcp))
        o2t = abs(self._target_pos[1] - obj[1])
        tcp_closed = max(obs[3], 0.0)
        near = tolerance(tcp_to_obj, bounds=(0, 0.05), margin=tcp_to_obj_init, sigmoid="long_tail")
        pressed = tolerance(o2t, bounds=(0, 0.005), margin=self.max_dist, sigmoid="long_tail")
        reward = 2 * hamacher_product(tcp_closed, near)
        if tcp_to_obj <= 0.05:
            reward += 8 * pressed
        return reward, tcp_to_obj, obs[3], o2t, near, pressed


class CoffeePull(_Coffee):
    """metaworld/envs/sawyer_coffee_pull_v3.py"""
    obj_low, obj_high = (-0.05, 0.7, -0.001), (0.05, 0.75, 0.001)
    goal_low, goal_high = (-0.1, 0.55, -0.001), (0.1, 0.65, 0.001)
    machine_from_goal = False

    def setup(self):
        self.obj_init_pos = A([0, 0.75, 0.0])
        self.hand_init_pos = A([0.0, 0.4, 0.2])

    def random_reset_space(self):
        return np.hstack((self.obj_low, self.goal_low)), np.hstack((self.obj_high, self.goal_high))

    @property
    def _target_site_config(self):
        return [("mug_goal" if not self.machine_from_goal else "coffee_goal", self._target_pos)]

    def _get_id_main_object(self):
        return self.data.geom("mug").id

    def _get_pos_objects(self):
        return self.get_body_com("obj")

    def _get_quat_objects(self):
        return mat2quat_xyzw(self.data.geom("mug").xmat)

    def reset_model(self):
        self._reset_hand()
        pos_mug_init, pos_mug_goal = np.split(self._get_state_rand_vec(), 2)
        self._set_obj_xyz(pos_mug_init)
        self.obj_init_pos = pos_mug_init
        self.model.body("coffee_machine").pos = (pos_mug_goal if self.machine_from_goal else pos_mug_init) + A([0.0, 0.22, 0.0])
        self._target_pos = pos_mug_goal
        self.model.site("mug_goal").pos = self._target_pos
        return self._get_obs()

    def evaluate_state(self, obs, action):
        reward, tcp_to_obj, tcp_open, obj_to_target, grasp_reward, in_place = self.compute_reward(action, obs)
        gs = float(self.touching_main_object and (tcp_open > 0))
        return reward, dict(success=float(obj_to_target <= 0.07), near_object=float(tcp_to_obj <= 0.03), grasp_success=gs,
                            grasp_reward=grasp_reward, in_place_reward=in_place, obj_to_target=obj_to_target, unscaled_reward=reward)

    def compute_reward(self, action, obs):
        obj, target = obs[4:7], self._target_pos.copy()
        scale = A([2.0, 2.0, 1.0])
        t2o = norm((obj - target) * scale)
        t2o_init = norm((self.obj_init_pos - target) * scale)
        in_place = tolerance(t2o, bounds=(0, 0.05), margin=t2o_init, sigmoid="long_tail")
        tcp_opened = obs[3]
        tcp_to_obj = float(norm(obj - self.tcp_center))
        g = self._gripper_caging_reward(action, obj, object_reach_radius=0.04, obj_radius=0.02, pad_success_thresh=0.05, xz_thresh=0.05,
                                        desired_gripper_effort=0.7, medium_density=True)
        reward = hamacher_product(g, in_place)
        if tcp_to_obj < 0.04 and tcp_opened > 0:
            reward += 1.0 + 5.0 * in_place
        if t2o < 0.05:
            reward = 10.0
        return reward, tcp_to_obj, tcp_opened, float(norm(obj - target)), g, in_place


class CoffeePush(CoffeePull):
    """metaworld/envs/sawyer_coffee_push_v3.py"""
    obj_low, obj_high = (-0.1, 0.55, -0.001), (0.1, 0.65, 0.001)
    goal_low, goal_high = (-0.05, 0.7, -0.001), (0.05, 0.75, 0.001)
    machine_from_goal = True

    def setup(self):
        self.obj_init_pos = A([0.0, 0.6, 0.0])
        self.hand_init_pos = A([0.0, 0.4, 0.2])


class DialTurn(SawyerXYZEnv):
    """metaworld/envs/sawyer_dial_turn_v3.py"""
    xml = "sawyer_dial"
    hand_low, hand_high = (-0.5, 0.40, 0.05), (0.5, 1, 0.5)
    obj_low, obj_high = (-0.1, 0.7, 0.0), (0.1, 0.8, 0.0)
    goal_low, goal_high = (-0.1, 0.73, 0.0299), (0.1, 0.83, 0.0301)
    TARGET_RADIUS = 0.07

    def setup(self):
        self.obj_init_pos = A([0, 0.7, 0.0])
        self.hand_init_pos = A([0, 0.6, 0.2], dtype=np.float32)

    def random_reset_space(self):
        return A(self.obj_low), A(self.obj_high)

    def _get_pos_objects(self):
        c = self.get_body_com("dial").copy()
        th = self.data.joint("knob_Joint_1").qpos
        return c + 0.05 * A([np.sin(th).item(), -np.cos(th).item(), 0.0])

    def _get_quat_objects(self):
        return self.data.body("dial").xquat

    def reset_model(self):
        self._reset_hand()
        goal_pos = self._get_state_rand_vec()
        self.obj_init_pos = goal_pos[:3]
        self._target_pos = goal_pos.copy() + A([0, 0.03, 0.03])
        self.model.body("dial").pos = self.obj_init_pos
        self.dial_push_position = self._get_pos_objects() + A([0.05, 0.02, 0.09])
        self.model.site("goal").pos = self._target_pos
        return self._get_obs()

    def evaluate_state(self, obs, action):
        reward, tcp_to_obj, _, target_to_obj, object_grasped, in_place = self.compute_reward(action, obs)
        return reward, dict(success=float(target_to_obj <= self.TARGET_RADIUS), near_object=float(tcp_to_obj <= 0.01), grasp_success=1.0,
                            grasp_reward=object_grasped, in_place_reward=in_place, obj_to_target=target_to_obj, unscaled_reward=reward)

    def compute_reward(self, action, obs):
        obj = self._get_pos_objects()
        push = self._get_pos_objects() + A([0.05, 0.02, 0.09])
        tcp, target = self.tcp_center, self._target_pos.copy()
        t2o = float(norm(obj - target))
        t2o_init = norm(self.dial_push_position - target)
        in_place = tolerance(t2o, bounds=(0, self.TARGET_RADIUS), margin=abs(t2o_init - self.TARGET_RADIUS), sigmoid="long_tail")
        tcp_to_obj = float(norm(push - tcp))
        tcp_to_obj_init = float(norm(self.dial_push_position - self.init_tcp))
        reach = tolerance(tcp_to_obj, bounds=(0, 0.005), margin=abs(tcp_to_obj_init - 0.005), sigmoid="gaussian")
        reach = hamacher_product(reach, min(max(0, action[-1]), 1))
        reward = 10 * hamacher_product(reach, in_place)
        return reward, tcp_to_obj, 0, t2o, reach, in_place


class DoorClose(DoorOpen):
    """metaworld/envs/sawyer_door_close_v3.py"""
    goal_low, goal_high = (0.2, 0.65, 0.1499), (0.3, 0.75, 0.1501)
    _target_site_config = property(lambda self: [("goal", self._target_pos)])

    def setup(self):
        DoorOpen.setup(self)
        self.hand_init_pos = A([-0.5, 0.6, 0.2], dtype=np.float32)

    def reset_model(self):
        self._reset_hand()
        self.obj_init_pos = self._get_state_rand_vec()
        self._target_pos = self.obj_init_pos.copy() + A([0.2, -0.2, 0.0])
        self.model.body("door").pos = self.obj_init_pos
        self.model.site("goal").pos = self._target_pos
        self._set_obj_xyz(A(-1.5708))
        return self._get_obs()

    def evaluate_state(self, obs, action):
        reward, obj_to_target, in_place = self.compute_reward(action, obs)
        return reward, dict(obj_to_target=obj_to_target, in_place_reward=in_place, success=float(obj_to_target <= 0.08), near_object=0.0,
                            grasp_success=1.0, grasp_reward=1.0, unscaled_reward=reward)

    def compute_reward(self, actions, obs):
        tcp, obj, target = self.tcp_center, obs[4:7], self._target_pos
        tcp_to_target = float(norm(tcp - target))
        o2t = float(norm(obj - target))
        in_place = tolerance(o2t, bounds=(0, 0.05), margin=norm(self.obj_init_pos - target), sigmoid="gaussian")
        hand_margin = float(norm(self.hand_init_pos - obj)) + 0.1
        hand_in_place = tolerance(tcp_to_target, bounds=(0, 0.0125), margin=hand_margin, sigmoid="gaussian")
        reward = 3 * hand_in_place + 6 * in_place
        if o2t < 0.05:
            reward = 10
        return reward, o2t, hand_in_place


class DoorLock(SawyerXYZEnv):
    """metaworld/envs/sawyer_door_lock_v3.py"""
    xml = "sawyer_door_lock"
    hand_low, hand_high = (-0.5, 0.40, -0.15), (0.5, 1, 0.5)
    obj_low, obj_high = (-0.1, 0.8, 0.15), (0.1, 0.85, 0.15)
    goal_low, goal_high = hand_low, hand_high
    _lock_length = 0.1
    handle_site = "lockStartLock"

    def setup(self):
        self.obj_init_pos = A([0, 0.85, 0.15], dtype=np.float32)
        self.hand_init_pos = A([0, 0.6, 0.2], dtype=np.float32)

    def random_reset_space(self):
        return A(self.obj_low), A(self.obj_high)

    @property
    def _target_site_config(self):
        return [("goal_lock", self._target_pos), ("goal_unlock", A([10.0, 10.0, 10.0]))]

    def _get_pos_objects(self):
        return self._get_site_pos(self.handle_site)

    def _get_quat_objects(self):
        return self.data.body("door_link").xquat

    def reset_model(self):
        from oracle import mjphys as P
        self._reset_hand()
        self.model.body("door").pos = self._get_state_rand_vec()
        for _ in range(self.frame_skip):
            P.mj_step(self.model, self.data)
        self.obj_init_pos = self.data.body("lock_link").xpos      # live view, as in the reference
        self._target_pos = self.obj_init_pos + A([0.0, -0.04, -0.1])
        return self._get_obs()

    evaluate_state = _Button.evaluate_state
    success_thr = 0.02

    def compute_reward(self, action, obs):
        obj = obs[4:7]
        tcp = self.get_body_com("leftpad")
        scale = A([0.25, 1.0, 0.5])
        tcp_to_obj = float(norm((obj - tcp) * scale))
        tcp_to_obj_init = float(norm((obj - self.init_left_pad) * scale))
        o2t = abs(self._target_pos[2] - obj[2])
        tcp_opened = max(obs[3], 0.0)
        near = tolerance(tcp_to_obj, bounds=(0, 0.01), margin=tcp_to_obj_init, sigmoid="long_tail")
        pressed = tolerance(o2t, bounds=(0, 0.005), margin=self._lock_length, sigmoid="long_tail")
        reward = 2 * hamacher_product(tcp_opened, near) + 8 * pressed
        return reward, tcp_to_obj, obs[3], o2t, near, pressed


class DoorUnlock(DoorLock):
    """metaworld/envs/sawyer_door_unlock_v3.py"""
    goal_low, goal_high = (0.0, 0.64, 0.2100), (0.2, 0.7, 0.2111)
    handle_site = "lockStartUnlock"

    def setup(self):
        self.obj_init_pos = A([0, 0.85, 0.15])
        self.hand_init_pos = A([0, 0.6, 0.2], dtype=np.float32)

    @property
    def _target_site_config(self):
        return [("goal_unlock", self._target_pos), ("goal_lock", A([10.0, 10.0, 10.0]))]

    def _set_obj_xyz(self, pos):
        qpos, qvel = self.data.qpos.flat.copy(), self.data.qvel.flat.copy()
        qpos[9] = pos
        qvel[9] = 0
        self.set_state(qpos, qvel)

    def reset_model(self):
        self._reset_hand()
        self.model.body("door").pos = self._get_state_rand_vec()
        self._set_obj_xyz(A(1.5708))
        self.obj_init_pos = self.data.body("lock_link").xpos      # live view
        self._target_pos = self.obj_init_pos + A([0.1, -0.04, 0.0])
        return self._get_obs()

    def compute_reward(self, action, obs):
        gripper, lock = obs[:3], obs[4:7]
        offset, scale = A([0.0, 0.055, 0.07]), A([0.25, 1.0, 0.5])
        s2l = (gripper + offset - lock) * scale
        s2l_init = (self.init_tcp + offset - self.obj_init_pos) * scale
        ready = tolerance(float(norm(s2l)), bounds=(0, 0.02), margin=norm(s2l_init), sigmoid="long_tail")
        o2t = abs(float(self._target_pos[0] - lock[0]))
        pushed = tolerance(o2t, bounds=(0, 0.005), margin=self._lock_length, sigmoid="long_tail")
        reward = 2 * ready + 8 * pushed
        return reward, float(norm(s2l)), obs[3], o2t, ready, pushed



def _grab_info(reward, grab, ready, in_place, success):
    return reward, dict(success=float(success), near_object=ready, grasp_success=grab >= 0.5, grasp_reward=grab, in_place_reward=in_place,
                        obj_to_target=0, unscaled_reward=reward)


class Assembly(SawyerXYZEnv):
    """metaworld/envs/sawyer_assembly_peg_v3.py"""
    xml = "sawyer_assembly_peg"
    hand_low, hand_high = (-0.5, 0.40, 0.05), (0.5, 1, 0.5)
    obj_low, obj_high = (0, 0.6, 0.02), (0, 0.6, 0.02)
    rgoal_low, rgoal_high = (-0.1, 0.75, 0.1), (0.1, 0.85, 0.1)
    goal_low, goal_high = rgoal_low, rgoal_high
    WRENCH_HANDLE_LENGTH = 0.02

    def setup(self):
        self.obj_init_pos = A([0, 0.6, 0.02], dtype=np.float32)
        self.hand_init_pos = A((0, 0.6, 0.2), dtype=np.float32)

    def random_reset_space(self):
        return np.hstack((self.obj_low, self.rgoal_low)), np.hstack((self.obj_high, self.rgoal_high))

    @property
    def _target_site_config(self):
        return [("pegTop", self._target_pos)]

    def _get_id_main_object(self):
        return self.data.geom("WrenchHandle").id

    def _get_pos_objects(self):
        return self.data.site("RoundNut-8").xpos

    def _get_quat_objects(self):
        return self.data.body("RoundNut").xquat

    def reset_model(self):
        self._reset_hand()
        goal_pos = self._get_state_rand_vec()
        self.obj_init_pos = goal_pos[:3]
        self._target_pos = goal_pos[-3:]
        peg_pos = self._target_pos - A([0.0, 0.0, 0.05])
        self._set_obj_xyz(self.obj_init_pos)
        self.model.body("peg").pos = peg_pos
        self.model.site("pegTop").pos = self._target_pos
        return self._get_obs()

    def evaluate_state(self, obs, action):
        reward, grab, ready, in_place, success = self.compute_reward(action, obs)
        return _grab_info(reward, grab, ready, in_place, success)

    @staticmethod
    def _reward_quat(obs, div=0.4, ideal=(0.707, 0, 0, 0.707)):
        return max(1.0 - float(norm(obs[7:11] - A(ideal))) / div, 0.0)

    @staticmethod
    def _reward_pos(wrench_center, target_pos):
        pos_error = target_pos - wrench_center
        radius = norm(pos_error[:2])
        success = bool(radius < 0.02 and pos_error[2] > 0.0)
        threshold = 0.02 if success else 0.01
        target_height = 0.02 * np.log(radius - threshold) + 0.2 if radius > threshold else 0.0
        pos_error[2] = target_height - wrench_center[2]
        lifted = wrench_center[2] > 0.02 or radius < threshold
        in_place = 0.1 * float(lifted) + 0.9 * tolerance(float(norm(pos_error * A([1.0, 1.0, 3.0]))), bounds=(0, 0.02), margin=0.4, sigmoid="long_tail")
        return in_place, success

    density = dict(medium_density=True)

    def _grab(self, actions, obs):
        hand, wrench = obs[:3], obs[4:7]
        threshed = wrench.copy()
        if abs(wrench[0] - hand[0]) < self.WRENCH_HANDLE_LENGTH / 2.0:
            threshed[0] = hand[0]
        return self._gripper_caging_reward(actions, threshed, object_reach_radius=0.01, obj_radius=0.015, pad_success_thresh=0.02,
                                           xz_thresh=0.01, **self.density)

    def compute_reward(self, actions, obs):
        wrench_center = self._get_site_pos("RoundNut")
        rq = self._reward_quat(obs)
        grab = self._grab(actions, obs)
        in_place, success = self._reward_pos(wrench_center, self._target_pos)
        reward = (2.0 * grab + 6.0 * in_place) * rq
        if success:
            reward = 10.0
        return reward, grab, rq, in_place, success


class Disassemble(Assembly):
    """metaworld/envs/sawyer_disassemble_peg_v3.py"""
    obj_low, obj_high = (0.0, 0.6, 0.025), (0.1, 0.75, 0.02501)
    rgoal_low, rgoal_high = (-0.1, 0.6, 0.1699), (0.1, 0.75, 0.1701)
    goal_low, goal_high = (-0.1, 0.6, 0.1749), (0.1, 0.75, 0.1751)
    density = dict(high_density=True)

    def setup(self):
        self.obj_init_pos = A([0, 0.7, 0.025])
        self.hand_init_pos = A((0, 0.4, 0.2), dtype=np.float32)

    def _get_pos_objects(self):
        return self._get_site_pos("RoundNut-8")

    def reset_model(self):
        from oracle import mjphys as P
        self._reset_hand()
        goal_pos = self._get_state_rand_vec()
        self.obj_init_pos = goal_pos[:3]
        self._target_pos = goal_pos[:3] + A([0, 0, 0.15])
        self.model.body("peg").pos = self.obj_init_pos + A([0.0, 0.0, 0.03])
        self.model.site("pegTop").pos = self.obj_init_pos + A([0.0, 0.0, 0.08])
        P.mj_forward(self.model, self.data)
        self._set_obj_xyz(self.obj_init_pos)
        return self._get_obs()

    def compute_reward(self, actions, obs):
        wrench_center = self._get_site_pos("RoundNut")
        rq = self._reward_quat(obs)
        grab = self._grab(actions, obs)
        pos_error = self._target_pos + A([0.0, 0.0, 0.1]) - wrench_center
        in_place = 0.1 * float(wrench_center[2] > 0.02) + 0.9 * tolerance(float(norm(pos_error)), bounds=(0, 0.02), margin=0.2, sigmoid="long_tail")
        reward = (2.0 * grab + 6.0 * in_place) * rq
        success = obs[6] > self._target_pos[2]
        if success:
            reward = 10.0
        return reward, grab, rq, in_place, success


class Basketball(SawyerXYZEnv):
    """metaworld/envs/sawyer_basketball_v3.py"""
    xml = "sawyer_basketball"
    hand_low, hand_high = (-0.5, 0.40, 0.05), (0.5, 1, 0.5)
    obj_low, obj_high = (-0.1, 0.6, 0.0299), (0.1, 0.7, 0.0301)
    rgoal_low, rgoal_high = (-0.1, 0.85, 0.0), (0.1, 0.9 + 1e-7, 0.0)
    goal_low, goal_high = (-0.1, 0.767, 0.2499), (0.1, 0.817 + 1e-7, 0.2501)
    TARGET_RADIUS = 0.08

    def setup(self):
        self.obj_init_pos = A([0, 0.6, 0.03], dtype=np.float32)
        self.hand_init_pos = A((0, 0.6, 0.2), dtype=np.float32)

    def random_reset_space(self):
        return np.hstack((self.obj_low, self.rgoal_low)), np.hstack((self.obj_high, self.rgoal_high))

    def _get_pos_objects(self):
        return self.get_body_com("bsktball")

    def _get_quat_objects(self):
        return self.data.body("bsktball").xquat

    def reset_model(self):
        self._reset_hand()
        goal_pos = self._get_state_rand_vec()
        basket_pos = goal_pos[3:]
        self.obj_init_pos = np.concatenate([goal_pos[:2], [self.obj_init_pos[-1]]])
        self.model.body("basket_goal").pos = basket_pos
        self._target_pos = self.data.site("goal").xpos          # live view
        self._set_obj_xyz(self.obj_init_pos)
        self.model.site("goal").pos = self._target_pos
        return self._get_obs()

    def evaluate_state(self, obs, action):
        obj = obs[4:7]
        reward, tcp_to_obj, tcp_open, obj_to_target, grasp_reward, in_place_reward = self.compute_reward(action, obs)
        return reward, dict(success=float(obj_to_target <= self.TARGET_RADIUS), near_object=float(tcp_to_obj <= 0.05),
                            grasp_success=float((tcp_open > 0) and (obj[2] - 0.03 > self.obj_init_pos[2])), grasp_reward=grasp_reward,
                            in_place_reward=in_place_reward, obj_to_target=obj_to_target, unscaled_reward=reward)

    def compute_reward(self, action, obs):
        obj = obs[4:7]
        target = self._target_pos.copy()
        target[2] = 0.3
        scale = A([1.0, 1.0, 2.0])
        t2o = float(norm((obj - target) * scale))
        t2o_init = norm((self.obj_init_pos - target) * scale)
        in_place = tolerance(t2o, bounds=(0, self.TARGET_RADIUS), margin=t2o_init, sigmoid="long_tail")
        tcp_opened = float(obs[3])
        tcp_to_obj = float(norm(obj - self.tcp_center))
        g = self._gripper_caging_reward(action, obj, object_reach_radius=0.01, obj_radius=0.025, pad_success_thresh=0.06, xz_thresh=0.005,
                                        high_density=True)
        lifted = tcp_to_obj < 0.035 and tcp_opened > 0 and obj[2] - 0.01 > self.obj_init_pos[2]
        if lifted:
            g = 1.0
        reward = hamacher_product(g, in_place)
        if lifted:
            reward += 1.0 + 5.0 * in_place
        if t2o < self.TARGET_RADIUS:
            reward = 10.0
        return reward, tcp_to_obj, tcp_opened, t2o, g, in_place


class BinPicking(SawyerXYZEnv):
    """metaworld/envs/sawyer_bin_picking_v3.py"""
    xml = "sawyer_bin_picking"
    hand_low, hand_high = (-0.5, 0.40, 0.07), (0.5, 1, 0.5)
    obj_low, obj_high = (-0.21, 0.65, 0.02), (-0.03, 0.75, 0.02)
    goal_low, goal_high = (0.1199, 0.699, -0.001), (0.1201, 0.701, 0.001)
    _target_site_config = []

    def setup(self):
        self.obj_init_pos = A([-0.12, 0.7, 0.02])
        self.hand_init_pos = A((0, 0.6, 0.2))
        self.goal = A([0.12, 0.7, 0.02])
        self._target_to_obj_init = None

    def random_reset_space(self):
        return np.hstack((self.obj_low, self.goal_low)), np.hstack((self.obj_high, self.goal_high))

    def _get_pos_objects(self):
        return self.get_body_com("obj")

    def _get_quat_objects(self):
        return self.data.body("obj").xquat

    def reset_model(self):
        self._reset_hand()
        obj_height = self.get_body_com("obj")[2]
        self.obj_init_pos = np.concatenate([self._get_state_rand_vec()[:2], [obj_height]])
        self._set_obj_xyz(self.obj_init_pos)
        self._target_pos = self.get_body_com("bin_goal")
        self._target_to_obj_init = None
        return self._get_obs()

    def evaluate_state(self, obs, action):
        reward, near_object, grasp_success, obj_to_target, grasp_reward, in_place_reward = self.compute_reward(action, obs)
        return reward, dict(success=float(obj_to_target <= 0.05), near_object=float(near_object), grasp_success=float(grasp_success),
                            grasp_reward=grasp_reward, in_place_reward=in_place_reward, obj_to_target=obj_to_target, unscaled_reward=reward)

    def compute_reward(self, action, obs):
        hand, obj = obs[:3], obs[4:7]
        t2o = float(norm(obj - self._target_pos))
        if self._target_to_obj_init is None:
            self._target_to_obj_init = t2o
        in_place = tolerance(t2o, bounds=(0, self.TARGET_RADIUS), margin=self._target_to_obj_init, sigmoid="long_tail")
        threshold = 0.03
        radii = [norm(hand[:2] - self.obj_init_pos[:2]), norm(hand[:2] - self._target_pos[:2])]
        floor = min([0.02 * np.log(r - threshold) + 0.2 if r > threshold else 0.0 for r in radii])
        above_floor = 1.0 if hand[2] >= floor else tolerance(max(floor - hand[2], 0.0), bounds=(0.0, 0.01), margin=0.05, sigmoid="long_tail")
        g = self._gripper_caging_reward(action, obj, obj_radius=0.015, pad_success_thresh=0.05, object_reach_radius=0.01, xz_thresh=0.01,
                                        desired_gripper_effort=0.7, high_density=True)
        reward = hamacher_product(g, in_place)
        near_object = bool(norm(obj - hand) < 0.04)
        grasp_success = near_object and bool(obj[2] - 0.02 > self.obj_init_pos[2]) and not bool(obs[3] < 0.43)
        if grasp_success:
            reward += 1.0 + 5.0 * hamacher_product(above_floor, in_place)
        if t2o < self.TARGET_RADIUS:
            reward = 10.0
        return reward, near_object, grasp_success, t2o, g, in_place


class BoxClose(SawyerXYZEnv):
    """metaworld/envs/sawyer_box_close_v3.py"""
    xml = "sawyer_box"
    hand_low, hand_high = (-0.5, 0.40, 0.05), (0.5, 1, 0.5)
    obj_low, obj_high = (-0.05, 0.5, 0.02), (0.05, 0.55, 0.02)
    goal_low, goal_high = (-0.1, 0.7, 0.133), (0.1, 0.8, 0.133)
    _target_site_config = []

    def setup(self):
        self.obj_init_pos = A([0, 0.55, 0.02], dtype=np.float32)
        self.hand_init_pos = A((0, 0.6, 0.2), dtype=np.float32)

    def random_reset_space(self):
        return np.hstack((self.obj_low, self.goal_low)), np.hstack((self.obj_high, self.goal_high))

    def _get_id_main_object(self):
        return self.data.geom("BoxHandleGeom").id

    def _get_pos_objects(self):
        return self.get_body_com("top_link")

    def _get_quat_objects(self):
        return self.data.body("top_link").xquat

    def reset_model(self):
        from oracle import mjphys as P
        self._reset_hand()
        box_height = self.get_body_com("boxbody")[2]
        goal_pos = self._get_state_rand_vec()
        self.obj_init_pos = np.concatenate([goal_pos[:2], [self.obj_init_pos[-1]]])
        self._target_pos = goal_pos[-3:]
        self.model.body("boxbody").pos = np.concatenate([self._target_pos[:2], [box_height]])
        for _ in range(self.frame_skip):
            P.mj_step(self.model, self.data)
        self._set_obj_xyz(self.obj_init_pos)
        self.model.site("goal").pos = self._target_pos
        return self._get_obs()

    def evaluate_state(self, obs, action):
        reward, grab, ready, lifted, success = self.compute_reward(action, obs)
        return _grab_info(reward, grab, ready, lifted, success)

    def compute_reward(self, actions, obs):
        grab = float(np.clip(((np.clip(actions[3], -1, 1) + 1.0) / 2.0), 0.0, 1.0))
        rq = Assembly._reward_quat(obs, div=0.2)
        hand = obs[:3]
        lid = obs[4:7] + A([0.0, 0.0, 0.02])
        radius = norm(hand[:2] - lid[:2])
        floor = 0.0 if radius <= 0.02 else 0.04 * np.log(radius - 0.02) + 0.4
        above_floor = 1.0 if hand[2] >= floor else tolerance(floor - hand[2], bounds=(0.0, 0.01), margin=floor / 2.0, sigmoid="long_tail")
        in_place = tolerance(float(norm(hand - lid)), bounds=(0, 0.02), margin=0.5, sigmoid="long_tail")
        ready = hamacher_product(above_floor, in_place)
        lifted = 0.2 * float(lid[2] > 0.04) + 0.8 * tolerance(float(norm((self._target_pos - lid) * A([1.0, 1.0, 3.0]))), bounds=(0, 0.05), margin=0.25, sigmoid="long_tail")
        reward = 2.0 * hamacher_product(grab, ready) + 8.0 * lifted
        success = bool(norm(obs[4:7] - self._target_pos) < 0.08)
        if success:
            reward = 10.0
        reward *= rq
        return reward, grab, ready, lifted, success


class FaucetOpen(SawyerXYZEnv):
    """metaworld/envs/sawyer_faucet_open_v3.py"""
    xml = "sawyer_faucet"
    hand_low, hand_high = (-0.5, 0.40, -0.15), (0.5, 1, 0.5)
    obj_low, obj_high = (-0.05, 0.8, 0.0), (0.05, 0.85, 0.0)
    goal_low, goal_high = hand_low, hand_high
    handle_site, sign, obj_offset, do_forward = "handleStartOpen", +1.0, (-0.04, 0.0, 0.03), False

    def setup(self):
        self.obj_init_pos = A([0, 0.8, 0.0])
        self.hand_init_pos = A([0.0, 0.4, 0.2])

    def random_reset_space(self):
        return A(self.obj_low), A(self.obj_high)

    @property
    def _target_site_config(self):
        a, b = ("goal_open", "goal_close") if self.sign > 0 else ("goal_close", "goal_open")
        return [(a, self._target_pos), (b, A([10.0, 10.0, 10.0]))]

    def _get_pos_objects(self):
        return self._get_site_pos(self.handle_site) + A([0.0, 0.0, -0.01])

    def _get_quat_objects(self):
        return self.data.body("faucetBase").xquat

    def reset_model(self):
        from oracle import mjphys as P
        self._reset_hand()
        self.obj_init_pos = self._get_state_rand_vec()
        self.model.body("faucetBase").pos = self.obj_init_pos
        self._target_pos = self.obj_init_pos + A([self.sign * 0.175, 0.0, 0.125])
        if self.do_forward:
            P.mj_forward(self.model, self.data)
        return self._get_obs()

    def evaluate_state(self, obs, action):
        reward, tcp_to_obj, _, target_to_obj, object_grasped, in_place = self.compute_reward(action, obs)
        return reward, dict(success=float(target_to_obj <= 0.07), near_object=float(tcp_to_obj <= 0.01), grasp_success=1.0,
                            grasp_reward=object_grasped, in_place_reward=in_place, obj_to_target=target_to_obj, unscaled_reward=reward)

    def compute_reward(self, action, obs):
        obj = obs[4:7] + A(self.obj_offset)
        tcp, target = self.tcp_center, self._target_pos.copy()
        t2o = norm(obj - target)
        t2o_init = norm(self.obj_init_pos - target)
        in_place = tolerance(t2o, bounds=(0, 0.07), margin=abs(t2o_init - 0.07), sigmoid="long_tail")
        tcp_to_obj = float(norm(obj - tcp))
        tcp_to_obj_init = norm(self.obj_init_pos - self.init_tcp)
        reach = tolerance(tcp_to_obj, bounds=(0, 0.01), margin=abs(tcp_to_obj_init - 0.01), sigmoid="gaussian")
        reward = 2 * (2 * reach + 3 * in_place)
        reward = 10 if t2o <= 0.07 else reward
        return reward, tcp_to_obj, 0, t2o, reach, in_place


class FaucetClose(FaucetOpen):
    """metaworld/envs/sawyer_faucet_close_v3.py"""
    obj_low, obj_high = (-0.1, 0.8, 0.0), (0.1, 0.85, 0.0)
    handle_site, sign, obj_offset, do_forward = "handleStartClose", -1.0, (0.0, 0.0, 0.0), True


class Hammer(SawyerXYZEnv):
    """metaworld/envs/sawyer_hammer_v3.py"""
    xml = "sawyer_hammer"
    hand_low, hand_high = (-0.5, 0.40, 0.05), (0.5, 1, 0.5)
    obj_low, obj_high = (-0.1, 0.4, 0.0), (0.1, 0.5, 0.0)
    goal_low, goal_high = (0.2399, 0.7399, 0.109), (0.2401, 0.7401, 0.111)
    HAMMER_HANDLE_LENGTH = 0.14

    def setup(self):
        self.hammer_init_pos = A([0, 0.5, 0.0])
        self.obj_init_pos = self.hammer_init_pos.copy()
        self.hand_init_pos = A([0, 0.4, 0.2])

    def random_reset_space(self):
        return A(self.obj_low), A(self.obj_high)

    def _get_id_main_object(self):
        return self.data.geom("HammerHandle").id

    def _get_pos_objects(self):
        return np.hstack((self.get_body_com("hammer").copy(), self.get_body_com("nail_link").copy()))

    def _get_quat_objects(self):
        return np.hstack((self.data.body("hammer").xquat, self.data.body("nail_link").xquat))

    def reset_model(self):
        self._reset_hand()
        self.model.body("box").pos = A([0.24, 0.85, 0.0])
        self._target_pos = self._get_site_pos("goal")
        self.hammer_init_pos = self._get_state_rand_vec()
        self.obj_init_pos = self.hammer_init_pos.copy()
        self._set_obj_xyz(self.hammer_init_pos)
        return self._get_obs()

    def evaluate_state(self, obs, action):
        reward, grab, ready, in_place, success = self.compute_reward(action, obs)
        return _grab_info(reward, grab, ready, in_place, success)

    def compute_reward(self, actions, obs):
        hand, hammer = obs[:3], obs[4:7]
        head = hammer + A([0.16, 0.06, 0.0])
        threshed = hammer.copy()
        if abs(hammer[0] - hand[0]) < self.HAMMER_HANDLE_LENGTH / 2.0:
            threshed[0] = hand[0]
        rq = Assembly._reward_quat(obs, div=0.4, ideal=(1.0, 0.0, 0.0, 0.0))
        grab = self._gripper_caging_reward(actions, threshed, object_reach_radius=0.01, obj_radius=0.015, pad_success_thresh=0.02, xz_thresh=0.01,
                                           high_density=True)
        in_place = 0.1 * float(head[2] > 0.02) + 0.9 * tolerance(norm(self._target_pos - head), bounds=(0, 0.02), margin=0.2, sigmoid="long_tail")
        reward = (2.0 * grab + 6.0 * in_place) * rq
        success = bool(self.data.joint("NailSlideJoint").qpos > 0.09)
        if success and reward > 5.0:
            reward = 10.0
        return reward, grab, rq, in_place, success



class _Handle(SawyerXYZEnv):
    """Shared parts of sawyer_handle_press_v3.py / _press_side / _pull / _pull_side."""
    hand_low, hand_high = (-0.5, 0.40, 0.05), (0.5, 1.0, 0.5)
    _target_site_config = []
    handle_site, goal_site, q0 = "handleStart", "goalPress", -0.001

    def setup(self):
        self.obj_init_pos = A([0, 0.9, 0.0])
        self.hand_init_pos = A((0, 0.6, 0.2))

    def random_reset_space(self):
        return A(self.obj_low), A(self.obj_high)

    def _get_pos_objects(self):
        return self._get_site_pos(self.handle_site)

    def _get_quat_objects(self):
        return np.zeros(4)

    def _set_obj_xyz(self, pos):
        qpos, qvel = self.data.qpos.flat.copy(), self.data.qvel.flat.copy()
        qpos[9] = pos
        qvel[9] = 0
        self.set_state(qpos, qvel)


class HandlePress(_Handle):
    """metaworld/envs/sawyer_handle_press_v3.py"""
    xml = "sawyer_handle_press"
    obj_low, obj_high = (-0.1, 0.8, -0.001), (0.1, 0.9, 0.001)
    goal_low, goal_high = (-0.1, 0.55, 0.04), (0.1, 0.70, 0.08)
    TARGET_RADIUS = 0.02

    def reset_model(self):
        self._reset_hand()
        self.obj_init_pos = self._get_state_rand_vec()
        self.model.body("box").pos = self.obj_init_pos
        self._set_obj_xyz(A(-0.001))
        self._target_pos = self._get_site_pos("goalPress")
        self._handle_init_pos = self._get_pos_objects()
        return self._get_obs()

    def evaluate_state(self, obs, action):
        reward, tcp_to_obj, _, target_to_obj, object_grasped, in_place = self.compute_reward(action, obs)
        return reward, dict(success=float(target_to_obj <= self.TARGET_RADIUS), near_object=float(tcp_to_obj <= 0.05), grasp_success=1.0,
                            grasp_reward=object_grasped, in_place_reward=in_place, obj_to_target=target_to_obj, unscaled_reward=reward)

    def compute_reward(self, actions, obs):
        obj, tcp, target = self._get_pos_objects(), self.tcp_center, self._target_pos.copy()
        t2o = abs(obj[2] - target[2])
        t2o_init = abs(self._handle_init_pos[2] - target[2])
        in_place = tolerance(t2o, bounds=(0, self.TARGET_RADIUS), margin=abs(t2o_init - self.TARGET_RADIUS), sigmoid="long_tail")
        tcp_to_obj = float(norm(obj - tcp))
        tcp_to_obj_init = norm(self._handle_init_pos - self.init_tcp)
        reach = tolerance(tcp_to_obj, bounds=(0, 0.02), margin=abs(tcp_to_obj_init - 0.02), sigmoid="long_tail")
        reward = hamacher_product(reach, in_place)
        reward = 1.0 if t2o <= self.TARGET_RADIUS else reward
        reward *= 10
        return reward, tcp_to_obj, 0, t2o, reach, in_place


class HandlePressSide(HandlePress):
    """metaworld/envs/sawyer_handle_press_side_v3.py"""
    xml = "sawyer_handle_press_sideways"
    obj_low, obj_high = (-0.35, 0.65, -0.001), (-0.25, 0.75, 0.001)
    goal_low, goal_high = _Handle.hand_low, _Handle.hand_high

    def setup(self):
        self.obj_init_pos = A([-0.3, 0.7, 0.0])
        self.hand_init_pos = A((0, 0.6, 0.2))


class HandlePull(_Handle):
    """metaworld/envs/sawyer_handle_pull_v3.py"""
    xml = "sawyer_handle_press"
    obj_low, obj_high = (-0.1, 0.8, -0.001), (0.1, 0.9, 0.001)
    goal_low, goal_high = (-0.1, 0.55, 0.04), (0.1, 0.70, 0.18)
    handle_site = "handleRight"
    side = False

    def reset_model(self):
        self._reset_hand()
        self.obj_init_pos = self._get_state_rand_vec()
        self.model.body("box").pos = self.obj_init_pos
        self._set_obj_xyz(A(-0.1))
        self._target_pos = self._get_site_pos("goalPull")
        if self.side:
            self.obj_init_pos = self._get_pos_objects()
        return self._get_obs()

    def evaluate_state(self, obs, action):
        obj = obs[4:7]
        reward, tcp_to_obj, tcp_open, obj_to_target, grasp_reward, in_place_reward = self.compute_reward(action, obs)
        return reward, dict(success=float(obj_to_target <= (0.08 if self.side else self.TARGET_RADIUS)), near_object=float(tcp_to_obj <= 0.05),
                            grasp_success=float((tcp_open > 0) and (obj[2] - 0.03 > self.obj_init_pos[2])), grasp_reward=grasp_reward,
                            in_place_reward=in_place_reward, obj_to_target=obj_to_target, unscaled_reward=reward)

    def compute_reward(self, action, obs):
        obj, target = obs[4:7], self._target_pos.copy()
        if self.side:
            t2o = norm(obj - target)
            t2o_init = norm(self.obj_init_pos - target)
            g = self._gripper_caging_reward(action, obj, pad_success_thresh=0.06, obj_radius=0.032, object_reach_radius=0.01, xz_thresh=0.01, high_density=True)
        else:
            t2o = abs(target[2] - obj[2])
            t2o_init = abs(target[2] - self.obj_init_pos[2])
            g = self._gripper_caging_reward(action, obj, pad_success_thresh=0.05, obj_radius=0.022, object_reach_radius=0.01, xz_thresh=0.01, high_density=True)
        in_place = tolerance(t2o, bounds=(0, self.TARGET_RADIUS), margin=t2o_init, sigmoid="long_tail")
        reward = hamacher_product(g, in_place)
        tcp_opened = obs[3]
        tcp_to_obj = float(norm(obj - self.tcp_center))
        if tcp_to_obj < 0.035 and tcp_opened > 0 and obj[2 if self.side else 1] - 0.01 > self.obj_init_pos[2]:
            reward += 1.0 + 5.0 * in_place
        if t2o < self.TARGET_RADIUS:
            reward = 10.0
        return reward, tcp_to_obj, tcp_opened, t2o, g, in_place


class HandlePullSide(HandlePull):
    """metaworld/envs/sawyer_handle_pull_side_v3.py"""
    xml = "sawyer_handle_press_sideways"
    obj_low, obj_high = (-0.35, 0.65, 0.0), (-0.25, 0.75, 0.0)
    goal_low, goal_high = _Handle.hand_low, _Handle.hand_high
    handle_site = "handleCenter"
    side = True

    def setup(self):
        self.obj_init_pos = A([-0.3, 0.7, 0.0])
        self.hand_init_pos = A((0, 0.6, 0.2))


class LeverPull(SawyerXYZEnv):
    """metaworld/envs/sawyer_lever_pull_v3.py"""
    xml = "sawyer_lever_pull"
    hand_low, hand_high = (-0.5, 0.40, -0.15), (0.5, 1, 0.5)
    obj_low, obj_high = (-0.1, 0.7, 0.0), (0.1, 0.8, 0.0)
    goal_low, goal_high = hand_low, hand_high
    LEVER_RADIUS = 0.2

    def setup(self):
        self.obj_init_pos = A([0, 0.7, 0.0])
        self.hand_init_pos = A([0, 0.4, 0.2], dtype=np.float32)

    def random_reset_space(self):
        return A(self.obj_low), A(self.obj_high)

    def _get_pos_objects(self):
        return self._get_site_pos("leverStart")

    def _get_quat_objects(self):
        return mat2quat_xyzw(self.data.geom("objGeom").xmat)

    def reset_model(self):
        self._reset_hand()
        self.obj_init_pos = self._get_state_rand_vec()
        self.model.body("lever").pos = self.obj_init_pos
        self._lever_pos_init = self.obj_init_pos + A([0.12, -self.LEVER_RADIUS, 0.25])
        self._target_pos = self.obj_init_pos + A([0.12, 0.0, 0.25 + self.LEVER_RADIUS])
        self.model.site("goal").pos = self._target_pos
        return self._get_obs()

    def evaluate_state(self, obs, action):
        reward, s2l, ready, lever_error, engagement = self.compute_reward(action, obs)
        return reward, dict(success=float(lever_error <= np.pi / 24), near_object=float(s2l < 0.03), grasp_success=float(ready > 0.9),
                            grasp_reward=ready, in_place_reward=engagement, obj_to_target=s2l, unscaled_reward=reward)

    def compute_reward(self, action, obs):
        gripper, lever = obs[:3], obs[4:7]
        scale, offset = A([4.0, 1.0, 4.0]), A([0.0, 0.055, 0.07])
        s2l = (gripper + offset - lever) * scale
        s2l_init = (self.init_tcp + offset - self._lever_pos_init) * scale
        ready = tolerance(float(norm(s2l)), bounds=(0, 0.02), margin=norm(s2l_init), sigmoid="long_tail")
        lever_angle = float(-self.data.joint("LeverAxis").qpos.item())
        lever_error = abs(lever_angle - np.pi / 2.0)
        engagement = tolerance(lever_error, bounds=(0, np.pi / 48.0), margin=(np.pi / 2.0) - (np.pi / 12.0), sigmoid="long_tail")
        target = self._target_pos
        in_place = tolerance(float(norm(lever - target)), bounds=(0, 0.04), margin=float(norm(self._lever_pos_init - target)), sigmoid="long_tail")
        reward = 10.0 * hamacher_product(ready, in_place)
        return reward, float(norm(s2l)), ready, lever_error, engagement


class PegUnplugSide(SawyerXYZEnv):
    """metaworld/envs/sawyer_peg_unplug_side_v3.py"""
    xml = "sawyer_peg_unplug_side"
    hand_low, hand_high = (-0.5, 0.40, 0.05), (0.5, 1, 0.5)
    obj_low, obj_high = (-0.25, 0.6, -0.001), (-0.15, 0.8, 0.001)
    goal_low, goal_high = (-0.056, 0.6, 0.13), (0.044, 0.8, 0.132)

    def setup(self):
        self.obj_init_pos = A([-0.225, 0.6, 0.05])
        self.hand_init_pos = A((0, 0.6, 0.2))

    def random_reset_space(self):
        return A(self.obj_low), A(self.obj_high)

    def _get_pos_objects(self):
        return self._get_site_pos("pegEnd")

    def _get_quat_objects(self):
        return self.data.body("plug1").xquat

    def _set_obj_xyz(self, pos):
        qpos, qvel = self.data.qpos.flat.copy(), self.data.qvel.flat.copy()
        qpos[9:12] = pos
        qpos[12:16] = A([1.0, 0.0, 0.0, 0.0])
        qvel[9:12] = 0
        self.set_state(qpos, qvel)

    def reset_model(self):
        self._reset_hand()
        pos_box = self._get_state_rand_vec()
        self.model.body("box").pos = pos_box
        pos_plug = pos_box + A([0.044, 0.0, 0.131])
        self._set_obj_xyz(pos_plug)
        self.obj_init_pos = self._get_site_pos("pegEnd")
        self._target_pos = pos_plug + A([0.15, 0.0, 0.0])
        self.model.site("goal").pos = self._target_pos
        return self._get_obs()

    def evaluate_state(self, obs, action):
        reward, tcp_to_obj, tcp_open, obj_to_target, grasp_reward, in_place_reward, grasp_success = self.compute_reward(action, obs)
        return reward, dict(success=float(obj_to_target <= 0.07), near_object=float(tcp_to_obj <= 0.03), grasp_success=grasp_success,
                            grasp_reward=grasp_reward, in_place_reward=in_place_reward, obj_to_target=obj_to_target, unscaled_reward=reward)

    def compute_reward(self, action, obs):
        tcp, obj, tcp_opened, target = self.tcp_center, obs[4:7], obs[3], self._target_pos
        tcp_to_obj = float(norm(obj - tcp))
        o2t = float(norm(obj - target))
        g = self._gripper_caging_reward(action, obj, object_reach_radius=0.01, obj_radius=0.025, pad_success_thresh=0.05, xz_thresh=0.005,
                                        desired_gripper_effort=0.8, high_density=True)
        in_place = tolerance(o2t, bounds=(0, 0.05), margin=float(norm(self.obj_init_pos - target)), sigmoid="long_tail")
        grasp_success = tcp_opened > 0.5 and (obj[0] - self.obj_init_pos[0] > 0.015)
        reward = 2 * g
        if grasp_success and tcp_to_obj < 0.035:
            reward = 1 + 2 * g + 5 * in_place
        if o2t <= 0.05:
            reward = 10.0
        return reward, tcp_to_obj, tcp_opened, o2t, g, in_place, float(grasp_success)


class PlateSlide(SawyerXYZEnv):
    """metaworld/envs/sawyer_plate_slide_v3.py (+ _side, _back, _back_side variants below)"""
    xml = "sawyer_plate_slide"
    hand_low, hand_high = (-0.5, 0.40, 0.05), (0.5, 1, 0.5)
    obj_low, obj_high = (0.0, 0.6, 0.0), (0.0, 0.6, 0.0)
    goal_low, goal_high = (-0.1, 0.85, 0.0), (0.1, 0.9, 0.0)
    q_init = (0.0, 0.0)
    variant_b = False
    goal_body = "model_target"       # how puck_goal is moved: model.body pos := target / data xpos (transient) / model.body pos := obj_init

    def setup(self):
        self.obj_init_pos = A([0.0, 0.6, 0.0], dtype=np.float32)
        self.hand_init_pos = A((0, 0.6, 0.2), dtype=np.float32)

    def random_reset_space(self):
        return np.hstack((self.obj_low, self.goal_low)), np.hstack((self.obj_high, self.goal_high))

    def _get_pos_objects(self):
        return self.data.geom("puck").xpos

    def _get_quat_objects(self):
        return mat2quat_xyzw(self.data.geom("puck").xmat)

    def _set_obj_xyz(self, pos):
        qpos, qvel = self.data.qpos.flat.copy(), self.data.qvel.flat.copy()
        qpos[9:11] = pos
        self.set_state(qpos, qvel)

    def reset_model(self):
        self._reset_hand()
        rand_vec = self._get_state_rand_vec()
        self.init_tcp = self.tcp_center
        self.obj_init_pos = rand_vec[:3]
        self._target_pos = rand_vec[3:]
        if self.goal_body == "model_target":
            self.model.body("puck_goal").pos = self._target_pos
        elif self.goal_body == "model_obj":
            self.model.body("puck_goal").pos = self.obj_init_pos
        else:
            self.data.body("puck_goal").xpos = self._target_pos      # transient: overwritten by the next forward
        self._set_obj_xyz(A(self.q_init))
        self.model.site("goal").pos = self._target_pos
        return self._get_obs()

    def evaluate_state(self, obs, action):
        reward, tcp_to_obj, tcp_opened, obj_to_target, object_grasped, in_place = self.compute_reward(action, obs)
        return reward, dict(success=float(obj_to_target <= 0.07), near_object=float(tcp_to_obj <= 0.03), grasp_reward=object_grasped, grasp_success=0.0,
                            in_place_reward=in_place, obj_to_target=obj_to_target, unscaled_reward=reward)

    def compute_reward(self, action, obs):
        R = 0.05
        tcp, obj, tcp_opened, target = self.tcp_center, obs[4:7], obs[3], self._target_pos
        o2t = float(norm(obj - target))
        m1 = float(norm(self.obj_init_pos - target)) - (R if self.variant_b else 0.0)
        in_place = tolerance(o2t, bounds=(0, R), margin=m1, sigmoid="long_tail")
        tcp_to_obj = float(norm(tcp - obj))
        m2 = float(norm(self.init_tcp - self.obj_init_pos)) - (R if self.variant_b else 0.0)
        grasped = tolerance(tcp_to_obj, bounds=(0, R), margin=m2, sigmoid="long_tail")
        if self.variant_b:
            reward = 1.5 * grasped
            if tcp[2] <= 0.03 and tcp_to_obj < 0.07:
                reward = 2.0 + (7.0 * in_place)
        else:
            reward = 8 * hamacher_product(grasped, in_place)
        if o2t < R:
            reward = 10.0
        return reward, tcp_to_obj, tcp_opened, o2t, grasped, in_place


class PlateSlideSide(PlateSlide):
    xml = "sawyer_plate_slide_sideway"
    goal_low, goal_high = (-0.3, 0.54, 0.0), (-0.25, 0.66, 0.0)
    variant_b, goal_body = True, "data"


class PlateSlideBack(PlateSlide):
    obj_low, obj_high = (0.0, 0.85, 0.0), (0.0, 0.85, 0.0)
    goal_low, goal_high = (-0.1, 0.6, 0.015), (0.1, 0.6, 0.015)
    q_init, variant_b, goal_body = (0.0, 0.15), True, "data"


class PlateSlideBackSide(PlateSlide):
    xml = "sawyer_plate_slide_sideway"
    obj_low, obj_high = (-0.25, 0.6, 0.0), (-0.25, 0.6, 0.0)
    goal_low, goal_high = (-0.05, 0.6, 0.015), (0.15, 0.6, 0.015)
    q_init, variant_b, goal_body = (-0.15, 0.0), True, "model_obj"


class ShelfPlace(_FreeObjMixin, SawyerXYZEnv):
    """metaworld/envs/sawyer_shelf_place_v3.py"""
    xml = "sawyer_shelf_placing"
    hand_low, hand_high = (-0.5, 0.40, 0.05), (0.5, 1, 0.5)
    obj_low, obj_high = (-0.1, 0.5, 0.019), (0.1, 0.6, 0.021)
    goal_low, goal_high = (-0.1, 0.8, 0.299), (0.1, 0.9, 0.301)

    def setup(self):
        self.init_config = dict(obj_init_pos=A([0, 0.6, 0.02]))
        self.obj_init_pos = self.init_config["obj_init_pos"]
        self.hand_init_pos = A([0, 0.6, 0.2], dtype=np.float32)

    def random_reset_space(self):
        return np.hstack((self.obj_low, self.goal_low)), np.hstack((self.obj_high, self.goal_high))

    def reset_model(self):
        from oracle import mjphys as P
        self._reset_hand()
        z = self.get_body_com("obj")[-1]
        goal_pos = self._get_state_rand_vec()
        base_shelf_pos = goal_pos - A([0, 0, 0, 0, 0, 0.3])
        self.obj_init_pos = np.concatenate((base_shelf_pos[:2], [z]))
        self.model.body("shelf").pos = base_shelf_pos[-3:]
        P.mj_forward(self.model, self.data)
        self._target_pos = self.model.site("goal").pos + self.model.body("shelf").pos
        self._set_obj_xyz(self.obj_init_pos)
        self._set_pos_site("goal", self._target_pos)
        return self._get_obs()

    evaluate_state = PushWall.evaluate_state

    def compute_reward(self, action, obs):
        R = 0.05
        tcp, obj, tcp_opened, target = self.tcp_center, obs[4:7], obs[3], self._target_pos
        o2t = float(norm(obj - target))
        tcp_to_obj = float(norm(obj - tcp))
        in_place = tolerance(o2t, bounds=(0, R), margin=norm(self.obj_init_pos - target), sigmoid="long_tail")
        g = self._gripper_caging_reward(action=action, obj_pos=obj, obj_radius=0.02, pad_success_thresh=0.05, object_reach_radius=0.01,
                                        xz_thresh=0.01, high_density=False)
        reward = hamacher_product(g, in_place)
        if 0.0 < obj[2] < 0.24 and (target[0] - 0.15 < obj[0] < target[0] + 0.15) and ((target[1] - 3 * R) < obj[1] < target[1]):
            z_scaling = (0.24 - obj[2]) / 0.24
            y_scaling = (obj[1] - (target[1] - 3 * R)) / (3 * R)
            in_place = np.clip(in_place - hamacher_product(y_scaling, z_scaling), 0.0, 1.0)
        if (0.0 < obj[2] < 0.24) and (target[0] - 0.15 < obj[0] < target[0] + 0.15) and (obj[1] > target[1]):
            in_place = 0.0
        if tcp_to_obj < 0.025 and (tcp_opened > 0) and (obj[2] - 0.01 > self.obj_init_pos[2]):
            reward += 1.0 + 5.0 * in_place
        if o2t < R:
            reward = 10.0
        return reward, tcp_to_obj, tcp_opened, o2t, g, in_place


class Soccer(SawyerXYZEnv):
    """metaworld/envs/sawyer_soccer_v3.py"""
    xml = "sawyer_soccer"
    hand_low, hand_high = (-0.5, 0.40, 0.05), (0.5, 1, 0.5)
    obj_low, obj_high = (-0.1, 0.6, 0.03), (0.1, 0.7, 0.03)
    goal_low, goal_high = (-0.1, 0.8, 0.0), (0.1, 0.9, 0.0)
    OBJ_RADIUS, TARGET_RADIUS = 0.013, 0.07

    def setup(self):
        self.obj_init_pos = A([0, 0.6, 0.03])
        self.hand_init_pos = A([0.0, 0.6, 0.2])

    def random_reset_space(self):
        return np.hstack((self.obj_low, self.goal_low)), np.hstack((self.obj_high, self.goal_high))

    def _get_id_main_object(self):
        return self.data.geom("objGeom").id

    def _get_pos_objects(self):
        return self.get_body_com("soccer_ball")

    def _get_quat_objects(self):
        return mat2quat_xyzw(self.data.body("soccer_ball").xmat)

    def reset_model(self):
        self._reset_hand()
        goal_pos = self._get_state_rand_vec()
        self._target_pos = goal_pos[3:]
        self.obj_init_pos = np.concatenate([goal_pos[:2], [self.obj_init_pos[-1]]])
        self.model.body("goal_whole").pos = self._target_pos
        self._set_obj_xyz(self.obj_init_pos)
        self.model.site("goal").pos = self._target_pos
        return self._get_obs()

    def evaluate_state(self, obs, action):
        obj = obs[4:7]
        reward, tcp_to_obj, tcp_opened, target_to_obj, object_grasped, in_place = self.compute_reward(action, obs)
        gs = float(self.touching_main_object and (tcp_opened > 0) and (obj[2] - 0.02 > self.obj_init_pos[2]))
        return reward, dict(success=float(target_to_obj <= 0.07), near_object=float(tcp_to_obj <= 0.03), grasp_success=gs, grasp_reward=object_grasped,
                            in_place_reward=in_place, obj_to_target=target_to_obj, unscaled_reward=reward)

    def compute_reward(self, action, obs):
        obj, tcp_opened = obs[4:7], obs[3]
        xs = A([3.0, 1.0, 1.0])
        tcp_to_obj = float(norm(obj - self.tcp_center))
        t2o = float(norm((obj - self._target_pos) * xs))
        t2o_init = float(norm((obj - self.obj_init_pos) * xs))
        in_place = tolerance(t2o, bounds=(0, self.TARGET_RADIUS), margin=t2o_init, sigmoid="long_tail")
        goal_line = self._target_pos[1] - 0.1
        if obj[1] > goal_line and abs(obj[0] - self._target_pos[0]) > 0.10:
            in_place = np.clip(in_place - 2 * ((obj[1] - goal_line) / (1 - goal_line)), 0.0, 1.0)
        g = _grip_caging(self, action, obj, self.OBJ_RADIUS, 0.01, 0.005)
        reward = (3 * g) + (6.5 * in_place)
        if t2o < self.TARGET_RADIUS:
            reward = 10.0
        return reward, tcp_to_obj, tcp_opened, float(norm(obj - self._target_pos)), g, in_place


class StickPull(SawyerXYZEnv):
    """metaworld/envs/sawyer_stick_pull_v3.py"""
    xml = "sawyer_stick_obj"
    hand_low, hand_high = (-0.5, 0.35, 0.05), (0.5, 1, 0.5)
    obj_low, obj_high = (-0.1, 0.55, 0.000), (0.0, 0.65, 0.001)
    goal_low, goal_high = (0.35, 0.45, 0.0199), (0.45, 0.55, 0.0201)
    obj_q0 = (0.0, 0.09)
    second_offset = (0.0, 0.0, 0.0)

    def setup(self):
        self.stick_init_pos = A([0, 0.6, 0.02])
        self.hand_init_pos = A([0, 0.6, 0.2])
        self.obj_init_pos = A([0.2, 0.69, 0.0])

    def random_reset_space(self):
        return np.hstack((self.obj_low, self.goal_low)), np.hstack((self.obj_high, self.goal_high))

    def _get_id_main_object(self):
        return self.data.geom("objGeom").id

    def _get_pos_objects(self):
        return np.hstack((self.get_body_com("stick").copy(), self._get_site_pos("insertion") + A(self.second_offset)))

    def _get_quat_objects(self):
        return np.hstack((mat2quat_xyzw(self.data.body("stick").xmat), A([0.0, 0.0, 0.0, 0.0])))

    def _set_stick_xyz(self, pos):
        qpos, qvel = self.data.qpos.flat.copy(), self.data.qvel.flat.copy()
        qpos[9:12] = pos.copy()
        qvel[9:15] = 0
        self.set_state(qpos, qvel)

    def _set_obj_xyz(self, pos):
        qpos, qvel = self.data.qpos.flat.copy(), self.data.qvel.flat.copy()
        qpos[16:18] = pos.copy()
        qvel[16:18] = 0          # nv = 17: this slice only reaches dof 16 (reference quirk)
        self.set_state(qpos, qvel)

    def reset_model(self):
        self._reset_hand()
        goal_pos = self._get_state_rand_vec()
        self.stick_init_pos = np.concatenate([goal_pos[:2], [0.02]])
        self._target_pos = np.concatenate([goal_pos[-3:-1], [self.target_z()]])
        self._set_stick_xyz(self.stick_init_pos)
        self._set_obj_xyz(A(self.obj_q0))
        self.obj_init_pos = self.get_body_com("object").copy()
        self.model.site("goal").pos = self._target_pos
        return self._get_obs()

    def target_z(self):
        return 0.02

    @staticmethod
    def _stick_is_inserted(handle, end_of_stick):
        return (end_of_stick[0] >= handle[0]) and (abs(end_of_stick[1] - handle[1]) <= 0.040) and (abs(end_of_stick[2] - handle[2]) <= 0.060)

    def evaluate_state(self, obs, action):
        stick, handle = obs[4:7], obs[11:14]
        end_of_stick = self._get_site_pos("stick_end")
        reward, tcp_to_obj, tcp_open, container_to_target, grasp_reward, stick_in_place = self.compute_reward(action, obs)
        success = float((norm(handle - self._target_pos) <= 0.12) and self._stick_is_inserted(handle, end_of_stick))
        gs = float(self.touching_main_object and (tcp_open > 0) and (stick[2] - 0.02 > self.obj_init_pos[2]))
        return reward, dict(success=success, near_object=float(tcp_to_obj <= 0.03), grasp_success=gs, grasp_reward=grasp_reward,
                            in_place_reward=stick_in_place, obj_to_target=container_to_target, unscaled_reward=reward)

    def compute_reward(self, action, obs):
        R = 0.05
        tcp, stick = self.tcp_center, obs[4:7]
        end_of_stick = self._get_site_pos("stick_end")
        container = obs[11:14] + A([0.05, 0.0, 0.0])
        container_init = self.obj_init_pos + A([0.05, 0.0, 0.0])
        handle, tcp_opened, target = obs[11:14], obs[3], self._target_pos
        tcp_to_stick = float(norm(stick - tcp))
        h2t = float(norm(handle - target))
        yz = A([1.0, 1.0, 2.0])
        sip = tolerance(float(norm((stick - container) * yz)), bounds=(0, R), margin=float(norm((self.stick_init_pos - container_init) * yz)), sigmoid="long_tail")
        sip2 = tolerance(float(norm(stick - target)), bounds=(0, R), margin=float(norm(self.stick_init_pos - target)), sigmoid="long_tail")
        cip = tolerance(float(norm(container - target)), bounds=(0, R), margin=float(norm(self.obj_init_pos - target)), sigmoid="long_tail")
        g = self._gripper_caging_reward(action=action, obj_pos=stick, obj_radius=0.014, pad_success_thresh=0.05, object_reach_radius=0.01,
                                        xz_thresh=0.01, high_density=True)
        grasp_success = tcp_to_stick < 0.02 and (tcp_opened > 0) and (stick[2] - 0.01 > self.stick_init_pos[2])
        g = 1 if grasp_success else g
        ipg = hamacher_product(g, sip)
        reward = ipg
        if grasp_success:
            reward = 1.0 + ipg + 5.0 * sip
            if self._stick_is_inserted(handle, end_of_stick):
                reward = 1.0 + ipg + 5.0 + 2.0 * sip2 + 1.0 * cip
                if h2t <= 0.12:
                    reward = 10.0
        return reward, tcp_to_stick, tcp_opened, h2t, g, sip


class StickPush(StickPull):
    """metaworld/envs/sawyer_stick_push_v3.py"""
    hand_low, hand_high = (-0.5, 0.40, 0.05), (0.5, 1, 0.5)
    obj_low, obj_high = (-0.08, 0.58, 0.000), (-0.03, 0.62, 0.001)
    goal_low, goal_high = (0.399, 0.55, 0.1319), (0.401, 0.6, 0.1321)
    obj_q0 = (0.0, 0.0)
    second_offset = (0.0, 0.09, 0.0)

    def setup(self):
        self.stick_init_pos = A([-0.1, 0.6, 0.02])
        self.hand_init_pos = A([0, 0.6, 0.2])
        self.obj_init_pos = A([0.2, 0.6, 0.0])

    def target_z(self):
        return self._get_site_pos("insertion")[-1]

    def evaluate_state(self, obs, action):
        stick, container = obs[4:7], obs[11:14]
        reward, tcp_to_obj, tcp_open, container_to_target, grasp_reward, stick_in_place = self.compute_reward(action, obs)
        success = float(norm(container - self._target_pos) <= 0.12)
        gs = float(self.touching_main_object and (tcp_open > 0) and (stick[2] - 0.01 > self.stick_init_pos[2]))
        return reward, dict(success=gs and success, near_object=float(tcp_to_obj <= 0.03), grasp_success=gs, grasp_reward=grasp_reward,
                            in_place_reward=stick_in_place, obj_to_target=container_to_target, unscaled_reward=reward)

    def _gripper_caging_reward(self, action, obj_pos, obj_radius, pad_success_thresh, object_reach_radius, xz_thresh, desired_gripper_effort=1.0,
                               high_density=False, medium_density=False):
        saved = self.obj_init_pos
        self.obj_init_pos = self.stick_init_pos            # the override is the shared reward with stick_init_pos in place of obj_init_pos
        try:
            return SawyerXYZEnv._gripper_caging_reward(self, action, obj_pos, obj_radius, pad_success_thresh, object_reach_radius, xz_thresh,
                                                       desired_gripper_effort, high_density, medium_density)
        finally:
            self.obj_init_pos = saved

    def compute_reward(self, action, obs):
        R = 0.12
        tcp = self.tcp_center
        stick = obs[4:7] + A([0.015, 0.0, 0.0])
        container, tcp_opened, target = obs[11:14], obs[3], self._target_pos
        tcp_to_stick = float(norm(stick - tcp))
        sip = tolerance(float(norm(stick - target)), bounds=(0, R), margin=float(norm(self.stick_init_pos - target) - R), sigmoid="long_tail")
        c2t = float(norm(container - target))
        cip = tolerance(c2t, bounds=(0, R), margin=float(norm(self.obj_init_pos - target) - R), sigmoid="long_tail")
        g = self._gripper_caging_reward(action=action, obj_pos=stick, obj_radius=0.04, pad_success_thresh=0.05, object_reach_radius=0.01,
                                        xz_thresh=0.01, high_density=True)
        reward = g
        if tcp_to_stick < 0.02 and (tcp_opened > 0) and (stick[2] - 0.01 > self.stick_init_pos[2]):
            g = 1
            reward = 2.0 + 5.0 * sip + 3.0 * cip
            if c2t <= R:
                reward = 10.0
        return reward, tcp_to_stick, tcp_opened, c2t, g, sip


TASKS = {"reach-v3": Reach, "push-v3": Push, "pick-place-v3": PickPlace, "door-open-v3": DoorOpen,
         "drawer-open-v3": DrawerOpen, "drawer-close-v3": DrawerClose, "button-press-topdown-v3": ButtonPressTopdown,
         "peg-insert-side-v3": PegInsertSide, "window-open-v3": WindowOpen, "window-close-v3": WindowClose,
         "reach-wall-v3": ReachWall, "push-wall-v3": PushWall, "pick-place-wall-v3": PickPlaceWall, "push-back-v3": PushBack,
         "sweep-v3": Sweep, "sweep-into-v3": SweepInto, "hand-insert-v3": HandInsert, "pick-out-of-hole-v3": PickOutOfHole,
         "button-press-v3": ButtonPress, "button-press-wall-v3": ButtonPressWall, "button-press-topdown-wall-v3": ButtonPressTopdownWall,
         "coffee-button-v3": CoffeeButton, "coffee-pull-v3": CoffeePull, "coffee-push-v3": CoffeePush, "dial-turn-v3": DialTurn,
         "door-close-v3": DoorClose, "door-lock-v3": DoorLock, "door-unlock-v3": DoorUnlock,
         "assembly-v3": Assembly, "disassemble-v3": Disassemble, "basketball-v3": Basketball, "bin-picking-v3": BinPicking,
         "box-close-v3": BoxClose, "faucet-open-v3": FaucetOpen, "faucet-close-v3": FaucetClose, "hammer-v3": Hammer,
         "handle-press-v3": HandlePress, "handle-press-side-v3": HandlePressSide, "handle-pull-v3": HandlePull, "handle-pull-side-v3": HandlePullSide,
         "lever-pull-v3": LeverPull, "peg-unplug-side-v3": PegUnplugSide, "plate-slide-v3": PlateSlide, "plate-slide-side-v3": PlateSlideSide,
         "plate-slide-back-v3": PlateSlideBack, "plate-slide-back-side-v3": PlateSlideBackSide, "shelf-place-v3": ShelfPlace, "soccer-v3": Soccer,
         "stick-push-v3": StickPush, "stick-pull-v3": StickPull}
