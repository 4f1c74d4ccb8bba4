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


TASKS = {"reach-v3": Reach, "push-v3": Push, "pick-place-v3": PickPlace, "door-open-v3": DoorOpen,
         "drawer-open-v3": DrawerOpen, "drawer-close-v3": DrawerClose, "button-press-topdown-v3": ButtonPressTopdown,
         "peg-insert-side-v3": PegInsertSide, "window-open-v3": WindowOpen, "window-close-v3": WindowClose,
         "reach-wall-v3": ReachWall, "push-wall-v3": PushWall, "pick-place-wall-v3": PickPlaceWall, "push-back-v3": PushBack,
         "sweep-v3": Sweep, "sweep-into-v3": SweepInto, "hand-insert-v3": HandInsert, "pick-out-of-hole-v3": PickOutOfHole}
