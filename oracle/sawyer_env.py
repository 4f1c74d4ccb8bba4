"""CPU oracle: Python restatement of the reference's single-env hot path.

TEST INFRASTRUCTURE -- PARITY UNPINNED (physics oracle is a restatement, see
oracle/mjphys.h).  Restates ``SawyerMocapBase`` / ``SawyerXYZEnv``
(metaworld/sawyer_xyz_env.py:26-858 in the reference) on top of the oracle
physics (oracle/mjphys.py) with the same method names, call order and numpy
aliasing behaviour, so the parity tests read like the reference's own tests.
Task classes live in oracle/tasks.py.
"""
from __future__ import annotations

import os

import numpy as np

from metaworld_b200 import mjcf
from oracle import mjphys as P

_MODEL_CACHE: dict = {}

HAND_LOW = np.array([-0.525, 0.348, -0.0525])   # _HAND_SPACE, sawyer_xyz_env.py:146-150
HAND_HIGH = np.array([+0.525, 1.025, 0.7])


def load_model(xml_name):
    if xml_name not in _MODEL_CACHE:
        from metaworld_b200 import modelzoo
        _MODEL_CACHE[xml_name] = modelzoo.full_model(xml_name)
    return _MODEL_CACHE[xml_name]


# ---- reward utilities (restating metaworld/utils/reward_utils.py:27-244) ----
_VAM = 0.1


def tolerance(x, bounds=(0.0, 0.0), margin=0.0, sigmoid="gaussian", value_at_margin=_VAM):
    lower, upper = bounds
    if lower > upper:
        raise ValueError("Lower bound must be <= upper bound.")
    if margin < 0:
        raise ValueError(f"`margin` must be non-negative. Current value: {margin}")
    if lower <= x <= upper:
        return 1.0
    if margin == 0:
        return 0.0
    d = (lower - x if x < lower else x - upper) / margin
    if sigmoid == "gaussian":
        scale = np.sqrt(-2 * np.log(value_at_margin))
        return float(np.exp(-0.5 * (d * scale) ** 2))
    if sigmoid == "long_tail":
        scale = np.sqrt(1 / value_at_margin - 1)
        return float(1 / ((d * scale) ** 2 + 1))
    raise ValueError(sigmoid)


def hamacher_product(a, b):
    if not ((0.0 <= a <= 1.0) and (0.0 <= b <= 1.0)):
        raise ValueError(f"a ({a}) and b ({b}) must range between 0 and 1")
    den = a + b - a * b
    return (a * b) / den if den > 0 else 0


def rect_prism_tolerance(curr, zero, one):
    def in_range(a, b, c):
        return float(b <= a <= c) if c >= b else float(c <= a <= b)

    if in_range(curr[0], zero[0], one[0]) and in_range(curr[1], zero[1], one[1]) and in_range(curr[2], zero[2], one[2]):
        diff = one - zero
        return ((curr[0] - zero[0]) / diff[0]) * ((curr[1] - zero[1]) / diff[1]) * ((curr[2] - zero[2]) / diff[2])
    return 1.0


def mat2quat_xyzw(mat9):
    """scipy ``Rotation.from_matrix(m).as_quat()`` (xyzw, scipy's branch + sign convention)."""
    from scipy.spatial.transform import Rotation

    return Rotation.from_matrix(np.asarray(mat9).reshape(3, 3)).as_quat()


class SawyerXYZEnv:
    """Restates sawyer_xyz_env.py:143-858.  Subclasses define the per-task parts."""

    max_path_length = 500
    TARGET_RADIUS = 0.05
    frame_skip = 5
    action_scale = 1.0 / 100
    xml = None
    hand_low = (-0.2, 0.55, 0.05)
    hand_high = (0.2, 0.75, 0.3)
    mocap_low = None
    mocap_high = None

    def __init__(self, reward_function_version="v2"):
        self.reward_function_version = reward_function_version
        self.hand_low = np.array(self.hand_low, dtype=np.float64)
        self.hand_high = np.array(self.hand_high, dtype=np.float64)
        self.mocap_low = self.hand_low if self.mocap_low is None else np.array(self.mocap_low, dtype=np.float64)
        self.mocap_high = self.hand_high if self.mocap_high is None else np.array(self.mocap_high, dtype=np.float64)
        self.curr_path_length = 0
        self._freeze_rand_vec = True
        self._last_rand_vec = None
        self._partially_observable = True
        self._set_task_called = False
        self.obj_init_pos = None
        self._target_pos = None
        self.model = P.OModel(load_model(self.xml))
        self.data = P.OData(self.model)
        self.reset_mocap_welds()
        P.mj_forward(self.model, self.data)                    # sawyer_xyz_env.py:231
        self.init_left_pad = self.get_body_com("leftpad")      # live views, as in the reference (:236-237)
        self.init_right_pad = self.get_body_com("rightpad")
        self._obs_obj_max_len = 14
        self.init_qpos = np.copy(self.data.qpos)
        self.init_qvel = np.copy(self.data.qvel)
        self.setup()
        self._prev_obs = self._get_curr_obs_combined_no_goal()

    # ---- per-task hooks
    def setup(self):
        raise NotImplementedError

    # ---- MujocoEnv glue [3P gymnasium.envs.mujoco.MujocoEnv]
    def do_simulation(self, ctrl, n_frames):
        self.data.ctrl = np.asarray(ctrl, dtype=np.float64)
        P.mj_step(self.model, self.data, n_frames)

    def set_state(self, qpos, qvel):
        self.data.qpos = qpos
        self.data.qvel = qvel
        P.mj_forward(self.model, self.data)

    def get_body_com(self, name):
        return self.data.body(name).xpos

    # ---- SawyerMocapBase
    def reset_mocap_welds(self):
        for i in range(self.model.eq_data.shape[0]):
            self.model.eq_data[i] = np.array([0.0, 0.0, 0.0, 0.0, 0.0, 0.0, -1.0, 0.0, 0.0, 0.0, 5.0])

    def get_endeff_pos(self):
        return self.data.body("hand").xpos

    @property
    def tcp_center(self):
        return (self.data.site("rightEndEffector").xpos + self.data.site("leftEndEffector").xpos) / 2.0

    # ---- SawyerXYZEnv
    def set_task_vec(self, rand_vec, partially_observable=False):
        """Equivalent of ``set_task(Task)`` after unpickling (sawyer_xyz_env.py:298-318)."""
        self._set_task_called = True
        self._freeze_rand_vec = True
        self._last_rand_vec = np.asarray(rand_vec, dtype=np.float64)
        self._partially_observable = bool(partially_observable)

    def set_xyz_action(self, action):
        action = np.clip(action, -1, 1)
        pos_delta = action * self.action_scale
        new_mocap_pos = self.data.mocap_pos + pos_delta[None]
        new_mocap_pos[0, :] = np.clip(new_mocap_pos[0, :], self.mocap_low, self.mocap_high)
        self.data.mocap_pos = new_mocap_pos
        self.data.mocap_quat = np.array([1, 0, 1, 0])

    def _set_obj_xyz(self, pos):
        qpos = self.data.qpos.flat.copy()
        qvel = self.data.qvel.flat.copy()
        qpos[9:12] = pos.copy()
        qvel[9:15] = 0
        self.set_state(qpos, qvel)

    def _get_site_pos(self, name):
        return self.data.site(name).xpos.copy()

    def _set_pos_site(self, name, pos):
        self.data.site(name).xpos = pos[:3]

    @property
    def _target_site_config(self):
        return [("goal", self._target_pos)]

    def touching_object(self, object_geom_id):
        lid = self.data.geom("leftpad_geom").id
        rid = self.data.geom("rightpad_geom").id
        lf = rf = 0.0
        for c in self.data.contact:
            if c.efc_address < 0:
                continue
            pair = (c.geom1, c.geom2)
            if lid in pair and object_geom_id in pair:
                lf += self.data.efc_force[c.efc_address]
            if rid in pair and object_geom_id in pair:
                rf += self.data.efc_force[c.efc_address]
        return 0 < lf and 0 < rf

    def _get_id_main_object(self):
        return self.data.geom("objGeom").id

    @property
    def touching_main_object(self):
        return self.touching_object(self._get_id_main_object())

    def _get_pos_goal(self):
        return self._target_pos

    def _get_curr_obs_combined_no_goal(self):
        pos_hand = self.get_endeff_pos()
        fr, fl = self.data.body("rightclaw"), self.data.body("leftclaw")
        g = np.clip(np.linalg.norm(fr.xpos - fl.xpos) / 0.1, 0.0, 1.0)
        padded = np.zeros(self._obs_obj_max_len)
        obj_pos = np.asarray(self._get_pos_objects())
        obj_quat = np.asarray(self._get_quat_objects())
        ps = np.split(obj_pos, len(obj_pos) // 3)
        qs = np.split(obj_quat, len(obj_quat) // 4)
        flat = np.hstack([np.hstack((p, q)) for p, q in zip(ps, qs)])
        padded[: len(flat)] = flat
        return np.hstack((pos_hand, g, padded))

    def _get_obs(self):
        pos_goal = self._get_pos_goal()
        if self._partially_observable:
            pos_goal = np.zeros_like(pos_goal)
        curr = self._get_curr_obs_combined_no_goal()
        obs = np.hstack((curr, self._prev_obs, pos_goal))
        self._prev_obs = curr
        return obs

    def obs_bounds(self):
        inf = np.full(14, np.inf)
        if self._partially_observable:
            gl = gh = np.zeros(3)
        else:
            gl, gh = np.array(self.goal_low, dtype=np.float64), np.array(self.goal_high, dtype=np.float64)
        low = np.hstack((HAND_LOW, -1.0, -inf, HAND_LOW, -1.0, -inf, gl))
        high = np.hstack((HAND_HIGH, 1.0, inf, HAND_HIGH, 1.0, inf, gh))
        return low, high

    def step(self, action):
        assert self._set_task_called, "You must call env.set_task before using env.step"
        action = np.asarray(action)
        assert len(action) == 4
        self.set_xyz_action(action[:3])
        if self.curr_path_length >= self.max_path_length:
            raise ValueError("You must reset the env manually once truncate==True")
        self.do_simulation([action[-1], -action[-1]], n_frames=self.frame_skip)
        self.curr_path_length += 1
        for site in self._target_site_config:
            self._set_pos_site(*site)
        P.mj_forward(self.model, self.data)
        obs = self._get_obs()
        low, high = self.obs_bounds()
        obs = np.clip(obs, low, high)
        self._last_stable_obs = obs
        reward, info = self.evaluate_state(obs, action)
        truncate = self.curr_path_length == self.max_path_length
        return np.array(obs, dtype=np.float64), reward, False, truncate, info

    def reset(self):
        self.curr_path_length = 0
        self.reset_model()
        P.mj_resetData(self.model, self.data)       # MujocoEnv.reset [3P]
        obs = self.reset_model()
        self._prev_obs = obs[:18].copy()
        obs[18:36] = self._prev_obs
        return obs.astype(np.float64), {}

    def _reset_hand(self, steps=50):
        for _ in range(steps):
            self.data.mocap_pos[0][:] = self.hand_init_pos
            self.data.mocap_quat[0][:] = np.array([1, 0, 1, 0])
            self.do_simulation([-1, 1], self.frame_skip)
        self.init_tcp = self.tcp_center

    def _get_state_rand_vec(self):
        if self._freeze_rand_vec:
            assert self._last_rand_vec is not None
            return self._last_rand_vec
        lo, hi = self.random_reset_space()
        rand_vec = np.random.uniform(lo, hi, size=lo.size).astype(np.float64)
        self._last_rand_vec = rand_vec
        return rand_vec

    def random_reset_space(self):
        raise NotImplementedError

    # ---- shared caging reward (sawyer_xyz_env.py:721-858)
    def _gripper_caging_reward(self, action, obj_pos, obj_radius, pad_success_thresh, object_reach_radius, xz_thresh,
                               desired_gripper_effort=1.0, high_density=False, medium_density=False):
        left_pad = self.get_body_com("leftpad")
        right_pad = self.get_body_com("rightpad")
        pad_y_lr = np.hstack((left_pad[1], right_pad[1]))
        pad_to_obj_lr = np.abs(pad_y_lr - obj_pos[1])
        pad_to_objinit_lr = np.abs(pad_y_lr - self.obj_init_pos[1])
        caging_lr_margin = np.abs(pad_to_objinit_lr - pad_success_thresh)
        caging_lr = [tolerance(pad_to_obj_lr[i], bounds=(obj_radius, pad_success_thresh), margin=caging_lr_margin[i],
                               sigmoid="long_tail") for i in range(2)]
        caging_y = hamacher_product(*caging_lr)
        tcp = self.tcp_center
        xz = [0, 2]
        caging_xz_margin = np.linalg.norm(self.obj_init_pos[xz] - self.init_tcp[xz]) - xz_thresh
        caging_xz = tolerance(np.linalg.norm(tcp[xz] - obj_pos[xz]), bounds=(0, xz_thresh), margin=caging_xz_margin,
                              sigmoid="long_tail")
        gripper_closed = min(max(0, action[-1]), desired_gripper_effort) / desired_gripper_effort
        caging = hamacher_product(caging_y, float(caging_xz))
        gripping = gripper_closed if caging > 0.97 else 0.0
        caging_and_gripping = hamacher_product(caging, gripping)
        if high_density:
            caging_and_gripping = (caging_and_gripping + caging) / 2
        if medium_density:
            tcp_to_obj = np.linalg.norm(obj_pos - tcp)
            tcp_to_obj_init = np.linalg.norm(self.obj_init_pos - self.init_tcp)
            reach_margin = abs(tcp_to_obj_init - object_reach_radius)
            reach = tolerance(tcp_to_obj, bounds=(0, object_reach_radius), margin=reach_margin, sigmoid="long_tail")
            caging_and_gripping = (caging_and_gripping + float(reach)) / 2
        return caging_and_gripping
