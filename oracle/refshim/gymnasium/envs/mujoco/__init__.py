"""gymnasium.envs.mujoco.MujocoEnv restated (gymnasium 1.1 `mujoco_env.py`), without rendering."""
from __future__ import annotations

import os

import mujoco
import numpy as np

from ... import spaces
from ...core import Env


class MujocoEnv(Env):
    def __init__(self, model_path, frame_skip, observation_space, render_mode=None, width=480, height=480,
                 camera_id=None, camera_name=None, default_camera_config=None, max_geom=1000, visual_options={}):
        self.fullpath = os.path.expanduser(str(model_path))
        if not os.path.exists(self.fullpath):
            raise OSError(f"File {self.fullpath} does not exist")
        self.width, self.height = width, height
        self.model, self.data = self._initialize_simulation()
        self.init_qpos = self.data.qpos.ravel().copy()
        self.init_qvel = self.data.qvel.ravel().copy()
        self.frame_skip = frame_skip
        if "render_fps" in self.metadata:
            assert int(np.round(1.0 / self.dt)) == self.metadata["render_fps"], \
                f'Expected value: {int(np.round(1.0 / self.dt))}, Actual value: {self.metadata["render_fps"]}'
        if observation_space is not None:
            self.observation_space = observation_space
        self._set_action_space()
        self.render_mode = render_mode
        self.camera_name, self.camera_id = camera_name, camera_id

    def _set_action_space(self):
        bounds = self.model.actuator_ctrlrange.copy().astype(np.float32)
        low, high = bounds.T
        self.action_space = spaces.Box(low=low, high=high, dtype=np.float32)
        return self.action_space

    def _initialize_simulation(self):
        model = mujoco.MjModel.from_xml_path(self.fullpath)
        data = mujoco.MjData(model)
        return model, data

    def set_state(self, qpos, qvel):
        assert qpos.shape == (self.model.nq,) and qvel.shape == (self.model.nv,)
        self.data.qpos[:] = np.copy(qpos)
        self.data.qvel[:] = np.copy(qvel)
        mujoco.mj_forward(self.model, self.data)

    def _step_mujoco_simulation(self, ctrl, n_frames):
        self.data.ctrl[:] = ctrl
        mujoco.mj_step(self.model, self.data, nstep=n_frames)
        mujoco.mj_rnePostConstraint(self.model, self.data)

    def do_simulation(self, ctrl, n_frames):
        if np.array(ctrl).shape != (self.model.nu,):
            raise ValueError(f"Action dimension mismatch. Expected {(self.model.nu,)}, found {np.array(ctrl).shape}")
        self._step_mujoco_simulation(ctrl, n_frames)

    def render(self):
        raise NotImplementedError("rendering is outside the hot path")

    def close(self):
        pass

    def reset(self, *, seed=None, options=None):
        super().reset(seed=seed)
        mujoco.mj_resetData(self.model, self.data)
        ob = self.reset_model()
        info = self._get_reset_info()
        return ob, info

    def reset_model(self):
        raise NotImplementedError

    def _get_reset_info(self):
        return {}

    @property
    def dt(self):
        return self.model.opt.timestep * self.frame_skip

    def get_body_com(self, body_name):
        return self.data.body(body_name).xpos

    def state_vector(self):
        return np.concatenate([self.data.qpos.flat, self.data.qvel.flat])
