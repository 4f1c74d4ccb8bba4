"""Drop-in vector environment over the CUDA engine.

Stands where the reference puts ``gymnasium.vector.SyncVectorEnv([...partial(_init_each_env ...)])``
(metaworld/__init__.py:460-604): same construction kwargs, same ``reset`` / ``step`` return shapes and dtypes,
SAME_STEP autoreset with ``final_obs`` / ``final_info``, the per-env wrapper stack folded in
(TimeLimit, AutoTerminateOnSuccessWrapper, OneHotWrapper, RecordEpisodeStatistics,
Random/PseudoRandomTaskSelectWrapper, CheckpointWrapper -- metaworld/wrappers.py) and the ``call`` / ``get_attr`` /
``set_attr`` names that ``metaworld/evaluation.py`` and the reference tests use.

Extension over the reference: ``num_envs`` may be any multiple of the number of env types (the reference
ignores it); env ``e`` has type ``e % n_types`` (task ids interleaved) and replica ``e // n_types``.  Replica
``r`` seeds its task-selection RNG with ``seed + r`` so replica 0 reproduces the reference stream.

The numpy API (`reset`, `step`) moves actions host->device and results device->host every call; the
``*_torch`` variants keep everything on the GPU and never synchronise.

Task selection on autoreset.  The reference draws the next task inside ``reset`` (wrappers.py:116-119), i.e. after the
terminal step.  The kernel needs the snapshot id of the next episode BEFORE the step that may end the episode, so the
host draws one task ahead ("pending").  The draw is speculative: whenever the sampler state becomes observable
(checkpoint, ``toggle_sample_tasks_on_reset``, ``sample_tasks``, an explicit ``reset``) it is either consumed as the
draw the reference would make at that point or rewound, so the task sequence and the RNG stream are the reference's.
"""
from __future__ import annotations

import base64

import numpy as np

from . import _gym
from .benchmarks import Task, reference_env_id
from .engine import ENVSTATE_DTYPE, INFO_KEYS, Engine
from .tasks import TASKS

MAX_PATH_LENGTH = 500     # SawyerXYZEnv.max_path_length (sawyer_xyz_env.py:152): truncates whatever TimeLimit says


def _serialize_task(task: Task) -> dict:          # metaworld/wrappers.py:35-39
    return {"env_name": task.env_name, "data": base64.b64encode(task.data).decode("ascii")}


def _deserialize_task(d: dict) -> Task:            # metaworld/wrappers.py:42-47
    assert "env_name" in d and "data" in d
    return Task(env_name=d["env_name"], data=base64.b64decode(d["data"]))


def _rng(seed):
    return np.random.Generator(np.random.PCG64(seed)) if seed is not None else np.random.default_rng()


class _SubEnv:
    """Host mirror of one sub-env's wrapper state (task list, RNGs, flags)."""

    def __init__(self, name, tasks, seed, pseudorandom, env_id):
        self.task_name = name
        self.env_id = env_id
        self.tasks = list(tasks)
        # env.seed(seed) (metaworld/__init__.py:428, sawyer_xyz_env.py:274-292) seeds env.np_random -- which the
        # task-select wrappers use (Wrapper.np_random is the wrapped env's) -- and the three spaces with the same seed
        self.np_random = _rng(seed)
        self.space_rng = {k: _rng(seed) for k in ("action_space", "obs_space", "goal_space")}
        self.pseudorandom = pseudorandom
        self.sample_tasks_on_reset = not pseudorandom
        self.current_task_idx = -1
        self.current_task: Task | None = None
        self.pending: Task | None = None
        self._pre = None          # sampler state before the speculative draw of `pending`

    def next_task(self):
        if self.pseudorandom:     # PseudoRandomTaskSelectWrapper._set_pseudo_random_task (wrappers.py:156-160)
            self.current_task_idx = (self.current_task_idx + 1) % len(self.tasks)
            if self.current_task_idx == 0:
                self.np_random.shuffle(self.tasks)
            return self.tasks[self.current_task_idx]
        idx = self.np_random.choice(len(self.tasks))   # RandomTaskSelectWrapper._set_random_task (wrappers.py:98-100)
        return self.tasks[idx]

    def draw_pending(self):
        """Task of the next episode, drawn one episode early (see module docstring)."""
        if self.sample_tasks_on_reset:
            self._pre = (list(self.tasks), self.current_task_idx, self.np_random.bit_generator.state)
            self.pending = self.next_task()
        else:
            self._pre = None
            self.pending = self.current_task

    def rewind_pending(self):
        if self._pre is not None:
            self.tasks, self.current_task_idx, self.np_random.bit_generator.state = self._pre
        self._pre = None
        self.pending = None

    def take_for_reset(self):
        """What the task-select wrapper's ``reset`` does (wrappers.py:116-119 / 181-184)."""
        if self.sample_tasks_on_reset:
            if self._pre is not None:          # the speculative draw IS the draw the reference makes now
                self.current_task, self._pre, self.pending = self.pending, None, None
            else:
                self.current_task = self.next_task()
        elif self.current_task is None:
            self.current_task = self.next_task()


class MetaWorldVecEnv(_gym.VectorEnvBase):
    metadata = {"render_modes": [], "autoreset_mode": "same_step"}

    def __init__(self, env_names, tasks_per_env, num_envs=None, seed=None, use_one_hot=False, num_tasks=None,
                 env_ids=None, max_episode_steps=None, terminate_on_success=False, task_select="random",
                 reward_function_version="v2", device=0, engine=None, recurrent_info_in_obs=False,
                 normalize_reward_in_recurrent_info=True, reward_normalization_method=None, reward_alpha=0.001,
                 normalize_observations=False, checkpoint_env_ids=None, type_seeds=None, **unused):
        if reward_function_version != "v2":
            raise NotImplementedError("only the default v2 rewards are implemented on the device")
        n_types = len(env_names)
        num_envs = n_types if num_envs is None else int(num_envs)
        if num_envs < n_types:
            raise ValueError(f"num_envs ({num_envs}) must be at least the number of env types ({n_types})")
        self.num_envs = num_envs
        self.n_types = n_types
        self.env_names = list(env_names)
        self.max_episode_steps = int(max_episode_steps or MAX_PATH_LENGTH)
        self.terminate_on_success = bool(terminate_on_success)
        self.use_one_hot = bool(use_one_hot)
        self.num_tasks = int(num_tasks or n_types)
        self.env_ids = list(range(n_types)) if env_ids is None else list(env_ids)
        self._seed = 0 if seed is None else int(seed)
        # one model slot per distinct env name
        uniq = list(dict.fromkeys(env_names))
        self.engine = engine or Engine(uniq, device=device)
        self._own_engine = engine is None
        torch = self.engine.torch
        self.torch = torch
        self.device = self.engine.device
        self._slot = [uniq.index(n) for n in env_names]
        self._slot_of_name = {n: uniq.index(n) for n in uniq}
        # snapshots: one per distinct (model slot, rand_vec, partially_observable)
        self._snap_of: dict = {}
        self._task_of_snap: dict = {}
        self._ensure_snapshots([tk for tasks in tasks_per_env for tk in tasks])
        # CheckpointWrapper ids (metaworld/__init__.py:455: f"{env_cls}_{env_id}"; env_id is None for the ML benchmarks)
        ck_ids = checkpoint_env_ids if checkpoint_env_ids is not None else self.env_ids
        self.sub = []
        for e in range(num_envs):
            t_i, rep = e % n_types, e // n_types
            # every env type shares `seed` (metaworld/__init__.py:497); the custom-mt entry point gives type i seed + i (:762)
            s = (None if seed is None else seed + rep) if type_seeds is None else (None if type_seeds[t_i] is None else type_seeds[t_i] + rep)
            eid = reference_env_id(env_names[t_i], ck_ids[t_i]) if env_names[t_i] in TASKS else f"{env_names[t_i]}_{ck_ids[t_i]}"
            if rep:       # replicas are an extension (the reference has one sub-env per id): keep their ids distinct
                eid += f".r{rep}"
            self.sub.append(_SubEnv(env_names[t_i], tasks_per_env[t_i], s, task_select != "random", eid))
        self.engine.set_envs([self._slot[e % n_types] for e in range(num_envs)])
        self._set_engine_options()
        # spaces
        T = self.num_tasks if self.use_one_hot else 0
        self.obs_dim = 39 + T
        self.obs_dtype = np.float32 if self.use_one_hot else np.float64
        inf = np.full(14, np.inf)
        hl, hh = np.array([-0.525, 0.348, -0.0525]), np.array([0.525, 1.025, 0.7])
        # goal bounds are [0, 0]: the reference builds the space at construction, when every env is still partially
        # observable (sawyer_xyz_env.py:208,546-558), and the wrappers keep that Box (wrappers.py:19-30)
        lo = np.hstack((hl, -1.0, -inf, hl, -1.0, -inf, np.zeros(3), np.zeros(T)))
        hi = np.hstack((hh, 1.0, inf, hh, 1.0, inf, np.zeros(3), np.ones(T)))
        self.single_observation_space = _gym.Box(lo.astype(self.obs_dtype), hi.astype(self.obs_dtype), dtype=self.obs_dtype)
        self.single_action_space = _gym.Box(-np.ones(4, np.float32), np.ones(4, np.float32), dtype=np.float32, seed=seed)
        self.observation_space = _gym.batch_space(self.single_observation_space, num_envs)
        self.action_space = _gym.batch_space(self.single_action_space, num_envs, seed=seed)
        # device buffers
        N = num_envs
        dev = self.device
        self.d_obs = torch.zeros(N, self.obs_dim, device=dev)
        if self.use_one_hot:
            ids = torch.tensor([self.env_ids[e % n_types] for e in range(N)], device=dev)
            self.d_obs[torch.arange(N, device=dev), 39 + ids] = 1.0
        self.d_final_obs = self.d_obs.clone()
        # the numpy API's outputs: the 39 columns k_step writes, packed (stride 39): the constant one-hot columns never cross
        # the bus, and the host converts a contiguous [N, 39] block
        self.d_obs39 = torch.zeros(N, 39, device=dev)
        self.d_final_obs39 = torch.zeros(N, 39, device=dev)
        self.d_reward = torch.zeros(N, device=dev)
        self.d_term = torch.zeros(N, dtype=torch.uint8, device=dev)
        self.d_trunc = torch.zeros(N, dtype=torch.uint8, device=dev)
        self.d_small = torch.zeros(N, 9, device=dev)        # info[7], reward, flags (terminated + 2 truncated): one D2H record
        self.d_info = self.d_small[:, :7]
        self.d_final_info = torch.zeros(N, 8, device=dev)
        self.d_actions = torch.zeros(N, 4, device=dev)
        self.d_next = torch.zeros(N, dtype=torch.int32, device=dev)
        self.d_cur = torch.zeros(N, dtype=torch.int32, device=dev)
        self.d_idx = torch.zeros(N, dtype=torch.int64, device=dev)
        on_gpu = self.device.type == "cuda"      # (the host-logic tests drive this class with a CPU stand-in for the engine)
        pin = (lambda t: t.pin_memory()) if on_gpu else (lambda t: t)
        self.h_actions = pin(torch.zeros(N, 4))
        self.h_next = pin(torch.zeros(N, dtype=torch.int32))
        self.h_idx = pin(torch.zeros(N, dtype=torch.int64))
        self.h_obs = pin(torch.zeros(N, self.obs_dim))
        self.h_obs39 = pin(torch.zeros(N, 39))
        self.h_final_obs39 = pin(torch.zeros(N, 39))
        self.h_small = pin(torch.zeros(N, 9))
        self.h_final_info = pin(torch.zeros(N, 8))
        self._next_ids = np.zeros(N, dtype=np.int32)
        self._ep_len = np.zeros(N, dtype=np.int64)
        self._closed = False
        self._needs_reset = True
        self._device_sampler = False
        # optional per-sub-env wrappers of the reference that sit above the one-hot wrapper (metaworld/__init__.py:437-444)
        from .post import StepPost
        self.post = StepPost(N, recurrent_info_in_obs, normalize_reward_in_recurrent_info, reward_normalization_method, reward_alpha,
                             normalize_observations)
        if self.post.recurrent:     # RNNBasedMetaRLWrapper's space: unbounded float32 of obs + action + reward + done (wrappers.py:55-62)
            D = self.obs_dim + self.post.extra
            self.obs_dtype = np.float32
            self.single_observation_space = _gym.Box(np.full(D, -np.inf, np.float32), np.full(D, np.inf, np.float32), dtype=np.float32)
            self.observation_space = _gym.batch_space(self.single_observation_space, num_envs)
        if self.post.norm_obs:      # gymnasium.wrappers.NormalizeObservation: unbounded float32 space of the same shape
            D = self.obs_dim + self.post.extra
            self.single_observation_space = _gym.Box(np.full(D, -np.inf, np.float32), np.full(D, np.inf, np.float32), dtype=np.float32)
            self.observation_space = _gym.batch_space(self.single_observation_space, num_envs)

    # ------------------------------------------------------------------ helpers
    def _set_engine_options(self):
        # SawyerXYZEnv truncates at its own max_path_length = 500 whatever the TimeLimit wrapper says (sawyer_xyz_env.py:634)
        self.engine.set_options(min(self.max_episode_steps, MAX_PATH_LENGTH), self.terminate_on_success, self._seed)

    @staticmethod
    def _task_key(slot, d):
        return (slot, np.asarray(d["rand_vec"], dtype=np.float64).tobytes(), bool(d["partially_observable"]))

    def _ensure_snapshots(self, tasks):
        """Episode-start snapshots for tasks the engine has not seen yet (construction, checkpoints with other goals)."""
        mi, rvs, po, keys = [], [], [], []
        for tk in tasks:
            d = tk.unpack()
            key = self._task_key(self._slot_of_name[tk.env_name], d)
            if key in self._snap_of or key in keys:
                continue
            v = np.asarray(d["rand_vec"], dtype=np.float64)
            rv = np.zeros(6)
            rv[: len(v)] = v
            keys.append(key); mi.append(key[0]); rvs.append(rv); po.append(key[2])
            self._task_of_snap[key] = tk
        if keys:
            ids = self.engine.build_snapshots(mi, np.array(rvs), po)
            for k, i in zip(keys, ids):
                self._snap_of[k] = int(i)
                self._task_of_snap[int(i)] = self._task_of_snap.pop(k)

    def _snap(self, task: Task) -> int:
        i = task.__dict__.get("_snap_id")
        if i is None or task.__dict__.get("_snap_owner") is not self:
            i = self._snap_of[self._task_key(self._slot_of_name[task.env_name], task.unpack())]
            task.__dict__["_snap_id"], task.__dict__["_snap_owner"] = i, self
        return i

    def _push_next(self):
        self.h_next.copy_(self.torch.from_numpy(self._next_ids))
        self.d_next.copy_(self.h_next, non_blocking=True)

    def _draw_pending(self, e):
        s = self.sub[e]
        s.draw_pending()
        self._next_ids[e] = self._snap(s.pending)

    def _redraw_all_pending(self):
        for e, s in enumerate(self.sub):
            s.rewind_pending()
            self._draw_pending(e)
        self._push_next()

    # ------------------------------------------------------------------ VectorEnv API
    def reset(self, *, seed=None, options=None):
        """Every sub-env: (task-select wrapper) pick a task, then SawyerXYZEnv.reset (its `seed` argument is ignored,
        sawyer_xyz_env.py:670)."""
        N = self.num_envs
        cur = np.zeros(N, dtype=np.int32)
        for e, s in enumerate(self.sub):
            s.take_for_reset()
            cur[e] = self._snap(s.current_task)
        for e in range(N):
            self._draw_pending(e)
        self.d_cur.copy_(self.torch.from_numpy(cur))
        self._push_next()
        self.engine.reset(self.d_cur, self.d_obs)
        self._ep_len[:] = 0
        self._needs_reset = False
        self.h_obs.copy_(self.d_obs)
        obs = self.h_obs.numpy().astype(self.obs_dtype)
        self._obs_template = obs.copy()        # constant columns (one-hot task id) of every later observation; see step()
        self._obs_template[:, :39] = 0
        if self.post.active:
            obs = self.post.on_reset(obs)
        return obs, {}

    def _advance_streams(self, idx):
        """The autoreset's task-select draw of sub-envs `idx` (their episode ends with this step): the speculative draw
        becomes the running task and the draw for the episode after it is made."""
        for e in idx:
            s = self.sub[e]
            s.current_task, s._pre = s.pending, None
            self._draw_pending(e)
        self._push_next()

    def step(self, actions):
        if self._needs_reset:
            raise RuntimeError("reset() must be called before step()")
        if self._device_sampler:
            raise RuntimeError("the device-side task sampler is active (step_torch was used): the numpy step API and its host "
                               "task streams are no longer in sync; call disable_device_sampler() + reset() first")
        t = self.torch
        N = self.num_envs
        a = np.ascontiguousarray(actions, dtype=np.float32).reshape(N, 4)
        self.h_actions.copy_(t.from_numpy(a))
        self.d_actions.copy_(self.h_actions, non_blocking=True)
        self.engine.step(self.d_actions, self.d_obs39, self.d_reward, self.d_term, self.d_trunc, self.d_small,
                         self.d_final_obs39, self.d_final_info, self.d_next)
        self.h_small.copy_(self.d_small, non_blocking=True)
        self.h_obs39.copy_(self.d_obs39, non_blocking=True)
        # ---- while the kernel runs: everything that does not need its results.
        # Truncations are known in advance (the host mirrors the episode lengths): the terminal rows of those envs are
        # fetched in the same batch of copies, their task streams are advanced and the snapshot ids of the episodes after
        # the coming ones are queued behind the kernel (stream order: k_step reads d_next before this copy overwrites it).
        pred = np.nonzero(self._ep_len + 1 >= min(self.max_episode_steps, MAX_PATH_LENGTH))[0]
        npred = len(pred)
        if npred:
            self.h_idx[:npred] = t.from_numpy(pred)
            d_pred = self.d_idx[:npred]
            d_pred.copy_(self.h_idx[:npred], non_blocking=True)
            self.h_final_obs39[:npred].copy_(self.d_final_obs39.index_select(0, d_pred), non_blocking=True)
            self.h_final_info[:npred].copy_(self.d_final_info.index_select(0, d_pred), non_blocking=True)
            self._advance_streams(pred)
        # fresh arrays every step, like the reference (:637).  The array is allocated here, off the critical path, as a copy
        # of a template that already holds the constant columns (the one-hot task id): only the 39 columns the kernel writes
        # remain to be converted once the results are there
        obs = self._obs_template.copy()
        if self.device.type == "cuda":
            t.cuda.current_stream(self.device).synchronize()
        # ---- results (single-threaded numpy on purpose: torch's parallel host copies are faster when idle but collapse
        # under a cgroup CPU quota smaller than the machine's core count)
        np.copyto(obs[:, :39], self.h_obs39.numpy())
        sm = np.ascontiguousarray(self.h_small.numpy().T, dtype=np.float64)      # [9, N]: rows are contiguous per-key arrays
        reward = sm[7]
        flags = sm[8].astype(np.int8)
        terminated, truncated = (flags & 1).astype(bool), (flags & 2).astype(bool)
        self._ep_len += 1
        done = terminated | truncated
        idx = np.nonzero(done)[0]
        any_done = len(idx) > 0
        # SAME_STEP (gymnasium SyncVectorEnv): a finished env's step info moves to `final_info` and its slot in the
        # top-level arrays is the (empty) reset info -> value 0, mask False; keys vanish when every env finished
        infos = {}
        if len(idx) < N:
            live = ~done
            for i, k in enumerate(INFO_KEYS):
                v = sm[i]
                if any_done:
                    v[idx] = 0.0
                infos[k] = v
                infos["_" + k] = live.copy()
        fo = ep_r = None
        if any_done:
            # terminal observations / infos of the finished envs only (a few rows per step in steady state)
            if np.array_equal(idx, pred):            # exactly the predicted truncations (always, unless success terminates)
                rows39, rows_i = self.h_final_obs39[:npred].numpy(), self.h_final_info[:npred].numpy().copy()
            else:
                d_idx = t.from_numpy(idx).to(self.device, non_blocking=True)
                rows39 = self.d_final_obs39.index_select(0, d_idx).cpu().numpy()
                rows_i = self.d_final_info.index_select(0, d_idx).cpu().numpy()
            rows_o = self._obs_template[idx]
            rows_o[:, :39] = rows39
        if self.post.active:
            if any_done:
                fo = np.zeros((N, self.obs_dim), dtype=self.obs_dtype)
                fo[idx] = rows_o
            obs, reward, fo, ep_r = self.post.on_step(obs, a, reward, terminated, truncated, final_obs=fo)
            if any_done:
                rows_o = fo[idx]
        if any_done:
            fi = np.zeros((8, N))
            fi[:, idx] = rows_i.T
            if ep_r is not None:
                fi[7] = ep_r             # RecordEpisodeStatistics sits outside the reward normalisation
            final_obs = np.full(N, None, dtype=object)
            for j, e in enumerate(idx):
                final_obs[e] = rows_o[j]
            final_info = {}
            for i, k in enumerate(INFO_KEYS):
                final_info[k] = fi[i]                              # rows of unfinished envs are 0
                final_info["_" + k] = done.copy()
            final_info["episode"] = {"r": fi[7], "l": np.where(done, self._ep_len, 0),
                                     "t": np.zeros(N), "_r": done.copy(), "_l": done.copy(), "_t": done.copy()}
            final_info["_episode"] = done.copy()
            infos["final_obs"], infos["_final_obs"] = final_obs, done.copy()
            infos["final_info"], infos["_final_info"] = final_info, done.copy()
            self._ep_len[idx] = 0
            if npred < len(idx):                     # terminations nobody could predict (terminate_on_success)
                self._advance_streams(np.setdiff1d(idx, pred))
        return obs, reward, terminated, truncated, infos

    def step_async(self, actions):
        self._pending_actions = actions

    def step_wait(self):
        return self.step(self._pending_actions)

    # GPU-resident variants (no host synchronisation; task re-sampling on autoreset happens on the device).  The optional
    # recurrent-obs / reward-normalisation wrappers are applied on the device as well (post.StepPostTorch).
    def enable_device_sampler(self):
        """Autoreset draws the next goal on the device: uniform over the env's own task list, a counter-based hash of
        (seed, env, episode) -- the distribution of RandomTaskSelectWrapper, not its PCG64 stream.  The host task mirrors
        stop being advanced; `get_attr("_last_rand_vec")` etc. then read the snapshot id back from the device."""
        first, count = [], []
        for e, s in enumerate(self.sub):
            ids = sorted(self._snap(tk) for tk in s.tasks)
            assert ids == list(range(ids[0], ids[0] + len(ids))), "device sampler needs contiguous snapshot ranges"
            first.append(ids[0]); count.append(len(ids))
        self.engine.set_goal_sets(first, count)
        self._device_sampler = True

    def disable_device_sampler(self):
        self._device_sampler = False
        self._needs_reset = True

    def _post_torch(self):
        if getattr(self, "_ptorch", None) is None:
            from .post import StepPostTorch
            p = self.post
            self._ptorch = StepPostTorch(self.torch, self.device, self.num_envs, self.obs_dim, p.recurrent, p.norm_in_obs,
                                         "exponential" if p.exponential else "gymnasium" if p.gym_reward else None, p.alpha, p.norm_obs)
            self._ptorch.load_host_state(p)
        return self._ptorch

    def reset_torch(self):
        self.reset()
        return self._post_torch().on_reset(self.d_obs) if self.post.active else self.d_obs

    def step_torch(self, actions):
        """`actions`: float32 CUDA tensor [num_envs, 4] on this env's device.  Returns device tensors (obs [N, obs_dim],
        reward [N], terminated u8 [N], truncated u8 [N], info [N, 7]) that are overwritten by the next call."""
        t = self.torch
        if self._needs_reset:
            raise RuntimeError("reset() must be called before step_torch()")
        if not (isinstance(actions, t.Tensor) and actions.dtype == t.float32 and actions.device == self.device
                and tuple(actions.shape) == (self.num_envs, 4) and actions.is_contiguous()):
            raise ValueError(f"step_torch needs a contiguous float32 tensor of shape ({self.num_envs}, 4) on {self.device}")
        if not self._device_sampler and any(s.sample_tasks_on_reset for s in self.sub):
            self.enable_device_sampler()      # without it every autoreset would restart the same pre-drawn goal
        nxt = None if self._device_sampler else self.d_next
        self.engine.step(actions, self.d_obs, self.d_reward, self.d_term, self.d_trunc, self.d_small, self.d_final_obs,
                         self.d_final_info, nxt)
        if self.post.active:       # RNNBasedMetaRLWrapper / NormalizeRewardsExponential on the device; the terminal observation
            obs, rew, self.d_final_obs_post, self.d_episode_return_post = self._post_torch().on_step(    # and episode returns stay available
                self.d_obs, actions, self.d_reward, self.d_term, self.d_trunc, self.d_final_obs)
            return obs, rew, self.d_term, self.d_trunc, self.d_info
        return self.d_obs, self.d_reward, self.d_term, self.d_trunc, self.d_info

    # attribute RPC used by metaworld/evaluation.py:48-169 and the reference tests
    def _current_tasks(self):
        if self._device_sampler:       # the device chose the goals: read the snapshot id of every env's running episode
            snap = self.engine.get_state()["snapshot"].astype(np.int64)
            return [self._task_of_snap[int(i)] for i in snap]
        return [s.current_task for s in self.sub]

    def get_attr(self, name):
        if name == "terminate_on_success":
            return tuple([self.terminate_on_success] * self.num_envs)
        if name == "task_name":
            return tuple(s.task_name for s in self.sub)
        if name == "tasks":
            return tuple(s.tasks for s in self.sub)
        if name in ("_last_rand_vec", "_partially_observable"):
            key = "rand_vec" if name == "_last_rand_vec" else "partially_observable"
            return tuple(None if tk is None else tk.unpack()[key] for tk in self._current_tasks())
        if name == "max_path_length":
            return tuple([MAX_PATH_LENGTH] * self.num_envs)
        if name == "curr_path_length":
            return tuple(int(x) for x in self._ep_len)
        if name == "sample_tasks_on_reset":
            return tuple(s.sample_tasks_on_reset for s in self.sub)
        if name == "env_id":
            return tuple(s.env_id for s in self.sub)
        raise AttributeError(name)

    def set_attr(self, name, values):
        vals = values if isinstance(values, (list, tuple)) else [values] * self.num_envs
        if name == "terminate_on_success":
            self.call("toggle_terminate_on_success", bool(vals[0]))
        elif name == "sample_tasks_on_reset":
            for s in self.sub:
                s.rewind_pending()
            for s, v in zip(self.sub, vals):
                s.sample_tasks_on_reset = bool(v)
            if not self._needs_reset:
                self._redraw_all_pending()
        else:
            raise AttributeError(name)

    def call(self, name, *args, **kwargs):
        if name == "toggle_terminate_on_success":
            self.terminate_on_success = bool(args[0])
            self._set_engine_options()
            return tuple([None] * self.num_envs)
        if name == "toggle_sample_tasks_on_reset":
            self.set_attr("sample_tasks_on_reset", bool(args[0]))
            return tuple([None] * self.num_envs)
        if name == "sample_tasks":        # wrappers.py:121-123 / 186-188: draw a task, then reset
            for s in self.sub:
                s.rewind_pending()
                s.current_task = s.next_task()
            saved = [s.sample_tasks_on_reset for s in self.sub]
            for s in self.sub:
                s.sample_tasks_on_reset = False
            obs, info = self.reset()
            for s, v in zip(self.sub, saved):
                s.sample_tasks_on_reset = v
            self._redraw_all_pending()
            return tuple((obs[e], {}) for e in range(self.num_envs))
        if name == "get_checkpoint":
            return self.get_checkpoint()
        if name == "load_checkpoint":
            self.load_checkpoint(args[0])
            return tuple([None] * self.num_envs)
        return self.get_attr(name)

    # ------------------------------------------------------------------ checkpoint (metaworld/wrappers.py:125-142,190-204,275-322)
    def get_checkpoint(self, physics=True):
        """Tuple of the reference's per-sub-env ``CheckpointWrapper.get_checkpoint()`` results: ``(env_id, dict)`` with
        ``tasks`` (base64), ``rng_state`` (random select) or ``current_task_idx`` (pseudorandom), ``sample_tasks_on_reset``
        and ``env_rng_state``.  Extension (ignored by the reference's loader): ``mw_b200`` holds the env's 512-byte device
        record (qpos/qvel/warm start/mocap/frame stack/reward latches/episode counters) and the host mirrors, so a resumed
        run continues bit-identically mid-episode."""
        st = self.engine.get_state() if (physics and not self._needs_reset) else None
        out = []
        for e, s in enumerate(self.sub):
            tasks, idx, rng_state = s._pre if s._pre is not None else (s.tasks, s.current_task_idx, s.np_random.bit_generator.state)
            ck = {"tasks": [_serialize_task(t) for t in tasks]}
            if s.pseudorandom:
                ck["current_task_idx"] = idx
            else:
                ck["rng_state"] = rng_state
            ck["sample_tasks_on_reset"] = s.sample_tasks_on_reset
            ck["env_rng_state"] = {"np_random_state": rng_state,
                                   "action_space_rng_state": s.space_rng["action_space"].bit_generator.state,
                                   "obs_space_rng_state": s.space_rng["obs_space"].bit_generator.state,
                                   "goal_space_rng_state": s.space_rng["goal_space"].bit_generator.state}
            ext = {"ep_len": int(self._ep_len[e]),
                   "current_task": None if s.current_task is None else _serialize_task(s.current_task)}
            if st is not None:
                ext["state"] = base64.b64encode(st[e].tobytes()).decode("ascii")
            ck["mw_b200"] = ext
            out.append((s.env_id, ck))
        return tuple(out)

    def load_checkpoint(self, ckpts):
        """Accepts what the reference's ``envs.call("load_checkpoint", ckpts)`` is given: the list of ``(env_id, dict)``
        tuples; every sub-env takes the entry with its own env_id (k-th env with an id takes the k-th entry with that id:
        the ML benchmarks give every sub-env the id ``..._None``)."""
        ckpts = list(ckpts)
        used = [False] * len(ckpts)
        mine = []
        for s in self.sub:
            hit = None
            for i, (env_id, ck) in enumerate(ckpts):
                if env_id == s.env_id and not used[i]:
                    hit = i
                    break
            if hit is None:
                raise ValueError(f"Could not load checkpoint, no checkpoint found with id {s.env_id}. Checkpoint IDs: ",
                                 [env_id for env_id, _ in ckpts])
            used[hit] = True
            mine.append(ckpts[hit][1])
        new_tasks = []
        for ck in mine:
            for k in ("tasks", "sample_tasks_on_reset", "env_rng_state"):
                assert k in ck
            new_tasks += [_deserialize_task(t) for t in ck["tasks"]]
        self._ensure_snapshots(new_tasks)
        st = None
        for e, (s, ck) in enumerate(zip(self.sub, mine)):
            s.pending, s._pre = None, None
            s.tasks = [_deserialize_task(t) for t in ck["tasks"]]
            if s.pseudorandom:
                assert "current_task_idx" in ck
                s.current_task_idx = ck["current_task_idx"]
            else:
                assert "rng_state" in ck
            s.sample_tasks_on_reset = ck["sample_tasks_on_reset"]
            ers = ck["env_rng_state"]
            s.np_random.bit_generator.state = ers["np_random_state"] if s.pseudorandom else ck["rng_state"]
            s.space_rng["action_space"].bit_generator.state = ers["action_space_rng_state"]
            s.space_rng["obs_space"].bit_generator.state = ers["obs_space_rng_state"]
            s.space_rng["goal_space"].bit_generator.state = ers["goal_space_rng_state"]
            ext = ck.get("mw_b200")
            if ext is not None:
                self._ep_len[e] = ext["ep_len"]
                s.current_task = None if ext["current_task"] is None else _deserialize_task(ext["current_task"])
                if "state" in ext:
                    if st is None:
                        st = self.engine.get_state()
                    st[e] = np.frombuffer(base64.b64decode(ext["state"]), dtype=ENVSTATE_DTYPE)[0]
        if st is not None:
            # snapshot ids are engine-local: re-point every restored record at this engine's id for the same task
            for e, s in enumerate(self.sub):
                if s.current_task is not None:
                    self._ensure_snapshots([s.current_task])
                    st[e]["snapshot"] = self._snap(s.current_task)
            self.engine.set_state(st)
            self._needs_reset = False
        if not self._needs_reset:
            for e, s in enumerate(self.sub):
                if s.current_task is None:
                    self._needs_reset = True
            if not self._needs_reset:
                for e in range(self.num_envs):
                    self._draw_pending(e)
                self._push_next()

    def close(self, **kwargs):
        if not self._closed and self._own_engine:
            self.engine.close()
        self._closed = True


def make_mt_envs(name, seed=None, num_tasks=None, num_envs=None, **kwargs):
    """``make_mt_envs`` (metaworld/__init__.py:460-513) -> MetaWorldVecEnv.  ``vector_strategy`` is accepted
    and ignored (there is one strategy: the GPU).  For a task name the reference returns ONE wrapped env, not a vector
    (:470-478): pass ``single=True`` (what ``gym.make("Meta-World/MT1", ...)`` does) to get that object."""
    from . import benchmarks as B

    if kwargs.pop("single", False):
        from .single_env import MetaWorldSingleEnv
        return MetaWorldSingleEnv(make_mt_envs(name, seed=seed, num_tasks=num_tasks, num_envs=1, **kwargs))

    kwargs.pop("vector_strategy", None); kwargs.pop("autoreset_mode", None)
    bench = B.make_benchmark(name, seed, kwargs.pop("num_goals", B.N_GOALS))
    names = list(bench.train_classes)
    default = {"MT10": 10, "MT25": 25, "MT50": 50}.get(name, 1)
    tasks = [[t for t in bench.train_tasks if t.env_name == n] for n in names]
    if name in TASKS:       # MT1: _init_each_env is called without env_id (metaworld/__init__.py:471-477)
        kwargs.setdefault("checkpoint_env_ids", [None])
    return MetaWorldVecEnv(names, tasks, num_envs=num_envs, seed=seed, num_tasks=num_tasks or default, **kwargs)


def make_ml_envs(name, seed=None, meta_batch_size=20, total_tasks_per_cls=None, split="train", num_envs=None, **kwargs):
    """``make_ml_envs`` / ``_make_ml_envs_inner`` (metaworld/__init__.py:515-604)."""
    from . import benchmarks as B

    kwargs.pop("vector_strategy", None); kwargs.pop("autoreset_mode", None)
    ng = kwargs.pop("num_goals", B.N_GOALS)
    bench = B.ML1(name, seed, ng) if name in TASKS else B.make_benchmark(name, seed, ng)
    classes = list(bench.train_classes if split == "train" else bench.test_classes)
    all_tasks = bench.train_tasks if split == "train" else bench.test_tasks
    assert meta_batch_size % len(classes) == 0, "meta_batch_size must be divisible by envs_per_task"
    per = meta_batch_size // len(classes)
    names, tasks = [], []
    for n in classes:
        ts = [t for t in all_tasks if t.env_name == n]
        if total_tasks_per_cls is not None:
            ts = ts[:total_tasks_per_cls]
        for i in range(per):
            names.append(n); tasks.append(ts[i::per])
    kwargs.setdefault("task_select", "pseudorandom")
    kwargs.setdefault("checkpoint_env_ids", [None] * len(names))     # _init_each_env gets no env_id here (:548-560)
    return MetaWorldVecEnv(names, tasks, num_envs=num_envs, seed=seed, **kwargs)


def make_custom_mt_envs(envs_list, seed=None, use_one_hot=False, num_envs=None, **kwargs):
    """``Meta-World/custom-mt-envs`` (metaworld/__init__.py:741-783): sub-env i is ``make_mt_envs(envs_list[i],
    num_tasks=len(envs_list), env_id=i, seed=seed + i)``, i.e. an MT1 benchmark of its own with its own seed."""
    from . import benchmarks as B

    kwargs.pop("vector_strategy", None); kwargs.pop("autoreset_mode", None)
    ng = kwargs.pop("num_goals", B.N_GOALS)
    seeds = [None if not seed else seed + i for i in range(len(envs_list))]
    tasks = [B.MT1(n, sd, ng).train_tasks for n, sd in zip(envs_list, seeds)]
    return MetaWorldVecEnv(list(envs_list), tasks, num_envs=num_envs, seed=seed, use_one_hot=use_one_hot, num_tasks=len(envs_list),
                           type_seeds=seeds, **kwargs)


def make_custom_ml_envs(train_envs, test_envs, seed=None, meta_batch_size=20, total_tasks_per_cls=None, split="train", num_envs=None, **kwargs):
    """``Meta-World/custom-ml-envs`` (metaworld/__init__.py:370-395,785-820): ``CustomML`` + ``_make_ml_envs_inner``."""
    from . import benchmarks as B

    if set(train_envs) & set(test_envs):
        raise ValueError("The test tasks cannot contain any of the train tasks.")
    kwargs.pop("vector_strategy", None); kwargs.pop("autoreset_mode", None)
    bench = B.Benchmark(train_envs, test_envs, True, seed, n_goals=kwargs.pop("num_goals", B.N_GOALS))
    classes = list(bench.train_classes if split == "train" else bench.test_classes)
    all_tasks = bench.train_tasks if split == "train" else bench.test_tasks
    assert meta_batch_size % len(classes) == 0, "meta_batch_size must be divisible by envs_per_task"
    per = meta_batch_size // len(classes)
    names, tasks = [], []
    for n in classes:
        ts = [t for t in all_tasks if t.env_name == n]
        if total_tasks_per_cls is not None:
            ts = ts[:total_tasks_per_cls]
        for i in range(per):
            names.append(n); tasks.append(ts[i::per])
    # _make_ml_envs_inner is reached without the pseudorandom partial here: _init_each_env's default task_select="random"
    kwargs.setdefault("checkpoint_env_ids", [None] * len(names))
    return MetaWorldVecEnv(names, tasks, num_envs=num_envs, seed=seed, **kwargs)
