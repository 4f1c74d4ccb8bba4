"""Drop-in vector environment over the CUDA engine.

Stands where the reference puts ``gymnasium.vector.SyncVectorEnv([...partial(_init_each_env ...)])``
(metaworld/__init__.py:460-604): same construction kwargs, same ``reset`` / ``step`` return shapes and dtypes,
SAME_STEP autoreset with ``final_obs`` / ``final_info``, the per-env wrapper stack folded in
(TimeLimit, AutoTerminateOnSuccessWrapper, OneHotWrapper, RecordEpisodeStatistics,
Random/PseudoRandomTaskSelectWrapper -- metaworld/wrappers.py) and the ``call`` / ``get_attr`` / ``set_attr``
names that ``metaworld/evaluation.py`` and the reference tests use.

Extension over the reference: ``num_envs`` may be any multiple of the number of env types (the reference
ignores it); env ``e`` has type ``e % n_types`` (task ids interleaved) and replica ``e // n_types``.  Replica
``r`` seeds its task-selection RNG with ``seed + r`` so replica 0 reproduces the reference stream.

The numpy API (`reset`, `step`) moves actions host->device and results device->host every call; the
``*_torch`` variants keep everything on the GPU and never synchronise.
"""
from __future__ import annotations

import numpy as np

from . import _gym
from .benchmarks import Task
from .engine import INFO_KEYS, Engine
from .tasks import TASKS


class _SubEnv:
    """Host mirror of one sub-env's wrapper state (task list, RNG, flags)."""

    def __init__(self, name, tasks, seed, pseudorandom):
        self.task_name = name
        self.tasks = list(tasks)
        self.np_random = np.random.Generator(np.random.PCG64(seed)) if seed is not None else np.random.default_rng()
        self.pseudorandom = pseudorandom
        self.sample_tasks_on_reset = not pseudorandom
        self.current_task_idx = -1
        self.current_task: Task | None = None

    def next_task(self):
        if self.pseudorandom:     # PseudoRandomTaskSelectWrapper._set_pseudo_random_task (wrappers.py:156-160)
            self.current_task_idx = (self.current_task_idx + 1) % len(self.tasks)
            if self.current_task_idx == 0:
                self.np_random.shuffle(self.tasks)
            return self.tasks[self.current_task_idx]
        idx = self.np_random.choice(len(self.tasks))   # RandomTaskSelectWrapper._set_random_task (wrappers.py:98-100)
        return self.tasks[idx]


class MetaWorldVecEnv(_gym.VectorEnvBase):
    metadata = {"render_modes": [], "autoreset_mode": "same_step"}

    def __init__(self, env_names, tasks_per_env, num_envs=None, seed=None, use_one_hot=False, num_tasks=None,
                 env_ids=None, max_episode_steps=None, terminate_on_success=False, task_select="random",
                 reward_function_version="v2", device=0, engine=None, recurrent_info_in_obs=False,
                 normalize_reward_in_recurrent_info=True, reward_normalization_method=None, reward_alpha=0.001,
                 normalize_observations=False, **unused):
        if reward_function_version != "v2":
            raise NotImplementedError("only the default v2 rewards are implemented on the device")
        if normalize_observations:
            raise NotImplementedError("normalize_observations relies on gymnasium.wrappers.NormalizeObservation; not provided")
        n_types = len(env_names)
        num_envs = n_types if num_envs is None else int(num_envs)
        if num_envs < n_types:
            raise ValueError(f"num_envs ({num_envs}) must be at least the number of env types ({n_types})")
        self.num_envs = num_envs
        self.n_types = n_types
        self.env_names = list(env_names)
        self.max_episode_steps = int(max_episode_steps or 500)
        self.terminate_on_success = bool(terminate_on_success)
        self.use_one_hot = bool(use_one_hot)
        self.num_tasks = int(num_tasks or n_types)
        self.env_ids = list(range(n_types)) if env_ids is None else list(env_ids)
        # one model slot per distinct env name
        uniq = list(dict.fromkeys(env_names))
        self.engine = engine or Engine(uniq, device=device)
        self._own_engine = engine is None
        torch = self.engine.torch
        self.torch = torch
        self.device = self.engine.device
        self._slot = [uniq.index(n) for n in env_names]
        # snapshots: one per distinct Task object
        self._snap_of = {}
        mi, rvs, po, keys = [], [], [], []
        for t_i, tasks in enumerate(tasks_per_env):
            for tk in tasks:
                if id(tk) in self._snap_of:
                    continue
                d = tk.unpack()
                self._snap_of[id(tk)] = len(keys)
                keys.append(tk)
                v = np.asarray(d["rand_vec"], dtype=np.float64)
                rv = np.zeros(6)
                rv[: len(v)] = v
                mi.append(self._slot[t_i])
                rvs.append(rv)
                po.append(bool(d["partially_observable"]))
        base = self.engine.build_snapshots(mi, np.array(rvs), po)
        self._snap_base = int(base[0])
        self.sub = []
        for e in range(num_envs):
            t_i, rep = e % n_types, e // n_types
            s = None if seed is None else seed + rep
            self.sub.append(_SubEnv(env_names[t_i], tasks_per_env[t_i], s, task_select != "random"))
        self.engine.set_envs([self._slot[e % n_types] for e in range(num_envs)])
        self.engine.set_options(self.max_episode_steps, self.terminate_on_success, 0 if seed is None else seed)
        # spaces
        T = self.num_tasks if self.use_one_hot else 0
        self.obs_dim = 39 + T
        self.obs_dtype = np.float32 if self.use_one_hot else np.float64
        inf = np.full(14, np.inf)
        hl, hh = np.array([-0.525, 0.348, -0.0525]), np.array([0.525, 1.025, 0.7])
        lo = np.hstack((hl, -1.0, -inf, hl, -1.0, -inf, np.full(3, -np.inf), np.zeros(T)))
        hi = np.hstack((hh, 1.0, inf, hh, 1.0, inf, np.full(3, np.inf), np.ones(T)))
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
        self.d_reward = torch.zeros(N, device=dev)
        self.d_term = torch.zeros(N, dtype=torch.uint8, device=dev)
        self.d_trunc = torch.zeros(N, dtype=torch.uint8, device=dev)
        self.d_info = torch.zeros(N, 7, device=dev)
        self.d_final_info = torch.zeros(N, 8, device=dev)
        self.d_actions = torch.zeros(N, 4, device=dev)
        self.d_next = torch.zeros(N, dtype=torch.int32, device=dev)
        self.d_cur = torch.zeros(N, dtype=torch.int32, device=dev)
        on_gpu = self.device.type == "cuda"      # (the host-logic tests drive this class with a CPU stand-in for the engine)
        pin = (lambda t: t.pin_memory()) if on_gpu else (lambda t: t)
        self.h_actions = pin(torch.zeros(N, 4))
        self.h_next = pin(torch.zeros(N, dtype=torch.int32))
        self.h_obs = pin(torch.zeros(N, self.obs_dim))
        self.h_small = pin(torch.zeros(N, 9))      # reward, info[7], flags
        self._ep_len = np.zeros(N, dtype=np.int64)
        self._closed = False
        self._needs_reset = True
        # optional per-sub-env wrappers of the reference that sit above the one-hot wrapper (metaworld/__init__.py:437-444)
        from .post import StepPost
        self.post = StepPost(N, recurrent_info_in_obs, normalize_reward_in_recurrent_info, reward_normalization_method, reward_alpha)
        if self.post.recurrent:     # RNNBasedMetaRLWrapper's space: unbounded float32 of obs + action + reward + done (wrappers.py:55-62)
            D = self.obs_dim + self.post.extra
            self.obs_dtype = np.float32
            self.single_observation_space = _gym.Box(np.full(D, -np.inf, np.float32), np.full(D, np.inf, np.float32), dtype=np.float32)
            self.observation_space = _gym.batch_space(self.single_observation_space, num_envs)

    # ------------------------------------------------------------------ helpers
    def _snap(self, task: Task) -> int:
        return self._snap_base + self._snap_of[id(task)]

    def _push_next(self):
        self.h_next.copy_(self.torch.from_numpy(self._next_ids))
        self.d_next.copy_(self.h_next, non_blocking=True)

    def _draw_pending(self, e):
        s = self.sub[e]
        if s.sample_tasks_on_reset:
            s.pending = s.next_task()
        else:
            s.pending = s.current_task
        self._next_ids[e] = self._snap(s.pending)

    # ------------------------------------------------------------------ VectorEnv API
    def reset(self, *, seed=None, options=None):
        """Every sub-env: (task-select wrapper) pick a task, then SawyerXYZEnv.reset."""
        N = self.num_envs
        cur = np.zeros(N, dtype=np.int32)
        self._next_ids = np.zeros(N, dtype=np.int32)
        for e, s in enumerate(self.sub):
            if s.sample_tasks_on_reset or s.current_task is None:
                s.current_task = s.next_task()
            cur[e] = self._snap(s.current_task)
        for e in range(N):
            self._draw_pending(e)
        self.d_cur.copy_(self.torch.from_numpy(cur))
        self._push_next()
        self.engine.reset(self.d_cur, self.d_obs)
        self._ep_len[:] = 0
        self._needs_reset = False
        obs = self.d_obs.cpu().numpy().astype(self.obs_dtype)
        if self.post.active:
            obs = self.post.on_reset(obs)
        return obs, {}

    def step(self, actions):
        if self._needs_reset:
            raise RuntimeError("reset() must be called before step()")
        t = self.torch
        a = np.ascontiguousarray(actions, dtype=np.float32).reshape(self.num_envs, 4)
        self.h_actions.copy_(t.from_numpy(a))
        self.d_actions.copy_(self.h_actions, non_blocking=True)
        self.engine.step(self.d_actions, self.d_obs, self.d_reward, self.d_term, self.d_trunc, self.d_info,
                         self.d_final_obs, self.d_final_info, self.d_next)
        small = t.cat([self.d_reward[:, None], self.d_info, (self.d_term + 2 * self.d_trunc).float()[:, None]], dim=1)
        self.h_small.copy_(small, non_blocking=True)
        self.h_obs.copy_(self.d_obs, non_blocking=True)
        if self.device.type == "cuda":
            t.cuda.current_stream(self.device).synchronize()
        sm = self.h_small.numpy()
        obs = self.h_obs.numpy().astype(self.obs_dtype)
        reward = sm[:, 0].astype(np.float64)
        flags = sm[:, 8].astype(np.int64)
        terminated, truncated = (flags & 1).astype(bool), (flags & 2).astype(bool)
        infos = {}
        for i, k in enumerate(INFO_KEYS):
            infos[k] = sm[:, 1 + i].astype(np.float64)
            infos["_" + k] = np.ones(self.num_envs, dtype=bool)
        self._ep_len += 1
        done = terminated | truncated
        fo = ep_r = None
        if done.any():
            fo = self.d_final_obs.cpu().numpy().astype(self.obs_dtype)
        if self.post.active:
            obs, reward, fo, ep_r = self.post.on_step(obs, a, reward, terminated, truncated, final_obs=fo)
        if done.any():
            fi = self.d_final_info.cpu().numpy()
            if ep_r is not None:
                fi = fi.copy(); fi[:, 7] = ep_r          # RecordEpisodeStatistics sits outside the reward normalisation
            final_obs = np.full(self.num_envs, None, dtype=object)
            for e in np.nonzero(done)[0]:
                final_obs[e] = fo[e]
            final_info = {}
            for i, k in enumerate(INFO_KEYS):
                final_info[k] = np.where(done, fi[:, i], 0.0)
                final_info["_" + k] = done.copy()
            final_info["episode"] = {"r": np.where(done, fi[:, 7], 0.0), "l": np.where(done, self._ep_len, 0),
                                     "t": np.zeros(self.num_envs), "_r": done.copy(), "_l": done.copy(), "_t": done.copy()}
            final_info["_episode"] = done.copy()
            infos["final_obs"], infos["_final_obs"] = final_obs, done.copy()
            infos["final_info"], infos["_final_info"] = final_info, done.copy()
            for e in np.nonzero(done)[0]:
                s = self.sub[e]
                s.current_task = s.pending
                self._draw_pending(e)
                self._ep_len[e] = 0
            self._push_next()
        return obs, reward, terminated, truncated, infos

    def step_async(self, actions):
        self._pending_actions = actions

    def step_wait(self):
        return self.step(self._pending_actions)

    # GPU-resident variants (no host synchronisation; task re-sampling on autoreset happens on the device).  They return the
    # engine's raw outputs: the optional recurrent-obs / reward-normalisation post-processing (post.py) is a numpy-path feature.
    def enable_device_sampler(self):
        first, count = [], []
        for e, s in enumerate(self.sub):
            ids = sorted(self._snap(tk) for tk in s.tasks)
            assert ids == list(range(ids[0], ids[0] + len(ids))), "device sampler needs contiguous snapshot ranges"
            first.append(ids[0]); count.append(len(ids))
        self.engine.set_goal_sets(first, count)
        self._device_sampler = True

    def reset_torch(self):
        self.reset()
        return self.d_obs

    def step_torch(self, actions):
        nxt = None if getattr(self, "_device_sampler", False) else self.d_next
        self.engine.step(actions, self.d_obs, self.d_reward, self.d_term, self.d_trunc, self.d_info, self.d_final_obs,
                         self.d_final_info, nxt)
        return self.d_obs, self.d_reward, self.d_term, self.d_trunc, self.d_info

    # attribute RPC used by metaworld/evaluation.py:48-169 and the reference tests
    def get_attr(self, name):
        if name == "terminate_on_success":
            return tuple([self.terminate_on_success] * self.num_envs)
        if name == "task_name":
            return tuple(s.task_name for s in self.sub)
        if name == "tasks":
            return tuple(s.tasks for s in self.sub)
        if name in ("_last_rand_vec", "_partially_observable"):
            key = "rand_vec" if name == "_last_rand_vec" else "partially_observable"
            return tuple(None if s.current_task is None else s.current_task.unpack()[key] for s in self.sub)
        if name == "max_path_length":
            return tuple([500] * self.num_envs)
        if name == "sample_tasks_on_reset":
            return tuple(s.sample_tasks_on_reset for s in self.sub)
        raise AttributeError(name)

    def set_attr(self, name, values):
        vals = values if isinstance(values, (list, tuple)) else [values] * self.num_envs
        if name == "terminate_on_success":
            self.call("toggle_terminate_on_success", bool(vals[0]))
        elif name == "sample_tasks_on_reset":
            for s, v in zip(self.sub, vals):
                s.sample_tasks_on_reset = bool(v)
        else:
            raise AttributeError(name)

    def call(self, name, *args, **kwargs):
        if name == "toggle_terminate_on_success":
            self.terminate_on_success = bool(args[0])
            self.engine.set_options(self.max_episode_steps, self.terminate_on_success, 0)
            return tuple([None] * self.num_envs)
        if name == "toggle_sample_tasks_on_reset":
            for s in self.sub:
                s.sample_tasks_on_reset = bool(args[0])
            if not self._needs_reset:
                for e in range(self.num_envs):
                    self._draw_pending(e)
                self._push_next()
            return tuple([None] * self.num_envs)
        if name == "sample_tasks":
            for s in self.sub:
                s.current_task = s.next_task()
            saved = [s.sample_tasks_on_reset for s in self.sub]
            for s in self.sub:
                s.sample_tasks_on_reset = False
            obs, info = self.reset()
            for s, v in zip(self.sub, saved):
                s.sample_tasks_on_reset = v
            for e in range(self.num_envs):
                self._draw_pending(e)
            self._push_next()
            return tuple((obs[e], {}) for e in range(self.num_envs))
        if name == "get_checkpoint":
            return tuple(self._checkpoint(e) for e in range(self.num_envs))
        if name == "load_checkpoint":
            for e, ck in enumerate(args[0]):
                self._load_checkpoint(e, ck)
            return tuple([None] * self.num_envs)
        return self.get_attr(name)

    def _checkpoint(self, e):
        s = self.sub[e]
        return dict(tasks=[(t.env_name, t.data) for t in s.tasks], rng_state=s.np_random.bit_generator.state,
                    current_task_idx=s.current_task_idx, sample_tasks_on_reset=s.sample_tasks_on_reset)

    def _load_checkpoint(self, e, ck):
        s = self.sub[e]
        by_data = {t.data: t for t in s.tasks}
        s.tasks = [by_data[d] for _, d in ck["tasks"]]
        s.np_random.bit_generator.state = ck["rng_state"]
        s.current_task_idx = ck["current_task_idx"]
        s.sample_tasks_on_reset = ck["sample_tasks_on_reset"]

    def close(self, **kwargs):
        if not self._closed and self._own_engine:
            self.engine.close()
        self._closed = True


def make_mt_envs(name, seed=None, num_tasks=None, num_envs=None, **kwargs):
    """``make_mt_envs`` (metaworld/__init__.py:460-513) -> MetaWorldVecEnv.  ``vector_strategy`` is accepted
    and ignored (there is one strategy: the GPU)."""
    from . import benchmarks as B

    kwargs.pop("vector_strategy", None); kwargs.pop("autoreset_mode", None)
    bench = B.make_benchmark(name, seed)
    names = bench.train_classes
    default = {"MT10": 10, "MT25": 25, "MT50": 50}.get(name, 1)
    tasks = [[t for t in bench.train_tasks if t.env_name == n] for n in names]
    return MetaWorldVecEnv(names, tasks, num_envs=num_envs, seed=seed, num_tasks=num_tasks or default, **kwargs)


def make_ml_envs(name, seed=None, meta_batch_size=20, total_tasks_per_cls=None, split="train", num_envs=None, **kwargs):
    """``make_ml_envs`` / ``_make_ml_envs_inner`` (metaworld/__init__.py:515-604)."""
    from . import benchmarks as B

    kwargs.pop("vector_strategy", None); kwargs.pop("autoreset_mode", None)
    bench = B.ML1(name, seed) if name in TASKS else B.make_benchmark(name, seed)
    classes = bench.train_classes if split == "train" else bench.test_classes
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
    return MetaWorldVecEnv(names, tasks, num_envs=num_envs, seed=seed, **kwargs)
