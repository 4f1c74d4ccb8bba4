"""Multi-GPU layout: environments shard embarrassingly, one process per GPU, no collective on the step path.

The reference's only parallelism is env data-parallelism through ``SyncVectorEnv`` / ``AsyncVectorEnv``
(metaworld/__init__.py:481-509).  Here rank r of W owns a contiguous block of the global interleaved-by-task
ordering (env e has type e % n_types), so every rank sees every task type and the per-rank work is balanced.
The optional epilogue collective concatenates per-rank observations on rank 0 (`gather_to_rank0`) with
``torch.distributed.all_gather_into_tensor`` / ``gather`` (NCCL over NVLink on GPUs, gloo on CPU in the tests).
"""
from __future__ import annotations

import numpy as np


def shard_env_ids(num_envs_total: int, rank: int, world: int) -> np.ndarray:
    """Global env ids owned by `rank`: a contiguous block (consecutive ids cycle through the task types)."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    base, rem = divmod(num_envs_total, world)
    start = rank * base + min(rank, rem)
    return np.arange(start, start + base + (1 if rank < rem else 0), dtype=np.int64)


def env_type(global_env_id, n_types: int):
    return np.asarray(global_env_id) % n_types


def gather_to_rank0(local, num_envs_total: int, rank: int, world: int, group=None):
    """Gathers per-rank rows (torch tensor [n_local, ...]) into global env order on rank 0; returns None elsewhere.
    Uneven shards are padded to the largest shard for the collective."""
    import torch
    import torch.distributed as dist

    if world == 1:
        return local
    n_max = (num_envs_total + world - 1) // world
    pad = torch.zeros((n_max,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    out = torch.empty((world * n_max,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, pad, group=group)
    if rank != 0:
        return None
    if num_envs_total == world * n_max:
        return out                      # even shards: the gathered buffer already is the global env order (contiguous blocks)
    # uneven shards: one index op drops the padding rows (shard r's rows sit at [r * n_max, r * n_max + len_r))
    base, rem = divmod(num_envs_total, world)
    g = torch.arange(num_envs_total, device=local.device)
    r_of = torch.where(g < rem * (base + 1), g // (base + 1), rem + (g - rem * (base + 1)) // max(base, 1))
    start = r_of * base + torch.clamp(r_of, max=rem)
    return out.index_select(0, r_of * n_max + (g - start))
