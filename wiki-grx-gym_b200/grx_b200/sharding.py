"""Host-side sharding arithmetic for data-parallel training: contiguous env blocks per rank, terrain types by
GLOBAL env index (legged_robot.py:1177-1180), and the reductions that make W ranks equivalent to one GPU with N_total envs:
advantage moments [sum, sum of squares, count] (base_storage.py:140-141) and the gradient / KL sums of a minibatch."""
from __future__ import annotations

import numpy as np


def env_block(rank, world_size, num_envs_total):
    """[start, stop) of the envs simulated by `rank`."""
    assert num_envs_total % world_size == 0
    n = num_envs_total // world_size
    return rank * n, (rank + 1) * n


def terrain_types_for(rank, world_size, num_envs_total, num_cols):
    a, b = env_block(rank, world_size, num_envs_total)
    return np.floor(np.arange(a, b) / (num_envs_total / num_cols)).astype(np.int64)


def moments(x):
    """[sum, sum of squares, count] of a tensor (what gae_kernel leaves in adv_moments)."""
    import torch
    x = x.double().reshape(-1)
    return torch.stack([x.sum(), (x * x).sum(), torch.tensor(float(x.numel()), dtype=torch.float64)])


def normalize_with_moments(x, m):
    """(x - mean) / (unbiased std + 1e-8) from (all-reduced) moments: normalize_adv_kernel."""
    cnt = m[2]
    mean = m[0] / cnt
    var = ((m[1] - cnt * mean * mean) / (cnt - 1.0)).clamp(min=0.0)
    return (x - mean.float()) / (var.sqrt().float() + 1e-8)


def combine_minibatch(reduce_buf_sum, world_size, nparam):
    """After the all-reduce(sum) of reduce_buf: mean gradient, global KL mean, global loss means (prep_apply_kernel)."""
    g = reduce_buf_sum[:nparam] / world_size
    tail = reduce_buf_sum[nparam:]
    cnt = tail[1]
    return g, tail[0] / cnt, tail[2] / cnt, tail[3] / cnt
