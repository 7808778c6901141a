"""N > 1 host logic on CPU: two gloo ranks reproduce the single-process result for the env partition, the global advantage
normalisation and the per-minibatch gradient / KL reduction (the three places ranks communicate)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from grx_b200 import sharding as sh


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(0)
    N, T, P = 64, 8, 37
    adv_all = torch.randn(T, N, generator=g)                       # raw advantages of the whole job
    a, b = sh.env_block(rank, world, N)
    local = adv_all[:, a:b]
    m = sh.moments(local)
    dist.all_reduce(m)                                              # the one 3-double all-reduce per iteration
    norm_local = sh.normalize_with_moments(local, m)
    ref = (adv_all - adv_all.mean()) / (adv_all.std() + 1e-8)      # base_storage.py:141 on one process
    ok_adv = torch.allclose(norm_local, ref[:, a:b], atol=1e-6)
    # minibatch: per-rank gradient of a mean over local rows + tail sums -> averaged gradient == gradient of the global mean
    x = torch.randn(N, P, generator=g)
    w = torch.randn(P, generator=g)
    rows = x[a:b]
    grad_local = (2 * (rows @ w)).unsqueeze(1).mul(rows).mean(0)   # d/dw mean((x w)^2) over local rows
    kl_local = (rows @ w).abs()
    buf = torch.cat([grad_local, torch.tensor([kl_local.sum(), float(rows.shape[0]), 0.0, 0.0])])
    dist.all_reduce(buf)                                            # the ONE all-reduce per minibatch
    gmean, kl_mean, _, _ = sh.combine_minibatch(buf, world, P)
    grad_ref = (2 * (x @ w)).unsqueeze(1).mul(x).mean(0)
    ok_grad = torch.allclose(gmean, grad_ref, atol=1e-5) and abs(float(kl_mean) - float((x @ w).abs().mean())) < 1e-6
    types = sh.terrain_types_for(rank, world, N, 4)
    q.put((rank, bool(ok_adv), bool(ok_grad), types.tolist()))
    dist.destroy_process_group()


def test_two_ranks_equal_one_process():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in ps:
        p.join(30)
    assert all(r[1] and r[2] for r in res), res
    # terrain columns stay balanced and contiguous across the rank boundary (global index, legged_robot.py:1177-1180)
    assert res[0][3] + res[1][3] == np.floor(np.arange(64) / 16).astype(int).tolist()


def test_env_block_partition():
    for W in (1, 2, 4, 8):
        blocks = [sh.env_block(r, W, 32768) for r in range(W)]
        assert blocks[0][0] == 0 and blocks[-1][1] == 32768
        assert all(blocks[i][1] == blocks[i + 1][0] for i in range(W - 1))
        assert all(b - a == 32768 // W for a, b in blocks)
