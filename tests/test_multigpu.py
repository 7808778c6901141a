"""Two B200s, one process per GPU (skipped with fewer than 2 GPUs — the driver's 1-GPU box; run with `gpurun --gpus 2`, log kept
under profiles/).  Covers the TRANSPORTS; the W-rank == 1-rank arithmetic is also covered on one GPU by tests/test_sharding_gpu.py.

 * the per-minibatch gradient all-reduce over NVLink peer memory (allreduce_kernel inside the update's CUDA graph) gives the same
   update as the NCCL all-reduce path, replicas bit-identical, no flag time-out;
 * 2 ranks x N envs (NVLink graph update, global-id action noise, all-reduced advantage moments) == 1 rank x 2N envs.

Tolerance (stated, and measured in the same test): the backward accumulates split-K partial sums with red.global.add / TMA reduce-add,
whose order changes from run to run, so even the SAME transport differs from itself by a relative ~1e-7 in the gradients; Adam's
g / (sqrt(v) + eps) turns that into up to 2 lr per step on elements whose gradient is ~0.  The test therefore measures the
run-to-run floor (NVLink path twice) and requires the cross-transport / cross-sharding difference of the UPDATE (w_final - w_init) to
stay within max(4 x floor, 1 % of the update norm) per tensor; LR sequence, Adam step count and mean KL must agree.
"""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

N, T, NMB, NEP = 256, 16, 4, 2


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _job_inputs(W):
    """Inputs of the GLOBAL job (W * N envs), generated identically in every process."""
    g = torch.Generator().manual_seed(100)
    steps = []
    for _ in range(T):
        steps.append((torch.randn(W * N, 39, generator=g), torch.randn(W * N, 168, generator=g), 0.1 * torch.randn(W * N, generator=g),
                      torch.rand(W * N, generator=g) < 0.02))
    last = torch.randn(W * N, 168, generator=g)
    B = (N * T) // NMB
    perms = [torch.randperm(NMB * B, generator=g) for _ in range(W)]
    return steps, last, perms, B


def _make(n, off, world, dev):
    from grx_b200.config import make_train_cfg
    from grx_b200.ppo import PPO, ActorCriticMLP
    tc = make_train_cfg()
    torch.manual_seed(5)
    ac = ActorCriticMLP(39, 168, 10, **tc["policy"])
    alg = PPO(ac, device=dev, world_size=world, seed=9, env_id_offset=off,
              **dict(tc["algorithm"], num_mini_batches=NMB, num_learning_epochs=NEP))
    alg.init_storage(n, T)
    return alg, ac


def _iteration(alg, steps, last, idx, sl, dev):
    for obs, cobs, rew, dn in steps:
        alg.act(obs[sl].to(dev), cobs[sl].to(dev))                   # fast mode: Philox keyed by (seed, GLOBAL env id, step)
        alg.process_env_step(rew[sl].to(dev), dn[sl].to(dev), {})
    alg.compute_returns(last[sl].to(dev))
    alg.update(indices=idx)
    torch.cuda.synchronize()
    alg.check_comm(wait=True)


def _worker(rank, world, port, comm, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    os.environ["GRX_COMM"] = comm
    torch.cuda.set_device(rank)
    dev = f"cuda:{rank}"
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device(dev))
    steps, last, perms, B = _job_inputs(world)
    alg, ac = _make(N, rank * N, world, dev)
    init = alg.params.cpu().numpy().copy()
    _iteration(alg, steps, last, perms[rank], slice(rank * N, (rank + 1) * N), dev)
    out = dict(rank=rank, comm=comm, used_comm=bool(alg._comm), init=init, params=alg.params.cpu().numpy(), adam_m=alg.adam_m.cpu().numpy(),
               stats=alg.minibatch_stats(), log=alg.mb_log.cpu().numpy(), actions=alg.storage.actions.cpu().numpy(),
               adv=alg.storage.advantages.cpu().numpy())
    alg.close()
    if rank == 0 and comm == "nvlink":   # the same GLOBAL job on one GPU: minibatch k = union of the ranks' minibatches k
        big, bac = _make(world * N, 0, 1, dev)
        glob = []
        for k in range(NMB):
            for r in range(world):
                loc = perms[r][k * B:(k + 1) * B]
                glob.append((loc // N) * (world * N) + r * N + loc % N)
        _iteration(big, steps, last, torch.cat(glob), slice(0, world * N), dev)
        out["big"] = dict(params=big.params.cpu().numpy(), log=big.mb_log.cpu().numpy(), actions=big.storage.actions.cpu().numpy(),
                          adv=big.storage.advantages.cpu().numpy(), step=big.adam_step)
        big.close()
    q.put(out)
    dist.barrier()
    dist.destroy_process_group()


def _run(world, comm):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q, port = ctx.Queue(), _free_port()
    ps = [ctx.Process(target=_worker, args=(r, world, port, comm, q)) for r in range(world)]
    for p in ps:
        p.start()
    res = sorted((q.get(timeout=600) for _ in range(world)), key=lambda r: r["rank"])
    for p in ps:
        p.join(60)
    return res


def _rel(a, b, init):
    """||(a - init) - (b - init)|| / ||b - init||: difference of two updates relative to the update."""
    return float(np.linalg.norm(a - b) / (np.linalg.norm(b - init) + 1e-30))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_gpus_nvlink_nccl_and_one_rank_equivalence():
    nv1, nv2, nc = _run(2, "nvlink"), _run(2, "nvlink"), _run(2, "nccl")
    assert all(r["used_comm"] for r in nv1 + nv2) and not any(r["used_comm"] for r in nc)
    assert all(r["stats"]["comm_error"] == 0 for r in nv1 + nv2 + nc), "a peer flag wait timed out"
    # replicas hold bit-identical parameters after the update (replicated optimiser on identical summed gradients)
    for res in (nv1, nv2, nc):
        np.testing.assert_array_equal(res[0]["params"], res[1]["params"])
        np.testing.assert_array_equal(res[0]["adam_m"], res[1]["adam_m"])
    init = nv1[0]["init"]
    floor = _rel(nv1[0]["params"], nv2[0]["params"], init)                             # the same transport, twice
    cross = _rel(nv1[0]["params"], nc[0]["params"], init)
    big = nv1[0]["big"]
    shard = _rel(nv1[0]["params"], big["params"], init)
    print(f"2-GPU update differences relative to the update norm: run-to-run floor {floor:.2e}, NVLink vs NCCL {cross:.2e}, "
          f"2 ranks vs 1 rank {shard:.2e}")
    bound = max(4.0 * floor, 1e-2)
    assert cross <= bound, (cross, floor)
    assert shard <= bound, (shard, floor)
    # LR sequence / KL per minibatch / step count
    for other in (nv2[0]["log"], nc[0]["log"], big["log"]):
        np.testing.assert_allclose(other[:, 1], nv1[0]["log"][:, 1], rtol=1e-6)        # learning-rate sequence
        np.testing.assert_allclose(other[:, 0], nv1[0]["log"][:, 0], rtol=5e-3, atol=1e-6)   # mean KL per minibatch
    assert nv1[0]["stats"]["step"] == nc[0]["stats"]["step"] == big["step"] == NMB * NEP
    # rollout: the shards' action noise and the globally normalised advantages are the global job's
    for r in range(2):
        np.testing.assert_allclose(nv1[r]["actions"], big["actions"][:, r * N:(r + 1) * N], rtol=0, atol=1e-5)
        np.testing.assert_allclose(nv1[r]["adv"], big["adv"][:, r * N:(r + 1) * N], rtol=1e-4, atol=1e-5)
    assert not np.allclose(nv1[0]["actions"], nv1[1]["actions"])                       # ranks do not duplicate each other's noise
