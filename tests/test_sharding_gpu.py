"""W ranks == ONE rank with W x N envs ("as if one GPU with N_total envs"), on the GPU through the C ABI.

These tests run on ONE device: the two "ranks" are two env / PPO objects with rank=r, world_size=2 on cuda:0 and the collectives are
done by hand between them (sum of the two reduce_bufs / the two advantage-moment vectors), i.e. exactly what NCCL or the NVLink
kernel deliver.  The transports themselves are covered on real GPUs by tests/test_multigpu.py.
"""
import ctypes as C

import numpy as np
import pytest
import torch

from grx_b200.config import make_cfg, make_train_cfg

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("mesh", ["plane", "heightfield"])
def test_env_shards_equal_slices_of_one_big_env(mesh):
    """Fast mode (in-kernel Philox keyed by (seed, GLOBAL env id, step)): rank r of 2 simulating envs [rN, (r+1)N) produces bit-identical
    trajectories (obs, privileged obs, rewards, resets, carried state, terrain levels) to the same global envs of a single 2N-env
    instance — terrain types, initial levels, domain randomisation and every random draw are functions of the global index."""
    from grx_b200.env import GRXVecEnv
    N2, W = 512, 2
    def cfg():
        c = make_cfg("GR1T1", N2, mesh)
        if mesh == "heightfield":
            c.terrain.num_rows, c.terrain.num_cols, c.terrain.max_init_terrain_level = 3, 4, 2
        return c
    big = GRXVecEnv(cfg(), sim_device="cuda:0")
    shards = [GRXVecEnv(cfg(), sim_device="cuda:0", rank=r, world_size=W) for r in range(W)]
    assert [s.env_id_offset for s in shards] == [0, N2 // W]
    g = torch.Generator().manual_seed(4)
    outs_b = big.reset()
    outs_s = [s.reset() for s in shards]
    steps = 40
    n_reset = 0
    for t in range(steps):
        a = (0.4 * torch.randn(N2, 10, generator=g)).cuda()
        ob, pb, rb, db, _ = big.step(a)
        n_reset += int(db.sum())
        for r, s in enumerate(shards):
            sl = slice(r * N2 // W, (r + 1) * N2 // W)
            o, p, rw, d, _ = s.step(a[sl].contiguous())
            msg = f"{mesh} step {t} rank {r}"
            assert torch.equal(o, ob[sl]), msg + " obs"
            assert torch.equal(p, pb[sl]), msg + " pri_obs"
            assert torch.equal(rw, rb[sl]), msg + " rew"
            assert torch.equal(d, db[sl]), msg + " reset"
            assert torch.equal(s.time_out_buf, big.time_out_buf[sl]), msg
    torch.cuda.synchronize()
    assert n_reset > 0                                                                # resets (and their draws) were exercised
    for r, s in enumerate(shards):
        sl = slice(r * N2 // W, (r + 1) * N2 // W)
        assert torch.equal(s.records, big.records[sl])                                # every carried quantity, bit for bit
        if mesh == "heightfield":
            assert torch.equal(s.terrain_levels, big.terrain_levels[sl]) and torch.equal(s.terrain_types, big.terrain_types[sl])


def test_two_emulated_ranks_equal_one_rank_ppo_update():
    """2 ranks x N envs == 1 rank x 2N envs for the whole PPO iteration at the registered width on the tensor-core path:
    action noise (global-id Philox), rollout storage, GAE with GLOBAL advantage moments, and 8 optimiser steps where each global
    minibatch is the union of the two ranks' minibatches.  Stated tolerance: the two sides sum the same per-row gradients in a
    different order (split-K atomics, two partial sums instead of one), so gradients agree to ~1e-6 relative; Adam turns a sign
    flip of a near-zero gradient element into <= 2 lr, hence ||dW_2rank - dW_1rank|| <= 2 % of ||dW|| per tensor, LR sequences equal."""
    from grx_b200 import _lib as L
    from grx_b200.ppo import PPO, ActorCriticMLP
    N, T, W, nmb, nep = 256, 8, 2, 4, 2
    tc = make_train_cfg()

    def make(n, off, world):
        torch.manual_seed(17)
        ac = ActorCriticMLP(39, 168, 10, **tc["policy"])
        # world_size > 1 without an initialised torch.distributed process group: the C side divides gradient / KL sums by world_size and
        # the CALLER performs the collectives through the stepwise entries (what a non-torch host does through the C ABI)
        alg = PPO(ac, device="cuda:0", seed=3, env_id_offset=off, world_size=world,
                  **dict(tc["algorithm"], num_mini_batches=nmb, num_learning_epochs=nep))
        alg.init_storage(n, T)
        return alg, ac
    big, bac = make(W * N, 0, 1)
    ranks = [make(N, r * N, W) for r in range(W)]
    init = {k: v.detach().cpu().clone() for k, v in bac.state_dict().items()}
    g = torch.Generator().manual_seed(8)
    for t in range(T):
        obs, cobs = torch.randn(W * N, 39, generator=g).cuda(), torch.randn(W * N, 168, generator=g).cuda()
        rew, dn = (0.1 * torch.randn(W * N, generator=g)).cuda(), (torch.rand(W * N, generator=g) < 0.03).cuda()
        ab = big.act(obs, cobs)                                                       # fast mode: Philox keyed by global env id
        big.process_env_step(rew, dn, {})
        for r, (alg, _) in enumerate(ranks):
            sl = slice(r * N, (r + 1) * N)
            a = alg.act(obs[sl].contiguous(), cobs[sl].contiguous())
            assert torch.allclose(a, ab[sl], atol=1e-5), "shard actions differ from the global job"
            alg.process_env_step(rew[sl].contiguous(), dn[sl].contiguous(), {})
    last = torch.randn(W * N, 168, generator=g).cuda()
    big.compute_returns(last)
    for r, (alg, _) in enumerate(ranks):                                             # local GAE, then the 3-double all-reduce by hand
        L.check(alg.lib.grx_ppo_compute_returns_local(alg._h, C.c_void_p(last[r * N:(r + 1) * N].contiguous().data_ptr()), alg._stream()))
    msum = sum(alg._moments[:3].clone() for alg, _ in ranks)
    for alg, _ in ranks:
        alg._moments[:3].copy_(msum)
        L.check(alg.lib.grx_ppo_normalize_advantages(alg._h, alg._stream()))
    torch.cuda.synchronize()
    for r, (alg, _) in enumerate(ranks):
        np.testing.assert_allclose(alg.storage.advantages.cpu().numpy(), big.storage.advantages[:, r * N:(r + 1) * N].cpu().numpy(), rtol=1e-4, atol=1e-5)
        np.testing.assert_allclose(alg.storage.returns.cpu().numpy(), big.storage.returns[:, r * N:(r + 1) * N].cpu().numpy(), rtol=1e-5, atol=1e-5)
    # per-rank permutations; the global minibatch k is the union of the ranks' minibatches k
    B = (N * T) // nmb
    perms = [torch.randperm(nmb * B, generator=g) for _ in range(W)]
    glob = []
    for k in range(nmb):
        for r in range(W):
            loc = perms[r][k * B:(k + 1) * B]
            tt, nn = loc // N, loc % N
            glob.append(tt * (W * N) + r * N + nn)
    big._indices.copy_(torch.cat(glob).cuda())
    for r, (alg, _) in enumerate(ranks):
        alg._indices.copy_(perms[r].cuda())
    lr_big, lr_rk = [], []
    for ep in range(nep):
        for mb in range(nmb):
            L.check(big.lib.grx_ppo_minibatch_grads(big._h, C.c_void_p(big._indices.data_ptr()), mb, big._stream()))
            L.check(big.lib.grx_ppo_minibatch_apply(big._h, big._stream()))
            for alg, _ in ranks:
                L.check(alg.lib.grx_ppo_minibatch_grads(alg._h, C.c_void_p(alg._indices.data_ptr()), mb, alg._stream()))
            tot = sum(alg.reduce_buf.clone() for alg, _ in ranks)                     # the ONE all-reduce(sum) per minibatch
            if ep == 0 and mb == 0:
                gb = big.reduce_buf.clone()
                np.testing.assert_allclose((tot[:-8] / W).cpu().numpy(), gb[:-8].cpu().numpy(), rtol=0, atol=2e-5 * float(gb[:-8].abs().max()))
                np.testing.assert_allclose(tot[-8:-4].cpu().numpy(), gb[-8:-4].cpu().numpy(), rtol=1e-4)   # KL / count / loss sums
            for alg, _ in ranks:
                alg.reduce_buf.copy_(tot)
                L.check(alg.lib.grx_ppo_minibatch_apply(alg._h, alg._stream()))
            lr_big.append(big.minibatch_stats()["lr"]); lr_rk.append(ranks[0][0].minibatch_stats()["lr"])
    torch.cuda.synchronize()
    assert lr_big == pytest.approx(lr_rk, rel=1e-6)
    assert torch.equal(ranks[0][0].params, ranks[1][0].params)                        # replicas stay bit-identical
    assert ranks[0][0].adam_step == big.adam_step == nmb * nep
    worst = 0.0
    for key, v in bac.state_dict().items():
        w0 = init[key].numpy()
        d1, d2 = v.cpu().numpy() - w0, ranks[0][1].state_dict()[key].cpu().numpy() - w0
        rel = np.linalg.norm(d2 - d1) / (np.linalg.norm(d1) + 1e-30)
        worst = max(worst, rel)
        assert rel < 0.02, f"{key}: 2-rank update differs from the 1-rank update by {rel:.4f} of its norm"
    print(f"2 emulated ranks vs 1 rank: worst per-tensor relative update difference {worst:.5f}")
    for alg, _ in ranks + [(big, None)]:
        alg.close()
