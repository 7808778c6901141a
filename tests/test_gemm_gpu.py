"""tcgen05 TF32 dense-layer kernel vs the fp32 SIMT kernel and torch fp32, through the C ABI (grx_gemm_debug):
forward (K-major x K-major), input gradient (K-major x MN-major), weight gradient (MN-major x MN-major, split-K)."""
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

# TF32 rounds the operands to 10 mantissa bits (rel 2^-11 each); fp32 accumulate.  Error of a length-K dot product of O(1)
# terms grows ~ sqrt(K) * 2^-11: tolerance relative to the output scale.
TF32_RTOL = 2e-3


def _run(variant, epi, M, N, K, use_tc, splits=1, seed=0):
    from grx_b200 import _lib as L
    lib = L.lib()
    g = torch.Generator(device="cuda").manual_seed(seed)
    r = lambda *s: torch.randn(*s, device="cuda", generator=g)
    if variant == 0:
        A, B = r(M, K), r(N, K) / K ** 0.5
        bias, aux, bias_out = r(N), None, None
        C_ = torch.zeros(M, N, device="cuda")
        ref = A @ B.t() + bias
        if epi == 1:
            ref = torch.nn.functional.elu(ref)
    elif variant == 1:
        A, B = r(M, K), r(K, N) / K ** 0.5
        bias, bias_out = None, None
        aux = torch.nn.functional.elu(r(M, N))
        C_ = torch.zeros(M, N, device="cuda")
        ref = (A @ B) * torch.where(aux > 0, torch.ones_like(aux), aux + 1)
    else:
        A, B = r(K, M), r(K, N) / K ** 0.5
        bias, aux = None, None
        bias_out = torch.zeros(M, device="cuda")
        C_ = torch.zeros(M, N, device="cuda")
        ref = A.t() @ B
    p = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None
    L.check(lib.grx_gemm_debug(variant, epi, M, N, K, p(A), p(B), p(C_), p(bias), p(aux), p(bias_out), splits, use_tc,
                               C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    torch.cuda.synchronize()
    extra = (bias_out, A.sum(0)) if variant == 2 else None
    return C_, ref, extra


SHAPES_FWD = [(10485, 512, 168), (10485, 256, 512), (10485, 128, 256), (4096, 512, 168), (300, 64, 64), (129, 48, 40), (128, 16, 8)]


@pytest.mark.parametrize("M,N,K", SHAPES_FWD)
@pytest.mark.parametrize("epi", [0, 1])
def test_forward_layer(M, N, K, epi):
    got, ref, _ = _run(0, epi, M, N, K, use_tc=1)
    simt, _, _ = _run(0, epi, M, N, K, use_tc=0)
    scale = float(ref.abs().max())
    assert float((simt - ref).abs().max()) <= 2e-5 * scale + 1e-5
    assert float((got - ref).abs().max()) <= TF32_RTOL * scale, float((got - ref).abs().max()) / scale


@pytest.mark.parametrize("M,N,K", [(10485, 512, 256), (10485, 256, 128), (777, 168, 512), (128, 64, 64)])
def test_input_gradient_layer(M, N, K):
    got, ref, _ = _run(1, 2, M, N, K, use_tc=1)
    scale = float(ref.abs().max())
    assert float((got - ref).abs().max()) <= TF32_RTOL * scale, float((got - ref).abs().max()) / scale


@pytest.mark.parametrize("M,N,K,splits", [(512, 168, 10485, 37), (256, 512, 10485, 20), (128, 256, 10485, 74), (128, 64, 333, 3), (512, 168, 64, 1)])
def test_weight_gradient_layer(M, N, K, splits):
    got, ref, (bo, bo_ref) = _run(2, 3, M, N, K, use_tc=1, splits=splits)
    scale = float(ref.abs().max())
    assert float((got - ref).abs().max()) <= TF32_RTOL * scale, float((got - ref).abs().max()) / scale
    np.testing.assert_allclose(bo.cpu().numpy(), bo_ref.cpu().numpy(), rtol=1e-4, atol=1e-3)


def test_tensor_core_update_matches_simt_update():
    """Whole PPO minibatch at the registered width through the tcgen05 path vs the fp32 SIMT path: gradients agree to TF32 accuracy."""
    from grx_b200.config import make_train_cfg
    from grx_b200.ppo import PPO, ActorCriticMLP
    from grx_b200 import _lib as L
    tc = make_train_cfg()
    O, P, A, N, T = 39, 168, 10, 512, 8
    grads = []
    for use_tc in (0, 1):
        torch.manual_seed(5)
        ac = ActorCriticMLP(O, P, A, **tc["policy"])
        alg = PPO(ac, device="cuda:0", use_tensor_cores=use_tc, **dict(tc["algorithm"], num_mini_batches=2, num_learning_epochs=1))
        alg.init_storage(N, T)
        g = torch.Generator().manual_seed(7)
        for s in range(T):
            alg.act(torch.randn(N, O, generator=g).cuda(), torch.randn(N, P, generator=g).cuda(), eps=torch.randn(N, A, generator=g).cuda())
            alg.process_env_step((0.1 * torch.randn(N, generator=g)).cuda(), (torch.rand(N, generator=g) < 0.1).cuda(), {})
        alg.compute_returns(torch.randn(N, P, generator=g).cuda())
        with torch.no_grad():
            for k, v in ac.state_dict().items():
                v.add_(0.02 * torch.randn(v.shape, generator=g).cuda() * (v.abs().mean() + 0.05))
        alg._indices.copy_(torch.randperm(N * T, generator=g).cuda())
        L.check(alg.lib.grx_ppo_minibatch_grads(alg._h, C.c_void_p(alg._indices.data_ptr()), 0, alg._stream()))
        torch.cuda.synchronize()
        grads.append({k: ac.view_of(alg.grads, k).clone() for k in ac.state_dict()})
    g0, g1 = grads
    worst = {}
    for k in g0:
        a, b = g0[k], g1[k]
        worst[k] = float((a - b).abs().max()) / (float(a.abs().max()) + 1e-12)
    # three TF32 layers forward and three backward: operand rounding 2^-11 per product, accumulated over both passes
    assert max(worst.values()) <= 1e-2, worst


@pytest.mark.parametrize("tile", [(1, 32), (1, 64), (1, 128), (1, 256), (2, 128), (2, 256)])
def test_every_macro_tile_configuration(tile):
    """The persistent kernel's macro tiles (row blocks x BN) are normally picked by a cost model; force each one and check the
    forward (bias + ELU, TMA store, ragged M / N), input-gradient (ELU', column sums) and split-K weight-gradient (TMA reduce-add)
    paths against fp32 torch."""
    from grx_b200 import _lib as L
    lib = L.lib()
    L.check(lib.grx_gemm_debug_tile(tile[0], tile[1]))
    try:
        for variant, epi, M, N, K, splits in [(0, 1, 1000, 328, 168, 1), (0, 0, 10485, 512, 40, 1), (1, 2, 777, 256, 128, 1),
                                              (2, 3, 256, 512, 3000, 7), (2, 3, 512, 168, 2048, 4)]:
            got, ref, _ = _run(variant, epi, M, N, K, use_tc=1, splits=splits, seed=3)
            scale = float(ref.abs().max())
            assert float((got - ref).abs().max()) <= TF32_RTOL * scale, (tile, variant, M, N, K, float((got - ref).abs().max()) / scale)
    finally:
        L.check(lib.grx_gemm_debug_tile(0, 0))


@pytest.mark.parametrize("M,store", [(4096, False), (10485, True), (333, True), (128, False)])
def test_chained_forward_equals_layerwise(M, store):
    """The three hidden layers as ONE chained tcgen05 kernel (csrc/grx_mlp_chain.cuh: activations handed from the epilogue warps to the
    MMA warp as shared-memory A operands) == the same layers as three grouped GEMM launches (same K order on the tensor core, same bias /
    ELU arithmetic: bit-identical is expected, 1e-6 relative allowed), and == torch fp32 within the TF32 tolerance.  store=False is the
    rollout variant (only the last hidden layer is written to HBM)."""
    from grx_b200 import _lib as L
    from grx_b200.config import make_train_cfg
    from grx_b200.ppo import PPO, ActorCriticMLP
    lib = L.lib()
    tc = make_train_cfg()
    torch.manual_seed(31)
    N, T = (M, 4) if not store else (M * 4 // 4, 4)           # update path: minibatch of N*T/4 = M rows
    ac = ActorCriticMLP(39, 168, 10, **tc["policy"])
    alg = PPO(ac, device="cuda:0", **dict(tc["algorithm"], num_mini_batches=4, num_learning_epochs=1))
    alg.init_storage(N, T)
    g = torch.Generator().manual_seed(M)
    obs, cobs, eps = torch.randn(N, 39, generator=g).cuda(), torch.randn(N, 168, generator=g).cuda(), torch.randn(N, 10, generator=g).cuda()
    adv, idx = torch.randn(T, N, 1, generator=g).cuda(), torch.randperm(N * T, generator=g).cuda()
    res = {}
    for fused in (1, 0):   # chained (A operands of layers 1 and 2 in tensor memory) / layerwise
        old = lib.grx_ppo_debug_fused(fused)
        alg.step = 0
        if not store:   # rollout: PPO.act
            a = alg.act(obs, cobs, eps=eps).clone()
            res[fused] = (a, alg.storage.mu[0].clone(), alg.storage.values[0].clone(), alg.storage.actions_log_prob[0].clone())
        else:           # update: one minibatch's forward + backward through the stepwise entry
            for t in range(T):
                alg.act(obs, cobs, eps=eps)
                alg.process_env_step(torch.zeros(N, device="cuda"), torch.zeros(N, dtype=torch.bool, device="cuda"), {})
            alg.compute_returns(cobs)
            alg.storage.advantages.copy_(adv)
            alg._indices.copy_(idx)
            import ctypes as C
            L.check(lib.grx_ppo_minibatch_grads(alg._h, C.c_void_p(alg._indices.data_ptr()), 2, alg._stream()))
            res[fused] = (alg.grads.clone(), alg.reduce_buf[-8:].clone())
        torch.cuda.synchronize()
        assert alg.minibatch_stats()["chain_error"] == 0, "a barrier wait inside the chained kernel timed out"
        lib.grx_ppo_debug_fused(old)
    for variant in (1,):
        for x, y in zip(res[variant], res[0]):
            scale = float(y.abs().max()) + 1e-30
            assert float((x - y).abs().max()) <= 1e-5 * scale, (variant, M, store, float((x - y).abs().max()), scale)
    if not store:   # and against plain torch fp32
        with torch.no_grad():
            mu = ac.actor.to_module()(obs.cpu())
        np.testing.assert_allclose(res[1][1].cpu().numpy(), mu.numpy(), rtol=0, atol=6e-3 * float(mu.abs().max()))
    alg.close()


@pytest.mark.parametrize("M,store", [(4096, False), (10485, True), (333, True), (128, False), (700, False)])
def test_layer_pipelined_launch_equals_layerwise(M, store):
    """Dependent dense layers in ONE persistent launch (tc::launch_pipe: the three hidden layers of both networks of the forward pass, the two
    wide input-gradient layers of the backward pass; a tile's loads wait on the row-block counter of the layer that produces its A operand)
    == the same layers as one grouped launch per layer.  Same contraction order per output element, so the forward is expected bit-identical
    (1e-6 relative allowed); the gradients additionally carry the run-to-run noise of the split-K / column-sum atomics (1e-5 of the scale,
    as for the chained kernel above).  store=False is the rollout (PPO.act), store=True one minibatch's forward + backward."""
    import ctypes as C
    from grx_b200 import _lib as L
    from grx_b200.config import make_train_cfg
    from grx_b200.ppo import PPO, ActorCriticMLP
    lib = L.lib()
    lib.grx_ppo_debug_pipe.argtypes = [C.c_int32]
    tc = make_train_cfg()
    torch.manual_seed(37)
    N, T = M, 4
    ac = ActorCriticMLP(39, 168, 10, **tc["policy"])
    alg = PPO(ac, device="cuda:0", **dict(tc["algorithm"], num_mini_batches=4, num_learning_epochs=1))
    alg.init_storage(N, T)
    g = torch.Generator().manual_seed(M + 1)
    obs, cobs, eps = torch.randn(N, 39, generator=g).cuda(), torch.randn(N, 168, generator=g).cuda(), torch.randn(N, 10, generator=g).cuda()
    adv, idx = torch.randn(T, N, 1, generator=g).cuda(), torch.randperm(N * T, generator=g).cuda()
    res = {}
    old_fused = lib.grx_ppo_debug_fused(0)
    try:
        for pipe in (3, 0, 3):   # 3 = forward AND input-gradient chains pipelined
            old = lib.grx_ppo_debug_pipe(pipe)
            alg.step = 0
            if not store:
                a = alg.act(obs, cobs, eps=eps).clone()
                out = (a, alg.storage.mu[0].clone(), alg.storage.values[0].clone(), alg.storage.actions_log_prob[0].clone())
            else:
                for t in range(T):
                    alg.act(obs, cobs, eps=eps)
                    alg.process_env_step(torch.zeros(N, device="cuda"), torch.zeros(N, dtype=torch.bool, device="cuda"), {})
                alg.compute_returns(cobs)
                alg.storage.advantages.copy_(adv)
                alg._indices.copy_(idx)
                L.check(lib.grx_ppo_minibatch_grads(alg._h, C.c_void_p(alg._indices.data_ptr()), 2, alg._stream()))
                out = (alg.grads.clone(), alg.reduce_buf[-8:].clone())
            torch.cuda.synchronize()
            assert alg.minibatch_stats()["chain_error"] == 0, "a dependency wait inside the pipelined launch timed out"
            res.setdefault(pipe, []).append(out)
            lib.grx_ppo_debug_pipe(old)
    finally:
        lib.grx_ppo_debug_fused(old_fused)
    tol = 1e-6 if not store else 1e-5
    for run in res[3]:   # both pipelined runs (the second one starts from the counters the first one left behind) vs the layerwise run
        for x, y in zip(run, res[0][0]):
            scale = float(y.abs().max()) + 1e-30
            assert float((x - y).abs().max()) <= tol * scale, (M, store, float((x - y).abs().max()), scale)
    if not store:
        with torch.no_grad():
            mu = ac.actor.to_module()(obs.cpu())
        np.testing.assert_allclose(res[3][0][1].cpu().numpy(), mu.numpy(), rtol=0, atol=6e-3 * float(mu.abs().max()))
    alg.close()
