"""Shared by the GPU parity tests and __graft_entry__.smoke(): build the CUDA env from a golden fixture and compare."""
import numpy as np

from golden_util import init_state, load_fixture, phys_oracle, setup_from_fixture, step_items  # noqa: F401

# post-physics map (the reference's own arithmetic): fp32 on both sides, different exp/sqrt implementations and summation trees
POST_TOL = dict(rtol=1e-5, atol=3e-6)
# full step incl. 10 substeps of our dynamics spec: CUDA fp32 (Delassus-space PGS, branch-sparse Cholesky, fused multiply-add)
# vs the C oracle in fp32 (velocity-space PGS, dense Cholesky, no contraction) — same equations, different rounding.
# MEASURED on B200 (gpurun_out/r2b_pytest.log, envs with identical active sets only):
#   one 2 ms substep (plane64_dec1):            positions 5e-6, joint angles 3e-7, joint rates 1.7e-4
#   one policy step = 10 substeps, plane:       positions 5e-5, joint angles 6e-6, joint rates 4e-3, contact forces 0.23 N
#   10 substeps, heightfield / trimesh:         positions 2e-4, joint angles 7e-5, joint rates 3-5e-2, contact forces 0.5-3.2 N
# Velocity-level quantities carry the error: a contact row targets v_n = -d / dt (dt = 2 ms), i.e. position rounding is amplified by
# 500 1/s per substep, and on the rough terrains the robots stand 25-60 m from the origin where one fp32 ulp of a world coordinate is
# 4-8e-6 m (=> 2-4e-3 m/s per contact row per substep, x lever arms of 0.1-0.4 m for joint rates).  PhysX carries state in fp32 too.
PHYS_TOL = dict(rtol=2e-3, atol=2e-3)           # positions, orientations, joint angles, rewards, non-velocity observation columns
PHYS_VEL_TOL = dict(rtol=5e-3, atol=8e-2)       # joint rates (rad/s, range +-20) and base velocities incl. their observation columns
# PD torques of the LAST substep = Kp (target - q) - Kd qd with Kp up to 251 N m/rad, Kd up to 14.7 (gr1t1_lower_limb_config.py): the state
# tolerance propagated through the gains; torque limits are 48-130 N m
PHYS_TORQUE_TOL = dict(rtol=2e-3, atol=0.3)
PHYS_FORCE_TOL = dict(rtol=2e-2, atol=5.0)      # net contact forces (N; body weight 518 N): impulse / dt, i.e. mass x velocity error / 2 ms
OBS_VEL_COLS = list(range(3, 6)) + list(range(19, 29))          # base angular velocity, joint rates
PRI_VEL_COLS = OBS_VEL_COLS + list(range(39, 42))               # + base linear velocity


def vel_cols(nd):
    """(obs, pri_obs) columns that carry velocities for a robot with nd DOF (obs = cmd 3, ang vel 3, gravity 3, q nd, qd nd, actions nd)."""
    o = list(range(3, 6)) + list(range(9 + nd, 9 + 2 * nd))
    return o, o + list(range(9 + 3 * nd, 12 + 3 * nd))


assert vel_cols(10) == (OBS_VEL_COLS, PRI_VEL_COLS)


def make_gpu_env(fx, generic=False, **kw):
    """generic=True routes the (lower-limb) model to the generic-topology kernels (GRX_ENV_GENERIC=1, read by grx_env_create); full-body
    fixtures always run there."""
    import os
    import torch
    from grx_b200.env import GRXVecEnv
    old_env = os.environ.get("GRX_ENV_GENERIC")
    os.environ["GRX_ENV_GENERIC"] = "1" if generic else "0"
    try:
        return _make_gpu_env(fx, torch, GRXVecEnv, **kw)
    finally:
        if old_env is None:
            del os.environ["GRX_ENV_GENERIC"]
        else:
            os.environ["GRX_ENV_GENERIC"] = old_env


def _make_gpu_env(fx, torch, GRXVecEnv, **kw):
    cfg, model, tables, consts, terrain = setup_from_fixture(fx)
    cfg.env.num_envs = len(consts["friction"])
    if model["nd"] <= 10:
        kw.setdefault("self_collision", False)   # the lower-limb fixtures' physics (the C oracle) ran without self-contact rows
    params = dict(friction=consts["friction"], restitution=consts["restitution"], motor_strength=consts["motor_strength"],
                  base_inertial=consts["base_inertial"])
    tkw = {}
    if terrain is not None:
        tkw = dict(terrain=dict(heights=terrain["heights"], terrain_origins=consts["terrain_origins"]),
                   terrain_levels=consts["terrain_levels"], terrain_types=consts["terrain_types"])
    env = GRXVecEnv(cfg, sim_device="cuda:0", params=params, env_origins=consts["env_origins"], **tkw, **kw)
    env.load_state(init_state(fx))
    torch.cuda.synchronize()
    return env, cfg, model, tables, consts, terrain


def make_oracle_env(fx):
    from oracle.env_oracle import EnvOracle
    cfg, model, tables, consts, terrain = setup_from_fixture(fx)
    phys = phys_oracle(cfg, model, tables, terrain)
    env = EnvOracle(cfg, tables, consts, phys, terrain)
    env.load_state(init_state(fx))
    return env


def smoke_check(np_, torch):
    """One policy step of 64 robots on cuda:0 through the C ABI, checked against the CPU oracle (env + physics)."""
    fx = load_fixture("plane64_dec1")
    env = make_gpu_env(fx)[0]
    ora = make_oracle_env(fx)
    for t in range(2):
        pre = f"step{t:02d}/"
        a, U, delay = fx[pre + "actions"], fx[pre + "U"], float(fx[pre + "delay"])
        obs, pri, rew, reset, _ = env.step(torch.from_numpy(a).cuda(), U=torch.from_numpy(U).cuda(), delay=delay)
        torch.cuda.synchronize()
        o_obs, o_pri, o_rew, o_reset, _ = ora.step(a, U, delay)
        np_.testing.assert_array_equal(reset.cpu().numpy(), o_reset.numpy())
        np_.testing.assert_allclose(obs.cpu().numpy(), o_obs.numpy(), **PHYS_TOL)
        np_.testing.assert_allclose(pri.cpu().numpy(), o_pri.numpy(), **PHYS_TOL)
        np_.testing.assert_allclose(rew.cpu().numpy(), o_rew.numpy(), **PHYS_TOL)
    print("smoke: env step on cuda:0 matches the CPU oracle (64 envs, 2 steps)")
    ppo_gradient_check(np_, torch, use_tc=1, atol=6e-3)
    print("smoke: PPO minibatch (tcgen05 TF32 layers, fused heads, backward) on cuda:0 matches the CPU oracle")


def tf32_emulated_grads(torch, p, b, clip, vcoef, ecoef):
    """The minibatch's gradients with every hidden-layer GEMM operand (forward, input gradient, weight gradient) TRUNCATED to TF32's 10-bit
    mantissa — what tcgen05 kind::tf32 does to fp32 operands (tools/tf32_update_error.py: truncation reproduces the measured update error) —
    fp32 accumulate, output heads in fp32: the arithmetic the tensor-core path performs, on the CPU through autograd.  Separates the error that is
    inherent to TF32 from implementation error where a gradient is ill-conditioned (the std gradient at 32 actions: large cancelling terms)."""
    def tf32(x):
        return (x.contiguous().view(torch.int32) & ~0x1FFF).view(torch.float32)

    class MM(torch.autograd.Function):
        @staticmethod
        def forward(ctx, a, w):          # a [M, K], w [N, K] -> a w^T
            ctx.save_for_backward(a, w)
            return tf32(a) @ tf32(w).t()

        @staticmethod
        def backward(ctx, g):
            a, w = ctx.saved_tensors
            return tf32(g) @ tf32(w), tf32(g).t() @ tf32(a)
    pa = {k: v.clone().requires_grad_(True) for k, v in p.items()}

    def mlp(net, x):
        for i in range(4):
            w, bias = pa[f"{net}.model.{2 * i}.weight"], pa[f"{net}.model.{2 * i}.bias"]
            x = (MM.apply(x, w) if i < 3 else x @ w.t()) + bias
            if i < 3:
                x = torch.nn.functional.elu(x)
        return x
    mu, v = mlp("actor", b["obs"]), mlp("critic", b["critic_obs"])
    dist = torch.distributions.Normal(mu, mu * 0.0 + pa["std"])
    lp = dist.log_prob(b["actions"]).sum(-1)
    ratio, A_ = torch.exp(lp - b["old_log_prob"].squeeze(-1)), b["advantages"].squeeze(-1)
    sl = torch.max(-A_ * ratio, -A_ * torch.clamp(ratio, 1.0 - clip, 1.0 + clip)).mean()
    vc = b["values"] + (v - b["values"]).clamp(-clip, clip)
    vl = torch.max((v - b["returns"]).pow(2), (vc - b["returns"]).pow(2)).mean()
    (sl + vcoef * vl - ecoef * dist.entropy().sum(-1).mean()).backward()
    return {k: t.grad for k, t in pa.items()}


def ppo_gradient_check(np_, torch, use_tc, atol, dims=(39, 168, 10)):
    """One PPO minibatch (gather -> forward -> fused heads / losses -> backward) at the registered network width through the C ABI,
    every gradient tensor and the KL / loss sums vs the hand-derived CPU oracle (oracle/ppo_oracle.py)."""
    import ctypes as C
    from grx_b200 import _lib as L
    from grx_b200.config import make_train_cfg
    from grx_b200.ppo import PPO, ActorCriticMLP
    from oracle import ppo_oracle as po
    torch.manual_seed(5)
    tc = make_train_cfg()
    (O, P, A), N, T = dims, 256, 8   # dims = (39, 168, 10) registered lower-limb task; (105, 234, 32) full-body task (un-fused output heads)
    ac = ActorCriticMLP(O, P, A, **tc["policy"])
    alg = PPO(ac, device="cuda:0", use_tensor_cores=use_tc, **dict(tc["algorithm"], num_mini_batches=2, num_learning_epochs=1))
    alg.init_storage(N, T)
    g = torch.Generator().manual_seed(7)
    obs, cobs = torch.randn(T, N, O, generator=g), torch.randn(T, N, P, generator=g)
    for s in range(T):
        alg.act(obs[s].cuda(), cobs[s].cuda(), eps=torch.randn(N, A, generator=g).cuda())
        alg.process_env_step((0.1 * torch.randn(N, generator=g)).cuda(), (torch.rand(N, generator=g) < 0.1).cuda(), {})
    alg.compute_returns(torch.randn(N, P, generator=g).cuda())
    with torch.no_grad():   # perturb the policy so that ratio != 1 and KL > 0
        for k, v in ac.state_dict().items():
            v.add_(0.02 * torch.randn(v.shape, generator=g).cuda() * (v.abs().mean() + 0.05))
    p_new = {k: v.cpu().clone() for k, v in ac.state_dict().items()}
    idx = torch.randperm(N * T, generator=g)
    alg._indices.copy_(idx.cuda())
    L.check(alg.lib.grx_ppo_minibatch_grads(alg._h, C.c_void_p(alg._indices.data_ptr()), 1, alg._stream()))
    torch.cuda.synchronize()
    st = alg.storage
    flat = lambda x: x.cpu().flatten(0, 1)
    sel = idx[alg.mini_batch_size:2 * alg.mini_batch_size]
    b = dict(obs=flat(st.obs)[sel], critic_obs=flat(st.critic_obs)[sel], actions=flat(st.actions)[sel], values=flat(st.values)[sel],
             advantages=flat(st.advantages)[sel], returns=flat(st.returns)[sel], old_log_prob=flat(st.actions_log_prob)[sel],
             old_mu=flat(st.mu)[sel], old_sigma=flat(st.sigma)[sel])
    stats, grads = po.minibatch_loss_and_grads(p_new, b, 0.2, 1.0, 0.01, True)
    got = alg.grads.cpu()
    emu = tf32_emulated_grads(torch, p_new, b, 0.2, 1.0, 0.01) if use_tc else None
    for k in p_new:
        gk = ac.view_of(got, k)
        scale = float(grads[k].abs().max()) + 1e-12
        err = float((gk - grads[k]).abs().max()) / scale
        if err <= atol:
            continue
        # beyond the tolerance against the fp32 oracle: accepted only where the SAME deviation is inherent to TF32 operands, i.e. the CUDA result
        # agrees (within the same tolerance) with the CPU emulation of the tensor-core arithmetic, and the emulation itself is that far from fp32
        assert emu is not None, (k, err)
        err_emu = float((gk - emu[k].reshape(gk.shape)).abs().max()) / scale
        inherent = float((emu[k].reshape(gk.shape) - grads[k]).abs().max()) / scale
        print(f"ppo_gradient_check {dims} {k}: {err:.2e} of scale vs the fp32 oracle, {err_emu:.2e} vs the TF32 emulation (emulation vs fp32: {inherent:.2e})")
        assert err_emu <= atol and inherent >= 0.5 * err, (k, err, err_emu, inherent)
    tail = alg.reduce_buf[-8:].cpu()
    rt = 1e-3 if not use_tc else 5e-3
    np_.testing.assert_allclose(float(tail[0] / tail[1]), float(stats["kl_mean"]), rtol=rt)
    np_.testing.assert_allclose(float(tail[2] / tail[1]), float(stats["surrogate_loss"]), rtol=rt, atol=1e-5)
    np_.testing.assert_allclose(float(tail[3] / tail[1]), float(stats["value_loss"]), rtol=rt)
    alg.close()
