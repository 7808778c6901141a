"""Shared by the GPU parity tests and __graft_entry__.smoke(): build the CUDA env from a golden fixture and compare."""
import numpy as np

from golden_util import init_state, load_fixture, setup_from_fixture, step_items  # noqa: F401

# post-physics map (the reference's own arithmetic): fp32 on both sides, different exp/sqrt implementations and summation trees
POST_TOL = dict(rtol=1e-5, atol=3e-6)
# full step incl. 10 substeps of our dynamics spec: CUDA fp32 (Delassus-space PGS, branch-sparse Cholesky, fused multiply-add)
# vs the C oracle in fp32 (velocity-space PGS, dense Cholesky, no contraction) — same equations, different rounding
PHYS_TOL = dict(rtol=2e-3, atol=2e-3)


def make_gpu_env(fx, **kw):
    import torch
    from grx_b200.env import GRXVecEnv
    cfg, model, tables, consts, terrain = setup_from_fixture(fx)
    cfg.env.num_envs = len(consts["friction"])
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
    from oracle.phys import PhysOracle
    cfg, model, tables, consts, terrain = setup_from_fixture(fx)
    phys = PhysOracle(model, tables, terrain, dtype=np.float32,
                      sim=dict(dt=cfg.sim.dt, decimation=cfg.control.decimation, action_scale=cfg.control.action_scale))
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
