"""The full-body 32-DOF GR1T1 / GR1T2 (the reference's unregistered GR1T1Cfg / GR1T2Cfg, gr1t1_config.py:10-307: 33 bodies, 32 actions,
observation 9 + 3 * 32 = 105, privileged 234) with robot self-collision (legged_robot_config.py:121) behind the SAME GRXVecEnv /
OnPolicyRunner surface, on the generic-topology kernels.  Parity against the reference classes and the C oracle is in tests/test_env_gpu.py
(fixtures plane_gr1t1_full / hf_gr1t2_full); here: behaviour at size, the physical envelope, and PPO end to end."""
import numpy as np
import pytest
import torch

from grx_b200.config import make_full_body_cfg, make_train_cfg

pytestmark = pytest.mark.gpu


def _env(robot, N, mesh, **kw):
    from grx_b200.env import GRXVecEnv
    cfg = make_full_body_cfg(robot, N, mesh)
    return GRXVecEnv(cfg, sim_device="cuda:0", **kw), cfg


@pytest.mark.parametrize("robot,mesh", [("GR1T1", "heightfield"), ("GR1T2", "trimesh")])
def test_full_body_properties_at_size(robot, mesh):
    N = 2048
    env, cfg = _env(robot, N, mesh)
    assert env.generic and env.num_actions == 32 and env.rng_k == 28 + 4 * 32 and len(env.self_collision_pairs) > 0
    obs, pri = env.reset()
    assert obs.shape == (N, 105) and pri.shape == (N, 234)
    g = torch.Generator(device="cuda").manual_seed(0)
    n_reset, n_steps = 0, 100
    lvl0 = env.terrain_levels.clone()
    for _ in range(n_steps):
        obs, pri, rew, reset, extras = env.step(0.2 * torch.randn(N, 32, device="cuda", generator=g))
        n_reset += int(reset.sum())
    torch.cuda.synchronize()
    for t in (obs, pri, rew, env.root_states, env.dof_pos, env.dof_vel, env.torques, env.contact_forces):
        assert torch.isfinite(t).all()
    assert obs.abs().max() <= 100.0 and pri.abs().max() <= 100.0
    assert 0 < n_reset < N * n_steps // 4
    noise = (obs - pri[:, :105]).abs().max(0).values.cpu().numpy()
    bound = np.array([0] * 3 + [0.05] * 3 + [0.03] * 3 + [0.04] * 32 + [0.2] * 32 + [0] * 32) + 1e-6
    assert (noise <= bound).all() and noise[3:73].min() > 0.0
    assert torch.allclose(env.root_states[:, 3:7].norm(dim=1), torch.ones(N, device="cuda"), atol=1e-4)
    lim_lo = torch.tensor(env.model["dof_lower"], device="cuda", dtype=torch.float32) - 0.05
    lim_hi = torch.tensor(env.model["dof_upper"], device="cuda", dtype=torch.float32) + 0.05
    assert ((env.dof_pos >= lim_lo) & (env.dof_pos <= lim_hi)).all(1).float().mean() > 0.9
    assert (env.terrain_levels != lvl0).any()
    tq = torch.tensor(env.model["dof_effort"], device="cuda", dtype=torch.float32)
    assert (env.torques.abs() <= tq + 1e-4).all()                                       # torque clip (legged_robot_fftai.py:64)
    assert set(extras["episode"].keys()) == {"rew_" + n for n in env.reward_names} | {"terrain_level"}


def test_full_body_stands_and_balances_its_weight():
    """Physical envelope: the full-body robot holding its default pose on the plane (zero actions = PD to the default angles) stays up and its
    contact forces carry its weight; with self-collision on, the default pose itself produces no self-contact forces (no spurious pairs).
    Window 0.5 s: with the full-body gains of gr1t1_config.py:158-178 (ankle pitch 11 N m/rad, ankle roll 0.25 — an order of magnitude below
    the registered lower-limb task's) the passive stance is only marginally stable; CUDA kernel and C oracle agree that most robots topple
    after ~1.1 s without a policy."""
    from grx_b200.robot import nominal_params
    from grx_b200.urdf import builtin_model
    from grx_b200.env import GRXVecEnv
    cfg = make_full_body_cfg("GR1T1", 32, "plane")
    cfg.noise.add_noise = False
    cfg.domain_rand.push_robots = cfg.domain_rand.randomize_init_dof_pos = cfg.domain_rand.randomize_init_base_velocity = False
    env = GRXVecEnv(cfg, sim_device="cuda:0", params=nominal_params(builtin_model("GR1T1_full"), 32))
    env.reset()
    alive = torch.ones(32, dtype=torch.bool, device="cuda")
    for _ in range(25):
        _, _, _, reset, _ = env.step(torch.zeros(32, 32, device="cuda"), delay=0.0)
        alive &= ~reset
    torch.cuda.synchronize()
    alive = alive.cpu().numpy()
    assert alive.mean() > 0.7, alive.mean()
    fz = env.contact_forces[:, :, 2].sum(1).cpu().numpy()
    mg = float(env.model["mass"][1:].sum() + env.params["base_inertial"][0, 0]) * 9.81
    np.testing.assert_allclose(fz[alive], mg, rtol=0.15)
    z = env.root_states[:, 2].cpu().numpy()
    assert (z[alive] > 0.8).all()
    feet = set(int(i) for i in env.feet_indices.cpu().numpy())
    others = [l for l in range(env.contact_forces.shape[1]) if l not in feet]
    assert float(env.contact_forces[torch.from_numpy(alive).cuda()][:, others].abs().max()) < 1.0   # only the feet touch anything


def test_full_body_runner_learns(tmp_path):
    """OnPolicyRunner on the 32-action task: actor 105 -> 512 -> 256 -> 128 -> 32, critic 234 -> ... -> 1 (the generic-width PPO kernels)."""
    from grx_b200.runner import OnPolicyRunner
    torch.manual_seed(1)
    env, cfg = _env("GR1T1", 512, "plane")
    tc = make_train_cfg()
    tc["runner"]["num_steps_per_env"] = 16
    tc["algorithm"]["num_mini_batches"] = 4
    tc["algorithm"]["num_learning_epochs"] = 2
    r = OnPolicyRunner(env, tc, log_dir=str(tmp_path), device="cuda:0")
    w0 = r.algorithm.params.clone()
    r.learn(3, init_at_random_ep_len=True)
    torch.cuda.synchronize()
    assert torch.isfinite(r.algorithm.params).all() and not torch.equal(w0, r.algorithm.params)
    sd = r.algorithm.actor_critic.state_dict()
    assert sd["std"].shape == (32,) and sd["actor.model.0.weight"].shape == (512, 105) and sd["critic.model.0.weight"].shape == (512, 234)
    a = r.get_inference_policy()(env.get_observations())
    assert a.shape == (512, 32) and torch.isfinite(a).all()
    assert np.isfinite(r.last_scalars["Loss/value_function"]) and np.isfinite(r.last_scalars["Loss/surrogate"])
