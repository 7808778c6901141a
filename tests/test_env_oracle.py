"""Pin the env oracle (oracle/env_oracle.py) to the reference: golden trajectories were produced by the
UNMODIFIED reference classes (GR1T1 / GR1T2 over FakeGym, oracle/ref_harness/gen_golden.py)."""
import numpy as np
import pytest
import torch

from golden_util import ENV_FIXTURES, FULL_BODY_FIXTURES, init_state, load_fixture, phys_oracle, setup_from_fixture, step_items
from oracle.env_oracle import EnvOracle

# torch CPU fp32 on both sides, same op order -> near bit-exact; physics = the same C code
TOL = dict(rtol=1e-6, atol=1e-6)


def _make(fx):
    cfg, model, tables, consts, terrain = setup_from_fixture(fx)
    phys = phys_oracle(cfg, model, tables, terrain)
    env = EnvOracle(cfg, tables, consts, phys, terrain)
    env.load_state(init_state(fx))
    return env


@pytest.mark.parametrize("name", ENV_FIXTURES + FULL_BODY_FIXTURES)
def test_full_step_matches_reference(name):
    fx = load_fixture(name)
    env = _make(fx)
    n_reset = 0
    for t in range(int(fx["meta/steps"])):
        pre = f"step{t:02d}/"
        env.step(fx[pre + "actions"], fx[pre + "U"], float(fx[pre + "delay"]))
        out, st = step_items(fx, t, "out"), step_items(fx, t, "state")
        np.testing.assert_array_equal(env.reset_buf.numpy(), out["reset_buf"].astype(bool), err_msg=f"{name} t={t} reset")
        np.testing.assert_array_equal(env.time_out_buf.numpy(), out["time_out_buf"].astype(bool))
        for k, v in (("obs_buf", env.obs_buf), ("pri_obs_buf", env.pri_obs_buf), ("rew_buf", env.rew_buf),
                     ("torques", env.torques), ("base_lin_vel", env.dbg["base_lin_vel"]),
                     ("feet_height", env.dbg["feet_height"]), ("measured_heights", env.dbg["measured_heights"])):
            ref = out[k]
            if ref.ndim == 0:      # measured_heights == 0 scalar on plane
                continue
            np.testing.assert_allclose(v.numpy(), ref, err_msg=f"{name} t={t} {k}", **TOL)
        for k in EnvOracle.CARRIED:
            np.testing.assert_allclose(getattr(env, k).numpy().astype(np.float64), st[k].astype(np.float64).reshape(getattr(env, k).shape),
                                       err_msg=f"{name} t={t} state {k}", **TOL)
        if "terrain_levels" in st:
            np.testing.assert_array_equal(env.terrain_levels.numpy(), st["terrain_levels"])
            np.testing.assert_allclose(env.env_origins.numpy(), st["env_origins"], **TOL)
        if pre + "extras_episode" in fx:
            got = np.array([float(env.extras["episode"]["rew_" + n]) for n in env.reward_names], np.float32)
            np.testing.assert_allclose(got, fx[pre + "extras_episode"], rtol=1e-5, atol=1e-7)
        n_reset += int(env.reset_buf.sum())
    assert n_reset >= 3


@pytest.mark.parametrize("name", ENV_FIXTURES + FULL_BODY_FIXTURES)
def test_post_physics_only_matches_reference(name):
    """Same, but with the reference's own physics outputs injected (isolates LR/FF/G1 arithmetic)."""
    fx = load_fixture(name)
    env = _make(fx)
    for t in range(int(fx["meta/steps"])):
        pre = f"step{t:02d}/"
        ph = {k: torch.from_numpy(v) for k, v in step_items(fx, t, "phys").items()}
        env.actions = torch.clip(torch.from_numpy(fx[pre + "actions"]), env.clip_min, env.clip_max)
        env.root_states.copy_(ph["root_states_phys"]); env.dof_pos.copy_(ph["dof_pos_phys"]); env.dof_vel.copy_(ph["dof_vel_phys"])
        ph["torques"] = ph["torques_phys"]
        env.post_physics(ph, torch.from_numpy(fx[pre + "U"]))
        out = step_items(fx, t, "out")
        np.testing.assert_allclose(env.obs_buf.numpy(), out["obs_buf"], **TOL)
        np.testing.assert_allclose(env.pri_obs_buf.numpy(), out["pri_obs_buf"], **TOL)
        np.testing.assert_allclose(env.rew_buf.numpy(), out["rew_buf"], **TOL)
        np.testing.assert_array_equal(env.reset_buf.numpy(), out["reset_buf"].astype(bool))


@pytest.mark.parametrize("name", ENV_FIXTURES + FULL_BODY_FIXTURES)
def test_task_tables_match_reference_constants(name):
    """PD gains / limits / body indices our host code derives == what the reference env derived (LR:176-192, 594-616, G1:18-113)."""
    fx = load_fixture(name)
    cfg, model, tb, consts, terrain = setup_from_fixture(fx)
    np.testing.assert_allclose(tb["kp"], fx["const/p_gains"], rtol=1e-6)
    np.testing.assert_allclose(tb["kd"], fx["const/d_gains"], rtol=1e-6)
    np.testing.assert_allclose(tb["default_pos"], fx["const/default_dof_pos"], rtol=1e-6)
    np.testing.assert_allclose(tb["torque_limits"], fx["const/torque_limits"], rtol=1e-6)
    np.testing.assert_allclose(tb["dof_vel_limits"], fx["const/dof_vel_limits"], rtol=1e-6)
    np.testing.assert_allclose(np.stack([tb["soft_lower"], tb["soft_upper"]], 1), fx["const/dof_pos_limits"], rtol=1e-5, atol=1e-6)
    assert list(tb["foot_links"]) == list(fx["const/feet_indices"])
    assert list(tb["torso_links"]) == list(fx["const/torso_indices"])
    assert list(tb["termination_links"]) == list(fx["const/termination_contact_indices"])
    env = _make(fx)
    np.testing.assert_allclose(env.noise_scale_vec.numpy(), fx["const/noise_scale_vec"], rtol=1e-6)
    np.testing.assert_allclose([env.reward_scales[n] for n in env.reward_names], fx["const/reward_scales"], rtol=1e-12)
    assert env.max_episode_length == float(fx["const/max_episode_length"]) and env.push_interval == float(fx["const/push_interval"])
