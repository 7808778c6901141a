"""TEST INFRASTRUCTURE — generate tests/golden/env_*.npz from the UNMODIFIED reference env classes.

Build-container only:  python -m oracle.ref_harness.gen_golden
Each fixture = one scenario: constants, the carried state before the first step, and per step the
inputs (actions, U, delay), the physics outputs seen at post_physics_step entry, and every output /
carried buffer after the step.  Physics inside is the C oracle (stand-in for PhysX); everything else is
the reference's own arithmetic (legged_robot.py / legged_robot_fftai.py / gr1t1.py).
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "wiki-grx-gym_b200"))

from grx_b200 import rng_layout as L  # noqa: E402
from oracle.ref_harness.driver import dump_state, env_constants, injected_step, make_reference_env  # noqa: E402

CARRIED = ("root_states", "dof_pos", "dof_vel", "last_dof_vel", "last_actions", "last_last_actions", "commands",
           "base_heights_offset", "feet_air_time", "feet_land_time", "feet_contact_last", "episode_length_buf",
           "episode_sums", "common_step_counter", "terrain_levels", "env_origins")
OUTPUTS = ("obs_buf", "pri_obs_buf", "rew_buf", "reset_buf", "time_out_buf", "base_lin_vel", "base_ang_vel",
           "base_projected_gravity", "feet_height", "measured_heights", "torques")


def scenario(name, task, num_envs, mesh_type, steps, seed, mutate_cfg=None, rows=None, cols=None, full_body=False):
    env, cfg = make_reference_env(task, num_envs, mesh_type, seed=seed, mutate_cfg=mutate_cfg, terrain_rows=rows,
                                  terrain_cols=cols, full_body=full_body)
    K = L.layout(env.num_actions).K
    env.reset()
    g = torch.Generator().manual_seed(1000 + seed)
    N = num_envs
    # ---- spread the state for coverage: time-outs, command resampling, pushes, falls
    env.episode_length_buf[:] = torch.randint(0, int(env.max_episode_length), (N,), generator=g)
    env.episode_length_buf[0:3] = int(env.max_episode_length) - torch.tensor([0, 1, 2])      # time-outs in steps 1..3
    env.episode_length_buf[3:6] = 500 - torch.tensor([1, 2, 3])                                # command resample
    env.common_step_counter = int(env.cfg.domain_rand.push_interval) - 3                       # push at step 3
    tilt = torch.tensor([0.5, 0.0, 0.0, 0.866])                                               # 60 deg roll: falls soon
    env.root_states[6, 3:7] = tilt
    env.root_states[7, 3:7] = torch.tensor([0.6428, 0.0, 0.0, 0.7660])                        # 80 deg: |g_z| < 0.33 at once
    if hasattr(env, "terrain_levels"):
        env.root_states[8:10, 0] += 4.5                                                        # walked far: move_up
    consts = env_constants(env)
    snap = {}
    orig_pps = env.post_physics_step

    def hooked():
        snap["root_states_phys"] = env.root_states.clone().numpy()
        snap["dof_pos_phys"] = env.dof_pos.clone().numpy()
        snap["dof_vel_phys"] = env.dof_vel.clone().numpy()
        snap["torques_phys"] = env.torques.clone().numpy()
        snap["foot_state"] = env.rigid_body_states[:, env.feet_indices].clone().numpy()
        snap["torso_quat"] = env.rigid_body_states[:, env.torso_indices][:, 0, 3:7].clone().numpy()
        snap["contact_forces"] = env.contact_forces.clone().numpy()
        snap["avg_feet_contact_force"] = env.avg_feet_contact_force.clone().numpy()
        snap["avg_feet_speed_xyz"] = env.avg_feet_speed_xyz.clone().numpy()
        snap["avg_feet_speed_rpy"] = env.avg_feet_speed_rpy.clone().numpy()
        return orig_pps()
    env.post_physics_step = hooked
    out = {("const/" + k): np.asarray(v) for k, v in consts.items() if k != "reward_names"}
    out["const/reward_names"] = np.array(consts["reward_names"])
    out["meta/task"], out["meta/mesh_type"], out["meta/steps"] = np.array(task + ("_full" if full_body else "")), np.array(mesh_type), np.array(steps)
    out["meta/decimation"] = np.array(env.cfg.control.decimation)
    out["meta/flags"] = np.array([int(env.cfg.noise.add_noise), int(env.cfg.domain_rand.push_robots),
                                  int(env.cfg.terrain.curriculum), int(env.cfg.domain_rand.randomize_init_dof_pos),
                                  int(env.cfg.domain_rand.randomize_init_base_velocity)])
    if hasattr(env, "terrain"):
        out["meta/terrain_rows_cols"] = np.array([env.cfg.terrain.num_rows, env.cfg.terrain.num_cols])
    s0 = dump_state(env)
    for k in CARRIED:
        if k in s0:
            out["init/" + k] = s0[k]
    n_reset = 0
    for t in range(steps):
        actions = 0.3 * torch.randn(N, env.num_actions, generator=g)
        actions[:, 3] += 0.2
        if t % 4 == 1:
            actions[0] = 5.0        # exercises the per-joint action clip
        U = torch.rand(N, K, generator=g)
        delay = max(0.0, float(5 + 2 * torch.randn(1, generator=g)))
        injected_step(env, actions, U, delay)
        s = dump_state(env)
        pre = f"step{t:02d}/"
        out[pre + "actions"], out[pre + "U"], out[pre + "delay"] = actions.numpy(), U.numpy(), np.array(delay)
        for k, v in snap.items():
            out[pre + "phys/" + k] = v
        for k in OUTPUTS:
            if k in s:
                out[pre + "out/" + k] = s[k]
        for k in CARRIED:
            if k in s:
                out[pre + "state/" + k] = s[k]
        if "episode" in env.extras and s["reset_buf"].any():
            names = consts["reward_names"]
            out[pre + "extras_episode"] = np.array([float(env.extras["episode"]["rew_" + n]) for n in names], np.float32)
            if "terrain_level" in env.extras["episode"]:
                out[pre + "extras_terrain_level"] = np.array(float(env.extras["episode"]["terrain_level"]), np.float32)
        n_reset += int(s["reset_buf"].sum())
    path = os.path.join(ROOT, "tests", "golden", f"env_{name}.npz")
    np.savez_compressed(path, **out)
    print(f"{name}: N={N} steps={steps} resets={n_reset}  -> {path} ({os.path.getsize(path) / 1e3:.0f} kB)")


def _dec1(cfg):   # BASELINE config #1: plane, 64 envs, 1 substep, noise/push/DR off
    cfg.control.decimation = 1
    cfg.noise.add_noise = False
    dr = cfg.domain_rand
    dr.push_robots = dr.randomize_friction = dr.randomize_restitution = dr.randomize_base_mass = False
    dr.randomize_base_com = dr.randomize_motor_strength = False
    dr.randomize_init_dof_pos = dr.randomize_init_base_velocity = False


def main():
    scenario("plane64_dec1", "GR1T1", 64, "plane", steps=6, seed=1, mutate_cfg=_dec1)
    scenario("plane_gr1t1", "GR1T1", 32, "plane", steps=10, seed=2)
    scenario("hf_gr1t1", "GR1T1", 32, "heightfield", steps=10, seed=3, rows=3, cols=4)
    scenario("hf_gr1t2_dr", "GR1T2", 32, "heightfield", steps=8, seed=4, rows=3, cols=4)
    scenario("tm_gr1t1", "GR1T1", 32, "trimesh", steps=8, seed=5, rows=3, cols=4)   # mesh_type = 'trimesh' + curriculum (BASELINE config #5)
    # the reference classes on the UNREGISTERED full-body 32-DOF configuration (driver.full_body_cfg), robot self-collision on
    scenario("plane_gr1t1_full", "GR1T1", 16, "plane", steps=8, seed=6, full_body=True)
    scenario("hf_gr1t2_full", "GR1T2", 16, "heightfield", steps=8, seed=7, rows=3, cols=4, full_body=True)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "full":   # only the full-body fixtures
        scenario("plane_gr1t1_full", "GR1T1", 16, "plane", steps=8, seed=6, full_body=True)
        scenario("hf_gr1t2_full", "GR1T2", 16, "heightfield", steps=8, seed=7, rows=3, cols=4, full_body=True)
    else:
        main()
