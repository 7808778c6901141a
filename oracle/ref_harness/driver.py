"""TEST INFRASTRUCTURE — drive the reference's unmodified GR1T1/GR1T2 env over FakeGym with injected
random draws, and dump every buffer.  Build-container only (imports /root/reference).

RNG injection: the reference draws with torch / numpy generators inside the hot path with
data-dependent shapes.  For parity the draws are replaced by table look-ups
``U[step][slot][env]`` (layout: grx_b200/rng_layout.py), by patching the *generator functions only*
(torch_rand_float in the env module namespaces, torch.rand_like, torch.randint_like,
numpy.random.normal).  The env code itself is untouched.
"""
from __future__ import annotations

import argparse
import contextlib
import copy
import io
import sys

import numpy as np
import torch

from . import stub


def full_body_cfg(task="GR1T1"):
    """The reference's UNREGISTERED full-body configuration (gr1t1_config.py:10-307 GR1T1Cfg / gr1t2_config.py GR1T2Cfg: 32 DOF) made runnable:
    its stale `num_obs = 121` replaced by what compute_observations (gr1t1.py:281-313) really produces (9 + 3 * 32 = 105; pri = 105 + 8 + 121),
    and — because upstream leaves every full-body reward scale at zero — the reward scales of the registered lower-limb task."""
    import legged_gym.envs  # noqa: F401
    if task == "GR1T2":
        from legged_gym.envs.gr1t2.gr1t2_config import GR1T2Cfg as Full
        from legged_gym.envs.gr1t2.gr1t2_lower_limb_config import GR1T2LowerLimbCfg as Low
    else:
        from legged_gym.envs.gr1t1.gr1t1_config import GR1T1Cfg as Full
        from legged_gym.envs.gr1t1.gr1t1_lower_limb_config import GR1T1LowerLimbCfg as Low
    cfg, low = Full(), Low()
    cfg.env.num_obs = 9 + 3 * cfg.env.num_actions
    cfg.env.num_pri_obs = cfg.env.num_obs + 8 + len(cfg.terrain.measured_points_x) * len(cfg.terrain.measured_points_y)
    for k in dir(low.rewards.scales):
        if not k.startswith("_"):
            setattr(cfg.rewards.scales, k, getattr(low.rewards.scales, k))
    return cfg


def make_reference_env(task="GR1T1", num_envs=64, mesh_type="plane", seed=1, mutate_cfg=None, quiet=True,
                       terrain_rows=None, terrain_cols=None, full_body=False):
    """Instantiate the reference task class over FakeGym (CPU).  Returns (env, env_cfg)."""
    stub.install()
    out = io.StringIO()
    with contextlib.redirect_stdout(out if quiet else sys.stdout):
        import legged_gym.envs  # noqa: F401  registers GR1T1 / GR1T2
        from legged_gym.utils.helpers import class_to_dict, parse_sim_params, set_seed
        from legged_gym.utils.task_registry import task_registry
        if full_body:
            env_cfg = full_body_cfg(task)
        else:
            env_cfg, _ = task_registry.get_cfgs(task)
        env_cfg = copy.deepcopy(env_cfg)
        env_cfg.env.num_envs = num_envs
        env_cfg.terrain.mesh_type = mesh_type
        if terrain_rows is not None:
            env_cfg.terrain.num_rows = terrain_rows
            env_cfg.terrain.max_init_terrain_level = terrain_rows - 1
        if terrain_cols is not None:
            env_cfg.terrain.num_cols = terrain_cols
        if mutate_cfg is not None:
            mutate_cfg(env_cfg)
        args = argparse.Namespace(physics_engine=1, use_gpu=False, subscenes=0, use_gpu_pipeline=False, num_threads=0,
                                  device="cpu", sim_device="cpu", headless=True)
        set_seed(seed)
        sim_params = parse_sim_params(args, {"sim": class_to_dict(env_cfg.sim)})
        env = task_registry.get_task_class(task)(cfg=env_cfg, sim_params=sim_params, physics_engine=1,
                                                 sim_device="cpu", headless=True)
    return env, env_cfg


class _Injector:
    """Context manager replacing the generator functions by table look-ups for one env.step()."""

    def __init__(self, env, U, delay):
        from grx_b200 import rng_layout
        L = rng_layout.layout(env.num_actions)   # slots of the draw table for this DOF count (== the module constants for 10 DOF)
        self.env, self.U, self.delay, self.L = env, U, float(delay), L
        self.saved = []

    def _caller_env_ids(self, depth=2):
        f = sys._getframe(depth)
        return f.f_code.co_name, f.f_locals.get("env_ids", None), f.f_back.f_code.co_name

    def __enter__(self):
        import legged_gym.envs.base.legged_robot as LR
        L, U = self.L, self.U
        counters = {}

        def rand_float(lower, upper, shape, device):
            fn, env_ids, parent = self._caller_env_ids()
            k = counters.get((fn, parent), 0)
            counters[(fn, parent)] = k + 1
            if fn == "_resample_commands":
                base = L.CMD_RESET if parent == "reset_idx" else L.CMD_TIME
                u = U[env_ids, base + k:base + k + 1]
            elif fn == "_reset_dofs":
                u = U[env_ids, L.RESET_DOF:L.RESET_DOF + shape[1]]
            elif fn == "_reset_root_states":
                base, width = [(L.RESET_XY, 2), (L.RESET_YAW, 1), (L.RESET_VEL, 6)][k if self.env.custom_origins else k + 1]
                u = U[env_ids, base:base + width]
            elif fn == "_push_robots":
                u = U[:, L.PUSH:L.PUSH + 2]
            else:
                raise RuntimeError(f"unexpected torch_rand_float call site {fn}")
            assert tuple(u.shape) == tuple(shape), (fn, u.shape, shape)
            return (upper - lower) * u + lower

        def rand_like(t, **kw):
            assert t.shape[1] == L.NOISE_N
            return U[:, L.NOISE:L.NOISE + t.shape[1]].clone()

        def randint_like(t, high, **kw):
            fn, env_ids, parent = self._caller_env_ids()
            assert fn == "_update_terrain_curriculum"
            return torch.floor(U[env_ids, L.CURRICULUM] * high).to(t.dtype).clamp(max=high - 1)

        def np_normal(loc=0.0, scale=1.0, size=None):
            return np.array([self.delay])

        self.saved = [(LR, "torch_rand_float", LR.torch_rand_float), (torch, "rand_like", torch.rand_like),
                      (torch, "randint_like", torch.randint_like), (np.random, "normal", np.random.normal)]
        LR.torch_rand_float = rand_float
        torch.rand_like = rand_like
        torch.randint_like = randint_like
        np.random.normal = np_normal
        return self

    def __exit__(self, *exc):
        for obj, name, val in self.saved:
            setattr(obj, name, val)


def injected_step(env, actions, U, delay):
    """env.step(actions) with the step's random draws taken from U [N, K] and the action delay = ``delay``."""
    with _Injector(env, U, delay):
        return env.step(actions)


STATE_KEYS = ("root_states", "dof_pos", "dof_vel", "last_dof_vel", "actions", "last_actions", "last_last_actions",
              "commands", "base_heights_offset", "feet_air_time", "feet_land_time", "feet_contact",
              "feet_contact_last", "feet_contact_filt", "episode_length_buf", "torques", "rigid_body_states",
              "contact_forces", "base_lin_vel", "base_ang_vel", "base_projected_gravity", "obs_buf", "pri_obs_buf",
              "rew_buf", "reset_buf", "time_out_buf", "feet_height", "avg_feet_contact_force",
              "avg_feet_speed_xyz", "avg_feet_speed_rpy", "dof_acc", "surround_heights_offset", "measured_heights")


def dump_state(env):
    d = {}
    for k in STATE_KEYS:
        v = getattr(env, k, None)
        if isinstance(v, torch.Tensor):
            d[k] = v.detach().clone().numpy()
        elif isinstance(v, (int, float)):
            d[k] = np.array(v)
    d["episode_sums"] = np.stack([env.episode_sums[n].numpy().copy() for n in env.reward_names], axis=1)
    d["common_step_counter"] = np.array(env.common_step_counter)
    if hasattr(env, "terrain_levels"):
        d["terrain_levels"] = env.terrain_levels.numpy().copy()
        d["terrain_types"] = env.terrain_types.numpy().copy()
    d["env_origins"] = env.env_origins.numpy().copy()
    return d


def env_constants(env):
    """Everything fixed at construction that the new env needs to start from the same place."""
    g = env.gym
    c = dict(friction=g.friction.copy(), restitution=g.restitution.copy(), base_inertial=g.base_inertial.copy(),
             motor_strength=env.motor_strength_scales.numpy().copy(), reward_names=list(env.reward_names),
             reward_scales=np.array([env.reward_scales[n] for n in env.reward_names], np.float64),
             noise_scale_vec=env.noise_scale_vec.numpy().copy(), env_origins=env.env_origins.numpy().copy(),
             p_gains=env.p_gains.numpy().copy(), d_gains=env.d_gains.numpy().copy(),
             default_dof_pos=env.default_dof_pos.numpy().copy()[0], torque_limits=env.torque_limits.numpy().copy(),
             dof_pos_limits=env.dof_pos_limits.numpy().copy(), dof_vel_limits=env.dof_vel_limits.numpy().copy(),
             feet_indices=env.feet_indices.numpy().copy(), torso_indices=env.torso_indices.numpy().copy(),
             termination_contact_indices=env.termination_contact_indices.numpy().copy(),
             max_episode_length=np.array(float(env.max_episode_length)), dt=np.array(env.dt),
             push_interval=np.array(float(env.cfg.domain_rand.push_interval)),
             resample_interval=np.array(int(env.cfg.commands.resample_command_interval)),
             custom_origins=np.array(bool(env.custom_origins)))
    if hasattr(env, "terrain_origins"):
        c["terrain_origins"] = env.terrain_origins.numpy().copy()
        c["height_samples"] = env.height_samples.numpy().copy()
        c["terrain_types"] = env.terrain_types.numpy().copy()
    return c
