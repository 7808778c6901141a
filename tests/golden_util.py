"""Helpers shared by the parity tests: rebuild configs / oracles from a golden fixture."""
import os

import numpy as np

from grx_b200.config import make_cfg, make_full_body_cfg
from grx_b200.robot import self_collision_pairs, task_tables
from grx_b200.urdf import builtin_model

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
ENV_FIXTURES = ["plane64_dec1", "plane_gr1t1", "hf_gr1t1", "hf_gr1t2_dr", "tm_gr1t1"]   # tm = mesh_type "trimesh" (BASELINE config #5)
# the reference classes on the unregistered full-body 32-DOF configuration (gr1t1_config.py:10-307; oracle/ref_harness/driver.py:full_body_cfg)
FULL_BODY_FIXTURES = ["plane_gr1t1_full", "hf_gr1t2_full"]
MAX_SELF_CONTACTS = 4


def load_fixture(name):
    return dict(np.load(os.path.join(GOLDEN, f"env_{name}.npz"), allow_pickle=False))


def cfg_from_fixture(fx):
    task, mesh = str(fx["meta/task"]), str(fx["meta/mesh_type"])
    N = fx["const/friction"].shape[0]
    cfg = make_full_body_cfg(task[:-5], N, mesh) if task.endswith("_full") else make_cfg(task, N, mesh)
    cfg.control.decimation = int(fx["meta/decimation"])
    fl = fx["meta/flags"]
    cfg.noise.add_noise, cfg.domain_rand.push_robots = bool(fl[0]), bool(fl[1])
    cfg.terrain.curriculum = bool(fl[2])
    cfg.domain_rand.randomize_init_dof_pos, cfg.domain_rand.randomize_init_base_velocity = bool(fl[3]), bool(fl[4])
    if "meta/terrain_rows_cols" in fx:
        cfg.terrain.num_rows, cfg.terrain.num_cols = [int(v) for v in fx["meta/terrain_rows_cols"]]
        cfg.terrain.max_init_terrain_level = cfg.terrain.num_rows - 1
    return cfg


def setup_from_fixture(fx):
    cfg = cfg_from_fixture(fx)
    model = builtin_model(str(fx["meta/task"]))
    tables = task_tables(model, cfg)
    consts = dict(friction=fx["const/friction"], restitution=fx["const/restitution"],
                  base_inertial=fx["const/base_inertial"], motor_strength=fx["const/motor_strength"],
                  env_origins=fx["init/env_origins"])
    terrain = None
    if "const/height_samples" in fx:
        consts.update(terrain_origins=fx["const/terrain_origins"], terrain_levels=fx["init/terrain_levels"],
                      terrain_types=fx["const/terrain_types"])
        terrain = dict(heights=fx["const/height_samples"], hscale=cfg.terrain.horizontal_scale,
                       vscale=cfg.terrain.vertical_scale, border=float(cfg.terrain.border_size),
                       friction=cfg.terrain.static_friction, restitution=cfg.terrain.restitution)
        if str(fx["meta/mesh_type"]) == "trimesh":   # structured trimesh: the vertex shifts of the reference's conversion (terrain.py is SHA-pinned to it)
            from grx_b200.terrain import heightfield_to_trimesh
            from oracle.phys import moves_from_vertices
            hs = np.asarray(fx["const/height_samples"], np.int16)
            verts, _ = heightfield_to_trimesh(hs, cfg.terrain.horizontal_scale, cfg.terrain.vertical_scale, cfg.terrain.slope_treshold)
            terrain["moves"] = moves_from_vertices(verts, hs.shape[0], hs.shape[1], cfg.terrain.horizontal_scale)
    return cfg, model, tables, consts, terrain


def phys_oracle(cfg, model, tables, terrain, dtype=np.float32):
    """The C physics oracle for a fixture's model: full-body trees run with robot self-collision (legged_robot_config.py:121), with the same
    candidate pairs the product derives."""
    from oracle.phys import PhysOracle
    sim = dict(dt=cfg.sim.dt, decimation=cfg.control.decimation, action_scale=cfg.control.action_scale)
    ctl = dict(tables)
    if model["nd"] > 10:
        ctl["self_pairs"] = self_collision_pairs(model, tables)
        sim["max_self_contacts"] = MAX_SELF_CONTACTS
    return PhysOracle(model, ctl, terrain, dtype=dtype, sim=sim)


def init_state(fx):
    return {k[len("init/"):]: v for k, v in fx.items() if k.startswith("init/")}


def step_items(fx, t, group):
    pre = f"step{t:02d}/{group}/"
    return {k[len(pre):]: v for k, v in fx.items() if k.startswith(pre)}


def wide_inputs(seed, N, T, O, P, A):
    """Inputs of the registered-width fixture, regenerated (not stored) by the test from the same CPU generator stream; the fixture keeps
    a SHA-256 of them so that a drifting generator fails loudly instead of silently comparing different rollouts."""
    import hashlib
    import torch
    g = torch.Generator().manual_seed(seed + 100)
    d = dict(eps=torch.randn(T, N, A, generator=g), obs=torch.randn(T, N, O, generator=g), critic_obs=torch.randn(T, N, P, generator=g),
             rewards=0.1 * torch.randn(T, N, generator=g), dones=torch.rand(T, N, generator=g) < 0.05)
    d["time_outs"] = d["dones"] & (torch.rand(T, N, generator=g) < 0.5)
    d["last_critic_obs"] = torch.randn(N, P, generator=g)
    d["indices"] = torch.randperm(N * T, generator=g)
    h = hashlib.sha256()
    for k in sorted(d):
        h.update(d[k].numpy().tobytes())
    return d, h.hexdigest()
