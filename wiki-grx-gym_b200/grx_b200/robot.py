"""Task-level tables derived from (model, cfg): what the reference computes with Python
loops over names at env creation (legged_robot.py:176-192 PD gains by substring,
:594-616 DOF limits and soft limits, :1092-1161 + gr1t1.py:18-113 body indices by
substring, gr1t1.py:127-253 joint indices by substring)."""
from __future__ import annotations

import numpy as np

from .urdf import base_inertial_for, base_inertials_batch


def _get(obj, name, default=None):
    return getattr(obj, name, default)


def task_tables(model, cfg):
    """Return a dict of numpy arrays / index lists for the env kernel and the oracles."""
    nd = model["nd"]
    dof_names, link_names = model["dof_names"], model["link_names"]
    kp, kd, q0 = np.zeros(nd), np.zeros(nd), np.zeros(nd)
    for i, name in enumerate(dof_names):
        q0[i] = cfg.init_state.default_joint_angles[name]           # legged_robot.py:179-180
        for key in cfg.control.stiffness.keys():                     # legged_robot.py:182-186 (last match wins)
            if key in name:
                kp[i], kd[i] = cfg.control.stiffness[key], cfg.control.damping[key]
    lo, hi = model["dof_lower"].copy(), model["dof_upper"].copy()
    mid, rng = (lo + hi) / 2, hi - lo                                 # legged_robot.py:607-610
    soft = cfg.rewards.soft_dof_pos_limit
    soft_lo, soft_hi = mid - 0.5 * rng * soft, mid + 0.5 * rng * soft
    a = cfg.asset
    feet = [i for i, s in enumerate(link_names) if a.foot_name in s]                  # gr1t1.py:30,67-69
    torso = [i for i, s in enumerate(link_names) if a.torso_name in s]                # gr1t1.py:21,39-41
    term = []
    for name in a.terminate_after_contacts_on:                                         # legged_robot.py:1148-1159
        term.extend(i for i, s in enumerate(link_names) if name in s)
    ankle = [i for i, s in enumerate(dof_names) if _get(a, "ankle_name", "ankle") in s]  # gr1t1.py:219-220
    # contact-sphere priority: spheres on termination links first (they must never lose their slot to the 8-contact cap,
    # or a fall would go undetected), then the rest in model order
    ts = set(term)
    sph_order = sorted(range(len(model["sph_rad"])), key=lambda s: (0 if int(model["sph_link"][s]) in ts else 1, s))
    return dict(sph_order=np.array(sph_order, dtype=np.int32), kp=kp, kd=kd, default_pos=q0, torque_limits=model["dof_effort"].copy(),
                dof_vel_limits=model["dof_velocity"].copy(), hard_lower=lo, hard_upper=hi,
                soft_lower=soft_lo, soft_upper=soft_hi, foot_links=feet, torso_links=torso,
                termination_links=term, ankle_dofs=ankle)


def _axang(a, th):
    c, s_, t = np.cos(th), np.sin(th), 1 - np.cos(th)
    return np.array([[c + a[0] * a[0] * t, a[0] * a[1] * t - a[2] * s_, a[0] * a[2] * t + a[1] * s_],
                     [a[1] * a[0] * t + a[2] * s_, c + a[1] * a[1] * t, a[1] * a[2] * t - a[0] * s_],
                     [a[2] * a[0] * t - a[1] * s_, a[2] * a[1] * t + a[0] * s_, c + a[2] * a[2] * t]])


def forward_kinematics(model, q):
    """Body frames (R [nb, 3, 3], o [nb, 3]) of the dynamic tree with the base at the origin, joint angles q [nd] (host-side helper)."""
    nb = model["nb"]
    R, o = np.zeros((nb, 3, 3)), np.zeros((nb, 3))
    R[0] = np.eye(3)
    jrot = np.asarray(model["jrot"]).reshape(nb, 3, 3)
    for b in range(1, nb):
        p = int(model["parent"][b])
        R[b] = R[p] @ jrot[b] @ _axang(np.asarray(model["axis"]).reshape(nb, 3)[b], q[b - 1])
        o[b] = o[p] + R[p] @ np.asarray(model["jpos"]).reshape(nb, 3)[b]
    return R, o


def self_collision_pairs(model, tables, max_pairs=64, min_rest_gap=0.005):
    """Candidate sphere pairs for robot self-collision (legged_robot_config.py:121 `self_collisions = 0` = enabled; every actor is created
    with collision_filter 0, legged_robot.py:1022-1028, so all of its shapes collide with each other except those on directly connected
    links).  Indices refer to the contact-PRIORITY order of the spheres (tables['sph_order']).  Kept: pairs on different, non-adjacent
    dynamic bodies that are apart by more than `min_rest_gap` in the default pose (shapes that touch by construction would otherwise push the
    robot apart at rest); ordered by that rest gap (the closest pairs — feet, shanks, hands vs thighs — first), at most `max_pairs`."""
    order = np.asarray(tables["sph_order"])
    body, pos, rad = np.asarray(model["sph_body"])[order], np.asarray(model["sph_pos"]).reshape(-1, 3)[order], np.asarray(model["sph_rad"])[order]
    parent = np.asarray(model["parent"])
    R, o = forward_kinematics(model, np.asarray(tables["default_pos"], float))
    x = np.stack([o[b] + R[b] @ p for b, p in zip(body, pos)])
    cand = []
    for a in range(len(rad)):
        for b in range(a + 1, len(rad)):
            ba, bb = int(body[a]), int(body[b])
            if ba == bb or parent[ba] == bb or parent[bb] == ba:
                continue
            gap = float(np.linalg.norm(x[a] - x[b]) - rad[a] - rad[b])
            if gap > min_rest_gap:
                cand.append((gap, a, b))
    cand.sort()
    return np.array([[a, b] for _, a, b in cand[:max_pairs]], dtype=np.int32).reshape(-1, 2)


def sample_domain_rand(model, cfg, num_envs, rng_np, rng_torch_cpu=None):
    """Per-env randomised physical parameters, vectorised.

    Same distributions and the same generators' *kinds* as the reference loop
    (legged_robot.py:538-580 friction / restitution: 64 buckets + randint bucket ids;
    :618-648 base mass / COM with numpy; :1060-1064 motor strength), but drawn in
    bulk instead of inside an O(num_envs) Python loop, so streams differ.
    Returns dict(friction[N], restitution[N], motor_strength[N,nd], base_inertial[N,10]).
    """
    dr = cfg.domain_rand
    N, nd = num_envs, model["nd"]
    fr = np.ones(N)
    if dr.randomize_friction:
        buckets = rng_np.uniform(dr.friction_range[0], dr.friction_range[1], 64)
        fr = buckets[rng_np.integers(0, 64, N)]
    # NOTE: un-randomised shapes keep the importer default; Isaac Gym's default shape friction is 1.0, restitution 0.
    rs = np.zeros(N)
    if dr.randomize_restitution:
        buckets = rng_np.uniform(dr.restitution_range[0], dr.restitution_range[1], 64)
        rs = buckets[rng_np.integers(0, 64, N)]
    ms = np.ones((N, nd))
    if dr.randomize_motor_strength:
        ms = rng_np.uniform(dr.multiply_motor_strength[0], dr.multiply_motor_strength[1], (N, nd))
    scale = np.ones(N)
    if dr.randomize_base_mass:
        scale = rng_np.uniform(dr.multiply_base_mass_range[0], dr.multiply_base_mass_range[1], N)
    off = np.zeros((N, 3))
    if dr.randomize_base_com:
        off = np.stack([rng_np.uniform(*dr.add_base_com_range_x, N), rng_np.uniform(*dr.add_base_com_range_y, N),
                        rng_np.uniform(*dr.add_base_com_range_z, N)], axis=1)
    bi = base_inertials_batch(model, scale, off)                                      # no O(num_envs) Python loop (the reference: legged_robot.py:1008-1082)
    return dict(friction=fr, restitution=rs, motor_strength=ms, base_inertial=bi)


def nominal_params(model, num_envs):
    m, c, I6 = base_inertial_for(model)
    bi = np.tile(np.concatenate([[m], c, I6]), (num_envs, 1))
    return dict(friction=np.ones(num_envs), restitution=np.zeros(num_envs),
                motor_strength=np.ones((num_envs, model["nd"])), base_inertial=bi)
