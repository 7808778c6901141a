"""Physics oracle (GRX-dyn v1) self-consistency: the reference pins nothing at the physics boundary
(SURVEY.md §8c — PhysX is binary-only), so the spec is validated through physical invariants."""
import numpy as np
import pytest

from grx_b200.config import make_cfg
from grx_b200.robot import nominal_params, task_tables
from grx_b200.urdf import builtin_model
from oracle.phys import PhysOracle


@pytest.fixture(scope="module", params=["GR1T1", "GR1T2"])
def setup(request):
    model = builtin_model(request.param)
    cfg = make_cfg(request.param)
    return model, cfg, task_tables(model, cfg)


def _rand_state(rng, tt, n):
    root = np.zeros((n, 13))
    root[:, :3] = rng.uniform(-1, 1, (n, 3)) + [0, 0, 3.0]
    qt = rng.normal(size=(n, 4))
    root[:, 3:7] = qt / np.linalg.norm(qt, axis=1, keepdims=True)
    root[:, 7:13] = rng.uniform(-1, 1, (n, 6))
    q = rng.uniform(tt["hard_lower"] * 0.8, tt["hard_upper"] * 0.8, (n, 10))
    qd = rng.uniform(-2, 2, (n, 10))
    return root, q, qd


def test_mass_matrix_matches_body_kinetic_energy(setup):
    model, cfg, tt = setup
    po = PhysOracle(model, tt, None, dtype=np.float64)
    bi = nominal_params(model, 1)["base_inertial"][0]
    rng = np.random.default_rng(0)
    root, q, qd = _rand_state(rng, tt, 20)
    for e in range(20):
        M, h, en = po.dynamics_terms(bi, root[e], q[e], qd[e])
        u = np.concatenate([qd[e], root[e, 7:13]])
        assert np.abs(M - M.T).max() < 1e-12
        assert np.linalg.eigvalsh(M).min() > 0
        assert abs(0.5 * u @ M @ u - en[0]) < 1e-10 * max(1.0, en[0])
        assert abs(M[10, 10] - model["mass"].sum()) < 1e-9


def test_free_flight_conserves_energy_and_momentum(setup):
    """tau = 0, no contact (robot 3 m up): energy drift -> 0 linearly in dt; linear momentum follows gravity;
    angular momentum about the COM is conserved.  A wrong Coriolis/centrifugal term breaks this at O(1)."""
    model, cfg, tt = setup
    bi = nominal_params(model, 1)
    rng = np.random.default_rng(1)
    root0, q0, qd0 = _rand_state(rng, tt, 1)
    qd0 *= 0.5
    zero_gain = dict(tt, kp=np.zeros(10), kd=np.zeros(10))
    drifts = []
    for dt in (1e-3, 5e-4):
        po = PhysOracle(model, zero_gain, None, sim=dict(dt=dt, decimation=int(round(0.05 / dt))), dtype=np.float64)
        root, q, qd = root0.copy(), q0.copy(), qd0.copy()
        M, h, e0 = po.dynamics_terms(bi["base_inertial"][0], root[0], q[0], qd[0])
        u0 = np.concatenate([qd[0], root[0, 7:13]])
        p0 = M[10:13] @ u0
        po.step(root, q, qd, np.zeros((1, 10)), np.zeros((1, 10)), 0.0, bi["motor_strength"], bi["base_inertial"],
                bi["friction"], bi["restitution"])
        M, h, e1 = po.dynamics_terms(bi["base_inertial"][0], root[0], q[0], qd[0])
        u1 = np.concatenate([qd[0], root[0, 7:13]])
        p1 = M[10:13] @ u1
        mtot = model["mass"].sum()
        assert np.allclose(p1 - p0, [0, 0, -9.81 * mtot * 0.05], atol=5e-5 * mtot)  # O(dt) integrator error
        drifts.append(abs(e1.sum() - e0.sum()) / max(1.0, abs(e0[0])))
    assert drifts[0] < 2e-2 and drifts[1] < 0.6 * drifts[0] + 1e-9, drifts


def test_static_stand_supports_weight(setup):
    """Standing under PD control on the plane: mean of the net vertical contact force over the feet = m g."""
    model, cfg, tt = setup
    po = PhysOracle(model, tt, None, dtype=np.float64)
    n = 2
    par = nominal_params(model, n)
    root = np.zeros((n, 13)); root[:, 2] = 0.95; root[:, 6] = 1.0
    q = np.tile(tt["default_pos"], (n, 1)); qd = np.zeros((n, 10))
    act = np.zeros((n, 10))
    fz = []
    for step in range(30):
        out = po.step(root, q, qd, act, act, 0.0, par["motor_strength"], par["base_inertial"], par["friction"], par["restitution"])
        if step >= 10:
            fz.append(out["contact_force"][0, :, 2].sum())
        non_feet = [l for l in range(len(model["link_names"])) if l not in tt["foot_links"]]
        assert np.abs(out["contact_force"][:, non_feet]).max() == 0.0
    mg = model["mass"].sum() * 9.81
    assert abs(np.mean(fz) - mg) < 0.05 * mg
    assert 0.80 < root[0, 2] < 0.95


def test_f32_tracks_f64(setup):
    model, cfg, tt = setup
    n = 8
    par = nominal_params(model, n)
    rng = np.random.default_rng(3)
    res = {}
    for dt in (np.float32, np.float64):
        po = PhysOracle(model, tt, None, dtype=dt)
        root = np.zeros((n, 13), dt); root[:, 2] = 0.95; root[:, 6] = 1.0
        q = np.tile(tt["default_pos"], (n, 1)).astype(dt); qd = np.zeros((n, 10), dt)
        act = (0.1 * np.random.default_rng(5).normal(size=(n, 10))).astype(dt)
        for step in range(5):
            out = po.step(root, q, qd, act, act, 0.0, par["motor_strength"], par["base_inertial"], par["friction"], par["restitution"])
        res[dt] = (root.copy(), q.copy(), qd.copy())
    assert np.abs(res[np.float32][0][:, :7] - res[np.float64][0][:, :7]).max() < 2e-4
    assert np.abs(res[np.float32][1] - res[np.float64][1]).max() < 5e-4


def test_heightfield_slope_contact():
    """A sphere-footed robot on a tilted heightfield gets contact normals from the triangle under each sphere."""
    model = builtin_model("GR1T1"); cfg = make_cfg("GR1T1"); tt = task_tables(model, cfg)
    rows = cols = 200
    hs = np.zeros((rows, cols), np.int16)
    hs += (np.arange(rows)[:, None] * 4).astype(np.int16)   # 4*0.005/0.1 = 0.2 slope along x
    terr = dict(heights=hs, hscale=0.1, vscale=0.005, border=10.0, friction=1.0, restitution=0.0)
    po = PhysOracle(model, tt, terr, dtype=np.float64)
    par = nominal_params(model, 1)
    root = np.zeros((1, 13)); root[0, 2] = 0.95 + 0.2 * 10.0; root[0, 6] = 1.0
    q = tt["default_pos"][None].copy(); qd = np.zeros((1, 10)); act = np.zeros((1, 10))
    for step in range(15):
        out = po.step(root, q, qd, act, act, 0.0, par["motor_strength"], par["base_inertial"], par["friction"], par["restitution"])
    f = out["contact_force"][0].sum(0)
    assert f[2] > 300 and f[0] < 0           # support force has a down-slope-opposing (−x) normal component
