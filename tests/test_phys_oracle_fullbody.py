"""Dynamics spec GRX-dyn v1 on the FULL-BODY GR1T1 / GR1T2 trees (33 bodies, 32 revolute DOF: legs, waist, head, arms — the
unregistered `GR1T1Cfg` / `GR1T2Cfg` models of gr1t1_config.py:10-307).  The C oracle is topology-generic; the spec itself is validated here on the
deeper tree with the same physical invariants as tests/test_phys_oracle.py (CPU only).  The CUDA side of the full-body task (generic-topology
kernels, csrc/grx_phys_generic.cu) is compared against this oracle in tests/test_physg_gpu.py and tests/test_env_gpu.py."""
import numpy as np
import pytest

from grx_b200.urdf import builtin_model
from grx_b200.robot import nominal_params
from oracle.phys import PhysOracle


@pytest.fixture(scope="module", params=["GR1T1_full", "GR1T2_full"])
def setup(request):
    model = builtin_model(request.param)
    nd = model["nd"]
    # gentle generic gains: the light distal links (wrists, head) make an explicit PD law stiff at dt = 2 ms with leg-sized gains
    ctl = dict(kp=np.full(nd, 20.0), kd=np.full(nd, 1.0), default_pos=np.zeros(nd), foot_links=[model["link_names"].index(n) for n in
               model["link_names"] if "foot_roll" in n], sph_order=np.arange(len(model["sph_rad"])))
    return model, ctl


def _rand_state(rng, model, n):
    nd = model["nd"]
    root = np.zeros((n, 13))
    root[:, :3] = rng.uniform(-1, 1, (n, 3)) + [0, 0, 3.0]
    qt = rng.normal(size=(n, 4))
    root[:, 3:7] = qt / np.linalg.norm(qt, axis=1, keepdims=True)
    root[:, 7:13] = rng.uniform(-1, 1, (n, 6))
    lo, hi = np.asarray(model["dof_lower"]), np.asarray(model["dof_upper"])
    q = rng.uniform(0.8 * lo, 0.8 * hi, (n, nd))
    qd = rng.uniform(-2, 2, (n, nd))
    return root, q, qd


def test_topology(setup):
    model, _ = setup
    assert model["nb"] == 33 and model["nd"] == 32 and len(model["link_names"]) == 37
    parent = np.asarray(model["parent"])
    assert parent[0] == -1 and (parent[1:] < np.arange(1, model["nb"])).all()          # topologically ordered tree
    assert np.bincount(parent[1:], minlength=model["nb"]).max() >= 3                      # branching: legs + waist chain from the base
    assert 50.0 < float(np.sum(model["mass"])) < 60.0                                     # 52.83 kg (GR1T1) / 56.91 kg (GR1T2), SURVEY App. D


def test_mass_matrix_is_spd_and_matches_kinetic_energy(setup):
    model, ctl = setup
    po = PhysOracle(model, ctl, None, dtype=np.float64)
    bi = nominal_params(model, 1)["base_inertial"][0]
    nd = model["nd"]
    rng = np.random.default_rng(0)
    root, q, qd = _rand_state(rng, model, 10)
    for e in range(10):
        M, h, en = po.dynamics_terms(bi, root[e], q[e], qd[e])
        u = np.concatenate([qd[e], root[e, 7:13]])
        assert M.shape == (nd + 6, nd + 6) and np.abs(M - M.T).max() < 1e-11
        assert np.linalg.eigvalsh(M).min() > 0
        assert abs(0.5 * u @ M @ u - en[0]) < 1e-9 * max(1.0, en[0])
        assert abs(M[nd, nd] - np.sum(model["mass"])) < 1e-9


def test_free_flight_conserves_momentum_and_energy(setup):
    model, ctl = setup
    nd = model["nd"]
    par = nominal_params(model, 1)
    rng = np.random.default_rng(1)
    root0, q0, qd0 = _rand_state(rng, model, 1)
    qd0 *= 0.3
    free = dict(ctl, kp=np.zeros(nd), kd=np.zeros(nd))
    drifts = []
    for dt in (1e-3, 5e-4):
        po = PhysOracle(model, free, None, sim=dict(dt=dt, decimation=int(round(0.04 / dt))), dtype=np.float64)
        root, q, qd = root0.copy(), q0.copy(), qd0.copy()
        M, h, e0 = po.dynamics_terms(par["base_inertial"][0], root[0], q[0], qd[0])
        p0 = M[nd:nd + 3] @ np.concatenate([qd[0], root[0, 7:13]])
        po.step(root, q, qd, np.zeros((1, nd)), np.zeros((1, nd)), 0.0, np.ones((1, nd)), par["base_inertial"], par["friction"],
                par["restitution"])
        M, h, e1 = po.dynamics_terms(par["base_inertial"][0], root[0], q[0], qd[0])
        p1 = M[nd:nd + 3] @ np.concatenate([qd[0], root[0, 7:13]])
        mtot = float(np.sum(model["mass"]))
        assert np.allclose(p1 - p0, [0, 0, -9.81 * mtot * 0.04], atol=1e-4 * mtot)
        drifts.append(abs(e1.sum() - e0.sum()) / max(1.0, abs(e0[0])))
    assert drifts[0] < 3e-2 and drifts[1] < 0.6 * drifts[0] + 1e-9, drifts


def test_f32_tracks_f64_full_body(setup):
    model, ctl = setup
    nd, n = model["nd"], 4
    par = nominal_params(model, n)
    res = {}
    for dt in (np.float32, np.float64):
        po = PhysOracle(model, ctl, None, dtype=dt)
        root = np.zeros((n, 13), dt); root[:, 2] = 1.5; root[:, 6] = 1.0                 # in the air: articulated dynamics + PD only
        q = np.zeros((n, nd), dt); qd = np.zeros((n, nd), dt)
        act = (0.1 * np.random.default_rng(5).normal(size=(n, nd))).astype(dt)
        for _ in range(3):
            po.step(root, q, qd, act, act, 0.0, np.ones((n, nd), dt), par["base_inertial"], par["friction"], par["restitution"])
        res[dt] = (root.copy(), q.copy(), qd.copy())
    assert np.abs(res[np.float32][0][:, :7] - res[np.float64][0][:, :7]).max() < 5e-4
    assert np.abs(res[np.float32][1] - res[np.float64][1]).max() < 2e-3
