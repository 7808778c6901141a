"""GPU parity tests of the fused env kernel, through the C ABI (libgrx_b200.so -> grx_b200.env.GRXVecEnv).

 * post-physics half vs golden trajectories of the UNMODIFIED reference classes (reference-pinned, tight tolerance)
 * full step (our dynamics spec) vs the golden full steps, whose physics was the C oracle in fp32 (teacher-forced per step)
 * mass matrix / bias vector vs the fp64 oracle
 * size-independent properties at BASELINE.json's full size (4096 envs, rough heightfield)
"""
import numpy as np
import pytest
import torch

from golden_util import ENV_FIXTURES, FULL_BODY_FIXTURES, load_fixture, step_items
from parity_util import PHYS_FORCE_TOL, PHYS_TOL, PHYS_TORQUE_TOL, PHYS_VEL_TOL, POST_TOL, make_gpu_env, vel_cols

# (fixture, generic): every lower-limb fixture on the specialised fused kernel AND on the generic-topology kernels (GRX_ENV_GENERIC=1: same
# task code from grx_task.cuh, dynamics of grx_phys_generic.cu); the full-body 32-DOF fixtures (always generic)
ENV_CASES = [(n, False) for n in ENV_FIXTURES] + [(n, True) for n in ENV_FIXTURES] + [(n, True) for n in FULL_BODY_FIXTURES]
_case_id = lambda c: c[0] + ("-generic" if c[1] else "")

pytestmark = pytest.mark.gpu

CARRIED_F = ("root_states", "dof_pos", "dof_vel", "last_dof_vel", "last_actions", "last_last_actions", "commands",
             "base_heights_offset", "feet_air_time", "feet_land_time")


def _cmp_state(env, st, tol, msg):
    for k in CARRIED_F:
        got = getattr(env, k).cpu().numpy()
        np.testing.assert_allclose(got, st[k].reshape(got.shape), err_msg=f"{msg} state {k}", **tol)
    np.testing.assert_array_equal(env.feet_contact_last.cpu().numpy() != 0, st["feet_contact_last"].astype(bool), err_msg=msg)
    np.testing.assert_array_equal(env.episode_length_buf.cpu().numpy(), st["episode_length_buf"], err_msg=msg)
    np.testing.assert_allclose(env.episode_sums_buf.cpu().numpy(), st["episode_sums"], err_msg=f"{msg} episode_sums", rtol=tol["rtol"], atol=10 * tol["atol"])
    if "terrain_levels" in st and env.custom_origins:
        np.testing.assert_array_equal(env.terrain_levels.cpu().numpy(), st["terrain_levels"], err_msg=msg)
        np.testing.assert_allclose(env.env_origins.cpu().numpy(), st["env_origins"], err_msg=msg, **tol)


@pytest.mark.parametrize("case", ENV_CASES, ids=_case_id)
def test_post_physics_matches_reference(case):
    """obs / pri_obs / rew / reset / time_out / carried state / extras of the CUDA post-physics path == the unmodified
    reference GR1T1 / GR1T2 classes (golden), given the same physics outputs and the same random draws."""
    name, generic = case
    fx = load_fixture(name)
    env = make_gpu_env(fx, generic=generic, sync_extras=True)[0]
    assert env.generic == generic
    dev = env.device
    n_reset = 0
    for t in range(int(fx["meta/steps"])):
        pre = f"step{t:02d}/"
        ph = step_items(fx, t, "phys")
        env.root_states.copy_(torch.from_numpy(ph["root_states_phys"]).to(dev))
        env.dof_pos.copy_(torch.from_numpy(ph["dof_pos_phys"]).to(dev))
        env.dof_vel.copy_(torch.from_numpy(ph["dof_vel_phys"]).to(dev))
        inj = dict(torques=torch.from_numpy(ph["torques_phys"]), foot_state=torch.from_numpy(ph["foot_state"]),
                   torso_quat=torch.from_numpy(ph["torso_quat"]), contact_forces=torch.from_numpy(ph["contact_forces"]),
                   avg_foot_force=torch.from_numpy(ph["avg_feet_contact_force"]), avg_foot_linvel=torch.from_numpy(ph["avg_feet_speed_xyz"]))
        U = torch.from_numpy(fx[pre + "U"]).to(dev)
        obs, pri, rew, reset, extras = env.post_physics_injected(torch.from_numpy(fx[pre + "actions"]), U, inj)
        out, st = step_items(fx, t, "out"), step_items(fx, t, "state")
        msg = f"{name} t={t}"
        np.testing.assert_array_equal(reset.cpu().numpy(), out["reset_buf"].astype(bool), err_msg=msg)
        np.testing.assert_array_equal(env.time_out_buf.cpu().numpy(), out["time_out_buf"].astype(bool), err_msg=msg)
        np.testing.assert_allclose(obs.cpu().numpy(), out["obs_buf"], err_msg=msg + " obs", **POST_TOL)
        np.testing.assert_allclose(pri.cpu().numpy(), out["pri_obs_buf"], err_msg=msg + " pri_obs", **POST_TOL)
        np.testing.assert_allclose(rew.cpu().numpy(), out["rew_buf"], err_msg=msg + " rew", **POST_TOL)
        _cmp_state(env, st, POST_TOL, msg)
        if pre + "extras_episode" in fx:
            got = np.array([float(extras["episode"]["rew_" + n]) for n in env.reward_names], np.float32)
            np.testing.assert_allclose(got, fx[pre + "extras_episode"], rtol=1e-4, atol=1e-6, err_msg=msg + " extras")
            if pre + "extras_terrain_level" in fx:
                np.testing.assert_allclose(float(extras["episode"]["terrain_level"]), float(fx[pre + "extras_terrain_level"]), rtol=1e-6)
        n_reset += int(reset.sum())
    assert n_reset >= 3


def _full_step_compare(name, report, generic=False):
    """Teacher-forced whole fused step (PD torque -> dynamics -> contact -> integrate x decimation -> post-physics) vs the C oracle run
    live on the same golden state, with ACTIVE-SET ATTRIBUTION: both sides export, per env and substep, a hash of the discrete decisions
    of the dynamics (accepted contact spheres, terrain triangle under each, restitution branch, active joint limits).
      * envs whose signatures agree on every substep took the same decisions: every output must agree within the stated fp32 tolerance;
      * envs whose signatures differ made a different contact / limit decision at a threshold (legitimate under different rounding:
        Delassus-space vs velocity-space PGS, sparse vs dense Cholesky, FMA contraction): they are COUNTED and bounded, not compared.
    No env may differ without a differing signature."""
    from golden_util import init_state
    from parity_util import make_oracle_env
    fx = load_fixture(name)
    env = make_gpu_env(fx, generic=generic)[0]
    assert env.generic == generic
    ora = make_oracle_env(fx)
    sig = env.debug_active_sig(True)
    obs_vel, pri_vel = vel_cols(env.num_actions)
    dev, N, dec = env.device, env.num_envs, int(fx["meta/decimation"])
    n_flip = 0
    for t in range(int(fx["meta/steps"])):
        pre = f"step{t:02d}/"
        st = init_state(fx) if t == 0 else step_items(fx, t - 1, "state")
        env.load_state(st); ora.load_state(st)
        a, U, delay = fx[pre + "actions"], fx[pre + "U"], float(fx[pre + "delay"])
        obs, pri, rew, reset, _ = env.step(torch.from_numpy(a).to(dev), U=torch.from_numpy(U).to(dev), delay=delay)
        torch.cuda.synchronize()
        o_obs, o_pri, o_rew, o_reset, _ = ora.step(a, U, delay)
        same = (sig[:, :dec].cpu().numpy().view(np.uint64) == ora.last_active_sig).all(1)
        n_flip += int((~same).sum())
        msg = f"{name} t={t}"
        assert (~same).sum() <= max(1, N // 16), f"{msg}: {(~same).sum()} of {N} envs took a different contact/limit decision"
        for nm, got, ref, tol in (("torques", env.torques, ora.torques, PHYS_TORQUE_TOL), ("obs", obs, o_obs, PHYS_TOL), ("pri_obs", pri, o_pri, PHYS_TOL),
                                  ("rew", rew, o_rew, PHYS_TOL), ("root_states", env.root_states, ora.root_states, PHYS_TOL),
                                  ("dof_pos", env.dof_pos, ora.dof_pos, PHYS_TOL), ("dof_vel", env.dof_vel, ora.dof_vel, PHYS_VEL_TOL),
                                  ("contact_forces", env.contact_forces, ora.contact_forces, PHYS_FORCE_TOL)):
            g_, r_ = got.cpu().numpy()[same], np.asarray(ref.numpy() if hasattr(ref, "numpy") else ref)[same]
            if g_.size:
                err = np.abs(g_ - r_)
                report[nm] = max(report.get(nm, 0.0), float(err.max()))
                # worst error in units of the stated tolerance (<= 1 passes); all quantities are reported before anything fails
                bound = tol["atol"] + tol["rtol"] * np.abs(r_)
                vcols = {"obs": obs_vel, "pri_obs": pri_vel, "root_states": list(range(7, 13))}.get(nm)
                if vcols is not None:   # velocity columns of a mixed array carry the velocity tolerance
                    bound[:, vcols] = PHYS_VEL_TOL["atol"] + PHYS_VEL_TOL["rtol"] * np.abs(r_[:, vcols])
                report["tolfrac/" + nm] = max(report.get("tolfrac/" + nm, 0.0), float((err / bound).max()))
        np.testing.assert_array_equal(reset.cpu().numpy()[same], o_reset.numpy()[same], err_msg=msg)
        for tns in (obs, pri, rew, env.root_states):                                      # the flipped envs still hold sane values
            assert bool(torch.isfinite(tns).all())
    report["flipped_env_steps"] = n_flip
    report["env_steps"] = N * int(fx["meta/steps"])
    return report


@pytest.mark.parametrize("case", ENV_CASES, ids=_case_id)
def test_full_step_matches_oracle(case):
    name, generic = case
    rep = _full_step_compare(name, {}, generic)
    print(f"\n{name}: max |CUDA - C oracle| over envs with identical active sets: "
          + ", ".join(f"{k} {v:.2e} ({rep['tolfrac/' + k]:.2f} of tol)" for k, v in rep.items() if "/" not in k and k not in ("flipped_env_steps", "env_steps"))
          + f"; {rep['flipped_env_steps']} of {rep['env_steps']} env-steps took a different contact/limit decision")
    assert rep["flipped_env_steps"] <= rep["env_steps"] // 32
    bad = {k: v for k, v in rep.items() if k.startswith("tolfrac/") and v > 1.0}
    assert not bad, f"{name}: envs with IDENTICAL active sets differ beyond the stated tolerance: {bad}" 


def test_single_substep_error():
    """Per-SUBSTEP error (fixture plane64_dec1: decimation 1, one 2 ms substep per policy step) reported separately from the
    10-substep error above; stated bound 2e-4 abs on the state, an order of magnitude below the per-policy-step tolerance."""
    rep = _full_step_compare("plane64_dec1", {})
    print("\nsingle substep: " + ", ".join(f"{k} {v:.2e}" for k, v in rep.items() if "/" not in k))
    for k in ("root_states", "dof_pos"):
        assert rep[k] < 5e-5, (k, rep[k])               # measured 5e-6 / 3e-7
    assert rep["dof_vel"] < 1e-3 and rep["obs"] < 1e-3   # measured 1.7e-4 (joint rates and their observation columns)


@pytest.mark.parametrize("robot", ["GR1T1", "GR1T2"])
def test_mass_matrix_and_bias_match_fp64_oracle(robot):
    from grx_b200.config import make_cfg
    from grx_b200.env import GRXVecEnv
    from oracle.phys import PhysOracle
    cfg = make_cfg(robot, 8, "plane")
    env = GRXVecEnv(cfg, sim_device="cuda:0")
    g = torch.Generator().manual_seed(3)
    root = torch.zeros(8, 13)
    root[:, 2] = 1.0
    q = torch.randn(8, 4, generator=g)
    root[:, 3:7] = q / q.norm(dim=1, keepdim=True)
    root[:, 7:13] = torch.randn(8, 6, generator=g)
    dq = 0.3 * torch.randn(8, 10, generator=g)
    dqd = 2.0 * torch.randn(8, 10, generator=g)
    env.root_states.copy_(root.cuda()); env.dof_pos.copy_(dq.cuda()); env.dof_vel.copy_(dqd.cuda())
    torch.cuda.synchronize()
    ora = PhysOracle(env.model, env.tables, None, dtype=np.float64)
    for i in range(8):
        M, h = env.debug_dynamics(i)
        Mo, ho, _ = ora.dynamics_terms(env.params["base_inertial"][i], root[i].numpy(), dq[i].numpy(), dqd[i].numpy())
        np.testing.assert_allclose(M, Mo, rtol=2e-4, atol=2e-4)
        np.testing.assert_allclose(h, ho, rtol=2e-4, atol=2e-3)


@pytest.mark.parametrize("robot,N,mesh", [("GR1T1", 4096, "heightfield"),     # BASELINE config #2
                                          ("GR1T2", 8192, "heightfield"),     # config #3: GR1T2 + full domain randomisation
                                          ("GR1T1", 4096, "trimesh")])        # config #5: one rank's share of the trimesh + curriculum job
def test_full_size_properties(robot, N, mesh):
    """BASELINE config shapes at full per-GPU size: rough terrain, curriculum, domain randomisation, fast-mode (in-kernel Philox)
    draws; size-independent properties instead of an oracle run."""
    from grx_b200.config import make_cfg
    from grx_b200.env import GRXVecEnv
    cfg = make_cfg(robot, N, mesh)
    env = GRXVecEnv(cfg, sim_device="cuda:0")
    obs, pri = env.reset()
    assert obs.shape == (N, 39) and pri.shape == (N, 168)
    g = torch.Generator(device="cuda").manual_seed(0)
    n_reset, n_steps = 0, 150
    lvl0 = env.terrain_levels.clone()
    for _ in range(n_steps):
        a = 0.3 * torch.randn(N, 10, device="cuda", generator=g)
        obs, pri, rew, reset, extras = env.step(a)
        n_reset += int(reset.sum())
    torch.cuda.synchronize()
    for t in (obs, pri, rew, env.root_states, env.dof_pos, env.dof_vel):
        assert torch.isfinite(t).all()
    assert obs.abs().max() <= 100.0 and pri.abs().max() <= 100.0                       # clip_observations
    assert 0 < n_reset < N * n_steps // 4                                           # random actions: robots fall, but not at once
    assert torch.allclose(pri[:, 3:9], obs[:, 3:9], atol=0.06)                         # obs = pri[:39] + bounded noise
    noise = (obs - pri[:, :39]).abs().max(0).values.cpu().numpy()
    bound = np.array([0] * 3 + [0.05] * 3 + [0.03] * 3 + [0.04] * 10 + [0.2] * 10 + [0] * 10) + 1e-6
    assert (noise <= bound).all() and noise[3:29].min() > 0.0
    q = env.root_states[:, 3:7]
    assert torch.allclose(q.norm(dim=1), torch.ones(N, device="cuda"), atol=1e-4)   # integrator keeps the quaternion unit
    lim_lo = torch.tensor(env.model["dof_lower"], device="cuda", dtype=torch.float32) - 0.05
    lim_hi = torch.tensor(env.model["dof_upper"], device="cuda", dtype=torch.float32) + 0.05
    inside = ((env.dof_pos >= lim_lo) & (env.dof_pos <= lim_hi)).all(1)
    assert inside.float().mean() > 0.95                                                # joint-limit rows hold (reset draws may start outside)
    assert (env.terrain_levels != lvl0).any()                                          # curriculum moved somebody
    ep = extras["episode"]
    assert set(ep.keys()) == {"rew_" + n for n in env.reward_names} | {"terrain_level"}
    assert all(torch.isfinite(v) for v in ep.values())


def test_standing_contact_force_balances_weight():
    """Physical envelope: a robot holding its default pose on the plane settles with sum F_z ~= m g."""
    from grx_b200.config import make_cfg
    from grx_b200.env import GRXVecEnv
    from grx_b200.robot import nominal_params
    from grx_b200.urdf import builtin_model
    cfg = make_cfg("GR1T1", 16, "plane")
    cfg.noise.add_noise = False
    cfg.domain_rand.push_robots = cfg.domain_rand.randomize_init_dof_pos = cfg.domain_rand.randomize_init_base_velocity = False
    env = GRXVecEnv(cfg, sim_device="cuda:0", params=nominal_params(builtin_model("GR1T1"), 16))
    env.reset()
    for _ in range(50):
        env.step(torch.zeros(16, 10, device="cuda"), delay=0.0)
    torch.cuda.synchronize()
    fz = env.contact_forces[:, :, 2].sum(1).cpu().numpy()
    mg = float(env.model["mass"][1:].sum() + env.params["base_inertial"][0, 0]) * 9.81
    alive = ~env.reset_buf.cpu().numpy()
    assert alive.mean() > 0.7
    np.testing.assert_allclose(fz[alive], mg, rtol=0.15)
    assert abs(mg - 52.83 * 9.81) < 1.0                                               # GR1T1 mass


def test_play_py_surface():
    """Attributes legged_gym/scripts/play.py reads while logging (play.py:110-123): env.dt, dof_pos, dof_vel, torques, commands,
    base_lin_vel, base_ang_vel, contact_forces[:, feet_indices, 2], cfg.control.action_scale; obs after reset come from get_observations."""
    from grx_b200.config import make_cfg
    from grx_b200.env import GRXVecEnv
    cfg = make_cfg("GR1T1", 16, "plane")
    env = GRXVecEnv(cfg, sim_device="cuda:0")
    env.reset()
    obs = env.get_observations()
    for _ in range(5):
        obs, _, rew, dones, infos = env.step(torch.zeros(16, 10, device="cuda"))
    torch.cuda.synchronize()
    assert abs(env.dt - 0.02) < 1e-9 and env.cfg.control.action_scale == 1.0
    for name, shape in (("dof_pos", (16, 10)), ("dof_vel", (16, 10)), ("torques", (16, 10)), ("commands", (16, 3)),
                        ("base_lin_vel", (16, 3)), ("base_ang_vel", (16, 3)), ("projected_gravity", (16, 3))):
        t = getattr(env, name)
        assert tuple(t.shape) == shape and bool(torch.isfinite(t).all()), name
    fz = env.contact_forces[0, env.feet_indices, 2]
    assert fz.shape == (2,) and float(env.dof_pos[0, 3].item()) == float(env.dof_pos[0, 3])
    # obs[3:6] is the base angular velocity the kernel computed one step earlier than the live root state; the gravity direction
    # of an upright robot is (0, 0, -1) in both
    up = env.projected_gravity[:, 2] < -0.9
    assert bool(up.any()) and torch.allclose(env.projected_gravity.norm(dim=1), torch.ones(16, device="cuda"), atol=1e-4)


def test_trimesh_entry_accepts_only_the_structured_conversion():
    """grx_env_set_terrain_trimesh (gym.add_triangle_mesh, legged_robot.py:903-924): the mesh must be the reference's structured
    conversion of the heightfield (terrain_utils.py:286-350); anything else is rejected with GRX_E_INVALID."""
    import ctypes as C
    from grx_b200 import _lib as L
    from grx_b200.config import make_cfg
    from grx_b200.env import GRXVecEnv
    from grx_b200.terrain import heightfield_to_trimesh
    cfg = make_cfg("GR1T1", 16, "trimesh")
    cfg.terrain.num_rows, cfg.terrain.num_cols, cfg.terrain.max_init_terrain_level = 2, 2, 1
    env = GRXVecEnv(cfg, sim_device="cuda:0", trimesh_builder="host")   # goes through the vertex / triangle entry with the Terrain's own mesh
    env.reset()
    for _ in range(3):
        env.step(torch.zeros(16, 10, device="cuda"))
    torch.cuda.synchronize()
    assert bool(torch.isfinite(env.obs_buf).all())
    hs = np.ascontiguousarray(env.terrain.heightsamples, np.int16)
    v, t = heightfield_to_trimesh(hs, 0.1, 0.005, 0.75)
    args = lambda vv, tt: (env._h, vv.ctypes.data_as(L.PF), vv.shape[0], tt.ctypes.data_as(C.POINTER(C.c_uint32)), tt.shape[0],
                           hs.ctypes.data_as(C.POINTER(C.c_int16)), hs.shape[0], hs.shape[1], C.c_float(0.1), C.c_float(0.005),
                           C.c_float(25.0), C.c_float(1.0), C.c_float(0.0))
    assert env.lib.grx_env_set_terrain_trimesh(*args(v, t)) == 0
    assert env.lib.grx_env_set_terrain_trimesh(*args(v[:-1].copy(), t)) == -1                     # wrong vertex count
    bad = v.copy(); bad[5, 2] += 0.5                                                              # a vertex off the heightfield
    assert env.lib.grx_env_set_terrain_trimesh(*args(bad, t)) == -1
    assert b"does not match" in env.lib.grx_last_error()
    bad = v.copy(); bad[7, 0] += 0.04                                                             # a non-integral sideways shift
    assert env.lib.grx_env_set_terrain_trimesh(*args(bad, t)) == -1
    badt = t.copy(); badt[len(t) // 2, [1, 2]] = badt[len(t) // 2, [2, 1]]                        # one triangle in the MIDDLE with swapped indices
    assert env.lib.grx_env_set_terrain_trimesh(*args(v, badt)) == -1 and b"structured pair" in env.lib.grx_last_error()
    assert env.lib.grx_env_set_terrain_trimesh(*args(v, t)) == 0


def test_device_trimesh_builder_equals_reference_conversion():
    """grx_env_set_terrain_trimesh_hf: the steep-edge vertex snapping computed ON THE DEVICE from the int16 sample grid (one thread per
    vertex) gives exactly the shifts of the reference's numpy conversion (terrain_utils.py:315-328 via grx_b200.terrain, SHA-pinned to the
    reference), on a full 10 x 20-tile rough terrain; and the env built on it steps on the mesh's top surface: on a stairs tile the
    contact height under a robot standing next to a riser is a flat tread, not the heightfield's ramp."""
    from grx_b200.config import make_cfg
    from grx_b200.env import GRXVecEnv
    from grx_b200.terrain import heightfield_to_trimesh
    cfg = make_cfg("GR1T1", 64, "trimesh")
    env = GRXVecEnv(cfg, sim_device="cuda:0")                                       # device builder (default)
    hs = np.ascontiguousarray(env.terrain.heightsamples, np.int16)
    verts, _ = heightfield_to_trimesh(hs, cfg.terrain.horizontal_scale, cfg.terrain.vertical_scale, cfg.terrain.slope_treshold)
    R, Cc = hs.shape
    v = verts.astype(np.float64).reshape(R, Cc, 3)
    gi, gj = np.meshgrid(np.arange(R), np.arange(Cc), indexing="ij")
    want = np.stack([np.rint(v[..., 0] / cfg.terrain.horizontal_scale - gi), np.rint(v[..., 1] / cfg.terrain.horizontal_scale - gj)], -1).astype(np.int8)
    got = env._view("terrain_moves").cpu().numpy().view(np.int8)
    assert got.shape == want.shape and int((want != 0).sum()) > 1000                  # stairs / obstacles do shift vertices
    np.testing.assert_array_equal(got, want)
    env.reset()
    for _ in range(5):
        env.step(torch.zeros(64, 10, device="cuda"))
    torch.cuda.synchronize()
    assert bool(torch.isfinite(env.obs_buf).all())


def test_step_host_equals_device_step():
    """grx_env_step_host (H2D actions -> fused step -> D2H obs / privileged obs / rewards / resets -> stream sync, the host-buffer
    entry a non-torch caller of VecEnv.step uses, timed by bench.py as e2e.env_step_host) returns exactly what the device-pointer
    entry grx_env_step leaves in the device buffers for the same state, actions, delay and step index."""
    import ctypes as C
    from grx_b200 import _lib as L
    from grx_b200.config import make_cfg
    from grx_b200.env import GRXVecEnv
    N = 300
    envs = [GRXVecEnv(make_cfg("GR1T1", N, "plane"), sim_device="cuda:0") for _ in range(2)]
    for e in envs:
        e.reset()
    g = torch.Generator().manual_seed(6)
    h_o, h_p = np.zeros((N, 39), np.float32), np.zeros((N, 168), np.float32)
    h_r, h_d = np.zeros(N, np.float32), np.zeros(N, np.uint8)
    P = lambda a: a.ctypes.data_as(C.c_void_p)
    for t in range(12):
        a = (0.3 * torch.randn(N, 10, generator=g)).numpy()
        idx = 1000 + t
        L.check(envs[0].lib.grx_env_step_host(envs[0]._h, P(a), C.c_float(3.0), 0, C.c_uint64(idx), P(h_o), P(h_p), P(h_r), P(h_d), None))
        L.check(envs[1].lib.grx_env_step(envs[1]._h, C.c_void_p(torch.from_numpy(a).cuda().data_ptr()), None, C.c_float(3.0), 0, C.c_uint64(idx), None))
        torch.cuda.synchronize()
        np.testing.assert_array_equal(h_o, envs[1].obs_buf.cpu().numpy())
        np.testing.assert_array_equal(h_p, envs[1].pri_obs_buf.cpu().numpy())
        np.testing.assert_array_equal(h_r, envs[1].rew_buf.cpu().numpy())
        np.testing.assert_array_equal(h_d, envs[1]._reset_u8.cpu().numpy())
    assert np.isfinite(h_o).all() and np.abs(h_o).max() > 0


def test_compat_exports_rigid_body_states_dof_state_episode_length():
    """SURVEY §8(b) lists rigid_body_states / dof_state among the tensors the boundary exports (legged_robot.py:110-135).  They are compat
    exports (off until first requested): rigid_body_states [N, 37, 13] must agree with the C oracle's link states of the same post-physics
    state and contain the feet rows the kernel reports separately; dof_state is the interleaved (pos, vel) mirror; episode_length_buf is
    int64 like the reference buffer (base_task.py:71-72) and follows the live counter."""
    from grx_b200.config import make_cfg
    from grx_b200.env import GRXVecEnv
    from grx_b200.robot import nominal_params
    from grx_b200.urdf import builtin_model
    from oracle.phys import PhysOracle
    N = 64
    cfg = make_cfg("GR1T1", N, "plane")
    cfg.noise.add_noise = False
    model = builtin_model("GR1T1")
    env = GRXVecEnv(cfg, sim_device="cuda:0", params=nominal_params(model, N))
    env.reset()
    rbs, ds, ep = env.rigid_body_states, env.dof_state, env.episode_length_buf        # first access switches the exports on
    assert rbs.shape == (N, len(model["link_names"]), 13) and ds.shape == (N, 10, 2) and ep.dtype == torch.int64
    g = torch.Generator().manual_seed(2)
    for t in range(6):
        pre = ep.clone()
        obs, pri, rew, reset, _ = env.step((0.2 * torch.randn(N, 10, generator=g)).cuda(), delay=2.0)
        torch.cuda.synchronize()
        alive = ~reset
        assert bool((ep[alive] == pre[alive] + 1).all()) and bool((ep[reset] == 0).all())       # int64 mirror follows the counter
        assert torch.equal(ep, env._episode_length.to(torch.int64))
        assert torch.equal(ds[..., 0], env.dof_pos) and torch.equal(ds[..., 1], env.dof_vel)
        np.testing.assert_allclose(rbs[:, env.feet_indices].cpu().numpy(), env.foot_state.cpu().numpy(), rtol=1e-6, atol=1e-6)
        # against the oracle's forward kinematics of the same state (envs that did not reset: their records still hold the post-physics state)
        ora = PhysOracle(model, env.tables, None, dtype=np.float64)
        a = alive.cpu().numpy()
        ls = ora.link_states(env.root_states.cpu().numpy()[a], env.dof_pos.cpu().numpy()[a], env.dof_vel.cpu().numpy()[a], env.params["base_inertial"][a])
        got = rbs.cpu().numpy()[a]
        np.testing.assert_allclose(got[..., :3], ls[..., :3], rtol=0, atol=2e-5)
        np.testing.assert_allclose(got[..., 7:], ls[..., 7:], rtol=1e-4, atol=2e-4)
        dot = np.abs((got[..., 3:7] * ls[..., 3:7]).sum(-1))                          # same rotation up to quaternion sign
        assert (dot > 1 - 1e-5).all()
