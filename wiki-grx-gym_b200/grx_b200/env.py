"""GRXVecEnv — the reference's VecEnv surface (rsl_rl/rsl_rl/env/vec_env.py:7-40 as concretised by
legged_gym/legged_gym/envs/base/base_task.py:39-121 and legged_robot.py:53-246) over the fused B200 env kernel.

Every public tensor is a zero-copy (possibly strided) view of device memory owned by libgrx_b200.so, the same
contract gymtorch.wrap_tensor gives (gymtorch.py:61-95).  One ``step()`` is ONE kernel launch on torch's current
stream; nothing in it synchronises with the host (the reference syncs 3x per step: legged_robot.py:292, 317, 387).

There is no CPU fallback: without the CUDA library and a CUDA device the constructor raises.
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np
import torch

from . import _lib as L
from . import rng_layout as RL
from .config import make_cfg
from . import sharding
from .robot import nominal_params, sample_domain_rand, self_collision_pairs, task_tables
from .terrain import DeviceTerrain, Terrain
from .urdf import builtin_model

REWARD_NAMES = list(L._SIGMAS[:21]) + ["on_the_air", "pose_offset", "stand_still"]   # alphabetical
assert REWARD_NAMES == sorted(REWARD_NAMES) and len(REWARD_NAMES) == 24

# grx_buffer.dtype (GymTensorDataType codes + Int32 / Int64) -> numpy typestr; u64 buffers are exposed as int64 bit patterns (torch has no uint64 arithmetic)
_TYPESTR = {L.DT_F32: "<f4", L.DT_U32: "<i4", L.DT_U64: "<i8", L.DT_U8: "|u1", L.DT_I16: "<i2", L.DT_I32: "<i4", L.DT_I64: "<i8"}
_ITEMSIZE = {L.DT_F32: 4, L.DT_U32: 4, L.DT_U64: 8, L.DT_U8: 1, L.DT_I16: 2, L.DT_I32: 4, L.DT_I64: 8}


class _DevArray:
    """__cuda_array_interface__ holder so torch can alias a raw device pointer (role of gymtorch.wrap_tensor)."""

    def __init__(self, buf: L.Buffer, owner):
        nd = buf.ndim
        item = _ITEMSIZE[buf.dtype]
        self.owner = owner
        self.__cuda_array_interface__ = {
            "shape": tuple(int(buf.dims[i]) for i in range(nd)),
            "strides": tuple(int(buf.strides[i]) * item for i in range(nd)),
            "typestr": _TYPESTR[buf.dtype], "data": (int(buf.data), False), "version": 2}


def _f32p(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a, a.ctypes.data_as(L.PF)


def _i32p(a):
    a = np.ascontiguousarray(a, dtype=np.int32)
    return a, a.ctypes.data_as(L.PI)


def model_desc(model, tables):
    """Flatten (model, task tables) into the C ``grx_model_desc`` (keeps the numpy arrays alive in ``keep``)."""
    keep, d = [], L.ModelDesc()
    order = np.asarray(tables["sph_order"])
    d.nb, d.nd, d.nl, d.ns = model["nb"], model["nd"], len(model["link_names"]), len(model["sph_rad"])
    d.nf, d.nterm, d.nankle = len(tables["foot_links"]), len(tables["termination_links"]), len(tables["ankle_dofs"])

    def F(name, arr):
        a, p = _f32p(arr); keep.append(a); setattr(d, name, p)

    def I(name, arr):
        a, p = _i32p(arr); keep.append(a); setattr(d, name, p)
    I("parent", model["parent"])
    for k in ("jpos", "jrot", "axis", "mass", "com", "inertia", "dof_lower", "dof_upper", "dof_effort", "link_pos", "link_rot"):
        F(k, model[k])
    F("dof_vel_limit", model["dof_velocity"])
    F("soft_lower", tables["soft_lower"]); F("soft_upper", tables["soft_upper"])
    F("kp", tables["kp"]); F("kd", tables["kd"]); F("default_pos", tables["default_pos"])
    I("link_body", model["link_body"])
    I("sph_body", model["sph_body"][order]); I("sph_link", model["sph_link"][order])
    F("sph_pos", model["sph_pos"][order]); F("sph_rad", model["sph_rad"][order])
    I("foot_links", tables["foot_links"]); I("term_links", tables["termination_links"]); I("ankle_dofs", tables["ankle_dofs"])
    d.torso_link = int(tables["torso_links"][0])
    return d, keep


def task_cfg(cfg, tables, seed=1, env_id_offset=0):
    """The C ``grx_task_cfg`` from a task config (ours or the reference's cfg object): what _parse_cfg
    (legged_robot.py:91-104), _prepare_reward_function (:840-866) and compute_noise_scale_vec_profile
    (gr1t1.py:315-336) derive at construction."""
    t = L.TaskCfg()
    # options of the reference the fused kernel does not implement must not be silently ignored
    if getattr(cfg.control, "control_type", "P") != "P":
        raise L.GrxError(f"control_type={cfg.control.control_type!r}: the fused kernel implements the 'P' law of the registered GRx tasks (legged_robot.py:691-700)")
    if getattr(cfg.rewards, "only_positive_rewards", False):
        raise L.GrxError("rewards.only_positive_rewards=True is not implemented (registered GRx tasks: False, gr1t1_config.py:188)")
    px = cfg.sim.physx
    t.sim_dt, t.gravity = cfg.sim.dt, cfg.sim.gravity[2]
    t.contact_offset, t.bounce_threshold, t.max_depen_vel = px.contact_offset, px.bounce_threshold_velocity, px.max_depenetration_velocity
    t.erp, t.solver_iters, t.decimation = 0.2, px.num_position_iterations, cfg.control.decimation
    t.action_scale = cfg.control.action_scale
    nd = len(tables["kp"])
    H = len(cfg.terrain.measured_points_x) * len(cfg.terrain.measured_points_y)
    t.num_obs, t.num_pri_obs, t.num_actions, t.num_height_points = cfg.env.num_obs, cfg.env.num_pri_obs, cfg.env.num_actions, H
    nz = cfg.normalization
    for i in range(nd):
        t.clip_actions_min[i], t.clip_actions_max[i] = float(nz.clip_actions_min[i]), float(nz.clip_actions_max[i])
    t.clip_observations = nz.clip_observations
    dt = cfg.control.decimation * cfg.sim.dt
    t.max_episode_length = float(np.ceil(cfg.env.episode_length_s / dt))
    t.max_episode_length_s = cfg.env.episode_length_s
    t.resample_interval = int(cfg.commands.resampling_command_interval_s / dt)
    r = cfg.commands.ranges
    for i, rng in enumerate((r.lin_vel_x, r.lin_vel_y, r.ang_vel_yaw)):
        t.cmd_range[i][0], t.cmd_range[i][1] = rng[0], rng[1]
    dr = cfg.domain_rand
    t.max_push_vel_xy = dr.max_push_vel_xy
    rough = cfg.terrain.mesh_type in ("heightfield", "trimesh")
    t.add_noise = int(cfg.noise.add_noise)
    t.randomize_init_dof_pos, t.randomize_init_base_velocity = int(dr.randomize_init_dof_pos), int(dr.randomize_init_base_velocity)
    t.curriculum, t.custom_origins = int(bool(cfg.terrain.curriculum) and rough), int(rough)
    t.measure_heights = int(cfg.terrain.measure_heights)
    ns, os_, nl = cfg.noise.noise_scales, nz.obs_scales, cfg.noise.noise_level
    nv = np.zeros(9 + 3 * nd)
    nv[3:6], nv[6:9] = ns.ang_vel * nl * os_.ang_vel, ns.gravity * nl * os_.gravity
    nv[9:9 + nd], nv[9 + nd:9 + 2 * nd] = ns.dof_pos * nl * os_.dof_pos, ns.dof_vel * nl * os_.dof_vel
    nv[9 + 2 * nd:] = ns.action * nl * os_.action
    for i, v in enumerate(nv):
        t.noise_scale_vec[i] = v
    t.obs_scale_lin_vel, t.obs_scale_ang_vel, t.obs_scale_gravity = os_.lin_vel, os_.ang_vel, os_.gravity
    t.obs_scale_dof_pos, t.obs_scale_dof_vel, t.obs_scale_action = os_.dof_pos, os_.dof_vel, os_.action
    t.obs_scale_height = os_.height_measurements
    ini = cfg.init_state
    for i, v in enumerate(list(ini.pos) + list(ini.rot) + list(ini.lin_vel) + list(ini.ang_vel)):
        t.base_init_state[i] = v
    for i, v in enumerate(cfg.terrain.measured_points_x):
        t.measured_points_x[i] = v
    for i, v in enumerate(cfg.terrain.measured_points_y):
        t.measured_points_y[i] = v
    t.n_points_x, t.n_points_y = len(cfg.terrain.measured_points_x), len(cfg.terrain.measured_points_y)
    t.terrain_env_length = getattr(cfg.terrain, "terrain_length", 8.0)
    rw = cfg.rewards
    active = [n for n in sorted(k for k in dir(rw.scales) if not k.startswith("_") and k != "to_dict")
              if getattr(rw.scales, n) != 0 and n != "termination"]
    if active != REWARD_NAMES:
        raise L.GrxError(f"the fused kernel implements exactly the 24 reward terms of the registered GRx tasks; cfg enables {active}")
    for i, n in enumerate(REWARD_NAMES):
        t.reward_scale[i] = getattr(rw.scales, n) * dt                                 # legged_robot.py:849-850
    for k in ("base_height_target", "swing_feet_height_target", "feet_stumble_ratio", "feet_air_time_target",
              "feet_land_time_max", "soft_dof_vel_limit", "soft_torque_limit"):
        setattr(t, k, getattr(rw, k))
    for n in L._SIGMAS:
        setattr(t, "sigma_" + n, getattr(rw, "sigma_" + n))
    t.seed, t.env_id_offset = seed, env_id_offset
    return t


def quat_rotate_inverse(q, v):
    """isaacgym/torch_utils.py:72-81 (xyzw quaternions): rotate world-frame vectors v [N, 3] into the frames q [N, 4]."""
    qw, qv = q[:, 3:4], q[:, :3]
    a = v * (2.0 * qw ** 2 - 1.0)
    b = torch.cross(qv, v, dim=-1) * qw * 2.0
    c = qv * (qv * v).sum(-1, keepdim=True) * 2.0
    return a - b + c


class _EpisodeInfo(dict):
    """extras["episode"] (legged_robot.py:420-427) evaluated lazily from the device accumulator ring: building 25 0-d tensors eagerly
    would cost 25+ launches per step for values that are read once per iteration.  On a step where nothing reset, the reference leaves
    the PREVIOUS non-empty dict in extras (reset_idx returns early, legged_robot.py:387-388); evaluating lazily, the same is obtained by
    walking the ring back from this launch's slot to the most recent slot with a non-zero reset count (ring = 256 launches deep)."""

    def __init__(self, ring, slot, names, inv_len_s, num_envs_total, curriculum):
        super().__init__()
        self._ring, self._slot, self._names, self._s, self._n = ring, slot, names, inv_len_s, num_envs_total
        keys = ["rew_" + n for n in names] + (["terrain_level"] if curriculum else [])
        for k in keys:
            dict.__setitem__(self, k, None)
        self._done = False

    def _fill(self):
        if not self._done:
            ring = self._ring.clone()
            n = ring.shape[0]
            order = (self._slot - torch.arange(n, device=ring.device)) % n            # this slot, the previous one, ...
            cnts = ring[order, 24]
            nz = torch.nonzero(cnts > 0)
            src = ring[order[nz[0, 0]]] if nz.numel() else ring[self._slot]
            cnt = torch.clamp(src[24], min=1.0)
            for i, nme in enumerate(self._names):
                dict.__setitem__(self, "rew_" + nme, src[i] / cnt * self._s)
            if "terrain_level" in self.keys():
                dict.__setitem__(self, "terrain_level", ring[self._slot][25] / self._n)   # computed on every step (legged_robot.py:427 runs with the resets; the level sum is over ALL envs)
            self._done = True

    def __getitem__(self, k):
        self._fill()
        return dict.__getitem__(self, k)

    def items(self):
        self._fill()
        return dict.items(self)

    def values(self):
        self._fill()
        return dict.values(self)


class GRXVecEnv:
    """Drop-in for ``GR1T1(cfg, sim_params, physics_engine, sim_device, headless)`` as seen by ``OnPolicyRunner`` / play.py."""

    def __init__(self, cfg=None, sim_params=None, physics_engine=None, sim_device="cuda:0", headless=True, *,
                 rank=0, world_size=1, params=None, terrain=None, env_origins=None, terrain_levels=None, terrain_types=None,
                 parity_rng=False, sync_extras=False, trimesh_builder="device", self_collision=None,
                 max_self_contacts=4, terrain_generator="device"):
        """cfg: grx_b200.config.make_cfg(...) or the reference's GR1T1LowerLimbCfg()/GR1T2LowerLimbCfg() object
        (cfg.env.num_envs is the GLOBAL env count; this rank simulates the contiguous block rank*N/W..(rank+1)*N/W).
        A full-body cfg (grx_b200.config.make_full_body_cfg: 32 DOF, num_obs 105) runs on the generic-topology kernels behind the same
        calls, with robot self-collision unless cfg.asset.self_collisions != 0 or self_collision=False.
        params / terrain / env_origins...: override the sampled per-env parameters (tests)."""
        if cfg is None:
            cfg = make_cfg("GR1T1")
        if not torch.cuda.is_available():
            raise L.GrxError("GRXVecEnv needs a CUDA device (no CPU fallback)")
        self.lib = L.lib()
        self.cfg = cfg
        self.headless = headless
        self.device = torch.device(sim_device)
        self.rank, self.world_size = rank, world_size
        n_total = int(cfg.env.num_envs)
        assert n_total % world_size == 0, "num_envs must divide evenly over ranks"
        self.num_envs_total, self.num_envs = n_total, n_total // world_size
        self.env_id_offset = sharding.env_block(rank, world_size, n_total)[0]
        self.num_obs, self.num_pri_obs, self.num_actions = cfg.env.num_obs, cfg.env.num_pri_obs, cfg.env.num_actions
        self.num_privileged_obs = self.num_pri_obs
        robot = getattr(cfg, "robot", None) or ("GR1T2" if "GR1T2" in getattr(cfg.asset, "file", getattr(cfg.asset, "name", "")) else "GR1T1")
        self.model = builtin_model(robot)
        self.tables = task_tables(self.model, cfg)
        self.sim_dt = cfg.sim.dt
        self.dt = cfg.control.decimation * cfg.sim.dt                                  # legged_robot.py:92
        self.max_episode_length_s = cfg.env.episode_length_s
        self.max_episode_length = np.ceil(self.max_episode_length_s / self.dt)         # legged_robot.py:101 (numpy float)
        self.push_interval = np.ceil(cfg.domain_rand.push_interval_s / self.dt)        # legged_robot.py:103
        rough = cfg.terrain.mesh_type in ("heightfield", "trimesh")
        if not rough:
            cfg.terrain.curriculum = False                                              # legged_robot.py:97-98
        self.custom_origins = rough
        seed = int(getattr(cfg, "seed", 1))
        self.parity_rng, self.sync_extras = parity_rng, sync_extras
        self.reward_names = list(REWARD_NAMES)
        N = self.num_envs

        # ---- terrain (legged_robot.py:513-526) — every rank builds the same grid from the same numpy seed
        self.terrain = None
        if rough and terrain is None:
            st = np.random.get_state()
            np.random.seed(seed)
            # "device": the grid is generated by one kernel (grx_terrain_generate) and handed to the env without a host round trip;
            # "host": the numpy generator (bit-identical arrays, tests/test_terrain_gpu.py)
            if terrain_generator == "device":
                self.terrain = DeviceTerrain(cfg.terrain, n_total, self.device)
                terrain = dict(heights=None, heights_dev=self.terrain.heights_dev, terrain_origins=self.terrain.env_origins)
            else:
                self.terrain = Terrain(cfg.terrain, n_total)
                terrain = dict(heights=self.terrain.heightsamples, terrain_origins=self.terrain.env_origins)
            np.random.set_state(st)
        # ---- env origins (legged_robot.py:1163-1195), keyed by GLOBAL env index
        g = np.random.default_rng(seed + 7919)
        if rough:
            t_org = np.asarray(terrain["terrain_origins"], np.float32)
            rows, cols = t_org.shape[:2]
            if terrain_levels is None:
                max_init = cfg.terrain.max_init_terrain_level if cfg.terrain.curriculum else rows - 1
                terrain_levels = g.integers(0, max_init + 1, n_total)[self.env_id_offset:self.env_id_offset + N]
            if terrain_types is None:
                terrain_types = sharding.terrain_types_for(rank, world_size, n_total, cols)
            if env_origins is None:
                env_origins = t_org[np.asarray(terrain_levels), np.asarray(terrain_types)]
        elif env_origins is None:
            ncol = np.floor(np.sqrt(n_total))
            gi = np.arange(self.env_id_offset, self.env_id_offset + N)
            env_origins = np.zeros((N, 3), np.float32)
            env_origins[:, 0] = cfg.env.env_spacing * (gi // ncol)
            env_origins[:, 1] = cfg.env.env_spacing * (gi % ncol)
        # ---- per-env physical parameters (legged_robot.py:538-648, 1060-1064)
        if params is None:
            full = sample_domain_rand(self.model, cfg, n_total, np.random.default_rng(seed + 104729))
            params = {k: v[self.env_id_offset:self.env_id_offset + N] for k, v in full.items()}
        self.params = params

        # ---- C objects
        self._md, self._keep = model_desc(self.model, self.tables)
        self._tc = task_cfg(cfg, self.tables, seed=seed, env_id_offset=self.env_id_offset)
        self._h = C.c_void_p()
        dev_index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        L.check(self.lib.grx_env_create(C.byref(self._md), C.byref(self._tc), N, dev_index, C.byref(self._h)))
        self.generic = bool(self.lib.grx_env_info(self._h, 4))        # generic-topology kernels (full-body 32-DOF trees; GRX_ENV_GENERIC=1)
        self.rng_k = int(self.lib.grx_env_info(self._h, 0))          # width of a parity-mode uniform-draw row (rng_layout.layout(nd).K)
        self.self_collision_pairs = np.zeros((0, 2), np.int32)
        if self.generic and self_collision is not False and int(getattr(cfg.asset, "self_collisions", 0)) == 0:   # legged_robot_config.py:121: 0 = enabled
            self.self_collision_pairs = self_collision_pairs(self.model, self.tables)
            pr = np.ascontiguousarray(self.self_collision_pairs, np.int32)
            L.check(self.lib.grx_env_set_self_collision(self._h, pr.ctypes.data_as(L.PI), len(pr), int(max_self_contacts)))
        tc = cfg.terrain
        if rough and terrain.get("heights_dev") is not None and (tc.mesh_type != "trimesh" or trimesh_builder == "device"):
            hd = terrain["heights_dev"]
            L.check(self.lib.grx_env_set_terrain_device(self._h, C.c_void_p(hd.data_ptr()), hd.shape[0], hd.shape[1], C.c_float(tc.horizontal_scale),
                                                        C.c_float(tc.vertical_scale), C.c_float(tc.border_size),
                                                        C.c_float(tc.slope_treshold if tc.mesh_type == "trimesh" else -1.0),
                                                        C.c_float(tc.static_friction), C.c_float(tc.restitution)))
        elif rough:
            hs = np.ascontiguousarray(terrain["heights"] if terrain.get("heights") is not None else terrain["heights_dev"].cpu().numpy(), np.int16)
            scal = (C.c_float(tc.horizontal_scale), C.c_float(tc.vertical_scale), C.c_float(tc.border_size),
                    C.c_float(tc.static_friction), C.c_float(tc.restitution))
            if tc.mesh_type == "trimesh":                                               # legged_robot.py:903-924 (_create_trimesh)
                if trimesh_builder == "device":
                    # the steep-edge vertex snapping (terrain_utils.py:315-328) runs as a kernel over the sample grid: no host mesh
                    L.check(self.lib.grx_env_set_terrain_trimesh_hf(self._h, hs.ctypes.data_as(C.POINTER(C.c_int16)), hs.shape[0], hs.shape[1],
                                                                    scal[0], scal[1], scal[2], C.c_float(tc.slope_treshold), scal[3], scal[4]))
                else:   # "host": upload the reference's own vertices / triangles (validated against the sample grid)
                    if self.terrain is not None and getattr(self.terrain, "vertices", None) is not None:
                        verts, tris = self.terrain.vertices, self.terrain.triangles
                    else:
                        from .terrain import heightfield_to_trimesh
                        verts, tris = heightfield_to_trimesh(hs, tc.horizontal_scale, tc.vertical_scale, tc.slope_treshold)
                    verts = np.ascontiguousarray(verts, np.float32)
                    tris = np.ascontiguousarray(tris, np.uint32)
                    L.check(self.lib.grx_env_set_terrain_trimesh(self._h, verts.ctypes.data_as(L.PF), verts.shape[0],
                                                                 tris.ctypes.data_as(C.POINTER(C.c_uint32)), tris.shape[0],
                                                                 hs.ctypes.data_as(C.POINTER(C.c_int16)), hs.shape[0], hs.shape[1], *scal))
            else:
                L.check(self.lib.grx_env_set_terrain_heightfield(self._h, hs.ctypes.data_as(C.POINTER(C.c_int16)), hs.shape[0], hs.shape[1], *scal))
        else:
            L.check(self.lib.grx_env_set_terrain_plane(self._h, C.c_float(tc.static_friction), C.c_float(tc.restitution)))
        fr, p_fr = _f32p(params["friction"]); rs, p_rs = _f32p(params["restitution"])
        ms, p_ms = _f32p(params["motor_strength"]); bi, p_bi = _f32p(params["base_inertial"])
        eo, p_eo = _f32p(env_origins)
        if rough:
            lv, p_lv = _i32p(terrain_levels); ty, p_ty = _i32p(terrain_types)
            to, p_to = _f32p(t_org)
            L.check(self.lib.grx_env_set_params(self._h, p_fr, p_rs, p_ms, p_bi, p_eo, p_lv, p_ty, p_to, t_org.shape[0], t_org.shape[1]))
            self.terrain_origins = torch.from_numpy(to.copy()).to(self.device)
            self.max_terrain_level = t_org.shape[0]
        else:
            L.check(self.lib.grx_env_set_params(self._h, p_fr, p_rs, p_ms, p_bi, p_eo, None, None, None, 0, 0))

        # ---- zero-copy views (legged_robot.py:106-203)
        v = self._view
        self.root_states, self.dof_pos, self.dof_vel = v("root_states"), v("dof_pos"), v("dof_vel")
        self.last_dof_vel, self.last_actions, self.last_last_actions = v("last_dof_vel"), v("last_actions"), v("last_last_actions")
        self.commands, self.base_heights_offset = v("commands"), v("base_heights_offset")
        self.feet_air_time, self.feet_land_time, self.feet_contact_last = v("feet_air_time"), v("feet_land_time"), v("feet_contact_last")
        self._episode_length = v("episode_length").squeeze(1)
        self._episode_length_i64 = self._rigid_body_states = self._dof_state = None   # compat exports, created on first access
        self.terrain_levels, self.terrain_types = v("terrain_levels").squeeze(1), v("terrain_types").squeeze(1)
        self.env_origins, self.episode_sums_buf = v("env_origins"), v("episode_sums")
        self.obs_buf, self.pri_obs_buf, self.rew_buf = v("obs"), v("pri_obs"), v("rew")
        self.privileged_obs_buf = self.pri_obs_buf
        self._reset_u8, self._time_out_u8 = v("reset"), v("time_out")
        self.reset_buf, self.time_out_buf = self._reset_u8.view(torch.bool), self._time_out_u8.view(torch.bool)
        self.torques, self.contact_forces, self.foot_state = v("torques"), v("contact_forces"), v("foot_state")
        self.records, self._accum = v("records"), v("episode_accum")
        self.base_quat = self.root_states[:, 3:7]
        self.feet_indices = torch.tensor(self.tables["foot_links"], device=self.device)
        self.default_dof_pos = torch.tensor(self.tables["default_pos"], dtype=torch.float32, device=self.device).unsqueeze(0)
        self.commands_scale = torch.ones(N, 3, device=self.device)
        self.extras = {}
        self.common_step_counter = 0
        self._step_index = 0
        self._delay_rng = np.random.default_rng(seed + 15485863)
        self._zero_actions = torch.zeros(N, self.num_actions, device=self.device)
        self.init_done = True

    # ------------------------------------------------------------------ plumbing
    def _view(self, name):
        b = L.Buffer()
        L.check(self.lib.grx_env_get_buffer(self._h, name.encode(), C.byref(b)))
        return torch.as_tensor(_DevArray(b, self), device=self.device)

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    # derived base-frame quantities the reference keeps as buffers (legged_robot.py:308-311); play.py logs them (play.py:110-123).
    # The kernel computes them internally for obs / rewards; here they are evaluated on demand from the live root state.
    @property
    def base_lin_vel(self):
        return quat_rotate_inverse(self.root_states[:, 3:7], self.root_states[:, 7:10])

    @property
    def base_ang_vel(self):
        return quat_rotate_inverse(self.root_states[:, 3:7], self.root_states[:, 10:13])

    @property
    def projected_gravity(self):
        g = torch.zeros(self.num_envs, 3, device=self.device)
        g[:, 2] = -1.0
        return quat_rotate_inverse(self.root_states[:, 3:7], g)

    @property
    def episode_length_buf(self):
        """int64 [N] like the reference buffer (base_task.py:71-72): a device mirror the kernel refreshes every step from the live int32
        counter inside the state record.  Whole-tensor assignment (what OnPolicyRunner.learn does) goes through the setter."""
        if self._episode_length_i64 is None:
            self._episode_length_i64 = self._view("episode_length_i64")
            self._episode_length_i64.copy_(self._episode_length)
        return self._episode_length_i64

    @episode_length_buf.setter
    def episode_length_buf(self, value):   # OnPolicyRunner.learn assigns a new tensor (on_policy_runner.py:125-127)
        v = torch.as_tensor(value).to(self.device)
        self._episode_length.copy_(v.to(self._episode_length.dtype))
        if self._episode_length_i64 is not None:
            self._episode_length_i64.copy_(v.to(torch.int64))

    @property
    def rigid_body_states(self):
        """[N, num_bodies, 13] world state of every URDF link (legged_robot.py:135) — compat export, switched on by the first access and
        refreshed by every following step (like the reference's tensor it is stale between a reset and the next step)."""
        if self._rigid_body_states is None:
            self._rigid_body_states = self._view("rigid_body_states")
        return self._rigid_body_states

    @property
    def dof_state(self):
        """[N, num_dof, 2] interleaved (pos, vel) mirror (legged_robot.py:122-125) — compat export, read-only: the live state is dof_pos / dof_vel."""
        if self._dof_state is None:
            self._dof_state = self._view("dof_state")
        return self._dof_state

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self.lib.grx_env_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ VecEnv surface
    def get_observations(self):                                                        # base_task.py:107-108
        return self.obs_buf

    def get_privileged_observations(self):                                             # base_task.py:110-111
        return self.pri_obs_buf

    def reset_idx(self, env_ids, U=None):                                              # legged_robot.py:377-440
        ids = torch.as_tensor(env_ids, device=self.device).to(torch.int32).contiguous()
        if ids.numel() == 0:
            return
        pU = C.c_void_p(U.data_ptr()) if U is not None else None
        self._step_index += 1
        L.check(self.lib.grx_env_reset_idx(self._h, C.c_void_p(ids.data_ptr()), ids.numel(), pU, int(self.init_done),
                                           C.c_uint64(self._step_index), self._stream()))
        self._publish_extras(True)

    def reset(self):                                                                   # base_task.py:117-121
        self.reset_idx(torch.arange(self.num_envs, device=self.device))
        obs, pri, _, _, _ = self.step(self._zero_actions)
        return obs, pri

    def sample_delay(self):
        """legged_robot_fftai.py:53-54: one N(5, 2) draw per step shared by all envs, clamped at 0 (host RNG in the reference too)."""
        return max(0.0, float(self._delay_rng.normal(5.0, 2.0)))

    def step(self, actions, U=None, delay=None):                                       # legged_robot.py:222-246
        """actions [N, num_actions] fp32 on this env's device.  U / delay: parity-mode overrides (tests)."""
        a = actions if (actions.dtype == torch.float32 and actions.is_contiguous()) else actions.float().contiguous()
        if delay is None:
            delay = self.sample_delay()
        self.common_step_counter += 1                                                   # legged_robot.py:281
        push = int(bool(self.cfg.domain_rand.push_robots) and (self.common_step_counter % self.push_interval == 0))
        self._step_index += 1
        pU = C.c_void_p(U.data_ptr()) if U is not None else None
        L.check(self.lib.grx_env_step(self._h, C.c_void_p(a.data_ptr()), pU, C.c_float(delay), push,
                                      C.c_uint64(self._step_index), self._stream()))
        self._publish_extras(False)
        return self.obs_buf, self.pri_obs_buf, self.rew_buf, self.reset_buf, self.extras

    def _publish_extras(self, from_reset):
        slot = int(self.lib.grx_env_accum_slot(self._h))
        if self.sync_extras:   # reference-exact staleness: rebuild only when something reset; costs a host sync
            if not from_reset and not bool(self._reset_u8.any()):
                return
        self.extras["episode"] = _EpisodeInfo(self._accum, slot, self.reward_names, 1.0 / self.max_episode_length_s,
                                              float(self.num_envs), bool(self.cfg.terrain.curriculum))
        if self.cfg.env.send_timeouts:
            self.extras["time_outs"] = self.time_out_buf

    # ------------------------------------------------------------------ state in/out (tests, checkpointing of env state)
    CARRIED = ("root_states", "dof_pos", "dof_vel", "last_dof_vel", "last_actions", "last_last_actions", "commands",
               "base_heights_offset", "feet_air_time", "feet_land_time", "feet_contact_last")

    def load_state(self, d):
        for k in self.CARRIED:
            dst = getattr(self, k)
            dst.copy_(torch.as_tensor(np.asarray(d[k]).astype(np.float32)).reshape(dst.shape).to(self.device))
        self.episode_length_buf = torch.as_tensor(np.asarray(d["episode_length_buf"]).astype(np.int64))
        self.episode_sums_buf.copy_(torch.as_tensor(np.asarray(d["episode_sums"], np.float32)).to(self.device))
        self.common_step_counter = int(d["common_step_counter"])
        if self.custom_origins and "terrain_levels" in d:
            self.terrain_levels.copy_(torch.as_tensor(np.asarray(d["terrain_levels"]).astype(np.int32)).to(self.device))
            self.env_origins.copy_(torch.as_tensor(np.asarray(d["env_origins"], np.float32)).to(self.device))

    def post_physics_injected(self, actions, U, inj):
        """Test entry (grx_env_post_physics): run only the post-physics half on injected physics outputs."""
        ip = L.InjectedPhysics()
        keep = {}
        for k in ("torques", "foot_state", "torso_quat", "contact_forces", "avg_foot_force", "avg_foot_linvel"):
            keep[k] = inj[k].to(self.device, torch.float32).contiguous()
            setattr(ip, k, keep[k].data_ptr())
        self.common_step_counter += 1
        push = int(bool(self.cfg.domain_rand.push_robots) and (self.common_step_counter % self.push_interval == 0))
        self._step_index += 1
        a = actions.to(self.device, torch.float32).contiguous()
        L.check(self.lib.grx_env_post_physics(self._h, C.c_void_p(a.data_ptr()), C.c_void_p(U.data_ptr()) if U is not None else None,
                                              C.byref(ip), push, C.c_uint64(self._step_index), self._stream()))
        torch.cuda.synchronize(self.device)
        self._publish_extras(False)
        return self.obs_buf, self.pri_obs_buf, self.rew_buf, self.reset_buf, self.extras

    def debug_active_sig(self, enable=True):
        """Parity tests: export the per-substep active-set signature (include/grx_b200.h grx_env_debug_active_sig); returns the
        [N, 16] int64 view (bit patterns of the u64 hashes) or None when switched off."""
        L.check(self.lib.grx_env_debug_active_sig(self._h, int(bool(enable))))
        self.active_sig = self._view("active_sig") if enable else None
        return self.active_sig

    def debug_dynamics(self, index):
        nv = self.model["nd"] + 6
        M, h = np.zeros((nv, nv), np.float32), np.zeros(nv, np.float32)
        L.check(self.lib.grx_env_debug_dynamics(self._h, index, M.ctypes.data_as(L.PF), h.ctypes.data_as(L.PF)))
        return M, h
