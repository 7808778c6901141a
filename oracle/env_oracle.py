"""TEST INFRASTRUCTURE — CPU restatement (torch fp32 on CPU) of the reference env step for the
registered GR1T1 / GR1T2 tasks.  Checker for the CUDA env kernel, and the `port` CPU baseline.

Pinned against the reference itself: tests/test_env_oracle.py compares every buffer with golden
trajectories produced by the UNMODIFIED reference classes driven over FakeGym
(oracle/ref_harness/gen_golden.py -> tests/golden/env_*.npz).

Each method cites the reference lines it restates:
  LR = legged_gym/legged_gym/envs/base/legged_robot.py
  FF = legged_gym/legged_gym/envs/fftai/legged_robot_fftai.py
  G1 = legged_gym/legged_gym/envs/gr1t1/gr1t1.py
  TU = IsaacGym_Preview_4_Package/isaacgym/python/isaacgym/torch_utils.py
Physics (gym.simulate) is oracle/phys_impl.h (our spec; parity unpinned there).
"""
from __future__ import annotations

import math

import numpy as np
import torch

from grx_b200 import rng_layout as L

REWARD_NAMES = ["action_diff", "action_diff_diff", "cmd_diff_ang_vel_yaw", "cmd_diff_base_height",
                "cmd_diff_base_orient", "cmd_diff_lin_vel_x", "cmd_diff_lin_vel_y", "cmd_diff_lin_vel_z",
                "cmd_diff_torso_orient", "dof_acc_new", "dof_tor_ankle_feet_lift_up", "dof_tor_new",
                "feet_air_force", "feet_air_height", "feet_air_time", "feet_land_time",
                "feet_speed_xy_close_to_ground", "feet_stumble", "limits_dof_pos", "limits_dof_tor",
                "limits_dof_vel", "on_the_air", "pose_offset", "stand_still"]


def quat_rotate_inverse(q, v):  # TU:72-81
    q_w = q[:, -1]
    q_vec = q[:, :3]
    a = v * (2.0 * q_w ** 2 - 1.0).unsqueeze(-1)
    b = torch.cross(q_vec, v, dim=-1) * q_w.unsqueeze(-1) * 2.0
    c = q_vec * torch.bmm(q_vec.view(-1, 1, 3), v.view(-1, 3, 1)).squeeze(-1) * 2.0
    return a - b + c


def quat_apply(a, b):  # TU:49-56
    xyz = a[:, :3]
    t = xyz.cross(b, dim=-1) * 2
    return b + a[:, 3:] * t + xyz.cross(t, dim=-1)


def quat_from_euler_xyz(roll, pitch, yaw):  # TU:177-190
    cy, sy = torch.cos(yaw * 0.5), torch.sin(yaw * 0.5)
    cr, sr = torch.cos(roll * 0.5), torch.sin(roll * 0.5)
    cp, sp = torch.cos(pitch * 0.5), torch.sin(pitch * 0.5)
    qw = cy * cr * cp + sy * sr * sp
    qx = cy * sr * cp - sy * cr * sp
    qy = cy * cr * sp + sy * sr * cp
    qz = sy * cr * cp - cy * sr * sp
    return torch.stack([qx, qy, qz, qw], dim=-1)


class EnvOracle:
    def __init__(self, cfg, tables, consts, phys, terrain=None):
        """cfg: task config (grx_b200.config.make_cfg or the reference's cfg object);
        tables: grx_b200.robot.task_tables; consts: per-env parameters
        (friction, restitution, base_inertial, motor_strength, env_origins, [terrain_origins, terrain_levels, terrain_types]);
        phys: oracle.phys.PhysOracle (float32); terrain: None or dict(heights int16 [rows, cols], hscale, vscale, border)."""
        f = lambda a: torch.as_tensor(np.asarray(a), dtype=torch.float32)
        self.cfg, self.tb, self.phys = cfg, tables, phys
        self.N = N = len(consts["friction"])
        self.nd = nd = len(tables["kp"])
        self.L = L.layout(nd)   # slots of the uniform-draw table for this DOF count
        self.dt = cfg.control.decimation * cfg.sim.dt                                  # LR:92
        self.max_episode_length_s = cfg.env.episode_length_s
        self.max_episode_length = float(np.ceil(self.max_episode_length_s / self.dt))  # LR:101
        self.push_interval = float(np.ceil(cfg.domain_rand.push_interval_s / self.dt))  # LR:103
        self.resample_interval = int(cfg.commands.resampling_command_interval_s / self.dt)  # LR:104
        self.friction, self.restitution = np.asarray(consts["friction"], np.float32), np.asarray(consts["restitution"], np.float32)
        self.base_inertial = np.asarray(consts["base_inertial"], np.float32)
        self.motor_strength = np.asarray(consts["motor_strength"], np.float32)
        self.default_dof_pos = f(tables["default_pos"]).unsqueeze(0)
        self.torque_limits = f(tables["torque_limits"])
        self.dof_vel_limits = f(tables["dof_vel_limits"])
        self.dof_pos_limits = torch.stack([f(tables["soft_lower"]), f(tables["soft_upper"])], dim=1)
        self.clip_max = torch.tensor(np.asarray(cfg.normalization.clip_actions_max)).to(torch.float32)   # FF:172-173
        self.clip_min = torch.tensor(np.asarray(cfg.normalization.clip_actions_min)).to(torch.float32)
        self.feet, self.term, self.torso = list(tables["foot_links"]), list(tables["termination_links"]), list(tables["torso_links"])
        self.ankle = list(tables["ankle_dofs"])
        # reward scales x dt, zero scales dropped, alphabetical (dir()) order   LR:845-866, helpers.py:46
        sc = cfg.rewards.scales
        self.reward_names = [n for n in sorted(k for k in dir(sc) if not k.startswith("_") and k != "to_dict")
                             if getattr(sc, n) != 0 and n != "termination"]
        assert self.reward_names == REWARD_NAMES, self.reward_names
        self.reward_scales = {n: getattr(sc, n) * self.dt for n in self.reward_names}
        ns, os_, nl = cfg.noise.noise_scales, cfg.normalization.obs_scales, cfg.noise.noise_level
        nv = torch.zeros(9 + 3 * nd)                                                   # G1:315-336
        nv[3:6] = ns.ang_vel * nl * os_.ang_vel
        nv[6:9] = ns.gravity * nl * os_.gravity
        nv[9:9 + nd] = ns.dof_pos * nl * os_.dof_pos
        nv[9 + nd:9 + 2 * nd] = ns.dof_vel * nl * os_.dof_vel
        nv[9 + 2 * nd:9 + 3 * nd] = ns.action * nl * os_.action
        self.noise_scale_vec = nv
        self.obs_scales = os_
        self.mesh_type = cfg.terrain.mesh_type
        self.curriculum = bool(cfg.terrain.curriculum) and self.mesh_type in ("heightfield", "trimesh")   # LR:97-98
        self.custom_origins = self.mesh_type in ("heightfield", "trimesh")             # LR:1167-1168
        y = torch.tensor(cfg.terrain.measured_points_y)
        x = torch.tensor(cfg.terrain.measured_points_x)
        gx, gy = torch.meshgrid(x, y, indexing="ij")                                   # LR:1225-1233
        self.num_height_points = gx.numel()
        self.height_points = torch.zeros(N, self.num_height_points, 3)
        self.height_points[:, :, 0] = gx.flatten()
        self.height_points[:, :, 1] = gy.flatten()
        if terrain is not None and terrain.get("heights") is not None:
            self.height_samples = torch.tensor(np.asarray(terrain["heights"], np.int16))
            self.hscale, self.vscale, self.border = terrain["hscale"], terrain["vscale"], terrain["border"]
        else:
            self.height_samples = None
        self.env_origins = f(consts["env_origins"]).clone()
        if self.custom_origins:
            self.terrain_origins = f(consts["terrain_origins"])
            self.terrain_levels = torch.as_tensor(np.asarray(consts["terrain_levels"])).long().clone()
            self.terrain_types = torch.as_tensor(np.asarray(consts["terrain_types"])).long().clone()
            self.max_terrain_level = cfg.terrain.num_rows
            self.env_length = cfg.terrain.terrain_length
        self.base_init_state = torch.tensor(list(cfg.init_state.pos) + list(cfg.init_state.rot) + list(cfg.init_state.lin_vel)
                                            + list(cfg.init_state.ang_vel), dtype=torch.float32)   # LR:991-995
        self.gravity_vec = torch.tensor([0.0, 0.0, -1.0]).repeat(N, 1)                 # LR:141
        z = lambda *s: torch.zeros(*s)
        # ---- state carried between steps
        self.root_states = z(N, 13); self.root_states[:, 6] = 1
        self.dof_pos, self.dof_vel, self.last_dof_vel = z(N, nd), z(N, nd), z(N, nd)
        self.last_actions, self.last_last_actions = z(N, nd), z(N, nd)
        self.commands = z(N, 3)
        self.base_heights_offset = z(N)
        self.feet_air_time, self.feet_land_time = z(N, 2), z(N, 2)
        self.feet_contact_last = torch.zeros(N, 2, dtype=torch.bool)
        self.episode_length_buf = torch.zeros(N, dtype=torch.long)
        self.episode_sums = z(N, len(self.reward_names))
        self.common_step_counter = 0
        self.extras = {}

    # ------------------------------------------------------------------ state in/out
    CARRIED = ("root_states", "dof_pos", "dof_vel", "last_dof_vel", "last_actions", "last_last_actions", "commands",
               "base_heights_offset", "feet_air_time", "feet_land_time", "feet_contact_last", "episode_length_buf",
               "episode_sums")

    def load_state(self, d):
        for k in self.CARRIED:
            src = d["feet_contact_last"] if k == "feet_contact_last" else d[k]
            t = torch.as_tensor(np.asarray(src))
            getattr(self, k).copy_(t.to(getattr(self, k).dtype).reshape(getattr(self, k).shape))
        self.common_step_counter = int(d["common_step_counter"])
        if self.custom_origins and "terrain_levels" in d:
            self.terrain_levels.copy_(torch.as_tensor(d["terrain_levels"]).long())
            self.env_origins.copy_(torch.as_tensor(d["env_origins"]).float())

    # ------------------------------------------------------------------ step
    def step(self, actions, U, delay):
        """actions [N, nd] fp32; U [N, K] uniform draws (rng_layout); delay: the scalar of FF:53-54."""
        actions = torch.as_tensor(actions, dtype=torch.float32)
        U = torch.as_tensor(U, dtype=torch.float32)
        self.actions = torch.clip(actions, self.clip_min, self.clip_max)               # FF:171-177
        ph = self.physics(self.actions, delay)
        return self.post_physics(ph, U)

    def physics(self, actions, delay):
        """FF:46-88 with gym.simulate == physics oracle.  Returns the physics outputs the task code reads."""
        root, q, qd = self.root_states.numpy(), self.dof_pos.numpy(), self.dof_vel.numpy()
        out = self.phys.step(root, q, qd, actions.numpy(), self.last_actions.numpy(), float(delay), self.motor_strength,
                             self.base_inertial, self.friction, self.restitution)
        t = lambda a: torch.from_numpy(np.ascontiguousarray(a))
        self.last_active_sig = out["active_sig"]          # [N, decimation] uint64 (compared with the CUDA kernel's by the parity tests)
        return dict(torques=t(out["torques"]), foot_state=t(out["link_state"][:, self.feet]),
                    torso_quat=t(out["link_state"][:, self.torso[0], 3:7]),
                    contact_forces=t(out["contact_force"]), avg_feet_contact_force=t(out["avg_foot_force"]),
                    avg_feet_speed_xyz=t(out["avg_foot_linvel"]), avg_feet_speed_rpy=t(out["avg_foot_angvel"]))

    def _get_heights(self):                                                            # LR:1235-1274
        N = self.N
        if self.mesh_type == "plane" or self.height_samples is None:
            return torch.zeros(N, self.num_height_points)
        quat = self.root_states[:, 3:7].repeat(1, self.num_height_points)
        quat_yaw = quat.clone().view(-1, 4)                                            # utils/math.py:38-42
        quat_yaw[:, :2] = 0.0
        quat_yaw = quat_yaw / quat_yaw.norm(p=2, dim=-1).clamp(min=1e-9).unsqueeze(-1)
        pts = quat_apply(quat_yaw, self.height_points.view(-1, 3)).view(N, -1, 3) + self.root_states[:, :3].unsqueeze(1)
        pts = pts + self.border
        pts = (pts / self.hscale).long()
        px = torch.clip(pts[:, :, 0].reshape(-1), 0, self.height_samples.shape[0] - 2)
        py = torch.clip(pts[:, :, 1].reshape(-1), 0, self.height_samples.shape[1] - 2)
        h = torch.min(torch.min(self.height_samples[px, py], self.height_samples[px + 1, py]), self.height_samples[px, py + 1])
        return h.view(N, -1) * self.vscale

    def _resample_commands(self, ids, U, base):                                        # LR:650-677
        r = self.cfg.commands.ranges
        lo, hi = r.lin_vel_x
        self.commands[ids, 0] = ((hi - lo) * U[ids, base:base + 1] + lo).squeeze(1)
        lo, hi = r.lin_vel_y
        self.commands[ids, 1] = ((hi - lo) * U[ids, base + 1:base + 2] + lo).squeeze(1)
        self.commands[ids, :2] *= (torch.norm(self.commands[ids, :2], dim=1) > 0.1).unsqueeze(1)
        lo, hi = r.ang_vel_yaw
        self.commands[ids, 2] = ((hi - lo) * U[ids, base + 2:base + 3] + lo).squeeze(1)

    def post_physics(self, ph, U):
        """LR:269-305 (+ FF:90-133) on physics outputs ``ph``; returns the step() tuple of LR:240-246."""
        cfg, N, dt = self.cfg, self.N, self.dt
        rw = cfg.rewards
        contact_forces = ph["contact_forces"]
        self.contact_forces = contact_forces            # kept for the parity tests (net_contact_force tensor, LR:267)
        self.torques = ph["torques"]
        self.common_step_counter += 1                                                  # LR:281-282
        self.episode_length_buf += 1
        # ---- post_physics_step_update_state                                          LR:307-334
        base_quat = self.root_states[:, 3:7]
        base_lin_vel = quat_rotate_inverse(base_quat, self.root_states[:, 7:10])
        base_ang_vel = quat_rotate_inverse(base_quat, self.root_states[:, 10:13])
        g_proj = quat_rotate_inverse(base_quat, self.gravity_vec)
        ids = (self.episode_length_buf % self.resample_interval == 0).nonzero(as_tuple=False).flatten()
        self._resample_commands(ids, U, self.L.CMD_TIME)
        measured_heights = self._get_heights()
        if cfg.domain_rand.push_robots and (self.common_step_counter % self.push_interval == 0):   # LR:333-334, 786-797
            mv = cfg.domain_rand.max_push_vel_xy
            self.root_states[:, 7:9] = (mv - -mv) * U[:, self.L.PUSH:self.L.PUSH + 2] + -mv
        # FF:108-133
        feet_contact = contact_forces[:, self.feet, 2] > 1.0
        contact_filt = torch.logical_or(feet_contact, self.feet_contact_last)
        first_contact = (self.feet_air_time > 0) * contact_filt
        self.feet_air_time += dt
        foot_z = ph["foot_state"][:, :, 2]
        feet_height = torch.zeros(N, 2)
        for i in range(2):
            feet_height[:, i] = torch.mean(foot_z[:, i].unsqueeze(1) - measured_heights, dim=1)
        self.feet_land_time += dt
        self.feet_land_time = self.feet_land_time * feet_contact
        feet_contact_f = feet_contact.clone()   # becomes the (aliased) feet_contact / feet_contact_last tensor, FF:128-129
        # ---- check_termination                                                        LR:336-353
        reset_buf = torch.any(torch.norm(contact_forces[:, self.term, :], dim=-1) > 1.0, dim=1)
        reset_buf |= torch.abs(g_proj[:, 2]) < 0.33
        time_out = self.episode_length_buf > self.max_episode_length
        reset_buf |= time_out
        # ---- compute_reward                                                           LR:355-375
        terms = self._rewards(ph, base_lin_vel, base_ang_vel, g_proj, feet_height, feet_contact, first_contact, measured_heights)
        rew = torch.zeros(N)
        for k, name in enumerate(self.reward_names):
            r = terms[name] * self.reward_scales[name]
            rew += r
            self.episode_sums[:, k] += r
        # ---- reset_idx                                                                LR:377-440, FF:137-146
        env_ids = reset_buf.nonzero(as_tuple=False).flatten()
        if len(env_ids) > 0:
            if self.curriculum:
                self._update_terrain_curriculum(env_ids, U)
            if cfg.domain_rand.randomize_init_dof_pos:                                 # LR:725-734
                self.dof_pos[env_ids] = ((1.5 - 0.5) * U[env_ids, self.L.RESET_DOF:self.L.RESET_DOF + self.nd] + 0.5) * self.default_dof_pos
            else:
                self.dof_pos[env_ids] = self.default_dof_pos
            self.dof_vel[env_ids] = 0.0
            self.root_states[env_ids] = self.base_init_state                           # LR:750-779
            self.root_states[env_ids, :3] += self.env_origins[env_ids]
            if self.custom_origins:
                self.root_states[env_ids, :2] += (1.0 - -1.0) * U[env_ids, self.L.RESET_XY:self.L.RESET_XY + 2] + -1.0
            yaw = ((2 * np.pi - -2 * np.pi) * U[env_ids, self.L.RESET_YAW:self.L.RESET_YAW + 1] + -2 * np.pi).squeeze(1)
            zer = torch.zeros(len(env_ids))
            self.root_states[env_ids, 3:7] = quat_from_euler_xyz(zer, zer, yaw)
            if cfg.domain_rand.randomize_init_base_velocity:
                self.root_states[env_ids, 7:13] = (0.5 - -0.5) * U[env_ids, self.L.RESET_VEL:self.L.RESET_VEL + 6] + -0.5
            else:
                self.root_states[env_ids, 7:13] = 0.0
            self._resample_commands(env_ids, U, self.L.CMD_RESET)
            self.last_actions[env_ids] = 0.0
            self.last_dof_vel[env_ids] = 0.0
            self.feet_air_time[env_ids] = 0.0
            self.feet_land_time[env_ids] = 0.0
            self.episode_length_buf[env_ids] = 0
            self.extras["episode"] = {}
            for k, name in enumerate(self.reward_names):
                self.extras["episode"]["rew_" + name] = torch.mean(self.episode_sums[env_ids, k]) / self.max_episode_length_s
                self.episode_sums[env_ids, k] = 0.0
            if self.curriculum:
                self.extras["episode"]["terrain_level"] = torch.mean(self.terrain_levels.float())
            if cfg.env.send_timeouts:
                self.extras["time_outs"] = time_out
            feet_contact_f[env_ids] = False                                            # FF:141 (aliases feet_contact_last)
            self.last_last_actions[env_ids] = 0.0
        # ---- compute_observations                                                     LR:442-452, FF:148-167, G1:281-313
        hm = self.obs_scales.height_measurements
        off = torch.clip(self.root_states[:, 2].unsqueeze(1) - rw.base_height_target - measured_heights, min=-1.0, max=1.0) * hm
        self.base_heights_offset = torch.mean(off, dim=1)
        surround = off
        dof_pos_offset = self.dof_pos - self.default_dof_pos
        os_ = self.obs_scales
        obs = torch.cat((self.commands[:, :3] * 1.0, base_ang_vel * os_.ang_vel, g_proj * os_.gravity,
                         dof_pos_offset * os_.dof_pos, self.dof_vel * os_.dof_vel, self.actions * os_.action), dim=-1)
        pri = torch.cat((obs, base_lin_vel * os_.lin_vel, self.base_heights_offset.unsqueeze(1) * hm, feet_contact_f,
                         feet_height * hm, surround * hm), dim=-1)
        if cfg.noise.add_noise:                                                         # LR:478-481
            obs = obs + (2 * U[:, self.L.NOISE:self.L.NOISE + obs.shape[1]] - 1) * self.noise_scale_vec
        self.last_actions[:] = self.actions[:]                                         # LR:299-300
        self.last_dof_vel[:] = self.dof_vel[:]
        self.last_last_actions[:] = self.last_actions[:]                               # FF:94
        self.feet_air_time = self.feet_air_time * (~contact_filt)                      # FF:97
        self.feet_contact_last = feet_contact_f
        co = cfg.normalization.clip_observations                                       # LR:240-244
        self.obs_buf, self.pri_obs_buf = torch.clip(obs, -co, co), torch.clip(pri, -co, co)
        self.rew_buf, self.reset_buf, self.time_out_buf = rew, reset_buf, time_out
        self.dbg = dict(base_lin_vel=base_lin_vel, base_ang_vel=base_ang_vel, base_projected_gravity=g_proj,
                        feet_height=feet_height, measured_heights=measured_heights, reward_terms=terms)
        return self.obs_buf, self.pri_obs_buf, rew, reset_buf, self.extras

    def _update_terrain_curriculum(self, env_ids, U):                                  # LR:799-826
        distance = torch.norm(self.root_states[env_ids, :2] - self.env_origins[env_ids, :2], dim=1)
        move_up = distance > self.env_length / 2
        move_down = (distance < torch.norm(self.commands[env_ids, :2], dim=1) * self.max_episode_length_s * 0.5) * ~move_up
        self.terrain_levels[env_ids] += 1 * move_up - 1 * move_down
        rnd = torch.floor(U[env_ids, self.L.CURRICULUM] * self.max_terrain_level).long().clamp(max=self.max_terrain_level - 1)
        self.terrain_levels[env_ids] = torch.where(self.terrain_levels[env_ids] >= self.max_terrain_level, rnd,
                                                   torch.clip(self.terrain_levels[env_ids], 0))
        self.env_origins[env_ids] = self.terrain_origins[self.terrain_levels[env_ids], self.terrain_types[env_ids]]

    def _rewards(self, ph, v, w, g, feet_height, feet_contact, first_contact, measured_heights):
        """The 24 active terms, each from the cited reference method."""
        rw, cfg = self.cfg.rewards, self.cfg
        a, la, lla = self.actions, self.last_actions, self.last_last_actions
        cmd, tq, q, qd = self.commands, self.torques, self.dof_pos, self.dof_vel
        asc = cfg.control.action_scale
        nz = torch.norm(cmd[:, :2], dim=1) > 0.1
        cf = ph["contact_forces"]
        t = {}
        e = torch.sum(torch.abs((la - a) * asc), dim=1)                                 # FF:257-263
        t["action_diff"] = 1 - torch.exp(rw.sigma_action_diff * e)
        e = torch.sum(torch.abs((la - a) * asc - (lla - la) * asc), dim=1)              # FF:265-272
        t["action_diff_diff"] = 1 - torch.exp(rw.sigma_action_diff_diff * e)
        t["cmd_diff_ang_vel_yaw"] = torch.exp(rw.sigma_cmd_diff_ang_vel_yaw * torch.abs(cmd[:, 2] - w[:, 2]))   # FF:236-239
        bh = self.base_heights_offset                                                   # FF:241-245 (one step stale, App. B-21)
        t["cmd_diff_base_height"] = torch.exp(rw.sigma_cmd_diff_base_height * (torch.abs(bh) * (bh < 0)))
        t["cmd_diff_base_orient"] = torch.exp(rw.sigma_cmd_diff_base_orient * torch.sum(torch.abs(g[:, :2]), dim=1))  # FF:249-253
        t["cmd_diff_lin_vel_x"] = torch.exp(rw.sigma_cmd_diff_lin_vel_x * torch.abs(cmd[:, 0] - v[:, 0]))     # FF:211-214
        t["cmd_diff_lin_vel_y"] = torch.exp(rw.sigma_cmd_diff_lin_vel_y * torch.abs(cmd[:, 1] - v[:, 1]))     # FF:216-219
        t["cmd_diff_lin_vel_z"] = torch.exp(rw.sigma_cmd_diff_lin_vel_z * torch.abs(0 - v[:, 2]))             # FF:221-224
        tg = quat_rotate_inverse(ph["torso_quat"], self.gravity_vec)                    # G1:340-349
        t["cmd_diff_torso_orient"] = torch.exp(rw.sigma_cmd_diff_torso_orient * torch.sum(torch.abs(tg[:, :2]), dim=1))
        e = torch.sum(torch.abs((qd - self.last_dof_vel) / self.dt), dim=1)             # FF:284-288
        t["dof_acc_new"] = 1 - torch.exp(rw.sigma_dof_acc_new * e)
        lfh, rfh = feet_height[:, 0], feet_height[:, 1]                                 # G1:398-421 (same means as FF:118-124)
        tgt = rw.swing_feet_height_target
        half = len(self.ankle) // 2
        el = torch.sum(torch.abs(tq[:, self.ankle[:half]]), dim=1) * torch.abs(lfh) * (lfh > (tgt / 2))
        er = torch.sum(torch.abs(tq[:, self.ankle[half:]]), dim=1) * torch.abs(rfh) * (rfh > (tgt / 2))
        t["dof_tor_ankle_feet_lift_up"] = 1 - torch.exp(rw.sigma_dof_tor_ankle_feet_lift_up * (el + er))
        t["dof_tor_new"] = 1 - torch.exp(rw.sigma_dof_tor_new * torch.sum(torch.abs(tq), dim=1))               # FF:292-296
        mid = torch.abs(self.feet_air_time - rw.feet_air_time_target / 2)               # G1:534-549
        e = torch.sum(mid * ph["avg_feet_contact_force"], dim=1)
        t["feet_air_force"] = torch.exp(rw.sigma_feet_air_force * e) * nz
        mn = torch.min(torch.stack((lfh, rfh)), dim=0)[0]                               # G1:502-532
        eh = torch.stack([torch.abs(lfh - mn - tgt), torch.abs(rfh - mn - tgt)], dim=1)
        t["feet_air_height"] = torch.exp(rw.sigma_feet_air_height * torch.sum(mid * eh, dim=1)) * nz
        e = torch.exp(rw.sigma_feet_air_time * torch.abs(self.feet_air_time - rw.feet_air_time_target))        # G1:490-500
        t["feet_air_time"] = torch.sum(e * first_contact, dim=1) * nz
        e = (self.feet_land_time - rw.feet_land_time_max) * (self.feet_land_time > rw.feet_land_time_max)      # G1:551-560
        t["feet_land_time"] = torch.sum(1 - torch.exp(rw.sigma_feet_land_time * e), dim=1) * nz
        sxyz = ph["avg_feet_speed_xyz"]                                                 # G1:425-454
        q4 = tgt / 4
        cl = torch.abs(lfh - q4) * (lfh < q4) / q4
        cr = torch.abs(rfh - q4) * (rfh < q4) / q4
        e = torch.norm(sxyz[:, 0, :2], dim=1) * cl + torch.norm(sxyz[:, 1, :2], dim=1) * cr
        t["feet_speed_xy_close_to_ground"] = torch.exp(rw.sigma_feet_speed_xy_close_to_ground * e)
        ff = cf[:, self.feet]                                                           # G1:571-589
        el = torch.norm(ff[:, 0, :2], dim=1) - rw.feet_stumble_ratio * torch.abs(ff[:, 0, 2])
        er = torch.norm(ff[:, 1, :2], dim=1) - rw.feet_stumble_ratio * torch.abs(ff[:, 1, 2])
        el, er = el * (el > 0), er * (er > 0)
        t["feet_stumble"] = (1 - torch.exp(rw.sigma_feet_stumble * el)) + (1 - torch.exp(rw.sigma_feet_stumble * er))
        ool = -(q - self.dof_pos_limits[:, 0]).clip(max=0.0)                            # FF:322-334
        ool = ool + (q - self.dof_pos_limits[:, 1]).clip(min=0.0)
        t["limits_dof_pos"] = 1 - torch.exp(rw.sigma_limits_dof_pos * torch.sum(torch.abs(ool), dim=1))
        e = torch.sum((torch.abs(tq) - self.torque_limits * rw.soft_torque_limit).clip(min=0.0), dim=1)        # FF:345-352
        t["limits_dof_tor"] = 1 - torch.exp(rw.sigma_limits_dof_tor * e)
        e = torch.sum((torch.abs(qd) - self.dof_vel_limits * rw.soft_dof_vel_limit).clip(min=0.0, max=1.0), dim=1)  # FF:336-343
        t["limits_dof_vel"] = 1 - torch.exp(rw.sigma_limits_dof_vel * e)
        t["on_the_air"] = (torch.sum(feet_contact, dim=1) == 0)                         # G1:562-567
        e = torch.sum(torch.abs(q - self.default_dof_pos), dim=1)
        t["pose_offset"] = torch.exp(rw.sigma_pose_offset * e)                          # FF:300-304
        t["stand_still"] = torch.exp(rw.sigma_stand_still * e) * (torch.norm(cmd[:, :2], dim=1) < 0.1)         # FF:196-207
        return t
