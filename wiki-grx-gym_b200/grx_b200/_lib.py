"""ctypes binding of libgrx_b200.so (include/grx_b200.h).  The product path has NO CPU fallback: a missing
library or a missing CUDA device raises."""
from __future__ import annotations

import ctypes as C
import os

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(os.path.dirname(_PKG), "libgrx_b200.so")

F, I32, I64, U64, P = C.c_float, C.c_int32, C.c_int64, C.c_uint64, C.c_void_p
PF, PI = C.POINTER(C.c_float), C.POINTER(C.c_int32)


class Buffer(C.Structure):   # == grx_buffer: carb::gym::GymTensor's fields + element strides
    _fields_ = [("device", I32), ("dtype", I32), ("ndim", I32), ("own_data", I32), ("dims", I64 * 8), ("strides", I64 * 8), ("data", P)]


# dtype codes == GymTensorDataType (GymTensor.h:20-27) + Int32 / Int64
DT_NONE, DT_F32, DT_U32, DT_U64, DT_U8, DT_I16, DT_I32, DT_I64 = range(8)


class ModelDesc(C.Structure):
    _fields_ = [("nb", I32), ("nd", I32), ("nl", I32), ("ns", I32), ("nf", I32), ("nterm", I32), ("nankle", I32),
                ("parent", PI), ("jpos", PF), ("jrot", PF), ("axis", PF), ("mass", PF), ("com", PF), ("inertia", PF),
                ("dof_lower", PF), ("dof_upper", PF), ("dof_vel_limit", PF), ("dof_effort", PF),
                ("soft_lower", PF), ("soft_upper", PF), ("kp", PF), ("kd", PF), ("default_pos", PF),
                ("link_body", PI), ("link_pos", PF), ("link_rot", PF), ("sph_body", PI), ("sph_link", PI),
                ("sph_pos", PF), ("sph_rad", PF), ("foot_links", PI), ("term_links", PI), ("ankle_dofs", PI),
                ("torso_link", I32)]


_SIGMAS = ["action_diff", "action_diff_diff", "cmd_diff_ang_vel_yaw", "cmd_diff_base_height", "cmd_diff_base_orient",
           "cmd_diff_lin_vel_x", "cmd_diff_lin_vel_y", "cmd_diff_lin_vel_z", "cmd_diff_torso_orient", "dof_acc_new",
           "dof_tor_ankle_feet_lift_up", "dof_tor_new", "feet_air_force", "feet_air_height", "feet_air_time",
           "feet_land_time", "feet_speed_xy_close_to_ground", "feet_stumble", "limits_dof_pos", "limits_dof_tor",
           "limits_dof_vel", "pose_offset", "stand_still"]


class TaskCfg(C.Structure):
    _fields_ = ([("sim_dt", F), ("gravity", F), ("contact_offset", F), ("bounce_threshold", F), ("max_depen_vel", F),
                 ("erp", F), ("solver_iters", I32), ("decimation", I32), ("action_scale", F),
                 ("num_obs", I32), ("num_pri_obs", I32), ("num_actions", I32), ("num_height_points", I32),
                 ("clip_actions_min", F * 32), ("clip_actions_max", F * 32), ("clip_observations", F),
                 ("max_episode_length", F), ("max_episode_length_s", F), ("resample_interval", I32),
                 ("cmd_range", (F * 2) * 3), ("max_push_vel_xy", F),
                 ("add_noise", I32), ("randomize_init_dof_pos", I32), ("randomize_init_base_velocity", I32),
                 ("curriculum", I32), ("custom_origins", I32), ("measure_heights", I32),
                 ("noise_scale_vec", F * 128),
                 ("obs_scale_lin_vel", F), ("obs_scale_ang_vel", F), ("obs_scale_gravity", F), ("obs_scale_dof_pos", F),
                 ("obs_scale_dof_vel", F), ("obs_scale_action", F), ("obs_scale_height", F),
                 ("base_init_state", F * 13), ("measured_points_x", F * 16), ("measured_points_y", F * 16),
                 ("n_points_x", I32), ("n_points_y", I32), ("terrain_env_length", F),
                 ("reward_scale", F * 24),
                 ("base_height_target", F), ("swing_feet_height_target", F), ("feet_stumble_ratio", F),
                 ("feet_air_time_target", F), ("feet_land_time_max", F), ("soft_dof_vel_limit", F),
                 ("soft_torque_limit", F)]
                + [("sigma_" + n, F) for n in _SIGMAS]
                + [("seed", U64), ("env_id_offset", I32)])


class InjectedPhysics(C.Structure):
    _fields_ = [("torques", P), ("foot_state", P), ("torso_quat", P), ("contact_forces", P), ("avg_foot_force", P),
                ("avg_foot_linvel", P)]


class PPOCfg(C.Structure):
    _fields_ = [("num_envs", I32), ("num_steps", I32), ("num_obs", I32), ("num_pri_obs", I32), ("num_actions", I32),
                ("actor_hidden", I32 * 3), ("critic_hidden", I32 * 3), ("num_learning_epochs", I32),
                ("num_mini_batches", I32), ("clip_param", F), ("gamma", F), ("lam", F), ("value_loss_coef", F),
                ("entropy_coef", F), ("learning_rate", F), ("learning_rate_min", F), ("learning_rate_max", F),
                ("desired_kl", F), ("max_grad_norm", F), ("adaptive_schedule", I32), ("use_clipped_value_loss", I32),
                ("init_noise_std", F), ("use_tensor_cores", I32), ("world_size", I32),
                ("env_id_offset", I32), ("comm_timeout_ms", I32), ("seed", U64)]


class PhysGCfg(C.Structure):
    _fields_ = [("sim_dt", F), ("gravity", F), ("contact_offset", F), ("bounce_threshold", F), ("max_depen_vel", F), ("erp", F),
                ("solver_iters", I32), ("decimation", I32), ("action_scale", F), ("max_contacts", I32), ("max_self_contacts", I32)]


class TerrainTile(C.Structure):
    _fields_ = [("kind", I32), ("slope_peak", I32), ("plat_lo", I32), ("plat_hi", I32), ("coarse_index", I32), ("step_width", I32), ("step_height", I32),
                ("num_rings", I32), ("rect_index", I32), ("num_rects", I32)]


class TerrainGrid(C.Structure):
    _fields_ = [("num_rows", I32), ("num_cols", I32), ("tile_width", I32), ("tile_length", I32), ("border", I32), ("coarse_nx", I32), ("coarse_ny", I32),
                ("origin_x1", I32), ("origin_x2", I32), ("origin_y1", I32), ("origin_y2", I32)]


class GrxError(RuntimeError):
    pass


_lib = None


def lib():
    """Load libgrx_b200.so (built by __graft_entry__.build()).  Raises if it is missing — there is no fallback path."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise GrxError(f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'`")
        _lib = C.CDLL(LIB_PATH)
        _lib.grx_last_error.restype = C.c_char_p
        _lib.grx_env_accum_slot.restype = C.c_int64
        _lib.grx_env_accum_slot.argtypes = [C.c_void_p]
        _lib.grx_debug_launch_count.restype = C.c_uint64
        _lib.grx_env_info.restype = C.c_int64
        _lib.grx_env_info.argtypes = [C.c_void_p, C.c_int32]
        sizes = (I32 * 5)()
        _lib.grx_abi_sizes(sizes, 5)
        want = [C.sizeof(Buffer), C.sizeof(ModelDesc), C.sizeof(TaskCfg), C.sizeof(InjectedPhysics), C.sizeof(PPOCfg)]
        if list(sizes) != want:
            raise GrxError(f"ABI mismatch between grx_b200/_lib.py and libgrx_b200.so: {list(sizes)} vs {want}")
    return _lib


def check(rc):
    if rc != 0:
        raise GrxError(f"grx error {rc}: {lib().grx_last_error().decode()}")


EXPORTED = ["grx_last_error", "grx_version", "grx_abi_sizes", "grx_debug_launch_count", "grx_env_create", "grx_env_destroy",
            "grx_env_set_terrain_plane", "grx_env_set_terrain_heightfield", "grx_env_set_terrain_trimesh", "grx_env_set_terrain_trimesh_hf", "grx_env_set_params", "grx_env_get_buffer",
            "grx_env_step", "grx_env_reset_idx", "grx_env_accum_slot", "grx_env_post_physics", "grx_env_step_host", "grx_env_debug_dynamics", "grx_env_debug_active_sig", "grx_env_set_self_collision", "grx_env_info", "grx_terrain_generate", "grx_env_set_terrain_device",
            "grx_ppo_create", "grx_ppo_destroy", "grx_ppo_get_buffer", "grx_ppo_act", "grx_ppo_process_env_step",
            "grx_ppo_compute_returns", "grx_ppo_compute_returns_local", "grx_ppo_normalize_advantages",
            "grx_ppo_minibatch_grads", "grx_ppo_minibatch_apply", "grx_ppo_update", "grx_ppo_comm_handle", "grx_ppo_comm_open",
            "grx_ppo_minibatch_apply_comm", "grx_ppo_debug_timing", "grx_ppo_act_inference", "grx_gemm_debug", "grx_gemm_debug_stamps", "grx_gemm_debug_tile", "grx_ppo_debug_fused", "grx_ppo_debug_pipe", "grx_gemm_debug_dw_plan", "grx_debug_stamps32",
            "grx_physg_create", "grx_physg_destroy", "grx_physg_set_terrain_plane", "grx_physg_set_terrain_heightfield", "grx_physg_step"]
