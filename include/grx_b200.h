/*
 * grx_b200 — C ABI of the B200-native GRx locomotion-RL hot path (libgrx_b200.so).
 *
 * Plain pointers and sizes only; no torch types.  Every entry returns 0 on success or a
 * negative GRX_E_* code; the message is available from grx_last_error().  Device work is
 * enqueued on the caller's stream (cudaStream_t passed as void*, NULL = legacy default
 * stream); nothing in the step / update entries synchronises with the host.
 *
 * Reference interfaces replaced (paths under the reference repo, FFTAI/Wiki-GRx-Gym):
 *   env side  — the `gym` object + gymtorch tensor interop used by
 *               legged_gym/legged_gym/envs/base/legged_robot.py (create_sim/load_asset/create_actor:
 *               507-528, 926-1090; acquire_*_tensor + wrap_tensor: 110-135; set_dof_actuation_force_tensor /
 *               simulate / refresh_*: legged_robot_fftai.py:67-76; set_*_tensor_indexed: 737-740, 782-784)
 *               and the task arithmetic on top of it (step: legged_robot.py:222-246).
 *               GymTensor descriptor: IsaacGym_Preview_4_Package/isaacgym/python/isaacgym/_bindings/src/gymtorch/GymTensor.h:20-41.
 *   PPO side  — rsl_rl/rsl_rl/algorithms/ppo.py (act 144-171, process_env_step 177-196, compute_returns 198-205,
 *               update 215-321), storage/base_storage.py:80-141, storage/rollout_storage.py:56-112,
 *               modules/actor_critic_mlp.py:165-231.
 */
#ifndef GRX_B200_H
#define GRX_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GRX_OK 0
#define GRX_E_INVALID (-1)  /* bad argument / unsupported model topology */
#define GRX_E_CUDA (-2)     /* CUDA runtime error (see grx_last_error) */
#define GRX_E_NOTFOUND (-3) /* unknown buffer name */
#define GRX_E_STATE (-4)    /* call order violated (e.g. step before set_params) */

/* dtype codes == carb::gym::GymTensorDataType (GymTensor.h:20-27: None 0, Fp32 1, Uint32 2, Uint64 3, Uint8 4, Int16 5), extended with
 * Int32 6 and Int64 7 */
enum { GRX_NONE = 0, GRX_F32 = 1, GRX_U32 = 2, GRX_U64 = 3, GRX_U8 = 4, GRX_I16 = 5, GRX_I32 = 6, GRX_I64 = 7 };
#define GRX_MAX_DIMS 8 /* == GYM_TENSOR_MAX_DIMENSIONS (GymTensor.h:11) */

/* Non-owning description of a device buffer: the fields of carb::gym::GymTensor (GymTensor.h:32-41: device, dtype, numDims, dims[8], data,
 * ownData) plus element strides, because per-env state lives in one array-of-records and is exported as strided views. */
typedef struct {
    int32_t device;     /* CUDA device ordinal (GymTensor::device; never -1: there is no CPU pipeline) */
    int32_t dtype;      /* GRX_* */
    int32_t ndim;
    int32_t own_data;   /* always 0: views stay valid for the lifetime of the env / ppo object (GymTensor::ownData == false, gymtorch.cpp:163-166) */
    int64_t dims[GRX_MAX_DIMS];
    int64_t strides[GRX_MAX_DIMS]; /* in elements */
    void *data;
} grx_buffer;

/* Flat dynamic model (host pointers, copied at create).  Produced by grx_b200/urdf.py + robot.py from the URDF
 * (what gym.load_asset + the name-matching loops of legged_robot.py:176-192,594-616,1092-1161 produce). */
typedef struct {
    int32_t nb, nd, nl, ns, nf, nterm, nankle;
    const int32_t *parent;                                       /* [nb] */
    const float *jpos, *jrot, *axis;                             /* [nb*3] [nb*9] [nb*3] */
    const float *mass, *com, *inertia;                           /* [nb] [nb*3] [nb*6] */
    const float *dof_lower, *dof_upper, *dof_vel_limit, *dof_effort; /* [nd] hard limits (URDF) */
    const float *soft_lower, *soft_upper;                        /* [nd] reward soft limits */
    const float *kp, *kd, *default_pos;                          /* [nd] */
    const int32_t *link_body;                                    /* [nl] */
    const float *link_pos, *link_rot;                            /* [nl*3] [nl*9] */
    const int32_t *sph_body, *sph_link;                          /* [ns] already in contact-priority order */
    const float *sph_pos, *sph_rad;                              /* [ns*3] [ns] */
    const int32_t *foot_links, *term_links, *ankle_dofs;         /* [nf] [nterm] [nankle] */
    int32_t torso_link;
} grx_model_desc;

/* Simulation + task parameters (legged_robot_config.py / gr1t1_config.py values after _parse_cfg, legged_robot.py:91-104). */
typedef struct {
    /* sim */
    float sim_dt, gravity, contact_offset, bounce_threshold, max_depen_vel, erp;
    int32_t solver_iters, decimation;
    float action_scale;
    /* task */
    int32_t num_obs, num_pri_obs, num_actions, num_height_points; /* 39, 168, 10, 121 */
    float clip_actions_min[32], clip_actions_max[32], clip_observations;
    float max_episode_length;        /* ceil(episode_length_s / dt) */
    float max_episode_length_s;
    int32_t resample_interval;       /* steps */
    float cmd_range[3][2];           /* lin_vel_x, lin_vel_y, ang_vel_yaw */
    float max_push_vel_xy;
    int32_t add_noise, randomize_init_dof_pos, randomize_init_base_velocity, curriculum, custom_origins, measure_heights;
    float noise_scale_vec[128];      /* [num_obs] (39 lower-limb, 105 full body) */
    float obs_scale_lin_vel, obs_scale_ang_vel, obs_scale_gravity, obs_scale_dof_pos, obs_scale_dof_vel,
          obs_scale_action, obs_scale_height;
    float base_init_state[13];
    float measured_points_x[16], measured_points_y[16];
    int32_t n_points_x, n_points_y;
    float terrain_env_length;        /* terrain.env_length (curriculum move_up threshold = /2) */
    /* rewards: scale*dt per term in alphabetical order of the 24 active terms, then parameters */
    float reward_scale[24];
    float base_height_target, swing_feet_height_target, feet_stumble_ratio, feet_air_time_target, feet_land_time_max;
    float soft_dof_vel_limit, soft_torque_limit;
    float sigma_action_diff, sigma_action_diff_diff, sigma_cmd_diff_ang_vel_yaw, sigma_cmd_diff_base_height,
          sigma_cmd_diff_base_orient, sigma_cmd_diff_lin_vel_x, sigma_cmd_diff_lin_vel_y, sigma_cmd_diff_lin_vel_z,
          sigma_cmd_diff_torso_orient, sigma_dof_acc_new, sigma_dof_tor_ankle_feet_lift_up, sigma_dof_tor_new,
          sigma_feet_air_force, sigma_feet_air_height, sigma_feet_air_time, sigma_feet_land_time,
          sigma_feet_speed_xy_close_to_ground, sigma_feet_stumble, sigma_limits_dof_pos, sigma_limits_dof_tor,
          sigma_limits_dof_vel, sigma_pose_offset, sigma_stand_still;
    uint64_t seed;                   /* Philox key for fast-mode draws */
    int32_t env_id_offset;           /* global index of local env 0 (multi-GPU sharding; RNG streams keyed by global id) */
} grx_task_cfg;

typedef struct grx_env grx_env;

const char *grx_last_error(void);
int grx_version(void);
/* sizeof() of {grx_buffer, grx_model_desc, grx_task_cfg, grx_injected_physics, grx_ppo_cfg}, for FFI bindings to self-check their layouts */
int grx_abi_sizes(int32_t *out, int32_t n);

/* Launch accounting (bench.py's `gpu_launches`): kernels this library has launched in the calling process so far; a CUDA-graph
 * replay counts its kernel nodes. */
uint64_t grx_debug_launch_count(void);

/* ---- environment -------------------------------------------------------------------------------------------- */
int grx_env_create(const grx_model_desc *model, const grx_task_cfg *cfg, int32_t num_envs, int32_t device, grx_env **out);
int grx_env_destroy(grx_env *env);

/* terrain: replaces gym.add_ground (legged_robot.py:868-876) / gym.add_heightfield (878-901); samples = host int16
 * [rows, cols], x = row axis, world x = row*hscale - border (same convention as height_samples at legged_robot.py:899-901) */
int grx_env_set_terrain_plane(grx_env *env, float friction, float restitution);
int grx_env_set_terrain_heightfield(grx_env *env, const int16_t *samples, int32_t rows, int32_t cols, float hscale,
                                    float vscale, float border, float friction, float restitution);

/* replaces gym.add_triangle_mesh (legged_robot.py:903-924): vertices [nv, 3] fp32 / triangles [nt, 3] u32 exactly as
 * terrain_utils.convert_heightfield_to_trimesh (terrain_utils.py:286-350) produces them from `samples`.  The mesh is checked to be this structured
 * conversion — every vertex (height == sample * vscale, x / y an integral shift of at most one cell from its grid position) and every triangle
 * (the two index triples of its cell) — anything else is rejected with GRX_E_INVALID.  Contacts are resolved on the mesh's TOP SURFACE: the
 * sample grid with the shifted vertices of the steep-edge snapping (flat treads + vertical walls instead of the heightfield's ramps); a
 * vertical wall carries no lateral contact. */
int grx_env_set_terrain_trimesh(grx_env *env, const float *vertices, int32_t nv, const uint32_t *triangles, int32_t nt,
                                const int16_t *samples, int32_t rows, int32_t cols, float hscale, float vscale, float border,
                                float friction, float restitution);
/* The same terrain built on the device from the sample grid alone (the vertex snapping of terrain_utils.py:315-328 as a kernel; bit-identical
 * shifts, buffer "terrain_moves"): what GRXVecEnv uses for mesh_type = 'trimesh' instead of materialising the 2.7 M-vertex host mesh. */
int grx_env_set_terrain_trimesh_hf(grx_env *env, const int16_t *samples, int32_t rows, int32_t cols, float hscale, float vscale,
                                   float border, float slope_threshold, float friction, float restitution);

/* per-env parameters (host pointers): replaces the O(num_envs) create_actor loop, legged_robot.py:1008-1082.
 * terrain_* may be NULL when custom_origins == 0.  terrain_origins = [t_rows, t_cols, 3]. */
int grx_env_set_params(grx_env *env, const float *friction, const float *restitution, const float *motor_strength,
                       const float *base_inertial, const float *env_origins, const int32_t *terrain_levels,
                       const int32_t *terrain_types, const float *terrain_origins, int32_t t_rows, int32_t t_cols);

/* zero-copy views of device state, by name (what acquire_*_tensor + gymtorch.wrap_tensor give, legged_robot.py:110-135):
 * root_states dof_pos dof_vel last_dof_vel last_actions last_last_actions commands base_heights_offset feet_air_time
 * feet_land_time feet_contact_last episode_length (i32, the live counter) terrain_levels terrain_types env_origins episode_sums
 * obs pri_obs rew reset time_out torques contact_forces foot_state episode_accum params records active_sig
 * Compat exports, switched on by the FIRST request and refreshed by every following step (they cost extra stores per step):
 *   rigid_body_states [N, nl, 13]  pos3 quat4(xyzw) linvel3 angvel3 of every URDF link (acquire_rigid_body_state_tensor, legged_robot.py:135)
 *   dof_state         [N, nd, 2]   interleaved (pos, vel) mirror of the DOF state (acquire_dof_state_tensor, legged_robot.py:122-125); read-only
 *   episode_length_i64 [N]         int64 mirror of the episode counter (the reference buffer's dtype, base_task.py:71-72) */
int grx_env_get_buffer(grx_env *env, const char *name, grx_buffer *out);

/* One policy step = legged_robot.py:222-246 with the GR1T1 MRO: action clip, `decimation` substeps of
 * PD torque + articulated dynamics + contact + integration, then the whole post-physics path, in ONE kernel launch.
 *   d_actions [N, num_actions] device fp32;  d_uniform: [N, GRX_RNG_K] device fp32 draws (parity mode) or NULL (fast
 *   mode: counter-based Philox in-kernel);  delay: the scalar of legged_robot_fftai.py:53-54;  push != 0 on steps where
 *   common_step_counter % push_interval == 0 (legged_robot.py:333-334);  step_index feeds the Philox counter. */
int grx_env_step(grx_env *env, const float *d_actions, const float *d_uniform, float delay, int32_t push,
                 uint64_t step_index, void *stream);

/* Host-invoked reset_idx (legged_robot.py:377-440 as reached from BaseTask.reset(), base_task.py:117-121): d_ids = device int32
 * env indices (NULL = all, then n must be num_envs); d_uniform as in grx_env_step; curriculum_active = the reference's
 * `init_done` guard (legged_robot.py:806-808). */
int grx_env_reset_idx(grx_env *env, const int32_t *d_ids, int32_t n, const float *d_uniform, int32_t curriculum_active,
                      uint64_t step_index, void *stream);

/* extras["episode"] (legged_robot.py:420-427): every step / reset launch accumulates, over the envs it resets, the 24
 * episode sums [0..23], the reset count [24] and the sum of terrain levels over ALL envs [25] into one 32-float slot of the
 * ring buffer `episode_accum` [256, 32].  Returns the slot the most recent launch used (no host sync needed to read it
 * later on the stream). */
int64_t grx_env_accum_slot(grx_env *env);

/* Test entry: the post-physics half only, on injected physics outputs (device pointers); the env records must already
 * hold the post-physics root / dof state.  Isolates the reference's own arithmetic (LR/FF/G1) from our dynamics spec. */
typedef struct {
    const float *torques;         /* [N, nd] */
    const float *foot_state;      /* [N, nf, 13] */
    const float *torso_quat;      /* [N, 4] */
    const float *contact_forces;  /* [N, nl, 3] */
    const float *avg_foot_force;  /* [N, nf] */
    const float *avg_foot_linvel; /* [N, nf, 3] */
} grx_injected_physics;
int grx_env_post_physics(grx_env *env, const float *d_actions, const float *d_uniform, const grx_injected_physics *inj,
                         int32_t push, uint64_t step_index, void *stream);

/* Host-buffer convenience used for end-to-end timing: H2D of actions (pinned or pageable host memory), step,
 * D2H of obs / pri_obs / rew / reset, then a stream synchronise.  Any output pointer may be NULL. */
int grx_env_step_host(grx_env *env, const float *h_actions, float delay, int32_t push, uint64_t step_index,
                      float *h_obs, float *h_pri_obs, float *h_rew, uint8_t *h_reset, void *stream);

/* Debug / parity: mass matrix [nv*nv] and bias vector [nv] (internal velocity order: joints, base linear, base angular)
 * of env `index` for its current state, computed by the same device code the step uses.  Host output pointers. */
int grx_env_debug_dynamics(grx_env *env, int32_t index, float *h_M, float *h_h);

/* Debug / parity: export, for every env and substep of the following steps, the ACTIVE-SET SIGNATURE of the dynamics (a 64-bit hash of
 * the discrete decisions: which contact spheres were accepted, the terrain triangle under each, the restitution branch, which joint
 * limits were active and on which side) into the device buffer "active_sig" [N, 16] u64.  oracle/phys_impl.h computes the same hash, so a
 * row that differs from the oracle beyond rounding can be attributed to a differing decision (contact-threshold flip) or flagged. */
int grx_env_debug_active_sig(grx_env *env, int32_t enable);

#define GRX_RNG_K 68 /* == grx_b200/rng_layout.py K of the registered 10-DOF tasks; in general 28 + 4 * num_dof (grx_env_info(env, 0)) */

/* Models other than the registered lower-limb tree — any floating-base revolute tree with <= 36 bodies in depth-first order, <= 32 DOF, <= 48
 * links, <= 32 contact spheres, 2 feet: the full-body 32-DOF GR1T1 / GR1T2 of gr1t1_config.py:10-307 (num_obs = 9 + 3 * 32 = 105, num_pri_obs =
 * 105 + 8 + H) — run behind the SAME grx_env_* entries on the generic-topology kernels (csrc/grx_phys_generic.cu); GRX_ENV_GENERIC=1 in the
 * environment routes the lower-limb model there as well.
 * Robot self-collision (legged_robot_config.py:121 self_collisions = 0 = enabled; create_actor(..., collision_filter = 0),
 * legged_robot.py:1022-1028): pairs = [npairs, 2] candidate sphere pairs (indices into the model's sphere arrays, priority order;
 * grx_b200/robot.py:self_collision_pairs), at most max_self_contacts (<= 4) sphere-sphere contacts per robot and substep.  Generic-topology
 * envs only (GRX_E_INVALID on the specialised lower-limb kernel, which carries no self-contact rows). */
int grx_env_set_self_collision(grx_env *env, const int32_t *pairs, int32_t npairs, int32_t max_self_contacts);
/* Shape facts a binding needs before it allocates: what = 0 uniform draws per env and step (d_uniform row width), 1 floats per state record,
 * 2 floats per parameter record, 3 actuated DOF, 4 = 1 when the generic-topology kernels run this env; -1 on a bad argument. */
int64_t grx_env_info(grx_env *env, int32_t what);

/* ---- generic-topology dynamics (full-body 32-DOF GR1T1 / GR1T2, robot self-collision) --------------------------------------------------
 * The fused env kernel above is specialised to the registered lower-limb tree.  grx_physg runs the same dynamics spec for ANY revolute tree
 * with a floating base (<= 36 bodies in depth-first order, <= 32 DOF, <= 48 links, <= 32 contact spheres) incl. robot self-collision
 * (legged_robot_config.py:121 self_collisions = 0 = enabled; create_actor(..., collision_filter = 0), legged_robot.py:1022-1028): one policy
 * step of physics = the body of during_physics_step (legged_robot_fftai.py:51-88) — decimation x [PD torque (legged_robot.py:679-715) ->
 * articulated dynamics -> ground / self / joint-limit constraints -> integrate] + the foot averages.  grx_physg_step has the signature of the
 * CPU oracle's grx_oracle_physics_step on DEVICE pointers (state in place), so the two are compared call for call. */
typedef struct {
    float sim_dt, gravity, contact_offset, bounce_threshold, max_depen_vel, erp;
    int32_t solver_iters, decimation;
    float action_scale;
    int32_t max_contacts;        /* ground contacts per robot and substep (<= 8) */
    int32_t max_self_contacts;   /* sphere-sphere self-contacts per robot and substep (<= 4), 0 = self-collision off */
} grx_physg_cfg;
typedef struct grx_physg grx_physg;
/* self_pairs: [npairs, 2] candidate sphere pairs (indices into the model's sphere arrays, priority order; grx_b200/robot.py:self_collision_pairs) */
int grx_physg_create(const grx_model_desc *model, const int32_t *self_pairs, int32_t npairs, const grx_physg_cfg *cfg, int32_t num_envs,
                     int32_t device, grx_physg **out);
int grx_physg_destroy(grx_physg *p);
int grx_physg_set_terrain_plane(grx_physg *p, float friction, float restitution);
int grx_physg_set_terrain_heightfield(grx_physg *p, const int16_t *samples, int32_t rows, int32_t cols, float hscale, float vscale,
                                      float border, float friction, float restitution);
/* d_root [N,13], d_dof_pos / d_dof_vel [N,nd] advanced in place; d_actions / d_last_actions [N,nd] (already clipped), delay as in
 * legged_robot_fftai.py:53-61; outputs: d_torques [N,nd] (last substep), d_link_state [N,nl,13], d_contact_force [N,nl,3],
 * d_avg_foot_force [N,nf], d_avg_foot_linvel / d_avg_foot_angvel [N,nf,3]; d_active_sig [N, decimation] u64 or NULL. */
int grx_physg_step(grx_physg *p, float *d_root, float *d_dof_pos, float *d_dof_vel, const float *d_actions, const float *d_last_actions,
                   float delay, const float *d_motor_strength, const float *d_base_inertial, const float *d_friction,
                   const float *d_restitution, float *d_torques, float *d_link_state, float *d_contact_force, float *d_avg_foot_force,
                   float *d_avg_foot_linvel, float *d_avg_foot_angvel, uint64_t *d_active_sig, void *stream);

/* ---- terrain generation on the device ------------------------------------------------------------------------------------------
 * Replaces the array work of Terrain.__init__ (legged_gym/utils/terrain.py:38-164) and the isaacgym/terrain_utils.py generators it calls
 * (pyramid_sloped_terrain :74-106, random_uniform_terrain :17-51, pyramid_stairs_terrain :195-227, discrete_obstacles_terrain :109-149): ONE
 * kernel writes the whole int16 sample grid (border included) into d_samples [tot_rows, tot_cols], a second one reduces the per-tile origin
 * heights (terrain.py:159-163).  The numpy random stream stays on the host — grx_b200/terrain.py draws it in the reference's call order and
 * fills the descriptors — so the grid is bit-identical to the reference's for the same np.random.seed (tests/test_terrain_gpu.py). */
typedef struct {
    int32_t kind;         /* 0 smooth pyramid slope, 1 rough slope (slope + random-uniform roughness), 2 pyramid stairs, 3 discrete obstacles */
    int32_t slope_peak;   /* kinds 0/1: int(slope * (horizontal_scale / vertical_scale) * (width / 2)), terrain_utils.py:93 */
    int32_t plat_lo, plat_hi; /* kinds 0/1: the platform corner sample x1 == y1 (:100-101); kind 3: the cleared centre platform [lo, hi) (:145-148) */
    int32_t coarse_index; /* kind 1: which [coarse_nx, coarse_ny] block of h_coarse holds this tile's np.random.choice levels (:37) */
    int32_t step_width, step_height, num_rings; /* kind 2: samples per step, height units per step (signed), number of rings painted (:213-226) */
    int32_t rect_index, num_rects; /* kind 3: range of h_rects rows (start_i, start_j, width, length, height) in draw order (:135-143) */
} grx_terrain_tile;
typedef struct {
    int32_t num_rows, num_cols;       /* tiles: difficulty rows x type columns (terrain.py:51) */
    int32_t tile_width, tile_length;  /* samples per tile (width_per_env_pixels, length_per_env_pixels; must be equal, as upstream) */
    int32_t border;                   /* samples */
    int32_t coarse_nx, coarse_ny;     /* size of one coarse level block (random_uniform_terrain's down-sampled grid) */
    int32_t origin_x1, origin_x2, origin_y1, origin_y2; /* centre window of the origin height (terrain.py:159-162) */
} grx_terrain_grid;
/* h_tiles [num_rows * num_cols] (row-major: tile (i, j) at i * num_cols + j); h_rx [tile_length] / h_ry [tile_width] = the float64 ramps of
 * pyramid_sloped_terrain; h_up_i0 / h_up_fx [tile_length], h_up_j0 / h_up_fy [tile_width] = cell index and fraction of every fine sample in the
 * coarse grid (interp2d, :41-48); h_coarse [num_coarse, coarse_nx, coarse_ny] int16; h_rects [num_rects_total, 5] int32.
 * d_samples: DEVICE int16 [num_rows * tile_length + 2 border, num_cols * tile_width + 2 border]; h_origin_zmax: host int32 [num_rows * num_cols]
 * (max sample of each tile's centre window; origin z = that * vertical_scale) or NULL.  Synchronises the stream before returning. */
int grx_terrain_generate(const grx_terrain_grid *grid, const grx_terrain_tile *h_tiles, const double *h_rx, const double *h_ry,
                         const int32_t *h_up_i0, const double *h_up_fx, const int32_t *h_up_j0, const double *h_up_fy,
                         const int16_t *h_coarse, int32_t num_coarse, const int32_t *h_rects, int32_t num_rects_total,
                         int16_t *d_samples, int32_t *h_origin_zmax, int32_t device, void *stream);
/* grx_env_set_terrain_heightfield / _trimesh_hf with the samples already ON THE DEVICE (e.g. written by grx_terrain_generate): device-to-device
 * copy, no host round trip.  slope_threshold < 0: heightfield; >= 0: structured trimesh built on the device (as grx_env_set_terrain_trimesh_hf). */
int grx_env_set_terrain_device(grx_env *env, const int16_t *d_samples, int32_t rows, int32_t cols, float hscale, float vscale, float border,
                               float slope_threshold, float friction, float restitution);

/* ---- PPO ---------------------------------------------------------------------------------------------------- */
typedef struct {
    int32_t num_envs, num_steps;             /* N (local), T */
    int32_t num_obs, num_pri_obs, num_actions;
    int32_t actor_hidden[3], critic_hidden[3];
    int32_t num_learning_epochs, num_mini_batches;
    float clip_param, gamma, lam, value_loss_coef, entropy_coef;
    float learning_rate, learning_rate_min, learning_rate_max, desired_kl, max_grad_norm;
    int32_t adaptive_schedule, use_clipped_value_loss;
    float init_noise_std;
    int32_t use_tensor_cores;                /* 1: tcgen05 kind::tf32 GEMMs for the dense layers, 0: fp32 SIMT GEMMs */
    int32_t world_size;                      /* gradient / KL sums are divided by this after the caller's all-reduce */
    int32_t env_id_offset;                   /* global index of local env 0: the rollout's Normal.sample() stream is keyed by
                                              * (seed, global env id, step) so W shards draw what one GPU with all envs would */
    int32_t comm_timeout_ms;                 /* NVLink all-reduce: give up waiting for a peer after this long (0 = 10000) and report
                                              * through the control block (`comm_error`); the optimiser step is then NOT applied */
    uint64_t seed;                           /* Philox key of the action noise (fast mode, d_eps == NULL); the task seed */
} grx_ppo_cfg;

typedef struct grx_ppo grx_ppo;

int grx_ppo_create(const grx_ppo_cfg *cfg, int32_t device, grx_ppo **out);
int grx_ppo_destroy(grx_ppo *ppo);
/* views: params grads adam_m adam_v (flat, reference state_dict order: std, actor.model.{0,2,4,6}.{weight,bias},
 * critic...), obs critic_obs actions values rewards dones actions_log_prob mu sigma returns advantages,
 * reduce_buf (grads + [kl_sum, count, surrogate_sum, value_loss_sum] tail; what the caller all-reduces), gsum, adv_moments,
 * ctl (device control block: lr, clip coefficient, skip, Adam step, ..., comm_error), mb_log [epochs*minibatches, 4] =
 * (mean KL, learning rate, loss, gradient norm) of every minibatch of the most recent update (ppo.py:262-268, 308-309) */
int grx_ppo_get_buffer(grx_ppo *ppo, const char *name, grx_buffer *out);
/* PPO.act (ppo.py:144-171) for rollout step t: actor+critic forward, a = mu + sigma*eps, log-prob; everything stored
 * into row t of the rollout storage (base_storage.py:80-100).  d_eps [N, A] standard normal draws or NULL (Philox). */
int grx_ppo_act(grx_ppo *ppo, const float *d_obs, const float *d_critic_obs, const float *d_eps, int32_t t,
                float *d_actions_out, uint64_t step_index, void *stream);
/* PPO.process_env_step (ppo.py:177-196): rewards += gamma * V * time_out; store rewards / dones in row t. */
int grx_ppo_process_env_step(grx_ppo *ppo, const float *d_rewards, const uint8_t *d_dones, const uint8_t *d_time_outs,
                             int32_t t, void *stream);
/* PPO.compute_returns (ppo.py:198-205, base_storage.py:120-141): critic(last obs), reverse GAE scan, advantage
 * normalisation.  For multi-GPU the three moments [sum, sum of squares, count] are left in `adv_moments` between
 * grx_ppo_compute_returns_local and grx_ppo_normalize_advantages so the caller can all-reduce them. */
int grx_ppo_compute_returns(grx_ppo *ppo, const float *d_last_critic_obs, void *stream);
int grx_ppo_compute_returns_local(grx_ppo *ppo, const float *d_last_critic_obs, void *stream);
int grx_ppo_normalize_advantages(grx_ppo *ppo, void *stream);
/* One minibatch of PPO.update (ppo.py:244-305), split so a gradient all-reduce can sit in between:
 *   grads: gather rows d_indices[mb*B .. (mb+1)*B) of the rollout, forward, losses, backward -> reduce_buf
 *   apply: KL -> adaptive LR (device scalar, no .item()), NaN-skip, global-norm clip, Adam */
int grx_ppo_minibatch_grads(grx_ppo *ppo, const int64_t *d_indices, int32_t mb, void *stream);
int grx_ppo_minibatch_apply(grx_ppo *ppo, void *stream);
/* Whole PPO.update (ppo.py:215-321): epochs x minibatches of grads+apply replayed from one CUDA graph.  With world_size > 1
 * it needs the NVLink all-reduce (grx_ppo_comm_open) and must be called by every rank. */
int grx_ppo_update(grx_ppo *ppo, const int64_t *d_indices, void *stream);
/* Multi-GPU gradient all-reduce over NVLink peer memory (no reference counterpart: the reference is single-GPU).
 *   grx_ppo_comm_handle: writes this rank's 64-byte cudaIpcMemHandle_t of its comm block [reduce_buf | gsum | flags];
 *   grx_ppo_comm_open:   handles = world x 64 bytes in rank order (exchanged by the caller over any host channel); maps every
 *                        peer's block.  All ranks must have returned from it (host barrier) before the first apply.
 *   grx_ppo_minibatch_apply_comm: apply step whose first kernel loads this rank's slice of the gradient from every peer, sums it
 *                        in rank order, pushes the sum into every peer's gsum and reduces the gradient norm on the way. */
int grx_ppo_comm_handle(grx_ppo *ppo, void *out64);
int grx_ppo_comm_open(grx_ppo *ppo, int32_t rank, int32_t world, const void *handles);
int grx_ppo_minibatch_apply_comm(grx_ppo *ppo, void *stream);
/* actor-only forward for play.py / get_inference_policy (actor_critic_mlp.py:209-217) */
int grx_ppo_act_inference(grx_ppo *ppo, const float *d_obs, int32_t n, float *d_actions_out, void *stream);

/* Profiling (GRX_PPO_TIMING=1 in the environment at create): mean microseconds per minibatch of the stepwise grads + apply
 * entries, split [memset+gather, actor forward, critic forward, heads/loss, actor backward, critic backward, apply]; resets. */
int grx_ppo_debug_timing(grx_ppo *ppo, float *out7, int32_t *count);

/* Debug / parity: one dense-layer GEMM on device pointers through the fp32 SIMT kernel (use_tc = 0) or the tcgen05 TF32 kernel.
 *   variant 0: C[M,N] = A[M,K] B[N,K]^T + bias (epi 0) or elu(...) (epi 1)      — forward (mlp.py:40-41)
 *   variant 1: C[M,N] = (A[M,K] B[K,N]) * ELU'(aux[M,N])                         — input gradient
 *   variant 2: C[M,N] += A[K,M]^T B[K,N], bias_out[M] += column sums of A        — weight / bias gradient (split-K) */
int grx_gemm_debug(int32_t variant, int32_t epi, int32_t M, int32_t N, int32_t K, const float *A, const float *B, float *C,
                   const float *bias, const float *aux, float *bias_out, int32_t splits, int32_t use_tc, void *stream);

/* Test / profiling: the three hidden layers of the registered policy run as ONE chained tcgen05 kernel per forward pass (csrc/grx_mlp_chain.cuh);
 * fwd_chain = 0 falls back to one grouped GEMM launch per layer, 1 enables, -1 only queries.  Returns the previous setting. */
int grx_ppo_debug_fused(int32_t fwd_chain);
/* Test / profiling: the layer-pipelined launches (several dependent dense layers in ONE persistent tcgen05 launch; GRX_LAYER_PIPE) on (1) / off (0)
 * at run time, -1 = query; returns the previous setting. */
int grx_ppo_debug_pipe(int32_t on);
/* Test (host arithmetic only, no GPU needed): the plan of a split-K weight-gradient group of np problems [M_i x N_i] with contraction length K on
 * `sms` SMs — macro tile and split count chosen together so that the tile list fills whole rounds; out3 = {row blocks per tile (1|2), columns per
 * tile (128|256), splits}. */
int grx_gemm_debug_dw_plan(const int32_t *M, const int32_t *N, int32_t K, int32_t np, int32_t sms, int32_t *out3);

/* Test / profiling: force the macro tile of the tensor-core GEMM (row_blocks in {1, 2} x bn in {32, 64, 128, 256}; 0, 0 = cost model). */
int grx_gemm_debug_tile(int32_t row_blocks, int32_t bn);

/* Profiling: %globaltimer stamps (ns) of CTA 0 of the most recent tensor-core GEMM launch: entry, setup done, first TMA issued,
 * first stage landed, last MMA committed, first accumulator complete, epilogue done, unused, then epilogue detail of warp 0's
 * first chunk: TMEM load done, staging written, proxy fence done, TMA store issued, all tiles stored.  16 values.  Synchronises. */
int grx_gemm_debug_stamps(uint64_t *out16);
/* Profiling: the same 16 values + 16 more; [16..23] = the gradient all-reduce kernel of the most recent minibatch, per phase: entry, ready
 * barrier passed, slices reduced and pushed, done flags published.  Synchronises. */
int grx_debug_stamps32(uint64_t *out32);

#ifdef __cplusplus
}
#endif
#endif /* GRX_B200_H */
