// grx_task.cuh — the TASK half of one policy step (everything of LeggedRobot.step that is not the dynamics), shared by the fused lower-limb env
// kernel (grx_env.cu, compile-time record layout LayC<10>) and the generic-topology env kernel (grx_phys_generic.cu, run-time layout LayR for
// any DOF count <= 32, e.g. the full-body 32-DOF GR1T1 / GR1T2 of gr1t1_config.py:10-307):
//   post_physics_step (legged_robot.py:269-305, legged_robot_fftai.py:90-133) = state update, _get_heights (LR:1235-1274), push (LR:786-797),
//   air / land timers, check_termination (LR:336-353), compute_reward with the 24 active terms (LR:355-375, gr1t1.py:338-589),
//   reset_idx + terrain curriculum (LR:377-440, 799-826, FF:137-146), compute_observations + noise (LR:442-481, FF:148-167, G1:281-336).
// One warp per robot; `rec` is the robot's state record staged in shared memory.  The same source serves both kernels, so the parity of the
// registered task against the reference goldens (tests/test_env_gpu.py) carries over to every other DOF count.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "grx_b200.h"
#include "grx_terrain.cuh"

namespace grx {   // plain data shared across translation units

struct LayR {
    int nd, root, dofpos, dofvel, lastdofvel, lastact, lastlastact, cmd, bho, air, land, clast, eplen, tlevel, origin, ttype, sums, rec_f;
    int c_motor, c_bi, c_fric, c_rest, cst_f;
    int u_noise, u_reset_dof, u_reset_xy, u_reset_yaw, u_reset_vel, u_cmd_time, u_cmd_reset, u_push, u_curriculum, rng_k;
};
// Arguments of one step / reset launch (both env kernels)
struct EnvArgs {
    float *rec;
    const float *cst;
    const void *model;             // ModelDev (grx_env.cu) or GModel (grx_phys_generic.cu)
    TerrainDev terrain;
    const float *terrain_origins;  // [t_rows, t_cols, 3]
    int t_rows, t_cols;
    int N;
    const float *actions;
    const float *U;  // nullable
    float delay;
    int push;
    unsigned long long step_index;
    float *obs, *pri_obs, *rew, *torques, *contact_forces, *foot_state;
    float *episode_accum, *episode_accum_next;  // this launch's slot of the extras ring, and the slot to clear for the next launch
    unsigned char *reset, *time_out;
    grx_injected_physics inj;
    float *dbg_M, *dbg_h;  // debug_dynamics
    int dbg_index;
    float *rigid_body_states, *dof_state;   // compat exports or nullptr
    long long *ep_len64;
    const int *link_body;                   // [nl] per-link tables (global memory; only read when rigid_body_states is exported)
    const float *link_pos, *link_rot;       // [nl, 3] [nl, 9]
    unsigned long long *dbg_sig;   // [N, dbg_sig_stride] active-set signature per substep, or nullptr (grx_env_debug_active_sig)
    int dbg_sig_stride;
};

}  // namespace grx

namespace {
using grx::EnvArgs;
using grx::LayR;

constexpr int TK_NREW = 24, TK_NF = 2;
constexpr int ACC_RING = 256, ACC_W = 32;   // extras["episode"] accumulators: one 32-float slot per launch, ring of 256
constexpr unsigned TK_FULL = 0xffffffffu;

// ---- per-env state record (floats; ints stored bit-wise), parameter record and uniform-draw slots (grx_b200/rng_layout.py) for `nd` DOF.
// nd = 10: 108-float record (432 B), 24-float parameters, 68 draws; nd = 32: 216 / 44 / 156.
template <int ND_>
struct LayC {
    static constexpr int nd = ND_;
    static constexpr int root = 0, dofpos = 16, dofvel = 16 + ND_, lastdofvel = 16 + 2 * ND_, lastact = 16 + 3 * ND_, lastlastact = 16 + 4 * ND_,
                         cmd = 16 + 5 * ND_, bho = cmd + 3, air = bho + 1, land = air + 2, clast = land + 2, eplen = clast + 2, tlevel = eplen + 1,
                         origin = tlevel + 1, ttype = origin + 3, sums = ttype + 1, rec_f = (sums + TK_NREW + 3) & ~3;
    static constexpr int c_motor = 0, c_bi = ND_, c_fric = ND_ + 10, c_rest = ND_ + 11, cst_f = (ND_ + 12 + 3) & ~3;
    static constexpr int u_noise = 0, u_reset_dof = 9 + 3 * ND_, u_reset_xy = u_reset_dof + ND_, u_reset_yaw = u_reset_xy + 2, u_reset_vel = u_reset_yaw + 1,
                         u_cmd_time = u_reset_vel + 6, u_cmd_reset = u_cmd_time + 3, u_push = u_cmd_reset + 3, u_curriculum = u_push + 2,
                         rng_k = u_curriculum + 2;
};
__host__ __device__ inline LayR make_layout(int nd) {
    LayR l;
    l.nd = nd; l.root = 0; l.dofpos = 16; l.dofvel = 16 + nd; l.lastdofvel = 16 + 2 * nd; l.lastact = 16 + 3 * nd; l.lastlastact = 16 + 4 * nd;
    l.cmd = 16 + 5 * nd; l.bho = l.cmd + 3; l.air = l.bho + 1; l.land = l.air + 2; l.clast = l.land + 2; l.eplen = l.clast + 2; l.tlevel = l.eplen + 1;
    l.origin = l.tlevel + 1; l.ttype = l.origin + 3; l.sums = l.ttype + 1; l.rec_f = (l.sums + TK_NREW + 3) & ~3;
    l.c_motor = 0; l.c_bi = nd; l.c_fric = nd + 10; l.c_rest = nd + 11; l.cst_f = (nd + 12 + 3) & ~3;
    l.u_noise = 0; l.u_reset_dof = 9 + 3 * nd; l.u_reset_xy = l.u_reset_dof + nd; l.u_reset_yaw = l.u_reset_xy + 2; l.u_reset_vel = l.u_reset_yaw + 1;
    l.u_cmd_time = l.u_reset_vel + 6; l.u_cmd_reset = l.u_cmd_time + 3; l.u_push = l.u_cmd_reset + 3; l.u_curriculum = l.u_push + 2;
    l.rng_k = l.u_curriculum + 2;
    return l;
}

__device__ __forceinline__ float tk_warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(TK_FULL, v, o);
    return v;
}

// ---- Philox4x32-10 (counter-based draws for fast mode)
__device__ __noinline__ void philox4(uint32_t k0, uint32_t k1, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t *out) {
#pragma unroll
    for (int r = 0; r < 10; r++) {
        uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
struct Draw {
    const float *U;  // row of this env or nullptr
    uint32_t k0, k1, gid, step_lo, step_hi;
    __device__ __forceinline__ float operator()(int slot) const {
        if (U) return U[slot];
        uint32_t r[4];
        philox4(k0, k1, gid, step_lo, step_hi, (uint32_t)(slot >> 2), r);
        return (float)(r[slot & 3] >> 8) * (1.0f / 16777216.0f);  // 24-bit mantissa uniform in [0,1), like torch.rand
    }
};

// torch_utils.py:72-81 quat_rotate_inverse
__device__ __forceinline__ void quat_rotate_inverse(const float *q, const float *v, float *o) {
    float qw = q[3];
    float s = 2.0f * qw * qw - 1.0f;
    const float cx[3] = {q[1] * v[2] - q[2] * v[1], q[2] * v[0] - q[0] * v[2], q[0] * v[1] - q[1] * v[0]};
    float d = q[0] * v[0] + q[1] * v[1] + q[2] * v[2];
#pragma unroll
    for (int k = 0; k < 3; k++) o[k] = v[k] * s - cx[k] * qw * 2.0f + q[k] * d * 2.0f;
}

template <class LAY>
__device__ __forceinline__ void resample_commands(float *rec, const LAY &L, const grx_task_cfg &cfg, const Draw &draw, int base) {
    // legged_robot.py:650-677 ((upper - lower) * u + lower; small xy commands zeroed)
    float cx = (cfg.cmd_range[0][1] - cfg.cmd_range[0][0]) * draw(base) + cfg.cmd_range[0][0];
    float cy = (cfg.cmd_range[1][1] - cfg.cmd_range[1][0]) * draw(base + 1) + cfg.cmd_range[1][0];
    const float keep = sqrtf(cx * cx + cy * cy) > 0.1f ? 1.f : 0.f;
    rec[L.cmd] = cx * keep;
    rec[L.cmd + 1] = cy * keep;
    rec[L.cmd + 2] = (cfg.cmd_range[2][1] - cfg.cmd_range[2][0]) * draw(base + 2) + cfg.cmd_range[2][0];
}

// ---- reset of one env (warp-cooperative): _update_terrain_curriculum LR:799-826, _reset_dofs LR:717-740,
// _reset_root_states LR:742-784, _resample_commands LR:402, buffer zeroing LR:405-415 + FF:137-146, episode sums -> extras LR:420-424
template <class LAY, class MODEL>
__device__ __forceinline__ void reset_env(float *rec, const LAY &L, const MODEL &m, const EnvArgs &A, const grx_task_cfg &cfg, const Draw &draw,
                                          int lane, float cnorm, bool curriculum_active) {
    if (cfg.curriculum && curriculum_active) {
        if (lane == 0) {
            const float dx = rec[L.root] - rec[L.origin], dy = rec[L.root + 1] - rec[L.origin + 1];
            const float distance = sqrtf(dx * dx + dy * dy);
            const bool up = distance > cfg.terrain_env_length / 2.f;
            const bool down = (distance < cnorm * cfg.max_episode_length_s * 0.5f) && !up;
            int level = __float_as_int(rec[L.tlevel]) + (up ? 1 : 0) - (down ? 1 : 0);
            if (level >= A.t_rows) level = min((int)floorf(draw(L.u_curriculum) * (float)A.t_rows), A.t_rows - 1);
            else level = max(level, 0);
            rec[L.tlevel] = __int_as_float(level);
            const int type = __float_as_int(rec[L.ttype]);
            const float *org = A.terrain_origins + ((size_t)level * A.t_cols + type) * 3;
            rec[L.origin] = org[0]; rec[L.origin + 1] = org[1]; rec[L.origin + 2] = org[2];
        }
        __syncwarp();
    }
    if (lane < L.nd) {                                                                  // _reset_dofs LR:717-734
        rec[L.dofpos + lane] = cfg.randomize_init_dof_pos ? ((1.5f - 0.5f) * draw(L.u_reset_dof + lane) + 0.5f) * m.q0[lane] : m.q0[lane];
        rec[L.dofvel + lane] = 0.f;
        rec[L.lastact + lane] = 0.f;
        rec[L.lastdofvel + lane] = 0.f;
        rec[L.lastlastact + lane] = 0.f;
    }
    if (lane == 0) {                                                                  // _reset_root_states LR:742-779
        float *rt = rec + L.root;
#pragma unroll
        for (int k = 0; k < 13; k++) rt[k] = cfg.base_init_state[k];
#pragma unroll
        for (int k = 0; k < 3; k++) rt[k] += rec[L.origin + k];
        if (cfg.custom_origins) {
            rt[0] += (1.0f - -1.0f) * draw(L.u_reset_xy) + -1.0f;
            rt[1] += (1.0f - -1.0f) * draw(L.u_reset_xy + 1) + -1.0f;
        }
        const float yaw = 12.566370614359172f * draw(L.u_reset_yaw) + -6.283185307179586f;
        float sy, cy;
        sincosf(yaw * 0.5f, &sy, &cy);
        rt[3] = 0.f; rt[4] = 0.f; rt[5] = sy; rt[6] = cy;
        if (cfg.randomize_init_base_velocity) {
#pragma unroll
            for (int k = 0; k < 6; k++) rt[7 + k] = (0.5f - -0.5f) * draw(L.u_reset_vel + k) + -0.5f;
        }
        resample_commands(rec, L, cfg, draw, L.u_cmd_reset);                                 // LR:402
        rec[L.eplen] = __int_as_float(0);
    }
    if (lane < TK_NF) { rec[L.air + lane] = 0.f; rec[L.land + lane] = 0.f; rec[L.clast + lane] = 0.f; }
    if (lane < TK_NREW) {                                                                // extras["episode"] sums LR:420-424
        atomicAdd(A.episode_accum + lane, rec[L.sums + lane]);
        rec[L.sums + lane] = 0.f;
    }
    if (lane == TK_NREW) atomicAdd(A.episode_accum + TK_NREW, 1.0f);
    __syncwarp();
}

// Everything after the physics for one robot (see the file header).  Inputs: the clipped actions (lane j < nd: action j / the previous one), the
// substep averages of the foot force / foot velocity and the foot height (lane f < 2: foot f), the torso quaternion, the last substep's torques
// tau[nd] and net contact forces cf[nl * 3] (shared memory).  Scratch (shared memory, per warp): mh[H] measured heights, ob[num_obs] noise-free
// observation, pri[num_pri_obs] privileged-observation row (left there for the caller to store), rterm[24].  Writes obs, rew, reset, time_out,
// torques and the compat exports; updates rec in place (the caller stores it).
template <class LAY, class MODEL>
__device__ __forceinline__ void task_post_physics(float *rec, const LAY &L, const MODEL &m, const EnvArgs &A, const grx_task_cfg &cfg, const Draw &draw,
                                                  const int lane, const int e, const float act_l, const float last_act_l, const float ff_acc,
                                                  const float (&fl_acc)[3], const float foot_z, const float (&torso_q)[4], const float *tau,
                                                  const float *cf, float *mh, float *ob, float *pri, float *rterm) {
    const int nd = L.nd;
    // =====================================================================================================
    // post_physics_step (legged_robot.py:269-305, legged_robot_fftai.py:90-133)
    // =====================================================================================================
    const float dtp = (float)cfg.decimation * cfg.sim_dt;
    int ep_len = __float_as_int(rec[L.eplen]) + 1;                                   // LR:282
    float base_quat[4], v_b[3], w_b[3], g_b[3];
    const float gvec[3] = {0.f, 0.f, -1.f};
#pragma unroll
    for (int k = 0; k < 4; k++) base_quat[k] = rec[L.root + 3 + k];
    quat_rotate_inverse(base_quat, rec + L.root + 7, v_b);                           // LR:309-311
    quat_rotate_inverse(base_quat, rec + L.root + 10, w_b);
    quat_rotate_inverse(base_quat, gvec, g_b);
    const float root_pos[3] = {rec[L.root], rec[L.root + 1], rec[L.root + 2]};
    __syncwarp();
    if (ep_len % cfg.resample_interval == 0) {                                        // LR:317-318
        if (lane == 0) resample_commands(rec, L, cfg, draw, L.u_cmd_time);
    }
    __syncwarp();
    // ---- _get_heights (legged_robot.py:1235-1274): trunc-to-int grid index, min of 3 samples
    const int H = cfg.num_height_points;
    if (cfg.measure_heights && A.terrain.type != 0) {
        float qz = base_quat[2], qw = base_quat[3];
        const float nrm = fmaxf(sqrtf(qz * qz + qw * qw), 1e-9f);
        qz /= nrm; qw /= nrm;
        for (int k = lane; k < H; k += 32) {
            const float px_ = cfg.measured_points_x[k / cfg.n_points_y], py_ = cfg.measured_points_y[k % cfg.n_points_y];
            // quat_apply with q = (0, 0, qz, qw), b = (px, py, 0): t = 2 * cross(xyz, b); out = b + w t + cross(xyz, t)
            const float tx = -qz * py_ * 2.f, ty = qz * px_ * 2.f;
            const float ox = px_ + qw * tx + (-qz * ty), oy = py_ + qw * ty + (qz * tx);
            float fx = ox + root_pos[0], fy = oy + root_pos[1];
            fx += A.terrain.border; fy += A.terrain.border;
            long long ix = (long long)(fx / A.terrain.hscale), iy = (long long)(fy / A.terrain.hscale);
            ix = min(max(ix, 0LL), (long long)A.terrain.rows - 2);
            iy = min(max(iy, 0LL), (long long)A.terrain.cols - 2);
            const short *p = A.terrain.h + (size_t)ix * A.terrain.cols + iy;
            const short h1 = __ldg(p), h2 = __ldg(p + A.terrain.cols), h3 = __ldg(p + 1);
            const short hm = min(min(h1, h2), h3);
            mh[k] = (float)hm * A.terrain.vscale;
        }
    } else {
        for (int k = lane; k < H; k += 32) mh[k] = 0.f;
    }
    __syncwarp();
    if (A.push && lane < 2) {                                                         // LR:333-334, 786-797
        const float mv = cfg.max_push_vel_xy;
        rec[L.root + 7 + lane] = (mv - -mv) * draw(L.u_push + lane) + -mv;
    }
    // ---- feet bookkeeping (FF:108-133): lane f < TK_NF owns foot f
    bool contact = false, filt = false, first = false;
    float air = 0.f, land = 0.f, fh_sum = 0.f, fxy = 0.f, fz = 0.f;
    {
        // sum_k (foot_z - mh[k]) for both feet, and sum_k clip(z - target - mh[k]) for the base: lanes stride over k
        const float fz0 = __shfl_sync(TK_FULL, foot_z, 0), fz1 = __shfl_sync(TK_FULL, foot_z, 1);
        float s0 = 0.f, s1 = 0.f;
        for (int k = lane; k < H; k += 32) { s0 += fz0 - mh[k]; s1 += fz1 - mh[k]; }
        s0 = tk_warp_sum(s0); s1 = tk_warp_sum(s1);
        fh_sum = lane == 0 ? s0 : s1;
    }
    const float feet_h = fh_sum / (float)H;   // valid on lanes 0,1
    if (lane < TK_NF) {
        const float *f = cf + 3 * m.foot_link[lane];
        fz = f[2];
        fxy = sqrtf(f[0] * f[0] + f[1] * f[1]);
        contact = fz > 1.0f;
        const bool lastc = rec[L.clast + lane] != 0.f;
        filt = contact || lastc;
        air = rec[L.air + lane];
        first = (air > 0.f) && filt;
        air += dtp;
        land = (rec[L.land + lane] + dtp) * (contact ? 1.f : 0.f);
    }
    // ---- check_termination (LR:336-353)
    bool term = false;
    for (int l = lane; l < m.nl; l += 32) {
        if ((m.term_mask >> l) & 1ull) {
            const float *f = cf + 3 * l;
            term |= sqrtf(f[0] * f[0] + f[1] * f[1] + f[2] * f[2]) > 1.0f;
        }
    }
    bool reset = __any_sync(TK_FULL, term);
    reset |= fabsf(g_b[2]) < 0.33f;
    const bool time_out = (float)ep_len > cfg.max_episode_length;
    reset |= time_out;

    // ---- compute_reward (LR:355-375): per-DOF sums by warp reduction, then the 24 terms
    float sa = 0, sadd = 0, sacc = 0, stor = 0, spose = 0, slpos = 0, slvel = 0, sltor = 0;
    float q_l = 0, qd_l = 0, tau_l = 0;
    if (lane < nd) {
        q_l = rec[L.dofpos + lane]; qd_l = rec[L.dofvel + lane]; tau_l = tau[lane];
        const float la = last_act_l, lla = rec[L.lastlastact + lane], asc = cfg.action_scale;
        const float e1 = (la - act_l) * asc, e0 = (lla - la) * asc;
        sa = fabsf(e1);
        sadd = fabsf(e1 - e0);
        sacc = fabsf((qd_l - rec[L.lastdofvel + lane]) / dtp);
        stor = fabsf(tau_l);
        spose = fabsf(q_l - m.q0[lane]);
        float ool = -fminf(q_l - m.soft_lower[lane], 0.f);
        ool += fmaxf(q_l - m.soft_upper[lane], 0.f);
        slpos = fabsf(ool);
        slvel = fminf(fmaxf(fabsf(qd_l) - m.dof_vel_limit[lane] * cfg.soft_dof_vel_limit, 0.f), 1.f);
        sltor = fmaxf(fabsf(tau_l) - m.dof_effort[lane] * cfg.soft_torque_limit, 0.f);
    }
    sa = tk_warp_sum(sa); sadd = tk_warp_sum(sadd); sacc = tk_warp_sum(sacc); stor = tk_warp_sum(stor);
    spose = tk_warp_sum(spose); slpos = tk_warp_sum(slpos); slvel = tk_warp_sum(slvel); sltor = tk_warp_sum(sltor);
    // foot quantities broadcast to every lane
    const float lfh = __shfl_sync(TK_FULL, feet_h, 0), rfh = __shfl_sync(TK_FULL, feet_h, 1);
    const float air0 = __shfl_sync(TK_FULL, air, 0), air1 = __shfl_sync(TK_FULL, air, 1);
    const float land0 = __shfl_sync(TK_FULL, land, 0), land1 = __shfl_sync(TK_FULL, land, 1);
    const float ff0 = __shfl_sync(TK_FULL, ff_acc, 0), ff1 = __shfl_sync(TK_FULL, ff_acc, 1);
    const float vx0 = __shfl_sync(TK_FULL, fl_acc[0], 0), vy0 = __shfl_sync(TK_FULL, fl_acc[1], 0);
    const float vx1 = __shfl_sync(TK_FULL, fl_acc[0], 1), vy1 = __shfl_sync(TK_FULL, fl_acc[1], 1);
    const float fxy0 = __shfl_sync(TK_FULL, fxy, 0), fxy1 = __shfl_sync(TK_FULL, fxy, 1);
    const float fz0 = __shfl_sync(TK_FULL, fz, 0), fz1 = __shfl_sync(TK_FULL, fz, 1);
    const unsigned cbal = __ballot_sync(TK_FULL, contact), fbal = __ballot_sync(TK_FULL, first);
    const float tau_a0 = m.ankle_torque(tau, 0), tau_a1 = m.ankle_torque(tau, 1);   // sum |tau| over the left / right half of the ankle DOF (gr1t1.py:406-411)
    const float cmdx = rec[L.cmd], cmdy = rec[L.cmd + 1], cmdw = rec[L.cmd + 2];
    const float cnorm = sqrtf(cmdx * cmdx + cmdy * cmdy);
    const float nz = cnorm > 0.1f ? 1.f : 0.f;
    float rew = 0.f;
    if (lane == 0) {
        float r[TK_NREW];
        const float bh = rec[L.bho];   // one step stale on purpose
        float tg[3];
        quat_rotate_inverse(torso_q, gvec, tg);
        const float tgt = cfg.swing_feet_height_target, q4 = tgt / 4.f;
        const float mid0 = fabsf(air0 - cfg.feet_air_time_target / 2.f), mid1 = fabsf(air1 - cfg.feet_air_time_target / 2.f);
        const float mn = fminf(lfh, rfh);
        r[0] = 1.f - expf(cfg.sigma_action_diff * sa);                                                    // action_diff
        r[1] = 1.f - expf(cfg.sigma_action_diff_diff * sadd);                                             // action_diff_diff
        r[2] = expf(cfg.sigma_cmd_diff_ang_vel_yaw * fabsf(cmdw - w_b[2]));                               // cmd_diff_ang_vel_yaw
        r[3] = expf(cfg.sigma_cmd_diff_base_height * (fabsf(bh) * (bh < 0.f ? 1.f : 0.f)));               // cmd_diff_base_height
        r[4] = expf(cfg.sigma_cmd_diff_base_orient * (fabsf(g_b[0]) + fabsf(g_b[1])));                    // cmd_diff_base_orient
        r[5] = expf(cfg.sigma_cmd_diff_lin_vel_x * fabsf(cmdx - v_b[0]));                                 // cmd_diff_lin_vel_x
        r[6] = expf(cfg.sigma_cmd_diff_lin_vel_y * fabsf(cmdy - v_b[1]));                                 // cmd_diff_lin_vel_y
        r[7] = expf(cfg.sigma_cmd_diff_lin_vel_z * fabsf(0.f - v_b[2]));                                  // cmd_diff_lin_vel_z
        r[8] = expf(cfg.sigma_cmd_diff_torso_orient * (fabsf(tg[0]) + fabsf(tg[1])));                     // cmd_diff_torso_orient
        r[9] = 1.f - expf(cfg.sigma_dof_acc_new * sacc);                                                  // dof_acc_new
        {
            const float el = tau_a0 * fabsf(lfh) * (lfh > tgt / 2.f ? 1.f : 0.f);
            const float er = tau_a1 * fabsf(rfh) * (rfh > tgt / 2.f ? 1.f : 0.f);
            r[10] = 1.f - expf(cfg.sigma_dof_tor_ankle_feet_lift_up * (el + er));                         // dof_tor_ankle_feet_lift_up
        }
        r[11] = 1.f - expf(cfg.sigma_dof_tor_new * stor);                                                 // dof_tor_new
        r[12] = expf(cfg.sigma_feet_air_force * (mid0 * ff0 + mid1 * ff1)) * nz;                          // feet_air_force
        r[13] = expf(cfg.sigma_feet_air_height * (mid0 * fabsf(lfh - mn - tgt) + mid1 * fabsf(rfh - mn - tgt))) * nz;  // feet_air_height
        r[14] = (expf(cfg.sigma_feet_air_time * fabsf(air0 - cfg.feet_air_time_target)) * ((fbal & 1u) ? 1.f : 0.f) +
                 expf(cfg.sigma_feet_air_time * fabsf(air1 - cfg.feet_air_time_target)) * ((fbal & 2u) ? 1.f : 0.f)) * nz;  // feet_air_time
        {
            const float e0 = (land0 - cfg.feet_land_time_max) * (land0 > cfg.feet_land_time_max ? 1.f : 0.f);
            const float e1 = (land1 - cfg.feet_land_time_max) * (land1 > cfg.feet_land_time_max ? 1.f : 0.f);
            r[15] = ((1.f - expf(cfg.sigma_feet_land_time * e0)) + (1.f - expf(cfg.sigma_feet_land_time * e1))) * nz;  // feet_land_time
        }
        {
            const float cl = fabsf(lfh - q4) * (lfh < q4 ? 1.f : 0.f) / q4, cr = fabsf(rfh - q4) * (rfh < q4 ? 1.f : 0.f) / q4;
            const float e_ = sqrtf(vx0 * vx0 + vy0 * vy0) * cl + sqrtf(vx1 * vx1 + vy1 * vy1) * cr;
            r[16] = expf(cfg.sigma_feet_speed_xy_close_to_ground * e_);                                   // feet_speed_xy_close_to_ground
        }
        {
            float el = fxy0 - cfg.feet_stumble_ratio * fabsf(fz0), er = fxy1 - cfg.feet_stumble_ratio * fabsf(fz1);
            el = el * (el > 0.f ? 1.f : 0.f); er = er * (er > 0.f ? 1.f : 0.f);
            r[17] = (1.f - expf(cfg.sigma_feet_stumble * el)) + (1.f - expf(cfg.sigma_feet_stumble * er));  // feet_stumble
        }
        r[18] = 1.f - expf(cfg.sigma_limits_dof_pos * slpos);                                             // limits_dof_pos
        r[19] = 1.f - expf(cfg.sigma_limits_dof_tor * sltor);                                             // limits_dof_tor
        r[20] = 1.f - expf(cfg.sigma_limits_dof_vel * slvel);                                             // limits_dof_vel
        r[21] = (cbal & 3u) == 0u ? 1.f : 0.f;                                                            // on_the_air
        r[22] = expf(cfg.sigma_pose_offset * spose);                                                      // pose_offset
        r[23] = expf(cfg.sigma_stand_still * spose) * (cnorm < 0.1f ? 1.f : 0.f);                         // stand_still
#pragma unroll
        for (int k = 0; k < TK_NREW; k++) {
            const float v = r[k] * cfg.reward_scale[k];
            rew += v;                   // alphabetical summation order (App. B-15)
            rterm[k] = v;
        }
    }
    __syncwarp();
    if (lane < TK_NREW) rec[L.sums + lane] += rterm[lane];                             // LR:366
    __syncwarp();

    // ---- reset_idx (LR:377-440, FF:137-146), curriculum (LR:799-826)
    bool contact_for_obs = contact;
    if (reset) {
        reset_env(rec, L, m, A, cfg, draw, lane, cnorm, true);
        ep_len = 0;
        if (lane < TK_NF) { air = 0.f; land = 0.f; contact_for_obs = false; }
    }
    if (lane == 0 && cfg.curriculum) atomicAdd(A.episode_accum + TK_NREW + 1, (float)__float_as_int(rec[L.tlevel]));   // LR:427
    __syncwarp();

    // ---- compute_observations (LR:442-452, FF:148-167, G1:281-313) — after the reset, with stale base quantities (App. B-2)
    const float hm = cfg.obs_scale_height;
    const float z_new = rec[L.root + 2];
    float bsum = 0.f;
    const int O = cfg.num_obs;
    for (int k = lane; k < H; k += 32) {
        const float off = fminf(fmaxf(z_new - cfg.base_height_target - mh[k], -1.f), 1.f) * hm;
        bsum += off;
        pri[O + 8 + k] = off * hm;                                                    // surround_heights_offset * 5 (x25 net, App. B-7)
    }
    bsum = tk_warp_sum(bsum);
    const float bho = bsum / (float)H;
    if (lane < 3) {
        ob[lane] = rec[L.cmd + lane] * 1.0f;                                         // commands * commands_scale (ones, G1:125)
        ob[3 + lane] = w_b[lane] * cfg.obs_scale_ang_vel;
        ob[6 + lane] = g_b[lane] * cfg.obs_scale_gravity;
    }
    if (lane < nd) {
        ob[9 + lane] = (rec[L.dofpos + lane] - m.q0[lane]) * cfg.obs_scale_dof_pos;
        ob[9 + nd + lane] = rec[L.dofvel + lane] * cfg.obs_scale_dof_vel;
        ob[9 + 2 * nd + lane] = act_l * cfg.obs_scale_action;
    }
    __syncwarp();
    const float co = cfg.clip_observations;
    for (int i = lane; i < O; i += 32) {
        const float v = ob[i];
        pri[i] = fminf(fmaxf(v, -co), co);                                            // privileged obs embeds the noise-free obs
        float vn = v;
        if (cfg.add_noise) vn += (2.f * draw(L.u_noise + i) - 1.f) * cfg.noise_scale_vec[i];   // LR:478-481
        A.obs[(size_t)e * O + i] = fminf(fmaxf(vn, -co), co);                        // LR:240-241
    }
    if (lane < 3) pri[O + lane] = fminf(fmaxf(v_b[lane] * cfg.obs_scale_lin_vel, -co), co);
    if (lane == 3) pri[O + 3] = fminf(fmaxf(bho * hm, -co), co);
    if (lane < TK_NF) {
        pri[O + 4 + lane] = contact_for_obs ? 1.f : 0.f;
        pri[O + 6 + lane] = fminf(fmaxf(feet_h * hm, -co), co);
    }
    __syncwarp();
    for (int k = lane; k < H; k += 32) pri[O + 8 + k] = fminf(fmaxf(pri[O + 8 + k], -co), co);

    // ---- carry-over (LR:299-300, FF:94-97) and outputs
    if (lane < nd) {
        rec[L.lastact + lane] = act_l;
        rec[L.lastlastact + lane] = act_l;                                            // == last_actions (App. B-1)
        rec[L.lastdofvel + lane] = rec[L.dofvel + lane];
        A.torques[(size_t)e * nd + lane] = tau[lane];
    }
    if (lane < TK_NF) {
        rec[L.air + lane] = air * (filt ? 0.f : 1.f);
        rec[L.land + lane] = land;
        rec[L.clast + lane] = contact_for_obs ? 1.f : 0.f;
    }
    if (lane == 0) {
        rec[L.bho] = bho;
        rec[L.eplen] = __int_as_float(ep_len);
        A.rew[e] = rew;
        A.reset[e] = reset ? 1 : 0;
        A.time_out[e] = time_out ? 1 : 0;
    }
    if (A.contact_forces != nullptr)
        for (int i = lane; i < m.nl * 3; i += 32) A.contact_forces[(size_t)e * m.nl * 3 + i] = cf[i];
    if (A.dof_state != nullptr && lane < nd) {   // compat export: interleaved (pos, vel), post-reset like the reference's dof_state after reset_idx
        reinterpret_cast<float2 *>(A.dof_state)[(size_t)e * nd + lane] = make_float2(rec[L.dofpos + lane], rec[L.dofvel + lane]);
    }
    if (A.ep_len64 != nullptr && lane == 0) A.ep_len64[e] = (long long)ep_len;
}

}  // namespace
