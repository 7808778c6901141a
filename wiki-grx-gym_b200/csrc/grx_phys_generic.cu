// grx_phys_generic.cu — dynamics spec "GRX-dyn v1" for ANY revolute tree (floating base, <= 36 bodies / 32 DOF), sm_100a: the full-body
// 32-DOF GR1T1 / GR1T2 (33 bodies: legs 2 x 6, waist 3, head 3, arms 2 x 7; gr1t1_config.py:10-307) incl. robot SELF-COLLISION
// (legged_robot_config.py:121 self_collisions = 0 = enabled, create_actor(..., collision_filter = 0) legged_robot.py:1022-1028).
//
// The fused env kernel (grx_env.cu) is specialised to the registered lower-limb topology (2 chains of 5, 16 velocity DOF: Cholesky in the
// registers of a half-warp, <= 31 constraint rows = one lane each).  This kernel is the general statement of the same equations, one warp
// per robot, everything in shared memory, in the structure of the CPU oracle (oracle/phys_impl.h: composite-rigid-body mass matrix about the
// base origin, dense Cholesky, velocity-space projected Gauss-Seidel over contact + self-contact + joint-limit rows, semi-implicit Euler).
// It stands behind the same reference calls: legged_robot_fftai.py:51-88 (substep loop, action delay, foot averages),
// legged_robot.py:679-715 (_compute_torques), legged_robot_fftai.py:67-76 (set_dof_actuation_force_tensor / simulate / refresh_*).
//
// C ABI (include/grx_b200.h): grx_physg_create / _set_terrain_* / _step / _destroy; _step has the signature of the oracle's
// grx_oracle_physics_step on device pointers, so parity tests compare the two call for call.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <string.h>

#include <string>
#include <vector>

#include "grx_b200.h"
#include "grx_count.h"
#include "grx_terrain.cuh"
#include "grx_task.cuh"
#include "grx_envg.h"

int grx_set_error(int code, const std::string &msg);   // grx_env.cu

#define CK(call)                                                                                             \
    do {                                                                                                     \
        cudaError_t err__ = (call);                                                                          \
        if (err__ != cudaSuccess)                                                                            \
            return grx_set_error(GRX_E_CUDA, std::string(#call) + ": " + cudaGetErrorString(err__));        \
    } while (0)

namespace {

constexpr int GB = 36, GD = 32, GV = 38, GS = 32, GL = 48, GP = 64, GDEPTH = 12;
constexpr int GKC = 8, GKS = 4, GKL = 8, GROWS = 3 * (GKC + GKS) + GKL;   // 44 constraint rows at most
constexpr int GPITCH = 41;            // row pitch of J / Y (odd: per-lane rows stay conflict-free)
constexpr int GWARPS = 10;            // warps (robots) per CTA at most: 10 x 22.2 KB of workspace fit the 227 KB of an SM (4096 robots on 148 SMs = 3 waves)
constexpr unsigned FULL = 0xffffffffu;

struct GModel {   // device copy of the model, in global memory (read through the read-only path)
    int nb, nd, nl, ns, nf, npairs, maxdepth;
    int parent[GB], depth[GB];
    signed char path[GB][GDEPTH];        // bodies from the first joint below the base down to the body itself
    unsigned long long anc[GB];          // bit j set: joint j (body j + 1) is on the path base -> body
    float jpos[GB][3], jrot[GB][9], axis[GB][3], mass[GB], com[GB][3], inertia[GB][6];
    float dof_lower[GD], dof_upper[GD], dof_vel_limit[GD], dof_effort[GD], kp[GD], kd[GD], q0[GD];
    int link_body[GL];
    float link_pos[GL][3], link_rot[GL][9];
    int sph_body[GS], sph_link[GS];
    float sph_pos[GS][3], sph_rad[GS];
    int foot_link[4];
    int pair_a[GP], pair_b[GP];
    // task tables (the env kernel only; grx_task.cuh reads them through the same member names as the lower-limb ModelDev)
    float soft_lower[GD], soft_upper[GD];
    unsigned long long term_mask;
    int ankle_dof[4], nankle;   // first half: left leg, second half: right leg (gr1t1.py:406-411); full body: ankle pitch + roll per leg
    int torso_link;
    __device__ __forceinline__ float ankle_torque(const float *tau, int side) const {
        const int half = nankle >> 1;
        float s = 0.f;
        for (int k = side * half; k < (side + 1) * half; k++) s += fabsf(tau[ankle_dof[k]]);
        return s;
    }
};

struct GArgs {
    const GModel *m;
    TerrainDev terrain;
    grx_physg_cfg cfg;
    int N;
    float *root, *q, *qd;
    const float *actions, *last_actions, *motor, *binert, *friction, *restitution;
    float delay;
    float *torques, *link_state, *contact_force, *avg_ff, *avg_fl, *avg_fa;
    unsigned long long *sig;
};

// Lower triangle of the symmetric (nv x nv) mass matrix / its Cholesky factor, packed by rows: element (i, j), i >= j
__device__ __forceinline__ int tri(int i, int j) { return ((i * (i + 1)) >> 1) + j; }

struct alignas(16) GWS {
    float root[16], q[GD], qd[GD], tau[GD], bin[12];
    // Lifetimes inside a substep: the kinematics arrays are dead once the constraint Jacobians have been built, the composite / joint-space
    // scratch once the mass matrix exists; Y = M^-1 J^T is written after both (the right-hand side row Y[GROWS - 1] lands on `sub`, which is
    // dead after g_mass_and_bias) -> they share storage (9.2 KB vs 7.2 KB).
    union {
        struct {
            float R[GB][9], o[GB][3], a[GB][3], c[GB][3], w[GB][3], vo[GB][3], al[GB][3], ao[GB][3], Iw[GB][6];
            float sub[GB][16];                   // per body, then subtree: mass, first moment 3, inertia 6, bias force 3, bias moment 3
            float Sl[GB][3], Sa[GB][3], Ff[GB][3], Fn[GB][3];
        };
        float Y[GROWS][GPITCH];
    };
    float M[GV * (GV + 1) / 2];           // packed lower triangle, see tri()
    float h[GV + 2], u[GV + 2];
    float J[GROWS][GPITCH];
    float Ad[GROWS], bias[GROWS], lam[GROWS];
    float cfr[GKC + GKS][9], cpt[GKC + GKS][4];
    int cbody[GKC + GKS], cbody2[GKC + GKS], clink[GKC + GKS], clink2[GKC + GKS];
    float cf[GL * 3];
    float facc[4][8];                    // per foot: sum |F|, sum |v| 3, sum |w| 3
};
static_assert(sizeof(float) * GB * (9 + 7 * 3 + 6) <= sizeof(float) * (GROWS - 1) * GPITCH, "the right-hand side row Y[GROWS - 1] must not land on the kinematics arrays (live until the Jacobians are built)");
static_assert(GWARPS * sizeof(GWS) <= 232448, "shared memory per CTA");

__device__ __forceinline__ void cross3(const float *a, const float *b, float *o) {
    float x = a[1] * b[2] - a[2] * b[1], y = a[2] * b[0] - a[0] * b[2], z = a[0] * b[1] - a[1] * b[0];
    o[0] = x; o[1] = y; o[2] = z;
}
__device__ __forceinline__ float dot3(const float *a, const float *b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
__device__ __forceinline__ void m3v(const float *R, const float *v, float *o) {
    float x = R[0] * v[0] + R[1] * v[1] + R[2] * v[2], y = R[3] * v[0] + R[4] * v[1] + R[5] * v[2], z = R[6] * v[0] + R[7] * v[1] + R[8] * v[2];
    o[0] = x; o[1] = y; o[2] = z;
}
__device__ __forceinline__ void m3m(const float *A, const float *B, float *C) {
    float t[9];
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) t[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
#pragma unroll
    for (int i = 0; i < 9; i++) C[i] = t[i];
}
__device__ __forceinline__ void quat2mat(const float *q, float *R) {
    float x = q[0], y = q[1], z = q[2], w = q[3];
    R[0] = 1 - 2 * (y * y + z * z); R[1] = 2 * (x * y - z * w);     R[2] = 2 * (x * z + y * w);
    R[3] = 2 * (x * y + z * w);     R[4] = 1 - 2 * (x * x + z * z); R[5] = 2 * (y * z - x * w);
    R[6] = 2 * (x * z - y * w);     R[7] = 2 * (y * z + x * w);     R[8] = 1 - 2 * (x * x + y * y);
}
__device__ __forceinline__ void mat2quat(const float *R, float *q) {
    float tr = R[0] + R[4] + R[8];
    if (tr > 0) { float s = sqrtf(tr + 1) * 2; q[3] = s / 4; q[0] = (R[7] - R[5]) / s; q[1] = (R[2] - R[6]) / s; q[2] = (R[3] - R[1]) / s; }
    else if (R[0] > R[4] && R[0] > R[8]) { float s = sqrtf(1 + R[0] - R[4] - R[8]) * 2; q[3] = (R[7] - R[5]) / s; q[0] = s / 4; q[1] = (R[1] + R[3]) / s; q[2] = (R[2] + R[6]) / s; }
    else if (R[4] > R[8]) { float s = sqrtf(1 + R[4] - R[0] - R[8]) * 2; q[3] = (R[2] - R[6]) / s; q[0] = (R[1] + R[3]) / s; q[1] = s / 4; q[2] = (R[5] + R[7]) / s; }
    else { float s = sqrtf(1 + R[8] - R[0] - R[4]) * 2; q[3] = (R[3] - R[1]) / s; q[0] = (R[2] + R[6]) / s; q[1] = (R[5] + R[7]) / s; q[2] = s / 4; }
}
__device__ __forceinline__ void axang2mat(const float *a, float th, float *R) {
    float s, c;
    sincosf(th, &s, &c);
    float t = 1 - c;
    R[0] = c + a[0] * a[0] * t;        R[1] = a[0] * a[1] * t - a[2] * s; R[2] = a[0] * a[2] * t + a[1] * s;
    R[3] = a[1] * a[0] * t + a[2] * s; R[4] = c + a[1] * a[1] * t;        R[5] = a[1] * a[2] * t - a[0] * s;
    R[6] = a[2] * a[0] * t - a[1] * s; R[7] = a[2] * a[1] * t + a[0] * s; R[8] = c + a[2] * a[2] * t;
}
__device__ __forceinline__ void sym6v(const float *S, const float *v, float *o) {
    float x = S[0] * v[0] + S[3] * v[1] + S[4] * v[2], y = S[3] * v[0] + S[1] * v[1] + S[5] * v[2], z = S[4] * v[0] + S[5] * v[1] + S[2] * v[2];
    o[0] = x; o[1] = y; o[2] = z;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}

// ---- forward kinematics + body velocities + velocity-product accelerations (oracle kinematics()): every body walks its own path from the base
__device__ void g_kinematics(GWS &s, const GModel &m, int lane) {
    for (int b = lane; b < m.nb; b += 32) {
        float R[9], o[3], w[3], vo[3], al[3] = {0, 0, 0}, ao[3] = {0, 0, 0}, a[3] = {0, 0, 0};
        quat2mat(s.root + 3, R);
#pragma unroll
        for (int k = 0; k < 3; k++) { o[k] = s.root[k]; vo[k] = s.root[7 + k]; w[k] = s.root[10 + k]; }
        const int depth = m.depth[b];
        for (int d = 0; d < depth; d++) {
            const int j = m.path[b][d];
            const float qj = s.q[j - 1], qdj = s.qd[j - 1];
            float Rj[9], Rq[9], r[3], an[3], t1[3], t2[3], t3[3], t4[3];
            m3m(R, m.jrot[j], Rj);
            axang2mat(m.axis[j], qj, Rq);
            m3v(R, m.jpos[j], r);
            m3m(Rj, Rq, R);
            m3v(R, m.axis[j], an);
            cross3(w, r, t1); cross3(al, r, t2); cross3(w, t1, t3); cross3(w, an, t4);
#pragma unroll
            for (int i = 0; i < 3; i++) {
                o[i] += r[i]; vo[i] += t1[i]; ao[i] += t2[i] + t3[i];
                al[i] += t4[i] * qdj; w[i] += an[i] * qdj; a[i] = an[i];
            }
        }
        const float *com = b == 0 ? s.bin + 1 : m.com[b];
        const float *I6 = b == 0 ? s.bin + 4 : m.inertia[b];
        float rc[3];
        m3v(R, com, rc);
        float I[9] = {I6[0], I6[3], I6[4], I6[3], I6[1], I6[5], I6[4], I6[5], I6[2]}, T[9], Rt[9];
        m3m(R, I, T);
#pragma unroll
        for (int i = 0; i < 3; i++)
#pragma unroll
            for (int j = 0; j < 3; j++) Rt[3 * i + j] = R[3 * j + i];
        m3m(T, Rt, T);
#pragma unroll
        for (int i = 0; i < 9; i++) s.R[b][i] = R[i];
#pragma unroll
        for (int i = 0; i < 3; i++) {
            s.o[b][i] = o[i]; s.a[b][i] = a[i]; s.c[b][i] = o[i] + rc[i];
            s.w[b][i] = w[i]; s.vo[b][i] = vo[i]; s.al[b][i] = al[i]; s.ao[b][i] = ao[i];
        }
        s.Iw[b][0] = T[0]; s.Iw[b][1] = T[4]; s.Iw[b][2] = T[8]; s.Iw[b][3] = T[1]; s.Iw[b][4] = T[2]; s.Iw[b][5] = T[5];
    }
    __syncwarp();
}

// ---- mass matrix (internal order: joints, base linear, base angular) + bias vector (oracle mass_and_bias())
__device__ void g_mass_and_bias(GWS &s, const GModel &m, float gravity, int lane) {
    const int nb = m.nb, nd = m.nd, nv = nd + 6;
    for (int i = lane; i < GV * (GV + 1) / 2; i += 32) s.M[i] = 0.0f;
    for (int b = lane; b < nb; b += 32) {
        const float mass = b == 0 ? s.bin[0] : m.mass[b];
        float r[3], rc[3];
#pragma unroll
        for (int k = 0; k < 3; k++) { r[k] = s.c[b][k] - s.o[0][k]; rc[k] = s.c[b][k] - s.o[b][k]; }
        const float rr = dot3(r, r);
        float *x = s.sub[b];
        x[0] = mass; x[1] = mass * r[0]; x[2] = mass * r[1]; x[3] = mass * r[2];
        x[4] = s.Iw[b][0] + mass * (rr - r[0] * r[0]); x[5] = s.Iw[b][1] + mass * (rr - r[1] * r[1]); x[6] = s.Iw[b][2] + mass * (rr - r[2] * r[2]);
        x[7] = s.Iw[b][3] - mass * r[0] * r[1]; x[8] = s.Iw[b][4] - mass * r[0] * r[2]; x[9] = s.Iw[b][5] - mass * r[1] * r[2];
        float t1[3], t2[3], ac[3], Iw_[3], Ial[3], g3[3], f[3];
        cross3(s.al[b], rc, t1); cross3(s.w[b], rc, t2); cross3(s.w[b], t2, t2);
#pragma unroll
        for (int k = 0; k < 3; k++) ac[k] = s.ao[b][k] + t1[k] + t2[k];
        ac[2] -= gravity;
#pragma unroll
        for (int k = 0; k < 3; k++) f[k] = mass * ac[k];
        sym6v(s.Iw[b], s.w[b], Iw_); sym6v(s.Iw[b], s.al[b], Ial);
        cross3(s.w[b], Iw_, g3); cross3(r, f, t1);
#pragma unroll
        for (int k = 0; k < 3; k++) { x[10 + k] = f[k]; x[13 + k] = Ial[k] + g3[k] + t1[k]; }
    }
    __syncwarp();
    for (int b = nb - 1; b >= 1; b--) {   // subtree sums in the oracle's order (children have larger indices than their parents)
        if (lane < 16) s.sub[m.parent[b]][lane] += s.sub[b][lane];
        __syncwarp();
    }
    for (int j = 1 + lane; j < nb; j += 32) {   // joint columns
        const float *ci = s.sub[j];
        float d[3], Sl[3], Sa[3], t1[3], t2[3], t3[3];
#pragma unroll
        for (int k = 0; k < 3; k++) { d[k] = s.o[0][k] - s.o[j][k]; Sa[k] = s.a[j][k]; }
        cross3(Sa, d, Sl);
        cross3(Sa, ci + 1, t1);
        cross3(ci + 1, Sl, t2);
        sym6v(ci + 4, Sa, t3);
#pragma unroll
        for (int k = 0; k < 3; k++) {
            s.Sl[j][k] = Sl[k]; s.Sa[j][k] = Sa[k];
            s.Ff[j][k] = ci[0] * Sl[k] + t1[k]; s.Fn[j][k] = t2[k] + t3[k];
        }
        s.h[j - 1] = dot3(Sl, ci + 10) + dot3(Sa, ci + 13);
    }
    __syncwarp();
    for (int j = 1 + lane; j < nb; j += 32) {
        for (int i = j; i >= 1; i = m.parent[i]) {   // i ancestor-or-self of j
            const float v = dot3(s.Sl[i], s.Ff[j]) + dot3(s.Sa[i], s.Fn[j]);
            s.M[tri(j - 1, i - 1)] = v;   // i is an ancestor of (or) j: i <= j
        }
#pragma unroll
        for (int k = 0; k < 3; k++) {
            s.M[tri(nd + k, j - 1)] = s.Ff[j][k];
            s.M[tri(nd + 3 + k, j - 1)] = s.Fn[j][k];
        }
    }
    if (lane == 0) {
        const float *ci = s.sub[0];
        const float mm = ci[0], *hh = ci + 1, *I = ci + 4;
        const int L = nd, A = nd + 3;
        const float hx[9] = {0, -hh[2], hh[1], hh[2], 0, -hh[0], -hh[1], hh[0], 0};
        for (int k = 0; k < 3; k++) s.M[tri(L + k, L + k)] = mm;
        for (int i = 0; i < 3; i++)
            for (int j = 0; j < 3; j++) s.M[tri(A + i, L + j)] = hx[3 * i + j];
        s.M[tri(A, A)] = I[0]; s.M[tri(A + 1, A + 1)] = I[1]; s.M[tri(A + 2, A + 2)] = I[2];
        s.M[tri(A + 1, A)] = I[3];
        s.M[tri(A + 2, A)] = I[4];
        s.M[tri(A + 2, A + 1)] = I[5];
        for (int k = 0; k < 6; k++) s.h[L + k] = ci[10 + k];
    }
    (void)nv;
    __syncwarp();
}

// ---- dense Cholesky in shared memory (lower factor in place), right-looking; lane owns rows lane, lane + 32
__device__ void g_cholesky(GWS &s, int n, int lane) {
    for (int k = 0; k < n; k++) {
        const float piv = sqrtf(s.M[tri(k, k)]);
        __syncwarp();
        for (int i = k + lane; i < n; i += 32) s.M[tri(i, k)] = i == k ? piv : s.M[tri(i, k)] / piv;
        __syncwarp();
        for (int i = k + 1 + lane; i < n; i += 32) {
            const float lik = s.M[tri(i, k)];
            float *row = s.M + tri(i, 0);
            for (int j = k + 1; j <= i; j++) row[j] -= lik * s.M[tri(j, k)];
        }
        __syncwarp();
    }
}
// x <- M^-1 x for the row `x` (one lane per right-hand side; reads of the factor are warp broadcasts)
__device__ void g_chol_solve(const GWS &s, int n, float *x) {
    for (int i = 0; i < n; i++) {
        float t = x[i];
        const float *row = s.M + tri(i, 0);
        for (int p = 0; p < i; p++) t -= row[p] * x[p];
        x[i] = t / row[i];
    }
    for (int i = n - 1; i >= 0; i--) {
        float t = x[i];
        for (int p = i + 1; p < n; p++) t -= s.M[tri(p, i)] * x[p];
        x[i] = t / s.M[tri(i, i)];
    }
}

// Jacobian row of world point x on body b along direction d, written by the whole warp into row[0 .. nv) (scaled by sgn, accumulated when acc)
__device__ void g_point_jac_row(const GWS &s, const GModel &m, int b, const float *x, const float *d, float sgn, bool acc, float *row, int lane) {
    const int nd = m.nd;
    const unsigned long long anc = m.anc[b];
    for (int j = lane; j < nd; j += 32) {
        float v = 0.f;
        if ((anc >> j) & 1ull) {
            float r[3], t[3];
#pragma unroll
            for (int k = 0; k < 3; k++) r[k] = x[k] - s.o[j + 1][k];
            cross3(s.a[j + 1], r, t);
            v = dot3(t, d);
        }
        row[j] = (acc ? row[j] : 0.f) + sgn * v;
    }
    if (lane < 6) {
        float r[3], t[3];
#pragma unroll
        for (int k = 0; k < 3; k++) r[k] = x[k] - s.o[0][k];
        cross3(r, d, t);
        const float v = lane < 3 ? d[lane] : t[lane - 3];
        row[nd + lane] = (acc ? row[nd + lane] : 0.f) + sgn * v;
    }
    __syncwarp();
}

__device__ void g_substep(GWS &s, const GModel &m, const GArgs &A, int lane, float mu_env, float rest_env, unsigned long long *sig_out) {
    const grx_physg_cfg &cfg = A.cfg;
    const int nd = m.nd, nv = nd + 6;
    const float dt = cfg.sim_dt;
    g_mass_and_bias(s, m, cfg.gravity, lane);
    g_cholesky(s, nv, lane);
    // ---- unconstrained update u* = u + dt M^-1 (tau - h): the right-hand side goes into row GROWS - 1 of Y and is solved further down TOGETHER with
    // the constraint rows (one lane per right-hand side) — a lone single-lane solve here cost as much as the whole batch
    float *rhs = s.Y[GROWS - 1];
    for (int i = lane; i < nv; i += 32) rhs[i] = (i < nd ? s.tau[i] : 0.f) - s.h[i];
    __syncwarp();
    const float mu = 0.5f * (mu_env + A.terrain.friction), rest = 0.5f * (rest_env + A.terrain.restitution);
    unsigned long long sig_item = 0ull;
    // ---- ground contacts: lane = sphere
    bool act = false;
    float n[3] = {0, 0, 1}, xs[3] = {0, 0, 0}, dist = 0, rad = 0;
    int sb = 0, cell[3] = {0, 0, 0};
    if (lane < m.ns) {
        sb = m.sph_body[lane];
        rad = m.sph_rad[lane];
        m3v(s.R[sb], m.sph_pos[lane], xs);
#pragma unroll
        for (int k = 0; k < 3; k++) xs[k] += s.o[sb][k];
        float hgt;
        terrain_query(A.terrain, xs[0], xs[1], hgt, n, cell);
        dist = (xs[2] - hgt) * n[2] - rad;
        act = dist < cfg.contact_offset;
    }
    const unsigned bal = __ballot_sync(FULL, act);
    const int rank = __popc(bal & ((1u << lane) - 1u));
    const int nc = min(__popc(bal), min(cfg.max_contacts, GKC));
    auto frame_and_target = [&](const float *nn, float d_, float vn0, float rest_, float *fr, float &target, bool &bounce, bool &usey) {
        usey = nn[0] > 0.9f || nn[0] < -0.9f;
        const float e[3] = {usey ? 0.f : 1.f, usey ? 1.f : 0.f, 0.f};
        const float dn = dot3(e, nn);
        float t1[3], t2[3];
#pragma unroll
        for (int k = 0; k < 3; k++) t1[k] = e[k] - dn * nn[k];
        const float inv = 1.0f / sqrtf(dot3(t1, t1));
#pragma unroll
        for (int k = 0; k < 3; k++) t1[k] *= inv;
        cross3(nn, t1, t2);
#pragma unroll
        for (int k = 0; k < 3; k++) { fr[k] = nn[k]; fr[3 + k] = t1[k]; fr[6 + k] = t2[k]; }
        if (d_ > 0) target = -d_ / dt;
        else { target = -d_ * cfg.erp / dt; if (target > cfg.max_depen_vel) target = cfg.max_depen_vel; }
        bounce = vn0 < -cfg.bounce_threshold && -rest_ * vn0 > target;
        if (bounce) target = -rest_ * vn0;
    };
    if (act && rank < nc) {
        float xc[3], rv[3], vc[3];
#pragma unroll
        for (int k = 0; k < 3; k++) { xc[k] = xs[k] - n[k] * rad; rv[k] = xc[k] - s.o[sb][k]; }
        cross3(s.w[sb], rv, vc);
        const float vn0 = dot3(n, s.vo[sb]) + dot3(n, vc);
        float target; bool bounce, usey;
        frame_and_target(n, dist, vn0, rest, s.cfr[rank], target, bounce, usey);
        sig_item = mix64((1ull << 56) | (unsigned long long)lane | ((unsigned long long)cell[0] << 6) | ((unsigned long long)cell[1] << 18) |
                         ((unsigned long long)cell[2] << 30) | ((unsigned long long)(bounce ? 1 : 0) << 33) | ((unsigned long long)(usey ? 1 : 0) << 34));
#pragma unroll
        for (int k = 0; k < 3; k++) s.cpt[rank][k] = xc[k];
        s.cpt[rank][3] = target;
        s.cbody[rank] = sb; s.cbody2[rank] = -1; s.clink[rank] = m.sph_link[lane]; s.clink2[rank] = -1;
    }
    __syncwarp();
    // ---- robot self-collision: lane = candidate pair (two passes for up to 64 pairs), the first max_self_contacts in pair order
    int nself = 0;
    const int max_self = min(cfg.max_self_contacts, GKS);
    for (int base = 0; base < m.npairs && nself < max_self; base += 32) {
        const int pi = base + lane;
        bool hit = false;
        float xa[3], xb[3], nn[3] = {0, 0, 1}, d_ = 0, dd = 1;
        int ba = 0, bb = 0, sa = 0, sb2 = 0;
        if (pi < m.npairs) {
            sa = m.pair_a[pi]; sb2 = m.pair_b[pi]; ba = m.sph_body[sa]; bb = m.sph_body[sb2];
            m3v(s.R[ba], m.sph_pos[sa], xa); m3v(s.R[bb], m.sph_pos[sb2], xb);
            float dv[3];
#pragma unroll
            for (int k = 0; k < 3; k++) { xa[k] += s.o[ba][k]; xb[k] += s.o[bb][k]; dv[k] = xa[k] - xb[k]; }
            dd = sqrtf(dot3(dv, dv));
            d_ = dd - m.sph_rad[sa] - m.sph_rad[sb2];
            hit = d_ < cfg.contact_offset && dd > 1e-9f;
#pragma unroll
            for (int k = 0; k < 3; k++) nn[k] = dv[k] / dd;
        }
        const unsigned hb = __ballot_sync(FULL, hit);
        const int hr = nself + __popc(hb & ((1u << lane) - 1u));
        if (hit && hr < max_self) {
            const int c = nc + hr;
            const float mid = m.sph_rad[sb2] + 0.5f * d_;
            float xc[3], ra[3], rb[3], va[3], vb[3];
#pragma unroll
            for (int k = 0; k < 3; k++) { xc[k] = xb[k] + nn[k] * mid; ra[k] = xc[k] - s.o[ba][k]; rb[k] = xc[k] - s.o[bb][k]; }
            cross3(s.w[ba], ra, va); cross3(s.w[bb], rb, vb);
            const float vn0 = dot3(nn, s.vo[ba]) + dot3(nn, va) - dot3(nn, s.vo[bb]) - dot3(nn, vb);
            float target; bool bounce, usey;
            frame_and_target(nn, d_, vn0, rest_env, s.cfr[c], target, bounce, usey);
            sig_item += mix64((3ull << 56) | (unsigned long long)pi | ((unsigned long long)(bounce ? 1 : 0) << 33) | ((unsigned long long)(usey ? 1 : 0) << 34));
#pragma unroll
            for (int k = 0; k < 3; k++) s.cpt[c][k] = xc[k];
            s.cpt[c][3] = target;
            s.cbody[c] = ba; s.cbody2[c] = bb; s.clink[c] = m.sph_link[sa]; s.clink2[c] = m.sph_link[sb2];
        }
        nself = min(nself + __popc(hb), max_self);
    }
    __syncwarp();
    const int ncon = nc + nself;
    // ---- joint-limit rows (predicted with the pre-step rate): lane = joint
    float lsgn = 0, ltgt = 0;
    if (lane < nd) {
        const float qq = s.q[lane], qn = qq + dt * s.qd[lane];
        if (qn < m.dof_lower[lane]) { lsgn = 1.f; ltgt = (m.dof_lower[lane] - qq) / dt; }
        else if (qn > m.dof_upper[lane]) { lsgn = -1.f; ltgt = (qq - m.dof_upper[lane]) / dt; }
    }
    const unsigned lbal = __ballot_sync(FULL, lsgn != 0.f);
    const int lrank = __popc(lbal & ((1u << lane) - 1u));
    const int nlim = min(__popc(lbal), 7);   // MAXLIM of the oracle
    const int nrow = 3 * ncon + nlim;
    if (lsgn != 0.f && lrank < nlim) {
        const int r = 3 * ncon + lrank;
        for (int i = 0; i < nv; i++) s.J[r][i] = 0.f;
        s.J[r][lane] = lsgn;
        s.bias[r] = ltgt;
        sig_item += mix64((2ull << 56) | (unsigned long long)lane | ((unsigned long long)(lsgn < 0.f ? 1 : 0) << 6));
    }
    if (sig_out != nullptr) {
        unsigned long long tot = sig_item;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const unsigned l2 = __shfl_xor_sync(FULL, (unsigned)tot, o), h2 = __shfl_xor_sync(FULL, (unsigned)(tot >> 32), o);
            tot += ((unsigned long long)h2 << 32) | l2;
        }
        if (lane == 0) *sig_out = tot;
    }
    __syncwarp();
    // ---- contact rows: J = J_a (- J_b for a self-contact), three directions per contact
    for (int c = 0; c < ncon; c++) {
        for (int k = 0; k < 3; k++) {
            float *row = s.J[3 * c + k];
            g_point_jac_row(s, m, s.cbody[c], s.cpt[c], s.cfr[c] + 3 * k, 1.f, false, row, lane);
            if (s.cbody2[c] >= 0) g_point_jac_row(s, m, s.cbody2[c], s.cpt[c], s.cfr[c] + 3 * k, -1.f, true, row, lane);
            if (lane == 0) s.bias[3 * c + k] = k == 0 ? s.cpt[c][3] : 0.f;
        }
    }
    __syncwarp();
    // ---- Y = M^-1 J^T (one lane per row, two passes), Ad = J . Y; "row" nrow = the unconstrained right-hand side (nrow <= GROWS - 1: its own row)
    for (int r = lane; r <= nrow; r += 32) {
        const bool con = r < nrow;
        float *y = con ? s.Y[r] : rhs;
        if (con)
            for (int i = 0; i < nv; i++) y[i] = s.J[r][i];
        g_chol_solve(s, nv, y);   // ONE call site: the lane of the right-hand side runs converged with the constraint lanes
        if (con) {
            float a = 0.f;
            for (int i = 0; i < nv; i++) a += s.J[r][i] * y[i];
            s.Ad[r] = a;
            s.lam[r] = 0.f;
        }
    }
    __syncwarp();
    for (int i = lane; i < nv; i += 32) s.u[i] = (i < nd ? s.qd[i] : s.root[7 + i - nd]) + dt * rhs[i];
    __syncwarp();
    // ---- projected Gauss-Seidel in velocity space (the oracle's sweep: contacts — normal, then 2 friction rows — then the limits).  The velocity
    // vector lives in registers (lane i: u[i] and, for i < nv - 32, u[32 + i]); every lane computes the same multiplier update from the
    // warp-uniform row product, so lam needs no cross-lane hand-over (each lane reads back its own identical store) and the sweep runs without
    // a single warp barrier; the next row's J / Y elements are fetched while the butterfly of the current row is in flight.
    {
        const bool hi = lane < nv - 32;
        float u0 = lane < nv ? s.u[lane] : 0.f, u1 = hi ? s.u[32 + lane] : 0.f;
        float j0 = 0.f, j1 = 0.f, y0 = 0.f, y1 = 0.f;
        if (nrow > 0) { j0 = lane < nv ? s.J[0][lane] : 0.f; j1 = hi ? s.J[0][32 + lane] : 0.f; y0 = lane < nv ? s.Y[0][lane] : 0.f; y1 = hi ? s.Y[0][32 + lane] : 0.f; }
        for (int it = 0; it < cfg.solver_iters; it++) {
            float lam_n = 0.f;   // multiplier of the current contact's normal row (bounds its two friction rows)
            for (int r = 0; r < nrow; r++) {
                float v = j0 * u0 + j1 * u1;
                const float cy0 = y0, cy1 = y1;
                const int rn = r + 1 < nrow ? r + 1 : 0;   // next row (wraps to the first row of the next sweep)
                j0 = lane < nv ? s.J[rn][lane] : 0.f; j1 = hi ? s.J[rn][32 + lane] : 0.f;
                y0 = lane < nv ? s.Y[rn][lane] : 0.f; y1 = hi ? s.Y[rn][32 + lane] : 0.f;
                v = warp_sum(v);
                const float l0 = s.lam[r];
                float ln = l0 - (v - s.bias[r]) / s.Ad[r];
                if (r < 3 * ncon) {
                    const int c = r / 3, k = r - 3 * c;
                    if (k == 0) { ln = fmaxf(ln, 0.f); lam_n = ln; }
                    else { const float lim = (s.cbody2[c] >= 0 ? mu_env : mu) * lam_n; ln = fminf(fmaxf(ln, -lim), lim); }
                } else ln = fmaxf(ln, 0.f);
                const float dl = ln - l0;
                s.lam[r] = ln;
                u0 += cy0 * dl; u1 += cy1 * dl;
            }
        }
        __syncwarp();
        if (lane < nv) s.u[lane] = u0;
        if (hi) s.u[32 + lane] = u1;
        __syncwarp();
    }
    // ---- net contact force per URDF link (world frame, on the body) = impulse / dt; a self-contact reacts on its second link
    for (int i = lane; i < m.nl * 3; i += 32) s.cf[i] = 0.f;
    __syncwarp();
    if (lane < 3) {
        for (int c = 0; c < ncon; c++) {
            const float *f = s.cfr[c];
            const float fk = (f[lane] * s.lam[3 * c] + f[3 + lane] * s.lam[3 * c + 1] + f[6 + lane] * s.lam[3 * c + 2]) / dt;
            s.cf[3 * s.clink[c] + lane] += fk;
            if (s.clink2[c] >= 0) s.cf[3 * s.clink2[c] + lane] -= fk;
        }
    }
    // ---- joint-rate limit + integrate
    if (lane < nd) {
        const float vl = m.dof_vel_limit[lane];
        const float v = fminf(fmaxf(s.u[lane], -vl), vl);
        s.qd[lane] = v;
        s.q[lane] += dt * v;
    }
    if (lane < 6) s.root[7 + lane] = s.u[nd + lane];
    __syncwarp();
    if (lane == 0) {
        float *rt = s.root;
#pragma unroll
        for (int k = 0; k < 3; k++) rt[k] += dt * rt[7 + k];
        const float wx = rt[10], wy = rt[11], wz = rt[12];
        const float wn = sqrtf(wx * wx + wy * wy + wz * wz), th = wn * dt;
        float sn, cs;
        sincosf(0.5f * th, &sn, &cs);
        const float sc = wn > 1e-9f ? sn / wn : 0.5f * dt;
        const float dq[4] = {wx * sc, wy * sc, wz * sc, cs};
        float *p = rt + 3;
        const float qx = dq[3] * p[0] + dq[0] * p[3] + dq[1] * p[2] - dq[2] * p[1];
        const float qy = dq[3] * p[1] - dq[0] * p[2] + dq[1] * p[3] + dq[2] * p[0];
        const float qz = dq[3] * p[2] + dq[0] * p[1] - dq[1] * p[0] + dq[2] * p[3];
        const float qw = dq[3] * p[3] - dq[0] * p[0] - dq[1] * p[1] - dq[2] * p[2];
        const float nn = 1.0f / sqrtf(qx * qx + qy * qy + qz * qz + qw * qw);
        p[0] = qx * nn; p[1] = qy * nn; p[2] = qz * nn; p[3] = qw * nn;
    }
    __syncwarp();
}

__device__ __forceinline__ void g_link_state(const GWS &s, const GModel &m, int l, float *out) {
    const int b = m.link_body[l];
    float r[3], R[9], t[3];
    m3v(s.R[b], m.link_pos[l], r);
    m3m(s.R[b], m.link_rot[l], R);
    mat2quat(R, out + 3);
    cross3(s.w[b], r, t);
#pragma unroll
    for (int k = 0; k < 3; k++) { out[k] = s.o[b][k] + r[k]; out[7 + k] = s.vo[b][k] + t[k]; out[10 + k] = s.w[b][k]; }
}

// One policy step of physics (the body of during_physics_step, legged_robot_fftai.py:51-88): decimation x [PD torque -> substep], foot averages.
__global__ void __launch_bounds__(GWARPS * 32, 1) physg_step_kernel(const __grid_constant__ GArgs A) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    GWS &s = *reinterpret_cast<GWS *>(smem_raw + (size_t)warp * sizeof(GWS));
    const GModel &m = *A.m;
    const int e = blockIdx.x * (blockDim.x >> 5) + warp;
    if (e >= A.N) return;
    const int nd = m.nd, nl = m.nl, nf = m.nf;
    if (lane < 13) s.root[lane] = A.root[(size_t)e * 13 + lane];
    if (lane < nd) { s.q[lane] = A.q[(size_t)e * nd + lane]; s.qd[lane] = A.qd[(size_t)e * nd + lane]; }
    if (lane < 10) s.bin[lane] = A.binert[(size_t)e * 10 + lane];
    if (lane < 4 * 8) (&s.facc[0][0])[lane] = 0.f;
    __syncwarp();
    const float mu_env = A.friction[e], rest_env = A.restitution[e];
    const grx_physg_cfg &cfg = A.cfg;
    g_kinematics(s, m, lane);
    for (int deci = 0; deci < cfg.decimation; deci++) {
        if (lane < nd) {   // _compute_torques (legged_robot.py:691-713) with the action delay of legged_robot_fftai.py:58-61
            const float a = ((float)deci < A.delay ? A.last_actions : A.actions)[(size_t)e * nd + lane];
            float t = m.kp[lane] * (a * cfg.action_scale + m.q0[lane] - s.q[lane]) - m.kd[lane] * s.qd[lane];
            t *= A.motor[(size_t)e * nd + lane];
            const float lim = m.dof_effort[lane];
            s.tau[lane] = fminf(fmaxf(t, -lim), lim);
        }
        __syncwarp();
        g_substep(s, m, A, lane, mu_env, rest_env, A.sig ? A.sig + (size_t)e * cfg.decimation + deci : nullptr);
        g_kinematics(s, m, lane);
        if (lane < nf) {   // foot statistics of the substep just integrated (legged_robot_fftai.py:79-88)
            float ls[13];
            const int l = m.foot_link[lane];
            g_link_state(s, m, l, ls);
            const float *f = s.cf + 3 * l;
            s.facc[lane][0] += sqrtf(f[0] * f[0] + f[1] * f[1] + f[2] * f[2]);
#pragma unroll
            for (int k = 0; k < 3; k++) { s.facc[lane][1 + k] += fabsf(ls[7 + k]); s.facc[lane][4 + k] += fabsf(ls[10 + k]); }
        }
        __syncwarp();
    }
    // ---- outputs
    const float invd = 1.0f / (float)cfg.decimation;
    if (lane < 13) A.root[(size_t)e * 13 + lane] = s.root[lane];
    if (lane < nd) {
        A.q[(size_t)e * nd + lane] = s.q[lane]; A.qd[(size_t)e * nd + lane] = s.qd[lane];
        A.torques[(size_t)e * nd + lane] = s.tau[lane];
    }
    if (lane < nf) {
        A.avg_ff[(size_t)e * nf + lane] = s.facc[lane][0] * invd;
#pragma unroll
        for (int k = 0; k < 3; k++) {
            A.avg_fl[((size_t)e * nf + lane) * 3 + k] = s.facc[lane][1 + k] * invd;
            A.avg_fa[((size_t)e * nf + lane) * 3 + k] = s.facc[lane][4 + k] * invd;
        }
    }
    for (int l = lane; l < nl; l += 32) {
        float ls[13];
        g_link_state(s, m, l, ls);
        float *o = A.link_state + ((size_t)e * nl + l) * 13;
#pragma unroll
        for (int k = 0; k < 13; k++) o[k] = ls[k];
    }
    for (int i = lane; i < nl * 3; i += 32) A.contact_force[(size_t)e * nl * 3 + i] = s.cf[i];
}


// =========================================================================================================
// The whole policy step for any topology (the generic counterpart of env_step_kernel in grx_env.cu): clip_actions (legged_robot_fftai.py:171-177)
// -> decimation x [PD torque -> g_substep] with the foot averages (FF:51-88) -> the task half of grx_task.cuh.  One warp per robot.  The state
// record stays in global memory during the physics (only root / q / qd / the previous actions are needed) and is staged over the dead
// constraint Jacobian afterwards: J[0..127] measured heights, J[128..255] noise-free observation, J[256..] the record; Y = privileged row.
// =========================================================================================================
template <bool PHYS>
__global__ void __launch_bounds__(GWARPS * 32, 1) envg_step_kernel(const __grid_constant__ EnvArgs A, const __grid_constant__ grx_task_cfg cfg,
                                                                   const __grid_constant__ GArgs G, const __grid_constant__ LayR L) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    GWS &s = *reinterpret_cast<GWS *>(smem_raw + (size_t)warp * sizeof(GWS));
    const GModel &m = *G.m;
    const int e = blockIdx.x * (blockDim.x >> 5) + warp;
    if (blockIdx.x == 0 && threadIdx.x < ACC_W) A.episode_accum_next[threadIdx.x] = 0.f;   // nobody accumulates into the next slot during this launch
    if (e >= A.N) return;
    const int nd = L.nd;
    float *grec = A.rec + (size_t)e * L.rec_f;
    const float *gcst = A.cst + (size_t)e * L.cst_f;
    Draw draw;
    draw.U = A.U ? A.U + (size_t)e * L.rng_k : nullptr;
    draw.k0 = (uint32_t)cfg.seed; draw.k1 = (uint32_t)(cfg.seed >> 32);
    draw.gid = (uint32_t)(cfg.env_id_offset + e);
    draw.step_lo = (uint32_t)A.step_index; draw.step_hi = (uint32_t)(A.step_index >> 32);

    float act_l = 0.f, last_act_l = 0.f;
    if (lane < nd) {   // clip_actions
        act_l = fminf(fmaxf(A.actions[(size_t)e * nd + lane], cfg.clip_actions_min[lane]), cfg.clip_actions_max[lane]);
        last_act_l = grec[L.lastact + lane];
    }
    float ff_acc = 0.f, fl_acc[3] = {0, 0, 0}, foot_z = 0.f;
    float torso_q[4] = {0, 0, 0, 1};
    if (PHYS) {
        if (lane < 13) s.root[lane] = grec[L.root + lane];
        float motor_l = 1.f;
        if (lane < nd) { s.q[lane] = grec[L.dofpos + lane]; s.qd[lane] = grec[L.dofvel + lane]; motor_l = gcst[L.c_motor + lane]; }
        if (lane < 10) s.bin[lane] = gcst[L.c_bi + lane];
        __syncwarp();
        const float mu_env = gcst[L.c_fric], rest_env = gcst[L.c_rest];
        g_kinematics(s, m, lane);
        for (int deci = 0; deci < cfg.decimation; deci++) {
            if (lane < nd) {   // _compute_torques (legged_robot.py:691-713) with the action delay of FF:58-61
                const float a = ((float)deci < A.delay) ? last_act_l : act_l;
                float t = m.kp[lane] * (a * cfg.action_scale + m.q0[lane] - s.q[lane]) - m.kd[lane] * s.qd[lane];
                t *= motor_l;
                const float lim = m.dof_effort[lane];
                s.tau[lane] = fminf(fmaxf(t, -lim), lim);
            }
            __syncwarp();
            if (A.dbg_M != nullptr) {   // grx_env_debug_dynamics: mass matrix + bias of env dbg_index before the factorisation
                g_mass_and_bias(s, m, G.cfg.gravity, lane);
                const int nv = nd + 6;
                if (e == A.dbg_index) {
                    for (int i = lane; i < nv * nv; i += 32) { const int r = i / nv, c2 = i % nv; A.dbg_M[i] = s.M[r >= c2 ? tri(r, c2) : tri(c2, r)]; }
                    for (int i = lane; i < nv; i += 32) A.dbg_h[i] = s.h[i];
                }
                return;
            }
            g_substep(s, m, G, lane, mu_env, rest_env, A.dbg_sig ? A.dbg_sig + (size_t)e * A.dbg_sig_stride + deci : nullptr);
            g_kinematics(s, m, lane);
            if (lane < TK_NF) {   // foot statistics of the substep just integrated (FF:79-81)
                float ls[13];
                const int l = m.foot_link[lane];
                g_link_state(s, m, l, ls);
                const float *f = s.cf + 3 * l;
                ff_acc += sqrtf(f[0] * f[0] + f[1] * f[1] + f[2] * f[2]);
#pragma unroll
                for (int k = 0; k < 3; k++) fl_acc[k] += fabsf(ls[7 + k]);
                foot_z = ls[2];
            }
            __syncwarp();
        }
        const float invd = 1.0f / (float)cfg.decimation;
        ff_acc *= invd;
#pragma unroll
        for (int k = 0; k < 3; k++) fl_acc[k] *= invd;
        {
            float ls[13];
            g_link_state(s, m, m.torso_link, ls);
#pragma unroll
            for (int k = 0; k < 4; k++) torso_q[k] = ls[3 + k];
        }
        if (A.rigid_body_states != nullptr) {   // compat export (refresh_rigid_body_state_tensor, FF:75)
            for (int l = lane; l < m.nl; l += 32) {
                float ls[13];
                g_link_state(s, m, l, ls);
                float *o = A.rigid_body_states + ((size_t)e * m.nl + l) * 13;
#pragma unroll
                for (int k = 0; k < 13; k++) o[k] = ls[k];
            }
        }
        if (A.foot_state != nullptr && lane < TK_NF) {
            float ls[13];
            g_link_state(s, m, m.foot_link[lane], ls);
            float *o = A.foot_state + ((size_t)e * TK_NF + lane) * 13;
#pragma unroll
            for (int k = 0; k < 13; k++) o[k] = ls[k];
        }
        __syncwarp();
    } else {   // injected physics outputs (grx_env_post_physics)
        if (lane < nd) s.tau[lane] = A.inj.torques[(size_t)e * nd + lane];
        for (int i = lane; i < m.nl * 3; i += 32) s.cf[i] = A.inj.contact_forces[(size_t)e * m.nl * 3 + i];
        if (lane < TK_NF) {
            foot_z = A.inj.foot_state[((size_t)e * TK_NF + lane) * 13 + 2];
            ff_acc = A.inj.avg_foot_force[(size_t)e * TK_NF + lane];
#pragma unroll
            for (int k = 0; k < 3; k++) fl_acc[k] = A.inj.avg_foot_linvel[((size_t)e * TK_NF + lane) * 3 + k];
        }
#pragma unroll
        for (int k = 0; k < 4; k++) torso_q[k] = A.inj.torso_quat[(size_t)e * 4 + k];
    }
    float *scratch = &s.J[0][0];
    float *mh = scratch, *ob = scratch + 128, *rec = scratch + 256, *pri = &s.Y[0][0];
    for (int i = lane; i < L.rec_f; i += 32) rec[i] = grec[i];
    __syncwarp();
    if (PHYS) {
        if (lane < 13) rec[L.root + lane] = s.root[lane];
        if (lane < nd) { rec[L.dofpos + lane] = s.q[lane]; rec[L.dofvel + lane] = s.qd[lane]; }
        __syncwarp();
    }
    task_post_physics(rec, L, m, A, cfg, draw, lane, e, act_l, last_act_l, ff_acc, fl_acc, foot_z, torso_q, s.tau, s.cf, mh, ob, pri, s.Ad);
    __syncwarp();
    for (int i = lane; i < L.rec_f; i += 32) grec[i] = rec[i];
    float *gpri = A.pri_obs + (size_t)e * cfg.num_pri_obs;
    for (int i = lane; i < cfg.num_pri_obs; i += 32) gpri[i] = pri[i];
}
static_assert(GROWS * GPITCH >= 256 + 16 + 6 * GD + 48 + 8 && GROWS >= TK_NREW, "post-physics scratch fits over J / Ad");

// host-invoked reset_idx (legged_robot.py:377-440 as reached from BaseTask.reset()): one warp per listed env
__global__ void __launch_bounds__(256) envg_reset_kernel(const __grid_constant__ EnvArgs A, const __grid_constant__ grx_task_cfg cfg,
                                                         const __grid_constant__ GArgs G, const __grid_constant__ LayR L, const int *ids, int n,
                                                         int curriculum_active) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float *rec = reinterpret_cast<float *>(smem_raw) + (size_t)warp * L.rec_f;
    const GModel &m = *G.m;
    if (blockIdx.x == 0 && threadIdx.x < ACC_W) A.episode_accum_next[threadIdx.x] = 0.f;
    const int w = blockIdx.x * 8 + warp;
    if (w >= n) return;
    const int e = ids ? ids[w] : w;
    if (e < 0 || e >= A.N) return;
    float *g = A.rec + (size_t)e * L.rec_f;
    for (int i = lane; i < L.rec_f; i += 32) rec[i] = g[i];
    __syncwarp();
    Draw draw;
    draw.U = A.U ? A.U + (size_t)e * L.rng_k : nullptr;
    draw.k0 = (uint32_t)cfg.seed ^ 0x5bd1e995u; draw.k1 = (uint32_t)(cfg.seed >> 32);
    draw.gid = (uint32_t)(cfg.env_id_offset + e);
    draw.step_lo = (uint32_t)A.step_index; draw.step_hi = (uint32_t)(A.step_index >> 32);
    const float cx = rec[L.cmd], cy = rec[L.cmd + 1];
    reset_env(rec, L, m, A, cfg, draw, lane, sqrtf(cx * cx + cy * cy), curriculum_active != 0);
    if (lane == 0 && cfg.curriculum) atomicAdd(A.episode_accum + TK_NREW + 1, (float)__float_as_int(rec[L.tlevel]));
    __syncwarp();
    for (int i = lane; i < L.rec_f; i += 32) g[i] = rec[i];
}

}  // namespace

// =========================================================================================================
// Host side
// =========================================================================================================
// One CTA per SM is resident (shared memory): take the number of waves of the widest CTA and shrink the CTA until those waves are evenly filled
static int g_warps_per_cta(int N) {
    static int sms = 0;
    if (sms == 0) { int dev = 0; cudaGetDevice(&dev); if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148; }
    const int waves = (N + GWARPS * sms - 1) / (GWARPS * sms);
    int w = (N + waves * sms - 1) / (waves * sms);
    return w < 1 ? 1 : (w > GWARPS ? GWARPS : w);
}
static bool gmodel_fits(const grx_model_desc *md, int npairs) {
    return !(md->nb < 1 || md->nb > GB || md->nd != md->nb - 1 || md->nd > GD || md->nl > GL || md->ns > GS || md->nf > 4 || npairs > GP);
}
static bool build_gmodel(const grx_model_desc *md, const int32_t *self_pairs, int npairs, GModel &m, std::string &err) {
    memset(&m, 0, sizeof(m));
    m.nb = md->nb; m.nd = md->nd; m.nl = md->nl; m.ns = md->ns; m.nf = md->nf; m.npairs = npairs;
    for (int b = 0; b < md->nb; b++) {
        m.parent[b] = md->parent[b];
        if (b > 0 && (md->parent[b] < 0 || md->parent[b] >= b)) { err = "bodies must be in depth-first order (parent index < body index)"; return false; }
        memcpy(m.jpos[b], md->jpos + 3 * b, 12); memcpy(m.jrot[b], md->jrot + 9 * b, 36); memcpy(m.axis[b], md->axis + 3 * b, 12);
        m.mass[b] = md->mass[b]; memcpy(m.com[b], md->com + 3 * b, 12); memcpy(m.inertia[b], md->inertia + 6 * b, 24);
        // path base -> b and the ancestor bit mask
        int chain[GB], n = 0;
        for (int i = b; i >= 1; i = md->parent[i]) chain[n++] = i;
        if (n > GDEPTH) { err = "kinematic tree deeper than 12 joints"; return false; }
        m.depth[b] = n;
        if (n > m.maxdepth) m.maxdepth = n;
        for (int k = 0; k < n; k++) { m.path[b][k] = (signed char)chain[n - 1 - k]; m.anc[b] |= 1ull << (chain[k] - 1); }
    }
    for (int j = 0; j < md->nd; j++) {
        m.dof_lower[j] = md->dof_lower[j]; m.dof_upper[j] = md->dof_upper[j]; m.dof_vel_limit[j] = md->dof_vel_limit[j]; m.dof_effort[j] = md->dof_effort[j];
        m.kp[j] = md->kp[j]; m.kd[j] = md->kd[j]; m.q0[j] = md->default_pos[j];
    }
    for (int l = 0; l < md->nl; l++) { m.link_body[l] = md->link_body[l]; memcpy(m.link_pos[l], md->link_pos + 3 * l, 12); memcpy(m.link_rot[l], md->link_rot + 9 * l, 36); }
    for (int s = 0; s < md->ns; s++) { m.sph_body[s] = md->sph_body[s]; m.sph_link[s] = md->sph_link[s]; memcpy(m.sph_pos[s], md->sph_pos + 3 * s, 12); m.sph_rad[s] = md->sph_rad[s]; }
    for (int f = 0; f < md->nf; f++) m.foot_link[f] = md->foot_links[f];
    for (int i = 0; i < npairs; i++) {
        m.pair_a[i] = self_pairs[2 * i]; m.pair_b[i] = self_pairs[2 * i + 1];
        if (m.pair_a[i] < 0 || m.pair_a[i] >= md->ns || m.pair_b[i] < 0 || m.pair_b[i] >= md->ns) { err = "self-collision pair out of range"; return false; }
    }
    // task tables
    for (int j = 0; j < md->nd; j++) { m.soft_lower[j] = md->soft_lower ? md->soft_lower[j] : md->dof_lower[j]; m.soft_upper[j] = md->soft_upper ? md->soft_upper[j] : md->dof_upper[j]; }
    m.term_mask = 0;
    for (int k = 0; k < md->nterm; k++) m.term_mask |= 1ull << md->term_links[k];
    m.nankle = md->nankle;
    for (int k = 0; k < md->nankle && k < 4; k++) m.ankle_dof[k] = md->ankle_dofs[k];
    m.torso_link = md->torso_link;
    return true;
}

struct grx_physg {
    int N = 0, device = 0;
    GModel hm;
    GModel *dm = nullptr;
    grx_physg_cfg cfg;
    TerrainDev terrain;
    short *heights = nullptr;
};

extern "C" int grx_physg_create(const grx_model_desc *md, const int32_t *self_pairs, int32_t npairs, const grx_physg_cfg *cfg, int32_t num_envs,
                                int32_t device, grx_physg **out) {
    if (!md || !cfg || !out || num_envs <= 0 || (npairs > 0 && !self_pairs)) return grx_set_error(GRX_E_INVALID, "grx_physg_create: bad arguments");
    if (md->nb < 1 || md->nb > GB || md->nd != md->nb - 1 || md->nd > GD || md->nl > GL || md->ns > GS || md->nf > 4 || npairs > GP)
        return grx_set_error(GRX_E_INVALID, "grx_physg_create: model exceeds the generic kernel's limits (<= 36 bodies, 32 DOF, 48 links, 32 contact spheres, 64 self-collision pairs)");
    if (cfg->decimation < 1 || cfg->solver_iters < 1) return grx_set_error(GRX_E_INVALID, "grx_physg_create: decimation / solver_iters must be >= 1");
    CK(cudaSetDevice(device));
    grx_physg *p = new grx_physg();
    p->N = num_envs; p->device = device; p->cfg = *cfg;
    {
        std::string err;
        if (!build_gmodel(md, self_pairs, npairs, p->hm, err)) { delete p; return grx_set_error(GRX_E_INVALID, "grx_physg_create: " + err); }
    }
    CK(cudaMalloc((void **)&p->dm, sizeof(GModel)));
    CK(cudaMemcpy(p->dm, &p->hm, sizeof(GModel), cudaMemcpyHostToDevice));
    memset(&p->terrain, 0, sizeof(p->terrain));
    p->terrain.hscale = 1.f; p->terrain.vscale = 1.f; p->terrain.friction = 1.f;
    CK(cudaFuncSetAttribute(physg_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(GWARPS * sizeof(GWS))));
    *out = p;
    return GRX_OK;
}

extern "C" int grx_physg_destroy(grx_physg *p) {
    if (!p) return GRX_OK;
    cudaSetDevice(p->device);
    if (p->dm) cudaFree(p->dm);
    if (p->heights) cudaFree(p->heights);
    delete p;
    return GRX_OK;
}

extern "C" int grx_physg_set_terrain_plane(grx_physg *p, float friction, float restitution) {
    if (!p) return grx_set_error(GRX_E_INVALID, "null physg");
    p->terrain.type = 0; p->terrain.friction = friction; p->terrain.restitution = restitution;
    return GRX_OK;
}

extern "C" int grx_physg_set_terrain_heightfield(grx_physg *p, const int16_t *samples, int32_t rows, int32_t cols, float hscale, float vscale,
                                                 float border, float friction, float restitution) {
    if (!p || !samples || rows < 2 || cols < 2) return grx_set_error(GRX_E_INVALID, "grx_physg_set_terrain_heightfield: bad arguments");
    CK(cudaSetDevice(p->device));
    if (p->heights) { cudaFree(p->heights); p->heights = nullptr; }
    CK(cudaMalloc((void **)&p->heights, (size_t)rows * cols * 2));
    CK(cudaMemcpy(p->heights, samples, (size_t)rows * cols * 2, cudaMemcpyHostToDevice));
    p->terrain.type = 1; p->terrain.rows = rows; p->terrain.cols = cols; p->terrain.h = p->heights; p->terrain.mv = nullptr; p->terrain.near_mv = nullptr;
    p->terrain.hscale = hscale; p->terrain.vscale = vscale; p->terrain.border = border; p->terrain.friction = friction; p->terrain.restitution = restitution;
    return GRX_OK;
}

extern "C" int grx_physg_step(grx_physg *p, float *d_root, float *d_dof_pos, float *d_dof_vel, const float *d_actions, const float *d_last_actions,
                              float delay, const float *d_motor_strength, const float *d_base_inertial, const float *d_friction,
                              const float *d_restitution, float *d_torques, float *d_link_state, float *d_contact_force, float *d_avg_foot_force,
                              float *d_avg_foot_linvel, float *d_avg_foot_angvel, uint64_t *d_active_sig, void *stream) {
    if (!p || !d_root || !d_dof_pos || !d_dof_vel || !d_actions || !d_last_actions || !d_motor_strength || !d_base_inertial || !d_friction ||
        !d_restitution || !d_torques || !d_link_state || !d_contact_force || !d_avg_foot_force || !d_avg_foot_linvel || !d_avg_foot_angvel)
        return grx_set_error(GRX_E_INVALID, "grx_physg_step: null argument");
    GArgs A;
    memset(&A, 0, sizeof(A));
    A.m = p->dm; A.terrain = p->terrain; A.cfg = p->cfg; A.N = p->N;
    A.root = d_root; A.q = d_dof_pos; A.qd = d_dof_vel; A.actions = d_actions; A.last_actions = d_last_actions; A.delay = delay;
    A.motor = d_motor_strength; A.binert = d_base_inertial; A.friction = d_friction; A.restitution = d_restitution;
    A.torques = d_torques; A.link_state = d_link_state; A.contact_force = d_contact_force; A.avg_ff = d_avg_foot_force;
    A.avg_fl = d_avg_foot_linvel; A.avg_fa = d_avg_foot_angvel; A.sig = reinterpret_cast<unsigned long long *>(d_active_sig);
    grx_count_launch();
    const int wpc = g_warps_per_cta(p->N);
    physg_step_kernel<<<(p->N + wpc - 1) / wpc, wpc * 32, wpc * sizeof(GWS), (cudaStream_t)stream>>>(A);
    CK(cudaGetLastError());
    return GRX_OK;
}

// =========================================================================================================
// Generic-topology env (grx_envg.h): what grx_env_* dispatches to for every model that is not the registered lower-limb tree
// =========================================================================================================
namespace {
struct EnvG {
    int device = 0;
    GModel hm;
    GModel *dm = nullptr;
    grx_physg_cfg pcfg;
};
GArgs envg_args(const EnvG *g, const EnvArgs &A, const grx_task_cfg &cfg) {
    GArgs G;
    memset(&G, 0, sizeof(G));
    G.m = g->dm; G.terrain = A.terrain; G.cfg = g->pcfg; G.N = A.N;
    G.cfg.sim_dt = cfg.sim_dt; G.cfg.gravity = cfg.gravity; G.cfg.contact_offset = cfg.contact_offset; G.cfg.bounce_threshold = cfg.bounce_threshold;
    G.cfg.max_depen_vel = cfg.max_depen_vel; G.cfg.erp = cfg.erp; G.cfg.solver_iters = cfg.solver_iters; G.cfg.decimation = cfg.decimation;
    G.cfg.action_scale = cfg.action_scale;
    return G;
}
}  // namespace

int grx::envg_create(const grx_model_desc *md, const grx_task_cfg *cfg, int device, void **out) {
    if (!gmodel_fits(md, 0) || md->nf != TK_NF || md->nl > GL)
        return grx_set_error(GRX_E_INVALID, "grx_env_create: model exceeds the generic kernel's limits (<= 36 bodies in depth-first order, 32 DOF, 48 links, 32 contact spheres, 2 feet)");
    if (md->torso_link < 0 || md->torso_link >= md->nl) return grx_set_error(GRX_E_INVALID, "grx_env_create: torso_link out of range");
    if (md->nankle != 2 && md->nankle != 4) return grx_set_error(GRX_E_INVALID, "grx_env_create: need 2 or 4 ankle DOF (left half, right half)");
    CK(cudaSetDevice(device));
    EnvG *g = new EnvG();
    g->device = device;
    std::string err;
    if (!build_gmodel(md, nullptr, 0, g->hm, err)) { delete g; return grx_set_error(GRX_E_INVALID, "grx_env_create: " + err); }
    memset(&g->pcfg, 0, sizeof(g->pcfg));
    g->pcfg.max_contacts = GKC; g->pcfg.max_self_contacts = 0;
    CK(cudaMalloc((void **)&g->dm, sizeof(GModel)));
    CK(cudaMemcpy(g->dm, &g->hm, sizeof(GModel), cudaMemcpyHostToDevice));
    CK(cudaFuncSetAttribute(envg_step_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(GWARPS * sizeof(GWS))));
    CK(cudaFuncSetAttribute(envg_step_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(GWARPS * sizeof(GWS))));
    (void)cfg;
    *out = g;
    return GRX_OK;
}

void grx::envg_destroy(void *h) {
    EnvG *g = static_cast<EnvG *>(h);
    if (!g) return;
    cudaSetDevice(g->device);
    if (g->dm) cudaFree(g->dm);
    delete g;
}

int grx::envg_set_self_collision(void *h, const int32_t *pairs, int32_t npairs, int32_t max_self_contacts) {
    EnvG *g = static_cast<EnvG *>(h);
    if (!g || npairs < 0 || npairs > GP || (npairs > 0 && !pairs) || max_self_contacts < 0 || max_self_contacts > GKS)
        return grx_set_error(GRX_E_INVALID, "grx_env_set_self_collision: bad arguments (<= 64 pairs, <= 4 self-contacts per substep)");
    for (int i = 0; i < npairs; i++)
        if (pairs[2 * i] < 0 || pairs[2 * i] >= g->hm.ns || pairs[2 * i + 1] < 0 || pairs[2 * i + 1] >= g->hm.ns)
            return grx_set_error(GRX_E_INVALID, "grx_env_set_self_collision: pair index out of range");
    CK(cudaSetDevice(g->device));
    CK(cudaDeviceSynchronize());
    for (int i = 0; i < npairs; i++) { g->hm.pair_a[i] = pairs[2 * i]; g->hm.pair_b[i] = pairs[2 * i + 1]; }
    g->hm.npairs = npairs;
    g->pcfg.max_self_contacts = npairs > 0 ? max_self_contacts : 0;
    CK(cudaMemcpy(g->dm, &g->hm, sizeof(GModel), cudaMemcpyHostToDevice));
    return GRX_OK;
}

int grx::envg_launch_step(void *h, const EnvArgs &A, const grx_task_cfg &cfg, const LayR &L, bool phys, cudaStream_t st) {
    EnvG *g = static_cast<EnvG *>(h);
    const GArgs G = envg_args(g, A, cfg);
    const int wpc = g_warps_per_cta(A.N), grid = (A.N + wpc - 1) / wpc;
    grx_count_launch();
    if (phys) envg_step_kernel<true><<<grid, wpc * 32, wpc * sizeof(GWS), st>>>(A, cfg, G, L);
    else envg_step_kernel<false><<<grid, wpc * 32, wpc * sizeof(GWS), st>>>(A, cfg, G, L);
    CK(cudaGetLastError());
    return GRX_OK;
}

int grx::envg_launch_reset(void *h, const EnvArgs &A, const grx_task_cfg &cfg, const LayR &L, const int *ids, int n, int curriculum_active, cudaStream_t st) {
    EnvG *g = static_cast<EnvG *>(h);
    const GArgs G = envg_args(g, A, cfg);
    grx_count_launch();
    envg_reset_kernel<<<(n + 7) / 8, 256, 8 * (size_t)L.rec_f * sizeof(float), st>>>(A, cfg, G, L, ids, n, curriculum_active);
    CK(cudaGetLastError());
    return GRX_OK;
}
