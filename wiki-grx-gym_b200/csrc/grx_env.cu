// grx_env.cu — the fused GRx env-step kernel for sm_100a and its C ABI (include/grx_b200.h).
//
// One warp per robot.  One launch = one policy step of the reference's LeggedRobot.step() with the GR1T1 MRO
// (legged_robot.py:222-246, legged_robot_fftai.py:46-133, gr1t1.py:281-589):
//   action clip -> `decimation` x [PD torque -> articulated-body dynamics -> contact/limit solve -> integrate]
//   -> state update, height sampling, termination, 24 reward terms, reset + curriculum, observations.
// The per-env state record (432 B) and parameter record (96 B) are staged into shared memory with TMA bulk copies
// (cp.async.bulk + mbarrier) and the updated record / privileged-observation row go back the same way; the small
// outputs are written with coalesced lane-strided stores.  Dynamics = spec "GRX-dyn v1" (oracle/phys_impl.h is the
// fp64/fp32 CPU statement of the same equations): Kane's equations with composite-rigid-body mass matrix about the
// base origin, branch-sparse Cholesky, one lane per constraint row (<= 8 contacts x 3 + <= 8 joint limits = 32 rows),
// Delassus-space projected Gauss-Seidel.
//
// Topology handled by this kernel: floating base + 2 serial chains of 5 revolute joints (the registered lower-limb
// GR1T1 / GR1T2 tasks); grx_env_create rejects anything else.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "grx_b200.h"
#include "grx_count.h"
#include "grx_terrain.cuh"
#include "grx_task.cuh"
#include "grx_envg.h"

std::atomic<unsigned long long> g_grx_launches{0};
extern "C" uint64_t grx_debug_launch_count(void) { return (uint64_t)g_grx_launches.load(); }

namespace {

constexpr int NB = 11, ND = 10, NV = 16, CH = 5, NLMAX = 40, NSMAX = 32, NF = 2, KC = 8, KLIM = 7;
constexpr int YS = 16;   // row stride of WS::Y (floats); the four float4 slots of a row are XOR-swizzled with (row >> 1) & 3 (conflict-free per-lane float4 stores)
constexpr int ASP = 31;  // pitch of the Delassus block: 31 rows x 31 columns (at most 31 constraint rows; lane 31 solves the unconstrained update)
constexpr int NREW = 24, NHMAX = 128;
constexpr int WARPS_PER_CTA = 16;   // warps per CTA of the 128-register build and of the reset kernel (one CTA per SM; its warps re-converge at every substep so
                                    // the warps of a scheduler share instruction-cache lines).  The launch picks warps_per_cta so that the CTAs fill whole waves
constexpr int WARPS_PER_CTA_WIDE = 28;   // the 72-register build: 4096 robots on 148 SMs = ONE wave of 28 warps per SM (needs sizeof(WS) <= 8.1 KB)
constexpr int SIG_STRIDE = 16;               // substep slots per env in the active-set signature export (decimation <= 16)
constexpr unsigned FULL = 0xffffffffu;

// ---- per-env state record (floats; ints stored bit-wise) — 108 floats = 432 B, 16-B aligned
enum { R_ROOT = 0, R_DOFPOS = 16, R_DOFVEL = 26, R_LASTDOFVEL = 36, R_LASTACT = 46, R_LASTLASTACT = 56, R_CMD = 66,
       R_BHO = 69, R_AIR = 70, R_LAND = 72, R_CLAST = 74, R_EPLEN = 76, R_TLEVEL = 77, R_ORIGIN = 78, R_TTYPE = 81,
       R_SUMS = 82, REC_F = 108 };
// ---- per-env parameter record — 24 floats = 96 B
enum { C_MOTOR = 0, C_BI = 10, C_FRIC = 20, C_REST = 21, CST_F = 24 };
// ---- uniform-draw slots (grx_b200/rng_layout.py)
enum { U_NOISE = 0, U_RESET_DOF = 39, U_RESET_XY = 49, U_RESET_YAW = 51, U_RESET_VEL = 52, U_CMD_TIME = 58,
       U_CMD_RESET = 61, U_PUSH = 64, U_CURRICULUM = 66 };
static_assert(GRX_RNG_K == 68, "rng layout");
using Lay10 = LayC<ND>;   // grx_task.cuh: the same offsets as the enums above, as a layout policy of the shared task code
static_assert(Lay10::dofpos == R_DOFPOS && Lay10::lastlastact == R_LASTLASTACT && Lay10::cmd == R_CMD && Lay10::bho == R_BHO && Lay10::clast == R_CLAST &&
              Lay10::eplen == R_EPLEN && Lay10::ttype == R_TTYPE && Lay10::sums == R_SUMS && Lay10::rec_f == REC_F, "record layout");
static_assert(Lay10::c_bi == C_BI && Lay10::c_fric == C_FRIC && Lay10::c_rest == C_REST && Lay10::cst_f == CST_F, "parameter record layout");
static_assert(Lay10::u_reset_dof == U_RESET_DOF && Lay10::u_cmd_time == U_CMD_TIME && Lay10::u_push == U_PUSH && Lay10::u_curriculum == U_CURRICULUM &&
              Lay10::rng_k == GRX_RNG_K, "uniform-draw layout");

struct ModelDev {
    float jpos[NB][3], jrot[NB][9], axis[NB][3], mass[NB], com[NB][3], inertia[NB][6];
    float dof_lower[ND], dof_upper[ND], dof_vel_limit[ND], dof_effort[ND], soft_lower[ND], soft_upper[ND];
    float kp[ND], kd[ND], q0[ND];
    int ns, nl;
    int sph_body[NSMAX], sph_link[NSMAX];
    float sph_pos[NSMAX][3], sph_rad[NSMAX];
    int foot_link[NF], foot_body[NF];
    float foot_pos[NF][3];
    float torso_rot[9];
    int torso_body;
    unsigned long long term_mask;
    int ankle_dof[2];
    __device__ __forceinline__ float ankle_torque(const float *tau, int side) const { return fabsf(tau[ankle_dof[side]]); }   // one ankle DOF per leg
    int jrot_nonident;   // bit b set: joint frame rotation of body b is not the identity (GR1T1 / GR1T2: none) -> one 3x3 product per joint saved
};




// per-warp shared-memory workspace
struct alignas(16) WS {
    float rec[REC_F];
    float cst[CST_F];
    // Lifetimes inside a substep: kinematics (R .. ao) and the mass-matrix scratch (bi .. M) are dead once the constraint Jacobians have been
    // built and M^-1 applied; the Delassus block As is written after that and read only by the Gauss-Seidel sweep -> they share storage.
    union {
        struct {
            float R[NB][9], o[NB][3], a[NB][3], c[NB][3], Iw[NB][6], w[NB][3], vo[NB][3], al[NB][3], ao[NB][3];
            float bi[NB][10], bw[NB][6];
            float S[ND][6], F[ND][6];
            float M[NV][NV + 1];
        };
        float As[ASP * ASP + 3];   // Delassus matrix, As[r * ASP + c] = J_c . Y_r (symmetric); after the physics: measured heights + obs staging
    };
    float invd[NV], h[NV], u[NV], tau[NV];
    float Y[32][YS];     // M^-1 J^T, one row per constraint (swizzled, see y_slot); after the physics: privileged-observation row staging (bulk-stored)
    float rowc[32][2];   // 1 / A_rr, bias
    float cfr[KC][9];    // contact frame n, t1, t2
    float cpt[KC][4];    // contact point xyz, target velocity
    int cbody[KC], clink[KC];
    int limj[8];
    float lims[8], limt[8];
    float cf[NLMAX * 3];
    float rterm[NREW];
    unsigned long long mbar;
};
static_assert(sizeof(float) * 32 * YS >= 168 * 4, "pri_obs staging aliases Y");
static_assert(sizeof(float) * (NB * (9 + 7 * 3 + 6) + NB * 16 + ND * 12 + NV * (NV + 1)) >= sizeof(float) * ASP * ASP, "the Delassus block fits over the dead dynamics scratch");
static_assert(ASP * ASP >= NHMAX + 64, "heights + obs staging alias the Delassus block");
static_assert(sizeof(WS) <= 8128, "28 warps per CTA need the per-warp workspace to stay near 8 KB");
__device__ __forceinline__ int y_slot(int row, int i4) { return (i4 ^ ((row >> 1) & 3)) << 2; }   // float offset of float4 slot i4 of row `row` in WS::Y

__device__ __forceinline__ void cross3(const float *a, const float *b, float *o) {
    float x = a[1] * b[2] - a[2] * b[1], y = a[2] * b[0] - a[0] * b[2], z = a[0] * b[1] - a[1] * b[0];
    o[0] = x; o[1] = y; o[2] = z;
}
__device__ __forceinline__ float dot3(const float *a, const float *b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
__device__ __forceinline__ void m3v(const float *R, const float *v, float *o) {
    float x = R[0] * v[0] + R[1] * v[1] + R[2] * v[2];
    float y = R[3] * v[0] + R[4] * v[1] + R[5] * v[2];
    float z = R[6] * v[0] + R[7] * v[1] + R[8] * v[2];
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
    __sincosf(th, &s, &c);   // |joint angle| < pi: abs error < 5e-7; the libm slow path alone is ~1000 instructions
    float t = 1 - c;
    R[0] = c + a[0] * a[0] * t;        R[1] = a[0] * a[1] * t - a[2] * s; R[2] = a[0] * a[2] * t + a[1] * s;
    R[3] = a[1] * a[0] * t + a[2] * s; R[4] = c + a[1] * a[1] * t;        R[5] = a[1] * a[2] * t - a[0] * s;
    R[6] = a[2] * a[0] * t - a[1] * s; R[7] = a[2] * a[1] * t + a[0] * s; R[8] = c + a[2] * a[2] * t;
}
__device__ __forceinline__ void sym6v(const float *S, const float *v, float *o) {
    float x = S[0] * v[0] + S[3] * v[1] + S[4] * v[2];
    float y = S[3] * v[0] + S[1] * v[1] + S[5] * v[2];
    float z = S[4] * v[0] + S[5] * v[1] + S[2] * v[2];
    o[0] = x; o[1] = y; o[2] = z;
}


// ---- TMA bulk copies (1-D) + mbarrier
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, uint32_t parity) {
    uint32_t ok = 0;
    while (!ok) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    }
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, unsigned long long *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_s2g(void *dst, const void *src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(src)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit_wait() {
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }



// ---- terrain query (oracle/phys_impl.h terrain_query)
// ---- forward kinematics + body velocities + velocity-product accelerations; lane b < NB owns body b and walks its own chain
template <int MAXW>   // (a separate copy per kernel build: its register allocation follows that build's launch bounds)
__device__ __noinline__ void kinematics(WS &s, const ModelDev &m, int lane) {
    if (lane < NB) {
        const float *rt = s.rec + R_ROOT;
        float R[9], o[3], w[3], vo[3], al[3] = {0, 0, 0}, ao[3] = {0, 0, 0}, a[3] = {0, 0, 0};
        quat2mat(rt + 3, R);
#pragma unroll
        for (int k = 0; k < 3; k++) { o[k] = rt[k]; vo[k] = rt[7 + k]; w[k] = rt[10 + k]; }
        const int leg = (lane - 1) / CH, depth = lane == 0 ? 0 : (lane - 1) % CH + 1;
#pragma unroll 1
        for (int k = 0; k < depth; k++) {
            const int j = 1 + leg * CH + k;
            const float q = s.rec[R_DOFPOS + j - 1], qd = s.rec[R_DOFVEL + j - 1];
            float Rj[9], Rq[9], r[3], an[3], t1[3], t2[3], t3[3], t4[3];
            if ((m.jrot_nonident >> j) & 1) m3m(R, m.jrot[j], Rj);
            else {
#pragma unroll
                for (int i = 0; i < 9; i++) Rj[i] = R[i];
            }
            axang2mat(m.axis[j], q, Rq);
            m3v(R, m.jpos[j], r);
            m3m(Rj, Rq, R);
            m3v(R, m.axis[j], an);
            cross3(w, r, t1);
            cross3(al, r, t2);
            cross3(w, t1, t3);
            cross3(w, an, t4);
#pragma unroll
            for (int i = 0; i < 3; i++) {
                o[i] += r[i]; vo[i] += t1[i]; ao[i] += t2[i] + t3[i];
                al[i] += t4[i] * qd; w[i] += an[i] * qd; a[i] = an[i];
            }
        }
        const float *com = lane == 0 ? s.cst + C_BI + 1 : m.com[lane];
        const float *I6 = lane == 0 ? s.cst + C_BI + 4 : m.inertia[lane];
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
        for (int i = 0; i < 9; i++) s.R[lane][i] = R[i];
#pragma unroll
        for (int i = 0; i < 3; i++) {
            s.o[lane][i] = o[i]; s.a[lane][i] = a[i]; s.c[lane][i] = o[i] + rc[i];
            s.w[lane][i] = w[i]; s.vo[lane][i] = vo[i]; s.al[lane][i] = al[i]; s.ao[lane][i] = ao[i];
        }
        s.Iw[lane][0] = T[0]; s.Iw[lane][1] = T[4]; s.Iw[lane][2] = T[8];
        s.Iw[lane][3] = T[1]; s.Iw[lane][4] = T[2]; s.Iw[lane][5] = T[5];
    }
    __syncwarp();
}

// ---- mass matrix (internal order: joints 0..9, base linear 10..12, base angular 13..15) + bias vector
__device__ __forceinline__ void mass_and_bias(WS &s, const ModelDev &m, float gravity, int lane) {
    for (int i = lane; i < NV * (NV + 1); i += 32) (&s.M[0][0])[i] = 0.0f;
    if (lane < NB) {
        const int b = lane;
        const float mass = b == 0 ? s.cst[C_BI] : m.mass[b];
        float r[3], rc[3];
#pragma unroll
        for (int k = 0; k < 3; k++) { r[k] = s.c[b][k] - s.o[0][k]; rc[k] = s.c[b][k] - s.o[b][k]; }
        const float rr = dot3(r, r);
        float *bi = s.bi[b];
        bi[0] = mass; bi[1] = mass * r[0]; bi[2] = mass * r[1]; bi[3] = mass * r[2];
        bi[4] = s.Iw[b][0] + mass * (rr - r[0] * r[0]); bi[5] = s.Iw[b][1] + mass * (rr - r[1] * r[1]);
        bi[6] = s.Iw[b][2] + mass * (rr - r[2] * r[2]);
        bi[7] = s.Iw[b][3] - mass * r[0] * r[1]; bi[8] = s.Iw[b][4] - mass * r[0] * r[2]; bi[9] = s.Iw[b][5] - mass * r[1] * r[2];
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
        for (int k = 0; k < 3; k++) { s.bw[b][k] = f[k]; s.bw[b][3 + k] = Ial[k] + g3[k] + t1[k]; }
    }
    __syncwarp();
    if (lane <= ND) {
        // lanes 0..9: composite of the chain suffix starting at joint body lane+1; lane 10: whole robot = base body + the two
        // leg composites (lanes 0 and CH), fetched by shuffles so that the summation loop runs CH trips instead of NB
        int first, last;
        if (lane < ND) { first = lane + 1; last = 1 + (lane / CH) * CH + CH - 1; }
        else { first = 0; last = 0; }
        float ci[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0}, cw[6] = {0, 0, 0, 0, 0, 0};
        for (int b = first; b <= last; b++) {
#pragma unroll
            for (int k = 0; k < 10; k++) ci[k] += s.bi[b][k];
#pragma unroll
            for (int k = 0; k < 6; k++) cw[k] += s.bw[b][k];
        }
        {
            const unsigned msk = (1u << (ND + 1)) - 1u;   // the lanes inside this branch
#pragma unroll
            for (int k = 0; k < 10; k++) {
                const float l1 = __shfl_sync(msk, ci[k], 0), l2 = __shfl_sync(msk, ci[k], CH);
                if (lane == ND) ci[k] += l1 + l2;
            }
#pragma unroll
            for (int k = 0; k < 6; k++) {
                const float l1 = __shfl_sync(msk, cw[k], 0), l2 = __shfl_sync(msk, cw[k], CH);
                if (lane == ND) cw[k] += l1 + l2;
            }
        }
        if (lane < ND) {
            const int j = lane + 1;
            float d[3], Sl[3], Sa[3], t1[3], t2[3], t3[3];
#pragma unroll
            for (int k = 0; k < 3; k++) { d[k] = s.o[0][k] - s.o[j][k]; Sa[k] = s.a[j][k]; }
            cross3(Sa, d, Sl);
            cross3(Sa, ci + 1, t1);
            cross3(ci + 1, Sl, t2);
            sym6v(ci + 4, Sa, t3);
#pragma unroll
            for (int k = 0; k < 3; k++) {
                s.S[lane][k] = Sl[k]; s.S[lane][3 + k] = Sa[k];
                s.F[lane][k] = ci[0] * Sl[k] + t1[k]; s.F[lane][3 + k] = t2[k] + t3[k];
            }
            s.h[lane] = dot3(Sl, cw) + dot3(Sa, cw + 3);
        } else {
            const float mm = ci[0], *hh = ci + 1, *I = ci + 4;
            const int L = ND, A = ND + 3;
            const float hx[9] = {0, -hh[2], hh[1], hh[2], 0, -hh[0], -hh[1], hh[0], 0};
#pragma unroll
            for (int k = 0; k < 3; k++) s.M[L + k][L + k] = mm;
#pragma unroll
            for (int i = 0; i < 3; i++)
#pragma unroll
                for (int j = 0; j < 3; j++) { s.M[A + i][L + j] = hx[3 * i + j]; s.M[L + j][A + i] = hx[3 * i + j]; }
            s.M[A][A] = I[0]; s.M[A + 1][A + 1] = I[1]; s.M[A + 2][A + 2] = I[2];
            s.M[A][A + 1] = s.M[A + 1][A] = I[3];
            s.M[A][A + 2] = s.M[A + 2][A] = I[4];
            s.M[A + 1][A + 2] = s.M[A + 2][A + 1] = I[5];
#pragma unroll
            for (int k = 0; k < 6; k++) s.h[L + k] = cw[k];
        }
    }
    __syncwarp();
    if (lane < 30) {  // joint-joint entries: 15 (ancestor, descendant) pairs per leg
        const int leg = lane / 15, idx = lane % 15;
        int kj = 0, acc = 0;
#pragma unroll
        for (int t = 0; t < CH; t++) if (idx >= acc + t + 1) { acc += t + 1; kj = t + 1; }
        const int ki = idx - acc;
        const int i = leg * CH + ki, j = leg * CH + kj;
        const float v = dot3(s.S[i], s.F[j]) + dot3(s.S[i] + 3, s.F[j] + 3);
        s.M[i][j] = v; s.M[j][i] = v;
    }
    for (int e = lane; e < 6 * ND; e += 32) {  // base-joint coupling: M[10+k][j] = F_j[k]
        const int j = e % ND, k = e / ND;
        const float v = s.F[j][k];
        s.M[ND + k][j] = v; s.M[j][ND + k] = v;
    }
    __syncwarp();
}

// ---- Cholesky of M in place (lower factor), right-looking, in REGISTERS: lane (i = lane & 15) holds row i of M; column k of L
// travels by warp shuffles (pivot from lane k, L[j][k] from lane j), so the factorisation is ~120 shuffle + FMA pairs with no
// shared-memory round trips or warp barriers inside.  The zero block between the two legs stays exactly zero (no fill): those
// updates are skipped here and chol_solve exploits the same structure.  Both half-warps compute the same rows (mirror).
template <int MAXW>
__device__ __noinline__ void cholesky(WS &s, int lane) {
    const int i = lane & 15;
    float row[NV];
#pragma unroll
    for (int j = 0; j < NV; j++) row[j] = s.M[i][j];
    float myinv = 0.f;
#pragma unroll
    for (int k = 0; k < NV; k++) {
        const float inv = rsqrtf(__shfl_sync(FULL, row[k], k));   // pivot (already updated by steps < k) from lane k
        const float lik = row[k] * inv;                            // L[i][k] for i >= k (lane k: sqrt of the pivot)
        row[k] = lik;
        if (i == k) myinv = inv;
#pragma unroll
        for (int j = k + 1; j < NV; j++) {
            if (k < CH && j >= CH && j < ND) continue;             // leg-1 column x leg-2 row: structurally zero
            row[j] -= lik * __shfl_sync(FULL, lik, j);             // L[j][k] from lane j; meaningful for i >= j
        }
    }
    if (lane < NV) {
#pragma unroll
        for (int j = 0; j < NV; j++) if (j <= i) s.M[i][j] = row[j];
        s.invd[i] = myinv;
    }
    __syncwarp();
}

// x <- M^-1 x with the factor in s.M (every lane solves its own right-hand side; reads of L are warp-broadcasts)
__device__ __forceinline__ void chol_solve(const WS &s, float *x) {
#pragma unroll
    for (int leg = 0; leg < 2; leg++) {
#pragma unroll
        for (int i = 0; i < CH; i++) {
            float t = x[leg * CH + i];
#pragma unroll
            for (int p = 0; p < i; p++) t -= s.M[leg * CH + i][leg * CH + p] * x[leg * CH + p];
            x[leg * CH + i] = t * s.invd[leg * CH + i];
        }
    }
#pragma unroll
    for (int i = ND; i < NV; i++) {
        float t = x[i];
#pragma unroll
        for (int p = 0; p < i; p++) t -= s.M[i][p] * x[p];
        x[i] = t * s.invd[i];
    }
#pragma unroll
    for (int i = NV - 1; i >= ND; i--) {
        float t = x[i];
#pragma unroll
        for (int p = i + 1; p < NV; p++) t -= s.M[p][i] * x[p];
        x[i] = t * s.invd[i];
    }
#pragma unroll
    for (int leg = 0; leg < 2; leg++) {
#pragma unroll
        for (int i = CH - 1; i >= 0; i--) {
            float t = x[leg * CH + i];
#pragma unroll
            for (int p = i + 1; p < CH; p++) t -= s.M[leg * CH + p][leg * CH + i] * x[leg * CH + p];
#pragma unroll
            for (int p = ND; p < NV; p++) t -= s.M[p][leg * CH + i] * x[p];
            x[leg * CH + i] = t * s.invd[leg * CH + i];
        }
    }
}

// ---- one dt: contacts + limits + solve + integrate.  Needs kinematics() of the current state in s.
// Constraint rows: 3 per contact (<= KC contacts) then <= KLIM joint-limit rows, one lane per row (<= 31); lane 31 solves the
// unconstrained update.  The projected Gauss-Seidel sweep runs in constraint space on the Delassus matrix A = J M^-1 J^T
// (same iterates as the velocity-space sweep of oracle/phys_impl.h): all loops are rolled to keep the instruction footprint small.
__device__ __forceinline__ void cta_align(int nthreads) {   // phase alignment of the CTA's warps (instruction-cache locality); no data is shared
    if (nthreads > 0) asm volatile("bar.sync 1, %0;" ::"r"(nthreads) : "memory");
}

template <int MAXW>
__device__ __noinline__ void substep(WS &s, const ModelDev &m, const EnvArgs &A, const grx_task_cfg &cfg, int lane, int env, int deci) {
    const float dt = cfg.sim_dt;
    mass_and_bias(s, m, cfg.gravity, lane);
    if (A.dbg_M != nullptr && A.dbg_index == (int)(blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5))) {  // debug_dynamics: export before factorisation
        for (int i = lane; i < NV * NV; i += 32) A.dbg_M[i] = s.M[i / NV][i % NV];
        if (lane < NV) A.dbg_h[lane] = s.h[lane];
        __syncwarp();
    }
    cholesky<MAXW>(s, lane);
    const float mu = 0.5f * (s.cst[C_FRIC] + A.terrain.friction), rest = 0.5f * (s.cst[C_REST] + A.terrain.restitution);
    // ---- contact detection: lane = sphere
    bool act = false;
    float n[3] = {0, 0, 1}, xs[3] = {0, 0, 0}, dist = 0, rad = 0;
    int sb = 0;
    int cell[3] = {0, 0, 0};
    unsigned long long sig_item = 0ull;   // this lane's contribution to the active-set signature (debug export, see phys_impl.h substep)
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
    const int nc = min(__popc(bal), KC);
    if (act && rank < KC) {
        float xc[3], t1[3], t2[3];
#pragma unroll
        for (int k = 0; k < 3; k++) xc[k] = xs[k] - n[k] * rad;
        const bool usey = n[0] > 0.9f || n[0] < -0.9f;
        const float e[3] = {usey ? 0.f : 1.f, usey ? 1.f : 0.f, 0.f};
        const float dn = dot3(e, n);
#pragma unroll
        for (int k = 0; k < 3; k++) t1[k] = e[k] - dn * n[k];
        const float inv = 1.0f / sqrtf(dot3(t1, t1));
#pragma unroll
        for (int k = 0; k < 3; k++) t1[k] *= inv;
        cross3(n, t1, t2);
        float rv[3], vc[3];
#pragma unroll
        for (int k = 0; k < 3; k++) rv[k] = xc[k] - s.o[sb][k];
        cross3(s.w[sb], rv, vc);
        const float vn0 = dot3(n, s.vo[sb]) + dot3(n, vc);
        float target;
        if (dist > 0) target = -dist / dt;
        else { target = -dist * cfg.erp / dt; if (target > cfg.max_depen_vel) target = cfg.max_depen_vel; }
        const bool bounce = vn0 < -cfg.bounce_threshold && -rest * vn0 > target;
        if (bounce) target = -rest * vn0;
        sig_item = mix64((1ull << 56) | (unsigned long long)lane | ((unsigned long long)cell[0] << 6) | ((unsigned long long)cell[1] << 18) |
                         ((unsigned long long)cell[2] << 30) | ((unsigned long long)(bounce ? 1 : 0) << 33) | ((unsigned long long)(usey ? 1 : 0) << 34));
#pragma unroll
        for (int k = 0; k < 3; k++) { s.cfr[rank][k] = n[k]; s.cfr[rank][3 + k] = t1[k]; s.cfr[rank][6 + k] = t2[k]; s.cpt[rank][k] = xc[k]; }
        s.cpt[rank][3] = target;
        s.cbody[rank] = sb;
        s.clink[rank] = m.sph_link[lane];
    }
    // ---- joint-limit rows (predicted with the pre-step rate): lane = joint
    float lsgn = 0, ltgt = 0;
    if (lane < ND) {
        const float q = s.rec[R_DOFPOS + lane], qn = q + dt * s.rec[R_DOFVEL + lane];
        if (qn < m.dof_lower[lane]) { lsgn = 1.f; ltgt = (m.dof_lower[lane] - q) / dt; }
        else if (qn > m.dof_upper[lane]) { lsgn = -1.f; ltgt = (q - m.dof_upper[lane]) / dt; }
    }
    const unsigned lbal = __ballot_sync(FULL, lsgn != 0.f);
    const int lrank = __popc(lbal & ((1u << lane) - 1u));
    const int nlim = min(__popc(lbal), KLIM);
    if (lsgn != 0.f && lrank < KLIM) {
        s.limj[lrank] = lane; s.lims[lrank] = lsgn; s.limt[lrank] = ltgt;
        sig_item += mix64((2ull << 56) | (unsigned long long)lane | ((unsigned long long)(lsgn < 0.f ? 1 : 0) << 6));
    }
    if (A.dbg_sig != nullptr) {   // debug export of the discrete decisions of this substep (parity tests)
        unsigned long long tot = sig_item;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const unsigned l2 = __shfl_xor_sync(FULL, (unsigned)tot, o), h2 = __shfl_xor_sync(FULL, (unsigned)(tot >> 32), o);
            tot += ((unsigned long long)h2 << 32) | l2;
        }
        if (lane == 0) A.dbg_sig[(size_t)env * A.dbg_sig_stride + deci] = tot;
    }
    __syncwarp();
    const int nrow = 3 * nc + nlim;  // <= 31
    // ---- build row lane's Jacobian, solve Y = M^-1 J^T ; lane 31 solves the unconstrained update
    float J[NV], x[NV], bias = 0.f;
#pragma unroll
    for (int i = 0; i < NV; i++) J[i] = 0.f;
    if (lane < 3 * nc) {
        const int c = lane / 3, k = lane % 3, b = s.cbody[c];
        const float *dir = s.cfr[c] + 3 * k, *xc = s.cpt[c];
        const int leg = (b - 1) / CH, depth = b == 0 ? 0 : (b - 1) % CH + 1;
#pragma unroll
        for (int jj = 0; jj < ND; jj++) {
            const bool on = b > 0 && (jj / CH) == leg && (jj % CH) < depth;
            float r[3], t[3];
#pragma unroll
            for (int i = 0; i < 3; i++) r[i] = xc[i] - s.o[jj + 1][i];
            cross3(s.a[jj + 1], r, t);
            J[jj] = on ? dot3(t, dir) : 0.f;
        }
        float r[3], t[3];
#pragma unroll
        for (int i = 0; i < 3; i++) r[i] = xc[i] - s.o[0][i];
        cross3(r, dir, t);
#pragma unroll
        for (int i = 0; i < 3; i++) { J[ND + i] = dir[i]; J[ND + 3 + i] = t[i]; }
        bias = k == 0 ? xc[3] : 0.f;
    } else if (lane < nrow) {
        const int l = lane - 3 * nc, j = s.limj[l];
        const float sg = s.lims[l];
#pragma unroll
        for (int jj = 0; jj < ND; jj++) J[jj] = jj == j ? sg : 0.f;
        bias = s.limt[l];
    }
    const bool rhs_lane = lane == 31;
#pragma unroll
    for (int i = 0; i < NV; i++) x[i] = rhs_lane ? ((i < ND ? s.tau[i] : 0.f) - s.h[i]) : J[i];
    chol_solve(s, x);
    float invA = 0.f;
    if (lane < nrow) {
        float arr = 0.f;
#pragma unroll
        for (int i = 0; i < NV; i++) arr += J[i] * x[i];
        invA = 1.0f / arr;
#pragma unroll
        for (int i = 0; i < NV; i += 4) *reinterpret_cast<float4 *>(&s.Y[lane][y_slot(lane, i >> 2)]) = make_float4(x[i], x[i + 1], x[i + 2], x[i + 3]);
        s.rowc[lane][0] = invA;
        s.rowc[lane][1] = bias;
    }
    const float *rs = s.rec + R_ROOT;
    if (rhs_lane) {
#pragma unroll
        for (int i = 0; i < NV; i++) s.u[i] = (i < ND ? s.rec[R_DOFVEL + i] : rs[7 + i - ND]) + dt * x[i];
    }
    __syncwarp();
    // ---- constraint-space velocity w = J u*, Delassus column block As[r][lane] = J_lane . Y_r
    float wv = 0.f;
#pragma unroll
    for (int i = 0; i < NV; i++) wv += J[i] * s.u[i];
#pragma unroll 2
    for (int r = 0; r < nrow; r++) {
        float acc = 0.f;
#pragma unroll
        for (int i = 0; i < NV; i += 4) {
            const float4 y = *reinterpret_cast<const float4 *>(&s.Y[r][y_slot(r, i >> 2)]);
            acc += J[i] * y.x + J[i + 1] * y.y + J[i + 2] * y.z + J[i + 3] * y.w;
        }
        if (lane < ASP) s.As[r * ASP + lane] = acc;
    }
    __syncwarp();
    // ---- projected Gauss-Seidel in constraint space.  Row r is owned by lane r; every lane evaluates the row update from the
    // owner's (w, lambda) obtained with two independent shuffles, so the multiplier change d is warp-uniform without a third one.
    // Rows are swept contact by contact (normal, then the two friction rows bounded by mu * lambda_n), then the joint limits.
    float lam = 0.f;
    const float *Ac = &s.As[min(lane, ASP - 1)];
#pragma unroll 1
    for (int it = 0; it < cfg.solver_iters; it++) {
        int r = 0;
#pragma unroll 1
        for (int c = 0; c < nc; c++) {
            float lim;
            {
                const float wr = __shfl_sync(FULL, wv, r), l0 = __shfl_sync(FULL, lam, r);
                const float2 rc = *reinterpret_cast<const float2 *>(s.rowc[r]);
                const float ln = fmaxf(l0 - (wr - rc.y) * rc.x, 0.f);
                if (lane == r) lam = ln;
                wv += Ac[r * ASP] * (ln - l0);
                lim = mu * ln;
                r++;
            }
#pragma unroll
            for (int k = 0; k < 2; k++, r++) {
                const float wr = __shfl_sync(FULL, wv, r), l0 = __shfl_sync(FULL, lam, r);
                const float2 rc = *reinterpret_cast<const float2 *>(s.rowc[r]);
                const float ln = fminf(fmaxf(l0 - (wr - rc.y) * rc.x, -lim), lim);
                if (lane == r) lam = ln;
                wv += Ac[r * ASP] * (ln - l0);
            }
        }
#pragma unroll 1
        for (; r < nrow; r++) {
            const float wr = __shfl_sync(FULL, wv, r), l0 = __shfl_sync(FULL, lam, r);
            const float2 rc = *reinterpret_cast<const float2 *>(s.rowc[r]);
            const float ln = fmaxf(l0 - (wr - rc.y) * rc.x, 0.f);
            if (lane == r) lam = ln;
            wv += Ac[r * ASP] * (ln - l0);
        }
    }
    // ---- u = u* + sum_r Y_r lam_r (lane i < NV owns component i)
    float unew = lane < NV ? s.u[lane] : 0.f;
#pragma unroll 1
    for (int r = 0; r < nrow; r++) {
        const float lr = __shfl_sync(FULL, lam, r);
        if (lane < NV) unew += s.Y[r][y_slot(r, lane >> 2) + (lane & 3)] * lr;
    }
    for (int i = lane; i < m.nl * 3; i += 32) s.cf[i] = 0.f;
    __syncwarp();
    {   // net contact force per URDF link (world frame, on the body) = impulse / dt
        const float l0 = __shfl_sync(FULL, lam, min(3 * lane, 31)), l1 = __shfl_sync(FULL, lam, min(3 * lane + 1, 31)),
                    l2 = __shfl_sync(FULL, lam, min(3 * lane + 2, 31));
        if (lane < nc) {
            const float *f = s.cfr[lane];
            const int link = s.clink[lane];
#pragma unroll
            for (int k = 0; k < 3; k++) atomicAdd(&s.cf[3 * link + k], (f[k] * l0 + f[3 + k] * l1 + f[6 + k] * l2) / dt);
        }
    }
    // ---- joint-rate limit + integrate
    if (lane < ND) {
        const float vl = m.dof_vel_limit[lane];
        const float v = fminf(fmaxf(unew, -vl), vl);
        s.rec[R_DOFVEL + lane] = v;
        s.rec[R_DOFPOS + lane] += dt * v;
    } else if (lane < NV) {
        s.rec[R_ROOT + 7 + lane - ND] = unew;
    }
    __syncwarp();
    if (lane == 0) {
        float *rt = s.rec + R_ROOT;
#pragma unroll
        for (int k = 0; k < 3; k++) rt[k] += dt * rt[7 + k];
        const float wx = rt[10], wy = rt[11], wz = rt[12];
        const float wn = sqrtf(wx * wx + wy * wy + wz * wz), th = wn * dt;
        float sn, cs;
        __sincosf(0.5f * th, &sn, &cs);
        const float sc = wn > 1e-9f ? sn / wn : 0.5f * dt;
        const float dq[4] = {wx * sc, wy * sc, wz * sc, cs};
        float *p = rt + 3;
        const float qx = dq[3] * p[0] + dq[0] * p[3] + dq[1] * p[2] - dq[2] * p[1];
        const float qy = dq[3] * p[1] - dq[0] * p[2] + dq[1] * p[3] + dq[2] * p[0];
        const float qz = dq[3] * p[2] + dq[0] * p[1] - dq[1] * p[0] + dq[2] * p[3];
        const float qw = dq[3] * p[3] - dq[0] * p[0] - dq[1] * p[1] - dq[2] * p[2];
        const float nn = rsqrtf(qx * qx + qy * qy + qz * qz + qw * qw);
        p[0] = qx * nn; p[1] = qy * nn; p[2] = qz * nn; p[3] = qw * nn;
    }
    __syncwarp();
}



// MAXW: warps per CTA the build is compiled for — 16 (128 registers per thread) or 28 (72 registers: one wave for 4096 robots on 148 SMs)
template <bool PHYS, int MAXW>
__global__ void __launch_bounds__(MAXW * 32, 1) env_step_kernel(const __grid_constant__ EnvArgs A,
                                                                const __grid_constant__ grx_task_cfg cfg) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    ModelDev &m = *reinterpret_cast<ModelDev *>(smem_raw);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    WS &s = *reinterpret_cast<WS *>(smem_raw + ((sizeof(ModelDev) + 15) & ~15) + (size_t)warp * sizeof(WS));
    const int wpc = blockDim.x >> 5;   // warps (= robots) per CTA, chosen by the launch
    const int e = blockIdx.x * wpc + warp;
    {   // model tables -> shared memory (divergent per-body indexing would serialise on the constant bank)
        const int4 *src = reinterpret_cast<const int4 *>(A.model);
        int4 *dst = reinterpret_cast<int4 *>(&m);
        for (int i = threadIdx.x; i < (int)(sizeof(ModelDev) / 16); i += blockDim.x) dst[i] = src[i];
    }
    if (blockIdx.x == 0 && threadIdx.x < ACC_W) A.episode_accum_next[threadIdx.x] = 0.f;   // nobody accumulates into the next slot during this launch
    const bool valid = e < A.N;
    if (valid) {
        if (lane == 0) {
            mbar_init(&s.mbar, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
        if (lane == 0) {   // stage this robot's state + parameter records with TMA bulk copies
            mbar_expect_tx(&s.mbar, (REC_F + CST_F) * 4);
            bulk_g2s(s.rec, A.rec + (size_t)e * REC_F, REC_F * 4, &s.mbar);
            bulk_g2s(s.cst, A.cst + (size_t)e * CST_F, CST_F * 4, &s.mbar);
        }
    }
    __syncthreads();  // model copy visible
    if (!valid) return;
    mbar_wait(&s.mbar, 0);

    const int nd = ND;
    Draw draw;
    draw.U = A.U ? A.U + (size_t)e * GRX_RNG_K : nullptr;
    draw.k0 = (uint32_t)cfg.seed; draw.k1 = (uint32_t)(cfg.seed >> 32);
    draw.gid = (uint32_t)(cfg.env_id_offset + e);
    draw.step_lo = (uint32_t)A.step_index; draw.step_hi = (uint32_t)(A.step_index >> 32);

    // ---- clip_actions (legged_robot_fftai.py:171-177)
    float act_l = 0.f, last_act_l = 0.f;
    if (lane < nd) {
        act_l = fminf(fmaxf(A.actions[(size_t)e * nd + lane], cfg.clip_actions_min[lane]), cfg.clip_actions_max[lane]);
        last_act_l = s.rec[R_LASTACT + lane];
    }
    float ff_acc = 0.f, fl_acc[3] = {0, 0, 0};  // lane f < NF: substep sums of |F_foot|, |v_foot| (FF:79-81)
    float foot_z = 0.f;
    float torso_q[4] = {0, 0, 0, 1};
    if (PHYS) {
        // Re-converge the CTA's warps once per substep: the substep is ~60 KB of mostly straight-line code, and with every warp
        // of a scheduler in a different place the instruction caches thrash (ncu: 28 % of warp samples `no_instructions`);
        // aligned warps share the fetched lines (measured 536 -> 445 us per launch).  More barriers per substep cost more than they save.
        const int cta_valid = min(wpc, A.N - (int)blockIdx.x * wpc) * 32;   // threads that did not exit above
#pragma unroll 1
        for (int deci = 0; deci <= cfg.decimation; deci++) {
            cta_align(A.dbg_M == nullptr ? cta_valid : 0);
            kinematics<MAXW>(s, m, lane);
            if (deci > 0 && lane < NF) {   // foot statistics of the substep just integrated (FF:79-81)
                const int l = m.foot_link[lane], b = m.foot_body[lane];
                float r[3], t[3];
                m3v(s.R[b], m.foot_pos[lane], r);
                cross3(s.w[b], r, t);
                const float *f = s.cf + 3 * l;
                ff_acc += sqrtf(f[0] * f[0] + f[1] * f[1] + f[2] * f[2]);
#pragma unroll
                for (int k = 0; k < 3; k++) fl_acc[k] += fabsf(s.vo[b][k] + t[k]);
                foot_z = s.o[b][2] + r[2];
            }
            if (deci == cfg.decimation) break;
            if (lane < nd) {   // _compute_torques (legged_robot.py:691-713) with the action delay of FF:58-61
                const float a = ((float)deci < A.delay) ? last_act_l : act_l;
                float t = m.kp[lane] * (a * cfg.action_scale + m.q0[lane] - s.rec[R_DOFPOS + lane]) - m.kd[lane] * s.rec[R_DOFVEL + lane];
                t *= s.cst[C_MOTOR + lane];
                const float lim = m.dof_effort[lane];
                s.tau[lane] = fminf(fmaxf(t, -lim), lim);
            }
            __syncwarp();
            substep<MAXW>(s, m, A, cfg, lane, e, deci);
        }
        if (A.dbg_M != nullptr) return;
        const float invd = 1.0f / (float)cfg.decimation;
        ff_acc *= invd;
#pragma unroll
        for (int k = 0; k < 3; k++) fl_acc[k] *= invd;
        float Rt[9];
        m3m(s.R[m.torso_body], m.torso_rot, Rt);
        mat2quat(Rt, torso_q);
        if (A.rigid_body_states != nullptr) {   // compat export: every URDF link's state after the last substep (refresh_rigid_body_state_tensor, FF:75)
            for (int l = lane; l < m.nl; l += 32) {
                const int b = __ldg(A.link_body + l);
                float lp[3], lr[9], r[3], t[3], Rl[9], ql[4];
#pragma unroll
                for (int k = 0; k < 3; k++) lp[k] = __ldg(A.link_pos + 3 * l + k);
#pragma unroll
                for (int k = 0; k < 9; k++) lr[k] = __ldg(A.link_rot + 9 * l + k);
                m3v(s.R[b], lp, r);
                cross3(s.w[b], r, t);
                m3m(s.R[b], lr, Rl);
                mat2quat(Rl, ql);
                float *o = A.rigid_body_states + ((size_t)e * m.nl + l) * 13;
#pragma unroll
                for (int k = 0; k < 3; k++) { o[k] = s.o[b][k] + r[k]; o[7 + k] = s.vo[b][k] + t[k]; o[10 + k] = s.w[b][k]; }
#pragma unroll
                for (int k = 0; k < 4; k++) o[3 + k] = ql[k];
            }
        }
        if (A.foot_state != nullptr && lane < NF) {   // compat export: feet link states (play.py / tests)
            float *fs = A.foot_state + ((size_t)e * NF + lane) * 13;
            const int b = m.foot_body[lane];
            float r[3], t[3], Rf[9], qf[4];
            m3v(s.R[b], m.foot_pos[lane], r);
            cross3(s.w[b], r, t);
            m3m(s.R[b], m.torso_rot, Rf);   // foot links share their body's frame up to the link rotation (identity for GRx)
            mat2quat(s.R[b], qf);
#pragma unroll
            for (int k = 0; k < 3; k++) { fs[k] = s.o[b][k] + r[k]; fs[7 + k] = s.vo[b][k] + t[k]; fs[10 + k] = s.w[b][k]; }
#pragma unroll
            for (int k = 0; k < 4; k++) fs[3 + k] = qf[k];
        }
    } else {
        // injected physics outputs (grx_env_post_physics)
        if (lane < nd) s.tau[lane] = A.inj.torques[(size_t)e * nd + lane];
        for (int i = lane; i < m.nl * 3; i += 32) s.cf[i] = A.inj.contact_forces[(size_t)e * m.nl * 3 + i];
        if (lane < NF) {
            foot_z = A.inj.foot_state[((size_t)e * NF + lane) * 13 + 2];
            ff_acc = A.inj.avg_foot_force[(size_t)e * NF + lane];
#pragma unroll
            for (int k = 0; k < 3; k++) fl_acc[k] = A.inj.avg_foot_linvel[((size_t)e * NF + lane) * 3 + k];
        }
#pragma unroll
        for (int k = 0; k < 4; k++) torso_q[k] = A.inj.torso_quat[(size_t)e * 4 + k];
        __syncwarp();
    }

    // =====================================================================================================
    // post_physics_step (legged_robot.py:269-305, legged_robot_fftai.py:90-133) .. compute_observations: grx_task.cuh
    // =====================================================================================================
    // scratch: measured heights + the noise-free obs over the Delassus block, the privileged-observation row over Y (both dead after the physics)
    task_post_physics(s.rec, Lay10(), m, A, cfg, draw, lane, e, act_l, last_act_l, ff_acc, fl_acc, foot_z, torso_q, s.tau, s.cf, &s.As[0],
                      &s.As[0] + NHMAX, &s.Y[0][0], s.rterm);
    __syncwarp();
    fence_async_smem();   // generic-proxy writes to smem -> visible to the bulk-copy (async) proxy
    __syncwarp();
    if (lane == 0) {
        bulk_s2g(A.rec + (size_t)e * REC_F, s.rec, REC_F * 4);
        bulk_s2g(A.pri_obs + (size_t)e * cfg.num_pri_obs, &s.Y[0][0], (uint32_t)cfg.num_pri_obs * 4);
        bulk_commit_wait();
    }
}

// ---- structured-trimesh builder on the device (terrain_utils.py:315-328): per vertex, the shift (in cells) of the steep-edge snapping, straight
// from the int16 sample grid.  The differences wrap in int16 like the reference's numpy arithmetic on the int16 array; thr = slope_threshold *
// horizontal_scale / vertical_scale is compared in double like numpy does.  One thread per vertex.
__global__ void trimesh_moves_kernel(const short *h, int rows, int cols, double thr, signed char *mv) {
    const int id = blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= rows * cols) return;
    const int i = id / cols, j = id % cols;
    auto H = [&](int a, int b) { return h[(size_t)a * cols + b]; };
    auto steep = [&](short hi, short lo) { return (double)(short)(hi - lo) > thr; };   // hf[hi] - hf[lo] > thr, int16 wrap-around
    int mx = 0, my = 0, mc = 0;
    if (i < rows - 1 && steep(H(i + 1, j), H(i, j))) mx += 1;
    if (i >= 1 && steep(H(i - 1, j), H(i, j))) mx -= 1;
    if (j < cols - 1 && steep(H(i, j + 1), H(i, j))) my += 1;
    if (j >= 1 && steep(H(i, j - 1), H(i, j))) my -= 1;
    if (i < rows - 1 && j < cols - 1 && steep(H(i + 1, j + 1), H(i, j))) mc += 1;
    if (i >= 1 && j >= 1 && steep(H(i - 1, j - 1), H(i, j))) mc -= 1;
    mv[2 * (size_t)id] = (signed char)(mx + (mx == 0 ? mc : 0));
    mv[2 * (size_t)id + 1] = (signed char)(my + (my == 0 ? mc : 0));
}
// per cell: does any vertex of the 3 x 3 cells around it (vertices i-1 .. i+2, j-1 .. j+2) carry a shift?
__global__ void trimesh_near_kernel(const signed char *mv, int rows, int cols, unsigned char *near_mv) {
    const int id = blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= rows * cols) return;
    const int i = id / cols, j = id % cols;
    int any = 0;
    for (int a = max(i - 1, 0); a <= min(i + 2, rows - 1); a++)
        for (int b = max(j - 1, 0); b <= min(j + 2, cols - 1); b++) any |= mv[2 * ((size_t)a * cols + b)] | mv[2 * ((size_t)a * cols + b) + 1];
    near_mv[id] = any ? 1 : 0;
}

// Host-invoked reset_idx (BaseTask.reset(), base_task.py:117-121): one warp per listed env (ids == nullptr: all envs).
__global__ void __launch_bounds__(WARPS_PER_CTA * 32) env_reset_kernel(const __grid_constant__ EnvArgs A, const __grid_constant__ grx_task_cfg cfg,
                                                                      const int *ids, int n, int curriculum_active) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    ModelDev &m = *reinterpret_cast<ModelDev *>(smem_raw);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    WS &s = *reinterpret_cast<WS *>(smem_raw + ((sizeof(ModelDev) + 15) & ~15) + (size_t)warp * sizeof(WS));
    {
        const int4 *src = reinterpret_cast<const int4 *>(A.model);
        int4 *dst = reinterpret_cast<int4 *>(&m);
        for (int i = threadIdx.x; i < (int)(sizeof(ModelDev) / 16); i += blockDim.x) dst[i] = src[i];
    }
    if (blockIdx.x == 0 && threadIdx.x < ACC_W) A.episode_accum_next[threadIdx.x] = 0.f;
    __syncthreads();
    const int w = blockIdx.x * WARPS_PER_CTA + warp;
    if (w >= n) return;
    const int e = ids ? ids[w] : w;
    if (e < 0 || e >= A.N) return;
    float *g = A.rec + (size_t)e * REC_F;
    for (int i = lane; i < REC_F; i += 32) s.rec[i] = g[i];
    __syncwarp();
    Draw draw;
    draw.U = A.U ? A.U + (size_t)e * GRX_RNG_K : nullptr;
    draw.k0 = (uint32_t)cfg.seed ^ 0x5bd1e995u; draw.k1 = (uint32_t)(cfg.seed >> 32);
    draw.gid = (uint32_t)(cfg.env_id_offset + e);
    draw.step_lo = (uint32_t)A.step_index; draw.step_hi = (uint32_t)(A.step_index >> 32);
    const float cx = s.rec[R_CMD], cy = s.rec[R_CMD + 1];
    reset_env(s.rec, Lay10(), m, A, cfg, draw, lane, sqrtf(cx * cx + cy * cy), curriculum_active != 0);
    if (lane == 0 && cfg.curriculum) atomicAdd(A.episode_accum + NREW + 1, (float)__float_as_int(s.rec[R_TLEVEL]));
    for (int i = lane; i < REC_F; i += 32) g[i] = s.rec[i];
}

}  // namespace

// =========================================================================================================
// Host side: C ABI (include/grx_b200.h)
// =========================================================================================================
static thread_local std::string g_err;
extern "C" const char *grx_last_error(void) { return g_err.c_str(); }
extern "C" int grx_version(void) { return 100; }
int grx_set_error(int code, const std::string &msg) { g_err = msg; return code; }

#define CK(call)                                                                                             \
    do {                                                                                                     \
        cudaError_t err__ = (call);                                                                          \
        if (err__ != cudaSuccess)                                                                            \
            return grx_set_error(GRX_E_CUDA, std::string(#call) + ": " + cudaGetErrorString(err__));        \
    } while (0)

struct grx_env {
    int N = 0, device = 0, nl = 0;
    grx_task_cfg cfg;
    ModelDev hmodel;
    ModelDev *dmodel = nullptr;
    float *rec = nullptr, *cst = nullptr, *obs = nullptr, *pri_obs = nullptr, *rew = nullptr, *torques = nullptr;
    float *contact_forces = nullptr, *foot_state = nullptr, *episode_accum = nullptr, *terrain_origins = nullptr;
    float *actions_stage = nullptr;
    unsigned long long *active_sig = nullptr;   // debug export, allocated by grx_env_debug_active_sig
    float *rigid_body_states = nullptr, *dof_state = nullptr;   // compat exports (allocated on first grx_env_get_buffer request)
    long long *ep_len64 = nullptr;
    int *d_link_body = nullptr;                                  // per-link tables for the rigid_body_states export
    float *d_link_pos = nullptr, *d_link_rot = nullptr;
    unsigned char *reset = nullptr, *time_out = nullptr;
    short *heights = nullptr;
    signed char *moves = nullptr;         // structured trimesh: vertex shifts
    unsigned char *near_moved = nullptr;
    TerrainDev terrain;
    int t_rows = 1, t_cols = 1;
    bool params_set = false;
    size_t smem = 0;
    int warps_per_cta = WARPS_PER_CTA;   // robots per CTA of env_step_kernel: whole waves of one CTA per SM (env_warps_per_cta)
    uint64_t launches = 0;   // step / reset launches so far; launch k accumulates extras into ring slot k % ACC_RING
    // Record layout of this env (grx_task.cuh) and, for every model that is not the registered lower-limb tree, the generic-topology kernels'
    // handle (grx_envg.h; e.g. the full-body 32-DOF GR1T1 / GR1T2 with robot self-collision)
    LayR lay;
    int nd = ND;
    void *g = nullptr;
};

// One CTA per SM is resident (registers + shared memory); with w warps per CTA a launch of N robots takes ceil(N / (w * SMs)) waves.
// Take the number of waves of the widest CTA and shrink the CTA until those waves are evenly filled (4096 robots on 148 SMs:
// 2 waves of 14 warps instead of one wave of 16 and one 73 % full).
static int env_warps_per_cta(int N, int sms, int maxw) {
    const int waves = (N + maxw * sms - 1) / (maxw * sms);
    int w = (N + waves * sms - 1) / (waves * sms);
    if (w < 1) w = 1;
    if (w > maxw) w = maxw;
    return w;
}
static size_t env_smem_bytes(int warps) { return ((sizeof(ModelDev) + 15) & ~(size_t)15) + (size_t)warps * sizeof(WS); }
// The two builds of the step kernel: <= 16 warps per CTA at 128 registers per thread, or <= 28 warps at 72 registers (more spills, but
// twice the warps to hide latency with and HALF THE WAVES: 4096 robots = one wave of 28 warps per SM instead of two of 14).
template <bool PHYS>
static void launch_env_step(const grx_env *e, const EnvArgs &A, const grx_task_cfg &cfg, cudaStream_t st);

extern "C" int grx_env_create(const grx_model_desc *md, const grx_task_cfg *cfg, int32_t num_envs, int32_t device, grx_env **out) {
    if (!md || !cfg || !out || num_envs <= 0) return grx_set_error(GRX_E_INVALID, "grx_env_create: null argument or num_envs <= 0");
    // The registered lower-limb tree (floating base + 2 chains of 5) runs on the specialised fused kernel; every other revolute tree (<= 36
    // bodies / 32 DOF: the full-body GR1T1 / GR1T2) — or every model, with GRX_ENV_GENERIC=1 — on the generic-topology kernels.
    static const int want_parent[NB] = {-1, 0, 1, 2, 3, 4, 0, 6, 7, 8, 9};
    bool generic = md->nb != NB || md->nd != ND || md->nl > NLMAX || md->ns > NSMAX;
    for (int b = 0; b < NB && !generic; b++) generic = md->parent[b] != want_parent[b];
    if (const char *gen = getenv("GRX_ENV_GENERIC")) generic = generic || atoi(gen) != 0;
    generic = generic || md->nankle != 2;
    if (md->nf != NF || (md->nankle != 2 && md->nankle != 4) || md->nd < 1 || md->nd > 32)
        return grx_set_error(GRX_E_INVALID, "grx_env_create: the task needs 2 feet, 2 or 4 ankle DOF and 1..32 actuated DOF");
    const int H = cfg->num_height_points, nd_ = md->nd;
    if (cfg->num_actions != nd_ || cfg->num_obs != 9 + 3 * nd_ || H > NHMAX || cfg->num_pri_obs != cfg->num_obs + 8 + H ||
        H != cfg->n_points_x * cfg->n_points_y || cfg->n_points_x > 16 || cfg->n_points_y > 16 || (!generic && (cfg->num_pri_obs * 4) % 16 != 0))
        return grx_set_error(GRX_E_INVALID, "grx_env_create: observation layout mismatch (need obs = 9+3*nd, pri_obs = obs+8+H, H <= 128)");
    if (cfg->decimation < 1 || cfg->solver_iters < 1 || cfg->resample_interval < 1)
        return grx_set_error(GRX_E_INVALID, "grx_env_create: decimation / solver_iters / resample_interval must be >= 1");
    CK(cudaSetDevice(device));
    grx_env *e = new grx_env();
    e->N = num_envs; e->device = device; e->cfg = *cfg; e->nl = md->nl; e->nd = nd_; e->lay = make_layout(nd_);
    if (generic) {
        int rc = grx::envg_create(md, cfg, device, &e->g);
        if (rc) { delete e; return rc; }
    }
    ModelDev &m = e->hmodel;
    memset(&m, 0, sizeof(m));
    for (int b = 0; b < NB && !generic; b++) {
        memcpy(m.jpos[b], md->jpos + 3 * b, 12); memcpy(m.jrot[b], md->jrot + 9 * b, 36); memcpy(m.axis[b], md->axis + 3 * b, 12);
        m.mass[b] = md->mass[b]; memcpy(m.com[b], md->com + 3 * b, 12); memcpy(m.inertia[b], md->inertia + 6 * b, 24);
    }
    for (int j = 0; j < ND && !generic; j++) {
        m.dof_lower[j] = md->dof_lower[j]; m.dof_upper[j] = md->dof_upper[j]; m.dof_vel_limit[j] = md->dof_vel_limit[j];
        m.dof_effort[j] = md->dof_effort[j]; m.soft_lower[j] = md->soft_lower[j]; m.soft_upper[j] = md->soft_upper[j];
        m.kp[j] = md->kp[j]; m.kd[j] = md->kd[j]; m.q0[j] = md->default_pos[j];
    }
    m.ns = md->ns; m.nl = md->nl;
    for (int s = 0; s < md->ns && !generic; s++) {
        m.sph_body[s] = md->sph_body[s]; m.sph_link[s] = md->sph_link[s]; memcpy(m.sph_pos[s], md->sph_pos + 3 * s, 12);
        m.sph_rad[s] = md->sph_rad[s];
    }
    for (int f = 0; f < NF; f++) {
        const int l = md->foot_links[f];
        m.foot_link[f] = l; m.foot_body[f] = md->link_body[l]; memcpy(m.foot_pos[f], md->link_pos + 3 * l, 12);
    }
    memcpy(m.torso_rot, md->link_rot + 9 * md->torso_link, 36);
    m.torso_body = md->link_body[md->torso_link];
    m.term_mask = 0;
    for (int t = 0; t < md->nterm; t++) m.term_mask |= 1ull << md->term_links[t];
    m.ankle_dof[0] = md->ankle_dofs[0]; m.ankle_dof[1] = md->ankle_dofs[1];
    m.jrot_nonident = 0;
    for (int b = 0; b < NB; b++) {
        static const float I3[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
        for (int k = 0; k < 9; k++)
            if (m.jrot[b][k] != I3[k]) m.jrot_nonident |= 1 << b;
    }
    const size_t N = num_envs;
#define ALLOC(ptr, count) CK(cudaMalloc((void **)&(ptr), (count))); CK(cudaMemset((ptr), 0, (count)))
    ALLOC(e->dmodel, sizeof(ModelDev));
    const LayR &L = e->lay;
    ALLOC(e->rec, N * L.rec_f * 4); ALLOC(e->cst, N * L.cst_f * 4);
    ALLOC(e->obs, N * cfg->num_obs * 4); ALLOC(e->pri_obs, N * cfg->num_pri_obs * 4);
    ALLOC(e->rew, N * 4); ALLOC(e->torques, N * nd_ * 4); ALLOC(e->contact_forces, N * md->nl * 3 * 4);
    ALLOC(e->foot_state, N * NF * 13 * 4); ALLOC(e->episode_accum, ACC_RING * ACC_W * 4); ALLOC(e->actions_stage, N * nd_ * 4);
    ALLOC(e->reset, N); ALLOC(e->time_out, N);
    ALLOC(e->terrain_origins, 3 * 4);
#undef ALLOC
    CK(cudaMemcpy(e->dmodel, &m, sizeof(ModelDev), cudaMemcpyHostToDevice));
    CK(cudaMalloc((void **)&e->d_link_body, (size_t)md->nl * 4)); CK(cudaMalloc((void **)&e->d_link_pos, (size_t)md->nl * 12)); CK(cudaMalloc((void **)&e->d_link_rot, (size_t)md->nl * 36));
    CK(cudaMemcpy(e->d_link_body, md->link_body, (size_t)md->nl * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(e->d_link_pos, md->link_pos, (size_t)md->nl * 12, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(e->d_link_rot, md->link_rot, (size_t)md->nl * 36, cudaMemcpyHostToDevice));
    {   // identity root quaternion
        std::vector<float> h(N * L.rec_f, 0.f);
        for (size_t i = 0; i < N; i++) h[i * L.rec_f + L.root + 6] = 1.f;
        CK(cudaMemcpy(e->rec, h.data(), h.size() * 4, cudaMemcpyHostToDevice));
    }
    e->terrain.type = 0; e->terrain.rows = e->terrain.cols = 0; e->terrain.h = nullptr; e->terrain.mv = nullptr; e->terrain.near_mv = nullptr;
    e->terrain.hscale = 1.f; e->terrain.vscale = 1.f; e->terrain.border = 0.f; e->terrain.friction = 1.f; e->terrain.restitution = 0.f;
    {
        int sms = 0;
        if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device) != cudaSuccess || sms <= 0) sms = 148;
        // Measured on B200 (profiles/r2f_env_*): the wide build runs 4096 robots as ONE wave of 28 warps per SM in 415 us vs 424 us for two waves of
        // 14 warps at 128 registers (IPC 2.06 vs 1.89) — but its 72-register budget spills 544 B per thread per substep, which turn into 411 MB
        // of DRAM writes per launch (3.9 MB for the 128-register build).  +2 % is not worth 100x the memory traffic: opt-in (GRX_ENV_WIDE=1).
        int maxw = WARPS_PER_CTA;
        if (const char *w = getenv("GRX_ENV_WIDE")) maxw = atoi(w) ? WARPS_PER_CTA_WIDE : WARPS_PER_CTA;
        e->warps_per_cta = env_warps_per_cta(num_envs, sms, maxw);
        if (const char *w = getenv("GRX_ENV_WARPS")) { const int v = atoi(w); if (v >= 1 && v <= WARPS_PER_CTA_WIDE) e->warps_per_cta = v; }
    }
    e->smem = env_smem_bytes(e->warps_per_cta > WARPS_PER_CTA ? WARPS_PER_CTA_WIDE : WARPS_PER_CTA);
    CK((cudaFuncSetAttribute(env_step_kernel<true, WARPS_PER_CTA>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)env_smem_bytes(WARPS_PER_CTA))));
    CK((cudaFuncSetAttribute(env_step_kernel<false, WARPS_PER_CTA>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)env_smem_bytes(WARPS_PER_CTA))));
    CK((cudaFuncSetAttribute(env_step_kernel<true, WARPS_PER_CTA_WIDE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)env_smem_bytes(WARPS_PER_CTA_WIDE))));
    CK((cudaFuncSetAttribute(env_step_kernel<false, WARPS_PER_CTA_WIDE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)env_smem_bytes(WARPS_PER_CTA_WIDE))));
    CK(cudaFuncSetAttribute(env_reset_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)env_smem_bytes(WARPS_PER_CTA)));
    *out = e;
    return GRX_OK;
}

template <bool PHYS>
static void launch_env_step(const grx_env *e, const EnvArgs &A, const grx_task_cfg &cfg, cudaStream_t st) {
    if (e->g) { grx::envg_launch_step(e->g, A, cfg, e->lay, PHYS, st); return; }
    const int wpc = e->warps_per_cta, grid = (e->N + wpc - 1) / wpc;
    grx_count_launch();
    if (wpc > WARPS_PER_CTA) env_step_kernel<PHYS, WARPS_PER_CTA_WIDE><<<grid, wpc * 32, env_smem_bytes(wpc), st>>>(A, cfg);
    else env_step_kernel<PHYS, WARPS_PER_CTA><<<grid, wpc * 32, env_smem_bytes(wpc), st>>>(A, cfg);
}

extern "C" int grx_env_destroy(grx_env *e) {
    if (!e) return GRX_OK;
    cudaSetDevice(e->device);
    void *ptrs[] = {e->dmodel, e->rec, e->cst, e->obs, e->pri_obs, e->rew, e->torques, e->contact_forces, e->foot_state,
                    e->episode_accum, e->terrain_origins, e->actions_stage, e->reset, e->time_out, e->heights, e->active_sig,
                    e->rigid_body_states, e->dof_state, e->ep_len64, e->d_link_body, e->d_link_pos, e->d_link_rot, e->moves, e->near_moved};
    for (void *p : ptrs) if (p) cudaFree(p);
    grx::envg_destroy(e->g);
    delete e;
    return GRX_OK;
}

extern "C" int grx_env_set_terrain_plane(grx_env *e, float friction, float restitution) {
    if (!e) return grx_set_error(GRX_E_INVALID, "null env");
    e->terrain.type = 0; e->terrain.friction = friction; e->terrain.restitution = restitution;
    return GRX_OK;
}

extern "C" int grx_env_set_terrain_heightfield(grx_env *e, const int16_t *samples, int32_t rows, int32_t cols, float hscale,
                                               float vscale, float border, float friction, float restitution) {
    if (!e || !samples || rows < 2 || cols < 2) return grx_set_error(GRX_E_INVALID, "grx_env_set_terrain_heightfield: bad arguments");
    CK(cudaSetDevice(e->device));
    if (e->heights) { cudaFree(e->heights); e->heights = nullptr; }
    CK(cudaMalloc((void **)&e->heights, (size_t)rows * cols * 2));
    CK(cudaMemcpy(e->heights, samples, (size_t)rows * cols * 2, cudaMemcpyHostToDevice));
    e->terrain.type = 1; e->terrain.rows = rows; e->terrain.cols = cols; e->terrain.h = e->heights;
    e->terrain.mv = nullptr; e->terrain.near_mv = nullptr;
    e->terrain.hscale = hscale; e->terrain.vscale = vscale; e->terrain.border = border;
    e->terrain.friction = friction; e->terrain.restitution = restitution;
    return GRX_OK;
}

static int finish_trimesh(grx_env *e, const signed char *h_moves, double thr, int32_t rows, int32_t cols);

// The same with the samples already on the device (written by grx_terrain_generate): no host round trip
extern "C" int grx_env_set_terrain_device(grx_env *e, const int16_t *d_samples, int32_t rows, int32_t cols, float hscale, float vscale, float border,
                                          float slope_threshold, float friction, float restitution) {
    if (!e || !d_samples || rows < 2 || cols < 2) return grx_set_error(GRX_E_INVALID, "grx_env_set_terrain_device: bad arguments");
    CK(cudaSetDevice(e->device));
    if (e->heights) { cudaFree(e->heights); e->heights = nullptr; }
    CK(cudaMalloc((void **)&e->heights, (size_t)rows * cols * 2));
    CK(cudaMemcpy(e->heights, d_samples, (size_t)rows * cols * 2, cudaMemcpyDeviceToDevice));
    e->terrain.type = 1; e->terrain.rows = rows; e->terrain.cols = cols; e->terrain.h = e->heights;
    e->terrain.mv = nullptr; e->terrain.near_mv = nullptr;
    e->terrain.hscale = hscale; e->terrain.vscale = vscale; e->terrain.border = border;
    e->terrain.friction = friction; e->terrain.restitution = restitution;
    if (slope_threshold >= 0.f) return finish_trimesh(e, nullptr, (double)slope_threshold * (double)hscale / (double)vscale, rows, cols);
    return GRX_OK;
}

// Structured trimesh = the heightfield's sample grid + per-vertex shifts (moves).  finish_trimesh uploads / derives the device tables.
static int finish_trimesh(grx_env *e, const signed char *h_moves, double thr, int32_t rows, int32_t cols) {
    const size_t nvx = (size_t)rows * cols;
    if (e->moves) { cudaFree(e->moves); e->moves = nullptr; }
    if (e->near_moved) { cudaFree(e->near_moved); e->near_moved = nullptr; }
    CK(cudaMalloc((void **)&e->moves, 2 * nvx));
    CK(cudaMalloc((void **)&e->near_moved, nvx));
    const int blocks = (int)((nvx + 255) / 256);
    if (h_moves) CK(cudaMemcpy(e->moves, h_moves, 2 * nvx, cudaMemcpyHostToDevice));
    else { grx_count_launch(); trimesh_moves_kernel<<<blocks, 256>>>(e->heights, rows, cols, thr, e->moves); }
    grx_count_launch();
    trimesh_near_kernel<<<blocks, 256>>>(e->moves, rows, cols, e->near_moved);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    e->terrain.type = 2; e->terrain.mv = e->moves; e->terrain.near_mv = e->near_moved;
    return GRX_OK;
}

// Replaces gym.add_triangle_mesh (legged_robot.py:903-924).  The mesh the reference uploads is the structured conversion of its heightfield
// (terrain_utils.convert_heightfield_to_trimesh: one vertex per sample, two triangles per cell, steep-edge vertices shifted sideways by one
// cell).  The entry checks that the mesh IS such a conversion of `samples` — every vertex (height, and x / y within one cell of its grid
// position, shifts integral) and EVERY triangle (the two index triples of its cell) — and rejects arbitrary meshes; contacts are then
// resolved on the mesh's top surface (grid + shifts, terrain_query).
extern "C" int grx_env_set_terrain_trimesh(grx_env *e, const float *vertices, int32_t nv, const uint32_t *triangles, int32_t nt,
                                           const int16_t *samples, int32_t rows, int32_t cols, float hscale, float vscale, float border,
                                           float friction, float restitution) {
    if (!e || !vertices || !triangles || !samples || rows < 2 || cols < 2)
        return grx_set_error(GRX_E_INVALID, "grx_env_set_terrain_trimesh: bad arguments");
    if ((long long)nv != (long long)rows * cols || (long long)nt != 2ll * (rows - 1) * (cols - 1))
        return grx_set_error(GRX_E_INVALID, "grx_env_set_terrain_trimesh: not the structured conversion of the heightfield (need nv == rows*cols, nt == 2(rows-1)(cols-1))");
    std::vector<signed char> mv(2 * (size_t)nv);
    for (long long i = 0; i < (long long)nv; i++) {
        const float z = vertices[3 * i + 2], want = (float)samples[i] * vscale;
        const float fx = vertices[3 * i] / hscale - (float)(i / cols), fy = vertices[3 * i + 1] / hscale - (float)(i % cols);
        const float rx = roundf(fx), ry = roundf(fy);
        if (fabsf(z - want) > 1e-4f + 1e-6f * fabsf(want) || fabsf(rx) > 1.f || fabsf(ry) > 1.f || fabsf(fx - rx) > 1e-3f || fabsf(fy - ry) > 1e-3f)
            return grx_set_error(GRX_E_INVALID, "grx_env_set_terrain_trimesh: vertex " + std::to_string(i) + " does not match the heightfield sample grid");
        mv[2 * i] = (signed char)rx; mv[2 * i + 1] = (signed char)ry;
    }
    for (long long c = 0; c < (long long)(rows - 1) * (cols - 1); c++) {   // terrain_utils.py:333-348: (ind0, ind3, ind1) and (ind0, ind2, ind3) per cell
        const uint32_t i0 = (uint32_t)((c / (cols - 1)) * cols + c % (cols - 1)), i1 = i0 + 1, i2 = i0 + cols, i3 = i2 + 1;
        const uint32_t *t0 = triangles + 6 * c, *t1 = t0 + 3;
        if (t0[0] != i0 || t0[1] != i3 || t0[2] != i1 || t1[0] != i0 || t1[1] != i2 || t1[2] != i3)
            return grx_set_error(GRX_E_INVALID, "grx_env_set_terrain_trimesh: triangles of cell " + std::to_string(c) + " are not the structured pair");
    }
    int rc = grx_env_set_terrain_heightfield(e, samples, rows, cols, hscale, vscale, border, friction, restitution);
    if (rc) return rc;
    return finish_trimesh(e, mv.data(), 0.0, rows, cols);
}

// The same terrain built ON THE DEVICE from the sample grid alone: the steep-edge snapping of terrain_utils.py:315-328 as one kernel over the
// vertices (no 2.7 M-vertex host mesh: the reference builds 33 MB of vertices + 65 MB of indices in numpy for the default 1300 x 2100 grid).
extern "C" int grx_env_set_terrain_trimesh_hf(grx_env *e, const int16_t *samples, int32_t rows, int32_t cols, float hscale, float vscale,
                                              float border, float slope_threshold, float friction, float restitution) {
    if (!e || !samples || rows < 2 || cols < 2) return grx_set_error(GRX_E_INVALID, "grx_env_set_terrain_trimesh_hf: bad arguments");
    int rc = grx_env_set_terrain_heightfield(e, samples, rows, cols, hscale, vscale, border, friction, restitution);
    if (rc) return rc;
    return finish_trimesh(e, nullptr, (double)slope_threshold * (double)hscale / (double)vscale, rows, cols);
}

extern "C" int grx_env_set_params(grx_env *e, const float *friction, const float *restitution, const float *motor_strength,
                                  const float *base_inertial, const float *env_origins, const int32_t *terrain_levels,
                                  const int32_t *terrain_types, const float *terrain_origins, int32_t t_rows, int32_t t_cols) {
    if (!e || !friction || !restitution || !motor_strength || !base_inertial || !env_origins)
        return grx_set_error(GRX_E_INVALID, "grx_env_set_params: null argument");
    CK(cudaSetDevice(e->device));
    const size_t N = e->N;
    const LayR &L = e->lay;
    const int nd = e->nd;
    std::vector<float> c(N * L.cst_f, 0.f), r(N * L.rec_f);
    CK(cudaMemcpy(r.data(), e->rec, r.size() * 4, cudaMemcpyDeviceToHost));
    for (size_t i = 0; i < N; i++) {
        for (int j = 0; j < nd; j++) c[i * L.cst_f + L.c_motor + j] = motor_strength[i * nd + j];
        for (int j = 0; j < 10; j++) c[i * L.cst_f + L.c_bi + j] = base_inertial[i * 10 + j];
        c[i * L.cst_f + L.c_fric] = friction[i]; c[i * L.cst_f + L.c_rest] = restitution[i];
        for (int k = 0; k < 3; k++) r[i * L.rec_f + L.origin + k] = env_origins[i * 3 + k];
        int lv = terrain_levels ? terrain_levels[i] : 0, ty = terrain_types ? terrain_types[i] : 0;
        memcpy(&r[i * L.rec_f + L.tlevel], &lv, 4); memcpy(&r[i * L.rec_f + L.ttype], &ty, 4);
    }
    CK(cudaMemcpy(e->cst, c.data(), c.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(e->rec, r.data(), r.size() * 4, cudaMemcpyHostToDevice));
    if (terrain_origins && t_rows > 0 && t_cols > 0) {
        cudaFree(e->terrain_origins);
        CK(cudaMalloc((void **)&e->terrain_origins, (size_t)t_rows * t_cols * 3 * 4));
        CK(cudaMemcpy(e->terrain_origins, terrain_origins, (size_t)t_rows * t_cols * 3 * 4, cudaMemcpyHostToDevice));
        e->t_rows = t_rows; e->t_cols = t_cols;
    } else if (e->cfg.curriculum) {
        return grx_set_error(GRX_E_INVALID, "grx_env_set_params: curriculum needs terrain_origins");
    }
    e->params_set = true;
    return GRX_OK;
}

static int g_buf_device = 0;   // device ordinal stamped into the descriptors built by set_buf (set by the get_buffer entries)
static void set_buf(grx_buffer *b, void *data, int dtype, int ndim, int64_t d0, int64_t d1, int64_t d2, int64_t s0, int64_t s1, int64_t s2) {
    b->device = g_buf_device; b->own_data = 0;
    b->data = data; b->dtype = dtype; b->ndim = ndim;
    for (int i = 0; i < GRX_MAX_DIMS; i++) { b->dims[i] = 1; b->strides[i] = 1; }
    b->dims[0] = d0; b->dims[1] = d1; b->dims[2] = d2;
    b->strides[0] = s0; b->strides[1] = s1; b->strides[2] = s2;
}

extern "C" int grx_env_get_buffer(grx_env *e, const char *name, grx_buffer *b) {
    if (!e || !name || !b) return grx_set_error(GRX_E_INVALID, "grx_env_get_buffer: null argument");
    const int64_t N = e->N;
    const std::string n(name);
    g_buf_device = e->device;
    const LayR &L = e->lay;
    const int nd = e->nd;
    struct RecView { const char *name; int off, width, dtype; };
    const RecView views[] = {
        {"root_states", L.root, 13, GRX_F32}, {"dof_pos", L.dofpos, nd, GRX_F32}, {"dof_vel", L.dofvel, nd, GRX_F32},
        {"last_dof_vel", L.lastdofvel, nd, GRX_F32}, {"last_actions", L.lastact, nd, GRX_F32},
        {"last_last_actions", L.lastlastact, nd, GRX_F32}, {"commands", L.cmd, 3, GRX_F32},
        {"base_heights_offset", L.bho, 1, GRX_F32}, {"feet_air_time", L.air, NF, GRX_F32}, {"feet_land_time", L.land, NF, GRX_F32},
        {"feet_contact_last", L.clast, NF, GRX_F32}, {"episode_length", L.eplen, 1, GRX_I32}, {"terrain_levels", L.tlevel, 1, GRX_I32},
        {"terrain_types", L.ttype, 1, GRX_I32}, {"env_origins", L.origin, 3, GRX_F32}, {"episode_sums", L.sums, NREW, GRX_F32},
        {"records", 0, L.rec_f, GRX_F32}};
    for (const RecView &v : views)
        if (n == v.name) { set_buf(b, e->rec + v.off, v.dtype, 2, N, v.width, 1, L.rec_f, 1, 1); return GRX_OK; }
    if (n == "obs") { set_buf(b, e->obs, GRX_F32, 2, N, e->cfg.num_obs, 1, e->cfg.num_obs, 1, 1); return GRX_OK; }
    if (n == "pri_obs") { set_buf(b, e->pri_obs, GRX_F32, 2, N, e->cfg.num_pri_obs, 1, e->cfg.num_pri_obs, 1, 1); return GRX_OK; }
    if (n == "rew") { set_buf(b, e->rew, GRX_F32, 1, N, 1, 1, 1, 1, 1); return GRX_OK; }
    if (n == "reset") { set_buf(b, e->reset, GRX_U8, 1, N, 1, 1, 1, 1, 1); return GRX_OK; }
    if (n == "time_out") { set_buf(b, e->time_out, GRX_U8, 1, N, 1, 1, 1, 1, 1); return GRX_OK; }
    if (n == "torques") { set_buf(b, e->torques, GRX_F32, 2, N, nd, 1, nd, 1, 1); return GRX_OK; }
    if (n == "contact_forces") { set_buf(b, e->contact_forces, GRX_F32, 3, N, e->nl, 3, (int64_t)e->nl * 3, 3, 1); return GRX_OK; }
    if (n == "foot_state") { set_buf(b, e->foot_state, GRX_F32, 3, N, NF, 13, NF * 13, 13, 1); return GRX_OK; }
    if (n == "episode_accum") { set_buf(b, e->episode_accum, GRX_F32, 2, ACC_RING, ACC_W, 1, ACC_W, 1, 1); return GRX_OK; }
    if (n == "params") { set_buf(b, e->cst, GRX_F32, 2, N, L.cst_f, 1, L.cst_f, 1, 1); return GRX_OK; }
    if (n == "rigid_body_states" || n == "dof_state" || n == "episode_length_i64") {   // compat exports: allocated + switched on by the first request
        CK(cudaSetDevice(e->device));
        if (n == "rigid_body_states") {
            if (!e->rigid_body_states) { CK(cudaMalloc((void **)&e->rigid_body_states, (size_t)N * e->nl * 13 * 4)); CK(cudaMemset(e->rigid_body_states, 0, (size_t)N * e->nl * 13 * 4)); }
            set_buf(b, e->rigid_body_states, GRX_F32, 3, N, e->nl, 13, (int64_t)e->nl * 13, 13, 1);
        } else if (n == "dof_state") {
            if (!e->dof_state) { CK(cudaMalloc((void **)&e->dof_state, (size_t)N * nd * 2 * 4)); CK(cudaMemset(e->dof_state, 0, (size_t)N * nd * 2 * 4)); }
            set_buf(b, e->dof_state, GRX_F32, 3, N, nd, 2, nd * 2, 2, 1);
        } else {
            if (!e->ep_len64) { CK(cudaMalloc((void **)&e->ep_len64, (size_t)N * 8)); CK(cudaMemset(e->ep_len64, 0, (size_t)N * 8)); }
            set_buf(b, e->ep_len64, GRX_I64, 1, N, 1, 1, 1, 1, 1);
        }
        return GRX_OK;
    }
    if (n == "terrain_moves") {   // structured trimesh: per-vertex shifts [rows, cols, 2] int8 (exported as u8 bit patterns)
        if (!e->moves) return grx_set_error(GRX_E_STATE, "grx_env_get_buffer: no trimesh terrain set");
        set_buf(b, e->moves, GRX_U8, 3, e->terrain.rows, e->terrain.cols, 2, (int64_t)e->terrain.cols * 2, 2, 1); return GRX_OK;
    }
    if (n == "active_sig") {
        if (!e->active_sig) return grx_set_error(GRX_E_STATE, "grx_env_get_buffer: call grx_env_debug_active_sig(env, 1) first");
        set_buf(b, e->active_sig, GRX_U64, 2, N, SIG_STRIDE, 1, SIG_STRIDE, 1, 1); return GRX_OK;
    }
    return grx_set_error(GRX_E_NOTFOUND, "grx_env_get_buffer: unknown buffer '" + n + "'");
}

static EnvArgs make_args(grx_env *e, const float *d_actions, const float *d_uniform, float delay, int push, uint64_t step_index) {
    EnvArgs A;
    memset(&A, 0, sizeof(A));
    A.rec = e->rec; A.cst = e->cst; A.model = e->dmodel; A.terrain = e->terrain; A.terrain_origins = e->terrain_origins;
    A.t_rows = e->t_rows; A.t_cols = e->t_cols; A.N = e->N; A.actions = d_actions; A.U = d_uniform; A.delay = delay; A.push = push;
    A.step_index = step_index; A.obs = e->obs; A.pri_obs = e->pri_obs; A.rew = e->rew; A.torques = e->torques;
    A.contact_forces = e->contact_forces; A.foot_state = e->foot_state;
    A.episode_accum = e->episode_accum + (size_t)(e->launches % ACC_RING) * ACC_W;
    A.episode_accum_next = e->episode_accum + (size_t)((e->launches + 1) % ACC_RING) * ACC_W;
    A.reset = e->reset; A.time_out = e->time_out;
    A.dbg_sig = e->active_sig; A.dbg_sig_stride = SIG_STRIDE;
    A.rigid_body_states = e->rigid_body_states; A.dof_state = e->dof_state; A.ep_len64 = e->ep_len64;
    A.link_body = e->d_link_body; A.link_pos = e->d_link_pos; A.link_rot = e->d_link_rot;
    return A;
}

extern "C" int grx_env_step(grx_env *e, const float *d_actions, const float *d_uniform, float delay, int32_t push,
                            uint64_t step_index, void *stream) {
    if (!e || !d_actions) return grx_set_error(GRX_E_INVALID, "grx_env_step: null argument");
    if (!e->params_set) return grx_set_error(GRX_E_STATE, "grx_env_step: call grx_env_set_params first");
    EnvArgs A = make_args(e, d_actions, d_uniform, delay, push, step_index);
    launch_env_step<true>(e, A, e->cfg, (cudaStream_t)stream);
    CK(cudaGetLastError());
    e->launches++;
    return GRX_OK;
}

extern "C" int grx_env_post_physics(grx_env *e, const float *d_actions, const float *d_uniform, const grx_injected_physics *inj,
                                    int32_t push, uint64_t step_index, void *stream) {
    if (!e || !d_actions || !inj) return grx_set_error(GRX_E_INVALID, "grx_env_post_physics: null argument");
    if (!e->params_set) return grx_set_error(GRX_E_STATE, "grx_env_post_physics: call grx_env_set_params first");
    EnvArgs A = make_args(e, d_actions, d_uniform, 0.f, push, step_index);
    A.inj = *inj;
    launch_env_step<false>(e, A, e->cfg, (cudaStream_t)stream);
    CK(cudaGetLastError());
    e->launches++;
    return GRX_OK;
}

extern "C" int grx_env_reset_idx(grx_env *e, const int32_t *d_ids, int32_t n, const float *d_uniform, int32_t curriculum_active,
                                 uint64_t step_index, void *stream) {
    if (!e || n < 0 || (d_ids == nullptr && n != e->N)) return grx_set_error(GRX_E_INVALID, "grx_env_reset_idx: bad arguments (ids == NULL needs n == num_envs)");
    if (!e->params_set) return grx_set_error(GRX_E_STATE, "grx_env_reset_idx: call grx_env_set_params first");
    if (n == 0) return GRX_OK;                                                        // LR:387-388
    EnvArgs A = make_args(e, nullptr, d_uniform, 0.f, 0, step_index);
    if (e->g) {
        int rc = grx::envg_launch_reset(e->g, A, e->cfg, e->lay, d_ids, n, curriculum_active, (cudaStream_t)stream);
        if (rc) return rc;
    } else {
        const int grid = (n + WARPS_PER_CTA - 1) / WARPS_PER_CTA;
        grx_count_launch();
        env_reset_kernel<<<grid, WARPS_PER_CTA * 32, env_smem_bytes(WARPS_PER_CTA), (cudaStream_t)stream>>>(A, e->cfg, d_ids, n, curriculum_active);
    }
    CK(cudaGetLastError());
    e->launches++;
    return GRX_OK;
}

extern "C" int64_t grx_env_accum_slot(grx_env *e) { return e ? (int64_t)((e->launches + ACC_RING - 1) % ACC_RING) : -1; }

extern "C" int grx_env_step_host(grx_env *e, const float *h_actions, float delay, int32_t push, uint64_t step_index,
                                 float *h_obs, float *h_pri_obs, float *h_rew, uint8_t *h_reset, void *stream) {
    if (!e || !h_actions) return grx_set_error(GRX_E_INVALID, "grx_env_step_host: null argument");
    cudaStream_t st = (cudaStream_t)stream;
    const size_t N = e->N;
    CK(cudaMemcpyAsync(e->actions_stage, h_actions, N * e->nd * 4, cudaMemcpyHostToDevice, st));
    int rc = grx_env_step(e, e->actions_stage, nullptr, delay, push, step_index, stream);
    if (rc) return rc;
    if (h_obs) CK(cudaMemcpyAsync(h_obs, e->obs, N * e->cfg.num_obs * 4, cudaMemcpyDeviceToHost, st));
    if (h_pri_obs) CK(cudaMemcpyAsync(h_pri_obs, e->pri_obs, N * e->cfg.num_pri_obs * 4, cudaMemcpyDeviceToHost, st));
    if (h_rew) CK(cudaMemcpyAsync(h_rew, e->rew, N * 4, cudaMemcpyDeviceToHost, st));
    if (h_reset) CK(cudaMemcpyAsync(h_reset, e->reset, N, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return GRX_OK;
}

// Debug / parity: switch the per-substep active-set signature export on (allocates [N, 16] u64, buffer name "active_sig") or off.
extern "C" int grx_env_debug_active_sig(grx_env *e, int32_t enable) {
    if (!e) return grx_set_error(GRX_E_INVALID, "null env");
    CK(cudaSetDevice(e->device));
    if (enable && e->cfg.decimation > SIG_STRIDE) return grx_set_error(GRX_E_INVALID, "grx_env_debug_active_sig: decimation > 16");
    if (enable && !e->active_sig) {
        CK(cudaMalloc((void **)&e->active_sig, (size_t)e->N * SIG_STRIDE * 8));
        CK(cudaMemset(e->active_sig, 0, (size_t)e->N * SIG_STRIDE * 8));
    } else if (!enable && e->active_sig) {
        CK(cudaDeviceSynchronize());
        cudaFree(e->active_sig); e->active_sig = nullptr;
    }
    return GRX_OK;
}

extern "C" int grx_env_debug_dynamics(grx_env *e, int32_t index, float *h_M, float *h_h) {
    if (!e || !h_M || !h_h || index < 0 || index >= e->N) return grx_set_error(GRX_E_INVALID, "grx_env_debug_dynamics: bad argument");
    if (!e->params_set) return grx_set_error(GRX_E_STATE, "grx_env_debug_dynamics: call grx_env_set_params first");
    CK(cudaSetDevice(e->device));
    float *tmp_rec = nullptr, *dM = nullptr, *dh = nullptr, *dact = nullptr;
    const int nv = e->nd + 6;   // == NV for the lower-limb tree
    CK(cudaMalloc((void **)&tmp_rec, (size_t)e->N * e->lay.rec_f * 4));
    CK(cudaMalloc((void **)&dM, nv * nv * 4)); CK(cudaMalloc((void **)&dh, nv * 4));
    CK(cudaMalloc((void **)&dact, (size_t)e->N * e->nd * 4)); CK(cudaMemset(dact, 0, (size_t)e->N * e->nd * 4));
    CK(cudaMemcpy(tmp_rec, e->rec, (size_t)e->N * e->lay.rec_f * 4, cudaMemcpyDeviceToDevice));
    EnvArgs A = make_args(e, dact, nullptr, 0.f, 0, 0);
    A.rec = tmp_rec; A.dbg_M = dM; A.dbg_h = dh; A.dbg_index = index;
    grx_task_cfg c = e->cfg;
    c.decimation = 1;
    launch_env_step<true>(e, A, c, nullptr);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(h_M, dM, nv * nv * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(h_h, dh, nv * 4, cudaMemcpyDeviceToHost));
    cudaFree(tmp_rec); cudaFree(dM); cudaFree(dh); cudaFree(dact);
    return GRX_OK;
}

// ABI self-check for foreign-function bindings: sizes of the public structs
extern "C" int grx_abi_sizes(int32_t *out, int32_t n) {
    const int32_t v[5] = {(int32_t)sizeof(grx_buffer), (int32_t)sizeof(grx_model_desc), (int32_t)sizeof(grx_task_cfg),
                          (int32_t)sizeof(grx_injected_physics), (int32_t)sizeof(grx_ppo_cfg)};
    for (int i = 0; i < n && i < 5; i++) out[i] = v[i];
    return 5;
}

// Robot self-collision (legged_robot_config.py:121 self_collisions = 0 = enabled; create_actor(..., collision_filter = 0), legged_robot.py:1022-1028):
// candidate sphere pairs of a generic-topology env.  The specialised lower-limb kernel carries no self-contact rows -> GRX_E_INVALID there
// (create the env with GRX_ENV_GENERIC=1 to run the lower-limb model with self-collision on the generic kernels).
extern "C" int grx_env_set_self_collision(grx_env *e, const int32_t *pairs, int32_t npairs, int32_t max_self_contacts) {
    if (!e) return grx_set_error(GRX_E_INVALID, "null env");
    if (!e->g) return grx_set_error(GRX_E_INVALID, "grx_env_set_self_collision: the specialised lower-limb kernel has no self-contact rows (GRX_ENV_GENERIC=1 selects the generic kernels)");
    return grx::envg_set_self_collision(e->g, pairs, npairs, max_self_contacts);
}

// Shape facts of an env a binding needs before it allocates: 0 = uniform draws per env and step (the K of grx_b200/rng_layout.py: 68 for 10 DOF,
// 156 for 32), 1 = floats per state record, 2 = floats per parameter record, 3 = actuated DOF, 4 = 1 if the generic-topology kernels run it
extern "C" int64_t grx_env_info(grx_env *e, int32_t what) {
    if (!e) return -1;
    switch (what) {
        case 0: return e->lay.rng_k;
        case 1: return e->lay.rec_f;
        case 2: return e->lay.cst_f;
        case 3: return e->nd;
        case 4: return e->g ? 1 : 0;
        default: return -1;
    }
}
