// grx_ppo.cu — the rsl_rl side of the hot path for sm_100a (C ABI in include/grx_b200.h):
//   ActorCriticMLP forward (rollout + update), GAE reverse scan (warp-shuffle linear-recurrence scan), the PPO
//   clipped-surrogate / clipped-value / entropy loss with its hand-derived backward, global-norm clip + Adam with the
//   adaptive-KL learning rate decided on the device (the reference syncs with .item() three times per minibatch,
//   ppo.py:264, 308-309).  Reference arithmetic: rsl_rl/rsl_rl/algorithms/ppo.py:144-321,
//   storage/base_storage.py:80-141, storage/rollout_storage.py:63-112, modules/actor_critic_mlp.py:160-231,
//   modules/mlp.py:7-42.  oracle/ppo_oracle.py is the CPU statement of the same equations.
//
// Dense layers: gemm_kernel (fp32 SIMT, any shape) or, for the 128-aligned hidden layers when
// cfg.use_tensor_cores != 0, the tcgen05 kind::tf32 kernels in grx_gemm_tc.cuh.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <string.h>

#include <string>
#include <vector>

#include "grx_b200.h"
#include "grx_count.h"
#include "grx_gemm_tc.cuh"
#include "grx_mlp_chain.cuh"

int grx_set_error(int code, const std::string &msg);   // grx_env.cu

#define CK(call)                                                                                             \
    do {                                                                                                     \
        cudaError_t err__ = (call);                                                                          \
        if (err__ != cudaSuccess)                                                                            \
            return grx_set_error(GRX_E_CUDA, std::string(#call) + ": " + cudaGetErrorString(err__));        \
    } while (0)

namespace {

constexpr int MAXA = 32;          // max action dim
constexpr int TAIL = 8;           // reduce_buf tail: [kl_sum, count, surrogate_sum, value_loss_sum, 0...]
constexpr float LOG_SQRT_2PI = 0.91893853320467274178f;
constexpr unsigned FULL = 0xffffffffu;

__device__ __forceinline__ float elu(float x) { return x > 0.f ? x : expm1f(x); }          // nn.ELU(alpha=1)
__device__ __forceinline__ float elu_grad_from_out(float h) { return h > 0.f ? 1.f : h + 1.f; }   // ELU'(z) = e^z = h + 1 for z <= 0

// =========================================================================================================
// fp32 SIMT GEMM, C[m,n] (+)= sum_k A(m,k) B(k,n), 64x64x16 tiles, 256 threads, 4x4 outputs per thread.
//   A_KC: A(m,k) = A[m*lda + k] (contraction index contiguous) else A[k*lda + m]
//   B_KC: B(k,n) = B[n*ldb + k]                                else B[k*ldb + n]
// Epilogues: 0 C = acc + bias[n]; 1 C = elu(acc + bias[n]); 2 C = acc * ELU'(aux[m,n]) ; 3 split-K: atomicAdd(C, acc) and
// (n-tile 0 only) bias_out[m] += sum_k A(m,k)   [dW = dY^T H with db = column sums of dY in the same pass]
// =========================================================================================================
struct GemmArgs {
    const float *A, *B;
    float *C;
    const float *bias;   // epi 0/1
    const float *aux;    // epi 2 (same layout as C)
    float *bias_out;     // epi 3
    int M, N, K, lda, ldb, ldc, kchunk;
};

constexpr int BM = 64, BN = 64, BK = 16;

template <bool A_KC, bool B_KC, int EPI>
__global__ void __launch_bounds__(256) gemm_kernel(const GemmArgs g) {
    __shared__ __align__(16) float As[BK][BM + 4];
    __shared__ __align__(16) float Bs[BK][BN + 4];
    const int t = threadIdx.x, tx = t & 15, ty = t >> 4;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    const int kbeg = blockIdx.z * g.kchunk, kend = min(g.K, kbeg + g.kchunk);
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j] = 0.f;
    float bsum = 0.f;
    for (int k0 = kbeg; k0 < kend; k0 += BK) {
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const int idx = t + i * 256;
            {
                const int m = A_KC ? idx / BK : idx % BM, k = A_KC ? idx % BK : idx / BM;
                const int gm = m0 + m, gk = k0 + k;
                float v = 0.f;
                if (gm < g.M && gk < kend) v = A_KC ? g.A[(size_t)gm * g.lda + gk] : g.A[(size_t)gk * g.lda + gm];
                As[k][m] = v;
            }
            {
                const int n = B_KC ? idx / BK : idx % BN, k = B_KC ? idx % BK : idx / BN;
                const int gn = n0 + n, gk = k0 + k;
                float v = 0.f;
                if (gn < g.N && gk < kend) v = B_KC ? g.B[(size_t)gn * g.ldb + gk] : g.B[(size_t)gk * g.ldb + gn];
                Bs[k][n] = v;
            }
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < BK; k++) {
            const float4 a = *reinterpret_cast<const float4 *>(&As[k][ty * 4]);
            const float4 b = *reinterpret_cast<const float4 *>(&Bs[k][tx * 4]);
            const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 4; i++)
#pragma unroll
                for (int j = 0; j < 4; j++) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        if (EPI == 3 && blockIdx.x == 0 && t < BM) {
#pragma unroll
            for (int k = 0; k < BK; k++) bsum += As[k][t];
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const int m = m0 + ty * 4 + i;
        if (m >= g.M) continue;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int n = n0 + tx * 4 + j;
            if (n >= g.N) continue;
            const size_t o = (size_t)m * g.ldc + n;
            if (EPI == 0) g.C[o] = acc[i][j] + g.bias[n];
            else if (EPI == 1) g.C[o] = elu(acc[i][j] + g.bias[n]);
            else if (EPI == 2) g.C[o] = acc[i][j] * elu_grad_from_out(g.aux[o]);
            else atomicAdd(&g.C[o], acc[i][j]);
        }
    }
    if (EPI == 3 && g.bias_out != nullptr && blockIdx.x == 0 && t < BM && m0 + t < g.M) atomicAdd(&g.bias_out[m0 + t], bsum);
}

template <bool A_KC, bool B_KC, int EPI>
void launch_gemm(const GemmArgs &g, int splits, cudaStream_t st) {
    GemmArgs a = g;
    a.kchunk = ((g.K + splits - 1) / splits + BK - 1) / BK * BK;
    const int z = (g.K + a.kchunk - 1) / a.kchunk;
    dim3 grid((g.N + BN - 1) / BN, (g.M + BM - 1) / BM, z);
    grx_count_launch();
    gemm_kernel<A_KC, B_KC, EPI><<<grid, 256, 0, st>>>(a);
}

// column sums of a row-major [M, N] matrix into out[N] (+=): bias gradients on the tensor-core path
__global__ void __launch_bounds__(256) colsum_kernel(const float *X, int M, int N, int ld, int rows_per_block, float *out) {
    const int n = blockIdx.x * 32 + (threadIdx.x & 31), w = threadIdx.x >> 5;
    const int r0 = blockIdx.y * rows_per_block, r1 = min(M, r0 + rows_per_block);
    float s = 0.f;
    if (n < N)
        for (int r = r0 + w; r < r1; r += 8) s += X[(size_t)r * ld + n];
    __shared__ float red[8][33];
    red[w][threadIdx.x & 31] = s;
    __syncthreads();
    if (w == 0 && n < N) {
        float t = 0.f;
#pragma unroll
        for (int k = 0; k < 8; k++) t += red[k][threadIdx.x];
        atomicAdd(&out[n], t);
    }
}

thread_local cudaError_t g_launch_err = cudaSuccess;   // first failed tensor-core launch (reported by the C entry points)

// Dense-layer dispatch for a GROUP of problems with the same operand layout and epilogue (actor + critic layer of the same
// depth; all weight gradients of a backward pass): ONE persistent tcgen05 TF32 launch when enabled and every shape / alignment
// allows, else the fp32 SIMT kernel per problem.
template <bool A_KC, bool B_KC, int EPI>
void dense_group(const GemmArgs *gs, const int *splits, int np, bool use_tc, cudaStream_t st) {
    // members the tensor-core kernel takes go into ONE grouped launch; the others (narrow output heads of a non-registered policy, unaligned
    // shapes) run on the fp32 SIMT kernel one by one — a single narrow member must not drag the whole group off the tensor cores
    bool on_tc[tc::MAXP > 8 ? tc::MAXP : 8] = {false};
    if (use_tc && np <= tc::MAXP) {
        tc::Problem ps[tc::MAXP];
        int sp_tc[tc::MAXP], ntc = 0;
        for (int i = 0; i < np; i++) {
            const GemmArgs &g = gs[i];
            tc::Problem a;
            a.A = g.A; a.B = g.B; a.C = g.C; a.bias = g.bias; a.aux = g.aux; a.colsum = EPI == 2 ? g.bias_out : nullptr;
            a.M = g.M; a.N = g.N; a.K = g.K; a.lda = g.lda; a.ldb = g.ldb; a.ldc = g.ldc;
            bool ok = tc::supported<A_KC, B_KC>(a);
            if (EPI == 3) ok = ok && g.M >= 64 && g.N >= 32;   // tiny weight gradients (the output heads) stay on the SIMT kernel
            if (EPI == 2) ok = ok && g.K >= 64;
            if (!ok) continue;
            on_tc[i] = true;
            sp_tc[ntc] = splits ? splits[i] : 1;
            ps[ntc++] = a;
            if (EPI == 3 && g.bias_out) {
                const int rpb = 512;
                grx_count_launch();
                colsum_kernel<<<dim3((g.M + 31) / 32, (g.K + rpb - 1) / rpb), 256, 0, st>>>(g.A, g.K, g.M, g.lda, rpb, g.bias_out);   // A = dY [rows, out]
            }
        }
        if (ntc > 0) {
            const cudaError_t e = tc::launch_group<A_KC, B_KC, EPI>(ps, ntc, splits ? sp_tc : nullptr, st);
            if (e != cudaSuccess && g_launch_err == cudaSuccess) g_launch_err = e;
        }
    }
    for (int i = 0; i < np; i++) {
        if (on_tc[i]) continue;
        const GemmArgs &g = gs[i];
        int sp = splits ? splits[i] : 1;
        if (EPI == 3) {   // SIMT split-K: enough 64 x 64 tiles to fill the GPU
            const int tiles = ((g.M + BM - 1) / BM) * ((g.N + BN - 1) / BN);
            sp = (2 * 148 + tiles - 1) / tiles;
            const int maxs = (g.K + 255) / 256;
            if (sp > maxs) sp = maxs;
            if (sp < 1) sp = 1;
        }
        launch_gemm<A_KC, B_KC, EPI>(g, sp, st);
        if (EPI == 2 && g.bias_out) {   // column sums of the produced gradient = bias gradient of the layer below
            const int rpb = 512;
            grx_count_launch();
            colsum_kernel<<<dim3((g.N + 31) / 32, (g.M + rpb - 1) / rpb), 256, 0, st>>>(g.C, g.M, g.N, g.ldc, rpb, g.bias_out);
        }
    }
}
template <bool A_KC, bool B_KC, int EPI>
void dense(const GemmArgs &g, int splits, bool use_tc, cudaStream_t st) { dense_group<A_KC, B_KC, EPI>(&g, &splits, 1, use_tc, st); }

// A CHAIN of dependent layers (gs[i] reads the output of gs[deps[i]], deps[i] < i or -1) as ONE layer-pipelined tcgen05 launch
// (tc::launch_pipe).  Returns false when the chain does not qualify (then the caller launches layer by layer).
template <bool A_KC, bool B_KC, int EPI>
bool dense_pipe(const GemmArgs *gs, const int *deps, int np, bool use_tc, int *sync, int *err, cudaStream_t st) {
    if (!use_tc || sync == nullptr || np > tc::MAXP) return false;
    tc::Problem ps[tc::MAXP];
    for (int i = 0; i < np; i++) {
        const GemmArgs &g = gs[i];
        tc::Problem &a = ps[i];
        a.A = g.A; a.B = g.B; a.C = g.C; a.bias = g.bias; a.aux = g.aux; a.colsum = EPI == 2 ? g.bias_out : nullptr;
        a.M = g.M; a.N = g.N; a.K = g.K; a.lda = g.lda; a.ldb = g.ldb; a.ldc = g.ldc;
    }
    if (!tc::pipe_supported<A_KC, B_KC, EPI>(ps, np, deps)) return false;
    const cudaError_t e = tc::launch_pipe<A_KC, B_KC, EPI>(ps, np, deps, sync, err, st);
    if (e != cudaSuccess && g_launch_err == cudaSuccess) g_launch_err = e;
    return true;
}

// =========================================================================================================
// Philox4x32-10 + Box-Muller for the rollout's Normal.sample() in fast mode (actor_critic_mlp.py:192-194)
// =========================================================================================================
__device__ __forceinline__ void philox4(uint32_t k0, uint32_t k1, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t *out) {
#pragma unroll
    for (int r = 0; r < 10; r++) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// PPO.act tail (ppo.py:150-171, base_storage.py:80-100): a = mu + sigma * eps, log-prob, and the transition's row of the
// rollout storage (obs, critic_obs, actions, values, log-prob, mu, sigma).  One thread per env.
struct ActArgs {
    const float *obs, *critic_obs, *mu, *value, *std, *eps;
    float *actions_out;
    float *s_obs, *s_cobs, *s_act, *s_val, *s_logp, *s_mu, *s_sigma;   // storage rows of step t
    int N, O, P, A;
    uint64_t seed, step_index;
    int env_id_offset;
};
__global__ void act_sample_store_kernel(const ActArgs a) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n < a.N) {
        float lp = 0.f;
        for (int j0 = 0; j0 < a.A; j0 += 4) {
            float z[4];
            if (a.eps) {
                for (int r = 0; r < 4; r++) z[r] = j0 + r < a.A ? a.eps[(size_t)n * a.A + j0 + r] : 0.f;
            } else {
                uint32_t rnd[4];
                philox4((uint32_t)a.seed ^ 0x2545F491u, (uint32_t)(a.seed >> 32), (uint32_t)(a.env_id_offset + n), (uint32_t)a.step_index,
                        (uint32_t)(a.step_index >> 32), (uint32_t)(j0 >> 2), rnd);
                for (int r = 0; r < 4; r += 2) {   // Box-Muller on (0,1] x [0,1)
                    const float u1 = ((float)(rnd[r] >> 8) + 1.0f) * (1.0f / 16777216.0f), u2 = (float)(rnd[r + 1] >> 8) * (1.0f / 16777216.0f);
                    const float rad = sqrtf(-2.0f * logf(u1));
                    float sn, cs;
                    sincospif(2.0f * u2, &sn, &cs);
                    z[r] = rad * cs; z[r + 1] = rad * sn;
                }
            }
            for (int r = 0; r < 4 && j0 + r < a.A; r++) {
                const int j = j0 + r;
                const float mu = a.mu[(size_t)n * a.A + j], sg = a.std[j];                          // sigma = mu * 0 + std (ACM:179-181)
                const float act = mu + sg * z[r];
                const float d = act - mu;
                lp += -(d * d) / (2.f * sg * sg) - logf(sg) - LOG_SQRT_2PI;                           // ACM:205
                a.actions_out[(size_t)n * a.A + j] = act;
                a.s_act[(size_t)n * a.A + j] = act; a.s_mu[(size_t)n * a.A + j] = mu; a.s_sigma[(size_t)n * a.A + j] = sg;
            }
        }
        a.s_logp[n] = lp;
        a.s_val[n] = a.value[n];
    }
    // coalesced copies of the observation rows into the storage (grid-stride over the flat arrays)
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (size_t)gridDim.x * blockDim.x;
    for (size_t i = tid; i < (size_t)a.N * a.O; i += nth) a.s_obs[i] = a.obs[i];
    for (size_t i = tid; i < (size_t)a.N * a.P; i += nth) a.s_cobs[i] = a.critic_obs[i];
}

// Rollout fast path for the registered policy (A == NA, last hidden width 128): the two output heads (mu = W3a h3a + b, V = W3c h3c + b)
// folded into the sampling / storage kernel — one warp per env, lane = 4 hidden columns, the 11 dot products reduced by the same
// 16-shuffle halving butterfly as ppo_heads_kernel, lane j owns action j.  Saves the two narrow SIMT GEMM launches per policy step.
struct ActHeadsArgs {
    const float *h3a, *h3c, *W3a, *b3a, *W3c, *b3c;
    ActArgs a;
};
template <int NA>
__global__ void __launch_bounds__(256) act_heads_kernel(const ActHeadsArgs q) {
    constexpr int H = 128;
    const ActArgs &a = q.a;
    const int lane = threadIdx.x & 31, warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
    const int c0 = lane * 4;
    const bool own = lane < NA;
    float4 wa[NA];   // parameters: not written by the forward chain in front of this kernel -> loaded before the dependency wait (see ppo_heads_kernel)
#pragma unroll
    for (int j = 0; j < NA; j++) wa[j] = *reinterpret_cast<const float4 *>(q.W3a + j * H + c0);
    const float4 wc = *reinterpret_cast<const float4 *>(q.W3c + c0);
    const float sg = own ? a.std[lane] : 1.f, ba = own ? q.b3a[lane] : 0.f, bc = q.b3c[0];
    tc::pdl_wait();
    for (int n = warp; n < a.N; n += nwarps) {
        const float4 ha = *reinterpret_cast<const float4 *>(q.h3a + (size_t)n * H + c0);
        const float4 hc = *reinterpret_cast<const float4 *>(q.h3c + (size_t)n * H + c0);
        float t[16];
#pragma unroll
        for (int j = 0; j < 16; j++) t[j] = 0.f;
#pragma unroll
        for (int j = 0; j < NA; j++) t[j] = ha.x * wa[j].x + ha.y * wa[j].y + ha.z * wa[j].z + ha.w * wa[j].w;
        t[NA] = hc.x * wc.x + hc.y * wc.y + hc.z * wc.z + hc.w * wc.w;
#pragma unroll
        for (int w = 8, bit = 8; w >= 1; w >>= 1, bit >>= 1) {
            const bool up = (lane & bit) != 0;
#pragma unroll
            for (int i = 0; i < w; i++) {
                const float send = up ? t[i] : t[i + w], keep = up ? t[i + w] : t[i];
                t[i] = keep + __shfl_xor_sync(FULL, send, bit);
            }
        }
        t[0] += __shfl_xor_sync(FULL, t[0], 16);
        const float mu = t[0] + ba;                              // lanes < NA
        const float v = __shfl_sync(FULL, t[0], NA) + bc;
        float z = 0.f;
        if (own) {
            if (a.eps) z = a.eps[(size_t)n * NA + lane];
            else {   // same counter / pairing as act_sample_store_kernel: element j uses draw (j >> 2), Box-Muller pair (j & 2)
                uint32_t rnd[4];
                philox4((uint32_t)a.seed ^ 0x2545F491u, (uint32_t)(a.seed >> 32), (uint32_t)(a.env_id_offset + n), (uint32_t)a.step_index,
                        (uint32_t)(a.step_index >> 32), (uint32_t)(lane >> 2), rnd);
                const int r = lane & 2;
                const float u1 = ((float)(rnd[r] >> 8) + 1.0f) * (1.0f / 16777216.0f), u2 = (float)(rnd[r + 1] >> 8) * (1.0f / 16777216.0f);
                const float rad = sqrtf(-2.0f * logf(u1));
                float sn, cs;
                sincospif(2.0f * u2, &sn, &cs);
                z = (lane & 1) ? rad * sn : rad * cs;
            }
        }
        const float act = mu + sg * z, d = act - mu;
        float lp = own ? -(d * d) / (2.f * sg * sg) - logf(sg) - LOG_SQRT_2PI : 0.f;      // ACM:205
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) lp += __shfl_xor_sync(FULL, lp, o);
        if (own) {
            const size_t o = (size_t)n * NA + lane;
            a.actions_out[o] = act; a.s_act[o] = act; a.s_mu[o] = mu; a.s_sigma[o] = sg;
        }
        if (lane == 0) { a.s_logp[n] = lp; a.s_val[n] = v; }
    }
    // coalesced copies of the observation rows into the storage (grid-stride over the flat arrays)
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (size_t)gridDim.x * blockDim.x;
    for (size_t i = tid; i < (size_t)a.N * a.O; i += nth) a.s_obs[i] = a.obs[i];
    for (size_t i = tid; i < (size_t)a.N * a.P; i += nth) a.s_cobs[i] = a.critic_obs[i];
}

// PPO.process_env_step (ppo.py:186-194): r += gamma * V * time_out; store rewards / dones of step t
__global__ void process_env_step_kernel(const float *rew, const uint8_t *dones, const uint8_t *time_outs, const float *values, float gamma,
                                        float *s_rew, uint8_t *s_done, int N) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    float r = rew[n];
    if (time_outs) r += gamma * (values[n] * (time_outs[n] ? 1.f : 0.f));
    s_rew[n] = r;
    s_done[n] = dones[n] ? 1 : 0;
}

// =========================================================================================================
// GAE (base_storage.py:120-141).  A_t = delta_t + c_t A_{t+1} is a linear recurrence; composing the affine maps
// f_t(x) = delta_t + c_t x is associative, so the reverse scan over T runs as a warp-shuffle scan: one warp per env,
// lane l owns the time steps of chunk l (T / 32 consecutive steps), tiles of 32 envs are transposed through shared
// memory so that HBM reads / writes stay coalesced along the env axis.  Also accumulates sum / sum of squares of the
// raw advantages (double) for the global normalisation.
// =========================================================================================================
constexpr int GAE_TMAX = 256;
__global__ void __launch_bounds__(1024) gae_kernel(const float *rew, const uint8_t *dones, const float *values, const float *last_values,
                                                   float gamma, float lam, float *returns, float *adv, double *moments, int T, int N) {
    extern __shared__ float gae_smem[];   // 3 x [T][33]
    float (*s_d)[33] = reinterpret_cast<float (*)[33]>(gae_smem);
    float (*s_c)[33] = s_d + T;
    float (*s_v)[33] = s_c + T;
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31, n0 = blockIdx.x * 32;
    const int nw = blockDim.x >> 5;
    for (int t = w; t < T; t += nw) {   // coalesced along n: delta_t and c_t per (t, n)
        const int n = n0 + lane;
        if (n < N) {
            const float v = values[(size_t)t * N + n];
            const float nv = t == T - 1 ? last_values[n] : values[(size_t)(t + 1) * N + n];
            const float nt = 1.0f - (dones[(size_t)t * N + n] ? 1.f : 0.f);
            s_d[t][lane] = rew[(size_t)t * N + n] + nt * gamma * nv - v;
            s_c[t][lane] = nt * gamma * lam;
            s_v[t][lane] = v;
        }
    }
    __syncthreads();
    double sum = 0.0, sq = 0.0;
    for (int e = w; e < 32; e += nw) {   // warp e scans env n0 + e; lane l owns steps [l*per, (l+1)*per)
        if (n0 + e >= N) continue;
        const int per = (T + 31) / 32, tb = lane * per, te = min(T, tb + per);
        // compose the chunk's maps from its last step backwards: A_tb = D + C * A_te
        float D = 0.f, Cc = 1.f;
        for (int t = te - 1; t >= tb; t--) { D = s_d[t][e] + s_c[t][e] * D; Cc = s_c[t][e] * Cc; }
        // exclusive reverse scan over lanes: incoming A for lane l = A_{te(l)} = composition of lanes l+1.. applied to 0
        float sd = D, sc = Cc;   // inclusive suffix composition
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const float od = __shfl_down_sync(FULL, sd, o), oc = __shfl_down_sync(FULL, sc, o);
            if (lane + o < 32) { sd = sd + sc * od; sc = sc * oc; }
        }
        float a = __shfl_down_sync(FULL, sd, 1);
        if (lane == 31) a = 0.f;
        for (int t = te - 1; t >= tb; t--) {
            a = s_d[t][e] + s_c[t][e] * a;
            s_d[t][e] = a;   // raw advantage; returns = a + v
            sum += (double)a; sq += (double)a * (double)a;
        }
    }
    __syncthreads();
    for (int t = w; t < T; t += nw) {
        const int n = n0 + lane;
        if (n < N) {
            const float a = s_d[t][lane];
            returns[(size_t)t * N + n] = a + s_v[t][lane];
            adv[(size_t)t * N + n] = (a + s_v[t][lane]) - s_v[t][lane];   // advantages = returns - values (BS:140)
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { sum += __shfl_xor_sync(FULL, sum, o); sq += __shfl_xor_sync(FULL, sq, o); }
    if (lane == 0) { atomicAdd(&moments[0], sum); atomicAdd(&moments[1], sq); }
    if (threadIdx.x == 0 && blockIdx.x == 0) atomicAdd(&moments[2], (double)T * (double)N);
}
// (A - mean) / (unbiased std + 1e-8) over all T*N (BS:141), moments = [sum, sum of squares, count] (all-reduced by the caller when sharded)
__global__ void normalize_adv_kernel(float *adv, const double *moments, size_t n) {
    const double cnt = moments[2], mean = moments[0] / cnt;
    const double var = fmax((moments[1] - cnt * mean * mean) / (cnt - 1.0), 0.0);
    const float m = (float)mean, inv = 1.0f / ((float)sqrt(var) + 1e-8f);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) adv[i] = (adv[i] - m) * inv;
}

// =========================================================================================================
// Minibatch gather (rollout_storage.py:77-112): rows d_indices[mb*B + r] of the flat [T*N, .] storage
// =========================================================================================================
struct GatherArgs {
    const int64_t *indices;
    const int *mb_counter;   // device-side minibatch index (so the same CUDA graph replays for every minibatch), or NULL
    int mb_off;              // added to the device counter: 1 when the gather is PREFETCHED during the previous minibatch
    int mb, nmb, B, O, P, A, Opad, Ppad;
    const float *s_obs, *s_cobs, *s_act, *s_val, *s_ret, *s_adv, *s_logp, *s_mu, *s_sigma;
    float *xa, *xc, *act, *val, *ret, *adv, *logp, *mu, *sigma;
    float4 *zero;        // gradient / reduction buffer of the minibatch, cleared here (saves a memset node between apply and gather)
    int nzero4;
};
__global__ void __launch_bounds__(256) gather_kernel(const GatherArgs g) {
    // flat element-parallel copy (a warp-per-row loop leaves the row loads of one warp serialised behind its index load)
    tc::pdl_wait();                // launched as a programmatic dependent of the previous minibatch's apply kernel
    tc::pdl_launch_dependents();
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < g.nzero4; i += gridDim.x * blockDim.x) g.zero[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    const int mb = g.mb_counter ? ((*g.mb_counter + g.mb_off) % g.nmb) : g.mb;
    const int64_t *idx = g.indices + (size_t)mb * g.B;
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (size_t)gridDim.x * blockDim.x;
    if ((g.P & 3) == 0) {   // critic rows are 16-byte aligned in the storage and in the staging buffer: float4 copies
        const int P4 = g.P >> 2, L4 = g.Ppad >> 2;
        const size_t tot = (size_t)g.B * P4;
        for (size_t e0 = tid; e0 < tot; e0 += 4 * nth) {   // 4 independent row fetches in flight per thread
            float4 v[4];
            size_t d[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const size_t e = e0 + u * nth;
                if (e < tot) {
                    const int r = (int)(e / P4), c = (int)(e % P4);
                    v[u] = __ldg(reinterpret_cast<const float4 *>(g.s_cobs) + (size_t)idx[r] * P4 + c);
                    d[u] = (size_t)r * L4 + c;
                }
            }
#pragma unroll
            for (int u = 0; u < 4; u++)
                if (e0 + u * nth < tot) reinterpret_cast<float4 *>(g.xc)[d[u]] = v[u];
        }
    } else {
        for (size_t e = tid; e < (size_t)g.B * g.P; e += nth) {
            const int r = (int)(e / g.P), c = (int)(e % g.P);
            g.xc[(size_t)r * g.Ppad + c] = __ldg(g.s_cobs + (size_t)idx[r] * g.P + c);
        }
    }
    for (size_t e = tid; e < (size_t)g.B * g.O; e += nth) {
        const int r = (int)(e / g.O), c = (int)(e % g.O);
        g.xa[(size_t)r * g.Opad + c] = __ldg(g.s_obs + (size_t)idx[r] * g.O + c);
    }
    for (size_t e = tid; e < (size_t)g.B * g.A; e += nth) {
        const int r = (int)(e / g.A), c = (int)(e % g.A);
        const size_t s = (size_t)idx[r] * g.A + c;
        g.act[e] = __ldg(g.s_act + s); g.mu[e] = __ldg(g.s_mu + s); g.sigma[e] = __ldg(g.s_sigma + s);
    }
    for (size_t r = tid; r < (size_t)g.B; r += nth) {
        const size_t s = (size_t)idx[r];
        g.val[r] = __ldg(g.s_val + s); g.ret[r] = __ldg(g.s_ret + s); g.adv[r] = __ldg(g.s_adv + s); g.logp[r] = __ldg(g.s_logp + s);
    }
}

// =========================================================================================================
// PPO losses + their gradients w.r.t. the network outputs (ppo.py:246-295; backward as in oracle/ppo_oracle.py)
//   per row: log-prob, ratio, clipped surrogate, clipped value loss, KL(old || new); outputs dmu [B,A], dv [B];
//   block-reduced sums -> tail[kl_sum, count, surrogate_sum, value_loss_sum] and grad(std)
// =========================================================================================================
struct LossArgs {
    const float *mu, *v, *std;                                   // new policy outputs
    const float *act, *old_mu, *old_sigma, *old_logp, *adv, *ret, *old_v;
    float *dmu, *dv, *gstd, *tail;
    int B, A;
    float clip, vcoef, ecoef;
    int clipped_value;
};
__global__ void __launch_bounds__(256) ppo_loss_kernel(const LossArgs a) {
    __shared__ float red[8][MAXA + 4];
    const int r = blockIdx.x * blockDim.x + threadIdx.x, lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    float kl = 0.f, surr = 0.f, vl = 0.f, cnt = 0.f;
    float gs[MAXA];
#pragma unroll
    for (int j = 0; j < MAXA; j++) gs[j] = 0.f;
    if (r < a.B) {
        const float invB = 1.0f / (float)a.B;
        float lp = 0.f;
        for (int j = 0; j < a.A; j++) {
            const float mu = a.mu[(size_t)r * a.A + j], sg = a.std[j], d = a.act[(size_t)r * a.A + j] - mu;
            lp += -(d * d) / (2.f * sg * sg) - logf(sg) - LOG_SQRT_2PI;
            const float os = a.old_sigma[(size_t)r * a.A + j], om = a.old_mu[(size_t)r * a.A + j];
            kl += logf(sg / os + 1.0e-5f) + (os * os + (om - mu) * (om - mu)) / (2.0f * sg * sg) - 0.5f;   // ppo.py:257-261
        }
        const float A_ = a.adv[r];
        const float ratio = expf(lp - a.old_logp[r]);
        const float s1 = -A_ * ratio, s2 = -A_ * fminf(fmaxf(ratio, 1.0f - a.clip), 1.0f + a.clip);
        surr = fmaxf(s1, s2);                                                                            // ppo.py:271-277
        const bool use1 = s1 >= s2, in_clip = ratio >= 1.0f - a.clip && ratio <= 1.0f + a.clip;
        const float dratio = (use1 || in_clip ? -A_ : 0.f) * invB;
        const float dlp = dratio * ratio;
        for (int j = 0; j < a.A; j++) {
            const float mu = a.mu[(size_t)r * a.A + j], sg = a.std[j], d = a.act[(size_t)r * a.A + j] - mu;
            a.dmu[(size_t)r * a.A + j] = dlp * d / (sg * sg);
            gs[j] = dlp * (d * d / (sg * sg * sg) - 1.0f / sg);
        }
        const float v = a.v[r], R = a.ret[r], V0 = a.old_v[r];
        float dv;
        if (a.clipped_value) {                                                                           // ppo.py:280-285
            const float vc = V0 + fminf(fmaxf(v - V0, -a.clip), a.clip);
            const float l1 = (v - R) * (v - R), l2 = (vc - R) * (vc - R);
            vl = fmaxf(l1, l2);
            const bool inc = (v - V0) >= -a.clip && (v - V0) <= a.clip;
            dv = l1 >= l2 ? 2.f * (v - R) : (inc ? 2.f * (vc - R) : 0.f);
        } else {
            vl = (R - v) * (R - v);
            dv = -2.f * (R - v);
        }
        a.dv[r] = dv * (a.vcoef * invB);
        cnt = 1.f;
    }
    // block reduction: 4 scalars + A std-gradients
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        kl += __shfl_xor_sync(FULL, kl, o); surr += __shfl_xor_sync(FULL, surr, o);
        vl += __shfl_xor_sync(FULL, vl, o); cnt += __shfl_xor_sync(FULL, cnt, o);
    }
    for (int j = 0; j < a.A; j++) {
        float x = gs[j];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(FULL, x, o);
        if (lane == 0) red[w][4 + j] = x;
    }
    if (lane == 0) { red[w][0] = kl; red[w][1] = cnt; red[w][2] = surr; red[w][3] = vl; }
    __syncthreads();
    if (threadIdx.x < 4 + a.A) {
        float x = 0.f;
        for (int k = 0; k < 8; k++) x += red[k][threadIdx.x];
        if (threadIdx.x < 4) atomicAdd(&a.tail[threadIdx.x], x);
        else {
            const int j = threadIdx.x - 4;
            if (blockIdx.x == 0) x += -a.ecoef * (1.0f / a.std[j]);   // d(-c_e * mean entropy)/d std_j, once per rank
            atomicAdd(&a.gstd[j], x);
        }
    }
}

// =========================================================================================================
// Fused output heads for the update: for each minibatch row (one warp per row)
//   mu = W3a h3a + b3a, v = W3c h3c + b3c   (the N = 10 / N = 1 output layers: far too narrow for a tensor-core tile)
//   -> the PPO losses and d(loss)/d(mu, v) exactly as ppo_loss_kernel
//   -> dh3 = (d(out) W3) * ELU'(h3) for both nets, and the head gradients dW3 += d(out) (x) h3, db3 += d(out), db2 += colsum(dh3)
// Replaces 2 head GEMMs + loss + 2 head dW GEMMs + 2 head dX GEMMs + 2 column-sum passes.  Memory-bound on h3 / dh3.
// =========================================================================================================
constexpr int HEADS_H = 128;       // hidden width of the last layer on the fused path (one float4 per lane)
constexpr int HEADS_THREADS = 384; // 12 warps per CTA, one CTA per SM (168 registers per thread: every accumulator stays in registers)
struct HeadsArgs {
    const float *h3a, *h3c;            // [B, 128] last hidden activations of actor / critic
    float *dh3a, *dh3c;                // same shapes
    const float *W3a, *b3a, *W3c, *b3c, *std;
    float *gW3a, *gb3a, *gW3c, *gb3c, *gb2a, *gb2c, *gstd, *tail;
    const float *act, *old_mu, *old_sigma, *old_logp, *adv, *ret, *old_v;
    int B;
    float clip, vcoef, ecoef;
    int clipped_value;
};
// One warp per minibatch row; a lane owns the same 4 hidden columns for the whole kernel, so its slices of W3a / W3c and all its
// gradient accumulators live in registers.  Lane j < NA additionally owns action j (its mu_j, log-prob / KL term, d mu_j, grad std_j).
template <int NA>
__global__ void __launch_bounds__(HEADS_THREADS, 1) ppo_heads_kernel(const HeadsArgs a) {
    static_assert(NA < 16, "the head butterfly reduces 16 products: NA actions + the value");
    constexpr int H = HEADS_H;
    constexpr int NACC = NA * H + 3 * H + 2 * NA + 8;   // gW3a | gW3c | gb2a | gb2c | gb3a[NA] | gstd[NA] | gb3c, kl, cnt, surr, vl
    __shared__ float acc[NACC];
    for (int i = threadIdx.x; i < NACC; i += blockDim.x) acc[i] = 0.f;
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wpb = blockDim.x >> 5, nwarps = gridDim.x * wpb;
    const int c0 = lane * 4;
    // the head weights / std were written by the PREVIOUS minibatch's Adam kernel, which completed before the forward chain in front of this
    // kernel could start (every kernel of the chain releases its dependents only after its own griddepcontrol.wait): loading them before this
    // kernel's wait hides their latency under the tail of the last forward layer
    float4 wa[NA], gWa[NA];
#pragma unroll
    for (int j = 0; j < NA; j++) { wa[j] = *reinterpret_cast<const float4 *>(a.W3a + j * H + c0); gWa[j] = make_float4(0.f, 0.f, 0.f, 0.f); }
    const float4 wc = *reinterpret_cast<const float4 *>(a.W3c + c0);
    float4 gWc = make_float4(0.f, 0.f, 0.f, 0.f), g2a = gWc, g2c = gWc;
    const bool own = lane < NA;
    const float sg = own ? a.std[lane] : 1.f, ba = own ? a.b3a[lane] : 0.f, bc = a.b3c[0];
    const float inv2s2 = 1.f / (2.f * sg * sg), logsg = logf(sg);
    tc::pdl_wait();
    tc::pdl_launch_dependents();
    float gba = 0.f, gsd = 0.f, gbc = 0.f, kl_s = 0.f, surr_s = 0.f, vl_s = 0.f, cnt_s = 0.f;
    const float invB = 1.0f / (float)a.B;
    // software-pipelined over rows: the loads of the warp's next row are issued before the arithmetic of the current one
    int r = blockIdx.x * wpb + warp;
    float4 n_ha = make_float4(0.f, 0.f, 0.f, 0.f), n_hc = n_ha;
    float n_ac = 0.f, n_om = 0.f, n_os = 1.f, n_A = 0.f, n_olp = 0.f, n_R = 0.f, n_V0 = 0.f;
    auto fetch = [&](int rr) {
        n_ha = *reinterpret_cast<const float4 *>(a.h3a + (size_t)rr * H + c0);
        n_hc = *reinterpret_cast<const float4 *>(a.h3c + (size_t)rr * H + c0);
        if (own) { n_ac = a.act[(size_t)rr * NA + lane]; n_om = a.old_mu[(size_t)rr * NA + lane]; n_os = a.old_sigma[(size_t)rr * NA + lane]; }
        n_A = a.adv[rr]; n_olp = a.old_logp[rr]; n_R = a.ret[rr]; n_V0 = a.old_v[rr];
    };
    if (r < a.B) fetch(r);
    for (; r < a.B; r += nwarps) {
        const float4 ha = n_ha, hc = n_hc;
        const float ac = n_ac, om = n_om, os = n_os, A_ = n_A, olp = n_olp, R = n_R, V0 = n_V0;
        if (r + nwarps < a.B) fetch(r + nwarps);
        // ---- heads: mu_j = W3a[j] . h3a + b, v = W3c . h3c + b   (butterfly sums; lane j keeps mu_j)
        // 16 partial dot products per lane (10 actions, the value, padding) reduced by a halving butterfly: 16 shuffles instead of 55;
        // afterwards lane l (and its mirror l + 16) holds the complete sum of product l, i.e. lane j already owns mu_j
        float t[16];
#pragma unroll
        for (int j = 0; j < 16; j++) t[j] = 0.f;
#pragma unroll
        for (int j = 0; j < NA; j++) t[j] = ha.x * wa[j].x + ha.y * wa[j].y + ha.z * wa[j].z + ha.w * wa[j].w;
        t[NA] = hc.x * wc.x + hc.y * wc.y + hc.z * wc.z + hc.w * wc.w;
#pragma unroll
        for (int w = 8, bit = 8; w >= 1; w >>= 1, bit >>= 1) {
            const bool up = (lane & bit) != 0;
#pragma unroll
            for (int i = 0; i < w; i++) {
                const float send = up ? t[i] : t[i + w], keep = up ? t[i + w] : t[i];
                t[i] = keep + __shfl_xor_sync(FULL, send, bit);
            }
        }
        t[0] += __shfl_xor_sync(FULL, t[0], 16);
        float mu = own ? t[0] : 0.f;
        float v = __shfl_sync(FULL, t[0], NA);
        v += bc;
        mu += ba;
        // ---- losses (ppo.py:246-295), same arithmetic as ppo_loss_kernel; lane j holds the terms of action j
        const float d = ac - mu;
        float lp = own ? -(d * d) * inv2s2 - logsg - LOG_SQRT_2PI : 0.f;
        float kl = own ? logf(sg / os + 1.0e-5f) + (os * os + (om - mu) * (om - mu)) * inv2s2 - 0.5f : 0.f;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { lp += __shfl_xor_sync(FULL, lp, o); kl += __shfl_xor_sync(FULL, kl, o); }
        const float ratio = expf(lp - olp);
        const float s1 = -A_ * ratio, s2 = -A_ * fminf(fmaxf(ratio, 1.0f - a.clip), 1.0f + a.clip);
        const bool use1 = s1 >= s2, in_clip = ratio >= 1.0f - a.clip && ratio <= 1.0f + a.clip;
        const float dlp = (use1 || in_clip ? -A_ : 0.f) * invB * ratio;
        float dv, vl;
        if (a.clipped_value) {
            const float vc = V0 + fminf(fmaxf(v - V0, -a.clip), a.clip);
            const float l1 = (v - R) * (v - R), l2 = (vc - R) * (vc - R);
            vl = fmaxf(l1, l2);
            const bool inc = (v - V0) >= -a.clip && (v - V0) <= a.clip;
            dv = l1 >= l2 ? 2.f * (v - R) : (inc ? 2.f * (vc - R) : 0.f);
        } else { vl = (R - v) * (R - v); dv = -2.f * (R - v); }
        dv *= a.vcoef * invB;
        kl_s += kl; surr_s += fmaxf(s1, s2); vl_s += vl; cnt_s += 1.f; gbc += dv;      // identical on every lane; lane 0 publishes
        const float dmu = own ? dlp * d / (sg * sg) : 0.f;
        if (own) { gba += dmu; gsd += dlp * (d * d / (sg * sg * sg) - 1.0f / sg); }
        // ---- head backward: dh3 = (d out . W3) * ELU'(h3); dW3 += d out (x) h3; db2 += dh3
        float4 da = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int j = 0; j < NA; j++) {
            const float dj = __shfl_sync(FULL, dmu, j);
            da.x += dj * wa[j].x; da.y += dj * wa[j].y; da.z += dj * wa[j].z; da.w += dj * wa[j].w;
            gWa[j].x += dj * ha.x; gWa[j].y += dj * ha.y; gWa[j].z += dj * ha.z; gWa[j].w += dj * ha.w;
        }
        da.x *= elu_grad_from_out(ha.x); da.y *= elu_grad_from_out(ha.y); da.z *= elu_grad_from_out(ha.z); da.w *= elu_grad_from_out(ha.w);
        *reinterpret_cast<float4 *>(a.dh3a + (size_t)r * H + c0) = da;
        g2a.x += da.x; g2a.y += da.y; g2a.z += da.z; g2a.w += da.w;
        const float4 dc = make_float4(dv * wc.x * elu_grad_from_out(hc.x), dv * wc.y * elu_grad_from_out(hc.y),
                                      dv * wc.z * elu_grad_from_out(hc.z), dv * wc.w * elu_grad_from_out(hc.w));
        *reinterpret_cast<float4 *>(a.dh3c + (size_t)r * H + c0) = dc;
        g2c.x += dc.x; g2c.y += dc.y; g2c.z += dc.z; g2c.w += dc.w;
        gWc.x += dv * hc.x; gWc.y += dv * hc.y; gWc.z += dv * hc.z; gWc.w += dv * hc.w;
    }
    // ---- CTA reduction in shared memory (16 warps), then one red.global per element per CTA
    float *aWa = acc, *aWc = aWa + NA * H, *ab2a = aWc + H, *ab2c = ab2a + H, *agba = ab2c + H, *agsd = agba + NA, *amisc = agsd + NA;
#pragma unroll
    for (int j = 0; j < NA; j++) {
        atomicAdd(&aWa[j * H + c0], gWa[j].x); atomicAdd(&aWa[j * H + c0 + 1], gWa[j].y);
        atomicAdd(&aWa[j * H + c0 + 2], gWa[j].z); atomicAdd(&aWa[j * H + c0 + 3], gWa[j].w);
    }
    atomicAdd(&aWc[c0], gWc.x); atomicAdd(&aWc[c0 + 1], gWc.y); atomicAdd(&aWc[c0 + 2], gWc.z); atomicAdd(&aWc[c0 + 3], gWc.w);
    atomicAdd(&ab2a[c0], g2a.x); atomicAdd(&ab2a[c0 + 1], g2a.y); atomicAdd(&ab2a[c0 + 2], g2a.z); atomicAdd(&ab2a[c0 + 3], g2a.w);
    atomicAdd(&ab2c[c0], g2c.x); atomicAdd(&ab2c[c0 + 1], g2c.y); atomicAdd(&ab2c[c0 + 2], g2c.z); atomicAdd(&ab2c[c0 + 3], g2c.w);
    if (own) { atomicAdd(&agba[lane], gba); atomicAdd(&agsd[lane], gsd); }
    if (lane == 0) { atomicAdd(&amisc[0], gbc); atomicAdd(&amisc[1], kl_s); atomicAdd(&amisc[2], cnt_s); atomicAdd(&amisc[3], surr_s); atomicAdd(&amisc[4], vl_s); }
    __syncthreads();
    // one vector reduction (red.global.add.v4.f32, sm_90+) per four elements: 148 CTAs x 1664 elements would otherwise be 246 K scalar atomics
    // on 1664 addresses; every destination starts on a 16-byte boundary (layout_net) and H % 4 == 0
    auto red4 = [](float *dst, const float *src) {
        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(src[0]), "f"(src[1]), "f"(src[2]), "f"(src[3]) : "memory");
    };
    for (int i = threadIdx.x * 4; i < NA * H; i += blockDim.x * 4) red4(a.gW3a + i, aWa + i);
    for (int i = threadIdx.x * 4; i < 3 * H; i += blockDim.x * 4) {
        const int k = i / H, j = i % H;
        red4((k == 0 ? a.gW3c : k == 1 ? a.gb2a : a.gb2c) + j, (k == 0 ? aWc : k == 1 ? ab2a : ab2c) + j);
    }
    if (threadIdx.x < NA) {
        atomicAdd(&a.gb3a[threadIdx.x], agba[threadIdx.x]);
        float gs = agsd[threadIdx.x];
        if (blockIdx.x == 0) gs += -a.ecoef * (1.0f / a.std[threadIdx.x]);   // d(-c_e * mean entropy)/d std_j, once per rank
        atomicAdd(&a.gstd[threadIdx.x], gs);
    }
    if (threadIdx.x == 0) {
        atomicAdd(&a.gb3c[0], amisc[0]);
        atomicAdd(&a.tail[0], amisc[1]); atomicAdd(&a.tail[1], amisc[2]); atomicAdd(&a.tail[2], amisc[3]); atomicAdd(&a.tail[3], amisc[4]);
    }
}

// =========================================================================================================
// apply: grad-norm -> control block (adaptive LR from KL, NaN skip, clip coefficient, Adam bias corrections) -> Adam
// =========================================================================================================
struct Ctl {            // device control block
    float lr;           // current learning rate (persistent)
    float coef;         // clip coefficient / world_size for this step
    int skip;           // NaN loss -> skip the optimiser step (ppo.py:297-299)
    int step;           // Adam step counter (persistent, shared by all parameters)
    float bc1, bc2s;    // 1 - beta1^t, sqrt(1 - beta2^t)
    float kl_mean, loss, value_loss, surrogate_loss, grad_norm;
    float sum_value_loss, sum_surrogate_loss;   // accumulated over the update (ppo.py:308-309)
    int mb_counter;     // minibatches processed in this update (device-side so a replayed graph advances by itself)
    double sumsq;       // scratch: sum of squares of the gradient
    int comm_epoch;     // NVLink all-reduce epochs completed (flags in peer memory are monotonic epoch numbers)
    int comm_error;     // set when a peer did not show up within the spin budget (results of that step are invalid)
    unsigned comm_arrive;   // scratch: blocks of allreduce_kernel that finished their slice
    unsigned apply_arrive;  // scratch: grid barrier of apply_kernel (blocks that contributed their partial gradient norm)
    unsigned apply_done;    // scratch: blocks of apply_kernel that have read the barrier results (the last one resets the scratch)
    int chain_error;        // set by the chained-layer kernels (grx_mlp_chain.cuh) when a barrier wait timed out (protocol error; results invalid)
    int pad_[2];
    double pow1, pow2;      // beta1^step, beta2^step as running products (fp64 pow is a ~10 us dependent chain on this part)
};

// =========================================================================================================
// Gradient all-reduce over NVLink peer memory, fused with the gradient-norm reduction (multi-GPU).
// Every rank maps every peer's comm block (cudaIpc): [reduce_buf | gsum | flags].  Per minibatch, ONE kernel per rank:
//   1. ready barrier: flags[peer].ready[rank] = epoch (st.release.sys into peer memory); spin on the local ready[] row;
//   2. two-shot: rank r owns float4 slice r of the buffer — it LOADS that slice from all W ranks' reduce_buf over NVLink, sums
//      in rank order (every element is summed by exactly one rank, so all ranks hold bit-identical results), and STORES the
//      sum into every rank's gsum (push all-gather); the squared norm of the slice is reduced on the way;
//   3. the last block to finish publishes the slice's sum of squares and done[rank] = epoch to every peer.
// apply_kernel then waits for all done[] flags, adds the W partial norms in rank order and proceeds exactly as on one GPU.
// No NCCL call, no host involvement: the kernel sits in the same CUDA graph as the rest of the minibatch.
// =========================================================================================================
constexpr int MAXW = 8;
// flag block of a rank, per PHASE (two phases per minibatch, see allreduce_kernel), offsets in ints:
//   ready[8]                 epoch numbers written by every peer ("my gradients of this phase are final")
//   done counter (1 uint)    the last CTA of every rank's allreduce_kernel adds 1 after the rank's slices are pushed and visible: monotonic, target = epoch * world
//                            (one-shot phase 1: every rank adds 1 once its apply_kernel has finished READING the peers' gradients: same target)
//   ticket (1 uint)          local: CTAs of this rank's allreduce_kernel that have finished (the last one publishes and resets it)
//   sumsq                    [world] doubles: squared norm of the slices rank r reduced (one-shot phase 1: [apply grid] local partials), summed in a fixed order by apply_kernel
//   local partials           [NBLK] doubles: per-CTA squared norms of this rank's allreduce_kernel, added by its last CTA
constexpr int NPHASE = 2, AR_NBLK = 128;   // CTAs of allreduce_kernel (every phase)
constexpr int FLAG_READY = 0, FLAG_DONE = 16, FLAG_TICKET = 17, FLAG_SUMSQ = 32, FLAG_LOCAL = 32 + 2 * 256, FLAG_PHASE = FLAG_LOCAL + 2 * AR_NBLK;
static_assert(MAXW <= 256 && 148 <= 256, "sumsq table: world entries (two-shot) or one per apply_kernel block (one-shot)");
constexpr size_t FLAG_BYTES = (size_t)NPHASE * FLAG_PHASE * 4;
constexpr int MAXRANGE = 4;
struct Ranges {         // float4 index ranges [lo, hi) of the gradient block one all-reduce phase covers
    int lo[MAXRANGE], hi[MAXRANGE], n;
};

struct CommDev {
    float *grads[MAXW];
    float *gsum[MAXW];
    int *flags[MAXW];
    int rank, world;
};
__device__ __forceinline__ void st_release_sys(int *p, int v) { asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ int ld_acquire_sys(const int *p) {
    int v;
    asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
// spin until *p >= epoch; gives up after `budget_ns` (cfg.comm_timeout_ms) and reports through ctl->comm_error, which is STICKY:
// apply_kernel does not touch the parameters once it is set and the host raises (grx_ppo_update / grx_ppo_check) — never hangs the GPU
__device__ __forceinline__ void wait_flag(const int *p, int epoch, Ctl *ctl, unsigned long long budget_ns) {
    const unsigned long long t0 = global_ns();
    while ((int)((unsigned)ld_acquire_sys(p) - (unsigned)epoch) < 0) {   // wrap-safe: the done counters grow by CTAs x ranks per minibatch
        if (global_ns() - t0 > budget_ns) { atomicExch(&ctl->comm_error, 1); break; }
        __nanosleep(64);
    }
}
__global__ void __launch_bounds__(256) allreduce_kernel(const CommDev c, const Ranges rg, int phase, int nparam4, Ctl *ctl, unsigned long long budget_ns) {
    const int epoch = ctl->comm_epoch + 1;
    const int fo = phase * FLAG_PHASE;
    if (threadIdx.x == 0) tc::stamp(16 + 4 * phase);
    if (blockIdx.x == 0 && threadIdx.x < c.world) st_release_sys(c.flags[threadIdx.x] + fo + FLAG_READY + c.rank, epoch);
    if (threadIdx.x < c.world) wait_flag(c.flags[c.rank] + fo + FLAG_READY + threadIdx.x, epoch, ctl, budget_ns);
    __syncthreads();
    if (threadIdx.x == 0) tc::stamp(17 + 4 * phase);
    float ss = 0.f;
    for (int q = 0; q < rg.n; q++) {
        const int len = rg.hi[q] - rg.lo[q], per = (len + c.world - 1) / c.world;
        const int lo = rg.lo[q] + c.rank * per, hi = min(rg.hi[q], lo + per);
        for (int i = lo + blockIdx.x * blockDim.x + threadIdx.x; i < hi; i += gridDim.x * blockDim.x) {
            float4 v[MAXW];
#pragma unroll
            for (int r = 0; r < MAXW; r++)   // all peer loads in flight at once (one NVLink round trip, not `world` of them)
                if (r < c.world) v[r] = reinterpret_cast<const float4 *>(c.grads[r])[i];
            float4 s = v[0];
#pragma unroll
            for (int r = 1; r < MAXW; r++)
                if (r < c.world) { s.x += v[r].x; s.y += v[r].y; s.z += v[r].z; s.w += v[r].w; }   // rank order: bit-identical on every rank
#pragma unroll
            for (int r = 0; r < MAXW; r++)
                if (r < c.world) reinterpret_cast<float4 *>(c.gsum[r])[i] = s;
            if (i < nparam4) ss += s.x * s.x + s.y * s.y + s.z * s.z + s.w * s.w;   // the tail (KL / loss sums) is not part of the gradient
        }
    }
    // publication: every CTA leaves the squared-norm partial of its slice in a LOCAL table, makes its pushes visible system-wide (one system
    // fence by thread 0, cumulative over the CTA's stores through the block barrier) and takes a ticket on a local counter; the LAST CTA of the
    // rank adds the partials in a fixed order, stores the rank's total into every rank's table and adds 1 to every rank's done counter
    // (release: ordered after this thread's stores and, by cumulativity, after everything the tickets made visible).  Remote traffic per phase
    // and rank: W stores + W atomics — per-CTA remote atomics (W x 128 per counter) do not scale: measured 45.6 ms per update on 8 GPUs
    // against 38.7 ms on 2.
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(FULL, ss, o);
    __shared__ float red[8];
    __shared__ int s_last;
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
    __syncthreads();
    int *loc = c.flags[c.rank] + fo;
    if (threadIdx.x == 0) {
        tc::stamp(18 + 4 * phase);
        double x = 0.0;
        for (int k = 0; k < 8; k++) x += (double)red[k];
        reinterpret_cast<volatile double *>(loc + FLAG_LOCAL)[blockIdx.x] = x;
        __threadfence_system();
        s_last = atomicAdd(reinterpret_cast<unsigned *>(loc + FLAG_TICKET), 1u) == gridDim.x - 1 ? 1 : 0;
    }
    __syncthreads();
    if (s_last && threadIdx.x < 32) {
        __threadfence();
        double t = 0.0;
        for (int i = threadIdx.x; i < (int)gridDim.x; i += 32) t += reinterpret_cast<const volatile double *>(loc + FLAG_LOCAL)[i];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(FULL, t, o);   // fixed tree: the same total on every run
        if ((int)threadIdx.x < c.world) {
            reinterpret_cast<volatile double *>(c.flags[threadIdx.x] + fo + FLAG_SUMSQ)[c.rank] = t;
            asm volatile("red.release.sys.global.add.u32 [%0], %1;" ::"l"(c.flags[threadIdx.x] + fo + FLAG_DONE), "r"(1u) : "memory");
        }
        if (threadIdx.x == 0) { *reinterpret_cast<volatile unsigned *>(loc + FLAG_TICKET) = 0u; tc::stamp(19 + 4 * phase); }
    }
}
struct PrepArgs {
    Ctl *ctl;
    const float *tail, *std;
    int A, adaptive, world_size;
    float desired_kl, lr_min, lr_max, max_grad_norm, vcoef, ecoef;
    int *comm_flags;   // this rank's flag block when the NVLink all-reduce is active (norm partials come from the peers), else NULL
    int comm_rank;
    unsigned long long budget_ns;   // peer-flag wait budget
    float4 *zero;      // gradient block to clear once the step is applied (graph mode with prefetched gathers), or NULL
    int nzero4;
    float *mb_log;     // [mb_log_cap][4] = (kl_mean, lr, loss, grad_norm) of minibatch ctl.mb_counter of this update (what ppo.py:262-268, 308-309 would log)
    int mb_log_cap;
    // ONE-SHOT phase 1 (the input-layer weight gradients, the last ones produced): instead of a third kernel + its push / fence / counter chain,
    // apply_kernel itself publishes "ready", loads the phase's ranges from ALL ranks over NVLink, sums them in rank order (bit-identical on
    // every rank) into its local gsum and takes their part of the gradient norm through its own grid barrier.
    int one_shot;
    CommDev comm;
    Ranges r1;
    float *gsum_w;     // == g, writable
};
// ONE kernel for the apply step (ppo.py:262-268, 297-305): gradient norm -> [grid barrier] -> adaptive LR from the mean KL,
// NaN skip, clip coefficient, Adam bias corrections (evaluated identically by every block from the same inputs) -> Adam on the
// block's slice; block 0 persists the control block.  The grid (<= one block per SM) is co-resident, so the barrier is an
// arrival counter in the control block; the last block to have read the results resets the scratch for the next launch.
// torch.optim.Adam semantics (ppo.py:81; betas 0.9/0.999, eps 1e-8, no weight decay).
__global__ void __launch_bounds__(1024) apply_kernel(const PrepArgs a, float *__restrict__ p, const float *__restrict__ g, float *__restrict__ m,
                                                      float *__restrict__ v, int n) {
    Ctl &c = *a.ctl;
    __shared__ float red[32];
    __shared__ double s_total;
    __shared__ int s_comm_bad;
    if (threadIdx.x == 0) tc::stamp(0);
    const float old_lr = c.lr;
    const int old_step = c.step;
    const double old_pow1 = c.pow1, old_pow2 = c.pow2;
    float ent = 0.f;   // read std BEFORE the grid barrier: after it, other blocks' Adam may already be updating the parameters (std = params[0..A))
    for (int j = 0; j < a.A; j++) ent += 0.5f + 0.5f * 1.8378770664093453f + logf(a.std[j]);   // ACM:160-163
    if (a.comm_flags == nullptr) {
        float s = 0.f;
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < (n >> 2); i += gridDim.x * blockDim.x) {   // n % 4 == 0, 16-byte aligned
            const float4 x = reinterpret_cast<const float4 *>(g)[i];
            s += x.x * x.x + x.y * x.y + x.z * x.z + x.w * x.w;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(FULL, s, o);
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    } else if (!a.one_shot) {
        if (threadIdx.x < NPHASE)   // wait until every CTA of every rank has pushed its slice of the summed gradient (and its partial norm), both phases
            wait_flag(a.comm_flags + threadIdx.x * FLAG_PHASE + FLAG_DONE, (c.comm_epoch + 1) * a.world_size, a.ctl, a.budget_ns);
    } else {
        const int epoch = c.comm_epoch + 1;
        // this kernel is a normal (fully serialised) launch behind the last weight-gradient GEMM: this rank's input-layer gradients are final
        if (blockIdx.x == 0 && threadIdx.x < a.world_size) st_release_sys(a.comm.flags[threadIdx.x] + FLAG_PHASE + FLAG_READY + a.comm_rank, epoch);
        if (threadIdx.x == 0) { if (a.one_shot == 1) wait_flag(a.comm_flags + FLAG_DONE, epoch * a.world_size, a.ctl, a.budget_ns); }   // phase 0 pushed by everybody (one_shot == 2: there is no phase 0, this kernel pulls everything)
        else if ((int)threadIdx.x <= a.world_size) wait_flag(a.comm_flags + FLAG_PHASE + FLAG_READY + (threadIdx.x - 1), epoch, a.ctl, a.budget_ns);
    }
    __syncthreads();
    if (threadIdx.x == 0) tc::stamp(1);
    if (a.comm_flags != nullptr && a.one_shot) {   // phase 1: pull the ranges from every rank, rank-order sum -> local gsum, squared norm of this block's share
        float s = 0.f;
        for (int q = 0; q < a.r1.n; q++) {
            for (int i = a.r1.lo[q] + blockIdx.x * blockDim.x + threadIdx.x; i < a.r1.hi[q]; i += gridDim.x * blockDim.x) {
                float4 v[MAXW];
#pragma unroll
                for (int r = 0; r < MAXW; r++)   // all peer loads in flight at once
                    if (r < a.world_size) v[r] = reinterpret_cast<const float4 *>(a.comm.grads[r])[i];
                float4 x = v[0];
#pragma unroll
                for (int r = 1; r < MAXW; r++)
                    if (r < a.world_size) { x.x += v[r].x; x.y += v[r].y; x.z += v[r].z; x.w += v[r].w; }
                reinterpret_cast<float4 *>(a.gsum_w)[i] = x;
                if (i < (n >> 2)) s += x.x * x.x + x.y * x.y + x.z * x.z + x.w * x.w;   // the tail (KL / loss sums) is not part of the gradient
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(FULL, s, o);
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        if (a.comm_flags == nullptr) {
            float x = 0.f;
            for (int k = 0; k < (int)(blockDim.x >> 5); k++) x += red[k];
            atomicAdd(&c.sumsq, (double)x);
        } else if (a.one_shot) {   // this block's partial -> its slot of the LOCAL phase-1 table (every rank computes the same partials: same elements, same sums)
            double x = 0.0;
            for (int k = 0; k < (int)(blockDim.x >> 5); k++) x += (double)red[k];
            reinterpret_cast<volatile double *>(a.comm_flags + FLAG_PHASE + FLAG_SUMSQ)[blockIdx.x] = x;
        }
        __threadfence();
        atomicAdd(&c.apply_arrive, 1u);
        const long long t0 = clock64();
        while (*reinterpret_cast<volatile unsigned *>(&c.apply_arrive) < gridDim.x) {
            if (clock64() - t0 > 8000000000ll) { atomicExch(&c.comm_error, 2); break; }   // grid not co-resident (should not happen): never hang the GPU
            __nanosleep(32);
        }
        __threadfence();
        if (a.comm_flags == nullptr) s_total = *reinterpret_cast<volatile double *>(&c.sumsq);
        // sticky: a peer-flag / grid-barrier wait that timed out (now or in an earlier minibatch) leaves the summed gradient undefined.
        // Every block reads the flag after the grid barrier, so all of them take the same decision: do not touch the parameters.
        s_comm_bad = *reinterpret_cast<volatile int *>(&c.comm_error);
    }
    __syncthreads();
    if (a.comm_flags != nullptr) {
        if (threadIdx.x < 32) {   // squared norm = sum of the partials of all ranks and phases, in an order that is the same everywhere
            const int per_phase = a.world_size;   // one total per rank
            double t = 0.0;
            for (int ph = a.one_shot == 2 ? 1 : 0; ph < NPHASE; ph++) {
                const volatile double *tab = reinterpret_cast<const volatile double *>(a.comm_flags + ph * FLAG_PHASE + FLAG_SUMSQ);
                const int cnt = (ph == 1 && a.one_shot) ? (int)gridDim.x : per_phase;   // one-shot phase 1: the partials of this kernel's own blocks
                for (int i = threadIdx.x; i < cnt; i += 32) t += tab[i];
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(FULL, t, o);
            if (threadIdx.x == 0) s_total = t;
        }
        // one-shot: every block of this rank is past the grid barrier, i.e. done reading the peers' gradients -> tell them (they wait for it before
        // their gradient block is cleared / accumulated into again)
        if (a.one_shot && blockIdx.x == 0 && threadIdx.x >= 32 && (int)threadIdx.x < 32 + a.world_size)
            asm volatile("red.release.sys.global.add.u32 [%0], %1;" ::"l"(a.comm.flags[threadIdx.x - 32] + FLAG_PHASE + FLAG_DONE), "r"(1u) : "memory");
        __syncthreads();
    }
    if (threadIdx.x == 0) tc::stamp(2);
    // ---- control scalars, identical in every block
    const float cnt = a.tail[1];
    const float kl_mean = a.tail[0] / cnt;
    float lr = old_lr;
    if (a.adaptive) {                                                                  // ppo.py:262-268, 207-213
        if (kl_mean > a.desired_kl * 2.0f) lr = fmaxf(a.lr_min, lr / 1.5f);
        else if (kl_mean < a.desired_kl / 2.0f && kl_mean > 0.0f) lr = fminf(a.lr_max, lr * 1.5f);
    }
    const float surr = a.tail[2] / cnt, vl = a.tail[3] / cnt;
    const float loss = surr + a.vcoef * vl - a.ecoef * ent;
    const bool skip = isnan(loss) || s_comm_bad != 0;
    const float W = (float)a.world_size;
    const float total = (float)sqrt(s_total) / W;                                      // grads are sums over ranks
    const float coef = fminf(a.max_grad_norm / (total + 1e-6f), 1.0f) / W;             // clip_grad_norm_ (ppo.py:304)
    const int step = old_step + (skip ? 0 : 1);
    const double pow1 = skip ? old_pow1 : old_pow1 * 0.9, pow2 = skip ? old_pow2 : old_pow2 * 0.999;
    const float bc1 = (float)(1.0 - pow1), bc2s = sqrtf((float)(1.0 - pow2));
    if (threadIdx.x == 0) tc::stamp(3);
    if (!skip) {
        const float step_size = lr / bc1;
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < (n >> 2); i += gridDim.x * blockDim.x) {
            const float4 g4 = reinterpret_cast<const float4 *>(g)[i];
            float4 m4 = reinterpret_cast<float4 *>(m)[i], v4 = reinterpret_cast<float4 *>(v)[i], p4 = reinterpret_cast<float4 *>(p)[i];
            const float gg[4] = {g4.x, g4.y, g4.z, g4.w};
            float *mm = &m4.x, *vv = &v4.x, *pp = &p4.x;
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const float gi = gg[k] * coef;
                const float mi = 0.9f * mm[k] + (1.0f - 0.9f) * gi;
                const float vi = 0.999f * vv[k] + (1.0f - 0.999f) * gi * gi;
                mm[k] = mi; vv[k] = vi;
                const float denom = sqrtf(vi) / bc2s + 1e-8f;
                pp[k] -= step_size * (mi / denom);
            }
            reinterpret_cast<float4 *>(m)[i] = m4; reinterpret_cast<float4 *>(v)[i] = v4; reinterpret_cast<float4 *>(p)[i] = p4;
        }
    }
    if (a.comm_flags != nullptr && a.one_shot) {   // the peers' one-shot reads of this rank's input-layer gradients must be over before they are cleared or accumulated into again
        if (threadIdx.x == 0) wait_flag(a.comm_flags + FLAG_PHASE + FLAG_DONE, (c.comm_epoch + 1) * a.world_size, a.ctl, a.budget_ns);
        __syncthreads();
    }
    if (a.zero != nullptr)   // every reader of this minibatch's gradients (this kernel; the peers' all-reduce, whose completion was awaited above) is done
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < a.nzero4; i += gridDim.x * blockDim.x) a.zero[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (threadIdx.x == 0) tc::stamp(4);
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned done = atomicAdd(&c.apply_done, 1u);
        if (done == gridDim.x - 1) {   // everybody has read the barrier results and the old control values: persist + reset scratch
            c.lr = lr; c.coef = coef; c.skip = skip ? 1 : 0; c.step = step; c.bc1 = bc1; c.bc2s = bc2s; c.pow1 = pow1; c.pow2 = pow2;
            c.kl_mean = kl_mean; c.loss = loss; c.value_loss = vl; c.surrogate_loss = surr; c.grad_norm = total;
            if (a.mb_log != nullptr && c.mb_counter >= 0 && c.mb_counter < a.mb_log_cap) {
                float *row = a.mb_log + 4 * c.mb_counter;
                row[0] = kl_mean; row[1] = lr; row[2] = loss; row[3] = total;
            }
            c.mb_counter += 1;
            if (!skip) { c.sum_value_loss += vl; c.sum_surrogate_loss += surr; }
            if (a.comm_flags != nullptr) c.comm_epoch += 1;
            c.sumsq = 0.0; c.apply_arrive = 0u; c.apply_done = 0u;
        }
    }
}

struct Net {            // one MLP: offsets into the flat parameter vector
    int dims[5];        // in, h1, h2, h3, out
    int ld[4];          // leading dimension of weight l = in-dim rounded up to 4 floats (16-byte rows for TMA / vector loads)
    size_t w[4], b[4];
};

}  // namespace

// =========================================================================================================
// Host side
// =========================================================================================================
struct grx_ppo {
    grx_ppo_cfg cfg;
    int device = 0;
    int Opad = 0, Ppad = 0;   // row stride of the staged actor / critic inputs (O, P rounded up to 4 floats)
    int N = 0, T = 0, O = 0, P = 0, A = 0, B = 0, MR = 0;   // B = minibatch rows, MR = workspace rows = max(N, B)
    size_t nparam = 0;
    Net actor, critic;
    float *params = nullptr, *reduce_buf = nullptr, *adam_m = nullptr, *adam_v = nullptr;
    // storage [T, N, .]
    float *s_obs = nullptr, *s_cobs = nullptr, *s_act = nullptr, *s_val = nullptr, *s_rew = nullptr, *s_logp = nullptr, *s_mu = nullptr,
          *s_sigma = nullptr, *s_ret = nullptr, *s_adv = nullptr;
    uint8_t *s_done = nullptr;
    // workspace
    float *ha[4] = {nullptr, nullptr, nullptr, nullptr}, *hc[4] = {nullptr, nullptr, nullptr, nullptr};     // activations (h1..h3, out)
    float *da[4] = {nullptr, nullptr, nullptr, nullptr}, *dc[4] = {nullptr, nullptr, nullptr, nullptr};     // gradients w.r.t. them
    float *xa = nullptr, *xc = nullptr, *mb_act = nullptr, *mb_val = nullptr, *mb_ret = nullptr, *mb_adv = nullptr, *mb_logp = nullptr,
          *mb_mu = nullptr, *mb_sigma = nullptr, *last_values = nullptr;
    struct MbIn { float *xa, *xc, *act, *val, *ret, *adv, *logp, *mu, *sigma; } mbin[2] = {};   // minibatch inputs, double-buffered: the gather of
                                                                                             // minibatch k+1 runs beside the kernels of minibatch k
    cudaStream_t gstream = nullptr;
    cudaEvent_t ev_gfork = nullptr, ev_gjoin = nullptr;
    double *moments = nullptr;
    Ctl *ctl = nullptr;
    float *mb_log = nullptr;   // per-minibatch (kl, lr, loss, grad norm) of the current update
    int mb_log_cap = 0;
    std::vector<void *> allocs;
    cudaGraphExec_t graph = nullptr;
    const int64_t *graph_indices = nullptr;
    unsigned long long graph_kernels = 0;   // kernel nodes of one epoch graph
    // phase timing of the stepwise minibatch entries (GRX_PPO_TIMING=1; profiling only: it synchronises after every minibatch)
    bool timing = false;
    cudaEvent_t tev[9] = {};
    double tacc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    int tcount = 0;
    // NVLink peer-memory all-reduce (world_size > 1): comm block = [reduce_buf | gsum | flags], one cudaIpc handle per rank
    void *comm_block = nullptr;
    size_t comm_bytes = 0;
    float *gsum = nullptr;
    int *flags = nullptr;
    bool comm_open = false;
    int apply_grid = 148;   // co-resident grid of apply_kernel (<= SM count)
    unsigned long long comm_budget_ns = 10000000000ull;   // cfg.comm_timeout_ms (GRX_COMM_TIMEOUT_MS overrides)
    int *pipe_sync = nullptr;   // row-block counters of the layer-pipelined launches (tc::launch_pipe), zero between launches
    // Gradient all-reduce protocol (fixed for the object's lifetime; GRX_COMM_ONESHOT overrides): 0 = two allreduce_kernel phases; 1 = phase 0 as a
    // kernel hidden behind the input-layer weight-gradient launch, phase 1 pulled inside apply_kernel; 2 = no all-reduce kernel at all: one merged
    // weight-gradient launch (as on one GPU) and apply_kernel pulls the WHOLE gradient block from every rank (bandwidth-bound: (W - 1) x 1.75 MB
    // over NVLink, one round trip of latency instead of a flag / push / fence / counter chain per phase)
    int one_shot = 1;
    CommDev comm;
    std::vector<void *> peer_maps;
    // the all-reduce runs in two phases so that most of it hides behind the last weight-gradient launch:
    //   phase 0 = every gradient that is final once the hidden-layer weight gradients are done (all but the two input-layer matrices) + the KL / loss tail,
    //   phase 1 = the input-layer weight gradients (24 % of the bytes)
    Ranges phase_ranges[NPHASE];
    cudaStream_t side = nullptr;           // phase 0 is forked onto this stream (inside the update's CUDA graph: a parallel branch)
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    bool phase0_launched = false;
};

static int ppo_alloc(grx_ppo *p, void **ptr, size_t bytes) {
    CK(cudaMalloc(ptr, bytes));
    CK(cudaMemset(*ptr, 0, bytes));
    p->allocs.push_back(*ptr);
    return GRX_OK;
}
#define PALLOC(ptr, count) do { int rc__ = ppo_alloc(p, (void **)&(ptr), (size_t)(count)); if (rc__) return rc__; } while (0)

static void layout_net(Net &n, const int *dims, size_t &off) {
    for (int i = 0; i < 5; i++) n.dims[i] = dims[i];
    for (int l = 0; l < 4; l++) {   // every tensor starts on a 16-byte boundary (cp.async / vector loads in the dense kernels)
        n.ld[l] = (dims[l] + 3) & ~3;
        off = (off + 3) & ~(size_t)3; n.w[l] = off; off += (size_t)dims[l + 1] * n.ld[l];
        off = (off + 3) & ~(size_t)3; n.b[l] = off; off += (size_t)dims[l + 1];
    }
}

extern "C" int grx_ppo_create(const grx_ppo_cfg *cfg, int32_t device, grx_ppo **out) {
    if (!cfg || !out) return grx_set_error(GRX_E_INVALID, "grx_ppo_create: null argument");
    if (cfg->num_envs <= 0 || cfg->num_steps <= 0 || cfg->num_steps > GAE_TMAX || cfg->num_actions > MAXA || cfg->num_actions <= 0 ||
        cfg->num_mini_batches <= 0 || cfg->num_learning_epochs <= 0 || cfg->world_size <= 0)
        return grx_set_error(GRX_E_INVALID, "grx_ppo_create: bad sizes (need num_steps <= 256, num_actions <= 32)");
    CK(cudaSetDevice(device));
    grx_ppo *p = new grx_ppo();
    p->cfg = *cfg; p->device = device;
    p->N = cfg->num_envs; p->T = cfg->num_steps; p->O = cfg->num_obs; p->P = cfg->num_pri_obs; p->A = cfg->num_actions;
    p->B = (int)(((size_t)p->N * p->T) / cfg->num_mini_batches);                       // rollout_storage.py:71-72
    if (p->B <= 0) { delete p; return grx_set_error(GRX_E_INVALID, "grx_ppo_create: fewer transitions than minibatches"); }
    p->MR = p->N > p->B ? p->N : p->B;
    p->Opad = (p->O + 3) & ~3; p->Ppad = (p->P + 3) & ~3;
    // flat parameter vector in the reference's state_dict order: std, actor.model.{0,2,4,6}.{weight,bias}, critic...
    size_t off = (size_t)p->A;
    const int da[5] = {p->O, cfg->actor_hidden[0], cfg->actor_hidden[1], cfg->actor_hidden[2], p->A};
    const int dc[5] = {p->P, cfg->critic_hidden[0], cfg->critic_hidden[1], cfg->critic_hidden[2], 1};
    layout_net(p->actor, da, off);
    layout_net(p->critic, dc, off);
    p->nparam = (off + 3) & ~(size_t)3;
    {   // all-reduce phases (float4 index ranges of the gradient block)
        const int aw_lo = (int)(p->actor.w[0] / 4), aw_hi = (int)((p->actor.w[0] + (size_t)p->actor.dims[1] * p->actor.ld[0]) / 4);
        const int cw_lo = (int)(p->critic.w[0] / 4), cw_hi = (int)((p->critic.w[0] + (size_t)p->critic.dims[1] * p->critic.ld[0]) / 4);
        const int n4 = (int)((p->nparam + TAIL) / 4);
        Ranges &r0 = p->phase_ranges[0], &r1 = p->phase_ranges[1];
        r0.n = 3; r0.lo[0] = 0; r0.hi[0] = aw_lo; r0.lo[1] = aw_hi; r0.hi[1] = cw_lo; r0.lo[2] = cw_hi; r0.hi[2] = n4;
        r1.n = 2; r1.lo[0] = aw_lo; r1.hi[0] = aw_hi; r1.lo[1] = cw_lo; r1.hi[1] = cw_hi;
    }
    const size_t TN = (size_t)p->T * p->N, MR = p->MR;
    PALLOC(p->params, p->nparam * 4); PALLOC(p->adam_m, p->nparam * 4); PALLOC(p->adam_v, p->nparam * 4);
    {   // comm block: its own cudaMalloc so that one cudaIpc handle covers exactly [reduce_buf | gsum | flags]
        const size_t nb = ((p->nparam + TAIL) * 4 + 255) & ~(size_t)255;
        p->comm_bytes = 2 * nb + FLAG_BYTES;
        PALLOC(p->comm_block, p->comm_bytes);
        p->reduce_buf = (float *)p->comm_block;
        p->gsum = (float *)((char *)p->comm_block + nb);
        p->flags = (int *)((char *)p->comm_block + 2 * nb);
        memset(&p->comm, 0, sizeof(p->comm));
    }
    PALLOC(p->s_obs, TN * p->O * 4); PALLOC(p->s_cobs, TN * p->P * 4); PALLOC(p->s_act, TN * p->A * 4); PALLOC(p->s_val, TN * 4);
    PALLOC(p->s_rew, TN * 4); PALLOC(p->s_logp, TN * 4); PALLOC(p->s_mu, TN * p->A * 4); PALLOC(p->s_sigma, TN * p->A * 4);
    PALLOC(p->s_ret, TN * 4); PALLOC(p->s_adv, TN * 4); PALLOC(p->s_done, TN);
    for (int l = 0; l < 4; l++) {
        PALLOC(p->ha[l], MR * da[l + 1] * 4); PALLOC(p->da[l], MR * da[l + 1] * 4);
        PALLOC(p->hc[l], MR * dc[l + 1] * 4); PALLOC(p->dc[l], MR * dc[l + 1] * 4);
    }
    PALLOC(p->xa, MR * p->Opad * 4); PALLOC(p->xc, MR * p->Ppad * 4); PALLOC(p->mb_act, MR * p->A * 4); PALLOC(p->mb_val, MR * 4);
    PALLOC(p->mb_ret, MR * 4); PALLOC(p->mb_adv, MR * 4); PALLOC(p->mb_logp, MR * 4); PALLOC(p->mb_mu, MR * p->A * 4);
    PALLOC(p->mb_sigma, MR * p->A * 4); PALLOC(p->last_values, (size_t)p->N * 4);
    PALLOC(p->pipe_sync, (size_t)tc::SYNC_INTS * 4);
    p->mbin[0] = {p->xa, p->xc, p->mb_act, p->mb_val, p->mb_ret, p->mb_adv, p->mb_logp, p->mb_mu, p->mb_sigma};
    {   // second set (only minibatch rows)
        const size_t Bm = (size_t)p->B;
        grx_ppo::MbIn &m = p->mbin[1];
        PALLOC(m.xa, Bm * p->Opad * 4); PALLOC(m.xc, Bm * p->Ppad * 4); PALLOC(m.act, Bm * p->A * 4); PALLOC(m.val, Bm * 4); PALLOC(m.ret, Bm * 4);
        PALLOC(m.adv, Bm * 4); PALLOC(m.logp, Bm * 4); PALLOC(m.mu, Bm * p->A * 4); PALLOC(m.sigma, Bm * p->A * 4);
    }
    CK(cudaStreamCreateWithFlags(&p->gstream, cudaStreamNonBlocking));
    CK(cudaEventCreateWithFlags(&p->ev_gfork, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&p->ev_gjoin, cudaEventDisableTiming));
    PALLOC(p->moments, 4 * sizeof(double)); PALLOC(p->ctl, sizeof(Ctl));
    p->mb_log_cap = cfg->num_learning_epochs * cfg->num_mini_batches;
    PALLOC(p->mb_log, (size_t)p->mb_log_cap * 4 * sizeof(float));
    CK(cudaFuncSetAttribute(gae_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 3 * GAE_TMAX * 33 * (int)sizeof(float)));
    {   // std = init_noise_std (ACM:79-82); weights are loaded by the host (same nn.Linear init stream as the reference); lr
        std::vector<float> s(p->A, cfg->init_noise_std);
        CK(cudaMemcpy(p->params, s.data(), p->A * 4, cudaMemcpyHostToDevice));
        Ctl c; memset(&c, 0, sizeof(c));
        c.lr = cfg->learning_rate;
        c.pow1 = 1.0; c.pow2 = 1.0;
        CK(cudaMemcpy(p->ctl, &c, sizeof(c), cudaMemcpyHostToDevice));
    }
    {
        int sms = 0;
        if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device) == cudaSuccess && sms > 0) p->apply_grid = sms;
    }
    if (cfg->comm_timeout_ms > 0) p->comm_budget_ns = (unsigned long long)cfg->comm_timeout_ms * 1000000ull;
    if (const char *e = getenv("GRX_COMM_TIMEOUT_MS")) { const long v = atol(e); if (v > 0) p->comm_budget_ns = (unsigned long long)v * 1000000ull; }
    if (const char *e = getenv("GRX_PPO_TIMING")) p->timing = atoi(e) != 0;
    // measured on one 8-GPU box (profiles/bench_r3h_*, bench_r3f_*, bench_r3e_*; M env-steps/s): protocol 2 / protocol 1 = 7.82 / 7.65 on 2 GPUs,
    // 15.30 / 15.03 on 4, 28.88 / 29.39 on 8 — pulling (W - 1) x 1.75 MB per rank is bandwidth-bound and overtakes the latency-bound chain at W > 4
    p->one_shot = cfg->world_size <= 4 ? 2 : 1;
    if (const char *e = getenv("GRX_COMM_ONESHOT")) { const int v = atoi(e); if (v >= 0 && v <= 2) p->one_shot = v; }
    if (p->timing) for (int i = 0; i < 9; i++) CK(cudaEventCreate(&p->tev[i]));
    *out = p;
    return GRX_OK;
}

extern "C" int grx_ppo_destroy(grx_ppo *p) {
    if (!p) return GRX_OK;
    cudaSetDevice(p->device);
    if (p->graph) cudaGraphExecDestroy(p->graph);
    if (p->side) cudaStreamDestroy(p->side);
    if (p->gstream) cudaStreamDestroy(p->gstream);
    if (p->ev_gfork) cudaEventDestroy(p->ev_gfork);
    if (p->ev_gjoin) cudaEventDestroy(p->ev_gjoin);
    if (p->ev_fork) cudaEventDestroy(p->ev_fork);
    if (p->ev_join) cudaEventDestroy(p->ev_join);
    for (void *q : p->peer_maps) cudaIpcCloseMemHandle(q);
    for (void *q : p->allocs) cudaFree(q);
    delete p;
    return GRX_OK;
}

static int g_ppo_buf_device = 0;
static void set_buf(grx_buffer *b, void *data, int dtype, int ndim, int64_t d0, int64_t d1, int64_t d2) {
    b->device = g_ppo_buf_device; b->own_data = 0;
    b->data = data; b->dtype = dtype; b->ndim = ndim;
    for (int i = 0; i < GRX_MAX_DIMS; i++) { b->dims[i] = 1; b->strides[i] = 1; }
    b->dims[0] = d0; b->dims[1] = d1; b->dims[2] = d2;
    b->strides[0] = d1 * d2; b->strides[1] = d2; b->strides[2] = 1;
}

extern "C" int grx_ppo_get_buffer(grx_ppo *p, const char *name, grx_buffer *b) {
    if (!p || !name || !b) return grx_set_error(GRX_E_INVALID, "grx_ppo_get_buffer: null argument");
    const std::string n(name);
    g_ppo_buf_device = p->device;
    const int64_t T = p->T, N = p->N, np_ = (int64_t)p->nparam;
    if (n == "params") { set_buf(b, p->params, GRX_F32, 1, np_, 1, 1); return GRX_OK; }
    if (n == "grads") { set_buf(b, p->reduce_buf, GRX_F32, 1, np_, 1, 1); return GRX_OK; }
    if (n == "reduce_buf") { set_buf(b, p->reduce_buf, GRX_F32, 1, np_ + TAIL, 1, 1); return GRX_OK; }
    if (n == "adam_m") { set_buf(b, p->adam_m, GRX_F32, 1, np_, 1, 1); return GRX_OK; }
    if (n == "adam_v") { set_buf(b, p->adam_v, GRX_F32, 1, np_, 1, 1); return GRX_OK; }
    if (n == "obs") { set_buf(b, p->s_obs, GRX_F32, 3, T, N, p->O); return GRX_OK; }
    if (n == "critic_obs") { set_buf(b, p->s_cobs, GRX_F32, 3, T, N, p->P); return GRX_OK; }
    if (n == "actions") { set_buf(b, p->s_act, GRX_F32, 3, T, N, p->A); return GRX_OK; }
    if (n == "mu") { set_buf(b, p->s_mu, GRX_F32, 3, T, N, p->A); return GRX_OK; }
    if (n == "sigma") { set_buf(b, p->s_sigma, GRX_F32, 3, T, N, p->A); return GRX_OK; }
    if (n == "values") { set_buf(b, p->s_val, GRX_F32, 3, T, N, 1); return GRX_OK; }
    if (n == "rewards") { set_buf(b, p->s_rew, GRX_F32, 3, T, N, 1); return GRX_OK; }
    if (n == "actions_log_prob") { set_buf(b, p->s_logp, GRX_F32, 3, T, N, 1); return GRX_OK; }
    if (n == "returns") { set_buf(b, p->s_ret, GRX_F32, 3, T, N, 1); return GRX_OK; }
    if (n == "advantages") { set_buf(b, p->s_adv, GRX_F32, 3, T, N, 1); return GRX_OK; }
    if (n == "dones") { set_buf(b, p->s_done, GRX_U8, 3, T, N, 1); return GRX_OK; }
    if (n == "adv_moments") { set_buf(b, p->moments, GRX_U64, 1, 3, 1, 1); return GRX_OK; }   // 3 doubles (bit pattern)
    if (n == "gsum") { set_buf(b, p->gsum, GRX_F32, 1, np_ + TAIL, 1, 1); return GRX_OK; }
    if (n == "ctl") { set_buf(b, p->ctl, GRX_F32, 1, sizeof(Ctl) / 4, 1, 1); return GRX_OK; }
    if (n == "mb_log") { set_buf(b, p->mb_log, GRX_F32, 2, p->mb_log_cap, 4, 1); return GRX_OK; }
    return grx_set_error(GRX_E_NOTFOUND, "grx_ppo_get_buffer: unknown buffer '" + n + "'");
}

// MLP forward over M rows for `nn` networks at once (x has row stride ldx): h[l] = elu(h[l-1] W_l^T + b_l) for l < nlayers, last of 4
// without ELU (mlp.py:26-41).  Layer l of every network goes into ONE grouped launch.
struct NetIO {
    const Net *net;
    const float *x;
    int ldx;
    float *const *h;
    float *const *d;
};
static void mlp_forward(grx_ppo *p, const NetIO *io, int nn, int M, int nlayers, cudaStream_t st, bool store_hidden = true) {
    int l0 = 0;
    if (p->cfg.use_tensor_cores != 0 && tc::chain::fwd_enabled() && nlayers >= 3 && nn <= 2) {
        // registered widths: the three hidden layers as ONE chained tcgen05 kernel (grx_mlp_chain.cuh); activations stay on chip between layers
        tc::chain::FwdProblem ps[2];
        bool ok = true;
        for (int i = 0; i < nn; i++) {
            const Net &net = *io[i].net;
            ok = ok && net.dims[1] == tc::chain::D1 && net.dims[2] == tc::chain::D2 && net.dims[3] == tc::chain::D3;
            tc::chain::FwdProblem &q = ps[i];
            q.X = io[i].x; q.K0 = net.dims[0]; q.ldx = io[i].ldx;
            q.W0 = p->params + net.w[0]; q.b0 = p->params + net.b[0]; q.ld0 = net.ld[0];
            q.W1 = p->params + net.w[1]; q.b1 = p->params + net.b[1];
            q.W2 = p->params + net.w[2]; q.b2 = p->params + net.b[2];
            q.H1 = io[i].h[0]; q.H2 = io[i].h[1]; q.H3 = io[i].h[2];
        }
        if (ok && tc::chain::fwd_supported(ps, nn)) {
            const cudaError_t e = tc::chain::launch_fwd(ps, nn, M, store_hidden, &p->ctl->chain_error, st);
            if (e != cudaSuccess && g_launch_err == cudaSuccess) g_launch_err = e;
            l0 = 3;
        }
    }
    auto layer_args = [&](int l, int i, GemmArgs &g) {
        const Net &net = *io[i].net;
        memset(&g, 0, sizeof(GemmArgs));
        g.A = l == 0 ? io[i].x : io[i].h[l - 1]; g.B = p->params + net.w[l]; g.C = io[i].h[l]; g.bias = p->params + net.b[l];
        g.M = M; g.N = net.dims[l + 1]; g.K = net.dims[l]; g.lda = l == 0 ? io[i].ldx : net.dims[l]; g.ldb = net.ld[l]; g.ldc = g.N;
    };
    const int nhid = nlayers < 3 ? nlayers : 3;   // layers with the bias + ELU epilogue
    if (l0 == 0 && nhid >= 2 && nn * nhid <= tc::MAXP && (tc::pipe_flag() & 1)) {   // the hidden layers of all networks as ONE layer-pipelined launch (tc::launch_pipe)
        GemmArgs g[tc::MAXP];
        int deps[tc::MAXP];
        for (int l = 0; l < nhid; l++)
            for (int i = 0; i < nn; i++) { layer_args(l, i, g[l * nn + i]); deps[l * nn + i] = l == 0 ? -1 : (l - 1) * nn + i; }
        if (dense_pipe<true, true, 1>(g, deps, nn * nhid, p->cfg.use_tensor_cores != 0, p->pipe_sync, &p->ctl->chain_error, st)) l0 = nhid;
    }
    for (int l = l0; l < nlayers; l++) {
        GemmArgs g[2];
        for (int i = 0; i < nn; i++) layer_args(l, i, g[i]);
        if (l < 3) dense_group<true, true, 1>(g, nullptr, nn, p->cfg.use_tensor_cores != 0, st);
        else dense_group<true, true, 0>(g, nullptr, nn, false, st);   // output heads in fp32 (as on the fused-heads path): mu feeds exp((a - mu)^2 / 2 sigma^2), TF32 rounding
    }                                                                 // of the head itself costs percents on the ratio and the std gradient (measured), and the GEMM is tiny
}
// MLP backward from layer `top` down for `nn` networks: d[top] holds dL/d(output of layer top).  First the input-gradient chain
// d[l-1] = (d[l] W_l) * ELU'(h[l-1]) (+ db_{l-1} = column sums), one grouped launch per layer; then ALL weight gradients
// dW_l = d[l]^T h[l-1] (split-K, accumulated with red.global.add) in grouped launches.  Bias grads: l < top from the column sums,
// l == top only when top == 3 (with the fused heads kernel top == 2 and db_2, dW_3, db_3 are already done).
static void launch_allreduce_phase(grx_ppo *p, int phase, cudaStream_t st) {
    grx_count_launch();
    allreduce_kernel<<<AR_NBLK, 256, 0, st>>>(p->comm, p->phase_ranges[phase], phase, (int)(p->nparam / 4), p->ctl, p->comm_budget_ns);
}
static void mlp_backward(grx_ppo *p, const NetIO *io, int nn, float *grads, int M, int top, cudaStream_t st, bool overlap_comm = false) {
    const bool tcu = p->cfg.use_tensor_cores != 0;
    auto dx_args = [&](int l, int i, GemmArgs &g) {
        const Net &net = *io[i].net;
        memset(&g, 0, sizeof(GemmArgs));
        g.A = io[i].d[l]; g.B = p->params + net.w[l]; g.C = io[i].d[l - 1]; g.aux = io[i].h[l - 1]; g.bias_out = grads + net.b[l - 1];
        g.M = M; g.N = net.dims[l]; g.K = net.dims[l + 1]; g.lda = net.dims[l + 1]; g.ldb = net.ld[l]; g.ldc = net.dims[l];
    };
    int ltop = top;
    if (top >= 2 && top <= 3 && nn * 2 <= tc::MAXP && (tc::pipe_flag() & 2)) {   // the two widest input-gradient layers (l = 2 -> 1) as ONE layer-pipelined launch; l = 3 (heads, K < 64) stays apart
        if (top == 3) {
            GemmArgs g3[2];
            for (int i = 0; i < nn; i++) dx_args(3, i, g3[i]);
            dense_group<true, false, 2>(g3, nullptr, nn, tcu, st);
        }
        GemmArgs g[tc::MAXP];
        int deps[tc::MAXP];
        for (int s = 0; s < 2; s++)
            for (int i = 0; i < nn; i++) { dx_args(2 - s, i, g[s * nn + i]); deps[s * nn + i] = s == 0 ? -1 : i; }
        if (dense_pipe<true, false, 2>(g, deps, nn * 2, tcu, p->pipe_sync, &p->ctl->chain_error, st)) ltop = 0;
        else if (top == 3) ltop = 2;
    }
    for (int l = ltop; l >= 1; l--) {
        GemmArgs g[2];
        for (int i = 0; i < nn; i++) dx_args(l, i, g[i]);
        dense_group<true, false, 2>(g, nullptr, nn, tcu, st);
    }
    GemmArgs g[8];
    int splits[8], n = 0;
    auto flush = [&]() { if (n) dense_group<false, false, 3>(g, splits, n, tcu, st); n = 0; };
    // Single GPU: all weight gradients in ONE grouped launch of 128-column tiles (the narrow actor input layer pads its one tile): measured 180.9 ->
    // 174.7 us per minibatch (GRX_DW_MERGE=0 restores the two launches).  Multi-GPU keeps two launches: phase 0 of the all-reduce overlaps the second.
    static const bool merge_env = [] { const char *e = getenv("GRX_DW_MERGE"); return e ? atoi(e) != 0 : true; }();
    const bool merge = merge_env && !overlap_comm;
    for (int pass = 0; pass < (merge ? 1 : 2); pass++) {   // pass 0: hidden layers (wide N), pass 1: input layer (narrow N) -> their own tile shape
        for (int l = top; l >= 0; l--) {     // (one launch for all six with the cost model free to pick BN = 64 was measured slower)
            if (!merge && (l == 0) != (pass == 1)) continue;
            for (int i = 0; i < nn; i++) {
                const Net &net = *io[i].net;
                GemmArgs &a = g[n];
                memset(&a, 0, sizeof(GemmArgs));
                a.A = io[i].d[l]; a.B = l == 0 ? io[i].x : io[i].h[l - 1]; a.C = grads + net.w[l];
                a.bias_out = (l == top && top == 3) ? grads + net.b[l] : nullptr;
                a.M = net.dims[l + 1]; a.N = l == 0 ? net.ld[0] : net.dims[l]; a.K = M; a.lda = net.dims[l + 1]; a.ldb = l == 0 ? io[i].ldx : net.dims[l];
                a.ldc = net.ld[l];
                static const int dw_chunk = [] { const char *e = getenv("GRX_DW_CHUNK"); const int v = e ? atoi(e) : 512; return v >= 32 ? v : 512; }();
                splits[n] = (M + dw_chunk - 1) / dw_chunk;   // contraction rows per split
                if (++n == tc::MAXP) flush();
            }
        }
        flush();
        if (overlap_comm) {
            if (pass == 0) {   // fork: phase 0 of the gradient all-reduce runs beside the input-layer weight-gradient launch (64 small CTAs fit next to the GEMM's one CTA per SM)
                cudaEventRecord(p->ev_fork, st);
                cudaStreamWaitEvent(p->side, p->ev_fork, 0);
                launch_allreduce_phase(p, 0, p->side);
                cudaEventRecord(p->ev_join, p->side);
                p->phase0_launched = true;
            } else {
                cudaStreamWaitEvent(st, p->ev_join, 0);   // join
            }
        }
    }
}
// stage caller-provided observation rows (any row stride >= width) into the padded, 16-byte aligned input buffers
static cudaError_t stage_rows(float *dst, int ld_dst, const float *src, int width, int rows, cudaStream_t st) {
    return cudaMemcpy2DAsync(dst, (size_t)ld_dst * 4, src, (size_t)width * 4, (size_t)width * 4, rows, cudaMemcpyDeviceToDevice, st);
}

extern "C" int grx_ppo_act(grx_ppo *p, const float *d_obs, const float *d_critic_obs, const float *d_eps, int32_t t, float *d_actions_out,
                           uint64_t step_index, void *stream) {
    if (!p || !d_obs || !d_critic_obs || !d_actions_out || t < 0 || t >= p->T) return grx_set_error(GRX_E_INVALID, "grx_ppo_act: bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    // rows that are already 16-byte aligned with a 16-byte pitch (the registered critic input: 168 floats) are read in place by TMA; the others
    // (actor: 39 floats) are staged into the padded input buffers
    const float *xa = p->xa, *xc = p->xc;
    if (p->O == p->Opad && tc::aligned16(d_obs)) xa = d_obs; else CK(stage_rows(p->xa, p->Opad, d_obs, p->O, p->N, st));
    if (p->P == p->Ppad && tc::aligned16(d_critic_obs)) xc = d_critic_obs; else CK(stage_rows(p->xc, p->Ppad, d_critic_obs, p->P, p->N, st));
    const NetIO io[2] = {{&p->actor, xa, p->Opad, p->ha, p->da}, {&p->critic, xc, p->Ppad, p->hc, p->dc}};
    const bool fused_heads = p->A == 10 && p->critic.dims[4] == 1 && p->actor.dims[3] == HEADS_H && p->critic.dims[3] == HEADS_H;
    mlp_forward(p, io, 2, p->N, fused_heads ? 3 : 4, st, false);   // rollout: only the last hidden layer leaves the chip
    ActArgs a; memset(&a, 0, sizeof(a));
    const size_t row = (size_t)t * p->N;
    a.obs = d_obs; a.critic_obs = d_critic_obs; a.mu = p->ha[3]; a.value = p->hc[3]; a.std = p->params; a.eps = d_eps;
    a.actions_out = d_actions_out;
    a.s_obs = p->s_obs + row * p->O; a.s_cobs = p->s_cobs + row * p->P; a.s_act = p->s_act + row * p->A; a.s_val = p->s_val + row;
    a.s_logp = p->s_logp + row; a.s_mu = p->s_mu + row * p->A; a.s_sigma = p->s_sigma + row * p->A;
    a.N = p->N; a.O = p->O; a.P = p->P; a.A = p->A; a.seed = p->cfg.seed ^ 0x9E3779B97F4A7C15ull; a.step_index = step_index; a.env_id_offset = p->cfg.env_id_offset;   // Philox key (task seed), counter (GLOBAL env id, step)
    if (fused_heads) {
        ActHeadsArgs q;
        q.h3a = p->ha[2]; q.h3c = p->hc[2];
        q.W3a = p->params + p->actor.w[3]; q.b3a = p->params + p->actor.b[3]; q.W3c = p->params + p->critic.w[3]; q.b3c = p->params + p->critic.b[3];
        q.a = a;
        CK(tc::launch_kernel(act_heads_kernel<10>, dim3((p->N + 7) / 8 < 1184 ? max((p->N + 7) / 8, 148) : 1184), dim3(256), 0, st, true, q));
    } else {
        grx_count_launch();
        act_sample_store_kernel<<<max((p->N + 127) / 128, 592), 128, 0, st>>>(a);
    }
    CK(cudaGetLastError());
    return GRX_OK;
}

extern "C" int grx_ppo_process_env_step(grx_ppo *p, const float *d_rewards, const uint8_t *d_dones, const uint8_t *d_time_outs, int32_t t,
                                        void *stream) {
    if (!p || !d_rewards || !d_dones || t < 0 || t >= p->T) return grx_set_error(GRX_E_INVALID, "grx_ppo_process_env_step: bad argument");
    const size_t row = (size_t)t * p->N;
    grx_count_launch();
    process_env_step_kernel<<<(p->N + 255) / 256, 256, 0, (cudaStream_t)stream>>>(d_rewards, d_dones, d_time_outs, p->s_val + row, p->cfg.gamma,
                                                                                   p->s_rew + row, p->s_done + row, p->N);
    CK(cudaGetLastError());
    return GRX_OK;
}

extern "C" int grx_ppo_compute_returns_local(grx_ppo *p, const float *d_last_critic_obs, void *stream) {
    if (!p || !d_last_critic_obs) return grx_set_error(GRX_E_INVALID, "grx_ppo_compute_returns: null argument");
    cudaStream_t st = (cudaStream_t)stream;
    const float *xc = p->xc;
    if (p->P == p->Ppad && tc::aligned16(d_last_critic_obs)) xc = d_last_critic_obs; else CK(stage_rows(p->xc, p->Ppad, d_last_critic_obs, p->P, p->N, st));
    const NetIO io = {&p->critic, xc, p->Ppad, p->hc, p->dc};
    mlp_forward(p, &io, 1, p->N, 4, st, false);                                        // ppo.py:204
    CK(cudaMemcpyAsync(p->last_values, p->hc[3], (size_t)p->N * 4, cudaMemcpyDeviceToDevice, st));
    CK(cudaMemsetAsync(p->moments, 0, 4 * sizeof(double), st));
    grx_count_launch();
    gae_kernel<<<(p->N + 31) / 32, 1024, (size_t)3 * p->T * 33 * sizeof(float), st>>>(p->s_rew, p->s_done, p->s_val, p->last_values, p->cfg.gamma, p->cfg.lam, p->s_ret, p->s_adv,
                                                  p->moments, p->T, p->N);
    CK(cudaGetLastError());
    return GRX_OK;
}
extern "C" int grx_ppo_normalize_advantages(grx_ppo *p, void *stream) {
    if (!p) return grx_set_error(GRX_E_INVALID, "null ppo");
    grx_count_launch();
    normalize_adv_kernel<<<296, 256, 0, (cudaStream_t)stream>>>(p->s_adv, p->moments, (size_t)p->T * p->N);
    CK(cudaGetLastError());
    return GRX_OK;
}
extern "C" int grx_ppo_compute_returns(grx_ppo *p, const float *d_last_critic_obs, void *stream) {
    int rc = grx_ppo_compute_returns_local(p, d_last_critic_obs, stream);
    if (rc) return rc;
    return grx_ppo_normalize_advantages(p, stream);
}

// minibatch rows of the rollout -> input set `set`; zero: also clear the gradient block (inline gathers; a prefetched gather runs beside the
// previous minibatch's accumulation and must not)
static void launch_gather(grx_ppo *p, const int64_t *d_indices, int mb, bool device_counter, int mb_off, int set, bool zero, cudaStream_t st, bool pdl) {
    GatherArgs g; memset(&g, 0, sizeof(g));
    const grx_ppo::MbIn &m = p->mbin[set];
    g.indices = d_indices; g.mb_counter = device_counter ? &p->ctl->mb_counter : nullptr; g.mb_off = mb_off; g.mb = mb; g.nmb = p->cfg.num_mini_batches;
    g.B = p->B; g.O = p->O; g.P = p->P; g.A = p->A; g.Opad = p->Opad; g.Ppad = p->Ppad;
    g.s_obs = p->s_obs; g.s_cobs = p->s_cobs; g.s_act = p->s_act; g.s_val = p->s_val; g.s_ret = p->s_ret; g.s_adv = p->s_adv;
    g.s_logp = p->s_logp; g.s_mu = p->s_mu; g.s_sigma = p->s_sigma;
    g.xa = m.xa; g.xc = m.xc; g.act = m.act; g.val = m.val; g.ret = m.ret; g.adv = m.adv; g.logp = m.logp; g.mu = m.mu; g.sigma = m.sigma;
    g.zero = reinterpret_cast<float4 *>(p->reduce_buf); g.nzero4 = zero ? (int)((p->nparam + TAIL) / 4) : 0;
    const cudaError_t e = tc::launch_kernel(gather_kernel, dim3(148 * 8), dim3(256), 0, st, pdl, g);
    if (e != cudaSuccess && g_launch_err == cudaSuccess) g_launch_err = e;
}

// set: which minibatch-input set the forward / backward read; gather_inline: run the gather (and clear the gradient block) first, in-stream
static int minibatch_grads(grx_ppo *p, const int64_t *d_indices, int mb, bool device_counter, cudaStream_t st, bool overlap_comm = false, int set = 0,
                           bool gather_inline = true) {
    const int B = p->B;
    const bool tm = p->timing && !device_counter;
#define TMARK(i) do { if (tm) cudaEventRecord(p->tev[i], st); } while (0)
    TMARK(0);
    if (gather_inline) launch_gather(p, d_indices, mb, device_counter, 0, set, true, st, true);
    TMARK(1);
    const grx_ppo::MbIn &mi = p->mbin[set];
    const Net &na = p->actor, &nc = p->critic;
    const bool fused_heads = p->A == 10 && nc.dims[4] == 1 && na.dims[3] == HEADS_H && nc.dims[3] == HEADS_H;   // the registered GRx policy; other shapes take the unfused path
    const NetIO io[2] = {{&na, mi.xa, p->Opad, p->ha, p->da}, {&nc, mi.xc, p->Ppad, p->hc, p->dc}};
    TMARK(2);
    mlp_forward(p, io, 2, B, fused_heads ? 3 : 4, st);                                 // ppo.py:244-248
    TMARK(3);
    float *gr = p->reduce_buf;
    if (fused_heads) {
        HeadsArgs a; memset(&a, 0, sizeof(a));
        a.h3a = p->ha[2]; a.h3c = p->hc[2]; a.dh3a = p->da[2]; a.dh3c = p->dc[2];
        a.W3a = p->params + na.w[3]; a.b3a = p->params + na.b[3]; a.W3c = p->params + nc.w[3]; a.b3c = p->params + nc.b[3]; a.std = p->params;
        a.gW3a = gr + na.w[3]; a.gb3a = gr + na.b[3]; a.gW3c = gr + nc.w[3]; a.gb3c = gr + nc.b[3]; a.gb2a = gr + na.b[2]; a.gb2c = gr + nc.b[2];
        a.gstd = gr; a.tail = gr + p->nparam;
        a.act = mi.act; a.old_mu = mi.mu; a.old_sigma = mi.sigma; a.old_logp = mi.logp; a.adv = mi.adv; a.ret = mi.ret; a.old_v = mi.val;
        a.B = B;
        a.clip = p->cfg.clip_param; a.vcoef = p->cfg.value_loss_coef; a.ecoef = p->cfg.entropy_coef; a.clipped_value = p->cfg.use_clipped_value_loss;
        { const cudaError_t e = tc::launch_kernel(ppo_heads_kernel<10>, dim3(148), dim3(HEADS_THREADS), 0, st, true, a); if (e != cudaSuccess && g_launch_err == cudaSuccess) g_launch_err = e; }
        TMARK(4);
        TMARK(5);
        mlp_backward(p, io, 2, gr, B, 2, st, overlap_comm);
        TMARK(6);
    } else {
        LossArgs a; memset(&a, 0, sizeof(a));
        a.mu = p->ha[3]; a.v = p->hc[3]; a.std = p->params; a.act = mi.act; a.old_mu = mi.mu; a.old_sigma = mi.sigma;
        a.old_logp = mi.logp; a.adv = mi.adv; a.ret = mi.ret; a.old_v = mi.val;
        a.dmu = p->da[3]; a.dv = p->dc[3]; a.gstd = gr; a.tail = gr + p->nparam;
        a.B = B; a.A = p->A; a.clip = p->cfg.clip_param; a.vcoef = p->cfg.value_loss_coef; a.ecoef = p->cfg.entropy_coef;
        a.clipped_value = p->cfg.use_clipped_value_loss;
        grx_count_launch();
        ppo_loss_kernel<<<(B + 255) / 256, 256, 0, st>>>(a);
        TMARK(4);
        TMARK(5);
        mlp_backward(p, io, 2, gr, B, 3, st, overlap_comm);
        TMARK(6);
    }
    if (g_launch_err != cudaSuccess) { const cudaError_t e = g_launch_err; g_launch_err = cudaSuccess; return grx_set_error(GRX_E_CUDA, std::string("tensor-core GEMM launch: ") + cudaGetErrorString(e)); }
    CK(cudaGetLastError());
    return GRX_OK;
}
static bool prefetch_enabled() {   // GRX_GATHER_PREFETCH=0: every gather in-stream at the head of its minibatch (A-B comparison)
    static const int on = [] { const char *e = getenv("GRX_GATHER_PREFETCH"); return e ? atoi(e) : 1; }();
    return on != 0;
}
static bool overlap_enabled() {   // GRX_COMM_OVERLAP=0: both all-reduce phases after the last backward launch (profiling / A-B comparison)
    static const int on = [] { const char *e = getenv("GRX_COMM_OVERLAP"); return e ? atoi(e) : 1; }();
    return on != 0;
}
static int minibatch_apply(grx_ppo *p, cudaStream_t st, bool use_comm, bool zero_grads = false) {
    const float *gsrc = p->reduce_buf;
    if (use_comm) {   // NVLink all-reduce + norm; the summed gradient lands in gsum on every rank
        if (p->one_shot != 2 && !p->phase0_launched) launch_allreduce_phase(p, 0, st);   // not forked by minibatch_grads (stepwise entry): both phases back to back
        if (p->one_shot == 0) launch_allreduce_phase(p, 1, st);
        p->phase0_launched = false;
        gsrc = p->gsum;
    }
    PrepArgs a; memset(&a, 0, sizeof(a));
    a.comm_flags = use_comm ? p->flags : nullptr; a.comm_rank = p->comm.rank; a.budget_ns = p->comm_budget_ns;
    a.one_shot = use_comm ? p->one_shot : 0; a.comm = p->comm; a.r1 = p->phase_ranges[1]; a.gsum_w = p->gsum;
    if (a.one_shot == 2) { a.r1.n = 1; a.r1.lo[0] = 0; a.r1.hi[0] = (int)((p->nparam + TAIL) / 4); }   // the whole block incl. the KL / loss tail
    a.mb_log = p->mb_log; a.mb_log_cap = p->mb_log_cap;
    a.zero = zero_grads ? reinterpret_cast<float4 *>(p->reduce_buf) : nullptr; a.nzero4 = (int)((p->nparam + TAIL) / 4);
    a.ctl = p->ctl; a.tail = gsrc + p->nparam; a.std = p->params; a.A = p->A; a.adaptive = p->cfg.adaptive_schedule;
    a.world_size = p->cfg.world_size; a.desired_kl = p->cfg.desired_kl; a.lr_min = p->cfg.learning_rate_min; a.lr_max = p->cfg.learning_rate_max;
    a.max_grad_norm = p->cfg.max_grad_norm; a.vcoef = p->cfg.value_loss_coef; a.ecoef = p->cfg.entropy_coef;
    // (a normal launch: as a programmatic dependent the 1024-thread blocks queue up behind the last GEMM's CTAs and the grid barrier
    // then waits on the stragglers — measured slower)
    CK(tc::launch_kernel(apply_kernel, dim3(p->apply_grid), dim3(1024), 0, st, false, a, p->params, (const float *)gsrc, p->adam_m, p->adam_v, (int)p->nparam));
    CK(cudaGetLastError());
    if (p->timing) {
        cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
        cudaStreamIsCapturing(st, &cs);
        if (cs == cudaStreamCaptureStatusNone) {
            cudaEventRecord(p->tev[7], st);
            CK(cudaEventSynchronize(p->tev[7]));
            for (int i = 0; i < 7; i++) { float ms = 0.f; if (cudaEventElapsedTime(&ms, p->tev[i], p->tev[i + 1]) == cudaSuccess) p->tacc[i] += ms; }
            p->tcount++;
        }
    }
    return GRX_OK;
}
/* profiling: mean microseconds per minibatch of [memset+gather, actor fwd, critic fwd, heads/loss, actor bwd, critic bwd, apply] */
extern "C" int grx_ppo_debug_timing(grx_ppo *p, float *out7, int32_t *count) {
    if (!p || !out7) return grx_set_error(GRX_E_INVALID, "grx_ppo_debug_timing: null argument");
    for (int i = 0; i < 7; i++) out7[i] = p->tcount ? (float)(p->tacc[i] / p->tcount * 1e3) : 0.f;
    if (count) *count = p->tcount;
    for (int i = 0; i < 8; i++) p->tacc[i] = 0;
    p->tcount = 0;
    return GRX_OK;
}

extern "C" int grx_ppo_minibatch_grads(grx_ppo *p, const int64_t *d_indices, int32_t mb, void *stream) {
    if (!p || !d_indices || mb < 0 || mb >= p->cfg.num_mini_batches) return grx_set_error(GRX_E_INVALID, "grx_ppo_minibatch_grads: bad argument");
    return minibatch_grads(p, d_indices, mb, false, (cudaStream_t)stream);
}
extern "C" int grx_ppo_minibatch_apply(grx_ppo *p, void *stream) {
    if (!p) return grx_set_error(GRX_E_INVALID, "null ppo");
    return minibatch_apply(p, (cudaStream_t)stream, false);   // the caller all-reduced reduce_buf itself (NCCL path) or world_size == 1
}

// ---- NVLink all-reduce setup: exchange grx_ppo_comm_handle() blobs between the ranks (any host channel), then open them
extern "C" int grx_ppo_comm_handle(grx_ppo *p, void *out64) {
    if (!p || !out64) return grx_set_error(GRX_E_INVALID, "grx_ppo_comm_handle: null argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
    cudaIpcMemHandle_t h;
    CK(cudaSetDevice(p->device));
    CK(cudaIpcGetMemHandle(&h, p->comm_block));
    memcpy(out64, &h, sizeof(h));
    return GRX_OK;
}
extern "C" int grx_ppo_comm_open(grx_ppo *p, int32_t rank, int32_t world, const void *handles) {
    if (!p || !handles || world < 2 || world > MAXW || rank < 0 || rank >= world || world != p->cfg.world_size)
        return grx_set_error(GRX_E_INVALID, "grx_ppo_comm_open: need 2 <= world == cfg.world_size <= 8 and 0 <= rank < world");
    if (p->comm_open) return grx_set_error(GRX_E_STATE, "grx_ppo_comm_open: already open");
    CK(cudaSetDevice(p->device));
    const size_t nb = ((p->nparam + TAIL) * 4 + 255) & ~(size_t)255;
    for (int r = 0; r < world; r++) {
        void *base = p->comm_block;
        if (r != rank) {
            cudaIpcMemHandle_t h;
            memcpy(&h, (const char *)handles + (size_t)r * sizeof(h), sizeof(h));
            CK(cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess));
            p->peer_maps.push_back(base);
        }
        p->comm.grads[r] = (float *)base;
        p->comm.gsum[r] = (float *)((char *)base + nb);
        p->comm.flags[r] = (int *)((char *)base + 2 * nb);
    }
    p->comm.rank = rank; p->comm.world = world;
    if (!p->side) {
        CK(cudaStreamCreateWithFlags(&p->side, cudaStreamNonBlocking));
        CK(cudaEventCreateWithFlags(&p->ev_fork, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&p->ev_join, cudaEventDisableTiming));
    }
    p->comm_open = true;
    if (p->graph) { cudaGraphExecDestroy(p->graph); p->graph = nullptr; }
    return GRX_OK;
}
/* apply step of the NVLink path: all-reduce (peer memory) + norm + LR / clip + Adam; every rank must call it once per minibatch */
extern "C" int grx_ppo_minibatch_apply_comm(grx_ppo *p, void *stream) {
    if (!p || !p->comm_open) return grx_set_error(GRX_E_STATE, "grx_ppo_minibatch_apply_comm: call grx_ppo_comm_open first");
    return minibatch_apply(p, (cudaStream_t)stream, true);
}

// Whole PPO.update (ppo.py:215-321): one EPOCH (num_mini_batches x [grads + apply]) is captured once as a CUDA graph whose gather
// kernels read the minibatch index from the device control block, then replayed num_learning_epochs times.
extern "C" int grx_ppo_update(grx_ppo *p, const int64_t *d_indices, void *stream) {
    if (!p || !d_indices) return grx_set_error(GRX_E_INVALID, "grx_ppo_update: null argument");
    if (p->cfg.world_size > 1 && !p->comm_open)
        return grx_set_error(GRX_E_STATE, "grx_ppo_update: world_size > 1 needs grx_ppo_comm_open (NVLink all-reduce) or the stepwise grads / all-reduce / apply entries");
    cudaStream_t st = (cudaStream_t)stream;
    // reset the per-update accumulators (mb_counter, loss sums); lr / Adam step persist
    CK(cudaMemsetAsync(&p->ctl->sum_value_loss, 0, 2 * sizeof(float) + sizeof(int), st));
    if (!p->graph || p->graph_indices != d_indices) {
        if (p->graph) { cudaGraphExecDestroy(p->graph); p->graph = nullptr; }
        cudaStream_t cs;
        CK(cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking));
        cudaGraph_t gr;
        const unsigned long long before_capture = g_grx_launches.load();
        CK(cudaStreamBeginCapture(cs, cudaStreamCaptureModeThreadLocal));
        int rc = 0;
        // one graph = one epoch (the gathers read the device-side minibatch counter).  The gather of minibatch k+1 is PREFETCHED: forked onto a
        // side branch at the start of minibatch k (into the other input set), it runs beside the dense layers of minibatch k (two of those
        // launches leave 66 SMs idle) and is joined before apply(k) advances the counter; apply clears the gradient block instead of the gather.
        const bool prefetch = prefetch_enabled();
        const int nmb = p->cfg.num_mini_batches;
        for (int mb = 0; mb < nmb && !rc; mb++) {
            const int set = prefetch ? (mb & 1) : 0;
            const bool inline_gather = !prefetch || mb == 0;
            if (prefetch && mb + 1 < nmb) {
                if (inline_gather) launch_gather(p, d_indices, 0, true, 0, set, true, cs, true);   // (minibatch 0: its own gather first, so the fork comes after it)
                cudaEventRecord(p->ev_gfork, cs);
                cudaStreamWaitEvent(p->gstream, p->ev_gfork, 0);
                launch_gather(p, d_indices, 0, true, 1, set ^ 1, false, p->gstream, false);
                cudaEventRecord(p->ev_gjoin, p->gstream);
                rc = minibatch_grads(p, d_indices, 0, true, cs, p->comm_open && overlap_enabled() && p->one_shot != 2, set, false);
                cudaStreamWaitEvent(cs, p->ev_gjoin, 0);
            } else {
                rc = minibatch_grads(p, d_indices, 0, true, cs, p->comm_open && overlap_enabled() && p->one_shot != 2, set, inline_gather);
            }
            if (!rc) rc = minibatch_apply(p, cs, p->comm_open, prefetch);
        }
        cudaError_t ce = cudaStreamEndCapture(cs, &gr);
        cudaStreamDestroy(cs);
        p->graph_kernels = g_grx_launches.load() - before_capture;   // recorded, not executed: counted per replay below
        g_grx_launches.store(before_capture);
        if (rc) return rc;
        CK(ce);
        CK(cudaGraphInstantiate(&p->graph, gr, 0));
        cudaGraphDestroy(gr);
        p->graph_indices = d_indices;
    }
    for (int i = 0; i < p->cfg.num_learning_epochs; i++) { CK(cudaGraphLaunch(p->graph, st)); grx_count_launch(p->graph_kernels); }
    return GRX_OK;
}

extern "C" int grx_ppo_act_inference(grx_ppo *p, const float *d_obs, int32_t n, float *d_actions_out, void *stream) {
    if (!p || !d_obs || !d_actions_out || n <= 0 || n > p->MR) return grx_set_error(GRX_E_INVALID, "grx_ppo_act_inference: bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    CK(stage_rows(p->xa, p->Opad, d_obs, p->O, n, st));
    const NetIO io = {&p->actor, p->xa, p->Opad, p->ha, p->da};
    mlp_forward(p, &io, 1, n, 4, st, false);
    CK(cudaMemcpyAsync(d_actions_out, p->ha[3], (size_t)n * p->A * 4, cudaMemcpyDeviceToDevice, st));
    return GRX_OK;
}

// Test / profiling: switch the chained-layer kernels (grx_mlp_chain.cuh) on / off at run time; returns the previous setting of the forward chain.
extern "C" int grx_ppo_debug_fused(int32_t fwd_chain) {
    const int old = tc::chain::fwd_flag();
    if (fwd_chain >= 0) tc::chain::fwd_flag() = fwd_chain;
    return old;
}

// Test / profiling: layer-pipelined dense-layer launches (tc::launch_pipe) on / off at run time; returns the previous setting.  A captured update
// graph keeps the setting it was captured with.
extern "C" int grx_ppo_debug_pipe(int32_t on) {
    const int old = tc::pipe_flag();
    if (on >= 0) tc::pipe_flag() = on;
    return old;
}

// Test: the plan of a split-K weight-gradient group (tc::dw_plan: macro tile and split count chosen together so that the tile list fills whole rounds
// of `sms` SMs).  Host arithmetic only — callable without a GPU.  out3 = {row blocks per tile, columns per tile, splits}.
extern "C" int grx_gemm_debug_dw_plan(const int32_t *M, const int32_t *N, int32_t K, int32_t np, int32_t sms, int32_t *out3) {
    if (!M || !N || !out3 || np < 1 || np > tc::MAXP || K < 1 || sms < 1) return grx_set_error(GRX_E_INVALID, "grx_gemm_debug_dw_plan: bad arguments");
    tc::Problem ps[tc::MAXP];
    memset(ps, 0, sizeof(ps));
    for (int i = 0; i < np; i++) { ps[i].M = M[i]; ps[i].N = N[i]; ps[i].K = K; }
    int bt = 1, bb = 128, z = 1;
    tc::dw_plan(ps, np, sms, bt, bb, z);
    out3[0] = bt; out3[1] = bb; out3[2] = z;
    return GRX_OK;
}

// Test / profiling: force the macro tile of the tensor-core GEMM ({0, 0} = cost model).  Configurations that do not apply to a launch
// (ELU' epilogue with more than 128 x 128, padding-only tiles) fall back to the cost model.
extern "C" int grx_gemm_debug_tile(int32_t row_blocks, int32_t bn) {
    tc::forced_tile()[0] = row_blocks; tc::forced_tile()[1] = bn;
    return GRX_OK;
}

// Profiling: %globaltimer stamps (ns) of CTA 0 of the most recent tensor-core GEMM launch (see tc::g_stamps).  Synchronises.
extern "C" int grx_gemm_debug_stamps(uint64_t *out8) {   /* 16 values */
    if (!out8) return grx_set_error(GRX_E_INVALID, "grx_gemm_debug_stamps: null argument");
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpyFromSymbol(out8, tc::g_stamps, 16 * sizeof(uint64_t)));
    return GRX_OK;
}
// Profiling: all 32 stamps; [16..23] = allreduce_kernel of the most recent minibatch: per phase {entry, ready barrier passed, slices reduced
// and pushed, done flags published}.  Synchronises.
extern "C" int grx_debug_stamps32(uint64_t *out32) {
    if (!out32) return grx_set_error(GRX_E_INVALID, "grx_debug_stamps32: null argument");
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpyFromSymbol(out32, tc::g_stamps, 32 * sizeof(uint64_t)));
    return GRX_OK;
}

// Debug / parity entry: one dense-layer GEMM on device pointers, through either implementation.
//   variant 0: C[M,N] = A[M,K] B[N,K]^T (+bias, epi 0/1)   1: C[M,N] = (A[M,K] B[K,N]) * ELU'(aux) (epi 2)   2: C[M,N] += A[K,M]^T B[K,N] (epi 3)
extern "C" int grx_gemm_debug(int32_t variant, int32_t epi, int32_t M, int32_t N, int32_t K, const float *A, const float *B, float *C,
                              const float *bias, const float *aux, float *bias_out, int32_t splits, int32_t use_tc, void *stream) {
    GemmArgs g; memset(&g, 0, sizeof(g));
    g.A = A; g.B = B; g.C = C; g.bias = bias; g.aux = aux; g.bias_out = bias_out; g.M = M; g.N = N; g.K = K; g.ldc = N;
    cudaStream_t st = (cudaStream_t)stream;
    if (variant == 0) { g.lda = K; g.ldb = K; if (epi == 1) dense<true, true, 1>(g, 1, use_tc != 0, st); else dense<true, true, 0>(g, 1, use_tc != 0, st); }
    else if (variant == 1) { g.lda = K; g.ldb = N; dense<true, false, 2>(g, 1, use_tc != 0, st); }
    else if (variant == 2) { g.lda = M; g.ldb = N; dense<false, false, 3>(g, splits, use_tc != 0, st); }
    else return grx_set_error(GRX_E_INVALID, "grx_gemm_debug: variant must be 0, 1 or 2");
    if (g_launch_err != cudaSuccess) { const cudaError_t e = g_launch_err; g_launch_err = cudaSuccess; return grx_set_error(GRX_E_CUDA, std::string("tensor-core GEMM launch: ") + cudaGetErrorString(e)); }
    CK(cudaGetLastError());
    return GRX_OK;
}
