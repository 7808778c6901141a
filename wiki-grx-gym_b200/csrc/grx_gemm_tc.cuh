// grx_gemm_tc.cuh — tcgen05 (5th-gen tensor core) TF32 GEMM for the actor / critic dense layers, sm_100a only.
//
//   C[m,n] (op)= sum_k A(m,k) B(k,n)        fp32 in HBM, TF32 multiply, fp32 accumulate in TMEM
//
// One CTA (8 warps) per 128 x BN output tile (BN = 128 / 64 / 32, two CTAs per SM so one tile's epilogue overlaps the other's
// main loop).  The contraction runs in chunks of 32: every thread stages its fixed slots of the A / B chunk into shared
// memory with 16-byte cp.async (zero-filled at the edges) directly in the UMMA canonical layouts below, one elected thread
// issues 4 x tcgen05.mma.kind::tf32 (K = 8 each) per chunk and commits them to an mbarrier that releases the stage; the
// accumulator never leaves TMEM until the epilogue reads it back with tcgen05.ld (32 lanes x 32 bit x 16 columns per
// instruction; warp w reads lane quadrant w % 4, column half w / 4) and applies bias / ELU / ELU' / split-K reduction.
// Both operand majors are supported through the shared-memory descriptors, so the three GEMM shapes of an MLP layer
// (forward X W^T, input gradient dY W, weight gradient dY^T X) run on the same kernel without transposed copies:
//   A_KMAJ: A(m,k) = A[m*lda + k]  (contraction contiguous)   else  A[k*lda + m]
//   B_KMAJ: B(k,n) = B[n*ldb + k]                              else  B[k*ldb + n]
// Shared-memory layouts (bytes; rows = 128 for A, BN for B; one chunk = 32 k):
//   K-major : off(r,k) = (r%8)*16 + (r/8)*128 + (k/4)*LBO + (k%4)*4,  LBO = rows*16 + 16 (the +16 staggers banks), SBO = 128
//   MN-major: TF32 operands that are contiguous along M/N exist in ONE canonical form only, SWIZZLE_128B_BASE32B: atoms of
//             32 (mn) x 4 (k) elements = 4 rows of 128 B with the 32-byte units of a row XOR-ed with the row index
//             (byte address bits [5,7) ^= bits [7,9)); atoms are LBO = 512 B apart along mn and SBO = (rows/32)*512 B apart along k:
//             off(r,k) = (k/4)*SBO + (r/32)*512 + (k%4)*128 + ((((r%32)/8) ^ (k%4)) * 32) + (r%8)*4   (rows padded to 32)
// Requirements checked by the host launcher: lda/ldb/ldc % 4 == 0, 16-byte aligned bases, N-extent of MN-major operands
// and K-extent of K-major operands multiples of 4, BN % 16 == 0, 16 <= BN <= 256.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

namespace tc {

constexpr int TM = 128, TK = 32, NTHREADS = 256, NTHREADS_CTA = 288;   // 8 producer / epilogue warps + 1 MMA-issuing warp

struct Args {
    const float *A, *B;
    float *C;
    const float *bias;   // EPI 0/1: [N]
    const float *aux;    // EPI 2: same layout as C
    int M, N, K, lda, ldb, ldc;
    int kchunk;          // contraction elements per blockIdx.z (multiple of 32)
    int stages;
    int dbg;             // profiling switches (GRX_TC_DEBUG): 1 skip MMA issue, 2 skip operand loads, 4 skip epilogue stores
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cp_async16(uint32_t dst, const void *src, int src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_init(uint32_t bar, int count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok = 0;
    while (!ok) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    }
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc], kind::tf32, issued by ONE thread for the CTA
__device__ __forceinline__ void tc_mma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tc_ld16(uint32_t taddr, float *v) {
    uint32_t r[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                   "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; i++) v[i] = __uint_as_float(r[i]);
}
// shared-memory matrix descriptor, SWIZZLE_NONE, version 1 (sm_100): start[0,14) LBO[16,30) SBO[32,46) (all >> 4)
// layout type [61,64): 0 = SWIZZLE_NONE (K-major tiles here), 1 = SWIZZLE_128B_BASE32B (MN-major TF32 tiles)
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t layout_type) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46) |
           ((uint64_t)layout_type << 61);
}
// ELU on the tensor-core path: ex2.approx based exp; abs error ~1e-7, far below the TF32 operand rounding (2^-11 relative)
__device__ __forceinline__ float elu_f(float x) { return x > 0.f ? x : __expf(x) - 1.f; }

template <bool KMAJ> __host__ __device__ constexpr uint32_t tile_bytes(int rows) {   // multiple of 1024 (swizzle atoms need 512-byte aligned tiles)
    return KMAJ ? ((8u * (uint32_t)(rows * 16 + 16) + 1023u) & ~1023u) : (uint32_t)((rows + 31) / 32 * 32) * 128u;
}

// One operand's share of a 32-deep contraction chunk for this thread: CNT fixed 16-byte slots (source pointer, smem offset).
template <bool KMAJ, int ROWS>
struct Slots {
    static constexpr int CNT = ROWS * 8 / NTHREADS;
    const float *src[CNT];
    uint32_t dst[CNT];
    int kofs[CNT];      // k offset of the slot inside a chunk
    bool ok[CNT];       // row (M / N extent) in range
    size_t step;        // source advance per chunk (elements)
    const float *base;  // always-valid address for zero-filled slots
    __device__ __forceinline__ void init(const float *G, int ld, int row0, int rows_total, int kbeg, int tid) {
        base = G;
#pragma unroll
        for (int i = 0; i < CNT; i++) {
            const int idx = tid + i * NTHREADS;
            if (KMAJ) {
                const int r = idx >> 3, c = idx & 7, gr = row0 + r;
                ok[i] = gr < rows_total; kofs[i] = c * 4;
                src[i] = G + (size_t)(ok[i] ? gr : 0) * ld + kbeg + c * 4;
                dst[i] = (uint32_t)((r & 7) * 16 + (r >> 3) * 128) + (uint32_t)c * (uint32_t)(ROWS * 16 + 16);
            } else {
                constexpr int cpr = ROWS / 4;
                const int j = idx % cpr, k = idx / cpr, gr = row0 + j * 4;
                ok[i] = gr < rows_total; kofs[i] = k;
                src[i] = G + (size_t)(kbeg + k) * ld + (ok[i] ? gr : 0);
                dst[i] = (uint32_t)(k >> 2) * (uint32_t)(ROWS / 32 * 512) + (uint32_t)(j >> 3) * 512u + (uint32_t)(k & 3) * 128u +
                         ((uint32_t)((((j & 7) >> 1) ^ (k & 3))) << 5) + (uint32_t)(j & 1) * 16u;
            }
        }
        step = KMAJ ? (size_t)TK : (size_t)TK * ld;
    }
    __device__ __forceinline__ void issue(uint32_t tile, int k0, int kend) {
#pragma unroll
        for (int i = 0; i < CNT; i++) {
            const bool v = ok[i] && (k0 + kofs[i] < kend);
            cp_async16(tile + dst[i], v ? src[i] : base, v ? 16 : 0);
            src[i] += step;
        }
    }
};

// EPI: 0 C = acc + bias[n] | 1 C = elu(acc + bias[n]) | 2 C = acc * ELU'(aux[m,n]) | 3 split-K: C += acc (red.global.add)
template <bool A_KMAJ, bool B_KMAJ, int EPI, int BN, int S>
__global__ void __launch_bounds__(NTHREADS_CTA, 2) gemm_tf32_kernel(const Args g) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ __align__(8) unsigned long long full_bar[S], empty_bar[S], accum_bar;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int m0 = blockIdx.y * TM, n0 = blockIdx.x * BN;
    const int kbeg = blockIdx.z * g.kchunk, kend = min(g.K, kbeg + g.kchunk);
    const int nchunks = (kend - kbeg + TK - 1) / TK;
    constexpr uint32_t a_bytes = tile_bytes<A_KMAJ>(TM), b_bytes = tile_bytes<B_KMAJ>(BN), stage_bytes = a_bytes + b_bytes;
    const uint32_t smem0 = smem_u32(smem);

    if (tid == 0) {
        for (int i = 0; i < S; i++) { mbar_init(smem_u32(&full_bar[i]), 8); mbar_init(smem_u32(&empty_bar[i]), 1); }
        mbar_init(smem_u32(&accum_bar), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {   // TMEM allocation: BN fp32 accumulator columns (power of two >= 32), one warp
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "n"(BN) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;

    // instruction descriptor: D fp32 [4,6)=1, A/B tf32 [7,10)=[10,13)=2, a_major bit 15, b_major bit 16 (1 = MN-major), N>>3 [17,23), M>>4 [24,29)
    constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((A_KMAJ ? 0u : 1u) << 15) | ((B_KMAJ ? 0u : 1u) << 16) |
                               ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);
    constexpr uint32_t a_lbo = A_KMAJ ? (uint32_t)(TM * 16 + 16) : 512u, a_sbo = A_KMAJ ? 128u : (uint32_t)(TM / 32) * 512u;
    constexpr uint32_t b_lbo = B_KMAJ ? (uint32_t)(BN * 16 + 16) : 512u, b_sbo = B_KMAJ ? 128u : (uint32_t)(BN / 32) * 512u;
    constexpr uint32_t a_step = A_KMAJ ? 2u * a_lbo : 2u * a_sbo, b_step = B_KMAJ ? 2u * b_lbo : 2u * b_sbo;   // advance per K = 8
    constexpr uint32_t a_type = A_KMAJ ? 0u : 1u, b_type = B_KMAJ ? 0u : 1u;

    // Warp-specialised main loop, no CTA-wide barrier inside:
    //   warps 0-7 (256 threads) = producers: wait for the stage to be free, issue their cp.async slots, and let the stage's `full`
    //     mbarrier track their completion (cp.async.mbarrier.arrive.noinc) — they run up to S chunks ahead of the tensor core;
    //   warp 8, one lane     = MMA issuer: wait `full`, proxy fence, 4 x tcgen05.mma, tcgen05.commit -> `empty` (frees the stage).
    if (warp < 8) {
        Slots<A_KMAJ, TM> sa;
        Slots<B_KMAJ, BN> sb;
        sa.init(g.A, g.lda, m0, g.M, kbeg, tid);
        sb.init(g.B, g.ldb, n0, g.N, kbeg, tid);
        // Each thread keeps up to S - 1 of its own cp.async groups in flight; once the group of chunk kb - (S-1) has landed for
        // every lane of the warp, ONE lane arrives on that stage's `full` barrier (8 arrivals per stage instead of 256).
        constexpr int D = S - 1;
        for (int kb = 0; kb < nchunks + D; kb++) {
            if (kb < nchunks) {
                const int st = kb % S;
                if (kb >= S) mbar_wait(smem_u32(&empty_bar[st]), (uint32_t)(((kb / S) - 1) & 1));
                const uint32_t ta = smem0 + (uint32_t)st * stage_bytes;
                const int k0 = kbeg + kb * TK;
                if (!(g.dbg & 2)) {
                    sa.issue(ta, k0, kend);
                    sb.issue(ta + a_bytes, k0, kend);
                }
            }
            cp_async_commit();
            if (kb >= D) {
                cp_async_wait<D>();      // this thread's group for chunk kb - D is complete
                fence_proxy_async();     // its smem writes -> visible to the async proxy (tensor core)
                __syncwarp();
                if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&full_bar[(kb - D) % S])) : "memory");
            }
        }
    } else if (lane == 0) {
        for (int kb = 0; kb < nchunks; kb++) {
            const int st = kb % S;
            mbar_wait(smem_u32(&full_bar[st]), (uint32_t)((kb / S) & 1));
            fence_proxy_async();   // generic-proxy (cp.async) writes -> visible to the tensor core (async proxy)
            tc_fence_after();
            const uint32_t ta = smem0 + (uint32_t)st * stage_bytes, tb = ta + a_bytes;
#pragma unroll
            for (int j = 0; j < TK / 8; j++) {
                const uint64_t da = smem_desc(ta + (uint32_t)j * a_step, a_lbo, a_sbo, a_type);
                const uint64_t db = smem_desc(tb + (uint32_t)j * b_step, b_lbo, b_sbo, b_type);
                if (!(g.dbg & 1)) tc_mma_tf32(tmem, da, db, idesc, (kb > 0 || j > 0) ? 1u : 0u);
            }
            tc_commit(smem_u32(&empty_bar[st]));                      // frees the stage when these MMAs have read it
            if (kb == nchunks - 1) tc_commit(smem_u32(&accum_bar));   // accumulator complete
        }
    }
    if (warp < 8) {
    // ---- epilogue: warp w owns TMEM lanes [32 (w%4), +32) = output rows m0 + 32 (w%4) + lane, and column half w/4
    if (nchunks > 0) mbar_wait(smem_u32(&accum_bar), 0);
    tc_fence_after();
    const int quad = warp & 3, chalf = warp >> 2;
    const int m = m0 + quad * 32 + lane;
#pragma unroll 1
    for (int c0 = chalf * (BN / 2); c0 < (chalf + 1) * (BN / 2); c0 += 16) {
        if (n0 + c0 >= g.N) break;   // warp-uniform
        float v[16];
        if (nchunks > 0) tc_ld16(tmem + ((uint32_t)(quad * 32) << 16) + (uint32_t)c0, v);
        else {
#pragma unroll
            for (int i = 0; i < 16; i++) v[i] = 0.f;
        }
        if (m < g.M && !(g.dbg & 4)) {
            float *crow = g.C + (size_t)m * g.ldc + n0 + c0;
            const int nvalid = min(16, g.N - (n0 + c0));
            if (EPI == 3) {
                if (nvalid == 16) {
#pragma unroll
                    for (int i = 0; i < 16; i += 4)
                        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(crow + i), "f"(v[i]), "f"(v[i + 1]), "f"(v[i + 2]), "f"(v[i + 3]) : "memory");
                } else {
                    for (int i = 0; i < nvalid; i++) atomicAdd(crow + i, v[i]);
                }
            } else {
                if (EPI == 0 || EPI == 1) {
                    if (nvalid == 16) {
#pragma unroll
                        for (int i = 0; i < 16; i += 4) {
                            const float4 b = __ldg(reinterpret_cast<const float4 *>(g.bias + n0 + c0 + i));
                            v[i] += b.x; v[i + 1] += b.y; v[i + 2] += b.z; v[i + 3] += b.w;
                        }
                        if (EPI == 1) {
#pragma unroll
                            for (int i = 0; i < 16; i++) v[i] = elu_f(v[i]);
                        }
                    } else {
                        for (int i = 0; i < nvalid; i++) { const float x = v[i] + __ldg(g.bias + n0 + c0 + i); v[i] = EPI == 1 ? elu_f(x) : x; }
                    }
                } else {
                    const float *arow = g.aux + (size_t)m * g.ldc + n0 + c0;
                    if (nvalid == 16) {
#pragma unroll
                        for (int i = 0; i < 16; i += 4) {
                            const float4 h = *reinterpret_cast<const float4 *>(arow + i);
                            v[i] *= h.x > 0.f ? 1.f : h.x + 1.f; v[i + 1] *= h.y > 0.f ? 1.f : h.y + 1.f;
                            v[i + 2] *= h.z > 0.f ? 1.f : h.z + 1.f; v[i + 3] *= h.w > 0.f ? 1.f : h.w + 1.f;
                        }
                    } else {
                        for (int i = 0; i < nvalid; i++) { const float h = arow[i]; v[i] *= h > 0.f ? 1.f : h + 1.f; }
                    }
                }
                if (nvalid == 16) {
#pragma unroll
                    for (int i = 0; i < 16; i += 4) *reinterpret_cast<float4 *>(crow + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
                } else {
                    for (int i = 0; i < nvalid; i++) crow[i] = v[i];
                }
            }
        }
    }
    }   // warp < 8
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(BN) : "memory");
}

inline bool aligned16(const void *p) { return ((uintptr_t)p & 15u) == 0; }

// true if this problem can run on the tensor-core kernel
template <bool A_KMAJ, bool B_KMAJ>
inline bool supported(const Args &g) {
    if (!aligned16(g.A) || !aligned16(g.B) || !aligned16(g.C) || (g.lda & 3) || (g.ldb & 3) || (g.ldc & 3)) return false;
    if (A_KMAJ ? (g.K & 3) : (g.M & 3)) return false;
    if (B_KMAJ ? (g.K & 3) : (g.N & 3)) return false;
    if (g.N < 16 || (g.N & 3) || g.M < 1 || g.K < 1) return false;
    if (g.aux && !aligned16(g.aux)) return false;
    if (g.bias && !aligned16(g.bias)) return false;
    return true;
}

template <bool A_KMAJ, bool B_KMAJ, int EPI, int BN, int S>
inline cudaError_t launch_bn(Args g, int z, cudaStream_t st) {
    constexpr size_t smem = (size_t)(tile_bytes<A_KMAJ>(TM) + tile_bytes<B_KMAJ>(BN)) * S;
    static bool attr_done = false;   // per template instantiation
    if (!attr_done) {
        cudaError_t e = cudaFuncSetAttribute(gemm_tf32_kernel<A_KMAJ, B_KMAJ, EPI, BN, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        attr_done = true;
    }
    dim3 grid((g.N + BN - 1) / BN, (g.M + TM - 1) / TM, z);
    gemm_tf32_kernel<A_KMAJ, B_KMAJ, EPI, BN, S><<<grid, NTHREADS_CTA, smem, st>>>(g);
    return cudaGetLastError();
}

template <bool A_KMAJ, bool B_KMAJ, int EPI>
inline cudaError_t launch(Args g, int splits, cudaStream_t st) {
    if (splits < 1) splits = 1;
    { static const char *e = getenv("GRX_TC_DEBUG"); g.dbg = e ? atoi(e) : 0; }
    g.kchunk = ((g.K + splits - 1) / splits + TK - 1) / TK * TK;
    const int z = (g.K + g.kchunk - 1) / g.kchunk;
    // N tile: these GEMMs are small, parallelism first: 128 only if it divides N and still yields >= ~1 CTA per SM slot
    const long tiles_m = (g.M + TM - 1) / TM;
    if (g.N % 128 == 0 && tiles_m * (g.N / 128) * z >= 148) return launch_bn<A_KMAJ, B_KMAJ, EPI, 128, 3>(g, z, st);
    if (g.N > 32 && tiles_m * ((g.N + 63) / 64) * z >= 100) return launch_bn<A_KMAJ, B_KMAJ, EPI, 64, 4>(g, z, st);
    if (g.N > 32 && g.N % 64 == 0) return launch_bn<A_KMAJ, B_KMAJ, EPI, 64, 4>(g, z, st);
    return launch_bn<A_KMAJ, B_KMAJ, EPI, 32, 5>(g, z, st);
}

}  // namespace tc
