// grx_gemm_tc.cuh — tcgen05 (5th-gen tensor core) TF32 GEMM for the actor / critic dense layers, sm_100a only.
//
//   C[m,n] (op)= sum_k A(m,k) B(k,n)        fp32 in HBM, TF32 multiply, fp32 accumulate in TMEM
//
// One CTA (4 warps) per 128 x BN output tile.  The contraction runs in chunks of 32: every thread stages the A / B chunk
// into shared memory with 16-byte cp.async (zero-filled at the edges) directly in the UMMA canonical no-swizzle
// "core matrix" layout, one elected thread issues 4 x tcgen05.mma.kind::tf32 (K = 8 each) per chunk and commits them to
// an mbarrier that releases the shared-memory stage; the accumulator never leaves TMEM until the epilogue reads it back
// with tcgen05.ld (32 lanes x 32 bit x 16 columns per instruction) and applies bias / ELU / ELU' / split-K reduction.
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

namespace tc {

constexpr int TM = 128, TK = 32, NTHREADS = 128;

struct Args {
    const float *A, *B;
    float *C;
    const float *bias;   // EPI 0/1: [N]
    const float *aux;    // EPI 2: same layout as C
    int M, N, K, lda, ldb, ldc;
    int BN;              // N tile (multiple of 16, <= 256)
    int kchunk;          // contraction elements per blockIdx.z (multiple of 32)
    int stages;
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
__device__ __forceinline__ float elu_f(float x) { return x > 0.f ? x : expm1f(x); }

template <bool KMAJ> __host__ __device__ constexpr uint32_t tile_bytes(int rows) {   // multiple of 1024 (swizzle atoms need 512-byte aligned tiles)
    return KMAJ ? ((8u * (uint32_t)(rows * 16 + 16) + 1023u) & ~1023u) : (uint32_t)((rows + 31) / 32 * 32) * 128u;
}

// stage one 32-deep contraction chunk of an operand (rows = tile extent along its M/N dimension)
template <bool KMAJ>
__device__ __forceinline__ void load_tile(uint32_t tile, const float *G, int ld, int row0, int rows, int rows_total, int k0, int kend, int tid) {
    if (KMAJ) {
        const uint32_t lbo = (uint32_t)(rows * 16 + 16);
        for (int idx = tid; idx < rows * 8; idx += NTHREADS) {
            const int r = idx >> 3, c = idx & 7, gk = k0 + c * 4, gr = row0 + r;
            const bool ok = gr < rows_total && gk < kend;
            const float *src = ok ? G + (size_t)gr * ld + gk : G;
            cp_async16(tile + (uint32_t)((r & 7) * 16 + (r >> 3) * 128) + (uint32_t)c * lbo, src, ok ? 16 : 0);
        }
    } else {
        const int rows32 = (rows + 31) / 32 * 32, cpr = rows32 >> 2;   // 16-byte chunks per k row (tile padded to 32 rows)
        const uint32_t sbo = (uint32_t)(rows32 / 32) * 512u;
        for (int idx = tid; idx < cpr * TK; idx += NTHREADS) {
            const int j = idx % cpr, k = idx / cpr, gk = k0 + k, gr = row0 + j * 4;
            const bool ok = gk < kend && gr < rows_total;
            const float *src = ok ? G + (size_t)gk * ld + gr : G;
            const uint32_t off = (uint32_t)(k >> 2) * sbo + (uint32_t)(j >> 3) * 512u + (uint32_t)(k & 3) * 128u +
                                 ((uint32_t)((((j & 7) >> 1) ^ (k & 3))) << 5) + (uint32_t)(j & 1) * 16u;
            cp_async16(tile + off, src, ok ? 16 : 0);
        }
    }
}

// EPI: 0 C = acc + bias[n] | 1 C = elu(acc + bias[n]) | 2 C = acc * ELU'(aux[m,n]) | 3 split-K: C += acc (red.global.add)
template <bool A_KMAJ, bool B_KMAJ, int EPI>
__global__ void __launch_bounds__(NTHREADS) gemm_tf32_kernel(const Args g) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ __align__(8) unsigned long long bars[5];   // stage-free barriers [0..3], accumulator-ready [4]
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int BN = g.BN, S = g.stages;
    const int m0 = blockIdx.y * TM, n0 = blockIdx.x * BN;
    const int kbeg = blockIdx.z * g.kchunk, kend = min(g.K, kbeg + g.kchunk);
    const int nchunks = (kend - kbeg + TK - 1) / TK;
    const uint32_t a_bytes = tile_bytes<A_KMAJ>(TM), b_bytes = tile_bytes<B_KMAJ>(BN);
    const uint32_t stage_bytes = a_bytes + b_bytes;   // both multiples of 1024
    const uint32_t smem0 = smem_u32(smem);
    const uint32_t ncols = BN <= 32 ? 32u : BN <= 64 ? 64u : BN <= 128 ? 128u : 256u;

    if (tid == 0) {
        for (int i = 0; i < 5; i++) mbar_init(smem_u32(&bars[i]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {   // TMEM allocation (power of two >= 32 columns), one warp
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(ncols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;

    // instruction descriptor: D fp32 [4,6)=1, A/B tf32 [7,10)=[10,13)=2, a_major bit 15, b_major bit 16 (1 = MN-major), N>>3 [17,23), M>>4 [24,29)
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((A_KMAJ ? 0u : 1u) << 15) | ((B_KMAJ ? 0u : 1u) << 16) |
                           ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);
    const uint32_t bn32 = (uint32_t)((BN + 31) / 32);
    const uint32_t a_lbo = A_KMAJ ? (uint32_t)(TM * 16 + 16) : 512u, a_sbo = A_KMAJ ? 128u : (uint32_t)(TM / 32) * 512u;
    const uint32_t b_lbo = B_KMAJ ? (uint32_t)(BN * 16 + 16) : 512u, b_sbo = B_KMAJ ? 128u : bn32 * 512u;
    const uint32_t a_step = A_KMAJ ? 2u * a_lbo : 2u * a_sbo, b_step = B_KMAJ ? 2u * b_lbo : 2u * b_sbo;   // advance per K = 8
    const uint32_t a_type = A_KMAJ ? 0u : 1u, b_type = B_KMAJ ? 0u : 1u;

    auto issue_load = [&](int kb) {
        const int s = kb % S;
        const uint32_t ta = smem0 + (uint32_t)s * stage_bytes, tb = ta + a_bytes;
        const int k0 = kbeg + kb * TK;
        load_tile<A_KMAJ>(ta, g.A, g.lda, m0, TM, g.M, k0, kend, tid);
        load_tile<B_KMAJ>(tb, g.B, g.ldb, n0, BN, g.N, k0, kend, tid);
    };
    // prologue: chunks 0 .. S-2
    for (int kb = 0; kb < S - 1; kb++) {
        if (kb < nchunks) issue_load(kb);
        cp_async_commit();
    }
    for (int kb = 0; kb < nchunks; kb++) {
        const int nxt = kb + S - 1;
        if (nxt < nchunks) {
            if (nxt >= S) mbar_wait(smem_u32(&bars[nxt % S]), (uint32_t)(((nxt / S) - 1) & 1));   // MMAs of chunk nxt - S released the stage
            issue_load(nxt);
        }
        cp_async_commit();
        // chunk kb has landed once at most S-1 newer groups are pending
        if (S == 2) cp_async_wait<1>(); else if (S == 3) cp_async_wait<2>(); else cp_async_wait<3>();
        fence_proxy_async();   // generic-proxy (cp.async) writes -> visible to the tensor core (async proxy)
        __syncthreads();
        if (tid == 0) {
            tc_fence_after();
            const int s = kb % S;
            const uint32_t ta = smem0 + (uint32_t)s * stage_bytes, tb = ta + a_bytes;
#pragma unroll
            for (int j = 0; j < TK / 8; j++) {
                const uint64_t da = smem_desc(ta + (uint32_t)j * a_step, a_lbo, a_sbo, a_type);
                const uint64_t db = smem_desc(tb + (uint32_t)j * b_step, b_lbo, b_sbo, b_type);
                tc_mma_tf32(tmem, da, db, idesc, (kb > 0 || j > 0) ? 1u : 0u);
            }
            tc_commit(smem_u32(&bars[s]));                       // frees stage s when these MMAs have read it
            if (kb == nchunks - 1) tc_commit(smem_u32(&bars[4]));   // accumulator complete
        }
    }
    cp_async_wait<0>();
    // ---- epilogue: warp w owns TMEM lanes [32w, 32w+32) = output rows m0 + 32w + lane
    if (nchunks > 0) mbar_wait(smem_u32(&bars[4]), 0);
    tc_fence_after();
    const int m = m0 + warp * 32 + lane;
    for (int c0 = 0; c0 < BN; c0 += 16) {
        if (n0 + c0 >= g.N) break;   // warp-uniform
        float v[16];
        if (nchunks > 0) tc_ld16(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
        else {
#pragma unroll
            for (int i = 0; i < 16; i++) v[i] = 0.f;
        }
        if (m < g.M) {
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
#pragma unroll
                    for (int i = 0; i < 16; i++) {
                        if (i < nvalid) {
                            const float x = v[i] + __ldg(g.bias + n0 + c0 + i);
                            v[i] = EPI == 1 ? elu_f(x) : x;
                        }
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
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(ncols) : "memory");
}

inline bool aligned16(const void *p) { return ((uintptr_t)p & 15u) == 0; }

// true if this problem can run on the tensor-core kernel
template <bool A_KMAJ, bool B_KMAJ>
inline bool supported(const Args &g) {
    if (!aligned16(g.A) || !aligned16(g.B) || !aligned16(g.C) || (g.lda & 3) || (g.ldb & 3) || (g.ldc & 3)) return false;
    if (A_KMAJ ? (g.K & 3) : (g.M & 3)) return false;
    if (B_KMAJ ? (g.K & 3) : (g.N & 3)) return false;
    if (g.N < 16 || g.M < 1 || g.K < 1) return false;
    if (g.aux && !aligned16(g.aux)) return false;
    return true;
}

template <bool A_KMAJ, bool B_KMAJ, int EPI>
inline cudaError_t launch(Args g, int splits, cudaStream_t st) {
    // N tile: whole N if it fits one tile (rounded up to 16), else 256 / 128 by divisibility
    int BN = g.N <= 256 ? (g.N + 15) / 16 * 16 : (g.N % 256 == 0 ? 256 : (g.N % 128 == 0 ? 128 : 256));
    g.BN = BN;
    const uint32_t stage = tile_bytes<A_KMAJ>(TM) + tile_bytes<B_KMAJ>(BN);
    g.stages = stage * 3 <= 110 * 1024 ? 3 : 2;
    if (splits < 1) splits = 1;
    g.kchunk = ((g.K + splits - 1) / splits + TK - 1) / TK * TK;
    const int z = (g.K + g.kchunk - 1) / g.kchunk;
    const size_t smem = (size_t)stage * g.stages;
    static bool attr_done = false;   // per template instantiation
    if (!attr_done) {
        cudaError_t e = cudaFuncSetAttribute(gemm_tf32_kernel<A_KMAJ, B_KMAJ, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        if (e != cudaSuccess) return e;
        attr_done = true;
    }
    dim3 grid((g.N + BN - 1) / BN, (g.M + TM - 1) / TM, z);
    gemm_tf32_kernel<A_KMAJ, B_KMAJ, EPI><<<grid, NTHREADS, smem, st>>>(g);
    return cudaGetLastError();
}

}  // namespace tc
