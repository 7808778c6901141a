// grx_gemm_tc.cuh — tcgen05 (5th-gen tensor core) TF32 GEMM for the actor / critic dense layers, sm_100a only.
//
//   C[m,n] (op)= sum_k A(m,k) B(k,n)        fp32 in HBM, TF32 multiply, fp32 accumulate in TMEM
//
// One CTA per 128 x BN output tile (BN = 128 / 64 / 32; two CTAs per SM so one tile's epilogue overlaps the other's main
// loop), 10 warps with fixed roles:
//   warp 8, one lane  TMA producer: per 32-deep contraction chunk it arms the stage's `full` mbarrier with the byte count
//                     (mbarrier.arrive.expect_tx) and issues cp.async.bulk.tensor.2d loads for the A and B boxes; TMA writes the
//                     boxes straight into the UMMA canonical shared-memory layouts and zero-fills everything out of bounds
//                     (ragged M / N / K need no predication anywhere);
//   warp 9, one lane  MMA issuer: waits `full`, issues 4 x tcgen05.mma.cta_group::1.kind::tf32 (K = 8 each) on shared-memory
//                     descriptors, then tcgen05.commit -> the stage's `empty` mbarrier (and the accumulator barrier at the end);
//   warps 0-7         epilogue: tcgen05.ld 32x32b.x16 from TMEM (warp w: lane quadrant w % 4, column half w / 4), then
//                     bias / ELU / ELU' / split-K red.global.add.v4.f32, 16-byte stores.
// Both operand majors are supported, so the three GEMM shapes of an MLP layer (forward X W^T, input gradient dY W, weight
// gradient dY^T X) run on this kernel without transposed copies:
//   A_KMAJ: A(m,k) = A[m*lda + k]  (contraction contiguous)   else  A[k*lda + m]
//   B_KMAJ: B(k,n) = B[n*ldb + k]                              else  B[k*ldb + n]
// Shared-memory layouts (one chunk = 32 k; rows = 128 for A, BN for B):
//   K-major : TMA box {32 k, rows}, CU_TENSOR_MAP_SWIZZLE_128B  == UMMA SWIZZLE_128B K-major: rows of 128 B, the 16-byte units of a
//             row XOR-ed with (row % 8); SBO = 1024 B (8 rows); one MMA (K = 8) advances the descriptor start by 32 B.
//   MN-major: TF32 operands contiguous along M/N exist in ONE canonical form, UMMA SWIZZLE_128B_BASE32B == TMA box {32 mn, 32 k},
//             CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B: k rows of 128 B (32 mn), the 32-byte units of a row XOR-ed with (row % 4);
//             one box per 32 mn columns (4 KB): atoms of 4 k rows are SBO = 512 B apart, mn blocks LBO = 4096 B apart; one MMA
//             (K = 8 = two atoms) advances the start by 1024 B.
// Requirements (checked by supported()): 16-byte aligned bases, leading dimensions % 4 == 0, N % 4 == 0.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include <mutex>
#include <unordered_map>

namespace tc {

constexpr int TM = 128, TK = 32, NTHREADS_CTA = 320;   // 8 epilogue warps + TMA warp + MMA warp

struct Args {
    float *C;
    const float *bias;   // EPI 0/1: [N]
    const float *aux;    // EPI 2: same layout as C
    float *colsum;       // EPI 2, optional: colsum[n] += sum over rows of the stored C (bias gradient of the layer below)
    int M, N, K, ldc;
    int kchunk;          // contraction elements per blockIdx.z (multiple of 32)
    int dbg;             // profiling switches (GRX_TC_DEBUG): 1 skip MMA issue, 2 skip operand loads, 4 skip epilogue stores
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, int count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok = 0;
    while (!ok) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    }
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
// TMA: 2-D tiled box load global -> shared, completion counted in bytes on an mbarrier
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *map, int c0, int c1, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(bar) : "memory");
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
// shared-memory matrix descriptor, version 1 (sm_100): start[0,14) LBO[16,30) SBO[32,46) (all >> 4), layout type [61,64):
// 2 = SWIZZLE_128B (K-major tiles), 1 = SWIZZLE_128B_BASE32B (MN-major TF32 tiles)
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t layout_type) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46) |
           ((uint64_t)layout_type << 61);
}
// ELU on the tensor-core path: ex2.approx based exp; abs error ~1e-7, far below the TF32 operand rounding (2^-11 relative)
__device__ __forceinline__ float elu_f(float x) { return x > 0.f ? x : __expf(x) - 1.f; }

__host__ __device__ constexpr uint32_t tile_bytes(int rows) { return (uint32_t)rows * 128u; }   // rows x 32 fp32, either major

// EPI: 0 C = acc + bias[n] | 1 C = elu(acc + bias[n]) | 2 C = acc * ELU'(aux[m,n]) | 3 split-K: C += acc (red.global.add)
template <bool A_KMAJ, bool B_KMAJ, int EPI, int BN, int S>
__global__ void __launch_bounds__(NTHREADS_CTA, 2) gemm_tf32_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                                                                    const Args g) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ __align__(8) unsigned long long full_bar[S], empty_bar[S], accum_bar;
    __shared__ uint32_t tmem_slot;
    __shared__ float colsum_s[BN];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int m0 = blockIdx.y * TM, n0 = blockIdx.x * BN;
    const int kbeg = blockIdx.z * g.kchunk, kend = min(g.K, kbeg + g.kchunk);
    const int nchunks = (kend - kbeg + TK - 1) / TK;
    constexpr uint32_t a_bytes = tile_bytes(TM), b_bytes = tile_bytes(BN), stage_bytes = a_bytes + b_bytes;
    const uint32_t smem0 = (smem_u32(smem) + 1023u) & ~1023u;   // swizzle atoms need 1024-byte aligned tiles

    if (tid == 0) {
        for (int i = 0; i < S; i++) { mbar_init(smem_u32(&full_bar[i]), 1); mbar_init(smem_u32(&empty_bar[i]), 1); }
        mbar_init(smem_u32(&accum_bar), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {   // TMEM allocation: BN fp32 accumulator columns (power of two >= 32), one warp
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "n"(BN) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (warp == 8 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;

    if (warp == 8) {
        if (lane == 0) {   // ---- TMA producer
            for (int kb = 0; kb < nchunks; kb++) {
                const int st = kb % S;
                if (kb >= S) mbar_wait(smem_u32(&empty_bar[st]), (uint32_t)(((kb / S) - 1) & 1));
                const uint32_t ta = smem0 + (uint32_t)st * stage_bytes, tb = ta + a_bytes, bar = smem_u32(&full_bar[st]);
                const int k0 = kbeg + kb * TK;
                if (g.dbg & 2) { mbar_arrive(bar); continue; }
                mbar_expect_tx(bar, stage_bytes);
                if (A_KMAJ) tma_load_2d(ta, &map_a, k0, m0, bar);
                else {
#pragma unroll
                    for (int j = 0; j < TM / 32; j++) tma_load_2d(ta + (uint32_t)j * 4096u, &map_a, m0 + 32 * j, k0, bar);
                }
                if (B_KMAJ) tma_load_2d(tb, &map_b, k0, n0, bar);
                else {
#pragma unroll
                    for (int j = 0; j < BN / 32; j++) tma_load_2d(tb + (uint32_t)j * 4096u, &map_b, n0 + 32 * j, k0, bar);
                }
            }
        }
    } else if (warp == 9) {
        if (lane == 0) {   // ---- MMA issuer
            // instruction descriptor: D fp32 [4,6)=1, A/B tf32 [7,10)=[10,13)=2, a_major bit 15, b_major bit 16 (1 = MN-major), N>>3 [17,23), M>>4 [24,29)
            constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((A_KMAJ ? 0u : 1u) << 15) | ((B_KMAJ ? 0u : 1u) << 16) |
                                       ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);
            constexpr uint32_t a_lbo = A_KMAJ ? 16u : 4096u, a_sbo = A_KMAJ ? 1024u : 512u, a_step = A_KMAJ ? 32u : 1024u, a_type = A_KMAJ ? 2u : 1u;
            constexpr uint32_t b_lbo = B_KMAJ ? 16u : 4096u, b_sbo = B_KMAJ ? 1024u : 512u, b_step = B_KMAJ ? 32u : 1024u, b_type = B_KMAJ ? 2u : 1u;
            for (int kb = 0; kb < nchunks; kb++) {
                const int st = kb % S;
                mbar_wait(smem_u32(&full_bar[st]), (uint32_t)((kb / S) & 1));
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
    } else {
        // ---- epilogue, two phases so that every global access is a full coalesced row segment:
        //  1. TMEM -> registers -> shared (warp w reads TMEM lanes [32 (w%4), +32) = tile rows, column half w/4; the pipeline
        //     stages are free by now and hold the 128 x BN fp32 tile with a 16-byte row pad: conflict-free float4 stores);
        //  2. each warp streams 16 tile rows shared -> global with lane = 4 consecutive columns: bias / ELU / ELU' (aux read with
        //     the same coalesced pattern) / split-K red.global.add.v4.f32.
        if (nchunks > 0) mbar_wait(smem_u32(&accum_bar), 0);
        tc_fence_after();
        constexpr int PITCH = BN + 4;   // floats
        float *tile = reinterpret_cast<float *>(smem + (smem0 - smem_u32(smem)));
        {
            const int quad = warp & 3, chalf = warp >> 2, r = quad * 32 + lane;
#pragma unroll 1
            for (int c0 = chalf * (BN / 2); c0 < (chalf + 1) * (BN / 2); c0 += 16) {
                float v[16];
                if (nchunks > 0) tc_ld16(tmem + ((uint32_t)(quad * 32) << 16) + (uint32_t)c0, v);
                else {
#pragma unroll
                    for (int i = 0; i < 16; i++) v[i] = 0.f;
                }
#pragma unroll
                for (int i = 0; i < 16; i += 4) *reinterpret_cast<float4 *>(tile + r * PITCH + c0 + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
            }
        }
        if (EPI == 2 && tid < BN) colsum_s[tid] = 0.f;
        asm volatile("bar.sync 1, 256;" ::: "memory");   // the 8 epilogue warps only
        float4 csum = make_float4(0.f, 0.f, 0.f, 0.f);    // EPI 2: a lane always owns the same 4 columns (32 % (BN/4) == 0)
        if (!(g.dbg & 4)) {
            constexpr int V4 = BN / 4;                    // float4 per tile row
#pragma unroll 1
            for (int idx = lane; idx < 16 * V4; idx += 32) {
                const int r = warp * 16 + idx / V4, c = (idx % V4) * 4;
                const int m = m0 + r, n = n0 + c;
                if (m >= g.M || n >= g.N) continue;       // N % 4 == 0: a float4 is either fully valid or fully out
                float4 v = *reinterpret_cast<const float4 *>(tile + r * PITCH + c);
                float *dst = g.C + (size_t)m * g.ldc + n;
                if (EPI == 3) {
                    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
                } else {
                    if (EPI == 0 || EPI == 1) {
                        const float4 bb = __ldg(reinterpret_cast<const float4 *>(g.bias + n));
                        v.x += bb.x; v.y += bb.y; v.z += bb.z; v.w += bb.w;
                        if (EPI == 1) { v.x = elu_f(v.x); v.y = elu_f(v.y); v.z = elu_f(v.z); v.w = elu_f(v.w); }
                    } else {
                        const float4 h = *reinterpret_cast<const float4 *>(g.aux + (size_t)m * g.ldc + n);
                        v.x *= h.x > 0.f ? 1.f : h.x + 1.f; v.y *= h.y > 0.f ? 1.f : h.y + 1.f;
                        v.z *= h.z > 0.f ? 1.f : h.z + 1.f; v.w *= h.w > 0.f ? 1.f : h.w + 1.f;
                        csum.x += v.x; csum.y += v.y; csum.z += v.z; csum.w += v.w;
                    }
                    *reinterpret_cast<float4 *>(dst) = v;
                }
            }
        }
        if (EPI == 2 && g.colsum != nullptr) {            // 8 warps -> shared -> one global atomic per column per CTA
            const int c = (lane % (BN / 4)) * 4;
            atomicAdd(&colsum_s[c], csum.x); atomicAdd(&colsum_s[c + 1], csum.y); atomicAdd(&colsum_s[c + 2], csum.z); atomicAdd(&colsum_s[c + 3], csum.w);
            asm volatile("bar.sync 1, 256;" ::: "memory");
            if (tid < BN && n0 + tid < g.N) atomicAdd(&g.colsum[n0 + tid], colsum_s[tid]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(BN) : "memory");
}

// ---------------------------------------------------------------------------------------------------------------------
// Host side: tensor-map cache + launcher
// ---------------------------------------------------------------------------------------------------------------------
struct Problem {
    const float *A, *B;
    float *C;
    const float *bias, *aux;
    float *colsum;
    int M, N, K, lda, ldb, ldc;
};

inline bool aligned16(const void *p) { return ((uintptr_t)p & 15u) == 0; }

template <bool A_KMAJ, bool B_KMAJ>
inline bool supported(const Problem &g) {
    if (!aligned16(g.A) || !aligned16(g.B) || !aligned16(g.C) || (g.lda & 3) || (g.ldb & 3) || (g.ldc & 3)) return false;
    if (g.N < 16 || (g.N & 3) || g.M < 1 || g.K < 1) return false;
    if (g.aux && !aligned16(g.aux)) return false;
    if (g.bias && !aligned16(g.bias)) return false;
    return true;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                  const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// Tensor map of a row-major fp32 matrix [rows, cols] (leading dimension ld) with box {32 (inner), box_rows}
inline cudaError_t get_tensor_map(const float *base, int rows, int cols, int ld, int box_rows, bool kmajor_use, CUtensorMap *out) {
    struct Key {
        const void *p; int rows, cols, ld, box, km;
        bool operator==(const Key &o) const { return p == o.p && rows == o.rows && cols == o.cols && ld == o.ld && box == o.box && km == o.km; }
    };
    struct Hash {
        size_t operator()(const Key &k) const {
            size_t h = (size_t)k.p;
            for (int v : {k.rows, k.cols, k.ld, k.box, k.km}) h = h * 1000003u ^ (size_t)v;
            return h;
        }
    };
    static std::unordered_map<Key, CUtensorMap, Hash> cache;
    static std::mutex mu;
    static EncodeTiledFn encode = nullptr;
    std::lock_guard<std::mutex> lock(mu);
    const Key key{base, rows, cols, ld, box_rows, kmajor_use ? 1 : 0};
    auto it = cache.find(key);
    if (it != cache.end()) { *out = it->second; return cudaSuccess; }
    if (!encode) {
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
        if (e != cudaSuccess) return e;
        if (q != cudaDriverEntryPointSuccess || !fn) return cudaErrorNotSupported;
        encode = (EncodeTiledFn)fn;
    }
    CUtensorMap m;
    const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
    const cuuint32_t box[2] = {32u, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1u, 1u};
    const CUresult r = encode(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void *)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                              kmajor_use ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return cudaErrorInvalidValue;
    cache.emplace(key, m);
    *out = m;
    return cudaSuccess;
}

template <bool A_KMAJ, bool B_KMAJ, int EPI, int BN, int S>
inline cudaError_t launch_bn(const Problem &p, Args g, int z, cudaStream_t st) {
    constexpr size_t smem = (size_t)(tile_bytes(TM) + tile_bytes(BN)) * S + 1024;
    static_assert((size_t)TM * (BN + 4) * 4 <= (size_t)(tile_bytes(TM) + tile_bytes(BN)) * S, "epilogue tile must fit in the pipeline stages");
    static bool attr_done = false;   // per template instantiation
    if (!attr_done) {
        cudaError_t e = cudaFuncSetAttribute(gemm_tf32_kernel<A_KMAJ, B_KMAJ, EPI, BN, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        attr_done = true;
    }
    CUtensorMap ma, mb;
    // K-major operand: matrix [extent, K] -> box {32 k, tile rows}.  MN-major operand: matrix [K, extent] -> box {32 mn, 32 k}.
    cudaError_t e = A_KMAJ ? get_tensor_map(p.A, p.M, p.K, p.lda, TM, true, &ma) : get_tensor_map(p.A, p.K, p.M, p.lda, 32, false, &ma);
    if (e != cudaSuccess) return e;
    e = B_KMAJ ? get_tensor_map(p.B, p.N, p.K, p.ldb, BN, true, &mb) : get_tensor_map(p.B, p.K, p.N, p.ldb, 32, false, &mb);
    if (e != cudaSuccess) return e;
    dim3 grid((p.N + BN - 1) / BN, (p.M + TM - 1) / TM, z);
    gemm_tf32_kernel<A_KMAJ, B_KMAJ, EPI, BN, S><<<grid, NTHREADS_CTA, smem, st>>>(ma, mb, g);
    return cudaGetLastError();
}

template <bool A_KMAJ, bool B_KMAJ, int EPI>
inline cudaError_t launch(const Problem &p, int splits, cudaStream_t st) {
    Args g;
    g.C = p.C; g.bias = p.bias; g.aux = p.aux; g.colsum = p.colsum; g.M = p.M; g.N = p.N; g.K = p.K; g.ldc = p.ldc;
    if (splits < 1) splits = 1;
    { static const char *e = getenv("GRX_TC_DEBUG"); g.dbg = e ? atoi(e) : 0; }
    g.kchunk = ((p.K + splits - 1) / splits + TK - 1) / TK * TK;
    const int z = (p.K + g.kchunk - 1) / g.kchunk;
    // N tile: these GEMMs are small, parallelism first: 128 only if it divides N and still yields >= ~1 CTA per SM slot
    const long tiles_m = (p.M + TM - 1) / TM;
    if (p.N % 128 == 0 && tiles_m * (p.N / 128) * z >= 148) return launch_bn<A_KMAJ, B_KMAJ, EPI, 128, 3>(p, g, z, st);
    if (p.N > 32 && tiles_m * ((p.N + 63) / 64) * z >= 100) return launch_bn<A_KMAJ, B_KMAJ, EPI, 64, 4>(p, g, z, st);
    if (p.N > 32 && p.N % 64 == 0) return launch_bn<A_KMAJ, B_KMAJ, EPI, 64, 4>(p, g, z, st);
    return launch_bn<A_KMAJ, B_KMAJ, EPI, 32, 4>(p, g, z, st);
}

}  // namespace tc
