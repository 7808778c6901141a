// grx_gemm_tc.cuh — tcgen05 (5th-gen tensor core) TF32 GEMM for the actor / critic dense layers, sm_100a only.
//
//   C[m,n] (op)= sum_k A(m,k) B(k,n)        fp32 in HBM, TF32 multiply, fp32 accumulate in TMEM
//
// PERSISTENT, GROUPED kernel: one CTA per SM walks a static round-robin list of 128 x BN output tiles that may span several
// independent problems of the same operand layout (actor + critic layer of the same depth, or all weight-gradient GEMMs of a
// backward pass), so one launch covers both networks and the pipelines never drain between tiles.  10 warps, fixed roles:
//   warp 8, one lane  TMA producer: runs ahead over the whole tile list; per 32-deep contraction chunk it arms the stage's `full`
//                     mbarrier with the byte count and issues cp.async.bulk.tensor.2d loads for the A and B boxes; TMA writes
//                     the boxes straight into the UMMA canonical shared-memory layouts and zero-fills out of bounds
//                     (ragged M / N / K need no predication anywhere);
//   warp 9, one lane  MMA issuer: waits `full`, issues 4 x tcgen05.mma.cta_group::1.kind::tf32 (K = 8 each) on shared-memory
//                     descriptors into one of TWO TMEM accumulators, tcgen05.commit -> the stage's `empty` mbarrier, and after
//                     the tile's last chunk -> that accumulator's `acc_full` mbarrier;
//   warps 0-7         epilogue of tile i while the MMA warp already works on tile i+1; the warps are independent (no block
//                     barrier): tcgen05.ld 32x32b.x32 from TMEM (lane = tile row, 32 consecutive columns = one 128-byte row
//                     segment) -> bias / ELU / ELU' (+ bias-gradient column sums by a halving butterfly) / split-K
//                     red.global.add.v4.f32 -> global; then `acc_empty` releases the accumulator.
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
#include <string.h>

#include <mutex>
#include <unordered_map>

#include "grx_count.h"

namespace tc {

constexpr int TM = 128, TK = 32, NTHREADS_CTA = 320;   // 8 epilogue warps + TMA warp + MMA warp
constexpr int MAXP = 6;                                 // problems per grouped launch (all weight gradients of actor + critic)

// profiling stamps of CTA 0 (ns, %globaltimer): [0] entry, [1] setup done, [2] first TMA issued, [3] first stage landed,
// [4] last MMA committed, [5] first accumulator complete (epilogue starts), [6] epilogue of the last tile done
__device__ unsigned long long g_stamps[32];   // [16..23]: allreduce_kernel, phase * 4 + {entry, ready barrier passed, slices reduced + pushed, done flags published}
__device__ __forceinline__ void stamp(int i) {
    if (blockIdx.x == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
        g_stamps[i] = t;
    }
}

struct Args {
    float *C;
    const float *bias;   // EPI 0/1: [N]
    const float *aux;    // EPI 2: same layout as C
    float *colsum;       // EPI 2, optional: colsum[n] += sum over rows of the stored C (bias gradient of the layer below)
    int M, N, K, ldc;
    int kchunk;          // contraction elements per split (multiple of 32)
    int nt_m, nt_n;      // tiles along M / N
    int tile_begin;      // first linear tile id of this problem in the group (tiles = nt_m * nt_n * splits)
    // layer pipelining inside ONE launch (launch_pipe): problem `dep` (an earlier member of the group, or -1) produces this problem's A operand;
    // a tile of row block tm may load A once sync[dep][tm] has reached `need` (= column tiles of dep x its epilogue warps); `signal` != 0:
    // every epilogue warp adds 1 to sync[this][tm] when its stores of a tile have completed
    int dep, need, signal;
};
constexpr int SYNC_RB = 256;                          // row-block counters per problem
constexpr int SYNC_INTS = MAXP * SYNC_RB + 4;         // + exit counter
struct Group {
    Args g[MAXP];
    int np, total_tiles;
    int dbg;   // profiling switches (GRX_TC_DEBUG): 1 skip the MMA issue, 2 skip the operand loads (results are garbage)
    int *sync; // launch_pipe: [MAXP][SYNC_RB] row-block counters + exit counter (all zero between launches: the last CTA to leave clears them), else NULL
    int *err;  // set to 3 when a dependency wait timed out (protocol error; results invalid)
};
struct Maps {
    CUtensorMap a[MAXP], b[MAXP];
    CUtensorMap c[MAXP];     // output, box {32 columns, 32 rows}, SWIZZLE_128B (TMA store / reduce-add)
    CUtensorMap aux[MAXP];   // EPI 2: the activation the gradient is masked with, same box
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
// Programmatic dependent launch: a kernel launched with the programmatic-serialization attribute may start while its predecessor in
// the stream / graph is still draining; everything before pdl_wait() (barrier init, TMEM allocation, descriptor prefetch) overlaps
// the predecessor's tail, everything after it sees the predecessor's memory.  Both are no-ops for a normal launch.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
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
// TMA: 2-D box store / reduce-add shared -> global (bulk async-group completion); out-of-bounds rows / columns are clipped
__device__ __forceinline__ void tma_store_2d(const CUtensorMap *map, int c0, int c1, uint32_t src) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];" ::"l"(map), "r"(c0), "r"(c1), "r"(src) : "memory");
}
__device__ __forceinline__ void tma_reduce_add_2d(const CUtensorMap *map, int c0, int c1, uint32_t src) {
    asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%1, %2}], [%3];" ::"l"(map), "r"(c0), "r"(c1), "r"(src) : "memory");
}
__device__ __forceinline__ void tma_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
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
__device__ __forceinline__ void tc_ld32(uint32_t taddr, float *v) {
    uint32_t r[32];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                 "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                   "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
                   "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
                   "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; i++) v[i] = __uint_as_float(r[i]);
}
// Column sums over the 32 lanes of a warp of CW per-lane values (lane = tile row) by a halving butterfly: CW - 1 (+1) shuffles
// instead of 5 CW; afterwards lane l (l < CW, and its mirror l + 16 when CW == 16) holds the total of column l in v[0].
template <int CW>
__device__ __forceinline__ void warp_colsum(float *v, int lane) {
#pragma unroll
    for (int w = CW / 2, bit = CW / 2; w >= 1; w >>= 1, bit >>= 1) {
        const bool up = (lane & bit) != 0;
#pragma unroll
        for (int i = 0; i < w; i++) {
            const float send = up ? v[i] : v[i + w], keep = up ? v[i + w] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, bit);
        }
    }
    if (CW == 16) v[0] += __shfl_xor_sync(0xffffffffu, v[0], 16);
}
// shared-memory matrix descriptor, version 1 (sm_100): start[0,14) LBO[16,30) SBO[32,46) (all >> 4), layout type [61,64):
// 2 = SWIZZLE_128B (K-major tiles), 1 = SWIZZLE_128B_BASE32B (MN-major TF32 tiles)
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t layout_type) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46) |
           ((uint64_t)layout_type << 61);
}
// ELU on the tensor-core path: branch-free ex2.approx.ftz (no denormal fix-up code, no predicate chains); abs error ~1e-7, far
// below the TF32 operand rounding (2^-11 relative).  For x > 0 the (possibly infinite) exponential is discarded by the select.
__device__ __forceinline__ float elu_f(float x) {
    float e;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * 1.4426950408889634f));
    return x > 0.f ? x : e - 1.f;
}

__host__ __device__ constexpr uint32_t tile_bytes(int rows) { return (uint32_t)rows * 128u; }   // rows x 32 fp32, either major

// EPI: 0 C = acc + bias[n] | 1 C = elu(acc + bias[n]) | 2 C = acc * ELU'(aux[m,n]) | 3 split-K: C += acc (TMA reduce-add)
// Macro tile: TMT (1 or 2) row blocks of 128 x BN columns per CTA tile; all row blocks share the B tile of a stage, so operand
// traffic per output element drops from 1/128 + 1/BN to 1/(128 TMT) + 1/BN.
template <int TMT, int BN>
struct TileCfg {
    static constexpr int NACC = 2 * TMT * BN <= 512 ? 2 : 1;         // accumulator sets in TMEM (2: epilogue overlaps the next tile)
    static constexpr int ACC_COLS = TMT * BN;
    static constexpr int EPW = BN >= 64 ? 8 : 4;                      // epilogue warps that own columns
    static constexpr int WCOLS = BN >= 64 ? BN / 2 : 32;              // columns per epilogue warp
    static constexpr int NCHW = WCOLS / 32;                           // 32-column chunks per warp and row block
    static constexpr int CT = TMT * NCHW;                             // chunks per warp and tile
    static constexpr int NBOX = CT < 2 ? CT : 2;                      // staging boxes per warp (ring when CT > NBOX)
    static constexpr uint32_t a_bytes = (uint32_t)TMT * 16384u, b_bytes = (uint32_t)BN * 128u, stage_bytes = a_bytes + b_bytes;
    static constexpr uint32_t staging_bytes = (uint32_t)EPW * NBOX * 4096u, bias_bytes = (uint32_t)EPW * WCOLS * 4u;
    static constexpr size_t smem(int S) { return (size_t)stage_bytes * S + staging_bytes + bias_bytes + 1024; }
};

template <bool A_KMAJ, bool B_KMAJ, int EPI, int TMT, int BN, int S>
__global__ void __launch_bounds__(NTHREADS_CTA, 1) gemm_tf32_kernel(const __grid_constant__ Maps maps, const __grid_constant__ Group grp) {
    using Cfg = TileCfg<TMT, BN>;
    constexpr int NACC = Cfg::NACC, ACC_COLS = Cfg::ACC_COLS, EPW = Cfg::EPW, WCOLS = Cfg::WCOLS, NCHW = Cfg::NCHW, CT = Cfg::CT, NBOX = Cfg::NBOX;
    static_assert(EPI != 2 || CT <= NBOX, "EPI 2 prefetches the activation boxes of a whole tile: needs one staging box per chunk");
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ __align__(8) unsigned long long full_bar[S], empty_bar[S], acc_full[2], acc_empty[2], aux_bar[8][2];
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) stamp(0);
    constexpr uint32_t a_bytes = Cfg::a_bytes, stage_bytes = Cfg::stage_bytes;
    const uint32_t smem0 = (smem_u32(smem) + 1023u) & ~1023u;   // swizzle atoms need 1024-byte aligned tiles

    if (tid == 0) {
        for (int i = 0; i < S; i++) { mbar_init(smem_u32(&full_bar[i]), 1); mbar_init(smem_u32(&empty_bar[i]), 1); }
        for (int i = 0; i < 2; i++) { mbar_init(smem_u32(&acc_full[i]), 1); mbar_init(smem_u32(&acc_empty[i]), EPW); }
        for (int i = 0; i < 16; i++) mbar_init(smem_u32(&aux_bar[i >> 1][i & 1]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {   // TMEM: NACC accumulator sets of TMT x BN fp32 columns (power of two >= 32 columns), allocated by one warp
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "n"(NACC * ACC_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (warp == 8 && lane == 0) {
        for (int p = 0; p < grp.np; p++) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.a[p]) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.b[p]) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.c[p]) : "memory");
            if (EPI == 2) asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.aux[p]) : "memory");
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    pdl_wait();                 // the set-up above overlapped the previous kernel's tail; its results are visible from here on
    pdl_launch_dependents();    // let the next kernel's CTAs queue up behind ours (they block in their own pdl_wait)
    if (tid == 0) stamp(1);

    // tile id -> (problem, m0, n0, contraction range); every role walks the same sequence t = blockIdx.x, + gridDim.x, ...
    struct Tile { int p, m0, n0, kbeg, nchunks, tm; };
    auto decode = [&](int t) {
        Tile T;
        int p = 0;
        while (p + 1 < grp.np && t >= grp.g[p + 1].tile_begin) p++;
        const Args &g = grp.g[p];
        const int local = t - g.tile_begin;
        const int tn = local % g.nt_n, rest = local / g.nt_n, tm = rest % g.nt_m, z = rest / g.nt_m;
        T.p = p; T.m0 = tm * (TM * TMT); T.n0 = tn * BN; T.kbeg = z * g.kchunk; T.tm = tm;
        const int kend = min(g.K, T.kbeg + g.kchunk);
        T.nchunks = (kend - T.kbeg + TK - 1) / TK;
        return T;
    };

    if (warp == 8) {
        if (lane == 0) {   // ---- TMA producer
            uint32_t it = 0;   // chunk counter across tiles -> stage / phase
            for (int t = blockIdx.x; t < grp.total_tiles; t += gridDim.x) {
                const Tile T = decode(t);
                const CUtensorMap *ma = &maps.a[T.p], *mb = &maps.b[T.p];
                if (grp.g[T.p].dep >= 0) {   // layer pipelining: the rows of A this tile reads are outputs of an earlier problem of this launch
                    const int *cnt = grp.sync + grp.g[T.p].dep * SYNC_RB + T.tm;
                    const int need = grp.g[T.p].need;
                    const long long t0 = clock64();
                    for (;;) {
                        int v;
                        asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(cnt) : "memory");
                        if (v >= need) break;
                        if (clock64() - t0 > 6000000000ll) { atomicExch(grp.err, 3); break; }   // never hang the GPU; the host raises on the flag
                        __nanosleep(32);
                    }
                    asm volatile("fence.proxy.async;" ::: "memory");   // the producers' stores (seen through the acquire) before this thread's TMA reads
                }
                for (int kb = 0; kb < T.nchunks; kb++, it++) {
                    const uint32_t st = it % S;
                    if (it >= S) mbar_wait(smem_u32(&empty_bar[st]), ((it / S) - 1) & 1);
                    const uint32_t ta = smem0 + st * stage_bytes, tb = ta + a_bytes, bar = smem_u32(&full_bar[st]);
                    const int k0 = T.kbeg + kb * TK;
                    if (grp.dbg & 2) { mbar_arrive(bar); continue; }
                    mbar_expect_tx(bar, stage_bytes);
                    if (A_KMAJ) {
#pragma unroll
                        for (int i = 0; i < TMT; i++) tma_load_2d(ta + (uint32_t)i * 16384u, ma, k0, T.m0 + TM * i, bar);
                    } else {
#pragma unroll
                        for (int j = 0; j < TMT * (TM / 32); j++) tma_load_2d(ta + (uint32_t)j * 4096u, ma, T.m0 + 32 * j, k0, bar);
                    }
                    if (B_KMAJ) tma_load_2d(tb, mb, k0, T.n0, bar);
                    else {
#pragma unroll
                        for (int j = 0; j < BN / 32; j++) tma_load_2d(tb + (uint32_t)j * 4096u, mb, T.n0 + 32 * j, k0, bar);
                    }
                    if (it == 0) stamp(2);
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
            uint32_t it = 0, ti = 0;
            for (int t = blockIdx.x; t < grp.total_tiles; t += gridDim.x, ti++) {
                const Tile T = decode(t);
                const uint32_t buf = ti % NACC;
                if (ti >= (uint32_t)NACC) { mbar_wait(smem_u32(&acc_empty[buf]), ((ti / NACC) - 1) & 1); tc_fence_after(); }   // epilogue drained this set
                const uint32_t dacc = tmem + buf * (uint32_t)ACC_COLS;
                for (int kb = 0; kb < T.nchunks; kb++, it++) {
                    const uint32_t st = it % S;
                    mbar_wait(smem_u32(&full_bar[st]), (it / S) & 1);
                    tc_fence_after();
                    if (it == 0) stamp(3);
                    const uint32_t ta = smem0 + st * stage_bytes, tb = ta + a_bytes;
#pragma unroll
                    for (int i = 0; i < TMT; i++) {
#pragma unroll
                        for (int j = 0; j < TK / 8; j++) {
                            const uint64_t da = smem_desc(ta + (uint32_t)i * 16384u + (uint32_t)j * a_step, a_lbo, a_sbo, a_type);
                            const uint64_t db = smem_desc(tb + (uint32_t)j * b_step, b_lbo, b_sbo, b_type);
                            if (!(grp.dbg & 1)) tc_mma_tf32(dacc + (uint32_t)(i * BN), da, db, idesc, (kb > 0 || j > 0) ? 1u : 0u);
                        }
                    }
                    tc_commit(smem_u32(&empty_bar[st]));                          // frees the stage when these MMAs have read it
                    if (kb == T.nchunks - 1) tc_commit(smem_u32(&acc_full[buf]));   // accumulator set complete
                }
            }
            stamp(4);
        }
    } else {
        // ---- epilogue (independent warps, no block-level synchronisation): warp w owns TMEM lanes [32 (w%4), +32) = rows of every
        // row block and the column half w/4.  Per 32-column chunk: tcgen05.ld 32x32b.x32 -> registers (lane = row) -> bias / ELU /
        // ELU' -> one of the warp's 4 KB staging boxes in shared memory (128-byte rows, 16-byte units XOR-swizzled with row % 8:
        // conflict-free and exactly the SWIZZLE_128B box layout) -> ONE TMA store (or TMA reduce-add for split-K) per chunk;
        // ragged edges are clipped by the TMA unit.  The boxes form a ring (a box is rewritten once the bulk group that read it
        // has drained).  Bias slice of the tile: staged per warp in shared memory while the accumulator is still being produced.
        // EPI 2: the activation boxes are TMA-loaded into the staging boxes at tile start and overwritten in place; bias-gradient
        // column sums by a halving butterfly.
        if (warp < EPW) {
            const int quad = warp & 3, chalf = warp >> 2;
            const uint32_t stg0 = smem0 + (uint32_t)S * stage_bytes + (uint32_t)warp * (NBOX * 4096u);
            const uint32_t bias0 = smem0 + (uint32_t)S * stage_bytes + Cfg::staging_bytes + (uint32_t)warp * (WCOLS * 4u);
            const uint32_t rowoff = (uint32_t)lane * 128u, sw = (uint32_t)(lane & 7);
            uint32_t ti = 0, cc = 0;   // tiles / chunks processed by this warp
            // layer pipelining: a finished tile is SIGNALLED (its row-block counter incremented) only once its bulk stores have completed.  Waiting
            // for that right after the tile would put one store latency per tile on the epilogue's critical path, so the signal lags: it is sent after
            // the first chunk of the warp's next tile has been committed (wait_group 1: everything but the newest group has completed) — or at once
            // when that next tile's accumulator is not ready yet (the warp would idle anyway; this also keeps the protocol free of cycles: a
            // pending signal never waits on a tile that may itself depend on it).
            int *pend = nullptr;
            auto flush_pending = [&](bool all) {
                if (pend != nullptr && lane == 0) {
                    if (all) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
                    else asm volatile("cp.async.bulk.wait_group 1;" ::: "memory");
                    asm volatile("fence.proxy.async;" ::: "memory");
                    __threadfence();
                    asm volatile("red.release.gpu.global.add.s32 [%0], %1;" ::"l"(pend), "r"(1) : "memory");
                }
                pend = nullptr;
            };
            for (int t = blockIdx.x; t < grp.total_tiles; t += gridDim.x, ti++) {
                const Tile T = decode(t);
                const Args &g = grp.g[T.p];
                const uint32_t buf = ti % NACC;
                if (pend != nullptr) {   // accumulator of this tile not complete yet -> use the idle time to publish the previous tile
                    uint32_t ready = 0;
                    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                                 : "=r"(ready) : "r"(smem_u32(&acc_full[buf])), "r"((ti / NACC) & 1) : "memory");
                    if (!ready) flush_pending(true);
                }
                const int ncol0 = T.n0 + chalf * WCOLS;
                if (EPI == 2) {
                    if (lane == 0) {
                        tma_wait_read0();   // the stores of the previous tile have finished reading the staging boxes
#pragma unroll
                        for (int c = 0; c < CT; c++) {
                            const uint32_t bar = smem_u32(&aux_bar[warp][c]);
                            mbar_expect_tx(bar, 4096u);
                            tma_load_2d(stg0 + c * 4096u, &maps.aux[T.p], ncol0 + (c % NCHW) * 32, T.m0 + (c / NCHW) * TM + quad * 32, bar);
                        }
                    }
                } else if (EPI == 0 || EPI == 1) {
                    if (lane < WCOLS / 4) {
                        const int nn = ncol0 + 4 * lane;
                        const float4 bb = nn < g.N ? __ldg(reinterpret_cast<const float4 *>(g.bias + nn)) : make_float4(0.f, 0.f, 0.f, 0.f);
                        asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(bias0 + 16u * lane), "f"(bb.x), "f"(bb.y), "f"(bb.z), "f"(bb.w) : "memory");
                    }
                }
                __syncwarp();
                mbar_wait(smem_u32(&acc_full[buf]), (ti / NACC) & 1);
                tc_fence_after();
                if (ti == 0 && tid == 0) stamp(5);
#pragma unroll 1
                for (int c = 0; c < CT; c++, cc++) {
                    const int rb = c / NCHW, ch = c % NCHW;
                    const int n = ncol0 + ch * 32, mrow0 = T.m0 + rb * TM + quad * 32;
                    const uint32_t bx = CT <= NBOX ? (uint32_t)c : cc % NBOX;
                    float v[32];
                    tc_ld32(tmem + ((uint32_t)(quad * 32) << 16) + buf * (uint32_t)ACC_COLS + (uint32_t)(rb * BN + chalf * WCOLS + ch * 32), v);
                    if (ti == 0 && tid == 0 && c == 0) stamp(8);
                    if (c == CT - 1) {   // this warp's last TMEM read of the accumulator set: hand it back to the MMA warp
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(smem_u32(&acc_empty[buf]));
                    }
                    if (EPI != 2) {      // ring: the bulk group that read this box (NBOX chunks ago) must have drained
                        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(NBOX - 1) : "memory");
                        __syncwarp();
                    }
                    const uint32_t box = stg0 + bx * 4096u + rowoff;
                    if (EPI == 2) mbar_wait(smem_u32(&aux_bar[warp][c]), ti & 1);
                    // pass 1: all shared loads of the chunk (bias slice / activation box) are issued back to back; interleaving them
                    // with the volatile shared stores below would serialise one latency per float4
                    float4 q[8];
                    if (EPI == 0 || EPI == 1) {
#pragma unroll
                        for (int i = 0; i < 8; i++)
                            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(q[i].x), "=f"(q[i].y), "=f"(q[i].z), "=f"(q[i].w) : "r"(bias0 + (uint32_t)(ch * 128 + i * 16)));
                    } else if (EPI == 2) {
#pragma unroll
                        for (int i = 0; i < 8; i++) {
                            const uint32_t addr = box + (((uint32_t)i ^ sw) << 4);
                            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(q[i].x), "=f"(q[i].y), "=f"(q[i].z), "=f"(q[i].w) : "r"(addr));
                        }
                    }
#pragma unroll
                    for (int i = 0; i < 32; i += 4) {
                        const uint32_t addr = box + ((((uint32_t)i >> 2) ^ sw) << 4);
                        float4 x = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
                        if (EPI == 0 || EPI == 1) {
                            const float4 bb = q[i >> 2];
                            x.x += bb.x; x.y += bb.y; x.z += bb.z; x.w += bb.w;
                            if (EPI == 1) { x.x = elu_f(x.x); x.y = elu_f(x.y); x.z = elu_f(x.z); x.w = elu_f(x.w); }
                        } else if (EPI == 2) {
                            const float4 h = q[i >> 2];
                            x.x *= h.x > 0.f ? 1.f : h.x + 1.f; x.y *= h.y > 0.f ? 1.f : h.y + 1.f;     // out-of-bounds box elements were zero-filled:
                            x.z *= h.z > 0.f ? 1.f : h.z + 1.f; x.w *= h.w > 0.f ? 1.f : h.w + 1.f;     // the accumulator is zero there too (operands zero-filled)
                            v[i] = x.x; v[i + 1] = x.y; v[i + 2] = x.z; v[i + 3] = x.w;
                        }
                        asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(x.x), "f"(x.y), "f"(x.z), "f"(x.w) : "memory");
                    }
                    if (ti == 0 && tid == 0 && c == 0) stamp(9);
                    fence_proxy_async();
                    __syncwarp();
                    if (ti == 0 && tid == 0 && c == 0) stamp(10);
                    if (lane == 0) {
                        if (EPI == 3) tma_reduce_add_2d(&maps.c[T.p], n, mrow0, stg0 + bx * 4096u);
                        else tma_store_2d(&maps.c[T.p], n, mrow0, stg0 + bx * 4096u);
                        tma_commit();
                        if (ti == 0 && tid == 0 && c == 0) stamp(11);
                    }
                    if (EPI == 2 && g.colsum != nullptr) {
                        warp_colsum<32>(v, lane);
                        if (n + lane < g.N) atomicAdd(&g.colsum[n + lane], v[0]);
                    }
                    if (c == 0 && pend != nullptr) flush_pending(false);   // the previous tile's groups are all older than the one just committed
                }
                if (g.signal) pend = grp.sync + T.p * SYNC_RB + T.tm;
            }
            flush_pending(true);
            if (ti >= 1 && tid == 0) stamp(12);
            if (lane == 0) tma_wait_read0();
        }
        if (tid == 0) stamp(6);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(NACC * ACC_COLS) : "memory");
    if (grp.sync != nullptr) {   // layer pipelining: the last CTA to leave clears the counters for the next launch (which touches them only after its pdl_wait)
        __shared__ int s_last;
        if (tid == 0) {
            __threadfence();
            s_last = atomicAdd(reinterpret_cast<unsigned *>(grp.sync + MAXP * SYNC_RB), 1u) == gridDim.x - 1 ? 1 : 0;
        }
        __syncthreads();
        if (s_last) {
            for (int i = tid; i < grp.np * SYNC_RB; i += NTHREADS_CTA) grp.sync[i] = 0;
            if (tid == 0) grp.sync[MAXP * SYNC_RB] = 0;
        }
    }
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

// test / profiling override of the macro-tile choice: {row blocks, BN}, {0, 0} = cost model (GRX_TC_TILE=2x256 sets it at load)
inline int *forced_tile() {
    static int t[2] = {0, 0};
    static bool init = false;
    if (!init) {
        init = true;
        if (const char *e = getenv("GRX_TC_TILE")) { t[0] = atoi(e); const char *x = strchr(e, 'x'); t[1] = x ? atoi(x + 1) : 0; }
    }
    return t;
}
// macro tile of the split-K weight-gradient groups (EPI 3), {0, 0} = cost model; GRX_DW_TILE=2x256 sets it at load
inline int *dw_tile() {
    static int t[2] = {0, 0};
    static bool init = false;
    if (!init) {
        init = true;
        if (const char *e = getenv("GRX_DW_TILE")) { t[0] = atoi(e); const char *x = strchr(e, 'x'); t[1] = x ? atoi(x + 1) : 0; }
    }
    return t;
}
// cudaLaunchKernelEx wrapper; pdl = launch with the programmatic stream serialization attribute (see pdl_wait)
inline bool pdl_enabled() {
    static const int on = [] { const char *e = getenv("GRX_PDL"); return e ? atoi(e) : 1; }();
    return on != 0;
}
template <typename... KArgs, typename... Args>
inline cudaError_t launch_kernel(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, bool pdl, Args... args) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = pdl && pdl_enabled() ? 1 : 0;
    grx_count_launch();
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}
inline int sm_count() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    }
    return n;
}

template <bool A_KMAJ, bool B_KMAJ, int EPI, int TMT, int BN, int S>
inline cudaError_t launch_cfg(const Problem *ps, int np, const int *splits, cudaStream_t st, const int *deps = nullptr, int *sync = nullptr, int *err = nullptr) {
    constexpr size_t smem = TileCfg<TMT, BN>::smem(S);
    static_assert(smem <= 226 * 1024, "shared memory budget (227 KB per CTA minus the static barriers)");
    static bool attr_done = false;   // per template instantiation
    if (!attr_done) {
        cudaError_t e = cudaFuncSetAttribute(gemm_tf32_kernel<A_KMAJ, B_KMAJ, EPI, TMT, BN, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        attr_done = true;
    }
    Maps maps;
    Group grp;
    memset(&grp, 0, sizeof(grp));
    grp.np = np;
    int total = 0;
    for (int i = 0; i < np; i++) {
        const Problem &p = ps[i];
        // K-major operand: matrix [extent, K] -> box {32 k, tile rows}.  MN-major operand: matrix [K, extent] -> box {32 mn, 32 k}.
        cudaError_t e = A_KMAJ ? get_tensor_map(p.A, p.M, p.K, p.lda, TM, true, &maps.a[i]) : get_tensor_map(p.A, p.K, p.M, p.lda, 32, false, &maps.a[i]);
        if (e != cudaSuccess) return e;
        e = B_KMAJ ? get_tensor_map(p.B, p.N, p.K, p.ldb, BN, true, &maps.b[i]) : get_tensor_map(p.B, p.K, p.N, p.ldb, 32, false, &maps.b[i]);
        if (e != cudaSuccess) return e;
        e = get_tensor_map(p.C, p.M, p.N, p.ldc, 32, true, &maps.c[i]);
        if (e != cudaSuccess) return e;
        if (EPI == 2) {
            e = get_tensor_map(p.aux, p.M, p.N, p.ldc, 32, true, &maps.aux[i]);
            if (e != cudaSuccess) return e;
        } else maps.aux[i] = maps.c[i];
        Args &g = grp.g[i];
        g.C = p.C; g.bias = p.bias; g.aux = p.aux; g.colsum = p.colsum; g.M = p.M; g.N = p.N; g.K = p.K; g.ldc = p.ldc;
        const int sp = splits && splits[i] > 1 ? splits[i] : 1;
        g.kchunk = ((p.K + sp - 1) / sp + TK - 1) / TK * TK;
        const int z = (p.K + g.kchunk - 1) / g.kchunk;
        g.nt_m = (p.M + TM * TMT - 1) / (TM * TMT); g.nt_n = (p.N + BN - 1) / BN;
        g.tile_begin = total;
        total += g.nt_m * g.nt_n * z;
        g.dep = -1; g.need = 0; g.signal = 0;
    }
    grp.sync = nullptr; grp.err = err;
    if (deps != nullptr && sync != nullptr) {   // layer pipelining (launch_pipe): deps[i] < i
        grp.sync = sync;
        for (int i = 0; i < np; i++) {
            if (deps[i] < 0) continue;
            if (deps[i] >= i || grp.g[i].nt_m != grp.g[deps[i]].nt_m || grp.g[i].nt_m > SYNC_RB || (splits && splits[i] > 1)) return cudaErrorInvalidValue;
            grp.g[i].dep = deps[i];
            grp.g[i].need = grp.g[deps[i]].nt_n * TileCfg<TMT, BN>::EPW;
            grp.g[deps[i]].signal = 1;
        }
    }
    for (int i = np; i < MAXP; i++) { maps.a[i] = maps.a[0]; maps.b[i] = maps.b[0]; maps.c[i] = maps.c[0]; maps.aux[i] = maps.aux[0]; grp.g[i].tile_begin = total; grp.g[i].dep = -1; }
    grp.total_tiles = total;
    { static const char *e = getenv("GRX_TC_DEBUG"); grp.dbg = e ? atoi(e) : 0; }
    if (total == 0) return cudaSuccess;
    const int grid = total < sm_count() ? total : sm_count();
    return launch_kernel(gemm_tf32_kernel<A_KMAJ, B_KMAJ, EPI, TMT, BN, S>, dim3(grid), dim3(NTHREADS_CTA), smem, st, true, maps, grp);
}

// One launch for `np` (<= MAXP) problems of the same operand layout and epilogue.  The macro tile is picked by a small cost model:
// these GEMMs are bound by operand bytes pulled into each SM (~70 B/ns per SM through TMA, measured) and the tile list is static
// round-robin, so cost = rounds x (operand bytes of one tile / 70 B/ns + the epilogue when it cannot overlap the next tile).
// Split-K weight-gradient groups (EPI 3): macro tile AND split count chosen together so that the tile list fills whole rounds of the SM count.
// These launches are bound by the operand bytes every SM pulls in (rounds x (128 TMT + BN) x rows per split x 4 B at ~70 B/ns) plus the
// reduce-add epilogues (one per split), and a tile list of 2.5 rounds costs 3.  Measured on the registered minibatch (10 485 rows, six gradients
// in one launch, profiles/r3a_dw_tile_experiments.txt): 1 x 128 tiles x 512-row splits 175.6 us per minibatch, 2 x 256 x 768-row splits
// (140 tiles = one round) 167.3 us.  GRX_DW_AUTO=0 keeps the caller's split count and the generic cost model.
inline bool dw_auto() {
    static const int on = [] { const char *e = getenv("GRX_DW_AUTO"); return e ? atoi(e) : 1; }();
    return on != 0;
}
inline void dw_plan(const Problem *ps, int np, int sms, int &bt, int &bb, int &z_out) {
    struct Cand { int tmt, bn; };
    static const Cand cands[4] = {{2, 256}, {2, 128}, {1, 256}, {1, 128}};
    double best = -1.0;
    int K = 0;
    for (int i = 0; i < np; i++) K = ps[i].K > K ? ps[i].K : K;
    for (int ci = 0; ci < 4; ci++) {
        const int tmt = cands[ci].tmt, bn = cands[ci].bn;
        long tiles_mn = 0;
        for (int i = 0; i < np; i++) tiles_mn += (long)((ps[i].M + TM * tmt - 1) / (TM * tmt)) * ((ps[i].N + bn - 1) / bn);
        const int ct = tmt * (bn / 64);
        const bool overlap = 2 * tmt * bn <= 512;
        for (int rounds = 1; rounds <= 4; rounds++) {
            int z = (int)((long)rounds * sms / tiles_mn);
            if (z < 1) continue;
            int kchunk = ((K + z - 1) / z + TK - 1) / TK * TK;
            if (kchunk < 4 * TK) kchunk = 4 * TK;
            z = (K + kchunk - 1) / kchunk;
            const long r = (tiles_mn * z + sms - 1) / sms;
            const double t_tile = (double)(TM * tmt + bn) * kchunk * 4.0 / 70.0 + (overlap ? 100.0 : 190.0 * ct);
            const double cost = r * t_tile + 190.0 * ct + 40.0 * z;   // + the reduce-add traffic grows with the number of splits
            if (best < 0 || cost < best) { best = cost; bt = tmt; bb = bn; z_out = z; }
        }
    }
}

template <bool A_KMAJ, bool B_KMAJ, int EPI>
inline cudaError_t launch_group(const Problem *ps, int np, const int *splits, cudaStream_t st) {
    if (np < 1 || np > MAXP) return cudaErrorInvalidValue;
    const int sms = sm_count();
    if (EPI == 3 && splits != nullptr && dw_auto() && dw_tile()[0] == 0 && forced_tile()[0] == 0) {
        int bt = 1, bb = 128, z = 1, sp[MAXP];
        dw_plan(ps, np, sms, bt, bb, z);
        for (int i = 0; i < np; i++) sp[i] = z;
        if (bt == 2 && bb == 256) return launch_cfg<A_KMAJ, B_KMAJ, EPI == 3 ? 3 : 0, 2, 256, 2>(ps, np, sp, st);
        if (bt == 2 && bb == 128) return launch_cfg<A_KMAJ, B_KMAJ, EPI == 3 ? 3 : 0, 2, 128, 3>(ps, np, sp, st);
        if (bt == 1 && bb == 256) return launch_cfg<A_KMAJ, B_KMAJ, EPI == 3 ? 3 : 0, 1, 256, 3>(ps, np, sp, st);
        return launch_cfg<A_KMAJ, B_KMAJ, EPI == 3 ? 3 : 0, 1, 128, 4>(ps, np, sp, st);
    }
    struct Cand { int tmt, bn; };
    static const Cand cands[6] = {{2, 256}, {2, 128}, {1, 256}, {1, 128}, {1, 64}, {1, 32}};
    double best = -1.0;
    int bt = 1, bb = 32;
    for (int ci = 0; ci < 6; ci++) {
        const int tmt = cands[ci].tmt, bn = cands[ci].bn;
        if (EPI == 2 && (tmt > 1 || bn > 128)) continue;   // ELU' needs one staging box per chunk of the tile
        long tiles = 0;
        double bytes = 0.0;
        bool ok = true;
        for (int i = 0; i < np; i++) {
            const int sp = splits && splits[i] > 1 ? splits[i] : 1;
            const int kchunk = ((ps[i].K + sp - 1) / sp + TK - 1) / TK * TK;
            const int z = (ps[i].K + kchunk - 1) / kchunk;
            if (bn > 32 && ps[i].N <= bn / 2 && !(EPI == 3 && np > 4 && bn == 128)) ok = false;   // mostly padding along N (tolerated for one narrow member of a merged weight-gradient group)
            if (tmt > 1 && ps[i].M <= TM) ok = false;                      // second row block would be padding
            const long ti = (long)((ps[i].M + TM * tmt - 1) / (TM * tmt)) * ((ps[i].N + bn - 1) / bn) * z;
            tiles += ti;
            bytes += (double)ti * (TM * tmt + bn) * kchunk * 4.0;
        }
        if (!ok && !(tmt == 1 && bn == 32) && !(EPI == 3 && dw_tile()[0] == tmt && dw_tile()[1] == bn)) continue;
        const long rounds = (tiles + sms - 1) / sms;
        const int ct = tmt * (bn >= 64 ? bn / 64 : 1);
        const bool overlap = 2 * tmt * bn <= 512;
        const double t_tile = bytes / (double)tiles / 70.0 + (overlap ? 100.0 : 350.0 * ct);
        double cost = rounds * t_tile + 350.0 * ct + 1500.0;
        if (forced_tile()[0] == tmt && forced_tile()[1] == bn) cost = 0.0;   // profiling / test override (grx_gemm_debug_tile)
        if (EPI == 3 && dw_tile()[0] == tmt && dw_tile()[1] == bn) { bt = tmt; bb = bn; break; }   // weight-gradient groups: tile chosen by measurement (see dw_tile)
        if (best < 0 || cost < best) { best = cost; bt = tmt; bb = bn; }
    }
    if (EPI != 2) {
        if (bt == 2 && bb == 256) return launch_cfg<A_KMAJ, B_KMAJ, EPI == 2 ? 0 : EPI, 2, 256, 2>(ps, np, splits, st);
        if (bt == 2 && bb == 128) return launch_cfg<A_KMAJ, B_KMAJ, EPI == 2 ? 0 : EPI, 2, 128, 3>(ps, np, splits, st);
        if (bt == 1 && bb == 256) return launch_cfg<A_KMAJ, B_KMAJ, EPI == 2 ? 0 : EPI, 1, 256, 3>(ps, np, splits, st);
    }
    if (bb == 128) return launch_cfg<A_KMAJ, B_KMAJ, EPI, 1, 128, 4>(ps, np, splits, st);
    if (bb == 64) return launch_cfg<A_KMAJ, B_KMAJ, EPI, 1, 64, 7>(ps, np, splits, st);
    return launch_cfg<A_KMAJ, B_KMAJ, EPI, 1, 32, 8>(ps, np, splits, st);
}

// Layer pipelining: several DEPENDENT layers (problem i reads, as its A operand, the output of problem deps[i] < i; equal M) in ONE persistent
// launch.  The static tile list is ordered by problem, every CTA walks it in order, and a tile's loads wait on the row-block counter of the
// producing problem — so layer l + 1 starts on the row blocks that are complete while the stragglers of layer l finish, and the launch gap,
// the set-up, the first TMA round trip and the un-overlapped last epilogue are paid once per CHAIN instead of once per layer.  Measured gain: small
// (see pipe_flag) — with programmatic dependent launch the per-layer boundary is already mostly hidden, and the chain's own tail (the last tiles
// of layer l + 1 wait for the last tiles of layer l) replaces it.  One tile shape for the whole chain: 2 x 128 rows x 128 columns (bias / ELU epilogues),
// 1 x 128 x 128 (ELU' epilogue).  Requires every CTA to be co-resident (grid <= SM count, one CTA per SM: true for this kernel).
inline int &pipe_flag() {
    // bit 0: forward chains, bit 1: input-gradient chains.  Measured per minibatch of the update graph (profiles/r2r_pipe_variants.txt): none 176.4 us,
    // input-gradient chain 175.0, forward chain 180.5 (its one tile shape costs layer 2 the 2 x 256 macro tile), both 178.8 -> default 2.
    static int on = [] { const char *e = getenv("GRX_LAYER_PIPE"); return e ? atoi(e) : 2; }();
    return on;
}
template <bool A_KMAJ, bool B_KMAJ, int EPI>
inline bool pipe_supported(const Problem *ps, int np, const int *deps) {
    if (np < 2 || np > MAXP || pipe_flag() == 0) return false;
    for (int i = 0; i < np; i++) {
        if (!supported<A_KMAJ, B_KMAJ>(ps[i]) || ps[i].N < 65 || ps[i].M != ps[0].M) return false;
        if (deps[i] >= i) return false;
        if (EPI == 2 && ps[i].K < 64) return false;
    }
    return (ps[0].M + TM - 1) / TM <= SYNC_RB;
}
template <bool A_KMAJ, bool B_KMAJ, int EPI>
inline cudaError_t launch_pipe(const Problem *ps, int np, const int *deps, int *sync, int *err, cudaStream_t st) {
    if (EPI == 2) return launch_cfg<A_KMAJ, B_KMAJ, EPI, 1, 128, 4>(ps, np, nullptr, st, deps, sync, err);
    if (ps[0].M > TM) return launch_cfg<A_KMAJ, B_KMAJ, EPI == 2 ? 0 : EPI, 2, 128, 3>(ps, np, nullptr, st, deps, sync, err);
    return launch_cfg<A_KMAJ, B_KMAJ, EPI, 1, 128, 4>(ps, np, nullptr, st, deps, sync, err);
}

template <bool A_KMAJ, bool B_KMAJ, int EPI>
inline cudaError_t launch(const Problem &p, int splits, cudaStream_t st) {
    return launch_group<A_KMAJ, B_KMAJ, EPI>(&p, 1, &splits, st);
}

}  // namespace tc
