// grx_mlp_chain.cuh — the three hidden layers of the registered GRx policy (in -> 512 -> 256 -> 128, ELU; rsl_rl/modules/mlp.py:26-41) as ONE
// persistent tcgen05 kernel per forward pass, sm_100a only: an experiment in removing the per-launch fixed costs of the layerwise path
// (three dependent grouped GEMM launches, grx_gemm_tc.cuh).  OPT-IN (GRX_FUSED_FWD=1): parity-green on B200, not faster (see fwd_flag()).
//
// One CTA tile = 128 rows of one network.  The hidden activations never travel through HBM between the layers: the epilogue warps turn an
// accumulator chunk into bias + ELU'd fp32 values and write them BACK INTO TENSOR MEMORY, in place, where the MMA warp uses them as the
// A operand of the next layer (tcgen05.mma with A in TMEM: accumulator layout == A layout).  The same values are TMA-stored to H1 / H2 / H3
// in HBM when the backward pass needs them (update) — or only H3 (rollout).  X and the weights stream through a ring of ten uniform 16 KB
// stages (128 rows x 32 k, SWIZZLE_128B K-major).
//   warps: 0-7 epilogue (quad = warp % 4 owns TMEM lanes / tile rows 32 quad .. +31, group = warp / 4 owns every other 32-column chunk),
//          8 = TMA producer (one lane), 9 = MMA issuer (one lane)
//   MMA issue order per tile:  L0(0) L0(1) L1(0) L0(2) L1(1) L0(3) L1(2) L1(3) L2 — layer 0 is produced in four 128-column chunks; L0(c+1)
//          runs on the tensor core while the epilogue warps turn chunk c into the A operand of L1(c); the producer streams the stages in
//          the same order.
// Every mbarrier wait is bounded (~0.1 s): a protocol error sets an error flag (reported to the host) instead of hanging the GPU.
#pragma once
#include "grx_gemm_tc.cuh"

namespace tc {
namespace chain {

constexpr int D1 = 512, D2 = 256, D3 = 128;      // hidden widths this kernel is built for (checked by the host)
constexpr int NKX = 8;                            // input width <= 256 (k-blocks of 32)
constexpr int NTHREADS = 320;
constexpr uint32_t T16K = 16384u;

struct FwdNet {
    const float *b0, *b1, *b2;
    int K0;
};
struct FwdArgs {
    FwdNet net[2];
    int np, M, tiles_per_net, total_tiles;
    int store_hidden;      // 1: H1 / H2 go to HBM too (update: the backward pass reads them), 0: only H3 (rollout)
    int *err;              // device int, set to 1 when a barrier wait timed out
};
struct FwdMaps {
    CUtensorMap x[2], w0[2], w1[2], w2[2], h1[2], h2[2], h3[2];
};

__device__ __forceinline__ void mbar_wait_a(uint32_t bar, uint32_t parity, volatile int *abort_flag) {
    uint32_t ok = 0;
    int spins = 0;
    long long t0 = 0;
    while (true) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        if (ok) return;
        if (++spins >= 64) {
            if (*abort_flag) return;
            if (t0 == 0) t0 = clock64();
            else if (clock64() - t0 > 200000000ll) { *abort_flag = 1; return; }
        }
    }
}

// =====================================================================================================================
// TS variant: the ELU'd chunk goes BACK INTO TENSOR MEMORY, in place over the accumulator it came from, and the next layer's MMAs take
// their A operand from TMEM (tcgen05.mma [d], [a_tmem], b_desc: accumulator layout == A layout, lane = row, one tf32 per 32-bit column).
// With kind::tf32 an SS-mode MMA of N = 128 reads 4 KB of A + 4 KB of B from shared memory per 64 tensor cycles = the whole 128 B/clk of
// the SM's shared-memory port, on top of the TMA writes of the next stages: measured 3.7 us per layer-0 chunk instead of 1.6.  Taking A
// from TMEM halves the shared-memory traffic of the tensor pipe.  Shared memory then only stages the TMA stores (2 pieces per warp).
//   TMEM: [0,128) / [128,256) layer-0 chunk ping-pong (accumulator, then A operand of layer 1) | [256,512) layer-1 accumulator, then A
//   operand of layer 2 | layer-2 accumulator re-uses [0,128).  Ordering between an MMA that reads a buffer as A and a later MMA that
//   overwrites it needs no barrier: one thread issues all MMAs and the tensor pipe executes them in issue order.
// =====================================================================================================================
constexpr int NSB_TS = 10, NSTG = 4;   // operand stage ring (16 KB each: 160 KB in flight); store staging = 8 warps x 2 pieces x 4 KB = 4 x 16 KB
constexpr size_t SMEM_BYTES_TS = (size_t)(NSTG + NSB_TS) * T16K + 1024;

__device__ __forceinline__ void tc_mma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
                 ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tc_st32(uint32_t taddr, const float *v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
                 "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
                 ::"r"(taddr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
                   "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])),
                   "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
                   "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15])),
                   "r"(__float_as_uint(v[16])), "r"(__float_as_uint(v[17])), "r"(__float_as_uint(v[18])), "r"(__float_as_uint(v[19])),
                   "r"(__float_as_uint(v[20])), "r"(__float_as_uint(v[21])), "r"(__float_as_uint(v[22])), "r"(__float_as_uint(v[23])),
                   "r"(__float_as_uint(v[24])), "r"(__float_as_uint(v[25])), "r"(__float_as_uint(v[26])), "r"(__float_as_uint(v[27])),
                   "r"(__float_as_uint(v[28])), "r"(__float_as_uint(v[29])), "r"(__float_as_uint(v[30])), "r"(__float_as_uint(v[31]))
                 : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

__global__ void __launch_bounds__(NTHREADS, 1) mlp_fwd_chain_ts_kernel(const __grid_constant__ FwdMaps maps, const __grid_constant__ FwdArgs args) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ __align__(8) unsigned long long full_bar[NSB_TS], empty_bar[NSB_TS], h1_ready[2][4], h2_ready[8], acc0_full[2], acc1_full, acc2_full,
        acc2_empty;
    __shared__ uint32_t tmem_slot;
    __shared__ int s_abort;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) stamp(0);
    const uint32_t smem0 = (smem_u32(smem) + 1023u) & ~1023u;
    const uint32_t stgreg = smem0, breg = stgreg + (uint32_t)NSTG * T16K;
    volatile int *abortp = &s_abort;

    if (tid == 0) {
        for (int i = 0; i < NSB_TS; i++) { mbar_init(smem_u32(&full_bar[i]), 1); mbar_init(smem_u32(&empty_bar[i]), 1); }
        for (int i = 0; i < 8; i++) { mbar_init(smem_u32(&h1_ready[i >> 2][i & 3]), 4); mbar_init(smem_u32(&h2_ready[i]), 4); }
        for (int i = 0; i < 2; i++) mbar_init(smem_u32(&acc0_full[i]), 1);
        mbar_init(smem_u32(&acc1_full), 1); mbar_init(smem_u32(&acc2_full), 1); mbar_init(smem_u32(&acc2_empty), 8);
        s_abort = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "n"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (warp == 8 && lane == 0) {
        for (int p = 0; p < args.np; p++) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.x[p]) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.w0[p]) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.w1[p]) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.w2[p]) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.h3[p]) : "memory");
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    pdl_wait();
    pdl_launch_dependents();
    if (tid == 0) stamp(1);
    constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    // Tile list: the LAST network (the critic: widest input, most layer-0 work) first.  CTA i takes tile i, then tiles from the END of
    // the second round downwards, so the extra tiles go to the CTAs whose first tile was the cheaper network.
    const int G = (int)gridDim.x;
    auto tile_of = [&](int ti) -> int {   // ti-th tile of this CTA, or -1
        if (ti == 0) return (int)blockIdx.x < args.total_tiles ? (int)blockIdx.x : -1;
        const int t = ti * G + (G - 1 - (int)blockIdx.x);
        return t < args.total_tiles ? t : -1;
    };
    auto net_of = [&](int t) -> int { return args.np - 1 - t / args.tiles_per_net; };

    if (warp == 8) {
        if (lane == 0) {   // ------------------------------------------------------------------ TMA producer
            uint32_t it = 0, ti = 0;
            auto stage = [&](const CUtensorMap *m, int k0, int r0) {
                const uint32_t st = it % NSB_TS;
                if (it >= NSB_TS) mbar_wait_a(smem_u32(&empty_bar[st]), ((it / NSB_TS) - 1) & 1, abortp);
                const uint32_t bar = smem_u32(&full_bar[st]);
                mbar_expect_tx(bar, T16K);
                tma_load_2d(breg + st * T16K, m, k0, r0, bar);
                it++;
            };
            for (int t = tile_of(0); t >= 0; t = tile_of(++ti)) {
                const int p = net_of(t), m0 = (t % args.tiles_per_net) * 128;
                const int nk0 = (args.net[p].K0 + 31) / 32;
                // layer 0 streams the X k-block with every weight k-block (X is re-read per 128-column chunk: keeping it resident would cost
                // up to 96 KB of the shared memory that the operand ring needs to cover the ~2 us L2 latency under load)
                auto L0 = [&](int c) { for (int kb = 0; kb < nk0; kb++) { stage(&maps.x[p], kb * 32, m0); stage(&maps.w0[p], kb * 32, c * 128); } };
                auto L1 = [&](int c) { for (int j = 0; j < 4; j++) for (int nh = 0; nh < 2; nh++) stage(&maps.w1[p], (4 * c + j) * 32, nh * 128); };
                L0(0); L0(1); L1(0); L0(2); L1(1); L0(3); L1(2); L1(3);
                for (int kb = 0; kb < 8; kb++) stage(&maps.w2[p], kb * 32, 0);
            }
        }
    } else if (warp == 9) {
        if (lane == 0) {   // ------------------------------------------------------------------ MMA issuer
            uint32_t it = 0, ti = 0;
            auto stage_wait = [&]() -> uint32_t {
                const uint32_t st = it % NSB_TS;
                mbar_wait_a(smem_u32(&full_bar[st]), (it / NSB_TS) & 1, abortp);
                tc_fence_after();
                return breg + st * T16K;
            };
            auto stage_done = [&]() { tc_commit(smem_u32(&empty_bar[it % NSB_TS])); it++; };
            for (int t = tile_of(0); t >= 0; t = tile_of(++ti)) {
                const int p = net_of(t);
                const int nk0 = (args.net[p].K0 + 31) / 32;
                auto L0 = [&](int c) {
                    const int b = c & 1;
                    if (c == 0 && ti >= 1) { mbar_wait_a(smem_u32(&acc2_empty), (ti - 1) & 1, abortp); tc_fence_after(); }   // buffer 0 hosted the previous tile's layer-2 accumulator
                    for (int kb = 0; kb < nk0; kb++) {
                        const uint32_t ta = stage_wait();          // X k-block
                        const uint32_t sa = it % NSB_TS;
                        it++;
                        const uint32_t tb = stage_wait();          // W0 k-block of this chunk
                        if (ti == 0 && c == 0 && kb == 0) stamp(2);
#pragma unroll
                        for (int j = 0; j < 4; j++)
                            tc_mma_tf32(tmem + (uint32_t)(b * 128), smem_desc(ta + (uint32_t)j * 32u, 16u, 1024u, 2u), smem_desc(tb + (uint32_t)j * 32u, 16u, 1024u, 2u), IDESC,
                                        (kb == 0 && j == 0) ? 0u : 1u);
                        tc_commit(smem_u32(&empty_bar[sa]));
                        stage_done();
                    }
                    tc_commit(smem_u32(&acc0_full[b]));
                    if (ti == 0 && c == 0) stamp(3);
                };
                auto L1 = [&](int c) {
                    const int b = c & 1;
                    for (int j = 0; j < 4; j++) {
                        mbar_wait_a(smem_u32(&h1_ready[b][j]), (uint32_t)(c >> 1) & 1, abortp);   // two completions per tile: chunks b and b + 2
                        tc_fence_after();
                        if (ti == 0 && c == 0 && j == 0) stamp(4);
                        for (int nh = 0; nh < 2; nh++) {
                            const uint32_t tb = stage_wait();
#pragma unroll
                            for (int jj = 0; jj < 4; jj++)
                                tc_mma_tf32_ts(tmem + (uint32_t)(256 + nh * 128), tmem + (uint32_t)(b * 128 + j * 32 + jj * 8), smem_desc(tb + (uint32_t)jj * 32u, 16u, 1024u, 2u),
                                               IDESC, (c == 0 && j == 0 && jj == 0) ? 0u : 1u);
                            stage_done();
                        }
                    }
                    if (c == 3) { tc_commit(smem_u32(&acc1_full)); if (ti == 0) stamp(5); }
                };
                L0(0); L0(1); L1(0); L0(2); L1(1); L0(3); L1(2); L1(3);
                for (int kb = 0; kb < 8; kb++) {
                    mbar_wait_a(smem_u32(&h2_ready[kb]), ti & 1, abortp);
                    tc_fence_after();
                    const uint32_t tb = stage_wait();
#pragma unroll
                    for (int jj = 0; jj < 4; jj++)
                        tc_mma_tf32_ts(tmem, tmem + (uint32_t)(256 + kb * 32 + jj * 8), smem_desc(tb + (uint32_t)jj * 32u, 16u, 1024u, 2u), IDESC, (kb == 0 && jj == 0) ? 0u : 1u);
                    stage_done();
                }
                tc_commit(smem_u32(&acc2_full));
                if (ti == 0) stamp(6);
            }
        }
    } else {
        // ---------------------------------------------------------------------------------- epilogue warps
        const int quad = warp & 3, grp = warp >> 2;
        const uint32_t lane_base = ((uint32_t)(quad * 32)) << 16;
        const uint32_t sw = (uint32_t)(lane & 7), row_off = (uint32_t)lane * 128u;
        const uint32_t my_stg = stgreg + (uint32_t)warp * 8192u;   // two 4 KB pieces per warp
        uint32_t ti = 0, f0[2] = {0, 0}, nst = 0;
        auto load_bias = [&](float4 *bq, const float *bias32) {
#pragma unroll
            for (int i = 0; i < 8; i++) bq[i] = __ldg(reinterpret_cast<const float4 *>(bias32) + i);
        };
        auto bias_elu = [&](float *v, const float4 *bq) {
#pragma unroll
            for (int i = 0; i < 8; i++) {
                v[4 * i] = elu_f(v[4 * i] + bq[i].x); v[4 * i + 1] = elu_f(v[4 * i + 1] + bq[i].y);
                v[4 * i + 2] = elu_f(v[4 * i + 2] + bq[i].z); v[4 * i + 3] = elu_f(v[4 * i + 3] + bq[i].w);
            }
        };
        for (int t = tile_of(0); t >= 0; t = tile_of(++ti)) {
            const int p = net_of(t), m0 = (t % args.tiles_per_net) * 128;
            const FwdNet &net = args.net[p];
            const int row0 = m0 + quad * 32;
            auto store_chunk = [&](const float *v, const CUtensorMap *hmap, int col) {   // registers -> this warp's staging piece -> TMA store
                const uint32_t piece = my_stg + (nst & 1u) * 4096u;
                nst++;
                if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");   // the store that used this piece two stores ago has read it
                __syncwarp();
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    const uint32_t addr = piece + row_off + ((((uint32_t)i) ^ sw) << 4);
                    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v[4 * i]), "f"(v[4 * i + 1]), "f"(v[4 * i + 2]), "f"(v[4 * i + 3]) : "memory");
                }
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) { tma_store_2d(hmap, col, row0, piece); tma_commit(); }
            };
            for (int c = 0; c < 4; c++) {
                const int b = c & 1;
                mbar_wait_a(smem_u32(&acc0_full[b]), f0[b] & 1, abortp);
                f0[b]++;
                tc_fence_after();
                if (ti == 0 && c == 0 && tid == 0) stamp(7);
#pragma unroll 1
                for (int jj = 0; jj < 2; jj++) {
                    const int j = grp + 2 * jj, col = c * 128 + j * 32;
                    float4 bq[8];
                    load_bias(bq, net.b0 + col);
                    float v[32];
                    const uint32_t ta = tmem + lane_base + (uint32_t)(b * 128 + j * 32);
                    tc_ld32(ta, v);
                    bias_elu(v, bq);
                    tc_st32(ta, v);                                   // in place: the accumulator chunk becomes the A operand of layer 1
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(smem_u32(&h1_ready[b][j]));
                    if (args.store_hidden) store_chunk(v, &maps.h1[p], col);
                }
                if (ti == 0 && tid == 0 && (c == 0 || c == 3)) stamp(c == 0 ? 8 : 9);
            }
            mbar_wait_a(smem_u32(&acc1_full), ti & 1, abortp);
            tc_fence_after();
            if (ti == 0 && tid == 0) stamp(10);
#pragma unroll 1
            for (int jj = 0; jj < 4; jj++) {
                const int kb = grp + 2 * jj;
                float4 bq[8];
                load_bias(bq, net.b1 + kb * 32);
                float v[32];
                const uint32_t ta = tmem + lane_base + (uint32_t)(256 + kb * 32);
                tc_ld32(ta, v);
                bias_elu(v, bq);
                tc_st32(ta, v);
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(smem_u32(&h2_ready[kb]));
                if (args.store_hidden) store_chunk(v, &maps.h2[p], kb * 32);
            }
            if (ti == 0 && tid == 0) stamp(11);
            mbar_wait_a(smem_u32(&acc2_full), ti & 1, abortp);
            tc_fence_after();
            if (ti == 0 && tid == 0) stamp(12);
#pragma unroll 1
            for (int jj = 0; jj < 2; jj++) {
                const int j = grp + 2 * jj;
                float4 bq[8];
                load_bias(bq, net.b2 + j * 32);
                float v[32];
                tc_ld32(tmem + lane_base + (uint32_t)(j * 32), v);
                if (jj == 1) { tc_fence_before(); __syncwarp(); if (lane == 0) mbar_arrive(smem_u32(&acc2_empty)); }
                bias_elu(v, bq);
                store_chunk(v, &maps.h3[p], j * 32);
            }
        }
        if (lane == 0) tma_wait_read0();
        if (tid == 0) stamp(13);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512) : "memory");
    if (tid == 0 && s_abort && args.err) *args.err = 1;
}

// ---------------------------------------------------------------------------------------------------------------------
// Host side
// ---------------------------------------------------------------------------------------------------------------------
struct FwdProblem {
    const float *X;           // [M, K0], row stride ldx (multiple of 4 floats, 16-byte aligned)
    int K0, ldx;
    const float *W0, *b0;     // [512, K0] row stride ld0
    int ld0;
    const float *W1, *b1;     // [256, 512]
    const float *W2, *b2;     // [128, 256]
    float *H1, *H2, *H3;      // [M, 512] [M, 256] [M, 128]
};

inline bool fwd_supported(const FwdProblem *ps, int np) {
    if (np < 1 || np > 2) return false;
    for (int i = 0; i < np; i++) {
        const FwdProblem &p = ps[i];
        if (p.K0 < 1 || p.K0 > NKX * 32 || (p.ldx & 3) || (p.ld0 & 3)) return false;
        for (const void *q : {(const void *)p.X, (const void *)p.W0, (const void *)p.W1, (const void *)p.W2, (const void *)p.b0, (const void *)p.b1,
                              (const void *)p.b2, (const void *)p.H1, (const void *)p.H2, (const void *)p.H3})
            if (!aligned16(q)) return false;
    }
    return true;
}
// GRX_FUSED_FWD / grx_ppo_debug_fused: 0 one grouped GEMM launch per layer (DEFAULT), 1 the chained kernel.  Measured on B200 (round 2,
// profiles/r2_chain_kernel_timeline.txt): the chained kernel is numerically identical but NOT faster — 192 vs 185 us per minibatch, 50.6 vs
// 45.4 us per policy step — because every 128-row tile re-reads all three weight matrices (0.75-1.4 MB) through one SM's ~85 GB/s L2 port
// (17-29 us per tile, and 164 tiles on 148 SMs need two rounds), while the layerwise kernels share each weight tile between 256 rows.
inline int &fwd_flag() {
    static int on = [] { const char *e = getenv("GRX_FUSED_FWD"); return e ? atoi(e) : 0; }();
    return on;
}
inline bool fwd_enabled() { return fwd_flag() != 0; }

inline cudaError_t launch_fwd(const FwdProblem *ps, int np, int M, bool store_hidden, int *d_err, cudaStream_t st) {
    static bool attr_done = false;
    if (!attr_done) {
        cudaError_t e = cudaFuncSetAttribute(mlp_fwd_chain_ts_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES_TS);
        if (e != cudaSuccess) return e;
        attr_done = true;
    }
    FwdMaps maps;
    FwdArgs a;
    memset(&a, 0, sizeof(a));
    a.np = np; a.M = M; a.tiles_per_net = (M + 127) / 128; a.total_tiles = np * a.tiles_per_net; a.store_hidden = store_hidden ? 1 : 0; a.err = d_err;
    for (int i = 0; i < 2; i++) {
        const FwdProblem &p = ps[i < np ? i : 0];
        cudaError_t e;
        if ((e = get_tensor_map(p.X, M, p.K0, p.ldx, 128, true, &maps.x[i])) != cudaSuccess) return e;
        if ((e = get_tensor_map(p.W0, D1, p.K0, p.ld0, 128, true, &maps.w0[i])) != cudaSuccess) return e;
        if ((e = get_tensor_map(p.W1, D2, D1, D1, 128, true, &maps.w1[i])) != cudaSuccess) return e;
        if ((e = get_tensor_map(p.W2, D3, D2, D2, 128, true, &maps.w2[i])) != cudaSuccess) return e;
        if ((e = get_tensor_map(p.H1, M, D1, D1, 32, true, &maps.h1[i])) != cudaSuccess) return e;
        if ((e = get_tensor_map(p.H2, M, D2, D2, 32, true, &maps.h2[i])) != cudaSuccess) return e;
        if ((e = get_tensor_map(p.H3, M, D3, D3, 32, true, &maps.h3[i])) != cudaSuccess) return e;
        a.net[i].b0 = p.b0; a.net[i].b1 = p.b1; a.net[i].b2 = p.b2; a.net[i].K0 = p.K0;
    }
    if (a.total_tiles == 0) return cudaSuccess;
    const int grid = a.total_tiles < sm_count() ? a.total_tiles : sm_count();
    return launch_kernel(mlp_fwd_chain_ts_kernel, dim3(grid), dim3(NTHREADS), SMEM_BYTES_TS, st, true, maps, a);
}

}  // namespace chain
}  // namespace tc
