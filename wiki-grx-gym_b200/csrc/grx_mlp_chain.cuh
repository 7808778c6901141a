// grx_mlp_chain.cuh — the three hidden layers of the registered GRx policy (in -> 512 -> 256 -> 128, ELU; rsl_rl/modules/mlp.py:26-41) as ONE
// persistent tcgen05 kernel per forward pass, sm_100a only.  Replaces three dependent grouped GEMM launches (grx_gemm_tc.cuh) whose fixed
// costs (set-up, first TMA round trip, un-overlapped last epilogue, launch gap: ~6 us each) were as large as their main loops.
//
// One CTA tile = 128 rows of one network.  The hidden activations never travel through HBM between the layers: the epilogue warps turn
// an accumulator chunk (TMEM) into bias + ELU'd fp32 values and write them into shared memory in EXACTLY the layout of a K-major
// SWIZZLE_128B UMMA A operand (the same 128-byte-row, 16-byte-unit XOR (row % 8) box the TMA store of grx_gemm_tc.cuh uses), where the
// MMA warp picks them up as the A operand of the next layer; the same box is also TMA-stored to H1 / H2 / H3 in HBM when the backward
// pass needs them (update) — or only H3 (rollout).  Weights stream through a ring of uniform 16 KB B stages (128 output rows x 32 k).
//
//   TMEM (512 columns):  [0,128) / [128,256)  layer-0 accumulator chunk, ping-pong (layer 0 is produced in four 128-column chunks)
//                        [256,512)            layer-1 accumulator (128 x 256), fed k-block by k-block as layer-0 chunks are ELU'd
//                        [0,128)              layer-2 accumulator (re-uses the ping buffer once layer 0 is done)
//   shared memory:       X tile resident (<= 6 k-blocks x 16 KB) | ring of 4 A-operand boxes (16 KB) | ring of 4 B stages (16 KB)
//   warps:               0-7 epilogue (quad = warp % 4 owns TMEM lanes / tile rows 32 quad .. 32 quad + 31, group = warp / 4 owns every
//                        other 32-column chunk), 8 = TMA producer (one lane), 9 = MMA issuer (one lane)
//   MMA issue order per tile:  L0(0) L0(1) L1(0) L0(2) L1(1) L0(3) L1(2) L1(3) L2   — L0(c+1) runs on the tensor core while the epilogue
//                        warps turn chunk c into the A operand of L1(c); the producer streams the weight stages in the same order.
// Every mbarrier wait is bounded (~0.1 s): a protocol error sets an error flag (reported to the host) instead of hanging the GPU.
#pragma once
#include "grx_gemm_tc.cuh"

namespace tc {
namespace chain {

constexpr int D1 = 512, D2 = 256, D3 = 128;      // hidden widths this kernel is built for (checked by the host)
constexpr int NKX = 6;                            // resident X k-blocks: input width <= 192
constexpr int NA2 = 4, NSB = 4;                   // A-operand box ring, B stage ring
constexpr int NTHREADS = 320;
constexpr uint32_t T16K = 16384u;
constexpr size_t SMEM_BYTES = (size_t)(NKX + NA2 + NSB) * T16K + 1024;
constexpr int BOXES_PER_TILE = 24;                // MMA-consumed A boxes per tile: 16 (layer-1 k-blocks) + 8 (layer-2 k-blocks)

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

// one 32-column chunk of an accumulator: TMEM -> registers -> + bias -> ELU -> this warp's 32-row piece of an A-operand box
__device__ __forceinline__ void chunk_to_box(const float *v, const float *__restrict__ bias32, uint32_t piece_row, uint32_t sw) {
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const float4 bb = __ldg(reinterpret_cast<const float4 *>(bias32) + i);
        float4 x = make_float4(elu_f(v[4 * i] + bb.x), elu_f(v[4 * i + 1] + bb.y), elu_f(v[4 * i + 2] + bb.z), elu_f(v[4 * i + 3] + bb.w));
        const uint32_t addr = piece_row + ((((uint32_t)i) ^ sw) << 4);
        asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(x.x), "f"(x.y), "f"(x.z), "f"(x.w) : "memory");
    }
}

__global__ void __launch_bounds__(NTHREADS, 1) mlp_fwd_chain_kernel(const __grid_constant__ FwdMaps maps, const __grid_constant__ FwdArgs args) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ __align__(8) unsigned long long full_bar[NSB], empty_bar[NSB], a2_full[NA2], a2_empty[NA2], acc0_full[2], acc0_empty[2],
        acc1_full, acc1_empty, acc2_full, x_full, x_empty;
    __shared__ uint32_t tmem_slot;
    __shared__ int s_abort;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t smem0 = (smem_u32(smem) + 1023u) & ~1023u;
    const uint32_t xreg = smem0, a2reg = smem0 + (uint32_t)NKX * T16K, breg = a2reg + (uint32_t)NA2 * T16K;
    volatile int *abortp = &s_abort;

    if (tid == 0) {
        for (int i = 0; i < NSB; i++) { mbar_init(smem_u32(&full_bar[i]), 1); mbar_init(smem_u32(&empty_bar[i]), 1); }
        for (int i = 0; i < NA2; i++) { mbar_init(smem_u32(&a2_full[i]), 4); mbar_init(smem_u32(&a2_empty[i]), 1); }
        for (int i = 0; i < 2; i++) { mbar_init(smem_u32(&acc0_full[i]), 1); mbar_init(smem_u32(&acc0_empty[i]), 8); }
        mbar_init(smem_u32(&acc1_full), 1); mbar_init(smem_u32(&acc1_empty), 8); mbar_init(smem_u32(&acc2_full), 1);
        mbar_init(smem_u32(&x_full), 1); mbar_init(smem_u32(&x_empty), 1);
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

    // instruction descriptor: D fp32, A / B tf32, both K-major, N = 128, M = 128
    constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);

    if (warp == 8) {
        if (lane == 0) {   // ------------------------------------------------------------------ TMA producer
            uint32_t it = 0, ti = 0;
            auto stage = [&](const CUtensorMap *m, int k0, int r0) {
                const uint32_t st = it % NSB;
                if (it >= NSB) mbar_wait_a(smem_u32(&empty_bar[st]), ((it / NSB) - 1) & 1, abortp);
                const uint32_t bar = smem_u32(&full_bar[st]);
                mbar_expect_tx(bar, T16K);
                tma_load_2d(breg + st * T16K, m, k0, r0, bar);
                it++;
            };
            for (int t = blockIdx.x; t < args.total_tiles; t += gridDim.x, ti++) {
                const int p = t / args.tiles_per_net, m0 = (t % args.tiles_per_net) * 128;
                const int nk0 = (args.net[p].K0 + 31) / 32;
                if (ti >= 1) mbar_wait_a(smem_u32(&x_empty), (ti - 1) & 1, abortp);   // layer 0 of the previous tile has read X
                mbar_expect_tx(smem_u32(&x_full), (uint32_t)nk0 * T16K);
                for (int kb = 0; kb < nk0; kb++) tma_load_2d(xreg + (uint32_t)kb * T16K, &maps.x[p], kb * 32, m0, smem_u32(&x_full));
                auto L0 = [&](int c) { for (int kb = 0; kb < nk0; kb++) stage(&maps.w0[p], kb * 32, c * 128); };
                auto L1 = [&](int c) { for (int j = 0; j < 4; j++) for (int nh = 0; nh < 2; nh++) stage(&maps.w1[p], (4 * c + j) * 32, nh * 128); };
                L0(0); L0(1); L1(0); L0(2); L1(1); L0(3); L1(2); L1(3);
                for (int kb = 0; kb < 8; kb++) stage(&maps.w2[p], kb * 32, 0);
            }
        }
    } else if (warp == 9) {
        if (lane == 0) {   // ------------------------------------------------------------------ MMA issuer
            uint32_t it = 0, ti = 0, gbase = 0;
            uint32_t e0[2] = {0, 0};   // fills of the layer-0 accumulator buffers so far (buffer 0 also hosts the layer-2 accumulator)
            auto mma4 = [&](uint32_t a_tile, uint32_t d_col, bool first) {   // one 32-deep k-block: 4 x (K = 8) on the B stage `it`
                const uint32_t st = it % NSB;
                mbar_wait_a(smem_u32(&full_bar[st]), (it / NSB) & 1, abortp);
                tc_fence_after();
                const uint32_t tb = breg + st * T16K;
#pragma unroll
                for (int j = 0; j < 4; j++)
                    tc_mma_tf32(tmem + d_col, smem_desc(a_tile + (uint32_t)j * 32u, 16u, 1024u, 2u), smem_desc(tb + (uint32_t)j * 32u, 16u, 1024u, 2u), IDESC,
                                (first && j == 0) ? 0u : 1u);
                tc_commit(smem_u32(&empty_bar[st]));
                it++;
            };
            for (int t = blockIdx.x; t < args.total_tiles; t += gridDim.x, ti++) {
                const int p = t / args.tiles_per_net;
                const int nk0 = (args.net[p].K0 + 31) / 32;
                mbar_wait_a(smem_u32(&x_full), ti & 1, abortp);
                tc_fence_after();
                auto acquire0 = [&](int b) {   // the epilogue has drained the previous contents of layer-0 buffer b
                    if (e0[b] >= 1) { mbar_wait_a(smem_u32(&acc0_empty[b]), (e0[b] - 1) & 1, abortp); tc_fence_after(); }
                    e0[b]++;
                };
                auto L0 = [&](int c) {
                    const int b = c & 1;
                    acquire0(b);
                    for (int kb = 0; kb < nk0; kb++) mma4(xreg + (uint32_t)kb * T16K, (uint32_t)(b * 128), kb == 0);
                    if (c == 3) tc_commit(smem_u32(&x_empty));
                    tc_commit(smem_u32(&acc0_full[b]));
                };
                auto L1 = [&](int c) {
                    if (c == 0 && ti >= 1) { mbar_wait_a(smem_u32(&acc1_empty), (ti - 1) & 1, abortp); tc_fence_after(); }
                    for (int j = 0; j < 4; j++) {
                        const uint32_t kb = 4 * c + j, g = gbase + kb, slot = g % NA2;
                        mbar_wait_a(smem_u32(&a2_full[slot]), (g / NA2) & 1, abortp);
                        tc_fence_after();
                        for (int nh = 0; nh < 2; nh++) mma4(a2reg + slot * T16K, (uint32_t)(256 + nh * 128), kb == 0);
                        tc_commit(smem_u32(&a2_empty[slot]));
                    }
                    if (c == 3) tc_commit(smem_u32(&acc1_full));
                };
                L0(0); L0(1); L1(0); L0(2); L1(1); L0(3); L1(2); L1(3);
                acquire0(0);   // layer-2 accumulator aliases layer-0 buffer 0
                for (int kb = 0; kb < 8; kb++) {
                    const uint32_t g = gbase + 16 + kb, slot = g % NA2;
                    mbar_wait_a(smem_u32(&a2_full[slot]), (g / NA2) & 1, abortp);
                    tc_fence_after();
                    mma4(a2reg + slot * T16K, 0u, kb == 0);
                    tc_commit(smem_u32(&a2_empty[slot]));
                }
                tc_commit(smem_u32(&acc2_full));
                gbase += BOXES_PER_TILE;
            }
        }
    } else {
        // ---------------------------------------------------------------------------------- epilogue warps
        const int quad = warp & 3, grp = warp >> 2;
        const uint32_t lane_base = ((uint32_t)(quad * 32)) << 16;
        const uint32_t sw = (uint32_t)(lane & 7);
        const uint32_t piece_off = (uint32_t)quad * 4096u, row_off = (uint32_t)lane * 128u;
        uint32_t ti = 0, gbase = 0, f0[2] = {0, 0};
        for (int t = blockIdx.x; t < args.total_tiles; t += gridDim.x, ti++) {
            const int p = t / args.tiles_per_net, m0 = (t % args.tiles_per_net) * 128;
            const FwdNet &net = args.net[p];
            const int row0 = m0 + quad * 32;
            // box g (index among the MMA-consumed boxes): wait until its slot is free, fill this warp's piece, publish, store
            auto emit = [&](const float *v, const float *bias32, uint32_t g, bool consumed, const CUtensorMap *hmap, int col, bool store) {
                const uint32_t slot = g % NA2;
                if (consumed && g >= (uint32_t)NA2) mbar_wait_a(smem_u32(&a2_empty[slot]), ((g / NA2) - 1) & 1, abortp);   // MMAs of the slot's previous box done
                if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");                               // my store from this piece (2 boxes ago) has read it
                __syncwarp();
                const uint32_t piece = a2reg + slot * T16K + piece_off;
                chunk_to_box(v, bias32, piece + row_off, sw);
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) {
                    if (consumed) mbar_arrive(smem_u32(&a2_full[slot]));
                    if (store) tma_store_2d(hmap, col, row0, piece);
                    tma_commit();
                }
            };
            for (int c = 0; c < 4; c++) {   // layer-0 chunks -> H1 columns [128 c, +128) = layer-1 k-blocks 4c .. 4c+3
                const int b = c & 1;
                mbar_wait_a(smem_u32(&acc0_full[b]), f0[b] & 1, abortp);
                f0[b]++;
                tc_fence_after();
#pragma unroll 1
                for (int jj = 0; jj < 2; jj++) {
                    const int j = grp + 2 * jj;
                    float v[32];
                    tc_ld32(tmem + lane_base + (uint32_t)(b * 128 + j * 32), v);
                    if (jj == 1) { tc_fence_before(); __syncwarp(); if (lane == 0) mbar_arrive(smem_u32(&acc0_empty[b])); }
                    const int col = c * 128 + j * 32;
                    emit(v, net.b0 + col, gbase + (uint32_t)(4 * c + j), true, &maps.h1[p], col, args.store_hidden != 0);
                }
            }
            mbar_wait_a(smem_u32(&acc1_full), ti & 1, abortp);
            tc_fence_after();
#pragma unroll 1
            for (int jj = 0; jj < 4; jj++) {   // layer-1 accumulator -> H2 = layer-2 k-blocks
                const int kb = grp + 2 * jj;
                float v[32];
                tc_ld32(tmem + lane_base + (uint32_t)(256 + kb * 32), v);
                if (jj == 3) { tc_fence_before(); __syncwarp(); if (lane == 0) mbar_arrive(smem_u32(&acc1_empty)); }
                emit(v, net.b1 + kb * 32, gbase + 16u + (uint32_t)kb, true, &maps.h2[p], kb * 32, args.store_hidden != 0);
            }
            mbar_wait_a(smem_u32(&acc2_full), ti & 1, abortp);
            tc_fence_after();
#pragma unroll 1
            for (int jj = 0; jj < 2; jj++) {   // layer-2 accumulator -> H3 (always stored); staged in slots grp, grp + 2 (free: all MMAs of the tile are done)
                const int j = grp + 2 * jj;
                float v[32];
                tc_ld32(tmem + lane_base + (uint32_t)(j * 32), v);
                if (jj == 1) { tc_fence_before(); __syncwarp(); if (lane == 0) mbar_arrive(smem_u32(&acc0_empty[0])); }
                emit(v, net.b2 + j * 32, (uint32_t)j, false, &maps.h3[p], j * 32, true);
            }
            gbase += BOXES_PER_TILE;
        }
        if (lane == 0) tma_wait_read0();
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
inline int &fwd_flag() {   // GRX_FUSED_FWD=0 (or grx_ppo_debug_fused) falls back to one grouped GEMM launch per layer
    static int on = [] { const char *e = getenv("GRX_FUSED_FWD"); return e ? atoi(e) : 1; }();
    return on;
}
inline bool fwd_enabled() { return fwd_flag() != 0; }

inline cudaError_t launch_fwd(const FwdProblem *ps, int np, int M, bool store_hidden, int *d_err, cudaStream_t st) {
    static bool attr_done = false;
    if (!attr_done) {
        cudaError_t e = cudaFuncSetAttribute(mlp_fwd_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES);
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
    return launch_kernel(mlp_fwd_chain_kernel, dim3(grid), dim3(NTHREADS), SMEM_BYTES, st, true, maps, a);
}

}  // namespace chain
}  // namespace tc
