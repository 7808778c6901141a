// grx_terrain_gen.cu — the procedural terrain of the GRx tasks generated ON THE DEVICE: the int16 sample grid of Terrain.__init__
// (legged_gym/utils/terrain.py:38-164) with the tile generators of isaacgym/terrain_utils.py (pyramid_sloped_terrain :74-106,
// random_uniform_terrain :17-51, pyramid_stairs_terrain :195-227, discrete_obstacles_terrain :109-149) for the tile types reachable with the
// reference's 5-entry terrain_proportions, and the per-tile origin heights (terrain.py:159-164).
//
// Bit-identical to the reference for the same numpy seed (tests/test_terrain_gpu.py vs grx_b200/terrain.py, which is SHA-pinned to the
// reference classes): the numpy random stream is drawn on the host in the reference's call order (a few thousand draws), the array arithmetic —
// float64 slope products truncated to int16, bilinear up-sampling of the coarse random levels + round-half-even, stair rings, rectangle
// painting in draw order, platform clearing, border — runs here, one thread per grid sample, written once, coalesced.  float64 products are
// issued with explicit round-to-nearest intrinsics (no FMA contraction) so that they round like numpy's.
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

#include "grx_b200.h"
#include "grx_count.h"

int grx_set_error(int code, const std::string &msg);   // grx_env.cu

#define CK(call)                                                                                             \
    do {                                                                                                     \
        cudaError_t err__ = (call);                                                                          \
        if (err__ != cudaSuccess)                                                                            \
            return grx_set_error(GRX_E_CUDA, std::string(#call) + ": " + cudaGetErrorString(err__));        \
    } while (0)

namespace {

__device__ __forceinline__ short slope_sample(const grx_terrain_tile &t, const double *rx, const double *ry, int x, int y) {
    // (max_height * xx * yy).astype(int16): two float64 products left to right, C truncation (terrain_utils.py:92-98)
    const double v = __dmul_rn(__dmul_rn((double)t.slope_peak, rx[x]), ry[y]);
    return (short)(int)v;
}

__global__ void terrain_gen_kernel(const grx_terrain_tile *__restrict__ tiles, int num_rows, int num_cols, int W, int Lp, int border, int tot_rows,
                                   int tot_cols, const double *__restrict__ rx, const double *__restrict__ ry, const int *__restrict__ up_i0,
                                   const double *__restrict__ up_fx, const int *__restrict__ up_j0, const double *__restrict__ up_fy,
                                   const short *__restrict__ coarse, int cnx, int cny, const int *__restrict__ rects, short *__restrict__ out) {
    const int id = blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= tot_rows * tot_cols) return;
    const int gi = id / tot_cols, gj = id % tot_cols;
    const int ti = gi - border, tj = gj - border;
    short h = 0;
    if (ti >= 0 && tj >= 0 && ti < num_rows * Lp && tj < num_cols * W) {
        const int i = ti / Lp, j = tj / W, x = ti % Lp, y = tj % W;   // tile (i, j), sample (x, y) of its height_field_raw[W, W]
        const grx_terrain_tile &t = tiles[i * num_cols + j];
        if (t.kind == 0 || t.kind == 1) {
            h = slope_sample(t, rx, ry, x, y);
            const short corner = slope_sample(t, rx, ry, t.plat_lo, t.plat_lo);                       // terrain_utils.py:100-105
            const short lo = corner < 0 ? corner : (short)0, hi = corner > 0 ? corner : (short)0;
            h = h < lo ? lo : (h > hi ? hi : h);
            if (t.kind == 1) {   // + rint(bilinear(coarse levels)) (terrain_utils.py:41-51)
                const short *z = coarse + (size_t)t.coarse_index * cnx * cny;
                const int i0 = up_i0[x], j0 = up_j0[y];
                const double fx = up_fx[x], fy = up_fy[y];
                const double z00 = (double)z[i0 * cny + j0], z10 = (double)z[(i0 + 1) * cny + j0];
                const double z01 = (double)z[i0 * cny + j0 + 1], z11 = (double)z[(i0 + 1) * cny + j0 + 1];
                const double ofx = __dsub_rn(1.0, fx), ofy = __dsub_rn(1.0, fy);
                const double a = __dadd_rn(__dmul_rn(z00, ofx), __dmul_rn(z10, fx));
                const double b = __dadd_rn(__dmul_rn(z01, ofx), __dmul_rn(z11, fx));
                const double f = __dadd_rn(__dmul_rn(a, ofy), __dmul_rn(b, fy));
                h = (short)(h + (short)(int)rint(f));                                                    // int16 += int16 (wraps like numpy)
            }
        } else if (t.kind == 2) {   // concentric rings, ring n covers [n sw, W - n sw)^2 and is n steps high (terrain_utils.py:213-226)
            int n = min(min(x, Lp - 1 - x), min(y, W - 1 - y)) / t.step_width;
            n = min(n, t.num_rings);
            h = (short)(n * t.step_height);
        } else if (t.kind == 3) {   // rectangles painted in draw order: the last one covering the sample wins (terrain_utils.py:135-143)
            const int *r = rects + (size_t)t.rect_index * 5;
            for (int k = 0; k < t.num_rects; k++) {
                const int si = r[5 * k], sj = r[5 * k + 1], w = r[5 * k + 2], l = r[5 * k + 3];
                if (x >= si && x < si + w && y >= sj && y < sj + l) h = (short)r[5 * k + 4];
            }
            if (x >= t.plat_lo && x < t.plat_hi && y >= t.plat_lo && y < t.plat_hi) h = 0;             // terrain_utils.py:145-148
        }
    }
    out[id] = h;
}

// env_origins z (terrain.py:159-163): max of the tile's centre window; one warp per tile
__global__ void terrain_origin_kernel(const short *__restrict__ grid, int num_rows, int num_cols, int W, int Lp, int border, int tot_cols, int x1, int x2,
                                      int y1, int y2, int *__restrict__ zmax) {
    const int tile = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (tile >= num_rows * num_cols) return;
    const int i = tile / num_cols, j = tile % num_cols;
    int m = -32768;
    const int nw = (x2 - x1) * (y2 - y1);
    for (int k = lane; k < nw; k += 32) {
        const int x = x1 + k / (y2 - y1), y = y1 + k % (y2 - y1);
        m = max(m, (int)grid[(size_t)(border + i * Lp + x) * tot_cols + border + j * W + y]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (lane == 0) zmax[tile] = m;
}

}  // namespace

extern "C" int grx_terrain_generate(const grx_terrain_grid *g, const grx_terrain_tile *h_tiles, const double *h_rx, const double *h_ry,
                                    const int32_t *h_up_i0, const double *h_up_fx, const int32_t *h_up_j0, const double *h_up_fy,
                                    const int16_t *h_coarse, int32_t num_coarse, const int32_t *h_rects, int32_t num_rects_total,
                                    int16_t *d_samples, int32_t *h_origin_zmax, int32_t device, void *stream) {
    if (!g || !h_tiles || !h_rx || !h_ry || !d_samples) return grx_set_error(GRX_E_INVALID, "grx_terrain_generate: null argument");
    const int W = g->tile_width, Lp = g->tile_length, nt = g->num_rows * g->num_cols;
    if (g->num_rows < 1 || g->num_cols < 1 || W < 2 || Lp != W || g->border < 0)
        return grx_set_error(GRX_E_INVALID, "grx_terrain_generate: bad grid (tiles are square, width x width samples, as upstream terrain.py:112-116)");
    if (num_coarse > 0 && (!h_coarse || !h_up_i0 || !h_up_fx || !h_up_j0 || !h_up_fy || g->coarse_nx < 2 || g->coarse_ny < 2))
        return grx_set_error(GRX_E_INVALID, "grx_terrain_generate: rough tiles need the coarse level tables");
    int coarse_used = 0, rects_used = 0;
    for (int k = 0; k < nt; k++) {
        const grx_terrain_tile &t = h_tiles[k];
        if (t.kind < 0 || t.kind > 3) return grx_set_error(GRX_E_INVALID, "grx_terrain_generate: tile kind must be 0..3 (smooth slope, rough slope, stairs, obstacles)");
        if (t.kind <= 1 && (t.plat_lo < 0 || t.plat_lo >= W)) return grx_set_error(GRX_E_INVALID, "grx_terrain_generate: platform corner outside the tile");
        if (t.kind == 1) { if (t.coarse_index < 0 || t.coarse_index >= num_coarse) return grx_set_error(GRX_E_INVALID, "grx_terrain_generate: coarse_index out of range"); coarse_used++; }
        if (t.kind == 2 && t.step_width < 1) return grx_set_error(GRX_E_INVALID, "grx_terrain_generate: step_width < 1 sample");
        if (t.kind == 3) { if (t.rect_index < 0 || t.num_rects < 0 || t.rect_index + t.num_rects > num_rects_total) return grx_set_error(GRX_E_INVALID, "grx_terrain_generate: rectangle range out of bounds"); rects_used += t.num_rects; }
    }
    (void)coarse_used; (void)rects_used;
    CK(cudaSetDevice(device));
    cudaStream_t st = (cudaStream_t)stream;
    const int tot_rows = g->num_rows * Lp + 2 * g->border, tot_cols = g->num_cols * W + 2 * g->border;
    // one staging allocation for all the small host tables
    const size_t csz = (size_t)num_coarse * g->coarse_nx * g->coarse_ny;
    size_t off[10], total = 0;
    const size_t bytes[9] = {sizeof(grx_terrain_tile) * nt, 8 * (size_t)Lp, 8 * (size_t)W, 4 * (size_t)Lp, 8 * (size_t)Lp, 4 * (size_t)W, 8 * (size_t)W,
                             2 * csz, 20 * (size_t)num_rects_total};
    for (int k = 0; k < 9; k++) { off[k] = total; total += (bytes[k] + 15) & ~(size_t)15; }
    off[9] = total; total += 4 * (size_t)nt;
    unsigned char *d = nullptr;
    CK(cudaMalloc((void **)&d, total));
    const void *src[9] = {h_tiles, h_rx, h_ry, h_up_i0, h_up_fx, h_up_j0, h_up_fy, h_coarse, h_rects};
    for (int k = 0; k < 9; k++)
        if (bytes[k] && src[k]) CK(cudaMemcpyAsync(d + off[k], src[k], bytes[k], cudaMemcpyHostToDevice, st));
    const int n = tot_rows * tot_cols;
    grx_count_launch();
    terrain_gen_kernel<<<(n + 255) / 256, 256, 0, st>>>(reinterpret_cast<const grx_terrain_tile *>(d + off[0]), g->num_rows, g->num_cols, W, Lp, g->border,
                                                        tot_rows, tot_cols, reinterpret_cast<const double *>(d + off[1]), reinterpret_cast<const double *>(d + off[2]),
                                                        reinterpret_cast<const int *>(d + off[3]), reinterpret_cast<const double *>(d + off[4]),
                                                        reinterpret_cast<const int *>(d + off[5]), reinterpret_cast<const double *>(d + off[6]),
                                                        reinterpret_cast<const short *>(d + off[7]), g->coarse_nx, g->coarse_ny,
                                                        reinterpret_cast<const int *>(d + off[8]), d_samples);
    CK(cudaGetLastError());
    if (h_origin_zmax) {
        grx_count_launch();
        terrain_origin_kernel<<<(nt + 7) / 8, 256, 0, st>>>(d_samples, g->num_rows, g->num_cols, W, Lp, g->border, tot_cols, g->origin_x1, g->origin_x2,
                                                            g->origin_y1, g->origin_y2, reinterpret_cast<int *>(d + off[9]));
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(h_origin_zmax, d + off[9], 4 * (size_t)nt, cudaMemcpyDeviceToHost, st));
    }
    CK(cudaStreamSynchronize(st));
    cudaFree(d);
    return GRX_OK;
}
