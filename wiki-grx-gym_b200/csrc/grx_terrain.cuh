// grx_terrain.cuh — terrain description on the device and the contact query shared by the fused lower-limb env kernel (grx_env.cu) and the
// generic-topology dynamics kernel (grx_phys_generic.cu): plane, heightfield (cells split along the diagonal of the reference's trimesh
// conversion) and structured trimesh (heightfield + the shifted vertices of the steep-edge snapping).  oracle/phys_impl.h terrain_query is the
// CPU statement of the same arithmetic.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace grx {   // plain data shared across translation units

struct TerrainDev {
    int type, rows, cols;   // 0 plane, 1 heightfield, 2 structured trimesh (heightfield + snapped vertices)
    const short *h;
    float hscale, vscale, border, friction, restitution;
    const signed char *mv;          // type 2: [rows, cols, 2] vertex shifts (cells) of the steep-edge snapping (terrain_utils.py:315-328)
    const unsigned char *near_mv;   // type 2: [rows, cols] != 0 where a vertex of the 3 x 3 cells around cell (i, j) is shifted
};

}  // namespace grx

namespace {
using grx::TerrainDev;

// splitmix64 finaliser: item hash of the active-set signature (same function in oracle/phys_impl.h)
__device__ __forceinline__ unsigned long long mix64(unsigned long long x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
// cell (debug signature only): grid cell (i, j) and which of its two triangles the query landed in
__device__ __forceinline__ void terrain_query(const TerrainDev &t, float x, float y, float &h, float *n, int *cell) {
    if (t.type == 0) { h = 0; n[0] = 0; n[1] = 0; n[2] = 1; cell[0] = cell[1] = cell[2] = 0; return; }
    float gx = (x + t.border) / t.hscale, gy = (y + t.border) / t.hscale;
    int i = (int)floorf(gx), j = (int)floorf(gy);
    if (i < 0) { i = 0; gx = 0; }
    if (j < 0) { j = 0; gy = 0; }
    if (i > t.rows - 2) { i = t.rows - 2; gx = (float)(t.rows - 1); }
    if (j > t.cols - 2) { j = t.cols - 2; gy = (float)(t.cols - 1); }
    float fx = gx - (float)i, fy = gy - (float)j;
    const short *p = t.h + (size_t)i * t.cols + j;
    float h00 = __ldg(p) * t.vscale, h01 = __ldg(p + 1) * t.vscale;
    float h10 = __ldg(p + t.cols) * t.vscale, h11 = __ldg(p + t.cols + 1) * t.vscale;
    float dhx, dhy;
    if (fx >= fy) { dhx = h10 - h00; dhy = h11 - h10; }
    else { dhx = h11 - h01; dhy = h01 - h00; }
    cell[0] = i; cell[1] = j; cell[2] = fx >= fy ? 0 : 1;
    h = h00 + dhx * fx + dhy * fy;
    float sx = -dhx / t.hscale, sy = -dhy / t.hscale;
    float inv = 1.0f / sqrtf(sx * sx + sy * sy + 1.0f);
    n[0] = sx * inv; n[1] = sy * inv; n[2] = inv;
    if (t.type != 2 || !__ldg(t.near_mv + (size_t)i * t.cols + j)) return;
    // Structured trimesh (gym.add_triangle_mesh of terrain_utils.convert_heightfield_to_trimesh): vertices next to a steep edge are shifted
    // sideways by one cell — a flat tread + a vertical wall instead of the heightfield's ramp.  Top surface under (x, y) = the highest of the
    // triangles of the 3 x 3 cells around the nominal cell whose projection contains the point (oracle/phys_impl.h terrain_query, same arithmetic).
    float best = -1e30f;
#pragma unroll 1
    for (int d = 0; d < 9; d++) {
        const int ci = i + d / 3 - 1, cj = j + d % 3 - 1;
        if (ci < 0 || cj < 0 || ci > t.rows - 2 || cj > t.cols - 2) continue;
        float P[4][3];   // P00 P10 P01 P11
#pragma unroll
        for (int v = 0; v < 4; v++) {
            const int vi = ci + (v & 1), vj = cj + (v >> 1);
            const size_t id = (size_t)vi * t.cols + vj;
            P[v][0] = (float)(vi + (int)__ldg(t.mv + 2 * id)); P[v][1] = (float)(vj + (int)__ldg(t.mv + 2 * id + 1)); P[v][2] = __ldg(t.h + id) * t.vscale;
        }
#pragma unroll
        for (int k = 0; k < 2; k++) {   // k = 0: (P00, P11, P01), k = 1: (P00, P10, P11)
            const float *a = P[0], *b = k ? P[1] : P[3], *c = k ? P[3] : P[2];
            const float e1x = b[0] - a[0], e1y = b[1] - a[1], e2x = c[0] - a[0], e2y = c[1] - a[1];
            const float det = e1x * e2y - e2x * e1y;
            if (fabsf(det) < 1e-6f) continue;
            const float px = gx - a[0], py = gy - a[1];
            const float u = (px * e2y - e2x * py) / det, w = (e1x * py - px * e1y) / det;
            if (u < -1e-5f || w < -1e-5f || u + w > 1.00001f) continue;
            const float hh = a[2] + u * (b[2] - a[2]) + w * (c[2] - a[2]);
            if (hh > best) {
                best = hh;
                float nx = (e1y * (c[2] - a[2]) - (b[2] - a[2]) * e2y) * t.hscale, ny = ((b[2] - a[2]) * e2x - e1x * (c[2] - a[2])) * t.hscale,
                      nz = det * t.hscale * t.hscale;
                if (nz < 0) { nx = -nx; ny = -ny; nz = -nz; }
                const float il = 1.0f / sqrtf(nx * nx + ny * ny + nz * nz);
                h = hh; n[0] = nx * il; n[1] = ny * il; n[2] = nz * il;
                cell[0] = ci; cell[1] = cj; cell[2] = 2 + k;
            }
        }
    }
}

}  // namespace
