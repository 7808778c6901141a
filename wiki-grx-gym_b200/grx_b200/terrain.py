"""Procedural terrain for the GRx tasks: int16 heightfield, per-tile env origins, optional trimesh.

Produces the same arrays as the reference's ``Terrain`` (legged_gym/utils/terrain.py:38-164) built on
Isaac Gym's ``terrain_utils`` (isaacgym/terrain_utils.py:17-350) — including the *same sequence of
numpy.random calls*, so that with the same ``np.random.seed`` the heightfield is bit-identical (pinned by
tests/test_terrain.py against tests/golden/terrain_*.npz generated from the reference).  The code is a
vectorised restatement: tiles are written straight into the global grid, the trimesh is built without
per-row Python loops.  Only the tile types reachable with the reference's 5-entry
``terrain_proportions`` are implemented (smooth slope, rough slope, stairs up/down, discrete obstacles).
"""
from __future__ import annotations

import numpy as np


class SubTile:
    """One tile of the grid (terrain_utils.py:353-360 ``SubTerrain``): int16 ``height_field_raw[width, length]``."""

    def __init__(self, width, length, vertical_scale, horizontal_scale):
        self.width, self.length = width, length
        self.vertical_scale, self.horizontal_scale = vertical_scale, horizontal_scale
        self.height_field_raw = np.zeros((width, length), dtype=np.int16)


def _pyramid_slope(t, slope, platform_size):
    """terrain_utils.py:74-106."""
    cx, cy = int(t.width / 2), int(t.length / 2)
    rx = ((cx - np.abs(cx - np.arange(t.width))) / cx).reshape(t.width, 1)
    ry = ((cy - np.abs(cy - np.arange(t.length))) / cy).reshape(1, t.length)
    peak = int(slope * (t.horizontal_scale / t.vertical_scale) * (t.width / 2))
    t.height_field_raw += (peak * rx * ry).astype(t.height_field_raw.dtype)
    half = int(platform_size / t.horizontal_scale / 2)
    x1, y1 = t.width // 2 - half, t.length // 2 - half
    corner = t.height_field_raw[x1, y1]
    t.height_field_raw = np.clip(t.height_field_raw, min(corner, 0), max(corner, 0))


def _bilinear_upsample(z, n_out_x, n_out_y, span_x, span_y):
    """Bilinear interpolation of samples z[i, j] placed on linspace(0, span, n) grids, evaluated on
    linspace(0, span, n_out) grids: what scipy.interpolate.interp2d(kind='linear') computed at terrain_utils.py:41-48."""
    nx, ny = z.shape
    gx = np.linspace(0, span_x, n_out_x) / span_x * (nx - 1)
    gy = np.linspace(0, span_y, n_out_y) / span_y * (ny - 1)
    i0 = np.clip(np.floor(gx).astype(int), 0, nx - 2)
    j0 = np.clip(np.floor(gy).astype(int), 0, ny - 2)
    fx, fy = (gx - i0)[:, None], (gy - j0)[None, :]
    z00, z10 = z[i0][:, j0], z[i0 + 1][:, j0]
    z01, z11 = z[i0][:, j0 + 1], z[i0 + 1][:, j0 + 1]
    return (z00 * (1 - fx) + z10 * fx) * (1 - fy) + (z01 * (1 - fx) + z11 * fx) * fy


def _random_uniform(t, min_height, max_height, step, downsampled_scale):
    """terrain_utils.py:17-51; one np.random.choice call of shape (w*hs/ds, l*hs/ds)."""
    lo, hi, st = int(min_height / t.vertical_scale), int(max_height / t.vertical_scale), int(step / t.vertical_scale)
    levels = np.arange(lo, hi + st, st)
    nx = int(t.width * t.horizontal_scale / downsampled_scale)
    ny = int(t.length * t.horizontal_scale / downsampled_scale)
    coarse = np.random.choice(levels, (nx, ny))
    fine = _bilinear_upsample(coarse.astype(np.float64), t.width, t.length, t.width * t.horizontal_scale,
                              t.length * t.horizontal_scale)
    t.height_field_raw += np.rint(fine).astype(np.int16)


def _pyramid_stairs(t, step_width, step_height, platform_size):
    """terrain_utils.py:195-227: concentric square rings, each one step higher (or lower)."""
    sw, sh = int(step_width / t.horizontal_scale), int(step_height / t.vertical_scale)
    plat = int(platform_size / t.horizontal_scale)
    x0, x1, y0, y1, h = 0, t.width, 0, t.length, 0
    while (x1 - x0) > plat and (y1 - y0) > plat:
        x0, x1, y0, y1, h = x0 + sw, x1 - sw, y0 + sw, y1 - sw, h + sh
        t.height_field_raw[x0:x1, y0:y1] = h


def _discrete_obstacles(t, max_height, min_size, max_size, num_rects, platform_size):
    """terrain_utils.py:109-149; five np.random.choice calls per rectangle, in the reference's order."""
    mh = int(max_height / t.vertical_scale)
    lo, hi = int(min_size / t.horizontal_scale), int(max_size / t.horizontal_scale)
    plat = int(platform_size / t.horizontal_scale)
    ni, nj = t.height_field_raw.shape
    heights = [-mh, -mh // 2, mh // 2, mh]
    sizes = range(lo, hi, 4)
    for _ in range(num_rects):
        w = np.random.choice(sizes)
        l = np.random.choice(sizes)
        si = np.random.choice(range(0, ni - w, 4))
        sj = np.random.choice(range(0, nj - l, 4))
        t.height_field_raw[si:si + w, sj:sj + l] = np.random.choice(heights)
    x1, x2 = (t.width - plat) // 2, (t.width + plat) // 2
    y1, y2 = (t.length - plat) // 2, (t.length + plat) // 2
    t.height_field_raw[x1:x2, y1:y2] = 0


class Terrain:
    """Same public attributes as the reference class: ``height_field_raw`` / ``heightsamples`` (int16
    [tot_rows, tot_cols]), ``env_origins`` [num_rows, num_cols, 3], ``tot_rows``, ``tot_cols``, ``border``,
    ``env_length``, ``env_width`` and, for ``mesh_type == 'trimesh'``, ``vertices`` / ``triangles``."""

    def __init__(self, cfg, num_robots):
        self.cfg, self.num_robots, self.type = cfg, num_robots, cfg.mesh_type
        if self.type in ("none", "plane"):
            return
        self.env_length, self.env_width = cfg.terrain_length, cfg.terrain_width
        self.proportions = [np.sum(cfg.terrain_proportions[:i + 1]) for i in range(len(cfg.terrain_proportions))]
        self.env_origins = np.zeros((cfg.num_rows, cfg.num_cols, 3))
        self.width_per_env_pixels = int(self.env_width / cfg.horizontal_scale)
        self.length_per_env_pixels = int(self.env_length / cfg.horizontal_scale)
        self.border = int(cfg.border_size / cfg.horizontal_scale)
        self.tot_cols = int(cfg.num_cols * self.width_per_env_pixels) + 2 * self.border
        self.tot_rows = int(cfg.num_rows * self.length_per_env_pixels) + 2 * self.border
        self.height_field_raw = np.zeros((self.tot_rows, self.tot_cols), dtype=np.int16)
        if cfg.curriculum:
            # terrain.py:85-92: difficulty by row, type by column; column-major visiting order fixes the RNG stream
            for j in range(cfg.num_cols):
                for i in range(cfg.num_rows):
                    self._place(self._make_tile(j / cfg.num_cols + 0.001, i / cfg.num_rows), i, j)
        elif getattr(cfg, "selected", False):
            raise NotImplementedError("terrain.selected is broken upstream (terrain.py:94-107) and not supported")
        else:
            # terrain.py:75-83
            for k in range(cfg.num_rows * cfg.num_cols):
                i, j = np.unravel_index(k, (cfg.num_rows, cfg.num_cols))
                choice = np.random.uniform(0, 1)
                difficulty = np.random.choice([0.5, 0.75, 0.9])
                self._place(self._make_tile(choice, difficulty), i, j)
        self.heightsamples = self.height_field_raw
        if self.type == "trimesh":
            self.vertices, self.triangles = heightfield_to_trimesh(self.height_field_raw, cfg.horizontal_scale,
                                                                   cfg.vertical_scale, cfg.slope_treshold)

    def _make_tile(self, choice, difficulty):
        """terrain.py:109-145 (tile is width x width, as upstream)."""
        c = self.cfg
        t = SubTile(self.width_per_env_pixels, self.width_per_env_pixels, c.vertical_scale, c.horizontal_scale)
        slope, step_h, obst_h = difficulty * 0.4, 0.05 + 0.18 * difficulty, 0.05 + difficulty * 0.2
        p = self.proportions
        if choice < p[0]:
            _pyramid_slope(t, -slope if choice < p[0] / 2 else slope, 3.0)
        elif choice < p[1]:
            _pyramid_slope(t, slope, 3.0)
            _random_uniform(t, -0.05, 0.05, 0.005, 0.2)
        elif choice < p[3]:
            _pyramid_stairs(t, 0.31, -step_h if choice < p[2] else step_h, 3.0)
        elif choice < p[4]:
            _discrete_obstacles(t, obst_h, 1.0, 2.0, 20, 3.0)
        else:
            raise NotImplementedError("tile types beyond terrain_proportions[4] (stepping stones / gap / pit)")
        return t

    def _place(self, t, i, j):
        """terrain.py:147-164."""
        L, W, hs = self.length_per_env_pixels, self.width_per_env_pixels, t.horizontal_scale
        self.height_field_raw[self.border + i * L: self.border + (i + 1) * L,
                              self.border + j * W: self.border + (j + 1) * W] = t.height_field_raw
        x1, x2 = int((self.env_length / 2.0 - 1) / hs), int((self.env_length / 2.0 + 1) / hs)
        y1, y2 = int((self.env_width / 2.0 - 1) / hs), int((self.env_width / 2.0 + 1) / hs)
        z = np.max(t.height_field_raw[x1:x2, y1:y2]) * t.vertical_scale
        self.env_origins[i, j] = [(i + 0.5) * self.env_length, (j + 0.5) * self.env_width, z]


class DeviceTerrain(Terrain):
    """The same terrain generated ON THE GPU (csrc/grx_terrain_gen.cu, C ABI grx_terrain_generate): the numpy random stream is drawn here in
    the reference's call order (a few thousand draws, so the grid stays bit-identical to the reference's for the same np.random.seed), the
    array arithmetic runs as one kernel over the grid.  ``heights_dev`` is the int16 CUDA tensor [tot_rows, tot_cols] (handed to the env
    without a host round trip, grx_env_set_terrain_device); ``heightsamples`` / ``height_field_raw`` are fetched lazily for callers that
    want the numpy array.  ``env_origins`` as in the host class.  No CPU fallback: raises without the CUDA library / a CUDA device."""

    def __init__(self, cfg, num_robots, device="cuda:0"):
        import ctypes as C
        import torch
        from . import _lib as L
        self.cfg, self.num_robots, self.type = cfg, num_robots, cfg.mesh_type
        if self.type in ("none", "plane"):
            return
        if not torch.cuda.is_available():
            raise L.GrxError("DeviceTerrain needs a CUDA device (no CPU fallback; grx_b200.terrain.Terrain is the host generator)")
        self.env_length, self.env_width = cfg.terrain_length, cfg.terrain_width
        self.proportions = [np.sum(cfg.terrain_proportions[:i + 1]) for i in range(len(cfg.terrain_proportions))]
        self.env_origins = np.zeros((cfg.num_rows, cfg.num_cols, 3))
        W = self.width_per_env_pixels = int(self.env_width / cfg.horizontal_scale)
        Lp = self.length_per_env_pixels = int(self.env_length / cfg.horizontal_scale)
        self.border = int(cfg.border_size / cfg.horizontal_scale)
        self.tot_cols = int(cfg.num_cols * W) + 2 * self.border
        self.tot_rows = int(cfg.num_rows * Lp) + 2 * self.border
        hs, vs = cfg.horizontal_scale, cfg.vertical_scale
        tiles = (L.TerrainTile * (cfg.num_rows * cfg.num_cols))()
        coarse, rects = [], []
        nx, ny = int(W * hs / 0.2), int(W * hs / 0.2)                                    # random_uniform_terrain(downsampled_scale=0.2)

        def draw(choice, difficulty, i, j):
            """_make_tile (terrain.py:109-145) reduced to its random draws + integer parameters."""
            d = tiles[i * cfg.num_cols + j]
            slope, step_h, obst_h = difficulty * 0.4, 0.05 + 0.18 * difficulty, 0.05 + difficulty * 0.2
            p = self.proportions
            if choice < p[1]:
                s = -slope if choice < p[0] / 2 else slope
                d.kind = 0 if choice < p[0] else 1
                d.slope_peak = int(s * (hs / vs) * (W / 2))
                d.plat_lo = W // 2 - int(3.0 / hs / 2)
                if d.kind == 1:
                    lo, hi, st = int(-0.05 / vs), int(0.05 / vs), int(0.005 / vs)
                    d.coarse_index = len(coarse)
                    coarse.append(np.random.choice(np.arange(lo, hi + st, st), (nx, ny)))
            elif choice < p[3]:
                d.kind = 2
                sw, sh = int(0.31 / hs), int((-step_h if choice < p[2] else step_h) / vs)
                plat = int(3.0 / hs)
                x0, x1, n = 0, W, 0
                while (x1 - x0) > plat:
                    x0, x1, n = x0 + sw, x1 - sw, n + 1
                d.step_width, d.step_height, d.num_rings = sw, sh, n
            elif choice < p[4]:
                d.kind = 3
                mh = int(obst_h / vs)
                lo, hi, plat = int(1.0 / hs), int(2.0 / hs), int(3.0 / hs)
                heights, sizes = [-mh, -mh // 2, mh // 2, mh], range(lo, hi, 4)
                d.rect_index, d.num_rects = len(rects), 20
                for _ in range(20):
                    w = np.random.choice(sizes)
                    l = np.random.choice(sizes)
                    si = np.random.choice(range(0, W - w, 4))
                    sj = np.random.choice(range(0, W - l, 4))
                    rects.append((si, sj, w, l, np.random.choice(heights)))
                d.plat_lo, d.plat_hi = (W - plat) // 2, (W + plat) // 2
            else:
                raise NotImplementedError("tile types beyond terrain_proportions[4] (stepping stones / gap / pit)")

        if cfg.curriculum:                                                                # terrain.py:85-92 (column-major visiting order)
            for j in range(cfg.num_cols):
                for i in range(cfg.num_rows):
                    draw(j / cfg.num_cols + 0.001, i / cfg.num_rows, i, j)
        elif getattr(cfg, "selected", False):
            raise NotImplementedError("terrain.selected is broken upstream (terrain.py:94-107) and not supported")
        else:                                                                             # terrain.py:75-83
            for k in range(cfg.num_rows * cfg.num_cols):
                i, j = np.unravel_index(k, (cfg.num_rows, cfg.num_cols))
                choice = np.random.uniform(0, 1)
                difficulty = np.random.choice([0.5, 0.75, 0.9])
                draw(choice, difficulty, i, j)
        # float64 tables of pyramid_sloped_terrain / interp2d, exactly as the host generator computes them
        cx = int(W / 2)
        ramp = np.ascontiguousarray((cx - np.abs(cx - np.arange(W))) / cx, np.float64)
        g = np.linspace(0, W * hs, W) / (W * hs) * (nx - 1)
        i0 = np.clip(np.floor(g).astype(int), 0, nx - 2)
        fx = np.ascontiguousarray(g - i0, np.float64)
        i0 = np.ascontiguousarray(i0, np.int32)
        co = np.ascontiguousarray(np.stack(coarse) if coarse else np.zeros((0, nx, ny)), np.int16)
        rc = np.ascontiguousarray(np.array(rects, np.int32).reshape(-1, 5))
        grid = L.TerrainGrid(cfg.num_rows, cfg.num_cols, W, Lp, self.border, nx, ny,
                             int((self.env_length / 2.0 - 1) / hs), int((self.env_length / 2.0 + 1) / hs),
                             int((self.env_width / 2.0 - 1) / hs), int((self.env_width / 2.0 + 1) / hs))
        dev = torch.device(device)
        self.heights_dev = torch.empty(self.tot_rows, self.tot_cols, dtype=torch.int16, device=dev)
        zmax = np.zeros(cfg.num_rows * cfg.num_cols, np.int32)
        PD = C.POINTER(C.c_double)
        L.check(L.lib().grx_terrain_generate(C.byref(grid), tiles, ramp.ctypes.data_as(PD), ramp.ctypes.data_as(PD), i0.ctypes.data_as(L.PI),
                                             fx.ctypes.data_as(PD), i0.ctypes.data_as(L.PI), fx.ctypes.data_as(PD),
                                             co.ctypes.data_as(C.POINTER(C.c_int16)), len(co), rc.ctypes.data_as(L.PI), len(rc),
                                             C.c_void_p(self.heights_dev.data_ptr()), zmax.ctypes.data_as(L.PI),
                                             dev.index if dev.index is not None else torch.cuda.current_device(),
                                             C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)))
        for i in range(cfg.num_rows):
            for j in range(cfg.num_cols):
                z = np.int16(zmax[i * cfg.num_cols + j]) * vs
                self.env_origins[i, j] = [(i + 0.5) * self.env_length, (j + 0.5) * self.env_width, z]
        self._host = None

    @property
    def heightsamples(self):
        if self._host is None:
            self._host = self.heights_dev.cpu().numpy()
        return self._host

    height_field_raw = heightsamples

    @property
    def vertices(self):
        raise AttributeError("DeviceTerrain builds the structured trimesh on the device (grx_env_set_terrain_device); use heightfield_to_trimesh(heightsamples, ...) for the host mesh")


def heightfield_to_trimesh(hf, horizontal_scale, vertical_scale, slope_threshold=None):
    """Structured 2-triangles-per-cell mesh with the reference's steep-edge vertex snapping
    (terrain_utils.py:286-350).  Returns (vertices float32 [R*C, 3], triangles uint32 [2(R-1)(C-1), 3])."""
    R, C = hf.shape
    xx = np.repeat(np.linspace(0, (R - 1) * horizontal_scale, R)[:, None], C, axis=1)
    yy = np.repeat(np.linspace(0, (C - 1) * horizontal_scale, C)[None, :], R, axis=0)
    if slope_threshold is not None:
        thr = slope_threshold * horizontal_scale / vertical_scale
        h = hf   # differences stay in the array dtype (int16), as upstream
        mx, my, mc = np.zeros((R, C)), np.zeros((R, C)), np.zeros((R, C))
        mx[:R - 1, :] += (h[1:, :] - h[:R - 1, :] > thr)
        mx[1:, :] -= (h[:R - 1, :] - h[1:, :] > thr)
        my[:, :C - 1] += (h[:, 1:] - h[:, :C - 1] > thr)
        my[:, 1:] -= (h[:, :C - 1] - h[:, 1:] > thr)
        mc[:R - 1, :C - 1] += (h[1:, 1:] - h[:R - 1, :C - 1] > thr)
        mc[1:, 1:] -= (h[:R - 1, :C - 1] - h[1:, 1:] > thr)
        xx += (mx + mc * (mx == 0)) * horizontal_scale
        yy += (my + mc * (my == 0)) * horizontal_scale
    v = np.zeros((R * C, 3), dtype=np.float32)
    v[:, 0], v[:, 1], v[:, 2] = xx.reshape(-1), yy.reshape(-1), hf.reshape(-1) * vertical_scale
    i0 = (np.arange(R - 1)[:, None] * C + np.arange(C - 1)[None, :]).reshape(-1)
    tri = np.empty((2 * (R - 1) * (C - 1), 3), dtype=np.uint32)
    tri[0::2] = np.stack([i0, i0 + C + 1, i0 + 1], axis=1)
    tri[1::2] = np.stack([i0, i0 + C, i0 + C + 1], axis=1)
    return v, tri
