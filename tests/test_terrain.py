"""Our terrain generator == the reference's Terrain / terrain_utils output for the same numpy seed
(golden hashes from oracle/ref_harness/gen_terrain_golden.py, made with the reference's own classes)."""
import hashlib
import os

import numpy as np
import pytest

from grx_b200.config import make_cfg
from grx_b200.terrain import Terrain

G = dict(np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "terrain.npz")))


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.mark.parametrize("name", ["curr_full", "curr_small_trimesh", "random_small"])
def test_terrain_bit_identical_to_reference(name):
    rows, cols, curriculum, seed = [int(v) for v in G[name + "/params"]]
    cfg = make_cfg("GR1T1", 64, str(G[name + "/mesh"])).terrain
    cfg.num_rows, cfg.num_cols, cfg.curriculum = rows, cols, bool(curriculum)
    np.random.seed(seed)
    t = Terrain(cfg, 64)
    assert t.heightsamples.dtype == np.int16 and tuple(t.heightsamples.shape) == tuple(G[name + "/hf_shape"])
    np.testing.assert_array_equal(t.heightsamples[::7, ::7], G[name + "/hf_sub"])
    assert sha(t.heightsamples) == str(G[name + "/hf_sha"])
    np.testing.assert_array_equal(t.env_origins, G[name + "/env_origins"])
    if name + "/vert_sha" in G:
        assert t.vertices.dtype == np.float32 and t.triangles.dtype == np.uint32
        assert sha(t.vertices) == str(G[name + "/vert_sha"])
        assert sha(t.triangles) == str(G[name + "/tri_sha"])
