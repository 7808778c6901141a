"""TEST INFRASTRUCTURE — golden terrain fixtures from the reference's Terrain class (build container only).
python -m oracle.ref_harness.gen_terrain_golden"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "wiki-grx-gym_b200"))
from oracle.ref_harness import stub  # noqa: E402


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    stub.install()
    import contextlib, io
    with contextlib.redirect_stdout(io.StringIO()):
        import legged_gym.envs  # noqa: F401 (breaks the import cycle)
    from legged_gym.utils.terrain import Terrain as RefTerrain
    from grx_b200.config import make_cfg
    out = {}
    for name, mesh, rows, cols, curriculum, seed in [("curr_full", "heightfield", 10, 20, True, 1),
                                                     ("curr_small_trimesh", "trimesh", 3, 5, True, 7),
                                                     ("random_small", "heightfield", 4, 6, False, 3)]:
        cfg = make_cfg("GR1T1", 64, mesh).terrain
        cfg.num_rows, cfg.num_cols, cfg.curriculum = rows, cols, curriculum
        np.random.seed(seed)
        t = RefTerrain(cfg, 64)
        out[name + "/params"] = np.array([rows, cols, int(curriculum), seed])
        out[name + "/mesh"] = np.array(mesh)
        out[name + "/hf_sha"] = np.array(sha(t.heightsamples))
        out[name + "/hf_shape"] = np.array(t.heightsamples.shape)
        out[name + "/hf_sub"] = t.heightsamples[::7, ::7].copy()
        out[name + "/env_origins"] = t.env_origins.copy()
        if mesh == "trimesh":
            out[name + "/vert_sha"] = np.array(sha(t.vertices))
            out[name + "/tri_sha"] = np.array(sha(t.triangles))
        print(name, t.heightsamples.shape, out[name + "/hf_sha"])
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "terrain.npz"), **out)


if __name__ == "__main__":
    main()
