"""Terrain generated on the device (csrc/grx_terrain_gen.cu through grx_terrain_generate) == the reference's Terrain / terrain_utils output for the
same numpy seed: bit-identical int16 grid and env origins vs the golden hashes made with the reference's own classes
(oracle/ref_harness/gen_terrain_golden.py) and vs the host generator (grx_b200/terrain.py) on further seeds / shapes."""
import hashlib
import os

import numpy as np
import pytest
import torch

from grx_b200.config import make_cfg
from grx_b200.terrain import DeviceTerrain, Terrain

pytestmark = pytest.mark.gpu

G = dict(np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "terrain.npz")))


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.mark.parametrize("name", ["curr_full", "curr_small_trimesh", "random_small"])
def test_device_terrain_bit_identical_to_reference(name):
    rows, cols, curriculum, seed = [int(v) for v in G[name + "/params"]]
    cfg = make_cfg("GR1T1", 64, str(G[name + "/mesh"])).terrain
    cfg.num_rows, cfg.num_cols, cfg.curriculum = rows, cols, bool(curriculum)
    np.random.seed(seed)
    t = DeviceTerrain(cfg, 64, "cuda:0")
    assert t.heights_dev.is_cuda and t.heights_dev.dtype == torch.int16
    hs = t.heightsamples
    assert tuple(hs.shape) == tuple(G[name + "/hf_shape"])
    np.testing.assert_array_equal(hs[::7, ::7], G[name + "/hf_sub"])
    assert sha(hs) == str(G[name + "/hf_sha"])
    np.testing.assert_array_equal(t.env_origins, G[name + "/env_origins"])


@pytest.mark.parametrize("seed,rows,cols,curriculum", [(3, 10, 20, True), (11, 4, 7, False), (12, 5, 5, False), (99, 2, 9, True)])
def test_device_terrain_equals_host_generator(seed, rows, cols, curriculum):
    """Every tile type (smooth / rough slope, stairs up / down, obstacles) at several difficulties; the random draws after generation must be
    at the same position of the numpy stream as after the host generator (same number of draws consumed)."""
    cfg = make_cfg("GR1T1", 64, "heightfield").terrain
    cfg.num_rows, cfg.num_cols, cfg.curriculum = rows, cols, curriculum
    np.random.seed(seed)
    h = Terrain(cfg, 64)
    after_host = np.random.uniform()
    np.random.seed(seed)
    d = DeviceTerrain(cfg, 64, "cuda:0")
    after_dev = np.random.uniform()
    np.testing.assert_array_equal(d.heightsamples, h.heightsamples)
    np.testing.assert_array_equal(d.env_origins, h.env_origins)
    assert after_host == after_dev


def test_env_on_device_terrain_equals_env_on_host_terrain():
    """GRXVecEnv with the device generator + device trimesh builder steps exactly like the env fed with the host arrays."""
    from grx_b200.env import GRXVecEnv
    outs = []
    for gen in ("device", "host"):
        cfg = make_cfg("GR1T1", 256, "trimesh")
        cfg.terrain.num_rows, cfg.terrain.num_cols = 4, 6
        cfg.terrain.max_init_terrain_level = 3
        env = GRXVecEnv(cfg, sim_device="cuda:0", terrain_generator=gen)
        env.reset()
        g = torch.Generator(device="cuda").manual_seed(0)
        for _ in range(20):
            obs, pri, rew, reset, _ = env.step(0.3 * torch.randn(256, 10, device="cuda", generator=g), delay=3.0)
        torch.cuda.synchronize()
        outs.append((obs.clone(), pri.clone(), rew.clone(), env.root_states.clone()))
        env.close()
    for a, b in zip(*outs):
        assert torch.equal(a, b)
