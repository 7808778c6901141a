"""The C-ABI library loads on a CPU-only box and exports every symbol include/grx_b200.h declares; the product path
fails loudly without a GPU (no CPU fallback)."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "grx_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(grx_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from grx_b200 import _lib
    lib = _lib.lib()
    names = _declared()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/grx_b200.h but not exported"
    assert sorted(_lib.EXPORTED) == names
    assert lib.grx_version() >= 100


def test_struct_layouts_match_header():
    from grx_b200 import _lib
    lib = _lib.lib()
    sizes = (ctypes.c_int32 * 5)()
    assert lib.grx_abi_sizes(sizes, 5) == 5
    assert list(sizes) == [ctypes.sizeof(_lib.Buffer), ctypes.sizeof(_lib.ModelDesc), ctypes.sizeof(_lib.TaskCfg),
                           ctypes.sizeof(_lib.InjectedPhysics), ctypes.sizeof(_lib.PPOCfg)]


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback():
    from grx_b200 import _lib
    from grx_b200.config import make_cfg
    from grx_b200.env import GRXVecEnv
    from grx_b200.ppo import PPO, ActorCriticMLP
    with pytest.raises(_lib.GrxError):
        GRXVecEnv(make_cfg("GR1T1", 8, "plane"))
    with pytest.raises(_lib.GrxError):
        PPO(ActorCriticMLP(39, 168, 10))
    # error convention of the C ABI: negative code + message
    assert _lib.lib().grx_env_get_buffer(None, b"obs", None) == -1
    assert b"null" in _lib.lib().grx_last_error()


def test_quat_rotate_inverse_host_helper():
    """GRXVecEnv.base_lin_vel / base_ang_vel (what play.py logs, play.py:110-123) use a host-side restatement of
    torch_utils.quat_rotate_inverse (torch_utils.py:72-81): check it against the rotation-matrix form R(q)^T v."""
    import numpy as np
    from grx_b200.env import quat_rotate_inverse
    g = torch.Generator().manual_seed(0)
    q = torch.randn(64, 4, generator=g, dtype=torch.float64)
    q = q / q.norm(dim=1, keepdim=True)
    v = torch.randn(64, 3, generator=g, dtype=torch.float64)
    x, y, z, w = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    R = torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w),
                     2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w),
                     2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)], dim=1).view(-1, 3, 3)
    want = torch.einsum("nji,nj->ni", R, v)
    np.testing.assert_allclose(quat_rotate_inverse(q, v).numpy(), want.numpy(), atol=1e-12)


def test_weight_gradient_plan_fills_whole_rounds():
    """Host logic of the tensor-core dispatch (tc::dw_plan), no GPU needed: for the registered minibatch (10 485 rows; the six weight gradients of
    actor + critic in one launch) the plan is 2 x 256 macro tiles x 14 splits = 140 tiles = ONE round of the 148 SMs (measured optimum,
    profiles/r3a_dw_tile_experiments.txt); in general the tile list never exceeds the planned number of whole rounds and every split is >= 128 rows."""
    import ctypes as C
    import numpy as np
    from grx_b200 import _lib as L
    lib = L.lib()

    def plan(shapes, K, sms=148):
        M = (C.c_int32 * len(shapes))(*[s[0] for s in shapes]); N = (C.c_int32 * len(shapes))(*[s[1] for s in shapes])
        out = (C.c_int32 * 3)()
        L.check(lib.grx_gemm_debug_dw_plan(M, N, K, len(shapes), sms, out))
        return tuple(out)
    six = [(128, 256), (128, 256), (256, 512), (256, 512), (512, 40), (512, 168)]
    assert plan(six, 10485) == (2, 256, 14)
    rng = np.random.default_rng(0)
    for _ in range(200):
        shapes = [six[i] for i in rng.choice(6, size=int(rng.integers(1, 7)), replace=True)]
        K, sms = int(rng.integers(200, 40000)), int(rng.choice([132, 148, 160]))
        tmt, bn, z = plan(shapes, K, sms)
        assert tmt in (1, 2) and bn in (128, 256) and z >= 1
        kchunk = -(-(-(-K // z)) // 32) * 32
        assert kchunk >= 128 or z == 1
        tiles = sum(-(-m // (128 * tmt)) * -(-n // bn) for m, n in shapes) * -(-K // kchunk)
        assert tiles <= 4 * sms
