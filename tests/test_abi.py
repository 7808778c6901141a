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
