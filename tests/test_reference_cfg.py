"""Drop-in boundary, host side: the C task config derived from the REFERENCE's own cfg objects (GR1T1LowerLimbCfg /
GR1T2LowerLimbCfg, built by its task registry) equals the one derived from grx_b200.config.make_cfg — i.e. a maintainer can hand
`GRXVecEnv` the cfg object `task_registry.get_cfgs()` returns.  Needs the reference tree, which exists only in
the build container: skipped on the GPU box."""
import ctypes as C
import os

import numpy as np
import pytest

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "legged_gym")), reason="reference tree not present")


def _struct_bytes(s):
    return bytes(C.string_at(C.addressof(s), C.sizeof(s)))


@pytest.mark.parametrize("task", ["GR1T1", "GR1T2"])
@pytest.mark.parametrize("mesh", ["plane", "heightfield"])
def test_task_cfg_from_reference_cfg_object(task, mesh):
    from oracle.ref_harness import stub
    stub.install()
    import legged_gym.envs  # noqa: F401  (registers the tasks; isaacgym is the harness stub)
    from legged_gym.utils.task_registry import task_registry
    from grx_b200 import _lib as L
    from grx_b200.config import make_cfg
    from grx_b200.env import task_cfg
    from grx_b200.robot import task_tables
    from grx_b200.urdf import builtin_model
    ref_cfg, ref_train = task_registry.get_cfgs(task)
    ref_cfg.terrain.mesh_type = mesh
    ours = make_cfg(task, ref_cfg.env.num_envs, mesh)
    model = builtin_model(task)
    t_ref = task_cfg(ref_cfg, task_tables(model, ref_cfg), seed=1)
    t_our = task_cfg(ours, task_tables(model, ours), seed=1)
    for name, _ in L.TaskCfg._fields_:
        a, b = getattr(t_ref, name), getattr(t_our, name)
        if hasattr(a, "__len__"):
            np.testing.assert_allclose(np.ctypeslib.as_array(a), np.ctypeslib.as_array(b), rtol=1e-7, err_msg=name)
        else:
            assert a == pytest.approx(b, rel=1e-7), name
    assert _struct_bytes(t_ref) == _struct_bytes(t_our)
    # the PPO / runner dict too (gr1t1_config.py:310-345 via class_to_dict)
    from legged_gym.utils.helpers import class_to_dict
    from grx_b200.config import make_train_cfg
    d_ref, d_our = class_to_dict(ref_train), make_train_cfg(task)
    for sect in ("algorithm", "policy"):
        for k, v in d_our[sect].items():
            assert d_ref[sect][k] == v or (isinstance(v, float) and abs(d_ref[sect][k] - v) < 1e-12), (sect, k, d_ref[sect][k], v)
    assert d_ref["runner"]["num_steps_per_env"] == d_our["runner"]["num_steps_per_env"] == 64
