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


@pytest.mark.parametrize("task", ["GR1T1", "GR1T2"])
def test_full_body_task_cfg_from_reference_cfg_object(task):
    """The same for the UNREGISTERED full-body configuration: the reference's own `GR1T1Cfg()` / `GR1T2Cfg()` object (gr1t1_config.py:10-307,
    gr1t2_config.py; observation sizes corrected and reward scales set by oracle/ref_harness/driver.py:full_body_cfg, because upstream leaves
    num_obs stale and every scale at zero) yields the same C task config, model tables and self-collision pairs as
    grx_b200.config.make_full_body_cfg — i.e. INTEGRATION.md's note on the full-body task holds for the reference's cfg object itself."""
    from oracle.ref_harness import stub
    stub.install()
    from oracle.ref_harness.driver import full_body_cfg
    from grx_b200 import _lib as L
    from grx_b200.config import make_full_body_cfg
    from grx_b200.env import task_cfg
    from grx_b200.robot import self_collision_pairs, task_tables
    from grx_b200.urdf import builtin_model
    ref_cfg = full_body_cfg(task)
    ours = make_full_body_cfg(task, ref_cfg.env.num_envs, ref_cfg.terrain.mesh_type)
    assert (ref_cfg.env.num_obs, ref_cfg.env.num_pri_obs, ref_cfg.env.num_actions) == (105, 234, 32) == (ours.env.num_obs, ours.env.num_pri_obs, ours.env.num_actions)
    model = builtin_model(task + "_full")
    tb_ref, tb_our = task_tables(model, ref_cfg), task_tables(model, ours)
    for k in tb_our:
        np.testing.assert_allclose(np.asarray(tb_ref[k], float), np.asarray(tb_our[k], float), rtol=1e-7, err_msg=k)
    np.testing.assert_array_equal(self_collision_pairs(model, tb_ref), self_collision_pairs(model, tb_our))
    t_ref, t_our = task_cfg(ref_cfg, tb_ref, seed=1), task_cfg(ours, tb_our, seed=1)
    for name, _ in L.TaskCfg._fields_:
        a, b = getattr(t_ref, name), getattr(t_our, name)
        if hasattr(a, "__len__"):
            np.testing.assert_allclose(np.ctypeslib.as_array(a), np.ctypeslib.as_array(b), rtol=1e-6, err_msg=name)
        else:
            assert a == pytest.approx(b, rel=1e-6), name
