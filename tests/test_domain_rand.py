"""Vectorised domain randomisation (grx_b200/robot.py:sample_domain_rand) vs the reference's per-env loop semantics
(legged_gym/envs/base/legged_robot.py:538-580 friction / restitution buckets, :618-648 base mass / COM, :1060-1064 motor strength):
same distributions (ranges, bucket structure, independence), checked statistically; the batched inertia composition equals the per-env one."""
import numpy as np
import pytest

from grx_b200.config import make_cfg
from grx_b200.robot import sample_domain_rand
from grx_b200.urdf import base_inertial_for, base_inertials_batch, builtin_model


@pytest.mark.parametrize("robot", ["GR1T1", "GR1T2"])
def test_batched_base_inertial_equals_per_env(robot):
    model = builtin_model(robot)
    g = np.random.default_rng(3)
    scale, off = g.uniform(0.8, 1.2, 50), g.uniform(-0.1, 0.1, (50, 3))
    bi = base_inertials_batch(model, scale, off)
    for e in range(50):
        m, c, I6 = base_inertial_for(model, scale[e], off[e])
        np.testing.assert_allclose(bi[e], np.concatenate([[m], c, I6]), rtol=1e-12, atol=1e-14)


def test_distributions_match_reference_ranges():
    cfg = make_cfg("GR1T2", 20000, "heightfield")
    dr = cfg.domain_rand
    for k in ("randomize_friction", "randomize_restitution", "randomize_motor_strength", "randomize_base_mass", "randomize_base_com"):
        assert getattr(dr, k), k                                    # BASELINE config #3: every flag on
    model = builtin_model("GR1T2")
    N = 20000
    p = sample_domain_rand(model, cfg, N, np.random.default_rng(11))
    # friction / restitution: at most 64 distinct bucket values inside the range, every env takes one of them (LR:538-580)
    for key, rng in (("friction", dr.friction_range), ("restitution", dr.restitution_range)):
        v = p[key]
        assert v.shape == (N,) and v.min() >= rng[0] and v.max() <= rng[1]
        u = np.unique(v)
        assert 48 <= len(u) <= 64                                   # 64 buckets, all but a few hit by 20000 uniform bucket ids
        counts = np.array([(v == x).sum() for x in u])
        assert abs(counts.mean() - N / len(u)) < 1e-6 and counts.max() < 2.0 * N / 64      # uniform bucket ids
    # motor strength: independent U(lo, hi) per env and DOF (LR:1060-1064)
    ms, (lo, hi) = p["motor_strength"], dr.multiply_motor_strength
    assert ms.shape == (N, model["nd"]) and ms.min() >= lo and ms.max() <= hi
    assert abs(ms.mean() - (lo + hi) / 2) < 2e-3 and abs(ms.var() - (hi - lo) ** 2 / 12) < 2e-4
    assert abs(np.corrcoef(ms[:, 0], ms[:, 1])[0, 1]) < 0.03
    # base mass x U(range), base COM + U(range) per axis, inertia recomputed (LR:618-648 + recomputeInertia at LR:1080)
    bi = p["base_inertial"]
    m0, c0, _ = base_inertial_for(model)
    root_m = model["root_link_inertial"][0]
    scale = (bi[:, 0] - (m0 - root_m)) / root_m
    lo, hi = dr.multiply_base_mass_range
    assert scale.min() >= lo - 1e-9 and scale.max() <= hi + 1e-9 and abs(scale.mean() - (lo + hi) / 2) < 2e-3
    rest_m = model["root_rest_inertial"][0]
    for ax, rng in enumerate((dr.add_base_com_range_x, dr.add_base_com_range_y, dr.add_base_com_range_z)):
        # composite COM = (m_root (c_root + off) + m_rest c_rest) / m  ->  off recovered exactly
        off = (bi[:, 1 + ax] * bi[:, 0] - rest_m * model["root_rest_inertial"][1 + ax]) / (root_m * scale) - model["root_link_inertial"][1 + ax]
        assert off.min() >= rng[0] - 1e-9 and off.max() <= rng[1] + 1e-9
        assert abs(off.mean() - (rng[0] + rng[1]) / 2) < 3e-3 and abs(off.var() - (rng[1] - rng[0]) ** 2 / 12) < 3e-4
    assert (np.linalg.eigvalsh(np.array([[bi[0, 4], bi[0, 7], bi[0, 8]], [bi[0, 7], bi[0, 5], bi[0, 9]], [bi[0, 8], bi[0, 9], bi[0, 6]]])) > 0).all()


def test_disabled_flags_give_nominal_parameters():
    cfg = make_cfg("GR1T1", 64, "plane")
    dr = cfg.domain_rand
    dr.randomize_friction = dr.randomize_restitution = dr.randomize_motor_strength = dr.randomize_base_mass = dr.randomize_base_com = False
    model = builtin_model("GR1T1")
    p = sample_domain_rand(model, cfg, 64, np.random.default_rng(0))
    m, c, I6 = base_inertial_for(model)
    assert (p["friction"] == 1.0).all() and (p["restitution"] == 0.0).all() and (p["motor_strength"] == 1.0).all()
    np.testing.assert_allclose(p["base_inertial"], np.tile(np.concatenate([[m], c, I6]), (64, 1)), rtol=1e-12)
