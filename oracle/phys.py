"""TEST INFRASTRUCTURE — ctypes front end of the C physics oracle (oracle/phys_impl.h).

PARITY UNPINNED for physics (closed-source PhysX in the reference; see phys_impl.h).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference arm import this.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force=False):
    so = os.path.join(_HERE, "libgrx_oracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("phys_oracle.c", "phys_impl.h")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
    return _LIB


def near_moved_cells(moves):
    """[rows, cols] uint8: 1 where any vertex of the 3 x 3 cells around cell (i, j) — vertices (i-1 .. i+2, j-1 .. j+2) — is shifted."""
    m = (np.asarray(moves) != 0).any(-1)
    R, Cc = m.shape
    pad = np.zeros((R + 3, Cc + 3), bool)
    pad[1:R + 1, 1:Cc + 1] = m
    out = np.zeros((R, Cc), bool)
    for a in range(4):
        for b in range(4):
            out |= pad[a:a + R, b:b + Cc]
    return np.ascontiguousarray(out.astype(np.uint8))


def moves_from_vertices(vertices, rows, cols, hscale):
    """Vertex shifts (in cells) of a structured trimesh [rows * cols, 3] relative to its sample grid: the reference's steep-edge snapping."""
    v = np.asarray(vertices, np.float64).reshape(rows, cols, 3)
    gi, gj = np.meshgrid(np.arange(rows), np.arange(cols), indexing="ij")
    mx, my = np.rint(v[..., 0] / hscale - gi), np.rint(v[..., 1] / hscale - gj)
    assert np.abs(mx).max() <= 1 and np.abs(my).max() <= 1
    return np.stack([mx, my], -1).astype(np.int8)


def _mk_structs(real):
    P = C.POINTER(real)
    PI = C.POINTER(C.c_int)

    class Model(C.Structure):
        _fields_ = [("nb", C.c_int), ("nd", C.c_int), ("nl", C.c_int), ("ns", C.c_int), ("nf", C.c_int),
                    ("parent", PI), ("jpos", P), ("jrot", P), ("axis", P), ("mass", P), ("com", P), ("inertia", P),
                    ("dof_lower", P), ("dof_upper", P), ("dof_vel_limit", P), ("dof_effort", P),
                    ("link_body", PI), ("link_pos", P), ("link_rot", P),
                    ("sph_body", PI), ("sph_link", PI), ("sph_pos", P), ("sph_rad", P),
                    ("foot_links", PI), ("kp", P), ("kd", P), ("default_pos", P),
                    ("npairs", C.c_int), ("pair_a", PI), ("pair_b", PI)]

    class Terrain(C.Structure):
        _fields_ = [("type", C.c_int), ("rows", C.c_int), ("cols", C.c_int), ("heights", C.POINTER(C.c_short)),
                    ("hscale", real), ("vscale", real), ("border", real), ("friction", real), ("restitution", real),
                    ("moves", C.POINTER(C.c_byte)), ("near_moved", C.POINTER(C.c_ubyte))]

    class SimCfg(C.Structure):
        _fields_ = [("dt", real), ("gravity", real), ("contact_offset", real), ("bounce_threshold", real),
                    ("max_depen_vel", real), ("erp", real), ("solver_iters", C.c_int), ("decimation", C.c_int),
                    ("action_scale", real), ("max_contacts", C.c_int), ("max_self_contacts", C.c_int)]
    return Model, Terrain, SimCfg


class PhysOracle:
    """Physics oracle for one robot model / terrain / sim config, f32 or f64."""

    def __init__(self, model, ctl, terrain=None, sim=None, dtype=np.float64):
        """model: dict from grx_b200.urdf; ctl: dict(kp, kd, default_pos, foot_links);
        terrain: None (plane) or dict(heights int16 [rows, cols], hscale, vscale, border, friction, restitution);
        sim: dict overriding dt / decimation / solver_iters / erp / ..."""
        self.np_real = np.dtype(dtype)
        self.real = C.c_double if self.np_real == np.float64 else C.c_float
        self.sfx = "_f64" if self.np_real == np.float64 else "_f32"
        Model, Terrain, SimCfg = _mk_structs(self.real)
        self._keep = []
        m = Model()
        m.nb, m.nd, m.nl, m.ns = model["nb"], model["nd"], len(model["link_names"]), len(model["sph_rad"])
        fl = np.asarray(ctl["foot_links"], dtype=np.int32)
        m.nf = len(fl)
        order = np.asarray(ctl.get("sph_order", np.arange(len(model["sph_rad"]))))   # contact priority (robot.task_tables)
        model = dict(model, sph_body=model["sph_body"][order], sph_link=model["sph_link"][order],
                     sph_pos=model["sph_pos"][order], sph_rad=model["sph_rad"][order])
        for k in ("parent", "link_body", "sph_body", "sph_link"):
            setattr(m, k, self._iptr(model[k]))
        m.foot_links = self._iptr(fl)
        for k in ("jpos", "jrot", "axis", "mass", "com", "inertia", "dof_lower", "dof_upper", "dof_effort",
                  "link_pos", "link_rot", "sph_pos", "sph_rad"):
            setattr(m, k, self._rptr(model[k]))
        m.dof_vel_limit = self._rptr(model["dof_velocity"])
        m.kp, m.kd, m.default_pos = self._rptr(ctl["kp"]), self._rptr(ctl["kd"]), self._rptr(ctl["default_pos"])
        pairs = np.asarray(ctl.get("self_pairs", np.zeros((0, 2), np.int32)), np.int32).reshape(-1, 2)   # robot.self_collision_pairs (sphere indices in priority order)
        m.npairs = len(pairs)
        m.pair_a, m.pair_b = self._iptr(pairs[:, 0].copy()), self._iptr(pairs[:, 1].copy())
        self.model, self.m = model, m
        self.nd, self.nl, self.nf = m.nd, m.nl, m.nf
        t = Terrain()
        if terrain is None:
            t.type, t.friction, t.restitution = 0, 1.0, 0.0
            t.hscale, t.vscale, t.border = 1.0, 1.0, 0.0
        elif terrain.get("heights") is None:
            t.type, t.friction, t.restitution = 0, terrain.get("friction", 1.0), terrain.get("restitution", 0.0)
            t.hscale, t.vscale, t.border = 1.0, 1.0, 0.0
        else:
            hs = np.ascontiguousarray(terrain["heights"], dtype=np.int16)
            self._keep.append(hs)
            t.type, t.rows, t.cols = 1, hs.shape[0], hs.shape[1]
            t.heights = hs.ctypes.data_as(C.POINTER(C.c_short))
            if terrain.get("moves") is not None:   # structured trimesh: vertex shifts of the steep-edge snapping + the cells that can see a shifted vertex
                mv = np.ascontiguousarray(terrain["moves"], dtype=np.int8)
                assert mv.shape == hs.shape + (2,)
                near = near_moved_cells(mv)
                self._keep += [mv, near]
                t.type = 2
                t.moves, t.near_moved = mv.ctypes.data_as(C.POINTER(C.c_byte)), near.ctypes.data_as(C.POINTER(C.c_ubyte))
            t.hscale, t.vscale, t.border = terrain["hscale"], terrain["vscale"], terrain["border"]
            t.friction, t.restitution = terrain.get("friction", 1.0), terrain.get("restitution", 0.0)
        self.t = t
        s = SimCfg()
        d = dict(dt=0.002, gravity=-9.81, contact_offset=0.01, bounce_threshold=0.5, max_depen_vel=1.0, erp=0.2,
                 solver_iters=4, decimation=10, action_scale=1.0, max_contacts=8, max_self_contacts=0)
        d.update(sim or {})
        for k, v in d.items():
            setattr(s, k, v)
        self.s = s
        self.sim = d

    def _rptr(self, a):
        a = np.ascontiguousarray(a, dtype=self.np_real)
        self._keep.append(a)
        return a.ctypes.data_as(C.POINTER(self.real))

    def _iptr(self, a):
        a = np.ascontiguousarray(a, dtype=np.int32)
        self._keep.append(a)
        return a.ctypes.data_as(C.POINTER(C.c_int))

    def _p(self, a):
        assert a.dtype == self.np_real and a.flags.c_contiguous, (a.dtype, a.flags)
        return a.ctypes.data_as(C.POINTER(self.real))

    def step(self, root, dof_pos, dof_vel, actions, last_actions, delay, motor_strength, base_inertial, friction,
             restitution):
        """In place on root [N,13], dof_pos/dof_vel [N,nd].  Returns dict of outputs."""
        N = root.shape[0]
        r = self.np_real
        c = lambda a: np.ascontiguousarray(a, dtype=r)
        actions, last_actions, motor_strength = c(actions), c(last_actions), c(motor_strength)
        base_inertial, friction, restitution = c(base_inertial), c(friction), c(restitution)
        out = dict(torques=np.zeros((N, self.nd), r), link_state=np.zeros((N, self.nl, 13), r),
                   contact_force=np.zeros((N, self.nl, 3), r), avg_foot_force=np.zeros((N, self.nf), r),
                   avg_foot_linvel=np.zeros((N, self.nf, 3), r), avg_foot_angvel=np.zeros((N, self.nf, 3), r))
        out["active_sig"] = np.zeros((N, int(self.sim["decimation"])), np.uint64)   # active-set signature per substep (phys_impl.h substep)
        fn = getattr(lib(), "grx_oracle_physics_step_sig" + self.sfx)
        fn.restype = C.c_int
        err = fn(C.byref(self.m), C.byref(self.t), C.byref(self.s), C.c_int(N), self._p(root), self._p(dof_pos),
                 self._p(dof_vel), self._p(actions), self._p(last_actions), self.real(delay), self._p(motor_strength),
                 self._p(base_inertial), self._p(friction), self._p(restitution), self._p(out["torques"]),
                 self._p(out["link_state"]), self._p(out["contact_force"]), self._p(out["avg_foot_force"]),
                 self._p(out["avg_foot_linvel"]), self._p(out["avg_foot_angvel"]),
                 out["active_sig"].ctypes.data_as(C.POINTER(C.c_uint64)))
        if err:
            raise RuntimeError(f"physics oracle failed (code {err})")
        return out

    def substep(self, root, dof_pos, dof_vel, tau, base_inertial, friction, restitution):
        """One dt with given torques, in place.  Returns (link_state [N,nl,13], contact_force [N,nl,3])."""
        N = root.shape[0]
        r = self.np_real
        c = lambda a: np.ascontiguousarray(a, dtype=r)
        ls, cf = np.zeros((N, self.nl, 13), r), np.zeros((N, self.nl, 3), r)
        fn = getattr(lib(), "grx_oracle_substep" + self.sfx)
        fn.restype = C.c_int
        err = fn(C.byref(self.m), C.byref(self.t), C.byref(self.s), C.c_int(N), self._p(root), self._p(dof_pos),
                 self._p(dof_vel), self._p(c(tau)), self._p(c(base_inertial)), self._p(c(friction)),
                 self._p(c(restitution)), self._p(ls), self._p(cf))
        if err:
            raise RuntimeError(f"physics oracle failed (code {err})")
        return ls, cf

    def link_states(self, root, dof_pos, dof_vel, base_inertial):
        N = root.shape[0]
        r = self.np_real
        c = lambda a: np.ascontiguousarray(a, dtype=r)
        ls = np.zeros((N, self.nl, 13), r)
        fn = getattr(lib(), "grx_oracle_link_states" + self.sfx)
        fn(C.byref(self.m), C.c_int(N), self._p(c(root)), self._p(c(dof_pos)), self._p(c(dof_vel)),
           self._p(c(base_inertial)), self._p(ls))
        return ls

    def dynamics_terms(self, base_inertial, root, q, qd):
        r = self.np_real
        nv = self.nd + 6
        Mq, h, en = np.zeros((nv, nv), r), np.zeros(nv, r), np.zeros(2, r)
        c = lambda a: np.ascontiguousarray(a, dtype=r)
        fn = getattr(lib(), "grx_oracle_dynamics_terms" + self.sfx)
        fn(C.byref(self.m), C.byref(self.s), self._p(c(base_inertial)), self._p(c(root)), self._p(c(q)), self._p(c(qd)),
           self._p(Mq), self._p(h), self._p(en))
        return Mq, h, en
