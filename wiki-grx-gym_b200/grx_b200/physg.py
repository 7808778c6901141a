"""Generic-topology dynamics on the GPU (csrc/grx_phys_generic.cu) behind a small Python handle: the full-body 32-DOF GR1T1 / GR1T2 trees
(gr1t1_config.py:10-307) with robot self-collision (legged_robot_config.py:121), or any other revolute tree the model compiler produces.
One ``step`` = the body of ``during_physics_step`` (legged_robot_fftai.py:51-88) for all envs.  No CPU fallback."""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib as L
from .env import model_desc
from .robot import self_collision_pairs


class PhysG:
    def __init__(self, model, tables, num_envs, device="cuda:0", terrain=None, sim=None, self_collision=True, max_self_contacts=4):
        """model / tables: grx_b200.urdf.builtin_model(...) + robot.task_tables(...); terrain: None (plane) or dict(heights int16 [rows, cols],
        hscale, vscale, border, friction, restitution); sim: overrides of dt / decimation / solver_iters / ..."""
        if not torch.cuda.is_available():
            raise L.GrxError("PhysG needs a CUDA device (no CPU fallback)")
        self.lib = L.lib()
        self.device = torch.device(device)
        self.model, self.tables, self.N = model, tables, num_envs
        self.nd, self.nl, self.nf = model["nd"], len(model["link_names"]), len(tables["foot_links"])
        d = dict(dt=0.002, gravity=-9.81, contact_offset=0.01, bounce_threshold=0.5, max_depen_vel=1.0, erp=0.2, solver_iters=4, decimation=10,
                 action_scale=1.0, max_contacts=8)
        d.update(sim or {})
        self.sim = d
        c = L.PhysGCfg()
        c.sim_dt, c.gravity, c.contact_offset, c.bounce_threshold = d["dt"], d["gravity"], d["contact_offset"], d["bounce_threshold"]
        c.max_depen_vel, c.erp, c.solver_iters, c.decimation = d["max_depen_vel"], d["erp"], d["solver_iters"], d["decimation"]
        c.action_scale, c.max_contacts = d["action_scale"], d["max_contacts"]
        self.pairs = self_collision_pairs(model, tables) if self_collision else np.zeros((0, 2), np.int32)
        c.max_self_contacts = max_self_contacts if self_collision else 0
        self.max_self_contacts = int(c.max_self_contacts)
        self._md, self._keep = model_desc(model, tables)
        pr = np.ascontiguousarray(self.pairs, np.int32)
        self._h = C.c_void_p()
        dev = self.device.index if self.device.index is not None else torch.cuda.current_device()
        L.check(self.lib.grx_physg_create(C.byref(self._md), pr.ctypes.data_as(L.PI), len(pr), C.byref(c), num_envs, dev, C.byref(self._h)))
        if terrain is not None and terrain.get("heights") is not None:
            hs = np.ascontiguousarray(terrain["heights"], np.int16)
            L.check(self.lib.grx_physg_set_terrain_heightfield(self._h, hs.ctypes.data_as(C.POINTER(C.c_int16)), hs.shape[0], hs.shape[1],
                                                               C.c_float(terrain["hscale"]), C.c_float(terrain["vscale"]), C.c_float(terrain["border"]),
                                                               C.c_float(terrain.get("friction", 1.0)), C.c_float(terrain.get("restitution", 0.0))))
        else:
            fr, rs = (terrain or {}).get("friction", 1.0), (terrain or {}).get("restitution", 0.0)
            L.check(self.lib.grx_physg_set_terrain_plane(self._h, C.c_float(fr), C.c_float(rs)))
        z = lambda *s: torch.zeros(*s, device=self.device)
        N, nd, nl, nf = num_envs, self.nd, self.nl, self.nf
        self.out = dict(torques=z(N, nd), link_state=z(N, nl, 13), contact_force=z(N, nl, 3), avg_foot_force=z(N, nf), avg_foot_linvel=z(N, nf, 3),
                        avg_foot_angvel=z(N, nf, 3))
        self.active_sig = torch.zeros(N, d["decimation"], dtype=torch.int64, device=self.device)

    def step(self, root, dof_pos, dof_vel, actions, last_actions, delay, motor_strength, base_inertial, friction, restitution):
        """All arguments: contiguous fp32 CUDA tensors; root / dof_pos / dof_vel are advanced in place.  Returns the dict of outputs."""
        p = lambda t: C.c_void_p(t.data_ptr())
        for t in (root, dof_pos, dof_vel, actions, last_actions, motor_strength, base_inertial, friction, restitution):
            assert t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()
        o = self.out
        L.check(self.lib.grx_physg_step(self._h, p(root), p(dof_pos), p(dof_vel), p(actions), p(last_actions), C.c_float(delay), p(motor_strength),
                                        p(base_inertial), p(friction), p(restitution), p(o["torques"]), p(o["link_state"]), p(o["contact_force"]),
                                        p(o["avg_foot_force"]), p(o["avg_foot_linvel"]), p(o["avg_foot_angvel"]), p(self.active_sig),
                                        C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)))
        return o

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self.lib.grx_physg_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
