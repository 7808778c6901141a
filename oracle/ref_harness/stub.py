"""TEST INFRASTRUCTURE — a fake ``isaacgym`` package so that the reference's UNMODIFIED env classes
(/root/reference/legged_gym/legged_gym/envs/**) can be imported and driven on CPU.

Runs only in the build container (needs /root/reference); its products are the committed golden
fixtures under tests/golden/.  ``FakeGym`` implements the ~45 gym methods the env touches
(SURVEY.md Appendix E); ``simulate()`` delegates to the C physics oracle (oracle/phys_impl.h), which
is a stand-in for PhysX — the *post-physics* arithmetic is the reference's own code.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

import numpy as np
import torch

REF = os.environ.get("GRX_REFERENCE_ROOT", "/root/reference")
IG = os.path.join(REF, "IsaacGym_Preview_4_Package", "isaacgym", "python", "isaacgym")


class Vec3:
    def __init__(self, x=0.0, y=0.0, z=0.0):
        self.x, self.y, self.z = float(x), float(y), float(z)


class Quat:
    def __init__(self, x=0.0, y=0.0, z=0.0, w=1.0):
        self.x, self.y, self.z, self.w = x, y, z, w


class Transform:
    def __init__(self, p=None, r=None):
        self.p = p if p is not None else Vec3()
        self.r = r if r is not None else Quat()


class _Bag:
    def __init__(self, **kw):
        self.__dict__.update(kw)


class SimParams(_Bag):
    def __init__(self):
        super().__init__(dt=1 / 60, substeps=2, up_axis=1, gravity=Vec3(0, 0, -9.81), use_gpu_pipeline=False,
                         physx=_Bag(use_gpu=False, num_subscenes=0, num_threads=0, solver_type=1,
                                    num_position_iterations=4, num_velocity_iterations=1, contact_offset=0.02,
                                    rest_offset=0.0, bounce_threshold_velocity=0.2, max_depenetration_velocity=100.0,
                                    max_gpu_contact_pairs=1024 * 1024, default_buffer_size_multiplier=2.0,
                                    contact_collection=2),
                         flex=_Bag())


class PlaneParams(_Bag):
    def __init__(self):
        super().__init__(normal=Vec3(0, 0, 1), static_friction=1.0, dynamic_friction=1.0, restitution=0.0)


class HeightFieldParams(_Bag):
    def __init__(self):
        super().__init__(column_scale=1.0, row_scale=1.0, vertical_scale=1.0, nbRows=0, nbColumns=0,
                         transform=Transform(), static_friction=1.0, dynamic_friction=1.0, restitution=0.0)


class TriangleMeshParams(_Bag):
    def __init__(self):
        super().__init__(nb_vertices=0, nb_triangles=0, transform=Transform(), static_friction=1.0,
                         dynamic_friction=1.0, restitution=0.0)


class AssetOptions(_Bag):
    pass


class CameraProperties(_Bag):
    pass


class _ShapeProps:
    def __init__(self):
        self.friction, self.restitution = 1.0, 0.0


class _BodyProps:
    def __init__(self, mass, com):
        self.mass, self.invMass, self.com = mass, (1.0 / mass if mass > 0 else 0.0), Vec3(*com)


class FakeGym:
    """CPU stand-in for the object returned by gymapi.acquire_gym() (base_task.py:42)."""

    def __init__(self):
        self.terrain = None
        self.envs = []
        self.tau = None
        self.phys = None
        self.sim_params = None
        self.n_simulate = 0

    # -- sim / terrain
    def create_sim(self, dev, gfx, engine, params):
        self.sim_params = params
        return "sim"

    def add_ground(self, sim, p):
        self.terrain = dict(heights=None, friction=p.static_friction, restitution=p.restitution)

    def add_heightfield(self, sim, samples, p):
        # legged_robot.py:881-898: nbRows = tot_cols, nbColumns = tot_rows (column-major); samples = [tot_rows, tot_cols]
        self.terrain = dict(heights=np.array(samples, dtype=np.int16).reshape(p.nbColumns, p.nbRows),
                            hscale=p.row_scale, vscale=p.vertical_scale, border=-p.transform.p.x,
                            friction=p.static_friction, restitution=p.restitution)

    def add_triangle_mesh(self, sim, verts, tris, p):
        self.terrain = dict(trimesh=(np.array(verts).reshape(-1, 3), np.array(tris).reshape(-1, 3)),
                            border=-p.transform.p.x, friction=p.static_friction, restitution=p.restitution)
        # The physics stand-in (like the product) resolves contacts of a structured trimesh on its sample grid + the vertex
        # shifts of the steep-edge snapping (top surface of the mesh; vertical walls carry no lateral contact); the caller is LeggedRobot._create_trimesh (legged_robot.py:906-921), whose Terrain object holds it.
        caller = sys._getframe(1).f_locals.get("self")
        ter = getattr(caller, "terrain", None)
        if ter is not None and hasattr(ter, "heightsamples"):
            from oracle.phys import moves_from_vertices
            hs_ = np.asarray(ter.heightsamples, dtype=np.int16)
            self.trimesh_as_heightfield = dict(heights=hs_, hscale=ter.cfg.horizontal_scale,
                                               vscale=ter.cfg.vertical_scale, border=-p.transform.p.x,
                                               friction=p.static_friction, restitution=p.restitution,
                                               # the vertex shifts of the mesh the reference actually uploaded (steep-edge snapping, terrain_utils.py:315-328)
                                               moves=moves_from_vertices(np.array(verts).reshape(-1, 3), hs_.shape[0], hs_.shape[1], ter.cfg.horizontal_scale))

    # -- asset
    def load_asset(self, sim, root, file, opts):
        sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..", "wiki-grx-gym_b200"))
        from grx_b200.urdf import compile_urdf
        self.model = compile_urdf(os.path.join(root, file))
        self.asset_opts = opts
        import xml.etree.ElementTree as ET
        self.n_shapes = len(ET.parse(os.path.join(root, file)).getroot().findall(".//collision"))
        self.shape_props = [_ShapeProps() for _ in range(self.n_shapes)]
        return "asset"

    def get_asset_dof_count(self, a):
        return self.model["nd"]

    def get_asset_rigid_body_count(self, a):
        return len(self.model["link_names"])

    def get_asset_dof_properties(self, a):
        m = self.model
        dt = np.dtype([("hasLimits", "?"), ("lower", "f4"), ("upper", "f4"), ("driveMode", "i4"), ("velocity", "f4"),
                       ("effort", "f4"), ("stiffness", "f4"), ("damping", "f4"), ("friction", "f4"), ("armature", "f4")])
        p = np.zeros(m["nd"], dtype=dt)
        p["hasLimits"], p["lower"], p["upper"] = True, m["dof_lower"], m["dof_upper"]
        p["velocity"], p["effort"], p["driveMode"] = m["dof_velocity"], m["dof_effort"], 3
        return p

    def get_asset_rigid_shape_properties(self, a):
        return self.shape_props

    def set_asset_rigid_shape_properties(self, a, props):
        self.shape_props = props

    def get_asset_rigid_body_names(self, a):
        return list(self.model["link_names"])

    def get_asset_dof_names(self, a):
        return list(self.model["dof_names"])

    # -- envs / actors
    def create_env(self, sim, lo, hi, per_row):
        self.envs.append(dict())
        return len(self.envs) - 1

    def create_actor(self, env, asset, pose, name, group, filt, seg):
        e = self.envs[env]
        e["pose"] = (pose.p.x, pose.p.y, pose.p.z)
        e["friction"] = float(self.shape_props[0].friction)
        e["restitution"] = float(self.shape_props[0].restitution)
        e["mass_scale"], e["com_offset"] = 1.0, (0.0, 0.0, 0.0)
        return 0

    def get_actor_name(self, e, a):
        return "actor"

    def get_actor_rigid_body_names(self, e, a):
        return list(self.model["link_names"])

    def get_actor_rigid_body_dict(self, e, a):
        return {n: i for i, n in enumerate(self.model["link_names"])}

    def get_actor_dof_names(self, e, a):
        return list(self.model["dof_names"])

    def get_actor_dof_dict(self, e, a):
        return {n: i for i, n in enumerate(self.model["dof_names"])}

    def set_actor_dof_properties(self, e, a, props):
        return True

    def get_actor_rigid_body_properties(self, e, a):
        r = self.model["root_link_inertial"]
        props = [_BodyProps(r[0], r[1:4])]
        props += [_BodyProps(1.0, (0, 0, 0)) for _ in range(len(self.model["link_names"]) - 1)]
        self._nominal_root = (r[0], tuple(r[1:4]))
        return props

    def set_actor_rigid_body_properties(self, e, a, props, recomputeInertia=False):
        m0, c0 = self._nominal_root
        self.envs[e]["mass_scale"] = props[0].mass / m0
        self.envs[e]["com_offset"] = (props[0].com.x - c0[0], props[0].com.y - c0[1], props[0].com.z - c0[2])
        return True

    def find_actor_rigid_body_handle(self, e, a, name):
        return self.model["link_names"].index(name)

    # -- tensors
    def prepare_sim(self, sim):
        from grx_b200.urdf import base_inertial_for
        from oracle.phys import PhysOracle
        N, nd, nl = len(self.envs), self.model["nd"], len(self.model["link_names"])
        self.root = torch.zeros(N, 13)
        self.root[:, 6] = 1.0
        for i, e in enumerate(self.envs):
            self.root[i, 0:3] = torch.tensor(e["pose"])
        self.dof_state = torch.zeros(N * nd, 2)
        self.contact = torch.zeros(N * nl, 3)
        self.rb = torch.zeros(N * nl, 13)
        self.friction = np.array([e["friction"] for e in self.envs], np.float32)
        self.restitution = np.array([e["restitution"] for e in self.envs], np.float32)
        bi = np.zeros((N, 10), np.float32)
        for i, e in enumerate(self.envs):
            m, c, I6 = base_inertial_for(self.model, e["mass_scale"], e["com_offset"])
            bi[i, 0], bi[i, 1:4], bi[i, 4:10] = m, c, I6
        self.base_inertial = bi
        from grx_b200.config import full_body_tables, make_cfg
        from grx_b200.robot import self_collision_pairs, task_tables
        sim_extra = {}
        if nd > 10:   # full-body tree: same contact priority AND the same self-collision candidate pairs as the product (legged_robot_config.py:121)
            tb = full_body_tables(self.model)
            ctl = dict(kp=np.zeros(nd), kd=np.zeros(nd), default_pos=np.zeros(nd), foot_links=[0], sph_order=tb["sph_order"],
                       self_pairs=self_collision_pairs(self.model, tb))
            sim_extra = dict(max_self_contacts=4)
        else:
            order = task_tables(self.model, make_cfg(self.model["name"]))["sph_order"]   # same contact priority as the product
            ctl = dict(kp=np.zeros(nd), kd=np.zeros(nd), default_pos=np.zeros(nd), foot_links=[0], sph_order=order)
        sp = self.sim_params
        terr = self.terrain
        if terr is not None and "trimesh" in terr:
            terr = getattr(self, "trimesh_as_heightfield", None)
            assert terr is not None, "set FakeGym.trimesh_as_heightfield before prepare_sim for trimesh terrains"
        self.phys = PhysOracle(self.model, ctl, terr, dtype=np.float32,
                               sim=dict(dt=sp.dt, contact_offset=sp.physx.contact_offset,
                                        bounce_threshold=sp.physx.bounce_threshold_velocity,
                                        max_depen_vel=sp.physx.max_depenetration_velocity,
                                        solver_iters=sp.physx.num_position_iterations, **sim_extra))
        self._refresh_links()
        return True

    def _refresh_links(self):
        N, nd = len(self.envs), self.model["nd"]
        ds = self.dof_state.view(N, nd, 2)
        ls = self.phys.link_states(self.root.numpy(), ds[..., 0].contiguous().numpy(), ds[..., 1].contiguous().numpy(),
                                   self.base_inertial)
        self.rb.view(N, -1, 13).copy_(torch.from_numpy(ls))

    def acquire_actor_root_state_tensor(self, sim):
        return self.root

    def acquire_dof_state_tensor(self, sim):
        return self.dof_state

    def acquire_net_contact_force_tensor(self, sim):
        return self.contact

    def acquire_rigid_body_state_tensor(self, sim):
        return self.rb

    def refresh_dof_state_tensor(self, sim):
        pass

    refresh_actor_root_state_tensor = refresh_net_contact_force_tensor = refresh_rigid_body_state_tensor = refresh_dof_state_tensor

    def set_dof_actuation_force_tensor(self, sim, tau):
        self.tau = tau
        return True

    def simulate(self, sim):
        N, nd = len(self.envs), self.model["nd"]
        ds = self.dof_state.view(N, nd, 2)
        q = ds[..., 0].contiguous().numpy()
        qd = ds[..., 1].contiguous().numpy()
        root = self.root.numpy()          # shares memory
        tau = self.tau.detach().reshape(N, nd).numpy()
        ls, cf = self.phys.substep(root, q, qd, tau, self.base_inertial, self.friction, self.restitution)
        ds[..., 0] = torch.from_numpy(q)
        ds[..., 1] = torch.from_numpy(qd)
        self.rb.view(N, -1, 13).copy_(torch.from_numpy(ls))
        self.contact.view(N, -1, 3).copy_(torch.from_numpy(cf))
        self.n_simulate += 1

    def fetch_results(self, sim, wait):
        pass

    # setters: env tensors alias sim state, so the (deferred) writes are already in place (tensors.rst.txt:391-396)
    def set_dof_state_tensor_indexed(self, sim, t, ids, n):
        return True

    def set_actor_root_state_tensor_indexed(self, sim, t, ids, n):
        return True

    def set_actor_root_state_tensor(self, sim, t):
        return True


def _load_by_path(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def install():
    """Put the fake ``isaacgym`` (+ real torch_utils / terrain_utils loaded by path) and the reference's
    ``legged_gym`` / ``rsl_rl`` on sys.path.  Idempotent."""
    if "isaacgym" in sys.modules and getattr(sys.modules["isaacgym"], "_grx_fake", False):
        return sys.modules["isaacgym"]
    if not hasattr(np, "float"):
        np.float = float                      # torch_utils.py:135 default arg
    pkg = types.ModuleType("isaacgym")
    pkg._grx_fake = True
    pkg.__path__ = []
    sys.modules["isaacgym"] = pkg
    gymapi = types.ModuleType("isaacgym.gymapi")
    for k, v in dict(Vec3=Vec3, Quat=Quat, Transform=Transform, SimParams=SimParams, PlaneParams=PlaneParams,
                     HeightFieldParams=HeightFieldParams, TriangleMeshParams=TriangleMeshParams,
                     AssetOptions=AssetOptions, CameraProperties=CameraProperties, SIM_PHYSX=1, SIM_FLEX=0,
                     UP_AXIS_Y=0, UP_AXIS_Z=1, KEY_ESCAPE=0, KEY_V=1).items():
        setattr(gymapi, k, v)
    gymapi.UpAxis = int
    gymapi.ContactCollection = int
    gymapi.acquire_gym = lambda *a: FakeGym()
    gymtorch = types.ModuleType("isaacgym.gymtorch")
    gymtorch.wrap_tensor = lambda t: t

    def unwrap_tensor(t):
        if not t.is_contiguous():
            raise Exception("Input tensor must be contiguous")   # gymtorch.py:97-99
        return t
    gymtorch.unwrap_tensor = unwrap_tensor
    gymutil = types.ModuleType("isaacgym.gymutil")

    def parse_device_str(s):
        if s == "cpu":
            return "cpu", 0
        return s.split(":")[0], int(s.split(":")[1]) if ":" in s else 0
    gymutil.parse_device_str = parse_device_str

    def parse_sim_config(d, sim_params):
        for k, v in d.items():
            if k == "physx":
                for kk, vv in v.items():
                    setattr(sim_params.physx, kk, vv)
            elif k == "gravity":
                sim_params.gravity = Vec3(*v)
            elif k != "flex":
                setattr(sim_params, k, v)
    gymutil.parse_sim_config = parse_sim_config
    for name, mod in (("gymapi", gymapi), ("gymtorch", gymtorch), ("gymutil", gymutil)):
        sys.modules["isaacgym." + name] = mod
        setattr(pkg, name, mod)
    # the REAL Isaac Gym python helpers, loaded by path
    pkg.torch_utils = _load_by_path("isaacgym.torch_utils", os.path.join(IG, "torch_utils.py"))
    from scipy import interpolate
    interpolate.interp2d = _interp2d_linear       # removed in SciPy 1.14 (still present as a raising stub); terrain_utils.py:44
    pkg.terrain_utils = _load_by_path("isaacgym.terrain_utils", os.path.join(IG, "terrain_utils.py"))
    if "matplotlib" not in sys.modules:
        try:
            import matplotlib.pyplot  # noqa: F401
        except Exception:
            mpl = types.ModuleType("matplotlib")
            mpl.pyplot = types.ModuleType("matplotlib.pyplot")
            sys.modules["matplotlib"], sys.modules["matplotlib.pyplot"] = mpl, mpl.pyplot
    for p in (os.path.join(REF, "legged_gym"), os.path.join(REF, "rsl_rl")):
        if p not in sys.path:
            sys.path.insert(0, p)
    return pkg


def _interp2d_linear(x, y, z, kind="linear"):
    """scipy.interpolate.interp2d(kind='linear') on a regular grid == bilinear interpolation."""
    from scipy.interpolate import RectBivariateSpline
    assert kind == "linear"
    spl = RectBivariateSpline(y, x, z, kx=1, ky=1, s=0)   # z[j, i] = f(x[i], y[j])
    return lambda xn, yn: spl(yn, xn)
