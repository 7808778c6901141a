"""URDF -> flat dynamic-model arrays for the GRx env-step kernel.

Replaces what the reference obtains from the closed-source Isaac Gym URDF
importer (``gym.load_asset`` at legged_robot.py:966 with the asset options of
legged_robot_config.py:104-131): body/DOF enumeration, DOF limits
(legged_robot.py:594-604), per-link rigid bodies (``collapse_fixed_joints =
False`` -> every URDF link is reported) and primitive collision shapes.

Design (ours, not Isaac Gym's):
  * the *dynamic* model merges every fixed joint into its parent, giving a
    floating base + one body per revolute joint (lower-limb GR1T1: 11 bodies,
    16 velocity DOF);
  * every URDF link is still *reported*: it keeps (dynamic body, local pose) so
    pose / velocity / net contact force per link can be exported;
  * collision cylinders/spheres become contact spheres (see ``_spheres_of``).

Body/DOF order: depth-first from the root, children visited in alphabetical
order of the child link name.  This reproduces the DOF order implied by the
reference's action-limit tables (gr1t1_config.py:284-299: left leg, right leg,
waist, head, left arm, right arm).
"""
from __future__ import annotations

import json
import os
import xml.etree.ElementTree as ET

import numpy as np

__all__ = ["compile_urdf", "save_model", "load_model", "builtin_model", "MODEL_KEYS"]


def _floats(s, n):
    v = [float(x) for x in s.split()]
    assert len(v) == n, (s, n)
    return np.array(v, dtype=np.float64)


def rpy_to_mat(rpy):
    """URDF fixed-axis roll/pitch/yaw -> rotation matrix R = Rz(y) Ry(p) Rx(r)."""
    r, p, y = rpy
    cr, sr, cp, sp, cy, sy = np.cos(r), np.sin(r), np.cos(p), np.sin(p), np.cos(y), np.sin(y)
    return np.array([
        [cy * cp, cy * sp * sr - sy * cr, cy * sp * cr + sy * sr],
        [sy * cp, sy * sp * sr + cy * cr, sy * sp * cr - cy * sr],
        [-sp, cp * sr, cp * cr]])


def _origin(elem):
    o = elem.find("origin") if elem is not None else None
    if o is None:
        return np.zeros(3), np.eye(3)
    xyz = _floats(o.attrib.get("xyz", "0 0 0"), 3)
    rpy = _floats(o.attrib.get("rpy", "0 0 0"), 3)
    return xyz, rpy_to_mat(rpy)


def _sym(ixx, iyy, izz, ixy, ixz, iyz):
    return np.array([[ixx, ixy, ixz], [ixy, iyy, iyz], [ixz, iyz, izz]])


def _spheres_of(geom, pos, rot):
    """Contact-sphere stand-ins for a URDF collision primitive.

    cylinder (axis = local z): thin ones (L/2 > r) -> two spheres of radius r at
    the two ends of the axis; fat ones -> one sphere of radius r at the centre.
    sphere -> itself.  box -> 8 corner spheres of radius 0 are NOT needed by the
    GRx URDFs and are rejected loudly.
    """
    g = geom[0]
    if g.tag == "sphere":
        return [(pos, float(g.attrib["radius"]))]
    if g.tag == "cylinder":
        r = float(g.attrib["radius"])
        half = 0.5 * float(g.attrib["length"])
        if half > r:
            ax = rot[:, 2]
            return [(pos + ax * half, r), (pos - ax * half, r)]
        return [(pos, r)]
    raise ValueError(f"unsupported collision geometry <{g.tag}>")


def compile_urdf(path):
    """Parse ``path`` and return the flat model dict (numpy arrays + name lists)."""
    root = ET.parse(path).getroot()
    links = {}
    for l in root.findall("link"):
        name = l.attrib["name"]
        inert = l.find("inertial")
        if inert is not None:
            cpos, crot = _origin(inert)
            m = float(inert.find("mass").attrib["value"])
            ia = inert.find("inertia").attrib
            I = _sym(*[float(ia[k]) for k in ("ixx", "iyy", "izz", "ixy", "ixz", "iyz")])
            I = crot @ I @ crot.T
        else:
            cpos, m, I = np.zeros(3), 0.0, np.zeros((3, 3))
        sph = []
        for c in l.findall("collision"):
            p, R = _origin(c)
            sph += _spheres_of(c.find("geometry"), p, R)
        links[name] = dict(mass=m, com=cpos, inertia=I, spheres=sph)

    children = {n: [] for n in links}
    has_parent = set()
    for j in root.findall("joint"):
        p = j.find("parent").attrib["link"]
        c = j.find("child").attrib["link"]
        pos, rot = _origin(j)
        ax = j.find("axis")
        axis = _floats(ax.attrib["xyz"], 3) if ax is not None else np.array([1.0, 0, 0])
        lim = j.find("limit")
        jt = j.attrib["type"]
        if jt not in ("revolute", "fixed", "continuous"):
            raise ValueError(f"joint type {jt} not supported")
        children[p].append(dict(name=j.attrib["name"], type=jt, child=c, pos=pos, rot=rot, axis=axis,
                                lower=float(lim.attrib.get("lower", -1e30)) if lim is not None else -1e30,
                                upper=float(lim.attrib.get("upper", 1e30)) if lim is not None else 1e30,
                                effort=float(lim.attrib.get("effort", 0)) if lim is not None else 0.0,
                                velocity=float(lim.attrib.get("velocity", 1e30)) if lim is not None else 1e30))
        has_parent.add(c)
    roots = [n for n in links if n not in has_parent]
    assert len(roots) == 1, roots

    # ---- DFS (alphabetical children) over URDF links: dynamic bodies + reported links
    bodies = []      # dynamic bodies: dict(parent, jpos, jrot, axis, parts=[(m, com, I)])
    link_names, link_body, link_pos, link_rot = [], [], [], []
    dof = dict(names=[], lower=[], upper=[], effort=[], velocity=[])
    sph = dict(body=[], link=[], pos=[], rad=[])

    def visit(lname, body_idx, pos, rot):
        """pos/rot = pose of this link in the frame of dynamic body body_idx."""
        li = len(link_names)
        link_names.append(lname)
        link_body.append(body_idx)
        link_pos.append(pos.copy())
        link_rot.append(rot.copy())
        L = links[lname]
        if L["mass"] > 0:
            bodies[body_idx]["parts"].append((L["mass"], pos + rot @ L["com"], rot @ L["inertia"] @ rot.T, lname))
        for c, r in L["spheres"]:
            sph["body"].append(body_idx)
            sph["link"].append(li)
            sph["pos"].append(pos + rot @ c)
            sph["rad"].append(r)
        for j in sorted(children[lname], key=lambda d: d["child"]):
            jp, jr = pos + rot @ j["pos"], rot @ j["rot"]
            if j["type"] == "fixed":
                visit(j["child"], body_idx, jp, jr)
            else:
                ax = j["axis"] / np.linalg.norm(j["axis"])
                bodies.append(dict(parent=body_idx, jpos=jp, jrot=jr, axis=ax, parts=[]))
                dof["names"].append(j["name"])
                for k in ("lower", "upper", "effort", "velocity"):
                    dof[k].append(j[k])
                visit(j["child"], len(bodies) - 1, np.zeros(3), np.eye(3))

    bodies.append(dict(parent=-1, jpos=np.zeros(3), jrot=np.eye(3), axis=np.zeros(3), parts=[]))
    visit(roots[0], 0, np.zeros(3), np.eye(3))

    nb = len(bodies)
    mass, com, inertia = np.zeros(nb), np.zeros((nb, 3)), np.zeros((nb, 6))
    for b, B in enumerate(bodies):
        m, c, I = merge_inertials([(p[0], p[1], p[2]) for p in B["parts"]])
        mass[b], com[b] = m, c
        inertia[b] = [I[0, 0], I[1, 1], I[2, 2], I[0, 1], I[0, 2], I[1, 2]]
    # base split for domain randomisation of the root *link* only (legged_robot.py:618-648: props[0])
    root_part = [p for p in bodies[0]["parts"] if p[3] == roots[0]]
    rest_parts = [p for p in bodies[0]["parts"] if p[3] != roots[0]]
    rm, rc, rI = merge_inertials([(p[0], p[1], p[2]) for p in root_part])
    sm, sc, sI = merge_inertials([(p[0], p[1], p[2]) for p in rest_parts])

    model = dict(
        name=root.attrib.get("name", "robot"),
        nb=nb, nd=nb - 1, nv=nb + 5,
        parent=np.array([B["parent"] for B in bodies], dtype=np.int32),
        jpos=np.array([B["jpos"] for B in bodies]),
        jrot=np.array([B["jrot"].reshape(9) for B in bodies]),
        axis=np.array([B["axis"] for B in bodies]),
        mass=mass, com=com, inertia=inertia,
        root_link_inertial=np.concatenate([[rm], rc, _sym6(rI)]),
        root_rest_inertial=np.concatenate([[sm], sc, _sym6(sI)]),
        dof_names=dof["names"],
        dof_lower=np.array(dof["lower"]), dof_upper=np.array(dof["upper"]),
        dof_effort=np.array(dof["effort"]), dof_velocity=np.array(dof["velocity"]),
        link_names=link_names,
        link_body=np.array(link_body, dtype=np.int32),
        link_pos=np.array(link_pos), link_rot=np.array([r.reshape(9) for r in link_rot]),
        sph_body=np.array(sph["body"], dtype=np.int32), sph_link=np.array(sph["link"], dtype=np.int32),
        sph_pos=np.array(sph["pos"]).reshape(-1, 3), sph_rad=np.array(sph["rad"]),
    )
    return model


def _sym6(I):
    return np.array([I[0, 0], I[1, 1], I[2, 2], I[0, 1], I[0, 2], I[1, 2]])


def sym6_to_mat(s):
    return _sym(s[0], s[1], s[2], s[3], s[4], s[5])


def merge_inertials(parts):
    """Composite (mass, com, inertia about com) of rigidly attached parts."""
    m = sum(p[0] for p in parts)
    if m <= 0:
        return 0.0, np.zeros(3), np.zeros((3, 3))
    c = sum(p[0] * p[1] for p in parts) / m
    I = np.zeros((3, 3))
    for pm, pc, pI in parts:
        d = pc - c
        I += pI + pm * (d @ d * np.eye(3) - np.outer(d, d))
    return m, c, I


def base_inertial_for(model, mass_scale=1.0, com_offset=(0.0, 0.0, 0.0)):
    """Composite base inertial when the root *link* is randomised.

    Mirrors ``_process_rigid_body_props`` (legged_robot.py:618-648) followed by
    ``set_actor_rigid_body_properties(..., recomputeInertia=True)``
    (legged_robot.py:1080): root-link mass x scale (inertia scaled with it), COM
    shifted by ``com_offset``; the rest of the rigidly attached links unchanged.
    Returns (mass, com[3], inertia6) of dynamic body 0.
    """
    r, s = model["root_link_inertial"], model["root_rest_inertial"]
    parts = [(r[0] * mass_scale, r[1:4] + np.asarray(com_offset), sym6_to_mat(r[4:10]) * mass_scale)]
    if s[0] > 0:
        parts.append((s[0], s[1:4], sym6_to_mat(s[4:10])))
    m, c, I = merge_inertials(parts)
    return m, c, _sym6(I)


def base_inertials_batch(model, mass_scale, com_offset):
    """``base_inertial_for`` for N envs at once (no Python loop over envs): mass_scale [N], com_offset [N, 3] -> [N, 10] rows
    (mass, com xyz, inertia xx yy zz xy xz yz about the composite COM).  Same arithmetic, vectorised (parallel-axis theorem per part)."""
    mass_scale, com_offset = np.asarray(mass_scale, np.float64), np.asarray(com_offset, np.float64)
    N = mass_scale.shape[0]
    r, s = model["root_link_inertial"], model["root_rest_inertial"]
    m1 = r[0] * mass_scale                                               # [N]
    c1 = r[1:4][None, :] + com_offset                                    # [N, 3]
    I1 = sym6_to_mat(r[4:10])[None, :, :] * mass_scale[:, None, None]    # [N, 3, 3]
    m2, c2, I2 = float(s[0]), s[1:4][None, :], sym6_to_mat(s[4:10])[None, :, :]
    m = m1 + m2
    c = (m1[:, None] * c1 + m2 * c2) / m[:, None]
    eye = np.eye(3)[None, :, :]

    def shifted(pm, pc, pI):
        d = pc - c
        return pI + np.asarray(pm).reshape(-1, 1, 1) * ((d * d).sum(1)[:, None, None] * eye - d[:, :, None] * d[:, None, :])
    I = shifted(m1, c1, I1) + (shifted(np.full(N, m2), np.broadcast_to(c2, (N, 3)), I2) if m2 > 0 else 0.0)
    out = np.zeros((N, 10))
    out[:, 0], out[:, 1:4] = m, c
    out[:, 4], out[:, 5], out[:, 6], out[:, 7], out[:, 8], out[:, 9] = I[:, 0, 0], I[:, 1, 1], I[:, 2, 2], I[:, 0, 1], I[:, 0, 2], I[:, 1, 2]
    return out


MODEL_KEYS = ("parent", "jpos", "jrot", "axis", "mass", "com", "inertia", "root_link_inertial",
              "root_rest_inertial", "dof_lower", "dof_upper", "dof_effort", "dof_velocity", "link_body",
              "link_pos", "link_rot", "sph_body", "sph_link", "sph_pos", "sph_rad")


def save_model(model, path):
    out = {}
    for k, v in model.items():
        out[k] = v.tolist() if isinstance(v, np.ndarray) else v
    with open(path, "w") as f:
        json.dump(out, f, indent=0)


def load_model(path):
    with open(path) as f:
        d = json.load(f)
    for k in MODEL_KEYS:
        dt = np.int32 if k in ("parent", "link_body", "sph_body", "sph_link") else np.float64
        d[k] = np.array(d[k], dtype=dt)
    d["sph_pos"] = d["sph_pos"].reshape(-1, 3)
    return d


_ASSET_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "assets")


def builtin_model(name):
    """Pre-compiled model for 'GR1T1' / 'GR1T2' (lower-limb, the registered tasks;
    envs/__init__.py:40-55) or 'GR1T1_full' / 'GR1T2_full'."""
    fn = {"GR1T1": "gr1t1_lower_limb.json", "GR1T2": "gr1t2_lower_limb.json",
          "GR1T1_full": "gr1t1_full.json", "GR1T2_full": "gr1t2_full.json"}[name]
    return load_model(os.path.join(_ASSET_DIR, fn))
