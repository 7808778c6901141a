"""Compile the four GRx URDFs into the committed model assets.

Run in the build container (needs /root/reference or --root pointing at a
legged_gym/resources/robots directory):  python tools/compile_models.py
"""
import argparse, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "wiki-grx-gym_b200"))
from grx_b200.urdf import compile_urdf, save_model

ap = argparse.ArgumentParser()
ap.add_argument("--root", default="/root/reference/legged_gym/resources/robots")
a = ap.parse_args()
out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "wiki-grx-gym_b200", "grx_b200", "assets")
for robot, urdf, name in [("GR1T1", "GR1T1_lower_limb.urdf", "gr1t1_lower_limb.json"), ("GR1T2", "GR1T2_lower_limb.urdf", "gr1t2_lower_limb.json"),
                          ("GR1T1", "GR1T1.urdf", "gr1t1_full.json"), ("GR1T2", "GR1T2.urdf", "gr1t2_full.json")]:
    m = compile_urdf(os.path.join(a.root, robot, "urdf", urdf))
    save_model(m, os.path.join(out, name))
    print(name, "nb", m["nb"], "mass", m["mass"].sum(), "links", len(m["link_names"]), "spheres", len(m["sph_rad"]))
    print("   dofs:", m["dof_names"])
