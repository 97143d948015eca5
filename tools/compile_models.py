"""Regenerates mujoco_mjx_lab_b200/data/*.json from the reference's MJCF files (run in the build container, where
/root/reference exists; the GPU box only has the committed JSON)."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from mujoco_mjx_lab_b200 import modelc  # noqa: E402

REF = os.environ.get("MJXB_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "mujoco_mjx_lab_b200", "data")
for name in ("humanoid_mjx", "humanoid"):
    m = modelc.compile_mjcf(os.path.join(REF, "models", f"{name}.xml"))
    modelc.save_model(m, os.path.join(OUT, f"{name}.json"))
    print(name, {k: m[k] for k in ("nq", "nv", "nu", "nbody", "ngeom", "npair", "ncon", "nefc")}, "mass", m["body_mass"].sum())
