"""Writes tests/golden/oracle_regression.npz: float64-oracle outputs on fixed seeds. These are REGRESSION vectors of this
repo's own CPU restatement (they pin the oracle against accidental change), NOT reference outputs: real-MJX vectors come
from tools/dump_mjx_golden.py on a machine that has mujoco-mjx."""
import os
import sys

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers as H  # noqa: E402

model = H.load()
orc = H.make_oracle(model, H.env_config())
out = {}
for kind in ("free", "lean", "tumble"):
    q, v, w, c = H.make_states(model, 8, 900, kind)
    f = orc.forward(q, v, w, c, prec="f64", debug=("qacc", "efc_force", "con_dist", "qM", "sensordata", "solver_niter", "efc_active"))
    s = orc.physics_step(q, v, w, None, c, prec="f64")
    for k in ("qacc", "efc_force", "con_dist", "sensordata", "solver_niter", "efc_active"):
        out[f"{kind}_{k}"] = f[k]
    out[f"{kind}_qM_diag"] = np.einsum("nii->ni", f["qM"])
    out[f"{kind}_qpos1"], out[f"{kind}_qvel1"] = s["qpos"], s["qvel"]
keys = H.ppo_keys(42, 8)
st, obs = orc.env_reset(keys, prec="f32")
out["reset_qpos"], out["reset_qvel"], out["reset_aux"], out["reset_obs"] = st["qpos"], st["qvel"], st["aux"], obs
act = np.clip(np.random.default_rng(0).normal(size=(8, 21)), -1, 1)
st1, obs1, r, te, tr, _, _ = orc.env_step(st, act, prec="f64")
out["step_obs"], out["step_reward"], out["step_qpos"] = obs1, r, st1["qpos"]
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "oracle_regression.npz"), **out)
print("wrote", len(out), "arrays")
