"""Quick GPU-vs-oracle stage comparison (development aid; the real checks live in tests/)."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import helpers as H  # noqa: E402
from mujoco_mjx_lab_b200 import mjx  # noqa: E402

model = H.load()
orc = H.make_oracle(model)
sysm = mjx.put_model(model)
print("launch config", sysm.launch_config())
res = {}
for kind in ("free", "stand", "lean", "tumble"):
    n = 256
    q, v, w, c = H.make_states(model, n, 1, kind)
    ref = orc.forward(q, v, w, c, prec="f64", debug=True)
    ref32 = orc.forward(q, v, w, c, prec="f32", debug=True)
    t = lambda a: torch.tensor(a, dtype=torch.float32, device="cuda")
    d = mjx.Data(t(q), t(v), t(w), torch.zeros(n, device="cuda"), t(c))
    nd, out = mjx.forward(sysm, d, debug=True)
    torch.cuda.synchronize()
    r = {}
    for name in ("xpos", "xquat", "qM", "qfrc_bias", "qfrc_passive", "qfrc_actuator", "qacc_smooth", "con_dist", "con_pos",
                 "efc_pos", "efc_D", "efc_aref", "efc_force", "qacc", "qfrc_constraint", "sensordata"):
        g = out[name].double().cpu().numpy()
        a = ref[name]
        scale = np.maximum(1.0, np.abs(a))
        r[name] = [float(np.max(np.abs(g - a) / scale)), float(np.max(np.abs(ref32[name] - a) / scale))]
    ga = out["efc_active"].cpu().numpy()
    r["cand_mismatch"] = int(((ga & 1) != (ref["efc_active"] & 1)).sum())
    r["act_mismatch"] = int(((ga >> 1) != (ref["efc_active"] >> 1)).sum())
    r["cand_mean"] = float((ref["efc_active"] & 1).sum(1).mean())
    r["cand_max"] = int((ref["efc_active"] & 1).sum(1).max())
    r["niter_gpu"] = float(out["solver_niter"].float().mean())
    r["niter_ref64"] = float(ref["solver_niter"].mean())
    r["niter_ref32"] = float(ref32["solver_niter"].mean())
    r["status"] = int(out["status"].max())
    r["nan"] = bool(torch.isnan(out["qacc"]).any())
    res[kind] = r
    print(kind, json.dumps(r, indent=None))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/quick_parity.json", "w"), indent=1)
