"""Dump per-env features of consecutive steps (trajectory distribution) to study predictors of the next step's Newton iteration count
(the cost key of the work-sorted schedule). Writes gpurun_out/niter_dump.npz."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import helpers as H
from mujoco_mjx_lab_b200 import mjx, training_utils
model, cfg = H.load(), H.env_config()
env = training_utils.load_model_and_create_env("", cfg, model=model)
v_reset, v_step, nu = env[8], env[9], env[5]
n = 65536
g = torch.Generator(device="cuda").manual_seed(0)
state, obs = v_reset(torch.from_numpy(H.ppo_keys(42, n).view(np.int32)).cuda())
act = lambda: torch.randn(n, nu, device="cuda", generator=g).clamp_(-1, 1)
rk = lambda: torch.randint(-2**31, 2**31 - 1, (n, 2), device="cuda", dtype=torch.int32, generator=g)
for i in range(60): state, obs, r, te, tr = v_step.autoreset(state, act(), rk(), inplace=True)
rec = {k: [] for k in ("niter", "ncand", "nact", "z", "vnorm", "stance", "ep", "wsnorm", "done")}
for i in range(6):
    a = act()
    d, aux = state
    _, out = mjx.forward(v_step.sys, mjx.Data(d.qpos, d.qvel, d.qacc_warmstart, d.time, a), debug=True)
    rec["niter"].append(out["solver_niter"].cpu().numpy())
    rec["ncand"].append((out["efc_active"] & 1).sum(1).cpu().numpy())
    rec["nact"].append(((out["efc_active"] >> 1) & 1).sum(1).cpu().numpy())
    rec["z"].append(d.qpos[:, 2].cpu().numpy()); rec["vnorm"].append(d.qvel.norm(dim=1).cpu().numpy())
    rec["wsnorm"].append(d.qacc_warmstart.norm(dim=1).cpu().numpy())
    rec["stance"].append(aux[:, 5].cpu().numpy()); rec["ep"].append(aux[:, 8].cpu().numpy())
    state, obs, r, te, tr = v_step.autoreset(state, a, rk(), inplace=True)
    rec["done"].append(torch.maximum(te, tr).cpu().numpy())
os.makedirs("gpurun_out", exist_ok=True)
np.savez_compressed("gpurun_out/niter_dump.npz", **{k: np.stack(v) for k, v in rec.items()})
print("ok")
