"""Development perf probe: env-step throughput (trajectory distribution B, auto-reset) and the speed-test cold step."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import helpers as H  # noqa: E402
from mujoco_mjx_lab_b200 import mjx, training_utils  # noqa: E402

model = H.load()
cfg = H.env_config()
m, sysm, q0, nq, nv, nu, single_reset, single_step, v_reset, v_step = training_utils.load_model_and_create_env("", cfg, model=model)
print("launch", v_step.sys.launch_config())
sizes = [int(x) for x in (sys.argv[1].split(",") if len(sys.argv) > 1 else "4096,65536,262144".split(","))]
for n in sizes:
    keys = torch.from_numpy(H.ppo_keys(42, n).view(np.int32)).cuda()
    state, obs = v_reset(keys)
    g = torch.Generator(device="cuda").manual_seed(0)
    acts = [torch.randn(n, nu, device="cuda", generator=g).clamp_(-1, 1) for _ in range(8)]
    rk = [torch.randint(-2**31, 2**31 - 1, (n, 2), device="cuda", dtype=torch.int32, generator=g) for _ in range(8)]
    nwarm, nt = 60, 40
    resets = 0.0
    for i in range(nwarm):
        state, obs, r, te, tr = v_step.autoreset(state, acts[i % 8], rk[i % 8], inplace=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(nt):
        state, obs, r, te, tr = v_step.autoreset(state, acts[i % 8], rk[i % 8], inplace=True)
        resets += 0
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / nt
    done_frac = float(torch.maximum(te, tr).mean())
    print(f"N={n:7d} env_step_autoreset {ms:8.3f} ms/step  {n / ms * 1e3:12.0f} steps/s  done_frac={done_frac:.4f} nan={bool(torch.isnan(obs).any())}")
    # physics-only stats on the current states
    d = state[0]
    dd = mjx.Data(d.qpos, d.qvel, d.qacc_warmstart, d.time, acts[0])
    _, out = mjx.forward(v_step.sys, dd, debug=True)
    cand = (out["efc_active"] & 1).sum(1).float()
    print(f"          mean newton iters {out['solver_niter'].float().mean():.2f} max {int(out['solver_niter'].max())}  cand rows mean {cand.mean():.1f} max {int(cand.max())} "
          f"status>0: {int((out['status'] != 0).sum())}")
    vel = torch.linspace(0, 1, n, device="cuda")
    for it in (1, 10):
        pos = mjx.speed_test(sysm, vel, it)
        torch.cuda.synchronize()
        e0.record()
        pos = mjx.speed_test(sysm, vel, it)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        print(f"          speed_test iters={it:3d}: {ms:8.3f} ms  {n * it / ms * 1e3:12.0f} steps/s")
    hist = torch.bincount(out["solver_niter"].flatten().long(), minlength=11).float()
    print("          niter histogram %:", [round(float(x) * 100 / hist.sum().item(), 1) for x in hist])
    rows_hist = torch.bincount((cand / 4).long().clamp(max=12), minlength=13).float()
    print("          cand rows/4 histogram %:", [round(float(x) * 100 / rows_hist.sum().item(), 1) for x in rows_hist])
