"""BASELINE.json configs[0..1]: step-throughput sweep 64 / 1K ... 256K envs on one GPU, both input distributions of SURVEY.md 8d:
(A) speed-test cold step (mjx_humanoid_speed_test.py semantics, `iters` repeats in one launch), (B) trajectory (v_step + auto-reset).
Writes gpurun_out/sweep.json (copy it to profiles/)."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import helpers as H  # noqa: E402
from mujoco_mjx_lab_b200 import mjx, training_utils  # noqa: E402

model, cfg = H.load(), H.env_config()
m, sysm, q0, nq, nv, nu, single_reset, single_step, v_reset, v_step = training_utils.load_model_and_create_env("", cfg, model=model)
rows = []
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for n in (64, 1024, 2048, 4096, 8192, 16384, 32768, 65536, 131072, 262144):
    g = torch.Generator(device="cuda").manual_seed(0)
    state, obs = v_reset(torch.from_numpy(H.ppo_keys(42, n).view(np.int32)).cuda())
    acts = [torch.randn(n, nu, device="cuda", generator=g).clamp_(-1, 1) for _ in range(4)]
    rk = [torch.randint(-2**31, 2**31 - 1, (n, 2), device="cuda", dtype=torch.int32, generator=g) for _ in range(4)]
    for i in range(64):
        state, obs, r, te, tr = v_step.autoreset(state, acts[i % 4], rk[i % 4], inplace=True)
    torch.cuda.synchronize()
    nt = 200 if n <= 16384 else 50
    e0.record()
    for i in range(nt):
        state, obs, r, te, tr = v_step.autoreset(state, acts[i % 4], rk[i % 4], inplace=True)
    e1.record()
    torch.cuda.synchronize()
    ms_b = e0.elapsed_time(e1) / nt
    vel = torch.linspace(0, 1, n, device="cuda")
    iters = 200 if n <= 16384 else 20
    mjx.speed_test(sysm, vel, 3)
    torch.cuda.synchronize()
    e0.record()
    pos = mjx.speed_test(sysm, vel, iters)
    e1.record()
    torch.cuda.synchronize()
    ms_a = e0.elapsed_time(e1) / iters
    rows.append(dict(n_env=n, traj_ms_per_step=ms_b, traj_steps_per_s=n / ms_b * 1e3, speedtest_ms_per_step=ms_a, speedtest_steps_per_s=n / ms_a * 1e3))
    print(rows[-1])
os.makedirs("gpurun_out", exist_ok=True)
json.dump(dict(gpu=torch.cuda.get_device_name(0), rows=rows,
               note="(A) speed-test: in-kernel loop of cold steps from qpos0 with qvel[0]=linspace(0,1,N); (B) trajectory: v_step + fused auto-reset, "
                    "host launches one kernel triple per step (no CUDA graph)"), open("gpurun_out/sweep.json", "w"), indent=1)
