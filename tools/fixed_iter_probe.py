"""Probe: throughput when every env does exactly K Newton iterations (tolerance 0): what the lockstep barriers cost through iteration variance."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests")); sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import helpers as H
from mujoco_mjx_lab_b200 import training_utils
n = 262144
for K in [int(x) if x != "None" else None for x in (sys.argv[1].split(",") if len(sys.argv) > 1 else "None,2,3,4,5,6".split(","))]:
    model = H.load(overrides=None if K is None else dict(iterations=K, tolerance=-1.0))
    cfg = H.env_config()
    m, sysm, q0, nq, nv, nu, sr, ss, v_reset, v_step = training_utils.load_model_and_create_env("", cfg, model=model)
    g = torch.Generator(device="cuda").manual_seed(0)
    state, obs = v_reset(torch.from_numpy(H.ppo_keys(42, n).view(np.int32)).cuda())
    acts = [torch.randn(n, nu, device="cuda", generator=g).clamp_(-1, 1) for _ in range(4)]
    rk = [torch.randint(-2**31, 2**31 - 1, (n, 2), device="cuda", dtype=torch.int32, generator=g) for _ in range(4)]
    for i in range(40): state, obs, r, te, tr = v_step.autoreset(state, acts[i % 4], rk[i % 4], inplace=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(20): state, obs, r, te, tr = v_step.autoreset(state, acts[i % 4], rk[i % 4], inplace=True)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print(f"iterations={K}: {ms:.3f} ms/step {n / ms * 1e3 / 1e6:.2f} M steps/s")
    del v_step, v_reset, sysm
