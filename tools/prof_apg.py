"""Small driver for ncu: a few APG forward + reverse env steps (diff_step) at BASELINE configs[3]'s batch (2048 envs, CG 4/4)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from mujoco_mjx_lab_b200 import apg, parallel  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
cfg, env = apg.make_apg_env()
v_reset, v_step = env[8], env[9]
sysm = v_step.sys
state, _ = v_reset(torch.from_numpy(parallel.rank_keys(1, 0, n).view(np.int32)).cuda())
g = torch.Generator(device="cuda").manual_seed(0)
acts = [(0.3 * torch.randn(n, sysm.nu, device="cuda", generator=g)).requires_grad_() for _ in range(steps)]
total = 0.0
for t in range(steps):
    state, _, r, te, tr = apg.diff_step(sysm, state, acts[t])
    total = total + r.sum()
total.backward()
torch.cuda.synchronize()
print("done", float(total), float(acts[0].grad.abs().sum()))
