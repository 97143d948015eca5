"""Small driver for ncu: settles a batch into the trajectory distribution, then launches N env steps (and speed-test steps)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import helpers as H  # noqa: E402
from mujoco_mjx_lab_b200 import mjx, training_utils  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
settle = int(sys.argv[2]) if len(sys.argv) > 2 else 60
model, cfg = H.load(), H.env_config()
m, sysm, q0, nq, nv, nu, single_reset, single_step, v_reset, v_step = training_utils.load_model_and_create_env("", cfg, model=model)
g = torch.Generator(device="cuda").manual_seed(0)
state, obs = v_reset(torch.from_numpy(H.ppo_keys(42, n).view(np.int32)).cuda())
acts = [torch.randn(n, nu, device="cuda", generator=g).clamp_(-1, 1) for _ in range(4)]
rk = [torch.randint(-2**31, 2**31 - 1, (n, 2), device="cuda", dtype=torch.int32, generator=g) for _ in range(4)]
for i in range(settle + 3):
    state, obs, r, te, tr = v_step.autoreset(state, acts[i % 4], rk[i % 4], inplace=True)
torch.cuda.synchronize()
vel = torch.linspace(0, 1, n, device="cuda")
pos = mjx.speed_test(sysm, vel, 1)
torch.cuda.synchronize()
print("done", float(obs.sum()), float(pos.sum()))
