"""Small case for compute-sanitizer: reset + a few fused steps (with auto-reset and an overflowing env) + speed test on 96 envs."""
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
n = 96
state, obs = v_reset(H.ppo_keys(1, n))
state[1][:8, 8] = 999.0                                   # force truncation -> fused reset path
g = torch.Generator(device="cuda").manual_seed(0)
for t in range(3):
    act = torch.randn(n, nu, device="cuda", generator=g)
    state, obs, r, te, tr = v_step.autoreset(state, act, H.ppo_keys(2 + t, n))
q, v, w, c = H.make_states(model, n, 77, "tumble")       # includes envs that overflow the 32-row tile
T = lambda a: torch.tensor(a, dtype=torch.float32, device="cuda")
nd, out = mjx.step(sysm, mjx.Data(T(q), T(v), T(w), torch.zeros(n, device="cuda"), T(c)), debug=True)
pos = mjx.speed_test(sysm, torch.linspace(0, 1, n, device="cuda"), 2)
torch.cuda.synchronize()
print("sanitize case ok", float(obs.sum()), int((out["status"] & 2 != 0).sum()), float(pos.sum()))
