"""Soak run: many fused steps with auto-reset on a large batch; reports NaN / status flags / episode statistics."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import ctypes as C  # noqa: E402

import helpers as H  # noqa: E402
from mujoco_mjx_lab_b200 import _lib, training_utils  # noqa: E402
from mujoco_mjx_lab_b200.mjx import _stream, state_c  # noqa: E402

n, steps = int(sys.argv[1]) if len(sys.argv) > 1 else 65536, int(sys.argv[2]) if len(sys.argv) > 2 else 1500
model, cfg = H.load(), H.env_config()
m, sysm, q0, nq, nv, nu, single_reset, single_step, v_reset, v_step = training_utils.load_model_and_create_env("", cfg, model=model)
L, h = _lib.lib(), v_step.sys.handle
g = torch.Generator(device="cuda").manual_seed(0)
(d, aux), obs = v_reset(torch.from_numpy(H.ppo_keys(42, n).view(np.int32)).cuda())
status = torch.zeros(n, dtype=torch.int32, device="cuda")
rew, te, tr = (torch.empty(n, device="cuda") for _ in range(3))
mask = torch.zeros(n, dtype=torch.uint8, device="cuda")
acc = dict(nan=0, spill=0, maxiter=0, resets=0, term=0, trunc=0)
for t in range(steps):
    scale = 1.0 if t % 3 else 3.0                       # occasionally saturating actions
    act = (torch.randn(n, nu, device="cuda", generator=g) * scale).contiguous()
    keys = torch.randint(-2**31, 2**31 - 1, (n, 2), device="cuda", dtype=torch.int32, generator=g)
    st = state_c(d.qpos, d.qvel, d.qacc_warmstart, d.time, aux)
    _lib.check(L.mjxb_step_autoreset(h, n, st, act.data_ptr(), keys.data_ptr(), st, obs.data_ptr(), rew.data_ptr(), te.data_ptr(), tr.data_ptr(),
                                     mask.data_ptr(), status.data_ptr(), _stream()))
    acc["nan"] += int((status & 1).ne(0).sum()); acc["spill"] += int((status & 2).ne(0).sum()); acc["maxiter"] += int((status & 4).ne(0).sum())
    acc["resets"] += int(mask.sum()); acc["term"] += int(te.sum()); acc["trunc"] += int(tr.sum())
    if not (torch.isfinite(obs).all() and torch.isfinite(rew).all() and torch.isfinite(d.qvel).all()):
        print("NON-FINITE at step", t); break
print("env-steps", n * steps, acc, "mean eplen", n * steps / max(acc["resets"], 1), "max |qvel|", float(d.qvel.abs().max()))
