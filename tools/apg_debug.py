"""Debug driver: APG updates at 2048 x 128 with per-update diagnostics."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from mujoco_mjx_lab_b200 import apg, parallel  # noqa: E402

cfg, env = apg.make_apg_env()
cfg.horizon, cfg.hidden_size = 128, 32
tr = apg.APGTrainer(cfg, env[8], env[9], 2048, output_scale=0.01)
for it in range(4):
    keys = torch.from_numpy(parallel.rank_keys(1 + it, 0, tr.n).view(np.int32)).cuda()
    tr.opt.zero_grad(set_to_none=True)
    ret, obs_traj, mr = tr.rollout_return(keys, False)
    bad_obs = (~torch.isfinite(obs_traj)).any(-1).sum(1)           # per step: envs with non-finite state
    print(f"update {it}: return {float(ret):.4f} first step with non-finite state {int((bad_obs > 0).float().argmax()) if (bad_obs > 0).any() else -1}"
          f" envs non-finite at end {int(bad_obs[-1])}", flush=True)
    (-ret).backward()
    g = torch.cat([p.grad.reshape(-1) for p in tr.params])
    print(f"   grad finite {bool(torch.isfinite(g).all())} norm {float(g.norm()):.4g} max {float(g.abs().max()):.4g}", flush=True)
    torch.nn.utils.clip_grad_norm_(tr.params, 0.3)
    tr.opt.step()
    print(f"   params finite {all(bool(torch.isfinite(p).all()) for p in tr.params)}", flush=True)
