import os, sys
import numpy as np, torch
sys.path.insert(0, "/root/repo/tests"); sys.path.insert(0, "/root/repo")
import helpers as H
from mujoco_mjx_lab_b200 import mjx, training_utils
model, cfg = H.load(), H.env_config()
m, sysm, q0, nq, nv, nu, single_reset, single_step, v_reset, v_step = training_utils.load_model_and_create_env("", cfg, model=model)
n = 65536
g = torch.Generator(device="cuda").manual_seed(0)
state, obs = v_reset(torch.from_numpy(H.ppo_keys(42, n).view(np.int32)).cuda())
def act(): return torch.randn(n, nu, device="cuda", generator=g).clamp_(-1, 1)
def rk(): return torch.randint(-2**31, 2**31 - 1, (n, 2), device="cuda", dtype=torch.int32, generator=g)
for i in range(60): state, obs, r, te, tr = v_step.autoreset(state, act(), rk(), inplace=True)
its = []
for i in range(3):
    a = act()
    d = state[0]
    _, out = mjx.forward(v_step.sys, mjx.Data(d.qpos, d.qvel, d.qacc_warmstart, d.time, a), debug=True)
    its.append(out["solver_niter"].clone())
    state, obs, r, te, tr = v_step.autoreset(state, a, rk(), inplace=True)
a, b = its[0].float(), its[1].float()
print("corr", float(torch.corrcoef(torch.stack([a, b]))[0, 1]))
d = (its[1] - its[0]).abs()
print("P(|d|<=0,1,2):", float((d == 0).float().mean()), float((d <= 1).float().mean()), float((d <= 2).float().mean()))
# expected max-of-16 in random vs sorted-by-previous grouping
def emax(x, grp=16):
    x = x[: (len(x) // grp) * grp].view(-1, grp)
    return float(x.max(1).values.float().mean())
idx = torch.argsort(its[0])
print("mean", float(b.mean()), "E[max16] random", emax(its[1]), "sorted by prev", emax(its[1][idx]), "oracle-sorted", emax(torch.sort(its[1]).values))
