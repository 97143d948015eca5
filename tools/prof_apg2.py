"""Kernel-level breakdown of one APG update (2048 envs x 128 steps, CG 4/4): total GPU kernel time by kernel, launch counts."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from torch.profiler import profile, ProfilerActivity
from mujoco_mjx_lab_b200 import apg
from mujoco_mjx_lab_b200.config import APGConfig
cfg, env = apg.make_apg_env()
acfg = APGConfig(); acfg.horizon = 128
tr = apg.APGTrainer(acfg, env[8], env[9], 2048, output_scale=0.01, use_cuda_graph=False)
tr.update(); tr.update()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    out = tr.update()
    torch.cuda.synchronize()
ka = prof.key_averages()
tot = sum(k.self_device_time_total for k in ka)
n = sum(k.count for k in ka)
print(f"update_ms (eager) {out['update_ms']:.1f}  sum of kernel time {tot/1e3:.1f} ms over {n} launches")
for k in sorted(ka, key=lambda k: -k.self_device_time_total)[:28]:
    print(f"{k.self_device_time_total/1e3:8.2f} ms  {k.count:6d} x {k.self_device_time_total/max(k.count,1):7.1f} us  {k.key[:110]}")
