"""Times mjxb_step_autoreset_host (pinned buffers) against the device-resident step at one batch size; prints per-call ms.
Usage: python tools/e2e_probe.py [n_env] [steps]   (pipeline variants are selected with the MJXB_HOST_* / MJXB_DIRECT_* env vars)"""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, torch
from mujoco_mjx_lab_b200 import _lib, modelc, training_utils, parallel
from mujoco_mjx_lab_b200.config import EnvConfig

n = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
dev = torch.device("cuda:0")
model = modelc.builtin_model("humanoid_mjx")
env = training_utils.load_model_and_create_env("", EnvConfig(posture_penalty_weight=0.0, random_flip=True), model=model)
v_reset, v_step = env[8], env[9]
g = torch.Generator(device=dev).manual_seed(1234)
acts = [torch.randn(n, 21, device=dev, generator=g).clamp_(-1, 1) for _ in range(4)]
keys = [torch.randint(-2 ** 31, 2 ** 31 - 1, (n, 2), device=dev, dtype=torch.int32, generator=g) for _ in range(4)]
state, obs = v_reset(torch.from_numpy(parallel.rank_keys(42, 0, n).view(np.int32)).to(dev))
for i in range(60):
    state, obs, r, te, tr = v_step.autoreset(state, acts[i % 4], keys[i % 4], inplace=True)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(steps):
    state, obs, r, te, tr = v_step.autoreset(state, acts[i % 4], keys[i % 4], inplace=True)
e1.record(); torch.cuda.synchronize()
dev_ms = e0.elapsed_time(e1) / steps
L, h = _lib.lib(), v_step.sys.handle
pin = lambda *s, dt=torch.float32: torch.empty(*s, dtype=dt, pin_memory=True)
h_act = [pin(n, 21) for _ in range(2)]
for b in h_act:
    b.copy_(acts[0].cpu())
h_keys = pin(n, 2, dt=torch.int32); h_keys.copy_(keys[0].cpu())
h_obs, h_r, h_te, h_tr = pin(n, 54), pin(n), pin(n), pin(n)
d, aux = state
hs = [np.ascontiguousarray(t.detach().cpu().numpy()) for t in (d.qpos, d.qvel, d.qacc_warmstart, d.time, aux)]
_lib.check(L.mjxb_state_set_host(h, n, *[a.ctypes.data for a in hs]))
def step(i):
    _lib.check(L.mjxb_step_autoreset_host(h, n, h_act[i % 2].data_ptr(), h_keys.data_ptr(), h_obs.data_ptr(), h_r.data_ptr(), h_te.data_ptr(), h_tr.data_ptr()))
for i in range(3):
    step(i)
t0 = time.perf_counter()
for i in range(steps):
    step(i)
ms = (time.perf_counter() - t0) * 1e3 / steps
tag = " ".join(f"{k}={v}" for k, v in os.environ.items() if k.startswith("MJXB_"))
print(f"n={n} device {dev_ms:.3f} ms ({n/dev_ms/1e3:.2f} M/s)  host-api {ms:.3f} ms ({n/ms/1e3:.2f} M/s)  overhead {ms-dev_ms:.3f} ms  [{tag}]")
