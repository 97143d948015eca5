"""Timeline of CTA 0 of the fused policy kernel (profiling build libmjxb_trace.so, -DMJXB_POLICY_TRACE=1): MJXB_LIB=.../libmjxb_trace.so"""
import sys, os, ctypes as C
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from mujoco_mjx_lab_b200 import policy as PL, ppo as P, _lib
od, nu = 54, 21
n = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
g = torch.Generator(device="cuda").manual_seed(0)
params = [p.detach() for p in P._mlp_params(od, [(256, "tanh")] * 3, nu, g, "cuda")]
log_std = torch.zeros(nu, device="cuda")
fp = PL.FusedPolicy(params, log_std, od, nu)
obs = [torch.randn(n, od, device="cuda", generator=g) for _ in range(3)]
eps = [torch.randn(n, nu, device="cuda", generator=g) for _ in range(3)]
rm, rv = torch.zeros(od, device="cuda"), torch.ones(od, device="cuda")
L = _lib.lib()
buf = (C.c_ulonglong * 4096)()
for i in range(3):
    fp.act(obs[i], eps[i], rm, rv)
    L.mjxb_policy_trace(buf)
a = np.frombuffer(buf, dtype=np.uint64)
ev = []
for region in range(3):
    for i in range(640):
        tag, t = int(a[region * 1280 + 2 * i]), int(a[region * 1280 + 2 * i + 1])
        if t == 0:
            break
        ev.append((t / 1965.0 * 1e3, tag >> 32, tag & 0xffffffff))     # SM clock at 1.965 GHz -> ns
ev.sort()
t0 = ev[0][0]
names = {202: "epi loads issued", 203: "epi first vec done", 204: "epi scatter done", 100: "mma wait A", 110: "mma got A", 120: "mma issued", 200: "epi tile start", 201: "epi A0 ready", 210: "epi acc ready", 220: "epi done"}
for t, s, tag in ev[:int(sys.argv[2]) if len(sys.argv) > 2 else 120]:
    base = tag if 202 <= tag <= 204 else (tag - tag % 10 if tag < 223 else 220)
    print(f"{(t - t0) / 1e3:9.2f} us  slot {s}  {names.get(base, '?'):16s} L{tag % 10}")
