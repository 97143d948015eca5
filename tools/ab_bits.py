"""Bit-level A/B of two library builds: runs the same seeded rollout (reset + K auto-reset steps) and writes a digest of every output.
Usage: MJXB_LIB=path/to/lib.so python tools/ab_bits.py out.npz [n_env] [steps];  python tools/ab_bits.py --cmp a.npz b.npz"""
import os, sys
import numpy as np
if sys.argv[1] == "--cmp":
    a, b = np.load(sys.argv[2]), np.load(sys.argv[3])
    bad = 0
    for k in a.files:
        same = np.array_equal(a[k].view(np.uint32), b[k].view(np.uint32))
        if not same:
            d = np.abs(a[k].astype(np.float64) - b[k].astype(np.float64))
            print(f"{k}: DIFFERENT  mismatched {np.mean(a[k] != b[k]):.3e}  max abs diff {np.nanmax(d):.3e}")
            bad += 1
    print("bit-identical" if bad == 0 else f"{bad} arrays differ")
    sys.exit(0)
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from mujoco_mjx_lab_b200 import modelc, training_utils, parallel
from mujoco_mjx_lab_b200.config import EnvConfig
out = sys.argv[1]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 40
model = modelc.builtin_model("humanoid_mjx")
env = training_utils.load_model_and_create_env("", EnvConfig(posture_penalty_weight=0.0, random_flip=True), model=model)
v_reset, v_step = env[8], env[9]
g = torch.Generator(device="cuda").manual_seed(7)
acts = [torch.randn(n, 21, device="cuda", generator=g).clamp_(-1, 1) for _ in range(8)]
keys = [torch.randint(-2 ** 31, 2 ** 31 - 1, (n, 2), device="cuda", dtype=torch.int32, generator=g) for _ in range(8)]
state, obs = v_reset(torch.from_numpy(parallel.rank_keys(42, 0, n).view(np.int32)).cuda())
rs = torch.zeros(n, device="cuda")
for i in range(steps):
    state, obs, r, te, tr = v_step.autoreset(state, acts[i % 8], keys[i % 8], inplace=True)
    rs += r
d, aux = state
np.savez(out, qpos=d.qpos.cpu().numpy(), qvel=d.qvel.cpu().numpy(), warm=d.qacc_warmstart.cpu().numpy(), time=d.time.cpu().numpy(),
         aux=aux.cpu().numpy(), obs=obs.cpu().numpy(), rsum=rs.cpu().numpy(), te=te.cpu().numpy(), tr=tr.cpu().numpy())
print("wrote", out)
