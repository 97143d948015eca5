"""Per-stage latency of one env step, from clock() stamps written by a profiling build of the kernel.
Build the variant:  cd mujoco_mjx_lab_b200/csrc && nvcc <Makefile flags> -DMJXB_STAGE_CLOCK=1 -shared -o ../../variants/libmjxb_clock.so mjxb_abi.cu mjxb_policy.cu -lcudart
Run:                MJXB_LIB=variants/libmjxb_clock.so python tools/stage_clock.py [n_env]"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, torch
from mujoco_mjx_lab_b200 import _lib, modelc, training_utils, parallel, mjx
from mujoco_mjx_lab_b200.config import EnvConfig

n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
env = training_utils.load_model_and_create_env("", EnvConfig(posture_penalty_weight=0.0, random_flip=True), model=modelc.builtin_model("humanoid_mjx"))
v_reset, v_step = env[8], env[9]
g = torch.Generator(device="cuda").manual_seed(1)
acts = [torch.randn(n, 21, device="cuda", generator=g).clamp_(-1, 1) for _ in range(4)]
keys = [torch.randint(-2 ** 31, 2 ** 31 - 1, (n, 2), device="cuda", dtype=torch.int32, generator=g) for _ in range(4)]
state, obs = v_reset(torch.from_numpy(parallel.rank_keys(42, 0, n).view(np.int32)).cuda())
for i in range(80):
    state, obs, r, te, tr = v_step.autoreset(state, acts[i % 4], keys[i % 4], inplace=True)
torch.cuda.synchronize()
d = state[0]
_, dbg = mjx.forward(v_step.sys, mjx.Data(d.qpos.clone(), d.qvel.clone(), d.qacc_warmstart.clone(), d.time.clone(), acts[0]), debug=True)
niter = dbg["solver_niter"].cpu().numpy()
state, obs, r, te, tr = v_step.autoreset(state, acts[0], keys[0], inplace=True)
torch.cuda.synchronize()
m = min(n, 4096)
out = np.zeros((m, 32), dtype=np.int32)
L = _lib.lib()
L.mjxb_debug_stage_clock.argtypes = [C.c_void_p, C.c_int32]
assert L.mjxb_debug_stage_clock(out.ctypes.data, m) == 0
names = ["load+ctrl", "kinematics", "geoms/sites", "com/cinert/cdof", "com_vel/cacc/rne", "crb/M/bias", "collision", "constraint rows",
         "M^-1 + warm start", "newton loop", "sensors", "integrate", "env layer", "store"]
dt = np.diff(out[:, :14].astype(np.int64), axis=1)
tot = (out[:, 13] - out[:, 0]).astype(np.int64)
print(f"n={n}: total cycles/env-step mean {tot.mean():.0f} (max {tot.max()}), mean newton iters {niter[:m].mean():.2f}")
for i, nm in enumerate(names[1:]):
    print(f"  {nm:22s} mean {dt[:, i].mean():8.0f}  p95 {np.percentile(dt[:, i], 95):8.0f}  ({100 * dt[:, i].mean() / tot.mean():4.1f} %)")
print(f"  inside 'M^-1 + warm start': entry barrier + M factor/solve {np.median(out[:, 14] - out[:, 7]):.0f}, warm-start evaluation {np.median(out[:, 15] - out[:, 14]):.0f}, "
      f"first update_constraint {np.median(out[:, 8] - out[:, 15]):.0f} (medians)")
it = niter[:m]
for k in sorted(set(it.tolist())):
    sel = it == k
    print(f"  newton loop cycles at niter={k}: {dt[sel, 8].mean():8.0f}  (n={sel.sum()})")
inner = ["H build (M load + active rows)", "factor + forward subst.", "backward subst.", "M*search, J*search", "line search (breakpoints, alpha)",
         "apply step", "update_constraint + grad"]
di = np.diff(out[:, 16:24].astype(np.int64), axis=1)
ok = it >= 1
dense = ok & ((out[:, 18] - out[:, 17]) <= 0)
print(f"  envs on the dense-factorisation path (a candidate contact couples two limbs): {dense.sum()} of {ok.sum()} ({100.0 * dense.sum() / max(ok.sum(), 1):.2f} %)")
print(f"  whole first iteration, tree path {np.median((out[ok & ~dense, 23] - out[ok & ~dense, 16])):.0f} cycles, dense path {np.median((out[dense, 23] - out[dense, 16])) if dense.any() else float('nan'):.0f} cycles")
ok = ok & ~dense
print("  first Newton iteration, stage latencies (cycles):")
for i, nm in enumerate(inner):
    print(f"    {nm:34s} mean {di[ok, i].mean():7.0f}  p95 {np.percentile(di[ok, i], 95):7.0f}")
print(f"    {'whole iteration':34s} mean {(out[ok, 23] - out[ok, 16]).mean():7.0f}")
ls = ["qg1/qg2 sums", "breakpoint setup (div, store, sync)", "derivative sweep over rows", "bracket (redux min/max)", "quadratic piece loop",
      "3 sums + alpha"]
seq = np.stack([out[:, 20], out[:, 24], out[:, 25], out[:, 26], out[:, 27], out[:, 28], out[:, 21]], axis=1).astype(np.int64)
dl = np.diff(seq, axis=1)
print("  line search detail (first iteration, medians):")
for i, nm in enumerate(ls):
    print(f"    {nm:36s} median {np.median(dl[ok, i]):7.0f}")
for i, nm in enumerate(inner):
    print(f"    [median] {nm:34s} {np.median(di[ok, i]):7.0f}")
