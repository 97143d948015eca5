"""APG stability probe: 128 env steps of the CPU oracle (restated mjx.step, float32) with ZERO actions, at train_apg.py:101-105 solver settings (CG, 4
iterations, 4 line-search iterations) and at the Newton 10/20 defaults: with CG 4/4 a few of 2048 fallen humanoids diverge to non-finite states
by step ~116; with Newton none do.  python tools/apg_stability_probe.py [cg|newton]"""
import sys, numpy as np
import os; R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path[:0] = [R, os.path.join(R, 'tests')]
import helpers
solver = sys.argv[1] if len(sys.argv) > 1 else "cg"
over = dict(solver=1, iterations=4, ls_iterations=4) if solver == "cg" else {}
m = helpers.load(overrides=over)
cfg = helpers.env_config(posture_penalty_weight=0.6, random_flip=False)
orc = helpers.make_oracle(m, cfg)
n = 2048
st, obs = orc.env_reset(helpers.ppo_keys(1, n), prec="f32")
for t in range(128):
    act = np.zeros((n, 21))
    st, obs, r, te, tr, mask, _ = orc.env_step(st, act, prec="f32")
    bad = ~np.isfinite(st["qvel"]).all(1)
    if t % 16 == 15 or bad.any():
        print(t, "nan envs", bad.sum(), "max|qvel|", np.nanmax(np.abs(st["qvel"])), "reward min", np.nanmin(r), "done", int(np.maximum(te,tr).sum()), "min z", np.nanmin(st["qpos"][:,2]), flush=True)
    if bad.sum() > 20: break
