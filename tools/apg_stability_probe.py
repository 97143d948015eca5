"""APG stability probe: 128 env steps of the CPU oracle (restated mjx.step) at train_apg.py:101-105 solver settings (CG 4/4) under a
random tanh policy with Xavier-scale weights on [qpos, qvel]: velocity feedback drives some envs to non-finite states within ~30 steps."""
import sys, numpy as np
import os; R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path[:0] = [R, os.path.join(R, 'tests')]
import helpers
from mujoco_mjx_lab_b200.config import EnvConfig
m = helpers.load(overrides=dict(solver=1, iterations=4, ls_iterations=4))
cfg = helpers.env_config(posture_penalty_weight=0.6, random_flip=False)
orc = helpers.make_oracle(m, cfg)
n = 512
st, obs = orc.env_reset(helpers.ppo_keys(1, n), prec="f32")
rng = np.random.default_rng(0)
W = rng.normal(size=(55, 21)) * 0.16
for t in range(128):
    o = np.concatenate([st["qpos"], st["qvel"]], 1)
    act = np.tanh(np.tanh(o @ W))
    st, obs, r, te, tr, mask, _ = orc.env_step(st, act, prec="f32")
    bad = ~np.isfinite(st["qvel"]).all(1)
    if t % 16 == 0 or bad.any():
        print(t, "nan envs", bad.sum(), "max|qvel|", np.nanmax(np.abs(st["qvel"])), "reward min", np.nanmin(r), "done", int(np.maximum(te,tr).sum()), "min z", st["qpos"][:,2].min())
    if bad.any(): break
