"""How far apart do the graph-replayed and the eager PPO learner drift over 3 iterations (tests/test_gpu_ppo.py::test_graph_update_equals_eager)?
Nondeterministic reduction orders (atomics in the bias-gradient kernel, cuBLAS stream-K) are amplified by Adam's normalisation."""
import sys, os
sys.path[:0] = [os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."), os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests")]
import numpy as np, torch
import helpers
from mujoco_mjx_lab_b200 import ppo as ppo_mod, training_utils
from mujoco_mjx_lab_b200.config import PPOConfig
model = helpers.load()
env = training_utils.load_model_and_create_env("", helpers.env_config(), model=model)
def trainer(graph):
    cfg = PPOConfig(); cfg.rollout_length, cfg.minibatch_size, cfg.epochs = 16, 512, 4
    cfg.env_config = helpers.env_config()
    return ppo_mod.PPOTrainer(cfg, env[8], env[9], 128, seed=5, use_cuda_graph=graph)
worst = []
for rep in range(int(sys.argv[1]) if len(sys.argv) > 1 else 12):
    a, b = trainer(True), trainer(False)
    def copy_rollout():
        for name in ("obs_traj", "act_traj", "logp_traj", "r_traj", "term_traj", "trunc_traj", "obs"):
            getattr(b, name).copy_(getattr(a, name))
    b.collect_rollout = copy_rollout
    for _ in range(3):
        a.iteration(); b.iteration()
    m = 0.0
    for pa, pb in zip(a.policy + [a.log_std] + a.value, b.policy + [b.log_std] + b.value):
        d = (pa - pb).abs()
        m = max(m, float((d / (2e-4 + 2e-3 * pb.abs())).max()))
    worst.append(m)
    print(f"rep {rep}: worst |a-b| / (2e-4 + 2e-3 |b|) = {m:.3f}", flush=True)
print("max over reps", max(worst))
