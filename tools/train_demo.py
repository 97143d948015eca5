"""End-to-end sanity run of the PPO driver on the fused step (reference train_ppo.py defaults: 1024 envs x 256 steps, 4 epochs):
prints the mean per-env rollout return and the mean episode length every few iterations. Usage: python tools/train_demo.py [iters]"""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from mujoco_mjx_lab_b200 import ppo as P, modelc, training_utils
from mujoco_mjx_lab_b200.config import PPOConfig, EnvConfig

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 200
cfg = PPOConfig()
cfg.rollout_length, cfg.minibatch_size, cfg.epochs = 256, 65536, 4
cfg.env_config = EnvConfig(posture_penalty_weight=0.0, random_flip=True)
env = training_utils.load_model_and_create_env("", cfg.env_config, model=modelc.builtin_model("humanoid_mjx"))
torch.backends.cuda.matmul.allow_tf32 = True
tr = P.PPOTrainer(cfg, env[8], env[9], 1024)
t0 = time.perf_counter()
for it in range(iters):
    out = tr.iteration()
    if it % max(1, iters // 20) == 0 or it == iters - 1:
        print(f"iter {it:4d}  return/rollout {out['train_return_avg']:9.2f}  mean episode length {out['train_eplen_avg']:7.1f}  "
              f"rollout {out['rollout_ms']:.1f} ms update {out['update_ms']:.1f} ms  wall {time.perf_counter() - t0:.1f} s", flush=True)
