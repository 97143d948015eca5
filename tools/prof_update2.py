"""Kernel-level breakdown of ONE 65,536-sample PPO minibatch step (gather + both networks forward / backward + Adam), eager launches."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from torch.profiler import profile, ProfilerActivity
from mujoco_mjx_lab_b200 import ppo as P, modelc, training_utils
from mujoco_mjx_lab_b200.config import PPOConfig, EnvConfig
cfg = PPOConfig(); cfg.rollout_length, cfg.minibatch_size, cfg.epochs = 64, 65536, 1
cfg.env_config = EnvConfig(posture_penalty_weight=0.0, random_flip=True)
model = modelc.builtin_model("humanoid_mjx")
env = training_utils.load_model_and_create_env("", cfg.env_config, model=model)
torch.backends.cuda.matmul.allow_tf32 = True
tr = P.PPOTrainer(cfg, env[8], env[9], 4096, use_cuda_graph=False)
tr.iteration()
total, od = 64 * 4096, 54
obs_f = torch.randn(total, tr.kpad, device="cuda"); obs_f[:, od:] = 0; act_f = torch.randn(total, 21, device="cuda")
logp_f = torch.randn(total, device="cuda"); ret_f = torch.randn(total, device="cuda"); adv_f = torch.randn(total, device="cuda")
idx = torch.randperm(total, device="cuda")[:65536]
for _ in range(3):
    tr._minibatch_fb(obs_f, act_f, logp_f, ret_f, adv_f, idx, zero=True); tr._opt_step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(4):
        tr._minibatch_fb(obs_f, act_f, logp_f, ret_f, adv_f, idx, zero=True); tr._opt_step()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=40, max_name_column_width=90))
