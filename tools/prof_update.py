import sys, os
sys.path.insert(0, "/root/repo")
import torch
from torch.profiler import profile, ProfilerActivity
from mujoco_mjx_lab_b200 import ppo as P, modelc, training_utils
from mujoco_mjx_lab_b200.config import PPOConfig, EnvConfig
cfg = PPOConfig(); cfg.rollout_length, cfg.minibatch_size, cfg.epochs = 64, 65536, 1
cfg.env_config = EnvConfig(posture_penalty_weight=0.0, random_flip=True)
model = modelc.builtin_model("humanoid_mjx")
env = training_utils.load_model_and_create_env("", cfg.env_config, model=model)
torch.backends.cuda.matmul.allow_tf32 = True
tr = P.PPOTrainer(cfg, env[8], env[9], 1024, use_cuda_graph=False)
tr.iteration(); tr.iteration()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    tr.iteration()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=70))
