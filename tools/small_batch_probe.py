"""Per-step device time of the fused step at small batch sizes, free of host dispatch: K steps captured in one CUDA graph.
Also times the PPO rollout body pieces (policy MLP + sampling + log-prob + trajectory stores) the same way."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, torch
from mujoco_mjx_lab_b200 import modelc, training_utils, parallel, ppo as P
from mujoco_mjx_lab_b200.config import EnvConfig, PPOConfig

model = modelc.builtin_model("humanoid_mjx")
env = training_utils.load_model_and_create_env("", EnvConfig(posture_penalty_weight=0.0, random_flip=True), model=model)
v_reset, v_step = env[8], env[9]
K = 64
for n in [int(a) for a in sys.argv[1:]] or [64, 1024, 4096]:
    g = torch.Generator(device="cuda").manual_seed(1)
    acts = [torch.randn(n, 21, device="cuda", generator=g).clamp_(-1, 1) for _ in range(4)]
    keys = [torch.randint(-2 ** 31, 2 ** 31 - 1, (n, 2), device="cuda", dtype=torch.int32, generator=g) for _ in range(4)]
    state, obs = v_reset(torch.from_numpy(parallel.rank_keys(42, 0, n).view(np.int32)).cuda())
    for i in range(60):
        state, obs, r, te, tr = v_step.autoreset(state, acts[i % 4], keys[i % 4], inplace=True)
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        for i in range(K):
            v_step.autoreset(state, acts[i % 4], keys[i % 4], inplace=True)
    gr.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); gr.replay(); gr.replay(); e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / (2 * K)
    # eager (host-dispatched) for comparison
    e0.record()
    for i in range(K):
        v_step.autoreset(state, acts[i % 4], keys[i % 4], inplace=True)
    e1.record(); torch.cuda.synchronize()
    us_e = e0.elapsed_time(e1) * 1e3 / K
    # policy part of the rollout body
    cfg = PPOConfig(); cfg.rollout_length = K
    torch.backends.cuda.matmul.allow_tf32 = True
    tr_ = P.PPOTrainer(cfg, v_reset, v_step, n, use_cuda_graph=False)
    def policy_part():
        for t in range(K):
            obs_n = tr_.rms.normalize(tr_.obs)
            mean = P._mlp_apply(tr_.policy, obs_n, tr_.nh_p)
            eps = torch.randn(mean.shape, device="cuda")
            act = mean + torch.exp(tr_.log_std) * eps
            k2 = torch.randint(-2 ** 31, 2 ** 31 - 1, (n, 2), device="cuda", dtype=torch.int32)
            tr_.obs_traj[t].copy_(tr_.obs); tr_.act_traj[t].copy_(act)
            tr_.logp_traj[t].copy_(P.gaussian_logprob(mean, tr_.log_std, act))
            tr_.r_traj[t].copy_(tr_.logp_traj[t]); tr_.term_traj[t].copy_(tr_.logp_traj[t]); tr_.trunc_traj[t].copy_(tr_.logp_traj[t])
    with torch.no_grad():
        policy_part(); torch.cuda.synchronize()
        g2 = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g2):
            policy_part()
        g2.replay(); torch.cuda.synchronize()
        e0.record(); g2.replay(); g2.replay(); e1.record(); torch.cuda.synchronize()
    us_p = e0.elapsed_time(e1) * 1e3 / (2 * K)
    print(f"n={n:6d}  step in graph {us:7.1f} us ({n/us:.2f} M/s)   eager {us_e:7.1f} us   policy+sampling+stores in graph {us_p:6.1f} us")
