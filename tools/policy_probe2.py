"""Fused policy kernel: timing at several batch sizes (CUDA events, L2-flushing rotation of inputs) and a quick check against torch."""
import sys, os
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from mujoco_mjx_lab_b200 import policy as PL, ppo as P
od, nu = 54, 21
g = torch.Generator(device="cuda").manual_seed(0)
params = [p.detach() for p in P._mlp_params(od, [(256, "tanh")] * 3, nu, g, "cuda")]
for i in range(1, 8, 2): params[i].normal_(0, 0.1, generator=g)
log_std = torch.linspace(-0.5, 0.2, nu, device="cuda")
fp = PL.FusedPolicy(params, log_std, od, nu)
rm, rv = torch.randn(od, device="cuda", generator=g) * 0.3, torch.rand(od, device="cuda", generator=g) + 0.5
for n in [int(x) for x in (sys.argv[1].split(",") if len(sys.argv) > 1 else "128,1024,4096,65536,262144".split(","))]:
    nbuf = max(2, min(8, (256 << 20) // (n * od * 4)))
    obs = [torch.randn(n, od, device="cuda", generator=g) * 2 + 0.5 for _ in range(nbuf)]
    eps = [torch.randn(n, nu, device="cuda", generator=g) for _ in range(nbuf)]
    act, logp, mean = torch.empty(n, nu, device="cuda"), torch.empty(n, device="cuda"), torch.empty(n, nu, device="cuda")
    fp.act(obs[0], eps[0], rm, rv, act_out=act, logp_out=logp, mean_out=mean)
    torch.cuda.synchronize()
    x = torch.clamp((obs[0] - rm) / torch.sqrt(rv + 1e-8), -10, 10)
    ref = P._mlp_apply(params, x, 3)
    err = float((mean - ref).abs().max())
    for _ in range(5): fp.act(obs[1], eps[1], rm, rv, act_out=act, logp_out=logp)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 40
    torch.cuda.synchronize(); e0.record()
    for i in range(reps): fp.act(obs[i % nbuf], eps[i % nbuf], rm, rv, act_out=act, logp_out=logp)
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / reps * 1e3
    if n <= 8192:                                   # small batches: the eager loop is host-bound; replay 20 launches from a CUDA graph
        st = torch.cuda.Stream()
        with torch.cuda.stream(st):
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr, stream=st):
                for i in range(20): fp.act(obs[i % nbuf], eps[i % nbuf], rm, rv, act_out=act, logp_out=logp)
            gr.replay(); torch.cuda.synchronize()
            e0.record(st)
            for _ in range(10): gr.replay()
            e1.record(st); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / 200 * 1e3
    fl = n * 2 * (64 * 256 + 2 * 256 * 256 + 256 * 32)
    print(f"n={n:7d}  {us:8.2f} us  {fl / us / 1e6:8.1f} TFLOP/s (padded)  max|mean - f32 torch|={err:.4f} error_flag={int(fp.error)}")
