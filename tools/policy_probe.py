"""Checks the fused tcgen05 policy kernel against torch (float32 and a bf16-operand emulation) and times it."""
import os, sys, math
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from mujoco_mjx_lab_b200 import policy as PL, ppo as P

torch.manual_seed(0)
dev = "cuda"
od, nu = 54, 21
g = torch.Generator(device=dev).manual_seed(3)
params = P._mlp_params(od, [(256, "tanh")] * 3, nu, g, dev)
for i in range(1, 8, 2):
    params[i].data.normal_(0, 0.1, generator=g)
log_std = torch.full((nu,), -0.3, device=dev)
fp = PL.FusedPolicy([p.detach() for p in params], log_std, od, nu)
for n in [int(a) for a in sys.argv[1:]] or [100, 1024, 65536]:
    obs = torch.randn(n, od, device=dev, generator=g) * 2 + 0.5
    eps = torch.randn(n, nu, device=dev, generator=g)
    mean_r, var_r = torch.randn(od, device=dev, generator=g) * 0.3, torch.rand(od, device=dev, generator=g) + 0.5
    mean_k = torch.empty(n, nu, device=dev)
    act, logp = fp.act(obs, eps, mean_r, var_r, mean_out=mean_k)
    torch.cuda.synchronize()
    assert int(fp.error) == 0, "tensor-core completion not observed"
    with torch.no_grad():
        x = torch.clamp((obs - mean_r) / torch.sqrt(var_r + 1e-8), -10, 10)
        ref = P._mlp_apply([p.detach() for p in params], x, 3)                       # float32 reference
        h = x.bfloat16().float()                                                      # bf16-operand emulation (fp32 accumulate)
        for i in range(0, 8, 2):
            h = h @ params[i].detach().bfloat16().float() + params[i + 1].detach()
            if i < 6:
                h = torch.tanh(h).bfloat16().float()
        emu = h
        act_ref = mean_k + torch.exp(log_std) * eps
        logp_ref = P.gaussian_logprob(mean_k, log_std, act_ref)
    print(f"n={n}: |mean-emu| max {float((mean_k-emu).abs().max()):.2e}  |mean-f32| max {float((mean_k-ref).abs().max()):.2e} "
          f"(ref scale {float(ref.abs().mean()):.2f})  act {float((act-act_ref).abs().max()):.1e}  logp {float((logp-logp_ref).abs().max()):.1e}")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(3):
        fp.act(obs, eps, mean_r, var_r)
    e0.record()
    for _ in range(20):
        fp.act(obs, eps, mean_r, var_r)
    e1.record(); torch.cuda.synchronize()
    print(f"      fused kernel {e0.elapsed_time(e1) / 20 * 1e3:.1f} us per call (incl. launch)")
