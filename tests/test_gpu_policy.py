"""Fused tcgen05 policy-inference kernel (include/mjxb.h mjxb_policy_*; reference train_ppo.py:121-126,135-140, src/networks.py:55-61)
against plain PyTorch: float32 reference of the same op (tolerance of bf16 operands, stated below) and a bf16-operand emulation."""
import pytest
import torch

from mujoco_mjx_lab_b200 import policy as PL, ppo as P

pytestmark = pytest.mark.gpu


def _setup(seed, od=54, nu=21):
    g = torch.Generator(device="cuda").manual_seed(seed)
    params = [p.detach() for p in P._mlp_params(od, [(256, "tanh")] * 3, nu, g, "cuda")]
    for i in range(1, 8, 2):
        params[i].normal_(0, 0.1, generator=g)
    log_std = torch.linspace(-0.5, 0.2, nu, device="cuda")
    return g, params, log_std


@pytest.mark.parametrize("n", [1, 100, 128, 129, 1024, 20000])
def test_fused_policy_matches_torch(n):
    od, nu = 54, 21
    g, params, log_std = _setup(n)
    fp = PL.FusedPolicy(params, log_std, od, nu)
    obs = torch.randn(n, od, device="cuda", generator=g) * 2 + 0.5
    eps = torch.randn(n, nu, device="cuda", generator=g)
    rm, rv = torch.randn(od, device="cuda", generator=g) * 0.3, torch.rand(od, device="cuda", generator=g) + 0.5
    mean_k = torch.full((n, nu), float("nan"), device="cuda")
    act, logp = fp.act(obs, eps, rm, rv, mean_out=mean_k)
    torch.cuda.synchronize()
    assert int(fp.error) == 0
    x = torch.clamp((obs - rm) / torch.sqrt(rv + 1e-8), -10, 10)
    ref = P._mlp_apply(params, x, 3)                                     # plain float32 PyTorch reference of the same op
    h = x.bfloat16().float()                                             # the kernel's arithmetic: bf16 operands, fp32 accumulation
    for i in range(0, 8, 2):
        h = h @ params[i].bfloat16().float() + params[i + 1]
        if i < 6:
            h = torch.tanh(h).bfloat16().float()
    # bf16 operands (8-bit mantissa) through 3 hidden layers: 3e-2 absolute on O(0.5) outputs against float32, 1e-2 against the
    # emulation (tanh.approx, accumulation order, bf16 rounding flips)
    torch.testing.assert_close(mean_k, ref, rtol=0, atol=3e-2)
    torch.testing.assert_close(mean_k, h, rtol=0, atol=1e-2)
    torch.testing.assert_close(act, mean_k + torch.exp(log_std) * eps, rtol=1e-6, atol=1e-6)       # sampling: exact in float32
    torch.testing.assert_close(logp, P.gaussian_logprob(mean_k, log_std, act), rtol=1e-5, atol=2e-5)


def test_fused_policy_without_normalisation_and_repacking():
    od, nu, n = 54, 21, 300
    g, params, log_std = _setup(7)
    fp = PL.FusedPolicy(params, log_std, od, nu)
    obs = torch.randn(n, od, device="cuda", generator=g)
    eps = torch.zeros(n, nu, device="cuda")
    m1 = torch.empty(n, nu, device="cuda")
    fp.act(obs, eps, None, None, mean_out=m1)
    torch.testing.assert_close(m1, P._mlp_apply(params, obs, 3), rtol=0, atol=3e-2)
    params[0].mul_(0.5)                                                  # an optimiser step happened: pack() picks the new weights up
    fp.pack()
    m2 = torch.empty(n, nu, device="cuda")
    a2, lp2 = fp.act(obs, eps, None, None, mean_out=m2)
    torch.testing.assert_close(m2, P._mlp_apply(params, obs, 3), rtol=0, atol=3e-2)
    torch.testing.assert_close(a2, m2)                                   # eps = 0: the action is the mean
    assert not torch.allclose(m1, m2)


def test_fused_policy_unaligned_slices_and_many_tiles():
    """The kernel moves a tile's observations / noise / actions as single bulk copies when the blocks are 16-byte aligned and sized, and
    falls back to thread-per-row access otherwise: batch slices that start at an odd row (8-byte aligned only) and a batch of more than
    2 x 148 tiles (every CTA walks several tiles in both of its slots) must give the same numbers as the aligned call."""
    od, nu = 54, 21
    g, params, log_std = _setup(11)
    fp = PL.FusedPolicy(params, log_std, od, nu)
    n = 148 * 128 * 2 + 128 * 37 + 5
    obs_big = torch.randn(n + 1, od, device="cuda", generator=g)
    eps_big = torch.randn(n + 1, nu, device="cuda", generator=g)
    act_big = torch.empty(n + 1, nu, device="cuda")
    rm, rv = torch.zeros(od, device="cuda"), torch.ones(od, device="cuda")
    a0, lp0 = fp.act(obs_big[:n].contiguous(), eps_big[:n].contiguous(), rm, rv)
    x = torch.clamp((obs_big[:n] - rm) / torch.sqrt(rv + 1e-8), -10, 10)
    torch.testing.assert_close(a0, P._mlp_apply(params, x, 3) + torch.exp(log_std) * eps_big[:n], rtol=0, atol=3e-2)
    a1, lp1 = fp.act(obs_big[1:], eps_big[1:], rm, rv, act_out=act_big[1:])       # all three blocks start 8 bytes off a 16-byte boundary
    a2, lp2 = fp.act(obs_big[1:].contiguous(), eps_big[1:].contiguous(), rm, rv)
    torch.cuda.synchronize()
    assert int(fp.error) == 0
    assert torch.equal(a1, a2) and torch.equal(lp1, lp2)


def test_unsupported_shape_is_refused():
    g = torch.Generator(device="cuda").manual_seed(0)
    params = [p.detach() for p in P._mlp_params(54, [(128, "tanh")] * 3, 21, g, "cuda")]
    assert not PL.supported(params, 54, 21)
    with pytest.raises(ValueError):
        PL.FusedPolicy(params, torch.zeros(21, device="cuda"), 54, 21)
