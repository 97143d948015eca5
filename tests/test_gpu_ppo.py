"""PPO iteration driver (caller of the hot path; reference train_ppo.py:128-371): the CUDA-graph update / GAE equal the eager ones."""
import numpy as np
import pytest
import torch

import helpers
from mujoco_mjx_lab_b200 import ppo as ppo_mod
from mujoco_mjx_lab_b200.config import PPOConfig

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def env(model):
    from mujoco_mjx_lab_b200 import training_utils
    return training_utils.load_model_and_create_env("", helpers.env_config(), model=model)


def _trainer(env, graph):
    cfg = PPOConfig()
    cfg.rollout_length, cfg.minibatch_size, cfg.epochs = 16, 512, 4
    cfg.env_config = helpers.env_config()
    return ppo_mod.PPOTrainer(cfg, env[8], env[9], 128, seed=5, use_cuda_graph=graph)


def test_gae_matches_numpy(env):
    tr = _trainer(env, False)
    T, n = 16, 128
    g = torch.Generator(device="cuda").manual_seed(0)
    r, v = torch.randn(T, n, device="cuda", generator=g), torch.randn(T + 1, n, device="cuda", generator=g)
    te = (torch.rand(T, n, device="cuda", generator=g) < 0.1).float()
    trn = (torch.rand(T, n, device="cuda", generator=g) < 0.1).float()
    adv, ret = tr.compute_gae(r, v, te, trn)                        # mjxb_gae kernel
    adv_t, ret_t = tr.compute_gae(r, v, te, trn, force_torch=True)  # the torch scan it replaces
    torch.testing.assert_close(adv, adv_t, rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(ret, ret_t, rtol=1e-5, atol=1e-5)
    rn, vn, ten, trnn = (x.double().cpu().numpy() for x in (r, v, te, trn))
    gam, lam = tr.cfg.gamma, tr.cfg.lam
    ref, carry = np.zeros((T, n)), np.zeros(n)
    for t in range(T - 1, -1, -1):                                   # train_ppo.py:171-202
        delta = rn[t] + gam * vn[t + 1] * (1 - ten[t]) - vn[t]
        carry = delta + gam * lam * (1 - np.maximum(ten[t], trnn[t])) * carry
        ref[t] = carry
    np.testing.assert_allclose(adv.cpu().numpy(), ref, rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(ret.cpu().numpy(), ref + vn[:-1], rtol=1e-5, atol=1e-5)


def test_graph_update_equals_eager(env):
    """Same trajectories, same minibatch permutations: the graph-replayed minibatch steps (4 epochs x 4 minibatches x 3 iterations,
    the first three steps eager warm-up, the fourth captured) give the same parameters as the eager loop."""
    a, b = _trainer(env, True), _trainer(env, False)

    def copy_rollout():                                              # trainer b learns from trainer a's rollouts
        for name in ("obs_traj", "act_traj", "logp_traj", "r_traj", "term_traj", "trunc_traj", "obs"):
            getattr(b, name).copy_(getattr(a, name))
    b.collect_rollout = copy_rollout
    for _ in range(3):
        ra = a.iteration()
        rb = b.iteration()
        assert np.isfinite(ra["train_return_avg"]) and ra["minibatches"] == 16
        for pa, pb in zip(a.policy + [a.log_std] + a.value, b.policy + [b.log_std] + b.value):
            torch.testing.assert_close(pa, pb, rtol=2e-3, atol=2e-4)
    assert a.upd["fb"] is not None


def test_linear_act_backward_matches_autograd():
    """The learner's linear(+tanh) layer with the fused tanh-backward / bias-gradient kernel gives torch autograd's gradients."""
    g = torch.Generator(device="cuda").manual_seed(1)
    for n, k, c, act in [(1000, 54, 256, True), (4097, 256, 256, True), (3000, 256, 21, False), (512, 256, 1, False),
                         (16384, 256, 256, True), (8192, 54, 256, False)]:      # the last two take the split-row weight gradient
        x = torch.randn(n, k, device="cuda", generator=g, requires_grad=True)
        w = (torch.randn(k, c, device="cuda", generator=g) * 0.1).requires_grad_()
        b = torch.randn(c, device="cuda", generator=g).requires_grad_()
        up = torch.randn(n, c, device="cuda", generator=g)
        y = ppo_mod._LinearAct.apply(x, w, b, act)
        gx, gw, gb = torch.autograd.grad((y * up).sum(), (x, w, b))
        y2 = torch.addmm(b, x, w)
        y2 = torch.tanh(y2) if act else y2
        rx, rw, rb = torch.autograd.grad((y2 * up).sum(), (x, w, b))
        torch.testing.assert_close(y, y2, rtol=1e-5, atol=1e-5)
        torch.testing.assert_close(gx, rx, rtol=1e-3, atol=1e-3)
        torch.testing.assert_close(gw, rw, rtol=1e-3, atol=5e-2)        # TF32 products summed over n rows
        torch.testing.assert_close(gb, rb, rtol=1e-4, atol=1e-2)
