"""GPU reverse-mode env step (mjxb_step_fwd_tape / mjxb_step_vjp, the APG path: reference train_apg.py:161-209).

Oracles: (1) the host double-precision build of the SAME adjoint source (tests/adjoint_host.cpp), itself pinned to finite differences of
the float64 CPU oracle by tests/test_adjoint_cpu.py; (2) directly, finite differences of the float64 oracle over a two-step rollout,
which also covers the chaining done by the torch.autograd wrapper.  Stated float32 tolerance, per env e_k = |g_gpu - g_ref|_inf / max(1, |g_ref|_inf): median over envs <= 2e-3 and at least 97 % of the
envs <= 2e-2 (the pull-back goes through two 27x27 solves in float32 with fast-math; the remaining envs have a constraint row whose
activity Jaref < 0 is decided within float32 rounding of zero, which selects a different H -- they are counted and printed).
"""
import ctypes as C

import numpy as np
import pytest
import torch

import helpers
from mujoco_mjx_lab_b200 import _abi, _lib, apg, modelc, training_utils
from mujoco_mjx_lab_b200.mjx import Data, state_c

pytestmark = pytest.mark.gpu
P = lambda a: a.ctypes.data_as(C.c_void_p) if a is not None else None


def T(a):
    return torch.tensor(np.asarray(a), dtype=torch.float32, device="cuda")


def N(t):
    return t.detach().double().cpu().numpy()


def _env(model, posture=0.6, **kw):
    cfg = helpers.env_config(posture_penalty_weight=posture)
    return cfg, training_utils.load_model_and_create_env("", cfg, model=model, **kw)


def _gpu_vjp(sysm, q, v, aux, act, cot):
    n = q.shape[0]
    tq, tv, ta, tu = T(q), T(v), T(aux), T(act)
    warm, tm = torch.zeros(n, 27, device="cuda"), torch.zeros(n, device="cuda")
    q2, v2, w2, t2, a2 = torch.empty_like(tq), torch.empty_like(tv), torch.empty_like(tv), torch.empty_like(tm), torch.empty_like(ta)
    obs, r, te, tr = torch.empty(n, 54, device="cuda"), torch.empty(n, device="cuda"), torch.empty(n, device="cuda"), torch.empty(n, device="cuda")
    L = sysm.lib
    st = torch.cuda.current_stream().cuda_stream
    _lib.check(L.mjxb_step_fwd_tape(sysm.handle, n, state_c(tq, tv, warm, tm, ta), tu.data_ptr(), state_c(q2, v2, w2, t2, a2), obs.data_ptr(),
                                    r.data_ptr(), te.data_ptr(), tr.data_ptr(), w2.data_ptr(), None, C.c_void_p(st)))
    gq, gv, ga, gr = T(cot["gq"]), T(cot["gv"]), T(cot["ga"]), T(cot["gr"])
    oq, ov, oa, ou = torch.empty_like(tq), torch.empty_like(tv), torch.empty_like(ta), torch.empty_like(tu)
    status = torch.zeros(n, dtype=torch.int32, device="cuda")
    _lib.check(L.mjxb_step_vjp(sysm.handle, n, state_c(tq, tv, warm, tm, ta), tu.data_ptr(), w2.data_ptr(), gq.data_ptr(), gv.data_ptr(),
                               ga.data_ptr(), gr.data_ptr(), oq.data_ptr(), ov.data_ptr(), oa.data_ptr(), ou.data_ptr(), status.data_ptr(), C.c_void_p(st)))
    torch.cuda.synchronize()
    return dict(q=N(oq), v=N(ov), aux=N(oa), act=N(ou)), N(w2), status.cpu().numpy(), (N(q2), N(v2), N(a2), N(r))


def _host_vjp(model, cfg, q, v, aux, act, tape, cot):
    n = q.shape[0]
    cc = _abi.make_env_config_c(cfg, 28, 27, 21)
    L, blob = helpers.adjoint_host(), modelc.pack_blob(model)
    g = dict(q=np.zeros((n, 28)), v=np.zeros((n, 27)), aux=np.zeros((n, 9)), act=np.zeros((n, 21)))
    f = lambda a: np.ascontiguousarray(np.asarray(a, dtype=np.float32).astype(np.float64))
    q, v, aux, act, tape = f(q), f(v), f(aux), f(act), f(tape)
    assert L.adj_step_vjp(P(blob), C.byref(cc), 1, n, P(q), P(v), P(aux), P(act), P(tape), P(f(cot["gq"])), P(f(cot["gv"])), P(f(cot["ga"])),
                          P(f(cot["gr"])), P(g["q"]), P(g["v"]), P(g["aux"]), P(g["act"]), None) == 0
    return g


def _case(model, kind, n, seed):
    rng = np.random.default_rng(seed)
    q, v, _, _ = helpers.make_states(model, n, 40 + seed, kind)
    act = rng.normal(size=(n, 21)) * 0.6
    aux = np.zeros((n, 9))
    aux[:, 0] = rng.random(n) < 0.5
    aux[:, 1], aux[:, 2], aux[:, 3] = q[:, 0] + 2.0 + rng.normal(size=n) * 0.1, q[:, 1] + rng.normal(size=n) * 0.3, q[:, 2]
    aux[:, 7], aux[:, 8] = -2.0 / 0.005, 3
    cot = dict(gq=rng.normal(size=(n, 28)), gv=rng.normal(size=(n, 27)), ga=np.zeros((n, 9)), gr=rng.normal(size=n))
    cot["ga"][:, [1, 2, 3, 7]] = rng.normal(size=(n, 4)) * np.array([1, 1, 1, 0.01])
    return q, v, aux, act, cot


def _rel_err(g, ref):
    return {k: np.abs(g[k] - ref[k]).max(axis=1) / np.maximum(1.0, np.abs(ref[k]).max(axis=1)) for k in ref}


@pytest.mark.parametrize("kind", ["free", "stand", "lean", "tumble"])
def test_gpu_vjp_matches_host_double_build(model, kind):
    cfg, env = _env(model)
    sysm = env[9].sys
    n = 256
    q, v, aux, act, cot = _case(model, kind, n, 1)
    g, tape, status, _ = _gpu_vjp(sysm, q, v, aux, act, cot)
    ref = _host_vjp(model, cfg, q, v, aux, act, tape, cot)
    err = _rel_err(g, ref)
    for k, e in err.items():
        print(f"[{kind}] d/d{k}: median rel err {np.median(e):.2e}  p99 {np.percentile(e, 99):.2e}  max {e.max():.2e}")
        assert np.isfinite(g[k]).all()
        frac_bad = float((e > 2e-2).mean())
        print(f"[{kind}] d/d{k}: envs above 2e-2: {int((e > 2e-2).sum())} of {e.size}")
        assert np.median(e) <= 2e-3 and frac_bad <= 0.03, (kind, k, float(np.median(e)), frac_bad, float(e.max()))
    assert (status & 1 == 0).all()


def test_gpu_vjp_overflow_tier(model):
    """States folded far beyond the joint limits and sunk into the floor need more than the 64-row tile: the 320-row instantiation
    re-runs them inside the same call (status bit ROW_SPILL) with the same result as the host build, which always uses the large tile."""
    cfg, env = _env(model)
    sysm = env[9].sys
    n = 64
    q, v, aux, act, cot = _case(model, "crumple", n, 2)
    g, tape, status, _ = _gpu_vjp(sysm, q, v, aux, act, cot)
    assert ((status & 2) != 0).sum() >= n // 2, "the crumple family is expected to overflow the main tile"
    ref = _host_vjp(model, cfg, q, v, aux, act, tape, cot)
    err = _rel_err(g, ref)
    for k, e in err.items():
        print(f"[crumple] d/d{k}: median rel err {np.median(e):.2e} max {e.max():.2e}")
        assert np.isfinite(g[k]).all() and np.median(e) <= 1e-2


def test_autograd_two_step_rollout_against_finite_differences(model):
    """diff_step chained by torch.autograd over two env steps, loss = sum of rewards + a linear functional of the final state, against
    central finite differences of the float64 oracle (solver run to convergence) over the same two steps."""
    tight = helpers.load(overrides=dict(tolerance=1e-15, iterations=200, ls_iterations=100, ls_tolerance=1e-6))
    cfg = helpers.env_config(posture_penalty_weight=0.6)
    orc = helpers.make_oracle(tight, cfg)
    _, env = _env(model)
    sysm = env[9].sys
    n = 8
    q, v, aux, act, cot = _case(model, "lean", n, 3)
    act2 = np.random.default_rng(4).normal(size=(n, 21)) * 0.5

    def oracle_loss(v_, act_, act2_):
        st = dict(qpos=q.copy(), qvel=v_.copy(), qacc_warmstart=np.zeros_like(v_), time=np.zeros(n), aux=aux.copy())
        st, _, r1, *_ = orc.env_step(st, act_, prec="f64")
        st, _, r2, *_ = orc.env_step(st, act2_, prec="f64")
        return r1 + 0.99 * r2 + (st["qpos"] * cot["gq"]).sum(1) + (st["qvel"] * cot["gv"]).sum(1)
    tq, tv, ta = T(q), T(v).requires_grad_(), T(aux)
    tu1, tu2 = T(act).requires_grad_(), T(act2).requires_grad_()
    state = (Data(tq, tv, torch.zeros(n, 27, device="cuda"), torch.zeros(n, device="cuda")), ta)
    state, _, r1, _, _ = apg.diff_step(sysm, state, tu1)
    state, _, r2, _, _ = apg.diff_step(sysm, state, tu2)
    loss = (r1 + 0.99 * r2 + (state[0].qpos * T(cot["gq"])).sum(1) + (state[0].qvel * T(cot["gv"])).sum(1)).sum()
    loss.backward()
    eps = 1e-6
    for name, t, arr in (("qvel", tv, v), ("action1", tu1, act), ("action2", tu2, act2)):
        fd = np.zeros_like(arr)
        for i in range(arr.shape[1]):
            args = dict(v_=v.copy(), act_=act.copy(), act2_=act2.copy())
            key = {"qvel": "v_", "action1": "act_", "action2": "act2_"}[name]
            args[key][:, i] += eps
            lp = oracle_loss(**args)
            args[key][:, i] -= 2 * eps
            fd[:, i] = (lp - oracle_loss(**args)) / (2 * eps)
        e = np.abs(N(t.grad) - fd).max(axis=1) / np.maximum(1.0, np.abs(fd).max(axis=1))
        print(f"two-step rollout d/d{name}: median rel err {np.median(e):.2e} max {e.max():.2e}")
        assert np.median(e) <= 5e-3 and e.max() <= 5e-2, (name, e)


def test_apg_trainer_runs_and_reduces_nothing_to_nan(model):
    """The corrected APG driver (reference train_apg.py:96-209) on a tiny batch: CG 4/4 solver, finite loss, non-zero clipped gradient."""
    cfg, env = apg.make_apg_env(model)
    cfg.horizon, cfg.hidden_size = 8, 32
    tr = apg.APGTrainer(cfg, env[8], env[9], 64)
    out = [tr.update() for _ in range(3)]
    assert all(np.isfinite(o["loss"]) and np.isfinite(o["grad_norm"]) for o in out)
    assert out[0]["grad_norm"] > 0
    assert env[9].sys.opt.solver == 1 and env[9].sys.opt.iterations == 4 and env[9].sys.opt.ls_iterations == 4
