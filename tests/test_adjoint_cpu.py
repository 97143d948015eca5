"""The reverse-mode env step (csrc/mjxb_adjoint.cuh, APG path: reference train_apg.py:161-209) checked WITHOUT a GPU.

The adjoint source is single-source host/device; tests/adjoint_host.cpp instantiates it in double on the CPU.  Its oracle is central
finite differences of the float64 CPU oracle (oracle/oracle.hpp: the restatement of mjx.step + src/envs.py single_step), with the
oracle's solver run to convergence (tolerance 1e-15) so that the implicit-function derivative is the derivative of what it computes.
Three levels: contact geometry adjoints, the inverse-dynamics gradient lam^T d(M a + c)/d(q, v), and the whole env step.
Tolerance: 2e-5 relative to max(1, |gradient|_inf) -- the measured agreement is 1e-7 (finite-difference noise).
"""
import ctypes as C

import numpy as np
import pytest

import helpers
from mujoco_mjx_lab_b200 import _abi, modelc

P = lambda a: a.ctypes.data_as(C.c_void_p) if a is not None else None
TOL = 2e-5


@pytest.fixture(scope="module")
def tight_model():
    return helpers.load(overrides=dict(tolerance=1e-15, iterations=200, ls_iterations=100, ls_tolerance=1e-6))


def test_contact_geometry_adjoints(model):
    L, blob = helpers.adjoint_host(), modelc.pack_blob(model)
    rng = np.random.default_rng(1)
    ng = model["ngeom"]
    kinds = {}
    for i, pr in enumerate(model["pairs"]):
        kinds.setdefault(pr["kind"], []).append(i)
    assert set(kinds) == {modelc.PAIR_PLANE_CAPSULE, modelc.PAIR_CAPSULE_CAPSULE, modelc.PAIR_SPHERE_CAPSULE}
    for kind, plist in kinds.items():
        for trial in range(12):
            pi = plist[rng.integers(len(plist))]
            pr = model["pairs"][pi]
            g1, g2 = pr["g1"], pr["g2"]
            gpos = rng.normal(size=(ng, 3)) * 0.15
            gaxis = rng.normal(size=(ng, 3))
            gaxis /= np.linalg.norm(gaxis, axis=1, keepdims=True)
            if g1 == 0:
                gpos[0], gaxis[0] = 0, [0, 0, 1]
            for e in range(2 if kind == modelc.PAIR_PLANE_CAPSULE else 1):
                outbar = rng.normal(size=13)
                if pr["condim"] == 1:
                    outbar[7:] = 0
                out, gbar = np.zeros(13), np.zeros(12)
                assert L.adj_contact(P(blob), pi, e, P(gpos), P(gaxis), P(out), P(outbar), P(gbar)) == 0

                def f(gp, ga):
                    o = np.zeros(13)
                    L.adj_contact(P(blob), pi, e, P(gp), P(ga), P(o), None, None)
                    return float(o @ outbar)
                eps, fd = 1e-6, np.zeros(12)
                for which, (is_axis, g) in enumerate(((0, g1), (1, g1), (0, g2), (1, g2))):
                    for k in range(3):
                        a, b = [gpos.copy(), gaxis.copy()], [gpos.copy(), gaxis.copy()]
                        a[is_axis][g, k] += eps
                        b[is_axis][g, k] -= eps
                        fd[3 * which + k] = (f(*a) - f(*b)) / (2 * eps)
                if g1 == 0:
                    fd[:6] = gbar[:6] = 0            # the plane belongs to the world body
                for which, g in ((1, g1), (3, g2)):  # axes are unit vectors: only the tangential part of their gradient is defined
                    ax = gaxis[g]
                    fd[3 * which:3 * which + 3] -= ax * (ax @ fd[3 * which:3 * which + 3])
                    gbar[3 * which:3 * which + 3] -= ax * (ax @ gbar[3 * which:3 * which + 3])
                assert np.abs(fd - gbar).max() <= TOL * max(1.0, np.abs(fd).max()), (kind, pi, e)


@pytest.mark.parametrize("kind", ["free", "tumble"])
def test_inverse_dynamics_gradient(model, oracle, kind):
    """d/d(q, v) of lam^T (M(q) a + c(q, v)) in tangent coordinates against finite differences of the oracle's crb / rne."""
    L, blob = helpers.adjoint_host(), modelc.pack_blob(model)
    rng = np.random.default_rng(0)
    q, v, _, _ = helpers.make_states(model, 1, 3, kind)
    q, v = q[0], v[0]
    lam, aa = rng.normal(size=27), rng.normal(size=27) * 3

    def scalar(qq, vv):
        out = oracle.forward(qq[None], vv[None], None, None, prec="f64", debug=("qM", "qfrc_bias"))
        return float(lam @ (out["qM"][0] @ aa + out["qfrc_bias"][0]))
    gqt, gv = np.zeros(27), np.zeros(27)
    assert L.adj_idgrad(P(blob), P(q), P(lam), P(v), P(aa), 1, P(gqt), P(gv)) == 0
    eps = 1e-6
    fd_q = np.array([(scalar(helpers.tangent_perturb(model, q, d, eps), v) - scalar(helpers.tangent_perturb(model, q, d, -eps), v)) / (2 * eps)
                     for d in range(27)])
    fd_v = np.zeros(27)
    for d in range(27):
        vp, vm = v.copy(), v.copy()
        vp[d] += eps
        vm[d] -= eps
        fd_v[d] = (scalar(q, vp) - scalar(q, vm)) / (2 * eps)
    assert np.abs(gqt - fd_q).max() <= TOL * max(1.0, np.abs(fd_q).max())
    assert np.abs(gv - fd_v).max() <= TOL * max(1.0, np.abs(fd_v).max())
    assert np.abs(fd_q[3:]).max() > 1.0 and np.abs(fd_q[:3]).max() < 1e-6      # translation invariance of M and c


def _env_case(model, cfg, kind, n, seed):
    rng = np.random.default_rng(seed)
    q, v, _, _ = helpers.make_states(model, n, 11 + seed, kind)
    act = rng.normal(size=(n, 21)) * 0.6
    aux = np.zeros((n, 9))
    aux[:, 0] = rng.random(n) < 0.5
    aux[:, 1], aux[:, 2], aux[:, 3] = q[:, 0] + 2.0 + rng.normal(size=n) * 0.1, q[:, 1] + rng.normal(size=n) * 0.3, q[:, 2]
    aux[:, 7], aux[:, 8] = -2.0 / 0.005, 3
    cot = dict(gq=rng.normal(size=(n, 28)), gv=rng.normal(size=(n, 27)), ga=np.zeros((n, 9)), gr=rng.normal(size=n))
    cot["ga"][:, [1, 2, 3, 7]] = rng.normal(size=(n, 4)) * np.array([1, 1, 1, 0.01])
    return q, v, aux, act, cot


def env_step_fd_check(model, cfg, kind, n, seed, prec=1, tol=TOL):
    orc = helpers.make_oracle(model, cfg)
    cc = _abi.make_env_config_c(cfg, 28, 27, 21)
    L, blob = helpers.adjoint_host(), modelc.pack_blob(model)
    q, v, aux, act, cot = _env_case(model, cfg, kind, n, seed)

    def run(q, v, aux, act):
        st = dict(qpos=q.copy(), qvel=v.copy(), qacc_warmstart=np.zeros_like(v), time=np.zeros(len(q)), aux=aux.copy())
        st2, _, r, _, _, _, _ = orc.env_step(st, act, prec="f64")
        return st2, r

    def loss(**kw):
        st2, r = run(kw["q"], kw["v"], kw["aux"], kw["act"])
        return (st2["qpos"] * cot["gq"]).sum(1) + (st2["qvel"] * cot["gv"]).sum(1) + (st2["aux"] * cot["ga"]).sum(1) + r * cot["gr"]
    st2, _ = run(q, v, aux, act)
    tape = st2["qacc_warmstart"].copy()
    g = dict(q=np.zeros((n, 28)), v=np.zeros((n, 27)), aux=np.zeros((n, 9)), act=np.zeros((n, 21)))
    status = np.zeros(n, dtype=np.int32)
    assert L.adj_step_vjp(P(blob), C.byref(cc), prec, n, P(q), P(v), P(aux), P(act), P(tape), P(cot["gq"]), P(cot["gv"]), P(cot["ga"]),
                          P(cot["gr"]), P(g["q"]), P(g["v"]), P(g["aux"]), P(g["act"]), P(status)) == 0
    assert (status == 0).all()
    eps, worst = 1e-6, {}
    base = dict(q=q, v=v, aux=aux, act=act)
    for name, dim, idxs in (("v", 27, None), ("act", 21, None), ("q", 28, None), ("aux", 9, [1, 2, 3, 7])):
        fd = np.zeros((n, dim))
        for i in (range(dim) if idxs is None else idxs):
            a, b = {k: x.copy() for k, x in base.items()}, {k: x.copy() for k, x in base.items()}
            a[name][:, i] += eps
            b[name][:, i] -= eps
            fd[:, i] = (loss(**a) - loss(**b)) / (2 * eps)
        err = np.abs(fd - g[name]).max(axis=1) / np.maximum(1.0, np.abs(fd).max(axis=1))
        worst[name] = float(err.max())
        assert err.max() <= tol, (kind, name, err)
    return worst, g, (q, v, aux, act, cot, tape)


@pytest.mark.parametrize("kind,posture", [("free", 0.0), ("stand", 0.0), ("lean", 0.6), ("tumble", 0.6)])
def test_env_step_vjp_against_finite_differences(tight_model, kind, posture):
    cfg = helpers.env_config(posture_penalty_weight=posture)
    worst, _, _ = env_step_fd_check(tight_model, cfg, kind, 3, 5)
    print(kind, "worst relative error vs finite differences", worst)


def test_env_step_vjp_cg_solver_apg_settings(tight_model):
    """train_apg.py:101-105 switches the solver to CG. The converged solution, hence the implicit-function gradient, is the same function
    of the inputs; CG's own termination leaves qacc ~1e-5 from the minimiser (its finite differences are noise), so the gradient taken at
    the CG tape is compared with the Newton one (itself checked against finite differences above): they agree to that residual."""
    cfg = helpers.env_config(random_flip=False, posture_penalty_weight=0.6)
    _, g_newton, (q, v, aux, act, cot, tape_newton) = env_step_fd_check(tight_model, cfg, "lean", 2, 9)
    m = dict(tight_model)
    m["opt"] = dict(m["opt"], solver=1)
    orc = helpers.make_oracle(m, cfg)
    st = dict(qpos=q.copy(), qvel=v.copy(), qacc_warmstart=np.zeros_like(v), time=np.zeros(len(q)), aux=aux.copy())
    st2, *_ = orc.env_step(st, act, prec="f64")
    tape = st2["qacc_warmstart"].copy()
    assert 1e-9 < np.abs(tape - tape_newton).max() < 1e-3
    cc = _abi.make_env_config_c(cfg, 28, 27, 21)
    L, blob, n = helpers.adjoint_host(), modelc.pack_blob(m), len(q)
    g = dict(q=np.zeros((n, 28)), v=np.zeros((n, 27)), aux=np.zeros((n, 9)), act=np.zeros((n, 21)))
    assert L.adj_step_vjp(P(blob), C.byref(cc), 1, n, P(q), P(v), P(aux), P(act), P(tape), P(cot["gq"]), P(cot["gv"]), P(cot["ga"]),
                          P(cot["gr"]), P(g["q"]), P(g["v"]), P(g["aux"]), P(g["act"]), None) == 0
    for k in g:
        assert np.abs(g[k] - g_newton[k]).max() <= 1e-3 * max(1.0, np.abs(g_newton[k]).max()), k


@pytest.mark.parametrize("xml,kind", [("humanoid_mjx", "lean"), ("humanoid", "stand")])
def test_physics_only_vjp_against_finite_differences(xml, kind):
    """mjxb_step_vjp with in.aux == NULL: the physics step alone (mjx.step; `action` is ctrl, clipped by ctrlrange only), on both
    humanoid models (humanoid.xml integrates with Euler + eulerdamp, the same implicit-damping solve)."""
    model = helpers.load(xml, overrides=dict(tolerance=1e-15, iterations=200, ls_iterations=100, ls_tolerance=1e-6))
    orc = helpers.make_oracle(model)
    L, blob = helpers.adjoint_host(), modelc.pack_blob(model)
    nq, nv, nu = model["nq"], model["nv"], model["nu"]
    rng = np.random.default_rng(7)
    n = 2
    q, v, _, _ = helpers.make_states(model, n, 21, kind)
    ctrl = rng.uniform(-0.9, 0.9, size=(n, nu))
    ctrl[:, 0] = 1.7                                           # saturated by ctrlrange: zero gradient
    gq, gv = rng.normal(size=(n, nq)), rng.normal(size=(n, nv))

    def run(q, v, ctrl):
        out = orc.physics_step(q, v, None, None, ctrl, prec="f64")
        return out

    def loss(q, v, ctrl):
        o = run(q, v, ctrl)
        return (o["qpos"] * gq).sum(1) + (o["qvel"] * gv).sum(1)
    tape = run(q, v, ctrl)["qacc_warmstart"].copy()
    g = dict(q=np.zeros((n, nq)), v=np.zeros((n, nv)), act=np.zeros((n, nu)))
    status = np.zeros(n, dtype=np.int32)
    assert L.adj_step_vjp(P(blob), None, 1, n, P(q), P(v), None, P(ctrl), P(tape), P(gq), P(gv), None, None, P(g["q"]), P(g["v"]), None,
                          P(g["act"]), P(status)) == 0
    assert (status == 0).all()
    eps = 1e-6
    base = dict(q=q, v=v, ctrl=ctrl)
    for name, key, dim in (("v", "v", nv), ("act", "ctrl", nu), ("q", "q", nq)):
        fd = np.zeros((n, dim))
        for i in range(dim):
            a, b = {k: x.copy() for k, x in base.items()}, {k: x.copy() for k, x in base.items()}
            a[key][:, i] += eps
            b[key][:, i] -= eps
            fd[:, i] = (loss(**a) - loss(**b)) / (2 * eps)
        err = np.abs(fd - g[name]).max(axis=1) / np.maximum(1.0, np.abs(fd).max(axis=1))
        assert err.max() <= TOL, (xml, name, err)
    assert np.abs(g["act"][:, 0]).max() == 0.0
