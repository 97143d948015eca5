"""Other models / solver settings through the same kernels: humanoid.xml (Euler + eulerdamp, Newton 100/50, all-pairs
collisions: 159 pairs / 303 rows, plane-sphere and sphere-sphere primitives) and lighten_solver (1 iteration, MJX's truncated
line search kept verbatim)."""
import numpy as np
import pytest
import torch

import helpers
from mujoco_mjx_lab_b200 import mjx
from test_gpu_parity import N, T, assert_f32_equivalent, rel

pytestmark = pytest.mark.gpu


def _compare(model, kind, seed, n=256, tol_floor=1e-4):
    orc = helpers.make_oracle(model)
    sysm = mjx.put_model(model)
    q, v, w, c = helpers.make_states(model, n, seed, kind)
    ref = orc.physics_step(q, v, w, None, c, prec="f64", debug=("qacc", "con_dist", "efc_active", "solver_niter"))
    r32 = orc.physics_step(q, v, w, None, c, prec="f32", debug=("qacc",))
    d = mjx.Data(T(q), T(v), T(w), torch.zeros(n, device="cuda"), T(c))
    _, dbg = mjx.forward(sysm, d, debug=True)
    nd = mjx.step(sysm, d)
    assert np.abs(N(dbg["con_dist"]) - ref["con_dist"]).max() < 2e-6
    cand_g, cand_r = dbg["efc_active"].cpu().numpy() & 1, ref["efc_active"] & 1
    ok = (cand_g == cand_r).all(axis=1)
    assert ok.mean() > 0.97
    assert_f32_equivalent(rel(N(dbg["qacc"]), ref["qacc"])[ok].max(axis=1), rel(r32["qacc"], ref["qacc"])[ok].max(axis=1), tol_floor, "qacc")
    assert_f32_equivalent(np.abs(N(nd.qvel) - ref["qvel"])[ok].max(axis=1), np.abs(r32["qvel"] - ref["qvel"])[ok].max(axis=1), 1e-5, "qvel")
    assert_f32_equivalent(np.abs(N(nd.qpos) - ref["qpos"])[ok].max(axis=1), np.abs(r32["qpos"] - ref["qpos"])[ok].max(axis=1), 1e-6, "qpos")
    return dbg, ref


@pytest.mark.parametrize("kind", ["lean", "tumble"])
def test_humanoid_xml(kind):
    model = helpers.load("humanoid")
    dbg, ref = _compare(model, kind, 400)
    assert (ref["efc_active"] & 1).sum(axis=1).max() > 20


def test_lighten_solver_iterative_linesearch():
    """reference src/training_utils.py:95-98: iterations = ls_iterations = 1 (the APG configuration)."""
    model = helpers.load(overrides=dict(iterations=1, ls_iterations=1))
    dbg, ref = _compare(model, "lean", 500, tol_floor=1e-3)
    assert int(dbg["solver_niter"].max()) == 1 and int(ref["solver_niter"].max()) == 1


def test_cg_solver_apg_settings():
    """reference train_apg.py:101-105 solver_options: CG with iterations = ls_iterations = 4 (Polak-Ribiere, M-preconditioned)."""
    from mujoco_mjx_lab_b200 import modelc
    model = helpers.load(overrides=dict(solver=modelc.SOLVER_CG, iterations=4, ls_iterations=4))
    dbg, ref = _compare(model, "lean", 600, tol_floor=1e-3)
    assert int(ref["solver_niter"].max()) == 4 and int(dbg["solver_niter"].max()) == 4


def test_speculative_reset_is_bit_identical(model):
    """Batches of one round (<= 14 envs per SM) re-initialise an env that finishes its episode with a reset warp beside the step
    (mjxb_abi.cu launch(); the verdict is published right after the stepping warp's kinematics): same bits as the deferred-reset path,
    over steps that do reset envs -- more of them than a CTA has reset warps, so the surplus takes the deferred path -- at sizes on
    both sides of the switch-over."""
    import helpers
    from mujoco_mjx_lab_b200 import _lib, training_utils
    env_a = training_utils.load_model_and_create_env("", helpers.env_config(), model=model)
    env_b = training_utils.load_model_and_create_env("", helpers.env_config(), model=model, flags=_lib.FLAG_NO_SPEC_RESET)
    assert env_b[9].sys.lib.mjxb_model_flags(env_b[9].sys.handle) & _lib.FLAG_NO_SPEC_RESET
    assert env_a[9].sys.lib.mjxb_model_flags(env_a[9].sys.handle) & _lib.FLAG_NO_SPEC_RESET == 0
    # (<= 2072 envs: one round of stepping + reset warps; up to 12,432 envs: two to six such rounds; forced resets exceed the reset
    #  warps of some CTAs; 12,500 envs: seven rounds, deferred resets in both models)
    for n in (1, 37, 1024, 1184, 1185, 2048, 2072, 2073, 4100, 8288, 12432, 12500):
        keys = helpers.ppo_keys(n, n)
        sa, oa = env_a[8](keys)
        sb, ob = env_b[8](keys)
        sa[1][: max(1, n // 7), 8] = 999.0                     # force truncations -> resets in the very first step
        sb[1][: max(1, n // 7), 8] = 999.0
        g = torch.Generator(device="cuda").manual_seed(n)
        resets = 0
        for t in range(12):
            act = torch.randn(n, 21, device="cuda", generator=g) * (3.0 if t % 3 == 0 else 1.0)
            rk = helpers.ppo_keys(100 + t, n)
            sa, oa, ra, tea, tra = env_a[9].autoreset(sa, act, rk)
            sb, ob, rb, teb, trb = env_b[9].autoreset(sb, act, rk)
            resets += int(torch.maximum(tea, tra).sum())
            assert torch.equal(oa, ob) and torch.equal(ra, rb) and torch.equal(tea, teb) and torch.equal(tra, trb), (n, t)
            assert torch.equal(sa[0].qpos, sb[0].qpos) and torch.equal(sa[0].qvel, sb[0].qvel) and torch.equal(sa[1], sb[1]), (n, t)
            assert torch.equal(sa[0].qacc_warmstart, sb[0].qacc_warmstart) and torch.equal(sa[0].time, sb[0].time), (n, t)
        assert resets >= max(1, n // 7)


def test_work_sorted_schedule_is_bit_identical(model):
    """Large batches deal the envs to the CTAs by descending cost of their previous step (mjxb_abi.cu launch(), mjxb_sort_work_kernel)
    and let a CTA take its next group of envs from a device-wide counter (dynamic rounds) instead of a static (round, CTA) assignment:
    the order in which envs are processed changes, no result does -- device API over steps that reset envs
    (a batch size that is not a multiple of the sort segment), and the pinned host pipeline whose input chunks are the sort segments."""
    import helpers
    from mujoco_mjx_lab_b200 import _lib, training_utils
    env_a = training_utils.load_model_and_create_env("", helpers.env_config(), model=model)
    env_b = training_utils.load_model_and_create_env("", helpers.env_config(), model=model, flags=_lib.FLAG_NO_WORK_SORT | _lib.FLAG_NO_DYN_ROUNDS)
    assert env_b[9].sys.lib.mjxb_model_flags(env_b[9].sys.handle) & _lib.FLAG_NO_WORK_SORT
    assert env_b[9].sys.lib.mjxb_model_flags(env_b[9].sys.handle) & _lib.FLAG_NO_DYN_ROUNDS
    assert env_a[9].sys.lib.mjxb_model_flags(env_a[9].sys.handle) & _lib.FLAG_NO_DYN_ROUNDS == 0
    assert env_a[9].sys.lib.mjxb_model_flags(env_a[9].sys.handle) & _lib.FLAG_NO_WORK_SORT == 0
    n = 40000 + 123
    keys = helpers.ppo_keys(5, n)
    sa, oa = env_a[8](keys)
    sb, ob = env_b[8](keys)
    sa[1][::7, 8] = 999.0                                    # force truncations -> resets in the very first step
    sb[1][::7, 8] = 999.0
    launches0 = env_a[9].sys.lib.mjxb_launch_count()
    g = torch.Generator(device="cuda").manual_seed(3)
    resets = 0
    for t in range(24):
        act = torch.randn(n, 21, device="cuda", generator=g) * (3.0 if t % 3 == 0 else 1.0)
        rk = helpers.ppo_keys(200 + t, n)
        sa, oa, ra, tea, tra = env_a[9].autoreset(sa, act, rk)
        sb, ob, rb, teb, trb = env_b[9].autoreset(sb, act, rk)
        resets += int(torch.maximum(tea, tra).sum())
        assert torch.equal(oa, ob) and torch.equal(ra, rb) and torch.equal(tea, teb) and torch.equal(tra, trb), t
        assert torch.equal(sa[0].qpos, sb[0].qpos) and torch.equal(sa[0].qvel, sb[0].qvel) and torch.equal(sa[1], sb[1]), t
        assert torch.equal(sa[0].qacc_warmstart, sb[0].qacc_warmstart) and torch.equal(sa[0].time, sb[0].time), t
    assert resets > 0
    assert env_a[9].sys.lib.mjxb_launch_count() - launches0 == 24 * (4 + 3)      # sorted: 4 launches per step; plain: 3
