"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on identical seeded inputs.

Stated float32 tolerances (north_star: "within a stated float32 tolerance of mjx.step"; the oracle stands in for mjx.step,
PARITY UNPINNED against real MJX, see oracle/oracle.hpp):
  kinematics / mass matrix / smooth forces : abs 2e-6 (xpos, xquat); qM 1e-5*max(1,|ref|) and forces 2e-4*max(1,|ref|), each plus
      3x the float32 oracle's own error
  contact geometry                         : |dist| 2e-6, pos 1e-5 on near contacts
  constraint rows                          : efc_D 1e-3 rel, efc_aref 5e-3*max(1,|ref|)
  solver outputs (qacc, efc_force) and the integrated state: the GPU's error against the float64 oracle must be within
      3x the error the oracle's own float32 instantiation makes on the same input (+ 1e-5 + 1e-4|ref|), because the
      Newton solve amplifies rounding by the conditioning of H; medians are additionally bounded absolutely.
  integer outputs (candidate / active masks, terminated, truncated, reset mask): bit-exact, except rows/envs whose
      deciding float is within a few ulp of its threshold in the float64 oracle (counted and bounded).
"""
import numpy as np
import pytest
import torch

import helpers
from mujoco_mjx_lab_b200 import mjx, training_utils

pytestmark = pytest.mark.gpu
KINDS = ["free", "stand", "lean", "tumble"]


@pytest.fixture(scope="module")
def sysm(model):
    return mjx.put_model(model)


@pytest.fixture(scope="module")
def env(model):
    cfg = helpers.env_config()
    return training_utils.load_model_and_create_env("", cfg, model=model)


def T(a):
    return torch.tensor(np.asarray(a), dtype=torch.float32, device="cuda")


def N(t):
    return t.detach().double().cpu().numpy()


def rel(a, ref):
    return np.abs(a - ref) / np.maximum(1.0, np.abs(ref))


def assert_f32_equivalent(err_gpu, err_f32, floor, what=""):
    """The GPU's error against the float64 oracle is distributed like the float32 oracle's own error on the same inputs:
    median within 2x, 99th percentile within 3x (each plus `floor`), and at most 0.5 % of the samples (never more than a handful)
    above 5x the float32 oracle's maximum. Rounding order differs between a sequential CPU evaluation and warp reductions, and the
    Newton solve amplifies it by the conditioning of H, so single maxima are heavy-tailed in BOTH float32 evaluations."""
    eg, e32 = np.asarray(err_gpu).ravel(), np.asarray(err_f32).ravel()
    stats = [(float(np.median(eg)), float(np.median(e32)), 2.0), (float(np.percentile(eg, 99)), float(np.percentile(e32, 99)), 3.0)]
    for g, r, k in stats:
        assert g <= k * r + floor, (what, stats, float(eg.max()), float(e32.max()))
    outliers = int((eg > 5.0 * e32.max() + floor).sum())
    assert outliers <= max(1, int(0.005 * eg.size)), (what, "outliers", outliers, eg.size, float(eg.max()), float(e32.max()))


@pytest.mark.parametrize("kind", KINDS)
def test_forward_stages(model, oracle, sysm, kind):
    n = 512
    q, v, w, c = helpers.make_states(model, n, 100 + KINDS.index(kind), kind)
    ref = oracle.forward(q, v, w, c, prec="f64", debug=True)
    r32 = oracle.forward(q, v, w, c, prec="f32", debug=True)
    _, out = mjx.forward(sysm, mjx.Data(T(q), T(v), T(w), torch.zeros(n, device="cuda"), T(c)), debug=True)
    g = {k: (t.cpu().numpy() if t.dtype == torch.int32 else N(t)) for k, t in out.items()}
    assert not np.isnan(g["qacc"]).any()
    assert np.abs(g["xpos"] - ref["xpos"]).max() < 2e-6 and np.abs(g["xquat"] - ref["xquat"]).max() < 2e-6
    assert rel(g["qM"], ref["qM"]).max() < 3 * rel(r32["qM"], ref["qM"]).max() + 1e-5
    for name in ("qfrc_bias", "qfrc_passive", "qfrc_actuator"):
        assert rel(g[name], ref[name]).max() < 3 * rel(r32[name], ref[name]).max() + 2e-4, name
    assert np.abs(g["con_dist"] - ref["con_dist"]).max() < 2e-6
    near = ref["con_dist"] < 0.05
    # closest points of nearly parallel / nearly intersecting capsules are ill-conditioned: calibrate on the float32 oracle
    assert_f32_equivalent(np.abs(g["con_pos"] - ref["con_pos"])[near], np.abs(r32["con_pos"] - ref["con_pos"])[near], 1e-5, "con_pos")
    assert_f32_equivalent(np.abs(g["con_normal"] - ref["con_normal"])[near], np.abs(r32["con_normal"] - ref["con_normal"])[near], 1e-4, "con_normal")
    # candidate mask: bit-exact unless the float64 position is within 1e-6 of zero
    cand_g, cand_r = g["efc_active"] & 1, ref["efc_active"] & 1
    pos64 = np.where(cand_r == 1, ref["efc_pos"], 0.0)
    mism = cand_g != cand_r
    if mism.any():
        dist_rows = _row_positions(model, ref)
        assert (np.abs(dist_rows[mism]) < 1e-6).all(), "candidate mask differs away from the threshold"
    print(f"[{kind}] candidate-mask mismatches (all at the threshold): {int(mism.sum())} of {mism.size}")
    assert int(mism.sum()) <= 2                                      # measured: 0 on all four families (tools/parity_counts_probe.py)
    both = (cand_g == 1) & (cand_r == 1)
    assert rel(g["efc_pos"], ref["efc_pos"])[both].max() < 2e-6
    assert (np.abs(g["efc_D"] - ref["efc_D"])[both] / ref["efc_D"][both]).max() < 1e-3
    b32 = both & ((r32["efc_active"] & 1) == 1)
    assert_f32_equivalent(rel(g["efc_aref"], ref["efc_aref"])[b32], rel(r32["efc_aref"], ref["efc_aref"])[b32], 5e-3, "efc_aref")
    # solver outputs: calibrated against the float32 oracle's own error
    ok_env = ~mism.any(axis=1)
    for name in ("qacc_smooth", "qacc", "qfrc_constraint", "efc_force"):
        eg, e32 = rel(g[name], ref[name])[ok_env].max(axis=1), rel(r32[name], ref[name])[ok_env].max(axis=1)
        assert_f32_equivalent(eg, e32, 1e-4, name)
        assert np.median(eg) < 2e-3, (name, np.median(eg))
    # active mask at the solution: mismatches only where the float64 force is tiny relative to the env's largest force
    act_g, act_r = g["efc_active"] >> 1, ref["efc_active"] >> 1
    bad = (act_g != act_r) & ok_env[:, None]
    fmax = np.maximum(ref["efc_force"].max(axis=1, keepdims=True), 1.0)
    jar = ref["efc_J"] @ ref["qacc"][:, :, None]
    jar = jar[:, :, 0] - ref["efc_aref"]
    thresh = np.abs(ref["efc_D"] * jar)                       # |force| the row would carry / is carrying
    # measured (tools/parity_counts_probe.py): 0 / 1 / 0 / 0 rows of 512 x 187, the one being a row whose force is ~0 in the float64 oracle
    # and which the float32 ORACLE flips as well
    print(f"[{kind}] active-mask mismatches: {int(bad.sum())} of {bad.size}")
    assert int(bad.sum()) <= 3
    assert (thresh[bad] < 1e-3 * np.broadcast_to(fmax, thresh.shape)[bad]).all(), "active mask differs away from the threshold"
    assert (g["efc_force"] >= 0).all()
    sens_g, sens_r = g["sensordata"] > 0, ref["sensordata"] > 0
    print(f"[{kind}] touch-sensor sign mismatches: {int((sens_g != sens_r)[ok_env].sum())} of {sens_g[ok_env].size}")
    assert int((sens_g != sens_r)[ok_env].sum()) <= 1               # measured: 0
    assert (g["status"][ok_env] & 1 == 0).all()


def _row_positions(model, ref):
    """Unmasked float64 constraint position of every static row (joint limits, tendon limits, contact slots)."""
    n = ref["con_dist"].shape[0]
    rows = np.zeros((n, model["nefc"]))
    # only contact rows can sit near the threshold in these tests; limit rows are reconstructed as 'far' unless candidates
    rows[:] = 1.0
    rows[:, :model["nlimit"] + model["ntlimit"]] = np.where((ref["efc_active"][:, :23] & 1) == 1, ref["efc_pos"][:, :23], 1.0)
    for p in model["pairs"]:
        for e in range(p["ncon"]):
            d = ref["con_dist"][:, p["con_adr"] + e]
            if p["condim"] == 1:
                rows[:, p["efc_adr"] + e] = d
            else:
                rows[:, p["efc_adr"] + 4 * e: p["efc_adr"] + 4 * e + 4] = d[:, None]
    return rows


@pytest.mark.parametrize("kind", KINDS)
def test_one_step_state(model, oracle, sysm, kind):
    n = 512
    q, v, w, c = helpers.make_states(model, n, 200 + KINDS.index(kind), kind)
    ref = oracle.physics_step(q, v, w, None, c, prec="f64")
    r32 = oracle.physics_step(q, v, w, None, c, prec="f32")
    nd = mjx.step(sysm, mjx.Data(T(q), T(v), T(w), torch.zeros(n, device="cuda"), T(c)))
    for name, t in (("qpos", nd.qpos), ("qvel", nd.qvel), ("qacc_warmstart", nd.qacc_warmstart)):
        g = N(t)
        eg, e32 = np.abs(g - ref[name]), np.abs(r32[name] - ref[name])
        tol = 3.0 * e32.max() + 1e-5 + 1e-4 * np.abs(ref[name])
        assert (eg <= tol).all(), (name, eg.max(), e32.max())
        # medians measured (tools/parity_counts_probe.py): qvel 2e-5 .. 6e-5, qacc_warmstart 6e-3 .. 1.4e-2 (float32 oracle: 2e-5 .. 3e-5, 6e-3 .. 1e-2)
        assert np.median(eg.max(axis=1)) < {"qpos": 1e-6, "qvel": 1.5e-4, "qacc_warmstart": 3e-2}[name], (name, np.median(eg.max(axis=1)))
        assert np.median(eg.max(axis=1)) <= 2.5 * np.median(e32.max(axis=1)) + 1e-7, (name, np.median(eg.max(axis=1)), np.median(e32.max(axis=1)))
    np.testing.assert_allclose(N(nd.time), 0.005, rtol=1e-6)
    np.testing.assert_allclose(np.linalg.norm(N(nd.qpos)[:, 3:7], axis=1), 1.0, atol=1e-6)


def test_multi_step_matches_single_steps(model, sysm):
    """nsteps in one launch == the same number of single-step launches, bitwise."""
    n = 256
    q, v, w, c = helpers.make_states(model, n, 5, "lean")
    d = mjx.Data(T(q), T(v), T(w), torch.zeros(n, device="cuda"), T(c))
    a = mjx.step(sysm, d, nsteps=5)
    b = d
    for _ in range(5):
        b = mjx.step(sysm, b)
    assert torch.equal(a.qpos, b.qpos) and torch.equal(a.qvel, b.qvel) and torch.equal(a.qacc_warmstart, b.qacc_warmstart)


def test_free_running_128_steps(model, oracle, sysm):
    """Free-running 128 steps: free flight and quiet standing stay within 1e-3 of the float64 oracle."""
    n = 64
    for kind, tol in (("free", 1e-3), ("stand", 1e-3)):
        q, v, w, c = helpers.make_states(model, n, 300, kind)
        if kind == "stand":
            v *= 0.0
        c = np.zeros_like(c)
        ref = oracle.physics_step(q, v, None, None, c, nsteps=128, prec="f64")
        r32 = oracle.physics_step(q, v, None, None, c, nsteps=128, prec="f32")
        nd = mjx.step(sysm, mjx.Data(T(q), T(v), torch.zeros(n, 27, device="cuda"), torch.zeros(n, device="cuda"), T(c)), nsteps=128)
        eg, e32 = np.abs(N(nd.qpos) - ref["qpos"]).max(axis=1), np.abs(r32["qpos"] - ref["qpos"]).max(axis=1)
        print(f"128-step free-running [{kind}] max|dqpos| gpu {eg.max():.2e} (median {np.median(eg):.2e}); oracle-f32 {e32.max():.2e}")
        assert np.median(eg) < tol
        assert eg.max() < max(tol, 3 * e32.max())


def test_reset_parity(model, oracle, env):
    v_reset = env[8]
    n = 1024
    keys = helpers.ppo_keys(42, n)
    (d, aux), obs = v_reset(keys)
    st, o_ref = oracle.env_reset(keys, prec="f32")
    # the noise arithmetic is IEEE-identical: bit-exact initial state
    np.testing.assert_array_equal(N(d.qpos), st["qpos"])
    np.testing.assert_array_equal(N(d.qvel), st["qvel"])
    np.testing.assert_array_equal(N(aux)[:, 0], st["aux"][:, 0])          # flip draws
    assert (N(d.time) == 0).all()
    np.testing.assert_allclose(N(aux)[:, 1:5], st["aux"][:, 1:5], atol=1e-6)
    np.testing.assert_allclose(N(aux)[:, 7], st["aux"][:, 7], atol=2e-3)  # -dist/dt: 200x gain
    # stance state may legitimately differ when a foot sits exactly on the floor (f32 threshold): counted; measured 0 of 1024
    n_stance = int((N(aux)[:, 5] != st["aux"][:, 5]).sum())
    print(f"reset: stance-state mismatches {n_stance} of {n}")
    assert n_stance <= 2
    np.testing.assert_allclose(N(obs), o_ref, atol=1e-5)
    s64, _ = oracle.env_reset(keys, prec="f64")
    e_g = np.abs(N(d.qacc_warmstart) - s64["qacc_warmstart"]).max()
    e_32 = np.abs(st["qacc_warmstart"] - s64["qacc_warmstart"]).max()
    assert e_g <= 3 * e_32 + 1e-3 or int((np.abs(N(d.qacc_warmstart) - s64["qacc_warmstart"]).max(axis=1) > 3 * e_32 + 1e-3).sum()) <= 1


def test_env_step_resynchronised_128(model, oracle, env):
    """128 consecutive oracle states, each advanced one env step by both paths (SURVEY section 7 're-synchronised')."""
    v_step = env[9]
    n, steps = 64, 128
    rng = np.random.default_rng(0)
    st, _ = oracle.env_reset(helpers.ppo_keys(1, n), prec="f32")
    worst = dict(obs=0.0, reward=0.0, qpos=0.0, qvel=0.0)
    errs = dict(obs_g=[], obs_32=[], r_g=[], r_32=[])
    flags = 0
    resets = 0
    for t in range(steps):
        act = rng.normal(size=(n, 21))
        rk = helpers.ppo_keys(1000 + t, n)
        d = mjx.Data(T(st["qpos"]), T(st["qvel"]), T(st["qacc_warmstart"]), T(st["time"]))
        (d2, aux2), obs, rew, te, tr = v_step.autoreset((d, T(st["aux"])), T(act), rk)
        s64 = {k: v.copy() for k, v in st.items()}
        st, o_ref, r_ref, te_ref, tr_ref, mask, _ = oracle.env_step(st, act, prec="f32", reset_keys=rk)
        _, o64, r64, te64, tr64, _, dbg64 = oracle.env_step(s64, act, prec="f64", reset_keys=rk, debug=("xpos",))
        height64 = dbg64["xpos"][:, 4, 2]
        te_g, tr_g = N(te), N(tr)
        near = np.abs(height64 - 0.7) < 1e-5
        assert (te_g == te_ref)[~near].all() and (tr_g == tr_ref).all()
        flags += int(near.sum())
        same = (te_g == te_ref) & (np.maximum(te_ref, tr_ref) == 0)
        resets += int(mask.sum())
        # envs that were reset in both: state is the bit-exact reset state; others: one-step tolerance
        done = (np.maximum(te_g, tr_g) > 0) & (te_g == te_ref)
        if done.any():
            np.testing.assert_array_equal(N(d2.qpos)[done], st["qpos"][done])
        e_obs_g, e_obs_32 = np.abs(N(obs) - o64)[same], np.abs(o_ref - o64)[same]
        e_r_g, e_r_32 = np.abs(N(rew) - r64)[same], np.abs(r_ref - r64)[same]
        errs["obs_g"].append(e_obs_g.max(axis=1)); errs["obs_32"].append(e_obs_32.max(axis=1))
        errs["r_g"].append(e_r_g); errs["r_32"].append(e_r_32)
        worst["obs"] = max(worst["obs"], e_obs_g.max()); worst["reward"] = max(worst["reward"], e_r_g.max())
        worst["qpos"] = max(worst["qpos"], np.abs(N(d2.qpos) - st["qpos"])[same].max())
        worst["qvel"] = max(worst["qvel"], np.abs(N(d2.qvel) - st["qvel"])[same].max())
        np.testing.assert_allclose(N(aux2)[same][:, [0, 4, 8]], st["aux"][same][:, [0, 4, 8]])
    print("resynchronised 128 steps: worst errors", worst, "near-threshold terminations", flags, "resets", resets)
    assert resets > 0
    assert_f32_equivalent(np.concatenate(errs["obs_g"]), np.concatenate(errs["obs_32"]), 1e-4, "obs")       # obs: 1e-4 abs floor
    assert_f32_equivalent(np.concatenate(errs["r_g"]), np.concatenate(errs["r_32"]), 2e-3, "reward")        # reward: 1/dt gain on dist
    assert worst["qpos"] < 1e-4 and worst["qvel"] < 5e-2
    assert flags <= 2, "terminated flags that differ at the 0.7 m threshold (counted)"


def test_speed_test_semantics(model, oracle, sysm):
    """mjx_humanoid_speed_test.py step: cold start from qpos0 with qvel[0] = linspace(0,1,N); output qpos[0]."""
    n = 64                                                    # BASELINE.json configs[0]
    vel = np.linspace(0, 1, n).astype(np.float32)
    pos = N(mjx.speed_test(sysm, T(vel), iters=3))
    ref = oracle.speed_test(vel, iters=1, prec="f64")
    np.testing.assert_allclose(pos, ref, atol=2e-6)           # feet sit on the contact threshold here
    np.testing.assert_allclose(pos, vel.astype(np.float64) * 0.005, atol=5e-5)


def test_host_buffer_api(model, env):
    """mjxb_*_host (H2D -> launch -> D2H) equals the device-pointer API."""
    import ctypes as C
    from mujoco_mjx_lab_b200 import _lib
    v_reset, v_step = env[8], env[9]
    L, h = _lib.lib(), v_step.sys.handle
    n = 300
    keys = helpers.ppo_keys(5, n)
    obs_h = np.zeros((n, 54), dtype=np.float32)
    _lib.check(L.mjxb_reset_host(h, n, keys.ctypes.data, obs_h.ctypes.data))
    state, obs = v_reset(keys)
    np.testing.assert_array_equal(obs_h, obs.cpu().numpy())
    act = np.random.default_rng(0).normal(size=(n, 21)).astype(np.float32)
    rk = helpers.ppo_keys(6, n)
    r_h, te_h, tr_h = (np.zeros(n, dtype=np.float32) for _ in range(3))
    for _ in range(3):
        _lib.check(L.mjxb_step_autoreset_host(h, n, act.ctypes.data, rk.ctypes.data, obs_h.ctypes.data, r_h.ctypes.data,
                                              te_h.ctypes.data, tr_h.ctypes.data))
        state, obs, r, te, tr = v_step.autoreset(state, torch.from_numpy(act).cuda(), rk)
    np.testing.assert_array_equal(obs_h, obs.cpu().numpy())
    np.testing.assert_array_equal(r_h, r.cpu().numpy())
    qpos_h = np.zeros((n, 28), dtype=np.float32)
    _lib.check(L.mjxb_state_get_host(h, n, qpos_h.ctypes.data, None, None, None, None))
    np.testing.assert_array_equal(qpos_h, state[0].qpos.cpu().numpy())
    assert L.mjxb_step_host(h, n + 1, act.ctypes.data, obs_h.ctypes.data, r_h.ctypes.data, te_h.ctypes.data, tr_h.ctypes.data) == -1


@pytest.mark.parametrize("n", [20000, 5000])              # 5 input chunks of 4096 envs (work-sorted, deferred resets) / 2 chunks, reset warps
def test_host_buffer_api_pinned_direct(model, env, n):
    """Pinned caller buffers take the direct pipeline (one launch, inputs behind per-chunk ready flags, obs stored straight into the
    caller's mapped buffer): same bits as the device-pointer API, over several input chunks and several calls (flag epochs)."""
    from mujoco_mjx_lab_b200 import _lib
    v_reset, v_step = env[8], env[9]
    L, h = _lib.lib(), v_step.sys.handle
    keys = helpers.ppo_keys(9, n)
    pin = lambda *shape, dt=torch.float32: torch.zeros(*shape, dtype=dt).pin_memory()
    obs_h, r_h, te_h, tr_h, act_h = pin(n, 54), pin(n), pin(n), pin(n), pin(n, 21)
    keys_h = torch.from_numpy(keys.view(np.int32).copy()).pin_memory()
    _lib.check(L.mjxb_reset_host(h, n, keys_h.data_ptr(), obs_h.data_ptr()))
    state, obs = v_reset(keys)
    assert torch.equal(obs_h, obs.cpu())
    rng = np.random.default_rng(3)
    for it in range(4):
        act_h.copy_(torch.from_numpy(np.clip(rng.normal(size=(n, 21)), -1, 1).astype(np.float32)))
        rk = helpers.ppo_keys(20 + it, n)
        keys_h.copy_(torch.from_numpy(rk.view(np.int32).copy()))
        _lib.check(L.mjxb_step_autoreset_host(h, n, act_h.data_ptr(), keys_h.data_ptr(), obs_h.data_ptr(), r_h.data_ptr(),
                                              te_h.data_ptr(), tr_h.data_ptr()))
        state, obs, r, te, tr = v_step.autoreset(state, act_h.cuda(), rk)
        assert torch.equal(obs_h, obs.cpu()) and torch.equal(r_h, r.cpu())
        assert torch.equal(te_h, te.cpu()) and torch.equal(tr_h, tr.cpu())
    qpos_h = pin(n, 28)
    _lib.check(L.mjxb_state_get_host(h, n, qpos_h.data_ptr(), None, None, None, None))
    assert torch.equal(qpos_h, state[0].qpos.cpu())


def test_host_step_without_autoreset_and_odd_sizes(model, env):
    """mjxb_step_host (no auto-reset) through pinned buffers, at sizes around one launch wave and a single env."""
    from mujoco_mjx_lab_b200 import _lib
    v_reset, v_step = env[8], env[9]
    L, h = _lib.lib(), v_step.sys.handle
    for n in (1, 2369, 4097):
        keys = helpers.ppo_keys(3, n)
        pin = lambda *shape: torch.zeros(*shape, dtype=torch.float32).pin_memory()
        obs_h, r_h, te_h, tr_h, act_h = pin(n, 54), pin(n), pin(n), pin(n), pin(n, 21)
        keys_h = torch.from_numpy(keys.view(np.int32).copy()).pin_memory()
        _lib.check(L.mjxb_reset_host(h, n, keys_h.data_ptr(), obs_h.data_ptr()))
        state, obs = v_reset(keys)
        act_h.copy_(torch.from_numpy(np.random.default_rng(n).uniform(-1, 1, (n, 21)).astype(np.float32)))
        for _ in range(2):
            _lib.check(L.mjxb_step_host(h, n, act_h.data_ptr(), obs_h.data_ptr(), r_h.data_ptr(), te_h.data_ptr(), tr_h.data_ptr()))
            state, obs, r, te, tr = v_step(state, act_h.cuda())
        assert torch.equal(obs_h, obs.cpu()) and torch.equal(r_h, r.cpu()) and torch.equal(te_h, te.cpu()) and torch.equal(tr_h, tr.cpu())


def test_step_into_caller_buffers(env):
    """v_step.autoreset(..., out=(obs, reward, terminated, truncated)) writes the same bits into caller-owned buffers (rollout slices)."""
    v_reset, v_step = env[8], env[9]
    n = 700
    state, obs0 = v_reset(helpers.ppo_keys(11, n))
    act = torch.rand(n, 21, device="cuda") * 2 - 1
    rk = helpers.ppo_keys(12, n)
    s1, o1, r1, te1, tr1 = v_step.autoreset(state, act, rk)
    traj = torch.zeros(3, 4, n, device="cuda")
    obs_buf = torch.zeros(n, 54, device="cuda")
    s2, o2, r2, te2, tr2 = v_step.autoreset(state, act, rk, out=(obs_buf, traj[0, 1], traj[1, 1], traj[2, 1]))
    assert o2.data_ptr() == obs_buf.data_ptr() and r2.data_ptr() == traj[0, 1].data_ptr()
    assert torch.equal(o1, obs_buf) and torch.equal(r1, traj[0, 1]) and torch.equal(te1, traj[1, 1]) and torch.equal(tr1, traj[2, 1])
    assert torch.equal(s1[0].qpos, s2[0].qpos)
    with pytest.raises(ValueError):
        v_step.autoreset(state, act, rk, out=(obs_buf, traj[0, :, 0], traj[1, 1], traj[2, 1]))      # non-contiguous slice


def test_argument_errors(env):
    from mujoco_mjx_lab_b200 import _lib
    v_reset, v_step = env[8], env[9]
    state, obs = v_reset(helpers.ppo_keys(0, 4))
    with pytest.raises(ValueError):
        v_step(state, torch.zeros(4, 20, device="cuda"))
    with pytest.raises(TypeError):
        v_step(state, torch.zeros(4, 21, device="cuda", dtype=torch.float64))
    with pytest.raises(ValueError):
        v_reset(np.zeros((4, 3), dtype=np.uint32))
    L = _lib.lib()
    assert L.mjxb_speed_test(v_step.sys.handle, 0, None, None, 1, None) == -1


def test_full_size_properties(model, env):
    """BASELINE configs[1] largest size (262144 envs): size-independent properties instead of an oracle run."""
    v_reset, v_step = env[8], env[9]
    n = 262144
    keys = helpers.ppo_keys(42, n)
    state, obs = v_reset(keys)
    state_b, obs_b = v_reset(keys)
    assert torch.equal(obs, obs_b) and torch.equal(state[0].qacc_warmstart, state_b[0].qacc_warmstart)   # deterministic
    g = torch.Generator(device="cuda").manual_seed(1)
    perm = torch.randperm(n, device="cuda", generator=g)
    for t in range(24):
        act = torch.randn(n, 21, device="cuda", generator=g)
        rk = torch.randint(-2 ** 31, 2 ** 31 - 1, (n, 2), device="cuda", dtype=torch.int32, generator=g)
        if t == 23:   # permutation equivariance: envs are independent, order must not matter (bitwise)
            (dp, auxp), obsp, rp, tep, trp = v_step.autoreset(
                (mjx.Data(state[0].qpos[perm], state[0].qvel[perm], state[0].qacc_warmstart[perm], state[0].time[perm]), state[1][perm]),
                act[perm], rk[perm])
        state, obs, r, te, tr = v_step.autoreset(state, act, rk)
    assert torch.equal(obsp, obs[perm]) and torch.equal(rp, r[perm]) and torch.equal(dp.qpos, state[0].qpos[perm])
    d, aux = state
    assert torch.isfinite(obs).all() and torch.isfinite(r).all() and torch.isfinite(d.qvel).all()
    assert (torch.linalg.norm(d.qpos[:, 3:7], dim=1) - 1).abs().max() < 1e-5
    assert ((te == 0) | (te == 1)).all() and ((tr == 0) | (tr == 1)).all()
    done = torch.maximum(te, tr) > 0
    assert (d.time[done] == 0).all() and (aux[done, 8] == 0).all()                   # auto-reset envs restart their clocks
    assert torch.allclose(d.time[~done], aux[~done, 8] * 0.005, atol=1e-5)
    assert (aux[:, 8] <= 24).all() and done.float().mean() < 0.2
    assert ((aux[:, 0] == 0) | (aux[:, 0] == 1)).all() and 0.45 < aux[:, 0].mean() < 0.55
