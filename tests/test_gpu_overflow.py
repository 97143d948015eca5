"""Envs whose candidate rows exceed the 32-row shared-memory tile are re-run by the mid (64 rows) and big (320 rows) tiers: same results."""
import numpy as np
import pytest
import torch

import helpers
from mujoco_mjx_lab_b200 import mjx

pytestmark = pytest.mark.gpu


def test_overflow_envs_match_oracle(model, oracle):
    sysm = mjx.put_model(model)
    n = 512
    q, v, w, c = helpers.make_states(model, n, 77, "tumble")
    ref = oracle.forward(q, v, w, c, prec="f64", debug=True)
    r32 = oracle.forward(q, v, w, c, prec="f32", debug=True)
    ncand = (ref["efc_active"] & 1).sum(axis=1)
    big = ncand > 32
    assert big.sum() >= 3, "seed no longer produces overflowing envs"
    assert ((ncand > 32) & (ncand <= 60)).sum() >= 3 and (ncand > 66).sum() >= 1, "need envs for both the mid and the big tier"
    t = lambda a: torch.tensor(a, dtype=torch.float32, device="cuda")
    _, out = mjx.forward(sysm, mjx.Data(t(q), t(v), t(w), torch.zeros(n, device="cuda"), t(c)), debug=True)
    status = out["status"].cpu().numpy()
    assert ((status & 2) != 0)[ncand > 34].all() and ((status & 2) == 0)[ncand < 17].all()
    g = out["qacc"].double().cpu().numpy()
    rel = lambda a, b: np.abs(a - b) / np.maximum(1, np.abs(b))
    from test_gpu_parity import assert_f32_equivalent
    eg, e32 = rel(g, ref["qacc"])[big].max(axis=1), rel(r32["qacc"], ref["qacc"])[big].max(axis=1)
    assert_f32_equivalent(eg, e32, 1e-4, "qacc of overflowing envs")
    cand_g = out["efc_active"].cpu().numpy() & 1
    assert ((cand_g != (ref["efc_active"] & 1)).sum(axis=1)[big] <= 1).all()
    # and a second launch (counters were reset by the consuming pass) gives identical bits
    _, out2 = mjx.forward(sysm, mjx.Data(t(q), t(v), t(w), torch.zeros(n, device="cuda"), t(c)), debug=True)
    assert torch.equal(out["qacc"], out2["qacc"])


def test_more_than_32_candidate_contacts(model, oracle):
    """Crumpled, sunk poses: up to ~46 candidate contacts and > 100 rows per env -- the big tier's strip loops over contacts
    (row assignment, touch sensors); candidate set bit-exact, qacc and sensors f32-equivalent."""
    sysm = mjx.put_model(model)
    n = 256
    q, v, w, c = helpers.make_states(model, n, 1, "crumple")
    ref = oracle.forward(q, v, w, c, prec="f64", debug=True)
    r32 = oracle.forward(q, v, w, c, prec="f32", debug=True)
    ncon = (ref["con_dist"] < 0).sum(axis=1)
    many = ncon > 32
    assert many.sum() >= 3, "seed no longer produces envs with more than 32 candidate contacts"
    t = lambda a: torch.tensor(a, dtype=torch.float32, device="cuda")
    _, out = mjx.forward(sysm, mjx.Data(t(q), t(v), t(w), torch.zeros(n, device="cuda"), t(c)), debug=True)
    torch.cuda.synchronize()
    status = out["status"].cpu().numpy()
    assert ((status & 2) != 0)[many].all() and ((status & 1) == 0).all()
    near = np.abs(ref["con_dist"]) < 1e-5                          # sign of a ~0 distance may differ in f32
    assert ((((out["con_dist"].cpu().numpy() < 0) != (ref["con_dist"] < 0)) & ~near).sum(axis=1) == 0).all()
    from test_gpu_parity import assert_f32_equivalent
    rel = lambda a, b: np.abs(a - b) / np.maximum(1, np.abs(b))
    g = out["qacc"].double().cpu().numpy()
    assert_f32_equivalent(rel(g, ref["qacc"]).max(axis=1), rel(r32["qacc"], ref["qacc"]).max(axis=1), 1e-4, "qacc of crumpled envs")
    s = out["sensordata"].double().cpu().numpy()
    assert_f32_equivalent(rel(s, ref["sensordata"]).max(axis=1), rel(r32["sensordata"], ref["sensordata"]).max(axis=1), 1e-4,
                          "touch sensors of crumpled envs")


def test_single_overflow_tier_in_the_latency_regime(model):
    """Auto-resetting batches of up to six rounds run ONE overflow tier (the big one consumes the main tier's list directly: every kernel
    boundary costs ~6 us there, mjxb_abi.cu launch()). Tumbling states that do overflow the 32-row tile -- some beyond the mid tier's 64
    rows -- must come out bit-identical to the three-tier chain (MJXB_SKIP_MID=0 at model creation), over steps with resets."""
    import os
    from mujoco_mjx_lab_b200 import training_utils
    env_a = training_utils.load_model_and_create_env("", helpers.env_config(), model=model)
    os.environ["MJXB_SKIP_MID"] = "0"
    try:
        env_b = training_utils.load_model_and_create_env("", helpers.env_config(), model=model)
    finally:
        del os.environ["MJXB_SKIP_MID"]
    n = 512
    q, v, w, c = helpers.make_states(model, n, 77, "tumble")
    t = lambda a: torch.tensor(a, dtype=torch.float32, device="cuda")
    keys = helpers.ppo_keys(4, n)
    outs = []
    for env in (env_a, env_b):
        (d, aux), _ = env[8](keys)
        d.qpos.copy_(t(q)); d.qvel.copy_(t(v)); d.qacc_warmstart.copy_(t(w))
        state = (d, aux)
        status = torch.zeros(n, dtype=torch.int32, device="cuda")
        launches0 = env[9].sys.lib.mjxb_launch_count()
        rec = []
        for step in range(4):
            act = t(c) * (0.5 + 0.5 * step)
            state, obs, r, te, tr = env[9].autoreset(state, act, helpers.ppo_keys(50 + step, n))
            rec.append((obs.clone(), r.clone(), te.clone(), tr.clone(), state[0].qpos.clone(), state[0].qvel.clone(), state[1].clone()))
        outs.append((rec, env[9].sys.lib.mjxb_launch_count() - launches0))
    (ra, la), (rb, lb) = outs
    assert la == 4 * 2 and lb == 4 * 3                                   # main + big   against   main + mid + big
    for sa, sb in zip(ra, rb):
        for xa, xb in zip(sa, sb):
            assert torch.equal(xa, xb)
    assert float(torch.stack([s[2] for s in ra]).sum()) > 0             # episodes did end (resets inside the overflow tier as well)
    ref = helpers.make_oracle(model, helpers.env_config()) if False else None
    # the states do overflow: the forward pass reports the row spill for a good part of the batch, some beyond 64 rows
    sysm = mjx.put_model(model)
    _, dbg = mjx.forward(sysm, mjx.Data(t(q), t(v), t(w), torch.zeros(n, device="cuda"), t(c)), debug=True)
    ncand = (dbg["efc_active"] & 1).sum(1)
    assert int((ncand > 32).sum()) >= 3 and int((ncand > 64).sum()) >= 1
