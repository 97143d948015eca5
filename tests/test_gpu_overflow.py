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
