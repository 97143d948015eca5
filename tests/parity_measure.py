"""Parity measurements shared by tests/test_gpu_exact.py and tools/parity_table.py (which commits them as profiles/r2_parity.json).

Four evaluations of the same one-step problem are compared on identical seeded inputs:
    o64   CPU oracle, float64            (accuracy reference)
    o32   CPU oracle, float32            (MJX's arithmetic restated: sequential sums, MJX's bracketed line search)
    exact libmjxb_exact.so               (GPU, IEEE arithmetic without FMA contraction, MJX's bracketed line search)
    fast  libmjxb.so                     (GPU product: --use_fast_math + closed-form exact line search)
    fast_it libmjxb.so + MJXB_FLAG_LS_ITERATIVE (GPU product arithmetic, MJX's bracketed line search): isolates fast-math
so that `exact vs o32` measures what the warp formulation (summation order, candidate-row compaction, tree-ordered Cholesky) changes,
`fast_it vs exact` what --use_fast_math changes, and `fast vs fast_it` what the exact line search changes.

SURVEY.md section 7's stated one-step float32 tolerances:
    |dqpos|, |dqvel| <= 1e-5 + 1e-4 |ref|;  |dqacc| <= 1e-3 + 1e-3 |ref|;  efc_force <= 1e-3 max(1, |ref|)
    integer outputs (candidate / active masks, solver_niter) bit-exact except where the deciding float is at its threshold (counted).
"""
from __future__ import annotations

import numpy as np
import torch

import helpers
from mujoco_mjx_lab_b200 import _lib, mjx

KINDS = ["free", "stand", "lean", "tumble"]


def T(a):
    return torch.tensor(np.asarray(a), dtype=torch.float32, device="cuda")


def N(t):
    return t.detach().double().cpu().numpy()


TOL = {
    "qpos": lambda ref: 1e-5 + 1e-4 * np.abs(ref),
    "qvel": lambda ref: 1e-5 + 1e-4 * np.abs(ref),
    "qacc": lambda ref: 1e-3 + 1e-3 * np.abs(ref),
    "efc_force": lambda ref: 1e-3 * np.maximum(1.0, np.abs(ref)),
}


def systems(model):
    """The three GPU evaluations (device models)."""
    return {"fast": mjx.put_model(model), "fast_it": mjx.put_model(model, flags=_lib.FLAG_LS_ITERATIVE),
            "exact": mjx.put_model(model, variant="exact")}


def gpu_step(sysm, q, v, w, c):
    """One physics step with the stage outputs of its forward pass. Returns dict of float64 / int arrays."""
    n = q.shape[0]
    nd, out = mjx.step(sysm, mjx.Data(T(q), T(v), T(w), torch.zeros(n, device="cuda"), T(c)), debug=True)
    g = {k: (t.cpu().numpy() if t.dtype == torch.int32 else N(t)) for k, t in out.items()}
    g.update(qpos=N(nd.qpos), qvel=N(nd.qvel), qacc_warmstart=N(nd.qacc_warmstart))
    return g


def oracle_step(oracle, q, v, w, c, prec):
    return oracle.physics_step(q, v, w, None, c, prec=prec, debug=True)


def _stats(err, tol):
    """abs-error statistics and the counted exceedances of the stated tolerance."""
    err, tol = np.asarray(err), np.asarray(tol)
    ratio = err / tol
    env_bad = (ratio > 1.0).reshape(err.shape[0], -1).any(axis=1)
    return dict(median=float(np.median(err)), p99=float(np.percentile(err, 99)), max=float(err.max()),
                median_over_tol=float(np.median(ratio)), p99_over_tol=float(np.percentile(ratio, 99)), max_over_tol=float(ratio.max()),
                n=int(err.size), n_exceed=int((ratio > 1.0).sum()), n_env=int(err.shape[0]), n_env_exceed=int(env_bad.sum()))


def compare(a, b, ref64=None):
    """`a` against `b` (b is the reference of the pair): float outputs against the stated tolerances, integer outputs by counts.
    Mask mismatches are split into 'at threshold' (the deciding float of the float64 oracle is within `near` of zero relative to the
    env's scale) and 'away from threshold'."""
    out = {}
    for name in ("qpos", "qvel", "qacc", "efc_force"):
        out[name] = _stats(np.abs(a[name] - b[name]), TOL[name](b[name]))
    cand_a, cand_b = a["efc_active"] & 1, b["efc_active"] & 1
    act_a, act_b = a["efc_active"] >> 1, b["efc_active"] >> 1
    cm, am = cand_a != cand_b, (act_a != act_b) & (cand_a == cand_b)
    out["candidate_mask"] = dict(n=int(cm.size), mismatch=int(cm.sum()))
    out["active_mask"] = dict(n=int(am.size), mismatch=int(am.sum()), n_active_ref=int(act_b.sum()))
    if ref64 is not None:
        # the float that decides a row's activity at the solution is Jaref = J qacc - aref (force = -D Jaref): a mismatch is 'at
        # threshold' when that force is below 1e-3 of the env's largest constraint force in the float64 oracle
        jar = (ref64["efc_J"] @ ref64["qacc"][:, :, None])[:, :, 0] - ref64["efc_aref"]
        f_would = np.abs(ref64["efc_D"] * jar)
        fmax = np.maximum(ref64["efc_force"].max(axis=1, keepdims=True), 1.0)
        near = f_would < 1e-3 * fmax
        out["active_mask"]["mismatch_at_threshold"] = int((am & near).sum())
        out["active_mask"]["mismatch_away"] = int((am & ~near).sum())
        pos = ref64.get("row_pos")
        if pos is not None:
            nearc = np.abs(pos) < 1e-6
            out["candidate_mask"]["mismatch_at_threshold"] = int((cm & nearc).sum())
            out["candidate_mask"]["mismatch_away"] = int((cm & ~nearc).sum())
    ni_a, ni_b = a["solver_niter"], b["solver_niter"]
    out["solver_niter"] = dict(n=int(ni_a.size), equal=int((ni_a == ni_b).sum()), mean_a=float(ni_a.mean()), mean_b=float(ni_b.mean()),
                               max_abs_diff=int(np.abs(ni_a - ni_b).max()))
    sa, sb = a["sensordata"] > 0, b["sensordata"] > 0
    out["touch_sign"] = dict(n=int(sa.size), mismatch=int((sa != sb).sum()))
    return out


def row_positions(model, ref):
    """Unmasked float64 constraint position of every static row (joint limits, tendon limits, contact slots)."""
    n = ref["con_dist"].shape[0]
    nl = model["nlimit"] + model["ntlimit"]
    rows = np.ones((n, model["nefc"]))
    rows[:, :nl] = np.where((ref["efc_active"][:, :nl] & 1) == 1, ref["efc_pos"][:, :nl], 1.0)
    for p in model["pairs"]:
        for e in range(p["ncon"]):
            d = ref["con_dist"][:, p["con_adr"] + e]
            if p["condim"] == 1:
                rows[:, p["efc_adr"] + e] = d
            else:
                rows[:, p["efc_adr"] + 4 * e: p["efc_adr"] + 4 * e + 4] = d[:, None]
    return rows


def one_step_table(model, oracle, syss, kind, n=512, seed=None):
    """All pairwise comparisons for one state family."""
    q, v, w, c = helpers.make_states(model, n, 700 + KINDS.index(kind) if seed is None else seed, kind)
    o64, o32 = oracle_step(oracle, q, v, w, c, "f64"), oracle_step(oracle, q, v, w, c, "f32")
    o64["row_pos"] = row_positions(model, o64)
    g = {k: gpu_step(s, q, v, w, c) for k, s in syss.items()}
    pairs = {"exact_vs_o32": (g["exact"], o32), "fast_it_vs_exact": (g["fast_it"], g["exact"]), "fast_vs_fast_it": (g["fast"], g["fast_it"]),
             "fast_vs_exact": (g["fast"], g["exact"]), "fast_vs_o32": (g["fast"], o32), "o32_vs_o64": (o32, o64),
             "exact_vs_o64": (g["exact"], o64), "fast_vs_o64": (g["fast"], o64)}
    table = {name: compare(a, b, o64) for name, (a, b) in pairs.items()}
    table["mean_candidate_rows"] = float((o64["efc_active"] & 1).sum(1).mean())
    table["mean_active_rows"] = float((o64["efc_active"] >> 1).sum(1).mean())
    return table


def resync_table(model, oracle, envs, n=64, steps=128):
    """128 consecutive oracle (float32) env states, each advanced one env step by every GPU evaluation; at step 1 and step 128 the
    contact forces / active masks of the forward pass are compared as well (north_star: 'after 1 and 128 steps')."""
    rng = np.random.default_rng(0)
    st, _ = oracle.env_reset(helpers.ppo_keys(1, n), prec="f32")
    acc = {k: dict(qpos=[], qvel=[], obs=[], reward=[], term_mismatch=0, trunc_mismatch=0, stance_mismatch=0, near_height=0, n=0) for k in envs}
    force_cmp = {}
    for t in range(steps):
        act = rng.normal(size=(n, 21))
        rk = helpers.ppo_keys(1000 + t, n)
        s_in = {k: v.copy() for k, v in st.items()}
        st, o_ref, r_ref, te_ref, tr_ref, mask, dbg32 = oracle.env_step(st, act, prec="f32", reset_keys=rk, debug=True)
        _, _, _, _, _, _, dbg64 = oracle.env_step({k: v.copy() for k, v in s_in.items()}, act, prec="f64", reset_keys=rk, debug=True)
        near = np.abs(dbg64["xpos"][:, 4, 2] - 0.7) < 1e-5
        for k, (v_step, sysm) in envs.items():
            d = mjx.Data(T(s_in["qpos"]), T(s_in["qvel"]), T(s_in["qacc_warmstart"]), T(s_in["time"]))
            (d2, aux2), obs, rew, te, tr = v_step.autoreset((d, T(s_in["aux"])), T(act), rk)
            te_g, tr_g = N(te), N(tr)
            a = acc[k]
            a["term_mismatch"] += int(((te_g != te_ref) & ~near).sum())
            a["near_height"] += int(((te_g != te_ref) & near).sum())
            a["trunc_mismatch"] += int((tr_g != tr_ref).sum())
            same = (te_g == te_ref) & (np.maximum(te_ref, tr_ref) == 0)
            a["n"] += int(same.sum())
            a["stance_mismatch"] += int((N(aux2)[same][:, 5] != st["aux"][same][:, 5]).sum())
            a["qpos"].append((np.abs(N(d2.qpos) - st["qpos"]) / TOL["qpos"](st["qpos"]))[same].max(axis=1))
            a["qvel"].append((np.abs(N(d2.qvel) - st["qvel"]) / TOL["qvel"](st["qvel"]))[same].max(axis=1))
            a["obs"].append(np.abs(N(obs) - o_ref)[same].max(axis=1))
            a["reward"].append(np.abs(N(rew) - r_ref)[same])
            if t in (0, steps - 1):   # forward-pass contact forces on the same input state (flip/clip applied like single_step does)
                flip = s_in["aux"][:, 0] > 0.5
                cfgc = sysm.env_cfg_c
                ap = np.array(cfgc.act_perm[:21]); asg = np.array(cfgc.act_sign[:21])
                ctrl = np.clip(np.where(flip[:, None], act[:, ap] * asg, act), -1, 1)
                g = gpu_step(sysm, s_in["qpos"], s_in["qvel"], s_in["qacc_warmstart"], ctrl)
                dbg64["row_pos"] = row_positions(model, dbg64)
                force_cmp[f"{k}_step{t + 1}_vs_o32"] = {kk: vv for kk, vv in compare(g, dict(dbg32, qpos=g["qpos"], qvel=g["qvel"]), dbg64).items()
                                                        if kk in ("efc_force", "qacc", "active_mask", "candidate_mask", "solver_niter", "touch_sign")}
    out = {}
    for k, a in acc.items():
        cat = {f: np.concatenate(a[f]) for f in ("qpos", "qvel", "obs", "reward")}
        out[k] = dict(env_steps=a["n"], terminated_mismatch_away_from_threshold=a["term_mismatch"], terminated_mismatch_at_threshold=a["near_height"],
                      truncated_mismatch=a["trunc_mismatch"], stance_state_mismatch=a["stance_mismatch"],
                      qpos_over_tol=dict(median=float(np.median(cat["qpos"])), p99=float(np.percentile(cat["qpos"], 99)), max=float(cat["qpos"].max()),
                                         n_exceed=int((cat["qpos"] > 1).sum())),
                      qvel_over_tol=dict(median=float(np.median(cat["qvel"])), p99=float(np.percentile(cat["qvel"], 99)), max=float(cat["qvel"].max()),
                                         n_exceed=int((cat["qvel"] > 1).sum())),
                      obs_abs=dict(median=float(np.median(cat["obs"])), p99=float(np.percentile(cat["obs"], 99)), max=float(cat["obs"].max())),
                      reward_abs=dict(median=float(np.median(cat["reward"])), p99=float(np.percentile(cat["reward"], 99)), max=float(cat["reward"].max())))
    out["forces_at_step_1_and_128"] = force_cmp
    return out
