"""Breaks the common mode between the product and the oracle: both are fed modelc.pack_blob(), so a wrong inertia / invweight / solimp in
modelc.py would be invisible to every GPU-vs-oracle parity test.  Two derivations that share NO code with modelc.py:

  1. body mass, centre of mass and inertia tensor by numerical quadrature over the geoms' volumes, from the geometry numbers of
     /root/reference/models/humanoid_mjx.xml transcribed BY HAND below (fromto / pos / radius; density 1000, MuJoCo sums overlapping
     geoms without subtracting the overlap) -- modelc.py uses closed-form capsule / sphere formulas and an MJCF parser;
  2. dof_invweight0 / body_invweight0 / tendon_invweight0 / meaninertia from the ORACLE's own pipeline at qpos0 (kinematics -> crb mass
     matrix -> Cholesky; oracle_capi.cpp orc_set_const) -- modelc.py builds M from per-body point Jacobians in numpy;
  3. the contact-row invweight and solref / solimp mixing rules restated from the MuJoCo documentation on the pairs' hand-known values.
"""
import ctypes as C

import numpy as np

from mujoco_mjx_lab_b200 import modelc
from oracle import oracle as O

# body -> [(kind, p0, p1 | None, radius)] in the BODY frame, read off reference models/humanoid_mjx.xml:107-190 and its default classes
# (:44-68: thigh .06, shin fromto "0 0 0 0 0 -.3" .049, foot .027 with foot1 / foot2 fromto, arm_upper .04, arm_lower .031, hand sphere .04)
GEOMS = {
    "torso": [("capsule", (0, -.07, 0), (0, .07, 0), .07), ("capsule", (-.01, -.06, -.12), (-.01, .06, -.12), .06)],
    "head": [("sphere", (0, 0, 0), None, .09)],
    "waist_lower": [("capsule", (0, -.06, 0), (0, .06, 0), .06)],
    "pelvis": [("capsule", (-.02, -.07, 0), (-.02, .07, 0), .09)],
    "thigh_right": [("capsule", (0, 0, 0), (0, .01, -.34), .06)],
    "shin_right": [("capsule", (0, 0, 0), (0, 0, -.3), .049)],
    "foot_right": [("capsule", (-.07, -.01, 0), (.14, -.03, 0), .027), ("capsule", (-.07, .01, 0), (.14, .03, 0), .027)],
    "thigh_left": [("capsule", (0, 0, 0), (0, -.01, -.34), .06)],
    "shin_left": [("capsule", (0, 0, 0), (0, 0, -.3), .049)],
    "foot_left": [("capsule", (-.07, -.01, 0), (.14, -.03, 0), .027), ("capsule", (-.07, .01, 0), (.14, .03, 0), .027)],
    "upper_arm_right": [("capsule", (0, 0, 0), (.16, -.16, -.16), .04)],
    "lower_arm_right": [("capsule", (.01, .01, .01), (.17, .17, .17), .031)],
    "hand_right": [("sphere", (0, 0, 0), None, .04)],
    "upper_arm_left": [("capsule", (0, 0, 0), (.16, .16, -.16), .04)],
    "lower_arm_left": [("capsule", (.01, -.01, .01), (.17, -.17, .17), .031)],
    "hand_left": [("sphere", (0, 0, 0), None, .04)],
}
DENSITY = 1000.0


def _quadrature(geom, ngrid=120):
    """mass, first moment, second moment (about the body origin) of one geom by midpoint quadrature on a grid over its bounding box."""
    kind, p0, p1, r = geom
    p0 = np.array(p0, float)
    p1 = p0 if p1 is None else np.array(p1, float)
    lo, hi = np.minimum(p0, p1) - r, np.maximum(p0, p1) + r
    axes = [lo[k] + (np.arange(ngrid) + 0.5) * (hi[k] - lo[k]) / ngrid for k in range(3)]
    dv = np.prod((hi - lo) / ngrid)
    X, Y, Z = np.meshgrid(*axes, indexing="ij")
    P = np.stack([X.ravel(), Y.ravel(), Z.ravel()], 1)
    ab = p1 - p0
    t = np.zeros(len(P)) if kind == "sphere" else np.clip((P - p0) @ ab / (ab @ ab), 0, 1)
    inside = np.linalg.norm(P - (p0 + t[:, None] * ab), axis=1) <= r
    P = P[inside]
    m = DENSITY * dv * len(P)
    return m, DENSITY * dv * P.sum(0), DENSITY * dv * (P.T @ P)


def test_body_inertia_by_quadrature(model):
    for name, geoms in GEOMS.items():
        b = model["body_name"].index(name)
        m, s1, s2 = 0.0, np.zeros(3), np.zeros((3, 3))
        for g in geoms:
            gm, gs1, gs2 = _quadrature(g)
            m, s1, s2 = m + gm, s1 + gs1, s2 + gs2
        com = s1 / m
        # inertia about the centre of mass, body axes:  I = tr(S) E - S  with S the central second moment
        sc = s2 - m * np.outer(com, com)
        inertia = np.trace(sc) * np.eye(3) - sc
        assert abs(model["body_mass"][b] - m) < 4e-3 * m, (name, model["body_mass"][b], m)
        np.testing.assert_allclose(model["body_ipos"][b], com, atol=2e-4, err_msg=name)
        # modelc stores the tensor about ipos in the body frame (possibly rotated into principal axes by body_iquat): compare invariants
        # (the eigenvalues) and, via the world-frame tensor at qpos0 orientation-free form R I R^T, the tensor itself when no iquat is used
        got = np.asarray(model["body_inertia"][b], float)
        if got.shape == (3,):
            got = np.diag(got)
        ev_ref, ev_got = np.sort(np.linalg.eigvalsh(inertia)), np.sort(np.linalg.eigvalsh(got))
        np.testing.assert_allclose(ev_got, ev_ref, rtol=1.5e-2, atol=1e-6, err_msg=name)   # grid quadrature: ~1 % on thin capsules
        if "body_iquat" not in model or np.allclose(np.asarray(model["body_iquat"])[b], [1, 0, 0, 0]):
            np.testing.assert_allclose(got, inertia, rtol=0, atol=1.5e-2 * ev_ref.max(), err_msg=name)
    assert abs(sum(_quadrature(g)[0] for gs in GEOMS.values() for g in gs) - 40.844) < 0.1       # SURVEY A.3 total mass


def _oracle_set_const(model):
    O.build()
    blob = modelc.pack_blob(model)
    nv, nb, nt = model["nv"], model["nbody"], model["ntendon"]
    dof, body, ten, mean = np.zeros(nv), np.zeros((nb, 2)), np.zeros(max(nt, 1)), C.c_double(0)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    O.lib().orc_set_const(p(blob), p(dof), p(body), p(ten), C.byref(mean))
    return blob, dof, body, ten[:nt], mean.value


def test_invweights_from_the_oracles_own_mass_matrix(model):
    blob, dof, body, ten, mean = _oracle_set_const(model)
    nv, nb = model["nv"], model["nbody"]
    np.testing.assert_allclose(blob["dof_invweight0"][:nv], dof, rtol=2e-5)
    np.testing.assert_allclose(blob["body_invweight0"][:nb], body, rtol=2e-5, atol=1e-12)
    np.testing.assert_allclose([t["invweight0"] for t in model["tendons"]], ten, rtol=2e-5)
    assert abs(float(blob["meaninertia"]) - mean) < 2e-5 * mean
    # what the kernels read: limit rows use the dof's invweight, tendon rows the tendon's
    for i, j in enumerate(blob["lim_jnt"][: int(blob["nlimit"])]):
        assert abs(float(blob["dof_invweight0"][int(blob["jnt_dofadr"][j])]) - dof[int(blob["jnt_dofadr"][j])]) < 2e-5 * dof[int(blob["jnt_dofadr"][j])]


def test_contact_row_parameters_restated(model):
    """MuJoCo's contact parameter rules restated on their own: friction = max, solref / solimp mixed with weight solmix1/(solmix1+solmix2)
    (= 1/2 here), condim = max; elliptic->pyramidal invweight of a frictional row  (tran1 + tran2) (1 + mu^2) 2 mu^2 / impratio  and of a
    frictionless one  tran1 + tran2  (engine_core_constraint.c mj_makeImpedance / mj_instantiateContact; mjx constraint.py)."""
    _, _, body, _, _ = _oracle_set_const(model)
    floor_sol = dict(solref=(0.02, 1.0), solimp=(0.9, 0.95, 0.001, 0.5, 2.0), mu=1.0)          # MuJoCo defaults on the floor geom
    body_sol = dict(solref=(0.015, 1.0), solimp=(0.9, 0.99, 0.003, 0.5, 2.0), mu=0.7)          # class "body" (xml :44)
    for p in model["pairs"]:
        b1, b2 = model["geom_body"][p["g1"]], model["geom_body"][p["g2"]]
        tran = body[b1, 0] + body[b2, 0]
        if p["g1"] == 0:
            assert p["condim"] == 3
            mu = max(floor_sol["mu"], body_sol["mu"])
            np.testing.assert_allclose(p["solref"], 0.5 * (np.array(floor_sol["solref"]) + np.array(body_sol["solref"])), rtol=1e-6)
            np.testing.assert_allclose(p["solimp"], 0.5 * (np.array(floor_sol["solimp"]) + np.array(body_sol["solimp"])), rtol=1e-6)
            assert abs(p["mu"] - mu) < 1e-7
            want = tran * (1.0 + mu * mu) * 2.0 * mu * mu / 1.0
        else:
            assert p["condim"] == 1
            np.testing.assert_allclose(p["solref"], body_sol["solref"], rtol=1e-6)
            np.testing.assert_allclose(p["solimp"], body_sol["solimp"], rtol=1e-6)
            want = tran
        assert abs(p["invweight"] - want) < 3e-5 * want, (p["g1"], p["g2"], p["invweight"], want)
