"""MJCF compiler anchors and blob plumbing."""
import os

import numpy as np
import pytest

import helpers
from mujoco_mjx_lab_b200 import modelc

REF_XML = "/root/reference/models"


def test_counts_and_mass(model):
    assert (model["nq"], model["nv"], model["nu"], model["nbody"], model["njnt"], model["ngeom"]) == (28, 27, 21, 17, 22, 20)
    assert (model["nsite"], model["ntendon"], model["nsensor"]) == (2, 2, 2)
    assert (model["npair"], model["ncon"], model["nefc"], model["nlimit"], model["ntlimit"]) == (108, 116, 187, 21, 2)
    assert abs(model["body_mass"].sum() - 40.84402) < 1e-4
    np.testing.assert_allclose(model["body_mass"][[1, 2, 4, 5, 6, 7]], [5.85383, 3.05363, 6.61619, 4.75175, 2.75570, 1.13114], atol=1e-4)


def test_pair_filter(model):
    kinds = [(p["kind"], p["condim"]) for p in model["pairs"]]
    assert kinds.count((modelc.PAIR_PLANE_CAPSULE, 3)) == 8
    assert kinds.count((modelc.PAIR_CAPSULE_CAPSULE, 1)) == 76
    assert kinds.count((modelc.PAIR_SPHERE_CAPSULE, 1)) == 24
    floor_pairs = [model["geom_name"][p["g2"]] for p in model["pairs"] if p["g1"] == 0]
    assert sorted(floor_pairs) == sorted(f"{g}_{s}" for g in ("thigh", "shin", "foot1", "foot2") for s in ("right", "left"))
    # excluded / parent-child pairs never appear
    names = {(model["geom_name"][p["g1"]], model["geom_name"][p["g2"]]) for p in model["pairs"]}
    assert ("waist_lower", "thigh_right") not in names and ("butt", "thigh_left") not in names
    # contact slots: condim-1 first, then two per plane-capsule pair; rows: 21 + 2 + 100 + 16*4
    assert [p["con_adr"] for p in model["pairs"]][:3] == [0, 1, 2] and model["ncon1"] == 100
    assert model["pairs"][-1]["efc_adr"] == 23 + 100 + 4 * 14


def test_contact_parameter_mixing(model):
    floor = [p for p in model["pairs"] if p["g1"] == 0][0]
    np.testing.assert_allclose(floor["solref"], [0.0175, 1.0])
    np.testing.assert_allclose(floor["solimp"], [0.9, 0.97, 0.002, 0.5, 2.0])
    assert floor["mu"] == 1.0
    body = [p for p in model["pairs"] if p["g1"] != 0][0]
    np.testing.assert_allclose(body["solref"], [0.015, 1.0])
    np.testing.assert_allclose(body["solimp"], [0.9, 0.99, 0.003, 0.5, 2.0])


def test_pose_anchors(model):
    x = model["xpos0"]
    assert abs(x[4, 2] - 0.857) < 1e-12 and abs(x[2, 2] - 1.472) < 1e-12 and abs(x[7, 2] - 0.027) < 1e-12
    assert model["body_name"][4] == "pelvis" and model["body_name"][2] == "head"
    assert model["sensor_name"] == ["touch_foot_right", "touch_foot_left"]


def test_joint_tables(model):
    rng = {n: r for n, r in zip(model["jnt_name"], np.degrees(model["jnt_range"]))}
    np.testing.assert_allclose(rng["knee_right"], [-160, 2])
    np.testing.assert_allclose(rng["hip_y_left"], [-150, 20])
    np.testing.assert_allclose(rng["elbow_left"], [-100, 50])
    ax = dict(zip(model["jnt_name"], model["jnt_axis"]))
    np.testing.assert_allclose(ax["ankle_x_right"], np.array([1, 0, .5]) / np.linalg.norm([1, 0, .5]))
    np.testing.assert_allclose(model["act_gear"], [40, 40, 40, 40, 40, 120, 80, 20, 20, 40, 40, 120, 80, 20, 20, 20, 20, 40, 20, 20, 40])
    assert list(model["act_dof"]) == list(range(6, 27))
    st = dict(zip(model["jnt_name"], model["jnt_stiffness"]))
    assert st["abdomen_z"] == 20 and st["knee_left"] == 1 and st["elbow_right"] == 0 and st["ankle_y_left"] == 6
    # structural non-zeros of the lower triangle of M: 243
    nnz = sum(len(chain(model, d)) for d in range(27))
    assert nnz == 243


def chain(model, d):
    out = []
    while d >= 0:
        out.append(d)
        d = model["dof_parent"][d]
    return out


def test_humanoid_xml_counts():
    m = helpers.load("humanoid")
    assert (m["npair"], m["ncon"], m["nefc"]) == (159, 175, 303)      # SURVEY A.4 last paragraph
    assert m["opt"]["integrator"] == modelc.INT_EULER and m["opt"]["iterations"] == 100 and m["opt"]["ls_iterations"] == 50


def test_blob_roundtrip(model):
    blob = modelc.pack_blob(model)
    again = modelc.blob_from_json(modelc.blob_to_json(blob))
    assert blob.tobytes() == again.tobytes()
    header = open(os.path.join(helpers.ROOT, "include", "mjxb_model.h")).read()
    assert header == modelc.emit_c_header(), "include/mjxb_model.h is stale: regenerate with modelc.emit_c_header()"


@pytest.mark.skipif(not os.path.isdir(REF_XML), reason="reference checkout absent (GPU box)")
@pytest.mark.parametrize("name", ["humanoid_mjx", "humanoid"])
def test_builtin_json_matches_reference_xml(name):
    fresh = modelc.compile_mjcf(os.path.join(REF_XML, f"{name}.xml"))
    assert modelc.pack_blob(fresh).tobytes() == modelc.pack_blob(helpers.load(name)).tobytes()


def test_setconst_invweights(model):
    """dof_invweight0 / body_invweight0 / tendon_invweight0 recomputed independently from M(qpos0)."""
    M, (_, _, xmat, xipos, xanchor, xaxis, jacp, jacr) = modelc.np_mass_matrix(model, model["qpos0"])
    Minv = np.linalg.inv(M)
    np.testing.assert_allclose(model["dof_invweight0"][6:], np.diag(Minv)[6:], rtol=1e-10)
    np.testing.assert_allclose(model["dof_invweight0"][0], np.diag(Minv)[:3].mean(), rtol=1e-10)
    np.testing.assert_allclose(model["body_invweight0"][7, 0], np.trace(jacp[7] @ Minv @ jacp[7].T) / 3, rtol=1e-10)
    np.testing.assert_allclose(model["meaninertia"], np.trace(M) / 27, rtol=1e-12)
    jt = np.zeros(27); jt[11], jt[12] = 0.5, -0.5
    np.testing.assert_allclose(model["tendons"][0]["invweight0"], jt @ Minv @ jt, rtol=1e-10)
