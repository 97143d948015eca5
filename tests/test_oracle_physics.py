"""CPU oracle checks: closed-form stage tests and invariants (the reference ships no golden vectors for mjx.step)."""
import numpy as np
import pytest

import helpers
from mujoco_mjx_lab_b200 import modelc
from oracle import oracle as O


def _bodies(model):
    return range(1, model["nbody"])


def test_mass_matrix_and_gravity_bias_vs_jacobian_formula(model):
    """CRB mass matrix and RNE gravity term against the independent numpy Jacobian construction."""
    orc = helpers.make_oracle(model)
    q, v, w, c = helpers.make_states(model, 8, 3, "free")
    out = orc.forward(q, np.zeros_like(v), prec="f64", debug=True)
    for e in range(8):
        M, (xpos, _, _, _, _, _, jacp, _) = modelc.np_mass_matrix(model, q[e])
        np.testing.assert_allclose(out["qM"][e], M, atol=2e-6)          # constants are f32-rounded in the blob
        np.testing.assert_allclose(out["xpos"][e], xpos, atol=1e-6)
        g = sum(model["body_mass"][b] * 9.81 * jacp[b][2] for b in _bodies(model))
        np.testing.assert_allclose(out["qfrc_bias"][e], g, atol=2e-4)


def test_free_flight_root_acceleration(model):
    """No contacts, zero velocity, springs at rest: the root accelerates at exactly -g."""
    orc = helpers.make_oracle(model)
    q = np.tile(model["qpos0"], (1, 1)); q[0, 2] += 1.0
    out = orc.forward(q, np.zeros((1, 27)), prec="f64", debug=True)
    np.testing.assert_allclose(out["qacc"][0, :3], [0, 0, -9.81], atol=1e-9)
    np.testing.assert_allclose(out["qacc"][0, 3:], 0, atol=1e-8)
    assert (out["efc_active"] == 0).all() and (out["con_dist"] > 0).all()


def test_energy_and_momentum_conservation(model):
    """Coriolis/centrifugal terms: kinetic energy and momentum are conserved without gravity, springs, dampers, limits."""
    blob = modelc.pack_blob(model).copy()
    blob["dof_damping"][:] = 0; blob["dof_stiffness"][:] = 0
    blob["jnt_range"][:, 0] = -100; blob["jnt_range"][:, 1] = 100
    blob["ten_range"][:, 0] = -100; blob["ten_range"][:, 1] = 100
    blob["gravity"][:] = 0
    rng = np.random.default_rng(0)
    q = model["qpos0"].copy(); q[2] += 3.0; q[7:] += rng.uniform(-.5, .5, 21)
    qq = rng.normal(size=4); q[3:7] = qq / np.linalg.norm(qq)
    v = rng.normal(size=27)

    def invariants(qp, qv):
        M, (_, _, xmat, xipos, _, _, jacp, jacr) = modelc.np_mass_matrix(model, qp)
        p = sum(model["body_mass"][b] * jacp[b] @ qv for b in _bodies(model))
        return 0.5 * qv @ M @ qv, p

    drift = []
    for dt in (4e-4, 2e-4):
        blob["timestep"] = dt
        orc = O.Oracle(blob)
        st = orc.physics_step(q[None], v[None], nsteps=int(round(0.1 / dt)), prec="f64")
        e0, p0 = invariants(q, v)
        e1, p1 = invariants(st["qpos"][0], st["qvel"][0])
        drift.append(abs(e1 - e0) / e0)
        assert abs(np.linalg.norm(st["qpos"][0, 3:7]) - 1) < 1e-12
        assert np.abs(p1 - p0).max() / np.abs(p0).max() < 2e-4
    assert drift[0] < 2e-4 and drift[1] < 0.7 * drift[0]      # first-order integrator: error shrinks with dt


SPHERE_XML = """<mujoco><option timestep="0.005"/><worldbody>
  <geom name="floor" type="plane" size="0 0 .05"/>
  <body name="ball" pos="0 0 {z}"><freejoint/><geom name="ball" type="sphere" size="0.1"/></body>
</worldbody></mujoco>"""


@pytest.mark.parametrize("pen", [2e-4, 7e-4, 3e-3])
def test_sphere_on_plane_closed_form(tmp_path, pen):
    """Soft-contact known answer: pyramidal 4-row contact, a_z = (4 D aref - m g) / (m + 4 D)  (SURVEY B.7, B.12)."""
    path = tmp_path / "sphere.xml"
    path.write_text(SPHERE_XML.format(z=0.1 - pen))
    m = modelc.compile_mjcf(str(path))
    assert (m["nv"], m["npair"], m["ncon"], m["nefc"]) == (6, 1, 1, 4)
    mass = m["body_mass"][1]
    np.testing.assert_allclose(mass, 4 / 3 * np.pi * 0.1 ** 3 * 1000)
    np.testing.assert_allclose(m["body_invweight0"][1, 0], 1 / mass, rtol=1e-9)
    orc = O.Oracle(modelc.pack_blob(m))
    out = orc.forward(m["qpos0"][None], np.zeros((1, 6)), prec="f64", debug=True)
    # impedance / stiffness per SURVEY B.7
    dmin, dmax, width, mid, power = 0.9, 0.95, 0.001, 0.5, 2.0
    tc, dr = max(0.02, 2 * 0.005), 1.0
    k = 1 / (dmax ** 2 * tc ** 2 * dr ** 2)
    x = pen / width
    y = (x ** power) / mid ** (power - 1) if x < mid else 1 - (1 - x) ** power / (1 - mid) ** (power - 1)
    imp = dmax if x > 1 else dmin + y * (dmax - dmin)
    mu = 1.0
    invw = (1 / mass) * (1 + mu * mu) * 2 * mu * mu
    D = 1 / (invw * (1 - imp) / imp)
    aref = k * imp * pen
    az = (4 * D * aref - mass * 9.81) / (mass + 4 * D)
    np.testing.assert_allclose(out["con_dist"][0, 0], -pen, rtol=1e-5)
    np.testing.assert_allclose(out["efc_D"][0], D, rtol=1e-5)
    np.testing.assert_allclose(out["efc_aref"][0], aref, rtol=1e-5)
    np.testing.assert_allclose(out["qacc"][0, 2], az, rtol=1e-5)
    np.testing.assert_allclose(out["qacc"][0, [0, 1, 3, 4, 5]], 0, atol=1e-7)
    np.testing.assert_allclose(out["efc_force"][0], D * (aref - az), rtol=1e-5)


def test_joint_limit_row_closed_form(model):
    """One violated hinge limit in free flight: row sign, D and aref follow SURVEY B.7."""
    orc = helpers.make_oracle(model)
    q = model["qpos0"][None].copy(); q[0, 2] += 1.0
    j = model["jnt_name"].index("knee_right"); qa = model["jnt_qposadr"][j]
    q[0, qa] = model["jnt_range"][j, 1] + 0.004                      # 4 mm rad past the upper limit
    out = orc.forward(q, np.zeros((1, 27)), prec="f64", debug=True)
    row = list(model["lim_jnts"]).index(j)
    act = out["efc_active"][0]
    assert act[row] & 1 and (act & 1).sum() == 1
    pos = -0.004
    np.testing.assert_allclose(out["efc_pos"][0, row], pos, rtol=1e-4)
    x = 0.004 / 0.01; imp = 1e-4 + (x * x / 0.5) * (0.99 - 1e-4)        # solimplimit "0 .99 .01", x < mid
    D = 1 / (model["dof_invweight0"][model["jnt_dofadr"][j]] * (1 - imp) / imp)
    np.testing.assert_allclose(out["efc_D"][0, row], D, rtol=1e-4)
    np.testing.assert_allclose(out["efc_aref"][0, row], -(1 / (0.99 ** 2 * 0.02 ** 2)) * imp * pos, rtol=1e-4)
    assert out["efc_force"][0, row] > 0


def test_solver_invariants(model):
    """Forces are non-negative, active rows are candidates, f32 and f64 agree, and the solution is a stationary point."""
    orc = helpers.make_oracle(model)
    q, v, w, c = helpers.make_states(model, 64, 5, "lean")
    o64 = orc.forward(q, v, w, c, prec="f64", debug=("efc_force", "efc_active", "qacc", "qM", "efc_J", "efc_D", "efc_aref", "qfrc_smooth",
                                                     "qfrc_constraint", "solver_niter"))
    o32 = orc.forward(q, v, w, c, prec="f32", debug=("efc_force", "efc_active", "qacc"))
    assert (o64["efc_force"] >= 0).all() and (o32["efc_force"] >= 0).all()
    assert (((o64["efc_active"] >> 1) & 1) <= (o64["efc_active"] & 1)).all()
    assert (o64["efc_active"] & 1).sum() > 64 * 8
    # KKT residual of the converged Newton solve: M a - qfrc_smooth - J^T f = 0 with f = -D (J a - aref) on active rows
    for e in range(64):
        a = o64["qacc"][e]
        jar = o64["efc_J"][e] @ a - o64["efc_aref"][e]
        f = np.where(jar < 0, -o64["efc_D"][e] * jar, 0.0)
        res = o64["qM"][e] @ a - o64["qfrc_smooth"][e] - o64["efc_J"][e].T @ f
        assert np.abs(res).max() < 1e-5 * max(1.0, np.abs(o64["qfrc_smooth"][e]).max())
        np.testing.assert_allclose(f, o64["efc_force"][e], atol=1e-7 * max(1.0, f.max()))
    err = np.abs(o32["qacc"] - o64["qacc"]) / np.maximum(1, np.abs(o64["qacc"]))
    assert np.median(err) < 1e-3 and err.max() < 0.2
    assert o64["solver_niter"].max() <= 10


def test_mirror_symmetry(model):
    """Left/right mirrored state -> mirrored accelerations (the model is symmetric; exercises every stage)."""
    orc = helpers.make_oracle(model)
    q, v, w, c = helpers.make_states(model, 16, 11, "lean")
    names = model["jnt_name"]
    # Reflection y -> -y. A hinge with axis a and angle t maps to axis a' = (ax, -ay, az) and angle -t (rotations are pseudo-vectors).
    # Every left joint is declared with axis -a'(right), so right/left pairs swap and KEEP their value;
    # the unpaired abdomen joints map to themselves: abdomen_z (a' = a) and abdomen_x (a' = a) flip sign, abdomen_y (a' = -a) keeps it.
    dof_of = {n: model["jnt_dofadr"][i] for i, n in enumerate(names)}
    perm, sign = np.arange(27), np.ones(27)
    for n, d in dof_of.items():
        if n.endswith("_right"):
            o = dof_of[n[:-6] + "_left"]
            perm[d], perm[o] = o, d
    for n in ("abdomen_z", "abdomen_x"):
        sign[dof_of[n]] = -1
    # root: linear (x, -y, z); angular velocity is a pseudo-vector in the body frame: (-wx, wy, -wz)
    sign[1] = -1; sign[3] = -1; sign[5] = -1
    qm, vm, cm = q.copy(), v[:, perm] * sign, np.zeros_like(c)
    qm[:, 1] *= -1
    qm[:, 3:7] = q[:, 3:7] * np.array([1, -1, 1, -1])            # quaternion of the reflected rotation
    qm[:, 7:] = (q[:, 7:][:, perm[6:] - 6]) * sign[6:]
    cm[:] = (c[:, perm[6:] - 6]) * sign[6:]
    a = orc.forward(q, v, None, c, prec="f64", debug=("qacc", "con_dist"))
    b = orc.forward(qm, vm, None, cm, prec="f64", debug=("qacc", "con_dist"))
    np.testing.assert_allclose(np.sort(a["con_dist"], axis=1), np.sort(b["con_dist"], axis=1), atol=1e-6)
    np.testing.assert_allclose(b["qacc"], a["qacc"][:, perm] * sign, atol=2e-4 * np.abs(a["qacc"]).max())
