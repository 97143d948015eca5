// oracle/oracle.hpp -- CPU restatement of the reference hot path.  TEST INFRASTRUCTURE ONLY.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may use anything
// under oracle/.  The product path (mujoco_mjx_lab_b200 + csrc/) never includes, links or calls this.
//
// PARITY UNPINNED: the arithmetic being restated lives in the un-vendored third-party packages
// mujoco-mjx==3.3.6 / mujoco==3.3.6 (reference requirements.txt:26-27), absent from this image, and the
// reference ships no golden vectors for mjx.step.  This file restates the published
// MuJoCo/MJX algorithm (dense formulation, static 116 contact slots / 187 constraint rows, exactly the shapes
// MJX materialises) and is anchored on the reference's own call sites:
//    src/envs.py:108-113  single_pipeline_init = make_data -> replace(qpos,qvel) -> mjx.forward
//    src/envs.py:115-202  single_reset            src/envs.py:333-492  single_step (mjx.step at :345)
//    mjx_humanoid_speed_test.py:48-57             make_data -> qvel[0]=vel -> mjx.step -> qpos[0]
// tools/dump_mjx_golden.py produces real-MJX vectors on any machine that has mujoco-mjx; tests load them from
// tests/golden/mjx_*.npz when present.
//
// Everything is templated on the scalar type: float mirrors MJX's float32 arithmetic (sequential summation
// order), double is the accuracy reference.  Model constants are always the float32-rounded blob values
// (mjx.put_model casts the model to float32).
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <vector>

#include "mjxb.h"

namespace orc {

constexpr int MAXCON = 256;
constexpr int MAXEFC = 640;
constexpr double MINVAL = 1e-15;  // mjMINVAL
constexpr double MINIMP = 1e-4, MAXIMP = 0.9999;

// ---------------------------------------------------------------- small vector / quaternion helpers
template <class R> inline R dot3(const R* a, const R* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
template <class R> inline void cross3(R* r, const R* a, const R* b) {
  R x = a[1] * b[2] - a[2] * b[1], y = a[2] * b[0] - a[0] * b[2], z = a[0] * b[1] - a[1] * b[0];
  r[0] = x; r[1] = y; r[2] = z;
}
template <class R> inline R norm3(const R* a) { return std::sqrt(dot3(a, a)); }
// MJX math.normalize_with_norm: x / (n + 1e-6*(n==0))
template <class R> inline R normalize3(R* a) {
  R n = norm3(a);
  R d = n + (n == R(0) ? R(1e-6) : R(0));
  a[0] /= d; a[1] /= d; a[2] /= d;
  return n;
}
template <class R> inline void quat_mul(R* r, const R* a, const R* b) {
  R w = a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3];
  R x = a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2];
  R y = a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1];
  R z = a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0];
  r[0] = w; r[1] = x; r[2] = y; r[3] = z;
}
// MJX math.rotate(vec, quat)
template <class R> inline void rotate(R* r, const R* v, const R* q) {
  R s = q[0];
  const R* u = q + 1;
  R uv = dot3(u, v), uu = dot3(u, u);
  R c[3];
  cross3(c, u, v);
  for (int k = 0; k < 3; k++) r[k] = R(2) * (uv * u[k]) + (s * s - uu) * v[k] + R(2) * s * c[k];
}
template <class R> inline void quat_to_mat(R* m, const R* q) {
  R w = q[0], x = q[1], y = q[2], z = q[3];
  m[0] = w * w + x * x - y * y - z * z; m[1] = R(2) * (x * y - w * z); m[2] = R(2) * (x * z + w * y);
  m[3] = R(2) * (x * y + w * z); m[4] = w * w - x * x + y * y - z * z; m[5] = R(2) * (y * z - w * x);
  m[6] = R(2) * (x * z - w * y); m[7] = R(2) * (y * z + w * x); m[8] = w * w - x * x - y * y + z * z;
}
template <class R> inline void axis_angle_quat(R* q, const R* axis, R angle) {
  R s = std::sin(angle * R(0.5)), c = std::cos(angle * R(0.5));
  q[0] = c; q[1] = axis[0] * s; q[2] = axis[1] * s; q[3] = axis[2] * s;
}
template <class R> inline void mat_vec(R* r, const R* m, const R* v) {
  R x = m[0] * v[0] + m[1] * v[1] + m[2] * v[2];
  R y = m[3] * v[0] + m[4] * v[1] + m[5] * v[2];
  R z = m[6] * v[0] + m[7] * v[1] + m[8] * v[2];
  r[0] = x; r[1] = y; r[2] = z;
}
template <class R> inline void matT_vec(R* r, const R* m, const R* v) {
  R x = m[0] * v[0] + m[3] * v[1] + m[6] * v[2];
  R y = m[1] * v[0] + m[4] * v[1] + m[7] * v[2];
  R z = m[2] * v[0] + m[5] * v[1] + m[8] * v[2];
  r[0] = x; r[1] = y; r[2] = z;
}
// spatial helpers on 6-vectors [ang; lin] and 10-vector inertias (MJX math.inert_mul / motion_cross / motion_cross_force)
template <class R> inline void inert_mul(R* r, const R* i, const R* v) {
  // i = [xx yy zz xy xz yz, m*off(3), m]
  R ang[3] = {i[0] * v[0] + i[3] * v[1] + i[4] * v[2], i[3] * v[0] + i[1] * v[1] + i[5] * v[2],
              i[4] * v[0] + i[5] * v[1] + i[2] * v[2]};
  R c1[3], c2[3];
  cross3(c1, i + 6, v + 3);
  cross3(c2, i + 6, v);
  for (int k = 0; k < 3; k++) { r[k] = ang[k] + c1[k]; r[3 + k] = i[9] * v[3 + k] - c2[k]; }
}
template <class R> inline void motion_cross(R* r, const R* u, const R* v) {
  R a[3], b[3], c[3];
  cross3(a, u, v); cross3(b, u, v + 3); cross3(c, u + 3, v);
  for (int k = 0; k < 3; k++) { r[k] = a[k]; r[3 + k] = b[k] + c[k]; }
}
template <class R> inline void motion_cross_force(R* r, const R* v, const R* f) {
  R a[3], b[3], c[3];
  cross3(a, v, f); cross3(b, v + 3, f + 3); cross3(c, v, f + 3);
  for (int k = 0; k < 3; k++) { r[k] = a[k] + b[k]; r[3 + k] = c[k]; }
}

// ---------------------------------------------------------------- data
template <class R> struct Data {
  // state (the fields of mjx.Data the env persists)
  R qpos[MJXB_MAXQ], qvel[MJXB_MAXDOF], ctrl[MJXB_MAXU], qacc_warmstart[MJXB_MAXDOF], time;
  // fwd_position
  R xpos[MJXB_MAXBODY][3], xquat[MJXB_MAXBODY][4], xmat[MJXB_MAXBODY][9], xipos[MJXB_MAXBODY][3];
  R xanchor[MJXB_MAXJNT][3], xaxis[MJXB_MAXJNT][3];
  R geom_xpos[MJXB_MAXGEOM][3], geom_xmat[MJXB_MAXGEOM][9];
  R site_xpos[MJXB_MAXSITE][3], site_xmat[MJXB_MAXSITE][9];
  R subtree_com[MJXB_MAXBODY][3];
  R cinert[MJXB_MAXBODY][10], crb[MJXB_MAXBODY][10], cdof[MJXB_MAXDOF][6], cdof_dot[MJXB_MAXDOF][6];
  R cvel[MJXB_MAXBODY][6], cacc[MJXB_MAXBODY][6], cfrc[MJXB_MAXBODY][6];
  R qM[MJXB_MAXDOF][MJXB_MAXDOF], qL[MJXB_MAXDOF][MJXB_MAXDOF];
  R ten_length[MJXB_MAXTENDON];
  R con_dist[MAXCON], con_pos[MAXCON][3], con_frame[MAXCON][9];
  R efc_J[MAXEFC][MJXB_MAXDOF], efc_pos[MAXEFC], efc_D[MAXEFC], efc_aref[MAXEFC], efc_force[MAXEFC];
  int efc_cand[MAXEFC], efc_active[MAXEFC];
  R qfrc_passive[MJXB_MAXDOF], qfrc_bias[MJXB_MAXDOF], qfrc_actuator[MJXB_MAXDOF], qfrc_smooth[MJXB_MAXDOF];
  R qacc_smooth[MJXB_MAXDOF], qacc[MJXB_MAXDOF], qfrc_constraint[MJXB_MAXDOF];
  R sensordata[MJXB_MAXSENSOR];
  int solver_niter;
  long flops;  // counted multiply/add operations of the dense formulation (SURVEY.md 8d counting model)
  // the same step counted activity-aware (SURVEY.md 8d "activity-aware minimum"): only candidate rows are assembled, only rows active at
  // the current iterate enter J^T D J (756 = 2 * 27*28/2 flops per active row), tree-sparse factor_m; counting rules in DESIGN.md
  long flops_act;
};

// ---------------------------------------------------------------- dense Cholesky (jax.scipy.linalg.cho_factor/cho_solve)
template <class R> inline bool chol_factor(R (*L)[MJXB_MAXDOF], const R (*A)[MJXB_MAXDOF], int n) {
  bool ok = true;
  for (int j = 0; j < n; j++) {
    R s = A[j][j];
    for (int k = 0; k < j; k++) s -= L[j][k] * L[j][k];
    if (!(s > R(0))) ok = false;
    R d = std::sqrt(s);
    L[j][j] = d;
    for (int i = j + 1; i < n; i++) {
      R t = A[i][j];
      for (int k = 0; k < j; k++) t -= L[i][k] * L[j][k];
      L[i][j] = t / d;
    }
  }
  return ok;
}
template <class R> inline void chol_solve(R* x, const R (*L)[MJXB_MAXDOF], const R* b, int n) {
  R y[MJXB_MAXDOF];
  for (int i = 0; i < n; i++) {
    R s = b[i];
    for (int k = 0; k < i; k++) s -= L[i][k] * y[k];
    y[i] = s / L[i][i];
  }
  for (int i = n - 1; i >= 0; i--) {
    R s = y[i];
    for (int k = i + 1; k < n; k++) s -= L[k][i] * x[k];
    x[i] = s / L[i][i];
  }
}

// ---------------------------------------------------------------- fwd_position: kinematics (mjx smooth.kinematics)
template <class R> void kinematics(const mjxb_model_blob& m, Data<R>& d) {
  for (int k = 0; k < 3; k++) d.xpos[0][k] = 0;
  d.xquat[0][0] = 1; d.xquat[0][1] = d.xquat[0][2] = d.xquat[0][3] = 0;
  for (int b = 1; b < m.nbody; b++) {
    int p = m.body_parent[b];
    R bp[3] = {R(m.body_pos[b][0]), R(m.body_pos[b][1]), R(m.body_pos[b][2])};
    R bq[4] = {R(m.body_quat[b][0]), R(m.body_quat[b][1]), R(m.body_quat[b][2]), R(m.body_quat[b][3])};
    R pos[3], quat[4], t[3];
    rotate(t, bp, d.xquat[p]);
    for (int k = 0; k < 3; k++) pos[k] = d.xpos[p][k] + t[k];
    quat_mul(quat, d.xquat[p], bq);
    for (int j = m.body_jntadr[b]; j < m.body_jntadr[b] + m.body_jntnum[b]; j++) {
      int qa = m.jnt_qposadr[j];
      R jp[3] = {R(m.jnt_pos[j][0]), R(m.jnt_pos[j][1]), R(m.jnt_pos[j][2])};
      R ja[3] = {R(m.jnt_axis[j][0]), R(m.jnt_axis[j][1]), R(m.jnt_axis[j][2])};
      if (m.jnt_type[j] == 0) {  // free
        for (int k = 0; k < 3; k++) { d.xanchor[j][k] = d.qpos[qa + k]; pos[k] = d.qpos[qa + k]; }
        d.xaxis[j][0] = 0; d.xaxis[j][1] = 0; d.xaxis[j][2] = 1;
        R* q = d.qpos + qa + 3;
        R n = std::sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
        R dn = n + (n == R(0) ? R(1e-6) : R(0));
        for (int k = 0; k < 4; k++) { q[k] = q[k] / dn; quat[k] = q[k]; }  // kinematics normalises qpos in place
      } else {  // hinge
        rotate(t, jp, quat);
        for (int k = 0; k < 3; k++) d.xanchor[j][k] = t[k] + pos[k];
        rotate(d.xaxis[j], ja, quat);
        R ql[4], qn[4];
        axis_angle_quat(ql, ja, d.qpos[qa] - R(m.qpos0[qa]));
        quat_mul(qn, quat, ql);
        for (int k = 0; k < 4; k++) quat[k] = qn[k];
        rotate(t, jp, quat);
        for (int k = 0; k < 3; k++) pos[k] = d.xanchor[j][k] - t[k];
      }
    }
    for (int k = 0; k < 3; k++) d.xpos[b][k] = pos[k];
    for (int k = 0; k < 4; k++) d.xquat[b][k] = quat[k];
  }
  for (int b = 0; b < m.nbody; b++) {
    quat_to_mat(d.xmat[b], d.xquat[b]);
    R ip[3] = {R(m.body_ipos[b][0]), R(m.body_ipos[b][1]), R(m.body_ipos[b][2])}, t[3];
    rotate(t, ip, d.xquat[b]);
    for (int k = 0; k < 3; k++) d.xipos[b][k] = d.xpos[b][k] + t[k];
  }
  for (int g = 0; g < m.ngeom; g++) {
    int b = m.geom_body[g];
    R gp[3] = {R(m.geom_pos[g][0]), R(m.geom_pos[g][1]), R(m.geom_pos[g][2])};
    R gq[4] = {R(m.geom_quat[g][0]), R(m.geom_quat[g][1]), R(m.geom_quat[g][2]), R(m.geom_quat[g][3])};
    R t[3], q[4];
    rotate(t, gp, d.xquat[b]);
    for (int k = 0; k < 3; k++) d.geom_xpos[g][k] = d.xpos[b][k] + t[k];
    quat_mul(q, d.xquat[b], gq);
    quat_to_mat(d.geom_xmat[g], q);
  }
  for (int s = 0; s < m.nsite; s++) {
    int b = m.site_body[s];
    R sp[3] = {R(m.site_pos[s][0]), R(m.site_pos[s][1]), R(m.site_pos[s][2])};
    R sq[4] = {R(m.site_quat[s][0]), R(m.site_quat[s][1]), R(m.site_quat[s][2]), R(m.site_quat[s][3])};
    R t[3], q[4];
    rotate(t, sp, d.xquat[b]);
    for (int k = 0; k < 3; k++) d.site_xpos[s][k] = d.xpos[b][k] + t[k];
    quat_mul(q, d.xquat[b], sq);
    quat_to_mat(d.site_xmat[s], q);
  }
  d.flops += 4000;
}

// ---------------------------------------------------------------- com_pos (mjx smooth.com_pos)
template <class R> void com_pos(const mjxb_model_blob& m, Data<R>& d) {
  R spos[MJXB_MAXBODY][3], smass[MJXB_MAXBODY];
  for (int b = 0; b < m.nbody; b++) {
    smass[b] = R(m.body_mass[b]);
    for (int k = 0; k < 3; k++) spos[b][k] = d.xipos[b][k] * R(m.body_mass[b]);
  }
  for (int b = m.nbody - 1; b > 0; b--) {
    int p = m.body_parent[b];
    smass[p] += smass[b];
    for (int k = 0; k < 3; k++) spos[p][k] += spos[b][k];
  }
  for (int b = 0; b < m.nbody; b++)
    for (int k = 0; k < 3; k++)
      d.subtree_com[b][k] = smass[b] < R(MINVAL) ? d.xipos[b][k] : spos[b][k] / std::max(smass[b], R(MINVAL));
  // root of every body's kinematic tree
  int root[MJXB_MAXBODY];
  root[0] = 0;
  for (int b = 1; b < m.nbody; b++) root[b] = m.body_parent[b] == 0 ? b : root[m.body_parent[b]];
  for (int b = 0; b < m.nbody; b++) {
    R off[3];
    for (int k = 0; k < 3; k++) off[k] = d.xipos[b][k] - d.subtree_com[root[b]][k];
    R mass = R(m.body_mass[b]);
    // body inertia tensor (about ipos, body axes, full symmetric) rotated to world: Iw = X I X^T
    const float* bi = m.body_inertia[b];
    R I[9] = {R(bi[0]), R(bi[3]), R(bi[4]), R(bi[3]), R(bi[1]), R(bi[5]), R(bi[4]), R(bi[5]), R(bi[2])};
    const R* X = d.xmat[b];
    R XI[9], Iw[9];
    for (int r = 0; r < 3; r++)
      for (int c = 0; c < 3; c++) XI[3 * r + c] = X[3 * r] * I[c] + X[3 * r + 1] * I[3 + c] + X[3 * r + 2] * I[6 + c];
    for (int r = 0; r < 3; r++)
      for (int c = 0; c < 3; c++) Iw[3 * r + c] = XI[3 * r] * X[3 * c] + XI[3 * r + 1] * X[3 * c + 1] + XI[3 * r + 2] * X[3 * c + 2];
    // + mass * (|off|^2 E - off off^T)
    R oo = dot3(off, off);
    R* ci = d.cinert[b];
    ci[0] = Iw[0] + mass * (oo - off[0] * off[0]);
    ci[1] = Iw[4] + mass * (oo - off[1] * off[1]);
    ci[2] = Iw[8] + mass * (oo - off[2] * off[2]);
    ci[3] = Iw[1] - mass * off[0] * off[1];
    ci[4] = Iw[2] - mass * off[0] * off[2];
    ci[5] = Iw[5] - mass * off[1] * off[2];
    for (int k = 0; k < 3; k++) ci[6 + k] = mass * off[k];
    ci[9] = mass;
  }
  // cdof: [axis; axis x (com - anchor)]
  for (int j = 0; j < m.njnt; j++) {
    int b = m.jnt_body[j], da = m.jnt_dofadr[j];
    R off[3];
    for (int k = 0; k < 3; k++) off[k] = d.subtree_com[root[b]][k] - d.xanchor[j][k];
    if (m.jnt_type[j] == 0) {
      for (int i = 0; i < 3; i++) {
        for (int k = 0; k < 6; k++) d.cdof[da + i][k] = 0;
        d.cdof[da + i][3 + i] = 1;
        R ax[3] = {d.xmat[b][i], d.xmat[b][3 + i], d.xmat[b][6 + i]};  // column i of xmat (body-frame axis)
        for (int k = 0; k < 3; k++) d.cdof[da + 3 + i][k] = ax[k];
        cross3(d.cdof[da + 3 + i] + 3, ax, off);
      }
    } else {
      for (int k = 0; k < 3; k++) d.cdof[da][k] = d.xaxis[j][k];
      cross3(d.cdof[da] + 3, d.xaxis[j], off);
    }
  }
  d.flops += 3000;
}

// ---------------------------------------------------------------- tendon, crb, factor_m
template <class R> void tendon(const mjxb_model_blob& m, Data<R>& d) {
  for (int t = 0; t < m.ntendon; t++) {
    R len = 0;
    for (int w = 0; w < m.ten_nwrap[t]; w++) len += R(m.ten_coef[t][w]) * d.qpos[m.ten_qpos[t][w]];
    d.ten_length[t] = len;
  }
}
template <class R> void crb(const mjxb_model_blob& m, Data<R>& d) {
  for (int b = 0; b < m.nbody; b++)
    for (int k = 0; k < 10; k++) d.crb[b][k] = d.cinert[b][k];
  for (int b = m.nbody - 1; b > 0; b--) {
    int p = m.body_parent[b];
    for (int k = 0; k < 10; k++) d.crb[p][k] += d.crb[b][k];
  }
  for (int k = 0; k < 10; k++) d.crb[0][k] = 0;
  for (int i = 0; i < m.nv; i++)
    for (int j = 0; j < m.nv; j++) d.qM[i][j] = 0;
  for (int i = 0; i < m.nv; i++) {
    R f[6];
    inert_mul(f, d.crb[m.dof_body[i]], d.cdof[i]);
    for (int j = i; j >= 0; j = m.dof_parent[j]) {
      R s = 0;
      for (int k = 0; k < 6; k++) s += d.cdof[j][k] * f[k];
      d.qM[i][j] = s;
      d.qM[j][i] = s;
    }
    d.qM[i][i] += R(m.dof_armature[i]);
  }
  d.flops += 4000;
}

// ---------------------------------------------------------------- collision (mjx collision_primitive)
template <class R> inline void make_frame(R* frame, const R* nin) {
  R a[3] = {nin[0], nin[1], nin[2]};
  normalize3(a);
  R b[3] = {0, 0, 0};
  if (R(-0.5) < a[1] && a[1] < R(0.5)) b[1] = 1; else b[2] = 1;
  R ab = dot3(a, b);
  for (int k = 0; k < 3; k++) b[k] -= a[k] * ab;
  normalize3(b);
  bool any = (a[0] != R(0)) || (a[1] != R(0)) || (a[2] != R(0));
  if (!any) b[0] = b[1] = b[2] = 0;
  for (int k = 0; k < 3; k++) { frame[k] = a[k]; frame[3 + k] = b[k]; }
  cross3(frame + 6, a, b);
}
template <class R> inline void sphere_sphere(R& dist, R* pos, R* n, const R* p1, R r1, const R* p2, R r2) {
  for (int k = 0; k < 3; k++) n[k] = p2[k] - p1[k];
  R len = normalize3(n);
  if (len == R(0)) { n[0] = 1; n[1] = 0; n[2] = 0; }
  dist = len - (r1 + r2);
  for (int k = 0; k < 3; k++) pos[k] = p1[k] + n[k] * (r1 + dist * R(0.5));
}
template <class R> inline void closest_segment_point(R* out, const R* a, const R* b, const R* pt) {
  R ab[3], pa[3];
  for (int k = 0; k < 3; k++) { ab[k] = b[k] - a[k]; pa[k] = pt[k] - a[k]; }
  R t = dot3(pa, ab) / (dot3(ab, ab) + R(1e-6));
  t = std::min(std::max(t, R(0)), R(1));
  for (int k = 0; k < 3; k++) out[k] = a[k] + t * ab[k];
}
template <class R>
inline void closest_segment_to_segment(R* best_a, R* best_b, const R* a0, const R* a1, const R* b0, const R* b1) {
  R dir_a[3], dir_b[3];
  for (int k = 0; k < 3; k++) { dir_a[k] = a1[k] - a0[k]; dir_b[k] = b1[k] - b0[k]; }
  R len_a = normalize3(dir_a), len_b = normalize3(dir_b);
  R half_a = len_a * R(0.5), half_b = len_b * R(0.5);
  R a_mid[3], b_mid[3], trans[3];
  for (int k = 0; k < 3; k++) {
    a_mid[k] = a0[k] + dir_a[k] * half_a;
    b_mid[k] = b0[k] + dir_b[k] * half_b;
    trans[k] = a_mid[k] - b_mid[k];
  }
  R dd = dot3(dir_a, dir_b), da_t = dot3(dir_a, trans), db_t = dot3(dir_b, trans);
  R denom = R(1) - dd * dd;
  R orig_ta = (-da_t + dd * db_t) / (denom + R(1e-6));
  R orig_tb = db_t + orig_ta * dd;
  R ta = std::min(std::max(orig_ta, -half_a), half_a);
  R tb = std::min(std::max(orig_tb, -half_b), half_b);
  for (int k = 0; k < 3; k++) { best_a[k] = a_mid[k] + dir_a[k] * ta; best_b[k] = b_mid[k] + dir_b[k] * tb; }
  R new_a[3], new_b[3];
  closest_segment_point(new_a, a0, a1, best_b);
  closest_segment_point(new_b, b0, b1, best_a);
  R d1 = 0, d2 = 0;
  for (int k = 0; k < 3; k++) {
    d1 += (new_a[k] - best_b[k]) * (new_a[k] - best_b[k]);
    d2 += (best_a[k] - new_b[k]) * (best_a[k] - new_b[k]);
  }
  if (d1 < d2) { for (int k = 0; k < 3; k++) best_a[k] = new_a[k]; }
  else { for (int k = 0; k < 3; k++) best_b[k] = new_b[k]; }
}

template <class R> void collision(const mjxb_model_blob& m, Data<R>& d) {
  for (int p = 0; p < m.npair; p++) {
    int g1 = m.pair_g1[p], g2 = m.pair_g2[p], c = m.pair_conadr[p];
    const R* p1 = d.geom_xpos[g1];
    const R* p2 = d.geom_xpos[g2];
    const R* m1 = d.geom_xmat[g1];
    const R* m2 = d.geom_xmat[g2];
    R ax1[3] = {m1[2], m1[5], m1[8]}, ax2[3] = {m2[2], m2[5], m2[8]};
    R r1 = R(m.geom_size[g1][0]), r2 = R(m.geom_size[g2][0]);
    R l1 = R(m.geom_size[g1][1]), l2 = R(m.geom_size[g2][1]);
    R n[3];
    switch (m.pair_kind[p]) {
      case 0: {  // plane - sphere
        R diff[3] = {p2[0] - p1[0], p2[1] - p1[1], p2[2] - p1[2]};
        R dist = dot3(diff, ax1) - r2;
        d.con_dist[c] = dist;
        for (int k = 0; k < 3; k++) d.con_pos[c][k] = p2[k] - ax1[k] * (r2 + R(0.5) * dist);
        make_frame(d.con_frame[c], ax1);
      } break;
      case 1: {  // plane - capsule: two contacts at +segment, -segment
        R na = dot3(ax1, ax2);
        R b[3] = {ax2[0] - ax1[0] * na, ax2[1] - ax1[1] * na, ax2[2] - ax1[2] * na};
        R bn = normalize3(b);
        if (bn < R(0.5)) {
          b[0] = 0; b[1] = 0; b[2] = 0;
          if (R(-0.5) < ax1[1] && ax1[1] < R(0.5)) b[1] = 1; else b[2] = 1;
        }
        R fr[9];
        for (int k = 0; k < 3; k++) { fr[k] = ax1[k]; fr[3 + k] = b[k]; }
        cross3(fr + 6, ax1, b);
        for (int e = 0; e < 2; e++) {
          R sg = e == 0 ? R(1) : R(-1);
          R sp[3], diff[3];
          for (int k = 0; k < 3; k++) { sp[k] = p2[k] + sg * (ax2[k] * l2); diff[k] = sp[k] - p1[k]; }
          R dist = dot3(diff, ax1) - r2;
          d.con_dist[c + e] = dist;
          for (int k = 0; k < 3; k++) d.con_pos[c + e][k] = sp[k] - ax1[k] * (r2 + R(0.5) * dist);
          for (int k = 0; k < 9; k++) d.con_frame[c + e][k] = fr[k];
        }
      } break;
      case 2: {  // sphere - sphere
        sphere_sphere(d.con_dist[c], d.con_pos[c], n, p1, r1, p2, r2);
        make_frame(d.con_frame[c], n);
      } break;
      case 3: {  // sphere - capsule
        R a[3], b[3], pt[3];
        for (int k = 0; k < 3; k++) { a[k] = p2[k] - ax2[k] * l2; b[k] = p2[k] + ax2[k] * l2; }
        closest_segment_point(pt, a, b, p1);
        sphere_sphere(d.con_dist[c], d.con_pos[c], n, p1, r1, pt, r2);
        make_frame(d.con_frame[c], n);
      } break;
      case 4: {  // capsule - capsule
        R a0[3], a1[3], b0[3], b1[3], pa[3], pb[3];
        for (int k = 0; k < 3; k++) {
          a0[k] = p1[k] - ax1[k] * l1; a1[k] = p1[k] + ax1[k] * l1;
          b0[k] = p2[k] - ax2[k] * l2; b1[k] = p2[k] + ax2[k] * l2;
        }
        closest_segment_to_segment(pa, pb, a0, a1, b0, b1);
        sphere_sphere(d.con_dist[c], d.con_pos[c], n, pa, r1, pb, r2);
        make_frame(d.con_frame[c], n);
      } break;
    }
  }
  d.flops += 13000;
}

// ---------------------------------------------------------------- make_constraint (mjx constraint.py)
template <class R> inline void kbi(const mjxb_model_blob& m, const float* solref, const float* solimp, R pos, R& k, R& b, R& imp) {
  R timeconst = R(solref[0]), dampratio = R(solref[1]);
  timeconst = std::max(timeconst, R(2) * R(m.timestep));  // refsafe
  R dmin = std::min(std::max(R(solimp[0]), R(MINIMP)), R(MAXIMP));
  R dmax = std::min(std::max(R(solimp[1]), R(MINIMP)), R(MAXIMP));
  R width = std::max(R(MINVAL), R(solimp[2]));
  R mid = std::min(std::max(R(solimp[3]), R(MINIMP)), R(MAXIMP));
  R power = std::max(R(1), R(solimp[4]));
  k = R(1) / (dmax * dmax * timeconst * timeconst * dampratio * dampratio);
  b = R(2) / (dmax * timeconst);
  if (solref[0] <= 0) k = -R(solref[0]) / (dmax * dmax);
  if (solref[1] <= 0) b = -R(solref[1]) / dmax;
  R x = std::abs(pos) / width;
  R ia = (R(1) / std::pow(mid, power - R(1))) * std::pow(x, power);
  R ib = R(1) - (R(1) / std::pow(R(1) - mid, power - R(1))) * std::pow(R(1) - x, power);
  R y = x < mid ? ia : ib;
  imp = dmin + y * (dmax - dmin);
  imp = std::min(std::max(imp, dmin), dmax);
  if (x > R(1)) imp = dmax;
}

template <class R>
inline void finish_row(const mjxb_model_blob& m, Data<R>& d, int r, R pos, R invweight, const float* solref, const float* solimp, bool cand) {
  // J row r already holds the (unmasked) Jacobian; MJX masks inactive rows to zero.
  if (!cand) {
    for (int j = 0; j < m.nv; j++) d.efc_J[r][j] = 0;
    d.efc_pos[r] = 0; d.efc_D[r] = 0; d.efc_aref[r] = 0; d.efc_cand[r] = 0;
    return;
  }
  R k, b, imp;
  kbi(m, solref, solimp, pos, k, b, imp);
  R rr = std::max(invweight * (R(1) - imp) / imp, R(MINVAL));
  R vel = 0;
  for (int j = 0; j < m.nv; j++) vel += d.efc_J[r][j] * d.qvel[j];
  d.efc_pos[r] = pos;
  d.efc_D[r] = R(1) / rr;
  d.efc_aref[r] = -b * vel - k * imp * pos;
  d.efc_cand[r] = 1;
}

template <class R> inline void point_jac(const mjxb_model_blob& m, const Data<R>& d, R (*jac)[3], const R* point, int body) {
  // mjx support.jac: translational Jacobian of `point` attached to `body`; jac[dof][xyz]
  R off[3];
  int root = body;
  while (root > 0 && m.body_parent[root] != 0) root = m.body_parent[root];
  for (int k = 0; k < 3; k++) off[k] = point[k] - d.subtree_com[root][k];
  for (int j = 0; j < m.nv; j++) jac[j][0] = jac[j][1] = jac[j][2] = 0;
  // dofs on the chain body -> root
  int b = body;
  while (b > 0 && m.body_dofnum[b] == 0) b = m.body_parent[b];
  if (b == 0) return;
  int last = m.body_dofadr[b] + m.body_dofnum[b] - 1;
  for (int j = last; j >= 0; j = m.dof_parent[j]) {
    R c[3];
    cross3(c, d.cdof[j], off);
    for (int k = 0; k < 3; k++) jac[j][k] = d.cdof[j][3 + k] + c[k];
  }
}

template <class R> void make_constraint(const mjxb_model_blob& m, Data<R>& d) {
  int r = 0;
  // joint limits
  for (int i = 0; i < m.nlimit; i++, r++) {
    int j = m.lim_jnt[i], qa = m.jnt_qposadr[j], da = m.jnt_dofadr[j];
    R q = d.qpos[qa];
    R dmin = q - R(m.jnt_range[j][0]), dmax = R(m.jnt_range[j][1]) - q;
    R pos = std::min(dmin, dmax);
    for (int k = 0; k < m.nv; k++) d.efc_J[r][k] = 0;
    d.efc_J[r][da] = dmin < dmax ? R(1) : R(-1);
    finish_row(m, d, r, pos, R(m.dof_invweight0[da]), m.jnt_solref[j], m.jnt_solimp[j], pos < R(0));
  }
  // tendon limits
  for (int i = 0; i < m.ntlimit; i++, r++) {
    int t = m.lim_ten[i];
    R len = d.ten_length[t];
    R dmin = len - R(m.ten_range[t][0]), dmax = R(m.ten_range[t][1]) - len;
    R pos = std::min(dmin, dmax);
    R sg = dmin < dmax ? R(1) : R(-1);
    for (int k = 0; k < m.nv; k++) d.efc_J[r][k] = 0;
    for (int w = 0; w < m.ten_nwrap[t]; w++) d.efc_J[r][m.ten_dof[t][w]] = sg * R(m.ten_coef[t][w]);
    finish_row(m, d, r, pos, R(m.ten_invweight0[t]), m.ten_solref[t], m.ten_solimp[t], pos < R(0));
  }
  // contacts
  R jac1[MJXB_MAXDOF][3], jac2[MJXB_MAXDOF][3];
  for (int p = 0; p < m.npair; p++) {
    int b1 = m.geom_body[m.pair_g1[p]], b2 = m.geom_body[m.pair_g2[p]];
    int ncon = (m.pair_kind[p] == 1) ? 2 : 1;
    for (int e = 0; e < ncon; e++) {
      int c = m.pair_conadr[p] + e;
      R pos = d.con_dist[c];
      bool cand = pos < R(0);
      point_jac(m, d, jac1, d.con_pos[c], b1);
      point_jac(m, d, jac2, d.con_pos[c], b2);
      const R* fr = d.con_frame[c];
      if (m.pair_condim[p] == 1) {
        int row = m.pair_efcadr[p] + e;
        for (int j = 0; j < m.nv; j++) {
          R df[3] = {jac2[j][0] - jac1[j][0], jac2[j][1] - jac1[j][1], jac2[j][2] - jac1[j][2]};
          d.efc_J[row][j] = dot3(fr, df);
        }
        finish_row(m, d, row, pos, R(m.pair_invweight[p]), m.pair_solref[p], m.pair_solimp[p], cand);
      } else {
        int row0 = m.pair_efcadr[p] + 4 * e;
        R mu = R(m.pair_mu[p]);
        for (int j = 0; j < m.nv; j++) {
          R df[3] = {jac2[j][0] - jac1[j][0], jac2[j][1] - jac1[j][1], jac2[j][2] - jac1[j][2]};
          R jn = dot3(fr, df), j1 = dot3(fr + 3, df), j2 = dot3(fr + 6, df);
          d.efc_J[row0 + 0][j] = jn + j1 * mu;
          d.efc_J[row0 + 1][j] = jn - j1 * mu;
          d.efc_J[row0 + 2][j] = jn + j2 * mu;
          d.efc_J[row0 + 3][j] = jn - j2 * mu;
        }
        for (int q = 0; q < 4; q++)
          finish_row(m, d, row0 + q, pos, R(m.pair_invweight[p]), m.pair_solref[p], m.pair_solimp[p], cand);
      }
    }
  }
  d.flops += 98000;
}

// ---------------------------------------------------------------- fwd_velocity / actuation / acceleration
template <class R> void com_vel(const mjxb_model_blob& m, Data<R>& d) {
  for (int k = 0; k < 6; k++) d.cvel[0][k] = 0;
  for (int b = 1; b < m.nbody; b++) {
    R cvel[6];
    for (int k = 0; k < 6; k++) cvel[k] = d.cvel[m.body_parent[b]][k];
    for (int j = m.body_jntadr[b]; j < m.body_jntadr[b] + m.body_jntnum[b]; j++) {
      int da = m.jnt_dofadr[j];
      if (m.jnt_type[j] == 0) {
        for (int i = 0; i < 3; i++)
          for (int k = 0; k < 6; k++) cvel[k] += d.cdof[da + i][k] * d.qvel[da + i];
        for (int i = 0; i < 3; i++) {
          for (int k = 0; k < 6; k++) d.cdof_dot[da + i][k] = 0;
          motion_cross(d.cdof_dot[da + 3 + i], cvel, d.cdof[da + 3 + i]);
        }
        for (int i = 3; i < 6; i++)
          for (int k = 0; k < 6; k++) cvel[k] += d.cdof[da + i][k] * d.qvel[da + i];
      } else {
        motion_cross(d.cdof_dot[da], cvel, d.cdof[da]);
        for (int k = 0; k < 6; k++) cvel[k] += d.cdof[da][k] * d.qvel[da];
      }
    }
    for (int k = 0; k < 6; k++) d.cvel[b][k] = cvel[k];
  }
}
template <class R> void passive(const mjxb_model_blob& m, Data<R>& d) {
  for (int i = 0; i < m.nv; i++) d.qfrc_passive[i] = 0;
  for (int j = 0; j < m.njnt; j++) {
    if (m.jnt_type[j] == 0) continue;
    int qa = m.jnt_qposadr[j], da = m.jnt_dofadr[j];
    d.qfrc_passive[da] = -R(m.dof_stiffness[da]) * (d.qpos[qa] - R(m.qpos_spring[qa])) - R(m.dof_damping[da]) * d.qvel[da];
  }
}
template <class R> void rne(const mjxb_model_blob& m, Data<R>& d) {
  for (int k = 0; k < 3; k++) { d.cacc[0][k] = 0; d.cacc[0][3 + k] = -R(m.gravity[k]); }
  for (int b = 1; b < m.nbody; b++) {
    for (int k = 0; k < 6; k++) d.cacc[b][k] = d.cacc[m.body_parent[b]][k];
    for (int i = m.body_dofadr[b]; i >= 0 && i < m.body_dofadr[b] + m.body_dofnum[b]; i++)
      for (int k = 0; k < 6; k++) d.cacc[b][k] += d.cdof_dot[i][k] * d.qvel[i];
  }
  for (int b = 0; b < m.nbody; b++) {
    R f1[6], iv[6], f2[6];
    inert_mul(f1, d.cinert[b], d.cacc[b]);
    inert_mul(iv, d.cinert[b], d.cvel[b]);
    motion_cross_force(f2, d.cvel[b], iv);
    for (int k = 0; k < 6; k++) d.cfrc[b][k] = f1[k] + f2[k];
  }
  for (int b = m.nbody - 1; b > 0; b--)
    for (int k = 0; k < 6; k++) d.cfrc[m.body_parent[b]][k] += d.cfrc[b][k];
  for (int i = 0; i < m.nv; i++) {
    R s = 0;
    for (int k = 0; k < 6; k++) s += d.cdof[i][k] * d.cfrc[m.dof_body[i]][k];
    d.qfrc_bias[i] = s;
  }
  d.flops += 6000;
}
template <class R> void fwd_actuation(const mjxb_model_blob& m, Data<R>& d) {
  for (int i = 0; i < m.nv; i++) d.qfrc_actuator[i] = 0;
  for (int u = 0; u < m.nu; u++) {
    R c = d.ctrl[u];
    if (m.act_ctrllimited[u]) c = std::min(std::max(c, R(m.act_ctrlrange[u][0])), R(m.act_ctrlrange[u][1]));
    d.qfrc_actuator[m.act_dof[u]] += R(m.act_gear[u]) * c;
  }
}
template <class R> void fwd_acceleration(const mjxb_model_blob& m, Data<R>& d) {
  for (int i = 0; i < m.nv; i++) d.qfrc_smooth[i] = d.qfrc_passive[i] - d.qfrc_bias[i] + d.qfrc_actuator[i];
  chol_solve(d.qacc_smooth, d.qL, d.qfrc_smooth, m.nv);
  d.flops += 1500;
}

// ---------------------------------------------------------------- solver (mjx solver.py)
template <class R> struct SolverCtx {
  R qacc[MJXB_MAXDOF], Ma[MJXB_MAXDOF], grad[MJXB_MAXDOF], Mgrad[MJXB_MAXDOF], search[MJXB_MAXDOF];
  R qfrc_constraint[MJXB_MAXDOF];
  R Jaref[MAXEFC], efc_force[MAXEFC];
  int active[MAXEFC];
  R gauss, cost, prev_cost;
  int niter;
};
template <class R> struct LSPoint { R alpha, cost, deriv_0, deriv_1; };

template <class R> void update_constraint(const mjxb_model_blob& m, Data<R>& d, SolverCtx<R>& c) {
  R cs = 0;
  for (int r = 0; r < m.nefc; r++) {
    c.active[r] = c.Jaref[r] < R(0);
    c.efc_force[r] = d.efc_D[r] * -c.Jaref[r] * R(c.active[r]);
  }
  for (int j = 0; j < m.nv; j++) {
    R s = 0;
    for (int r = 0; r < m.nefc; r++) s += d.efc_J[r][j] * c.efc_force[r];
    c.qfrc_constraint[j] = s;
  }
  R g = 0;
  for (int j = 0; j < m.nv; j++) g += (c.Ma[j] - d.qfrc_smooth[j]) * (c.qacc[j] - d.qacc_smooth[j]);
  c.gauss = R(0.5) * g;
  for (int r = 0; r < m.nefc; r++) cs += d.efc_D[r] * c.Jaref[r] * c.Jaref[r] * R(c.active[r]);
  c.prev_cost = c.cost;
  c.cost = R(0.5) * cs + c.gauss;
  d.flops += 2 * m.nefc * m.nv + 6 * m.nefc;
}
template <class R> void update_gradient(const mjxb_model_blob& m, Data<R>& d, SolverCtx<R>& c) {
  for (int j = 0; j < m.nv; j++) c.grad[j] = c.Ma[j] - d.qfrc_smooth[j] - c.qfrc_constraint[j];
  if (m.solver == 1) {  // CG: Mgrad = M^-1 grad
    chol_solve(c.Mgrad, d.qL, c.grad, m.nv);
    return;
  }
  static thread_local R H[MJXB_MAXDOF][MJXB_MAXDOF], L[MJXB_MAXDOF][MJXB_MAXDOF];
  for (int i = 0; i < m.nv; i++)
    for (int j = 0; j < m.nv; j++) {
      R s = 0;
      for (int r = 0; r < m.nefc; r++) s += d.efc_J[r][i] * d.efc_D[r] * R(c.active[r]) * d.efc_J[r][j];
      H[i][j] = d.qM[i][j] + s;
    }
  chol_factor(L, H, m.nv);
  chol_solve(c.Mgrad, L, c.grad, m.nv);
  d.flops += 2L * m.nefc * m.nv * (m.nv + 1) / 2 + 6600 + 1500;
  {  // activity-aware: J^T D J over the rows active at this iterate + factor/solve + the iteration's matvecs and line search (20K)
    long nact = 0;
    for (int r = 0; r < m.nefc; r++) nact += c.active[r] ? 1 : 0;
    d.flops_act += 756L * nact + 20000L;
  }
}
template <class R> void ctx_create(const mjxb_model_blob& m, Data<R>& d, SolverCtx<R>& c, const R* qacc, bool grad) {
  for (int j = 0; j < m.nv; j++) c.qacc[j] = qacc[j];
  for (int r = 0; r < m.nefc; r++) {
    R s = 0;
    for (int j = 0; j < m.nv; j++) s += d.efc_J[r][j] * qacc[j];
    c.Jaref[r] = s - d.efc_aref[r];
  }
  for (int i = 0; i < m.nv; i++) {
    R s = 0;
    for (int j = 0; j < m.nv; j++) s += d.qM[i][j] * qacc[j];
    c.Ma[i] = s;
  }
  for (int j = 0; j < m.nv; j++) c.grad[j] = c.Mgrad[j] = c.search[j] = 0;
  c.gauss = 0; c.cost = std::numeric_limits<R>::infinity(); c.prev_cost = 0; c.niter = 0;
  update_constraint(m, d, c);
  if (grad) {
    update_gradient(m, d, c);
    for (int j = 0; j < m.nv; j++) c.search[j] = -c.Mgrad[j];
  }
  d.flops += 2L * m.nefc * m.nv + 2L * m.nv * m.nv;
}
template <class R>
inline LSPoint<R> ls_point(const mjxb_model_blob& m, Data<R>& d, const SolverCtx<R>& c, R alpha, const R* jv, const R (*quad)[3], const R* qg) {
  R q0 = qg[0], q1 = qg[1], q2 = qg[2];
  for (int r = 0; r < m.nefc; r++) {
    R x = c.Jaref[r] + alpha * jv[r];
    if (x < R(0)) { q0 += quad[r][0]; q1 += quad[r][1]; q2 += quad[r][2]; }
  }
  LSPoint<R> p;
  p.alpha = alpha;
  p.cost = alpha * alpha * q2 + alpha * q1 + q0;
  p.deriv_0 = R(2) * alpha * q2 + q1;
  p.deriv_1 = R(2) * q2 + (q2 == R(0) ? R(MINVAL) : R(0));
  d.flops += 8 * m.nefc;
  return p;
}
template <class R> void linesearch(const mjxb_model_blob& m, Data<R>& d, SolverCtx<R>& c) {
  R nrm = 0;
  for (int j = 0; j < m.nv; j++) nrm += c.search[j] * c.search[j];
  R smag = std::sqrt(nrm) * R(m.meaninertia) * R(std::max(1, m.nv));
  R gtol = R(m.tolerance) * R(m.ls_tolerance) * smag;
  R mv[MJXB_MAXDOF];
  static thread_local R jv[MAXEFC], quad[MAXEFC][3];
  for (int i = 0; i < m.nv; i++) {
    R s = 0;
    for (int j = 0; j < m.nv; j++) s += d.qM[i][j] * c.search[j];
    mv[i] = s;
  }
  for (int r = 0; r < m.nefc; r++) {
    R s = 0;
    for (int j = 0; j < m.nv; j++) s += d.efc_J[r][j] * c.search[j];
    jv[r] = s;
  }
  R s1 = 0, s2 = 0, s3 = 0;
  for (int j = 0; j < m.nv; j++) { s1 += c.search[j] * c.Ma[j]; s2 += c.search[j] * d.qfrc_smooth[j]; s3 += c.search[j] * mv[j]; }
  R qg[3] = {c.gauss, s1 - s2, R(0.5) * s3};
  for (int r = 0; r < m.nefc; r++) {
    quad[r][0] = R(0.5) * c.Jaref[r] * c.Jaref[r] * d.efc_D[r];
    quad[r][1] = jv[r] * c.Jaref[r] * d.efc_D[r];
    quad[r][2] = R(0.5) * jv[r] * jv[r] * d.efc_D[r];
  }
  auto point = [&](R a) { return ls_point(m, d, c, a, jv, quad, qg); };
  LSPoint<R> p0 = point(R(0));
  LSPoint<R> lo = point(p0.alpha - p0.deriv_0 / p0.deriv_1);
  LSPoint<R> hi;
  if (lo.deriv_0 < p0.deriv_0) { hi = p0; } else { hi = lo; lo = p0; }
  bool swap = true;
  int ls_iter = 0;
  while (true) {
    bool done = ls_iter >= m.ls_iterations;
    done |= !swap;
    done |= (lo.deriv_0 < R(0)) && (lo.deriv_0 > -gtol);
    done |= (hi.deriv_0 > R(0)) && (hi.deriv_0 < gtol);
    if (done) break;
    LSPoint<R> lo_next = point(lo.alpha - lo.deriv_0 / lo.deriv_1);
    LSPoint<R> hi_next = point(hi.alpha - hi.deriv_0 / hi.deriv_1);
    LSPoint<R> mid = point(R(0.5) * (lo.alpha + hi.alpha));
    bool swap_lo_next = (lo.deriv_0 > R(0)) || (lo.deriv_0 < lo_next.deriv_0);
    if (swap_lo_next) lo = lo_next;
    bool swap_lo_mid = (mid.deriv_0 < R(0)) && (lo.deriv_0 < mid.deriv_0);
    if (swap_lo_mid) lo = mid;
    bool swap_hi_next = (hi.deriv_0 < R(0)) || (hi.deriv_0 > hi_next.deriv_0);
    if (swap_hi_next) hi = hi_next;
    bool swap_hi_mid = (mid.deriv_0 > R(0)) && (hi.deriv_0 > mid.deriv_0);
    if (swap_hi_mid) hi = mid;
    swap = swap_lo_next || swap_lo_mid || swap_hi_next || swap_hi_mid;
    ls_iter++;
  }
  bool improved = (lo.cost < p0.cost) || (hi.cost < p0.cost);
  R alpha = lo.cost < hi.cost ? lo.alpha : hi.alpha;
  if (improved) {
    for (int j = 0; j < m.nv; j++) { c.qacc[j] += c.search[j] * alpha; c.Ma[j] += mv[j] * alpha; }
    for (int r = 0; r < m.nefc; r++) c.Jaref[r] += jv[r] * alpha;
  }
  d.flops += 2L * m.nefc * m.nv + 2L * m.nv * m.nv + 8 * m.nefc;
}
template <class R> void solve(const mjxb_model_blob& m, Data<R>& d) {
  static thread_local SolverCtx<R> c, w;
  // warm start: cheaper of qacc_warmstart and qacc_smooth
  ctx_create(m, d, w, d.qacc_warmstart, false);
  R cost_warm = w.cost;
  ctx_create(m, d, w, d.qacc_smooth, false);
  R cost_smooth = w.cost;
  const R* q0 = cost_warm < cost_smooth ? d.qacc_warmstart : d.qacc_smooth;
  ctx_create(m, d, c, q0, true);
  R scale = R(1) / (R(m.meaninertia) * R(std::max(1, m.nv)));
  auto body = [&]() {
    linesearch(m, d, c);
    R prev_grad[MJXB_MAXDOF], prev_Mgrad[MJXB_MAXDOF];
    for (int j = 0; j < m.nv; j++) { prev_grad[j] = c.grad[j]; prev_Mgrad[j] = c.Mgrad[j]; }
    update_constraint(m, d, c);
    update_gradient(m, d, c);
    if (m.solver == 2) {
      for (int j = 0; j < m.nv; j++) c.search[j] = -c.Mgrad[j];
    } else {
      R num = 0, den = 0;
      for (int j = 0; j < m.nv; j++) { num += c.grad[j] * (c.Mgrad[j] - prev_Mgrad[j]); den += prev_grad[j] * prev_Mgrad[j]; }
      R beta = std::max(R(0), num / std::max(R(MINVAL), den));
      for (int j = 0; j < m.nv; j++) c.search[j] = -c.Mgrad[j] + beta * c.search[j];
    }
    c.niter++;
  };
  if (m.iterations == 1) {
    body();
  } else {
    while (true) {
      R improvement = (c.prev_cost - c.cost) * scale;
      R gn = 0;
      for (int j = 0; j < m.nv; j++) gn += c.grad[j] * c.grad[j];
      R gradient = std::sqrt(gn) * scale;
      bool done = c.niter >= m.iterations;
      done |= improvement < R(m.tolerance);
      done |= gradient < R(m.tolerance);
      if (done) break;
      body();
    }
  }
  for (int j = 0; j < m.nv; j++) {
    d.qacc[j] = c.qacc[j];
    d.qacc_warmstart[j] = c.qacc[j];
    d.qfrc_constraint[j] = c.qfrc_constraint[j];
  }
  for (int r = 0; r < m.nefc; r++) { d.efc_force[r] = c.efc_force[r]; d.efc_active[r] = c.active[r]; }
  d.solver_niter = c.niter;
}

// ---------------------------------------------------------------- touch sensor (engine_sensor.c mjSENS_TOUCH / mjx sensor.py)
template <class R> inline R ray_box(const R* size, const R* pnt, const R* vec) {
  // nearest x >= 0 with pnt + x*vec on a face of the box, else -1 (mjx ray._ray_box / engine_ray.c ray_box)
  R best = -1;
  static const int ifa[3][2] = {{1, 2}, {0, 2}, {0, 1}};
  for (int i = 0; i < 3; i++) {
    if (std::abs(vec[i]) <= R(MINVAL)) continue;
    for (int side = -1; side <= 1; side += 2) {
      R sol = (R(side) * size[i] - pnt[i]) / vec[i];
      if (sol >= R(0)) {
        R p0 = pnt[ifa[i][0]] + sol * vec[ifa[i][0]];
        R p1 = pnt[ifa[i][1]] + sol * vec[ifa[i][1]];
        if (std::abs(p0) <= size[ifa[i][0]] && std::abs(p1) <= size[ifa[i][1]])
          if (best < R(0) || sol < best) best = sol;
      }
    }
  }
  return best;
}
template <class R> void sensor_touch(const mjxb_model_blob& m, Data<R>& d) {
  for (int s = 0; s < m.nsensor; s++) {
    int site = m.sensor_site[s], body = m.site_body[site];
    R total = 0;
    for (int p = 0; p < m.npair; p++) {
      int b1 = m.geom_body[m.pair_g1[p]], b2 = m.geom_body[m.pair_g2[p]];
      if (b1 != body && b2 != body) continue;
      int ncon = (m.pair_kind[p] == 1) ? 2 : 1;
      for (int e = 0; e < ncon; e++) {
        int c = m.pair_conadr[p] + e;
        if (!(d.con_dist[c] < R(0))) continue;
        R normal;
        if (m.pair_condim[p] == 1) normal = d.efc_force[m.pair_efcadr[p] + e];
        else {
          int r0 = m.pair_efcadr[p] + 4 * e;
          normal = d.efc_force[r0] + d.efc_force[r0 + 1] + d.efc_force[r0 + 2] + d.efc_force[r0 + 3];
        }
        if (!(normal > R(0))) continue;
        R ray[3] = {d.con_frame[c][0] * normal, d.con_frame[c][1] * normal, d.con_frame[c][2] * normal};
        normalize3(ray);
        if (body == b2) { ray[0] = -ray[0]; ray[1] = -ray[1]; ray[2] = -ray[2]; }
        R dp[3] = {d.con_pos[c][0] - d.site_xpos[site][0], d.con_pos[c][1] - d.site_xpos[site][1], d.con_pos[c][2] - d.site_xpos[site][2]};
        R lp[3], lv[3];
        matT_vec(lp, d.site_xmat[site], dp);
        matT_vec(lv, d.site_xmat[site], ray);
        R sz[3] = {R(m.site_size[site][0]), R(m.site_size[site][1]), R(m.site_size[site][2])};
        if (ray_box(sz, lp, lv) >= R(0)) total += normal;
      }
    }
    d.sensordata[s] = total;
  }
}

// ---------------------------------------------------------------- forward / step (mjx forward.py)
template <class R> void forward(const mjxb_model_blob& m, Data<R>& d) {
  kinematics(m, d);
  com_pos(m, d);
  tendon(m, d);
  crb(m, d);
  chol_factor(d.qL, d.qM, m.nv);  // factor_m (dense: jacobian="dense")
  d.flops += 6600;
  collision(m, d);
  make_constraint(m, d);
  com_vel(m, d);
  passive(m, d);
  rne(m, d);
  fwd_actuation(m, d);
  fwd_acceleration(m, d);
  solve(m, d);
  sensor_touch(m, d);
  {  // activity-aware fixed part: kinematics 4K, com/cinert/cdof 3K, crb 3K, tree-sparse factor_m 2.6K + solve 0.5K, 108 pair tests 13K,
     // rows of the candidates only (27*12+60 each) + their J*qvel (54 each), com_vel/rne 6K, sensor/obs/reward 1K, implicit integrate 3K
    long ncand = 0;
    for (int r = 0; r < m.nefc; r++) ncand += d.efc_cand[r] ? 1 : 0;
    d.flops_act += 4000 + 3000 + 3000 + 3100 + 13000 + ncand * (27 * 12 + 60 + 54) + 6000 + 1000 + 3000;
  }
}
template <class R> void integrate(const mjxb_model_blob& m, Data<R>& d) {
  // implicitfast (qDeriv = -diag(damping)) and Euler with eulerdamp solve the same system for this model family
  R qacc[MJXB_MAXDOF];
  bool damp = (m.integrator == 3) || (m.integrator == 0 && m.eulerdamp);
  if (damp) {
    static thread_local R A[MJXB_MAXDOF][MJXB_MAXDOF], L[MJXB_MAXDOF][MJXB_MAXDOF];
    for (int i = 0; i < m.nv; i++) {
      for (int j = 0; j < m.nv; j++) A[i][j] = d.qM[i][j];
      A[i][i] += R(m.timestep) * R(m.dof_damping[i]);
    }
    R f[MJXB_MAXDOF];
    for (int i = 0; i < m.nv; i++) f[i] = d.qfrc_smooth[i] + d.qfrc_constraint[i];
    chol_factor(L, A, m.nv);
    chol_solve(qacc, L, f, m.nv);
    d.flops += 8000;
  } else {
    for (int i = 0; i < m.nv; i++) qacc[i] = d.qacc[i];
  }
  R h = R(m.timestep);
  for (int i = 0; i < m.nv; i++) d.qvel[i] += qacc[i] * h;
  for (int j = 0; j < m.njnt; j++) {
    int qa = m.jnt_qposadr[j], da = m.jnt_dofadr[j];
    if (m.jnt_type[j] == 0) {
      for (int k = 0; k < 3; k++) d.qpos[qa + k] += h * d.qvel[da + k];
      R v[3] = {d.qvel[da + 3], d.qvel[da + 4], d.qvel[da + 5]};
      R nrm = normalize3(v);
      R qr[4], qn[4];
      axis_angle_quat(qr, v, h * nrm);
      quat_mul(qn, d.qpos + qa + 3, qr);
      R n = std::sqrt(qn[0] * qn[0] + qn[1] * qn[1] + qn[2] * qn[2] + qn[3] * qn[3]);
      R dn = n + (n == R(0) ? R(1e-6) : R(0));
      for (int k = 0; k < 4; k++) d.qpos[qa + 3 + k] = qn[k] / dn;
    } else {
      d.qpos[qa] += h * d.qvel[da];
    }
  }
  d.time += h;
}
template <class R> void step(const mjxb_model_blob& m, Data<R>& d) {
  forward(m, d);
  integrate(m, d);
}

// ---------------------------------------------------------------- threefry2x32 + the jax.random pieces single_reset uses
inline void threefry2x32(uint32_t k0, uint32_t k1, uint32_t c0, uint32_t c1, uint32_t* o0, uint32_t* o1) {
  static const int R0[4] = {13, 15, 26, 6}, R1[4] = {17, 29, 16, 24};
  uint32_t ks[3] = {k0, k1, k0 ^ k1 ^ 0x1BD11BDAu};
  uint32_t x0 = c0 + ks[0], x1 = c1 + ks[1];
  auto rotl = [](uint32_t v, int r) { return (v << r) | (v >> (32 - r)); };
  for (int i = 0; i < 5; i++) {
    const int* rr = (i % 2 == 0) ? R0 : R1;
    for (int j = 0; j < 4; j++) { x0 += x1; x1 = rotl(x1, rr[j]); x1 ^= x0; }
    x0 += ks[(i + 1) % 3];
    x1 += ks[(i + 2) % 3] + uint32_t(i + 1);
  }
  *o0 = x0; *o1 = x1;
}
// jax.random.split(key, n)[i] with jax_threefry_partitionable=True (default in jax 0.7.2)
inline void jax_split(const uint32_t key[2], int i, uint32_t out[2]) { threefry2x32(key[0], key[1], 0u, uint32_t(i), &out[0], &out[1]); }
// jax.random.bits(key, (n,), 32)[i]
inline uint32_t jax_bits(const uint32_t key[2], int i) {
  uint32_t a, b;
  threefry2x32(key[0], key[1], 0u, uint32_t(i), &a, &b);
  return a ^ b;
}
// jax.random.uniform(key, (n,), float32, minval, maxval)[i]
inline float jax_uniform(const uint32_t key[2], int i, float minval, float maxval) {
  uint32_t bits = (jax_bits(key, i) >> 9) | 0x3F800000u;
  float f;
  std::memcpy(&f, &bits, 4);
  f = f - 1.0f;
  f = f * (maxval - minval) + minval;
  return std::max(minval, f);
}

// ---------------------------------------------------------------- env layer (src/envs.py), transcribed per SURVEY.md Appendix D
template <class R> struct EnvOut { R obs[MJXB_MAXOBS]; R reward, terminated, truncated; };

template <class R> inline R stance_state(const mjxb_env_config& cfg, const R* sens) {  // src/envs.py:89-106
  bool r = sens[cfg.touch_sensor_right_id] > R(0), l = sens[cfg.touch_sensor_left_id] > R(0);
  return (r && l) ? R(0) : (r && !l) ? R(1) : (!r && l) ? R(2) : R(3);
}
template <class R> inline void rpy_from_quat(const R* q, R& roll, R& pitch, R& yaw) {  // src/envs.py:357-366
  R w = q[0], x = q[1], y = q[2], z = q[3];
  roll = std::atan2(R(2) * (w * x + y * z), R(1) - R(2) * (x * x + y * y));
  R sinp = R(2) * (w * y - z * x);
  pitch = std::asin(std::min(std::max(sinp, R(-1)), R(1)));
  yaw = std::atan2(R(2) * (w * z + x * y), R(1) - R(2) * (y * y + z * z));
}
template <class R>
inline void compute_obs(const mjxb_model_blob& m, const mjxb_env_config& cfg, const Data<R>& d, R flip, R height, R roll, R pitch,
                        R yaw, const R* pelvis_quat, const R* tgt, R* obs) {  // src/envs.py:274-331
  R raw[MJXB_MAXOBS];
  int o = 0;
  raw[o++] = height; raw[o++] = roll; raw[o++] = pitch; raw[o++] = yaw;
  for (int i = 7; i < m.nq; i++) raw[o++] = d.qpos[i];
  R w = pelvis_quat[0], x = pelvis_quat[1], y = pelvis_quat[2], z = pelvis_quat[3];
  R xx = x * x, yy = y * y, zz = z * z, xy = x * y, xz = x * z, yz = y * z, wx = w * x, wy = w * y, wz = w * z;
  R r00 = R(1) - R(2) * (yy + zz), r01 = R(2) * (xy - wz), r02 = R(2) * (xz + wy);
  R r10 = R(2) * (xy + wz), r11 = R(1) - R(2) * (xx + zz), r12 = R(2) * (yz - wx);
  R r20 = R(2) * (xz - wy), r21 = R(2) * (yz + wx), r22 = R(1) - R(2) * (xx + yy);
  for (int g = 0; g < 2; g++) {
    R lx = d.qvel[3 * g], ly = d.qvel[3 * g + 1], lz = d.qvel[3 * g + 2];
    raw[o++] = r00 * lx + r10 * ly + r20 * lz;
    raw[o++] = r01 * lx + r11 * ly + r21 * lz;
    raw[o++] = r02 * lx + r12 * ly + r22 * lz;
  }
  for (int i = 6; i < m.nv; i++) raw[o++] = d.qvel[i];
  raw[o++] = tgt[0]; raw[o++] = tgt[1];
  for (int i = 0; i < cfg.obs_dim; i++) obs[i] = flip > R(0.5) ? raw[cfg.obs_perm[i]] * R(cfg.obs_sign[i]) : raw[i];
}

// single_reset (src/envs.py:115-202). `d` receives the post-forward state, aux[9], obs.
template <class R>
void env_reset(const mjxb_model_blob& m, const mjxb_env_config& cfg, const uint32_t key[2], Data<R>& d, R* aux, R* obs) {
  uint32_t k1[2], k2[2], k3[2], k4[2];
  jax_split(key, 0, k1); jax_split(key, 1, k2); jax_split(key, 2, k3); jax_split(key, 3, k4);
  R flip = 0;
  if (cfg.random_flip) flip = jax_uniform(k3, 0, 0.0f, 1.0f) < 0.5f ? R(1) : R(0);  // random.bernoulli(k3, 0.5)
  for (int i = 0; i < m.nq; i++) d.qpos[i] = R(m.qpos0[i]);
  for (int i = 7; i < m.nq; i++) {
    float nz = jax_uniform(k1, i - 7, 0.0f, 1.0f) * 2.0f - 1.0f;
    d.qpos[i] = R(float(m.qpos0[i]) + cfg.random_joint_noise * nz);
  }
  for (int i = 0; i < m.nv; i++) {
    float nz = jax_uniform(k2, i, 0.0f, 1.0f) * 2.0f - 1.0f;
    d.qvel[i] = R(0.0f + cfg.random_vel_noise * nz);
  }
  for (int i = 0; i < m.nu; i++) d.ctrl[i] = 0;
  for (int i = 0; i < m.nv; i++) d.qacc_warmstart[i] = 0;
  d.time = 0;
  // first single_pipeline_init only feeds xpos[pelvis], which does not depend on qvel
  kinematics(m, d);
  R bx = d.xpos[cfg.pelvis_body_id][0], by = d.xpos[cfg.pelvis_body_id][1], bz = d.xpos[cfg.pelvis_body_id][2];
  R tx = bx + R(cfg.target_dist), ty = by, tz = bz;
  if (cfg.initial_velocity_max > 0.0f) {
    R dx = tx - bx, dy = ty - by;
    R dist_xy = std::sqrt(dx * dx + dy * dy);
    R vmag = R(jax_uniform(k4, 0, 0.0f, cfg.initial_velocity_max));
    d.qvel[0] = dist_xy > R(1e-6) ? vmag * dx / dist_xy : R(0);
    d.qvel[1] = dist_xy > R(1e-6) ? vmag * dy / dist_xy : R(0);
  }
  forward(m, d);
  const R* bp = d.xpos[cfg.pelvis_body_id];
  const R* hp = d.xpos[cfg.head_body_id];
  R dx_p = tx - bp[0], dy_p = ty - bp[1];
  R dist_p = std::sqrt(dx_p * dx_p + dy_p * dy_p);
  R dx_h = tx - hp[0], dy_h = ty - hp[1];
  R dist_h = std::sqrt(dx_h * dx_h + dy_h * dy_h);
  R dist = std::max(dist_p, dist_h);
  R last_pot = -dist / R(m.timestep);
  R st = stance_state(cfg, d.sensordata);
  aux[0] = flip; aux[1] = tx; aux[2] = ty; aux[3] = tz; aux[4] = 0; aux[5] = st; aux[6] = d.time; aux[7] = last_pot; aux[8] = 0;
  R roll, pitch, yaw;
  const R* q = d.xquat[cfg.pelvis_body_id];
  rpy_from_quat(q, roll, pitch, yaw);
  R angle = std::atan2(dy_p, dx_p) - yaw;
  R soft = dist / (R(1) + std::abs(dist));
  R tgt[2] = {soft * std::sin(angle), soft * std::cos(angle)};
  compute_obs(m, cfg, d, flip, bp[2], roll, pitch, yaw, q, tgt, obs);
}

// single_step (src/envs.py:333-492)
template <class R>
void env_step(const mjxb_model_blob& m, const mjxb_env_config& cfg, Data<R>& d, R* aux, const R* action, EnvOut<R>& out) {
  R flip = aux[0];
  for (int i = 0; i < m.nu; i++) {
    R a = flip > R(0.5) ? action[cfg.act_perm[i]] * R(cfg.act_sign[i]) : action[i];
    d.ctrl[i] = std::min(std::max(a, R(-1)), R(1));
  }
  step(m, d);  // xpos/xquat/sensordata/qfrc_actuator are those of the forward pass (pre-integration), qpos/qvel/time post
  const R* hp = d.xpos[cfg.head_body_id];
  const R* bp = d.xpos[cfg.pelvis_body_id];
  R height = bp[2];
  const R* q = d.xquat[cfg.pelvis_body_id];
  R roll, pitch, yaw;
  rpy_from_quat(q, roll, pitch, yaw);
  R tx = aux[1], ty = aux[2], tz = aux[3];
  R dx_p = tx - bp[0], dy_p = ty - bp[1];
  R dist_p = std::sqrt(dx_p * dx_p + dy_p * dy_p);
  R dx_h = tx - hp[0], dy_h = ty - hp[1];
  R dist_h = std::sqrt(dx_h * dx_h + dy_h * dy_h);
  R dist = std::max(dist_p, dist_h);
  R dt = R(m.timestep);
  R progress = (-dist / dt - aux[7]) * R(cfg.progress_weight);
  R pw = 0, st2 = 0;
  int nj = m.nv - 6;
  for (int i = 6; i < m.nv; i++) { pw += std::abs(d.qfrc_actuator[i] * d.qvel[i]); st2 += d.qfrc_actuator[i] * d.qfrc_actuator[i]; }
  R energy = R(cfg.electricity_cost) * (pw / R(nj)) + R(cfg.stall_torque_cost) * (st2 / R(nj));
  bool p_ok = (pitch > R(-0.087)) && (pitch < R(0.174));
  bool r_ok = (roll > R(-0.174)) && (roll < R(0.174));
  R posture = ((p_ok ? R(0) : std::abs(pitch)) + (r_ok ? R(0) : std::abs(roll))) * R(cfg.posture_penalty_weight);
  R tall = R(cfg.tall_bonus_weight) * (height > R(cfg.tall_height_threshold) ? R(1) : R(-1));
  R old_stance = aux[5], last_change = aux[6];
  R new_stance = stance_state(cfg, d.sensordata);
  bool changed = new_stance != old_stance;
  R duration = d.time - last_change;
  R stance_reward = (changed && duration > R(0.1)) ? R(cfg.stance_time_reward_weight) * duration / dt : R(0);
  R stance_upd = changed ? new_stance : old_stance;
  R stance_time = changed ? d.time : last_change;
  bool is_close = dist < R(cfg.target_threshold);
  R close_count = is_close ? aux[4] + R(1) : R(0);
  R target_bonus = is_close ? R(2) : R(0);
  bool advance = close_count >= R(cfg.stop_frames);
  if (advance) { tx = bp[0] + R(cfg.target_dist); ty = bp[1]; tz = bp[2]; close_count = 0; }
  R dx_p2 = tx - bp[0], dy_p2 = ty - bp[1];
  R dist_p2 = std::sqrt(dx_p2 * dx_p2 + dy_p2 * dy_p2);
  R dx_h2 = tx - hp[0], dy_h2 = ty - hp[1];
  R dist_h2 = std::sqrt(dx_h2 * dx_h2 + dy_h2 * dy_h2);
  R dist_new = std::max(dist_p2, dist_h2);
  R angle_new = std::atan2(dy_p2, dx_p2) - yaw;
  R soft = dist_new / (R(1) + std::abs(dist_new));
  R tgt[2] = {soft * std::sin(angle_new), soft * std::cos(angle_new)};
  R pot_new = -dist_new / dt;
  R reward = progress + target_bonus + stance_reward - energy + tall - posture - R(0);
  R ep = aux[8] + R(1);
  bool fallen = height < R(cfg.terminate_height);
  out.terminated = fallen ? R(1) : R(0);
  out.truncated = (cfg.max_episode_steps > 0 && ep >= R(cfg.max_episode_steps)) ? R(1) : R(0);
  if (fallen) reward = reward + R(cfg.terminate_reward);
  out.reward = reward;
  aux[0] = flip; aux[1] = tx; aux[2] = ty; aux[3] = tz; aux[4] = close_count; aux[5] = stance_upd; aux[6] = stance_time;
  aux[7] = pot_new; aux[8] = ep;
  compute_obs(m, cfg, d, flip, height, roll, pitch, yaw, q, tgt, out.obs);
}

}  // namespace orc
