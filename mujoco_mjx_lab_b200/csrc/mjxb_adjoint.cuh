// mjxb_adjoint.cuh -- reverse mode (vector-Jacobian product) of one env step, for analytic policy gradients.
//
// Replaces what JAX's autodiff does to `v_step` inside reference train_apg.py:161-209 (value_and_grad through
// lax.scan(jax.checkpoint(body))): the checkpointed body re-runs the step forward and then runs its transpose. Here the tape is the
// step's inputs plus the solver's qacc (= the output state's qacc_warmstart); this file recomputes the forward quantities from them and
// applies the adjoint.  SURVEY.md 8f rank 3 / BASELINE.json configs[3].
//
// What is differentiated:
//   (q', v', aux'[1,2,3,7], reward) = step(q, v, aux[1,2,3,7], action)     with  v' = v + h x,  q' = q (+) h v',
//   x = (M + h D)^-1 (qfrc_smooth + J^T f),  qacc a = argmin of MJX's convex constraint cost  <=>  M a - qfrc_smooth - J^T f(a) = 0.
// The solve is differentiated by the implicit function theorem at the converged solution (one extra solve with H = M + J^T D_act J)
// instead of through unrolled Newton / CG iterations; the active set is held fixed (it is piecewise constant). Rigid-body terms use
// spatial algebra about the fixed world point P = the current subtree COM (MJX's cdof / cinert reference point): a function of body poses
// has the tangent-space gradient  dL/dq_d = S_d . sum_{b in subtree(d)} wrench_b  (S_d = cdof_d), joint axes obey dS_e/dq_d = S_d x S_e for
// d acting on e, inertias dI_b/dq_d = S_d x* I_b - I_b S_d x.  Gradients with respect to the free joint's quaternion COMPONENTS are
// recovered from the body-frame tangent gradient g_w as (2/|q|) q^ (x) (0, g_w) (MJX normalises the quaternion inside kinematics).
// Not differentiated (piecewise constant): contact candidate / active sets, stance state, target advance, termination, clip saturation.
//
// Single source for host and device: written bulk-synchronously over a per-env workspace -- loops `MJA_FOR(i, n)` distribute items over
// the lanes of one warp on the GPU and run sequentially on the CPU; `X.sync()` is __syncwarp() / nothing. The CPU build exists so that
// the mathematics can be checked against finite differences of the float64 oracle without a GPU (tests/adjoint_host.cpp,
// tests/test_adjoint_cpu.py); the GPU build is checked against the same finite differences and against the CPU build.
#pragma once
#include <math.h>
#include <stdint.h>

#include "mjxb.h"
#include "mjxb_model_dev.h"

#if defined(__CUDACC__)
#define MJA_HD __host__ __device__ __forceinline__
#define MJA_HDN __host__ __device__ __noinline__
#else
#define MJA_HD inline
#define MJA_HDN inline
#endif

namespace mjxb {
namespace adj {

struct Lanes {
  int lane, n;
  MJA_HD void sync() const {
#if defined(__CUDA_ARCH__)
    __syncwarp();
#endif
  }
};
#define MJA_FOR(i, count) for (int i = X.lane; i < (count); i += X.n)

constexpr int NB = MJXB_MAXBODY, NG = MJXB_MAXGEOM, NJ = MJXB_MAXJNT;
enum { VJP_OK = 0, VJP_OVERFLOW = 1, VJP_NONFINITE = 2 };

// ------------------------------------------------------------------------------------------- small math (templated: float on the GPU)
template <class T> MJA_HD T dot3(const T* a, const T* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
template <class T> MJA_HD T dot6(const T* a, const T* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2] + a[3] * b[3] + a[4] * b[4] + a[5] * b[5]; }
template <class T> MJA_HD void cross3(T* r, const T* a, const T* b) {
  T x = a[1] * b[2] - a[2] * b[1], y = a[2] * b[0] - a[0] * b[2], z = a[0] * b[1] - a[1] * b[0];
  r[0] = x; r[1] = y; r[2] = z;
}
template <class T> MJA_HD T clampT(T x, T lo, T hi) { return x < lo ? lo : (x > hi ? hi : x); }
template <class T> MJA_HD T absT(T x) { return x < T(0) ? -x : x; }
template <class T> MJA_HD T maxT(T a, T b) { return a > b ? a : b; }
template <class T> MJA_HD void quat_mul(T* r, const T* a, const T* b) {
  T w = a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3];
  T x = a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2];
  T y = a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1];
  T z = a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0];
  r[0] = w; r[1] = x; r[2] = y; r[3] = z;
}
template <class T> MJA_HD void quat_conj(T* r, const T* a) { r[0] = a[0]; r[1] = -a[1]; r[2] = -a[2]; r[3] = -a[3]; }
template <class T> MJA_HD void rotq(T* r, const T* v, const T* q) {
  T s = q[0];
  const T* u = q + 1;
  T uv = dot3(u, v), uu = dot3(u, u), c[3];
  cross3(c, u, v);
  for (int k = 0; k < 3; k++) r[k] = T(2) * (uv * u[k]) + (s * s - uu) * v[k] + T(2) * s * c[k];
}
template <class T> MJA_HD void quat_to_mat(T* m, const T* q) {
  T w = q[0], x = q[1], y = q[2], z = q[3];
  m[0] = w * w + x * x - y * y - z * z; m[1] = T(2) * (x * y - w * z); m[2] = T(2) * (x * z + w * y);
  m[3] = T(2) * (x * y + w * z); m[4] = w * w - x * x + y * y - z * z; m[5] = T(2) * (y * z - w * x);
  m[6] = T(2) * (x * z - w * y); m[7] = T(2) * (y * z + w * x); m[8] = w * w - x * x - y * y + z * z;
}
// spatial algebra on [angular; linear] 6-vectors and MJX's 10-number inertias [xx yy zz xy xz yz, m*off(3), m]
template <class T> MJA_HD void inert_mul(T* r, const T* i, const T* v) {
  T c1[3], c2[3];
  cross3(c1, i + 6, v + 3);
  cross3(c2, i + 6, v);
  r[0] = i[0] * v[0] + i[3] * v[1] + i[4] * v[2] + c1[0];
  r[1] = i[3] * v[0] + i[1] * v[1] + i[5] * v[2] + c1[1];
  r[2] = i[4] * v[0] + i[5] * v[1] + i[2] * v[2] + c1[2];
  r[3] = i[9] * v[3] - c2[0]; r[4] = i[9] * v[4] - c2[1]; r[5] = i[9] * v[5] - c2[2];
}
template <class T> MJA_HD void mcross(T* r, const T* u, const T* v) {  // motion x motion
  T a[3], b[3], c[3];
  cross3(a, u, v); cross3(b, u, v + 3); cross3(c, u + 3, v);
  for (int k = 0; k < 3; k++) { r[k] = a[k]; r[3 + k] = b[k] + c[k]; }
}
template <class T> MJA_HD void mcrossf(T* r, const T* v, const T* f) {  // motion x* force
  T a[3], b[3], c[3];
  cross3(a, v, f); cross3(b, v + 3, f + 3); cross3(c, v, f + 3);
  for (int k = 0; k < 3; k++) { r[k] = a[k] + b[k]; r[3 + k] = c[k]; }
}
template <class T> MJA_HD T normalize3(T* a) {  // mjx math.normalize_with_norm
  T n = sqrt(dot3(a, a));
  T d = n + (n == T(0) ? T(1e-6) : T(0));
  a[0] /= d; a[1] /= d; a[2] /= d;
  return n;
}
// cotangent of u for r = u / |u| (|u| = n > 0, r given)
template <class T> MJA_HD void normalize3_adj(T* ubar, const T* r, T n, const T* rbar) {
  T rr = dot3(r, rbar);
  for (int k = 0; k < 3; k++) ubar[k] = (rbar[k] - r[k] * rr) / n;
}

// ------------------------------------------------------------------------------------------- per-env workspace
template <class T, int CAP_, int MAXCC_>
struct AdjS {
  static constexpr int CAP = CAP_, MAXCC = MAXCC_;
  // step inputs
  T q[32], v[32], a[32], u[32], uclip[32], aux[MJXB_AUX_DIM];
  T qnorm;
  // forward recompute
  T xpos[NB][3], xquat[NB][4], xipos[NB][3], com[3];
  T janchor[NJ][3], jaxis[NJ][3];
  T cinert[NB][10];
  T S[NV][6];                     // cdof
  T V[NV][6], Sd[NV][6], A[NV][6];  // inclusive velocity prefix, cdof_dot, inclusive acceleration prefix (of the current idgrad call)
  union {   // lifetimes do not overlap: the twist prefixes feed the contact-row cotangents, the per-body scratch belongs to idgrad (later)
    struct { T Zv[NV][6], Za[NV][6], Zl[NV][6]; };  // twist prefixes sum_{e<=d} S_e z_e for z = qvel, qacc, lambda
    struct { T tb0[NB][6], tb1[NB][6], tb2[NB][6], tb3[NB][6]; };   // per-body scratch
  };
  T M[NV][NVP];
  T L[NV][NV];
  T col[2][32];                   // GPU Cholesky: column broadcast buffer (double buffered)
  T dinvA[32];                    // 1 / diag of the factor of M + h D (kept for the second solve with that factor)
  T gpos[NG][3], gaxis[NG][3];
  unsigned char pflag[MJXB_MAXPAIR];
  // contacts
  T cc_n[MAXCC][3], cc_t1[MAXCC][3], cc_t2[MAXCC][3], cc_pos[MAXCC][3], cc_dist[MAXCC], cc_mu[MAXCC];
  int cc_pair[MAXCC], cc_row[MAXCC];
  T cc_w[MAXCC][12];              // pose wrenches (about com) on the two bodies of the contact
  // rows
  T J[CAP][NVP];
  T rpos[CAP], rD[CAP], rb[CAP], rk[CAP], rimp[CAP], rdimp[CAP], rinvw[CAP], raref[CAP], rjar[CAP], rf[CAP], rw[CAP], rposbar[CAP];
  T rphibar[CAP][6];
  int rinfo[CAP], ract[CAP];
  int ncc, nrow, overflow;
  // adjoint state
  T x[32], vnew[32];              // integrator acceleration, post-step velocity
  T xbar[32], abar[32], lam[32], y[32], wv[32];
  T gqt[32], gv[32], ubar[32], gqdirect[32];
  T Sbar[NV][6], Hacc[NV][6], Wb[NB][6];
  union {   // composite inertias are consumed by the mass matrix before any per-dof scratch is written
    struct { T td0[NV][6], td1[NV][6], td2[NV][6]; };  // per-dof scratch
    T crb[NB][10];
  };
  T gauxin[MJXB_AUX_DIM];
  T gbody_pelvis[3], gbody_head[3], gquat_pelvis[4];
  T red[32];
};

// pointers of one env (element type T: float on the GPU, double in the CPU check)
template <class T>
struct EnvIO {
  const T *qpos, *qvel, *aux, *action, *tape_qacc;
  const T *g_qpos_out, *g_qvel_out, *g_aux_out;
  T g_reward;
  T *g_qpos_in, *g_qvel_in, *g_aux_in, *g_action;
};

// ------------------------------------------------------------------------------------------- forward recompute
template <class T, class WS>
MJA_HD void fwd_kinematics(const DevModel& C, WS& W, Lanes X) {
  if (X.lane == 0) {
    for (int j = 0; j < C.njnt; j++)
      if (C.jnt_type[j] == 0) {
        const int qa = C.jnt_qposadr[j];
        T* qq = &W.q[qa + 3];
        T n = sqrt(qq[0] * qq[0] + qq[1] * qq[1] + qq[2] * qq[2] + qq[3] * qq[3]);
        W.qnorm = n;
        T dn = n + (n == T(0) ? T(1e-6) : T(0));
        for (int k = 0; k < 4; k++) qq[k] /= dn;
      }
    for (int k = 0; k < 3; k++) W.xpos[0][k] = T(0);
    W.xquat[0][0] = T(1); W.xquat[0][1] = W.xquat[0][2] = W.xquat[0][3] = T(0);
  }
  X.sync();
  int maxdepth = 0;
  for (int b = 0; b < C.nbody; b++) maxdepth = C.body_depth[b] > maxdepth ? C.body_depth[b] : maxdepth;
  for (int lvl = 1; lvl <= maxdepth; lvl++) {
    MJA_FOR(b, C.nbody) {
      if (C.body_depth[b] != lvl) continue;
      const int p = C.body_parent[b];
      T bp[3] = {T(C.body_pos[b][0]), T(C.body_pos[b][1]), T(C.body_pos[b][2])};
      T bq[4] = {T(C.body_quat[b][0]), T(C.body_quat[b][1]), T(C.body_quat[b][2]), T(C.body_quat[b][3])};
      T pos[3], quat[4], t[3];
      rotq(t, bp, W.xquat[p]);
      for (int k = 0; k < 3; k++) pos[k] = W.xpos[p][k] + t[k];
      quat_mul(quat, W.xquat[p], bq);
      for (int j = C.body_jntadr[b]; j < C.body_jntadr[b] + C.body_jntnum[b]; j++) {
        const int qa = C.jnt_qposadr[j];
        T jp[3] = {T(C.jnt_pos[j][0]), T(C.jnt_pos[j][1]), T(C.jnt_pos[j][2])};
        T ja[3] = {T(C.jnt_axis[j][0]), T(C.jnt_axis[j][1]), T(C.jnt_axis[j][2])};
        if (C.jnt_type[j] == 0) {
          for (int k = 0; k < 3; k++) { pos[k] = W.q[qa + k]; W.janchor[j][k] = pos[k]; }
          W.jaxis[j][0] = T(0); W.jaxis[j][1] = T(0); W.jaxis[j][2] = T(1);
          for (int k = 0; k < 4; k++) quat[k] = W.q[qa + 3 + k];
        } else {
          rotq(t, jp, quat);
          for (int k = 0; k < 3; k++) W.janchor[j][k] = t[k] + pos[k];
          rotq(W.jaxis[j], ja, quat);
          const T ang = W.q[qa] - T(C.qpos0[qa]);
          const T sn = sin(ang * T(0.5)), cs = cos(ang * T(0.5));
          T ql[4] = {cs, ja[0] * sn, ja[1] * sn, ja[2] * sn}, qn[4];
          quat_mul(qn, quat, ql);
          for (int k = 0; k < 4; k++) quat[k] = qn[k];
          rotq(t, jp, quat);
          for (int k = 0; k < 3; k++) pos[k] = W.janchor[j][k] - t[k];
        }
      }
      for (int k = 0; k < 3; k++) W.xpos[b][k] = pos[k];
      for (int k = 0; k < 4; k++) W.xquat[b][k] = quat[k];
    }
    X.sync();
  }
}

template <class T, class WS>
MJA_HD void fwd_com_cdof(const DevModel& C, WS& W, Lanes X) {
  MJA_FOR(b, C.nbody) {
    T ip[3] = {T(C.body_ipos[b][0]), T(C.body_ipos[b][1]), T(C.body_ipos[b][2])}, t[3];
    rotq(t, ip, W.xquat[b]);
    for (int k = 0; k < 3; k++) W.xipos[b][k] = W.xpos[b][k] + t[k];
  }
  X.sync();
  if (X.lane == 0) {
    T s[3] = {T(0), T(0), T(0)};
    for (int b = 0; b < C.nbody; b++)
      for (int k = 0; k < 3; k++) s[k] += T(C.body_mass[b]) * W.xipos[b][k];
    const T inv = T(1) / maxT(T(C.total_mass), T(1e-15));
    for (int k = 0; k < 3; k++) W.com[k] = s[k] * inv;
  }
  X.sync();
  MJA_FOR(b, C.nbody) {
    T R[9];
    quat_to_mat(R, W.xquat[b]);
    const T mass = T(C.body_mass[b]);
    T off[3] = {W.xipos[b][0] - W.com[0], W.xipos[b][1] - W.com[1], W.xipos[b][2] - W.com[2]};
    const float* bi = C.body_inertia[b];
    T I[9] = {T(bi[0]), T(bi[3]), T(bi[4]), T(bi[3]), T(bi[1]), T(bi[5]), T(bi[4]), T(bi[5]), T(bi[2])}, XI[9], Iw[9];
    for (int r = 0; r < 3; r++)
      for (int c = 0; c < 3; c++) XI[3 * r + c] = R[3 * r] * I[c] + R[3 * r + 1] * I[3 + c] + R[3 * r + 2] * I[6 + c];
    for (int r = 0; r < 3; r++)
      for (int c = 0; c < 3; c++) Iw[3 * r + c] = XI[3 * r] * R[3 * c] + XI[3 * r + 1] * R[3 * c + 1] + XI[3 * r + 2] * R[3 * c + 2];
    const T oo = dot3(off, off);
    T* ci = W.cinert[b];
    ci[0] = Iw[0] + mass * (oo - off[0] * off[0]);
    ci[1] = Iw[4] + mass * (oo - off[1] * off[1]);
    ci[2] = Iw[8] + mass * (oo - off[2] * off[2]);
    ci[3] = Iw[1] - mass * off[0] * off[1];
    ci[4] = Iw[2] - mass * off[0] * off[2];
    ci[5] = Iw[5] - mass * off[1] * off[2];
    ci[6] = mass * off[0]; ci[7] = mass * off[1]; ci[8] = mass * off[2];
    ci[9] = mass;
    if (b == 0) for (int k = 0; k < 10; k++) ci[k] = T(0);
  }
  MJA_FOR(d, NV) {
    const int j = C.dof_jnt[d], b = C.dof_body[d];
    T* cd = W.S[d];
    if (C.jnt_type[j] == 0) {
      const int k = d - C.jnt_dofadr[j];
      if (k < 3) {
        for (int i = 0; i < 6; i++) cd[i] = T(0);
        cd[3 + k] = T(1);
      } else {
        T R[9];
        quat_to_mat(R, W.xquat[b]);
        T ax[3] = {R[k - 3], R[3 + k - 3], R[6 + k - 3]};
        T off[3] = {W.com[0] - W.xpos[b][0], W.com[1] - W.xpos[b][1], W.com[2] - W.xpos[b][2]};
        cd[0] = ax[0]; cd[1] = ax[1]; cd[2] = ax[2];
        cross3(cd + 3, ax, off);
      }
    } else {
      T off[3] = {W.com[0] - W.janchor[j][0], W.com[1] - W.janchor[j][1], W.com[2] - W.janchor[j][2]};
      cd[0] = W.jaxis[j][0]; cd[1] = W.jaxis[j][1]; cd[2] = W.jaxis[j][2];
      cross3(cd + 3, W.jaxis[j], off);
    }
  }
  X.sync();
}

// out[d] = sum over the chain root..d of S_e z_e
template <class T, class WS>
MJA_HD void twist_prefix(const DevModel& C, WS& W, Lanes X, const T* z, T (*out)[6]) {
  MJA_FOR(d, NV) {
    T s[6] = {T(0), T(0), T(0), T(0), T(0), T(0)};
    for (int e = d; e >= 0; e = C.dof_parent[e])
      for (int k = 0; k < 6; k++) s[k] += W.S[e][k] * z[e];
    for (int k = 0; k < 6; k++) out[d][k] = s[k];
  }
  X.sync();
}

// V (velocity prefix), Sd (cdof_dot), A (acceleration prefix incl. -gravity) for generalised velocity vv and acceleration aa
template <class T, class WS>
MJA_HD void fwd_vel_acc(const DevModel& C, WS& W, Lanes X, const T* vv, const T* aa, bool grav) {
  twist_prefix<T>(C, W, X, vv, W.V);
  MJA_FOR(d, NV) {
    const int src = C.dof_cvel_src[d];
    if (src == -2) {
      for (int k = 0; k < 6; k++) W.Sd[d][k] = T(0);
    } else {
      T vs[6] = {T(0), T(0), T(0), T(0), T(0), T(0)};
      if (src >= 0) for (int k = 0; k < 6; k++) vs[k] = W.V[src][k];
      mcross(W.Sd[d], vs, W.S[d]);
    }
  }
  X.sync();
  MJA_FOR(d, NV) {
    T s[6] = {T(0), T(0), T(0), grav ? -T(C.gravity[0]) : T(0), grav ? -T(C.gravity[1]) : T(0), grav ? -T(C.gravity[2]) : T(0)};
    for (int e = d; e >= 0; e = C.dof_parent[e])
      for (int k = 0; k < 6; k++) s[k] += W.S[e][k] * aa[e] + W.Sd[e][k] * vv[e];
    for (int k = 0; k < 6; k++) W.A[d][k] = s[k];
  }
  X.sync();
}

template <class T, class WS>
MJA_HD void fwd_mass_matrix(const DevModel& C, WS& W, Lanes X) {
  MJA_FOR(b, C.nbody) {
    T s[10];
    for (int k = 0; k < 10; k++) s[k] = T(0);
    if (b >= 1)
      for (int c = b; c < C.body_subtree_end[b]; c++)
        for (int k = 0; k < 10; k++) s[k] += W.cinert[c][k];
    for (int k = 0; k < 10; k++) W.crb[b][k] = s[k];
  }
  MJA_FOR(i, NV * NVP) (&W.M[0][0])[i] = T(0);
  X.sync();
  MJA_FOR(i, NV) {
    T f[6];
    inert_mul(f, W.crb[C.dof_body[i]], W.S[i]);
    for (int j = i; j >= 0; j = C.dof_parent[j]) {
      T s = dot6(W.S[j], f);
      if (j == i) s += T(C.dof_armature[i]);
      W.M[i][j] = s;
      W.M[j][i] = s;
    }
  }
  X.sync();
}

// ------------------------------------------------------------------------------------------- contact geometry, forward and adjoint
template <class T> MJA_HD void make_tangents(const T* n, T* t1, T* t2, T* b0_out, T* unorm_out) {
  T b[3] = {T(0), T(0), T(0)};
  if (T(-0.5) < n[1] && n[1] < T(0.5)) b[1] = T(1); else b[2] = T(1);
  if (b0_out) { b0_out[0] = b[0]; b0_out[1] = b[1]; b0_out[2] = b[2]; }
  const T ab = dot3(n, b);
  b[0] -= n[0] * ab; b[1] -= n[1] * ab; b[2] -= n[2] * ab;
  const T un = normalize3(b);
  if (unorm_out) *unorm_out = un;
  t1[0] = b[0]; t1[1] = b[1]; t1[2] = b[2];
  cross3(t2, n, b);
}
// cotangent of n through t1 = normalize(b0 - n (n.b0)), t2 = n x t1
template <class T> MJA_HD void make_tangents_adj(const T* n, const T* t1, const T* b0, T un, const T* t1bar_in, const T* t2bar, T* nbar) {
  T t1bar[3] = {t1bar_in[0], t1bar_in[1], t1bar_in[2]}, c[3];
  cross3(c, t1, t2bar);   // t2 = n x t1 : nbar += t1 x t2bar, t1bar += t2bar x n
  for (int k = 0; k < 3; k++) nbar[k] += c[k];
  cross3(c, t2bar, n);
  for (int k = 0; k < 3; k++) t1bar[k] += c[k];
  if (un > T(0)) {
    T ub[3];
    normalize3_adj(ub, t1, un, t1bar);
    const T nb0 = dot3(n, b0), nub = dot3(n, ub);
    for (int k = 0; k < 3; k++) nbar[k] += -nb0 * ub[k] - b0[k] * nub;
  }
}
template <class T> MJA_HD void closest_segment_point(T* out, T& t_out, bool& free_t, const T* a, const T* b, const T* pt) {
  T ab[3] = {b[0] - a[0], b[1] - a[1], b[2] - a[2]}, pa[3] = {pt[0] - a[0], pt[1] - a[1], pt[2] - a[2]};
  const T t0 = dot3(pa, ab) / (dot3(ab, ab) + T(1e-6));
  const T t = clampT(t0, T(0), T(1));
  free_t = (t0 > T(0)) && (t0 < T(1));
  t_out = t;
  out[0] = a[0] + t * ab[0]; out[1] = a[1] + t * ab[1]; out[2] = a[2] + t * ab[2];
}
template <class T> MJA_HD void closest_segment_point_adj(const T* a, const T* b, const T* pt, const T* outbar, T* abar, T* bbar, T* ptbar) {
  T ab[3] = {b[0] - a[0], b[1] - a[1], b[2] - a[2]}, pa[3] = {pt[0] - a[0], pt[1] - a[1], pt[2] - a[2]};
  const T num = dot3(pa, ab), den = dot3(ab, ab) + T(1e-6);
  const T t0 = num / den, t = clampT(t0, T(0), T(1));
  T abbar[3] = {t * outbar[0], t * outbar[1], t * outbar[2]};
  for (int k = 0; k < 3; k++) abar[k] += outbar[k];
  if (t0 > T(0) && t0 < T(1)) {
    const T tbar = dot3(ab, outbar);
    const T numbar = tbar / den, denbar = -tbar * t0 / den;
    for (int k = 0; k < 3; k++) {
      ptbar[k] += numbar * ab[k];
      abar[k] -= numbar * ab[k];
      abbar[k] += numbar * pa[k] + T(2) * denbar * ab[k];
    }
  }
  for (int k = 0; k < 3; k++) { bbar[k] += abbar[k]; abar[k] -= abbar[k]; }
}
// sphere_sphere(pa, r1, pb, r2): n = (pb - pa)/len, dist = len - r1 - r2, pos = pa + n (r1 + dist/2)
template <class T> MJA_HD void sphere_sphere(T& dist, T* pos, T* n, T& len, const T* pa, T r1, const T* pb, T r2) {
  n[0] = pb[0] - pa[0]; n[1] = pb[1] - pa[1]; n[2] = pb[2] - pa[2];
  len = normalize3(n);
  if (len == T(0)) { n[0] = T(1); n[1] = T(0); n[2] = T(0); }
  dist = len - (r1 + r2);
  const T s = r1 + dist * T(0.5);
  for (int k = 0; k < 3; k++) pos[k] = pa[k] + n[k] * s;
}
template <class T>
MJA_HD void sphere_sphere_adj(const T* n, T len, T dist, T r1, T distbar, const T* posbar, const T* nbar_in, T* pabar, T* pbbar) {
  T nbar[3];
  const T s = r1 + dist * T(0.5);
  for (int k = 0; k < 3; k++) { nbar[k] = nbar_in[k] + posbar[k] * s; pabar[k] += posbar[k]; }
  const T lenbar = distbar + T(0.5) * dot3(n, posbar);
  if (len > T(0)) {
    T db[3];
    normalize3_adj(db, n, len, nbar);
    for (int k = 0; k < 3; k++) { db[k] += lenbar * n[k]; pbbar[k] += db[k]; pabar[k] -= db[k]; }
  }
}

// Contact `e` of geom pair p (mjx collision_primitive): forward values.
template <class T, class WS>
MJA_HD void contact_fwd(const DevModel& C, const WS& W, int p, int e, T& dist, T* pos, T* n, T* t1, T* t2) {
  const uint32_t w0 = C.pair_w0[p];
  const int g1 = w0 & 0xff, g2 = (w0 >> 8) & 0xff, kind = (w0 >> 16) & 0xff, condim = (w0 >> 24) & 0x7f;
  const T* p1 = W.gpos[g1]; const T* p2 = W.gpos[g2]; const T* ax1 = W.gaxis[g1]; const T* ax2 = W.gaxis[g2];
  const T r1 = T(C.geom_rad[g1]), r2 = T(C.geom_rad[g2]), l1 = T(C.geom_half[g1]), l2 = T(C.geom_half[g2]);
  for (int k = 0; k < 3; k++) { t1[k] = T(0); t2[k] = T(0); }
  if (kind == PAIR_PLANE_CAPSULE || kind == PAIR_PLANE_SPHERE) {
    for (int k = 0; k < 3; k++) n[k] = ax1[k];
    T sp[3] = {p2[0], p2[1], p2[2]};
    if (kind == PAIR_PLANE_CAPSULE) {
      const T na = dot3(ax1, ax2);
      T b[3] = {ax2[0] - ax1[0] * na, ax2[1] - ax1[1] * na, ax2[2] - ax1[2] * na};
      const T bn = normalize3(b);
      if (bn < T(0.5)) {
        b[0] = T(0); b[1] = T(0); b[2] = T(0);
        if (T(-0.5) < ax1[1] && ax1[1] < T(0.5)) b[1] = T(1); else b[2] = T(1);
      }
      for (int k = 0; k < 3; k++) t1[k] = b[k];
      cross3(t2, ax1, b);
      const T sg = e == 0 ? T(1) : T(-1);
      for (int k = 0; k < 3; k++) sp[k] = p2[k] + sg * (ax2[k] * l2);
    } else {
      make_tangents<T>(n, t1, t2, nullptr, nullptr);
    }
    T df[3] = {sp[0] - p1[0], sp[1] - p1[1], sp[2] - p1[2]};
    dist = dot3(df, ax1) - r2;
    const T s = r2 + T(0.5) * dist;
    for (int k = 0; k < 3; k++) pos[k] = sp[k] - ax1[k] * s;
    return;
  }
  T pa[3] = {p1[0], p1[1], p1[2]}, pb[3] = {p2[0], p2[1], p2[2]};
  if (kind == PAIR_SPHERE_CAPSULE) {
    T a[3], b[3], tt;
    bool fr;
    for (int k = 0; k < 3; k++) { a[k] = p2[k] - ax2[k] * l2; b[k] = p2[k] + ax2[k] * l2; }
    closest_segment_point(pb, tt, fr, a, b, p1);
  } else if (kind == PAIR_CAPSULE_CAPSULE) {
    T a0[3], a1[3], b0[3], b1[3];
    for (int k = 0; k < 3; k++) {
      a0[k] = p1[k] - ax1[k] * l1; a1[k] = p1[k] + ax1[k] * l1;
      b0[k] = p2[k] - ax2[k] * l2; b1[k] = p2[k] + ax2[k] * l2;
    }
    T dir_a[3] = {a1[0] - a0[0], a1[1] - a0[1], a1[2] - a0[2]}, dir_b[3] = {b1[0] - b0[0], b1[1] - b0[1], b1[2] - b0[2]};
    const T half_a = normalize3(dir_a) * T(0.5), half_b = normalize3(dir_b) * T(0.5);
    T a_mid[3], b_mid[3], trans[3];
    for (int k = 0; k < 3; k++) { a_mid[k] = a0[k] + dir_a[k] * half_a; b_mid[k] = b0[k] + dir_b[k] * half_b; trans[k] = a_mid[k] - b_mid[k]; }
    const T dd = dot3(dir_a, dir_b), da_t = dot3(dir_a, trans), db_t = dot3(dir_b, trans);
    const T denom = T(1) - dd * dd;
    const T orig_ta = (-da_t + dd * db_t) / (denom + T(1e-6));
    const T orig_tb = db_t + orig_ta * dd;
    const T ta = clampT(orig_ta, -half_a, half_a), tb = clampT(orig_tb, -half_b, half_b);
    T best_a[3], best_b[3], new_a[3], new_b[3], tt;
    bool fr;
    for (int k = 0; k < 3; k++) { best_a[k] = a_mid[k] + dir_a[k] * ta; best_b[k] = b_mid[k] + dir_b[k] * tb; }
    closest_segment_point(new_a, tt, fr, a0, a1, best_b);
    closest_segment_point(new_b, tt, fr, b0, b1, best_a);
    T d1 = T(0), d2 = T(0);
    for (int k = 0; k < 3; k++) {
      d1 += (new_a[k] - best_b[k]) * (new_a[k] - best_b[k]);
      d2 += (best_a[k] - new_b[k]) * (best_a[k] - new_b[k]);
    }
    if (d1 < d2) { for (int k = 0; k < 3; k++) { pa[k] = new_a[k]; pb[k] = best_b[k]; } }
    else { for (int k = 0; k < 3; k++) { pa[k] = best_a[k]; pb[k] = new_b[k]; } }
  }
  T len;
  sphere_sphere(dist, pos, n, len, pa, r1, pb, r2);
  if (condim > 1) make_tangents<T>(n, t1, t2, nullptr, nullptr);
}

// Adjoint of contact_fwd: cotangents of (dist, pos, n, t1, t2) -> cotangents of the two geoms' centres and axes (accumulated).
template <class T, class WS>
MJA_HD void contact_adj(const DevModel& C, const WS& W, int p, int e, T distbar, const T* posbar, const T* nbar_in, const T* t1bar,
                        const T* t2bar, T* gp1, T* ga1, T* gp2, T* ga2) {
  const uint32_t w0 = C.pair_w0[p];
  const int g1 = w0 & 0xff, g2 = (w0 >> 8) & 0xff, kind = (w0 >> 16) & 0xff, condim = (w0 >> 24) & 0x7f;
  const T* p1 = W.gpos[g1]; const T* p2 = W.gpos[g2]; const T* ax1 = W.gaxis[g1]; const T* ax2 = W.gaxis[g2];
  const T r1 = T(C.geom_rad[g1]), r2 = T(C.geom_rad[g2]), l1 = T(C.geom_half[g1]), l2 = T(C.geom_half[g2]);
  if (kind == PAIR_PLANE_CAPSULE || kind == PAIR_PLANE_SPHERE) {
    // dist = (sp - p1).ax1 - r2 ; pos = sp - ax1 (r2 + dist/2) ; n = ax1
    T sp[3] = {p2[0], p2[1], p2[2]}, spbar[3], ax1bar[3];
    const T sg = e == 0 ? T(1) : T(-1);
    if (kind == PAIR_PLANE_CAPSULE) for (int k = 0; k < 3; k++) sp[k] = p2[k] + sg * (ax2[k] * l2);
    T df[3] = {sp[0] - p1[0], sp[1] - p1[1], sp[2] - p1[2]};
    const T dist = dot3(df, ax1) - r2, s = r2 + T(0.5) * dist;
    const T dbar = distbar - T(0.5) * dot3(ax1, posbar);
    for (int k = 0; k < 3; k++) {
      spbar[k] = posbar[k] + dbar * ax1[k];
      ax1bar[k] = nbar_in[k] - s * posbar[k] + dbar * df[k];
      gp1[k] -= dbar * ax1[k];
    }
    if (kind == PAIR_PLANE_CAPSULE) {
      // t1 = b = normalize(ax2 - ax1 (ax1.ax2)), t2 = ax1 x b
      const T na = dot3(ax1, ax2);
      T b[3] = {ax2[0] - ax1[0] * na, ax2[1] - ax1[1] * na, ax2[2] - ax1[2] * na};
      const T bn = normalize3(b);
      if (bn >= T(0.5)) {
        T bbar[3], c[3], ub[3];
        cross3(c, t2bar, ax1);                  // t2 = ax1 x b : bbar += t2bar x ax1 ; ax1bar += b x t2bar
        for (int k = 0; k < 3; k++) bbar[k] = t1bar[k] + c[k];
        cross3(c, b, t2bar);
        for (int k = 0; k < 3; k++) ax1bar[k] += c[k];
        normalize3_adj(ub, b, bn, bbar);
        const T a1ub = dot3(ax1, ub);
        for (int k = 0; k < 3; k++) {
          ga2[k] += ub[k] - ax1[k] * a1ub;       // u = ax2 - ax1 (ax1.ax2)
          ax1bar[k] += -na * ub[k] - ax2[k] * a1ub;
        }
      } else {                                  // fixed b: only t2 = ax1 x b depends on ax1
        T bf[3] = {T(0), T(0), T(0)}, c[3];
        if (T(-0.5) < ax1[1] && ax1[1] < T(0.5)) bf[1] = T(1); else bf[2] = T(1);
        cross3(c, bf, t2bar);
        for (int k = 0; k < 3; k++) ax1bar[k] += c[k];
      }
      for (int k = 0; k < 3; k++) { gp2[k] += spbar[k]; ga2[k] += sg * l2 * spbar[k]; }
    } else {
      for (int k = 0; k < 3; k++) gp2[k] += spbar[k];
      T b0[3], un, t1f[3], t2f[3];
      make_tangents<T>(ax1, t1f, t2f, b0, &un);
      make_tangents_adj<T>(ax1, t1f, b0, un, t1bar, t2bar, ax1bar);
    }
    for (int k = 0; k < 3; k++) ga1[k] += ax1bar[k];
    return;
  }
  // sphere-type pairs: (pa, pb) closest points, then sphere_sphere
  T pa[3] = {p1[0], p1[1], p1[2]}, pb[3] = {p2[0], p2[1], p2[2]};
  T a0[3], a1[3], b0[3], b1[3], best_a[3], best_b[3];
  T dir_a[3], dir_b[3], half_a = T(0), half_b = T(0), dd = T(0), da_t = T(0), db_t = T(0), denom = T(1), orig_ta = T(0), orig_tb = T(0), ta = T(0), tb = T(0);
  T trans[3] = {T(0), T(0), T(0)};
  bool choose_new_a = false;
  if (kind == PAIR_SPHERE_CAPSULE) {
    T tt;
    bool fr;
    for (int k = 0; k < 3; k++) { b0[k] = p2[k] - ax2[k] * l2; b1[k] = p2[k] + ax2[k] * l2; }
    closest_segment_point(pb, tt, fr, b0, b1, p1);
  } else if (kind == PAIR_CAPSULE_CAPSULE) {
    for (int k = 0; k < 3; k++) {
      a0[k] = p1[k] - ax1[k] * l1; a1[k] = p1[k] + ax1[k] * l1;
      b0[k] = p2[k] - ax2[k] * l2; b1[k] = p2[k] + ax2[k] * l2;
      dir_a[k] = ax1[k]; dir_b[k] = ax2[k];            // = normalize(a1 - a0), normalize(b1 - b0) for unit axes
      trans[k] = p1[k] - p2[k];                        // a_mid - b_mid
    }
    half_a = l1; half_b = l2;
    dd = dot3(dir_a, dir_b); da_t = dot3(dir_a, trans); db_t = dot3(dir_b, trans);
    denom = T(1) - dd * dd;
    orig_ta = (-da_t + dd * db_t) / (denom + T(1e-6));
    orig_tb = db_t + orig_ta * dd;
    ta = clampT(orig_ta, -half_a, half_a); tb = clampT(orig_tb, -half_b, half_b);
    T new_a[3], new_b[3], tt;
    bool fr;
    for (int k = 0; k < 3; k++) { best_a[k] = p1[k] + dir_a[k] * ta; best_b[k] = p2[k] + dir_b[k] * tb; }
    closest_segment_point(new_a, tt, fr, a0, a1, best_b);
    closest_segment_point(new_b, tt, fr, b0, b1, best_a);
    T d1 = T(0), d2 = T(0);
    for (int k = 0; k < 3; k++) {
      d1 += (new_a[k] - best_b[k]) * (new_a[k] - best_b[k]);
      d2 += (best_a[k] - new_b[k]) * (best_a[k] - new_b[k]);
    }
    choose_new_a = d1 < d2;
    if (choose_new_a) { for (int k = 0; k < 3; k++) { pa[k] = new_a[k]; pb[k] = best_b[k]; } }
    else { for (int k = 0; k < 3; k++) { pa[k] = best_a[k]; pb[k] = new_b[k]; } }
  }
  T dist, pos[3], n[3], len;
  sphere_sphere(dist, pos, n, len, pa, r1, pb, r2);
  T nbar[3] = {nbar_in[0], nbar_in[1], nbar_in[2]};
  if (condim > 1) {
    T b0v[3], un, t1f[3], t2f[3];
    make_tangents<T>(n, t1f, t2f, b0v, &un);
    make_tangents_adj<T>(n, t1f, b0v, un, t1bar, t2bar, nbar);
  }
  T pabar[3] = {T(0), T(0), T(0)}, pbbar[3] = {T(0), T(0), T(0)};
  sphere_sphere_adj<T>(n, len, dist, r1, distbar, posbar, nbar, pabar, pbbar);
  if (kind == PAIR_SPHERE_SPHERE) {
    for (int k = 0; k < 3; k++) { gp1[k] += pabar[k]; gp2[k] += pbbar[k]; }
  } else if (kind == PAIR_SPHERE_CAPSULE) {
    T b0bar[3] = {T(0), T(0), T(0)}, b1bar[3] = {T(0), T(0), T(0)}, ptbar[3] = {T(0), T(0), T(0)};
    closest_segment_point_adj<T>(b0, b1, p1, pbbar, b0bar, b1bar, ptbar);
    for (int k = 0; k < 3; k++) {
      gp1[k] += pabar[k] + ptbar[k];
      gp2[k] += b0bar[k] + b1bar[k];
      ga2[k] += l2 * (b1bar[k] - b0bar[k]);
    }
  } else {  // capsule - capsule
    T a0bar[3] = {T(0), T(0), T(0)}, a1bar[3] = {T(0), T(0), T(0)}, b0bar[3] = {T(0), T(0), T(0)}, b1bar[3] = {T(0), T(0), T(0)};
    T bestabar[3] = {T(0), T(0), T(0)}, bestbbar[3] = {T(0), T(0), T(0)};
    if (choose_new_a) {   // pa = csp(a0, a1, best_b), pb = best_b
      closest_segment_point_adj<T>(a0, a1, best_b, pabar, a0bar, a1bar, bestbbar);
      for (int k = 0; k < 3; k++) bestbbar[k] += pbbar[k];
    } else {              // pa = best_a, pb = csp(b0, b1, best_a)
      closest_segment_point_adj<T>(b0, b1, best_a, pbbar, b0bar, b1bar, bestabar);
      for (int k = 0; k < 3; k++) bestabar[k] += pabar[k];
    }
    // best_a = p1 + dir_a ta ; best_b = p2 + dir_b tb
    T p1bar[3], p2bar[3], dabar[3], dbbar[3];
    T tabar = dot3(dir_a, bestabar), tbbar = dot3(dir_b, bestbbar);
    for (int k = 0; k < 3; k++) {
      p1bar[k] = bestabar[k]; p2bar[k] = bestbbar[k];
      dabar[k] = ta * bestabar[k]; dbbar[k] = tb * bestbbar[k];
    }
    T otabar = (orig_ta > -half_a && orig_ta < half_a) ? tabar : T(0);
    T otbbar = (orig_tb > -half_b && orig_tb < half_b) ? tbbar : T(0);
    // orig_tb = db_t + orig_ta dd
    T dbtbar = otbbar, ddbar = otbbar * orig_ta;
    otabar += otbbar * dd;
    // orig_ta = (-da_t + dd db_t) / (denom + 1e-6), denom = 1 - dd^2
    const T den = denom + T(1e-6);
    const T datbar = -otabar / den;
    dbtbar += otabar * dd / den;
    ddbar += otabar * db_t / den;
    const T denbar = -otabar * orig_ta / den;
    ddbar += denbar * (-T(2) * dd);
    T transbar[3];
    for (int k = 0; k < 3; k++) {
      dabar[k] += ddbar * dir_b[k] + datbar * trans[k];
      dbbar[k] += ddbar * dir_a[k] + dbtbar * trans[k];
      transbar[k] = datbar * dir_a[k] + dbtbar * dir_b[k];
      p1bar[k] += transbar[k]; p2bar[k] -= transbar[k];
    }
    for (int k = 0; k < 3; k++) {
      gp1[k] += p1bar[k] + a0bar[k] + a1bar[k];
      gp2[k] += p2bar[k] + b0bar[k] + b1bar[k];
      ga1[k] += dabar[k] + l1 * (a1bar[k] - a0bar[k]);
      ga2[k] += dbbar[k] + l2 * (b1bar[k] - b0bar[k]);
    }
  }
}

// ------------------------------------------------------------------------------------------- constraint impedance with derivative
template <class T>
MJA_HD void kbi_d(T timestep, const float* solref, const float* solimp, T pos, T& k, T& b, T& imp, T& dimp) {
  const T timeconst = maxT(T(solref[0]), T(2) * timestep), dampratio = T(solref[1]);
  const T dmin = clampT(T(solimp[0]), T(1e-4), T(0.9999)), dmax = clampT(T(solimp[1]), T(1e-4), T(0.9999));
  const T width = maxT(T(1e-15), T(solimp[2])), mid = clampT(T(solimp[3]), T(1e-4), T(0.9999)), power = maxT(T(1), T(solimp[4]));
  k = T(1) / (dmax * dmax * timeconst * timeconst * dampratio * dampratio);
  b = T(2) / (dmax * timeconst);
  if (solref[0] <= 0.0f) k = -T(solref[0]) / (dmax * dmax);
  if (solref[1] <= 0.0f) b = -T(solref[1]) / dmax;
  const T x = absT(pos) / width;
  T y, dy;
  if (power == T(2)) {
    if (x < mid) { y = (T(1) / mid) * (x * x); dy = T(2) * x / mid; }
    else { y = T(1) - (T(1) / (T(1) - mid)) * ((T(1) - x) * (T(1) - x)); dy = T(2) * (T(1) - x) / (T(1) - mid); }
  } else if (power == T(1)) {
    y = x; dy = T(1);
  } else {
    if (x < mid) { y = (T(1) / pow(mid, power - T(1))) * pow(x, power); dy = power * pow(x, power - T(1)) / pow(mid, power - T(1)); }
    else { y = T(1) - (T(1) / pow(T(1) - mid, power - T(1))) * pow(T(1) - x, power); dy = power * pow(T(1) - x, power - T(1)) / pow(T(1) - mid, power - T(1)); }
  }
  imp = clampT(dmin + y * (dmax - dmin), dmin, dmax);
  dimp = dy * (dmax - dmin) * (pos < T(0) ? T(-1) : T(1)) / width;
  if (x > T(1)) { imp = dmax; dimp = T(0); }
}

// ------------------------------------------------------------------------------------------- dense Cholesky on the workspace (W.L)
template <class T, class WS>
MJA_HD void chol_factor(WS& W, Lanes X) {  // W.L holds the full symmetric matrix; afterwards its lower triangle is the factor
  for (int k = 0; k < NV; k++) {
    if (X.lane == 0) W.L[k][k] = sqrt(W.L[k][k]);
    X.sync();
    MJA_FOR(i, NV) if (i > k) W.L[i][k] /= W.L[k][k];
    X.sync();
    MJA_FOR(i, NV) if (i > k) for (int j = k + 1; j <= i; j++) W.L[i][j] -= W.L[i][k] * W.L[j][k];
    X.sync();
  }
}
template <class T, class WS>
MJA_HD void chol_solve(WS& W, Lanes X, T* x /* in: rhs, out: solution (workspace vector) */) {
  if (X.lane == 0) {
    for (int i = 0; i < NV; i++) {
      T s = x[i];
      for (int k = 0; k < i; k++) s -= W.L[i][k] * x[k];
      x[i] = s / W.L[i][i];
    }
    for (int i = NV - 1; i >= 0; i--) {
      T s = x[i];
      for (int k = i + 1; k < NV; k++) s -= W.L[k][i] * x[k];
      x[i] = s / W.L[i][i];
    }
  }
  X.sync();
}

#if defined(__CUDA_ARCH__)
// GPU: row `lane` of the SPD matrix lives in registers (a[0..NV), lower part used); column k is broadcast through W.col. On return
// a[k] (k < lane) = L[lane][k], dinv = 1 / L[lane][lane], and the rows are stored to W.L for the backward substitutions.
template <class WS>
__device__ __forceinline__ void gpu_chol_rows(WS& W, int lane, float (&a)[NVP], float& dinv) {
  dinv = 1.0f;
#pragma unroll
  for (int k = 0; k < NV; k++) {
    const float akk = __shfl_sync(0xffffffffu, a[k], k);
    const float inv = rsqrtf(akk);
    const float lik = a[k] * inv;          // valid for lane > k; lane k gets sqrt(akk)
    if (lane == k) dinv = inv;
    a[k] = lik;
    if (k + 1 < NV) {
      float* col = W.col[k & 1];
      col[lane] = lik;
      __syncwarp();
#pragma unroll
      for (int j = k + 1; j < NV; j++) a[j] -= lik * col[j];
    }
  }
  __syncwarp();
  if (lane < NV) {
#pragma unroll
    for (int k = 0; k < NV; k++) W.L[lane][k] = (k < lane) ? a[k] : (k == lane ? 1.0f / dinv : 0.0f);
  }
  __syncwarp();
}
// solve (L L^T) x = b from the rows in W.L and the lane's 1 / L_ii
template <class WS>
__device__ __forceinline__ float gpu_chol_solve(const WS& W, int lane, float dinv, float b) {
  const int row = lane < NV ? lane : 0;
  float y = lane < NV ? b : 0.0f;
#pragma unroll
  for (int k = 0; k < NV; k++) {
    const float yk = __shfl_sync(0xffffffffu, y * dinv, k);
    if (lane > k && lane < NV) y -= W.L[row][k] * yk;
  }
  float x = y * dinv;
#pragma unroll
  for (int k = NV - 1; k >= 0; k--) {
    const float xk = __shfl_sync(0xffffffffu, x * dinv, k);
    if (lane < k) x -= W.L[k][row] * xk;
  }
  return x * dinv;
}
#endif

// ------------------------------------------------------------------------------------------- gradient of sign * lam^T ID(q, vv, aa)
// ID = inverse dynamics M(q) aa + c(q, vv) (gravity optional). Accumulates into W.Sbar (cotangents of the joint motion axes S_e),
// W.Hacc (inertia-variation pseudo-forces summed over the bodies each dof moves) and W.gv (gradient with respect to vv).
template <class T, class WS>
MJA_HDN void idgrad(const DevModel& C, WS& W, Lanes X, const T* lam, const T* vv, const T* aa, bool grav, T sign) {
  fwd_vel_acc<T>(C, W, X, vv, aa, grav);
  twist_prefix<T>(C, W, X, lam, W.td0);                 // td0[d] = W prefix (twist of the body after dof d under velocity lam)
  // per body: f, Abar = I W, Vbar, H  -> tb0..tb3
  MJA_FOR(b, C.nbody) {
    T f[6] = {T(0), T(0), T(0), T(0), T(0), T(0)}, ab[6] = {T(0), T(0), T(0), T(0), T(0), T(0)}, vb[6] = {T(0), T(0), T(0), T(0), T(0), T(0)},
      hb[6] = {T(0), T(0), T(0), T(0), T(0), T(0)};
    const int ld = C.body_lastdof[b];
    if (b >= 1 && ld >= 0) {
      const T* I = W.cinert[b];
      const T* Wt = W.td0[ld]; const T* Vb = W.V[ld]; const T* Ab = W.A[ld];
      T IW[6], IV[6], IA[6], t[6], vxw[6], ivxw[6];
      inert_mul(IW, I, Wt); inert_mul(IV, I, Vb); inert_mul(IA, I, Ab);
      mcrossf(t, Vb, IV);
      for (int k = 0; k < 6; k++) { f[k] = IA[k] + t[k]; ab[k] = IW[k]; }
      mcross(vxw, Vb, Wt);
      inert_mul(ivxw, I, vxw);
      mcrossf(t, Wt, IV);
      for (int k = 0; k < 6; k++) vb[k] = -t[k] - ivxw[k];
      // H = W x* IA + A x* IW - (VxW) x* IV - V x* I(VxW)
      mcrossf(t, Wt, IA); for (int k = 0; k < 6; k++) hb[k] = t[k];
      mcrossf(t, Ab, IW); for (int k = 0; k < 6; k++) hb[k] += t[k];
      mcrossf(t, vxw, IV); for (int k = 0; k < 6; k++) hb[k] -= t[k];
      mcrossf(t, Vb, ivxw); for (int k = 0; k < 6; k++) hb[k] -= t[k];
    }
    for (int k = 0; k < 6; k++) { W.tb0[b][k] = f[k]; W.tb1[b][k] = ab[k]; W.tb2[b][k] = vb[k]; W.tb3[b][k] = hb[k]; }
  }
  X.sync();
  // per dof: subtree sums over the bodies the dof moves
  MJA_FOR(e, NV) {
    T fs[6] = {T(0), T(0), T(0), T(0), T(0), T(0)}, as[6] = {T(0), T(0), T(0), T(0), T(0), T(0)}, hs[6] = {T(0), T(0), T(0), T(0), T(0), T(0)};
    for (int b = 1; b < C.nbody; b++)
      if ((C.body_dofmask[b] >> e) & 1u)
        for (int k = 0; k < 6; k++) { fs[k] += W.tb0[b][k]; as[k] += W.tb1[b][k]; hs[k] += W.tb3[b][k]; }
    T sb[6], sdbar[6], t[6];
    for (int k = 0; k < 6; k++) { sb[k] = lam[e] * fs[k] + aa[e] * as[k]; sdbar[k] = vv[e] * as[k]; }
    const int src = C.dof_cvel_src[e];
    for (int k = 0; k < 6; k++) W.td1[e][k] = T(0);
    if (src != -2) {
      if (src >= 0) {
        mcrossf(t, W.V[src], sdbar);                      // Sd = Vsrc x S : Sbar -= Vsrc x* Sdbar
        for (int k = 0; k < 6; k++) sb[k] -= t[k];
        mcrossf(W.td1[e], W.S[e], sdbar);                 // ... and Vbar[src] += S x* Sdbar
      }
    }
    for (int k = 0; k < 6; k++) { W.Sbar[e][k] += sign * sb[k]; W.Hacc[e][k] += sign * hs[k]; W.td2[e][k] = as[k]; }
  }
  X.sync();
  MJA_FOR(e, NV) {
    T vt[6] = {T(0), T(0), T(0), T(0), T(0), T(0)};
    for (int b = 1; b < C.nbody; b++)
      if ((C.body_dofmask[b] >> e) & 1u)
        for (int k = 0; k < 6; k++) vt[k] += W.tb2[b][k];
    for (int e2 = 0; e2 < NV; e2++) {
      const int src = C.dof_cvel_src[e2];
      if (src >= 0 && ((C.dof_ancmask[src] >> e) & 1u))
        for (int k = 0; k < 6; k++) vt[k] += W.td1[e2][k];
    }
    for (int k = 0; k < 6; k++) W.Sbar[e][k] += sign * vv[e] * vt[k];
    W.gv[e] += sign * (dot6(W.S[e], vt) + dot6(W.td2[e], W.Sd[e]));
  }
  X.sync();
}

// ------------------------------------------------------------------------------------------- the whole step, forward part
template <class T, class WS>
MJA_HDN void vjp_forward(const DevModel& C, const PairParam* pp, WS& W, Lanes X, const EnvIO<T>& io) {
  const mjxb_env_config& cfg = C.cfg;
  MJA_FOR(i, 32) {
    W.q[i] = i < C.nq ? io.qpos[i] : T(0);
    W.v[i] = i < NV ? io.qvel[i] : T(0);
    W.a[i] = i < NV ? io.tape_qacc[i] : T(0);
    W.u[i] = T(0); W.uclip[i] = T(0);
    if (i < MJXB_AUX_DIM) W.aux[i] = io.aux ? io.aux[i] : T(0);
  }
  X.sync();
  MJA_FOR(i, C.nu) {
    if (io.aux == nullptr) { W.u[i] = io.action[i]; W.uclip[i] = T(1); continue; }   // physics-only: `action` is ctrl (mjx.step)
    const bool flip = W.aux[0] > T(0.5);
    const T val = flip ? io.action[cfg.act_perm[i]] * T(cfg.act_sign[i]) : io.action[i];
    W.u[i] = clampT(val, T(-1), T(1));
    W.uclip[i] = (val > T(-1) && val < T(1)) ? T(1) : T(0);
  }
  X.sync();
  fwd_kinematics<T>(C, W, X);
  fwd_com_cdof<T>(C, W, X);
  fwd_mass_matrix<T>(C, W, X);
  twist_prefix<T>(C, W, X, W.v, W.Zv);
  twist_prefix<T>(C, W, X, W.a, W.Za);
  // geoms
  MJA_FOR(g, C.ngeom) {
    const int b = C.geom_body[g];
    T gp[3] = {T(C.geom_pos[g][0]), T(C.geom_pos[g][1]), T(C.geom_pos[g][2])}, ga[3] = {T(C.geom_axis[g][0]), T(C.geom_axis[g][1]), T(C.geom_axis[g][2])}, t[3];
    rotq(t, gp, W.xquat[b]);
    for (int k = 0; k < 3; k++) W.gpos[g][k] = W.xpos[b][k] + t[k];
    rotq(W.gaxis[g], ga, W.xquat[b]);
  }
  X.sync();
  // collision: candidate flags per pair, ordered compaction, contact data of the candidates
  MJA_FOR(p, C.npair) {
    const int kind = (C.pair_w0[p] >> 16) & 0xff;
    unsigned char fl = 0;
    T dist, pos[3], n[3], t1[3], t2[3];
    contact_fwd<T>(C, W, p, 0, dist, pos, n, t1, t2);
    if (dist < T(0)) fl |= 1;
    if (kind == PAIR_PLANE_CAPSULE) {
      contact_fwd<T>(C, W, p, 1, dist, pos, n, t1, t2);
      if (dist < T(0)) fl |= 2;
    }
    W.pflag[p] = fl;
  }
  X.sync();
#if defined(__CUDA_ARCH__)
  {  // ordered compaction by ballots: a strip of 32 pairs, endpoint 0 and endpoint 1 interleaved in pair order
    int ncc = 0, ovf = 0;
    const unsigned lt = (1u << X.lane) - 1u;
    for (int base = 0; base < C.npair; base += 32) {
      const int p = base + X.lane;
      const unsigned fl = p < C.npair ? W.pflag[p] : 0u;
      const unsigned m0 = __ballot_sync(0xffffffffu, fl & 1u), m1 = __ballot_sync(0xffffffffu, fl & 2u);
      const int before = ncc + __popc(m0 & lt) + __popc(m1 & lt);
      const int cd = p < C.npair ? (int)((C.pair_w0[p] >> 24) & 0x7f) << 20 : 0;
      if (fl & 1u) { if (before < WS::MAXCC) W.cc_pair[before] = p | cd; else ovf = 1; }
      if (fl & 2u) { const int i1 = before + ((fl & 1u) ? 1 : 0); if (i1 < WS::MAXCC) W.cc_pair[i1] = p | (1 << 16) | cd; else ovf = 1; }
      ncc += __popc(m0) + __popc(m1);
    }
    ovf = __any_sync(0xffffffffu, ovf) ? 1 : 0;
    if (X.lane == 0) { W.ncc = ncc > WS::MAXCC ? WS::MAXCC : ncc; W.overflow = ovf; }
  }
#else
  if (X.lane == 0) {
    int ncc = 0, ovf = 0;
    for (int p = 0; p < C.npair; p++)
      for (int e = 0; e < 2; e++)
        if (W.pflag[p] & (1 << e)) {
          if (ncc < WS::MAXCC) W.cc_pair[ncc] = p | (e << 16) | ((int)((C.pair_w0[p] >> 24) & 0x7f) << 20);
          else ovf = 1;
          ncc++;
        }
    W.ncc = ncc > WS::MAXCC ? WS::MAXCC : ncc;
    W.overflow = ovf;
  }
#endif
  X.sync();
  MJA_FOR(c, W.ncc) {
    const int pr = W.cc_pair[c], p = pr & 0xffff, e = (pr >> 16) & 0xf;
    contact_fwd<T>(C, W, p, e, W.cc_dist[c], W.cc_pos[c], W.cc_n[c], W.cc_t1[c], W.cc_t2[c]);
    W.cc_mu[c] = T(pp[p].mu);
  }
  X.sync();
  // row list (limits, tendon limits, contacts) in MJX's static order
  if (X.lane == 0) {
    int nrow = 0, ovf = W.overflow;
    for (int i = 0; i < C.nlimit; i++) {
      const T qq = W.q[C.lim_qadr[i]];
      const T dmin = qq - T(C.lim_range[i][0]), dmax = T(C.lim_range[i][1]) - qq;
      const T pos = dmin < dmax ? dmin : dmax;
      if (pos < T(0)) {
        if (nrow < WS::CAP) { W.rinfo[nrow] = ROW_LIMIT | (i << 2) | (dmin < dmax ? 0 : 1 << 10); W.rpos[nrow] = pos; nrow++; } else ovf = 1;
      }
    }
    for (int i = 0; i < C.ntlimit; i++) {
      T len = T(0);
      for (int w = 0; w < C.ten_nwrap[i]; w++) len += T(C.ten_coef[i][w]) * W.q[C.ten_qpos[i][w]];
      const T dmin = len - T(C.ten_range[i][0]), dmax = T(C.ten_range[i][1]) - len;
      const T pos = dmin < dmax ? dmin : dmax;
      if (pos < T(0)) {
        if (nrow < WS::CAP) { W.rinfo[nrow] = ROW_TENDON | (i << 2) | (dmin < dmax ? 0 : 1 << 10); W.rpos[nrow] = pos; nrow++; } else ovf = 1;
      }
    }
    for (int c = 0; c < W.ncc; c++) {
      const int nr = ((W.cc_pair[c] >> 20) > 1) ? 4 : 1;
      if (nrow + nr <= WS::CAP) {
        W.cc_row[c] = nrow;
        for (int s = 0; s < nr; s++) { W.rinfo[nrow] = (nr == 4 ? ROW_CON3 : ROW_CON1) | (c << 2) | (s << 11); W.rpos[nrow] = W.cc_dist[c]; nrow++; }
      } else { W.cc_row[c] = -1; ovf = 1; }
    }
    W.nrow = nrow;
    W.overflow = ovf;
  }
  X.sync();
  const T h = T(C.timestep);
  MJA_FOR(r, W.nrow) {
    const int info = W.rinfo[r], kind = info & 3, idx = (info >> 2) & 0xff;
    const T sg = (info & (1 << 10)) ? T(-1) : T(1);
    T* Jr = W.J[r];
    for (int d = 0; d < NVP; d++) Jr[d] = T(0);
    T invw;
    const float *solref, *solimp;
    if (kind == ROW_LIMIT) {
      Jr[C.lim_dof[idx]] = sg;
      invw = T(C.lim_invweight[idx]); solref = C.lim_solref[idx]; solimp = C.lim_solimp[idx];
    } else if (kind == ROW_TENDON) {
      for (int w = 0; w < C.ten_nwrap[idx]; w++) Jr[C.ten_dof[idx][w]] += sg * T(C.ten_coef[idx][w]);
      invw = T(C.ten_invweight[idx]); solref = C.ten_solref[idx]; solimp = C.ten_solimp[idx];
    } else {
      const int c = idx, pr = W.cc_pair[c], p = pr & 0xffff, sub = (info >> 11) & 3;
      const uint32_t w0 = C.pair_w0[p];
      const int b1 = C.geom_body[w0 & 0xff], b2 = C.geom_body[(w0 >> 8) & 0xff];
      T dir[3];
      const T mu = W.cc_mu[c];
      for (int k = 0; k < 3; k++) {
        dir[k] = W.cc_n[c][k];
        if (kind == ROW_CON3) dir[k] += (sub == 0 ? mu : sub == 1 ? -mu : T(0)) * W.cc_t1[c][k] + (sub == 2 ? mu : sub == 3 ? -mu : T(0)) * W.cc_t2[c][k];
      }
      T off[3] = {W.cc_pos[c][0] - W.com[0], W.cc_pos[c][1] - W.com[1], W.cc_pos[c][2] - W.com[2]};
      T phi[6];
      cross3(phi, off, dir);
      phi[3] = dir[0]; phi[4] = dir[1]; phi[5] = dir[2];
      for (int d = 0; d < NV; d++) {
        T sgn = T(0);
        if ((C.body_dofmask[b2] >> d) & 1u) sgn += T(1);
        if ((C.body_dofmask[b1] >> d) & 1u) sgn -= T(1);
        if (sgn != T(0)) Jr[d] = sgn * dot6(phi, W.S[d]);
      }
      invw = T(pp[p].invweight); solref = pp[p].solref; solimp = pp[p].solimp;
    }
    T kk, bb, imp, dimp;
    kbi_d<T>(h, solref, solimp, W.rpos[r], kk, bb, imp, dimp);
    const T rr = maxT(invw * (T(1) - imp) / imp, T(1e-15));
    T vel = T(0), ja = T(0);
    for (int d = 0; d < NV; d++) { vel += Jr[d] * W.v[d]; ja += Jr[d] * W.a[d]; }
    const T D = T(1) / rr, aref = -bb * vel - kk * imp * W.rpos[r];
    W.rD[r] = D; W.rb[r] = bb; W.rk[r] = kk; W.rimp[r] = imp; W.rdimp[r] = dimp; W.rinvw[r] = invw; W.raref[r] = aref;
    const T jar = ja - aref;
    W.rjar[r] = jar;
    W.ract[r] = jar < T(0) ? 1 : 0;
    W.rf[r] = jar < T(0) ? -D * jar : T(0);
  }
  X.sync();
  // integrator acceleration x = (M + h D)^-1 (M a)  [M a = qfrc_smooth + J^T f at the solution], post-step velocity
  const bool damp = C.damp_implicit != 0;
  MJA_FOR(i, NV) {
    T s = T(0);
    for (int j = 0; j < NV; j++) s += W.M[i][j] * W.a[j];
    W.x[i] = damp ? s : W.a[i];
  }
  if (damp) {
#if defined(__CUDA_ARCH__)
    {
      X.sync();
      const int lane = X.lane, row = lane < NV ? lane : 0;
      float a[NVP], dinv;
#pragma unroll
      for (int j = 0; j < NVP; j++) a[j] = (lane < NV && j < NV) ? (float)W.M[row][j] + (j == lane ? (float)h * C.dof_damping[row] : 0.0f) : (j == lane ? 1.0f : 0.0f);
      gpu_chol_rows(W, lane, a, dinv);
      W.dinvA[lane] = dinv;
      const float xs = gpu_chol_solve(W, lane, dinv, (float)W.x[row]);
      __syncwarp();
      if (lane < NV) W.x[lane] = xs;
    }
#else
    MJA_FOR(i, NV) for (int j = 0; j < NV; j++) W.L[i][j] = W.M[i][j] + (i == j ? h * T(C.dof_damping[i]) : T(0));
    X.sync();
    chol_factor<T>(W, X);
    chol_solve<T>(W, X, W.x);
#endif
  }
  X.sync();
  MJA_FOR(i, 32) W.vnew[i] = i < NV ? W.v[i] + h * W.x[i] : T(0);
  X.sync();
}

// ------------------------------------------------------------------------------------------- the whole step, reverse part
template <class T, class WS>
MJA_HDN int step_vjp_env(const DevModel& C, const PairParam* pp, WS& W, Lanes X, const EnvIO<T>& io) {
  vjp_forward<T>(C, pp, W, X, io);
  const mjxb_env_config& cfg = C.cfg;
  const T h = T(C.timestep);
  const bool env_layer = io.aux != nullptr;
  MJA_FOR(i, 32) { W.gqt[i] = T(0); W.gv[i] = T(0); W.ubar[i] = T(0); W.gqdirect[i] = T(0); W.xbar[i] = T(0); W.abar[i] = T(0); W.wv[i] = T(0); }
  MJA_FOR(i, NV * 6) { (&W.Sbar[0][0])[i] = T(0); (&W.Hacc[0][0])[i] = T(0); }
  MJA_FOR(i, NB * 6) (&W.Wb[0][0])[i] = T(0);
  X.sync();
  // ---- (1) env layer (src/envs.py:333-492): reward and aux'[1,2,3,7] -> pelvis / head frames, v', ctrl, aux[1,2,3,7]   (lane 0)
  // wv accumulates the cotangent of the post-step velocity v'
  MJA_FOR(i, NV) W.wv[i] = io.g_qvel_out ? io.g_qvel_out[i] : T(0);
  X.sync();
  if (X.lane == 0) {
    for (int k = 0; k < MJXB_AUX_DIM; k++) W.gauxin[k] = T(0);
    for (int k = 0; k < 3; k++) { W.gbody_pelvis[k] = T(0); W.gbody_head[k] = T(0); }
    for (int k = 0; k < 4; k++) W.gquat_pelvis[k] = T(0);
    if (env_layer) {
      const int pb = cfg.pelvis_body_id, hb = cfg.head_body_id;
      const T bx = W.xpos[pb][0], by = W.xpos[pb][1], bz = W.xpos[pb][2], hx = W.xpos[hb][0], hy = W.xpos[hb][1];
      const T gr = io.g_reward;
      T tx = W.aux[1], ty = W.aux[2];
      // distances to the target before a possible advance
      const T dxp = tx - bx, dyp = ty - by, dxh = tx - hx, dyh = ty - hy;
      const T dp = sqrt(dxp * dxp + dyp * dyp), dh = sqrt(dxh * dxh + dyh * dyh);
      const T dist = dp > dh ? dp : dh;
      // progress = (-dist / h - last_pot) * w
      const T gdist = -gr * T(cfg.progress_weight) / h;
      W.gauxin[7] += -gr * T(cfg.progress_weight);
      if (dist > T(0)) {
        if (dp > dh) { const T gx = gdist * dxp / dp, gy = gdist * dyp / dp; W.gauxin[1] += gx; W.gauxin[2] += gy; W.gbody_pelvis[0] -= gx; W.gbody_pelvis[1] -= gy; }
        else { const T gx = gdist * dxh / dh, gy = gdist * dyh / dh; W.gauxin[1] += gx; W.gauxin[2] += gy; W.gbody_head[0] -= gx; W.gbody_head[1] -= gy; }
      }
      // energy = c1 mean|tau v'| + c2 mean tau^2  (reward -= energy)
      const int nj = NV - 6;
      for (int d = 6; d < NV; d++) {
        const int uu = C.dof_act[d];
        if (uu < 0) continue;
        const T lo = T(C.dof_ctrl_lo[d]), hi = T(C.dof_ctrl_hi[d]);
        const T cu = clampT(W.u[uu], lo, hi);
        const T tau = T(C.dof_gear[d]) * cu;
        const T pw = tau * W.vnew[d];
        const T sgn = pw > T(0) ? T(1) : (pw < T(0) ? T(-1) : T(0));
        const T gtau = -gr * (T(cfg.electricity_cost) * sgn * W.vnew[d] / T(nj) + T(2) * T(cfg.stall_torque_cost) * tau / T(nj));
        W.wv[d] += -gr * T(cfg.electricity_cost) * sgn * tau / T(nj);
        if (W.u[uu] > lo && W.u[uu] < hi) W.ubar[uu] += T(C.dof_gear[d]) * gtau;
      }
      // posture = w (|pitch| [outside window] + |roll| [outside window])  (reward -= posture)
      const T qw = W.xquat[pb][0], qx = W.xquat[pb][1], qy = W.xquat[pb][2], qz = W.xquat[pb][3];
      if (cfg.posture_penalty_weight != 0.0f) {
        const T sr = T(2) * (qw * qx + qy * qz), cr = T(1) - T(2) * (qx * qx + qy * qy);
        const T roll = atan2(sr, cr);
        const T sp0 = T(2) * (qw * qy - qz * qx), sp = clampT(sp0, T(-1), T(1));
        const T pitch = asin(sp);
        const bool p_ok = (pitch > T(-0.087)) && (pitch < T(0.174)), r_ok = (roll > T(-0.174)) && (roll < T(0.174));
        const T gpitch = p_ok ? T(0) : -gr * T(cfg.posture_penalty_weight) * (pitch < T(0) ? T(-1) : T(1));
        const T groll = r_ok ? T(0) : -gr * T(cfg.posture_penalty_weight) * (roll < T(0) ? T(-1) : T(1));
        if (groll != T(0)) {
          const T den = sr * sr + cr * cr;
          const T gs = groll * cr / den, gc = -groll * sr / den;
          W.gquat_pelvis[0] += gs * T(2) * qx; W.gquat_pelvis[1] += gs * T(2) * qw + gc * (-T(4) * qx);
          W.gquat_pelvis[2] += gs * T(2) * qz + gc * (-T(4) * qy); W.gquat_pelvis[3] += gs * T(2) * qy;
        }
        if (gpitch != T(0) && sp0 > T(-1) && sp0 < T(1)) {
          const T gs = gpitch / sqrt(T(1) - sp * sp);
          W.gquat_pelvis[0] += gs * T(2) * qy; W.gquat_pelvis[1] += gs * (-T(2) * qz);
          W.gquat_pelvis[2] += gs * T(2) * qw; W.gquat_pelvis[3] += gs * (-T(2) * qx);
        }
      }
      // aux' = [.., t', .., last_pot' = -dist_obs / h, ..]: the target advances when close for stop_frames steps
      const bool is_close = dist < T(cfg.target_threshold);
      const T close_count = is_close ? W.aux[4] + T(1) : T(0);
      const bool advance = close_count >= T(cfg.stop_frames);
      T gt[3] = {io.g_aux_out ? io.g_aux_out[1] : T(0), io.g_aux_out ? io.g_aux_out[2] : T(0), io.g_aux_out ? io.g_aux_out[3] : T(0)};
      const T glp = io.g_aux_out ? io.g_aux_out[7] : T(0);
      if (advance) { tx = bx + T(cfg.target_dist); ty = by; }
      const T ex = tx - bx, ey = ty - by, fx = tx - hx, fy = ty - hy;
      const T d2p = sqrt(ex * ex + ey * ey), d2h = sqrt(fx * fx + fy * fy);
      const T gdo = -glp / h;
      if (d2p > d2h) { if (d2p > T(0)) { const T gx = gdo * ex / d2p, gy = gdo * ey / d2p; gt[0] += gx; gt[1] += gy; W.gbody_pelvis[0] -= gx; W.gbody_pelvis[1] -= gy; } }
      else if (d2h > T(0)) { const T gx = gdo * fx / d2h, gy = gdo * fy / d2h; gt[0] += gx; gt[1] += gy; W.gbody_head[0] -= gx; W.gbody_head[1] -= gy; }
      if (advance) { for (int k = 0; k < 3; k++) W.gbody_pelvis[k] += gt[k]; }
      else { for (int k = 0; k < 3; k++) W.gauxin[1 + k] += gt[k]; }
      // pose wrenches of the two bodies the env layer reads
      {
        T off[3] = {W.xpos[pb][0] - W.com[0], W.xpos[pb][1] - W.com[1], W.xpos[pb][2] - W.com[2]}, c[3];
        cross3(c, off, W.gbody_pelvis);
        for (int k = 0; k < 3; k++) { W.Wb[pb][k] += c[k]; W.Wb[pb][3 + k] += W.gbody_pelvis[k]; }
        // quaternion cotangent -> world torque: d xquat = 1/2 (0, dtheta) (x) xquat
        for (int k = 0; k < 3; k++) {
          T ek[4] = {T(0), k == 0 ? T(1) : T(0), k == 1 ? T(1) : T(0), k == 2 ? T(1) : T(0)}, pq[4];
          quat_mul(pq, ek, W.xquat[pb]);
          W.Wb[pb][k] += T(0.5) * (pq[0] * W.gquat_pelvis[0] + pq[1] * W.gquat_pelvis[1] + pq[2] * W.gquat_pelvis[2] + pq[3] * W.gquat_pelvis[3]);
        }
        T offh[3] = {W.xpos[hb][0] - W.com[0], W.xpos[hb][1] - W.com[1], W.xpos[hb][2] - W.com[2]};
        cross3(c, offh, W.gbody_head);
        for (int k = 0; k < 3; k++) { W.Wb[hb][k] += c[k]; W.Wb[hb][3 + k] += W.gbody_head[k]; }
      }
    }
    // ---- (2) integrator (mjx forward._advance): q' = q (+) h v', v' = v + h x
    for (int i = 0; i < C.nq; i++) {
      const int kind = C.qpos_kind[i], ax = C.qpos_aux[i];
      const T g = io.g_qpos_out ? io.g_qpos_out[i] : T(0);
      if (kind == QK_HINGE || kind == QK_FREEPOS) { W.gqdirect[i] += g; W.wv[ax] += h * g; }
    }
    for (int j = 0; j < C.njnt; j++) {
      if (C.jnt_type[j] != 0 || io.g_qpos_out == nullptr) continue;
      const int qa = C.jnt_qposadr[j] + 3, da = C.jnt_dofadr[j] + 3;
      const T* qq = &W.q[qa];                       // normalised quaternion
      T w[3] = {W.vnew[da], W.vnew[da + 1], W.vnew[da + 2]};
      const T nrm = sqrt(dot3(w, w));
      const T ang = h * nrm * T(0.5), sn = sin(ang), cs = cos(ang);
      T what[3] = {T(0), T(0), T(0)};
      if (nrm > T(0)) for (int k = 0; k < 3; k++) what[k] = w[k] / nrm;
      T r[4] = {cs, what[0] * sn, what[1] * sn, what[2] * sn}, qn[4];
      quat_mul(qn, qq, r);
      const T n2 = sqrt(qn[0] * qn[0] + qn[1] * qn[1] + qn[2] * qn[2] + qn[3] * qn[3]);
      T qo[4] = {qn[0] / n2, qn[1] / n2, qn[2] / n2, qn[3] / n2};
      const T* g = io.g_qpos_out + qa;
      const T gq = qo[0] * g[0] + qo[1] * g[1] + qo[2] * g[2] + qo[3] * g[3];
      T gqn[4] = {(g[0] - qo[0] * gq) / n2, (g[1] - qo[1] * gq) / n2, (g[2] - qo[2] * gq) / n2, (g[3] - qo[3] * gq) / n2};
      T rc[4], qc[4], gqq[4], gr4[4];
      quat_conj(rc, r); quat_conj(qc, qq);
      quat_mul(gqq, gqn, rc);                       // qn = q (x) r : qbar = qnbar (x) r*, rbar = q* (x) qnbar
      quat_mul(gr4, qc, gqn);
      for (int k = 0; k < 4; k++) W.gqdirect[qa + k] += gqq[k];   // cotangent of the NORMALISED quaternion
      // r = (cos(h n / 2), sin(h n / 2) w / n)
      T gw[3];
      if (nrm > T(1e-10)) {
        const T rv = dot3(what, gr4 + 1);
        for (int k = 0; k < 3; k++)
          gw[k] = -sn * (h * T(0.5)) * what[k] * gr4[0] + (sn / nrm) * (gr4[1 + k] - what[k] * rv) + cs * (h * T(0.5)) * what[k] * rv;
      } else {
        for (int k = 0; k < 3; k++) gw[k] = (h * T(0.5)) * gr4[1 + k];
      }
      for (int k = 0; k < 3; k++) W.wv[da + k] += gw[k];
    }
  }
  X.sync();
  MJA_FOR(i, NV) { W.gv[i] += W.wv[i]; W.xbar[i] = h * W.wv[i]; }
  X.sync();
  // ---- (3) x = (M + h D)^-1 (M a): y = (M + h D)^-1 xbar ; Mbar += y (a - x)^T ; abar = M y     (W.L still holds that factor)
  const bool damp = C.damp_implicit != 0;
  MJA_FOR(i, 32) W.y[i] = i < NV ? W.xbar[i] : T(0);
  X.sync();
  if (damp) {
#if defined(__CUDA_ARCH__)
    {
      const float ys = gpu_chol_solve(W, X.lane, (float)W.dinvA[X.lane], (float)W.y[X.lane < NV ? X.lane : 0]);
      __syncwarp();
      if (X.lane < NV) W.y[X.lane] = ys;
      __syncwarp();
    }
#else
    chol_solve<T>(W, X, W.y);
#endif
    MJA_FOR(i, NV) {
      T s = T(0);
      for (int j = 0; j < NV; j++) s += W.M[i][j] * W.y[j];
      W.abar[i] = s;
    }
  } else {
    MJA_FOR(i, NV) { W.abar[i] = W.xbar[i]; W.y[i] = T(0); }
  }
  X.sync();
  // ---- (4) implicit-function adjoint of the solve: H lam = abar, H = M + sum_active D_r J_r J_r^T
#if defined(__CUDA_ARCH__)
  {
    X.sync();
    const int lane = X.lane, row = lane < NV ? lane : 0;
    float a[NVP], dinv;
#pragma unroll
    for (int j = 0; j < NVP; j++) a[j] = (lane < NV && j < NV) ? (float)W.M[row][j] : (j == lane ? 1.0f : 0.0f);
    for (int r = 0; r < W.nrow; r++) {
      if (!W.ract[r]) continue;                         // warp-uniform
      const float w = lane < NV ? (float)(W.rD[r] * W.J[r][row]) : 0.0f;
#pragma unroll
      for (int j = 0; j < NV; j++) a[j] += w * (float)W.J[r][j];
    }
    gpu_chol_rows(W, lane, a, dinv);
    const float ls = gpu_chol_solve(W, lane, dinv, (float)W.abar[row]);
    __syncwarp();
    W.lam[lane] = lane < NV ? ls : 0.0f;
    __syncwarp();
  }
#else
  MJA_FOR(i, NV) {
    for (int j = 0; j < NV; j++) {
      T s = W.M[i][j];
      for (int r = 0; r < W.nrow; r++)
        if (W.ract[r]) s += W.rD[r] * W.J[r][i] * W.J[r][j];
      W.L[i][j] = s;
    }
  }
  MJA_FOR(i, 32) W.lam[i] = i < NV ? W.abar[i] : T(0);
  X.sync();
  chol_factor<T>(W, X);
  chol_solve<T>(W, X, W.lam);
#endif
  twist_prefix<T>(C, W, X, W.lam, W.Zl);
  // rows: w = J lam ; cotangents of aref / D / pos ; direct terms in v ; relative-twist form of the contact-row cotangent phibar
  MJA_FOR(r, W.nrow) {
    T w = T(0);
    for (int d = 0; d < NV; d++) w += W.J[r][d] * W.lam[d];
    W.rw[r] = w;
    T posbar = T(0);
    for (int k = 0; k < 6; k++) W.rphibar[r][k] = T(0);
    if (W.ract[r]) {
      const T D = W.rD[r], arefbar = D * w, Dbar = -w * W.rjar[r];
      const T imp = W.rimp[r];
      posbar += -W.rk[r] * arefbar * (imp + W.rpos[r] * W.rdimp[r]);
      const T rr = W.rinvw[r] * (T(1) - imp) / imp;
      if (rr > T(1e-15)) posbar += Dbar * W.rdimp[r] / (W.rinvw[r] * (T(1) - imp) * (T(1) - imp));
      const int info = W.rinfo[r], kind = info & 3;
      if (kind == ROW_CON1 || kind == ROW_CON3) {
        const int c = (info >> 2) & 0xff, p = W.cc_pair[c] & 0xffff;
        const uint32_t w0 = C.pair_w0[p];
        const int b1 = C.geom_body[w0 & 0xff], b2 = C.geom_body[(w0 >> 8) & 0xff];
        const int l1 = C.body_lastdof[b1], l2 = C.body_lastdof[b2];
        const T alpha = W.rf[r], beta = -D * w, bb = W.rb[r];
        for (int k = 0; k < 6; k++) {
          const T dl = (l2 >= 0 ? W.Zl[l2][k] : T(0)) - (l1 >= 0 ? W.Zl[l1][k] : T(0));
          const T da = (l2 >= 0 ? W.Za[l2][k] : T(0)) - (l1 >= 0 ? W.Za[l1][k] : T(0));
          const T dv = (l2 >= 0 ? W.Zv[l2][k] : T(0)) - (l1 >= 0 ? W.Zv[l1][k] : T(0));
          W.rphibar[r][k] = alpha * dl + beta * (da + bb * dv);
        }
      }
    }
    W.rposbar[r] = posbar;
  }
  X.sync();
  // vbar from aref = -b J v - ... :  gv -= J^T (b arefbar) ;  and Sbar from the contact rows: Sbar_d += sum_r sgn_rd Jbar_rd phi_r
  MJA_FOR(d, NV) {
    T s = T(0);
    T sb[6] = {T(0), T(0), T(0), T(0), T(0), T(0)};
    for (int r = 0; r < W.nrow; r++) {
      if (!W.ract[r]) continue;
      const T D = W.rD[r], w = W.rw[r];
      s -= W.rb[r] * D * w * W.J[r][d];
      const int info = W.rinfo[r], kind = info & 3;
      if (kind != ROW_CON1 && kind != ROW_CON3) continue;
      const int c = (info >> 2) & 0xff, p = W.cc_pair[c] & 0xffff, sub = (info >> 11) & 3;
      const uint32_t w0 = C.pair_w0[p];
      const int b1 = C.geom_body[w0 & 0xff], b2 = C.geom_body[(w0 >> 8) & 0xff];
      T sgn = T(0);
      if ((C.body_dofmask[b2] >> d) & 1u) sgn += T(1);
      if ((C.body_dofmask[b1] >> d) & 1u) sgn -= T(1);
      if (sgn == T(0)) continue;
      const T jbar = W.rf[r] * W.lam[d] - D * w * (W.a[d] + W.rb[r] * W.v[d]);
      const T mu = W.cc_mu[c];
      T dir[3], off[3] = {W.cc_pos[c][0] - W.com[0], W.cc_pos[c][1] - W.com[1], W.cc_pos[c][2] - W.com[2]}, phi[6];
      for (int k = 0; k < 3; k++) {
        dir[k] = W.cc_n[c][k];
        if (kind == ROW_CON3) dir[k] += (sub == 0 ? mu : sub == 1 ? -mu : T(0)) * W.cc_t1[c][k] + (sub == 2 ? mu : sub == 3 ? -mu : T(0)) * W.cc_t2[c][k];
      }
      cross3(phi, off, dir);
      phi[3] = dir[0]; phi[4] = dir[1]; phi[5] = dir[2];
      for (int k = 0; k < 6; k++) sb[k] += sgn * jbar * phi[k];
    }
    W.gv[d] += s;
    for (int k = 0; k < 6; k++) W.Sbar[d][k] += sb[k];
  }
  // limit / tendon rows: pos is linear in qpos
  if (X.lane == 0) {
    for (int r = 0; r < W.nrow; r++) {
      const int info = W.rinfo[r], kind = info & 3, idx = (info >> 2) & 0xff;
      const T sg = (info & (1 << 10)) ? T(-1) : T(1);
      if (kind == ROW_LIMIT) W.gqt[C.lim_dof[idx]] += sg * W.rposbar[r];
      else if (kind == ROW_TENDON)
        for (int w = 0; w < C.ten_nwrap[idx]; w++) W.gqt[C.ten_dof[idx][w]] += sg * T(C.ten_coef[idx][w]) * W.rposbar[r];
    }
  }
  X.sync();
  // contacts: cotangents of (dist, pos, n, t1, t2) -> geoms -> pose wrenches of the two bodies
  MJA_FOR(c, W.ncc) {
    for (int k = 0; k < 12; k++) W.cc_w[c][k] = T(0);
    const int rb = W.cc_row[c];
    if (rb < 0) continue;
    const int pr = W.cc_pair[c], p = pr & 0xffff, e = (pr >> 16) & 0xf, nr = ((pr >> 20) > 1) ? 4 : 1;
    const T mu = W.cc_mu[c];
    T distbar = T(0), posbar[3] = {T(0), T(0), T(0)}, nbar[3] = {T(0), T(0), T(0)}, t1bar[3] = {T(0), T(0), T(0)}, t2bar[3] = {T(0), T(0), T(0)};
    T off[3] = {W.cc_pos[c][0] - W.com[0], W.cc_pos[c][1] - W.com[1], W.cc_pos[c][2] - W.com[2]};
    for (int s = 0; s < nr; s++) {
      const int r = rb + s;
      distbar += W.rposbar[r];
      const T* pb = W.rphibar[r];
      T dir[3], c1[3], dirbar[3];
      for (int k = 0; k < 3; k++) {
        dir[k] = W.cc_n[c][k];
        if (nr == 4) dir[k] += (s == 0 ? mu : s == 1 ? -mu : T(0)) * W.cc_t1[c][k] + (s == 2 ? mu : s == 3 ? -mu : T(0)) * W.cc_t2[c][k];
      }
      cross3(c1, dir, pb);                 // phi_ang = off x dir : posbar += dir x phibar_ang ; dirbar = phibar_ang x off + phibar_lin
      for (int k = 0; k < 3; k++) posbar[k] += c1[k];
      cross3(dirbar, pb, off);
      for (int k = 0; k < 3; k++) {
        dirbar[k] += pb[3 + k];
        nbar[k] += dirbar[k];
        if (nr == 4) {
          t1bar[k] += (s == 0 ? mu : s == 1 ? -mu : T(0)) * dirbar[k];
          t2bar[k] += (s == 2 ? mu : s == 3 ? -mu : T(0)) * dirbar[k];
        }
      }
    }
    T gp1[3] = {T(0), T(0), T(0)}, ga1[3] = {T(0), T(0), T(0)}, gp2[3] = {T(0), T(0), T(0)}, ga2[3] = {T(0), T(0), T(0)};
    contact_adj<T>(C, W, p, e, distbar, posbar, nbar, t1bar, t2bar, gp1, ga1, gp2, ga2);
    const uint32_t w0 = C.pair_w0[p];
    const int g1 = w0 & 0xff, g2 = (w0 >> 8) & 0xff;
    T o1[3] = {W.gpos[g1][0] - W.com[0], W.gpos[g1][1] - W.com[1], W.gpos[g1][2] - W.com[2]};
    T o2[3] = {W.gpos[g2][0] - W.com[0], W.gpos[g2][1] - W.com[1], W.gpos[g2][2] - W.com[2]};
    T ca[3], cb[3];
    cross3(ca, o1, gp1); cross3(cb, W.gaxis[g1], ga1);
    for (int k = 0; k < 3; k++) { W.cc_w[c][k] = ca[k] + cb[k]; W.cc_w[c][3 + k] = gp1[k]; }
    cross3(ca, o2, gp2); cross3(cb, W.gaxis[g2], ga2);
    for (int k = 0; k < 3; k++) { W.cc_w[c][6 + k] = ca[k] + cb[k]; W.cc_w[c][9 + k] = gp2[k]; }
  }
  X.sync();
  MJA_FOR(b, C.nbody) {
    if (b == 0) continue;
    for (int c = 0; c < W.ncc; c++) {
      const int p = W.cc_pair[c] & 0xffff;
      const uint32_t w0 = C.pair_w0[p];
      const int b1 = C.geom_body[w0 & 0xff], b2 = C.geom_body[(w0 >> 8) & 0xff];
      if (b1 == b) for (int k = 0; k < 6; k++) W.Wb[b][k] += W.cc_w[c][k];
      if (b2 == b) for (int k = 0; k < 6; k++) W.Wb[b][k] += W.cc_w[c][6 + k];
    }
  }
  X.sync();
  // ---- (5) smooth dynamics: L_dyn = y^T M (a - x) - lam^T (M a + c(q, v)) + lam . (passive + actuation)
  idgrad<T>(C, W, X, W.lam, W.v, W.a, true, T(-1));
  if (damp) {
    MJA_FOR(i, 32) { W.red[i] = i < NV ? W.a[i] - W.x[i] : T(0); W.wv[i] = T(0); }
    X.sync();
    // second bilinear form: velocity 0 -> only the M part; gv contributions vanish (vv = 0)
    idgrad<T>(C, W, X, W.y, W.wv, W.red, false, T(1));
  }
  MJA_FOR(d, NV) {
    const int qa = C.dof_qadr[d];
    if (qa >= 0) { W.gqt[d] += -T(C.dof_stiffness[d]) * W.lam[d]; }
    W.gv[d] += -T(C.dof_damping[d]) * W.lam[d];
    const int uu = C.dof_act[d];
    if (uu >= 0 && W.u[uu] > T(C.dof_ctrl_lo[d]) && W.u[uu] < T(C.dof_ctrl_hi[d])) W.ubar[uu] += T(C.dof_gear[d]) * W.lam[d];
  }
  X.sync();
  // ---- (6) tangent-space gradient: S_d . (sum of g_e = S_e x* Sbar_e over the dofs d acts on  -  inertia terms  +  pose wrenches)
  MJA_FOR(e, NV) mcrossf(W.td0[e], W.S[e], W.Sbar[e]);
  X.sync();
  MJA_FOR(d, NV) {
    const int j = C.dof_jnt[d];
    int gsrc = d;
    if (C.jnt_type[j] == 0 && d - C.jnt_dofadr[j] >= 3) gsrc = C.jnt_dofadr[j] + 3;   // the three body-frame rotations act on each other
    T tot[6] = {T(0), T(0), T(0), T(0), T(0), T(0)};
    for (int e = 0; e < NV; e++)
      if ((C.dof_ancmask[e] >> gsrc) & 1u)
        for (int k = 0; k < 6; k++) tot[k] += W.td0[e][k];
    for (int b = 1; b < C.nbody; b++)
      if ((C.body_dofmask[b] >> d) & 1u)
        for (int k = 0; k < 6; k++) tot[k] += W.Wb[b][k];
    for (int k = 0; k < 6; k++) tot[k] -= W.Hacc[d][k];
    W.gqt[d] += dot6(W.S[d], tot);
  }
  X.sync();
  // ---- (7) outputs: tangent -> qpos components, action unflip / clip
  if (X.lane == 0) {
    for (int i = 0; i < C.nq; i++) {
      const int kind = C.qpos_kind[i], ax = C.qpos_aux[i];
      if (kind == QK_HINGE || kind == QK_FREEPOS) io.g_qpos_in[i] = W.gqt[ax] + W.gqdirect[i];
    }
    for (int j = 0; j < C.njnt; j++) {
      if (C.jnt_type[j] != 0) continue;
      const int qa = C.jnt_qposadr[j] + 3, da = C.jnt_dofadr[j] + 3;
      const T* qq = &W.q[qa];
      T gw[4] = {T(0), W.gqt[da], W.gqt[da + 1], W.gqt[da + 2]}, t[4];
      quat_mul(t, qq, gw);                          // 2 q^ (x) (0, g_w)
      T gh[4];
      for (int k = 0; k < 4; k++) gh[k] = T(2) * t[k] + W.gqdirect[qa + k];
      const T gq = qq[0] * gh[0] + qq[1] * gh[1] + qq[2] * gh[2] + qq[3] * gh[3];
      const T nn = W.qnorm > T(0) ? W.qnorm : T(1);
      for (int k = 0; k < 4; k++) io.g_qpos_in[qa + k] = (gh[k] - qq[k] * gq) / nn;
    }
    for (int k = 0; k < MJXB_AUX_DIM; k++) if (io.g_aux_in) io.g_aux_in[k] = W.gauxin[k];
  }
  MJA_FOR(i, NV) io.g_qvel_in[i] = W.gv[i];
  MJA_FOR(i, C.nu) io.g_action[i] = T(0);
  X.sync();
  if (X.lane == 0) {   // ctrl_i = clip(flip ? action[perm_i] * sign_i : action_i)
    const bool flip = env_layer && W.aux[0] > T(0.5);
    for (int i = 0; i < C.nu; i++) {
      const T g = W.ubar[i] * W.uclip[i];
      if (flip) io.g_action[cfg.act_perm[i]] += g * T(cfg.act_sign[i]);
      else io.g_action[i] += g;
    }
  }
  X.sync();
  // A state that has already diverged (non-finite inputs, or a pull-back that overflowed) must not poison the rest of the batch through
  // the optimiser: its cotangents are zeroed and the caller is told (VJP_NONFINITE -> MJXB_STATUS_NAN).
  if (X.lane == 0) {
    bool ok = true;
    for (int i = 0; i < C.nq; i++) ok = ok && (absT(io.g_qpos_in[i]) < T(1e30));
    for (int i = 0; i < NV; i++) ok = ok && (absT(io.g_qvel_in[i]) < T(1e30));
    for (int i = 0; i < C.nu; i++) ok = ok && (absT(io.g_action[i]) < T(1e30));
    if (io.g_aux_in) for (int i = 0; i < MJXB_AUX_DIM; i++) ok = ok && (absT(io.g_aux_in[i]) < T(1e30));
    W.red[0] = ok ? T(1) : T(0);
    if (!ok) {
      for (int i = 0; i < C.nq; i++) io.g_qpos_in[i] = T(0);
      for (int i = 0; i < NV; i++) io.g_qvel_in[i] = T(0);
      for (int i = 0; i < C.nu; i++) io.g_action[i] = T(0);
      if (io.g_aux_in) for (int i = 0; i < MJXB_AUX_DIM; i++) io.g_aux_in[i] = T(0);
    }
  }
  X.sync();
  if (W.red[0] == T(0)) return VJP_NONFINITE;
  return W.overflow ? VJP_OVERFLOW : VJP_OK;
}

}  // namespace adj
}  // namespace mjxb
