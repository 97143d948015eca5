// oracle/oracle_capi.cpp -- extern "C" surface of the CPU oracle (TEST INFRASTRUCTURE, see oracle.hpp header).
// All floating arrays cross this boundary as double; `prec` selects the arithmetic (0 = float32, 1 = float64).
#include <omp.h>

#include <cstdio>
#include <memory>

#include "oracle.hpp"

using namespace orc;

extern "C" {

typedef struct orc_debug {
  double* xpos; double* xquat; double* qM; double* qfrc_bias; double* qfrc_passive; double* qfrc_actuator;
  double* qacc_smooth; double* con_dist; double* con_pos; double* con_normal; double* efc_J; double* efc_pos;
  double* efc_D; double* efc_aref; double* efc_force; int32_t* efc_active; double* qacc; double* qfrc_constraint;
  double* sensordata; int32_t* solver_niter; int64_t* flops;
  double* cdof; double* cinert; double* subtree_com; double* qfrc_smooth;
  int64_t* flops_act;
} orc_debug;

size_t orc_blob_sizeof(void) { return sizeof(mjxb_model_blob); }
size_t orc_env_config_sizeof(void) { return sizeof(mjxb_env_config); }

void orc_threefry2x32(uint32_t k0, uint32_t k1, uint32_t c0, uint32_t c1, uint32_t* out) {
  threefry2x32(k0, k1, c0, c1, &out[0], &out[1]);
}
void orc_jax_split(const uint32_t* key, int n, uint32_t* out) {
  for (int i = 0; i < n; i++) jax_split(key, i, out + 2 * i);
}
void orc_jax_uniform(const uint32_t* key, int n, float minval, float maxval, float* out) {
  for (int i = 0; i < n; i++) out[i] = jax_uniform(key, i, minval, maxval);
}

}  // extern "C"

namespace {

template <class R> void load_state(const mjxb_model_blob& m, Data<R>& d, int e, const double* qpos, const double* qvel,
                                   const double* warm, const double* time, const double* ctrl) {
  for (int i = 0; i < m.nq; i++) d.qpos[i] = R(qpos[(size_t)e * m.nq + i]);
  for (int i = 0; i < m.nv; i++) d.qvel[i] = R(qvel[(size_t)e * m.nv + i]);
  for (int i = 0; i < m.nv; i++) d.qacc_warmstart[i] = warm ? R(warm[(size_t)e * m.nv + i]) : R(0);
  for (int i = 0; i < m.nu; i++) d.ctrl[i] = ctrl ? R(ctrl[(size_t)e * m.nu + i]) : R(0);
  d.time = time ? R(time[e]) : R(0);
}
template <class R> void store_state(const mjxb_model_blob& m, const Data<R>& d, int e, double* qpos, double* qvel, double* warm,
                                    double* time) {
  for (int i = 0; i < m.nq; i++) qpos[(size_t)e * m.nq + i] = double(d.qpos[i]);
  for (int i = 0; i < m.nv; i++) qvel[(size_t)e * m.nv + i] = double(d.qvel[i]);
  if (warm) for (int i = 0; i < m.nv; i++) warm[(size_t)e * m.nv + i] = double(d.qacc_warmstart[i]);
  if (time) time[e] = double(d.time);
}
template <class R> void store_debug(const mjxb_model_blob& m, const Data<R>& d, int e, const orc_debug* g) {
  if (!g) return;
  size_t nb = m.nbody, nv = m.nv, nc = m.ncon, ne = m.nefc;
  if (g->xpos) for (size_t b = 0; b < nb; b++) for (int k = 0; k < 3; k++) g->xpos[(e * nb + b) * 3 + k] = d.xpos[b][k];
  if (g->xquat) for (size_t b = 0; b < nb; b++) for (int k = 0; k < 4; k++) g->xquat[(e * nb + b) * 4 + k] = d.xquat[b][k];
  if (g->qM) for (size_t i = 0; i < nv; i++) for (size_t j = 0; j < nv; j++) g->qM[(e * nv + i) * nv + j] = d.qM[i][j];
  for (size_t i = 0; i < nv; i++) {
    if (g->qfrc_bias) g->qfrc_bias[e * nv + i] = d.qfrc_bias[i];
    if (g->qfrc_passive) g->qfrc_passive[e * nv + i] = d.qfrc_passive[i];
    if (g->qfrc_actuator) g->qfrc_actuator[e * nv + i] = d.qfrc_actuator[i];
    if (g->qfrc_smooth) g->qfrc_smooth[e * nv + i] = d.qfrc_smooth[i];
    if (g->qacc_smooth) g->qacc_smooth[e * nv + i] = d.qacc_smooth[i];
    if (g->qacc) g->qacc[e * nv + i] = d.qacc[i];
    if (g->qfrc_constraint) g->qfrc_constraint[e * nv + i] = d.qfrc_constraint[i];
    if (g->cdof) for (int k = 0; k < 6; k++) g->cdof[(e * nv + i) * 6 + k] = d.cdof[i][k];
  }
  if (g->cinert) for (size_t b = 0; b < nb; b++) for (int k = 0; k < 10; k++) g->cinert[(e * nb + b) * 10 + k] = d.cinert[b][k];
  if (g->subtree_com) for (size_t b = 0; b < nb; b++) for (int k = 0; k < 3; k++) g->subtree_com[(e * nb + b) * 3 + k] = d.subtree_com[b][k];
  for (size_t c = 0; c < nc; c++) {
    if (g->con_dist) g->con_dist[e * nc + c] = d.con_dist[c];
    if (g->con_pos) for (int k = 0; k < 3; k++) g->con_pos[(e * nc + c) * 3 + k] = d.con_pos[c][k];
    if (g->con_normal) for (int k = 0; k < 3; k++) g->con_normal[(e * nc + c) * 3 + k] = d.con_frame[c][k];
  }
  for (size_t r = 0; r < ne; r++) {
    if (g->efc_J) for (size_t j = 0; j < nv; j++) g->efc_J[(e * ne + r) * nv + j] = d.efc_J[r][j];
    if (g->efc_pos) g->efc_pos[e * ne + r] = d.efc_pos[r];
    if (g->efc_D) g->efc_D[e * ne + r] = d.efc_D[r];
    if (g->efc_aref) g->efc_aref[e * ne + r] = d.efc_aref[r];
    if (g->efc_force) g->efc_force[e * ne + r] = d.efc_force[r];
    if (g->efc_active) g->efc_active[e * ne + r] = (d.efc_cand[r] ? 1 : 0) | (d.efc_active[r] ? 2 : 0);
  }
  if (g->sensordata) for (int s = 0; s < m.nsensor; s++) g->sensordata[(size_t)e * m.nsensor + s] = d.sensordata[s];
  if (g->solver_niter) g->solver_niter[e] = d.solver_niter;
  if (g->flops) g->flops[e] = d.flops;
  if (g->flops_act) g->flops_act[e] = d.flops_act;
}

template <class R>
void physics_step_t(const mjxb_model_blob& m, int n, int nsteps, int do_integrate, double* qpos, double* qvel, double* warm,
                    double* time, const double* ctrl, const orc_debug* dbg, int nthreads) {
#pragma omp parallel num_threads(nthreads)
  {
    std::unique_ptr<Data<R>> dp(new Data<R>());
    Data<R>& d = *dp;
#pragma omp for schedule(dynamic, 4)
    for (int e = 0; e < n; e++) {
      load_state(m, d, e, qpos, qvel, warm, time, ctrl);
      for (int s = 0; s < nsteps; s++) {
        d.flops = 0; d.flops_act = 0;
        forward(m, d);
        if (s == nsteps - 1) store_debug(m, d, e, dbg);
        if (do_integrate) integrate(m, d);
      }
      store_state(m, d, e, qpos, qvel, warm, time);
    }
  }
}

template <class R>
void env_reset_t(const mjxb_model_blob& m, const mjxb_env_config& cfg, int n, const uint32_t* keys, double* qpos, double* qvel,
                 double* warm, double* time, double* aux, double* obs, int nthreads) {
#pragma omp parallel num_threads(nthreads)
  {
    std::unique_ptr<Data<R>> dp(new Data<R>());
    Data<R>& d = *dp;
#pragma omp for schedule(dynamic, 4)
    for (int e = 0; e < n; e++) {
      R a[MJXB_AUX_DIM], o[MJXB_MAXOBS];
      env_reset(m, cfg, keys + 2 * (size_t)e, d, a, o);
      store_state(m, d, e, qpos, qvel, warm, time);
      for (int k = 0; k < MJXB_AUX_DIM; k++) aux[(size_t)e * MJXB_AUX_DIM + k] = double(a[k]);
      for (int k = 0; k < cfg.obs_dim; k++) obs[(size_t)e * cfg.obs_dim + k] = double(o[k]);
    }
  }
}

template <class R>
void env_step_t(const mjxb_model_blob& m, const mjxb_env_config& cfg, int n, double* qpos, double* qvel, double* warm, double* time,
                double* aux, const double* action, double* obs, double* reward, double* terminated, double* truncated,
                const uint32_t* reset_keys, uint8_t* reset_mask, const orc_debug* dbg, int nthreads) {
#pragma omp parallel num_threads(nthreads)
  {
    std::unique_ptr<Data<R>> dp(new Data<R>());
    Data<R>& d = *dp;
#pragma omp for schedule(dynamic, 4)
    for (int e = 0; e < n; e++) {
      load_state(m, d, e, qpos, qvel, warm, time, (const double*)nullptr);
      R a[MJXB_AUX_DIM], act[MJXB_MAXU];
      for (int k = 0; k < MJXB_AUX_DIM; k++) a[k] = R(aux[(size_t)e * MJXB_AUX_DIM + k]);
      for (int k = 0; k < m.nu; k++) act[k] = R(action[(size_t)e * m.nu + k]);
      EnvOut<R> out;
      d.flops = 0; d.flops_act = 0;
      env_step(m, cfg, d, a, act, out);
      store_debug(m, d, e, dbg);
      reward[e] = double(out.reward); terminated[e] = double(out.terminated); truncated[e] = double(out.truncated);
      bool done = std::max(out.terminated, out.truncated) > R(0);
      if (reset_mask) reset_mask[e] = (reset_keys && done) ? 1 : 0;
      if (reset_keys && done) env_reset(m, cfg, reset_keys + 2 * (size_t)e, d, a, out.obs);  // train_ppo.py:150-161 merge
      store_state(m, d, e, qpos, qvel, warm, time);
      for (int k = 0; k < MJXB_AUX_DIM; k++) aux[(size_t)e * MJXB_AUX_DIM + k] = double(a[k]);
      for (int k = 0; k < cfg.obs_dim; k++) obs[(size_t)e * cfg.obs_dim + k] = double(out.obs[k]);
    }
  }
}

// mjx_humanoid_speed_test.py:48-57: make_data -> qvel[0] = vel -> mjx.step -> qpos[0]
template <class R> void speed_test_t(const mjxb_model_blob& m, int n, const double* vel, double* pos, int iters, int nthreads) {
#pragma omp parallel num_threads(nthreads)
  {
    std::unique_ptr<Data<R>> dp(new Data<R>());
    Data<R>& d = *dp;
    for (int it = 0; it < iters; it++) {
#pragma omp for schedule(dynamic, 4)
      for (int e = 0; e < n; e++) {
        for (int i = 0; i < m.nq; i++) d.qpos[i] = R(m.qpos0[i]);
        for (int i = 0; i < m.nv; i++) { d.qvel[i] = 0; d.qacc_warmstart[i] = 0; }
        for (int i = 0; i < m.nu; i++) d.ctrl[i] = 0;
        d.time = 0;
        d.qvel[0] = R(vel[e]);
        step(m, d);
        pos[e] = double(d.qpos[0]);
      }
    }
  }
}

}  // namespace

extern "C" {

int orc_max_threads(void) { return omp_get_max_threads(); }

// mj_setConst quantities at qpos0 (engine_setconst.c: dof_invweight0, body_invweight0, tendon_invweight0, stat.meaninertia), derived
// here from the oracle's OWN pipeline -- kinematics -> com_pos -> crb mass matrix -> Cholesky -> M^-1, body Jacobians from cdof -- i.e.
// independently of modelc.py's per-body point-Jacobian construction that produced the values stored in the blob. Float64 throughout.
void orc_set_const(const mjxb_model_blob* mp, double* dof_invweight0, double* body_invweight0 /*[nbody,2]*/, double* tendon_invweight0,
                   double* meaninertia) {
  const mjxb_model_blob& m = *mp;
  std::unique_ptr<Data<double>> dp(new Data<double>());
  Data<double>& d = *dp;
  for (int i = 0; i < m.nq; i++) d.qpos[i] = double(m.qpos0[i]);
  for (int i = 0; i < m.nv; i++) d.qvel[i] = 0;
  kinematics(m, d);
  com_pos(m, d);
  crb(m, d);
  chol_factor(d.qL, d.qM, m.nv);
  static double Minv[MJXB_MAXDOF][MJXB_MAXDOF];
  for (int j = 0; j < m.nv; j++) {
    double e[MJXB_MAXDOF] = {0}, x[MJXB_MAXDOF];
    e[j] = 1;
    chol_solve(x, d.qL, e, m.nv);
    for (int i = 0; i < m.nv; i++) Minv[i][j] = x[i];
  }
  double tr = 0;
  for (int i = 0; i < m.nv; i++) { dof_invweight0[i] = Minv[i][i]; tr += d.qM[i][i]; }
  *meaninertia = tr / m.nv;
  for (int j = 0; j < m.njnt; j++)
    if (m.jnt_type[j] == 0) {  // free joint: translational and rotational dofs share their mean
      int a = m.jnt_dofadr[j];
      double t = (Minv[a][a] + Minv[a + 1][a + 1] + Minv[a + 2][a + 2]) / 3, r = (Minv[a + 3][a + 3] + Minv[a + 4][a + 4] + Minv[a + 5][a + 5]) / 3;
      for (int k = 0; k < 3; k++) { dof_invweight0[a + k] = t; dof_invweight0[a + 3 + k] = r; }
    }
  for (int b = 0; b < m.nbody; b++) {
    body_invweight0[2 * b] = body_invweight0[2 * b + 1] = 0;
    if (b == 0) continue;
    double jp[MJXB_MAXDOF][3], jr[MJXB_MAXDOF][3];
    point_jac(m, d, jp, d.xipos[b], b);   // translational Jacobian of the body's inertial-frame origin (mjx support.jac)
    for (int j = 0; j < m.nv; j++) jr[j][0] = jr[j][1] = jr[j][2] = 0;
    int bb = b;
    while (bb > 0 && m.body_dofnum[bb] == 0) bb = m.body_parent[bb];
    if (bb > 0)
      for (int j = m.body_dofadr[bb] + m.body_dofnum[bb] - 1; j >= 0; j = m.dof_parent[j])
        for (int k = 0; k < 3; k++) jr[j][k] = d.cdof[j][k];   // angular part of cdof = world rotation axis of the dof
    double tp = 0, trr = 0;
    for (int k = 0; k < 3; k++)
      for (int i = 0; i < m.nv; i++)
        for (int j = 0; j < m.nv; j++) { tp += jp[i][k] * Minv[i][j] * jp[j][k]; trr += jr[i][k] * Minv[i][j] * jr[j][k]; }
    body_invweight0[2 * b] = std::max(MINVAL, tp / 3);
    body_invweight0[2 * b + 1] = std::max(MINVAL, trr / 3);
  }
  for (int t = 0; t < m.ntendon; t++) {
    double jt[MJXB_MAXDOF] = {0}, s = 0;
    for (int w = 0; w < m.ten_nwrap[t]; w++) jt[m.ten_dof[t][w]] += double(m.ten_coef[t][w]);
    for (int i = 0; i < m.nv; i++)
      for (int j = 0; j < m.nv; j++) s += jt[i] * Minv[i][j] * jt[j];
    tendon_invweight0[t] = s;
  }
}

void orc_physics_step(const mjxb_model_blob* m, int prec, int n, int nsteps, int do_integrate, double* qpos, double* qvel,
                      double* warm, double* time, const double* ctrl, const orc_debug* dbg, int nthreads) {
  if (nthreads <= 0) nthreads = omp_get_max_threads();
  if (prec == 0) physics_step_t<float>(*m, n, nsteps, do_integrate, qpos, qvel, warm, time, ctrl, dbg, nthreads);
  else physics_step_t<double>(*m, n, nsteps, do_integrate, qpos, qvel, warm, time, ctrl, dbg, nthreads);
}
void orc_env_reset(const mjxb_model_blob* m, const mjxb_env_config* cfg, int prec, int n, const uint32_t* keys, double* qpos,
                   double* qvel, double* warm, double* time, double* aux, double* obs, int nthreads) {
  if (nthreads <= 0) nthreads = omp_get_max_threads();
  if (prec == 0) env_reset_t<float>(*m, *cfg, n, keys, qpos, qvel, warm, time, aux, obs, nthreads);
  else env_reset_t<double>(*m, *cfg, n, keys, qpos, qvel, warm, time, aux, obs, nthreads);
}
void orc_env_step(const mjxb_model_blob* m, const mjxb_env_config* cfg, int prec, int n, double* qpos, double* qvel, double* warm,
                  double* time, double* aux, const double* action, double* obs, double* reward, double* terminated,
                  double* truncated, const uint32_t* reset_keys, uint8_t* reset_mask, const orc_debug* dbg, int nthreads) {
  if (nthreads <= 0) nthreads = omp_get_max_threads();
  if (prec == 0)
    env_step_t<float>(*m, *cfg, n, qpos, qvel, warm, time, aux, action, obs, reward, terminated, truncated, reset_keys, reset_mask, dbg, nthreads);
  else
    env_step_t<double>(*m, *cfg, n, qpos, qvel, warm, time, aux, action, obs, reward, terminated, truncated, reset_keys, reset_mask, dbg, nthreads);
}
void orc_speed_test(const mjxb_model_blob* m, int prec, int n, const double* vel, double* pos, int iters, int nthreads) {
  if (nthreads <= 0) nthreads = omp_get_max_threads();
  if (prec == 0) speed_test_t<float>(*m, n, vel, pos, iters, nthreads);
  else speed_test_t<double>(*m, n, vel, pos, iters, nthreads);
}

}  // extern "C"
