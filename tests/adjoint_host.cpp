// tests/adjoint_host.cpp -- host-compiled build of the reverse-mode step (mujoco_mjx_lab_b200/csrc/mjxb_adjoint.cuh), TEST HARNESS ONLY.
// The same source the GPU kernel instantiates in float is instantiated here in double (and float), run sequentially, so that
// tests/test_adjoint_cpu.py can check the adjoint mathematics against finite differences of the float64 oracle without a GPU.
// Built by tests/helpers.build_adjoint_host() with g++ into tests/_build/libadjoint_host.so.
#include <cstdio>
#include <cstring>
#include <memory>

#include "../mujoco_mjx_lab_b200/csrc/mjxb_adjoint.cuh"

using namespace mjxb;
using namespace mjxb::adj;

namespace {
struct HostModel {
  DevModel D;
  PairParam pp[MJXB_MAXPAIR];
};
int make_model(const mjxb_model_blob* blob, const mjxb_env_config* cfg, HostModel& hm) {
  memset(hm.pp, 0, sizeof(hm.pp));
  return build_dev_model(*blob, cfg, hm.D, hm.pp);
}
template <class T> using WSBig = AdjS<T, CAP_BIG, MAXCC_BIG>;
}  // namespace

extern "C" {

// Full step VJP for n envs (double arithmetic when prec = 1, float when 0; arrays are double either way).
int adj_step_vjp(const mjxb_model_blob* blob, const mjxb_env_config* cfg, int prec, int n, const double* qpos, const double* qvel,
                 const double* aux, const double* action, const double* tape_qacc, const double* g_qpos_out, const double* g_qvel_out,
                 const double* g_aux_out, const double* g_reward, double* g_qpos_in, double* g_qvel_in, double* g_aux_in,
                 double* g_action, int* status) {
  std::unique_ptr<HostModel> hm(new HostModel());
  int rc = make_model(blob, cfg, *hm);
  if (rc) return rc;
  const DevModel& C = hm->D;
  const int nq = C.nq, nu = C.nu;
  Lanes X{0, 1};
  if (prec == 1) {
    std::unique_ptr<WSBig<double>> W(new WSBig<double>());
    for (int e = 0; e < n; e++) {
      EnvIO<double> io;
      io.qpos = qpos + (size_t)e * nq; io.qvel = qvel + (size_t)e * NV; io.aux = aux ? aux + (size_t)e * MJXB_AUX_DIM : nullptr;
      io.action = action + (size_t)e * nu; io.tape_qacc = tape_qacc + (size_t)e * NV;
      io.g_qpos_out = g_qpos_out ? g_qpos_out + (size_t)e * nq : nullptr; io.g_qvel_out = g_qvel_out ? g_qvel_out + (size_t)e * NV : nullptr;
      io.g_aux_out = g_aux_out ? g_aux_out + (size_t)e * MJXB_AUX_DIM : nullptr; io.g_reward = g_reward ? g_reward[e] : 0.0;
      io.g_qpos_in = g_qpos_in + (size_t)e * nq; io.g_qvel_in = g_qvel_in + (size_t)e * NV;
      io.g_aux_in = g_aux_in ? g_aux_in + (size_t)e * MJXB_AUX_DIM : nullptr; io.g_action = g_action + (size_t)e * nu;
      const int st = step_vjp_env<double>(C, hm->pp, *W, X, io);
      if (status) status[e] = st;
    }
  } else {
    std::unique_ptr<WSBig<float>> W(new WSBig<float>());
    float bi[6][32], bo[4][32];
    for (int e = 0; e < n; e++) {
      EnvIO<float> io;
      auto ld = [&](float* dst, const double* src, int k) { if (src) for (int i = 0; i < k; i++) dst[i] = (float)src[i]; };
      float fq[32], fv[32], fa[16], fu[32], ft[32], gq[32], gv[32], ga[16];
      ld(fq, qpos + (size_t)e * nq, nq); ld(fv, qvel + (size_t)e * NV, NV); if (aux) ld(fa, aux + (size_t)e * MJXB_AUX_DIM, MJXB_AUX_DIM);
      ld(fu, action + (size_t)e * nu, nu); ld(ft, tape_qacc + (size_t)e * NV, NV);
      if (g_qpos_out) ld(gq, g_qpos_out + (size_t)e * nq, nq);
      if (g_qvel_out) ld(gv, g_qvel_out + (size_t)e * NV, NV);
      if (g_aux_out) ld(ga, g_aux_out + (size_t)e * MJXB_AUX_DIM, MJXB_AUX_DIM);
      io.qpos = fq; io.qvel = fv; io.aux = aux ? fa : nullptr; io.action = fu; io.tape_qacc = ft;
      io.g_qpos_out = g_qpos_out ? gq : nullptr; io.g_qvel_out = g_qvel_out ? gv : nullptr; io.g_aux_out = g_aux_out ? ga : nullptr;
      io.g_reward = g_reward ? (float)g_reward[e] : 0.0f;
      io.g_qpos_in = bo[0]; io.g_qvel_in = bo[1]; io.g_aux_in = bo[2]; io.g_action = bo[3];
      const int st = step_vjp_env<float>(C, hm->pp, *W, X, io);
      if (status) status[e] = st;
      for (int i = 0; i < nq; i++) g_qpos_in[(size_t)e * nq + i] = bo[0][i];
      for (int i = 0; i < NV; i++) g_qvel_in[(size_t)e * NV + i] = bo[1][i];
      if (g_aux_in) for (int i = 0; i < MJXB_AUX_DIM; i++) g_aux_in[(size_t)e * MJXB_AUX_DIM + i] = bo[2][i];
      for (int i = 0; i < nu; i++) g_action[(size_t)e * nu + i] = bo[3][i];
      (void)bi;
    }
  }
  return 0;
}

// Unit-test hook: tangent-space gradient (27 dofs) and velocity gradient of  lam^T ID(q, vv, aa)  (gravity on / off), one env.
int adj_idgrad(const mjxb_model_blob* blob, const double* qpos, const double* lam, const double* vv, const double* aa, int grav,
               double* gqt, double* gv) {
  std::unique_ptr<HostModel> hm(new HostModel());
  int rc = make_model(blob, nullptr, *hm);
  if (rc) return rc;
  const DevModel& C = hm->D;
  std::unique_ptr<WSBig<double>> Wp(new WSBig<double>());
  WSBig<double>& W = *Wp;
  Lanes X{0, 1};
  for (int i = 0; i < 32; i++) { W.q[i] = i < C.nq ? qpos[i] : 0.0; W.gv[i] = 0; W.gqt[i] = 0; }
  memset(W.Sbar, 0, sizeof(W.Sbar)); memset(W.Hacc, 0, sizeof(W.Hacc)); memset(W.Wb, 0, sizeof(W.Wb));
  fwd_kinematics<double>(C, W, X);
  fwd_com_cdof<double>(C, W, X);
  double l32[32] = {0}, v32[32] = {0}, a32[32] = {0};
  for (int i = 0; i < NV; i++) { l32[i] = lam[i]; v32[i] = vv[i]; a32[i] = aa[i]; }
  idgrad<double>(C, W, X, l32, v32, a32, grav != 0, 1.0);
  for (int e = 0; e < NV; e++) mcrossf(W.td0[e], W.S[e], W.Sbar[e]);
  for (int d = 0; d < NV; d++) {
    const int j = C.dof_jnt[d];
    int gsrc = d;
    if (C.jnt_type[j] == 0 && d - C.jnt_dofadr[j] >= 3) gsrc = C.jnt_dofadr[j] + 3;
    double tot[6] = {0, 0, 0, 0, 0, 0};
    for (int e = 0; e < NV; e++)
      if ((C.dof_ancmask[e] >> gsrc) & 1u) for (int k = 0; k < 6; k++) tot[k] += W.td0[e][k];
    for (int k = 0; k < 6; k++) tot[k] -= W.Hacc[d][k];
    gqt[d] = dot6(W.S[d], tot);
    gv[d] = W.gv[d];
  }
  return 0;
}

// Unit-test hook: contact geometry of pair p, contact e from explicit geom frames, and its adjoint.
int adj_contact(const mjxb_model_blob* blob, int p, int e, const double* gpos /*[ngeom,3]*/, const double* gaxis /*[ngeom,3]*/,
                double* out /*dist, pos3, n3, t1 3, t2 3 = 13*/, const double* outbar /*13*/, double* gbar /*gp1 3, ga1 3, gp2 3, ga2 3*/) {
  std::unique_ptr<HostModel> hm(new HostModel());
  int rc = make_model(blob, nullptr, *hm);
  if (rc) return rc;
  const DevModel& C = hm->D;
  std::unique_ptr<WSBig<double>> Wp(new WSBig<double>());
  WSBig<double>& W = *Wp;
  for (int g = 0; g < C.ngeom; g++) for (int k = 0; k < 3; k++) { W.gpos[g][k] = gpos[3 * g + k]; W.gaxis[g][k] = gaxis[3 * g + k]; }
  contact_fwd<double>(C, W, p, e, out[0], out + 1, out + 4, out + 7, out + 10);
  if (outbar && gbar) {
    for (int k = 0; k < 12; k++) gbar[k] = 0;
    contact_adj<double>(C, W, p, e, outbar[0], outbar + 1, outbar + 4, outbar + 7, outbar + 10, gbar, gbar + 3, gbar + 6, gbar + 9);
  }
  return 0;
}

}  // extern "C"
