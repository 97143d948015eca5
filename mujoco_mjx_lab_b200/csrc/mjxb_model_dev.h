// mjxb_model_dev.h -- the device-side model record (DevModel) and its host-side construction from the model blob.
// Plain C++ (no CUDA): included by the kernels (mjxb_device.cuh, mjxb_adjoint.cuh), by the C-ABI host code (mjxb_abi.cu) and by the
// host-compiled build of the reverse-mode step that the CPU tests check against finite differences (tests/adjoint_host.cpp).
#pragma once
#include <stdint.h>
#include <string.h>

#include "mjxb.h"

namespace mjxb {

constexpr int NV = 27;    // compile-time dof count: the register-resident factorisation is statically unrolled
constexpr int NVP = 28;   // row stride (floats) of M and J: 16-byte aligned rows for 128-bit loads
// Main instantiation: candidate constraint rows / candidate contacts held in shared memory per env. An env that needs
// more is appended to an overflow list and re-run by the BIG instantiation (capacity >= every static row of the model).
#ifndef MJXB_CAP_MAIN
#define MJXB_CAP_MAIN 32
#endif
#ifndef MJXB_MAXCC_MAIN
#define MJXB_MAXCC_MAIN 16
#endif
#ifndef MJXB_WARPS_MAIN
#define MJXB_WARPS_MAIN 16
#endif
constexpr int CAP_MAIN = MJXB_CAP_MAIN, MAXCC_MAIN = MJXB_MAXCC_MAIN, WARPS_MAIN = MJXB_WARPS_MAIN;
constexpr int CAP_MID = 64, MAXCC_MID = 24, WARPS_MID = 10;
constexpr int CAP_BIG = 320, MAXCC_BIG = 176, WARPS_BIG = 3;
constexpr float MINVAL = 1e-15f;

enum { MODE_ENV_STEP = 0, MODE_ENV_RESET = 1, MODE_PHYS_STEP = 2, MODE_FORWARD = 3, MODE_SPEED_TEST = 4 };
enum { PAIR_PLANE_SPHERE = 0, PAIR_PLANE_CAPSULE = 1, PAIR_SPHERE_SPHERE = 2, PAIR_SPHERE_CAPSULE = 3, PAIR_CAPSULE_CAPSULE = 4 };
enum { ROW_LIMIT = 0, ROW_TENDON = 1, ROW_CON1 = 2, ROW_CON3 = 3 };
enum { QK_HINGE = 0, QK_FREEPOS = 1, QK_FREEQUAT = 2 };

struct PairParam { float mu, invweight, solref[2], solimp[5]; };

// Device copy of the model: the fields of mjxb_model_blob the kernels need, plus derived tables.
struct DevModel {
  int nq, nv, nu, nbody, njnt, ngeom, nsite, ntendon, nsensor, npair, ncon, nefc, nlimit, ntlimit, ncon1;
  int solver, iterations, ls_iterations, damp_implicit, maxdepth, ls_exact;
  float timestep, gravity[3], tolerance, ls_tolerance, meaninertia, total_mass;
  int body_parent[MJXB_MAXBODY], body_depth[MJXB_MAXBODY], body_subtree_end[MJXB_MAXBODY], body_jntadr[MJXB_MAXBODY],
      body_jntnum[MJXB_MAXBODY];
  uint32_t body_dofmask[MJXB_MAXBODY];  // dofs that move the body (ancestors-or-self)
  float body_pos[MJXB_MAXBODY][3], body_quat[MJXB_MAXBODY][4], body_ipos[MJXB_MAXBODY][3], body_inertia[MJXB_MAXBODY][6],
      body_mass[MJXB_MAXBODY];
  int jnt_type[MJXB_MAXJNT], jnt_qposadr[MJXB_MAXJNT], jnt_dofadr[MJXB_MAXJNT];
  int jnt_parent[MJXB_MAXJNT], jnt_first[MJXB_MAXJNT], jnt_bodyid[MJXB_MAXJNT];  // joint tree (previous joint up the chain), first joint of its body
  int body_srcjnt[MJXB_MAXBODY], body_lastdof[MJXB_MAXBODY];  // joint whose frame carries the body; last dof moving the body (-1: none)
  float body_relpos[MJXB_MAXBODY][3], body_relquat[MJXB_MAXBODY][4];  // fixed offset of the body frame from that joint frame
  int dof_cvel_src[MJXB_MAXDOF];  // dof whose inclusive velocity prefix is 'cvel before this dof' (-1: zero, -2: cdof_dot = 0)
  int tree_steps;  // pointer-jumping rounds covering the deepest joint / dof chain
  int tree_chol_ok;  // the model's dof tree is the one mjxb_chol_tree.cuh was generated for
  uint32_t dof_ancmask[MJXB_MAXDOF];  // dofs on the chain root..d (self included): the dofs whose motion moves dof d's axis
  float jnt_pos[MJXB_MAXJNT][3], jnt_axis[MJXB_MAXJNT][3];
  int lim_dof[MJXB_MAXJNT], lim_qadr[MJXB_MAXJNT], lim_row[MJXB_MAXJNT];
  float lim_range[MJXB_MAXJNT][2], lim_invweight[MJXB_MAXJNT], lim_solref[MJXB_MAXJNT][2], lim_solimp[MJXB_MAXJNT][5];
  int dof_body[MJXB_MAXDOF], dof_jnt[MJXB_MAXDOF], dof_parent[MJXB_MAXDOF], dof_qadr[MJXB_MAXDOF], dof_act[MJXB_MAXDOF];
  float dof_armature[MJXB_MAXDOF], dof_damping[MJXB_MAXDOF], dof_stiffness[MJXB_MAXDOF], dof_gear[MJXB_MAXDOF],
      dof_ctrl_lo[MJXB_MAXDOF], dof_ctrl_hi[MJXB_MAXDOF];
  int qpos_kind[MJXB_MAXQ], qpos_aux[MJXB_MAXQ];  // integrator addressing: hinge -> dof; free pos -> dof; free quat -> (qadr | comp<<8 | dof<<16)
  float qpos0[MJXB_MAXQ], qpos_spring[MJXB_MAXQ];
  int geom_body[MJXB_MAXGEOM];
  float geom_pos[MJXB_MAXGEOM][3], geom_axis[MJXB_MAXGEOM][3], geom_rad[MJXB_MAXGEOM], geom_half[MJXB_MAXGEOM];
  uint32_t pair_w0[MJXB_MAXPAIR];  // g1 | g2<<8 | kind<<16 | condim<<24
  uint32_t pair_w1[MJXB_MAXPAIR];  // conadr | efcadr<<16
  int ten_nwrap[MJXB_MAXTENDON], ten_dof[MJXB_MAXTENDON][MJXB_MAXWRAP], ten_qpos[MJXB_MAXTENDON][MJXB_MAXWRAP], ten_row[MJXB_MAXTENDON];
  float ten_coef[MJXB_MAXTENDON][MJXB_MAXWRAP], ten_range[MJXB_MAXTENDON][2], ten_solref[MJXB_MAXTENDON][2],
      ten_solimp[MJXB_MAXTENDON][5], ten_invweight[MJXB_MAXTENDON];
  int site_body[MJXB_MAXSITE], sensor_site[MJXB_MAXSENSOR];
  float site_pos[MJXB_MAXSITE][3], site_quat[MJXB_MAXSITE][4], site_size[MJXB_MAXSITE][3];
  mjxb_env_config cfg;
  int pad_[3];
};

inline int build_dev_model(const mjxb_model_blob& b, const mjxb_env_config* cfg, DevModel& D, PairParam* pp) {
  memset(&D, 0, sizeof(D));
  if (b.nv != NV) return MJXB_EUNSUPPORTED;  // the in-register factorisation is compiled for nv = 27 (humanoid family)
  if (b.nbody > 32 || b.ngeom > 32 || b.nq > 32 || b.nlimit > 32 || b.ntlimit > 32 || b.nsensor > MJXB_MAXSENSOR) return MJXB_EUNSUPPORTED;
  if (b.solver != 2 && b.solver != 1) return MJXB_EUNSUPPORTED;
  D.nq = b.nq; D.nv = b.nv; D.nu = b.nu; D.nbody = b.nbody; D.njnt = b.njnt; D.ngeom = b.ngeom; D.nsite = b.nsite;
  D.ntendon = b.ntendon; D.nsensor = b.nsensor; D.npair = b.npair; D.ncon = b.ncon; D.nefc = b.nefc; D.nlimit = b.nlimit;
  D.ntlimit = b.ntlimit; D.ncon1 = b.ncon1; D.solver = b.solver; D.iterations = b.iterations; D.ls_iterations = b.ls_iterations;
  D.damp_implicit = (b.integrator == 3) || (b.integrator == 0 && b.eulerdamp);
  D.maxdepth = b.maxdepth;
  // exact line search whenever MJX's own search is run to convergence; the truncated settings keep MJX's iteration
  D.ls_exact = (b.ls_iterations >= 10) ? 1 : 0;
  D.tree_chol_ok = 0;  // set by the caller that knows the generated elimination tree (mjxb_abi.cu)
  D.timestep = b.timestep; D.tolerance = b.tolerance; D.ls_tolerance = b.ls_tolerance; D.meaninertia = b.meaninertia;
  for (int k = 0; k < 3; k++) D.gravity[k] = b.gravity[k];
  double tm = 0;
  for (int i = 0; i < b.nbody; i++) {
    D.body_parent[i] = b.body_parent[i]; D.body_depth[i] = b.body_depth[i]; D.body_subtree_end[i] = b.body_subtree_end[i];
    D.body_jntadr[i] = b.body_jntadr[i]; D.body_jntnum[i] = b.body_jntnum[i];
    for (int k = 0; k < 3; k++) { D.body_pos[i][k] = b.body_pos[i][k]; D.body_ipos[i][k] = b.body_ipos[i][k]; }
    for (int k = 0; k < 4; k++) D.body_quat[i][k] = b.body_quat[i][k];
    for (int k = 0; k < 6; k++) D.body_inertia[i][k] = b.body_inertia[i][k];
    D.body_mass[i] = b.body_mass[i];
    tm += b.body_mass[i];
    if (i >= 1) {  // single kinematic tree rooted at body 1 (one subtree_com reference point)
      int r = i;
      while (b.body_parent[r] != 0) r = b.body_parent[r];
      if (r != 1) return MJXB_EUNSUPPORTED;
    }
  }
  D.total_mass = (float)tm;
  for (int j = 0; j < b.njnt; j++) {
    D.jnt_type[j] = b.jnt_type[j]; D.jnt_qposadr[j] = b.jnt_qposadr[j]; D.jnt_dofadr[j] = b.jnt_dofadr[j];
    for (int k = 0; k < 3; k++) { D.jnt_pos[j][k] = b.jnt_pos[j][k]; D.jnt_axis[j][k] = b.jnt_axis[j][k]; }
    if (b.jnt_type[j] != 0 && b.jnt_type[j] != 3) return MJXB_EUNSUPPORTED;
  }
  for (int i = 0; i < b.nlimit; i++) {
    int j = b.lim_jnt[i];
    D.lim_dof[i] = b.jnt_dofadr[j]; D.lim_qadr[i] = b.jnt_qposadr[j]; D.lim_row[i] = i;
    D.lim_range[i][0] = b.jnt_range[j][0]; D.lim_range[i][1] = b.jnt_range[j][1];
    D.lim_invweight[i] = b.dof_invweight0[b.jnt_dofadr[j]];
    for (int k = 0; k < 2; k++) D.lim_solref[i][k] = b.jnt_solref[j][k];
    for (int k = 0; k < 5; k++) D.lim_solimp[i][k] = b.jnt_solimp[j][k];
  }
  for (int d = 0; d < MJXB_MAXDOF; d++) { D.dof_act[d] = -1; D.dof_qadr[d] = -1; D.dof_parent[d] = -1; }
  for (int d = 0; d < b.nv; d++) {
    D.dof_ancmask[d] = 0;
    for (int a = d; a >= 0; a = b.dof_parent[a]) D.dof_ancmask[d] |= 1u << a;
    D.dof_body[d] = b.dof_body[d]; D.dof_jnt[d] = b.dof_jnt[d]; D.dof_parent[d] = b.dof_parent[d];
    D.dof_armature[d] = b.dof_armature[d]; D.dof_damping[d] = b.dof_damping[d]; D.dof_stiffness[d] = b.dof_stiffness[d];
    int j = b.dof_jnt[d];
    if (b.jnt_type[j] == 3) D.dof_qadr[d] = b.jnt_qposadr[j];
  }
  for (int u = 0; u < b.nu; u++) {
    int d = b.act_dof[u];
    if (D.dof_act[d] >= 0) return MJXB_EUNSUPPORTED;  // one motor per dof
    D.dof_act[d] = u; D.dof_gear[d] = b.act_gear[u];
    D.dof_ctrl_lo[d] = b.act_ctrllimited[u] ? b.act_ctrlrange[u][0] : -3.0e38f;
    D.dof_ctrl_hi[d] = b.act_ctrllimited[u] ? b.act_ctrlrange[u][1] : 3.0e38f;
  }
  for (int i = 0; i < b.nq; i++) { D.qpos0[i] = b.qpos0[i]; D.qpos_spring[i] = b.qpos_spring[i]; }
  for (int j = 0; j < b.njnt; j++) {
    int qa = b.jnt_qposadr[j], da = b.jnt_dofadr[j];
    if (b.jnt_type[j] == 0) {
      for (int k = 0; k < 3; k++) { D.qpos_kind[qa + k] = QK_FREEPOS; D.qpos_aux[qa + k] = da + k; }
      for (int k = 0; k < 4; k++) { D.qpos_kind[qa + 3 + k] = QK_FREEQUAT; D.qpos_aux[qa + 3 + k] = (qa + 3) | (k << 8) | ((da + 3) << 16); }
    } else {
      D.qpos_kind[qa] = QK_HINGE; D.qpos_aux[qa] = da;
    }
  }
  // joint tree + body/dof tables for the prefix-composition kinematics
  {
    int body_lastjnt[MJXB_MAXBODY];
    for (int i = 0; i < b.nbody; i++) body_lastjnt[i] = b.body_jntnum[i] > 0 ? b.body_jntadr[i] + b.body_jntnum[i] - 1 : -1;
    int maxchain = 1;
    for (int j = 0; j < b.njnt; j++) {
      const int bd = b.jnt_body[j];
      D.jnt_bodyid[j] = bd;
      D.jnt_first[j] = (j == b.body_jntadr[bd]) ? 1 : 0;
      int par = -1;
      if (!D.jnt_first[j]) par = j - 1;
      else {
        int a = b.body_parent[bd];
        // fixed offsets of joint-less bodies between bd and its nearest jointed ancestor are not supported for a FIRST joint
        if (a > 0 && body_lastjnt[a] < 0) return MJXB_EUNSUPPORTED;
        par = a > 0 ? body_lastjnt[a] : -1;
      }
      if (b.jnt_type[j] == 0 && (par >= 0 || b.body_parent[bd] != 0)) return MJXB_EUNSUPPORTED;  // free joints only on top-level bodies
      D.jnt_parent[j] = par;
    }
    for (int j = 0; j < b.njnt; j++) { int n = 1; for (int a = D.jnt_parent[j]; a >= 0; a = D.jnt_parent[a]) n++; if (n > maxchain) maxchain = n; }
    for (int d = 0; d < b.nv; d++) { int n = 1; for (int a = b.dof_parent[d]; a >= 0; a = b.dof_parent[a]) n++; if (n > maxchain) maxchain = n; }
    D.tree_steps = 0;
    while ((1 << D.tree_steps) < maxchain) D.tree_steps++;
    for (int i = 0; i < b.nbody; i++) {
      // body frame = frame after joint srcjnt composed with (relpos, relquat); joint-less bodies accumulate their fixed offsets
      double rp[3] = {0, 0, 0}, rq[4] = {1, 0, 0, 0};
      int a = i;
      while (a > 0 && body_lastjnt[a] < 0) {  // prepend body a's offset: T_a o (rp, rq)
        const double w = b.body_quat[a][0], x = b.body_quat[a][1], y = b.body_quat[a][2], z = b.body_quat[a][3];
        const double R[9] = {w * w + x * x - y * y - z * z, 2 * (x * y - w * z), 2 * (x * z + w * y), 2 * (x * y + w * z), w * w - x * x + y * y - z * z,
                             2 * (y * z - w * x), 2 * (x * z - w * y), 2 * (y * z + w * x), w * w - x * x - y * y + z * z};
        const double np[3] = {b.body_pos[a][0] + R[0] * rp[0] + R[1] * rp[1] + R[2] * rp[2], b.body_pos[a][1] + R[3] * rp[0] + R[4] * rp[1] + R[5] * rp[2],
                              b.body_pos[a][2] + R[6] * rp[0] + R[7] * rp[1] + R[8] * rp[2]};
        const double nq[4] = {w * rq[0] - x * rq[1] - y * rq[2] - z * rq[3], w * rq[1] + x * rq[0] + y * rq[3] - z * rq[2],
                              w * rq[2] - x * rq[3] + y * rq[0] + z * rq[1], w * rq[3] + x * rq[2] - y * rq[1] + z * rq[0]};
        for (int k = 0; k < 3; k++) rp[k] = np[k];
        for (int k = 0; k < 4; k++) rq[k] = nq[k];
        a = b.body_parent[a];
      }
      D.body_srcjnt[i] = a > 0 ? body_lastjnt[a] : -1;
      for (int k = 0; k < 3; k++) D.body_relpos[i][k] = (float)rp[k];
      for (int k = 0; k < 4; k++) D.body_relquat[i][k] = (float)rq[k];
      int bb = i;
      while (bb > 0 && b.body_dofnum[bb] == 0) bb = b.body_parent[bb];
      D.body_lastdof[i] = bb > 0 ? b.body_dofadr[bb] + b.body_dofnum[bb] - 1 : -1;
    }
    for (int d = 0; d < MJXB_MAXDOF; d++) D.dof_cvel_src[d] = -2;
    for (int d = 0; d < b.nv; d++) {
      const int j = b.dof_jnt[d];
      if (b.jnt_type[j] == 0) {
        const int k = d - b.jnt_dofadr[j];
        D.dof_cvel_src[d] = k < 3 ? -2 : b.jnt_dofadr[j] + 2;   // linear: cdof_dot = 0; angular: velocity after the three linear dofs
      } else {
        D.dof_cvel_src[d] = b.dof_parent[d];                    // -1: nothing moves before this dof
      }
    }
  }
  // dofs that move each body: walk the dof-parent chain from the body's (or nearest jointed ancestor's) last dof
  for (int i = 1; i < b.nbody; i++) {
    int bb = i;
    while (bb > 0 && b.body_dofnum[bb] == 0) bb = b.body_parent[bb];
    uint32_t mask = 0;
    if (bb > 0)
      for (int d = b.body_dofadr[bb] + b.body_dofnum[bb] - 1; d >= 0; d = b.dof_parent[d]) mask |= 1u << d;
    D.body_dofmask[i] = mask;
  }
  for (int g = 0; g < b.ngeom; g++) {
    D.geom_body[g] = b.geom_body[g];
    for (int k = 0; k < 3; k++) D.geom_pos[g][k] = b.geom_pos[g][k];
    // local z axis of the geom frame (third column of the rotation of geom_quat), in double
    double w = b.geom_quat[g][0], x = b.geom_quat[g][1], y = b.geom_quat[g][2], z = b.geom_quat[g][3];
    D.geom_axis[g][0] = (float)(2 * (x * z + w * y)); D.geom_axis[g][1] = (float)(2 * (y * z - w * x));
    D.geom_axis[g][2] = (float)(w * w - x * x - y * y + z * z);
    D.geom_rad[g] = b.geom_size[g][0]; D.geom_half[g] = b.geom_size[g][1];
    if (b.geom_type[g] == 0) { D.geom_rad[g] = 0.0f; D.geom_half[g] = 0.0f; }
  }
  if (b.npair > MJXB_MAXPAIR || b.ncon > MAXCC_BIG || b.nefc > CAP_BIG) return MJXB_EUNSUPPORTED;
  for (int p = 0; p < b.npair; p++) {
    D.pair_w0[p] = (uint32_t)b.pair_g1[p] | ((uint32_t)b.pair_g2[p] << 8) | ((uint32_t)b.pair_kind[p] << 16) | ((uint32_t)b.pair_condim[p] << 24);
    {  // bit 31: the two bodies sit on different limbs (neither dof chain contains the other): such a row breaks the tree pattern of H
      const uint32_t m1 = D.body_dofmask[b.geom_body[b.pair_g1[p]]], m2 = D.body_dofmask[b.geom_body[b.pair_g2[p]]];
      if ((m1 & m2) != m1 && (m1 & m2) != m2) D.pair_w0[p] |= 0x80000000u;
    }
    D.pair_w1[p] = (uint32_t)b.pair_conadr[p] | ((uint32_t)b.pair_efcadr[p] << 16);
    pp[p].mu = b.pair_mu[p]; pp[p].invweight = b.pair_invweight[p];
    for (int k = 0; k < 2; k++) pp[p].solref[k] = b.pair_solref[p][k];
    for (int k = 0; k < 5; k++) pp[p].solimp[k] = b.pair_solimp[p][k];
    if (b.pair_condim[p] != 1 && b.pair_condim[p] != 3) return MJXB_EUNSUPPORTED;
  }
  for (int i = 0; i < b.ntlimit; i++) {
    int t = b.lim_ten[i];
    D.ten_nwrap[i] = b.ten_nwrap[t]; D.ten_row[i] = b.nlimit + i;
    for (int w = 0; w < MJXB_MAXWRAP; w++) { D.ten_dof[i][w] = b.ten_dof[t][w]; D.ten_qpos[i][w] = b.ten_qpos[t][w]; D.ten_coef[i][w] = b.ten_coef[t][w]; }
    for (int k = 0; k < 2; k++) { D.ten_range[i][k] = b.ten_range[t][k]; D.ten_solref[i][k] = b.ten_solref[t][k]; }
    for (int k = 0; k < 5; k++) D.ten_solimp[i][k] = b.ten_solimp[t][k];
    D.ten_invweight[i] = b.ten_invweight0[t];
  }
  for (int s = 0; s < b.nsite; s++) {
    D.site_body[s] = b.site_body[s];
    for (int k = 0; k < 3; k++) { D.site_pos[s][k] = b.site_pos[s][k]; D.site_size[s][k] = b.site_size[s][k]; }
    for (int k = 0; k < 4; k++) D.site_quat[s][k] = b.site_quat[s][k];
  }
  for (int s = 0; s < b.nsensor; s++) D.sensor_site[s] = b.sensor_site[s];
  if (cfg) {
    D.cfg = *cfg;
    if (cfg->obs_dim != 1 + 3 + (b.nq - 7) + b.nv + 2 || cfg->obs_dim > MJXB_MAXOBS) return MJXB_EINVAL;
    if (cfg->pelvis_body_id < 0 || cfg->pelvis_body_id >= b.nbody || cfg->head_body_id < 0 || cfg->head_body_id >= b.nbody) return MJXB_EINVAL;
    if (cfg->touch_sensor_right_id < 0 || cfg->touch_sensor_right_id >= b.nsensor || cfg->touch_sensor_left_id < 0 ||
        cfg->touch_sensor_left_id >= b.nsensor) return MJXB_EINVAL;
    for (int i = 0; i < b.nu; i++) if (cfg->act_perm[i] < 0 || cfg->act_perm[i] >= b.nu) return MJXB_EINVAL;
    for (int i = 0; i < cfg->obs_dim; i++) if (cfg->obs_perm[i] < 0 || cfg->obs_perm[i] >= cfg->obs_dim) return MJXB_EINVAL;
  } else {
    D.cfg.obs_dim = 1 + 3 + (b.nq - 7) + b.nv + 2;
    D.cfg.pelvis_body_id = 0; D.cfg.head_body_id = 0;
  }
  return MJXB_OK;
}

}  // namespace mjxb
