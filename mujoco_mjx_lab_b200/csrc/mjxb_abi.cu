// mjxb_abi.cu -- the extern "C" boundary declared in include/mjxb.h (host side: model upload, launches, host-buffer arena).
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>

#include "mjxb.h"
#include "mjxb_device.cuh"

using namespace mjxb;

namespace {

constexpr int kResetPad = 148 * 16 * 2;  // per-CTA reset queues are sized in whole rounds: n_env + (CTAs x warps) entries at most
thread_local char g_cuda_err[512] = "";

int cuda_fail(cudaError_t e, const char* what) {
  snprintf(g_cuda_err, sizeof(g_cuda_err), "%s: %s", what, cudaGetErrorString(e));
  return MJXB_ECUDA;
}
#define CU(call)                                         \
  do {                                                   \
    cudaError_t e_ = (call);                             \
    if (e_ != cudaSuccess) return cuda_fail(e_, #call);  \
  } while (0)

struct Arena {  // device-resident env batch for the *_host entry points
  int n = 0;
  float *qpos = nullptr, *qvel = nullptr, *warm = nullptr, *time = nullptr, *aux = nullptr;
  float *action = nullptr, *obs = nullptr, *reward = nullptr, *term = nullptr, *trunc = nullptr;
  uint32_t* keys = nullptr;
  cudaStream_t stream = nullptr;
  // chunked host pipeline: H2D / kernel / D2H of consecutive env chunks overlap on separate streams, each with its own overflow list
  static constexpr int kSlots = 3;
  cudaStream_t pipe[kSlots] = {nullptr, nullptr, nullptr};
  int* pipe_ovf[kSlots] = {nullptr, nullptr, nullptr};
  int chunk = 0;
  // direct pipeline (pinned caller buffers): one launch; inputs stream in behind per-chunk ready flags, obs go straight to the host
  static constexpr int kReadyMax = 64;
  unsigned* ready = nullptr;      // device flags [kReadyMax]
  unsigned* ready_src = nullptr;  // pinned source of the flag copies
  unsigned epoch = 0;
  int ready_shift = 0;
};

}  // namespace

struct mjxb_model {
  DevModel host;
  DevModel* dev = nullptr;
  PairParam* dev_pp = nullptr;
  int device = 0, num_sms = 0, warps = 0, warps_mid = 0, warps_big = 0;
  size_t smem = 0, smem_mid = 0, smem_big = 0;
  int* ovf = nullptr;   // [0] countA, [1] doneA, [2] countB, [3] doneB, [4..4+cap) listA (main -> mid), [4+cap..) listB (mid -> big)
  int ovf_cap = 0;
  Arena arena;
};

namespace {

void arena_free(Arena& a) {
  float** ps[] = {&a.qpos, &a.qvel, &a.warm, &a.time, &a.aux, &a.action, &a.obs, &a.reward, &a.term, &a.trunc};
  for (float** p : ps) { if (*p) cudaFree(*p); *p = nullptr; }
  if (a.keys) cudaFree(a.keys);
  a.keys = nullptr;
  if (a.ready) cudaFree(a.ready);
  if (a.ready_src) cudaFreeHost(a.ready_src);
  a.ready = nullptr; a.ready_src = nullptr;
  if (a.stream) cudaStreamDestroy(a.stream);
  a.stream = nullptr;
  for (int i = 0; i < Arena::kSlots; i++) {
    if (a.pipe[i]) cudaStreamDestroy(a.pipe[i]);
    if (a.pipe_ovf[i]) cudaFree(a.pipe_ovf[i]);
    a.pipe[i] = nullptr; a.pipe_ovf[i] = nullptr;
  }
  a.n = 0;
}

int arena_ensure(mjxb_model* m, int n) {
  Arena& a = m->arena;
  if (a.n == n) return MJXB_OK;
  arena_free(a);
  const DevModel& C = m->host;
  CU(cudaSetDevice(m->device));
  CU(cudaStreamCreateWithFlags(&a.stream, cudaStreamNonBlocking));
  size_t N = (size_t)n;
  CU(cudaMalloc(&a.qpos, N * C.nq * 4)); CU(cudaMalloc(&a.qvel, N * C.nv * 4)); CU(cudaMalloc(&a.warm, N * C.nv * 4));
  CU(cudaMalloc(&a.time, N * 4)); CU(cudaMalloc(&a.aux, N * MJXB_AUX_DIM * 4)); CU(cudaMalloc(&a.action, N * C.nu * 4));
  CU(cudaMalloc(&a.obs, N * C.cfg.obs_dim * 4)); CU(cudaMalloc(&a.reward, N * 4)); CU(cudaMalloc(&a.term, N * 4));
  CU(cudaMalloc(&a.trunc, N * 4)); CU(cudaMalloc(&a.keys, N * 8));
  CU(cudaMemsetAsync(a.qpos, 0, N * C.nq * 4, a.stream)); CU(cudaMemsetAsync(a.qvel, 0, N * C.nv * 4, a.stream));
  CU(cudaMemsetAsync(a.warm, 0, N * C.nv * 4, a.stream)); CU(cudaMemsetAsync(a.time, 0, N * 4, a.stream));
  CU(cudaMemsetAsync(a.aux, 0, N * MJXB_AUX_DIM * 4, a.stream));
  // chunks of >= 65536 envs, at most 4 per step (each chunk still fills every SM for many rounds; every chunk pays its own
  // overflow-consume launch, so more chunks stop paying off)
  int nchunk = n / 65536;
  if (nchunk < 1) nchunk = 1;
  if (nchunk > 4) nchunk = 4;
  { const char* e = getenv("MJXB_HOST_CHUNKS"); if (e && atoi(e) > 0) nchunk = atoi(e); }
  a.chunk = (n + nchunk - 1) / nchunk;
  for (int i = 0; i < Arena::kSlots; i++) {
    CU(cudaStreamCreateWithFlags(&a.pipe[i], cudaStreamNonBlocking));
    CU(cudaMalloc(&a.pipe_ovf[i], (3 * (size_t)a.chunk + 4 + kResetPad) * sizeof(int)));
    CU(cudaMemsetAsync(a.pipe_ovf[i], 0, 4 * sizeof(int), a.stream));
  }
  CU(cudaMalloc(&a.ready, Arena::kReadyMax * sizeof(unsigned)));
  CU(cudaMemsetAsync(a.ready, 0, Arena::kReadyMax * sizeof(unsigned), a.stream));
  CU(cudaHostAlloc(&a.ready_src, (Arena::kReadyMax + 1) * sizeof(unsigned), cudaHostAllocDefault));  // [kReadyMax]: timeout flag
  a.ready_src[Arena::kReadyMax] = 0u;
  a.epoch = 0;
  a.ready_shift = 12;  // input chunks of 2^shift envs: about 8 per step, never more than kReadyMax
  while (((n + (1 << a.ready_shift) - 1) >> a.ready_shift) > 8) a.ready_shift++;
  CU(cudaStreamSynchronize(a.stream));
  a.n = n;
  return MJXB_OK;
}

int build_dev_model(const mjxb_model_blob& b, const mjxb_env_config* cfg, DevModel& D, PairParam* pp) {
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
  D.tree_chol_ok = (b.nv == kTreeNV) ? 1 : 0;
  for (int d = 0; d < b.nv && d < kTreeNV; d++) if (b.dof_parent[d] != kTreeDofParent[d]) D.tree_chol_ok = 0;
  if (getenv("MJXB_DENSE_CHOL")) D.tree_chol_ok = 0;
  if (getenv("MJXB_LS_ITERATIVE")) D.ls_exact = 0;
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

using KMain = void (*)(const DevModel*, const PairParam*, StepArgs);

int launch(const mjxb_model* mc, const StepArgs& args_in, bool dbg, cudaStream_t stream, int* ovf_buf = nullptr, int ovf_buf_cap = 0) {
  mjxb_model* m = const_cast<mjxb_model*>(mc);  // the overflow list is library-owned scratch, grown on first use for a batch size
  int cur = 0;
  CU(cudaGetDevice(&cur));
  if (cur != m->device) CU(cudaSetDevice(m->device));
  if (ovf_buf == nullptr && m->ovf_cap < args_in.n_env) {
    if (m->ovf) { CU(cudaStreamSynchronize(stream)); CU(cudaFree(m->ovf)); m->ovf = nullptr; }
    CU(cudaMalloc(&m->ovf, (3 * (size_t)args_in.n_env + 4 + kResetPad) * sizeof(int)));
    CU(cudaMemsetAsync(m->ovf, 0, 4 * sizeof(int), stream));
    m->ovf_cap = args_in.n_env;
  }
  StepArgs args = args_in;
  int* ovf = ovf_buf ? ovf_buf : m->ovf;
  const int cap = ovf_buf ? ovf_buf_cap : m->ovf_cap;
  int* listA = ovf + 4;
  int* listB = ovf + 4 + cap;
  args.in_count = nullptr; args.in_list = nullptr; args.in_done = nullptr; args.out_count = ovf; args.out_list = listA;
  args.reset_list = (getenv("MJXB_INLINE_RESET") != nullptr) ? nullptr : ovf + 4 + 2 * (size_t)cap;
  // 1 (default): CTA barriers at the round top and at the solver entry / exit; 3: additionally at every factor/solve round (was the
  // better choice before the solver's dependent chains were shortened: 35.2 M against 36.4 M now); 0: none (profiling aid)
  { const char* e = getenv("MJXB_LOCKSTEP"); args.lockstep = e ? atoi(e) : 1; }
  { const char* e = getenv("MJXB_LOCKSTEP_GROUP"); args.lockstep_group = e ? atoi(e) : 0; }
  // small batches: spread the envs over every SM (fewer warps per CTA run faster than 16 sharing one SM's issue slots)
  int warps = m->warps;
  const int per_sm = (args.n_env + m->num_sms - 1) / m->num_sms;
  if (per_sm < warps) warps = per_sm < 1 ? 1 : per_sm;
  int grid = (args.n_env + warps - 1) / warps;
  if (grid > m->num_sms) grid = m->num_sms;
  const size_t smem_main = m->smem - (size_t)(m->warps - warps) * sizeof(WarpS<CAP_MAIN, MAXCC_MAIN>);
  args.reset_stride = ((args.n_env + grid * warps - 1) / (grid * warps)) * warps;
  const bool ls = m->host.ls_exact != 0 && m->host.solver == 2;  // fast instantiation: Newton + exact line search; else the general one
#define MJXB_LAUNCH(CAPv, CCv, Wv, G, B, SM)                                                                                   \
  do {                                                                                                                        \
    if (dbg && ls) mjxb_step_kernel<true, CAPv, CCv, Wv, true><<<G, B, SM, stream>>>(m->dev, m->dev_pp, args);                  \
    else if (dbg) mjxb_step_kernel<true, CAPv, CCv, Wv, false><<<G, B, SM, stream>>>(m->dev, m->dev_pp, args);                  \
    else if (ls) mjxb_step_kernel<false, CAPv, CCv, Wv, true><<<G, B, SM, stream>>>(m->dev, m->dev_pp, args);                   \
    else mjxb_step_kernel<false, CAPv, CCv, Wv, false><<<G, B, SM, stream>>>(m->dev, m->dev_pp, args);                          \
  } while (0)
  // hot instantiation: one step per launch, Newton + exact line search, resets deferred (or none requested)
  const bool single = !dbg && ls && args.nsteps == 1 && (!(args.mode == MODE_ENV_STEP && args.autoreset) || args.reset_list != nullptr);
  if (single) mjxb_step_kernel<false, CAP_MAIN, MAXCC_MAIN, WARPS_MAIN, true, true><<<grid, warps * 32, smem_main, stream>>>(m->dev, m->dev_pp, args);
  else MJXB_LAUNCH(CAP_MAIN, MAXCC_MAIN, WARPS_MAIN, grid, warps * 32, smem_main);
  cudaError_t e = cudaGetLastError();
  const bool sync_tiers = getenv("MJXB_SYNC_TIERS") != nullptr;  // debugging aid: attribute a device fault to its tier
  if (sync_tiers && e == cudaSuccess) { e = cudaStreamSynchronize(stream); if (e != cudaSuccess) fprintf(stderr, "[mjxb] main tier failed: %s\n", cudaGetErrorString(e)); }
  args.in_ready = nullptr;  // only the first pass waits for streamed inputs
  if (e == cudaSuccess) {  // mid tier (64 rows / 24 contacts) over the envs the main tile could not hold; usually few: exits at once when empty
    args.in_count = ovf; args.in_done = ovf + 1; args.in_list = listA; args.out_count = ovf + 2; args.out_list = listB;
    const int wm = m->warps_mid;
    int gridm = m->num_sms;
    if (gridm * wm > args.n_env) gridm = (args.n_env + wm - 1) / wm;
    args.reset_stride = ((args.n_env + gridm * wm - 1) / (gridm * wm)) * wm;
    MJXB_LAUNCH(CAP_MID, MAXCC_MID, WARPS_MID, gridm, wm * 32, m->smem_mid);
    e = cudaGetLastError();
    if (sync_tiers && e == cudaSuccess) { e = cudaStreamSynchronize(stream); if (e != cudaSuccess) fprintf(stderr, "[mjxb] mid tier failed: %s\n", cudaGetErrorString(e)); }
  }
  if (e == cudaSuccess) {  // big tier: holds every static row / contact slot of the model
    args.in_count = ovf + 2; args.in_done = ovf + 3; args.in_list = listB; args.out_count = nullptr; args.out_list = nullptr;
    const int wb = m->warps_big;
    int gridb = m->num_sms;
    if (gridb * wb > args.n_env) gridb = (args.n_env + wb - 1) / wb;
    args.reset_stride = ((args.n_env + gridb * wb - 1) / (gridb * wb)) * wb;
    MJXB_LAUNCH(CAP_BIG, MAXCC_BIG, WARPS_BIG, gridb, wb * 32, m->smem_big);
#undef MJXB_LAUNCH
    e = cudaGetLastError();
    if (sync_tiers && e == cudaSuccess) { e = cudaStreamSynchronize(stream); if (e != cudaSuccess) fprintf(stderr, "[mjxb] big tier failed: %s\n", cudaGetErrorString(e)); }
  }
  if (cur != m->device) cudaSetDevice(cur);
  if (e != cudaSuccess) return cuda_fail(e, "mjxb_step_kernel launch");
  return MJXB_OK;
}

bool state_ok(const mjxb_state& s, bool need_aux) {
  return s.qpos && s.qvel && s.qacc_warmstart && s.time && (!need_aux || s.aux);
}

}  // namespace

extern "C" {

int mjxb_abi_version(void) { return MJXB_ABI_VERSION; }
size_t mjxb_blob_sizeof(void) { return sizeof(mjxb_model_blob); }
size_t mjxb_env_config_sizeof(void) { return sizeof(mjxb_env_config); }
const char* mjxb_last_cuda_error(void) { return g_cuda_err; }
const char* mjxb_strerror(int code) {
  switch (code) {
    case MJXB_OK: return "ok";
    case MJXB_EINVAL: return "invalid argument";
    case MJXB_EBLOB: return "bad model blob (magic/version/size)";
    case MJXB_ECUDA: return "CUDA error (see mjxb_last_cuda_error)";
    case MJXB_ENOGPU: return "no CUDA device (there is no CPU fallback)";
    case MJXB_EUNSUPPORTED: return "model outside the compiled humanoid family (nv=27, hinge/free joints, Newton, pyramidal)";
    default: return "unknown error";
  }
}

int mjxb_model_create(const void* blob, size_t blob_bytes, const mjxb_env_config* cfg, int device, mjxb_model** out) {
  if (!blob || !out) return MJXB_EINVAL;
  *out = nullptr;
  if (blob_bytes != sizeof(mjxb_model_blob)) return MJXB_EBLOB;
  const mjxb_model_blob& b = *reinterpret_cast<const mjxb_model_blob*>(blob);
  if (b.magic != MJXB_BLOB_MAGIC || b.version != MJXB_BLOB_VERSION) return MJXB_EBLOB;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); return MJXB_ENOGPU; }
  if (device < 0 || device >= ndev) return MJXB_EINVAL;
  mjxb_model* m = new (std::nothrow) mjxb_model();
  if (!m) return MJXB_EINVAL;
  static PairParam pp[MJXB_MAXPAIR];
  static std::mutex mu;
  std::lock_guard<std::mutex> lock(mu);
  memset(pp, 0, sizeof(pp));
  int rc = build_dev_model(b, cfg, m->host, pp);
  if (rc != MJXB_OK) { delete m; return rc; }
  m->device = device;
  cudaError_t e;
#define CUX(call) if ((e = (call)) != cudaSuccess) { cuda_fail(e, #call); mjxb_model_destroy(m); return MJXB_ECUDA; }
  int cur = 0;
  CUX(cudaGetDevice(&cur));
  CUX(cudaSetDevice(device));
  cudaDeviceProp prop;
  CUX(cudaGetDeviceProperties(&prop, device));
  m->num_sms = prop.multiProcessorCount;
  const size_t model_bytes = (sizeof(DevModel) + 15) & ~size_t(15);
  const size_t avail = prop.sharedMemPerBlockOptin;
  using WSMain = WarpS<CAP_MAIN, MAXCC_MAIN>;
  using WSMid = WarpS<CAP_MID, MAXCC_MID>;
  using WSBig = WarpS<CAP_BIG, MAXCC_BIG>;
  int warps = (int)((avail - model_bytes) / sizeof(WSMain)), warps_mid = (int)((avail - model_bytes) / sizeof(WSMid)),
      warps_big = (int)((avail - model_bytes) / sizeof(WSBig));
  if (warps > WARPS_MAIN) warps = WARPS_MAIN;
  if (warps_mid > WARPS_MID) warps_mid = WARPS_MID;
  if (warps_big > WARPS_BIG) warps_big = WARPS_BIG;
  if (warps < 1 || warps_mid < 1 || warps_big < 1) { delete m; cudaSetDevice(cur); return MJXB_EUNSUPPORTED; }
  m->warps = warps; m->warps_mid = warps_mid; m->warps_big = warps_big;
  m->smem = model_bytes + (size_t)warps * sizeof(WSMain);
  m->smem_mid = model_bytes + (size_t)warps_mid * sizeof(WSMid);
  m->smem_big = model_bytes + (size_t)warps_big * sizeof(WSBig);
#define MJXB_SMEM_ATTR(DBGv, LSv)                                                                                                         \
  CUX(cudaFuncSetAttribute(mjxb_step_kernel<DBGv, CAP_MAIN, MAXCC_MAIN, WARPS_MAIN, LSv>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)m->smem)); \
  CUX(cudaFuncSetAttribute(mjxb_step_kernel<DBGv, CAP_MID, MAXCC_MID, WARPS_MID, LSv>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)m->smem_mid)); \
  CUX(cudaFuncSetAttribute(mjxb_step_kernel<DBGv, CAP_BIG, MAXCC_BIG, WARPS_BIG, LSv>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)m->smem_big));
  MJXB_SMEM_ATTR(false, true) MJXB_SMEM_ATTR(true, true) MJXB_SMEM_ATTR(false, false) MJXB_SMEM_ATTR(true, false)
  CUX(cudaFuncSetAttribute(mjxb_step_kernel<false, CAP_MAIN, MAXCC_MAIN, WARPS_MAIN, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)m->smem));
#undef MJXB_SMEM_ATTR
  CUX(cudaMalloc(&m->dev, sizeof(DevModel)));
  CUX(cudaMalloc(&m->dev_pp, sizeof(pp)));
  CUX(cudaMemcpy(m->dev, &m->host, sizeof(DevModel), cudaMemcpyHostToDevice));
  CUX(cudaMemcpy(m->dev_pp, pp, sizeof(pp), cudaMemcpyHostToDevice));
  cudaSetDevice(cur);
#undef CUX
  *out = m;
  return MJXB_OK;
}

void mjxb_model_destroy(mjxb_model* m) {
  if (!m) return;
  arena_free(m->arena);
  if (m->dev) cudaFree(m->dev);
  if (m->dev_pp) cudaFree(m->dev_pp);
  if (m->ovf) cudaFree(m->ovf);
  delete m;
}

int mjxb_model_dims(const mjxb_model* m, int32_t dims[8]) {
  if (!m || !dims) return MJXB_EINVAL;
  const DevModel& C = m->host;
  dims[0] = C.nq; dims[1] = C.nv; dims[2] = C.nu; dims[3] = C.nbody; dims[4] = C.ncon; dims[5] = C.nefc; dims[6] = C.nsensor;
  dims[7] = C.cfg.obs_dim;
  return MJXB_OK;
}

size_t mjxb_model_scratch_bytes(const mjxb_model* m) { return m ? (3 * (size_t)m->ovf_cap + 4 + kResetPad) * sizeof(int) : 0; }

#if MJXB_STAGE_CLOCK
int mjxb_debug_stage_clock(int32_t* host_out, int32_t n_env) {  // profiling variant only (not part of include/mjxb.h)
  return cudaMemcpyFromSymbol(host_out, ::g_stage_clock, sizeof(int) * 32 * (size_t)(n_env < 4096 ? n_env : 4096)) == cudaSuccess ? 0 : MJXB_ECUDA;
}
#endif

int mjxb_launch_config(const mjxb_model* m, int32_t cfg[4]) {  // warps per CTA, dynamic smem bytes, SM count, sizeof(WarpS)
  if (!m || !cfg) return MJXB_EINVAL;
  cfg[0] = m->warps; cfg[1] = (int32_t)m->smem; cfg[2] = m->num_sms; cfg[3] = (int32_t)sizeof(WarpS<CAP_MAIN, MAXCC_MAIN>);
  return MJXB_OK;
}

int mjxb_reset(const mjxb_model* m, int32_t n_env, const uint32_t* keys, mjxb_state out, float* obs, int32_t* status, void* stream) {
  if (!m || n_env <= 0 || !keys || !obs || !state_ok(out, true)) return MJXB_EINVAL;
  StepArgs a;
  memset(&a, 0, sizeof(a));
  a.n_env = n_env; a.mode = MODE_ENV_RESET; a.nsteps = 1; a.out = out; a.in = out; a.keys = keys; a.obs = obs; a.status = status;
  return launch(m, a, false, (cudaStream_t)stream);
}

int mjxb_step(const mjxb_model* m, int32_t n_env, mjxb_state in, const float* action, mjxb_state out, float* obs, float* reward,
              float* terminated, float* truncated, int32_t* status, void* stream) {
  if (!m || n_env <= 0 || !action || !obs || !reward || !terminated || !truncated || !state_ok(in, true) || !state_ok(out, true))
    return MJXB_EINVAL;
  StepArgs a;
  memset(&a, 0, sizeof(a));
  a.n_env = n_env; a.mode = MODE_ENV_STEP; a.nsteps = 1; a.in = in; a.out = out; a.action = action; a.obs = obs; a.reward = reward;
  a.terminated = terminated; a.truncated = truncated; a.status = status;
  return launch(m, a, false, (cudaStream_t)stream);
}

int mjxb_step_autoreset(const mjxb_model* m, int32_t n_env, mjxb_state in, const float* action, const uint32_t* keys, mjxb_state out,
                        float* obs, float* reward, float* terminated, float* truncated, uint8_t* reset_mask, int32_t* status,
                        void* stream) {
  if (!m || n_env <= 0 || !action || !keys || !obs || !reward || !terminated || !truncated || !state_ok(in, true) || !state_ok(out, true))
    return MJXB_EINVAL;
  StepArgs a;
  memset(&a, 0, sizeof(a));
  a.n_env = n_env; a.mode = MODE_ENV_STEP; a.nsteps = 1; a.autoreset = 1; a.in = in; a.out = out; a.action = action; a.keys = keys;
  a.obs = obs; a.reward = reward; a.terminated = terminated; a.truncated = truncated; a.reset_mask = reset_mask; a.status = status;
  return launch(m, a, false, (cudaStream_t)stream);
}

int mjxb_physics_step(const mjxb_model* m, int32_t n_env, mjxb_state io, const float* ctrl, int32_t nsteps, const mjxb_debug* dbg,
                      int32_t* status, void* stream) {
  if (!m || n_env <= 0 || nsteps <= 0 || !state_ok(io, false)) return MJXB_EINVAL;
  StepArgs a;
  memset(&a, 0, sizeof(a));
  a.n_env = n_env; a.mode = MODE_PHYS_STEP; a.nsteps = nsteps; a.in = io; a.out = io; a.action = ctrl; a.status = status;
  if (dbg) a.dbg = *dbg;
  return launch(m, a, dbg != nullptr, (cudaStream_t)stream);
}

int mjxb_forward(const mjxb_model* m, int32_t n_env, mjxb_state io, const float* ctrl, const mjxb_debug* dbg, int32_t* status,
                 void* stream) {
  if (!m || n_env <= 0 || !state_ok(io, false)) return MJXB_EINVAL;
  StepArgs a;
  memset(&a, 0, sizeof(a));
  a.n_env = n_env; a.mode = MODE_FORWARD; a.nsteps = 1; a.in = io; a.out = io; a.action = ctrl; a.status = status;
  if (dbg) a.dbg = *dbg;
  return launch(m, a, dbg != nullptr, (cudaStream_t)stream);
}

int mjxb_speed_test(const mjxb_model* m, int32_t n_env, const float* vel, float* pos, int32_t iters, void* stream) {
  if (!m || n_env <= 0 || !vel || !pos || iters <= 0) return MJXB_EINVAL;
  StepArgs a;
  memset(&a, 0, sizeof(a));
  a.n_env = n_env; a.mode = MODE_SPEED_TEST; a.nsteps = iters; a.vel = vel; a.pos = pos;
  return launch(m, a, false, (cudaStream_t)stream);
}

// ---------------------------------------------------------------- host-buffer variants (the end-to-end path)
static mjxb_state arena_state(const Arena& a) {
  mjxb_state s;
  s.qpos = a.qpos; s.qvel = a.qvel; s.qacc_warmstart = a.warm; s.time = a.time; s.aux = a.aux;
  return s;
}

int mjxb_reset_host(mjxb_model* m, int32_t n_env, const uint32_t* keys_host, float* obs_host) {
  if (!m || n_env <= 0 || !keys_host || !obs_host) return MJXB_EINVAL;
  int rc = arena_ensure(m, n_env);
  if (rc) return rc;
  Arena& a = m->arena;
  size_t N = (size_t)n_env;
  CU(cudaMemcpyAsync(a.keys, keys_host, N * 8, cudaMemcpyHostToDevice, a.stream));
  rc = mjxb_reset(m, n_env, a.keys, arena_state(a), a.obs, nullptr, a.stream);
  if (rc) return rc;
  CU(cudaMemcpyAsync(obs_host, a.obs, N * m->host.cfg.obs_dim * 4, cudaMemcpyDeviceToHost, a.stream));
  CU(cudaStreamSynchronize(a.stream));
  return MJXB_OK;
}

// device-visible alias of a pinned / registered host pointer (UVA); false for pageable memory
static bool host_dev_ptr(const void* p, void** dp) {
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
  if (at.type != cudaMemoryTypeHost || at.devicePointer == nullptr) return false;
  *dp = at.devicePointer;
  return true;
}

static int step_host_impl(mjxb_model* m, int32_t n_env, const float* action_host, const uint32_t* keys_host, float* obs_host,
                          float* reward_host, float* terminated_host, float* truncated_host) {
  if (!m || n_env <= 0 || !action_host || !obs_host || !reward_host || !terminated_host || !truncated_host) return MJXB_EINVAL;
  Arena& a = m->arena;
  if (a.n != n_env) return MJXB_EINVAL;  // reset_host / state_set_host must have created the batch
  const int nu = m->host.nu, od = m->host.cfg.obs_dim, nq = m->host.nq, nv = m->host.nv;
  CU(cudaStreamSynchronize(a.stream));  // state_set_host / reset_host ran on a.stream
  // Direct pipeline when the caller's buffers are pinned (device-visible): ONE launch over the whole batch; action / keys stream in
  // on the copy stream in ~8 chunks, each followed by a 4-byte flag copy the kernel waits on (copy engines only: a flag *kernel* could
  // never be scheduled beside the persistent grid); obs are stored by the kernel straight into the caller's mapped buffer.
  void *obs_dev = nullptr, *tmp = nullptr;
  const bool direct = getenv("MJXB_HOST_DIRECT") ? atoi(getenv("MJXB_HOST_DIRECT")) != 0 : true;
  if (direct && host_dev_ptr(obs_host, &obs_dev) && host_dev_ptr(action_host, &tmp) && (!keys_host || host_dev_ptr(keys_host, &tmp))) {
    cudaStream_t main_st = a.pipe[0], copy_st = a.pipe[1];
    a.epoch++;
    const int csz = 1 << a.ready_shift, nchunk = (n_env + csz - 1) / csz;
    for (int k = 0; k < nchunk; k++) {
      const size_t o = (size_t)k * csz, c = (size_t)((n_env - k * csz < csz) ? n_env - k * csz : csz);
      a.ready_src[k] = a.epoch;
      CU(cudaMemcpyAsync(a.action + o * nu, action_host + o * nu, c * nu * 4, cudaMemcpyHostToDevice, copy_st));
      if (keys_host) CU(cudaMemcpyAsync(a.keys + o * 2, keys_host + o * 2, c * 8, cudaMemcpyHostToDevice, copy_st));
      CU(cudaMemcpyAsync(a.ready + k, a.ready_src + k, sizeof(unsigned), cudaMemcpyHostToDevice, copy_st));
    }
    StepArgs sa;
    memset(&sa, 0, sizeof(sa));
    sa.n_env = n_env; sa.mode = MODE_ENV_STEP; sa.nsteps = 1; sa.autoreset = keys_host ? 1 : 0; sa.in = arena_state(a); sa.out = sa.in;
    sa.action = a.action; sa.keys = keys_host ? a.keys : nullptr; sa.obs = static_cast<float*>(obs_dev); sa.reward = a.reward;
    sa.terminated = a.term; sa.truncated = a.trunc;
    sa.in_ready = a.ready; sa.in_ready_shift = a.ready_shift; sa.in_ready_epoch = a.epoch;
    sa.in_timeout = a.ready_src + Arena::kReadyMax;  // pinned: the same address is valid on the device (UVA)
    const bool obs_direct = getenv("MJXB_DIRECT_OBS") ? atoi(getenv("MJXB_DIRECT_OBS")) != 0 : true;
    const bool sc_direct = getenv("MJXB_DIRECT_SCALARS") ? atoi(getenv("MJXB_DIRECT_SCALARS")) != 0 : true;
    void *rd = nullptr, *td = nullptr, *ud = nullptr;
    const bool scd = sc_direct && host_dev_ptr(reward_host, &rd) && host_dev_ptr(terminated_host, &td) && host_dev_ptr(truncated_host, &ud);
    if (!obs_direct) sa.obs = a.obs;
    if (scd) { sa.reward = (float*)rd; sa.terminated = (float*)td; sa.truncated = (float*)ud; }
    int rc = launch(m, sa, false, main_st);
    if (rc) { cudaStreamSynchronize(copy_st); return rc; }
    if (!obs_direct) CU(cudaMemcpyAsync(obs_host, a.obs, (size_t)n_env * od * 4, cudaMemcpyDeviceToHost, main_st));
    if (!scd) {
    CU(cudaMemcpyAsync(reward_host, a.reward, (size_t)n_env * 4, cudaMemcpyDeviceToHost, main_st));
    CU(cudaMemcpyAsync(terminated_host, a.term, (size_t)n_env * 4, cudaMemcpyDeviceToHost, main_st));
    CU(cudaMemcpyAsync(truncated_host, a.trunc, (size_t)n_env * 4, cudaMemcpyDeviceToHost, main_st));
    }
    CU(cudaStreamSynchronize(copy_st));
    CU(cudaStreamSynchronize(main_st));
    if (a.ready_src[Arena::kReadyMax] != 0u) {
      a.ready_src[Arena::kReadyMax] = 0u;
      snprintf(g_cuda_err, sizeof(g_cuda_err), "mjxb_step_host: an input chunk did not reach the device within the bounded wait");
      return MJXB_ECUDA;
    }
    return MJXB_OK;
  }
  int slot = 0;
  for (int lo = 0; lo < n_env; lo += a.chunk, slot = (slot + 1) % Arena::kSlots) {
    const int cn = (n_env - lo < a.chunk) ? n_env - lo : a.chunk;
    const size_t o = (size_t)lo, c = (size_t)cn;
    cudaStream_t st = a.pipe[slot];
    CU(cudaMemcpyAsync(a.action + o * nu, action_host + o * nu, c * nu * 4, cudaMemcpyHostToDevice, st));
    if (keys_host) CU(cudaMemcpyAsync(a.keys + o * 2, keys_host + o * 2, c * 8, cudaMemcpyHostToDevice, st));
    mjxb_state sv;
    sv.qpos = a.qpos + o * nq; sv.qvel = a.qvel + o * nv; sv.qacc_warmstart = a.warm + o * nv; sv.time = a.time + o; sv.aux = a.aux + o * MJXB_AUX_DIM;
    StepArgs sa;
    memset(&sa, 0, sizeof(sa));
    sa.n_env = cn; sa.mode = MODE_ENV_STEP; sa.nsteps = 1; sa.autoreset = keys_host ? 1 : 0; sa.in = sv; sa.out = sv;
    sa.action = a.action + o * nu; sa.keys = keys_host ? a.keys + o * 2 : nullptr; sa.obs = a.obs + o * od; sa.reward = a.reward + o;
    sa.terminated = a.term + o; sa.truncated = a.trunc + o;
    int rc = launch(m, sa, false, st, a.pipe_ovf[slot], a.chunk);
    if (rc) return rc;
    CU(cudaMemcpyAsync(obs_host + o * od, a.obs + o * od, c * od * 4, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(reward_host + o, a.reward + o, c * 4, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(terminated_host + o, a.term + o, c * 4, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(truncated_host + o, a.trunc + o, c * 4, cudaMemcpyDeviceToHost, st));
  }
  for (int i = 0; i < Arena::kSlots; i++) CU(cudaStreamSynchronize(a.pipe[i]));
  return MJXB_OK;
}

int mjxb_step_host(mjxb_model* m, int32_t n_env, const float* action_host, float* obs_host, float* reward_host, float* terminated_host,
                   float* truncated_host) {
  return step_host_impl(m, n_env, action_host, nullptr, obs_host, reward_host, terminated_host, truncated_host);
}

int mjxb_step_autoreset_host(mjxb_model* m, int32_t n_env, const float* action_host, const uint32_t* keys_host, float* obs_host,
                             float* reward_host, float* terminated_host, float* truncated_host) {
  if (!keys_host) return MJXB_EINVAL;
  return step_host_impl(m, n_env, action_host, keys_host, obs_host, reward_host, terminated_host, truncated_host);
}

int mjxb_state_get_host(mjxb_model* m, int32_t n_env, float* qpos, float* qvel, float* qacc_warmstart, float* time, float* aux) {
  if (!m || n_env <= 0 || m->arena.n != n_env) return MJXB_EINVAL;
  Arena& a = m->arena;
  size_t N = (size_t)n_env;
  if (qpos) CU(cudaMemcpyAsync(qpos, a.qpos, N * m->host.nq * 4, cudaMemcpyDeviceToHost, a.stream));
  if (qvel) CU(cudaMemcpyAsync(qvel, a.qvel, N * m->host.nv * 4, cudaMemcpyDeviceToHost, a.stream));
  if (qacc_warmstart) CU(cudaMemcpyAsync(qacc_warmstart, a.warm, N * m->host.nv * 4, cudaMemcpyDeviceToHost, a.stream));
  if (time) CU(cudaMemcpyAsync(time, a.time, N * 4, cudaMemcpyDeviceToHost, a.stream));
  if (aux) CU(cudaMemcpyAsync(aux, a.aux, N * MJXB_AUX_DIM * 4, cudaMemcpyDeviceToHost, a.stream));
  CU(cudaStreamSynchronize(a.stream));
  return MJXB_OK;
}

int mjxb_state_set_host(mjxb_model* m, int32_t n_env, const float* qpos, const float* qvel, const float* qacc_warmstart,
                        const float* time, const float* aux) {
  if (!m || n_env <= 0) return MJXB_EINVAL;
  int rc = arena_ensure(m, n_env);
  if (rc) return rc;
  Arena& a = m->arena;
  size_t N = (size_t)n_env;
  if (qpos) CU(cudaMemcpyAsync(a.qpos, qpos, N * m->host.nq * 4, cudaMemcpyHostToDevice, a.stream));
  if (qvel) CU(cudaMemcpyAsync(a.qvel, qvel, N * m->host.nv * 4, cudaMemcpyHostToDevice, a.stream));
  if (qacc_warmstart) CU(cudaMemcpyAsync(a.warm, qacc_warmstart, N * m->host.nv * 4, cudaMemcpyHostToDevice, a.stream));
  if (time) CU(cudaMemcpyAsync(a.time, time, N * 4, cudaMemcpyHostToDevice, a.stream));
  if (aux) CU(cudaMemcpyAsync(a.aux, aux, N * MJXB_AUX_DIM * 4, cudaMemcpyHostToDevice, a.stream));
  CU(cudaStreamSynchronize(a.stream));
  return MJXB_OK;
}

}  // extern "C"
