// mjxb_adjoint.cu -- GPU build of the reverse-mode env step (mjxb_adjoint.cuh) and its C-ABI entry points
// (mjxb_step_fwd_tape / mjxb_step_vjp, include/mjxb.h).  One warp per env; the per-env workspace lives in shared memory.
#include <cuda_runtime.h>

#include <mutex>

#include "mjxb.h"
#include "mjxb_adjoint.cuh"
#include "mjxb_internal.h"

namespace mjxb {

constexpr int VJP_CAP = 48, VJP_MAXCC = 20;   // main tile: rows / contacts of one env; envs that need more go to the CAP_BIG tile

struct VjpArgs {
  int n_env;
  mjxb_state in;
  const float *action, *tape_qacc, *g_qpos_out, *g_qvel_out, *g_aux_out, *g_reward;
  float *g_qpos_in, *g_qvel_in, *g_aux_in, *g_action;
  int32_t* status;
  int *in_count, *in_list, *in_done, *out_count, *out_list;
};

template <int CAP, int MAXCC>
__global__ void __launch_bounds__(256, 1) step_vjp_kernel(const DevModel* __restrict__ gmodel, const PairParam* __restrict__ pair_param, VjpArgs A) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  pdl_prologue();
  if (A.in_list != nullptr && *reinterpret_cast<volatile int*>(A.in_count) == 0) return;   // empty overflow list: leave at once
  DevModel& C = *reinterpret_cast<DevModel*>(smem_raw);
  {
    const int4* src = reinterpret_cast<const int4*>(gmodel);
    int4* dst = reinterpret_cast<int4*>(smem_raw);
    for (int i = threadIdx.x; i < (int)(sizeof(DevModel) / 16); i += blockDim.x) dst[i] = src[i];
  }
  __syncthreads();
  using WS = adj::AdjS<float, CAP, MAXCC>;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  WS& W = *reinterpret_cast<WS*>(smem_raw + ((sizeof(DevModel) + 15) & ~size_t(15)) + (size_t)warp * ((sizeof(WS) + 15) & ~size_t(15)));
  const adj::Lanes X{lane, 32};
  const bool consuming = A.in_list != nullptr;
  const int n_items = consuming ? *reinterpret_cast<volatile int*>(A.in_count) : A.n_env;
  const int nq = C.nq, nu = C.nu;
  for (int item = blockIdx.x * nwarp + warp; item < n_items; item += gridDim.x * nwarp) {
    const size_t env = consuming ? A.in_list[item] : item;
    adj::EnvIO<float> io;
    io.qpos = A.in.qpos + env * nq; io.qvel = A.in.qvel + env * NV; io.aux = A.in.aux ? A.in.aux + env * MJXB_AUX_DIM : nullptr;
    io.action = A.action + env * nu; io.tape_qacc = A.tape_qacc + env * NV;
    io.g_qpos_out = A.g_qpos_out ? A.g_qpos_out + env * nq : nullptr; io.g_qvel_out = A.g_qvel_out ? A.g_qvel_out + env * NV : nullptr;
    io.g_aux_out = A.g_aux_out ? A.g_aux_out + env * MJXB_AUX_DIM : nullptr; io.g_reward = A.g_reward ? A.g_reward[env] : 0.0f;
    io.g_qpos_in = A.g_qpos_in + env * nq; io.g_qvel_in = A.g_qvel_in + env * NV;
    io.g_aux_in = A.g_aux_in ? A.g_aux_in + env * MJXB_AUX_DIM : nullptr; io.g_action = A.g_action + env * nu;
    const int st = adj::step_vjp_env<float>(C, pair_param, W, X, io);
    if (lane == 0) {
      if (st == adj::VJP_OVERFLOW && A.out_list != nullptr) { const int slot = atomicAdd(A.out_count, 1); A.out_list[slot] = (int)env; }
      if (A.status) A.status[env] = (consuming ? MJXB_STATUS_ROW_SPILL : 0) | (st == adj::VJP_NONFINITE ? MJXB_STATUS_NAN : 0) |
                                    (st == adj::VJP_OVERFLOW && A.out_list == nullptr ? MJXB_STATUS_ROW_SPILL | MJXB_STATUS_NAN : 0);
    }
    __syncwarp();
  }
  if (consuming) {  // last CTA out resets the consumed list's counters
    __syncthreads();
    if (threadIdx.x == 0) {
      __threadfence();
      const int t = atomicAdd(A.in_done, 1);
      if (t == (int)gridDim.x - 1) { *A.in_count = 0; *A.in_done = 0; __threadfence(); }
    }
  }
}

namespace {
struct VjpGeom { int warps, warps_big; size_t smem, smem_big; bool ok; };
VjpGeom g_geom[64];
bool g_geom_set[64] = {};
std::mutex g_geom_mu;

int vjp_geometry(int device, VjpGeom* out) {
  std::lock_guard<std::mutex> lock(g_geom_mu);
  if (device < 0 || device >= 64) return MJXB_EINVAL;
  if (!g_geom_set[device]) {
    cudaDeviceProp prop;
    cudaError_t e = cudaGetDeviceProperties(&prop, device);
    if (e != cudaSuccess) return report_cuda_error(e, "cudaGetDeviceProperties");
    using WSMain = adj::AdjS<float, VJP_CAP, VJP_MAXCC>;
    using WSBig = adj::AdjS<float, CAP_BIG, MAXCC_BIG>;
    const size_t model_bytes = (sizeof(DevModel) + 15) & ~size_t(15), avail = prop.sharedMemPerBlockOptin;
    const size_t wsm = (sizeof(WSMain) + 15) & ~size_t(15), wsb = (sizeof(WSBig) + 15) & ~size_t(15);
    VjpGeom g;
    g.warps = (int)((avail - model_bytes) / wsm); g.warps_big = (int)((avail - model_bytes) / wsb);
    if (g.warps > 8) g.warps = 8;
    if (g.warps_big > 8) g.warps_big = 8;
    g.ok = g.warps >= 1 && g.warps_big >= 1;
    g.smem = model_bytes + g.warps * wsm; g.smem_big = model_bytes + g.warps_big * wsb;
    if (g.ok) {
      e = cudaFuncSetAttribute(step_vjp_kernel<VJP_CAP, VJP_MAXCC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.smem);
      if (e == cudaSuccess) e = cudaFuncSetAttribute(step_vjp_kernel<CAP_BIG, MAXCC_BIG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.smem_big);
      if (e != cudaSuccess) return report_cuda_error(e, "cudaFuncSetAttribute(step_vjp_kernel)");
    }
    g_geom[device] = g;
    g_geom_set[device] = true;
  }
  *out = g_geom[device];
  return out->ok ? MJXB_OK : MJXB_EUNSUPPORTED;
}
}  // namespace

}  // namespace mjxb

using namespace mjxb;

extern "C" {

int mjxb_step_fwd_tape(const mjxb_model* m, int32_t n_env, mjxb_state in, const float* action, mjxb_state out, float* obs, float* reward,
                       float* terminated, float* truncated, float* tape_qacc, int32_t* status, void* stream) {
  if (!tape_qacc) return MJXB_EINVAL;
  const int rc = mjxb_step(m, n_env, in, action, out, obs, reward, terminated, truncated, status, stream);
  if (rc != MJXB_OK) return rc;
  ModelView mv;
  if (model_view(m, &mv) != MJXB_OK) return MJXB_EINVAL;
  // the tape of the reverse pass is the solver's qacc, which the step stores as the next warm start
  if (tape_qacc != out.qacc_warmstart) {
    const cudaError_t e = cudaMemcpyAsync(tape_qacc, out.qacc_warmstart, (size_t)n_env * mv.host->nv * sizeof(float), cudaMemcpyDeviceToDevice,
                                          (cudaStream_t)stream);
    if (e != cudaSuccess) return report_cuda_error(e, "cudaMemcpyAsync(tape)");
  }
  return MJXB_OK;
}

int mjxb_step_vjp(const mjxb_model* m, int32_t n_env, mjxb_state in, const float* action, const float* tape_qacc, const float* g_qpos_out,
                  const float* g_qvel_out, const float* g_aux_out, const float* g_reward, float* g_qpos_in, float* g_qvel_in, float* g_aux_in,
                  float* g_action, int32_t* status, void* stream_) {
  if (!m || n_env <= 0 || !in.qpos || !in.qvel || !action || !tape_qacc || !g_qpos_in || !g_qvel_in || !g_action) return MJXB_EINVAL;
  if (in.aux != nullptr && g_aux_in == nullptr) return MJXB_EINVAL;
  cudaStream_t stream = (cudaStream_t)stream_;
  ModelView mv;
  if (model_view(m, &mv) != MJXB_OK) return MJXB_EINVAL;
  int cur = 0;
  cudaError_t e = cudaGetDevice(&cur);
  if (e != cudaSuccess) return report_cuda_error(e, "cudaGetDevice");
  if (cur != mv.device && (e = cudaSetDevice(mv.device)) != cudaSuccess) return report_cuda_error(e, "cudaSetDevice");
  VjpGeom g;
  int rc = vjp_geometry(mv.device, &g);
  int *buf = nullptr, cap = 0;
  if (rc == MJXB_OK) rc = model_scratch(m, stream, n_env, &buf, &cap);
  if (rc == MJXB_OK) {
    VjpArgs A;
    A.n_env = n_env; A.in = in; A.action = action; A.tape_qacc = tape_qacc; A.g_qpos_out = g_qpos_out; A.g_qvel_out = g_qvel_out;
    A.g_aux_out = g_aux_out; A.g_reward = g_reward; A.g_qpos_in = g_qpos_in; A.g_qvel_in = g_qvel_in; A.g_aux_in = g_aux_in;
    A.g_action = g_action; A.status = status;
    A.in_count = nullptr; A.in_list = nullptr; A.in_done = nullptr; A.out_count = buf; A.out_list = buf + 4;
    int grid = (n_env + g.warps - 1) / g.warps;
    if (grid > mv.num_sms) grid = mv.num_sms;
    launch_pdl(step_vjp_kernel<VJP_CAP, VJP_MAXCC>, dim3(grid), dim3(g.warps * 32), g.smem, stream, mv.dev, mv.dev_pp, A);
    A.in_count = buf; A.in_done = buf + 1; A.in_list = buf + 4; A.out_count = nullptr; A.out_list = nullptr;
    int gridb = (n_env + g.warps_big - 1) / g.warps_big;
    if (gridb > mv.num_sms) gridb = mv.num_sms;
    launch_pdl(step_vjp_kernel<CAP_BIG, MAXCC_BIG>, dim3(gridb), dim3(g.warps_big * 32), g.smem_big, stream, mv.dev, mv.dev_pp, A);
    g_mjxb_launches += 2;
    e = cudaGetLastError();
    if (e != cudaSuccess) rc = report_cuda_error(e, "step_vjp_kernel launch");
  }
  if (cur != mv.device) cudaSetDevice(cur);
  return rc;
}

}  // extern "C"
