// mjxb_learner.cu -- PPO learner helpers that replace chains of tiny elementwise / reduction launches in the minibatch step
// (reference train_ppo.py:204-252: gaussian_logprob, ratio, advantage normalisation, clipped surrogate, entropy bonus; optax.adam):
//   mjxb_ppo_loss : losses of one minibatch AND their gradients with respect to the policy mean / log_std in two launches
//   mjxb_adam     : Adam over ONE flat parameter / gradient buffer with a device-resident step counter (CUDA-graph replayable)
// Plain FP32 CUDA-core kernels: they are bandwidth-trivial (65536 x 21 floats); what they remove is ~45 launches per minibatch.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <string.h>

#include "mjxb.h"
#include "mjxb_internal.h"

namespace mjxbl {

constexpr int kMaxAct = 32;

// stats[0] = sum adv, stats[1] = sum adv^2  (zeroed by the caller's memset)
__global__ void adv_stats_kernel(int n, const float* __restrict__ adv, float* __restrict__ stats) {
  float s = 0.f, s2 = 0.f;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) { const float a = adv[i]; s += a; s2 += a * a; }
  for (int o = 16; o > 0; o >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, o); s2 += __shfl_xor_sync(0xffffffffu, s2, o); }
  __shared__ float sh[2][32];
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) { sh[0][w] = s; sh[1][w] = s2; }
  __syncthreads();
  if (w == 0) {
    s = l < (blockDim.x >> 5) ? sh[0][l] : 0.f; s2 = l < (blockDim.x >> 5) ? sh[1][l] : 0.f;
    for (int o = 16; o > 0; o >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, o); s2 += __shfl_xor_sync(0xffffffffu, s2, o); }
    if (l == 0) { atomicAdd(stats, s); atomicAdd(stats + 1, s2); }
  }
}

// one thread per sample: logp, ratio, clipped surrogate, d loss / d mean; block-reduced d loss / d log_std and loss sums
__global__ void ppo_loss_kernel(int n, int A, int ld, const float* __restrict__ mean, const float* __restrict__ log_std, const float* __restrict__ action,
                                const float* __restrict__ old_logp, const float* __restrict__ adv, const float* __restrict__ stats, float clip_eps,
                                float ent_coef, float* __restrict__ g_mean, float* __restrict__ g_log_std, float* __restrict__ loss_out) {
  __shared__ float s_ls[kMaxAct], s_ivar[kMaxAct], s_gls[kMaxAct];
  __shared__ float s_loss;
  if (threadIdx.x < kMaxAct) {
    const float ls = threadIdx.x < A ? log_std[threadIdx.x] : 0.f;
    s_ls[threadIdx.x] = ls; s_ivar[threadIdx.x] = expf(-2.0f * ls); s_gls[threadIdx.x] = 0.f;
  }
  if (threadIdx.x == 0) s_loss = 0.f;
  __syncthreads();
  const float inv_n = 1.0f / (float)n;
  const float amean = stats[0] * inv_n;
  const float avar = fmaxf(stats[1] * inv_n - amean * amean, 0.0f);          // population variance (torch .std(unbiased=False))
  const float ainv = 1.0f / (sqrtf(avar) + 1e-8f);
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  float gls[kMaxAct];
#pragma unroll
  for (int j = 0; j < kMaxAct; j++) gls[j] = 0.f;
  float lrow = 0.f;
  if (i < n) {
    float d[kMaxAct], q = 0.f;
#pragma unroll
    for (int j = 0; j < kMaxAct; j++) {
      d[j] = 0.f;
      if (j < A) {
        d[j] = action[(size_t)i * A + j] - mean[(size_t)i * ld + j];
        q += d[j] * d[j] * s_ivar[j] + 2.0f * s_ls[j] + 1.8378770664093453f;     // log(2 pi)
      }
    }
    const float logp = -0.5f * q;
    const float ratio = expf(logp - old_logp[i]);
    const float adn = (adv[i] - amean) * ainv;
    const float s1 = ratio * adn, cr = fminf(fmaxf(ratio, 1.0f - clip_eps), 1.0f + clip_eps), s2 = cr * adn;
    lrow = -fminf(s1, s2) * inv_n;
    // d(-min(s1, s2))/d ratio: s1 is taken when s1 <= s2 (ties share the same derivative inside the clip range)
    float dr;
    if (s1 < s2) dr = -adn;
    else if (s2 < s1) dr = (ratio > 1.0f - clip_eps && ratio < 1.0f + clip_eps) ? -adn : 0.0f;
    else dr = (ratio > 1.0f - clip_eps && ratio < 1.0f + clip_eps) ? -adn : -0.5f * adn;
    const float dlogp = dr * ratio * inv_n;
#pragma unroll
    for (int j = 0; j < kMaxAct; j++)
      if (j < A) {
        const float t = d[j] * s_ivar[j];
        g_mean[(size_t)i * ld + j] = dlogp * t;              // d logp / d mean_j = (a_j - mean_j) / var_j
        gls[j] = dlogp * (d[j] * t - 1.0f);                  // d logp / d log_std_j = (a_j - mean_j)^2 / var_j - 1
      } else if (j < ld) {
        g_mean[(size_t)i * ld + j] = 0.0f;                   // padding columns of a padded output layer carry no gradient
      }
  }
  // block reduction of the log_std gradient and the loss
#pragma unroll
  for (int j = 0; j < kMaxAct; j++) {
    if (j < A) {
      float v = gls[j];
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if ((threadIdx.x & 31) == 0) atomicAdd(&s_gls[j], v);
    }
  }
  for (int o = 16; o > 0; o >>= 1) lrow += __shfl_xor_sync(0xffffffffu, lrow, o);
  if ((threadIdx.x & 31) == 0) atomicAdd(&s_loss, lrow);
  __syncthreads();
  if (threadIdx.x < A) atomicAdd(g_log_std + threadIdx.x, s_gls[threadIdx.x] - (blockIdx.x == 0 ? ent_coef / (float)A : 0.0f));
  if (threadIdx.x == 0) {
    atomicAdd(loss_out, s_loss);
    if (blockIdx.x == 0) {   // entropy bonus: loss -= ent_coef * 0.5 * sum_j (1 + log 2 pi + 2 log_std_j) / A
      float ent = 0.f;
      for (int j = 0; j < A; j++) ent += 1.0f + 1.8378770664093453f + 2.0f * s_ls[j];
      atomicAdd(loss_out, -ent_coef * 0.5f * ent / (float)A);
    }
  }
}

// Adam (optax.adam / torch.optim.Adam semantics) over a flat buffer; elements [0, split) use lr0, the rest lr1; *step is the number
// of updates already applied and is advanced by the tick kernel that runs first
__global__ void adam_tick_kernel(float* step) { *step += 1.0f; }
__global__ void adam_kernel(int n, int split, float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                            const float* __restrict__ step, float lr0, float lr1, float b1, float b2, float eps, float gscale) {
  const float t = *step;
  const float c1 = (float)(1.0 - pow((double)b1, (double)t)), c2 = (float)(1.0 - pow((double)b2, (double)t));   // exact: fast-math powf is not
  const float isc2 = rsqrtf(c2);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const float gi = g[i] * gscale;
    const float mi = b1 * m[i] + (1.0f - b1) * gi, vi = b2 * v[i] + (1.0f - b2) * gi * gi;
    m[i] = mi; v[i] = vi;
    const float lr = i < split ? lr0 : lr1;
    p[i] -= (lr / c1) * mi / (sqrtf(vi) * isc2 + eps);
  }
}

}  // namespace mjxbl

extern "C" {

int mjxb_ppo_loss(int32_t n, int32_t act_dim, const float* mean, const float* log_std, const float* action, const float* old_logp,
                  const float* adv, float clip_eps, float ent_coef, float* scratch4, float* g_mean, float* g_log_std, float* loss_out,
                  void* stream_) {
  return mjxb_ppo_loss_ld(n, act_dim, act_dim, mean, log_std, action, old_logp, adv, clip_eps, ent_coef, scratch4, g_mean, g_log_std, loss_out,
                          stream_);
}

int mjxb_ppo_loss_ld(int32_t n, int32_t act_dim, int32_t ld_mean, const float* mean, const float* log_std, const float* action,
                     const float* old_logp, const float* adv, float clip_eps, float ent_coef, float* scratch4, float* g_mean, float* g_log_std,
                     float* loss_out, void* stream_) {
  if (n <= 0 || act_dim <= 0 || act_dim > mjxbl::kMaxAct || ld_mean < act_dim || ld_mean > mjxbl::kMaxAct || !mean || !log_std || !action || !old_logp || !adv || !scratch4 || !g_mean ||
      !g_log_std || !loss_out) return MJXB_EINVAL;
  cudaStream_t stream = (cudaStream_t)stream_;
  cudaError_t e = cudaMemsetAsync(scratch4, 0, 4 * sizeof(float), stream);
  if (e == cudaSuccess) e = cudaMemsetAsync(g_log_std, 0, act_dim * sizeof(float), stream);
  if (e == cudaSuccess) e = cudaMemsetAsync(loss_out, 0, sizeof(float), stream);
  if (e != cudaSuccess) return mjxb::report_cuda_error(e, "cudaMemsetAsync(ppo_loss)");
  int blocks = (n + 255) / 256;
  mjxbl::adv_stats_kernel<<<blocks > 592 ? 592 : blocks, 256, 0, stream>>>(n, adv, scratch4);
  mjxbl::ppo_loss_kernel<<<(n + 127) / 128, 128, 0, stream>>>(n, act_dim, ld_mean, mean, log_std, action, old_logp, adv, scratch4, clip_eps, ent_coef,
                                                             g_mean, g_log_std, loss_out);
  g_mjxb_launches += 2;
  e = cudaGetLastError();
  return e == cudaSuccess ? MJXB_OK : mjxb::report_cuda_error(e, "ppo_loss_kernel launch");
}

int mjxb_adam(int32_t n, int32_t split, float* param, const float* grad, float* m, float* v, float* step_dev, float lr0, float lr1, float b1,
              float b2, float eps, float grad_scale, void* stream_) {
  if (n <= 0 || split < 0 || split > n || !param || !grad || !m || !v || !step_dev) return MJXB_EINVAL;
  cudaStream_t stream = (cudaStream_t)stream_;
  mjxbl::adam_tick_kernel<<<1, 1, 0, stream>>>(step_dev);
  int blocks = (n + 255) / 256;
  mjxbl::adam_kernel<<<blocks > 1184 ? 1184 : blocks, 256, 0, stream>>>(n, split, param, grad, m, v, step_dev, lr0, lr1, b1, b2, eps, grad_scale);
  g_mjxb_launches += 2;
  const cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? MJXB_OK : mjxb::report_cuda_error(e, "adam_kernel launch");
}

}  // extern "C"

// ------------------------------------------------------------------------------------------------------------------------------
// Fused gradient all-reduce + Adam over NVLink peer memory (one process per GPU; the learner's only collective).
// Every rank owns a gradient buffer and a flag block allocated with cudaMalloc and exported through CUDA IPC; peers map them. One
// kernel per minibatch: (1) cross-GPU barrier: "my gradients are complete" (a release store of the epoch into every peer's flag
// block, then an acquire spin on the own block); (2) every rank reads ALL ranks' gradients element-wise straight from peer memory
// (P2P loads over NVLink / NVSwitch), sums them in rank order -- identical bits on every rank, so the replicas stay in step -- and
// applies Adam to its own copy of the parameters; (3) cross-GPU barrier: "I have finished reading" (so that no rank starts
// overwriting its gradients while a peer still reads them). No NCCL launch, no flatten / unflatten, no separate optimiser pass; the
// epoch lives on the device, so the kernel replays from a CUDA graph. All spins are bounded (a lost peer sets *error, never hangs).
namespace mjxbl {

constexpr int kMaxWorld = 16;

struct CommDev {                    // device-visible part (passed by value)
  int rank, world;
  const float* grad[kMaxWorld];     // every rank's gradient buffer (own included), mapped into this process
  unsigned* flags[kMaxWorld];       // every rank's flag block: [0, kMaxWorld) barrier-1 slots, [kMaxWorld, 2 kMaxWorld) barrier-2 slots
  unsigned* epoch;                  // own device counter (number of completed all-reduces)
  unsigned* done_blocks;            // own device counter
  int* error;                       // own device flag
};

__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) { asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ bool wait_all(const unsigned* local_flags, int world, unsigned e) {
  for (int r = 0; r < world; r++) {
    bool ok = false;
    for (int spin = 0; spin < (1 << 23); spin++) {               // bounded (~ seconds): a lost peer must not hang the device
      if ((int)(ld_acquire_sys(local_flags + r) - e) >= 0) { ok = true; break; }
      __nanosleep(64);
    }
    if (!ok) return false;
  }
  return true;
}

__global__ void __launch_bounds__(256) allreduce_adam_kernel(CommDev cm, int n, int split, float* __restrict__ p, float* __restrict__ m,
                                                             float* __restrict__ v, float* __restrict__ step, float lr0, float lr1, float b1,
                                                             float b2, float eps) {
  __shared__ int s_ok;
  const unsigned e = *reinterpret_cast<volatile unsigned*>(cm.epoch) + 1u;
  // ---- barrier 1: every rank's backward pass is complete
  if (threadIdx.x == 0) {
    if (blockIdx.x == 0) {
      __threadfence_system();
      for (int r = 0; r < cm.world; r++) st_release_sys(cm.flags[r] + cm.rank, e);
    }
    s_ok = wait_all(cm.flags[cm.rank], cm.world, e) ? 1 : 0;
    if (!s_ok) *cm.error = 1;
  }
  __syncthreads();
  if (s_ok) {
    const float t = *step + 1.0f;                                 // advanced by the last block below
    const float c1 = (float)(1.0 - pow((double)b1, (double)t)), c2 = (float)(1.0 - pow((double)b2, (double)t));
    const float isc2 = rsqrtf(c2), inv_w = 1.0f / (float)cm.world;
    const int n4 = n >> 2;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += gridDim.x * blockDim.x) {
      float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int r = 0; r < cm.world; r++) {                        // fixed rank order: the same bits on every rank
        const float4 x = __ldcv(reinterpret_cast<const float4*>(cm.grad[r]) + i);
        g.x += x.x; g.y += x.y; g.z += x.z; g.w += x.w;
      }
      float gi[4] = {g.x * inv_w, g.y * inv_w, g.z * inv_w, g.w * inv_w};
      float4 pm = reinterpret_cast<float4*>(m)[i], pv = reinterpret_cast<float4*>(v)[i], pp = reinterpret_cast<float4*>(p)[i];
      float mm[4] = {pm.x, pm.y, pm.z, pm.w}, vv[4] = {pv.x, pv.y, pv.z, pv.w}, ppp[4] = {pp.x, pp.y, pp.z, pp.w};
#pragma unroll
      for (int k = 0; k < 4; k++) {
        mm[k] = b1 * mm[k] + (1.0f - b1) * gi[k];
        vv[k] = b2 * vv[k] + (1.0f - b2) * gi[k] * gi[k];
        const float lr = (4 * i + k) < split ? lr0 : lr1;
        ppp[k] -= (lr / c1) * mm[k] / (sqrtf(vv[k]) * isc2 + eps);
      }
      reinterpret_cast<float4*>(m)[i] = make_float4(mm[0], mm[1], mm[2], mm[3]);
      reinterpret_cast<float4*>(v)[i] = make_float4(vv[0], vv[1], vv[2], vv[3]);
      reinterpret_cast<float4*>(p)[i] = make_float4(ppp[0], ppp[1], ppp[2], ppp[3]);
    }
    for (int i = 4 * n4 + blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {   // tail
      float g = 0.f;
      for (int r = 0; r < cm.world; r++) g += __ldcv(cm.grad[r] + i);
      g *= inv_w;
      const float mi = b1 * m[i] + (1.0f - b1) * g, vi = b2 * v[i] + (1.0f - b2) * g * g;
      m[i] = mi; v[i] = vi;
      p[i] -= ((i < split ? lr0 : lr1) / c1) * mi / (sqrtf(vi) * isc2 + eps);
    }
  }
  // ---- barrier 2: the last block of this rank tells the peers it has finished reading, and waits for theirs
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned done = atomicAdd(cm.done_blocks, 1u);
    if (done == gridDim.x - 1) {
      for (int r = 0; r < cm.world; r++) st_release_sys(cm.flags[r] + kMaxWorld + cm.rank, e);
      if (!wait_all(cm.flags[cm.rank] + kMaxWorld, cm.world, e)) *cm.error = 1;
      *step += 1.0f;
      *cm.done_blocks = 0u;
      *reinterpret_cast<volatile unsigned*>(cm.epoch) = e;
      __threadfence();
    }
  }
}

}  // namespace mjxbl

struct mjxb_comm {
  int rank = 0, world = 1, n = 0, device = 0;
  float* grad = nullptr;            // own gradient buffer (cudaMalloc, IPC-exported)
  unsigned* flags = nullptr;        // own flag block
  unsigned* counters = nullptr;     // [0] epoch, [1] done blocks, [2] error (own, not shared)
  void* peer_grad[mjxbl::kMaxWorld] = {};
  void* peer_flags[mjxbl::kMaxWorld] = {};
  bool connected = false;
  mjxbl::CommDev dev;
};

extern "C" {

int mjxb_comm_create(int32_t rank, int32_t world, int32_t n_floats, mjxb_comm** out) {
  if (!out || world < 1 || world > mjxbl::kMaxWorld || rank < 0 || rank >= world || n_floats <= 0) return MJXB_EINVAL;
  mjxb_comm* c = new mjxb_comm();
  c->rank = rank; c->world = world; c->n = n_floats;
  cudaError_t e = cudaGetDevice(&c->device);
  if (e == cudaSuccess) e = cudaMalloc(&c->grad, (size_t)((n_floats + 3) & ~3) * sizeof(float));
  if (e == cudaSuccess) e = cudaMalloc(&c->flags, 2 * mjxbl::kMaxWorld * sizeof(unsigned));
  if (e == cudaSuccess) e = cudaMalloc(&c->counters, 4 * sizeof(unsigned));
  if (e == cudaSuccess) e = cudaMemset(c->grad, 0, (size_t)((n_floats + 3) & ~3) * sizeof(float));
  if (e == cudaSuccess) e = cudaMemset(c->flags, 0, 2 * mjxbl::kMaxWorld * sizeof(unsigned));
  if (e == cudaSuccess) e = cudaMemset(c->counters, 0, 4 * sizeof(unsigned));
  if (e != cudaSuccess) { mjxb_comm_destroy(c); return mjxb::report_cuda_error(e, "mjxb_comm_create"); }
  *out = c;
  return MJXB_OK;
}

/* 2 x 64 bytes: the IPC handles of the gradient buffer and of the flag block */
int mjxb_comm_local_handles(mjxb_comm* c, void* handles_out) {
  if (!c || !handles_out) return MJXB_EINVAL;
  cudaIpcMemHandle_t h[2];
  cudaError_t e = cudaIpcGetMemHandle(&h[0], c->grad);
  if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h[1], c->flags);
  if (e != cudaSuccess) return mjxb::report_cuda_error(e, "cudaIpcGetMemHandle");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  memcpy(handles_out, h, sizeof(h));
  return MJXB_OK;
}

/* all_handles: world x 128 bytes in rank order (every rank's mjxb_comm_local_handles output) */
int mjxb_comm_connect(mjxb_comm* c, const void* all_handles) {
  if (!c || !all_handles) return MJXB_EINVAL;
  const cudaIpcMemHandle_t* h = reinterpret_cast<const cudaIpcMemHandle_t*>(all_handles);
  for (int r = 0; r < c->world; r++) {
    if (r == c->rank) { c->peer_grad[r] = c->grad; c->peer_flags[r] = c->flags; continue; }
    cudaError_t e = cudaIpcOpenMemHandle(&c->peer_grad[r], h[2 * r], cudaIpcMemLazyEnablePeerAccess);
    if (e == cudaSuccess) e = cudaIpcOpenMemHandle(&c->peer_flags[r], h[2 * r + 1], cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) return mjxb::report_cuda_error(e, "cudaIpcOpenMemHandle");
  }
  c->dev.rank = c->rank; c->dev.world = c->world;
  for (int r = 0; r < c->world; r++) { c->dev.grad[r] = (const float*)c->peer_grad[r]; c->dev.flags[r] = (unsigned*)c->peer_flags[r]; }
  c->dev.epoch = c->counters; c->dev.done_blocks = c->counters + 1; c->dev.error = (int*)(c->counters + 2);
  c->connected = true;
  return MJXB_OK;
}

float* mjxb_comm_grad_buffer(mjxb_comm* c) { return c ? c->grad : nullptr; }

/* 0 = healthy; 1 = a peer did not reach a barrier within the bounded wait (synchronises the stream's device) */
int mjxb_comm_error(mjxb_comm* c) {
  if (!c) return MJXB_EINVAL;
  int err = 0;
  if (cudaMemcpy(&err, c->counters + 2, sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess) { cudaGetLastError(); return MJXB_ECUDA; }
  return err;
}

int mjxb_allreduce_adam(mjxb_comm* c, int32_t n, int32_t split, float* param, float* m, float* v, float* step_dev, float lr0, float lr1,
                        float b1, float b2, float eps, void* stream) {
  if (!c || !c->connected || n <= 0 || n > c->n || split < 0 || split > n || !param || !m || !v || !step_dev) return MJXB_EINVAL;
  // one wave of co-resident blocks (the cross-GPU barriers are executed by every block): 64 blocks x 256 threads
  mjxbl::allreduce_adam_kernel<<<64, 256, 0, (cudaStream_t)stream>>>(c->dev, n, split, param, m, v, step_dev, lr0, lr1, b1, b2, eps);
  g_mjxb_launches += 1;
  const cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? MJXB_OK : mjxb::report_cuda_error(e, "allreduce_adam_kernel launch");
}

void mjxb_comm_destroy(mjxb_comm* c) {
  if (!c) return;
  for (int r = 0; r < c->world; r++) {
    if (r == c->rank) continue;
    if (c->peer_grad[r]) cudaIpcCloseMemHandle(c->peer_grad[r]);
    if (c->peer_flags[r]) cudaIpcCloseMemHandle(c->peer_flags[r]);
  }
  if (c->grad) cudaFree(c->grad);
  if (c->flags) cudaFree(c->flags);
  if (c->counters) cudaFree(c->counters);
  delete c;
}

}  // extern "C"
