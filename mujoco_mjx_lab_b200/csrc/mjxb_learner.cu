// mjxb_learner.cu -- PPO learner helpers that replace chains of tiny elementwise / reduction launches in the minibatch step
// (reference train_ppo.py:204-252: gaussian_logprob, ratio, advantage normalisation, clipped surrogate, entropy bonus; optax.adam):
//   mjxb_ppo_loss : losses of one minibatch AND their gradients with respect to the policy mean / log_std in two launches
//   mjxb_adam     : Adam over ONE flat parameter / gradient buffer with a device-resident step counter (CUDA-graph replayable)
// Plain FP32 CUDA-core kernels: they are bandwidth-trivial (65536 x 21 floats); what they remove is ~45 launches per minibatch.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

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
__global__ void ppo_loss_kernel(int n, int A, const float* __restrict__ mean, const float* __restrict__ log_std, const float* __restrict__ action,
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
        d[j] = action[(size_t)i * A + j] - mean[(size_t)i * A + j];
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
        g_mean[(size_t)i * A + j] = dlogp * t;               // d logp / d mean_j = (a_j - mean_j) / var_j
        gls[j] = dlogp * (d[j] * t - 1.0f);                  // d logp / d log_std_j = (a_j - mean_j)^2 / var_j - 1
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
  if (n <= 0 || act_dim <= 0 || act_dim > mjxbl::kMaxAct || !mean || !log_std || !action || !old_logp || !adv || !scratch4 || !g_mean ||
      !g_log_std || !loss_out) return MJXB_EINVAL;
  cudaStream_t stream = (cudaStream_t)stream_;
  cudaError_t e = cudaMemsetAsync(scratch4, 0, 4 * sizeof(float), stream);
  if (e == cudaSuccess) e = cudaMemsetAsync(g_log_std, 0, act_dim * sizeof(float), stream);
  if (e == cudaSuccess) e = cudaMemsetAsync(loss_out, 0, sizeof(float), stream);
  if (e != cudaSuccess) return mjxb::report_cuda_error(e, "cudaMemsetAsync(ppo_loss)");
  int blocks = (n + 255) / 256;
  mjxbl::adv_stats_kernel<<<blocks > 592 ? 592 : blocks, 256, 0, stream>>>(n, adv, scratch4);
  mjxbl::ppo_loss_kernel<<<(n + 127) / 128, 128, 0, stream>>>(n, act_dim, mean, log_std, action, old_logp, adv, scratch4, clip_eps, ent_coef,
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
