// mjxb_policy.cu -- fused policy inference for the rollout loop (reference train_ppo.py:135-140,
// src/networks.py:55-61,105-112): obs normalisation -> 54-256-256-256-21 tanh MLP -> Gaussian sample -> log-prob, ONE launch.
//
// sm_100a design (round 2: persistent, warp-specialised, two env tiles in flight per SM):
//   * one CTA per SM, 10 warps: warp 0 = weight producer, warp 1 = MMA issuer, warps 2-5 / 6-9 = the epilogue groups of tile slot 0 / 1;
//   * a tile is 128 envs (MMA M); each slot owns a 64 KB activation buffer in the canonical K-major layout and 256 TMEM columns;
//   * the four GEMMs of a tile run on the 5th-generation tensor cores (tcgen05.mma kind::f16, bf16 operands, fp32 accumulators in TMEM),
//     issued by ONE thread; the two slots are a layer apart, so the MMAs of one tile run while the other tile's epilogue group reads its
//     accumulator (tcgen05.ld), adds the bias, applies tanh and writes the next layer's A operand -- activations never leave the SM;
//   * weights are pre-packed (mjxb_policy_pack_weight) into 16 KB chunks, each a ready-to-use canonical K-major B operand
//     (hidden layers: 32 of K x 256 of N; output layer: 256 x 32), and stream through a 4-stage shared-memory ring with one bulk-copy
//     (cp.async.bulk -> mbarrier complete_tx) per chunk: the producer thread runs ahead of the MMAs by up to 64 KB;
//   * a tile's observations, noise and actions are contiguous blocks of global memory: each moves as ONE bulk copy through staging;
//   * pipelines are mbarriers only (ring full / empty, accumulator full, A-operand ready); no CTA-wide barrier after the prologue.
//
// Shared-memory operand layouts (no swizzle; core matrix = 8 rows x 16 bytes, stored contiguously):
//   A (activations, 128 rows x K):   byte(r,k) = (r%8)*16 + (r/8)*128 + (k/8)*2048 + (k%8)*2         -> SBO = 128, LBO = 2048
//   B chunk (W^T, N rows x Kc of K):  byte(n,k) = (n%8)*16 + (n/8)*(Kc/8)*128 + (k/8)*128 + (k%8)*2   -> SBO = 16*Kc, LBO = 128
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <atomic>
#include <mutex>
#include <stdint.h>
#include <stdio.h>

#include "mjxb.h"
#include "mjxb_internal.h"

namespace mjxbp {

constexpr int kTile = 128;      // envs per tile = MMA M
constexpr int kHid = 256;       // hidden width = MMA N of the hidden layers
constexpr int kInPad = 64;      // obs_dim 54 padded to a multiple of 16
constexpr int kOutPad = 32;     // action dim 21 padded to a multiple of 16
constexpr int kSlots = 2;       // env tiles in flight per CTA
constexpr int kEpiWarps = 4;    // per slot: one warp per TMEM lane quadrant
constexpr int kThreads = 32 * (2 + kSlots * kEpiWarps);   // producer + MMA issuer + 2 x 4 epilogue warps = 320
constexpr int kABytes = kTile * kHid * 2;                 // 64 KB per slot
constexpr int kChunkBytes = 16384;                        // one weight chunk = one ring stage
constexpr int kStages = 4;
constexpr int kBiasFloats = 3 * kHid + kOutPad;
constexpr int kOffRing = kSlots * kABytes;
constexpr int kOffBias = kOffRing + kStages * kChunkBytes;
constexpr int kOffNorm = kOffBias + kBiasFloats * 4;      // observation statistics: mean[64], 1 / sqrt(var + 1e-8)[64]
constexpr int kOffLogStd = kOffNorm + 2 * kInPad * 4;     // log_std[32], exp(log_std)[32], exp(-2 log_std)[32]
constexpr int kEpsCols = 24;                              // noise staging of one slot holds 128 x act_dim floats for act_dim <= 24
constexpr int kEpsBytes = kTile * kEpsCols * 4;           // (wider action vectors read their noise thread-per-row)
constexpr int kOffEps = kOffLogStd + 3 * kOutPad * 4;
constexpr int kOffBar = kOffEps + kSlots * kEpsBytes;
constexpr int kNumBar = 2 * kStages + 4 * kSlots;         // ring full / empty, accumulator full, A ready, observations / noise landed
constexpr int kOffMisc = kOffBar + kNumBar * 8;           // TMEM base slot, abort flag
constexpr int kSmemBytes = kOffMisc + 16;
constexpr uint32_t kTmemCols = 512;                       // 256 fp32 accumulator columns per slot
static_assert(kSmemBytes <= 232448, "exceeds the 227 KB opt-in shared memory of sm_100");
static_assert(kHid * 32 * 2 == kChunkBytes && kOutPad * kHid * 2 == kChunkBytes, "chunk geometry");

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFFu);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version 1 (sm_100); base offset 0, no swizzle
  return d;
}

// kind::f16 instruction descriptor: D = f32, A = B = bf16, both K-major, M = 128
__device__ __forceinline__ uint32_t umma_idesc(int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(kTile >> 4) << 24);
}

__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {   // arrives on `bar` when every MMA issued so far by this thread has retired
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(bar) : "memory");
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
        "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
        "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
        "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
}

__device__ __forceinline__ float tanh_fast(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// bounded mbarrier wait: a completion that never comes (a fault, a lost copy) sets the abort flag, every role leaves its loop, and the
// launch reports through *error instead of hanging the device
__device__ __forceinline__ bool mbar_wait(uint32_t bar, uint32_t parity, volatile int* abort_flag, int* error) {
  for (int spin = 0; spin < (1 << 22); spin++) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    if (ok) return true;
    if ((spin & 255) == 255 && *abort_flag) return false;
  }
  *abort_flag = 1;
  if (error) *error = 1;
  return false;
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(bar) : "memory");
}

#ifndef MJXB_POLICY_TRACE
#define MJXB_POLICY_TRACE 0
#endif
#if MJXB_POLICY_TRACE
__device__ unsigned long long g_trace[4096];
// CTA 0 only; three writer threads (MMA issuer = region 0, thread 0 of the epilogue group of slot s = region 1 + s), each with its own
// cursor (a register): one clock64 read and one fire-and-forget store per stamp
#define TRACE_DECL int trace_n = 0
__device__ __forceinline__ void trace(int region, int& n, int slot, unsigned long long tag) {
  if (blockIdx.x != 0 || n >= 640) return;
  g_trace[region * 1280 + 2 * n] = tag | ((unsigned long long)slot << 32);
  g_trace[region * 1280 + 2 * n + 1] = (unsigned long long)clock64();
  n++;
}
#define TRACE(slot, tag) trace(trace_region, trace_n, slot, tag)
#else
#define TRACE_DECL
#define TRACE(slot, tag)
#endif

struct PolicyArgs {
  int n_env, obs_dim, act_dim;
  const float* obs;        // [n, obs_dim]
  const float* rms_mean;   // [obs_dim] or NULL (no normalisation)
  const float* rms_var;    // [obs_dim]
  const void* w[4];        // packed bf16 weights (mjxb_policy_pack_weight)
  const float* b[4];       // biases, float32
  const float* log_std;    // [act_dim]
  const float* eps;        // [n, act_dim] standard normal noise
  float* act;              // [n, act_dim]
  float* logp;             // [n]
  float* mean;             // [n, act_dim] or NULL
  int* error;              // device flag: set to 1 if a pipeline completion was not observed within the bounded wait
};

// weight chunks of layer l: count, K extent of a chunk, N
__device__ __forceinline__ int layer_chunks(int l) { return l == 0 ? kInPad / 32 : (l == 3 ? 1 : kHid / 32); }
__device__ __forceinline__ int layer_kc(int l) { return l == 3 ? kHid : 32; }
__device__ __forceinline__ int layer_n(int l) { return l == 3 ? kOutPad : kHid; }

__global__ void __launch_bounds__(kThreads, 1) policy_act_kernel(PolicyArgs P) {
  extern __shared__ __align__(1024) unsigned char smem[];
  // programmatic dependent launch: staged behind the previous env step's kernels; TMEM allocation and barrier init run under their tail
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  float* sBias = reinterpret_cast<float*>(smem + kOffBias);
  float* sNorm = reinterpret_cast<float*>(smem + kOffNorm);
  float* sLogStd = reinterpret_cast<float*>(smem + kOffLogStd);
  const uint32_t bar0 = smem_u32(smem + kOffBar);
  auto bar_wfull = [&](int i) { return bar0 + 8u * (uint32_t)i; };
  auto bar_wempty = [&](int i) { return bar0 + 8u * (uint32_t)(kStages + i); };
  auto bar_accfull = [&](int s) { return bar0 + 8u * (uint32_t)(2 * kStages + s); };
  auto bar_aready = [&](int s) { return bar0 + 8u * (uint32_t)(2 * kStages + kSlots + s); };
  auto bar_obsfull = [&](int s) { return bar0 + 8u * (uint32_t)(2 * kStages + 2 * kSlots + s); };
  auto bar_epsfull = [&](int s) { return bar0 + 8u * (uint32_t)(2 * kStages + 3 * kSlots + s); };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + kOffMisc);
  volatile int* abort_flag = reinterpret_cast<volatile int*>(smem + kOffMisc + 4);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(tmem_slot)), "r"(kTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  if (tid == 32) {
    for (int i = 0; i < kStages; i++) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(bar_wfull(i)) : "memory");
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(bar_wempty(i)) : "memory");
    }
    for (int s = 0; s < kSlots; s++) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(bar_accfull(s)) : "memory");
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar_aready(s)), "r"(kEpiWarps * 32) : "memory");
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(bar_obsfull(s)) : "memory");
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(bar_epsfull(s)) : "memory");
    }
    *abort_flag = 0;
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  asm volatile("griddepcontrol.wait;" ::: "memory");   // from here on the predecessor's results (observations, parameters) are read
  for (int i = tid; i < kBiasFloats; i += kThreads) {
    const int l = i < 3 * kHid ? i / kHid : 3, j = i - l * kHid;
    sBias[i] = (l < 3 || j < P.act_dim) ? P.b[l][j] : 0.0f;
  }
  if (tid >= 64 && tid < 64 + kOutPad) {
    const int j = tid - 64;
    const float ls = j < P.act_dim ? P.log_std[j] : 0.0f;
    sLogStd[j] = ls; sLogStd[kOutPad + j] = __expf(ls); sLogStd[2 * kOutPad + j] = 1.0f / __expf(2.0f * ls);
  }
  if (tid < kInPad) {
    const bool on = P.rms_mean != nullptr && tid < P.obs_dim;
    sNorm[tid] = on ? P.rms_mean[tid] : 0.0f;
    sNorm[kInPad + tid] = on ? 1.0f / sqrtf(P.rms_var[tid] + 1e-8f) : 1.0f;
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  // tiles of this CTA: global tile t = j * gridDim.x + blockIdx.x for j = 0 .. nloc-1; slot s takes j = s, s + 2, ...
  const int n_tiles = (P.n_env + kTile - 1) / kTile;
  const int nloc = ((int)blockIdx.x < n_tiles) ? (n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
  int cnt[kSlots];
#pragma unroll
  for (int s = 0; s < kSlots; s++) cnt[s] = (nloc - s + 1) / 2;
  // The MMA issuer walks the slot-steps (slot s, its u-th step = layer u % 4 of its (u / 4)-th tile) in the order u-major, s-minor; the
  // producer streams the weight chunks in exactly that order.

  if (warp == 0) {
    // ================================================================ weight producer (one thread)
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      bool ok = true;
      for (int u = 0; ok && u < 4 * cnt[0]; u++) {
        const int l = u & 3;
        for (int s = 0; ok && s < kSlots; s++) {
          if (u >= 4 * cnt[s]) continue;
          const char* src = reinterpret_cast<const char*>(P.w[l]);
          const int nchunk = layer_chunks(l);
          for (int c = 0; c < nchunk; c++) {
            if (!(ok = mbar_wait(bar_wempty(stage), phase ^ 1u, abort_flag, P.error))) break;
            const uint32_t dst = smem_u32(smem + kOffRing + stage * kChunkBytes), fb = bar_wfull(stage);
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(fb), "r"(kChunkBytes) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(dst),
                         "l"(src + (size_t)c * kChunkBytes), "r"(kChunkBytes), "r"(fb)
                         : "memory");
            if (++stage == kStages) { stage = 0; phase ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================================================================ MMA issuer (one thread)
    if (lane == 0) {
      TRACE_DECL;
      const int trace_region = 0; (void)trace_region;
      int stage = 0;
      uint32_t phase = 0;
      bool ok = true;
      for (int u = 0; ok && u < 4 * cnt[0]; u++) {
        const int l = u & 3;
        const int nchunk = layer_chunks(l), kc = layer_kc(l);
        const uint32_t idesc = umma_idesc(layer_n(l));
        for (int s = 0; ok && s < kSlots; s++) {
          if (u >= 4 * cnt[s]) continue;
          TRACE(s, 100 + l);
          if (!(ok = mbar_wait(bar_aready(s), (uint32_t)(u & 1), abort_flag, P.error))) break;   // this layer's A operand is in smem, TMEM is drained
          TRACE(s, 110 + l);
          asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
          const uint32_t a0 = smem_u32(smem + s * kABytes), tmem_d = tmem_base + (uint32_t)(s * kHid);
          for (int c = 0; c < nchunk; c++) {
            if (!(ok = mbar_wait(bar_wfull(stage), phase, abort_flag, P.error))) break;
            asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
            const uint32_t b0 = smem_u32(smem + kOffRing + stage * kChunkBytes);
            for (int ks = 0; ks < kc / 16; ks++) {
              const int kk = c * kc + ks * 16;
              const uint64_t ad = umma_desc(a0 + (uint32_t)(kk >> 3) * 2048u, 2048, 128);
              const uint64_t bd = umma_desc(b0 + (uint32_t)ks * 256u, 128, 16 * kc);
              umma_bf16(tmem_d, ad, bd, idesc, (c | ks) ? 1u : 0u);
            }
            umma_commit(bar_wempty(stage));   // the ring stage is free once these MMAs have read it
            if (++stage == kStages) { stage = 0; phase ^= 1u; }
          }
          if (ok) umma_commit(bar_accfull(s));
          TRACE(s, 120 + l);
        }
      }
    }
  } else {
    // ================================================================ epilogue group of slot s (4 warps, thread <-> env row of the tile)
    const int s = (warp - 2) / kEpiWarps;
    const int quad = warp & 3;                       // a warp may only touch TMEM lanes 32 * (warp % 4) .. + 31
    const int row = 32 * quad + lane;
    const int gt = 32 * ((warp - 2) % kEpiWarps) + lane;   // thread index within the group
    unsigned char* sA = smem + s * kABytes;
    unsigned char* sArow = sA + (row & 7) * 16 + (row >> 3) * 128;
    const uint32_t tmem_row = tmem_base + ((uint32_t)(32 * quad) << 16) + (uint32_t)(s * kHid);
    // Global I/O of a tile is three CONTIGUOUS blocks (observations, noise, actions): they move as single bulk copies (cp.async.bulk)
    // through shared-memory staging, and a thread then reads / writes its own row there. (Thread-per-row access to global memory costs
    // one sector per lane per instruction: 3 us of load-issue time per tile for the noise alone.) Staging: observations in the part of
    // the slot's A buffer that layer 0 does not read (bytes 16384 ..), actions in its last 16 KB (both free between tiles), noise in
    // a buffer of its own (prefetched a whole tile ahead). A block that is not 16-byte aligned / sized (odd batch sizes) falls back to
    // thread-per-row access.
    const int od = P.obs_dim, ad = P.act_dim;
    float* sObs = reinterpret_cast<float*>(sA + kInPad * kTile * 2);
    float* sAct = reinterpret_cast<float*>(sA + kABytes - 16384);
    float* sEps = reinterpret_cast<float*>(smem + kOffEps + s * kEpsBytes);
    const bool obs_al = (reinterpret_cast<uintptr_t>(P.obs) & 15) == 0 && ((kTile * od * 4) & 15) == 0;
    const bool eps_al = (reinterpret_cast<uintptr_t>(P.eps) & 15) == 0 && ((kTile * ad * 4) & 15) == 0 && ad <= kEpsCols;
    const bool act_al = (reinterpret_cast<uintptr_t>(P.act) & 15) == 0 && ((kTile * ad * 4) & 15) == 0 && P.mean == nullptr;
    auto tile_of = [&](int i) { return (2 * i + s) * (int)gridDim.x + (int)blockIdx.x; };
    auto rows_of = [&](int i) { return min(kTile, P.n_env - tile_of(i) * kTile); };
    auto bulk_load = [&](float* dst, const float* src, int bytes, uint32_t bar) {
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(bytes) : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(smem_u32(dst)), "l"(src),
                   "r"(bytes), "r"(bar)
                   : "memory");
    };
    uint32_t obs_phase = 0, eps_phase = 0;
    bool ok = true;
    TRACE_DECL;
    const int trace_region = 1 + s; (void)trace_region;
    if (gt == 0 && cnt[s] > 0 && obs_al && ((rows_of(0) * od * 4) & 15) == 0)
      bulk_load(sObs, P.obs + (size_t)tile_of(0) * kTile * od, rows_of(0) * od * 4, bar_obsfull(s));
    for (int i = 0; ok && i < cnt[s]; i++) {
      const int tile = tile_of(i);
      const int env = tile * kTile + row;
      const bool live = env < P.n_env;
      const int rows_valid = rows_of(i);
      const bool obs_bulk = obs_al && ((rows_valid * od * 4) & 15) == 0;
      const bool eps_bulk = eps_al && ((rows_valid * ad * 4) & 15) == 0;
      const bool act_bulk = act_al && ((rows_valid * ad * 4) & 15) == 0;
      if (gt == 0) TRACE(s, 200);
      // the noise of this tile: needed after the output layer (the previous tile's readers passed the group barrier below)
      if (gt == 0 && eps_bulk) bulk_load(sEps, P.eps + (size_t)tile * kTile * ad, rows_valid * ad * 4, bar_epsfull(s));
      // ---- layer-0 A operand: normalised observations, bf16, K padded to 64
      {
        float xin[kInPad];                           // this env's observation row (zero beyond obs_dim / the batch)
        if (obs_bulk) {
          ok = mbar_wait(bar_obsfull(s), obs_phase, abort_flag, P.error);
          obs_phase ^= 1u;
          const float* orow = sObs + row * od;        // shared memory (bank = 22 * lane + k for obs_dim 54: two-way conflicts at worst)
#pragma unroll
          for (int k = 0; k < kInPad; k++) xin[k] = (k < od) ? orow[k] : 0.0f;
        } else {
          const float* orow = P.obs + (size_t)(live ? env : 0) * od;
#pragma unroll
          for (int k = 0; k < kInPad; k++) xin[k] = (k < od) ? __ldg(orow + k) : 0.0f;
        }
        if (gt == 0) TRACE(s, 202);
        const bool norm = P.rms_mean != nullptr;
#pragma unroll
        for (int g = 0; g < kInPad / 8; g++) {
          __align__(16) __nv_bfloat162 pk[4];
#pragma unroll
          for (int j = 0; j < 4; j++) {
            const int k = 8 * g + 2 * j;
            float x0 = xin[k], x1 = xin[k + 1];
            if (norm) {
              x0 = fminf(fmaxf((x0 - sNorm[k]) * sNorm[kInPad + k], -10.0f), 10.0f);
              x1 = fminf(fmaxf((x1 - sNorm[k + 1]) * sNorm[kInPad + k + 1], -10.0f), 10.0f);
            }
            pk[j] = __floats2bfloat162_rn(live ? x0 : 0.0f, live ? x1 : 0.0f);
          }
          *reinterpret_cast<int4*>(sArow + g * 2048) = *reinterpret_cast<const int4*>(pk);
        }
      }
      if (gt == 0) {
        TRACE(s, 204);
        asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory");   // the previous tile's action store has read its staging
      }
      asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");   // (this thread's TMEM reads of the previous tile are done)
      asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");        // generic-proxy smem writes -> visible to the tensor-core proxy
      mbar_arrive(bar_aready(s));
      if (gt == 0) TRACE(s, 201);
#pragma unroll 1
      for (int l = 0; l < 4; l++) {
        const int u = 4 * i + l;
        ok = mbar_wait(bar_accfull(s), (uint32_t)(u & 1), abort_flag, P.error) && ok;
        ok = __all_sync(0xffffffffu, ok);
        if (!ok) break;
        if (gt == 0) TRACE(s, 210 + l);
        asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
        if (l < 3) {
          const float* bias = sBias + l * kHid;
#pragma unroll 1
          for (int cb = 0; cb < kHid; cb += 32) {
            uint32_t v[32];
            tmem_ld32(tmem_row + (uint32_t)cb, v);
#pragma unroll
            for (int c = 0; c < 4; c++) {
              __align__(16) __nv_bfloat162 pk[4];
              const float4 b0 = *reinterpret_cast<const float4*>(bias + cb + 8 * c), b1 = *reinterpret_cast<const float4*>(bias + cb + 8 * c + 4);
              pk[0] = __floats2bfloat162_rn(tanh_fast(__uint_as_float(v[8 * c + 0]) + b0.x), tanh_fast(__uint_as_float(v[8 * c + 1]) + b0.y));
              pk[1] = __floats2bfloat162_rn(tanh_fast(__uint_as_float(v[8 * c + 2]) + b0.z), tanh_fast(__uint_as_float(v[8 * c + 3]) + b0.w));
              pk[2] = __floats2bfloat162_rn(tanh_fast(__uint_as_float(v[8 * c + 4]) + b1.x), tanh_fast(__uint_as_float(v[8 * c + 5]) + b1.y));
              pk[3] = __floats2bfloat162_rn(tanh_fast(__uint_as_float(v[8 * c + 6]) + b1.z), tanh_fast(__uint_as_float(v[8 * c + 7]) + b1.w));
              *reinterpret_cast<int4*>(sArow + ((cb >> 3) + c) * 2048) = *reinterpret_cast<const int4*>(pk);
            }
          }
          asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
          asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
          mbar_arrive(bar_aready(s));
          if (gt == 0) TRACE(s, 220 + l);
        } else {
          // ---- output layer: the slot's A buffer is free (its last reader, the layer-3 MMAs, retired): the next tile's observations
          // start to arrive while this tile samples
          if (gt == 0 && i + 1 < cnt[s] && obs_al && ((rows_of(i + 1) * od * 4) & 15) == 0)
            bulk_load(sObs, P.obs + (size_t)tile_of(i + 1) * kTile * od, rows_of(i + 1) * od * 4, bar_obsfull(s));
          // mean -> sample, log-prob (train_ppo.py:121-126,135-140), store
          uint32_t v[32];
          tmem_ld32(tmem_row, v);
          float e[kOutPad], a_out[kOutPad];
          if (eps_bulk) {
            ok = mbar_wait(bar_epsfull(s), eps_phase, abort_flag, P.error);
            eps_phase ^= 1u;
            const float* erow = sEps + row * ad;      // shared memory (odd row stride: conflict-free)
#pragma unroll
            for (int j = 0; j < kOutPad; j++) e[j] = (j < ad) ? erow[j] : 0.0f;
          } else {
            const float* erow = P.eps + (size_t)(live ? env : 0) * ad;
#pragma unroll
            for (int j = 0; j < kOutPad; j++) e[j] = (j < ad) ? __ldg(erow + j) : 0.0f;
          }
          float lp = 0.0f;
          const float* bias = sBias + 3 * kHid;
#pragma unroll
          for (int j = 0; j < kOutPad; j++) {
            const float mu = __uint_as_float(v[j]) + bias[j];
            const float a = mu + sLogStd[kOutPad + j] * e[j];
            const float d = a - mu;
            if (j < ad) lp += d * d * sLogStd[2 * kOutPad + j] + 2.0f * sLogStd[j] + 1.8378770664093453f;
            a_out[j] = a;
            v[j] = __float_as_uint(mu);
          }
          if (act_bulk) {
            float* arow = sAct + row * ad;
#pragma unroll
            for (int j = 0; j < kOutPad; j++) if (j < ad) arow[j] = a_out[j];
          } else if (live) {
            float* arow = P.act + (size_t)env * ad;
#pragma unroll
            for (int j = 0; j < kOutPad; j++) if (j < ad) arow[j] = a_out[j];
          }
          if (live) {
            if (P.mean) {
#pragma unroll
              for (int j = 0; j < kOutPad; j++) if (j < ad) P.mean[(size_t)env * ad + j] = __uint_as_float(v[j]);
            }
            P.logp[env] = -0.5f * lp;
          }
          asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
          asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");        // staged actions -> visible to the bulk store
          asm volatile("bar.sync %0, %1;\n" ::"r"(1 + s), "r"(kEpiWarps * 32) : "memory");   // the group: noise consumed, actions staged
          if (gt == 0 && act_bulk) {
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;\n" ::"l"(P.act + (size_t)tile * kTile * ad),
                         "r"(smem_u32(sAct)), "r"(rows_valid * ad * 4)
                         : "memory");
            asm volatile("cp.async.bulk.commit_group;\n" ::: "memory");
          }
          if (gt == 0) TRACE(s, 223);
        }
      }
    }
    if (gt == 0) asm volatile("cp.async.bulk.wait_group 0;\n" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
  }
}

// W [K, N] row-major float32 (x @ W convention) -> bf16 image of B = W^T, zero padded to [Np, Kp], as Kp / Kc chunks of 16 KB, each the
// canonical K-major layout of Kc = 16384 / (2 Np) columns of K (32 for the hidden layers, 256 for the output layer)
__global__ void pack_weight_kernel(const float* __restrict__ w, int K, int N, int Kp, int Np, __nv_bfloat16* __restrict__ out) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= Np * Kp) return;
  const int n = idx / Kp, k = idx % Kp;
  const int Kc = kChunkBytes / (2 * Np), chunk = k / Kc, kk = k % Kc;
  const float v = (n < N && k < K) ? w[(size_t)k * N + n] : 0.0f;
  const size_t byte = (size_t)chunk * kChunkBytes + (size_t)(n & 7) * 16 + (size_t)(n >> 3) * (Kc / 8) * 128 + (size_t)(kk >> 3) * 128 + (size_t)(kk & 7) * 2;
  out[byte / 2] = __float2bfloat16_rn(v);
}

// GAE reverse scan (reference train_ppo.py:171-202): one thread per env walks the rollout backwards; the loads of a step do not depend
// on the carried advantage, so the unrolled loop keeps eight steps of loads in flight (coalesced across envs)
__global__ void gae_kernel(int T, int n, const float* __restrict__ r, const float* __restrict__ v, const float* __restrict__ te,
                           const float* __restrict__ tr, float gamma, float lam, float* __restrict__ adv, float* __restrict__ ret) {
  const int env = blockIdx.x * blockDim.x + threadIdx.x;
  if (env >= n) return;
  float carry = 0.0f, vnext = v[(size_t)T * n + env];
#pragma unroll 8
  for (int t = T - 1; t >= 0; t--) {
    const size_t i = (size_t)t * n + env;
    const float vt = v[i], term = te[i], trunc = tr[i];
    const float delta = r[i] + gamma * vnext * (1.0f - term) - vt;
    carry = delta + gamma * lam * (1.0f - fmaxf(term, trunc)) * carry;
    adv[i] = carry;
    ret[i] = carry + vt;
    vnext = vt;
  }
}

// Learner backward helper: dz = dy * (1 - y^2) (tanh backward; y == NULL: dz = dy, no write) fused with the bias gradient
// db[c] += sum_rows dz[:, c]. Thread j owns column j of a row block (coalesced across the CTA); one atomicAdd per column per CTA.
__global__ void tanh_bwd_colsum_kernel(int n, int c, const float* __restrict__ dy, const float* __restrict__ y, float* __restrict__ dz,
                                       float* __restrict__ db) {
  const int rows_per_cta = (n + gridDim.x - 1) / gridDim.x;
  const int r0 = blockIdx.x * rows_per_cta, r1 = min(n, r0 + rows_per_cta);
  for (int col = threadIdx.x; col < c; col += blockDim.x) {
    float acc = 0.0f;
#pragma unroll 8
    for (int r = r0; r < r1; r++) {
      const size_t i = (size_t)r * c + col;
      float g = dy[i];
      if (y != nullptr) {
        const float t = y[i];
        g *= 1.0f - t * t;
        dz[i] = g;
      }
      acc += g;
    }
    if (r1 > r0) atomicAdd(db + col, acc);
  }
}

// FP32 FMA-pipe ceiling, measured: 8 independent FFMA chains per thread (no memory traffic), 1024 threads per CTA, 2 CTAs per SM.
__global__ void __launch_bounds__(1024, 2) ffma_peak_kernel(float* sink, int iters, float seed) {
  float a0 = seed + threadIdx.x, a1 = a0 + 1.f, a2 = a0 + 2.f, a3 = a0 + 3.f, a4 = a0 + 4.f, a5 = a0 + 5.f, a6 = a0 + 6.f, a7 = a0 + 7.f;
  const float m = 0.999999f, c = 1e-7f * seed;
#pragma unroll 1
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int u = 0; u < 16; u++) {
      a0 = fmaf(a0, m, c); a1 = fmaf(a1, m, c); a2 = fmaf(a2, m, c); a3 = fmaf(a3, m, c);
      a4 = fmaf(a4, m, c); a5 = fmaf(a5, m, c); a6 = fmaf(a6, m, c); a7 = fmaf(a7, m, c);
    }
  }
  const float r = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
  if (r == 123.456f) sink[0] = r;   // never true: keeps the chains alive
}

}  // namespace mjxbp

extern "C" {

int mjxb_ffma_peak(int32_t device, float* tflops_out, float* ms_out) {
  if (!tflops_out) return MJXB_EINVAL;
  int cur = 0, nsm = 0;
  if (cudaGetDevice(&cur) != cudaSuccess) { cudaGetLastError(); return MJXB_ENOGPU; }
  if (cudaSetDevice(device) != cudaSuccess) { cudaGetLastError(); return MJXB_EINVAL; }
  cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, device);
  float* sink = nullptr;
  cudaEvent_t e0, e1;
  if (cudaMalloc(&sink, 4) != cudaSuccess) { cudaGetLastError(); cudaSetDevice(cur); return MJXB_ECUDA; }
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int grid = nsm * 2, iters = 4096;
  float best = 1e30f;
  for (int rep = 0; rep < 6; rep++) {   // first repetitions warm up clocks / instruction cache; best of the rest
    cudaEventRecord(e0, 0);
    mjxbp::ffma_peak_kernel<<<grid, 1024>>>(sink, iters, 1.0f + rep);
    cudaEventRecord(e1, 0);
    cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    if (rep >= 2 && ms < best) best = ms;
  }
  const cudaError_t err = cudaGetLastError();
  cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(sink);
  cudaSetDevice(cur);
  if (err != cudaSuccess) return MJXB_ECUDA;
  const double flops = 2.0 * 8 * 16 * (double)iters * 1024.0 * grid;
  *tflops_out = (float)(flops / (best * 1e-3) / 1e12);
  if (ms_out) *ms_out = best;
  return MJXB_OK;
}

#if MJXB_POLICY_TRACE
int mjxb_policy_trace(unsigned long long* host_out) {   // profiling variant only (not part of include/mjxb.h)
  cudaDeviceSynchronize();
  const int rc = cudaMemcpyFromSymbol(host_out, mjxbp::g_trace, sizeof(unsigned long long) * 4096) == cudaSuccess ? 0 : MJXB_ECUDA;
  static unsigned long long zeros[4096];
  cudaMemcpyToSymbol(mjxbp::g_trace, zeros, sizeof(zeros));
  return rc;
}
#endif

int mjxb_tanh_bwd_colsum(int32_t n, int32_t c, const float* dy, const float* y, float* dz, float* db_zeroed, void* stream) {
  if (n <= 0 || c <= 0 || !dy || !db_zeroed || (y != nullptr && dz == nullptr)) return MJXB_EINVAL;
  int grid = (n + 63) / 64;
  if (grid > 148 * 8) grid = 148 * 8;
  const int threads = c >= 256 ? 256 : ((c + 31) / 32) * 32;
  g_mjxb_launches++;
  mjxbp::tanh_bwd_colsum_kernel<<<grid, threads, 0, (cudaStream_t)stream>>>(n, c, dy, y, dz, db_zeroed);
  return cudaGetLastError() == cudaSuccess ? MJXB_OK : MJXB_ECUDA;
}

int mjxb_gae(int32_t rollout_length, int32_t n_env, const float* reward, const float* value, const float* terminated,
             const float* truncated, float gamma, float lam, float* advantage, float* ret, void* stream) {
  if (rollout_length <= 0 || n_env <= 0 || !reward || !value || !terminated || !truncated || !advantage || !ret) return MJXB_EINVAL;
  g_mjxb_launches++;
  mjxbp::gae_kernel<<<(n_env + 127) / 128, 128, 0, (cudaStream_t)stream>>>(rollout_length, n_env, reward, value, terminated, truncated,
                                                                           gamma, lam, advantage, ret);
  return cudaGetLastError() == cudaSuccess ? MJXB_OK : MJXB_ECUDA;
}

int mjxb_policy_pack_weight(const float* w, int32_t k, int32_t n, int32_t k_pad, int32_t n_pad, void* out_bf16, void* stream) {
  if (!w || !out_bf16 || k <= 0 || n <= 0 || k_pad < k || n_pad < n || (k_pad % 16) || (n_pad % 16)) return MJXB_EINVAL;
  // the image is a sequence of 16 KB chunks (Kc = 8192 / n_pad columns of K each): the two shapes the kernel streams
  if (!((n_pad == mjxbp::kHid && k_pad % 32 == 0) || (n_pad == mjxbp::kOutPad && k_pad == mjxbp::kHid))) return MJXB_EUNSUPPORTED;
  const int total = n_pad * k_pad;
  g_mjxb_launches++;
  mjxbp::pack_weight_kernel<<<(total + 255) / 256, 256, 0, (cudaStream_t)stream>>>(w, k, n, k_pad, n_pad, (__nv_bfloat16*)out_bf16);
  return cudaGetLastError() == cudaSuccess ? MJXB_OK : MJXB_ECUDA;
}

int mjxb_policy_act(int32_t n_env, int32_t obs_dim, int32_t act_dim, const float* obs, const float* rms_mean, const float* rms_var,
                    const void* const* w_packed, const float* const* bias, const float* log_std, const float* eps, float* act,
                    float* logp, float* mean, int32_t* error_flag, void* stream) {
  if (n_env <= 0 || !obs || !w_packed || !bias || !log_std || !eps || !act || !logp) return MJXB_EINVAL;
  if (obs_dim <= 0 || obs_dim > mjxbp::kInPad || act_dim <= 0 || act_dim > mjxbp::kOutPad) return MJXB_EUNSUPPORTED;
  if ((rms_mean == nullptr) != (rms_var == nullptr)) return MJXB_EINVAL;
  int sms = 0;
  {  // the opt-in shared-memory size is a per-device function attribute: set it once per device, thread-safely
    static std::mutex mu;
    static bool attr_set[64] = {};
    static int num_sms[64] = {};
    int devid = 0;
    if (cudaGetDevice(&devid) != cudaSuccess || devid < 0 || devid >= 64) { cudaGetLastError(); return MJXB_ECUDA; }
    std::lock_guard<std::mutex> lock(mu);
    if (!attr_set[devid]) {
      if (cudaFuncSetAttribute(mjxbp::policy_act_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, mjxbp::kSmemBytes) != cudaSuccess) {
        cudaGetLastError();
        return MJXB_ECUDA;
      }
      if (cudaDeviceGetAttribute(&num_sms[devid], cudaDevAttrMultiProcessorCount, devid) != cudaSuccess || num_sms[devid] <= 0) {
        cudaGetLastError();
        return MJXB_ECUDA;
      }
      attr_set[devid] = true;
    }
    sms = num_sms[devid];
  }
  mjxbp::PolicyArgs P;
  P.n_env = n_env; P.obs_dim = obs_dim; P.act_dim = act_dim; P.obs = obs; P.rms_mean = rms_mean; P.rms_var = rms_var;
  for (int i = 0; i < 4; i++) {
    if (!w_packed[i] || !bias[i]) return MJXB_EINVAL;
    P.w[i] = w_packed[i]; P.b[i] = bias[i];
  }
  P.log_std = log_std; P.eps = eps; P.act = act; P.logp = logp; P.mean = mean; P.error = error_flag;
  // persistent: one CTA per SM, each walks its env tiles two at a time (a batch of <= 148 tiles runs one tile per CTA)
  int grid = (n_env + mjxbp::kTile - 1) / mjxbp::kTile;
  if (grid > sms) grid = sms;
  g_mjxb_launches++;
  mjxb::launch_pdl(mjxbp::policy_act_kernel, dim3(grid), dim3(mjxbp::kThreads), (size_t)mjxbp::kSmemBytes, (cudaStream_t)stream, P);
  return cudaGetLastError() == cudaSuccess ? MJXB_OK : MJXB_ECUDA;
}

}  // extern "C"
