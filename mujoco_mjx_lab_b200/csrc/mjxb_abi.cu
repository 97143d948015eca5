// mjxb_abi.cu -- the extern "C" boundary declared in include/mjxb.h (host side: model upload, launches, host-buffer arena).
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <atomic>
#include <mutex>
#include <new>
#include <vector>

#ifndef MJXB_EXACT
#define MJXB_EXACT 0
#endif
#include "mjxb.h"
#include "mjxb_device.cuh"
#include "mjxb_internal.h"

using namespace mjxb;

// every kernel this library launches is counted here (bench.py reports the count of its timed region as `gpu_launches`)
std::atomic<long long> g_mjxb_launches{0};

namespace {

constexpr int kMaxWarps = WARPS_MAIN > WARPS_MID ? (WARPS_MAIN > WARPS_BIG ? WARPS_MAIN : WARPS_BIG) : (WARPS_MID > WARPS_BIG ? WARPS_MID : WARPS_BIG);
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

// Tuning switches, read ONCE at mjxb_model_create (flags argument of mjxb_model_create_ex, or the MJXB_* environment variables for
// experiments): nothing on the launch path calls getenv.
struct Tuning {
  int lockstep = 1;        // 1: CTA barriers at the round top and at the solver entry / exit; 3: also at every factor/solve round; 0: none
  int lockstep_group = 0;  // warps per barrier group (0 = the whole CTA)
  bool inline_reset = false, sync_tiers = false, spec_reset = true;
  bool host_direct = true, direct_obs = true, direct_scalars = true;
  int host_chunks = 0;
  bool balance_rounds = true;
  bool dyn_spec = true;      // ... also in the rounds that run reset warps; MJXB_DYN_SPEC
  bool dyn_rounds = true;    // groups of envs taken from a device-wide counter (batches of >= 2 rounds per CTA without reset warps); MJXB_DYN_ROUNDS
  bool skip_mid = true;      // latency regime: the big tier consumes the main tier's overflow list directly; MJXB_SKIP_MID=0 keeps three tiers
  int spec_max_rounds = 6;   // concurrent auto-reset (reset warps beside the stepping warps) for batches of up to this many rounds; MJXB_SPEC_MAX_ROUNDS
  int sort_min_env = 16384;  // work-sorted scheduling from this batch size on (0 = never); MJXB_SORT_MIN_ENV
  int sort_seg_shift = 11;   // segment = 2^shift envs (one sort CTA each: 128 CTAs at 262,144 envs; 2^15-env segments cost 2-3 % at 65,536 envs:
                             // the sort's own latency); never larger than an input chunk of the host pipeline
};

// Per-stream launch scratch (overflow counters + lists + per-CTA reset queues). One entry per stream that has launched on the model, so
// launches on distinct streams never share counters; an entry is never freed or resized once handed to a launch (a larger batch gets a
// new buffer, the old one is retired until mjxb_model_destroy), so CUDA graphs captured earlier stay valid.
struct Scratch {
  cudaStream_t stream = nullptr;
  int* buf = nullptr;   // [0] countA, [1] doneA, [2] countB, [3] doneB, [4..4+cap) listA (main -> mid), [4+cap..) listB (mid -> big), then reset queues,
                        // then the work-sorted schedule: perm [cap] ints and the cost keys [cap] bytes
  int cap = 0;
};

struct mjxb_model {
  DevModel host;
  DevModel* dev = nullptr;
  PairParam* dev_pp = nullptr;
  int device = 0, num_sms = 0, warps = 0, warps_mid = 0, warps_big = 0;
  size_t smem = 0, smem_mid = 0, smem_big = 0;
  Tuning tune;
  mutable std::mutex scratch_mu;
  mutable std::vector<Scratch> scratch;      // live entries, one per stream
  mutable std::vector<int*> scratch_retired; // superseded buffers (freed at destroy)
  Arena arena;
  size_t sched_offset(int n_env) const { return 3 * (size_t)n_env + 4 + (size_t)num_sms * kMaxWarps * 2; }   // ints before perm
  size_t dyn_offset(int n_env) const { return sched_offset(n_env) + (size_t)n_env + ((size_t)n_env + 3) / 4; }   // ints before the group counter
  // ... then 4 counter ints and the reset queues of the dynamic rounds (twice as deep as the static ones: a CTA may take up to twice its share)
  size_t scratch_ints(int n_env) const { return dyn_offset(n_env) + 4 + 2 * (size_t)n_env + (size_t)num_sms * kMaxWarps * 4; }
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
  if (m->tune.host_chunks > 0) nchunk = m->tune.host_chunks;
  a.chunk = (n + nchunk - 1) / nchunk;
  for (int i = 0; i < Arena::kSlots; i++) {
    CU(cudaStreamCreateWithFlags(&a.pipe[i], cudaStreamNonBlocking));
    CU(cudaMalloc(&a.pipe_ovf[i], m->scratch_ints(a.chunk) * sizeof(int)));
    CU(cudaMemsetAsync(a.pipe_ovf[i], 0, 4 * sizeof(int), a.stream));
    CU(cudaMemsetAsync(a.pipe_ovf[i] + m->sched_offset(a.chunk) + a.chunk, 0, (size_t)a.chunk, a.stream));
    CU(cudaMemsetAsync(a.pipe_ovf[i] + m->dyn_offset(a.chunk), 0, 4 * sizeof(int), a.stream));
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


using KMain = void (*)(const DevModel*, const PairParam*, StepArgs);

// the stream's scratch entry, created / grown here (never while the stream is capturing: mjxb_model_reserve must have run before)
int scratch_for(const mjxb_model* m, cudaStream_t stream, int n_env, int** buf, int* cap) {
  std::lock_guard<std::mutex> lock(m->scratch_mu);
  Scratch* s = nullptr;
  for (Scratch& e : m->scratch) if (e.stream == stream) { s = &e; break; }
  if (s != nullptr && s->cap >= n_env) { *buf = s->buf; *cap = s->cap; return MJXB_OK; }
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(stream, &cs) != cudaSuccess) cudaGetLastError();
  if (cs != cudaStreamCaptureStatusNone) {
    snprintf(g_cuda_err, sizeof(g_cuda_err), "launch scratch for %d envs is missing on a capturing stream: call mjxb_model_reserve first", n_env);
    return MJXB_ECUDA;
  }
  int* nb = nullptr;
  CU(cudaMalloc(&nb, m->scratch_ints(n_env) * sizeof(int)));
  CU(cudaMemsetAsync(nb, 0, 4 * sizeof(int), stream));
  CU(cudaMemsetAsync(nb + m->sched_offset(n_env) + n_env, 0, (size_t)n_env, stream));   // cost keys: no hint yet
  CU(cudaMemsetAsync(nb + m->dyn_offset(n_env), 0, 4 * sizeof(int), stream));           // group counter of the dynamic rounds
  if (s == nullptr) { m->scratch.push_back(Scratch()); s = &m->scratch.back(); s->stream = stream; }
  else m->scratch_retired.push_back(s->buf);  // earlier launches / captured graphs may still reference it
  s->buf = nb; s->cap = n_env;
  *buf = nb; *cap = n_env;
  return MJXB_OK;
}

int launch(const mjxb_model* m, const StepArgs& args_in, bool dbg, cudaStream_t stream, int* ovf_buf = nullptr, int ovf_buf_cap = 0) {
  int cur = 0;
  CU(cudaGetDevice(&cur));
  if (cur != m->device) CU(cudaSetDevice(m->device));
  int* ovf = ovf_buf;
  int cap = ovf_buf_cap;
  if (ovf == nullptr) {
    const int rc = scratch_for(m, stream, args_in.n_env, &ovf, &cap);
    if (rc != MJXB_OK) { if (cur != m->device) cudaSetDevice(cur); return rc; }
  }
  StepArgs args = args_in;
  int* listA = ovf + 4;
  int* listB = ovf + 4 + cap;
  args.in_count = nullptr; args.in_list = nullptr; args.in_done = nullptr; args.out_count = ovf; args.out_list = listA;
  args.reset_list = m->tune.inline_reset ? nullptr : ovf + 4 + 2 * (size_t)cap;
  // 1 (default): CTA barriers at the round top and at the solver entry / exit; 3: additionally at every factor/solve round (was the
  // better choice before the solver's dependent chains were shortened: 35.2 M against 36.4 M now); 0: none (profiling aid)
  args.lockstep = m->tune.lockstep; args.lockstep_group = m->tune.lockstep_group;
  // small batches: spread the envs over every SM (fewer warps per CTA run faster than 16 sharing one SM's issue slots)
  int warps = m->warps;
  const int per_sm = (args.n_env + m->num_sms - 1) / m->num_sms;
  if (per_sm < warps) warps = per_sm < 1 ? 1 : per_sm;
  else if (m->tune.balance_rounds) {   // a batch of a few rounds: equal rounds (4096 envs: 2 x 14 warps per SM, not 16 + 12) -- fewer warps share the SM
    const int rounds = (per_sm + warps - 1) / warps;
    const int bal = (per_sm + rounds - 1) / rounds;
    if (bal < warps) warps = bal;
  }
  int grid = (args.n_env + warps - 1) / warps;
  if (grid > m->num_sms) grid = m->num_sms;
  const bool ls_ = m->host.ls_exact != 0 && m->host.solver == 2;
  const bool single_ = !dbg && ls_ && args.nsteps == 1 && (!(args.mode == MODE_ENV_STEP && args.autoreset) || args.reset_list != nullptr);
  // concurrent auto-reset for batches of a few rounds: `steppers` stepping warps + reset warps per CTA; an env that finishes its episode
  // is re-initialised by a reset warp beside the step round instead of in a packed pass after the step rounds (bit-identical results;
  // 1024 envs: 105.6 -> 81 us per step, 2048 envs: 142 -> 106 us, 4096 envs 21.5 -> 26.0 M env-steps/s, 8192 envs 25.1 -> 27.7 M). 10,240 / 12,288 envs
  // (5 / 6 rounds) +7 / +5 %. From seven rounds on (14,000 envs: 29.4 against 28.4 M) all 16 warps step and the resets run in packed rounds at the end (2 rounds in
  // 113 at 262,144 envs).
  args.spec_reset = 0;
  if (m->tune.spec_reset && single_ && args.mode == MODE_ENV_STEP && args.autoreset) {
    const int max_step = m->warps - 2;                                      // at least two reset warps
    const int rounds = (per_sm + max_step - 1) / (max_step < 1 ? 1 : max_step);
    if (max_step >= 1 && rounds <= m->tune.spec_max_rounds) {
      int steppers = (per_sm + rounds - 1) / (rounds < 1 ? 1 : rounds);
      if (steppers < 1) steppers = 1;
      int resetters = m->warps - steppers;
      if (resetters > steppers) resetters = steppers;
      args.spec_reset = steppers;
      warps = steppers + resetters;
      grid = (args.n_env + steppers - 1) / steppers;
      if (grid > m->num_sms) grid = m->num_sms;
    }
  }
  // work-sorted scheduling: the envs of every segment are dealt to the CTAs by descending cost key of their previous step
  args.perm = nullptr; args.work_out = nullptr;
  int extra_launches = 0;
  if (m->tune.sort_min_env > 0 && args.n_env >= m->tune.sort_min_env && (args.mode == MODE_ENV_STEP || args.mode == MODE_PHYS_STEP) && !dbg) {
    int* perm = ovf + m->sched_offset(cap);
    uint8_t* work = reinterpret_cast<uint8_t*>(perm + cap);
    int shift = m->tune.sort_seg_shift;
    if (args.in_ready != nullptr && args.in_ready_shift < shift) shift = args.in_ready_shift;   // streamed inputs: never reorder across chunks
    launch_pdl(mjxb_sort_work_kernel, dim3((args.n_env + (1 << shift) - 1) >> shift), dim3(1024), 0, stream, (const uint8_t*)work, perm, args.n_env, shift);
    args.perm = perm; args.work_out = work;
    extra_launches = 1;
  }
  const size_t smem_main = m->smem - (size_t)(m->warps - warps) * sizeof(WarpS<CAP_MAIN, MAXCC_MAIN>);
  args.reset_stride = ((args.n_env + grid * warps - 1) / (grid * warps)) * warps;
  // dynamic rounds: from two rounds per CTA on, groups of envs (one per stepping warp) are taken from a device-wide counter
  args.dyn_counter = nullptr; args.dyn_done = nullptr; args.dyn_max_rounds = 0;
  {
    const int per_round = grid * (args.spec_reset != 0 ? args.spec_reset : warps);
    const int rounds = (args.n_env + per_round - 1) / per_round;
    if (m->tune.dyn_rounds && single_ && (args.spec_reset == 0 || m->tune.dyn_spec) && rounds >= 2 &&
        (args.mode == MODE_ENV_STEP || args.mode == MODE_PHYS_STEP) && !dbg) {
      int* dynp = ovf + m->dyn_offset(cap);
      args.dyn_counter = dynp; args.dyn_done = dynp + 1;
      args.dyn_max_rounds = 2 * rounds;                       // grid * 2 * rounds * warps <= 2 * (n_env + grid * warps) queue slots
      args.reset_stride = args.dyn_max_rounds * warps;
      if (args.reset_list != nullptr) args.reset_list = dynp + 4;
    }
  }
  const bool ls = m->host.ls_exact != 0 && m->host.solver == 2;  // fast instantiation: Newton + exact line search; else the general one
#define MJXB_LAUNCH(CAPv, CCv, Wv, G, B, SM)                                                                                   \
  do {                                                                                                                        \
    if (dbg && ls) launch_pdl(mjxb_step_kernel<true, CAPv, CCv, Wv, true>, dim3(G), dim3(B), SM, stream, (const DevModel*)m->dev, (const PairParam*)m->dev_pp, args);   \
    else if (dbg) launch_pdl(mjxb_step_kernel<true, CAPv, CCv, Wv, false>, dim3(G), dim3(B), SM, stream, (const DevModel*)m->dev, (const PairParam*)m->dev_pp, args);   \
    else if (ls) launch_pdl(mjxb_step_kernel<false, CAPv, CCv, Wv, true>, dim3(G), dim3(B), SM, stream, (const DevModel*)m->dev, (const PairParam*)m->dev_pp, args);    \
    else launch_pdl(mjxb_step_kernel<false, CAPv, CCv, Wv, false>, dim3(G), dim3(B), SM, stream, (const DevModel*)m->dev, (const PairParam*)m->dev_pp, args);           \
  } while (0)
  // hot instantiation: one step per launch, Newton + exact line search, resets deferred (or none requested)
  const bool single = !dbg && ls && args.nsteps == 1 && (!(args.mode == MODE_ENV_STEP && args.autoreset) || args.reset_list != nullptr);
  if (single && args.dyn_counter != nullptr)
    launch_pdl(mjxb_step_kernel<false, CAP_MAIN, MAXCC_MAIN, WARPS_MAIN, true, true, true>, dim3(grid), dim3(warps * 32), smem_main, stream,
               (const DevModel*)m->dev, (const PairParam*)m->dev_pp, args);
  else if (single) launch_pdl(mjxb_step_kernel<false, CAP_MAIN, MAXCC_MAIN, WARPS_MAIN, true, true>, dim3(grid), dim3(warps * 32), smem_main, stream,
                         (const DevModel*)m->dev, (const PairParam*)m->dev_pp, args);
  else MJXB_LAUNCH(CAP_MAIN, MAXCC_MAIN, WARPS_MAIN, grid, warps * 32, smem_main);
  // Every kernel boundary of a step costs ~6 us at small batches (14 us at 4096 envs), also for an overflow tier that finds its list empty
  // and leaves at once (measured by skipping them: 1024 envs 81 -> 69 us per step). The latency regime (concurrent-reset launches,
  // <= 12,432 envs with auto-reset) therefore runs ONE overflow tier: the big one consumes the main tier's list directly. It holds every
  // row of the model, so results are the same; what is given up is the mid tier's higher throughput (10 instead of 3 warps per SM)
  // when many envs overflow, which an auto-resetting batch of that size does not have (5e-6 of the env-steps in the bench distribution).
  const bool skip_mid = args.spec_reset != 0 && m->tune.skip_mid;
  g_mjxb_launches += (skip_mid ? 2 : 3) + extra_launches;   // (the schedule sort) + main tier + the overflow tiers below (each leaves at once when its list is empty)
  cudaError_t e = cudaGetLastError();
  const bool sync_tiers = m->tune.sync_tiers;  // debugging aid: attribute a device fault to its tier
  if (sync_tiers && e == cudaSuccess) { e = cudaStreamSynchronize(stream); if (e != cudaSuccess) fprintf(stderr, "[mjxb] main tier failed: %s\n", cudaGetErrorString(e)); }
  args.in_ready = nullptr;  // only the first pass waits for streamed inputs
  args.perm = nullptr;      // (the overflow tiers read their own lists; they still leave cost keys)
  args.dyn_counter = nullptr; args.dyn_done = nullptr;
  if (e == cudaSuccess && !skip_mid) {  // mid tier (64 rows / 24 contacts) over the envs the main tile could not hold; usually few: exits at once when empty
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
    if (skip_mid) { args.in_count = ovf; args.in_done = ovf + 1; args.in_list = listA; }
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

namespace mjxb {
int model_view(const mjxb_model* m, ModelView* out) {
  if (!m || !out) return MJXB_EINVAL;
  out->host = &m->host; out->dev = m->dev; out->dev_pp = m->dev_pp; out->device = m->device; out->num_sms = m->num_sms;
  return MJXB_OK;
}
int model_scratch(const mjxb_model* m, cudaStream_t stream, int n_env, int** buf, int* cap) { return scratch_for(m, stream, n_env, buf, cap); }
int report_cuda_error(cudaError_t e, const char* what) { return cuda_fail(e, what); }
bool pdl_enabled() {
  static const bool on = [] { const char* e = getenv("MJXB_PDL"); return e ? atoi(e) != 0 : true; }();
  return on;
}
}  // namespace mjxb

extern "C" {

int mjxb_abi_version(void) { return MJXB_ABI_VERSION; }
long long mjxb_launch_count(void) { return g_mjxb_launches.load(); }
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

static int env_int(const char* name, int dflt) { const char* e = getenv(name); return e ? atoi(e) : dflt; }

int mjxb_model_create(const void* blob, size_t blob_bytes, const mjxb_env_config* cfg, int device, mjxb_model** out) {
  // the experiment switches of the environment are folded into flags here, once
  uint32_t flags = 0;
  if (getenv("MJXB_LS_ITERATIVE")) flags |= MJXB_FLAG_LS_ITERATIVE;
  if (getenv("MJXB_DENSE_CHOL")) flags |= MJXB_FLAG_DENSE_CHOL;
  if (getenv("MJXB_INLINE_RESET")) flags |= MJXB_FLAG_INLINE_RESET;
  return mjxb_model_create_ex(blob, blob_bytes, cfg, device, flags, out);
}

int mjxb_model_create_ex(const void* blob, size_t blob_bytes, const mjxb_env_config* cfg, int device, uint32_t flags, mjxb_model** out) {
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
  {  // the generated leaves-first elimination (mjxb_chol_tree.cuh) applies when the model's dof tree is the one it was generated for
    DevModel& D = m->host;
    D.tree_chol_ok = (b.nv == kTreeNV) ? 1 : 0;
    for (int d = 0; d < b.nv && d < kTreeNV; d++) if (b.dof_parent[d] != kTreeDofParent[d]) D.tree_chol_ok = 0;
    if (flags & MJXB_FLAG_DENSE_CHOL) D.tree_chol_ok = 0;
#if MJXB_EXACT
    D.ls_exact = 0;  // the reference-arithmetic build always runs MJX's bracketed line search
#endif
    if (flags & MJXB_FLAG_LS_ITERATIVE) D.ls_exact = 0;
  }
  m->tune.inline_reset = (flags & MJXB_FLAG_INLINE_RESET) != 0;
  m->tune.lockstep = env_int("MJXB_LOCKSTEP", 1);
  m->tune.lockstep_group = env_int("MJXB_LOCKSTEP_GROUP", 0);
  m->tune.sync_tiers = getenv("MJXB_SYNC_TIERS") != nullptr;
  m->tune.spec_reset = env_int("MJXB_SPEC_RESET", 1) != 0 && !m->tune.inline_reset && !(flags & MJXB_FLAG_NO_SPEC_RESET);
  m->tune.host_chunks = env_int("MJXB_HOST_CHUNKS", 0);
  m->tune.sort_min_env = (flags & MJXB_FLAG_NO_WORK_SORT) ? 0 : env_int("MJXB_SORT_MIN_ENV", m->tune.sort_min_env);
  m->tune.sort_seg_shift = env_int("MJXB_SORT_SEG_SHIFT", m->tune.sort_seg_shift);
  m->tune.balance_rounds = env_int("MJXB_BALANCE_ROUNDS", 1) != 0;
  m->tune.dyn_rounds = env_int("MJXB_DYN_ROUNDS", 1) != 0 && !(flags & MJXB_FLAG_NO_DYN_ROUNDS);
  m->tune.dyn_spec = env_int("MJXB_DYN_SPEC", 1) != 0;
  m->tune.spec_max_rounds = env_int("MJXB_SPEC_MAX_ROUNDS", m->tune.spec_max_rounds);
  m->tune.skip_mid = env_int("MJXB_SKIP_MID", 1) != 0;
  m->tune.host_direct = env_int("MJXB_HOST_DIRECT", 1) != 0;
  m->tune.direct_obs = env_int("MJXB_DIRECT_OBS", 1) != 0;
  m->tune.direct_scalars = env_int("MJXB_DIRECT_SCALARS", 1) != 0;
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
  CUX(cudaFuncSetAttribute(mjxb_step_kernel<false, CAP_MAIN, MAXCC_MAIN, WARPS_MAIN, true, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)m->smem));
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
  for (Scratch& e : m->scratch) if (e.buf) cudaFree(e.buf);
  for (int* p : m->scratch_retired) if (p) cudaFree(p);
  delete m;
}

int mjxb_model_dims(const mjxb_model* m, int32_t dims[8]) {
  if (!m || !dims) return MJXB_EINVAL;
  const DevModel& C = m->host;
  dims[0] = C.nq; dims[1] = C.nv; dims[2] = C.nu; dims[3] = C.nbody; dims[4] = C.ncon; dims[5] = C.nefc; dims[6] = C.nsensor;
  dims[7] = C.cfg.obs_dim;
  return MJXB_OK;
}

size_t mjxb_model_scratch_bytes(const mjxb_model* m) {
  if (!m) return 0;
  std::lock_guard<std::mutex> lock(m->scratch_mu);
  size_t total = 0;
  for (const Scratch& e : m->scratch) total += m->scratch_ints(e.cap) * sizeof(int);
  return total;
}

int mjxb_model_reserve(const mjxb_model* m, int32_t n_env, void* stream) {
  if (!m || n_env <= 0) return MJXB_EINVAL;
  int cur = 0;
  CU(cudaGetDevice(&cur));
  if (cur != m->device) CU(cudaSetDevice(m->device));
  int* buf = nullptr;
  int cap = 0;
  const int rc = scratch_for(m, (cudaStream_t)stream, n_env, &buf, &cap);
  if (cur != m->device) cudaSetDevice(cur);
  return rc;
}

int mjxb_model_flags(const mjxb_model* m) {
  if (!m) return MJXB_EINVAL;
  int f = 0;
  if (!m->host.ls_exact) f |= MJXB_FLAG_LS_ITERATIVE;
  if (!m->host.tree_chol_ok) f |= MJXB_FLAG_DENSE_CHOL;
  if (m->tune.inline_reset) f |= MJXB_FLAG_INLINE_RESET;
  if (!m->tune.spec_reset) f |= MJXB_FLAG_NO_SPEC_RESET;
  if (m->tune.sort_min_env <= 0) f |= MJXB_FLAG_NO_WORK_SORT;
  if (!m->tune.dyn_rounds) f |= MJXB_FLAG_NO_DYN_ROUNDS;
#if MJXB_EXACT
  f |= MJXB_FLAG_BUILD_EXACT;
#endif
  return f;
}

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
  const bool direct = m->tune.host_direct;
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
    const bool obs_direct = m->tune.direct_obs, sc_direct = m->tune.direct_scalars;
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
