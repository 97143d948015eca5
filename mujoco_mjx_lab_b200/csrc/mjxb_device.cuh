// mjxb_device.cuh -- the fused batched humanoid step for sm_100a: one warp per environment.
//
// Replaces the vmapped mjx.step / mjx.forward + env glue of reference src/envs.py:108-202,333-495 and
// mjx_humanoid_speed_test.py:48-57 (stage list: SURVEY.md Appendix B).  Design:
//   * lane <-> dof / body / geom pair / constraint row; no __syncthreads after the prologue, only __syncwarp;
//   * model constants staged once per CTA in shared memory; per-env working set ~13 KB of shared memory;
//   * only CANDIDATE constraint rows (pos < 0) are ever materialised (MJX keeps all 187, zero-masked);
//   * the Newton Hessian row of dof i lives in lane i's registers: assembled from 128-bit broadcast loads of the
//     Jacobian rows, factorised in registers (Cholesky, column broadcast through a 2x32-float smem buffer),
//     forward-substituted in registers; one code instance serves M, H and (M + h*damping);
//   * per-env early exit from the Newton and line-search loops (a vmapped while_loop cannot).
// FP32 CUDA cores only: nv = 27 contractions with data-dependent masks are not tensor-core shaped.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "mjxb.h"
#include "mjxb_model_dev.h"

namespace mjxb {

constexpr unsigned FULL = 0xffffffffu;

static_assert(sizeof(DevModel) % 16 == 0, "DevModel is staged into shared memory with 16-byte copies");

struct StepArgs {
  int n_env, mode, nsteps, autoreset;
  mjxb_state in, out;
  const float* action;   // [n, nu] (env modes: raw policy action; physics modes: ctrl) or NULL
  const uint32_t* keys;  // [n, 2]
  float *obs, *reward, *terminated, *truncated;
  uint8_t* reset_mask;
  int32_t* status;
  const float* vel;      // speed test
  float* pos;
  // overflow protocol (library-owned scratch): a pass iterates over `in_list` (NULL: all envs 0..n_env), appends the envs its
  // row tile cannot hold to `out_list` (NULL: none -- the last tier holds every static row), and resets its input list when done
  int* in_count; int* in_list; int* in_done;
  int* out_count; int* out_list;
  // deferred auto-reset: envs that finish an episode are queued per CTA and re-initialised in packed lockstep rounds after the
  // step rounds (a reset run inline would make the other 15 warps of its round wait for a whole extra pass)
  int* reset_list; int reset_stride;
  // concurrent auto-reset (batches of ONE round, <= 14 envs per SM): besides its `spec_reset` stepping warps a CTA runs reset warps
  // (blockDim / 32 - spec_reset of them). Whether an env finishes its episode in this step is known right after the stepping warp's
  // kinematics (termination reads the pelvis height of the step's forward pass, truncation the episode counter): it is published
  // then, and a reset warp re-initialises that env from its key (single_reset depends on the key alone) beside the step -- the reset
  // no longer runs as a second, serialised pass of the pipeline. Reset warps without work only answer the CTA barriers. More
  // finishing envs than reset warps in one CTA (rare): the surplus takes the deferred path below.
  int spec_reset;
  // work-sorted scheduling (large batches): `perm` lists the envs of each segment by descending predicted cost, so that the 16 envs a CTA
  // runs in lockstep need about the same number of Newton iterations; every env leaves its cost key (this step's iteration count) in
  // `work_out` for the next launch on the stream. A stale or missing hint only costs time: results do not depend on the order.
  const int* perm; uint8_t* work_out;
  // dynamic rounds (batches of several rounds without reset warps): a CTA takes its next group of envs from a device-wide counter when it
  // has finished a round, instead of the static (round, CTA) assignment -- CTAs whose rounds ran long take fewer of them. dyn_counter /
  // dyn_done live in the stream's scratch (zero between launches: the last CTA out resets them); a CTA takes at most dyn_max_rounds
  // groups (the capacity of its reset queue).
  int* dyn_counter; int* dyn_done; int dyn_max_rounds;
  int lockstep;          // CTA barriers keep the warps of an SM in the same code region (instruction-cache locality)
  int lockstep_group;    // warps per barrier group (0 = the whole CTA)
  // host-buffer pipeline: action / keys arrive in chunks of (1 << in_ready_shift) envs while the kernel already runs; the copy
  // stream writes in_ready_epoch into in_ready[chunk] after each chunk (NULL: every input is resident at launch)
  const unsigned* in_ready; int in_ready_shift; unsigned in_ready_epoch;
  unsigned* in_timeout;  // set to 1 (host-mapped) if a chunk did not arrive within the bounded wait
  mjxb_debug dbg;
};

// shared-memory vector slots (32 floats each)
enum { VQPOS = 0, VQVEL, VCTRL, VX, VY, VTMP, NVEC };

template <int CAP, int MAXCC>
struct __align__(16) WarpS {
  alignas(16) float M[NV * NVP];
  union {
    float L[NV * NV];  // Cholesky factor rows for the backward substitution
    struct {
      float xpos[MJXB_MAXBODY][3];
      float xquat[MJXB_MAXBODY][4];
      float cinert[MJXB_MAXBODY][10];  // later: composite inertia, in place
      float cvel[MJXB_MAXBODY][6];     // later: subtree-summed cfrc
      float cacc[MJXB_MAXBODY][6];     // later: body-local cfrc
    } a;
  };
  alignas(16) float J[CAP * NVP];
  float cdof[NV + 1][6];
  alignas(16) float vec[NVEC][32];
  alignas(16) float col[2][32];
  float rD[CAP], raref[CAP], rJaref[CAP], rjv[CAP], rforce[CAP];
  int rinfo[CAP];
  float gpos[MJXB_MAXGEOM][3], gaxis[MJXB_MAXGEOM][3];
  float cc_n[MAXCC][3], cc_t1[MAXCC][3], cc_t2[MAXCC][3], cc_pos[MAXCC][3], cc_dist[MAXCC];
  int cc_pair[MAXCC], cc_row[MAXCC];
  float site_xpos[MJXB_MAXSITE][3], site_xmat[MJXB_MAXSITE][9];
  float pelvis_pos[3], pelvis_quat[4], head_pos[3];
  float sens[MJXB_MAXSENSOR];
  float pad_[2];
};

// ------------------------------------------------------------------------------------------- small math
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
  return v;
}
__device__ __forceinline__ float dot3(const float* a, const float* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
__device__ __forceinline__ void cross3(float* r, const float* a, const float* b) {
  float x = a[1] * b[2] - a[2] * b[1], y = a[2] * b[0] - a[0] * b[2], z = a[0] * b[1] - a[1] * b[0];
  r[0] = x; r[1] = y; r[2] = z;
}
// x / (|x| + 1e-6*(|x|==0)), returns |x|   (mjx math.normalize_with_norm)
__device__ __forceinline__ float normalize3(float* a) {
  float n = sqrtf(dot3(a, a));
  float inv = 1.0f / (n + (n == 0.0f ? 1e-6f : 0.0f));
  a[0] *= inv; a[1] *= inv; a[2] *= inv;
  return n;
}
__device__ __forceinline__ void quat_mul(float* r, const float* a, const float* b) {
  float w = a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3];
  float x = a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2];
  float y = a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1];
  float z = a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0];
  r[0] = w; r[1] = x; r[2] = y; r[3] = z;
}
__device__ __forceinline__ void rotq(float* r, const float* v, const float* q) {  // mjx math.rotate
  float s = q[0];
  const float* u = q + 1;
  float uv = dot3(u, v), uu = dot3(u, u), c[3];
  cross3(c, u, v);
#pragma unroll
  for (int k = 0; k < 3; k++) r[k] = 2.0f * (uv * u[k]) + (s * s - uu) * v[k] + 2.0f * s * c[k];
}
__device__ __forceinline__ void quat_to_mat(float* m, const float* q) {
  float w = q[0], x = q[1], y = q[2], z = q[3];
  m[0] = w * w + x * x - y * y - z * z; m[1] = 2.0f * (x * y - w * z); m[2] = 2.0f * (x * z + w * y);
  m[3] = 2.0f * (x * y + w * z); m[4] = w * w - x * x + y * y - z * z; m[5] = 2.0f * (y * z - w * x);
  m[6] = 2.0f * (x * z - w * y); m[7] = 2.0f * (y * z + w * x); m[8] = w * w - x * x - y * y + z * z;
}
__device__ __forceinline__ void inert_mul(float* r, const float* i, const float* v) {
  float c1[3], c2[3];
  cross3(c1, i + 6, v + 3);
  cross3(c2, i + 6, v);
  r[0] = i[0] * v[0] + i[3] * v[1] + i[4] * v[2] + c1[0];
  r[1] = i[3] * v[0] + i[1] * v[1] + i[5] * v[2] + c1[1];
  r[2] = i[4] * v[0] + i[5] * v[1] + i[2] * v[2] + c1[2];
  r[3] = i[9] * v[3] - c2[0]; r[4] = i[9] * v[4] - c2[1]; r[5] = i[9] * v[5] - c2[2];
}
__device__ __forceinline__ void motion_cross(float* r, const float* u, const float* v) {
  float a[3], b[3], c[3];
  cross3(a, u, v); cross3(b, u, v + 3); cross3(c, u + 3, v);
#pragma unroll
  for (int k = 0; k < 3; k++) { r[k] = a[k]; r[3 + k] = b[k] + c[k]; }
}
__device__ __forceinline__ void motion_cross_force(float* r, const float* v, const float* f) {
  float a[3], b[3], c[3];
  cross3(a, v, f); cross3(b, v + 3, f + 3); cross3(c, v, f + 3);
#pragma unroll
  for (int k = 0; k < 3; k++) { r[k] = a[k] + b[k]; r[3 + k] = c[k]; }
}
__device__ __forceinline__ float clampf(float x, float lo, float hi) { return fminf(fmaxf(x, lo), hi); }

// ------------------------------------------------------------------------------------------- threefry / jax.random
__device__ __forceinline__ uint32_t rotl32(uint32_t v, int r) { return (v << r) | (v >> (32 - r)); }
__device__ __noinline__ uint2 threefry2x32v(uint32_t k0, uint32_t k1, uint32_t c0, uint32_t c1) {  // by value: reference outputs would force the callers' locals into local memory
  uint32_t ks0 = k0, ks1 = k1, ks2 = k0 ^ k1 ^ 0x1BD11BDAu;
  uint32_t x0 = c0 + ks0, x1 = c1 + ks1;
#define MJXB_TF_R(r) { x0 += x1; x1 = rotl32(x1, r); x1 ^= x0; }
  MJXB_TF_R(13) MJXB_TF_R(15) MJXB_TF_R(26) MJXB_TF_R(6)  x0 += ks1; x1 += ks2 + 1u;
  MJXB_TF_R(17) MJXB_TF_R(29) MJXB_TF_R(16) MJXB_TF_R(24) x0 += ks2; x1 += ks0 + 2u;
  MJXB_TF_R(13) MJXB_TF_R(15) MJXB_TF_R(26) MJXB_TF_R(6)  x0 += ks0; x1 += ks1 + 3u;
  MJXB_TF_R(17) MJXB_TF_R(29) MJXB_TF_R(16) MJXB_TF_R(24) x0 += ks1; x1 += ks2 + 4u;
  MJXB_TF_R(13) MJXB_TF_R(15) MJXB_TF_R(26) MJXB_TF_R(6)  x0 += ks2; x1 += ks0 + 5u;
#undef MJXB_TF_R
  return make_uint2(x0, x1);
}
__device__ __forceinline__ void threefry2x32(uint32_t k0, uint32_t k1, uint32_t c0, uint32_t c1, uint32_t& o0, uint32_t& o1) {
  const uint2 r = threefry2x32v(k0, k1, c0, c1);
  o0 = r.x; o1 = r.y;
}
// jax.random.uniform(key, (n,), f32, minval, maxval)[i] with threefry_partitionable
__device__ __forceinline__ float jax_uniform(uint32_t k0, uint32_t k1, uint32_t i, float minval, float maxval) {
  uint32_t a, b;
  threefry2x32(k0, k1, 0u, i, a, b);
  float f = __uint_as_float(((a ^ b) >> 9) | 0x3F800000u) - 1.0f;
  return fmaxf(minval, __fadd_rn(__fmul_rn(f, maxval - minval), minval));
}

// ------------------------------------------------------------------------------------------- constraint impedance (mjx constraint._kbi)
__device__ __noinline__ float2 kbi_general_power(float mid, float power, float x) {  // cold: solimp power not in {1, 2}
  return make_float2((1.0f / powf(mid, power - 1.0f)) * powf(x, power), 1.0f - (1.0f / powf(1.0f - mid, power - 1.0f)) * powf(1.0f - x, power));
}
__device__ __forceinline__ void kbi(float timestep, const float* solref, const float* solimp, float pos, float& k, float& b, float& imp) {
  float timeconst = fmaxf(solref[0], 2.0f * timestep), dampratio = solref[1];
  float dmin = clampf(solimp[0], 1e-4f, 0.9999f), dmax = clampf(solimp[1], 1e-4f, 0.9999f);
  float width = fmaxf(MINVAL, solimp[2]), mid = clampf(solimp[3], 1e-4f, 0.9999f), power = fmaxf(1.0f, solimp[4]);
  k = 1.0f / (dmax * dmax * timeconst * timeconst * dampratio * dampratio);
  b = 2.0f / (dmax * timeconst);
  if (solref[0] <= 0.0f) k = -solref[0] / (dmax * dmax);
  if (solref[1] <= 0.0f) b = -solref[1] / dmax;
  float x = fabsf(pos) / width;
  float ia, ib;
  if (power == 2.0f) { ia = (1.0f / mid) * (x * x); ib = 1.0f - (1.0f / (1.0f - mid)) * ((1.0f - x) * (1.0f - x)); }
  else if (power == 1.0f) { ia = x; ib = x; }
  else { const float2 r = kbi_general_power(mid, power, x); ia = r.x; ib = r.y; }
  float y = x < mid ? ia : ib;
  imp = clampf(dmin + y * (dmax - dmin), dmin, dmax);
  if (x > 1.0f) imp = dmax;
}

// ------------------------------------------------------------------------------------------- collision primitives (mjx collision_primitive / math)
__device__ __forceinline__ void sphere_sphere(float& dist, float* pos, float* n, const float* p1, float r1, const float* p2, float r2) {
  n[0] = p2[0] - p1[0]; n[1] = p2[1] - p1[1]; n[2] = p2[2] - p1[2];
  float len = normalize3(n);
  if (len == 0.0f) { n[0] = 1.0f; n[1] = 0.0f; n[2] = 0.0f; }
  dist = len - (r1 + r2);
  float s = r1 + dist * 0.5f;
  pos[0] = p1[0] + n[0] * s; pos[1] = p1[1] + n[1] * s; pos[2] = p1[2] + n[2] * s;
}
__device__ __forceinline__ void closest_segment_point(float* out, const float* a, const float* b, const float* pt) {
  float ab[3] = {b[0] - a[0], b[1] - a[1], b[2] - a[2]}, pa[3] = {pt[0] - a[0], pt[1] - a[1], pt[2] - a[2]};
  float t = clampf(dot3(pa, ab) / (dot3(ab, ab) + 1e-6f), 0.0f, 1.0f);
  out[0] = a[0] + t * ab[0]; out[1] = a[1] + t * ab[1]; out[2] = a[2] + t * ab[2];
}
__device__ __forceinline__ void closest_segment_to_segment(float* best_a, float* best_b, const float* a0, const float* a1,
                                                           const float* b0, const float* b1) {
  float dir_a[3] = {a1[0] - a0[0], a1[1] - a0[1], a1[2] - a0[2]}, dir_b[3] = {b1[0] - b0[0], b1[1] - b0[1], b1[2] - b0[2]};
  float half_a = normalize3(dir_a) * 0.5f, half_b = normalize3(dir_b) * 0.5f;
  float a_mid[3], b_mid[3], trans[3];
#pragma unroll
  for (int k = 0; k < 3; k++) {
    a_mid[k] = a0[k] + dir_a[k] * half_a;
    b_mid[k] = b0[k] + dir_b[k] * half_b;
    trans[k] = a_mid[k] - b_mid[k];
  }
  float dd = dot3(dir_a, dir_b), da_t = dot3(dir_a, trans), db_t = dot3(dir_b, trans);
  float denom = 1.0f - dd * dd;
  float orig_ta = (-da_t + dd * db_t) / (denom + 1e-6f);
  float orig_tb = db_t + orig_ta * dd;
  float ta = clampf(orig_ta, -half_a, half_a), tb = clampf(orig_tb, -half_b, half_b);
#pragma unroll
  for (int k = 0; k < 3; k++) { best_a[k] = a_mid[k] + dir_a[k] * ta; best_b[k] = b_mid[k] + dir_b[k] * tb; }
  float new_a[3], new_b[3];
  closest_segment_point(new_a, a0, a1, best_b);
  closest_segment_point(new_b, b0, b1, best_a);
  float d1 = 0.0f, d2 = 0.0f;
#pragma unroll
  for (int k = 0; k < 3; k++) {
    d1 += (new_a[k] - best_b[k]) * (new_a[k] - best_b[k]);
    d2 += (best_a[k] - new_b[k]) * (best_a[k] - new_b[k]);
  }
  if (d1 < d2) { best_a[0] = new_a[0]; best_a[1] = new_a[1]; best_a[2] = new_a[2]; }
  else { best_b[0] = new_b[0]; best_b[1] = new_b[1]; best_b[2] = new_b[2]; }
}
// mjx math.make_frame tangents for a unit normal
__device__ __forceinline__ void make_tangents(const float* n, float* t1, float* t2) {
  float b[3] = {0.0f, 0.0f, 0.0f};
  if (-0.5f < n[1] && n[1] < 0.5f) b[1] = 1.0f; else b[2] = 1.0f;
  float ab = dot3(n, b);
  b[0] -= n[0] * ab; b[1] -= n[1] * ab; b[2] -= n[2] * ab;
  normalize3(b);
  t1[0] = b[0]; t1[1] = b[1]; t1[2] = b[2];
  cross3(t2, n, b);
}
// nearest x >= 0 with pnt + x*vec on a face of the axis-aligned box `size`, else -1 (engine_ray.c ray_box / mjx ray._ray_box).
// One face pair; everything by value so that nothing is forced into local memory.
__device__ __forceinline__ float ray_box_axis(float best, float s_i, float p_i, float v_i, float s_a, float p_a, float v_a, float s_b, float p_b, float v_b) {
  if (fabsf(v_i) > MINVAL) {
#pragma unroll
    for (int side = -1; side <= 1; side += 2) {
      const float sol = ((float)side * s_i - p_i) / v_i;
      if (sol >= 0.0f) {
        const float qa = p_a + sol * v_a, qb = p_b + sol * v_b;
        if (fabsf(qa) <= s_a && fabsf(qb) <= s_b && (best < 0.0f || sol < best)) best = sol;
      }
    }
  }
  return best;
}
__device__ __forceinline__ float ray_box(float sx, float sy, float sz, float px, float py, float pz, float vx, float vy, float vz) {
  float best = -1.0f;
  best = ray_box_axis(best, sx, px, vx, sy, py, vy, sz, pz, vz);
  best = ray_box_axis(best, sy, py, vy, sx, px, vx, sz, pz, vz);
  best = ray_box_axis(best, sz, pz, vz, sx, px, vx, sy, py, vy);
  return best;
}

// ------------------------------------------------------------------------------------------- warp-level linear algebra on WarpS
// y_i = sum_j M[i][j] x_j ; x is published through vec[VX]
template <class WS>
__device__ __forceinline__ float matvec_M(WS& S, int lane, float x) {
  __syncwarp();
  S.vec[VX][lane] = (lane < NV) ? x : 0.0f;
  __syncwarp();
  float acc = 0.0f;
  if (lane < NV) {
    const float4* row = reinterpret_cast<const float4*>(&S.M[lane * NVP]);
    const float4* xv = reinterpret_cast<const float4*>(&S.vec[VX][0]);
#pragma unroll
    for (int g = 0; g < NVP / 4; g++) {
      float4 m = row[g], xx = xv[g];
      acc += m.x * xx.x; acc += m.y * xx.y; acc += m.z * xx.z; acc += m.w * xx.w;
    }
  }
  return acc;
}
// matvec_M and rows_times of the same x with one publication of x and the two 28-term dependent FMA chains of a lane (its row of M,
// its row of J) interleaved -- same per-chain order, so the results are bit-identical to the separate calls
template <class WS>
__device__ __forceinline__ float matvec_M_and_rows(WS& S, int lane, int nrow, float x, float* out /* smem [CAP] */) {
  __syncwarp();
  S.vec[VX][lane] = (lane < NV) ? x : 0.0f;
  __syncwarp();
  const float4* xv = reinterpret_cast<const float4*>(&S.vec[VX][0]);
  const float4* mrow = reinterpret_cast<const float4*>(&S.M[(lane < NV ? lane : 0) * NVP]);
  const float4* jrow = reinterpret_cast<const float4*>(&S.J[(lane < nrow ? lane : 0) * NVP]);
  float accm = 0.0f, accj = 0.0f;
#pragma unroll
  for (int g = 0; g < NVP / 4; g++) {
    const float4 xx = xv[g], m = mrow[g], j = jrow[g];
    accm += m.x * xx.x; accj += j.x * xx.x;
    accm += m.y * xx.y; accj += j.y * xx.y;
    accm += m.z * xx.z; accj += j.z * xx.z;
    accm += m.w * xx.w; accj += j.w * xx.w;
  }
  if (lane < nrow) out[lane] = accj;
  for (int r = 32 + lane; r < nrow; r += 32) {   // further strips (capacity tiers above 32 rows)
    const float4* row = reinterpret_cast<const float4*>(&S.J[r * NVP]);
    float acc = 0.0f;
#pragma unroll
    for (int g = 0; g < NVP / 4; g++) {
      float4 m = row[g], xx = xv[g];
      acc += m.x * xx.x; acc += m.y * xx.y; acc += m.z * xx.z; acc += m.w * xx.w;
    }
    out[r] = acc;
  }
  __syncwarp();
  return (lane < NV) ? accm : 0.0f;
}
// the same for two vectors at once (warm-start selection): M and J rows are loaded once, four dependent chains interleave
template <class WS>
__device__ __forceinline__ void matvec_M_and_rows2(WS& S, int lane, int nrow, float x1, float x2, float* out1, float* out2, float& m1,
                                                   float& m2) {
  __syncwarp();
  S.vec[VX][lane] = (lane < NV) ? x1 : 0.0f;
  S.vec[VTMP][lane] = (lane < NV) ? x2 : 0.0f;
  __syncwarp();
  const float4* xv = reinterpret_cast<const float4*>(&S.vec[VX][0]);
  const float4* yv = reinterpret_cast<const float4*>(&S.vec[VTMP][0]);
  const float4* mrow = reinterpret_cast<const float4*>(&S.M[(lane < NV ? lane : 0) * NVP]);
  float am1 = 0.0f, am2 = 0.0f;
#pragma unroll
  for (int g = 0; g < NVP / 4; g++) {
    const float4 xx = xv[g], yy = yv[g], m = mrow[g];
    am1 += m.x * xx.x; am2 += m.x * yy.x;
    am1 += m.y * xx.y; am2 += m.y * yy.y;
    am1 += m.z * xx.z; am2 += m.z * yy.z;
    am1 += m.w * xx.w; am2 += m.w * yy.w;
  }
  for (int r = lane; r < nrow; r += 32) {
    const float4* row = reinterpret_cast<const float4*>(&S.J[r * NVP]);
    float a1 = 0.0f, a2 = 0.0f;
#pragma unroll
    for (int g = 0; g < NVP / 4; g++) {
      const float4 j = row[g], xx = xv[g], yy = yv[g];
      a1 += j.x * xx.x; a2 += j.x * yy.x;
      a1 += j.y * xx.y; a2 += j.y * yy.y;
      a1 += j.z * xx.z; a2 += j.z * yy.z;
      a1 += j.w * xx.w; a2 += j.w * yy.w;
    }
    out1[r] = a1; out2[r] = a2;
  }
  __syncwarp();
  m1 = (lane < NV) ? am1 : 0.0f;
  m2 = (lane < NV) ? am2 : 0.0f;
}
// out[r] = sum_d J[r][d] x_d for all rows (lane-per-row strips); x published through vec[VX]
template <class WS>
__device__ __forceinline__ void rows_times(WS& S, int lane, int nrow, float x, float* out /* smem [CAP] */) {
  __syncwarp();
  S.vec[VX][lane] = (lane < NV) ? x : 0.0f;
  __syncwarp();
  const float4* xv = reinterpret_cast<const float4*>(&S.vec[VX][0]);
  for (int r = lane; r < nrow; r += 32) {
    const float4* row = reinterpret_cast<const float4*>(&S.J[r * NVP]);
    float acc = 0.0f;
#pragma unroll
    for (int g = 0; g < NVP / 4; g++) {
      float4 m = row[g], xx = xv[g];
      acc += m.x * xx.x; acc += m.y * xx.y; acc += m.z * xx.z; acc += m.w * xx.w;
    }
    out[r] = acc;
  }
  __syncwarp();
}
// y_d = sum_r J[r][d] f_r
template <class WS>
__device__ __forceinline__ float rowsT_times(const WS& S, int lane, int nrow, const float* f /* smem [CAP] */) {
  float acc = 0.0f;
  if (lane < NVP) {
    for (int r = 0; r < nrow; r++) acc += S.J[r * NVP + lane] * f[r];
  }
  return acc;
}

// In-register Cholesky of the SPD matrix whose row `lane` is a[0..NV) (lower part used). On return a[k] (k<lane) = L[lane][k],
// dinv = 1 / L[lane][lane]. Column k is broadcast through S.col (double buffered): 27 SHFL + 27 STS + ~100 LDS.128.
template <class WS>
__device__ __forceinline__ void chol_rows(WS& S, int lane, float (&a)[NVP], float& dinv) {
  dinv = 1.0f;
#pragma unroll
  for (int k = 0; k < NV; k++) {
    float akk = __shfl_sync(FULL, a[k], k);
    float inv = rsqrtf(akk);
    float lik = a[k] * inv;          // valid for lane > k; lane k gets sqrt(akk)
    if (lane == k) dinv = inv;
    a[k] = lik;
    if (k + 1 < NV) {
      float* col = S.col[k & 1];
      col[lane] = lik;
      __syncwarp();
      const float4* cv = reinterpret_cast<const float4*>(col);
#pragma unroll
      for (int g = (k + 1) / 4; g < NVP / 4; g++) {
        float4 c = cv[g];
        if (4 * g + 0 > k && 4 * g + 0 < NV) a[4 * g + 0] -= lik * c.x;
        if (4 * g + 1 > k && 4 * g + 1 < NV) a[4 * g + 1] -= lik * c.y;
        if (4 * g + 2 > k && 4 * g + 2 < NV) a[4 * g + 2] -= lik * c.z;
        if (4 * g + 3 > k && 4 * g + 3 < NV) a[4 * g + 3] -= lik * c.w;
      }
    }
  }
}
// solve (L L^T) x = b with the factor from chol_rows; L rows are spilled to S.L for the backward pass
template <class WS>
__device__ __forceinline__ float chol_solve_rows(WS& S, int lane, const float (&a)[NVP], float dinv, float b) {
  float y = b;
#pragma unroll
  for (int k = 0; k < NV; k++) {
    float yk = __shfl_sync(FULL, y * dinv, k);
    if (lane > k) y -= a[k] * yk;
  }
  y *= dinv;
  __syncwarp();
  if (lane < NV) {
#pragma unroll
    for (int k = 0; k < NV; k++)
      if (k < lane) S.L[lane * NV + k] = a[k];
  }
  __syncwarp();
  float x = y;
#pragma unroll
  for (int k = NV - 1; k >= 0; k--) {
    float xk = __shfl_sync(FULL, x * dinv, k);
    if (lane < k) x -= S.L[k * NV + lane] * xk;
  }
  return x * dinv;
}

struct LSPoint { float alpha, cost, d0, d1; };

}  // namespace mjxb
// Profiling aid (tools/stage_clock.py builds a variant with -DMJXB_STAGE_CLOCK=1): per-env clock() stamps at the stage boundaries
#ifndef MJXB_STAGE_CLOCK
#define MJXB_STAGE_CLOCK 0
#endif
#if MJXB_STAGE_CLOCK
__device__ int g_stage_clock[4096 * 32];
#define MJXB_STAMP(i) do { if (lane == 0 && valid && env < 4096) g_stage_clock[env * 32 + (i)] = (int)clock(); } while (0)
#else
#define MJXB_STAMP(i) do { } while (0)
#endif
#ifndef MJXB_WORK_KEY_ITER
// cost key of the work-sorted schedule = this step's Newton iterations * ITER + candidate rows * ROW. The candidate-row count of the
// previous step predicts the next step's iteration count better than the previous iteration count does (tools/niter_predict_dump.py on the
// trajectory distribution: mean over groups of 16 of the largest count 5.86 unsorted, 5.57 by iterations, 5.40 by rows or by this key,
// 5.2 for a boosted-tree regressor on every previous-step feature, 3.04 with hindsight), and it also sorts by the cost of an iteration.
#define MJXB_WORK_KEY_ITER 4
#define MJXB_WORK_KEY_ROW 2
#endif
#ifndef MJXB_FACTOR_REUSE
#define MJXB_FACTOR_REUSE 1
#endif
#include "mjxb_chol_tree.cuh"
namespace mjxb {

// barrier over a group of `g` consecutive warps (named barrier 1 + group index); g <= 0 or g >= CTA: the whole CTA
__device__ __forceinline__ void group_sync(int warp, int g) {
  const int nwarp = blockDim.x >> 5;
  if (g <= 0 || g >= nwarp) { __syncthreads(); return; }
  const int grp = warp / g;
  const int cnt = min(g, nwarp - grp * g) * 32;
  asm volatile("bar.sync %0, %1;" ::"r"(grp + 1), "r"(cnt) : "memory");
}

// Segmented counting sort of the envs by descending cost key (one CTA per segment of 2^seg_shift envs, 256 bins): perm[lo .. hi) = the envs
// of the segment, heaviest first. The order inside a bin is arbitrary (atomics) -- nothing downstream depends on it.
__global__ void __launch_bounds__(1024) mjxb_sort_work_kernel(const uint8_t* __restrict__ work, int* __restrict__ perm, int n, int seg_shift) {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
  __shared__ int hist[256];
  __shared__ int base[256];
  const int lo = blockIdx.x << seg_shift, hi = min(n, lo + (1 << seg_shift)), lane = threadIdx.x & 31;
  if (threadIdx.x < 256) hist[threadIdx.x] = 0;
  __syncthreads();
  for (int i0 = lo + (threadIdx.x & ~31); i0 < hi; i0 += blockDim.x) {   // warp-uniform trip count: match_any needs the whole warp
    const int i = i0 + lane;
    const int key = i < hi ? work[i] : 256 + lane;                       // out-of-range lanes get distinct dummy keys
    const unsigned peers = __match_any_sync(FULL, key);
    if (i < hi && lane == __ffs(peers) - 1) atomicAdd(&hist[key], __popc(peers));
  }
  __syncthreads();
  if (threadIdx.x < 32) {   // exclusive prefix over the bins in descending key order
    int run = 0;
    for (int b0 = 0; b0 < 256; b0 += 32) {
      const int bin = 255 - (b0 + lane);
      const int c = hist[bin];
      int inc = c;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(FULL, inc, o); if (lane >= o) inc += t; }
      base[bin] = run + inc - c;
      run += __shfl_sync(FULL, inc, 31);
    }
  }
  __syncthreads();
  for (int i0 = lo + (threadIdx.x & ~31); i0 < hi; i0 += blockDim.x) {
    const int i = i0 + lane;
    const int key = i < hi ? work[i] : 256 + lane;
    const unsigned peers = __match_any_sync(FULL, key);
    const int leader = __ffs(peers) - 1;
    int pos = 0;
    if (i < hi && lane == leader) pos = atomicAdd(&base[key], __popc(peers));
    pos = __shfl_sync(FULL, pos, leader) + __popc(peers & ((1u << lane) - 1u));
    if (i < hi) perm[lo + pos] = i;
  }
}

// lane-per-row loop over the candidate rows with a compile-time trip bound (NSTRIP = ceil(CAP / 32), 1 for the main tier): a plain
// `for (r = lane; r < nrow; r += 32)` is unrolled eightfold by the compiler with branchy prologues although it runs exactly once
#define MJXB_FOR_ROW_STRIPS(r) _Pragma("unroll") for (int r = lane, strip_ = 0; strip_ < NSTRIP && r < nrow; strip_++, r += 32)
// ------------------------------------------------------------------------------------------- the kernel
// SINGLE: one step per launch and resets always deferred -> no loop-carried per-env registers across the step / pass loops.
// DYN: groups of envs are taken from a device-wide counter (StepArgs::dyn_counter) instead of the static (round, CTA) assignment.
template <bool DBG, int CAP, int MAXCC, int MAXW, bool LS_EXACT, bool SINGLE = false, bool DYN = false>
__global__ void __launch_bounds__(MAXW * 32, 1) mjxb_step_kernel(const DevModel* __restrict__ gmodel, const PairParam* __restrict__ pair_param,
                                                            StepArgs A) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  // programmatic dependent launch (no-ops in a plain launch): let the successor be staged at once; wait for the predecessor's results
  // only after the work that does not depend on them (staging the model constants) -- except in an overflow tier, whose first read
  // (its input count) is a result of the predecessor
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  if (A.in_list != nullptr) {
    asm volatile("griddepcontrol.wait;" ::: "memory");
    // an overflow tier whose input list is empty (the common case) leaves before it touches shared memory
    if (*reinterpret_cast<volatile int*>(A.in_count) == 0) return;
  }
  DevModel& C = *reinterpret_cast<DevModel*>(smem_raw);
  {
    const int4* src = reinterpret_cast<const int4*>(gmodel);
    int4* dst = reinterpret_cast<int4*>(smem_raw);
    for (int i = threadIdx.x; i < (int)(sizeof(DevModel) / 16); i += blockDim.x) dst[i] = src[i];
  }
  if (A.in_list == nullptr) asm volatile("griddepcontrol.wait;" ::: "memory");
  __shared__ int s_reset_count;
  __shared__ int s_spec_done[32];
  __shared__ int s_spec_need[32];   // per stepping slot: 0 = the env goes on, 1 = it finishes and a reset warp serves it, 2 = finishes, unserved
  __shared__ int s_need_slot[32];   // the slots that need a reset, in order of publication
  __shared__ int s_pub_count, s_need_count;
  __shared__ int s_dyn_base;
  if (threadIdx.x == 0) { s_reset_count = 0; s_pub_count = 0; s_need_count = 0; }
  if (threadIdx.x < 32) s_spec_need[threadIdx.x] = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  using WS = WarpS<CAP, MAXCC>;
  constexpr int NSTRIP = (CAP + 31) / 32;
  WS& S = *reinterpret_cast<WS*>(smem_raw + ((sizeof(DevModel) + 15) & ~size_t(15)) + (size_t)warp * sizeof(WS));
  const unsigned lt_mask = (1u << lane) - 1u;
  const float h = C.timestep;
  const mjxb_env_config& cfg = C.cfg;
  const int nbody = C.nbody;

  const bool consuming = A.in_list != nullptr;
  const int n_items = consuming ? *reinterpret_cast<volatile int*>(A.in_count) : A.n_env;
  const bool spec = SINGLE && A.spec_reset != 0 && A.autoreset && A.mode == MODE_ENV_STEP && !consuming;   // CTA-uniform
  const int nslot = spec ? min(A.spec_reset, nwarp - 1) : nwarp;   // stepping warps = envs per CTA round
  int wslot = warp;                                        // (a reset warp takes the slot of the env it serves)
  const bool spec_warp = spec && warp >= nslot;            // this warp is a reset warp of the step round
  const int n_rounds = (n_items + gridDim.x * nslot - 1) / (gridDim.x * nslot);
  const bool defer = A.autoreset && A.reset_list != nullptr && A.mode == MODE_ENV_STEP;
  constexpr bool dyn = DYN;                                // (the host launches this instantiation only without reset warps / input lists)
  int total_rounds = n_rounds, n_reset = 0, reset_round0 = n_rounds;
  bool reset_phase = false;
  for (int round = 0;; round++) {
    if (dyn && !reset_phase) {   // take the next group of envs
      if (threadIdx.x == 0) s_dyn_base = round < A.dyn_max_rounds ? atomicAdd(A.dyn_counter, nslot) : n_items;
      __syncthreads();
      total_rounds = s_dyn_base < n_items ? round + 1 : round;
    }
    if (round == total_rounds) {  // CTA-uniform: after the step rounds, the queued resets run 16 per round
      if (reset_phase || !defer) break;
      __syncthreads();
      n_reset = s_reset_count;
      reset_phase = true;
      reset_round0 = round;
      total_rounds += (n_reset + nwarp - 1) / nwarp;
      if (round == total_rounds) break;
    }
    if (A.lockstep > 0 || spec) group_sync(warp, spec ? 0 : A.lockstep_group);
    const bool spec_partner = spec_warp && !reset_phase;
    if (spec_partner) {
      // wait until every stepping warp of the CTA has published its verdict (~3 K cycles in: right after its kinematics), then take
      // the (warp - nslot)-th env that needs a reset, if there is one; otherwise stay out of the SM's way and only answer the CTA
      // barriers of the round (1024 envs: 7 instead of 14 working warps per SM)
      int served = -1;
      if (lane == 0) {
        volatile int* pub = &s_pub_count;
#pragma unroll 1
        for (int spin = 0; spin < (1 << 20) && *pub < nslot; spin++) __nanosleep(200);
        __threadfence_block();
        const int k = warp - nslot;
        if (k < *reinterpret_cast<volatile int*>(&s_need_count)) served = *reinterpret_cast<volatile int*>(&s_need_slot[k]);
      }
      served = __shfl_sync(FULL, served, 0);
      if (served < 0) {
        if (A.lockstep > 0) group_sync(warp, A.lockstep_group);                                   // solver entry
        if (A.lockstep == 3) { while (__syncthreads_or(0)) {} }                                   // per-iteration barriers / solver exit
        else if (A.lockstep > 0 && A.lockstep != 2) group_sync(warp, A.lockstep_group);
        __syncthreads();                                                                          // the verdict barrier of the round's tail
        continue;
      }
      wslot = served;
    }
    // Warps without work in the last round (and envs that overflow the row tile) still run the whole pipeline -- on env 0 /
    // on a truncated row set -- with every global store suppressed, so that all warps of the CTA reach the same barriers.
    bool valid;
    int env;
    if (!reset_phase) {
      // sorted schedule: odd rounds deal the groups to the CTAs in reverse, so that a CTA's heavy groups are paired with light ones
      const int cta = (A.perm != nullptr && (round & 1)) ? (int)gridDim.x - 1 - (int)blockIdx.x : (int)blockIdx.x;
      const int item = dyn ? s_dyn_base + wslot : (round * gridDim.x + cta) * nslot + wslot;
      valid = item < n_items;
      env = valid ? (consuming ? A.in_list[item] : (A.perm != nullptr ? A.perm[item] : item)) : (consuming ? A.in_list[0] : 0);
    } else {
      const int idx = (round - reset_round0) * nwarp + warp;
      valid = idx < n_reset;
      env = A.reset_list[(size_t)blockIdx.x * A.reset_stride + (valid ? idx : 0)];
    }
    bool overflow = false, deferred = false;
    // ---------------------------------------------------------------- load state (lane d <-> qpos[d], qvel[d], ...)
    int mode = (reset_phase || spec_partner) ? MODE_ENV_RESET : A.mode;
    // Per-env values are NOT carried in registers across the pipeline (128 registers per thread, and the 24 KB of L1 left beside
    // 230 KB of shared memory cannot hold spills): they are published to shared memory here and re-read where they are used.
    float q = 0.0f, v = 0.0f, ws = 0.0f, ctrl = 0.0f, tm = 0.0f, aux = 0.0f;
    int status = consuming ? MJXB_STATUS_ROW_SPILL : 0;
    if (A.in_ready != nullptr && !reset_phase) {   // (the speculative partner waits too: its key arrives with the chunk)  // wait for this env's chunk of action / keys (copy engine -> flag, L2-coherent)
      if (lane == 0) {
        const unsigned* flag = A.in_ready + (env >> A.in_ready_shift);
#pragma unroll 1
        for (int spin = 0; spin < (1 << 22); spin++) {  // bounded (~2 s): a lost copy must not hang the device
          unsigned seen;
          asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(seen) : "l"(flag) : "memory");
          if ((int)(seen - A.in_ready_epoch) >= 0) break;
          __nanosleep(500);
          if (spin == (1 << 22) - 1 && A.in_timeout) *A.in_timeout = 1u;
        }
      }
      __syncwarp();
    }
    {
      float action = 0.0f;
      if (mode == MODE_ENV_STEP || mode == MODE_PHYS_STEP || mode == MODE_FORWARD) {
        if (lane < C.nq) q = A.in.qpos[(size_t)env * C.nq + lane];
        if (lane < NV) { v = A.in.qvel[(size_t)env * NV + lane]; ws = A.in.qacc_warmstart[(size_t)env * NV + lane]; }
        tm = A.in.time[env];
        if (A.action != nullptr && lane < C.nu) action = __ldcg(A.action + (size_t)env * C.nu + lane);
      }
      if (mode == MODE_ENV_STEP) {
        const float flip = A.in.aux[(size_t)env * MJXB_AUX_DIM];
        // src/envs.py:339-341 flip + clip
        int src = lane < C.nu ? cfg.act_perm[lane] : 0;
        float af = __shfl_sync(FULL, action, src) * (lane < C.nu ? cfg.act_sign[lane] : 0.0f);
        ctrl = clampf(flip > 0.5f ? af : action, -1.0f, 1.0f);
      } else {
        ctrl = action;
      }
    }
    // env-step outputs latched before an in-kernel auto-reset overwrites the state
    float out_reward = 0.0f, out_term = 0.0f, out_trunc = 0.0f;
    float tgt_x = 0.0f, tgt_y = 0.0f, tgt_z = 0.0f, flip_r = 0.0f, vmag_r = 0.0f;
    int did_reset = 0;

    const int nsteps = SINGLE ? 1 : A.nsteps;
    for (int istep = 0; istep < nsteps; istep++) {
      if (mode == MODE_SPEED_TEST) {  // mjx_humanoid_speed_test.py:50-53: make_data, qvel[0] = vel
        q = lane < C.nq ? C.qpos0[lane] : 0.0f;
        v = (lane == 0) ? A.vel[env] : 0.0f;
        ws = 0.0f; ctrl = 0.0f; tm = 0.0f;
      }
      for (int pass = 0; pass < (SINGLE ? 1 : 2); pass++) {  // pass 1 only for the inline auto-reset (no CTA barriers there)
      if (mode == MODE_ENV_RESET) {
        // ------------------------------------------------------------ single_reset state init (src/envs.py:117-131,147)
        const uint32_t key0 = __ldcg(A.keys + 2 * (size_t)env), key1 = __ldcg(A.keys + 2 * (size_t)env + 1);
        uint32_t k1a, k1b, k2a, k2b, k3a, k3b, k4a, k4b;
        threefry2x32(key0, key1, 0u, 0u, k1a, k1b);
        threefry2x32(key0, key1, 0u, 1u, k2a, k2b);
        threefry2x32(key0, key1, 0u, 2u, k3a, k3b);
        threefry2x32(key0, key1, 0u, 3u, k4a, k4b);
        flip_r = 0.0f;
        if (cfg.random_flip) flip_r = jax_uniform(k3a, k3b, 0u, 0.0f, 1.0f) < 0.5f ? 1.0f : 0.0f;
        q = lane < C.nq ? C.qpos0[lane] : 0.0f;
        if (lane >= 7 && lane < C.nq) {
          float nz = __fadd_rn(__fmul_rn(jax_uniform(k1a, k1b, (uint32_t)(lane - 7), 0.0f, 1.0f), 2.0f), -1.0f);
          q = __fadd_rn(q, __fmul_rn(cfg.random_joint_noise, nz));
        }
        v = 0.0f;
        if (lane < NV) {
          float nz = __fadd_rn(__fmul_rn(jax_uniform(k2a, k2b, (uint32_t)lane, 0.0f, 1.0f), 2.0f), -1.0f);
          v = __fmul_rn(cfg.random_vel_noise, nz);
        }
        if (cfg.initial_velocity_max > 0.0f) {
          // target = pelvis + (target_dist, 0): direction is +x up to rounding; vx = vmag*dx/|dx|, vy = vmag*0/|dx|
          float vmag = jax_uniform(k4a, k4b, 0u, 0.0f, cfg.initial_velocity_max);
          // dx = (bx + target_dist) - bx is evaluated after kinematics below (needs pelvis x)
          vmag_r = vmag;
        }
        ws = 0.0f; ctrl = 0.0f; tm = 0.0f;
      }
      // publish state vectors
      __syncwarp();
      S.vec[VQPOS][lane] = q;
      S.vec[VQVEL][lane] = (lane < NV) ? v : 0.0f;
      S.vec[VX][lane] = (lane < NV) ? ws : 0.0f;   // warm start, read back by the solver (VX is scratch from the solver on)
      if (istep == 0 || mode == MODE_ENV_RESET) S.vec[VCTRL][lane] = ctrl;  // constant over the in-kernel steps
      __syncwarp();

      MJXB_STAMP(0);
      // ---------------------------------------------------------------- kinematics: parallel prefix over the joint tree
      // Lane j holds the rigid transform of joint j relative to the frame before it (body offset folded into a body's first joint);
      // log2(depth) rounds of pointer jumping with shuffles compose it up the chain to the world frame after the joint. The joint
      // anchor and axis are invariant under the joint's own rotation, so they are read off that frame (mjx smooth.kinematics).
      {
        float ep[3] = {0.0f, 0.0f, 0.0f}, eq[4] = {1.0f, 0.0f, 0.0f, 0.0f};
        int jp = -1;
        if (lane < C.njnt) {
          const int j = lane, qa = C.jnt_qposadr[j];
          jp = C.jnt_parent[j];
          if (C.jnt_type[j] == 0) {
            ep[0] = S.vec[VQPOS][qa]; ep[1] = S.vec[VQPOS][qa + 1]; ep[2] = S.vec[VQPOS][qa + 2];
            float w = S.vec[VQPOS][qa + 3], x = S.vec[VQPOS][qa + 4], y = S.vec[VQPOS][qa + 5], z = S.vec[VQPOS][qa + 6];
            float n = sqrtf(w * w + x * x + y * y + z * z);
            float dn = 1.0f / (n + (n == 0.0f ? 1e-6f : 0.0f));
            eq[0] = w * dn; eq[1] = x * dn; eq[2] = y * dn; eq[3] = z * dn;
            S.vec[VQPOS][qa + 3] = eq[0]; S.vec[VQPOS][qa + 4] = eq[1]; S.vec[VQPOS][qa + 5] = eq[2]; S.vec[VQPOS][qa + 6] = eq[3];
          } else {
            float ang = S.vec[VQPOS][qa] - C.qpos0[qa], sn, cs, t[3];
            sincosf(ang * 0.5f, &sn, &cs);
            float ql[4] = {cs, C.jnt_axis[j][0] * sn, C.jnt_axis[j][1] * sn, C.jnt_axis[j][2] * sn};
            rotq(t, C.jnt_pos[j], ql);
            float dp[3] = {C.jnt_pos[j][0] - t[0], C.jnt_pos[j][1] - t[1], C.jnt_pos[j][2] - t[2]};  // rotation about the anchor
            if (C.jnt_first[j]) {
              const int b = C.jnt_bodyid[j];
              rotq(t, dp, C.body_quat[b]);
              ep[0] = C.body_pos[b][0] + t[0]; ep[1] = C.body_pos[b][1] + t[1]; ep[2] = C.body_pos[b][2] + t[2];
              quat_mul(eq, C.body_quat[b], ql);
            } else {
              ep[0] = dp[0]; ep[1] = dp[1]; ep[2] = dp[2];
              eq[0] = ql[0]; eq[1] = ql[1]; eq[2] = ql[2]; eq[3] = ql[3];
            }
          }
        }
        for (int st = 0; st < C.tree_steps; st++) {
          const int src = jp < 0 ? 0 : jp;
          float P[3], Q[4];
#pragma unroll
          for (int k = 0; k < 3; k++) P[k] = __shfl_sync(FULL, ep[k], src);
#pragma unroll
          for (int k = 0; k < 4; k++) Q[k] = __shfl_sync(FULL, eq[k], src);
          const int jp2 = __shfl_sync(FULL, jp, src);
          if (jp >= 0) {
            float t[3], qn[4];
            rotq(t, ep, Q);
            ep[0] = P[0] + t[0]; ep[1] = P[1] + t[1]; ep[2] = P[2] + t[2];
            quat_mul(qn, Q, eq);
            eq[0] = qn[0]; eq[1] = qn[1]; eq[2] = qn[2]; eq[3] = qn[3];
            jp = jp2;
          }
        }
        if (lane < C.njnt && C.jnt_type[lane] != 0) {  // hinge: stash world axis / anchor in its cdof slot (finalised after com)
          const int da = C.jnt_dofadr[lane];
          float t[3], axis[3];
          rotq(t, C.jnt_pos[lane], eq);
          rotq(axis, C.jnt_axis[lane], eq);
          S.cdof[da][0] = axis[0]; S.cdof[da][1] = axis[1]; S.cdof[da][2] = axis[2];
          S.cdof[da][3] = ep[0] + t[0]; S.cdof[da][4] = ep[1] + t[1]; S.cdof[da][5] = ep[2] + t[2];
        }
        {  // body frames: the frame after the body's last joint (or of the nearest jointed ancestor) composed with a fixed offset
          const int sj = lane < nbody ? C.body_srcjnt[lane] : -1;
          const int src = sj < 0 ? 0 : sj;
          float P[3], Q[4];
#pragma unroll
          for (int k = 0; k < 3; k++) P[k] = __shfl_sync(FULL, ep[k], src);
#pragma unroll
          for (int k = 0; k < 4; k++) Q[k] = __shfl_sync(FULL, eq[k], src);
          if (sj < 0) { P[0] = P[1] = P[2] = 0.0f; Q[0] = 1.0f; Q[1] = Q[2] = Q[3] = 0.0f; }
          if (lane < nbody) {
            float t[3], qn[4];
            rotq(t, C.body_relpos[lane], Q);
            quat_mul(qn, Q, C.body_relquat[lane]);
            S.a.xpos[lane][0] = P[0] + t[0]; S.a.xpos[lane][1] = P[1] + t[1]; S.a.xpos[lane][2] = P[2] + t[2];
            S.a.xquat[lane][0] = qn[0]; S.a.xquat[lane][1] = qn[1]; S.a.xquat[lane][2] = qn[2]; S.a.xquat[lane][3] = qn[3];
          }
        }
        __syncwarp();
      }

      if (mode == MODE_ENV_RESET && cfg.initial_velocity_max > 0.0f) {  // src/envs.py:136-152
        float bx = S.a.xpos[cfg.pelvis_body_id][0], by = S.a.xpos[cfg.pelvis_body_id][1];
        float tx = bx + cfg.target_dist, ty = by;
        float dx = tx - bx, dy = ty - by;
        float dist_xy = sqrtf(dx * dx + dy * dy);
        float vmag = vmag_r;
        float vx = dist_xy > 1e-6f ? __fdiv_rn(__fmul_rn(vmag, dx), dist_xy) : 0.0f;
        float vy = dist_xy > 1e-6f ? __fdiv_rn(__fmul_rn(vmag, dy), dist_xy) : 0.0f;
        __syncwarp();
        if (lane == 0) S.vec[VQVEL][0] = vx;
        if (lane == 1) S.vec[VQVEL][1] = vy;
        __syncwarp();
      }

      MJXB_STAMP(1);
      // ---------------------------------------------------------------- geoms, sites, frames the env layer reads
      if (lane < C.ngeom) {
        const int b = C.geom_body[lane];
        float bq[4] = {S.a.xquat[b][0], S.a.xquat[b][1], S.a.xquat[b][2], S.a.xquat[b][3]}, t[3];
        rotq(t, C.geom_pos[lane], bq);
        S.gpos[lane][0] = S.a.xpos[b][0] + t[0]; S.gpos[lane][1] = S.a.xpos[b][1] + t[1]; S.gpos[lane][2] = S.a.xpos[b][2] + t[2];
        rotq(t, C.geom_axis[lane], bq);
        S.gaxis[lane][0] = t[0]; S.gaxis[lane][1] = t[1]; S.gaxis[lane][2] = t[2];
      }
      if (lane < C.nsite) {
        const int b = C.site_body[lane];
        float bq[4] = {S.a.xquat[b][0], S.a.xquat[b][1], S.a.xquat[b][2], S.a.xquat[b][3]}, t[3], sq[4];
        rotq(t, C.site_pos[lane], bq);
        S.site_xpos[lane][0] = S.a.xpos[b][0] + t[0]; S.site_xpos[lane][1] = S.a.xpos[b][1] + t[1]; S.site_xpos[lane][2] = S.a.xpos[b][2] + t[2];
        quat_mul(sq, bq, C.site_quat[lane]);
        quat_to_mat(S.site_xmat[lane], sq);
      }
      if (lane == 31) {
        const int pb = cfg.pelvis_body_id, hb = cfg.head_body_id;
#pragma unroll
        for (int k = 0; k < 3; k++) { S.pelvis_pos[k] = S.a.xpos[pb][k]; S.head_pos[k] = S.a.xpos[hb][k]; }
#pragma unroll
        for (int k = 0; k < 4; k++) S.pelvis_quat[k] = S.a.xquat[pb][k];
        if (spec && !spec_warp && !reset_phase) {   // the verdict for the reset warps: the same predicates as the env layer evaluates at the end
          const bool fallen = S.a.xpos[pb][2] < cfg.terminate_height;
          const float ep1 = A.in.aux[(size_t)env * MJXB_AUX_DIM + 8] + 1.0f;
          const bool trunc = cfg.max_episode_steps > 0 && ep1 >= (float)cfg.max_episode_steps;
          int need = 0;
          if (valid && mode == MODE_ENV_STEP && (fallen || trunc)) {
            const int k = atomicAdd(&s_need_count, 1);
            s_need_slot[k] = wslot;
            need = k < nwarp - nslot ? 1 : 2;           // more finishing envs than reset warps: deferred reset after the round
          }
          s_spec_need[wslot] = need;
          __threadfence_block();
          atomicAdd(&s_pub_count, 1);
        }
      }
      if (DBG) {
        if (A.dbg.xpos && lane < nbody) for (int k = 0; k < 3; k++) A.dbg.xpos[((size_t)env * nbody + lane) * 3 + k] = S.a.xpos[lane][k];
        if (A.dbg.xquat && lane < nbody) for (int k = 0; k < 4; k++) A.dbg.xquat[((size_t)env * nbody + lane) * 4 + k] = S.a.xquat[lane][k];
      }

      MJXB_STAMP(2);
      // ---------------------------------------------------------------- com_pos: subtree com of the single tree, cinert, cdof
      float com[3];
      {
        float xip[3] = {0.0f, 0.0f, 0.0f}, mass = 0.0f, R[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
        if (lane < nbody) {
          quat_to_mat(R, S.a.xquat[lane]);
          const float* ip = C.body_ipos[lane];
          xip[0] = S.a.xpos[lane][0] + R[0] * ip[0] + R[1] * ip[1] + R[2] * ip[2];
          xip[1] = S.a.xpos[lane][1] + R[3] * ip[0] + R[4] * ip[1] + R[5] * ip[2];
          xip[2] = S.a.xpos[lane][2] + R[6] * ip[0] + R[7] * ip[1] + R[8] * ip[2];
          mass = C.body_mass[lane];
        }
        float inv_m = 1.0f / fmaxf(C.total_mass, MINVAL);
        com[0] = warp_sum(mass * xip[0]) * inv_m;
        com[1] = warp_sum(mass * xip[1]) * inv_m;
        com[2] = warp_sum(mass * xip[2]) * inv_m;
        if (lane < nbody) {
          float off[3] = {xip[0] - com[0], xip[1] - com[1], xip[2] - com[2]};
          const float* bi = C.body_inertia[lane];
          float I[9] = {bi[0], bi[3], bi[4], bi[3], bi[1], bi[5], bi[4], bi[5], bi[2]}, XI[9], Iw[9];
#pragma unroll
          for (int r = 0; r < 3; r++)
#pragma unroll
            for (int c = 0; c < 3; c++) XI[3 * r + c] = R[3 * r] * I[c] + R[3 * r + 1] * I[3 + c] + R[3 * r + 2] * I[6 + c];
#pragma unroll
          for (int r = 0; r < 3; r++)
#pragma unroll
            for (int c = 0; c < 3; c++) Iw[3 * r + c] = XI[3 * r] * R[3 * c] + XI[3 * r + 1] * R[3 * c + 1] + XI[3 * r + 2] * R[3 * c + 2];
          float oo = dot3(off, off);
          float* ci = S.a.cinert[lane];
          ci[0] = Iw[0] + mass * (oo - off[0] * off[0]);
          ci[1] = Iw[4] + mass * (oo - off[1] * off[1]);
          ci[2] = Iw[8] + mass * (oo - off[2] * off[2]);
          ci[3] = Iw[1] - mass * off[0] * off[1];
          ci[4] = Iw[2] - mass * off[0] * off[2];
          ci[5] = Iw[5] - mass * off[1] * off[2];
          ci[6] = mass * off[0]; ci[7] = mass * off[1]; ci[8] = mass * off[2];
          ci[9] = mass;
        }
        if (lane < NV) {
          const int j = C.dof_jnt[lane], b = C.dof_body[lane];
          float* cd = S.cdof[lane];
          if (C.jnt_type[j] == 0) {
            const int k = lane - C.jnt_dofadr[j];
            if (k < 3) {
              cd[0] = cd[1] = cd[2] = 0.0f;
              cd[3] = k == 0 ? 1.0f : 0.0f; cd[4] = k == 1 ? 1.0f : 0.0f; cd[5] = k == 2 ? 1.0f : 0.0f;
            } else {
              float Rb[9];
              quat_to_mat(Rb, S.a.xquat[b]);
              float ax[3];  // column (k-3) of the body rotation, selected without dynamic register indexing
              ax[0] = k == 3 ? Rb[0] : k == 4 ? Rb[1] : Rb[2];
              ax[1] = k == 3 ? Rb[3] : k == 4 ? Rb[4] : Rb[5];
              ax[2] = k == 3 ? Rb[6] : k == 4 ? Rb[7] : Rb[8];
              float off[3] = {com[0] - S.a.xpos[b][0], com[1] - S.a.xpos[b][1], com[2] - S.a.xpos[b][2]};
              cd[0] = ax[0]; cd[1] = ax[1]; cd[2] = ax[2];
              cross3(cd + 3, ax, off);
            }
          } else {
            float ax[3] = {cd[0], cd[1], cd[2]}, off[3] = {com[0] - cd[3], com[1] - cd[4], com[2] - cd[5]};
            cross3(cd + 3, ax, off);
          }
        }
      }
      __syncwarp();

      MJXB_STAMP(3);
      // ---------------------------------------------------------------- com_vel + cacc: two inclusive prefix sums over the dof tree
      // cvel[b] = sum of cdof_d*qvel_d over the dofs moving b; cdof_dot_d = (cvel before dof d) x cdof_d; cacc[b] = -g + sum cdof_dot_d*qvel_d
      // (mjx smooth.com_vel / rne forward pass; the free joint's angular dofs see the velocity after its three linear dofs only).
      {
        float I6[6], W6[6], cd[6] = {0, 0, 0, 0, 0, 0};
        const float qv = S.vec[VQVEL][lane];
        if (lane < NV) {
#pragma unroll
          for (int k = 0; k < 6; k++) cd[k] = S.cdof[lane][k];
        }
#pragma unroll
        for (int k = 0; k < 6; k++) I6[k] = cd[k] * qv;
        int dp_ = lane < NV ? C.dof_parent[lane] : -1;
        for (int st = 0; st < C.tree_steps; st++) {
          const int src = dp_ < 0 ? 0 : dp_;
          float g6[6];
#pragma unroll
          for (int k = 0; k < 6; k++) g6[k] = __shfl_sync(FULL, I6[k], src);
          const int p2 = __shfl_sync(FULL, dp_, src);
          if (dp_ >= 0) {
#pragma unroll
            for (int k = 0; k < 6; k++) I6[k] += g6[k];
            dp_ = p2;
          }
        }
        const int cs_ = lane < NV ? C.dof_cvel_src[lane] : -2;
        float cb[6];
#pragma unroll
        for (int k = 0; k < 6; k++) cb[k] = __shfl_sync(FULL, I6[k], cs_ < 0 ? 0 : cs_);
        if (cs_ < 0) {
#pragma unroll
          for (int k = 0; k < 6; k++) cb[k] = 0.0f;
        }
        float cdd[6];
        motion_cross(cdd, cb, cd);
#pragma unroll
        for (int k = 0; k < 6; k++) W6[k] = (cs_ == -2) ? 0.0f : cdd[k] * qv;
        dp_ = lane < NV ? C.dof_parent[lane] : -1;
        for (int st = 0; st < C.tree_steps; st++) {
          const int src = dp_ < 0 ? 0 : dp_;
          float g6[6];
#pragma unroll
          for (int k = 0; k < 6; k++) g6[k] = __shfl_sync(FULL, W6[k], src);
          const int p2 = __shfl_sync(FULL, dp_, src);
          if (dp_ >= 0) {
#pragma unroll
            for (int k = 0; k < 6; k++) W6[k] += g6[k];
            dp_ = p2;
          }
        }
        const int ld = lane < nbody ? C.body_lastdof[lane] : -1;
        float cv[6], ca[6];
#pragma unroll
        for (int k = 0; k < 6; k++) { cv[k] = __shfl_sync(FULL, I6[k], ld < 0 ? 0 : ld); ca[k] = __shfl_sync(FULL, W6[k], ld < 0 ? 0 : ld); }
        if (lane < nbody) {
#pragma unroll
          for (int k = 0; k < 6; k++) { S.a.cvel[lane][k] = ld < 0 ? 0.0f : cv[k]; S.a.cacc[lane][k] = ld < 0 ? 0.0f : ca[k]; }
          S.a.cacc[lane][3] -= C.gravity[0]; S.a.cacc[lane][4] -= C.gravity[1]; S.a.cacc[lane][5] -= C.gravity[2];
        }
      }
      __syncwarp();
      if (lane < nbody) {  // f = I*cacc + cvel x* (I*cvel), written over cacc[b]
        float ci[10], ca[6], cv[6], f1[6], iv[6], f2[6];
#pragma unroll
        for (int k = 0; k < 10; k++) ci[k] = S.a.cinert[lane][k];
#pragma unroll
        for (int k = 0; k < 6; k++) { ca[k] = S.a.cacc[lane][k]; cv[k] = S.a.cvel[lane][k]; }
        inert_mul(f1, ci, ca);
        inert_mul(iv, ci, cv);
        motion_cross_force(f2, cv, iv);
#pragma unroll
        for (int k = 0; k < 6; k++) S.a.cacc[lane][k] = (lane == 0) ? 0.0f : f1[k] + f2[k];
      }
      __syncwarp();
      {  // subtree sums (bodies are in DFS order: a subtree is the contiguous id range [b, subtree_end[b]))
        float crb[10], cf[6];
#pragma unroll
        for (int k = 0; k < 10; k++) crb[k] = 0.0f;
#pragma unroll
        for (int k = 0; k < 6; k++) cf[k] = 0.0f;
        if (lane >= 1 && lane < nbody) {
          for (int c = lane; c < C.body_subtree_end[lane]; c++) {
#pragma unroll
            for (int k = 0; k < 10; k++) crb[k] += S.a.cinert[c][k];
#pragma unroll
            for (int k = 0; k < 6; k++) cf[k] += S.a.cacc[c][k];
          }
        }
        __syncwarp();
        if (lane < nbody) {
#pragma unroll
          for (int k = 0; k < 10; k++) S.a.cinert[lane][k] = crb[k];
#pragma unroll
          for (int k = 0; k < 6; k++) S.a.cvel[lane][k] = cf[k];
        }
      }
      for (int i = lane; i < NV * NVP; i += 32) S.M[i] = 0.0f;
      __syncwarp();

      MJXB_STAMP(4);
      // ---------------------------------------------------------------- crb mass matrix, bias, passive, actuation
      float qfs = 0.0f;  // qfrc_smooth
      if (lane < NV) {
        float cd[6], f[6];
#pragma unroll
        for (int k = 0; k < 6; k++) cd[k] = S.cdof[lane][k];
        inert_mul(f, S.a.cinert[C.dof_body[lane]], cd);
        for (int j = lane; j >= 0; j = C.dof_parent[j]) {
          float s = 0.0f;
#pragma unroll
          for (int k = 0; k < 6; k++) s += S.cdof[j][k] * f[k];
          if (j == lane) s += C.dof_armature[lane];
          S.M[lane * NVP + j] = s;
          S.M[j * NVP + lane] = s;
        }
        float bias = 0.0f;
        const float* cf = S.a.cvel[C.dof_body[lane]];
#pragma unroll
        for (int k = 0; k < 6; k++) bias += cd[k] * cf[k];
        float passive = 0.0f;
        const int qa = C.dof_qadr[lane];
        if (qa >= 0) passive = -C.dof_stiffness[lane] * (S.vec[VQPOS][qa] - C.qpos_spring[qa]) - C.dof_damping[lane] * S.vec[VQVEL][lane];
        float qfrc_act = 0.0f;
        const int u = C.dof_act[lane];
        if (u >= 0) qfrc_act = C.dof_gear[lane] * clampf(S.vec[VCTRL][u], C.dof_ctrl_lo[lane], C.dof_ctrl_hi[lane]);
        qfs = passive - bias + qfrc_act;
        if (DBG) {
          if (A.dbg.qfrc_bias) A.dbg.qfrc_bias[(size_t)env * NV + lane] = bias;
          if (A.dbg.qfrc_passive) A.dbg.qfrc_passive[(size_t)env * NV + lane] = passive;
          if (A.dbg.qfrc_actuator) A.dbg.qfrc_actuator[(size_t)env * NV + lane] = qfrc_act;
        }
      }
      __syncwarp();
      if (DBG && A.dbg.qM) {
        for (int i = lane; i < NV * NV; i += 32) A.dbg.qM[(size_t)env * NV * NV + i] = S.M[(i / NV) * NVP + (i % NV)];
      }

      MJXB_STAMP(5);
      // ---------------------------------------------------------------- collision: lane per geom pair, candidates compacted by ballot
      int ncc = 0;
      bool tree_ok = C.tree_chol_ok != 0;
      // Broad phase: a pair whose bounding spheres (capsule: half length + radius around the geom centre; plane: centre height) are
      // apart cannot penetrate, so it cannot become a candidate (dist < 0). The survivors (typically 10-20 of 108) are compacted in
      // pair order -- the candidate order, hence every later rounding, is unchanged -- and the narrow phase below runs over one or two
      // strips of 32 instead of four. Debug builds keep every pair (they report all distances).
      int* pot_list = reinterpret_cast<int*>(&S.J[0]);   // J is not written before the constraint stage
      int npot = 0;
      for (int base = 0; base < C.npair; base += 32) {
        const int p = base + lane;
        bool pot = false;
        if (p < C.npair) {
          pot = true;
          if (!DBG) {
            const uint32_t w0 = C.pair_w0[p];
            const int g1 = w0 & 0xff, g2 = (w0 >> 8) & 0xff, kd = (w0 >> 16) & 0xff;
            const float dx = S.gpos[g2][0] - S.gpos[g1][0], dy = S.gpos[g2][1] - S.gpos[g1][1], dz = S.gpos[g2][2] - S.gpos[g1][2];
            const float e2 = C.geom_rad[g2] + C.geom_half[g2];
            if (kd == PAIR_PLANE_CAPSULE || kd == PAIR_PLANE_SPHERE) {
              pot = dx * S.gaxis[g1][0] + dy * S.gaxis[g1][1] + dz * S.gaxis[g1][2] - e2 <= 1e-5f;
            } else {
              const float reach = C.geom_rad[g1] + C.geom_half[g1] + e2;
              pot = dx * dx + dy * dy + dz * dz <= reach * reach * 1.0001f + 1e-6f;
            }
          }
        }
        const unsigned mk = __ballot_sync(FULL, pot);
        if (pot) pot_list[npot + __popc(mk & lt_mask)] = p;
        npot += __popc(mk);
      }
      __syncwarp();
      for (int base = 0; base < npot; base += 32) {
        const bool valid = base + lane < npot;
        const int p = valid ? pot_list[base + lane] : 0;
        float dist[2] = {1.0f, 1.0f}, cpos[2][3] = {{0, 0, 0}, {0, 0, 0}}, n[3] = {0, 0, 1}, t1[3] = {0, 0, 0}, t2[3] = {0, 0, 0};
        int kind = -1, condim = 1;
        bool cross_branch = false;
        if (valid) {
          const uint32_t w0 = C.pair_w0[p];
          const int g1 = w0 & 0xff, g2 = (w0 >> 8) & 0xff;
          kind = (w0 >> 16) & 0xff; condim = (w0 >> 24) & 0x7f; cross_branch = (w0 >> 31) != 0u;
          float p1[3] = {S.gpos[g1][0], S.gpos[g1][1], S.gpos[g1][2]}, p2[3] = {S.gpos[g2][0], S.gpos[g2][1], S.gpos[g2][2]};
          float ax1[3] = {S.gaxis[g1][0], S.gaxis[g1][1], S.gaxis[g1][2]}, ax2[3] = {S.gaxis[g2][0], S.gaxis[g2][1], S.gaxis[g2][2]};
          const float r1 = C.geom_rad[g1], r2 = C.geom_rad[g2], l1 = C.geom_half[g1], l2 = C.geom_half[g2];
          if (kind == PAIR_PLANE_CAPSULE || kind == PAIR_PLANE_SPHERE) {
            n[0] = ax1[0]; n[1] = ax1[1]; n[2] = ax1[2];
            if (kind == PAIR_PLANE_CAPSULE) {
              float na = dot3(ax1, ax2);
              float b[3] = {ax2[0] - ax1[0] * na, ax2[1] - ax1[1] * na, ax2[2] - ax1[2] * na};
              float bn = normalize3(b);
              if (bn < 0.5f) {
                b[0] = 0.0f; b[1] = 0.0f; b[2] = 0.0f;
                if (-0.5f < ax1[1] && ax1[1] < 0.5f) b[1] = 1.0f; else b[2] = 1.0f;
              }
              t1[0] = b[0]; t1[1] = b[1]; t1[2] = b[2];
              cross3(t2, ax1, b);
#pragma unroll
              for (int e = 0; e < 2; e++) {
                const float sg = e == 0 ? 1.0f : -1.0f;
                float sp[3] = {p2[0] + sg * (ax2[0] * l2), p2[1] + sg * (ax2[1] * l2), p2[2] + sg * (ax2[2] * l2)};
                float df[3] = {sp[0] - p1[0], sp[1] - p1[1], sp[2] - p1[2]};
                float d_ = dot3(df, ax1) - r2;
                dist[e] = d_;
                float s_ = r2 + 0.5f * d_;
                cpos[e][0] = sp[0] - ax1[0] * s_; cpos[e][1] = sp[1] - ax1[1] * s_; cpos[e][2] = sp[2] - ax1[2] * s_;
              }
            } else {
              float df[3] = {p2[0] - p1[0], p2[1] - p1[1], p2[2] - p1[2]};
              float d_ = dot3(df, ax1) - r2;
              dist[0] = d_;
              float s_ = r2 + 0.5f * d_;
              cpos[0][0] = p2[0] - ax1[0] * s_; cpos[0][1] = p2[1] - ax1[1] * s_; cpos[0][2] = p2[2] - ax1[2] * s_;
              make_tangents(n, t1, t2);
            }
          } else {
            float pa[3] = {p1[0], p1[1], p1[2]}, pb[3] = {p2[0], p2[1], p2[2]};
            if (kind == PAIR_SPHERE_CAPSULE) {
              float a[3] = {p2[0] - ax2[0] * l2, p2[1] - ax2[1] * l2, p2[2] - ax2[2] * l2};
              float b[3] = {p2[0] + ax2[0] * l2, p2[1] + ax2[1] * l2, p2[2] + ax2[2] * l2};
              closest_segment_point(pb, a, b, p1);
            } else if (kind == PAIR_CAPSULE_CAPSULE) {
              float a0[3] = {p1[0] - ax1[0] * l1, p1[1] - ax1[1] * l1, p1[2] - ax1[2] * l1};
              float a1[3] = {p1[0] + ax1[0] * l1, p1[1] + ax1[1] * l1, p1[2] + ax1[2] * l1};
              float b0[3] = {p2[0] - ax2[0] * l2, p2[1] - ax2[1] * l2, p2[2] - ax2[2] * l2};
              float b1[3] = {p2[0] + ax2[0] * l2, p2[1] + ax2[1] * l2, p2[2] + ax2[2] * l2};
              closest_segment_to_segment(pa, pb, a0, a1, b0, b1);
            }
            sphere_sphere(dist[0], cpos[0], n, pa, r1, pb, r2);
            if (condim > 1) make_tangents(n, t1, t2);
          }
          if (DBG) {
            const int c0 = C.pair_w1[p] & 0xffff;
            const int ne = (kind == PAIR_PLANE_CAPSULE) ? 2 : 1;
            for (int e = 0; e < ne; e++) {
              if (A.dbg.con_dist) A.dbg.con_dist[(size_t)env * C.ncon + c0 + e] = dist[e];
              if (A.dbg.con_pos) for (int k = 0; k < 3; k++) A.dbg.con_pos[((size_t)env * C.ncon + c0 + e) * 3 + k] = cpos[e][k];
              if (A.dbg.con_normal) for (int k = 0; k < 3; k++) A.dbg.con_normal[((size_t)env * C.ncon + c0 + e) * 3 + k] = n[k];
            }
          }
        }
#pragma unroll
        for (int e = 0; e < 2; e++) {
          const bool cand = valid && (e == 0 || kind == PAIR_PLANE_CAPSULE) && dist[e] < 0.0f;
          const unsigned mk = __ballot_sync(FULL, cand);
          if (__ballot_sync(FULL, cand && cross_branch) != 0u) tree_ok = false;  // a row will couple two limbs: dense factorisation
          const int idx = ncc + __popc(mk & lt_mask);
          if (cand && idx < MAXCC) {
#pragma unroll
            for (int k = 0; k < 3; k++) { S.cc_n[idx][k] = n[k]; S.cc_t1[idx][k] = t1[k]; S.cc_t2[idx][k] = t2[k]; S.cc_pos[idx][k] = cpos[e][k]; }
            S.cc_dist[idx] = dist[e];
            S.cc_pair[idx] = p | (e << 16) | (condim << 20);
          }
          ncc += __popc(mk);
        }
      }
      if (ncc > MAXCC) { overflow = true; ncc = MAXCC; }
      __syncwarp();

      MJXB_STAMP(6);
      // ---------------------------------------------------------------- constraint rows: joint limits, tendon limits, contacts
      int nrow = 0;
      {
        // joint limits
        bool cand = false;
        float pos = 0.0f, sgn = 1.0f;
        if (lane < C.nlimit) {
          float qq = S.vec[VQPOS][C.lim_qadr[lane]];
          float dmin = qq - C.lim_range[lane][0], dmax = C.lim_range[lane][1] - qq;
          pos = fminf(dmin, dmax);
          sgn = dmin < dmax ? 1.0f : -1.0f;
          cand = pos < 0.0f;
        }
        unsigned mk = __ballot_sync(FULL, cand);
        int r = nrow + __popc(mk & lt_mask);
        if (cand && r < CAP) { S.rinfo[r] = ROW_LIMIT | (lane << 2) | (sgn < 0.0f ? 1 << 10 : 0); S.rJaref[r] = pos; }
        nrow += __popc(mk);
        // tendon limits
        cand = false;
        if (lane < C.ntlimit) {
          float len = 0.0f;
          for (int w = 0; w < C.ten_nwrap[lane]; w++) len += C.ten_coef[lane][w] * S.vec[VQPOS][C.ten_qpos[lane][w]];
          float dmin = len - C.ten_range[lane][0], dmax = C.ten_range[lane][1] - len;
          pos = fminf(dmin, dmax);
          sgn = dmin < dmax ? 1.0f : -1.0f;
          cand = pos < 0.0f;
        }
        mk = __ballot_sync(FULL, cand);
        r = nrow + __popc(mk & lt_mask);
        if (cand && r < CAP) { S.rinfo[r] = ROW_TENDON | (lane << 2) | (sgn < 0.0f ? 1 << 10 : 0); S.rJaref[r] = pos; }
        nrow += __popc(mk);
        // contacts: exclusive scan of rows per candidate, 32 candidates per strip (the big tier holds up to 176)
        int total = 0;
#pragma unroll
        for (int cb = 0, cstrip_ = 0; cstrip_ < (MAXCC + 31) / 32 && cb < ncc; cstrip_++, cb += 32) {
          const int c = cb + lane;
          int nr = 0;
          if (c < ncc) nr = ((S.cc_pair[c] >> 20) > 1) ? 4 : 1;
          int scan = nr;
#pragma unroll
          for (int o = 1; o < 32; o <<= 1) {
            int t = __shfl_up_sync(FULL, scan, o);
            if (lane >= o) scan += t;
          }
          const int rbase = nrow + total + scan - nr;
          if (c < ncc) {
            S.cc_row[c] = (rbase + nr <= CAP) ? rbase : -1;
            S.rforce[c] = pair_param[S.cc_pair[c] & 0xffff].mu;   // friction of every candidate fetched in one round (global memory);
          }                                                        // rforce is free until the solver starts
          total += __shfl_sync(FULL, scan, 31);
        }
        if (nrow + total > CAP) overflow = true;
        if (overflow) { total = 0; ncc = 0; nrow = min(nrow, CAP); }  // results are discarded; keep the row set inside the tile
        const int nrow_lim = nrow;
        nrow += total;
        __syncwarp();
        // cooperative Jacobian rows: limits / tendons
        for (int rr = 0; rr < nrow_lim; rr++) {
          const int info = S.rinfo[rr];
          const int idx = (info >> 2) & 0xff;
          const float sg = (info & (1 << 10)) ? -1.0f : 1.0f;
          float val = 0.0f;
          if ((info & 3) == ROW_LIMIT) {
            val = (lane == C.lim_dof[idx]) ? sg : 0.0f;
          } else {
            for (int w = 0; w < C.ten_nwrap[idx]; w++)
              if (C.ten_dof[idx][w] == lane) val += sg * C.ten_coef[idx][w];
          }
          if (lane < NVP) S.J[rr * NVP + lane] = val;
        }
        // contact rows: lane d evaluates its column of the point Jacobian difference (mjx support.jac)
        float cdl[6] = {0, 0, 0, 0, 0, 0};
        if (lane < NV) {
#pragma unroll
          for (int k = 0; k < 6; k++) cdl[k] = S.cdof[lane][k];
        }
        for (int c = 0; c < ncc; c++) {
          const int pr = S.cc_pair[c], p = pr & 0xffff, rb = S.cc_row[c];
          const uint32_t w0 = C.pair_w0[p];
          const int b1 = C.geom_body[w0 & 0xff], b2 = C.geom_body[(w0 >> 8) & 0xff];
          const bool c3 = (pr >> 20) > 1;
          float off[3] = {S.cc_pos[c][0] - com[0], S.cc_pos[c][1] - com[1], S.cc_pos[c][2] - com[2]};
          float jc[3], sgnb = 0.0f;
          cross3(jc, cdl, off);
          jc[0] += cdl[3]; jc[1] += cdl[4]; jc[2] += cdl[5];
          if ((C.body_dofmask[b2] >> lane) & 1u) sgnb += 1.0f;
          if ((C.body_dofmask[b1] >> lane) & 1u) sgnb -= 1.0f;
          jc[0] *= sgnb; jc[1] *= sgnb; jc[2] *= sgnb;
          float jn = dot3(S.cc_n[c], jc);
          if (lane < NVP) {
            if (!c3) {
              S.J[rb * NVP + lane] = jn;
            } else {
              const float mu = S.rforce[c];
              float j1 = dot3(S.cc_t1[c], jc) * mu, j2 = dot3(S.cc_t2[c], jc) * mu;
              S.J[(rb + 0) * NVP + lane] = jn + j1;
              S.J[(rb + 1) * NVP + lane] = jn - j1;
              S.J[(rb + 2) * NVP + lane] = jn + j2;
              S.J[(rb + 3) * NVP + lane] = jn - j2;
            }
          }
          if (lane < (c3 ? 4 : 1)) {
            S.rinfo[rb + lane] = (c3 ? ROW_CON3 : ROW_CON1) | (c << 2) | (lane << 11);
            S.rJaref[rb + lane] = S.cc_dist[c];
          }
        }
        __syncwarp();
        // per-row impedance / reference acceleration (lane per row)
        const float4* qv4 = reinterpret_cast<const float4*>(&S.vec[VQVEL][0]);
        MJXB_FOR_ROW_STRIPS(rr) {
          const int info = S.rinfo[rr], kind = info & 3, idx = (info >> 2) & 0xff;
          const float rpos = S.rJaref[rr];
          float invw, solref[2], solimp[5];
          int efc_row;
          if (kind == ROW_LIMIT) {
            invw = C.lim_invweight[idx]; solref[0] = C.lim_solref[idx][0]; solref[1] = C.lim_solref[idx][1];
#pragma unroll
            for (int k = 0; k < 5; k++) solimp[k] = C.lim_solimp[idx][k];
            efc_row = C.lim_row[idx];
          } else if (kind == ROW_TENDON) {
            invw = C.ten_invweight[idx]; solref[0] = C.ten_solref[idx][0]; solref[1] = C.ten_solref[idx][1];
#pragma unroll
            for (int k = 0; k < 5; k++) solimp[k] = C.ten_solimp[idx][k];
            efc_row = C.ten_row[idx];
          } else {
            const int pr = S.cc_pair[idx], p = pr & 0xffff, e = (pr >> 16) & 0xf;
            const PairParam pp = pair_param[p];
            invw = pp.invweight; solref[0] = pp.solref[0]; solref[1] = pp.solref[1];
#pragma unroll
            for (int k = 0; k < 5; k++) solimp[k] = pp.solimp[k];
            const int efc0 = C.pair_w1[p] >> 16;
            efc_row = (kind == ROW_CON1) ? efc0 + e : efc0 + 4 * e + ((info >> 11) & 3);
          }
          float kk, bb, imp;
          kbi(h, solref, solimp, rpos, kk, bb, imp);
          float rr_ = fmaxf(invw * (1.0f - imp) / imp, MINVAL);
          const float4* row = reinterpret_cast<const float4*>(&S.J[rr * NVP]);
          float vel = 0.0f;
#pragma unroll
          for (int g = 0; g < NVP / 4; g++) {
            float4 m = row[g], xx = qv4[g];
            vel += m.x * xx.x; vel += m.y * xx.y; vel += m.z * xx.z; vel += m.w * xx.w;
          }
          const float D = 1.0f / rr_, aref = -bb * vel - kk * imp * rpos;
          S.rD[rr] = D;
          S.raref[rr] = aref;
          S.rinfo[rr] = (info & 0xffff) | (efc_row << 16);
          if (DBG) {
            if (A.dbg.efc_pos) A.dbg.efc_pos[(size_t)env * C.nefc + efc_row] = rpos;
            if (A.dbg.efc_D) A.dbg.efc_D[(size_t)env * C.nefc + efc_row] = D;
            if (A.dbg.efc_aref) A.dbg.efc_aref[(size_t)env * C.nefc + efc_row] = aref;
          }
        }
        __syncwarp();
      }
      if (A.lockstep > 0 && pass == 0) group_sync(warp, A.lockstep_group);  // all warps of the group enter the solver code together

      MJXB_STAMP(7);
      // ---------------------------------------------------------------- solve: one factor/solve code instance drives
      //   phase 0: qacc_smooth = M^-1 qfrc_smooth         (mjx smooth.factor_m/solve_m)
      //   phase 1: Newton  Mgrad = (M + J^T D_act J)^-1 grad (mjx solver._update_gradient), repeated
      //   phase 2: qacc'   = (M + h*damping)^-1 (qfrc_smooth + qfrc_constraint)   (mjx forward.implicit / euler)
      float qas = 0.0f, qacc = 0.0f, Ma = 0.0f, qfc = 0.0f, grad = 0.0f, search = 0.0f;
      float gauss = 0.0f, cost = 0.0f, prev_cost = 0.0f, prev_grad = 0.0f, prev_Mgrad = 0.0f, grad_sq = 0.0f;
      int niter = 0, phase = 0;
      bool factor_valid = false;           // S.L / dinv_keep hold the factor of H for the current active set (tree pattern only)
      float* dinv_keep = &S.gpos[0][0];    // geom frames are dead after the collision stage: 32 floats of 1/R_ii
      static_assert(sizeof(S.gpos) >= 32 * sizeof(float), "dinv_keep needs 32 floats");
      const float scale = 1.0f / (C.meaninertia * (float)max(1, C.nv));
      float qacc_int = 0.0f;
      const bool integrate_pass = (mode == MODE_ENV_STEP || mode == MODE_PHYS_STEP || mode == MODE_SPEED_TEST);

      // update_constraint (mjx solver._update_constraint) on the current Jaref/Ma/qacc
      auto update_constraint = [&]() {
        float cs = 0.0f;
        MJXB_FOR_ROW_STRIPS(r) {
          const float ja = S.rJaref[r], D = S.rD[r];
          const bool active = ja < 0.0f;
          S.rforce[r] = active ? -D * ja : 0.0f;
          cs += active ? D * ja * ja : 0.0f;
        }
        __syncwarp();
        qfc = rowsT_times(S, lane, nrow, S.rforce);
        grad = Ma - qfs - qfc;
        // three reductions side by side (the compiler interleaves the shuffle chains): gauss, cost, and the squared gradient norm
        // that the next termination test needs -- taken here it costs no dependent round of its own
        gauss = 0.5f * warp_sum((lane < NV) ? (Ma - qfs) * (qacc - qas) : 0.0f);
        grad_sq = warp_sum((lane < NV) ? grad * grad : 0.0f);
        prev_cost = cost;
        cost = 0.5f * warp_sum(cs) + gauss;
      };

      while (true) {
        if (A.lockstep == 3 && pass == 0) __syncthreads_or(1);  // re-align the warps of the CTA at every factor/solve round (finished warps answer below)
        asm volatile("" : "+r"(phase));  // keep `phase` opaque: the compiler otherwise clones the whole factor/solve body per phase
        if (phase == 1) {  // mjx solver.solve cond(): evaluated before paying for the next factorisation
          bool done;
          if (C.iterations == 1) {
            done = niter >= 1;
          } else {
            const float improvement = (prev_cost - cost) * scale;
            const float gradient = sqrtf(grad_sq) * scale;
            done = (niter >= C.iterations) || (improvement < C.tolerance) || (gradient < C.tolerance);
            if (niter >= C.iterations && !(improvement < C.tolerance) && !(gradient < C.tolerance)) status |= MJXB_STATUS_MAXITER;
          }
          if (done) {
            if (!integrate_pass) break;
            if (!C.damp_implicit) { qacc_int = qacc; break; }  // plain Euler without eulerdamp integrates the solver's qacc
            phase = 2;
          }
        }
        if (phase == 1 && niter == 0) MJXB_STAMP(16);
        // ---- assemble the rows of the system matrix in registers, factor, solve
        // Newton iterations whose active set did not change since the last factorisation (typically the final, polishing ones)
        // reuse that factor: H = M + J^T diag(D*active) J would be rebuilt bit for bit, so the solve reads the rows of R kept in S.L.
        const float rhs = (phase == 0) ? qfs : (phase == 1) ? grad : (qfs + qfc);
        float x;
        if (MJXB_FACTOR_REUSE && LS_EXACT && phase == 1 && factor_valid) {
          x = chol_tree_solve_lds(S, lane, dinv_keep[lane], rhs);
        } else {
          float a[NVP];
          {
            const float4* row = reinterpret_cast<const float4*>(&S.M[(lane < NV ? lane : 0) * NVP]);
#pragma unroll
            for (int g = 0; g < NVP / 4; g++) {
              float4 m = row[g];
              a[4 * g] = m.x; a[4 * g + 1] = m.y; a[4 * g + 2] = m.z; a[4 * g + 3] = m.w;
            }
          }
          if (phase == 2) {
            const float hd = (lane < NV) ? h * C.dof_damping[lane] : 0.0f;
#pragma unroll
            for (int j = 0; j < NV; j++) a[j] += (j == lane) ? hd : 0.0f;
          }
          if (phase == 1 && (LS_EXACT || C.solver == 2)) {  // Newton: H = M + J^T diag(D*active) J ; CG preconditions with M alone
            // rows active at the current iterate, as a bit mask per strip of 32: the loop below then depends on no shared-memory
            // load (row r+1's loads are issued under row r's FMAs) and inactive rows cost nothing
#pragma unroll 1
            for (int st = 0; st < NSTRIP; st++) {
              const int rl = st * 32 + lane;
              unsigned act = __ballot_sync(FULL, rl < nrow && S.rJaref[rl] < 0.0f);
              while (act != 0u) {
                const int r = st * 32 + __ffs(act) - 1;
                act &= act - 1u;
                const float w = S.rD[r] * S.J[r * NVP + (lane < NVP ? lane : 0)];
                const float4* jr = reinterpret_cast<const float4*>(&S.J[r * NVP]);
#pragma unroll
                for (int g = 0; g < NVP / 4; g++) {
                  float4 jj = jr[g];
                  a[4 * g] += w * jj.x; a[4 * g + 1] += w * jj.y; a[4 * g + 2] += w * jj.z; a[4 * g + 3] += w * jj.w;
                }
              }
            }
          }
          if (lane >= NV) {  // idle lanes: identity rows keep the arithmetic finite
#pragma unroll
            for (int j = 0; j < NVP; j++) a[j] = 0.0f;
          }
          float dinv;
          if (tree_ok || (phase != 1 && C.tree_chol_ok != 0)) {  // M and M + h*damping always follow the tree pattern; H does unless a row couples two limbs
            float zf = (lane < NV) ? rhs : 0.0f;   // forward substitution fused into the factorisation
            if (phase == 1 && niter == 0) MJXB_STAMP(17);
            chol_tree(S, lane, a, dinv, zf);
            if (phase == 1 && niter == 0) MJXB_STAMP(18);
            x = chol_tree_solve(S, lane, a, dinv, zf);
            if (phase == 1 && niter == 0) MJXB_STAMP(19);
            if (MJXB_FACTOR_REUSE && LS_EXACT && phase == 1) { dinv_keep[lane] = dinv; factor_valid = true; }
          } else {
            chol_rows(S, lane, a, dinv);
            x = chol_solve_rows(S, lane, a, dinv, (lane < NV) ? rhs : 0.0f);
          }
        }

        if (phase == 0) {
          MJXB_STAMP(14);
          qas = x;
          if (DBG && A.dbg.qacc_smooth && lane < NV) A.dbg.qacc_smooth[(size_t)env * NV + lane] = qas;
          // warm start (mjx solver.solve): cheaper of qacc_warmstart and qacc_smooth
          const float w0 = S.vec[VX][lane];  // qacc_warmstart (parked at the top of the pass)
          // both candidates (qacc_warmstart, qacc_smooth) are evaluated side by side: one pass over M and J, one loop over the rows,
          // four reductions next to each other (each sum keeps its own order: bit-identical to evaluating them one after the other)
          float Ma_w, Ma_s;
          matvec_M_and_rows2(S, lane, nrow, w0, qas, S.rjv, S.rJaref, Ma_w, Ma_s);
          float cs = 0.0f, cs2 = 0.0f;
          MJXB_FOR_ROW_STRIPS(r) {
            const float ar = S.raref[r], D = S.rD[r];
            const float ja = S.rjv[r] - ar, jb = S.rJaref[r] - ar;
            S.rjv[r] = ja;
            S.rJaref[r] = jb;
            cs += ja < 0.0f ? D * ja * ja : 0.0f;
            cs2 += jb < 0.0f ? D * jb * jb : 0.0f;
          }
          const float cost_w = 0.5f * warp_sum(cs) + 0.5f * warp_sum((lane < NV) ? (Ma_w - qfs) * (w0 - qas) : 0.0f);
          const float cost_s = 0.5f * warp_sum(cs2) + 0.5f * warp_sum((lane < NV) ? (Ma_s - qfs) * (qas - qas) : 0.0f);
          const bool use_warm = cost_w < cost_s;
          qacc = use_warm ? w0 : qas;
          Ma = use_warm ? Ma_w : Ma_s;
          if (use_warm) MJXB_FOR_ROW_STRIPS(r) S.rJaref[r] = S.rjv[r];
          __syncwarp();
          MJXB_STAMP(15);
          cost = __int_as_float(0x7f800000);  // Context.create: cost = inf, prev_cost = 0
          prev_cost = 0.0f;
          update_constraint();
          phase = 1;
          MJXB_STAMP(8);
          continue;
        }
        if (phase == 2) { qacc_int = x; break; }

        // ---- phase 1: x = Mgrad
        if (LS_EXACT || C.solver == 2 || niter == 0) {  // (the LS_EXACT instantiation is Newton-only)
          search = -x;
        } else {  // CG, Polak-Ribiere (mjx solver.solve body)
          const float num = warp_sum((lane < NV) ? grad * (x - prev_Mgrad) : 0.0f);
          const float den = warp_sum((lane < NV) ? prev_grad * prev_Mgrad : 0.0f);
          const float beta = fmaxf(0.0f, num / fmaxf(MINVAL, den));
          search = -x + beta * search;
        }
        if (!LS_EXACT) { prev_grad = grad; prev_Mgrad = x; }
        // ---- line search along `search`: minimise f(alpha) = gauss-quadratic + sum_r [Jaref_r + alpha jv_r < 0] D_r (Jaref_r + alpha jv_r)^2 / 2
        {
          const float mv = matvec_M_and_rows(S, lane, nrow, search, S.rjv);
          if (niter == 0) MJXB_STAMP(20);
          const float qg0 = gauss;
          const float qg1 = warp_sum((lane < NV) ? search * Ma : 0.0f) - warp_sum((lane < NV) ? search * qfs : 0.0f);
          const float qg2 = 0.5f * warp_sum((lane < NV) ? search * mv : 0.0f);
          float alpha_step;   // step taken (0 = stay)
          if (LS_EXACT) {
            // Exact 1-D minimiser. f' is continuous, piecewise linear and non-decreasing with breakpoints
            // alpha_r = -Jaref_r / jv_r; MJX's bracketed Newton iteration (solver._linesearch) converges to the same point.
            // Lane r evaluates f'(alpha_r) over all rows (128-bit broadcast loads, no shuffles), two warp min/max reductions
            // bracket the root, and the root of the linear piece inside the bracket is taken in closed form.
            if (niter == 0) MJXB_STAMP(24);
            float* rls = S.rforce;   // per-row D*jv, published for the sweep (rforce is rebuilt by update_constraint)
            float al[NSTRIP], gp[NSTRIP];
            bool ok[NSTRIP];
#pragma unroll
            for (int st = 0; st < NSTRIP; st++) {
              const int r = st * 32 + lane;
              al[st] = 0.0f; gp[st] = 0.0f; ok[st] = false;
              if (r < nrow) {
                const float ja = S.rJaref[r], jv = S.rjv[r];
                rls[r] = S.rD[r] * jv;
                const float a_ = -ja / jv;
                ok[st] = (jv != 0.0f) && (a_ > 0.0f) && (a_ < 3.0e38f);
                al[st] = ok[st] ? a_ : 0.0f;
              }
            }
            __syncwarp();
            if (niter == 0) MJXB_STAMP(25);
            for (int r2 = 0; r2 < nrow; r2++) {
              const float ja = S.rJaref[r2], jv = S.rjv[r2], dj = rls[r2];
#pragma unroll
              for (int st = 0; st < NSTRIP; st++) {
                const float x = fmaf(al[st], jv, ja);
                gp[st] += (x < 0.0f) ? dj * x : 0.0f;
              }
            }
            if (niter == 0) MJXB_STAMP(26);
            unsigned lo_b = 0u, hi_b = 0x7f800000u;  // positive floats order like their bit patterns
#pragma unroll
            for (int st = 0; st < NSTRIP; st++) {
              const float g_ = gp[st] + qg1 + 2.0f * al[st] * qg2;
              if (ok[st] && g_ < 0.0f) lo_b = max(lo_b, __float_as_uint(al[st]));
              if (ok[st] && !(g_ < 0.0f)) hi_b = min(hi_b, __float_as_uint(al[st]));
            }
            lo_b = __reduce_max_sync(FULL, lo_b);
            hi_b = __reduce_min_sync(FULL, hi_b);
            const float a_lo = __uint_as_float(lo_b), a_hi = __uint_as_float(hi_b);
            const float a_mid = (hi_b == 0x7f800000u) ? (2.0f * a_lo + 1.0f) : 0.5f * (a_lo + a_hi);
            if (niter == 0) MJXB_STAMP(27);
            float sa = 0.0f, sb = 0.0f, sc = 0.0f;  // quadratic piece on (a_lo, a_hi): f = C + alpha A + alpha^2 B
            MJXB_FOR_ROW_STRIPS(r) {
              const float ja = S.rJaref[r], jv = S.rjv[r], dj = rls[r];
              const bool on = fmaf(a_mid, jv, ja) < 0.0f;
              sa += on ? dj * ja : 0.0f;
              sb += on ? dj * jv : 0.0f;
              sc += on ? S.rD[r] * ja * ja : 0.0f;
            }
            if (niter == 0) MJXB_STAMP(28);
            const float Aq = qg1 + warp_sum(sa), Bq = qg2 + 0.5f * warp_sum(sb), Cq = qg0 + 0.5f * warp_sum(sc);
            float a_star = (Bq > 0.0f) ? -Aq / (2.0f * Bq) : 0.0f;
            a_star = fminf(fmaxf(a_star, a_lo), a_hi);
            // MJX moves only if the evaluated cost improves on f(0) (solver._linesearch tail); this is what lets the outer loop's
            // `improvement < tolerance` test fire once float32 can no longer resolve a decrease
            const float cost_star = a_star * a_star * Bq + a_star * Aq + Cq;
            alpha_step = (a_star > 0.0f && a_star < 3.0e38f && cost_star < cost) ? a_star : 0.0f;
          } else {
          // MJX's bracketed iteration, kept verbatim for truncated settings (lighten_solver: ls_iterations = 1)
          const float snorm = sqrtf(warp_sum((lane < NV) ? search * search : 0.0f));
          const float gtol = C.tolerance * C.ls_tolerance * snorm * C.meaninertia * (float)max(1, C.nv);
          float ja[NSTRIP], jv[NSTRIP], qa[NSTRIP], qb[NSTRIP], qc[NSTRIP];
#pragma unroll
          for (int s = 0; s < NSTRIP; s++) {
            const int r = s * 32 + lane;
            ja[s] = 1.0f; jv[s] = 0.0f; qa[s] = qb[s] = qc[s] = 0.0f;
            if (r < nrow) {
              const float D = S.rD[r];
              ja[s] = S.rJaref[r]; jv[s] = S.rjv[r];
              qa[s] = 0.5f * ja[s] * ja[s] * D; qb[s] = jv[s] * ja[s] * D; qc[s] = 0.5f * jv[s] * jv[s] * D;
            }
          }
          auto point = [&](float alpha) {
            float p0 = 0.0f, p1 = 0.0f, p2 = 0.0f;
#pragma unroll
            for (int s = 0; s < NSTRIP; s++) {
              const bool on = (ja[s] + alpha * jv[s]) < 0.0f;
              p0 += on ? qa[s] : 0.0f; p1 += on ? qb[s] : 0.0f; p2 += on ? qc[s] : 0.0f;
            }
            p0 = qg0 + warp_sum(p0); p1 = qg1 + warp_sum(p1); p2 = qg2 + warp_sum(p2);
            LSPoint pt;
            pt.alpha = alpha;
            pt.cost = alpha * alpha * p2 + alpha * p1 + p0;
            pt.d0 = 2.0f * alpha * p2 + p1;
            pt.d1 = 2.0f * p2 + (p2 == 0.0f ? MINVAL : 0.0f);
            return pt;
          };
          const LSPoint p0 = point(0.0f);
          LSPoint lo = point(p0.alpha - p0.d0 / p0.d1), hi;
          if (lo.d0 < p0.d0) { hi = p0; } else { hi = lo; lo = p0; }
          bool swap = true;
          int ls_iter = 0;
          while (true) {
            bool ls_done = ls_iter >= C.ls_iterations;
            ls_done |= !swap;
            ls_done |= (lo.d0 < 0.0f) && (lo.d0 > -gtol);
            ls_done |= (hi.d0 > 0.0f) && (hi.d0 < gtol);
            if (ls_done) break;
            const LSPoint lo_next = point(lo.alpha - lo.d0 / lo.d1);
            const LSPoint hi_next = point(hi.alpha - hi.d0 / hi.d1);
            const LSPoint mid = point(0.5f * (lo.alpha + hi.alpha));
            const bool swap_lo_next = (lo.d0 > 0.0f) || (lo.d0 < lo_next.d0);
            if (swap_lo_next) lo = lo_next;
            const bool swap_lo_mid = (mid.d0 < 0.0f) && (lo.d0 < mid.d0);
            if (swap_lo_mid) lo = mid;
            const bool swap_hi_next = (hi.d0 < 0.0f) || (hi.d0 > hi_next.d0);
            if (swap_hi_next) hi = hi_next;
            const bool swap_hi_mid = (mid.d0 > 0.0f) && (hi.d0 > mid.d0);
            if (swap_hi_mid) hi = mid;
            swap = swap_lo_next || swap_lo_mid || swap_hi_next || swap_hi_mid;
            ls_iter++;
          }
          const bool improved = (lo.cost < p0.cost) || (hi.cost < p0.cost);
          alpha_step = improved ? (lo.cost < hi.cost ? lo.alpha : hi.alpha) : 0.0f;
          }
          if (niter == 0) MJXB_STAMP(21);
          if (alpha_step != 0.0f) {
            qacc += search * alpha_step;
            Ma += mv * alpha_step;
            bool flip = false;
            MJXB_FOR_ROW_STRIPS(r) {
              const float ja0 = S.rJaref[r], ja1 = ja0 + S.rjv[r] * alpha_step;
              flip |= (ja0 < 0.0f) != (ja1 < 0.0f);
              S.rJaref[r] = ja1;
            }
            if (LS_EXACT && __any_sync(FULL, flip)) factor_valid = false;
          }
          __syncwarp();
        }
        if (niter == 0) MJXB_STAMP(22);
        update_constraint();
        if (niter == 0) MJXB_STAMP(23);
        niter++;
      }  // factor/solve loop

      if (A.lockstep == 3 && pass == 0) { while (__syncthreads_or(0)) {} }  // finished warps keep answering the per-round barrier
      else if (A.lockstep > 0 && A.lockstep != 2 && pass == 0) group_sync(warp, A.lockstep_group);  // ... and leave it together (early finishers would idle at the round barrier anyway)
      if (A.work_out != nullptr && valid && !spec_partner && lane == 0) A.work_out[env] = (uint8_t)min(niter * MJXB_WORK_KEY_ITER + nrow * MJXB_WORK_KEY_ROW, 255);
      if (DBG) {
        if (lane < NV) {
          if (A.dbg.qacc) A.dbg.qacc[(size_t)env * NV + lane] = qacc;
          if (A.dbg.qfrc_constraint) A.dbg.qfrc_constraint[(size_t)env * NV + lane] = qfc;
        }
        if (A.dbg.solver_niter && lane == 0) A.dbg.solver_niter[env] = niter;
        MJXB_FOR_ROW_STRIPS(r) {
          const int efc_row = S.rinfo[r] >> 16;
          if (A.dbg.efc_force) A.dbg.efc_force[(size_t)env * C.nefc + efc_row] = S.rforce[r];
          if (A.dbg.efc_active) A.dbg.efc_active[(size_t)env * C.nefc + efc_row] = 1 | (S.rJaref[r] < 0.0f ? 2 : 0);
        }
      }

      MJXB_STAMP(9);
      // ---------------------------------------------------------------- touch sensors (mjx sensor.sensor_acc / engine_sensor.c mjSENS_TOUCH)
      {
#pragma unroll 1
        for (int s = 0; s < C.nsensor; s++) {
          float contrib = 0.0f;
          const int site = C.sensor_site[s], sb = C.site_body[site];
#pragma unroll
          for (int c = lane, cstrip_ = 0; cstrip_ < (MAXCC + 31) / 32 && c < ncc; cstrip_++, c += 32) {
            const int pr = S.cc_pair[c], p = pr & 0xffff, rb = S.cc_row[c];
            const uint32_t w0 = C.pair_w0[p];
            const int b1 = C.geom_body[w0 & 0xff], b2 = C.geom_body[(w0 >> 8) & 0xff];
            if (sb != b1 && sb != b2) continue;
            float nf = S.rforce[rb];
            if ((pr >> 20) > 1) nf = ((S.rforce[rb] + S.rforce[rb + 1]) + S.rforce[rb + 2]) + S.rforce[rb + 3];
            if (!(nf > 0.0f)) continue;
            float ray[3] = {S.cc_n[c][0] * nf, S.cc_n[c][1] * nf, S.cc_n[c][2] * nf};
            normalize3(ray);
            if (sb == b2) { ray[0] = -ray[0]; ray[1] = -ray[1]; ray[2] = -ray[2]; }
            const float dx = S.cc_pos[c][0] - S.site_xpos[site][0], dy = S.cc_pos[c][1] - S.site_xpos[site][1],
                        dz = S.cc_pos[c][2] - S.site_xpos[site][2];
            const float* Rm = S.site_xmat[site];
            const float lpx = Rm[0] * dx + Rm[3] * dy + Rm[6] * dz, lpy = Rm[1] * dx + Rm[4] * dy + Rm[7] * dz,
                        lpz = Rm[2] * dx + Rm[5] * dy + Rm[8] * dz;
            const float lvx = Rm[0] * ray[0] + Rm[3] * ray[1] + Rm[6] * ray[2], lvy = Rm[1] * ray[0] + Rm[4] * ray[1] + Rm[7] * ray[2],
                        lvz = Rm[2] * ray[0] + Rm[5] * ray[1] + Rm[8] * ray[2];
            if (ray_box(C.site_size[site][0], C.site_size[site][1], C.site_size[site][2], lpx, lpy, lpz, lvx, lvy, lvz) >= 0.0f) contrib += nf;
          }
          contrib = warp_sum(contrib);
          if (lane == 0) S.sens[s] = contrib;
        }
        __syncwarp();
        if (DBG && A.dbg.sensordata && lane < C.nsensor) A.dbg.sensordata[(size_t)env * C.nsensor + lane] = S.sens[lane];
      }

      const float qacc_solver = qacc;  // qacc_warmstart <- solver qacc (mjx solver.solve tail)
      if (!(fabsf(qacc_solver) < 3.0e38f)) status |= MJXB_STATUS_NAN;

      MJXB_STAMP(10);
      // ---------------------------------------------------------------- integrate (mjx forward._advance)
      if (integrate_pass) {
        q = S.vec[VQPOS][lane];
        v = S.vec[VQVEL][lane] + qacc_int * h;
        __syncwarp();
        S.vec[VQVEL][lane] = (lane < NV) ? v : 0.0f;
        __syncwarp();
        if (lane < C.nq) {
          const int kind = C.qpos_kind[lane], ax = C.qpos_aux[lane];
          if (kind == QK_HINGE || kind == QK_FREEPOS) {
            q = q + h * S.vec[VQVEL][ax];
          } else {
            const int qa = ax & 0xff, comp = (ax >> 8) & 0xff, da = (ax >> 16) & 0xff;
            float qq[4] = {S.vec[VQPOS][qa], S.vec[VQPOS][qa + 1], S.vec[VQPOS][qa + 2], S.vec[VQPOS][qa + 3]};
            float w[3] = {S.vec[VQVEL][da], S.vec[VQVEL][da + 1], S.vec[VQVEL][da + 2]};
            float nrm = normalize3(w), sn, cs;
            sincosf(h * nrm * 0.5f, &sn, &cs);
            float qr[4] = {cs, w[0] * sn, w[1] * sn, w[2] * sn}, qn[4];
            quat_mul(qn, qq, qr);
            float n2 = sqrtf(qn[0] * qn[0] + qn[1] * qn[1] + qn[2] * qn[2] + qn[3] * qn[3]);
            float dn = n2 + (n2 == 0.0f ? 1e-6f : 0.0f);
            q = (comp == 0 ? qn[0] : comp == 1 ? qn[1] : comp == 2 ? qn[2] : qn[3]) / dn;
          }
        }
        tm = tm + h;
      } else {
        q = S.vec[VQPOS][lane];
        v = S.vec[VQVEL][lane];
      }
      ws = qacc_solver;
      if (!(fabsf(q) < 3.0e38f) || !(fabsf(v) < 3.0e38f)) status |= MJXB_STATUS_NAN;

      MJXB_STAMP(11);
      // ---------------------------------------------------------------- env layer (src/envs.py) on the forward-pass frames
      if (mode == MODE_ENV_STEP || mode == MODE_ENV_RESET) {
        __syncwarp();
        S.vec[VQPOS][lane] = q;       // post-integration qpos / qvel feed the observation
        S.vec[VQVEL][lane] = (lane < NV) ? v : 0.0f;
        __syncwarp();
        const float bx = S.pelvis_pos[0], by = S.pelvis_pos[1], bz = S.pelvis_pos[2];
        const float hx = S.head_pos[0], hy = S.head_pos[1];
        const float qw = S.pelvis_quat[0], qx = S.pelvis_quat[1], qy = S.pelvis_quat[2], qz = S.pelvis_quat[3];
        const float roll = atan2f(2.0f * (qw * qx + qy * qz), 1.0f - 2.0f * (qx * qx + qy * qy));
        const float pitch = asinf(clampf(2.0f * (qw * qy - qz * qx), -1.0f, 1.0f));
        const float yaw = atan2f(2.0f * (qw * qz + qx * qy), 1.0f - 2.0f * (qy * qy + qz * qz));
        const float sr = S.sens[cfg.touch_sensor_right_id], sl = S.sens[cfg.touch_sensor_left_id];
        const bool rcon = sr > 0.0f, lcon = sl > 0.0f;
        const float new_stance = (rcon && lcon) ? 0.0f : (rcon && !lcon) ? 1.0f : (!rcon && lcon) ? 2.0f : 3.0f;
        float flip, tx, ty, tz, close_count, stance, stance_time, last_pot, ep, dist_obs, dxo, dyo;
        if (mode == MODE_ENV_STEP) {
          aux = (lane < MJXB_AUX_DIM) ? A.in.aux[(size_t)env * MJXB_AUX_DIM + lane] : 0.0f;  // loaded late: nothing has overwritten it yet
          float qfrc_act = 0.0f;   // gear * clip(ctrl), recomputed rather than kept alive through the solver
          if (lane < NV && C.dof_act[lane] >= 0) qfrc_act = C.dof_gear[lane] * clampf(S.vec[VCTRL][C.dof_act[lane]], C.dof_ctrl_lo[lane], C.dof_ctrl_hi[lane]);
          flip = __shfl_sync(FULL, aux, 0); tx = __shfl_sync(FULL, aux, 1); ty = __shfl_sync(FULL, aux, 2); tz = __shfl_sync(FULL, aux, 3);
          close_count = __shfl_sync(FULL, aux, 4); stance = __shfl_sync(FULL, aux, 5); stance_time = __shfl_sync(FULL, aux, 6);
          last_pot = __shfl_sync(FULL, aux, 7); ep = __shfl_sync(FULL, aux, 8);
          const float dx_p = tx - bx, dy_p = ty - by, dx_h = tx - hx, dy_h = ty - hy;
          const float dist = fmaxf(sqrtf(dx_p * dx_p + dy_p * dy_p), sqrtf(dx_h * dx_h + dy_h * dy_h));
          const float progress = (-dist / h - last_pot) * cfg.progress_weight;
          const int nj = C.nv - 6;
          const float pw = warp_sum((lane >= 6 && lane < NV) ? fabsf(qfrc_act * v) : 0.0f);
          const float st2 = warp_sum((lane >= 6 && lane < NV) ? qfrc_act * qfrc_act : 0.0f);
          const float energy = cfg.electricity_cost * (pw / (float)nj) + cfg.stall_torque_cost * (st2 / (float)nj);
          const bool p_ok = (pitch > -0.087f) && (pitch < 0.174f), r_ok = (roll > -0.174f) && (roll < 0.174f);
          const float posture = ((p_ok ? 0.0f : fabsf(pitch)) + (r_ok ? 0.0f : fabsf(roll))) * cfg.posture_penalty_weight;
          const float tall = cfg.tall_bonus_weight * (bz > cfg.tall_height_threshold ? 1.0f : -1.0f);
          const bool changed = new_stance != stance;
          const float duration = tm - stance_time;
          const float stance_reward = (changed && duration > 0.1f) ? cfg.stance_time_reward_weight * duration / h : 0.0f;
          stance = changed ? new_stance : stance;
          stance_time = changed ? tm : stance_time;
          const bool is_close = dist < cfg.target_threshold;
          close_count = is_close ? close_count + 1.0f : 0.0f;
          const float target_bonus = is_close ? 2.0f : 0.0f;
          if (close_count >= (float)cfg.stop_frames) { tx = bx + cfg.target_dist; ty = by; tz = bz; close_count = 0.0f; }
          dxo = tx - bx; dyo = ty - by;
          const float dx_h2 = tx - hx, dy_h2 = ty - hy;
          dist_obs = fmaxf(sqrtf(dxo * dxo + dyo * dyo), sqrtf(dx_h2 * dx_h2 + dy_h2 * dy_h2));
          last_pot = -dist_obs / h;
          float reward = progress + target_bonus + stance_reward - energy + tall - posture - 0.0f;
          ep = ep + 1.0f;
          const bool fallen = bz < cfg.terminate_height;
          out_term = fallen ? 1.0f : 0.0f;
          out_trunc = (cfg.max_episode_steps > 0 && ep >= (float)cfg.max_episode_steps) ? 1.0f : 0.0f;
          if (fallen) reward = reward + cfg.terminate_reward;
          out_reward = reward;
        } else {  // reset (src/envs.py:136-174)
          flip = flip_r;
          tx = bx + cfg.target_dist; ty = by; tz = bz;
          dxo = tx - bx; dyo = ty - by;
          const float dx_h = tx - hx, dy_h = ty - hy;
          dist_obs = fmaxf(sqrtf(dxo * dxo + dyo * dyo), sqrtf(dx_h * dx_h + dy_h * dy_h));
          last_pot = -dist_obs / h;
          close_count = 0.0f; stance = new_stance; stance_time = tm; ep = 0.0f;
        }
        tgt_x = tx; tgt_y = ty; tgt_z = tz;
        // aux' (src/envs.py:174,484)
        aux = 0.0f;
        if (lane == 0) aux = flip; if (lane == 1) aux = tx; if (lane == 2) aux = ty; if (lane == 3) aux = tz;
        if (lane == 4) aux = close_count; if (lane == 5) aux = stance; if (lane == 6) aux = stance_time;
        if (lane == 7) aux = last_pot; if (lane == 8) aux = ep;

        const bool done_env = (mode == MODE_ENV_STEP) && (fmaxf(out_term, out_trunc) > 0.0f);
        if (done_env && A.autoreset) {  // train_ppo.py:150-161 fused
          did_reset = 1;
          if (defer || SINGLE) {
            deferred = true;             // queued below; the reset phase of this launch re-initialises it
          } else {
            mode = MODE_ENV_RESET;       // inline: re-run the pipeline in reset mode for this env
            continue;
          }
        }
        // observation (src/envs.py:274-331): [height, rpy, qpos[7:], R^T v_lin, R^T v_ang, qvel[6:], tgt]
        const float angle = atan2f(dyo, dxo) - yaw;
        const float soft = dist_obs / (1.0f + fabsf(dist_obs));
        const float tg0 = soft * sinf(angle), tg1 = soft * cosf(angle);
        const float xx = qx * qx, yy = qy * qy, zz = qz * qz, xy = qx * qy, xz = qx * qz, yz = qy * qz, wx = qw * qx, wy = qw * qy, wz = qw * qz;
        const float r00 = 1.0f - 2.0f * (yy + zz), r01 = 2.0f * (xy - wz), r02 = 2.0f * (xz + wy);
        const float r10 = 2.0f * (xy + wz), r11 = 1.0f - 2.0f * (xx + zz), r12 = 2.0f * (yz - wx);
        const float r20 = 2.0f * (xz - wy), r21 = 2.0f * (yz + wx), r22 = 1.0f - 2.0f * (xx + yy);
        const int nqj = C.nq - 7, od = cfg.obs_dim;
        for (int o = lane; o < od; o += 32) {
          const int src = flip > 0.5f ? cfg.obs_perm[o] : o;
          float val;
          if (src == 0) val = bz;
          else if (src == 1) val = roll;
          else if (src == 2) val = pitch;
          else if (src == 3) val = yaw;
          else if (src < 4 + nqj) val = S.vec[VQPOS][7 + (src - 4)];
          else if (src < 4 + nqj + 6) {
            const int k = src - 4 - nqj, gsel = k / 3, c = k % 3;
            const float lx = S.vec[VQVEL][3 * gsel], ly = S.vec[VQVEL][3 * gsel + 1], lz = S.vec[VQVEL][3 * gsel + 2];
            val = c == 0 ? r00 * lx + r10 * ly + r20 * lz : c == 1 ? r01 * lx + r11 * ly + r21 * lz : r02 * lx + r12 * ly + r22 * lz;
          } else if (src < 4 + nqj + C.nv) val = S.vec[VQVEL][6 + (src - 4 - nqj - 6)];
          else val = (src == od - 2) ? tg0 : tg1;
          if (flip > 0.5f) val *= cfg.obs_sign[o];
          if (spec_partner) S.J[o] = val;     // parked (J is dead by now); stored after the partner's verdict
          else if (valid && !overflow && !deferred) A.obs[(size_t)env * od + o] = val;
        }
      }
      break;
      }  // pass
      (void)tgt_x; (void)tgt_y; (void)tgt_z;
    }  // nsteps
    if (spec && !reset_phase) {   // the stepping warp tells the reset warp that served it whether the episode really ended (it does unless the env overflowed)
      if (!spec_partner && lane == 0) s_spec_done[wslot] = (valid && !overflow && deferred) ? 1 : 0;
      __syncthreads();
      if (threadIdx.x == 0) { s_pub_count = 0; s_need_count = 0; }   // re-armed for the next step round (its round-top barrier orders this)
      if (spec_partner) {
        const bool take = valid && !overflow && s_spec_done[wslot] != 0;
        if (take) {
          const int od = cfg.obs_dim;
          for (int o = lane; o < od; o += 32) A.obs[(size_t)env * od + o] = S.J[o];
          if (lane < C.nq) A.out.qpos[(size_t)env * C.nq + lane] = q;
          if (lane < NV) { A.out.qvel[(size_t)env * NV + lane] = v; A.out.qacc_warmstart[(size_t)env * NV + lane] = ws; }
          if (lane == 0) A.out.time[env] = tm;
          if (lane < MJXB_AUX_DIM) A.out.aux[(size_t)env * MJXB_AUX_DIM + lane] = aux;
        }
        __syncwarp();
        continue;
      }
    }
    if (!valid) continue;
    if (overflow) {  // leave the env untouched; the big-capacity pass redoes it from its inputs
      if (A.out_list != nullptr) {
        if (lane == 0) { const int slot = atomicAdd(A.out_count, 1); A.out_list[slot] = env; }
      } else if (A.status && lane == 0) {
        A.status[env] = status | MJXB_STATUS_ROW_SPILL;
      }
      __syncwarp();
      continue;
    }

      MJXB_STAMP(12);
    // ------------------------------------------------------------------ store
    if (deferred) {  // step results now, state / obs / aux from the reset phase
      if (lane == 0) {
        A.reward[env] = out_reward; A.terminated[env] = out_term; A.truncated[env] = out_trunc;
        if (A.reset_mask) A.reset_mask[env] = 1;
        if (A.status) A.status[env] = status;
        if (!spec || s_spec_need[wslot] != 1) {   // (concurrent-reset rounds: only an env that no reset warp could serve)
          const int slot = atomicAdd(&s_reset_count, 1);
          A.reset_list[(size_t)blockIdx.x * A.reset_stride + slot] = env;
        }
      }
      __syncwarp();
      continue;
    }
    if (A.mode == MODE_SPEED_TEST) {
      if (lane == 0) A.pos[env] = q;
    } else {
      if (lane < C.nq) A.out.qpos[(size_t)env * C.nq + lane] = q;
      if (lane < NV) { A.out.qvel[(size_t)env * NV + lane] = v; A.out.qacc_warmstart[(size_t)env * NV + lane] = ws; }
      if (lane == 0) A.out.time[env] = tm;
      if (A.mode == MODE_ENV_STEP || A.mode == MODE_ENV_RESET) {
        if (lane < MJXB_AUX_DIM) A.out.aux[(size_t)env * MJXB_AUX_DIM + lane] = aux;
      }
      if (A.mode == MODE_ENV_STEP && !reset_phase && lane == 0) {
        A.reward[env] = out_reward; A.terminated[env] = out_term; A.truncated[env] = out_trunc;
        if (A.reset_mask) A.reset_mask[env] = (uint8_t)did_reset;
      }
    }
    if (A.status && lane == 0) A.status[env] = reset_phase ? (A.status[env] | status) : status;
    MJXB_STAMP(13);
    __syncwarp();
  }
  if (dyn) {  // last CTA out re-arms the group counter for the next launch
    __syncthreads();
    if (threadIdx.x == 0) {
      __threadfence();
      const int t = atomicAdd(A.dyn_done, 1);
      if (t == (int)gridDim.x - 1) { *A.dyn_counter = 0; *A.dyn_done = 0; __threadfence(); }
    }
  }
  if (consuming) {  // last CTA out resets the consumed list's counters for the next step
    __syncthreads();
    if (threadIdx.x == 0) {
      __threadfence();
      const int t = atomicAdd(A.in_done, 1);
      if (t == (int)gridDim.x - 1) { *A.in_count = 0; *A.in_done = 0; __threadfence(); }
    }
  }
}

}  // namespace mjxb
