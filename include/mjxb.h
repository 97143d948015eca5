/* mjxb -- C ABI of the B200-native batched humanoid physics step.
 *
 * This is the drop-in boundary for the hot path that the reference drives through
 *   src/envs.py:494-495   v_reset = jit(vmap(single_reset)); v_step = jit(vmap(single_step))
 *   src/envs.py:108-113   single_pipeline_init  (mjx.make_data -> mjx.forward)
 *   src/envs.py:345       mjx.step(sys, d)
 *   mjx_humanoid_speed_test.py:48-57,88-93   vmap(make_data -> qvel[0]=v -> mjx.step -> qpos[0])
 * Nothing like it exists in the reference (pure Python on JAX/MJX); each entry point below names the
 * reference function it replaces.  INTEGRATION.md shows the XLA-FFI / ctypes binding a maintainer adds.
 *
 * Conventions
 *   - plain pointers and sizes only; no torch / jax types.
 *   - every array argument of the *device* entry points is a CUDA device pointer owned by the caller,
 *     row-major [n_env, k] float32 ("structure of arrays": one array per state field);
 *     all work is enqueued on the caller's stream (`stream` is a cudaStream_t passed as void*);
 *     no host synchronisation inside reset/step. The only allocation is the per-stream launch scratch (overflow lists, 12 B / env),
 *     made the first time a stream launches a batch larger than it has seen; mjxb_model_reserve(m, n_env, stream) makes it
 *     ahead of time -- mandatory before capturing the stream into a CUDA graph. A scratch buffer is never freed or moved before
 *     mjxb_model_destroy, so captured graphs stay valid.
 *   - the *_host entry points take host pointers and do H2D -> kernel -> D2H themselves (synchronous).
 *   - return value: 0 = ok, negative = MJXB_E* (argument / CUDA errors, reported synchronously).
 *     Numerical trouble is reported asynchronously per env through `status` (bit flags below).
 *   - the model constants are immutable after create; one model per device. Launches on distinct streams use distinct scratch
 *     and may run concurrently from different threads; launches on ONE stream are ordered by the stream.
 */
#ifndef MJXB_H_
#define MJXB_H_

#include <stddef.h>
#include <stdint.h>
#include "mjxb_model.h"

#ifdef __cplusplus
extern "C" {
#endif

#define MJXB_ABI_VERSION 1

#define MJXB_OK 0
#define MJXB_EINVAL (-1)   /* bad argument (NULL pointer, n_env <= 0, size mismatch) */
#define MJXB_EBLOB (-2)    /* model blob has wrong magic / version / size or exceeds compiled capacities */
#define MJXB_ECUDA (-3)    /* CUDA runtime error (mjxb_last_cuda_error() has the text) */
#define MJXB_ENOGPU (-4)   /* no CUDA device: there is deliberately no CPU fallback */
#define MJXB_EUNSUPPORTED (-5)

/* per-env status bits (int32 status[n_env], optional) */
#define MJXB_STATUS_NAN 1            /* non-finite qacc / state produced */
#define MJXB_STATUS_ROW_SPILL 2      /* candidate constraint rows exceeded the shared-memory tile; global scratch used */
#define MJXB_STATUS_MAXITER 4        /* Newton/CG hit opt.iterations without meeting the tolerance */

#define MJXB_AUX_DIM 9    /* src/envs.py:15  [flip,tx,ty,tz,close_count,stance_state,stance_last_change_time,last_pot,episode_step] */
#define MJXB_MAXOBS 64

/* EnvConfig (src/config.py:29-66) + the flip permutations create_env_functions derives (src/envs.py:49-74). */
typedef struct mjxb_env_config {
  float progress_weight, electricity_cost, stall_torque_cost, posture_penalty_weight;
  float tall_height_threshold, tall_bonus_weight, target_threshold, target_dist;
  float stance_time_reward_weight, random_joint_noise, random_vel_noise, initial_velocity_max;
  float terminate_height, terminate_reward;
  int32_t stop_frames, max_episode_steps, random_flip;
  int32_t pelvis_body_id, head_body_id, touch_sensor_right_id, touch_sensor_left_id;
  int32_t obs_dim;
  int32_t act_perm[MJXB_MAXU];
  float act_sign[MJXB_MAXU];
  int32_t obs_perm[MJXB_MAXOBS];
  float obs_sign[MJXB_MAXOBS];
} mjxb_env_config;

/* The persistent per-env state: the minimal replacement of `EnvState = (mjx.Data, AuxState)` (src/envs.py:13-17).
 * Everything else the env reads from mjx.Data (xpos, xquat, sensordata, qfrc_actuator) is recomputed in the step. */
typedef struct mjxb_state {
  float* qpos;           /* [n_env, nq]  */
  float* qvel;           /* [n_env, nv]  */
  float* qacc_warmstart; /* [n_env, nv]  */
  float* time;           /* [n_env]      */
  float* aux;            /* [n_env, 9]   */
} mjxb_state;

/* Optional per-stage outputs for parity tests (any pointer may be NULL). Shapes use the model's static sizes. */
typedef struct mjxb_debug {
  float* xpos;            /* [n, nbody, 3] */
  float* xquat;           /* [n, nbody, 4] */
  float* qM;              /* [n, nv, nv]   dense symmetric */
  float* qfrc_bias;       /* [n, nv] */
  float* qfrc_passive;    /* [n, nv] */
  float* qfrc_actuator;   /* [n, nv] */
  float* qacc_smooth;     /* [n, nv] */
  float* con_dist;        /* [n, ncon] */
  float* con_pos;         /* [n, ncon, 3] */
  float* con_normal;      /* [n, ncon, 3] */
  float* efc_pos;         /* [n, nefc]  (0 for rows that are not candidates, as MJX masks them) */
  float* efc_D;           /* [n, nefc] */
  float* efc_aref;        /* [n, nefc] */
  float* efc_force;       /* [n, nefc] */
  int32_t* efc_active;    /* [n, nefc]  bit0: candidate (pos<0); bit1: active at the solution (Jaref<0) */
  float* qacc;            /* [n, nv] */
  float* qfrc_constraint; /* [n, nv] */
  float* sensordata;      /* [n, nsensor] */
  int32_t* solver_niter;  /* [n] */
} mjxb_debug;

typedef struct mjxb_model mjxb_model;

int mjxb_abi_version(void);
/* number of kernels this library has launched in the process so far (a CUDA-graph replay of captured launches is not counted) */
long long mjxb_launch_count(void);
size_t mjxb_blob_sizeof(void);
size_t mjxb_env_config_sizeof(void);
const char* mjxb_strerror(int code);
const char* mjxb_last_cuda_error(void);

/* replaces mjx.put_model(m) + create_env_functions' closure over (sys, cfg, q0) (src/training_utils.py:105-112).
 * `blob` is the POD produced by modelc.pack_blob (struct mjxb_model_blob); copied to `device`. */
int mjxb_model_create(const void* blob, size_t blob_bytes, const mjxb_env_config* cfg, int device, mjxb_model** out);
/* the same with explicit option flags (mjxb_model_create derives them from MJXB_LS_ITERATIVE / MJXB_DENSE_CHOL / MJXB_INLINE_RESET in
 * the environment, read once at create; nothing on the launch path reads the environment) */
#define MJXB_FLAG_LS_ITERATIVE 1u  /* run MJX's bracketed line-search iteration (solver._linesearch) even where ls_iterations >= 10 would
                                      select the closed-form exact minimiser */
#define MJXB_FLAG_DENSE_CHOL 2u    /* dense right-looking Cholesky instead of the generated tree-ordered elimination */
#define MJXB_FLAG_INLINE_RESET 4u  /* auto-reset runs inline instead of in deferred packed rounds */
#define MJXB_FLAG_NO_SPEC_RESET 8u /* never run the auto-reset on reset warps beside the step rounds (batches of up to six rounds, <= 12,432 envs):
                                      always the deferred packed reset rounds after the step rounds */
#define MJXB_FLAG_NO_WORK_SORT 16u /* never deal the envs of a large batch (>= 16,384) to the CTAs in order of their previous step's Newton
                                      iteration count (the work-sorted schedule: one extra small launch per step, identical results) */
#define MJXB_FLAG_NO_DYN_ROUNDS 32u /* static (round, CTA) assignment of the env groups instead of a device-wide group counter (batches of
                                      several rounds per SM; identical results) */
#define MJXB_FLAG_BUILD_EXACT 256u /* (reported only) the library is the reference-arithmetic build: no fast-math, no FMA contraction */
int mjxb_model_create_ex(const void* blob, size_t blob_bytes, const mjxb_env_config* cfg, int device, uint32_t flags,
                         mjxb_model** out);
/* the options in effect for `m` (MJXB_FLAG_* bits) */
int mjxb_model_flags(const mjxb_model* m);
/* make the launch scratch of `stream` large enough for n_env envs now (needed before stream capture; optional otherwise) */
int mjxb_model_reserve(const mjxb_model* m, int32_t n_env, void* stream);
void mjxb_model_destroy(mjxb_model* m);
/* nq, nv, nu, nbody, ncon, nefc, nsensor, obs_dim */
int mjxb_model_dims(const mjxb_model* m, int32_t dims[8]);
/* bytes of launch scratch currently held for `m` over all streams. */
size_t mjxb_model_scratch_bytes(const mjxb_model* m);

/* v_reset (src/envs.py:115-202,494): keys u32[n,2] (JAX threefry key data) -> state, obs[n,obs_dim]. */
int mjxb_reset(const mjxb_model* m, int32_t n_env, const uint32_t* keys, mjxb_state out, float* obs,
               int32_t* status, void* stream);

/* v_step (src/envs.py:333-492,495): (state, action[n,nu]) -> state', obs, reward, terminated, truncated.
 * `in` and `out` may alias field by field (in-place update). */
int mjxb_step(const mjxb_model* m, int32_t n_env, mjxb_state in, const float* action, mjxb_state out,
              float* obs, float* reward, float* terminated, float* truncated, int32_t* status, void* stream);

/* v_step fused with the trainer's auto-reset glue (train_ppo.py:143-161): envs with max(terminated,truncated)>0
 * are re-initialised from keys[n,2] inside the same launch; reward/terminated/truncated are the step's,
 * state/obs are the merged ones; reset_mask[n] (u8, optional) reports which envs were reset. */
int mjxb_step_autoreset(const mjxb_model* m, int32_t n_env, mjxb_state in, const float* action,
                        const uint32_t* keys, mjxb_state out, float* obs, float* reward, float* terminated,
                        float* truncated, uint8_t* reset_mask, int32_t* status, void* stream);

/* mjx.step (src/envs.py:345; mjx_humanoid_speed_test.py:54) without the env layer: nsteps consecutive physics
 * steps with a fixed ctrl[n,nu] (may be NULL = zeros), state advanced in place (aux is ignored).
 * `dbg` (optional) receives the stage outputs of the LAST step's forward pass. */
int mjxb_physics_step(const mjxb_model* m, int32_t n_env, mjxb_state io, const float* ctrl, int32_t nsteps,
                      const mjxb_debug* dbg, int32_t* status, void* stream);

/* mjx.forward (src/envs.py:112): forward dynamics only (no integration); writes qacc_warmstart, dbg. */
int mjxb_forward(const mjxb_model* m, int32_t n_env, mjxb_state io, const float* ctrl, const mjxb_debug* dbg,
                 int32_t* status, void* stream);

/* mjx_humanoid_speed_test.py:48-57,88-93: for it in range(iters): pos = step(make_data with qvel[0]=vel[i]).qpos[0];
 * acc += sum(pos). vel[n] in, pos[n] out (last iteration), device pointers. */
int mjxb_speed_test(const mjxb_model* m, int32_t n_env, const float* vel, float* pos, int32_t iters, void* stream);

/* launch geometry of the step kernel, for reports: cfg = {env-warps per CTA, dynamic shared memory bytes per CTA, SM count,
 * shared memory bytes per env-warp} */
int mjxb_launch_config(const mjxb_model* m, int32_t cfg[4]);

/* Host-buffer variants (pinned or pageable host memory): H2D of the inputs, the launch, D2H of the outputs, sync.
 * The device state stays resident in a library-owned arena bound to `m` (created on first use for n_env); one caller thread
 * per model for these entry points.
 * Pinned (cudaHostAlloc / cudaHostRegister) buffers take the direct pipeline: ONE launch over the batch, action / keys copied in
 * ~8 chunks on a copy stream with a ready flag per chunk that the kernel waits on, obs / reward / terminated / truncated stored
 * by the kernel straight into the caller's buffers (visible when the call returns). Pageable buffers take a chunked three-stream
 * H2D -> launch -> D2H pipeline. Results are identical. A chunk that never arrives (2 s) makes the call return MJXB_ECUDA. */
int mjxb_reset_host(mjxb_model* m, int32_t n_env, const uint32_t* keys_host, float* obs_host);
int mjxb_step_host(mjxb_model* m, int32_t n_env, const float* action_host, float* obs_host, float* reward_host,
                   float* terminated_host, float* truncated_host);
int mjxb_step_autoreset_host(mjxb_model* m, int32_t n_env, const float* action_host, const uint32_t* keys_host,
                             float* obs_host, float* reward_host, float* terminated_host, float* truncated_host);
/* copy the resident arena state to / from host arrays (any pointer may be NULL) */
int mjxb_state_get_host(mjxb_model* m, int32_t n_env, float* qpos, float* qvel, float* qacc_warmstart, float* time,
                        float* aux);
int mjxb_state_set_host(mjxb_model* m, int32_t n_env, const float* qpos, const float* qvel,
                        const float* qacc_warmstart, const float* time, const float* aux);

/* ---- fused policy inference for the rollout loop (reference train_ppo.py:121-126,135-140, src/networks.py:55-61):
 * act = mean(obs_n) + exp(log_std) * eps, logp = Gaussian log-density of act, with obs_n = clip((obs - rms_mean) / sqrt(rms_var + 1e-8), +-10)
 * and mean = the 3 x 256 tanh MLP (obs_dim <= 64, act_dim <= 32). One launch: tcgen05 tensor cores (bf16 operands, fp32 accumulation in
 * TMEM), activations stay on the SM. Weights are given pre-packed by mjxb_policy_pack_weight (bf16, canonical K-major core-matrix layout,
 * zero padded: layer 0 [256 x 64], layers 1-2 [256 x 256], layer 3 [32 x 256]); biases float32. All pointers are device pointers.
 * `mean` and `error_flag` may be NULL; *error_flag is set to 1 if a tensor-core completion was not observed (bounded wait). */
int mjxb_policy_pack_weight(const float* w /*[k, n] row-major, x @ w*/, int32_t k, int32_t n, int32_t k_pad, int32_t n_pad,
                            void* out_bf16 /*[n_pad * k_pad] bf16*/, void* stream);
int mjxb_policy_act(int32_t n_env, int32_t obs_dim, int32_t act_dim, const float* obs, const float* rms_mean, const float* rms_var,
                    const void* const* w_packed /*[4]*/, const float* const* bias /*[4]*/, const float* log_std, const float* eps,
                    float* act, float* logp, float* mean, int32_t* error_flag, void* stream);

/* Learner backward helper for y = tanh(x W + b): dz = dy * (1 - y^2) and db += column sums of dz in one pass over [n, c] row-major
 * arrays (y == NULL: plain linear layer, dz is not written and db += column sums of dy). db must be zeroed by the caller. */
int mjxb_tanh_bwd_colsum(int32_t n, int32_t c, const float* dy, const float* y, float* dz, float* db_zeroed, void* stream);

/* Losses of one PPO minibatch and their gradients in two launches (reference train_ppo.py:204-252): logp = Gaussian log-density of
 * action under (mean, exp(log_std)); ratio = exp(logp - old_logp); adv_n = (adv - mean(adv)) / (std(adv) + 1e-8) over the minibatch;
 * loss = mean(-min(ratio adv_n, clip(ratio, 1 +- clip_eps) adv_n)) - ent_coef * entropy(log_std). Writes d loss / d mean [n, act_dim],
 * d loss / d log_std [act_dim] and the loss (loss_out[0]); scratch4 is 4 floats of device scratch. All pointers are device pointers. */
int mjxb_ppo_loss(int32_t n, int32_t act_dim, const float* mean, const float* log_std, const float* action, const float* old_logp,
                  const float* adv, float clip_eps, float ent_coef, float* scratch4, float* g_mean, float* g_log_std, float* loss_out,
                  void* stream);
/* the same for a mean / g_mean stored with a row stride of ld_mean >= act_dim floats (an output layer padded to a multiple of four
 * columns so that its GEMMs take the aligned tensor-core kernels); the padding columns of g_mean are written as zeros */
int mjxb_ppo_loss_ld(int32_t n, int32_t act_dim, int32_t ld_mean, const float* mean, const float* log_std, const float* action,
                     const float* old_logp, const float* adv, float clip_eps, float ent_coef, float* scratch4, float* g_mean,
                     float* g_log_std, float* loss_out, void* stream);
/* Adam (optax.adam semantics: bias-corrected moments, eps outside the square root) over one flat parameter / gradient buffer of n floats:
 * elements [0, split) use lr0, the rest lr1; grad is multiplied by grad_scale first (1 / world size after an all-reduce sum);
 * *step_dev (device float, 0 at the start) counts the updates, so the call can be replayed from a CUDA graph. */
int mjxb_adam(int32_t n, int32_t split, float* param, const float* grad, float* m, float* v, float* step_dev, float lr0, float lr1,
              float b1, float b2, float eps, float grad_scale, void* stream);

/* ---- the learner's collective, fused with the optimiser, over NVLink peer memory (one process per GPU; reference has no multi-GPU code:
 * SURVEY 8e places the only collective in the gradient all-reduce of train_ppo.py:240-246). Every rank owns a gradient buffer and a
 * flag block exported through CUDA IPC. mjxb_allreduce_adam is ONE kernel per minibatch: cross-GPU barrier (all backward passes
 * complete), element-wise sum of ALL ranks' gradients read straight from peer memory in rank order (identical bits on every rank),
 * division by the world size, Adam on the rank's own parameter copy, cross-GPU barrier (all ranks finished reading). The epoch lives on
 * the device (CUDA-graph replayable); every spin is bounded and reported through mjxb_comm_error, never a hang.
 * Set-up: create on every rank, exchange the 128-byte mjxb_comm_local_handles blobs (e.g. torch.distributed.all_gather_object),
 * connect with all of them in rank order; gradients are accumulated directly in mjxb_comm_grad_buffer. */
typedef struct mjxb_comm mjxb_comm;
int mjxb_comm_create(int32_t rank, int32_t world, int32_t n_floats, mjxb_comm** out);
int mjxb_comm_local_handles(mjxb_comm* c, void* handles_out /* 128 bytes */);
int mjxb_comm_connect(mjxb_comm* c, const void* all_handles /* world x 128 bytes, rank order */);
float* mjxb_comm_grad_buffer(mjxb_comm* c);
int mjxb_comm_error(mjxb_comm* c);
int mjxb_allreduce_adam(mjxb_comm* c, int32_t n, int32_t split, float* param, float* m, float* v, float* step_dev, float lr0,
                        float lr1, float b1, float b2, float eps, void* stream);
void mjxb_comm_destroy(mjxb_comm* c);

/* Generalised advantage estimation over a rollout (reference train_ppo.py:171-202): delta_t = r_t + gamma v_{t+1} (1 - terminated_t) - v_t,
 * adv_t = delta_t + gamma lam (1 - max(terminated_t, truncated_t)) adv_{t+1}, ret_t = adv_t + v_t. reward / terminated / truncated /
 * advantage / ret are [rollout_length, n_env], value is [rollout_length + 1, n_env]; device pointers, one launch. */
int mjxb_gae(int32_t rollout_length, int32_t n_env, const float* reward, const float* value, const float* terminated,
             const float* truncated, float gamma, float lam, float* advantage, float* ret, void* stream);

/* ---- analytic policy gradients (reference train_apg.py:161-209: value_and_grad through lax.scan(jax.checkpoint(v_step))).
 * mjxb_step_fwd_tape: v_step (src/envs.py:333-492, no auto-reset) that also records the tape of the reverse pass: the solver's qacc
 *   tape_qacc[n, nv] (may alias out.qacc_warmstart, which holds the same numbers). Everything else the reverse pass needs is recomputed
 *   from the step's inputs, which the caller keeps (as jax.checkpoint does).
 * mjxb_step_vjp: vector-Jacobian product of that step. Given the step's inputs (`in`, `action`), the tape, and the cotangents of its
 *   outputs -- g_qpos_out[n, nq], g_qvel_out[n, nv], g_aux_out[n, 9] (entries 1,2,3 = target and 7 = last_pot are used), g_reward[n]; any
 *   may be NULL = zero -- it writes the cotangents of its inputs: g_qpos_in[n, nq] (quaternion components included), g_qvel_in[n, nv],
 *   g_aux_in[n, 9], g_action[n, nu]. The constraint solve is differentiated by the implicit function theorem at the converged solution
 *   (active set held fixed); discrete outputs (done flags, stance, target advance) have zero gradient; qacc_warmstart has zero gradient.
 *   in.aux == NULL selects the physics step alone (mjx.step: `action` is ctrl, no env layer, g_aux_* / g_reward ignored).
 *   Envs whose candidate rows exceed the 64-row tile are re-run by a 320-row instantiation inside the same call. */
int mjxb_step_fwd_tape(const mjxb_model* m, int32_t n_env, mjxb_state in, const float* action, mjxb_state out, float* obs, float* reward,
                       float* terminated, float* truncated, float* tape_qacc, int32_t* status, void* stream);
int mjxb_step_vjp(const mjxb_model* m, int32_t n_env, mjxb_state in, const float* action, const float* tape_qacc,
                  const float* g_qpos_out, const float* g_qvel_out, const float* g_aux_out, const float* g_reward, float* g_qpos_in,
                  float* g_qvel_in, float* g_aux_in, float* g_action, int32_t* status, void* stream);

/* Measurement aid (bench.py's roofline_fp32 denominator): the FP32 FMA-pipe throughput of `device`, measured with a kernel of
 * independent FFMA chains and no memory traffic (best of several repetitions; synchronises the device). */
int mjxb_ffma_peak(int32_t device, float* tflops_out, float* ms_out);

#ifdef __cplusplus
}
#endif
#endif /* MJXB_H_ */
