// mjxb_ffi.cc -- XLA FFI custom-call adapter over the C ABI (include/mjxb.h): what the reference's JAX code binds so that
// v_reset / v_step stay callable under jit / lax.scan (reference src/envs.py:494-497, train_ppo.py:143,166-168) and, for APG,
// differentiable through jax.custom_vjp (reference train_apg.py:161-209).  Zero-copy: XLA hands over device buffers, the handlers pass
// their pointers and XLA's stream straight to libmjxb.so; state outputs may alias state inputs (input_output_aliases).
//
// Built only where jaxlib's headers exist:   make -C mujoco_mjx_lab_b200/csrc ffi
//     g++ -std=c++17 -shared -fPIC -I$(python -c "import jax.ffi; print(jax.ffi.include_dir())") -I../../include mjxb_ffi.cc \
//         -L.. -lmjxb -Wl,-rpath,'$ORIGIN' -o ../libmjxb_ffi.so
// jax / jaxlib are NOT installed in this image, so this file is compiled and tested nowhere here (tests/test_jax_ffi.py skips with that
// reason); the Python side is mujoco_mjx_lab_b200/jax_ffi.py.
#include <cstdint>

#include "mjxb.h"
#include "xla/ffi/api/ffi.h"

namespace ffi = xla::ffi;
using F32 = ffi::Buffer<ffi::F32>;
using U32 = ffi::Buffer<ffi::U32>;
using RF32 = ffi::ResultBuffer<ffi::F32>;
typedef struct CUstream_st* cudaStream_t;

static const mjxb_model* model_of(int64_t handle) { return reinterpret_cast<const mjxb_model*>(static_cast<intptr_t>(handle)); }
static ffi::Error status_of(int rc) {
  if (rc == MJXB_OK) return ffi::Error::Success();
  return ffi::Error::Internal(rc == MJXB_ECUDA ? mjxb_last_cuda_error() : mjxb_strerror(rc));
}
static int32_t batch_of(const F32& qpos) { return static_cast<int32_t>(qpos.dimensions()[0]); }

// v_reset(keys) -> (qpos, qvel, qacc_warmstart, time, aux), obs
static ffi::Error ResetImpl(cudaStream_t stream, int64_t model_handle, U32 keys, RF32 qpos, RF32 qvel, RF32 warm, RF32 time, RF32 aux, RF32 obs) {
  mjxb_state out{qpos->typed_data(), qvel->typed_data(), warm->typed_data(), time->typed_data(), aux->typed_data()};
  const int32_t n = static_cast<int32_t>(keys.dimensions()[0]);
  return status_of(mjxb_reset(model_of(model_handle), n, keys.typed_data(), out, obs->typed_data(), nullptr, stream));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(MjxbReset, ResetImpl,
                              ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>().Attr<int64_t>("model_handle").Arg<U32>()
                                  .Ret<F32>().Ret<F32>().Ret<F32>().Ret<F32>().Ret<F32>().Ret<F32>());

// v_step(state, action) -> state', obs, reward, terminated, truncated        (state' may alias state)
static ffi::Error StepImpl(cudaStream_t stream, int64_t model_handle, F32 qpos, F32 qvel, F32 warm, F32 time, F32 aux, F32 action,
                           RF32 qpos_o, RF32 qvel_o, RF32 warm_o, RF32 time_o, RF32 aux_o, RF32 obs, RF32 reward, RF32 terminated, RF32 truncated) {
  mjxb_state in{qpos.typed_data(), qvel.typed_data(), warm.typed_data(), time.typed_data(), aux.typed_data()};
  mjxb_state out{qpos_o->typed_data(), qvel_o->typed_data(), warm_o->typed_data(), time_o->typed_data(), aux_o->typed_data()};
  return status_of(mjxb_step(model_of(model_handle), batch_of(qpos), in, action.typed_data(), out, obs->typed_data(), reward->typed_data(),
                             terminated->typed_data(), truncated->typed_data(), nullptr, stream));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(MjxbStep, StepImpl,
                              ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>().Attr<int64_t>("model_handle")
                                  .Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>()
                                  .Ret<F32>().Ret<F32>().Ret<F32>().Ret<F32>().Ret<F32>().Ret<F32>().Ret<F32>().Ret<F32>().Ret<F32>());

// v_step fused with the trainer's reset-and-merge (train_ppo.py:143-161)
static ffi::Error StepAutoresetImpl(cudaStream_t stream, int64_t model_handle, F32 qpos, F32 qvel, F32 warm, F32 time, F32 aux, F32 action,
                                    U32 keys, RF32 qpos_o, RF32 qvel_o, RF32 warm_o, RF32 time_o, RF32 aux_o, RF32 obs, RF32 reward,
                                    RF32 terminated, RF32 truncated) {
  mjxb_state in{qpos.typed_data(), qvel.typed_data(), warm.typed_data(), time.typed_data(), aux.typed_data()};
  mjxb_state out{qpos_o->typed_data(), qvel_o->typed_data(), warm_o->typed_data(), time_o->typed_data(), aux_o->typed_data()};
  return status_of(mjxb_step_autoreset(model_of(model_handle), batch_of(qpos), in, action.typed_data(), keys.typed_data(), out, obs->typed_data(),
                                       reward->typed_data(), terminated->typed_data(), truncated->typed_data(), nullptr, nullptr, stream));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(MjxbStepAutoreset, StepAutoresetImpl,
                              ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>().Attr<int64_t>("model_handle")
                                  .Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>().Arg<U32>()
                                  .Ret<F32>().Ret<F32>().Ret<F32>().Ret<F32>().Ret<F32>().Ret<F32>().Ret<F32>().Ret<F32>().Ret<F32>());

// reverse mode of v_step (custom_vjp bwd): inputs + tape (= qacc_warmstart of the step's output) + output cotangents -> input cotangents
static ffi::Error StepVjpImpl(cudaStream_t stream, int64_t model_handle, F32 qpos, F32 qvel, F32 warm, F32 time, F32 aux, F32 action,
                              F32 tape_qacc, F32 g_qpos_o, F32 g_qvel_o, F32 g_aux_o, F32 g_reward, RF32 g_qpos, RF32 g_qvel, RF32 g_aux,
                              RF32 g_action) {
  mjxb_state in{qpos.typed_data(), qvel.typed_data(), warm.typed_data(), time.typed_data(), aux.typed_data()};
  return status_of(mjxb_step_vjp(model_of(model_handle), batch_of(qpos), in, action.typed_data(), tape_qacc.typed_data(), g_qpos_o.typed_data(),
                                 g_qvel_o.typed_data(), g_aux_o.typed_data(), g_reward.typed_data(), g_qpos->typed_data(), g_qvel->typed_data(),
                                 g_aux->typed_data(), g_action->typed_data(), nullptr, stream));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(MjxbStepVjp, StepVjpImpl,
                              ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>().Attr<int64_t>("model_handle")
                                  .Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>()
                                  .Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>()
                                  .Ret<F32>().Ret<F32>().Ret<F32>().Ret<F32>());
