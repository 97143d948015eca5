// mjxb_internal.h -- what the translation units of libmjxb.so share besides include/mjxb.h (not part of the ABI).
#pragma once
#include <cuda_runtime.h>

#include <atomic>

#include "mjxb.h"
#include "mjxb_model_dev.h"

extern std::atomic<long long> g_mjxb_launches;   // every kernel launch of the library is counted (mjxb_launch_count)

namespace mjxb {

struct ModelView {
  const DevModel* host;      // host copy of the device model record
  const DevModel* dev;       // device copy
  const PairParam* dev_pp;   // per-pair contact parameters (device)
  int device, num_sms;
};
int model_view(const mjxb_model* m, ModelView* out);
// the per-stream launch scratch of `m` (mjxb_abi.cu): [0] countA, [1] doneA, [2] countB, [3] doneB, [4 .. 4 + cap) listA, ...
int model_scratch(const mjxb_model* m, cudaStream_t stream, int n_env, int** buf, int* cap);
int report_cuda_error(cudaError_t e, const char* what);   // records the text for mjxb_last_cuda_error, returns MJXB_ECUDA

// Programmatic dependent launch (sm_90+): a kernel launched with the attribute may be staged while its predecessor in the stream still
// runs; every kernel of this library starts with pdl_prologue(), which lets ITS successor be staged and then waits until the
// predecessor grid has completed and its writes are visible. Between the dependent kernels of one env step (main tier -> overflow
// tiers -> policy -> next step) this removes the ~5 us launch gap per dependency. MJXB_PDL=0 disables it (read once).
bool pdl_enabled();
template <class Kernel, class... Args>
inline cudaError_t launch_pdl(Kernel kernel, dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, args...);
}
#if defined(__CUDACC__)
__device__ __forceinline__ void pdl_prologue() {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
}
#endif

}  // namespace mjxb
