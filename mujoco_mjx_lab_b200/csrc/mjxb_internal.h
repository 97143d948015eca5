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

}  // namespace mjxb
