// capi_util.h — helpers shared by the extern "C" translation units (capi.cu, pipeline.cu).
#pragma once
#include <cuda_runtime.h>

#include "kernels.h"

int pob_fail(const char* where, const char* what);
int pob_cuda_fail(const char* where, cudaError_t e);
int pob_sm_count_cached();
int pob_check_common(const char* where, const void* packed, int sh_deg, int precision);
pob::FwdParams pob_base_params(const void* packed, int sh_deg);

#define POB_CUDA(where, call)                               \
  do {                                                      \
    cudaError_t _e = (call);                                \
    if (_e != cudaSuccess) return pob_cuda_fail(where, _e); \
  } while (0)

// ---- instrumentation (bench.py): kernel launch counter and per-phase CUDA-event timing ---------
enum PobPhase { POB_PH_FWD = 0, POB_PH_BWD, POB_PH_WGRAD, POB_PH_RENDER, POB_PH_OPTIM, POB_PH_COUNT };
void pob_count_launch(int n = 1);
// records an event pair around [begin, end) of one kernel launch when timing is enabled
struct PobPhaseTimer {
  PobPhaseTimer(int phase, cudaStream_t st);
  ~PobPhaseTimer();
  int slot;
  cudaStream_t st;
};
