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
