// Shared helpers for the cmr_b200 CUDA sources (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/cmr_b200.h"

namespace cmr {

// Records the last CUDA error text for cmr_last_cuda_error().
void record_cuda_error(cudaError_t e, const char* file, int line);

#define CMR_CUDA_TRY(expr)                                   \
  do {                                                       \
    cudaError_t _e = (expr);                                 \
    if (_e != cudaSuccess) {                                 \
      ::cmr::record_cuda_error(_e, __FILE__, __LINE__);      \
      return CMR_ERR_CUDA;                                   \
    }                                                        \
  } while (0)

// Every kernel launch of the library passes through here: counts it (cmr_launch_count)
// and surfaces launch errors.
void count_launch();
#define CMR_LAUNCH_CHECK()               \
  do {                                   \
    ::cmr::count_launch();               \
    CMR_CUDA_TRY(cudaGetLastError());    \
  } while (0)

// Optional per-launch timing of the tensor-core kernels (cmr_prof_enable): a pair of
// CUDA events on the launching stream around the launch, plus the launch's algorithmic
// work (FLOPs; algorithmic bytes for ROIAlign), summed by cmr_prof_collect.  Skipped while a stream is being captured.
enum ProfKind {
  kProfConvGemm = 0, kProfWgrad = 1, kProfRoiAlign = 2, kProfRoiAlignBwd = 3,
  // secondary views of the kProfConvGemm launches, split by what bounds them (prof_tag)
  kProfConvTensorBound = 4,   // work = FLOPs
  kProfConvHbmBound = 5,      // work = algorithmic bytes
  // the reference-layout ROIAlign operator (cmr_roi_align_fwd_cl / _bwd_cl), bytes
  kProfRoiAlignApi = 6, kProfRoiAlignApiBwd = 7,
  kProfKinds = 8
};
void prof_begin(int kind, double work, cudaStream_t st);
void prof_tag(int kind2, double work2);   // between prof_begin and prof_end
void prof_end(cudaStream_t st);

#define CMR_REQUIRE(cond)                 \
  do {                                    \
    if (!(cond)) return CMR_ERR_INVALID_ARG; \
  } while (0)

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

// Number of SMs of the current device (148 on B200), cached per process.
int sm_count();

// Deterministic mode: floating-point reductions whose arrival order is not fixed (split
// weight gradients, ROIAlign backward, bias column sums) accumulate into 64-bit FIXED-POINT
// words instead -- integer addition is associative, so the sum does not depend on the order.
// One unit = 2^-40: +-8.4e6 of range, 9e-13 of resolution (a single fp32 term is converted
// exactly up to that resolution).
#ifdef __CUDACC__
__device__ __forceinline__ long long to_fixed(float v) {
  return __double2ll_rn((double)v * 1099511627776.0);
}
__device__ __forceinline__ void red_fixed(long long* addr, float v) {
  atomicAdd(reinterpret_cast<unsigned long long*>(addr), (unsigned long long)to_fixed(v));
}
#endif

__host__ __device__ inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
__host__ __device__ inline long long ceil_div_ll(long long a, long long b) {
  return (a + b - 1) / b;
}

}  // namespace cmr
