// Library-level entry points: version, status strings, error bookkeeping.
#include <string.h>

#include "common.cuh"

namespace cmr {
namespace {
thread_local char g_last_error[512] = "";
}

void record_cuda_error(cudaError_t e, const char* file, int line) {
  snprintf(g_last_error, sizeof(g_last_error), "%s (%s) at %s:%d", cudaGetErrorName(e),
           cudaGetErrorString(e), file, line);
}

int sm_count() {
  static int cached = 0;
  if (cached == 0) {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) == cudaSuccess &&
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
      cached = n;
    else
      cached = 148;
  }
  return cached;
}
}  // namespace cmr

extern "C" const char* cmr_status_string(int status) {
  switch (status) {
    case CMR_OK: return "ok";
    case CMR_ERR_INVALID_ARG: return "invalid argument";
    case CMR_ERR_CUDA: return "CUDA error";
    case CMR_ERR_WORKSPACE: return "workspace too small";
    case CMR_ERR_UNSUPPORTED: return "unsupported shape";
    default: return "unknown status";
  }
}

extern "C" int cmr_version(void) { return 1; }

extern "C" const char* cmr_last_cuda_error(void) { return cmr::g_last_error; }
