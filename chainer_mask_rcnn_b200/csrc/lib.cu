// Library-level entry points: version, status strings, error bookkeeping.
#include <string.h>

#include <atomic>
#include <mutex>
#include <vector>

#include "common.cuh"

namespace cmr {
namespace {
thread_local char g_last_error[512] = "";
std::atomic<long long> g_launches{0};

struct ProfRec {
  cudaEvent_t a, b;
  int kind;
  double work;
  int kind2;      // optional second classification of the same launch (prof_tag), -1 = none
  double work2;
};
std::mutex g_prof_mu;
bool g_prof_on = false;
std::vector<ProfRec> g_recs;       // completed (a, b) pairs awaiting collection
std::vector<cudaEvent_t> g_free;   // recycled events
bool g_open = false;               // a prof_begin without its prof_end yet
ProfRec g_cur;

cudaEvent_t take_event() {
  if (!g_free.empty()) {
    cudaEvent_t e = g_free.back();
    g_free.pop_back();
    return e;
  }
  cudaEvent_t e = nullptr;
  cudaEventCreate(&e);
  return e;
}
}  // namespace

void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

void prof_begin(int kind, double work, cudaStream_t st) {
  if (!g_prof_on) return;
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(st, &cs) != cudaSuccess || cs != cudaStreamCaptureStatusNone) return;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  g_cur.a = take_event();
  g_cur.b = take_event();
  g_cur.kind = kind;
  g_cur.work = work;
  g_cur.kind2 = -1;
  g_cur.work2 = 0.0;
  if (!g_cur.a || !g_cur.b) return;
  cudaEventRecord(g_cur.a, st);
  g_open = true;
}

void prof_tag(int kind2, double work2) {
  if (!g_prof_on) return;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  if (!g_open) return;
  g_cur.kind2 = kind2;
  g_cur.work2 = work2;
}

void prof_end(cudaStream_t st) {
  if (!g_prof_on) return;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  if (!g_open) return;
  cudaEventRecord(g_cur.b, st);
  g_recs.push_back(g_cur);
  g_open = false;
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

extern "C" int cmr_version(void) { return 5; }

extern "C" long long cmr_launch_count(void) { return cmr::g_launches.load(); }

extern "C" int cmr_prof_enable(int on) {
  std::lock_guard<std::mutex> lk(cmr::g_prof_mu);
  cmr::g_prof_on = on != 0;
  return CMR_OK;
}

extern "C" int cmr_prof_collect(int kind, double* total_ms, double* total_work,
                                long long* launches) {
  CMR_REQUIRE(kind >= 0 && kind < cmr::kProfKinds && total_ms && total_work && launches);
  std::lock_guard<std::mutex> lk(cmr::g_prof_mu);
  double ms = 0.0, work = 0.0;
  long long n = 0;
  std::vector<cmr::ProfRec> keep;
  for (cmr::ProfRec r : cmr::g_recs) {
    if (r.kind != kind && r.kind2 != kind) {
      keep.push_back(r);
      continue;
    }
    CMR_CUDA_TRY(cudaEventSynchronize(r.b));
    float t = 0.f;
    CMR_CUDA_TRY(cudaEventElapsedTime(&t, r.a, r.b));
    ms += t;
    ++n;
    if (r.kind2 == kind) {     // secondary view: the record stays for its primary kind
      work += r.work2;
      r.kind2 = -1;
      keep.push_back(r);
      continue;
    }
    work += r.work;
    cmr::g_free.push_back(r.a);
    cmr::g_free.push_back(r.b);
  }
  cmr::g_recs.swap(keep);
  *total_ms = ms;
  *total_work = work;
  *launches = n;
  return CMR_OK;
}

extern "C" const char* cmr_last_cuda_error(void) { return cmr::g_last_error; }
