// Host-side runtime of libcft_b200: error text, device check, launch counters, profiling.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <mutex>

#include "cft_common.cuh"

namespace cft {

static thread_local char g_err[512] = "";
static long long g_launches = 0;
static bool g_prof = false;
static double g_ms[CFT_K_COUNT];
static long long g_cnt[CFT_K_COUNT];
static std::mutex g_mu;

struct PendingEvent {
  int id;
  cudaEvent_t e0, e1;
};
static PendingEvent* g_pending = nullptr;
static int g_npending = 0, g_cap = 0;

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int fail_arg(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return CFT_E_ARG;
}

int check_cuda(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return CFT_OK;
  set_error("%s: %s", what, cudaGetErrorString(e));
  return CFT_E_CUDA;
}

int sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
  }
  return n;
}

static void drain_pending() {
  for (int i = 0; i < g_npending; ++i) {
    float ms = 0.f;
    cudaEventSynchronize(g_pending[i].e1);
    if (cudaEventElapsedTime(&ms, g_pending[i].e0, g_pending[i].e1) == cudaSuccess) {
      g_ms[g_pending[i].id] += ms;
      g_cnt[g_pending[i].id] += 1;
    }
    cudaEventDestroy(g_pending[i].e0);
    cudaEventDestroy(g_pending[i].e1);
  }
  g_npending = 0;
}

LaunchScope::LaunchScope(int kernel_id, cudaStream_t s) : id(kernel_id), stream(s) {
  if (g_prof) {
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0, stream);
  }
}

int LaunchScope::finish(const char* what) {
  cudaError_t e = cudaGetLastError();
  std::lock_guard<std::mutex> lk(g_mu);
  g_launches += 1;
  if (e0) {
    cudaEventRecord(e1, stream);
    if (g_npending == g_cap) {
      int ncap = g_cap ? g_cap * 2 : 1024;
      PendingEvent* np = (PendingEvent*)realloc(g_pending, sizeof(PendingEvent) * ncap);
      if (np) {
        g_pending = np;
        g_cap = ncap;
      }
    }
    if (g_npending < g_cap) g_pending[g_npending++] = PendingEvent{id, e0, e1};
  }
  return check_cuda(e, what);
}

}  // namespace cft

using namespace cft;

extern "C" int cft_abi_version(void) { return CFT_ABI_VERSION; }
extern "C" const char* cft_last_error(void) { return g_err; }

extern "C" int cft_check_device(int* sms, int* major, int* minor) {
  int dev = 0;
  int rc = check_cuda(cudaGetDevice(&dev), "cudaGetDevice");
  if (rc) return rc;
  int ma = 0, mi = 0, n = 0;
  cudaDeviceGetAttribute(&ma, cudaDevAttrComputeCapabilityMajor, dev);
  cudaDeviceGetAttribute(&mi, cudaDevAttrComputeCapabilityMinor, dev);
  cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  if (sms) *sms = n;
  if (major) *major = ma;
  if (minor) *minor = mi;
  if (ma != 10) {
    set_error("libcft_b200 is built for sm_100a only; device is sm_%d%d", ma, mi);
    return CFT_E_UNSUPPORTED;
  }
  return CFT_OK;
}

extern "C" int cft_prof_enable(int on) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (on) {
    drain_pending();
    memset(g_ms, 0, sizeof(g_ms));
    memset(g_cnt, 0, sizeof(g_cnt));
  }
  g_prof = on != 0;
  return CFT_OK;
}

extern "C" int cft_prof_get(int id, double* total_ms, long long* launches) {
  if (id < 0 || id >= CFT_K_COUNT) return fail_arg("cft_prof_get: bad kernel id %d", id);
  std::lock_guard<std::mutex> lk(g_mu);
  drain_pending();
  if (total_ms) *total_ms = g_ms[id];
  if (launches) *launches = g_cnt[id];
  return CFT_OK;
}

extern "C" long long cft_launch_count(void) { return g_launches; }
