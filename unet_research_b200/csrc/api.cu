// libb2u.so: error plumbing and device queries of the C ABI (include/b2u.h).
#include "b2u_common.cuh"

#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

static thread_local char g_err[1024] = "";

void b2u_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int b2u_num_sms() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int v = 0;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
    cached[dev] = v;
  }
  return cached[dev];
}

int b2u_pdl_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("B2U_PDL");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v;
}

extern "C" const char* b2u_last_error(void) { return g_err; }
extern "C" int b2u_version(void) { return B2U_VERSION; }

extern "C" int b2u_device_info(int* sm_count, int* max_threads_per_sm) {
  int dev = 0;
  B2U_CHECK_CUDA(cudaGetDevice(&dev));
  int sms = 0, mt = 0;
  B2U_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  B2U_CHECK_CUDA(cudaDeviceGetAttribute(&mt, cudaDevAttrMaxThreadsPerMultiProcessor, dev));
  if (sm_count) *sm_count = sms;
  if (max_threads_per_sm) *max_threads_per_sm = mt;
  return B2U_OK;
}
