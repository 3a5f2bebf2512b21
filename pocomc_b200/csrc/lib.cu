// Library plumbing: error string, device info.
#include "common.cuh"
#include <stdarg.h>

namespace pmc {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
int sm_count() {
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
}  // namespace pmc

extern "C" const char* pmc_last_error(void) { return pmc::g_err; }
extern "C" int pmc_version(void) { return 100; }
extern "C" int pmc_device_info(int32_t* sms, int32_t* major, int32_t* minor) {
  int dev = 0;
  PMC_TRY(cudaGetDevice(&dev));
  int a = 0, b = 0, c = 0;
  PMC_TRY(cudaDeviceGetAttribute(&a, cudaDevAttrMultiProcessorCount, dev));
  PMC_TRY(cudaDeviceGetAttribute(&b, cudaDevAttrComputeCapabilityMajor, dev));
  PMC_TRY(cudaDeviceGetAttribute(&c, cudaDevAttrComputeCapabilityMinor, dev));
  if (sms) *sms = a;
  if (major) *major = b;
  if (minor) *minor = c;
  return 0;
}
