// Library plumbing: error string, device info.
#include "common.cuh"
#include <stdarg.h>
#include <stdlib.h>

namespace pmc {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
bool pdl_enabled() {
  static const bool on = [] { const char* e = getenv("PMC_B200_PDL"); return !(e && e[0] == '0'); }();
  return on;
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

// ---- host <-> device staging of the per-step arrays (the likelihood is a host black box: x' leaves, logl' returns) ----
extern "C" int pmc_event_create(void** event_out) {
  PMC_REQUIRE(event_out, "pmc_event_create: null pointer");
  cudaEvent_t e;
  PMC_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  *event_out = e;
  return 0;
}
extern "C" int pmc_event_destroy(void* event) {
  if (event) PMC_TRY(cudaEventDestroy(static_cast<cudaEvent_t>(event)));
  return 0;
}
extern "C" int pmc_event_synchronize(void* event) {
  PMC_REQUIRE(event, "pmc_event_synchronize: null event");
  PMC_TRY(cudaEventSynchronize(static_cast<cudaEvent_t>(event)));
  return 0;
}
extern "C" int pmc_stream_synchronize(pmc_stream_t stream) {
  PMC_TRY(cudaStreamSynchronize(pmc::as_stream(stream)));
  return 0;
}
extern "C" int pmc_memcpy_async(void* dst, const void* src, int64_t bytes, pmc_stream_t stream) {
  if (bytes == 0) return 0;
  PMC_REQUIRE(dst && src && bytes > 0, "pmc_memcpy_async: bad arguments");
  PMC_TRY(cudaMemcpyAsync(dst, src, (size_t)bytes, cudaMemcpyDefault, pmc::as_stream(stream)));
  return 0;
}
extern "C" int pmc_download_rows(const double* x_dev, double* x_host, const uint8_t* flag_dev, uint8_t* flag_host, int64_t n,
                                 int32_t d, int32_t n_chunks, void* const* events, pmc_stream_t stream) {
  PMC_REQUIRE(x_dev && x_host && n >= 0 && d >= 1 && n_chunks >= 1 && events, "pmc_download_rows: bad arguments");
  PMC_REQUIRE((flag_dev == nullptr) == (flag_host == nullptr), "pmc_download_rows: flags need both pointers");
  cudaStream_t s = pmc::as_stream(stream);
  if (flag_dev && n > 0) PMC_TRY(cudaMemcpyAsync(flag_host, flag_dev, (size_t)n, cudaMemcpyDeviceToHost, s));
  const int64_t per = (n + n_chunks - 1) / n_chunks;
  for (int k = 0; k < n_chunks; ++k) {
    const int64_t a = (int64_t)k * per, b = a + per < n ? a + per : n;
    if (b > a) PMC_TRY(cudaMemcpyAsync(x_host + a * d, x_dev + a * d, (size_t)(b - a) * d * sizeof(double), cudaMemcpyDeviceToHost, s));
    PMC_REQUIRE(events[k], "pmc_download_rows: null event");
    PMC_TRY(cudaEventRecord(static_cast<cudaEvent_t>(events[k]), s));
  }
  return 0;
}
