// libgeobi: error channel and device queries.
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace geobi {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
}  // namespace geobi

extern "C" const char* geobi_last_error(void) { return geobi::g_err; }
extern "C" int geobi_version(void) { return 100; }

extern "C" int geobi_device_info(int* sm_count_host, int* cc_major_host, int* cc_minor_host) {
  int dev = 0;
  GEOBI_CUDA_OK(cudaGetDevice(&dev));
  cudaDeviceProp p;
  GEOBI_CUDA_OK(cudaGetDeviceProperties(&p, dev));
  if (sm_count_host) *sm_count_host = p.multiProcessorCount;
  if (cc_major_host) *cc_major_host = p.major;
  if (cc_minor_host) *cc_minor_host = p.minor;
  return GEOBI_OK;
}
