// libsininn: error reporting, device info.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>
#include "common.cuh"

namespace sininn {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

bool pdl_enabled() {
  static int mode = -1;
  if (mode < 0) {
    const char* e = getenv("SININN_PDL");
    mode = (e && atoi(e) == 0) ? 0 : 1;
  }
  return mode == 1;
}

int sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

}  // namespace sininn

extern "C" {

int sininn_version(void) { return 100; }

const char* sininn_last_error(void) { return sininn::g_err; }

int sininn_device_info(int* sm_count, int* cc_major, int* cc_minor) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) {
    sininn::set_error("cudaGetDevice: %s", cudaGetErrorString(e));
    return SININN_ECUDA;
  }
  cudaDeviceProp p;
  e = cudaGetDeviceProperties(&p, dev);
  if (e != cudaSuccess) {
    sininn::set_error("cudaGetDeviceProperties: %s", cudaGetErrorString(e));
    return SININN_ECUDA;
  }
  if (sm_count) *sm_count = p.multiProcessorCount;
  if (cc_major) *cc_major = p.major;
  if (cc_minor) *cc_minor = p.minor;
  return SININN_OK;
}

}  // extern "C"
