// Error string plumbing + version query for the C ABI.
#include "common.cuh"
#include <cstdarg>
#include <cstring>

static thread_local char g_err[512] = "";

extern "C" {
void vca_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
const char* vca_last_error() { return g_err; }
int vca_abi_version() { return 1; }
// 1 when the current device is an sm_100 part (the only one this library carries SASS for).
int vca_device_ok() {
  int dev = 0, major = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  return major == 10;
}
}
