// Library-wide pieces of the C ABI: version, error string, device query.
#include <cstring>

#include "sod_common.cuh"

namespace sod {

namespace {
thread_local char g_err[512] = {0};
}

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int device_sm_count() {
  int dev = 0, sms = 0;
  if (cudaGetDevice(&dev) != cudaSuccess ||
      cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) {
    set_error("cannot query the current CUDA device: %s", cudaGetErrorString(cudaGetLastError()));
    return -1;
  }
  return sms;
}

}  // namespace sod

extern "C" {

int sod_version(void) { return 1000; }

const char* sod_last_error(void) { return sod::g_err; }

int sod_device_sm_count(void) { return sod::device_sm_count(); }

}  // extern "C"
