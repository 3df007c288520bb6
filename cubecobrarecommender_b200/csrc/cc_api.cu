// Error convention and library-level entry points of libcubecobra_b200.so.
#include <stdarg.h>

#include "cc_common.cuh"

namespace cc {

static thread_local char g_error[1024] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what, const char* file, int line) {
  set_error("CUDA error %d (%s) at %s:%d: %s", int(e), cudaGetErrorName(e), file, line, what);
  // keep the sticky/last error readable by the caller's runtime too, but clear ours
  (void)cudaGetLastError();
  return CC_ERR_CUDA;
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

}  // namespace cc

extern "C" {

const char* cc_last_error(void) { return cc::g_error; }

int cc_version(void) { return CC_VERSION; }

// 0 when a CUDA device of compute capability 10.x is current, an error otherwise
// (the library carries sm_100a SASS only: there is no other code path).
int cc_device_check(void) {
  int dev = 0;
  CC_CHECK_CUDA(cudaGetDevice(&dev));
  int major = 0, minor = 0;
  CC_CHECK_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  CC_CHECK_CUDA(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev));
  if (major != 10) {
    cc::set_error("cubecobra_b200 needs an sm_100a device (B200); device %d is sm_%d%d", dev, major, minor);
    return CC_ERR_DEVICE;
  }
  return CC_OK;
}

}  // extern "C"
