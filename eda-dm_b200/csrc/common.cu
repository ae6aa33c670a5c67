// Error plumbing shared by every C-ABI entry point.
#include "common.cuh"
#include <stdarg.h>

namespace edadm {

char* last_error_buf() {
  static thread_local char buf[512] = {0};
  return buf;
}

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(last_error_buf(), 512, fmt, ap);
  va_end(ap);
  return code;
}

int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(EDADM_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
  return EDADM_OK;
}

int sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) {
      cudaGetLastError();
      n = 148;  // B200
    }
  }
  return n;
}

}  // namespace edadm

extern "C" const char* edadm_last_error(void) { return edadm::last_error_buf(); }
extern "C" int edadm_abi_version(void) { return 4; }
