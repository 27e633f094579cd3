// errors.cu -- last-error bookkeeping shared by every C-ABI entry point.
#include "common.cuh"
#include <stdarg.h>

namespace trl {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int check_launch(const char* what) {
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: launch failed: %s", what, cudaGetErrorString(e));
    return static_cast<int>(e);
  }
  return TRL_OK;
}

}  // namespace trl

TRL_API const char* trl_last_error(void) { return trl::g_err; }

TRL_API int trl_abi_version(void) { return 3; }

// Number of SMs / name of the device the calling thread is bound to (diagnostics).
TRL_API int trl_device_info(int* sm_count, int* cc_major, int* cc_minor) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) { trl::set_error("cudaGetDevice: %s", cudaGetErrorString(e)); return (int)e; }
  cudaDeviceProp prop;
  e = cudaGetDeviceProperties(&prop, dev);
  if (e != cudaSuccess) { trl::set_error("cudaGetDeviceProperties: %s", cudaGetErrorString(e)); return (int)e; }
  if (sm_count) *sm_count = prop.multiProcessorCount;
  if (cc_major) *cc_major = prop.major;
  if (cc_minor) *cc_minor = prop.minor;
  return TRL_OK;
}
