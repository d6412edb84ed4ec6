// api.cu -- ABI version, thread-local error string, device check.
#include "common.cuh"
#include <string.h>

namespace tsg {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
}  // namespace tsg

extern "C" int tsg_abi_version(void) { return TSG_ABI_VERSION; }
extern "C" const char* tsg_last_error(void) { return tsg::g_err; }
extern "C" int tsg_check_device(void) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) { tsg::set_error("no CUDA device: %s", cudaGetErrorString(e)); return TSG_EARCH; }
  int major = 0;
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  if (major != 10) { tsg::set_error("libtsg is built for sm_100a only; device has compute capability %d.x", major); return TSG_EARCH; }
  return TSG_OK;
}
