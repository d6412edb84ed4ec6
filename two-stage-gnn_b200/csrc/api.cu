// api.cu -- ABI version, thread-local error string, device check.
#include "common.cuh"
#include <string.h>
#include <mutex>

namespace tsg {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
}  // namespace tsg

namespace tsg {
// Ticket pool for the fused reduction tails (common.cuh partial_sum_tail): 4,096 zeroed counters per device, handed
// out round robin.  A counter returns to zero when its kernel ends, so a slot is clean again long before the 4,096
// calls that bring it round; kernels running concurrently on different streams get different slots.
static constexpr int TICKETS = 4096;
static unsigned* g_ticket_pool[64];
static unsigned g_ticket_cursor[64];
static std::mutex g_ticket_mu;

// Created by tsg_init_device() only (the ONE entry point that allocates and synchronises; call it once per device,
// outside any stream capture).  ticket_next() never allocates: without a pool it returns nullptr and the callers
// take their separate-launch second stage (or report TSG_EINVAL where no such variant exists).
int ticket_pool_init() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return -1;
  std::lock_guard<std::mutex> lock(g_ticket_mu);
  if (g_ticket_pool[dev]) return 0;
  unsigned* p = nullptr;
  if (cudaMalloc(&p, TICKETS * sizeof(unsigned)) != cudaSuccess) { cudaGetLastError(); return -1; }
  if (cudaMemset(p, 0, TICKETS * sizeof(unsigned)) != cudaSuccess) { cudaGetLastError(); cudaFree(p); return -1; }
  g_ticket_pool[dev] = p;
  return 0;
}

unsigned* ticket_next() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  std::lock_guard<std::mutex> lock(g_ticket_mu);
  if (!g_ticket_pool[dev]) return nullptr;
  return g_ticket_pool[dev] + (g_ticket_cursor[dev]++ % TICKETS);
}
}  // namespace tsg

extern "C" int tsg_abi_version(void) { return TSG_ABI_VERSION; }
extern "C" const char* tsg_last_error(void) { return tsg::g_err; }
extern "C" int tsg_init_device(void) {
  if (tsg::ticket_pool_init() != 0) { tsg::set_error("tsg_init_device: cannot create the per-device counter pool"); return TSG_ELAUNCH; }
  return TSG_OK;
}
extern "C" int tsg_check_device(void) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) { tsg::set_error("no CUDA device: %s", cudaGetErrorString(e)); return TSG_EARCH; }
  int major = 0;
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  if (major != 10) { tsg::set_error("libtsg is built for sm_100a only; device has compute capability %d.x", major); return TSG_EARCH; }
  return TSG_OK;
}
