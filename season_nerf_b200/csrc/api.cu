// Library-level entry points of the C ABI (version, launch counter, error strings).
#include <atomic>
#include "common.cuh"
#include "api.h"

namespace snb {
static std::atomic<long long> g_launches{0};
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int num_sms() {
  static std::atomic<int> cache[64];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  int n = cache[dev].load(std::memory_order_relaxed);
  if (n == 0) {
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cache[dev].store(n, std::memory_order_relaxed);
  }
  return n;
}
}  // namespace snb

extern "C" int snb_version(void) { return 200; }
extern "C" int snb_num_sms(void) { return snb::num_sms(); }
extern "C" long long snb_launch_count(void) { return snb::g_launches.load(); }
extern "C" const char* snb_error_string(int code) {
  if (code == SNB_OK) return "ok";
  if (code == SNB_ERR_ARG) return "season_nerf_b200: invalid argument";
  if (code == SNB_ERR_UNSUPPORTED) return "season_nerf_b200: unsupported configuration";
  if (code > 0) return cudaGetErrorString((cudaError_t)code);
  return "season_nerf_b200: unknown error";
}
