// abi.cu -- library-wide state of the C ABI (error text, version, launch accounting).
#include "common.cuh"

namespace osb {
thread_local std::string g_last_error;
std::atomic<long long> g_launches{0};
}  // namespace osb

extern "C" const char* osb_last_error(void) { return osb::g_last_error.c_str(); }
extern "C" const char* osb_version(void) { return "omniswarm_b200 0.1.0 (sm_100a)"; }
extern "C" int osb_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}
extern "C" int64_t osb_launch_count(void) { return (int64_t)osb::g_launches.load(); }
