// Error string, version and device queries for libfemb200.
#include "common.cuh"

namespace femb {
static thread_local std::string g_err;
void set_error(const std::string& msg) { g_err = msg; }
int num_sms() {
  static int sms = 0;
  if (!sms) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) sms = SMS;
  }
  return sms;
}
}  // namespace femb

extern "C" const char* femb_last_error(void) { return femb::g_err.c_str(); }
extern "C" int femb_version(void) { return 100; }
