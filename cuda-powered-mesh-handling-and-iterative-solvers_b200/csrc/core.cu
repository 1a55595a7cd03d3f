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

SolveCtx* solve_ctx(cudaStream_t user) {
  constexpr int MAX_DEV = 64;
  static thread_local SolveCtx table[MAX_DEV];
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess || dev < 0 || dev >= MAX_DEV) {
    set_error(std::string("solve_ctx: cudaGetDevice: ") + (e != cudaSuccess ? cudaGetErrorString(e) : "device index out of range"));
    return nullptr;
  }
  SolveCtx* c = &table[dev];
  if (!c->stream) {
    cudaStream_t s = nullptr;
    if ((e = cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking)) == cudaSuccess) e = cudaEventCreateWithFlags(&c->order_ev, cudaEventDisableTiming);
    for (int k = 0; k < 2 && e == cudaSuccess; ++k) {
      e = cudaEventCreateWithFlags(&c->poll_ev[k], cudaEventDisableTiming);
      if (e == cudaSuccess) e = cudaEventCreate(&c->time_ev[k]);
    }
    if (e == cudaSuccess) e = cudaMallocHost(&c->pinned, SOLVE_PINNED_BYTES);
    if (e != cudaSuccess) {
      set_error(std::string("solve_ctx: creating the solver stream / events: ") + cudaGetErrorString(e));
      return nullptr;
    }
    c->stream = s;
  }
  if ((e = cudaEventRecord(c->order_ev, user)) != cudaSuccess || (e = cudaStreamWaitEvent(c->stream, c->order_ev, 0)) != cudaSuccess) {
    set_error(std::string("solve_ctx: ordering the solver stream after the caller's stream: ") + cudaGetErrorString(e));
    return nullptr;
  }
  return c;
}
}  // namespace femb

extern "C" const char* femb_last_error(void) { return femb::g_err.c_str(); }
extern "C" int femb_version(void) { return 100; }
