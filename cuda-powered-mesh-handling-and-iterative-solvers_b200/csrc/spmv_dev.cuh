// Device-side building blocks shared by the single-GPU solver (krylov.cu) and the multi-GPU solver (dist.cu).
#pragma once
#include "common.cuh"

namespace femb {

constexpr int SPMV_THREADS = 256;
constexpr int STREAM_CAP = 5632;  // products per CTA (44 KB of static shared memory)

__device__ __forceinline__ double ld_stream(const double* p) {
  double v;
  asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ int ld_stream(const int* p) {
  int v;
  asm volatile("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}
// x gathers: read-only path when x is immutable during the kernel, plain (coherent at L2, L1-cached) loads when peers
// write the ghost part of x before this kernel's flag wait (multi-GPU)
template <bool NC>
__device__ __forceinline__ double ld_x(const double* p) {
  if (NC) return __ldg(p);
  return *p;
}

// CSR-stream SpMV over the rows [0,n): a CTA owns R = 256/LR consecutive rows.  Their nonzeros form one contiguous slice
// of val/col, which the whole CTA streams with fully coalesced loads, multiplies by the gathered x and parks in shared
// memory; then LR lanes per row add up that row's products (index order within a lane, fixed shuffle tree across lanes).
// Short FEM rows (~15 nonzeros for P1) therefore cost no idle lanes and no per-row pointer chasing while streaming.
// Returns this thread's partial of sum_r y_r * x_r (only when `fused`), with masked rows forced to zero.
struct NoHaloWait {
  __device__ __forceinline__ void operator()() const {}
};

// `halo_row`: rows >= halo_row may read entries of x that another GPU is still writing; `wait` (block-uniform, may
// __syncthreads) is called once, right before this CTA touches its first such tile, so interior tiles overlap the exchange.
template <int LR, bool NC, typename Wait = NoHaloWait>
__device__ __forceinline__ double spmv_stream_rows(long long n, const int* __restrict__ crow, const int* __restrict__ col,
                                                   const double* __restrict__ val, const double* __restrict__ x, double* __restrict__ y,
                                                   const unsigned char* __restrict__ mask, bool accumulate, bool fused,
                                                   long long halo_row = 0x7fffffffffffffffll, Wait wait = Wait()) {
  constexpr int R = SPMV_THREADS / LR;
  __shared__ double prod[STREAM_CAP];
  __shared__ int rp[R + 1];
  const int tid = threadIdx.x, sub = tid % LR, lr = tid / LR;
  double dot = 0.0;
  bool waited = false;
  for (long long r0 = (long long)blockIdx.x * R; r0 < n; r0 += (long long)gridDim.x * R) {
    const int nr = (int)min((long long)R, n - r0);
    if (!waited && r0 + nr > halo_row) {
      wait();
      waited = true;
    }
    for (int t = tid; t <= nr; t += SPMV_THREADS) rp[t] = __ldg(crow + r0 + t);
    __syncthreads();
    const int s = rp[0], e = rp[nr];
    const bool fits = (e - s) <= STREAM_CAP;
    if (fits) {
      int j = s + tid;
      for (; j + 3 * SPMV_THREADS < e; j += 4 * SPMV_THREADS) {
        int c[4];
        double v[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) c[q] = ld_stream(col + j + q * SPMV_THREADS), v[q] = ld_stream(val + j + q * SPMV_THREADS);
#pragma unroll
        for (int q = 0; q < 4; ++q) prod[j - s + q * SPMV_THREADS] = v[q] * ld_x<NC>(x + c[q]);
      }
      for (; j < e; j += SPMV_THREADS) prod[j - s] = ld_stream(val + j) * ld_x<NC>(x + ld_stream(col + j));
    }
    __syncthreads();
    double sum = 0.0;
    if (lr < nr) {
      const int a = rp[lr], b = rp[lr + 1];
      if (fits) {
        for (int j = a - s + sub; j < b - s; j += LR) sum += prod[j];
      } else {  // oversized slice (very long rows): read straight from global memory
        for (int j = a + sub; j < b; j += LR) sum += ld_stream(val + j) * ld_x<NC>(x + ld_stream(col + j));
      }
    }
#pragma unroll
    for (int o = LR / 2; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    if (lr < nr && sub == 0) {
      const long long r = r0 + lr;
      if (accumulate) sum += y[r];
      if (fused) {
        if (mask && !mask[r]) sum = 0.0;
        dot += sum * ld_x<NC>(x + r);
      }
      y[r] = sum;
    }
    __syncthreads();
  }
  return dot;
}

}  // namespace femb
