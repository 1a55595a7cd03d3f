// Tall-skinny building blocks of the subspace-iteration modal solver (reference solver/solver.py:1084-1312): the reference
// keeps the subspace X [n_dof, k] as a torch tensor and walks its columns with torch ops (norm, dot against all previous
// columns, X^T K X, X Z); here the k vectors are stored one after the other (column j at X + j*ld, contiguous, so each is an
// SpMV operand as it stands) and three kernels cover everything that touches n_dof-long data:
//   mv_gram   G[i][j] = sum_r X_i[r] w[r] Y_j[r]      (norms, projections, A_k = Y^T K Y, B_k = Y^T M Y) -- per-CTA partials
//             summed in index order by a second kernel, so the k x k results are deterministic
//   mv_update Y_j = beta Y_j + sum_i X_i C[i][j]       (normalise, subtract projections, rotate the basis X Z)
//   mv_scale_mask  X_j[r] *= scale[r], zeroed where mask[r] == 0   (M^-1 and the clamped boundary dofs)
// The k x k problems themselves (Gauss-Jordan inverse, Jacobi sweeps) are a few hundred flops and stay on the host, as the
// reference's `.item()`-driven Python loops effectively do.
#include "common.cuh"

namespace femb {

constexpr int MV_MAXK = 8;  // ki, kj <= 8: 64 accumulators per thread at most
constexpr int MV_THREADS = 256;

struct MvCoef {
  double c[MV_MAXK * MV_MAXK];
};

template <int KI, int KJ>
__global__ void __launch_bounds__(MV_THREADS) mv_gram_kernel(long long n, const double* __restrict__ X, long long ldx, const double* __restrict__ Y,
                                                             long long ldy, const double* __restrict__ w, double* __restrict__ partial) {
  double acc[KI * KJ];
#pragma unroll
  for (int t = 0; t < KI * KJ; ++t) acc[t] = 0.0;
  for (long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x; r < n; r += (long long)gridDim.x * blockDim.x) {
    double y[KJ];
    const double wr = w ? w[r] : 1.0;
#pragma unroll
    for (int j = 0; j < KJ; ++j) y[j] = Y[j * ldy + r] * wr;
#pragma unroll
    for (int i = 0; i < KI; ++i) {
      const double x = X[i * ldx + r];
#pragma unroll
      for (int j = 0; j < KJ; ++j) acc[i * KJ + j] += x * y[j];
    }
  }
#pragma unroll
  for (int t = 0; t < KI * KJ; ++t) {
    const double s = block_sum<MV_THREADS>(acc[t]);
    if (threadIdx.x == 0) partial[(size_t)blockIdx.x * (KI * KJ) + t] = s;
  }
}

__global__ void __launch_bounds__(MV_THREADS) mv_gram_finish(int nparts, int kk, const double* __restrict__ partial, double* __restrict__ G) {
  for (int t = 0; t < kk; ++t) {
    double a = 0.0;
    for (int p = threadIdx.x; p < nparts; p += MV_THREADS) a += partial[(size_t)p * kk + t];
    a = block_sum<MV_THREADS>(a);
    if (threadIdx.x == 0) G[t] = a;
  }
}

template <int KI>
__global__ void __launch_bounds__(MV_THREADS) mv_update_kernel(long long n, const double* __restrict__ X, long long ldx, int kj, MvCoef C, double beta,
                                                               double* __restrict__ Y, long long ldy) {
  for (long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x; r < n; r += (long long)gridDim.x * blockDim.x) {
    double x[KI];
#pragma unroll
    for (int i = 0; i < KI; ++i) x[i] = X[i * ldx + r];  // read first: Y may alias columns of X (in-place scale / rotation)
    double out[MV_MAXK];
    for (int j = 0; j < kj; ++j) {
      double s = beta != 0.0 ? beta * Y[j * ldy + r] : 0.0;
#pragma unroll
      for (int i = 0; i < KI; ++i) s += x[i] * C.c[i * kj + j];
      out[j] = s;
    }
    for (int j = 0; j < kj; ++j) Y[j * ldy + r] = out[j];
  }
}

__global__ void mv_scale_mask_kernel(long long n, int k, double* __restrict__ X, long long ldx, const double* __restrict__ scale,
                                     const unsigned char* __restrict__ mask) {
  for (long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x; r < n; r += (long long)gridDim.x * blockDim.x) {
    const double f = (mask && !mask[r]) ? 0.0 : (scale ? scale[r] : 1.0);
    const bool zero = mask && !mask[r];
    for (int j = 0; j < k; ++j) X[j * ldx + r] = zero ? 0.0 : X[j * ldx + r] * f;
  }
}

template <int KI>
static void launch_gram_kj(int kj, int grid, cudaStream_t s, long long n, const double* X, long long ldx, const double* Y, long long ldy,
                           const double* w, double* partial) {
  switch (kj) {
    case 1: mv_gram_kernel<KI, 1><<<grid, MV_THREADS, 0, s>>>(n, X, ldx, Y, ldy, w, partial); break;
    case 2: mv_gram_kernel<KI, 2><<<grid, MV_THREADS, 0, s>>>(n, X, ldx, Y, ldy, w, partial); break;
    case 3: mv_gram_kernel<KI, 3><<<grid, MV_THREADS, 0, s>>>(n, X, ldx, Y, ldy, w, partial); break;
    case 4: mv_gram_kernel<KI, 4><<<grid, MV_THREADS, 0, s>>>(n, X, ldx, Y, ldy, w, partial); break;
    case 5: mv_gram_kernel<KI, 5><<<grid, MV_THREADS, 0, s>>>(n, X, ldx, Y, ldy, w, partial); break;
    case 6: mv_gram_kernel<KI, 6><<<grid, MV_THREADS, 0, s>>>(n, X, ldx, Y, ldy, w, partial); break;
    case 7: mv_gram_kernel<KI, 7><<<grid, MV_THREADS, 0, s>>>(n, X, ldx, Y, ldy, w, partial); break;
    default: mv_gram_kernel<KI, 8><<<grid, MV_THREADS, 0, s>>>(n, X, ldx, Y, ldy, w, partial); break;
  }
}

}  // namespace femb

using namespace femb;

extern "C" int femb_mv_gram(int64_t n, int ki, const double* X, int64_t ldx, int kj, const double* Y, int64_t ldy, const double* w, double* G_host,
                            femb_stream stream) {
  FEMB_CHECK_ARG(n >= 0 && ki >= 1 && ki <= MV_MAXK && kj >= 1 && kj <= MV_MAXK && X && Y && G_host, "1 <= ki, kj <= 8, non-null pointers");
  FEMB_CHECK_ARG(ldx >= n && ldy >= n, "leading dimensions must be >= n");
  cudaStream_t s = as_stream(stream);
  const int kk = ki * kj;
  if (n == 0) {
    for (int t = 0; t < kk; ++t) G_host[t] = 0.0;
    return FEMB_OK;
  }
  const int grid = grid_for(n, MV_THREADS, 4);
  Scratch scr(s);
  double *partial, *G;
  FEMB_CUDA(scr.alloc(&partial, (size_t)grid * kk));
  FEMB_CUDA(scr.alloc(&G, (size_t)kk));
  switch (ki) {
    case 1: launch_gram_kj<1>(kj, grid, s, n, X, ldx, Y, ldy, w, partial); break;
    case 2: launch_gram_kj<2>(kj, grid, s, n, X, ldx, Y, ldy, w, partial); break;
    case 3: launch_gram_kj<3>(kj, grid, s, n, X, ldx, Y, ldy, w, partial); break;
    case 4: launch_gram_kj<4>(kj, grid, s, n, X, ldx, Y, ldy, w, partial); break;
    case 5: launch_gram_kj<5>(kj, grid, s, n, X, ldx, Y, ldy, w, partial); break;
    case 6: launch_gram_kj<6>(kj, grid, s, n, X, ldx, Y, ldy, w, partial); break;
    case 7: launch_gram_kj<7>(kj, grid, s, n, X, ldx, Y, ldy, w, partial); break;
    default: launch_gram_kj<8>(kj, grid, s, n, X, ldx, Y, ldy, w, partial); break;
  }
  mv_gram_finish<<<1, MV_THREADS, 0, s>>>(grid, kk, partial, G);
  FEMB_LAUNCH_CHECK();
  FEMB_CUDA(cudaMemcpyAsync(G_host, G, sizeof(double) * kk, cudaMemcpyDeviceToHost, s));
  FEMB_CUDA(cudaStreamSynchronize(s));
  return FEMB_OK;
}

extern "C" int femb_mv_update(int64_t n, int ki, const double* X, int64_t ldx, int kj, const double* C_host, double beta, double* Y, int64_t ldy,
                              femb_stream stream) {
  FEMB_CHECK_ARG(n >= 0 && ki >= 1 && ki <= MV_MAXK && kj >= 1 && kj <= MV_MAXK && X && Y && C_host, "1 <= ki, kj <= 8, non-null pointers");
  FEMB_CHECK_ARG(ldx >= n && ldy >= n, "leading dimensions must be >= n");
  if (n == 0) return FEMB_OK;
  cudaStream_t s = as_stream(stream);
  MvCoef C;
  memset(&C, 0, sizeof(C));
  for (int t = 0; t < ki * kj; ++t) C.c[t] = C_host[t];
  const int grid = grid_for(n, MV_THREADS, 8);
#define UPD(K) mv_update_kernel<K><<<grid, MV_THREADS, 0, s>>>(n, X, ldx, kj, C, beta, Y, ldy)
  switch (ki) {
    case 1: UPD(1); break;
    case 2: UPD(2); break;
    case 3: UPD(3); break;
    case 4: UPD(4); break;
    case 5: UPD(5); break;
    case 6: UPD(6); break;
    case 7: UPD(7); break;
    default: UPD(8); break;
  }
#undef UPD
  FEMB_LAUNCH_CHECK();
  return FEMB_OK;
}

extern "C" int femb_mv_scale_mask(int64_t n, int k, double* X, int64_t ldx, const double* scale, const uint8_t* mask, femb_stream stream) {
  FEMB_CHECK_ARG(n >= 0 && k >= 1 && X && ldx >= n, "n >= 0, k >= 1, ldx >= n");
  if (n == 0) return FEMB_OK;
  mv_scale_mask_kernel<<<grid_for(n, 256, 8), 256, 0, as_stream(stream)>>>(n, k, X, ldx, scale, mask);
  FEMB_LAUNCH_CHECK();
  return FEMB_OK;
}
