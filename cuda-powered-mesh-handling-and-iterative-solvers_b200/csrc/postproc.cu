// Post-processing kernels around the element path: shell displacements in element frames and shell stress post-processing
// (reference solver/shell.py:41-56, :104-160), tet face forces and their balance over shared faces (element.py:3343-3384),
// unit face normals of wedges (element.py:2377-2420).  One thread per output row, coalesced enough for their size: all of
// them are a single streaming pass over arrays the element kernels just produced.
#include "common.cuh"

namespace femb {

// local[m,i,0:3] = unit[m] g[0:3], local[m,i,3:6] = unit[m] g[3:6] with g = disp[conn[m,i]] (einsum 'mij,mkj->mik')
template <typename T, typename I>
__global__ void shell_local_disp_kernel(const I* __restrict__ conn, long long M, int nen, const T* __restrict__ disp, const T* __restrict__ unit,
                                        T* __restrict__ out) {
  const long long total = M * nen;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const long long m = t / nen;
    const long long n = ldidx(conn + t);
    T g[6], R[9];
    for (int k = 0; k < 6; ++k) g[k] = __ldg(disp + 6 * n + k);
    for (int k = 0; k < 9; ++k) R[k] = __ldg(unit + 9 * m + k);
    for (int h = 0; h < 2; ++h)
      for (int k = 0; k < 3; ++k) out[6 * t + 3 * h + k] = g[3 * h] * R[3 * k] + g[3 * h + 1] * R[3 * k + 1] + g[3 * h + 2] * R[3 * k + 2];
  }
}

// local[m,i,d] = (x[conn[m,i]] - x[conn[m,0]]) . unit[m,d,:]   (einsum 'mna,mda->mnd', shell.py:339-345, :641-647)
template <typename T, typename I>
__global__ void shell_local_coords_kernel(const T* __restrict__ coords, const I* __restrict__ conn, long long M, int nen, const T* __restrict__ unit,
                                          T* __restrict__ out) {
  const long long total = M * nen;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const long long m = t / nen;
    const long long n = ldidx(conn + t), n0 = ldidx(conn + m * nen);
    T v[3];
    for (int c = 0; c < 3; ++c) v[c] = __ldg(coords + 3 * n + c) - __ldg(coords + 3 * n0 + c);
    for (int d = 0; d < 3; ++d) out[3 * t + d] = v[0] * __ldg(unit + 9 * m + 3 * d) + v[1] * __ldg(unit + 9 * m + 3 * d + 1) + v[2] * __ldg(unit + 9 * m + 3 * d + 2);
  }
}

// out[8,M]: sx, sy, txy, s1, s2, theta_p, tau_max, vm (shell.py:122-160); NMQ row stride = ncol (6 or 8)
template <typename T>
__global__ void shell_postprocess_kernel(const T* __restrict__ nmq, long long M, int ncol, T f1, T f2, T* __restrict__ out) {
  for (long long m = blockIdx.x * (long long)blockDim.x + threadIdx.x; m < M; m += (long long)gridDim.x * blockDim.x) {
    const T* r = nmq + m * ncol;
    const T sx = r[0] * f1 + r[3] * f2, sy = r[1] * f1 + r[4] * f2, txy = r[2] * f1 + r[5] * f2;
    const T half = T(0.5) * (sx + sy), diff = T(0.5) * (sx - sy);
    const T R = sqrt(diff * diff + txy * txy);
    const T s1 = half + R, s2 = half - R;
    T den = sx - sy;
    if (den < T(1.0e-30)) den = T(1.0e-30);  // clamp(min=eps): only the positive half-plane is kept, as in the reference
    const T th = T(0.5) * atan2(T(2.0) * txy, den);
    out[m] = sx, out[M + m] = sy, out[2 * M + m] = txy, out[3 * M + m] = s1, out[4 * M + m] = s2, out[5 * M + m] = th;
    out[6 * M + m] = T(0.5) * (s1 - s2);
    out[7 * M + m] = sqrt(s1 * s1 - s1 * s2 + s2 * s2 + T(1.0e-30));
  }
}

// forces[m,f,:] = stress[m] normal[m,f,:]  (matmul of [M,1,3,3] with [M,nf,3,1])
template <typename T>
__global__ void face_forces_kernel(const T* __restrict__ normals, const T* __restrict__ stress, long long M, int nf, T* __restrict__ out) {
  const long long total = M * nf;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const long long m = t / nf;
    const T n0 = normals[3 * t], n1 = normals[3 * t + 1], n2 = normals[3 * t + 2];
    const T* S = stress + 9 * m;
    for (int i = 0; i < 3; ++i) out[3 * t + i] = S[3 * i] * n0 + S[3 * i + 1] * n1 + S[3 * i + 2] * n2;
  }
}

// out[s,:] = forces[e1,f1,:] + forces[e2,f2,:] for pairs[s] = ((e1,f1),(e2,f2))
template <typename T>
__global__ void shared_forces_kernel(const long long* __restrict__ pairs, long long S, const T* __restrict__ forces, int nf, T* __restrict__ out) {
  for (long long s = blockIdx.x * (long long)blockDim.x + threadIdx.x; s < S; s += (long long)gridDim.x * blockDim.x) {
    const long long a = (pairs[4 * s] * nf + pairs[4 * s + 1]) * 3, b = (pairs[4 * s + 2] * nf + pairs[4 * s + 3]) * 3;
    for (int c = 0; c < 3; ++c) out[3 * s + c] = forces[a + c] + forces[b + c];
  }
}

// wedge faces: quads (0,1,4,3) (1,2,5,4) (2,0,3,5) with edges 0->1, 0->3; triangles (0,2,1) (3,4,5) with edges 0->1, 0->2;
// UNIT normals, no orientation fix (element.py:2392-2420)
template <typename T, typename I>
__global__ void wedge_normals_kernel(const T* __restrict__ coords, const I* __restrict__ conn, long long M, int stride, T* __restrict__ out) {
  const int tab[5][3] = {{0, 1, 3}, {1, 2, 4}, {2, 0, 5}, {0, 2, 1}, {3, 4, 5}};  // origin, end of edge 1, end of edge 2
  const long long total = M * 5;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const long long e = t / 5;
    const int f = (int)(t - e * 5);
    const long long n0 = ldidx(conn + e * stride + tab[f][0]), n1 = ldidx(conn + e * stride + tab[f][1]), n2 = ldidx(conn + e * stride + tab[f][2]);
    T a[3], b[3];
    for (int c = 0; c < 3; ++c) {
      const T x0 = coords[3 * n0 + c];
      a[c] = coords[3 * n1 + c] - x0;
      b[c] = coords[3 * n2 + c] - x0;
    }
    const T n[3] = {a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]};
    const T len = sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
    for (int c = 0; c < 3; ++c) out[3 * t + c] = n[c] / len;
  }
}

// compute_s3_normal shell.py:184-203 (cross(x1-x0, x2-x0)/2) and compute_s4_normal :483-502 (cross(x1-x0, x3-x0))
template <typename T, typename I>
__global__ void shell_normal_kernel(const T* __restrict__ coords, const I* __restrict__ conn, long long M, int nen, T* __restrict__ out) {
  const int second = nen == 3 ? 2 : 3;
  const T sc = nen == 3 ? T(0.5) : T(1);
  for (long long m = blockIdx.x * (long long)blockDim.x + threadIdx.x; m < M; m += (long long)gridDim.x * blockDim.x) {
    const long long n0 = ldidx(conn + m * nen), n1 = ldidx(conn + m * nen + 1), n2 = ldidx(conn + m * nen + second);
    T a[3], b[3];
    for (int c = 0; c < 3; ++c) {
      const T x0 = coords[3 * n0 + c];
      a[c] = coords[3 * n1 + c] - x0;
      b[c] = coords[3 * n2 + c] - x0;
    }
    out[3 * m] = (a[1] * b[2] - a[2] * b[1]) * sc;
    out[3 * m + 1] = (a[2] * b[0] - a[0] * b[2]) * sc;
    out[3 * m + 2] = (a[0] * b[1] - a[1] * b[0]) * sc;
  }
}

// Element operator of a shell in global axes: Kg = T^T K T with T = blockdiag(R, R, ...) (R = unit, rows = local axes), i.e.
// every 3x3 block (p,q) of K becomes R^T K_pq R -- what compute_shell_nodal_forces (shell.py:58-102) applies implicitly by
// rotating the nodal vectors into the element frame and the forces back.  One thread per block.
template <typename T>
__global__ void shell_rotate_K_kernel(const T* __restrict__ K, const T* __restrict__ unit, long long M, int nb, T* __restrict__ out) {
  const long long total = M * nb * nb;
  const int nd = 3 * nb;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const long long m = t / (nb * nb);
    const int pq = (int)(t - m * nb * nb), p = pq / nb, q = pq - p * nb;
    T R[9], B[9], W[9];
    for (int k = 0; k < 9; ++k) R[k] = __ldg(unit + 9 * m + k);
    const T* src = K + (m * nd + 3 * p) * nd + 3 * q;
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) B[3 * i + j] = src[i * nd + j];
    for (int i = 0; i < 3; ++i)  // W = B R
      for (int b = 0; b < 3; ++b) W[3 * i + b] = B[3 * i] * R[b] + B[3 * i + 1] * R[3 + b] + B[3 * i + 2] * R[6 + b];
    T* dst = out + (m * nd + 3 * p) * nd + 3 * q;
    for (int a = 0; a < 3; ++a)  // out = R^T W
      for (int b = 0; b < 3; ++b) dst[a * nd + b] = R[a] * W[b] + R[3 + a] * W[3 + b] + R[6 + a] * W[6 + b];
  }
}

// subdivision.ipynb cell 15: every subdomain starts from the global load vector ...
template <typename T>
__global__ void subdomain_broadcast_kernel(const T* __restrict__ F, long long len, int n_sub, T* __restrict__ out) {
  const long long total = len * n_sub;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) out[t] = F[t % len];
}
// ... and every (subdomain, interface node) target receives its interface-force unknowns: + the one it shares with the next
// subdomain of the node's group, - the one it shares with the previous (targets are unique, so no two threads write one entry)
template <typename T>
__global__ void subdomain_interface_kernel(const T* __restrict__ F, long long N, const T* __restrict__ fv, long long ntgt,
                                           const long long* __restrict__ tgt, const int* __restrict__ plus, const int* __restrict__ minus,
                                           T* __restrict__ out) {
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < ntgt; t += (long long)gridDim.x * blockDim.x) {
    const long long o = tgt[t], node = o % N;
    const int a = plus[t], b = minus[t];
    for (int c = 0; c < 3; ++c) {
      const T f = F[3 * node + c];
      T v;
      if (a >= 0 && b >= 0) v = f + (fv[3ll * a + c] - fv[3ll * b + c]);
      else if (a >= 0) v = f + fv[3ll * a + c];
      else if (b >= 0) v = f - fv[3ll * b + c];
      else v = f;
      out[3 * o + c] = v;
    }
  }
}

}  // namespace femb

using namespace femb;

extern "C" int femb_shell_normal(const void* coords, int fp, const void* conn, int ib, int64_t M, int nen, void* out, femb_stream stream) {
  FEMB_CHECK_ARG((fp == 4 || fp == 8) && (ib == 4 || ib == 8) && M >= 0 && (nen == 3 || nen == 4), "fp in {4,8}, ib in {4,8}, nen in {3,4}");
  if (M == 0) return FEMB_OK;
  cudaStream_t s = as_stream(stream);
  const int grid = grid_for(M, 128);
#define SN(T, I) shell_normal_kernel<T, I><<<grid, 128, 0, s>>>((const T*)coords, (const I*)conn, M, nen, (T*)out)
  if (fp == 8) {
    if (ib == 8) SN(double, long long);
    else SN(double, int);
  } else {
    if (ib == 8) SN(float, long long);
    else SN(float, int);
  }
#undef SN
  FEMB_LAUNCH_CHECK();
  return FEMB_OK;
}

extern "C" int femb_shell_rotate_operator(const void* K, const void* unit, int fp, int64_t M, int nd, void* out, femb_stream stream) {
  FEMB_CHECK_ARG((fp == 4 || fp == 8) && M >= 0 && nd >= 3 && nd % 3 == 0 && K != out, "fp in {4,8}, nd a multiple of 3, distinct buffers");
  if (M == 0) return FEMB_OK;
  cudaStream_t s = as_stream(stream);
  const int nb = nd / 3;
  const int grid = grid_for(M * nb * nb, 128);
  if (fp == 8) shell_rotate_K_kernel<double><<<grid, 128, 0, s>>>((const double*)K, (const double*)unit, M, nb, (double*)out);
  else shell_rotate_K_kernel<float><<<grid, 128, 0, s>>>((const float*)K, (const float*)unit, M, nb, (float*)out);
  FEMB_LAUNCH_CHECK();
  return FEMB_OK;
}

extern "C" int femb_subdomain_forces(const void* F, int fp, int64_t N, int n_sub, const void* free_vars, int64_t ntgt, const int64_t* tgt,
                                     const int32_t* plus, const int32_t* minus, void* out, femb_stream stream) {
  FEMB_CHECK_ARG((fp == 4 || fp == 8) && N >= 0 && n_sub >= 1 && ntgt >= 0, "fp in {4,8}, N >= 0, n_sub >= 1, ntgt >= 0");
  if (N == 0) return FEMB_OK;
  cudaStream_t s = as_stream(stream);
  const int g1 = grid_for(3 * N * n_sub, 256), g2 = grid_for(ntgt, 128);
  if (fp == 8) {
    subdomain_broadcast_kernel<double><<<g1, 256, 0, s>>>((const double*)F, 3 * N, n_sub, (double*)out);
    if (ntgt) subdomain_interface_kernel<double><<<g2, 128, 0, s>>>((const double*)F, N, (const double*)free_vars, ntgt, (const long long*)tgt, plus, minus, (double*)out);
  } else {
    subdomain_broadcast_kernel<float><<<g1, 256, 0, s>>>((const float*)F, 3 * N, n_sub, (float*)out);
    if (ntgt) subdomain_interface_kernel<float><<<g2, 128, 0, s>>>((const float*)F, N, (const float*)free_vars, ntgt, (const long long*)tgt, plus, minus, (float*)out);
  }
  FEMB_LAUNCH_CHECK();
  return FEMB_OK;
}

extern "C" int femb_shell_local_displacement(const void* conn, int ib, int64_t M, int nen, const void* disp, const void* unit, int fp, void* out,
                                             femb_stream stream) {
  FEMB_CHECK_ARG((fp == 4 || fp == 8) && (ib == 4 || ib == 8) && M >= 0 && nen >= 1, "fp in {4,8}, ib in {4,8}, M >= 0, nen >= 1");
  if (M == 0) return FEMB_OK;
  cudaStream_t s = as_stream(stream);
  const int grid = grid_for(M * nen, 128);
#define LD(T, I) shell_local_disp_kernel<T, I><<<grid, 128, 0, s>>>((const I*)conn, M, nen, (const T*)disp, (const T*)unit, (T*)out)
  if (fp == 8) {
    if (ib == 8) LD(double, long long);
    else LD(double, int);
  } else {
    if (ib == 8) LD(float, long long);
    else LD(float, int);
  }
#undef LD
  FEMB_LAUNCH_CHECK();
  return FEMB_OK;
}

extern "C" int femb_shell_local_coordinates(const void* coords, int fp, const void* conn, int ib, int64_t M, int nen, const void* unit, void* out,
                                            femb_stream stream) {
  FEMB_CHECK_ARG((fp == 4 || fp == 8) && (ib == 4 || ib == 8) && M >= 0 && nen >= 1, "fp in {4,8}, ib in {4,8}, M >= 0, nen >= 1");
  if (M == 0) return FEMB_OK;
  cudaStream_t s = as_stream(stream);
  const int grid = grid_for(M * nen, 128);
#define LC(T, I) shell_local_coords_kernel<T, I><<<grid, 128, 0, s>>>((const T*)coords, (const I*)conn, M, nen, (const T*)unit, (T*)out)
  if (fp == 8) {
    if (ib == 8) LC(double, long long);
    else LC(double, int);
  } else {
    if (ib == 8) LC(float, long long);
    else LC(float, int);
  }
#undef LC
  FEMB_LAUNCH_CHECK();
  return FEMB_OK;
}

extern "C" int femb_shell_postprocess(const void* nmq, int fp, int64_t M, int ncol, double t, double z, void* out, femb_stream stream) {
  FEMB_CHECK_ARG((fp == 4 || fp == 8) && M >= 0 && ncol >= 6, "fp in {4,8}, M >= 0, at least 6 columns (N, M resultants)");
  if (M == 0) return FEMB_OK;
  cudaStream_t s = as_stream(stream);
  const double f1 = 1.0 / t, f2 = 6.0 * z / (t * t);
  const int grid = grid_for(M, 256);
  if (fp == 8) shell_postprocess_kernel<double><<<grid, 256, 0, s>>>((const double*)nmq, M, ncol, f1, f2, (double*)out);
  else shell_postprocess_kernel<float><<<grid, 256, 0, s>>>((const float*)nmq, M, ncol, (float)f1, (float)f2, (float*)out);
  FEMB_LAUNCH_CHECK();
  return FEMB_OK;
}

extern "C" int femb_face_forces(const void* normals, const void* stress, int fp, int64_t M, int nf, void* out, femb_stream stream) {
  FEMB_CHECK_ARG((fp == 4 || fp == 8) && M >= 0 && nf >= 1, "fp in {4,8}, M >= 0, nf >= 1");
  if (M == 0) return FEMB_OK;
  cudaStream_t s = as_stream(stream);
  const int grid = grid_for(M * nf, 128);
  if (fp == 8) face_forces_kernel<double><<<grid, 128, 0, s>>>((const double*)normals, (const double*)stress, M, nf, (double*)out);
  else face_forces_kernel<float><<<grid, 128, 0, s>>>((const float*)normals, (const float*)stress, M, nf, (float*)out);
  FEMB_LAUNCH_CHECK();
  return FEMB_OK;
}

extern "C" int femb_shared_face_forces_sum(const int64_t* pairs, int64_t S, const void* forces, int fp, int nf, void* out, femb_stream stream) {
  FEMB_CHECK_ARG((fp == 4 || fp == 8) && S >= 0 && nf >= 1, "fp in {4,8}, S >= 0, nf >= 1");
  if (S == 0) return FEMB_OK;
  cudaStream_t s = as_stream(stream);
  const int grid = grid_for(S, 128);
  if (fp == 8) shared_forces_kernel<double><<<grid, 128, 0, s>>>((const long long*)pairs, S, (const double*)forces, nf, (double*)out);
  else shared_forces_kernel<float><<<grid, 128, 0, s>>>((const long long*)pairs, S, (const float*)forces, nf, (float*)out);
  FEMB_LAUNCH_CHECK();
  return FEMB_OK;
}

extern "C" int femb_wedge_face_normals(const void* coords, int fp, const void* conn, int ib, int64_t M, int conn_stride, void* out,
                                       femb_stream stream) {
  FEMB_CHECK_ARG((fp == 4 || fp == 8) && (ib == 4 || ib == 8) && M >= 0 && conn_stride >= 6, "fp in {4,8}, ib in {4,8}, M >= 0, stride >= 6");
  if (M == 0) return FEMB_OK;
  cudaStream_t s = as_stream(stream);
  const int grid = grid_for(M * 5, 128);
#define WN(T, I) wedge_normals_kernel<T, I><<<grid, 128, 0, s>>>((const T*)coords, (const I*)conn, M, conn_stride, (T*)out)
  if (fp == 8) {
    if (ib == 8) WN(double, long long);
    else WN(double, int);
  } else {
    if (ib == 8) WN(float, long long);
    else WN(float, int);
  }
#undef WN
  FEMB_LAUNCH_CHECK();
  return FEMB_OK;
}
