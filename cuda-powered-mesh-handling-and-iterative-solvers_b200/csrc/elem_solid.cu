// Solid element kernels: C3D4 (closed form) and the isoparametric family C3D10 / C3D8 / C3D6.
//
// Reference behaviour being reproduced (solver/element.py): compute_c3d4_{B,K}_matrix :835-903,
// compute_c3d10_* :1026-1239, compute_c3d8_* :1601-1803, compute_c3d6_* :2482-2676, volumes :514, :1248, :2198.
// Nothing here materialises B for the K paths: with isotropic D the node-pair block is
//   K_ab[i][j] = lambda g_a[i] g_b[j] + mu g_a[j] g_b[i] + mu delta_ij (g_a . g_b)
// which is B_a^T D B_b written out (Voigt order xx,yy,zz,xy,yz,zx, engineering shear).
#include "common.cuh"

namespace femb {

// ------------------------------------------------------------------------------------------------
// C3D4: one thread per element, 256-bit stores of complete 32-byte sectors.
// ------------------------------------------------------------------------------------------------
template <typename T>
struct Tet {
  T g[4][3];
  T det;
};

template <typename T, typename I>
__device__ __forceinline__ void tet_setup(const T* __restrict__ coords, const I* __restrict__ conn, long long e, Tet<T>& t) {
  long long n[4];
  if (sizeof(I) == 8) {  // one 256-bit load of the element's four int64 ids
    asm volatile("ld.global.nc.L1::no_allocate.v4.s64 {%0,%1,%2,%3}, [%4];" : "=l"(n[0]), "=l"(n[1]), "=l"(n[2]), "=l"(n[3]) : "l"(conn + 4 * e));
  } else {
    const int4 q = __ldg(reinterpret_cast<const int4*>(conn) + e);
    n[0] = q.x, n[1] = q.y, n[2] = q.z, n[3] = q.w;
  }
  T x[4][3];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int k = 0; k < 3; ++k) x[a][k] = __ldg(coords + 3 * n[a] + k);
  T e1[3], e2[3], e3[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    e1[k] = x[1][k] - x[0][k];
    e2[k] = x[2][k] - x[0][k];
    e3[k] = x[3][k] - x[0][k];
  }
  // cofactor columns of inv([e1;e2;e3])
  T c1[3] = {e2[1] * e3[2] - e2[2] * e3[1], e2[2] * e3[0] - e2[0] * e3[2], e2[0] * e3[1] - e2[1] * e3[0]};
  T c2[3] = {e3[1] * e1[2] - e3[2] * e1[1], e3[2] * e1[0] - e3[0] * e1[2], e3[0] * e1[1] - e3[1] * e1[0]};
  T c3[3] = {e1[1] * e2[2] - e1[2] * e2[1], e1[2] * e2[0] - e1[0] * e2[2], e1[0] * e2[1] - e1[1] * e2[0]};
  t.det = e1[0] * c1[0] + e1[1] * c1[1] + e1[2] * c1[2];
  const T id = T(1) / t.det;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    t.g[1][k] = c1[k] * id;
    t.g[2][k] = c2[k] * id;
    t.g[3][k] = c3[k] * id;
    t.g[0][k] = -(t.g[1][k] + t.g[2][k] + t.g[3][k]);
  }
}

template <typename T>
__device__ __forceinline__ void store_row12(T* p, const T* v) {
#pragma unroll
  for (int k = 0; k < 12; ++k) p[k] = v[k];
}
template <>
__device__ __forceinline__ void store_row12<double>(double* p, const double* v) {
  st256(p, v[0], v[1], v[2], v[3]);
  st256(p + 4, v[4], v[5], v[6], v[7]);
  st256(p + 8, v[8], v[9], v[10], v[11]);
}

template <typename T>
__device__ __forceinline__ void store_row4(T* p, const T* v) {
#pragma unroll
  for (int k = 0; k < 4; ++k) p[k] = v[k];
}
template <>
__device__ __forceinline__ void store_row4<double>(double* p, const double* v) {
  st256(p, v[0], v[1], v[2], v[3]);
}

// WHAT: 0 gradients [M,4,3]; 1 B [M,6,12]; 2 K [M,12,12]; 3 Poisson [M,4,4]; 4 mass [M,12,12]; 5 volume [M]
template <typename T, typename I, int WHAT>
__global__ void __launch_bounds__(128) c3d4_kernel(const T* __restrict__ coords, const I* __restrict__ conn, long long M,
                                                   T lam, T mu, T* __restrict__ out, int* __restrict__ flag) {
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < M; e += (long long)gridDim.x * blockDim.x) {
    Tet<T> t;
    tet_setup(coords, conn, e, t);
    if (WHAT != 5 && WHAT != 4 && fabs((double)t.det) < 1e-12 && flag) *flag = 1;
    const T V = fabs(t.det) / T(6);
    if (WHAT == 5) {
      out[e] = V;
    } else if (WHAT == 0) {
      T* o = out + e * 12;
      T v[12];
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int k = 0; k < 3; ++k) v[3 * a + k] = t.g[a][k];
      store_row12(o, v);
    } else if (WHAT == 1) {
      T* o = out + e * 72;
      // rows: xx, yy, zz, xy, yz, zx (element.py:870-879)
      const int c0[6] = {0, -1, -1, 1, -1, 2};  // which gradient component lands in dof column 0 / 1 / 2
      const int c1[6] = {-1, 1, -1, 0, 2, -1};
      const int c2[6] = {-1, -1, 2, -1, 1, 0};
#pragma unroll
      for (int r = 0; r < 6; ++r) {
        T v[12];
#pragma unroll
        for (int a = 0; a < 4; ++a) {
          v[3 * a + 0] = c0[r] >= 0 ? t.g[a][c0[r]] : T(0);
          v[3 * a + 1] = c1[r] >= 0 ? t.g[a][c1[r]] : T(0);
          v[3 * a + 2] = c2[r] >= 0 ? t.g[a][c2[r]] : T(0);
        }
        store_row12(o + 12 * r, v);
      }
    } else if (WHAT == 2) {
      T* o = out + e * 144;
#pragma unroll
      for (int a = 0; a < 4; ++a) {
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          T v[12];
#pragma unroll
          for (int b = 0; b < 4; ++b) {
            const T dot = t.g[a][0] * t.g[b][0] + t.g[a][1] * t.g[b][1] + t.g[a][2] * t.g[b][2];
#pragma unroll
            for (int j = 0; j < 3; ++j) {
              T k = lam * t.g[a][i] * t.g[b][j] + mu * t.g[a][j] * t.g[b][i];
              if (i == j) k += mu * dot;
              v[3 * b + j] = k * V;
            }
          }
          store_row12(o + 12 * (3 * a + i), v);
        }
      }
    } else if (WHAT == 3) {
      T* o = out + e * 16;
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        T v[4];
#pragma unroll
        for (int b = 0; b < 4; ++b) v[b] = (t.g[a][0] * t.g[b][0] + t.g[a][1] * t.g[b][1] + t.g[a][2] * t.g[b][2]) * V;
        store_row4(o + 4 * a, v);
      }
    } else if (WHAT == 4) {
      T* o = out + e * 144;
      const T m = lam * V / T(20);  // lam carries rho
#pragma unroll
      for (int r = 0; r < 12; ++r) {
        T v[12];
#pragma unroll
        for (int c = 0; c < 12; ++c) v[c] = (r % 3 == c % 3) ? (r == c ? 2 * m : m) : T(0);
        store_row12(o + 12 * r, v);
      }
    }
  }
}

template <typename T, typename I>
static int c3d4_dispatch(int what, const void* coords, const void* conn, long long M, double E, double nu, void* out, int* flag,
                         cudaStream_t s) {
  if (M == 0) return FEMB_OK;
  const double c = E / ((1 + nu) * (1 - 2 * nu));
  T lam = (T)(c * nu), mu = (T)(c * (1 - 2 * nu) / 2);
  if (what == 4) lam = (T)E;
  const int grid = grid_for(M, 128);
  const T* X = static_cast<const T*>(coords);
  const I* C = static_cast<const I*>(conn);
  T* O = static_cast<T*>(out);
  switch (what) {
    case 0: c3d4_kernel<T, I, 0><<<grid, 128, 0, s>>>(X, C, M, lam, mu, O, flag); break;
    case 1: c3d4_kernel<T, I, 1><<<grid, 128, 0, s>>>(X, C, M, lam, mu, O, flag); break;
    case 2: c3d4_kernel<T, I, 2><<<grid, 128, 0, s>>>(X, C, M, lam, mu, O, flag); break;
    case 3: c3d4_kernel<T, I, 3><<<grid, 128, 0, s>>>(X, C, M, lam, mu, O, flag); break;
    case 4: c3d4_kernel<T, I, 4><<<grid, 128, 0, s>>>(X, C, M, lam, mu, O, flag); break;
    case 5: c3d4_kernel<T, I, 5><<<grid, 128, 0, s>>>(X, C, M, lam, mu, O, flag); break;
    default: set_error("femb_c3d4: unknown `what`"); return FEMB_ERR_ARG;
  }
  FEMB_LAUNCH_CHECK();
  return FEMB_OK;
}

// ------------------------------------------------------------------------------------------------
// Volumes through the reference's sub-tet tables
// ------------------------------------------------------------------------------------------------
__constant__ int c_vol_tets[2][6][4] = {
    {{0, 1, 3, 4}, {1, 2, 3, 6}, {1, 3, 4, 5}, {3, 4, 5, 7}, {3, 5, 6, 7}, {3, 5, 6, 1}},   // hex  (element.py:1282-1287)
    {{0, 1, 2, 3}, {1, 2, 4, 3}, {2, 4, 5, 3}, {0, 0, 0, 0}, {0, 0, 0, 0}, {0, 0, 0, 0}}};  // wedge (element.py:2226-2228)

template <typename T>
__device__ __forceinline__ T abs_tet_vol(const T (*x)[3], const int* t) {
  T a[3], b[3], c[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    a[k] = x[t[1]][k] - x[t[0]][k];
    b[k] = x[t[2]][k] - x[t[0]][k];
    c[k] = x[t[3]][k] - x[t[0]][k];
  }
  const T d = a[0] * (b[1] * c[2] - b[2] * c[1]) - a[1] * (b[0] * c[2] - b[2] * c[0]) + a[2] * (b[0] * c[1] - b[1] * c[0]);
  return fabs(d) / T(6);
}

template <typename T, typename I, int NEN>
__global__ void volumes_kernel(const T* __restrict__ coords, const I* __restrict__ conn, long long M, int stride, T* __restrict__ vol) {
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < M; e += (long long)gridDim.x * blockDim.x) {
    T x[NEN][3];
#pragma unroll
    for (int a = 0; a < NEN; ++a) {
      const long long n = ldidx(conn + e * stride + a);
#pragma unroll
      for (int k = 0; k < 3; ++k) x[a][k] = __ldg(coords + 3 * n + k);
    }
    if (NEN == 4) {
      const int t[4] = {0, 1, 2, 3};
      vol[e] = abs_tet_vol<T>(x, t);
    } else {
      const int which = NEN == 8 ? 0 : 1, nt = NEN == 8 ? 6 : 3;
      T v = 0;
      for (int s = 0; s < nt; ++s) v += abs_tet_vol<T>(x, c_vol_tets[which][s]);
      vol[e] = v;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Isoparametric solids.  One CTA = EPB elements.
//   phase A: thread (element, point)  -> J^-1, gradients, detJ*w into shared memory
//   phase B: thread (element, a<=b)   -> 3x3 node-pair block accumulated over the points, written (and mirrored)
//            into a shared K tile
//   phase C: the whole CTA streams the tile to global memory with coalesced 16-byte stores
// ------------------------------------------------------------------------------------------------
struct SolidTab {
  const double* dN;  // [nq][nen][3]
  const double* w;   // [nq]
};

template <typename T, typename I, int NEN, int EPB>
__global__ void __launch_bounds__(512) solid_K_kernel(const T* __restrict__ coords, const I* __restrict__ conn, long long M, SolidTab tab, int nq, int mode,
                               T lam, T mu, T* __restrict__ out) {
  constexpr int ND = 3 * NEN;
  constexpr int NPAIR = NEN * (NEN + 1) / 2;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* xs = reinterpret_cast<T*>(smem_raw);         // [EPB][NEN][3]
  T* gs = xs + EPB * NEN * 3;                     // [EPB][nq][NEN][3]
  T* wd = gs + (size_t)EPB * nq * NEN * 3;        // [EPB][nq]
  T* kt = wd + EPB * nq;                          // [EPB][ND][ND]
  const int tid = threadIdx.x;
  const int nslice = mode == 4 ? nq : 1;

  for (long long e0 = (long long)blockIdx.x * EPB; e0 < M; e0 += (long long)gridDim.x * EPB) {
    const int ne = (int)min((long long)EPB, M - e0);
    __syncthreads();
    for (int t = tid; t < ne * NEN; t += blockDim.x) {
      const long long n = ldidx(conn + e0 * NEN + t);
#pragma unroll
      for (int k = 0; k < 3; ++k) xs[t * 3 + k] = __ldg(coords + 3 * n + k);
    }
    __syncthreads();
    // ---- phase A
    for (int t = tid; t < ne * nq; t += blockDim.x) {
      const int el = t / nq, q = t - el * nq;
      const T* x = xs + el * NEN * 3;
      const double* dn = tab.dN + (size_t)q * NEN * 3;
      T J[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
      for (int a = 0; a < NEN; ++a)
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          const T d = (T)dn[a * 3 + i];
#pragma unroll
          for (int k = 0; k < 3; ++k) J[i * 3 + k] += d * x[a * 3 + k];
        }
      T Ji[9];
      const T det = inv3(J, Ji);
      T* g = gs + ((size_t)el * nq + q) * NEN * 3;
#pragma unroll
      for (int a = 0; a < NEN; ++a)
#pragma unroll
        for (int i = 0; i < 3; ++i)
          g[a * 3 + i] = Ji[i * 3 + 0] * (T)dn[a * 3 + 0] + Ji[i * 3 + 1] * (T)dn[a * 3 + 1] + Ji[i * 3 + 2] * (T)dn[a * 3 + 2];
      T wt;
      if (mode == 5) {  // C3D6 single=True: |volume| of the reference's 3-tet split (element.py:2226-2228, 2652-2656)
        T xx[6][3];
        for (int a = 0; a < 6 && a < NEN; ++a)
          for (int k = 0; k < 3; ++k) xx[a][k] = x[a * 3 + k];
        wt = 0;
        for (int s = 0; s < 3; ++s) wt += abs_tet_vol<T>(xx, c_vol_tets[1][s]);
      } else {
        wt = mode == 4 ? det : det * (T)tab.w[q];
      }
      wd[el * nq + q] = wt;
    }
    __syncthreads();
    for (int slice = 0; slice < nslice; ++slice) {
      // ---- phase B
      for (int t = tid; t < ne * NPAIR; t += blockDim.x) {
        const int el = t / NPAIR;
        int p = t - el * NPAIR;
        int a = 0;
        while (p >= NEN - a) {
          p -= NEN - a;
          ++a;
        }
        const int b = a + p;
        T acc[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
        const int q0 = mode == 4 ? slice : 0, q1 = mode == 4 ? slice + 1 : nq;
        for (int q = q0; q < q1; ++q) {
          const T* g = gs + ((size_t)el * nq + q) * NEN * 3;
          const T w = wd[el * nq + q];
          const T ga[3] = {g[a * 3], g[a * 3 + 1], g[a * 3 + 2]};
          const T gb[3] = {g[b * 3], g[b * 3 + 1], g[b * 3 + 2]};
          const T dot = mu * (ga[0] * gb[0] + ga[1] * gb[1] + ga[2] * gb[2]);
#pragma unroll
          for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) {
              T k = lam * ga[i] * gb[j] + mu * ga[j] * gb[i];
              if (i == j) k += dot;
              acc[i * 3 + j] += k * w;
            }
        }
        T* K = kt + (size_t)el * ND * ND;
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
          for (int j = 0; j < 3; ++j) {
            K[(3 * a + i) * ND + 3 * b + j] = acc[i * 3 + j];
            K[(3 * b + j) * ND + 3 * a + i] = acc[i * 3 + j];
          }
      }
      __syncthreads();
      // ---- phase C
      T* dst = out + ((size_t)slice * M + e0) * ND * ND;
      const int total = ne * ND * ND;
      if (sizeof(T) == 8) {
        for (int t = 2 * tid; t < total; t += 2 * blockDim.x)
          st128(reinterpret_cast<double*>(dst) + t, (double)kt[t], (double)kt[t + 1]);
      } else {
        for (int t = tid; t < total; t += blockDim.x) dst[t] = kt[t];
      }
      __syncthreads();
    }
  }
}

// what 0: J [M,3,3]; 1: gradients [M,NEN,3]; 2: B [M,6,3*NEN] -- single point, one thread per element
template <typename T, typename I, int NEN>
__global__ void solid_point_kernel(const T* __restrict__ coords, const I* __restrict__ conn, long long M, SolidTab tab, int what,
                                   T* __restrict__ out) {
  constexpr int ND = 3 * NEN;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < M; e += (long long)gridDim.x * blockDim.x) {
    T J[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int a = 0; a < NEN; ++a) {
      const long long n = ldidx(conn + e * NEN + a);
      T x[3] = {__ldg(coords + 3 * n), __ldg(coords + 3 * n + 1), __ldg(coords + 3 * n + 2)};
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        const T d = (T)tab.dN[a * 3 + i];
#pragma unroll
        for (int k = 0; k < 3; ++k) J[i * 3 + k] += d * x[k];
      }
    }
    if (what == 0) {
      for (int k = 0; k < 9; ++k) out[e * 9 + k] = J[k];
      continue;
    }
    T Ji[9];
    inv3(J, Ji);
    if (what == 2)
      for (int k = 0; k < 6 * ND; ++k) out[e * 6 * ND + k] = 0;
    for (int a = 0; a < NEN; ++a) {
      T g[3];
#pragma unroll
      for (int i = 0; i < 3; ++i)
        g[i] = Ji[i * 3 + 0] * (T)tab.dN[a * 3 + 0] + Ji[i * 3 + 1] * (T)tab.dN[a * 3 + 1] + Ji[i * 3 + 2] * (T)tab.dN[a * 3 + 2];
      if (what == 1) {
        for (int i = 0; i < 3; ++i) out[(e * NEN + a) * 3 + i] = g[i];
      } else {
        T* B = out + e * 6 * ND;
        B[0 * ND + 3 * a + 0] = g[0];
        B[1 * ND + 3 * a + 1] = g[1];
        B[2 * ND + 3 * a + 2] = g[2];
        B[3 * ND + 3 * a + 0] = g[1];
        B[3 * ND + 3 * a + 1] = g[0];
        B[4 * ND + 3 * a + 1] = g[2];
        B[4 * ND + 3 * a + 2] = g[1];
        B[5 * ND + 3 * a + 0] = g[2];
        B[5 * ND + 3 * a + 2] = g[0];
      }
    }
  }
}

// natural-coordinate derivative tables, evaluated on the host in fp64 exactly as written in the reference
static void dN_c3d10(const double* p, double* o) {  // element.py:1042-1055
  const double xi = p[0], eta = p[1], zeta = p[2], n3 = -4 * (1 - xi - eta - zeta) + 1;
  const double t[30] = {4 * xi - 1, 0, 0, 0, 4 * eta - 1, 0, 0, 0, 4 * zeta - 1, n3, n3, n3, 4 * eta, 4 * xi, 0,
                        0, 4 * zeta, 4 * eta, 4 * zeta, 0, 4 * xi, 4 * (1 - 2 * xi - eta - zeta), -4 * xi, -4 * xi,
                        -4 * eta, 4 * (1 - xi - 2 * eta - zeta), -4 * eta, -4 * zeta, -4 * zeta, 4 * (1 - xi - eta - 2 * zeta)};
  memcpy(o, t, sizeof(t));
}
static void dN_c3d8(const double* p, double* o) {  // element.py:1617-1626
  static const int sx[8] = {-1, 1, 1, -1, -1, 1, 1, -1}, sy[8] = {-1, -1, 1, 1, -1, -1, 1, 1}, sz[8] = {-1, -1, -1, -1, 1, 1, 1, 1};
  for (int a = 0; a < 8; ++a) {
    o[3 * a + 0] = 0.125 * sx[a] * (1 + sy[a] * p[1]) * (1 + sz[a] * p[2]);
    o[3 * a + 1] = 0.125 * sy[a] * (1 + sx[a] * p[0]) * (1 + sz[a] * p[2]);
    o[3 * a + 2] = 0.125 * sz[a] * (1 + sx[a] * p[0]) * (1 + sy[a] * p[1]);
  }
}
static void dN_c3d6(const double* p, double* o) {  // element.py:2499-2506
  const double r = p[0], s = p[1], t = p[2];
  const double v[18] = {-0.5 * (1 - t), -0.5 * (1 - t), -0.5 * (1 - r - s), 0.5 * (1 - t), 0.0, -0.5 * r,
                        0.0, 0.5 * (1 - t), -0.5 * s, -0.5 * (1 + t), -0.5 * (1 + t), 0.5 * (1 - r - s),
                        0.5 * (1 + t), 0.0, 0.5 * r, 0.0, 0.5 * (1 + t), 0.5 * s};
  memcpy(o, v, sizeof(v));
}

template <typename T, typename I, int NEN>
static int solid_dispatch(int what, const void* coords, const void* conn, long long M, const double* pts, int nq, double E, double nu,
                          void* out, cudaStream_t s) {
  if (M == 0) return FEMB_OK;
  FEMB_CHECK_ARG(nq >= 1 && nq <= 64, "1 <= nq <= 64");
  double tabh[64 * 20 * 3 + 64];
  for (int q = 0; q < nq; ++q) {
    double* o = tabh + (size_t)q * NEN * 3;
    if (NEN == 10) dN_c3d10(pts + 4 * q, o);
    if (NEN == 8) dN_c3d8(pts + 4 * q, o);
    if (NEN == 6) dN_c3d6(pts + 4 * q, o);
  }
  double* wh = tabh + (size_t)nq * NEN * 3;
  for (int q = 0; q < nq; ++q) wh[q] = pts[4 * q + 3];
  Scratch scr(s);
  double* tabd;
  const size_t ntab = (size_t)nq * NEN * 3 + nq;
  FEMB_CUDA(scr.alloc(&tabd, ntab));
  FEMB_CUDA(cudaMemcpyAsync(tabd, tabh, ntab * sizeof(double), cudaMemcpyHostToDevice, s));
  SolidTab tab{tabd, tabd + (size_t)nq * NEN * 3};
  const T* X = static_cast<const T*>(coords);
  const I* C = static_cast<const I*>(conn);
  T* O = static_cast<T*>(out);
  if (what <= 2) {
    solid_point_kernel<T, I, NEN><<<grid_for(M, 128), 128, 0, s>>>(X, C, M, tab, what, O);
    FEMB_LAUNCH_CHECK();
    return FEMB_OK;
  }
  FEMB_CHECK_ARG(what >= 3 && what <= 5, "femb_solid: what in 0..5");
  FEMB_CHECK_ARG(what != 5 || NEN == 6, "what=5 is the C3D6 single-point mode");
  const int nqk = what == 5 ? 1 : nq;
  const double c = E / ((1 + nu) * (1 - 2 * nu));
  const T lam = (T)(c * nu), mu = (T)(c * (1 - 2 * nu) / 2);
  constexpr int EPB = 8, ND = 3 * NEN, NPAIR = NEN * (NEN + 1) / 2;
  const size_t smem = sizeof(T) * ((size_t)EPB * NEN * 3 + (size_t)EPB * nqk * NEN * 3 + (size_t)EPB * nqk + (size_t)EPB * ND * ND);
  auto kern = solid_K_kernel<T, I, NEN, EPB>;
  FEMB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int threads = EPB * NPAIR;
  threads = ((threads + 31) / 32) * 32;
  if (threads > 512) threads = 512;
  const int grid = (int)std::min<long long>((M + EPB - 1) / EPB, (long long)SMS * 16);
  kern<<<grid, threads, smem, s>>>(X, C, M, tab, nqk, what, lam, mu, O);
  FEMB_LAUNCH_CHECK();
  return FEMB_OK;
}

template <typename I>
__global__ void to_c3d4_kernel(const I* __restrict__ conn, long long M, int nen, int k, const int* __restrict__ table, long long* __restrict__ out) {
  const long long total = M * k * 4;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const long long e = t / (k * 4);
    const int r = (int)(t - e * k * 4);
    out[t] = ldidx(conn + e * nen + table[r]);
  }
}

}  // namespace femb

using namespace femb;

#define DISPATCH_TI(fn, fp, ib, ...)                                        \
  ((fp) == 8 ? ((ib) == 8 ? fn<double, long long>(__VA_ARGS__) : fn<double, int>(__VA_ARGS__)) \
             : ((ib) == 8 ? fn<float, long long>(__VA_ARGS__) : fn<float, int>(__VA_ARGS__)))

extern "C" int femb_c3d4(int what, const void* coords, int fp, const void* conn, int ib, int64_t M, double E, double nu, void* out,
                         int32_t* flag, femb_stream stream) {
  FEMB_CHECK_ARG((fp == 4 || fp == 8) && (ib == 4 || ib == 8), "fp in {4,8}, ib in {4,8}");
  FEMB_CHECK_ARG(M >= 0, "M >= 0");
  return DISPATCH_TI(c3d4_dispatch, fp, ib, what, coords, conn, M, E, nu, out, flag, as_stream(stream));
}

template <typename T, typename I>
static int volumes_dispatch(int kind, const void* coords, const void* conn, long long M, int stride, void* vol, cudaStream_t s) {
  if (M == 0) return FEMB_OK;
  const int grid = grid_for(M, 128);
  const T* X = static_cast<const T*>(coords);
  const I* C = static_cast<const I*>(conn);
  if (kind == FEMB_C3D4 || kind == FEMB_C3D10) volumes_kernel<T, I, 4><<<grid, 128, 0, s>>>(X, C, M, stride, (T*)vol);
  else if (kind == FEMB_C3D8) volumes_kernel<T, I, 8><<<grid, 128, 0, s>>>(X, C, M, stride, (T*)vol);
  else if (kind == FEMB_C3D6) volumes_kernel<T, I, 6><<<grid, 128, 0, s>>>(X, C, M, stride, (T*)vol);
  else {
    set_error("femb_elem_volumes: kind must be C3D4/C3D10/C3D8/C3D6");
    return FEMB_ERR_ARG;
  }
  FEMB_LAUNCH_CHECK();
  return FEMB_OK;
}

extern "C" int femb_elem_volumes(int kind, const void* coords, int fp, const void* conn, int ib, int64_t M, int conn_stride, void* vol,
                                 femb_stream stream) {
  FEMB_CHECK_ARG((fp == 4 || fp == 8) && (ib == 4 || ib == 8), "fp in {4,8}, ib in {4,8}");
  return DISPATCH_TI(volumes_dispatch, fp, ib, kind, coords, conn, M, conn_stride, vol, as_stream(stream));
}

template <typename T, typename I>
static int solid_dispatch_kind(int kind, int what, const void* coords, const void* conn, long long M, const double* pts, int nq, double E,
                               double nu, void* out, cudaStream_t s) {
  switch (kind) {
    case FEMB_C3D10: return solid_dispatch<T, I, 10>(what, coords, conn, M, pts, nq, E, nu, out, s);
    case FEMB_C3D8: return solid_dispatch<T, I, 8>(what, coords, conn, M, pts, nq, E, nu, out, s);
    case FEMB_C3D6: return solid_dispatch<T, I, 6>(what, coords, conn, M, pts, nq, E, nu, out, s);
  }
  set_error("femb_solid: kind must be C3D10/C3D8/C3D6");
  return FEMB_ERR_ARG;
}

extern "C" int femb_solid(int kind, int what, const void* coords, int fp, const void* conn, int ib, int64_t M, const double* pts_host,
                          int nq, double E, double nu, void* out, femb_stream stream) {
  FEMB_CHECK_ARG((fp == 4 || fp == 8) && (ib == 4 || ib == 8), "fp in {4,8}, ib in {4,8}");
  FEMB_CHECK_ARG(pts_host != nullptr, "pts_host");
  return DISPATCH_TI(solid_dispatch_kind, fp, ib, kind, what, coords, conn, M, pts_host, nq, E, nu, out, as_stream(stream));
}

extern "C" int femb_default_points(int kind, double* p) {
  // fp32-rounded 1/sqrt(3) (quirk q1: element.py:2460-2463, shell.py:657-663)
  const double g32 = (double)(1.0f / sqrtf(3.0f));
  if (kind == FEMB_C3D10) {  // element.py:995-1024
    static const double pts[11][4] = {{.25, .25, .25, .1}, {.1, .1, .1, .05}, {.1, .1, .7, .05}, {.1, .7, .1, .05},
                                      {.7, .1, .1, .05},   {.1, .4, .4, .03}, {.4, .1, .4, .03}, {.4, .4, .1, .03},
                                      {.3, .3, .3, .02},   {.2, .2, .6, .02}, {.2, .6, .2, .02}};
    memcpy(p, pts, sizeof(pts));
    return 11;
  }
  if (kind == FEMB_C3D8) {  // element.py:1583-1599, exact in fp64
    const double g = 1.0 / sqrt(3.0);
    int q = 0;
    for (int a = -1; a <= 1; a += 2)
      for (int b = -1; b <= 1; b += 2)
        for (int c = -1; c <= 1; c += 2, ++q) {
          p[4 * q] = a * g, p[4 * q + 1] = b * g, p[4 * q + 2] = c * g, p[4 * q + 3] = 1.0;
        }
    return 8;
  }
  if (kind == FEMB_C3D6) {  // element.py:2448-2480
    const double tri[3][2] = {{1.0 / 6, 1.0 / 6}, {2.0 / 3, 1.0 / 6}, {1.0 / 6, 2.0 / 3}};
    int q = 0;
    for (int i = 0; i < 3; ++i)
      for (int j = -1; j <= 1; j += 2, ++q) {
        p[4 * q] = tri[i][0], p[4 * q + 1] = tri[i][1], p[4 * q + 2] = j * g32, p[4 * q + 3] = 1.0 / 3;
      }
    return 6;
  }
  if (kind == FEMB_S4) {  // shell.py:651-672
    const double s[4][2] = {{1, -1}, {1, 1}, {-1, 1}, {-1, -1}};
    for (int q = 0; q < 4; ++q) p[4 * q] = s[q][0] * g32, p[4 * q + 1] = s[q][1] * g32, p[4 * q + 2] = 0, p[4 * q + 3] = 1.0;
    return 4;
  }
  set_error("femb_default_points: unknown kind");
  return -1;
}

extern "C" int femb_to_c3d4(int kind, const void* conn, int ib, int64_t M, int64_t* out, femb_stream stream) {
  static const int t10[32] = {0, 4, 6, 7, 4, 1, 5, 8, 6, 5, 2, 9, 7, 8, 9, 3, 4, 6, 7, 5, 6, 7, 9, 5, 4, 7, 8, 5, 5, 8, 7, 9};  // :977-986
  static const int t8[24] = {0, 1, 3, 4, 1, 2, 3, 6, 1, 3, 4, 5, 3, 4, 5, 7, 3, 5, 6, 7, 3, 5, 6, 2};                          // :1567-1574
  static const int t6[12] = {0, 1, 2, 3, 1, 2, 3, 5, 1, 3, 4, 5};                                                              // :2435-2439
  const int* tab;
  int k, nen;
  if (kind == FEMB_C3D10) tab = t10, k = 8, nen = 10;
  else if (kind == FEMB_C3D8) tab = t8, k = 6, nen = 8;
  else if (kind == FEMB_C3D6) tab = t6, k = 3, nen = 6;
  else {
    set_error("femb_to_c3d4: kind must be C3D10/C3D8/C3D6");
    return FEMB_ERR_ARG;
  }
  if (M == 0) return FEMB_OK;
  cudaStream_t s = as_stream(stream);
  Scratch scr(s);
  int* dtab;
  FEMB_CUDA(scr.alloc(&dtab, (size_t)k * 4));
  FEMB_CUDA(cudaMemcpyAsync(dtab, tab, sizeof(int) * k * 4, cudaMemcpyHostToDevice, s));
  const int grid = grid_for(M * k * 4, 256);
  if (ib == 8) to_c3d4_kernel<long long><<<grid, 256, 0, s>>>((const long long*)conn, M, nen, k, dtab, (long long*)out);
  else to_c3d4_kernel<int><<<grid, 256, 0, s>>>((const int*)conn, M, nen, k, dtab, (long long*)out);
  FEMB_LAUNCH_CHECK();
  return FEMB_OK;
}
