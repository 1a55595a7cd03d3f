// Kirchhoff shell elements S3 / S4 (reference solver/shell.py:297-453 and :597-861).
// One thread per element: local frame -> 2x2 Jacobian -> gradients -> B / K.  The reference contracts the parametric
// derivatives with Jinv (not its transpose, shell.py:399-401 and :743-744); that is reproduced as is.
#include "common.cuh"

namespace femb {

template <typename T>
__device__ __forceinline__ T norm3(const T* v) { return sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]); }

// unit rows (e1,e2,e3); S3: orthogonalise b against the un-normalised a (shell.py:311-318),
// S4: normalise a first and take b from node 3 (shell.py:611-618)
template <typename T, int NEN>
__device__ __forceinline__ void shell_frame(const T (*x)[3], T* u) {
  T a[3], b[3];
  const int nb = NEN == 3 ? 2 : 3;
  for (int k = 0; k < 3; ++k) {
    a[k] = x[1][k] - x[0][k];
    b[k] = x[nb][k] - x[0][k];
  }
  if (NEN == 4) {
    const T na = norm3(a);
    for (int k = 0; k < 3; ++k) a[k] /= na;
  }
  const T f = (a[0] * b[0] + a[1] * b[1] + a[2] * b[2]) / (a[0] * a[0] + a[1] * a[1] + a[2] * a[2]);
  for (int k = 0; k < 3; ++k) b[k] -= f * a[k];
  if (NEN == 3) {
    const T na = norm3(a);
    for (int k = 0; k < 3; ++k) a[k] /= na;
  }
  const T nbn = norm3(b);
  for (int k = 0; k < 3; ++k) b[k] /= nbn;
  u[0] = a[0], u[1] = a[1], u[2] = a[2];
  u[3] = b[0], u[4] = b[1], u[5] = b[2];
  u[6] = a[1] * b[2] - a[2] * b[1];
  u[7] = a[2] * b[0] - a[0] * b[2];
  u[8] = a[0] * b[1] - a[1] * b[0];
}

// f = (1-xi, 1+xi, 1-eta, 1+eta): the four bilinear factors arrive evaluated by the host, because the reference forms them in
// the dtype of whatever it was handed -- python doubles on the K path (`.item()`, shell.py:841-842), 0-dim fp32 tensors when
// compute_s4_B_matrix iterates over its fp32 rule (shell.py:813-814 -> :683-695)
template <int NEN>
__device__ __forceinline__ void shell_dparam(const double* f, double* dxi, double* deta) {
  if (NEN == 3) {
    dxi[0] = -1, dxi[1] = 1, dxi[2] = 0;
    deta[0] = -1, deta[1] = 0, deta[2] = 1;
  } else {
    dxi[0] = 0.25 * -f[2], dxi[1] = 0.25 * f[2], dxi[2] = 0.25 * f[3], dxi[3] = 0.25 * -f[3];
    deta[0] = 0.25 * -f[0], deta[1] = 0.25 * -f[1], deta[2] = 0.25 * f[1], deta[3] = 0.25 * f[0];
  }
}

struct ShellPts {
  double f[16][4];  // 1-xi, 1+xi, 1-eta, 1+eta
  double w[16];
  double D[36];
};

// B [6, 6 NEN] from the nodal gradients (shell.py:417-437, :772-798); `stride` > 1 writes one point of a [.., nq] layout
template <typename T, int NEN>
__device__ __forceinline__ void shell_store_B(T* __restrict__ B, int stride, const T* gx, const T* gy) {
  constexpr int ND = 6 * NEN;
  for (int k = 0; k < 6 * ND; ++k) B[(size_t)k * stride] = 0;
  for (int a = 0; a < NEN; ++a) {
    B[(size_t)(0 * ND + 6 * a + 0) * stride] = gx[a];
    B[(size_t)(1 * ND + 6 * a + 1) * stride] = gy[a];
    B[(size_t)(2 * ND + 6 * a + 0) * stride] = gy[a];
    B[(size_t)(2 * ND + 6 * a + 1) * stride] = gx[a];
    B[(size_t)(3 * ND + 6 * a + 4) * stride] = -gx[a];
    B[(size_t)(4 * ND + 6 * a + 3) * stride] = gy[a];
    B[(size_t)(5 * ND + 6 * a + 3) * stride] = gy[a];
    B[(size_t)(5 * ND + 6 * a + 4) * stride] = gx[a];
  }
}

// what 0 unit [M,3,3]; 1 J [M,2,2]; 2 grads [M,NEN,2]; 3 B [M,6,6NEN]; 4 K [M,ND,ND]; 5 per-point K [M,ND,ND,nq];
//      7 sum_q w_q B_q [M,6,ND]; 8 per-point w_q B_q [M,6,ND,nq];
//      9 stress resultants [M,6] = D (sum_q w_q B_q) u_e with u_e = disp[conn] taken in GLOBAL axes as the reference does
template <typename T, typename I, int NEN>
__global__ void __launch_bounds__(64) shell_kernel(const T* __restrict__ coords, const I* __restrict__ conn, long long M, ShellPts sp, int nq,
                                                   int what, const T* __restrict__ disp, T* __restrict__ out) {
  constexpr int ND = 6 * NEN;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < M; e += (long long)gridDim.x * blockDim.x) {
    T x[NEN][3];
    for (int a = 0; a < NEN; ++a) {
      const long long n = ldidx(conn + e * NEN + a);
      for (int k = 0; k < 3; ++k) x[a][k] = __ldg(coords + 3 * n + k);
    }
    T u[9];
    shell_frame<T, NEN>(x, u);
    if (what == 0) {
      for (int k = 0; k < 9; ++k) out[e * 9 + k] = u[k];
      continue;
    }
    T lx[NEN], ly[NEN];
    for (int a = 0; a < NEN; ++a) {
      const T v[3] = {x[a][0] - x[0][0], x[a][1] - x[0][1], x[a][2] - x[0][2]};
      lx[a] = v[0] * u[0] + v[1] * u[1] + v[2] * u[2];
      ly[a] = v[0] * u[3] + v[1] * u[4] + v[2] * u[5];
    }
    T* Ko = out + e * (size_t)ND * ND * (what == 5 ? nq : 1);
    const int npts = (what >= 4) ? nq : 1;
    T Gx[NEN], Gy[NEN];  // what 7 / 9: weighted sums of the gradients over the rule
    for (int a = 0; a < NEN; ++a) Gx[a] = 0, Gy[a] = 0;
    for (int q = 0; q < npts; ++q) {
      double dxi_d[4], deta_d[4];
      shell_dparam<NEN>(sp.f[q], dxi_d, deta_d);
      T dxi[NEN], deta[NEN];
      for (int a = 0; a < NEN; ++a) dxi[a] = (T)dxi_d[a], deta[a] = (T)deta_d[a];
      T J[4] = {0, 0, 0, 0};
      for (int a = 0; a < NEN; ++a) {
        J[0] += dxi[a] * lx[a];
        J[1] += deta[a] * lx[a];
        J[2] += dxi[a] * ly[a];
        J[3] += deta[a] * ly[a];
      }
      if (what == 1) {
        for (int k = 0; k < 4; ++k) out[e * 4 + k] = J[k];
        break;
      }
      const T det = J[0] * J[3] - J[1] * J[2];
      const T Ji[4] = {J[3] / det, -J[1] / det, -J[2] / det, J[0] / det};
      T gx[NEN], gy[NEN];
      for (int a = 0; a < NEN; ++a) {
        gx[a] = Ji[0] * dxi[a] + Ji[1] * deta[a];
        gy[a] = Ji[2] * dxi[a] + Ji[3] * deta[a];
      }
      if (what == 2) {
        for (int a = 0; a < NEN; ++a) out[(e * NEN + a) * 2] = gx[a], out[(e * NEN + a) * 2 + 1] = gy[a];
        break;
      }
      if (what == 3) {
        shell_store_B<T, NEN>(out + e * 6 * ND, 1, gx, gy);
        break;
      }
      if (what >= 7) {
        const T wq = NEN == 3 ? T(1) : (T)sp.w[q];
        if (what == 8) {
          T wx[NEN], wy[NEN];
          for (int a = 0; a < NEN; ++a) wx[a] = gx[a] * wq, wy[a] = gy[a] * wq;
          shell_store_B<T, NEN>(out + e * 6 * ND * nq + q, nq, wx, wy);
        } else {
          for (int a = 0; a < NEN; ++a) Gx[a] += gx[a] * wq, Gy[a] += gy[a] * wq;
        }
        continue;
      }
      // K_ab = B_a^T D B_b * detJ * w ; only dof rows/cols {0,1,3,4} are populated
      const T wt = det * (NEN == 3 ? T(0.5) : (T)sp.w[q]);
      for (int a = 0; a < NEN; ++a) {
        // Ba columns for dofs 0,1,3,4 as 6-vectors
        const T Ba[4][6] = {{gx[a], 0, gy[a], 0, 0, 0}, {0, gy[a], gx[a], 0, 0, 0}, {0, 0, 0, 0, gy[a], gy[a]}, {0, 0, 0, -gx[a], 0, gx[a]}};
        for (int b = 0; b < NEN; ++b) {
          const T Bb[4][6] = {{gx[b], 0, gy[b], 0, 0, 0}, {0, gy[b], gx[b], 0, 0, 0}, {0, 0, 0, 0, gy[b], gy[b]}, {0, 0, 0, -gx[b], 0, gx[b]}};
          const int dof[4] = {0, 1, 3, 4};
          T blk[6][6];
          for (int i = 0; i < 6; ++i)
            for (int j = 0; j < 6; ++j) blk[i][j] = 0;
          for (int ci = 0; ci < 4; ++ci)
            for (int cj = 0; cj < 4; ++cj) {
              T s = 0;
              for (int r = 0; r < 6; ++r) {
                T db = 0;
                for (int c = 0; c < 6; ++c) db += (T)sp.D[r * 6 + c] * Bb[cj][c];
                s += Ba[ci][r] * db;
              }
              blk[dof[ci]][dof[cj]] = s * wt;
            }
          for (int i = 0; i < 6; ++i)
            for (int j = 0; j < 6; ++j) {
              const size_t idx = (size_t)(6 * a + i) * ND + 6 * b + j;
              if (what == 5) Ko[idx * nq + q] = blk[i][j];
              else if (q == 0) Ko[idx] = blk[i][j];
              else Ko[idx] += blk[i][j];
            }
        }
      }
    }
    if (what == 7) shell_store_B<T, NEN>(out + e * 6 * ND, 1, Gx, Gy);
    if (what == 9) {
      // strain = B u_e (rows: exx, eyy, gxy from u,v; kx, ky, kxy from thx, thy), stress = D strain (shell.py:455-481, :863-879)
      T st[6] = {0, 0, 0, 0, 0, 0};
      for (int a = 0; a < NEN; ++a) {
        const long long n = ldidx(conn + e * NEN + a);
        const T ux = __ldg(disp + 6 * n), uy = __ldg(disp + 6 * n + 1), tx = __ldg(disp + 6 * n + 3), ty = __ldg(disp + 6 * n + 4);
        st[0] += Gx[a] * ux;
        st[1] += Gy[a] * uy;
        st[2] += Gy[a] * ux + Gx[a] * uy;
        st[3] += -Gx[a] * ty;
        st[4] += Gy[a] * tx;
        st[5] += Gy[a] * tx + Gx[a] * ty;
      }
      for (int r = 0; r < 6; ++r) {
        T sg = 0;
        for (int c = 0; c < 6; ++c) sg += st[c] * (T)sp.D[r * 6 + c];
        out[e * 6 + r] = sg;
      }
    }
  }
}

// fac = [nq,5] rows (1-xi, 1+xi, 1-eta, 1+eta, w) on the host
template <typename T, typename I>
static int shell_dispatch(int kind, int what, const void* coords, const void* conn, long long M, const double* fac, int nq, const double* D6,
                          const void* disp, void* out, cudaStream_t s) {
  FEMB_CHECK_ARG(what >= 0 && what <= 9 && what != 6, "femb_shell: what in 0..5, 7..9");
  FEMB_CHECK_ARG(kind == FEMB_S3 || kind == FEMB_S4, "femb_shell: kind must be S3/S4");
  if (M == 0) return FEMB_OK;
  ShellPts sp;
  memset(&sp, 0, sizeof(sp));
  if (kind == FEMB_S4) {
    FEMB_CHECK_ARG(fac != nullptr && nq >= 1 && nq <= 16, "S4 needs 1..16 points");
    for (int q = 0; q < nq; ++q) {
      for (int k = 0; k < 4; ++k) sp.f[q][k] = fac[5 * q + k];
      sp.w[q] = fac[5 * q + 4];
    }
  } else {
    nq = 1;
    sp.w[0] = 1.0;
  }
  if (what == 4 || what == 5 || what == 9) {
    FEMB_CHECK_ARG(D6 != nullptr, "D6_host");
    memcpy(sp.D, D6, sizeof(sp.D));
  }
  FEMB_CHECK_ARG(what != 9 || disp != nullptr, "what = 9 needs the displacement [N,6]");
  const int grid = grid_for(M, 64);
  const T* X = static_cast<const T*>(coords);
  const I* C = static_cast<const I*>(conn);
  const T* U = static_cast<const T*>(disp);
  if (kind == FEMB_S3) {
    const int w3 = what == 5 ? 4 : what == 8 ? 7 : what;  // one point: the per-point layouts coincide with the summed ones
    shell_kernel<T, I, 3><<<grid, 64, 0, s>>>(X, C, M, sp, nq, w3, U, (T*)out);
  } else {
    shell_kernel<T, I, 4><<<grid, 64, 0, s>>>(X, C, M, sp, nq, what, U, (T*)out);
  }
  FEMB_LAUNCH_CHECK();
  return FEMB_OK;
}

static int shell_entry(int kind, int what, const void* coords, int fp, const void* conn, int ib, int64_t M, const double* fac, int nq,
                       const double* D6_host, const void* disp, void* out, femb_stream stream) {
  FEMB_CHECK_ARG((fp == 4 || fp == 8) && (ib == 4 || ib == 8), "fp in {4,8}, ib in {4,8}");
  cudaStream_t s = as_stream(stream);
  if (fp == 8) {
    if (ib == 8) return shell_dispatch<double, long long>(kind, what, coords, conn, M, fac, nq, D6_host, disp, out, s);
    return shell_dispatch<double, int>(kind, what, coords, conn, M, fac, nq, D6_host, disp, out, s);
  }
  if (ib == 8) return shell_dispatch<float, long long>(kind, what, coords, conn, M, fac, nq, D6_host, disp, out, s);
  return shell_dispatch<float, int>(kind, what, coords, conn, M, fac, nq, D6_host, disp, out, s);
}

}  // namespace femb

using namespace femb;

extern "C" int femb_shell(int kind, int what, const void* coords, int fp, const void* conn, int ib, int64_t M, const double* pts_host, int nq,
                          const double* D6_host, void* out, femb_stream stream) {
  FEMB_CHECK_ARG(what >= 0 && what <= 5, "femb_shell: what in 0..5 (7..9 go through femb_shell_ex)");
  double fac[16 * 5];
  if (kind == FEMB_S4) {
    FEMB_CHECK_ARG(pts_host != nullptr && nq >= 1 && nq <= 16, "S4 needs 1..16 points");
    for (int q = 0; q < nq; ++q) {
      const double xi = pts_host[4 * q], eta = pts_host[4 * q + 1];
      fac[5 * q] = 1 - xi, fac[5 * q + 1] = 1 + xi, fac[5 * q + 2] = 1 - eta, fac[5 * q + 3] = 1 + eta, fac[5 * q + 4] = pts_host[4 * q + 3];
    }
  }
  return shell_entry(kind, what, coords, fp, conn, ib, M, fac, nq, D6_host, nullptr, out, stream);
}

extern "C" int femb_shell_ex(int kind, int what, const void* coords, int fp, const void* conn, int ib, int64_t M, const double* fac_host,
                             int nq, const double* D6_host, const void* disp, void* out, femb_stream stream) {
  return shell_entry(kind, what, coords, fp, conn, ib, M, fac_host, nq, D6_host, disp, out, stream);
}
