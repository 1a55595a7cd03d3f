// Global assembly: element matrices -> CSR, deterministic, no atomics on values.
//
// Replaces the un-coalesced torch.sparse_coo_tensor of subdivision.ipynb cell 6 (reference) with
//   plan   : node->element incidence lists (stable radix sort of (node, flat slot)) and the node-level sparsity pattern
//            (segmented sort + unique of each node's candidate columns)
//   values : every CSR row has ONE owner (a thread for P1 tets, a warp for the generic path) that adds the row's
//            contributions in incidence order (ascending element id), i.e. a sort-based segmented reduction into the
//            precomputed pattern: bit-reproducible, no atomics.  Either from materialised Ke (any element type / dofs
//            per node) or fused from coordinates for P1 tets (Poisson and elasticity; Ke never exists).
#include <cstdlib>
#include <cub/cub.cuh>

#include "common.cuh"

#include "plan.cuh"

namespace femb {

template <typename I>
__global__ void conn_to_i32(const I* __restrict__ conn, long long L, long long N, int* __restrict__ out, int* __restrict__ iota,
                            int* __restrict__ bad) {
  bool oob = false;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < L; t += (long long)gridDim.x * blockDim.x) {
    const long long v = ldidx(conn + t);
    oob |= v < 0 || v >= N;  // the reference raises an index error here; the sorts below would index out of bounds
    out[t] = (int)v;
    iota[t] = (int)t;
  }
  if (oob) *bad = 1;
}

// inc_ptr from the sorted node keys (handles nodes without elements)
__global__ void ptr_from_sorted(const int* __restrict__ keys, long long L, long long N, int* __restrict__ ptr) {
  for (long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x; k < L; k += (long long)gridDim.x * blockDim.x) {
    const int prev = k == 0 ? -1 : keys[k - 1], cur = keys[k];
    for (int n = prev + 1; n <= cur; ++n) ptr[n] = (int)k;
    if (k == L - 1)
      for (long long n = cur + 1; n <= N; ++n) ptr[n] = (int)L;
  }
}

__global__ void emit_candidates(const int* __restrict__ conn32, const int* __restrict__ inc, long long L, int nen, int* __restrict__ cand) {
  const long long total = L * nen;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const long long k = t / nen;
    const int b = (int)(t - k * nen);
    const int e = inc[k] / nen;
    cand[t] = conn32[(long long)e * nen + b];
  }
}

__global__ void scale_offsets(const int* __restrict__ ptr, long long n, int nen, int* __restrict__ off) {
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) off[t] = ptr[t] * nen;
}

// one warp per node: count (fill==nullptr) or write the distinct values of its sorted candidate segment
__global__ void unique_rows(const int* __restrict__ cand, const int* __restrict__ off, long long N, int* __restrict__ cnt,
                            const int* __restrict__ node_ptr, int* __restrict__ node_col) {
  const int lane = threadIdx.x & 31;
  const long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5, nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long i = warp; i < N; i += nwarps) {
    const int s = off[i], t = off[i + 1];
    int base = node_col ? node_ptr[i] : 0, total = 0;
    for (int k0 = s; k0 < t; k0 += 32) {
      const int k = k0 + lane;
      bool head = false;
      int v = 0;
      if (k < t) {
        v = cand[k];
        head = (k == s) || (cand[k - 1] != v);
      }
      const unsigned m = __ballot_sync(0xffffffffu, head);
      if (node_col && head) node_col[base + total + __popc(m & ((1u << lane) - 1))] = v;
      total += __popc(m);
    }
    if (!node_col && lane == 0) cnt[i] = total;
  }
}

__global__ void inc_counts(const int* __restrict__ inc_ptr, long long N, int* __restrict__ out) {
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < N; t += (long long)gridDim.x * blockDim.x) out[t] = inc_ptr[t + 1] - inc_ptr[t];
}

__global__ void pattern_kernel(const int* __restrict__ node_ptr, const int* __restrict__ node_col, long long N, int d, int* __restrict__ crow,
                               int* __restrict__ col) {
  const int lane = threadIdx.x & 31;
  const long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5, nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long i = warp; i < N; i += nwarps) {
    const int s = node_ptr[i], len = node_ptr[i + 1] - s;
    const long long base = (long long)d * d * s;
    if (lane < d) crow[i * d + lane] = (int)(base + (long long)lane * len * d);
    if (i == N - 1 && lane == 0) crow[N * d] = (int)((long long)d * d * node_ptr[N]);
    const int per_row = len * d;
    for (int t = lane; t < d * per_row; t += 32) {
      const int within = t % per_row;
      col[base + t] = node_col[s + within / d] * d + within % d;
    }
  }
}

__device__ __forceinline__ int lower_bound_i32(const int* a, int n, int v) {
  int lo = 0, hi = n;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (a[mid] < v) lo = mid + 1;
    else hi = mid;
  }
  return lo;
}

// slot table: for incidence k (node i = keys[k], element e, local a) and every local b, where conn[e][b] sits in row i
__global__ void fill_slots(const int* __restrict__ conn32, const int* __restrict__ inc, const int* __restrict__ keys, const int* __restrict__ node_ptr,
                           const int* __restrict__ node_col, long long L, int nen, unsigned char* __restrict__ slots) {
  const long long total = L * nen;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const long long k = t / nen;
    const int b = (int)(t - k * nen), i = keys[k], e = inc[k] / nen;
    const int s = node_ptr[i];
    slots[t] = (unsigned char)lower_bound_i32(node_col + s, node_ptr[i + 1] - s, conn32[(long long)e * nen + b]);
  }
}

// ---- thread-per-row P1 Poisson assembly ------------------------------------------------------------
// One THREAD owns one node row and walks its incidence list in ascending element order, so the sum order is fixed without
// any cross-lane merge.  SRC 0: the element's cofactor vectors are rebuilt from the coordinates (fused path, Ke never
// exists); SRC 1: row `a` of a materialised Ke[M,4,4] is read as one 32-byte sector.  The row accumulator lives in shared
// memory as acc[slot][thread] (bank-conflict free, no dynamic register indexing).  The dependent load chain
// inc[k] -> conn[e] -> coords[n] is software-pipelined three deep (index data two incidences ahead, connectivity one
// ahead, coordinates one ahead of the arithmetic), which is what turns the kernel from latency-bound into fp64-bound.
struct P1Idx {
  int slot;
  unsigned int sl;
};

template <int SRC>
__device__ __forceinline__ void p1_load_payload(bool on, const int4& q, int slot, const double* __restrict__ coords, const double* __restrict__ Ke,
                                                double* x) {
  if (!on) return;
  if (SRC == 1) {
    const double* row = Ke + ((long long)(slot >> 2) * 4 + (slot & 3)) * 4;
    asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(x[0]), "=d"(x[1]), "=d"(x[2]), "=d"(x[3]) : "l"(row));
  } else {
    const int nd4[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
    for (int n = 0; n < 4; ++n)
#pragma unroll
      for (int t = 0; t < 3; ++t) x[3 * n + t] = __ldg(coords + 3ll * nd4[n] + t);
  }
}

template <int SRC, int BD>
__device__ __forceinline__ void p1_accumulate(const double* x, int slot, unsigned int sl, double* acc, int tid, int* flag) {
  double v[4];
  if (SRC == 1) {
    v[0] = x[0], v[1] = x[1], v[2] = x[2], v[3] = x[3];
  } else {
    const int a = slot & 3;
    double e1[3], e2[3], e3[3], c[4][3];
#pragma unroll
    for (int t = 0; t < 3; ++t) {
      e1[t] = x[3 + t] - x[t];
      e2[t] = x[6 + t] - x[t];
      e3[t] = x[9 + t] - x[t];
    }
    c[1][0] = e2[1] * e3[2] - e2[2] * e3[1], c[1][1] = e2[2] * e3[0] - e2[0] * e3[2], c[1][2] = e2[0] * e3[1] - e2[1] * e3[0];
    c[2][0] = e3[1] * e1[2] - e3[2] * e1[1], c[2][1] = e3[2] * e1[0] - e3[0] * e1[2], c[2][2] = e3[0] * e1[1] - e3[1] * e1[0];
    c[3][0] = e1[1] * e2[2] - e1[2] * e2[1], c[3][1] = e1[2] * e2[0] - e1[0] * e2[2], c[3][2] = e1[0] * e2[1] - e1[1] * e2[0];
    const double det = e1[0] * c[1][0] + e1[1] * c[1][1] + e1[2] * c[1][2];
    if (fabs(det) < 1e-12 && flag) *flag = 1;
#pragma unroll
    for (int t = 0; t < 3; ++t) c[0][t] = -(c[1][t] + c[2][t] + c[3][t]);
    const double scale = 1.0 / (6.0 * fabs(det));  // V g_a.g_b = c_a.c_b / (6|det|)
    double ca[3] = {c[0][0], c[0][1], c[0][2]};
#pragma unroll
    for (int n = 1; n < 4; ++n)
      if (a == n) ca[0] = c[n][0], ca[1] = c[n][1], ca[2] = c[n][2];
#pragma unroll
    for (int b = 0; b < 4; ++b) v[b] = (ca[0] * c[b][0] + ca[1] * c[b][1] + ca[2] * c[b][2]) * scale;
  }
#pragma unroll
  for (int b = 0; b < 4; ++b) acc[((sl >> (8 * b)) & 255u) * BD + tid] += v[b];
}

template <int SRC, int BD>
__global__ void __launch_bounds__(BD) assemble_p1_scalar_rows(const int* __restrict__ conn32, const int* __restrict__ inc_ptr,
                                                              const int* __restrict__ inc, const unsigned int* __restrict__ slots4,
                                                              const int* __restrict__ node_ptr, long long N, const double* __restrict__ coords,
                                                              const double* __restrict__ Ke, double* __restrict__ vals, int* __restrict__ flag) {
  extern __shared__ __align__(16) double acc[];  // [max_row][BD]
  const int tid = threadIdx.x;
  constexpr int NX = SRC == 1 ? 4 : 12;
  for (long long i0 = (long long)blockIdx.x * BD; i0 < N; i0 += (long long)gridDim.x * BD) {
    const long long i = i0 + tid;
    if (i >= N) continue;
    const int s = node_ptr[i], len = node_ptr[i + 1] - s;
    const int k0 = inc_ptr[i], k1 = inc_ptr[i + 1];
    auto load_idx = [&](int k) {
      P1Idx r{0, 0u};
      if (k < k1) r.slot = __ldg(inc + k), r.sl = __ldg(slots4 + k);
      return r;
    };
    auto load_conn = [&](int k, const P1Idx& id) {
      int4 q = make_int4(0, 0, 0, 0);
      if (SRC == 0 && k < k1) q = __ldg(reinterpret_cast<const int4*>(conn32) + (id.slot >> 2));
      return q;
    };
    // prologue: fill the pipeline
    P1Idx ia = load_idx(k0), ib = load_idx(k0 + 1), ic = load_idx(k0 + 2);
    int4 qa = load_conn(k0, ia), qb = load_conn(k0 + 1, ib);
    double xa[NX], xb[NX];
    p1_load_payload<SRC>(k0 < k1, qa, ia.slot, coords, Ke, xa);
    for (int p = 0; p < len; ++p) acc[p * BD + tid] = 0.0;
    for (int k = k0; k < k1; k += 2) {
      // even half: payload k+1 and index data further ahead go out before the arithmetic of k
      p1_load_payload<SRC>(k + 1 < k1, qb, ib.slot, coords, Ke, xb);
      qa = load_conn(k + 2, ic);
      P1Idx id = load_idx(k + 3);
      p1_accumulate<SRC, BD>(xa, ia.slot, ia.sl, acc, tid, flag);
      if (k + 1 >= k1) break;
      // odd half
      p1_load_payload<SRC>(k + 2 < k1, qa, ic.slot, coords, Ke, xa);
      qb = load_conn(k + 3, id);
      P1Idx ie = load_idx(k + 4);
      p1_accumulate<SRC, BD>(xb, ib.slot, ib.sl, acc, tid, flag);
      ia = ic, ib = id, ic = ie;
    }
    double* dst = vals + s;
    for (int p = 0; p < len; ++p) dst[p] = acc[p * BD + tid];
  }
}

// ---- tiled fused P1 Poisson assembly ----------------------------------------------------------------------
// Same ownership (thread = row, ascending element order) but every per-lane gather of index data is replaced by ONE
// coalesced 16-byte record per step: the plan stores, for each tile of 32 consecutive rows, step-major records holding the
// three OTHER nodes of the incident element and their row slots.  The row's own node is taken as local node 0 (the
// element matrix is invariant under the renumbering), so only three padded 32-byte coordinate loads remain per step.
// L1 wavefronts per warp-step drop from ~176 (three strided index streams + 12 strided 8-byte coordinate loads) to ~28.
__global__ void tile_degrees(const int* __restrict__ inc_ptr, long long N, long long ntiles, int* __restrict__ deg) {
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t <= ntiles; t += (long long)gridDim.x * blockDim.x) {
    int m = 0;
    if (t < ntiles)
      for (long long i = t * 32; i < min(N, t * 32 + 32); ++i) m = max(m, inc_ptr[i + 1] - inc_ptr[i]);
    deg[t] = m;
  }
}

__global__ void build_records(const int* __restrict__ conn32, const int* __restrict__ inc_ptr, const int* __restrict__ inc,
                              const unsigned char* __restrict__ slots, const int* __restrict__ tile_ptr, long long N, long long ntiles,
                              int4* __restrict__ rec, unsigned char* __restrict__ pdiag) {
  const int lane = threadIdx.x & 31;
  const long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5, nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long t = warp; t < ntiles; t += nwarps) {
    const long long i = t * 32 + lane;
    const int s0 = tile_ptr[t], steps = tile_ptr[t + 1] - s0;
    const int k0 = i < N ? inc_ptr[i] : 0, k1 = i < N ? inc_ptr[i + 1] : 0;
    for (int st = 0; st < steps; ++st) {
      int4 r = make_int4(-1, -1, -1, 0);
      const int k = k0 + st;
      if (k < k1) {
        const int slot = inc[k], e = slot >> 2, a = slot & 3;
        int ids[3], sl[3], n = 0;
        for (int b = 0; b < 4; ++b) {
          if (b == a) {
            if (st == 0) pdiag[i] = slots[4ll * k + b];
            continue;
          }
          ids[n] = conn32[4ll * e + b];
          sl[n] = slots[4ll * k + b];
          ++n;
        }
        r = make_int4(ids[0], ids[1], ids[2], sl[0] | (sl[1] << 8) | (sl[2] << 16));
      }
      rec[((long long)s0 + st) * 32 + lane] = r;
    }
  }
}

__global__ void pad_coords(const double* __restrict__ coords, long long N, double* __restrict__ out) {
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < N; t += (long long)gridDim.x * blockDim.x)
    st256(out + 4 * t, coords[3 * t], coords[3 * t + 1], coords[3 * t + 2], 0.0);
}

__device__ __forceinline__ void ld_xyz(const double* __restrict__ c4, int node, double* x) {
  [[maybe_unused]] double w;  // fourth lane of the padded record
  asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(x[0]), "=d"(x[1]), "=d"(x[2]), "=d"(w) : "l"(c4 + 4ll * node));
}

template <int BD, int OCC, int VARIANT>
__global__ void __launch_bounds__(BD, OCC) assemble_p1_poisson_tiles(const int4* __restrict__ rec, const int* __restrict__ tile_ptr,
                                                                const int* __restrict__ node_ptr, const unsigned char* __restrict__ pdiag,
                                                                long long N, long long ntiles, const double* __restrict__ c4,
                                                                double* __restrict__ vals, int* __restrict__ flag) {
  extern __shared__ __align__(16) double acc[];  // [max_row][BD]
  const int tid = threadIdx.x, lane = tid & 31;
  constexpr int WPB = BD / 32;
  for (long long t = (long long)blockIdx.x * WPB + (tid >> 5); t < ntiles; t += (long long)gridDim.x * WPB) {
    const long long i = t * 32 + lane;
    const bool live = i < N;
    const int s = live ? node_ptr[i] : 0, len = live ? node_ptr[i + 1] - s : 0;
    const int s0 = tile_ptr[t], steps = tile_ptr[t + 1] - s0;
    const int4* rp = rec + (long long)s0 * 32 + lane;
    double x0[3] = {0, 0, 0};
    if (live) ld_xyz(c4, (int)i, x0);
    const int pd = live ? pdiag[i] : 0;
    for (int p = 0; p < len; ++p) acc[p * BD + tid] = 0.0;
    double diag = 0.0;
    // pipeline: records are resident five steps ahead, the coordinates of step+1 are loaded during the arithmetic of `step`,
    // and the coordinates of steps +3/+4 are pulled into L1 with prefetch hints (no registers) so that load mostly hits
    auto rec_at = [&](int st) { return st < steps ? __ldg(rp + st * 32) : make_int4(-1, 0, 0, 0); };
    auto pf = [&](const int4& r) {
      if (r.x < 0) return;
      asm volatile("prefetch.global.L1 [%0];" ::"l"(c4 + 4ll * r.x));
      asm volatile("prefetch.global.L1 [%0];" ::"l"(c4 + 4ll * r.y));
      asm volatile("prefetch.global.L1 [%0];" ::"l"(c4 + 4ll * r.z));
    };
    int4 ra = rec_at(0), rb = rec_at(1), r2 = rec_at(2), r3 = rec_at(3), r4 = rec_at(4);
    double xa[9], xb[9];
    if (ra.x >= 0) ld_xyz(c4, ra.x, xa), ld_xyz(c4, ra.y, xa + 3), ld_xyz(c4, ra.z, xa + 6);
    pf(rb), pf(r2);
    // register-only part of one incidence: the row of the element matrix that belongs to local node 0 (= the row node)
    auto row_of = [&](const int4& r, const double* x, double* v) {  // v[0] diagonal, v[1..3] the three other nodes
      double e1[3], e2[3], e3[3], c1[3], c2[3], c3[3], c0[3];
#pragma unroll
      for (int q = 0; q < 3; ++q) e1[q] = x[q] - x0[q], e2[q] = x[3 + q] - x0[q], e3[q] = x[6 + q] - x0[q];
      c1[0] = e2[1] * e3[2] - e2[2] * e3[1], c1[1] = e2[2] * e3[0] - e2[0] * e3[2], c1[2] = e2[0] * e3[1] - e2[1] * e3[0];
      c2[0] = e3[1] * e1[2] - e3[2] * e1[1], c2[1] = e3[2] * e1[0] - e3[0] * e1[2], c2[2] = e3[0] * e1[1] - e3[1] * e1[0];
      c3[0] = e1[1] * e2[2] - e1[2] * e2[1], c3[1] = e1[2] * e2[0] - e1[0] * e2[2], c3[2] = e1[0] * e2[1] - e1[1] * e2[0];
      const double det = e1[0] * c1[0] + e1[1] * c1[1] + e1[2] * c1[2];
      if (r.x >= 0 && fabs(det) < 1e-12 && flag) *flag = 1;
#pragma unroll
      for (int q = 0; q < 3; ++q) c0[q] = -(c1[q] + c2[q] + c3[q]);
      // V g_a.g_b = c_a.c_b / (6|det|); reciprocal by MUFU seed + two Newton steps (branch-free, ~1 ulp)
      const double d6 = 6.0 * fabs(det);
      double rc;
      if (VARIANT & 2) {
        asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(rc) : "d"(d6));
        rc = rc * (2.0 - d6 * rc);
        rc = rc * (2.0 - d6 * rc);
      } else {
        rc = 1.0 / d6;
      }
      const double scale = r.x >= 0 ? rc : 0.0;  // padding steps contribute exactly zero
      v[0] = (c0[0] * c0[0] + c0[1] * c0[1] + c0[2] * c0[2]) * scale;
      v[1] = (c0[0] * c1[0] + c0[1] * c1[1] + c0[2] * c1[2]) * scale;
      v[2] = (c0[0] * c2[0] + c0[1] * c2[1] + c0[2] * c2[2]) * scale;
      v[3] = (c0[0] * c3[0] + c0[1] * c3[1] + c0[2] * c3[2]) * scale;
    };
    auto scatter = [&](const int4& r, const double* v) {
      if (r.x < 0) return;
      diag += v[0];
      acc[(r.w & 255) * BD + tid] += v[1];
      acc[((r.w >> 8) & 255) * BD + tid] += v[2];
      acc[((r.w >> 16) & 255) * BD + tid] += v[3];
    };
    for (int st = 0; st < steps; st += 2) {
      if (rb.x >= 0) ld_xyz(c4, rb.x, xb), ld_xyz(c4, rb.y, xb + 3), ld_xyz(c4, rb.z, xb + 6);
      const int4 r5 = rec_at(st + 5), r6 = rec_at(st + 6);
      pf(r3), pf(r4);
      // the arithmetic of two consecutive incidences is independent: issuing both before the ordered accumulation gives the
      // fp64 pipe two dependency chains to interleave (the sum order into acc[] is still step order)
      double va[4], vb[4];
      if (VARIANT & 1) {
        row_of(ra, xa, va);
        row_of(rb, xb, vb);
        scatter(ra, va);
        scatter(rb, vb);
      } else {
        row_of(ra, xa, va);
        scatter(ra, va);
      }
      if (r2.x >= 0) ld_xyz(c4, r2.x, xa), ld_xyz(c4, r2.y, xa + 3), ld_xyz(c4, r2.z, xa + 6);
      if (!(VARIANT & 1)) {
        row_of(rb, xb, vb);
        scatter(rb, vb);
      }
      ra = r2, rb = r3, r2 = r4, r3 = r5, r4 = r6;
    }
    if (live) {
      if (len > 0) acc[pd * BD + tid] += diag;
      double* dst = vals + s;
      for (int p = 0; p < len; ++p) dst[p] = acc[p * BD + tid];
    }
  }
}

// ---- tiled fused P1 elasticity assembly ---------------------------------------------------------------------------------
// Same records and ownership as the Poisson kernel; three warps share a tile of 32 nodes, warp alpha owning dof row
// (node, alpha).  Each thread rebuilds the cofactor vectors (3x redundant fp64 work, but the output is 9x larger than for
// Poisson, so the kernel stays within ~4x of its write roofline) and accumulates its row's 3*len entries in shared memory
// laid out like the CSR row: acc[(slot*3+beta)][thread].
__global__ void __launch_bounds__(96) assemble_p1_elasticity_tiles(const int4* __restrict__ rec, const int* __restrict__ tile_ptr,
                                                                   const int* __restrict__ node_ptr, const unsigned char* __restrict__ pdiag,
                                                                   long long N, long long ntiles, const double* __restrict__ c4, double lam,
                                                                   double mu, double* __restrict__ vals, int* __restrict__ flag) {
  constexpr int BD = 96;
  extern __shared__ __align__(16) double acc[];  // [max_row*3][BD]
  const int tid = threadIdx.x, lane = tid & 31, alpha = tid >> 5;
  for (long long t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const long long i = t * 32 + lane;
    const bool live = i < N;
    const int s = live ? node_ptr[i] : 0, len = live ? node_ptr[i + 1] - s : 0;
    const int s0 = tile_ptr[t], steps = tile_ptr[t + 1] - s0;
    const int4* rp = rec + (long long)s0 * 32 + lane;
    double x0[3] = {0, 0, 0};
    if (live) ld_xyz(c4, (int)i, x0);
    const int pd = live ? pdiag[i] : 0;
    for (int p = 0; p < 3 * len; ++p) acc[p * BD + tid] = 0.0;
    auto rec_at = [&](int st) { return st < steps ? __ldg(rp + st * 32) : make_int4(-1, 0, 0, 0); };
    int4 ra = rec_at(0), rb = rec_at(1), r2 = rec_at(2);
    double xa[9], xb[9];
    if (ra.x >= 0) ld_xyz(c4, ra.x, xa), ld_xyz(c4, ra.y, xa + 3), ld_xyz(c4, ra.z, xa + 6);
    auto contribute = [&](const int4& r, const double* x) {
      if (r.x < 0) return;
      double e1[3], e2[3], e3[3], c[4][3];
#pragma unroll
      for (int q = 0; q < 3; ++q) e1[q] = x[q] - x0[q], e2[q] = x[3 + q] - x0[q], e3[q] = x[6 + q] - x0[q];
      c[1][0] = e2[1] * e3[2] - e2[2] * e3[1], c[1][1] = e2[2] * e3[0] - e2[0] * e3[2], c[1][2] = e2[0] * e3[1] - e2[1] * e3[0];
      c[2][0] = e3[1] * e1[2] - e3[2] * e1[1], c[2][1] = e3[2] * e1[0] - e3[0] * e1[2], c[2][2] = e3[0] * e1[1] - e3[1] * e1[0];
      c[3][0] = e1[1] * e2[2] - e1[2] * e2[1], c[3][1] = e1[2] * e2[0] - e1[0] * e2[2], c[3][2] = e1[0] * e2[1] - e1[1] * e2[0];
      const double det = e1[0] * c[1][0] + e1[1] * c[1][1] + e1[2] * c[1][2];
      if (fabs(det) < 1e-12 && flag) *flag = 1;
#pragma unroll
      for (int q = 0; q < 3; ++q) c[0][q] = -(c[1][q] + c[2][q] + c[3][q]);
      const double scale = 1.0 / (6.0 * fabs(det));  // V g_a (x) g_b = c_a (x) c_b / (6|det|)
      const double ca = alpha == 0 ? c[0][0] : alpha == 1 ? c[0][1] : c[0][2];  // c0[alpha]
      const int slot[4] = {pd, r.w & 255, (r.w >> 8) & 255, (r.w >> 16) & 255};
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const double dot = mu * (c[0][0] * c[b][0] + c[0][1] * c[b][1] + c[0][2] * c[b][2]);
        const double cba = alpha == 0 ? c[b][0] : alpha == 1 ? c[b][1] : c[b][2];  // cb[alpha]
#pragma unroll
        for (int beta = 0; beta < 3; ++beta) {
          double v = lam * ca * c[b][beta] + mu * cba * c[0][beta];
          if (beta == alpha) v += dot;
          acc[(slot[b] * 3 + beta) * BD + tid] += v * scale;
        }
      }
    };
    for (int st = 0; st < steps; st += 2) {
      if (rb.x >= 0) ld_xyz(c4, rb.x, xb), ld_xyz(c4, rb.y, xb + 3), ld_xyz(c4, rb.z, xb + 6);
      const int4 r3 = rec_at(st + 3);
      contribute(ra, xa);
      if (r2.x >= 0) ld_xyz(c4, r2.x, xa), ld_xyz(c4, r2.y, xa + 3), ld_xyz(c4, r2.z, xa + 6);
      const int4 r4 = rec_at(st + 4);
      contribute(rb, xb);
      ra = r2, rb = r3, r2 = r4;
    }
    if (live) {
      double* dst = vals + 9ll * s + (long long)alpha * 3 * len;  // row (node i, alpha) of the dof-level CSR
      for (int p = 0; p < 3 * len; ++p) dst[p] = acc[p * BD + tid];
    }
  }
}

// ---- values from materialised Ke -----------------------------------------------------------------
// warp per node; shared accumulator laid out exactly like the node's d rows of the CSR value array
template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32) assemble_gather(const int* __restrict__ conn32, const int* __restrict__ inc_ptr,
                                                              const int* __restrict__ inc, const int* __restrict__ node_ptr,
                                                              const int* __restrict__ node_col, long long N, int nen, int d, int max_row,
                                                              const double* __restrict__ Ke, double* __restrict__ vals) {
  extern __shared__ __align__(16) double sm[];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  double* acc = sm + (size_t)w * max_row * d * d;
  const int nd = nen * d;
  const long long warp = (long long)blockIdx.x * WARPS + w, nwarps = (long long)gridDim.x * WARPS;
  for (long long i = warp; i < N; i += nwarps) {
    const int s = node_ptr[i], len = node_ptr[i + 1] - s;
    const int nacc = len * d * d;
    for (int t = lane; t < nacc; t += 32) acc[t] = 0.0;
    __syncwarp();
    const int* cols = node_col + s;
    for (int k = inc_ptr[i]; k < inc_ptr[i + 1]; ++k) {
      const int slot = inc[k], e = slot / nen, a = slot - e * nen;
      int pos = 0;
      if (lane < nen) pos = lower_bound_i32(cols, len, conn32[(long long)e * nen + lane]);
      const double* src = Ke + ((long long)e * nd + (long long)a * d) * nd;  // d consecutive rows of Ke
      for (int t0 = 0; t0 < d * nd; t0 += 32) {  // warp-uniform trip count: every lane takes part in the shuffle
        const int t = t0 + lane;
        const bool valid = t < d * nd;
        const int tt = valid ? t : 0;
        const int alpha = tt / nd, c = tt - alpha * nd, b = c / d, beta = c - b * d;
        const int p = __shfl_sync(0xffffffffu, pos, b);
        if (valid) acc[(alpha * len + p) * d + beta] += __ldg(src + t);
      }
      __syncwarp();
    }
    double* dst = vals + (long long)d * d * s;
    for (int t = lane; t < nacc; t += 32) dst[t] = acc[t];
    __syncwarp();
  }
}

// Same ownership and summation order, batched: ncu on the 2 M-tet P2 operator showed the kernel above waiting ~5 us per
// incidence on one dependent chain (inc -> conn -> binary search over the row's columns -> Ke), 23 % of the HBM peak.  Here the
// positions come from the plan's slot table (one coalesced byte load per batch instead of a search), the per-lane index
// arithmetic (alpha, b, beta of entry t) is hoisted out of the loops, and the Ke rows of U consecutive incidences are all in
// flight before the first is accumulated.  The adds into acc[] still happen incidence by incidence in ascending element
// order, so the values are bit-identical to the kernel above.  R = ceil(d*nd/32) entries per lane and incidence.
template <int WARPS, int R, int U>
__global__ void __launch_bounds__(WARPS * 32) assemble_gather_batched(const int* __restrict__ inc_ptr, const int* __restrict__ inc,
                                                                      const unsigned char* __restrict__ slots, const int* __restrict__ node_ptr,
                                                                      long long N, int nen, int d, int max_row, const double* __restrict__ Ke,
                                                                      double* __restrict__ vals) {
  extern __shared__ __align__(16) double sm[];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int nd = nen * d, per = d * nd;  // entries of the d consecutive Ke rows of one incidence
  double* acc = sm + (size_t)w * max_row * d * d;
  unsigned char* spos = reinterpret_cast<unsigned char*>(sm + (size_t)WARPS * max_row * d * d) + (size_t)w * 256;  // [U][nen] <= 256
  int bq[R], off[R];  // entry t = q*32+lane of an incidence: local column node b, and alpha*len*d + beta filled per row
  int al[R], be[R];
  bool valid[R];
#pragma unroll
  for (int q = 0; q < R; ++q) {
    const int t = q * 32 + lane;
    valid[q] = t < per;
    const int tt = valid[q] ? t : 0;
    al[q] = tt / nd;
    const int c = tt - al[q] * nd;
    bq[q] = c / d;
    be[q] = c - bq[q] * d;
  }
  const long long warp = (long long)blockIdx.x * WARPS + w, nwarps = (long long)gridDim.x * WARPS;
  for (long long i = warp; i < N; i += nwarps) {
    const int s = node_ptr[i], len = node_ptr[i + 1] - s;
    const int nacc = len * d * d;
    for (int t = lane; t < nacc; t += 32) acc[t] = 0.0;
#pragma unroll
    for (int q = 0; q < R; ++q) off[q] = al[q] * len * d + be[q];
    const int k0 = inc_ptr[i], k1 = inc_ptr[i + 1];
    for (int kb = k0; kb < k1; kb += U) {
      const int nb = min(U, k1 - kb);
      __syncwarp();  // acc zeroing / the previous batch's use of spos
      const int myslot = lane < nb ? __ldg(inc + kb + lane) : 0;
      for (int t = lane; t < nb * nen; t += 32) spos[t] = __ldg(slots + (long long)kb * nen + t);
      double v[U][R];
#pragma unroll
      for (int j = 0; j < U; ++j) {
        const int slot = __shfl_sync(0xffffffffu, myslot, j);
        const int e = slot / nen, a = slot - e * nen;
        const double* src = Ke + ((long long)e * nd + (long long)a * d) * nd;
#pragma unroll
        for (int q = 0; q < R; ++q) v[j][q] = (j < nb && valid[q]) ? __ldg(src + q * 32 + lane) : 0.0;
      }
      __syncwarp();  // spos visible
#pragma unroll
      for (int j = 0; j < U; ++j) {
        if (j < nb) {
#pragma unroll
          for (int q = 0; q < R; ++q)
            if (valid[q]) acc[off[q] + (int)spos[j * nen + bq[q]] * d] += v[j][q];
        }
        __syncwarp();  // the next incidence may touch the same entries from other lanes
      }
    }
    __syncwarp();
    double* dst = vals + (long long)d * d * s;
    for (int t = lane; t < nacc; t += 32) dst[t] = acc[t];
    __syncwarp();
  }
}

// ---- fused P1 tet assembly -------------------------------------------------------------------------
// warp per node, one lane per incident element (batches of 32).  Each lane rebuilds the element's cofactor vectors from
// the coordinates and produces row `a` of Ke; off-diagonal columns are merged in lane (= element) order through
// match/rank rounds, the diagonal through a fixed shuffle tree.  KIND 0: Poisson (d=1), 1: elasticity (d=3).
template <int KIND, int WARPS>
__global__ void __launch_bounds__(WARPS * 32) assemble_c3d4_fused(const int* __restrict__ conn32, const int* __restrict__ inc_ptr,
                                                                  const int* __restrict__ inc, const int* __restrict__ node_ptr,
                                                                  const int* __restrict__ node_col, long long N, int max_row,
                                                                  const double* __restrict__ coords, double lam, double mu,
                                                                  double* __restrict__ vals, int* __restrict__ flag) {
  constexpr int D = KIND == 0 ? 1 : 3, DD = D * D;
  extern __shared__ __align__(16) double sm[];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  double* acc = sm + (size_t)w * (max_row * DD + (max_row + 1) / 2);
  int* cols = reinterpret_cast<int*>(acc + (size_t)max_row * DD);
  const long long warp = (long long)blockIdx.x * WARPS + w, nwarps = (long long)gridDim.x * WARPS;
  for (long long i = warp; i < N; i += nwarps) {
    const int s = node_ptr[i], len = node_ptr[i + 1] - s;
    const int nacc = len * DD;
    for (int t = lane; t < nacc; t += 32) acc[t] = 0.0;
    for (int t = lane; t < len; t += 32) cols[t] = node_col[s + t];
    __syncwarp();
    const int pdiag = lower_bound_i32(cols, len, (int)i);
    const int k0 = inc_ptr[i], k1 = inc_ptr[i + 1];
    for (int kb = k0; kb < k1; kb += 32) {
      const int k = kb + lane;
      const bool on = k < k1;
      int nd4[4] = {0, 0, 0, 0}, a = 0;
      double c[4][3], scale = 0.0;
      if (on) {
        const int slot = inc[k], e = slot >> 2;
        a = slot & 3;
        const int4 q = __ldg(reinterpret_cast<const int4*>(conn32) + e);
        nd4[0] = q.x, nd4[1] = q.y, nd4[2] = q.z, nd4[3] = q.w;
        double x[4][3];
#pragma unroll
        for (int v = 0; v < 4; ++v)
#pragma unroll
          for (int t = 0; t < 3; ++t) x[v][t] = __ldg(coords + 3ll * nd4[v] + t);
        double e1[3], e2[3], e3[3];
#pragma unroll
        for (int t = 0; t < 3; ++t) {
          e1[t] = x[1][t] - x[0][t];
          e2[t] = x[2][t] - x[0][t];
          e3[t] = x[3][t] - x[0][t];
        }
        c[1][0] = e2[1] * e3[2] - e2[2] * e3[1], c[1][1] = e2[2] * e3[0] - e2[0] * e3[2], c[1][2] = e2[0] * e3[1] - e2[1] * e3[0];
        c[2][0] = e3[1] * e1[2] - e3[2] * e1[1], c[2][1] = e3[2] * e1[0] - e3[0] * e1[2], c[2][2] = e3[0] * e1[1] - e3[1] * e1[0];
        c[3][0] = e1[1] * e2[2] - e1[2] * e2[1], c[3][1] = e1[2] * e2[0] - e1[0] * e2[2], c[3][2] = e1[0] * e2[1] - e1[1] * e2[0];
        const double det = e1[0] * c[1][0] + e1[1] * c[1][1] + e1[2] * c[1][2];
        if (fabs(det) < 1e-12 && flag) *flag = 1;
#pragma unroll
        for (int t = 0; t < 3; ++t) c[0][t] = -(c[1][t] + c[2][t] + c[3][t]);
        scale = 1.0 / (6.0 * fabs(det));  // V g_a.g_b = c_a.c_b / (6 |det|)
      }
      // gradient-cofactor of this lane's own row node
      double ca[3] = {0, 0, 0};
#pragma unroll
      for (int v = 0; v < 4; ++v)
        if (v == a) ca[0] = c[v][0], ca[1] = c[v][1], ca[2] = c[v][2];
      // diagonal block (row node with itself): fixed shuffle tree over the batch
      {
        const double dot = ca[0] * ca[0] + ca[1] * ca[1] + ca[2] * ca[2];
#pragma unroll
        for (int r = 0; r < DD; ++r) {
          double v;
          if (KIND == 0) {
            v = dot;
          } else {
            v = (lam + mu) * ca[r / D] * ca[r % D];
            if (r / D == r % D) v += mu * dot;
          }
          v = warp_sum(on ? v * scale : 0.0);
          if (lane == 0) acc[((r / D) * len + pdiag) * D + (r % D)] += v;
        }
      }
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        double blk[DD];
        const bool off = on && (b != a);
        if (off) {
          const double dot = ca[0] * c[b][0] + ca[1] * c[b][1] + ca[2] * c[b][2];
          if (KIND == 0) {
            blk[0] = dot * scale;
          } else {
#pragma unroll
            for (int r = 0; r < 3; ++r)
#pragma unroll
              for (int t = 0; t < 3; ++t) {
                double v = lam * ca[r] * c[b][t] + mu * ca[t] * c[b][r];
                if (r == t) v += mu * dot;
                blk[r * D + t] = v * scale;
              }
          }
        } else {
#pragma unroll
          for (int r = 0; r < DD; ++r) blk[r] = 0.0;
        }
        // off-diagonal columns: equal columns are merged in lane order
        const int p = off ? lower_bound_i32(cols, len, nd4[b]) : -1 - lane;
        const unsigned peers = __match_any_sync(0xffffffffu, p);
        const int rank = __popc(peers & ((1u << lane) - 1));
        const int rounds = __reduce_max_sync(0xffffffffu, off ? __popc(peers) : 0);
        for (int t = 0; t < rounds; ++t) {
          if (off && rank == t) {
#pragma unroll
            for (int r = 0; r < DD; ++r) acc[((r / D) * len + p) * D + (r % D)] += blk[r];
          }
          __syncwarp();
        }
      }
    }
    __syncwarp();
    double* dst = vals + (long long)DD * s;
    for (int t = lane; t < nacc; t += 32) dst[t] = acc[t];
    __syncwarp();
  }
}

// ---- element-by-element operator (compute_nodal_forces / compute_shell_nodal_forces) -------------------
template <typename T>
__global__ void ebe_apply_kernel(const int* __restrict__ conn32, const int* __restrict__ inc_ptr, const int* __restrict__ inc, long long N,
                                 int nen, int d, const T* __restrict__ Ke, const T* __restrict__ u, T* __restrict__ y) {
  const int nd = nen * d;
  const long long rows = N * d;
  for (long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x; r < rows; r += (long long)gridDim.x * blockDim.x) {
    const long long i = r / d;
    const int alpha = (int)(r - i * d);
    T sum = 0;
    for (int k = inc_ptr[i]; k < inc_ptr[i + 1]; ++k) {
      const int slot = inc[k], e = slot / nen, a = slot - e * nen;
      const T* row = Ke + ((long long)e * nd + a * d + alpha) * nd;
      const int* cn = conn32 + (long long)e * nen;
      T part = 0;
      for (int b = 0; b < nen; ++b) {
        const T* ub = u + (long long)cn[b] * d;
        for (int beta = 0; beta < d; ++beta) part += __ldg(row + b * d + beta) * __ldg(ub + beta);
      }
      sum += part;
    }
    y[r] = sum;
  }
}

// shells: rotate each element's nodal vectors into its frame, apply Ke, rotate the row-node's force back (shell.py:58-102)
template <typename T>
__global__ void ebe_shell_kernel(const int* __restrict__ conn32, const int* __restrict__ inc_ptr, const int* __restrict__ inc, long long N,
                                 int nen, const T* __restrict__ Ke, const T* __restrict__ u, const T* __restrict__ unit, T* __restrict__ y) {
  const int nd = nen * 6;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < N; i += (long long)gridDim.x * blockDim.x) {
    T out[6] = {0, 0, 0, 0, 0, 0};
    for (int k = inc_ptr[i]; k < inc_ptr[i + 1]; ++k) {
      const int slot = inc[k], e = slot / nen, a = slot - e * nen;
      const T* R = unit + (long long)e * 9;
      T Rm[9];
      for (int t = 0; t < 9; ++t) Rm[t] = __ldg(R + t);
      T fl[6] = {0, 0, 0, 0, 0, 0};
      for (int b = 0; b < nen; ++b) {
        const T* ub = u + (long long)conn32[(long long)e * nen + b] * 6;
        T loc[6];
        for (int h = 0; h < 2; ++h)
          for (int dd = 0; dd < 3; ++dd)
            loc[3 * h + dd] = __ldg(ub + 3 * h) * Rm[dd * 3] + __ldg(ub + 3 * h + 1) * Rm[dd * 3 + 1] + __ldg(ub + 3 * h + 2) * Rm[dd * 3 + 2];
        for (int r = 0; r < 6; ++r) {
          const T* row = Ke + ((long long)e * nd + a * 6 + r) * nd + b * 6;
          T s = 0;
          for (int c = 0; c < 6; ++c) s += __ldg(row + c) * loc[c];
          fl[r] += s;
        }
      }
      for (int h = 0; h < 2; ++h)
        for (int g = 0; g < 3; ++g) out[3 * h + g] += Rm[g] * fl[3 * h] + Rm[3 + g] * fl[3 * h + 1] + Rm[6 + g] * fl[3 * h + 2];
    }
    for (int r = 0; r < 6; ++r) y[i * 6 + r] = out[r];
  }
}

static void plan_free(femb_csr_plan* p) {
  if (!p) return;
  cudaFree(p->conn32);
  cudaFree(p->inc_ptr);
  cudaFree(p->inc);
  cudaFree(p->node_ptr);
  cudaFree(p->node_col);
  cudaFree(p->inc_slots);
  cudaFree(p->rec);
  cudaFree(p->tile_ptr);
  cudaFree(p->pdiag);
  femb::block_plan_free(p->blk);
  delete p;
}

static int build_p1_records(femb_csr_plan* p, cudaStream_t s) {
  const long long N = p->N, nt = (N + 31) / 32;
  p->ntiles = nt;
  Scratch scr(s);
  int* deg;
  FEMB_CUDA(scr.alloc(&deg, nt + 1));
  FEMB_CUDA(cudaMalloc(&p->tile_ptr, sizeof(int) * (nt + 1)));
  FEMB_CUDA(cudaMalloc(&p->pdiag, (size_t)N));
  FEMB_CUDA(cudaMemsetAsync(p->pdiag, 0, (size_t)N, s));
  tile_degrees<<<grid_for(nt + 1, 256), 256, 0, s>>>(p->inc_ptr, N, nt, deg);
  FEMB_LAUNCH_CHECK();
  size_t tb = 0;
  FEMB_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tb, deg, p->tile_ptr, (int)(nt + 1), s));
  void* tmp;
  FEMB_CUDA(scr.alloc((char**)&tmp, tb));
  FEMB_CUDA(cub::DeviceScan::ExclusiveSum(tmp, tb, deg, p->tile_ptr, (int)(nt + 1), s));
  int total = 0;
  FEMB_CUDA(cudaMemcpyAsync(&total, p->tile_ptr + nt, sizeof(int), cudaMemcpyDeviceToHost, s));
  FEMB_CUDA(cudaStreamSynchronize(s));
  p->total_steps = total;
  FEMB_CUDA(cudaMalloc(&p->rec, sizeof(int4) * 32 * (size_t)std::max(total, 1)));
  build_records<<<grid_for(nt * 32, 256), 256, 0, s>>>(p->conn32, p->inc_ptr, p->inc, p->inc_slots, p->tile_ptr, N, nt, p->rec, p->pdiag);
  FEMB_LAUNCH_CHECK();
  return FEMB_OK;
}

template <typename I>
static int plan_build(const I* conn, femb_csr_plan* p, cudaStream_t s) {
  const long long M = p->M, N = p->N, L = M * p->nen;
  const int nen = p->nen;
  FEMB_CUDA(cudaMalloc(&p->conn32, sizeof(int) * std::max<long long>(L, 1)));
  FEMB_CUDA(cudaMalloc(&p->inc_ptr, sizeof(int) * (N + 1)));
  FEMB_CUDA(cudaMalloc(&p->inc, sizeof(int) * std::max<long long>(L, 1)));
  FEMB_CUDA(cudaMalloc(&p->node_ptr, sizeof(int) * (N + 1)));
  Scratch scr(s);
  int *iota, *keys_sorted, *cand, *cand_sorted, *off, *cnt;
  FEMB_CUDA(scr.alloc(&iota, L));
  FEMB_CUDA(scr.alloc(&keys_sorted, L));
  int* bad;
  FEMB_CUDA(scr.alloc(&bad, 1));
  FEMB_CUDA(cudaMemsetAsync(bad, 0, sizeof(int), s));
  conn_to_i32<I><<<grid_for(L, 256), 256, 0, s>>>(conn, L, N, p->conn32, iota, bad);
  FEMB_LAUNCH_CHECK();
  int hbad = 0;
  FEMB_CUDA(cudaMemcpyAsync(&hbad, bad, sizeof(int), cudaMemcpyDeviceToHost, s));
  FEMB_CUDA(cudaStreamSynchronize(s));
  FEMB_CHECK_ARG(hbad == 0, "connectivity holds a node index outside [0, n_nodes)");
  int bits = 1;
  while ((1ll << bits) < N) ++bits;
  size_t tb = 0;
  FEMB_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tb, p->conn32, keys_sorted, iota, p->inc, (int)L, 0, bits, s));
  void* tmp;
  FEMB_CUDA(scr.alloc((char**)&tmp, tb));
  FEMB_CUDA(cub::DeviceRadixSort::SortPairs(tmp, tb, p->conn32, keys_sorted, iota, p->inc, (int)L, 0, bits, s));
  if (L > 0) {
    ptr_from_sorted<<<grid_for(L, 256), 256, 0, s>>>(keys_sorted, L, N, p->inc_ptr);
  } else {
    FEMB_CUDA(cudaMemsetAsync(p->inc_ptr, 0, sizeof(int) * (N + 1), s));
  }
  FEMB_LAUNCH_CHECK();
  // candidate columns of every node, grouped by node; sorted per segment, then made unique
  const long long LC = L * nen;
  FEMB_CUDA(scr.alloc(&cand, LC));
  FEMB_CUDA(scr.alloc(&cand_sorted, LC));
  FEMB_CUDA(scr.alloc(&off, N + 1));
  FEMB_CUDA(scr.alloc(&cnt, N + 1));
  if (LC > 0) emit_candidates<<<grid_for(LC, 256), 256, 0, s>>>(p->conn32, p->inc, L, nen, cand);
  scale_offsets<<<grid_for(N + 1, 256), 256, 0, s>>>(p->inc_ptr, N + 1, nen, off);
  FEMB_LAUNCH_CHECK();
  size_t tb2 = 0;
  FEMB_CUDA(cub::DeviceSegmentedSort::SortKeys(nullptr, tb2, cand, cand_sorted, (int)LC, (int)N, off, off + 1, s));
  void* tmp2;
  FEMB_CUDA(scr.alloc((char**)&tmp2, tb2));
  FEMB_CUDA(cub::DeviceSegmentedSort::SortKeys(tmp2, tb2, cand, cand_sorted, (int)LC, (int)N, off, off + 1, s));
  FEMB_CUDA(cudaMemsetAsync(cnt, 0, sizeof(int) * (N + 1), s));
  unique_rows<<<grid_for(N * 32, 256), 256, 0, s>>>(cand_sorted, off, N, cnt, nullptr, nullptr);
  FEMB_LAUNCH_CHECK();
  size_t tb3 = 0;
  FEMB_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tb3, cnt, p->node_ptr, (int)(N + 1), s));
  void* tmp3;
  FEMB_CUDA(scr.alloc((char**)&tmp3, tb3));
  FEMB_CUDA(cub::DeviceScan::ExclusiveSum(tmp3, tb3, cnt, p->node_ptr, (int)(N + 1), s));
  // maxima for shared-memory sizing
  int *dmax, *inccnt;
  FEMB_CUDA(scr.alloc(&dmax, 2));
  FEMB_CUDA(scr.alloc(&inccnt, N));
  inc_counts<<<grid_for(N, 256), 256, 0, s>>>(p->inc_ptr, N, inccnt);
  size_t tb4 = 0, tb5 = 0;
  FEMB_CUDA(cub::DeviceReduce::Max(nullptr, tb4, cnt, dmax, (int)N, s));
  FEMB_CUDA(cub::DeviceReduce::Max(nullptr, tb5, inccnt, dmax + 1, (int)N, s));
  void* tmp4;
  FEMB_CUDA(scr.alloc((char**)&tmp4, std::max(tb4, tb5)));
  FEMB_CUDA(cub::DeviceReduce::Max(tmp4, tb4, cnt, dmax, (int)N, s));
  FEMB_CUDA(cub::DeviceReduce::Max(tmp4, tb5, inccnt, dmax + 1, (int)N, s));
  int hmax[2] = {0, 0}, hn = 0;
  FEMB_CUDA(cudaMemcpyAsync(hmax, dmax, sizeof(hmax), cudaMemcpyDeviceToHost, s));
  FEMB_CUDA(cudaMemcpyAsync(&hn, p->node_ptr + N, sizeof(int), cudaMemcpyDeviceToHost, s));
  FEMB_CUDA(cudaStreamSynchronize(s));
  p->max_row = hmax[0];
  p->max_inc = hmax[1];
  p->nnzn = hn;
  FEMB_CUDA(cudaMalloc(&p->node_col, sizeof(int) * std::max<long long>(p->nnzn, 1)));
  unique_rows<<<grid_for(N * 32, 256), 256, 0, s>>>(cand_sorted, off, N, nullptr, p->node_ptr, p->node_col);
  FEMB_LAUNCH_CHECK();
  if (p->max_row <= 255 && LC > 0) {
    FEMB_CUDA(cudaMalloc(&p->inc_slots, (size_t)LC));
    fill_slots<<<grid_for(LC, 256), 256, 0, s>>>(p->conn32, p->inc, keys_sorted, p->node_ptr, p->node_col, L, nen, p->inc_slots);
    FEMB_LAUNCH_CHECK();
  }
  FEMB_CUDA(cudaStreamSynchronize(s));
  return FEMB_OK;
}

}  // namespace femb

using namespace femb;

extern "C" int femb_csr_plan_create(const void* conn, int ib, int64_t M, int nen, int64_t n_nodes, femb_stream stream, femb_csr_plan** plan,
                                    int64_t* nnz_nodes) {
  FEMB_CHECK_ARG(ib == 4 || ib == 8, "ib in {4,8}");
  FEMB_CHECK_ARG(plan != nullptr && M >= 0 && nen >= 1 && nen <= 32 && n_nodes >= 1, "plan/M/nen(1..32)/n_nodes");
  FEMB_CHECK_ARG(n_nodes < (1ll << 31) - 1 && (long long)M * nen * nen < (1ll << 31) - 1, "index range: M*nen^2 and n_nodes must fit int32");
  auto* p = new femb_csr_plan();
  p->M = M, p->N = n_nodes, p->nen = nen;
  cudaStream_t s = as_stream(stream);
  const int rc = ib == 8 ? plan_build<long long>((const long long*)conn, p, s) : plan_build<int>((const int*)conn, p, s);
  if (rc != FEMB_OK) {
    plan_free(p);
    return rc;
  }
  *plan = p;
  if (nnz_nodes) *nnz_nodes = p->nnzn;
  return FEMB_OK;
}

extern "C" int femb_csr_plan_destroy(femb_csr_plan* plan) {
  plan_free(plan);
  return FEMB_OK;
}

extern "C" int femb_csr_plan_pattern(femb_csr_plan* p, int ndof, int32_t* crow, int32_t* col, femb_stream stream) {
  FEMB_CHECK_ARG(p && ndof >= 1 && ndof <= 6, "plan, 1 <= ndof <= 6");
  FEMB_CHECK_ARG(p->nnzn * ndof * ndof < (1ll << 31) - 1, "nnz must fit int32");
  pattern_kernel<<<grid_for(p->N * 32, 256), 256, 0, as_stream(stream)>>>(p->node_ptr, p->node_col, p->N, ndof, crow, col);
  FEMB_LAUNCH_CHECK();
  return FEMB_OK;
}

extern "C" int femb_csr_assemble(femb_csr_plan* p, int ndof, const double* Ke, double* vals, femb_stream stream) {
  FEMB_CHECK_ARG(p && ndof >= 1 && ndof <= 6, "plan, 1 <= ndof <= 6");
  const size_t per_warp = sizeof(double) * (size_t)p->max_row * ndof * ndof;
  FEMB_CHECK_ARG(per_warp <= 200 * 1024, "row too long for the shared-memory accumulator");
  cudaStream_t s = as_stream(stream);
  if (ndof == 1 && p->nen == 4 && p->inc_slots && (size_t)p->max_row * 8 * 128 <= 160 * 1024) {
    // scalar P1: one thread per row (a warp-per-row tile would leave 28 of 32 lanes idle on 4 entries per incidence)
    constexpr int BD = 128;
    const size_t smem = sizeof(double) * (size_t)p->max_row * BD;
    FEMB_CUDA(cudaFuncSetAttribute(assemble_p1_scalar_rows<1, BD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    assemble_p1_scalar_rows<1, BD><<<grid_for(p->N, BD, 16), BD, smem, s>>>(p->conn32, p->inc_ptr, p->inc, (const unsigned int*)p->inc_slots,
                                                                            p->node_ptr, p->N, nullptr, Ke, vals, nullptr);
    FEMB_LAUNCH_CHECK();
    return FEMB_OK;
  }
#define LAUNCH_GATHER(W)                                                                                            \
  {                                                                                                                  \
    FEMB_CUDA(cudaFuncSetAttribute(assemble_gather<W>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(per_warp * W))); \
    assemble_gather<W><<<grid_for(p->N, W, 16), W * 32, per_warp * W, s>>>(p->conn32, p->inc_ptr, p->inc, p->node_ptr, p->node_col, p->N, \
                                                                          p->nen, ndof, p->max_row, Ke, vals);      \
  }
  // batched variant: needs the slot table, at most 6 entries per lane and incidence, U*nen <= 256 position bytes
  static const bool gather_old = getenv("FEMB_GATHER_OLD") != nullptr;  // A/B switch
  const int R = (ndof * ndof * p->nen + 31) / 32;
  if (p->inc_slots && R <= 6 && p->nen <= 32 && per_warp * 8 + 8 * 256 <= 96 * 1024 && !gather_old) {
    const size_t smem = per_warp * 8 + 8 * 256;
#define LAUNCH_BATCHED(RV, UV)                                                                                                   \
  {                                                                                                                              \
    FEMB_CUDA(cudaFuncSetAttribute(assemble_gather_batched<8, RV, UV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    assemble_gather_batched<8, RV, UV><<<grid_for(p->N, 8, 16), 256, smem, s>>>(p->inc_ptr, p->inc, p->inc_slots, p->node_ptr, p->N, p->nen, \
                                                                             ndof, p->max_row, Ke, vals);                       \
  }
    switch (R) {
      case 1: LAUNCH_BATCHED(1, 8) break;
      case 2: LAUNCH_BATCHED(2, 6) break;
      case 3: LAUNCH_BATCHED(3, 4) break;
      case 4: LAUNCH_BATCHED(4, 3) break;
      case 5: LAUNCH_BATCHED(5, 2) break;
      default: LAUNCH_BATCHED(6, 2) break;
    }
#undef LAUNCH_BATCHED
    FEMB_LAUNCH_CHECK();
    return FEMB_OK;
  }
  if (per_warp * 8 <= 96 * 1024) LAUNCH_GATHER(8)
  else if (per_warp * 4 <= 200 * 1024) LAUNCH_GATHER(4)
  else LAUNCH_GATHER(1)
#undef LAUNCH_GATHER
  FEMB_LAUNCH_CHECK();
  return FEMB_OK;
}

extern "C" int femb_csr_assemble_c3d4(femb_csr_plan* p, int kind, const double* coords, double E, double nu, double* vals, int32_t* flag,
                                      femb_stream stream) {
  FEMB_CHECK_ARG(p && p->nen == 4 && (kind == 0 || kind == 1), "plan must be a 4-node plan; kind in {0,1}");
  const int dd = kind == 0 ? 1 : 9;
  constexpr int W = 8;
  const size_t per_warp = sizeof(double) * ((size_t)p->max_row * dd + (p->max_row + 1) / 2);
  FEMB_CHECK_ARG(per_warp * W <= 200 * 1024, "row too long for the shared-memory accumulator");
  const double c = E / ((1 + nu) * (1 - 2 * nu));
  cudaStream_t s = as_stream(stream);
  const int grid = grid_for(p->N, W, 16);
  static const bool no_tiles = getenv("FEMB_ASM_ROWS") != nullptr;  // A/B switch: previous thread-per-row kernel
  // Block-owned assembly (assembly_blocks.cu) is opt-in (FEMB_ASM_BLOCKS=1): on the 64 M-tet mesh it measured 2.65 ms against
  // 2.54 ms for the row-tile kernel below -- each element is computed once instead of four times, but the colour-ordered
  // shared-memory adds and coordinate gathers saturate the shared-memory pipe (ncu: l1tex 79 %, 43 % of the wavefronts are
  // bank conflicts; profiles/r02_ncu_assembly_blocks.txt).
  static const bool use_blocks = getenv("FEMB_ASM_BLOCKS") != nullptr && !no_tiles;
  if (kind == 0 && p->max_row <= 255 && p->M > 0 && use_blocks && !p->blk_failed) {
    if (!p->blk) {  // built at the first call: the clustering needs coordinates
      const int rc = block_plan_build(p, coords, s);
      if (rc != FEMB_OK) return rc;
    }
    if (p->blk) return block_assemble(p, coords, vals, flag, s);
  }
  if (kind == 0 && p->inc_slots && (size_t)p->max_row * 8 * 128 <= 160 * 1024 && !no_tiles) {
    constexpr int BD = 128;
    if (!p->rec) {  // lazily built, topology only
      const int rc = build_p1_records(p, s);
      if (rc != FEMB_OK) return rc;
    }
    Scratch scr(s);
    double* c4;
    FEMB_CUDA(scr.alloc(&c4, (size_t)4 * p->N));
    pad_coords<<<grid_for(p->N, 256), 256, 0, s>>>(coords, p->N, c4);
    const size_t smem = sizeof(double) * (size_t)p->max_row * BD;
    static const int occ = getenv("FEMB_ASM_VARIANT") ? atoi(getenv("FEMB_ASM_VARIANT")) : 2;  // measured on C4: 0 = 2.61 ms, 1 = 2.83, 2 = 2.55, 3 = 2.57; bit0: two incidences in flight, bit1: MUFU reciprocal
#define LAUNCH_TILES(V)                                                                                                          \
  {                                                                                                                              \
    FEMB_CUDA(cudaFuncSetAttribute(assemble_p1_poisson_tiles<BD, 4, V>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    assemble_p1_poisson_tiles<BD, 4, V><<<grid_for(p->ntiles, BD / 32, 16), BD, smem, s>>>(p->rec, p->tile_ptr, p->node_ptr, p->pdiag, p->N, \
                                                                                        p->ntiles, c4, vals, flag);             \
  }
    // register budget: the five-deep record pipeline needs all 128 registers; cutting it for 5 / 6 / 8 resident CTAs spills and
    // measured 3.4 / 6.1 / 8.2 ms against 2.55 ms
    if (occ == 1) LAUNCH_TILES(1) else if (occ == 2) LAUNCH_TILES(2) else if (occ == 3) LAUNCH_TILES(3) else LAUNCH_TILES(0)
#undef LAUNCH_TILES
  } else if (kind == 0 && p->inc_slots && (size_t)p->max_row * 8 * 128 <= 160 * 1024) {
    constexpr int BD = 128;
    const size_t smem = sizeof(double) * (size_t)p->max_row * BD;
    FEMB_CUDA(cudaFuncSetAttribute(assemble_p1_scalar_rows<0, BD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    assemble_p1_scalar_rows<0, BD><<<grid_for(p->N, BD, 16), BD, smem, s>>>(p->conn32, p->inc_ptr, p->inc, (const unsigned int*)p->inc_slots,
                                                                            p->node_ptr, p->N, coords, nullptr, vals, flag);
  } else if (kind == 1 && p->inc_slots && (size_t)p->max_row * 3 * 8 * 96 <= 200 * 1024 && !no_tiles) {
    if (!p->rec) {
      const int rc = build_p1_records(p, s);
      if (rc != FEMB_OK) return rc;
    }
    Scratch scr(s);
    double* c4;
    FEMB_CUDA(scr.alloc(&c4, (size_t)4 * p->N));
    pad_coords<<<grid_for(p->N, 256), 256, 0, s>>>(coords, p->N, c4);
    const size_t smem = sizeof(double) * (size_t)p->max_row * 3 * 96;
    FEMB_CUDA(cudaFuncSetAttribute(assemble_p1_elasticity_tiles, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    assemble_p1_elasticity_tiles<<<grid_for(p->ntiles, 1, 24), 96, smem, s>>>(p->rec, p->tile_ptr, p->node_ptr, p->pdiag, p->N, p->ntiles, c4, c * nu,
                                                                              c * (1 - 2 * nu) / 2, vals, flag);
  } else if (kind == 0) {
    FEMB_CUDA(cudaFuncSetAttribute(assemble_c3d4_fused<0, W>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(per_warp * W)));
    assemble_c3d4_fused<0, W><<<grid, W * 32, per_warp * W, s>>>(p->conn32, p->inc_ptr, p->inc, p->node_ptr, p->node_col, p->N, p->max_row, coords,
                                                                 0.0, 0.0, vals, flag);
  } else {
    FEMB_CUDA(cudaFuncSetAttribute(assemble_c3d4_fused<1, W>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(per_warp * W)));
    assemble_c3d4_fused<1, W><<<grid, W * 32, per_warp * W, s>>>(p->conn32, p->inc_ptr, p->inc, p->node_ptr, p->node_col, p->N, p->max_row, coords,
                                                                 c * nu, c * (1 - 2 * nu) / 2, vals, flag);
  }
  FEMB_LAUNCH_CHECK();
  return FEMB_OK;
}

namespace femb {
// compute_node_vm_stress element.py:466-504: node value = mean of the values of the elements containing the node
// (0 for isolated nodes).  The reference scatters with atomic index_add; here each node sums its incidence list in
// ascending element order, so the result is bit-reproducible.
template <typename T>
__global__ void node_average_kernel(const int* __restrict__ inc_ptr, const int* __restrict__ inc, long long N, int nen,
                                    const T* __restrict__ ev, T* __restrict__ out) {
  for (long long n = blockIdx.x * (long long)blockDim.x + threadIdx.x; n < N; n += (long long)gridDim.x * blockDim.x) {
    const int b = inc_ptr[n], e = inc_ptr[n + 1];
    T sum = 0;
    for (int k = b; k < e; ++k) sum += __ldg(ev + inc[k] / nen);
    out[n] = e > b ? sum / (T)(e - b) : T(0);
  }
}
}  // namespace femb

extern "C" int femb_node_average(femb_csr_plan* p, const void* elem_values, int fp, void* out, femb_stream stream) {
  FEMB_CHECK_ARG(p && (fp == 4 || fp == 8), "plan, fp in {4,8}");
  if (p->N == 0) return FEMB_OK;
  cudaStream_t s = as_stream(stream);
  const int grid = grid_for(p->N, 128);
  if (fp == 8) node_average_kernel<double><<<grid, 128, 0, s>>>(p->inc_ptr, p->inc, p->N, p->nen, (const double*)elem_values, (double*)out);
  else node_average_kernel<float><<<grid, 128, 0, s>>>(p->inc_ptr, p->inc, p->N, p->nen, (const float*)elem_values, (float*)out);
  FEMB_LAUNCH_CHECK();
  return FEMB_OK;
}

extern "C" int femb_ebe_apply(femb_csr_plan* p, int ndof, const void* Ke, const void* u, const void* unit, int fp, void* y, femb_stream stream) {
  FEMB_CHECK_ARG(p && (fp == 4 || fp == 8), "plan, fp in {4,8}");
  cudaStream_t s = as_stream(stream);
  if (unit) {
    FEMB_CHECK_ARG(ndof == 6, "shell operator has 6 dofs per node");
    const int grid = grid_for(p->N, 128);
    if (fp == 8) ebe_shell_kernel<double><<<grid, 128, 0, s>>>(p->conn32, p->inc_ptr, p->inc, p->N, p->nen, (const double*)Ke, (const double*)u, (const double*)unit, (double*)y);
    else ebe_shell_kernel<float><<<grid, 128, 0, s>>>(p->conn32, p->inc_ptr, p->inc, p->N, p->nen, (const float*)Ke, (const float*)u, (const float*)unit, (float*)y);
  } else {
    FEMB_CHECK_ARG(ndof >= 1 && ndof <= 6, "1 <= ndof <= 6");
    const int grid = grid_for(p->N * ndof, 128);
    if (fp == 8) ebe_apply_kernel<double><<<grid, 128, 0, s>>>(p->conn32, p->inc_ptr, p->inc, p->N, p->nen, ndof, (const double*)Ke, (const double*)u, (double*)y);
    else ebe_apply_kernel<float><<<grid, 128, 0, s>>>(p->conn32, p->inc_ptr, p->inc, p->N, p->nen, ndof, (const float*)Ke, (const float*)u, (float*)y);
  }
  FEMB_LAUNCH_CHECK();
  return FEMB_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
// shell_extrude (shell.py:885-983): per-node normals of a mid-surface mesh of triangles and quads, then bottom / top layers.
// The reference scatters unit face normals with index_add_ (triangles, then each quad's triangle (0,1,2), then its
// triangle (0,2,3)); here every node walks its incidence lists in that same order, so the sums are deterministic.
namespace femb {

template <typename T>
__device__ __forceinline__ void add_unit_normal(const T* __restrict__ X, int n0, int n1, int n2, T eps, T* s) {
  T a[3], b[3];
  for (int c = 0; c < 3; ++c) {
    const T x0 = __ldg(X + 3ll * n0 + c);
    a[c] = __ldg(X + 3ll * n1 + c) - x0;
    b[c] = __ldg(X + 3ll * n2 + c) - x0;
  }
  const T n[3] = {a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]};
  const T len = sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]) + eps;
  for (int c = 0; c < 3; ++c) s[c] += n[c] / len;
}

template <typename T>
__global__ void extrude_nodes_kernel(const int* __restrict__ tconn, const int* __restrict__ tptr, const int* __restrict__ tinc,
                                     const int* __restrict__ qconn, const int* __restrict__ qptr, const int* __restrict__ qinc,
                                     const T* __restrict__ X, long long N, T eps, T half_t, T* __restrict__ out) {
  for (long long n = blockIdx.x * (long long)blockDim.x + threadIdx.x; n < N; n += (long long)gridDim.x * blockDim.x) {
    T s[3] = {0, 0, 0};
    T cnt = 0;
    if (tptr) {
      for (int k = tptr[n]; k < tptr[n + 1]; ++k) {
        const int e = tinc[k] / 3;
        add_unit_normal(X, tconn[3 * e], tconn[3 * e + 1], tconn[3 * e + 2], eps, s);
        cnt += 1;
      }
    }
    if (qptr) {
      const int b = qptr[n], e1 = qptr[n + 1];
      for (int k = b; k < e1; ++k) {  // quad[:, :3]
        const int e = qinc[k] >> 2, a = qinc[k] & 3;
        if (a == 3) continue;
        add_unit_normal(X, qconn[4 * e], qconn[4 * e + 1], qconn[4 * e + 2], eps, s);
        cnt += 1;
      }
      for (int k = b; k < e1; ++k) {  // quad[:, [0,2,3]]
        const int e = qinc[k] >> 2, a = qinc[k] & 3;
        if (a == 1) continue;
        add_unit_normal(X, qconn[4 * e], qconn[4 * e + 2], qconn[4 * e + 3], eps, s);
        cnt += 1;
      }
    }
    T len = 0;
    for (int c = 0; c < 3; ++c) {
      s[c] = s[c] / (cnt + eps);
      len += s[c] * s[c];
    }
    len = sqrt(len) + eps;
    for (int c = 0; c < 3; ++c) {
      const T x = X[3 * n + c], d = half_t * (s[c] / len);
      out[3 * n + c] = x - d;
      out[3 * (n + N) + c] = x + d;
    }
  }
}

template <typename I>
__global__ void extrude_conn_kernel(const I* __restrict__ conn, long long M, int nen, long long N, I* __restrict__ out) {
  const long long total = M * 2 * nen;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const long long e = t / (2 * nen);
    const int a = (int)(t - e * 2 * nen);
    out[t] = a < nen ? conn[e * nen + a] : (I)(conn[e * nen + a - nen] + N);
  }
}

}  // namespace femb

extern "C" int femb_shell_extrude(femb_csr_plan* tri, femb_csr_plan* quad, const void* coords, int fp, int64_t N, double thickness, double eps,
                                  void* coords3d, femb_stream stream) {
  FEMB_CHECK_ARG((fp == 4 || fp == 8) && N >= 0 && coords && coords3d, "fp in {4,8}, N >= 0, non-null buffers");
  FEMB_CHECK_ARG(!tri || (tri->nen == 3 && tri->N == N), "triangle plan: nen = 3 and n_nodes = N");
  FEMB_CHECK_ARG(!quad || (quad->nen == 4 && quad->N == N), "quad plan: nen = 4 and n_nodes = N");
  if (N == 0) return FEMB_OK;
  cudaStream_t s = as_stream(stream);
  const int grid = grid_for(N, 128);
  const int *tc = tri ? tri->conn32 : nullptr, *tp = tri ? tri->inc_ptr : nullptr, *ti = tri ? tri->inc : nullptr;
  const int *qc = quad ? quad->conn32 : nullptr, *qp = quad ? quad->inc_ptr : nullptr, *qi = quad ? quad->inc : nullptr;
  if (fp == 8)
    extrude_nodes_kernel<double><<<grid, 128, 0, s>>>(tc, tp, ti, qc, qp, qi, (const double*)coords, N, eps, 0.5 * thickness, (double*)coords3d);
  else
    extrude_nodes_kernel<float><<<grid, 128, 0, s>>>(tc, tp, ti, qc, qp, qi, (const float*)coords, N, (float)eps, (float)(0.5 * thickness),
                                                     (float*)coords3d);
  FEMB_LAUNCH_CHECK();
  return FEMB_OK;
}

extern "C" int femb_extrude_connectivity(const void* conn, int ib, int64_t M, int nen, int64_t N, void* out, femb_stream stream) {
  FEMB_CHECK_ARG((ib == 4 || ib == 8) && M >= 0 && nen >= 1, "ib in {4,8}, M >= 0, nen >= 1");
  if (M == 0) return FEMB_OK;
  cudaStream_t s = as_stream(stream);
  const int grid = grid_for(M * 2 * nen, 256);
  if (ib == 8) extrude_conn_kernel<long long><<<grid, 256, 0, s>>>((const long long*)conn, M, nen, N, (long long*)out);
  else extrude_conn_kernel<int><<<grid, 256, 0, s>>>((const int*)conn, M, nen, N, (int*)out);
  FEMB_LAUNCH_CHECK();
  return FEMB_OK;
}

// Diagonal of sum_e K_e without assembling: out[node*d + alpha] = sum over the node's incidences (ascending element id) of
// Ke[e][a*d+alpha][a*d+alpha].  The lumped mass of vectorized_modal_solver (solver.py:1126-1131) and the Jacobi diagonal of
// compute_diagonal_preconditioner (solver.py:814-833) -- both index_add_ scatters in the reference.  col0 != 0 reproduces the
// reference preconditioner's strided-view bug, which sums column 0 of every row instead of the diagonal (solver.py:828).
namespace femb {
template <typename T>
__global__ void ebe_diag_kernel(const int* __restrict__ inc_ptr, const int* __restrict__ inc, long long N, int nen, int d, int col0,
                                const T* __restrict__ Ke, T* __restrict__ out) {
  const long long total = N * d;
  const int nd = nen * d;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const long long node = t / d;
    const int alpha = (int)(t - node * d);
    T sum = 0;
    for (int k = inc_ptr[node]; k < inc_ptr[node + 1]; ++k) {
      const int slot = inc[k], e = slot / nen, a = slot - e * nen;
      const int row = a * d + alpha;
      sum += __ldg(Ke + ((long long)e * nd + row) * nd + (col0 ? 0 : row));
    }
    out[t] = sum;
  }
}
}  // namespace femb

extern "C" int femb_ebe_diag(femb_csr_plan* p, int ndof, const void* Ke, int fp, int col0, void* out, femb_stream stream) {
  FEMB_CHECK_ARG(p && (fp == 4 || fp == 8) && ndof >= 1 && ndof <= 6 && Ke && out, "plan, fp in {4,8}, 1 <= ndof <= 6, non-null buffers");
  if (p->N == 0) return FEMB_OK;
  cudaStream_t s = as_stream(stream);
  const int grid = grid_for(p->N * ndof, 128);
  if (fp == 8) ebe_diag_kernel<double><<<grid, 128, 0, s>>>(p->inc_ptr, p->inc, p->N, p->nen, ndof, col0, (const double*)Ke, (double*)out);
  else ebe_diag_kernel<float><<<grid, 128, 0, s>>>(p->inc_ptr, p->inc, p->N, p->nen, ndof, col0, (const float*)Ke, (float*)out);
  FEMB_LAUNCH_CHECK();
  return FEMB_OK;
}
