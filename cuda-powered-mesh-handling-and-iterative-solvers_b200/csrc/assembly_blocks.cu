// Fused P1 Poisson assembly, block-owned (coords -> CSR values; the "assembled elems/s" kernel of the benchmark).
//
// The row-tile kernel in assembly.cu gives every CSR row to one thread that walks the row's incidence list: each element is
// recomputed by its 4 rows and the per-incidence records are a 4.1 GB stream (ncu: 6.2 GB of DRAM traffic against 3.6 GB
// algorithmic, fp64 pipe 32 %, 0.22 of the HBM peak).  Here a CTA owns a BLOCK of R rows that are close in space:
//   plan   nodes are ordered along a Morton curve of their coordinates and cut into blocks of R rows; a block lists every
//          element touching one of its rows (elements on a block boundary appear in up to 4 blocks: ~1.4x redundancy instead
//          of 4x), the nodes those elements reference outside the block ("halo"), and a 20-byte record per listed element:
//          4 block-local node indices + the 12 positions of its off-diagonal entries in the rows of its nodes.
//          Each block's elements are greedily coloured and sorted by colour: two elements of one colour never touch the same
//          OFF-DIAGONAL entry of an owned row (i.e. they share no edge with an owned end point; ~8 colours on a Kuhn mesh,
//          against >= 24 if they also had to avoid each other's diagonal).  The diagonal is not accumulated at all: every row of
//          the Laplace stiffness sums to zero (sum_b grad N_b = 0), so the write-out sets K_ii = -sum_{j != i} K_ij from the
//          finished off-diagonal sums (fixed shuffle tree), which agrees with the directly summed diagonal to rounding.
//   kernel the CTA stages the coordinates of its rows and halo nodes in shared memory, zeroes an accumulator laid out like
//          its rows' CSR segments, then goes colour by colour: one thread per element computes the 4x4 element matrix ONCE
//          from shared-memory coordinates and adds the rows of its owned nodes into the accumulator -- within a colour no
//          two threads touch the same row, across colours a barrier orders the adds, so every entry is summed in colour
//          order: deterministic without atomics.  Finally the rows are written out, 16 lanes per row.
// Bytes per element listed: 20 (record) + its share of the staged coordinates; per row: 8 (metadata) + 8 * len (values).
#include <cub/cub.cuh>

#include "plan.cuh"

namespace femb {

void block_plan_free(BlockPlan* b) {
  if (!b) return;
  cudaFree(b->blk_node);
  cudaFree(b->row_gstart);
  cudaFree(b->row_meta);
  cudaFree(b->blk_acc);
  cudaFree(b->blk_hptr);
  cudaFree(b->blk_halo);
  cudaFree(b->blk_eptr);
  cudaFree(b->blk_cptr);
  cudaFree(b->rec);
  cudaFree(b->rec2);
  delete b;
}

namespace {

// ---- plan kernels -------------------------------------------------------------------------------------------------------
__global__ void bbox_partial(const double* __restrict__ coords, long long N, double* __restrict__ out) {
  double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < N; i += (long long)gridDim.x * blockDim.x)
    for (int c = 0; c < 3; ++c) {
      const double v = coords[3 * i + c];
      lo[c] = fmin(lo[c], v), hi[c] = fmax(hi[c], v);
    }
  __shared__ double sh[6][256];
  for (int c = 0; c < 3; ++c) sh[c][threadIdx.x] = lo[c], sh[3 + c][threadIdx.x] = hi[c];
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s)
      for (int c = 0; c < 3; ++c) {
        sh[c][threadIdx.x] = fmin(sh[c][threadIdx.x], sh[c][threadIdx.x + s]);
        sh[3 + c][threadIdx.x] = fmax(sh[3 + c][threadIdx.x], sh[3 + c][threadIdx.x + s]);
      }
    __syncthreads();
  }
  if (threadIdx.x < 6) out[blockIdx.x * 6 + threadIdx.x] = sh[threadIdx.x][0];
}

__device__ __forceinline__ unsigned long long spread21(unsigned long long x) {  // bit i -> bit 3i
  x &= 0x1fffffull;
  x = (x | x << 32) & 0x1f00000000ffffull;
  x = (x | x << 16) & 0x1f0000ff0000ffull;
  x = (x | x << 8) & 0x100f00f00f00f00full;
  x = (x | x << 4) & 0x10c30c30c30c30c3ull;
  x = (x | x << 2) & 0x1249249249249249ull;
  return x;
}

struct Box {
  double lo[3], scale[3];
};

__global__ void morton_kernel(const double* __restrict__ coords, long long N, Box box, unsigned long long* __restrict__ keys, int* __restrict__ ids) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < N; i += (long long)gridDim.x * blockDim.x) {
    unsigned long long q[3];
    for (int c = 0; c < 3; ++c) {
      double t = (coords[3 * i + c] - box.lo[c]) * box.scale[c];
      t = fmin(fmax(t, 0.0), 2097151.0);
      q[c] = (unsigned long long)t;
    }
    keys[i] = spread21(q[0]) << 2 | spread21(q[1]) << 1 | spread21(q[2]);
    ids[i] = (int)i;
  }
}

__global__ void rank_kernel(const int* __restrict__ order, long long N, int* __restrict__ rank) {
  for (long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x; r < N; r += (long long)gridDim.x * blockDim.x) rank[order[r]] = (int)r;
}

__device__ __forceinline__ int lower_bound_dev(const int* a, int n, int v) {
  int lo = 0, hi = n;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (a[mid] < v) lo = mid + 1;
    else hi = mid;
  }
  return lo;
}

// one thread per block: row table (node, first CSR entry, accumulator offset, length, diagonal slot)
__global__ void block_rows_kernel(const int* __restrict__ order, long long N, int R, long long nblocks, const int* __restrict__ node_ptr,
                                  const int* __restrict__ node_col, int* __restrict__ blk_node, int* __restrict__ row_gstart,
                                  unsigned* __restrict__ row_meta, int* __restrict__ blk_acc) {
  for (long long b = blockIdx.x * (long long)blockDim.x + threadIdx.x; b < nblocks; b += (long long)gridDim.x * blockDim.x) {
    int off = 0;
    for (int j = 0; j < R; ++j) {
      const long long r = b * R + j;
      int node = -1, gs = 0, len = 0, diag = 0;
      if (r < N) {
        node = order[r];
        gs = node_ptr[node];
        len = node_ptr[node + 1] - gs;
        diag = lower_bound_dev(node_col + gs, len, node);
        if (diag >= len) diag = 0;  // a node without elements has an empty row
      }
      blk_node[r] = node;
      row_gstart[r] = gs;
      row_meta[r] = (unsigned)min(off, 65535) | (unsigned)len << 16 | (unsigned)diag << 24;
      off += len;
    }
    blk_acc[b] = off;
  }
}

// distinct blocks among an element's 4 nodes (first occurrences in local-node order)
__device__ __forceinline__ int element_blocks(const int* __restrict__ conn32, const int* __restrict__ rank, long long e, int R, int* blk) {
  const int4 q = __ldg(reinterpret_cast<const int4*>(conn32) + e);
  const int b4[4] = {rank[q.x] / R, rank[q.y] / R, rank[q.z] / R, rank[q.w] / R};
  int n = 0;
  for (int a = 0; a < 4; ++a) {
    bool seen = false;
    for (int k = 0; k < n; ++k) seen |= blk[k] == b4[a];
    if (!seen) blk[n++] = b4[a];
  }
  return n;
}

__global__ void count_pairs(const int* __restrict__ conn32, long long M, const int* __restrict__ rank, int R, int* __restrict__ cnt) {
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e <= M; e += (long long)gridDim.x * blockDim.x) {
    int blk[4];
    cnt[e] = e < M ? element_blocks(conn32, rank, e, R, blk) : 0;
  }
}

__global__ void emit_pairs(const int* __restrict__ conn32, long long M, const int* __restrict__ rank, int R, const long long* __restrict__ off,
                           unsigned long long* __restrict__ keys) {
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < M; e += (long long)gridDim.x * blockDim.x) {
    int blk[4];
    const int n = element_blocks(conn32, rank, e, R, blk);
    for (int k = 0; k < n; ++k) keys[off[e] + k] = (unsigned long long)blk[k] << 32 | (unsigned long long)e;
  }
}

// ptr[b] = first index whose key's high word is >= b (keys sorted)
__global__ void ptr_from_keys(const unsigned long long* __restrict__ keys, long long total, long long nblocks, int* __restrict__ ptr) {
  for (long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x; k <= total; k += (long long)gridDim.x * blockDim.x) {
    const long long prev = k == 0 ? -1 : (long long)(keys[k - 1] >> 32);
    const long long cur = k == total ? nblocks : (long long)(keys[k] >> 32);
    for (long long b = prev + 1; b <= cur; ++b) ptr[b] = (int)k;
  }
}

__global__ void count_halo(const unsigned long long* __restrict__ keys, long long total, const int* __restrict__ conn32,
                           const int* __restrict__ rank, int R, int* __restrict__ cnt) {
  for (long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x; k <= total; k += (long long)gridDim.x * blockDim.x) {
    int n = 0;
    if (k < total) {
      const int b = (int)(keys[k] >> 32);
      const int4 q = __ldg(reinterpret_cast<const int4*>(conn32) + (keys[k] & 0xffffffffull));
      n = (rank[q.x] / R != b) + (rank[q.y] / R != b) + (rank[q.z] / R != b) + (rank[q.w] / R != b);
    }
    cnt[k] = n;
  }
}

__global__ void emit_halo(const unsigned long long* __restrict__ keys, long long total, const int* __restrict__ conn32,
                          const int* __restrict__ rank, int R, const long long* __restrict__ off, unsigned long long* __restrict__ hkeys) {
  for (long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x; k < total; k += (long long)gridDim.x * blockDim.x) {
    const unsigned long long b = keys[k] >> 32;
    const int4 q = __ldg(reinterpret_cast<const int4*>(conn32) + (keys[k] & 0xffffffffull));
    const int nd[4] = {q.x, q.y, q.z, q.w};
    long long o = off[k];
    for (int a = 0; a < 4; ++a)
      if ((unsigned long long)(rank[nd[a]] / R) != b) hkeys[o++] = b << 32 | (unsigned long long)(unsigned)nd[a];
  }
}

__global__ void max_diff_kernel(const int* __restrict__ ptr, long long n, int* __restrict__ out) {
  int m = 0;
  for (long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x; k < n; k += (long long)gridDim.x * blockDim.x) m = max(m, ptr[k + 1] - ptr[k]);
  atomicMax(out, m);
}

__global__ void halo_ids(const unsigned long long* __restrict__ uniq, long long nh, int* __restrict__ out) {
  for (long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x; k < nh; k += (long long)gridDim.x * blockDim.x)
    out[k] = (int)(uniq[k] & 0xffffffffull);
}

// one thread per (block, element) pair, list order: the 20-byte record {4 block-local node indices, 12 slot bytes}
__global__ void record_kernel(const unsigned long long* __restrict__ keys, long long total, const int* __restrict__ conn32,
                              const int* __restrict__ rank, int R, const int* __restrict__ blk_hptr, const int* __restrict__ blk_halo,
                              const int* __restrict__ node_ptr, const int* __restrict__ node_col, uint4* __restrict__ rec,
                              unsigned* __restrict__ rec2, int* __restrict__ degenerate) {
  for (long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x; k < total; k += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(keys[k] >> 32);
    const int4 q = __ldg(reinterpret_cast<const int4*>(conn32) + (keys[k] & 0xffffffffull));
    const int nd[4] = {q.x, q.y, q.z, q.w};
    unsigned ln[4];
    unsigned char sl[12];
    const int h0 = blk_hptr[b], nh = blk_hptr[b + 1] - h0;
    for (int a = 0; a < 4; ++a) {
      const int r = rank[nd[a]];
      const bool owned = r / R == b;
      ln[a] = owned ? (unsigned)(r - b * R) : (unsigned)(R + lower_bound_dev(blk_halo + h0, nh, nd[a]));
      const int s = node_ptr[nd[a]], len = node_ptr[nd[a] + 1] - s;
      int t = 0;
      for (int c = 0; c < 4; ++c) {
        if (c == a) continue;
        sl[a * 3 + t] = owned ? (unsigned char)lower_bound_dev(node_col + s, len, nd[c]) : 0;
        ++t;
      }
    }
    auto pack = [&](int i) { return (unsigned)sl[i] | (unsigned)sl[i + 1] << 8 | (unsigned)sl[i + 2] << 16 | (unsigned)sl[i + 3] << 24; };
    if (nd[0] == nd[1] || nd[0] == nd[2] || nd[0] == nd[3] || nd[1] == nd[2] || nd[1] == nd[3] || nd[2] == nd[3]) *degenerate = 1;
    rec[k] = make_uint4(ln[0] | ln[1] << 16, ln[2] | ln[3] << 16, pack(0), pack(4));
    rec2[k] = pack(8);
  }
}

// Greedy colouring, one warp per block (lane 0 walks the block's elements in ascending id; the other lanes only help with
// the tables): colour = lowest bit not used by any off-diagonal entry the element adds to (entries of owned rows; an entry is
// identified by its position in the block accumulator).  Then a stable counting sort by colour: dest[k] = position of pair k
// inside its block's colour-sorted list.
__global__ void color_kernel(const uint4* __restrict__ rec, const unsigned* __restrict__ rec2, const int* __restrict__ blk_eptr,
                             const unsigned* __restrict__ row_meta, int R, int max_acc, long long nblocks, unsigned char* __restrict__ color,
                             unsigned short* __restrict__ blk_cptr, int* __restrict__ dest, int* __restrict__ stats) {
  extern __shared__ unsigned long long mask[];                     // [max_acc]
  unsigned* meta = reinterpret_cast<unsigned*>(mask + max_acc);    // [R]
  __shared__ int cnt[BLK_MAXC + 1];
  for (long long b = blockIdx.x; b < nblocks; b += gridDim.x) {
    for (int j = threadIdx.x; j < max_acc; j += blockDim.x) mask[j] = 0ull;
    for (int j = threadIdx.x; j < R; j += blockDim.x) meta[j] = row_meta[b * R + j];
    for (int c = threadIdx.x; c <= BLK_MAXC; c += blockDim.x) cnt[c] = 0;
    __syncthreads();
    const int e0 = blk_eptr[b], e1 = blk_eptr[b + 1];
    if (threadIdx.x == 0) {
      int maxc = 0;
      bool overflow = false;
      for (int k = e0; k < e1; ++k) {
        const uint4 q1 = rec[k];
        const unsigned slots[3] = {q1.z, q1.w, rec2[k]};
        const unsigned ln[4] = {q1.x & 0xffffu, q1.x >> 16, q1.y & 0xffffu, q1.y >> 16};
        int idx[12], n = 0;
        unsigned long long used = 0ull;
        for (int a = 0; a < 4; ++a) {
          if (ln[a] >= (unsigned)R) continue;
          const int base = (int)(meta[ln[a]] & 0xffffu);
          for (int t = 0; t < 3; ++t) {
            const int kk = a * 3 + t;
            idx[n] = base + (int)((slots[kk >> 2] >> (8 * (kk & 3))) & 255u);
            used |= mask[idx[n]];
            ++n;
          }
        }
        int c = __ffsll((long long)~used) - 1;
        if (c < 0) c = BLK_MAXC - 1, overflow = true;  // more than 64 colours: reported, the caller falls back
        for (int i = 0; i < n; ++i) mask[idx[i]] |= 1ull << c;
        color[k] = (unsigned char)c;
        cnt[c + 1] += 1;
        maxc = max(maxc, c + 1);
      }
      int largest = 0;
      for (int c = 0; c < BLK_MAXC; ++c) largest = max(largest, cnt[c + 1]), cnt[c + 1] += cnt[c];
      atomicMax(&stats[0], maxc);
      atomicMax(&stats[1], e1 - e0);
      if (overflow) atomicMax(&stats[2], 1);
    }
    __syncthreads();
    for (int c = threadIdx.x; c <= BLK_MAXC; c += blockDim.x) blk_cptr[b * (BLK_MAXC + 1) + c] = (unsigned short)min(cnt[c], 65535);
    __syncthreads();
    if (threadIdx.x == 0)
      for (int k = e0; k < e1; ++k) dest[k] = e0 + cnt[color[k]]++;
    __syncthreads();
  }
}

__global__ void permute_records(const uint4* __restrict__ rin, const unsigned* __restrict__ rin2, const int* __restrict__ dest, long long total,
                                uint4* __restrict__ rout, unsigned* __restrict__ rout2) {
  for (long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x; k < total; k += (long long)gridDim.x * blockDim.x) {
    const int d = dest[k];
    rout[d] = rin[k];
    rout2[d] = rin2[k];
  }
}

// ---- the assembly kernel -----------------------------------------------------------------------------------------------
struct BlockArgs {
  int R, max_acc, max_local;
  const int *blk_node, *row_gstart, *blk_acc, *blk_hptr, *blk_halo, *blk_eptr;
  const unsigned* row_meta;
  const unsigned short* blk_cptr;
  const uint4* rec;
  const unsigned* rec2;
};

template <int THREADS>
__global__ void __launch_bounds__(THREADS) assemble_p1_blocks(const BlockArgs A, const double* __restrict__ coords, double* __restrict__ vals,
                                                             int* __restrict__ flag) {
  extern __shared__ __align__(16) unsigned char smraw[];
  double* acc = reinterpret_cast<double*>(smraw);                        // [max_acc] laid out like the rows' CSR segments
  double* xs = acc + A.max_acc;                                          // [max_local][3]
  unsigned* meta = reinterpret_cast<unsigned*>(xs + 3 * (size_t)A.max_local);  // [R]
  __shared__ unsigned short cptr[BLK_MAXC + 1];
  const int tid = threadIdx.x, R = A.R;
  const long long b = blockIdx.x;
  const int e0 = A.blk_eptr[b];
  if (tid <= BLK_MAXC) cptr[tid] = A.blk_cptr[b * (BLK_MAXC + 1) + tid];
  // first colour's record is requested before anything else: its latency hides behind the staging below
  uint4 r1 = make_uint4(0, 0, 0, 0);
  unsigned r2 = 0;
  const int first_cnt = A.blk_cptr[b * (BLK_MAXC + 1) + 1];
  if (tid < first_cnt) r1 = __ldg(A.rec + e0 + tid), r2 = __ldg(A.rec2 + e0 + tid);
  {  // pull the whole record list of the block into L2 now: the per-colour loads below then see L2 latency, not DRAM latency
    const int cnt = A.blk_cptr[b * (BLK_MAXC + 1) + BLK_MAXC];
    const char* p1 = reinterpret_cast<const char*>(A.rec + e0);
    const char* p2 = reinterpret_cast<const char*>(A.rec2 + e0);
    for (int o = tid * 128; o < cnt * 16; o += THREADS * 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(p1 + o));
    for (int o = tid * 128; o < cnt * 4; o += THREADS * 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(p2 + o));
  }
  const int nacc = A.blk_acc[b];
  for (int t = tid; t < nacc; t += THREADS) acc[t] = 0.0;
  for (int j = tid; j < R; j += THREADS) {
    meta[j] = A.row_meta[b * R + j];
    const int node = A.blk_node[b * R + j];
    if (node >= 0) {
      xs[3 * j] = __ldg(coords + 3ll * node), xs[3 * j + 1] = __ldg(coords + 3ll * node + 1), xs[3 * j + 2] = __ldg(coords + 3ll * node + 2);
    }
  }
  {
    const int h0 = A.blk_hptr[b], nh = A.blk_hptr[b + 1] - h0;
    for (int h = tid; h < nh; h += THREADS) {
      const int node = __ldg(A.blk_halo + h0 + h);
      double* d = xs + 3 * (size_t)(R + h);
      d[0] = __ldg(coords + 3ll * node), d[1] = __ldg(coords + 3ll * node + 1), d[2] = __ldg(coords + 3ll * node + 2);
    }
  }
  __syncthreads();
  const int total = cptr[BLK_MAXC];
  for (int c = 0; c < BLK_MAXC; ++c) {
    const int lo = cptr[c], hi = cptr[c + 1];
    if (lo >= total) break;
    // this colour's record was requested one colour ago; request the next colour's before the arithmetic
    const uint4 c1 = r1;
    const unsigned c2 = r2;
    if (c + 1 < BLK_MAXC) {
      const int ni = hi + tid;
      if (ni < (int)cptr[min(c + 2, BLK_MAXC)]) r1 = __ldg(A.rec + e0 + ni), r2 = __ldg(A.rec2 + e0 + ni);
    }
    for (int i = lo + tid; i < hi; i += THREADS) {
      uint4 q1 = c1;
      unsigned q2 = c2;
      if (i != lo + tid) q1 = __ldg(A.rec + e0 + i), q2 = __ldg(A.rec2 + e0 + i);  // colours larger than the CTA (rare)
      const unsigned ln[4] = {q1.x & 0xffffu, q1.x >> 16, q1.y & 0xffffu, q1.y >> 16};
      const double* x0 = xs + 3 * ln[0];
      const double* x1 = xs + 3 * ln[1];
      const double* x2 = xs + 3 * ln[2];
      const double* x3 = xs + 3 * ln[3];
      double e1[3], e2[3], e3[3], cv[4][3];
#pragma unroll
      for (int t = 0; t < 3; ++t) {
        const double o = x0[t];
        e1[t] = x1[t] - o, e2[t] = x2[t] - o, e3[t] = x3[t] - o;
      }
      cv[1][0] = e2[1] * e3[2] - e2[2] * e3[1], cv[1][1] = e2[2] * e3[0] - e2[0] * e3[2], cv[1][2] = e2[0] * e3[1] - e2[1] * e3[0];
      cv[2][0] = e3[1] * e1[2] - e3[2] * e1[1], cv[2][1] = e3[2] * e1[0] - e3[0] * e1[2], cv[2][2] = e3[0] * e1[1] - e3[1] * e1[0];
      cv[3][0] = e1[1] * e2[2] - e1[2] * e2[1], cv[3][1] = e1[2] * e2[0] - e1[0] * e2[2], cv[3][2] = e1[0] * e2[1] - e1[1] * e2[0];
      const double det = e1[0] * cv[1][0] + e1[1] * cv[1][1] + e1[2] * cv[1][2];
      if (fabs(det) < 1e-12 && flag) *flag = 1;
#pragma unroll
      for (int t = 0; t < 3; ++t) cv[0][t] = -(cv[1][t] + cv[2][t] + cv[3][t]);
      // V g_a.g_b = c_a.c_b / (6|det|); reciprocal by MUFU seed + two Newton steps (branch-free, ~1 ulp), as the row-tile kernel
      const double d6 = 6.0 * fabs(det);
      double rc;
      asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(rc) : "d"(d6));
      rc = rc * (2.0 - d6 * rc);
      rc = rc * (2.0 - d6 * rc);
      double kk[4][4];  // off-diagonal entries only (6 unique)
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int bb = a + 1; bb < 4; ++bb) {
          const double v = (cv[a][0] * cv[bb][0] + cv[a][1] * cv[bb][1] + cv[a][2] * cv[bb][2]) * rc;
          kk[a][bb] = v, kk[bb][a] = v;
        }
      const unsigned slots[3] = {q1.z, q1.w, q2};
      // The four entries of a row are distinct (the plan rejects elements with a repeated node), so they are read together,
      // added and written back: one shared-memory round trip per row instead of four dependent ones.
      unsigned mrow[4];
#pragma unroll
      for (int a = 0; a < 4; ++a) mrow[a] = ln[a] < (unsigned)R ? meta[ln[a]] : 0xffffffffu;  // owned row: it lives in this block's accumulator
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        if (mrow[a] != 0xffffffffu) {
          double* row = acc + (mrow[a] & 0xffffu);
          int pos[3];
          double add[3];
          int t = 0;
#pragma unroll
          for (int bb = 0; bb < 4; ++bb) {
            if (bb == a) continue;
            const int k = a * 3 + t;
            pos[t] = (int)((slots[k >> 2] >> (8 * (k & 3))) & 255u);
            add[t] = kk[a][bb];
            ++t;
          }
          const double o0 = row[pos[0]], o1 = row[pos[1]], o2 = row[pos[2]];
          row[pos[0]] = o0 + add[0], row[pos[1]] = o1 + add[1], row[pos[2]] = o2 + add[2];
        }
      }
    }
    __syncthreads();
  }
  // write-out, 16 lanes per row: the diagonal is minus the sum of the row's off-diagonal entries (zero row sums of the Laplace
  // stiffness), summed lane-strided and then over a fixed shuffle tree -- the accumulator holds 0 at the diagonal position
  const int hl = tid & 15;
  for (int j0 = 0; j0 < R; j0 += THREADS / 16) {  // uniform trip count: the shuffles below need all 16 lanes of a row
    const int j = j0 + (tid >> 4);
    const unsigned m = j < R ? meta[j] : 0u;
    const int len = (m >> 16) & 255, diag = (int)(m >> 24);
    const double* row = acc + (m & 0xffffu);
    double* dst = j < R ? vals + A.row_gstart[b * R + j] : vals;
    double sum = 0.0;
    for (int t = hl; t < len; t += 16) sum += row[t];
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    for (int t = hl; t < len; t += 16) dst[t] = t == diag ? -sum : row[t];
  }
}

template <typename T>
int device_scan_ll(Scratch& scr, const int* in, long long* out, long long n, cudaStream_t s) {
  size_t tb = 0;
  FEMB_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tb, in, out, (int)n, s));
  void* tmp;
  FEMB_CUDA(scr.alloc((char**)&tmp, tb));
  FEMB_CUDA(cub::DeviceScan::ExclusiveSum(tmp, tb, in, out, (int)n, s));
  return FEMB_OK;
}

int sort_keys_u64(Scratch& scr, unsigned long long* in, unsigned long long* out, long long n, int end_bit, cudaStream_t s) {
  size_t tb = 0;
  FEMB_CUDA(cub::DeviceRadixSort::SortKeys(nullptr, tb, in, out, (int)n, 0, end_bit, s));
  void* tmp;
  FEMB_CUDA(scr.alloc((char**)&tmp, tb));
  FEMB_CUDA(cub::DeviceRadixSort::SortKeys(tmp, tb, in, out, (int)n, 0, end_bit, s));
  return FEMB_OK;
}

int bits_for(long long v) {
  int b = 1;
  while ((1ll << b) < v) ++b;
  return b;
}

constexpr size_t BLK_SMEM_LIMIT = 110 * 1024;  // two CTAs per SM

size_t block_smem(const BlockPlan* bp) { return sizeof(double) * ((size_t)bp->max_acc + 3 * (size_t)(bp->R + bp->max_halo)) + sizeof(unsigned) * bp->R; }

// FEMB_OK = built, BLK_TOO_BIG = this R does not fit (try a smaller one), BLK_DEGENERATE = an element repeats a node, else = error
constexpr int BLK_TOO_BIG = -100, BLK_DEGENERATE = -101;
int try_build(femb_csr_plan* p, const double* coords, int R, cudaStream_t s, BlockPlan** out) {
  const long long N = p->N, M = p->M;
  const long long nblocks = (N + R - 1) / R;
  BlockPlan* bp = new BlockPlan();
  bp->R = R, bp->nblocks = nblocks;
  struct Guard {
    BlockPlan*& b;
    bool keep = false;
    ~Guard() {
      if (!keep) block_plan_free(b), b = nullptr;
    }
  } guard{bp};
  Scratch scr(s);
  // ---- Morton order of the nodes
  const int gb = 256;
  double* part;
  FEMB_CUDA(scr.alloc(&part, (size_t)gb * 6));
  bbox_partial<<<gb, 256, 0, s>>>(coords, N, part);
  FEMB_LAUNCH_CHECK();
  static thread_local double hpart[256 * 6];
  FEMB_CUDA(cudaMemcpyAsync(hpart, part, sizeof(double) * gb * 6, cudaMemcpyDeviceToHost, s));
  FEMB_CUDA(cudaStreamSynchronize(s));
  Box box;
  for (int c = 0; c < 3; ++c) {
    double lo = 1e300, hi = -1e300;
    for (int g = 0; g < gb; ++g) lo = std::min(lo, hpart[g * 6 + c]), hi = std::max(hi, hpart[g * 6 + 3 + c]);
    box.lo[c] = lo;
    box.scale[c] = hi > lo ? 2097151.0 / (hi - lo) : 0.0;
  }
  {  // one common scale: the curve follows the geometry, not the bounding box's aspect ratio
    double smin = 1e300;
    for (int c = 0; c < 3; ++c)
      if (box.scale[c] > 0.0) smin = std::min(smin, box.scale[c]);
    for (int c = 0; c < 3; ++c)
      if (box.scale[c] > 0.0) box.scale[c] = smin;
  }
  unsigned long long *mk, *mk2;
  int *ids, *order, *rank;
  FEMB_CUDA(scr.alloc(&mk, (size_t)N));
  FEMB_CUDA(scr.alloc(&mk2, (size_t)N));
  FEMB_CUDA(scr.alloc(&ids, (size_t)N));
  FEMB_CUDA(scr.alloc(&order, (size_t)N));
  FEMB_CUDA(scr.alloc(&rank, (size_t)N));
  morton_kernel<<<grid_for(N, 256), 256, 0, s>>>(coords, N, box, mk, ids);
  FEMB_LAUNCH_CHECK();
  {
    size_t tb = 0;
    FEMB_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tb, mk, mk2, ids, order, (int)N, 0, 63, s));
    void* tmp;
    FEMB_CUDA(scr.alloc((char**)&tmp, tb));
    FEMB_CUDA(cub::DeviceRadixSort::SortPairs(tmp, tb, mk, mk2, ids, order, (int)N, 0, 63, s));
  }
  rank_kernel<<<grid_for(N, 256), 256, 0, s>>>(order, N, rank);
  FEMB_LAUNCH_CHECK();
  // ---- row tables
  FEMB_CUDA(cudaMalloc(&bp->blk_node, sizeof(int) * nblocks * R));
  FEMB_CUDA(cudaMalloc(&bp->row_gstart, sizeof(int) * nblocks * R));
  FEMB_CUDA(cudaMalloc(&bp->row_meta, sizeof(unsigned) * nblocks * R));
  FEMB_CUDA(cudaMalloc(&bp->blk_acc, sizeof(int) * nblocks));
  block_rows_kernel<<<grid_for(nblocks, 64), 64, 0, s>>>(order, N, R, nblocks, p->node_ptr, p->node_col, bp->blk_node, bp->row_gstart, bp->row_meta,
                                                         bp->blk_acc);
  FEMB_LAUNCH_CHECK();
  int* dmax;
  FEMB_CUDA(scr.alloc(&dmax, 4));
  {
    size_t tb = 0;
    FEMB_CUDA(cub::DeviceReduce::Max(nullptr, tb, bp->blk_acc, dmax, (int)nblocks, s));
    void* tmp;
    FEMB_CUDA(scr.alloc((char**)&tmp, tb));
    FEMB_CUDA(cub::DeviceReduce::Max(tmp, tb, bp->blk_acc, dmax, (int)nblocks, s));
  }
  FEMB_CUDA(cudaMemcpyAsync(&bp->max_acc, dmax, sizeof(int), cudaMemcpyDeviceToHost, s));
  FEMB_CUDA(cudaStreamSynchronize(s));
  if (bp->max_acc > 65535 || sizeof(double) * (size_t)bp->max_acc > BLK_SMEM_LIMIT) return BLK_TOO_BIG;
  // ---- (block, element) pairs, sorted by block then element id
  int* cnt;
  long long* off;
  FEMB_CUDA(scr.alloc(&cnt, (size_t)M + 1));
  FEMB_CUDA(scr.alloc(&off, (size_t)M + 1));
  count_pairs<<<grid_for(M + 1, 256), 256, 0, s>>>(p->conn32, M, rank, R, cnt);
  FEMB_LAUNCH_CHECK();
  if (device_scan_ll<int>(scr, cnt, off, M + 1, s) != FEMB_OK) return FEMB_ERR_CUDA;
  long long total = 0;
  FEMB_CUDA(cudaMemcpyAsync(&total, off + M, sizeof(long long), cudaMemcpyDeviceToHost, s));
  FEMB_CUDA(cudaStreamSynchronize(s));
  if (total >= (1ll << 31) - 1) return BLK_TOO_BIG;
  bp->total = total;
  unsigned long long *keys0, *keys;
  FEMB_CUDA(scr.alloc(&keys0, (size_t)std::max<long long>(total, 1)));
  FEMB_CUDA(scr.alloc(&keys, (size_t)std::max<long long>(total, 1)));
  if (M > 0) emit_pairs<<<grid_for(M, 256), 256, 0, s>>>(p->conn32, M, rank, R, off, keys0);
  FEMB_LAUNCH_CHECK();
  if (sort_keys_u64(scr, keys0, keys, total, 32 + bits_for(nblocks + 1), s) != FEMB_OK) return FEMB_ERR_CUDA;
  FEMB_CUDA(cudaMalloc(&bp->blk_eptr, sizeof(int) * (nblocks + 1)));
  ptr_from_keys<<<grid_for(total + 1, 256), 256, 0, s>>>(keys, total, nblocks, bp->blk_eptr);
  FEMB_LAUNCH_CHECK();
  // ---- halo lists: distinct non-owned nodes of each block's elements
  int* hcnt;
  long long* hoff;
  FEMB_CUDA(scr.alloc(&hcnt, (size_t)total + 1));
  FEMB_CUDA(scr.alloc(&hoff, (size_t)total + 1));
  count_halo<<<grid_for(total + 1, 256), 256, 0, s>>>(keys, total, p->conn32, rank, R, hcnt);
  FEMB_LAUNCH_CHECK();
  if (device_scan_ll<int>(scr, hcnt, hoff, total + 1, s) != FEMB_OK) return FEMB_ERR_CUDA;
  long long htotal = 0;
  FEMB_CUDA(cudaMemcpyAsync(&htotal, hoff + total, sizeof(long long), cudaMemcpyDeviceToHost, s));
  FEMB_CUDA(cudaStreamSynchronize(s));
  if (htotal >= (1ll << 31) - 1) return BLK_TOO_BIG;
  unsigned long long *hk0, *hk1, *huniq;
  long long* dnh;
  FEMB_CUDA(scr.alloc(&hk0, (size_t)std::max<long long>(htotal, 1)));
  FEMB_CUDA(scr.alloc(&hk1, (size_t)std::max<long long>(htotal, 1)));
  FEMB_CUDA(scr.alloc(&huniq, (size_t)std::max<long long>(htotal, 1)));
  FEMB_CUDA(scr.alloc(&dnh, 1));
  if (total > 0) emit_halo<<<grid_for(total, 256), 256, 0, s>>>(keys, total, p->conn32, rank, R, hoff, hk0);
  FEMB_LAUNCH_CHECK();
  if (sort_keys_u64(scr, hk0, hk1, htotal, 32 + bits_for(nblocks + 1), s) != FEMB_OK) return FEMB_ERR_CUDA;
  {
    size_t tb = 0;
    FEMB_CUDA(cub::DeviceSelect::Unique(nullptr, tb, hk1, huniq, dnh, (int)htotal, s));
    void* tmp;
    FEMB_CUDA(scr.alloc((char**)&tmp, tb));
    FEMB_CUDA(cub::DeviceSelect::Unique(tmp, tb, hk1, huniq, dnh, (int)htotal, s));
  }
  long long nh = 0;
  FEMB_CUDA(cudaMemcpyAsync(&nh, dnh, sizeof(long long), cudaMemcpyDeviceToHost, s));
  FEMB_CUDA(cudaStreamSynchronize(s));
  FEMB_CUDA(cudaMalloc(&bp->blk_hptr, sizeof(int) * (nblocks + 1)));
  FEMB_CUDA(cudaMalloc(&bp->blk_halo, sizeof(int) * std::max<long long>(nh, 1)));
  ptr_from_keys<<<grid_for(nh + 1, 256), 256, 0, s>>>(huniq, nh, nblocks, bp->blk_hptr);
  if (nh > 0) halo_ids<<<grid_for(nh, 256), 256, 0, s>>>(huniq, nh, bp->blk_halo);
  FEMB_LAUNCH_CHECK();
  int* stats;  // [0] colours, [1] elements per block, [2] colour overflow, [3] halo nodes per block
  FEMB_CUDA(scr.alloc(&stats, 4));
  FEMB_CUDA(cudaMemsetAsync(stats, 0, 4 * sizeof(int), s));
  max_diff_kernel<<<grid_for(nblocks, 256), 256, 0, s>>>(bp->blk_hptr, nblocks, stats + 3);
  FEMB_LAUNCH_CHECK();
  // ---- records in list order, colours from the records, then the records move to their colour-sorted positions
  int hstats[4];
  FEMB_CUDA(cudaMemcpyAsync(hstats, stats, sizeof(hstats), cudaMemcpyDeviceToHost, s));
  FEMB_CUDA(cudaStreamSynchronize(s));
  bp->max_halo = hstats[3];
  if (R + bp->max_halo > 65535 || block_smem(bp) > BLK_SMEM_LIMIT) return BLK_TOO_BIG;
  uint4* rtmp;
  unsigned* rtmp2;
  FEMB_CUDA(scr.alloc(&rtmp, (size_t)std::max<long long>(total, 1)));
  FEMB_CUDA(scr.alloc(&rtmp2, (size_t)std::max<long long>(total, 1)));
  if (total > 0)
    record_kernel<<<grid_for(total, 256), 256, 0, s>>>(keys, total, p->conn32, rank, R, bp->blk_hptr, bp->blk_halo, p->node_ptr, p->node_col, rtmp, rtmp2,
                                                       stats + 2);
  FEMB_LAUNCH_CHECK();
  FEMB_CUDA(cudaMemcpyAsync(hstats, stats, sizeof(hstats), cudaMemcpyDeviceToHost, s));
  FEMB_CUDA(cudaStreamSynchronize(s));
  if (hstats[2]) return BLK_DEGENERATE;  // an element with a repeated node: the batched row updates need distinct entries -> row-tile kernel
  unsigned char* color;
  int* dest;
  FEMB_CUDA(scr.alloc(&color, (size_t)std::max<long long>(total, 1)));
  FEMB_CUDA(scr.alloc(&dest, (size_t)std::max<long long>(total, 1)));
  FEMB_CUDA(cudaMalloc(&bp->blk_cptr, sizeof(unsigned short) * nblocks * (BLK_MAXC + 1)));
  const size_t csm = sizeof(unsigned long long) * bp->max_acc + sizeof(unsigned) * R;
  FEMB_CUDA(cudaFuncSetAttribute(color_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)csm));
  color_kernel<<<(int)std::min<long long>(nblocks, 148 * 16), 32, csm, s>>>(rtmp, rtmp2, bp->blk_eptr, bp->row_meta, R, bp->max_acc, nblocks, color,
                                                                         bp->blk_cptr, dest, stats);
  FEMB_LAUNCH_CHECK();
  FEMB_CUDA(cudaMemcpyAsync(hstats, stats, sizeof(hstats), cudaMemcpyDeviceToHost, s));
  FEMB_CUDA(cudaStreamSynchronize(s));
  bp->max_colors = hstats[0], bp->max_elems = hstats[1];
  if (hstats[2] || bp->max_elems > 65535) return BLK_TOO_BIG;
  FEMB_CUDA(cudaMalloc(&bp->rec, sizeof(uint4) * std::max<long long>(total, 1)));
  FEMB_CUDA(cudaMalloc(&bp->rec2, sizeof(unsigned) * std::max<long long>(total, 1)));
  if (total > 0) permute_records<<<grid_for(total, 256), 256, 0, s>>>(rtmp, rtmp2, dest, total, bp->rec, bp->rec2);
  FEMB_LAUNCH_CHECK();
  FEMB_CUDA(cudaStreamSynchronize(s));
  guard.keep = true;
  *out = bp;
  return FEMB_OK;
}

}  // namespace

// Builds the block plan (once per plan; needs coordinates for the clustering -- any coordinates give a CORRECT plan, nearby
// nodes in one block only make it efficient).  Returns FEMB_OK with p->blk set, or FEMB_OK with p->blk_failed when no block
// size fits the shared-memory budget (the caller keeps using the row-tile kernel).
int block_plan_build(femb_csr_plan* p, const double* coords, cudaStream_t s) {
  static const int r_env = getenv("FEMB_ASM_BLOCK_ROWS") ? atoi(getenv("FEMB_ASM_BLOCK_ROWS")) : 0;
  cudaEvent_t t0, t1;
  FEMB_CUDA(cudaEventCreate(&t0));
  FEMB_CUDA(cudaEventCreate(&t1));
  cudaEventRecord(t0, s);
  int rc = BLK_TOO_BIG;
  for (int R = r_env > 0 ? r_env : 256; R >= 64 && rc == BLK_TOO_BIG; R /= 2) {
    BlockPlan* bp = nullptr;
    rc = try_build(p, coords, R, s, &bp);
    if (rc == FEMB_OK) p->blk = bp;
    if (r_env > 0) break;
  }
  cudaEventRecord(t1, s);
  cudaEventSynchronize(t1);
  float ms = 0.f;
  cudaEventElapsedTime(&ms, t0, t1);
  cudaEventDestroy(t0);
  cudaEventDestroy(t1);
  if (rc == BLK_TOO_BIG || rc == BLK_DEGENERATE) p->blk_failed = true;
  else if (rc != FEMB_OK) return rc;
  if (p->blk) p->blk->build_ms = ms;
  if (getenv("FEMB_ASM_VERBOSE") && p->blk)
    fprintf(stderr, "[femb] block plan: R=%d blocks=%lld listed elements=%lld (%.3fx of %lld) colours<=%d elems/block<=%d halo<=%d acc<=%d smem=%zu B, built in %.1f ms\n",
            p->blk->R, p->blk->nblocks, p->blk->total, (double)p->blk->total / (double)std::max<long long>(p->M, 1), p->M, p->blk->max_colors,
            p->blk->max_elems, p->blk->max_halo, p->blk->max_acc, block_smem(p->blk), ms);
  return FEMB_OK;
}

int block_assemble(femb_csr_plan* p, const double* coords, double* vals, int* flag, cudaStream_t s) {
  const BlockPlan* bp = p->blk;
  BlockArgs A;
  A.R = bp->R, A.max_acc = bp->max_acc, A.max_local = bp->R + bp->max_halo;
  A.blk_node = bp->blk_node, A.row_gstart = bp->row_gstart, A.blk_acc = bp->blk_acc, A.blk_hptr = bp->blk_hptr, A.blk_halo = bp->blk_halo;
  A.blk_eptr = bp->blk_eptr, A.row_meta = bp->row_meta, A.blk_cptr = bp->blk_cptr, A.rec = bp->rec, A.rec2 = bp->rec2;
  const size_t smem = block_smem(bp);
  static const int t_env = getenv("FEMB_ASM_BLOCK_THREADS") ? atoi(getenv("FEMB_ASM_BLOCK_THREADS")) : 256;
#define FEMB_LAUNCH_BLOCKS(T)                                                                                          \
  {                                                                                                                    \
    FEMB_CUDA(cudaFuncSetAttribute(assemble_p1_blocks<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));    \
    assemble_p1_blocks<T><<<(unsigned)bp->nblocks, T, smem, s>>>(A, coords, vals, flag);                               \
  }
  if (t_env == 128) FEMB_LAUNCH_BLOCKS(128) else if (t_env == 512) FEMB_LAUNCH_BLOCKS(512) else FEMB_LAUNCH_BLOCKS(256)
#undef FEMB_LAUNCH_BLOCKS
  FEMB_LAUNCH_CHECK();
  return FEMB_OK;
}

}  // namespace femb
