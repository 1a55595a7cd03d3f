// P1 -> P2 mid-edge insertion (reference c3d4_to_c3d10, solver/element.py:777-833) as device kernels.
//
// The reference walks the elements in a Python loop with a dict: edge (min,max) -> new node id, ids handed out in
// first-encounter order over (element-major, edge slots (0,1),(1,2),(2,0),(0,3),(1,3),(2,3)).  Here:
//   1. every (element, slot) occurrence k = 6e + slot gets the key (lo << 32 | hi); ONE stable radix sort of (key, k)
//   2. in the sorted stream equal keys are adjacent and -- the sort being stable -- the first of a group is the edge's first
//      occurrence; a max-scan of the group-head positions gives every occurrence its group's head
//   3. flags at the first occurrences, in occurrence order, are prefix-summed: that rank IS the reference's numbering
//   4. fill: connectivity [M,10] int32, mid-point coordinates (x_lo + x_hi) / 2, and the edge list in new-id order
// Bit-exact against the reference's numbering (tests/golden/tets.npz).
#include <cub/cub.cuh>

#include "common.cuh"

struct femb_p2_plan {
  long long M = 0, N = 0, E = 0;
  int* mid = nullptr;    // [6M] new node id of every (element, slot) occurrence
  int* edges = nullptr;  // [E,2] (lo, hi) of the edge behind new node N + r
};

namespace femb {
namespace {

__constant__ int c_slots[6][2] = {{0, 1}, {1, 2}, {2, 0}, {0, 3}, {1, 3}, {2, 3}};

template <typename I>
__global__ void edge_keys(const I* __restrict__ conn, long long M, unsigned long long* __restrict__ keys, int* __restrict__ occ) {
  for (long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x; k < 6 * M; k += (long long)gridDim.x * blockDim.x) {
    const long long e = k / 6;
    const int s = (int)(k - 6 * e);
    const long long a = ldidx(conn + 4 * e + c_slots[s][0]), b = ldidx(conn + 4 * e + c_slots[s][1]);
    const unsigned long long lo = (unsigned long long)min(a, b), hi = (unsigned long long)max(a, b);
    keys[k] = lo << 32 | hi;
    occ[k] = (int)k;
  }
}

// head position of every sorted entry (its own position if it starts a group, 0 otherwise -> inclusive max-scan), and a flag
// at the first occurrence of every edge, indexed by occurrence
__global__ void group_heads(const unsigned long long* __restrict__ keys, const int* __restrict__ occ, long long L, int* __restrict__ headpos,
                            int* __restrict__ first_flag) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < L; i += (long long)gridDim.x * blockDim.x) {
    const bool head = i == 0 || keys[i] != keys[i - 1];
    headpos[i] = head ? (int)i : 0;
    if (head) first_flag[occ[i]] = 1;
  }
}

__global__ void assign_ids(const unsigned long long* __restrict__ keys, const int* __restrict__ occ, const int* __restrict__ headpos,
                           const int* __restrict__ rank, long long L, long long N, int* __restrict__ mid, int* __restrict__ edges) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < L; i += (long long)gridDim.x * blockDim.x) {
    const int h = headpos[i];
    const int r = rank[occ[h]];  // first-encounter rank of the edge
    mid[occ[i]] = (int)(N + r);
    if (h == i) edges[2ll * r] = (int)(keys[i] >> 32), edges[2ll * r + 1] = (int)(keys[i] & 0xffffffffull);
  }
}

template <typename I>
__global__ void fill_conn(const I* __restrict__ conn, const int* __restrict__ mid, long long M, int* __restrict__ out) {
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < 10 * M; t += (long long)gridDim.x * blockDim.x) {
    const long long e = t / 10;
    const int a = (int)(t - 10 * e);
    out[t] = a < 4 ? (int)ldidx(conn + 4 * e + a) : mid[6 * e + a - 4];
  }
}

template <typename T, typename O>
__global__ void fill_coords(const T* __restrict__ x, const int* __restrict__ edges, long long N, long long E, O* __restrict__ out) {
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < 3 * (N + E); t += (long long)gridDim.x * blockDim.x) {
    const long long n = t / 3;
    const int c = (int)(t - 3 * n);
    if (n < N) {
      out[t] = (O)x[t];
    } else {
      const long long r = n - N;
      out[t] = (O)((x[3ll * edges[2 * r] + c] + x[3ll * edges[2 * r + 1] + c]) / 2);  // (cA + cB) / 2 as the reference forms it
    }
  }
}

struct MaxOp {
  __device__ __forceinline__ int operator()(int a, int b) const { return a > b ? a : b; }
};

template <typename I>
int p2_build(femb_p2_plan* p, const I* conn, cudaStream_t s) {
  const long long M = p->M, L = 6 * M, N = p->N;
  Scratch scr(s);
  unsigned long long *k0, *k1;
  int *o0, *o1, *headpos, *flag, *rank;
  FEMB_CUDA(scr.alloc(&k0, L));
  FEMB_CUDA(scr.alloc(&k1, L));
  FEMB_CUDA(scr.alloc(&o0, L));
  FEMB_CUDA(scr.alloc(&o1, L));
  FEMB_CUDA(scr.alloc(&headpos, L));
  FEMB_CUDA(scr.alloc(&flag, L + 1));
  FEMB_CUDA(scr.alloc(&rank, L + 1));
  FEMB_CUDA(cudaMemsetAsync(flag, 0, sizeof(int) * (L + 1), s));
  edge_keys<I><<<grid_for(L, 256), 256, 0, s>>>(conn, M, k0, o0);
  FEMB_LAUNCH_CHECK();
  int bits = 1;
  while ((1ll << bits) < N) ++bits;
  size_t tb = 0, tb2 = 0, tb3 = 0;
  FEMB_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tb, k0, k1, o0, o1, (int)L, 0, 32 + bits, s));
  FEMB_CUDA(cub::DeviceScan::InclusiveScan(nullptr, tb2, headpos, headpos, MaxOp(), (int)L, s));
  FEMB_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tb3, flag, rank, (int)(L + 1), s));
  void* tmp;
  FEMB_CUDA(scr.alloc((char**)&tmp, std::max(tb, std::max(tb2, tb3))));
  FEMB_CUDA(cub::DeviceRadixSort::SortPairs(tmp, tb, k0, k1, o0, o1, (int)L, 0, 32 + bits, s));
  group_heads<<<grid_for(L, 256), 256, 0, s>>>(k1, o1, L, headpos, flag);
  FEMB_LAUNCH_CHECK();
  FEMB_CUDA(cub::DeviceScan::InclusiveScan(tmp, tb2, headpos, headpos, MaxOp(), (int)L, s));
  FEMB_CUDA(cub::DeviceScan::ExclusiveSum(tmp, tb3, flag, rank, (int)(L + 1), s));
  int E = 0;
  FEMB_CUDA(cudaMemcpyAsync(&E, rank + L, sizeof(int), cudaMemcpyDeviceToHost, s));
  FEMB_CUDA(cudaStreamSynchronize(s));
  p->E = E;
  FEMB_CHECK_ARG(N + (long long)E < (1ll << 31) - 1, "N + edges must fit int32 (the reference returns int32 connectivity)");
  FEMB_CUDA(cudaMalloc(&p->mid, sizeof(int) * std::max<long long>(L, 1)));
  FEMB_CUDA(cudaMalloc(&p->edges, sizeof(int) * 2 * std::max<long long>(E, 1)));
  assign_ids<<<grid_for(L, 256), 256, 0, s>>>(k1, o1, headpos, rank, L, N, p->mid, p->edges);
  FEMB_LAUNCH_CHECK();
  FEMB_CUDA(cudaStreamSynchronize(s));
  return FEMB_OK;
}

void p2_free(femb_p2_plan* p) {
  if (!p) return;
  cudaFree(p->mid);
  cudaFree(p->edges);
  delete p;
}

}  // namespace
}  // namespace femb

using namespace femb;

extern "C" int femb_p2_create(const void* conn, int ib, int64_t M, int64_t n_nodes, femb_stream stream, femb_p2_plan** plan, int64_t* n_edges) {
  FEMB_CHECK_ARG((ib == 4 || ib == 8) && plan && M >= 0 && n_nodes >= 1, "ib in {4,8}, plan, M >= 0, n_nodes >= 1");
  FEMB_CHECK_ARG(6 * (long long)M < (1ll << 31) - 2 && n_nodes < (1ll << 31) - 1, "6*M and n_nodes must fit int32");
  auto* p = new femb_p2_plan();
  p->M = M, p->N = n_nodes;
  int rc = FEMB_OK;
  if (M > 0) rc = ib == 8 ? p2_build<long long>(p, (const long long*)conn, as_stream(stream)) : p2_build<int>(p, (const int*)conn, as_stream(stream));
  if (rc != FEMB_OK) {
    p2_free(p);
    return rc;
  }
  *plan = p;
  if (n_edges) *n_edges = p->E;
  return FEMB_OK;
}

extern "C" int femb_p2_fill(femb_p2_plan* p, const void* conn, int ib, const void* coords, int fp_in, int fp_out, int32_t* conn10, void* coords_out,
                            int32_t* edges, femb_stream stream) {
  FEMB_CHECK_ARG(p && (ib == 4 || ib == 8) && (fp_in == 4 || fp_in == 8) && (fp_out == 4 || fp_out == 8), "plan, ib / fp in {4,8}");
  cudaStream_t s = as_stream(stream);
  if (conn10 && p->M > 0) {
    if (ib == 8) fill_conn<long long><<<grid_for(10 * p->M, 256), 256, 0, s>>>((const long long*)conn, p->mid, p->M, conn10);
    else fill_conn<int><<<grid_for(10 * p->M, 256), 256, 0, s>>>((const int*)conn, p->mid, p->M, conn10);
  }
  if (coords_out) {
    const int g = grid_for(3 * (p->N + p->E), 256);
#define FC(TI, TO) fill_coords<TI, TO><<<g, 256, 0, s>>>((const TI*)coords, p->edges, p->N, p->E, (TO*)coords_out)
    if (fp_in == 8 && fp_out == 8) FC(double, double);
    else if (fp_in == 8) FC(double, float);
    else if (fp_out == 8) FC(float, double);
    else FC(float, float);
#undef FC
  }
  if (edges && p->E > 0) FEMB_CUDA(cudaMemcpyAsync(edges, p->edges, sizeof(int) * 2 * p->E, cudaMemcpyDeviceToDevice, s));
  FEMB_LAUNCH_CHECK();
  return FEMB_OK;
}

extern "C" int femb_p2_destroy(femb_p2_plan* p) {
  p2_free(p);
  return FEMB_OK;
}
