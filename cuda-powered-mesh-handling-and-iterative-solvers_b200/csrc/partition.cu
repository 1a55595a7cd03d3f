// Element-graph partition by farthest-seed region growing (reference: subdivision.ipynb cells 8-9).
//
// The notebook builds the element face-adjacency matrix as a torch sparse COO tensor and runs every BFS level as a
// torch.sparse.mm with a dense [n_parts, M] frontier (pick_distant_seeds: one full BFS per seed; region_growing_partition: one
// SpMM per level), reading `.item()` / `.any()` back every level.  Here the adjacency is a CSR built by one radix sort of the
// directed edge keys, and a level is one kernel over the still-unlabelled elements (each looks at its <= 4 face neighbours);
// the host reads one "changed" flag per level.  Tie rule when several regions reach an element in the same level: the
// highest part index wins -- what the reference's `labels[idx[:,1]] = idx[:,0]` does on the CPU, where duplicate indices
// are written in row-major (part-major) order; on CUDA the reference's result is unspecified for such elements.
#include <cub/cub.cuh>

#include "common.cuh"

namespace femb {

__global__ void pairs_to_keys(const long long* __restrict__ pairs, long long S, int pair_stride, unsigned long long* __restrict__ keys) {
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < S; t += (long long)gridDim.x * blockDim.x) {
    const unsigned long long a = (unsigned long long)pairs[t * pair_stride], b = (unsigned long long)pairs[t * pair_stride + pair_stride / 2];
    keys[2 * t] = (a << 32) | b;
    keys[2 * t + 1] = (b << 32) | a;
  }
}

__global__ void keys_to_csr(const unsigned long long* __restrict__ keys, long long L, long long M, int* __restrict__ crow, int* __restrict__ col) {
  for (long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x; k < L; k += (long long)gridDim.x * blockDim.x) {
    const long long cur = (long long)(keys[k] >> 32), prev = k == 0 ? -1 : (long long)(keys[k - 1] >> 32);
    col[k] = (int)(keys[k] & 0xffffffffull);
    for (long long n = prev + 1; n <= cur; ++n) crow[n] = (int)k;
    if (k == L - 1)
      for (long long n = cur + 1; n <= M; ++n) crow[n] = (int)L;
  }
}

__global__ void fill_int(int* __restrict__ p, long long n, int v) {
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) p[t] = v;
}

// one BFS level: an unreached vertex adjacent to a vertex reached at `level - 1` gets dist = level (and, when labels are
// propagated, the largest label among those neighbours)
__global__ void bfs_level_kernel(const int* __restrict__ crow, const int* __restrict__ col, long long M, int level, int* __restrict__ dist,
                                 int* __restrict__ label, int* __restrict__ changed) {
  bool any = false;
  for (long long v = blockIdx.x * (long long)blockDim.x + threadIdx.x; v < M; v += (long long)gridDim.x * blockDim.x) {
    if (dist[v] >= 0) continue;
    int best = -1;
    bool hit = false;
    for (int k = crow[v]; k < crow[v + 1]; ++k) {
      const int w = col[k];
      if (dist[w] == level - 1) {
        hit = true;
        if (label) best = max(best, label[w]);
      }
    }
    if (hit) {
      dist[v] = level;  // neighbours test dist == level - 1, so writing `level` during the sweep cannot be picked up this level
      if (label) label[v] = best;
      any = true;
    }
  }
  if (any) *changed = 1;
}

__global__ void set_sources(const long long* __restrict__ src, int n, int* __restrict__ dist, int* __restrict__ label) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n) {
    dist[src[t]] = 0;
    if (label) label[src[t]] = t;
  }
}

}  // namespace femb

using namespace femb;

extern "C" int femb_graph_from_pairs(const int64_t* pairs, int64_t S, int pair_stride, int64_t M, int32_t* crow, int32_t* col,
                                     femb_stream stream) {
  FEMB_CHECK_ARG(S >= 0 && M >= 1 && M < (1ll << 31) && 2 * S < (1ll << 31) && crow && (col || S == 0) && (pair_stride == 2 || pair_stride == 4),
                 "S >= 0, 1 <= M < 2^31, 2S < 2^31, pair_stride in {2,4}");
  cudaStream_t s = as_stream(stream);
  const long long L = 2 * S;
  if (L == 0) {
    fill_int<<<grid_for(M + 1, 256), 256, 0, s>>>(crow, M + 1, 0);
    FEMB_LAUNCH_CHECK();
    return FEMB_OK;
  }
  Scratch scr(s);
  unsigned long long *k0, *k1;
  FEMB_CUDA(scr.alloc(&k0, (size_t)L));
  FEMB_CUDA(scr.alloc(&k1, (size_t)L));
  pairs_to_keys<<<grid_for(S, 256), 256, 0, s>>>(reinterpret_cast<const long long*>(pairs), S, pair_stride, k0);
  size_t tmp_bytes = 0;
  int bits = 32;
  while (bits < 64 && (M >> (bits - 32)) != 0) ++bits;
  cub::DeviceRadixSort::SortKeys(nullptr, tmp_bytes, k0, k1, (int)L, 0, bits, s);
  unsigned char* tmp;
  FEMB_CUDA(scr.alloc(&tmp, tmp_bytes));
  FEMB_CUDA(cub::DeviceRadixSort::SortKeys(tmp, tmp_bytes, k0, k1, (int)L, 0, bits, s));
  keys_to_csr<<<grid_for(L, 256), 256, 0, s>>>(k1, L, M, crow, col);
  FEMB_LAUNCH_CHECK();
  FEMB_CUDA(cudaStreamSynchronize(s));  // scratch keys are released when this returns
  return FEMB_OK;
}

extern "C" int femb_graph_bfs(const int32_t* crow, const int32_t* col, int64_t M, const int64_t* sources, int n_sources, int32_t* dist,
                              int32_t* label, int32_t* levels_host, femb_stream stream) {
  FEMB_CHECK_ARG(crow && col && M >= 1 && sources && n_sources >= 1 && dist, "crow/col/sources/dist, M >= 1, n_sources >= 1");
  cudaStream_t s = as_stream(stream);
  Scratch scr(s);
  int* changed;
  FEMB_CUDA(scr.alloc(&changed, 1));
  fill_int<<<grid_for(M, 256), 256, 0, s>>>(dist, M, -1);
  if (label) fill_int<<<grid_for(M, 256), 256, 0, s>>>(label, M, -1);
  set_sources<<<(n_sources + 127) / 128, 128, 0, s>>>(reinterpret_cast<const long long*>(sources), n_sources, dist, label);
  FEMB_LAUNCH_CHECK();
  int level = 1, h = 1;
  while (h) {
    FEMB_CUDA(cudaMemsetAsync(changed, 0, sizeof(int), s));
    bfs_level_kernel<<<grid_for(M, 256), 256, 0, s>>>(crow, col, M, level, dist, label, changed);
    FEMB_CUDA(cudaMemcpyAsync(&h, changed, sizeof(int), cudaMemcpyDeviceToHost, s));
    FEMB_CUDA(cudaStreamSynchronize(s));
    if (h) ++level;
  }
  if (levels_host) *levels_host = level - 1;
  return FEMB_OK;
}
