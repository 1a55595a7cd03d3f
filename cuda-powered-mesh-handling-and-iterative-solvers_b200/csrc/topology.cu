// Face / edge connectivity and surface extraction (bit-exact contracts of the reference):
//   compute_*_surface_faces_with_*  element.py:543-579, 1293-1334, 2234-2283; shell.py:261-295, 561-597
//   identify_*_shared_faces/edges   element.py:707-762, 1474-1532; shell.py:205-259, 504-559
// The reference groups canonical (ascending) node tuples with torch.unique(dim=0).  Here every entity gets its tuple
// packed into 64-bit words and the entity ids are ordered with stable LSD radix-sort passes (cub), least significant
// word first.  Run lengths in the sorted stream give "appears once" (surface) and "appears twice" (shared); outputs are
// produced with flag + exclusive-scan compaction, so their order is exactly the reference's.
#include <cub/cub.cuh>

#include "common.cuh"

namespace femb {

struct EntTable {
  int nf, nfn;
  int nodes[6][4];
  int surf_slot[6];
  int extra[6];
};

static const EntTable kTables[6] = {
    // tets: shared-face numbering (element.py:722-727); surface routine concatenates (012)(013)(023)(123) (:557-569)
    {4, 3, {{0, 1, 2, 0}, {0, 1, 3, 0}, {1, 2, 3, 0}, {0, 2, 3, 0}}, {0, 1, 3, 2}, {3, 2, 0, 1}},
    // hexes (element.py:1308-1324, 1489-1496)
    {6, 4, {{0, 1, 5, 4}, {1, 2, 6, 5}, {2, 3, 7, 6}, {0, 4, 7, 3}, {0, 3, 2, 1}, {4, 5, 6, 7}}, {0, 1, 2, 3, 4, 5}, {2, 0, 0, 1, 4, 0}},
    // wedge quads / tris (element.py:2248-2266)
    {3, 4, {{0, 1, 4, 3}, {1, 2, 5, 4}, {2, 0, 3, 5}}, {0, 1, 2}, {2, 0, 1}},
    {2, 3, {{0, 2, 1, 0}, {3, 4, 5, 0}}, {0, 1}, {3, 0}},
    // triangle edges (shell.py:220-224, 275-285), quad edges (shell.py:519-524, 575-587)
    {3, 2, {{0, 1, 0, 0}, {1, 2, 0, 0}, {2, 0, 0, 0}}, {0, 1, 2}, {2, 0, 1}},
    {4, 2, {{0, 1, 0, 0}, {1, 2, 0, 0}, {2, 3, 0, 0}, {3, 0, 0, 0}}, {0, 1, 2, 3}, {3, 0, 1, 2}},
};

}  // namespace femb

struct femb_entity_plan {
  femb::EntTable tab;
  const void* conn = nullptr;  // caller keeps it alive until destroy
  int ib = 8, stride = 0;
  long long M = 0, T = 0, K = 0, S = 0;
  int* order = nullptr;   // [T] entity ids (e*nf+f) in lexicographic order of their canonical tuples
  int* sscan = nullptr;   // [T+1] exclusive scan of "appears once" flags in the surface routine's slot-major order
  int* pscan = nullptr;   // [T+1] exclusive scan of "first of a pair" flags in sorted order
};

namespace femb {

template <typename I>
__device__ __forceinline__ void canon_tuple(const I* __restrict__ conn, int stride, const EntTable& tab, int ent, int* t) {
  const int e = ent / tab.nf, f = ent - e * tab.nf;
  for (int k = 0; k < tab.nfn; ++k) t[k] = (int)ldidx(conn + (long long)e * stride + tab.nodes[f][k]);
  // insertion sort, nfn <= 4
  for (int a = 1; a < tab.nfn; ++a) {
    const int v = t[a];
    int b = a - 1;
    while (b >= 0 && t[b] > v) {
      t[b + 1] = t[b];
      --b;
    }
    t[b + 1] = v;
  }
}

// key of one LSD pass: components [c0, c1) of the canonical tuple, most significant first
template <typename I>
__global__ void pack_keys(const I* __restrict__ conn, int stride, EntTable tab, const int* __restrict__ perm, long long T, int c0, int c1,
                          int bits, unsigned long long* __restrict__ keys, int* __restrict__ iota) {
  for (long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x; k < T; k += (long long)gridDim.x * blockDim.x) {
    const int ent = perm ? perm[k] : (int)k;
    int t[4];
    canon_tuple(conn, stride, tab, ent, t);
    unsigned long long key = 0;
    for (int c = c0; c < c1; ++c) key = (key << bits) | (unsigned long long)(unsigned)t[c];
    keys[k] = key;
    if (iota) iota[k] = (int)k;
  }
}

template <typename I>
__global__ void max_node_kernel(const I* __restrict__ conn, long long total, int* __restrict__ out) {
  int m = 0;
  for (long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x; k < total; k += (long long)gridDim.x * blockDim.x) m = max(m, (int)ldidx(conn + k));
  for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) atomicMax(out, m);
}

template <typename I>
__global__ void classify_runs(const I* __restrict__ conn, int stride, EntTable tab, const int* __restrict__ order, long long T, long long M,
                              int* __restrict__ sflag, int* __restrict__ pflag) {
  for (long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x; k < T; k += (long long)gridDim.x * blockDim.x) {
    int a[4], b[4];
    canon_tuple(conn, stride, tab, order[k], a);
    auto same = [&](long long j) {
      if (j < 0 || j >= T) return false;
      canon_tuple(conn, stride, tab, order[j], b);
      for (int c = 0; c < tab.nfn; ++c)
        if (a[c] != b[c]) return false;
      return true;
    };
    const bool head = !same(k - 1);
    const bool n1 = same(k + 1);
    const bool n2 = n1 && same(k + 2);
    const int ent = order[k], e = ent / tab.nf, f = ent - e * tab.nf;
    sflag[(long long)tab.surf_slot[f] * M + e] = (head && !n1) ? 1 : 0;
    pflag[k] = (head && n1 && !n2) ? 1 : 0;
  }
}

template <typename I>
__global__ void write_surface(const I* __restrict__ conn, int stride, EntTable tab, const int* __restrict__ sscan, long long T, long long M,
                              long long* __restrict__ faces, long long* __restrict__ extra) {
  for (long long pos = blockIdx.x * (long long)blockDim.x + threadIdx.x; pos < T; pos += (long long)gridDim.x * blockDim.x) {
    const int o = sscan[pos];
    if (sscan[pos + 1] == o) continue;
    const int slot = (int)(pos / M);
    const long long e = pos - (long long)slot * M;
    int f = 0;
    for (int q = 0; q < tab.nf; ++q)
      if (tab.surf_slot[q] == slot) f = q;
    for (int c = 0; c < tab.nfn; ++c) faces[(long long)o * tab.nfn + c] = ldidx(conn + e * stride + tab.nodes[f][c]);
    extra[o] = ldidx(conn + e * stride + tab.extra[f]);
  }
}

__global__ void write_shared(EntTable tab, const int* __restrict__ order, const int* __restrict__ pscan, long long T, long long* __restrict__ pairs) {
  for (long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x; k < T; k += (long long)gridDim.x * blockDim.x) {
    const int o = pscan[k];
    if (pscan[k + 1] == o) continue;
    const int a = order[k], b = order[k + 1];
    long long* p = pairs + 4ll * o;
    p[0] = a / tab.nf, p[1] = a % tab.nf, p[2] = b / tab.nf, p[3] = b % tab.nf;
  }
}

template <typename I>
static int entities_build(femb_entity_plan* p, cudaStream_t s) {
  const I* conn = static_cast<const I*>(p->conn);
  const long long T = p->T, M = p->M;
  const EntTable tab = p->tab;
  Scratch scr(s);
  int* dmax;
  FEMB_CUDA(scr.alloc(&dmax, 1));
  FEMB_CUDA(cudaMemsetAsync(dmax, 0, sizeof(int), s));
  max_node_kernel<I><<<grid_for(M * p->stride, 256), 256, 0, s>>>(conn, M * p->stride, dmax);
  int hmax = 0;
  FEMB_CUDA(cudaMemcpyAsync(&hmax, dmax, sizeof(int), cudaMemcpyDeviceToHost, s));
  FEMB_CUDA(cudaStreamSynchronize(s));
  int bits = 1;
  while ((1ll << bits) <= hmax) ++bits;
  const int per_pass = std::max(1, std::min(tab.nfn, 64 / bits));
  unsigned long long *keys, *keys_out;
  int *perm_a, *perm_b, *sflag, *pflag;
  FEMB_CUDA(scr.alloc(&keys, T));
  FEMB_CUDA(scr.alloc(&keys_out, T));
  FEMB_CUDA(scr.alloc(&perm_a, T));
  FEMB_CUDA(scr.alloc(&perm_b, T));
  size_t tb = 0;
  FEMB_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tb, keys, keys_out, perm_a, perm_b, (int)T, 0, 64, s));
  void* tmp;
  FEMB_CUDA(scr.alloc((char**)&tmp, tb));
  int* cur = nullptr;  // current permutation (nullptr = identity)
  int* nxt = perm_b;
  int* src = perm_a;
  for (int c1 = tab.nfn; c1 > 0; c1 -= per_pass) {
    const int c0 = std::max(0, c1 - per_pass);
    // keys of this pass in the current order; `src` receives the current permutation as sort payload
    pack_keys<I><<<grid_for(T, 256), 256, 0, s>>>(conn, p->stride, tab, cur, T, c0, c1, bits, keys, cur ? nullptr : src);
    FEMB_LAUNCH_CHECK();
    const int* payload = cur ? cur : src;
    int* dst = (payload == perm_a) ? perm_b : perm_a;
    FEMB_CUDA(cub::DeviceRadixSort::SortPairs(tmp, tb, keys, keys_out, payload, dst, (int)T, 0, bits * (c1 - c0), s));
    cur = dst;
    (void)nxt;
  }
  FEMB_CUDA(cudaMemcpyAsync(p->order, cur, sizeof(int) * T, cudaMemcpyDeviceToDevice, s));
  FEMB_CUDA(scr.alloc(&sflag, T + 1));
  FEMB_CUDA(scr.alloc(&pflag, T + 1));
  FEMB_CUDA(cudaMemsetAsync(sflag + T, 0, sizeof(int), s));
  FEMB_CUDA(cudaMemsetAsync(pflag + T, 0, sizeof(int), s));
  classify_runs<I><<<grid_for(T, 256), 256, 0, s>>>(conn, p->stride, tab, p->order, T, M, sflag, pflag);
  FEMB_LAUNCH_CHECK();
  size_t tb2 = 0;
  FEMB_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tb2, sflag, p->sscan, (int)(T + 1), s));
  void* tmp2;
  FEMB_CUDA(scr.alloc((char**)&tmp2, tb2));
  FEMB_CUDA(cub::DeviceScan::ExclusiveSum(tmp2, tb2, sflag, p->sscan, (int)(T + 1), s));
  FEMB_CUDA(cub::DeviceScan::ExclusiveSum(tmp2, tb2, pflag, p->pscan, (int)(T + 1), s));
  int hk = 0, hs = 0;
  FEMB_CUDA(cudaMemcpyAsync(&hk, p->sscan + T, sizeof(int), cudaMemcpyDeviceToHost, s));
  FEMB_CUDA(cudaMemcpyAsync(&hs, p->pscan + T, sizeof(int), cudaMemcpyDeviceToHost, s));
  FEMB_CUDA(cudaStreamSynchronize(s));
  p->K = hk;
  p->S = hs;
  return FEMB_OK;
}

template <typename T>
__global__ void surface_normals_kernel(const T* __restrict__ coords, const long long* __restrict__ faces, const long long* __restrict__ extra,
                                       long long K, int nfn, int second, T* __restrict__ out) {
  for (long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x; k < K; k += (long long)gridDim.x * blockDim.x) {
    T p[4][3], ctr[3] = {0, 0, 0};
    for (int a = 0; a < nfn; ++a)
      for (int c = 0; c < 3; ++c) {
        p[a][c] = coords[3 * faces[k * nfn + a] + c];
        ctr[c] += p[a][c];
      }
    T v1[3], v2[3];
    for (int c = 0; c < 3; ++c) v1[c] = p[1][c] - p[0][c], v2[c] = p[second][c] - p[0][c];
    T n[3] = {v1[1] * v2[2] - v1[2] * v2[1], v1[2] * v2[0] - v1[0] * v2[2], v1[0] * v2[1] - v1[1] * v2[0]};
    const T nn = sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
    T to[3];
    for (int c = 0; c < 3; ++c) {
      n[c] /= nn;
      to[c] = coords[3 * extra[k] + c] - ctr[c] / (T)nfn;
    }
    const T tn = sqrt(to[0] * to[0] + to[1] * to[1] + to[2] * to[2]);
    const T dot = n[0] * (to[0] / tn) + n[1] * (to[1] / tn) + n[2] * (to[2] / tn);
    const T sgn = dot > 0 ? T(-1) : T(1);
    for (int c = 0; c < 3; ++c) out[3 * k + c] = sgn * n[c];
  }
}

// per-element area-weighted face normals: tets cross(e01,e02)/2 (element.py:652-705), hexes cross(e01,e03) (:1418-1472)
template <typename T, typename I>
__global__ void face_normals_kernel(const T* __restrict__ coords, const I* __restrict__ conn, long long M, int stride, EntTable tab, int hex,
                                    T* __restrict__ out) {
  const int extra_elem_hex[6] = {2, 0, 0, 1, 6, 0};  // element.py:1463 (differs from the surface routine's table)
  const long long total = M * tab.nf;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const long long e = t / tab.nf;
    const int f = (int)(t - e * tab.nf);
    T p[4][3], ctr[3] = {0, 0, 0};
    for (int a = 0; a < tab.nfn; ++a) {
      const long long n = ldidx(conn + e * stride + tab.nodes[f][a]);
      for (int c = 0; c < 3; ++c) {
        p[a][c] = coords[3 * n + c];
        ctr[c] += p[a][c];
      }
    }
    const int second = hex ? 3 : 2;
    T v1[3], v2[3];
    for (int c = 0; c < 3; ++c) v1[c] = p[1][c] - p[0][c], v2[c] = p[second][c] - p[0][c];
    const T sc = hex ? T(1) : T(0.5);
    T n[3] = {(v1[1] * v2[2] - v1[2] * v2[1]) * sc, (v1[2] * v2[0] - v1[0] * v2[2]) * sc, (v1[0] * v2[1] - v1[1] * v2[0]) * sc};
    const long long ne = ldidx(conn + e * stride + (hex ? extra_elem_hex[f] : tab.extra[f]));
    T dot = 0;
    for (int c = 0; c < 3; ++c) dot += n[c] * (coords[3 * ne + c] - ctr[c] / (T)tab.nfn);
    const T sgn = dot > 0 ? T(-1) : T(1);
    for (int c = 0; c < 3; ++c) out[3 * t + c] = sgn * n[c];
  }
}

static void entities_free(femb_entity_plan* p) {
  if (!p) return;
  cudaFree(p->order);
  cudaFree(p->sscan);
  cudaFree(p->pscan);
  delete p;
}

}  // namespace femb

using namespace femb;

extern "C" int femb_entities_create(int ent_kind, const void* conn, int ib, int64_t M, int conn_stride, femb_stream stream,
                                    femb_entity_plan** plan, int64_t* n_surface, int64_t* n_shared) {
  FEMB_CHECK_ARG(ent_kind >= 0 && ent_kind <= 5, "ent_kind in 0..5");
  FEMB_CHECK_ARG((ib == 4 || ib == 8) && plan && M >= 0 && conn_stride >= 2, "ib / plan / M / conn_stride");
  auto* p = new femb_entity_plan();
  p->tab = kTables[ent_kind];
  p->conn = conn, p->ib = ib, p->stride = conn_stride, p->M = M, p->T = M * p->tab.nf;
  if (p->T >= (1ll << 31) - 2) {
    delete p;
    set_error("femb_entities_create: M * entities_per_element must fit int32");
    return FEMB_ERR_ARG;
  }
  int rc = FEMB_OK;
  if (M > 0) {
    if (cudaMalloc(&p->order, sizeof(int) * p->T) != cudaSuccess || cudaMalloc(&p->sscan, sizeof(int) * (p->T + 1)) != cudaSuccess ||
        cudaMalloc(&p->pscan, sizeof(int) * (p->T + 1)) != cudaSuccess) {
      entities_free(p);
      set_error("femb_entities_create: out of device memory");
      return FEMB_ERR_CUDA;
    }
    rc = ib == 8 ? entities_build<long long>(p, as_stream(stream)) : entities_build<int>(p, as_stream(stream));
  }
  if (rc != FEMB_OK) {
    entities_free(p);
    return rc;
  }
  *plan = p;
  if (n_surface) *n_surface = p->K;
  if (n_shared) *n_shared = p->S;
  return FEMB_OK;
}

extern "C" int femb_entities_surface(femb_entity_plan* p, int64_t* faces, int64_t* extra, femb_stream stream) {
  FEMB_CHECK_ARG(p != nullptr, "plan");
  if (p->T == 0 || p->K == 0) return FEMB_OK;
  cudaStream_t s = as_stream(stream);
  const int grid = grid_for(p->T, 256);
  if (p->ib == 8) write_surface<long long><<<grid, 256, 0, s>>>((const long long*)p->conn, p->stride, p->tab, p->sscan, p->T, p->M, (long long*)faces, (long long*)extra);
  else write_surface<int><<<grid, 256, 0, s>>>((const int*)p->conn, p->stride, p->tab, p->sscan, p->T, p->M, (long long*)faces, (long long*)extra);
  FEMB_LAUNCH_CHECK();
  return FEMB_OK;
}

extern "C" int femb_entities_shared(femb_entity_plan* p, int64_t* pairs, femb_stream stream) {
  FEMB_CHECK_ARG(p != nullptr, "plan");
  if (p->T == 0 || p->S == 0) return FEMB_OK;
  write_shared<<<grid_for(p->T, 256), 256, 0, as_stream(stream)>>>(p->tab, p->order, p->pscan, p->T, (long long*)pairs);
  FEMB_LAUNCH_CHECK();
  return FEMB_OK;
}

extern "C" int femb_entities_destroy(femb_entity_plan* p) {
  entities_free(p);
  return FEMB_OK;
}

extern "C" int femb_surface_normals(const void* coords, int fp, const int64_t* faces, const int64_t* extra, int64_t K, int nfn, int second,
                                    void* normals, femb_stream stream) {
  FEMB_CHECK_ARG((fp == 4 || fp == 8) && nfn >= 3 && nfn <= 4 && second >= 1 && second < nfn, "fp / nfn / second");
  if (K == 0) return FEMB_OK;
  cudaStream_t s = as_stream(stream);
  const int grid = grid_for(K, 128);
  if (fp == 8) surface_normals_kernel<double><<<grid, 128, 0, s>>>((const double*)coords, (const long long*)faces, (const long long*)extra, K, nfn, second, (double*)normals);
  else surface_normals_kernel<float><<<grid, 128, 0, s>>>((const float*)coords, (const long long*)faces, (const long long*)extra, K, nfn, second, (float*)normals);
  FEMB_LAUNCH_CHECK();
  return FEMB_OK;
}

extern "C" int femb_face_normals_area(int kind, const void* coords, int fp, const void* conn, int ib, int64_t M, int conn_stride, void* out,
                                      femb_stream stream) {
  FEMB_CHECK_ARG((fp == 4 || fp == 8) && (ib == 4 || ib == 8), "fp in {4,8}, ib in {4,8}");
  FEMB_CHECK_ARG(kind == FEMB_C3D4 || kind == FEMB_C3D8, "kind must be C3D4 or C3D8");
  if (M == 0) return FEMB_OK;
  const int hex = kind == FEMB_C3D8;
  const EntTable tab = kTables[hex ? 1 : 0];
  cudaStream_t s = as_stream(stream);
  const int grid = grid_for(M * tab.nf, 128);
#define FN(T, I) face_normals_kernel<T, I><<<grid, 128, 0, s>>>((const T*)coords, (const I*)conn, M, conn_stride, tab, hex, (T*)out)
  if (fp == 8) {
    if (ib == 8) FN(double, long long);
    else FN(double, int);
  } else {
    if (ib == 8) FN(float, long long);
    else FN(float, int);
  }
#undef FN
  FEMB_LAUNCH_CHECK();
  return FEMB_OK;
}
