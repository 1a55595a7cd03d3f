// Face / edge connectivity and surface extraction (bit-exact contracts of the reference):
//   compute_*_surface_faces_with_*  element.py:543-579, 1293-1334, 2234-2283; shell.py:261-295, 561-597
//   identify_*_shared_faces/edges   element.py:707-762, 1474-1532; shell.py:205-259, 504-559
// The reference groups canonical (ascending) node tuples with torch.unique(dim=0).  Here every entity gets its tuple
// packed into 64-bit words and the entity ids are ordered with stable LSD radix-sort passes (cub), least significant
// word first.  Run lengths in the sorted stream give "appears once" (surface) and "appears twice" (shared); outputs are
// produced with flag + exclusive-scan compaction, so their order is exactly the reference's.
//
// Default path (entities_build_buckets): nine radix passes over 256 M (key, id) pairs cost 55 ms on the 64 M-tet mesh, 60x
// the bytes the answer needs.  Instead every entity is dropped into the bucket of its SMALLEST node (one counting pass, one
// scatter pass: ~24 faces per bucket on a tet mesh), and duplicates are matched inside the bucket by one warp: a bitonic
// sort of the warp by the rest of the tuple (buckets above 32 entries: every lane ranks its entity against all others of the
// bucket, read as warp-wide broadcasts from shared memory).  Buckets are in node order and the in-bucket order is the
// lexicographic order of the remaining nodes, so shared pairs come out in exactly the order of the sorted tuples (lower
// entity id first); surface entities (rare) are appended to a list that is then sorted by their slot-major position.
// 6.9 ms on that mesh (DESIGN.md section 3.2 has the steps and their measured effect).  The scatter order inside a bucket is arbitrary (atomics), the
// ranks are not: outputs are deterministic.  Buckets larger than BUCKET_MAX (a node of extreme valence) fall back to the
// radix path.
#include <cstdlib>
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
  // bucket path: the answers themselves, compact
  bool buckets = false;
  int* surf_pos = nullptr;    // [K] slot-major positions (slot*M+e) of the entities that appear once, ascending
  uint2* pair_ents = nullptr; // [S] (entity a, entity b), a < b, in lexicographic order of the shared tuple
  // ... or, uncompacted, per bucket: pairs of bucket b = pairbuf[bptr[b] .. +pbase[b+1]-pbase[b]), output rows from pbase[b]
  uint2* pairbuf = nullptr;
  int *bptr = nullptr, *pbase = nullptr;
  long long nb = 0;
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

// out[0] = largest node id, out[1] = 1 if any id is negative or does not fit int32 (the buckets are indexed by node id)
template <typename I>
__global__ void max_node_kernel(const I* __restrict__ conn, long long total, int* __restrict__ out) {
  int m = 0;
  bool bad = false;
  for (long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x; k < total; k += (long long)gridDim.x * blockDim.x) {
    const long long v = ldidx(conn + k);
    bad |= v < 0 || v > 0x7ffffffell;
    m = max(m, (int)v);
  }
  for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) atomicMax(out, m);
  if (__any_sync(0xffffffffu, bad) && (threadIdx.x & 31) == 0) out[1] = 1;
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
  FEMB_CUDA(scr.alloc(&dmax, 2));
  FEMB_CUDA(cudaMemsetAsync(dmax, 0, 2 * sizeof(int), s));
  max_node_kernel<I><<<grid_for(M * p->stride, 256), 256, 0, s>>>(conn, M * p->stride, dmax);
  int hm[2] = {0, 0};
  FEMB_CUDA(cudaMemcpyAsync(hm, dmax, 2 * sizeof(int), cudaMemcpyDeviceToHost, s));
  FEMB_CUDA(cudaStreamSynchronize(s));
  FEMB_CHECK_ARG(hm[1] == 0, "connectivity holds a negative node id (or one beyond int32)");
  const int hmax = hm[0];
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

// ---- bucket path ---------------------------------------------------------------------------------------------------------
constexpr int BUCKET_MAX = 4096;   // largest bucket the in-bucket matcher accepts (its cost is quadratic in the bucket size)

// Thread per ELEMENT: its nodes are read once, the nf canonical tuples are formed in registers, and entities with the same
// smallest node share one atomic (a tet has three faces in the bucket of its smallest node and one in the bucket of its
// second smallest: 2 atomics instead of 4).
template <typename I>
__device__ __forceinline__ void element_tuples(const I* __restrict__ conn, int stride, const EntTable& tab, long long e, int (*t)[4]) {
  int nd[8];
  for (int k = 0; k < 8; ++k) nd[k] = k < stride ? (int)ldidx(conn + e * stride + k) : 0;
  for (int f = 0; f < tab.nf; ++f) {
    for (int k = 0; k < 4; ++k) t[f][k] = k < tab.nfn ? nd[tab.nodes[f][k]] : 0;
    for (int a = 1; a < tab.nfn; ++a) {  // insertion sort, nfn <= 4
      const int v = t[f][a];
      int b = a - 1;
      while (b >= 0 && t[f][b] > v) {
        t[f][b + 1] = t[f][b];
        --b;
      }
      t[f][b + 1] = v;
    }
  }
}

template <typename I>
__global__ void bucket_count(const I* __restrict__ conn, int stride, EntTable tab, long long M, int* __restrict__ cnt) {
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < M; e += (long long)gridDim.x * blockDim.x) {
    int t[6][4];
    element_tuples(conn, stride, tab, e, t);
    for (int f = 0; f < tab.nf; ++f) {
      bool first = true;
      int c = 0;
      for (int g = 0; g < tab.nf; ++g) {
        if (t[g][0] == t[f][0]) {
          if (g < f) first = false;
          ++c;
        }
      }
      if (first) atomicAdd(cnt + t[f][0], c);
    }
  }
}

// record of a scattered entity: key = (t1 << 32 | t2) (the bucket is t0), t3 (4-node entities only), entity id
template <typename I>
__global__ void bucket_scatter(const I* __restrict__ conn, int stride, EntTable tab, long long M, const int* __restrict__ bptr, int* __restrict__ cur,
                               unsigned long long* __restrict__ key, int* __restrict__ key3, int* __restrict__ ent) {
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < M; e += (long long)gridDim.x * blockDim.x) {
    int t[6][4];
    element_tuples(conn, stride, tab, e, t);
    for (int f = 0; f < tab.nf; ++f) {
      bool first = true;
      int c = 0;
      for (int g = 0; g < tab.nf; ++g) {
        if (t[g][0] == t[f][0]) {
          if (g < f) first = false;
          ++c;
        }
      }
      if (!first) continue;
      int pos = bptr[t[f][0]] + atomicAdd(cur + t[f][0], c);
      for (int g = f; g < tab.nf; ++g) {
        if (t[g][0] != t[f][0]) continue;
        key[pos] = (unsigned long long)(unsigned)t[g][1] << 32 | (unsigned long long)(unsigned)(tab.nfn > 2 ? t[g][2] : 0);
        if (key3) key3[pos] = t[g][3];
        ent[pos] = (int)(e * tab.nf + g);
        ++pos;
      }
    }
  }
}

// Tetrahedron faces -- the entity kind of the large meshes -- without the table-driven loops above (whose runtime indices
// put the tuples in local memory): the four sorted triples are formed in registers, and lanes of a warp whose elements have
// the same bucket share ONE atomic (neighbouring elements of a mesh generator's or a partitioner's ordering usually have the
// same smallest node: the six tets of a Kuhn cube all do).  Three faces of a tet hold its smallest node (all four when the
// element repeats that node), the fourth goes to the bucket of its own smallest node.
struct TetFaces {
  int t[4][3];   // ascending node triples of the faces (012)(013)(123)(023), element.py:722-727
  int m, tb;     // bucket of the faces holding the element's smallest node / bucket of the remaining face
  bool has_b;
};

template <typename I>
__device__ __forceinline__ TetFaces tet_faces(const I* __restrict__ conn, int stride, long long e) {
  int n[4];
  if (stride == 4 && sizeof(I) == 8) {
    long long q[4];
    asm volatile("ld.global.nc.L1::no_allocate.v4.s64 {%0,%1,%2,%3}, [%4];" : "=l"(q[0]), "=l"(q[1]), "=l"(q[2]), "=l"(q[3]) : "l"(conn + 4 * e));
#pragma unroll
    for (int k = 0; k < 4; ++k) n[k] = (int)q[k];
  } else if (stride == 4) {
    const int4 q = __ldg(reinterpret_cast<const int4*>(conn) + e);
    n[0] = q.x, n[1] = q.y, n[2] = q.z, n[3] = q.w;
  } else {
#pragma unroll
    for (int k = 0; k < 4; ++k) n[k] = (int)ldidx(conn + e * stride + k);
  }
  TetFaces r;
  constexpr int F[4][3] = {{0, 1, 2}, {0, 1, 3}, {1, 2, 3}, {0, 2, 3}};
#pragma unroll
  for (int f = 0; f < 4; ++f) {
    int a = n[F[f][0]], b = n[F[f][1]], c = n[F[f][2]];
    const int lo = min(a, b), hi = max(a, b);
    r.t[f][0] = min(lo, c);
    r.t[f][2] = max(hi, c);
    r.t[f][1] = max(lo, min(hi, c));
  }
  r.m = min(min(n[0], n[1]), min(n[2], n[3]));
  r.has_b = false, r.tb = 0;
#pragma unroll
  for (int f = 0; f < 4; ++f)
    if (r.t[f][0] != r.m) r.has_b = true, r.tb = r.t[f][0];   // at most one face lacks the smallest node
  return r;
}

template <typename I>
__global__ void __launch_bounds__(256) tet_bucket_count(const I* __restrict__ conn, int stride, long long M, int* __restrict__ cnt) {
  const int lane = threadIdx.x & 31;
  for (long long base = blockIdx.x * (long long)blockDim.x + threadIdx.x - lane; base < M; base += (long long)gridDim.x * blockDim.x) {
    const long long e = base + lane;
    const bool on = e < M;
    TetFaces r;
    if (on) r = tet_faces(conn, stride, e);
    const bool hb = on && r.has_b;
    // idle lanes carry distinct negative keys: they match nobody
    const unsigned ma = __match_any_sync(0xffffffffu, on ? r.m : ~lane), mb = __match_any_sync(0xffffffffu, hb ? r.tb : ~lane);
    const unsigned four = __ballot_sync(0xffffffffu, on && !r.has_b);
    if (on && lane == __ffs(ma) - 1) atomicAdd(cnt + r.m, 3 * __popc(ma) + __popc(ma & four));
    if (hb && lane == __ffs(mb) - 1) atomicAdd(cnt + r.tb, __popc(mb));
  }
}

template <typename I>
__global__ void __launch_bounds__(256) tet_bucket_scatter(const I* __restrict__ conn, int stride, long long M, const int* __restrict__ bptr,
                                                          int* __restrict__ cur, unsigned long long* __restrict__ key, int* __restrict__ ent) {
  const int lane = threadIdx.x & 31;
  const unsigned lt = (1u << lane) - 1u;
  for (long long base = blockIdx.x * (long long)blockDim.x + threadIdx.x - lane; base < M; base += (long long)gridDim.x * blockDim.x) {
    const long long e = base + lane;
    const bool on = e < M;
    TetFaces r;
    if (on) r = tet_faces(conn, stride, e);
    const bool hb = on && r.has_b;
    const unsigned ma = __match_any_sync(0xffffffffu, on ? r.m : ~lane), mb = __match_any_sync(0xffffffffu, hb ? r.tb : ~lane);
    const unsigned four = __ballot_sync(0xffffffffu, on && !r.has_b);
    // the lowest lane of a group reserves the group's slots
    const int la = __ffs(ma) - 1, lb = __ffs(mb) - 1;
    const int ga = __popc(ma);
    int pa = 0, pb = 0;
    if (on && lane == la) pa = bptr[r.m] + atomicAdd(cur + r.m, 3 * ga + __popc(ma & four));
    if (hb && lane == lb) pb = bptr[r.tb] + atomicAdd(cur + r.tb, __popc(mb));
    // slot of a lane's i-th face in the group's range: i * (lanes in the group) + (rank of the lane) -- one store instruction
    // of the warp then fills consecutive slots (1.93 -> 1.65 ms on the 64 M-tet mesh against three slots per lane); the fourth
    // faces of elements that repeat their smallest node come last.  Two elements per thread (connectivity of both requested
    // before the first atomic round trip) measured the same: not kept.
    const int base_a = __shfl_sync(0xffffffffu, pa, la);
    pa = base_a + __popc(ma & lt);
    const int pa4 = base_a + 3 * ga + __popc(ma & four & lt);
    pb = __shfl_sync(0xffffffffu, pb, lb) + __popc(mb & lt);
    if (!on) continue;
    int ia = 0;
#pragma unroll
    for (int f = 0; f < 4; ++f) {
      const bool in_a = r.t[f][0] == r.m;
      const int pos = in_a ? (ia < 3 ? pa + ia * ga : pa4) : pb;
      ia += in_a;
      key[pos] = (unsigned long long)(unsigned)r.t[f][1] << 32 | (unsigned long long)(unsigned)r.t[f][2];
      ent[pos] = (int)(e * 4 + f);
    }
  }
}

// One WARP per bucket.  Lane l owns the bucket's entries l, l+32, ...; the bucket is staged in the warp's shared-memory
// slice (buckets above BUCKET_STAGE entries are read from global memory instead), and every lane walks ALL entries of the
// bucket -- the same address for the whole warp, i.e. one broadcast read and no divergence -- to find for its own entry:
//   mult     how many entries carry the same tuple            lowest   whether it has the smallest entity id among them
//   partner  the other entry of its group                      rank     its position in the order (tuple, entity id)
// rank is a permutation of 0..cnt-1, so writing "is the first of exactly two" to flag[rank] sorts those flags by tuple; a
// head's output position inside the bucket is the number of heads before it (ballot + popc per 32 ranks).
//   appears once        -> appended to the surface list (slot-major position; the list is sorted afterwards)
//   first of exactly two -> pairbuf[bucket start + head rank] = (entity, partner entity)      lane 0: npairs[bucket]
constexpr int BUCKET_STAGE = 64;
constexpr int MATCH_WARPS = 8;

// Fast path of bucket_match: one entry per lane, bitonic sort of the warp by tuple (15 compare-exchange steps over
// shuffles); equal tuples are then neighbours: a group of one is a surface entity, a group of exactly two a shared pair
// (lower entity id first), and the heads are already in tuple order.
template <typename K, bool WIDE>
__device__ __forceinline__ void bucket_sort_emit(K key, int k3, int en, int lane, int cnt, int s0, long long b, const EntTable& tab, long long M,
                                                 int* __restrict__ npairs, uint2* __restrict__ pairbuf, int* __restrict__ surf_list,
                                                 int* __restrict__ surf_count) {
#pragma unroll
  for (int k = 2; k <= 32; k <<= 1) {
#pragma unroll
    for (int j = k >> 1; j > 0; j >>= 1) {
      const K ok = __shfl_xor_sync(0xffffffffu, key, j);
      const int oe = __shfl_xor_sync(0xffffffffu, en, j);
      const bool want_min = ((lane & k) == 0) == ((lane & j) == 0);
      if (!WIDE) {  // one min/max with the direction as its predicate; equal keys keep their own entity on both sides
        const K nk = want_min ? min(key, ok) : max(key, ok);
        en = nk != key ? oe : en;
        key = nk;
      } else {
        const int ok3 = __shfl_xor_sync(0xffffffffu, k3, j);
        const bool o_less = ok < key || (ok == key && ok3 < k3);
        const bool o_more = key < ok || (ok == key && k3 < ok3);
        if (want_min ? o_less : o_more) key = ok, k3 = ok3, en = oe;
      }
    }
  }
  const K kp1 = __shfl_up_sync(0xffffffffu, key, 1), kn1 = __shfl_down_sync(0xffffffffu, key, 1), kn2 = __shfl_down_sync(0xffffffffu, key, 2);
  const int tp1 = WIDE ? __shfl_up_sync(0xffffffffu, k3, 1) : 0, tn1 = WIDE ? __shfl_down_sync(0xffffffffu, k3, 1) : 0,
            tn2 = WIDE ? __shfl_down_sync(0xffffffffu, k3, 2) : 0;
  const int en1 = __shfl_down_sync(0xffffffffu, en, 1);
  const bool on = lane < cnt;
  const bool eq_prev = lane > 0 && kp1 == key && (!WIDE || tp1 == k3);
  const bool eq_n1 = lane + 1 < cnt && kn1 == key && (!WIDE || tn1 == k3);
  const bool eq_n2 = lane + 2 < cnt && kn2 == key && (!WIDE || tn2 == k3);
  const bool once = on && !eq_prev && !eq_n1, head = on && !eq_prev && eq_n1 && !eq_n2;
  const unsigned bal = __ballot_sync(0xffffffffu, head);
  if (head) pairbuf[s0 + __popc(bal & ((1u << lane) - 1u))] = make_uint2((unsigned)min(en, en1), (unsigned)max(en, en1));
  if (once) {
    const int e = en / tab.nf, f = en - e * tab.nf;
    surf_list[atomicAdd(surf_count, 1)] = (int)((long long)tab.surf_slot[f] * M + e);
  }
  if (lane == 0) npairs[b] = __popc(bal);
}

template <bool WIDE>
__global__ void __launch_bounds__(MATCH_WARPS * 32) bucket_match(const int* __restrict__ bptr, long long nb, const unsigned long long* __restrict__ gkey,
                                                                  const int* __restrict__ gkey3, const int* __restrict__ gent, EntTable tab, long long M,
                                                                  int* __restrict__ npairs, uint2* __restrict__ pairbuf, unsigned char* __restrict__ gflag,
                                                                  int* __restrict__ surf_list, int* __restrict__ surf_count, bool key32) {
  __shared__ unsigned long long skey[MATCH_WARPS][BUCKET_STAGE];
  __shared__ int skey3[WIDE ? MATCH_WARPS : 1][BUCKET_STAGE];
  __shared__ int sent[MATCH_WARPS][BUCKET_STAGE];
  __shared__ unsigned char sflag[MATCH_WARPS][BUCKET_STAGE];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const long long warp = (long long)blockIdx.x * MATCH_WARPS + w, nwarps = (long long)gridDim.x * MATCH_WARPS;
  // software pipeline over the warp's buckets: the bounds of bucket b + 2*nwarps and the records of bucket b + nwarps are
  // requested before bucket b is sorted (two dependent global loads per bucket would otherwise be fully exposed)
  int q0 = 0, q1 = 0, r0 = 0, r1 = 0;  // bounds of the next / next-but-one bucket
  if (warp < nb) q0 = bptr[warp], q1 = bptr[warp + 1];
  if (warp + nwarps < nb) r0 = bptr[warp + nwarps], r1 = bptr[warp + nwarps + 1];
  unsigned long long pkey = 0ull;
  int pk3 = 0, pen = 0;
  if (lane < q1 - q0 && q1 - q0 <= 32) {
    pkey = gkey[q0 + lane], pen = gent[q0 + lane];
    if (WIDE) pk3 = gkey3[q0 + lane];
  }
  for (long long b = warp; b < nb; b += nwarps) {
    const int s0 = q0, cnt = q1 - q0;
    const unsigned long long ckey = pkey;
    const int ck3 = pk3, cen = pen;
    q0 = r0, q1 = r1;
    if (b + 2 * nwarps < nb) r0 = bptr[b + 2 * nwarps], r1 = bptr[b + 2 * nwarps + 1];
    if (b + nwarps < nb && lane < q1 - q0 && q1 - q0 <= 32) {
      pkey = gkey[q0 + lane], pen = gent[q0 + lane];
      if (WIDE) pk3 = gkey3[q0 + lane];
    }
    if (cnt == 0) continue;
    if (cnt <= 32) {
      // Fast path (a tet mesh has ~24 faces per bucket): bucket_sort_emit, one entry per lane.
      // 3-node tuples whose two remaining nodes lie within 65534 ids of the bucket's node (every bucket of a mesh numbered
      // with any locality) sort on ONE 32-bit word -- two shuffles per compare-exchange step instead of three.
      bool narrow = false;
      if (!WIDE && key32) {
        const unsigned d2 = lane < cnt ? (unsigned)ckey - (unsigned)b : 0u, d1 = lane < cnt ? (unsigned)(ckey >> 32) - (unsigned)b : 0u;
        narrow = __reduce_max_sync(0xffffffffu, max(d1, d2)) < 65535u;
        if (narrow) bucket_sort_emit<unsigned, false>(lane < cnt ? (d1 << 16 | d2) : ~0u, 0, lane < cnt ? cen : 0, lane, cnt, s0, b, tab, M, npairs, pairbuf, surf_list, surf_count);
      }
      if (!narrow)
        bucket_sort_emit<unsigned long long, WIDE>(lane < cnt ? ckey : ~0ull, WIDE ? (lane < cnt ? ck3 : 0x7fffffff) : 0, lane < cnt ? cen : 0, lane, cnt, s0, b,
                                                   tab, M, npairs, pairbuf, surf_list, surf_count);
      continue;
    }
    const bool staged = cnt <= BUCKET_STAGE;
    const unsigned long long* kp = gkey + s0;
    const int* k3p = WIDE ? gkey3 + s0 : nullptr;
    const int* ep = gent + s0;
    unsigned char* fp = gflag + s0;
    if (staged) {
      for (int i = lane; i < cnt; i += 32) {
        skey[w][i] = kp[i], sent[w][i] = ep[i];
        if (WIDE) skey3[w][i] = k3p[i];
      }
      kp = skey[w], ep = sent[w], fp = sflag[w];
      if (WIDE) k3p = skey3[w];
    }
    __syncwarp();
    // pass 1: classify my entries, publish head flags by rank
    for (int i0 = 0; i0 < cnt; i0 += 32) {
      const int i = i0 + lane;
      const bool on = i < cnt;
      const unsigned long long mk = on ? kp[i] : 0ull;
      const int mk3 = (WIDE && on) ? k3p[i] : 0, me = on ? ep[i] : 0;
      int mult = 0, partner = 0, rank = 0;
      bool lowest = true;
      for (int j = 0; j < cnt; ++j) {
        const unsigned long long kj = kp[j];
        const int ej = ep[j];
        const int k3j = WIDE ? k3p[j] : 0;
        const bool same = kj == mk && (!WIDE || k3j == mk3);
        const bool before = kj < mk || (kj == mk && WIDE && k3j < mk3) || (same && ej < me);
        mult += same;
        rank += before;
        if (same && j != i) partner = j, lowest &= ej > me;
      }
      const bool head = on && mult == 2 && lowest;
      if (on) fp[rank] = head ? 1 : 0;
      if (on && mult == 1) {
        const int e = me / tab.nf, f = me - e * tab.nf;
        surf_list[atomicAdd(surf_count, 1)] = (int)((long long)tab.surf_slot[f] * M + e);
      }
      if (!staged) __threadfence_block();
      __syncwarp();
      (void)head, (void)partner;  // the flags are complete only after the last chunk -> pass 2
    }
    {
      // pass 2 (rare): recompute rank / head of my entries and count the heads before each of them from the published flags
      __syncwarp();
      int total = 0;
      for (int r0 = 0; r0 < cnt; r0 += 32) total += __popc(__ballot_sync(0xffffffffu, r0 + lane < cnt && fp[r0 + lane]));
      for (int i0 = 0; i0 < cnt; i0 += 32) {
        const int i = i0 + lane;
        const bool on = i < cnt;
        const unsigned long long mk = on ? kp[i] : 0ull;
        const int mk3 = (WIDE && on) ? k3p[i] : 0, me = on ? ep[i] : 0;
        int mult = 0, partner = 0, rank = 0;
        bool lowest = true;
        for (int j = 0; j < cnt; ++j) {
          const unsigned long long kj = kp[j];
          const int ej = ep[j];
          const int k3j = WIDE ? k3p[j] : 0;
          const bool same = kj == mk && (!WIDE || k3j == mk3);
          const bool before = kj < mk || (kj == mk && WIDE && k3j < mk3) || (same && ej < me);
          mult += same;
          rank += before;
          if (same && j != i) partner = j, lowest &= ej > me;
        }
        if (on && mult == 2 && lowest) {
          int hb = 0;
          for (int r = 0; r < rank; ++r) hb += fp[r];
          pairbuf[s0 + hb] = make_uint2((unsigned)me, (unsigned)ep[partner]);
        }
      }
      if (lane == 0) npairs[b] = total;
    }
    __syncwarp();
  }
}

// pairs of bucket b: pairbuf[bptr[b] .. +npairs[b]) -> out[pbase[b] ..); 8 lanes per bucket
__global__ void bucket_compact_pairs(const int* __restrict__ bptr, const int* __restrict__ npairs, const int* __restrict__ pbase, long long nb,
                                     const uint2* __restrict__ pairbuf, uint2* __restrict__ out) {
  const int sub = threadIdx.x & 7;
  for (long long b = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 3; b < nb; b += ((long long)gridDim.x * blockDim.x) >> 3) {
    const int np = npairs[b];
    const long long src = bptr[b], dst = pbase[b];
    for (int k = sub; k < np; k += 8) out[dst + k] = pairbuf[src + k];
  }
}

template <typename I>
__global__ void write_surface_list(const I* __restrict__ conn, int stride, EntTable tab, const int* __restrict__ surf_pos, long long K, long long M,
                                   long long* __restrict__ faces, long long* __restrict__ extra) {
  for (long long o = blockIdx.x * (long long)blockDim.x + threadIdx.x; o < K; o += (long long)gridDim.x * blockDim.x) {
    const long long pos = surf_pos[o];
    const int slot = (int)(pos / M);
    const long long e = pos - (long long)slot * M;
    int f = 0;
    for (int q = 0; q < tab.nf; ++q)
      if (tab.surf_slot[q] == slot) f = q;
    for (int c = 0; c < tab.nfn; ++c) faces[o * tab.nfn + c] = ldidx(conn + e * stride + tab.nodes[f][c]);
    extra[o] = ldidx(conn + e * stride + tab.extra[f]);
  }
}

// one 32-byte store per pair: a warp writes 1 KB of complete sectors (four 8-byte stores per lane would touch every line 4x)
template <int NF>
__global__ void write_shared_list(int nf_rt, const uint2* __restrict__ pair_ents, long long S, long long* __restrict__ pairs) {
  const unsigned nf = NF ? NF : nf_rt;
  for (long long o = blockIdx.x * (long long)blockDim.x + threadIdx.x; o < S; o += (long long)gridDim.x * blockDim.x) {
    const uint2 q = __ldg(pair_ents + o);
    const long long a = q.x / nf, b = q.x % nf, c = q.y / nf, d = q.y % nf;
    asm volatile("st.global.v4.s64 [%0], {%1,%2,%3,%4};" ::"l"(pairs + 4 * o), "l"(a), "l"(b), "l"(c), "l"(d) : "memory");
  }
}

// The same straight from the per-bucket pair lists (8 lanes per bucket, ~12 pairs per bucket on a tet mesh): no compaction
// pass, the 32-byte rows of consecutive buckets are consecutive in the output.
template <int NF>
__global__ void write_shared_buckets(int nf_rt, const int* __restrict__ bptr, const int* __restrict__ pbase, long long nb,
                                     const uint2* __restrict__ pairbuf, long long* __restrict__ pairs) {
  const unsigned nf = NF ? NF : nf_rt;
  const int sub = threadIdx.x & 7;
  for (long long b = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 3; b < nb; b += ((long long)gridDim.x * blockDim.x) >> 3) {
    const long long dst = pbase[b], src = bptr[b];
    const int np = pbase[b + 1] - (int)dst;
    for (int k = sub; k < np; k += 8) {
      const uint2 q = __ldg(pairbuf + src + k);
      const long long a = q.x / nf, bb = q.x % nf, c = q.y / nf, d = q.y % nf;
      asm volatile("st.global.v4.s64 [%0], {%1,%2,%3,%4};" ::"l"(pairs + 4 * (dst + k)), "l"(a), "l"(bb), "l"(c), "l"(d) : "memory");
    }
  }
}

// FEMB_OK = done, TOPO_NOT_APPLICABLE = a bucket is too large (use the radix path), anything else = error
constexpr int TOPO_NOT_APPLICABLE = -100;
template <typename I>
static int entities_build_buckets(femb_entity_plan* p, cudaStream_t s) {
  const I* conn = static_cast<const I*>(p->conn);
  const long long T = p->T, M = p->M;
  const EntTable tab = p->tab;
  Scratch scr(s);
  int* dmax;
  FEMB_CUDA(scr.alloc(&dmax, 3));   // [0] largest id, [1] bad-id flag, [2] largest bucket (below)
  FEMB_CUDA(cudaMemsetAsync(dmax, 0, 3 * sizeof(int), s));
  max_node_kernel<I><<<grid_for(M * p->stride, 256), 256, 0, s>>>(conn, M * p->stride, dmax);
  int hm[2] = {0, 0};
  FEMB_CUDA(cudaMemcpyAsync(hm, dmax, 2 * sizeof(int), cudaMemcpyDeviceToHost, s));
  FEMB_CUDA(cudaStreamSynchronize(s));
  FEMB_CHECK_ARG(hm[1] == 0, "connectivity holds a negative node id (or one beyond int32)");
  const long long nb = (long long)hm[0] + 1;
  // FEMB_TOPO_OPT (A/B switches, default all on): 1 = blocks walk the elements in order (no grid cap: the buckets a block
  // fills stay in L2 until their last contributor has passed), 2 = tet fast path, 4 = 32-bit in-bucket keys, 8 = the plan
  // keeps the per-bucket pair lists (no compaction pass)
  const int opt = getenv("FEMB_TOPO_OPT") ? atoi(getenv("FEMB_TOPO_OPT")) : 15;   // read per call: tools/topo_rate.py sweeps it
  const bool keep_lists = (opt & 8) != 0;
  int *cnt, *bptr, *cur, *npairs, *pbase, *ent, *surf_list;
  uint2* pairbuf;
  FEMB_CUDA(scr.alloc(&cnt, nb + 1));
  FEMB_CUDA(scr.alloc(&cur, nb + 1));
  FEMB_CUDA(scr.alloc(&npairs, nb + 1));
  if (keep_lists) {  // owned by the plan (released by entities_free, also on the error paths of the caller)
    FEMB_CUDA(cudaMallocAsync((void**)&p->bptr, sizeof(int) * (size_t)(nb + 1), s));
    FEMB_CUDA(cudaMallocAsync((void**)&p->pbase, sizeof(int) * (size_t)(nb + 1), s));
    bptr = p->bptr, pbase = p->pbase, p->nb = nb;
  } else {
    FEMB_CUDA(scr.alloc(&bptr, nb + 1));
    FEMB_CUDA(scr.alloc(&pbase, nb + 1));
  }
  FEMB_CUDA(cudaMemsetAsync(cnt, 0, sizeof(int) * (nb + 1), s));
  FEMB_CUDA(cudaMemsetAsync(cur, 0, sizeof(int) * (nb + 1), s));
  FEMB_CUDA(cudaMemsetAsync(npairs, 0, sizeof(int) * (nb + 1), s));
  const int egrid = grid_for(M, 256, (opt & 1) ? (1 << 20) : 64);
  const bool tets = (opt & 2) && tab.nf == 4 && tab.nfn == 3 && p->stride >= 4;
  if (tets) tet_bucket_count<I><<<egrid, 256, 0, s>>>(conn, p->stride, M, cnt);
  else bucket_count<I><<<egrid, 256, 0, s>>>(conn, p->stride, tab, M, cnt);
  FEMB_LAUNCH_CHECK();
  size_t tb = 0, tb2 = 0;
  FEMB_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tb, cnt, bptr, (int)(nb + 1), s));
  FEMB_CUDA(cub::DeviceReduce::Max(nullptr, tb2, cnt, dmax + 2, (int)nb, s));
  void* tmp;
  FEMB_CUDA(scr.alloc((char**)&tmp, std::max(tb, tb2)));
  FEMB_CUDA(cub::DeviceScan::ExclusiveSum(tmp, tb, cnt, bptr, (int)(nb + 1), s));
  FEMB_CUDA(cub::DeviceReduce::Max(tmp, tb2, cnt, dmax + 2, (int)nb, s));
  int maxb = 0;
  FEMB_CUDA(cudaMemcpyAsync(&maxb, dmax + 2, sizeof(int), cudaMemcpyDeviceToHost, s));
  FEMB_CUDA(cudaStreamSynchronize(s));
  if (maxb > BUCKET_MAX) return TOPO_NOT_APPLICABLE;
  const bool wide = tab.nfn > 3;
  unsigned long long* key;
  int* key3 = nullptr;
  unsigned char* gflag;
  FEMB_CUDA(scr.alloc(&key, T));
  if (wide) FEMB_CUDA(scr.alloc(&key3, T));
  FEMB_CUDA(scr.alloc(&ent, T));
  if (keep_lists) {
    FEMB_CUDA(cudaMallocAsync((void**)&p->pairbuf, sizeof(uint2) * (size_t)T, s));
    pairbuf = p->pairbuf;
  } else {
    FEMB_CUDA(scr.alloc(&pairbuf, T));
  }
  FEMB_CUDA(scr.alloc(&gflag, T));
  FEMB_CUDA(scr.alloc(&surf_list, T + 1));
  int* surf_count = surf_list + T;
  FEMB_CUDA(cudaMemsetAsync(surf_count, 0, sizeof(int), s));
  if (tets) tet_bucket_scatter<I><<<egrid, 256, 0, s>>>(conn, p->stride, M, bptr, cur, key, ent);
  else bucket_scatter<I><<<egrid, 256, 0, s>>>(conn, p->stride, tab, M, bptr, cur, key, key3, ent);
  FEMB_LAUNCH_CHECK();
  const int mgrid = grid_for(nb, MATCH_WARPS, 8);
  const bool key32 = (opt & 4) && tab.nfn == 3;
  if (wide) bucket_match<true><<<mgrid, MATCH_WARPS * 32, 0, s>>>(bptr, nb, key, key3, ent, tab, M, npairs, pairbuf, gflag, surf_list, surf_count, false);
  else bucket_match<false><<<mgrid, MATCH_WARPS * 32, 0, s>>>(bptr, nb, key, key3, ent, tab, M, npairs, pairbuf, gflag, surf_list, surf_count, key32);
  FEMB_LAUNCH_CHECK();
  FEMB_CUDA(cub::DeviceScan::ExclusiveSum(tmp, tb, npairs, pbase, (int)(nb + 1), s));
  int hs = 0, hk = 0;
  FEMB_CUDA(cudaMemcpyAsync(&hs, pbase + nb, sizeof(int), cudaMemcpyDeviceToHost, s));
  FEMB_CUDA(cudaMemcpyAsync(&hk, surf_count, sizeof(int), cudaMemcpyDeviceToHost, s));
  FEMB_CUDA(cudaStreamSynchronize(s));
  p->S = hs, p->K = hk;
  if (hs > 0 && !keep_lists) {
    FEMB_CUDA(cudaMallocAsync((void**)&p->pair_ents, sizeof(uint2) * (size_t)hs, s));
    bucket_compact_pairs<<<grid_for(nb * 8, 256), 256, 0, s>>>(bptr, npairs, pbase, nb, pairbuf, p->pair_ents);
    FEMB_LAUNCH_CHECK();
  }
  if (hk > 0) {
    FEMB_CUDA(cudaMallocAsync((void**)&p->surf_pos, sizeof(int) * (size_t)hk, s));
    int bits = 1;
    while ((1ll << bits) < T) ++bits;
    size_t tb3 = 0;
    FEMB_CUDA(cub::DeviceRadixSort::SortKeys(nullptr, tb3, surf_list, p->surf_pos, hk, 0, bits, s));
    void* tmp3;
    FEMB_CUDA(scr.alloc((char**)&tmp3, tb3));
    FEMB_CUDA(cub::DeviceRadixSort::SortKeys(tmp3, tb3, surf_list, p->surf_pos, hk, 0, bits, s));
  }
  p->buckets = true;
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
  cudaFree(p->surf_pos);
  cudaFree(p->pair_ents);
  cudaFree(p->pairbuf);
  cudaFree(p->bptr);
  cudaFree(p->pbase);
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
  static const bool force_radix = getenv("FEMB_TOPO_RADIX") != nullptr;  // A/B switch: the radix-sort path
  if (M > 0 && !force_radix) {
    rc = ib == 8 ? entities_build_buckets<long long>(p, as_stream(stream)) : entities_build_buckets<int>(p, as_stream(stream));
  }
  if (M > 0 && (force_radix || rc == TOPO_NOT_APPLICABLE)) {
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
  if (p->buckets) {
    const int g = grid_for(p->K, 256);
    if (p->ib == 8) write_surface_list<long long><<<g, 256, 0, s>>>((const long long*)p->conn, p->stride, p->tab, p->surf_pos, p->K, p->M, (long long*)faces, (long long*)extra);
    else write_surface_list<int><<<g, 256, 0, s>>>((const int*)p->conn, p->stride, p->tab, p->surf_pos, p->K, p->M, (long long*)faces, (long long*)extra);
    FEMB_LAUNCH_CHECK();
    return FEMB_OK;
  }
  const int grid = grid_for(p->T, 256);
  if (p->ib == 8) write_surface<long long><<<grid, 256, 0, s>>>((const long long*)p->conn, p->stride, p->tab, p->sscan, p->T, p->M, (long long*)faces, (long long*)extra);
  else write_surface<int><<<grid, 256, 0, s>>>((const int*)p->conn, p->stride, p->tab, p->sscan, p->T, p->M, (long long*)faces, (long long*)extra);
  FEMB_LAUNCH_CHECK();
  return FEMB_OK;
}

extern "C" int femb_entities_shared(femb_entity_plan* p, int64_t* pairs, femb_stream stream) {
  FEMB_CHECK_ARG(p != nullptr, "plan");
  if (p->T == 0 || p->S == 0) return FEMB_OK;
  if (p->buckets) {
    FEMB_CHECK_ARG((reinterpret_cast<uintptr_t>(pairs) & 31) == 0, "pairs must be 32-byte aligned");
    if (p->pairbuf) {
      const int gb = grid_for(p->nb * 8, 256, 1 << 20);
      if (p->tab.nf == 4) write_shared_buckets<4><<<gb, 256, 0, as_stream(stream)>>>(4, p->bptr, p->pbase, p->nb, p->pairbuf, (long long*)pairs);
      else write_shared_buckets<0><<<gb, 256, 0, as_stream(stream)>>>(p->tab.nf, p->bptr, p->pbase, p->nb, p->pairbuf, (long long*)pairs);
      FEMB_LAUNCH_CHECK();
      return FEMB_OK;
    }
    const int g = grid_for(p->S, 256, 1 << 20);
    if (p->tab.nf == 4) write_shared_list<4><<<g, 256, 0, as_stream(stream)>>>(4, p->pair_ents, p->S, (long long*)pairs);
    else write_shared_list<0><<<g, 256, 0, as_stream(stream)>>>(p->tab.nf, p->pair_ents, p->S, (long long*)pairs);
    FEMB_LAUNCH_CHECK();
    return FEMB_OK;
  }
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
