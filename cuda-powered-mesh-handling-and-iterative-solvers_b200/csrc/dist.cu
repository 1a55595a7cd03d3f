// Multi-GPU conjugate gradient over NVLink peer memory (one process per GPU, rows partitioned).
//
// The reference has no distributed code.  Each rank owns a block of rows (local numbering [owned | ghost]) and keeps its
// search direction p in a cudaIpc-shared "symmetric" buffer.  One CG iteration is four kernels, all in one CUDA graph:
//   push  boundary entries of p are stored straight into the neighbours' ghost slots (st.global on mapped peer pointers
//         through NVSwitch), then flag A is raised on every neighbour
//   k1    waits for flag A of its neighbours, CSR-stream SpMV on the owned rows, p.Ap partial; the last CTA stores the
//         partial into slot[rank] of EVERY rank's reduction array and raises flag B everywhere
//   k2    waits for flag B of all ranks, sums the P partials in rank order (deterministic, identical on every rank),
//         guards/alpha, u += alpha p, r -= alpha Ap, r.r partial -> slot[rank] everywhere, flag C
//   k3    waits for flag C, rs_new, convergence test / beta (identical on every rank), p = r + beta p
// There is no NCCL call and no host involvement inside the loop: the "collectives" are peer stores plus epoch flags, which
// costs a few microseconds instead of tens per all-reduce -- the difference between ~4x and >6x strong scaling at 8 GPUs
// when an iteration is ~85 us of compute.  All spin loops carry a timeout so a lost rank turns into an error, not a hang.
#include <cstdlib>

#include "spmv_dev.cuh"

namespace femb {

constexpr int MAXP = 16;
constexpr long long SPIN_TIMEOUT_CYCLES = 4000000000ll;  // ~2 s at 2 GHz

// layout of every rank's symmetric buffer (identical on all ranks)
struct SymHeader {
  volatile long long flagA[MAXP];  // halo arrived from rank q (epoch)
  volatile long long flagB[MAXP];  // p.Ap partial of rank q arrived
  volatile long long flagC[MAXP];  // r.r partial of rank q arrived
  volatile double redB[MAXP];
  volatile double redC[MAXP];
  double pad[MAXP * 3];
};
static_assert(sizeof(SymHeader) % 256 == 0, "header keeps p 256-byte aligned");

struct DistState {
  double rs_old, rs_new, pAp, alpha, beta;
  int it, stop, status, iterations;
  long long epochA, epochB, epochC;
  unsigned int ticket_push, ticket1, ticket2, ticket3;
  long long* trace;  // optional [TRACE_ITERS][12] globaltimer stamps (FEMB_DIST_TRACE=1)
};

constexpr int TRACE_ITERS = 64;
__device__ __forceinline__ void trace_stamp(DistState* st, int slot) {
  if (st->trace && st->it < TRACE_ITERS) {
    long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    st->trace[st->it * 12 + slot] = t;
  }
}

struct Peers {
  SymHeader* hdr[MAXP];  // every rank's header as mapped in this process (own included)
  int rank, P;
  int nnbr;
  int nbr[MAXP];
  int send_ptr[MAXP + 1];      // prefix into send_idx per neighbour
  long long ghost_off[MAXP];   // where my block starts inside neighbour's p (n_owned_q + recv_off_q[rank]), in doubles
};

__device__ __forceinline__ double* sym_p(SymHeader* h) { return reinterpret_cast<double*>(h + 1); }

__device__ __forceinline__ bool spin_until(volatile long long* flag, long long expect, DistState* st) {
  const long long t0 = clock64();
  while (*flag < expect) {
    if (clock64() - t0 > SPIN_TIMEOUT_CYCLES) {
      st->stop = 3, st->status = 3;
      return false;
    }
  }
  return true;
}

// block-level wait: thread 0 polls the listed ranks' flags in MY header, everyone else parks on the barrier
__device__ __forceinline__ void wait_flags(volatile long long* flags, const int* ranks, int count, long long expect, DistState* st) {
  if (threadIdx.x < count) spin_until(flags + ranks[threadIdx.x], expect, st);  // one lane per flag: the polls overlap
  __syncthreads();
}

// ---- push: p boundary -> neighbours' ghost slots, then flag A --------------------------------------------------------
__global__ void __launch_bounds__(256) dist_push_kernel(Peers pe, const int* __restrict__ send_idx, DistState* st) {
  if (st->stop) return;
  if (blockIdx.x == 0 && threadIdx.x == 0) trace_stamp(st, 0);
  const double* p = sym_p(pe.hdr[pe.rank]);
  const int total = pe.send_ptr[pe.nnbr];
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x) {
    int k = 0;
    while (t >= pe.send_ptr[k + 1]) ++k;
    double* dst = sym_p(pe.hdr[pe.nbr[k]]) + pe.ghost_off[k];
    dst[t - pe.send_ptr[k]] = p[send_idx[t]];
  }
  __threadfence_system();
  __shared__ bool last;
  __syncthreads();
  if (threadIdx.x == 0) last = atomicAdd(&st->ticket_push, 1u) == gridDim.x - 1;
  __syncthreads();
  if (last && threadIdx.x == 0) {
    st->ticket_push = 0;
    const long long e = st->epochA + 1;
    st->epochA = e;
    __threadfence_system();
    for (int k = 0; k < pe.nnbr; ++k) pe.hdr[pe.nbr[k]]->flagA[pe.rank] = e;
    trace_stamp(st, 1);
  }
}

// ---- k1: SpMV on owned rows + p.Ap partial -> everyone ---------------------------------------------------------------
template <int LR, bool FUSED>
__global__ void __launch_bounds__(SPMV_THREADS) dist_spmv_kernel(Peers pe, long long n_owned, const int* __restrict__ crow,
                                                                 const int* __restrict__ col, const double* __restrict__ val,
                                                                 double* __restrict__ y, const unsigned char* __restrict__ mask,
                                                                 double* __restrict__ partial, DistState* st) {
  if (st->stop) return;
  SymHeader* me = pe.hdr[pe.rank];
  if (FUSED && blockIdx.x == 0 && threadIdx.x == 0) trace_stamp(st, 2);
  wait_flags(me->flagA, pe.nbr, pe.nnbr, st->epochA, st);
  if (st->stop) return;
  if (FUSED && blockIdx.x == 0 && threadIdx.x == 0) trace_stamp(st, 3);
  const double* x = sym_p(me);
  const double dot = spmv_stream_rows<LR, false>(n_owned, crow, col, val, x, y, mask, false, FUSED);
  if (!FUSED) return;
  const double t = block_sum<SPMV_THREADS>(dot);
  __shared__ bool last;
  if (threadIdx.x == 0) {
    partial[blockIdx.x] = t;
    __threadfence();
    last = atomicAdd(&st->ticket1, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (last) {
    __threadfence();
    double a = 0.0;
    for (int k = threadIdx.x; k < (int)gridDim.x; k += SPMV_THREADS) a += ((volatile double*)partial)[k];
    a = block_sum<SPMV_THREADS>(a);
    if (threadIdx.x == 0) {
      st->ticket1 = 0;
      const long long e = st->epochB + 1;
      st->epochB = e;
      for (int q = 0; q < pe.P; ++q) pe.hdr[q]->redB[pe.rank] = a;
      __threadfence_system();
      for (int q = 0; q < pe.P; ++q) pe.hdr[q]->flagB[pe.rank] = e;
      trace_stamp(st, 4);
    }
  }
}

constexpr int DV_THREADS = 256;

__device__ __forceinline__ double sum_slots(volatile double* red, int P) {
  double a = 0.0;
  for (int q = 0; q < P; ++q) a += red[q];  // rank order: identical result on every rank
  return a;
}

// ---- k2: u += alpha p ; r -= alpha Ap ; r.r partial -> everyone -------------------------------------------------------
__global__ void __launch_bounds__(DV_THREADS) dist_update_kernel(Peers pe, long long n, double* __restrict__ u, double* __restrict__ r,
                                                                 const double* __restrict__ Ap, double* __restrict__ partial, DistState* st,
                                                                 double eps, int guards) {
  if (st->stop) return;
  SymHeader* me = pe.hdr[pe.rank];
  __shared__ int all[MAXP];
  if (threadIdx.x < MAXP) all[threadIdx.x] = threadIdx.x;
  __syncthreads();
  if (blockIdx.x == 0 && threadIdx.x == 0) trace_stamp(st, 5);
  wait_flags(me->flagB, all, pe.P, st->epochB, st);
  if (st->stop) return;
  if (blockIdx.x == 0 && threadIdx.x == 0) trace_stamp(st, 6);
  const double pAp = sum_slots(me->redB, pe.P);
  const double alpha = st->rs_old / (pAp + eps);
  if (guards && (fabs(pAp) < eps || pAp < 0.0 || !isfinite(alpha))) {  // solver.py:187-198, same verdict on every rank/CTA
    if (blockIdx.x == 0 && threadIdx.x == 0) st->stop = 1, st->status = 1, st->iterations = st->it + 1, st->pAp = pAp;
    return;
  }
  const double* p = sym_p(me);
  double dot = 0.0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const double ri = r[i] - alpha * Ap[i];
    u[i] += alpha * p[i];
    r[i] = ri;
    dot += ri * ri;
  }
  const double t = block_sum<DV_THREADS>(dot);
  __shared__ bool last;
  if (threadIdx.x == 0) {
    partial[blockIdx.x] = t;
    __threadfence();
    last = atomicAdd(&st->ticket2, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (last) {
    __threadfence();
    double a = 0.0;
    for (int k = threadIdx.x; k < (int)gridDim.x; k += DV_THREADS) a += ((volatile double*)partial)[k];
    a = block_sum<DV_THREADS>(a);
    if (threadIdx.x == 0) {
      st->ticket2 = 0;
      st->pAp = pAp, st->alpha = alpha;
      const long long e = st->epochC + 1;
      st->epochC = e;
      for (int q = 0; q < pe.P; ++q) pe.hdr[q]->redC[pe.rank] = a;
      __threadfence_system();
      for (int q = 0; q < pe.P; ++q) pe.hdr[q]->flagC[pe.rank] = e;
      trace_stamp(st, 7);
    }
  }
}

// ---- k3: convergence / beta ; p = r + beta p -------------------------------------------------------------------------
__global__ void __launch_bounds__(DV_THREADS) dist_direction_kernel(Peers pe, long long n, const double* __restrict__ r, DistState* st,
                                                                    double tol, double eps, int guards, int max_iter) {
  if (st->stop) return;
  SymHeader* me = pe.hdr[pe.rank];
  __shared__ int all[MAXP];
  if (threadIdx.x < MAXP) all[threadIdx.x] = threadIdx.x;
  __syncthreads();
  if (blockIdx.x == 0 && threadIdx.x == 0) trace_stamp(st, 8);
  wait_flags(me->flagC, all, pe.P, st->epochC, st);
  if (st->stop) return;
  if (blockIdx.x == 0 && threadIdx.x == 0) trace_stamp(st, 9);
  const double rs_new = sum_slots(me->redC, pe.P);
  const double beta = rs_new / (st->rs_old + eps);
  const bool conv = sqrt(rs_new) < tol;                  // solver.py:210-212
  const bool bad = guards && !isfinite(beta);            // solver.py:216-218
  if (conv || bad) {
    if (blockIdx.x == 0 && threadIdx.x == 0) st->stop = 1, st->status = conv ? 0 : 1, st->iterations = st->it + 1, st->rs_new = rs_new;
    return;
  }
  double* p = sym_p(me);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) p[i] = r[i] + beta * p[i];
  __shared__ bool last;
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    last = atomicAdd(&st->ticket3, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (last && threadIdx.x == 0) {  // every CTA has read rs_old by now
    st->ticket3 = 0;
    trace_stamp(st, 10);
    st->rs_new = rs_new, st->beta = beta, st->rs_old = rs_new;
    st->it += 1;
    if (st->it >= max_iter) st->stop = 2, st->status = 2, st->iterations = max_iter;
  }
}

// ---- setup ------------------------------------------------------------------------------------------------------------
__global__ void dist_load_p(Peers pe, long long n, double* __restrict__ u, const unsigned char* __restrict__ mask) {
  double* p = sym_p(pe.hdr[pe.rank]);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    if (mask && !mask[i]) u[i] = 0.0;
    p[i] = u[i];
  }
}

__global__ void __launch_bounds__(DV_THREADS) dist_init_kernel(Peers pe, long long n, const double* __restrict__ F, const double* __restrict__ Au,
                                                               const unsigned char* __restrict__ mask, double* __restrict__ r,
                                                               double* __restrict__ partial, DistState* st) {
  double* p = sym_p(pe.hdr[pe.rank]);
  double dot = 0.0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    double ri = F[i] - Au[i];
    if (mask && !mask[i]) ri = 0.0;
    r[i] = ri;
    p[i] = ri;
    dot += ri * ri;
  }
  const double t = block_sum<DV_THREADS>(dot);
  __shared__ bool last;
  if (threadIdx.x == 0) {
    partial[blockIdx.x] = t;
    __threadfence();
    last = atomicAdd(&st->ticket2, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (last) {
    __threadfence();
    double a = 0.0;
    for (int k = threadIdx.x; k < (int)gridDim.x; k += DV_THREADS) a += ((volatile double*)partial)[k];
    a = block_sum<DV_THREADS>(a);
    if (threadIdx.x == 0) {
      st->ticket2 = 0;
      const long long e = st->epochC + 1;
      st->epochC = e;
      for (int q = 0; q < pe.P; ++q) pe.hdr[q]->redC[pe.rank] = a;
      __threadfence_system();
      for (int q = 0; q < pe.P; ++q) pe.hdr[q]->flagC[pe.rank] = e;
    }
  }
}

__global__ void dist_init_finish(Peers pe, DistState* st, int max_iter) {
  SymHeader* me = pe.hdr[pe.rank];
  if (threadIdx.x == 0) {
    bool ok = true;
    for (int q = 0; q < pe.P && ok; ++q) ok = spin_until(me->flagC + q, st->epochC, st);
    if (ok) {
      const double a = sum_slots(me->redC, pe.P);
      st->rs_old = a, st->rs_new = a;
      if (max_iter <= 0) st->stop = 2, st->status = 2;
    }
  }
}

static int pick_lr(long long n, long long nnz) {
  const double avg = n > 0 ? (double)nnz / (double)n : 1.0;
  for (int lr = 1; lr <= 32; lr *= 2)
    if ((SPMV_THREADS / lr) * avg * 1.25 <= STREAM_CAP) return lr;
  return 32;
}

template <bool FUSED>
static void launch_dist_spmv(int lr, int grid, cudaStream_t s, const Peers& pe, long long n, const int* crow, const int* col, const double* val,
                             double* y, const unsigned char* mask, double* partial, DistState* st) {
  switch (lr) {
    case 1: dist_spmv_kernel<1, FUSED><<<grid, SPMV_THREADS, 0, s>>>(pe, n, crow, col, val, y, mask, partial, st); break;
    case 2: dist_spmv_kernel<2, FUSED><<<grid, SPMV_THREADS, 0, s>>>(pe, n, crow, col, val, y, mask, partial, st); break;
    case 4: dist_spmv_kernel<4, FUSED><<<grid, SPMV_THREADS, 0, s>>>(pe, n, crow, col, val, y, mask, partial, st); break;
    case 8: dist_spmv_kernel<8, FUSED><<<grid, SPMV_THREADS, 0, s>>>(pe, n, crow, col, val, y, mask, partial, st); break;
    case 16: dist_spmv_kernel<16, FUSED><<<grid, SPMV_THREADS, 0, s>>>(pe, n, crow, col, val, y, mask, partial, st); break;
    default: dist_spmv_kernel<32, FUSED><<<grid, SPMV_THREADS, 0, s>>>(pe, n, crow, col, val, y, mask, partial, st); break;
  }
}

static cudaStream_t dist_stream(cudaStream_t user) {
  static thread_local cudaStream_t s = nullptr;
  static thread_local cudaEvent_t ev = nullptr;
  if (!s) {
    if (cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking) != cudaSuccess) return nullptr;
    cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
  }
  cudaEventRecord(ev, user);
  cudaStreamWaitEvent(s, ev, 0);
  return s;
}

}  // namespace femb

using namespace femb;

extern "C" int femb_dist_header_bytes(void) { return (int)sizeof(SymHeader); }

extern "C" int femb_dist_alloc(int64_t bytes, void** ptr, void* ipc_handle64) {
  FEMB_CHECK_ARG(bytes > 0 && ptr && ipc_handle64, "bytes/ptr/handle");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "ipc handle size");
  FEMB_CUDA(cudaMalloc(ptr, (size_t)bytes));
  FEMB_CUDA(cudaMemset(*ptr, 0, (size_t)bytes));
  FEMB_CUDA(cudaIpcGetMemHandle(reinterpret_cast<cudaIpcMemHandle_t*>(ipc_handle64), *ptr));
  return FEMB_OK;
}

extern "C" int femb_dist_open(const void* ipc_handle64, void** ptr) {
  FEMB_CHECK_ARG(ptr && ipc_handle64, "ptr/handle");
  cudaIpcMemHandle_t h;
  memcpy(&h, ipc_handle64, sizeof(h));
  FEMB_CUDA(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return FEMB_OK;
}

extern "C" int femb_dist_close(void* ptr) {
  FEMB_CUDA(cudaIpcCloseMemHandle(ptr));
  return FEMB_OK;
}

extern "C" int femb_dist_free(void* ptr) {
  FEMB_CUDA(cudaFree(ptr));
  return FEMB_OK;
}

extern "C" int femb_dist_reset(void* own_sym, femb_stream stream) {
  FEMB_CUDA(cudaMemsetAsync(own_sym, 0, sizeof(SymHeader), as_stream(stream)));
  FEMB_CUDA(cudaStreamSynchronize(as_stream(stream)));
  return FEMB_OK;
}

extern "C" int femb_dist_cg_solve(int rank, int nranks, int64_t n_owned, int64_t nnz, const int32_t* crow, const int32_t* col, const double* val,
                                  const double* F, const uint8_t* mask, double* u, double* work, void* const* sym_host, int nnbr,
                                  const int32_t* nbr_host, const int32_t* send_ptr_host, const int32_t* send_idx,
                                  const int64_t* ghost_off_host, double tol, int max_iter, double eps, int check_every,
                                  femb_cg_result* result_host, femb_stream stream) {
  FEMB_CHECK_ARG(nranks >= 1 && nranks <= MAXP && rank >= 0 && rank < nranks && nnbr >= 0 && nnbr < MAXP, "rank/nranks/nnbr");
  FEMB_CHECK_ARG(n_owned > 0 && crow && col && val && F && u && work && sym_host && result_host, "null pointer / n_owned <= 0");
  if (check_every < 1) check_every = 16;
  cudaStream_t s = dist_stream(as_stream(stream));
  FEMB_CHECK_ARG(s != nullptr, "could not create the solver stream");
  Peers pe;
  memset(&pe, 0, sizeof(pe));
  pe.rank = rank, pe.P = nranks, pe.nnbr = nnbr;
  for (int q = 0; q < nranks; ++q) pe.hdr[q] = static_cast<SymHeader*>(sym_host[q]);
  for (int k = 0; k < nnbr; ++k) pe.nbr[k] = nbr_host[k], pe.ghost_off[k] = ghost_off_host[k];
  for (int k = 0; k <= nnbr; ++k) pe.send_ptr[k] = send_ptr_host[k];
  const long long n = n_owned;
  double *r = work, *Ap = work + n;
  const int lr = pick_lr(n, nnz);
  // one row tile per CTA up to 32k CTAs, then the smallest equal share: no CTA does one tile more than another
  const long long tiles = (n + SPMV_THREADS / lr - 1) / (SPMV_THREADS / lr);
  const long long per = (tiles + SMS * 32 - 1) / (SMS * 32);
  const int g1 = (int)((tiles + per - 1) / per);
  const int g2 = grid_for(n, DV_THREADS, 8);
  const int gp = std::max(1, std::min(64, (pe.send_ptr[nnbr] + 255) / 256));
  Scratch scr(s);
  double* partial;
  DistState* st;
  FEMB_CUDA(scr.alloc(&partial, (size_t)std::max(g1, g2)));
  FEMB_CUDA(scr.alloc(&st, 1));
  FEMB_CUDA(cudaMemsetAsync(st, 0, sizeof(DistState), s));
  long long* trace = nullptr;
  if (getenv("FEMB_DIST_TRACE")) {
    FEMB_CUDA(scr.alloc(&trace, (size_t)TRACE_ITERS * 12));
    FEMB_CUDA(cudaMemsetAsync(trace, 0, sizeof(long long) * TRACE_ITERS * 12, s));
    FEMB_CUDA(cudaMemcpyAsync(&st->trace, &trace, sizeof(trace), cudaMemcpyHostToDevice, s));
  }
  const int guards = 1;
  // ---- setup: p <- mask.*u, halo, Ap = A u, r = mask.*(F - Ap), p = r, rs_old = allreduce(r.r)   (solver.py:163-181)
  dist_load_p<<<g2, DV_THREADS, 0, s>>>(pe, n, u, mask);
  dist_push_kernel<<<gp, 256, 0, s>>>(pe, send_idx, st);
  launch_dist_spmv<false>(lr, g1, s, pe, n, crow, col, val, Ap, nullptr, nullptr, st);
  dist_init_kernel<<<g2, DV_THREADS, 0, s>>>(pe, n, F, Ap, mask, r, partial, st);
  dist_init_finish<<<1, 32, 0, s>>>(pe, st, max_iter);
  FEMB_LAUNCH_CHECK();
  // ---- iterations
  cudaGraph_t graph = nullptr;
  cudaGraphExec_t exec = nullptr;
  FEMB_CUDA(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
  for (int k = 0; k < check_every; ++k) {
    dist_push_kernel<<<gp, 256, 0, s>>>(pe, send_idx, st);
    launch_dist_spmv<true>(lr, g1, s, pe, n, crow, col, val, Ap, mask, partial, st);
    dist_update_kernel<<<g2, DV_THREADS, 0, s>>>(pe, n, u, r, Ap, partial, st, eps, guards);
    dist_direction_kernel<<<g2, DV_THREADS, 0, s>>>(pe, n, r, st, tol, eps, guards, max_iter);
  }
  cudaError_t ce = cudaStreamEndCapture(s, &graph);
  if (ce != cudaSuccess) {
    set_error(std::string("cudaStreamEndCapture: ") + cudaGetErrorString(ce));
    return FEMB_ERR_CUDA;
  }
  FEMB_CUDA(cudaGraphInstantiate(&exec, graph, 0));
  static thread_local DistState* hst = nullptr;
  static thread_local cudaEvent_t ev[2] = {nullptr, nullptr}, tev[2] = {nullptr, nullptr};
  if (!hst) {
    FEMB_CUDA(cudaMallocHost(&hst, 2 * sizeof(DistState)));
    for (int k = 0; k < 2; ++k) {
      FEMB_CUDA(cudaEventCreateWithFlags(&ev[k], cudaEventDisableTiming));
      FEMB_CUDA(cudaEventCreate(&tev[k]));
    }
  }
  int rc = FEMB_OK;
  const int launches = (max_iter + check_every - 1) / check_every;
  hst[0].stop = hst[1].stop = 0;
  bool stopped = false;
  cudaEventRecord(tev[0], s);
  for (int l = 0; l < launches && !stopped; ++l) {
    if (cudaGraphLaunch(exec, s) != cudaSuccess || cudaMemcpyAsync(&hst[l & 1], st, sizeof(DistState), cudaMemcpyDeviceToHost, s) != cudaSuccess ||
        cudaEventRecord(ev[l & 1], s) != cudaSuccess) {
      set_error(std::string("distributed CG graph launch: ") + cudaGetErrorString(cudaGetLastError()));
      rc = FEMB_ERR_CUDA;
      break;
    }
    if (l >= 1) {
      cudaEventSynchronize(ev[(l - 1) & 1]);
      stopped = hst[(l - 1) & 1].stop != 0;
    }
  }
  cudaEventRecord(tev[1], s);
  DistState* fin = &hst[0];
  if (rc == FEMB_OK && (cudaMemcpyAsync(fin, st, sizeof(DistState), cudaMemcpyDeviceToHost, s) != cudaSuccess || cudaStreamSynchronize(s) != cudaSuccess)) {
    set_error(std::string("distributed CG final state: ") + cudaGetErrorString(cudaGetLastError()));
    rc = FEMB_ERR_CUDA;
  }
  float ms = 0.f;
  if (rc == FEMB_OK) cudaEventElapsedTime(&ms, tev[0], tev[1]);
  cudaGraphExecDestroy(exec);
  cudaGraphDestroy(graph);
  if (rc != FEMB_OK) return rc;
  if (fin->status == 3) {
    set_error("distributed CG: timed out waiting for a peer rank (flag never arrived)");
    return FEMB_ERR_NCCL;
  }
  result_host->iterations = fin->stop ? fin->iterations : max_iter;
  result_host->status = fin->stop ? fin->status : 2;
  result_host->rs = fin->rs_new;
  result_host->loop_ms = ms;
  if (trace && fin->it > 20) {  // average phase times over iterations 10..min(it,TRACE_ITERS)-1 (microseconds)
    static long long h[TRACE_ITERS * 12];
    cudaMemcpy(h, trace, sizeof(h), cudaMemcpyDeviceToHost);
    const int i1 = std::min(fin->it, TRACE_ITERS) - 1;
    double acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int i = 10; i < i1; ++i) {
      const long long* t = h + i * 12;
      acc[0] += t[1] - t[0];                 // push
      acc[1] += t[3] - t[2];                 // k1 wait for halo
      acc[2] += t[4] - t[3];                 // k1 body + reduction + remote stores
      acc[3] += t[6] - t[5];                 // k2 wait for all-reduce B
      acc[4] += t[7] - t[6];                 // k2 body
      acc[5] += t[9] - t[8];                 // k3 wait for all-reduce C
      acc[6] += t[10] - t[9];                // k3 body
      acc[7] += (h + (i + 1) * 12)[0] - t[0];  // whole iteration
    }
    const double m = 1e-3 / (i1 - 10);
    fprintf(stderr, "[femb dist trace] rank %d: push %.1f | k1 wait %.1f body %.1f | k2 wait %.1f body %.1f | k3 wait %.1f body %.1f | iteration %.1f us\n",
            rank, acc[0] * m, acc[1] * m, acc[2] * m, acc[3] * m, acc[4] * m, acc[5] * m, acc[6] * m, acc[7] * m);
  }
  return FEMB_OK;
}
