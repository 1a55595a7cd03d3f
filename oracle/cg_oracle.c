/* CPU oracle, compiled part -- TEST / BASELINE INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * C restatement of the reference's projected CG loop (solver/solver.py:144-229) on the coalesced CSR operator of
 * subdivision.ipynb cell 6, multi-threaded with pthreads (this image's gcc has no libgomp): every thread owns a static
 * row range for the whole solve, partial dot products are combined in thread order behind barriers, and every thread takes
 * the same scalar decisions.  Used only as the CPU baseline of bench.py (`cpu_baseline`, `--impl reference`) and checked
 * against the numpy oracle in tests/test_oracle_c.py.  Parity status: pinned through the numpy oracle
 * (oracle/fem_oracle.py), which is pinned to the reference's own outputs.
 * Build: make -C oracle   ->   oracle/_build/libfemoracle.so
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <unistd.h>

#define MAXT 256

int oracle_threads(void) {
  long n = sysconf(_SC_NPROCESSORS_ONLN);
  const char* e = getenv("FEMB_ORACLE_THREADS");
  if (e && atoi(e) > 0) n = atoi(e);
  if (n < 1) n = 1;
  if (n > MAXT) n = MAXT;
  return (int)n;
}

typedef struct {
  int64_t n;
  const int64_t *crow, *col;
  const double *val, *F;
  const unsigned char* freed;
  double *u, *r, *p, *Ap;
  double tol, eps;
  int max_iter, T;
  pthread_barrier_t bar;
  double part[MAXT][8]; /* padded against false sharing */
  int iterations, status;
  const double* x_in; /* matvec-only mode */
  double* y_out;
} ctx_t;

typedef struct {
  ctx_t* c;
  int id;
} arg_t;

static void rows_of(const ctx_t* c, int id, int64_t* lo, int64_t* hi) {
  /* balance by nonzeros: thread id gets the rows whose crow falls in its share */
  const int64_t nnz = c->crow[c->n];
  const int64_t a = nnz / c->T * id, b = id == c->T - 1 ? nnz : nnz / c->T * (id + 1);
  int64_t l = 0, h = c->n;
  while (l < h) { int64_t m = (l + h) / 2; if (c->crow[m] < a) l = m + 1; else h = m; }
  *lo = id == 0 ? 0 : l;
  l = 0, h = c->n;
  while (l < h) { int64_t m = (l + h) / 2; if (c->crow[m] < b) l = m + 1; else h = m; }
  *hi = id == c->T - 1 ? c->n : l;
}

static void matvec_rows(const ctx_t* c, const double* x, double* y, int64_t lo, int64_t hi) {
  for (int64_t r = lo; r < hi; ++r) {
    double s = 0.0;
    for (int64_t j = c->crow[r]; j < c->crow[r + 1]; ++j) s += c->val[j] * x[c->col[j]];
    y[r] = s;
  }
}

static double combine(ctx_t* c, int id, double mine) {
  c->part[id][0] = mine;
  pthread_barrier_wait(&c->bar);
  double s = 0.0;
  for (int t = 0; t < c->T; ++t) s += c->part[t][0];
  pthread_barrier_wait(&c->bar); /* everyone has read the partials before they are overwritten */
  return s;
}

static void* matvec_worker(void* a_) {
  arg_t* a = (arg_t*)a_;
  int64_t lo, hi;
  rows_of(a->c, a->id, &lo, &hi);
  matvec_rows(a->c, a->c->x_in, a->c->y_out, lo, hi);
  return NULL;
}

static void* cg_worker(void* a_) {
  arg_t* a = (arg_t*)a_;
  ctx_t* c = a->c;
  const int id = a->id;
  int64_t lo, hi;
  rows_of(c, id, &lo, &hi);
  for (int64_t i = lo; i < hi; ++i)
    if (!c->freed[i]) c->u[i] = 0.0;
  pthread_barrier_wait(&c->bar);
  matvec_rows(c, c->u, c->Ap, lo, hi);
  double acc = 0.0;
  for (int64_t i = lo; i < hi; ++i) {
    c->r[i] = c->freed[i] ? c->F[i] - c->Ap[i] : 0.0;
    c->p[i] = c->r[i];
    acc += c->r[i] * c->r[i];
  }
  double rs_old = combine(c, id, acc);
  int it, status = 2, iterations = c->max_iter;
  for (it = 0; it < c->max_iter; ++it) {
    matvec_rows(c, c->p, c->Ap, lo, hi);
    acc = 0.0;
    for (int64_t i = lo; i < hi; ++i) acc += c->p[i] * c->Ap[i];
    const double pAp = combine(c, id, acc);
    const double alpha = rs_old / (pAp + c->eps);
    if (fabs(pAp) < c->eps || pAp < 0.0 || !isfinite(alpha)) { status = 1, iterations = it + 1; break; } /* solver.py:187-198 */
    acc = 0.0;
    for (int64_t i = lo; i < hi; ++i) {
      if (c->freed[i]) {
        c->u[i] += alpha * c->p[i];
        c->r[i] -= alpha * c->Ap[i];
      } else {
        c->u[i] = 0.0, c->r[i] = 0.0;
      }
      acc += c->r[i] * c->r[i];
    }
    const double rs_new = combine(c, id, acc);
    if (sqrt(rs_new) < c->tol) { status = 0, iterations = it + 1; break; } /* solver.py:210-212 */
    const double beta = rs_new / (rs_old + c->eps);
    if (!isfinite(beta)) { status = 1, iterations = it + 1; break; }
    for (int64_t i = lo; i < hi; ++i) c->p[i] = c->freed[i] ? c->r[i] + beta * c->p[i] : 0.0;
    rs_old = rs_new;
    pthread_barrier_wait(&c->bar); /* p complete before the next matvec gathers it */
  }
  if (id == 0) c->iterations = iterations, c->status = status;
  return NULL;
}

static void run(ctx_t* c, void* (*fn)(void*)) {
  pthread_t th[MAXT];
  arg_t args[MAXT];
  pthread_barrier_init(&c->bar, NULL, (unsigned)c->T);
  for (int t = 0; t < c->T; ++t) {
    args[t].c = c, args[t].id = t;
    pthread_create(&th[t], NULL, fn, &args[t]);
  }
  for (int t = 0; t < c->T; ++t) pthread_join(th[t], NULL);
  pthread_barrier_destroy(&c->bar);
}

void oracle_csr_matvec(int64_t n, const int64_t* crow, const int64_t* col, const double* val, const double* x, double* y) {
  ctx_t c;
  c.n = n, c.crow = crow, c.col = col, c.val = val, c.x_in = x, c.y_out = y;
  c.T = oracle_threads();
  if (c.T > n) c.T = n > 0 ? (int)n : 1;
  run(&c, matvec_worker);
}

/* freed[i] = 0 for dofs of fixed nodes.  Returns the iteration count the reference would print; *status: 0 converged,
 * 1 breakdown, 2 max_iter.  work = 3*n doubles (r, p, Ap). */
int oracle_cg_csr(int64_t n, const int64_t* crow, const int64_t* col, const double* val, const double* F, const unsigned char* freed,
                  double* u, double* work, double tol, int max_iter, double eps, int* status) {
  ctx_t c;
  c.n = n, c.crow = crow, c.col = col, c.val = val, c.F = F, c.freed = freed, c.u = u;
  c.r = work, c.p = work + n, c.Ap = work + 2 * n;
  c.tol = tol, c.eps = eps, c.max_iter = max_iter;
  c.T = oracle_threads();
  if (c.T > n) c.T = n > 0 ? (int)n : 1;
  c.iterations = max_iter, c.status = 2;
  run(&c, cg_worker);
  *status = c.status;
  return c.iterations;
}
