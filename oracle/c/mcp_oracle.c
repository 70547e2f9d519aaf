/*
 * mcp_oracle.c — TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C restatement of the reference's interior-point hot path, used (a) as a second oracle for
 * large-batch parity checks and (b) as the CPU baseline / `bench.py --impl reference` arm ("port": the
 * Julia reference cannot run in this image).  Only tests/, __graft_entry__.smoke() and bench.py may load
 * it; nothing under mcp_b200/ does.  PARITY PIN: checked against the reference's known answers through
 * tests/test_oracle.py (test/runtests.jl:30-38,112-115) and against the Python restatement
 * oracle/ip_oracle.py; beyond that "parity unpinned" (see that file's header).
 *
 * What follows what (all paths under /root/reference):
 *   solve_one()            src/solver.jl:35-122   (loop, tol·I shift, stale kkt_error, ϵ schedule, status)
 *   ftb_linesearch()       src/solver.jl:127-138
 *   eval_F / eval_J        src/mcp.jl:76-80, 97-120  (F = [G; H−s; s∘y−ϵ], CSC ∇F_z)
 *   splu_* (below)         stands in for UMFPACK via LinearSolve.jl (src/solver.jl:50,61,81-83): a
 *                          left-looking sparse LU with partial pivoting (Gilbert–Peierls), column
 *                          pre-ordering supplied by the caller (COLAMD from SuperLU, computed once per
 *                          pattern like UMFPACK's symbolic analysis).
 *   batch driver           OpenMP parallel-for over θ with one workspace per thread — the analogue of
 *                          Threads.@threads with a deepcopy(mcp) per thread (the reference's mcp holds a
 *                          shared result_buffer, src/solver.jl:53).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

enum { OP_CONST = 0, OP_X, OP_Y, OP_THETA, OP_ADD, OP_SUB, OP_MUL, OP_DIV, OP_NEG, OP_SQRT, OP_EXP, OP_LOG, OP_SIN, OP_COS, OP_POWI };

typedef struct {
  int32_t nx, ny, nt, n_nodes;
  const int32_t *op, *a, *b;
  const double* consts;
  const int32_t* gh_nodes;
  int32_t jz_nnz;
  const int32_t *jz_rows, *jz_cols, *jz_nodes;
  const int32_t* colperm; /* [n] fill-reducing column order, or NULL for natural */
  /* optional: F!/∇F_z! as COMPILED straight-line C generated from the tape (oracle/c_emit.py) — the analogue of the
   * reference's build_function output (src/mcp.jl:82-120); NULL ⇒ the tape is interpreted */
  void (*eval_fn)(const double* x, const double* y, const double* th, double* gh, double* jz, double* scratch /* n_nodes */);
} oracle_problem;

typedef struct {
  double tol;
  int32_t max_inner_iters, max_outer_iters;
  double tightening_rate, loosening_rate, min_stepsize;
} oracle_opts;

/* ---------------------------------------------------------------------------------------------- */
/* sparse LU, left-looking with partial pivoting                                                    */
/* ---------------------------------------------------------------------------------------------- */
typedef struct {
  int n;
  int *Lp, *Li, *Up, *Ui, *pinv;
  double *Lx, *Ux;
  int lcap, ucap;
  /* work */
  int *xi, *pstack, *mark;
  double* x;
} splu_t;

static void splu_init(splu_t* F, int n, int nnz) {
  F->n = n;
  F->lcap = F->ucap = 8 * nnz + 16 * n;
  F->Lp = malloc(sizeof(int) * (n + 1));
  F->Up = malloc(sizeof(int) * (n + 1));
  F->Li = malloc(sizeof(int) * F->lcap);
  F->Ui = malloc(sizeof(int) * F->ucap);
  F->Lx = malloc(sizeof(double) * F->lcap);
  F->Ux = malloc(sizeof(double) * F->ucap);
  F->pinv = malloc(sizeof(int) * n);
  F->xi = malloc(sizeof(int) * 2 * n);
  F->pstack = malloc(sizeof(int) * n);
  F->mark = malloc(sizeof(int) * n);
  F->x = calloc(n, sizeof(double));
}

static void splu_free(splu_t* F) {
  free(F->Lp); free(F->Up); free(F->Li); free(F->Ui); free(F->Lx); free(F->Ux);
  free(F->pinv); free(F->xi); free(F->pstack); free(F->mark); free(F->x);
}

static void grow(int** idx, double** val, int* cap, int need) {
  if (need <= *cap) return;
  while (*cap < need) *cap *= 2;
  *idx = realloc(*idx, sizeof(int) * *cap);
  *val = realloc(*val, sizeof(double) * *cap);
}

/* depth-first search of the graph of L from node j; L columns are indexed through pinv */
static int dfs(int j, const splu_t* F, int top, int* xi, int* pstack, int* mark, int stamp) {
  int head = 0;
  xi[0] = j;
  while (head >= 0) {
    j = xi[head];
    int jnew = F->pinv[j];
    if (mark[j] != stamp) {
      mark[j] = stamp;
      pstack[head] = (jnew < 0) ? 0 : F->Lp[jnew];
    }
    int done = 1;
    int p2 = (jnew < 0) ? 0 : F->Lp[jnew + 1];
    for (int p = pstack[head]; p < p2; ++p) {
      int i = F->Li[p];
      if (mark[i] == stamp) continue;
      pstack[head] = p;
      xi[++head] = i;
      done = 0;
      break;
    }
    if (done) {
      --head;
      xi[--top] = j;
    }
  }
  return top;
}

/* Factor A(:, q) = P' L U.  Returns 0 on success, 1 if structurally/numerically singular. */
static int splu_factor(splu_t* F, const int* Ap, const int* Ai, const double* Ax, const int* q) {
  const int n = F->n;
  int lnz = 0, unz = 0;
  int* xi = F->xi;
  double* x = F->x;
  for (int i = 0; i < n; ++i) { F->pinv[i] = -1; F->mark[i] = -1; x[i] = 0.0; }
  for (int k = 0; k < n; ++k) {
    F->Lp[k] = lnz;
    F->Up[k] = unz;
    grow(&F->Li, &F->Lx, &F->lcap, lnz + n);
    grow(&F->Ui, &F->Ux, &F->ucap, unz + n);
    const int col = q ? q[k] : k;
    /* reach of A(:,col) in the graph of L */
    int top = n;
    for (int p = Ap[col]; p < Ap[col + 1]; ++p)
      if (F->mark[Ai[p]] != k) top = dfs(Ai[p], F, top, xi, xi + n, F->mark, k);
    for (int p = top; p < n; ++p) x[xi[p]] = 0.0;
    for (int p = Ap[col]; p < Ap[col + 1]; ++p) x[Ai[p]] = Ax[p];
    /* x = L \ A(:,col) */
    for (int px = top; px < n; ++px) {
      const int j = xi[px];
      const int J = F->pinv[j];
      if (J < 0) continue;
      const double xj = x[j]; /* unit diagonal stored first */
      for (int p = F->Lp[J] + 1; p < F->Lp[J + 1]; ++p) x[F->Li[p]] -= F->Lx[p] * xj;
    }
    /* pivot: largest magnitude among rows not yet pivotal */
    int ipiv = -1;
    double a = -1.0;
    for (int p = top; p < n; ++p) {
      const int i = xi[p];
      if (F->pinv[i] < 0) {
        const double t = fabs(x[i]);
        if (t > a) { a = t; ipiv = i; }
      } else {
        F->Ui[unz] = F->pinv[i];
        F->Ux[unz++] = x[i];
      }
    }
    if (ipiv == -1 || !(a > 0.0) || !isfinite(a)) return 1;
    const double pivot = x[ipiv];
    F->Ui[unz] = k;
    F->Ux[unz++] = pivot;
    F->pinv[ipiv] = k;
    F->Li[lnz] = ipiv;
    F->Lx[lnz++] = 1.0;
    for (int p = top; p < n; ++p) {
      const int i = xi[p];
      if (F->pinv[i] < 0) {
        F->Li[lnz] = i;
        F->Lx[lnz++] = x[i] / pivot;
      }
      x[i] = 0.0;
    }
  }
  F->Lp[n] = lnz;
  F->Up[n] = unz;
  for (int p = 0; p < lnz; ++p) F->Li[p] = F->pinv[F->Li[p]];
  return 0;
}

/* solve A z = b using the factors; b is overwritten by work, result in out */
static void splu_solve(const splu_t* F, const int* q, const double* b, double* work, double* out) {
  const int n = F->n;
  for (int i = 0; i < n; ++i) work[F->pinv[i]] = b[i];
  for (int j = 0; j < n; ++j) { /* L: unit diagonal first in each column */
    const double xj = work[j];
    for (int p = F->Lp[j] + 1; p < F->Lp[j + 1]; ++p) work[F->Li[p]] -= F->Lx[p] * xj;
  }
  for (int j = n - 1; j >= 0; --j) { /* U: diagonal last in each column */
    work[j] /= F->Ux[F->Up[j + 1] - 1];
    const double xj = work[j];
    for (int p = F->Up[j]; p < F->Up[j + 1] - 1; ++p) work[F->Ui[p]] -= F->Ux[p] * xj;
  }
  for (int k = 0; k < n; ++k) out[q ? q[k] : k] = work[k];
}

/* ---------------------------------------------------------------------------------------------- */
/* problem set-up shared by all threads                                                             */
/* ---------------------------------------------------------------------------------------------- */
typedef struct {
  const oracle_problem* P;
  int n, nnzA;
  int *Ap, *Ai;        /* CSC pattern of ∇F + tol·I */
  int* slot_jz;        /* CSC slot of each IR Jacobian entry */
  int *slot_mI, *slot_S, *slot_Y, *slot_diag;
  int* need;           /* tape nodes to evaluate, in order */
  int n_need;
} setup_t;

typedef struct { int r, c, src; } trip_t;
static int trip_cmp(const void* a, const void* b) {
  const trip_t *x = a, *y = b;
  if (x->c != y->c) return x->c - y->c;
  return x->r - y->r;
}

static void setup_build(setup_t* S, const oracle_problem* P) {
  const int nx = P->nx, ny = P->ny, n = nx + 2 * ny;
  S->P = P;
  S->n = n;
  const int ntrip = P->jz_nnz + 3 * ny + n;
  trip_t* T = malloc(sizeof(trip_t) * ntrip);
  int t = 0;
  for (int k = 0; k < P->jz_nnz; ++k) T[t++] = (trip_t){P->jz_rows[k], P->jz_cols[k], k};
  for (int k = 0; k < ny; ++k) T[t++] = (trip_t){nx + k, nx + ny + k, -1 - 4 * k - 0};      /* −I        */
  for (int k = 0; k < ny; ++k) T[t++] = (trip_t){nx + ny + k, nx + k, -1 - 4 * k - 1};      /* diag(s)   */
  for (int k = 0; k < ny; ++k) T[t++] = (trip_t){nx + ny + k, nx + ny + k, -1 - 4 * k - 2}; /* diag(y)   */
  for (int i = 0; i < n; ++i) T[t++] = (trip_t){i, i, INT32_MIN + i};                       /* tol·I     */
  qsort(T, ntrip, sizeof(trip_t), trip_cmp);
  S->Ap = calloc(n + 1, sizeof(int));
  S->Ai = malloc(sizeof(int) * ntrip);
  S->slot_jz = malloc(sizeof(int) * (P->jz_nnz + 1));
  S->slot_mI = malloc(sizeof(int) * (ny + 1));
  S->slot_S = malloc(sizeof(int) * (ny + 1));
  S->slot_Y = malloc(sizeof(int) * (ny + 1));
  S->slot_diag = malloc(sizeof(int) * n);
  int nnz = 0;
  for (int i = 0; i < ntrip; ++i) {
    if (i == 0 || T[i].r != T[i - 1].r || T[i].c != T[i - 1].c) {
      S->Ai[nnz] = T[i].r;
      S->Ap[T[i].c + 1]++;
      ++nnz;
    }
    const int slot = nnz - 1, src = T[i].src;
    if (src >= 0) S->slot_jz[src] = slot;
    else if (src < -(1 << 30)) S->slot_diag[src - INT32_MIN] = slot;
    else {
      const int k = (-1 - src) / 4, w = (-1 - src) % 4;
      if (w == 0) S->slot_mI[k] = slot; else if (w == 1) S->slot_S[k] = slot; else S->slot_Y[k] = slot;
    }
  }
  for (int c = 0; c < n; ++c) S->Ap[c + 1] += S->Ap[c];
  S->nnzA = nnz;
  free(T);
  /* nodes needed by G, H and the Jacobian entries */
  char* mark = calloc(P->n_nodes, 1);
  for (int i = 0; i < nx + ny; ++i) mark[P->gh_nodes[i]] = 1;
  for (int k = 0; k < P->jz_nnz; ++k) mark[P->jz_nodes[k]] = 1;
  for (int v = P->n_nodes - 1; v >= 0; --v) {
    if (!mark[v]) continue;
    const int op = P->op[v];
    if (op >= OP_ADD && op <= OP_DIV) { mark[P->a[v]] = 1; mark[P->b[v]] = 1; }
    else if (op >= OP_NEG) mark[P->a[v]] = 1;
  }
  S->need = malloc(sizeof(int) * P->n_nodes);
  S->n_need = 0;
  for (int v = 0; v < P->n_nodes; ++v) if (mark[v]) S->need[S->n_need++] = v;
  free(mark);
}

static void setup_free(setup_t* S) {
  free(S->Ap); free(S->Ai); free(S->slot_jz); free(S->slot_mI); free(S->slot_S); free(S->slot_Y);
  free(S->slot_diag); free(S->need);
}

static double powi(double a, int n) {
  double r = 1.0;
  int neg = n < 0;
  if (neg) n = -n;
  while (n) { if (n & 1) r *= a; a *= a; n >>= 1; }
  return neg ? 1.0 / r : r;
}

/* the compiled callables F! and ∇F_z! of the reference evaluate this tape (src/mcp.jl:82-120) */
static void eval_tape(const setup_t* S, const double* x, const double* y, const double* th, double* v) {
  const oracle_problem* P = S->P;
  for (int i = 0; i < S->n_need; ++i) {
    const int n = S->need[i], a = P->a[n], b = P->b[n];
    switch (P->op[n]) {
      case OP_CONST: v[n] = P->consts[a]; break;
      case OP_X: v[n] = x[a]; break;
      case OP_Y: v[n] = y[a]; break;
      case OP_THETA: v[n] = th[a]; break;
      case OP_ADD: v[n] = v[a] + v[b]; break;
      case OP_SUB: v[n] = v[a] - v[b]; break;
      case OP_MUL: v[n] = v[a] * v[b]; break;
      case OP_DIV: v[n] = v[a] / v[b]; break;
      case OP_NEG: v[n] = -v[a]; break;
      case OP_SQRT: v[n] = sqrt(v[a]); break;
      case OP_EXP: v[n] = exp(v[a]); break;
      case OP_LOG: v[n] = log(v[a]); break;
      case OP_SIN: v[n] = sin(v[a]); break;
      case OP_COS: v[n] = cos(v[a]); break;
      case OP_POWI: v[n] = powi(v[a], b); break;
      default: v[n] = NAN;
    }
  }
}

/* src/solver.jl:127-138 */
static double ftb_linesearch(const double* v, const double* d, int n, double tol) {
  const double tau = 0.995, decay = 0.5, c = 1 - tau;
  double alpha = 1.0;
  for (;;) {
    int any = 0;
    for (int i = 0; i < n; ++i) if (v[i] + alpha * d[i] < c * v[i]) { any = 1; break; } /* :129 */
    if (!any) return alpha;
    if (alpha < tol) return NAN; /* :130-131 */
    alpha *= decay;              /* :134 */
    if (alpha == 0.0) return NAN; /* guard: the reference would spin forever for tol <= 0 */
  }
}

typedef struct {
  splu_t lu;
  double *vals, *Ax, *F, *dz, *work, *x, *y, *s, *gh, *jz;
} thread_ws;

static void ws_init(thread_ws* W, const setup_t* S) {
  const oracle_problem* P = S->P;
  splu_init(&W->lu, S->n, S->nnzA);
  W->vals = malloc(sizeof(double) * P->n_nodes);
  W->Ax = malloc(sizeof(double) * S->nnzA);
  W->F = malloc(sizeof(double) * S->n);
  W->dz = malloc(sizeof(double) * S->n);
  W->work = malloc(sizeof(double) * S->n);
  W->x = malloc(sizeof(double) * (P->nx + 1));
  W->y = malloc(sizeof(double) * (P->ny + 1));
  W->s = malloc(sizeof(double) * (P->ny + 1));
  W->gh = malloc(sizeof(double) * (P->nx + P->ny + 1));
  W->jz = malloc(sizeof(double) * (P->jz_nnz + 1));
}

static void ws_free(thread_ws* W) {
  splu_free(&W->lu);
  free(W->vals); free(W->Ax); free(W->F); free(W->dz); free(W->work); free(W->x); free(W->y); free(W->s);
  free(W->gh); free(W->jz);
}

/* `solve(::InteriorPoint, mcp, θ; …)` — src/solver.jl:35-122 */
static void solve_one(const setup_t* S, thread_ws* W, const oracle_opts* o, const double* th, const double* x0,
                      const double* y0, const double* s0, double* x_out, double* y_out, double* s_out,
                      double* kkt_out, double* eps_out, int32_t* outer_out, int32_t* status_out, int32_t* steps_out) {
  const oracle_problem* P = S->P;
  const int nx = P->nx, ny = P->ny, n = S->n;
  double *x = W->x, *y = W->y, *s = W->s, *F = W->F, *dz = W->dz;
  for (int i = 0; i < nx; ++i) x[i] = x0 ? x0[i] : 0.0; /* :39 */
  for (int i = 0; i < ny; ++i) { y[i] = y0 ? y0[i] : 1.0; s[i] = s0 ? s0[i] : 1.0; } /* :40-41 */
  const double tol = o->tol;
  double eps = 1.0, kkt = INFINITY; /* :67-68 */
  int status = 0, outer = 1, steps = 0; /* :69-70 */
  while (kkt > tol && eps > tol && outer < o->max_outer_iters) { /* :71 */
    int inner = 1; /* :72 */
    status = 0;    /* :73 */
    while (kkt > eps && inner < o->max_inner_iters) { /* :75 */
      if (P->eval_fn) {
        P->eval_fn(x, y, th, W->gh, W->jz, W->vals);
      } else {
        eval_tape(S, x, y, th, W->vals);
        for (int i = 0; i < nx + ny; ++i) W->gh[i] = W->vals[P->gh_nodes[i]];
        for (int k = 0; k < P->jz_nnz; ++k) W->jz[k] = W->vals[P->jz_nodes[k]];
      }
      /* F = [G; H − s; s∘y − ϵ]   (:79, src/mcp.jl:76-80) */
      for (int i = 0; i < nx; ++i) F[i] = W->gh[i];
      for (int k = 0; k < ny; ++k) {
        F[nx + k] = W->gh[nx + k] - s[k];
        F[nx + ny + k] = s[k] * y[k] - eps;
      }
      /* A = ∇F + tol·I   (:80-81) */
      memset(W->Ax, 0, sizeof(double) * S->nnzA);
      for (int k = 0; k < P->jz_nnz; ++k) W->Ax[S->slot_jz[k]] += W->jz[k];
      for (int k = 0; k < ny; ++k) {
        W->Ax[S->slot_mI[k]] += -1.0;
        W->Ax[S->slot_S[k]] += s[k];
        W->Ax[S->slot_Y[k]] += y[k];
      }
      for (int i = 0; i < n; ++i) W->Ax[S->slot_diag[i]] += tol;
      /* δz = A \ (−F)   (:82-83) */
      if (splu_factor(&W->lu, S->Ap, S->Ai, W->Ax, P->colperm)) { status = 1; break; } /* :84-88 */
      for (int i = 0; i < n; ++i) F[i] = -F[i];
      splu_solve(&W->lu, P->colperm, F, W->work, dz);
      const double a_s = ftb_linesearch(s, dz + nx + ny, ny, o->min_stepsize); /* :93 */
      const double a_y = ftb_linesearch(y, dz + nx, ny, o->min_stepsize);      /* :94 */
      if (isnan(a_s) || isnan(a_y)) { status = 1; break; }                     /* :96-100 */
      for (int i = 0; i < nx; ++i) x[i] += a_s * dz[i];                         /* :103 */
      for (int k = 0; k < ny; ++k) { s[k] += a_s * dz[nx + ny + k]; y[k] += a_y * dz[nx + k]; } /* :104-105 */
      double m = 0.0; /* :107 norm(F, Inf), NaN-propagating */
      for (int i = 0; i < n; ++i) { const double f = fabs(F[i]); if (isnan(f)) { m = NAN; break; } if (f > m) m = f; }
      kkt = m;
      ++inner; /* :108 */
      ++steps;
    }
    eps *= (status == 0) ? 1.0 - exp(-o->tightening_rate * inner) : 1.0 + exp(-o->loosening_rate * inner); /* :111-113 */
    ++outer; /* :114 */
  }
  if (outer == o->max_outer_iters) status = 1; /* :117-119 */
  memcpy(x_out, x, sizeof(double) * nx);
  memcpy(y_out, y, sizeof(double) * ny);
  memcpy(s_out, s, sizeof(double) * ny);
  *kkt_out = kkt; *eps_out = eps; *outer_out = outer; *status_out = status;
  if (steps_out) *steps_out = steps;
}

/* Batched driver.  theta is nt×B column-major; x0/y0/s0 may be NULL.  Returns the number of threads used. */
int mcp_oracle_solve_batch(const oracle_problem* P, int64_t B, const double* theta, const double* x0, const double* y0,
                           const double* s0, const oracle_opts* opts, double* x_out, double* y_out, double* s_out,
                           double* kkt_out, double* eps_out, int32_t* outer_out, int32_t* status_out,
                           int32_t* steps_out, int nthreads) {
  setup_t S;
  setup_build(&S, P);
  int used = 1;
#ifdef _OPENMP
  if (nthreads <= 0) nthreads = omp_get_max_threads();
  used = nthreads;
#pragma omp parallel num_threads(nthreads)
#endif
  {
    thread_ws W;
    ws_init(&W, &S);
#ifdef _OPENMP
#pragma omp for schedule(dynamic, 1)
#endif
    for (int64_t b = 0; b < B; ++b) {
      solve_one(&S, &W, opts, theta + b * P->nt, x0 ? x0 + b * P->nx : NULL, y0 ? y0 + b * P->ny : NULL,
                s0 ? s0 + b * P->ny : NULL, x_out + b * P->nx, y_out + b * P->ny, s_out + b * P->ny, kkt_out + b,
                eps_out + b, outer_out + b, status_out + b, steps_out ? steps_out + b : NULL);
    }
    ws_free(&W);
  }
  setup_free(&S);
  return used;
}

int mcp_oracle_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
