/* CPU ORACLE — TEST INFRASTRUCTURE ONLY (see cask_oracle.h for the contract and the pinning).
 * Plain C restatement of the reference algorithms; each function cites the reference lines
 * (relative to /root/reference) it follows.  Built with -ffp-contract=off so that a*b+c is a
 * separate multiply and add, as in the reference's scalar C++.
 */
#include "cask_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define EMPTY_FLAG ((int32_t)(1u << 31))

static int32_t round_up(int32_t v, int32_t to) { return (v % to == 0) ? v : (v / to + 1) * to; }

/* src/runtime/Spmv.cpp:25-40 — an input_width-lane reader whose lane cursor carries across rows;
 * every row, even an empty one, costs at least one cycle. */
int32_t oracle_count_compute_cycles(const int32_t* v, int32_t size, int32_t w) {
  int32_t cycles = 0, lane = 0;
  for (int32_t i = 0; i < size; i++) {
    int32_t left = v[i] - (i ? v[i - 1] : 0);
    for (;;) {
      int32_t take = w - lane < left ? w - lane : left;
      lane = (lane + take) % w;
      cycles++;
      left -= take;
      if (left <= 0) break;
    }
  }
  return cycles;
}

/* src/runtime/Spmv.hpp:213-238 — maximal runs of k empty rows become one entry k|1<<31;
 * non-empty rows keep their cumulative end offset. */
int32_t oracle_encode_empty_rows(const int32_t* v, int32_t size, int32_t* out) {
  int32_t len = 0, run = 0;
  for (int32_t i = 0; i < size; i++) {
    int32_t rowlen = v[i] - (i ? v[i - 1] : 0);
    if (rowlen == 0) { run++; continue; }
    if (run) out[len++] = run | EMPTY_FLAG;
    run = 0;
    out[len++] = v[i];
  }
  if (run) out[len++] = run | EMPTY_FLAG;
  return len;
}

/* src/runtime/SparseMatrix.hpp:426-443 (row stripe, columns untouched), :459-482 (column blocks:
 * n cumulative END offsets per block, block-local column index, original order kept) and
 * src/runtime/Spmv.cpp:42-107 (concatenate blocks, pad pairs to input_width, cycle model). */
int oracle_do_blocking(int32_t m, const int32_t* row_ptr, const int32_t* col_ind, const double* values,
                       int32_t start, int32_t n, int32_t bs, int32_t w, int arch, oracle_partition* out) {
  memset(out, 0, sizeof(*out));
  if (bs <= 0 || w <= 0 || m <= 0 || n < 0) return -1;
  int32_t nBlocks = m / bs + (m % bs == 0 ? 0 : 1);
  size_t cells = (size_t)nBlocks * (size_t)(n > 0 ? n : 1);
  int32_t* endoff = (int32_t*)calloc(cells, sizeof(int32_t)); /* [b][i]: count, then cumulative */
  if (!endoff) return -2;
  for (int32_t i = 0; i < n; i++)
    for (int32_t k = row_ptr[start + i]; k < row_ptr[start + i + 1]; k++)
      endoff[(size_t)(col_ind[k] / bs) * n + i]++;
  int64_t* pair_base = (int64_t*)calloc((size_t)nBlocks + 1, sizeof(int64_t));
  for (int32_t b = 0; b < nBlocks; b++) {
    int32_t* e = endoff + (size_t)b * n;
    for (int32_t i = 1; i < n; i++) e[i] += e[i - 1];
    int32_t nnz_b = n ? e[n - 1] : 0;
    pair_base[b + 1] = pair_base[b] + round_up(nnz_b, w); /* Utils.hpp:61-68 via Spmv.cpp:81-82 */
  }
  int32_t* colptr = (int32_t*)malloc(sizeof(int32_t) * (cells ? cells : 1));
  int64_t ncolptr = 0;
  int32_t cycles = 0, reduction = n * nBlocks, empty = 0; /* Spmv.cpp:66-69 */
  for (int32_t b = 0; b < nBlocks; b++) {
    const int32_t* e = endoff + (size_t)b * n;
    int32_t len;
    int encode = arch == ORACLE_ARCH_SKIPEMPTY && b != 0 && b != nBlocks - 1; /* Spmv.hpp:244 */
    if (encode) len = oracle_encode_empty_rows(e, n, colptr + ncolptr);
    else { memcpy(colptr + ncolptr, e, sizeof(int32_t) * (size_t)n); len = n; }
    int32_t diff = n - len; /* Spmv.cpp:76-79 */
    empty += diff;
    reduction -= diff;
    cycles += oracle_count_compute_cycles(e, n, w) - diff;
    ncolptr += len;
  }
  int64_t npairs = pair_base[nBlocks];
  oracle_pair* pairs = (oracle_pair*)calloc((size_t)(npairs ? npairs : 1), sizeof(oracle_pair));
  int64_t* cursor = (int64_t*)malloc(sizeof(int64_t) * (size_t)(nBlocks ? nBlocks : 1));
  for (int32_t b = 0; b < nBlocks; b++) cursor[b] = pair_base[b];
  for (int32_t i = 0; i < n; i++)
    for (int32_t k = row_ptr[start + i]; k < row_ptr[start + i + 1]; k++) {
      int32_t b = col_ind[k] / bs;
      oracle_pair* p = &pairs[cursor[b]++];
      p->value = values[k];
      p->indptr = col_ind[k] - b * bs; /* SparseMatrix.hpp:475 */
    }
  int32_t out_len = n ? round_up(n, 384 / 8) : 0; /* Spmv.cpp:91-92, burst = 384 B */
  int32_t v_len = round_up(m, bs);                 /* Spmv.cpp:93 */
  out->nBlocks = nBlocks;
  out->n = n;
  out->paddingCycles = out_len - n;
  out->totalCycles = cycles + v_len;
  out->vector_load_cycles = v_len / nBlocks;
  out->outSize = out_len * 8;
  out->emptyCycles = empty;
  out->reductionCycles = reduction;
  out->m_colptr_unpaddedLength = (int32_t)ncolptr;
  out->m_indptr_values_unpaddedLength = (int32_t)npairs;
  out->len_colptr = ncolptr;
  out->len_pairs = npairs;
  out->m_colptr = colptr;
  out->m_indptr_values = pairs;
  free(endoff); free(pair_base); free(cursor);
  return 0;
}

static void copy_partition(oracle_partition* d, const oracle_partition* s, int zero_values) {
  *d = *s;
  d->m_colptr = (int32_t*)malloc(sizeof(int32_t) * (size_t)(s->len_colptr ? s->len_colptr : 1));
  memcpy(d->m_colptr, s->m_colptr, sizeof(int32_t) * (size_t)s->len_colptr);
  d->m_indptr_values = (oracle_pair*)malloc(sizeof(oracle_pair) * (size_t)(s->len_pairs ? s->len_pairs : 1));
  memcpy(d->m_indptr_values, s->m_indptr_values, sizeof(oracle_pair) * (size_t)s->len_pairs);
  if (zero_values)
    for (int64_t i = 0; i < d->len_pairs; i++) d->m_indptr_values[i].value = 0;
}

/* src/runtime/Spmv.cpp:329-365 */
int oracle_preprocess(int32_t n, int32_t m, const int32_t* row_ptr, const int32_t* col_ind,
                      const double* values, int arch, int32_t P, int32_t cache, int32_t w,
                      oracle_partition* parts) {
  if (P <= 0) return -1;
  int32_t rpp = n / P;
  if (rpp == 0) { /* fewer rows than pipes: everyone gets the whole matrix, values zeroed on 1.. */
    int rc = oracle_do_blocking(m, row_ptr, col_ind, values, 0, n, cache, w, arch, &parts[0]);
    if (rc) return rc;
    for (int32_t p = 1; p < P; p++) copy_partition(&parts[p], &parts[0], 1);
    return 0;
  }
  int32_t start = 0;
  for (int32_t p = 0; p < P; p++) {
    int32_t rows = p == P - 1 ? n - start : rpp;
    int rc = oracle_do_blocking(m, row_ptr, col_ind, values, start, rows, cache, w, arch, &parts[p]);
    if (rc) return rc;
    start += rows;
  }
  return 0;
}

void oracle_free_partition(oracle_partition* p) {
  free(p->m_colptr);
  free(p->m_indptr_values);
  p->m_colptr = NULL;
  p->m_indptr_values = NULL;
}

/* src/runtime/SparseMatrix.hpp:255-264: rows independent; inside a row std::map order = ascending
 * column; result[row] += b[col] * value. Rows whose columns are already ascending are summed in
 * place, others through an insertion-sorted copy. */
void oracle_csr_dot(int32_t n, const int32_t* rp, const int32_t* ci, const double* va, const double* x,
                    double* y) {
  for (int32_t i = 0; i < n; i++) {
    int32_t b = rp[i], e = rp[i + 1];
    int sorted = 1;
    for (int32_t k = b + 1; k < e; k++)
      if (ci[k] <= ci[k - 1]) { sorted = 0; break; }
    double acc = 0.0;
    if (sorted) {
      for (int32_t k = b; k < e; k++) acc += x[ci[k]] * va[k];
    } else {
      int32_t len = e - b;
      int32_t* c = (int32_t*)malloc(sizeof(int32_t) * (size_t)len);
      double* v = (double*)malloc(sizeof(double) * (size_t)len);
      int32_t cnt = 0;
      for (int32_t k = b; k < e; k++) { /* dok[i][col] = value: later duplicates overwrite */
        int32_t pos = 0;
        while (pos < cnt && c[pos] < ci[k]) pos++;
        if (pos < cnt && c[pos] == ci[k]) { v[pos] = va[k]; continue; }
        for (int32_t q = cnt; q > pos; q--) { c[q] = c[q - 1]; v[q] = v[q - 1]; }
        c[pos] = ci[k]; v[pos] = va[k]; cnt++;
      }
      for (int32_t k = 0; k < cnt; k++) acc += x[c[k]] * v[k];
      free(c); free(v);
    }
    y[i] = acc;
  }
}

/* SURVEY.md section 3.3 (ParallelCsrReadControl.java:160-196, SpmvKernel.java:61-78,286-297):
 * what a device computes from the partition arrays.  input_width is recovered from the pair
 * stream: each block's pairs are padded to a multiple of it, so block bases are given by the
 * caller through the colptr walk plus `w`. */
static int partition_spmv_one(const oracle_partition* p, int32_t cache, int32_t w, const double* x,
                              int32_t m, double* y) {
  int64_t cp = 0, pr = 0;
  double* xc = (double*)malloc(sizeof(double) * (size_t)cache);
  for (int32_t b = 0; b < p->nBlocks; b++) {
    for (int32_t j = 0; j < cache; j++) {
      int64_t col = (int64_t)b * cache + j;
      xc[j] = col < m ? x[col] : 0.0; /* Spmv.cpp:211-213 pads x with zeros */
    }
    int32_t row = 0, prev = 0;
    int64_t blk_pairs = 0;
    while (row < p->n) {
      if (cp >= p->len_colptr) { free(xc); return -1; }
      int32_t e = p->m_colptr[cp++];
      if (e & EMPTY_FLAG) {
        int32_t skip = e & 0x7fffffff;
        if (b == 0) for (int32_t r = 0; r < skip && row + r < p->n; r++) y[row + r] = 0.0;
        row += skip;
        continue;
      }
      int32_t len = e - prev;
      prev = e;
      double acc = 0.0;
      for (int32_t k = 0; k < len; k++) {
        const oracle_pair* q = &p->m_indptr_values[pr + blk_pairs + k];
        acc += q->value * xc[q->indptr];
      }
      blk_pairs += len;
      if (b == 0) y[row] = acc; else y[row] += acc; /* first block overwrites, later accumulate */
      row++;
    }
    pr += round_up((int32_t)blk_pairs, w);
  }
  free(xc);
  return (cp == p->len_colptr && pr == p->len_pairs) ? 0 : -2;
}

int oracle_partition_spmv_w(const oracle_partition* parts, int32_t nparts, int32_t cache, int32_t w,
                            const double* x, int32_t m, double* y, int32_t n_total) {
  int64_t total = 0;
  for (int32_t p = 0; p < nparts; p++) total += parts[p].n;
  double* cat = (double*)calloc((size_t)(total ? total : 1), sizeof(double));
  int64_t off = 0;
  int rc = 0;
  for (int32_t p = 0; p < nparts && !rc; p++) {
    rc = partition_spmv_one(&parts[p], cache, w, x, m, cat + off);
    off += parts[p].n;
  }
  /* Spmv.cpp:303-326: concatenate nrows[i] doubles per pipe, then drop filler rows */
  for (int32_t i = 0; i < n_total; i++) y[i] = i < total ? cat[i] : 0.0;
  free(cat);
  return rc;
}

/* ---- solvers --------------------------------------------------------------------------- */
static void symv_lower(int32_t n, const int32_t* rp, const int32_t* ci, const double* va,
                       const double* x, double* y) {
  /* mkl_dcsrsymv('l', ...) at SparseLinearSolvers.hpp:189,206: y = A x, A symmetric, only the
   * lower triangle stored.  MKL's internal summation order is unknowable; restated row by row. */
  for (int32_t i = 0; i < n; i++) y[i] = 0.0;
  for (int32_t i = 0; i < n; i++)
    for (int32_t k = rp[i]; k < rp[i + 1]; k++) {
      int32_t j = ci[k];
      if (j > i) continue; /* 'l': entries above the diagonal are ignored */
      y[i] += va[k] * x[j];
      if (j != i) y[j] += va[k] * x[i];
    }
}

static void gemv_full(int32_t n, const int32_t* rp, const int32_t* ci, const double* va,
                      const double* x, double* y) {
  for (int32_t i = 0; i < n; i++) {
    double acc = 0.0;
    for (int32_t k = rp[i]; k < rp[i + 1]; k++) acc += va[k] * x[ci[k]];
    y[i] = acc;
  }
}

static double ddot(int32_t n, const double* a, const double* b) {
  double s = 0.0;
  for (int32_t i = 0; i < n; i++) s += a[i] * b[i];
  return s;
}

typedef void (*matvec_fn)(int32_t, const int32_t*, const int32_t*, const double*, const double*, double*);

/* src/runtime/SparseLinearSolvers.hpp:162-239 with Precon = IdentityPreconditioner (z = r). */
static int pcg_impl(matvec_fn mv, int32_t n, const int32_t* rp, const int32_t* ci, const double* va,
                    const double* rhs, double* x, int32_t* iterations, int32_t maxiters, double tol,
                    double* rs_final) {
  double* r = (double*)malloc(sizeof(double) * (size_t)n);
  double* p = (double*)malloc(sizeof(double) * (size_t)n);
  double* Ap = (double*)malloc(sizeof(double) * (size_t)n);
  int converged = 0;
  mv(n, rp, ci, va, x, r);                                  /* :189  r = A x      */
  for (int32_t i = 0; i < n; i++) r[i] = rhs[i] - r[i];     /* :190  r = b - r    */
  memcpy(p, r, sizeof(double) * (size_t)n);                 /* :193-195 z = r; p = z */
  double rsold = ddot(n, r, r);                             /* :198 */
  double rsnew = rsold;
  for (int32_t it = 0; it < maxiters; it++) {               /* :200 */
    mv(n, rp, ci, va, p, Ap);                               /* :206 */
    double alpha = rsold / ddot(n, p, Ap);                  /* :208 */
    for (int32_t i = 0; i < n; i++) x[i] += alpha * p[i];   /* :210 */
    for (int32_t i = 0; i < n; i++) r[i] = -alpha * Ap[i] + r[i]; /* :212 daxpby(-alpha, Ap, 1, r) */
    rsnew = ddot(n, r, r);                                  /* :215-218 */
    if (rsnew <= tol * tol) { converged = 1; break; }       /* :220-226 absolute test */
    double beta = rsnew / rsold;
    for (int32_t i = 0; i < n; i++) p[i] = r[i] + beta * p[i]; /* :229 */
    rsold = rsnew;                                          /* :230 */
    *iterations = it;                                       /* :231 — only when not converged */
  }
  if (rs_final) *rs_final = rsnew;
  free(r); free(p); free(Ap);
  return converged;
}

int oracle_pcg(int32_t n, const int32_t* rp, const int32_t* ci, const double* va, const double* rhs,
               double* x, int32_t* iterations, int32_t maxiters, double tol) {
  return pcg_impl(symv_lower, n, rp, ci, va, rhs, x, iterations, maxiters, tol, NULL);
}

int oracle_pcg_full(int32_t n, const int32_t* rp, const int32_t* ci, const double* va,
                    const double* rhs, double* x, int32_t* iterations, int32_t maxiters, double tol,
                    double* rs_final) {
  return pcg_impl(gemv_full, n, rp, ci, va, rhs, x, iterations, maxiters, tol, rs_final);
}

/* Eigen 3.3.1, Eigen/src/IterativeLinearSolvers/BiCGSTAB.h, bicgstab() with the default
 * DiagonalPreconditioner (BasicPreconditioners.h: invdiag = 1/a_jj, 1 where a_jj == 0); call sites
 * src/runtime/SparseLinearSolvers.cpp:18-26,62-67.  PARITY UNPINNED (Eigen absent; no reference test). */
int oracle_bicgstab(int32_t n, const int32_t* rp, const int32_t* ci, const double* va, const double* b,
                    double* x, int32_t* iters, double* tol_error) {
  const double tol = *tol_error;
  const int32_t maxit = *iters;
  size_t bytes = sizeof(double) * (size_t)n;
  double *r = malloc(bytes), *r0 = malloc(bytes), *v = calloc(n, sizeof(double)),
         *p = calloc(n, sizeof(double)), *y = malloc(bytes), *z = malloc(bytes), *s = malloc(bytes),
         *t = malloc(bytes), *invd = malloc(bytes);
  for (int32_t i = 0; i < n; i++) {
    double d = 0.0;
    for (int32_t k = rp[i]; k < rp[i + 1]; k++) if (ci[k] == i) d = va[k];
    invd[i] = d != 0.0 ? 1.0 / d : 1.0;
  }
  gemv_full(n, rp, ci, va, x, r);
  for (int32_t i = 0; i < n; i++) r[i] = b[i] - r[i];
  memcpy(r0, r, bytes);
  double r0_sq = ddot(n, r0, r0), rhs_sq = ddot(n, b, b);
  if (rhs_sq == 0.0) {
    memset(x, 0, bytes);
    *iters = 0; *tol_error = 0.0;
    goto done;
  }
  {
    double rho = 1, alpha = 1, w = 1;
    double tol2 = tol * tol * rhs_sq;
    double eps2 = 2.220446049250313e-16 * 2.220446049250313e-16;
    int32_t i = 0, restarts = 0;
    while (ddot(n, r, r) > tol2 && i < maxit) {
      double rho_old = rho;
      rho = ddot(n, r0, r);
      if (fabs(rho) < eps2 * r0_sq) {
        gemv_full(n, rp, ci, va, x, r);
        for (int32_t k = 0; k < n; k++) r[k] = b[k] - r[k];
        memcpy(r0, r, bytes);
        rho = r0_sq = ddot(n, r, r);
        if (restarts++ == 0) i = 0;
      }
      double beta = (rho / rho_old) * (alpha / w);
      for (int32_t k = 0; k < n; k++) p[k] = r[k] + beta * (p[k] - w * v[k]);
      for (int32_t k = 0; k < n; k++) y[k] = invd[k] * p[k];
      gemv_full(n, rp, ci, va, y, v);
      alpha = rho / ddot(n, r0, v);
      for (int32_t k = 0; k < n; k++) s[k] = r[k] - alpha * v[k];
      for (int32_t k = 0; k < n; k++) z[k] = invd[k] * s[k];
      gemv_full(n, rp, ci, va, z, t);
      double tt = ddot(n, t, t);
      w = tt > 0.0 ? ddot(n, t, s) / tt : 0.0;
      for (int32_t k = 0; k < n; k++) x[k] += alpha * y[k] + w * z[k];
      for (int32_t k = 0; k < n; k++) r[k] = s[k] - w * t[k];
      ++i;
    }
    *tol_error = sqrt(ddot(n, r, r) / rhs_sq);
    *iters = i;
  }
done:
  free(r); free(r0); free(v); free(p); free(y); free(z); free(s); free(t); free(invd);
  return 1;
}

/* ---- synthetic matrices ------------------------------------------- */
int64_t oracle_gen_poisson2d(int32_t N, int32_t* rp, int32_t* ci, double* va) {
  int64_t nnz = 0;
  for (int32_t i = 0; i < N; i++)
    for (int32_t j = 0; j < N; j++) {
      int32_t row = i * N + j;
      if (rp) rp[row] = (int32_t)nnz;
#define EMIT(c, v) do { if (rp) { ci[nnz] = (c); va[nnz] = (v); } nnz++; } while (0)
      if (i > 0) EMIT(row - N, -1.0);
      if (j > 0) EMIT(row - 1, -1.0);
      EMIT(row, 4.0);
      if (j < N - 1) EMIT(row + 1, -1.0);
      if (i < N - 1) EMIT(row + N, -1.0);
    }
  if (rp) rp[(int64_t)N * N] = (int32_t)nnz;
  return nnz;
}

int64_t oracle_gen_poisson3d27(int32_t N, int32_t* rp, int32_t* ci, double* va) {
  int64_t nnz = 0;
  for (int32_t z = 0; z < N; z++)
    for (int32_t y = 0; y < N; y++)
      for (int32_t x = 0; x < N; x++) {
        int32_t row = (z * N + y) * N + x;
        if (rp) rp[row] = (int32_t)nnz;
        for (int32_t dz = -1; dz <= 1; dz++)
          for (int32_t dy = -1; dy <= 1; dy++)
            for (int32_t dx = -1; dx <= 1; dx++) {
              int32_t zz = z + dz, yy = y + dy, xx = x + dx;
              if (zz < 0 || zz >= N || yy < 0 || yy >= N || xx < 0 || xx >= N) continue;
              EMIT((zz * N + yy) * N + xx, (dz == 0 && dy == 0 && dx == 0) ? 26.0 : -1.0);
            }
      }
  if (rp) rp[(int64_t)N * N * N] = (int32_t)nnz;
  return nnz;
}

/* 7-point diffusion (6,-1) + first-order upwind convection, cell Peclet 0.5 * (1, 0.5, 0.25):
 * the upwind (minus-side) neighbours get -1-c, the diagonal 6+cx+cy+cz. Nonsymmetric M-matrix. */
int64_t oracle_gen_convdiff3d7(int32_t N, int32_t* rp, int32_t* ci, double* va) {
  const double cx = 0.5, cy = 0.25, cz = 0.125;
  int64_t nnz = 0;
  for (int32_t z = 0; z < N; z++)
    for (int32_t y = 0; y < N; y++)
      for (int32_t x = 0; x < N; x++) {
        int32_t row = (z * N + y) * N + x;
        if (rp) rp[row] = (int32_t)nnz;
        if (z > 0) EMIT(row - N * N, -1.0 - cz);
        if (y > 0) EMIT(row - N, -1.0 - cy);
        if (x > 0) EMIT(row - 1, -1.0 - cx);
        EMIT(row, 6.0 + cx + cy + cz);
        if (x < N - 1) EMIT(row + 1, -1.0);
        if (y < N - 1) EMIT(row + N, -1.0);
        if (z < N - 1) EMIT(row + N * N, -1.0);
      }
  if (rp) rp[(int64_t)N * N * N] = (int32_t)nnz;
  return nnz;
}
#undef EMIT

static uint64_t splitmix64(uint64_t z) {
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

typedef struct { uint64_t key; uint64_t edge; } rmat_edge;
static int cmp_edge(const void* a, const void* b) {
  const rmat_edge *x = a, *y = b;
  if (x->key != y->key) return x->key < y->key ? -1 : 1;
  return x->edge < y->edge ? -1 : (x->edge > y->edge);
}

/* R-MAT with integer (16-bit) quadrant thresholds so that CPU and GPU generate identical edges:
 * per edge e and level group g, h = splitmix64(seed ^ (e*16+g)); four 16-bit draws per hash. */
int64_t oracle_gen_rmat(int32_t scale, int32_t edge_factor, uint64_t seed, int32_t* rp, int32_t* ci,
                        double* va) {
  const uint32_t ta = 37356, tb = ta + 12452, tc = tb + 12452;
  uint64_t E = (uint64_t)edge_factor << scale;
  int32_t n = 1 << scale;
  rmat_edge* ed = (rmat_edge*)malloc(sizeof(rmat_edge) * E);
  for (uint64_t e = 0; e < E; e++) {
    uint32_t row = 0, col = 0;
    uint64_t h = 0;
    for (int32_t l = 0; l < scale; l++) {
      if ((l & 3) == 0) h = splitmix64(seed ^ (e * 16 + (uint64_t)(l >> 2)));
      uint32_t u = (uint32_t)(h >> (16 * (l & 3))) & 0xffff;
      uint32_t rb = u >= tb, cb = (u >= ta && u < tb) || u >= tc;
      row = (row << 1) | rb;
      col = (col << 1) | cb;
    }
    ed[e].key = ((uint64_t)row << 32) | col;
    ed[e].edge = e;
  }
  qsort(ed, E, sizeof(rmat_edge), cmp_edge);
  int64_t nnz = 0;
  if (rp) memset(rp, 0, sizeof(int32_t) * ((size_t)n + 1));
  for (uint64_t i = 0; i < E;) {
    uint64_t j = i;
    double sum = 0.0;
    while (j < E && ed[j].key == ed[i].key) { /* duplicates summed in edge order */
      uint64_t hv = splitmix64((seed + 0x5851F42D4C957F2Dull) ^ ed[j].edge);
      sum += (double)(hv >> 11) * (2.0 / 9007199254740992.0) - 1.0;
      j++;
    }
    if (rp) {
      ci[nnz] = (int32_t)(ed[i].key & 0xffffffffu);
      va[nnz] = sum;
      rp[(ed[i].key >> 32) + 1]++;
    }
    nnz++;
    i = j;
  }
  if (rp) for (int32_t i = 0; i < n; i++) rp[i + 1] += rp[i];
  free(ed);
  return nnz;
}

/* ---- CPU baseline port ------------------------------------------------------------------ */
int oracle_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

int oracle_csr_spmv_omp(int64_t n, const int64_t* rp, const int32_t* ci, const double* va,
                        const double* x, double* y) {
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; i++) {
    double acc = 0.0;
    for (int64_t k = rp[i]; k < rp[i + 1]; k++) acc += va[k] * x[ci[k]];
    y[i] = acc;
  }
  return oracle_num_threads();
}

int oracle_csr_spmv_omp32(int32_t n, const int32_t* rp, const int32_t* ci, const double* va,
                          const double* x, double* y) {
#pragma omp parallel for schedule(static)
  for (int32_t i = 0; i < n; i++) {
    double acc = 0.0;
    for (int32_t k = rp[i]; k < rp[i + 1]; k++) acc += va[k] * x[ci[k]];
    y[i] = acc;
  }
  return oracle_num_threads();
}

/* ---- preconditioners ---------------------------------------------- */
/* position of (i, j) in row i, or -1 (DokMatrix::dok[i].count(j)) */
static int32_t find_entry(const int32_t* rp, const int32_t* ci, int32_t i, int32_t j) {
  for (int32_t k = rp[i]; k < rp[i + 1]; k++)
    if (ci[k] == j) return k;
  return -1;
}

/* ILUPreconditioner::ILUPreconditioner, src/runtime/SparseLinearSolvers.hpp:89-140: ILU(0) in IKJ order on
 * the pattern of `a` (rows in ascending column order, as CsrMatrix::toDok's std::map iterates them).
 * pc receives the factors in the pattern of `a`: strictly-lower entries = multipliers, the rest = U.
 * isNnz(k, j) is "entry exists AND its current value != 0" (SparseMatrix.hpp:209-215).
 * NB the reference's pcg hands the constructor whatever `a` it was given - in test/LinearSolvers.cpp:54-77
 * that is the LOWER TRIANGLE ONLY, for which no row update ever fires (row k has no column > k). */
int oracle_ilu0(int32_t n, const int32_t* rp, const int32_t* ci, const double* va, double* pc) {
  memcpy(pc, va, sizeof(double) * (size_t)rp[n]);
  for (int32_t i = 1; i < n; i++) {                               /* :93 */
    for (int32_t a = rp[i]; a < rp[i + 1]; a++) {                 /* :97  for (auto& p : pc.dok[i]) */
      const int32_t k = ci[a];
      if (k >= i) break;                                          /* :99-100 */
      const int32_t dk = find_entry(rp, ci, k, k);
      if (dk < 0 || pc[dk] == 0.0) continue;                      /* :101-102 !isNnz(k, k) */
      pc[a] = pc[a] / pc[dk];                                     /* :103 */
      const double beta = pc[a];                                  /* :104 */
      for (int32_t b = rp[i]; b < rp[i + 1]; b++) {               /* :106 */
        const int32_t j = ci[b];
        if (j < k + 1) continue;                                  /* :108-109 */
        const int32_t kj = find_entry(rp, ci, k, j);
        if (kj >= 0 && pc[kj] != 0.0) pc[b] = pc[b] - pc[kj] * beta; /* :111-113 */
      }
    }
  }
  return 0;
}

/* ILUPreconditioner::apply, SparseLinearSolvers.hpp:142-150: L y = x then U z = y with
 * mkl_dcsrtrsv(uplo, 'N', diag = 'N') (MklLayer.hpp:66-84) - NON-unit diagonal in BOTH solves: L is
 * getLowerTriangular() (j <= i, so it carries U's diagonal), U is getUpperTriangular() (i <= j).
 * MKL's internal summation order is unknowable; restated in ascending column order.
 * Returns 1 if a diagonal entry is missing or zero (MKL's behaviour is then undefined). */
int oracle_ilu_apply_mode(int32_t n, const int32_t* rp, const int32_t* ci, const double* pc, const double* x,
                          double* z, int unit_lower) {
  double* y = (double*)malloc(sizeof(double) * (size_t)(n > 0 ? n : 1));
  int rc = 0;
  for (int32_t i = 0; i < n; i++) {
    double acc = x[i], d = 0.0;
    for (int32_t k = rp[i]; k < rp[i + 1]; k++) {
      if (ci[k] < i) acc -= pc[k] * y[ci[k]];
      else if (ci[k] == i) d = pc[k];
    }
    if (unit_lower) d = 1.0;
    if (d == 0.0) rc = 1;
    y[i] = acc / d;
  }
  for (int32_t i = n - 1; i >= 0; i--) {
    double acc = y[i], d = 0.0;
    for (int32_t k = rp[i]; k < rp[i + 1]; k++) {
      if (ci[k] > i) acc -= pc[k] * z[ci[k]];
      else if (ci[k] == i) d = pc[k];
    }
    if (d == 0.0) rc = 1;
    z[i] = acc / d;
  }
  free(y);
  return rc;
}

int oracle_ilu_apply(int32_t n, const int32_t* rp, const int32_t* ci, const double* pc, const double* x,
                     double* z) {
  return oracle_ilu_apply_mode(n, rp, ci, pc, x, z, 0);
}

/* pcg<double, Precon>, SparseLinearSolvers.hpp:162-239, with the preconditioner left in:
 *   precon 0  IdentityPreconditioner (:64-73)
 *   precon 1  ILUPreconditioner (:77-156), built from the SAME arrays the product is taken from
 *   precon 2  Jacobi, z = r / a_ii (1 where a_ii is absent or 0) - NOT in the reference; SURVEY 8(f) rank 4
 *             proposes it as the first GPU preconditioner (it is what Eigen's BiCGSTAB defaults to)
 *   precon 3  the same ILU(0) factors applied with a UNIT lower solve, M = (I + L)(D + U) - the textbook ILU(0)
 *             preconditioner, symmetric for symmetric A.  NOT the reference: its non-unit lower solve makes M
 *             non-symmetric and its PCG stalls on every SPD stencil tried (tests/test_oracle_precond.py)
 * lower != 0: `a` is the stored lower triangle and the product is mkl_dcsrsymv('l') (the reference's call);
 * lower == 0: `a` is the full matrix (what the GPU path multiplies with). */
int oracle_pcg_precond(int32_t n, const int32_t* rp, const int32_t* ci, const double* va, const double* rhs,
                       double* x, int32_t* iterations, int32_t maxiters, double tol, int precon, int lower,
                       double* rs_final) {
  matvec_fn mv = lower ? symv_lower : gemv_full;
  const size_t nb = sizeof(double) * (size_t)(n > 0 ? n : 1);
  double *r = (double*)malloc(nb), *p = (double*)malloc(nb), *Ap = (double*)malloc(nb), *z = (double*)malloc(nb);
  double* pc = NULL;
  double* invd = NULL;
  int converged = 0, bad = 0;
  if (precon == 1 || precon == 3) {
    pc = (double*)malloc(sizeof(double) * (size_t)(rp[n] > 0 ? rp[n] : 1));
    oracle_ilu0(n, rp, ci, va, pc);
  } else if (precon == 2) {
    invd = (double*)malloc(nb);
    for (int32_t i = 0; i < n; i++) {
      const int32_t d = find_entry(rp, ci, i, i);
      invd[i] = (d >= 0 && va[d] != 0.0) ? 1.0 / va[d] : 1.0;
    }
  }
#define ORACLE_APPLY()                                                          \
  do {                                                                          \
    if (precon == 1 || precon == 3) bad |= oracle_ilu_apply_mode(n, rp, ci, pc, r, z, precon == 3); \
    else if (precon == 2) for (int32_t i_ = 0; i_ < n; i_++) z[i_] = invd[i_] * r[i_]; \
    else memcpy(z, r, nb);                                                      \
  } while (0)
  mv(n, rp, ci, va, x, r);                                        /* :189 */
  for (int32_t i = 0; i < n; i++) r[i] = rhs[i] - r[i];           /* :190 */
  ORACLE_APPLY();                                                 /* :193 */
  memcpy(p, z, nb);                                               /* :195 */
  double rsold = ddot(n, r, z), rsnew = rsold;                    /* :198 */
  for (int32_t it = 0; it < maxiters; it++) {                     /* :200 */
    mv(n, rp, ci, va, p, Ap);                                     /* :206 */
    const double alpha = rsold / ddot(n, p, Ap);                  /* :208 */
    for (int32_t i = 0; i < n; i++) x[i] += alpha * p[i];         /* :210 */
    for (int32_t i = 0; i < n; i++) r[i] = -alpha * Ap[i] + r[i]; /* :212 */
    ORACLE_APPLY();                                               /* :215 */
    rsnew = ddot(n, r, z);                                        /* :218 */
    if (rsnew <= tol * tol) { converged = 1; break; }             /* :220-226 */
    const double beta = rsnew / rsold;
    for (int32_t i = 0; i < n; i++) p[i] = z[i] + beta * p[i];    /* :229 daxpby(1, z, beta, p) */
    rsold = rsnew;                                                /* :230 */
    *iterations = it;                                             /* :231 */
  }
#undef ORACLE_APPLY
  if (rs_final) *rs_final = rsnew;
  free(r); free(p); free(Ap); free(z); free(pc); free(invd);
  return bad ? -1 : converged;
}
