/*
 * oracle/shim/mkl_shim.c -- open restatement of the oneMKL entry points LPM-C calls.
 *
 * TEST INFRASTRUCTURE ONLY (oracle).  See mkl.h for the call sites.  Nothing here is
 * derived from MKL sources; each routine restates the documented behaviour of the
 * corresponding MKL routine (Developer Reference: cblas_?nrm2, cblas_?gemv, cblas_?gemm,
 * LAPACKE_?gesv, mkl_sparse_?_mv, RCI ISS dcg_init/dcg_check/dcg/dcg_get).
 *
 * Parity note: summation orders are plain left-to-right loops (serial mode).  With
 * lpmb_shim_set_threads(t>1) the vector ops and the SpMV become OpenMP-parallel (used
 * only for the timed CPU baseline, never for parity fixtures).
 */
#include <math.h>
#include <omp.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "mkl.h"

static int g_threads = 1;
static long g_spmv_calls = 0;
static double g_spmv_seconds = 0.0;
static double g_setup_seconds = 0.0; /* threaded baseline: one-off work per matrix handle, part of g_spmv_seconds */
static int g_last_itercount = -1;
int lpmb_shim_last_itercount(void) { return g_last_itercount; }

void lpmb_shim_set_threads(int threads) { g_threads = threads < 1 ? 1 : threads; }
int lpmb_shim_get_threads(void) { return g_threads; }
long lpmb_shim_spmv_calls(void) { return g_spmv_calls; }
double lpmb_shim_spmv_seconds(void) { return g_spmv_seconds; }
double lpmb_shim_setup_seconds(void) { return g_setup_seconds; }
void lpmb_shim_reset_counters(void)
{
    g_spmv_calls = 0;
    g_spmv_seconds = 0.0;
    g_setup_seconds = 0.0;
}

/* ------------------------------------------------------------------ CBLAS */

/* Pairwise (cascade) summation, base blocks of 32 summed left to right: rounding error grows like
 * log2(n)*eps instead of n*eps.  MKL's own ddot is blocked/vectorised (not a naive loop) and its
 * exact order is unknowable here; a naive loop would put ~n*eps = 3e-12 of noise into every CG dot
 * product at n = 27 783, which CG amplifies to ~4e-9 in the bond forces after three load steps --
 * i.e. the oracle's own rounding noise would dominate the 1e-9 parity budget.  LPMB_SHIM_DOT=naive
 * restores the plain loop (used to measure exactly that sensitivity). */
static int g_dot_naive = -1;
static double dot_pairwise(const int n, const double *a, const double *b)
{
    if (n <= 32) {
        double s = 0.0;
        for (int i = 0; i < n; i++)
            s += a[i] * b[i];
        return s;
    }
    const int h = (n / 2 + 31) & ~31;
    return dot_pairwise(h, a, b) + dot_pairwise(n - h, a + h, b + h);
}

static double dot_n(const int n, const double *a, const double *b)
{
    double s = 0.0;
    if (g_dot_naive < 0) {
        const char *e = getenv("LPMB_SHIM_DOT");
        g_dot_naive = (e && strcmp(e, "naive") == 0) ? 1 : 0;
    }
    if (g_threads > 1) {
#pragma omp parallel for reduction(+ : s) num_threads(g_threads) schedule(static)
        for (int i = 0; i < n; i++)
            s += a[i] * b[i];
    } else if (g_dot_naive) {
        for (int i = 0; i < n; i++)
            s += a[i] * b[i];
    } else {
        s = dot_pairwise(n, a, b);
    }
    return s;
}

/* Euclidean norm, plain sum of squares (lpmc_project.c:412-413,463; boundary.c:88,176). */
double cblas_dnrm2(const MKL_INT n, const double *x, const MKL_INT incx)
{
    if (incx == 1 && n > 4096)
        return sqrt(dot_n(n, x, x));
    double s = 0.0;
    for (int i = 0; i < n; i++)
        s += x[i * incx] * x[i * incx];
    return sqrt(s);
}

/* y = alpha*op(A)*x + beta*y.  Only RowMajor/NoTrans is used (initialization.c:137...). */
void cblas_dgemv(const CBLAS_LAYOUT layout, const CBLAS_TRANSPOSE trans, const MKL_INT m, const MKL_INT n,
                 const double alpha, const double *a, const MKL_INT lda, const double *x, const MKL_INT incx,
                 const double beta, double *y, const MKL_INT incy)
{
    if (layout != CblasRowMajor || trans != CblasNoTrans) {
        fprintf(stderr, "mkl_shim: cblas_dgemv variant not supported\n");
        abort();
    }
    for (int i = 0; i < m; i++) {
        double s = 0.0;
        for (int j = 0; j < n; j++)
            s += a[i * lda + j] * x[j * incx];
        /* beta == 0 must not propagate NaN/garbage from y (BLAS convention). */
        y[i * incy] = (beta == 0.0) ? alpha * s : alpha * s + beta * y[i * incy];
    }
}

/* C = alpha*A*B + beta*C, RowMajor/NoTrans/NoTrans only (stiffness.c:23,56,185,216,248). */
void cblas_dgemm(const CBLAS_LAYOUT layout, const CBLAS_TRANSPOSE transa, const CBLAS_TRANSPOSE transb,
                 const MKL_INT m, const MKL_INT n, const MKL_INT k, const double alpha, const double *a,
                 const MKL_INT lda, const double *b, const MKL_INT ldb, const double beta, double *c,
                 const MKL_INT ldc)
{
    if (layout != CblasRowMajor || transa != CblasNoTrans || transb != CblasNoTrans) {
        fprintf(stderr, "mkl_shim: cblas_dgemm variant not supported\n");
        abort();
    }
    for (int i = 0; i < m; i++)
        for (int j = 0; j < n; j++) {
            double s = 0.0;
            for (int p = 0; p < k; p++)
                s += a[i * lda + p] * b[p * ldb + j];
            c[i * ldc + j] = (beta == 0.0) ? alpha * s : alpha * s + beta * c[i * ldc + j];
        }
}

/* ---------------------------------------------------------------- LAPACKE */

/* Row-major LU with partial pivoting (first row of maximal |a| wins, as LAPACK's idamax),
 * A overwritten by L\U, B by the solution; info = k+1 on an exactly zero pivot -- and then, as in
 * LAPACK's dgesv (dgetrf, then dgetrs only if info == 0), B is left exactly as it came in
 * (computeStrain, lpm_basic.c:209-242, reads it in that case). */
lapack_int LAPACKE_dgesv(int matrix_layout, lapack_int n, lapack_int nrhs, double *a, lapack_int lda,
                         lapack_int *ipiv, double *b, lapack_int ldb)
{
    if (matrix_layout != LAPACK_ROW_MAJOR) {
        fprintf(stderr, "mkl_shim: LAPACKE_dgesv column-major not supported\n");
        abort();
    }
    lapack_int info = 0;
    double b_in[64];
    const int keep = n * nrhs <= 64;
    if (keep)
        for (int i = 0; i < n; i++)
            for (int j = 0; j < nrhs; j++)
                b_in[i * nrhs + j] = b[i * ldb + j];
    for (int k = 0; k < n; k++) {
        int p = k;
        double amax = fabs(a[k * lda + k]);
        for (int i = k + 1; i < n; i++) {
            double v = fabs(a[i * lda + k]);
            if (v > amax) {
                amax = v;
                p = i;
            }
        }
        ipiv[k] = p + 1;
        if (a[p * lda + k] == 0.0) {
            if (info == 0)
                info = k + 1;
            continue;
        }
        if (p != k) {
            for (int j = 0; j < n; j++) {
                double t = a[k * lda + j];
                a[k * lda + j] = a[p * lda + j];
                a[p * lda + j] = t;
            }
            for (int j = 0; j < nrhs; j++) {
                double t = b[k * ldb + j];
                b[k * ldb + j] = b[p * ldb + j];
                b[p * ldb + j] = t;
            }
        }
        const double piv = a[k * lda + k];
        for (int i = k + 1; i < n; i++) {
            const double l = a[i * lda + k] / piv;
            a[i * lda + k] = l;
            if (l != 0.0) {
                for (int j = k + 1; j < n; j++)
                    a[i * lda + j] -= l * a[k * lda + j];
                for (int j = 0; j < nrhs; j++)
                    b[i * ldb + j] -= l * b[k * ldb + j];
            }
        }
    }
    if (info != 0) {
        if (keep)
            for (int i = 0; i < n; i++)
                for (int j = 0; j < nrhs; j++)
                    b[i * ldb + j] = b_in[i * nrhs + j];
        return info;
    }
    for (int j = 0; j < nrhs; j++)
        for (int i = n - 1; i >= 0; i--) {
            double s = b[i * ldb + j];
            for (int c = i + 1; c < n; c++)
                s -= a[i * lda + c] * b[c * ldb + j];
            b[i * ldb + j] = s / a[i * lda + i];
        }
    return 0;
}

/* ------------------------------------------------------------ Sparse BLAS */

struct lpmb_shim_sparse_matrix {
    int base, rows, cols;
    int *rs, *re, *col;
    double *val;
    /* threaded variant (timed CPU baseline only): row chunks of equal non-zero count, one per thread, and per-chunk
     * overflow buffers for the transposed contributions that land behind the chunk (at most `band` rows behind it) */
    int nchunk, band;
    int *chunk;   /* [nchunk+1] first row of every chunk */
    double *ovf;  /* [nchunk][band] */
};

sparse_status_t mkl_sparse_d_create_csr(sparse_matrix_t *A, const sparse_index_base_t indexing, const MKL_INT rows,
                                        const MKL_INT cols, MKL_INT *rows_start, MKL_INT *rows_end,
                                        MKL_INT *col_indx, double *values)
{
    struct lpmb_shim_sparse_matrix *m = (struct lpmb_shim_sparse_matrix *)calloc(1, sizeof(*m));
    m->base = (indexing == SPARSE_INDEX_BASE_ONE) ? 1 : 0;
    m->rows = rows;
    m->cols = cols;
    m->rs = rows_start;
    m->re = rows_end;
    m->col = col_indx;
    m->val = values;
    *A = m;
    return SPARSE_STATUS_SUCCESS;
}

sparse_status_t mkl_sparse_destroy(sparse_matrix_t A)
{
    if (A) {
        free(A->chunk);
        free(A->ovf);
        free(A);
    }
    return SPARSE_STATUS_SUCCESS;
}

/* Threaded symmetric SpMV works on the stored triangle directly (half the bytes of an expanded matrix, nothing to build
 * per handle but two small tables -- solver.c:206 creates a new handle for every solve).  Rows are cut into one chunk
 * per thread with equal non-zero counts.  A thread owns y on its chunk; the transposed products v*x_i that fall behind
 * its chunk (columns j >= chunk end; the lattice matrix is banded, j - i <= band) go into its private overflow buffer,
 * and after a barrier every thread adds the overflow parts that target its own rows, in chunk order. */
static void plan_threaded(struct lpmb_shim_sparse_matrix *m, int nt)
{
    const int n = m->rows, base = m->base;
    int band = 0;
#pragma omp parallel for reduction(max : band) num_threads(nt) schedule(static)
    for (int i = 0; i < n; i++)
        if (m->re[i] > m->rs[i]) {
            /* columns of a row are not assumed sorted */
            for (int k = m->rs[i] - base; k < m->re[i] - base; k++)
                if (m->col[k] - base - i > band)
                    band = m->col[k] - base - i;
        }
    m->band = band;
    m->nchunk = nt;
    m->chunk = (int *)malloc(((size_t)nt + 1) * sizeof(int));
    const long first = m->rs[0] - base, total = (long)(m->re[n - 1] - base) - first;
    m->chunk[0] = 0;
    int row = 0;
    for (int t = 1; t < nt; t++) {
        const long want = first + total * t / nt;
        while (row < n && m->rs[row] - base < want)
            row++;
        m->chunk[t] = row;
    }
    m->chunk[nt] = n;
    m->ovf = (double *)malloc((size_t)nt * (size_t)(band > 0 ? band : 1) * sizeof(double));
}

/* y = alpha*A*x + beta*y for SYMMETRIC/UPPER/NON_UNIT (solver.c:198-200,243). */
sparse_status_t mkl_sparse_d_mv(const sparse_operation_t operation, const double alpha, const sparse_matrix_t A,
                                const struct matrix_descr descr, const double *x, const double beta, double *y)
{
    if (operation != SPARSE_OPERATION_NON_TRANSPOSE || descr.type != SPARSE_MATRIX_TYPE_SYMMETRIC ||
        descr.mode != SPARSE_FILL_MODE_UPPER || descr.diag != SPARSE_DIAG_NON_UNIT) {
        fprintf(stderr, "mkl_shim: mkl_sparse_d_mv variant not supported\n");
        abort();
    }
    const double t0 = omp_get_wtime();
    struct lpmb_shim_sparse_matrix *m = A;
    const int n = m->rows, base = m->base;
    if (g_threads > 1 && n > 0) {
        if (!m->chunk || m->nchunk != g_threads) {
            free(m->chunk);
            free(m->ovf);
            plan_threaded(m, g_threads);
            g_setup_seconds += omp_get_wtime() - t0;
        }
        const int nt = m->nchunk, band = m->band;
#pragma omp parallel num_threads(nt)
        {
            /* one chunk per thread; loops over chunks so that a smaller team than requested still covers all of them */
            const int me = omp_get_thread_num(), team = omp_get_num_threads();
            for (int t = me; t < nt; t += team) {
                const int r0 = m->chunk[t], r1 = m->chunk[t + 1];
                double *ov = m->ovf + (size_t)t * (band > 0 ? band : 1);
                for (int j = 0; j < band; j++)
                    ov[j] = 0.0;
                for (int i = r0; i < r1; i++)
                    y[i] = (beta == 0.0) ? 0.0 : beta * y[i];
                for (int i = r0; i < r1; i++) {
                    double s = 0.0;
                    const double axi = alpha * x[i];
                    for (int k = m->rs[i] - base; k < m->re[i] - base; k++) {
                        const int j = m->col[k] - base;
                        const double v = m->val[k];
                        s += v * x[j];
                        if (j != i) {
                            if (j < r1)
                                y[j] += v * axi; /* j > i inside the own chunk (upper triangle) */
                            else
                                ov[j - r1] += v * axi;
                        }
                    }
                    y[i] += alpha * s;
                }
            }
#pragma omp barrier
            for (int t = me; t < nt; t += team) {
                const int r0 = m->chunk[t], r1 = m->chunk[t + 1];
                for (int s = 0; s < t; s++) { /* earlier chunks spill forward only */
                    const int e = m->chunk[s + 1];
                    const double *ov = m->ovf + (size_t)s * (band > 0 ? band : 1);
                    const int lo = r0 > e ? r0 : e, hi = r1 < e + band ? r1 : e + band;
                    for (int i = lo; i < hi; i++)
                        y[i] += ov[i - e];
                }
            }
        }
    } else {
        if (beta == 0.0)
            memset(y, 0, (size_t)n * sizeof(double));
        else
            for (int i = 0; i < n; i++)
                y[i] *= beta;
        for (int i = 0; i < n; i++) {
            double s = 0.0;
            const double xi = x[i];
            for (int k = m->rs[i] - base; k < m->re[i] - base; k++) {
                const int j = m->col[k] - base;
                const double v = m->val[k];
                s += v * x[j];
                if (j != i)
                    y[j] += alpha * (v * xi);
            }
            y[i] += alpha * s;
        }
    }
    g_spmv_calls++;
    g_spmv_seconds += omp_get_wtime() - t0;
    return SPARSE_STATUS_SUCCESS;
}

/* ---------------------------------------------------------------- RCI CG
 *
 * MKL RCI CG (documented scheme): tmp = [p | A*p | r | z], n each (solver.c:205,243).
 *   ipar[0]=n  ipar[3]=iteration counter  ipar[4]=max iterations
 *   ipar[7]=1 iteration-count stop test   ipar[8]=1 residual stop test
 *   ipar[9]=1 user stop test (rci_request=2)  ipar[10]=1 preconditioned (rci_request=3)
 *   dpar[0]=rel tol  dpar[1]=abs tol  dpar[2]=||r0||^2  dpar[3]=dpar[0]*dpar[2]+dpar[1]
 *   dpar[4]=||r_k||^2  dpar[5]=||r_{k-1}||^2  dpar[6]=alpha  dpar[7]=beta
 * Residual test (on SQUARED norms): dpar[4] <= dpar[3].
 * ipar[100] (private) holds the state-machine phase.
 */
#define PH ipar[100]

void dcg_init(const MKL_INT *n, const double *x, const double *b, MKL_INT *rci_request, MKL_INT *ipar, double *dpar,
              double *tmp)
{
    (void)x;
    (void)b;
    (void)tmp;
    memset(ipar, 0, 128 * sizeof(MKL_INT));
    memset(dpar, 0, 128 * sizeof(double));
    ipar[0] = *n;
    ipar[1] = 6;
    ipar[2] = 1;
    ipar[3] = 0;
    ipar[4] = (*n < 150) ? *n : 150;
    ipar[5] = 1;
    ipar[6] = 1;
    ipar[7] = 1;
    ipar[8] = 0;
    ipar[9] = 1;
    ipar[10] = 0;
    dpar[0] = 1.0e-6;
    dpar[1] = 0.0;
    PH = 0;
    *rci_request = 0;
}

void dcg_check(const MKL_INT *n, const double *x, const double *b, MKL_INT *rci_request, MKL_INT *ipar,
               double *dpar, double *tmp)
{
    (void)x;
    (void)b;
    (void)tmp;
    *rci_request = 0;
    if (ipar[0] != *n || ipar[4] < 0 || dpar[0] < 0.0 || dpar[1] < 0.0)
        *rci_request = -1100;
    if (ipar[10] != 0 || ipar[9] != 0) {
        /* the reference never enables these (solver.c:219-220); keep the shim honest */
        if (ipar[10] != 0)
            *rci_request = -1100;
    }
}

static void vec_axpy(const int n, const double a, const double *x, double *y)
{
    if (g_threads > 1) {
#pragma omp parallel for num_threads(g_threads) schedule(static)
        for (int i = 0; i < n; i++)
            y[i] += a * x[i];
    } else {
        for (int i = 0; i < n; i++)
            y[i] += a * x[i];
    }
}

void dcg(const MKL_INT *np, double *x, const double *b, MKL_INT *rci_request, MKL_INT *ipar, double *dpar,
         double *tmp)
{
    const int n = *np;
    double *p = tmp, *Ap = tmp + n, *r = tmp + 2 * (size_t)n;

    if (PH == 0) {
        /* first entry: ask for A*x0 */
        memcpy(p, x, (size_t)n * sizeof(double));
        PH = 1;
        *rci_request = 1;
        return;
    }
    if (PH == 1) {
        /* r0 = b - A*x0 ; p0 = r0 */
        for (int i = 0; i < n; i++)
            r[i] = b[i] - Ap[i];
        memcpy(p, r, (size_t)n * sizeof(double));
        dpar[2] = dot_n(n, r, r);
        dpar[3] = dpar[0] * dpar[2] + dpar[1];
        dpar[4] = dpar[2];
        ipar[3] = 0;
        if (ipar[8] && dpar[4] <= dpar[3]) { /* already converged (e.g. zero rhs) */
            PH = 3;
            *rci_request = 0;
            return;
        }
        PH = 2;
        *rci_request = 1;
        return;
    }
    if (PH == 2) {
        /* one CG iteration with Ap = A*p just delivered */
        const double pAp = dot_n(n, p, Ap);
        const double alpha = dpar[4] / pAp;
        dpar[6] = alpha;
        vec_axpy(n, alpha, p, x);
        vec_axpy(n, -alpha, Ap, r);
        dpar[5] = dpar[4];
        dpar[4] = dot_n(n, r, r);
        ipar[3] += 1;
        if (ipar[8] && dpar[4] <= dpar[3]) {
            PH = 3;
            *rci_request = 0;
            return;
        }
        if (ipar[7] && ipar[3] >= ipar[4]) {
            PH = 3;
            *rci_request = (ipar[8] || ipar[9]) ? -1 : 0;
            return;
        }
        const double beta = dpar[4] / dpar[5];
        dpar[7] = beta;
        if (g_threads > 1) {
#pragma omp parallel for num_threads(g_threads) schedule(static)
            for (int i = 0; i < n; i++)
                p[i] = r[i] + beta * p[i];
        } else {
            for (int i = 0; i < n; i++)
                p[i] = r[i] + beta * p[i];
        }
        *rci_request = 1;
        return;
    }
    *rci_request = 0;
}

void dcg_get(const MKL_INT *n, const double *x, const double *b, const MKL_INT *rci_request, const MKL_INT *ipar,
             const double *dpar, const double *tmp, MKL_INT *itercount)
{
    (void)n;
    (void)x;
    (void)b;
    (void)rci_request;
    (void)dpar;
    (void)tmp;
    *itercount = ipar[3];
    g_last_itercount = ipar[3];
}

/* ---------------------------------------------------------------- PARDISO */

void PARDISO(void *pt, const MKL_INT *maxfct, const MKL_INT *mnum, const MKL_INT *mtype, const MKL_INT *phase,
             const MKL_INT *n, const void *a, const MKL_INT *ia, const MKL_INT *ja, MKL_INT *perm,
             const MKL_INT *nrhs, MKL_INT *iparm, const MKL_INT *msglvl, void *b, void *x, MKL_INT *error)
{
    (void)pt; (void)maxfct; (void)mnum; (void)mtype; (void)phase; (void)n; (void)a; (void)ia; (void)ja;
    (void)perm; (void)nrhs; (void)iparm; (void)msglvl; (void)b; (void)x;
    fprintf(stderr, "mkl_shim: PARDISO is not part of the oracle (never selected by any driver)\n");
    *error = -1;
}

void mkl_free_buffers(void) {}
