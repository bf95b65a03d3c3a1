/*
 * oracle/shim/mkl.h -- open stand-in for the Intel oneMKL surface that LPM-C uses.
 *
 * TEST INFRASTRUCTURE ONLY.  Intel MKL is not vendored by the reference
 * (CMakeLists.txt:7 find_package(MKL), version unpinned, README.md:14 "2021.4+")
 * and is absent from this image, so the reference's unmodified sources are
 * compiled against this header + mkl_shim.c.  Only the 12 entry points the
 * reference calls are declared (call sites: solver.c:49-79,206-253,
 * stiffness.c:23,56,185,216,248, initialization.c:137.., constitutive.c:1213,1227,
 * lpm_basic.c:19,41,212, boundary.c:88,176, lpmc_project.c:412-413,463,558).
 *
 * "parity unpinned" at this boundary: there is no MKL binary and the reference
 * has no tests, so the summation order inside dcg / mkl_sparse_d_mv / dnrm2 is
 * defined, for this project, by mkl_shim.c's restatement of MKL's documented
 * algorithms.
 */
#ifndef LPMB_ORACLE_MKL_SHIM_H
#define LPMB_ORACLE_MKL_SHIM_H

#ifdef __cplusplus
extern "C" {
#endif

typedef int MKL_INT;

/* ---- CBLAS ---- */
typedef enum { CblasRowMajor = 101, CblasColMajor = 102 } CBLAS_LAYOUT;
typedef enum { CblasNoTrans = 111, CblasTrans = 112, CblasConjTrans = 113 } CBLAS_TRANSPOSE;

double cblas_dnrm2(const MKL_INT n, const double *x, const MKL_INT incx);
void cblas_dgemv(const CBLAS_LAYOUT layout, const CBLAS_TRANSPOSE trans, const MKL_INT m, const MKL_INT n,
                 const double alpha, const double *a, const MKL_INT lda, const double *x, const MKL_INT incx,
                 const double beta, double *y, const MKL_INT incy);
void cblas_dgemm(const CBLAS_LAYOUT layout, const CBLAS_TRANSPOSE transa, const CBLAS_TRANSPOSE transb,
                 const MKL_INT m, const MKL_INT n, const MKL_INT k, const double alpha, const double *a,
                 const MKL_INT lda, const double *b, const MKL_INT ldb, const double beta, double *c,
                 const MKL_INT ldc);

/* ---- LAPACKE ---- */
#define LAPACK_ROW_MAJOR 101
#define LAPACK_COL_MAJOR 102
typedef int lapack_int;
lapack_int LAPACKE_dgesv(int matrix_layout, lapack_int n, lapack_int nrhs, double *a, lapack_int lda,
                         lapack_int *ipiv, double *b, lapack_int ldb);

/* ---- Sparse BLAS (inspector-executor) ---- */
typedef enum { SPARSE_STATUS_SUCCESS = 0, SPARSE_STATUS_NOT_SUPPORTED = 6 } sparse_status_t;
typedef enum { SPARSE_INDEX_BASE_ZERO = 0, SPARSE_INDEX_BASE_ONE = 1 } sparse_index_base_t;
typedef enum {
    SPARSE_OPERATION_NON_TRANSPOSE = 10,
    SPARSE_OPERATION_TRANSPOSE = 11,
    SPARSE_OPERATION_CONJUGATE_TRANSPOSE = 12
} sparse_operation_t;
typedef enum {
    SPARSE_MATRIX_TYPE_GENERAL = 20,
    SPARSE_MATRIX_TYPE_SYMMETRIC = 21,
    SPARSE_MATRIX_TYPE_HERMITIAN = 22,
    SPARSE_MATRIX_TYPE_TRIANGULAR = 23,
    SPARSE_MATRIX_TYPE_DIAGONAL = 24
} sparse_matrix_type_t;
typedef enum { SPARSE_FILL_MODE_LOWER = 40, SPARSE_FILL_MODE_UPPER = 41, SPARSE_FILL_MODE_FULL = 42 } sparse_fill_mode_t;
typedef enum { SPARSE_DIAG_NON_UNIT = 50, SPARSE_DIAG_UNIT = 51 } sparse_diag_type_t;

struct matrix_descr {
    sparse_matrix_type_t type;
    sparse_fill_mode_t mode;
    sparse_diag_type_t diag;
};

struct lpmb_shim_sparse_matrix;
typedef struct lpmb_shim_sparse_matrix *sparse_matrix_t;

sparse_status_t mkl_sparse_d_create_csr(sparse_matrix_t *A, const sparse_index_base_t indexing, const MKL_INT rows,
                                        const MKL_INT cols, MKL_INT *rows_start, MKL_INT *rows_end,
                                        MKL_INT *col_indx, double *values);
sparse_status_t mkl_sparse_d_mv(const sparse_operation_t operation, const double alpha, const sparse_matrix_t A,
                                const struct matrix_descr descr, const double *x, const double beta, double *y);
sparse_status_t mkl_sparse_destroy(sparse_matrix_t A);

/* ---- RCI conjugate gradient ---- */
void dcg_init(const MKL_INT *n, const double *x, const double *b, MKL_INT *rci_request, MKL_INT *ipar, double *dpar,
              double *tmp);
void dcg_check(const MKL_INT *n, const double *x, const double *b, MKL_INT *rci_request, MKL_INT *ipar,
               double *dpar, double *tmp);
void dcg(const MKL_INT *n, double *x, const double *b, MKL_INT *rci_request, MKL_INT *ipar, double *dpar,
         double *tmp);
void dcg_get(const MKL_INT *n, const double *x, const double *b, const MKL_INT *rci_request, const MKL_INT *ipar,
             const double *dpar, const double *tmp, MKL_INT *itercount);

/* ---- PARDISO (symbol only; never selected: cal_method="cg", lpmc_project.c:282) ---- */
void PARDISO(void *pt, const MKL_INT *maxfct, const MKL_INT *mnum, const MKL_INT *mtype, const MKL_INT *phase,
             const MKL_INT *n, const void *a, const MKL_INT *ia, const MKL_INT *ja, MKL_INT *perm,
             const MKL_INT *nrhs, MKL_INT *iparm, const MKL_INT *msglvl, void *b, void *x, MKL_INT *error);

/* ---- service ---- */
void mkl_free_buffers(void);

/* ---- shim controls (not MKL): used by the timed CPU baseline ---- */
/* threads > 1 switches dnrm2 / dcg vector ops / sparse mv to OpenMP-parallel variants
 * (sparse mv then runs row-parallel on a full CSR expanded once per handle). */
void lpmb_shim_set_threads(int threads);
int lpmb_shim_get_threads(void);
/* counters for the harness */
long lpmb_shim_spmv_calls(void);
double lpmb_shim_spmv_seconds(void);
void lpmb_shim_reset_counters(void);
/* iteration count reported by the most recent dcg_get() (solver.c:253) */
int lpmb_shim_last_itercount(void);

#ifdef __cplusplus
}
#endif
#endif
