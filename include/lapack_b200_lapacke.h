/*
 * lapack_b200_lapacke.h -- LAPACKE C entry points exported by liblapack_b200.so for the hot path.
 *
 * Same names, argument lists and return conventions as the reference's LAPACKE/include/lapacke.h (line
 * numbers cited per group); lapack_int is 32-bit (lapack.h:97-102, LP64 build).  matrix_layout is
 * LAPACK_ROW_MAJOR (101) or LAPACK_COL_MAJOR (102) (lapacke.h:65-66).
 */
#ifndef LAPACK_B200_LAPACKE_H
#define LAPACK_B200_LAPACKE_H
#ifdef __cplusplus
extern "C" {
#endif

#ifndef lapack_int
#define lapack_int int
#endif
#ifndef LAPACK_ROW_MAJOR
#define LAPACK_ROW_MAJOR 101
#define LAPACK_COL_MAJOR 102
#define LAPACK_WORK_MEMORY_ERROR -1010
#define LAPACK_TRANSPOSE_MEMORY_ERROR -1011
#endif

void LAPACKE_xerbla(const char* name, lapack_int info);     /* LAPACKE/utils/lapacke_xerbla.c:36 */
void LAPACKE_set_nancheck(int flag);                         /* LAPACKE/src/lapacke_nancheck.c */
int LAPACKE_get_nancheck(void);

/* lapacke.h:1131, 1142, 6193 */
lapack_int LAPACKE_dgetrf(int matrix_layout, lapack_int m, lapack_int n, double* a, lapack_int lda, lapack_int* ipiv);
lapack_int LAPACKE_dgetrf_work(int matrix_layout, lapack_int m, lapack_int n, double* a, lapack_int lda, lapack_int* ipiv);
lapack_int LAPACKE_dgetrf2(int matrix_layout, lapack_int m, lapack_int n, double* a, lapack_int lda, lapack_int* ipiv);
lapack_int LAPACKE_dgetrf2_work(int matrix_layout, lapack_int m, lapack_int n, double* a, lapack_int lda, lapack_int* ipiv);
/* lapacke.h:1165, 6232 */
lapack_int LAPACKE_dgetrs(int matrix_layout, char trans, lapack_int n, lapack_int nrhs, const double* a, lapack_int lda,
                          const lapack_int* ipiv, double* b, lapack_int ldb);
lapack_int LAPACKE_dgetrs_work(int matrix_layout, char trans, lapack_int n, lapack_int nrhs, const double* a, lapack_int lda,
                               const lapack_int* ipiv, double* b, lapack_int ldb);
/* lapacke.h:939, 5957 */
lapack_int LAPACKE_dgesv(int matrix_layout, lapack_int n, lapack_int nrhs, double* a, lapack_int lda, lapack_int* ipiv,
                         double* b, lapack_int ldb);
lapack_int LAPACKE_dgesv_work(int matrix_layout, lapack_int n, lapack_int nrhs, double* a, lapack_int lda, lapack_int* ipiv,
                              double* b, lapack_int ldb);
/* lapacke.h:3124, 3115, 8701 */
lapack_int LAPACKE_dpotrf(int matrix_layout, char uplo, lapack_int n, double* a, lapack_int lda);
lapack_int LAPACKE_dpotrf_work(int matrix_layout, char uplo, lapack_int n, double* a, lapack_int lda);
lapack_int LAPACKE_dpotrf2(int matrix_layout, char uplo, lapack_int n, double* a, lapack_int lda);
lapack_int LAPACKE_dpotrf2_work(int matrix_layout, char uplo, lapack_int n, double* a, lapack_int lda);
/* lapacke.h:3143, 8720 */
lapack_int LAPACKE_dpotrs(int matrix_layout, char uplo, lapack_int n, lapack_int nrhs, const double* a, lapack_int lda,
                          double* b, lapack_int ldb);
lapack_int LAPACKE_dpotrs_work(int matrix_layout, char uplo, lapack_int n, lapack_int nrhs, const double* a, lapack_int lda,
                               double* b, lapack_int ldb);
/* lapacke.h:3030, 8589 */
lapack_int LAPACKE_dposv(int matrix_layout, char uplo, lapack_int n, lapack_int nrhs, double* a, lapack_int lda, double* b,
                         lapack_int ldb);
lapack_int LAPACKE_dposv_work(int matrix_layout, char uplo, lapack_int n, lapack_int nrhs, double* a, lapack_int lda,
                              double* b, lapack_int ldb);
/* lapacke.h:818, 5690, 807, 5675 */
lapack_int LAPACKE_dgeqrf(int matrix_layout, lapack_int m, lapack_int n, double* a, lapack_int lda, double* tau);
lapack_int LAPACKE_dgeqrf_work(int matrix_layout, lapack_int m, lapack_int n, double* a, lapack_int lda, double* tau,
                               double* work, lapack_int lwork);
lapack_int LAPACKE_dgeqr2(int matrix_layout, lapack_int m, lapack_int n, double* a, lapack_int lda, double* tau);
lapack_int LAPACKE_dgeqr2_work(int matrix_layout, lapack_int m, lapack_int n, double* a, lapack_int lda, double* tau,
                               double* work);
/* LAPACKE_dgelqf / LAPACKE_dormlq (+ _work), LAPACKE/src/lapacke_dgelqf*.c, lapacke_dormlq*.c */
lapack_int LAPACKE_dgelqf(int matrix_layout, lapack_int m, lapack_int n, double* a, lapack_int lda, double* tau);
lapack_int LAPACKE_dgelqf_work(int matrix_layout, lapack_int m, lapack_int n, double* a, lapack_int lda, double* tau,
                               double* work, lapack_int lwork);
lapack_int LAPACKE_dormlq(int matrix_layout, char side, char trans, lapack_int m, lapack_int n, lapack_int k,
                          const double* a, lapack_int lda, const double* tau, double* c, lapack_int ldc);
lapack_int LAPACKE_dormlq_work(int matrix_layout, char side, char trans, lapack_int m, lapack_int n, lapack_int k,
                               const double* a, lapack_int lda, const double* tau, double* c, lapack_int ldc, double* work,
                               lapack_int lwork);
/* LAPACKE_dgels / LAPACKE_dgels_work (LAPACKE/src/lapacke_dgels.c, lapacke_dgels_work.c) -- SURVEY 8f rank 4 */
lapack_int LAPACKE_dgels(int matrix_layout, char trans, lapack_int m, lapack_int n, lapack_int nrhs, double* a, lapack_int lda,
                         double* b, lapack_int ldb);
lapack_int LAPACKE_dgels_work(int matrix_layout, char trans, lapack_int m, lapack_int n, lapack_int nrhs, double* a,
                              lapack_int lda, double* b, lapack_int ldb, double* work, lapack_int lwork);
/* lapacke.h:11369, 11526 (DGEQRT), 11348, 11505 (DGEMQRT) -- SURVEY 8f rank 4; row-major T is nb x k, ldt >= k */
lapack_int LAPACKE_dgeqrt(int matrix_layout, lapack_int m, lapack_int n, lapack_int nb, double* a, lapack_int lda, double* t,
                          lapack_int ldt);
lapack_int LAPACKE_dgeqrt_work(int matrix_layout, lapack_int m, lapack_int n, lapack_int nb, double* a, lapack_int lda,
                               double* t, lapack_int ldt, double* work);
lapack_int LAPACKE_dgemqrt(int matrix_layout, char side, char trans, lapack_int m, lapack_int n, lapack_int k, lapack_int nb,
                           const double* v, lapack_int ldv, const double* t, lapack_int ldt, double* c, lapack_int ldc);
lapack_int LAPACKE_dgemqrt_work(int matrix_layout, char side, char trans, lapack_int m, lapack_int n, lapack_int k,
                                lapack_int nb, const double* v, lapack_int ldv, const double* t, lapack_int ldt, double* c,
                                lapack_int ldc, double* work);
/* lapacke.h:1153, 6216 (DGETRI) -- SURVEY 8f rank 2 */
lapack_int LAPACKE_dgetri(int matrix_layout, lapack_int n, double* a, lapack_int lda, const lapack_int* ipiv);
lapack_int LAPACKE_dgetri_work(int matrix_layout, lapack_int n, double* a, lapack_int lda, const lapack_int* ipiv,
                               double* work, lapack_int lwork);
/* lapacke.h:2664, 8163 (DORGQR), 2729, 8246 (DORMQR) -- SURVEY 8f rank 1 */
lapack_int LAPACKE_dorgqr(int matrix_layout, lapack_int m, lapack_int n, lapack_int k, double* a, lapack_int lda,
                          const double* tau);
lapack_int LAPACKE_dorgqr_work(int matrix_layout, lapack_int m, lapack_int n, lapack_int k, double* a, lapack_int lda,
                               const double* tau, double* work, lapack_int lwork);
lapack_int LAPACKE_dormqr(int matrix_layout, char side, char trans, lapack_int m, lapack_int n, lapack_int k,
                          const double* a, lapack_int lda, const double* tau, double* c, lapack_int ldc);
lapack_int LAPACKE_dormqr_work(int matrix_layout, char side, char trans, lapack_int m, lapack_int n, lapack_int k,
                               const double* a, lapack_int lda, const double* tau, double* c, lapack_int ldc, double* work,
                               lapack_int lwork);
/* lapacke.h:2495, 7976 */
lapack_int LAPACKE_dlarft(int matrix_layout, char direct, char storev, lapack_int n, lapack_int k, const double* v,
                          lapack_int ldv, const double* tau, double* t, lapack_int ldt);
lapack_int LAPACKE_dlarft_work(int matrix_layout, char direct, char storev, lapack_int n, lapack_int k, const double* v,
                               lapack_int ldv, const double* tau, double* t, lapack_int ldt);
/* lapacke.h:2462, 7939 */
lapack_int LAPACKE_dlarfb(int matrix_layout, char side, char trans, char direct, char storev, lapack_int m, lapack_int n,
                          lapack_int k, const double* v, lapack_int ldv, const double* t, lapack_int ldt, double* c,
                          lapack_int ldc);
lapack_int LAPACKE_dlarfb_work(int matrix_layout, char side, char trans, char direct, char storev, lapack_int m,
                               lapack_int n, lapack_int k, const double* v, lapack_int ldv, const double* t, lapack_int ldt,
                               double* c, lapack_int ldc, double* work, lapack_int ldwork);
/* lapacke.h:2577, 8061 */
lapack_int LAPACKE_dlaswp(int matrix_layout, lapack_int n, double* a, lapack_int lda, lapack_int k1, lapack_int k2,
                          const lapack_int* ipiv, lapack_int incx);
lapack_int LAPACKE_dlaswp_work(int matrix_layout, lapack_int n, double* a, lapack_int lda, lapack_int k1, lapack_int k2,
                               const lapack_int* ipiv, lapack_int incx);

#ifdef __cplusplus
}
#endif
#endif
