/*
 * oracle.h -- CPU restatement of the Reference-LAPACK hot path (TEST INFRASTRUCTURE ONLY).
 *
 * This directory is the parity oracle.  Nothing under lapack_b200/ may include, link or call it;
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs do.
 *
 * Every function restates one routine of /root/reference (file:line cited at the definition) in
 * plain C, single thread, same loop order and same floating-point operation order as the Fortran
 * (compile with -O2 -ffp-contract=off, no -ffast-math, so no FMA contraction is introduced).
 * Conventions: all matrices column-major, 0-based C indexing inside, but every *value* that the
 * reference defines as 1-based (IPIV entries, INFO, IDAMAX result) stays 1-based.
 * Scalars are passed by value (this is not the Fortran ABI; the ABI lives in include/).
 *
 * Pinning: the reference ships no golden vectors for this path.  The oracle is
 * pinned against the gfortran-compiled netlib LAPACK 3.12.0 routines inside scipy's bundled
 * OpenBLAS (tests/golden/make_golden.py -> tests/golden/ npz files) and the DLARNV known answer.
 */
#ifndef LAPACK_B200_ORACLE_H
#define LAPACK_B200_ORACLE_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif

/* ---- BLAS/SRC ---- */
int    ora_lsame(char a, char b);
int    ora_idamax(int n, const double *x, int incx);
void   ora_dscal(int n, double a, double *x, int incx);
void   ora_dswap(int n, double *x, int incx, double *y, int incy);
void   ora_daxpy(int n, double a, const double *x, int incx, double *y, int incy);
void   ora_dcopy(int n, const double *x, int incx, double *y, int incy);
double ora_ddot(int n, const double *x, int incx, const double *y, int incy);
double ora_dnrm2(int n, const double *x, int incx);
void   ora_dgemv(char trans, int m, int n, double alpha, const double *a, int lda,
                 const double *x, int incx, double beta, double *y, int incy);
void   ora_dger(int m, int n, double alpha, const double *x, int incx, const double *y, int incy,
                double *a, int lda);
void   ora_dtrmv(char uplo, char trans, char diag, int n, const double *a, int lda, double *x, int incx);
void   ora_dgemm(char transa, char transb, int m, int n, int k, double alpha, const double *a, int lda,
                 const double *b, int ldb, double beta, double *c, int ldc);
void   ora_dtrsm(char side, char uplo, char transa, char diag, int m, int n, double alpha,
                 const double *a, int lda, double *b, int ldb);
void   ora_dtrmm(char side, char uplo, char transa, char diag, int m, int n, double alpha,
                 const double *a, int lda, double *b, int ldb);
void   ora_dsyrk(char uplo, char trans, int n, int k, double alpha, const double *a, int lda,
                 double beta, double *c, int ldc);

/* ---- env / aux ---- */
double ora_dlamch(char cmach);
int    ora_ilaenv_nb(const char *name);      /* ISPEC=1 block size for DGETRF/DPOTRF/DGEQRF */
void   ora_set_nb(int nb_getrf, int nb_potrf, int nb_geqrf, int nx_geqrf); /* like TESTING/LIN/xlaenv.f */
double ora_dlapy2(double x, double y);
void   ora_dlaruv(int iseed[4], int n, double *x);
void   ora_dlarnv(int idist, int iseed[4], long n, double *x);
double ora_dlange(char norm, int m, int n, const double *a, int lda);
double ora_dlansy(char norm, char uplo, int n, const double *a, int lda);
void   ora_dlacpy(char uplo, int m, int n, const double *a, int lda, double *b, int ldb);
void   ora_dlaset(char uplo, int m, int n, double alpha, double beta, double *a, int lda);

/* ---- LU ---- */
void ora_dlaswp(int n, double *a, int lda, int k1, int k2, const int *ipiv, int incx);
void ora_dgetrf2(int m, int n, double *a, int lda, int *ipiv, int *info);
void ora_dgetrf(int m, int n, double *a, int lda, int *ipiv, int *info);
void ora_dgetrs(char trans, int n, int nrhs, const double *a, int lda, const int *ipiv,
                double *b, int ldb, int *info);
void ora_dgesv(int n, int nrhs, double *a, int lda, int *ipiv, double *b, int ldb, int *info);

/* ---- Cholesky ---- */
void ora_dpotrf2(char uplo, int n, double *a, int lda, int *info);
void ora_dpotrf(char uplo, int n, double *a, int lda, int *info);
void ora_dpotrs(char uplo, int n, int nrhs, const double *a, int lda, double *b, int ldb, int *info);
void ora_dposv(char uplo, int n, int nrhs, double *a, int lda, double *b, int ldb, int *info);

/* ---- QR ---- */
void ora_dlarfg(int n, double *alpha, double *x, int incx, double *tau);
int  ora_iladlc(int m, int n, const double *a, int lda);
int  ora_iladlr(int m, int n, const double *a, int lda);
void ora_dlarf1f(char side, int m, int n, const double *v, int incv, double tau, double *c, int ldc,
                 double *work);
void ora_dgeqr2(int m, int n, double *a, int lda, double *tau, double *work, int *info);
void ora_dlarft_lvl2(char direct, char storev, int n, int k, const double *v, int ldv,
                     const double *tau, double *t, int ldt);
void ora_dlarft(char direct, char storev, int n, int k, const double *v, int ldv,
                const double *tau, double *t, int ldt);
void ora_dlarfb(char side, char trans, char direct, char storev, int m, int n, int k,
                const double *v, int ldv, const double *t, int ldt, double *c, int ldc,
                double *work, int ldwork);
void ora_dgeqrf(int m, int n, double *a, int lda, double *tau, double *work, int lwork, int *info);
void ora_dorg2r(int m, int n, int k, double *a, int lda, const double *tau, double *work, int *info);
void ora_dtrti2(char uplo, char diag, int n, double *a, int lda, int *info);
void ora_dtrtri(char uplo, char diag, int n, double *a, int lda, int *info);
void ora_dgetri(int n, double *a, int lda, const int *ipiv, double *work, int lwork, int *info);
void ora_set_nb_getri(int nb);
void ora_dgeqrt3(int m, int n, double *a, int lda, double *t, int ldt, int *info);
void ora_dgeqrt(int m, int n, int nb, double *a, int lda, double *t, int ldt, double *work, int *info);
void ora_dgemqrt(char side, char trans, int m, int n, int k, int nb, const double *v, int ldv, const double *t, int ldt,
                 double *c, int ldc, double *work, int *info);
void ora_dlascl_g(double cfrom, double cto, int m, int n, double *a, int lda);
void ora_dtrtrs(char uplo, char trans, char diag, int n, int nrhs, const double *a, int lda, double *b, int ldb, int *info);
void ora_dgelq2(int m, int n, double *a, int lda, double *tau, double *work, int *info);
void ora_dorml2(char side, char trans, int m, int n, int k, const double *a, int lda, const double *tau, double *c, int ldc,
                double *work, int *info);
void ora_dgels(char trans, int m, int n, int nrhs, double *a, int lda, double *b, int ldb, double *work, int lwork, int *info);
void ora_dlacn2(int n, double *v, double *x, int *isgn, double *est, int *kase, int *isave);
/* ---- tall-skinny QR, L = 0 forms of the triangular-pentagonal kernels ---- */
void ora_dtpqrt2_l0(int m, int n, double *a, int lda, double *b, int ldb, double *t, int ldt);
void ora_dtprfb_ltfc_l0(int m, int n, int k, const double *v, int ldv, const double *t, int ldt, double *a, int lda, double *b,
                        int ldb, double *work, int ldwork);
void ora_dtpqrt_l0(int m, int n, int nb, double *a, int lda, double *b, int ldb, double *t, int ldt, double *work, int *info);
void ora_dlatsqr(int m, int n, int mb, int nb, double *a, int lda, double *t, int ldt, double *work, int *info);
/* ---- condition estimation / expert driver ---- */
void ora_dtrsv(char uplo, char trans, char diag, int n, const double *a, int lda, double *x);
void ora_drscl(int n, double sa, double *x);
void ora_dlatrs(char uplo, char trans, char diag, char normin, int n, const double *a, int lda, double *x, double *scale,
                double *cnorm, int *info);
void ora_dgecon(char norm, int n, const double *a, int lda, double anorm, double *rcond, double *work, int *iwork, int *info);
void ora_dgeequ(int m, int n, const double *a, int lda, double *r, double *c, double *rowcnd, double *colcnd, double *amax,
                int *info);
char ora_dlaqge(int m, int n, double *a, int lda, const double *r, const double *c, double rowcnd, double colcnd, double amax);
void ora_dgesvx(char fact, char trans, int n, int nrhs, double *a, int lda, double *af, int ldaf, int *ipiv, char *equed,
                double *r, double *c, double *b, int ldb, double *x, int ldx, double *rcond, double *ferr, double *berr,
                double *work, int *iwork, int *info);
void ora_dgerfs(char trans, int n, int nrhs, const double *a, int lda, const double *af, int ldaf, const int *ipiv,
                const double *b, int ldb, double *x, int ldx, double *ferr, double *berr, double *work, int *iwork, int *info);
void ora_dorm2r(char side, char trans, int m, int n, int k, const double *a, int lda, const double *tau, double *c, int ldc,
                double *work, int *info);
void ora_dormqr(char side, char trans, int m, int n, int k, const double *a, int lda, const double *tau, double *c, int ldc,
                double *work, int lwork, int *info);
void ora_dorgqr(int m, int n, int k, double *a, int lda, const double *tau, double *work, int lwork,
                int *info);

/* ---- TESTING/LIN checkers (residual ratios, pass iff < 30) ---- */
void ora_dget01(int m, int n, const double *a, int lda, double *afac, int ldafac, const int *ipiv,
                double *rwork, double *resid);
void ora_dget02(char trans, int m, int n, int nrhs, const double *a, int lda, const double *x, int ldx,
                double *b, int ldb, double *rwork, double *resid);
void ora_dget04(int n, int nrhs, const double *x, int ldx, const double *xact, int ldxact,
                double rcond, double *resid);
void ora_dpot01(char uplo, int n, const double *a, int lda, double *afac, int ldafac, double *rwork,
                double *resid);
void ora_dpot02(char uplo, int n, int nrhs, const double *a, int lda, const double *x, int ldx,
                double *b, int ldb, double *rwork, double *resid);
void ora_dqrt01(int m, int n, const double *a, const double *af, double *q, double *r, int lda,
                const double *tau, double *work, int lwork, double *rwork, double *result);

#ifdef __cplusplus
}
#endif
#endif
