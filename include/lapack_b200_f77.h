/*
 * lapack_b200_f77.h -- Fortran-77 ABI entry points exported by liblapack_b200.so.
 *
 * These are the link-time drop-in symbols for the reference's hot path: same names
 * (lower case + trailing underscore = LAPACK_GLOBAL, LAPACKE/include/lapacke_mangling_with_flags.h.in:4-13),
 * same argument lists as LAPACKE/include/lapack.h (cited per prototype), every argument by reference,
 * hidden CHARACTER lengths appended as size_t (LAPACK_FORTRAN_STRLEN_END, lapack.h:20-24).
 * Matrix / vector pointers may be host or device memory.
 */
#ifndef LAPACK_B200_F77_H
#define LAPACK_B200_F77_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif

/* BLAS/SRC/xerbla.f:59 (weak: an application-supplied XERBLA takes precedence) and BLAS/SRC/lsame.f */
void xerbla_(const char* srname, const int* info, size_t srname_len);
int lsame_(const char* ca, const char* cb, size_t, size_t);
/* behaviour of the built-in xerbla_: 0 = print + STOP like the reference, 1 = print + return, 2 = record only */
void lb200_set_xerbla_mode(int mode);
int lb200_last_xerbla(char* name_out /* >= 33 bytes */, int* info_out);   /* returns number of calls so far */
void lb200_clear_xerbla(void);

/* BLAS/SRC/dgemm.f:187, dsyrk.f:168, dtrsm.f:180, dtrmm.f:176   (CBLAS/include/cblas_f77.h:440-448) */
void dgemm_(const char* transa, const char* transb, const int* m, const int* n, const int* k, const double* alpha,
            const double* A, const int* lda, const double* B, const int* ldb, const double* beta, double* C,
            const int* ldc, size_t, size_t);
void dsyrk_(const char* uplo, const char* trans, const int* n, const int* k, const double* alpha, const double* A,
            const int* lda, const double* beta, double* C, const int* ldc, size_t, size_t);
void dtrsm_(const char* side, const char* uplo, const char* transa, const char* diag, const int* m, const int* n,
            const double* alpha, const double* A, const int* lda, double* B, const int* ldb, size_t, size_t, size_t, size_t);
void dtrmm_(const char* side, const char* uplo, const char* transa, const char* diag, const int* m, const int* n,
            const double* alpha, const double* A, const int* lda, double* B, const int* ldb, size_t, size_t, size_t, size_t);

/* LAPACKE/include/lapack.h:4349 (dgetrf), :4373 (dgetrf2), :11326 (dlaswp), :4436 (dgetrs), :3706 (dgesv) */
void dgetrf_(const int* m, const int* n, double* A, const int* lda, int* ipiv, int* info);
void dgetrf2_(const int* m, const int* n, double* A, const int* lda, int* ipiv, int* info);
void dlaswp_(const int* n, double* A, const int* lda, const int* k1, const int* k2, const int* ipiv, const int* incx);
void dgetrs_(const char* trans, const int* n, const int* nrhs, const double* A, const int* lda, const int* ipiv, double* B,
             const int* ldb, int* info, size_t);
void dgesv_(const int* n, const int* nrhs, double* A, const int* lda, int* ipiv, double* B, const int* ldb, int* info);

/* lapack.h:13786 (dpotrf), dpotrf2, :13979 (dpotrs), :13394 (dposv) */
void dpotrf_(const char* uplo, const int* n, double* A, const int* lda, int* info, size_t);
void dpotrf2_(const char* uplo, const int* n, double* A, const int* lda, int* info, size_t);
void dpotrs_(const char* uplo, const int* n, const int* nrhs, const double* A, const int* lda, double* B, const int* ldb,
             int* info, size_t);
void dposv_(const char* uplo, const int* n, const int* nrhs, double* A, const int* lda, double* B, const int* ldb,
            int* info, size_t);

/* lapack.h:2991 (dgeqrf), dgeqr2, :10946 (dlarft), :10847 (dlarfb) */
void dgeqrf_(const int* m, const int* n, double* A, const int* lda, double* tau, double* work, const int* lwork, int* info);
void dgeqr2_(const int* m, const int* n, double* A, const int* lda, double* tau, double* work, int* info);
/* SRC/dgelqf.f:142 DGELQF, SRC/dormlq.f:165 DORMLQ, SRC/dgels.f:189 DGELS(TRANS,M,N,NRHS,A,LDA,B,LDB,WORK,LWORK,INFO) */
void dgelqf_(const int* m, const int* n, double* A, const int* lda, double* tau, double* work, const int* lwork, int* info);
void dormlq_(const char* side, const char* trans, const int* m, const int* n, const int* k, const double* A, const int* lda,
             const double* tau, double* C, const int* ldc, double* work, const int* lwork, int* info, size_t, size_t);
void dgels_(const char* trans, const int* m, const int* n, const int* nrhs, double* A, const int* lda, double* B, const int* ldb,
            double* work, const int* lwork, int* info, size_t);
/* SRC/dgeqrt.f:139 DGEQRT(M,N,NB,A,LDA,T,LDT,WORK,INFO); SRC/dgemqrt.f:166 DGEMQRT(SIDE,TRANS,M,N,K,NB,V,LDV,T,LDT,C,LDC,WORK,INFO) */
void dgeqrt_(const int* m, const int* n, const int* nb, double* A, const int* lda, double* T, const int* ldt, double* work,
             int* info);
/* SRC/dlatsqr.f:170 DLATSQR(M,N,MB,NB,A,LDA,T,LDT,WORK,LWORK,INFO) (LAPACKE/include/lapack.h dlatsqr_) */
void dlatsqr_(const int* m, const int* n, const int* mb, const int* nb, double* A, const int* lda, double* T, const int* ldt,
              double* work, const int* lwork, int* info);
/* SRC/dgeqrt3.f:129 DGEQRT3(M,N,A,LDA,T,LDT,INFO) (LAPACKE/include/lapack.h dgeqrt3_) */
void dgeqrt3_(const int* m, const int* n, double* A, const int* lda, double* T, const int* ldt, int* info);
void dgemqrt_(const char* side, const char* trans, const int* m, const int* n, const int* k, const int* nb, const double* V,
              const int* ldv, const double* T, const int* ldt, double* C, const int* ldc, double* work, int* info, size_t, size_t);
/* SRC/dgerfs.f:183 DGERFS(TRANS,N,NRHS,A,LDA,AF,LDAF,IPIV,B,LDB,X,LDX,FERR,BERR,WORK,IWORK,INFO) */
void dgerfs_(const char* trans, const int* n, const int* nrhs, const double* A, const int* lda, const double* AF, const int* ldaf,
             const int* ipiv, const double* B, const int* ldb, double* X, const int* ldx, double* ferr, double* berr, double* work,
             int* iwork, int* info, size_t);
/* SRC/dgetri.f:114 DGETRI(N,A,LDA,IPIV,WORK,LWORK,INFO) (lapack.h LAPACK_dgetri) */
void dgetri_(const int* n, double* A, const int* lda, const int* ipiv, double* work, const int* lwork, int* info);
/* SRC/dorgqr.f:126 DORGQR(M,N,K,A,LDA,TAU,WORK,LWORK,INFO); SRC/dormqr.f:165 DORMQR(SIDE,TRANS,M,N,K,A,LDA,TAU,C,LDC,WORK,
   LWORK,INFO) (lapack.h:11807-11813, 12061-12078) */
void dorgqr_(const int* m, const int* n, const int* k, double* A, const int* lda, const double* tau, double* work,
             const int* lwork, int* info);
void dormqr_(const char* side, const char* trans, const int* m, const int* n, const int* k, const double* A, const int* lda,
             const double* tau, double* C, const int* ldc, double* work, const int* lwork, int* info, size_t, size_t);
void dlarft_(const char* direct, const char* storev, const int* n, const int* k, const double* V, const int* ldv,
             const double* tau, double* T, const int* ldt, size_t, size_t);
void dlarfb_(const char* side, const char* trans, const char* direct, const char* storev, const int* m, const int* n,
             const int* k, const double* V, const int* ldv, const double* T, const int* ldt, double* C, const int* ldc,
             double* work, const int* ldwork, size_t, size_t, size_t, size_t);

/* ---- condition estimation and the expert driver ---- */
/* SRC/dlatrs.f:238 DLATRS(UPLO,TRANS,DIAG,NORMIN,N,A,LDA,X,SCALE,CNORM,INFO) (lapack.h dlatrs_) */
void dlatrs_(const char* uplo, const char* trans, const char* diag, const char* normin, const int* n, const double* A, const int* lda,
             double* x, double* scale, double* cnorm, int* info, size_t, size_t, size_t, size_t);
/* SRC/dgecon.f:128 DGECON(NORM,N,A,LDA,ANORM,RCOND,WORK,IWORK,INFO) (lapack.h dgecon_) */
void dgecon_(const char* norm, const int* n, const double* A, const int* lda, const double* anorm, double* rcond, double* work,
             int* iwork, int* info, size_t);
/* SRC/dgeequ.f:139 DGEEQU(M,N,A,LDA,R,C,ROWCND,COLCND,AMAX,INFO); SRC/dlaqge.f:140 DLAQGE(M,N,A,LDA,R,C,ROWCND,COLCND,AMAX,EQUED) */
void dgeequ_(const int* m, const int* n, const double* A, const int* lda, double* r, double* c, double* rowcnd, double* colcnd,
             double* amax, int* info);
void dlaqge_(const int* m, const int* n, double* A, const int* lda, const double* r, const double* c, const double* rowcnd,
             const double* colcnd, const double* amax, char* equed, size_t);
/* SRC/dgesvx.f:344 DGESVX(FACT,TRANS,N,NRHS,A,LDA,AF,LDAF,IPIV,EQUED,R,C,B,LDB,X,LDX,RCOND,FERR,BERR,WORK,IWORK,INFO) (lapack.h dgesvx_) */
void dgesvx_(const char* fact, const char* trans, const int* n, const int* nrhs, double* A, const int* lda, double* AF, const int* ldaf,
             int* ipiv, char* equed, double* R, double* C, double* B, const int* ldb, double* X, const int* ldx, double* rcond,
             double* ferr, double* berr, double* work, int* iwork, int* info, size_t, size_t, size_t);

#ifdef __cplusplus
}
#endif
#endif
