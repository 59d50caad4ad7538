/*
 * lapack_b200_f77_64.h -- the "_64" extended API of the same entry points (64-bit INTEGER arguments).
 *
 * The reference builds every routine a second time with 64-bit integers and a _64 suffix when BUILD_INDEX64_EXT_API is
 * on (the default: CMakeLists.txt:121-123, BLAS/SRC/CMakeLists.txt:126-134, SRC/CMakeLists.txt:543-554); C callers
 * reach them through LAPACK_GLOBAL_SUFFIX with API_SUFFIX(a) = a##_64 (LAPACKE/include/lapack.h:150-155), i.e. the
 * symbols dgetrf_64_, dgemm_64_, ...  Here they are thin forwarders to the 32-bit-index implementation: every
 * dimension must fit in 32 bits (the device kernels index rows with 32-bit integers and 64-bit offsets); a larger
 * value is reported as an illegal argument through xerbla_64_.  IPIV is a 64-bit array on this interface.
 */
#ifndef LAPACK_B200_F77_64_H
#define LAPACK_B200_F77_64_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* weak: an application-supplied xerbla_64_ takes precedence; the default forwards to xerbla_ */
void xerbla_64_(const char* srname, const int64_t* info, size_t srname_len);

void dgemm_64_(const char* transa, const char* transb, const int64_t* m, const int64_t* n, const int64_t* k, const double* alpha,
               const double* A, const int64_t* lda, const double* B, const int64_t* ldb, const double* beta, double* C,
               const int64_t* ldc, size_t, size_t);
void dsyrk_64_(const char* uplo, const char* trans, const int64_t* n, const int64_t* k, const double* alpha, const double* A,
               const int64_t* lda, const double* beta, double* C, const int64_t* ldc, size_t, size_t);
void dtrsm_64_(const char* side, const char* uplo, const char* transa, const char* diag, const int64_t* m, const int64_t* n,
               const double* alpha, const double* A, const int64_t* lda, double* B, const int64_t* ldb, size_t, size_t, size_t,
               size_t);
void dtrmm_64_(const char* side, const char* uplo, const char* transa, const char* diag, const int64_t* m, const int64_t* n,
               const double* alpha, const double* A, const int64_t* lda, double* B, const int64_t* ldb, size_t, size_t, size_t,
               size_t);

void dgetrf_64_(const int64_t* m, const int64_t* n, double* A, const int64_t* lda, int64_t* ipiv, int64_t* info);
void dgetrf2_64_(const int64_t* m, const int64_t* n, double* A, const int64_t* lda, int64_t* ipiv, int64_t* info);
void dlaswp_64_(const int64_t* n, double* A, const int64_t* lda, const int64_t* k1, const int64_t* k2, const int64_t* ipiv,
                const int64_t* incx);
void dgetrs_64_(const char* trans, const int64_t* n, const int64_t* nrhs, const double* A, const int64_t* lda, const int64_t* ipiv,
                double* B, const int64_t* ldb, int64_t* info, size_t);
void dgesv_64_(const int64_t* n, const int64_t* nrhs, double* A, const int64_t* lda, int64_t* ipiv, double* B, const int64_t* ldb,
               int64_t* info);

void dpotrf_64_(const char* uplo, const int64_t* n, double* A, const int64_t* lda, int64_t* info, size_t);
void dpotrf2_64_(const char* uplo, const int64_t* n, double* A, const int64_t* lda, int64_t* info, size_t);
void dpotrs_64_(const char* uplo, const int64_t* n, const int64_t* nrhs, const double* A, const int64_t* lda, double* B,
                const int64_t* ldb, int64_t* info, size_t);
void dposv_64_(const char* uplo, const int64_t* n, const int64_t* nrhs, double* A, const int64_t* lda, double* B,
               const int64_t* ldb, int64_t* info, size_t);

void dgeqrf_64_(const int64_t* m, const int64_t* n, double* A, const int64_t* lda, double* tau, double* work,
                const int64_t* lwork, int64_t* info);
void dgeqr2_64_(const int64_t* m, const int64_t* n, double* A, const int64_t* lda, double* tau, double* work, int64_t* info);
void dlarft_64_(const char* direct, const char* storev, const int64_t* n, const int64_t* k, const double* V, const int64_t* ldv,
                const double* tau, double* T, const int64_t* ldt, size_t, size_t);
void dlarfb_64_(const char* side, const char* trans, const char* direct, const char* storev, const int64_t* m, const int64_t* n,
                const int64_t* k, const double* V, const int64_t* ldv, const double* T, const int64_t* ldt, double* C,
                const int64_t* ldc, double* work, const int64_t* ldwork, size_t, size_t, size_t, size_t);

#ifdef __cplusplus
}
#endif
#endif
