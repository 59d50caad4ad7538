/* lapack_b200_cblas.h -- CBLAS entry points over the B200 Level-3 kernels.
 *
 * Same prototypes and enum values as the reference's CBLAS/include/cblas.h (CBLAS_LAYOUT :43, CBLAS_TRANSPOSE :44,
 * CBLAS_UPLO :45, CBLAS_DIAG :46, CBLAS_SIDE :47; cblas_dgemm :545, cblas_dsyrk :556, cblas_dtrmm :565,
 * cblas_dtrsm :570).  They replace CBLAS/src/cblas_dgemm.c, cblas_dsyrk.c, cblas_dtrmm.c, cblas_dtrsm.c: the
 * row-major cases are mapped onto the column-major Fortran symbols exactly as those files do (operands swapped,
 * UPLO / SIDE / TRANS flipped), so no data is transposed.  Pointers may be host or device memory.
 */
#ifndef LAPACK_B200_CBLAS_H
#define LAPACK_B200_CBLAS_H
#ifdef __cplusplus
extern "C" {
#endif

#ifndef CBLAS_H
typedef enum CBLAS_LAYOUT { CblasRowMajor = 101, CblasColMajor = 102 } CBLAS_LAYOUT;
typedef enum CBLAS_TRANSPOSE { CblasNoTrans = 111, CblasTrans = 112, CblasConjTrans = 113 } CBLAS_TRANSPOSE;
typedef enum CBLAS_UPLO { CblasUpper = 121, CblasLower = 122 } CBLAS_UPLO;
typedef enum CBLAS_DIAG { CblasNonUnit = 131, CblasUnit = 132 } CBLAS_DIAG;
typedef enum CBLAS_SIDE { CblasLeft = 141, CblasRight = 142 } CBLAS_SIDE;
#endif

void cblas_dgemm(CBLAS_LAYOUT layout, CBLAS_TRANSPOSE TransA, CBLAS_TRANSPOSE TransB, const int M, const int N, const int K,
                 const double alpha, const double* A, const int lda, const double* B, const int ldb, const double beta,
                 double* C, const int ldc);
void cblas_dsyrk(CBLAS_LAYOUT layout, CBLAS_UPLO Uplo, CBLAS_TRANSPOSE Trans, const int N, const int K, const double alpha,
                 const double* A, const int lda, const double beta, double* C, const int ldc);
void cblas_dtrmm(CBLAS_LAYOUT layout, CBLAS_SIDE Side, CBLAS_UPLO Uplo, CBLAS_TRANSPOSE TransA, CBLAS_DIAG Diag, const int M,
                 const int N, const double alpha, const double* A, const int lda, double* B, const int ldb);
void cblas_dtrsm(CBLAS_LAYOUT layout, CBLAS_SIDE Side, CBLAS_UPLO Uplo, CBLAS_TRANSPOSE TransA, CBLAS_DIAG Diag, const int M,
                 const int N, const double alpha, const double* A, const int lda, double* B, const int ldb);
/* weak: an application-supplied cblas_xerbla (CBLAS/src/cblas_xerbla.c) takes precedence */
void cblas_xerbla(int p, const char* rout, const char* form, ...);

#ifdef __cplusplus
}
#endif
#endif
