// cblas_api.cu -- CBLAS Level-3 entry points over the Fortran-77 symbols of this library.
//
// Replaces CBLAS/src/cblas_dgemm.c:38-108, cblas_dsyrk.c:36-107, cblas_dtrmm.c:39-152, cblas_dtrsm.c:39-153: option
// enums are translated to the Fortran characters; for CblasRowMajor the row-major problem is rewritten as the
// equivalent column-major one on the same memory (C^T = op(B)^T op(A)^T: operands and M/N swapped; triangular and
// symmetric operands: UPLO flipped, SIDE flipped, TRANS flipped for SYRK), which is what the reference does.
// Illegal enum values go to cblas_xerbla with the reference's argument positions; dimension / leading-dimension
// errors are detected by the Fortran-level routine and reported through xerbla_ with its (column-major) positions.
#include <cstdarg>
#include <cstddef>
#include <cstdio>

#include "../../include/lapack_b200_cblas.h"
#include "../../include/lapack_b200_f77.h"

extern "C" {

__attribute__((weak)) void cblas_xerbla(int p, const char* rout, const char* form, ...) {
    va_list ap;
    va_start(ap, form);
    if (p) fprintf(stderr, "Parameter %d to routine %s was incorrect\n", p, rout);
    vfprintf(stderr, form, ap);
    va_end(ap);
}

static bool tr_char(CBLAS_TRANSPOSE t, char* c) {
    if (t == CblasNoTrans) *c = 'N'; else if (t == CblasTrans) *c = 'T'; else if (t == CblasConjTrans) *c = 'C'; else return false;
    return true;
}

void cblas_dgemm(CBLAS_LAYOUT layout, CBLAS_TRANSPOSE TransA, CBLAS_TRANSPOSE TransB, const int M, const int N, const int K,
                 const double alpha, const double* A, const int lda, const double* B, const int ldb, const double beta,
                 double* C, const int ldc) {
    char TA, TB;
    if (layout != CblasColMajor && layout != CblasRowMajor) { cblas_xerbla(1, "cblas_dgemm", "Illegal layout setting, %d\n", layout); return; }
    if (!tr_char(TransA, &TA)) { cblas_xerbla(2, "cblas_dgemm", "Illegal TransA setting, %d\n", TransA); return; }
    if (!tr_char(TransB, &TB)) { cblas_xerbla(3, "cblas_dgemm", "Illegal TransB setting, %d\n", TransB); return; }
    if (layout == CblasColMajor) dgemm_(&TA, &TB, &M, &N, &K, &alpha, A, &lda, B, &ldb, &beta, C, &ldc, 1, 1);
    else dgemm_(&TB, &TA, &N, &M, &K, &alpha, B, &ldb, A, &lda, &beta, C, &ldc, 1, 1);       // cblas_dgemm.c:78-104
}

void cblas_dsyrk(CBLAS_LAYOUT layout, CBLAS_UPLO Uplo, CBLAS_TRANSPOSE Trans, const int N, const int K, const double alpha,
                 const double* A, const int lda, const double beta, double* C, const int ldc) {
    char UL, TR;
    if (layout != CblasColMajor && layout != CblasRowMajor) { cblas_xerbla(1, "cblas_dsyrk", "Illegal layout setting, %d\n", layout); return; }
    const bool row = (layout == CblasRowMajor);
    if (Uplo == CblasUpper) UL = row ? 'L' : 'U'; else if (Uplo == CblasLower) UL = row ? 'U' : 'L';
    else { cblas_xerbla(2, "cblas_dsyrk", "Illegal Uplo setting, %d\n", Uplo); return; }
    if (Trans == CblasNoTrans) TR = row ? 'T' : 'N'; else if (Trans == CblasTrans || Trans == CblasConjTrans) TR = row ? 'N' : 'T';
    else { cblas_xerbla(3, "cblas_dsyrk", "Illegal Trans setting, %d\n", Trans); return; }
    dsyrk_(&UL, &TR, &N, &K, &alpha, A, &lda, &beta, C, &ldc, 1, 1);                          // cblas_dsyrk.c:48-104
}

static bool tri_opts(const char* name, bool row, CBLAS_SIDE Side, CBLAS_UPLO Uplo, CBLAS_TRANSPOSE TransA, CBLAS_DIAG Diag, char* SD,
                     char* UL, char* TA, char* DI) {
    if (Side == CblasRight) *SD = row ? 'L' : 'R'; else if (Side == CblasLeft) *SD = row ? 'R' : 'L';
    else { cblas_xerbla(2, name, "Illegal Side setting, %d\n", Side); return false; }
    if (Uplo == CblasUpper) *UL = row ? 'L' : 'U'; else if (Uplo == CblasLower) *UL = row ? 'U' : 'L';
    else { cblas_xerbla(3, name, "Illegal Uplo setting, %d\n", Uplo); return false; }
    if (!tr_char(TransA, TA)) { cblas_xerbla(4, name, "Illegal Trans setting, %d\n", TransA); return false; }
    if (Diag == CblasUnit) *DI = 'U'; else if (Diag == CblasNonUnit) *DI = 'N';
    else { cblas_xerbla(5, name, "Illegal Diag setting, %d\n", Diag); return false; }
    return true;
}

void cblas_dtrmm(CBLAS_LAYOUT layout, CBLAS_SIDE Side, CBLAS_UPLO Uplo, CBLAS_TRANSPOSE TransA, CBLAS_DIAG Diag, const int M,
                 const int N, const double alpha, const double* A, const int lda, double* B, const int ldb) {
    char SD, UL, TA, DI;
    if (layout != CblasColMajor && layout != CblasRowMajor) { cblas_xerbla(1, "cblas_dtrmm", "Illegal layout setting, %d\n", layout); return; }
    const bool row = (layout == CblasRowMajor);
    if (!tri_opts("cblas_dtrmm", row, Side, Uplo, TransA, Diag, &SD, &UL, &TA, &DI)) return;
    if (!row) dtrmm_(&SD, &UL, &TA, &DI, &M, &N, &alpha, A, &lda, B, &ldb, 1, 1, 1, 1);
    else dtrmm_(&SD, &UL, &TA, &DI, &N, &M, &alpha, A, &lda, B, &ldb, 1, 1, 1, 1);             // cblas_dtrmm.c:96-149
}

void cblas_dtrsm(CBLAS_LAYOUT layout, CBLAS_SIDE Side, CBLAS_UPLO Uplo, CBLAS_TRANSPOSE TransA, CBLAS_DIAG Diag, const int M,
                 const int N, const double alpha, const double* A, const int lda, double* B, const int ldb) {
    char SD, UL, TA, DI;
    if (layout != CblasColMajor && layout != CblasRowMajor) { cblas_xerbla(1, "cblas_dtrsm", "Illegal layout setting, %d\n", layout); return; }
    const bool row = (layout == CblasRowMajor);
    if (!tri_opts("cblas_dtrsm", row, Side, Uplo, TransA, Diag, &SD, &UL, &TA, &DI)) return;
    if (!row) dtrsm_(&SD, &UL, &TA, &DI, &M, &N, &alpha, A, &lda, B, &ldb, 1, 1, 1, 1);
    else dtrsm_(&SD, &UL, &TA, &DI, &N, &M, &alpha, A, &lda, B, &ldb, 1, 1, 1, 1);             // cblas_dtrsm.c:96-150
}

}  // extern "C"
