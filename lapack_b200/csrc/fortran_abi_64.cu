// fortran_abi_64.cu -- "_64" extended API (64-bit INTEGER) of the Fortran-77 entry points: see include/lapack_b200_f77_64.h.
// Thin forwarders: arguments are narrowed to 32 bits (a value that does not fit is an illegal argument, reported through
// xerbla_64_ with the position the 32-bit routine would use), IPIV arrays are converted, INFO is widened.
#include <cstdint>
#include <cstring>
#include <climits>
#include <vector>
#include <cuda_runtime.h>

#include "../../include/lapack_b200_f77.h"
#include "../../include/lapack_b200_f77_64.h"

extern "C" __attribute__((weak)) void xerbla_64_(const char* srname, const int64_t* info, size_t len) {
    const int i32 = (int)*info;
    xerbla_(srname, &i32, len);
}

namespace {

struct Narrow {
    const char* name;
    bool ok = true;
    explicit Narrow(const char* n) : name(n) {}
    // pos = 1-based argument position reported if the value does not fit
    int operator()(const int64_t* v, int pos) {
        if (*v > INT_MAX || *v < INT_MIN) {
            if (ok) { int64_t p = pos; xerbla_64_(name, &p, strlen(name)); }
            ok = false;
            return 0;
        }
        return (int)*v;
    }
};

// IPIV conversion through the host (cudaMemcpyDefault accepts host and device pointers alike)
std::vector<int> ipiv_in(const int64_t* p, size_t n) {
    std::vector<int64_t> w(n);
    std::vector<int> r(n);
    if (n) cudaMemcpy(w.data(), p, n * sizeof(int64_t), cudaMemcpyDefault);
    for (size_t i = 0; i < n; ++i) r[i] = (int)w[i];
    return r;
}
void ipiv_out(int64_t* p, const std::vector<int>& r) {
    std::vector<int64_t> w(r.size());
    for (size_t i = 0; i < r.size(); ++i) w[i] = r[i];
    if (!r.empty()) cudaMemcpy(p, w.data(), r.size() * sizeof(int64_t), cudaMemcpyDefault);
}
inline size_t zmin(int a, int b) { int v = a < b ? a : b; return v > 0 ? (size_t)v : 0; }

}  // namespace

extern "C" {

void dgemm_64_(const char* transa, const char* transb, const int64_t* m, const int64_t* n, const int64_t* k, const double* alpha,
               const double* A, const int64_t* lda, const double* B, const int64_t* ldb, const double* beta, double* C,
               const int64_t* ldc, size_t l1, size_t l2) {
    Narrow nw("DGEMM ");
    const int m_ = nw(m, 3), n_ = nw(n, 4), k_ = nw(k, 5), lda_ = nw(lda, 8), ldb_ = nw(ldb, 10), ldc_ = nw(ldc, 13);
    if (nw.ok) dgemm_(transa, transb, &m_, &n_, &k_, alpha, A, &lda_, B, &ldb_, beta, C, &ldc_, l1, l2);
}
void dsyrk_64_(const char* uplo, const char* trans, const int64_t* n, const int64_t* k, const double* alpha, const double* A,
               const int64_t* lda, const double* beta, double* C, const int64_t* ldc, size_t l1, size_t l2) {
    Narrow nw("DSYRK ");
    const int n_ = nw(n, 3), k_ = nw(k, 4), lda_ = nw(lda, 7), ldc_ = nw(ldc, 10);
    if (nw.ok) dsyrk_(uplo, trans, &n_, &k_, alpha, A, &lda_, beta, C, &ldc_, l1, l2);
}
void dtrsm_64_(const char* side, const char* uplo, const char* transa, const char* diag, const int64_t* m, const int64_t* n,
               const double* alpha, const double* A, const int64_t* lda, double* B, const int64_t* ldb, size_t l1, size_t l2,
               size_t l3, size_t l4) {
    Narrow nw("DTRSM ");
    const int m_ = nw(m, 5), n_ = nw(n, 6), lda_ = nw(lda, 9), ldb_ = nw(ldb, 11);
    if (nw.ok) dtrsm_(side, uplo, transa, diag, &m_, &n_, alpha, A, &lda_, B, &ldb_, l1, l2, l3, l4);
}
void dtrmm_64_(const char* side, const char* uplo, const char* transa, const char* diag, const int64_t* m, const int64_t* n,
               const double* alpha, const double* A, const int64_t* lda, double* B, const int64_t* ldb, size_t l1, size_t l2,
               size_t l3, size_t l4) {
    Narrow nw("DTRMM ");
    const int m_ = nw(m, 5), n_ = nw(n, 6), lda_ = nw(lda, 9), ldb_ = nw(ldb, 11);
    if (nw.ok) dtrmm_(side, uplo, transa, diag, &m_, &n_, alpha, A, &lda_, B, &ldb_, l1, l2, l3, l4);
}

static void getrf64(bool rec, const int64_t* m, const int64_t* n, double* A, const int64_t* lda, int64_t* ipiv, int64_t* info) {
    Narrow nw(rec ? "DGETRF2" : "DGETRF");
    const int m_ = nw(m, 1), n_ = nw(n, 2), lda_ = nw(lda, 4);
    if (!nw.ok) { *info = (*m > INT_MAX || *m < INT_MIN) ? -1 : (*n > INT_MAX || *n < INT_MIN) ? -2 : -4; return; }
    std::vector<int> p(zmin(m_, n_));
    int i32 = 0;
    if (rec) dgetrf2_(&m_, &n_, A, &lda_, p.data(), &i32); else dgetrf_(&m_, &n_, A, &lda_, p.data(), &i32);
    if (i32 >= 0) ipiv_out(ipiv, p);
    *info = i32;
}
void dgetrf_64_(const int64_t* m, const int64_t* n, double* A, const int64_t* lda, int64_t* ipiv, int64_t* info) {
    getrf64(false, m, n, A, lda, ipiv, info);
}
void dgetrf2_64_(const int64_t* m, const int64_t* n, double* A, const int64_t* lda, int64_t* ipiv, int64_t* info) {
    getrf64(true, m, n, A, lda, ipiv, info);
}
void dlaswp_64_(const int64_t* n, double* A, const int64_t* lda, const int64_t* k1, const int64_t* k2, const int64_t* ipiv,
                const int64_t* incx) {
    Narrow nw("DLASWP");
    const int n_ = nw(n, 1), lda_ = nw(lda, 3), k1_ = nw(k1, 4), k2_ = nw(k2, 5), inc_ = nw(incx, 7);
    if (!nw.ok || n_ <= 0 || inc_ == 0 || k2_ < k1_) return;
    const int ainc = inc_ > 0 ? inc_ : -inc_;
    const size_t np = (size_t)(k1_ - 1) + (size_t)(k2_ - k1_) * ainc + 1;      // entries read (dlaswp.f:99-103)
    std::vector<int> p = ipiv_in(ipiv, np);
    dlaswp_(&n_, A, &lda_, &k1_, &k2_, p.data(), &inc_);
}
void dgetrs_64_(const char* trans, const int64_t* n, const int64_t* nrhs, const double* A, const int64_t* lda, const int64_t* ipiv,
                double* B, const int64_t* ldb, int64_t* info, size_t l1) {
    Narrow nw("DGETRS");
    const int n_ = nw(n, 2), nrhs_ = nw(nrhs, 3), lda_ = nw(lda, 5), ldb_ = nw(ldb, 8);
    if (!nw.ok) { *info = -2; return; }
    std::vector<int> p = ipiv_in(ipiv, n_ > 0 ? (size_t)n_ : 0);
    int i32 = 0;
    dgetrs_(trans, &n_, &nrhs_, A, &lda_, p.data(), B, &ldb_, &i32, l1);
    *info = i32;
}
void dgesv_64_(const int64_t* n, const int64_t* nrhs, double* A, const int64_t* lda, int64_t* ipiv, double* B, const int64_t* ldb,
               int64_t* info) {
    Narrow nw("DGESV ");
    const int n_ = nw(n, 1), nrhs_ = nw(nrhs, 2), lda_ = nw(lda, 4), ldb_ = nw(ldb, 7);
    if (!nw.ok) { *info = -1; return; }
    std::vector<int> p(n_ > 0 ? (size_t)n_ : 0);
    int i32 = 0;
    dgesv_(&n_, &nrhs_, A, &lda_, p.data(), B, &ldb_, &i32);
    if (i32 >= 0) ipiv_out(ipiv, p);
    *info = i32;
}

void dpotrf_64_(const char* uplo, const int64_t* n, double* A, const int64_t* lda, int64_t* info, size_t l1) {
    Narrow nw("DPOTRF");
    const int n_ = nw(n, 2), lda_ = nw(lda, 4);
    if (!nw.ok) { *info = -2; return; }
    int i32 = 0;
    dpotrf_(uplo, &n_, A, &lda_, &i32, l1);
    *info = i32;
}
void dpotrf2_64_(const char* uplo, const int64_t* n, double* A, const int64_t* lda, int64_t* info, size_t l1) {
    Narrow nw("DPOTRF2");
    const int n_ = nw(n, 2), lda_ = nw(lda, 4);
    if (!nw.ok) { *info = -2; return; }
    int i32 = 0;
    dpotrf2_(uplo, &n_, A, &lda_, &i32, l1);
    *info = i32;
}
void dpotrs_64_(const char* uplo, const int64_t* n, const int64_t* nrhs, const double* A, const int64_t* lda, double* B,
                const int64_t* ldb, int64_t* info, size_t l1) {
    Narrow nw("DPOTRS");
    const int n_ = nw(n, 2), nrhs_ = nw(nrhs, 3), lda_ = nw(lda, 5), ldb_ = nw(ldb, 7);
    if (!nw.ok) { *info = -2; return; }
    int i32 = 0;
    dpotrs_(uplo, &n_, &nrhs_, A, &lda_, B, &ldb_, &i32, l1);
    *info = i32;
}
void dposv_64_(const char* uplo, const int64_t* n, const int64_t* nrhs, double* A, const int64_t* lda, double* B,
               const int64_t* ldb, int64_t* info, size_t l1) {
    Narrow nw("DPOSV ");
    const int n_ = nw(n, 2), nrhs_ = nw(nrhs, 3), lda_ = nw(lda, 5), ldb_ = nw(ldb, 7);
    if (!nw.ok) { *info = -2; return; }
    int i32 = 0;
    dposv_(uplo, &n_, &nrhs_, A, &lda_, B, &ldb_, &i32, l1);
    *info = i32;
}

void dgeqrf_64_(const int64_t* m, const int64_t* n, double* A, const int64_t* lda, double* tau, double* work,
                const int64_t* lwork, int64_t* info) {
    Narrow nw("DGEQRF");
    const int m_ = nw(m, 1), n_ = nw(n, 2), lda_ = nw(lda, 4);
    // a workspace larger than 2^31-1 doubles is simply "large enough" for the 32-bit routine
    const int lwork_ = (*lwork > INT_MAX) ? INT_MAX : nw(lwork, 7);
    if (!nw.ok) { *info = -1; return; }
    int i32 = 0;
    dgeqrf_(&m_, &n_, A, &lda_, tau, work, &lwork_, &i32);
    *info = i32;
}
void dgeqr2_64_(const int64_t* m, const int64_t* n, double* A, const int64_t* lda, double* tau, double* work, int64_t* info) {
    Narrow nw("DGEQR2");
    const int m_ = nw(m, 1), n_ = nw(n, 2), lda_ = nw(lda, 4);
    if (!nw.ok) { *info = -1; return; }
    int i32 = 0;
    dgeqr2_(&m_, &n_, A, &lda_, tau, work, &i32);
    *info = i32;
}
void dlarft_64_(const char* direct, const char* storev, const int64_t* n, const int64_t* k, const double* V, const int64_t* ldv,
                const double* tau, double* T, const int64_t* ldt, size_t l1, size_t l2) {
    Narrow nw("DLARFT");
    const int n_ = nw(n, 3), k_ = nw(k, 4), ldv_ = nw(ldv, 6), ldt_ = nw(ldt, 9);
    if (nw.ok) dlarft_(direct, storev, &n_, &k_, V, &ldv_, tau, T, &ldt_, l1, l2);
}
void dlarfb_64_(const char* side, const char* trans, const char* direct, const char* storev, const int64_t* m, const int64_t* n,
                const int64_t* k, const double* V, const int64_t* ldv, const double* T, const int64_t* ldt, double* C,
                const int64_t* ldc, double* work, const int64_t* ldwork, size_t l1, size_t l2, size_t l3, size_t l4) {
    Narrow nw("DLARFB");
    const int m_ = nw(m, 5), n_ = nw(n, 6), k_ = nw(k, 7), ldv_ = nw(ldv, 9), ldt_ = nw(ldt, 11), ldc_ = nw(ldc, 13);
    const int ldw_ = (*ldwork > INT_MAX) ? INT_MAX : nw(ldwork, 15);
    if (nw.ok) dlarfb_(side, trans, direct, storev, &m_, &n_, &k_, V, &ldv_, T, &ldt_, C, &ldc_, work, &ldw_, l1, l2, l3, l4);
}

}  // extern "C"
