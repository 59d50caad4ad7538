// capi.cu -- device-pointer C ABI (lb200_* symbols declared in include/lapack_b200.h).
//
// Thin `extern "C"` shims over the internal lb:: API: plain pointers and sizes, no C++ or torch types.
// `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).  Calls are asynchronous on
// that stream; INFO words are device ints.  Return value: 0, or a negative CUDA-error code (-1000 - e).
#include "lb_internal.h"
#include "../../include/lapack_b200.h"

namespace lb {
double fp64_peak(cudaStream_t s, int kind, int warps_per_cta, int ctas_per_sm, int iters);
void gemm_set_config(int cfg);
void trsm_set_inverse_enabled(int on);
void gemm_set_splitk_balance(int on);
void gemm_set_tma(int on);
void gemm_profile(int enable);
void gemm_profile_read(double* total_ms, double* total_flops, long long* launches);
void getrf_set_params(int nb, int leaf, int lookahead);
void getrf_set_cluster_max(int c);
void getrf_set_big_leaf(int v);
void getrf_set_tall_rows(int r);
void getrf_set_cluster_fat(int v);
void getrf_set_thin(int mode, int min_rows);
void getrf_set_super(int nb);
void getrf_set_defer_left(int on, int tail_rows);
void batched_set_mode(int m);
void laswp_set_bulk(int on);
void geqrf_set_cluster_max(int c);
void potrf_set_params(int nb, int lookahead);
void geqrf_set_params(int nb, int lookahead);
void trsm_set_fewrhs_mode(int mode);
}  // namespace lb

static inline cudaStream_t S(void* s) { return (cudaStream_t)s; }
// return code of ONE call: only errors raised since the previous lb200_* call returned are reported, and they are consumed
static inline int rc() {
    const int e = lb::take_cuda_error();
    return e ? -1000 - e : 0;
}

extern "C" {

int lb200_version(void) { return 100; }
int lb200_last_cuda_error(void) { return lb::last_cuda_error(); }
void lb200_clear_cuda_error(void) { lb::clear_cuda_error(); }
unsigned long long lb200_launch_count(void) { return lb::g_launches; }
void lb200_reset_launch_count(void) { lb::g_launches = 0; }

double lb200_fp64_peak_tflops(void* stream, int kind, int warps_per_cta, int ctas_per_sm, int iters) {
    return lb::fp64_peak(S(stream), kind, warps_per_cta, ctas_per_sm, iters);
}
void lb200_set_gemm_config(int cfg) { lb::gemm_set_config(cfg); }
void lb200_set_trsm_inverse(int on) { lb::trsm_set_inverse_enabled(on); }
void lb200_set_gemm_splitk_balance(int on) { lb::gemm_set_splitk_balance(on); }
void lb200_set_gemm_tma(int on) { lb::gemm_set_tma(on); }
void lb200_profile_gemm(int enable) { lb::gemm_profile(enable); }
void lb200_profile_gemm_read(double* total_ms, double* total_flops, long long* launches) {
    lb::gemm_profile_read(total_ms, total_flops, launches);
}
#ifndef LB_MINIMAL
void lb200_set_getrf_params(int nb, int leaf, int lookahead) { lb::getrf_set_params(nb, leaf, lookahead); }
void lb200_set_getrf_cluster_max(int ctas) { lb::getrf_set_cluster_max(ctas); }
void lb200_set_getrf_big_leaf(int rows4) { lb::getrf_set_big_leaf(rows4); }
void lb200_set_getrf_tall_rows(int rows_per_cta) { lb::getrf_set_tall_rows(rows_per_cta); }
void lb200_set_getrf_cluster_fat(int on) { lb::getrf_set_cluster_fat(on); }
void lb200_set_getrf_thin(int mode, int min_rows) { lb::getrf_set_thin(mode, min_rows); }
void lb200_set_getrf_super(int nb) { lb::getrf_set_super(nb); }
void lb200_set_getrf_defer_left(int on, int tail_rows) { lb::getrf_set_defer_left(on, tail_rows); }
void lb200_set_batched_mode(int mode) { lb::batched_set_mode(mode); }
void lb200_set_laswp_bulk(int on) { lb::laswp_set_bulk(on); }
void lb200_set_geqrf_cluster_max(int ctas) { lb::geqrf_set_cluster_max(ctas); }
void lb200_set_potrf_params(int nb, int lookahead) { lb::potrf_set_params(nb, lookahead); }
void lb200_set_geqrf_params(int nb, int lookahead) { lb::geqrf_set_params(nb, lookahead); }
void lb200_set_fewrhs_mode(int mode) { lb::trsm_set_fewrhs_mode(mode); }
int lb200_set_l2_fetch_granularity(int bytes) { return (int)cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)bytes); }
#endif

int lb200_dgemm(void* stream, char transa, char transb, int m, int n, int k, double alpha, const double* dA,
                long long lda, const double* dB, long long ldb, double beta, double* dC, long long ldc) {
    lb::gemm(S(stream), transa, transb, m, n, k, alpha, dA, lda, dB, ldb, beta, dC, ldc, 0);
    return rc();
}
int lb200_dsyrk(void* stream, char uplo, char trans, int n, int k, double alpha, const double* dA, long long lda,
                double beta, double* dC, long long ldc) {
    lb::syrk(S(stream), uplo, trans, n, k, alpha, dA, lda, beta, dC, ldc);
    return rc();
}
#ifndef LB_MINIMAL
int lb200_dtrsm(void* stream, char side, char uplo, char trans, char diag, int m, int n, double alpha,
                const double* dA, long long lda, double* dB, long long ldb) {
    lb::trsm(S(stream), side, uplo, trans, diag, m, n, alpha, dA, lda, dB, ldb);
    return rc();
}
int lb200_dtrmm(void* stream, char side, char uplo, char trans, char diag, int m, int n, double alpha,
                const double* dA, long long lda, double* dB, long long ldb) {
    lb::trmm(S(stream), side, uplo, trans, diag, m, n, alpha, dA, lda, dB, ldb);
    return rc();
}
int lb200_dlaswp(void* stream, int n, double* dA, long long lda, int k1, int k2, const int* dipiv, int incx) {
    lb::laswp(S(stream), n, dA, lda, k1, k2, dipiv, incx);
    return rc();
}
int lb200_dgetrf(void* stream, int m, int n, double* dA, long long lda, int* dipiv, int* dinfo) {
    lb::getrf(S(stream), m, n, dA, lda, dipiv, dinfo);
    return rc();
}
int lb200_dgetrf2(void* stream, int m, int n, double* dA, long long lda, int* dipiv, int* dinfo) {
    lb::getrf2(S(stream), m, n, dA, lda, dipiv, dinfo);
    return rc();
}
int lb200_dgetrs(void* stream, char trans, int n, int nrhs, const double* dA, long long lda, const int* dipiv,
                 double* dB, long long ldb) {
    lb::getrs(S(stream), trans, n, nrhs, dA, lda, dipiv, dB, ldb);
    return rc();
}
int lb200_dpotrf(void* stream, char uplo, int n, double* dA, long long lda, int* dinfo) {
    lb::potrf(S(stream), uplo, n, dA, lda, dinfo);
    return rc();
}
int lb200_dpotrf2(void* stream, char uplo, int n, double* dA, long long lda, int* dinfo) {
    lb::potrf2(S(stream), uplo, n, dA, lda, dinfo);
    return rc();
}
int lb200_dpotrs(void* stream, char uplo, int n, int nrhs, const double* dA, long long lda, double* dB,
                 long long ldb) {
    lb::potrs(S(stream), uplo, n, nrhs, dA, lda, dB, ldb);
    return rc();
}
int lb200_dgeqrf(void* stream, int m, int n, double* dA, long long lda, double* dtau) {
    lb::geqrf(S(stream), m, n, dA, lda, dtau);
    return rc();
}
int lb200_dgeqr2(void* stream, int m, int n, double* dA, long long lda, double* dtau) {
    lb::geqr2(S(stream), m, n, dA, lda, dtau);
    return rc();
}
int lb200_dlarft(void* stream, int n, int k, const double* dV, long long ldv, const double* dtau, double* dT,
                 long long ldt) {
    lb::larft(S(stream), n, k, dV, ldv, dtau, dT, ldt);
    return rc();
}
int lb200_dlarfb(void* stream, char side, char trans, int m, int n, int k, const double* dV, long long ldv,
                 const double* dT, long long ldt, double* dC, long long ldc) {
    lb::larfb(S(stream), side, trans, m, n, k, dV, ldv, dT, ldt, dC, ldc);
    return rc();
}
int lb200_dgetri(void* stream, int n, double* dA, long long lda, const int* dipiv, int* dinfo) {
    lb::getri(S(stream), n, dA, lda, dipiv, dinfo);
    return rc();
}
int lb200_dgeqrt(void* stream, int m, int n, int nb, double* dA, long long lda, double* dT, long long ldt) {
    lb::geqrt(S(stream), m, n, nb, dA, lda, dT, ldt);
    return rc();
}
int lb200_dgemqrt(void* stream, char side, char trans, int m, int n, int k, int nb, const double* dV, long long ldv,
                  const double* dT, long long ldt, double* dC, long long ldc) {
    lb::gemqrt(S(stream), side, trans, m, n, k, nb, dV, ldv, dT, ldt, dC, ldc);
    return rc();
}
int lb200_dormqr(void* stream, char side, char trans, int m, int n, int k, const double* dA, long long lda, const double* dtau,
                 double* dC, long long ldc) {
    lb::ormqr(S(stream), side, trans, m, n, k, dA, lda, dtau, dC, ldc);
    return rc();
}
int lb200_dorgqr(void* stream, int m, int n, int k, double* dA, long long lda, const double* dtau) {
    lb::orgqr(S(stream), m, n, k, dA, lda, dtau);
    return rc();
}
int lb200_dgetrf_batched32(void* stream, long long batch, double* dA, int* dipiv, int* dinfo) {
    lb::getrf_batched_32(S(stream), batch, dA, dipiv, dinfo);
    return rc();
}
int lb200_dpotrf_batched32(void* stream, char uplo, long long batch, double* dA, int* dinfo) {
    lb::potrf_batched_32(S(stream), uplo, batch, dA, dinfo);
    return rc();
}
int lb200_dlarnv_matrix(void* stream, const int iseed[4], long long stream_offset, int m, int n, double* dA,
                        long long lda) {
    lb::larnv_matrix(S(stream), iseed, stream_offset, m, n, dA, lda);
    return rc();
}
int lb200_dlarnv_submatrix(void* stream, const int iseed[4], long long stream_offset, long long stream_ld, int m, int n,
                           double* dA, long long lda) {
    lb::larnv_submatrix(S(stream), iseed, stream_offset, stream_ld, m, n, dA, lda);
    return rc();
}
int lb200_laswp_compose(void* stream, int np, const int* dipiv_rel, int* dsrc_top, int* dinv_top) {
    if (np > 2048) return -1;
    lb::laswp_compose(S(stream), np, dipiv_rel, dsrc_top, dinv_top);
    return rc();
}
int lb200_gather_rows(void* stream, int nidx, const int* didx, const double* dA, long long lda, int ncols, double* dW, long long ldw) {
    lb::gather_rows(S(stream), nidx, didx, dA, lda, ncols, dW, ldw);
    return rc();
}
int lb200_scatter_rows(void* stream, int nidx, const int* didx, const double* dW, long long ldw, int ncols, double* dA, long long lda) {
    lb::scatter_rows(S(stream), nidx, didx, dW, ldw, ncols, dA, lda);
    return rc();
}
int lb200_make_spd(void* stream, int n, double* dA, long long lda, double shift) {
    lb::make_spd(S(stream), n, dA, lda, shift);
    return rc();
}
int lb200_dlacpy(void* stream, char uplo, int m, int n, const double* dA, long long lda, double* dB, long long ldb) {
    lb::lacpy(S(stream), uplo, m, n, dA, lda, dB, ldb);
    return rc();
}
int lb200_transpose(void* stream, int m, int n, const double* dA, long long lda, double* dB, long long ldb) {
    lb::transpose(S(stream), m, n, dA, lda, dB, ldb);
    return rc();
}

#endif  // LB_MINIMAL

}  // extern "C"
