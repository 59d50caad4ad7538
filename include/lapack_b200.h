/*
 * lapack_b200.h -- C ABI of liblapack_b200.so (B200 / sm_100a implementation of LAPACK's blocked
 * one-sided factorizations).  Three layers, all `extern "C"`, plain pointers and sizes:
 *
 *  1. lb200_*  : device-pointer API.  Every matrix/vector/IPIV/INFO pointer is a DEVICE pointer,
 *                `stream` is a cudaStream_t passed as void*; calls are asynchronous on that stream.
 *  2. Fortran-77 ABI symbols (dgetrf_, dpotrf_, dgeqrf_, ... see lapack_b200_f77.h): drop-in for the
 *                reference's SRC/ and BLAS/SRC symbols; pointers may be host or device memory.
 *  3. LAPACKE_* : drop-in for the reference's LAPACKE C entry points (lapack_b200_lapacke.h).
 *
 * Each declaration cites the reference interface it replaces (paths relative to the reference tree).
 */
#ifndef LAPACK_B200_H
#define LAPACK_B200_H
#ifdef __cplusplus
extern "C" {
#endif

int lb200_version(void);
int lb200_last_cuda_error(void);
void lb200_clear_cuda_error(void);
unsigned long long lb200_launch_count(void);      /* kernels launched by this library so far */
void lb200_reset_launch_count(void);

/* measured FP64 pipe peak (TFLOP/s): kind 0 = DMMA.8x8x4 issue rate, 1 = DFMA */
double lb200_fp64_peak_tflops(void* stream, int kind, int warps_per_cta, int ctas_per_sm, int iters);

/* per-launch CUDA-event timing of the large trailing-update GEMMs (used by bench.py's roofline leg):
 * enable(1) resets the counters; read() returns the summed durations / algorithmic flops / launch count */
void lb200_profile_gemm(int enable);
void lb200_profile_gemm_read(double* total_ms, double* total_flops, long long* launches);

/* tuning knobs (testing / benchmarking; defaults are chosen per problem size) */
void lb200_set_gemm_splitk_balance(int on);   /* 1 (default): long-K GEMMs whose tile count leaves a partly filled last wave are split along K */
void lb200_set_trsm_inverse(int on);   /* 1 (default): the panel solve inside DPOTRF uses inverted 32x32 diagonal blocks (DMMA leaves; 377.5 -> 373.9 ms at n=32768); 0: substitution leaves everywhere */
void lb200_set_gemm_config(int cfg);   /* -1 auto, 0/1/2 force a cp.async tile shape, 3 force the TMA kernel */
void lb200_set_gemm_tma(int on);       /* 0 disables the TMA fast path (falls back to cp.async) */
void lb200_set_getrf_params(int nb, int leaf, int lookahead);
/* panels of at most ctas*1024 rows are factored by the thread-block-cluster leaf kernel (default 8, 0 = never, max 16) */
void lb200_set_getrf_cluster_max(int ctas);
/* panels too tall for one cluster: 1 (default) = 256-thread x 4-row leaf kernel, 0 = 1024-thread x 1-row kernel */
void lb200_set_getrf_big_leaf(int rows4);
/* rows per CTA (1024 default / 2048 / 4096) of the leaf kernel for panels too tall for one cluster: fewer, fatter CTAs hold fewer SMs */
void lb200_set_getrf_tall_rows(int rows_per_cta);
/* 1: panels of 16385..32768 rows are factored by ONE thread-block cluster of <= 16 CTAs with 8 rows per thread (measured neutral: 852 vs 850 ms) */
void lb200_set_getrf_cluster_fat(int on);
/* thin LU leaves for panels taller than min_rows: 0 = off, 1 = 128 threads x 2 rows per CTA, 2 = 64 threads x 4 rows (one GEMM-CTA slot each) */
void lb200_set_getrf_thin(int mode, int min_rows);
/* DGETRF two-level driver: outer block size (default 4096; 0 = single level); used for min(m,n) >= 3*nb, m <= 56320, device-resident callers */
void lb200_set_getrf_super(int nb);
/* DGETRF: 1 (default) = the interchanges left of the panel are applied in the tail of the factorization (trailing matrix <= tail_rows rows) */
void lb200_set_getrf_defer_left(int on, int tail_rows);
/* batched 32x32 DGETRF: 2 (default) = two matrices per warp, 0 = one matrix per warp, 1 = persistent software-pipelined one-matrix kernel (measured slower) */
void lb200_set_batched_mode(int mode);
/* DLASWP apply kernel: 1 = scattered row reads as 16-byte cp.async.bulk copies instead of LDG (experiment) */
void lb200_set_laswp_bulk(int on);
void lb200_set_geqrf_cluster_max(int ctas);
void lb200_set_potrf_params(int nb, int lookahead);
void lb200_set_geqrf_params(int nb, int lookahead);
/* triangular solves with <= 8 right-hand sides: 1 (default) = one persistent streaming kernel, 0 = leaf/GEMV recursion */
void lb200_set_fewrhs_mode(int mode);
/* experiment knob: cudaLimitMaxL2FetchGranularity (32 / 64 / 128 bytes) for the LDA-strided row interchanges */
int lb200_set_l2_fetch_granularity(int bytes);

/* BLAS/SRC/dgemm.f:187 DGEMM(TRANSA,TRANSB,M,N,K,ALPHA,A,LDA,B,LDB,BETA,C,LDC) */
int lb200_dgemm(void* stream, char transa, char transb, int m, int n, int k, double alpha, const double* dA,
                long long lda, const double* dB, long long ldb, double beta, double* dC, long long ldc);
/* BLAS/SRC/dsyrk.f:168 DSYRK(UPLO,TRANS,N,K,ALPHA,A,LDA,BETA,C,LDC) */
int lb200_dsyrk(void* stream, char uplo, char trans, int n, int k, double alpha, const double* dA, long long lda,
                double beta, double* dC, long long ldc);
/* BLAS/SRC/dtrsm.f:180 DTRSM(SIDE,UPLO,TRANSA,DIAG,M,N,ALPHA,A,LDA,B,LDB) */
int lb200_dtrsm(void* stream, char side, char uplo, char trans, char diag, int m, int n, double alpha,
                const double* dA, long long lda, double* dB, long long ldb);
/* BLAS/SRC/dtrmm.f:176 DTRMM(SIDE,UPLO,TRANSA,DIAG,M,N,ALPHA,A,LDA,B,LDB) */
int lb200_dtrmm(void* stream, char side, char uplo, char trans, char diag, int m, int n, double alpha,
                const double* dA, long long lda, double* dB, long long ldb);
/* SRC/dlaswp.f:112 DLASWP(N,A,LDA,K1,K2,IPIV,INCX); ipiv is a device array of 1-based rows */
int lb200_dlaswp(void* stream, int n, double* dA, long long lda, int k1, int k2, const int* dipiv, int incx);

/* SRC/dgetrf.f:105 DGETRF(M,N,A,LDA,IPIV,INFO); SRC/dgetrf2.f:112 */
int lb200_dgetrf(void* stream, int m, int n, double* dA, long long lda, int* dipiv, int* dinfo);
int lb200_dgetrf2(void* stream, int m, int n, double* dA, long long lda, int* dipiv, int* dinfo);
/* SRC/dgetrs.f:118 DGETRS(TRANS,N,NRHS,A,LDA,IPIV,B,LDB,INFO) */
int lb200_dgetrs(void* stream, char trans, int n, int nrhs, const double* dA, long long lda, const int* dipiv,
                 double* dB, long long ldb);
/* SRC/dpotrf.f:104 DPOTRF(UPLO,N,A,LDA,INFO); SRC/dpotrf2.f:105; SRC/dpotrs.f:107 */
int lb200_dpotrf(void* stream, char uplo, int n, double* dA, long long lda, int* dinfo);
int lb200_dpotrf2(void* stream, char uplo, int n, double* dA, long long lda, int* dinfo);
int lb200_dpotrs(void* stream, char uplo, int n, int nrhs, const double* dA, long long lda, double* dB,
                 long long ldb);
/* SRC/dgeqrf.f:145 DGEQRF(M,N,A,LDA,TAU,WORK,LWORK,INFO) (device scratch replaces WORK); SRC/dgeqr2.f:127 */
int lb200_dgeqrf(void* stream, int m, int n, double* dA, long long lda, double* dtau);
int lb200_dgeqr2(void* stream, int m, int n, double* dA, long long lda, double* dtau);
/* SRC/dlarft.f:160 DLARFT('F','C',N,K,V,LDV,TAU,T,LDT);  SRC/dlarfb.f:192 DLARFB(SIDE,TRANS,'F','C',...) */
int lb200_dlarft(void* stream, int n, int k, const double* dV, long long ldv, const double* dtau, double* dT,
                 long long ldt);
int lb200_dlarfb(void* stream, char side, char trans, int m, int n, int k, const double* dV, long long ldv,
                 const double* dT, long long ldt, double* dC, long long ldc);
/* SRC/dgetri.f:114 DGETRI on device pointers: A holds the DGETRF factors, overwritten by inv(A); *dinfo = i if U(i,i) == 0 */
int lb200_dgetri(void* stream, int n, double* dA, long long lda, const int* dipiv, int* dinfo);
/* SRC/dgeqrt.f:139 DGEQRT (T is nb x min(m,n)), SRC/dgemqrt.f:166 DGEMQRT on device pointers */
int lb200_dgeqrt(void* stream, int m, int n, int nb, double* dA, long long lda, double* dT, long long ldt);
int lb200_dgemqrt(void* stream, char side, char trans, int m, int n, int k, int nb, const double* dV, long long ldv,
                  const double* dT, long long ldt, double* dC, long long ldc);
/* SRC/dormqr.f:165, SRC/dorgqr.f:126 on device pointers */
int lb200_dormqr(void* stream, char side, char trans, int m, int n, int k, const double* dA, long long lda, const double* dtau,
                 double* dC, long long ldc);
int lb200_dorgqr(void* stream, int m, int n, int k, double* dA, long long lda, const double* dtau);
/* batched 32x32 (config C5b): matrices contiguous, stride 1024 doubles; ipiv 32 ints per matrix */
int lb200_dgetrf_batched32(void* stream, long long batch, double* dA, int* dipiv, int* dinfo);
int lb200_dpotrf_batched32(void* stream, char uplo, long long batch, double* dA, int* dinfo);

/* SRC/dlarnv.f:97 + SRC/dlaruv.f:95: fill A (m x n, column by column) with U(-1,1) from the 48-bit LCG,
 * starting `stream_offset` draws after `iseed` (O(log n) jump-ahead, bit-identical to DLARNV(2)). */
int lb200_dlarnv_matrix(void* stream, const int iseed[4], long long stream_offset, int m, int n, double* dA,
                        long long lda);
/* window of a global DLARNV matrix with stream_ld rows: A(i,j) = draw (stream_offset + j*stream_ld + i) */
int lb200_dlarnv_submatrix(void* stream, const int iseed[4], long long stream_offset, long long stream_ld, int m, int n,
                           double* dA, long long lda);
/* building blocks of the P x Q distributed DGETRF (lapack_b200/dist2d.py): one panel's interchanges (SRC/dlaswp.f:152-167 with
 * DGETRF's pivots, relative to the panel's first row, 1-based, np <= 2048) composed into src_top[t] = relative row whose
 * content ends in panel row t and inv_top[t] = relative row where the original panel row t ends; row gather / scatter
 * between a column-major matrix and a packed np x ncols buffer (idx[t] < 0 skips the row) */
int lb200_laswp_compose(void* stream, int np, const int* dipiv_rel, int* dsrc_top, int* dinv_top);
int lb200_gather_rows(void* stream, int nidx, const int* didx, const double* dA, long long lda, int ncols, double* dW, long long ldw);
int lb200_scatter_rows(void* stream, int nidx, const int* didx, const double* dW, long long ldw, int ncols, double* dA, long long lda);
int lb200_make_spd(void* stream, int n, double* dA, long long lda, double shift); /* A := (A+A')/2 + shift*I */
int lb200_dlacpy(void* stream, char uplo, int m, int n, const double* dA, long long lda, double* dB, long long ldb);
int lb200_transpose(void* stream, int m, int n, const double* dA, long long lda, double* dB, long long ldb);

#ifdef __cplusplus
}
#endif
#endif
