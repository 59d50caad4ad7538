// lb_internal.h -- internal device-side API of liblapack_b200 (sm_100a only).
//
// Everything here works on DEVICE pointers, column-major, and is asynchronous on the given
// stream.  The Fortran-77 ABI symbols (fortran_abi.cu), the LAPACKE entry points (lapacke_api.cu)
// and the device-pointer C API (capi.cu) sit on top of these.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <mutex>

#define LB_CUDA_CHECK(expr)                                                                       \
    do {                                                                                          \
        cudaError_t _e = (expr);                                                                  \
        if (_e != cudaSuccess) {                                                                  \
            fprintf(stderr, "lapack_b200: CUDA error %s at %s:%d: %s\n", cudaGetErrorName(_e),    \
                    __FILE__, __LINE__, cudaGetErrorString(_e));                                  \
            lb::record_cuda_error(_e);                                                            \
        }                                                                                         \
    } while (0)

namespace lb {

typedef long long i64;

void record_cuda_error(cudaError_t e);
int last_cuda_error();        // 0 if none since last clear
int take_cuda_error();        // error raised since the last take (0 if none); consumes it
int pending_cuda_error();     // same without consuming
void clear_cuda_error();

inline __host__ __device__ i64 idx2(i64 i, i64 j, i64 ld) { return i + j * ld; }
inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// ------------------------------------------------------------------ Level-3 BLAS (device)
// C := alpha*op(A)*op(B) + beta*C.  tri: 0 = full C, 1 = only the lower triangle of C is
// read/written (row >= col), 2 = only the upper triangle (used for DSYRK; C must be square).
void gemm(cudaStream_t s, char transa, char transb, int m, int n, int k, double alpha, const double* A,
          i64 lda, const double* B, i64 ldb, double beta, double* C, i64 ldc, int tri = 0);
void syrk(cudaStream_t s, char uplo, char trans, int n, int k, double alpha, const double* A, i64 lda,
          double beta, double* C, i64 ldc);
// op(A)*X = alpha*B or X*op(A) = alpha*B, X overwrites B.
void trsm(cudaStream_t s, char side, char uplo, char trans, char diag, int m, int n, double alpha,
          const double* A, i64 lda, double* B, i64 ldb);
// Drivers' large solves only (thread-local switch): 32 x 32 leaves as in-place DMMA products with the inverted diagonal blocks.
void trsm_set_inverse_leaves(int on);
int trsm_inverse_enabled();                 // library-wide knob (lb200_set_trsm_inverse), default 1
// B := alpha*op(A)*B or alpha*B*op(A), A triangular.  Needs a scratch copy of B (from the pool).
void trmm(cudaStream_t s, char side, char uplo, char trans, char diag, int m, int n, double alpha,
          const double* A, i64 lda, double* B, i64 ldb);

// ------------------------------------------------------------------ aux kernels (device)
void laswp(cudaStream_t s, int n, double* A, i64 lda, int k1, int k2, const int* ipiv, int incx);
void lacpy(cudaStream_t s, char uplo, int m, int n, const double* A, i64 lda, double* B, i64 ldb);
void laset(cudaStream_t s, char uplo, int m, int n, double alpha, double beta, double* A, i64 lda);
void transpose(cudaStream_t s, int m, int n, const double* A, i64 lda, double* B, i64 ldb);  // B = A^T
void larnv_fill(cudaStream_t s, int idist, const int iseed[4], i64 offset, i64 count, double* x);
void larnv_matrix(cudaStream_t s, const int iseed[4], i64 stream_offset, int m, int n, double* A, i64 lda);
void make_spd(cudaStream_t s, int n, double* A, i64 lda, double shift);  // A := (A+A^T)/2 + shift*I
void larnv_submatrix(cudaStream_t s, const int iseed[4], i64 stream_offset, i64 stream_ld, int m, int n, double* A, i64 lda);
void laswp_compose(cudaStream_t s, int np, const int* ipiv_rel, int* src_top, int* inv_top);
void gather_rows(cudaStream_t s, int nidx, const int* idx, const double* A, i64 lda, int ncols, double* W, i64 ldw);
void scatter_rows(cudaStream_t s, int nidx, const int* idx, const double* W, i64 ldw, int ncols, double* A, i64 lda);
void iadd(cudaStream_t s, int n, int* x, int v);
void info_max_offset(cudaStream_t s, int* info, const int* iinfo, int offset);  // LU: first nonzero wins

// ------------------------------------------------------------------ factorizations (device)
// All take a device int* info (single int, must be zeroed or is zeroed inside as documented).
void getrf(cudaStream_t s, int m, int n, double* A, i64 lda, int* ipiv, int* info);
void getrf2(cudaStream_t s, int m, int n, double* A, i64 lda, int* ipiv, int* info);
void getrs(cudaStream_t s, char trans, int n, int nrhs, const double* A, i64 lda, const int* ipiv,
           double* B, i64 ldb);
void getri(cudaStream_t s, int n, double* A, i64 lda, const int* ipiv, int* info);   // info: device word, set inside
void potrf(cudaStream_t s, char uplo, int n, double* A, i64 lda, int* info);
void potrf2(cudaStream_t s, char uplo, int n, double* A, i64 lda, int* info);
void potrs(cudaStream_t s, char uplo, int n, int nrhs, const double* A, i64 lda, double* B, i64 ldb);
void geqrf(cudaStream_t s, int m, int n, double* A, i64 lda, double* tau);
void geqr2(cudaStream_t s, int m, int n, double* A, i64 lda, double* tau);
void larft(cudaStream_t s, int n, int k, const double* V, i64 ldv, const double* tau, double* T, i64 ldt);
void larfb(cudaStream_t s, char side, char trans, int m, int n, int k, const double* V, i64 ldv,
           const double* T, i64 ldt, double* C, i64 ldc);
// all four DIRECT ('F'/'B') x STOREV ('C'/'R') storage schemes of dlarft.f / dlarfb.f
void larft_general(cudaStream_t s, bool backward, bool rowwise, int n, int k, const double* V, i64 ldv, const double* tau, double* T,
                   i64 ldt);
void larfb_general(cudaStream_t s, char side, char trans, bool backward, bool rowwise, int m, int n, int k, const double* V, i64 ldv,
                   const double* T, i64 ldt, double* C, i64 ldc);
void geqrt(cudaStream_t s, int m, int n, int nb, double* A, i64 lda, double* T, i64 ldt);
void latsqr(cudaStream_t s, int m, int n, int mb, int nb, double* A, i64 lda, double* T, i64 ldt);
void gemqrt(cudaStream_t s, char side, char trans, int m, int n, int k, int nb, const double* V, i64 ldv, const double* T,
            i64 ldt, double* C, i64 ldc);
void gelqf(cudaStream_t s, int m, int n, double* A, i64 lda, double* tau);
void ormlq(cudaStream_t s, char side, char trans, int m, int n, int k, const double* A, i64 lda, const double* tau, double* C,
           i64 ldc);
double amax_abs(cudaStream_t s, int m, int n, const double* A, i64 lda);   // host value; synchronises the stream
void scale_matrix(cudaStream_t s, int m, int n, double alpha, double* B, i64 ldb);
void ormqr(cudaStream_t s, char side, char trans, int m, int n, int k, const double* A, i64 lda, const double* tau,
           double* C, i64 ldc);
void orgqr(cudaStream_t s, int m, int n, int k, double* A, i64 lda, const double* tau);
void getrf_batched_32(cudaStream_t s, i64 batch, double* A, int* ipiv, int* info);
void potrf_batched_32(cudaStream_t s, char uplo, i64 batch, double* A, int* info);

// ------------------------------------------------------------------ workspace pool
// Stream-ordered scratch (cudaMallocAsync on the library's pool); freed with ws_free on the same stream.
void* ws_alloc(cudaStream_t s, size_t bytes);
void ws_free(cudaStream_t s, void* p);

// Streaming download hook for host-resident callers (Cholesky, QR): as soon as a block column (or block row) of
// the factor is final, the factorization copies it to the caller's host matrix on a separate stream, so that the
// device->host transfer overlaps with the rest of the factorization instead of following it.
struct StreamOut {
    double* host = nullptr;      // caller's matrix (element (0,0)), pinned host memory
    i64 ldh = 0;
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev = nullptr;
    int done_cols = 0;           // potrf: columns [0, done_cols) (rows for UPLO='U') have been queued for download;
                                 // getrf: block rows [0, done_cols) of U (right of their diagonal blocks)
};
StreamOut*& stream_out();         // nullptr when no streaming download is requested
int getrf_block();                // outer block size of getrf (host path: which U block rows were streamed)
int potrf_block();                // outer block size of potrf (streamed host path needs it to divide its upload chunk)

// launch counter (bench.py's gpu_launches claim)
extern unsigned long long g_launches;
inline void count_launch(int n = 1) { g_launches += (unsigned long long)n; }

// side streams / events for look-ahead
struct Aux {
    cudaStream_t panel_stream = nullptr;   // high priority
    cudaStream_t update_stream = nullptr;
    cudaStream_t side_stream = nullptr;    // low priority: work that is off the critical path (left-of-panel interchanges)
    cudaStream_t prep_stream = nullptr;    // medium priority: memory-bound preparation (interchanges + U12 solve) of the next chunk
    cudaEvent_t ev[32];
    bool ready = false;
};
Aux& aux(int level = 0);    // level 1: second stream set for the outer level of the two-level LU driver

int num_sms();
std::recursive_mutex& driver_mutex();    // serialises the drivers that share lb::aux()'s streams and events
const int* kernel_guard();                 // see runtime.cu: device INFO word that disables queued Level-3 kernels once non-zero
void set_kernel_guard(const int* p);

}  // namespace lb
