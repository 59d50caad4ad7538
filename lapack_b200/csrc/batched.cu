// batched.cu -- batched factorization of many 32 x 32 matrices (config C5b): one matrix per warp.
//
// Semantics per matrix = SRC/dgetrf2.f (partial pivoting, IDAMAX first-index tie-break, reciprocal
// scaling unless |pivot| < SFMIN, INFO = first zero pivot) and SRC/dpotrf2.f.  Layout: matrices are
// contiguous, column-major, stride 1024 doubles; IPIV is 32 ints per matrix; INFO one int per matrix.
//
// Each lane holds one matrix row in registers (32 doubles).  LU uses *implicit* pivoting: rows are never
// exchanged between lanes; every lane tracks the position its row would have after LAPACK's interchanges,
// the arg-max is tie-broken on that position (which is what IDAMAX sees), and rows are written back to
// their final positions.  HBM-bound: 8192 B read + 8192 B written + 128 B IPIV per matrix.
#include "lb_internal.h"
#include <cfloat>
#include <math_constants.h>

namespace lb {

constexpr int BW = 32;

__global__ void __launch_bounds__(256) getrf_batched32_kernel(i64 batch, double* __restrict__ A, int* __restrict__ ipiv,
                                                              int* __restrict__ info) {
    const int lane = threadIdx.x & 31;
    const i64 id = (i64)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (id >= batch) return;
    double* M = A + id * (BW * BW);
    double a[BW];
#pragma unroll
    for (int c = 0; c < BW; ++c) a[c] = M[lane + BW * c];
    int mypos = lane;        // current position of my row under LAPACK's explicit interchanges
    bool done = false;       // my row has already been used as a pivot row
    int myipiv = 0, minfo = 0;
#pragma unroll
    for (int c = 0; c < BW; ++c) {
        double key = -1.0;
        if (!done) {
            key = fabs(a[c]);
            if (key != key) key = (mypos == c) ? CUDART_INF : -1.0;    // idamax.f:103: NaN only wins in first place
        }
        int pos = done ? 0x7fffffff : mypos;
        int who = lane;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            double ok = __shfl_xor_sync(0xffffffffu, key, off);
            int op = __shfl_xor_sync(0xffffffffu, pos, off);
            int ow = __shfl_xor_sync(0xffffffffu, who, off);
            if (ok > key || (ok == key && op < pos)) { key = ok; pos = op; who = ow; }
        }
        // `who` = lane holding the pivot row, `pos` = its current position
        if (lane == c) myipiv = pos + 1;
        const double pivot = __shfl_sync(0xffffffffu, a[c], who);
        // interchange bookkeeping: the row sitting at position c moves to `pos`
        if (!done && mypos == c && lane != who) mypos = pos;
        if (lane == who) { mypos = c; done = true; }
        if (pivot == 0.0) {
            if (minfo == 0) minfo = c + 1;                              // dgetrf2.f:212-214
        } else {
            double l = 0.0;
            if (!done) {
                if (fabs(pivot) >= DBL_MIN) l = a[c] * (1.0 / pivot);   // dgetrf2.f:204-205
                else l = a[c] / pivot;                                  // dgetrf2.f:207-209
                a[c] = l;
            }
#pragma unroll
            for (int q = c + 1; q < BW; ++q) {
                double pq = __shfl_sync(0xffffffffu, a[q], who);
                if (!done) a[q] = fma(-l, pq, a[q]);
            }
        }
    }
#pragma unroll
    for (int c = 0; c < BW; ++c) M[mypos + BW * c] = a[c];
    ipiv[id * BW + lane] = myipiv;
    if (lane == 0) info[id] = minfo;
}

void getrf_batched_32(cudaStream_t s, i64 batch, double* A, int* ipiv, int* info) {
    if (batch <= 0) return;
    const int wpb = 8;
    getrf_batched32_kernel<<<(unsigned)((batch + wpb - 1) / wpb), wpb * 32, 0, s>>>(batch, A, ipiv, info);
    count_launch();
    LB_CUDA_CHECK(cudaGetLastError());
}

__global__ void __launch_bounds__(256) potrf_batched32_kernel(i64 batch, double* __restrict__ A, bool upper,
                                                              int* __restrict__ info) {
    const int lane = threadIdx.x & 31;
    const i64 id = (i64)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (id >= batch) return;
    double* M = A + id * (BW * BW);
    // lane i holds row i of L (for UPLO='U': L = U^T, i.e. column i of the stored matrix)
    double a[BW];
#pragma unroll
    for (int k = 0; k < BW; ++k) a[k] = (k <= lane) ? (upper ? M[k + BW * lane] : M[lane + BW * k]) : 0.0;
    int fail = 0;
#pragma unroll
    for (int k = 0; k < BW; ++k) {
        const double d = __shfl_sync(0xffffffffu, a[k], k);
        if (fail == 0 && (d <= 0.0 || d != d)) fail = k + 1;            // dpotrf2.f:169-172 (warp-uniform)
        if (fail == 0) {
            const double sq = sqrt(d);
            if (lane == k) a[k] = sq;
            else if (lane > k) a[k] = a[k] / sq;
#pragma unroll
            for (int j = k + 1; j < BW; ++j) {
                const double ljk = __shfl_sync(0xffffffffu, a[k], j);
                if (lane >= j) a[j] = fma(-a[k], ljk, a[j]);
            }
        }
    }
#pragma unroll
    for (int k = 0; k < BW; ++k)
        if (k <= lane) { if (upper) M[k + BW * lane] = a[k]; else M[lane + BW * k] = a[k]; }
    if (lane == 0) info[id] = fail;
}

void potrf_batched_32(cudaStream_t s, char uplo, i64 batch, double* A, int* info) {
    if (batch <= 0) return;
    const int wpb = 8;
    potrf_batched32_kernel<<<(unsigned)((batch + wpb - 1) / wpb), wpb * 32, 0, s>>>(batch, A, uplo == 'U' || uplo == 'u', info);
    count_launch();
    LB_CUDA_CHECK(cudaGetLastError());
}

}  // namespace lb
