// trsm_f64.cu -- triangular solve / multiply with many right-hand sides (DTRSM, DTRMM) for sm_100a.
//
// Replaces BLAS/SRC/dtrsm.f:257-405 (all 16 SIDE/UPLO/TRANS/DIAG combinations) and BLAS/SRC/dtrmm.f.
// Call sites on the hot path: SRC/dgetrf.f:204 and dgetrf2.f:240 (L,L,N,U), dpotrf.f:229 (R,L,T,N) and
// :199 (L,U,T,N), dgetrs.f:191-217, dpotrs.f:174-195.
//
// Structure: the triangle is split recursively; the off-diagonal part of every split is a DMMA GEMM
// (gemm_f64.cu), so for a 512-wide triangle 7/8 of the flops run on the FP64 tensor pipe.  The leaves
// (<= 32 x 32 triangle) are solved by substitution with one thread per right-hand side, the triangle
// held in shared memory (broadcast reads) and the right-hand side held in registers; divisions by the
// diagonal are kept as divisions like the reference (dtrsm.f:282,294,...).
#include "lb_internal.h"

namespace lb {

constexpr int TB = 32;   // leaf triangle size

// T(i,k) = op(A)(i,k) for the leaf, as a dense TB x TB array in shared memory (only the triangle is read
// from global memory; the rest is zero).
__device__ __forceinline__ void load_leaf_tri(double (*T)[TB + 1], const double* __restrict__ A, i64 lda, int nb,
                                              bool upper, bool trans) {
    for (int idx = threadIdx.x; idx < TB * TB; idx += blockDim.x) {
        int i = idx % TB, k = idx / TB;   // stored element A(i,k)
        double v = 0.0;
        if (i < nb && k < nb && (upper ? i <= k : i >= k)) v = A[i + (i64)k * lda];
        if (trans) T[k][i] = v; else T[i][k] = v;
    }
}

// Left side: solve T * X = B, T = op(A) (nb x nb), B is nb x n; one thread per column of B.
//   LOWER=true  : T lower triangular -> forward substitution   (dtrsm.f:290-300 / 307-316)
//   LOWER=false : T upper triangular -> backward substitution  (dtrsm.f:278-288 / 318-327)
template <bool LOWER>
__global__ void __launch_bounds__(128) trsm_left_leaf_kernel(int nb, int n, const double* __restrict__ A, i64 lda,
                                                             bool a_upper, bool trans, bool unit, double* __restrict__ B,
                                                             i64 ldb) {
    __shared__ double T[TB][TB + 1];
    load_leaf_tri(T, A, lda, nb, a_upper, trans);
    __syncthreads();
    int col = blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= n) return;
    double* b = B + (i64)col * ldb;
    double x[TB];
#pragma unroll
    for (int i = 0; i < TB; ++i) x[i] = (i < nb) ? b[i] : 0.0;
    if (LOWER) {
#pragma unroll
        for (int k = 0; k < TB; ++k) {
            if (k < nb) {
                if (!unit) x[k] = x[k] / T[k][k];
#pragma unroll
                for (int i = k + 1; i < TB; ++i) x[i] = x[i] - x[k] * T[i][k];
            }
        }
    } else {
#pragma unroll
        for (int k = TB - 1; k >= 0; --k) {
            if (k < nb) {
                if (!unit) x[k] = x[k] / T[k][k];
#pragma unroll
                for (int i = 0; i < k; ++i) x[i] = x[i] - x[k] * T[i][k];
            }
        }
    }
#pragma unroll
    for (int i = 0; i < TB; ++i)
        if (i < nb) b[i] = x[i];
}

// Right side: solve X * T = B, T = op(A) (nb x nb), B is m x nb; one thread per row of B (coalesced).
//   UPPER_T=true  : T upper triangular -> columns left to right  (dtrsm.f:337-354 / 403-424)
//   UPPER_T=false : T lower triangular -> columns right to left  (dtrsm.f:356-373 / 380-401)
template <bool UPPER_T>
__global__ void __launch_bounds__(128) trsm_right_leaf_kernel(int m, int nb, const double* __restrict__ A, i64 lda,
                                                              bool a_upper, bool trans, bool unit, double* __restrict__ B,
                                                              i64 ldb) {
    __shared__ double T[TB][TB + 1];
    load_leaf_tri(T, A, lda, nb, a_upper, trans);
    __syncthreads();
    int row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= m) return;
    double* b = B + row;
    double x[TB];
#pragma unroll
    for (int j = 0; j < TB; ++j) x[j] = (j < nb) ? b[(i64)j * ldb] : 0.0;
    if (UPPER_T) {
#pragma unroll
        for (int j = 0; j < TB; ++j) {
            if (j < nb) {
#pragma unroll
                for (int k = 0; k < j; ++k) x[j] = x[j] - T[k][j] * x[k];
                if (!unit) x[j] = x[j] / T[j][j];
            }
        }
    } else {
#pragma unroll
        for (int j = TB - 1; j >= 0; --j) {
            if (j < nb) {
#pragma unroll
                for (int k = j + 1; k < TB; ++k) x[j] = x[j] - T[k][j] * x[k];
                if (!unit) x[j] = x[j] / T[j][j];
            }
        }
    }
#pragma unroll
    for (int j = 0; j < TB; ++j)
        if (j < nb) b[(i64)j * ldb] = x[j];
}

__global__ void scale_full_kernel(int m, int n, double alpha, double* B, i64 ldb) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    for (int j = blockIdx.y; j < n; j += gridDim.y) {
        double* b = B + i + (i64)j * ldb;
        *b = (alpha == 0.0) ? 0.0 : alpha * (*b);
    }
}
static void scale_full(cudaStream_t s, int m, int n, double alpha, double* B, i64 ldb) {
    dim3 grid(ceil_div(m, 256), (unsigned)min(n, 8192));
    scale_full_kernel<<<grid, 256, 0, s>>>(m, n, alpha, B, ldb);
    count_launch();
}

// ----------------------------------------------------------------------------------------------
// Fused small left-lower solve (no transpose): T*X = B with T (m x m, m <= 256) lower triangular, B m x n.
// Used inside the LU panel recursion (dgetrf2.f:240) where the recursive version would cost ~2 launches per
// 32 rows on the critical path.  One CTA owns TS_COLS columns of B (kept in shared memory); per 32-row
// block: (1) warp-per-column forward substitution with shuffles, (2) all threads update the rows below,
// streaming the 32-column strip of T from L2.
constexpr int TS_COLS = 8;
constexpr int TS_MAXM = 256;
__global__ void __launch_bounds__(256) trsm_left_lower_small_kernel(int m, int n, const double* __restrict__ A, i64 lda,
                                                                    bool unit, double* __restrict__ B, i64 ldb) {
    __shared__ double sB[TS_COLS][TS_MAXM + 2];
    __shared__ double sT[32][33];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int c0 = blockIdx.x * TS_COLS;
    const int nc = min(TS_COLS, n - c0);
    for (int q = tid; q < nc * m; q += 256) {
        int c = q / m, r = q - c * m;
        sB[c][r] = B[r + (i64)(c0 + c) * ldb];
    }
    __syncthreads();
    for (int kb = 0; kb < m; kb += 32) {
        const int bs = min(32, m - kb);
        // diagonal block -> shared (only the lower triangle is read)
        for (int q = tid; q < 32 * 32; q += 256) {
            int i = q & 31, k = q >> 5;
            sT[i][k] = (i < bs && k < bs && i >= k) ? A[(kb + i) + (i64)(kb + k) * lda] : 0.0;
        }
        __syncthreads();
        // (1) forward substitution: warp w solves column w of the slab, lane = row inside the block
        if (warp < nc) {
            double x = (lane < bs) ? sB[warp][kb + lane] : 0.0;
            for (int k = 0; k < bs; ++k) {
                if (!unit && lane == k) x = x / sT[k][k];
                double xk = __shfl_sync(0xffffffffu, x, k);
                if (lane > k) x = x - xk * sT[lane][k];
            }
            if (lane < bs) sB[warp][kb + lane] = x;
        }
        __syncthreads();
        // (2) rows below the block: B(r,:) -= T(r, kb:kb+bs) * X(kb:kb+bs, :)
        for (int r = kb + 32 + tid; r < m; r += 256) {
            double acc[TS_COLS];
#pragma unroll
            for (int c = 0; c < TS_COLS; ++c) acc[c] = 0.0;
            const double* trow = A + r + (i64)kb * lda;
#pragma unroll 8
            for (int k = 0; k < 32; ++k) {
                if (k < bs) {
                    double t = trow[(i64)k * lda];
#pragma unroll
                    for (int c = 0; c < TS_COLS; ++c) acc[c] = fma(t, sB[c][kb + k], acc[c]);
                }
            }
#pragma unroll
            for (int c = 0; c < TS_COLS; ++c) sB[c][r] -= acc[c];
        }
        __syncthreads();
    }
    for (int q = tid; q < nc * m; q += 256) {
        int c = q / m, r = q - c * m;
        B[r + (i64)(c0 + c) * ldb] = sB[c][r];
    }
}

static inline bool is(char c, char u) { return c == u || c == (char)(u + 32); }

// split point: largest multiple of TB (power-of-two times TB preferred) not exceeding half, at least TB
static int split_point(int n) {
    int h = TB;
    while (h * 2 < n) h *= 2;
    return h;   // TB <= h < n, h = TB * 2^j
}

static void trsm_left_rec(cudaStream_t s, bool upper, bool trans, bool unit, int m, int n, const double* A, i64 lda,
                          double* B, i64 ldb) {
    const bool eff_lower = (upper == trans);   // (L,N) or (U,T)
    if (m <= TB) {
        int threads = 128;
        if (eff_lower) trsm_left_leaf_kernel<true><<<ceil_div(n, threads), threads, 0, s>>>(m, n, A, lda, upper, trans, unit, B, ldb);
        else trsm_left_leaf_kernel<false><<<ceil_div(n, threads), threads, 0, s>>>(m, n, A, lda, upper, trans, unit, B, ldb);
        count_launch();
        return;
    }
    int m1 = split_point(m), m2 = m - m1;
    const double* A11 = A;
    const double* A22 = A + m1 + (i64)m1 * lda;
    const double* A21 = A + m1;                    // stored (m2 x m1) block below the diagonal
    const double* A12 = A + (i64)m1 * lda;         // stored (m1 x m2) block right of the diagonal
    double* B1 = B;
    double* B2 = B + m1;
    if (eff_lower) {
        trsm_left_rec(s, upper, trans, unit, m1, n, A11, lda, B1, ldb);
        if (!trans) gemm(s, 'N', 'N', m2, n, m1, -1.0, A21, lda, B1, ldb, 1.0, B2, ldb);
        else gemm(s, 'T', 'N', m2, n, m1, -1.0, A12, lda, B1, ldb, 1.0, B2, ldb);
        trsm_left_rec(s, upper, trans, unit, m2, n, A22, lda, B2, ldb);
    } else {
        trsm_left_rec(s, upper, trans, unit, m2, n, A22, lda, B2, ldb);
        if (!trans) gemm(s, 'N', 'N', m1, n, m2, -1.0, A12, lda, B2, ldb, 1.0, B1, ldb);
        else gemm(s, 'T', 'N', m1, n, m2, -1.0, A21, lda, B2, ldb, 1.0, B1, ldb);
        trsm_left_rec(s, upper, trans, unit, m1, n, A11, lda, B1, ldb);
    }
}

static void trsm_right_rec(cudaStream_t s, bool upper, bool trans, bool unit, int m, int n, const double* A, i64 lda,
                           double* B, i64 ldb) {
    const bool eff_upper = (upper != trans);   // (U,N) or (L,T)
    if (n <= TB) {
        int threads = 128;
        if (eff_upper) trsm_right_leaf_kernel<true><<<ceil_div(m, threads), threads, 0, s>>>(m, n, A, lda, upper, trans, unit, B, ldb);
        else trsm_right_leaf_kernel<false><<<ceil_div(m, threads), threads, 0, s>>>(m, n, A, lda, upper, trans, unit, B, ldb);
        count_launch();
        return;
    }
    int n1 = split_point(n), n2 = n - n1;
    const double* A11 = A;
    const double* A22 = A + n1 + (i64)n1 * lda;
    const double* A21 = A + n1;
    const double* A12 = A + (i64)n1 * lda;
    double* B1 = B;
    double* B2 = B + (i64)n1 * ldb;
    if (eff_upper) {
        trsm_right_rec(s, upper, trans, unit, m, n1, A11, lda, B1, ldb);
        if (!trans) gemm(s, 'N', 'N', m, n2, n1, -1.0, B1, ldb, A12, lda, 1.0, B2, ldb);
        else gemm(s, 'N', 'T', m, n2, n1, -1.0, B1, ldb, A21, lda, 1.0, B2, ldb);
        trsm_right_rec(s, upper, trans, unit, m, n2, A22, lda, B2, ldb);
    } else {
        trsm_right_rec(s, upper, trans, unit, m, n2, A22, lda, B2, ldb);
        if (!trans) gemm(s, 'N', 'N', m, n1, n2, -1.0, B2, ldb, A21, lda, 1.0, B1, ldb);
        else gemm(s, 'N', 'T', m, n1, n2, -1.0, B2, ldb, A12, lda, 1.0, B1, ldb);
        trsm_right_rec(s, upper, trans, unit, m, n1, A11, lda, B1, ldb);
    }
}

void trsm(cudaStream_t s, char side, char uplo, char trans, char diag, int m, int n, double alpha, const double* A,
          i64 lda, double* B, i64 ldb) {
    if (m <= 0 || n <= 0) return;
    if (alpha != 1.0) {   // alpha == 0: B := 0 without touching A (dtrsm.f:261-268)
        scale_full(s, m, n, alpha, B, ldb);
        if (alpha == 0.0) return;
    }
    const bool left = is(side, 'L'), upper = is(uplo, 'U'), tr = !is(trans, 'N'), unit = is(diag, 'U');
    if (left && !upper && !tr && m <= TS_MAXM && m > TB && n <= 4096) {
        // latency-critical in-panel solve: one fused kernel instead of the recursion
        trsm_left_lower_small_kernel<<<ceil_div(n, TS_COLS), 256, 0, s>>>(m, n, A, lda, unit, B, ldb);
        count_launch();
    } else if (left) trsm_left_rec(s, upper, tr, unit, m, n, A, lda, B, ldb);
    else trsm_right_rec(s, upper, tr, unit, m, n, A, lda, B, ldb);
    LB_CUDA_CHECK(cudaGetLastError());
}

// ----------------------------------------------------------------------------------------------
// DTRMM: B := alpha*op(A)*B or alpha*B*op(A).  The triangle is expanded into a dense scratch matrix
// (explicit zeros, explicit unit diagonal) and the product runs as one DMMA GEMM into a scratch copy.
__global__ void expand_tri_kernel(int n, const double* __restrict__ A, i64 lda, bool upper, bool unit,
                                  double* __restrict__ T, i64 ldt) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    for (int j = blockIdx.y; j < n; j += gridDim.y) {
        double v = 0.0;
        if (i == j) v = unit ? 1.0 : A[i + (i64)j * lda];
        else if (upper ? i < j : i > j) v = A[i + (i64)j * lda];
        T[i + (i64)j * ldt] = v;
    }
}
void expand_tri(cudaStream_t s, int n, const double* A, i64 lda, bool upper, bool unit, double* T, i64 ldt) {
    dim3 grid(ceil_div(n, 128), (unsigned)min(n, 8192));
    expand_tri_kernel<<<grid, 128, 0, s>>>(n, A, lda, upper, unit, T, ldt);
    count_launch();
}

void trmm(cudaStream_t s, char side, char uplo, char trans, char diag, int m, int n, double alpha, const double* A,
          i64 lda, double* B, i64 ldb) {
    if (m <= 0 || n <= 0) return;
    if (alpha == 0.0) { scale_full(s, m, n, 0.0, B, ldb); return; }
    const bool left = is(side, 'L'), upper = is(uplo, 'U'), tr = !is(trans, 'N'), unit = is(diag, 'U');
    const int na = left ? m : n;
    i64 ldt = (na + 1) & ~1;
    i64 ldw = (m + 1) & ~1;
    double* T = (double*)ws_alloc(s, sizeof(double) * ldt * na);
    double* W = (double*)ws_alloc(s, sizeof(double) * ldw * n);
    expand_tri(s, na, A, lda, upper, unit, T, ldt);
    lacpy(s, 'A', m, n, B, ldb, W, ldw);
    if (left) gemm(s, tr ? 'T' : 'N', 'N', m, n, m, alpha, T, ldt, W, ldw, 0.0, B, ldb);
    else gemm(s, 'N', tr ? 'T' : 'N', m, n, n, alpha, W, ldw, T, ldt, 0.0, B, ldb);
    ws_free(s, T);
    ws_free(s, W);
}

}  // namespace lb
