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
#include <cfloat>

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
                                                             i64 ldb, const int* guard) {
    __shared__ double T[TB][TB + 1];
    if (guard && *guard != 0) return;
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
                                                              i64 ldb, const int* guard) {
    __shared__ double T[TB][TB + 1];
    if (guard && *guard != 0) return;
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
            double tv[32];                                   // the whole 32-entry strip in flight before the first FMA
#pragma unroll
            for (int k = 0; k < 32; ++k) tv[k] = (k < bs) ? __ldg(trow + (i64)k * lda) : 0.0;
#pragma unroll
            for (int k = 0; k < 32; ++k) {
                if (k < bs) {                                // rows beyond the block are uninitialised shared memory
#pragma unroll
                    for (int c = 0; c < TS_COLS; ++c) acc[c] = fma(tv[k], sB[c][kb + k], acc[c]);
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

void scale_matrix(cudaStream_t s, int m, int n, double alpha, double* B, i64 ldb) { scale_full(s, m, n, alpha, B, ldb); }

static inline bool is(char c, char u) { return c == u || c == (char)(u + 32); }

// ----------------------------------------------------------------------------------------------
// Inverted diagonal blocks (opt-in, used by DPOTRF for its panel solve; measured neutral for U12 = inv(L11) A12 of DGETRF, where
// it is therefore not used).  The 32 x 32 substitution leaves above are latency-bound CUDA-core kernels that take SM slots from the concurrent
// trailing GEMM; with the inverse of every 32 x 32 diagonal block formed once per triangle (one small launch), a leaf becomes the
// in-place DMMA product B := inv(T_kk) B (or B inv(T_kk)), K = 32, which only streams B.  The blocks are diagonal blocks of a Cholesky factor, so inv(T_kk) is benign (factor difference against the
// substitution leaves at n = 32768: 6e-18 absolute); the public DTRSM keeps
// the substitution leaves (dtrsm.f semantics entry by entry).
struct InvBlocks { const double* base = nullptr; int blk0 = 0; };       // 32 x 32 dense blocks, ld 32; blk0 = block index of A(0,0)
static thread_local int g_trsm_inverse_leaves = 0;
static int g_trsm_inverse_enabled = 1;
void trsm_set_inverse_leaves(int on) { g_trsm_inverse_leaves = on && g_trsm_inverse_enabled; }
int trsm_inverse_enabled() { return g_trsm_inverse_enabled; }
void trsm_set_inverse_enabled(int on) { g_trsm_inverse_enabled = on; }

// one CTA (32 threads) per diagonal block: thread j solves T x = e_j by substitution; partial last block padded with the identity
__global__ void __launch_bounds__(32) trtri32_blocks_kernel(int n, const double* __restrict__ A, i64 lda, bool upper, bool unit,
                                                            double* __restrict__ Tinv, const int* guard) {
    __shared__ double T[TB][TB + 1];
    if (guard && *guard != 0) return;
    const int b = blockIdx.x, r0 = b * TB, nb = min(TB, n - r0), j = threadIdx.x;
    for (int i = 0; i < TB; ++i) {
        double v = (i == j) ? 1.0 : 0.0;
        if (i < nb && j < nb && (upper ? i <= j : i >= j) && !(unit && i == j)) v = A[(r0 + i) + (i64)(r0 + j) * lda];
        T[i][j] = v;
    }
    __syncwarp();
    double x[TB];
#pragma unroll
    for (int i = 0; i < TB; ++i) x[i] = (i == j) ? 1.0 : 0.0;
    if (!upper) {
#pragma unroll
        for (int k = 0; k < TB; ++k) {
            x[k] = x[k] / T[k][k];
#pragma unroll
            for (int i = k + 1; i < TB; ++i) x[i] = fma(-x[k], T[i][k], x[i]);
        }
    } else {
#pragma unroll
        for (int k = TB - 1; k >= 0; --k) {
            x[k] = x[k] / T[k][k];
#pragma unroll
            for (int i = 0; i < k; ++i) x[i] = fma(-x[k], T[i][k], x[i]);
        }
    }
    double* out = Tinv + (size_t)b * TB * TB + (size_t)j * TB;
#pragma unroll
    for (int i = 0; i < TB; ++i) out[i] = x[i];
}

// split point: largest multiple of TB (power-of-two times TB preferred) not exceeding half, at least TB
static int split_point(int n) {
    int h = TB;
    while (h * 2 < n) h *= 2;
    return h;   // TB <= h < n, h = TB * 2^j
}

static void trsm_left_rec(cudaStream_t s, bool upper, bool trans, bool unit, int m, int n, const double* A, i64 lda,
                          double* B, i64 ldb, InvBlocks iv = InvBlocks()) {
    const bool eff_lower = (upper == trans);   // (L,N) or (U,T)
    if (m <= TB && iv.base) {
        // B := op(inv(T_kk)) B in place: a CTA's 64-column output tile reads exactly its own columns of B, all of them before it writes
        gemm(s, trans ? 'T' : 'N', 'N', m, n, m, 1.0, iv.base + (size_t)iv.blk0 * TB * TB, TB, B, ldb, 0.0, B, ldb);
        return;
    }
    if (m <= TB) {
        int threads = 128;
        if (eff_lower) trsm_left_leaf_kernel<true><<<ceil_div(n, threads), threads, 0, s>>>(m, n, A, lda, upper, trans, unit, B, ldb, kernel_guard());
        else trsm_left_leaf_kernel<false><<<ceil_div(n, threads), threads, 0, s>>>(m, n, A, lda, upper, trans, unit, B, ldb, kernel_guard());
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
    InvBlocks iv2 = iv;
    iv2.blk0 = iv.blk0 + m1 / TB;
    if (eff_lower) {
        trsm_left_rec(s, upper, trans, unit, m1, n, A11, lda, B1, ldb, iv);
        if (!trans) gemm(s, 'N', 'N', m2, n, m1, -1.0, A21, lda, B1, ldb, 1.0, B2, ldb);
        else gemm(s, 'T', 'N', m2, n, m1, -1.0, A12, lda, B1, ldb, 1.0, B2, ldb);
        trsm_left_rec(s, upper, trans, unit, m2, n, A22, lda, B2, ldb, iv2);
    } else {
        trsm_left_rec(s, upper, trans, unit, m2, n, A22, lda, B2, ldb, iv2);
        if (!trans) gemm(s, 'N', 'N', m1, n, m2, -1.0, A12, lda, B2, ldb, 1.0, B1, ldb);
        else gemm(s, 'T', 'N', m1, n, m2, -1.0, A21, lda, B2, ldb, 1.0, B1, ldb);
        trsm_left_rec(s, upper, trans, unit, m1, n, A11, lda, B1, ldb, iv);
    }
}

static void trsm_right_rec(cudaStream_t s, bool upper, bool trans, bool unit, int m, int n, const double* A, i64 lda,
                           double* B, i64 ldb, InvBlocks iv = InvBlocks()) {
    const bool eff_upper = (upper != trans);   // (U,N) or (L,T)
    if (n <= TB && iv.base) {
        // B := B op(inv(T_kk)) in place: a CTA's 64-row output tile reads exactly its own rows of B, all of them before it writes
        gemm(s, 'N', trans ? 'T' : 'N', m, n, n, 1.0, B, ldb, iv.base + (size_t)iv.blk0 * TB * TB, TB, 0.0, B, ldb);
        return;
    }
    if (n <= TB) {
        int threads = 128;
        if (eff_upper) trsm_right_leaf_kernel<true><<<ceil_div(m, threads), threads, 0, s>>>(m, n, A, lda, upper, trans, unit, B, ldb, kernel_guard());
        else trsm_right_leaf_kernel<false><<<ceil_div(m, threads), threads, 0, s>>>(m, n, A, lda, upper, trans, unit, B, ldb, kernel_guard());
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
    InvBlocks iv2 = iv;
    iv2.blk0 = iv.blk0 + n1 / TB;
    if (eff_upper) {
        trsm_right_rec(s, upper, trans, unit, m, n1, A11, lda, B1, ldb, iv);
        if (!trans) gemm(s, 'N', 'N', m, n2, n1, -1.0, B1, ldb, A12, lda, 1.0, B2, ldb);
        else gemm(s, 'N', 'T', m, n2, n1, -1.0, B1, ldb, A21, lda, 1.0, B2, ldb);
        trsm_right_rec(s, upper, trans, unit, m, n2, A22, lda, B2, ldb, iv2);
    } else {
        trsm_right_rec(s, upper, trans, unit, m, n2, A22, lda, B2, ldb, iv2);
        if (!trans) gemm(s, 'N', 'N', m, n1, n2, -1.0, B2, ldb, A21, lda, 1.0, B1, ldb);
        else gemm(s, 'N', 'T', m, n1, n2, -1.0, B2, ldb, A12, lda, 1.0, B1, ldb);
        trsm_right_rec(s, upper, trans, unit, m, n1, A11, lda, B1, ldb, iv);
    }
}

// ----------------------------------------------------------------------------------------------
// Few right-hand sides (DGETRS / DPOTRS with NRHS <= 8): the solve is memory-bound (the triangle is read once, 8n^2
// bytes) and latency-bound along the diagonal; the generic recursion would issue thousands of tiny launches and run
// its off-diagonal updates as 64-column GEMM tiles with one useful column.  Here
//   * the recursion stops at FR_LEAF = 128 rows; a leaf is ONE CTA that first pulls its whole diagonal block and its
//     right-hand sides into shared memory (all loads in flight at once), then walks the block in 32-row steps: warp 0
//     substitutes inside the 32 x 32 sub-block (lane = row, right-hand sides in registers), the other warps update
//     the rest of the leaf;
//   * the off-diagonal updates are streaming GEMV kernels (coalesced, all SMs, fixed summation order).
constexpr int FR_LEAF = 128, FR_MAXRHS = 8;
void trsv_stream(cudaStream_t s, bool upper, bool trans, bool unit, int n, int nrhs, const double* A, i64 lda, double* B,
                 i64 ldb);                      // trsv_stream.cu: one persistent kernel per solve
static int g_fewrhs_mode = 1;                   // 1 = persistent streaming kernel (default), 0 = leaf/GEMV recursion
void trsm_set_fewrhs_mode(int mode) { g_fewrhs_mode = mode; }

template <bool eff_lower>     // (L,N) / (U,T): forward substitution; otherwise backward
__global__ void __launch_bounds__(256) trsm_left_fewrhs_kernel(int m, int nrhs, const double* __restrict__ A, i64 lda,
                                                               bool trans, bool unit, double* __restrict__ B, i64 ldb) {
    extern __shared__ double fr_smem[];
    double (*S)[FR_LEAF + 1] = reinterpret_cast<double (*)[FR_LEAF + 1]>(fr_smem);          // S[i][j] = T(i,j), T = op(A)
    double (*bs)[FR_MAXRHS] = reinterpret_cast<double (*)[FR_MAXRHS]>(fr_smem + FR_LEAF * (FR_LEAF + 1));
    double* rinv = fr_smem + FR_LEAF * (FR_LEAF + 1) + FR_LEAF * FR_MAXRHS;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // ---- one shot: the m x m block and the right-hand sides
    {
        // thread = (row of A, column parity): every warp-level load is 32 consecutive rows of one column of A; sixteen
        // loads are in flight per thread before the first store
        const int ar = tid & (FR_LEAF - 1), ac0 = tid >> 7;
        if (ar < m) {
            const double* src = A + ar;
#pragma unroll 1
            for (int c0 = ac0; c0 < m; c0 += 32) {
                double v[16];
#pragma unroll
                for (int u = 0; u < 16; ++u) v[u] = (c0 + 2 * u < m) ? __ldg(src + (i64)(c0 + 2 * u) * lda) : 0.0;
#pragma unroll
                for (int u = 0; u < 16; ++u) {
                    const int ac = c0 + 2 * u;
                    if (ac < m) { if (!trans) S[ar][ac] = v[u]; else S[ac][ar] = v[u]; }      // T(i,j) = A(j,i)
                }
            }
        }
        for (int idx = tid; idx < FR_LEAF * FR_MAXRHS; idx += 256) {
            const int i = idx & (FR_LEAF - 1), r = idx >> 7;
            if (i < m && r < nrhs) bs[i][r] = B[i + (i64)r * ldb];
        }
    }
    __syncthreads();
    if (tid < m) rinv[tid] = 1.0 / S[tid][tid];
    __syncthreads();
    const int nsb = (m + 31) / 32;
    for (int sidx = 0; sidx < nsb; ++sidx) {
        const int sb = eff_lower ? sidx : nsb - 1 - sidx;
        const int r0 = sb * 32;
        const int rows = min(32, m - r0);
        if (warp == 0) {
            double x[FR_MAXRHS];
#pragma unroll
            for (int r = 0; r < FR_MAXRHS; ++r) x[r] = (r < nrhs && lane < rows) ? bs[r0 + lane][r] : 0.0;
            const int li = min(lane, rows - 1);
#pragma unroll 4
            for (int jj = 0; jj < rows; ++jj) {
                const int j = eff_lower ? jj : rows - 1 - jj;
                const double tij = S[r0 + li][r0 + j];
                const double d = S[r0 + j][r0 + j], ri = rinv[r0 + j];
                const bool use_rcp = fabs(d) >= DBL_MIN;
                const bool below = eff_lower ? (lane > j) : (lane < j);
#pragma unroll
                for (int r = 0; r < FR_MAXRHS; ++r) {
                    if (r < nrhs) {
                        double xj = __shfl_sync(0xffffffffu, x[r], j);
                        // dtrsm.f: B(k,j) = B(k,j)/A(k,k).  The n divisions of a solve form one dependent chain, so the
                        // reciprocal (computed off the chain) is used unless it could overflow -- DGETRF2's own rule
                        // (dgetrf2.f:204-210); at most one extra rounding per entry
                        if (!unit) xj = use_rcp ? xj * ri : xj / d;
                        if (lane == j) x[r] = xj;
                        else if (below) x[r] = x[r] - xj * tij;       // dtrsm.f: B(i,j) = B(i,j) - B(k,j)*A(i,k)
                    }
                }
            }
#pragma unroll
            for (int r = 0; r < FR_MAXRHS; ++r)
                if (r < nrhs && lane < rows) bs[r0 + lane][r] = x[r];
        }
        __syncthreads();
        // rest of the leaf: rows after (forward) or before (backward) the sub-block
        const int lo = eff_lower ? r0 + rows : 0, hi = eff_lower ? m : r0;
        const int i = lo + tid - 32;
        if (warp != 0 && i < hi) {
            double acc[FR_MAXRHS];
#pragma unroll
            for (int r = 0; r < FR_MAXRHS; ++r) acc[r] = 0.0;
            for (int j = 0; j < rows; ++j) {
                const double t = S[i][r0 + j];
#pragma unroll
                for (int r = 0; r < FR_MAXRHS; ++r)
                    if (r < nrhs) acc[r] = fma(t, bs[r0 + j][r], acc[r]);
            }
#pragma unroll
            for (int r = 0; r < FR_MAXRHS; ++r)
                if (r < nrhs) bs[i][r] -= acc[r];
        }
        __syncthreads();
    }
    for (int idx = tid; idx < FR_LEAF * FR_MAXRHS; idx += 256) {
        const int i = idx & (FR_LEAF - 1), r = idx >> 7;
        if (i < m && r < nrhs) B[i + (i64)r * ldb] = bs[i][r];
    }
}

// C(m x nr) -= A(m x k) X(k x nr), A column-major: a CTA owns 32 rows, its 8 warps take the columns k0, k0+8, ... (every
// warp-level load is 32 consecutive rows of one column), partial sums are combined in warp order.
__global__ void __launch_bounds__(256) gemv_n_fewrhs_kernel(int m, int k, int nr, const double* __restrict__ A, i64 lda,
                                                            const double* __restrict__ X, i64 ldx, double* __restrict__ C, i64 ldc) {
    __shared__ double part[8][32][FR_MAXRHS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int i = blockIdx.x * 32 + lane;
    double acc[FR_MAXRHS];
#pragma unroll
    for (int r = 0; r < FR_MAXRHS; ++r) acc[r] = 0.0;
    if (i < m) {
        const double* a = A + i;
        int j = warp;
        for (; j + 56 < k; j += 64) {                 // 8 loads in flight per thread
            double v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) v[u] = __ldcs(a + (i64)(j + 8 * u) * lda);
#pragma unroll
            for (int u = 0; u < 8; ++u) {
#pragma unroll
                for (int r = 0; r < FR_MAXRHS; ++r)
                    if (r < nr) acc[r] = fma(v[u], __ldg(X + (j + 8 * u) + (i64)r * ldx), acc[r]);
            }
        }
        for (; j < k; j += 8) {
            const double v = __ldcs(a + (i64)j * lda);
#pragma unroll
            for (int r = 0; r < FR_MAXRHS; ++r)
                if (r < nr) acc[r] = fma(v, __ldg(X + j + (i64)r * ldx), acc[r]);
        }
    }
#pragma unroll
    for (int r = 0; r < FR_MAXRHS; ++r) part[warp][lane][r] = acc[r];
    __syncthreads();
    if (warp == 0 && i < m) {
#pragma unroll
        for (int r = 0; r < FR_MAXRHS; ++r) {
            if (r < nr) {
                double t = 0.0;
#pragma unroll
                for (int w = 0; w < 8; ++w) t += part[w][lane][r];
                C[i + (i64)r * ldc] -= t;
            }
        }
    }
}

// C(m x nr) -= A^T X with A stored k x m (column i of A is contiguous): one warp per output row, lanes stride the
// column, fixed-order shuffle reduction.
__global__ void __launch_bounds__(256) gemv_t_fewrhs_kernel(int m, int k, int nr, const double* __restrict__ A, i64 lda,
                                                            const double* __restrict__ X, i64 ldx, double* __restrict__ C, i64 ldc) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int i = blockIdx.x * 8 + warp;
    if (i >= m) return;
    const double* a = A + (i64)i * lda;
    double acc[FR_MAXRHS];
#pragma unroll
    for (int r = 0; r < FR_MAXRHS; ++r) acc[r] = 0.0;
    int j = lane;
    for (; j + 96 < k; j += 128) {
        double v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) v[u] = __ldcs(a + j + 32 * u);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
#pragma unroll
            for (int r = 0; r < FR_MAXRHS; ++r)
                if (r < nr) acc[r] = fma(v[u], __ldg(X + (j + 32 * u) + (i64)r * ldx), acc[r]);
        }
    }
    for (; j < k; j += 32) {
        const double v = __ldcs(a + j);
#pragma unroll
        for (int r = 0; r < FR_MAXRHS; ++r)
            if (r < nr) acc[r] = fma(v, __ldg(X + j + (i64)r * ldx), acc[r]);
    }
#pragma unroll
    for (int r = 0; r < FR_MAXRHS; ++r) {
        if (r < nr) {
            double t = acc[r];
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) t += __shfl_xor_sync(0xffffffffu, t, off);
            if (lane == 0) C[i + (i64)r * ldc] -= t;
        }
    }
}

static void gemv_fewrhs(cudaStream_t s, bool a_trans, int m, int k, int nr, const double* A, i64 lda, const double* X, i64 ldx,
                        double* C, i64 ldc) {
    if (m <= 0 || k <= 0) return;
    if (!a_trans) gemv_n_fewrhs_kernel<<<ceil_div(m, 32), 256, 0, s>>>(m, k, nr, A, lda, X, ldx, C, ldc);
    else gemv_t_fewrhs_kernel<<<ceil_div(m, 8), 256, 0, s>>>(m, k, nr, A, lda, X, ldx, C, ldc);
    count_launch();
}

static void trsm_left_fewrhs_rec(cudaStream_t s, bool upper, bool trans, bool unit, int m, int n, const double* A, i64 lda,
                                 double* B, i64 ldb) {
    const bool eff_lower = (upper == trans);
    if (m <= FR_LEAF) {
        const size_t smem = sizeof(double) * (FR_LEAF * (FR_LEAF + 1) + FR_LEAF * FR_MAXRHS + FR_LEAF);
        static bool attr = false;
        if (!attr) {
            LB_CUDA_CHECK(cudaFuncSetAttribute(trsm_left_fewrhs_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            LB_CUDA_CHECK(cudaFuncSetAttribute(trsm_left_fewrhs_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            attr = true;
        }
        if (eff_lower) trsm_left_fewrhs_kernel<true><<<1, 256, smem, s>>>(m, n, A, lda, trans, unit, B, ldb);
        else trsm_left_fewrhs_kernel<false><<<1, 256, smem, s>>>(m, n, A, lda, trans, unit, B, ldb);
        count_launch();
        return;
    }
    int m1 = FR_LEAF;
    while (m1 * 2 < m) m1 *= 2;
    const int m2 = m - m1;
    const double* A11 = A;
    const double* A22 = A + m1 + (i64)m1 * lda;
    const double* A21 = A + m1;                    // stored m2 x m1
    const double* A12 = A + (i64)m1 * lda;         // stored m1 x m2
    double* B1 = B;
    double* B2 = B + m1;
    if (eff_lower) {
        trsm_left_fewrhs_rec(s, upper, trans, unit, m1, n, A11, lda, B1, ldb);
        if (!trans) gemv_fewrhs(s, false, m2, m1, n, A21, lda, B1, ldb, B2, ldb);       // B2 -= A21 B1
        else gemv_fewrhs(s, true, m2, m1, n, A12, lda, B1, ldb, B2, ldb);               // B2 -= A12^T B1
        trsm_left_fewrhs_rec(s, upper, trans, unit, m2, n, A22, lda, B2, ldb);
    } else {
        trsm_left_fewrhs_rec(s, upper, trans, unit, m2, n, A22, lda, B2, ldb);
        if (!trans) gemv_fewrhs(s, false, m1, m2, n, A12, lda, B2, ldb, B1, ldb);       // B1 -= A12 B2
        else gemv_fewrhs(s, true, m1, m2, n, A21, lda, B2, ldb, B1, ldb);               // B1 -= A21^T B2
        trsm_left_fewrhs_rec(s, upper, trans, unit, m1, n, A11, lda, B1, ldb);
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
    } else if (left && n <= FR_MAXRHS && m > FR_LEAF) {
        if (g_fewrhs_mode == 1) trsv_stream(s, upper, tr, unit, m, n, A, lda, B, ldb);
        else trsm_left_fewrhs_rec(s, upper, tr, unit, m, n, A, lda, B, ldb);
    } else {
        const int nt = left ? m : n;                       // order of the triangle
        InvBlocks iv;
        double* scratch = nullptr;
        if (g_trsm_inverse_leaves && nt >= 2 * TB && (left ? n : m) >= 1024) {
            const int nblk = ceil_div(nt, TB);
            scratch = (double*)ws_alloc(s, sizeof(double) * (size_t)nblk * TB * TB);
            trtri32_blocks_kernel<<<nblk, 32, 0, s>>>(nt, A, lda, upper, unit, scratch, kernel_guard());
            count_launch();
            iv.base = scratch;
        }
        if (left) trsm_left_rec(s, upper, tr, unit, m, n, A, lda, B, ldb, iv);
        else trsm_right_rec(s, upper, tr, unit, m, n, A, lda, B, ldb, iv);
        if (scratch) ws_free(s, scratch);
    }
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
