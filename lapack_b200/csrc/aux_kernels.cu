// aux_kernels.cu -- memory-bound helper kernels: row interchanges (DLASWP), copies, transposes,
// the DLARNV/DLARUV generator with jump-ahead, and small integer fix-ups for IPIV/INFO.
#include "lb_internal.h"

namespace lb {

// ----------------------------------------------------------------------------------------------
// DLASWP (SRC/dlaswp.f:138-183).  Rows are LDA-strided in column-major storage, so a row swap touches
// one 32-byte sector per element.  Variant A: one thread per column walks the pivot list (held in
// shared memory) -- used when there are many columns.  Variant B: one CTA per column stages the column
// in shared memory, one thread applies the interchanges, all threads write back -- used for few columns
// (DGETRS right-hand sides).  Algorithmic bytes: 32 B per swapped pair per column.
constexpr int LASWP_MAX_PIV = 2048;

__global__ void laswp_cols_kernel(int n, double* __restrict__ A, i64 lda, int k1, int k2, const int* __restrict__ ipiv,
                                  int incx) {
    extern __shared__ int spiv[];
    const int np = k2 - k1 + 1;
    // pivot for row i (k1<=i<=k2) lives at ipiv[ix0 + (i-k1)*incx - 1] (forward) -- dlaswp.f:138-150
    const int ix0 = incx > 0 ? k1 : k1 + (k1 - k2) * incx;
    for (int t = threadIdx.x; t < np; t += blockDim.x) {
        // spiv[t] = pivot of the t-th interchange IN ORDER OF APPLICATION
        int i = incx > 0 ? k1 + t : k2 - t;
        int ix = ix0 + t * incx;
        (void)i;
        spiv[t] = ipiv[ix - 1];
    }
    __syncthreads();
    int col = blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= n) return;
    double* a = A + (i64)col * lda;
    for (int t = 0; t < np; ++t) {
        int i = incx > 0 ? k1 + t : k2 - t;
        int ip = spiv[t];
        if (ip != i) {
            double tmp = a[i - 1];
            a[i - 1] = a[ip - 1];
            a[ip - 1] = tmp;
        }
    }
}

__global__ void laswp_colsmem_kernel(int rows, double* __restrict__ A, i64 lda, int k1, int k2,
                                     const int* __restrict__ ipiv, int incx) {
    extern __shared__ double scol[];
    double* a = A + (i64)blockIdx.x * lda;
    for (int i = threadIdx.x; i < rows; i += blockDim.x) scol[i] = a[i];
    __syncthreads();
    if (threadIdx.x == 0) {
        const int np = k2 - k1 + 1;
        const int ix0 = incx > 0 ? k1 : k1 + (k1 - k2) * incx;
        for (int t = 0; t < np; ++t) {
            int i = incx > 0 ? k1 + t : k2 - t;
            int ip = ipiv[ix0 + t * incx - 1];
            if (ip != i) { double tmp = scol[i - 1]; scol[i - 1] = scol[ip - 1]; scol[ip - 1] = tmp; }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < rows; i += blockDim.x) a[i] = scol[i];
}

// Variant C (default for more than a handful of interchanges): the sequence of transpositions is first
// composed into a list of independent moves (dst_row <- src_row) by one small CTA, then every column applies
// that list with all loads in flight at once (gather into shared memory, barrier, scatter).  This removes the
// dependent load->store->load chain of variant A (about one memory round trip per interchange).
struct MoveList {
    int count;
    int pad;
    int2 mv[1];   // (dst_row, src_row), 0-based
};

__global__ void laswp_build_kernel(int np, int k1, int k2, const int* __restrict__ ipiv, int incx, MoveList* out) {
    extern __shared__ int sm[];
    int* piv = sm;              // [np]   pivot row (1-based) of the t-th interchange in application order
    int* slot = sm + np;        // [np]   slot of the pivot row: < np inside the block, >= np outside
    int* idx = sm + 2 * np;     // [2np]  which original slot currently sits in each slot
    int* rowof = sm + 4 * np;   // [2np]  global row (1-based) of each slot, 0 = unused
    __shared__ int s_count;
    const int ix0 = incx > 0 ? k1 : k1 + (k1 - k2) * incx;
    for (int t = threadIdx.x; t < np; t += blockDim.x) piv[t] = ipiv[ix0 + t * incx - 1];
    for (int q = threadIdx.x; q < 2 * np; q += blockDim.x) { idx[q] = q; rowof[q] = (q < np) ? k1 + q : 0; }
    if (threadIdx.x == 0) s_count = 0;
    __syncthreads();
    for (int t = threadIdx.x; t < np; t += blockDim.x) {
        int ip = piv[t];
        int sl;
        if (ip >= k1 && ip <= k2) sl = ip - k1;
        else {
            int first = t;
            for (int u = 0; u < t; ++u)
                if (piv[u] == ip) { first = u; break; }
            sl = np + first;
            if (first == t) rowof[sl] = ip;
        }
        slot[t] = sl;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int t = 0; t < np; ++t) {
            int i = incx > 0 ? k1 + t : k2 - t;     // row being interchanged at step t
            int a = i - k1, b = slot[t];
            int tmp = idx[a]; idx[a] = idx[b]; idx[b] = tmp;
        }
    }
    __syncthreads();
    for (int q = threadIdx.x; q < 2 * np; q += blockDim.x) {
        if (rowof[q] != 0 && idx[q] != q) {
            int pos = atomicAdd(&s_count, 1);
            out->mv[pos] = make_int2(rowof[q] - 1, rowof[idx[q]] - 1);
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) out->count = s_count;
}

constexpr int LASWP_CW = 4;   // columns per CTA in the apply kernel
__global__ void __launch_bounds__(256) laswp_apply_kernel(int n, double* __restrict__ A, i64 lda, const MoveList* __restrict__ ml) {
    extern __shared__ double sval[];
    const int cnt = ml->count;
    if (cnt == 0) return;
    const int c0 = blockIdx.x * LASWP_CW;
    const int nc = min(LASWP_CW, n - c0);
    const int total = cnt * nc;
    for (int q = threadIdx.x; q < total; q += blockDim.x) {
        int c = q / cnt, mvi = q - c * cnt;
        sval[q] = A[(i64)(c0 + c) * lda + ml->mv[mvi].y];
    }
    __syncthreads();
    for (int q = threadIdx.x; q < total; q += blockDim.x) {
        int c = q / cnt, mvi = q - c * cnt;
        A[(i64)(c0 + c) * lda + ml->mv[mvi].x] = sval[q];
    }
}

// max row touched must be known for variant B; the caller passes `rows_hint` (0 = unknown)
static void laswp_impl(cudaStream_t s, int n, double* A, i64 lda, int k1, int k2, const int* ipiv, int incx,
                       int rows_hint) {
    if (n <= 0 || incx == 0 || k2 < k1) return;
    (void)rows_hint;
    const int np_all = k2 - k1 + 1;
    if (np_all >= 4) {
        // process in chunks of at most 2048 interchanges (shared-memory bound of the build kernel)
        const int abs_inc = incx > 0 ? incx : -incx;
        static bool attr = false;
        if (!attr) {
            LB_CUDA_CHECK(cudaFuncSetAttribute(laswp_build_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 6 * LASWP_MAX_PIV * 4));
            LB_CUDA_CHECK(cudaFuncSetAttribute(laswp_apply_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                               2 * LASWP_MAX_PIV * LASWP_CW * 8));
            attr = true;
        }
        for (int done = 0; done < np_all; done += LASWP_MAX_PIV) {
            int cnt = min(LASWP_MAX_PIV, np_all - done);
            int ck1, ck2;
            const int* cpiv;
            if (incx > 0) { ck1 = k1 + done; ck2 = ck1 + cnt - 1; cpiv = ipiv + (i64)(ck1 - k1) * (incx - 1); }
            else { ck2 = k2 - done; ck1 = ck2 - cnt + 1; cpiv = ipiv + (i64)(k1 - ck1) + (i64)(k2 - ck2) * abs_inc; }
            MoveList* ml = (MoveList*)ws_alloc(s, sizeof(MoveList) + sizeof(int2) * 2 * cnt);
            laswp_build_kernel<<<1, 256, (size_t)6 * cnt * sizeof(int), s>>>(cnt, ck1, ck2, cpiv, incx, ml);
            laswp_apply_kernel<<<ceil_div(n, LASWP_CW), 256, (size_t)2 * cnt * LASWP_CW * sizeof(double), s>>>(n, A, lda, ml);
            count_launch(2);
            ws_free(s, ml);
        }
        LB_CUDA_CHECK(cudaGetLastError());
        return;
    }
    // chunk the pivot list so that it fits in shared memory, preserving the application order
    int abs_inc = incx > 0 ? incx : -incx;
    for (int done = 0; done < k2 - k1 + 1; done += LASWP_MAX_PIV) {
        int cnt = min(LASWP_MAX_PIV, k2 - k1 + 1 - done);
        int ck1, ck2;
        const int* cpiv = ipiv;
        if (incx > 0) { ck1 = k1 + done; ck2 = ck1 + cnt - 1; }
        else { ck2 = k2 - done; ck1 = ck2 - cnt + 1; }
        // For chunked calls the pivot of row i must still be found at the same address as in the full call.
        // Forward: address(i) = ipiv[(k1 + (i-k1)*incx) - 1]; a sub-call with k1'=ck1 uses ipiv'[(ck1 + (i-ck1)*incx) - 1],
        // so shift the base pointer by (ck1-k1)*(incx-1).  Reverse: address(i) = ipiv[(k1 + (k2-i)*abs) - 1]
        // (dlaswp.f:143-146); a sub-call uses ipiv'[(ck1 + (ck2-i)*abs) - 1], so shift by (k1-ck1) + (k2-ck2)*abs.
        if (incx > 0) cpiv = ipiv + (i64)(ck1 - k1) * (incx - 1);
        else cpiv = ipiv + (i64)(k1 - ck1) + (i64)(k2 - ck2) * abs_inc;
        int threads = 128;
        laswp_cols_kernel<<<ceil_div(n, threads), threads, (size_t)cnt * sizeof(int), s>>>(n, A, lda, ck1, ck2, cpiv, incx);
        count_launch();
    }
    LB_CUDA_CHECK(cudaGetLastError());
}

void laswp(cudaStream_t s, int n, double* A, i64 lda, int k1, int k2, const int* ipiv, int incx) {
    laswp_impl(s, n, A, lda, k1, k2, ipiv, incx, 0);
}
// Plan once, apply to several column ranges (the blocked LU driver applies one panel's interchanges to the
// look-ahead slab, the rest of the trailing matrix and the columns on the left).  k2-k1+1 <= LASWP_MAX_PIV.
void* laswp_plan(cudaStream_t s, int k1, int k2, const int* ipiv, int incx) {
    const int cnt = k2 - k1 + 1;
    if (cnt <= 0 || cnt > LASWP_MAX_PIV) return nullptr;
    static bool attr = false;
    if (!attr) {
        LB_CUDA_CHECK(cudaFuncSetAttribute(laswp_build_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 6 * LASWP_MAX_PIV * 4));
        LB_CUDA_CHECK(cudaFuncSetAttribute(laswp_apply_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           2 * LASWP_MAX_PIV * LASWP_CW * 8));
        attr = true;
    }
    MoveList* ml = (MoveList*)ws_alloc(s, sizeof(MoveList) + sizeof(int2) * 2 * cnt);
    laswp_build_kernel<<<1, 256, (size_t)6 * cnt * sizeof(int), s>>>(cnt, k1, k2, ipiv, incx, ml);
    count_launch();
    return ml;
}
void laswp_apply_plan(cudaStream_t s, int n, double* A, i64 lda, const void* plan, int npiv) {
    if (n <= 0 || !plan) return;
    laswp_apply_kernel<<<ceil_div(n, LASWP_CW), 256, (size_t)2 * npiv * LASWP_CW * sizeof(double), s>>>(n, A, lda, (const MoveList*)plan);
    count_launch();
}
void laswp_plan_free(cudaStream_t s, void* plan) { ws_free(s, plan); }
// variant with a known row extent (all pivots < rows): lets few-column calls use the staged kernel
void laswp_rows(cudaStream_t s, int n, int rows, double* A, i64 lda, int k1, int k2, const int* ipiv, int incx) {
    laswp_impl(s, n, A, lda, k1, k2, ipiv, incx, rows);
}

// ----------------------------------------------------------------------------------------------
__global__ void lacpy_kernel(int uplo, int m, int n, const double* __restrict__ A, i64 lda, double* __restrict__ B,
                             i64 ldb) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    for (int j = blockIdx.y; j < n; j += gridDim.y) {
        if (uplo == 1 && i > j) continue;     // upper: i <= j
        if (uplo == 2 && i < j) continue;     // lower: i >= j
        B[i + (i64)j * ldb] = A[i + (i64)j * lda];
    }
}
void lacpy(cudaStream_t s, char uplo, int m, int n, const double* A, i64 lda, double* B, i64 ldb) {
    if (m <= 0 || n <= 0) return;
    int u = (uplo == 'U' || uplo == 'u') ? 1 : (uplo == 'L' || uplo == 'l') ? 2 : 0;
    dim3 grid(ceil_div(m, 256), (unsigned)min(n, 8192));
    lacpy_kernel<<<grid, 256, 0, s>>>(u, m, n, A, lda, B, ldb);
    count_launch();
}

__global__ void laset_kernel(int uplo, int m, int n, double alpha, double beta, double* __restrict__ A, i64 lda) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    for (int j = blockIdx.y; j < n; j += gridDim.y) {
        if (i == j) { A[i + (i64)j * lda] = beta; continue; }
        if (uplo == 1 && i > j) continue;
        if (uplo == 2 && i < j) continue;
        A[i + (i64)j * lda] = alpha;
    }
}
void laset(cudaStream_t s, char uplo, int m, int n, double alpha, double beta, double* A, i64 lda) {
    if (m <= 0 || n <= 0) return;
    int u = (uplo == 'U' || uplo == 'u') ? 1 : (uplo == 'L' || uplo == 'l') ? 2 : 0;
    dim3 grid(ceil_div(m, 256), (unsigned)min(n, 8192));
    laset_kernel<<<grid, 256, 0, s>>>(u, m, n, alpha, beta, A, lda);
    count_launch();
}

// B (n x m) = A(m x n)^T through a padded 32x32 shared tile (coalesced on both sides)
__global__ void transpose_kernel(int m, int n, const double* __restrict__ A, i64 lda, double* __restrict__ B, i64 ldb) {
    __shared__ double tile[32][33];
    int bi = blockIdx.x * 32, bj = blockIdx.y * 32;
    for (int jj = threadIdx.y; jj < 32; jj += blockDim.y) {
        int i = bi + threadIdx.x, j = bj + jj;
        if (i < m && j < n) tile[jj][threadIdx.x] = A[i + (i64)j * lda];
    }
    __syncthreads();
    for (int ii = threadIdx.y; ii < 32; ii += blockDim.y) {
        int j = bj + threadIdx.x, i = bi + ii;
        if (i < m && j < n) B[j + (i64)i * ldb] = tile[threadIdx.x][ii];
    }
}
void transpose(cudaStream_t s, int m, int n, const double* A, i64 lda, double* B, i64 ldb) {
    if (m <= 0 || n <= 0) return;
    dim3 grid(ceil_div(m, 32), ceil_div(n, 32)), block(32, 8);
    transpose_kernel<<<grid, block, 0, s>>>(m, n, A, lda, B, ldb);
    count_launch();
}

// ----------------------------------------------------------------------------------------------
// DLARUV/DLARNV (SRC/dlaruv.f:401-447, SRC/dlarnv.f:140-170): x_k = seed * a^k mod 2^48, a = 33952834046453.
// DLARNV consumes the stream in order, so draw k of the whole stream is independent of the 64/128 chunking.
__device__ __forceinline__ unsigned long long lcg_pow(unsigned long long k) {
    const unsigned long long MASK = (1ULL << 48) - 1;
    unsigned long long r = 1, b = 33952834046453ULL;
    while (k) {
        if (k & 1) r = (r * b) & MASK;
        b = (b * b) & MASK;
        k >>= 1;
    }
    return r;
}
constexpr int LARNV_CHUNK = 8;
__global__ void larnv_matrix_kernel(unsigned long long seed, i64 offset, int m, int n, double* __restrict__ A, i64 lda) {
    const unsigned long long MASK = (1ULL << 48) - 1, AMUL = 33952834046453ULL;
    i64 total = (i64)m * n;
    i64 c0 = ((i64)blockIdx.x * blockDim.x + threadIdx.x) * LARNV_CHUNK;
    if (c0 >= total) return;
    unsigned long long st = (seed * lcg_pow((unsigned long long)(offset + c0))) & MASK;   // state before draw c0+1
#pragma unroll
    for (int q = 0; q < LARNV_CHUNK; ++q) {
        i64 e = c0 + q;
        if (e >= total) break;
        st = (st * AMUL) & MASK;
        double u = (double)st * (1.0 / 281474976710656.0);
        i64 j = e / m, i = e - j * m;
        A[i + j * lda] = 2.0 * u - 1.0;
    }
}
void larnv_matrix(cudaStream_t s, const int iseed[4], i64 stream_offset, int m, int n, double* A, i64 lda) {
    if (m <= 0 || n <= 0) return;
    unsigned long long seed = ((unsigned long long)iseed[0] << 36) | ((unsigned long long)iseed[1] << 24) |
                              ((unsigned long long)iseed[2] << 12) | (unsigned long long)iseed[3];
    i64 total = (i64)m * n;
    i64 threads = (total + LARNV_CHUNK - 1) / LARNV_CHUNK;
    larnv_matrix_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, s>>>(seed, stream_offset, m, n, A, lda);
    count_launch();
}
void larnv_fill(cudaStream_t s, int idist, const int iseed[4], i64 offset, i64 count, double* x) {
    (void)idist;   // only IDIST=2 (uniform(-1,1)) is generated on the device
    i64 done = 0;
    while (done < count) {   // rows limited to int
        int chunk = (int)((count - done) < (1LL << 30) ? (count - done) : (1LL << 30));
        larnv_matrix(s, iseed, offset + done, chunk, 1, x + done, chunk);
        done += chunk;
    }
}

// A := (A + A^T)/2 + shift*I, in place, tile pairs (bi >= bj)
__global__ void make_spd_kernel(int n, double* __restrict__ A, i64 lda, double shift) {
    __shared__ double t1[32][33], t2[32][33];
    int bi = blockIdx.x, bj = blockIdx.y;
    if (bi < bj) return;
    for (int jj = threadIdx.y; jj < 32; jj += blockDim.y) {
        int i = bi * 32 + threadIdx.x, j = bj * 32 + jj;
        if (i < n && j < n) t1[jj][threadIdx.x] = A[i + (i64)j * lda];        // A(bi-block, bj-block)
        int i2 = bj * 32 + threadIdx.x, j2 = bi * 32 + jj;
        if (i2 < n && j2 < n) t2[jj][threadIdx.x] = A[i2 + (i64)j2 * lda];    // A(bj-block, bi-block)
    }
    __syncthreads();
    for (int jj = threadIdx.y; jj < 32; jj += blockDim.y) {
        int i = bi * 32 + threadIdx.x, j = bj * 32 + jj;
        if (i < n && j < n) {
            double v = (t1[jj][threadIdx.x] + t2[threadIdx.x][jj]) * 0.5;
            if (i == j) v += shift;
            A[i + (i64)j * lda] = v;
        }
        int i2 = bj * 32 + threadIdx.x, j2 = bi * 32 + jj;
        if (bi != bj && i2 < n && j2 < n) {
            double v = (t2[jj][threadIdx.x] + t1[threadIdx.x][jj]) * 0.5;
            A[i2 + (i64)j2 * lda] = v;
        }
    }
}
void make_spd(cudaStream_t s, int n, double* A, i64 lda, double shift) {
    if (n <= 0) return;
    dim3 grid(ceil_div(n, 32), ceil_div(n, 32)), block(32, 8);
    make_spd_kernel<<<grid, block, 0, s>>>(n, A, lda, shift);
    count_launch();
}

__global__ void iadd_kernel(int n, int* x, int v) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) x[i] += v;
}
void iadd(cudaStream_t s, int n, int* x, int v) {
    if (n <= 0 || v == 0) return;
    iadd_kernel<<<ceil_div(n, 256), 256, 0, s>>>(n, x, v);
    count_launch();
}

// LU INFO rule (dgetrf.f:185-186, dgetrf2.f:229-230,254-255): keep the first non-zero, shifted by the block offset
__global__ void info_first_kernel(int* info, const int* iinfo, int offset) {
    if (*info == 0 && *iinfo > 0) *info = *iinfo + offset;
}
void info_max_offset(cudaStream_t s, int* info, const int* iinfo, int offset) {
    info_first_kernel<<<1, 1, 0, s>>>(info, iinfo, offset);
    count_launch();
}

}  // namespace lb
