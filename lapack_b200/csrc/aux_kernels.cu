// aux_kernels.cu -- memory-bound helper kernels: row interchanges (DLASWP), copies, transposes,
// the DLARNV/DLARUV generator with jump-ahead, and small integer fix-ups for IPIV/INFO.
#include "lb_internal.h"
#include <cstdint>

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

// Variant C (default for more than a handful of interchanges).  A sequence of transpositions is first COMPOSED into one
// permutation by a small CTA (O(np): one thread walks the sequence on an index array in shared memory; rows outside the
// pivot block K1..K2 live in a small hash table), which yields independent moves in two groups:
//   group 1: block row K1+t  <- original row srcA[t]      (destinations contiguous: coalesced writes)
//   group 2: outside row dst <- original row src          (sorted by src, which is a block row for LU pivots: coalesced reads)
// The apply kernel then gathers every moved element of a few columns into shared memory with all loads in flight at once,
// and scatters.  DRAM traffic per interchanged pair and column: one 32-B sector read + written back for the row outside
// the block (the write hits the sector the same CTA has just read), 8 + 8 B for the block row -- 80 B for 32 algorithmic
// bytes, which is what LDA-strided rows allow.
struct SwapPlan {
    int np;       // block rows K1 .. K1+np-1
    int k1;       // 1-based
    int nout;     // entries of group 2
    int nmoved;   // block rows whose content changes (informational)
    int data[1];  // [np] srcA (0-based row, == K1-1+t when unchanged), then up to 2*np (dst, src) pairs (second half: scratch)
};

constexpr int LASWP_HASH = 8192;      // slots for rows outside the block (<= LASWP_MAX_PIV of them), power of two
__device__ __forceinline__ int swp_hash(int row) { return (int)(((unsigned)row * 2654435761u) >> 19) & (LASWP_HASH - 1); }

__global__ void __launch_bounds__(256) laswp_build_kernel(int np, int k1, int k2, const int* __restrict__ ipiv, int incx, SwapPlan* out) {
    extern __shared__ int sm[];
    int* piv = sm;                  // [np]   pivot row (1-based) of the t-th interchange in application order
    int* cur = sm + np;             // [np]   original row (1-based) currently stored in block row K1+q
    int* inv = sm + 2 * np;         // [np]   final row (1-based) of the original block row K1+q
    int* hkey = sm + 3 * np;        // [HASH] outside row (1-based), 0 = empty
    int* hval = hkey + LASWP_HASH;  // [HASH] original row currently stored there
    __shared__ int s_cnt[256];
    __shared__ int s_extra;
    const int tid = threadIdx.x;
    const int ix0 = incx > 0 ? k1 : k1 + (k1 - k2) * incx;
    for (int t = tid; t < np; t += 256) { piv[t] = ipiv[ix0 + t * incx - 1]; cur[t] = k1 + t; inv[t] = 0; }
    for (int q = tid; q < LASWP_HASH; q += 256) hkey[q] = 0;
    if (tid == 0) s_extra = 0;
    __syncthreads();
    if (tid == 0) {
        for (int t = 0; t < np; ++t) {
            const int i = incx > 0 ? k1 + t : k2 - t;     // row interchanged at step t (dlaswp.f:152-167)
            const int ip = piv[t];
            if (ip == i) continue;
            const int a = cur[i - k1];
            if (ip >= k1 && ip <= k2) {
                cur[i - k1] = cur[ip - k1];
                cur[ip - k1] = a;
            } else {
                int h = swp_hash(ip);
                while (hkey[h] != 0 && hkey[h] != ip) h = (h + 1) & (LASWP_HASH - 1);
                if (hkey[h] == 0) { hkey[h] = ip; hval[h] = ip; }
                cur[i - k1] = hval[h];
                hval[h] = a;
            }
        }
    }
    __syncthreads();
    // group 1 and the inverse map of the block-origin rows
    int* srcA = out->data;
    int2* pairs = reinterpret_cast<int2*>(out->data + np + (np & 1));
    int moved = 0;
    for (int t = tid; t < np; t += 256) {
        const int o = cur[t];
        srcA[t] = o - 1;
        if (o != k1 + t) ++moved;
        if (o >= k1 && o <= k2) inv[o - k1] = k1 + t;
    }
    for (int q = tid; q < LASWP_HASH; q += 256) {
        const int r = hkey[q];
        if (r != 0 && hval[q] != r) {
            const int o = hval[q];
            if (o >= k1 && o <= k2) inv[o - k1] = -r;             // block-origin row that ends outside the block
            else { const int pos = atomicAdd(&s_extra, 1); pairs[np + pos] = make_int2(r - 1, o - 1); }   // outside -> outside (general pivots only)
        }
    }
    __syncthreads();
    // group 2 in the order of the source row: per-thread counts over a contiguous range, exclusive scan, write
    const int per = (np + 255) / 256;
    int cnt = 0;
    for (int q = tid * per; q < min(np, (tid + 1) * per); ++q) cnt += inv[q] < 0;
    s_cnt[tid] = cnt;
    __syncthreads();
    if (tid == 0) { int run = 0; for (int q = 0; q < 256; ++q) { const int c = s_cnt[q]; s_cnt[q] = run; run += c; } out->nout = run + s_extra; }
    __syncthreads();
    int pos = s_cnt[tid];
    for (int q = tid * per; q < min(np, (tid + 1) * per); ++q)
        if (inv[q] < 0) pairs[pos++] = make_int2(-inv[q] - 1, k1 + q - 1);
    __syncthreads();
    // the outside -> outside entries were parked behind the first np slots; move them behind the sorted ones
    if (tid == 0) {
        const int base = out->nout - s_extra;
        for (int e = 0; e < s_extra; ++e) pairs[base + e] = pairs[np + e];
        out->np = np; out->k1 = k1;
    }
    // (counting `moved` is informational only)
    if (moved) atomicAdd(&out->nmoved, moved);
}

// one CTA = CW consecutive columns; shared memory: (np + nout) * CW doubles
__global__ void __launch_bounds__(256) laswp_apply_kernel(int n, double* __restrict__ A, i64 lda, const SwapPlan* __restrict__ pl, int cw) {
    extern __shared__ double sval[];
    const int np = pl->np, nout = pl->nout, k1 = pl->k1;
    if (pl->nmoved == 0 && nout == 0) return;
    const int* __restrict__ srcA = pl->data;
    const int2* __restrict__ pairs = reinterpret_cast<const int2*>(pl->data + np + (np & 1));
    const int c0 = blockIdx.x * cw;
    const int nc = min(cw, n - c0);
    double* Ac = A + (i64)c0 * lda;
    const int tot = np + nout;
    // gather: thread = move, inner loop = columns (all loads of a thread in flight together)
    for (int q = threadIdx.x; q < tot; q += 256) {
        int src;
        bool need;
        if (q < np) { src = srcA[q]; need = src != k1 - 1 + q; } else { src = pairs[q - np].y; need = true; }
        if (need) {
#pragma unroll 4
            for (int c = 0; c < nc; ++c) sval[(size_t)c * tot + q] = Ac[(i64)c * lda + src];
        }
    }
    __syncthreads();
    for (int q = threadIdx.x; q < tot; q += 256) {
        int dst;
        bool need;
        if (q < np) { dst = k1 - 1 + q; need = srcA[q] != dst; } else { dst = pairs[q - np].x; need = true; }
        if (need) {
#pragma unroll 4
            for (int c = 0; c < nc; ++c) Ac[(i64)c * lda + dst] = sval[(size_t)c * tot + q];
        }
    }
}

// Variant with the SCATTERED reads issued as 16-byte bulk copies (cp.async.bulk, SASS UBLKCP: the TMA engine's linear mode) into a
// shared-memory stage, completion counted on one mbarrier.  Motivation (profiles/r02_laswp_ncu.txt): an LDG that misses L2 on an
// LDA-strided row is filled as a full 128-byte line, 16x the 8 useful bytes; the bulk-copy path fetches the 16-byte pair it asks for.
// Needs 16-byte aligned pairs: A 16-byte aligned and lda even (else the LDG kernel above is used).
__device__ __forceinline__ unsigned swp_smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__global__ void __launch_bounds__(256) laswp_apply_bulk_kernel(int n, double* __restrict__ A, i64 lda, const SwapPlan* __restrict__ pl, int cw) {
    extern __shared__ __align__(16) double sval[];         // [cw][np] 16-byte slots (group 1), then [cw][nout] doubles (group 2)
    __shared__ __align__(8) unsigned long long mbar;
    const int np = pl->np, nout = pl->nout, k1 = pl->k1;
    if (pl->nmoved == 0 && nout == 0) return;
    const int* __restrict__ srcA = pl->data;
    const int2* __restrict__ pairs = reinterpret_cast<const int2*>(pl->data + np + (np & 1));
    const int c0 = blockIdx.x * cw;
    const int nc = min(cw, n - c0);
    double* Ac = A + (i64)c0 * lda;
    double2* s1 = reinterpret_cast<double2*>(sval);
    double* s2 = sval + (size_t)2 * np * cw;
    const unsigned bar = swp_smem_u32(&mbar);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned bytes = (unsigned)pl->nmoved * (unsigned)nc * 16u;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(bytes) : "memory");
    }
    // group 1: block row k1-1+q <- row srcA[q]; one 16-byte bulk copy of the aligned row pair per (move, column)
    for (int q = threadIdx.x; q < np; q += 256) {
        const int src = srcA[q];
        if (src != k1 - 1 + q) {
            const double* g = Ac + (src & ~1);
#pragma unroll 4
            for (int c = 0; c < nc; ++c) {
                const unsigned dst = swp_smem_u32(s1 + (size_t)c * np + q);
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], 16, [%2];\n"
                             ::"r"(dst), "l"(g + (i64)c * lda), "r"(bar) : "memory");
            }
        }
    }
    // group 2: outside row <- (mostly) block row: sources contiguous, plain loads
    for (int q = threadIdx.x; q < nout; q += 256) {
        const int src = pairs[q].y;
#pragma unroll 4
        for (int c = 0; c < nc; ++c) s2[(size_t)c * nout + q] = Ac[(i64)c * lda + src];
    }
    {   // wait for the bulk copies (phase 0)
        unsigned done = 0;
        while (!done) {
            asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
                         : "=r"(done) : "r"(bar), "r"(0) : "memory");
        }
    }
    __syncthreads();
    for (int q = threadIdx.x; q < np; q += 256) {
        const int src = srcA[q], dst = k1 - 1 + q;
        if (src != dst) {
#pragma unroll 4
            for (int c = 0; c < nc; ++c) {
                const double2 v = s1[(size_t)c * np + q];
                Ac[(i64)c * lda + dst] = (src & 1) ? v.y : v.x;
            }
        }
    }
    for (int q = threadIdx.x; q < nout; q += 256) {
        const int dst = pairs[q].x;
#pragma unroll 4
        for (int c = 0; c < nc; ++c) Ac[(i64)c * lda + dst] = s2[(size_t)c * nout + q];
    }
}
static int g_laswp_bulk = 0;     // 1: scattered reads through cp.async.bulk (experiment knob, lb200_set_laswp_bulk)
void laswp_set_bulk(int on) { g_laswp_bulk = on; }

static size_t swap_plan_bytes(int np) { return sizeof(SwapPlan) + sizeof(int) * ((size_t)np + 1 + 4 * (size_t)np + 2); }
static void laswp_attr() {
    static bool attr = false;
    if (!attr) {
        LB_CUDA_CHECK(cudaFuncSetAttribute(laswp_build_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (3 * LASWP_MAX_PIV + 2 * LASWP_HASH) * 4));
        LB_CUDA_CHECK(cudaFuncSetAttribute(laswp_apply_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * LASWP_MAX_PIV * 4 * 8));
        attr = true;
    }
}
static SwapPlan* swap_plan_build(cudaStream_t s, int k1, int k2, const int* ipiv, int incx) {
    const int cnt = k2 - k1 + 1;
    laswp_attr();
    SwapPlan* pl = (SwapPlan*)ws_alloc(s, swap_plan_bytes(cnt));
    LB_CUDA_CHECK(cudaMemsetAsync(pl, 0, sizeof(SwapPlan), s));
    laswp_build_kernel<<<1, 256, (size_t)(3 * cnt + 2 * LASWP_HASH) * sizeof(int), s>>>(cnt, k1, k2, ipiv, incx, pl);
    count_launch();
    return pl;
}
static void swap_plan_apply(cudaStream_t s, int n, double* A, i64 lda, const SwapPlan* pl, int npiv) {
    if (n <= 0 || !pl) return;
    // columns per CTA: as many as fit next to 2*npiv staged rows (<= 128 KB), at most 8, and enough CTAs to fill the GPU
    int cw = (int)((size_t)(2 * LASWP_MAX_PIV * 4) / (size_t)(2 * npiv));
    cw = max(1, min(8, cw));
    while (cw > 1 && ceil_div(n, cw) < 2 * num_sms()) cw >>= 1;
    if (g_laswp_bulk && (lda & 1) == 0 && (((uintptr_t)A) & 15) == 0) {
        // 16-byte slots for group 1 + 8-byte slots for group 2: (2 np + np) * cw doubles
        int cwb = max(1, min(cw, (int)((size_t)(2 * LASWP_MAX_PIV * 4) / (size_t)(3 * npiv))));
        static bool attr = false;
        if (!attr) {
            LB_CUDA_CHECK(cudaFuncSetAttribute(laswp_apply_bulk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * LASWP_MAX_PIV * 4 * 8));
            attr = true;
        }
        laswp_apply_bulk_kernel<<<ceil_div(n, cwb), 256, (size_t)3 * npiv * cwb * sizeof(double), s>>>(n, A, lda, pl, cwb);
        count_launch();
        return;
    }
    laswp_apply_kernel<<<ceil_div(n, cw), 256, (size_t)2 * npiv * cw * sizeof(double), s>>>(n, A, lda, pl, cw);
    count_launch();
}

// ---- few columns, many interchanges (DGETRS right-hand sides: all n pivots on nrhs columns; dgetrs.f:187,217).  The whole
// sequence is composed on an index array in shared memory by ONE thread (rows up to LASWP_LONG_MAX), then every column is
// permuted by a gather into scratch and a copy back.
constexpr int LASWP_LONG_MAX = 55 * 1024;
constexpr int LASWP_LONG_TILE = 1024;
__global__ void __launch_bounds__(256) laswp_long_build_kernel(int rows, int k1, int k2, const int* __restrict__ ipiv, int incx,
                                                               int* __restrict__ src_out) {
    extern __shared__ int cur[];      // cur[r] = original row (0-based) stored in row r; then one tile of pivots
    int* ptile = cur + rows;
    for (int r = threadIdx.x; r < rows; r += 256) cur[r] = r;
    const int np = k2 - k1 + 1;
    const int ix0 = incx > 0 ? k1 : k1 + (k1 - k2) * incx;
    for (int t0 = 0; t0 < np; t0 += LASWP_LONG_TILE) {
        const int tn = min(LASWP_LONG_TILE, np - t0);
        __syncthreads();
        for (int q = threadIdx.x; q < tn; q += 256) ptile[q] = ipiv[ix0 + (t0 + q) * incx - 1];
        __syncthreads();
        if (threadIdx.x == 0) {
            for (int q = 0; q < tn; ++q) {
                const int t = t0 + q;
                const int i = incx > 0 ? k1 + t : k2 - t;          // dlaswp.f:152-167
                const int ip = ptile[q];
                if (ip != i) { const int a = cur[i - 1]; cur[i - 1] = cur[ip - 1]; cur[ip - 1] = a; }
            }
        }
    }
    __syncthreads();
    for (int r = threadIdx.x; r < rows; r += 256) src_out[r] = cur[r];
}
__global__ void laswp_long_gather_kernel(int rows, int n, const double* __restrict__ A, i64 lda, const int* __restrict__ src,
                                         double* __restrict__ W) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    const int sr = src[r];
    for (int c = blockIdx.y; c < n; c += gridDim.y) W[r + (i64)c * rows] = A[sr + (i64)c * lda];
}
__global__ void laswp_long_copy_kernel(int rows, int n, const double* __restrict__ W, const int* __restrict__ src, double* __restrict__ A, i64 lda) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    if (src[r] == r) return;
    for (int c = blockIdx.y; c < n; c += gridDim.y) A[r + (i64)c * lda] = W[r + (i64)c * rows];
}

static void laswp_stream_permute(cudaStream_t s, int rows, int rmin, int n, double* A, i64 lda, const int* src);
static void laswp_impl(cudaStream_t s, int n, double* A, i64 lda, int k1, int k2, const int* ipiv, int incx,
                       int rows_hint) {
    if (n <= 0 || incx == 0 || k2 < k1) return;
    const int np_all = k2 - k1 + 1;
    const int abs_inc = incx > 0 ? incx : -incx;
    if (rows_hint > 0 && rows_hint <= LASWP_LONG_MAX && np_all > LASWP_MAX_PIV / 4 && n <= 64) {
        static bool attr = false;
        if (!attr) {
            LB_CUDA_CHECK(cudaFuncSetAttribute(laswp_long_build_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                               (LASWP_LONG_MAX + LASWP_LONG_TILE) * 4));
            attr = true;
        }
        int* src = (int*)ws_alloc(s, sizeof(int) * (size_t)rows_hint);
        double* W = (double*)ws_alloc(s, sizeof(double) * (size_t)rows_hint * n);
        laswp_long_build_kernel<<<1, 256, (size_t)(rows_hint + LASWP_LONG_TILE) * sizeof(int), s>>>(rows_hint, k1, k2, ipiv, incx, src);
        dim3 grid(ceil_div(rows_hint, 256), (unsigned)n);
        laswp_long_gather_kernel<<<grid, 256, 0, s>>>(rows_hint, n, A, lda, src, W);
        laswp_long_copy_kernel<<<grid, 256, 0, s>>>(rows_hint, n, W, src, A, lda);
        count_launch(3);
        ws_free(s, src);
        ws_free(s, W);
        LB_CUDA_CHECK(cudaGetLastError());
        return;
    }
    if (rows_hint > 0 && rows_hint <= LASWP_LONG_MAX && np_all >= 2 * LASWP_MAX_PIV && n > 64) {
        // many columns AND a long pivot list (the two interchange sweeps of the recursive host-streamed DGETRF, DGETRS with many
        // right-hand sides): compose the whole sequence once, then stream every column through its permutation (one pass over the
        // touched rows instead of np/2048 plans of scattered line fills each)
        static bool attr = false;
        if (!attr) {
            LB_CUDA_CHECK(cudaFuncSetAttribute(laswp_long_build_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                               (LASWP_LONG_MAX + LASWP_LONG_TILE) * 4));
            attr = true;
        }
        int* src = (int*)ws_alloc(s, sizeof(int) * (size_t)rows_hint);
        laswp_long_build_kernel<<<1, 256, (size_t)(rows_hint + LASWP_LONG_TILE) * sizeof(int), s>>>(rows_hint, k1, k2, ipiv, incx, src);
        count_launch();
        laswp_stream_permute(s, rows_hint, k1 - 1, n, A, lda, src);
        ws_free(s, src);
        LB_CUDA_CHECK(cudaGetLastError());
        return;
    }
    if (np_all >= 4) {
        // process in chunks of at most 2048 interchanges (shared-memory bound of the build kernel)
        for (int done = 0; done < np_all; done += LASWP_MAX_PIV) {
            int cnt = min(LASWP_MAX_PIV, np_all - done);
            int ck1, ck2;
            if (incx > 0) { ck1 = k1 + done; ck2 = ck1 + cnt - 1; }
            else { ck2 = k2 - done; ck1 = ck2 - cnt + 1; }
            // the pivot of row i sits at ipiv[(k1 + (i-k1)*|incx|) - 1] in both directions (dlaswp.f:138-150); a sub-call with
            // k1' = ck1 looks at ipiv'[(ck1 + (i-ck1)*|incx|) - 1], so the base pointer moves by (ck1-k1)*(|incx|-1)
            const int* cpiv = ipiv + (i64)(ck1 - k1) * (abs_inc - 1);
            SwapPlan* pl = swap_plan_build(s, ck1, ck2, cpiv, incx);
            swap_plan_apply(s, n, A, lda, pl, cnt);
            ws_free(s, pl);
        }
        LB_CUDA_CHECK(cudaGetLastError());
        return;
    }
    // a handful of interchanges: one thread per column walks the list
    {
        int threads = 128;
        laswp_cols_kernel<<<ceil_div(n, threads), threads, (size_t)np_all * sizeof(int), s>>>(n, A, lda, k1, k2, ipiv, incx);
        count_launch();
    }
    LB_CUDA_CHECK(cudaGetLastError());
}

void laswp(cudaStream_t s, int n, double* A, i64 lda, int k1, int k2, const int* ipiv, int incx) {
    laswp_impl(s, n, A, lda, k1, k2, ipiv, incx, 0);
}
// Plan once, apply to several column ranges (the blocked LU driver applies one panel's interchanges to the
// look-ahead slab, the rest of the trailing matrix and the columns on the left).  k2-k1+1 <= LASWP_MAX_PIV.
// ---- deferred interchanges left of the panels (getrf.cu): the plans of D consecutive panels, plan d to be applied to the columns
// [0, (d+1)*nb), are COMPOSED per block column -- block column b receives plans b .. D-1 in that order -- into one row permutation
// sigma_b (new row i <- old row sigma_b(i)), and every column is then permuted ONCE: gathered through its permutation (scattered
// 8-byte reads, but of a column that one CTA is streaming through L2 anyway) and written back.  DRAM traffic ~ the touched part of
// the columns once, instead of one 128-byte line fill per moved element and panel (profiles/r02_laswp_ncu.txt).
//   compose: one CTA; t = the current plan's map tau in shared memory (identity elsewhere), sig <- tau(sig), from the last plan
//   to the first: sigma_b(i) = tau_{b}(sigma_{b+1}(i)).
constexpr int LASWP_CHAIN_MAX = 120;
struct SwapPlanList { const SwapPlan* p[LASWP_CHAIN_MAX]; };
__global__ void __launch_bounds__(1024) laswp_chain_compose_kernel(int m, int nplans, const SwapPlanList plans, int* __restrict__ sigma) {
    extern __shared__ int tmap[];                    // [m]
    const int tid = threadIdx.x;
    for (int i = tid; i < m; i += 1024) tmap[i] = i;
    __syncthreads();
    for (int d = nplans - 1; d >= 0; --d) {
        const SwapPlan* pl = plans.p[d];
        const int np = pl->np, nout = pl->nout, k1 = pl->k1;
        const int* srcA = pl->data;
        const int2* pairs = reinterpret_cast<const int2*>(pl->data + np + (np & 1));
        for (int q = tid; q < np; q += 1024) tmap[k1 - 1 + q] = srcA[q];
        for (int q = tid; q < nout; q += 1024) tmap[pairs[q].x] = pairs[q].y;
        __syncthreads();
        int* sig = sigma + (size_t)d * m;
        const int* prev = (d + 1 < nplans) ? sigma + (size_t)(d + 1) * m : nullptr;
        for (int i = tid; i < m; i += 1024) sig[i] = tmap[prev ? prev[i] : i];
        __syncthreads();
        for (int q = tid; q < np; q += 1024) tmap[k1 - 1 + q] = k1 - 1 + q;
        for (int q = tid; q < nout; q += 1024) tmap[pairs[q].x] = pairs[q].x;
        __syncthreads();
    }
}
// persistent CTAs, one column at a time: gather the moved rows into this CTA's scratch column, then write them back
// nb > 0: column c belongs to block column b = c / nb with its own permutation sigma + b*m, rows below (b+1)*nb;
// nb == 0: one permutation for all columns, rows from rmin on (long pivot sequences, laswp_impl)
__global__ void __launch_bounds__(1024) laswp_chain_apply_kernel(int m, int nb, int ncols, double* __restrict__ A, i64 lda,
                                                                 const int* __restrict__ sigma, double* __restrict__ W, int rmin) {
    double* w = W + (size_t)blockIdx.x * m;
    for (int c = blockIdx.x; c < ncols; c += gridDim.x) {
        const int b = nb > 0 ? c / nb : 0;
        const int r0 = nb > 0 ? (b + 1) * nb : rmin; // rows above belong to earlier panels: untouched by plans b ..
        const int* sig = sigma + (size_t)b * m;
        double* col = A + (i64)c * lda;
        for (int i = r0 + threadIdx.x; i < m; i += 1024) {
            const int src = sig[i];
            if (src != i) w[i] = col[src];
        }
        __syncthreads();
        for (int i = r0 + threadIdx.x; i < m; i += 1024)
            if (sig[i] != i) col[i] = w[i];
        __syncthreads();
    }
}
static void laswp_stream_permute(cudaStream_t s, int rows, int rmin, int n, double* A, i64 lda, const int* src) {
    const int grid = min(2 * num_sms(), n);
    double* W = (double*)ws_alloc(s, sizeof(double) * (size_t)grid * rows);
    laswp_chain_apply_kernel<<<grid, 1024, 0, s>>>(rows, 0, n, A, lda, src, W, rmin);
    count_launch();
    ws_free(s, W);
}
// plans[d] (device SwapPlan pointers, host array) applies to columns [0, (d+1)*nb); false = not applicable (caller falls back)
bool laswp_apply_chain(cudaStream_t s, int m, int nb, int nplans, void* const* plans_host, double* A, i64 lda) {
    if (nplans <= 0) return true;
    const size_t smem = sizeof(int) * (size_t)m;
    if (smem > 200 * 1024 || nplans > LASWP_CHAIN_MAX) return false;
    static bool attr = false;
    if (!attr) {
        LB_CUDA_CHECK(cudaFuncSetAttribute(laswp_chain_compose_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        attr = true;
    }
    const int grid = 2 * num_sms();
    SwapPlanList list;
    for (int d = 0; d < nplans; ++d) list.p[d] = (const SwapPlan*)plans_host[d];
    int* sigma = (int*)ws_alloc(s, sizeof(int) * (size_t)nplans * m);
    double* W = (double*)ws_alloc(s, sizeof(double) * (size_t)grid * m);
    laswp_chain_compose_kernel<<<1, 1024, smem, s>>>(m, nplans, list, sigma);
    laswp_chain_apply_kernel<<<grid, 1024, 0, s>>>(m, nb, nplans * nb, A, lda, sigma, W, 0);
    count_launch(2);
    ws_free(s, W); ws_free(s, sigma);
    LB_CUDA_CHECK(cudaGetLastError());
    return true;
}

void* laswp_plan(cudaStream_t s, int k1, int k2, const int* ipiv, int incx) {
    const int cnt = k2 - k1 + 1;
    if (cnt <= 0 || cnt > LASWP_MAX_PIV) return nullptr;
    return swap_plan_build(s, k1, k2, ipiv, incx);
}
void laswp_apply_plan(cudaStream_t s, int n, double* A, i64 lda, const void* plan, int npiv) {
    swap_plan_apply(s, n, A, lda, (const SwapPlan*)plan, npiv);
}
void laswp_plan_free(cudaStream_t s, void* plan) { ws_free(s, plan); }
// variant with a known row extent (every pivot <= rows): lets few-column calls with long pivot lists use the composed path
void laswp_rows(cudaStream_t s, int n, int rows, double* A, i64 lda, int k1, int k2, const int* ipiv, int incx) {
    laswp_impl(s, n, A, lda, k1, k2, ipiv, incx, rows);
}

// ---- pieces of the row interchanges for the P x Q distributed LU (lapack_b200/dist2d.py), where the rows of one panel's
// interchanges live on different GPUs.  compose: the np transpositions (pivots relative to the block, 1-based, ip >= i as
// DGETRF produces them) as two maps over the np block rows: src_top[t] = original relative row whose content ends in block
// row t;  inv_top[t] = relative row where the original block row t ends (0-based, relative to the first block row).
__global__ void __launch_bounds__(256) laswp_compose_kernel(int np, const int* __restrict__ ipiv, int* __restrict__ src_top,
                                                            int* __restrict__ inv_top) {
    extern __shared__ int sm[];
    int* cur = sm;                    // [np]
    int* hkey = sm + np;              // [HASH] relative row + 1, 0 = empty
    int* hval = hkey + LASWP_HASH;
    const int tid = threadIdx.x;
    for (int t = tid; t < np; t += 256) { cur[t] = t; inv_top[t] = -1; }
    for (int q = tid; q < LASWP_HASH; q += 256) hkey[q] = 0;
    __syncthreads();
    if (tid == 0) {
        for (int t = 0; t < np; ++t) {
            const int ip = ipiv[t] - 1;          // relative, 0-based
            if (ip == t) continue;
            const int a = cur[t];
            if (ip < np) { cur[t] = cur[ip]; cur[ip] = a; }
            else {
                int h = swp_hash(ip);
                while (hkey[h] != 0 && hkey[h] != ip + 1) h = (h + 1) & (LASWP_HASH - 1);
                if (hkey[h] == 0) { hkey[h] = ip + 1; hval[h] = ip; }
                cur[t] = hval[h];
                hval[h] = a;
            }
        }
    }
    __syncthreads();
    for (int t = tid; t < np; t += 256) {
        const int o = cur[t];
        src_top[t] = o;
        if (o < np) inv_top[o] = t;
    }
    __syncthreads();
    for (int q = tid; q < LASWP_HASH; q += 256)
        if (hkey[q] != 0 && hval[q] < np) inv_top[hval[q]] = hkey[q] - 1;
}
void laswp_compose(cudaStream_t s, int np, const int* ipiv, int* src_top, int* inv_top) {
    if (np <= 0) return;
    laswp_attr();
    static bool attr = false;
    if (!attr) {
        LB_CUDA_CHECK(cudaFuncSetAttribute(laswp_compose_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (LASWP_MAX_PIV + 2 * LASWP_HASH) * 4));
        attr = true;
    }
    laswp_compose_kernel<<<1, 256, (size_t)(np + 2 * LASWP_HASH) * sizeof(int), s>>>(np, ipiv, src_top, inv_top);
    count_launch();
    LB_CUDA_CHECK(cudaGetLastError());
}
// W(t, c) = A(idx[t], c) for idx[t] >= 0 (W column-major nidx x ncols, ldw); rows with idx < 0 are left untouched
__global__ void gather_rows_kernel(int nidx, const int* __restrict__ idx, const double* __restrict__ A, i64 lda, int ncols,
                                   double* __restrict__ W, i64 ldw) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nidx) return;
    const int r = idx[t];
    if (r < 0) return;
    for (int c = blockIdx.y * 8; c < ncols; c += gridDim.y * 8) {
        double v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = (c + u < ncols) ? A[r + (i64)(c + u) * lda] : 0.0;
#pragma unroll
        for (int u = 0; u < 8; ++u) if (c + u < ncols) W[t + (i64)(c + u) * ldw] = v[u];
    }
}
__global__ void scatter_rows_kernel(int nidx, const int* __restrict__ idx, const double* __restrict__ W, i64 ldw, int ncols,
                                    double* __restrict__ A, i64 lda) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nidx) return;
    const int r = idx[t];
    if (r < 0) return;
    for (int c = blockIdx.y * 8; c < ncols; c += gridDim.y * 8) {
        double v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = (c + u < ncols) ? W[t + (i64)(c + u) * ldw] : 0.0;
#pragma unroll
        for (int u = 0; u < 8; ++u) if (c + u < ncols) A[r + (i64)(c + u) * lda] = v[u];
    }
}
void gather_rows(cudaStream_t s, int nidx, const int* idx, const double* A, i64 lda, int ncols, double* W, i64 ldw) {
    if (nidx <= 0 || ncols <= 0) return;
    dim3 grid(ceil_div(nidx, 128), (unsigned)min(ceil_div(ncols, 8), 16384));
    gather_rows_kernel<<<grid, 128, 0, s>>>(nidx, idx, A, lda, ncols, W, ldw);
    count_launch();
}
void scatter_rows(cudaStream_t s, int nidx, const int* idx, const double* W, i64 ldw, int ncols, double* A, i64 lda) {
    if (nidx <= 0 || ncols <= 0) return;
    dim3 grid(ceil_div(nidx, 128), (unsigned)min(ceil_div(ncols, 8), 16384));
    scatter_rows_kernel<<<grid, 128, 0, s>>>(nidx, idx, W, ldw, ncols, A, lda);
    count_launch();
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
// A(i,j) = draw number (offset + j*stream_ld + i) of the stream: an m x n window of a global column-major DLARNV matrix with
// stream_ld rows (2D block-cyclic pieces of ONE global matrix)
__global__ void larnv_submatrix_kernel(unsigned long long seed, i64 offset, i64 stream_ld, int m, int n, double* __restrict__ A, i64 lda) {
    const unsigned long long MASK = (1ULL << 48) - 1, AMUL = 33952834046453ULL;
    const int chunks = (m + LARNV_CHUNK - 1) / LARNV_CHUNK;
    const i64 t = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (i64)chunks * n) return;
    const int j = (int)(t / chunks), i0 = (int)(t - (i64)j * chunks) * LARNV_CHUNK;
    unsigned long long st = (seed * lcg_pow((unsigned long long)(offset + (i64)j * stream_ld + i0))) & MASK;
#pragma unroll
    for (int q = 0; q < LARNV_CHUNK; ++q) {
        if (i0 + q >= m) break;
        st = (st * AMUL) & MASK;
        A[i0 + q + (i64)j * lda] = 2.0 * ((double)st * (1.0 / 281474976710656.0)) - 1.0;
    }
}
void larnv_submatrix(cudaStream_t s, const int iseed[4], i64 stream_offset, i64 stream_ld, int m, int n, double* A, i64 lda) {
    if (m <= 0 || n <= 0) return;
    unsigned long long seed = ((unsigned long long)iseed[0] << 36) | ((unsigned long long)iseed[1] << 24) |
                              ((unsigned long long)iseed[2] << 12) | (unsigned long long)iseed[3];
    const i64 threads = (i64)((m + LARNV_CHUNK - 1) / LARNV_CHUNK) * n;
    larnv_submatrix_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, s>>>(seed, stream_offset, stream_ld, m, n, A, lda);
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
