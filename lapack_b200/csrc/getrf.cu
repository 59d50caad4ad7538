// getrf.cu -- LU with partial pivoting on one B200: DGETRF / DGETRF2 / DGETRS.
//
// Reference path: SRC/dgetrf.f:164-219 (right-looking blocked driver), SRC/dgetrf2.f:170-265 (recursive
// panel; leaf = IDAMAX + swap + scale, BLAS/SRC/idamax.f:93-123, dscal.f:104-135), SRC/dlaswp.f:138-183,
// SRC/dgetrs.f:181-218.
//
// B200 design.
//  * Outer block NB (default 512, not ILAENV's 64): the trailing update C -= L21*U12 then has arithmetic
//    intensity NB/8 flop/B, far above the HBM ridge, and runs as DMMA GEMM tiles (gemm_f64.cu).
//  * The NB-wide panel is factored recursively (same splitting idea as DGETRF2) so that all but the
//    narrowest level is again TRSM/GEMM on the tensor pipe.  The recursion stops at a W(=16)-column leaf.
//  * Leaf kernel (getrf_leaf_kernel): cooperative multi-CTA kernel, one panel row per thread held in
//    registers for all W columns.  Per column: block-level arg-max (warp shuffles), candidates + their
//    rows published to global scratch, ONE grid barrier, every CTA then picks the winner (first index on
//    ties, IDAMAX semantics), swaps through the published copies, scales by the reciprocal (or divides
//    when |pivot| < SFMIN, dgetrf2.f:204-210) and applies the rank-1 update to its rows.  Memory-bound:
//    the panel is read once and written once; algorithmic bytes 16*M*W per leaf.
//  * Look-ahead: panel k+1 is factored on a high-priority stream as soon as its columns have received
//    update k, while the rest of update k runs on another stream.
#include "lb_internal.h"
#include <cfloat>
#include <math_constants.h>
#include <mutex>
#include <vector>

namespace lb {

void laswp_rows(cudaStream_t s, int n, int rows, double* A, i64 lda, int k1, int k2, const int* ipiv, int incx);
void* laswp_plan(cudaStream_t s, int k1, int k2, const int* ipiv, int incx);
void laswp_apply_plan(cudaStream_t s, int n, double* A, i64 lda, const void* plan, int npiv);
void laswp_plan_free(cudaStream_t s, void* plan);
bool laswp_apply_chain(cudaStream_t s, int m, int nb, int nplans, void* const* plans_host, double* A, i64 lda);

static int g_nb = 512, g_lookahead = 1;
static int g_cluster_max = 16;      // panels of up to this many 1024-row CTAs use the cluster leaf (0 = never)
int getrf_block() { return min(g_nb, 2048); }
void getrf_set_cluster_max(int c) { g_cluster_max = c < 0 ? 0 : (c > 16 ? 16 : c); }
static int g_big_leaf_rows4 = 1;    // panels too tall for a cluster: 1 = 256 threads x 4 rows kernel, 0 = 1024 x 1 row kernel
void getrf_set_big_leaf(int v) { g_big_leaf_rows4 = v ? 1 : 0; }
static int g_cluster_fat = 0;       // 1: panels of 16385..32768 rows use ONE cluster of <= 16 CTAs with 8 rows per thread (2048 rows per CTA)
void getrf_set_cluster_fat(int v) { g_cluster_fat = v ? 1 : 0; }
static int g_tall_rows = 1024;      // rows per CTA of the global-packet leaf for panels too tall for one cluster: 1024 / 2048 / 4096
void getrf_set_tall_rows(int r) { g_tall_rows = (r == 2048 || r == 4096) ? r : 1024; }
// Thin leaves (see getrf_leaf_cluster_kernel, MINB): 0 = off, 1 = 128 threads x 2 rows, 2 = 64 threads x 4 rows (256 rows per CTA either
// way, one GEMM-CTA slot each); used for panels of more than g_thin_min_rows rows, i.e. while the panel is hidden behind the update.
static int g_defer_left = 1, g_defer_tail_rows = 512;     // deferred interchanges left of the panel: 0 off, 1 composed streaming pass, 2 plan by plan
void getrf_set_defer_left(int on, int tail_rows) { g_defer_left = on; if (tail_rows > 0) g_defer_tail_rows = tail_rows; }
static int g_thin_mode = 0, g_thin_min_rows = 16384;
void getrf_set_thin(int mode, int min_rows) { g_thin_mode = mode; if (min_rows >= 0) g_thin_min_rows = min_rows; }
void getrf_set_params(int nb, int leaf, int lookahead) {
    (void)leaf;
    if (nb > 0) g_nb = nb;
    if (lookahead >= 0) g_lookahead = lookahead;
}

// ------------------------------------------------------------------------------------------------
// Leaf kernel communication.  No central barrier counter: every CTA publishes, per column step, a tagged
// packet {tag = epoch<<32 | row, key, the 16 row values}; the tag is stored last with st.release and every
// CTA polls all G tags with ld.acquire, so "barrier" and "data exchange" cost one L2 round trip together.
// Packets are double-buffered on the step parity; a CTA can only overwrite slot (c&1) at step c+2 after it
// has seen every step-(c+1) packet, i.e. after every other CTA finished reading step c.
struct LeafPacket {
    unsigned long long tag;
    double key;
    double rowdata[16];
    double pad[14];                  // 256-byte stride
};
struct LeafParams {
    int m, n;
    double* A;
    i64 lda;
    int* ipiv;        // panel-relative, 1-based on output
    int* info;        // device word: first zero pivot (absolute, 1-based) -- written only if still 0
    int info_off;     // absolute column offset of this leaf
    int piv_base;     // added to every pivot index written (row offset of the leaf inside the caller's matrix)
    double sfmin;
    unsigned epoch_base;
    LeafPacket* cand;       // [2][G]
    LeafPacket* top;        // [2]     current row c (published by its owner)
    unsigned long long* hist;   // [W]  winner of every step, tagged, for the swap CTAs
    int G;            // work CTAs; CTAs >= G apply the interchanges to the other panel columns
    double* SW;       // element (leaf top row, panel column 0)
    int sw_left, sw_right;   // panel columns left / right of the leaf
};

__device__ __forceinline__ bool cand_better(double k1, int r1, double k2, int r2) {
    return (k1 > k2) || (k1 == k2 && r1 < r2);
}
__device__ __forceinline__ void st_release_u64(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.gpu.global.u64 [%0], %1;\n" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];\n" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long poll_tag(const unsigned long long* p, unsigned epoch) {
    unsigned long long v;
    do { v = ld_acquire_u64(p); } while ((unsigned)(v >> 32) != epoch);
    return v;
}

// The column loop is NOT unrolled: after step c every thread stores its column-c value (final L or U entry)
// and rotates its register window one column to the left, so that the active column is always a[0].  The loop
// body therefore exists once in the instruction stream.
template <int W, int THREADS>
__global__ void __launch_bounds__(THREADS, 1) getrf_leaf_kernel(LeafParams p) {
    constexpr int NWARP = THREADS / 32;
    __shared__ double s_key[NWARP];
    __shared__ int s_row[NWARP];
    __shared__ double s_prow[W], s_trow[W];
    __shared__ int s_win_row;
    __shared__ int s_loc_row;
    __shared__ double s_loc_key;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = blockIdx.x;
    const int kmax = min(p.m, p.n);

    if (g >= p.G) {
        // ===== interchange CTAs: one panel column per thread, follow the winner history (dgetrf2.f:236,263).
        // Columns of the leaf itself are included once they are final (column q < c has been stored by the work
        // CTAs before their step-c packets, which CTA 0 acquired before releasing hist[c]).
        const int t = (g - p.G) * THREADS + tid;
        const int width = p.sw_left + p.n + p.sw_right;
        const bool valid = t < width;
        const int own = t - p.sw_left;                        // index inside the leaf, if 0 <= own < n
        double* colp = p.SW + (i64)t * p.lda;
#pragma unroll 1
        for (int c = 0; c < kmax; ++c) {
            unsigned long long v = 0;
            if (lane == 0) v = poll_tag(p.hist + c, p.epoch_base + c + 1);
            v = __shfl_sync(0xffffffffu, v, 0);
            const int prow = (int)(unsigned)(v & 0xffffffffu);
            const bool mine = valid && (own < 0 || own >= p.n || own < c);
            if (mine && prow != c) {
                double x = colp[c], y = colp[prow];
                colp[c] = y;
                colp[prow] = x;
            }
        }
        return;
    }

    const int row = g * THREADS + tid;          // panel-relative row owned by this thread
    const bool have = row < p.m;
    double a[W];                                 // a[q] = A(row, c + q): window starting at the active column c
#pragma unroll
    for (int q = 0; q < W; ++q) a[q] = (have && q < p.n) ? p.A[row + (i64)q * p.lda] : 0.0;

#pragma unroll 1
    for (int c = 0; c < kmax; ++c) {
        const int slot = c & 1;
        const unsigned epoch = p.epoch_base + c + 1;
        // (0) the owner of row c publishes the current top row
        if (have && row == c) {
            LeafPacket* tp = p.top + slot;
#pragma unroll
            for (int q = 0; q < W; ++q) tp->rowdata[q] = a[q];
            st_release_u64(&tp->tag, ((unsigned long long)epoch << 32));
        }
        // (1) local arg-max over active rows (row >= c); IDAMAX semantics: first index of the max,
        //     NaN never wins unless it is the very first element (idamax.f:103 strict '>').
        double key = -1.0;
        int krow = 0x7fffffff;
        if (have && row >= c) {
            double v = fabs(a[0]);
            if (v != v) v = (row == c) ? CUDART_INF : -1.0;
            key = v;
            krow = row;
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            double ok = __shfl_xor_sync(0xffffffffu, key, off);
            int orow = __shfl_xor_sync(0xffffffffu, krow, off);
            if (cand_better(ok, orow, key, krow)) { key = ok; krow = orow; }
        }
        if (lane == 0) { s_key[warp] = key; s_row[warp] = krow; }
        __syncthreads();
        if (warp == 0) {
            double k2 = (lane < NWARP) ? s_key[lane] : -1.0;
            int r2 = (lane < NWARP) ? s_row[lane] : 0x7fffffff;
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                double ok = __shfl_xor_sync(0xffffffffu, k2, off);
                int orow = __shfl_xor_sync(0xffffffffu, r2, off);
                if (cand_better(ok, orow, k2, r2)) { k2 = ok; r2 = orow; }
            }
            if (lane == 0) { s_loc_row = r2; s_loc_key = k2; }
        }
        __syncthreads();
        // (2) the thread owning the CTA's best row publishes it (data first, tag last with release)
        if (have && row == s_loc_row) {
            LeafPacket* cp = p.cand + slot * p.G + g;
            cp->key = s_loc_key;
#pragma unroll
            for (int q = 0; q < W; ++q) cp->rowdata[q] = a[q];
            st_release_u64(&cp->tag, ((unsigned long long)epoch << 32) | (unsigned)row);
        }
        // (3)+(4) warp 0 waits for all G packets, picks the winner and fetches the two rows
        if (warp == 0) {
            double k2 = -2.0;
            int r2 = 0x7fffffff, g2 = 0;
            for (int q = lane; q < p.G; q += 32) {
                const LeafPacket* cp = p.cand + slot * p.G + q;
                unsigned long long tag = poll_tag(&cp->tag, epoch);
                double ck = __ldcg(&cp->key);
                int cr = (int)(unsigned)(tag & 0xffffffffu);
                if (cand_better(ck, cr, k2, r2)) { k2 = ck; r2 = cr; g2 = q; }
            }
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                double ok = __shfl_xor_sync(0xffffffffu, k2, off);
                int orow = __shfl_xor_sync(0xffffffffu, r2, off);
                int og = __shfl_xor_sync(0xffffffffu, g2, off);
                if (cand_better(ok, orow, k2, r2)) { k2 = ok; r2 = orow; g2 = og; }
            }
            if (lane < W) {
                (void)poll_tag(&p.cand[slot * p.G + g2].tag, epoch);        // own acquire for the data below
                (void)poll_tag(&p.top[slot].tag, epoch);
                s_prow[lane] = __ldcg(&p.cand[slot * p.G + g2].rowdata[lane]);
                s_trow[lane] = __ldcg(&p.top[slot].rowdata[lane]);
            }
            if (lane == 0) {
                s_win_row = r2;
                if (g == 0) st_release_u64(p.hist + c, ((unsigned long long)epoch << 32) | (unsigned)r2);
            }
        }
        __syncthreads();
        const int prow = s_win_row;
        const double pivot = s_prow[0];
        // (5) interchange through the published copies (dgetrf2.f:196-200); columns left of c were already
        //     stored and are exchanged in global memory by the interchange CTAs
        if (prow != c) {
            if (have && row == prow) {
#pragma unroll
                for (int q = 0; q < W; ++q) a[q] = s_trow[q];
            } else if (have && row == c) {
#pragma unroll
                for (int q = 0; q < W; ++q) a[q] = s_prow[q];
            }
        }
        // (6) pivot index / singularity flag (dgetrf2.f:191-192,212-214)
        if (g == 0 && tid == 0) {
            p.ipiv[c] = prow + 1 + p.piv_base;
            if (pivot == 0.0 && *p.info == 0) *p.info = p.info_off + c + 1;
        }
        // (7) scale (reciprocal unless |pivot| < SFMIN, dgetrf2.f:204-210) and rank-1 update
        if (pivot != 0.0 && have && row > c) {
            double l;
            if (fabs(pivot) >= p.sfmin) l = a[0] * (1.0 / pivot);
            else l = a[0] / pivot;
            a[0] = l;
#pragma unroll
            for (int q = 1; q < W; ++q) a[q] = fma(-l, s_prow[q], a[q]);
        }
        // (8) column c is final for every row: store it and slide the window
        if (have) p.A[row + (i64)c * p.lda] = a[0];
#pragma unroll
        for (int q = 0; q + 1 < W; ++q) a[q] = a[q + 1];
        a[W - 1] = 0.0;
    }
    // wide leaf (m < n): the columns right of the last pivot are still in the window
    if (have) {
#pragma unroll
        for (int q = 0; q < W; ++q)
            if (kmax + q < p.n) p.A[row + (i64)(kmax + q) * p.lda] = a[q];
    }
}

// ------------------------------------------------------------------------------------------------
// Cluster leaf (panels of at most CL_MAX_CTAS*1024 rows): the work CTAs form ONE thread-block cluster, so the
// per-column exchange goes through distributed shared memory and the hardware cluster barrier instead of L2
// round trips.  Per column: CTA-local arg-max, the best row and (in CTA 0) the current top row are written
// to the CTA's own shared memory, barrier.cluster, warp 0 of every CTA reads the C candidate heads remotely,
// picks the winner and pulls the two rows; one __syncthreads later every thread swaps/scales/updates.
// All W columns stay in registers until the end (the window rotates cyclically), so an interchange moves whole
// leaf rows and only the panel columns OUTSIDE the leaf are left to the interchange CTAs (second cluster).
struct ClusterCand {
    double key;
    int row;
    int pad;
    double rowdata[16];
};
__device__ __forceinline__ void cluster_arrive_release() { asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory"); }
__device__ __forceinline__ void cluster_wait_acquire() { asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory"); }
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ unsigned mapa_u32(unsigned addr, unsigned rank) {
    unsigned r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;\n" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ double ld_dsmem_f64(unsigned addr) {
    double v;
    asm volatile("ld.shared::cluster.f64 %0, [%1];\n" : "=d"(v) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ int ld_dsmem_s32(unsigned addr) {
    int v;
    asm volatile("ld.shared::cluster.s32 %0, [%1];\n" : "=r"(v) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_u64(unsigned long long* p, unsigned long long v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;\n" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_relaxed_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];\n" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// Warp arg-max of (key, row) with IDAMAX tie-breaking.  key >= 0 for candidates (its high word then orders like
// an unsigned integer), key < 0 = no candidate.  Fast path: one redux.sync on the high words and a ballot; only
// when several lanes agree in the high word does the full butterfly run.
__device__ __forceinline__ void warp_argmax(double& key, int& krow) {
    const unsigned full = 0xffffffffu;
    const bool cand = key >= 0.0;
    const unsigned hi = cand ? (unsigned)__double2hiint(key) + 1u : 0u;      // +1: a candidate 0.0 still beats "none"
    const unsigned mh = __reduce_max_sync(full, hi);
    const unsigned tie = __ballot_sync(full, hi == mh);
    if (__popc(tie) == 1) {
        const int src = __ffs(tie) - 1;
        key = __shfl_sync(full, key, src);
        krow = __shfl_sync(full, krow, src);
        return;
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        double ok = __shfl_xor_sync(full, key, off);
        int orow = __shfl_xor_sync(full, krow, off);
        if (cand_better(ok, orow, key, krow)) { key = ok; krow = orow; }
    }
}

// R rows per thread (row = g*R*THREADS + r*THREADS + tid): the per-step bookkeeping (reductions, barriers, the
// pull of the pivot row) is paid once per thread, so few fat threads beat many thin ones -- the kernel is bound
// by instruction issue, not by latency.
// MINB > 1 ("thin" leaves): the CTA is sized to fit into the resources ONE trailing-update GEMM CTA gives back (128 threads x 128
// registers, or 64 x 255), so a look-ahead panel launched next to a running update is placed as GEMM CTAs retire instead of waiting
// for whole SMs to drain (the block scheduler holds back every lower-priority CTA while a higher-priority one is pending).
template <int W, int THREADS, int R, bool CLUSTER, int MINB = 1>
__global__ void __launch_bounds__(THREADS, MINB) getrf_leaf_cluster_kernel(LeafParams p) {
    constexpr int NWARP = THREADS / 32;
    constexpr int ROWS = THREADS * R;
    __shared__ double s_key[NWARP];
    __shared__ int s_row[NWARP];
    __shared__ __align__(16) ClusterCand s_cand[2];
    __shared__ __align__(16) double s_top[2][W];
    __shared__ __align__(16) double s_prow[W];
    __shared__ __align__(16) double s_trow[W];
    __shared__ int s_win_row;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = blockIdx.x;
    const int C = p.G;                                        // cluster size == number of work CTAs
    const int kmax = min(p.m, p.n);

    if (g >= C) {
        // ===== interchange CTAs (second cluster): the panel columns outside the leaf (dgetrf2.f:236,263)
        const int width = p.sw_left + p.n + p.sw_right;
        if ((g - C) * THREADS >= width) return;
        const int t = (g - C) * THREADS + tid;
        const int own = t - p.sw_left;
        const bool mine = t < width && (own < 0 || own >= p.n);
        double* colp = p.SW + (i64)t * p.lda;
#pragma unroll 1
        for (int c = 0; c < kmax; ++c) {
            unsigned long long v = 0;
            if (lane == 0) {
                const unsigned epoch = p.epoch_base + c + 1;
                do { v = ld_relaxed_u64(p.hist + c); } while ((unsigned)(v >> 32) != epoch);
            }
            v = __shfl_sync(0xffffffffu, v, 0);
            const int prow = (int)(unsigned)(v & 0xffffffffu);
            if (mine && prow != c) {
                double x = colp[c], y = colp[prow];
                colp[c] = y;
                colp[prow] = x;
            }
        }
        return;
    }

    const int row0 = g * ROWS + tid;             // rows row0 + r*THREADS
    double a[R][W];                              // a[r][q] = A(row_r, (c + q) mod W) at step c
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const int row = row0 + r * THREADS;
#pragma unroll
        for (int q = 0; q < W; ++q) a[r][q] = (row < p.m && q < p.n) ? p.A[row + (i64)q * p.lda] : 0.0;
    }
    const unsigned cand_base = smem_u32(&s_cand[0]);
    const unsigned top_base = smem_u32(&s_top[0][0]);

#pragma unroll 1
    for (int c = 0; c < kmax; ++c) {
        const int slot = c & 1;
        // (1) CTA-local arg-max over the active rows, IDAMAX semantics (idamax.f:95-106)
        double key = -1.0;
        int krow = 0x7fffffff;
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int row = row0 + r * THREADS;
            if (row < p.m && row >= c) {
                double v = fabs(a[r][0]);
                if (v != v) v = (row == c) ? CUDART_INF : -1.0;
                if (v > key) { key = v; krow = row; }          // rows ascend with r: strict '>' keeps the first
            }
        }
        if (key < 0.0) krow = 0x7fffffff;
        warp_argmax(key, krow);
        if (lane == 0) { s_key[warp] = key; s_row[warp] = krow; }
        __syncthreads();
        {   // every warp reduces the NWARP partial results itself (no second CTA barrier)
            double k2 = (lane < NWARP) ? s_key[lane] : -1.0;
            int r2 = (lane < NWARP) ? s_row[lane] : 0x7fffffff;
#pragma unroll
            for (int off = NWARP / 2; off > 0; off >>= 1) {
                double ok = __shfl_xor_sync(0xffffffffu, k2, off);
                int orow = __shfl_xor_sync(0xffffffffu, r2, off);
                if (cand_better(ok, orow, k2, r2)) { k2 = ok; r2 = orow; }
            }
            key = __shfl_sync(0xffffffffu, k2, 0);
            krow = __shfl_sync(0xffffffffu, r2, 0);
        }
        if (CLUSTER) {
        // (2) publish the CTA's best row (and the top row) in this CTA's shared memory
        if (krow == 0x7fffffff) {
            if (tid == 0) { s_cand[slot].key = -2.0; s_cand[slot].row = 0x7fffffff; }     // no active row here
        } else if (((krow - row0) % THREADS) == 0 && krow >= row0 && krow < row0 + R * THREADS) {
#pragma unroll
            for (int r = 0; r < R; ++r)
                if (krow == row0 + r * THREADS) {
                    s_cand[slot].key = key;
                    s_cand[slot].row = krow;
#pragma unroll
                    for (int q = 0; q < W; ++q) s_cand[slot].rowdata[q] = a[r][q];
                }
        }
        if (row0 == c) {                                       // row c always lives in CTA 0, r = 0
#pragma unroll
            for (int q = 0; q < W; ++q) s_top[slot][q] = a[0][q];
        }
        cluster_arrive_release();
        cluster_wait_acquire();
        // (3) warp 0: winner over the C candidates, then pull the winner's row and the top row
        if (warp == 0) {
            double k2 = -3.0;
            int r2 = 0x7fffffff, g2 = 0;
            if (lane < C) {
                const unsigned ra = mapa_u32(cand_base + slot * (unsigned)sizeof(ClusterCand), (unsigned)lane);
                k2 = ld_dsmem_f64(ra);
                r2 = ld_dsmem_s32(ra + 8);
                g2 = lane;
            }
#pragma unroll
            for (int off = 8; off > 0; off >>= 1) {
                double ok = __shfl_xor_sync(0xffffffffu, k2, off);
                int orow = __shfl_xor_sync(0xffffffffu, r2, off);
                int og = __shfl_xor_sync(0xffffffffu, g2, off);
                if (cand_better(ok, orow, k2, r2)) { k2 = ok; r2 = orow; g2 = og; }
            }
            if (lane < W) {
                const unsigned ra = mapa_u32(cand_base + slot * (unsigned)sizeof(ClusterCand) + 16 + lane * 8, (unsigned)g2);
                s_prow[lane] = ld_dsmem_f64(ra);
            } else if (lane < 2 * W) {
                const unsigned ta = mapa_u32(top_base + (slot * W + (lane - W)) * 8, 0u);
                s_trow[lane - W] = ld_dsmem_f64(ta);
            }
            if (lane == 0) s_win_row = r2;
        }
        } else {
        // (2') no cluster (taller panels): tagged packets in global memory, as in getrf_leaf_kernel -- data first, the
        //      tag {epoch, row} last with st.release; every CTA's warp 0 polls all G tags with ld.acquire
        const unsigned epoch = p.epoch_base + c + 1;
        if (krow == 0x7fffffff) {
            if (tid == 0) {
                LeafPacket* cp = p.cand + slot * p.G + g;
                cp->key = -2.0;
                st_release_u64(&cp->tag, ((unsigned long long)epoch << 32) | 0x7fffffffu);
            }
        } else if (((krow - row0) % THREADS) == 0 && krow >= row0 && krow < row0 + R * THREADS) {
#pragma unroll
            for (int r = 0; r < R; ++r)
                if (krow == row0 + r * THREADS) {
                    LeafPacket* cp = p.cand + slot * p.G + g;
                    cp->key = key;
#pragma unroll
                    for (int q = 0; q < W; ++q) cp->rowdata[q] = a[r][q];
                    st_release_u64(&cp->tag, ((unsigned long long)epoch << 32) | (unsigned)krow);
                }
        }
        if (row0 == c) {
            LeafPacket* tp = p.top + slot;
#pragma unroll
            for (int q = 0; q < W; ++q) tp->rowdata[q] = a[0][q];
            st_release_u64(&tp->tag, ((unsigned long long)epoch << 32));
        }
        if (warp == 0) {
            double k2 = -3.0;
            int r2 = 0x7fffffff, g2 = 0;
            // up to four packets per lane are polled together (their loads overlap): thin leaves have up to 256 work CTAs
            for (int q0 = 0; q0 < p.G; q0 += 128) {
                unsigned long long tg[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int q = q0 + lane + 32 * i;
                    tg[i] = (q < p.G) ? ld_acquire_u64(&p.cand[slot * p.G + q].tag) : 0ull;
                }
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int q = q0 + lane + 32 * i;
                    if (q < p.G)
                        while ((unsigned)(tg[i] >> 32) != epoch) tg[i] = ld_acquire_u64(&p.cand[slot * p.G + q].tag);
                }
                double ck[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int q = q0 + lane + 32 * i;
                    ck[i] = (q < p.G) ? __ldcg(&p.cand[slot * p.G + q].key) : -3.0;
                }
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int q = q0 + lane + 32 * i;
                    const int cr = (int)(unsigned)(tg[i] & 0xffffffffu);
                    if (q < p.G && cand_better(ck[i], cr, k2, r2)) { k2 = ck[i]; r2 = cr; g2 = q; }
                }
            }
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                double ok = __shfl_xor_sync(0xffffffffu, k2, off);
                int orow = __shfl_xor_sync(0xffffffffu, r2, off);
                int og = __shfl_xor_sync(0xffffffffu, g2, off);
                if (cand_better(ok, orow, k2, r2)) { k2 = ok; r2 = orow; g2 = og; }
            }
            if (lane < W) {
                (void)poll_tag(&p.cand[slot * p.G + g2].tag, epoch);        // own acquire for the data below
                s_prow[lane] = __ldcg(&p.cand[slot * p.G + g2].rowdata[lane]);
            } else if (lane < 2 * W) {
                (void)poll_tag(&p.top[slot].tag, epoch);
                s_trow[lane - W] = __ldcg(&p.top[slot].rowdata[lane - W]);
            }
            if (lane == 0) s_win_row = r2;
        }
        }
        __syncthreads();
        const int prow = s_win_row;
        double pr[W];
#pragma unroll
        for (int q = 0; q < W; q += 2) {
            const double2 v = *reinterpret_cast<const double2*>(&s_prow[q]);
            pr[q] = v.x; pr[q + 1] = v.y;
        }
        const double pivot = pr[0];
        // (4) interchange whole leaf rows through the published copies (dgetrf2.f:196-200)
        if (prow != c) {
            if (((prow - row0) % THREADS) == 0 && prow >= row0 && prow < row0 + R * THREADS) {
#pragma unroll
                for (int r = 0; r < R; ++r)
                    if (prow == row0 + r * THREADS) {
#pragma unroll
                        for (int q = 0; q < W; ++q) a[r][q] = s_trow[q];
                    }
            }
            if (row0 == c) {
#pragma unroll
                for (int q = 0; q < W; ++q) a[0][q] = pr[q];
            }
        }
        // (5) pivot index / singularity flag (dgetrf2.f:191-192,212-214); winner history for the interchange CTAs
        if (g == 0 && tid == 32) {
            p.ipiv[c] = prow + 1 + p.piv_base;
            if (pivot == 0.0 && *p.info == 0) *p.info = p.info_off + c + 1;
            st_relaxed_u64(p.hist + c, ((unsigned long long)(p.epoch_base + c + 1) << 32) | (unsigned)prow);
        }
        // (6) scale (reciprocal unless |pivot| < SFMIN, dgetrf2.f:204-210) and rank-1 update of the live columns
        if (pivot != 0.0) {
            const bool use_rcp = fabs(pivot) >= p.sfmin;
            const double rcp = 1.0 / pivot;
            const int live = W - c;                           // window positions 1..live-1 hold columns > c
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const int row = row0 + r * THREADS;
                if (row < p.m && row > c) {
                    const double l = use_rcp ? a[r][0] * rcp : a[r][0] / pivot;
                    a[r][0] = l;
#pragma unroll
                    for (int q = 1; q < W; ++q)
                        if (q < live) a[r][q] = fma(-l, pr[q], a[r][q]);
                }
            }
        }
        // (7) rotate the window: column c goes to the back
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const double t0 = a[r][0];
#pragma unroll
            for (int q = 0; q + 1 < W; ++q) a[r][q] = a[r][q + 1];
            a[r][W - 1] = t0;
        }
    }
    // window position q now holds column (q + kmax) mod W
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const int row = row0 + r * THREADS;
        if (row < p.m) {
#pragma unroll
            for (int q = 0; q < W; ++q) {
                const int col = (q + kmax) % W;
                if (col < p.n) p.A[row + (i64)col * p.lda] = a[r][q];
            }
        }
    }
    // nobody may exit while its shared memory can still be read by a peer
    if (CLUSTER) {
        cluster_arrive_release();
        cluster_wait_acquire();
    }
}

// ------------------------------------------------------------------------------------------------
constexpr int LEAF_W = 16;
constexpr int LEAF_THREADS = 1024;
constexpr int CL_THREADS = 256, CL_R = 4;     // cluster leaf: 256 threads x 4 rows = 1024 rows per CTA
static_assert(CL_THREADS * CL_R == LEAF_THREADS, "both leaf kernels cover 1024 rows per CTA");

struct LeafWs {
    unsigned epoch = 0;
    LeafPacket* cand = nullptr;
    LeafPacket* top = nullptr;
    unsigned long long* hist = nullptr;
    int maxG = 0;
};
static LeafWs& leaf_ws() {
    static LeafWs w;
    if (!w.cand) {
        w.maxG = 256;
        LB_CUDA_CHECK(cudaMalloc(&w.cand, sizeof(LeafPacket) * 2 * w.maxG));
        LB_CUDA_CHECK(cudaMemset(w.cand, 0, sizeof(LeafPacket) * 2 * w.maxG));
        LB_CUDA_CHECK(cudaMalloc(&w.top, sizeof(LeafPacket) * 2));
        LB_CUDA_CHECK(cudaMemset(w.top, 0, sizeof(LeafPacket) * 2));
        LB_CUDA_CHECK(cudaMalloc(&w.hist, sizeof(unsigned long long) * LEAF_W));
        LB_CUDA_CHECK(cudaMemset(w.hist, 0, sizeof(unsigned long long) * LEAF_W));

    }
    return w;
}

// The panel the leaves belong to: P = element (0,0) of the panel, width columns; a leaf at column offset
// `off` (== its row offset) interchanges rows in all the other panel columns as it goes.
struct PanelCtx {
    double* P;
    int width;
    int piv_base;     // row offset of the panel inside the caller's matrix (dgetrf.f:183-186 shift, done at the source)
};

// factor an m x n (n <= LEAF_W) leaf at panel offset `off`; ipiv relative (1-based); *info set to
// info_off + col + 1 on the first exact zero pivot (only if still zero)
// largest cluster this device/driver schedules for the cluster leaf (16 is a non-portable size)
static int cluster_hw_max() {
    static int hw_max = 0;
    if (!hw_max) {
        auto kern = getrf_leaf_cluster_kernel<LEAF_W, CL_THREADS, CL_R, true>;
        hw_max = 8;
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess) {
            cudaLaunchConfig_t q = {};
            q.gridDim = dim3(16); q.blockDim = dim3(CL_THREADS);
            cudaLaunchAttribute qa[1];
            qa[0].id = cudaLaunchAttributeClusterDimension;
            qa[0].val.clusterDim.x = 16; qa[0].val.clusterDim.y = 1; qa[0].val.clusterDim.z = 1;
            q.attrs = qa; q.numAttrs = 1;
            int nact = 0;
            if (cudaOccupancyMaxActiveClusters(&nact, kern, &q) == cudaSuccess && nact > 0) hw_max = 16;
        }
        (void)cudaGetLastError();
    }
    return hw_max;
}

static void getrf_leaf(cudaStream_t s, const PanelCtx& pc, int off, int m, int n, double* A, i64 lda, int* ipiv, int* info,
                       int info_off) {
    // pivots are written relative to the top of the matrix the panel belongs to: pc.piv_base + row offset of the leaf
    LeafWs& w = leaf_ws();
    LeafParams p;
    p.m = m; p.n = n; p.A = A; p.lda = lda; p.ipiv = ipiv; p.info = info; p.info_off = info_off;
    p.piv_base = pc.piv_base + off;
    p.sfmin = DBL_MIN;   // DLAMCH('S') (INSTALL/dlamch.f:111-122)
    p.G = ceil_div(m, LEAF_THREADS);
    p.epoch_base = w.epoch; p.cand = w.cand; p.top = w.top; p.hist = w.hist;
    p.SW = pc.P + off;
    p.sw_left = off;
    p.sw_right = pc.width - off - n;
    const int S = ceil_div(pc.width, LEAF_THREADS);
    if (g_thin_mode && m > g_thin_min_rows) {
        const int thr = (g_thin_mode == 1) ? 128 : 64;
        const int Gt = ceil_div(m, 256);
        const int St = (pc.width > n) ? ceil_div(pc.width, thr) : 0;
        if (Gt <= w.maxG && Gt + St <= 4 * num_sms() - 16) {
            p.G = Gt;
            if (g_thin_mode == 1) getrf_leaf_cluster_kernel<LEAF_W, 128, 2, false, 4><<<Gt + St, 128, 0, s>>>(p);
            else getrf_leaf_cluster_kernel<LEAF_W, 64, 4, false, 4><<<Gt + St, 64, 0, s>>>(p);
            count_launch();
            w.epoch += (unsigned)min(m, n);
            LB_CUDA_CHECK(cudaGetLastError());
            return;
        }
    }
    // All work CTAs of a leaf must be co-resident (they spin on each other's packets).  Panels taller than
    // (#SMs - interchange CTAs) x 1024 rows use 8 or 16 rows per thread -- slower (the row window spills to local
    // memory) but the same algorithm and results; 256 x 16 x 146 = 598,016 rows is the limit.
    {
        const int S4 = (pc.width > n) ? ceil_div(pc.width, CL_THREADS) : 0;
        const int cap = min(w.maxG, num_sms() - S4 - 2);
        const bool beyond_cluster = !(p.G <= g_cluster_max && p.G <= cluster_hw_max());
        if (p.G > cap || (beyond_cluster && g_tall_rows > 1024)) {
            int rows_per_cta = (p.G > cap) ? 2048 : g_tall_rows;
            if (ceil_div(m, rows_per_cta) > cap) rows_per_cta = 4096;
            p.G = ceil_div(m, rows_per_cta);
            if (p.G > cap) {
                fprintf(stderr, "lapack_b200: panel of %d rows exceeds the cooperative leaf capacity (%d rows)\n", m, cap * 4096);
                record_cuda_error(cudaErrorInvalidValue);
                return;
            }
            if (rows_per_cta == 2048) getrf_leaf_cluster_kernel<LEAF_W, CL_THREADS, 8, false><<<p.G + S4, CL_THREADS, 0, s>>>(p);
            else getrf_leaf_cluster_kernel<LEAF_W, CL_THREADS, 16, false><<<p.G + S4, CL_THREADS, 0, s>>>(p);
            count_launch();
            w.epoch += (unsigned)min(m, n);
            LB_CUDA_CHECK(cudaGetLastError());
            return;
        }
    }
    if (g_cluster_fat && p.G > min(g_cluster_max, cluster_hw_max()) && ceil_div(m, 2048) <= min(g_cluster_max, cluster_hw_max())) {
        // fat cluster leaf: 8 rows per thread, 2048 rows per CTA, DSMEM exchange -- half the SMs of the global-packet leaf
        auto kern = getrf_leaf_cluster_kernel<LEAF_W, CL_THREADS, 8, true>;
        static bool attr = false;
        if (!attr) { (void)cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1); (void)cudaGetLastError(); attr = true; }
        p.G = ceil_div(m, 2048);
        const bool outside = pc.width > n;
        int nclusters = 1;
        if (outside) nclusters += ceil_div(ceil_div(pc.width, CL_THREADS), p.G);
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)(p.G * nclusters));
        cfg.blockDim = dim3(CL_THREADS);
        cfg.dynamicSmemBytes = 0;
        cfg.stream = s;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = (unsigned)p.G; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        LB_CUDA_CHECK(cudaLaunchKernelEx(&cfg, kern, p));
        count_launch();
        w.epoch += (unsigned)min(m, n);
        return;
    }
    if (p.G <= g_cluster_max && p.G <= cluster_hw_max()) {
        // one cluster of G work CTAs (+ clusters of interchange CTAs when the panel is wider than the leaf)
        auto kern = getrf_leaf_cluster_kernel<LEAF_W, CL_THREADS, CL_R, true>;
        const bool outside = pc.width > n;
        int nclusters = 1;
        if (outside) nclusters += ceil_div(ceil_div(pc.width, CL_THREADS), p.G);
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)(p.G * nclusters));
        cfg.blockDim = dim3(CL_THREADS);
        cfg.dynamicSmemBytes = 0;
        cfg.stream = s;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = (unsigned)p.G; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        LB_CUDA_CHECK(cudaLaunchKernelEx(&cfg, kern, p));
        count_launch();
        w.epoch += (unsigned)min(m, n);
        return;
    }
    if (g_big_leaf_rows4) {
        // taller panels: same 256 x 4-row kernel, exchange through global packets (all CTAs must be co-resident)
        const int S4 = ceil_div(pc.width, CL_THREADS);
        getrf_leaf_cluster_kernel<LEAF_W, CL_THREADS, CL_R, false><<<p.G + (pc.width > n ? S4 : 0), CL_THREADS, 0, s>>>(p);
        count_launch();
        w.epoch += (unsigned)min(m, n);
        LB_CUDA_CHECK(cudaGetLastError());
        return;
    }
    getrf_leaf_kernel<LEAF_W, LEAF_THREADS><<<p.G + S, LEAF_THREADS, 0, s>>>(p);
    count_launch();
    w.epoch += (unsigned)min(m, n);
    LB_CUDA_CHECK(cudaGetLastError());
}

// recursive panel (the DGETRF2 recursion, dgetrf2.f:216-263) down to LEAF_W columns.  The row interchanges
// of dgetrf2.f:236 and :263 are applied by the leaves themselves to every other column of the panel.
static void getrf_panel_rec(cudaStream_t s, const PanelCtx& pc, int off, int m, int n, double* A, i64 lda, int* ipiv,
                            int* info, int info_off) {
    if (m <= 0 || n <= 0) return;
    if (n <= LEAF_W) { getrf_leaf(s, pc, off, m, n, A, lda, ipiv, info, info_off); return; }
    const int mn = min(m, n);
    int n1 = LEAF_W;
    while (n1 * 2 < mn) n1 *= 2;                 // power-of-two multiple of the leaf width, < mn
    if (n1 >= mn) n1 = max(1, mn / 2);
    if (n1 > mn) n1 = mn;
    const int n2 = n - n1;
    getrf_panel_rec(s, pc, off, m, n1, A, lda, ipiv, info, info_off);
    double* A12 = A + (i64)n1 * lda;
    double* A21 = A + n1;
    double* A22 = A + n1 + (i64)n1 * lda;
    trsm(s, 'L', 'L', 'N', 'U', n1, n2, 1.0, A, lda, A12, lda);               // dgetrf2.f:240
    if (m > n1) {
        gemm(s, 'N', 'N', m - n1, n2, n1, -1.0, A21, lda, A12, lda, 1.0, A22, lda);   // dgetrf2.f:245
        // dgetrf2.f:250; the index shift of dgetrf2.f:257-259 is applied by the leaves themselves (piv_base)
        getrf_panel_rec(s, pc, off + n1, m - n1, n2, A22, lda, ipiv + n1, info, info_off + n1);
    }
}
// piv_base = info_off = row/column offset of the panel inside the caller's matrix
static void getrf_panel(cudaStream_t s, int m, int n, double* A, i64 lda, int* ipiv, int* info, int info_off) {
    PanelCtx pc{A, n, info_off};
    getrf_panel_rec(s, pc, 0, m, n, A, lda, ipiv, info, info_off);
}

// one mutex for every driver that uses the look-ahead streams / events of lb::aux() (runtime.cu): concurrent host threads calling
// different factorizations through the device API must not interleave their event joins
static std::recursive_mutex& g_lib_mutex = driver_mutex();

void getrf2(cudaStream_t s, int m, int n, double* A, i64 lda, int* ipiv, int* info) {
    std::lock_guard<std::recursive_mutex> lock(g_lib_mutex);
    LB_CUDA_CHECK(cudaMemsetAsync(info, 0, sizeof(int), s));
    getrf_panel(s, m, n, A, lda, ipiv, info, 0);
}

// Optional timeline of the blocked driver (LB200_TRACE_LU=1): timing events around every chunk GEMM and every panel;
// after the factorization the per-step intervals are printed to stderr (the call becomes synchronous).
struct LuTrace {
    struct Rec { int step, kind; cudaEvent_t e0, e1; };     // kind 0..4 = GEMM chunk, 10..14 = preparation of chunk, 20 = panel
    std::vector<Rec> recs;
    cudaEvent_t origin = nullptr;
    void mark(cudaStream_t st, int step, int kind, bool begin) {
        if (begin) {
            Rec r{step, kind, nullptr, nullptr};
            cudaEventCreate(&r.e0); cudaEventCreate(&r.e1);
            cudaEventRecord(r.e0, st);
            recs.push_back(r);
        } else {
            for (int i = (int)recs.size() - 1; i >= 0; --i)
                if (recs[i].step == step && recs[i].kind == kind) { cudaEventRecord(recs[i].e1, st); break; }
        }
    }
    void dump() {
        cudaDeviceSynchronize();
        fprintf(stderr, "LU_TRACE step kind start_ms end_ms\n");
        for (auto& r : recs) {
            float a = 0, b = 0;
            cudaEventElapsedTime(&a, origin, r.e0);
            cudaEventElapsedTime(&b, origin, r.e1);
            fprintf(stderr, "LU_TRACE %d %d %.3f %.3f\n", r.step, r.kind, a, b);
            cudaEventDestroy(r.e0); cudaEventDestroy(r.e1);
        }
        recs.clear();
    }
};

// Two-level driver.  level 0 = the blocked right-looking driver described above (NB = 512, panel = recursive DGETRF2 kernels).
// level 1 (lb200_set_getrf_super(4096); OFF by default) = the SAME driver with NB = 4096 whose "panel" is the level-0 driver on its
// own stream set: the trailing updates then have K = 4096, the 4096 interchanges of a step are composed and streamed (laswp_rows),
// and the level-0 factorization of the next 4096 columns runs beside the level-1 update.  Measured at n = 32768
// (profiles/r02c_two_level_lu.txt): same IPIV, the update GEMMs drop from 735 to 598 ms as intended, but the level-0 factorization
// beside GEMM CTAs that live 8x longer takes 154 instead of ~55 ms per block, the 4096-wide U12 solves (8% of the flops instead of
// 1%) leave 7-8 ms gaps before every chunk, and the first block has nothing to hide behind: 832-848 ms against 794.  A is the (base, *) corner of the caller's matrix: pivots, INFO and the interchange plans use rows of the CALLER's matrix
// (base + local row), ipiv0 is the caller's pivot array.
static int g_super_nb = 0;                   // 0 = single level (default: the two-level driver was measured slower, see below)
void getrf_set_super(int nb) { g_super_nb = nb; }
static int g_driver_top_level = 0;           // level of the outermost getrf_driver call in flight (under g_lib_mutex)
static void getrf_driver(int level, cudaStream_t s, int m, int n, double* A, i64 lda, int* ipiv0, int base, int* info) {
    static const bool trace_on = getenv("LB200_TRACE_LU") != nullptr;
    // Profiling aid (WRONG RESULTS, timing only): LB200_ABLATE bit 0 = skip the interchanges left of the panel, bit 1 = skip the
    // interchanges of the trailing columns, bit 2 = skip the U12 solve -- measures what each memory-bound stage costs the update GEMMs.
    static const int ablate = getenv("LB200_ABLATE") ? atoi(getenv("LB200_ABLATE")) : 0;
    LuTrace trace;
    LuTrace* tr = (trace_on && g_driver_top_level == level && base == 0) ? &trace : nullptr;    // the top-level call only
    if (tr) { cudaEventCreate(&tr->origin); cudaEventRecord(tr->origin, s); }
    if (m <= 0 || n <= 0) return;
    const int mn = min(m, n);
    const int nb = level == 0 ? min(g_nb, 2048) : g_super_nb;
    int* const ipiv = ipiv0 + base;              // pivots of the local rows
    double* const Arow0 = A - base;              // row 0 of the caller's matrix in local column 0 (interchanges use caller rows)
    auto factor_panel = [&](cudaStream_t st, int pm, int pn, double* P, int off) {
        // panel at local offset (off, off): level 0 = the recursive kernels, level 1 = the level-0 driver
        if (level == 0) getrf_panel(st, pm, pn, P, lda, ipiv + off, info, base + off);
        else getrf_driver(0, st, pm, pn, P, lda, ipiv0, base + off, info);
    };
    if (nb >= mn) { factor_panel(s, m, n, A, 0); return; }

    const bool la = g_lookahead != 0;
    Aux& ax = aux(level);
    StreamOut* so = (level == 0 && base == 0) ? stream_out() : nullptr;
    // Streams (look-ahead on): sp = panel (highest priority), sq = memory-bound preparation of the trailing
    // columns (interchanges + U12 solve, medium priority), su = trailing GEMMs (low priority), sl = interchanges
    // left of the panel (low priority, off the critical path).  The trailing columns are processed in up to five
    // chunks -- the next panel's columns, the panel after that, 2048 more, a quarter of the rest, the remainder -- so that the
    // preparation of chunk c+1 overlaps with the GEMM of chunk c, and the preparation of the NEXT step's first chunk
    // (this step's chunk 1) can start while this step's big chunks are still running.
    cudaStream_t sp = la ? ax.panel_stream : s;
    cudaStream_t sq = la ? ax.prep_stream : s;
    cudaStream_t su = la ? ax.update_stream : s;
    cudaStream_t sl = la ? ax.side_stream : s;
    cudaEvent_t ev_panel = ax.ev[0], ev_next = ax.ev[1], ev_join = ax.ev[2], ev_plan = ax.ev[8], ev_left = ax.ev[9];
    cudaEvent_t ev_gemm = ax.ev[10];
    cudaEvent_t ev_prep[5] = {ax.ev[11], ax.ev[12], ax.ev[13], ax.ev[14], ax.ev[15]};
    cudaEvent_t ev_upd[5] = {ax.ev[16], ax.ev[17], ax.ev[18], ax.ev[19], ax.ev[20]};      // recorded after the GEMM of chunk q
    if (la) {
        LB_CUDA_CHECK(cudaEventRecord(ev_join, s));
        LB_CUDA_CHECK(cudaStreamWaitEvent(sp, ev_join, 0));
        LB_CUDA_CHECK(cudaStreamWaitEvent(sq, ev_join, 0));
        LB_CUDA_CHECK(cudaStreamWaitEvent(su, ev_join, 0));
        LB_CUDA_CHECK(cudaStreamWaitEvent(sl, ev_join, 0));
    }
    // first panel
    factor_panel(sp, m, min(nb, mn), A, 0);
    if (la) LB_CUDA_CHECK(cudaEventRecord(ev_panel, sp));
    bool gemm_recorded = false;
    struct DeferredSwap { void* plan; int ncols, npiv; };
    std::vector<DeferredSwap> deferred;          // interchanges left of the panel that wait for the tail (see below)
    // deferred[d] is the plan of panel d+1 and applies to the columns [0, (d+1)*nb): composed per block column and applied in one
    // streaming pass (laswp_apply_chain); plan by plan when the composition is not applicable (uneven panels, very tall matrices)
    auto flush_deferred = [&]() {
        if (deferred.empty()) return;
        bool chain_ok = g_defer_left == 1;
        std::vector<void*> pl;
        for (size_t d = 0; d < deferred.size(); ++d) {
            pl.push_back(deferred[d].plan);
            if (deferred[d].ncols != (int)(d + 1) * nb || deferred[d].npiv != nb) chain_ok = false;
        }
        // (with base > 0 the permutations cover the caller's rows 0 .. base + m; they are the identity above base)
        if (!(chain_ok && laswp_apply_chain(sl, base + m, nb, (int)pl.size(), pl.data(), Arow0, lda)))
            for (const DeferredSwap& d : deferred) laswp_apply_plan(sl, d.ncols, Arow0, lda, d.plan, d.npiv);
        for (const DeferredSwap& d : deferred) laswp_plan_free(sl, d.plan);
        deferred.clear();
    };
    int pc_lo[5], pc_hi[5], npc = 0;     // column ranges of the previous step's GEMM chunks (their completion = ev_upd[q])

    for (int j = 0; j < mn; j += nb) {
        const int jb = min(nb, mn - j);
        const int jn = j + jb;                       // first column after this panel
        double* Ajj = A + j + (i64)j * lda;
        const int jb2 = (jn < mn) ? min(nb, mn - jn) : 0;
        // chunk boundaries (columns): [c[0],c[1]) = next panel's columns (or everything if there is no next panel),
        // [c[1],c[2]) = the panel after that, a narrow third chunk (its preparation can only start after the whole update j-1 and
        // must be finished when the GEMMs of the first two chunks are), then a quarter of the rest, then the remainder
        int c[6];
        int nchunk = 0;
        c[0] = jn;
        if (jn < n) {
            const int w1 = (jb2 > 0) ? jb2 : (n - jn);
            c[++nchunk] = jn + w1;
            int rest = n - jn - w1;
            if (rest > 0) {
                const int w2 = min(nb, rest);
                c[nchunk + 1] = c[nchunk] + w2;
                ++nchunk;
                rest -= w2;
            }
            if (rest >= 8192) {
                c[nchunk + 1] = c[nchunk] + 2048;
                ++nchunk;
                rest -= 2048;
            }
            if (rest > 0) {
                int h1 = (rest >= 4096) ? ((rest / 4 + 63) / 64) * 64 : rest;
                c[nchunk + 1] = c[nchunk] + h1;
                ++nchunk;
                if (h1 < rest) c[++nchunk] = n;
            }
        }
        if (la) LB_CUDA_CHECK(cudaStreamWaitEvent(sq, ev_panel, 0));                 // panel j is factored
        // one plan for this panel's interchanges (dgetrf.f:193,199), applied to all column ranges
        // level 1: 4096 interchanges per step are beyond one plan; they are composed and streamed per column range (laswp_rows)
        void* plan = (level == 0) ? laswp_plan(sq, base + j + 1, base + jn, ipiv0, 1) : nullptr;
        auto swap_cols = [&](cudaStream_t st, int ncols, double* cols_row0) {       // cols_row0: caller row 0 of the first column
            if (plan) laswp_apply_plan(st, ncols, cols_row0, lda, plan, jb);
            else laswp_rows(st, ncols, base + m, cols_row0, lda, base + j + 1, base + jn, ipiv0, 1);
        };
        // preparation on sq: interchanges + block row of U for every chunk, in order
        for (int q = 0; q < nchunk; ++q) {
            const int w = c[q + 1] - c[q];
            // the columns of this chunk must have received update j-1: wait for exactly the GEMM chunks of the previous step that
            // cover them (chunk 0 here was chunk 1 there, chunk 1 here lies inside chunk 2 there, ...), so that the preparation
            // of the first chunks of step j overlaps with the big GEMMs of step j-1
            if (la)
                for (int r = 0; r < npc; ++r)
                    if (pc_lo[r] < c[q + 1] && c[q] < pc_hi[r]) LB_CUDA_CHECK(cudaStreamWaitEvent(sq, ev_upd[r], 0));
            if (tr) tr->mark(sq, j / nb, 10 + q, true);
            if (!(ablate & 2)) swap_cols(sq, w, Arow0 + (i64)c[q] * lda);                            // dgetrf.f:199
            // (inverted 32 x 32 diagonal blocks -- trsm_set_inverse_leaves, used by DPOTRF -- were measured neutral here: 806 vs 807 ms)
            if (!(ablate & 4)) trsm(sq, 'L', 'L', 'N', 'U', jb, w, 1.0, Ajj, lda, A + j + (i64)c[q] * lda, lda);   // dgetrf.f:204
            if (tr) tr->mark(sq, j / nb, 10 + q, false);
            if (la) LB_CUDA_CHECK(cudaEventRecord(ev_prep[q], sq));
        }
        // the interchanges left of the panel (below) and the next step's work on sq come after the whole update j-1
        if (la && gemm_recorded) LB_CUDA_CHECK(cudaStreamWaitEvent(sq, ev_gemm, 0));
        if (la) LB_CUDA_CHECK(cudaEventRecord(ev_plan, sq));
        if (so && la && jn < n) {
            // block row j of U (rows j..jn, columns jn..n) is final once every chunk has been prepared: start its
            // download now (host-resident callers, lb::StreamOut)
            LB_CUDA_CHECK(cudaStreamWaitEvent(so->copy_stream, ev_plan, 0));
            LB_CUDA_CHECK(cudaMemcpy2DAsync(so->host + j + (i64)jn * so->ldh, so->ldh * 8, A + j + (i64)jn * lda, lda * 8,
                                            (size_t)jb * 8, n - jn, cudaMemcpyDeviceToHost, so->copy_stream));
            so->done_cols = jn;
        }
        // GEMMs on su; the next panel is factored right after the first chunk
        for (int q = 0; q < nchunk; ++q) {
            const int w = c[q + 1] - c[q];
            if (la) LB_CUDA_CHECK(cudaStreamWaitEvent(su, ev_prep[q], 0));
            if (tr) tr->mark(su, j / nb, q, true);
            if (jn < m)
                gemm(su, 'N', 'N', m - jn, w, jb, -1.0, A + jn + (i64)j * lda, lda, A + j + (i64)c[q] * lda, lda, 1.0,
                     A + jn + (i64)c[q] * lda, lda);                                                   // dgetrf.f:212
            if (tr) tr->mark(su, j / nb, q, false);
            if (la) LB_CUDA_CHECK(cudaEventRecord(ev_upd[q], su));
            if (q == 0 && jb2 > 0) {
                if (la) {
                    LB_CUDA_CHECK(cudaEventRecord(ev_next, su));
                    LB_CUDA_CHECK(cudaStreamWaitEvent(sp, ev_next, 0));
                }
                // factor the next panel (overlaps with the rest of this update when look-ahead is on)
                // the leaves write absolute pivot rows (dgetrf.f:187-189 shift applied at the source)
                if (tr) tr->mark(sp, j / nb, 20, true);
                factor_panel(sp, m - jn, jb2, A + jn + (i64)jn * lda, jn);
                if (tr) tr->mark(sp, j / nb, 20, false);
                if (la) LB_CUDA_CHECK(cudaEventRecord(ev_panel, sp));
            }
        }
        if (la) { LB_CUDA_CHECK(cudaEventRecord(ev_gemm, su)); gemm_recorded = true; }
        npc = nchunk;
        for (int q = 0; q < nchunk; ++q) { pc_lo[q] = c[q]; pc_hi[q] = c[q + 1]; }
        // interchanges to the left of the panel (dgetrf.f:193).  Those columns are final L columns whose only
        // remaining reader was the trailing GEMM of the previous step (sq waited for it above), so they run on
        // a low-priority side stream concurrently with this step's trailing update.
        // These interchanges are scattered DRAM accesses (one 128-byte line fill per moved element) that slow the concurrent GEMMs
        // down by about their own duration (LB200_ABLATE=1: 809 -> 791 ms at n = 32768), and moving them into the tail of the
        // factorization does not hide them either (808 ms).  So they are DEFERRED and COMPOSED: the plans are kept until the trailing
        // matrix is down to g_defer_tail_rows rows (default: the last panel), then every block column receives its whole chain of
        // interchanges as ONE row permutation in a streaming pass (laswp_apply_chain): 809 -> 793 ms, bit-identical results.
        const bool defer = la && plan && g_defer_left && (m - jn) > g_defer_tail_rows && jn < mn;
        if (defer) {
            if (j > 0 && !(ablate & 1)) deferred.push_back(DeferredSwap{plan, j, jb});
            else laswp_plan_free(sl, plan);
        } else {
            if (la) LB_CUDA_CHECK(cudaStreamWaitEvent(sl, ev_plan, 0));
            flush_deferred();
            if (j > 0 && !(ablate & 1)) swap_cols(sl, j, Arow0);
            if (plan) laswp_plan_free(sl, plan);
        }
    }
    if (!deferred.empty()) {        // (not reached: the last step is never deferred)
        if (la) LB_CUDA_CHECK(cudaStreamWaitEvent(sl, ev_plan, 0));
        flush_deferred();
    }
    if (la) {
        LB_CUDA_CHECK(cudaEventRecord(ev_left, sl));
        LB_CUDA_CHECK(cudaStreamWaitEvent(s, ev_left, 0));
        LB_CUDA_CHECK(cudaEventRecord(ev_join, su));
        LB_CUDA_CHECK(cudaStreamWaitEvent(s, ev_join, 0));
        LB_CUDA_CHECK(cudaEventRecord(ev_next, sp));
        LB_CUDA_CHECK(cudaStreamWaitEvent(s, ev_next, 0));
        LB_CUDA_CHECK(cudaEventRecord(ev_plan, sq));
        LB_CUDA_CHECK(cudaStreamWaitEvent(s, ev_plan, 0));
    }
    if (tr) { tr->dump(); cudaEventDestroy(tr->origin); }
}

void getrf(cudaStream_t s, int m, int n, double* A, i64 lda, int* ipiv, int* info) {
    std::lock_guard<std::recursive_mutex> lock(g_lib_mutex);
    LB_CUDA_CHECK(cudaMemsetAsync(info, 0, sizeof(int), s));
    if (m <= 0 || n <= 0) return;
    // level 1 needs the composed long-list interchanges (rows <= 55K, laswp_impl) and is only worth it with >= 3 outer blocks;
    // host-streamed callers (lb::StreamOut) keep the single-level driver, whose block rows of U they download as they finish
    const bool super = g_super_nb >= 1024 && g_lookahead != 0 && min(m, n) >= 3 * g_super_nb && m <= 55 * 1024 && !stream_out() &&
                       g_super_nb > min(g_nb, 2048);
    g_driver_top_level = super ? 1 : 0;
    getrf_driver(super ? 1 : 0, s, m, n, A, lda, ipiv, 0, info);
}

// DGETRS (SRC/dgetrs.f:181-218); argument checks live in the Fortran-ABI layer
void getrs(cudaStream_t s, char trans, int n, int nrhs, const double* A, i64 lda, const int* ipiv, double* B, i64 ldb) {
    if (n <= 0 || nrhs <= 0) return;
    const bool notran = (trans == 'N' || trans == 'n');
    if (notran) {
        laswp_rows(s, nrhs, n, B, ldb, 1, n, ipiv, 1);
        trsm(s, 'L', 'L', 'N', 'U', n, nrhs, 1.0, A, lda, B, ldb);
        trsm(s, 'L', 'U', 'N', 'N', n, nrhs, 1.0, A, lda, B, ldb);
    } else {
        trsm(s, 'L', 'U', 'T', 'N', n, nrhs, 1.0, A, lda, B, ldb);
        trsm(s, 'L', 'L', 'T', 'U', n, nrhs, 1.0, A, lda, B, ldb);
        laswp_rows(s, nrhs, n, B, ldb, 1, n, ipiv, -1);
    }
}

// ------------------------------------------------------------------------------------------------
// DGETRI (SRC/dgetri.f:150-259): inverse from the LU factors.  The reference forms inv(U) in place
// (DTRTRI) and then solves inv(A)*L = inv(U) block column by block column; on the GPU both steps are triangular
// solves against the identity in a scratch matrix W (they run as DMMA GEMMs through the recursive DTRSM):
//     W = I;  U W = W  (W = inv(U));  W L = W  (W = inv(U) inv(L));  A(:, c) = W(:, perm(c))
// where perm composes the column interchanges of dgetri.f:250-255 (reverse order).  As in DTRTRI, an exactly zero
// U(i,i) gives INFO = i and leaves A untouched (dtrtri.f:169-175, dgetri.f:181-183) -- decided on the device.
__global__ void getri_diag_check_kernel(int n, const double* __restrict__ A, i64 lda, int* info) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && A[i + (i64)i * lda] == 0.0) atomicMin((unsigned*)info, (unsigned)(i + 1));
}
__global__ void getri_info_init_kernel(int* info) { *info = 0x7fffffff; }
__global__ void getri_info_fix_kernel(int* info) { if (*info == 0x7fffffff) *info = 0; }
__global__ void getri_perm_kernel(int n, const int* __restrict__ ipiv, int* __restrict__ perm) {
    // single thread: the interchanges are applied sequentially, last to first (dgetri.f:250-255)
    for (int c = 0; c < n; ++c) perm[c] = c;
    for (int j = n - 2; j >= 0; --j) {
        const int jp = ipiv[j] - 1;
        if (jp != j) { int t = perm[j]; perm[j] = perm[jp]; perm[jp] = t; }
    }
}
__global__ void getri_gather_kernel(int n, const double* __restrict__ W, i64 ldw, const int* __restrict__ perm,
                                    const int* __restrict__ info, double* __restrict__ A, i64 lda) {
    if (*info != 0) return;                                  // singular: A keeps the LU factors
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    for (int c = blockIdx.y; c < n; c += gridDim.y) A[i + (i64)c * lda] = W[i + (i64)perm[c] * ldw];
}

void getri(cudaStream_t s, int n, double* A, i64 lda, const int* ipiv, int* info) {
    if (n <= 0) { LB_CUDA_CHECK(cudaMemsetAsync(info, 0, sizeof(int), s)); return; }
    const i64 ldw = ((i64)n + 1) & ~1LL;
    double* W = (double*)ws_alloc(s, sizeof(double) * (size_t)ldw * n);
    int* perm = (int*)ws_alloc(s, sizeof(int) * (size_t)n);
    getri_info_init_kernel<<<1, 1, 0, s>>>(info);
    getri_diag_check_kernel<<<ceil_div(n, 256), 256, 0, s>>>(n, A, lda, info);
    getri_info_fix_kernel<<<1, 1, 0, s>>>(info);
    getri_perm_kernel<<<1, 1, 0, s>>>(n, ipiv, perm);
    count_launch(4);
    laset(s, 'A', n, n, 0.0, 1.0, W, ldw);
    trsm(s, 'L', 'U', 'N', 'N', n, n, 1.0, A, lda, W, ldw);              // W = inv(U)          (dgetri.f:181)
    trsm(s, 'R', 'L', 'N', 'U', n, n, 1.0, A, lda, W, ldw);              // W = inv(U) inv(L)   (dgetri.f:198-248)
    dim3 grid(ceil_div(n, 256), (unsigned)min(n, 4096));
    getri_gather_kernel<<<grid, 256, 0, s>>>(n, W, ldw, perm, info, A, lda);
    count_launch();
    ws_free(s, W);
    ws_free(s, perm);
    LB_CUDA_CHECK(cudaGetLastError());
}

}  // namespace lb
