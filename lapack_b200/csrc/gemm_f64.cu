// gemm_f64.cu -- FP64 tensor-core GEMM / SYRK for sm_100a.
//
// Replaces the reference triple loops BLAS/SRC/dgemm.f:322-383 and dsyrk.f:278-355 on the trailing
// updates of DGETRF (SRC/dgetrf.f:212), DPOTRF (SRC/dpotrf.f:216,225; VARIANTS/cholesky/RL) and DLARFB
// (SRC/dlarfb.f:270,287).
//
// Hardware mapping.  On sm_100a the FP64 tensor pipe is only reachable through the warp-level
// `mma.sync.m8n8k4.f64` (SASS DMMA.8x8x4); tcgen05 has no f64 kind.  Each warp owns a 64(m) x 32(n)
// block of C held in registers (64 doubles/thread) and issues 32 independent DMMAs per k-step of 4.
// The MMA is issued "transposed" (first operand = op(B) fragment, second = op(A) fragment) so that
// every thread ends up with two consecutive rows of one column of C, i.e. a 16-byte column-major store.
// Operand tiles are staged global->shared by a multi-stage cp.async (LDGSTS) ring; the shared layouts
// are padded (+4 doubles) so that the 8-byte fragment loads are bank-conflict free.
//
// Roofline: 2*M*N*K flops per launch against the FP64 DMMA peak; minimum HBM traffic
// 8*(2*M*N + M*K + K*N) bytes.
#include "lb_internal.h"
#include <vector>

namespace lb {

unsigned long long g_launches = 0;

struct GemmParams {
    int M, N, K;
    double alpha, beta;
    const double* A; i64 lda;
    const double* B; i64 ldb;
    double* C; i64 ldc;
    int tri;              // 0 full, 1 lower, 2 upper
    int tiles_m, tiles_n;
    int kchunk;           // split-K: blockIdx.y owns k in [y*kchunk, (y+1)*kchunk); C advances by c_zstride
    i64 c_zstride;
    const int* guard;     // device word: when non-null and non-zero the launch does nothing (Cholesky after INFO > 0)
};

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, int src_bytes) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async8(void* smem, const void* gmem, int src_bytes) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(s), "l"(gmem), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

// D(8x8) += X(8x4,row) * Y(4x8,col); thread t supplies X[t/4][t%4], Y[t%4][t/4], holds D[t/4][2(t%4)+{0,1}].
__device__ __forceinline__ void dmma884(double& d0, double& d1, double x, double y) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(d0), "+d"(d1)
                 : "d"(x), "d"(y));
}

constexpr int BK = 16;      // default k-tile
constexpr int PAD = 4;

// Shared-memory tile of one operand: EXT = extent along m (or n), BK along k.
//   K-major  (KMAJ=true):  element (e,k) at e*(BK+PAD) + k      (k contiguous in global memory)
//   MN-major (KMAJ=false): element (e,k) at k*(EXT+PAD) + e      (e contiguous in global memory)
template <int EXT, bool KMAJ, int BK = 16>
struct TileLayout {
    static constexpr int ELEMS = KMAJ ? EXT * (BK + PAD) : BK * (EXT + PAD);
    __device__ static __forceinline__ int off(int e, int k) { return KMAJ ? e * (BK + PAD) + k : k * (EXT + PAD) + e; }
};

// Per-thread state for the cp.async copies of one operand tile.  All address arithmetic that does not
// depend on the k-tile index is done once in init(); issue() only adds the k offset and clamps at the K
// edge (zero fill), so the steady-state loop spends its issue slots on LDS/DMMA.
//   g points at element (e=0,k=0) of the whole operand; e0 = tile origin, elim/klim = operand extents.
template <int EXT, bool KMAJ, bool AL16, int NTHREADS, int BK = 16>
struct TileLoader {
    static constexpr int N = AL16 ? (EXT * BK / 2) / NTHREADS : (EXT * BK) / NTHREADS;
    static_assert((AL16 ? (EXT * BK / 2) : (EXT * BK)) % NTHREADS == 0, "tile/thread mismatch");
    const double* gp[N];   // global address of the chunk for k-tile 0 (nullptr-safe: clamped to g)
    int so[N];             // shared offset (doubles) inside the stage
    int ebytes[N];         // bytes available along e (AL16 && !KMAJ), or 0/full validity flag
    int kloc[N];           // k index of the chunk inside the tile
    i64 kstep;             // pointer increment per k-tile (doubles)
    const double* g0;

    __device__ __forceinline__ void init(const double* __restrict__ g, i64 ld, int e0, int elim, int tid) {
        g0 = g;
        kstep = KMAJ ? (i64)BK : (i64)BK * ld;
#pragma unroll
        for (int i = 0; i < N; ++i) {
            int idx = tid + i * NTHREADS;
            int e, k;
            if (AL16) {
                if (KMAJ) { e = idx / (BK / 2); k = 2 * (idx % (BK / 2)); }
                else { k = idx / (EXT / 2); e = 2 * (idx % (EXT / 2)); }
            } else {
                if (KMAJ) { e = idx / BK; k = idx % BK; }
                else { k = idx / EXT; e = idx % EXT; }
            }
            int ge = e0 + e;
            kloc[i] = k;
            so[i] = KMAJ ? e * (BK + PAD) + k : k * (EXT + PAD) + e;
            int rem = elim - ge;
            if (AL16 && !KMAJ) ebytes[i] = rem >= 2 ? 16 : (rem == 1 ? 8 : 0);
            else ebytes[i] = rem >= 1 ? (AL16 ? 16 : 8) : 0;
            gp[i] = ebytes[i] ? (KMAJ ? g + (i64)ge * ld + k : g + (i64)k * ld + ge) : g;
        }
    }
    __device__ __forceinline__ void issue(double* sm, int kt, int klim) {
        const int krem = klim - kt * BK;      // valid k's left from the start of this tile
        const i64 koff = (i64)kt * kstep;
#pragma unroll
        for (int i = 0; i < N; ++i) {
            int bytes;
            if (AL16 && KMAJ) {
                int r = krem - kloc[i];
                bytes = ebytes[i] ? (r >= 2 ? 16 : (r == 1 ? 8 : 0)) : 0;
            } else {
                bytes = (kloc[i] < krem) ? ebytes[i] : 0;
            }
            const double* src = bytes ? gp[i] + koff : g0;
            if (AL16) cp_async16(sm + so[i], src, bytes);
            else cp_async8(sm + so[i], src, bytes);
        }
    }
};

// Tile columns [c0, c0+nc) that a group of tile rows [first_m, first_m+gsz) needs when only one triangle of C is computed
// (tri 1: lower, tiles with some m >= n; tri 2: upper, tiles with some m <= n).
constexpr int TRI_GROUP = 16;
__host__ __device__ __forceinline__ void tri_group_cols(int tri, int first_m, int gsz, int bm, int bn, int tiles_n, int& c0, int& nc) {
    if (tri == 1) {
        c0 = 0;
        const int last = ((first_m + gsz) * bm - 1) / bn;        // last tile column touched by the lowest row of the group
        nc = (last + 1 < tiles_n) ? last + 1 : tiles_n;
    } else {
        c0 = (first_m * bm) / bn;                                 // first tile column touched by the top row of the group
        if (c0 > tiles_n) c0 = tiles_n;
        nc = tiles_n - c0;
    }
}

// Loader for INTERIOR tiles (no M/N/K edge, 16-byte aligned operands): the chunks of one thread are a fixed stride apart in global and
// in shared memory, so a k-tile costs N LDGSTS + N pointer additions and no predicates (the general loader above spends ~150 integer
// instructions per k-tile on clamping).
template <int EXT, bool KMAJ, int NTHREADS, int BK = 16>
struct LeanLoader {
    static constexpr int N = (EXT * BK / 2) / NTHREADS;
    static constexpr int PER = KMAJ ? NTHREADS / (BK / 2) : NTHREADS / (EXT / 2);        // e (KMAJ) or k (MN-major) distance between chunks
    static constexpr int SSTRIDE = KMAJ ? PER * (BK + PAD) : PER * (EXT + PAD);           // shared distance (doubles)
    static_assert(KMAJ ? (NTHREADS % (BK / 2) == 0) : (NTHREADS % (EXT / 2) == 0), "lean loader: thread count");
    const double* g;       // chunk 0 of the next k-tile
    i64 gstride, kstep;
    int s0;
    __device__ __forceinline__ void init(const double* __restrict__ gbase, i64 ld, int e0, int tid) {
        int e, k;
        if (KMAJ) { e = tid / (BK / 2); k = 2 * (tid % (BK / 2)); }
        else { k = tid / (EXT / 2); e = 2 * (tid % (EXT / 2)); }
        s0 = KMAJ ? e * (BK + PAD) + k : k * (EXT + PAD) + e;
        g = KMAJ ? gbase + (i64)(e0 + e) * ld + k : gbase + (i64)k * ld + (e0 + e);
        gstride = (i64)PER * ld;
        kstep = KMAJ ? (i64)BK : (i64)BK * ld;
    }
    __device__ __forceinline__ void issue(double* sm) {
        const double* q = g;
        double* d = sm + s0;
#pragma unroll
        for (int i = 0; i < N; ++i) {
            cp_async16(d + i * SSTRIDE, q, 16);
            q += gstride;
        }
        g += kstep;
    }
};

template <int BM, int BN, int WARPS_M, int WARPS_N, bool A_KMAJ, bool B_KMAJ, bool AL16, int STAGES, int BK = 16, int VAR = 0>
__global__ void __launch_bounds__(WARPS_M* WARPS_N * 32,
                                  ((BM / WARPS_M) * (BN / WARPS_N) <= 1024) ? (512 / (WARPS_M * WARPS_N * 32) > 0 ? 512 / (WARPS_M * WARPS_N * 32) : 1)
                                                                            : ((WARPS_M * WARPS_N <= 4) ? 2 : 1))
    gemm_f64_dmma_kernel(GemmParams p) {
    constexpr int NTHREADS = WARPS_M * WARPS_N * 32;
    constexpr int WM = BM / WARPS_M, WN = BN / WARPS_N;
    constexpr int MT = WM / 8, NT = WN / 8;
    using LA = TileLayout<BM, A_KMAJ, BK>;
    using LB = TileLayout<BN, B_KMAJ, BK>;
    constexpr int STAGE_ELEMS = LA::ELEMS + LB::ELEMS;

    extern __shared__ __align__(16) double smem[];

    // grouped rasterisation: 16 tile-rows at a time so that a wave of CTAs covers a squarish patch of C
    constexpr int GROUP = TRI_GROUP;
    int pid = blockIdx.x;
    int pid_m, pid_n;
    if (p.tri == 0) {
        int width = GROUP * p.tiles_n;
        int group_id = pid / width;
        int first_m = group_id * GROUP;
        int gsz = min(p.tiles_m - first_m, GROUP);
        pid_m = first_m + (pid % width) % gsz;
        pid_n = (pid % width) / gsz;
    } else {
        // one triangle of C (DSYRK and the look-ahead block column of DPOTRF): the grid holds, per group of tile rows, only the tile
        // columns that can touch the triangle (tri_group_cols, same formula on the host), not tiles_m x tiles_n CTAs of which half
        // exit at once
        int first_m = 0, gsz, c0, nc;
        for (;;) {
            gsz = min(p.tiles_m - first_m, GROUP);
            tri_group_cols(p.tri, first_m, gsz, BM, BN, p.tiles_n, c0, nc);
            if (pid < gsz * nc) break;
            pid -= gsz * nc;
            first_m += GROUP;
        }
        pid_m = first_m + pid % gsz;
        pid_n = c0 + pid / gsz;
    }
    const int m0 = pid_m * BM, n0 = pid_n * BN;
    if (p.guard && *p.guard != 0) return;
    if (p.tri == 1 && m0 + BM - 1 < n0) return;       // tile strictly above the diagonal
    if (p.tri == 2 && n0 + BN - 1 < m0) return;       // tile strictly below the diagonal

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int warp_m = warp % WARPS_M, warp_n = warp / WARPS_M;
    const int g4 = lane >> 2, t4 = lane & 3;

    // split-K slice owned by this CTA (gridDim.y == 1 and kchunk == K for an ordinary launch)
    {
        const int kb = blockIdx.y * p.kchunk;
        p.A += A_KMAJ ? (i64)kb : (i64)kb * p.lda;
        p.B += B_KMAJ ? (i64)kb : (i64)kb * p.ldb;
        p.C += (i64)blockIdx.y * p.c_zstride;
        p.K = min(p.K - kb, p.kchunk);
    }

    // pull the C tile towards L2 while the main loop runs (the epilogue is a read-modify-write)
    if (p.beta != 0.0) {
        constexpr int LINES_PER_COL = BM / 16;
        for (int l = tid; l < BN * LINES_PER_COL; l += NTHREADS) {
            int n = n0 + l / LINES_PER_COL, m = m0 + (l % LINES_PER_COL) * 16;
            if (n < p.N && m < p.M) {
                const double* ptr = p.C + (i64)n * p.ldc + m;
                asm volatile("prefetch.global.L2 [%0];\n" ::"l"(ptr));
            }
        }
    }

    double acc[MT][NT][2];
#pragma unroll
    for (int a = 0; a < MT; ++a)
#pragma unroll
        for (int b = 0; b < NT; ++b) acc[a][b][0] = acc[a][b][1] = 0.0;

    const int KT = (p.K + BK - 1) / BK;
    // fragment offsets inside a stage (k added per step)
    int aoff[MT], boff[NT];
#pragma unroll
    for (int a = 0; a < MT; ++a) aoff[a] = LA::off(warp_m * WM + a * 8 + g4, t4);
#pragma unroll
    for (int b = 0; b < NT; ++b) boff[b] = LB::off(warp_n * WN + b * 8 + g4, t4);
    constexpr int AKS = A_KMAJ ? 1 : (BM + PAD);     // shared stride of one k step
    constexpr int BKS = B_KMAJ ? 1 : (BN + PAD);
    auto compute = [&](const double* sa, const double* sb) {
#pragma unroll
        for (int kk = 0; kk < BK; kk += 4) {
            double af[MT], bf[NT];
#pragma unroll
            for (int a = 0; a < MT; ++a) af[a] = sa[aoff[a] + kk * AKS];
#pragma unroll
            for (int b = 0; b < NT; ++b) bf[b] = sb[boff[b] + kk * BKS];
#pragma unroll
            for (int a = 0; a < MT; ++a)
#pragma unroll
                for (int b = 0; b < NT; ++b) dmma884(acc[a][b][0], acc[a][b][1], bf[b], af[a]);
        }
    };

    const bool lean = (VAR == 1) && AL16 && (m0 + BM <= p.M) && (n0 + BN <= p.N) && (p.K % BK == 0) && (KT >= STAGES);
    if (lean) {
        LeanLoader<BM, A_KMAJ, NTHREADS, BK> la;
        LeanLoader<BN, B_KMAJ, NTHREADS, BK> lb_;
        la.init(p.A, p.lda, m0, tid);
        lb_.init(p.B, p.ldb, n0, tid);
#pragma unroll
        for (int s = 0; s < STAGES - 1; ++s) {
            la.issue(smem + s * STAGE_ELEMS);
            lb_.issue(smem + s * STAGE_ELEMS + LA::ELEMS);
            cp_async_commit();
        }
        int st_c = 0, st_l = STAGES - 1;             // stage computed / stage loaded next
        for (int kt = 0; kt < KT; ++kt) {
            cp_async_wait<STAGES - 2>();
            __syncthreads();
            if (kt + STAGES - 1 < KT) {
                la.issue(smem + st_l * STAGE_ELEMS);
                lb_.issue(smem + st_l * STAGE_ELEMS + LA::ELEMS);
            }
            cp_async_commit();
            const double* sa = smem + st_c * STAGE_ELEMS;
            compute(sa, sa + LA::ELEMS);
            st_c = (st_c + 1 == STAGES) ? 0 : st_c + 1;
            st_l = (st_l + 1 == STAGES) ? 0 : st_l + 1;
        }
    } else {
    TileLoader<BM, A_KMAJ, AL16, NTHREADS, BK> ldA;
    TileLoader<BN, B_KMAJ, AL16, NTHREADS, BK> ldB;
    ldA.init(p.A, p.lda, m0, p.M, tid);
    ldB.init(p.B, p.ldb, n0, p.N, tid);

    auto issue = [&](int kt) {
        if (kt < KT) {
            double* sa = smem + (kt % STAGES) * STAGE_ELEMS;
            ldA.issue(sa, kt, p.K);
            ldB.issue(sa + LA::ELEMS, kt, p.K);
        }
        cp_async_commit();
    };

#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) issue(s);

    for (int kt = 0; kt < KT; ++kt) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        issue(kt + STAGES - 1);
        const double* sa = smem + (kt % STAGES) * STAGE_ELEMS;
        compute(sa, sa + LA::ELEMS);
    }
    }
    cp_async_wait<0>();

    // epilogue: thread holds C(m = mrow + {0,1}, n = ncol) for every (a,b).  All loads of one column
    // batch are issued before its stores so that the read-modify-write costs one memory round trip per
    // batch instead of one per element.
    const bool vec_ok = AL16 && ((p.ldc & 1) == 0) && ((((uintptr_t)p.C) & 15) == 0);
    const double alpha = p.alpha, beta = p.beta;
    const int mbase = m0 + warp_m * WM + 2 * t4;
#pragma unroll
    for (int b = 0; b < NT; ++b) {
        int n = n0 + warp_n * WN + b * 8 + g4;
        if (n >= p.N) continue;
        double* ccol = p.C + (i64)n * p.ldc;
        double c0[MT], c1[MT];
        bool ok0[MT], ok1[MT];
#pragma unroll
        for (int a = 0; a < MT; ++a) {
            int m = mbase + a * 8;
            ok0[a] = (m < p.M);
            ok1[a] = (m + 1 < p.M);
            if (p.tri == 1) { ok0[a] = ok0[a] && (m >= n); ok1[a] = ok1[a] && (m + 1 >= n); }
            if (p.tri == 2) { ok0[a] = ok0[a] && (m <= n); ok1[a] = ok1[a] && (m + 1 <= n); }
            c0[a] = c1[a] = 0.0;
            if (beta != 0.0) {
                if (vec_ok && ok0[a] && ok1[a]) {
                    double2 c = *reinterpret_cast<const double2*>(ccol + m);
                    c0[a] = c.x; c1[a] = c.y;
                } else {
                    if (ok0[a]) c0[a] = ccol[m];
                    if (ok1[a]) c1[a] = ccol[m + 1];
                }
            }
        }
#pragma unroll
        for (int a = 0; a < MT; ++a) {
            int m = mbase + a * 8;
            double v0 = alpha * acc[a][b][0], v1 = alpha * acc[a][b][1];
            if (beta != 0.0) { v0 += beta * c0[a]; v1 += beta * c1[a]; }
            if (vec_ok && ok0[a] && ok1[a]) {
                *reinterpret_cast<double2*>(ccol + m) = make_double2(v0, v1);
            } else {
                if (ok0[a]) ccol[m] = v0;
                if (ok1[a]) ccol[m + 1] = v1;
            }
        }
    }
}

// C := beta*C (full or one triangle); beta == 0 writes zeros without reading (dgemm.f:303-318).
__global__ void scale_matrix_kernel(int m, int n, double beta, double* C, i64 ldc, int tri) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    int j = blockIdx.y;
    if (i >= m) return;
    for (; j < n; j += gridDim.y) {
        if (tri == 1 && i < j) continue;
        if (tri == 2 && i > j) continue;
        double* c = C + i + (i64)j * ldc;
        *c = (beta == 0.0) ? 0.0 : beta * (*c);
    }
}

// C := alpha * sum_z P[z] + beta*C  (deterministic split-K reduction; P slices are dense m x n, ld = m)
__global__ void splitk_reduce_kernel(int m, int n, int nz, double alpha, double beta, const double* __restrict__ P,
                                     double* __restrict__ C, i64 ldc, int tri) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    for (int j = blockIdx.y; j < n; j += gridDim.y) {
        if (tri == 1 && i < j) continue;
        if (tri == 2 && i > j) continue;
        double acc = 0.0;
        const double* q = P + i + (i64)j * m;
        const i64 zs = (i64)m * n;
        int z = 0;
        for (; z + 8 <= nz; z += 8) {          // eight loads in flight, added in slice order
            double v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) v[u] = __ldcs(q + (i64)(z + u) * zs);
#pragma unroll
            for (int u = 0; u < 8; ++u) acc += v[u];
        }
        for (; z < nz; ++z) acc += __ldcs(q + (i64)z * zs);
        double* c = C + i + (i64)j * ldc;
        *c = (beta == 0.0) ? alpha * acc : alpha * acc + beta * (*c);
    }
}

static int g_splitk_balance = 1;   // wave-balancing split-K of long-K GEMMs (lb200_set_gemm_splitk_balance)
void gemm_set_splitk_balance(int on) { g_splitk_balance = on; }
static int g_gemm_cfg = -1;   // -1 auto, 0 = 128x128x8w, 1 = 128x64x4w, 2 = 64x64 (2 warps)

template <int BM, int BN, int WMW, int WNW, int STAGES, int BKT = 16, int VAR = 0>
static void launch_cfg(cudaStream_t s, bool a_k, bool b_k, bool al16, const GemmParams& p0) {
    GemmParams p = p0;
    p.tiles_m = ceil_div(p.M, BM);
    p.tiles_n = ceil_div(p.N, BN);
    constexpr int NTH = WMW * WNW * 32;
    size_t smem = 0;
    i64 nctas = (i64)p.tiles_m * p.tiles_n;
    if (p.tri != 0) {
        nctas = 0;
        for (int first_m = 0; first_m < p.tiles_m; first_m += TRI_GROUP) {
            const int gsz = min(p.tiles_m - first_m, TRI_GROUP);
            int c0, nc;
            tri_group_cols(p.tri, first_m, gsz, BM, BN, p.tiles_n, c0, nc);
            nctas += (i64)gsz * nc;
        }
        if (nctas == 0) return;
    }
    dim3 grid((unsigned)nctas, (unsigned)ceil_div(p.K, p.kchunk));
#define LB_LAUNCH(AK, BKM, AL)                                                                                  \
    {                                                                                                           \
        auto kern = gemm_f64_dmma_kernel<BM, BN, WMW, WNW, AK, BKM, AL, STAGES, BKT, VAR>;                             \
        smem = sizeof(double) * STAGES * (TileLayout<BM, AK, BKT>::ELEMS + TileLayout<BN, BKM, BKT>::ELEMS);     \
        static bool attr_set = false;                                                                           \
        if (!attr_set) {                                                                                        \
            LB_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));  \
            attr_set = true;                                                                                    \
        }                                                                                                       \
        kern<<<grid, NTH, smem, s>>>(p);                                                              \
    }
    if (a_k) {
        if (b_k) { if (al16) LB_LAUNCH(true, true, true) else LB_LAUNCH(true, true, false) }
        else     { if (al16) LB_LAUNCH(true, false, true) else LB_LAUNCH(true, false, false) }
    } else {
        if (b_k) { if (al16) LB_LAUNCH(false, true, true) else LB_LAUNCH(false, true, false) }
        else     { if (al16) LB_LAUNCH(false, false, true) else LB_LAUNCH(false, false, false) }
    }
#undef LB_LAUNCH
    count_launch();
    LB_CUDA_CHECK(cudaGetLastError());
}

void gemm_set_config(int cfg) { g_gemm_cfg = cfg; }
bool gemm_tma_try(cudaStream_t s, bool a_k, bool b_k, int m, int n, int k, double alpha, const double* A, i64 lda,
                  const double* B, i64 ldb, double beta, double* C, i64 ldc, int tri);
bool gemm_tma64_try(cudaStream_t s, bool a_k, bool b_k, int m, int n, int k, double alpha, const double* A, i64 lda,
                    const double* B, i64 ldb, double beta, double* C, i64 ldc, int tri);

// ---- optional per-launch timing of the large GEMMs (bench.py's roofline leg) ----------------------
struct GemmProf {
    bool on = false;
    std::vector<cudaEvent_t> e0, e1;
    std::vector<double> flops;
    size_t used = 0;
};
static GemmProf g_prof;
void gemm_profile(int enable) {
    g_prof.on = enable != 0;
    if (enable) g_prof.used = 0;
}
void gemm_profile_read(double* total_ms, double* total_flops, long long* launches) {
    double ms = 0.0, fl = 0.0;
    for (size_t i = 0; i < g_prof.used; ++i) {
        float t = 0.f;
        cudaEventSynchronize(g_prof.e1[i]);
        if (cudaEventElapsedTime(&t, g_prof.e0[i], g_prof.e1[i]) == cudaSuccess) { ms += t; fl += g_prof.flops[i]; }
    }
    *total_ms = ms; *total_flops = fl; *launches = (long long)g_prof.used;
}
static void gemm_impl(cudaStream_t s, char transa, char transb, int m, int n, int k, double alpha, const double* A, i64 lda,
                      const double* B, i64 ldb, double beta, double* C, i64 ldc, int tri);

void gemm(cudaStream_t s, char transa, char transb, int m, int n, int k, double alpha, const double* A, i64 lda,
          const double* B, i64 ldb, double beta, double* C, i64 ldc, int tri) {
    double fl = 2.0 * m * n * (double)k * (tri ? 0.5 : 1.0);
    bool prof = g_prof.on && fl >= 2e9 && g_prof.used < 16384;
    size_t idx = 0;
    if (prof) {
        idx = g_prof.used++;
        if (idx >= g_prof.e0.size()) {
            cudaEvent_t a, b;
            cudaEventCreate(&a); cudaEventCreate(&b);
            g_prof.e0.push_back(a); g_prof.e1.push_back(b); g_prof.flops.push_back(0.0);
        }
        g_prof.flops[idx] = fl;
        cudaEventRecord(g_prof.e0[idx], s);
    }
    gemm_impl(s, transa, transb, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, tri);
    if (prof) cudaEventRecord(g_prof.e1[idx], s);
}

static void gemm_impl(cudaStream_t s, char transa, char transb, int m, int n, int k, double alpha, const double* A, i64 lda,
                      const double* B, i64 ldb, double beta, double* C, i64 ldc, int tri) {
    if (m <= 0 || n <= 0) return;
    if (alpha == 0.0 || k <= 0) {
        if (beta == 1.0) return;
        dim3 grid(ceil_div(m, 256), (unsigned)min(n, 4096));
        scale_matrix_kernel<<<grid, 256, 0, s>>>(m, n, beta, C, ldc, tri);
        count_launch();
        return;
    }
    const bool ta = (transa == 'T' || transa == 't' || transa == 'C' || transa == 'c');
    const bool tb = (transb == 'T' || transb == 't' || transb == 'C' || transb == 'c');
    const bool a_k = ta;        // op(A)=A^T: A stored K x M, k contiguous
    const bool b_k = !tb;       // op(B)=B  : B stored K x N, k contiguous
    const bool al16 = ((((uintptr_t)A) & 15) == 0) && ((((uintptr_t)B) & 15) == 0) && ((lda & 1) == 0) &&
                      ((ldb & 1) == 0);
    GemmParams p;
    p.M = m; p.N = n; p.K = k; p.alpha = alpha; p.beta = beta;
    p.A = A; p.lda = lda; p.B = B; p.ldb = ldb; p.C = C; p.ldc = ldc; p.tri = tri;
    p.tiles_m = p.tiles_n = 0;
    p.kchunk = k;
    p.c_zstride = 0;
    p.guard = kernel_guard();
    int cfg = g_gemm_cfg;
    // Reduce-shaped products (few output tiles, long K: V^T*C in the QR panel, V^T*V) are split along K;
    // the slices go to scratch and are summed in a fixed order (deterministic).
    {
        i64 t64 = (i64)ceil_div(m, 64) * ceil_div(n, 64);
        if (t64 * 4 <= num_sms() && k >= 1024) {
            // slices of at least 256 in K, about two CTAs per SM in total
            int nz = (int)min((i64)(k / 256), (i64)(2 * num_sms()) / t64);
            if (nz >= 2) {
                int kchunk = ceil_div(ceil_div(k, nz), BK) * BK;
                nz = ceil_div(k, kchunk);
                double* P = (double*)ws_alloc(s, sizeof(double) * (size_t)m * n * nz);
                GemmParams q = p;
                q.alpha = 1.0; q.beta = 0.0; q.C = P; q.ldc = m; q.tri = 0;
                q.kchunk = kchunk; q.c_zstride = (i64)m * n;
                launch_cfg<64, 64, 1, 2, 4>(s, a_k, b_k, al16, q);
                dim3 rgrid(ceil_div(m, 128), (unsigned)min(n, 4096));
                splitk_reduce_kernel<<<rgrid, 128, 0, s>>>(m, n, nz, alpha, beta, P, C, ldc, tri);
                count_launch();
                ws_free(s, P);
                return;
            }
        }
    }
    // Long-K products with few output tiles per wave (W = V^T C of DLARFB: 8 tile rows, K = panel height): the launch is a handful
    // of waves of 4 x #SMs CTAs and the last, partly filled wave costs up to 15%.  Split K so that the CTA count fills whole waves;
    // slices are summed in a fixed order (deterministic).  K = NB updates of LU / Cholesky never take this path (k >= 2048).
    if (g_splitk_balance && k >= 2048 && (cfg < 0 || cfg == 13 || cfg == 14)) {
        const i64 t64 = (i64)ceil_div(m, 64) * ceil_div(n, 64);
        const i64 slots = (cfg == 14 ? 3 : 4) * (i64)num_sms();
        auto eff = [&](int z) {
            const i64 tot = t64 * z, wv = (tot + slots - 1) / slots;
            return (double)tot / (double)(wv * slots) - 0.004 * (z - 1);
        };
        int best = 1;
        double be = eff(1);
        if (be < 0.95 && t64 <= 16 * slots)
            for (int z = 2; z <= min(8, k / 512); ++z)
                if (eff(z) > be + 0.01) { best = z; be = eff(z); }
        if (best > 1) {
            int kchunk = ceil_div(ceil_div(k, best), BK) * BK;
            const int nz = ceil_div(k, kchunk);
            double* P = (double*)ws_alloc(s, sizeof(double) * (size_t)m * n * nz);
            GemmParams q = p;
            q.alpha = 1.0; q.beta = 0.0; q.C = P; q.ldc = m; q.tri = 0;
            q.kchunk = kchunk; q.c_zstride = (i64)m * n;
            if (cfg == 14) launch_cfg<64, 64, 2, 2, 3, 16, 1>(s, a_k, b_k, al16, q);
            else launch_cfg<64, 64, 2, 2, 2, 16, 1>(s, a_k, b_k, al16, q);
            dim3 rgrid(ceil_div(m, 128), (unsigned)min(n, 4096));
            splitk_reduce_kernel<<<rgrid, 128, 0, s>>>(m, n, nz, alpha, beta, P, C, ldc, tri);
            count_launch();
            ws_free(s, P);
            return;
        }
    }
    if (cfg == 3) {
        // persistent warp-specialised TMA kernel (gemm_tma.cu): one CTA per SM for the whole GEMM, so it must not
        // be used while a look-ahead panel needs SMs; measured 31.3 TFLOP/s at K=512 vs 33.6 for the default below.
        if (k >= 32 && gemm_tma_try(s, a_k, b_k, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, tri)) return;
        cfg = -1;
    }
    // default: 64x64 tiles, 4 warps of 32x32, 2-stage cp.async ring, 4 CTAs per SM, interior tiles through the lean loader
    // (measured on B200: 34.8 TFLOP/s at K=512, 35.6 at 8192^3, against 33.8 / 34.4 for the general loader alone -- cfg 8 -- and
    // 34.9 / 35.5 for cuBLAS; profiles/r02_gemm_lean_ab.txt)
    if (cfg == 12) {   // 64x64 TMA-fed kernel (same tile shape as cfg 8, operands moved by cp.async.bulk.tensor)
        if (k >= 16 && gemm_tma64_try(s, a_k, b_k, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, tri)) return;
        cfg = -1;
    }
    if (cfg < 0) cfg = 13;
    if (cfg == 4) launch_cfg<128, 64, 4, 2, 3>(s, a_k, b_k, al16, p);          // 8 warps of 32x32, 2 CTAs/SM
    else if (cfg == 5) launch_cfg<128, 128, 4, 4, 4>(s, a_k, b_k, al16, p);     // 16 warps of 32x32, 1 CTA/SM
    else if (cfg == 6) launch_cfg<64, 64, 2, 2, 4>(s, a_k, b_k, al16, p);       // 4 warps of 32x32, 2 CTAs/SM (smem)
    else if (cfg == 7) launch_cfg<64, 64, 2, 2, 3>(s, a_k, b_k, al16, p);       // 3 CTAs/SM
    else if (cfg == 8) launch_cfg<64, 64, 2, 2, 2>(s, a_k, b_k, al16, p);       // 4 CTAs/SM
    else if (cfg == 9) launch_cfg<64, 64, 2, 2, 3, 32>(s, a_k, b_k, al16, p);   // BK=32
    else if (cfg == 10) launch_cfg<64, 128, 2, 2, 3>(s, a_k, b_k, al16, p);     // 4 warps of 32x64
    else if (cfg == 11) launch_cfg<128, 64, 2, 2, 3, 32>(s, a_k, b_k, al16, p); // cfg1 with BK=32
    else if (cfg == 13) launch_cfg<64, 64, 2, 2, 2, 16, 1>(s, a_k, b_k, al16, p);   // cfg 8 + lean interior loader
    else if (cfg == 14) launch_cfg<64, 64, 2, 2, 3, 16, 1>(s, a_k, b_k, al16, p);   // same with 3 stages (3 CTAs/SM)
    else if (cfg == 0) launch_cfg<128, 128, 2, 4, 4>(s, a_k, b_k, al16, p);
    else if (cfg == 1) launch_cfg<128, 64, 2, 2, 3>(s, a_k, b_k, al16, p);
    else launch_cfg<64, 64, 1, 2, 4>(s, a_k, b_k, al16, p);
}

// DSYRK (BLAS/SRC/dsyrk.f:168): C := alpha*A*A^T + beta*C ('N') or alpha*A^T*A + beta*C ('T'), one triangle.
void syrk(cudaStream_t s, char uplo, char trans, int n, int k, double alpha, const double* A, i64 lda, double beta,
          double* C, i64 ldc) {
    const bool upper = (uplo == 'U' || uplo == 'u');
    const bool notr = (trans == 'N' || trans == 'n');
    if (notr) gemm(s, 'N', 'T', n, n, k, alpha, A, lda, A, lda, beta, C, ldc, upper ? 2 : 1);
    else gemm(s, 'T', 'N', n, n, k, alpha, A, lda, A, lda, beta, C, ldc, upper ? 2 : 1);
}

}  // namespace lb
