// gemm_tma.cu -- the fast path of the FP64 GEMM: persistent, warp-specialised, TMA-fed DMMA kernel.
//
// Same math and same fragment-to-thread mapping idea as gemm_f64.cu, different data movement:
//   * one producer warp issues cp.async.bulk.tensor (TMA, SASS UTMALDG) loads of the op(A)/op(B) tiles
//     into a 6-stage shared-memory ring, signalling per-stage "full" mbarriers with a transaction count;
//   * eight consumer warps (64x32 of C each, accumulators in registers) wait on "full", issue
//     LDS + DMMA.8x8x4, and release the stage through an "empty" mbarrier -- no __syncthreads in the loop,
//     no address arithmetic in the consumers, so the FP64 tensor pipe stays fed;
//   * the kernel is persistent (one CTA per SM, static tile schedule in grouped raster order): the producer
//     runs ahead into the next tile while the consumers are still in the read-modify-write epilogue.
// Shared-memory layout: TMA SWIZZLE_128B boxes.  An operand whose contiguous global dimension is m (or n)
// is fetched as [BK][16] boxes, one whose contiguous dimension is k as a single [rows][16] box; with the
// row permutations below every 8-byte fragment load of a half-warp hits 16 distinct bank pairs.
// Requirements: 16-byte aligned A/B base, even lda/ldb.  Anything else uses gemm_f64.cu.
#include "lb_internal.h"
#include <cuda.h>
#include <cstring>

namespace lb {

namespace tma {

constexpr int BM = 128, BN = 128, BK = 16, STAGES = 6;
constexpr int NCONS = 8;                       // consumer warps: 2 (m) x 4 (n), 64 x 32 each
constexpr int NPROD = 4;                       // producer warpgroup (one warp issues TMA; the group donates registers)
constexpr int NTHREADS = (NCONS + NPROD) * 32;
constexpr int A_BYTES = BM * BK * 8, B_BYTES = BN * BK * 8;
constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;

struct Params {
    int M, N, K;
    double alpha, beta;
    double* C; i64 ldc;
    int tri;
    int tiles_m, tiles_n;
};

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(unsigned dst, const CUtensorMap* map, int c0, int c1, unsigned bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];\n"
                 ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void dmma(double& d0, double& d1, double x, double y) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(d0), "+d"(d1) : "d"(x), "d"(y));
}

// ---- fragment addressing inside one operand stage (byte offsets; see the header comment) ----------
// mn-major operand: 16-wide chunks -> boxes of [BK][16]; fragment f (8 rows) = chunk f/2, half f%2.
//   row of lane group g inside the 16-chunk: perm(g) + 4*(f&1), perm(g) = (g&1) + 8*((g>>1)&1) + 2*(g>>2)
__device__ __forceinline__ int mn_frag_off(int frag, int g4, int t4) {
    int c0 = 4 * ((g4 >> 1) & 1) + (g4 >> 2) + 2 * (frag & 1);            // 16-byte chunk index before the swizzle
    return (frag >> 1) * (BK * 128) + t4 * 128 + ((c0 ^ t4) << 4) + (g4 & 1) * 8;
}
__device__ __forceinline__ int mn_frag_row(int frag, int g) {               // row (0..) of lane-group g in the tile
    return (frag >> 1) * 16 + 4 * (frag & 1) + (g & 1) + 8 * ((g >> 1) & 1) + 2 * (g >> 2);
}
// k-major operand: one box [rows][16 k]; fragment f = rows 8f..8f+7, lane group g -> row 8f + eperm(g)
__device__ __forceinline__ int eperm(int g) { return 2 * (g & 3) + (g >> 2); }
__device__ __forceinline__ int k_frag_off(int frag, int g4, int t4) {
    int e = 8 * frag + eperm(g4);
    return e * 128 + ((((t4 >> 1) ^ eperm(g4)) & 7) << 4) + (t4 & 1) * 8;
}
__device__ __forceinline__ int k_frag_row(int frag, int g) { return 8 * frag + eperm(g); }

template <bool A_KMAJ, bool B_KMAJ>
__global__ void __launch_bounds__(NTHREADS, 1)
    gemm_f64_tma_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, Params p) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    unsigned long long* bars = (unsigned long long*)(smem + STAGES * STAGE_BYTES);
    const unsigned full0 = smem_u32(bars), empty0 = smem_u32(bars + STAGES);
    const unsigned sbase = smem_u32(smem);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, NCONS); }
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    }
    __syncthreads();

    const int KT = (p.K + BK - 1) / BK;
    const int ntiles = p.tiles_m * p.tiles_n;
    constexpr int GROUP = 16;

    auto tile_coords = [&](int t, int& m0, int& n0) {
        int width = GROUP * p.tiles_n;
        int group_id = t / width;
        int first_m = group_id * GROUP;
        int gsz = min(p.tiles_m - first_m, GROUP);
        m0 = (first_m + (t % width) % gsz) * BM;
        n0 = ((t % width) / gsz) * BN;
    };
    auto tile_skipped = [&](int m0, int n0) {
        return (p.tri == 1 && m0 + BM - 1 < n0) || (p.tri == 2 && n0 + BN - 1 < m0);
    };

    if (warp < NPROD) {
        // ===================== producer warpgroup: give registers away, one elected lane issues TMA ==========
        asm volatile("setmaxnreg.dec.sync.aligned.u32 40;\n");
        if (warp == 0 && lane == 0) {
            int it = 0;
            for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
                int m0, n0;
                tile_coords(t, m0, n0);
                if (tile_skipped(m0, n0)) continue;
                for (int kt = 0; kt < KT; ++kt, ++it) {
                    const int s = it % STAGES;
                    const unsigned ph = (it / STAGES) & 1;
                    mbar_wait(empty0 + 8 * s, ph ^ 1);
                    mbar_expect_tx(full0 + 8 * s, STAGE_BYTES);
                    const unsigned sa = sbase + s * STAGE_BYTES, sb = sa + A_BYTES;
                    const int k0 = kt * BK;
                    if (A_KMAJ) tma_load_2d(sa, &mapA, k0, m0, full0 + 8 * s);
                    else {
#pragma unroll
                        for (int b = 0; b < BM / 16; ++b) tma_load_2d(sa + b * (BK * 128), &mapA, m0 + 16 * b, k0, full0 + 8 * s);
                    }
                    if (B_KMAJ) tma_load_2d(sb, &mapB, k0, n0, full0 + 8 * s);
                    else {
#pragma unroll
                        for (int b = 0; b < BN / 16; ++b) tma_load_2d(sb + b * (BK * 128), &mapB, n0 + 16 * b, k0, full0 + 8 * s);
                    }
                }
            }
        }
        return;
    }

    // ===================== consumer warps (two warpgroups) =====================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 232;\n");
    const int cwarp = warp - NPROD;
    const int ctid = tid - NPROD * 32;
    const int warp_m = cwarp & 1, warp_n = cwarp >> 1;
    const int g4 = lane >> 2, t4 = lane & 3;
    constexpr int MT = 8, NT = 4;

    int aoff[MT], boff[NT];
#pragma unroll
    for (int a = 0; a < MT; ++a) aoff[a] = A_KMAJ ? k_frag_off(warp_m * 8 + a, g4, t4) : mn_frag_off(warp_m * 8 + a, g4, t4);
#pragma unroll
    for (int b = 0; b < NT; ++b) boff[b] = A_BYTES + (B_KMAJ ? k_frag_off(warp_n * 4 + b, g4, t4) : mn_frag_off(warp_n * 4 + b, g4, t4));

    int it = 0;
    for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
        int m0, n0;
        tile_coords(t, m0, n0);
        if (tile_skipped(m0, n0)) continue;

        if (p.beta != 0.0) {   // pull the C tile towards L2 for the read-modify-write epilogue
            for (int l = ctid; l < BN * (BM / 16); l += NCONS * 32) {
                int n = n0 + l / (BM / 16), m = m0 + (l % (BM / 16)) * 16;
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

        for (int kt = 0; kt < KT; ++kt, ++it) {
            const int s = it % STAGES;
            const unsigned ph = (it / STAGES) & 1;
            mbar_wait(full0 + 8 * s, ph);
            const unsigned char* st = smem + s * STAGE_BYTES;
#pragma unroll
            for (int kk = 0; kk < BK; kk += 4) {
                double af[MT], bf[NT];
#pragma unroll
                for (int a = 0; a < MT; ++a) {
                    int off = A_KMAJ ? (aoff[a] ^ (kk * 8)) : ((aoff[a] + kk * 128) ^ ((kk & 4) << 4));
                    af[a] = *reinterpret_cast<const double*>(st + off);
                }
#pragma unroll
                for (int b = 0; b < NT; ++b) {
                    int off = B_KMAJ ? (boff[b] ^ (kk * 8)) : ((boff[b] + kk * 128) ^ ((kk & 4) << 4));
                    bf[b] = *reinterpret_cast<const double*>(st + off);
                }
#pragma unroll
                for (int a = 0; a < MT; ++a)
#pragma unroll
                    for (int b = 0; b < NT; ++b) dmma(acc[a][b][0], acc[a][b][1], bf[b], af[a]);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(empty0 + 8 * s);
        }

        // ---- epilogue: C = alpha*acc + beta*C, loads of one column batch issued before its stores ----
        const bool vec_ok = !A_KMAJ && ((p.ldc & 1) == 0) && ((((uintptr_t)p.C) & 15) == 0);
        const double alpha = p.alpha, beta = p.beta;
#pragma unroll
        for (int b = 0; b < NT; ++b) {
            const int n = n0 + (B_KMAJ ? k_frag_row(warp_n * 4 + b, g4) : mn_frag_row(warp_n * 4 + b, g4));
            if (n >= p.N) continue;
            double* ccol = p.C + (i64)n * p.ldc;
            double c0[MT], c1[MT];
            int r0[MT], r1[MT];
            bool ok0[MT], ok1[MT];
#pragma unroll
            for (int a = 0; a < MT; ++a) {
                r0[a] = m0 + (A_KMAJ ? k_frag_row(warp_m * 8 + a, 2 * t4) : mn_frag_row(warp_m * 8 + a, 2 * t4));
                r1[a] = m0 + (A_KMAJ ? k_frag_row(warp_m * 8 + a, 2 * t4 + 1) : mn_frag_row(warp_m * 8 + a, 2 * t4 + 1));
                ok0[a] = r0[a] < p.M;
                ok1[a] = r1[a] < p.M;
                if (p.tri == 1) { ok0[a] = ok0[a] && (r0[a] >= n); ok1[a] = ok1[a] && (r1[a] >= n); }
                if (p.tri == 2) { ok0[a] = ok0[a] && (r0[a] <= n); ok1[a] = ok1[a] && (r1[a] <= n); }
                c0[a] = c1[a] = 0.0;
                if (beta != 0.0) {
                    if (vec_ok && ok0[a] && ok1[a]) {
                        double2 c = *reinterpret_cast<const double2*>(ccol + r0[a]);
                        c0[a] = c.x; c1[a] = c.y;
                    } else {
                        if (ok0[a]) c0[a] = ccol[r0[a]];
                        if (ok1[a]) c1[a] = ccol[r1[a]];
                    }
                }
            }
#pragma unroll
            for (int a = 0; a < MT; ++a) {
                double v0 = alpha * acc[a][b][0], v1 = alpha * acc[a][b][1];
                if (beta != 0.0) { v0 += beta * c0[a]; v1 += beta * c1[a]; }
                if (vec_ok && ok0[a] && ok1[a]) {
                    *reinterpret_cast<double2*>(ccol + r0[a]) = make_double2(v0, v1);
                } else {
                    if (ok0[a]) ccol[r0[a]] = v0;
                    if (ok1[a]) ccol[r1[a]] = v1;
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Small-tile TMA kernel: 64 x 64 tile, 4 warps of 32 x 32, 3-stage ring, 4 CTAs per SM, NOT persistent.
// This is the shape that measured best on B200 for the trailing updates (many small CTAs keep the DMMA pipe
// busy across each other's barriers and epilogues, and retire continuously so that a high-priority look-ahead
// panel can get SMs).  Thread 0 issues the TMA loads for the stage that has just been consumed; everybody
// waits on the stage's mbarrier.  No per-thread address arithmetic, no cp.async bookkeeping.
namespace t64 {
constexpr int TM = 64, TN = 64, TSTAGES = 3, TTHREADS = 128;
constexpr int TA_BYTES = TM * BK * 8, TB_BYTES = TN * BK * 8, TSTAGE_BYTES = TA_BYTES + TB_BYTES;
constexpr int TSMEM_BYTES = TSTAGES * TSTAGE_BYTES + 1024 + 64;
}  // namespace t64

template <bool A_KMAJ, bool B_KMAJ>
__global__ void __launch_bounds__(t64::TTHREADS, 4)
    gemm_f64_tma64_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, Params p) {
    using namespace t64;
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    unsigned long long* bars = (unsigned long long*)(smem + TSTAGES * TSTAGE_BYTES);
    const unsigned full0 = smem_u32(bars);
    const unsigned sbase = smem_u32(smem);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    constexpr int GROUP = 16;
    const int t = blockIdx.x;
    const int width = GROUP * p.tiles_n;
    const int group_id = t / width;
    const int first_m = group_id * GROUP;
    const int gsz = min(p.tiles_m - first_m, GROUP);
    const int m0 = (first_m + (t % width) % gsz) * TM;
    const int n0 = ((t % width) / gsz) * TN;
    if ((p.tri == 1 && m0 + TM - 1 < n0) || (p.tri == 2 && n0 + TN - 1 < m0)) return;

    if (tid == 0) {
        for (int s = 0; s < TSTAGES; ++s) mbar_init(full0 + 8 * s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    }
    __syncthreads();

    const int KT = (p.K + BK - 1) / BK;
    auto issue = [&](int kt) {
        const int s = kt % TSTAGES;
        const unsigned bar = full0 + 8 * s;
        mbar_expect_tx(bar, TSTAGE_BYTES);
        const unsigned sa = sbase + s * TSTAGE_BYTES, sb = sa + TA_BYTES;
        const int k0 = kt * BK;
        if (A_KMAJ) tma_load_2d(sa, &mapA, k0, m0, bar);
        else {
#pragma unroll
            for (int b = 0; b < TM / 16; ++b) tma_load_2d(sa + b * (BK * 128), &mapA, m0 + 16 * b, k0, bar);
        }
        if (B_KMAJ) tma_load_2d(sb, &mapB, k0, n0, bar);
        else {
#pragma unroll
            for (int b = 0; b < TN / 16; ++b) tma_load_2d(sb + b * (BK * 128), &mapB, n0 + 16 * b, k0, bar);
        }
    };
    if (tid == 0) {
        for (int kt = 0; kt < TSTAGES && kt < KT; ++kt) issue(kt);
    }
    if (p.beta != 0.0) {   // pull the C tile towards L2 for the read-modify-write epilogue
        for (int l = tid; l < TN * (TM / 16); l += TTHREADS) {
            int n = n0 + l / (TM / 16), m = m0 + (l % (TM / 16)) * 16;
            if (n < p.N && m < p.M) {
                const double* ptr = p.C + (i64)n * p.ldc + m;
                asm volatile("prefetch.global.L2 [%0];\n" ::"l"(ptr));
            }
        }
    }

    const int warp_m = warp & 1, warp_n = warp >> 1;
    const int g4 = lane >> 2, t4 = lane & 3;
    constexpr int MT = 4, NT = 4;
    int aoff[MT], boff[NT];
#pragma unroll
    for (int a = 0; a < MT; ++a) aoff[a] = A_KMAJ ? k_frag_off(warp_m * 4 + a, g4, t4) : mn_frag_off(warp_m * 4 + a, g4, t4);
#pragma unroll
    for (int b = 0; b < NT; ++b) boff[b] = TA_BYTES + (B_KMAJ ? k_frag_off(warp_n * 4 + b, g4, t4) : mn_frag_off(warp_n * 4 + b, g4, t4));

    double acc[MT][NT][2];
#pragma unroll
    for (int a = 0; a < MT; ++a)
#pragma unroll
        for (int b = 0; b < NT; ++b) acc[a][b][0] = acc[a][b][1] = 0.0;

    for (int kt = 0; kt < KT; ++kt) {
        const int s = kt % TSTAGES;
        mbar_wait(full0 + 8 * s, (kt / TSTAGES) & 1);
        const unsigned char* st = smem + s * TSTAGE_BYTES;
#pragma unroll
        for (int kk = 0; kk < BK; kk += 4) {
            double af[MT], bf[NT];
#pragma unroll
            for (int a = 0; a < MT; ++a) {
                int off = A_KMAJ ? (aoff[a] ^ (kk * 8)) : ((aoff[a] + kk * 128) ^ ((kk & 4) << 4));
                af[a] = *reinterpret_cast<const double*>(st + off);
            }
#pragma unroll
            for (int b = 0; b < NT; ++b) {
                int off = B_KMAJ ? (boff[b] ^ (kk * 8)) : ((boff[b] + kk * 128) ^ ((kk & 4) << 4));
                bf[b] = *reinterpret_cast<const double*>(st + off);
            }
#pragma unroll
            for (int a = 0; a < MT; ++a)
#pragma unroll
                for (int b = 0; b < NT; ++b) dmma(acc[a][b][0], acc[a][b][1], bf[b], af[a]);
        }
        __syncthreads();                                   // the stage has been read by all four warps
        if (tid == 0 && kt + TSTAGES < KT) issue(kt + TSTAGES);
    }

    const bool vec_ok = !A_KMAJ && ((p.ldc & 1) == 0) && ((((uintptr_t)p.C) & 15) == 0);
    const double alpha = p.alpha, beta = p.beta;
#pragma unroll
    for (int b = 0; b < NT; ++b) {
        const int n = n0 + (B_KMAJ ? k_frag_row(warp_n * 4 + b, g4) : mn_frag_row(warp_n * 4 + b, g4));
        if (n >= p.N) continue;
        double* ccol = p.C + (i64)n * p.ldc;
        double c0[MT], c1[MT];
        int r0[MT], r1[MT];
        bool ok0[MT], ok1[MT];
#pragma unroll
        for (int a = 0; a < MT; ++a) {
            r0[a] = m0 + (A_KMAJ ? k_frag_row(warp_m * 4 + a, 2 * t4) : mn_frag_row(warp_m * 4 + a, 2 * t4));
            r1[a] = m0 + (A_KMAJ ? k_frag_row(warp_m * 4 + a, 2 * t4 + 1) : mn_frag_row(warp_m * 4 + a, 2 * t4 + 1));
            ok0[a] = r0[a] < p.M;
            ok1[a] = r1[a] < p.M;
            if (p.tri == 1) { ok0[a] = ok0[a] && (r0[a] >= n); ok1[a] = ok1[a] && (r1[a] >= n); }
            if (p.tri == 2) { ok0[a] = ok0[a] && (r0[a] <= n); ok1[a] = ok1[a] && (r1[a] <= n); }
            c0[a] = c1[a] = 0.0;
            if (beta != 0.0) {
                if (vec_ok && ok0[a] && ok1[a]) {
                    double2 c = *reinterpret_cast<const double2*>(ccol + r0[a]);
                    c0[a] = c.x; c1[a] = c.y;
                } else {
                    if (ok0[a]) c0[a] = ccol[r0[a]];
                    if (ok1[a]) c1[a] = ccol[r1[a]];
                }
            }
        }
#pragma unroll
        for (int a = 0; a < MT; ++a) {
            double v0 = alpha * acc[a][b][0], v1 = alpha * acc[a][b][1];
            if (beta != 0.0) { v0 += beta * c0[a]; v1 += beta * c1[a]; }
            if (vec_ok && ok0[a] && ok1[a]) {
                *reinterpret_cast<double2*>(ccol + r0[a]) = make_double2(v0, v1);
            } else {
                if (ok0[a]) ccol[r0[a]] = v0;
                if (ok1[a]) ccol[r1[a]] = v1;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
        else (void)cudaGetLastError();
    }
    return fn;
}

// operand with `ext` rows/cols along m (or n) and K along k.  kmaj: k contiguous (element (e,k) at k + e*ld)
static bool make_map(CUtensorMap* map, const double* base, bool kmaj, int ext, int K, i64 ld, int box_ext) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return false;
    cuuint64_t dims[2], strides[1];
    cuuint32_t box[2], estr[2] = {1, 1};
    if (kmaj) { dims[0] = (cuuint64_t)K; dims[1] = (cuuint64_t)ext; box[0] = 16; box[1] = (cuuint32_t)box_ext; }
    else { dims[0] = (cuuint64_t)ext; dims[1] = (cuuint64_t)K; box[0] = 16; box[1] = BK; }
    strides[0] = (cuuint64_t)ld * 8;
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, (void*)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

}  // namespace tma

// 64x64 non-persistent TMA kernel (default for aligned operands)
bool gemm_tma64_try(cudaStream_t s, bool a_k, bool b_k, int m, int n, int k, double alpha, const double* A, i64 lda,
                    const double* B, i64 ldb, double beta, double* C, i64 ldc, int tri) {
    using namespace tma;
    using namespace tma::t64;
    if ((((uintptr_t)A) & 15) || (((uintptr_t)B) & 15) || (lda & 1) || (ldb & 1)) return false;
    if (lda * 8 >= (1LL << 40) || ldb * 8 >= (1LL << 40)) return false;
    CUtensorMap mapA, mapB;
    if (!make_map(&mapA, A, a_k, m, k, lda, TM)) return false;
    if (!make_map(&mapB, B, b_k, n, k, ldb, TN)) return false;
    Params p;
    p.M = m; p.N = n; p.K = k; p.alpha = alpha; p.beta = beta; p.C = C; p.ldc = ldc; p.tri = tri;
    p.tiles_m = ceil_div(m, TM); p.tiles_n = ceil_div(n, TN);
    const i64 grid = (i64)p.tiles_m * p.tiles_n;
#define LB_TMA64_LAUNCH(AK, BKM)                                                                                     \
    {                                                                                                                \
        auto kern = gemm_f64_tma64_kernel<AK, BKM>;                                                                  \
        static bool attr_set = false;                                                                                \
        if (!attr_set) {                                                                                             \
            LB_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, TSMEM_BYTES));     \
            attr_set = true;                                                                                         \
        }                                                                                                            \
        kern<<<(unsigned)grid, TTHREADS, TSMEM_BYTES, s>>>(mapA, mapB, p);                                           \
    }
    if (a_k) { if (b_k) LB_TMA64_LAUNCH(true, true) else LB_TMA64_LAUNCH(true, false) }
    else     { if (b_k) LB_TMA64_LAUNCH(false, true) else LB_TMA64_LAUNCH(false, false) }
#undef LB_TMA64_LAUNCH
    count_launch();
    LB_CUDA_CHECK(cudaGetLastError());
    return true;
}

static int g_tma_enabled = 1;
void gemm_set_tma(int on) { g_tma_enabled = on; }

// returns false if the problem is not eligible (caller falls back to the cp.async kernel)
bool gemm_tma_try(cudaStream_t s, bool a_k, bool b_k, int m, int n, int k, double alpha, const double* A, i64 lda,
                  const double* B, i64 ldb, double beta, double* C, i64 ldc, int tri) {
    using namespace tma;
    if (!g_tma_enabled) return false;
    if ((((uintptr_t)A) & 15) || (((uintptr_t)B) & 15) || (lda & 1) || (ldb & 1)) return false;
    if (lda * 8 >= (1LL << 40) || ldb * 8 >= (1LL << 40)) return false;
    CUtensorMap mapA, mapB;
    if (!make_map(&mapA, A, a_k, m, k, lda, BM)) return false;
    if (!make_map(&mapB, B, b_k, n, k, ldb, BN)) return false;
    Params p;
    p.M = m; p.N = n; p.K = k; p.alpha = alpha; p.beta = beta; p.C = C; p.ldc = ldc; p.tri = tri;
    p.tiles_m = ceil_div(m, BM); p.tiles_n = ceil_div(n, BN);
    const int ntiles = p.tiles_m * p.tiles_n;
    const int grid = min(ntiles, num_sms());
#define LB_TMA_LAUNCH(AK, BKM)                                                                                       \
    {                                                                                                                \
        auto kern = gemm_f64_tma_kernel<AK, BKM>;                                                                    \
        static bool attr_set = false;                                                                                \
        if (!attr_set) {                                                                                             \
            LB_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));      \
            attr_set = true;                                                                                         \
        }                                                                                                            \
        kern<<<grid, NTHREADS, SMEM_BYTES, s>>>(mapA, mapB, p);                                                      \
    }
    if (a_k) { if (b_k) LB_TMA_LAUNCH(true, true) else LB_TMA_LAUNCH(true, false) }
    else     { if (b_k) LB_TMA_LAUNCH(false, true) else LB_TMA_LAUNCH(false, false) }
#undef LB_TMA_LAUNCH
    count_launch();
    LB_CUDA_CHECK(cudaGetLastError());
    return true;
}

}  // namespace lb
