// geqrf.cu -- Householder QR on one B200: DGEQRF / DGEQR2 / DLARFT / DLARFB.
//
// Reference path: SRC/dgeqrf.f:180-278 (blocked driver), SRC/dgeqr2.f:169-184 (unblocked panel),
// SRC/dlarfg.f:140-186 (one reflector), SRC/dlarf1f.f:192-290 (apply), SRC/dlarft.f:207-349 and
// SRC/dlarft_lvl2.f:199-258 (triangular factor T), SRC/dlarfb.f:231-345 (block reflector, Forward/Columnwise).
//
// B200 design.
//  * Outer block NB (default 256; the reference's 32 gives K=32 GEMMs that sit on the HBM ridge).  The
//    trailing update C := (I - V T^T V^T) C is three DMMA GEMMs on an explicit "clean" copy Vc of the
//    reflectors (unit diagonal, zeros above), so no triangular special cases are needed:
//        W = Vc^T C   (reduce-shaped, split-K when narrow),  W2 = T^T W,  C -= Vc W2.
//  * The panel is factored recursively (the Elmroth-Gustavson recursion the reference itself uses in
//    dlarft.f / dgeqrt3.f): left half, apply to right half, right half, T12 = -T1 (V1^T V2) T2.
//  * Leaf kernel (geqr2_leaf_kernel, W=16 columns): cooperative multi-CTA kernel, one row per thread held
//    in registers.  Per column ONE fused reduction (sum x^2, x^T C for the remaining columns and the
//    V^T v dot products needed for T) and ONE grid barrier; then every CTA forms beta/tau exactly as
//    DLARFG does (DLAPY2 form) and applies the reflector to its rows.  Memory-bound: 16*M*W bytes per leaf.
//  * Over/underflow safety (DNRM2's scaled accumulators, DLARFG's rescale loop) is obtained by an exact
//    power-of-two pre-scaling of the whole matrix when max|A| is outside [2^-400, 2^400]; V and tau are
//    invariant under it and R is scaled back.
#include "lb_internal.h"
#include <cfloat>
#include <mutex>
#include <cstring>

namespace lb {

static inline double __longlong_as_double_host(unsigned long long v) { double d; memcpy(&d, &v, 8); return d; }

void expand_tri(cudaStream_t s, int n, const double* A, i64 lda, bool upper, bool unit, double* T, i64 ldt);

static int g_qr_nb = 256, g_qr_lookahead = 1;
void geqrf_set_params(int nb, int lookahead) {
    if (nb > 0) g_qr_nb = nb;
    if (lookahead >= 0) g_qr_lookahead = lookahead;
}

constexpr int QW = 16;          // leaf width
constexpr int QTHREADS = 1024;  // rows per CTA

struct QrLeafParams {
    int m, n;
    double* A; i64 lda;
    double* tau;
    double* Vc; i64 ldvc;      // clean reflectors (m x n)
    double* T; i64 ldt;        // leaf T block (n x n upper triangular), lower part untouched
    unsigned* bar; unsigned bar_base;
    double* part;              // [2][G][QW]
    double* toprow;            // [2][QW]
    int G;
};

__device__ __forceinline__ void qr_grid_barrier(unsigned* bar, unsigned target, int G) {
    __syncthreads();
    if (G > 1) {
        if (threadIdx.x == 0) {
            __threadfence();
            atomicAdd(bar, 1u);
            unsigned v;
            do {
                asm volatile("ld.acquire.gpu.global.u32 %0, [%1];\n" : "=r"(v) : "l"(bar) : "memory");
            } while ((int)(v - target) < 0);
            __threadfence();
        }
        __syncthreads();
    }
}

// SRC/dlapy2.f:96-112 for finite inputs
__device__ __forceinline__ double dev_dlapy2(double x, double y) {
    double xa = fabs(x), ya = fabs(y);
    double w = fmax(xa, ya), z = fmin(xa, ya);
    if (z == 0.0 || w > DBL_MAX) return w;
    double q = z / w;
    return w * sqrt(1.0 + q * q);
}

template <int W, int THREADS>
__global__ void __launch_bounds__(THREADS, 1) geqr2_leaf_kernel(QrLeafParams p) {
    constexpr int NWARP = THREADS / 32;
    __shared__ double s_red[NWARP][W];
    __shared__ double s_tot[W];
    __shared__ double s_trow[W];
    __shared__ double s_G[W][W];      // s_G[i][c] = V(:,i)^T v_c, i < c
    __shared__ double s_tau[W];
    __shared__ double s_T[W][W];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = blockIdx.x;
    const int row = g * THREADS + tid;
    const bool have = row < p.m;
    const int kmax = min(p.m, p.n);

    double a[W];
#pragma unroll
    for (int c = 0; c < W; ++c) a[c] = (have && c < p.n) ? p.A[row + (i64)c * p.lda] : 0.0;

    // step c = 0..kmax-1: reflector c; step c = kmax: only the trailing V^T v reduction for column kmax-1
#pragma unroll
    for (int c = 0; c <= W; ++c) {
        if (c <= kmax) {
            const int slot = c & 1;
            double red[W];
#pragma unroll
            for (int q = 0; q < W; ++q) red[q] = 0.0;
            if (c < kmax && c < W) {
                if (have && row > c) {
                    red[0] = a[c < W ? c : 0] * a[c < W ? c : 0];
#pragma unroll
                    for (int q = c + 1; q < W; ++q) red[1 + (q - c - 1)] = a[c < W ? c : 0] * a[q];
                }
            }
            if (c >= 1) {
                // dots of the finished reflector pc = c-1 with the earlier ones (unit diagonal at row pc)
                const int pc = c - 1;
                if (have && row >= pc) {
#pragma unroll
                    for (int i = 0; i + 1 < c; ++i) {   // i in 0..pc-1
                        double vi = a[i];
                        red[W - c + i] = (row == pc) ? vi : vi * a[pc];
                    }
                }
            }
            // block reduction of the W partial sums
#pragma unroll
            for (int q = 0; q < W; ++q) {
                double v = red[q];
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
                if (lane == 0) s_red[warp][q] = v;
            }
            __syncthreads();
            if (tid < W) {
                double v = 0.0;
#pragma unroll 8
                for (int w = 0; w < NWARP; ++w) v += s_red[w][tid];
                p.part[((i64)slot * p.G + g) * W + tid] = v;
            }
            if (c < kmax && have && row == c) {
                double* dst = p.toprow + slot * W;
#pragma unroll
                for (int q = 0; q < W; ++q) dst[q] = a[q];
            }
            qr_grid_barrier(p.bar, p.bar_base + (unsigned)(c + 1) * (unsigned)p.G, p.G);
            if (tid < W) {
                double v = 0.0;
                for (int q = 0; q < p.G; ++q) v += __ldcg(p.part + ((i64)slot * p.G + q) * W + tid);
                s_tot[tid] = v;
                if (c < kmax) s_trow[tid] = __ldcg(p.toprow + slot * W + tid);
            }
            __syncthreads();
            if (c >= 1) {
                const int pc = c - 1;
                if (tid < pc) s_G[tid][pc] = s_tot[W - c + tid];
            }
            if (c < kmax && c < W) {
                // DLARFG (dlarfg.f:140-186) with xnorm^2 = s_tot[0]
                const double alpha = s_trow[c];
                const double xnorm = sqrt(s_tot[0]);
                double tau = 0.0, beta = alpha, scale = 0.0;
                if (xnorm != 0.0) {
                    beta = -copysign(dev_dlapy2(alpha, xnorm), alpha);
                    tau = (beta - alpha) / beta;
                    scale = 1.0 / (alpha - beta);
                }
                if (tid == 0) {
                    s_tau[c] = tau;
                    if (g == 0) p.tau[c] = tau;
                }
                if (tau != 0.0) {
                    if (have && row > c) {
                        const double v = a[c] * scale;           // DSCAL by 1/(alpha-beta), dlarfg.f:178
                        a[c] = v;
#pragma unroll
                        for (int q = c + 1; q < W; ++q) {
                            // w(q) = C(1,q) + v2^T C2(:,q)  (dlarf1f.f:247-250);  C -= tau*v*w^T (dlarf1f.f:256-258)
                            const double wq = s_trow[q] + scale * s_tot[1 + (q - c - 1)];
                            a[q] = fma(-tau * wq, v, a[q]);
                        }
                    } else if (have && row == c) {
                        a[c] = beta;
#pragma unroll
                        for (int q = c + 1; q < W; ++q) {
                            const double wq = s_trow[q] + scale * s_tot[1 + (q - c - 1)];
                            a[q] = a[q] - tau * wq;
                        }
                    }
                }
            }
            __syncthreads();
        }
    }

    // T for the leaf (dlarft_lvl2.f:199-258): T(0:c,c) = T(0:c,0:c) * (-tau_c * G(0:c,c)), T(c,c) = tau_c
    if (g == 0) {
        if (warp == 0) {
            for (int c = 0; c < kmax; ++c) {
                const double tc = s_tau[c];
                double tmp = (lane < c) ? -tc * s_G[lane][c] : 0.0;
                // upper-triangular matvec: out[i] = sum_{j=i..c-1} T[i][j]*tmp[j]
                double out = 0.0;
                for (int j = 0; j < c; ++j) {
                    double tj = __shfl_sync(0xffffffffu, tmp, j);
                    if (lane <= j && lane < c) out += s_T[lane][j] * tj;
                }
                if (lane < c) s_T[lane][c] = (tc == 0.0) ? 0.0 : out;
                if (lane == c) s_T[c][c] = tc;
                __syncwarp();
            }
            for (int idx = lane; idx < kmax * kmax; idx += 32) {
                int i = idx % kmax, j = idx / kmax;
                if (i <= j) p.T[i + (i64)j * p.ldt] = s_T[i][j];
            }
        }
    }
    if (have) {
#pragma unroll
        for (int c = 0; c < W; ++c) {
            if (c < p.n) {
                p.A[row + (i64)c * p.lda] = a[c];
                if (p.Vc) {
                    double v = a[c];
                    if (c >= kmax || row < c) v = 0.0;
                    else if (row == c) v = 1.0;
                    p.Vc[row + (i64)c * p.ldvc] = v;
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Cluster leaf (panels of at most 16 x 1024 rows): the CTAs form one thread-block cluster and the per-column
// reduction travels through distributed shared memory + the hardware cluster barrier.  R rows per thread
// (row = g*R*THREADS + r*THREADS + tid); the W-column window rotates so that the active column is a[.][0] and the
// column loop exists once in the instruction stream.  At step c the window holds, at positions
//   0          : column c (the vector being reduced),
//   1..live-1  : columns c+1..W-1 (to be updated),            live = W - c
//   live..W-1  : the finished reflectors v_0..v_{c-1} (position W-1 = v_{c-1}).
// One fused reduction per column gives  sum x^2,  x^T C(:,q)  and the dots  v_i^T v_{c-1}  that DLARFT needs.
__device__ __forceinline__ void qcl_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory"); }
__device__ __forceinline__ void qcl_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory"); }
__device__ __forceinline__ unsigned qcl_smem(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ double qcl_ld(unsigned addr, unsigned rank) {
    unsigned ra;
    double v;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;\n" : "=r"(ra) : "r"(addr), "r"(rank));
    asm volatile("ld.shared::cluster.f64 %0, [%1];\n" : "=d"(v) : "r"(ra) : "memory");
    return v;
}

// Sum 16 per-lane values over the 32 lanes of a warp with 16 shuffles instead of 80: every round each lane gives
// away half of its values.  On return v[0] in lane l is the warp total of value index l >> 1.
__device__ __forceinline__ void warp_reduce16(double (&v)[16], int lane) {
    const unsigned full = 0xffffffffu;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const bool hi = lane & 16;
        const double send = hi ? v[j] : v[j + 8], keep = hi ? v[j + 8] : v[j];
        v[j] = keep + __shfl_xor_sync(full, send, 16);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const bool hi = lane & 8;
        const double send = hi ? v[j] : v[j + 4], keep = hi ? v[j + 4] : v[j];
        v[j] = keep + __shfl_xor_sync(full, send, 8);
    }
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        const bool hi = lane & 4;
        const double send = hi ? v[j] : v[j + 2], keep = hi ? v[j + 2] : v[j];
        v[j] = keep + __shfl_xor_sync(full, send, 4);
    }
    {
        const bool hi = lane & 2;
        const double send = hi ? v[0] : v[1], keep = hi ? v[1] : v[0];
        v[0] = keep + __shfl_xor_sync(full, send, 2);
    }
    v[0] += __shfl_xor_sync(full, v[0], 1);
}

template <int W, int THREADS, int R, bool CLUSTER>
__global__ void __launch_bounds__(THREADS, 1) geqr2_leaf_cluster_kernel(QrLeafParams p) {
    static_assert(W == 16, "warp_reduce16 handles 16 values");
    constexpr int NWARP = THREADS / 32;
    constexpr int ROWS = THREADS * R;
    __shared__ double s_red[NWARP][W];
    __shared__ __align__(16) double s_part[2][W];
    __shared__ __align__(16) double s_top[2][W];
    __shared__ double s_tot[W];
    __shared__ double s_trow[W];
    __shared__ double s_hh[4];        // tau, beta, scale of the current reflector
    __shared__ double s_G[W][W];      // s_G[i][c] = V(:,i)^T v_c, i < c
    __shared__ double s_tau[W];
    __shared__ double s_T[W][W];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = blockIdx.x;
    const int C = p.G;
    const int row0 = g * ROWS + tid;
    const int kmax = min(p.m, p.n);
    const unsigned part_base = qcl_smem(&s_part[0][0]);
    const unsigned top_base = qcl_smem(&s_top[0][0]);

    double a[R][W];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const int row = row0 + r * THREADS;
#pragma unroll
        for (int q = 0; q < W; ++q) a[r][q] = (row < p.m && q < p.n) ? p.A[row + (i64)q * p.lda] : 0.0;
    }

    // step c = 0..kmax-1: reflector c; step c = kmax: only the v_i^T v_{kmax-1} dots
#pragma unroll 1
    for (int c = 0; c <= kmax; ++c) {
        const int slot = c & 1;
        const int live = W - c;
        const int pc = c - 1;
        const bool main_step = c < kmax;
        double red[W];
#pragma unroll
        for (int q = 0; q < W; ++q) red[q] = 0.0;
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int row = row0 + r * THREADS;
            if (row < p.m) {
                const double x = (main_step && row > c) ? a[r][0] : 0.0;
                // finished reflector v_pc sits at position W-1 (unit diagonal at row pc, zero above)
                const double vp = (c >= 1 && row >= pc) ? ((row == pc) ? 1.0 : a[r][W - 1]) : 0.0;
                red[0] = fma(main_step ? x : vp, a[r][0], red[0]);   // x^2, or v_0^T v_pc at the extra step of a full leaf
#pragma unroll
                for (int q = 1; q < W; ++q) red[q] = fma((q < live) ? x : vp, a[r][q], red[q]);   // q = W-1 is live only at c = 0
            }
        }
        warp_reduce16(red, lane);
        if ((lane & 1) == 0) s_red[warp][lane >> 1] = red[0];
        __syncthreads();
        if (CLUSTER) {
        if (tid < W) {
            double v = 0.0;
#pragma unroll
            for (int w = 0; w < NWARP; ++w) v += s_red[w][tid];
            s_part[slot][tid] = v;
        }
        if (main_step && row0 == c) {                          // row c always lives in CTA 0, r = 0
#pragma unroll
            for (int q = 0; q < W; ++q) s_top[slot][q] = a[0][q];
        }
        qcl_arrive();
        qcl_wait();
        if (warp == 0) {
            if (lane < W) {
                // all C remote loads are issued before the first add (fixed summation order, same in every CTA)
                double pv[16];
#pragma unroll
                for (int q = 0; q < 16; ++q) pv[q] = (q < C) ? qcl_ld(part_base + (slot * W + lane) * 8, (unsigned)q) : 0.0;
                double v = 0.0;
#pragma unroll
                for (int q = 0; q < 16; ++q) v += pv[q];
                s_tot[lane] = v;
                // DLARFG (dlarfg.f:140-186) with xnorm^2 = total[0]; every lane of the half-warp gets total[0]
                const double t0 = __shfl_sync(0x0000ffffu, v, 0);
                const double alpha = main_step ? qcl_ld(top_base + (slot * W + 0) * 8, 0u) : 0.0;
                if (lane == 0 && main_step) {
                    const double xnorm = sqrt(t0);
                    double tau = 0.0, beta = alpha, scale = 0.0;
                    if (xnorm != 0.0) {
                        beta = -copysign(dev_dlapy2(alpha, xnorm), alpha);
                        tau = (beta - alpha) / beta;
                        scale = 1.0 / (alpha - beta);
                    }
                    s_hh[0] = tau; s_hh[1] = beta; s_hh[2] = scale;
                    s_tau[c] = tau;
                    if (g == 0) p.tau[c] = tau;
                }
            } else if (main_step) {
                s_trow[lane - W] = qcl_ld(top_base + (slot * W + (lane - W)) * 8, 0u);
            }
        }
        } else {
        // no cluster (panels taller than 16 CTAs): partial sums and the top row travel through global memory and one
        // grid barrier per column, as in geqr2_leaf_kernel; all CTAs must be co-resident
        if (tid < W) {
            double v = 0.0;
#pragma unroll
            for (int w = 0; w < NWARP; ++w) v += s_red[w][tid];
            p.part[((i64)slot * p.G + g) * W + tid] = v;
        }
        if (main_step && row0 == c) {
            double* dst = p.toprow + slot * W;
#pragma unroll
            for (int q = 0; q < W; ++q) dst[q] = a[0][q];
        }
        qr_grid_barrier(p.bar, p.bar_base + (unsigned)(c + 1) * (unsigned)p.G, p.G);
        if (warp == 0) {
            if (lane < W) {
                double v = 0.0;
                for (int q = 0; q < p.G; ++q) v += __ldcg(p.part + ((i64)slot * p.G + q) * W + lane);
                s_tot[lane] = v;
                const double t0 = __shfl_sync(0x0000ffffu, v, 0);
                const double alpha = main_step ? __ldcg(p.toprow + slot * W) : 0.0;
                if (lane == 0 && main_step) {
                    const double xnorm = sqrt(t0);
                    double tau = 0.0, beta = alpha, scale = 0.0;
                    if (xnorm != 0.0) {
                        beta = -copysign(dev_dlapy2(alpha, xnorm), alpha);
                        tau = (beta - alpha) / beta;
                        scale = 1.0 / (alpha - beta);
                    }
                    s_hh[0] = tau; s_hh[1] = beta; s_hh[2] = scale;
                    s_tau[c] = tau;
                    if (g == 0) p.tau[c] = tau;
                }
            } else if (main_step) {
                s_trow[lane - W] = __ldcg(p.toprow + slot * W + (lane - W));
            }
        }
        }
        __syncthreads();
        if (c >= 1 && tid < pc) s_G[tid][pc] = s_tot[live + tid];      // position live+i holds v_i
        if (main_step) {
            const double tau = s_hh[0], beta = s_hh[1], scale = s_hh[2];
            if (tau != 0.0) {
                // w(q) = C(1,q) + v2^T C2(:,q)  (dlarf1f.f:247-250);  C -= tau*v*w^T (dlarf1f.f:256-258)
                double tw[W];
#pragma unroll
                for (int q = 1; q < W; ++q) tw[q] = -tau * (s_trow[q] + scale * s_tot[q]);
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    const int row = row0 + r * THREADS;
                    if (row < p.m && row > c) {
                        const double v = a[r][0] * scale;            // DSCAL by 1/(alpha-beta), dlarfg.f:178
                        a[r][0] = v;
#pragma unroll
                        for (int q = 1; q < W; ++q)
                            if (q < live) a[r][q] = fma(tw[q], v, a[r][q]);
                    } else if (row == c) {
                        a[r][0] = beta;
#pragma unroll
                        for (int q = 1; q < W; ++q)
                            if (q < live) a[r][q] = a[r][q] + tw[q];
                    }
                }
            }
            // rotate the window: column c goes to the back
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const double t0 = a[r][0];
#pragma unroll
                for (int q = 0; q + 1 < W; ++q) a[r][q] = a[r][q + 1];
                a[r][W - 1] = t0;
            }
        }
    }
    __syncthreads();

    // T for the leaf (dlarft_lvl2.f:199-258): T(0:c,c) = T(0:c,0:c) * (-tau_c * G(0:c,c)), T(c,c) = tau_c
    if (g == 0 && warp == 0) {
        for (int c = 0; c < kmax; ++c) {
            const double tc = s_tau[c];
            double tmp = (lane < c) ? -tc * s_G[lane][c] : 0.0;
            double out = 0.0;
            for (int j = 0; j < c; ++j) {
                double tj = __shfl_sync(0xffffffffu, tmp, j);
                if (lane <= j && lane < c) out += s_T[lane][j] * tj;
            }
            if (lane < c) s_T[lane][c] = (tc == 0.0) ? 0.0 : out;
            if (lane == c) s_T[c][c] = tc;
            __syncwarp();
        }
        for (int idx = lane; idx < kmax * kmax; idx += 32) {
            int i = idx % kmax, j = idx / kmax;
            if (i <= j) p.T[i + (i64)j * p.ldt] = s_T[i][j];
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
                if (col < p.n) {
                    p.A[row + (i64)col * p.lda] = a[r][q];
                    if (p.Vc) {
                        double v = a[r][q];
                        if (col >= kmax || row < col) v = 0.0;
                        else if (row == col) v = 1.0;
                        p.Vc[row + (i64)col * p.ldvc] = v;
                    }
                }
            }
        }
    }
    if (CLUSTER) {
        qcl_arrive();
        qcl_wait();
    }
}

struct QrWs {
    unsigned* bar = nullptr;
    unsigned base = 0;
    double* part = nullptr;
    double* toprow = nullptr;
    int maxG = 0;
};
static QrWs& qr_ws() {
    static QrWs w;
    if (!w.bar) {
        w.maxG = 1024;
        LB_CUDA_CHECK(cudaMalloc(&w.bar, 256));
        LB_CUDA_CHECK(cudaMemset(w.bar, 0, 256));
        LB_CUDA_CHECK(cudaMalloc(&w.part, sizeof(double) * 2 * w.maxG * QW));
        LB_CUDA_CHECK(cudaMalloc(&w.toprow, sizeof(double) * 2 * QW));
    }
    return w;
}

constexpr int QCL_THREADS = 256, QCL_R = 4;       // cluster leaf: 256 threads x 4 rows = 1024 rows per CTA
static_assert(QCL_THREADS * QCL_R == QTHREADS, "both leaf kernels cover 1024 rows per CTA");
static int g_qr_cluster_max = 16;
void geqrf_set_cluster_max(int c) { g_qr_cluster_max = c < 0 ? 0 : (c > 16 ? 16 : c); }
static int qr_cluster_hw_max() {
    static int hw_max = 0;
    if (!hw_max) {
        auto kern = geqr2_leaf_cluster_kernel<QW, QCL_THREADS, QCL_R, true>;
        hw_max = 8;
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess) {
            cudaLaunchConfig_t q = {};
            q.gridDim = dim3(16); q.blockDim = dim3(QCL_THREADS);
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

// leaf: m x n (n <= QW); writes A (R and v), tau, Vc (clean, may be null) and the n x n T block
static void geqr2_leaf(cudaStream_t s, int m, int n, double* A, i64 lda, double* tau, double* Vc, i64 ldvc, double* T,
                       i64 ldt) {
    QrWs& w = qr_ws();
    QrLeafParams p;
    p.m = m; p.n = n; p.A = A; p.lda = lda; p.tau = tau; p.Vc = Vc; p.ldvc = ldvc; p.T = T; p.ldt = ldt;
    p.G = ceil_div(m, QTHREADS);
    p.bar = w.bar; p.bar_base = w.base; p.part = w.part; p.toprow = w.toprow;
    {
        // all CTAs of a leaf must be co-resident (grid barrier); taller panels use 8 or 16 rows per thread (slower:
        // the row window spills to local memory), 256 x 16 x 146 = 598,016 rows is the limit
        const int cap = min(w.maxG, num_sms() - 2);
        if (p.G > cap) {
            int rows_per_cta = 2048;
            if (ceil_div(m, rows_per_cta) > cap) rows_per_cta = 4096;
            p.G = ceil_div(m, rows_per_cta);
            if (p.G > cap) {
                fprintf(stderr, "lapack_b200: QR panel of %d rows exceeds the cooperative leaf capacity (%d rows)\n", m, cap * 4096);
                record_cuda_error(cudaErrorInvalidValue);
                return;
            }
            if (rows_per_cta == 2048) geqr2_leaf_cluster_kernel<QW, QCL_THREADS, 8, false><<<p.G, QCL_THREADS, 0, s>>>(p);
            else geqr2_leaf_cluster_kernel<QW, QCL_THREADS, 16, false><<<p.G, QCL_THREADS, 0, s>>>(p);
            count_launch();
            if (p.G > 1) w.base += (unsigned)(min(m, n) + 1) * (unsigned)p.G;
            LB_CUDA_CHECK(cudaGetLastError());
            return;
        }
    }
    if (p.G <= g_qr_cluster_max && p.G <= qr_cluster_hw_max()) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)p.G);
        cfg.blockDim = dim3(QCL_THREADS);
        cfg.stream = s;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = (unsigned)p.G; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        LB_CUDA_CHECK(cudaLaunchKernelEx(&cfg, geqr2_leaf_cluster_kernel<QW, QCL_THREADS, QCL_R, true>, p));
        count_launch();
        return;
    }
    geqr2_leaf_kernel<QW, QTHREADS><<<p.G, QTHREADS, 0, s>>>(p);
    count_launch();
    if (p.G > 1) w.base += (unsigned)(min(m, n) + 1) * (unsigned)p.G;
    LB_CUDA_CHECK(cudaGetLastError());
}

// Recursive panel: A (m x n) -> R, V (in A), tau, clean Vc (m x n), T (n x n upper, lower part zero on entry).
// tmp: scratch of at least n*n doubles (+ n x ncols for the in-panel block reflector application).
struct PanelWs {
    double* W1;   // (n/2) x (n/2)
    double* W2;
};
static void geqrf_panel(cudaStream_t s, int m, int n, double* A, i64 lda, double* tau, double* Vc, i64 ldvc, double* T,
                        i64 ldt, double* W1, double* W2) {
    if (m <= 0 || n <= 0) return;
    if (n <= QW) { geqr2_leaf(s, m, n, A, lda, tau, Vc, ldvc, T, ldt); return; }
    int n1 = QW;
    while (n1 * 2 < n) n1 *= 2;
    if (n1 > m) n1 = m;          // cannot have more reflectors than rows in the left part
    const int n2 = n - n1;
    // left half
    geqrf_panel(s, m, n1, A, lda, tau, Vc, ldvc, T, ldt, W1, W2);
    const int k1 = min(m, n1);
    // apply H1^T to the right half: C = A(:, n1:n)
    double* C = A + (i64)n1 * lda;
    i64 ldw = (k1 + 1) & ~1;
    gemm(s, 'T', 'N', k1, n2, m, 1.0, Vc, ldvc, C, lda, 0.0, W1, ldw);            // W1 = V1^T C      (dlarfb.f:257-275)
    gemm(s, 'T', 'N', k1, n2, k1, 1.0, T, ldt, W1, ldw, 0.0, W2, ldw);            // W2 = T1^T W1     (dlarfb.f:277)
    gemm(s, 'N', 'N', m, n2, k1, -1.0, Vc, ldvc, W2, ldw, 1.0, C, lda);           // C -= V1 W2       (dlarfb.f:287-304)
    if (m > n1) {
        // right half on the trailing rows
        geqrf_panel(s, m - n1, n2, A + n1 + (i64)n1 * lda, lda, tau + n1, Vc + n1 + (i64)n1 * ldvc, ldvc,
                    T + n1 + (i64)n1 * ldt, ldt, W1, W2);
        // rows 0..n1-1 of the right half of Vc are zero
        laset(s, 'A', n1, n2, 0.0, 0.0, Vc + (i64)n1 * ldvc, ldvc);
        const int k2 = min(m - n1, n2);
        // T12 = -T1 (V1^T V2) T2      (dlarft.f:308-349)
        gemm(s, 'T', 'N', k1, k2, m - n1, 1.0, Vc + n1, ldvc, Vc + n1 + (i64)n1 * ldvc, ldvc, 0.0, W1, ldw);
        gemm(s, 'N', 'N', k1, k2, k1, -1.0, T, ldt, W1, ldw, 0.0, W2, ldw);
        gemm(s, 'N', 'N', k1, k2, k2, 1.0, W2, ldw, T + n1 + (i64)n1 * ldt, ldt, 0.0, T + (i64)n1 * ldt, ldt);
    } else {
        laset(s, 'A', m, n2, 0.0, 0.0, Vc + (i64)n1 * ldvc, ldvc);
    }
}

// max |A| of a device matrix (finite entries only), as the bit pattern of a non-negative double
__global__ void amax_kernel(int m, int n, const double* __restrict__ A, i64 lda, unsigned long long* amax_bits) {
    double v = 0.0;
    for (i64 j = blockIdx.x; j < n; j += gridDim.x)
        for (int i = threadIdx.x; i < m; i += blockDim.x) {
            double x = fabs(A[i + j * lda]);
            if (x == x && x <= DBL_MAX) v = fmax(v, x);   // ignore NaN/Inf: they propagate on their own
        }
    for (int off = 16; off > 0; off >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, off));
    if ((threadIdx.x & 31) == 0) atomicMax(amax_bits, (unsigned long long)__double_as_longlong(v));
}

// Exact power-of-two scaling PER COLUMN against over/underflow in the fused sum-of-squares reduction of the leaves (the reference
// gets the same protection from DNRM2's scaled accumulation and DLARFG's rescaling loop, dlarfg.f:159-176).  QR commutes with
// positive column scalings: A D = Q (R D) with the same reflectors and the same tau, so a column whose largest entry lies outside
// 2^+-300 is multiplied by 2^(1-e) before the factorization and its part of R (rows <= column) by the inverse afterwards.  Powers of
// two: every other column's result is bit-identical to the unscaled computation.
__global__ void col_scale_decide_kernel(int m, int n, const double* __restrict__ A, i64 lda, double* __restrict__ dcol, int* __restrict__ any) {
    const int j = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (j >= n) return;
    const double* col = A + (i64)j * lda;
    double v = 0.0;
    for (int i = lane; i < m; i += 32) {
        const double x = fabs(col[i]);
        if (x == x && x <= DBL_MAX) v = fmax(v, x);       // ignore NaN/Inf: they propagate on their own
    }
    for (int off = 16; off > 0; off >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, off));
    if (lane == 0) {
        double d = 1.0;
        if (v > 0.0) {
            int e;
            frexp(v, &e);
            if (e > 300 || e < -300) { d = ldexp(1.0, 1 - e); atomicOr(any, 1); }
        }
        dcol[j] = d;
    }
}
// A(:,j) *= d(j) (inverse = 0) or A(0:j, j) /= d(j) (inverse = 1: un-scale R), flagged columns only
__global__ void col_scale_apply_kernel(int m, int n, double* __restrict__ A, i64 lda, const double* __restrict__ dcol, const int* __restrict__ any,
                                       int inverse) {
    if (*any == 0) return;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    for (int j = blockIdx.y; j < n; j += gridDim.y) {
        const double d = dcol[j];
        if (d == 1.0) continue;
        if (inverse) { if (i <= j) A[i + (i64)j * lda] *= (1.0 / d); }
        else A[i + (i64)j * lda] *= d;
    }
}

// one mutex for every driver that uses the look-ahead streams / events of lb::aux() (runtime.cu): concurrent host threads calling
// different factorizations through the device API must not interleave their event joins
static std::recursive_mutex& g_qr_mutex = driver_mutex();

// C := (I - V T^T V^T) C for the m x nc block C, V/T given as clean copies  (dlarfb.f:248-304 as three GEMMs)
static void apply_block_reflector(cudaStream_t s, int m, int nc, int k, const double* Vc, i64 ldvc, const double* T, i64 ldt,
                                  double* C, i64 ldc, double* W1, double* W2, i64 ldw) {
    if (nc <= 0) return;
    gemm(s, 'T', 'N', k, nc, m, 1.0, Vc, ldvc, C, ldc, 0.0, W1, ldw);       // W1 = V^T C
    gemm(s, 'T', 'N', k, nc, k, 1.0, T, ldt, W1, ldw, 0.0, W2, ldw);        // W2 = T^T W1
    gemm(s, 'N', 'N', m, nc, k, -1.0, Vc, ldvc, W2, ldw, 1.0, C, ldc);      // C -= V W2
}

// Tout != nullptr (DGEQRT): the block size is the caller's and the upper triangle of every panel's T factor is copied to
// Tout(1:jb, j:j+jb) -- the layout of SRC/dgeqrt.f:196-198.
static void geqrf_impl(cudaStream_t s, int m, int n, double* A, i64 lda, double* tau, int nb, bool lookahead,
                       double* Tout = nullptr, i64 ldtout = 0) {
    if (m <= 0 || n <= 0) return;
    const int k = min(m, n);
    nb = min(nb, k);
    if (!Tout && nb < QW) nb = min(QW, k);
    const bool la = lookahead && k > nb;
    // scratch: two (Vc, T) sets so that panel j+1 can be built while update j still reads set j
    const i64 ldvc = (m + 1) & ~1;
    const i64 ldt = (nb + 1) & ~1;
    const i64 ldw = (nb + 1) & ~1;
    double* Vc[2];
    double* T[2];
    Vc[0] = (double*)ws_alloc(s, sizeof(double) * ldvc * nb);
    T[0] = (double*)ws_alloc(s, sizeof(double) * ldt * nb);
    Vc[1] = la ? (double*)ws_alloc(s, sizeof(double) * ldvc * nb) : Vc[0];
    T[1] = la ? (double*)ws_alloc(s, sizeof(double) * ldt * nb) : T[0];
    double* W1p = (double*)ws_alloc(s, sizeof(double) * ldw * nb);            // in-panel scratch (panel stream)
    double* W2p = (double*)ws_alloc(s, sizeof(double) * ldw * nb);
    double* W1u = (double*)ws_alloc(s, sizeof(double) * ldw * max(n, nb));    // trailing-update scratch
    double* W2u = (double*)ws_alloc(s, sizeof(double) * ldw * max(n, nb));
    double* dcol = (double*)ws_alloc(s, sizeof(double) * (size_t)n + 64);
    int* any_scaled = (int*)(dcol + n);

    // exact power-of-two pre-scaling of out-of-range columns (see col_scale_decide_kernel)
    LB_CUDA_CHECK(cudaMemsetAsync(any_scaled, 0, sizeof(int), s));
    col_scale_decide_kernel<<<ceil_div(n, 8), 256, 0, s>>>(m, n, A, lda, dcol, any_scaled);
    dim3 sgrid(ceil_div(m, 256), (unsigned)min(n, 4096));
    col_scale_apply_kernel<<<sgrid, 256, 0, s>>>(m, n, A, lda, dcol, any_scaled, 0);
    count_launch(2);

    Aux& ax = aux();
    cudaStream_t sp = la ? ax.panel_stream : s;
    cudaStream_t su = la ? ax.update_stream : s;
    cudaEvent_t ev_panel = ax.ev[6], ev_next = ax.ev[7], ev_join = ax.ev[2];
    if (la) {
        LB_CUDA_CHECK(cudaEventRecord(ev_join, s));
        LB_CUDA_CHECK(cudaStreamWaitEvent(sp, ev_join, 0));
        LB_CUDA_CHECK(cudaStreamWaitEvent(su, ev_join, 0));
    }
    auto panel = [&](int j, int jb, int set) {
        LB_CUDA_CHECK(cudaMemsetAsync(T[set], 0, sizeof(double) * ldt * nb, sp));
        geqrf_panel(sp, m - j, jb, A + j + (i64)j * lda, lda, tau + j, Vc[set], ldvc, T[set], ldt, W1p, W2p);
        if (Tout) lacpy(sp, 'U', jb, jb, T[set], ldt, Tout + (i64)j * ldtout, ldtout);
    };
    panel(0, min(nb, k), 0);
    if (la) LB_CUDA_CHECK(cudaEventRecord(ev_panel, sp));
    int set = 0;
    for (int j = 0; j < k; j += nb, set ^= (la ? 1 : 0)) {
        const int jb = min(nb, k - j);
        const int jn = j + jb;
        const int mj = m - j;
        if (jn >= n) break;
        if (la) LB_CUDA_CHECK(cudaStreamWaitEvent(su, ev_panel, 0));
        const int jb2 = (jn < k) ? min(nb, k - jn) : 0;
        if (jb2 > 0) {
            // look-ahead: next panel's columns first, then factor it while the rest is updated
            apply_block_reflector(su, mj, jb2, jb, Vc[set], ldvc, T[set], ldt, A + j + (i64)jn * lda, lda, W1u, W2u, ldw);
            if (la) { LB_CUDA_CHECK(cudaEventRecord(ev_next, su)); LB_CUDA_CHECK(cudaStreamWaitEvent(sp, ev_next, 0)); }
            if (la) {
                panel(jn, jb2, set ^ 1);
                LB_CUDA_CHECK(cudaEventRecord(ev_panel, sp));
            }
            apply_block_reflector(su, mj, n - jn - jb2, jb, Vc[set], ldvc, T[set], ldt, A + j + (i64)(jn + jb2) * lda, lda, W1u,
                                  W2u, ldw);
            if (!la) panel(jn, jb2, set);     // sequential mode: same scratch set, after the whole update
        } else {
            apply_block_reflector(su, mj, n - jn, jb, Vc[set], ldvc, T[set], ldt, A + j + (i64)jn * lda, lda, W1u, W2u, ldw);
        }
    }
    if (la) {
        LB_CUDA_CHECK(cudaEventRecord(ev_join, su));
        LB_CUDA_CHECK(cudaStreamWaitEvent(s, ev_join, 0));
        LB_CUDA_CHECK(cudaEventRecord(ev_next, sp));
        LB_CUDA_CHECK(cudaStreamWaitEvent(s, ev_next, 0));
    }
    col_scale_apply_kernel<<<sgrid, 256, 0, s>>>(m, n, A, lda, dcol, any_scaled, 1);   // un-scale R only
    count_launch();
    ws_free(s, Vc[0]); ws_free(s, T[0]);
    if (la) { ws_free(s, Vc[1]); ws_free(s, T[1]); }
    ws_free(s, W1p); ws_free(s, W2p); ws_free(s, W1u); ws_free(s, W2u); ws_free(s, dcol);
    LB_CUDA_CHECK(cudaGetLastError());
}

void geqrf(cudaStream_t s, int m, int n, double* A, i64 lda, double* tau) {
    std::lock_guard<std::recursive_mutex> lock(g_qr_mutex);
    geqrf_impl(s, m, n, A, lda, tau, g_qr_nb, g_qr_lookahead != 0);
}
// DGEQR2: same factorization, panel-only code path (one recursive panel per 64 columns)
void geqr2(cudaStream_t s, int m, int n, double* A, i64 lda, double* tau) {
    std::lock_guard<std::recursive_mutex> lock(g_qr_mutex);
    geqrf_impl(s, m, n, A, lda, tau, 64, false);
}

// ------------------------------------------------------------------------------------------------
// clean dense copy Vc (n x k, one reflector per column, explicit unit entry and explicit zeros) of the reflectors stored in V
// for every DIRECT / STOREV combination of dlarft.f:100-150 / dlarfb.f:150-190:
//   forward : v_j has its unit entry in row j and zeros above;      backward: unit entry in row n-k+j, zeros below
//   columnwise: v_j = V(:, j) (V is n x k);                          rowwise : v_j = V(j, :) (V is k x n)
__global__ void clean_v_kernel(int n, int k, const double* __restrict__ V, i64 ldv, double* __restrict__ Vc, i64 ldvc, bool backward,
                               bool rowwise) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    for (int j = blockIdx.y; j < k; j += gridDim.y) {
        const int u = backward ? n - k + j : j;                 // row of the unit entry
        double v;
        if (i == u) v = 1.0;
        else if (backward ? i > u : i < u) v = 0.0;
        else v = rowwise ? V[j + (i64)i * ldv] : V[i + (i64)j * ldv];
        Vc[i + (i64)j * ldvc] = v;
    }
}
static void clean_v(cudaStream_t s, int n, int k, const double* V, i64 ldv, double* Vc, i64 ldvc, bool backward = false,
                    bool rowwise = false) {
    dim3 grid(ceil_div(n, 256), (unsigned)min(k, 4096));
    clean_v_kernel<<<grid, 256, 0, s>>>(n, k, V, ldv, Vc, ldvc, backward, rowwise);
    count_launch();
}
// index reversal of a k x k matrix / a k-vector (backward DLARFT is forward DLARFT on the reflectors in reverse order)
__global__ void flip_square_kernel(int k, const double* __restrict__ A, i64 lda, double* __restrict__ B, i64 ldb, int tri) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y;
    if (i >= k) return;
    if (tri == 2 && i < j) return;                              // write the lower triangle only
    B[i + (i64)j * ldb] = A[(k - 1 - i) + (i64)(k - 1 - j) * lda];
}
__global__ void flip_vec_kernel(int k, const double* __restrict__ x, double* __restrict__ y) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < k) y[i] = x[k - 1 - i];
}

// T block (len <= 32) from G = V^T V and tau: dlarft_lvl2.f:199-258
__global__ void larft_leaf_kernel(int len, const double* __restrict__ G, i64 ldg, const double* __restrict__ tau,
                                  double* __restrict__ T, i64 ldt) {
    __shared__ double sT[32][33];
    const int lane = threadIdx.x;
    for (int c = 0; c < len; ++c) {
        const double tc = tau[c];
        double tmp = (lane < c) ? -tc * G[lane + (i64)c * ldg] : 0.0;
        double out = 0.0;
        for (int j = 0; j < c; ++j) {
            double tj = __shfl_sync(0xffffffffu, tmp, j);
            if (lane <= j && lane < c) out += sT[lane][j] * tj;
        }
        if (lane < c) sT[lane][c] = (tc == 0.0) ? 0.0 : out;
        if (lane == c) sT[c][c] = tc;
        __syncwarp();
    }
    for (int idx = lane; idx < len * len; idx += 32) {
        int i = idx % len, j = idx / len;
        if (i <= j) T[i + (i64)j * ldt] = sT[i][j];
    }
}
static void larft_rec(cudaStream_t s, int k, const double* G, i64 ldg, const double* tau, double* T, i64 ldt, double* W1,
                      double* W2, i64 ldw) {
    if (k <= 32) {
        larft_leaf_kernel<<<1, 32, 0, s>>>(k, G, ldg, tau, T, ldt);
        count_launch();
        return;
    }
    int k1 = 32;
    while (k1 * 2 < k) k1 *= 2;
    const int k2 = k - k1;
    larft_rec(s, k1, G, ldg, tau, T, ldt, W1, W2, ldw);
    larft_rec(s, k2, G + k1 + (i64)k1 * ldg, ldg, tau + k1, T + k1 + (i64)k1 * ldt, ldt, W1, W2, ldw);
    // T12 = -T1 * G12 * T2; T1/T2 are read through expanded (zero-filled) copies
    double* E1 = W1;                       // k1 x k1
    double* E2 = W1 + ldw * k1;            // k2 x k2
    double* Tm = W2;                       // k1 x k2
    expand_tri(s, k1, T, ldt, true, false, E1, ldw);
    expand_tri(s, k2, T + k1 + (i64)k1 * ldt, ldt, true, false, E2, ldw);
    gemm(s, 'N', 'N', k1, k2, k1, -1.0, E1, ldw, G + (i64)k1 * ldg, ldg, 0.0, Tm, ldw);
    gemm(s, 'N', 'N', k1, k2, k2, 1.0, Tm, ldw, E2, ldw, 0.0, T + (i64)k1 * ldt, ldt);
}

// DLARFT(DIRECT, STOREV, n, k, V, ldv, tau, T, ldt) (SRC/dlarft.f:160; H = H(1)...H(k) forward, H(k)...H(1) backward; H = I - V T V^T
// columnwise, I - V^T T V rowwise): only the upper (forward) / lower (backward) triangle of T is written.
void larft_general(cudaStream_t s, bool backward, bool rowwise, int n, int k, const double* V, i64 ldv, const double* tau, double* T,
                   i64 ldt) {
    if (n <= 0 || k <= 0) return;
    const i64 ldvc = (n + 1) & ~1, ldg = (k + 1) & ~1, ldw = (k + 1) & ~1;
    double* Vc = (double*)ws_alloc(s, sizeof(double) * ldvc * k);
    double* G = (double*)ws_alloc(s, sizeof(double) * ldg * k);
    double* W1 = (double*)ws_alloc(s, sizeof(double) * ldw * (k + 2));
    double* W2 = (double*)ws_alloc(s, sizeof(double) * ldw * (k + 2));
    clean_v(s, n, k, V, ldv, Vc, ldvc, backward, rowwise);
    gemm(s, 'T', 'N', k, k, n, 1.0, Vc, ldvc, Vc, ldvc, 0.0, G, ldg);
    if (!backward) larft_rec(s, k, G, ldg, tau, T, ldt, W1, W2, ldw);
    else {
        // reflectors in reverse order: G' = J G J, tau' = J tau, T' upper from the forward recurrence, T = J T' J (lower)
        double* Gf = (double*)ws_alloc(s, sizeof(double) * ldg * k);
        double* Tf = (double*)ws_alloc(s, sizeof(double) * ldg * k);
        double* tf = (double*)ws_alloc(s, sizeof(double) * (size_t)k);
        dim3 grid(ceil_div(k, 128), (unsigned)k);
        flip_square_kernel<<<grid, 128, 0, s>>>(k, G, ldg, Gf, ldg, 0);
        flip_vec_kernel<<<ceil_div(k, 128), 128, 0, s>>>(k, tau, tf);
        LB_CUDA_CHECK(cudaMemsetAsync(Tf, 0, sizeof(double) * ldg * k, s));
        larft_rec(s, k, Gf, ldg, tf, Tf, ldg, W1, W2, ldw);
        flip_square_kernel<<<grid, 128, 0, s>>>(k, Tf, ldg, T, ldt, 2);
        count_launch(3);
        ws_free(s, Gf); ws_free(s, Tf); ws_free(s, tf);
    }
    ws_free(s, Vc); ws_free(s, G); ws_free(s, W1); ws_free(s, W2);
}
void larft(cudaStream_t s, int n, int k, const double* V, i64 ldv, const double* tau, double* T, i64 ldt) {
    larft_general(s, false, false, n, k, V, ldv, tau, T, ldt);
}

// DLARFB(SIDE, TRANS, DIRECT, STOREV, m, n, k, V, T, C) (SRC/dlarfb.f:192): C := H C, H^T C, C H or C H^T with the block
// reflector H = I - Vc T Vc^T, Vc the dense n x k reflector matrix of any storage scheme, T upper (forward) or lower
// (backward) triangular.  Two GEMMs with the long dimension and a k x k one in between (dlarfb.f:248-304 and the other 7 cases).
void larfb_general(cudaStream_t s, char side, char trans, bool backward, bool rowwise, int m, int n, int k, const double* V, i64 ldv,
                   const double* T, i64 ldt, double* C, i64 ldc) {
    if (m <= 0 || n <= 0 || k <= 0) return;
    const bool left = (side == 'L' || side == 'l');
    const bool tr = !(trans == 'N' || trans == 'n');
    const int nv = left ? m : n;
    const i64 ldvc = (nv + 1) & ~1, lde = (k + 1) & ~1;
    double* Vc = (double*)ws_alloc(s, sizeof(double) * ldvc * k);
    double* E = (double*)ws_alloc(s, sizeof(double) * lde * k);
    clean_v(s, nv, k, V, ldv, Vc, ldvc, backward, rowwise);
    expand_tri(s, k, T, ldt, !backward, false, E, lde);
    if (left) {
        const i64 ldw = (k + 1) & ~1;
        double* W1 = (double*)ws_alloc(s, sizeof(double) * ldw * n);
        double* W2 = (double*)ws_alloc(s, sizeof(double) * ldw * n);
        gemm(s, 'T', 'N', k, n, m, 1.0, Vc, ldvc, C, ldc, 0.0, W1, ldw);                 // W1 = V^T C
        gemm(s, tr ? 'T' : 'N', 'N', k, n, k, 1.0, E, lde, W1, ldw, 0.0, W2, ldw);       // W2 = op(T) W1
        gemm(s, 'N', 'N', m, n, k, -1.0, Vc, ldvc, W2, ldw, 1.0, C, ldc);                // C -= V W2
        ws_free(s, W1); ws_free(s, W2);
    } else {
        const i64 ldw = (m + 1) & ~1;
        double* W1 = (double*)ws_alloc(s, sizeof(double) * ldw * k);
        double* W2 = (double*)ws_alloc(s, sizeof(double) * ldw * k);
        gemm(s, 'N', 'N', m, k, n, 1.0, C, ldc, Vc, ldvc, 0.0, W1, ldw);                 // W1 = C V
        gemm(s, 'N', tr ? 'T' : 'N', m, k, k, 1.0, W1, ldw, E, lde, 0.0, W2, ldw);       // W2 = W1 op(T)
        gemm(s, 'N', 'T', m, n, k, -1.0, W2, ldw, Vc, ldvc, 1.0, C, ldc);                // C -= W2 V^T
        ws_free(s, W1); ws_free(s, W2);
    }
    ws_free(s, Vc); ws_free(s, E);
}
void larfb(cudaStream_t s, char side, char trans, int m, int n, int k, const double* V, i64 ldv, const double* T, i64 ldt,
           double* C, i64 ldc) {
    larfb_general(s, side, trans, false, false, m, n, k, V, ldv, T, ldt, C, ldc);
}

// ------------------------------------------------------------------------------------------------
// DORMQR (SRC/dormqr.f:283-333): C := Q C, Q^T C, C Q or C Q^T with Q = H(1)...H(k) from DGEQRF.  Same block loop as
// the reference (forward for (L,T) and (R,N), backward otherwise) with NB = 256 instead of 32: DLARFT + DLARFB per block.
void ormqr(cudaStream_t s, char side, char trans, int m, int n, int k, const double* A, i64 lda, const double* tau,
           double* C, i64 ldc) {
    if (m <= 0 || n <= 0 || k <= 0) return;
    const bool left = (side == 'L' || side == 'l');
    const bool notran = (trans == 'N' || trans == 'n');
    const int nq = left ? m : n;
    const int nb = 256;
    const bool forward = (left && !notran) || (!left && notran);
    const i64 ldt = nb;
    double* T = (double*)ws_alloc(s, sizeof(double) * ldt * nb);
    const int nblk = ceil_div(k, nb);
    for (int b = 0; b < nblk; ++b) {
        const int i = forward ? b * nb : (nblk - 1 - b) * nb;
        const int ib = min(nb, k - i);
        const double* Aii = A + i + (i64)i * lda;
        larft(s, nq - i, ib, Aii, lda, tau + i, T, ldt);
        if (left) larfb(s, 'L', trans, m - i, n, ib, Aii, lda, T, ldt, C + i, ldc);
        else larfb(s, 'R', trans, m, n - i, ib, Aii, lda, T, ldt, C + (i64)i * ldc, ldc);
    }
    ws_free(s, T);
}

__global__ void add_diag_kernel(int n, double* A, i64 lda, double v) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) A[i + (i64)i * lda] += v;
}

// DORGQR (SRC/dorgqr.f:220-281): the first n columns of Q = H(1)...H(k), in place over the reflectors.  Blocks are
// processed backwards like the reference; for each block H is applied to the columns on its right (DLARFB 'L','N'),
// then the block's own columns become  H [I; 0] = [I; 0] - V (T V1^T)  (what DORG2R computes column by column).
void orgqr(cudaStream_t s, int m, int n, int k, double* A, i64 lda, const double* tau) {
    if (n <= 0) return;
    if (n > k) {
        // columns k..n-1 start as columns of the unit matrix (dorg2r.f:151-158)
        laset(s, 'A', k, n - k, 0.0, 0.0, A + (i64)k * lda, lda);
        laset(s, 'A', m - k, n - k, 0.0, 1.0, A + k + (i64)k * lda, lda);
    }
    if (k <= 0) return;
    const int nb = 256;
    const i64 ldt = nb, ldvc = ((i64)m + 1) & ~1LL;
    double* T = (double*)ws_alloc(s, sizeof(double) * ldt * nb);
    double* E = (double*)ws_alloc(s, sizeof(double) * ldt * nb);
    double* W = (double*)ws_alloc(s, sizeof(double) * ldt * nb);
    double* Vc = (double*)ws_alloc(s, sizeof(double) * ldvc * nb);
    const int nblk = ceil_div(k, nb);
    for (int b = nblk - 1; b >= 0; --b) {
        const int i = b * nb;
        const int ib = min(nb, k - i);
        const int mi = m - i;
        double* Aii = A + i + (i64)i * lda;
        larft(s, mi, ib, Aii, lda, tau + i, T, ldt);
        if (i + ib < n) larfb(s, 'L', 'N', mi, n - i - ib, ib, Aii, lda, T, ldt, A + i + (i64)(i + ib) * lda, lda);
        clean_v(s, mi, ib, Aii, lda, Vc, ldvc);
        expand_tri(s, ib, T, ldt, true, false, E, ldt);
        gemm(s, 'N', 'T', ib, ib, ib, 1.0, E, ldt, Vc, ldvc, 0.0, W, ldt);          // W = T V1^T
        gemm(s, 'N', 'N', mi, ib, ib, -1.0, Vc, ldvc, W, ldt, 0.0, Aii, lda);       // block := -V W
        add_diag_kernel<<<ceil_div(ib, 128), 128, 0, s>>>(ib, Aii, lda, 1.0);       //          + [I; 0]
        count_launch();
        if (i > 0) laset(s, 'A', i, ib, 0.0, 0.0, A + (i64)i * lda, lda);           // rows above the block (dorgqr.f:270-274)
    }
    ws_free(s, T); ws_free(s, E); ws_free(s, W); ws_free(s, Vc);
}

// ------------------------------------------------------------------------------------------------
// DGEQRT / DGEMQRT (SRC/dgeqrt.f:166-211, SRC/dgemqrt.f:199-287): the blocked QR above with the
// caller's block size, keeping the T factors; and the application of Q / Q^T from those stored factors.
void geqrt(cudaStream_t s, int m, int n, int nb, double* A, i64 lda, double* T, i64 ldt) {
    const int k = min(m, n);
    if (k <= 0) return;
    std::lock_guard<std::recursive_mutex> lock(g_qr_mutex);
    double* tau = (double*)ws_alloc(s, sizeof(double) * (size_t)k);
    geqrf_impl(s, m, n, A, lda, tau, nb, g_qr_lookahead != 0, T, ldt);
    ws_free(s, tau);
}

// DLATSQR (SRC/dlatsqr.f:185-290): tall-skinny QR by row blocks of MB rows.  After DGEQRT of the first block
// every further block B (MB-N rows) is combined with the current R by DTPQRT(L = 0) -- the QR factorization of the stacked matrix
// [R; B].  Because the rows of R below the diagonal are exactly zero, Householder QR of the stacked (N + rows) x N matrix produces
// precisely DTPQRT's reflectors (v = [e_i; b_i], dtpqrt2.f:214-232), the same T blocks and the same updated R, so each step is
// the blocked DGEQRT path of this file on a stacked scratch copy; R goes back to A(1:N,1:N), the reflector block to the rows of A,
// the T blocks to T(1, CTR*N+1).
void latsqr(cudaStream_t s, int m, int n, int mb, int nb, double* A, i64 lda, double* T, i64 ldt) {
    if (min(m, n) <= 0) return;
    if (mb <= n || mb >= m) { geqrt(s, m, n, nb, A, lda, T, ldt); return; }          // dlatsqr.f:252-255
    const int kk = (m - n) % (mb - n), ii = m - kk;
    geqrt(s, mb, n, nb, A, lda, T, ldt);                                             // dlatsqr.f:261
    const i64 lds = ((i64)mb + 1) & ~1LL;
    double* S = (double*)ws_alloc(s, sizeof(double) * (size_t)lds * n);
    auto tp = [&](int i, int rows, int ctr) {
        laset(s, 'A', n, n, 0.0, 0.0, S, lds);
        lacpy(s, 'U', n, n, A, lda, S, lds);
        lacpy(s, 'A', rows, n, A + i, lda, S + n, lds);
        geqrt(s, n + rows, n, nb, S, lds, T + (i64)ctr * n * ldt, ldt);              // == DTPQRT(rows, N, 0, NB, ...)
        lacpy(s, 'U', n, n, S, lds, A, lda);
        lacpy(s, 'A', rows, n, S + n, lds, A + i, lda);
    };
    int ctr = 1;
    for (int i = mb; i <= ii - mb + n; i += mb - n) tp(i, mb - n, ctr++);            // dlatsqr.f:264-269
    if (ii < m) tp(ii, kk, ctr);                                                     // dlatsqr.f:273-277
    ws_free(s, S);
}

void gemqrt(cudaStream_t s, char side, char trans, int m, int n, int k, int nb, const double* V, i64 ldv, const double* T,
            i64 ldt, double* C, i64 ldc) {
    if (m <= 0 || n <= 0 || k <= 0) return;
    const bool left = (side == 'L' || side == 'l');
    const bool notran = (trans == 'N' || trans == 'n');
    const bool forward = (left && !notran) || (!left && notran);               // dgemqrt.f:242-285
    const int nblk = ceil_div(k, nb);
    for (int b = 0; b < nblk; ++b) {
        const int i = forward ? b * nb : (nblk - 1 - b) * nb;
        const int ib = min(nb, k - i);
        const double* Vi = V + i + (i64)i * ldv;
        const double* Ti = T + (i64)i * ldt;
        if (left) larfb(s, 'L', trans, m - i, n, ib, Vi, ldv, Ti, ldt, C + i, ldc);
        else larfb(s, 'R', trans, m, n - i, ib, Vi, ldv, Ti, ldt, C + (i64)i * ldc, ldc);
    }
}

// ------------------------------------------------------------------------------------------------
// LQ through QR of the transpose (for DGELS): A = L Q  <=>  A^T = Q^T L^T, so DGEQRF of A^T gives
// R = L^T and the very reflectors DGELQF produces (SRC/dgelq2.f:165-183 runs DLARFG on the rows of A exactly as
// DGEQR2 does on the columns of A^T); transposing the factored array back yields DGELQF's storage: L on and below the
// diagonal, reflector i in row i to the right of the diagonal, same TAU.  The two transposes are HBM-bound passes.
void gelqf(cudaStream_t s, int m, int n, double* A, i64 lda, double* tau) {
    if (m <= 0 || n <= 0) return;
    const i64 ldt = ((i64)n + 1) & ~1LL;
    double* At = (double*)ws_alloc(s, sizeof(double) * (size_t)ldt * m);
    transpose(s, m, n, A, lda, At, ldt);
    geqrf(s, n, m, At, ldt, tau);
    transpose(s, n, m, At, ldt, A, lda);
    ws_free(s, At);
}

// DORMLQ (SRC/dormlq.f): Q = H(k) ... H(1) with the reflectors in the rows of A (k x nq).  With V = A(1:k,1:nq)^T the
// same reflectors are the columns of V and Q = (H(1) ... H(k))^T, so Q C = DORMQR('T') and Q^T C = DORMQR('N') on V.
void ormlq(cudaStream_t s, char side, char trans, int m, int n, int k, const double* A, i64 lda, const double* tau, double* C,
           i64 ldc) {
    if (m <= 0 || n <= 0 || k <= 0) return;
    const bool left = (side == 'L' || side == 'l');
    const bool notran = (trans == 'N' || trans == 'n');
    const int nq = left ? m : n;
    const i64 ldv = ((i64)nq + 1) & ~1LL;
    double* V = (double*)ws_alloc(s, sizeof(double) * (size_t)ldv * k);
    transpose(s, k, nq, A, lda, V, ldv);
    ormqr(s, side, notran ? 'T' : 'N', m, n, k, V, ldv, tau, C, ldc);
    ws_free(s, V);
}

// max |a_ij| (DLANGE 'M' for finite data), returned on the host; synchronises the stream
double amax_abs(cudaStream_t s, int m, int n, const double* A, i64 lda) {
    if (m <= 0 || n <= 0) return 0.0;
    unsigned long long* d = (unsigned long long*)ws_alloc(s, 64);
    LB_CUDA_CHECK(cudaMemsetAsync(d, 0, 8, s));
    amax_kernel<<<min(n, 4 * num_sms()), 256, 0, s>>>(m, n, A, lda, d);
    count_launch();
    unsigned long long h = 0;
    LB_CUDA_CHECK(cudaMemcpyAsync(&h, d, 8, cudaMemcpyDeviceToHost, s));
    LB_CUDA_CHECK(cudaStreamSynchronize(s));
    ws_free(s, d);
    return __longlong_as_double_host(h);
}

}  // namespace lb
