// batched.cu -- batched factorization of many 32 x 32 matrices (config C5b): one matrix per warp (DPOTRF, and the DGETRF
// kernel kept for comparison) or per half-warp (default DGETRF kernel, further down).
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

// One matrix per warp, lane = row, the row's 32 entries in registers (column loop fully unrolled so every index
// is static).  Pivot search: |a(:,c)| is non-negative, so its bit pattern orders like an unsigned integer and the
// warp maximum is two 32-bit redux.sync (high word, then low word among the leaders) plus a ballot; ties go to
// the smallest row position like IDAMAX (idamax.f:95-106).  The pivot row reaches the other lanes through a
// 256-byte shared-memory line per warp (one 16-byte store per column pair by the owner, broadcast loads by
// everybody) instead of two shuffles per entry; the line is double buffered on the step parity so one
// __syncwarp per step orders everything.  Interchanges are implicit: a row never moves, only its position does.
constexpr int BWARPS = 4;
__global__ void __launch_bounds__(BWARPS * 32, 5) getrf_batched32_kernel(i64 batch, double* __restrict__ A, int* __restrict__ ipiv,
                                                                       int* __restrict__ info) {
    __shared__ __align__(16) double rowbuf[BWARPS][2][BW];
    __shared__ int posbuf[BWARPS][2];
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const i64 id = (i64)blockIdx.x * BWARPS + w;
    if (id >= batch) return;
    double* M = A + id * (BW * BW);
    double a[BW];
#pragma unroll
    for (int q = 0; q < BW; ++q) a[q] = M[lane + BW * q];
    int mypos = lane;        // current position of my row under LAPACK's explicit interchanges
    bool done = false;       // my row has already been used as a pivot row
    int myipiv = 0, minfo = 0;
#pragma unroll
    for (int c = 0; c < BW; ++c) {
        // ---- pivot search
        unsigned long long kb = (unsigned long long)__double_as_longlong(fabs(a[c]));
        bool cand = !done;
        if (a[c] != a[c]) {                                   // NaN only wins from the first place (idamax.f:103)
            if (mypos == c) kb = 0x7ff0000000000000ULL; else cand = false;
        }
        const unsigned hi = (unsigned)(kb >> 32), lo = (unsigned)kb;
        const unsigned mh = __reduce_max_sync(full, cand ? hi : 0u);
        const bool c1 = cand && hi == mh;
        unsigned tie = __ballot_sync(full, c1);
        if (__popc(tie) != 1) {                                // leaders agree in the high word: compare the low words
            const unsigned ml = __reduce_max_sync(full, c1 ? lo : 0u);
            const bool c2 = c1 && lo == ml;
            tie = __ballot_sync(full, c2);
            if (__popc(tie) != 1) {                            // exact tie: smallest position wins
                const unsigned mp = __reduce_min_sync(full, c2 ? (unsigned)mypos : 0xffffffffu);
                tie = __ballot_sync(full, c2 && (unsigned)mypos == mp);
            }
        }
        const int who = __ffs(tie) - 1;                        // lane holding the pivot row
        // ---- publish the pivot row (columns >= c, in aligned pairs) and its current position
        double* rb = rowbuf[w][c & 1];
        if (lane == who) {
            posbuf[w][c & 1] = mypos;
#pragma unroll
            for (int q = c & ~1; q < BW; q += 2) *reinterpret_cast<double2*>(rb + q) = make_double2(a[q], a[q + 1]);
        }
        __syncwarp();
        const int pos = posbuf[w][c & 1];
        if (lane == c) myipiv = pos + 1;
        // interchange bookkeeping: the row sitting at position c moves to `pos`
        if (!done && mypos == c && lane != who) mypos = pos;
        if (lane == who) { mypos = c; done = true; }
        const double pivot = rb[c];
        if (pivot == 0.0) {
            if (minfo == 0) minfo = c + 1;                              // dgetrf2.f:212-214
        } else if (!done) {
            double l;
            if (fabs(pivot) >= DBL_MIN) l = a[c] * (1.0 / pivot);       // dgetrf2.f:204-205
            else l = a[c] / pivot;                                      // dgetrf2.f:207-209
            a[c] = l;
            if (c + 1 < BW) {
                if ((c & 1) == 0) a[c + 1] = fma(-l, rb[c + 1], a[c + 1]);
#pragma unroll
                for (int q = (c + 2) & ~1; q < BW; q += 2) {
                    const double2 p = *reinterpret_cast<const double2*>(rb + q);
                    a[q] = fma(-l, p.x, a[q]);
                    a[q + 1] = fma(-l, p.y, a[q + 1]);
                }
            }
        }
    }
#pragma unroll
    for (int q = 0; q < BW; ++q) M[mypos + BW * q] = a[q];
    ipiv[id * BW + lane] = myipiv;
    if (lane == 0) info[id] = minfo;
}

// Persistent, software-pipelined form of the one-matrix-per-warp algorithm (lb200_set_batched_mode(1); NOT the default).  Idea: every warp
// loops over matrices, the NEXT matrix is already in flight into a shared-memory stage (cp.async) while the current one is
// factored from registers.  Measured SLOWER than the one-shot kernel (9.12 vs 7.18 ms per 1M matrices, tools/bench_batched2.py):
// the one-shot kernel is not phase-locked on HBM vs LSU as assumed -- it is issue/latency-bound (3286 warp instructions per
// matrix at IPC 1.77: 651 DFMA, 320 LDS, 304 predicated STS of the pivot-row publication, ~1000 instructions of pivot search
// and interchange bookkeeping, 12% of the stall samples are instruction-cache misses of the 52 KB unrolled body;
// profiles/r02_batched_getrf32_ncu_source.txt), and the pipelined form has fewer resident warps (158 registers, 12 warps
// per SM instead of 20) to hide the same dependent chain.
constexpr int PWARPS = 4;
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(gsrc) : "memory");
}
__global__ void __launch_bounds__(PWARPS * 32, 3) getrf_batched32_pipe_kernel(i64 batch, double* __restrict__ A, int* __restrict__ ipiv,
                                                                            int* __restrict__ info) {
    extern __shared__ __align__(16) double pstage[];          // [PWARPS][2][1024] staged matrices
    __shared__ __align__(16) double rowbuf[PWARPS][2][BW];
    __shared__ int posbuf[PWARPS][2];
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const i64 nw = (i64)gridDim.x * PWARPS;
    i64 id = (i64)blockIdx.x * PWARPS + w;
    double* st = pstage + (size_t)w * 2 * (BW * BW);
    if (id < batch) {
        const double* src = A + id * (BW * BW);
#pragma unroll
        for (int q = 0; q < 16; ++q) cp_async16(st + (q * 32 + lane) * 2, src + (q * 32 + lane) * 2);
    }
    asm volatile("cp.async.commit_group;\n" ::: "memory");
    for (int it = 0; id < batch; id += nw, ++it) {
        double* cur = st + (it & 1) * (BW * BW);
        double* nxt = st + ((it + 1) & 1) * (BW * BW);
        const i64 idn = id + nw;
        if (idn < batch) {
            const double* src = A + idn * (BW * BW);
#pragma unroll
            for (int q = 0; q < 16; ++q) cp_async16(nxt + (q * 32 + lane) * 2, src + (q * 32 + lane) * 2);
        }
        asm volatile("cp.async.commit_group;\n" ::: "memory");
        asm volatile("cp.async.wait_group 1;\n" ::: "memory");
        __syncwarp();
        double* M = A + id * (BW * BW);
        double a[BW];
#pragma unroll
        for (int q = 0; q < BW; ++q) a[q] = cur[lane + BW * q];
        __syncwarp();                                          // the stage may be refilled two iterations from now
        int mypos = lane;
        bool done = false;
        int myipiv = 0, minfo = 0;
#pragma unroll
        for (int c = 0; c < BW; ++c) {
            unsigned long long kb = (unsigned long long)__double_as_longlong(fabs(a[c]));
            bool cand = !done;
            if (a[c] != a[c]) {                                   // NaN only wins from the first place (idamax.f:103)
                if (mypos == c) kb = 0x7ff0000000000000ULL; else cand = false;
            }
            const unsigned hi = (unsigned)(kb >> 32), lo = (unsigned)kb;
            const unsigned mh = __reduce_max_sync(full, cand ? hi : 0u);
            const bool c1 = cand && hi == mh;
            unsigned tie = __ballot_sync(full, c1);
            if (__popc(tie) != 1) {
                const unsigned ml = __reduce_max_sync(full, c1 ? lo : 0u);
                const bool c2 = c1 && lo == ml;
                tie = __ballot_sync(full, c2);
                if (__popc(tie) != 1) {
                    const unsigned mp = __reduce_min_sync(full, c2 ? (unsigned)mypos : 0xffffffffu);
                    tie = __ballot_sync(full, c2 && (unsigned)mypos == mp);
                }
            }
            const int who = __ffs(tie) - 1;
            double* rb = rowbuf[w][c & 1];
            if (lane == who) {
                posbuf[w][c & 1] = mypos;
#pragma unroll
                for (int q = c & ~1; q < BW; q += 2) *reinterpret_cast<double2*>(rb + q) = make_double2(a[q], a[q + 1]);
            }
            __syncwarp();
            const int pos = posbuf[w][c & 1];
            if (lane == c) myipiv = pos + 1;
            if (!done && mypos == c && lane != who) mypos = pos;
            if (lane == who) { mypos = c; done = true; }
            const double pivot = rb[c];
            if (pivot == 0.0) {
                if (minfo == 0) minfo = c + 1;                              // dgetrf2.f:212-214
            } else if (!done) {
                double l;
                if (fabs(pivot) >= DBL_MIN) l = a[c] * (1.0 / pivot);       // dgetrf2.f:204-205
                else l = a[c] / pivot;                                      // dgetrf2.f:207-209
                a[c] = l;
                if (c + 1 < BW) {
                    if ((c & 1) == 0) a[c + 1] = fma(-l, rb[c + 1], a[c + 1]);
#pragma unroll
                    for (int q = (c + 2) & ~1; q < BW; q += 2) {
                        const double2 pr = *reinterpret_cast<const double2*>(rb + q);
                        a[q] = fma(-l, pr.x, a[q]);
                        a[q + 1] = fma(-l, pr.y, a[q + 1]);
                    }
                }
            }
        }
#pragma unroll
        for (int q = 0; q < BW; ++q) __stcs(M + mypos + BW * q, a[q]);
        ipiv[id * BW + lane] = myipiv;
        if (lane == 0) info[id] = minfo;
        __syncwarp();                                          // rowbuf / posbuf are reused by the next matrix
    }
}

// ------------------------------------------------------------------------------------------------
// Two matrices per warp (default; lb200_set_batched_mode(2)): a half-warp owns one matrix and every lane holds TWO rows
// (l16, l16+16; 64 doubles).  Every broadcast load of the pivot row feeds two rows, so the shared-memory -> register return volume of
// the one-row kernel (496 doubles x 32 lanes = 127 KB per matrix) halves, and the per-step bookkeeping (pivot search, publication,
// reciprocal, interchange bookkeeping) is issued once for two matrices.  Same arithmetic per element (the same FMAs in the same
// order): factors, IPIV and INFO are bit-identical to the one-row kernel (tools/bench_batched3.py, special cases included).
// Measured 6.67 ms per 1M matrices against 7.16 (profiles/r02_batched_two_per_warp_ab.txt).
constexpr int B2WARPS = 4;
// maximum over each half-warp of a per-lane value: two full-warp reductions (each half contributes 0 to the other one's); a
// redux.sync with a half-warp member mask works too but compiles to a WARPSYNC / ENDCOLLECTIVE pair and was measured slower
__device__ __forceinline__ unsigned half_max_u32(int h, unsigned v) {
    const unsigned m0 = __reduce_max_sync(0xffffffffu, h == 0 ? v : 0u);
    const unsigned m1 = __reduce_max_sync(0xffffffffu, h == 0 ? 0u : v);
    return h == 0 ? m0 : m1;
}
__device__ __forceinline__ unsigned half_min_u32(int h, unsigned v) {
    const unsigned m0 = __reduce_min_sync(0xffffffffu, h == 0 ? v : 0xffffffffu);
    const unsigned m1 = __reduce_min_sync(0xffffffffu, h == 0 ? 0xffffffffu : v);
    return h == 0 ? m0 : m1;
}
// General pivot choice for one half-warp (rare path: the leaders agree in the high word, or an Inf/NaN is among the candidates):
// largest |a|, smallest current position on ties, NaN only from the first place (idamax.f:95-106).  Returns who | slot << 8.
__device__ __noinline__ int batched2_pick_slow(int h, double x0, double x1, bool d0, bool d1, int p0, int p1, int c) {
    const int sh = h * 16;
    unsigned long long k0 = (unsigned long long)__double_as_longlong(fabs(x0)), k1 = (unsigned long long)__double_as_longlong(fabs(x1));
    bool c0 = !d0, c1 = !d1;
    if (x0 != x0) { if (p0 == c) k0 = 0x7ff0000000000000ULL; else c0 = false; }
    if (x1 != x1) { if (p1 == c) k1 = 0x7ff0000000000000ULL; else c1 = false; }
    const bool use1 = c1 && (!c0 || k1 > k0 || (k1 == k0 && p1 < p0));
    const unsigned long long kb = use1 ? k1 : k0;
    const int pb = use1 ? p1 : p0;
    const bool cb = c0 || c1;
    const unsigned hi = (unsigned)(kb >> 32), lo = (unsigned)kb;
    const unsigned mh = half_max_u32(h, cb ? hi : 0u);
    bool t = cb && hi == mh;
    const unsigned ml = half_max_u32(h, t ? lo : 0u);
    t = t && lo == ml;
    const unsigned mp = half_min_u32(h, t ? (unsigned)pb : 0xffffffffu);
    t = t && (unsigned)pb == mp;
    const unsigned b = (__ballot_sync(0xffffffffu, t) >> sh) & 0xffffu;
    const int who = __ffs(b) - 1;
    const int slot = __shfl_sync(0xffffffffu, use1 ? 1 : 0, who + sh);
    return who | (slot << 8);
}
__device__ __noinline__ double batched2_div(double a, double b) { return a / b; }      // |pivot| < SFMIN only (dgetrf2.f:207-209)

// elimination of columns [Q0, Q0 + 8) (those right of c+1): the pivot-row entries are loaded ONCE and used by both row slots; the
// two small blocks are if-converted to predicated DFMAs
template <int Q0, int C>
__device__ __forceinline__ void batched2_update8(double (&a0)[BW], double (&a1)[BW], const double* rb, double l0, double l1, bool d0, bool d1) {
    if (Q0 + 8 <= ((C + 2) & ~1)) return;                       // nothing right of column c+1 in this chunk
    constexpr int QS = (Q0 > ((C + 2) & ~1)) ? Q0 : ((C + 2) & ~1);
    double2 u[4];
#pragma unroll
    for (int q = QS; q < Q0 + 8; q += 2) u[(q - Q0) >> 1] = *reinterpret_cast<const double2*>(rb + q);
    if (!d0) {
#pragma unroll
        for (int q = QS; q < Q0 + 8; q += 2) {
            a0[q] = fma(-l0, u[(q - Q0) >> 1].x, a0[q]);
            a0[q + 1] = fma(-l0, u[(q - Q0) >> 1].y, a0[q + 1]);
        }
    }
    if (!d1) {
#pragma unroll
        for (int q = QS; q < Q0 + 8; q += 2) {
            a1[q] = fma(-l1, u[(q - Q0) >> 1].x, a1[q]);
            a1[q + 1] = fma(-l1, u[(q - Q0) >> 1].y, a1[q + 1]);
        }
    }
}

template <int C>
__device__ __forceinline__ void batched2_step(double (&a0)[BW], double (&a1)[BW], int& pos0, int& pos1, bool& d0, bool& d1, int& ip0, int& ip1,
                                              int& minfo, double* rowbuf_wh, int* posbuf_wh, int h, int l16) {
    const unsigned full = 0xffffffffu;
    const int sh = h * 16;
    constexpr int c = C;
    // ---- pivot search: high words of |a(:,c)| first
    const unsigned h0 = d0 ? 0u : ((unsigned)__double2hiint(a0[c]) & 0x7fffffffu);
    const unsigned h1 = d1 ? 0u : ((unsigned)__double2hiint(a1[c]) & 0x7fffffffu);
    const unsigned mh = half_max_u32(h, max(h0, h1));
    const bool e0 = !d0 && h0 == mh, e1 = !d1 && h1 == mh;
    const unsigned b0 = (__ballot_sync(full, e0) >> sh) & 0xffffu, b1 = (__ballot_sync(full, e1) >> sh) & 0xffffu;
    const bool uniq = (__popc(b0) + __popc(b1) == 1) && mh < 0x7ff00000u;
    int who, slot;
    if (__all_sync(full, uniq)) {
        who = __ffs(b0 | b1) - 1;
        slot = b1 != 0u;
    } else {
        const int r = batched2_pick_slow(h, a0[c], a1[c], d0, d1, pos0, pos1, c);
        who = r & 0xff;
        slot = r >> 8;
    }
    // ---- publish the pivot row (columns >= c, in aligned pairs) and its current position
    double* rb = rowbuf_wh + (c & 1) * BW;
    if (l16 == who) {
        if (slot == 0) {
            posbuf_wh[c & 1] = pos0;
#pragma unroll
            for (int q = c & ~1; q < BW; q += 2) *reinterpret_cast<double2*>(rb + q) = make_double2(a0[q], a0[q + 1]);
            pos0 = c; d0 = true;
        } else {
            posbuf_wh[c & 1] = pos1;
#pragma unroll
            for (int q = c & ~1; q < BW; q += 2) *reinterpret_cast<double2*>(rb + q) = make_double2(a1[q], a1[q + 1]);
            pos1 = c; d1 = true;
        }
    }
    __syncwarp();
    const int pos = posbuf_wh[c & 1];
    if (l16 == (c & 15)) { if (c < 16) ip0 = pos + 1; else ip1 = pos + 1; }
    // interchange bookkeeping: the row sitting at position c moves to `pos` (the pivot row itself is already marked)
    if (!d0 && pos0 == c) pos0 = pos;
    if (!d1 && pos1 == c) pos1 = pos;
    const double2 pv = *reinterpret_cast<const double2*>(rb + (c & ~1));
    const double pivot = (c & 1) ? pv.y : pv.x;
    if (pivot == 0.0) {
        if (minfo == 0) minfo = c + 1;                              // dgetrf2.f:212-214
    } else {
        const double rcp = 1.0 / pivot;
        double l0 = a0[c] * rcp, l1 = a1[c] * rcp;                  // dgetrf2.f:204-205
        if (!(fabs(pivot) >= DBL_MIN)) { l0 = batched2_div(a0[c], pivot); l1 = batched2_div(a1[c], pivot); }
        if (!d0) { a0[c] = l0; if (c + 1 < BW && (c & 1) == 0) a0[c + 1] = fma(-l0, pv.y, a0[c + 1]); }
        if (!d1) { a1[c] = l1; if (c + 1 < BW && (c & 1) == 0) a1[c + 1] = fma(-l1, pv.y, a1[c + 1]); }
        batched2_update8<0, C>(a0, a1, rb, l0, l1, d0, d1);
        batched2_update8<8, C>(a0, a1, rb, l0, l1, d0, d1);
        batched2_update8<16, C>(a0, a1, rb, l0, l1, d0, d1);
        batched2_update8<24, C>(a0, a1, rb, l0, l1, d0, d1);
    }
}
template <int C>
struct Batched2Steps {
    __device__ __forceinline__ static void run(double (&a0)[BW], double (&a1)[BW], int& pos0, int& pos1, bool& d0, bool& d1, int& ip0, int& ip1,
                                               int& minfo, double* rowbuf_wh, int* posbuf_wh, int h, int l16) {
        batched2_step<C>(a0, a1, pos0, pos1, d0, d1, ip0, ip1, minfo, rowbuf_wh, posbuf_wh, h, l16);
        Batched2Steps<C + 1>::run(a0, a1, pos0, pos1, d0, d1, ip0, ip1, minfo, rowbuf_wh, posbuf_wh, h, l16);
    }
};
template <>
struct Batched2Steps<BW> {
    __device__ __forceinline__ static void run(double (&)[BW], double (&)[BW], int&, int&, bool&, bool&, int&, int&, int&, double*, int*, int, int) {}
};

__global__ void __launch_bounds__(B2WARPS * 32, 3) getrf_batched32_two_kernel(i64 batch, double* __restrict__ A, int* __restrict__ ipiv,
                                                                             int* __restrict__ info) {
    __shared__ __align__(16) double rowbuf[B2WARPS][2][2][BW];
    __shared__ int posbuf[B2WARPS][2][2];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int h = lane >> 4, l16 = lane & 15;
    const i64 id = ((i64)blockIdx.x * B2WARPS + w) * 2 + h;
    const bool valid = id < batch;
    if (((i64)blockIdx.x * B2WARPS + w) * 2 >= batch) return;          // whole warp idle
    double* M = A + (valid ? id : 0) * (BW * BW);
    double a0[BW], a1[BW];
#pragma unroll
    for (int q = 0; q < BW; ++q) {
        a0[q] = valid ? M[l16 + BW * q] : 0.0;
        a1[q] = valid ? M[l16 + 16 + BW * q] : 0.0;
    }
    int pos0 = l16, pos1 = l16 + 16;             // current positions of my rows under LAPACK's explicit interchanges
    bool d0 = false, d1 = false;                 // row already used as a pivot row
    int ip0 = 0, ip1 = 0, minfo = 0;
    Batched2Steps<0>::run(a0, a1, pos0, pos1, d0, d1, ip0, ip1, minfo, &rowbuf[w][h][0][0], &posbuf[w][h][0], h, l16);
    if (valid) {
#pragma unroll
        for (int q = 0; q < BW; ++q) { M[pos0 + BW * q] = a0[q]; M[pos1 + BW * q] = a1[q]; }
        ipiv[id * BW + l16] = ip0;
        ipiv[id * BW + 16 + l16] = ip1;
        if (l16 == 0) info[id] = minfo;
    }
}

static int g_batched_mode = 2;     // 2 = two matrices per warp (default), 0 = one matrix per warp, 1 = persistent pipelined one-matrix kernel
void batched_set_mode(int m) { g_batched_mode = m; }

void getrf_batched_32(cudaStream_t s, i64 batch, double* A, int* ipiv, int* info) {
    if (batch <= 0) return;
    if (g_batched_mode == 1) {
        const size_t smem = sizeof(double) * PWARPS * 2 * BW * BW;     // 64 KB
        static bool attr = false;
        if (!attr) {
            LB_CUDA_CHECK(cudaFuncSetAttribute(getrf_batched32_pipe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            attr = true;
        }
        const i64 want = (batch + PWARPS - 1) / PWARPS;
        const int grid = (int)(want < (i64)num_sms() * 3 ? want : (i64)num_sms() * 3);
        getrf_batched32_pipe_kernel<<<grid, PWARPS * 32, smem, s>>>(batch, A, ipiv, info);
    } else if (g_batched_mode == 2) {
        const i64 per_cta = 2 * B2WARPS;
        getrf_batched32_two_kernel<<<(unsigned)((batch + per_cta - 1) / per_cta), B2WARPS * 32, 0, s>>>(batch, A, ipiv, info);
    } else {
        const int wpb = BWARPS;
        getrf_batched32_kernel<<<(unsigned)((batch + wpb - 1) / wpb), wpb * 32, 0, s>>>(batch, A, ipiv, info);
    }
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
