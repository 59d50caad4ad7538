// trsv_stream.cu -- ONE persistent kernel per triangular solve with few right-hand sides (DGETRS / DPOTRS at
// NRHS <= 8; SRC/dgetrs.f:187-198 and :205-217, SRC/dpotrs.f:168-196; arithmetic of BLAS/SRC/dtrsm.f:278-327).
//
// The solve is HBM-bound (the triangle is read once: 4 n^2 bytes) with a latency-bound dependent chain along the
// diagonal.  Formulation: the unknowns are cut into blocks of 128; a CTA claims block i through an atomic counter (in
// chain order, which makes the spin-waits below deadlock-free: a waiting CTA only ever waits for a block that is
// already running), streams the whole block row (block column for op(A) = A^T) of the triangle through registers while
// the earlier unknowns x_j become available, reduces, solves its 128 x 128 diagonal block from shared memory and
// publishes x_i.  All SMs stream different block rows at the same time, so the triangle is read once at full
// bandwidth while the diagonal chain advances; only the LAST 128 columns of a block row are on the chain, and their
// matrix entries are already in registers when x_{i-1} arrives (look-ahead).
//
// Exchange: x travels in self-validating 16-byte packets {hi32(x) : tag, lo32(x) : tag}; both 8-byte halves carry the
// launch tag, so a reader needs no separate flag and no ordering between the halves -- one L2 round trip per
// hand-over.  The packet array is zeroed before the launch (tag 0 never matches).  Per CTA ONE feeder warp polls global
// memory (first a single probe packet, with a back-off that grows with the distance from the head of the chain, so the
// L2 slice that holds the newest packets is not hammered by 148 x 16 warps) and hands the unknowns to the 16 worker warps
// through a shared-memory ring.
//
// Diagonal block (the only work on the chain).  Off the chain, while the block row streams, the CTA inverts the two
// 64 x 64 triangular diagonal sub-blocks in shared memory.  On the chain each half is then
//     x0 = inv(T) r;   x = x0 + inv(T) (r - T x0)          (one step of iterative refinement against T itself)
// i.e. three 64 x 64 matrix-vector products spread over all 512 worker threads instead of 64 dependent substitution
// steps; the refinement step restores the componentwise backward stability of substitution (Skeel).  Blocks whose
// diagonal holds a zero, a denormal or a non-finite entry (or whose inverse overflows) take the plain substitution
// path with dtrsm.f's divisions, so Inf/NaN propagate exactly like the reference.
#include "lb_internal.h"
#include <cfloat>

namespace lb {
namespace {

constexpr int SV_BS = 128;        // unknowns per chain step / per CTA claim
constexpr int SV_H = 64;          // inverted diagonal sub-block
constexpr int SV_WORKERS = 512;   // 16 worker warps
constexpr int SV_THREADS = SV_WORKERS + 32;   // + 1 feeder warp
constexpr int SV_DEPTH = 4;       // ring of 128-unknown units between the feeder and the workers
constexpr int SV_LDS = SV_BS + 1; // leading dimension of the diagonal block in shared memory
constexpr int SV_LDI = 72;        // leading dimension of the inverted 64 x 64 blocks (conflict-free 8-lane row reads)

struct SvParams {
    int n, nrhs, nblk, unit;
    const double* A;
    i64 lda;
    double* B;
    i64 ldb;
    unsigned tag;
    unsigned* counter;
    ulonglong2* pkt;      // [nrhs][n]
};

__device__ __forceinline__ bool try_packet(const ulonglong2* p, unsigned tag, double& x) {
    unsigned long long a, b;
    asm volatile("ld.relaxed.gpu.global.v2.u64 {%0,%1}, [%2];\n" : "=l"(a), "=l"(b) : "l"(p) : "memory");
    x = __hiloint2double((int)(a >> 32), (int)(b >> 32));
    return (unsigned)a == tag && (unsigned)b == tag;
}
__device__ __forceinline__ void publish_packet(ulonglong2* p, double x, unsigned tag) {
    const unsigned long long a = ((unsigned long long)(unsigned)__double2hiint(x) << 32) | tag;
    const unsigned long long b = ((unsigned long long)(unsigned)__double2loint(x) << 32) | tag;
    asm volatile("st.relaxed.gpu.global.v2.u64 [%0], {%1,%2};\n" ::"l"(p), "l"(a), "l"(b) : "memory");
}
__device__ __forceinline__ void worker_bar() { asm volatile("bar.sync 1, %0;\n" ::"n"(SV_WORKERS) : "memory"); }
__device__ __forceinline__ int ld_volatile_s32(const int* p) { return *reinterpret_cast<const volatile int*>(p); }
__device__ __forceinline__ void st_volatile_s32(int* p, int v) { *reinterpret_cast<volatile int*>(p) = v; }

// 64 x 64 matrix-vector product from shared memory by the 512 worker threads: thread = (row = wt >> 3, part = wt & 7),
// part p takes the columns p, p + 8, ...; the 8 partial sums of a row are combined by xor-shuffles (fixed order).
template <int NR>
__device__ __forceinline__ void gemv64(const double* __restrict__ M, int ldm, const double (*v)[NR], int row, int part, double out[NR]) {
#pragma unroll
    for (int r = 0; r < NR; ++r) out[r] = 0.0;
    const double* mr = M + row * ldm + part;
#pragma unroll
    for (int kk = 0; kk < 8; ++kk) {
        const double m = mr[8 * kk];
#pragma unroll
        for (int r = 0; r < NR; ++r) out[r] = fma(m, v[part + 8 * kk][r], out[r]);
    }
#pragma unroll
    for (int r = 0; r < NR; ++r) {
        out[r] += __shfl_xor_sync(0xffffffffu, out[r], 1);
        out[r] += __shfl_xor_sync(0xffffffffu, out[r], 2);
        out[r] += __shfl_xor_sync(0xffffffffu, out[r], 4);
    }
}

// TRANS = false: op(A) = A, FWD <=> A lower.   TRANS = true: op(A) = A^T, FWD <=> A upper.
// Inside a block the unknowns are indexed through perm(i) = FWD ? i : 127 - i, which turns every diagonal block into a
// LOWER triangular system, so the on-chain code exists once.
template <int NR, bool TRANS, bool FWD>
__global__ void __launch_bounds__(SV_THREADS, 1) trsv_stream_kernel(SvParams p) {
    extern __shared__ double sv_smem[];
    double* S = sv_smem;                                                         // [128][129]  S[pi][pj] = op(A)(i,j)
    double* Ti = S + SV_BS * SV_LDS;                                             // [2][64][72] inverses of the diagonal 64-blocks
    double (*bs)[NR] = reinterpret_cast<double (*)[NR]>(Ti + 2 * SV_H * SV_LDI); // [128] right-hand side -> solution (permuted index)
    double (*xt)[NR] = bs + SV_BS;                                               // [64] x0 of the current half
    double (*rt)[NR] = xt + SV_H;                                                // [64] residual of the current half
    double* rinv = reinterpret_cast<double*>(rt + SV_H);                         // [128]
    double* red = rinv + SV_BS;                                                  // [4][128][NR] partial sums (op = N)
    double (*xs)[SV_BS][NR] = reinterpret_cast<double (*)[SV_BS][NR]>(red + 4 * SV_BS * NR);   // [DEPTH][128] ring
    __shared__ int s_blk, s_avail, s_exc;
    __shared__ int s_prog[16];
    constexpr bool STORED_LOWER = (FWD != TRANS);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n = p.n;
    const i64 lda = p.lda;
    auto perm = [](int i) { return FWD ? i : SV_BS - 1 - i; };

    for (;;) {
        __syncthreads();
        if (tid == 0) { s_blk = (int)atomicAdd(p.counter, 1u); s_avail = 0; s_exc = 0; }
        if (tid < 16) s_prog[tid] = 0;
        __syncthreads();
        const int blk = s_blk;
        if (blk >= p.nblk) break;
        const int bi = FWD ? blk : p.nblk - 1 - blk;
        const int r0 = bi * SV_BS;
        const int mrows = min(SV_BS, n - r0);
        const int nunits = blk;                       // blocks earlier in chain order: blk of them
        // unit u of this block row = chain block u = matrix block (FWD ? u : nblk-1-u)

        if (warp == 16) {
            // ================================================= feeder warp
            for (int u = 0; u < nunits; ++u) {
                for (;;) {                                                    // ring slot free?
                    int m = ld_volatile_s32(&s_prog[lane & 15]);
                    m = __reduce_min_sync(0xffffffffu, m);
                    if (u - m < SV_DEPTH) break;
                    __nanosleep(40);
                }
                const int jb = FWD ? u : p.nblk - 1 - u;
                const int j0 = jb * SV_BS;
                const int jrows = min(SV_BS, n - j0);
                // probe ONE packet until it carries the tag (all 128 are published together); back off with the distance
                // from the head of the chain -- only the next block in chain order is latency-critical
                {
                    const int dist = blk - u;
                    const unsigned nap = dist <= 1 ? 0u : (unsigned)min(2000, 100 * (dist - 1));
                    double dummy;
                    while (!try_packet(p.pkt + j0 + jrows - 1, p.tag, dummy)) {
                        if (nap) __nanosleep(nap);
                    }
                }
                double (*slot)[NR] = xs[u % SV_DEPTH];
#pragma unroll
                for (int r = 0; r < NR; ++r) {
                    if (r < p.nrhs) {
                        double xv[4];
                        bool ok[4];
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const int row = lane + 32 * q;
                            ok[q] = row >= jrows || try_packet(p.pkt + (i64)r * n + j0 + row, p.tag, xv[q]);
                            if (row >= jrows) xv[q] = 0.0;
                        }
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const int row = lane + 32 * q;
                            while (!ok[q]) ok[q] = try_packet(p.pkt + (i64)r * n + j0 + row, p.tag, xv[q]);
                            slot[row][r] = xv[q];
                        }
                    } else {
#pragma unroll
                        for (int q = 0; q < 4; ++q) slot[lane + 32 * q][r] = 0.0;
                    }
                }
                __syncwarp();
                __threadfence_block();
                if (lane == 0) st_volatile_s32(&s_avail, u + 1);
            }
            continue;      // next claim (joins the workers at the barrier on top of the loop)
        }

        // ===================================================== worker warps
        // ---- own right-hand side rows (registers) and the diagonal block (shared memory); both are off the chain
        double bown = 0.0;
        if (tid < SV_BS * NR) {
            const int row = tid & (SV_BS - 1), r = tid >> 7;
            if (row < mrows && r < p.nrhs) bown = p.B[(r0 + row) + (i64)r * p.ldb];
        }
        {
            const double* Ad = p.A + r0 + (i64)r0 * lda;
#pragma unroll 8
            for (int idx = tid; idx < SV_BS * SV_BS; idx += SV_WORKERS) {
                const int ii = idx & (SV_BS - 1), kk = idx >> 7;      // stored element A(ii, kk) of the block
                double v = 0.0;
                const bool inb = ii < mrows && kk < mrows;
                if (ii == kk) v = (p.unit || !inb) ? 1.0 : __ldg(Ad + ii + (i64)kk * lda);     // padding rows: identity
                else if (inb && (STORED_LOWER ? ii > kk : ii < kk)) v = __ldg(Ad + ii + (i64)kk * lda);
                const int oi = TRANS ? kk : ii, oj = TRANS ? ii : kk;                         // op(A)(oi, oj)
                S[perm(oi) * SV_LDS + perm(oj)] = v;
            }
        }
        worker_bar();
        if (tid < SV_BS) {
            const double d = S[tid * SV_LDS + tid];
            rinv[tid] = 1.0 / d;
            if (!(fabs(d) >= DBL_MIN) || !(fabs(d) <= DBL_MAX)) s_exc = 1;      // zero, denormal, Inf or NaN on the diagonal
        }
        worker_bar();
        // inverses of the two 64 x 64 lower-triangular diagonal blocks: thread = column c of the inverse (warps 0..3),
        // forward substitution against the unit vector e_c
        if (tid < 2 * SV_H && !s_exc) {
            const int h = tid >> 6, c = tid & (SV_H - 1);
            const double* M = S + (h * SV_H) * SV_LDS + h * SV_H;
            double* Z = Ti + h * SV_H * SV_LDI;
            const double* ri = rinv + h * SV_H;
            bool bad = false;
            for (int k = 0; k < SV_H; ++k) {
                double a0 = 0.0, a1 = 0.0;
                int j = 0;
                for (; j + 1 < k; j += 2) {
                    a0 = fma(M[k * SV_LDS + j], Z[j * SV_LDI + c], a0);
                    a1 = fma(M[k * SV_LDS + j + 1], Z[(j + 1) * SV_LDI + c], a1);
                }
                if (j < k) a0 = fma(M[k * SV_LDS + j], Z[j * SV_LDI + c], a0);
                const double z = ((k == c ? 1.0 : 0.0) - (a0 + a1)) * ri[k];
                Z[k * SV_LDI + c] = z;
                bad |= !(fabs(z) <= DBL_MAX);
            }
            if (bad) s_exc = 1;
        }

        // ---- stream the off-diagonal part: one 128-unknown unit at a time, in the order in which the units become available
        if (!TRANS) {
            // thread = (row of the block, group g of 16 consecutive columns of each 64-column half-unit)
            const int row = tid & (SV_BS - 1), g = tid >> 7;
            const bool rok = row < mrows;
            double acc[NR];
#pragma unroll
            for (int r = 0; r < NR; ++r) acc[r] = 0.0;
            const double* arow = p.A + (r0 + row);
#pragma unroll 1
            for (int u = 0; u < nunits; ++u) {
                const int j0 = (FWD ? u : p.nblk - 1 - u) * SV_BS;
                double v[2][16];
#pragma unroll
                for (int ch = 0; ch < 2; ++ch) {
                    const int c0 = j0 + ch * 64 + g * 16;
                    const int nv = n - c0;
                    const double* ap = arow + (i64)c0 * lda;
#pragma unroll
                    for (int k = 0; k < 16; ++k) v[ch][k] = (rok && k < nv) ? __ldcs(ap + (i64)k * lda) : 0.0;
                }
                while (ld_volatile_s32(&s_avail) <= u) __nanosleep(20);
                const double (*slot)[NR] = xs[u % SV_DEPTH];
#pragma unroll
                for (int ch = 0; ch < 2; ++ch) {
#pragma unroll
                    for (int k = 0; k < 16; ++k) {
#pragma unroll
                        for (int r = 0; r < NR; ++r) acc[r] = fma(v[ch][k], slot[ch * 64 + g * 16 + k][r], acc[r]);
                    }
                }
                __syncwarp();
                if (lane == 0) st_volatile_s32(&s_prog[warp], u + 1);
            }
#pragma unroll
            for (int r = 0; r < NR; ++r) red[(g * SV_BS + row) * NR + r] = acc[r];
            worker_bar();
            if (tid < SV_BS * NR) {
                const int rw = tid & (SV_BS - 1), r = tid >> 7;
                double t = red[(0 * SV_BS + rw) * NR + r];
#pragma unroll
                for (int gg = 1; gg < 4; ++gg) t += red[(gg * SV_BS + rw) * NR + r];
                bs[perm(rw)][r] = bown - t;
            }
        } else {
            // warp = 8 consecutive columns of the block (= unknowns), lane = rows lane, lane + 32, ... of the unit
            const int cb = warp * 8;
            double acc[8][NR];
#pragma unroll
            for (int k = 0; k < 8; ++k)
#pragma unroll
                for (int r = 0; r < NR; ++r) acc[k][r] = 0.0;
            const double* acol = p.A + (i64)(r0 + cb) * lda;
#pragma unroll 1
            for (int u = 0; u < nunits; ++u) {
                const int j0 = (FWD ? u : p.nblk - 1 - u) * SV_BS;
#pragma unroll
                for (int hf = 0; hf < 2; ++hf) {                  // two half-units keep the register count at 16 loads in flight
                    const int ra = j0 + hf * 64 + lane, rb = ra + 32;
                    const bool oka = ra < n, okb = rb < n;
                    double va[8], vb[8];
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const bool cok = cb + k < mrows;
                        va[k] = (oka && cok) ? __ldcs(acol + ra + (i64)k * lda) : 0.0;
                        vb[k] = (okb && cok) ? __ldcs(acol + rb + (i64)k * lda) : 0.0;
                    }
                    if (hf == 0) { while (ld_volatile_s32(&s_avail) <= u) __nanosleep(20); }
                    const double (*slot)[NR] = xs[u % SV_DEPTH];
                    double xa[NR], xb[NR];
#pragma unroll
                    for (int r = 0; r < NR; ++r) { xa[r] = slot[hf * 64 + lane][r]; xb[r] = slot[hf * 64 + lane + 32][r]; }
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
#pragma unroll
                        for (int r = 0; r < NR; ++r) acc[k][r] = fma(vb[k], xb[r], fma(va[k], xa[r], acc[k][r]));
                    }
                }
                __syncwarp();
                if (lane == 0) st_volatile_s32(&s_prog[warp], u + 1);
            }
#pragma unroll
            for (int k = 0; k < 8; ++k) {
#pragma unroll
                for (int r = 0; r < NR; ++r) {
                    double t = acc[k][r];
#pragma unroll
                    for (int off = 16; off > 0; off >>= 1) t += __shfl_xor_sync(0xffffffffu, t, off);
                    if (lane == 0) red[(cb + k) * NR + r] = t;
                }
            }
            worker_bar();
            if (tid < SV_BS * NR) {
                const int rw = tid & (SV_BS - 1), r = tid >> 7;
                bs[perm(rw)][r] = bown - red[rw * NR + r];
            }
        }
        worker_bar();

        // ---- diagonal block (lower triangular in the permuted index), ON THE CHAIN
        if (!s_exc) {
            const int row = tid >> 3, part = tid & 7;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int q0 = h * SV_H;
                const double* Z = Ti + h * SV_H * SV_LDI;
                double t[NR];
                gemv64<NR>(Z, SV_LDI, bs + q0, row, part, t);                              // x0 = inv(T) r
                if (part == 0) {
#pragma unroll
                    for (int r = 0; r < NR; ++r) xt[row][r] = t[r];
                }
                worker_bar();
                gemv64<NR>(S + q0 * SV_LDS + q0, SV_LDS, xt, row, part, t);                // T x0
                if (part == 0) {
#pragma unroll
                    for (int r = 0; r < NR; ++r) rt[row][r] = bs[q0 + row][r] - t[r];
                }
                worker_bar();
                gemv64<NR>(Z, SV_LDI, rt, row, part, t);                                   // correction
                if (part == 0) {
#pragma unroll
                    for (int r = 0; r < NR; ++r) bs[q0 + row][r] = xt[row][r] + t[r];
                }
                worker_bar();
                if (h == 0) {
                    gemv64<NR>(S + SV_H * SV_LDS, SV_LDS, bs, row, part, t);               // r2 -= T21 x1
                    if (part == 0) {
#pragma unroll
                        for (int r = 0; r < NR; ++r) bs[SV_H + row][r] -= t[r];
                    }
                    worker_bar();
                }
            }
        } else {
            // exceptional diagonal: substitution in 32-row steps with dtrsm.f's divisions (warp 0), the other warps update
            // the rest of the block
            for (int sb = 0; sb < SV_BS / 32; ++sb) {
                const int q0 = sb * 32;
                if (warp == 0) {
                    double x[NR];
#pragma unroll
                    for (int r = 0; r < NR; ++r) x[r] = bs[q0 + lane][r];
#pragma unroll 4
                    for (int j = 0; j < 32; ++j) {
                        const double tij = S[(q0 + lane) * SV_LDS + q0 + j];
                        const double d = S[(q0 + j) * SV_LDS + q0 + j];
#pragma unroll
                        for (int r = 0; r < NR; ++r) {
                            double xj = __shfl_sync(0xffffffffu, x[r], j);
                            if (!p.unit) xj = xj / d;                       // dtrsm.f: B(k,j) = B(k,j)/A(k,k)
                            if (lane == j) x[r] = xj;
                            else if (lane > j) x[r] = x[r] - xj * tij;      // dtrsm.f: B(i,j) = B(i,j) - B(k,j)*A(i,k)
                        }
                    }
#pragma unroll
                    for (int r = 0; r < NR; ++r) bs[q0 + lane][r] = x[r];
                }
                worker_bar();
                const int i = q0 + 32 + tid - 32;
                if (warp != 0 && i < SV_BS) {
                    double a2[NR];
#pragma unroll
                    for (int r = 0; r < NR; ++r) a2[r] = 0.0;
                    for (int j = 0; j < 32; ++j) {
                        const double t = S[i * SV_LDS + q0 + j];
#pragma unroll
                        for (int r = 0; r < NR; ++r) a2[r] = fma(t, bs[q0 + j][r], a2[r]);
                    }
#pragma unroll
                    for (int r = 0; r < NR; ++r) bs[i][r] -= a2[r];
                }
                worker_bar();
            }
        }

        // ---- publish x_i (packets for the other CTAs, B for the caller)
        if (tid < SV_BS * NR) {
            const int row = tid & (SV_BS - 1), r = tid >> 7;
            if (row < mrows && r < p.nrhs) {
                const double x = bs[perm(row)][r];
                publish_packet(p.pkt + (i64)r * n + r0 + row, x, p.tag);
                p.B[(r0 + row) + (i64)r * p.ldb] = x;
            }
        }
    }
}

template <int NR>
constexpr size_t sv_smem_bytes() {
    return sizeof(double) * (SV_BS * SV_LDS + 2 * SV_H * SV_LDI + SV_BS * NR + 2 * SV_H * NR + SV_BS + 4 * SV_BS * NR +
                             SV_DEPTH * SV_BS * NR);
}

template <int NR, bool TRANS, bool FWD>
void launch_sv(cudaStream_t s, const SvParams& p) {
    const size_t smem = sv_smem_bytes<NR>();
    static bool attr = false;
    if (!attr) {
        LB_CUDA_CHECK(cudaFuncSetAttribute(trsv_stream_kernel<NR, TRANS, FWD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr = true;
    }
    const int grid = min(p.nblk, num_sms());
    trsv_stream_kernel<NR, TRANS, FWD><<<grid, SV_THREADS, smem, s>>>(p);
    count_launch();
}

template <int NR>
void dispatch_sv(cudaStream_t s, bool trans, bool fwd, const SvParams& p) {
    if (!trans) { if (fwd) launch_sv<NR, false, true>(s, p); else launch_sv<NR, false, false>(s, p); }
    else { if (fwd) launch_sv<NR, true, true>(s, p); else launch_sv<NR, true, false>(s, p); }
}

}  // namespace

// op(A) X = B, A n x n triangular, B n x nrhs (nrhs small); X overwrites B.  Right-hand sides go two at a time.
void trsv_stream(cudaStream_t s, bool upper, bool trans, bool unit, int n, int nrhs, const double* A, i64 lda, double* B,
                 i64 ldb) {
    if (n <= 0 || nrhs <= 0) return;
    const size_t pkt_bytes = sizeof(ulonglong2) * 2 * (size_t)n;
    char* ws = (char*)ws_alloc(s, pkt_bytes + 64);
    const bool fwd = (upper == trans);       // (L,N) and (U,T) substitute forward
    for (int r = 0; r < nrhs; r += 2) {
        const int nr = min(2, nrhs - r);
        LB_CUDA_CHECK(cudaMemsetAsync(ws, 0, pkt_bytes + 64, s));
        SvParams p;
        p.n = n; p.nrhs = nr; p.nblk = (n + SV_BS - 1) / SV_BS; p.unit = unit ? 1 : 0;
        p.A = A; p.lda = lda; p.B = B + (i64)r * ldb; p.ldb = ldb;
        p.tag = 1u;
        p.pkt = (ulonglong2*)ws;
        p.counter = (unsigned*)(ws + pkt_bytes);
        if (nr == 1) dispatch_sv<1>(s, trans, fwd, p); else dispatch_sv<2>(s, trans, fwd, p);
    }
    ws_free(s, ws);
    LB_CUDA_CHECK(cudaGetLastError());
}

}  // namespace lb
