// gecon.cu -- condition estimation and equilibration on top of the LU factors:
//   DLATRS (SRC/dlatrs.f:250-850)  triangular solve with scaling against overflow,
//   DGECON (SRC/dgecon.f:128-285)  reciprocal condition number from DGETRF's factors (Hager/Higham estimator DLACN2),
//   DGEEQU / DLAQGE (SRC/dgeequ.f:160-310, SRC/dlaqge.f:160-230) row / column equilibration, and the norms DGESVX needs.
//
// Split between host and device: every O(n^2) pass (column norms of the triangle, the solves, the scalings, the norms) runs
// on the device; the O(n) scalar logic -- DLATRS's growth bound that decides between the plain solve and the scaled one,
// DLACN2's reverse-communication state machine -- runs on the host on n-vectors, exactly as the reference orders it.
// The plain solve is the persistent streaming kernel of trsv_stream.cu; the scaled Level-1 algorithm (needed only when the
// bound says the solution could overflow) is one CTA walking the columns with the reference's rescaling rules.
#include "lb_internal.h"
#include <cfloat>
#include <cmath>
#include <vector>

namespace lb {
void trsv_stream(cudaStream_t s, bool upper, bool trans, bool unit, int n, int nrhs, const double* A, i64 lda, double* B, i64 ldb);

namespace {

// cnorm[j] = tscal * sum of |A(i,j)| over the strictly upper / lower part of column j (dlatrs.f:300-317, :368-384); one warp per
// column; with a mask only the flagged columns are recomputed
__global__ void tri_cnorm_kernel(int n, const double* __restrict__ A, i64 lda, bool upper, double tscal, const unsigned char* __restrict__ mask,
                                 double* __restrict__ cnorm) {
    const int j = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (j >= n) return;
    if (mask && !mask[j]) return;
    const int i0 = upper ? 0 : j + 1, i1 = upper ? j : n;
    const double* col = A + (i64)j * lda;
    double acc = 0.0;
    for (int i = i0 + lane; i < i1; i += 32) acc += tscal * fabs(col[i]);
    for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
    if (lane == 0) cnorm[j] = acc;
}
// max |A(i,j)| over the strict triangle (dlatrs.f:343-360), bit pattern of a non-negative double orders like an integer
__global__ void tri_amax_kernel(int n, const double* __restrict__ A, i64 lda, bool upper, unsigned long long* out) {
    const int j = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (j >= n) return;
    const int i0 = upper ? 0 : j + 1, i1 = upper ? j : n;
    const double* col = A + (i64)j * lda;
    double m = 0.0;
    for (int i = i0 + lane; i < i1; i += 32) { const double t = fabs(col[i]); if (m < t || t != t) m = t; }
    for (int off = 16; off > 0; off >>= 1) { const double t = __shfl_xor_sync(0xffffffffu, m, off); if (m < t || t != t) m = t; }
    if (lane == 0) atomicMax(out, (unsigned long long)__double_as_longlong(m));
}
__global__ void diag_gather_kernel(int n, const double* __restrict__ A, i64 lda, double* __restrict__ d) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) d[i] = A[i + (i64)i * lda];
}
__global__ void vec_scal_kernel(int n, double a, double* __restrict__ x) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) x[i] *= a;
}

// ---- scaled Level-1 solve (dlatrs.f:568-840): ONE CTA, x in global memory, scalars in shared memory
constexpr int LT_THREADS = 1024;
__device__ double block_max(double v, double* red) {
    for (int off = 16; off > 0; off >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, off));
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double r = red[0];
    for (int w = 1; w < LT_THREADS / 32; ++w) r = fmax(r, red[w]);
    return r;
}
__device__ double block_sum(double v, double* red) {
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double r = 0.0;
    for (int w = 0; w < LT_THREADS / 32; ++w) r += red[w];
    return r;
}
struct LatrsParams {
    int n, upper, notran, nounit;
    const double* A;
    i64 lda;
    double* x;
    const double* cnorm;
    double tscal, smlnum, bignum;
    double* scale_out;
};
__global__ void __launch_bounds__(LT_THREADS) latrs_scaled_kernel(LatrsParams p) {
    __shared__ double red[LT_THREADS / 32];
    __shared__ double s_mul, s_xj, s_uscal, s_tjjs;
    __shared__ int s_zero, s_divmode;
    const int tid = threadIdx.x, n = p.n;
    double* x = p.x;
    double scale = 1.0;                                    // replicated in every thread (all decisions are uniform)
    double lm = 0.0;
    for (int i = tid; i < n; i += LT_THREADS) lm = fmax(lm, fabs(x[i]));
    double xmax = block_max(lm, red);
    if (xmax > p.bignum) {                                 // dlatrs.f:572-580
        scale = p.bignum / xmax;
        for (int i = tid; i < n; i += LT_THREADS) x[i] *= scale;
        xmax = p.bignum;
    }
    __syncthreads();
    const bool fwd = p.notran ? !p.upper : p.upper;        // column order: dlatrs.f:423-431 / :502-510
    for (int jj = 0; jj < n; ++jj) {
        const int j = fwd ? jj : n - 1 - jj;
        const double* col = p.A + (i64)j * p.lda;
        const double cj = p.cnorm[j];
        // the entries of x this step combines with column j: the part of the column inside the triangle
        const int i0 = p.upper ? 0 : j + 1, i1 = p.upper ? j : n;
        if (p.notran) {
            // ---- x(j) := x(j) / A(j,j) with rescaling of the whole vector (dlatrs.f:588-661), then the overflow guard of the
            // update (:666-682); one thread decides, everybody applies
            if (tid == 0) {
                double mul = 1.0, xjv = x[j], xj = fabs(xjv), xm = xmax;
                int zero = 0;
                double tjjs = p.nounit ? col[j] * p.tscal : p.tscal;
                const bool skip = !p.nounit && p.tscal == 1.0;
                if (!skip) {
                    const double tjj = fabs(tjjs);
                    if (tjj > p.smlnum) {
                        if (tjj < 1.0 && xj > tjj * p.bignum) { const double rec = 1.0 / xj; mul *= rec; xjv *= rec; xm *= rec; }
                        xjv = xjv / tjjs;
                        xj = fabs(xjv);
                    } else if (tjj > 0.0) {
                        if (xj > tjj * p.bignum) {
                            double rec = (tjj * p.bignum) / xj;
                            if (cj > 1.0) rec = rec / cj;
                            mul *= rec; xjv *= rec; xm *= rec;
                        }
                        xjv = xjv / tjjs;
                        xj = fabs(xjv);
                    } else { zero = 1; xjv = 1.0; xj = 1.0; xm = 0.0; }
                }
                if (!zero) {
                    if (xj > 1.0) {
                        double rec = 1.0 / xj;
                        if (cj > (p.bignum - xm) * rec) { rec *= 0.5; mul *= rec; xjv *= rec; }
                    } else if (xj * cj > p.bignum - xm) { mul *= 0.5; xjv *= 0.5; }
                }
                s_mul = mul; s_xj = xjv; s_zero = zero;
            }
            __syncthreads();
            const double mul = s_mul, xjv = s_xj;
            const int zero = s_zero;
            if (zero) {                                     // A(j,j) = 0: x := e_j, scale := 0 (dlatrs.f:652-660)
                for (int i = tid; i < n; i += LT_THREADS) x[i] = (i == j) ? 1.0 : 0.0;
                scale = 0.0;
            } else if (mul != 1.0) {
                for (int i = tid; i < n; i += LT_THREADS) if (i != j) x[i] *= mul;
                scale *= mul;
            }
            __syncthreads();
            if (tid == 0) x[j] = xjv;
            // ---- x(i0:i1) -= x(j) * tscal * A(i0:i1, j), xmax := max |x(i0:i1)| (dlatrs.f:684-706)
            lm = 0.0;
            const double f = -xjv * p.tscal;
            for (int i = i0 + tid; i < i1; i += LT_THREADS) { const double v = fma(f, col[i], x[i]); x[i] = v; lm = fmax(lm, fabs(v)); }
            if (i1 > i0) xmax = block_max(lm, red); else __syncthreads();
        } else {
            // ---- x(j) := (x(j) - sum_k A(k,j) x(k)) / A(j,j) (dlatrs.f:712-838)
            if (tid == 0) {
                const double xj = fabs(x[j]);
                double uscal = p.tscal, mul = 1.0, tjjs = p.nounit ? col[j] * p.tscal : p.tscal;
                double rec = 1.0 / fmax(xmax, 1.0);
                if (cj > (p.bignum - xj) * rec) {
                    rec *= 0.5;
                    const double tjj = fabs(tjjs);
                    if (tjj > 1.0) { rec = fmin(1.0, rec * tjj); uscal = uscal / tjjs; }
                    if (rec < 1.0) mul = rec;
                }
                s_mul = mul; s_uscal = uscal; s_tjjs = tjjs;
            }
            __syncthreads();
            double mul = s_mul;
            const double uscal = s_uscal;
            if (mul != 1.0) {
                for (int i = tid; i < n; i += LT_THREADS) x[i] *= mul;
                scale *= mul; xmax *= mul;
                __syncthreads();
            }
            double ls = 0.0;
            if (uscal == 1.0) { for (int i = i0 + tid; i < i1; i += LT_THREADS) ls = fma(col[i], x[i], ls); }
            else { for (int i = i0 + tid; i < i1; i += LT_THREADS) ls = fma(col[i] * uscal, x[i], ls); }
            const double sumj = block_sum(ls, red);
            if (tid == 0) {
                double m2 = 1.0, xjv;
                int zero = 0;
                if (uscal == p.tscal) {
                    xjv = x[j] - sumj;
                    const double xj = fabs(xjv), tjjs = s_tjjs;
                    const bool skip = !p.nounit && p.tscal == 1.0;
                    if (!skip) {
                        const double tjj = fabs(tjjs);
                        if (tjj > p.smlnum) {
                            if (tjj < 1.0 && xj > tjj * p.bignum) { const double rec = 1.0 / xj; m2 = rec; xjv *= rec; }
                            xjv = xjv / tjjs;
                        } else if (tjj > 0.0) {
                            if (xj > tjj * p.bignum) { const double rec = (tjj * p.bignum) / xj; m2 = rec; xjv *= rec; }
                            xjv = xjv / tjjs;
                        } else { zero = 1; xjv = 1.0; }
                    }
                } else xjv = x[j] / s_tjjs - sumj;
                s_mul = m2; s_xj = xjv; s_zero = zero;
            }
            __syncthreads();
            mul = s_mul;
            const double xjv = s_xj;
            if (s_zero) {
                for (int i = tid; i < n; i += LT_THREADS) x[i] = (i == j) ? 1.0 : 0.0;
                scale = 0.0; xmax = 0.0;
            } else if (mul != 1.0) {
                for (int i = tid; i < n; i += LT_THREADS) if (i != j) x[i] *= mul;
                scale *= mul; xmax *= mul;
            }
            __syncthreads();
            if (tid == 0) x[j] = xjv;
            xmax = fmax(xmax, fabs(xjv));
        }
        __syncthreads();
    }
    (void)s_divmode;
    if (tid == 0) *p.scale_out = scale / p.tscal;           // dlatrs.f:841
}

// ---- equilibration helpers
// r(i) = max_j |A(i,j)| : thread per row, coalesced over rows (dgeequ.f:196-204)
__global__ void row_amax_kernel(int m, int n, const double* __restrict__ A, i64 lda, double* __restrict__ r) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    double v = 0.0;
    for (int j = 0; j < n; ++j) v = fmax(v, fabs(A[i + (i64)j * lda]));
    r[i] = v;
}
// c(j) = max_i |A(i,j)| r(i) : warp per column (dgeequ.f:246-256); r == nullptr means r = 1
__global__ void col_amax_kernel(int m, int n, const double* __restrict__ A, i64 lda, const double* __restrict__ r, double* __restrict__ c) {
    const int j = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (j >= n) return;
    const double* col = A + (i64)j * lda;
    double v = 0.0;
    for (int i = lane; i < m; i += 32) v = fmax(v, fabs(col[i]) * (r ? r[i] : 1.0));
    for (int off = 16; off > 0; off >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, off));
    if (lane == 0) c[j] = v;
}
// A(i,j) := c(j) * r(i) * A(i,j) with the multiplication order of dlaqge.f:186-219 (mode 1 = C, 2 = R, 3 = B)
__global__ void laqge_kernel(int m, int n, double* __restrict__ A, i64 lda, const double* __restrict__ r, const double* __restrict__ c, int mode) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    for (int j = blockIdx.y; j < n; j += gridDim.y) {
        double* a = A + i + (i64)j * lda;
        if (mode == 1) *a = c[j] * (*a);
        else if (mode == 2) *a = r[i] * (*a);
        else *a = c[j] * r[i] * (*a);
    }
}
// B(i,j) := d(i) * B(i,j)
__global__ void row_scale_kernel(int m, int n, double* __restrict__ B, i64 ldb, const double* __restrict__ d) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    const double s = d[i];
    for (int j = blockIdx.y; j < n; j += gridDim.y) B[i + (i64)j * ldb] = s * B[i + (i64)j * ldb];
}
// column sums of |A| (1-norm pieces) and row sums (inf-norm pieces)
__global__ void col_asum_kernel(int m, int n, const double* __restrict__ A, i64 lda, double* __restrict__ c) {
    const int j = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (j >= n) return;
    const double* col = A + (i64)j * lda;
    double v = 0.0;
    for (int i = lane; i < m; i += 32) v += fabs(col[i]);
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    if (lane == 0) c[j] = v;
}
__global__ void row_asum_kernel(int m, int n, const double* __restrict__ A, i64 lda, double* __restrict__ r) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    double v = 0.0;
    for (int j = 0; j < n; ++j) v += fabs(A[i + (i64)j * lda]);
    r[i] = v;
}
// max |A(i,j)| over i <= j, i < m (DLANTR 'M','U','N', dlantr.f:190-200), per column
__global__ void upper_col_amax_kernel(int m, int n, const double* __restrict__ A, i64 lda, double* __restrict__ c) {
    const int j = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (j >= n) return;
    const double* col = A + (i64)j * lda;
    const int rows = min(m, j + 1);
    double v = 0.0;
    for (int i = lane; i < rows; i += 32) { const double t = fabs(col[i]); if (v < t || t != t) v = t; }
    for (int off = 16; off > 0; off >>= 1) { const double t = __shfl_xor_sync(0xffffffffu, v, off); if (v < t || t != t) v = t; }
    if (lane == 0) c[j] = v;
}

std::vector<double> download(cudaStream_t s, const double* d, int n) {
    std::vector<double> h((size_t)n);
    LB_CUDA_CHECK(cudaMemcpyAsync(h.data(), d, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost, s));
    LB_CUDA_CHECK(cudaStreamSynchronize(s));
    return h;
}
double nanmax(const std::vector<double>& v) {           // DLANGE semantics: a NaN wins
    double r = 0.0;
    for (double t : v) if (r < t || t != t) r = t;
    return r;
}

}  // namespace

// ------------------------------------------------------------------------------------------------ DLATRS
// x (device, n) := solution of op(A) x = scale * b, returns scale.  cnorm (device, n) is computed when !normin_y and reused otherwise.
// hdiag: host copy of diag(A) (filled on first use when empty).  Synchronises the stream (host decisions).
double latrs(cudaStream_t s, bool upper, bool notran, bool nounit, bool normin_y, int n, const double* A, i64 lda, double* x,
             double* cnorm, std::vector<double>& hdiag) {
    if (n <= 0) return 1.0;
    const double ovfl = DBL_MAX;
    const double smlnum = DBL_MIN / DBL_EPSILON, bignum = 1.0 / smlnum;      // DLAMCH('S') / DLAMCH('P') (dlatrs.f:293-294)
    const int wpb = 8;
    if (!normin_y) {
        tri_cnorm_kernel<<<ceil_div(n, wpb), wpb * 32, 0, s>>>(n, A, lda, upper, 1.0, nullptr, cnorm);
        count_launch();
    }
    if (nounit && hdiag.empty()) {
        double* dd = (double*)ws_alloc(s, sizeof(double) * (size_t)n);
        diag_gather_kernel<<<ceil_div(n, 256), 256, 0, s>>>(n, A, lda, dd);
        count_launch();
        hdiag = download(s, dd, n);
        ws_free(s, dd);
    }
    std::vector<double> hc = download(s, cnorm, n);
    std::vector<double> hx = download(s, x, n);
    // ---- TSCAL (dlatrs.f:322-392)
    double tmax = 0.0;
    for (int j = 0; j < n; ++j) if (hc[j] > tmax) tmax = hc[j];               // IDAMAX on non-negative entries
    for (int j = 0; j < n; ++j) if (hc[j] != hc[j]) { tmax = hc[j]; break; }
    double tscal = 1.0;
    bool cnorm_scaled = false;
    if (!(tmax <= bignum)) {
        if (tmax <= ovfl) {
            tscal = 1.0 / (smlnum * tmax);
            for (int j = 0; j < n; ++j) hc[j] *= tscal;
            cnorm_scaled = true;
        } else {
            unsigned long long* dmax = (unsigned long long*)ws_alloc(s, 64);
            LB_CUDA_CHECK(cudaMemsetAsync(dmax, 0, 64, s));
            tri_amax_kernel<<<ceil_div(n, wpb), wpb * 32, 0, s>>>(n, A, lda, upper, dmax);
            count_launch();
            unsigned long long bits = 0;
            LB_CUDA_CHECK(cudaMemcpyAsync(&bits, dmax, 8, cudaMemcpyDeviceToHost, s));
            LB_CUDA_CHECK(cudaStreamSynchronize(s));
            ws_free(s, dmax);
            double t2;
            memcpy(&t2, &bits, 8);
            if (t2 <= ovfl) {
                tscal = 1.0 / (smlnum * t2);
                std::vector<unsigned char> mask((size_t)n, 0);
                bool any = false;
                for (int j = 0; j < n; ++j) { if (hc[j] <= ovfl) hc[j] *= tscal; else { mask[j] = 1; any = true; } }
                if (any) {                                                     // recompute without forming Inf (dlatrs.f:368-384)
                    unsigned char* dm = (unsigned char*)ws_alloc(s, (size_t)n);
                    double* dc = (double*)ws_alloc(s, sizeof(double) * (size_t)n);
                    LB_CUDA_CHECK(cudaMemcpyAsync(dm, mask.data(), (size_t)n, cudaMemcpyHostToDevice, s));
                    tri_cnorm_kernel<<<ceil_div(n, wpb), wpb * 32, 0, s>>>(n, A, lda, upper, tscal, dm, dc);
                    count_launch();
                    std::vector<double> h2 = download(s, dc, n);
                    for (int j = 0; j < n; ++j) if (mask[j]) hc[j] = h2[j];
                    ws_free(s, dm); ws_free(s, dc);
                }
                cnorm_scaled = true;
            } else {                                                           // Inf / NaN in A: let the plain solve propagate them
                trsv_stream(s, upper, !notran, !nounit, n, 1, A, lda, x, n);
                return 1.0;
            }
        }
    }
    // ---- bound on the solution (dlatrs.f:398-560)
    double xmax = 0.0;
    for (int i = 0; i < n; ++i) { const double t = fabs(hx[i]); if (t > xmax) xmax = t; }
    double xbnd = xmax, grow;
    const bool fwd = notran ? !upper : upper;
    auto J = [&](int jj) { return fwd ? jj : n - 1 - jj; };
    if (tscal != 1.0) grow = 0.0;
    else if (notran) {
        if (nounit) {
            grow = 1.0 / fmax(xbnd, smlnum);
            xbnd = grow;
            bool broke = false;
            for (int jj = 0; jj < n; ++jj) {
                const int j = J(jj);
                if (grow <= smlnum) { broke = true; break; }
                const double tjj = fabs(hdiag[j]);
                xbnd = fmin(xbnd, fmin(1.0, tjj) * grow);
                if (tjj + hc[j] >= smlnum) grow = grow * (tjj / (tjj + hc[j])); else grow = 0.0;
            }
            if (!broke) grow = xbnd;
        } else {
            grow = fmin(1.0, 1.0 / fmax(xbnd, smlnum));
            for (int jj = 0; jj < n; ++jj) { if (grow <= smlnum) break; grow = grow * (1.0 / (1.0 + hc[J(jj)])); }
        }
    } else {
        if (nounit) {
            grow = 1.0 / fmax(xbnd, smlnum);
            xbnd = grow;
            bool broke = false;
            for (int jj = 0; jj < n; ++jj) {
                const int j = J(jj);
                if (grow <= smlnum) { broke = true; break; }
                const double xj = 1.0 + hc[j];
                grow = fmin(grow, xbnd / xj);
                const double tjj = fabs(hdiag[j]);
                if (xj > tjj) xbnd = xbnd * (tjj / xj);
            }
            if (!broke) grow = fmin(grow, xbnd);
        } else {
            grow = fmin(1.0, 1.0 / fmax(xbnd, smlnum));
            for (int jj = 0; jj < n; ++jj) { if (grow <= smlnum) break; grow = grow / (1.0 + hc[J(jj)]); }
        }
    }
    double scale = 1.0;
    if (grow * tscal > smlnum) {
        trsv_stream(s, upper, !notran, !nounit, n, 1, A, lda, x, n);           // dlatrs.f:562-567 (DTRSV)
    } else {
        double* dcn = cnorm;
        double* tmpc = nullptr;
        if (cnorm_scaled) {                                                    // the kernel needs the SCALED norms
            tmpc = (double*)ws_alloc(s, sizeof(double) * (size_t)n);
            LB_CUDA_CHECK(cudaMemcpyAsync(tmpc, hc.data(), sizeof(double) * (size_t)n, cudaMemcpyHostToDevice, s));
            dcn = tmpc;
        }
        double* dscale = (double*)ws_alloc(s, 64);
        LatrsParams p;
        p.n = n; p.upper = upper; p.notran = notran; p.nounit = nounit; p.A = A; p.lda = lda; p.x = x; p.cnorm = dcn;
        p.tscal = tscal; p.smlnum = smlnum; p.bignum = bignum; p.scale_out = dscale;
        latrs_scaled_kernel<<<1, LT_THREADS, 0, s>>>(p);
        count_launch();
        LB_CUDA_CHECK(cudaMemcpyAsync(&scale, dscale, sizeof(double), cudaMemcpyDeviceToHost, s));
        LB_CUDA_CHECK(cudaStreamSynchronize(s));
        ws_free(s, dscale);
        if (tmpc) ws_free(s, tmpc);
    }
    LB_CUDA_CHECK(cudaGetLastError());
    return scale;
}

// ------------------------------------------------------------------------------------------------ DLACN2 (host, SRC/dlacn2.f:166-293)
static void host_lacn2(int n, double* v, double* x, int* isgn, double* est, int* kase, int* isave) {
    const int itmax = 5;
    auto iamax = [&](const double* y) { int j = 0; double m = fabs(y[0]); for (int i = 1; i < n; ++i) if (fabs(y[i]) > m) { m = fabs(y[i]); j = i; } return j + 1; };
    auto asum = [&](const double* y) { double t = 0.0; for (int i = 0; i < n; ++i) t += fabs(y[i]); return t; };
    auto sign_step = [&]() { for (int i = 0; i < n; ++i) { x[i] = (x[i] >= 0.0) ? 1.0 : -1.0; isgn[i] = (int)x[i]; } };
    auto unit_step = [&]() { for (int i = 0; i < n; ++i) x[i] = 0.0; x[isave[1] - 1] = 1.0; *kase = 1; isave[0] = 3; };
    auto alt_step = [&]() { double a = 1.0; for (int i = 0; i < n; ++i) { x[i] = a * (1.0 + (double)i / (double)(n - 1)); a = -a; } *kase = 1; isave[0] = 5; };
    if (*kase == 0) { for (int i = 0; i < n; ++i) x[i] = 1.0 / (double)n; *kase = 1; isave[0] = 1; return; }
    switch (isave[0]) {
    case 1:
        if (n == 1) { v[0] = x[0]; *est = fabs(v[0]); *kase = 0; return; }
        *est = asum(x);
        sign_step();
        *kase = 2; isave[0] = 2;
        return;
    case 2:
        isave[1] = iamax(x); isave[2] = 2;
        unit_step();
        return;
    case 3: {
        for (int i = 0; i < n; ++i) v[i] = x[i];
        const double estold = *est;
        *est = asum(v);
        bool changed = false;
        for (int i = 0; i < n; ++i) { const int xs = (x[i] >= 0.0) ? 1 : -1; if (xs != isgn[i]) { changed = true; break; } }
        if (!changed || *est <= estold) { alt_step(); return; }
        sign_step();
        *kase = 2; isave[0] = 4;
        return;
    }
    case 4: {
        const int jlast = isave[1];
        isave[1] = iamax(x);
        if (x[jlast - 1] != fabs(x[isave[1] - 1]) && isave[2] < itmax) { isave[2] += 1; unit_step(); return; }
        alt_step();
        return;
    }
    case 5: {
        const double temp = 2.0 * (asum(x) / (double)(3 * n));
        if (temp > *est) { for (int i = 0; i < n; ++i) v[i] = x[i]; *est = temp; }
        *kase = 0;
        return;
    }
    }
}

// SRC/drscl.f:120-170 on a host vector
static void host_rscl(int n, double sa, double* x) {
    const double smlnum = DBL_MIN, bignum = 1.0 / smlnum;
    double cden = sa, cnum = 1.0, mul;
    bool done;
    do {
        const double cden1 = cden * smlnum, cnum1 = cnum / bignum;
        if (fabs(cden1) > fabs(cnum) && cnum != 0.0) { mul = smlnum; done = false; cden = cden1; }
        else if (fabs(cnum1) > fabs(cden)) { mul = bignum; done = false; cnum = cnum1; }
        else { mul = cnum / cden; done = true; }
        for (int i = 0; i < n; ++i) x[i] *= mul;
    } while (!done);
}

// ------------------------------------------------------------------------------------------------ DGECON
// A: DGETRF factors on the device.  Returns INFO (0 / 1 as dgecon.f:272-282); argument checks live in the ABI layer.
int gecon(cudaStream_t s, bool onenrm, int n, const double* A, i64 lda, double anorm, double* rcond) {
    *rcond = 0.0;
    const double hugeval = DBL_MAX, smlnum = DBL_MIN;
    double* dx = (double*)ws_alloc(s, sizeof(double) * (size_t)n);
    double* cl = (double*)ws_alloc(s, sizeof(double) * (size_t)n);
    double* cu = (double*)ws_alloc(s, sizeof(double) * (size_t)n);
    std::vector<double> hx((size_t)n), hv((size_t)n), hdiag, nodiag;
    std::vector<int> isgn((size_t)n);
    double ainvnm = 0.0;
    bool normin = false;
    const int kase1 = onenrm ? 1 : 2;
    int kase = 0, isave[3] = {0, 0, 0};
    bool bailed = false;
    for (;;) {
        host_lacn2(n, hv.data(), hx.data(), isgn.data(), &ainvnm, &kase, isave);
        if (kase == 0) break;
        LB_CUDA_CHECK(cudaMemcpyAsync(dx, hx.data(), sizeof(double) * (size_t)n, cudaMemcpyHostToDevice, s));
        double sl, su;
        if (kase == kase1) {
            sl = latrs(s, false, true, false, normin, n, A, lda, dx, cl, nodiag);      // inv(L)
            su = latrs(s, true, true, true, normin, n, A, lda, dx, cu, hdiag);         // inv(U)
        } else {
            su = latrs(s, true, false, true, normin, n, A, lda, dx, cu, hdiag);        // inv(U**T)
            sl = latrs(s, false, false, false, normin, n, A, lda, dx, cl, nodiag);     // inv(L**T)
        }
        LB_CUDA_CHECK(cudaMemcpyAsync(hx.data(), dx, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost, s));
        LB_CUDA_CHECK(cudaStreamSynchronize(s));
        const double scale = sl * su;
        normin = true;
        if (scale != 1.0) {
            int ix = 0;
            for (int i = 1; i < n; ++i) if (fabs(hx[i]) > fabs(hx[ix])) ix = i;
            if (scale < fabs(hx[ix]) * smlnum || scale == 0.0) { bailed = true; break; }   // dgecon.f:260-261 (RCOND stays 0)
            host_rscl(n, scale, hx.data());
        }
    }
    ws_free(s, dx); ws_free(s, cl); ws_free(s, cu);
    if (bailed) return 0;
    if (ainvnm != 0.0) *rcond = (1.0 / ainvnm) / anorm;
    else return 1;
    if (*rcond != *rcond || *rcond > hugeval) return 1;
    return 0;
}

// ------------------------------------------------------------------------------------------------ norms / equilibration (device passes, host scalars)
// DLANGE('1' / 'I' / 'M') of a device matrix; synchronises
double lange(cudaStream_t s, char norm, int m, int n, const double* A, i64 lda) {
    if (m <= 0 || n <= 0) return 0.0;
    const bool one = (norm == '1' || norm == 'O' || norm == 'o'), inf = (norm == 'I' || norm == 'i');
    const int len = inf ? m : n;
    double* d = (double*)ws_alloc(s, sizeof(double) * (size_t)len);
    if (one) col_asum_kernel<<<ceil_div(n, 8), 256, 0, s>>>(m, n, A, lda, d);
    else if (inf) row_asum_kernel<<<ceil_div(m, 128), 128, 0, s>>>(m, n, A, lda, d);
    else col_amax_kernel<<<ceil_div(n, 8), 256, 0, s>>>(m, n, A, lda, nullptr, d);
    count_launch();
    std::vector<double> h = download(s, d, len);
    ws_free(s, d);
    return nanmax(h);
}
// DLANTR('M','U','N', m, n)
double lantr_max_upper(cudaStream_t s, int m, int n, const double* A, i64 lda) {
    if (m <= 0 || n <= 0) return 0.0;
    double* d = (double*)ws_alloc(s, sizeof(double) * (size_t)n);
    upper_col_amax_kernel<<<ceil_div(n, 8), 256, 0, s>>>(m, n, A, lda, d);
    count_launch();
    std::vector<double> h = download(s, d, n);
    ws_free(s, d);
    return nanmax(h);
}
// DGEEQU: r, c are HOST vectors (the scalar post-processing of dgeequ.f:206-236, :258-288 runs on them); returns INFO
int geequ(cudaStream_t s, int m, int n, const double* A, i64 lda, double* r, double* c, double* rowcnd, double* colcnd, double* amax) {
    if (m == 0 || n == 0) { *rowcnd = 1.0; *colcnd = 1.0; *amax = 0.0; return 0; }
    const double smlnum = DBL_MIN, bignum = 1.0 / smlnum;
    double* dr = (double*)ws_alloc(s, sizeof(double) * (size_t)m);
    double* dc = (double*)ws_alloc(s, sizeof(double) * (size_t)n);
    row_amax_kernel<<<ceil_div(m, 128), 128, 0, s>>>(m, n, A, lda, dr);
    count_launch();
    std::vector<double> hr = download(s, dr, m);
    double rcmin = bignum, rcmax = 0.0;
    for (int i = 0; i < m; ++i) { r[i] = hr[i]; rcmax = fmax(rcmax, r[i]); rcmin = fmin(rcmin, r[i]); }
    *amax = rcmax;
    int info = 0;
    if (rcmin == 0.0) { for (int i = 0; i < m; ++i) if (r[i] == 0.0) { info = i + 1; break; } }
    else {
        for (int i = 0; i < m; ++i) r[i] = 1.0 / fmin(fmax(r[i], smlnum), bignum);
        *rowcnd = fmax(rcmin, smlnum) / fmin(rcmax, bignum);
    }
    if (info == 0) {
        LB_CUDA_CHECK(cudaMemcpyAsync(dr, r, sizeof(double) * (size_t)m, cudaMemcpyHostToDevice, s));
        col_amax_kernel<<<ceil_div(n, 8), 256, 0, s>>>(m, n, A, lda, dr, dc);
        count_launch();
        std::vector<double> hc = download(s, dc, n);
        rcmin = bignum; rcmax = 0.0;
        for (int j = 0; j < n; ++j) { c[j] = hc[j]; rcmin = fmin(rcmin, c[j]); rcmax = fmax(rcmax, c[j]); }
        if (rcmin == 0.0) { for (int j = 0; j < n; ++j) if (c[j] == 0.0) { info = m + j + 1; break; } }
        else {
            for (int j = 0; j < n; ++j) c[j] = 1.0 / fmin(fmax(c[j], smlnum), bignum);
            *colcnd = fmax(rcmin, smlnum) / fmin(rcmax, bignum);
        }
    }
    ws_free(s, dr); ws_free(s, dc);
    return info;
}
// DLAQGE on a device matrix with HOST scale vectors; returns EQUED
char laqge(cudaStream_t s, int m, int n, double* A, i64 lda, const double* r, const double* c, double rowcnd, double colcnd, double amax) {
    const double thresh = 0.1;
    if (m <= 0 || n <= 0) return 'N';
    const double small_ = DBL_MIN / DBL_EPSILON, large_ = 1.0 / small_;
    int mode;
    char equed;
    if (rowcnd >= thresh && amax >= small_ && amax <= large_) {
        if (colcnd >= thresh) return 'N';
        mode = 1; equed = 'C';
    } else if (colcnd >= thresh) { mode = 2; equed = 'R'; }
    else { mode = 3; equed = 'B'; }
    double* dr = (double*)ws_alloc(s, sizeof(double) * (size_t)m);
    double* dc = (double*)ws_alloc(s, sizeof(double) * (size_t)n);
    LB_CUDA_CHECK(cudaMemcpyAsync(dr, r, sizeof(double) * (size_t)m, cudaMemcpyHostToDevice, s));
    LB_CUDA_CHECK(cudaMemcpyAsync(dc, c, sizeof(double) * (size_t)n, cudaMemcpyHostToDevice, s));
    dim3 grid(ceil_div(m, 256), (unsigned)min(n, 8192));
    laqge_kernel<<<grid, 256, 0, s>>>(m, n, A, lda, dr, dc, mode);
    count_launch();
    ws_free(s, dr); ws_free(s, dc);
    return equed;
}
// B := diag(d) B with d a HOST vector
void scale_rows(cudaStream_t s, int m, int n, double* B, i64 ldb, const double* d) {
    if (m <= 0 || n <= 0) return;
    double* dd = (double*)ws_alloc(s, sizeof(double) * (size_t)m);
    LB_CUDA_CHECK(cudaMemcpyAsync(dd, d, sizeof(double) * (size_t)m, cudaMemcpyHostToDevice, s));
    dim3 grid(ceil_div(m, 256), (unsigned)min(n, 8192));
    row_scale_kernel<<<grid, 256, 0, s>>>(m, n, B, ldb, dd);
    count_launch();
    ws_free(s, dd);
}

}  // namespace lb
