// fortran_abi.cu -- the drop-in boundary: Fortran-77 ABI symbols of the reference's LU / Cholesky / QR
// hot path and the Level-3 BLAS under it (declared in include/lapack_b200_f77.h).
//
// Contract (reference LAPACKE/include/lapack.h): every argument by reference,
// INTEGER = 32-bit int, CHARACTER*1 = char* plus a hidden trailing size_t length (only the first character
// is read, case-insensitively like BLAS/SRC/lsame.f), column-major, in place, synchronous on return.
// Argument checks, their order, the XERBLA name/position and quick returns are the reference's
// (SRC/dgetrf.f:144-160, dgetrs.f:158-180, dgesv.f:148-162, dpotrf.f:145-162, dpotrs.f:147-166,
// dposv.f:160-176, dgeqrf.f:180-212, dgeqr2.f:150-162, BLAS/SRC/dgemm.f:253-299, dtrsm.f:230-260,
// dsyrk.f:210-240).  Errors go through the EXTERNAL symbol xerbla_ so that a caller's own XERBLA
// (e.g. TESTING/LIN/xerbla.f) intercepts them.
//
// Pointers may be host or device memory (cudaPointerGetAttributes).  Host operands are staged through
// device scratch on an internal stream (pinned host memory gets full-rate async DMA); device operands are
// used in place on the legacy default stream.  There is no CPU fallback: without a usable CUDA device the
// routines report INFO = -1001 - cudaError and return.
#include "lb_internal.h"
#include "../../include/lapack_b200_f77.h"
#include <cfloat>
#include <cmath>
#include <cstring>
#include <mutex>
#include <vector>

namespace lb {
void laswp_rows(cudaStream_t s, int n, int rows, double* A, i64 lda, int k1, int k2, const int* ipiv, int incx);
// gecon.cu
double latrs(cudaStream_t s, bool upper, bool notran, bool nounit, bool normin_y, int n, const double* A, i64 lda, double* x,
             double* cnorm, std::vector<double>& hdiag);
int gecon(cudaStream_t s, bool onenrm, int n, const double* A, i64 lda, double anorm, double* rcond);
double lange(cudaStream_t s, char norm, int m, int n, const double* A, i64 lda);
double lantr_max_upper(cudaStream_t s, int m, int n, const double* A, i64 lda);
int geequ(cudaStream_t s, int m, int n, const double* A, i64 lda, double* r, double* c, double* rowcnd, double* colcnd, double* amax);
char laqge(cudaStream_t s, int m, int n, double* A, i64 lda, const double* r, const double* c, double rowcnd, double colcnd, double amax);
void scale_rows(cudaStream_t s, int m, int n, double* B, i64 ldb, const double* d);
}

// ------------------------------------------------------------------------------------------------ XERBLA
static int g_xerbla_mode = -1;        // 0 stop (reference), 1 print+return, 2 silent (record only)
static char g_xerbla_name[33] = {0};
static int g_xerbla_info = 0;
static int g_xerbla_count = 0;

extern "C" void lb200_set_xerbla_mode(int mode) { g_xerbla_mode = mode; }
extern "C" int lb200_last_xerbla(char* name_out, int* info_out) {
    if (name_out) strcpy(name_out, g_xerbla_name);
    if (info_out) *info_out = g_xerbla_info;
    return g_xerbla_count;
}
extern "C" void lb200_clear_xerbla(void) { g_xerbla_count = 0; g_xerbla_info = 0; g_xerbla_name[0] = 0; }

// BLAS/SRC/xerbla.f:59-83.  Weak: a XERBLA supplied by the application or test harness wins.
extern "C" __attribute__((weak)) void xerbla_(const char* srname, const int* info, size_t srname_len) {
    size_t len = srname_len;
    if (len == 0 || len > 32) len = strnlen(srname, 32);
    memcpy(g_xerbla_name, srname, len);
    g_xerbla_name[len] = 0;
    while (len > 0 && g_xerbla_name[len - 1] == ' ') g_xerbla_name[--len] = 0;   // LEN_TRIM
    g_xerbla_info = *info;
    g_xerbla_count++;
    if (g_xerbla_mode < 0) {
        const char* e = getenv("LAPACK_B200_XERBLA");
        g_xerbla_mode = (e && !strcmp(e, "return")) ? 1 : (e && !strcmp(e, "silent")) ? 2 : 0;
    }
    if (g_xerbla_mode != 2)
        printf(" ** On entry to %s parameter number %2d had an illegal value\n", g_xerbla_name, *info);
    if (g_xerbla_mode == 0) {
        fflush(stdout);
        exit(0);   // Fortran STOP
    }
}

static void call_xerbla(const char* name6, int pos) { xerbla_(name6, &pos, strlen(name6)); }

// BLAS/SRC/lsame.f
extern "C" int lsame_(const char* ca, const char* cb, size_t, size_t) {
    char a = *ca, b = *cb;
    if (a >= 'a' && a <= 'z') a = (char)(a - 32);
    if (b >= 'a' && b <= 'z') b = (char)(b - 32);
    return a == b;
}
static inline bool same(const char* p, char u) { char c = *p; if (c >= 'a' && c <= 'z') c = (char)(c - 32); return c == u; }
static inline int imax(int a, int b) { return a > b ? a : b; }
static inline int imin(int a, int b) { return a < b ? a : b; }

// ------------------------------------------------------------------------------------------------ staging
namespace {

enum PtrKind { PK_HOST = 0, PK_PINNED = 1, PK_DEVICE = 2 };

PtrKind ptr_kind(const void* p) {
    cudaPointerAttributes at;
    if (!p) return PK_HOST;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { (void)cudaGetLastError(); return PK_HOST; }
    if (at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged) return PK_DEVICE;
    if (at.type == cudaMemoryTypeHost) return PK_PINNED;
    return PK_HOST;
}

std::mutex g_abi_mutex;     // one Fortran-ABI call at a time (the reference is re-entrant; we serialise)

// A large PAGEABLE host matrix (what a Fortran caller's ALLOCATE or malloc gives) is page-locked for the duration of the
// call, so that the streamed paths below (asynchronous DMA overlapped with the factorization) apply to it too; without
// this every cudaMemcpyAsync from it degrades to a synchronous staged copy.  LAPACK_B200_HOST_REGISTER=0 disables it.
struct TempPin {
    void* base = nullptr;
    // ONE registration for the whole matrix.  Measured (tools/hostreg_probe.py, 8 GiB, pages already touched): cudaHostRegister
    // 179 ms + cudaHostUnregister 140 ms on one thread; splitting the range over 2 / 4 / 8 / 16 threads is SLOWER (532 / 225 / 851 /
    // 1951 ms) and a 2-D copy that spans two registrations is rejected (cudaErrorInvalidValue), so the range is never split.
    TempPin(const void* p, size_t bytes, size_t min_bytes = (size_t)256 << 20) {
        static int enabled = -1;
        if (enabled < 0) { const char* e = getenv("LAPACK_B200_HOST_REGISTER"); enabled = (e && e[0] == '0') ? 0 : 1; }
        if (!enabled || !p || bytes < min_bytes || ptr_kind(p) != PK_HOST) return;
        if (cudaHostRegister(const_cast<void*>(p), bytes, cudaHostRegisterDefault) == cudaSuccess) base = const_cast<void*>(p);
        else (void)cudaGetLastError();
    }
    ~TempPin() { if (base && cudaHostUnregister(base) != cudaSuccess) (void)cudaGetLastError(); }
    TempPin(const TempPin&) = delete;
    TempPin& operator=(const TempPin&) = delete;
};

cudaStream_t host_stream() {
    static cudaStream_t s = nullptr;
    if (!s) LB_CUDA_CHECK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    return s;
}

bool device_ok(int* info) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0) {
        (void)cudaGetLastError();
        fprintf(stderr, "lapack_b200: no usable CUDA device (%s); this library has no CPU fallback\n",
                e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
        if (info) *info = -1001 - (int)e;
        return false;
    }
    return true;
}

// Call context: picks the stream (legacy default stream if any operand is device memory, else the internal
// stream), owns the staged copies, copies results back and synchronises in finish().
struct Ctx {
    cudaStream_t s = nullptr;
    bool any_device = false;
    struct Item { void* dev; void* host; size_t width_bytes; size_t rows_or_cols; size_t hpitch; size_t dpitch; bool out; };
    std::vector<Item> items;

    void scan(std::initializer_list<const void*> ptrs) {
        for (const void* p : ptrs) if (p && ptr_kind(p) == PK_DEVICE) any_device = true;
        s = any_device ? (cudaStream_t)0 : host_stream();
    }
    // column-major rows x cols matrix with leading dimension ld
    double* mat(const double* p, int rows, int cols, lb::i64 ld, bool in, bool out, lb::i64* ld_dev) {
        if (rows <= 0 || cols <= 0) { *ld_dev = imax(1, rows); return const_cast<double*>(p); }
        if (ptr_kind(p) == PK_DEVICE) { *ld_dev = ld; return const_cast<double*>(p); }
        lb::i64 ldd = ((lb::i64)rows + 1) & ~1LL;
        double* d = (double*)lb::ws_alloc(s, sizeof(double) * (size_t)ldd * cols);
        if (in) LB_CUDA_CHECK(cudaMemcpy2DAsync(d, ldd * 8, p, ld * 8, (size_t)rows * 8, cols, cudaMemcpyHostToDevice, s));
        items.push_back({d, (void*)p, (size_t)rows * 8, (size_t)cols, (size_t)ld * 8, (size_t)ldd * 8, out});
        *ld_dev = ldd;
        return d;
    }
    template <typename T>
    T* vec(const T* p, size_t n, bool in, bool out) {
        if (n == 0) return const_cast<T*>(p);
        if (ptr_kind(p) == PK_DEVICE) return const_cast<T*>(p);
        T* d = (T*)lb::ws_alloc(s, sizeof(T) * n);
        if (in) LB_CUDA_CHECK(cudaMemcpyAsync(d, p, sizeof(T) * n, cudaMemcpyHostToDevice, s));
        items.push_back({d, (void*)p, sizeof(T) * n, 1, sizeof(T) * n, sizeof(T) * n, out});
        return d;
    }
    int* dev_info() {
        int* d = (int*)lb::ws_alloc(s, 64);
        LB_CUDA_CHECK(cudaMemsetAsync(d, 0, 64, s));
        return d;
    }
    // copies outputs back, frees scratch, synchronises; returns the device INFO word if given
    int finish(int* dinfo = nullptr) {
        int hinfo = 0;
        for (auto& it : items)
            if (it.out)
                LB_CUDA_CHECK(cudaMemcpy2DAsync(it.host, it.hpitch, it.dev, it.dpitch, it.width_bytes, it.rows_or_cols,
                                                cudaMemcpyDeviceToHost, s));
        if (dinfo) LB_CUDA_CHECK(cudaMemcpyAsync(&hinfo, dinfo, sizeof(int), cudaMemcpyDeviceToHost, s));
        for (auto& it : items) lb::ws_free(s, it.dev);
        if (dinfo) lb::ws_free(s, dinfo);
        LB_CUDA_CHECK(cudaStreamSynchronize(s));
        int e = lb::take_cuda_error();
        if (e) return -1001 - e;
        return hinfo;
    }
};

}  // namespace

// ---- device helpers of DGERFS ------------------------------------------------------------------------
namespace {
// w(i) = |b(i)| + sum_k |A(i,k)| |x(k)|        (dgerfs.f:296-305), one thread per row, coalesced over rows
__global__ void abs_gemv_n_kernel(int n, const double* __restrict__ A, lb::i64 lda, const double* __restrict__ x,
                                  const double* __restrict__ b, double* __restrict__ w) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double acc = fabs(b[i]);
    for (int k = 0; k < n; ++k) acc += fabs(A[i + (lb::i64)k * lda]) * fabs(x[k]);
    w[i] = acc;
}
// w(k) = |b(k)| + sum_i |A(i,k)| |x(i)|        (dgerfs.f:306-314), one warp per column
__global__ void abs_gemv_t_kernel(int n, const double* __restrict__ A, lb::i64 lda, const double* __restrict__ x,
                                  const double* __restrict__ b, double* __restrict__ w) {
    const int k = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (k >= n) return;
    double acc = 0.0;
    for (int i = lane; i < n; i += 32) acc += fabs(A[i + (lb::i64)k * lda]) * fabs(x[i]);
    for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
    if (lane == 0) w[k] = fabs(b[k]) + acc;
}
__global__ void vec_add_kernel(int n, const double* __restrict__ d, double* __restrict__ x) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) x[i] += d[i];
}

// DLACN2 (SRC/dlacn2.f:166-293) on host vectors; same state machine as the reference (isave[3], labels 1..5)
void host_dlacn2(int n, double* v, double* x, int* isgn, double* est, int* kase, int* isave) {
    const int itmax = 5;
    auto iamax = [&](const double* y) { int j = 0; double m = fabs(y[0]); for (int i = 1; i < n; ++i) if (fabs(y[i]) > m) { m = fabs(y[i]); j = i; } return j + 1; };
    auto asum = [&](const double* y) { double t = 0.0; for (int i = 0; i < n; ++i) t += fabs(y[i]); return t; };
    if (*kase == 0) {
        for (int i = 0; i < n; ++i) x[i] = 1.0 / (double)n;
        *kase = 1; isave[0] = 1;
        return;
    }
    int state = isave[0];
    if (state == 1) {
        if (n == 1) { v[0] = x[0]; *est = fabs(v[0]); *kase = 0; return; }
        *est = asum(x);
        for (int i = 0; i < n; ++i) { x[i] = (x[i] >= 0.0) ? 1.0 : -1.0; isgn[i] = (int)x[i]; }
        *kase = 2; isave[0] = 2;
        return;
    }
    bool to50 = false, to120 = false;
    if (state == 2) { isave[1] = iamax(x); isave[2] = 2; to50 = true; }
    else if (state == 3) {
        memcpy(v, x, sizeof(double) * (size_t)n);
        const double estold = *est;
        *est = asum(v);
        bool changed = false;
        for (int i = 0; i < n; ++i) { const int xs = (x[i] >= 0.0) ? 1 : -1; if (xs != isgn[i]) { changed = true; break; } }
        if (!changed || *est <= estold) to120 = true;
        else {
            for (int i = 0; i < n; ++i) { x[i] = (x[i] >= 0.0) ? 1.0 : -1.0; isgn[i] = (int)x[i]; }
            *kase = 2; isave[0] = 4;
            return;
        }
    } else if (state == 4) {
        const int jlast = isave[1];
        isave[1] = iamax(x);
        if (x[jlast - 1] != fabs(x[isave[1] - 1]) && isave[2] < itmax) { isave[2] += 1; to50 = true; }
        else to120 = true;
    } else if (state == 5) {
        const double temp = 2.0 * (asum(x) / (double)(3 * n));
        if (temp > *est) { memcpy(v, x, sizeof(double) * (size_t)n); *est = temp; }
        *kase = 0;
        return;
    }
    if (to50) {
        for (int i = 0; i < n; ++i) x[i] = 0.0;
        x[isave[1] - 1] = 1.0;
        *kase = 1; isave[0] = 3;
        return;
    }
    if (to120) {
        double altsgn = 1.0;
        for (int i = 0; i < n; ++i) { x[i] = altsgn * (1.0 + (double)i / (double)(n - 1)); altsgn = -altsgn; }
        *kase = 1; isave[0] = 5;
        return;
    }
}
}  // namespace

extern "C" {

// ================================================================================================ BLAS 3
void dgemm_(const char* transa, const char* transb, const int* m, const int* n, const int* k, const double* alpha,
            const double* A, const int* lda, const double* B, const int* ldb, const double* beta, double* C,
            const int* ldc, size_t, size_t) {
    const bool nota = same(transa, 'N'), notb = same(transb, 'N');
    const int nrowa = nota ? *m : *k, nrowb = notb ? *k : *n;
    int info = 0;
    if (!nota && !same(transa, 'C') && !same(transa, 'T')) info = 1;
    else if (!notb && !same(transb, 'C') && !same(transb, 'T')) info = 2;
    else if (*m < 0) info = 3;
    else if (*n < 0) info = 4;
    else if (*k < 0) info = 5;
    else if (*lda < imax(1, nrowa)) info = 8;
    else if (*ldb < imax(1, nrowb)) info = 10;
    else if (*ldc < imax(1, *m)) info = 13;
    if (info) { call_xerbla("DGEMM ", info); return; }
    if (*m == 0 || *n == 0 || ((*alpha == 0.0 || *k == 0) && *beta == 1.0)) return;
    if (!device_ok(nullptr)) return;
    std::lock_guard<std::mutex> lock(g_abi_mutex);
    Ctx c; c.scan({A, B, C});
    lb::i64 la, lbb, lc;
    const bool need_ab = (*alpha != 0.0 && *k > 0);
    const int acols = nota ? *k : *m, bcols = notb ? *n : *k;
    const double* dA = need_ab ? c.mat(A, nrowa, acols, *lda, true, false, &la) : A;
    const double* dB = need_ab ? c.mat(B, nrowb, bcols, *ldb, true, false, &lbb) : B;
    if (!need_ab) { la = *lda; lbb = *ldb; }
    double* dC = c.mat(C, *m, *n, *ldc, *beta != 0.0, true, &lc);
    lb::gemm(c.s, nota ? 'N' : 'T', notb ? 'N' : 'T', *m, *n, need_ab ? *k : 0, *alpha, dA, la, dB, lbb, *beta, dC, lc, 0);
    c.finish();
}

void dsyrk_(const char* uplo, const char* trans, const int* n, const int* k, const double* alpha, const double* A,
            const int* lda, const double* beta, double* C, const int* ldc, size_t, size_t) {
    const bool notr = same(trans, 'N');
    const int nrowa = notr ? *n : *k;
    const bool upper = same(uplo, 'U');
    int info = 0;
    if (!upper && !same(uplo, 'L')) info = 1;
    else if (!notr && !same(trans, 'T') && !same(trans, 'C')) info = 2;
    else if (*n < 0) info = 3;
    else if (*k < 0) info = 4;
    else if (*lda < imax(1, nrowa)) info = 7;
    else if (*ldc < imax(1, *n)) info = 10;
    if (info) { call_xerbla("DSYRK ", info); return; }
    if (*n == 0 || ((*alpha == 0.0 || *k == 0) && *beta == 1.0)) return;
    if (!device_ok(nullptr)) return;
    std::lock_guard<std::mutex> lock(g_abi_mutex);
    Ctx c; c.scan({A, C});
    lb::i64 la = *lda, lc;
    const bool need_a = (*alpha != 0.0 && *k > 0);
    const double* dA = need_a ? c.mat(A, nrowa, notr ? *k : *n, *lda, true, false, &la) : A;
    double* dC = c.mat(C, *n, *n, *ldc, true, true, &lc);   // the untouched triangle travels both ways unchanged
    lb::syrk(c.s, upper ? 'U' : 'L', notr ? 'N' : 'T', *n, need_a ? *k : 0, *alpha, dA, la, *beta, dC, lc);
    c.finish();
}

static void trxm_common(bool solve, const char* side, const char* uplo, const char* transa, const char* diag, const int* m,
                        const int* n, const double* alpha, const double* A, const int* lda, double* B, const int* ldb) {
    const bool lside = same(side, 'L');
    const int nrowa = lside ? *m : *n;
    int info = 0;
    if (!lside && !same(side, 'R')) info = 1;
    else if (!same(uplo, 'U') && !same(uplo, 'L')) info = 2;
    else if (!same(transa, 'N') && !same(transa, 'T') && !same(transa, 'C')) info = 3;
    else if (!same(diag, 'U') && !same(diag, 'N')) info = 4;
    else if (*m < 0) info = 5;
    else if (*n < 0) info = 6;
    else if (*lda < imax(1, nrowa)) info = 9;
    else if (*ldb < imax(1, *m)) info = 11;
    if (info) { call_xerbla(solve ? "DTRSM " : "DTRMM ", info); return; }
    if (*m == 0 || *n == 0) return;
    if (!device_ok(nullptr)) return;
    std::lock_guard<std::mutex> lock(g_abi_mutex);
    Ctx c; c.scan({A, B});
    lb::i64 la = *lda, lbb;
    const double* dA = (*alpha != 0.0) ? c.mat(A, nrowa, nrowa, *lda, true, false, &la) : A;
    double* dB = c.mat(B, *m, *n, *ldb, *alpha != 0.0, true, &lbb);
    const char tr = same(transa, 'N') ? 'N' : 'T';
    if (solve) lb::trsm(c.s, lside ? 'L' : 'R', same(uplo, 'U') ? 'U' : 'L', tr, same(diag, 'U') ? 'U' : 'N', *m, *n, *alpha, dA, la, dB, lbb);
    else lb::trmm(c.s, lside ? 'L' : 'R', same(uplo, 'U') ? 'U' : 'L', tr, same(diag, 'U') ? 'U' : 'N', *m, *n, *alpha, dA, la, dB, lbb);
    c.finish();
}
void dtrsm_(const char* side, const char* uplo, const char* transa, const char* diag, const int* m, const int* n,
            const double* alpha, const double* A, const int* lda, double* B, const int* ldb, size_t, size_t, size_t, size_t) {
    trxm_common(true, side, uplo, transa, diag, m, n, alpha, A, lda, B, ldb);
}
void dtrmm_(const char* side, const char* uplo, const char* transa, const char* diag, const int* m, const int* n,
            const double* alpha, const double* A, const int* lda, double* B, const int* ldb, size_t, size_t, size_t, size_t) {
    trxm_common(false, side, uplo, transa, diag, m, n, alpha, A, lda, B, ldb);
}

// ================================================================================================ LU
// Host-resident square DGETRF with transfer/compute overlap: one level of the DGETRF2 recursion (dgetrf2.f:216-263)
// at the top.  The left n1 columns are uploaded first and factored (blocked DGETRF) while the right n - n1 columns
// are still crossing PCIe; then A12/A22 get the accumulated interchanges, the triangular solve and ONE large-K GEMM,
// A22 is factored, and its interchanges go back to the left columns.  Everything that is final early (L11/U11, U12,
// the block rows of U inside A22) is downloaded while the factorization continues.
static int getrf_host_streamed(int n, double* A, int lda, int* ipiv) {
    static cudaStream_t copy_stream = nullptr, up_stream = nullptr;
    static cudaEvent_t ev_up = nullptr, ev_u12 = nullptr;
    if (!copy_stream) {
        LB_CUDA_CHECK(cudaStreamCreateWithFlags(&copy_stream, cudaStreamNonBlocking));
        LB_CUDA_CHECK(cudaStreamCreateWithFlags(&up_stream, cudaStreamNonBlocking));
        LB_CUDA_CHECK(cudaEventCreateWithFlags(&ev_up, cudaEventDisableTiming));
        LB_CUDA_CHECK(cudaEventCreateWithFlags(&ev_u12, cudaEventDisableTiming));
    }
    cudaStream_t s = host_stream();
    const lb::i64 ldd = ((lb::i64)n + 1) & ~1LL;
    double* dA = (double*)lb::ws_alloc(s, sizeof(double) * (size_t)ldd * n);
    int* dp = (int*)lb::ws_alloc(s, sizeof(int) * (size_t)n);
    int* dinfo = (int*)lb::ws_alloc(s, 64);          // dinfo[0] = left part / result, dinfo[8] = A22
    // split so that factoring the left part takes about as long as uploading the right part (measured: 3n/8)
    const int n1 = imin(n - 512, (int)((((lb::i64)n * 3 / 8) + 511) / 512) * 512), n2 = n - n1;
    double* dA12 = dA + (lb::i64)n1 * ldd;
    double* dA22 = dA12 + n1;
    // make sure the scratch exists before the copy stream touches it
    LB_CUDA_CHECK(cudaEventRecord(ev_up, s));
    LB_CUDA_CHECK(cudaStreamWaitEvent(up_stream, ev_up, 0));
    LB_CUDA_CHECK(cudaStreamWaitEvent(copy_stream, ev_up, 0));
    LB_CUDA_CHECK(cudaMemcpy2DAsync(dA, ldd * 8, A, (size_t)lda * 8, (size_t)n * 8, n1, cudaMemcpyHostToDevice, s));
    LB_CUDA_CHECK(cudaMemcpy2DAsync(dA12, ldd * 8, A + (lb::i64)n1 * lda, (size_t)lda * 8, (size_t)n * 8, n2,
                                    cudaMemcpyHostToDevice, up_stream));
    LB_CUDA_CHECK(cudaEventRecord(ev_up, up_stream));
    lb::getrf(s, n, n1, dA, ldd, dp, dinfo);                                            // dgetrf2.f:231
    // rows 0..n1 of the left part (L11 and U11) are final: later interchanges only touch rows >= n1
    LB_CUDA_CHECK(cudaEventRecord(ev_u12, s));
    LB_CUDA_CHECK(cudaStreamWaitEvent(copy_stream, ev_u12, 0));
    LB_CUDA_CHECK(cudaMemcpy2DAsync(A, (size_t)lda * 8, dA, ldd * 8, (size_t)n1 * 8, n1, cudaMemcpyDeviceToHost, copy_stream));
    LB_CUDA_CHECK(cudaStreamWaitEvent(s, ev_up, 0));
    lb::laswp_rows(s, n2, n, dA12, ldd, 1, n1, dp, 1);                                  // dgetrf2.f:236 (composed, one streaming pass)
    lb::trsm(s, 'L', 'L', 'N', 'U', n1, n2, 1.0, dA, ldd, dA12, ldd);                   // dgetrf2.f:240
    LB_CUDA_CHECK(cudaEventRecord(ev_u12, s));
    LB_CUDA_CHECK(cudaStreamWaitEvent(copy_stream, ev_u12, 0));
    LB_CUDA_CHECK(cudaMemcpy2DAsync(A + (lb::i64)n1 * lda, (size_t)lda * 8, dA12, ldd * 8, (size_t)n1 * 8, n2,
                                    cudaMemcpyDeviceToHost, copy_stream));              // U12 is final
    lb::gemm(s, 'N', 'N', n2, n2, n1, -1.0, dA + n1, ldd, dA12, ldd, 1.0, dA22, ldd);   // dgetrf2.f:245
    // A22: the blocked driver streams every finished block row of U itself (lb::StreamOut)
    lb::StreamOut so;
    so.host = A + n1 + (lb::i64)n1 * lda; so.ldh = lda; so.copy_stream = copy_stream; so.ev = ev_u12; so.done_cols = 0;
    lb::stream_out() = &so;
    lb::getrf(s, n2, n2, dA22, ldd, dp + n1, dinfo + 8);                                // dgetrf2.f:250
    lb::stream_out() = nullptr;
    lb::info_max_offset(s, dinfo, dinfo + 8, n1);                                       // dgetrf2.f:251-252
    lb::iadd(s, n2, dp + n1, n1);                                                       // dgetrf2.f:257-259
    lb::laswp_rows(s, n1, n, dA, ldd, n1 + 1, n, dp, 1);                                // dgetrf2.f:263 (composed, one streaming pass)
    // what is left of A22: the block lower trapezoids (L and the diagonal blocks), or everything if nothing was streamed
    {
        const int nb = lb::getrf_block();
        if (so.done_cols == 0) {
            LB_CUDA_CHECK(cudaMemcpy2DAsync(A + n1 + (lb::i64)n1 * lda, (size_t)lda * 8, dA22, ldd * 8, (size_t)n2 * 8, n2,
                                            cudaMemcpyDeviceToHost, s));
        } else {
            for (int j0 = 0; j0 < n2; j0 += nb) {
                const int w = imin(nb, n2 - j0);
                LB_CUDA_CHECK(cudaMemcpy2DAsync(A + (n1 + j0) + (lb::i64)(n1 + j0) * lda, (size_t)lda * 8,
                                                dA22 + j0 + (lb::i64)j0 * ldd, ldd * 8, (size_t)(n2 - j0) * 8, w,
                                                cudaMemcpyDeviceToHost, s));
            }
        }
    }
    // rows n1..n of the left part (L21 after the interchanges of A22)
    LB_CUDA_CHECK(cudaMemcpy2DAsync(A + n1, (size_t)lda * 8, dA + n1, ldd * 8, (size_t)n2 * 8, n1, cudaMemcpyDeviceToHost, s));
    LB_CUDA_CHECK(cudaMemcpyAsync(ipiv, dp, sizeof(int) * (size_t)n, cudaMemcpyDeviceToHost, s));
    int hinfo = 0;
    LB_CUDA_CHECK(cudaMemcpyAsync(&hinfo, dinfo, sizeof(int), cudaMemcpyDeviceToHost, s));
    LB_CUDA_CHECK(cudaStreamSynchronize(s));
    LB_CUDA_CHECK(cudaStreamSynchronize(copy_stream));
    lb::ws_free(s, dA);
    lb::ws_free(s, dp);
    lb::ws_free(s, dinfo);
    LB_CUDA_CHECK(cudaStreamSynchronize(s));
    int e = lb::take_cuda_error();
    if (e) return -1001 - e;
    return hinfo;
}

static void getrf_common(bool recursive, const int* m, const int* n, double* A, const int* lda, int* ipiv, int* info) {
    *info = 0;
    if (*m < 0) *info = -1;
    else if (*n < 0) *info = -2;
    else if (*lda < imax(1, *m)) *info = -4;
    if (*info != 0) { call_xerbla(recursive ? "DGETRF2" : "DGETRF", -*info); return; }
    if (*m == 0 || *n == 0) return;
    if (!device_ok(info)) return;
    std::lock_guard<std::mutex> lock(g_abi_mutex);
    TempPin pin(A, ((size_t)*lda * (size_t)(*n - 1) + (size_t)*m) * sizeof(double));
    if (!recursive && *m == *n && *n >= 8192 && ptr_kind(A) == PK_PINNED && ptr_kind(ipiv) != PK_DEVICE) {
        *info = getrf_host_streamed(*n, A, *lda, ipiv);
        return;
    }
    Ctx c; c.scan({A, ipiv});
    lb::i64 la;
    double* dA = c.mat(A, *m, *n, *lda, true, true, &la);
    int* dp = c.vec<int>(ipiv, (size_t)imin(*m, *n), false, true);
    int* dinfo = c.dev_info();
    if (recursive) lb::getrf2(c.s, *m, *n, dA, la, dp, dinfo);
    else lb::getrf(c.s, *m, *n, dA, la, dp, dinfo);
    *info = c.finish(dinfo);
}
void dgetrf_(const int* m, const int* n, double* A, const int* lda, int* ipiv, int* info) {
    getrf_common(false, m, n, A, lda, ipiv, info);
}
void dgetrf2_(const int* m, const int* n, double* A, const int* lda, int* ipiv, int* info) {
    getrf_common(true, m, n, A, lda, ipiv, info);
}

// SRC/dlaswp.f:112 -- no argument checking, no INFO
void dlaswp_(const int* n, double* A, const int* lda, const int* k1, const int* k2, const int* ipiv, const int* incx) {
    if (*n <= 0 || *incx == 0 || *k2 < *k1) return;
    if (!device_ok(nullptr)) return;
    std::lock_guard<std::mutex> lock(g_abi_mutex);
    // rows touched: K1..K2 and every pivot target; pivots are read at positions K1..K1+(K2-K1)*|INCX|
    const int ainc = *incx > 0 ? *incx : -*incx;
    const size_t npiv = (size_t)(*k1 - 1) + (size_t)(*k2 - *k1) * ainc + 1;
    Ctx c; c.scan({A, ipiv});
    int rows = *k2;
    std::vector<int> hp;
    if (ptr_kind(ipiv) != PK_DEVICE) {
        for (int i = *k1; i <= *k2; ++i) { int v = ipiv[(*k1 - 1) + (size_t)(i - *k1) * ainc]; if (v > rows) rows = v; }
    } else {
        hp.resize(npiv);
        cudaMemcpy(hp.data(), ipiv, npiv * sizeof(int), cudaMemcpyDeviceToHost);
        for (int i = *k1; i <= *k2; ++i) { int v = hp[(*k1 - 1) + (size_t)(i - *k1) * ainc]; if (v > rows) rows = v; }
    }
    if (rows > *lda) rows = *lda;
    lb::i64 la;
    double* dA = c.mat(A, rows, *n, *lda, true, true, &la);
    int* dp = c.vec<int>(ipiv, npiv, true, false);
    lb::laswp(c.s, *n, dA, la, *k1, *k2, dp, *incx);
    c.finish();
}

void dgetrs_(const char* trans, const int* n, const int* nrhs, const double* A, const int* lda, const int* ipiv, double* B,
             const int* ldb, int* info, size_t) {
    *info = 0;
    const bool notran = same(trans, 'N');
    if (!notran && !same(trans, 'T') && !same(trans, 'C')) *info = -1;
    else if (*n < 0) *info = -2;
    else if (*nrhs < 0) *info = -3;
    else if (*lda < imax(1, *n)) *info = -5;
    else if (*ldb < imax(1, *n)) *info = -8;
    if (*info != 0) { call_xerbla("DGETRS", -*info); return; }
    if (*n == 0 || *nrhs == 0) return;
    if (!device_ok(info)) return;
    std::lock_guard<std::mutex> lock(g_abi_mutex);
    Ctx c; c.scan({A, ipiv, B});
    lb::i64 la, lbb;
    const double* dA = c.mat(A, *n, *n, *lda, true, false, &la);
    const int* dp = c.vec<int>(ipiv, (size_t)*n, true, false);
    double* dB = c.mat(B, *n, *nrhs, *ldb, true, true, &lbb);
    lb::getrs(c.s, notran ? 'N' : 'T', *n, *nrhs, dA, la, dp, dB, lbb);
    int r = c.finish();
    if (r) *info = r;
}

void dgesv_(const int* n, const int* nrhs, double* A, const int* lda, int* ipiv, double* B, const int* ldb, int* info) {
    *info = 0;
    if (*n < 0) *info = -1;
    else if (*nrhs < 0) *info = -2;
    else if (*lda < imax(1, *n)) *info = -4;
    else if (*ldb < imax(1, *n)) *info = -7;
    if (*info != 0) { call_xerbla("DGESV ", -*info); return; }
    if (*n == 0) return;
    if (!device_ok(info)) return;
    std::lock_guard<std::mutex> lock(g_abi_mutex);
    Ctx c; c.scan({A, ipiv, B});
    lb::i64 la, lbb;
    double* dA = c.mat(A, *n, *n, *lda, true, true, &la);
    int* dp = c.vec<int>(ipiv, (size_t)*n, false, true);
    double* dB = c.mat(B, *n, *nrhs, *ldb, true, true, &lbb);
    int* dinfo = c.dev_info();
    lb::getrf(c.s, *n, *n, dA, la, dp, dinfo);
    // DGESV solves only if INFO == 0 (dgesv.f:166); the factorization's INFO is needed on the host first
    int hinfo = 0;
    LB_CUDA_CHECK(cudaMemcpyAsync(&hinfo, dinfo, sizeof(int), cudaMemcpyDeviceToHost, c.s));
    LB_CUDA_CHECK(cudaStreamSynchronize(c.s));
    if (hinfo == 0 && *nrhs > 0) lb::getrs(c.s, 'N', *n, *nrhs, dA, la, dp, dB, lbb);
    else if (hinfo != 0) { for (auto& it : c.items) if (it.dev == dB) it.out = false; }
    int r = c.finish(dinfo);
    *info = r;
}

// ================================================================================================ Cholesky
// Host-resident Cholesky with transfer/compute overlap: only the UPLO triangle crosses PCIe (block-column
// trapezoids), and every finished block column of the factor is downloaded on a copy stream while the rest of
// the factorization is still running (lb::StreamOut).  Needs pinned host memory for truly asynchronous copies.
// upload the UPLO triangle of columns [c0, c1) of a host matrix as block-column trapezoids; rows are clipped to
// [rlo, rhi) (used to send the top block row / the trailing block separately)
static void upload_triangle(cudaStream_t s, bool upper, int n, const double* A, int lda, double* dA, lb::i64 ldd, int cb,
                            int c0, int c1, int rlo, int rhi) {
    for (int j0 = c0; j0 < c1; j0 += cb) {
        const int w = imin(cb, c1 - j0);
        int r0 = upper ? 0 : j0, r1 = upper ? j0 + w : n;
        if (r0 < rlo) r0 = rlo;
        if (r1 > rhi) r1 = rhi;
        if (r1 <= r0) continue;
        LB_CUDA_CHECK(cudaMemcpy2DAsync(dA + r0 + (lb::i64)j0 * ldd, ldd * 8, A + r0 + (lb::i64)j0 * lda, (size_t)lda * 8,
                                        (size_t)(r1 - r0) * 8, w, cudaMemcpyHostToDevice, s));
    }
}

// one streamed factorization of the nn x nn diagonal block at (o, o): lb::potrf with the StreamOut hook, then whatever
// the hook did not send (small blocks), in upload-shaped trapezoids
static void potrf_streamed_block(cudaStream_t s, cudaStream_t copy_stream, cudaEvent_t ev, bool upper, int nn, int o, double* A,
                                 int lda, double* dA, lb::i64 ldd, int* dinfo, int cb) {
    lb::StreamOut so;
    so.host = A + o + (lb::i64)o * lda; so.ldh = lda; so.copy_stream = copy_stream; so.ev = ev; so.done_cols = 0;
    lb::stream_out() = &so;
    lb::potrf(s, upper ? 'U' : 'L', nn, dA + o + (lb::i64)o * ldd, ldd, dinfo);
    lb::stream_out() = nullptr;
    for (int j0 = so.done_cols; j0 < nn; j0 += cb) {
        const int w = imin(cb, nn - j0);
        const int r0 = upper ? so.done_cols : j0, r1 = upper ? j0 + w : nn;
        LB_CUDA_CHECK(cudaMemcpy2DAsync(A + (o + r0) + (lb::i64)(o + j0) * lda, (size_t)lda * 8,
                                        dA + (o + r0) + (lb::i64)(o + j0) * ldd, ldd * 8, (size_t)(r1 - r0) * 8, w,
                                        cudaMemcpyDeviceToHost, s));
    }
}

// DPOTRF, or DPOSV when nrhs > 0 (factor, then solve only if INFO = 0, dposv.f:176-183), for a pinned host matrix.
// Only the UPLO triangle crosses PCIe.  n >= 8192: one level of the DPOTRF2 recursion (dpotrf2.f:181-228) at the top --
// the leading n1 block columns (rows, for 'U') are uploaded first and factored / solved while the trailing block is
// still in flight; then ONE large-K DSYRK and the factorization of the trailing block.  Every finished block column
// of the factor is downloaded while the rest is still being computed (lb::StreamOut).
static int potrf_host_streamed(bool upper, int n, double* A, int lda, int nrhs = 0, double* B = nullptr, int ldb = 0) {
    static cudaStream_t copy_stream = nullptr, up_stream = nullptr;
    static cudaEvent_t ev = nullptr, ev_up = nullptr, ev_x = nullptr;
    if (!copy_stream) {
        LB_CUDA_CHECK(cudaStreamCreateWithFlags(&copy_stream, cudaStreamNonBlocking));
        LB_CUDA_CHECK(cudaStreamCreateWithFlags(&up_stream, cudaStreamNonBlocking));
        LB_CUDA_CHECK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        LB_CUDA_CHECK(cudaEventCreateWithFlags(&ev_up, cudaEventDisableTiming));
        LB_CUDA_CHECK(cudaEventCreateWithFlags(&ev_x, cudaEventDisableTiming));
    }
    cudaStream_t s = host_stream();
    const lb::i64 ldd = ((lb::i64)n + 1) & ~1LL;
    double* dA = (double*)lb::ws_alloc(s, sizeof(double) * (size_t)ldd * n);
    double* dB = nrhs > 0 ? (double*)lb::ws_alloc(s, sizeof(double) * (size_t)ldd * nrhs) : nullptr;
    int* dinfo = (int*)lb::ws_alloc(s, 64);
    const int cb = 2048;
    const char ul = upper ? 'U' : 'L';
    int n1 = n;
    if (n >= 8192) n1 = imin(n - cb, ((n / 4 + cb - 1) / cb) * cb);
    const int n2 = n - n1;
    if (n2 == 0) {
        upload_triangle(s, upper, n, A, lda, dA, ldd, cb, 0, n, 0, n);
        if (nrhs > 0)
            LB_CUDA_CHECK(cudaMemcpy2DAsync(dB, ldd * 8, B, (size_t)ldb * 8, (size_t)n * 8, nrhs, cudaMemcpyHostToDevice, s));
        potrf_streamed_block(s, copy_stream, ev, upper, n, 0, A, lda, dA, ldd, dinfo, cb);
    } else {
        // the scratch must exist before the upload stream touches it
        LB_CUDA_CHECK(cudaEventRecord(ev_x, s));
        LB_CUDA_CHECK(cudaStreamWaitEvent(up_stream, ev_x, 0));
        if (!upper) {
            upload_triangle(s, false, n, A, lda, dA, ldd, cb, 0, n1, 0, n);                 // block columns 0..n1 (L11, A21)
            upload_triangle(up_stream, false, n, A, lda, dA, ldd, cb, n1, n, 0, n);         // A22
        } else {
            upload_triangle(s, true, n, A, lda, dA, ldd, cb, 0, n, 0, n1);                  // rows 0..n1 (U11, A12)
            upload_triangle(up_stream, true, n, A, lda, dA, ldd, cb, n1, n, n1, n);         // A22
        }
        if (nrhs > 0)
            LB_CUDA_CHECK(cudaMemcpy2DAsync(dB, ldd * 8, B, (size_t)ldb * 8, (size_t)n * 8, nrhs, cudaMemcpyHostToDevice, up_stream));
        LB_CUDA_CHECK(cudaEventRecord(ev_up, up_stream));
        potrf_streamed_block(s, copy_stream, ev, upper, n1, 0, A, lda, dA, ldd, dinfo, cb);                 // dpotrf2.f:185
        double* dA22 = dA + n1 + (lb::i64)n1 * ldd;
        if (!upper) {
            double* dA21 = dA + n1;
            lb::trsm(s, 'R', 'L', 'T', 'N', n2, n1, 1.0, dA, ldd, dA21, ldd);                               // dpotrf2.f:217
            LB_CUDA_CHECK(cudaEventRecord(ev_x, s));
            LB_CUDA_CHECK(cudaStreamWaitEvent(copy_stream, ev_x, 0));
            LB_CUDA_CHECK(cudaMemcpy2DAsync(A + n1, (size_t)lda * 8, dA21, ldd * 8, (size_t)n2 * 8, n1, cudaMemcpyDeviceToHost,
                                            copy_stream));                                                 // L21 is final
            LB_CUDA_CHECK(cudaStreamWaitEvent(s, ev_up, 0));
            lb::syrk(s, 'L', 'N', n2, n1, -1.0, dA21, ldd, 1.0, dA22, ldd);                                 // dpotrf2.f:222
        } else {
            double* dA12 = dA + (lb::i64)n1 * ldd;
            lb::trsm(s, 'L', 'U', 'T', 'N', n1, n2, 1.0, dA, ldd, dA12, ldd);                               // dpotrf2.f:201
            LB_CUDA_CHECK(cudaEventRecord(ev_x, s));
            LB_CUDA_CHECK(cudaStreamWaitEvent(copy_stream, ev_x, 0));
            LB_CUDA_CHECK(cudaMemcpy2DAsync(A + (lb::i64)n1 * lda, (size_t)lda * 8, dA12, ldd * 8, (size_t)n1 * 8, n2,
                                            cudaMemcpyDeviceToHost, copy_stream));                         // U12 is final
            LB_CUDA_CHECK(cudaStreamWaitEvent(s, ev_up, 0));
            lb::syrk(s, 'U', 'T', n2, n1, -1.0, dA12, ldd, 1.0, dA22, ldd);                                 // dpotrf2.f:206
        }
        potrf_streamed_block(s, copy_stream, ev, upper, n2, n1, A, lda, dA, ldd, dinfo + 8, cb);            // dpotrf2.f:227
        lb::info_max_offset(s, dinfo, dinfo + 8, n1);                                                       // dpotrf2.f:228-231
    }
    int hinfo = 0;
    LB_CUDA_CHECK(cudaMemcpyAsync(&hinfo, dinfo, sizeof(int), cudaMemcpyDeviceToHost, s));
    if (nrhs > 0) {
        LB_CUDA_CHECK(cudaStreamSynchronize(s));               // INFO decides whether the solve happens
        if (hinfo == 0 && lb::pending_cuda_error() == 0) {
            LB_CUDA_CHECK(cudaStreamWaitEvent(s, ev_up, 0));
            lb::potrs(s, ul, n, nrhs, dA, ldd, dB, ldd);
            LB_CUDA_CHECK(cudaMemcpy2DAsync(B, (size_t)ldb * 8, dB, ldd * 8, (size_t)n * 8, nrhs, cudaMemcpyDeviceToHost, s));
        }
    }
    LB_CUDA_CHECK(cudaStreamSynchronize(copy_stream));
    LB_CUDA_CHECK(cudaStreamSynchronize(up_stream));
    lb::ws_free(s, dA);
    if (dB) lb::ws_free(s, dB);
    lb::ws_free(s, dinfo);
    LB_CUDA_CHECK(cudaStreamSynchronize(s));
    int e = lb::take_cuda_error();
    if (e) return -1001 - e;
    return hinfo;
}

static void potrf_common(bool recursive, const char* uplo, const int* n, double* A, const int* lda, int* info) {
    *info = 0;
    const bool upper = same(uplo, 'U');
    if (!upper && !same(uplo, 'L')) *info = -1;
    else if (*n < 0) *info = -2;
    else if (*lda < imax(1, *n)) *info = -4;
    if (*info != 0) { call_xerbla(recursive ? "DPOTRF2" : "DPOTRF", -*info); return; }
    if (*n == 0) return;
    if (!device_ok(info)) return;
    std::lock_guard<std::mutex> lock(g_abi_mutex);
    TempPin pin(A, ((size_t)*lda * (size_t)(*n - 1) + (size_t)*n) * sizeof(double));
    if (!recursive && *n >= 2048 && ptr_kind(A) == PK_PINNED && 2048 % lb::potrf_block() == 0) {
        *info = potrf_host_streamed(upper, *n, A, *lda);
        return;
    }
    Ctx c; c.scan({A});
    lb::i64 la;
    double* dA = c.mat(A, *n, *n, *lda, true, true, &la);
    int* dinfo = c.dev_info();
    if (recursive) lb::potrf2(c.s, upper ? 'U' : 'L', *n, dA, la, dinfo);
    else lb::potrf(c.s, upper ? 'U' : 'L', *n, dA, la, dinfo);
    *info = c.finish(dinfo);
}
void dpotrf_(const char* uplo, const int* n, double* A, const int* lda, int* info, size_t) { potrf_common(false, uplo, n, A, lda, info); }
void dpotrf2_(const char* uplo, const int* n, double* A, const int* lda, int* info, size_t) { potrf_common(true, uplo, n, A, lda, info); }

void dpotrs_(const char* uplo, const int* n, const int* nrhs, const double* A, const int* lda, double* B, const int* ldb,
             int* info, size_t) {
    *info = 0;
    const bool upper = same(uplo, 'U');
    if (!upper && !same(uplo, 'L')) *info = -1;
    else if (*n < 0) *info = -2;
    else if (*nrhs < 0) *info = -3;
    else if (*lda < imax(1, *n)) *info = -5;
    else if (*ldb < imax(1, *n)) *info = -7;
    if (*info != 0) { call_xerbla("DPOTRS", -*info); return; }
    if (*n == 0 || *nrhs == 0) return;
    if (!device_ok(info)) return;
    std::lock_guard<std::mutex> lock(g_abi_mutex);
    Ctx c; c.scan({A, B});
    lb::i64 la, lbb;
    const double* dA = c.mat(A, *n, *n, *lda, true, false, &la);
    double* dB = c.mat(B, *n, *nrhs, *ldb, true, true, &lbb);
    lb::potrs(c.s, upper ? 'U' : 'L', *n, *nrhs, dA, la, dB, lbb);
    int r = c.finish();
    if (r) *info = r;
}

void dposv_(const char* uplo, const int* n, const int* nrhs, double* A, const int* lda, double* B, const int* ldb,
            int* info, size_t) {
    *info = 0;
    const bool upper = same(uplo, 'U');
    if (!upper && !same(uplo, 'L')) *info = -1;
    else if (*n < 0) *info = -2;
    else if (*nrhs < 0) *info = -3;
    else if (*lda < imax(1, *n)) *info = -5;
    else if (*ldb < imax(1, *n)) *info = -7;
    if (*info != 0) { call_xerbla("DPOSV ", -*info); return; }
    if (*n == 0) return;
    if (!device_ok(info)) return;
    std::lock_guard<std::mutex> lock(g_abi_mutex);
    TempPin pin(A, ((size_t)*lda * (size_t)(*n - 1) + (size_t)*n) * sizeof(double));
    if (*n >= 2048 && ptr_kind(A) == PK_PINNED && (*nrhs == 0 || ptr_kind(B) == PK_PINNED || ptr_kind(B) == PK_HOST) &&
        2048 % lb::potrf_block() == 0) {
        *info = potrf_host_streamed(upper, *n, A, *lda, *nrhs, B, *ldb);
        return;
    }
    Ctx c; c.scan({A, B});
    lb::i64 la, lbb;
    double* dA = c.mat(A, *n, *n, *lda, true, true, &la);
    double* dB = c.mat(B, *n, *nrhs, *ldb, true, true, &lbb);
    int* dinfo = c.dev_info();
    lb::potrf(c.s, upper ? 'U' : 'L', *n, dA, la, dinfo);
    int hinfo = 0;
    LB_CUDA_CHECK(cudaMemcpyAsync(&hinfo, dinfo, sizeof(int), cudaMemcpyDeviceToHost, c.s));
    LB_CUDA_CHECK(cudaStreamSynchronize(c.s));
    if (hinfo == 0 && *nrhs > 0) lb::potrs(c.s, upper ? 'U' : 'L', *n, *nrhs, dA, la, dB, lbb);   // dposv.f:180-183
    else if (hinfo != 0) { for (auto& it : c.items) if (it.dev == dB) it.out = false; }
    *info = c.finish(dinfo);
}

// ================================================================================================ QR
// Reference block sizes that only matter for the WORK(1) protocol (SRC/ilaenv.f:296-302, 623-630)
static const int REF_NB_GEQRF = 32, REF_NX_GEQRF = 128;

void dgeqrf_(const int* m, const int* n, double* A, const int* lda, double* tau, double* work, const int* lwork, int* info) {
    const int k = imin(*m, *n);
    *info = 0;
    const bool lquery = (*lwork == -1);
    if (*m < 0) *info = -1;
    else if (*n < 0) *info = -2;
    else if (*lda < imax(1, *m)) *info = -4;
    else if (!lquery) { if (*lwork <= 0 || (*m > 0 && *lwork < imax(1, *n))) *info = -7; }
    if (*info != 0) { call_xerbla("DGEQRF", -*info); return; }
    if (lquery) { work[0] = (k == 0) ? 1.0 : (double)*n * REF_NB_GEQRF; return; }    // dgeqrf.f:197-204
    if (k == 0) { work[0] = 1.0; return; }
    if (!device_ok(info)) return;
    {
        std::lock_guard<std::mutex> lock(g_abi_mutex);
        Ctx c; c.scan({A, tau});
        lb::i64 la;
        double* dA = c.mat(A, *m, *n, *lda, true, true, &la);
        double* dt = c.vec<double>(tau, (size_t)k, false, true);
        lb::geqrf(c.s, *m, *n, dA, la, dt);
        int r = c.finish();
        if (r) *info = r;
    }
    // WORK(1) = IWS as the reference would report it (dgeqrf.f:216-236,278); device scratch replaces WORK
    int iws = *n;
    if (REF_NB_GEQRF > 1 && REF_NB_GEQRF < k && REF_NX_GEQRF < k) iws = *n * REF_NB_GEQRF;
    if (ptr_kind(work) != PK_DEVICE) work[0] = (double)iws;
}

void dgeqr2_(const int* m, const int* n, double* A, const int* lda, double* tau, double* work, int* info) {
    (void)work;
    *info = 0;
    if (*m < 0) *info = -1;
    else if (*n < 0) *info = -2;
    else if (*lda < imax(1, *m)) *info = -4;
    if (*info != 0) { call_xerbla("DGEQR2", -*info); return; }
    const int k = imin(*m, *n);
    if (k == 0) return;
    if (!device_ok(info)) return;
    std::lock_guard<std::mutex> lock(g_abi_mutex);
    Ctx c; c.scan({A, tau});
    lb::i64 la;
    double* dA = c.mat(A, *m, *n, *lda, true, true, &la);
    double* dt = c.vec<double>(tau, (size_t)k, false, true);
    lb::geqr2(c.s, *m, *n, dA, la, dt);
    int r = c.finish();
    if (r) *info = r;
}

// SRC/dlarft.f:160 -- all four DIRECT x STOREV schemes (the reference's own DORGLQ / DGERQF / DGEQLF / DORMxx call the others
// through this symbol when the library is preloaded); no argument checking, no INFO, like the reference
void dlarft_(const char* direct, const char* storev, const int* n, const int* k, const double* V, const int* ldv,
             const double* tau, double* T, const int* ldt, size_t, size_t) {
    if (*n == 0 || *k == 0) return;
    const bool backward = !same(direct, 'F'), rowwise = !same(storev, 'C');
    if (!device_ok(nullptr)) return;
    std::lock_guard<std::mutex> lock(g_abi_mutex);
    Ctx c; c.scan({V, tau, T});
    lb::i64 lv, lt;
    const double* dV = rowwise ? c.mat(V, *k, *n, *ldv, true, false, &lv) : c.mat(V, *n, *k, *ldv, true, false, &lv);
    const double* dtau = c.vec<double>(tau, (size_t)*k, true, false);
    double* dT = c.mat(T, *k, *k, *ldt, true, true, &lt);   // the other triangle travels unchanged
    lb::larft_general(c.s, backward, rowwise, *n, *k, dV, lv, dtau, dT, lt);
    c.finish();
}

// SRC/dlarfb.f:192 -- all SIDE / TRANS / DIRECT / STOREV combinations; WORK is not used (device scratch)
void dlarfb_(const char* side, const char* trans, const char* direct, const char* storev, const int* m, const int* n,
             const int* k, const double* V, const int* ldv, const double* T, const int* ldt, double* C, const int* ldc,
             double* work, const int* ldwork, size_t, size_t, size_t, size_t) {
    (void)work; (void)ldwork;
    if (*m <= 0 || *n <= 0) return;
    if (*k <= 0) return;
    const bool backward = !same(direct, 'F'), rowwise = !same(storev, 'C');
    if (!device_ok(nullptr)) return;
    std::lock_guard<std::mutex> lock(g_abi_mutex);
    Ctx c; c.scan({V, T, C});
    lb::i64 lv, lt, lc;
    const bool left = same(side, 'L');
    const int nv = left ? *m : *n;
    const double* dV = rowwise ? c.mat(V, *k, nv, *ldv, true, false, &lv) : c.mat(V, nv, *k, *ldv, true, false, &lv);
    const double* dT = c.mat(T, *k, *k, *ldt, true, false, &lt);
    double* dC = c.mat(C, *m, *n, *ldc, true, true, &lc);
    lb::larfb_general(c.s, left ? 'L' : 'R', same(trans, 'N') ? 'N' : 'T', backward, rowwise, *m, *n, *k, dV, lv, dT, lt, dC, lc);
    c.finish();
}

// DORGQR / DORMQR.  Reference block sizes for the WORK(1) protocol: ilaenv.f:416-436 (NB=32),
// dormqr.f:189-190 (NBMAX=64, LDT=65).
void dorgqr_(const int* m, const int* n, const int* k, double* A, const int* lda, const double* tau, double* work,
             const int* lwork, int* info) {
    *info = 0;
    const int nb = 32, nx = 128;
    const bool lquery = (*lwork == -1);
    if (*m < 0) *info = -1;
    else if (*n < 0 || *n > *m) *info = -2;
    else if (*k < 0 || *k > *n) *info = -3;
    else if (*lda < imax(1, *m)) *info = -5;
    else if (*lwork < imax(1, *n) && !lquery) *info = -8;
    if (*info != 0) { call_xerbla("DORGQR", -*info); return; }
    if (lquery) { work[0] = (double)(imax(1, *n) * nb); return; }                   // dorgqr.f:163-165
    if (*n <= 0) { work[0] = 1.0; return; }
    if (!device_ok(info)) return;
    {
        std::lock_guard<std::mutex> lock(g_abi_mutex);
        Ctx c; c.scan({A, tau});
        lb::i64 la;
        double* dA = c.mat(A, *m, *n, *lda, true, true, &la);
        const double* dt = c.vec<double>(const_cast<double*>(tau), (size_t)*k, true, false);
        lb::orgqr(c.s, *m, *n, *k, dA, la, dt);
        int r = c.finish();
        if (r) *info = r;
    }
    int iws = *n;
    if (nb > 1 && nb < *k && nx < *k) iws = *n * nb;                               // dorgqr.f:200-207,277
    if (ptr_kind(work) != PK_DEVICE) work[0] = (double)iws;
}

void dormqr_(const char* side, const char* trans, const int* m, const int* n, const int* k, const double* A, const int* lda,
             const double* tau, double* C, const int* ldc, double* work, const int* lwork, int* info, size_t, size_t) {
    *info = 0;
    const bool left = same(side, 'L'), notran = same(trans, 'N');
    const bool lquery = (*lwork == -1);
    const int nq = left ? *m : *n, nw = left ? imax(1, *n) : imax(1, *m);
    if (!left && !same(side, 'R')) *info = -1;
    else if (!notran && !same(trans, 'T')) *info = -2;
    else if (*m < 0) *info = -3;
    else if (*n < 0) *info = -4;
    else if (*k < 0 || *k > nq) *info = -5;
    else if (*lda < imax(1, nq)) *info = -7;
    else if (*ldc < imax(1, *m)) *info = -10;
    else if (*lwork < nw && !lquery) *info = -12;
    const int lwkopt = nw * 32 + 65 * 32;                                          // dormqr.f:240-244
    if (*info != 0) { call_xerbla("DORMQR", -*info); return; }
    if (ptr_kind(work) != PK_DEVICE) work[0] = (double)lwkopt;
    if (lquery) return;
    if (*m == 0 || *n == 0 || *k == 0) { if (ptr_kind(work) != PK_DEVICE) work[0] = 1.0; return; }
    if (!device_ok(info)) return;
    std::lock_guard<std::mutex> lock(g_abi_mutex);
    Ctx c; c.scan({A, tau, C});
    lb::i64 la, lc;
    const double* dA = c.mat(const_cast<double*>(A), nq, *k, *lda, true, false, &la);
    const double* dt = c.vec<double>(const_cast<double*>(tau), (size_t)*k, true, false);
    double* dC = c.mat(C, *m, *n, *ldc, true, true, &lc);
    lb::ormqr(c.s, left ? 'L' : 'R', notran ? 'N' : 'T', *m, *n, *k, dA, la, dt, dC, lc);
    int r = c.finish();
    if (r) *info = r;
}

// DGETRI (SRC/dgetri.f:114).  WORK(1) protocol with the reference block size 64 (ilaenv.f:361-367).
void dgetri_(const int* n, double* A, const int* lda, const int* ipiv, double* work, const int* lwork, int* info) {
    *info = 0;
    const int nb = 64;
    const bool lquery = (*lwork == -1);
    if (*n < 0) *info = -1;
    else if (*lda < imax(1, *n)) *info = -3;
    else if (*lwork < imax(1, *n) && !lquery) *info = -6;
    if (*info != 0) { call_xerbla("DGETRI", -*info); return; }
    if (ptr_kind(work) != PK_DEVICE) work[0] = (double)imax(1, *n * nb);          // dgetri.f:153-155
    if (lquery || *n == 0) return;
    if (!device_ok(info)) return;
    {
        std::lock_guard<std::mutex> lock(g_abi_mutex);
        Ctx c; c.scan({A, ipiv});
        lb::i64 la;
        double* dA = c.mat(A, *n, *n, *lda, true, true, &la);
        const int* dp = c.vec<int>(ipiv, (size_t)*n, true, false);
        int* dinfo = c.dev_info();
        lb::getri(c.s, *n, dA, la, dp, dinfo);
        *info = c.finish(dinfo);
    }
    int iws = *n;
    if (nb > 1 && nb < *n) iws = imax(*n * nb, 1);                                 // dgetri.f:186-196,257
    if (*info == 0 && ptr_kind(work) != PK_DEVICE) work[0] = (double)iws;
}

// DGEQRT / DGEMQRT (SRC/dgeqrt.f:139, SRC/dgemqrt.f:166).  WORK is not used (device scratch).
void dgeqrt_(const int* m, const int* n, const int* nb, double* A, const int* lda, double* T, const int* ldt, double* work,
             int* info) {
    (void)work;
    *info = 0;
    const int k = imin(*m, *n);
    if (*m < 0) *info = -1;
    else if (*n < 0) *info = -2;
    else if (*nb < 1 || (*nb > k && k > 0)) *info = -3;
    else if (*lda < imax(1, *m)) *info = -5;
    else if (*ldt < *nb) *info = -7;
    if (*info != 0) { call_xerbla("DGEQRT", -*info); return; }
    if (k == 0) return;
    if (!device_ok(info)) return;
    std::lock_guard<std::mutex> lock(g_abi_mutex);
    Ctx c; c.scan({A, T});
    lb::i64 la, lt;
    double* dA = c.mat(A, *m, *n, *lda, true, true, &la);
    double* dT = c.mat(T, *nb, k, *ldt, true, true, &lt);          // in + out: the strictly lower parts of the T blocks stay
    lb::geqrt(c.s, *m, *n, *nb, dA, la, dT, lt);
    int r = c.finish();
    if (r) *info = r;
}

// SRC/dgeqrt3.f:129 DGEQRT3(M,N,A,LDA,T,LDT,INFO): recursive QR with the full N x N compact-WY factor T (M >= N).  The panel
// recursion of csrc/geqrf.cu IS this algorithm (dgeqrt3.f:157-250: split N1 = N/2, factor left, apply to the right, factor
// the right, T12 = -T1 (V1^T V2) T2), so the routine is DGEQRT with one block of width N.
void dgeqrt3_(const int* m, const int* n, double* A, const int* lda, double* T, const int* ldt, int* info) {
    *info = 0;
    if (*n < 0) *info = -2;
    else if (*m < *n) *info = -1;
    else if (*lda < imax(1, *m)) *info = -4;
    else if (*ldt < imax(1, *n)) *info = -6;
    if (*info != 0) { call_xerbla("DGEQRT3", -*info); return; }
    if (*n == 0) return;
    if (!device_ok(info)) return;
    std::lock_guard<std::mutex> lock(g_abi_mutex);
    Ctx c; c.scan({A, T});
    lb::i64 la, lt;
    double* dA = c.mat(A, *m, *n, *lda, true, true, &la);
    double* dT = c.mat(T, *n, *n, *ldt, true, true, &lt);           // the part below the diagonal is not used and stays
    lb::geqrt(c.s, *m, *n, *n, dA, la, dT, lt);
    int r = c.finish();
    if (r) *info = r;
}

// SRC/dlatsqr.f:170 DLATSQR(M,N,MB,NB,A,LDA,T,LDT,WORK,LWORK,INFO); WORK is not used (device scratch), WORK(1) = N*NB
void dlatsqr_(const int* m, const int* n, const int* mb, const int* nb, double* A, const int* lda, double* T, const int* ldt,
              double* work, const int* lwork, int* info) {
    *info = 0;
    const bool lquery = (*lwork == -1);
    const int minmn = imin(*m, *n);
    const int lwmin = (minmn == 0) ? 1 : *n * *nb;
    if (*m < 0) *info = -1;
    else if (*n < 0 || *m < *n) *info = -2;
    else if (*mb < 1) *info = -3;
    else if (*nb < 1 || (*nb > *n && *n > 0)) *info = -4;
    else if (*lda < imax(1, *m)) *info = -6;
    else if (*ldt < *nb) *info = -8;
    else if (*lwork < lwmin && !lquery) *info = -10;
    if (*info == 0 && ptr_kind(work) != PK_DEVICE) work[0] = (double)lwmin;
    if (*info != 0) { call_xerbla("DLATSQR", -*info); return; }
    if (lquery || minmn == 0) return;
    if (!device_ok(info)) return;
    std::lock_guard<std::mutex> lock(g_abi_mutex);
    int tcols = *n;
    if (!(*mb <= *n || *mb >= *m)) tcols = *n * ((*m - *n + (*mb - *n) - 1) / (*mb - *n));     // dlatsqr.f:100-104
    Ctx c; c.scan({A, T});
    lb::i64 la, lt;
    double* dA = c.mat(A, *m, *n, *lda, true, true, &la);
    double* dT = c.mat(T, *nb, tcols, *ldt, true, true, &lt);
    lb::latsqr(c.s, *m, *n, *mb, *nb, dA, la, dT, lt);
    int r = c.finish();
    if (r) *info = r;
}

void dgemqrt_(const char* side, const char* trans, const int* m, const int* n, const int* k, const int* nb, const double* V,
              const int* ldv, const double* T, const int* ldt, double* C, const int* ldc, double* work, int* info, size_t, size_t) {
    (void)work;
    *info = 0;
    const bool left = same(side, 'L'), right = same(side, 'R'), tran = same(trans, 'T'), notran = same(trans, 'N');
    const int q = left ? *m : *n;
    if (!left && !right) *info = -1;
    else if (!tran && !notran) *info = -2;
    else if (*m < 0) *info = -3;
    else if (*n < 0) *info = -4;
    else if (*k < 0 || *k > q) *info = -5;
    else if (*nb < 1 || (*nb > *k && *k > 0)) *info = -6;
    else if (*ldv < imax(1, q)) *info = -8;
    else if (*ldt < *nb) *info = -10;
    else if (*ldc < imax(1, *m)) *info = -12;
    if (*info != 0) { call_xerbla("DGEMQRT", -*info); return; }
    if (*m == 0 || *n == 0 || *k == 0) return;
    if (!device_ok(info)) return;
    std::lock_guard<std::mutex> lock(g_abi_mutex);
    Ctx c; c.scan({V, T, C});
    lb::i64 lv, lt, lc;
    const double* dV = c.mat(const_cast<double*>(V), q, *k, *ldv, true, false, &lv);
    const double* dT = c.mat(const_cast<double*>(T), *nb, *k, *ldt, true, false, &lt);
    double* dC = c.mat(C, *m, *n, *ldc, true, true, &lc);
    lb::gemqrt(c.s, left ? 'L' : 'R', notran ? 'N' : 'T', *m, *n, *k, *nb, dV, lv, dT, lt, dC, lc);
    int r = c.finish();
    if (r) *info = r;
}

// ---- DGELQF / DORMLQ / DGELS (the main consumer of the QR path) ------------------------------------
void dgelqf_(const int* m, const int* n, double* A, const int* lda, double* tau, double* work, const int* lwork, int* info) {
    const int k = imin(*m, *n), nb = 32;
    *info = 0;
    const bool lquery = (*lwork == -1);
    if (*m < 0) *info = -1;
    else if (*n < 0) *info = -2;
    else if (*lda < imax(1, *m)) *info = -4;
    else if (!lquery) { if (*lwork <= 0 || (*n > 0 && *lwork < imax(1, *m))) *info = -7; }
    if (*info != 0) { call_xerbla("DGELQF", -*info); return; }
    if (lquery) { work[0] = (k == 0) ? 1.0 : (double)*m * nb; return; }              // dgelqf.f:189-197
    if (k == 0) { work[0] = 1.0; return; }
    if (!device_ok(info)) return;
    {
        std::lock_guard<std::mutex> lock(g_abi_mutex);
        Ctx c; c.scan({A, tau});
        lb::i64 la;
        double* dA = c.mat(A, *m, *n, *lda, true, true, &la);
        double* dt = c.vec<double>(tau, (size_t)k, false, true);
        lb::gelqf(c.s, *m, *n, dA, la, dt);
        int r = c.finish();
        if (r) *info = r;
    }
    int iws = *m;
    if (nb > 1 && nb < k && 128 < k) iws = *m * nb;                                     // dgelqf.f:207-231, NX = 128
    if (ptr_kind(work) != PK_DEVICE) work[0] = (double)iws;
}

void dormlq_(const char* side, const char* trans, const int* m, const int* n, const int* k, const double* A, const int* lda,
             const double* tau, double* C, const int* ldc, double* work, const int* lwork, int* info, size_t, size_t) {
    *info = 0;
    const bool left = same(side, 'L'), notran = same(trans, 'N');
    const bool lquery = (*lwork == -1);
    const int nq = left ? *m : *n, nw = left ? imax(1, *n) : imax(1, *m);
    if (!left && !same(side, 'R')) *info = -1;
    else if (!notran && !same(trans, 'T')) *info = -2;
    else if (*m < 0) *info = -3;
    else if (*n < 0) *info = -4;
    else if (*k < 0 || *k > nq) *info = -5;
    else if (*lda < imax(1, *k)) *info = -7;
    else if (*ldc < imax(1, *m)) *info = -10;
    else if (*lwork < nw && !lquery) *info = -12;
    if (*info != 0) { call_xerbla("DORMLQ", -*info); return; }
    if (ptr_kind(work) != PK_DEVICE) work[0] = (double)(nw * 32 + 65 * 32);            // dormlq.f:238-242
    if (lquery) return;
    if (*m == 0 || *n == 0 || *k == 0) { if (ptr_kind(work) != PK_DEVICE) work[0] = 1.0; return; }
    if (!device_ok(info)) return;
    std::lock_guard<std::mutex> lock(g_abi_mutex);
    Ctx c; c.scan({A, tau, C});
    lb::i64 la, lc;
    const double* dA = c.mat(const_cast<double*>(A), *k, nq, *lda, true, false, &la);
    const double* dt = c.vec<double>(const_cast<double*>(tau), (size_t)*k, true, false);
    double* dC = c.mat(C, *m, *n, *ldc, true, true, &lc);
    lb::ormlq(c.s, left ? 'L' : 'R', notran ? 'N' : 'T', *m, *n, *k, dA, la, dt, dC, lc);
    int r = c.finish();
    if (r) *info = r;
}

// DLASCL 'G' (SRC/dlascl.f:233-285): the multipliers are computed on the host exactly as the reference does, each step is one
// scaling pass on the device
static void dev_lascl(cudaStream_t s, double cfrom, double cto, int m, int n, double* dA, lb::i64 lda) {
    if (m == 0 || n == 0) return;
    const double smlnum = 2.2250738585072014e-308, bignum = 1.0 / smlnum;
    double cfromc = cfrom, ctoc = cto;
    for (;;) {
        double cfrom1 = cfromc * smlnum, mul;
        bool done;
        if (cfrom1 == cfromc) { mul = ctoc / cfromc; done = true; }
        else {
            double cto1 = ctoc / bignum;
            if (cto1 == ctoc) { mul = ctoc; done = true; cfromc = 1.0; }
            else if (fabs(cfrom1) > fabs(ctoc) && ctoc != 0.0) { mul = smlnum; done = false; cfromc = cfrom1; }
            else if (fabs(cto1) > fabs(cfromc)) { mul = bignum; done = false; ctoc = cto1; }
            else { mul = ctoc / cfromc; done = true; if (mul == 1.0) return; }
        }
        lb::scale_matrix(s, m, n, mul, dA, lda);
        if (done) break;
    }
}

// first exactly-zero diagonal entry of a device triangular matrix (DTRTRS, dtrtrs.f:214-220); 0 if none.  Synchronises.
static int dev_first_zero_diag(cudaStream_t s, int n, const double* dA, lb::i64 lda) {
    std::vector<double> d((size_t)n);
    LB_CUDA_CHECK(cudaMemcpy2DAsync(d.data(), 8, dA, (size_t)(lda + 1) * 8, 8, (size_t)n, cudaMemcpyDeviceToHost, s));
    LB_CUDA_CHECK(cudaStreamSynchronize(s));
    for (int i = 0; i < n; ++i) if (d[(size_t)i] == 0.0) return i + 1;
    return 0;
}

void dgels_(const char* trans, const int* m, const int* n, const int* nrhs, double* A, const int* lda, double* B, const int* ldb,
            double* work, const int* lwork, int* info, size_t) {
    const int mn = imin(*m, *n);
    const bool lquery = (*lwork == -1);
    const bool tn = same(trans, 'N');
    *info = 0;
    if (!tn && !same(trans, 'T')) *info = -1;
    else if (*m < 0) *info = -2;
    else if (*n < 0) *info = -3;
    else if (*nrhs < 0) *info = -4;
    else if (*lda < imax(1, *m)) *info = -6;
    else if (*ldb < imax(1, imax(*m, *n))) *info = -8;
    else if (*lwork < imax(1, mn + imax(mn, *nrhs)) && !lquery) *info = -10;
    const int wsize = imax(1, mn + imax(mn, *nrhs) * 32);                               // dgels.f:262-286, NB = 32
    if ((*info == 0 || *info == -10) && ptr_kind(work) != PK_DEVICE) work[0] = (double)wsize;
    if (*info != 0) { call_xerbla("DGELS ", -*info); return; }
    if (lquery) return;
    const int mx = imax(*m, *n);
    if (!device_ok(info)) return;
    std::lock_guard<std::mutex> lock(g_abi_mutex);
    Ctx c; c.scan({A, B});
    lb::i64 la, lbb;
    double* dB = c.mat(B, mx, *nrhs, *ldb, true, true, &lbb);
    if (imin(mn, *nrhs) == 0) {                                                          // dgels.f:297-300
        if (mx > 0 && *nrhs > 0) lb::laset(c.s, 'A', mx, *nrhs, 0.0, 0.0, dB, lbb);
        int r = c.finish();
        if (r) *info = r;
        return;
    }
    double* dA = c.mat(A, *m, *n, *lda, true, true, &la);
    const bool tpsd = !tn;
    const double smlnum = 2.2250738585072014e-308 / 2.220446049250313e-16, bignum = 1.0 / smlnum;   // DLAMCH('S')/DLAMCH('P')
    const double anrm = lb::amax_abs(c.s, *m, *n, dA, la);
    int iascl = 0, ibscl = 0, scllen = 0, hinfo = 0;
    bool solved = true;
    if (anrm > 0.0 && anrm < smlnum) { dev_lascl(c.s, anrm, smlnum, *m, *n, dA, la); iascl = 1; }
    else if (anrm > bignum) { dev_lascl(c.s, anrm, bignum, *m, *n, dA, la); iascl = 2; }
    else if (anrm == 0.0) {                                                              // dgels.f:326-333
        lb::laset(c.s, 'A', mx, *nrhs, 0.0, 0.0, dB, lbb);
        int r = c.finish();
        if (r) *info = r;
        if (ptr_kind(work) != PK_DEVICE) work[0] = (double)wsize;
        return;
    }
    const int brow = tpsd ? *n : *m;
    const double bnrm = lb::amax_abs(c.s, brow, *nrhs, dB, lbb);
    if (bnrm > 0.0 && bnrm < smlnum) { dev_lascl(c.s, bnrm, smlnum, brow, *nrhs, dB, lbb); ibscl = 1; }
    else if (bnrm > bignum) { dev_lascl(c.s, bnrm, bignum, brow, *nrhs, dB, lbb); ibscl = 2; }
    double* dtau = (double*)lb::ws_alloc(c.s, sizeof(double) * (size_t)mn);
    if (*m >= *n) {
        lb::geqrf(c.s, *m, *n, dA, la, dtau);                                           // dgels.f:359
        if (!tpsd) {
            lb::ormqr(c.s, 'L', 'T', *m, *nrhs, *n, dA, la, dtau, dB, lbb);            // dgels.f:371
            hinfo = dev_first_zero_diag(c.s, *n, dA, la);                               // DTRTRS, dgels.f:379
            if (hinfo == 0) lb::trsm(c.s, 'L', 'U', 'N', 'N', *n, *nrhs, 1.0, dA, la, dB, lbb); else solved = false;
            scllen = *n;
        } else {
            hinfo = dev_first_zero_diag(c.s, *n, dA, la);                               // dgels.f:395
            if (hinfo == 0) {
                lb::trsm(c.s, 'L', 'U', 'T', 'N', *n, *nrhs, 1.0, dA, la, dB, lbb);
                if (*m > *n) lb::laset(c.s, 'A', *m - *n, *nrhs, 0.0, 0.0, dB + *n, lbb);               // dgels.f:404-408
                lb::ormqr(c.s, 'L', 'N', *m, *nrhs, *n, dA, la, dtau, dB, lbb);        // dgels.f:412
            } else solved = false;
            scllen = *m;
        }
    } else {
        lb::gelqf(c.s, *m, *n, dA, la, dtau);                                           // dgels.f:426
        if (!tpsd) {
            hinfo = dev_first_zero_diag(c.s, *m, dA, la);                               // dgels.f:438
            if (hinfo == 0) {
                lb::trsm(c.s, 'L', 'L', 'N', 'N', *m, *nrhs, 1.0, dA, la, dB, lbb);
                lb::laset(c.s, 'A', *n - *m, *nrhs, 0.0, 0.0, dB + *m, lbb);                             // dgels.f:448-452
                lb::ormlq(c.s, 'L', 'T', *n, *nrhs, *m, dA, la, dtau, dB, lbb);        // dgels.f:456
            } else solved = false;
            scllen = *n;
        } else {
            lb::ormlq(c.s, 'L', 'N', *n, *nrhs, *m, dA, la, dtau, dB, lbb);            // dgels.f:470
            hinfo = dev_first_zero_diag(c.s, *m, dA, la);                               // dgels.f:478
            if (hinfo == 0) lb::trsm(c.s, 'L', 'L', 'T', 'N', *m, *nrhs, 1.0, dA, la, dB, lbb); else solved = false;
            scllen = *m;
        }
    }
    if (solved) {                                                                        // dgels.f:494-509
        if (iascl == 1) dev_lascl(c.s, anrm, smlnum, scllen, *nrhs, dB, lbb);
        else if (iascl == 2) dev_lascl(c.s, anrm, bignum, scllen, *nrhs, dB, lbb);
        if (ibscl == 1) dev_lascl(c.s, smlnum, bnrm, scllen, *nrhs, dB, lbb);
        else if (ibscl == 2) dev_lascl(c.s, bignum, bnrm, scllen, *nrhs, dB, lbb);
    }
    lb::ws_free(c.s, dtau);
    int r = c.finish();
    *info = r ? r : hinfo;
    if (solved && ptr_kind(work) != PK_DEVICE) work[0] = (double)wsize;
}

// DGERFS (SRC/dgerfs.f:235-440): iterative refinement with BERR / FERR.  The O(n^2) pieces -- residual, |A||x|, the solves
// with the factors -- run on the device; the scalar logic and DLACN2's state machine run on the host on length-n vectors.
// DGERFS on device operands (dgerfs.f:270-420): residual, |A||x| and the solves on the device, BERR / FERR logic and DLACN2 on
// the host.  ferr / berr are host arrays.
static void gerfs_device(cudaStream_t st, bool notran, int N, int nrhs, const double* dA, lb::i64 la, const double* dAF, lb::i64 laf,
                         const int* dp, const double* dB, lb::i64 lbb, double* dX, lb::i64 lx, double* ferr, double* berr) {
    double* dr = (double*)lb::ws_alloc(st, sizeof(double) * (size_t)N);
    double* dw = (double*)lb::ws_alloc(st, sizeof(double) * (size_t)N);
    std::vector<double> r((size_t)N), w((size_t)N), v((size_t)N), xh((size_t)N);
    std::vector<int> isgn((size_t)N);
    const char tr = notran ? 'N' : 'T', trt = notran ? 'T' : 'N';
    const int itmax = 5, nz = N + 1;
    const double eps = 1.1102230246251565e-16, safmin = 2.2250738585072014e-308;
    const double safe1 = nz * safmin, safe2 = safe1 / eps;
    auto solve = [&](char t) { lb::getrs(st, t, N, 1, dAF, laf, dp, dr, N); };
    auto to_host = [&](const double* d, std::vector<double>& h) {
        LB_CUDA_CHECK(cudaMemcpyAsync(h.data(), d, sizeof(double) * (size_t)N, cudaMemcpyDeviceToHost, st));
        LB_CUDA_CHECK(cudaStreamSynchronize(st));
    };
    auto to_dev = [&](const std::vector<double>& h, double* d) {
        LB_CUDA_CHECK(cudaMemcpyAsync(d, h.data(), sizeof(double) * (size_t)N, cudaMemcpyHostToDevice, st));
        LB_CUDA_CHECK(cudaStreamSynchronize(st));
    };
    for (int j = 0; j < nrhs; ++j) {
        const double* dbj = dB + (lb::i64)j * lbb;
        double* dxj = dX + (lb::i64)j * lx;
        int count = 1;
        double lstres = 3.0;
        for (;;) {
            LB_CUDA_CHECK(cudaMemcpyAsync(dr, dbj, sizeof(double) * (size_t)N, cudaMemcpyDeviceToDevice, st));
            lb::gemm(st, tr, 'N', N, 1, N, -1.0, dA, la, dxj, N, 1.0, dr, N);                       // dgerfs.f:285-288
            if (notran) abs_gemv_n_kernel<<<(N + 127) / 128, 128, 0, st>>>(N, dA, la, dxj, dbj, dw);
            else abs_gemv_t_kernel<<<(N + 7) / 8, 256, 0, st>>>(N, dA, la, dxj, dbj, dw);
            to_host(dr, r);
            to_host(dw, w);
            double sm = 0.0;
            for (int i = 0; i < N; ++i) {
                if (w[(size_t)i] > safe2) sm = fmax(sm, fabs(r[(size_t)i]) / w[(size_t)i]);
                else sm = fmax(sm, (fabs(r[(size_t)i]) + safe1) / (w[(size_t)i] + safe1));
            }
            berr[j] = sm;
            if (berr[j] > eps && 2.0 * berr[j] <= lstres && count <= itmax) {                            // dgerfs.f:340-352
                solve(tr);
                vec_add_kernel<<<(N + 255) / 256, 256, 0, st>>>(N, dr, dxj);
                lstres = berr[j];
                ++count;
                continue;
            }
            break;
        }
        for (int i = 0; i < N; ++i) {                                                                    // dgerfs.f:376-384
            if (w[(size_t)i] > safe2) w[(size_t)i] = fabs(r[(size_t)i]) + nz * eps * w[(size_t)i];
            else w[(size_t)i] = fabs(r[(size_t)i]) + nz * eps * w[(size_t)i] + safe1;
        }
        int kase = 0, isave[3] = {0, 0, 0};
        for (;;) {
            host_dlacn2(N, v.data(), r.data(), isgn.data(), &ferr[j], &kase, isave);
            if (kase == 0) break;
            if (kase == 1) {                                                                             // dgerfs.f:392-398
                to_dev(r, dr);
                solve(trt);
                to_host(dr, r);
                for (int i = 0; i < N; ++i) r[(size_t)i] = w[(size_t)i] * r[(size_t)i];
            } else {                                                                                     // dgerfs.f:403-409
                for (int i = 0; i < N; ++i) r[(size_t)i] = w[(size_t)i] * r[(size_t)i];
                to_dev(r, dr);
                solve(tr);
                to_host(dr, r);
            }
        }
        to_host(dxj, xh);
        lstres = 0.0;
        for (int i = 0; i < N; ++i) lstres = fmax(lstres, fabs(xh[(size_t)i]));
        if (lstres != 0.0) ferr[j] = ferr[j] / lstres;
    }
    lb::ws_free(st, dr);
    lb::ws_free(st, dw);
}

void dgerfs_(const char* trans, const int* n, const int* nrhs, const double* A, const int* lda, const double* AF, const int* ldaf,
             const int* ipiv, const double* B, const int* ldb, double* X, const int* ldx, double* ferr, double* berr, double* work,
             int* iwork, int* info, size_t) {
    (void)work; (void)iwork;
    *info = 0;
    const bool notran = same(trans, 'N');
    if (!notran && !same(trans, 'T') && !same(trans, 'C')) *info = -1;
    else if (*n < 0) *info = -2;
    else if (*nrhs < 0) *info = -3;
    else if (*lda < imax(1, *n)) *info = -5;
    else if (*ldaf < imax(1, *n)) *info = -7;
    else if (*ldb < imax(1, *n)) *info = -10;
    else if (*ldx < imax(1, *n)) *info = -12;
    if (*info != 0) { call_xerbla("DGERFS", -*info); return; }
    if (*n == 0 || *nrhs == 0) { for (int j = 0; j < *nrhs; ++j) { ferr[j] = 0.0; berr[j] = 0.0; } return; }
    if (!device_ok(info)) return;
    std::lock_guard<std::mutex> lock(g_abi_mutex);
    const int N = *n;
    Ctx c; c.scan({A, AF, ipiv, B, X});
    lb::i64 la, laf, lbb, lx;
    const double* dA = c.mat(const_cast<double*>(A), N, N, *lda, true, false, &la);
    const double* dAF = c.mat(const_cast<double*>(AF), N, N, *ldaf, true, false, &laf);
    const int* dp = c.vec<int>(ipiv, (size_t)N, true, false);
    const double* dB = c.mat(const_cast<double*>(B), N, *nrhs, *ldb, true, false, &lbb);
    double* dX = c.mat(X, N, *nrhs, *ldx, true, true, &lx);
    gerfs_device(c.s, notran, N, *nrhs, dA, la, dAF, laf, dp, dB, lbb, dX, lx, ferr, berr);
    int rr = c.finish();
    if (rr) *info = rr;
}

// ================================================================================================ condition number / expert driver
// SRC/dlatrs.f:238 DLATRS(UPLO,TRANS,DIAG,NORMIN,N,A,LDA,X,SCALE,CNORM,INFO)
void dlatrs_(const char* uplo, const char* trans, const char* diag, const char* normin, const int* n, const double* A, const int* lda,
             double* x, double* scale, double* cnorm, int* info, size_t, size_t, size_t, size_t) {
    *info = 0;
    const bool upper = same(uplo, 'U'), notran = same(trans, 'N'), nounit = same(diag, 'N');
    if (!upper && !same(uplo, 'L')) *info = -1;
    else if (!notran && !same(trans, 'T') && !same(trans, 'C')) *info = -2;
    else if (!nounit && !same(diag, 'U')) *info = -3;
    else if (!same(normin, 'Y') && !same(normin, 'N')) *info = -4;
    else if (*n < 0) *info = -5;
    else if (*lda < imax(1, *n)) *info = -7;
    if (*info != 0) { call_xerbla("DLATRS", -*info); return; }
    *scale = 1.0;
    if (*n == 0) return;
    if (!device_ok(info)) return;
    std::lock_guard<std::mutex> lock(g_abi_mutex);
    Ctx c; c.scan({A, x, cnorm});
    lb::i64 la;
    const double* dA = c.mat(A, *n, *n, *lda, true, false, &la);
    double* dx = c.vec<double>(x, (size_t)*n, true, true);
    double* dc = c.vec<double>(cnorm, (size_t)*n, same(normin, 'Y'), true);
    std::vector<double> hdiag;
    *scale = lb::latrs(c.s, upper, notran, nounit, same(normin, 'Y'), *n, dA, la, dx, dc, hdiag);
    int r = c.finish();
    if (r) *info = r;
}

// SRC/dgecon.f:128 DGECON(NORM,N,A,LDA,ANORM,RCOND,WORK,IWORK,INFO); WORK / IWORK are not used (device scratch)
void dgecon_(const char* norm, const int* n, const double* A, const int* lda, const double* anorm, double* rcond, double* work,
             int* iwork, int* info, size_t) {
    (void)work; (void)iwork;
    *info = 0;
    const bool onenrm = same(norm, '1') || same(norm, 'O');
    if (!onenrm && !same(norm, 'I')) *info = -1;
    else if (*n < 0) *info = -2;
    else if (*lda < imax(1, *n)) *info = -4;
    else if (*anorm < 0.0) *info = -5;
    if (*info != 0) { call_xerbla("DGECON", -*info); return; }
    *rcond = 0.0;                                                       // dgecon.f:195-209
    if (*n == 0) { *rcond = 1.0; return; }
    else if (*anorm == 0.0) return;
    else if (*anorm != *anorm) { *rcond = *anorm; *info = -5; return; }
    else if (*anorm > DBL_MAX) { *info = -5; return; }
    if (!device_ok(info)) return;
    std::lock_guard<std::mutex> lock(g_abi_mutex);
    Ctx c; c.scan({A});
    lb::i64 la;
    const double* dA = c.mat(A, *n, *n, *lda, true, false, &la);
    const int ci = lb::gecon(c.s, onenrm, *n, dA, la, *anorm, rcond);
    int r = c.finish();
    *info = r ? r : ci;
}

// SRC/dgeequ.f:139 DGEEQU(M,N,A,LDA,R,C,ROWCND,COLCND,AMAX,INFO); R, C are host vectors
void dgeequ_(const int* m, const int* n, const double* A, const int* lda, double* r, double* c, double* rowcnd, double* colcnd,
             double* amax, int* info) {
    *info = 0;
    if (*m < 0) *info = -1;
    else if (*n < 0) *info = -2;
    else if (*lda < imax(1, *m)) *info = -4;
    if (*info != 0) { call_xerbla("DGEEQU", -*info); return; }
    if (*m == 0 || *n == 0) { *rowcnd = 1.0; *colcnd = 1.0; *amax = 0.0; return; }
    if (!device_ok(info)) return;
    std::lock_guard<std::mutex> lock(g_abi_mutex);
    Ctx cx; cx.scan({A});
    lb::i64 la;
    const double* dA = cx.mat(A, *m, *n, *lda, true, false, &la);
    const int gi = lb::geequ(cx.s, *m, *n, dA, la, r, c, rowcnd, colcnd, amax);
    int rr = cx.finish();
    *info = rr ? rr : gi;
}

// SRC/dlaqge.f:140 DLAQGE(M,N,A,LDA,R,C,ROWCND,COLCND,AMAX,EQUED); no INFO
void dlaqge_(const int* m, const int* n, double* A, const int* lda, const double* r, const double* c, const double* rowcnd,
             const double* colcnd, const double* amax, char* equed, size_t) {
    if (*m <= 0 || *n <= 0) { *equed = 'N'; return; }
    if (!device_ok(nullptr)) return;
    std::lock_guard<std::mutex> lock(g_abi_mutex);
    Ctx cx; cx.scan({A});
    lb::i64 la;
    double* dA = cx.mat(A, *m, *n, *lda, true, true, &la);
    *equed = lb::laqge(cx.s, *m, *n, dA, la, r, c, *rowcnd, *colcnd, *amax);
    if (*equed == 'N') { for (auto& it : cx.items) it.out = false; }      // nothing was scaled: leave A alone
    cx.finish();
}

// SRC/dgesvx.f:344 DGESVX: equilibrate (FACT='E'), factor, estimate the condition number, solve, refine, bound the errors.
// R, C, FERR, BERR are host vectors; WORK(1) returns the reciprocal pivot growth when WORK is host memory; IWORK is not used.
void dgesvx_(const char* fact, const char* trans, const int* n, const int* nrhs, double* A, const int* lda, double* AF, const int* ldaf,
             int* ipiv, char* equed, double* R, double* Cs, double* B, const int* ldb, double* X, const int* ldx, double* rcond,
             double* ferr, double* berr, double* work, int* iwork, int* info, size_t, size_t, size_t) {
    (void)iwork;
    *info = 0;
    const bool nofact = same(fact, 'N'), equil = same(fact, 'E'), notran = same(trans, 'N');
    const int N = *n;
    bool rowequ, colequ;
    const double smlnum = DBL_MIN, bignum = 1.0 / smlnum;
    double rowcnd = 1.0, colcnd = 1.0, amax = 0.0;
    if (nofact || equil) { *equed = 'N'; rowequ = false; colequ = false; }
    else { rowequ = same(equed, 'R') || same(equed, 'B'); colequ = same(equed, 'C') || same(equed, 'B'); }
    if (!nofact && !equil && !same(fact, 'F')) *info = -1;                 // dgesvx.f:390-447
    else if (!notran && !same(trans, 'T') && !same(trans, 'C')) *info = -2;
    else if (N < 0) *info = -3;
    else if (*nrhs < 0) *info = -4;
    else if (*lda < imax(1, N)) *info = -6;
    else if (*ldaf < imax(1, N)) *info = -8;
    else if (same(fact, 'F') && !(rowequ || colequ || same(equed, 'N'))) *info = -10;
    else {
        if (rowequ) {
            double rcmin = bignum, rcmax = 0.0;
            for (int j = 0; j < N; ++j) { rcmin = fmin(rcmin, R[j]); rcmax = fmax(rcmax, R[j]); }
            if (rcmin <= 0.0) *info = -11;
            else if (N > 0) rowcnd = fmax(rcmin, smlnum) / fmin(rcmax, bignum);
        }
        if (colequ && *info == 0) {
            double rcmin = bignum, rcmax = 0.0;
            for (int j = 0; j < N; ++j) { rcmin = fmin(rcmin, Cs[j]); rcmax = fmax(rcmax, Cs[j]); }
            if (rcmin <= 0.0) *info = -12;
            else if (N > 0) colcnd = fmax(rcmin, smlnum) / fmin(rcmax, bignum);
        }
        if (*info == 0) {
            if (*ldb < imax(1, N)) *info = -14;
            else if (*ldx < imax(1, N)) *info = -16;
        }
    }
    if (*info != 0) { call_xerbla("DGESVX", -*info); return; }
    if (!device_ok(info)) return;
    std::lock_guard<std::mutex> lock(g_abi_mutex);
    Ctx c; c.scan({A, AF, ipiv, B, X});
    lb::i64 la, laf, lbb, lx;
    double* dA = c.mat(A, N, N, *lda, true, equil, &la);
    double* dAF = c.mat(AF, N, N, *ldaf, !(nofact || equil), nofact || equil, &laf);
    int* dp = c.vec<int>(ipiv, (size_t)N, !(nofact || equil), nofact || equil);
    double* dB = c.mat(B, N, *nrhs, *ldb, true, true, &lbb);
    double* dX = c.mat(X, N, *nrhs, *ldx, false, true, &lx);
    auto set_out = [&](void* dev, bool out) { for (auto& it : c.items) if (it.dev == dev) it.out = out; };
    if (equil && N > 0) {                                                   // dgesvx.f:453-467
        const int infequ = lb::geequ(c.s, N, N, dA, la, R, Cs, &rowcnd, &colcnd, &amax);
        if (infequ == 0) {
            *equed = lb::laqge(c.s, N, N, dA, la, R, Cs, rowcnd, colcnd, amax);
            rowequ = (*equed == 'R' || *equed == 'B');
            colequ = (*equed == 'C' || *equed == 'B');
        }
    }
    if (!rowequ && !colequ) set_out(dA, false);                             // A was not scaled
    bool b_changed = false;
    if (notran) { if (rowequ) { lb::scale_rows(c.s, N, *nrhs, dB, lbb, R); b_changed = true; } }        // dgesvx.f:471-487
    else if (colequ) { lb::scale_rows(c.s, N, *nrhs, dB, lbb, Cs); b_changed = true; }
    if (!b_changed) set_out(dB, false);
    double rpvgrw;
    if (nofact || equil) {                                                  // dgesvx.f:489-513
        if (N > 0) lb::lacpy(c.s, 'A', N, N, dA, la, dAF, laf);
        int* dinfo = c.dev_info();
        if (N > 0) lb::getrf(c.s, N, N, dAF, laf, dp, dinfo);
        int hinfo = 0;
        LB_CUDA_CHECK(cudaMemcpyAsync(&hinfo, dinfo, sizeof(int), cudaMemcpyDeviceToHost, c.s));
        LB_CUDA_CHECK(cudaStreamSynchronize(c.s));
        lb::ws_free(c.s, dinfo);
        if (hinfo > 0) {
            rpvgrw = lb::lantr_max_upper(c.s, hinfo, hinfo, dAF, laf);
            if (rpvgrw == 0.0) rpvgrw = 1.0;
            else rpvgrw = lb::lange(c.s, 'M', N, hinfo, dA, la) / rpvgrw;
            if (ptr_kind(work) != PK_DEVICE) work[0] = rpvgrw;
            *rcond = 0.0;
            set_out(dX, false);
            int r = c.finish();
            *info = r ? r : hinfo;
            return;
        }
    }
    const char norm = notran ? '1' : 'I';
    const double anorm = lb::lange(c.s, norm, N, N, dA, la);
    rpvgrw = lb::lantr_max_upper(c.s, N, N, dAF, laf);
    if (rpvgrw == 0.0) rpvgrw = 1.0;
    else rpvgrw = lb::lange(c.s, 'M', N, N, dA, la) / rpvgrw;
    // DGECON (its own quick returns, dgecon.f:195-209)
    *rcond = 0.0;
    if (N == 0) *rcond = 1.0;
    else if (anorm == 0.0) {}
    else if (anorm != anorm) { *rcond = anorm; *info = -5; }
    else if (anorm > DBL_MAX) *info = -5;
    else *info = lb::gecon(c.s, notran, N, dAF, laf, anorm, rcond);
    if (N > 0 && *nrhs > 0) {
        lb::lacpy(c.s, 'A', N, *nrhs, dB, lbb, dX, lx);
        lb::getrs(c.s, notran ? 'N' : 'T', N, *nrhs, dAF, laf, dp, dX, lx);
        gerfs_device(c.s, notran, N, *nrhs, dA, la, dAF, laf, dp, dB, lbb, dX, lx, ferr, berr);
    } else for (int j = 0; j < *nrhs; ++j) { ferr[j] = 0.0; berr[j] = 0.0; }
    if (notran) {                                                           // dgesvx.f:556-581
        if (colequ) { lb::scale_rows(c.s, N, *nrhs, dX, lx, Cs); for (int j = 0; j < *nrhs; ++j) ferr[j] /= colcnd; }
    } else if (rowequ) { lb::scale_rows(c.s, N, *nrhs, dX, lx, R); for (int j = 0; j < *nrhs; ++j) ferr[j] /= rowcnd; }
    if (ptr_kind(work) != PK_DEVICE) work[0] = rpvgrw;
    int r = c.finish();
    *info = r;
    if (r == 0 && *rcond < 1.1102230246251565e-16) *info = N + 1;           // dgesvx.f:587-588
}

}  // extern "C"
