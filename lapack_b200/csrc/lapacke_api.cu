// lapacke_api.cu -- LAPACKE C entry points for the hot path (declared in include/lapack_b200_lapacke.h).
//
// Drop-in for LAPACKE/src/lapacke_d{getrf,getrf2,getrs,gesv,potrf,potrf2,potrs,posv,geqrf,geqr2,larfb,larft,
// laswp}{,_work}.c: same names, argument lists (LAPACKE/include/lapacke.h:807-3143, 5675-8720), return
// codes and error behaviour: info<0 from the Fortran layer is shifted by one for the extra matrix_layout
// argument (lapacke_dgetrf_work.c:42-44); row-major leading-dimension errors and bad layouts call
// LAPACKE_xerbla (LAPACKE/utils/lapacke_xerbla.c:36-45); the optional NaN pre-check (LAPACKE_NANCHECK
// environment variable, lapacke_nancheck.c:99-113) returns -(argument position).
//
// Difference in mechanism, not behaviour: the reference transposes row-major inputs on the host into a
// malloc'ed column-major copy; here the row-major block is uploaded as it is and transposed on the GPU
// (coalesced 32x32 shared-memory tiles), then the device-resident Fortran-ABI routine runs on it.
#include "lb_internal.h"
#include "../../include/lapack_b200_f77.h"
#include "../../include/lapack_b200_lapacke.h"
#include <cstring>
#include <thread>
#include <vector>
#include <atomic>

namespace {

inline int imax(int a, int b) { return a > b ? a : b; }
inline int imin(int a, int b) { return a < b ? a : b; }
inline bool lsame(char a, char b) {
    if (a >= 'a' && a <= 'z') a = (char)(a - 32);
    if (b >= 'a' && b <= 'z') b = (char)(b - 32);
    return a == b;
}
bool is_dev(const void* p) {
    cudaPointerAttributes at;
    if (!p || cudaPointerGetAttributes(&at, p) != cudaSuccess) { (void)cudaGetLastError(); return false; }
    return at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged;
}

// LAPACKE/src/lapacke_nancheck.c:40-113
int g_nancheck = -1;
int get_nancheck() {
    if (g_nancheck != -1) return g_nancheck;
    const char* e = getenv("LAPACKE_NANCHECK");
    g_nancheck = (e == nullptr) ? 1 : (atoi(e) ? 1 : 0);
    return g_nancheck;
}

__global__ void nan_scan_kernel(long long rows, long long cols, const double* a, long long ld, int tri, int* flag) {
    // tri: 0 full, 1 lower incl diag, 2 upper incl diag, 3 strictly-lower + everything below row `cols`
    for (long long j = blockIdx.x; j < cols; j += gridDim.x)
        for (long long i = threadIdx.x; i < rows; i += blockDim.x) {
            if (tri == 1 && i < j) continue;
            if (tri == 2 && i > j) continue;
            if (tri == 3 && i <= j) continue;
            double v = a[i + j * ld];
            if (v != v) *flag = 1;
        }
}

// generic strided NaN scan of a `rows x cols` column-major view (host: multi-threaded; device: kernel)
bool nan_scan(const double* a, long long rows, long long cols, long long ld, int tri) {
    if (!a || rows <= 0 || cols <= 0) return false;
    if (is_dev(a)) {
        int* flag = nullptr;
        int h = 0;
        cudaMalloc(&flag, sizeof(int));
        cudaMemset(flag, 0, sizeof(int));
        nan_scan_kernel<<<(unsigned)(cols < 2048 ? cols : 2048), 256>>>(rows, cols, a, ld, tri, flag);
        cudaMemcpy(&h, flag, sizeof(int), cudaMemcpyDeviceToHost);
        cudaFree(flag);
        return h != 0;
    }
    std::atomic<int> found{0};
    unsigned nt = std::thread::hardware_concurrency();
    if (nt == 0) nt = 1;
    if (nt > 16) nt = 16;
    if (rows * cols < (1 << 20)) nt = 1;
    auto work = [&](long long j0, long long j1) {
        for (long long j = j0; j < j1 && !found.load(std::memory_order_relaxed); ++j) {
            long long i0 = 0, i1 = rows;
            if (tri == 1) i0 = j;
            if (tri == 2) i1 = (j + 1 < rows) ? j + 1 : rows;
            if (tri == 3) i0 = j + 1;
            const double* col = a + j * ld;
            for (long long i = i0; i < i1; ++i)
                if (col[i] != col[i]) { found.store(1); return; }
        }
    };
    if (nt == 1) work(0, cols);
    else {
        std::vector<std::thread> th;
        long long chunk = (cols + nt - 1) / nt;
        for (unsigned t = 0; t < nt; ++t) {
            long long j0 = t * chunk, j1 = (j0 + chunk < cols) ? j0 + chunk : cols;
            if (j0 < j1) th.emplace_back(work, j0, j1);
        }
        for (auto& t : th) t.join();
    }
    return found.load() != 0;
}
// LAPACKE_dge_nancheck: row-major m x n with lda == column-major n x m with ld = lda
bool dge_nan(int layout, int m, int n, const double* a, int lda) {
    if (layout == LAPACK_COL_MAJOR) return nan_scan(a, m, n, lda, 0);
    return nan_scan(a, n, m, lda, 0);
}
// LAPACKE_dpo_nancheck: only the UPLO triangle (row-major lower == column-major upper of the transpose)
bool dpo_nan(int layout, char uplo, int n, const double* a, int lda) {
    bool lower = lsame(uplo, 'L');
    if (layout == LAPACK_ROW_MAJOR) lower = !lower;
    return nan_scan(a, n, n, lda, lower ? 1 : 2);
}

void lapacke_xerbla(const char* name, int info) { LAPACKE_xerbla(name, info); }

// A row-major (rows x cols, row stride ld) host/device matrix staged as a DEVICE column-major copy.
struct RowMajor {
    double* dev_rm = nullptr;   // device image of the row-major block == column-major cols x rows, ld = ldr
    double* dev_cm = nullptr;   // column-major rows x cols, ld = ldc
    long long ldr = 0, ldc = 0;
    int rows = 0, cols = 0;
    double* user = nullptr;
    int user_ld = 0;
    cudaStream_t s = nullptr;

    bool in(const double* a, int rows_, int cols_, int ld) {
        rows = rows_; cols = cols_; user = const_cast<double*>(a); user_ld = ld;
        if (rows <= 0 || cols <= 0) return true;
        ldr = ((long long)cols + 1) & ~1LL;
        ldc = ((long long)rows + 1) & ~1LL;
        if (cudaMalloc(&dev_rm, sizeof(double) * ldr * rows) != cudaSuccess) return false;
        if (cudaMalloc(&dev_cm, sizeof(double) * ldc * cols) != cudaSuccess) { cudaFree(dev_rm); dev_rm = nullptr; return false; }
        cudaMemcpy2DAsync(dev_rm, ldr * 8, a, (size_t)ld * 8, (size_t)cols * 8, rows, cudaMemcpyDefault, s);
        lb::transpose(s, cols, rows, dev_rm, ldr, dev_cm, ldc);
        cudaStreamSynchronize(s);
        return true;
    }
    void out() {
        if (rows <= 0 || cols <= 0 || !dev_cm) return;
        lb::transpose(s, rows, cols, dev_cm, ldc, dev_rm, ldr);
        cudaMemcpy2DAsync(user, (size_t)user_ld * 8, dev_rm, ldr * 8, (size_t)cols * 8, rows, cudaMemcpyDefault, s);
        cudaStreamSynchronize(s);
    }
    ~RowMajor() { if (dev_rm) cudaFree(dev_rm); if (dev_cm) cudaFree(dev_cm); }
};

}  // namespace

extern "C" {

// LAPACKE/utils/lapacke_xerbla.c:36-45
void LAPACKE_xerbla(const char* name, lapack_int info) {
    if (info == LAPACK_WORK_MEMORY_ERROR) printf("Not enough memory to allocate work array in %s\n", name);
    else if (info == LAPACK_TRANSPOSE_MEMORY_ERROR) printf("Not enough memory to transpose matrix in %s\n", name);
    else if (info < 0) printf("Wrong parameter %d in %s\n", -(int)info, name);
}
void LAPACKE_set_nancheck(int flag) { g_nancheck = flag ? 1 : 0; }
int LAPACKE_get_nancheck(void) { return get_nancheck(); }

#define LB_LAYOUT_OK(l) ((l) == LAPACK_COL_MAJOR || (l) == LAPACK_ROW_MAJOR)
#define LB_ADJ(info) do { if ((info) < 0) (info) = (info)-1; } while (0)

// ------------------------------------------------------------------------------------------------ dgetrf / dgetrf2
static lapack_int getrf_work(const char* nm, bool rec, int layout, lapack_int m, lapack_int n, double* a, lapack_int lda,
                             lapack_int* ipiv) {
    lapack_int info = 0;
    if (layout == LAPACK_COL_MAJOR) {
        if (rec) dgetrf2_(&m, &n, a, &lda, ipiv, &info); else dgetrf_(&m, &n, a, &lda, ipiv, &info);
        LB_ADJ(info);
    } else if (layout == LAPACK_ROW_MAJOR) {
        if (lda < n) { info = -5; lapacke_xerbla(nm, info); return info; }
        RowMajor r;
        if (!r.in(a, m, n, lda)) { info = LAPACK_TRANSPOSE_MEMORY_ERROR; lapacke_xerbla(nm, info); return info; }
        lapack_int lda_t = (lapack_int)imax(1, (int)r.ldc);
        double* at = r.dev_cm ? r.dev_cm : a;
        if (rec) dgetrf2_(&m, &n, at, &lda_t, ipiv, &info); else dgetrf_(&m, &n, at, &lda_t, ipiv, &info);
        LB_ADJ(info);
        r.out();
    } else { info = -1; lapacke_xerbla(nm, info); }
    return info;
}
lapack_int LAPACKE_dgetrf_work(int layout, lapack_int m, lapack_int n, double* a, lapack_int lda, lapack_int* ipiv) {
    return getrf_work("LAPACKE_dgetrf_work", false, layout, m, n, a, lda, ipiv);
}
lapack_int LAPACKE_dgetrf2_work(int layout, lapack_int m, lapack_int n, double* a, lapack_int lda, lapack_int* ipiv) {
    return getrf_work("LAPACKE_dgetrf2_work", true, layout, m, n, a, lda, ipiv);
}
lapack_int LAPACKE_dgetrf(int layout, lapack_int m, lapack_int n, double* a, lapack_int lda, lapack_int* ipiv) {
    if (!LB_LAYOUT_OK(layout)) { lapacke_xerbla("LAPACKE_dgetrf", -1); return -1; }
    if (get_nancheck() && dge_nan(layout, m, n, a, lda)) return -4;
    return LAPACKE_dgetrf_work(layout, m, n, a, lda, ipiv);
}
lapack_int LAPACKE_dgetrf2(int layout, lapack_int m, lapack_int n, double* a, lapack_int lda, lapack_int* ipiv) {
    if (!LB_LAYOUT_OK(layout)) { lapacke_xerbla("LAPACKE_dgetrf2", -1); return -1; }
    if (get_nancheck() && dge_nan(layout, m, n, a, lda)) return -4;
    return LAPACKE_dgetrf2_work(layout, m, n, a, lda, ipiv);
}

// ------------------------------------------------------------------------------------------------ dgetrs / dgesv
lapack_int LAPACKE_dgetrs_work(int layout, char trans, lapack_int n, lapack_int nrhs, const double* a, lapack_int lda,
                               const lapack_int* ipiv, double* b, lapack_int ldb) {
    lapack_int info = 0;
    if (layout == LAPACK_COL_MAJOR) {
        dgetrs_(&trans, &n, &nrhs, a, &lda, ipiv, b, &ldb, &info, 1);
        LB_ADJ(info);
    } else if (layout == LAPACK_ROW_MAJOR) {
        if (lda < n) { info = -6; lapacke_xerbla("LAPACKE_dgetrs_work", info); return info; }
        if (ldb < nrhs) { info = -9; lapacke_xerbla("LAPACKE_dgetrs_work", info); return info; }
        RowMajor ra, rb;
        if (!ra.in(a, n, n, lda) || !rb.in(b, n, nrhs, ldb)) { info = LAPACK_TRANSPOSE_MEMORY_ERROR; lapacke_xerbla("LAPACKE_dgetrs_work", info); return info; }
        lapack_int lda_t = imax(1, (int)ra.ldc), ldb_t = imax(1, (int)rb.ldc);
        dgetrs_(&trans, &n, &nrhs, ra.dev_cm ? ra.dev_cm : a, &lda_t, ipiv, rb.dev_cm ? rb.dev_cm : b, &ldb_t, &info, 1);
        LB_ADJ(info);
        rb.out();
    } else { info = -1; lapacke_xerbla("LAPACKE_dgetrs_work", info); }
    return info;
}
lapack_int LAPACKE_dgetrs(int layout, char trans, lapack_int n, lapack_int nrhs, const double* a, lapack_int lda,
                          const lapack_int* ipiv, double* b, lapack_int ldb) {
    if (!LB_LAYOUT_OK(layout)) { lapacke_xerbla("LAPACKE_dgetrs", -1); return -1; }
    if (get_nancheck()) {
        if (dge_nan(layout, n, n, a, lda)) return -5;
        if (dge_nan(layout, n, nrhs, b, ldb)) return -8;
    }
    return LAPACKE_dgetrs_work(layout, trans, n, nrhs, a, lda, ipiv, b, ldb);
}
lapack_int LAPACKE_dgesv_work(int layout, lapack_int n, lapack_int nrhs, double* a, lapack_int lda, lapack_int* ipiv,
                              double* b, lapack_int ldb) {
    lapack_int info = 0;
    if (layout == LAPACK_COL_MAJOR) {
        dgesv_(&n, &nrhs, a, &lda, ipiv, b, &ldb, &info);
        LB_ADJ(info);
    } else if (layout == LAPACK_ROW_MAJOR) {
        if (lda < n) { info = -5; lapacke_xerbla("LAPACKE_dgesv_work", info); return info; }
        if (ldb < nrhs) { info = -8; lapacke_xerbla("LAPACKE_dgesv_work", info); return info; }
        RowMajor ra, rb;
        if (!ra.in(a, n, n, lda) || !rb.in(b, n, nrhs, ldb)) { info = LAPACK_TRANSPOSE_MEMORY_ERROR; lapacke_xerbla("LAPACKE_dgesv_work", info); return info; }
        lapack_int lda_t = imax(1, (int)ra.ldc), ldb_t = imax(1, (int)rb.ldc);
        dgesv_(&n, &nrhs, ra.dev_cm ? ra.dev_cm : a, &lda_t, ipiv, rb.dev_cm ? rb.dev_cm : b, &ldb_t, &info);
        LB_ADJ(info);
        ra.out();
        rb.out();
    } else { info = -1; lapacke_xerbla("LAPACKE_dgesv_work", info); }
    return info;
}
lapack_int LAPACKE_dgesv(int layout, lapack_int n, lapack_int nrhs, double* a, lapack_int lda, lapack_int* ipiv, double* b,
                         lapack_int ldb) {
    if (!LB_LAYOUT_OK(layout)) { lapacke_xerbla("LAPACKE_dgesv", -1); return -1; }
    if (get_nancheck()) {
        if (dge_nan(layout, n, n, a, lda)) return -4;
        if (dge_nan(layout, n, nrhs, b, ldb)) return -7;
    }
    return LAPACKE_dgesv_work(layout, n, nrhs, a, lda, ipiv, b, ldb);
}

// ------------------------------------------------------------------------------------------------ dpotrf / dpotrf2 / dpotrs / dposv
static lapack_int potrf_work(const char* nm, bool rec, int layout, char uplo, lapack_int n, double* a, lapack_int lda) {
    lapack_int info = 0;
    if (layout == LAPACK_COL_MAJOR) {
        if (rec) dpotrf2_(&uplo, &n, a, &lda, &info, 1); else dpotrf_(&uplo, &n, a, &lda, &info, 1);
        LB_ADJ(info);
    } else if (layout == LAPACK_ROW_MAJOR) {
        if (lda < n) { info = -5; lapacke_xerbla(nm, info); return info; }
        // A row-major UPLO triangle is the opposite triangle of the same buffer read column-major, and
        // A = L L^T (row-major lower) <=> buffer-as-column-major = U^T U: no transposition needed at all.
        char u2 = lsame(uplo, 'L') ? 'U' : (lsame(uplo, 'U') ? 'L' : uplo);
        if (rec) dpotrf2_(&u2, &n, a, &lda, &info, 1); else dpotrf_(&u2, &n, a, &lda, &info, 1);
        LB_ADJ(info);
    } else { info = -1; lapacke_xerbla(nm, info); }
    return info;
}
lapack_int LAPACKE_dpotrf_work(int layout, char uplo, lapack_int n, double* a, lapack_int lda) {
    return potrf_work("LAPACKE_dpotrf_work", false, layout, uplo, n, a, lda);
}
lapack_int LAPACKE_dpotrf2_work(int layout, char uplo, lapack_int n, double* a, lapack_int lda) {
    return potrf_work("LAPACKE_dpotrf2_work", true, layout, uplo, n, a, lda);
}
lapack_int LAPACKE_dpotrf(int layout, char uplo, lapack_int n, double* a, lapack_int lda) {
    if (!LB_LAYOUT_OK(layout)) { lapacke_xerbla("LAPACKE_dpotrf", -1); return -1; }
    if (get_nancheck() && dpo_nan(layout, uplo, n, a, lda)) return -4;
    return LAPACKE_dpotrf_work(layout, uplo, n, a, lda);
}
lapack_int LAPACKE_dpotrf2(int layout, char uplo, lapack_int n, double* a, lapack_int lda) {
    if (!LB_LAYOUT_OK(layout)) { lapacke_xerbla("LAPACKE_dpotrf2", -1); return -1; }
    if (get_nancheck() && dpo_nan(layout, uplo, n, a, lda)) return -4;
    return LAPACKE_dpotrf2_work(layout, uplo, n, a, lda);
}
lapack_int LAPACKE_dpotrs_work(int layout, char uplo, lapack_int n, lapack_int nrhs, const double* a, lapack_int lda,
                               double* b, lapack_int ldb) {
    lapack_int info = 0;
    if (layout == LAPACK_COL_MAJOR) {
        dpotrs_(&uplo, &n, &nrhs, a, &lda, b, &ldb, &info, 1);
        LB_ADJ(info);
    } else if (layout == LAPACK_ROW_MAJOR) {
        if (lda < n) { info = -6; lapacke_xerbla("LAPACKE_dpotrs_work", info); return info; }
        if (ldb < nrhs) { info = -8; lapacke_xerbla("LAPACKE_dpotrs_work", info); return info; }
        RowMajor rb;
        if (!rb.in(b, n, nrhs, ldb)) { info = LAPACK_TRANSPOSE_MEMORY_ERROR; lapacke_xerbla("LAPACKE_dpotrs_work", info); return info; }
        char u2 = lsame(uplo, 'L') ? 'U' : (lsame(uplo, 'U') ? 'L' : uplo);
        lapack_int ldb_t = imax(1, (int)rb.ldc);
        dpotrs_(&u2, &n, &nrhs, a, &lda, rb.dev_cm ? rb.dev_cm : b, &ldb_t, &info, 1);
        LB_ADJ(info);
        rb.out();
    } else { info = -1; lapacke_xerbla("LAPACKE_dpotrs_work", info); }
    return info;
}
lapack_int LAPACKE_dpotrs(int layout, char uplo, lapack_int n, lapack_int nrhs, const double* a, lapack_int lda, double* b,
                          lapack_int ldb) {
    if (!LB_LAYOUT_OK(layout)) { lapacke_xerbla("LAPACKE_dpotrs", -1); return -1; }
    if (get_nancheck()) {
        if (dpo_nan(layout, uplo, n, a, lda)) return -5;
        if (dge_nan(layout, n, nrhs, b, ldb)) return -7;
    }
    return LAPACKE_dpotrs_work(layout, uplo, n, nrhs, a, lda, b, ldb);
}
lapack_int LAPACKE_dposv_work(int layout, char uplo, lapack_int n, lapack_int nrhs, double* a, lapack_int lda, double* b,
                              lapack_int ldb) {
    lapack_int info = 0;
    if (layout == LAPACK_COL_MAJOR) {
        dposv_(&uplo, &n, &nrhs, a, &lda, b, &ldb, &info, 1);
        LB_ADJ(info);
    } else if (layout == LAPACK_ROW_MAJOR) {
        if (lda < n) { info = -6; lapacke_xerbla("LAPACKE_dposv_work", info); return info; }
        if (ldb < nrhs) { info = -8; lapacke_xerbla("LAPACKE_dposv_work", info); return info; }
        RowMajor rb;
        if (!rb.in(b, n, nrhs, ldb)) { info = LAPACK_TRANSPOSE_MEMORY_ERROR; lapacke_xerbla("LAPACKE_dposv_work", info); return info; }
        char u2 = lsame(uplo, 'L') ? 'U' : (lsame(uplo, 'U') ? 'L' : uplo);
        lapack_int ldb_t = imax(1, (int)rb.ldc);
        dposv_(&u2, &n, &nrhs, a, &lda, rb.dev_cm ? rb.dev_cm : b, &ldb_t, &info, 1);
        LB_ADJ(info);
        rb.out();
    } else { info = -1; lapacke_xerbla("LAPACKE_dposv_work", info); }
    return info;
}
lapack_int LAPACKE_dposv(int layout, char uplo, lapack_int n, lapack_int nrhs, double* a, lapack_int lda, double* b,
                         lapack_int ldb) {
    if (!LB_LAYOUT_OK(layout)) { lapacke_xerbla("LAPACKE_dposv", -1); return -1; }
    if (get_nancheck()) {
        if (dpo_nan(layout, uplo, n, a, lda)) return -5;
        if (dge_nan(layout, n, nrhs, b, ldb)) return -7;
    }
    return LAPACKE_dposv_work(layout, uplo, n, nrhs, a, lda, b, ldb);
}

// ------------------------------------------------------------------------------------------------ dgeqrf / dgeqr2
lapack_int LAPACKE_dgeqrf_work(int layout, lapack_int m, lapack_int n, double* a, lapack_int lda, double* tau, double* work,
                               lapack_int lwork) {
    lapack_int info = 0;
    if (layout == LAPACK_COL_MAJOR) {
        dgeqrf_(&m, &n, a, &lda, tau, work, &lwork, &info);
        LB_ADJ(info);
    } else if (layout == LAPACK_ROW_MAJOR) {
        if (lda < n) { info = -5; lapacke_xerbla("LAPACKE_dgeqrf_work", info); return info; }
        lapack_int lda_t = imax(1, m);
        if (lwork == -1) {   // workspace query (lapacke_dgeqrf_work.c:55-59)
            dgeqrf_(&m, &n, a, &lda_t, tau, work, &lwork, &info);
            LB_ADJ(info);
            return info;
        }
        RowMajor r;
        if (!r.in(a, m, n, lda)) { info = LAPACK_TRANSPOSE_MEMORY_ERROR; lapacke_xerbla("LAPACKE_dgeqrf_work", info); return info; }
        lda_t = imax(1, (int)r.ldc);
        dgeqrf_(&m, &n, r.dev_cm ? r.dev_cm : a, &lda_t, tau, work, &lwork, &info);
        LB_ADJ(info);
        r.out();
    } else { info = -1; lapacke_xerbla("LAPACKE_dgeqrf_work", info); }
    return info;
}
lapack_int LAPACKE_dgeqrf(int layout, lapack_int m, lapack_int n, double* a, lapack_int lda, double* tau) {
    if (!LB_LAYOUT_OK(layout)) { lapacke_xerbla("LAPACKE_dgeqrf", -1); return -1; }
    if (get_nancheck() && dge_nan(layout, m, n, a, lda)) return -4;
    double wq = 0.0;
    lapack_int info = LAPACKE_dgeqrf_work(layout, m, n, a, lda, tau, &wq, -1);
    if (info != 0) return info;
    lapack_int lwork = (lapack_int)wq;
    double* work = (double*)malloc(sizeof(double) * (size_t)imax(1, lwork));
    if (!work) { lapacke_xerbla("LAPACKE_dgeqrf", LAPACK_WORK_MEMORY_ERROR); return LAPACK_WORK_MEMORY_ERROR; }
    info = LAPACKE_dgeqrf_work(layout, m, n, a, lda, tau, work, lwork);
    free(work);
    return info;
}
lapack_int LAPACKE_dgeqr2_work(int layout, lapack_int m, lapack_int n, double* a, lapack_int lda, double* tau, double* work) {
    lapack_int info = 0;
    if (layout == LAPACK_COL_MAJOR) {
        dgeqr2_(&m, &n, a, &lda, tau, work, &info);
        LB_ADJ(info);
    } else if (layout == LAPACK_ROW_MAJOR) {
        if (lda < n) { info = -5; lapacke_xerbla("LAPACKE_dgeqr2_work", info); return info; }
        RowMajor r;
        if (!r.in(a, m, n, lda)) { info = LAPACK_TRANSPOSE_MEMORY_ERROR; lapacke_xerbla("LAPACKE_dgeqr2_work", info); return info; }
        lapack_int lda_t = imax(1, (int)r.ldc);
        dgeqr2_(&m, &n, r.dev_cm ? r.dev_cm : a, &lda_t, tau, work, &info);
        LB_ADJ(info);
        r.out();
    } else { info = -1; lapacke_xerbla("LAPACKE_dgeqr2_work", info); }
    return info;
}
lapack_int LAPACKE_dgeqr2(int layout, lapack_int m, lapack_int n, double* a, lapack_int lda, double* tau) {
    if (!LB_LAYOUT_OK(layout)) { lapacke_xerbla("LAPACKE_dgeqr2", -1); return -1; }
    if (get_nancheck() && dge_nan(layout, m, n, a, lda)) return -4;
    double* work = (double*)malloc(sizeof(double) * (size_t)imax(1, n));
    if (!work) { lapacke_xerbla("LAPACKE_dgeqr2", LAPACK_WORK_MEMORY_ERROR); return LAPACK_WORK_MEMORY_ERROR; }
    lapack_int info = LAPACKE_dgeqr2_work(layout, m, n, a, lda, tau, work);
    free(work);
    return info;
}

// ------------------------------------------------------------------------------------------------ dgetri
// LAPACKE/src/lapacke_dgetri_work.c:40-88, lapacke_dgetri.c:36-75
lapack_int LAPACKE_dgetri_work(int layout, lapack_int n, double* a, lapack_int lda, const lapack_int* ipiv, double* work,
                               lapack_int lwork) {
    lapack_int info = 0;
    if (layout == LAPACK_COL_MAJOR) {
        dgetri_(&n, a, &lda, ipiv, work, &lwork, &info);
        LB_ADJ(info);
    } else if (layout == LAPACK_ROW_MAJOR) {
        lapack_int lda_t = imax(1, n);
        if (lda < n) { info = -4; lapacke_xerbla("LAPACKE_dgetri_work", info); return info; }
        if (lwork == -1) {
            dgetri_(&n, a, &lda_t, ipiv, work, &lwork, &info);
            LB_ADJ(info);
            return info;
        }
        RowMajor r;
        if (!r.in(a, n, n, lda)) { info = LAPACK_TRANSPOSE_MEMORY_ERROR; lapacke_xerbla("LAPACKE_dgetri_work", info); return info; }
        lda_t = imax(1, (int)r.ldc);
        dgetri_(&n, r.dev_cm ? r.dev_cm : a, &lda_t, ipiv, work, &lwork, &info);
        LB_ADJ(info);
        r.out();
    } else { info = -1; lapacke_xerbla("LAPACKE_dgetri_work", info); }
    return info;
}
lapack_int LAPACKE_dgetri(int layout, lapack_int n, double* a, lapack_int lda, const lapack_int* ipiv) {
    if (!LB_LAYOUT_OK(layout)) { lapacke_xerbla("LAPACKE_dgetri", -1); return -1; }
    if (get_nancheck() && dge_nan(layout, n, n, a, lda)) return -3;
    double wq = 0.0;
    lapack_int info = LAPACKE_dgetri_work(layout, n, a, lda, ipiv, &wq, -1);
    if (info != 0) return info;
    lapack_int lwork = (lapack_int)wq;
    double* work = (double*)malloc(sizeof(double) * (size_t)imax(1, lwork));
    if (!work) { lapacke_xerbla("LAPACKE_dgetri", LAPACK_WORK_MEMORY_ERROR); return LAPACK_WORK_MEMORY_ERROR; }
    info = LAPACKE_dgetri_work(layout, n, a, lda, ipiv, work, lwork);
    free(work);
    return info;
}

// ------------------------------------------------------------------------------------------------ dgelqf / dormlq
// LAPACKE/src/lapacke_dgelqf_work.c, lapacke_dgelqf.c, lapacke_dormlq_work.c, lapacke_dormlq.c (same structure as the QR pair)
lapack_int LAPACKE_dgelqf_work(int layout, lapack_int m, lapack_int n, double* a, lapack_int lda, double* tau, double* work,
                               lapack_int lwork) {
    lapack_int info = 0;
    if (layout == LAPACK_COL_MAJOR) {
        dgelqf_(&m, &n, a, &lda, tau, work, &lwork, &info);
        LB_ADJ(info);
    } else if (layout == LAPACK_ROW_MAJOR) {
        if (lda < n) { info = -5; lapacke_xerbla("LAPACKE_dgelqf_work", info); return info; }
        lapack_int lda_t = imax(1, m);
        if (lwork == -1) {
            dgelqf_(&m, &n, a, &lda_t, tau, work, &lwork, &info);
            LB_ADJ(info);
            return info;
        }
        RowMajor r;
        if (!r.in(a, m, n, lda)) { info = LAPACK_TRANSPOSE_MEMORY_ERROR; lapacke_xerbla("LAPACKE_dgelqf_work", info); return info; }
        lda_t = imax(1, (int)r.ldc);
        dgelqf_(&m, &n, r.dev_cm ? r.dev_cm : a, &lda_t, tau, work, &lwork, &info);
        LB_ADJ(info);
        r.out();
    } else { info = -1; lapacke_xerbla("LAPACKE_dgelqf_work", info); }
    return info;
}
lapack_int LAPACKE_dgelqf(int layout, lapack_int m, lapack_int n, double* a, lapack_int lda, double* tau) {
    if (!LB_LAYOUT_OK(layout)) { lapacke_xerbla("LAPACKE_dgelqf", -1); return -1; }
    if (get_nancheck() && dge_nan(layout, m, n, a, lda)) return -4;
    double wq = 0.0;
    lapack_int info = LAPACKE_dgelqf_work(layout, m, n, a, lda, tau, &wq, -1);
    if (info != 0) return info;
    lapack_int lwork = (lapack_int)wq;
    double* work = (double*)malloc(sizeof(double) * (size_t)imax(1, lwork));
    if (!work) { lapacke_xerbla("LAPACKE_dgelqf", LAPACK_WORK_MEMORY_ERROR); return LAPACK_WORK_MEMORY_ERROR; }
    info = LAPACKE_dgelqf_work(layout, m, n, a, lda, tau, work, lwork);
    free(work);
    return info;
}
lapack_int LAPACKE_dormlq_work(int layout, char side, char trans, lapack_int m, lapack_int n, lapack_int k, const double* a,
                               lapack_int lda, const double* tau, double* c, lapack_int ldc, double* work, lapack_int lwork) {
    lapack_int info = 0;
    if (layout == LAPACK_COL_MAJOR) {
        dormlq_(&side, &trans, &m, &n, &k, a, &lda, tau, c, &ldc, work, &lwork, &info, 1, 1);
        LB_ADJ(info);
    } else if (layout == LAPACK_ROW_MAJOR) {
        const lapack_int r = lsame(side, 'l') ? m : n;
        lapack_int lda_t = imax(1, k), ldc_t = imax(1, m);
        if (lda < r) { info = -8; lapacke_xerbla("LAPACKE_dormlq_work", info); return info; }
        if (ldc < n) { info = -11; lapacke_xerbla("LAPACKE_dormlq_work", info); return info; }
        if (lwork == -1) {
            dormlq_(&side, &trans, &m, &n, &k, a, &lda_t, tau, c, &ldc_t, work, &lwork, &info, 1, 1);
            LB_ADJ(info);
            return info;
        }
        RowMajor ra, rc;
        if (!ra.in(a, k, r, lda) || !rc.in(c, m, n, ldc)) {
            info = LAPACK_TRANSPOSE_MEMORY_ERROR; lapacke_xerbla("LAPACKE_dormlq_work", info); return info;
        }
        lda_t = imax(1, (int)ra.ldc); ldc_t = imax(1, (int)rc.ldc);
        dormlq_(&side, &trans, &m, &n, &k, ra.dev_cm ? ra.dev_cm : a, &lda_t, tau, rc.dev_cm ? rc.dev_cm : c, &ldc_t, work, &lwork,
                &info, 1, 1);
        LB_ADJ(info);
        rc.out();
    } else { info = -1; lapacke_xerbla("LAPACKE_dormlq_work", info); }
    return info;
}
lapack_int LAPACKE_dormlq(int layout, char side, char trans, lapack_int m, lapack_int n, lapack_int k, const double* a,
                          lapack_int lda, const double* tau, double* c, lapack_int ldc) {
    if (!LB_LAYOUT_OK(layout)) { lapacke_xerbla("LAPACKE_dormlq", -1); return -1; }
    if (get_nancheck()) {
        const lapack_int r = lsame(side, 'l') ? m : n;
        if (dge_nan(layout, k, r, a, lda)) return -7;
        if (dge_nan(layout, m, n, c, ldc)) return -10;
        if (nan_scan(tau, k, 1, imax(1, k), 0)) return -9;
    }
    double wq = 0.0;
    lapack_int info = LAPACKE_dormlq_work(layout, side, trans, m, n, k, a, lda, tau, c, ldc, &wq, -1);
    if (info != 0) return info;
    lapack_int lwork = (lapack_int)wq;
    double* work = (double*)malloc(sizeof(double) * (size_t)imax(1, lwork));
    if (!work) { lapacke_xerbla("LAPACKE_dormlq", LAPACK_WORK_MEMORY_ERROR); return LAPACK_WORK_MEMORY_ERROR; }
    info = LAPACKE_dormlq_work(layout, side, trans, m, n, k, a, lda, tau, c, ldc, work, lwork);
    free(work);
    return info;
}

// ------------------------------------------------------------------------------------------------ dgels
// LAPACKE/src/lapacke_dgels_work.c:41-105, lapacke_dgels.c:36-82
lapack_int LAPACKE_dgels_work(int layout, char trans, lapack_int m, lapack_int n, lapack_int nrhs, double* a, lapack_int lda,
                              double* b, lapack_int ldb, double* work, lapack_int lwork) {
    lapack_int info = 0;
    if (layout == LAPACK_COL_MAJOR) {
        dgels_(&trans, &m, &n, &nrhs, a, &lda, b, &ldb, work, &lwork, &info, 1);
        LB_ADJ(info);
    } else if (layout == LAPACK_ROW_MAJOR) {
        const lapack_int mx = imax(m, n);
        lapack_int lda_t = imax(1, m), ldb_t = imax(1, mx);
        if (lda < n) { info = -7; lapacke_xerbla("LAPACKE_dgels_work", info); return info; }
        if (ldb < nrhs) { info = -9; lapacke_xerbla("LAPACKE_dgels_work", info); return info; }
        if (lwork == -1) {
            dgels_(&trans, &m, &n, &nrhs, a, &lda_t, b, &ldb_t, work, &lwork, &info, 1);
            LB_ADJ(info);
            return info;
        }
        RowMajor ra, rb;
        if (!ra.in(a, m, n, lda) || !rb.in(b, mx, nrhs, ldb)) {
            info = LAPACK_TRANSPOSE_MEMORY_ERROR; lapacke_xerbla("LAPACKE_dgels_work", info); return info;
        }
        lda_t = imax(1, (int)ra.ldc); ldb_t = imax(1, (int)rb.ldc);
        dgels_(&trans, &m, &n, &nrhs, ra.dev_cm ? ra.dev_cm : a, &lda_t, rb.dev_cm ? rb.dev_cm : b, &ldb_t, work, &lwork, &info, 1);
        LB_ADJ(info);
        ra.out();
        rb.out();
    } else { info = -1; lapacke_xerbla("LAPACKE_dgels_work", info); }
    return info;
}
lapack_int LAPACKE_dgels(int layout, char trans, lapack_int m, lapack_int n, lapack_int nrhs, double* a, lapack_int lda, double* b,
                         lapack_int ldb) {
    if (!LB_LAYOUT_OK(layout)) { lapacke_xerbla("LAPACKE_dgels", -1); return -1; }
    if (get_nancheck()) {
        if (dge_nan(layout, m, n, a, lda)) return -6;
        if (dge_nan(layout, imax(m, n), nrhs, b, ldb)) return -8;
    }
    double wq = 0.0;
    lapack_int info = LAPACKE_dgels_work(layout, trans, m, n, nrhs, a, lda, b, ldb, &wq, -1);
    if (info != 0) return info;
    lapack_int lwork = (lapack_int)wq;
    double* work = (double*)malloc(sizeof(double) * (size_t)imax(1, lwork));
    if (!work) { lapacke_xerbla("LAPACKE_dgels", LAPACK_WORK_MEMORY_ERROR); return LAPACK_WORK_MEMORY_ERROR; }
    info = LAPACKE_dgels_work(layout, trans, m, n, nrhs, a, lda, b, ldb, work, lwork);
    free(work);
    return info;
}

// ------------------------------------------------------------------------------------------------ dgeqrt / dgemqrt
// LAPACKE/src/lapacke_dgeqrt_work.c:40-98, lapacke_dgeqrt.c:36-68, lapacke_dgemqrt_work.c:40-115, lapacke_dgemqrt.c:36-80.
// Row-major operands are transposed on the GPU; row-major T is nb x k with ldt >= k (for DGEMQRT the reference checks
// ldt < nb and transposes an "ldt x nb" array, lapacke_dgemqrt_work.c:62-66,90 -- the consistent layout is used here).
lapack_int LAPACKE_dgeqrt_work(int layout, lapack_int m, lapack_int n, lapack_int nb, double* a, lapack_int lda, double* t,
                               lapack_int ldt, double* work) {
    lapack_int info = 0;
    if (layout == LAPACK_COL_MAJOR) {
        dgeqrt_(&m, &n, &nb, a, &lda, t, &ldt, work, &info);
        LB_ADJ(info);
    } else if (layout == LAPACK_ROW_MAJOR) {
        const lapack_int k = imin(m, n);
        if (lda < n) { info = -6; lapacke_xerbla("LAPACKE_dgeqrt_work", info); return info; }
        if (ldt < k) { info = -8; lapacke_xerbla("LAPACKE_dgeqrt_work", info); return info; }
        RowMajor ra, rt;
        if (!ra.in(a, m, n, lda) || !rt.in(t, nb, k, ldt)) {
            info = LAPACK_TRANSPOSE_MEMORY_ERROR; lapacke_xerbla("LAPACKE_dgeqrt_work", info); return info;
        }
        lapack_int lda_t = imax(1, (int)ra.ldc), ldt_t = imax(1, (int)rt.ldc);
        dgeqrt_(&m, &n, &nb, ra.dev_cm ? ra.dev_cm : a, &lda_t, rt.dev_cm ? rt.dev_cm : t, &ldt_t, work, &info);
        LB_ADJ(info);
        ra.out();
        rt.out();
    } else { info = -1; lapacke_xerbla("LAPACKE_dgeqrt_work", info); }
    return info;
}
lapack_int LAPACKE_dgeqrt(int layout, lapack_int m, lapack_int n, lapack_int nb, double* a, lapack_int lda, double* t,
                          lapack_int ldt) {
    if (!LB_LAYOUT_OK(layout)) { lapacke_xerbla("LAPACKE_dgeqrt", -1); return -1; }
    if (get_nancheck() && dge_nan(layout, m, n, a, lda)) return -5;
    double* work = (double*)malloc(sizeof(double) * (size_t)imax(1, nb) * (size_t)imax(1, n));
    if (!work) { lapacke_xerbla("LAPACKE_dgeqrt", LAPACK_WORK_MEMORY_ERROR); return LAPACK_WORK_MEMORY_ERROR; }
    lapack_int info = LAPACKE_dgeqrt_work(layout, m, n, nb, a, lda, t, ldt, work);
    free(work);
    return info;
}
lapack_int LAPACKE_dgemqrt_work(int layout, char side, char trans, lapack_int m, lapack_int n, lapack_int k, lapack_int nb,
                                const double* v, lapack_int ldv, const double* t, lapack_int ldt, double* c, lapack_int ldc,
                                double* work) {
    lapack_int info = 0;
    if (layout == LAPACK_COL_MAJOR) {
        dgemqrt_(&side, &trans, &m, &n, &k, &nb, v, &ldv, t, &ldt, c, &ldc, work, &info, 1, 1);
        LB_ADJ(info);
    } else if (layout == LAPACK_ROW_MAJOR) {
        const lapack_int q = lsame(side, 'l') ? m : n;
        if (ldc < n) { info = -13; lapacke_xerbla("LAPACKE_dgemqrt_work", info); return info; }
        if (ldt < k) { info = -11; lapacke_xerbla("LAPACKE_dgemqrt_work", info); return info; }
        if (ldv < k) { info = -9; lapacke_xerbla("LAPACKE_dgemqrt_work", info); return info; }
        RowMajor rv, rt, rc;
        if (!rv.in(v, q, k, ldv) || !rt.in(t, nb, k, ldt) || !rc.in(c, m, n, ldc)) {
            info = LAPACK_TRANSPOSE_MEMORY_ERROR; lapacke_xerbla("LAPACKE_dgemqrt_work", info); return info;
        }
        lapack_int ldv_t = imax(1, (int)rv.ldc), ldt_t = imax(1, (int)rt.ldc), ldc_t = imax(1, (int)rc.ldc);
        dgemqrt_(&side, &trans, &m, &n, &k, &nb, rv.dev_cm ? rv.dev_cm : v, &ldv_t, rt.dev_cm ? rt.dev_cm : t, &ldt_t,
                 rc.dev_cm ? rc.dev_cm : c, &ldc_t, work, &info, 1, 1);
        LB_ADJ(info);
        rc.out();
    } else { info = -1; lapacke_xerbla("LAPACKE_dgemqrt_work", info); }
    return info;
}
lapack_int LAPACKE_dgemqrt(int layout, char side, char trans, lapack_int m, lapack_int n, lapack_int k, lapack_int nb,
                           const double* v, lapack_int ldv, const double* t, lapack_int ldt, double* c, lapack_int ldc) {
    if (!LB_LAYOUT_OK(layout)) { lapacke_xerbla("LAPACKE_dgemqrt", -1); return -1; }
    if (get_nancheck()) {
        const lapack_int q = lsame(side, 'l') ? m : (lsame(side, 'r') ? n : 0);
        if (dge_nan(layout, m, n, c, ldc)) return -12;
        if (dge_nan(layout, nb, k, t, ldt)) return -10;
        if (dge_nan(layout, q, k, v, ldv)) return -8;
    }
    double* work = (double*)malloc(sizeof(double) * (size_t)imax(1, m) * (size_t)imax(1, nb));
    if (!work) { lapacke_xerbla("LAPACKE_dgemqrt", LAPACK_WORK_MEMORY_ERROR); return LAPACK_WORK_MEMORY_ERROR; }
    lapack_int info = LAPACKE_dgemqrt_work(layout, side, trans, m, n, k, nb, v, ldv, t, ldt, c, ldc, work);
    free(work);
    return info;
}

// ------------------------------------------------------------------------------------------------ dorgqr / dormqr
// LAPACKE/src/lapacke_dorgqr_work.c:41-88, lapacke_dorgqr.c:36-78, lapacke_dormqr_work.c:41-110, lapacke_dormqr.c:36-90
lapack_int LAPACKE_dorgqr_work(int layout, lapack_int m, lapack_int n, lapack_int k, double* a, lapack_int lda, const double* tau,
                               double* work, lapack_int lwork) {
    lapack_int info = 0;
    if (layout == LAPACK_COL_MAJOR) {
        dorgqr_(&m, &n, &k, a, &lda, tau, work, &lwork, &info);
        LB_ADJ(info);
    } else if (layout == LAPACK_ROW_MAJOR) {
        lapack_int lda_t = imax(1, m);
        if (lda < n) { info = -6; lapacke_xerbla("LAPACKE_dorgqr_work", info); return info; }
        if (lwork == -1) {
            dorgqr_(&m, &n, &k, a, &lda_t, tau, work, &lwork, &info);
            LB_ADJ(info);
            return info;
        }
        RowMajor r;
        if (!r.in(a, m, n, lda)) { info = LAPACK_TRANSPOSE_MEMORY_ERROR; lapacke_xerbla("LAPACKE_dorgqr_work", info); return info; }
        lda_t = imax(1, (int)r.ldc);
        dorgqr_(&m, &n, &k, r.dev_cm ? r.dev_cm : a, &lda_t, tau, work, &lwork, &info);
        LB_ADJ(info);
        r.out();
    } else { info = -1; lapacke_xerbla("LAPACKE_dorgqr_work", info); }
    return info;
}
lapack_int LAPACKE_dorgqr(int layout, lapack_int m, lapack_int n, lapack_int k, double* a, lapack_int lda, const double* tau) {
    if (!LB_LAYOUT_OK(layout)) { lapacke_xerbla("LAPACKE_dorgqr", -1); return -1; }
    if (get_nancheck()) {
        if (dge_nan(layout, m, n, a, lda)) return -5;
        if (nan_scan(tau, k, 1, imax(1, k), 0)) return -7;
    }
    double wq = 0.0;
    lapack_int info = LAPACKE_dorgqr_work(layout, m, n, k, a, lda, tau, &wq, -1);
    if (info != 0) return info;
    lapack_int lwork = (lapack_int)wq;
    double* work = (double*)malloc(sizeof(double) * (size_t)imax(1, lwork));
    if (!work) { lapacke_xerbla("LAPACKE_dorgqr", LAPACK_WORK_MEMORY_ERROR); return LAPACK_WORK_MEMORY_ERROR; }
    info = LAPACKE_dorgqr_work(layout, m, n, k, a, lda, tau, work, lwork);
    free(work);
    return info;
}
lapack_int LAPACKE_dormqr_work(int layout, char side, char trans, lapack_int m, lapack_int n, lapack_int k, const double* a,
                               lapack_int lda, const double* tau, double* c, lapack_int ldc, double* work, lapack_int lwork) {
    lapack_int info = 0;
    if (layout == LAPACK_COL_MAJOR) {
        dormqr_(&side, &trans, &m, &n, &k, a, &lda, tau, c, &ldc, work, &lwork, &info, 1, 1);
        LB_ADJ(info);
    } else if (layout == LAPACK_ROW_MAJOR) {
        const lapack_int r = lsame(side, 'l') ? m : n;
        lapack_int lda_t = imax(1, r), ldc_t = imax(1, m);
        if (lda < k) { info = -8; lapacke_xerbla("LAPACKE_dormqr_work", info); return info; }
        if (ldc < n) { info = -11; lapacke_xerbla("LAPACKE_dormqr_work", info); return info; }
        if (lwork == -1) {
            dormqr_(&side, &trans, &m, &n, &k, a, &lda_t, tau, c, &ldc_t, work, &lwork, &info, 1, 1);
            LB_ADJ(info);
            return info;
        }
        RowMajor ra, rc;
        if (!ra.in(a, r, k, lda) || !rc.in(c, m, n, ldc)) {
            info = LAPACK_TRANSPOSE_MEMORY_ERROR; lapacke_xerbla("LAPACKE_dormqr_work", info); return info;
        }
        lda_t = imax(1, (int)ra.ldc); ldc_t = imax(1, (int)rc.ldc);
        dormqr_(&side, &trans, &m, &n, &k, ra.dev_cm ? ra.dev_cm : a, &lda_t, tau, rc.dev_cm ? rc.dev_cm : c, &ldc_t, work, &lwork,
                &info, 1, 1);
        LB_ADJ(info);
        rc.out();
    } else { info = -1; lapacke_xerbla("LAPACKE_dormqr_work", info); }
    return info;
}
lapack_int LAPACKE_dormqr(int layout, char side, char trans, lapack_int m, lapack_int n, lapack_int k, const double* a,
                          lapack_int lda, const double* tau, double* c, lapack_int ldc) {
    if (!LB_LAYOUT_OK(layout)) { lapacke_xerbla("LAPACKE_dormqr", -1); return -1; }
    if (get_nancheck()) {
        const lapack_int r = lsame(side, 'l') ? m : n;
        if (dge_nan(layout, r, k, a, lda)) return -7;
        if (dge_nan(layout, m, n, c, ldc)) return -10;
        if (nan_scan(tau, k, 1, imax(1, k), 0)) return -9;
    }
    double wq = 0.0;
    lapack_int info = LAPACKE_dormqr_work(layout, side, trans, m, n, k, a, lda, tau, c, ldc, &wq, -1);
    if (info != 0) return info;
    lapack_int lwork = (lapack_int)wq;
    double* work = (double*)malloc(sizeof(double) * (size_t)imax(1, lwork));
    if (!work) { lapacke_xerbla("LAPACKE_dormqr", LAPACK_WORK_MEMORY_ERROR); return LAPACK_WORK_MEMORY_ERROR; }
    info = LAPACKE_dormqr_work(layout, side, trans, m, n, k, a, lda, tau, c, ldc, work, lwork);
    free(work);
    return info;
}

// ------------------------------------------------------------------------------------------------ dlarft / dlarfb (Forward, Columnwise)
lapack_int LAPACKE_dlarft_work(int layout, char direct, char storev, lapack_int n, lapack_int k, const double* v,
                               lapack_int ldv, const double* tau, double* t, lapack_int ldt) {
    lapack_int info = 0;
    if (layout == LAPACK_COL_MAJOR) {
        dlarft_(&direct, &storev, &n, &k, v, &ldv, tau, t, &ldt, 1, 1);
    } else if (layout == LAPACK_ROW_MAJOR) {
        const bool col = lsame(storev, 'c');
        const lapack_int nrows_v = col ? n : k, ncols_v = col ? k : n;
        if (ldt < k) { info = -10; lapacke_xerbla("LAPACKE_dlarft_work", info); return info; }
        if (ldv < ncols_v) { info = -7; lapacke_xerbla("LAPACKE_dlarft_work", info); return info; }
        RowMajor rv, rt;
        if (!rv.in(v, nrows_v, ncols_v, ldv) || !rt.in(t, k, k, ldt)) { info = LAPACK_TRANSPOSE_MEMORY_ERROR; lapacke_xerbla("LAPACKE_dlarft_work", info); return info; }
        lapack_int ldv_t = imax(1, (int)rv.ldc), ldt_t = imax(1, (int)rt.ldc);
        dlarft_(&direct, &storev, &n, &k, rv.dev_cm ? rv.dev_cm : v, &ldv_t, tau, rt.dev_cm ? rt.dev_cm : t, &ldt_t, 1, 1);
        rt.out();
    } else { info = -1; lapacke_xerbla("LAPACKE_dlarft_work", info); }
    return info;
}
lapack_int LAPACKE_dlarft(int layout, char direct, char storev, lapack_int n, lapack_int k, const double* v, lapack_int ldv,
                          const double* tau, double* t, lapack_int ldt) {
    if (!LB_LAYOUT_OK(layout)) { lapacke_xerbla("LAPACKE_dlarft", -1); return -1; }
    if (get_nancheck()) {
        const bool col = lsame(storev, 'c');
        const lapack_int nrows_v = col ? n : (lsame(storev, 'r') ? k : 1), ncols_v = col ? k : (lsame(storev, 'r') ? n : 1);
        if (nan_scan(tau, k, 1, imax(1, k), 0)) return -8;
        if (dge_nan(layout, nrows_v, ncols_v, v, ldv)) return -6;
    }
    return LAPACKE_dlarft_work(layout, direct, storev, n, k, v, ldv, tau, t, ldt);
}
lapack_int LAPACKE_dlarfb_work(int layout, char side, char trans, char direct, char storev, lapack_int m, lapack_int n,
                               lapack_int k, const double* v, lapack_int ldv, const double* t, lapack_int ldt, double* c,
                               lapack_int ldc, double* work, lapack_int ldwork) {
    lapack_int info = 0;
    if (layout == LAPACK_COL_MAJOR) {
        dlarfb_(&side, &trans, &direct, &storev, &m, &n, &k, v, &ldv, t, &ldt, c, &ldc, work, &ldwork, 1, 1, 1, 1);
    } else if (layout == LAPACK_ROW_MAJOR) {
        const bool left = lsame(side, 'l'), col = lsame(storev, 'c');
        const lapack_int nrows_v = (col && left) ? m : ((col && !left) ? n : (!col ? k : 1));
        const lapack_int ncols_v = (!col && left) ? m : ((!col && !left) ? n : (col ? k : 1));
        if (ldc < n) { info = -14; lapacke_xerbla("LAPACKE_dlarfb_work", info); return info; }
        if (ldt < k) { info = -12; lapacke_xerbla("LAPACKE_dlarfb_work", info); return info; }
        if (ldv < ncols_v) { info = -10; lapacke_xerbla("LAPACKE_dlarfb_work", info); return info; }
        if ((col && k > nrows_v) || (!col && k > ncols_v)) { info = -8; lapacke_xerbla("LAPACKE_dlarfb_work", info); return info; }
        RowMajor rv, rt, rc;
        if (!rv.in(v, nrows_v, ncols_v, ldv) || !rt.in(t, k, k, ldt) || !rc.in(c, m, n, ldc)) {
            info = LAPACK_TRANSPOSE_MEMORY_ERROR; lapacke_xerbla("LAPACKE_dlarfb_work", info); return info;
        }
        lapack_int ldv_t = imax(1, (int)rv.ldc), ldt_t = imax(1, (int)rt.ldc), ldc_t = imax(1, (int)rc.ldc);
        dlarfb_(&side, &trans, &direct, &storev, &m, &n, &k, rv.dev_cm ? rv.dev_cm : v, &ldv_t, rt.dev_cm ? rt.dev_cm : t, &ldt_t,
                rc.dev_cm ? rc.dev_cm : c, &ldc_t, work, &ldwork, 1, 1, 1, 1);
        rc.out();
    } else { info = -1; lapacke_xerbla("LAPACKE_dlarfb_work", info); }
    return info;
}
lapack_int LAPACKE_dlarfb(int layout, char side, char trans, char direct, char storev, lapack_int m, lapack_int n,
                          lapack_int k, const double* v, lapack_int ldv, const double* t, lapack_int ldt, double* c,
                          lapack_int ldc) {
    if (!LB_LAYOUT_OK(layout)) { lapacke_xerbla("LAPACKE_dlarfb", -1); return -1; }
    const bool left = lsame(side, 'l'), col = lsame(storev, 'c');
    if (get_nancheck()) {
        const lapack_int nrows_v = (col && left) ? m : ((col && !left) ? n : (!col ? k : 1));
        const lapack_int ncols_v = (!col && left) ? m : ((!col && !left) ? n : (col ? k : 1));
        if ((col && k > nrows_v) || (!col && k > ncols_v)) { lapacke_xerbla("LAPACKE_dlarfb", -8); return -8; }
        // unit lower trapezoid of V (Forward/Columnwise): strictly lower part of the leading k x k plus the rest
        if (col) {
            bool bad = (layout == LAPACK_COL_MAJOR) ? nan_scan(v, nrows_v, ncols_v, ldv, 3)
                                                    : nan_scan(v, ncols_v, nrows_v, ldv, 2) && nan_scan(v, ncols_v, nrows_v, ldv, 0);
            if (bad) return -9;
        }
        if (dge_nan(layout, k, k, t, ldt)) return -11;
        if (dge_nan(layout, m, n, c, ldc)) return -13;
    }
    const lapack_int ldwork = left ? n : (lsame(side, 'r') ? m : 1);
    return LAPACKE_dlarfb_work(layout, side, trans, direct, storev, m, n, k, v, ldv, t, ldt, c, ldc, nullptr, ldwork);
}

// ------------------------------------------------------------------------------------------------ dlaswp
lapack_int LAPACKE_dlaswp_work(int layout, lapack_int n, double* a, lapack_int lda, lapack_int k1, lapack_int k2,
                               const lapack_int* ipiv, lapack_int incx) {
    lapack_int info = 0;
    if (layout == LAPACK_COL_MAJOR) {
        dlaswp_(&n, a, &lda, &k1, &k2, ipiv, &incx);
    } else if (layout == LAPACK_ROW_MAJOR) {
        lapack_int rows = imax(1, k2);
        const lapack_int ainc = incx < 0 ? -incx : incx;
        for (lapack_int i = k1; i <= k2; ++i) rows = imax(rows, ipiv[k1 + (i - k1) * ainc - 1]);
        if (lda < n) { info = -4; lapacke_xerbla("LAPACKE_dlaswp_work", info); return info; }
        RowMajor r;
        if (!r.in(a, rows, n, lda)) { info = LAPACK_TRANSPOSE_MEMORY_ERROR; lapacke_xerbla("LAPACKE_dlaswp_work", info); return info; }
        lapack_int lda_t = imax(1, (int)r.ldc);
        dlaswp_(&n, r.dev_cm ? r.dev_cm : a, &lda_t, &k1, &k2, ipiv, &incx);
        r.out();
    } else { info = -1; lapacke_xerbla("LAPACKE_dlaswp_work", info); }
    return info;
}
lapack_int LAPACKE_dlaswp(int layout, lapack_int n, double* a, lapack_int lda, lapack_int k1, lapack_int k2,
                          const lapack_int* ipiv, lapack_int incx) {
    if (!LB_LAYOUT_OK(layout)) { lapacke_xerbla("LAPACKE_dlaswp", -1); return -1; }
    return LAPACKE_dlaswp_work(layout, n, a, lda, k1, k2, ipiv, incx);   // no NaN check (lapacke_dlaswp.c:45-58)
}

}  // extern "C"
