// potrf.cu -- Cholesky factorization on one B200: DPOTRF / DPOTRF2 / DPOTRS.
//
// Reference path: SRC/dpotrf.f:166-240 (blocked; left-looking by block column in the main tree, the
// right-looking ordering used here is the reference's own SRC/VARIANTS/cholesky/RL/dpotrf.f:205-229),
// SRC/dpotrf2.f:165-230 (recursive diagonal block; leaf = test "<= 0 or NaN", then SQRT), SRC/dpotrs.f:168-196.
//
// B200 design: outer block NB (default 512) so the trailing DSYRK (C -= L21*L21^T, lower tiles only)
// runs as DMMA tiles; the NB x NB diagonal block is factored by the DPOTRF2 recursion down to a 32 x 32
// leaf held in shared memory by one CTA; the panel solve is the recursive DTRSM.  Look-ahead: the next
// block column is updated first and its diagonal block + panel solve run on a high-priority stream while
// the rest of the trailing update proceeds.  Only the UPLO triangle is ever read or written
// (dpotrf.f:65-71; TESTING/LIN/dchkpo.f leaves stale data in the other triangle).
#include "lb_internal.h"
#include <mutex>

namespace lb {

static int g_po_nb = 512, g_po_lookahead = 1;
void potrf_set_params(int nb, int lookahead) {
    if (nb > 0) g_po_nb = nb;
    if (lookahead >= 0) g_po_lookahead = lookahead;
}

int potrf_block() { return g_po_nb; }

constexpr int PL = 32;   // leaf size

// One CTA (32 x 32 threads) factors an n x n (n <= 32) block in shared memory, lower form.
// For UPLO='U' the block is read/written transposed so the same code produces U = L^T.
__global__ void __launch_bounds__(PL* PL) potrf_leaf_kernel(int n, double* __restrict__ A, i64 lda, bool upper, int* info,
                                                            int info_off) {
    __shared__ double L[PL][PL + 1];
    __shared__ int s_fail;
    const int i = threadIdx.x, j = threadIdx.y;     // element (i,j) of the lower triangle, i >= j
    const bool mine = (i < n && j < n && i >= j);
    if (*info != 0) return;                         // an earlier leaf failed: the factorization has been abandoned (dpotrf.f:239-240)
    if (mine) L[i][j] = upper ? A[j + (i64)i * lda] : A[i + (i64)j * lda];
    if (i == 0 && j == 0) s_fail = 0;
    __syncthreads();
    for (int k = 0; k < n; ++k) {
        if (i == 0 && j == 0) {
            double d = L[k][k];
            if (d <= 0.0 || d != d) {                // dpotrf2.f:169-172
                s_fail = k + 1;
            } else {
                L[k][k] = sqrt(d);                   // dpotrf2.f:176
            }
        }
        __syncthreads();
        if (s_fail) break;
        if (j == k && i > k && i < n) L[i][k] = L[i][k] / L[k][k];   // DTRSM R,L,T,N: division by the diagonal
        __syncthreads();
        if (mine && j > k) L[i][j] = fma(-L[i][k], L[j][k], L[i][j]);
        __syncthreads();
    }
    if (s_fail) {
        if (i == 0 && j == 0 && *info == 0) *info = info_off + s_fail;
        // columns < fail-1 are final and stored; the rest is left partially updated like the reference
    }
    if (mine) {
        if (upper) A[j + (i64)i * lda] = L[i][j]; else A[i + (i64)j * lda] = L[i][j];
    }
}

static void potrf_leaf(cudaStream_t s, bool upper, int n, double* A, i64 lda, int* info, int info_off) {
    dim3 block(PL, PL);
    potrf_leaf_kernel<<<1, block, 0, s>>>(n, A, lda, upper, info, info_off);
    count_launch();
}

// DPOTRF2 recursion (dpotrf2.f:181-228)
static void potrf_rec(cudaStream_t s, bool upper, int n, double* A, i64 lda, int* info, int info_off) {
    if (n <= 0) return;
    if (n <= PL) { potrf_leaf(s, upper, n, A, lda, info, info_off); return; }
    int n1 = PL;
    while (n1 * 2 < n) n1 *= 2;
    const int n2 = n - n1;
    potrf_rec(s, upper, n1, A, lda, info, info_off);
    double* A22 = A + n1 + (i64)n1 * lda;
    if (upper) {
        double* A12 = A + (i64)n1 * lda;
        trsm(s, 'L', 'U', 'T', 'N', n1, n2, 1.0, A, lda, A12, lda);          // dpotrf2.f:201
        syrk(s, 'U', 'T', n2, n1, -1.0, A12, lda, 1.0, A22, lda);            // dpotrf2.f:206
    } else {
        double* A21 = A + n1;
        trsm(s, 'R', 'L', 'T', 'N', n2, n1, 1.0, A, lda, A21, lda);          // dpotrf2.f:217
        syrk(s, 'L', 'N', n2, n1, -1.0, A21, lda, 1.0, A22, lda);            // dpotrf2.f:222
    }
    potrf_rec(s, upper, n2, A22, lda, info, info_off + n1);
}

// one mutex for every driver that uses the look-ahead streams / events of lb::aux() (runtime.cu): concurrent host threads calling
// different factorizations through the device API must not interleave their event joins
static std::recursive_mutex& g_po_mutex = driver_mutex();

// every Level-3 kernel queued by the Cholesky drivers carries the INFO word as a guard (runtime.cu)
struct GuardScope {
    explicit GuardScope(const int* p) { set_kernel_guard(p); }
    ~GuardScope() { set_kernel_guard(nullptr); }
};

void potrf2(cudaStream_t s, char uplo, int n, double* A, i64 lda, int* info) {
    std::lock_guard<std::recursive_mutex> lock(g_po_mutex);
    LB_CUDA_CHECK(cudaMemsetAsync(info, 0, sizeof(int), s));
    GuardScope guard(info);
    potrf_rec(s, uplo == 'U' || uplo == 'u', n, A, lda, info, 0);
}

void potrf(cudaStream_t s, char uplo, int n, double* A, i64 lda, int* info) {
    std::lock_guard<std::recursive_mutex> lock(g_po_mutex);
    LB_CUDA_CHECK(cudaMemsetAsync(info, 0, sizeof(int), s));
    if (n <= 0) return;
    GuardScope guard(info);
    const bool upper = (uplo == 'U' || uplo == 'u');
    const int nb = g_po_nb;
    if (nb >= n) { potrf_rec(s, upper, n, A, lda, info, 0); return; }

    const bool la = g_po_lookahead != 0;
    Aux& ax = aux();
    cudaStream_t sp = la ? ax.panel_stream : s;
    cudaStream_t su = la ? ax.update_stream : s;
    cudaEvent_t ev_panel = ax.ev[3], ev_next = ax.ev[4], ev_join = ax.ev[5];
    if (la) {
        LB_CUDA_CHECK(cudaEventRecord(ev_join, s));
        LB_CUDA_CHECK(cudaStreamWaitEvent(sp, ev_join, 0));
        LB_CUDA_CHECK(cudaStreamWaitEvent(su, ev_join, 0));
    }
    // panel(j): factor the diagonal block and solve for the block column / block row
    StreamOut* so = stream_out();
    auto panel = [&](int j, int jb) {
        double* Ajj = A + j + (i64)j * lda;
        potrf_rec(sp, upper, jb, Ajj, lda, info, j);
        const int rest = n - j - jb;
        if (rest > 0) {
            trsm_set_inverse_leaves(1);          // panel solve with DMMA leaves (diagonal blocks of a Cholesky factor)
            if (upper) trsm(sp, 'L', 'U', 'T', 'N', jb, rest, 1.0, Ajj, lda, A + j + (i64)(j + jb) * lda, lda);
            else trsm(sp, 'R', 'L', 'T', 'N', rest, jb, 1.0, Ajj, lda, A + (j + jb) + (i64)j * lda, lda);
            trsm_set_inverse_leaves(0);
        }
        if (so) {
            // block column j (lower) / block row j (upper) of the factor is final: start its download now
            LB_CUDA_CHECK(cudaEventRecord(so->ev, sp));
            LB_CUDA_CHECK(cudaStreamWaitEvent(so->copy_stream, so->ev, 0));
            if (upper)
                LB_CUDA_CHECK(cudaMemcpy2DAsync(so->host + j + (i64)j * so->ldh, so->ldh * 8, Ajj, lda * 8, (size_t)jb * 8, n - j,
                                                cudaMemcpyDeviceToHost, so->copy_stream));
            else
                LB_CUDA_CHECK(cudaMemcpy2DAsync(so->host + j + (i64)j * so->ldh, so->ldh * 8, Ajj, lda * 8, (size_t)(n - j) * 8, jb,
                                                cudaMemcpyDeviceToHost, so->copy_stream));
            so->done_cols = j + jb;
        }
    };
    panel(0, min(nb, n));
    if (la) LB_CUDA_CHECK(cudaEventRecord(ev_panel, sp));

    for (int j = 0; j < n; j += nb) {
        const int jb = min(nb, n - j);
        const int jn = j + jb;
        if (jn >= n) break;
        if (la) LB_CUDA_CHECK(cudaStreamWaitEvent(su, ev_panel, 0));
        const int jb2 = min(nb, n - jn);
        const int rest = n - jn - jb2;
        if (upper) {
            // U12 = A(j:jn, jn:n).  next block row: A(jn:jn+jb2, jn:n) -= U12(:,0:jb2)^T * U12   (upper part)
            const double* U12 = A + j + (i64)jn * lda;
            gemm(su, 'T', 'N', jb2, n - jn, jb, -1.0, U12, lda, U12, lda, 1.0, A + jn + (i64)jn * lda, lda, 2);
            if (la) { LB_CUDA_CHECK(cudaEventRecord(ev_next, su)); LB_CUDA_CHECK(cudaStreamWaitEvent(sp, ev_next, 0)); }
            panel(jn, jb2);
            if (la) LB_CUDA_CHECK(cudaEventRecord(ev_panel, sp));
            if (rest > 0) {
                const double* U13 = A + j + (i64)(jn + jb2) * lda;
                syrk(su, 'U', 'T', rest, jb, -1.0, U13, lda, 1.0, A + (jn + jb2) + (i64)(jn + jb2) * lda, lda);
            }
        } else {
            // L21 = A(jn:n, j:jn).  next block column: A(jn:n, jn:jn+jb2) -= L21 * L21(0:jb2,:)^T   (lower part)
            const double* L21 = A + jn + (i64)j * lda;
            gemm(su, 'N', 'T', n - jn, jb2, jb, -1.0, L21, lda, L21, lda, 1.0, A + jn + (i64)jn * lda, lda, 1);
            if (la) { LB_CUDA_CHECK(cudaEventRecord(ev_next, su)); LB_CUDA_CHECK(cudaStreamWaitEvent(sp, ev_next, 0)); }
            panel(jn, jb2);
            if (la) LB_CUDA_CHECK(cudaEventRecord(ev_panel, sp));
            if (rest > 0) {
                const double* L31 = A + (jn + jb2) + (i64)j * lda;
                syrk(su, 'L', 'N', rest, jb, -1.0, L31, lda, 1.0, A + (jn + jb2) + (i64)(jn + jb2) * lda, lda);
            }
        }
    }
    if (la) {
        LB_CUDA_CHECK(cudaEventRecord(ev_join, su));
        LB_CUDA_CHECK(cudaStreamWaitEvent(s, ev_join, 0));
        LB_CUDA_CHECK(cudaEventRecord(ev_next, sp));
        LB_CUDA_CHECK(cudaStreamWaitEvent(s, ev_next, 0));
    }
}

// DPOTRS (SRC/dpotrs.f:168-196)
void potrs(cudaStream_t s, char uplo, int n, int nrhs, const double* A, i64 lda, double* B, i64 ldb) {
    if (n <= 0 || nrhs <= 0) return;
    if (uplo == 'U' || uplo == 'u') {
        trsm(s, 'L', 'U', 'T', 'N', n, nrhs, 1.0, A, lda, B, ldb);
        trsm(s, 'L', 'U', 'N', 'N', n, nrhs, 1.0, A, lda, B, ldb);
    } else {
        trsm(s, 'L', 'L', 'N', 'N', n, nrhs, 1.0, A, lda, B, ldb);
        trsm(s, 'L', 'L', 'T', 'N', n, nrhs, 1.0, A, lda, B, ldb);
    }
}

}  // namespace lb
