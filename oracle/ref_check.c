/*
 * ref_check.c -- CPU restatement of the TESTING/LIN residual checkers used for the hot path.
 * TEST INFRASTRUCTURE ONLY (see oracle.h).  A result passes iff the ratio is < 30
 * (TESTING/dtest.in:13).
 */
#include "oracle.h"
#include <math.h>
#include <stdlib.h>

#define A_(i, j) a[(size_t)(i) + (size_t)(j) * lda]
#define B_(i, j) b[(size_t)(i) + (size_t)(j) * ldb]
#define X_(i, j) x[(size_t)(i) + (size_t)(j) * ldx]
#define AF_(i, j) afac[(size_t)(i) + (size_t)(j) * ldafac]
static int imin(int a, int b) { return a < b ? a : b; }
static int imax(int a, int b) { return a > b ? a : b; }

/* BLAS/SRC/dasum.f */
static double ref_dasum(int n, const double *x)
{
    double t = 0.0;   /* the 6-way unrolled sum in dasum.f is one left-to-right chain */
    for (int i = 0; i < n; ++i) t = t + fabs(x[i]);
    return t;
}

/* BLAS/SRC/dsyr.f, UPLO='L', unit stride */
static void ref_dsyr_lower(int n, double alpha, const double *x, double *a, int lda)
{
    for (int j = 0; j < n; ++j)
        if (x[j] != 0.0) {
            double temp = alpha * x[j];
            for (int i = j; i < n; ++i) A_(i, j) = A_(i, j) + x[i] * temp;
        }
}

/* BLAS/SRC/dsymm.f, SIDE='L' */
static void ref_dsymm_left(char uplo, int m, int n, double alpha, const double *a, int lda, const double *b,
                           int ldb, double beta, double *c, int ldc)
{
#define C_(i, j) c[(size_t)(i) + (size_t)(j) * ldc]
    if (ora_lsame(uplo, 'U')) {
        for (int j = 0; j < n; ++j)
            for (int i = 0; i < m; ++i) {
                double temp1 = alpha * B_(i, j), temp2 = 0.0;
                for (int k = 0; k < i; ++k) {
                    C_(k, j) = C_(k, j) + temp1 * A_(k, i);
                    temp2 = temp2 + B_(k, j) * A_(k, i);
                }
                if (beta == 0.0) C_(i, j) = temp1 * A_(i, i) + alpha * temp2;
                else C_(i, j) = beta * C_(i, j) + temp1 * A_(i, i) + alpha * temp2;
            }
    } else {
        for (int j = 0; j < n; ++j)
            for (int i = m - 1; i >= 0; --i) {
                double temp1 = alpha * B_(i, j), temp2 = 0.0;
                for (int k = i + 1; k < m; ++k) {
                    C_(k, j) = C_(k, j) + temp1 * A_(k, i);
                    temp2 = temp2 + B_(k, j) * A_(k, i);
                }
                if (beta == 0.0) C_(i, j) = temp1 * A_(i, i) + alpha * temp2;
                else C_(i, j) = beta * C_(i, j) + temp1 * A_(i, i) + alpha * temp2;
            }
    }
#undef C_
}

/* TESTING/LIN/dget01.f:155-206:  ||L*U - P*A|| / (N * ||A|| * eps), 1-norm.  afac is overwritten. */
void ora_dget01(int m, int n, const double *a, int lda, double *afac, int ldafac, const int *ipiv,
                double *rwork, double *resid)
{
    (void)rwork;
    if (m <= 0 || n <= 0) { *resid = 0.0; return; }
    double eps = ora_dlamch('E');
    double anorm = ora_dlange('1', m, n, a, lda);
    for (int k = n; k >= 1; --k) {
        if (k > m) {
            ora_dtrmv('L', 'N', 'U', m, afac, ldafac, &AF_(0, k - 1), 1);
        } else {
            double t = AF_(k - 1, k - 1);
            if (k + 1 <= m) {
                ora_dscal(m - k, t, &AF_(k, k - 1), 1);
                ora_dgemv('N', m - k, k - 1, 1.0, &AF_(k, 0), ldafac, &AF_(0, k - 1), 1, 1.0, &AF_(k, k - 1), 1);
            }
            AF_(k - 1, k - 1) = t + ora_ddot(k - 1, &AF_(k - 1, 0), ldafac, &AF_(0, k - 1), 1);
            ora_dtrmv('L', 'N', 'U', k - 1, afac, ldafac, &AF_(0, k - 1), 1);
        }
    }
    ora_dlaswp(n, afac, ldafac, 1, imin(m, n), ipiv, -1);
    for (int j = 0; j < n; ++j)
        for (int i = 0; i < m; ++i) AF_(i, j) = AF_(i, j) - A_(i, j);
    *resid = ora_dlange('1', m, n, afac, ldafac);
    if (anorm <= 0.0) { if (*resid != 0.0) *resid = 1.0 / eps; }
    else *resid = ((*resid / (double)n) / anorm) / eps;
}

/* TESTING/LIN/dget02.f:191-217:  ||B - op(A) X|| / (||A|| ||X|| eps).  b is overwritten. */
void ora_dget02(char trans, int m, int n, int nrhs, const double *a, int lda, const double *x, int ldx,
                double *b, int ldb, double *rwork, double *resid)
{
    (void)rwork;
    if (m <= 0 || n <= 0 || nrhs == 0) { *resid = 0.0; return; }
    int tr = ora_lsame(trans, 'T') || ora_lsame(trans, 'C');
    int n1 = tr ? n : m, n2 = tr ? m : n;
    double eps = ora_dlamch('E');
    double anorm = ora_lsame(trans, 'N') ? ora_dlange('1', m, n, a, lda) : ora_dlange('I', m, n, a, lda);
    if (anorm <= 0.0) { *resid = 1.0 / eps; return; }
    ora_dgemm(trans, 'N', n1, nrhs, n2, -1.0, a, lda, x, ldx, 1.0, b, ldb);
    *resid = 0.0;
    for (int j = 0; j < nrhs; ++j) {
        double bnorm = ref_dasum(n1, &B_(0, j)), xnorm = ref_dasum(n2, &X_(0, j));
        if (xnorm <= 0.0) *resid = 1.0 / eps;
        else { double r = ((bnorm / anorm) / xnorm) / eps; if (r > *resid) *resid = r; }
    }
}

/* TESTING/LIN/dget04.f:155-171:  max_j ||x - xact||_inf / ||xact||_inf * rcond / eps */
void ora_dget04(int n, int nrhs, const double *x, int ldx, const double *xact, int ldxact, double rcond,
                double *resid)
{
    if (n <= 0 || nrhs <= 0) { *resid = 0.0; return; }
    double eps = ora_dlamch('E');
    if (rcond < 0.0) { *resid = 1.0 / eps; return; }
    *resid = 0.0;
    for (int j = 0; j < nrhs; ++j) {
        const double *xa = xact + (size_t)j * ldxact;
        int ix = ora_idamax(n, xa, 1);
        double xnorm = fabs(xa[ix - 1]), diffnm = 0.0;
        for (int i = 0; i < n; ++i) { double d = fabs(X_(i, j) - xa[i]); if (d > diffnm) diffnm = d; }
        if (xnorm <= 0.0) { if (diffnm > 0.0) *resid = 1.0 / eps; }
        else { double r = (diffnm / xnorm) * rcond; if (r > *resid) *resid = r; }
    }
    if (*resid * eps < 1.0) *resid = *resid / eps;
}

/* TESTING/LIN/dpot01.f:151-215:  ||L L' - A|| / (N ||A|| eps)  (or U'U).  afac is overwritten. */
void ora_dpot01(char uplo, int n, const double *a, int lda, double *afac, int ldafac, double *rwork,
                double *resid)
{
    (void)rwork;
    if (n <= 0) { *resid = 0.0; return; }
    double eps = ora_dlamch('E');
    double anorm = ora_dlansy('1', uplo, n, a, lda);
    if (anorm <= 0.0) { *resid = 1.0 / eps; return; }
    if (ora_lsame(uplo, 'U')) {
        for (int k = n; k >= 1; --k) {
            double t = ora_ddot(k, &AF_(0, k - 1), 1, &AF_(0, k - 1), 1);
            AF_(k - 1, k - 1) = t;
            ora_dtrmv('U', 'T', 'N', k - 1, afac, ldafac, &AF_(0, k - 1), 1);
        }
        for (int j = 0; j < n; ++j) for (int i = 0; i <= j; ++i) AF_(i, j) = AF_(i, j) - A_(i, j);
    } else {
        for (int k = n; k >= 1; --k) {
            if (k + 1 <= n) ref_dsyr_lower(n - k, 1.0, &AF_(k, k - 1), &AF_(k, k), ldafac);
            double t = AF_(k - 1, k - 1);
            ora_dscal(n - k + 1, t, &AF_(k - 1, k - 1), 1);
        }
        for (int j = 0; j < n; ++j) for (int i = j; i < n; ++i) AF_(i, j) = AF_(i, j) - A_(i, j);
    }
    *resid = ora_dlansy('1', uplo, n, afac, ldafac);
    *resid = ((*resid / (double)n) / anorm) / eps;
}

/* TESTING/LIN/dpot02.f:174-196.  b is overwritten with B - A X. */
void ora_dpot02(char uplo, int n, int nrhs, const double *a, int lda, const double *x, int ldx, double *b,
                int ldb, double *rwork, double *resid)
{
    (void)rwork;
    if (n <= 0 || nrhs <= 0) { *resid = 0.0; return; }
    double eps = ora_dlamch('E');
    double anorm = ora_dlansy('1', uplo, n, a, lda);
    if (anorm <= 0.0) { *resid = 1.0 / eps; return; }
    ref_dsymm_left(uplo, n, nrhs, -1.0, a, lda, x, ldx, 1.0, b, ldb);
    *resid = 0.0;
    for (int j = 0; j < nrhs; ++j) {
        double bnorm = ref_dasum(n, &B_(0, j)), xnorm = ref_dasum(n, &X_(0, j));
        if (xnorm <= 0.0) *resid = 1.0 / eps;
        else { double r = ((bnorm / anorm) / xnorm) / eps; if (r > *resid) *resid = r; }
    }
}

/* TESTING/LIN/dqrt01.f:181-223, split so the factorization under test is supplied by the caller:
 * `af`/`tau` hold a DGEQRF result for `a` (m x n, all leading dimensions = lda >= m).
 * q (lda x m), r (lda x max(m,n)) are scratch.  result[0] = ||R - Q'A||/(M ||A|| eps),
 * result[1] = ||I - Q'Q||/(M eps). */
void ora_dqrt01(int m, int n, const double *a, const double *af, double *q, double *r, int lda,
                const double *tau, double *work, int lwork, double *rwork, double *result)
{
    (void)rwork;
    const double rogue = -1.0e10;
    int minmn = imin(m, n), info;
    double eps = ora_dlamch('E');
    ora_dlaset('F', m, m, rogue, rogue, q, lda);
    if (m > 1) ora_dlacpy('L', m - 1, n, af + 1, lda, q + 1, lda);
    ora_dorgqr(m, m, minmn, q, lda, tau, work, lwork, &info);
    ora_dlaset('F', m, n, 0.0, 0.0, r, lda);
    ora_dlacpy('U', m, n, af, lda, r, lda);
    ora_dgemm('T', 'N', m, n, m, -1.0, q, lda, a, lda, 1.0, r, lda);
    double anorm = ora_dlange('1', m, n, a, lda);
    double resid = ora_dlange('1', m, n, r, lda);
    result[0] = anorm > 0.0 ? ((resid / (double)imax(1, m)) / anorm) / eps : 0.0;
    ora_dlaset('F', m, m, 0.0, 1.0, r, lda);
    ora_dsyrk('U', 'T', m, m, -1.0, q, lda, 1.0, r, lda);
    resid = ora_dlansy('1', 'U', m, r, lda);
    result[1] = (resid / (double)imax(1, m)) / eps;
}
