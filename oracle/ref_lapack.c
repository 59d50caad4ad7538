/*
 * ref_lapack.c -- CPU restatement of the reference LAPACK routines on the one-sided factorization
 * hot path (LU / Cholesky / Householder QR) and the few aux routines they need.
 * TEST INFRASTRUCTURE ONLY (see oracle.h).  Each function cites the /root/reference file:line it
 * follows; operation order is the reference's (build with -ffp-contract=off).
 */
#include "oracle.h"
#include <math.h>
#include <float.h>
#include <stdlib.h>
#include <string.h>

#define A_(i, j) a[(size_t)(i) + (size_t)(j) * lda]
#define B_(i, j) b[(size_t)(i) + (size_t)(j) * ldb]
#define C_(i, j) c[(size_t)(i) + (size_t)(j) * ldc]
#define V_(i, j) v[(size_t)(i) + (size_t)(j) * ldv]
#define T_(i, j) t[(size_t)(i) + (size_t)(j) * ldt]
#define W_(i, j) work[(size_t)(i) + (size_t)(j) * ldwork]
static int imin(int a, int b) { return a < b ? a : b; }
static int imax(int a, int b) { return a > b ? a : b; }

/* ------------------------------------------------------------------------------------------ */
/* INSTALL/dlamch.f:99-122 for IEEE binary64: eps = 2^-53 (rounding), sfmin = tiny.            */
double ora_dlamch(char cmach)
{
    if (ora_lsame(cmach, 'E')) return DBL_EPSILON * 0.5;
    if (ora_lsame(cmach, 'S')) return DBL_MIN;              /* 1/huge < tiny, so sfmin = tiny */
    if (ora_lsame(cmach, 'B')) return 2.0;
    if (ora_lsame(cmach, 'P')) return DBL_EPSILON * 0.5 * 2.0;
    if (ora_lsame(cmach, 'N')) return 53.0;
    if (ora_lsame(cmach, 'R')) return 1.0;
    if (ora_lsame(cmach, 'M')) return -1021.0;
    if (ora_lsame(cmach, 'U')) return DBL_MIN;
    if (ora_lsame(cmach, 'L')) return 1024.0;
    if (ora_lsame(cmach, 'O')) return DBL_MAX;
    return 0.0;
}

/* SRC/ilaenv.f:289-302,374-380 (NB) and :623-630 (NX); overridable like TESTING/LIN/xlaenv.f:101. */
static int g_nb_getrf = 64, g_nb_potrf = 64, g_nb_geqrf = 32, g_nx_geqrf = 128, g_nb_getri = 64;
void ora_set_nb_getri(int nb) { g_nb_getri = nb; }
void ora_set_nb(int nb_getrf, int nb_potrf, int nb_geqrf, int nx_geqrf)
{
    g_nb_getrf = nb_getrf; g_nb_potrf = nb_potrf; g_nb_geqrf = nb_geqrf; g_nx_geqrf = nx_geqrf;
}
int ora_ilaenv_nb(const char *name)
{
    if (!strcmp(name, "DGETRF")) return g_nb_getrf;
    if (!strcmp(name, "DPOTRF")) return g_nb_potrf;
    if (!strcmp(name, "DGEQRF") || !strcmp(name, "DORGQR") || !strcmp(name, "DORMQR")) return g_nb_geqrf;   /* ilaenv.f:416-436: 32 */
    if (!strcmp(name, "DGETRI") || !strcmp(name, "DTRTRI")) return g_nb_getri;                               /* ilaenv.f:361-367, 475-481: 64 */
    return 1;
}

/* SRC/dlapy2.f:96-112 */
double ora_dlapy2(double x, double y)
{
    int xnan = (x != x), ynan = (y != y);
    double r = 0.0;
    if (xnan) r = x;
    if (ynan) r = y;
    if (!(xnan || ynan)) {
        double xa = fabs(x), ya = fabs(y);
        double w = xa > ya ? xa : ya, z = xa < ya ? xa : ya;
        if (z == 0.0 || w > DBL_MAX) r = w;
        else { double q = z / w; r = w * sqrt(1.0 + q * q); }
    }
    return r;
}

/* SRC/dlaruv.f:401-447.  The MM(i,:) table rows are the base-4096 digits of a**i mod 2**48 with
 * a = 33952834046453 (row 1 = 494,322,2508,2549), so x(i) = (seed * a**i mod 2**48) / 2**48 and
 * the seed leaves as seed * a**n.  The digit-wise float sum at dlaruv.f:417-418 is exact in
 * binary64 (48 significant bits), so the "== 1.0" retry at :420-433 can never trigger here. */
void ora_dlaruv(int iseed[4], int n, double *x)
{
    const unsigned long long MASK = (1ULL << 48) - 1, AMUL = 33952834046453ULL;
    if (n < 1) return;
    if (n > 128) n = 128;
    unsigned long long s = ((unsigned long long)iseed[0] << 36) | ((unsigned long long)iseed[1] << 24) |
                           ((unsigned long long)iseed[2] << 12) | (unsigned long long)iseed[3];
    unsigned long long p = s;
    for (int i = 0; i < n; ++i) {
        p = (p * AMUL) & MASK;
        x[i] = (double)p * (1.0 / 281474976710656.0);
    }
    iseed[0] = (int)((p >> 36) & 4095); iseed[1] = (int)((p >> 24) & 4095);
    iseed[2] = (int)((p >> 12) & 4095); iseed[3] = (int)(p & 4095);
}

/* SRC/dlarnv.f:140-170; idist 1: U(0,1), 2: U(-1,1), 3: N(0,1). */
void ora_dlarnv(int idist, int iseed[4], long n, double *x)
{
    const double twopi = 6.28318530717958647692528676655900576839;
    double u[128];
    for (long iv = 0; iv < n; iv += 64) {
        int il = (int)((n - iv) < 64 ? (n - iv) : 64);
        int il2 = (idist == 3) ? 2 * il : il;
        ora_dlaruv(iseed, il2, u);
        if (idist == 1) for (int i = 0; i < il; ++i) x[iv + i] = u[i];
        else if (idist == 2) for (int i = 0; i < il; ++i) x[iv + i] = 2.0 * u[i] - 1.0;
        else if (idist == 3)
            for (int i = 0; i < il; ++i) x[iv + i] = sqrt(-2.0 * log(u[2 * i])) * cos(twopi * u[2 * i + 1]);
    }
}

/* SRC/dlange.f:144-190 ('M', '1'/'O', 'I'; NaN-propagating max via DISNAN). */
double ora_dlange(char norm, int m, int n, const double *a, int lda)
{
    double value = 0.0;
    if (imin(m, n) == 0) return 0.0;
    if (ora_lsame(norm, 'M')) {
        for (int j = 0; j < n; ++j)
            for (int i = 0; i < m; ++i) { double t = fabs(A_(i, j)); if (value < t || t != t) value = t; }
    } else if (ora_lsame(norm, 'O') || norm == '1') {
        for (int j = 0; j < n; ++j) {
            double sum = 0.0;
            for (int i = 0; i < m; ++i) sum = sum + fabs(A_(i, j));
            if (value < sum || sum != sum) value = sum;
        }
    } else if (ora_lsame(norm, 'I')) {
        double *w = (double *)calloc((size_t)m, sizeof(double));
        for (int j = 0; j < n; ++j)
            for (int i = 0; i < m; ++i) w[i] = w[i] + fabs(A_(i, j));
        for (int i = 0; i < m; ++i) { double t = w[i]; if (value < t || t != t) value = t; }
        free(w);
    } else {   /* 'F': plain two-pass scaled form (not on the checker path) */
        double scale = 0.0, ssq = 1.0;
        for (int j = 0; j < n; ++j)
            for (int i = 0; i < m; ++i) {
                double t = fabs(A_(i, j));
                if (t != 0.0) {
                    if (scale < t) { ssq = 1.0 + ssq * (scale / t) * (scale / t); scale = t; }
                    else ssq = ssq + (t / scale) * (t / scale);
                }
            }
        value = scale * sqrt(ssq);
    }
    return value;
}

/* SRC/dlansy.f:150-199 ('M', and '1'='O'='I' for a symmetric matrix). */
double ora_dlansy(char norm, char uplo, int n, const double *a, int lda)
{
    double value = 0.0;
    if (n == 0) return 0.0;
    int upper = ora_lsame(uplo, 'U');
    if (ora_lsame(norm, 'M')) {
        for (int j = 0; j < n; ++j) {
            int i0 = upper ? 0 : j, i1 = upper ? j + 1 : n;
            for (int i = i0; i < i1; ++i) { double t = fabs(A_(i, j)); if (value < t || t != t) value = t; }
        }
        return value;
    }
    double *w = (double *)calloc((size_t)n, sizeof(double));
    if (upper) {
        for (int j = 0; j < n; ++j) {
            double sum = 0.0;
            for (int i = 0; i < j; ++i) { double t = fabs(A_(i, j)); sum = sum + t; w[i] = w[i] + t; }
            w[j] = sum + fabs(A_(j, j));
        }
        for (int i = 0; i < n; ++i) { double sum = w[i]; if (value < sum || sum != sum) value = sum; }
    } else {
        for (int j = 0; j < n; ++j) {
            double sum = w[j] + fabs(A_(j, j));
            for (int i = j + 1; i < n; ++i) { double t = fabs(A_(i, j)); sum = sum + t; w[i] = w[i] + t; }
            if (value < sum || sum != sum) value = sum;
        }
    }
    free(w);
    return value;
}

/* SRC/dlacpy.f */
void ora_dlacpy(char uplo, int m, int n, const double *a, int lda, double *b, int ldb)
{
    if (ora_lsame(uplo, 'U')) {
        for (int j = 0; j < n; ++j) for (int i = 0; i < imin(j + 1, m); ++i) B_(i, j) = A_(i, j);
    } else if (ora_lsame(uplo, 'L')) {
        for (int j = 0; j < n; ++j) for (int i = j; i < m; ++i) B_(i, j) = A_(i, j);
    } else {
        for (int j = 0; j < n; ++j) for (int i = 0; i < m; ++i) B_(i, j) = A_(i, j);
    }
}

/* SRC/dlaset.f: off-diagonals := alpha, diagonal := beta */
void ora_dlaset(char uplo, int m, int n, double alpha, double beta, double *a, int lda)
{
    if (ora_lsame(uplo, 'U')) {
        for (int j = 1; j < n; ++j) for (int i = 0; i < imin(j, m); ++i) A_(i, j) = alpha;
    } else if (ora_lsame(uplo, 'L')) {
        for (int j = 0; j < imin(m, n); ++j) for (int i = j + 1; i < m; ++i) A_(i, j) = alpha;
    } else {
        for (int j = 0; j < n; ++j) for (int i = 0; i < m; ++i) A_(i, j) = alpha;
    }
    for (int i = 0; i < imin(m, n); ++i) A_(i, i) = beta;
}

/* ------------------------------------------------------------------------------------------ */
/* LU                                                                                          */

/* SRC/dlaswp.f:138-183.  k1,k2 1-based, ipiv entries 1-based; 32-column strips then remainder. */
void ora_dlaswp(int n, double *a, int lda, int k1, int k2, const int *ipiv, int incx)
{
    int ix0, i1, i2, inc;
    if (incx > 0) { ix0 = k1; i1 = k1; i2 = k2; inc = 1; }
    else if (incx < 0) { ix0 = k1 + (k1 - k2) * incx; i1 = k2; i2 = k1; inc = -1; }
    else return;
    int n32 = (n / 32) * 32;
    for (int j = 0; j < n32; j += 32) {
        int ix = ix0;
        for (int i = i1; inc > 0 ? i <= i2 : i >= i2; i += inc, ix += incx) {
            int ip = ipiv[ix - 1];
            if (ip != i)
                for (int k = j; k < j + 32; ++k) {
                    double temp = A_(i - 1, k); A_(i - 1, k) = A_(ip - 1, k); A_(ip - 1, k) = temp;
                }
        }
    }
    if (n32 != n) {
        int ix = ix0;
        for (int i = i1; inc > 0 ? i <= i2 : i >= i2; i += inc, ix += incx) {
            int ip = ipiv[ix - 1];
            if (ip != i)
                for (int k = n32; k < n; ++k) {
                    double temp = A_(i - 1, k); A_(i - 1, k) = A_(ip - 1, k); A_(ip - 1, k) = temp;
                }
        }
    }
}

/* SRC/dgetrf2.f:170-265 (recursive panel LU). */
void ora_dgetrf2(int m, int n, double *a, int lda, int *ipiv, int *info)
{
    *info = 0;
    if (m < 0) *info = -1; else if (n < 0) *info = -2; else if (lda < imax(1, m)) *info = -4;
    if (*info != 0) return;
    if (m == 0 || n == 0) return;
    if (m == 1) {                                           /* :170-177 */
        ipiv[0] = 1;
        if (A_(0, 0) == 0.0) *info = 1;
    } else if (n == 1) {                                    /* :179-214 */
        double sfmin = ora_dlamch('S');
        int i = ora_idamax(m, a, 1);
        ipiv[0] = i;
        if (A_(i - 1, 0) != 0.0) {
            if (i != 1) { double temp = A_(0, 0); A_(0, 0) = A_(i - 1, 0); A_(i - 1, 0) = temp; }
            if (fabs(A_(0, 0)) >= sfmin) ora_dscal(m - 1, 1.0 / A_(0, 0), &A_(1, 0), 1);
            else for (int r = 1; r < m; ++r) A_(r, 0) = A_(r, 0) / A_(0, 0);
        } else *info = 1;
    } else {                                                /* :216-263 */
        int n1 = imin(m, n) / 2, n2 = n - n1, iinfo;
        ora_dgetrf2(m, n1, a, lda, ipiv, &iinfo);
        if (*info == 0 && iinfo > 0) *info = iinfo;
        ora_dlaswp(n2, &A_(0, n1), lda, 1, n1, ipiv, 1);
        ora_dtrsm('L', 'L', 'N', 'U', n1, n2, 1.0, a, lda, &A_(0, n1), lda);
        ora_dgemm('N', 'N', m - n1, n2, n1, -1.0, &A_(n1, 0), lda, &A_(0, n1), lda, 1.0, &A_(n1, n1), lda);
        ora_dgetrf2(m - n1, n2, &A_(n1, n1), lda, ipiv + n1, &iinfo);
        if (*info == 0 && iinfo > 0) *info = iinfo + n1;
        for (int i = n1; i < imin(m, n); ++i) ipiv[i] = ipiv[i] + n1;
        ora_dlaswp(n1, a, lda, n1 + 1, imin(m, n), ipiv, 1);
    }
}

/* SRC/dgetrf.f:144-219 (right-looking blocked LU, NB from ILAENV). */
void ora_dgetrf(int m, int n, double *a, int lda, int *ipiv, int *info)
{
    *info = 0;
    if (m < 0) *info = -1; else if (n < 0) *info = -2; else if (lda < imax(1, m)) *info = -4;
    if (*info != 0) return;
    if (m == 0 || n == 0) return;
    int nb = ora_ilaenv_nb("DGETRF"), mn = imin(m, n);
    if (nb <= 1 || nb >= mn) { ora_dgetrf2(m, n, a, lda, ipiv, info); return; }
    for (int j = 0; j < mn; j += nb) {                      /* j is 0-based; reference J = j+1 */
        int jb = imin(mn - j, nb), iinfo;
        ora_dgetrf2(m - j, jb, &A_(j, j), lda, ipiv + j, &iinfo);
        if (*info == 0 && iinfo > 0) *info = iinfo + j;
        for (int i = j; i < imin(m, j + jb); ++i) ipiv[i] = j + ipiv[i];
        ora_dlaswp(j, a, lda, j + 1, j + jb, ipiv, 1);
        if (j + jb < n) {
            ora_dlaswp(n - j - jb, &A_(0, j + jb), lda, j + 1, j + jb, ipiv, 1);
            ora_dtrsm('L', 'L', 'N', 'U', jb, n - j - jb, 1.0, &A_(j, j), lda, &A_(j, j + jb), lda);
            if (j + jb < m)
                ora_dgemm('N', 'N', m - j - jb, n - j - jb, jb, -1.0, &A_(j + jb, j), lda, &A_(j, j + jb), lda,
                          1.0, &A_(j + jb, j + jb), lda);
        }
    }
}

/* SRC/dgetrs.f:158-218 */
void ora_dgetrs(char trans, int n, int nrhs, const double *a, int lda, const int *ipiv, double *b, int ldb,
                int *info)
{
    *info = 0;
    int notran = ora_lsame(trans, 'N');
    if (!notran && !ora_lsame(trans, 'T') && !ora_lsame(trans, 'C')) *info = -1;
    else if (n < 0) *info = -2;
    else if (nrhs < 0) *info = -3;
    else if (lda < imax(1, n)) *info = -5;
    else if (ldb < imax(1, n)) *info = -8;
    if (*info != 0) return;
    if (n == 0 || nrhs == 0) return;
    if (notran) {
        ora_dlaswp(nrhs, b, ldb, 1, n, ipiv, 1);
        ora_dtrsm('L', 'L', 'N', 'U', n, nrhs, 1.0, a, lda, b, ldb);
        ora_dtrsm('L', 'U', 'N', 'N', n, nrhs, 1.0, a, lda, b, ldb);
    } else {
        ora_dtrsm('L', 'U', 'T', 'N', n, nrhs, 1.0, a, lda, b, ldb);
        ora_dtrsm('L', 'L', 'T', 'U', n, nrhs, 1.0, a, lda, b, ldb);
        ora_dlaswp(nrhs, b, ldb, 1, n, ipiv, -1);
    }
}

/* SRC/dgesv.f:148-172 */
void ora_dgesv(int n, int nrhs, double *a, int lda, int *ipiv, double *b, int ldb, int *info)
{
    *info = 0;
    if (n < 0) *info = -1; else if (nrhs < 0) *info = -2; else if (lda < imax(1, n)) *info = -4;
    else if (ldb < imax(1, n)) *info = -7;
    if (*info != 0) return;
    ora_dgetrf(n, n, a, lda, ipiv, info);
    if (*info == 0) ora_dgetrs('N', n, nrhs, a, lda, ipiv, b, ldb, info);
}

/* ------------------------------------------------------------------------------------------ */
/* Cholesky                                                                                    */

/* SRC/dpotrf2.f:165-230 */
void ora_dpotrf2(char uplo, int n, double *a, int lda, int *info)
{
    *info = 0;
    int upper = ora_lsame(uplo, 'U');
    if (!upper && !ora_lsame(uplo, 'L')) *info = -1; else if (n < 0) *info = -2;
    else if (lda < imax(1, n)) *info = -4;
    if (*info != 0) return;
    if (n == 0) return;
    if (n == 1) {
        if (A_(0, 0) <= 0.0 || A_(0, 0) != A_(0, 0)) { *info = 1; return; }
        A_(0, 0) = sqrt(A_(0, 0));
    } else {
        int n1 = n / 2, n2 = n - n1, iinfo;
        ora_dpotrf2(uplo, n1, a, lda, &iinfo);
        if (iinfo != 0) { *info = iinfo; return; }
        if (upper) {
            ora_dtrsm('L', 'U', 'T', 'N', n1, n2, 1.0, a, lda, &A_(0, n1), lda);
            ora_dsyrk(uplo, 'T', n2, n1, -1.0, &A_(0, n1), lda, 1.0, &A_(n1, n1), lda);
        } else {
            ora_dtrsm('R', 'L', 'T', 'N', n2, n1, 1.0, a, lda, &A_(n1, 0), lda);
            ora_dsyrk(uplo, 'N', n2, n1, -1.0, &A_(n1, 0), lda, 1.0, &A_(n1, n1), lda);
        }
        ora_dpotrf2(uplo, n2, &A_(n1, n1), lda, &iinfo);
        if (iinfo != 0) { *info = iinfo + n1; return; }
    }
}

/* SRC/dpotrf.f:145-240 (left-looking by block column). */
void ora_dpotrf(char uplo, int n, double *a, int lda, int *info)
{
    *info = 0;
    int upper = ora_lsame(uplo, 'U');
    if (!upper && !ora_lsame(uplo, 'L')) *info = -1; else if (n < 0) *info = -2;
    else if (lda < imax(1, n)) *info = -4;
    if (*info != 0) return;
    if (n == 0) return;
    int nb = ora_ilaenv_nb("DPOTRF");
    if (nb <= 1 || nb >= n) { ora_dpotrf2(uplo, n, a, lda, info); return; }
    for (int j = 0; j < n; j += nb) {
        int jb = imin(nb, n - j);
        if (upper) {
            ora_dsyrk('U', 'T', jb, j, -1.0, &A_(0, j), lda, 1.0, &A_(j, j), lda);
            ora_dpotrf2('U', jb, &A_(j, j), lda, info);
            if (*info != 0) { *info = *info + j; return; }
            if (j + jb < n) {
                ora_dgemm('T', 'N', jb, n - j - jb, j, -1.0, &A_(0, j), lda, &A_(0, j + jb), lda, 1.0,
                          &A_(j, j + jb), lda);
                ora_dtrsm('L', 'U', 'T', 'N', jb, n - j - jb, 1.0, &A_(j, j), lda, &A_(j, j + jb), lda);
            }
        } else {
            ora_dsyrk('L', 'N', jb, j, -1.0, &A_(j, 0), lda, 1.0, &A_(j, j), lda);
            ora_dpotrf2('L', jb, &A_(j, j), lda, info);
            if (*info != 0) { *info = *info + j; return; }
            if (j + jb < n) {
                ora_dgemm('N', 'T', n - j - jb, jb, j, -1.0, &A_(j + jb, 0), lda, &A_(j, 0), lda, 1.0,
                          &A_(j + jb, j), lda);
                ora_dtrsm('R', 'L', 'T', 'N', n - j - jb, jb, 1.0, &A_(j, j), lda, &A_(j + jb, j), lda);
            }
        }
    }
}

/* SRC/dpotrs.f:147-196 */
void ora_dpotrs(char uplo, int n, int nrhs, const double *a, int lda, double *b, int ldb, int *info)
{
    *info = 0;
    int upper = ora_lsame(uplo, 'U');
    if (!upper && !ora_lsame(uplo, 'L')) *info = -1; else if (n < 0) *info = -2; else if (nrhs < 0) *info = -3;
    else if (lda < imax(1, n)) *info = -5; else if (ldb < imax(1, n)) *info = -7;
    if (*info != 0) return;
    if (n == 0 || nrhs == 0) return;
    if (upper) {
        ora_dtrsm('L', 'U', 'T', 'N', n, nrhs, 1.0, a, lda, b, ldb);
        ora_dtrsm('L', 'U', 'N', 'N', n, nrhs, 1.0, a, lda, b, ldb);
    } else {
        ora_dtrsm('L', 'L', 'N', 'N', n, nrhs, 1.0, a, lda, b, ldb);
        ora_dtrsm('L', 'L', 'T', 'N', n, nrhs, 1.0, a, lda, b, ldb);
    }
}

/* SRC/dposv.f:160-183 */
void ora_dposv(char uplo, int n, int nrhs, double *a, int lda, double *b, int ldb, int *info)
{
    *info = 0;
    if (!ora_lsame(uplo, 'U') && !ora_lsame(uplo, 'L')) *info = -1; else if (n < 0) *info = -2;
    else if (nrhs < 0) *info = -3; else if (lda < imax(1, n)) *info = -5; else if (ldb < imax(1, n)) *info = -7;
    if (*info != 0) return;
    ora_dpotrf(uplo, n, a, lda, info);
    if (*info == 0) ora_dpotrs(uplo, n, nrhs, a, lda, b, ldb, info);
}

/* ------------------------------------------------------------------------------------------ */
/* Householder QR                                                                              */

/* SRC/dlarfg.f:140-186 */
void ora_dlarfg(int n, double *alpha, double *x, int incx, double *tau)
{
    if (n <= 1) { *tau = 0.0; return; }
    double xnorm = ora_dnrm2(n - 1, x, incx);
    if (xnorm == 0.0) { *tau = 0.0; return; }
    double beta = -copysign(ora_dlapy2(*alpha, xnorm), *alpha);
    double safmin = ora_dlamch('S') / ora_dlamch('E');
    int knt = 0;
    if (fabs(beta) < safmin) {
        double rsafmn = 1.0 / safmin;
        do {
            knt = knt + 1;
            ora_dscal(n - 1, rsafmn, x, incx);
            beta = beta * rsafmn;
            *alpha = *alpha * rsafmn;
        } while (fabs(beta) < safmin && knt < 20);
        xnorm = ora_dnrm2(n - 1, x, incx);
        beta = -copysign(ora_dlapy2(*alpha, xnorm), *alpha);
    }
    *tau = (beta - *alpha) / beta;
    ora_dscal(n - 1, 1.0 / (*alpha - beta), x, incx);
    for (int j = 0; j < knt; ++j) beta = beta * safmin;
    *alpha = beta;
}

/* SRC/iladlc.f:101-112 -- last non-zero column (1-based count). */
int ora_iladlc(int m, int n, const double *a, int lda)
{
    if (n == 0) return n;
    if (A_(0, n - 1) != 0.0 || A_(m - 1, n - 1) != 0.0) return n;
    for (int j = n; j >= 1; --j)
        for (int i = 0; i < m; ++i)
            if (A_(i, j - 1) != 0.0) return j;
    return 0;
}

/* SRC/iladlr.f:101-114 -- last non-zero row (1-based count). */
int ora_iladlr(int m, int n, const double *a, int lda)
{
    if (m == 0) return m;
    if (A_(m - 1, 0) != 0.0 || A_(m - 1, n - 1) != 0.0) return m;
    int r = 0;
    for (int j = 0; j < n; ++j) {
        int i = m;
        while (i >= 1 && A_(imax(i, 1) - 1, j) == 0.0) i = i - 1;
        r = imax(r, i);
    }
    return r;
}

/* SRC/dlarf1f.f:192-290 -- apply H = I - tau v v**T with v(1)=1 implicit (v[0] is NOT read). */
void ora_dlarf1f(char side, int m, int n, const double *v, int incv, double tau, double *c, int ldc,
                 double *work)
{
    int applyleft = ora_lsame(side, 'L');
    int lastv = 1, lastc = 0;
    ptrdiff_t i = 0;                                         /* 0-based position in v of element LASTV */
    if (tau != 0.0) {
        lastv = applyleft ? m : n;
        i = incv > 0 ? (ptrdiff_t)(lastv - 1) * incv : 0;
        while (lastv > 1 && v[i] == 0.0) { lastv = lastv - 1; i = i - incv; }
        lastc = applyleft ? ora_iladlc(lastv, n, c, ldc) : ora_iladlr(m, lastv, c, ldc);
        if (incv > 0) i = incv;                             /* -> V(2) */
    }
    if (lastc == 0) return;
    if (applyleft) {
        if (lastv == 1) {
            ora_dscal(lastc, 1.0 - tau, c, ldc);
        } else {
            ora_dgemv('T', lastv - 1, lastc, 1.0, &C_(1, 0), ldc, &v[i], incv, 0.0, work, 1);
            ora_daxpy(lastc, 1.0, c, ldc, work, 1);
            ora_daxpy(lastc, -tau, work, 1, c, ldc);
            ora_dger(lastv - 1, lastc, -tau, &v[i], incv, work, 1, &C_(1, 0), ldc);
        }
    } else {
        if (lastv == 1) {
            ora_dscal(lastc, 1.0 - tau, c, 1);
        } else {
            ora_dgemv('N', lastc, lastv - 1, 1.0, &C_(0, 1), ldc, &v[i], incv, 0.0, work, 1);
            ora_daxpy(lastc, 1.0, c, 1, work, 1);
            ora_daxpy(lastc, -tau, work, 1, c, 1);
            ora_dger(lastc, lastv - 1, -tau, work, 1, &v[i], incv, &C_(0, 1), ldc);
        }
    }
}

/* SRC/dgeqr2.f:150-184 */
void ora_dgeqr2(int m, int n, double *a, int lda, double *tau, double *work, int *info)
{
    *info = 0;
    if (m < 0) *info = -1; else if (n < 0) *info = -2; else if (lda < imax(1, m)) *info = -4;
    if (*info != 0) return;
    int k = imin(m, n);
    for (int i = 0; i < k; ++i) {
        ora_dlarfg(m - i, &A_(i, i), &A_(imin(i + 1, m - 1), i), 1, &tau[i]);
        if (i < n - 1) ora_dlarf1f('L', m - i, n - i - 1, &A_(i, i), 1, tau[i], &A_(i, i + 1), lda, work);
    }
}

/* SRC/dlarft_lvl2.f:199-258 -- DIRECT='F', STOREV='C' only (the QR case). */
void ora_dlarft_lvl2(char direct, char storev, int n, int k, const double *v, int ldv, const double *tau,
                     double *t, int ldt)
{
    (void)direct; (void)storev;
    if (n == 0) return;
    int prevlastv = n;                                      /* 1-based like the reference */
    for (int i = 1; i <= k; ++i) {
        prevlastv = imax(i, prevlastv);
        if (tau[i - 1] == 0.0) {
            for (int j = 1; j <= i; ++j) T_(j - 1, i - 1) = 0.0;
        } else {
            int lastv;
            for (lastv = n; lastv >= i + 1; --lastv)
                if (V_(lastv - 1, i - 1) != 0.0) break;
            for (int j = 1; j <= i - 1; ++j) T_(j - 1, i - 1) = -tau[i - 1] * V_(i - 1, j - 1);
            int j = imin(lastv, prevlastv);
            ora_dgemv('T', j - i, i - 1, -tau[i - 1], &V_(i, 0), ldv, &V_(i, i - 1), 1, 1.0, &T_(0, i - 1), 1);
            ora_dtrmv('U', 'N', 'N', i - 1, t, ldt, &T_(0, i - 1), 1);
            T_(i - 1, i - 1) = tau[i - 1];
            if (i > 1) prevlastv = imax(prevlastv, lastv); else prevlastv = lastv;
        }
    }
}

/* SRC/dlarft.f:207-349 -- recursive compact-WY T, QR case (DIRECT='F', STOREV='C');
 * crossover NX = 64 (SRC/ilaenv.f:679-682). */
void ora_dlarft(char direct, char storev, int n, int k, const double *v, int ldv, const double *tau,
                double *t, int ldt)
{
    if (n == 0 || k == 0) return;
    if (n == 1 || k == 1) { T_(0, 0) = tau[0]; return; }
    const int nx = 64;
    if (k < nx) { ora_dlarft_lvl2(direct, storev, n, k, v, ldv, tau, t, ldt); return; }
    int l = k / 2;
    ora_dlarft(direct, storev, n, l, v, ldv, tau, t, ldt);
    ora_dlarft(direct, storev, n - l, k - l, &V_(l, l), ldv, tau + l, &T_(l, l), ldt);
    for (int j = 0; j < l; ++j)
        for (int i = 0; i < k - l; ++i) T_(j, l + i) = V_(l + i, j);
    ora_dtrmm('R', 'L', 'N', 'U', l, k - l, 1.0, &V_(l, l), ldv, &T_(0, l), ldt);
    ora_dgemm('T', 'N', l, k - l, n - k, 1.0, &V_(k, 0), ldv, &V_(k, l), ldv, 1.0, &T_(0, l), ldt);
    ora_dtrmm('L', 'U', 'N', 'N', l, k - l, -1.0, t, ldt, &T_(0, l), ldt);
    ora_dtrmm('R', 'U', 'N', 'N', l, k - l, 1.0, &T_(l, l), ldt, &T_(0, l), ldt);
}

/* SRC/dlarfb.f:231-345 -- DIRECT='F', STOREV='C'; SIDE L or R; TRANS N or T. */
void ora_dlarfb(char side, char trans, char direct, char storev, int m, int n, int k, const double *v, int ldv,
                const double *t, int ldt, double *c, int ldc, double *work, int ldwork)
{
    (void)direct; (void)storev;
    if (m <= 0 || n <= 0) return;
    char transt = ora_lsame(trans, 'N') ? 'T' : 'N';
    if (ora_lsame(side, 'L')) {
        /* W := C**T V = (C1**T V1 + C2**T V2) : dlarfb.f:257-275 */
        for (int j = 0; j < k; ++j) ora_dcopy(n, &C_(j, 0), ldc, &W_(0, j), 1);
        ora_dtrmm('R', 'L', 'N', 'U', n, k, 1.0, v, ldv, work, ldwork);
        if (m > k) ora_dgemm('T', 'N', n, k, m - k, 1.0, &C_(k, 0), ldc, &V_(k, 0), ldv, 1.0, work, ldwork);
        ora_dtrmm('R', 'U', transt, 'N', n, k, 1.0, t, ldt, work, ldwork);
        if (m > k) ora_dgemm('N', 'T', m - k, n, k, -1.0, &V_(k, 0), ldv, work, ldwork, 1.0, &C_(k, 0), ldc);
        ora_dtrmm('R', 'L', 'T', 'U', n, k, 1.0, v, ldv, work, ldwork);
        for (int j = 0; j < k; ++j)
            for (int i = 0; i < n; ++i) C_(j, i) = C_(j, i) - W_(i, j);
    } else {
        /* W := C V : dlarfb.f:307-345 */
        for (int j = 0; j < k; ++j) ora_dcopy(m, &C_(0, j), 1, &W_(0, j), 1);
        ora_dtrmm('R', 'L', 'N', 'U', m, k, 1.0, v, ldv, work, ldwork);
        if (n > k) ora_dgemm('N', 'N', m, k, n - k, 1.0, &C_(0, k), ldc, &V_(k, 0), ldv, 1.0, work, ldwork);
        ora_dtrmm('R', 'U', trans, 'N', m, k, 1.0, t, ldt, work, ldwork);
        if (n > k) ora_dgemm('N', 'T', m, n - k, k, -1.0, work, ldwork, &V_(k, 0), ldv, 1.0, &C_(0, k), ldc);
        ora_dtrmm('R', 'L', 'T', 'U', m, k, 1.0, v, ldv, work, ldwork);
        for (int j = 0; j < k; ++j)
            for (int i = 0; i < m; ++i) C_(i, j) = C_(i, j) - W_(i, j);
    }
}

/* SRC/dgeqrf.f:180-278.  work must hold max(1,n)*nb doubles (lwork is honoured like the reference). */
void ora_dgeqrf(int m, int n, double *a, int lda, double *tau, double *work, int lwork, int *info)
{
    int k = imin(m, n);
    *info = 0;
    int nb = ora_ilaenv_nb("DGEQRF");
    int lquery = (lwork == -1);
    if (m < 0) *info = -1; else if (n < 0) *info = -2; else if (lda < imax(1, m)) *info = -4;
    else if (!lquery) { if (lwork <= 0 || (m > 0 && lwork < imax(1, n))) *info = -7; }
    if (*info != 0) return;
    if (lquery) { work[0] = (k == 0) ? 1.0 : (double)n * nb; return; }
    if (k == 0) { work[0] = 1.0; return; }
    int nbmin = 2, nx = 0, iws = n, ldwork = n, i = 0, iinfo;
    if (nb > 1 && nb < k) {
        nx = imax(0, g_nx_geqrf);
        if (nx < k) {
            ldwork = n;
            iws = ldwork * nb;
            if (lwork < iws) { nb = lwork / ldwork; nbmin = 2; }
        }
    }
    if (nb >= nbmin && nb < k && nx < k) {
        for (i = 0; i < k - nx; i += nb) {
            int ib = imin(k - i, nb);
            ora_dgeqr2(m - i, ib, &A_(i, i), lda, &tau[i], work, &iinfo);
            if (i + ib < n) {
                ora_dlarft('F', 'C', m - i, ib, &A_(i, i), lda, &tau[i], work, ldwork);
                ora_dlarfb('L', 'T', 'F', 'C', m - i, n - i - ib, ib, &A_(i, i), lda, work, ldwork,
                           &A_(i, i + ib), lda, work + ib, ldwork);
            }
        }
    } else i = 0;
    if (i < k) ora_dgeqr2(m - i, n - i, &A_(i, i), lda, &tau[i], work, &iinfo);
    work[0] = (double)iws;
}

/* SRC/dorg2r.f:130-168 */
void ora_dorg2r(int m, int n, int k, double *a, int lda, const double *tau, double *work, int *info)
{
    *info = 0;
    if (m < 0) *info = -1; else if (n < 0 || n > m) *info = -2; else if (k < 0 || k > n) *info = -3;
    else if (lda < imax(1, m)) *info = -5;
    if (*info != 0) return;
    if (n <= 0) return;
    for (int j = k; j < n; ++j) {
        for (int l = 0; l < m; ++l) A_(l, j) = 0.0;
        A_(j, j) = 1.0;
    }
    for (int i = k - 1; i >= 0; --i) {
        if (i < n - 1) ora_dlarf1f('L', m - i, n - i - 1, &A_(i, i), 1, tau[i], &A_(i, i + 1), lda, work);
        if (i < m - 1) ora_dscal(m - i - 1, -tau[i], &A_(i + 1, i), 1);
        A_(i, i) = 1.0 - tau[i];
        for (int l = 0; l < i; ++l) A_(l, i) = 0.0;
    }
}

/* SRC/dorm2r.f:195-271 -- unblocked application of Q or Q**T from DGEQRF (work: n if SIDE='L', m if 'R'). */
void ora_dorm2r(char side, char trans, int m, int n, int k, const double *a, int lda, const double *tau, double *c, int ldc,
                double *work, int *info)
{
    int left = ora_lsame(side, 'L'), notran = ora_lsame(trans, 'N');
    int nq = left ? m : n;
    *info = 0;
    if (!left && !ora_lsame(side, 'R')) *info = -1; else if (!notran && !ora_lsame(trans, 'T')) *info = -2;
    else if (m < 0) *info = -3; else if (n < 0) *info = -4; else if (k < 0 || k > nq) *info = -5;
    else if (lda < imax(1, nq)) *info = -7; else if (ldc < imax(1, m)) *info = -10;
    if (*info != 0) return;
    if (m == 0 || n == 0 || k == 0) return;
    int forward = (left && !notran) || (!left && notran);          /* dorm2r.f:226-235 */
    for (int t = 0; t < k; ++t) {
        int i = forward ? t : k - 1 - t;                           /* 0-based reflector index */
        if (left) ora_dlarf1f('L', m - i, n, &A_(i, i), 1, tau[i], &c[i], ldc, work);                  /* C(i:m,1:n) */
        else ora_dlarf1f('R', m, n - i, &A_(i, i), 1, tau[i], &c[(size_t)i * ldc], ldc, work);         /* C(1:m,i:n) */
    }
}

/* SRC/dormqr.f:203-336 (NB = min(64, ILAENV) = 32, LDT = 65; blocked DLARFT + DLARFB, DORM2R when NB >= K). */
void ora_dormqr(char side, char trans, int m, int n, int k, const double *a, int lda, const double *tau, double *c, int ldc,
                double *work, int lwork, int *info)
{
    const int nbmax = 64, ldt = nbmax + 1;
    int left = ora_lsame(side, 'L'), notran = ora_lsame(trans, 'N');
    int lquery = (lwork == -1);
    int nq = left ? m : n, nw = left ? imax(1, n) : imax(1, m);
    *info = 0;
    if (!left && !ora_lsame(side, 'R')) *info = -1; else if (!notran && !ora_lsame(trans, 'T')) *info = -2;
    else if (m < 0) *info = -3; else if (n < 0) *info = -4; else if (k < 0 || k > nq) *info = -5;
    else if (lda < imax(1, nq)) *info = -7; else if (ldc < imax(1, m)) *info = -10;
    else if (lwork < nw && !lquery) *info = -12;
    int nb = imin(nbmax, ora_ilaenv_nb("DORMQR"));
    int lwkopt = nw * nb + ldt * nb;
    if (*info == 0) work[0] = (double)lwkopt;
    if (*info != 0 || lquery) return;
    if (m == 0 || n == 0 || k == 0) { work[0] = 1.0; return; }
    int nbmin = 2, ldwork = nw, iinfo;
    if (nb > 1 && nb < k && lwork < lwkopt) { nb = lwork / (ldwork + ldt); nbmin = 2; }
    if (nb < nbmin || nb >= k) {
        ora_dorm2r(side, trans, m, n, k, a, lda, tau, c, ldc, work, &iinfo);
    } else {
        double *tw = work + (size_t)nw * nb;                       /* IWT = 1 + NW*NB */
        int forward = (left && !notran) || (!left && notran);
        int nblk = (k + nb - 1) / nb;
        for (int b = 0; b < nblk; ++b) {
            int i = forward ? b * nb : (nblk - 1 - b) * nb;
            int ib = imin(nb, k - i);
            ora_dlarft('F', 'C', nq - i, ib, &A_(i, i), lda, tau + i, tw, ldt);
            if (left) ora_dlarfb('L', trans, 'F', 'C', m - i, n, ib, &A_(i, i), lda, tw, ldt, &c[i], ldc, work, ldwork);
            else ora_dlarfb('R', trans, 'F', 'C', m, n - i, ib, &A_(i, i), lda, tw, ldt, &c[(size_t)i * ldc], ldc, work, ldwork);
        }
    }
    work[0] = (double)lwkopt;
}

/* SRC/dorgqr.f:160-277 (NB=32, NX=128 from ilaenv.f for xORGQR; work >= n*nb). */
void ora_dorgqr(int m, int n, int k, double *a, int lda, const double *tau, double *work, int lwork, int *info)
{
    *info = 0;
    int nb = ora_ilaenv_nb("DORGQR");
    int lquery = (lwork == -1);
    if (m < 0) *info = -1; else if (n < 0 || n > m) *info = -2; else if (k < 0 || k > n) *info = -3;
    else if (lda < imax(1, m)) *info = -5; else if (lwork < imax(1, n) && !lquery) *info = -8;
    if (*info != 0) return;
    if (lquery) { work[0] = (double)(imax(1, n) * nb); return; }
    if (n <= 0) { work[0] = 1.0; return; }
    int nbmin = 2, nx = 0, iws = n, ldwork = n, ki = 0, kk, iinfo;
    if (nb > 1 && nb < k) {
        nx = imax(0, g_nx_geqrf);
        if (nx < k) {
            ldwork = n;
            iws = ldwork * nb;
            if (lwork < iws) { nb = lwork / ldwork; nbmin = 2; }
        }
    }
    if (nb >= nbmin && nb < k && nx < k) {
        ki = ((k - nx - 1) / nb) * nb;
        kk = imin(k, ki + nb);
        for (int j = kk; j < n; ++j) for (int i = 0; i < kk; ++i) A_(i, j) = 0.0;
    } else kk = 0;
    if (kk < n) ora_dorg2r(m - kk, n - kk, k - kk, &A_(kk, kk), lda, tau + kk, work, &iinfo);
    if (kk > 0) {
        for (int i = ki; i >= 0; i -= nb) {                 /* i 0-based; reference I = i+1 */
            int ib = imin(nb, k - i);
            if (i + ib < n) {
                ora_dlarft('F', 'C', m - i, ib, &A_(i, i), lda, tau + i, work, ldwork);
                ora_dlarfb('L', 'N', 'F', 'C', m - i, n - i - ib, ib, &A_(i, i), lda, work, ldwork,
                           &A_(i, i + ib), lda, work + ib, ldwork);
            }
            ora_dorg2r(m - i, ib, ib, &A_(i, i), lda, tau + i, work, &iinfo);
            for (int j = i; j < i + ib; ++j) for (int l = 0; l < i; ++l) A_(l, j) = 0.0;
        }
    }
    work[0] = (double)iws;
}

/* SRC/dtrti2.f:150-213 -- unblocked inverse of a triangular matrix. */
void ora_dtrti2(char uplo, char diag, int n, double *a, int lda, int *info)
{
    int upper = ora_lsame(uplo, 'U'), nounit = ora_lsame(diag, 'N');
    *info = 0;
    if (!upper && !ora_lsame(uplo, 'L')) *info = -1; else if (!nounit && !ora_lsame(diag, 'U')) *info = -2;
    else if (n < 0) *info = -3; else if (lda < imax(1, n)) *info = -5;
    if (*info != 0) return;
    if (upper) {
        for (int j = 0; j < n; ++j) {
            double ajj;
            if (nounit) { A_(j, j) = 1.0 / A_(j, j); ajj = -A_(j, j); } else ajj = -1.0;
            ora_dtrmv('U', 'N', diag, j, a, lda, &A_(0, j), 1);               /* elements 1:j-1 of column j */
            ora_dscal(j, ajj, &A_(0, j), 1);
        }
    } else {
        for (int j = n - 1; j >= 0; --j) {
            double ajj;
            if (nounit) { A_(j, j) = 1.0 / A_(j, j); ajj = -A_(j, j); } else ajj = -1.0;
            if (j < n - 1) {
                ora_dtrmv('L', 'N', diag, n - 1 - j, &A_(j + 1, j + 1), lda, &A_(j + 1, j), 1);
                ora_dscal(n - 1 - j, ajj, &A_(j + 1, j), 1);
            }
        }
    }
}

/* SRC/dtrtri.f:150-243 (NB = 64, ilaenv.f:475-481); INFO = i if A(i,i) is exactly zero (nothing is computed then). */
void ora_dtrtri(char uplo, char diag, int n, double *a, int lda, int *info)
{
    int upper = ora_lsame(uplo, 'U'), nounit = ora_lsame(diag, 'N');
    *info = 0;
    if (!upper && !ora_lsame(uplo, 'L')) *info = -1; else if (!nounit && !ora_lsame(diag, 'U')) *info = -2;
    else if (n < 0) *info = -3; else if (lda < imax(1, n)) *info = -5;
    if (*info != 0 || n == 0) return;
    if (nounit) {
        for (int i = 0; i < n; ++i) if (A_(i, i) == 0.0) { *info = i + 1; return; }
    }
    int nb = ora_ilaenv_nb("DTRTRI");
    if (nb <= 1 || nb >= n) { ora_dtrti2(uplo, diag, n, a, lda, info); return; }
    if (upper) {
        for (int j = 0; j < n; j += nb) {
            int jb = imin(nb, n - j);
            ora_dtrmm('L', 'U', 'N', diag, j, jb, 1.0, a, lda, &A_(0, j), lda);
            ora_dtrsm('R', 'U', 'N', diag, j, jb, -1.0, &A_(j, j), lda, &A_(0, j), lda);
            ora_dtrti2('U', diag, jb, &A_(j, j), lda, info);
        }
    } else {
        int nn = ((n - 1) / nb) * nb;                                       /* 0-based start of the last block */
        for (int j = nn; j >= 0; j -= nb) {
            int jb = imin(nb, n - j);
            if (j + jb < n) {
                ora_dtrmm('L', 'L', 'N', diag, n - j - jb, jb, 1.0, &A_(j + jb, j + jb), lda, &A_(j + jb, j), lda);
                ora_dtrsm('R', 'L', 'N', diag, n - j - jb, jb, -1.0, &A_(j, j), lda, &A_(j + jb, j), lda);
            }
            ora_dtrti2('L', diag, jb, &A_(j, j), lda, info);
        }
    }
}

/* SRC/dgetri.f:150-259 -- inverse from the LU factors: inv(U), then inv(A)*L = inv(U), then column interchanges.
   work: at least n*nb doubles for the blocked path (lwork honoured like the reference). */
void ora_dgetri(int n, double *a, int lda, const int *ipiv, double *work, int lwork, int *info)
{
    *info = 0;
    int nb = ora_ilaenv_nb("DGETRI");
    int lquery = (lwork == -1);
    work[0] = (double)imax(1, n * nb);
    if (n < 0) *info = -1; else if (lda < imax(1, n)) *info = -3; else if (lwork < imax(1, n) && !lquery) *info = -6;
    if (*info != 0 || lquery) return;
    if (n == 0) return;
    ora_dtrtri('U', 'N', n, a, lda, info);
    if (*info > 0) return;
    int nbmin = 2, ldwork = n, iws;
    if (nb > 1 && nb < n) {
        iws = imax(ldwork * nb, 1);
        if (lwork < iws) { nb = lwork / ldwork; nbmin = 2; }
    } else iws = n;
    if (nb < nbmin || nb >= n) {
        for (int j = n - 1; j >= 0; --j) {
            for (int i = j + 1; i < n; ++i) { work[i] = A_(i, j); A_(i, j) = 0.0; }
            if (j < n - 1) ora_dgemv('N', n, n - 1 - j, -1.0, &A_(0, j + 1), lda, &work[j + 1], 1, 1.0, &A_(0, j), 1);
        }
    } else {
        int nn = ((n - 1) / nb) * nb;
        for (int j = nn; j >= 0; j -= nb) {
            int jb = imin(nb, n - j);
            for (int jj = j; jj < j + jb; ++jj)
                for (int i = jj + 1; i < n; ++i) { work[i + (size_t)(jj - j) * ldwork] = A_(i, jj); A_(i, jj) = 0.0; }
            if (j + jb < n)
                ora_dgemm('N', 'N', n, jb, n - j - jb, -1.0, &A_(0, j + jb), lda, &work[j + jb], ldwork, 1.0, &A_(0, j), lda);
            ora_dtrsm('R', 'L', 'N', 'U', n, jb, 1.0, &work[j], ldwork, &A_(0, j), lda);
        }
    }
    for (int j = n - 2; j >= 0; --j) {
        int jp = ipiv[j] - 1;
        if (jp != j) ora_dswap(n, &A_(0, j), 1, &A_(0, jp), 1);
    }
    work[0] = (double)iws;
}

/* SRC/dgeqrt3.f:157-250 -- recursive QR of an m x n (m >= n) block with the compact-WY factor T (n x n upper). */
void ora_dgeqrt3(int m, int n, double *a, int lda, double *t, int ldt, int *info)
{
    *info = 0;
    if (n < 0) *info = -2; else if (m < n) *info = -1; else if (lda < imax(1, m)) *info = -4; else if (ldt < imax(1, n)) *info = -6;
    if (*info != 0) return;
    if (n == 1) {
        ora_dlarfg(m, &A_(0, 0), &A_(imin(1, m - 1), 0), 1, &T_(0, 0));
        return;
    }
    int n1 = n / 2, n2 = n - n1, j1 = imin(n1, n - 1), i1 = imin(n, m - 1), iinfo;     /* 0-based J1, I1 */
    ora_dgeqrt3(m, n1, a, lda, t, ldt, &iinfo);
    /* A(1:M,J1:N) = Q1^T A(1:M,J1:N), workspace T(1:N1,J1:N) */
    for (int j = 0; j < n2; ++j) for (int i = 0; i < n1; ++i) T_(i, j + n1) = A_(i, j + n1);
    ora_dtrmm('L', 'L', 'T', 'U', n1, n2, 1.0, a, lda, &T_(0, j1), ldt);
    ora_dgemm('T', 'N', n1, n2, m - n1, 1.0, &A_(j1, 0), lda, &A_(j1, j1), lda, 1.0, &T_(0, j1), ldt);
    ora_dtrmm('L', 'U', 'T', 'N', n1, n2, 1.0, t, ldt, &T_(0, j1), ldt);
    ora_dgemm('N', 'N', m - n1, n2, n1, -1.0, &A_(j1, 0), lda, &T_(0, j1), ldt, 1.0, &A_(j1, j1), lda);
    ora_dtrmm('L', 'L', 'N', 'U', n1, n2, 1.0, a, lda, &T_(0, j1), ldt);
    for (int j = 0; j < n2; ++j) for (int i = 0; i < n1; ++i) A_(i, j + n1) = A_(i, j + n1) - T_(i, j + n1);
    ora_dgeqrt3(m - n1, n2, &A_(j1, j1), lda, &T_(j1, j1), ldt, &iinfo);
    /* T3 = T(1:N1,J1:N) = -T1 Y1^T Y2 T2 */
    for (int i = 0; i < n1; ++i) for (int j = 0; j < n2; ++j) T_(i, j + n1) = A_(j + n1, i);
    ora_dtrmm('R', 'L', 'N', 'U', n1, n2, 1.0, &A_(j1, j1), lda, &T_(0, j1), ldt);
    ora_dgemm('T', 'N', n1, n2, m - n, 1.0, &A_(i1, 0), lda, &A_(i1, j1), lda, 1.0, &T_(0, j1), ldt);
    ora_dtrmm('L', 'U', 'N', 'N', n1, n2, -1.0, t, ldt, &T_(0, j1), ldt);
    ora_dtrmm('R', 'U', 'N', 'N', n1, n2, 1.0, &T_(j1, j1), ldt, &T_(0, j1), ldt);
}

/* SRC/dgeqrt.f:166-211 (USE_RECURSIVE_QR = .TRUE.): blocked QR keeping the T factors, T is nb x min(m,n). work: nb*n. */
void ora_dgeqrt(int m, int n, int nb, double *a, int lda, double *t, int ldt, double *work, int *info)
{
    *info = 0;
    if (m < 0) *info = -1; else if (n < 0) *info = -2;
    else if (nb < 1 || (nb > imin(m, n) && imin(m, n) > 0)) *info = -3;
    else if (lda < imax(1, m)) *info = -5; else if (ldt < nb) *info = -7;
    if (*info != 0) return;
    int k = imin(m, n), iinfo;
    if (k == 0) return;
    for (int i = 0; i < k; i += nb) {
        int ib = imin(k - i, nb);
        ora_dgeqrt3(m - i, ib, &A_(i, i), lda, &T_(0, i), ldt, &iinfo);
        if (i + ib < n)
            ora_dlarfb('L', 'T', 'F', 'C', m - i, n - i - ib, ib, &A_(i, i), lda, &T_(0, i), ldt, &A_(i, i + ib), lda, work,
                       n - i - ib);
    }
}

/* SRC/dgemqrt.f:199-287 -- apply Q or Q^T from DGEQRT.  work: n*nb (SIDE='L') or m*nb ('R'). */
void ora_dgemqrt(char side, char trans, int m, int n, int k, int nb, const double *v, int ldv, const double *t, int ldt,
                 double *c, int ldc, double *work, int *info)
{
    int left = ora_lsame(side, 'L'), right = ora_lsame(side, 'R'), tran = ora_lsame(trans, 'T'), notran = ora_lsame(trans, 'N');
    int ldwork = left ? imax(1, n) : imax(1, m), q = left ? m : n;
    *info = 0;
    if (!left && !right) *info = -1; else if (!tran && !notran) *info = -2; else if (m < 0) *info = -3; else if (n < 0) *info = -4;
    else if (k < 0 || k > q) *info = -5; else if (nb < 1 || (nb > k && k > 0)) *info = -6;
    else if (ldv < imax(1, q)) *info = -8; else if (ldt < nb) *info = -10; else if (ldc < imax(1, m)) *info = -12;
    if (*info != 0) return;
    if (m == 0 || n == 0 || k == 0) return;
    int forward = (left && tran) || (right && notran);
    int nblk = (k + nb - 1) / nb;
    for (int b = 0; b < nblk; ++b) {
        int i = forward ? b * nb : (nblk - 1 - b) * nb;
        int ib = imin(nb, k - i);
        const double *vi = v + (size_t)i + (size_t)i * ldv, *ti = t + (size_t)i * ldt;
        if (left) ora_dlarfb('L', trans, 'F', 'C', m - i, n, ib, vi, ldv, ti, ldt, &c[i], ldc, work, ldwork);
        else ora_dlarfb('R', trans, 'F', 'C', m, n - i, ib, vi, ldv, ti, ldt, &c[(size_t)i * ldc], ldc, work, ldwork);
    }
}

/* SRC/dlascl.f:233-285, TYPE = 'G' only: A := A * (cto/cfrom) without over/underflow (stepwise by SMLNUM / BIGNUM). */
void ora_dlascl_g(double cfrom, double cto, int m, int n, double *a, int lda)
{
    if (m == 0 || n == 0) return;
    const double smlnum = 2.2250738585072014e-308, bignum = 1.0 / smlnum;       /* DLAMCH('S') */
    double cfromc = cfrom, ctoc = cto;
    for (;;) {
        double cfrom1 = cfromc * smlnum, mul;
        int done;
        if (cfrom1 == cfromc) { mul = ctoc / cfromc; done = 1; }
        else {
            double cto1 = ctoc / bignum;
            if (cto1 == ctoc) { mul = ctoc; done = 1; cfromc = 1.0; }
            else if (fabs(cfrom1) > fabs(ctoc) && ctoc != 0.0) { mul = smlnum; done = 0; cfromc = cfrom1; }
            else if (fabs(cto1) > fabs(cfromc)) { mul = bignum; done = 0; ctoc = cto1; }
            else { mul = ctoc / cfromc; done = 1; if (mul == 1.0) return; }
        }
        for (int j = 0; j < n; ++j) for (int i = 0; i < m; ++i) A_(i, j) = A_(i, j) * mul;
        if (done) break;
    }
}

/* SRC/dtrtrs.f:181-224: INFO = i if A(i,i) is exactly zero (non-unit), otherwise one DTRSM. */
void ora_dtrtrs(char uplo, char trans, char diag, int n, int nrhs, const double *a, int lda, double *b, int ldb, int *info)
{
    *info = 0;
    if (n == 0) return;
    if (ora_lsame(diag, 'N'))
        for (int i = 0; i < n; ++i) if (A_(i, i) == 0.0) { *info = i + 1; return; }
    ora_dtrsm('L', uplo, trans, diag, n, nrhs, 1.0, a, lda, b, ldb);
}

/* SRC/dgelq2.f:140-188 -- unblocked LQ.  (SRC/dgelqf.f switches to the blocked DLARFT/DLARFB 'Rowwise' form for
   k > NX = 128; the factorization is the same up to rounding, so the oracle uses this form at every size.) */
void ora_dgelq2(int m, int n, double *a, int lda, double *tau, double *work, int *info)
{
    *info = 0;
    if (m < 0) *info = -1; else if (n < 0) *info = -2; else if (lda < imax(1, m)) *info = -4;
    if (*info != 0) return;
    int k = imin(m, n);
    for (int i = 0; i < k; ++i) {
        ora_dlarfg(n - i, &A_(i, i), &A_(i, imin(i + 1, n - 1)), lda, &tau[i]);
        if (i < m - 1) ora_dlarf1f('R', m - i - 1, n - i, &A_(i, i), lda, tau[i], &A_(i + 1, i), lda, work);
    }
}

/* SRC/dorml2.f:215-266 -- apply Q or Q^T from DGELQF, Q = H(k) ... H(1), reflectors stored in the rows of A. */
void ora_dorml2(char side, char trans, int m, int n, int k, const double *a, int lda, const double *tau, double *c, int ldc,
                double *work, int *info)
{
    int left = ora_lsame(side, 'L'), notran = ora_lsame(trans, 'N');
    int nq = left ? m : n;
    *info = 0;
    if (!left && !ora_lsame(side, 'R')) *info = -1; else if (!notran && !ora_lsame(trans, 'T')) *info = -2;
    else if (m < 0) *info = -3; else if (n < 0) *info = -4; else if (k < 0 || k > nq) *info = -5;
    else if (lda < imax(1, k)) *info = -7; else if (ldc < imax(1, m)) *info = -10;
    if (*info != 0) return;
    if (m == 0 || n == 0 || k == 0) return;
    int forward = (left && notran) || (!left && !notran);
    for (int t = 0; t < k; ++t) {
        int i = forward ? t : k - 1 - t;
        if (left) ora_dlarf1f('L', m - i, n, &A_(i, i), lda, tau[i], &c[i], ldc, work);
        else ora_dlarf1f('R', m, n - i, &A_(i, i), lda, tau[i], &c[(size_t)i * ldc], ldc, work);
    }
}

/* SRC/dgels.f:238-514 -- least squares / minimum norm solution with QR (m >= n) or LQ (m < n), full rank assumed.
   work: mn + max(mn, nrhs, n, m) doubles are enough for the unblocked kernels used here. */
void ora_dgels(char trans, int m, int n, int nrhs, double *a, int lda, double *b, int ldb, double *work, int lwork, int *info)
{
    int mn = imin(m, n), lquery = (lwork == -1);
    int tn = ora_lsame(trans, 'N');
    *info = 0;
    if (!tn && !ora_lsame(trans, 'T')) *info = -1; else if (m < 0) *info = -2; else if (n < 0) *info = -3;
    else if (nrhs < 0) *info = -4; else if (lda < imax(1, m)) *info = -6; else if (ldb < imax(1, imax(m, n))) *info = -8;
    else if (lwork < imax(1, mn + imax(mn, nrhs)) && !lquery) *info = -10;
    int nb = 32;                                                       /* ilaenv.f: DGEQRF/DGELQF/DORMQR/DORMLQ */
    int wsize = imax(1, mn + imax(mn, nrhs) * nb);
    if (*info == 0 || *info == -10) work[0] = (double)wsize;
    if (*info != 0 || lquery) return;
    int tpsd = !tn;
#define B_(i, j) b[(size_t)(i) + (size_t)(j) * ldb]
    if (imin(imin(m, n), nrhs) == 0) {
        for (int j = 0; j < nrhs; ++j) for (int i = 0; i < imax(m, n); ++i) B_(i, j) = 0.0;
        return;
    }
    const double smlnum = 2.2250738585072014e-308 / 2.220446049250313e-16, bignum = 1.0 / smlnum;   /* 'S' / 'P' */
    double anrm = ora_dlange('M', m, n, a, lda);
    int iascl = 0, ibscl = 0, scllen = 0, iinfo;
    if (anrm > 0.0 && anrm < smlnum) { ora_dlascl_g(anrm, smlnum, m, n, a, lda); iascl = 1; }
    else if (anrm > bignum) { ora_dlascl_g(anrm, bignum, m, n, a, lda); iascl = 2; }
    else if (anrm == 0.0) {
        for (int j = 0; j < nrhs; ++j) for (int i = 0; i < imax(m, n); ++i) B_(i, j) = 0.0;
        work[0] = (double)wsize;
        return;
    }
    int brow = tpsd ? n : m;
    double bnrm = ora_dlange('M', brow, nrhs, b, ldb);
    if (bnrm > 0.0 && bnrm < smlnum) { ora_dlascl_g(bnrm, smlnum, brow, nrhs, b, ldb); ibscl = 1; }
    else if (bnrm > bignum) { ora_dlascl_g(bnrm, bignum, brow, nrhs, b, ldb); ibscl = 2; }
    double *tau = work, *w2 = work + mn;
    int lw2 = lwork - mn;
    if (m >= n) {
        ora_dgeqrf(m, n, a, lda, tau, w2, lw2, &iinfo);
        if (!tpsd) {
            ora_dormqr('L', 'T', m, nrhs, n, a, lda, tau, b, ldb, w2, lw2, &iinfo);
            ora_dtrtrs('U', 'N', 'N', n, nrhs, a, lda, b, ldb, info);
            if (*info > 0) return;
            scllen = n;
        } else {
            ora_dtrtrs('U', 'T', 'N', n, nrhs, a, lda, b, ldb, info);
            if (*info > 0) return;
            for (int j = 0; j < nrhs; ++j) for (int i = n; i < m; ++i) B_(i, j) = 0.0;
            ora_dormqr('L', 'N', m, nrhs, n, a, lda, tau, b, ldb, w2, lw2, &iinfo);
            scllen = m;
        }
    } else {
        ora_dgelq2(m, n, a, lda, tau, w2, &iinfo);
        if (!tpsd) {
            ora_dtrtrs('L', 'N', 'N', m, nrhs, a, lda, b, ldb, info);
            if (*info > 0) return;
            for (int j = 0; j < nrhs; ++j) for (int i = m; i < n; ++i) B_(i, j) = 0.0;
            ora_dorml2('L', 'T', n, nrhs, m, a, lda, tau, b, ldb, w2, &iinfo);
            scllen = n;
        } else {
            ora_dorml2('L', 'N', n, nrhs, m, a, lda, tau, b, ldb, w2, &iinfo);
            ora_dtrtrs('L', 'T', 'N', m, nrhs, a, lda, b, ldb, info);
            if (*info > 0) return;
            scllen = m;
        }
    }
    if (iascl == 1) ora_dlascl_g(anrm, smlnum, scllen, nrhs, b, ldb);
    else if (iascl == 2) ora_dlascl_g(anrm, bignum, scllen, nrhs, b, ldb);
    if (ibscl == 1) ora_dlascl_g(smlnum, bnrm, scllen, nrhs, b, ldb);
    else if (ibscl == 2) ora_dlascl_g(bignum, bnrm, scllen, nrhs, b, ldb);
#undef B_
    work[0] = (double)wsize;
}

/* SRC/dlacn2.f:166-293 -- Hager/Higham 1-norm estimator, reverse communication (state in isave[3], 1-based labels). */
void ora_dlacn2(int n, double *v, double *x, int *isgn, double *est, int *kase, int *isave)
{
    const int itmax = 5;
    if (*kase == 0) {
        for (int i = 0; i < n; ++i) x[i] = 1.0 / (double)n;
        *kase = 1; isave[0] = 1;
        return;
    }
    int jlast;
    double estold, temp, altsgn;
    switch (isave[0]) {
    case 1:
        if (n == 1) { v[0] = x[0]; *est = fabs(v[0]); goto L150; }
        *est = 0.0;
        for (int i = 0; i < n; ++i) *est += fabs(x[i]);                      /* DASUM */
        for (int i = 0; i < n; ++i) { x[i] = (x[i] >= 0.0) ? 1.0 : -1.0; isgn[i] = (int)x[i]; }
        *kase = 2; isave[0] = 2;
        return;
    case 2:
        isave[1] = ora_idamax(n, x, 1);
        isave[2] = 2;
    L50:
        for (int i = 0; i < n; ++i) x[i] = 0.0;
        x[isave[1] - 1] = 1.0;
        *kase = 1; isave[0] = 3;
        return;
    case 3: {
        int changed = 0;
        ora_dcopy(n, x, 1, v, 1);
        estold = *est;
        *est = 0.0;
        for (int i = 0; i < n; ++i) *est += fabs(v[i]);
        for (int i = 0; i < n; ++i) {
            int xs = (x[i] >= 0.0) ? 1 : -1;
            if (xs != isgn[i]) { changed = 1; break; }
        }
        if (!changed) goto L120;                                             /* repeated sign vector: converged */
        if (*est <= estold) goto L120;
        for (int i = 0; i < n; ++i) { x[i] = (x[i] >= 0.0) ? 1.0 : -1.0; isgn[i] = (int)x[i]; }
        *kase = 2; isave[0] = 4;
        return;
    }
    case 4:
        jlast = isave[1];
        isave[1] = ora_idamax(n, x, 1);
        if (x[jlast - 1] != fabs(x[isave[1] - 1]) && isave[2] < itmax) { isave[2] += 1; goto L50; }
    L120:
        altsgn = 1.0;
        for (int i = 0; i < n; ++i) { x[i] = altsgn * (1.0 + (double)i / (double)(n - 1)); altsgn = -altsgn; }
        *kase = 1; isave[0] = 5;
        return;
    case 5:
        temp = 0.0;
        for (int i = 0; i < n; ++i) temp += fabs(x[i]);
        temp = 2.0 * (temp / (double)(3 * n));
        if (temp > *est) { ora_dcopy(n, x, 1, v, 1); *est = temp; }
        goto L150;
    }
L150:
    *kase = 0;
}

/* SRC/dgerfs.f:235-440 -- iterative refinement of DGETRS solutions with componentwise backward error BERR and forward
   error bound FERR.  work: 3n doubles, iwork: n ints. */
void ora_dgerfs(char trans, int n, int nrhs, const double *a, int lda, const double *af, int ldaf, const int *ipiv,
                const double *b, int ldb, double *x, int ldx, double *ferr, double *berr, double *work, int *iwork, int *info)
{
    const int itmax = 5;
    int notran = ora_lsame(trans, 'N');
    *info = 0;
    if (!notran && !ora_lsame(trans, 'T') && !ora_lsame(trans, 'C')) *info = -1; else if (n < 0) *info = -2;
    else if (nrhs < 0) *info = -3; else if (lda < imax(1, n)) *info = -5; else if (ldaf < imax(1, n)) *info = -7;
    else if (ldb < imax(1, n)) *info = -10; else if (ldx < imax(1, n)) *info = -12;
    if (*info != 0) return;
    if (n == 0 || nrhs == 0) { for (int j = 0; j < nrhs; ++j) { ferr[j] = 0.0; berr[j] = 0.0; } return; }
    char transt = notran ? 'T' : 'N';
    const int nz = n + 1;
    const double eps = 1.1102230246251565e-16, safmin = 2.2250738585072014e-308;   /* DLAMCH('Epsilon'), ('Safe minimum') */
    const double safe1 = nz * safmin, safe2 = safe1 / eps;
    for (int j = 0; j < nrhs; ++j) {
        const double *bj = b + (size_t)j * ldb;
        double *xj = x + (size_t)j * ldx;
        int count = 1, iinfo;
        double lstres = 3.0;
        for (;;) {
            ora_dcopy(n, bj, 1, work + n, 1);
            ora_dgemv(trans, n, n, -1.0, a, lda, xj, 1, 1.0, work + n, 1);
            for (int i = 0; i < n; ++i) work[i] = fabs(bj[i]);
            if (notran) {
                for (int k = 0; k < n; ++k) { double xk = fabs(xj[k]); for (int i = 0; i < n; ++i) work[i] += fabs(A_(i, k)) * xk; }
            } else {
                for (int k = 0; k < n; ++k) { double s = 0.0; for (int i = 0; i < n; ++i) s += fabs(A_(i, k)) * fabs(xj[i]); work[k] += s; }
            }
            double s = 0.0;
            for (int i = 0; i < n; ++i) {
                if (work[i] > safe2) s = fmax(s, fabs(work[n + i]) / work[i]);
                else s = fmax(s, (fabs(work[n + i]) + safe1) / (work[i] + safe1));
            }
            berr[j] = s;
            if (berr[j] > eps && 2.0 * berr[j] <= lstres && count <= itmax) {
                ora_dgetrs(trans, n, 1, af, ldaf, ipiv, work + n, n, &iinfo);
                ora_daxpy(n, 1.0, work + n, 1, xj, 1);
                lstres = berr[j];
                ++count;
                continue;
            }
            break;
        }
        for (int i = 0; i < n; ++i) {
            if (work[i] > safe2) work[i] = fabs(work[n + i]) + nz * eps * work[i];
            else work[i] = fabs(work[n + i]) + nz * eps * work[i] + safe1;
        }
        int kase = 0, isave[3] = {0, 0, 0};
        for (;;) {
            ora_dlacn2(n, work + 2 * n, work + n, iwork, &ferr[j], &kase, isave);
            if (kase == 0) break;
            if (kase == 1) {
                ora_dgetrs(transt, n, 1, af, ldaf, ipiv, work + n, n, &iinfo);
                for (int i = 0; i < n; ++i) work[n + i] = work[i] * work[n + i];
            } else {
                for (int i = 0; i < n; ++i) work[n + i] = work[i] * work[n + i];
                ora_dgetrs(trans, n, 1, af, ldaf, ipiv, work + n, n, &iinfo);
            }
        }
        lstres = 0.0;
        for (int i = 0; i < n; ++i) lstres = fmax(lstres, fabs(xj[i]));
        if (lstres != 0.0) ferr[j] = ferr[j] / lstres;
    }
}

/* ======================================================================================================================
 * Condition estimation and the expert driver: DLATRS, DRSCL, DGECON, DGEEQU, DLAQGE, DGESVX.
 * ====================================================================================================================== */

/* BLAS/SRC/dtrsv.f:180-330 (INCX = 1) */
void ora_dtrsv(char uplo, char trans, char diag, int n, const double *a, int lda, double *x)
{
    const int upper = ora_lsame(uplo, 'U'), notran = ora_lsame(trans, 'N'), nounit = ora_lsame(diag, 'N');
    if (n == 0) return;
    if (notran) {
        if (upper) {
            for (int j = n - 1; j >= 0; --j)
                if (x[j] != 0.0) {
                    if (nounit) x[j] = x[j] / A_(j, j);
                    double temp = x[j];
                    for (int i = j - 1; i >= 0; --i) x[i] = x[i] - temp * A_(i, j);
                }
        } else {
            for (int j = 0; j < n; ++j)
                if (x[j] != 0.0) {
                    if (nounit) x[j] = x[j] / A_(j, j);
                    double temp = x[j];
                    for (int i = j + 1; i < n; ++i) x[i] = x[i] - temp * A_(i, j);
                }
        }
    } else {
        if (upper) {
            for (int j = 0; j < n; ++j) {
                double temp = x[j];
                for (int i = 0; i < j; ++i) temp = temp - A_(i, j) * x[i];
                if (nounit) temp = temp / A_(j, j);
                x[j] = temp;
            }
        } else {
            for (int j = n - 1; j >= 0; --j) {
                double temp = x[j];
                for (int i = n - 1; i > j; --i) temp = temp - A_(i, j) * x[i];
                if (nounit) temp = temp / A_(j, j);
                x[j] = temp;
            }
        }
    }
}

/* SRC/drscl.f:120-170: x := x / sa without overflow / underflow of the reciprocal */
void ora_drscl(int n, double sa, double *x)
{
    if (n <= 0) return;
    const double smlnum = ora_dlamch('S'), bignum = 1.0 / smlnum;
    double cden = sa, cnum = 1.0, mul;
    int done;
    do {
        double cden1 = cden * smlnum, cnum1 = cnum / bignum;
        if (fabs(cden1) > fabs(cnum) && cnum != 0.0) { mul = smlnum; done = 0; cden = cden1; }
        else if (fabs(cnum1) > fabs(cden)) { mul = bignum; done = 0; cnum = cnum1; }
        else { mul = cnum / cden; done = 1; }
        ora_dscal(n, mul, x, 1);
    } while (!done);
}

static double asum_(int n, const double *x) { double s = 0.0; for (int i = 0; i < n; ++i) s += fabs(x[i]); return s; }

/* SRC/dlatrs.f:250-850 -- triangular solve with scaling to prevent overflow: A x = scale b or A^T x = scale b. */
void ora_dlatrs(char uplo, char trans, char diag, char normin, int n, const double *a, int lda, double *x, double *scale,
                double *cnorm, int *info)
{
    const int upper = ora_lsame(uplo, 'U'), notran = ora_lsame(trans, 'N'), nounit = ora_lsame(diag, 'N');
    *info = 0;
    if (!upper && !ora_lsame(uplo, 'L')) *info = -1;
    else if (!notran && !ora_lsame(trans, 'T') && !ora_lsame(trans, 'C')) *info = -2;
    else if (!nounit && !ora_lsame(diag, 'U')) *info = -3;
    else if (!ora_lsame(normin, 'Y') && !ora_lsame(normin, 'N')) *info = -4;
    else if (n < 0) *info = -5;
    else if (lda < imax(1, n)) *info = -7;
    if (*info != 0) return;
    *scale = 1.0;
    if (n == 0) return;
    const double ovfl = ora_dlamch('O');
    const double smlnum = ora_dlamch('S') / ora_dlamch('P'), bignum = 1.0 / smlnum;
    if (ora_lsame(normin, 'N')) {                                            /* dlatrs.f:300-317 */
        if (upper) for (int j = 0; j < n; ++j) cnorm[j] = asum_(j, &A_(0, j));
        else { for (int j = 0; j < n - 1; ++j) cnorm[j] = asum_(n - j - 1, &A_(j + 1, j)); cnorm[n - 1] = 0.0; }
    }
    int imaxi = ora_idamax(n, cnorm, 1);
    double tmax = cnorm[imaxi - 1], tscal;
    if (tmax <= bignum) tscal = 1.0;
    else if (tmax <= ovfl) { tscal = 1.0 / (smlnum * tmax); ora_dscal(n, tscal, cnorm, 1); }
    else {                                                                   /* dlatrs.f:338-392 */
        tmax = 0.0;
        if (upper) { for (int j = 1; j < n; ++j) for (int i = 0; i < j; ++i) tmax = fmax(fabs(A_(i, j)), tmax); }
        else { for (int j = 0; j < n - 1; ++j) for (int i = j + 1; i < n; ++i) tmax = fmax(fabs(A_(i, j)), tmax); }
        if (tmax <= ovfl) {
            tscal = 1.0 / (smlnum * tmax);
            for (int j = 0; j < n; ++j) {
                if (cnorm[j] <= ovfl) cnorm[j] = cnorm[j] * tscal;
                else {
                    cnorm[j] = 0.0;
                    if (upper) for (int i = 0; i < j; ++i) cnorm[j] += tscal * fabs(A_(i, j));
                    else for (int i = j + 1; i < n; ++i) cnorm[j] += tscal * fabs(A_(i, j));
                }
            }
        } else { ora_dtrsv(uplo, trans, diag, n, a, lda, x); return; }
    }
    int j = ora_idamax(n, x, 1);
    double xmax = fabs(x[j - 1]), xbnd = xmax, grow;
    int jfirst, jlast, jinc;
    if (notran) {
        if (upper) { jfirst = n - 1; jlast = 0; jinc = -1; } else { jfirst = 0; jlast = n - 1; jinc = 1; }
        if (tscal != 1.0) grow = 0.0;
        else if (nounit) {
            grow = 1.0 / fmax(xbnd, smlnum);
            xbnd = grow;
            int broke = 0;
            for (j = jfirst; jinc > 0 ? j <= jlast : j >= jlast; j += jinc) {
                if (grow <= smlnum) { broke = 1; break; }
                double tjj = fabs(A_(j, j));
                xbnd = fmin(xbnd, fmin(1.0, tjj) * grow);
                if (tjj + cnorm[j] >= smlnum) grow = grow * (tjj / (tjj + cnorm[j]));
                else grow = 0.0;
            }
            if (!broke) grow = xbnd;
        } else {
            grow = fmin(1.0, 1.0 / fmax(xbnd, smlnum));
            for (j = jfirst; jinc > 0 ? j <= jlast : j >= jlast; j += jinc) {
                if (grow <= smlnum) break;
                grow = grow * (1.0 / (1.0 + cnorm[j]));
            }
        }
    } else {
        if (upper) { jfirst = 0; jlast = n - 1; jinc = 1; } else { jfirst = n - 1; jlast = 0; jinc = -1; }
        if (tscal != 1.0) grow = 0.0;
        else if (nounit) {
            grow = 1.0 / fmax(xbnd, smlnum);
            xbnd = grow;
            int broke = 0;
            for (j = jfirst; jinc > 0 ? j <= jlast : j >= jlast; j += jinc) {
                if (grow <= smlnum) { broke = 1; break; }
                double xj = 1.0 + cnorm[j];
                grow = fmin(grow, xbnd / xj);
                double tjj = fabs(A_(j, j));
                if (xj > tjj) xbnd = xbnd * (tjj / xj);
            }
            if (!broke) grow = fmin(grow, xbnd);
        } else {
            grow = fmin(1.0, 1.0 / fmax(xbnd, smlnum));
            for (j = jfirst; jinc > 0 ? j <= jlast : j >= jlast; j += jinc) {
                if (grow <= smlnum) break;
                double xj = 1.0 + cnorm[j];
                grow = grow / xj;
            }
        }
    }
    if (grow * tscal > smlnum) {
        ora_dtrsv(uplo, trans, diag, n, a, lda, x);                          /* dlatrs.f:562-567 */
    } else {
        double rec, tjjs = 0.0, tjj, xj;
        if (xmax > bignum) { *scale = bignum / xmax; ora_dscal(n, *scale, x, 1); xmax = bignum; }
        if (notran) {
            for (j = jfirst; jinc > 0 ? j <= jlast : j >= jlast; j += jinc) {
                xj = fabs(x[j]);
                int skip = 0;
                if (nounit) tjjs = A_(j, j) * tscal;
                else { tjjs = tscal; if (tscal == 1.0) skip = 1; }
                if (!skip) {
                    tjj = fabs(tjjs);
                    if (tjj > smlnum) {
                        if (tjj < 1.0 && xj > tjj * bignum) { rec = 1.0 / xj; ora_dscal(n, rec, x, 1); *scale *= rec; xmax *= rec; }
                        x[j] = x[j] / tjjs;
                        xj = fabs(x[j]);
                    } else if (tjj > 0.0) {
                        if (xj > tjj * bignum) {
                            rec = (tjj * bignum) / xj;
                            if (cnorm[j] > 1.0) rec = rec / cnorm[j];
                            ora_dscal(n, rec, x, 1); *scale *= rec; xmax *= rec;
                        }
                        x[j] = x[j] / tjjs;
                        xj = fabs(x[j]);
                    } else {
                        for (int i = 0; i < n; ++i) x[i] = 0.0;
                        x[j] = 1.0; xj = 1.0; *scale = 0.0; xmax = 0.0;
                    }
                }
                if (xj > 1.0) {
                    rec = 1.0 / xj;
                    if (cnorm[j] > (bignum - xmax) * rec) { rec *= 0.5; ora_dscal(n, rec, x, 1); *scale *= rec; }
                } else if (xj * cnorm[j] > bignum - xmax) { ora_dscal(n, 0.5, x, 1); *scale *= 0.5; }
                if (upper) {
                    if (j > 0) {
                        ora_daxpy(j, -x[j] * tscal, &A_(0, j), 1, x, 1);
                        int i = ora_idamax(j, x, 1);
                        xmax = fabs(x[i - 1]);
                    }
                } else if (j < n - 1) {
                    ora_daxpy(n - j - 1, -x[j] * tscal, &A_(j + 1, j), 1, x + j + 1, 1);
                    int i = j + ora_idamax(n - j - 1, x + j + 1, 1);
                    xmax = fabs(x[i]);
                }
            }
        } else {
            for (j = jfirst; jinc > 0 ? j <= jlast : j >= jlast; j += jinc) {
                xj = fabs(x[j]);
                double uscal = tscal, sumj;
                rec = 1.0 / fmax(xmax, 1.0);
                if (cnorm[j] > (bignum - xj) * rec) {
                    rec *= 0.5;
                    if (nounit) tjjs = A_(j, j) * tscal; else tjjs = tscal;
                    tjj = fabs(tjjs);
                    if (tjj > 1.0) { rec = fmin(1.0, rec * tjj); uscal = uscal / tjjs; }
                    if (rec < 1.0) { ora_dscal(n, rec, x, 1); *scale *= rec; xmax *= rec; }
                }
                sumj = 0.0;
                if (uscal == 1.0) {
                    if (upper) sumj = ora_ddot(j, &A_(0, j), 1, x, 1);
                    else if (j < n - 1) sumj = ora_ddot(n - j - 1, &A_(j + 1, j), 1, x + j + 1, 1);
                } else {
                    if (upper) for (int i = 0; i < j; ++i) sumj = sumj + (A_(i, j) * uscal) * x[i];
                    else for (int i = j + 1; i < n; ++i) sumj = sumj + (A_(i, j) * uscal) * x[i];
                }
                if (uscal == tscal) {
                    x[j] = x[j] - sumj;
                    xj = fabs(x[j]);
                    int skip = 0;
                    if (nounit) tjjs = A_(j, j) * tscal;
                    else { tjjs = tscal; if (tscal == 1.0) skip = 1; }
                    if (!skip) {
                        tjj = fabs(tjjs);
                        if (tjj > smlnum) {
                            if (tjj < 1.0 && xj > tjj * bignum) { rec = 1.0 / xj; ora_dscal(n, rec, x, 1); *scale *= rec; xmax *= rec; }
                            x[j] = x[j] / tjjs;
                        } else if (tjj > 0.0) {
                            if (xj > tjj * bignum) { rec = (tjj * bignum) / xj; ora_dscal(n, rec, x, 1); *scale *= rec; xmax *= rec; }
                            x[j] = x[j] / tjjs;
                        } else {
                            for (int i = 0; i < n; ++i) x[i] = 0.0;
                            x[j] = 1.0; *scale = 0.0; xmax = 0.0;
                        }
                    }
                } else x[j] = x[j] / tjjs - sumj;
                xmax = fmax(xmax, fabs(x[j]));
            }
        }
        *scale = *scale / tscal;
    }
    if (tscal != 1.0) ora_dscal(n, 1.0 / tscal, cnorm, 1);
}

/* SRC/dgecon.f:128-285.  work: 4n doubles, iwork: n ints. */
void ora_dgecon(char norm, int n, const double *a, int lda, double anorm, double *rcond, double *work, int *iwork, int *info)
{
    const double hugeval = ora_dlamch('O');
    *info = 0;
    const int onenrm = ora_lsame(norm, '1') || ora_lsame(norm, 'O');
    if (!onenrm && !ora_lsame(norm, 'I')) *info = -1;
    else if (n < 0) *info = -2;
    else if (lda < imax(1, n)) *info = -4;
    else if (anorm < 0.0) *info = -5;
    if (*info != 0) return;
    *rcond = 0.0;
    if (n == 0) { *rcond = 1.0; return; }
    else if (anorm == 0.0) return;
    else if (anorm != anorm) { *rcond = anorm; *info = -5; return; }
    else if (anorm > hugeval) { *info = -5; return; }
    const double smlnum = ora_dlamch('S');
    double ainvnm = 0.0, sl, su, scale;
    char normin = 'N';
    const int kase1 = onenrm ? 1 : 2;
    int kase = 0, isave[3] = {0, 0, 0}, iinfo;
    for (;;) {
        ora_dlacn2(n, work + n, work, iwork, &ainvnm, &kase, isave);
        if (kase == 0) break;
        if (kase == kase1) {
            ora_dlatrs('L', 'N', 'U', normin, n, a, lda, work, &sl, work + 2 * n, &iinfo);
            ora_dlatrs('U', 'N', 'N', normin, n, a, lda, work, &su, work + 3 * n, &iinfo);
        } else {
            ora_dlatrs('U', 'T', 'N', normin, n, a, lda, work, &su, work + 3 * n, &iinfo);
            ora_dlatrs('L', 'T', 'U', normin, n, a, lda, work, &sl, work + 2 * n, &iinfo);
        }
        scale = sl * su;
        normin = 'Y';
        if (scale != 1.0) {
            int ix = ora_idamax(n, work, 1);
            if (scale < fabs(work[ix - 1]) * smlnum || scale == 0.0) return;
            ora_drscl(n, scale, work);
        }
    }
    if (ainvnm != 0.0) *rcond = (1.0 / ainvnm) / anorm;
    else { *info = 1; return; }
    if (*rcond != *rcond || *rcond > hugeval) *info = 1;
}

/* SRC/dgeequ.f:160-310 */
void ora_dgeequ(int m, int n, const double *a, int lda, double *r, double *c, double *rowcnd, double *colcnd, double *amax,
                int *info)
{
    *info = 0;
    if (m < 0) *info = -1; else if (n < 0) *info = -2; else if (lda < imax(1, m)) *info = -4;
    if (*info != 0) return;
    if (m == 0 || n == 0) { *rowcnd = 1.0; *colcnd = 1.0; *amax = 0.0; return; }
    const double smlnum = ora_dlamch('S'), bignum = 1.0 / smlnum;
    for (int i = 0; i < m; ++i) r[i] = 0.0;
    for (int j = 0; j < n; ++j) for (int i = 0; i < m; ++i) r[i] = fmax(r[i], fabs(A_(i, j)));
    double rcmin = bignum, rcmax = 0.0;
    for (int i = 0; i < m; ++i) { rcmax = fmax(rcmax, r[i]); rcmin = fmin(rcmin, r[i]); }
    *amax = rcmax;
    if (rcmin == 0.0) { for (int i = 0; i < m; ++i) if (r[i] == 0.0) { *info = i + 1; return; } }
    else {
        for (int i = 0; i < m; ++i) r[i] = 1.0 / fmin(fmax(r[i], smlnum), bignum);
        *rowcnd = fmax(rcmin, smlnum) / fmin(rcmax, bignum);
    }
    for (int j = 0; j < n; ++j) c[j] = 0.0;
    for (int j = 0; j < n; ++j) for (int i = 0; i < m; ++i) c[j] = fmax(c[j], fabs(A_(i, j)) * r[i]);
    rcmin = bignum; rcmax = 0.0;
    for (int j = 0; j < n; ++j) { rcmin = fmin(rcmin, c[j]); rcmax = fmax(rcmax, c[j]); }
    if (rcmin == 0.0) { for (int j = 0; j < n; ++j) if (c[j] == 0.0) { *info = m + j + 1; return; } }
    else {
        for (int j = 0; j < n; ++j) c[j] = 1.0 / fmin(fmax(c[j], smlnum), bignum);
        *colcnd = fmax(rcmin, smlnum) / fmin(rcmax, bignum);
    }
}

/* SRC/dlaqge.f:160-230; returns EQUED */
char ora_dlaqge(int m, int n, double *a, int lda, const double *r, const double *c, double rowcnd, double colcnd, double amax)
{
    const double thresh = 0.1;
    if (m <= 0 || n <= 0) return 'N';
    const double small = ora_dlamch('S') / ora_dlamch('P'), large = 1.0 / small;
    if (rowcnd >= thresh && amax >= small && amax <= large) {
        if (colcnd >= thresh) return 'N';
        for (int j = 0; j < n; ++j) { double cj = c[j]; for (int i = 0; i < m; ++i) A_(i, j) = cj * A_(i, j); }
        return 'C';
    } else if (colcnd >= thresh) {
        for (int j = 0; j < n; ++j) for (int i = 0; i < m; ++i) A_(i, j) = r[i] * A_(i, j);
        return 'R';
    }
    for (int j = 0; j < n; ++j) { double cj = c[j]; for (int i = 0; i < m; ++i) A_(i, j) = cj * r[i] * A_(i, j); }
    return 'B';
}

static double lantr_max_upper_(int m, int n, const double *a, int lda)        /* DLANTR('M','U','N', m, n) (dlantr.f:190-200) */
{
    double v = 0.0;
    for (int j = 0; j < n; ++j)
        for (int i = 0; i < imin(m, j + 1); ++i) { double t = fabs(A_(i, j)); if (v < t || t != t) v = t; }
    return v;
}

/* SRC/dgesvx.f:344-600.  work: 4n doubles (work[0] returns RPVGRW), iwork: n ints.  equed is in/out. */
void ora_dgesvx(char fact, char trans, int n, int nrhs, double *a, int lda, double *af, int ldaf, int *ipiv, char *equed,
                double *r, double *c, double *b, int ldb, double *x, int ldx, double *rcond, double *ferr, double *berr,
                double *work, int *iwork, int *info)
{
    *info = 0;
    const int nofact = ora_lsame(fact, 'N'), equil = ora_lsame(fact, 'E'), notran = ora_lsame(trans, 'N');
    int rowequ, colequ;
    double smlnum = ora_dlamch('S'), bignum = 1.0 / smlnum, rowcnd = 1.0, colcnd = 1.0, amax, rcmin, rcmax, rpvgrw;
    if (nofact || equil) { *equed = 'N'; rowequ = 0; colequ = 0; }
    else { rowequ = ora_lsame(*equed, 'R') || ora_lsame(*equed, 'B'); colequ = ora_lsame(*equed, 'C') || ora_lsame(*equed, 'B'); }
    if (!nofact && !equil && !ora_lsame(fact, 'F')) *info = -1;
    else if (!notran && !ora_lsame(trans, 'T') && !ora_lsame(trans, 'C')) *info = -2;
    else if (n < 0) *info = -3;
    else if (nrhs < 0) *info = -4;
    else if (lda < imax(1, n)) *info = -6;
    else if (ldaf < imax(1, n)) *info = -8;
    else if (ora_lsame(fact, 'F') && !(rowequ || colequ || ora_lsame(*equed, 'N'))) *info = -10;
    else {
        if (rowequ) {
            rcmin = bignum; rcmax = 0.0;
            for (int j = 0; j < n; ++j) { rcmin = fmin(rcmin, r[j]); rcmax = fmax(rcmax, r[j]); }
            if (rcmin <= 0.0) *info = -11; else if (n > 0) rowcnd = fmax(rcmin, smlnum) / fmin(rcmax, bignum); else rowcnd = 1.0;
        }
        if (colequ && *info == 0) {
            rcmin = bignum; rcmax = 0.0;
            for (int j = 0; j < n; ++j) { rcmin = fmin(rcmin, c[j]); rcmax = fmax(rcmax, c[j]); }
            if (rcmin <= 0.0) *info = -12; else if (n > 0) colcnd = fmax(rcmin, smlnum) / fmin(rcmax, bignum); else colcnd = 1.0;
        }
        if (*info == 0) { if (ldb < imax(1, n)) *info = -14; else if (ldx < imax(1, n)) *info = -16; }
    }
    if (*info != 0) return;
    if (equil) {
        int infequ;
        ora_dgeequ(n, n, a, lda, r, c, &rowcnd, &colcnd, &amax, &infequ);
        if (infequ == 0) {
            *equed = ora_dlaqge(n, n, a, lda, r, c, rowcnd, colcnd, amax);
            rowequ = ora_lsame(*equed, 'R') || ora_lsame(*equed, 'B');
            colequ = ora_lsame(*equed, 'C') || ora_lsame(*equed, 'B');
        }
    }
    if (notran) { if (rowequ) for (int j = 0; j < nrhs; ++j) for (int i = 0; i < n; ++i) b[i + (size_t)j * ldb] *= r[i]; }
    else if (colequ) for (int j = 0; j < nrhs; ++j) for (int i = 0; i < n; ++i) b[i + (size_t)j * ldb] *= c[i];
    if (nofact || equil) {
        ora_dlacpy('F', n, n, a, lda, af, ldaf);
        ora_dgetrf(n, n, af, ldaf, ipiv, info);
        if (*info > 0) {
            rpvgrw = lantr_max_upper_(*info, *info, af, ldaf);
            if (rpvgrw == 0.0) rpvgrw = 1.0; else rpvgrw = ora_dlange('M', n, *info, a, lda) / rpvgrw;
            work[0] = rpvgrw;
            *rcond = 0.0;
            return;
        }
    }
    char norm = notran ? '1' : 'I';
    double anorm = ora_dlange(norm, n, n, a, lda);
    rpvgrw = lantr_max_upper_(n, n, af, ldaf);
    if (rpvgrw == 0.0) rpvgrw = 1.0; else rpvgrw = ora_dlange('M', n, n, a, lda) / rpvgrw;
    ora_dgecon(norm, n, af, ldaf, anorm, rcond, work, iwork, info);
    ora_dlacpy('F', n, nrhs, b, ldb, x, ldx);
    ora_dgetrs(trans, n, nrhs, af, ldaf, ipiv, x, ldx, info);
    ora_dgerfs(trans, n, nrhs, a, lda, af, ldaf, ipiv, b, ldb, x, ldx, ferr, berr, work, iwork, info);
    if (notran) {
        if (colequ) {
            for (int j = 0; j < nrhs; ++j) for (int i = 0; i < n; ++i) x[i + (size_t)j * ldx] *= c[i];
            for (int j = 0; j < nrhs; ++j) ferr[j] /= colcnd;
        }
    } else if (rowequ) {
        for (int j = 0; j < nrhs; ++j) for (int i = 0; i < n; ++i) x[i + (size_t)j * ldx] *= r[i];
        for (int j = 0; j < nrhs; ++j) ferr[j] /= rowcnd;
    }
    work[0] = rpvgrw;
    if (*rcond < ora_dlamch('E')) *info = n + 1;
}

/* ======================================================================================================================
 * Tall-skinny QR: DLATSQR and the triangular-pentagonal kernels it calls, restated for L = 0 -- the only
 * value DLATSQR passes (dlatsqr.f:265-277): the pentagonal block B is then a plain M x N rectangle under the N x N triangle A.
 * ====================================================================================================================== */
#define B_(i, j) b[(size_t)(i) + (size_t)(j) * ldb]
#define T_(i, j) t[(size_t)(i) + (size_t)(j) * ldt]

/* SRC/dtpqrt2.f:214-300 with L = 0 */
void ora_dtpqrt2_l0(int m, int n, double *a, int lda, double *b, int ldb, double *t, int ldt)
{
    if (n == 0 || m == 0) return;
    for (int i = 0; i < n; ++i) {
        /* generate H(i) to annihilate B(:,i) (dtpqrt2.f:221) */
        ora_dlarfg(m + 1, &A_(i, i), &B_(0, i), 1, &T_(i, 0));
        if (i < n - 1) {
            for (int j = 0; j < n - i - 1; ++j) T_(j, n - 1) = A_(i, i + 1 + j);
            ora_dgemv('T', m, n - i - 1, 1.0, &B_(0, i + 1), ldb, &B_(0, i), 1, 1.0, &T_(0, n - 1), 1);
            double alpha = -T_(i, 0);
            for (int j = 0; j < n - i - 1; ++j) A_(i, i + 1 + j) = A_(i, i + 1 + j) + alpha * T_(j, n - 1);
            ora_dger(m, n - i - 1, alpha, &B_(0, i), 1, &T_(0, n - 1), 1, &B_(0, i + 1), ldb);
        }
    }
    for (int i = 1; i < n; ++i) {
        double alpha = -T_(i, 0);
        for (int j = 0; j < i; ++j) T_(j, i) = 0.0;
        /* P = 0: no triangular part of B2; the rectangular part of B2 has L = 0 rows: y := 0 (dtpqrt2.f:271-285) */
        ora_dgemv('T', m, i, alpha, b, ldb, &B_(0, i), 1, 1.0, &T_(0, i), 1);                  /* B1 (dtpqrt2.f:289) */
        ora_dtrmv('U', 'N', 'N', i, t, ldt, &T_(0, i), 1);                                     /* dtpqrt2.f:294 */
        T_(i, i) = T_(i, 0);
        T_(i, 0) = 0.0;
    }
}

/* SRC/dtprfb.f:322-383 ('L','T','F','C') with L = 0: [A; B] := H^T [A; B], H = I - W T W^T, W = [I; V] */
void ora_dtprfb_ltfc_l0(int m, int n, int k, const double *v, int ldv, const double *t, int ldt, double *a, int lda, double *b,
                        int ldb, double *work, int ldwork)
{
    if (m <= 0 || n <= 0 || k <= 0) return;
    ora_dgemm('T', 'N', k, n, m, 1.0, v, ldv, b, ldb, 0.0, work, ldwork);
    for (int j = 0; j < n; ++j) for (int i = 0; i < k; ++i) work[i + (size_t)j * ldwork] += A_(i, j);
    ora_dtrmm('L', 'U', 'T', 'N', k, n, 1.0, t, ldt, work, ldwork);
    for (int j = 0; j < n; ++j) for (int i = 0; i < k; ++i) A_(i, j) -= work[i + (size_t)j * ldwork];
    ora_dgemm('N', 'N', m, n, k, -1.0, v, ldv, work, ldwork, 1.0, b, ldb);
}

/* SRC/dtpqrt.f:200-270 with L = 0.  work: nb*n doubles */
void ora_dtpqrt_l0(int m, int n, int nb, double *a, int lda, double *b, int ldb, double *t, int ldt, double *work, int *info)
{
    *info = 0;
    if (m < 0) *info = -1; else if (n < 0) *info = -2; else if (nb < 1 || (nb > n && n > 0)) *info = -4;
    else if (lda < imax(1, n)) *info = -6; else if (ldb < imax(1, m)) *info = -8; else if (ldt < nb) *info = -10;
    if (*info != 0 || m == 0 || n == 0) return;
    for (int i = 0; i < n; i += nb) {
        const int ib = imin(n - i, nb);
        ora_dtpqrt2_l0(m, ib, &A_(i, i), lda, &B_(0, i), ldb, &T_(0, i), ldt);
        if (i + ib < n)
            ora_dtprfb_ltfc_l0(m, n - i - ib, ib, &B_(0, i), ldb, &T_(0, i), ldt, &A_(i, i + ib), lda, &B_(0, i + ib), ldb, work, ib);
    }
}

/* SRC/dlatsqr.f:185-290.  t: nb x n*ceil((m-n)/(mb-n)) ; work: nb*n doubles */
void ora_dlatsqr(int m, int n, int mb, int nb, double *a, int lda, double *t, int ldt, double *work, int *info)
{
    *info = 0;
    if (m < 0) *info = -1; else if (n < 0 || m < n) *info = -2; else if (mb < 1) *info = -3;
    else if (nb < 1 || (nb > n && n > 0)) *info = -4; else if (lda < imax(1, m)) *info = -6; else if (ldt < nb) *info = -8;
    if (*info != 0) return;
    if (imin(m, n) == 0) return;
    if (mb <= n || mb >= m) { ora_dgeqrt(m, n, nb, a, lda, t, ldt, work, info); return; }
    const int kk = (m - n) % (mb - n), ii = m - kk;                       /* ii: 0-based first row of the last block */
    ora_dgeqrt(mb, n, nb, a, lda, t, ldt, work, info);
    int ctr = 1;
    for (int i = mb; i <= ii - mb + n; i += mb - n) {                    /* DO I = MB+1, II-MB+N, MB-N (1-based; II = ii+1) */
        ora_dtpqrt_l0(mb - n, n, nb, a, lda, &A_(i, 0), lda, &T_(0, ctr * n), ldt, work, info);
        ++ctr;
    }
    if (ii < m) ora_dtpqrt_l0(kk, n, nb, a, lda, &A_(ii, 0), lda, &T_(0, ctr * n), ldt, work, info);
}
