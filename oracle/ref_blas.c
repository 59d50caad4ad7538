/*
 * ref_blas.c -- CPU restatement of the reference BLAS routines on the one-sided factorization
 * hot path.  TEST INFRASTRUCTURE ONLY (see oracle.h).  Same loop nests and the same order of
 * floating-point operations as /root/reference/BLAS/SRC (cited per function); 0-based indices.
 */
#include "oracle.h"
#include <math.h>
#include <float.h>

#define A_(i, j) a[(size_t)(i) + (size_t)(j) * lda]
#define B_(i, j) b[(size_t)(i) + (size_t)(j) * ldb]
#define C_(i, j) c[(size_t)(i) + (size_t)(j) * ldc]

/* BLAS/SRC/lsame.f:  case-insensitive single character compare (ASCII branch). */
int ora_lsame(char a, char b)
{
    if (a >= 'a' && a <= 'z') a = (char)(a - 32);
    if (b >= 'a' && b <= 'z') b = (char)(b - 32);
    return a == b;
}

/* BLAS/SRC/idamax.f:93-123.  1-based result; first index of max |x| (strict >), 0 if n<1 or incx<=0. */
int ora_idamax(int n, const double *x, int incx)
{
    if (n < 1 || incx <= 0) return 0;
    int best = 1;
    if (n == 1) return best;
    double dmax = fabs(x[0]);
    size_t ix = (size_t)incx;
    for (int i = 2; i <= n; ++i, ix += (size_t)incx) {
        if (fabs(x[ix]) > dmax) { best = i; dmax = fabs(x[ix]); }
    }
    return best;
}

/* BLAS/SRC/dscal.f:104-135.  Note the early return when da == 1. */
void ora_dscal(int n, double da, double *x, int incx)
{
    if (n <= 0 || incx <= 0 || da == 1.0) return;
    for (int i = 0; i < n; ++i) x[(size_t)i * incx] = da * x[(size_t)i * incx];
}

/* BLAS/SRC/dswap.f */
void ora_dswap(int n, double *x, int incx, double *y, int incy)
{
    if (n <= 0) return;
    ptrdiff_t ix = incx < 0 ? (ptrdiff_t)(1 - n) * incx : 0;
    ptrdiff_t iy = incy < 0 ? (ptrdiff_t)(1 - n) * incy : 0;
    for (int i = 0; i < n; ++i, ix += incx, iy += incy) { double t = x[ix]; x[ix] = y[iy]; y[iy] = t; }
}

/* BLAS/SRC/daxpy.f:  y += a*x, returns early for a == 0. */
void ora_daxpy(int n, double da, const double *x, int incx, double *y, int incy)
{
    if (n <= 0 || da == 0.0) return;
    ptrdiff_t ix = incx < 0 ? (ptrdiff_t)(1 - n) * incx : 0;
    ptrdiff_t iy = incy < 0 ? (ptrdiff_t)(1 - n) * incy : 0;
    for (int i = 0; i < n; ++i, ix += incx, iy += incy) y[iy] = y[iy] + da * x[ix];
}

/* BLAS/SRC/dcopy.f */
void ora_dcopy(int n, const double *x, int incx, double *y, int incy)
{
    if (n <= 0) return;
    ptrdiff_t ix = incx < 0 ? (ptrdiff_t)(1 - n) * incx : 0;
    ptrdiff_t iy = incy < 0 ? (ptrdiff_t)(1 - n) * incy : 0;
    for (int i = 0; i < n; ++i, ix += incx, iy += incy) y[iy] = x[ix];
}

/* BLAS/SRC/ddot.f (the 5-way unrolled sum is a single left-to-right accumulation chain). */
double ora_ddot(int n, const double *x, int incx, const double *y, int incy)
{
    double t = 0.0;
    if (n <= 0) return t;
    if (incx == 1 && incy == 1) {
        int m = n % 5;
        for (int i = 0; i < m; ++i) t = t + x[i] * y[i];
        if (n < 5) return t;
        for (int i = m; i < n; i += 5)
            t = t + x[i] * y[i] + x[i + 1] * y[i + 1] + x[i + 2] * y[i + 2] + x[i + 3] * y[i + 3] +
                x[i + 4] * y[i + 4];
        return t;
    }
    ptrdiff_t ix = incx < 0 ? (ptrdiff_t)(1 - n) * incx : 0;
    ptrdiff_t iy = incy < 0 ? (ptrdiff_t)(1 - n) * incy : 0;
    for (int i = 0; i < n; ++i, ix += incx, iy += incy) t = t + x[ix] * y[iy];
    return t;
}

/* BLAS/SRC/dnrm2.f90:126-198.  Blue's three-accumulator sum of squares. */
double ora_dnrm2(int n, const double *x, int incx)
{
    if (n <= 0) return 0.0;
    /* radix 2, minexponent -1021, maxexponent 1024, digits 53 (dnrm2.f90:103-110) */
    const double tsml = ldexp(1.0, -511);   /* 2**ceiling((-1021-1)*0.5)       */
    const double tbig = ldexp(1.0, 486);    /* 2**floor((1024-53+1)*0.5)       */
    const double ssml = ldexp(1.0, 537);    /* 2**(-floor((-1021-53)*0.5))     */
    const double sbig = ldexp(1.0, -538);   /* 2**(-ceiling((1024+53-1)*0.5))  */
    const double maxN = DBL_MAX;
    double scl = 1.0, sumsq = 0.0, asml = 0.0, amed = 0.0, abig = 0.0;
    int notbig = 1;
    ptrdiff_t ix = incx < 0 ? (ptrdiff_t)(1 - n) * incx : 0;
    for (int i = 0; i < n; ++i, ix += incx) {
        double ax = fabs(x[ix]);
        if (ax > tbig) { double s = ax * sbig; abig = abig + s * s; notbig = 0; }
        else if (ax < tsml) { if (notbig) { double s = ax * ssml; asml = asml + s * s; } }
        else amed = amed + ax * ax;
    }
    if (abig > 0.0) {
        if (amed > 0.0 || amed > maxN || amed != amed) abig = abig + (amed * sbig) * sbig;
        scl = 1.0 / sbig;
        sumsq = abig;
    } else if (asml > 0.0) {
        if (amed > 0.0 || amed > maxN || amed != amed) {
            double ymin, ymax;
            amed = sqrt(amed);
            asml = sqrt(asml) / ssml;
            if (asml > amed) { ymin = amed; ymax = asml; } else { ymin = asml; ymax = amed; }
            scl = 1.0;
            double q = ymin / ymax;
            sumsq = ymax * ymax * (1.0 + q * q);
        } else {
            scl = 1.0 / ssml;
            sumsq = asml;
        }
    } else {
        scl = 1.0;
        sumsq = amed;
    }
    return scl * sqrt(sumsq);
}

/* BLAS/SRC/dgemv.f:247-323 */
void ora_dgemv(char trans, int m, int n, double alpha, const double *a, int lda, const double *x, int incx,
               double beta, double *y, int incy)
{
    if (m == 0 || n == 0 || (alpha == 0.0 && beta == 1.0)) return;
    int notrans = ora_lsame(trans, 'N');
    int lenx = notrans ? n : m, leny = notrans ? m : n;
    ptrdiff_t kx = incx > 0 ? 0 : -(ptrdiff_t)(lenx - 1) * incx;
    ptrdiff_t ky = incy > 0 ? 0 : -(ptrdiff_t)(leny - 1) * incy;
    if (beta != 1.0) {
        ptrdiff_t iy = ky;
        for (int i = 0; i < leny; ++i, iy += incy) y[iy] = (beta == 0.0) ? 0.0 : beta * y[iy];
    }
    if (alpha == 0.0) return;
    if (notrans) {
        ptrdiff_t jx = kx;
        for (int j = 0; j < n; ++j, jx += incx) {
            double temp = alpha * x[jx];
            ptrdiff_t iy = ky;
            for (int i = 0; i < m; ++i, iy += incy) y[iy] = y[iy] + temp * A_(i, j);
        }
    } else {
        ptrdiff_t jy = ky;
        for (int j = 0; j < n; ++j, jy += incy) {
            double temp = 0.0;
            ptrdiff_t ix = kx;
            for (int i = 0; i < m; ++i, ix += incx) temp = temp + A_(i, j) * x[ix];
            y[jy] = y[jy] + alpha * temp;
        }
    }
}

/* BLAS/SRC/dger.f:182-219  (skips columns with y(j) == 0) */
void ora_dger(int m, int n, double alpha, const double *x, int incx, const double *y, int incy, double *a,
              int lda)
{
    if (m == 0 || n == 0 || alpha == 0.0) return;
    ptrdiff_t jy = incy > 0 ? 0 : -(ptrdiff_t)(n - 1) * incy;
    ptrdiff_t kx = incx > 0 ? 0 : -(ptrdiff_t)(m - 1) * incx;
    for (int j = 0; j < n; ++j, jy += incy) {
        if (y[jy] != 0.0) {
            double temp = alpha * y[jy];
            ptrdiff_t ix = kx;
            for (int i = 0; i < m; ++i, ix += incx) A_(i, j) = A_(i, j) + x[ix] * temp;
        }
    }
}

/* BLAS/SRC/dtrmv.f:  x := op(A) x, A triangular (this tree has no zero-skip tests). */
void ora_dtrmv(char uplo, char trans, char diag, int n, const double *a, int lda, double *x, int incx)
{
    if (n == 0) return;
    int nounit = ora_lsame(diag, 'N');
    int upper = ora_lsame(uplo, 'U');
    ptrdiff_t kx = incx > 0 ? 0 : -(ptrdiff_t)(n - 1) * incx;
#define X_(i) x[kx + (ptrdiff_t)(i) * incx]
    if (ora_lsame(trans, 'N')) {
        if (upper) {
            for (int j = 0; j < n; ++j) {
                double temp = X_(j);
                for (int i = 0; i < j; ++i) X_(i) = X_(i) + temp * A_(i, j);
                if (nounit) X_(j) = X_(j) * A_(j, j);
            }
        } else {
            for (int j = n - 1; j >= 0; --j) {
                double temp = X_(j);
                for (int i = n - 1; i > j; --i) X_(i) = X_(i) + temp * A_(i, j);
                if (nounit) X_(j) = X_(j) * A_(j, j);
            }
        }
    } else {
        if (upper) {
            for (int j = n - 1; j >= 0; --j) {
                double temp = X_(j);
                if (nounit) temp = temp * A_(j, j);
                for (int i = j - 1; i >= 0; --i) temp = temp + A_(i, j) * X_(i);
                X_(j) = temp;
            }
        } else {
            for (int j = 0; j < n; ++j) {
                double temp = X_(j);
                if (nounit) temp = temp * A_(j, j);
                for (int i = j + 1; i < n; ++i) temp = temp + A_(i, j) * X_(i);
                X_(j) = temp;
            }
        }
    }
#undef X_
}

/* BLAS/SRC/dgemm.f:298-383.  (j,l,i) axpy order for op(A)=A; dot-product form for op(A)=A**T. */
void ora_dgemm(char transa, char transb, int m, int n, int k, double alpha, const double *a, int lda,
               const double *b, int ldb, double beta, double *c, int ldc)
{
    int nota = ora_lsame(transa, 'N'), notb = ora_lsame(transb, 'N');
    if (m == 0 || n == 0 || ((alpha == 0.0 || k == 0) && beta == 1.0)) return;
    if (alpha == 0.0) {
        for (int j = 0; j < n; ++j)
            for (int i = 0; i < m; ++i) C_(i, j) = (beta == 0.0) ? 0.0 : beta * C_(i, j);
        return;
    }
    if (notb) {
        if (nota) {
            for (int j = 0; j < n; ++j) {
                if (beta == 0.0) { for (int i = 0; i < m; ++i) C_(i, j) = 0.0; }
                else if (beta != 1.0) { for (int i = 0; i < m; ++i) C_(i, j) = beta * C_(i, j); }
                for (int l = 0; l < k; ++l) {
                    double temp = alpha * B_(l, j);
                    for (int i = 0; i < m; ++i) C_(i, j) = C_(i, j) + temp * A_(i, l);
                }
            }
        } else {
            for (int j = 0; j < n; ++j)
                for (int i = 0; i < m; ++i) {
                    double temp = 0.0;
                    for (int l = 0; l < k; ++l) temp = temp + A_(l, i) * B_(l, j);
                    C_(i, j) = (beta == 0.0) ? alpha * temp : alpha * temp + beta * C_(i, j);
                }
        }
    } else {
        if (nota) {
            for (int j = 0; j < n; ++j) {
                if (beta == 0.0) { for (int i = 0; i < m; ++i) C_(i, j) = 0.0; }
                else if (beta != 1.0) { for (int i = 0; i < m; ++i) C_(i, j) = beta * C_(i, j); }
                for (int l = 0; l < k; ++l) {
                    double temp = alpha * B_(j, l);
                    for (int i = 0; i < m; ++i) C_(i, j) = C_(i, j) + temp * A_(i, l);
                }
            }
        } else {
            for (int j = 0; j < n; ++j)
                for (int i = 0; i < m; ++i) {
                    double temp = 0.0;
                    for (int l = 0; l < k; ++l) temp = temp + A_(l, i) * B_(j, l);
                    C_(i, j) = (beta == 0.0) ? alpha * temp : alpha * temp + beta * C_(i, j);
                }
        }
    }
}

/* BLAS/SRC/dtrsm.f:257-405 -- all 16 option combinations, loop nests as in this tree's dtrsm.f
 * (unconditional alpha scaling, no zero-skip tests, divisions by the diagonal). */
void ora_dtrsm(char side, char uplo, char transa, char diag, int m, int n, double alpha, const double *a,
               int lda, double *b, int ldb)
{
    int lside = ora_lsame(side, 'L'), nounit = ora_lsame(diag, 'N'), upper = ora_lsame(uplo, 'U');
    int notr = ora_lsame(transa, 'N');
    if (m == 0 || n == 0) return;
    if (alpha == 0.0) {                                    /* dtrsm.f:261-268 */
        for (int j = 0; j < n; ++j)
            for (int i = 0; i < m; ++i) B_(i, j) = 0.0;
        return;
    }
    if (lside) {
        if (notr) {
            if (upper) {                                   /* L,U,N */
                for (int j = 0; j < n; ++j) {
                    for (int i = 0; i < m; ++i) B_(i, j) = alpha * B_(i, j);
                    for (int k = m - 1; k >= 0; --k) {
                        if (nounit) B_(k, j) = B_(k, j) / A_(k, k);
                        for (int i = 0; i < k; ++i) B_(i, j) = B_(i, j) - B_(k, j) * A_(i, k);
                    }
                }
            } else {                                       /* L,L,N */
                for (int j = 0; j < n; ++j) {
                    for (int i = 0; i < m; ++i) B_(i, j) = alpha * B_(i, j);
                    for (int k = 0; k < m; ++k) {
                        if (nounit) B_(k, j) = B_(k, j) / A_(k, k);
                        for (int i = k + 1; i < m; ++i) B_(i, j) = B_(i, j) - B_(k, j) * A_(i, k);
                    }
                }
            }
        } else {
            if (upper) {                                   /* L,U,T */
                for (int j = 0; j < n; ++j)
                    for (int i = 0; i < m; ++i) {
                        double temp = alpha * B_(i, j);
                        for (int k = 0; k < i; ++k) temp = temp - A_(k, i) * B_(k, j);
                        if (nounit) temp = temp / A_(i, i);
                        B_(i, j) = temp;
                    }
            } else {                                       /* L,L,T */
                for (int j = 0; j < n; ++j)
                    for (int i = m - 1; i >= 0; --i) {
                        double temp = alpha * B_(i, j);
                        for (int k = i + 1; k < m; ++k) temp = temp - A_(k, i) * B_(k, j);
                        if (nounit) temp = temp / A_(i, i);
                        B_(i, j) = temp;
                    }
            }
        }
    } else {
        if (notr) {
            if (upper) {                                   /* R,U,N */
                for (int j = 0; j < n; ++j) {
                    for (int i = 0; i < m; ++i) B_(i, j) = alpha * B_(i, j);
                    for (int k = 0; k < j; ++k)
                        for (int i = 0; i < m; ++i) B_(i, j) = B_(i, j) - A_(k, j) * B_(i, k);
                    if (nounit) for (int i = 0; i < m; ++i) B_(i, j) = B_(i, j) / A_(j, j);
                }
            } else {                                       /* R,L,N */
                for (int j = n - 1; j >= 0; --j) {
                    for (int i = 0; i < m; ++i) B_(i, j) = alpha * B_(i, j);
                    for (int k = j + 1; k < n; ++k)
                        for (int i = 0; i < m; ++i) B_(i, j) = B_(i, j) - A_(k, j) * B_(i, k);
                    if (nounit) for (int i = 0; i < m; ++i) B_(i, j) = B_(i, j) / A_(j, j);
                }
            }
        } else {
            if (upper) {                                   /* R,U,T */
                for (int k = n - 1; k >= 0; --k) {
                    if (nounit) for (int i = 0; i < m; ++i) B_(i, k) = B_(i, k) / A_(k, k);
                    for (int j = 0; j < k; ++j)
                        for (int i = 0; i < m; ++i) B_(i, j) = B_(i, j) - A_(j, k) * B_(i, k);
                    for (int i = 0; i < m; ++i) B_(i, k) = alpha * B_(i, k);
                }
            } else {                                       /* R,L,T */
                for (int k = 0; k < n; ++k) {
                    if (nounit) for (int i = 0; i < m; ++i) B_(i, k) = B_(i, k) / A_(k, k);
                    for (int j = k + 1; j < n; ++j)
                        for (int i = 0; i < m; ++i) B_(i, j) = B_(i, j) - A_(j, k) * B_(i, k);
                    for (int i = 0; i < m; ++i) B_(i, k) = alpha * B_(i, k);
                }
            }
        }
    }
}

/* BLAS/SRC/dtrmm.f -- all option combinations (this tree: no zero-skip tests). */
void ora_dtrmm(char side, char uplo, char transa, char diag, int m, int n, double alpha, const double *a,
               int lda, double *b, int ldb)
{
    int lside = ora_lsame(side, 'L'), nounit = ora_lsame(diag, 'N'), upper = ora_lsame(uplo, 'U');
    int notr = ora_lsame(transa, 'N');
    if (m == 0 || n == 0) return;
    if (alpha == 0.0) {
        for (int j = 0; j < n; ++j)
            for (int i = 0; i < m; ++i) B_(i, j) = 0.0;
        return;
    }
    if (lside) {
        if (notr) {
            if (upper) {
                for (int j = 0; j < n; ++j)
                    for (int k = 0; k < m; ++k) {
                        double temp = alpha * B_(k, j);
                        for (int i = 0; i < k; ++i) B_(i, j) = B_(i, j) + temp * A_(i, k);
                        if (nounit) temp = temp * A_(k, k);
                        B_(k, j) = temp;
                    }
            } else {
                for (int j = 0; j < n; ++j)
                    for (int k = m - 1; k >= 0; --k) {
                        double temp = alpha * B_(k, j);
                        B_(k, j) = temp;
                        if (nounit) B_(k, j) = B_(k, j) * A_(k, k);
                        for (int i = k + 1; i < m; ++i) B_(i, j) = B_(i, j) + temp * A_(i, k);
                    }
            }
        } else {
            if (upper) {
                for (int j = 0; j < n; ++j)
                    for (int i = m - 1; i >= 0; --i) {
                        double temp = B_(i, j);
                        if (nounit) temp = temp * A_(i, i);
                        for (int k = 0; k < i; ++k) temp = temp + A_(k, i) * B_(k, j);
                        B_(i, j) = alpha * temp;
                    }
            } else {
                for (int j = 0; j < n; ++j)
                    for (int i = 0; i < m; ++i) {
                        double temp = B_(i, j);
                        if (nounit) temp = temp * A_(i, i);
                        for (int k = i + 1; k < m; ++k) temp = temp + A_(k, i) * B_(k, j);
                        B_(i, j) = alpha * temp;
                    }
            }
        }
    } else {
        if (notr) {
            if (upper) {
                for (int j = n - 1; j >= 0; --j) {
                    double temp = alpha;
                    if (nounit) temp = temp * A_(j, j);
                    for (int i = 0; i < m; ++i) B_(i, j) = temp * B_(i, j);
                    for (int k = 0; k < j; ++k) {
                        temp = alpha * A_(k, j);
                        for (int i = 0; i < m; ++i) B_(i, j) = B_(i, j) + temp * B_(i, k);
                    }
                }
            } else {
                for (int j = 0; j < n; ++j) {
                    double temp = alpha;
                    if (nounit) temp = temp * A_(j, j);
                    for (int i = 0; i < m; ++i) B_(i, j) = temp * B_(i, j);
                    for (int k = j + 1; k < n; ++k) {
                        temp = alpha * A_(k, j);
                        for (int i = 0; i < m; ++i) B_(i, j) = B_(i, j) + temp * B_(i, k);
                    }
                }
            }
        } else {
            if (upper) {
                for (int k = 0; k < n; ++k) {
                    for (int j = 0; j < k; ++j) {
                        double temp = alpha * A_(j, k);
                        for (int i = 0; i < m; ++i) B_(i, j) = B_(i, j) + temp * B_(i, k);
                    }
                    double temp = alpha;
                    if (nounit) temp = temp * A_(k, k);
                    if (temp != 1.0) for (int i = 0; i < m; ++i) B_(i, k) = temp * B_(i, k);
                }
            } else {
                for (int k = n - 1; k >= 0; --k) {
                    for (int j = k + 1; j < n; ++j) {
                        double temp = alpha * A_(j, k);
                        for (int i = 0; i < m; ++i) B_(i, j) = B_(i, j) + temp * B_(i, k);
                    }
                    double temp = alpha;
                    if (nounit) temp = temp * A_(k, k);
                    if (temp != 1.0) for (int i = 0; i < m; ++i) B_(i, k) = temp * B_(i, k);
                }
            }
        }
    }
}

/* BLAS/SRC/dsyrk.f:238-355 */
void ora_dsyrk(char uplo, char trans, int n, int k, double alpha, const double *a, int lda, double beta,
               double *c, int ldc)
{
    int upper = ora_lsame(uplo, 'U');
    if (n == 0 || ((alpha == 0.0 || k == 0) && beta == 1.0)) return;
    if (alpha == 0.0) {
        for (int j = 0; j < n; ++j) {
            int i0 = upper ? 0 : j, i1 = upper ? j + 1 : n;
            for (int i = i0; i < i1; ++i) C_(i, j) = (beta == 0.0) ? 0.0 : beta * C_(i, j);
        }
        return;
    }
    if (ora_lsame(trans, 'N')) {
        for (int j = 0; j < n; ++j) {
            int i0 = upper ? 0 : j, i1 = upper ? j + 1 : n;
            if (beta == 0.0) { for (int i = i0; i < i1; ++i) C_(i, j) = 0.0; }
            else if (beta != 1.0) { for (int i = i0; i < i1; ++i) C_(i, j) = beta * C_(i, j); }
            for (int l = 0; l < k; ++l) {
                if (A_(j, l) != 0.0) {
                    double temp = alpha * A_(j, l);
                    for (int i = i0; i < i1; ++i) C_(i, j) = C_(i, j) + temp * A_(i, l);
                }
            }
        }
    } else {
        for (int j = 0; j < n; ++j) {
            int i0 = upper ? 0 : j, i1 = upper ? j + 1 : n;
            for (int i = i0; i < i1; ++i) {
                double temp = 0.0;
                for (int l = 0; l < k; ++l) temp = temp + A_(l, i) * A_(l, j);
                C_(i, j) = (beta == 0.0) ? alpha * temp : alpha * temp + beta * C_(i, j);
            }
        }
    }
}
