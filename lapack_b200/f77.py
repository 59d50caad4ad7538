"""ctypes bindings of the Fortran-77 ABI symbols (include/lapack_b200_f77.h) for numpy host arrays.

Argument meaning, in-place behaviour, IPIV/INFO conventions and error handling are the reference's
(SRC/dgetrf.f, dgetrs.f, dgesv.f, dpotrf.f, dpotrs.f, dposv.f, dgeqrf.f, dlarft.f, dlarfb.f, dlaswp.f,
BLAS/SRC/dgemm.f, dtrsm.f, dsyrk.f).  Arrays must be float64 / int32, column-major (order='F'); they are
modified in place.  Every function returns INFO where the reference has one.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import lib

_i = C.c_int
_d = C.c_double


def _p(a):
    return C.c_void_p(a.ctypes.data) if isinstance(a, np.ndarray) else C.c_void_p(int(a))


def _ld(a):
    if a.ndim == 1:
        return max(1, a.shape[0])
    if a.size == 0:
        return max(1, a.shape[0])
    assert a.shape[0] <= 1 or a.strides[0] == a.itemsize, "column-major (order='F') array expected"
    return max(1, a.strides[1] // a.itemsize, a.shape[0]) if a.shape[1] > 1 else max(1, a.shape[0])


def _r(v):
    return C.byref(_i(v))


def _c(ch):
    return C.c_char_p(ch.encode())


def dgemm(transa, transb, m, n, k, alpha, a, lda, b, ldb, beta, c, ldc):
    lib().dgemm_(_c(transa), _c(transb), _r(m), _r(n), _r(k), C.byref(_d(alpha)), _p(a), _r(lda), _p(b), _r(ldb),
                 C.byref(_d(beta)), _p(c), _r(ldc), C.c_size_t(1), C.c_size_t(1))


def dsyrk(uplo, trans, n, k, alpha, a, lda, beta, c, ldc):
    lib().dsyrk_(_c(uplo), _c(trans), _r(n), _r(k), C.byref(_d(alpha)), _p(a), _r(lda), C.byref(_d(beta)), _p(c), _r(ldc),
                 C.c_size_t(1), C.c_size_t(1))


def dtrsm(side, uplo, transa, diag, m, n, alpha, a, lda, b, ldb):
    lib().dtrsm_(_c(side), _c(uplo), _c(transa), _c(diag), _r(m), _r(n), C.byref(_d(alpha)), _p(a), _r(lda), _p(b), _r(ldb),
                 C.c_size_t(1), C.c_size_t(1), C.c_size_t(1), C.c_size_t(1))


def dtrmm(side, uplo, transa, diag, m, n, alpha, a, lda, b, ldb):
    lib().dtrmm_(_c(side), _c(uplo), _c(transa), _c(diag), _r(m), _r(n), C.byref(_d(alpha)), _p(a), _r(lda), _p(b), _r(ldb),
                 C.c_size_t(1), C.c_size_t(1), C.c_size_t(1), C.c_size_t(1))


def dgetrf(m, n, a, lda, ipiv, recursive=False):
    info = _i(0)
    fn = lib().dgetrf2_ if recursive else lib().dgetrf_
    fn(_r(m), _r(n), _p(a), _r(lda), _p(ipiv), C.byref(info))
    return info.value


def dlaswp(n, a, lda, k1, k2, ipiv, incx):
    lib().dlaswp_(_r(n), _p(a), _r(lda), _r(k1), _r(k2), _p(ipiv), _r(incx))


def dgetrs(trans, n, nrhs, a, lda, ipiv, b, ldb):
    info = _i(0)
    lib().dgetrs_(_c(trans), _r(n), _r(nrhs), _p(a), _r(lda), _p(ipiv), _p(b), _r(ldb), C.byref(info), C.c_size_t(1))
    return info.value


def dgesv(n, nrhs, a, lda, ipiv, b, ldb):
    info = _i(0)
    lib().dgesv_(_r(n), _r(nrhs), _p(a), _r(lda), _p(ipiv), _p(b), _r(ldb), C.byref(info))
    return info.value


def dpotrf(uplo, n, a, lda, recursive=False):
    info = _i(0)
    fn = lib().dpotrf2_ if recursive else lib().dpotrf_
    fn(_c(uplo), _r(n), _p(a), _r(lda), C.byref(info), C.c_size_t(1))
    return info.value


def dpotrs(uplo, n, nrhs, a, lda, b, ldb):
    info = _i(0)
    lib().dpotrs_(_c(uplo), _r(n), _r(nrhs), _p(a), _r(lda), _p(b), _r(ldb), C.byref(info), C.c_size_t(1))
    return info.value


def dposv(uplo, n, nrhs, a, lda, b, ldb):
    info = _i(0)
    lib().dposv_(_c(uplo), _r(n), _r(nrhs), _p(a), _r(lda), _p(b), _r(ldb), C.byref(info), C.c_size_t(1))
    return info.value


def dgeqrf(m, n, a, lda, tau, work, lwork):
    info = _i(0)
    lib().dgeqrf_(_r(m), _r(n), _p(a), _r(lda), _p(tau), _p(work), _r(lwork), C.byref(info))
    return info.value


def dgeqr2(m, n, a, lda, tau, work):
    info = _i(0)
    lib().dgeqr2_(_r(m), _r(n), _p(a), _r(lda), _p(tau), _p(work), C.byref(info))
    return info.value


def dlarft(direct, storev, n, k, v, ldv, tau, t, ldt):
    lib().dlarft_(_c(direct), _c(storev), _r(n), _r(k), _p(v), _r(ldv), _p(tau), _p(t), _r(ldt), C.c_size_t(1), C.c_size_t(1))


def dlarfb(side, trans, direct, storev, m, n, k, v, ldv, t, ldt, c, ldc, work, ldwork):
    lib().dlarfb_(_c(side), _c(trans), _c(direct), _c(storev), _r(m), _r(n), _r(k), _p(v), _r(ldv), _p(t), _r(ldt), _p(c),
                  _r(ldc), _p(work), _r(ldwork), C.c_size_t(1), C.c_size_t(1), C.c_size_t(1), C.c_size_t(1))


# ---- convenience forms used by the parity tests (shapes taken from the arrays) -----------------
def getrf(a, recursive=False):
    m, n = a.shape
    ipiv = np.zeros(max(1, min(m, n)), dtype=np.int32)
    info = dgetrf(m, n, a, _ld(a), ipiv, recursive)
    return ipiv[:min(m, n)], info


def getrs(trans, a, ipiv, b):
    return dgetrs(trans, a.shape[0], b.shape[1], a, _ld(a), np.ascontiguousarray(ipiv, dtype=np.int32), b, _ld(b))


def gesv(a, b):
    n = a.shape[0]
    ipiv = np.zeros(max(1, n), dtype=np.int32)
    info = dgesv(n, b.shape[1], a, _ld(a), ipiv, b, _ld(b))
    return ipiv[:n], info


def potrf(uplo, a, recursive=False):
    return dpotrf(uplo, a.shape[0], a, _ld(a), recursive)


def potrs(uplo, a, b):
    return dpotrs(uplo, a.shape[0], b.shape[1], a, _ld(a), b, _ld(b))


def posv(uplo, a, b):
    return dposv(uplo, a.shape[0], b.shape[1], a, _ld(a), b, _ld(b))


def geqrf(a):
    m, n = a.shape
    tau = np.zeros(max(1, min(m, n)))
    wq = np.zeros(1)
    info = dgeqrf(m, n, a, _ld(a), tau, wq, -1)
    if info != 0:
        return tau[:min(m, n)], info, wq[0]
    lwork = max(1, int(wq[0]))
    work = np.zeros(lwork)
    info = dgeqrf(m, n, a, _ld(a), tau, work, lwork)
    return tau[:min(m, n)], info, work[0]


def dgelqf(m, n, a, lda, tau, work, lwork):
    info = _i(0)
    lib().dgelqf_(_r(m), _r(n), _p(a), _r(lda), _p(tau), _p(work), _r(lwork), C.byref(info))
    return info.value


def dormlq(side, trans, m, n, k, a, lda, tau, c, ldc, work, lwork):
    info = _i(0)
    lib().dormlq_(_c(side), _c(trans), _r(m), _r(n), _r(k), _p(a), _r(lda), _p(tau), _p(c), _r(ldc), _p(work), _r(lwork),
                  C.byref(info), C.c_size_t(1), C.c_size_t(1))
    return info.value


def dgels(trans, m, n, nrhs, a, lda, b, ldb, work, lwork):
    info = _i(0)
    lib().dgels_(_c(trans), _r(m), _r(n), _r(nrhs), _p(a), _r(lda), _p(b), _r(ldb), _p(work), _r(lwork), C.byref(info),
                 C.c_size_t(1))
    return info.value


def gelqf(a):
    m, n = a.shape
    tau = np.zeros(max(1, min(m, n)))
    wq = np.zeros(1)
    info = dgelqf(m, n, a, _ld(a), tau, wq, -1)
    if info != 0:
        return tau[:min(m, n)], info
    work = np.zeros(max(1, int(wq[0])))
    return tau[:min(m, n)], dgelqf(m, n, a, _ld(a), tau, work, len(work))


def ormlq(side, trans, a, tau, c):
    m, n = c.shape
    wq = np.zeros(1)
    info = dormlq(side, trans, m, n, len(tau), a, _ld(a), tau, c, _ld(c), wq, -1)
    if info != 0:
        return info
    work = np.zeros(max(1, int(wq[0])))
    return dormlq(side, trans, m, n, len(tau), a, _ld(a), tau, c, _ld(c), work, len(work))


def gels(trans, a, b):
    """a (m x n) is overwritten by its QR / LQ factors, b (max(m,n) x nrhs) by the solution; workspace query like the reference"""
    m, n = a.shape
    wq = np.zeros(1)
    info = dgels(trans, m, n, b.shape[1], a, _ld(a), b, _ld(b), wq, -1)
    if info != 0:
        return info
    work = np.zeros(max(1, int(wq[0])))
    return dgels(trans, m, n, b.shape[1], a, _ld(a), b, _ld(b), work, len(work))


def dgeqrt(m, n, nb, a, lda, t, ldt, work):
    info = _i(0)
    lib().dgeqrt_(_r(m), _r(n), _r(nb), _p(a), _r(lda), _p(t), _r(ldt), _p(work), C.byref(info))
    return info.value


def dlatsqr(m, n, mb, nb, a, lda, t, ldt, work, lwork):
    info = _i(0)
    lib().dlatsqr_(_r(m), _r(n), _r(mb), _r(nb), _p(a), _r(lda), _p(t), _r(ldt), _p(work), _r(lwork), C.byref(info))
    return info.value


def dgeqrt3(m, n, a, lda, t, ldt):
    info = _i(0)
    lib().dgeqrt3_(_r(m), _r(n), _p(a), _r(lda), _p(t), _r(ldt), C.byref(info))
    return info.value


def dgemqrt(side, trans, m, n, k, nb, v, ldv, t, ldt, c, ldc, work):
    info = _i(0)
    lib().dgemqrt_(_c(side), _c(trans), _r(m), _r(n), _r(k), _r(nb), _p(v), _r(ldv), _p(t), _r(ldt), _p(c), _r(ldc), _p(work),
                   C.byref(info), C.c_size_t(1), C.c_size_t(1))
    return info.value


def geqrt(a, nb):
    """blocked QR keeping the T factors (DGEQRT); returns (t [nb x min(m,n)], info)"""
    m, n = a.shape
    t = np.zeros((nb, max(1, min(m, n))), order="F")
    work = np.zeros(max(1, nb * n))
    return t, dgeqrt(m, n, nb, a, _ld(a), t, _ld(t), work)


def gemqrt(side, trans, v, t, c, nb, k=None):
    m, n = c.shape
    k = t.shape[1] if k is None else k
    work = np.zeros(max(1, (n if side.upper() == "L" else m) * nb))
    return dgemqrt(side, trans, m, n, k, nb, v, _ld(v), t, _ld(t), c, _ld(c), work)


def dgerfs(trans, n, nrhs, a, lda, af, ldaf, ipiv, b, ldb, x, ldx, ferr, berr, work, iwork):
    info = _i(0)
    lib().dgerfs_(_c(trans), _r(n), _r(nrhs), _p(a), _r(lda), _p(af), _r(ldaf), _p(ipiv), _p(b), _r(ldb), _p(x), _r(ldx), _p(ferr),
                  _p(berr), _p(work), _p(iwork), C.byref(info), C.c_size_t(1))
    return info.value


def gerfs(trans, a, af, ipiv, b, x):
    """refine x in place; returns (ferr, berr, info)"""
    n, nrhs = a.shape[0], b.shape[1]
    ferr, berr = np.zeros(max(1, nrhs)), np.zeros(max(1, nrhs))
    work, iwork = np.zeros(max(1, 3 * n)), np.zeros(max(1, n), dtype=np.int32)
    info = dgerfs(trans, n, nrhs, a, _ld(a), af, _ld(af), np.ascontiguousarray(ipiv, dtype=np.int32), b, _ld(b), x, _ld(x), ferr,
                  berr, work, iwork)
    return ferr[:nrhs], berr[:nrhs], info


def dgetri(n, a, lda, ipiv, work, lwork):
    info = _i(0)
    lib().dgetri_(_r(n), _p(a), _r(lda), _p(ipiv), _p(work), _r(lwork), C.byref(info))
    return info.value


def getri(a, ipiv):
    """a (DGETRF factors) := inv(A); workspace query protocol like the reference"""
    n = a.shape[0]
    ipiv = np.ascontiguousarray(ipiv, dtype=np.int32)
    wq = np.zeros(1)
    info = dgetri(n, a, _ld(a), ipiv, wq, -1)
    if info != 0:
        return info
    work = np.zeros(max(1, int(wq[0])))
    return dgetri(n, a, _ld(a), ipiv, work, len(work))


def dorgqr(m, n, k, a, lda, tau, work, lwork):
    info = _i(0)
    lib().dorgqr_(_r(m), _r(n), _r(k), _p(a), _r(lda), _p(tau), _p(work), _r(lwork), C.byref(info))
    return info.value


def dormqr(side, trans, m, n, k, a, lda, tau, c, ldc, work, lwork):
    info = _i(0)
    lib().dormqr_(_c(side), _c(trans), _r(m), _r(n), _r(k), _p(a), _r(lda), _p(tau), _p(c), _r(ldc), _p(work), _r(lwork),
                  C.byref(info), C.c_size_t(1), C.c_size_t(1))
    return info.value


def orgqr(a, tau):
    """a (m x n, holding the reflectors) := first n columns of Q; workspace query protocol like the reference"""
    m, n = a.shape
    wq = np.zeros(1)
    info = dorgqr(m, n, len(tau), a, _ld(a), tau, wq, -1)
    if info != 0:
        return info
    work = np.zeros(max(1, int(wq[0])))
    return dorgqr(m, n, len(tau), a, _ld(a), tau, work, len(work))


def ormqr(side, trans, a, tau, c):
    m, n = c.shape
    wq = np.zeros(1)
    info = dormqr(side, trans, m, n, len(tau), a, _ld(a), tau, c, _ld(c), wq, -1)
    if info != 0:
        return info
    work = np.zeros(max(1, int(wq[0])))
    return dormqr(side, trans, m, n, len(tau), a, _ld(a), tau, c, _ld(c), work, len(work))


def geqr2(a):
    m, n = a.shape
    tau = np.zeros(max(1, min(m, n)))
    work = np.zeros(max(1, n))
    info = dgeqr2(m, n, a, _ld(a), tau, work)
    return tau[:min(m, n)], info


def larft(v, tau):
    n, k = v.shape
    t = np.zeros((k, k), order="F")
    dlarft("F", "C", n, k, v, _ld(v), np.ascontiguousarray(tau), t, _ld(t))
    return t


def larfb(side, trans, v, t, c):
    m, n = c.shape
    k = t.shape[0]
    ldw = n if side.upper() == "L" else m
    work = np.zeros((max(1, ldw), max(1, k)), order="F")
    dlarfb(side, trans, "F", "C", m, n, k, v, _ld(v), t, _ld(t), c, _ld(c), work, _ld(work))


# ---- condition estimation / expert driver ------------------------------------
def dlatrs(uplo, trans, diag, normin, a, x, cnorm):
    """x (1-D) := solution of op(A) x = scale*b; cnorm in/out; returns (scale, info)"""
    n = a.shape[0]
    scale, info = _d(0.0), _i(0)
    lib().dlatrs_(_c(uplo), _c(trans), _c(diag), _c(normin), _r(n), _p(a), _r(_ld(a)), _p(x), C.byref(scale), _p(cnorm), C.byref(info),
                  C.c_size_t(1), C.c_size_t(1), C.c_size_t(1), C.c_size_t(1))
    return scale.value, info.value


def dgecon(norm, a, anorm, n=None, lda=None):
    n = a.shape[0] if n is None else n
    lda = _ld(a) if lda is None else lda
    work, iwork = np.zeros(max(1, 4 * n)), np.zeros(max(1, n), dtype=np.int32)
    rcond, info = _d(0.0), _i(0)
    lib().dgecon_(_c(norm), _r(n), _p(a), _r(lda), C.byref(_d(anorm)), C.byref(rcond), _p(work), _p(iwork), C.byref(info), C.c_size_t(1))
    return rcond.value, info.value


def dgeequ(a):
    m, n = a.shape
    r, c = np.zeros(max(1, m)), np.zeros(max(1, n))
    rowcnd, colcnd, amax, info = _d(0.0), _d(0.0), _d(0.0), _i(0)
    lib().dgeequ_(_r(m), _r(n), _p(a), _r(_ld(a)), _p(r), _p(c), C.byref(rowcnd), C.byref(colcnd), C.byref(amax), C.byref(info))
    return r[:m], c[:n], rowcnd.value, colcnd.value, amax.value, info.value


def dgesvx(fact, trans, a, af, ipiv, equed, r, c, b, n=None, nrhs=None, lda=None, ldaf=None, ldb=None):
    """SRC/dgesvx.f; a, af, ipiv, r, c, b in/out like the reference.  Returns dict(x, rcond, ferr, berr, rpvgrw, equed, info)."""
    n = a.shape[0] if n is None else n
    nrhs = b.shape[1] if nrhs is None else nrhs
    x = np.zeros((max(1, n), max(1, nrhs)), order="F")
    ferr, berr = np.zeros(max(1, nrhs)), np.zeros(max(1, nrhs))
    work, iwork = np.zeros(max(1, 4 * n)), np.zeros(max(1, n), dtype=np.int32)
    rcond, info = _d(0.0), _i(0)
    eq = C.create_string_buffer(equed.encode(), 2)
    lib().dgesvx_(_c(fact), _c(trans), _r(n), _r(nrhs), _p(a), _r(_ld(a) if lda is None else lda), _p(af), _r(_ld(af) if ldaf is None else ldaf),
                  _p(ipiv), eq, _p(r), _p(c), _p(b), _r(_ld(b) if ldb is None else ldb), _p(x), _r(max(1, n)), C.byref(rcond), _p(ferr),
                  _p(berr), _p(work), _p(iwork), C.byref(info), C.c_size_t(1), C.c_size_t(1), C.c_size_t(1))
    return dict(x=x[:n, :nrhs], rcond=rcond.value, ferr=ferr[:nrhs], berr=berr[:nrhs], rpvgrw=work[0], equed=eq.value.decode()[:1],
                info=info.value)
