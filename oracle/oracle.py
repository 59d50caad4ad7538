"""ctypes front-end to the CPU parity oracle (oracle/liboracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package lapack_b200/ never imports this.

All matrices are numpy float64 arrays in Fortran (column-major) order and are modified in place,
exactly like the reference routines.  IPIV / INFO are 1-based like LAPACK's.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


def build(force: bool = False) -> str:
    """Compile liboracle.so (and _ref/ when /root/reference exists) with the committed Makefile."""
    so = os.path.join(_HERE, "liboracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("ref_blas.c", "ref_lapack.c", "ref_check.c", "oracle.h")]
    stale = (not os.path.exists(so)) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs)
    if force or stale:
        subprocess.run(["make", "-C", _HERE, "all"], check=True, capture_output=True)
    return so


def lib() -> C.CDLL:
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.ora_dnrm2.restype = C.c_double
        _LIB.ora_dlange.restype = C.c_double
        _LIB.ora_dlansy.restype = C.c_double
        _LIB.ora_dlamch.restype = C.c_double
        _LIB.ora_dlapy2.restype = C.c_double
        _LIB.ora_ddot.restype = C.c_double
    return _LIB


def _d(a: np.ndarray):
    assert a.dtype == np.float64
    return a.ctypes.data_as(_dp)


def _i(a: np.ndarray):
    assert a.dtype == np.int32
    return a.ctypes.data_as(_ip)


def _ld(a: np.ndarray) -> int:
    """Leading dimension of a Fortran-ordered 2-D array (or a column-major view of one)."""
    if a.ndim == 1:
        return max(1, a.shape[0])
    if a.size == 0:
        return max(1, a.shape[0])
    assert a.shape[0] <= 1 or a.strides[0] == 8, "oracle expects column-major (order='F') arrays"
    return max(1, a.strides[1] // 8, a.shape[0]) if a.shape[1] > 1 else max(1, a.shape[0])


def _c(ch: str):
    return C.c_char(ch.encode())


def fmat(m: int, n: int) -> np.ndarray:
    return np.zeros((m, n), dtype=np.float64, order="F")


# ---------------------------------------------------------------- RNG / generators
def dlarnv(idist: int, iseed, n: int):
    seed = (C.c_int * 4)(*iseed)
    x = np.empty(n, dtype=np.float64)
    lib().ora_dlarnv(C.c_int(idist), seed, C.c_long(n), _d(x))
    return x, list(seed)


def random_matrix(m: int, n: int, iseed=(1988, 1989, 1990, 1991)):
    """U(-1,1) matrix filled column by column from DLARNV(2)."""
    x, seed = dlarnv(2, iseed, m * n)
    return np.asfortranarray(x.reshape((n, m)).T), seed


def spd_matrix(n: int, iseed=(1988, 1989, 1990, 1991)):
    r, seed = random_matrix(n, n, iseed)
    s = (r + r.T) * 0.5
    s[np.arange(n), np.arange(n)] += n
    return np.asfortranarray(s), seed


# ---------------------------------------------------------------- BLAS
def dgemm(ta, tb, m, n, k, alpha, a, b, beta, c):
    lib().ora_dgemm(_c(ta), _c(tb), m, n, k, C.c_double(alpha), _d(a), _ld(a), _d(b), _ld(b),
                    C.c_double(beta), _d(c), _ld(c))


def dtrsm(side, uplo, trans, diag, m, n, alpha, a, b):
    lib().ora_dtrsm(_c(side), _c(uplo), _c(trans), _c(diag), m, n, C.c_double(alpha), _d(a), _ld(a),
                    _d(b), _ld(b))


def dtrmm(side, uplo, trans, diag, m, n, alpha, a, b):
    lib().ora_dtrmm(_c(side), _c(uplo), _c(trans), _c(diag), m, n, C.c_double(alpha), _d(a), _ld(a),
                    _d(b), _ld(b))


def dsyrk(uplo, trans, n, k, alpha, a, beta, c):
    lib().ora_dsyrk(_c(uplo), _c(trans), n, k, C.c_double(alpha), _d(a), _ld(a), C.c_double(beta),
                    _d(c), _ld(c))


def idamax(x):
    return lib().ora_idamax(len(x), _d(x), 1)


def dnrm2(x):
    return lib().ora_dnrm2(len(x), _d(x), 1)


# ---------------------------------------------------------------- LU
def set_nb(getrf=64, potrf=64, geqrf=32, nx=128):
    lib().ora_set_nb(getrf, potrf, geqrf, nx)


def dlaswp(a, k1, k2, ipiv, incx):
    lib().ora_dlaswp(a.shape[1], _d(a), _ld(a), k1, k2, _i(ipiv), incx)


def dgetrf2(a):
    m, n = a.shape
    ipiv = np.zeros(min(m, n), dtype=np.int32)
    info = C.c_int(0)
    lib().ora_dgetrf2(m, n, _d(a), _ld(a), _i(ipiv), C.byref(info))
    return ipiv, info.value


def dgetrf(a):
    m, n = a.shape
    ipiv = np.zeros(min(m, n), dtype=np.int32)
    info = C.c_int(0)
    lib().ora_dgetrf(m, n, _d(a), _ld(a), _i(ipiv), C.byref(info))
    return ipiv, info.value


def dgetrs(trans, a, ipiv, b):
    n = a.shape[0]
    nrhs = b.shape[1]
    info = C.c_int(0)
    lib().ora_dgetrs(_c(trans), n, nrhs, _d(a), _ld(a), _i(ipiv), _d(b), _ld(b), C.byref(info))
    return info.value


def dgesv(a, b):
    n = a.shape[0]
    ipiv = np.zeros(n, dtype=np.int32)
    info = C.c_int(0)
    lib().ora_dgesv(n, b.shape[1], _d(a), _ld(a), _i(ipiv), _d(b), _ld(b), C.byref(info))
    return ipiv, info.value


# ---------------------------------------------------------------- Cholesky
def dpotrf2(uplo, a):
    info = C.c_int(0)
    lib().ora_dpotrf2(_c(uplo), a.shape[0], _d(a), _ld(a), C.byref(info))
    return info.value


def dpotrf(uplo, a):
    info = C.c_int(0)
    lib().ora_dpotrf(_c(uplo), a.shape[0], _d(a), _ld(a), C.byref(info))
    return info.value


def dpotrs(uplo, a, b):
    info = C.c_int(0)
    lib().ora_dpotrs(_c(uplo), a.shape[0], b.shape[1], _d(a), _ld(a), _d(b), _ld(b), C.byref(info))
    return info.value


def dposv(uplo, a, b):
    info = C.c_int(0)
    lib().ora_dposv(_c(uplo), a.shape[0], b.shape[1], _d(a), _ld(a), _d(b), _ld(b), C.byref(info))
    return info.value


# ---------------------------------------------------------------- QR
def dlarfg(alpha: float, x: np.ndarray):
    """Returns (beta, tau); x is scaled in place to v(2:n)."""
    al = C.c_double(alpha)
    tau = C.c_double(0.0)
    lib().ora_dlarfg(len(x) + 1, C.byref(al), _d(x), 1, C.byref(tau))
    return al.value, tau.value


def dgeqr2(a):
    m, n = a.shape
    tau = np.zeros(min(m, n))
    work = np.zeros(max(1, n))
    info = C.c_int(0)
    lib().ora_dgeqr2(m, n, _d(a), _ld(a), _d(tau), _d(work), C.byref(info))
    return tau, info.value


def dgeqrf(a, lwork=None):
    m, n = a.shape
    tau = np.zeros(max(1, min(m, n)))
    nb = lib().ora_ilaenv_nb(b"DGEQRF")
    if lwork is None:
        lwork = max(1, n) * nb
    work = np.zeros(max(1, lwork))
    info = C.c_int(0)
    lib().ora_dgeqrf(m, n, _d(a), _ld(a), _d(tau), _d(work), lwork, C.byref(info))
    return tau[:min(m, n)], info.value, work[0]


def dlarft(v, tau):
    n, k = v.shape
    t = fmat(k, k)
    lib().ora_dlarft(_c("F"), _c("C"), n, k, _d(v), _ld(v), _d(tau), _d(t), _ld(t))
    return t


def dlarfb(side, trans, v, t, c):
    m, n = c.shape
    k = t.shape[0]
    ldw = n if side.upper() == "L" else m
    work = fmat(max(1, ldw), max(1, k))
    lib().ora_dlarfb(_c(side), _c(trans), _c("F"), _c("C"), m, n, k, _d(v), _ld(v), _d(t), _ld(t),
                     _d(c), _ld(c), _d(work), _ld(work))


def dorgqr(a, tau, k=None):
    m, n = a.shape
    k = len(tau) if k is None else k
    lwork = max(1, n) * 32
    work = np.zeros(lwork)
    info = C.c_int(0)
    lib().ora_dorgqr(m, n, k, _d(a), _ld(a), _d(tau), _d(work), lwork, C.byref(info))
    return info.value


# ---------------------------------------------------------------- checkers (ratios; pass iff < 30)
THRESH = 30.0


def dgeqrt(a, nb):
    """blocked QR with stored T factors (SRC/dgeqrt.f); returns (t [nb x min(m,n)], info)"""
    m, n = a.shape
    t = fmat(nb, max(1, min(m, n)))
    work = np.zeros(max(1, nb * n))
    info = C.c_int(0)
    lib().ora_dgeqrt(m, n, nb, _d(a), _ld(a), _d(t), _ld(t), _d(work), C.byref(info))
    return t, info.value


def dgemqrt(side, trans, v, t, c, nb, k=None):
    m, n = c.shape
    k = t.shape[1] if k is None else k
    work = np.zeros(max(1, (n if side.upper() == "L" else m) * nb))
    info = C.c_int(0)
    lib().ora_dgemqrt(_c(side), _c(trans), m, n, k, nb, _d(v), _ld(v), _d(t), _ld(t), _d(c), _ld(c), _d(work), C.byref(info))
    return info.value


def dgerfs(trans, a, af, ipiv, b, x):
    """iterative refinement (SRC/dgerfs.f): x is improved in place; returns (ferr, berr, info)"""
    n, nrhs = a.shape[0], b.shape[1]
    ferr, berr = np.zeros(max(1, nrhs)), np.zeros(max(1, nrhs))
    work, iwork = np.zeros(max(1, 3 * n)), np.zeros(max(1, n), dtype=np.int32)
    info = C.c_int(0)
    lib().ora_dgerfs(_c(trans), n, nrhs, _d(a), _ld(a), _d(af), _ld(af), _i(np.ascontiguousarray(ipiv, dtype=np.int32)), _d(b), _ld(b),
                     _d(x), _ld(x), _d(ferr), _d(berr), _d(work), _i(iwork), C.byref(info))
    return ferr[:nrhs], berr[:nrhs], info.value


def latsqr_tcols(m, n, mb):
    """number of columns of DLATSQR's T array: N * ceil((M-N)/(MB-N)) (dlatsqr.f:100-104), N when a single DGEQRT is used"""
    if mb <= n or mb >= m:
        return max(1, n)
    return n * (-(-(m - n) // (mb - n)))


def dlatsqr(a, mb, nb):
    """tall-skinny QR (SRC/dlatsqr.f); returns (t [nb x N*blocks], info)"""
    m, n = a.shape
    t = fmat(nb, latsqr_tcols(m, n, mb))
    work = np.zeros(max(1, nb * n))
    info = C.c_int(0)
    lib().ora_dlatsqr(m, n, mb, nb, _d(a), _ld(a), _d(t), _ld(t), _d(work), C.byref(info))
    return t, info.value


def dlatrs(uplo, trans, diag, normin, a, x, cnorm):
    """SRC/dlatrs.f: x (1-D) := solution of op(A) x = scale*b; returns (scale, info); cnorm is in/out"""
    n = a.shape[0]
    scale, info = C.c_double(0.0), C.c_int(0)
    lib().ora_dlatrs(_c(uplo), _c(trans), _c(diag), _c(normin), n, _d(a), _ld(a), _d(x), C.byref(scale), _d(cnorm), C.byref(info))
    return scale.value, info.value


def dgecon(norm, a, anorm):
    """SRC/dgecon.f on DGETRF factors; returns (rcond, info)"""
    n = a.shape[0]
    work, iwork = np.zeros(max(1, 4 * n)), np.zeros(max(1, n), dtype=np.int32)
    rcond, info = C.c_double(0.0), C.c_int(0)
    lib().ora_dgecon(_c(norm), n, _d(a), _ld(a), C.c_double(anorm), C.byref(rcond), _d(work), _i(iwork), C.byref(info))
    return rcond.value, info.value


def dgeequ(a):
    m, n = a.shape
    r, c = np.zeros(max(1, m)), np.zeros(max(1, n))
    rowcnd, colcnd, amax, info = C.c_double(0.0), C.c_double(0.0), C.c_double(0.0), C.c_int(0)
    lib().ora_dgeequ(m, n, _d(a), _ld(a), _d(r), _d(c), C.byref(rowcnd), C.byref(colcnd), C.byref(amax), C.byref(info))
    return r[:m], c[:n], rowcnd.value, colcnd.value, amax.value, info.value


def dgesvx(fact, trans, a, af, ipiv, equed, r, c, b):
    """SRC/dgesvx.f; a, af, ipiv, r, c, b are in/out as in the reference.  Returns dict(x, rcond, ferr, berr, rpvgrw, equed, info)"""
    n, nrhs = a.shape[0], b.shape[1]
    x = np.zeros((max(1, n), max(1, nrhs)), order="F")
    ferr, berr = np.zeros(max(1, nrhs)), np.zeros(max(1, nrhs))
    work, iwork = np.zeros(max(1, 4 * n)), np.zeros(max(1, n), dtype=np.int32)
    rcond, info = C.c_double(0.0), C.c_int(0)
    eq = C.c_char(equed.encode())
    lib().ora_dgesvx(_c(fact), _c(trans), n, nrhs, _d(a), _ld(a), _d(af), _ld(af), _i(ipiv), C.byref(eq), _d(r), _d(c), _d(b), _ld(b),
                     _d(x), max(1, n), C.byref(rcond), _d(ferr), _d(berr), _d(work), _i(iwork), C.byref(info))
    return dict(x=x[:n, :nrhs], rcond=rcond.value, ferr=ferr[:nrhs], berr=berr[:nrhs], rpvgrw=work[0], equed=eq.value.decode(),
                info=info.value)


def dgels(trans, a, b):
    """least squares / minimum norm solve (SRC/dgels.f); a is overwritten by its QR / LQ factors, b (max(m,n) x nrhs) by the
    solution"""
    m, n = a.shape
    nrhs = b.shape[1]
    mn = min(m, n)
    lwork = max(1, mn + max(mn, nrhs, m, n) * 32)
    work = np.zeros(lwork)
    info = C.c_int(0)
    lib().ora_dgels(_c(trans), m, n, nrhs, _d(a), _ld(a), _d(b), _ld(b), _d(work), lwork, C.byref(info))
    return info.value


def dgelq2(a):
    m, n = a.shape
    tau = np.zeros(max(1, min(m, n)))
    work = np.zeros(max(1, m))
    info = C.c_int(0)
    lib().ora_dgelq2(m, n, _d(a), _ld(a), _d(tau), _d(work), C.byref(info))
    return tau[:min(m, n)], info.value


def dorml2(side, trans, a, tau, c):
    m, n = c.shape
    work = np.zeros(max(1, m, n))
    info = C.c_int(0)
    lib().ora_dorml2(_c(side), _c(trans), m, n, len(tau), _d(a), _ld(a), _d(np.ascontiguousarray(tau)), _d(c), _ld(c), _d(work),
                     C.byref(info))
    return info.value


def dtrtri(uplo, diag, a):
    info = C.c_int(0)
    lib().ora_dtrtri(_c(uplo), _c(diag), a.shape[0], _d(a), _ld(a), C.byref(info))
    return info.value


def dgetri(a, ipiv, nb=None):
    """inverse from dgetrf(a) output, in place (SRC/dgetri.f)"""
    n = a.shape[0]
    if nb is not None:
        lib().ora_set_nb_getri(nb)
    try:
        lwork = max(1, n) * 64
        work = np.zeros(lwork)
        info = C.c_int(0)
        lib().ora_dgetri(n, _d(a), _ld(a), _i(np.ascontiguousarray(ipiv, dtype=np.int32)), _d(work), lwork, C.byref(info))
    finally:
        if nb is not None:
            lib().ora_set_nb_getri(64)
    return info.value


def dormqr(side, trans, a, tau, c, k=None):
    """C := Q C, Q^T C, C Q or C Q^T with Q from dgeqrf(a, tau) (SRC/dormqr.f); c is overwritten"""
    m, n = c.shape
    k = len(tau) if k is None else k
    nw = max(1, n) if side.upper() == "L" else max(1, m)
    lwork = nw * 64 + 65 * 64
    work = np.zeros(lwork)
    info = C.c_int(0)
    lib().ora_dormqr(_c(side), _c(trans), m, n, k, _d(a), _ld(a), _d(np.ascontiguousarray(tau)), _d(c), _ld(c),
                     _d(work), lwork, C.byref(info))
    return info.value


def dget01(a, afac, ipiv):
    m, n = a.shape
    af = np.array(afac, order="F", copy=True)
    r = C.c_double(0.0)
    lib().ora_dget01(m, n, _d(a), _ld(a), _d(af), _ld(af), _i(ipiv), None, C.byref(r))
    return r.value


def dget02(trans, a, x, b):
    m, n = a.shape
    bb = np.array(b, order="F", copy=True)
    r = C.c_double(0.0)
    lib().ora_dget02(_c(trans), m, n, x.shape[1], _d(a), _ld(a), _d(x), _ld(x), _d(bb), _ld(bb), None,
                     C.byref(r))
    return r.value


def dget04(x, xact, rcond):
    r = C.c_double(0.0)
    lib().ora_dget04(x.shape[0], x.shape[1], _d(x), _ld(x), _d(xact), _ld(xact), C.c_double(rcond),
                     C.byref(r))
    return r.value


def dpot01(uplo, a, afac):
    af = np.array(afac, order="F", copy=True)
    r = C.c_double(0.0)
    lib().ora_dpot01(_c(uplo), a.shape[0], _d(a), _ld(a), _d(af), _ld(af), None, C.byref(r))
    return r.value


def dpot02(uplo, a, x, b):
    bb = np.array(b, order="F", copy=True)
    r = C.c_double(0.0)
    lib().ora_dpot02(_c(uplo), a.shape[0], x.shape[1], _d(a), _ld(a), _d(x), _ld(x), _d(bb), _ld(bb),
                     None, C.byref(r))
    return r.value


def dqrt01(a, af, tau):
    """Two ratios of TESTING/LIN/dqrt01.f for a supplied DGEQRF result (af, tau) of a (m >= n or m < n)."""
    m, n = a.shape
    lda = max(1, m)
    a_ = np.array(a, order="F", copy=True)
    af_ = np.array(af, order="F", copy=True)
    q = fmat(lda, max(1, m))
    r = fmat(lda, max(1, m, n))
    lwork = max(1, m) * 32
    work = np.zeros(lwork)
    res = np.zeros(2)
    lib().ora_dqrt01(m, n, _d(a_), _d(af_), _d(q), _d(r), lda, _d(np.ascontiguousarray(tau)), _d(work), lwork,
                     None, _d(res))
    return res
