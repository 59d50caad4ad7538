"""LAPACKE layer on the GPU: our LAPACKE_* entry points and the REFERENCE's own LAPACKE C layer (compiled in place
from /root/reference into oracle/_ref/liblapacke_ref.so, bound to our Fortran symbols) must agree with the oracle for
both matrix layouts (LAPACKE/src/lapacke_dgetrf_work.c:45-70 row-major semantics)."""
import ctypes as C
import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from oracle import oracle as O  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SEED = (1988, 1989, 1990, 1991)
ROW, COL = 101, 102


def vp(x):
    return x.ctypes.data_as(C.c_void_p)


@pytest.fixture(scope="module")
def libs():
    import lapack_b200
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    ours = lapack_b200.lib()
    ours.lb200_set_xerbla_mode(2)
    out = [("ours", ours)]
    so = os.path.join(ROOT, "oracle", "_ref", "liblapacke_ref.so")
    if os.path.exists(so):
        out.append(("reference-lapacke", C.CDLL(so, mode=C.RTLD_GLOBAL)))
    return out


@pytest.mark.parametrize("layout", [COL, ROW])
def test_lapacke_dgetrf_dgetrs_dgesv(libs, layout):
    for name, L in libs:
        for (m, n) in ((40, 40), (300, 200), (200, 300)):
            a, seed = O.random_matrix(m, n, SEED)
            ref = a.copy(order="F")
            ipiv_ref, _ = O.dgetrf(ref)
            buf = np.array(a, order="C" if layout == ROW else "F", copy=True)
            lda = n if layout == ROW else m
            ipiv = np.zeros(min(m, n), dtype=np.int32)
            assert L.LAPACKE_dgetrf(layout, m, n, vp(buf), lda, vp(ipiv)) == 0, name
            assert np.array_equal(ipiv, ipiv_ref), name
            assert np.max(np.abs(buf - ref)) < 1e-11, name
        n, nrhs = 250, 3
        a, seed = O.random_matrix(n, n, SEED)
        x, _ = O.random_matrix(n, nrhs, seed)
        b = a @ x
        abuf = np.array(a, order="C" if layout == ROW else "F", copy=True)
        bbuf = np.array(b, order="C" if layout == ROW else "F", copy=True)
        ipiv = np.zeros(n, dtype=np.int32)
        assert L.LAPACKE_dgesv(layout, n, nrhs, vp(abuf), n, vp(ipiv), vp(bbuf), nrhs if layout == ROW else n) == 0
        assert np.max(np.abs(bbuf - x)) < 1e-10, name
        b2 = np.array(b, order="C" if layout == ROW else "F", copy=True)
        assert L.LAPACKE_dgetrs(layout, C.c_char(b"N"), n, nrhs, vp(abuf), n, vp(ipiv), vp(b2), nrhs if layout == ROW else n) == 0
        assert np.max(np.abs(b2 - x)) < 1e-10, name


@pytest.mark.parametrize("layout", [COL, ROW])
@pytest.mark.parametrize("uplo", ["L", "U"])
def test_lapacke_dpotrf_dposv(libs, layout, uplo):
    for name, L in libs:
        n, nrhs = 220, 2
        s, seed = O.spd_matrix(n, SEED)
        ref = s.copy(order="F")
        assert O.dpotrf(uplo, ref) == 0
        buf = np.array(s, order="C" if layout == ROW else "F", copy=True)
        assert L.LAPACKE_dpotrf(layout, C.c_char(uplo.encode()), n, vp(buf), n) == 0, name
        tri = np.tril if uplo == "L" else np.triu
        assert np.max(np.abs(tri(buf) - tri(ref))) < 1e-12, name
        x, _ = O.random_matrix(n, nrhs, seed)
        b = s @ x
        abuf = np.array(s, order="C" if layout == ROW else "F", copy=True)
        bbuf = np.array(b, order="C" if layout == ROW else "F", copy=True)
        assert L.LAPACKE_dposv(layout, C.c_char(uplo.encode()), n, nrhs, vp(abuf), n, vp(bbuf), nrhs if layout == ROW else n) == 0
        assert np.max(np.abs(bbuf - x)) < 1e-10, name


@pytest.mark.parametrize("layout", [COL, ROW])
def test_lapacke_dgeqrf(libs, layout):
    for name, L in libs:
        m, n = 260, 180
        a, _ = O.random_matrix(m, n, SEED)
        ref = a.copy(order="F")
        tau_ref, _, _ = O.dgeqrf(ref)
        buf = np.array(a, order="C" if layout == ROW else "F", copy=True)
        tau = np.zeros(n)
        assert L.LAPACKE_dgeqrf(layout, m, n, vp(buf), n if layout == ROW else m, vp(tau)) == 0, name
        assert np.max(np.abs(buf - ref)) < 1e-10 and np.max(np.abs(tau - tau_ref)) < 1e-11, name


def test_lapacke_nancheck_and_singular(libs):
    for name, L in libs:
        a, _ = O.random_matrix(50, 50, SEED)
        a[:, 9] = 0.0
        ipiv = np.zeros(50, dtype=np.int32)
        buf = a.copy(order="F")
        assert L.LAPACKE_dgetrf(COL, 50, 50, vp(buf), 50, vp(ipiv)) == 10, name        # INFO = zero-pivot column
        buf = a.copy(order="F")
        buf[3, 4] = np.nan
        assert L.LAPACKE_dgetrf(COL, 50, 50, vp(buf), 50, vp(ipiv)) == -4, name        # NaN pre-check


@pytest.mark.parametrize("layout", [COL, ROW])
def test_lapacke_dorgqr_dormqr(libs, layout):
    """LAPACKE_dorgqr / LAPACKE_dormqr (ours, and the reference's wrappers bound to our dorgqr_ / dormqr_) vs the oracle."""
    dp = C.POINTER(C.c_double)
    m, n, nc = 130, 70, 9
    a, _ = O.random_matrix(m, n, SEED)
    af = a.copy(order="F")
    tau, info, _ = O.dgeqrf(af)
    q_ref = af.copy(order="F")
    assert O.dorgqr(q_ref, tau) == 0
    c0, _ = O.random_matrix(m, nc, (3, 5, 7, 9))
    order = "C" if layout == ROW else "F"
    for name, L in libs:
        L.LAPACKE_dorgqr.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, dp, C.c_int, dp]
        L.LAPACKE_dormqr.argtypes = [C.c_int, C.c_char, C.c_char, C.c_int, C.c_int, C.c_int, dp, C.c_int, dp, dp, C.c_int]
        q = np.array(af, order=order, copy=True)
        assert L.LAPACKE_dorgqr(layout, m, n, n, q.ctypes.data_as(dp), n if layout == ROW else m, tau.ctypes.data_as(dp)) == 0, name
        assert np.max(np.abs(q - q_ref)) < 1e-12, name
        for side, trans in (("L", "T"), ("L", "N")):
            c_ref = c0.copy(order="F")
            assert O.dormqr(side, trans, af, tau, c_ref) == 0
            abuf = np.array(af, order=order, copy=True)
            cbuf = np.array(c0, order=order, copy=True)
            rc = L.LAPACKE_dormqr(layout, side.encode(), trans.encode(), m, nc, n, abuf.ctypes.data_as(dp),
                                  n if layout == ROW else m, tau.ctypes.data_as(dp), cbuf.ctypes.data_as(dp),
                                  nc if layout == ROW else m)
            assert rc == 0, name
            assert np.max(np.abs(cbuf - c_ref)) < 1e-12, (name, side, trans)


@pytest.mark.parametrize("layout", [COL, ROW])
def test_lapacke_dgetri(libs, layout):
    n = 180
    a, _ = O.random_matrix(n, n, SEED)
    lu = a.copy(order="F")
    ipiv, _ = O.dgetrf(lu)
    ref = lu.copy(order="F")
    assert O.dgetri(ref, ipiv) == 0
    for name, L in libs:
        buf = np.array(lu, order="C" if layout == ROW else "F", copy=True)
        assert L.LAPACKE_dgetri(layout, n, vp(buf), n, vp(ipiv)) == 0, name
        assert np.max(np.abs(buf - ref)) < 1e-10 * np.max(np.abs(ref)), name


@pytest.mark.parametrize("layout", [COL, ROW])
def test_lapacke_dgels(libs, layout):
    for (m, n, nrhs, trans) in ((150, 80, 3, b"N"), (80, 150, 2, b"N"), (150, 80, 2, b"T")):
        a, _ = O.random_matrix(m, n, SEED)
        b, _ = O.random_matrix(max(m, n), nrhs, (3, 5, 7, 9))
        a_ref, x_ref = a.copy(order="F"), b.copy(order="F")
        assert O.dgels(trans.decode(), a_ref, x_ref) == 0
        rows = n if trans == b"N" else m
        order = "C" if layout == ROW else "F"
        for name, L in libs:
            L.LAPACKE_dgels.argtypes = [C.c_int, C.c_char, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int]
            abuf, bbuf = np.array(a, order=order, copy=True), np.array(b, order=order, copy=True)
            rc = L.LAPACKE_dgels(layout, trans, m, n, nrhs, vp(abuf), n if layout == ROW else m, vp(bbuf),
                                 nrhs if layout == ROW else max(m, n))
            assert rc == 0, name
            assert np.max(np.abs(bbuf[:rows] - x_ref[:rows])) < 1e-10 * np.max(np.abs(x_ref[:rows])), (name, m, n, trans)


@pytest.mark.parametrize("layout", [COL, ROW])
def test_lapacke_dgelqf_dormlq(libs, layout):
    dp = C.POINTER(C.c_double)
    m, n, nc = 70, 130, 9
    a, _ = O.random_matrix(m, n, SEED)
    ref = a.copy(order="F")
    tau_ref, _ = O.dgelq2(ref)
    c0, _ = O.random_matrix(n, nc, (3, 5, 7, 9))
    c_ref = c0.copy(order="F")
    assert O.dorml2("L", "T", ref, tau_ref, c_ref) == 0
    order = "C" if layout == ROW else "F"
    for name, L in libs:
        L.LAPACKE_dgelqf.argtypes = [C.c_int, C.c_int, C.c_int, dp, C.c_int, dp]
        L.LAPACKE_dormlq.argtypes = [C.c_int, C.c_char, C.c_char, C.c_int, C.c_int, C.c_int, dp, C.c_int, dp, dp, C.c_int]
        buf = np.array(a, order=order, copy=True)
        tau = np.zeros(m)
        assert L.LAPACKE_dgelqf(layout, m, n, buf.ctypes.data_as(dp), n if layout == ROW else m, tau.ctypes.data_as(dp)) == 0, name
        assert np.max(np.abs(buf - ref)) < 1e-12 and np.max(np.abs(tau - tau_ref)) < 1e-12, name
        cbuf = np.array(c0, order=order, copy=True)
        rc = L.LAPACKE_dormlq(layout, b"L", b"T", n, nc, m, buf.ctypes.data_as(dp), n if layout == ROW else m,
                              tau.ctypes.data_as(dp), cbuf.ctypes.data_as(dp), nc if layout == ROW else n)
        assert rc == 0, name
        assert np.max(np.abs(cbuf - c_ref)) < 1e-12, name
