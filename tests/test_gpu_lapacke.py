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


def _ref_lapacke(libs):
    for name, L in libs:
        if name == "reference-lapacke":
            return L
    pytest.skip("oracle/_ref/liblapacke_ref.so not built")


@pytest.mark.parametrize("layout", [COL, ROW])
def test_reference_lapacke_dgecon_dgeequ_on_our_symbols(libs, layout):
    """The reference's OWN LAPACKE_dgecon / LAPACKE_dgeequ (LAPACKE/src/lapacke_dgecon.c, lapacke_dgeequ.c), compiled in place, call our
    dgecon_ / dgeequ_ (row-major goes through their transposition): same RCOND as the oracle's DGECON."""
    L = _ref_lapacke(libs)
    L.LAPACKE_dgecon.argtypes = [C.c_int, C.c_char, C.c_int, C.c_void_p, C.c_int, C.c_double, C.c_void_p]
    n = 260
    a, _ = O.random_matrix(n, n, SEED)
    lu = a.copy(order="F")
    O.dgetrf(lu)
    buf = np.array(lu, order="C" if layout == ROW else "F", copy=True)
    for norm in (b"1", b"I"):
        anorm = float(np.linalg.norm(a, 1 if norm == b"1" else np.inf))
        rc_ref, _ = O.dgecon(norm.decode(), lu, anorm)
        rcond = C.c_double(0.0)
        assert L.LAPACKE_dgecon(layout, norm, n, vp(buf), n, anorm, C.byref(rcond)) == 0
        assert abs(rcond.value - rc_ref) <= 1e-9 * rc_ref
    m2, n2 = 70, 50
    b = np.asfortranarray(np.random.default_rng(4).uniform(-1, 1, (m2, n2)) * (10.0 ** np.linspace(-4, 4, m2))[:, None])
    want = O.dgeequ(b)
    bb = np.array(b, order="C" if layout == ROW else "F", copy=True)
    r, c = np.zeros(m2), np.zeros(n2)
    rowcnd, colcnd, amax = C.c_double(0), C.c_double(0), C.c_double(0)
    assert L.LAPACKE_dgeequ(layout, m2, n2, vp(bb), n2 if layout == ROW else m2, vp(r), vp(c), C.byref(rowcnd), C.byref(colcnd), C.byref(amax)) == 0
    assert np.allclose(r, want[0], rtol=1e-15) and np.allclose(c, want[1], rtol=1e-15)
    assert (rowcnd.value, colcnd.value, amax.value) == pytest.approx(want[2:5], rel=1e-15)


@pytest.mark.parametrize("layout", [COL, ROW])
def test_reference_lapacke_dgesvx_dgerfs_on_our_symbols(libs, layout):
    L = _ref_lapacke(libs)
    n, nrhs = 180, 2
    a0, seed = O.random_matrix(n, n, SEED)
    a0 *= (10.0 ** np.linspace(-4, 4, n))[:, None]
    xact, _ = O.random_matrix(n, nrhs, seed)
    b0 = np.asfortranarray(a0 @ xact)
    order = "C" if layout == ROW else "F"
    a, b = np.array(a0, order=order, copy=True), np.array(b0, order=order, copy=True)
    af, x = np.zeros((n, n), order=order), np.zeros((n, nrhs), order=order)
    ipiv, r, c = np.zeros(n, dtype=np.int32), np.zeros(n), np.zeros(n)
    ferr, berr, rpiv = np.zeros(nrhs), np.zeros(nrhs), np.zeros(1)
    equed, rcond = C.create_string_buffer(b"N", 2), C.c_double(0.0)
    ldb = nrhs if layout == ROW else n
    L.LAPACKE_dgesvx.argtypes = [C.c_int, C.c_char, C.c_char, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p,
                                 C.c_char_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                 C.c_void_p, C.c_void_p]
    info = L.LAPACKE_dgesvx(layout, b"E", b"N", n, nrhs, vp(a), n, vp(af), n, vp(ipiv), equed, vp(r), vp(c), vp(b), ldb, vp(x), ldb,
                            C.byref(rcond), vp(ferr), vp(berr), vp(rpiv))
    a2, b2 = a0.copy(order="F"), b0.copy(order="F")
    af2, ipiv2, r2, c2 = np.zeros((n, n), order="F"), np.zeros(n, dtype=np.int32), np.zeros(n), np.zeros(n)
    ref = O.dgesvx("E", "N", a2, af2, ipiv2, "N", r2, c2, b2)
    assert info == ref["info"] == 0 and equed.value.decode()[:1] == ref["equed"]
    assert np.array_equal(ipiv, ipiv2)
    assert abs(rcond.value - ref["rcond"]) <= 1e-8 * ref["rcond"]
    assert abs(rpiv[0] - ref["rpvgrw"]) <= 1e-11 * ref["rpvgrw"]
    assert np.max(np.abs(x - xact)) / np.max(np.abs(xact)) <= max(1e-9, 2 * np.max(ferr))
    assert np.all(berr <= 4 * 2.0 ** -53 * (n + 1))
    # LAPACKE_dgerfs on the equilibrated system returned above.  The reference's row-major wrapper writes the scaled B back only for
    # FACT = 'F' (lapacke_dgesvx_work.c:124-128), so the equilibrated right-hand side is formed here for both layouts.
    b = np.array(b0 * r[:, None] if ref["equed"] in "RB" else b0, order=order, copy=True)
    xs = x / c[:, None] if ref["equed"] in "CB" else x            # DGESVX returned x scaled back; refine in the equilibrated variables
    x3 = np.array(xs * (1 + 1e-8), order=order, copy=True)
    L.LAPACKE_dgerfs.argtypes = [C.c_int, C.c_char, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                 C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    assert L.LAPACKE_dgerfs(layout, b"N", n, nrhs, vp(a), n, vp(af), n, vp(ipiv), vp(b), ldb, vp(x3), ldb, vp(ferr), vp(berr)) == 0
    assert np.max(np.abs(x3 - xs)) / np.max(np.abs(xs)) <= max(1e-9, 2 * np.max(ferr))
    assert np.all(berr <= 4 * 2.0 ** -53 * (n + 1))


@pytest.mark.parametrize("layout", [COL, ROW])
def test_reference_lapacke_dgeqrt3_on_our_symbols(libs, layout):
    L = _ref_lapacke(libs)
    m, n = 150, 60
    a, _ = O.random_matrix(m, n, SEED)
    ref = a.copy(order="F")
    t_ref, _ = O.dgeqrt(ref, n)
    order = "C" if layout == ROW else "F"
    buf, t = np.array(a, order=order, copy=True), np.zeros((n, n), order=order)
    assert L.LAPACKE_dgeqrt3(layout, m, n, vp(buf), n if layout == ROW else m, vp(t), n) == 0
    assert np.max(np.abs(buf - ref)) < 1e-11
    assert np.max(np.abs(np.triu(t) - np.triu(t_ref[:n, :n]))) < 1e-11
