"""CBLAS entry points (include/lapack_b200_cblas.h) on the GPU: row-major and column-major calls against the oracle's
reference BLAS (BLAS/SRC/dgemm.f, dsyrk.f, dtrsm.f, dtrmm.f) on the same data, like CBLAS/testing/c_dblat3.f runs
every case in both layouts."""
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import oracle as O  # noqa: E402

ROW, COL = 101, 102
NOTR, TR = 111, 112
UP, LO = 121, 122
NONUNIT, UNIT = 131, 132
LEFT, RIGHT = 141, 142
dp = C.POINTER(C.c_double)


@pytest.fixture(scope="module")
def L():
    import lapack_b200
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    lib = lapack_b200.lib()
    lib.cblas_dgemm.argtypes = [C.c_int] * 6 + [C.c_double, dp, C.c_int, dp, C.c_int, C.c_double, dp, C.c_int]
    lib.cblas_dsyrk.argtypes = [C.c_int] * 5 + [C.c_double, dp, C.c_int, C.c_double, dp, C.c_int]
    lib.cblas_dtrsm.argtypes = [C.c_int] * 7 + [C.c_double, dp, C.c_int, dp, C.c_int]
    lib.cblas_dtrmm.argtypes = [C.c_int] * 7 + [C.c_double, dp, C.c_int, dp, C.c_int]
    for f in (lib.cblas_dgemm, lib.cblas_dsyrk, lib.cblas_dtrsm, lib.cblas_dtrmm):
        f.restype = None
    return lib


def P(x):
    return x.ctypes.data_as(dp)


def rnd(m, n, seed):
    return O.random_matrix(m, n, seed)[0]


def rel(x, y):
    return float(np.max(np.abs(x - y))) / max(1e-300, float(np.max(np.abs(y))))


@pytest.mark.parametrize("ta", [NOTR, TR])
@pytest.mark.parametrize("tb", [NOTR, TR])
def test_cblas_dgemm_both_layouts(L, ta, tb):
    m, n, k = 150, 97, 210
    a = rnd(m, k, (1, 2, 3, 5)) if ta == NOTR else rnd(k, m, (1, 2, 3, 5))
    b = rnd(k, n, (7, 2, 3, 5)) if tb == NOTR else rnd(n, k, (7, 2, 3, 5))
    c0 = rnd(m, n, (9, 2, 3, 5))
    ref = 0.7 * (a if ta == NOTR else a.T) @ (b if tb == NOTR else b.T) - 1.3 * c0
    # column-major (Fortran-ordered arrays, ld = rows)
    c = c0.copy(order="F")
    L.cblas_dgemm(COL, ta, tb, m, n, k, 0.7, P(a), a.shape[0], P(b), b.shape[0], -1.3, P(c), m)
    assert rel(c, ref) < 1e-13
    # row-major (C-ordered arrays, ld = columns)
    ar, br, cr = np.ascontiguousarray(a), np.ascontiguousarray(b), np.array(c0, order="C", copy=True)
    L.cblas_dgemm(ROW, ta, tb, m, n, k, 0.7, P(ar), ar.shape[1], P(br), br.shape[1], -1.3, P(cr), n)
    assert rel(cr, ref) < 1e-13


@pytest.mark.parametrize("uplo", [UP, LO])
@pytest.mark.parametrize("trans", [NOTR, TR])
def test_cblas_dsyrk_both_layouts(L, uplo, trans):
    n, k = 130, 75
    a = rnd(n, k, (1, 2, 3, 5)) if trans == NOTR else rnd(k, n, (1, 2, 3, 5))
    c0 = rnd(n, n, (9, 2, 3, 5))
    full = 0.5 * (a @ a.T if trans == NOTR else a.T @ a) + 2.0 * c0
    tri = np.triu if uplo == UP else np.tril
    other = (lambda x: np.tril(x, -1)) if uplo == UP else (lambda x: np.triu(x, 1))
    c = c0.copy(order="F")
    L.cblas_dsyrk(COL, uplo, trans, n, k, 0.5, P(a), a.shape[0], 2.0, P(c), n)
    assert rel(tri(c), tri(full)) < 1e-13 and np.array_equal(other(c), other(c0))     # other triangle untouched
    ar, cr = np.ascontiguousarray(a), np.array(c0, order="C", copy=True)
    L.cblas_dsyrk(ROW, uplo, trans, n, k, 0.5, P(ar), ar.shape[1], 2.0, P(cr), n)
    assert rel(tri(cr), tri(full)) < 1e-13 and np.array_equal(other(cr), other(c0))


@pytest.mark.parametrize("side", [LEFT, RIGHT])
@pytest.mark.parametrize("uplo", [UP, LO])
@pytest.mark.parametrize("trans", [NOTR, TR])
@pytest.mark.parametrize("diag", [NONUNIT, UNIT])
def test_cblas_dtrsm_dtrmm_both_layouts(L, side, uplo, trans, diag):
    m, n = 90, 61
    na = m if side == LEFT else n
    a = rnd(na, na, (1, 2, 3, 5)) + 4.0 * np.eye(na)
    b0 = rnd(m, n, (9, 2, 3, 5))
    t = np.triu(a) if uplo == UP else np.tril(a)
    if diag == UNIT:
        t = t - np.diag(np.diag(t)) + np.eye(na)
    op = t if trans == NOTR else t.T
    mm = 1.5 * (op @ b0 if side == LEFT else b0 @ op)
    sm = 1.5 * (np.linalg.solve(op, b0) if side == LEFT else np.linalg.solve(op.T, b0.T).T)
    for layout in (COL, ROW):
        aa = a.copy(order="F") if layout == COL else np.array(a, order="C", copy=True)
        for fn, want in ((L.cblas_dtrmm, mm), (L.cblas_dtrsm, sm)):
            b = b0.copy(order="F") if layout == COL else np.array(b0, order="C", copy=True)
            fn(layout, side, uplo, trans, diag, m, n, 1.5, P(aa), na, P(b), m if layout == COL else n)
            assert rel(b, want) < 1e-11, (layout, fn)


def test_reference_cblas_wrappers_give_identical_results(L):
    """The reference's own CBLAS wrappers (CBLAS/src/cblas_dgemm.c etc., compiled in place into oracle/_ref/libcblas_ref.so,
    their dgemm_/dsyrk_/dtrmm_/dtrsm_ resolved by liblapack_b200.so) must issue exactly the Fortran call ours does:
    results bit-identical in both layouts."""
    import os
    so = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "libcblas_ref.so")
    if not os.path.exists(so):
        pytest.skip("oracle/_ref/libcblas_ref.so not built (reference tree absent at build time)")
    R = C.CDLL(so)
    R.cblas_dgemm.argtypes = L.cblas_dgemm.argtypes
    R.cblas_dsyrk.argtypes = L.cblas_dsyrk.argtypes
    R.cblas_dtrsm.argtypes = L.cblas_dtrsm.argtypes
    R.cblas_dtrmm.argtypes = L.cblas_dtrmm.argtypes
    for f in (R.cblas_dgemm, R.cblas_dsyrk, R.cblas_dtrsm, R.cblas_dtrmm):
        f.restype = None
    m, n, k = 77, 53, 91
    for layout in (ROW, COL):
        order = "C" if layout == ROW else "F"
        ld = (lambda x: x.shape[1]) if layout == ROW else (lambda x: x.shape[0])
        for ta in (NOTR, TR):
            for tb in (NOTR, TR):
                a = np.array(rnd(m, k, (1, 2, 3, 5)) if ta == NOTR else rnd(k, m, (1, 2, 3, 5)), order=order)
                b = np.array(rnd(k, n, (7, 2, 3, 5)) if tb == NOTR else rnd(n, k, (7, 2, 3, 5)), order=order)
                c0 = np.array(rnd(m, n, (9, 2, 3, 5)), order=order)
                c1, c2 = c0.copy(order=order), c0.copy(order=order)
                L.cblas_dgemm(layout, ta, tb, m, n, k, 0.7, P(a), ld(a), P(b), ld(b), -1.3, P(c1), ld(c1))
                R.cblas_dgemm(layout, ta, tb, m, n, k, 0.7, P(a), ld(a), P(b), ld(b), -1.3, P(c2), ld(c2))
                assert np.array_equal(c1, c2), (layout, ta, tb)
        for uplo in (UP, LO):
            for trans in (NOTR, TR):
                a = np.array(rnd(n, k, (1, 2, 3, 5)) if trans == NOTR else rnd(k, n, (1, 2, 3, 5)), order=order)
                c0 = np.array(rnd(n, n, (9, 2, 3, 5)), order=order)
                c1, c2 = c0.copy(order=order), c0.copy(order=order)
                L.cblas_dsyrk(layout, uplo, trans, n, k, 0.5, P(a), ld(a), 2.0, P(c1), n)
                R.cblas_dsyrk(layout, uplo, trans, n, k, 0.5, P(a), ld(a), 2.0, P(c2), n)
                assert np.array_equal(c1, c2), (layout, uplo, trans)
        for side in (LEFT, RIGHT):
            na = m if side == LEFT else n
            a = np.array(rnd(na, na, (1, 2, 3, 5)) + 4.0 * np.eye(na), order=order)
            for uplo in (UP, LO):
                for trans in (NOTR, TR):
                    for diag in (NONUNIT, UNIT):
                        for ours, theirs in ((L.cblas_dtrsm, R.cblas_dtrsm), (L.cblas_dtrmm, R.cblas_dtrmm)):
                            b0 = np.array(rnd(m, n, (9, 2, 3, 5)), order=order)
                            b1, b2 = b0.copy(order=order), b0.copy(order=order)
                            ours(layout, side, uplo, trans, diag, m, n, 1.5, P(a), na, P(b1), ld(b1))
                            theirs(layout, side, uplo, trans, diag, m, n, 1.5, P(a), na, P(b2), ld(b2))
                            assert np.array_equal(b1, b2), (layout, side, uplo, trans, diag)
