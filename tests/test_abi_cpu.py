"""CPU-only checks of the drop-in boundary: the shared library loads without a GPU, exports every symbol the
headers declare, and reproduces the reference's argument checking / XERBLA positions (TESTING/LIN/derrge.f:133-184,
derrpo.f:133-178, derrqr.f:118-130; BLAS/TESTING/dblat3.f DCHKE).  No compute call is made here."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lb():
    import lapack_b200
    L = lapack_b200.lib()
    L.lb200_set_xerbla_mode(2)
    return lapack_b200


def declared_symbols():
    names = set()
    for h in ("lapack_b200.h", "lapack_b200_f77.h", "lapack_b200_f77_64.h", "lapack_b200_lapacke.h", "lapack_b200_cblas.h"):
        src = open(os.path.join(ROOT, "include", h)).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        for m in re.finditer(r"\b(lb200_\w+|LAPACKE_\w+|cblas_\w+|[a-z0-9]+(?:_64)?_)\s*\(", src):
            n = m.group(1)
            if n.startswith(("lb200_", "LAPACKE_", "cblas_")) or re.fullmatch(r"(d[a-z0-9]+|xerbla|lsame)(_64)?_", n):
                names.add(n)
    return sorted(names)


def test_library_exports_every_declared_symbol(lb):
    L = lb.lib()
    names = declared_symbols()
    assert len(names) > 70
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing


def test_reference_lapacke_binds_to_our_symbols(lb):
    """The reference's own LAPACKE C layer (compiled from /root/reference by oracle/Makefile) resolves its
    Fortran symbols against liblapack_b200.so -- the link-time drop-in mechanism of SRC/VARIANTS/README:66-78."""
    so = os.path.join(ROOT, "oracle", "_ref", "liblapacke_ref.so")
    if not os.path.exists(so):
        pytest.skip("oracle/_ref not built (reference tree absent)")
    lb.lib()                                  # RTLD_GLOBAL: dgetrf_ ... become visible
    R = C.CDLL(so, mode=C.RTLD_GLOBAL)
    a = np.zeros((2, 2), order="F")
    ipiv = np.zeros(2, dtype=np.int32)
    # illegal lda -> dgetrf_ INFO=-4 -> LAPACKE shifts to -5 (lapacke_dgetrf_work.c:42-44); no GPU work happens
    assert R.LAPACKE_dgetrf_work(102, 2, 2, a.ctypes.data_as(C.c_void_p), 1, ipiv.ctypes.data_as(C.c_void_p)) == -5
    name = C.create_string_buffer(40)
    info = C.c_int(0)
    lb.lib().lb200_last_xerbla(name, C.byref(info))
    assert name.value == b"DGETRF" and info.value == 4


def _last(lb):
    name = C.create_string_buffer(40)
    info = C.c_int(0)
    cnt = lb.lib().lb200_last_xerbla(name, C.byref(info))
    return name.value.decode(), info.value, cnt


def expect(lb, fn, routine, pos):
    lb.lib().lb200_clear_xerbla()
    ret = fn()
    name, info, cnt = _last(lb)
    assert cnt == 1 and name == routine and info == pos, (routine, pos, name, info, cnt)
    return ret


A = np.zeros((4, 4), order="F")
B = np.zeros((4, 4), order="F")
IP = np.zeros(4, dtype=np.int32)
W = np.zeros(16)
TAU = np.zeros(4)


def test_lu_error_exits(lb):
    f = lb.f77
    assert expect(lb, lambda: f.dgetrf(-1, 0, A, 1, IP), "DGETRF", 1) == -1
    assert expect(lb, lambda: f.dgetrf(0, -1, A, 1, IP), "DGETRF", 2) == -2
    assert expect(lb, lambda: f.dgetrf(2, 1, A, 1, IP), "DGETRF", 4) == -4
    assert expect(lb, lambda: f.dgetrf(-1, 0, A, 1, IP, True), "DGETRF2", 1) == -1
    assert expect(lb, lambda: f.dgetrf(2, 1, A, 1, IP, True), "DGETRF2", 4) == -4
    assert expect(lb, lambda: f.dgetrs("/", 0, 0, A, 1, IP, B, 1), "DGETRS", 1) == -1
    assert expect(lb, lambda: f.dgetrs("N", -1, 0, A, 1, IP, B, 1), "DGETRS", 2) == -2
    assert expect(lb, lambda: f.dgetrs("N", 0, -1, A, 1, IP, B, 1), "DGETRS", 3) == -3
    assert expect(lb, lambda: f.dgetrs("N", 2, 1, A, 1, IP, B, 2), "DGETRS", 5) == -5
    assert expect(lb, lambda: f.dgetrs("N", 2, 1, A, 2, IP, B, 1), "DGETRS", 8) == -8
    assert expect(lb, lambda: f.dgesv(-1, 0, A, 1, IP, B, 1), "DGESV", 1) == -1
    assert expect(lb, lambda: f.dgesv(0, -1, A, 1, IP, B, 1), "DGESV", 2) == -2
    assert expect(lb, lambda: f.dgesv(2, 1, A, 1, IP, B, 2), "DGESV", 4) == -4
    assert expect(lb, lambda: f.dgesv(2, 1, A, 2, IP, B, 1), "DGESV", 7) == -7


def test_cholesky_error_exits(lb):
    f = lb.f77
    assert expect(lb, lambda: f.dpotrf("/", 0, A, 1), "DPOTRF", 1) == -1
    assert expect(lb, lambda: f.dpotrf("U", -1, A, 1), "DPOTRF", 2) == -2
    assert expect(lb, lambda: f.dpotrf("U", 2, A, 1), "DPOTRF", 4) == -4
    assert expect(lb, lambda: f.dpotrf("U", 2, A, 1, True), "DPOTRF2", 4) == -4
    assert expect(lb, lambda: f.dpotrs("/", 0, 0, A, 1, B, 1), "DPOTRS", 1) == -1
    assert expect(lb, lambda: f.dpotrs("U", -1, 0, A, 1, B, 1), "DPOTRS", 2) == -2
    assert expect(lb, lambda: f.dpotrs("U", 0, -1, A, 1, B, 1), "DPOTRS", 3) == -3
    assert expect(lb, lambda: f.dpotrs("U", 2, 1, A, 1, B, 2), "DPOTRS", 5) == -5
    assert expect(lb, lambda: f.dpotrs("U", 2, 1, A, 2, B, 1), "DPOTRS", 7) == -7
    assert expect(lb, lambda: f.dposv("/", 0, 0, A, 1, B, 1), "DPOSV", 1) == -1
    assert expect(lb, lambda: f.dposv("U", 2, 0, A, 1, B, 2), "DPOSV", 5) == -5
    assert expect(lb, lambda: f.dposv("U", 2, 0, A, 2, B, 1), "DPOSV", 7) == -7


def test_qr_error_exits_and_query(lb):
    f = lb.f77
    assert expect(lb, lambda: f.dgeqrf(-1, 0, A, 1, TAU, W, 1), "DGEQRF", 1) == -1
    assert expect(lb, lambda: f.dgeqrf(0, -1, A, 1, TAU, W, 1), "DGEQRF", 2) == -2
    assert expect(lb, lambda: f.dgeqrf(2, 1, A, 1, TAU, W, 1), "DGEQRF", 4) == -4
    assert expect(lb, lambda: f.dgeqrf(1, 2, A, 1, TAU, W, 1), "DGEQRF", 7) == -7
    assert expect(lb, lambda: f.dgeqr2(-1, 0, A, 1, TAU, W), "DGEQR2", 1) == -1
    assert expect(lb, lambda: f.dgeqr2(2, 1, A, 1, TAU, W), "DGEQR2", 4) == -4
    wq = np.zeros(1)
    assert f.dgeqrf(300, 200, A, 300, TAU, wq, -1) == 0 and wq[0] == 200 * 32    # dgeqrf.f:197-204
    assert f.dgeqrf(0, 5, A, 1, TAU, wq, -1) == 0 and wq[0] == 1


def test_orgqr_ormqr_error_exits_and_query(lb):
    """TESTING/LIN/derrqr.f:177-259: XERBLA positions of DORGQR and DORMQR."""
    f = lb.f77
    X, AF = TAU, B
    for args, pos in (((-1, 0, 0, A, 1, X, W, 1), 1), ((0, -1, 0, A, 1, X, W, 1), 2), ((1, 2, 0, A, 1, X, W, 2), 2),
                      ((0, 0, -1, A, 1, X, W, 1), 3), ((1, 1, 2, A, 1, X, W, 1), 3), ((2, 2, 0, A, 1, X, W, 2), 5),
                      ((2, 2, 0, A, 2, X, W, 1), 8)):
        assert expect(lb, lambda a=args: f.dorgqr(*a), "DORGQR", pos) == -pos
    for args, pos in ((("/", "N", 0, 0, 0, A, 1, X, AF, 1, W, 1), 1), (("L", "/", 0, 0, 0, A, 1, X, AF, 1, W, 1), 2),
                      (("L", "N", -1, 0, 0, A, 1, X, AF, 1, W, 1), 3), (("L", "N", 0, -1, 0, A, 1, X, AF, 1, W, 1), 4),
                      (("L", "N", 0, 0, -1, A, 1, X, AF, 1, W, 1), 5), (("L", "N", 0, 1, 1, A, 1, X, AF, 1, W, 1), 5),
                      (("R", "N", 1, 0, 1, A, 1, X, AF, 1, W, 1), 5), (("L", "N", 2, 1, 0, A, 1, X, AF, 2, W, 1), 7),
                      (("R", "N", 1, 2, 0, A, 1, X, AF, 1, W, 1), 7), (("L", "N", 2, 1, 0, A, 2, X, AF, 1, W, 1), 10),
                      (("L", "N", 1, 2, 0, A, 1, X, AF, 1, W, 1), 12), (("R", "N", 2, 1, 0, A, 1, X, AF, 2, W, 1), 12)):
        assert expect(lb, lambda a=args: f.dormqr(*a), "DORMQR", pos) == -pos
    wq = np.zeros(1)
    assert f.dorgqr(300, 200, 200, A, 300, TAU, wq, -1) == 0 and wq[0] == 200 * 32         # dorgqr.f:163-165
    assert f.dormqr("L", "T", 300, 7, 200, A, 300, TAU, B, 300, wq, -1) == 0 and wq[0] == 7 * 32 + 65 * 32   # dormqr.f:240-244
    assert f.dormqr("R", "N", 7, 300, 200, A, 300, TAU, B, 7, wq, -1) == 0 and wq[0] == 7 * 32 + 65 * 32
    assert f.dormqr("L", "N", 0, 0, 0, A, 1, TAU, B, 1, wq, 1) == 0 and wq[0] == 1             # quick return, no GPU needed


def test_geqrt_gemqrt_error_exits(lb):
    """TESTING/LIN/derrqrt.f:117-130 (DGEQRT) and :168-199 (DGEMQRT)."""
    f = lb.f77
    T = np.zeros((4, 4), order="F")
    for args, pos in (((-1, 0, 1, A, 1, T, 1, W), 1), ((0, -1, 1, A, 1, T, 1, W), 2), ((0, 0, 0, A, 1, T, 1, W), 3),
                      ((2, 1, 1, A, 1, T, 1, W), 5), ((2, 2, 2, A, 2, T, 1, W), 7)):
        assert expect(lb, lambda a=args: f.dgeqrt(*a), "DGEQRT", pos) == -pos
    for args, pos in ((("/", "N", 0, 0, 0, 1, A, 1, T, 1, B, 1, W), 1), (("L", "/", 0, 0, 0, 1, A, 1, T, 1, B, 1, W), 2),
                      (("L", "N", -1, 0, 0, 1, A, 1, T, 1, B, 1, W), 3), (("L", "N", 0, -1, 0, 1, A, 1, T, 1, B, 1, W), 4),
                      (("L", "N", 0, 0, -1, 1, A, 1, T, 1, B, 1, W), 5), (("R", "N", 0, 0, -1, 1, A, 1, T, 1, B, 1, W), 5),
                      (("L", "N", 0, 0, 0, 0, A, 1, T, 1, B, 1, W), 6), (("R", "N", 1, 2, 1, 1, A, 1, T, 1, B, 1, W), 8),
                      (("L", "N", 2, 1, 1, 1, A, 1, T, 1, B, 1, W), 8), (("R", "N", 1, 1, 1, 1, A, 1, T, 0, B, 1, W), 10),
                      (("L", "N", 1, 1, 1, 1, A, 1, T, 1, B, 0, W), 12)):
        assert expect(lb, lambda a=args: f.dgemqrt(*a), "DGEMQRT", pos) == -pos


def test_gels_gelqf_ormlq_error_exits_and_query(lb):
    """TESTING/LIN/derrls.f:117-139 (DGELS), derrlq.f (DGELQF / DORMLQ positions as in dgelqf.f:172-186, dormlq.f:218-236)."""
    f = lb.f77
    for args, pos in ((("/", 0, 0, 0, A, 1, B, 1, W, 1), 1), (("N", -1, 0, 0, A, 1, B, 1, W, 1), 2),
                      (("N", 0, -1, 0, A, 1, B, 1, W, 1), 3), (("N", 0, 0, -1, A, 1, B, 1, W, 1), 4),
                      (("N", 2, 0, 0, A, 1, B, 2, W, 2), 6), (("N", 2, 0, 0, A, 2, B, 1, W, 2), 8),
                      (("N", 0, 2, 0, A, 1, B, 1, W, 2), 8), (("N", 1, 1, 0, A, 1, B, 1, W, 1), 10)):
        assert expect(lb, lambda a=args: f.dgels(*a), "DGELS", pos) == -pos
    assert expect(lb, lambda: f.dgelqf(-1, 0, A, 1, TAU, W, 1), "DGELQF", 1) == -1
    assert expect(lb, lambda: f.dgelqf(0, -1, A, 1, TAU, W, 1), "DGELQF", 2) == -2
    assert expect(lb, lambda: f.dgelqf(2, 1, A, 1, TAU, W, 2), "DGELQF", 4) == -4
    assert expect(lb, lambda: f.dgelqf(2, 1, A, 2, TAU, W, 1), "DGELQF", 7) == -7
    assert expect(lb, lambda: f.dormlq("/", "N", 0, 0, 0, A, 1, TAU, B, 1, W, 1), "DORMLQ", 1) == -1
    assert expect(lb, lambda: f.dormlq("L", "N", 0, 1, 1, A, 1, TAU, B, 1, W, 1), "DORMLQ", 5) == -5
    assert expect(lb, lambda: f.dormlq("L", "N", 2, 0, 2, A, 1, TAU, B, 2, W, 1), "DORMLQ", 7) == -7
    assert expect(lb, lambda: f.dormlq("L", "N", 2, 1, 0, A, 1, TAU, B, 1, W, 1), "DORMLQ", 10) == -10
    assert expect(lb, lambda: f.dormlq("L", "N", 1, 2, 0, A, 1, TAU, B, 1, W, 1), "DORMLQ", 12) == -12
    wq = np.zeros(1)
    assert f.dgels("N", 300, 200, 7, A, 300, B, 300, wq, -1) == 0 and wq[0] == 200 + 200 * 32      # dgels.f:284-285
    assert f.dgelqf(200, 300, A, 200, TAU, wq, -1) == 0 and wq[0] == 200 * 32
    assert f.dormlq("L", "T", 300, 7, 200, A, 200, TAU, B, 300, wq, -1) == 0 and wq[0] == 7 * 32 + 65 * 32


def test_gerfs_error_exits(lb):
    """TESTING/LIN/derrge.f:189-216."""
    f = lb.f77
    R1, R2, IW = np.zeros(4), np.zeros(4), np.zeros(4, dtype=np.int32)
    X = np.zeros((4, 4), order="F")
    for args, pos in ((("/", 0, 0, A, 1, B, 1, IP, B, 1, X, 1), 1), (("N", -1, 0, A, 1, B, 1, IP, B, 1, X, 1), 2),
                      (("N", 0, -1, A, 1, B, 1, IP, B, 1, X, 1), 3), (("N", 2, 1, A, 1, B, 2, IP, B, 2, X, 2), 5),
                      (("N", 2, 1, A, 2, B, 1, IP, B, 2, X, 2), 7), (("N", 2, 1, A, 2, B, 2, IP, B, 1, X, 2), 10),
                      (("N", 2, 1, A, 2, B, 2, IP, B, 2, X, 1), 12)):
        assert expect(lb, lambda a=args: f.dgerfs(*a, R1, R2, W, IW), "DGERFS", pos) == -pos
    R1[:] = 5.0
    assert f.dgerfs("N", 0, 2, A, 1, B, 1, IP, B, 1, X, 1, R1, R2, W, IW) == 0 and np.all(R1[:2] == 0.0)     # quick return


def test_getri_error_exits_and_query(lb):
    """TESTING/LIN/derrge.f:157-165 (positions 1 and 3) plus the LWORK check and query of dgetri.f:152-170."""
    f = lb.f77
    assert expect(lb, lambda: f.dgetri(-1, A, 1, IP, W, 1), "DGETRI", 1) == -1
    assert expect(lb, lambda: f.dgetri(2, A, 1, IP, W, 2), "DGETRI", 3) == -3
    assert expect(lb, lambda: f.dgetri(2, A, 2, IP, W, 1), "DGETRI", 6) == -6
    wq = np.zeros(1)
    assert f.dgetri(300, A, 300, IP, wq, -1) == 0 and wq[0] == 300 * 64
    assert f.dgetri(0, A, 1, IP, wq, 1) == 0 and wq[0] == 1


def test_quick_returns_need_no_gpu(lb):
    """M==0 or N==0 return before any device work (dgetrf.f:159, dpotrf.f:161, dgeqrf.f:209-212)."""
    f = lb.f77
    lb.lib().lb200_clear_xerbla()
    assert f.dgetrf(0, 4, A, 1, IP) == 0 and f.dgetrf(4, 0, A, 4, IP) == 0
    assert f.dpotrf("L", 0, A, 1) == 0
    assert f.dgetrs("N", 0, 1, A, 1, IP, B, 1) == 0 and f.dgetrs("N", 4, 0, A, 4, IP, B, 4) == 0
    w = np.zeros(1)
    assert f.dgeqrf(0, 3, A, 1, TAU, w, 3) == 0 and w[0] == 1
    assert _last(lb)[2] == 0


def test_blas3_error_exits(lb):
    f = lb.f77
    expect(lb, lambda: f.dgemm("/", "N", 0, 0, 0, 1.0, A, 1, B, 1, 0.0, W, 1), "DGEMM", 1)
    expect(lb, lambda: f.dgemm("N", "/", 0, 0, 0, 1.0, A, 1, B, 1, 0.0, W, 1), "DGEMM", 2)
    expect(lb, lambda: f.dgemm("N", "N", -1, 0, 0, 1.0, A, 1, B, 1, 0.0, W, 1), "DGEMM", 3)
    expect(lb, lambda: f.dgemm("N", "N", 0, -1, 0, 1.0, A, 1, B, 1, 0.0, W, 1), "DGEMM", 4)
    expect(lb, lambda: f.dgemm("N", "N", 0, 0, -1, 1.0, A, 1, B, 1, 0.0, W, 1), "DGEMM", 5)
    expect(lb, lambda: f.dgemm("N", "N", 2, 0, 0, 1.0, A, 1, B, 1, 0.0, W, 2), "DGEMM", 8)
    expect(lb, lambda: f.dgemm("N", "N", 0, 0, 2, 1.0, A, 1, B, 1, 0.0, W, 1), "DGEMM", 10)
    expect(lb, lambda: f.dgemm("N", "N", 2, 0, 0, 1.0, A, 2, B, 1, 0.0, W, 1), "DGEMM", 13)
    expect(lb, lambda: f.dtrsm("/", "U", "N", "N", 0, 0, 1.0, A, 1, B, 1), "DTRSM", 1)
    expect(lb, lambda: f.dtrsm("L", "/", "N", "N", 0, 0, 1.0, A, 1, B, 1), "DTRSM", 2)
    expect(lb, lambda: f.dtrsm("L", "U", "/", "N", 0, 0, 1.0, A, 1, B, 1), "DTRSM", 3)
    expect(lb, lambda: f.dtrsm("L", "U", "N", "/", 0, 0, 1.0, A, 1, B, 1), "DTRSM", 4)
    expect(lb, lambda: f.dtrsm("L", "U", "N", "N", -1, 0, 1.0, A, 1, B, 1), "DTRSM", 5)
    expect(lb, lambda: f.dtrsm("L", "U", "N", "N", 0, -1, 1.0, A, 1, B, 1), "DTRSM", 6)
    expect(lb, lambda: f.dtrsm("L", "U", "N", "N", 2, 0, 1.0, A, 1, B, 2), "DTRSM", 9)
    expect(lb, lambda: f.dtrsm("L", "U", "N", "N", 2, 0, 1.0, A, 2, B, 1), "DTRSM", 11)
    expect(lb, lambda: f.dtrmm("R", "U", "N", "N", 0, 2, 1.0, A, 1, B, 1), "DTRMM", 9)
    expect(lb, lambda: f.dsyrk("/", "N", 0, 0, 1.0, A, 1, 0.0, W, 1), "DSYRK", 1)
    expect(lb, lambda: f.dsyrk("U", "/", 0, 0, 1.0, A, 1, 0.0, W, 1), "DSYRK", 2)
    expect(lb, lambda: f.dsyrk("U", "N", -1, 0, 1.0, A, 1, 0.0, W, 1), "DSYRK", 3)
    expect(lb, lambda: f.dsyrk("U", "N", 0, -1, 1.0, A, 1, 0.0, W, 1), "DSYRK", 4)
    expect(lb, lambda: f.dsyrk("U", "N", 2, 0, 1.0, A, 1, 0.0, W, 2), "DSYRK", 7)
    expect(lb, lambda: f.dsyrk("U", "N", 2, 0, 1.0, A, 2, 0.0, W, 1), "DSYRK", 10)


def test_lapacke_layer_errors(lb):
    L = lb.lib()
    vp = lambda x: x.ctypes.data_as(C.c_void_p)
    assert L.LAPACKE_dgetrf(100, 2, 2, vp(A), 2, vp(IP)) == -1                      # bad layout
    assert L.LAPACKE_dgetrf_work(101, 2, 3, vp(A), 2, vp(IP)) == -5                 # row-major lda < n
    assert L.LAPACKE_dgetrf_work(102, 2, 2, vp(A), 1, vp(IP)) == -5                 # Fortran -4 shifted by one
    assert L.LAPACKE_dpotrf_work(101, C.c_char(b"L"), 3, vp(A), 2) == -5
    assert L.LAPACKE_dgetrs_work(101, C.c_char(b"N"), 2, 3, vp(A), 2, vp(IP), vp(B), 2) == -9
    assert L.LAPACKE_dgesv_work(101, 2, 3, vp(A), 1, vp(IP), vp(B), 3) == -5
    assert L.LAPACKE_dgeqrf_work(101, 2, 3, vp(A), 2, vp(TAU), vp(W), 16) == -5
    assert L.LAPACKE_dlaswp_work(101, 3, vp(A), 2, 1, 1, vp(np.ones(1, dtype=np.int32)), 1) == -4
    wq = np.zeros(1)
    assert L.LAPACKE_dgeqrf_work(102, 300, 200, vp(A), 300, vp(TAU), vp(wq), -1) == 0 and wq[0] == 6400
    nan = np.full((2, 2), np.nan, order="F")
    L.LAPACKE_set_nancheck(1)
    assert L.LAPACKE_dgetrf(102, 2, 2, vp(nan), 2, vp(IP)) == -4                    # NaN pre-check (lapacke_dgetrf.c:42-49)
    assert L.LAPACKE_dpotrf(102, C.c_char(b"L"), 2, vp(nan), 2) == -4
    assert L.LAPACKE_dgesv(102, 2, 1, vp(A), 2, vp(IP), vp(nan), 2) == -7


def test_cblas_illegal_enums_need_no_gpu(lb, capfd):
    """CBLAS/src/cblas_dgemm.c:52-76, cblas_dtrsm.c:46-84: illegal option values are reported and nothing is computed."""
    L = lb.lib()
    dp = C.POINTER(C.c_double)
    L.cblas_dgemm.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, dp, C.c_int, dp, C.c_int,
                              C.c_double, dp, C.c_int]
    L.cblas_dtrsm.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, dp, C.c_int, dp, C.c_int]
    L.cblas_dgemm.restype = None
    L.cblas_dtrsm.restype = None
    c = np.full((2, 2), 5.0)
    p = lambda x: x.ctypes.data_as(dp)
    L.cblas_dgemm(102, 999, 111, 2, 2, 2, 1.0, p(A), 2, p(B), 2, 0.0, p(c), 2)
    L.cblas_dgemm(100, 111, 111, 2, 2, 2, 1.0, p(A), 2, p(B), 2, 0.0, p(c), 2)
    L.cblas_dtrsm(101, 141, 121, 111, 555, 2, 2, 1.0, p(A), 2, p(c), 2)
    err = capfd.readouterr().err
    assert "Parameter 2 to routine cblas_dgemm" in err and "Parameter 1 to routine cblas_dgemm" in err
    assert "Parameter 5 to routine cblas_dtrsm" in err
    assert np.all(c == 5.0)


def test_ilp64_entry_points_narrow_and_report(lb):
    """include/lapack_b200_f77_64.h: the _64 symbols forward to the 32-bit-index routines; the reference's own argument
    errors come back widened, and a dimension that does not fit in 32 bits is reported at its argument position."""
    L = lb.lib()
    i64 = C.c_int64
    a = np.zeros((4, 4), order="F")
    ip = np.zeros(4, dtype=np.int64)
    info = i64(7)

    def getrf64(m, n, lda):
        L.dgetrf_64_(C.byref(i64(m)), C.byref(i64(n)), a.ctypes.data_as(C.c_void_p), C.byref(i64(lda)),
                     ip.ctypes.data_as(C.c_void_p), C.byref(info))
        return info.value

    assert expect(lb, lambda: getrf64(-1, 0, 1), "DGETRF", 1) == -1
    assert expect(lb, lambda: getrf64(2, 1, 1), "DGETRF", 4) == -4
    assert expect(lb, lambda: getrf64(1 << 40, 1, 1 << 40), "DGETRF", 1) == -1          # does not fit in 32 bits
    assert getrf64(0, 0, 1) == 0                                                        # quick return, no GPU needed
    L.lb200_clear_xerbla()
    one = C.c_double(1.0)
    L.dgemm_64_(b"N", b"N", C.byref(i64(1 << 33)), C.byref(i64(1)), C.byref(i64(1)), C.byref(one), a.ctypes.data_as(C.c_void_p),
                C.byref(i64(1 << 33)), a.ctypes.data_as(C.c_void_p), C.byref(i64(1)), C.byref(one), a.ctypes.data_as(C.c_void_p),
                C.byref(i64(1 << 33)), C.c_size_t(1), C.c_size_t(1))
    name, pos, cnt = _last(lb)
    assert (name, pos, cnt) == ("DGEMM", 3, 1)
