"""GPU parity: the persistent streaming triangular solve (csrc/trsv_stream.cu) behind DTRSM / DGETRS / DPOTRS with
few right-hand sides, against the oracle's DTRSM (BLAS/SRC/dtrsm.f:278-327), and BASELINE configs[0] at its own size
(DGESV n=4096, nrhs=1; SRC/dgesv.f:165-172): IPIV bit-exact, solution within 1e-10 of the oracle's."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from oracle import oracle as O  # noqa: E402

SEED = (1988, 1989, 1990, 1991)


@pytest.fixture(scope="module")
def lb():
    import lapack_b200
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    lapack_b200.lib().lb200_set_xerbla_mode(2)
    return lapack_b200


def rel(x, y):
    return float(np.max(np.abs(x - y))) / max(1e-300, float(np.max(np.abs(y))))


@pytest.mark.parametrize("uplo", "LU")
@pytest.mark.parametrize("trans", "NT")
@pytest.mark.parametrize("diag", "NU")
def test_fewrhs_stream_vs_oracle(lb, uplo, trans, diag):
    rng = np.random.default_rng(11)
    for (m, nrhs, pad) in ((129, 1, 0), (200, 2, 3), (256, 1, 0), (257, 3, 1), (1000, 1, 0), (1333, 5, 7), (2049, 8, 0)):
        lda = m + pad
        abuf = np.full((lda, m), 7.5e300, order="F")                   # padding rows / the other triangle must never be read
        t = rng.uniform(-1, 1, (m, m)) / np.sqrt(m)
        t += np.eye(m) * (1.0 if diag == "N" else 0.0) * 2.0
        if uplo == "L":
            abuf[:m][np.tril_indices(m)] = t[np.tril_indices(m)]
        else:
            abuf[:m][np.triu_indices(m)] = t[np.triu_indices(m)]
        if diag == "U":
            abuf[:m][np.diag_indices(m)] = 9.9e300                     # a unit diagonal is not referenced (dtrsm.f:283)
        aref = np.asfortranarray(np.where(np.abs(abuf[:m]) > 1e300, 0.0, abuf[:m]))
        if diag == "U":
            aref[np.diag_indices(m)] = 1.0
        ldb = m + 2
        bbuf = np.full((ldb, nrhs), -3.0e10, order="F")
        bbuf[:m] = rng.uniform(-1, 1, (m, nrhs))
        want = np.asfortranarray(bbuf[:m].copy())
        O.dtrsm("L", uplo, trans, diag, m, nrhs, 1.0, aref, want)
        for mode in (1, 0):                                            # streaming kernel, then the leaf/GEMV recursion
            lb.lib().lb200_set_fewrhs_mode(mode)
            got = bbuf.copy(order="F")
            lb.f77.dtrsm("L", uplo, trans, diag, m, nrhs, 1.0, abuf, lda, got, ldb)
            lb.lib().lb200_set_fewrhs_mode(1)
            assert np.all(got[m:] == -3.0e10)
            assert rel(got[:m], want) < 1e-10, (mode, m, nrhs, uplo, trans, diag)


def test_fewrhs_stream_zero_diagonal_and_tiny(lb):
    """Division semantics of dtrsm.f:282: an exactly zero diagonal gives Inf/NaN, a denormal diagonal is divided by."""
    m = 300
    rng = np.random.default_rng(5)
    a = np.asfortranarray(np.tril(rng.uniform(-1, 1, (m, m)) / m) + 2 * np.eye(m))
    a[250, 250] = 0.0
    b = np.asfortranarray(rng.uniform(-1, 1, (m, 1)))
    want = b.copy(order="F")
    with np.errstate(all="ignore"):
        O.dtrsm("L", "L", "N", "N", m, 1, 1.0, a, want)
    got = b.copy(order="F")
    lb.f77.dtrsm("L", "L", "N", "N", m, 1, 1.0, a, m, got, m)
    assert np.allclose(got[:250], want[:250], rtol=1e-10, atol=0)
    assert not np.isfinite(got[250, 0]) and not np.isfinite(want[250, 0])
    a[250, 250] = 1e-310                                               # denormal pivot: divide, do not multiply by Inf
    want = b.copy(order="F")
    with np.errstate(all="ignore"):
        O.dtrsm("L", "L", "N", "N", m, 1, 1.0, a, want)
    got = b.copy(order="F")
    lb.f77.dtrsm("L", "L", "N", "N", m, 1, 1.0, a, m, got, m)
    fin = np.isfinite(want[:, 0])
    assert np.array_equal(np.isfinite(got[:, 0]), fin)
    assert np.allclose(got[fin], want[fin], rtol=1e-9, atol=0)


def test_c1_dgesv_n4096_vs_oracle(lb):
    """BASELINE configs[0] at its own size: DGESV n=4096, 1 RHS, DLARNV(2) input -- IPIV bit-exact with the oracle."""
    n = 4096
    a, seed = O.random_matrix(n, n, SEED)
    xact, _ = O.random_matrix(n, 1, seed)
    b = np.asfortranarray(a @ xact)
    lu_ref, x_ref = a.copy(order="F"), b.copy(order="F")
    ipiv_ref, info_ref = O.dgesv(lu_ref, x_ref)
    lu, x = a.copy(order="F"), b.copy(order="F")
    ipiv, info = lb.f77.gesv(lu, x)
    assert info == info_ref == 0
    assert np.array_equal(ipiv, ipiv_ref)
    assert rel(lu, lu_ref) < 1e-11                                     # the factors themselves agree to ~1e-12
    assert O.dget02("N", a, x, b) < O.THRESH
    # Agreement of the SOLUTIONS is limited by conditioning, not by the code: at this size the oracle's own forward error
    # against XACT is ~5e-11 (cond_1(A) ~ 1e6), and two backward-stable solvers (different summation order) differ by a
    # small multiple of eps*cond.  So the bar is the reference's own DGET04 criterion (TESTING/LIN/dget04.f:155-171):
    # |x - x_ref| / (|x_ref| * cond * eps) below the suite threshold -- plus the flat 1e-10 whenever conditioning allows it.
    cond1 = np.linalg.cond(a, 1)
    eps = 2.0 ** -53
    err_ref = rel(x_ref, xact)
    assert rel(x, xact) / (cond1 * eps) < O.THRESH
    assert rel(x, x_ref) / (cond1 * eps) < O.THRESH
    assert rel(x, x_ref) < max(1e-10, 10.0 * err_ref), (rel(x, x_ref), err_ref, cond1)
    xt_ref, xt = b.copy(order="F"), b.copy(order="F")
    O.dgetrs("T", lu_ref, ipiv_ref, xt_ref)
    assert lb.f77.getrs("T", lu, ipiv, xt) == 0
    assert rel(xt, xt_ref) / (cond1 * eps) < O.THRESH
    assert O.dget02("T", a, xt, b) < O.THRESH


def test_dposv_n4096_vs_oracle(lb):
    n = 3000
    for uplo in "LU":
        s, seed = O.spd_matrix(n, SEED)
        xact, _ = O.random_matrix(n, 2, seed)
        b = np.asfortranarray(s @ xact)
        f_ref, x_ref = s.copy(order="F"), b.copy(order="F")
        assert O.dposv(uplo, f_ref, x_ref) == 0
        f, x = s.copy(order="F"), b.copy(order="F")
        assert lb.f77.posv(uplo, f, x) == 0
        assert rel(x, x_ref) < 1e-10
        assert O.dpot02(uplo, s, x, b) < O.THRESH


def test_dlaswp_long_pivot_lists(lb):
    """More interchanges than one pivot chunk (2048) holds, forward and reverse, INCX = +-1 and +-2, few and many columns
    (SRC/dlaswp.f:138-150: with INCX < 0 the pivots are read from IPIV(K1 + (K2-K1)*|INCX|) downwards)."""
    rng = np.random.default_rng(21)
    m = 5000
    for n in (1, 3, 40):
        a = np.asfortranarray(rng.uniform(-1, 1, (m, n)))
        for (k1, k2, ainc) in ((1, 4500, 1), (7, 4700, 1), (1, 2300, 2)):
            ipiv = np.zeros(k1 + (k2 - k1) * ainc + 4, dtype=np.int32)
            for i in range(k1, k2 + 1):
                ipiv[k1 + (i - k1) * ainc - 1] = rng.integers(i, m + 1)
            for sgn in (1, -1):
                want = a.copy(order="F")
                O.dlaswp(want, k1, k2, ipiv, sgn * ainc)
                got = a.copy(order="F")
                lb.f77.dlaswp(n, got, m, k1, k2, ipiv, sgn * ainc)
                assert np.array_equal(got, want), (n, k1, k2, sgn * ainc)


def test_dlaswp_general_pivots_many_columns(lb):
    """Composed-permutation path: pivots anywhere in the matrix (also above K1 and repeated), many columns, both directions."""
    rng = np.random.default_rng(22)
    m, n = 3000, 333
    a = np.asfortranarray(rng.uniform(-1, 1, (m, n)))
    for (k1, k2) in ((1, 512), (100, 700), (2000, 2999), (5, 8)):
        for kind in ("lu", "any", "few"):
            ipiv = np.zeros(k2 + 2, dtype=np.int32)
            for i in range(k1, k2 + 1):
                if kind == "lu":
                    ipiv[i - 1] = rng.integers(i, m + 1)
                elif kind == "any":
                    ipiv[i - 1] = rng.integers(1, m + 1)
                else:
                    ipiv[i - 1] = rng.choice([i, 7, 2999, 1500])
            for incx in (1, -1):
                want = a.copy(order="F")
                O.dlaswp(want, k1, k2, ipiv, incx)
                got = a.copy(order="F")
                lb.f77.dlaswp(n, got, m, k1, k2, ipiv, incx)
                assert np.array_equal(got, want), (k1, k2, kind, incx)
