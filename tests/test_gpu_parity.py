"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on identical seeded inputs.

Bars (BASELINE.json north_star): IPIV bit-exact on random well-conditioned matrices; LAPACK TESTING/LIN residual
ratios < 30 (TESTING/dtest.in:13); solutions within 1e-10 relative of the oracle's; BLAS results within the
BLAS/TESTING/dblat3.in threshold of 16 (ratio |err| / (eps * sum|a||b|)).
"""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from oracle import oracle as O  # noqa: E402

SEED = (1988, 1989, 1990, 1991)
EPS = 2.0 ** -53


@pytest.fixture(scope="module")
def lb():
    import lapack_b200
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    lapack_b200.lib().lb200_set_xerbla_mode(2)
    return lapack_b200


def rel(x, y):
    s = max(1e-300, float(np.max(np.abs(y)))) if y.size else 1.0
    return float(np.max(np.abs(x - y))) / s if x.size else 0.0


def dev_to_np(t):
    return np.asfortranarray(t.cpu().numpy())


def np_to_dev(lb, a):
    t = lb.dev.colmajor(a.shape[0], a.shape[1])
    t.copy_(torch.from_numpy(np.ascontiguousarray(a)))
    return t


# ------------------------------------------------------------------------------------------- generator
def test_larnv_bit_exact(lb):
    for (m, n, off) in ((1000, 37, 0), (129, 5, 12345), (1, 1, 7), (64, 64, 2 ** 33 + 5)):
        a = lb.dev.larnv_matrix(m, n, SEED, off)
        x, _ = O.dlarnv(2, SEED, off + m * n) if off < 10 ** 6 else (None, None)
        got = dev_to_np(a)
        if x is not None:
            want = np.asfortranarray(x[off:].reshape((n, m)).T)
            assert np.array_equal(got, want)
        else:
            assert np.all(np.abs(got) <= 1.0) and len(np.unique(got)) == got.size


def test_spd_generator(lb):
    n = 300
    a = lb.dev.larnv_matrix(n, n, SEED)
    lb.dev.make_spd(a, float(n))
    want, _ = O.spd_matrix(n, SEED)
    assert np.array_equal(dev_to_np(a), want)


# ------------------------------------------------------------------------------------------- BLAS 3
@pytest.mark.parametrize("ta", "NT")
@pytest.mark.parametrize("tb", "NT")
def test_dgemm_vs_oracle(lb, ta, tb):
    rng = np.random.default_rng(5)
    for (m, n, k, alpha, beta) in ((9, 5, 3, 0.7, 1.3), (130, 67, 45, -1.0, 1.0), (64, 64, 200, 1.0, 0.0), (5, 9, 0, 1.0, 0.7),
                                   (33, 1, 17, 0.0, 1.3), (257, 129, 64, -1.0, 1.0)):
        a = np.asfortranarray(rng.uniform(-1, 1, (m, k) if ta == "N" else (k, m)))
        b = np.asfortranarray(rng.uniform(-1, 1, (k, n) if tb == "N" else (n, k)))
        c = np.asfortranarray(rng.uniform(-1, 1, (m, n)))
        want = c.copy(order="F")
        O.dgemm(ta, tb, m, n, k, alpha, a if a.size else np.zeros((1, 1), order="F"), b if b.size else np.zeros((1, 1), order="F"), beta, want)
        got = c.copy(order="F")
        lda = max(1, a.shape[0])
        ldb = max(1, b.shape[0])
        lb.f77.dgemm(ta, tb, m, n, k, alpha, a if a.size else np.zeros(1), lda, b if b.size else np.zeros(1), ldb, beta, got, m)
        opa = np.abs(a if ta == "N" else a.T)
        opb = np.abs(b if tb == "N" else b.T)
        g = abs(alpha) * (opa @ opb if k else 0.0) + abs(beta) * np.abs(c) + 1e-300
        assert np.max(np.abs(got - want) / g) / EPS < 16.0


def test_dsyrk_triangle_only(lb):
    rng = np.random.default_rng(6)
    for uplo in "LU":
        for tr in "NT":
            n, k = 77, 40
            a = np.asfortranarray(rng.uniform(-1, 1, (n, k) if tr == "N" else (k, n)))
            c = np.asfortranarray(rng.uniform(-1, 1, (n + 1, n)))      # ld = n+1 like dblat3
            want = c.copy(order="F")
            O.dsyrk(uplo, tr, n, k, -1.0, a, 1.0, want[:n, :])
            got = c.copy(order="F")
            lb.f77.dsyrk(uplo, tr, n, k, -1.0, a, a.shape[0], 1.0, got, n + 1)
            assert rel(got[:n], want[:n]) < 1e-13
            other = np.triu(np.ones((n, n), bool), 1) if uplo == "L" else np.tril(np.ones((n, n), bool), -1)
            assert np.array_equal(got[:n][other], c[:n][other])          # other triangle untouched
            assert np.array_equal(got[n], c[n])                          # padding row untouched


@pytest.mark.parametrize("side", "LR")
@pytest.mark.parametrize("uplo", "LU")
@pytest.mark.parametrize("trans", "NT")
@pytest.mark.parametrize("diag", "NU")
def test_dtrsm_dtrmm_vs_oracle(lb, side, uplo, trans, diag):
    rng = np.random.default_rng(7)
    for (m, n) in ((1, 1), (5, 9), (33, 70), (130, 45), (70, 200)):
        na = m if side == "L" else n
        a = np.asfortranarray(rng.uniform(-1, 1, (na, na)) + na * np.eye(na) * 0.5)
        b = np.asfortranarray(rng.uniform(-1, 1, (m, n)))
        for alpha in (1.0, 0.7):
            want = b.copy(order="F")
            O.dtrsm(side, uplo, trans, diag, m, n, alpha, a, want)
            got = b.copy(order="F")
            lb.f77.dtrsm(side, uplo, trans, diag, m, n, alpha, a, na, got, m)
            assert rel(got, want) < 1e-11, ("trsm", m, n, alpha)
            want = b.copy(order="F")
            O.dtrmm(side, uplo, trans, diag, m, n, alpha, a, want)
            got = b.copy(order="F")
            lb.f77.dtrmm(side, uplo, trans, diag, m, n, alpha, a, na, got, m)
            assert rel(got, want) < 1e-13, ("trmm", m, n, alpha)


def test_dlaswp_vs_oracle(lb):
    rng = np.random.default_rng(8)
    m, n = 60, 45
    a = np.asfortranarray(rng.uniform(-1, 1, (m, n)))
    ipiv = np.array([rng.integers(i + 1, m + 1) for i in range(20)], dtype=np.int32)
    for incx in (1, -1):
        for (k1, k2) in ((1, 20), (3, 11), (7, 7)):
            want = a.copy(order="F")
            O.dlaswp(want, k1, k2, ipiv, incx)
            got = a.copy(order="F")
            lb.f77.dlaswp(n, got, m, k1, k2, ipiv, incx)
            assert np.array_equal(got, want), (incx, k1, k2)


# ------------------------------------------------------------------------------------------- LU
LU_SHAPES = [(1, 1), (1, 5), (9, 1), (5, 3), (3, 5), (16, 16), (17, 17), (50, 50), (100, 37), (37, 100), (300, 300),
             (1100, 260), (1500, 1500), (2100, 700)]


@pytest.mark.parametrize("shape", LU_SHAPES)
@pytest.mark.parametrize("recursive", [False, True])
def test_dgetrf_ipiv_and_residual(lb, shape, recursive):
    m, n = shape
    a, _ = O.random_matrix(m, n, SEED)
    ref = a.copy(order="F")
    ipiv_ref, info_ref = O.dgetrf(ref) if not recursive else O.dgetrf2(ref)
    got = a.copy(order="F")
    ipiv, info = lb.f77.getrf(got, recursive)
    assert info == info_ref == 0
    assert np.array_equal(ipiv, ipiv_ref)                         # IPIV bit-exact
    assert rel(got, ref) < 1e-10
    assert O.dget01(a, got, ipiv) < O.THRESH


@pytest.mark.parametrize("nb,la", [(16, 0), (16, 1), (32, 1), (64, 1), (128, 0), (128, 1)])
def test_dgetrf_blocked_lookahead(lb, nb, la):
    """Force the blocked right-looking driver (with and without look-ahead streams) at test sizes."""
    L = lb.lib()
    L.lb200_set_getrf_params(nb, 0, la)
    try:
        for (m, n) in ((300, 300), (700, 450), (450, 700), (1300, 1300)):
            a, _ = O.random_matrix(m, n, SEED)
            ref = a.copy(order="F")
            ipiv_ref, _ = O.dgetrf(ref)
            got = a.copy(order="F")
            ipiv, info = lb.f77.getrf(got)
            assert info == 0
            assert np.array_equal(ipiv, ipiv_ref), (m, n)
            assert O.dget01(a, got, ipiv) < O.THRESH
            assert rel(got, ref) < 1e-10
    finally:
        L.lb200_set_getrf_params(512, 0, 1)


@pytest.mark.parametrize("cluster_max,big_leaf", [(0, 0), (0, 1), (2, 0), (2, 1), (8, 1), (16, 1)])
def test_dgetrf_leaf_kernels_agree(lb, cluster_max, big_leaf):
    """The global-exchange leaf (cluster_max=0) and the thread-block-cluster leaf give the same pivots and factors
    as the reference, including panels that span several CTAs, ties, zero columns and NaN entries."""
    L = lb.lib()
    L.lb200_set_getrf_cluster_max(cluster_max)
    L.lb200_set_getrf_big_leaf(big_leaf)
    try:
        for (m, n) in ((3000, 40), (2500, 300), (1025, 1025), (5000, 17)):
            a, _ = O.random_matrix(m, n, SEED)
            a[:, 3] = np.round(a[:, 3] * 4.0) / 4.0                # many exact ties in one column
            a[m // 2:, 5] = a[m // 2, 5]                           # a long run of equal entries
            a[:, 7] = 0.0                                          # exactly singular column -> INFO = 8 (if reached)
            ref = a.copy(order="F")
            ipiv_ref, info_ref = O.dgetrf2(ref)
            got = a.copy(order="F")
            ipiv, info = lb.f77.getrf(got, True)
            assert info == info_ref, (m, n)
            assert np.array_equal(ipiv, ipiv_ref), (m, n)
            assert rel(got, ref) < 1e-10
        # NaN handling of IDAMAX (idamax.f:103): a NaN is never selected unless it sits in the first place
        a, _ = O.random_matrix(2100, 20, SEED)
        a[1500, 2] = np.nan
        a[4, 4] = np.nan
        ref = a.copy(order="F")
        ipiv_ref, info_ref = O.dgetrf2(ref)
        got = a.copy(order="F")
        ipiv, info = lb.f77.getrf(got, True)
        assert info == info_ref
        assert np.array_equal(ipiv, ipiv_ref)
        assert np.array_equal(np.isnan(got), np.isnan(ref))
    finally:
        L.lb200_set_getrf_cluster_max(16)
        L.lb200_set_getrf_big_leaf(1)


def test_dgetrf_pinned_host_streamed(lb):
    """Pinned host caller, square n >= 8192: split upload / top-level recursion (fortran_abi.cu getrf_host_streamed)."""
    n, lda = 8200, 8208
    a, _ = O.random_matrix(n, n, SEED)
    buf = torch.empty((n, lda), dtype=torch.float64).pin_memory()
    h = buf.numpy().T                                               # (lda, n) column-major view
    h[:] = 7.0
    h[:n, :] = a
    ipiv = np.zeros(n, dtype=np.int32)
    assert lb.f77.dgetrf(n, n, h, lda, ipiv) == 0
    got = np.asfortranarray(h[:n, :])
    assert np.all(h[n:, :] == 7.0)
    # same pivots and factors as the device-resident blocked driver
    ad = np_to_dev(lb, a)
    piv_d, info_d = lb.dev.getrf(ad)
    assert np.array_equal(ipiv, piv_d.cpu().numpy())
    assert rel(got, dev_to_np(ad)) < 1e-10
    # residual by a GPU product (the oracle's O(n^3) checker is too slow at this size)
    L = torch.tril(ad, -1) + torch.eye(n, dtype=torch.float64, device=ad.device)
    U = torch.triu(ad)
    pa = torch.from_numpy(a).to(ad.device)
    for i in range(n):                                              # apply the interchanges to A
        p = int(ipiv[i]) - 1
        if p != i:
            tmp = pa[i].clone(); pa[i] = pa[p]; pa[p] = tmp
    resid = float((L @ U - pa).abs().sum(dim=0).max()) / (n * float(pa.abs().sum(dim=0).max()) * EPS)
    assert resid < O.THRESH
    # exactly singular column in the right part -> INFO from the second half, shifted (dgetrf2.f:251-252)
    h[:n, :] = a
    h[:n, 5000] = 0.0
    assert lb.f77.dgetrf(n, n, h, lda, ipiv) == 5001


@pytest.mark.parametrize("m,n", [(200000, 24), (400000, 20)])
def test_very_tall_panels(lb, m, n):
    """Panels taller than (#SMs x 1024) rows: the 8- and 16-rows-per-thread leaf kernels (LU and QR)."""
    a, _ = O.random_matrix(m, n, SEED)
    ref = a.copy(order="F")
    ipiv_ref, info_ref = O.dgetrf2(ref)
    got = a.copy(order="F")
    ipiv, info = lb.f77.getrf(got)
    assert info == info_ref == 0
    assert np.array_equal(ipiv, ipiv_ref)
    assert rel(got, ref) < 1e-10
    qr_ref = a.copy(order="F")
    tau_ref, _, _ = O.dgeqrf(qr_ref)
    qr = a.copy(order="F")
    tau, info, _ = lb.f77.geqrf(qr)
    assert info == 0
    assert rel(tau, tau_ref) < 1e-10
    assert rel(np.triu(qr[:n, :]), np.triu(qr_ref[:n, :])) < 1e-10
    assert rel(qr, qr_ref) < 1e-9


def test_dgetrf_singular_info(lb):
    """TESTING/LIN/dchkge.f:328-347: zero a column -> INFO = that column, factorization completes."""
    n = 120
    a, _ = O.random_matrix(n, n, SEED)
    for izero in (1, n, n // 2 + 1):
        b = a.copy(order="F")
        b[:, izero - 1] = 0.0
        ref = b.copy(order="F")
        ipiv_ref, info_ref = O.dgetrf(ref)
        got = b.copy(order="F")
        ipiv, info = lb.f77.getrf(got)
        assert info == info_ref == izero
        assert np.array_equal(ipiv, ipiv_ref)
        assert O.dget01(b, got, ipiv) < O.THRESH
    b = a.copy(order="F")
    b[:, n // 2:] = 0.0                                            # type 7: last half zero
    ref = b.copy(order="F")
    ipiv_ref, info_ref = O.dgetrf(ref)
    got = b.copy(order="F")
    ipiv, info = lb.f77.getrf(got)
    assert info == info_ref == n // 2 + 1
    assert np.array_equal(ipiv, ipiv_ref)


def test_dgetrf_tiny_pivot_and_lda(lb):
    """|pivot| < SFMIN takes the divide branch (dgetrf2.f:207-209); lda > m leaves the padding untouched."""
    m, n, lda = 40, 30, 47
    a, _ = O.random_matrix(m, n, SEED)
    a[:, 0] *= 1e-310
    buf = np.full((lda, n), -1.0e10, order="F")
    buf[:m] = a
    ref = a.copy(order="F")
    ipiv_ref, _ = O.dgetrf(ref)
    ipiv = np.zeros(n, dtype=np.int32)
    info = lb.f77.dgetrf(m, n, buf, lda, ipiv)
    assert info == 0 and np.array_equal(ipiv, ipiv_ref)
    assert rel(buf[:m], ref) < 1e-10
    assert np.all(buf[m:] == -1.0e10)


def test_dgesv_dgetrs_solution(lb):
    for (n, nrhs) in ((1, 1), (33, 2), (200, 1), (777, 15), (1500, 3)):
        a, seed = O.random_matrix(n, n, SEED)
        xact, _ = O.random_matrix(n, nrhs, seed)
        b = np.asfortranarray(a @ xact)
        lu_ref, x_ref = a.copy(order="F"), b.copy(order="F")
        ipiv_ref, info_ref = O.dgesv(lu_ref, x_ref)
        lu, x = a.copy(order="F"), b.copy(order="F")
        ipiv, info = lb.f77.gesv(lu, x)
        assert info == info_ref == 0
        assert np.array_equal(ipiv, ipiv_ref)
        assert rel(x, x_ref) < 1e-10                               # within 1e-10 relative of the reference solution
        assert O.dget02("N", a, x, b) < O.THRESH
        xt_ref = b.copy(order="F")
        O.dgetrs("T", lu_ref, ipiv_ref, xt_ref)
        xt = b.copy(order="F")
        assert lb.f77.getrs("T", lu, ipiv, xt) == 0
        assert rel(xt, xt_ref) < 1e-9
        assert O.dget02("T", a, xt, b) < O.THRESH


# ------------------------------------------------------------------------------------------- Cholesky
@pytest.mark.parametrize("uplo", "LU")
@pytest.mark.parametrize("n", [1, 2, 31, 32, 33, 100, 513, 1400])
@pytest.mark.parametrize("recursive", [False, True])
def test_dpotrf_vs_oracle(lb, uplo, n, recursive):
    s, _ = O.spd_matrix(n, SEED)
    rogue = s.copy(order="F")
    mask = np.triu(np.ones((n, n), bool), 1) if uplo == "L" else np.tril(np.ones((n, n), bool), -1)
    rogue[mask] = -1.0e10                                          # the other triangle must not be referenced
    ref = rogue.copy(order="F")
    assert (O.dpotrf2(uplo, ref) if recursive else O.dpotrf(uplo, ref)) == 0
    got = rogue.copy(order="F")
    assert lb.f77.potrf(uplo, got, recursive) == 0
    assert np.all(got[mask] == -1.0e10)
    tri = np.tril if uplo == "L" else np.triu
    assert rel(tri(got), tri(ref)) < 1e-12
    assert O.dpot01(uplo, s, got) < O.THRESH


@pytest.mark.parametrize("nb,la", [(32, 0), (32, 1), (64, 1), (128, 1)])
def test_dpotrf_blocked_lookahead(lb, nb, la):
    L = lb.lib()
    L.lb200_set_potrf_params(nb, la)
    try:
        for uplo in "LU":
            for n in (200, 333, 1025):
                s, _ = O.spd_matrix(n, SEED)
                got = s.copy(order="F")
                assert lb.f77.potrf(uplo, got) == 0
                assert O.dpot01(uplo, s, got) < O.THRESH
                ref = s.copy(order="F")
                O.dpotrf(uplo, ref)
                tri = np.tril if uplo == "L" else np.triu
                assert rel(tri(got), tri(ref)) < 1e-12
    finally:
        L.lb200_set_potrf_params(512, 1)


@pytest.mark.parametrize("uplo", "LU")
def test_dpotrf_pinned_host_streamed(lb, uplo):
    """Pinned host caller, n >= 2048: triangle-only upload + streamed download of finished block columns."""
    n, lda = 2600, 2610
    s, _ = O.spd_matrix(n, SEED)
    buf = torch.empty((n, lda), dtype=torch.float64).pin_memory()   # row-major (n, lda) == column-major lda x n
    h = buf.numpy().T                                               # (lda, n) column-major view
    mask = np.triu(np.ones((n, n), bool), 1) if uplo == "L" else np.tril(np.ones((n, n), bool), -1)
    h[:] = 7.0
    h[:n, :] = s
    h[:n, :][mask] = -1.0e10
    assert lb.f77.dpotrf(uplo, n, h, lda) == 0
    got = np.asfortranarray(h[:n, :])
    assert np.all(got[mask] == -1.0e10)
    assert np.all(h[n:, :] == 7.0)
    assert O.dpot01(uplo, s, got) < O.THRESH
    ref = s.copy(order="F")
    O.dpotrf(uplo, ref)
    tri = np.tril if uplo == "L" else np.triu
    assert rel(tri(got), tri(ref)) < 1e-12
    # not positive definite: INFO and no hang
    h[:n, :] = s
    h[2100, 2100] = -5.0
    assert lb.f77.dpotrf(uplo, n, h, lda) == 2101
    # DPOSV through the same streamed path (pinned A, pageable B): solution vs the reference, B untouched on failure
    h[:n, :] = s
    h[:n, :][mask] = -1.0e10
    x_true, _ = O.random_matrix(n, 2, (7, 8, 9, 11))
    b = np.asfortranarray(s @ x_true)
    b_ref = b.copy(order="F")
    a_ref = s.copy(order="F")
    assert O.dposv(uplo, a_ref, b_ref) == 0
    assert lb.f77.dposv(uplo, n, 2, h, lda, b, n) == 0
    assert rel(b, b_ref) < 1e-10
    assert np.all(h[:n, :][mask] == -1.0e10)
    h[:n, :] = s
    h[5, 5] = 0.0
    b2 = b_ref.copy(order="F")
    assert lb.f77.dposv(uplo, n, 2, h, lda, b2, n) == 6
    assert np.array_equal(b2, b_ref)


@pytest.mark.parametrize("uplo", "LU")
def test_dpotrf_pinned_host_split_upload(lb, uplo):
    """Pinned host caller, n >= 8192: leading block factored while the trailing block is still uploading (one level of
    the DPOTRF2 recursion), streamed download; compared with the device-resident factorization."""
    n, lda = 8300, 8304
    dev = torch.device("cuda")
    s_dev = lb.dev.larnv_matrix(n, n, SEED)
    lb.dev.make_spd(s_dev, float(n))
    s = dev_to_np(s_dev)
    buf = torch.empty((n, lda), dtype=torch.float64).pin_memory()
    h = buf.numpy().T
    mask = np.triu(np.ones((n, n), bool), 1) if uplo == "L" else np.tril(np.ones((n, n), bool), -1)
    h[:] = 7.0
    h[:n, :] = s
    h[:n, :][mask] = -1.0e10
    assert lb.f77.dpotrf(uplo, n, h, lda) == 0
    got = np.asfortranarray(h[:n, :])
    assert np.all(got[mask] == -1.0e10) and np.all(h[n:, :] == 7.0)
    ref_dev = s_dev.clone()
    assert int(lb.dev.potrf(uplo, ref_dev).item()) == 0
    ref = dev_to_np(ref_dev)
    tri = np.tril if uplo == "L" else np.triu
    assert rel(tri(got), tri(ref)) < 1e-11
    f = torch.from_numpy(tri(got)).to(dev)
    prod = f @ f.T if uplo == "L" else f.T @ f
    resid = float((prod - s_dev).abs().sum(dim=0).max()) / (n * float(s_dev.abs().sum(dim=0).max()) * EPS)
    assert resid < O.THRESH
    # failure inside the trailing block: INFO is shifted by the size of the leading block (dpotrf2.f:228-231)
    h[:n, :] = s
    h[5000, 5000] = -3.0
    assert lb.f77.dpotrf(uplo, n, h, lda) == 5001
    h[:n, :] = s
    h[100, 100] = -3.0
    assert lb.f77.dpotrf(uplo, n, h, lda) == 101
    # DPOSV through the same path
    h[:n, :] = s
    x_true, _ = O.random_matrix(n, 2, (7, 8, 9, 11))
    b = np.asfortranarray(s @ x_true)
    assert lb.f77.dposv(uplo, n, 2, h, lda, b, n) == 0
    assert rel(b, x_true) < 1e-10


def test_dpotrf_not_positive_definite(lb):
    """TESTING/LIN/dchkpo.f:313-344: zero row+column IZERO -> INFO = IZERO."""
    n = 150
    s, _ = O.spd_matrix(n, SEED)
    for uplo in "LU":
        for izero in (1, n, n // 2 + 1):
            b = s.copy(order="F")
            b[izero - 1, :] = 0.0
            b[:, izero - 1] = 0.0
            ref = b.copy(order="F")
            info_ref = O.dpotrf(uplo, ref)
            got = b.copy(order="F")
            assert lb.f77.potrf(uplo, got) == info_ref == izero
    b = s.copy(order="F")
    b[40, 40] = np.nan
    assert lb.f77.potrf("L", b.copy(order="F")) == 41              # NaN on the diagonal (dpotrf2.f:169)


def test_dpotrf_abort_leaves_finite_matrix(lb):
    """dpotrf.f:219-220,239-240: the factorization stops at the first non-positive leading minor.  Everything queued behind the
    failing block (DTRSM / DSYRK / later panels, block-column downloads of the streamed host path) must not run on the unfactored
    block: no Inf/NaN may appear, and the block columns finished before the failure are the oracle's."""
    n, izero = 2600, 1301
    s, _ = O.spd_matrix(n, SEED)
    s[izero - 1, :] = 0.0
    s[:, izero - 1] = 0.0
    ref = s.copy(order="F")
    assert O.dpotrf("L", ref) == izero
    nb = 512
    done = ((izero - 1) // nb) * nb                                 # block columns completely factored before the failure
    for pinned in (False, True):
        if pinned:
            buf = torch.empty((n, n), dtype=torch.float64).pin_memory()
            got = buf.numpy().T                                     # column-major view of pinned memory
            got[:, :] = s
        else:
            got = s.copy(order="F")
        assert lb.f77.dpotrf("L", n, got if not pinned else buf.data_ptr(), n) == izero
        g = np.tril(got)
        assert np.all(np.isfinite(g)), pinned
        assert np.max(np.abs(g[:, :done] - np.tril(ref)[:, :done])) < 1e-10 * np.max(np.abs(ref))
        assert np.array_equal(np.triu(got, 1), np.triu(s, 1))       # the other triangle is never touched


def test_dposv_dpotrs_solution(lb):
    for uplo in "LU":
        for (n, nrhs) in ((1, 1), (150, 3), (900, 1)):
            s, seed = O.spd_matrix(n, SEED)
            xact, _ = O.random_matrix(n, nrhs, seed)
            b = np.asfortranarray(s @ xact)
            f_ref, x_ref = s.copy(order="F"), b.copy(order="F")
            assert O.dposv(uplo, f_ref, x_ref) == 0
            f, x = s.copy(order="F"), b.copy(order="F")
            assert lb.f77.posv(uplo, f, x) == 0
            assert rel(x, x_ref) < 1e-10
            assert O.dpot02(uplo, s, x, b) < O.THRESH
            x2 = b.copy(order="F")
            assert lb.f77.potrs(uplo, f, x2) == 0
            assert rel(x2, x_ref) < 1e-10


# ------------------------------------------------------------------------------------------- QR
QR_SHAPES = [(1, 1), (5, 3), (3, 5), (16, 16), (17, 40), (90, 40), (50, 50), (30, 45), (300, 200), (1100, 300), (2100, 2100)]


@pytest.mark.parametrize("shape", QR_SHAPES)
def test_dgeqrf_vs_oracle(lb, shape):
    m, n = shape
    a, _ = O.random_matrix(m, n, SEED)
    ref = a.copy(order="F")
    tau_ref, info_ref, _ = O.dgeqrf(ref)
    got = a.copy(order="F")
    tau, info, w1 = lb.f77.geqrf(got)
    assert info == 0
    assert rel(got, ref) < 1e-10
    assert rel(tau, tau_ref) < 1e-10
    if m <= 1200:
        res = O.dqrt01(a, got, tau)
        assert res[0] < O.THRESH and res[1] < O.THRESH


def test_dgeqrf_workspace_protocol(lb):
    """LWORK=-1 query returns N*NB (dgeqrf.f:197-204); LWORK < N is argument 7."""
    m, n = 200, 150
    a, _ = O.random_matrix(m, n, SEED)
    tau = np.zeros(n)
    wq = np.zeros(1)
    assert lb.f77.dgeqrf(m, n, a, m, tau, wq, -1) == 0 and wq[0] == n * 32
    assert lb.f77.dgeqrf(m, n, a, m, tau, np.zeros(10), 10) == -7


def test_dgeqrf_scaled_inputs(lb):
    """Near-overflow / near-underflow scalings (dlatb4.f QR types 7-8) must not break the fused norm reduction."""
    m, n = 120, 80
    a, _ = O.random_matrix(m, n, SEED)
    for scale in (1e292, 1e-292):
        b = np.asfortranarray(a * scale)
        got = b.copy(order="F")
        tau, info, _ = lb.f77.geqrf(got)
        assert info == 0 and np.all(np.isfinite(got))
        res = O.dqrt01(b, got, tau)
        assert res[0] < O.THRESH and res[1] < O.THRESH


def test_dgeqrf_graded_columns(lb):
    """Columns of wildly different magnitude inside one matrix: a column of entries ~1e-200 next to O(1) columns must
    keep its relative accuracy -- the reference gets it from DNRM2's scaled sum and DLARFG's rescaling loop (dlarfg.f:159-176)."""
    m, n = 150, 90
    a, _ = O.random_matrix(m, n, SEED)
    d = 10.0 ** np.where(np.arange(n) % 3 == 0, -200.0, np.where(np.arange(n) % 3 == 1, 0.0, 180.0))
    b = np.asfortranarray(a * d[None, :])
    ref = b.copy(order="F")
    tau_ref, _, _ = O.dgeqrf(ref)
    got = b.copy(order="F")
    tau, info, _ = lb.f77.geqrf(got)
    assert info == 0 and np.all(np.isfinite(got))
    assert np.max(np.abs(tau - tau_ref)) < 1e-12
    v_got, v_ref = np.tril(got, -1), np.tril(ref, -1)
    assert np.max(np.abs(v_got - v_ref)) < 1e-11                              # reflectors are scale-free
    r_got, r_ref = np.triu(got[:n]), np.triu(ref[:n])
    colmax = np.max(np.abs(r_ref), axis=0)
    assert np.all(colmax > 0)
    assert np.max(np.abs(r_got - r_ref) / colmax[None, :]) < 1e-11           # every column of R to ITS OWN scale


def test_dlarft_dlarfb_vs_oracle(lb):
    for (m, k, n) in ((90, 40, 17), (300, 100, 64), (64, 64, 10)):
        a, seed = O.random_matrix(m, k, SEED)
        tau, _ = O.dgeqr2(a)
        t_ref = O.dlarft(a, tau)
        t = lb.f77.larft(a, tau)
        assert rel(np.triu(t), np.triu(t_ref)) < 1e-11
        c0, _ = O.random_matrix(m, n, seed)
        for trans in "TN":
            want = c0.copy(order="F")
            O.dlarfb("L", trans, a, t_ref, want)
            got = c0.copy(order="F")
            lb.f77.larfb("L", trans, a, t_ref, got)
            assert rel(got, want) < 1e-11
        c1 = np.asfortranarray(c0.T.copy())
        for trans in "TN":
            want = c1.copy(order="F")
            O.dlarfb("R", trans, a, t_ref, want)
            got = c1.copy(order="F")
            lb.f77.larfb("R", trans, a, t_ref, got)
            assert rel(got, want) < 1e-11


# ------------------------------------------------------------------------------------------- DGELQF / DORMLQ / DGELS
@pytest.mark.parametrize("shape", [(40, 40), (170, 300), (300, 170), (520, 1500)])
def test_dgelqf_dormlq_vs_oracle(lb, shape):
    m, n = shape
    a, _ = O.random_matrix(m, n, SEED)
    ref = a.copy(order="F")
    tau_ref, info = O.dgelq2(ref)
    got = a.copy(order="F")
    tau, info = lb.f77.gelqf(got)
    assert info == 0
    assert rel(tau, tau_ref) < 1e-11 and rel(got, ref) < 1e-11
    nc = 7
    k = min(m, n)
    for side in "LR":
        c0, _ = O.random_matrix(n if side == "L" else nc, nc if side == "L" else n, (3, 5, 7, 9))
        for trans in "NT":
            c = c0.copy(order="F")
            assert lb.f77.ormlq(side, trans, got[:k, :], tau, c) == 0
            c_ref = c0.copy(order="F")
            assert O.dorml2(side, trans, np.asfortranarray(ref[:k, :]), tau_ref, c_ref) == 0
            assert rel(c, c_ref) < 1e-11, (side, trans)


@pytest.mark.parametrize("m,n,nrhs", [(300, 170, 3), (170, 300, 2), (200, 200, 1), (1500, 520, 4), (520, 1500, 4)])
@pytest.mark.parametrize("trans", "NT")
def test_dgels_vs_oracle(lb, m, n, nrhs, trans):
    a, _ = O.random_matrix(m, n, SEED)
    b, _ = O.random_matrix(max(m, n), nrhs, (3, 5, 7, 9))
    a_ref, x_ref = a.copy(order="F"), b.copy(order="F")
    assert O.dgels(trans, a_ref, x_ref) == 0
    a_got, x = a.copy(order="F"), b.copy(order="F")
    assert lb.f77.gels(trans, a_got, x) == 0
    rows = n if trans == "N" else m
    assert rel(x[:rows], x_ref[:rows]) < 1e-10
    assert rel(a_got, a_ref) < 1e-10                                  # the QR / LQ factors are returned in A
    # least-squares optimality / consistency, independent of the oracle
    op = a if trans == "N" else a.T
    rhs = b[:op.shape[0]]
    r = op @ x[:rows] - rhs
    if op.shape[0] >= op.shape[1]:
        assert np.max(np.abs(op.T @ r)) < 1e-9 * np.max(np.abs(op)) * np.max(np.abs(rhs)) * max(m, n)    # normal equations
    else:
        assert np.max(np.abs(r)) < 1e-10 * np.max(np.abs(rhs)) * max(m, n)                               # exact solve


def test_dgels_scaling_and_rank_deficiency(lb):
    m, n, nrhs = 200, 90, 2
    a, _ = O.random_matrix(m, n, SEED)
    b, _ = O.random_matrix(m, nrhs, (3, 5, 7, 9))
    for sa, sb in ((1e-300, 1.0), (1e300, 1.0), (1.0, 1e-300), (1.0, 1e300)):       # dgels.f:309-353
        a_ref, x_ref = (a * sa).copy(order="F"), (b * sb).copy(order="F")
        assert O.dgels("N", a_ref, x_ref) == 0
        a_got, x = (a * sa).copy(order="F"), (b * sb).copy(order="F")
        assert lb.f77.gels("N", a_got, x) == 0
        assert rel(x[:n], x_ref[:n]) < 1e-10, (sa, sb)
    a0 = a.copy(order="F")
    a0[:, 6] = 0.0                                                                  # R(7,7) = 0 exactly: INFO = 7 from DTRTRS
    assert lb.f77.gels("N", a0, b.copy(order="F")) == 7
    z = np.zeros((m, n), order="F")                                                 # zero matrix: zero solution (dgels.f:326-333)
    x = b.copy(order="F")
    assert lb.f77.gels("N", z, x) == 0 and np.all(x == 0.0)


# ------------------------------------------------------------------------------------------- DGEQRT / DGEMQRT
@pytest.mark.parametrize("m,n,nb", [(40, 40, 40), (300, 170, 32), (700, 700, 100), (1500, 520, 7), (200, 350, 64)])
def test_dgeqrt_dgemqrt_vs_oracle(lb, m, n, nb):
    a, _ = O.random_matrix(m, n, SEED)
    k = min(m, n)
    ref = a.copy(order="F")
    t_ref, info = O.dgeqrt(ref, nb)
    assert info == 0
    got = a.copy(order="F")
    t = np.full((nb, k), -1.0e10, order="F")
    work = np.zeros(nb * n)
    assert lb.f77.dgeqrt(m, n, nb, got, m, t, nb, work) == 0
    assert rel(got, ref) < 1e-11
    for i in range(0, k, nb):                                        # upper triangles of the T blocks; the rest is untouched
        ib = min(nb, k - i)
        blk, blk_ref = t[:ib, i:i + ib], t_ref[:ib, i:i + ib]
        assert rel(np.triu(blk), np.triu(blk_ref)) < 1e-11
        assert np.all(blk[np.tril_indices(ib, -1)] == -1.0e10)
    nc = 9
    for side in "LR":
        c0, _ = O.random_matrix(m if side == "L" else nc, nc if side == "L" else m, (3, 5, 7, 9))
        kk = k
        for trans in "NT":
            c = c0.copy(order="F")
            assert lb.f77.gemqrt(side, trans, got, t, c, nb, kk) == 0
            c_ref = c0.copy(order="F")
            assert O.dgemqrt(side, trans, ref, t_ref, c_ref, nb, kk) == 0
            assert rel(c, c_ref) < 1e-11, (side, trans)


# ------------------------------------------------------------------------------------------- DGERFS
@pytest.mark.parametrize("n,nrhs", [(1, 1), (40, 2), (300, 3), (1100, 2)])
@pytest.mark.parametrize("trans", "NT")
def test_dgerfs_vs_oracle(lb, n, nrhs, trans):
    a, seed = O.random_matrix(n, n, SEED)
    xt, _ = O.random_matrix(n, nrhs, seed)
    b = np.asfortranarray((a if trans == "N" else a.T) @ xt)
    af = a.copy(order="F")
    ipiv, info = O.dgetrf(af)
    x0 = np.asfortranarray(xt * (1.0 + 1e-9))                       # slightly wrong on purpose
    x_ref = x0.copy(order="F")
    ferr_ref, berr_ref, info_ref = O.dgerfs(trans, a, af, ipiv, b, x_ref)
    x = x0.copy(order="F")
    ferr, berr, info = lb.f77.gerfs(trans, a, af, ipiv, b, x)
    assert info == info_ref == 0
    assert rel(x, x_ref) < 1e-11
    assert np.all(berr < 4e-16 * max(1, n) ** 0.5) and np.all(berr_ref < 4e-16 * max(1, n) ** 0.5)
    assert np.all(np.abs(ferr - ferr_ref) <= 0.1 * ferr_ref + 1e-16)
    err = np.max(np.abs(x - xt), axis=0) / np.max(np.abs(x), axis=0)
    assert np.all(err <= ferr + 1e-16)                               # FERR is a bound on the true forward error (dgerfs.f:96-103)


# ------------------------------------------------------------------------------------------- DGETRI
@pytest.mark.parametrize("n", [1, 33, 300, 1100])
def test_dgetri_vs_oracle(lb, n):
    a, _ = O.random_matrix(n, n, SEED)
    lu = a.copy(order="F")
    ipiv, info = O.dgetrf(lu)
    ref = lu.copy(order="F")
    assert O.dgetri(ref, ipiv) == 0
    got = lu.copy(order="F")
    assert lb.f77.getri(got, ipiv) == 0
    scale = float(np.max(np.abs(ref)))
    assert np.max(np.abs(got - ref)) < 1e-10 * scale
    assert np.max(np.abs(got @ a - np.eye(n))) < 1e-9
    # factor + invert entirely on the device
    d = np_to_dev(lb, a)
    piv_d, _ = lb.dev.getrf(d)
    assert int(lb.dev.getri(d, piv_d).item()) == 0
    assert np.max(np.abs(dev_to_np(d) - ref)) < 1e-10 * scale
    if n >= 33:
        # exactly singular U: INFO = i and A keeps its factors (dtrtri.f:169-175, dgetri.f:181-183)
        sing = lu.copy(order="F")
        sing[20, 20] = 0.0
        before = sing.copy()
        assert lb.f77.getri(sing, ipiv) == 21
        assert np.array_equal(sing, before)


# ------------------------------------------------------------------------------------------- DORGQR / DORMQR
@pytest.mark.parametrize("shape", [(1, 1), (40, 40), (300, 170), (700, 700), (1500, 520)])
def test_dorgqr_dormqr_vs_oracle(lb, shape):
    m, n = shape
    a, _ = O.random_matrix(m, n, SEED)
    af = a.copy(order="F")
    tau, info, _ = lb.f77.geqrf(af)
    assert info == 0
    k = min(m, n)
    # DORGQR: first n columns of Q, vs the oracle on the same reflectors; orthogonality and A = Q R
    q = af.copy(order="F")
    assert lb.f77.orgqr(q, tau) == 0
    q_ref = af.copy(order="F")
    assert O.dorgqr(q_ref, tau) == 0
    assert rel(q, q_ref) < 1e-12
    assert np.max(np.abs(q.T @ q - np.eye(n))) < 1e-12 * max(m, 10)
    assert rel(q @ np.triu(af[:n, :]), a) < 1e-12 * max(n, 10)
    # partial generation (n_q < k columns, k_q < n_q reflectors) like dqrt02.f
    if n >= 8:
        nq, kq = n - 3, n - 5
        q2 = af[:, :nq].copy(order="F")
        assert lb.f77.orgqr(q2, tau[:kq]) == 0
        q2_ref = af[:, :nq].copy(order="F")
        assert O.dorgqr(q2_ref, tau[:kq]) == 0
        assert rel(q2, q2_ref) < 1e-12
    # DORMQR: all four SIDE/TRANS combinations against the oracle (dqrt03.f checks the same products)
    nc = 11
    for side in "LR":
        c0, _ = O.random_matrix(m if side == "L" else nc, nc if side == "L" else m, (3, 5, 7, 9))
        for trans in "NT":
            c = c0.copy(order="F")
            assert lb.f77.ormqr(side, trans, af, tau, c) == 0
            c_ref = c0.copy(order="F")
            assert O.dormqr(side, trans, af, tau, c_ref) == 0
            assert rel(c, c_ref) < 1e-12, (side, trans)
    # Q^T A = R: the use the reference's own checker makes of DORMQR (dqrt01.f:191-200 forms it with DORGQR)
    c = a.copy(order="F")
    assert lb.f77.ormqr("L", "T", af, tau, c) == 0
    assert rel(np.triu(c[:k, :]), np.triu(af[:k, :])) < 1e-12 * max(n, 10)
    assert np.max(np.abs(np.tril(c, -1))) < 1e-12 * max(m, 10) * np.max(np.abs(a))


def test_lapacke_dorgqr_dormqr_row_major(lb):
    import ctypes as C
    L = lb.lib()
    m, n, nc = 120, 70, 9
    a, _ = O.random_matrix(m, n, SEED)
    af = a.copy(order="F")
    tau, info, _ = lb.f77.geqrf(af)
    dp = C.POINTER(C.c_double)
    ptr = lambda x: x.ctypes.data_as(dp)
    L.LAPACKE_dorgqr.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, dp, C.c_int, dp]
    L.LAPACKE_dormqr.argtypes = [C.c_int, C.c_char, C.c_char, C.c_int, C.c_int, C.c_int, dp, C.c_int, dp, dp, C.c_int]
    q_ref = af.copy(order="F")
    O.dorgqr(q_ref, tau)
    q_rm = np.array(af, order="C", copy=True)                       # row-major storage, lda = n
    assert L.LAPACKE_dorgqr(101, m, n, n, ptr(q_rm), n, ptr(tau)) == 0
    assert rel(q_rm, q_ref) < 1e-12
    c0, _ = O.random_matrix(m, nc, (3, 5, 7, 9))
    c_ref = c0.copy(order="F")
    O.dormqr("L", "T", af, tau, c_ref)
    a_rm = np.array(af, order="C", copy=True)
    c_rm = np.array(c0, order="C", copy=True)
    assert L.LAPACKE_dormqr(101, b"L", b"T", m, nc, n, ptr(a_rm), n, ptr(tau), ptr(c_rm), nc) == 0
    assert rel(c_rm, c_ref) < 1e-12
    c_cm = c0.copy(order="F")
    assert L.LAPACKE_dormqr(102, b"L", b"T", m, nc, n, ptr(af), m, ptr(tau), ptr(c_cm), m) == 0
    assert rel(c_cm, c_ref) < 1e-12
    assert L.LAPACKE_dormqr(101, b"L", b"T", m, nc, n, ptr(a_rm), n - 1, ptr(tau), ptr(c_rm), nc) == -8   # lda < k


# ------------------------------------------------------------------------------------------- batched 32x32
def test_batched_getrf32(lb):
    batch = 3000
    a = lb.dev.larnv_matrix(32, 32 * batch, SEED)                  # column-major 32 x (32*batch) == batch matrices
    a_np = dev_to_np(a).copy()
    mats = a.t().contiguous().view(batch, 32, 32)                  # [b][col][row]
    # make a few matrices singular / tiny
    mats[5, 3, :] = 0.0
    mats[6, :, :] = 0.0
    mats[7, 0, :] *= 1e-310
    ref_in = mats.cpu().numpy().copy()
    ipiv, info = lb.dev.getrf_batched32(mats)
    torch.cuda.synchronize()
    out = mats.cpu().numpy()
    ipiv = ipiv.cpu().numpy()
    info = info.cpu().numpy()
    for b in list(range(0, 40)) + list(range(batch - 20, batch)):
        x = np.asfortranarray(ref_in[b].T)
        r = x.copy(order="F")
        ipiv_ref, info_ref = O.dgetrf2(r)
        assert info[b] == info_ref, b
        assert np.array_equal(ipiv[b], ipiv_ref), b
        assert rel(out[b].T, r) < 1e-12, b
        if info_ref == 0:
            assert O.dget01(x, np.asfortranarray(out[b].T), ipiv[b]) < O.THRESH
    assert a_np.shape == (32, 32 * batch)


def test_batched_potrf32(lb):
    batch = 1000
    a = lb.dev.larnv_matrix(32, 32 * batch, SEED)
    mats = a.t().contiguous().view(batch, 32, 32)
    mats.copy_((mats + mats.transpose(1, 2)) * 0.5 + 32.0 * torch.eye(32, device=mats.device, dtype=mats.dtype))
    mats[3, 10, 10] = -1.0
    ref_in = mats.cpu().numpy().copy()
    for uplo in "LU":
        work = mats.clone()
        info = lb.dev.potrf_batched32(uplo, work).cpu().numpy()
        out = work.cpu().numpy()
        for b in range(0, 30):
            x = np.asfortranarray(ref_in[b].T)
            r = x.copy(order="F")
            info_ref = O.dpotrf2(uplo, r)
            assert info[b] == info_ref, (uplo, b)
            if info_ref == 0:
                tri = np.tril if uplo == "L" else np.triu
                assert rel(tri(out[b].T), tri(r)) < 1e-13
                assert O.dpot01(uplo, x, np.asfortranarray(out[b].T)) < O.THRESH


# ------------------------------------------------------------------------------------------- device-pointer API
def test_device_api_getrf_getrs(lb):
    n = 1000
    a = lb.dev.larnv_matrix(n, n, SEED)
    a0 = dev_to_np(a)
    want, _ = O.random_matrix(n, n, SEED)
    assert np.array_equal(a0, want)
    ipiv, info = lb.dev.getrf(a)
    ref = want.copy(order="F")
    ipiv_ref, _ = O.dgetrf(ref)
    assert int(info.item()) == 0
    assert np.array_equal(ipiv.cpu().numpy(), ipiv_ref)
    b = lb.dev.larnv_matrix(n, 2, SEED, offset=n * n)
    b0 = dev_to_np(b)
    lb.dev.getrs("N", a, ipiv, b)
    assert O.dget02("N", want, dev_to_np(b), b0) < O.THRESH


def test_fortran_abi_from_several_host_threads(lb):
    """SURVEY 8b 'Threading': the replacement must be callable from multiple host threads and synchronous on return.
    ctypes releases the GIL during the call, so these calls really overlap; the library serialises them."""
    import threading
    n = 700
    mats = [O.random_matrix(n, n, (11 + 2 * t, 3, 5, 7))[0] for t in range(6)]
    refs = []
    for a in mats:
        r = a.copy(order="F")
        refs.append((r, O.dgetrf(r)[0]))
    out = [None] * len(mats)

    def work(t):
        got = mats[t].copy(order="F")
        if t % 2 == 0:
            ipiv, info = lb.f77.getrf(got)
            out[t] = (got, ipiv, info)
        else:
            s = np.asfortranarray(mats[t] @ mats[t].T + n * np.eye(n))
            f = s.copy(order="F")
            info = lb.f77.potrf("L", f)
            out[t] = (s, f, info)

    th = [threading.Thread(target=work, args=(t,)) for t in range(len(mats))]
    for x in th:
        x.start()
    for x in th:
        x.join()
    for t in range(len(mats)):
        if t % 2 == 0:
            got, ipiv, info = out[t]
            assert info == 0 and np.array_equal(ipiv, refs[t][1]) and rel(got, refs[t][0]) < 1e-10
        else:
            s, f, info = out[t]
            assert info == 0 and O.dpot01("L", s, f) < O.THRESH


def test_ilp64_api_matches_32bit(lb):
    """_64 entry points (64-bit INTEGER, 64-bit IPIV) give exactly the results of the 32-bit ones."""
    import ctypes as C
    L = lb.lib()
    i64 = C.c_int64
    vp = lambda x: x.ctypes.data_as(C.c_void_p)
    n, nrhs = 300, 3
    a, seed = O.random_matrix(n, n, SEED)
    xt, _ = O.random_matrix(n, nrhs, seed)
    b = np.asfortranarray(a @ xt)
    lu32 = a.copy(order="F")
    piv32, info32 = lb.f77.getrf(lu32)
    lu64, piv64, info = a.copy(order="F"), np.zeros(n, dtype=np.int64), i64(-9)
    L.dgetrf_64_(C.byref(i64(n)), C.byref(i64(n)), vp(lu64), C.byref(i64(n)), vp(piv64), C.byref(info))
    assert info.value == info32 == 0 and np.array_equal(piv64, piv32) and np.array_equal(lu64, lu32)
    x64 = b.copy(order="F")
    L.dgetrs_64_(b"N", C.byref(i64(n)), C.byref(i64(nrhs)), vp(lu64), C.byref(i64(n)), vp(piv64), vp(x64), C.byref(i64(n)),
                 C.byref(info), C.c_size_t(1))
    assert info.value == 0 and rel(x64, xt) < 1e-10
    s, _ = O.spd_matrix(n, SEED)
    f32, f64 = s.copy(order="F"), s.copy(order="F")
    assert lb.f77.potrf("L", f32) == 0
    L.dpotrf_64_(b"L", C.byref(i64(n)), vp(f64), C.byref(i64(n)), C.byref(info), C.c_size_t(1))
    assert info.value == 0 and np.array_equal(f64, f32)
    c32, c64 = np.zeros((n, n), order="F"), np.zeros((n, n), order="F")
    one, zero = C.c_double(1.0), C.c_double(0.0)
    lb.f77.dgemm("N", "T", n, n, n, 1.0, a, n, s, n, 0.0, c32, n)
    L.dgemm_64_(b"N", b"T", C.byref(i64(n)), C.byref(i64(n)), C.byref(i64(n)), C.byref(one), vp(a), C.byref(i64(n)), vp(s),
                C.byref(i64(n)), C.byref(zero), vp(c64), C.byref(i64(n)), C.c_size_t(1), C.c_size_t(1))
    assert np.array_equal(c64, c32)
    q32, q64 = a.copy(order="F"), a.copy(order="F")
    tau32, _, _ = lb.f77.geqrf(q32)
    tau64, work = np.zeros(n), np.zeros(n * 32)
    L.dgeqrf_64_(C.byref(i64(n)), C.byref(i64(n)), vp(q64), C.byref(i64(n)), vp(tau64), vp(work), C.byref(i64(len(work))),
                 C.byref(info))
    assert info.value == 0 and np.array_equal(q64, q32) and np.array_equal(tau64, tau32)
