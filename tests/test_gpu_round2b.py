"""GPU tests for the kernels changed in the second half of round 2: the lean interior GEMM loader, the triangular CTA grid of
one-triangle GEMMs, the wave-balancing split-K, and the two-matrices-per-warp batched DGETRF.  Each new path is compared bit for bit
with the path it replaced (same arithmetic order), and the ones whose summation order changes (split-K) against the oracle."""
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


def _bits_equal(x, y):
    return bool(torch.equal(x.view(torch.int64), y.view(torch.int64)))


@pytest.mark.parametrize("ta", "NT")
@pytest.mark.parametrize("tb", "NT")
def test_lean_loader_equals_general_loader(lb, ta, tb):
    """cfg 13 / 14 (interior tiles through the lean loader, edge tiles through the general one) == cfg 8 (general loader only),
    bit for bit, on shapes with and without edges, odd leading dimensions (no 16-byte path) and K that is not a multiple of 16"""
    L = lb.lib()
    try:
        for (m, n, k, pad) in ((256, 192, 64, 0), (200, 130, 48, 0), (320, 256, 40, 0), (257, 129, 64, 1), (64, 64, 16, 0), (640, 64, 512, 2)):
            ar, ac = (m, k) if ta == "N" else (k, m)
            br, bc = (k, n) if tb == "N" else (n, k)
            a = lb.dev.colmajor(ar, ac, ar + pad); a.normal_()
            b = lb.dev.colmajor(br, bc, br + pad); b.normal_()
            c0 = lb.dev.colmajor(m, n, m + pad); c0.normal_()
            outs = []
            for cfg in (8, 13, 14):
                L.lb200_set_gemm_config(cfg)
                c = c0.clone()
                lb.dev.gemm(ta, tb, -1.0, a, b, 1.0, c)
                torch.cuda.synchronize()
                outs.append(c)
            assert _bits_equal(outs[0], outs[1]) and _bits_equal(outs[0], outs[2]), (m, n, k, pad)
    finally:
        L.lb200_set_gemm_config(-1)


@pytest.mark.parametrize("uplo", "LU")
@pytest.mark.parametrize("trans", "NT")
def test_syrk_triangular_grid(lb, uplo, trans):
    """one-triangle launches enumerate only the tiles of the triangle: every entry of the triangle equals the full product computed
    by DGEMM on the same kernel, the other triangle and the padding rows keep their values (sizes around the 16-tile-row groups)"""
    for n in (64, 65, 1000, 1024, 1088, 2100):
        k = 96
        a = lb.dev.colmajor(n, k) if trans == "N" else lb.dev.colmajor(k, n)
        a.normal_()
        c0 = lb.dev.colmajor(n + 3, n); c0.normal_()
        c = c0.clone()
        lb.dev.syrk(uplo, trans, -1.0, a, 1.0, c[:n, :])
        full = c0.clone()
        if trans == "N":
            lb.dev.gemm("N", "T", -1.0, a, a, 1.0, full[:n, :])
        else:
            lb.dev.gemm("T", "N", -1.0, a, a, 1.0, full[:n, :])
        torch.cuda.synchronize()
        tri = torch.tril(torch.ones(n, n, dtype=torch.bool, device=c.device)) if uplo == "L" else torch.triu(torch.ones(n, n, dtype=torch.bool, device=c.device))
        assert _bits_equal(c[:n][tri], full[:n][tri]), n
        assert _bits_equal(c[:n][~tri], c0[:n][~tri]), n
        assert _bits_equal(c[n:], c0[n:]), n


def test_rectangular_triangle_gemm_in_cholesky(lb):
    """the look-ahead block column of DPOTRF is a rectangular one-triangle GEMM (tri = 1 / 2 with M != N): Cholesky of both triangles
    at a size that is not a multiple of the tile group, against the oracle"""
    n = 1100
    for uplo in "LU":
        a = lb.dev.larnv_matrix(n, n, SEED)
        lb.dev.make_spd(a, float(n))
        s, _ = O.spd_matrix(n, SEED)
        lb.lib().lb200_set_potrf_params(128, 1)
        try:
            info = lb.dev.potrf(uplo, a)
        finally:
            lb.lib().lb200_set_potrf_params(512, 1)
        torch.cuda.synchronize()
        assert int(info) == 0
        got = np.asfortranarray(a.cpu().numpy())
        assert O.dpot01(uplo, s, got) < O.THRESH


def test_splitk_balance_long_k(lb):
    """long-K GEMMs with a partly filled last wave are split along K (deterministic slice order): against the oracle, and the same
    bits on a second run"""
    rng = np.random.default_rng(11)
    m, n, k = 128, 4608, 2304            # 2 x 72 = 144 tiles: 0.24 of a wave
    a = np.asfortranarray(rng.uniform(-1, 1, (k, m)))
    b = np.asfortranarray(rng.uniform(-1, 1, (k, n)))
    c = np.asfortranarray(rng.uniform(-1, 1, (m, n)))
    want = c.copy(order="F")
    want[:] = 0.7 * (a.T @ b) + 1.3 * c
    da = lb.dev.colmajor(k, m); da.copy_(torch.from_numpy(np.ascontiguousarray(a)))
    db = lb.dev.colmajor(k, n); db.copy_(torch.from_numpy(np.ascontiguousarray(b)))
    outs = []
    for on in (1, 1, 0):
        lb.lib().lb200_set_gemm_splitk_balance(on)
        dc = lb.dev.colmajor(m, n); dc.copy_(torch.from_numpy(np.ascontiguousarray(c)))
        lb.dev.gemm("T", "N", 0.7, da, db, 1.3, dc)
        torch.cuda.synchronize()
        outs.append(dc)
    lb.lib().lb200_set_gemm_splitk_balance(1)
    g = 0.7 * (np.abs(a).T @ np.abs(b)) + 1.3 * np.abs(c)
    for o in outs:
        assert np.max(np.abs(o.cpu().numpy() - want) / g) / EPS < 16.0
    assert _bits_equal(outs[0], outs[1])


def test_batched_two_per_warp_equals_one_row_kernel(lb):
    """the default batched DGETRF (two matrices per warp) against the one-row kernel, bit for bit, on random matrices and on the
    special cases (zero matrix / column, ties, NaN, Inf, denormal pivots, rank one, odd batch), and against the oracle on samples"""
    L = lb.lib()
    g = torch.Generator(device="cpu"); g.manual_seed(11)
    sp = torch.randn(4097, 32, 32, dtype=torch.float64, generator=g)
    sp[0] = 0.0
    sp[1, :, 5] = 0.0
    sp[2] = 1.0
    sp[3, 7, 0] = float("nan")
    sp[4, 0, 0] = float("nan")
    sp[5, 9, 3] = float("inf")
    sp[6] = torch.randint(-2, 3, (32, 32), generator=g).double()
    sp[7] = sp[7] * 1e-310
    sp[8, :, :] = torch.arange(32, dtype=torch.float64).view(32, 1)
    for k in range(9, 64):
        sp[k] = torch.randint(-1, 2, (32, 32), generator=g).double()
    a0 = sp.transpose(1, 2).contiguous().cuda()                     # kernel layout: [b][col][row]
    res = {}
    try:
        for mode in (0, 2):
            L.lb200_set_batched_mode(mode)
            a = a0.clone()
            ipiv, info = lb.dev.getrf_batched32(a)
            torch.cuda.synchronize()
            res[mode] = (a, ipiv.clone(), info.clone())
    finally:
        L.lb200_set_batched_mode(2)
    assert _bits_equal(res[0][0], res[2][0])
    assert torch.equal(res[0][1], res[2][1]) and torch.equal(res[0][2], res[2][2])
    out = res[2][0].cpu().numpy()
    ipiv = res[2][1].cpu().numpy()
    info = res[2][2].cpu().numpy()
    inp = a0.cpu().numpy()
    # against the oracle: INFO everywhere; IPIV and the factors on the nonsingular samples (rank-deficient inputs eliminate to
    # rounding noise, where FMA vs separate multiply-add legitimately differ -- for those the bit-for-bit comparison above is the check)
    for b in (0, 1, 64, 100, 1000, 4096):
        x = np.asfortranarray(inp[b].T)
        r = x.copy(order="F")
        ipiv_ref, info_ref = O.dgetrf2(r)
        assert info[b] == info_ref, b
        if info_ref == 0 and b >= 64:
            assert np.array_equal(ipiv[b], ipiv_ref), b
            assert np.max(np.abs(out[b].T - r)) < 1e-11 * np.max(np.abs(r)), b


@pytest.mark.parametrize("shape", [(3000, 3000), (4100, 2300), (2300, 4100), (2600, 300)])
def test_dgetrf_deferred_composed_interchanges(lb, shape):
    """the interchanges left of the panels are kept, composed per block column and applied in one streaming pass at the end:
    IPIV and factors bit-identical to the immediate plan-by-plan application and to the plan-by-plan deferred one, and DGET01 passes"""
    m, n = shape
    L = lb.lib()
    a0 = lb.dev.larnv_matrix(m, n, SEED)
    outs = []
    try:
        L.lb200_set_getrf_params(128, 0, 1)
        for mode, tail in ((0, 512), (1, 128), (1, 1024), (2, 128)):
            L.lb200_set_getrf_defer_left(mode, tail)
            a = a0.clone()
            piv, info = lb.dev.getrf(a)
            torch.cuda.synchronize()
            assert int(info) == 0
            outs.append((piv.clone(), a))
    finally:
        L.lb200_set_getrf_params(512, 0, 1)
        L.lb200_set_getrf_defer_left(1, 512)
    for piv, a in outs[1:]:
        assert torch.equal(piv, outs[0][0]) and _bits_equal(a, outs[0][1])
    x = np.asfortranarray(a0.cpu().numpy())
    got = np.asfortranarray(outs[1][1].cpu().numpy())
    assert O.dget01(x, got, outs[1][0].cpu().numpy()) < O.THRESH


def test_long_pivot_lists_on_many_columns(lb):
    """DLASWP with a long pivot list on many columns is composed once and streamed (laswp_impl): the interchanges of DGETRS with
    many right-hand sides, forward and reverse, against an independent application by the short-list path, and the two interchange
    sweeps of the host-streamed square DGETRF (n1 >= 4096) against the device-resident driver"""
    n, nrhs = 5000, 96
    a = lb.dev.larnv_matrix(n, n, SEED)
    piv, info = lb.dev.getrf(a)
    assert int(info) == 0
    b0 = lb.dev.larnv_matrix(n, nrhs, SEED, n * n)
    for trans in "NT":
        x = b0.clone()
        lb.dev.getrs(trans, a, piv, x)                         # nrhs > 64, n pivots: composed + streamed interchanges
        # reference: the same solve in column groups of 48 (n pivots on <= 64 columns: gather / copy-back path)
        y = b0.clone()
        for c0 in range(0, nrhs, 48):
            lb.dev.getrs(trans, a, piv, y[:, c0:c0 + 48])
        torch.cuda.synchronize()
        # (the triangular solves may split K differently for 96 and 48 columns: agreement to rounding, not bit for bit)
        assert float((x - y).abs().max()) < 1e-11 * float(y.abs().max()), trans
    # host-streamed DGETRF with both sweeps long
    n = 12288
    a0 = lb.dev.larnv_matrix(n, n, SEED)
    buf = torch.empty((n, n), dtype=torch.float64).pin_memory()
    h = buf.numpy().T
    h[:] = a0.cpu().numpy()
    ipiv = np.zeros(n, dtype=np.int32)
    assert lb.f77.dgetrf(n, n, h, n, ipiv) == 0
    pd, infod = lb.dev.getrf(a0)
    torch.cuda.synchronize()
    assert np.array_equal(ipiv, pd.cpu().numpy())
    got = torch.from_numpy(np.ascontiguousarray(h)).to(a0.device)
    scale = float(a0.abs().max())
    assert float((got - a0).abs().max()) < 1e-10 * scale


@pytest.mark.parametrize("shape", [(3300, 3300), (4200, 3100)])
def test_dgetrf_two_level_driver(lb, shape):
    """lb200_set_getrf_super: the outer level (NB = 1024 here) whose panel is the NB = 256 driver gives LAPACK's pivots and, up to
    rounding (different K-blocking of the updates), the factors of the single-level driver; DGET01 passes"""
    m, n = shape
    L = lb.lib()
    a0 = lb.dev.larnv_matrix(m, n, SEED)
    try:
        L.lb200_set_getrf_params(256, 0, 1)
        L.lb200_set_getrf_super(0)
        a1 = a0.clone(); p1, i1 = lb.dev.getrf(a1)
        L.lb200_set_getrf_super(1024)
        a2 = a0.clone(); p2, i2 = lb.dev.getrf(a2)
        torch.cuda.synchronize()
    finally:
        L.lb200_set_getrf_super(0)
        L.lb200_set_getrf_params(512, 0, 1)
    assert int(i1) == 0 and int(i2) == 0
    assert torch.equal(p1, p2)
    assert float((a1 - a2).abs().max()) < 1e-11 * float(a1.abs().max())
    x = np.asfortranarray(a0.cpu().numpy())
    assert O.dget01(x, np.asfortranarray(a2.cpu().numpy()), p2.cpu().numpy()) < O.THRESH


@pytest.mark.parametrize("uplo", "LU")
def test_dpotrf_inverted_diagonal_block_leaves(lb, uplo):
    """DPOTRF's panel solve uses inverted 32 x 32 diagonal blocks (DMMA leaves) once the panel has >= 1024 rows / columns: both
    triangles at a size that reaches it, against the substitution leaves (rounding-level agreement) and DPOT01"""
    n = 2300
    L = lb.lib()
    a0 = lb.dev.larnv_matrix(n, n, SEED)
    lb.dev.make_spd(a0, float(n))
    s, _ = O.spd_matrix(n, SEED)
    outs = []
    try:
        for inv in (1, 0):
            L.lb200_set_trsm_inverse(inv)
            a = a0.clone()
            info = lb.dev.potrf(uplo, a)
            torch.cuda.synchronize()
            assert int(info) == 0
            outs.append(a)
    finally:
        L.lb200_set_trsm_inverse(1)
    tri = torch.tril if uplo == "L" else torch.triu
    assert float((tri(outs[0]) - tri(outs[1])).abs().max()) < 1e-12 * float(tri(outs[1]).abs().max())
    assert O.dpot01(uplo, s, np.asfortranarray(outs[0].cpu().numpy())) < O.THRESH
