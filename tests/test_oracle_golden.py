"""Pins the CPU oracle (oracle/) against the committed netlib-3.12.0 golden vectors
(tests/golden/make_golden.py) and the DLARNV known answer from SURVEY.md section 7.

Tolerances: IPIV / INFO / RNG bit-exact; factor entries 1e-12 relative (the golden side ran on
OpenBLAS BLAS, a different summation order than the reference BLAS triple loops the oracle restates).
"""
import os

import numpy as np
import pytest

from oracle import oracle as O

SEED = (1988, 1989, 1990, 1991)
RTOL = 1e-12


def relerr(x, y):
    d = np.max(np.abs(x - y)) if x.size else 0.0
    s = max(1.0, np.max(np.abs(y))) if y.size else 1.0
    return d / s


def test_dlarnv_known_answer():
    x, seed = O.dlarnv(2, SEED, 5)
    assert list(x) == [-0.5217827788720584, -0.08059010722889326, -0.46509858502694357,
                       0.07496184221979973, -0.745637131925811]
    assert seed == [520, 3830, 1597, 2307]


@pytest.mark.parametrize("idist", [1, 2, 3])
def test_dlarnv_stream(golden, idist):
    x, seed = O.dlarnv(idist, SEED, 300)
    ref = golden[f"larnv{idist}_x"]
    if idist == 3:   # log/cos come from different libm builds
        assert np.allclose(x, ref, rtol=1e-14, atol=1e-15)
    else:
        assert np.array_equal(x, ref)
    assert seed == list(golden[f"larnv{idist}_seed"])


@pytest.mark.parametrize("tag", ["tall", "sq", "wide", "one", "col", "sing"])
def test_dgetrf2(golden, tag):
    a = np.array(golden[f"getrf2_{tag}_a"], order="F")
    if tag != "sing":
        regen, _ = O.random_matrix(*a.shape, SEED)
        assert np.array_equal(regen, a)      # generator parity with the golden inputs
    ipiv, info = O.dgetrf2(a)
    assert info == int(golden[f"getrf2_{tag}_info"])
    assert np.array_equal(ipiv, golden[f"getrf2_{tag}_ipiv"])
    assert relerr(a, golden[f"getrf2_{tag}_lu"]) < RTOL


@pytest.mark.parametrize("nb", [1, 3, 20, 64])
def test_dgetrf_blocked_equals_recursive_pivots(golden, nb):
    a = np.array(golden["getrf2_tall_a"], order="F")
    O.set_nb(getrf=nb)
    try:
        ipiv, info = O.dgetrf(a)
    finally:
        O.set_nb()
    assert info == 0
    assert np.array_equal(ipiv, golden["getrf2_tall_ipiv"])
    assert relerr(a, golden["getrf2_tall_lu"]) < RTOL


@pytest.mark.parametrize("uplo", ["L", "U"])
def test_dpotrf2(golden, uplo):
    a = np.array(golden[f"potrf2_{uplo}_a"], order="F")
    a0 = a.copy(order="F")
    info = O.dpotrf2(uplo, a)
    assert info == int(golden[f"potrf2_{uplo}_info"]) == 0
    tri = np.tril if uplo == "L" else np.triu
    assert relerr(tri(a), tri(golden[f"potrf2_{uplo}_f"])) < RTOL
    other = (lambda x: np.triu(x, 1)) if uplo == "L" else (lambda x: np.tril(x, -1))
    assert np.array_equal(other(a), other(a0))          # opposite triangle untouched
    for nb in (1, 3, 20):
        b = a0.copy(order="F")
        O.set_nb(potrf=nb)
        try:
            assert O.dpotrf(uplo, b) == 0
        finally:
            O.set_nb()
        assert relerr(tri(b), tri(golden[f"potrf2_{uplo}_f"])) < RTOL


def test_dpotrf2_not_spd(golden):
    a = np.array(golden["potrf2_bad_a"], order="F")
    assert O.dpotrf2("L", a) == int(golden["potrf2_bad_info"]) == 11
    b = np.array(golden["potrf2_bad_a"], order="F")
    O.set_nb(potrf=4)
    try:
        assert O.dpotrf("L", b) == 11
    finally:
        O.set_nb()


def test_dpotrs(golden):
    f = np.array(golden["potrf2_L_f"], order="F")
    x = np.array(golden["potrs_b"], order="F")
    assert O.dpotrs("L", f, x) == 0
    assert relerr(x, golden["potrs_x"]) < RTOL


@pytest.mark.parametrize("tag", ["", "tiny_"])
def test_dlarfg(golden, tag):
    v = np.array(golden[f"larfg_{tag}in"])
    x = v[1:].copy()
    beta, tau = O.dlarfg(float(v[0]), x)
    assert abs(beta - golden[f"larfg_{tag}beta"]) <= 4e-16 * abs(golden[f"larfg_{tag}beta"])
    assert abs(tau - golden[f"larfg_{tag}tau"]) <= 1e-15
    assert relerr(x, golden[f"larfg_{tag}v"]) < 1e-14


@pytest.mark.parametrize("tag", ["tall", "sq", "wide"])
def test_dgeqr2(golden, tag):
    a = np.array(golden[f"geqrf_{tag}_a"], order="F")
    tau, info = O.dgeqr2(a)
    assert info == 0
    assert relerr(a, golden[f"geqr2_{tag}_qr"]) < RTOL
    assert relerr(tau, golden[f"geqr2_{tag}_tau"]) < RTOL


@pytest.mark.parametrize("tag", ["tall", "sq", "wide", "big"])
def test_dgeqrf(golden, tag):
    a = np.array(golden[f"geqrf_{tag}_a"], order="F")
    m, n = a.shape
    tau, info, w1 = O.dgeqrf(a)
    assert info == 0
    assert relerr(a, golden[f"geqrf_{tag}_qr"]) < RTOL
    assert relerr(tau, golden[f"geqrf_{tag}_tau"]) < RTOL
    if tag == "big":
        assert w1 == n * 32        # k=150 > NX=128: the blocked DLARFT/DLARFB path ran (IWS = N*NB)
    res = O.dqrt01(np.array(golden[f"geqrf_{tag}_a"], order="F"), a, tau)
    assert res[0] < O.THRESH and res[1] < O.THRESH


def test_dgeqrf_small_blocks(golden):
    """Force the blocked path (NB=8, NX=0) like TESTING/LIN does through xlaenv."""
    a = np.array(golden["geqrf_tall_a"], order="F")
    O.set_nb(geqrf=8, nx=0)
    try:
        tau, info, _ = O.dgeqrf(a)
    finally:
        O.set_nb()
    assert relerr(a, golden["geqrf_tall_qr"]) < RTOL
    assert relerr(tau, golden["geqrf_tall_tau"]) < RTOL


def test_dlarft_dlarfb_dorgqr(golden):
    qr = np.array(golden["geqrf_tall_qr"], order="F")
    tau = np.array(golden["geqrf_tall_tau"])
    t = O.dlarft(qr, tau)
    assert relerr(np.triu(t), golden["larft_t"]) < RTOL
    for trans in ("T", "N"):
        c = np.array(golden["larfb_c_in"], order="F")
        O.dlarfb("L", trans, qr, t, c)
        assert relerr(c, golden[f"larfb_L{trans}_c"]) < RTOL
    m, k = qr.shape
    q = np.zeros((m, m), order="F")
    q[:, :k] = np.tril(qr, -1)[:, :k]
    assert O.dorgqr(q, tau, k) == 0
    assert relerr(q, golden["orgqr_q"]) < RTOL
    assert np.max(np.abs(q.T @ q - np.eye(m))) < 1e-13


def test_dlarft_recursive_branch():
    """k >= 64 takes the recursive Level-3 branch of dlarft.f:308-349; compare with the Level-2 one."""
    a, _ = O.random_matrix(150, 80, SEED)
    tau, info = O.dgeqr2(a)
    t_rec = O.dlarft(a, tau)
    t_l2 = O.fmat(80, 80)
    import ctypes as C
    O.lib().ora_dlarft_lvl2(C.c_char(b"F"), C.c_char(b"C"), 150, 80, O._d(a), 150, O._d(tau), O._d(t_l2), 80)
    assert relerr(np.triu(t_rec), np.triu(t_l2)) < 1e-13


def test_checkers_flag_bad_factorizations(golden):
    a = np.array(golden["getrf2_sq_a"], order="F")
    lu = np.array(golden["getrf2_sq_lu"], order="F")
    ipiv = np.array(golden["getrf2_sq_ipiv"])
    assert O.dget01(a, lu, ipiv) < 1.0
    lu[5, 7] += 1e-6
    assert O.dget01(a, lu, ipiv) > O.THRESH
    s = np.array(golden["potrf2_L_a"], order="F")
    f = np.array(golden["potrf2_L_f"], order="F")
    assert O.dpot01("L", s, f) < 1.0
    assert O.dpot01("U", s, np.array(golden["potrf2_U_f"], order="F")) < 1.0
    f[20, 3] += 1e-6
    assert O.dpot01("L", s, f) > O.THRESH


def test_dgesv_solution_and_residuals():
    n = 200
    a, seed = O.random_matrix(n, n, SEED)
    xact, _ = O.random_matrix(n, 2, seed)
    b = np.asfortranarray(a @ xact)
    lu = a.copy(order="F")
    x = b.copy(order="F")
    ipiv, info = O.dgesv(lu, x)
    assert info == 0
    assert O.dget01(a, lu, ipiv) < O.THRESH
    assert O.dget02("N", a, x, b) < O.THRESH
    rcond = 1.0 / np.linalg.cond(a, 1)
    assert O.dget04(x, xact, rcond) < O.THRESH
    xt = b.copy(order="F")
    assert O.dgetrs("T", lu, ipiv, xt) == 0
    assert O.dget02("T", a, xt, b) < O.THRESH


def test_dposv_solution_and_residuals():
    n = 150
    s, seed = O.spd_matrix(n, SEED)
    xact, _ = O.random_matrix(n, 3, seed)
    b = np.asfortranarray(s @ xact)
    for uplo in ("L", "U"):
        f = s.copy(order="F")
        x = b.copy(order="F")
        assert O.dposv(uplo, f, x) == 0
        assert O.dpot01(uplo, s, f) < O.THRESH
        assert O.dpot02(uplo, s, x, b) < O.THRESH
        assert np.max(np.abs(x - xact)) / np.max(np.abs(xact)) < 1e-12


@pytest.mark.parametrize("tag", ["tall", "sq"])
def test_dormqr_dorgqr_netlib(tag):
    """ora_dormqr / ora_dorgqr vs netlib 3.12.0 (tests/golden/make_golden_ormqr.py)."""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "netlib_golden_ormqr.npz"))
    qr, tau = np.asfortranarray(g[f"{tag}_qr"]), g[f"{tag}_tau"]
    for side, c0 in (("L", g[f"{tag}_cl"]), ("R", g[f"{tag}_cr"])):
        for trans in "NT":
            c = np.asfortranarray(c0.copy())
            assert O.dormqr(side, trans, qr, tau, c) == 0
            assert np.max(np.abs(c - g[f"{tag}_ormqr_{side}{trans}"])) < 1e-12
    q = qr.copy(order="F")
    assert O.dorgqr(q, tau) == 0
    assert np.max(np.abs(q - g[f"{tag}_q"])) < 1e-12
    # unblocked path (K <= NB) agrees with the blocked one
    O.set_nb(geqrf=128)
    try:
        c = np.asfortranarray(g[f"{tag}_cl"].copy())
        assert O.dormqr("L", "T", qr, tau, c) == 0
        assert np.max(np.abs(c - g[f"{tag}_ormqr_LT"])) < 1e-12
    finally:
        O.set_nb()


@pytest.mark.parametrize("n", [7, 70])
def test_dgetri_dtrtri_netlib(n):
    """ora_dgetri / ora_dtrtri vs netlib 3.12.0 (tests/golden/make_golden_getri.py); blocked and unblocked paths."""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "netlib_golden_getri.npz"))
    lu, ipiv, inv_ref, a = np.asfortranarray(g[f"lu{n}"]), g[f"ipiv{n}"], g[f"inv{n}"], g[f"a{n}"]
    scale = np.max(np.abs(inv_ref))
    for nb in (None, 1, 8):
        x = lu.copy(order="F")
        assert O.dgetri(x, ipiv, nb=nb) == 0
        assert np.max(np.abs(x - inv_ref)) < 1e-11 * scale
        assert np.max(np.abs(x @ a - np.eye(n))) < 1e-10
    for uplo in "UL":
        t = np.asfortranarray(g[f"tri{uplo}{n}"].copy())
        assert O.dtrtri(uplo, "N", t) == 0
        assert np.max(np.abs(t - g[f"triinv{uplo}{n}"])) < 1e-12
    # exactly singular U: INFO = i, nothing computed (dtrtri.f:169-175, dgetri.f:181-183)
    x = lu.copy(order="F")
    x[3, 3] = 0.0
    before = x.copy()
    assert O.dgetri(x, ipiv) == 4
    assert np.array_equal(x, before)


@pytest.mark.parametrize("tag", ["tall", "sq", "wide"])
def test_dgeqrt_dgemqrt_netlib(tag):
    """ora_dgeqrt (recursive DGEQRT3 panels) / ora_dgemqrt vs netlib 3.12.0 (tests/golden/make_golden_geqrt.py)."""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "netlib_golden_geqrt.npz"))
    a, nb = np.asfortranarray(g[f"{tag}_a"]), int(g[f"{tag}_nb"])
    m, n = a.shape
    k = min(m, n)
    x = a.copy(order="F")
    t, info = O.dgeqrt(x, nb)
    assert info == 0
    assert np.max(np.abs(x - g[f"{tag}_qr"])) < 1e-12
    for i in range(0, k, nb):                                    # only the upper triangles of the T blocks are defined
        ib = min(nb, k - i)
        assert np.max(np.abs(np.triu(t[:ib, i:i + ib]) - np.triu(g[f"{tag}_t"][:ib, i:i + ib]))) < 1e-12
    for side, c0 in (("L", g[f"{tag}_cl"]), ("R", g[f"{tag}_cr"])):
        for trans in "NT":
            c = np.asfortranarray(c0.copy())
            assert O.dgemqrt(side, trans, x, t, c, nb, k) == 0
            assert np.max(np.abs(c - g[f"{tag}_gemqrt_{side}{trans}"])) < 1e-12
    # same R, V as DGEQRF (the factorization is unique up to rounding)
    y = a.copy(order="F")
    tau, _, _ = O.dgeqrf(y)
    assert np.max(np.abs(x - y)) < 1e-12


def test_dgels_netlib():
    """ora_dgels (QR and LQ paths, both TRANS, scaling branches, rank deficiency) vs netlib 3.12.0
    (tests/golden/make_golden_gels.py)."""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "netlib_golden_gels.npz"))
    for tag in ("tall", "wide", "sq"):
        a, b = np.asfortranarray(g[f"{tag}_a"]), np.asfortranarray(g[f"{tag}_b"])
        m, n = a.shape
        for trans in "NT":
            af, x = a.copy(order="F"), b.copy(order="F")
            assert O.dgels(trans, af, x) == 0
            rows = n if trans == "N" else m
            assert np.max(np.abs(x[:rows] - g[f"{tag}_{trans}_x"][:rows])) < 1e-12
            assert np.max(np.abs(af - g[f"{tag}_{trans}_af"])) < 1e-12        # QR / LQ factors (DGELQ2 == DGELQF up to rounding)
    a, b = np.asfortranarray(g["tall_a"]), np.asfortranarray(g["tall_b"])
    for k in range(4):
        sa, sb = g[f"scale{k}_s"]
        af, x = (a * sa).copy(order="F"), (b * sb).copy(order="F")
        assert O.dgels("N", af, x) == 0
        ref = g[f"scale{k}_x"][:40]
        assert np.max(np.abs(x[:40] - ref)) <= 1e-12 * np.max(np.abs(ref))
    a0 = a.copy(order="F")
    a0[:, 4] = 0.0
    assert O.dgels("N", a0, b.copy(order="F")) == int(g["rankdef_info"]) == 5


@pytest.mark.parametrize("n", [7, 80])
def test_dgerfs_netlib(n):
    """ora_dgerfs / ora_dlacn2 vs netlib 3.12.0 (tests/golden/make_golden_gerfs.py): refined solution to rounding, FERR within
    a few percent (its value depends on the rounding of the residual), BERR at the rounding level in both."""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "netlib_golden_gerfs.npz"))
    a, af, ipiv = np.asfortranarray(g[f"a{n}"]), np.asfortranarray(g[f"af{n}"]), g[f"ipiv{n}"]
    for trans in "NT":
        b, x = np.asfortranarray(g[f"b{n}{trans}"]), np.asfortranarray(g[f"x0_{n}{trans}"].copy())
        ferr, berr, info = O.dgerfs(trans, a, af, ipiv, b, x)
        assert info == 0
        assert np.max(np.abs(x - g[f"x{n}{trans}"])) < 1e-12 * np.max(np.abs(x))
        assert np.all(np.abs(ferr - g[f"ferr{n}{trans}"]) <= 0.05 * g[f"ferr{n}{trans}"])
        assert np.all(berr < 1e-15) and np.all(g[f"berr{n}{trans}"] < 1e-15)


# ------------------------------------------------------------------------------------------- DLATRS / DGECON / DGESVX
@pytest.fixture(scope="module")
def golden_gecon():
    return np.load(os.path.join(os.path.dirname(__file__), "golden", "netlib_golden_gecon.npz"))


def _close(x, y, tol):
    x, y = np.asarray(x, dtype=float), np.asarray(y, dtype=float)
    fin = np.isfinite(y)
    if not np.array_equal(np.isfinite(x), fin):
        return False
    if not fin.any():
        return True
    s = max(1e-300, float(np.max(np.abs(y[fin]))))
    return float(np.max(np.abs(x[fin] - y[fin]))) / s < tol


def test_dlatrs_matches_netlib(golden_gecon):
    """SRC/dlatrs.f incl. the scaled Level-1 branch (tiny / zero diagonal, column norms beyond BIGNUM): SCALE exactly, x and CNORM
    to rounding, NORMIN='Y' reusing the norms."""
    g = golden_gecon
    seen_scaled = 0
    for k in range(int(g["n_trs"][0])):
        uplo, trans, diag = (chr(v) for v in g[f"trs{k}_meta"])
        a, x = np.asfortranarray(g[f"trs{k}_a"]), g[f"trs{k}_x"].copy()
        cn = np.zeros(len(x))
        scale, info = O.dlatrs(uplo, trans, diag, "N", a, x, cn)
        assert info == 0
        s_ref, s_ref2 = g[f"trs{k}_scale"]
        assert scale == s_ref or abs(scale - s_ref) <= 1e-13 * abs(s_ref), (k, uplo, trans, diag, scale, s_ref)
        assert _close(x, g[f"trs{k}_xo"], 1e-11), (k, uplo, trans, diag)
        assert _close(cn, g[f"trs{k}_cn"], 1e-14)
        seen_scaled += scale != 1.0
        x2 = g[f"trs{k}_x"].copy()
        scale2, _ = O.dlatrs(uplo, trans, diag, "Y", a, x2, cn)
        assert scale2 == s_ref2 or abs(scale2 - s_ref2) <= 1e-13 * abs(s_ref2)
        assert _close(x2, g[f"trs{k}_xo2"], 1e-11)
    assert seen_scaled >= 8                                              # the scaled branch really ran


def test_dgecon_matches_netlib(golden_gecon):
    g = golden_gecon
    for k in range(int(g["n_con"][0])):
        lu = np.asfortranarray(g[f"con{k}_lu"])
        for norm in "1I":
            anorm, rc_ref, info_ref = g[f"con{k}_{norm}"]
            rc, info = O.dgecon(norm, lu, anorm)
            assert info == int(info_ref)
            # the estimate goes through triangular solves (OpenBLAS DTRSV on the golden side): rounding shows up amplified
            # for numerically singular factors (Hilbert n=60, rcond ~ 3e-20)
            tol = 1e-11 if rc_ref > 1e-12 else 1e-6
            assert rc == rc_ref or abs(rc - rc_ref) <= tol * abs(rc_ref), (k, norm, rc, rc_ref)


def test_dgesvx_matches_netlib(golden_gecon):
    g = golden_gecon
    for k in range(int(g["n_svx"][0])):
        fact, trans, equed_ref, info_ref = g[f"svx{k}_meta"]
        fact, trans, equed_ref = chr(fact), chr(trans), chr(equed_ref)
        a, b = np.asfortranarray(g[f"svx{k}_a"].copy()), np.asfortranarray(g[f"svx{k}_b"].copy())
        n = a.shape[0]
        af, ipiv, r, c = np.zeros((n, n), order="F"), np.zeros(n, dtype=np.int32), np.zeros(n), np.zeros(n)
        res = O.dgesvx(fact, trans, a, af, ipiv, "N", r, c, b)
        assert res["info"] == int(info_ref) and res["equed"] == equed_ref, (k, res["info"], info_ref, res["equed"], equed_ref)
        assert _close(a, g[f"svx{k}_a_out"], 1e-14) and _close(b, g[f"svx{k}_b_out"], 1e-14)
        rc_ref, rpv_ref = g[f"svx{k}_scal"]
        assert abs(res["rpvgrw"] - rpv_ref) <= 1e-12 * abs(rpv_ref)
        if equed_ref in "RB":
            assert _close(r, g[f"svx{k}_r"], 1e-15)
        if equed_ref in "CB":
            assert _close(c, g[f"svx{k}_c"], 1e-15)
        if 0 < int(info_ref) <= n:
            assert res["rcond"] == 0.0
            continue
        assert np.array_equal(ipiv, g[f"svx{k}_ipiv"])
        assert abs(res["rcond"] - rc_ref) <= 1e-10 * abs(rc_ref)
        for j in range(res["x"].shape[1]):                                # two correct solvers agree within their own FERR bounds
            tol = max(1e-9, 2.0 * (res["ferr"][j] + g[f"svx{k}_ferr"][j]))
            assert _close(res["x"][:, j], g[f"svx{k}_x"][:, j], tol), (k, j, tol)
        assert np.all(res["berr"] <= 4 * 2.0 ** -53 * (n + 1)) and np.all(g[f"svx{k}_berr"] <= 4 * 2.0 ** -53 * (n + 1))
        assert np.all(res["ferr"] <= 10 * g[f"svx{k}_ferr"]) and np.all(g[f"svx{k}_ferr"] <= 10 * res["ferr"])


# ------------------------------------------------------------------------------------------- blocked drivers at >= 1000^2
@pytest.fixture(scope="module")
def golden_large():
    return np.load(os.path.join(os.path.dirname(__file__), "golden", "netlib_golden_large.npz"))


def _sample(a):
    return np.ascontiguousarray(a[::7, ::7]), a.sum(axis=1)


@pytest.mark.parametrize("tag", ["sq", "tall", "wide"])
def test_blocked_dgetrf_large_matches_netlib(golden_large, tag):
    """ora_dgetrf (SRC/dgetrf.f:164-219, NB = 64, with its DLASWP / DTRSM / DGEMM) at >= 1000^2 against netlib's recursive DGETRF2 on
    the same DLARNV input: IPIV identical, factors to rounding."""
    g = golden_large
    m, n = (int(v) for v in g[f"lu_{tag}_shape"])
    a, seed = O.random_matrix(m, n, (1988, 1989, 1990, 1991))
    lu = a.copy(order="F")
    ipiv, info = O.dgetrf(lu)
    assert info == int(g[f"lu_{tag}_info"][0]) == 0
    assert np.array_equal(ipiv, g[f"lu_{tag}_ipiv"])
    smp, rs = _sample(lu)
    assert np.max(np.abs(smp - g[f"lu_{tag}_sample"])) < 1e-11 * np.max(np.abs(smp))
    assert np.max(np.abs(rs - g[f"lu_{tag}_rowsum"])) < 1e-10 * np.max(np.abs(rs))
    if tag == "sq":
        x, _ = O.random_matrix(n, 3, seed)
        for tr in "NT":
            sol = np.asfortranarray((a if tr == "N" else a.T) @ x)
            O.dgetrs(tr, lu, ipiv, sol)                                 # SRC/dgetrs.f:187-217
            ref = g[f"getrs_{tr}"]
            assert np.max(np.abs(sol - ref)) / np.max(np.abs(ref)) < 1e-9
            assert np.max(np.abs(sol - x)) / np.max(np.abs(x)) < 1e-8


@pytest.mark.parametrize("uplo", "LU")
def test_blocked_dpotrf_large_matches_netlib(golden_large, uplo):
    g = golden_large
    n = 1200
    s, _ = O.spd_matrix(n, (1988, 1989, 1990, 1991))
    f = s.copy(order="F")
    assert O.dpotrf(uplo, f) == int(g[f"po_{uplo}_info"][0]) == 0       # SRC/dpotrf.f:166-240, NB = 64
    tri = np.tril(f) if uplo == "L" else np.triu(f)
    smp, rs = _sample(tri)
    assert np.max(np.abs(smp - g[f"po_{uplo}_sample"])) < 1e-12 * np.max(np.abs(smp))
    assert np.max(np.abs(rs - g[f"po_{uplo}_rowsum"])) < 1e-11 * np.max(np.abs(rs))


def test_dlaswp_long_lists_match_netlib(golden_large):
    g = golden_large
    m, ncol = 1100, 40
    enc = np.asfortranarray(np.arange(m)[:, None] * 1000.0 + np.arange(ncol)[None, :])
    for k in range(int(g["n_swp"][0])):
        k1, k2, incx = (int(v) for v in g[f"swp{k}_args"])
        a = enc.copy(order="F")
        O.dlaswp(a, k1, k2, g[f"swp{k}_ipiv"], incx)                    # SRC/dlaswp.f:138-183 (32-column strips + remainder)
        assert np.array_equal(a[:, 0], g[f"swp{k}_col0"])
        assert np.array_equal(a, a[:, :1] + np.arange(ncol)[None, :])


# ------------------------------------------------------------------------------------------- DLATSQR
def test_dlatsqr_matches_netlib():
    """SRC/dlatsqr.f:185-290 with DTPQRT / DTPQRT2 / DTPRFB (L = 0): A (R and the reflector blocks) and every T block to rounding"""
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "netlib_golden_latsqr.npz"))
    for k, (m, n, mb, nb) in enumerate(g["cases"]):
        a, _ = O.random_matrix(int(m), int(n), (1988, 1989, 1990, 1991))
        t, info = O.dlatsqr(a, int(mb), int(nb))
        assert info == 0
        assert np.max(np.abs(a - g[f"a{k}"])) < 1e-12 * max(1.0, np.max(np.abs(g[f"a{k}"])))
        ref_t = g[f"t{k}"]
        n, nb = int(n), int(nb)
        for g0 in range(0, ref_t.shape[1], n):                         # one N-wide group of T blocks per row block (dlatsqr.f:100-104)
            for i in range(0, n, nb):                                  # only the upper triangle of every IB x IB block is defined
                w = min(nb, n - i)
                c0 = g0 + i
                assert np.max(np.abs(np.triu(t[:w, c0:c0 + w]) - np.triu(ref_t[:w, c0:c0 + w]))) < 1e-12
