"""GPU parity for the condition-estimation / expert-driver layer: DLATRS, DGECON, DGEEQU, DGESVX and DGEQRT3
through the Fortran-77 ABI, against the committed netlib golden vectors (tests/golden/make_golden_gecon.py) and the oracle."""
import os

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


@pytest.fixture(scope="module")
def g():
    return np.load(os.path.join(os.path.dirname(__file__), "golden", "netlib_golden_gecon.npz"))


def close(x, y, tol):
    x, y = np.asarray(x, dtype=float), np.asarray(y, dtype=float)
    fin = np.isfinite(y)
    if not np.array_equal(np.isfinite(x), fin):
        return False
    if not fin.any():
        return True
    return float(np.max(np.abs(x[fin] - y[fin]))) / max(1e-300, float(np.max(np.abs(y[fin])))) < tol


def test_dlatrs_vs_netlib_golden(lb, g):
    """plain DTRSV branch and the scaled Level-1 branch (dlatrs.f:568-840): SCALE, x, CNORM; NORMIN='Y' reuses the norms"""
    scaled = 0
    for k in range(int(g["n_trs"][0])):
        uplo, trans, diag = (chr(v) for v in g[f"trs{k}_meta"])
        a, x = np.asfortranarray(g[f"trs{k}_a"]), g[f"trs{k}_x"].copy()
        cn = np.zeros(len(x))
        scale, info = lb.f77.dlatrs(uplo, trans, diag, "N", a, x, cn)
        s_ref, s_ref2 = g[f"trs{k}_scale"]
        assert info == 0
        assert scale == s_ref or abs(scale - s_ref) <= 1e-12 * abs(s_ref), (k, uplo, trans, diag, scale, s_ref)
        assert close(x, g[f"trs{k}_xo"], 1e-10), (k, uplo, trans, diag)
        assert close(cn, g[f"trs{k}_cn"], 1e-13)
        scaled += scale != 1.0
        x2 = g[f"trs{k}_x"].copy()
        scale2, _ = lb.f77.dlatrs(uplo, trans, diag, "Y", a, x2, cn)
        assert scale2 == s_ref2 or abs(scale2 - s_ref2) <= 1e-12 * abs(s_ref2)
        assert close(x2, g[f"trs{k}_xo2"], 1e-10)
    assert scaled >= 8


def test_dgecon_vs_netlib_golden_and_oracle(lb, g):
    for k in range(int(g["n_con"][0])):
        lu = np.asfortranarray(g[f"con{k}_lu"])
        for norm in "1I":
            anorm, rc_ref, info_ref = g[f"con{k}_{norm}"]
            rc, info = lb.f77.dgecon(norm, lu, anorm)
            assert info == int(info_ref)
            tol = 1e-10 if rc_ref > 1e-12 else 1e-5
            assert rc == rc_ref or abs(rc - rc_ref) <= tol * abs(rc_ref), (k, norm, rc, rc_ref)
    # larger, through our own factorization: same estimate as the oracle's DGECON on the oracle's factors
    for n in (300, 1100):
        a, _ = O.random_matrix(n, n, SEED)
        lu_ref = a.copy(order="F")
        O.dgetrf(lu_ref)
        lu = a.copy(order="F")
        lb.f77.getrf(lu)
        for norm in "1I":
            anorm = float(np.linalg.norm(a, 1 if norm == "1" else np.inf))
            rc_ref, _ = O.dgecon(norm, lu_ref, anorm)
            rc, info = lb.f77.dgecon(norm, lu, anorm)
            assert info == 0 and abs(rc - rc_ref) <= 1e-8 * rc_ref, (n, norm, rc, rc_ref)
            true_rc = 1.0 / (anorm * np.linalg.norm(np.linalg.inv(a), 1 if norm == "1" else np.inf))
            assert true_rc * 0.999 <= rc <= 10 * true_rc          # the estimator never overestimates ||inv(A)|| (dlacn2.f)


def test_dgecon_error_exits_and_quick_returns(lb):
    a = np.asfortranarray(np.eye(3))
    assert lb.f77.dgecon("X", a, 1.0)[1] == -1
    assert lb.f77.dgecon("1", a, 1.0, n=-1)[1] == -2
    assert lb.f77.dgecon("1", a, 1.0, n=3, lda=2)[1] == -4
    assert lb.f77.dgecon("1", a, -1.0)[1] == -5
    assert lb.f77.dgecon("1", a, 0.0) == (0.0, 0)                  # dgecon.f:199-200
    rc, info = lb.f77.dgecon("1", a, float("nan"))
    assert np.isnan(rc) and info == -5                             # dgecon.f:201-203
    assert lb.f77.dgecon("1", a, float("inf"))[1] == -5
    z = np.asfortranarray(np.diag([1.0, 0.0, 2.0]))               # exactly singular U: RCOND = 0 (inv-norm estimate overflows)
    rc, info = lb.f77.dgecon("1", z, 2.0)
    assert rc == 0.0


def test_dgeequ_vs_oracle(lb):
    rng = np.random.default_rng(3)
    for (m, n) in ((7, 5), (100, 130), (400, 300)):
        a = np.asfortranarray(rng.uniform(-1, 1, (m, n)) * (10.0 ** rng.uniform(-6, 6, m))[:, None])
        want = O.dgeequ(a)
        got = lb.f77.dgeequ(a)
        assert got[5] == want[5] == 0
        assert close(got[0], want[0], 1e-15) and close(got[1], want[1], 1e-15)
        assert got[2:5] == pytest.approx(want[2:5], rel=1e-15)
    a[3, :] = 0.0
    assert lb.f77.dgeequ(a)[5] == O.dgeequ(a)[5] == 4              # zero row: INFO = i (dgeequ.f:224-229)


def test_dgesvx_vs_netlib_golden(lb, g):
    for k in range(int(g["n_svx"][0])):
        fact, trans, equed_ref, info_ref = g[f"svx{k}_meta"]
        fact, trans, equed_ref = chr(fact), chr(trans), chr(equed_ref)
        a, b = np.asfortranarray(g[f"svx{k}_a"].copy()), np.asfortranarray(g[f"svx{k}_b"].copy())
        n = a.shape[0]
        af, ipiv, r, c = np.zeros((n, n), order="F"), np.zeros(n, dtype=np.int32), np.zeros(n), np.zeros(n)
        res = lb.f77.dgesvx(fact, trans, a, af, ipiv, "N", r, c, b)
        assert res["info"] == int(info_ref) and res["equed"] == equed_ref, (k, res["info"], info_ref, res["equed"], equed_ref)
        assert close(a, g[f"svx{k}_a_out"], 1e-14) and close(b, g[f"svx{k}_b_out"], 1e-14)      # A, B come back equilibrated
        rc_ref, rpv_ref = g[f"svx{k}_scal"]
        assert abs(res["rpvgrw"] - rpv_ref) <= 1e-11 * abs(rpv_ref)
        if equed_ref in "RB":
            assert close(r, g[f"svx{k}_r"], 1e-15)
        if equed_ref in "CB":
            assert close(c, g[f"svx{k}_c"], 1e-15)
        if 0 < int(info_ref) <= n:
            assert res["rcond"] == 0.0
            continue
        assert np.array_equal(ipiv, g[f"svx{k}_ipiv"])
        assert close(af, g[f"svx{k}_af"], 1e-11)
        assert abs(res["rcond"] - rc_ref) <= 1e-9 * abs(rc_ref)
        for j in range(res["x"].shape[1]):
            tol = max(1e-9, 2.0 * (res["ferr"][j] + g[f"svx{k}_ferr"][j]))
            assert close(res["x"][:, j], g[f"svx{k}_x"][:, j], tol), (k, j)
        assert np.all(res["berr"] <= 4 * 2.0 ** -53 * (n + 1))
        assert np.all(res["ferr"] <= 10 * g[f"svx{k}_ferr"]) and np.all(g[f"svx{k}_ferr"] <= 10 * res["ferr"])


def test_dgesvx_fact_f_and_large(lb):
    """FACT='F' reuses the caller's factors / scalings; a larger system against the oracle's DGESVX"""
    n, nrhs = 700, 2
    a0, seed = O.random_matrix(n, n, SEED)
    a0 *= (10.0 ** np.linspace(-5, 5, n))[:, None]
    xact, _ = O.random_matrix(n, nrhs, seed)
    b0 = np.asfortranarray(a0 @ xact)
    a, b = a0.copy(order="F"), b0.copy(order="F")
    af, ipiv, r, c = np.zeros((n, n), order="F"), np.zeros(n, dtype=np.int32), np.zeros(n), np.zeros(n)
    res = lb.f77.dgesvx("E", "N", a, af, ipiv, "N", r, c, b)
    a2, b2 = a0.copy(order="F"), b0.copy(order="F")
    af2, ipiv2, r2, c2 = np.zeros((n, n), order="F"), np.zeros(n, dtype=np.int32), np.zeros(n), np.zeros(n)
    ref = O.dgesvx("E", "N", a2, af2, ipiv2, "N", r2, c2, b2)
    assert res["info"] == ref["info"] == 0 and res["equed"] == ref["equed"] and res["equed"] in "RB"
    assert np.array_equal(ipiv, ipiv2)
    assert abs(res["rcond"] - ref["rcond"]) <= 1e-7 * ref["rcond"]
    assert np.max(np.abs(res["x"] - xact)) / np.max(np.abs(xact)) <= np.max(res["ferr"])      # FERR really bounds the error
    assert close(res["x"], ref["x"], max(1e-9, 2 * (np.max(res["ferr"]) + np.max(ref["ferr"]))))
    # FACT = 'F': hand the factors and scalings back in with a new right-hand side (already equilibrated A)
    b3 = np.asfortranarray(a0 @ xact[:, :1] * 2.0)
    res3 = lb.f77.dgesvx("F", "N", a, af, ipiv, res["equed"], r, c, b3)
    assert res3["info"] == 0 and res3["equed"] == res["equed"]
    assert np.max(np.abs(res3["x"] - 2.0 * xact[:, :1])) / 2.0 <= max(1e-9, 4 * res3["ferr"][0])


def test_dgeqrt3_vs_oracle(lb):
    for (m, n) in ((1, 1), (9, 4), (130, 64), (400, 150)):
        a, _ = O.random_matrix(m, n, SEED)
        ref = a.copy(order="F")
        t_ref, info_ref = O.dgeqrt(ref, n)                           # one block of width n == DGEQRT3 (dgeqrt.f:196-199)
        got = a.copy(order="F")
        t = np.full((n + 1, n), 4.5e77, order="F")
        assert lb.f77.dgeqrt3(m, n, got, m, t, n + 1) == info_ref == 0
        assert np.max(np.abs(got - ref)) < 1e-11 * max(1.0, np.max(np.abs(ref)))
        assert np.max(np.abs(np.triu(t[:n]) - np.triu(t_ref[:n, :n]))) < 1e-11
        assert np.all(t[n] == 4.5e77) and np.all(t[:n][np.tril_indices(n, -1)] == 4.5e77)       # below the diagonal: not used
    assert lb.f77.dgeqrt3(3, 5, np.zeros((3, 5), order="F"), 3, np.zeros((5, 5), order="F"), 5) == -1   # M < N (dgeqrt3.f:162)


def test_dlatsqr_vs_netlib_golden(lb):
    """DLATSQR (SRC/dlatsqr.f:185-290): R, the reflector blocks and every T block against netlib's DLATSQR on the same DLARNV input;
    Q^T A = R checked through the stored blocks for the largest case."""
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "netlib_golden_latsqr.npz"))
    for k, (m, n, mb, nb) in enumerate(g["cases"]):
        m, n, mb, nb = int(m), int(n), int(mb), int(nb)
        a, _ = O.random_matrix(m, n, SEED)
        tc = O.latsqr_tcols(m, n, mb)
        t = np.zeros((nb, tc), order="F")
        wq = np.zeros(1)
        assert lb.f77.dlatsqr(m, n, mb, nb, a, m, t, nb, wq, -1) == 0 and wq[0] == n * nb          # workspace query (dlatsqr.f:213-217)
        work = np.zeros(n * nb)
        assert lb.f77.dlatsqr(m, n, mb, nb, a, m, t, nb, work, n * nb) == 0
        ref_a, ref_t = g[f"a{k}"], g[f"t{k}"]
        assert np.max(np.abs(a - ref_a)) < 1e-11 * max(1.0, np.max(np.abs(ref_a))), (m, n, mb, nb)
        for g0 in range(0, tc, n):
            for i in range(0, n, nb):
                w = min(nb, n - i)
                c0 = g0 + i
                assert np.max(np.abs(np.triu(t[:w, c0:c0 + w]) - np.triu(ref_t[:w, c0:c0 + w]))) < 1e-11, (m, n, mb, nb, c0)
    # error exits (dlatsqr.f:222-238)
    a = np.zeros((10, 4), order="F"); t = np.zeros((2, 8), order="F"); w = np.zeros(8)
    assert lb.f77.dlatsqr(3, 4, 6, 2, a, 10, t, 2, w, 8) == -2        # M < N
    assert lb.f77.dlatsqr(10, 4, 0, 2, a, 10, t, 2, w, 8) == -3
    assert lb.f77.dlatsqr(10, 4, 6, 5, a, 10, t, 5, w, 20) == -4      # NB > N
    assert lb.f77.dlatsqr(10, 4, 6, 2, a, 9, t, 2, w, 8) == -6
    assert lb.f77.dlatsqr(10, 4, 6, 2, a, 10, t, 1, w, 8) == -8
    assert lb.f77.dlatsqr(10, 4, 6, 2, a, 10, t, 2, w, 7) == -10
