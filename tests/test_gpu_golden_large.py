"""GPU path against the committed netlib golden vectors at >= 1000^2 (tests/golden/make_golden_large.py): DGETRF / DGETRS / DPOTRF /
DLASWP through the Fortran-77 ABI.  IPIV and the interchange results must be identical, factors agree to rounding."""
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
    return np.load(os.path.join(os.path.dirname(__file__), "golden", "netlib_golden_large.npz"))


def sample(a):
    return np.ascontiguousarray(a[::7, ::7]), a.sum(axis=1)


@pytest.mark.parametrize("tag", ["sq", "tall", "wide"])
def test_dgetrf_dgetrs_vs_netlib_golden(lb, g, tag):
    m, n = (int(v) for v in g[f"lu_{tag}_shape"])
    a, seed = O.random_matrix(m, n, SEED)
    lu = a.copy(order="F")
    ipiv, info = lb.f77.getrf(lu)
    assert info == 0 and np.array_equal(ipiv, g[f"lu_{tag}_ipiv"])           # IPIV bit-exact with netlib's DGETRF2
    smp, rs = sample(lu)
    assert np.max(np.abs(smp - g[f"lu_{tag}_sample"])) < 1e-11 * np.max(np.abs(smp))
    assert np.max(np.abs(rs - g[f"lu_{tag}_rowsum"])) < 1e-10 * np.max(np.abs(rs))
    if tag == "sq":
        x, _ = O.random_matrix(n, 3, seed)
        for tr in "NT":
            sol = np.asfortranarray((a if tr == "N" else a.T) @ x)
            assert lb.f77.getrs(tr, lu, ipiv, sol) == 0
            ref = g[f"getrs_{tr}"]
            assert np.max(np.abs(sol - ref)) / np.max(np.abs(ref)) < 1e-9


@pytest.mark.parametrize("uplo", "LU")
def test_dpotrf_vs_netlib_golden(lb, g, uplo):
    n = 1200
    s, _ = O.spd_matrix(n, SEED)
    f = s.copy(order="F")
    assert lb.f77.potrf(uplo, f) == 0
    tri = np.tril(f) if uplo == "L" else np.triu(f)
    smp, rs = sample(tri)
    assert np.max(np.abs(smp - g[f"po_{uplo}_sample"])) < 1e-12 * np.max(np.abs(smp))
    assert np.max(np.abs(rs - g[f"po_{uplo}_rowsum"])) < 1e-11 * np.max(np.abs(rs))


def test_dlaswp_vs_netlib_golden(lb, g):
    m, ncol = 1100, 40
    enc = np.asfortranarray(np.arange(m)[:, None] * 1000.0 + np.arange(ncol)[None, :])
    for k in range(int(g["n_swp"][0])):
        k1, k2, incx = (int(v) for v in g[f"swp{k}_args"])
        a = enc.copy(order="F")
        lb.f77.dlaswp(ncol, a, m, k1, k2, g[f"swp{k}_ipiv"], incx)
        assert np.array_equal(a[:, 0], g[f"swp{k}_col0"]), (k1, k2, incx)
        assert np.array_equal(a, a[:, :1] + np.arange(ncol)[None, :])
