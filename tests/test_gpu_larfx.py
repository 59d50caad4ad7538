"""GPU parity: DLARFT / DLARFB for ALL DIRECT x STOREV storage schemes (SRC/dlarft.f:100-150, SRC/dlarfb.f:150-190) against the
definition H = H(1)...H(k) (forward) / H(k)...H(1) (backward), H(i) = I - tau(i) v(i) v(i)^T, evaluated densely in numpy.  The
reference's DORGLQ/DORMLQ use ('F','R'), DGERQF/DORMRQ ('B','R'), DGEQLF/DORMQL ('B','C'): with the library preloaded those callers
reach these symbols, so every combination must be served."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

SENT = 3.0e33      # entries of V that the storage scheme defines implicitly (unit / zero) must never be read


@pytest.fixture(scope="module")
def lb():
    import lapack_b200
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    lapack_b200.lib().lb200_set_xerbla_mode(2)
    return lapack_b200


def make_reflectors(rng, direct, storev, nv, k):
    """returns (V as stored, dense reflector matrix Vc nv x k, tau)"""
    vc = rng.uniform(-1, 1, (nv, k))
    stored = np.full((nv, k), SENT)
    for j in range(k):
        u = j if direct == "F" else nv - k + j
        if direct == "F":
            vc[:u, j] = 0.0
            stored[u + 1:, j] = vc[u + 1:, j]
        else:
            vc[u + 1:, j] = 0.0
            stored[:u, j] = vc[:u, j]
        vc[u, j] = 1.0
    tau = rng.uniform(0.5, 1.0, k) * 2.0 / np.sum(vc * vc, axis=0)   # ||H(i)|| <= 1 (tau = 2/v'v is the orthogonal case)
    tau[k // 2] = 0.0                                              # H(i) = I (dlarft.f: tau(i) = 0 column of T is zero)
    v = np.asfortranarray(stored if storev == "C" else stored.T)
    return v, vc, tau


def dense_h(direct, vc, tau):
    nv, k = vc.shape
    h = np.eye(nv)
    for j in range(k):
        hj = np.eye(nv) - tau[j] * np.outer(vc[:, j], vc[:, j])
        h = h @ hj if direct == "F" else hj @ h
    return h


@pytest.mark.parametrize("direct", "FB")
@pytest.mark.parametrize("storev", "CR")
def test_dlarft_all_schemes(lb, direct, storev):
    rng = np.random.default_rng(31)
    for (nv, k) in ((7, 3), (40, 40), (130, 33), (300, 70), (65, 64)):
        v, vc, tau = make_reflectors(rng, direct, storev, nv, k)
        t = np.full((k + 1, k), -7.0e22, order="F")
        lb.f77.dlarft(direct, storev, nv, k, v, v.shape[0], tau, t, k + 1)
        tt = t[:k]
        if direct == "F":
            assert np.all(tt[np.tril_indices(k, -1)] == -7.0e22)       # only the upper triangle is written (dlarft.f:41-44)
            tri = np.triu(tt)
        else:
            assert np.all(tt[np.triu_indices(k, 1)] == -7.0e22)
            tri = np.tril(tt)
        assert np.all(t[k] == -7.0e22)
        h = dense_h(direct, vc, tau)
        assert np.max(np.abs(np.eye(nv) - vc @ tri @ vc.T - h)) < 1e-12 * k, (direct, storev, nv, k)


@pytest.mark.parametrize("direct", "FB")
@pytest.mark.parametrize("storev", "CR")
@pytest.mark.parametrize("side", "LR")
@pytest.mark.parametrize("trans", "NT")
def test_dlarfb_all_schemes(lb, direct, storev, side, trans):
    rng = np.random.default_rng(32)
    for (m, n, k) in ((9, 5, 3), (70, 45, 33), (200, 130, 64)):
        nv = m if side == "L" else n
        if k > nv:
            continue
        v, vc, tau = make_reflectors(rng, direct, storev, nv, k)
        t = np.zeros((k, k), order="F")
        lb.f77.dlarft(direct, storev, nv, k, v, v.shape[0], tau, t, k)
        t += SENT * (np.tril(np.ones((k, k)), -1) if direct == "F" else np.triu(np.ones((k, k)), 1))   # other triangle: not referenced
        c = np.asfortranarray(rng.uniform(-1, 1, (m, n)))
        h = dense_h(direct, vc, tau)
        hh = h if trans == "N" else h.T
        want = hh @ c if side == "L" else c @ hh
        ldc = m + 3
        got = np.full((ldc, n), 5.5e11, order="F")
        got[:m] = c
        work = np.zeros((max(m, n), k), order="F")
        lb.f77.dlarfb(side, trans, direct, storev, m, n, k, v, v.shape[0], t, k, got, ldc, work, max(m, n))
        assert np.all(got[m:] == 5.5e11)
        assert np.max(np.abs(got[:m] - want)) < 1e-12 * k, (direct, storev, side, trans, m, n, k)
