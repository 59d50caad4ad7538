"""CUDA path (through the Fortran-77 ABI) on the reference test suite's matrix types: DLATB4/DLATMS matrices from
tests/golden/suite_matrices.npz, block sizes 1 / 3 / 20 / default like the NBVAL loop of TESTING/LIN/dchkge.f,
same gates as the reference's test programs (ratios < 30, INFO = IZERO) plus agreement with the oracle."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import oracle as O  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SUITE = np.load(os.path.join(ROOT, "tests", "golden", "suite_matrices.npz"))
GE = sorted(k for k in SUITE.files if k.startswith("ge_"))
PO = sorted(k for k in SUITE.files if k.startswith("po_"))
QR = sorted(k for k in SUITE.files if k.startswith("qr_"))


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


@pytest.mark.parametrize("nb", [1, 3, 20, 512])
def test_ge_types(lb, nb):
    L = lb.lib()
    L.lb200_set_getrf_params(nb, 0, 1)
    try:
        for key in GE:
            a = np.asfortranarray(SUITE[key])
            well = not key.endswith(("_t8", "_t9"))                      # types 8, 9 are ill-conditioned on purpose
            ref = a.copy(order="F")
            ipiv_ref, info_ref = O.dgetrf(ref)
            got = a.copy(order="F")
            ipiv, info = lb.f77.getrf(got)
            assert info == info_ref == 0, key
            assert O.dget01(a, got, ipiv) < O.THRESH, key
            if well:
                assert np.array_equal(ipiv, ipiv_ref), key
                assert rel(got, ref) < 1e-10, key
            if key.endswith("_t4"):
                m, n = a.shape
                mn = min(m, n)
                for izero in (1, mn, mn // 2 + 1):                        # types 5, 6, 7 (dchkge.f:328-347)
                    b = a.copy(order="F")
                    if izero == mn and izero < n:
                        b[:, izero - 1:] = 0.0
                    else:
                        b[:, izero - 1] = 0.0
                    ref = b.copy(order="F")
                    ipiv_ref, info_ref = O.dgetrf(ref)
                    got = b.copy(order="F")
                    ipiv, info = lb.f77.getrf(got)
                    assert info == info_ref == izero, (key, izero)
                    assert np.array_equal(ipiv, ipiv_ref), (key, izero)
                    assert O.dget01(b, got, ipiv) < O.THRESH
                # solve with the factors: DGETRS vs the oracle's solution (dchkge.f:420-450, DGET02 gate)
                if m == n:
                    x, _ = O.random_matrix(n, 2, (5, 6, 7, 9))
                    b0 = np.asfortranarray(a @ x)
                    for trans in "NT":
                        bb = np.asfortranarray(a @ x if trans == "N" else a.T @ x)
                        f = a.copy(order="F")
                        ipiv, _ = lb.f77.getrf(f)
                        sol = bb.copy(order="F")
                        assert lb.f77.getrs(trans, f, ipiv, sol) == 0
                        assert O.dget02(trans, a, sol, bb) < O.THRESH
                        assert rel(sol, x) < 1e-10
                    del b0
    finally:
        L.lb200_set_getrf_params(512, 0, 1)


@pytest.mark.parametrize("nb", [1, 3, 20, 512])
def test_po_types(lb, nb):
    L = lb.lib()
    L.lb200_set_potrf_params(nb, 1)
    try:
        for key in PO:
            a = np.asfortranarray(SUITE[key])
            n = a.shape[0]
            for uplo in "UL":
                ref = a.copy(order="F")
                assert O.dpotrf(uplo, ref) == 0
                got = a.copy(order="F")
                assert lb.f77.potrf(uplo, got) == 0, (key, uplo)
                assert O.dpot01(uplo, a, got) < O.THRESH, (key, uplo)
                tri = np.tril if uplo == "L" else np.triu
                if not key.endswith(("_t6", "_t7")):
                    assert rel(tri(got), tri(ref)) < 1e-10, (key, uplo)
                if key.endswith("_t2"):
                    for izero in (1, n, n // 2 + 1):                      # types 3, 4, 5 (dchkpo.f:313-344)
                        b = a.copy(order="F")
                        b[izero - 1, :] = 0.0
                        b[:, izero - 1] = 0.0
                        assert lb.f77.potrf(uplo, b.copy(order="F")) == izero, (key, uplo, izero)
    finally:
        L.lb200_set_potrf_params(512, 1)


@pytest.mark.parametrize("nb", [1, 3, 20, 256])
def test_qr_types(lb, nb):
    L = lb.lib()
    L.lb200_set_geqrf_params(nb, 1)
    try:
        for key in QR:
            a = np.asfortranarray(SUITE[key])
            got = a.copy(order="F")
            tau, info, _ = lb.f77.geqrf(got)
            assert info == 0, key
            r1, r2 = O.dqrt01(a, got, tau)                               # ||R - Q'A|| and ||I - Q'Q|| ratios
            assert r1 < O.THRESH and r2 < O.THRESH, (key, r1, r2)
            ref = a.copy(order="F")
            tau_ref, _, _ = O.dgeqrf(ref)
            if not key.endswith(("_t5", "_t6")):
                assert rel(np.triu(got), np.triu(ref)) < 1e-10, key
                assert rel(tau, tau_ref) < 1e-10, key
    finally:
        L.lb200_set_geqrf_params(256, 1)


@pytest.mark.parametrize("nb", [1, 3, 20])
def test_tiny_shape_sweep(lb, nb):
    """The reference test programs sweep M, N in {0, 1, 2, 3, 5, 10, 50} with NB in {1, 3, 20} (TESTING/dtest.in); every
    combination through LU, QR (+ DORGQR), LQ and, for square sizes, Cholesky, against the oracle."""
    L = lb.lib()
    L.lb200_set_getrf_params(nb, 0, 1)
    L.lb200_set_potrf_params(nb, 1)
    L.lb200_set_geqrf_params(nb, 1)
    try:
        for m in (0, 1, 2, 3, 5, 10, 50):
            for n in (0, 1, 2, 3, 5, 10, 50):
                a, _ = O.random_matrix(m, n, (1988, 1989, 1990, 1991))
                k = min(m, n)
                # LU
                ref = a.copy(order="F")
                ipiv_ref, info_ref = O.dgetrf2(ref)
                got = a.copy(order="F")
                ipiv, info = lb.f77.getrf(got)
                assert info == info_ref, (m, n)
                if k > 0:
                    assert np.array_equal(ipiv[:k], ipiv_ref[:k]), (m, n)
                    assert rel(got, ref) < 1e-11, (m, n)
                # QR and the first min(m, n) columns of Q
                ref = a.copy(order="F")
                tau_ref, _, _ = O.dgeqrf(ref)
                got = a.copy(order="F")
                tau, info, _ = lb.f77.geqrf(got)
                assert info == 0
                if k > 0:
                    assert rel(got, ref) < 1e-11 and rel(tau[:k], tau_ref[:k]) < 1e-11, (m, n)
                    q = np.asfortranarray(got[:, :k])
                    assert lb.f77.orgqr(q, tau[:k]) == 0
                    assert np.max(np.abs(q.T @ q - np.eye(k))) < 1e-12 * max(m, 10), (m, n)
                # LQ
                if k > 0:
                    ref = a.copy(order="F")
                    tau_ref, _ = O.dgelq2(ref)
                    got = a.copy(order="F")
                    tau, info = lb.f77.gelqf(got)
                    assert info == 0 and rel(got, ref) < 1e-11 and rel(tau, tau_ref) < 1e-11, (m, n)
                # Cholesky
                if m == n and n > 0:
                    s, _ = O.spd_matrix(n, (1988, 1989, 1990, 1991))
                    for uplo in "UL":
                        ref = s.copy(order="F")
                        assert O.dpotrf(uplo, ref) == 0
                        got = s.copy(order="F")
                        assert lb.f77.potrf(uplo, got) == 0
                        tri = np.triu if uplo == "U" else np.tril
                        assert rel(tri(got), tri(ref)) < 1e-12, (n, uplo)
    finally:
        L.lb200_set_getrf_params(512, 0, 1)
        L.lb200_set_potrf_params(512, 1)
        L.lb200_set_geqrf_params(256, 1)
