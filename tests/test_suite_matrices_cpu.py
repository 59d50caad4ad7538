"""The oracle on the reference test suite's matrix types (tests/golden/suite_matrices.npz, DLATB4 + DLATMS):
the same gates the reference's own test programs apply -- residual ratios below THRESH = 30
(TESTING/LIN/dchkge.f:369-383, dchkpo.f:362-376, dchkqr.f) and INFO = IZERO for the singular types
(dchkge.f:328-347, dchkpo.f:313-344) -- for several block sizes (dchkge.f NBVAL loop)."""
import os

import numpy as np
import pytest

from oracle import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SUITE = np.load(os.path.join(ROOT, "tests", "golden", "suite_matrices.npz"))
GE = sorted(k for k in SUITE.files if k.startswith("ge_"))
PO = sorted(k for k in SUITE.files if k.startswith("po_"))
QR = sorted(k for k in SUITE.files if k.startswith("qr_"))


@pytest.mark.parametrize("key", GE)
@pytest.mark.parametrize("nb", [1, 3, 20])
def test_oracle_ge(key, nb):
    a = np.asfortranarray(SUITE[key])
    O.set_nb(getrf=nb)
    try:
        f = a.copy(order="F")
        ipiv, info = O.dgetrf(f)
        assert info == 0
        if info == 0:
            assert O.dget01(a, f, ipiv) < O.THRESH
        if key.endswith("_t4"):
            m, n = a.shape
            mn = min(m, n)
            for izero in (1, mn, mn // 2 + 1):                            # types 5, 6, 7
                b = a.copy(order="F")
                if izero == mn and izero < n:
                    b[:, izero - 1:] = 0.0
                else:
                    b[:, izero - 1] = 0.0
                f = b.copy(order="F")
                ipiv, info = O.dgetrf(f)
                assert info == izero, (key, izero)
                assert O.dget01(b, f, ipiv) < O.THRESH
    finally:
        O.set_nb()


@pytest.mark.parametrize("key", PO)
@pytest.mark.parametrize("uplo", "UL")
@pytest.mark.parametrize("nb", [1, 3, 20])
def test_oracle_po(key, uplo, nb):
    a = np.asfortranarray(SUITE[key])
    n = a.shape[0]
    O.set_nb(potrf=nb)
    try:
        f = a.copy(order="F")
        assert O.dpotrf(uplo, f) == 0
        assert O.dpot01(uplo, a, f) < O.THRESH
        if key.endswith("_t2"):
            for izero in (1, n, n // 2 + 1):                              # types 3, 4, 5
                b = a.copy(order="F")
                b[izero - 1, :] = 0.0
                b[:, izero - 1] = 0.0
                assert O.dpotrf(uplo, b.copy(order="F")) == izero
    finally:
        O.set_nb()


@pytest.mark.parametrize("key", QR)
@pytest.mark.parametrize("nb", [1, 3, 20])
def test_oracle_qr(key, nb):
    a = np.asfortranarray(SUITE[key])
    O.set_nb(geqrf=nb, nx=1)
    try:
        f = a.copy(order="F")
        tau, info, _ = O.dgeqrf(f)
        assert info == 0
        r1, r2 = O.dqrt01(a, f, tau)
        assert r1 < O.THRESH and r2 < O.THRESH
    finally:
        O.set_nb()
