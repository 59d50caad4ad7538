#!/usr/bin/env python
"""Generate tests/golden/suite_matrices.npz -- the reference test suite's matrix types at small sizes.

TESTING/LIN/dchkge.f:300-316, dchkpo.f:240-262 and dchkqr.f build their inputs with DLATB4 (parameter table,
TESTING/LIN/dlatb4.f:170-176 constants, :196-224 QR, :437-477 GE, :540-574 PO) + DLATMS, seed 1988..1991.  The
Fortran generator cannot be compiled here; the gfortran-compiled netlib DLATMS inside scipy's OpenBLAS is called
instead (same routine, same DLARNV stream).  The zero-column / zero-row types (GE 5-7, PO 3-5) are derived from
the stored type-4 / type-2 matrices by the tests, as the test programs do.

Run:  python tests/golden/make_suite_matrices.py
"""
import ctypes as C
import glob
import os

import numpy as np
import scipy

HERE = os.path.dirname(os.path.abspath(__file__))
_so = glob.glob(os.path.join(os.path.dirname(scipy.__file__), "..", "scipy.libs", "libscipy_openblas*.so"))[0]
L = C.CDLL(_so)
dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int)

EPS = 2.0 ** -52                     # DLAMCH('Precision') = eps*base
SFMIN = 2.2250738585072014e-308
BADC2 = 0.1 / EPS
BADC1 = np.sqrt(BADC2)
SMALL = 0.25 * (SFMIN / EPS)
LARGE = 1.0 / SMALL


def dlatms(m, n, sym, cond, anorm, kl, ku, seed):
    a = np.zeros((m, n), order="F")
    d = np.zeros(min(m, n))
    work = np.zeros(3 * max(m, n, 1))
    iseed = np.array(seed, dtype=np.int32)
    info = C.c_int(0)
    ci = lambda v: C.byref(C.c_int(v))
    cd = lambda v: C.byref(C.c_double(v))
    L.scipy_dlatms_(ci(m), ci(n), C.c_char_p(b"S"), iseed.ctypes.data_as(ip), C.c_char_p(sym.encode()),
                    d.ctypes.data_as(dp), ci(3), cd(cond), cd(anorm), ci(kl), ci(ku), C.c_char_p(b"N"),
                    a.ctypes.data_as(dp), ci(max(1, m)), work.ctypes.data_as(dp), C.byref(info),
                    C.c_size_t(1), C.c_size_t(1), C.c_size_t(1))
    assert info.value == 0, info.value
    return a, tuple(int(x) for x in iseed)


out = {}
seed = (1988, 1989, 1990, 1991)
# ---- GE (dlatb4.f:437-477): 1 diagonal, 2 upper, 3 lower, 4 random, 8 cond BADC1, 9 cond BADC2, 10 tiny, 11 huge
for (m, n) in ((40, 40), (40, 10), (10, 40), (66, 66)):
    for imat in (1, 2, 3, 4, 8, 9, 10, 11):
        kl = 0 if imat in (1, 2) else max(m - 1, 0)
        ku = 0 if imat in (1, 3) else max(n - 1, 0)
        cond = BADC1 if imat == 8 else BADC2 if imat == 9 else 2.0
        anorm = SMALL if imat == 10 else LARGE if imat == 11 else 1.0
        a, seed = dlatms(m, n, "N", cond, anorm, kl, ku, seed)
        out[f"ge_{m}x{n}_t{imat}"] = a
# ---- PO (dlatb4.f:540-574): 1 diagonal, 2 random, 6 BADC1, 7 BADC2, 8 tiny, 9 huge
for n in (40, 66):
    for imat in (1, 2, 6, 7, 8, 9):
        kl = 0 if imat == 1 else max(n - 1, 0)
        cond = BADC1 if imat == 6 else BADC2 if imat == 7 else 2.0
        anorm = SMALL if imat == 8 else LARGE if imat == 9 else 1.0
        a, seed = dlatms(n, n, "P", cond, anorm, kl, kl, seed)
        out[f"po_{n}_t{imat}"] = a
# ---- QR (dlatb4.f:196-224): 1 diagonal, 2 upper, 3 lower, 4 random, 5 BADC1, 6 BADC2, 7 tiny, 8 huge
for (m, n) in ((40, 40), (66, 34), (34, 66)):
    for imat in range(1, 9):
        kl = 0 if imat in (1, 2) else max(m - 1, 0)
        ku = 0 if imat in (1, 3) else max(n - 1, 0)
        cond = BADC1 if imat == 5 else BADC2 if imat == 6 else 2.0
        anorm = SMALL if imat == 7 else LARGE if imat == 8 else 1.0
        a, seed = dlatms(m, n, "N", cond, anorm, kl, ku, seed)
        out[f"qr_{m}x{n}_t{imat}"] = a
np.savez_compressed(os.path.join(HERE, "suite_matrices.npz"), **out)
print(len(out), "matrices,", sum(v.size for v in out.values()), "doubles")
