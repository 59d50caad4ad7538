#!/usr/bin/env python
"""Generate tests/golden/netlib_golden_getri.npz: DGETRI / DTRTRI outputs of the netlib LAPACK 3.12.0 drivers inside
scipy's OpenBLAS on DGETRF2 factors (same source of truth and caveats as make_golden.py), to pin oracle/ora_dgetri and
ora_dtrtri.   Run:  python tests/golden/make_golden_getri.py"""
import ctypes as C
import glob
import os

import numpy as np
import scipy

HERE = os.path.dirname(os.path.abspath(__file__))
_so = glob.glob(os.path.join(os.path.dirname(scipy.__file__), "..", "scipy.libs", "libscipy_openblas*.so"))[0]
L = C.CDLL(_so)
vp = lambda a: a.ctypes.data_as(C.c_void_p)
ci = lambda v: C.byref(C.c_int(v))


def rand(m, n, seed):
    s = np.array(seed, dtype=np.int32)
    x = np.empty(m * n)
    L.scipy_dlarnv_(ci(2), vp(s), ci(m * n), vp(x))
    return np.asfortranarray(x.reshape((n, m)).T)


out = {}
for n in (7, 70):
    a = rand(n, n, (1988, 1989, 1990, 1991))
    lu = a.copy(order="F")
    ipiv = np.zeros(n, dtype=np.int32)
    info = C.c_int(0)
    L.scipy_dgetrf2_(ci(n), ci(n), vp(lu), ci(n), vp(ipiv), C.byref(info))
    inv = lu.copy(order="F")
    work = np.zeros(64 * n)
    L.scipy_dgetri_(ci(n), vp(inv), ci(n), vp(ipiv), vp(work), ci(len(work)), C.byref(info))
    assert info.value == 0
    out[f"a{n}"], out[f"lu{n}"], out[f"ipiv{n}"], out[f"inv{n}"] = a, lu, ipiv, inv
    for uplo in "UL":
        t = (np.triu(a) if uplo == "U" else np.tril(a)) + 3.0 * np.eye(n)
        ti = np.asfortranarray(t.copy())
        L.scipy_dtrtri_(C.c_char_p(uplo.encode()), C.c_char_p(b"N"), ci(n), vp(ti), ci(n), C.byref(info), C.c_size_t(1), C.c_size_t(1))
        assert info.value == 0
        out[f"tri{uplo}{n}"], out[f"triinv{uplo}{n}"] = np.asfortranarray(t), ti
np.savez_compressed(os.path.join(HERE, "netlib_golden_getri.npz"), **out)
print(sorted(out))
