#!/usr/bin/env python
"""Generate tests/golden/netlib_golden_geqrt.npz: DGEQRT / DGEMQRT outputs of netlib LAPACK 3.12.0 (scipy's OpenBLAS build,
same caveats as make_golden.py), to pin oracle/ora_dgeqrt3, ora_dgeqrt, ora_dgemqrt.  Run: python tests/golden/make_golden_geqrt.py"""
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
for tag, (m, n, nb) in {"tall": (90, 40, 8), "sq": (70, 70, 32), "wide": (30, 45, 7)}.items():
    a = rand(m, n, (1988, 1989, 1990, 1991))
    k = min(m, n)
    qr = a.copy(order="F")
    t = np.zeros((nb, k), order="F")
    work = np.zeros(nb * max(m, n))
    info = C.c_int(0)
    L.scipy_dgeqrt_(ci(m), ci(n), ci(nb), vp(qr), ci(m), vp(t), ci(nb), vp(work), C.byref(info))
    assert info.value == 0
    out[f"{tag}_a"], out[f"{tag}_qr"], out[f"{tag}_t"], out[f"{tag}_nb"] = a, qr, t, np.int32(nb)
    cl, cr = rand(m, 6, (3, 5, 7, 9)), rand(6, m, (13, 15, 17, 19))
    out[f"{tag}_cl"], out[f"{tag}_cr"] = cl, cr
    for side, c0 in (("L", cl), ("R", cr)):
        for trans in "NT":
            c = c0.copy(order="F")
            L.scipy_dgemqrt_(C.c_char_p(side.encode()), C.c_char_p(trans.encode()), ci(c.shape[0]), ci(c.shape[1]), ci(k), ci(nb),
                             vp(qr), ci(m), vp(t), ci(nb), vp(c), ci(c.shape[0]), vp(work), C.byref(info), C.c_size_t(1), C.c_size_t(1))
            assert info.value == 0
            out[f"{tag}_gemqrt_{side}{trans}"] = c
np.savez_compressed(os.path.join(HERE, "netlib_golden_geqrt.npz"), **out)
print(len(out), "arrays")
