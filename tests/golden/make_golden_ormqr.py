#!/usr/bin/env python
"""Generate tests/golden/netlib_golden_ormqr.npz: DORMQR / DORGQR outputs of the gfortran-compiled netlib LAPACK
3.12.0 inside scipy's OpenBLAS (same source of truth and caveats as make_golden.py), to pin oracle/ora_dormqr,
ora_dorm2r and ora_dorgqr.   Run:  python tests/golden/make_golden_ormqr.py"""
import ctypes as C
import glob
import os

import numpy as np
import scipy

HERE = os.path.dirname(os.path.abspath(__file__))
_so = glob.glob(os.path.join(os.path.dirname(scipy.__file__), "..", "scipy.libs", "libscipy_openblas*.so"))[0]
L = C.CDLL(_so)
dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int)
D = lambda a: a.ctypes.data_as(dp)
ci = lambda v: C.byref(C.c_int(v))


def rand(m, n, seed):
    s = np.array(seed, dtype=np.int32)
    x = np.empty(m * n)
    L.scipy_dlarnv_(ci(2), s.ctypes.data_as(ip), ci(m * n), D(x))
    return np.asfortranarray(x.reshape((n, m)).T)


out = {}
for tag, (m, n) in {"tall": (90, 40), "sq": (70, 70)}.items():
    a = rand(m, n, (1988, 1989, 1990, 1991))
    k = min(m, n)
    qr = a.copy(order="F")
    tau = np.zeros(k)
    work = np.zeros(64 * max(m, n) + 65 * 64)
    info = C.c_int(0)
    L.scipy_dgeqrf_(ci(m), ci(n), D(qr), ci(m), D(tau), D(work), ci(len(work)), C.byref(info))
    out[f"{tag}_a"], out[f"{tag}_qr"], out[f"{tag}_tau"] = a, qr, tau
    cl = rand(m, 9, (3, 5, 7, 9))          # for SIDE='L': Q is m x m
    cr = rand(9, m, (13, 15, 17, 19))      # for SIDE='R'
    for side, c0 in (("L", cl), ("R", cr)):
        for trans in "NT":
            c = c0.copy(order="F")
            L.scipy_dormqr_(C.c_char_p(side.encode()), C.c_char_p(trans.encode()), ci(c.shape[0]), ci(c.shape[1]), ci(k),
                            D(qr), ci(m), D(tau), D(c), ci(c.shape[0]), D(work), ci(len(work)), C.byref(info),
                            C.c_size_t(1), C.c_size_t(1))
            assert info.value == 0
            out[f"{tag}_ormqr_{side}{trans}"] = c
    out[f"{tag}_cl"], out[f"{tag}_cr"] = cl, cr
    q = qr.copy(order="F")
    L.scipy_dorgqr_(ci(m), ci(n), ci(k), D(q), ci(m), D(tau), D(work), ci(len(work)), C.byref(info))
    assert info.value == 0
    out[f"{tag}_q"] = q
np.savez_compressed(os.path.join(HERE, "netlib_golden_ormqr.npz"), **out)
print(sorted(out))
