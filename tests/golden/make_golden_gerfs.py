#!/usr/bin/env python
"""Generate tests/golden/netlib_golden_gerfs.npz: DGERFS outputs of netlib LAPACK 3.12.0 (scipy's OpenBLAS build; same caveats
as make_golden.py) on perturbed DGETRS solutions, to pin oracle/ora_dgerfs and ora_dlacn2.
Run:  python tests/golden/make_golden_gerfs.py"""
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
for n, nrhs in ((7, 2), (80, 3)):
    a = rand(n, n, (1988, 1989, 1990, 1991))
    xt = rand(n, nrhs, (3, 5, 7, 9))
    af = a.copy(order="F")
    ipiv = np.zeros(n, dtype=np.int32)
    info = C.c_int(0)
    L.scipy_dgetrf2_(ci(n), ci(n), vp(af), ci(n), vp(ipiv), C.byref(info))
    out[f"a{n}"], out[f"af{n}"], out[f"ipiv{n}"] = a, af, ipiv
    for trans in "NT":
        b = np.asfortranarray((a if trans == "N" else a.T) @ xt)
        x0 = np.asfortranarray(xt * (1.0 + 1e-9))                 # a slightly wrong solution: refinement has work to do
        x = x0.copy(order="F")
        ferr, berr = np.zeros(nrhs), np.zeros(nrhs)
        work, iwork = np.zeros(3 * n), np.zeros(n, dtype=np.int32)
        L.scipy_dgerfs_(C.c_char_p(trans.encode()), ci(n), ci(nrhs), vp(a), ci(n), vp(af), ci(n), vp(ipiv), vp(b), ci(n), vp(x), ci(n),
                        vp(ferr), vp(berr), vp(work), vp(iwork), C.byref(info), C.c_size_t(1))
        assert info.value == 0
        out[f"b{n}{trans}"], out[f"x0_{n}{trans}"], out[f"x{n}{trans}"] = b, x0, x
        out[f"ferr{n}{trans}"], out[f"berr{n}{trans}"] = ferr, berr
np.savez_compressed(os.path.join(HERE, "netlib_golden_gerfs.npz"), **out)
print(len(out), "arrays")
