#!/usr/bin/env python
"""Generate tests/golden/netlib_golden_gels.npz: DGELS (all four M>=N / M<N x TRANS cases, plus the scaling branches) and
DGELQF outputs of netlib LAPACK 3.12.0 (scipy's OpenBLAS build; same caveats as make_golden.py), to pin oracle/ora_dgels,
ora_dgelq2, ora_dorml2, ora_dtrtrs, ora_dlascl_g.   Run:  python tests/golden/make_golden_gels.py"""
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


def gels(trans, a, b):
    m, n = a.shape
    a, b = a.copy(order="F"), b.copy(order="F")
    work = np.zeros(64 * (max(m, n) + b.shape[1]) + min(m, n))
    info = C.c_int(0)
    L.scipy_dgels_(C.c_char_p(trans.encode()), ci(m), ci(n), ci(b.shape[1]), vp(a), ci(m), vp(b), ci(max(m, n)), vp(work),
                   ci(len(work)), C.byref(info), C.c_size_t(1))
    return a, b, info.value


out = {}
for tag, (m, n, nrhs) in {"tall": (90, 40, 3), "wide": (40, 90, 2), "sq": (50, 50, 1)}.items():
    a = rand(m, n, (1988, 1989, 1990, 1991))
    b = rand(max(m, n), nrhs, (3, 5, 7, 9))
    out[f"{tag}_a"], out[f"{tag}_b"] = a, b
    for trans in "NT":
        af, x, info = gels(trans, a, b)
        assert info == 0
        out[f"{tag}_{trans}_af"], out[f"{tag}_{trans}_x"] = af, x
# scaling branches (dgels.f:309-353) and exact rank deficiency (INFO = i from DTRTRS)
a, b = out["tall_a"], out["tall_b"]
for k, (sa, sb) in enumerate(((1e-300, 1.0), (1e300, 1.0), (1.0, 1e-300), (1.0, 1e300))):
    _, x, info = gels("N", a * sa, b * sb)
    assert info == 0
    out[f"scale{k}_x"] = x
    out[f"scale{k}_s"] = np.array([sa, sb])
a0 = a.copy(order="F")
a0[:, 4] = 0.0
_, _, info = gels("N", a0, b)
out["rankdef_info"] = np.int32(info)
np.savez_compressed(os.path.join(HERE, "netlib_golden_gels.npz"), **out)
print(len(out), "arrays; rank-deficient INFO =", info)
