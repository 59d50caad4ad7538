#!/usr/bin/env python
"""Generate tests/golden/netlib_golden_gecon.npz: netlib 3.12.0 (gfortran, inside scipy's OpenBLAS; see make_golden.py) outputs of
DLATRS (plain DTRSV branch AND the scaled Level-1 branch, all UPLO/TRANS/DIAG), DGECON, DGEEQU and DGESVX on seeded inputs.
tests/test_oracle_golden.py replays them through oracle/ref_lapack.c.   Run: python tests/golden/make_golden_gecon.py"""
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
I = lambda a: a.ctypes.data_as(ip)
ci = lambda v: C.byref(C.c_int(v))
cc = lambda ch: C.c_char_p(ch.encode())
one = C.c_size_t(1)

out = {}
rng = np.random.default_rng(20261017)


def dlatrs(uplo, trans, diag, normin, a, x, cnorm):
    n = a.shape[0]
    scale, info = C.c_double(0.0), C.c_int(0)
    L.scipy_dlatrs_(cc(uplo), cc(trans), cc(diag), cc(normin), ci(n), D(a), ci(a.shape[0]), D(x), C.byref(scale), D(cnorm), C.byref(info),
                    one, one, one, one)
    return scale.value, info.value


def tri_case(kind, n, uplo):
    a = rng.uniform(-1, 1, (n, n))
    if kind == "plain":
        a += 4 * np.eye(n)
    elif kind == "tiny_diag":                      # growth bound fails -> scaled Level-1 branch with rescaling of x
        a[np.diag_indices(n)] = rng.uniform(1, 2, n) * 1e-300
    elif kind == "zero_diag":                      # exact zero pivot -> scale = 0, null-vector branch
        a += 4 * np.eye(n)
        a[n // 2, n // 2] = 0.0
    elif kind == "big_offdiag":                    # column norms beyond BIGNUM -> TSCAL != 1
        a *= 1e300
        a[np.diag_indices(n)] = rng.uniform(1, 2, n)
    a = np.triu(a) if uplo == "U" else np.tril(a)
    return np.asfortranarray(a)


k = 0
for kind in ("plain", "tiny_diag", "zero_diag", "big_offdiag"):
    for uplo in "UL":
        for trans in "NT":
            for diag in "NU":
                n = 23
                a = tri_case(kind, n, uplo)
                x = rng.uniform(-1, 1, n) * (1e10 if kind == "tiny_diag" else 1.0)
                cn = np.zeros(n)
                xo = x.copy()
                scale, info = dlatrs(uplo, trans, diag, "N", a, xo, cn)
                x2 = x.copy()
                cn2 = cn.copy()
                scale2, _ = dlatrs(uplo, trans, diag, "Y", a, x2, cn2)       # NORMIN = 'Y' reuses the norms
                out[f"trs{k}_meta"] = np.array([ord(uplo), ord(trans), ord(diag)], dtype=np.int32)
                out[f"trs{k}_a"], out[f"trs{k}_x"], out[f"trs{k}_xo"], out[f"trs{k}_cn"] = a, x, xo, cn
                out[f"trs{k}_scale"] = np.array([scale, scale2])
                out[f"trs{k}_xo2"] = x2
                k += 1
out["n_trs"] = np.array([k], dtype=np.int32)


def dgetrf(a):
    n = a.shape[0]
    ipiv, info = np.zeros(n, dtype=np.int32), C.c_int(0)
    L.scipy_dgetrf_(ci(n), ci(n), D(a), ci(n), I(ipiv), C.byref(info))
    return ipiv, info.value


k = 0
for (n, kind) in ((1, "rand"), (17, "rand"), (150, "rand"), (60, "hilbert"), (40, "graded"), (30, "singular")):
    if kind == "hilbert":
        a = 1.0 / (np.arange(n)[:, None] + np.arange(n)[None, :] + 1.0)
    elif kind == "graded":
        a = rng.uniform(-1, 1, (n, n)) * (10.0 ** np.linspace(0, -280, n))[None, :]
    else:
        a = rng.uniform(-1, 1, (n, n))
        if kind == "singular":
            a[:, 7] = 0.0
    a = np.asfortranarray(a)
    lu = a.copy(order="F")
    ipiv, info = dgetrf(lu)
    for norm in "1I":
        anorm = float(np.linalg.norm(a, 1 if norm == "1" else np.inf))
        work, iwork = np.zeros(4 * n), np.zeros(n, dtype=np.int32)
        rcond, cinfo = C.c_double(0.0), C.c_int(0)
        L.scipy_dgecon_(cc(norm), ci(n), D(lu), ci(n), C.byref(C.c_double(anorm)), C.byref(rcond), D(work), I(iwork), C.byref(cinfo), one)
        out[f"con{k}_{norm}"] = np.array([anorm, rcond.value, cinfo.value])
    out[f"con{k}_lu"] = lu
    k += 1
out["n_con"] = np.array([k], dtype=np.int32)

k = 0
for (n, nrhs, kind, fact, trans) in ((12, 2, "rand", "N", "N"), (50, 3, "badrow", "E", "N"), (50, 1, "badcol", "E", "T"), (33, 2, "both", "E", "N"),
                                     (33, 2, "both", "E", "T"), (20, 1, "singular", "N", "N"), (25, 2, "rand", "E", "N")):
    a = rng.uniform(-1, 1, (n, n))
    if kind in ("badrow", "both"):
        a *= (10.0 ** rng.uniform(-8, 8, n))[:, None]
    if kind in ("badcol", "both"):
        a *= (10.0 ** rng.uniform(-8, 8, n))[None, :]
    if kind == "singular":
        a[:, 5] = 0.0
    a = np.asfortranarray(a)
    xact = rng.uniform(-1, 1, (n, nrhs))
    b = np.asfortranarray((a if trans == "N" else a.T) @ xact)
    a_io, b_io = a.copy(order="F"), b.copy(order="F")
    af, ipiv = np.zeros((n, n), order="F"), np.zeros(n, dtype=np.int32)
    r, c = np.zeros(n), np.zeros(n)
    x = np.zeros((n, nrhs), order="F")
    ferr, berr, work, iwork = np.zeros(nrhs), np.zeros(nrhs), np.zeros(4 * n), np.zeros(n, dtype=np.int32)
    equed = C.create_string_buffer(b"N", 2)
    rcond, info = C.c_double(0.0), C.c_int(0)
    L.scipy_dgesvx_(cc(fact), cc(trans), ci(n), ci(nrhs), D(a_io), ci(n), D(af), ci(n), I(ipiv), equed, D(r), D(c), D(b_io), ci(n), D(x), ci(n),
                    C.byref(rcond), D(ferr), D(berr), D(work), I(iwork), C.byref(info), one, one, one)
    out[f"svx{k}_meta"] = np.array([ord(fact), ord(trans), ord(equed.value.decode()[0]), info.value], dtype=np.int32)
    out[f"svx{k}_a"], out[f"svx{k}_b"] = a, b
    out[f"svx{k}_a_out"], out[f"svx{k}_b_out"], out[f"svx{k}_af"], out[f"svx{k}_ipiv"] = a_io, b_io, af, ipiv
    out[f"svx{k}_r"], out[f"svx{k}_c"], out[f"svx{k}_x"] = r, c, x
    out[f"svx{k}_scal"] = np.array([rcond.value, work[0]])
    out[f"svx{k}_ferr"], out[f"svx{k}_berr"] = ferr, berr
    k += 1
out["n_svx"] = np.array([k], dtype=np.int32)
np.savez_compressed(os.path.join(HERE, "netlib_golden_gecon.npz"), **out)
print("wrote", len(out), "arrays")
