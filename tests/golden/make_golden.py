#!/usr/bin/env python
"""Generate tests/golden/netlib_golden.npz -- reference outputs that pin the oracle.

The reference tree ships no golden vectors for this path and its Fortran cannot be
compiled in this image (no Fortran compiler).  The closest thing to "the reference run here" is
the gfortran-compiled netlib LAPACK 3.12.0 bundled inside scipy's OpenBLAS, whose Fortran-ABI
symbols are reachable through ctypes with a `scipy_` prefix.  This script calls those routines on
seeded inputs and stores inputs + outputs; tests/test_oracle_golden.py replays the inputs through
oracle/ and compares.

Caveat recorded with the data: the BLAS underneath those netlib routines is OpenBLAS (different
summation order), so factor entries agree to rounding (~1e-13 relative) rather than bit-for-bit;
IPIV, INFO and the DLARNV stream must agree exactly.

Run:  python tests/golden/make_golden.py      (needs scipy; writes next to this file)
"""
import ctypes as C
import glob
import os

import numpy as np
import scipy

HERE = os.path.dirname(os.path.abspath(__file__))
_so = glob.glob(os.path.join(os.path.dirname(scipy.__file__), "..", "scipy.libs", "libscipy_openblas*.so"))[0]
L = C.CDLL(_so)
dp = C.POINTER(C.c_double)
ip = C.POINTER(C.c_int)


def D(a):
    return a.ctypes.data_as(dp)


def I(a):
    return a.ctypes.data_as(ip)


def ci(v):
    return C.byref(C.c_int(v))


def dlarnv(idist, iseed, n):
    seed = np.array(iseed, dtype=np.int32)
    x = np.empty(n)
    L.scipy_dlarnv_(ci(idist), I(seed), ci(n), D(x))
    return x, seed


def rand(m, n, seed):
    x, s = dlarnv(2, seed, m * n)
    return np.asfortranarray(x.reshape((n, m)).T), s


out = {}
ver = (C.c_int(), C.c_int(), C.c_int())
L.scipy_ilaver_(C.byref(ver[0]), C.byref(ver[1]), C.byref(ver[2]))
out["ilaver"] = np.array([v.value for v in ver], dtype=np.int32)

# --- DLARNV known answers (bit exact) -------------------------------------------------------
for idist in (1, 2, 3):
    x, s = dlarnv(idist, (1988, 1989, 1990, 1991), 300)
    out[f"larnv{idist}_x"] = x
    out[f"larnv{idist}_seed"] = s

# --- DGETRF2 (recursive LU) on tall, square and wide inputs ---------------------------------
for tag, (m, n) in {"tall": (70, 40), "sq": (64, 64), "wide": (30, 50), "one": (1, 5), "col": (9, 1)}.items():
    a, _ = rand(m, n, (1988, 1989, 1990, 1991))
    lu = a.copy(order="F")
    ipiv = np.zeros(min(m, n), dtype=np.int32)
    info = C.c_int(0)
    L.scipy_dgetrf2_(ci(m), ci(n), D(lu), ci(m), I(ipiv), C.byref(info))
    out[f"getrf2_{tag}_a"] = a
    out[f"getrf2_{tag}_lu"] = lu
    out[f"getrf2_{tag}_ipiv"] = ipiv
    out[f"getrf2_{tag}_info"] = np.int32(info.value)

# singular: zero column 3 -> INFO = 3, factorization continues
a, _ = rand(20, 20, (7, 8, 9, 11))
a[:, 2] = 0.0
lu = a.copy(order="F")
ipiv = np.zeros(20, dtype=np.int32)
info = C.c_int(0)
L.scipy_dgetrf2_(ci(20), ci(20), D(lu), ci(20), I(ipiv), C.byref(info))
out["getrf2_sing_a"], out["getrf2_sing_lu"], out["getrf2_sing_ipiv"] = a, lu, ipiv
out["getrf2_sing_info"] = np.int32(info.value)

# --- DPOTRF2 (recursive Cholesky), both triangles, and a non-SPD case -----------------------
for uplo in ("L", "U"):
    r, _ = rand(48, 48, (1988, 1989, 1990, 1991))
    s = np.asfortranarray((r + r.T) * 0.5 + 48 * np.eye(48))
    f = s.copy(order="F")
    info = C.c_int(0)
    L.scipy_dpotrf2_(C.c_char_p(uplo.encode()), ci(48), D(f), ci(48), C.byref(info), C.c_size_t(1))
    out[f"potrf2_{uplo}_a"], out[f"potrf2_{uplo}_f"], out[f"potrf2_{uplo}_info"] = s, f, np.int32(info.value)
s2 = s.copy(order="F")
s2[10, 10] = -1.0
f = s2.copy(order="F")
info = C.c_int(0)
L.scipy_dpotrf2_(C.c_char_p(b"L"), ci(48), D(f), ci(48), C.byref(info), C.c_size_t(1))
out["potrf2_bad_a"], out["potrf2_bad_info"] = s2, np.int32(info.value)

# --- DPOTRS with the netlib routine ---------------------------------------------------------
b, _ = rand(48, 3, (5, 6, 7, 9))
x = b.copy(order="F")
info = C.c_int(0)
L.scipy_dpotrs_(C.c_char_p(b"L"), ci(48), ci(3), D(out["potrf2_L_f"]), ci(48), D(x), ci(48), C.byref(info),
                C.c_size_t(1))
out["potrs_b"], out["potrs_x"] = b, x

# --- DLARFG ---------------------------------------------------------------------------------
v, _ = dlarnv(2, (11, 12, 13, 15), 33)
alpha = C.c_double(v[0])
xx = v[1:].copy()
tau = C.c_double(0)
L.scipy_dlarfg_(ci(33), C.byref(alpha), D(xx), ci(1), C.byref(tau))
out["larfg_in"], out["larfg_beta"], out["larfg_tau"], out["larfg_v"] = v, np.float64(alpha.value), np.float64(tau.value), xx
# tiny-norm rescale branch (dlarfg.f:159-176)
vt = v * 1e-300
alpha = C.c_double(vt[0])
xx = vt[1:].copy()
tau = C.c_double(0)
L.scipy_dlarfg_(ci(33), C.byref(alpha), D(xx), ci(1), C.byref(tau))
out["larfg_tiny_in"], out["larfg_tiny_beta"], out["larfg_tiny_tau"], out["larfg_tiny_v"] = vt, np.float64(alpha.value), np.float64(tau.value), xx

# --- DGEQR2 / DGEQRF / DLARFT / DLARFB / DORGQR ---------------------------------------------
for tag, (m, n) in {"tall": (90, 40), "sq": (50, 50), "wide": (30, 45), "big": (200, 150)}.items():
    a, _ = rand(m, n, (1988, 1989, 1990, 1991))
    k = min(m, n)
    qr = a.copy(order="F")
    tau = np.zeros(k)
    work = np.zeros(max(1, n) * 64)
    info = C.c_int(0)
    L.scipy_dgeqrf_(ci(m), ci(n), D(qr), ci(m), D(tau), D(work), ci(len(work)), C.byref(info))
    out[f"geqrf_{tag}_a"], out[f"geqrf_{tag}_qr"], out[f"geqrf_{tag}_tau"] = a, qr, tau
    if tag != "big":
        q2 = a.copy(order="F")
        tau2 = np.zeros(k)
        L.scipy_dgeqr2_(ci(m), ci(n), D(q2), ci(m), D(tau2), D(work), C.byref(info))
        out[f"geqr2_{tag}_qr"], out[f"geqr2_{tag}_tau"] = q2, tau2

qr, tau = out["geqrf_tall_qr"], out["geqrf_tall_tau"]
m, k = 90, 40
t = np.zeros((k, k), order="F")
L.scipy_dlarft_(C.c_char_p(b"F"), C.c_char_p(b"C"), ci(m), ci(k), D(qr), ci(m), D(tau), D(t), ci(k),
                C.c_size_t(1), C.c_size_t(1))
out["larft_t"] = np.triu(t)
cmat, _ = rand(m, 17, (21, 22, 23, 25))
for trans in ("T", "N"):
    c2 = cmat.copy(order="F")
    work = np.zeros((17, k), order="F")
    L.scipy_dlarfb_(C.c_char_p(b"L"), C.c_char_p(trans.encode()), C.c_char_p(b"F"), C.c_char_p(b"C"), ci(m), ci(17),
                    ci(k), D(qr), ci(m), D(t), ci(k), D(c2), ci(m), D(work), ci(17), C.c_size_t(1), C.c_size_t(1),
                    C.c_size_t(1), C.c_size_t(1))
    out[f"larfb_L{trans}_c"] = c2
out["larfb_c_in"] = cmat
q = np.zeros((m, m), order="F")
q[:, :k] = np.tril(qr, -1)[:, :k]
work = np.zeros(m * 64)
info = C.c_int(0)
L.scipy_dorgqr_(ci(m), ci(m), ci(k), D(q), ci(m), D(tau), D(work), ci(len(work)), C.byref(info))
out["orgqr_q"] = q

np.savez_compressed(os.path.join(HERE, "netlib_golden.npz"), **out)
print("wrote", os.path.join(HERE, "netlib_golden.npz"), "LAPACK", out["ilaver"], "keys", len(out))
