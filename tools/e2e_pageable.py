"""dgetrf_ / dposv_ on PAGEABLE host arrays at order N (run twice: LAPACK_B200_HOST_REGISTER=0 and =1): python tools/e2e_pageable.py [N]"""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import lapack_b200 as lb
n = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
a0 = lb.dev.larnv_matrix(n, n)
s0 = a0.clone(); lb.dev.make_spd(s0, float(n))
h = np.empty((n, n), dtype=np.float64)             # pageable; row-major (n,n) == column-major n x n
b = np.ones((1, n))
ipiv = np.zeros(n, dtype=np.int32)
def t(fn, src, reps=2):
    out = []
    for _ in range(reps):
        torch.from_numpy(h).copy_(src.t()); torch.cuda.synchronize()
        t0 = time.perf_counter(); r = fn(); out.append((time.perf_counter() - t0) * 1e3)
        assert r == 0, r
    return out
print("HOST_REGISTER =", os.environ.get("LAPACK_B200_HOST_REGISTER", "1 (default)"))
print("dgetrf_ pageable ms:", t(lambda: lb.f77.dgetrf(n, n, h.ctypes.data, n, ipiv), a0))
print("dposv_ L pageable ms:", t(lambda: lb.f77.dposv("L", n, 1, h.ctypes.data, n, b.ctypes.data, n), s0))
