"""Host-pointer (pinned) timings of the individual Fortran-ABI calls at order N: python tools/e2e_parts.py [N]"""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import lapack_b200 as lb
n = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
a0 = lb.dev.larnv_matrix(n, n)
s0 = a0.clone(); lb.dev.make_spd(s0, float(n))
h = torch.empty((n, n), dtype=torch.float64).pin_memory()
ipiv = np.zeros(n, dtype=np.int32)
b = torch.ones((1, n), dtype=torch.float64).pin_memory()
def t(fn, src, reps=3):
    best = 1e30
    for _ in range(reps):
        h.copy_(src.t()); torch.cuda.synchronize()
        t0 = time.perf_counter(); r = fn(); dt = time.perf_counter() - t0
        assert r == 0, r
        best = min(best, dt)
    return best * 1e3
print("dgetrf_ pinned", t(lambda: lb.f77.dgetrf(n, n, h.data_ptr(), n, ipiv), a0), "ms")
print("dpotrf_ L pinned", t(lambda: lb.f77.dpotrf("L", n, h.data_ptr(), n), s0), "ms")
print("dpotrf_ U pinned", t(lambda: lb.f77.dpotrf("U", n, h.data_ptr(), n), s0), "ms")
print("dposv_ L pinned", t(lambda: lb.f77.dposv("L", n, 1, h.data_ptr(), n, b.data_ptr(), n), s0), "ms")
tau = np.zeros(n); wq = np.zeros(1)
lb.f77.dgeqrf(n, n, h.data_ptr(), n, tau, wq, -1); work = np.zeros(int(wq[0]))
print("dgeqrf_ pinned", t(lambda: lb.f77.dgeqrf(n, n, h.data_ptr(), n, tau, work, len(work)), a0, reps=2), "ms")
