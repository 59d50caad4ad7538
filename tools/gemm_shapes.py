"""Timing of the default GEMM on the shapes the factorizations issue (all transposition combos, SYRK masks)."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import lapack_b200 as lb

def t(fn, reps=5):
    fn(); torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best

for (m, n, k) in ((16384, 16384, 512), (16384, 16384, 256), (24576, 24576, 512), (8192, 8192, 512), (32768, 512, 512), (16384, 16384, 8192)):
    for ta, tb in (("N", "N"), ("N", "T"), ("T", "N")):
        ar, ac = (m, k) if ta == "N" else (k, m); br, bc = (k, n) if tb == "N" else (n, k)
        A = lb.dev.colmajor(ar, ac); A.normal_(); B = lb.dev.colmajor(br, bc); B.normal_(); C = lb.dev.colmajor(m, n); C.normal_()
        ms = t(lambda: lb.dev.gemm(ta, tb, -1.0, A, B, 1.0, C))
        print(f"gemm {ta}{tb} {m}x{n}x{k}: {ms:.3f} ms {2*m*n*k/ms*1e-9:.2f} TF/s", flush=True)
        del A, B, C
    if m == n:
        A = lb.dev.colmajor(n, k); A.normal_(); C = lb.dev.colmajor(n, n); C.normal_()
        ms = t(lambda: lb.dev.syrk("L", "N", -1.0, A, 1.0, C))
        print(f"syrk LN {n}x{k}: {ms:.3f} ms {n*n*k/ms*1e-9:.2f} TF/s (n^2 k flops)", flush=True)
        del A, C
