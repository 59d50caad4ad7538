"""Run a few launches of one GEMM configuration (for ncu): python tools/prof_gemm.py M N K CFG [TA TB]"""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import lapack_b200 as lb
m, n, k, cfg = [int(x) for x in sys.argv[1:5]]
ta, tb = (sys.argv[5], sys.argv[6]) if len(sys.argv) > 6 else ("N", "N")
L = lb.lib(); L.lb200_set_gemm_config(cfg)
ar, ac = (m, k) if ta == "N" else (k, m); br, bc = (k, n) if tb == "N" else (n, k)
A = lb.dev.colmajor(ar, ac); A.normal_(); B = lb.dev.colmajor(br, bc); B.normal_(); Cm = lb.dev.colmajor(m, n); Cm.normal_()
for _ in range(3):
    lb.dev.gemm(ta, tb, -1.0, A, B, 1.0, Cm)
torch.cuda.synchronize()
