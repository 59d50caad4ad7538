"""One DPOTRS('L'), DGETRS('N'), DGETRS('T') with 1 RHS on device-resident factors (for ncu): python tools/prof_solve.py [N] [reps]"""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import lapack_b200 as lb
n = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
a = lb.dev.larnv_matrix(n, n); lb.dev.make_spd(a, float(n)); lb.dev.potrf("L", a)
lu = lb.dev.larnv_matrix(n, n); piv, info = lb.dev.getrf(lu)
b = lb.dev.larnv_matrix(n, 1)
torch.cuda.synchronize()
for _ in range(reps):
    lb.dev.potrs("L", a, b)
    lb.dev.getrs("N", lu, piv, b)
    lb.dev.getrs("T", lu, piv, b)
torch.cuda.synchronize()
print("done")
