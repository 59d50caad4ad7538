"""Outer block size sweep for DGETRF / DPOTRF at order N: python tools/nb_sweep.py [N]"""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import lapack_b200 as lb
L = lb.lib()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
a0 = lb.dev.larnv_matrix(n, n); a = a0.clone()
def t(fn, restore, reps=2):
    best = 1e30
    for _ in range(reps):
        restore(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best
for nb in (384, 512, 640, 768, 1024):
    L.lb200_set_getrf_params(nb, 0, 1)
    print("getrf nb", nb, f"{t(lambda: lb.dev.getrf(a), lambda: a.copy_(a0)):.1f} ms", flush=True)
L.lb200_set_getrf_params(512, 0, 1)
lb.dev.make_spd(a0, float(n))
for nb in (256, 512, 768, 1024):
    L.lb200_set_potrf_params(nb, 1)
    print("potrf nb", nb, f"{t(lambda: lb.dev.potrf('L', a), lambda: a.copy_(a0)):.1f} ms", flush=True)
L.lb200_set_potrf_params(512, 1)
