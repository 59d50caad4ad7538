"""DPOTRS / DGETRS timing with few right-hand sides on device-resident factors: python tools/solve_time.py [N]"""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import lapack_b200 as lb
L = lb.lib()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
a = lb.dev.larnv_matrix(n, n); lb.dev.make_spd(a, float(n)); lb.dev.potrf("L", a)
lu = lb.dev.larnv_matrix(n, n); piv, info = lb.dev.getrf(lu)
for nrhs in (1, 4, 8, 16):
    b = lb.dev.larnv_matrix(n, nrhs)
    for name, fn in (("potrs", lambda: lb.dev.potrs("L", a, b)), ("getrs", lambda: lb.dev.getrs("N", lu, piv, b))):
        fn(); torch.cuda.synchronize()
        l0 = L.lb200_launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        print(f"{name} n={n} nrhs={nrhs}: {e0.elapsed_time(e1):.2f} ms, {L.lb200_launch_count() - l0} launches", flush=True)
