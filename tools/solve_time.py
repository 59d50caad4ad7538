"""DPOTRS / DGETRS timing with few right-hand sides on device-resident factors: python tools/solve_time.py [N]
Both few-RHS modes: 1 = persistent streaming kernel (trsv_stream.cu), 0 = leaf/GEMV recursion."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import lapack_b200 as lb
L = lb.lib()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
a = lb.dev.larnv_matrix(n, n); lb.dev.make_spd(a, float(n)); lb.dev.potrf("L", a)
lu = lb.dev.larnv_matrix(n, n); piv, info = lb.dev.getrf(lu)
for mode in (1, 0):
    L.lb200_set_fewrhs_mode(mode)
    for nrhs in (1, 2, 4, 8):
        b0 = lb.dev.larnv_matrix(n, nrhs)
        b = b0.clone()
        for name, fn in (("potrs L", lambda: lb.dev.potrs("L", a, b)), ("getrs N", lambda: lb.dev.getrs("N", lu, piv, b)),
                         ("getrs T", lambda: lb.dev.getrs("T", lu, piv, b))):
            best = 1e9
            for rep in range(3):
                b.copy_(b0); torch.cuda.synchronize()
                l0 = L.lb200_launch_count()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); fn(); e1.record(); torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1))
            gbs = 8.0 * n * n / (best * 1e-3) * 1e-9
            print(f"mode={mode} {name} n={n} nrhs={nrhs}: {best:.3f} ms, {L.lb200_launch_count() - l0} launches, {gbs:.0f} GB/s (8n^2 B)", flush=True)
L.lb200_set_fewrhs_mode(1)
