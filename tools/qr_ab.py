"""A/B of the QR cluster leaf: timings with cluster_max = 0 / 16 and agreement of R, tau."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import lapack_b200 as lb
L = lb.lib()

def run(n, ncols, cm, reps=3):
    L.lb200_set_geqrf_cluster_max(cm)
    a0 = lb.dev.larnv_matrix(n, ncols)
    a = a0.clone()
    best, tau = 1e30, None
    for _ in range(reps):
        a.copy_(a0)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        tau = lb.dev.geqrf(a)
        e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best, tau.clone(), a.clone()

for n, ncols in ((3000, 16), (8192, 256), (16384, 256), (4096, 4096), (8192, 8192), (16384, 16384), (32768, 32768)):
    ref = None
    for cm in (0, 16):
        ms, tau, a = run(n, ncols, cm)
        same = ""
        if ref is None: ref = (tau, a)
        else: same = f"tau_maxdiff={float((tau - ref[0]).abs().max()):.3e} a_maxdiff={float((a - ref[1]).abs().max()):.3e} amax={float(a.abs().max()):.3e}"
        print("geqrf", n, ncols, "cluster_max", cm, f"{ms:.3f} ms", same, flush=True)
        del a, tau
    ref = None
    torch.cuda.empty_cache()
L.lb200_set_geqrf_cluster_max(16)
