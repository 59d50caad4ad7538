"""A/B of the cluster leaf: python tools/cluster_ab.py  (panel and full-LU timings with cluster_max = 0 / 8 / 16)"""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import lapack_b200 as lb
L = lb.lib()

def run(which, n, cm, reps=3):
    L.lb200_set_getrf_cluster_max(cm)
    a0 = lb.dev.larnv_matrix(n, 512 if which == "panel" else n)
    a = a0.clone()
    best, piv = 1e30, None
    for _ in range(reps):
        a.copy_(a0)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        piv, info = lb.dev.getrf(a, recursive=(which == "panel"))
        e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best, piv.clone(), a.clone()

for which, n in (("panel", 2048), ("panel", 8192), ("panel", 16384), ("getrf", 4096), ("getrf", 8192), ("getrf", 16384), ("getrf", 32768)):
    ref = None
    for cm in (0, 8, 16):
        try:
            ms, piv, a = run(which, n, cm)
        except Exception as e:
            print(which, n, "cluster_max", cm, "FAILED", e, flush=True)
            continue
        same = ""
        if ref is None: ref = (piv, a)
        else: same = f"ipiv_same={bool((piv == ref[0]).all())} maxdiff={float((a - ref[1]).abs().max()):.3e}"
        print(which, n, "cluster_max", cm, f"{ms:.3f} ms", same, flush=True)
L.lb200_set_getrf_cluster_max(16)
