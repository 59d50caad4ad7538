import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import lapack_b200 as lb
L = lb.lib()
def run(which, n, big, reps=3):
    L.lb200_set_getrf_big_leaf(big)
    a0 = lb.dev.larnv_matrix(n, 512 if which == "panel" else n)
    a = a0.clone(); best = 1e30
    for _ in range(reps):
        a.copy_(a0); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); piv, info = lb.dev.getrf(a, recursive=(which == "panel")); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best, piv.clone(), a
for which, n in (("panel", 32768), ("panel", 131072), ("getrf", 24576), ("getrf", 32768)):
    ref = None
    for big in (0, 1):
        ms, piv, a = run(which, n, big)
        same = "" if ref is None else f"ipiv_same={bool((piv == ref[0]).all())} maxdiff={float((a - ref[1]).abs().max()):.2e}"
        if ref is None: ref = (piv, a.clone())
        print(which, n, "big_leaf", big, f"{ms:.3f} ms", same, flush=True)
        del a
    ref = None; torch.cuda.empty_cache()
