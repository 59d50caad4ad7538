"""A/B of LU driver knobs on one GPU: python tools/lu_knobs.py [n]"""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import lapack_b200 as lb
L = lb.lib()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
a0 = lb.dev.larnv_matrix(n, n)
a = a0.clone()
def run(tag):
    best = 1e9
    for _ in range(3):
        a.copy_(a0); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); piv, info = lb.dev.getrf(a); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    print(f"{tag}: {best:.1f} ms  {(2*n**3/3)/best*1e-9:.2f} TFLOP/s", flush=True)
    return piv.clone()
p0 = run("default")
L.lb200_set_getrf_cluster_fat(1)
p = run("cluster_fat (8 rows/thread, one 16-CTA cluster up to 32768 rows)")
print("  ipiv equal:", bool((p == p0).all()))
L.lb200_set_getrf_cluster_fat(0)
if len(sys.argv) > 2: sys.exit(0)
for rows in (2048, 4096):
    L.lb200_set_getrf_tall_rows(rows)
    p = run(f"tall_rows={rows}")
    print("  ipiv equal:", bool((p == p0).all()))
L.lb200_set_getrf_tall_rows(1024)
for cm in (8, 4):
    L.lb200_set_getrf_cluster_max(cm)
    for rows in (1024, 2048, 4096):
        L.lb200_set_getrf_tall_rows(rows)
        p = run(f"cluster_max={cm} tall_rows={rows}")
        print("  ipiv equal:", bool((p == p0).all()))
L.lb200_set_getrf_cluster_max(16); L.lb200_set_getrf_tall_rows(1024)
