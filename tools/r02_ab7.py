"""two-level DGETRF: outer GEMM at 3 CTAs/SM (cfg 14) and thin inner leaves, so that the level-0 factorization finds free slots"""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import lapack_b200 as lb
L = lb.lib()
n = 32768
a0 = lb.dev.larnv_matrix(n, n)
a = a0.clone()
def run(tag, reps=2):
    best = 1e9
    for _ in range(reps):
        a.copy_(a0); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); piv, info = lb.dev.getrf(a); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    print(f"{tag}: {best:.1f} ms {(2*n**3/3)/best*1e-9:.2f} TFLOP/s", flush=True)
    return piv.clone()
L.lb200_set_getrf_super(0); p0 = run("single level", 3)
for snb in (4096, 2048):
    L.lb200_set_getrf_super(snb)
    for cfg in (-1, 14):
        L.lb200_set_gemm_config(cfg)
        for thin in ((0, 0), (2, 0), (1, 0), (2, 4096)):
            L.lb200_set_getrf_thin(thin[0], thin[1])
            p = run(f"super {snb} gemm cfg {cfg} thin {thin}")
            if not bool((p == p0).all()): print("   IPIV DIFFERS", flush=True)
L.lb200_set_getrf_super(4096); L.lb200_set_gemm_config(-1); L.lb200_set_getrf_thin(0, 16384)
