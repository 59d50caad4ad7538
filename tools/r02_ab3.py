"""A/B: update GEMM at 3 CTAs/SM (cfg 14) so that one quarter of every SM stays free for the panel kernels, with thin LU leaves."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import lapack_b200 as lb
L = lb.lib()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 32768

def timed(fn, restore, reps=2):
    best = 1e9; out = None
    for _ in range(reps):
        restore(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); out = fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best, out

a0 = lb.dev.larnv_matrix(n, n)
a = a0.clone()
p0 = None
for cfg in (-1, 14):
    L.lb200_set_gemm_config(cfg)
    for mode, mr in ((0, 0), (1, 16384), (2, 16384), (1, 8192), (2, 8192), (2, 4096)):
        L.lb200_set_getrf_thin(mode, mr)
        ms, (piv, info) = timed(lambda: lb.dev.getrf(a), lambda: a.copy_(a0))
        if p0 is None: p0 = piv.clone()
        print(f"DGETRF n={n} gemm cfg {cfg} thin {mode}/{mr}: {ms:.1f} ms {(2*n**3/3)/ms*1e-9:.2f} TFLOP/s ipiv equal {bool((piv == p0).all())}", flush=True)
L.lb200_set_getrf_thin(0, 16384)
for cfg in (-1, 14):
    L.lb200_set_gemm_config(cfg)
    ms, tau = timed(lambda: lb.dev.geqrf(a), lambda: a.copy_(a0))
    print(f"DGEQRF n={n} gemm cfg {cfg}: {ms:.1f} ms {(4*n**3/3)/ms*1e-9:.2f} TFLOP/s", flush=True)
lb.dev.make_spd(a0, float(n))
for cfg in (-1, 14):
    L.lb200_set_gemm_config(cfg)
    ms, info = timed(lambda: lb.dev.potrf("L", a), lambda: a.copy_(a0))
    print(f"DPOTRF n={n} gemm cfg {cfg}: {ms:.1f} ms {(n**3/3)/ms*1e-9:.2f} TFLOP/s info={int(info)}", flush=True)
L.lb200_set_gemm_config(-1)
