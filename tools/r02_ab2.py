"""A/B after the lean GEMM became the default: LU (per-chunk events), Cholesky, QR with / without the wave-balancing split-K."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import lapack_b200 as lb
L = lb.lib()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 32768

def timed(fn, restore, reps=3):
    best = 1e9; out = None
    for _ in range(reps):
        restore(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); out = fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best, out

a0 = lb.dev.larnv_matrix(n, n)
a = a0.clone()
ms, (piv, info) = timed(lambda: lb.dev.getrf(a), lambda: a.copy_(a0))
print(f"DGETRF n={n}: {ms:.1f} ms {(2*n**3/3)/ms*1e-9:.2f} TFLOP/s info={int(info)}", flush=True)
L.lb200_set_gemm_config(8)
ms8, (piv8, info8) = timed(lambda: lb.dev.getrf(a), lambda: a.copy_(a0), 2)
print(f"DGETRF n={n} gemm cfg 8 (general loader): {ms8:.1f} ms  ipiv equal {bool((piv8 == piv).all())}", flush=True)
L.lb200_set_gemm_config(-1)
# QR
for sk in (1, 0):
    L.lb200_set_gemm_splitk_balance(sk)
    ms, tau = timed(lambda: lb.dev.geqrf(a), lambda: a.copy_(a0), 2)
    print(f"DGEQRF n={n} splitk_balance={sk}: {ms:.1f} ms {(4*n**3/3)/ms*1e-9:.2f} TFLOP/s", flush=True)
    if sk == 1: r1 = a.clone(); t1 = tau.clone()
    else: print(f"   max|R diff| {float((a - r1).abs().max()):.3e} (|R|max {float(a.abs().max()):.3e})  max|tau diff| {float((tau - t1).abs().max()):.3e}", flush=True)
L.lb200_set_gemm_splitk_balance(1)
del r1
# Cholesky
lb.dev.make_spd(a0, float(n))
ms, info = timed(lambda: lb.dev.potrf("L", a), lambda: a.copy_(a0))
print(f"DPOTRF n={n}: {ms:.1f} ms {(n**3/3)/ms*1e-9:.2f} TFLOP/s info={int(info)}", flush=True)
L.lb200_set_gemm_config(8)
ms, info = timed(lambda: lb.dev.potrf("L", a), lambda: a.copy_(a0), 2)
print(f"DPOTRF n={n} gemm cfg 8: {ms:.1f} ms", flush=True)
L.lb200_set_gemm_config(-1)
