"""A/B of the inverted-diagonal-block TRSM leaves inside DGETRF / DPOTRF (lb200_set_trsm_inverse): time, IPIV, factor agreement."""
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
ref = None
for inv in (1, 0):
    L.lb200_set_trsm_inverse(inv)
    ms, (piv, info) = timed(lambda: lb.dev.getrf(a), lambda: a.copy_(a0))
    print(f"DGETRF n={n} trsm_inverse={inv}: {ms:.1f} ms {(2*n**3/3)/ms*1e-9:.2f} TFLOP/s info={int(info)}", flush=True)
    if ref is None: ref = (piv.clone(), a.clone())
    else:
        d = (a - ref[1]).abs().max().item(); s = ref[1].abs().max().item()
        print(f"   ipiv equal {bool((piv == ref[0]).all())}  max|LU diff| {d:.3e} (max|LU| {s:.3e})", flush=True)
del ref
lb.dev.make_spd(a0, float(n))
ref = None
for inv in (1, 0):
    L.lb200_set_trsm_inverse(inv)
    ms, info = timed(lambda: lb.dev.potrf("L", a), lambda: a.copy_(a0))
    print(f"DPOTRF n={n} trsm_inverse={inv}: {ms:.1f} ms {(n**3/3)/ms*1e-9:.2f} TFLOP/s info={int(info)}", flush=True)
    if ref is None: ref = torch.tril(a).clone()
    else:
        d = (torch.tril(a) - ref).abs().max().item(); s = ref.abs().max().item()
        print(f"   max|L diff| {d:.3e} (max|L| {s:.3e})", flush=True)
L.lb200_set_trsm_inverse(1)
