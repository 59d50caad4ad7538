"""A/B of the two-level DGETRF driver (lb200_set_getrf_super): time, IPIV equality and factor agreement with the single-level driver."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import lapack_b200 as lb
L = lb.lib()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
sizes = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [0, 4096, 2048, 8192]

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
for snb in sizes:
    L.lb200_set_getrf_super(snb)
    ms, (piv, info) = timed(lambda: lb.dev.getrf(a), lambda: a.copy_(a0), 2 if ref is not None else 3)
    line = f"DGETRF n={n} super_nb={snb}: {ms:.1f} ms {(2*n**3/3)/ms*1e-9:.2f} TFLOP/s info={int(info)}"
    if ref is None: ref = (piv.clone(), a.clone())
    else:
        d = (a - ref[1]).abs().max().item(); sc = ref[1].abs().max().item()
        line += f"  ipiv equal {bool((piv == ref[0]).all())}  max|LU diff| {d:.3e} (max|LU| {sc:.3e})"
    print(line, flush=True)
# residual check of the last run by a randomized product: || P A x - L U x || / (n ||A|| ||x|| eps)
x = torch.randn(n, 1, dtype=torch.float64, device=a.device)
Ux = torch.triu(a) @ x
LUx = torch.tril(a, -1) @ Ux + Ux
pa = a0.clone()
pv = piv.cpu().numpy()
perm = list(range(n))
for i in range(n):
    p = int(pv[i]) - 1
    if p != i: perm[i], perm[p] = perm[p], perm[i]
PAx = (a0 @ x)[torch.tensor(perm, device=a.device)]
r = (PAx - LUx).abs().max().item() / (n * a0.abs().max().item() * x.abs().max().item() * 2.0 ** -53)
print(f"randomized residual ratio (last configuration): {r:.3e}", flush=True)
L.lb200_set_getrf_super(4096)
