"""A/B of the deferred interchanges left of the panel (lb200_set_getrf_defer_left): time, and bit-for-bit agreement of IPIV / factors."""
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
for on, tail in ((0, 10240), (1, 512), (1, 1024), (1, 2048), (1, 3072), (1, 4096)):
    L.lb200_set_getrf_defer_left(on, tail)
    ms, (piv, info) = timed(lambda: lb.dev.getrf(a), lambda: a.copy_(a0))
    if ref is None: ref = (piv.clone(), a.clone())
    same = bool(torch.equal(piv, ref[0]) and torch.equal(a, ref[1]))
    print(f"DGETRF n={n} defer_left={on} tail_rows={tail}: {ms:.1f} ms {(2*n**3/3)/ms*1e-9:.2f} TFLOP/s  bitwise == immediate: {same}", flush=True)
L.lb200_set_getrf_defer_left(1, 10240)
# odd shapes (tall, wide, small): deferred == immediate
for (m, nn) in ((20000, 12000), (12000, 20000), (9000, 9000), (30000, 1500), (8192, 8192), (60000, 4096)):
    x0 = lb.dev.larnv_matrix(m, nn); outs = []
    for on in (0, 1):
        L.lb200_set_getrf_defer_left(on, 2048)
        x = x0.clone(); piv, info = lb.dev.getrf(x); torch.cuda.synchronize(); outs.append((piv.clone(), x))
    print(f"shape {m}x{nn}: bitwise equal {bool(torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1]))}", flush=True)
L.lb200_set_getrf_defer_left(1, 10240)
