"""batched 32x32 DGETRF: one matrix per warp (mode 0) vs two matrices per warp (mode 2, the default): timing on random matrices and
bit-for-bit agreement, including the special cases (zero / NaN / Inf entries, exact ties, odd batch).  The other variants listed in
profiles/r02_batched_two_per_warp_ab.txt were measured with this script and then removed from the library."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import lapack_b200 as lb
L = lb.lib()
batch = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
modes = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [0, 2]

def run(a0, mode, reps):
    L.lb200_set_batched_mode(mode)
    a = a0.clone(); best = 1e9
    for _ in range(reps):
        a.copy_(a0); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ipiv, info = lb.dev.getrf_batched32(a); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best, a.clone(), ipiv.clone(), info.clone()

def same(x, y):   # bitwise, NaN-safe
    return bool(torch.equal(x.view(torch.int64), y.view(torch.int64))) if x.dtype == torch.float64 else bool(torch.equal(x, y))

# special cases: 4097 matrices (odd batch)
g = torch.Generator(device="cpu"); g.manual_seed(11)
sp = torch.randn(4097, 32, 32, dtype=torch.float64, generator=g)
sp[0] = 0.0                                   # zero matrix
sp[1, :, 5] = 0.0                             # zero column
sp[2] = 1.0                                   # all ties
sp[3, 7, 0] = float("nan")                    # NaN below the first place
sp[4, 0, 0] = float("nan")                    # NaN in the first place
sp[5, 9, 3] = float("inf")
sp[6] = torch.randint(-2, 3, (32, 32), generator=g).double()       # many exact ties
sp[7] = sp[7] * 1e-310                        # denormals (|pivot| < SFMIN branch)
sp[8, :, :] = torch.arange(32, dtype=torch.float64).view(32, 1)   # rank one
for k in range(9, 64): sp[k] = torch.randint(-1, 2, (32, 32), generator=g).double()
spd = sp.transpose(1, 2).contiguous().cuda()   # kernel layout: column-major 32x32 per matrix
ref = None
for mode in modes:
    ms, a, ipiv, info = run(spd, mode, 1)
    if ref is None: ref = (a, ipiv, info)
    else: print(f"special cases mode {mode}: factors {same(a, ref[0])} ipiv {same(ipiv, ref[1])} info {same(info, ref[2])}", flush=True)
a0 = lb.dev.larnv_matrix(32, 32 * batch).t().contiguous().view(batch, 32, 32)
ref = None
for mode in modes:
    ms, a, ipiv, info = run(a0, mode, 5)
    print(f"mode {mode}: {ms:.3f} ms  {batch * 16512 / ms * 1e-6:.0f} GB/s = {batch * 16512 / ms * 1e-6 / 6454.3:.3f} of HBM peak", flush=True)
    if ref is None: ref = (a, ipiv, info)
    else: print(f"   identical to mode {modes[0]}: factors {same(a, ref[0])} ipiv {same(ipiv, ref[1])} info {same(info, ref[2])}", flush=True)
L.lb200_set_batched_mode(2)
