"""batched 32x32 DGETRF: one-shot vs persistent pipelined kernel.  python tools/bench_batched2.py [batch]"""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import lapack_b200 as lb
L = lb.lib()
batch = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
a0 = lb.dev.larnv_matrix(32, 32 * batch).t().contiguous().view(batch, 32, 32)
res = {}
for mode in (0, 1):
    L.lb200_set_batched_mode(mode)
    a = a0.clone(); best = 1e9
    for _ in range(5):
        a.copy_(a0); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ipiv, info = lb.dev.getrf_batched32(a); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    res[mode] = (a.clone(), ipiv.clone(), info.clone())
    print(f"mode {mode}: {best:.3f} ms  {batch * 16512 / best * 1e-6:.0f} GB/s", flush=True)
print("identical results:", bool(torch.equal(res[0][0], res[1][0]) and torch.equal(res[0][1], res[1][1]) and torch.equal(res[0][2], res[1][2])))
L.lb200_set_batched_mode(2)
