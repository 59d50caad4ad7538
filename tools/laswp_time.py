"""DLASWP apply timing: one panel's interchanges (np pivots, LU-like: ip uniform in [i, m]) on n columns of an m-row matrix.
Algorithmic bytes = 32 B per interchanged pair and column.  python tools/laswp_time.py [m] [n] [np]"""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import lapack_b200 as lb
if os.environ.get('BULK'): lb.lib().lb200_set_laswp_bulk(int(os.environ['BULK']))
if os.environ.get('L2G'): print('L2 fetch granularity ->', os.environ['L2G'], lb.lib().lb200_set_l2_fetch_granularity(int(os.environ['L2G'])))
m = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
n = int(sys.argv[2]) if len(sys.argv) > 2 else 32768
for npiv in ([int(sys.argv[3])] if len(sys.argv) > 3 else [64, 512, 2048]):
    a = lb.dev.larnv_matrix(m, n)
    g = torch.Generator(device="cpu"); g.manual_seed(3)
    k1 = 4096
    piv = torch.zeros(k1 + npiv, dtype=torch.int32)
    for t in range(npiv):
        i = k1 + t
        piv[i - 1] = int(torch.randint(i, m + 1, (1,), generator=g))
    dp = piv.cuda()
    moved = int((piv[k1 - 1:k1 - 1 + npiv] != torch.arange(k1, k1 + npiv, dtype=torch.int32)).sum())
    best = 1e9
    for rep in range(4):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); lb.dev.laswp(a, k1, k1 + npiv - 1, dp, 1); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    print(f"laswp m={m} n={n} np={npiv}: {best:.3f} ms, {32.0 * moved * n / (best * 1e-3) * 1e-9:.0f} GB/s algorithmic", flush=True)
