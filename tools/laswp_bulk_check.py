"""bit-for-bit comparison of the two DLASWP apply kernels (LDG vs cp.async.bulk) on one LU-like pivot panel"""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import lapack_b200 as lb
L = lb.lib()
m, n, npiv, k1 = 9000, 3001, 512, 1025
g = torch.Generator(device="cpu"); g.manual_seed(5)
piv = torch.zeros(k1 + npiv, dtype=torch.int32)
for t in range(npiv):
    i = k1 + t
    piv[i - 1] = int(torch.randint(i, m + 1, (1,), generator=g))
dp = piv.cuda()
a0 = lb.dev.larnv_matrix(m, n)
outs = []
for mode in (0, 1):
    L.lb200_set_laswp_bulk(mode)
    a = a0.clone()
    lb.dev.laswp(a, k1, k1 + npiv - 1, dp, 1)
    lb.dev.laswp(a, k1, k1 + npiv - 1, dp, -1)      # reverse application must restore the matrix
    b = a0.clone(); lb.dev.laswp(b, k1, k1 + npiv - 1, dp, 1)
    torch.cuda.synchronize()
    outs.append(b)
    print("mode", mode, "forward+reverse restores:", bool(torch.equal(a, a0)))
print("modes identical:", bool(torch.equal(outs[0], outs[1])))
L.lb200_set_laswp_bulk(0)
