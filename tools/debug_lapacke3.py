import ctypes as C, os, sys, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import lapack_b200 as lb
from oracle import oracle as O
L = lb.lib(); L.lb200_set_xerbla_mode(2)
R = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "liblapacke_ref.so"), mode=C.RTLD_GLOBAL)
vp = lambda x: x.ctypes.data_as(C.c_void_p)
n, nrhs = 250, 3
a, seed = O.random_matrix(n, n); x, _ = O.random_matrix(n, nrhs, seed); b = a @ x
ref = a.copy(order="F"); pr, _ = O.dgetrf(ref)
xr = np.asfortranarray(b.copy()); O.dgetrs("N", ref, pr, xr)
abuf = np.ascontiguousarray(ref)
for name, lib_ in (("ref-lapacke", R), ("ours", L), ("ref-lapacke", R), ("ours", L), ("ours", L)):
    bb = np.ascontiguousarray(b)
    r = lib_.LAPACKE_dgetrs_work(101, C.c_char(b"N"), n, nrhs, vp(abuf), n, vp(pr), vp(bb), nrhs)
    print(name, "rc", r, "err", np.max(np.abs(bb - xr)), "abuf intact", np.array_equal(abuf, np.ascontiguousarray(ref)), "piv intact", np.array_equal(pr, O.dgetrf(a.copy(order='F'))[0]))
# python emulation of the glue
for rep in range(3):
    dA_rm = torch.from_numpy(abuf).cuda()              # (n,n) row-major == col-major A^T
    dAcm = lb.dev.transpose(dA_rm.t()[:n, :n]) if False else None
    A_rm_view = torch.as_strided(dA_rm, (n, n), (1, n))       # col-major view of the buffer = LU^T
    A_cm = lb.dev.transpose(A_rm_view)                        # col-major LU
    dB_rm = torch.from_numpy(np.ascontiguousarray(b)).cuda()  # (n,3) row-major = col-major 3 x n
    B_rm_view = torch.as_strided(dB_rm, (nrhs, n), (1, nrhs))
    B_cm = lb.dev.transpose(B_rm_view)                        # n x 3 col-major
    torch.cuda.synchronize()
    info = lb.f77.dgetrs("N", n, nrhs, A_cm.data_ptr(), lb.dev.ld(A_cm), pr, B_cm.data_ptr(), lb.dev.ld(B_cm))
    torch.cuda.synchronize()
    print("python glue rep", rep, "info", info, "err", np.max(np.abs(B_cm.cpu().numpy() - xr)))
