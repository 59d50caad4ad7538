import ctypes as C, os, sys, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import lapack_b200 as lb
from oracle import oracle as O
L = lb.lib(); L.lb200_set_xerbla_mode(2)
vp = lambda x: x.ctypes.data_as(C.c_void_p)
n, nrhs = 250, 3
a, seed = O.random_matrix(n, n); x, _ = O.random_matrix(n, nrhs, seed); b = a @ x
ref = a.copy(order="F"); pr, _ = O.dgetrf(ref)
xr = np.asfortranarray(b.copy()); O.dgetrs("N", ref, pr, xr)
abuf = np.ascontiguousarray(ref)          # row-major LU
for nr in (3, 4, 1, 8):
    bb = np.ascontiguousarray(b[:, :min(nr, 3)]) if nr <= 3 else np.ascontiguousarray(np.hstack([b, b[:, :nr-3]]))
    r = L.LAPACKE_dgetrs_work(101, C.c_char(b"N"), n, bb.shape[1], vp(abuf), n, vp(pr), vp(bb), bb.shape[1])
    want = np.hstack([xr, xr[:, :max(0, nr-3)]])[:, :bb.shape[1]]
    print("nrhs", bb.shape[1], "rc", r, "err", np.max(np.abs(bb - want)))
# square B to see if it is the B transposition
bb = np.ascontiguousarray(np.hstack([b] * 84)[:, :250])
r = L.LAPACKE_dgetrs_work(101, C.c_char(b"N"), n, 250, vp(abuf), n, vp(pr), vp(bb), 250)
print("nrhs 250 rc", r, "err", np.max(np.abs(bb[:, :3] - xr)))
# row-major with padded lda/ldb
ap = np.zeros((n, n + 6)); ap[:, :n] = ref
bp = np.zeros((n, 8)); bp[:, :3] = b
r = L.LAPACKE_dgetrs_work(101, C.c_char(b"N"), n, 3, vp(ap), n + 6, vp(pr), vp(bp), 8)
print("padded rc", r, "err", np.max(np.abs(bp[:, :3] - xr)))
