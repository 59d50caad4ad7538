import ctypes as C, os, sys, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import lapack_b200 as lb
from oracle import oracle as O
L = lb.lib(); L.lb200_set_xerbla_mode(2)
vp = lambda x: x.ctypes.data_as(C.c_void_p)
n, nrhs = 250, 3
a, seed = O.random_matrix(n, n); x, _ = O.random_matrix(n, nrhs, seed); b = a @ x
for layout in (102, 101):
    abuf = np.ascontiguousarray(a) if layout == 101 else a.copy(order="F")
    ipiv = np.zeros(n, dtype=np.int32)
    print("getrf", L.LAPACKE_dgetrf(layout, n, n, vp(abuf), n, vp(ipiv)))
    ref = a.copy(order="F"); pr, _ = O.dgetrf(ref)
    print(" lu err", np.max(np.abs(abuf - ref)), "piv ok", np.array_equal(ipiv, pr))
    for tr in (b"N", b"T"):
        b2 = np.ascontiguousarray(b) if layout == 101 else np.asfortranarray(b)
        ldb = nrhs if layout == 101 else n
        r = L.LAPACKE_dgetrs(layout, C.c_char(tr), n, nrhs, vp(abuf), n, vp(ipiv), vp(b2), ldb)
        xr = np.asfortranarray(b.copy()); O.dgetrs(tr.decode(), ref, pr, xr)
        print(" layout", layout, tr, "rc", r, "err vs oracle", np.max(np.abs(b2 - xr)))
    # work variant
    b3 = np.ascontiguousarray(b) if layout == 101 else np.asfortranarray(b)
    r = L.LAPACKE_dgetrs_work(layout, C.c_char(b"N"), n, nrhs, vp(abuf), n, vp(ipiv), vp(b3), nrhs if layout == 101 else n)
    print(" work rc", r, np.max(np.abs(b3 - x)))

# --- localize: Fortran ABI dgetrs_ with device A/B and host / device ipiv
print("---- f77 dgetrs_ mixed pointers")
ref = a.copy(order="F"); pr, _ = O.dgetrf(ref)
xr = np.asfortranarray(b.copy()); O.dgetrs("N", ref, pr, xr)
dA = lb.dev.colmajor(n, n); dA.copy_(torch.from_numpy(ref))
for piv_mode in ("host", "device"):
    for rep in range(2):
        dB = lb.dev.colmajor(n, nrhs); dB.copy_(torch.from_numpy(b)); torch.cuda.synchronize()
        piv = pr if piv_mode == "host" else torch.from_numpy(pr).cuda()
        info = lb.f77.dgetrs("N", n, nrhs, dA.data_ptr(), n, piv if piv_mode == "host" else piv.data_ptr(), dB.data_ptr(), n)
        torch.cuda.synchronize()
        print(piv_mode, rep, "info", info, "err", np.max(np.abs(dB.cpu().numpy() - xr)))
# same through the device API
dB = lb.dev.colmajor(n, nrhs); dB.copy_(torch.from_numpy(b))
lb.dev.getrs("N", dA, torch.from_numpy(pr).cuda(), dB); torch.cuda.synchronize()
print("dev api err", np.max(np.abs(dB.cpu().numpy() - xr)))
# transposes alone
t = lb.dev.colmajor(3, 250); t.normal_()
tt = lb.dev.transpose(t); torch.cuda.synchronize()
print("transpose 3x250 err", (tt - t.t()).abs().max().item())
