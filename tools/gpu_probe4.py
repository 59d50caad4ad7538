"""GPU probe #4: GEMM tile-shape sweep (more warps per SM) and LU n=32768 with non-persistent trailing GEMM."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import lapack_b200 as lb
L = lb.lib()
dev = torch.device("cuda:0")
def timeit(fn, reps=4, warm=1):
    for _ in range(warm): fn()
    torch.cuda.synchronize(); ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1) * 1e-3)
    return min(ts)
for (m, n, k) in ((16384, 16384, 512), (8192, 8192, 8192), (16384, 16384, 256), (32768, 256, 256), (4096, 4096, 512)):
    for cfg in (6, 7, 8, 9, 10, 11):
        L.lb200_set_gemm_config(cfg)
        A = lb.dev.colmajor(m, k); A.normal_(); B = lb.dev.colmajor(k, n); B.normal_(); Cm = lb.dev.colmajor(m, n); Cm.normal_()
        ref = None
        if (m, n, k) == (4096, 4096, 512):
            ref = Cm - A @ B
        t = timeit(lambda: lb.dev.gemm("N", "N", -1.0, A, B, 1.0, Cm), reps=3, warm=0 if ref is not None else 1)
        err = ""
        if ref is not None:
            Cm.copy_(ref + A @ B); lb.dev.gemm("N", "N", -1.0, A, B, 1.0, Cm); torch.cuda.synchronize()
            err = f" maxerr={(Cm-ref).abs().max().item():.2e}"
        print(f"cfg{cfg} NN {m}x{n}x{k}: {2.0*m*n*k/t*1e-12:.2f} TF/s{err}", flush=True)
        del A, B, Cm
L.lb200_set_gemm_config(-1)
n = 32768
a0 = lb.dev.larnv_matrix(n, n)
a = a0.clone()
for cfg in (6, 7):
    L.lb200_set_gemm_config(cfg)
    for rep in range(2):
        a.copy_(a0); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ipiv, info = lb.dev.getrf(a); e1.record(); torch.cuda.synchronize()
        t = e0.elapsed_time(e1) * 1e-3
    print(f"DGETRF n={n} gemm cfg {cfg}: {t*1e3:.1f} ms {(2*n**3/3)/t*1e-12:.2f} TF/s", flush=True)
L.lb200_set_getrf_params(512, 0, 0)
L.lb200_set_gemm_config(-1)
a.copy_(a0); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); ipiv, info = lb.dev.getrf(a); e1.record(); torch.cuda.synchronize()
print(f"DGETRF n={n} no look-ahead: {e0.elapsed_time(e1):.1f} ms", flush=True)
