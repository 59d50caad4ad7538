"""GPU probe #1: FP64 pipe peaks, library baselines (cuBLAS/cuSOLVER through torch) and the first
correctness + speed check of the DMMA GEMM.  Run on the B200 box:  python tools/gpu_probe1.py
Writes gpurun_out/probe1.json."""
import ctypes as C
import json
import os
import subprocess
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
L = C.CDLL(os.path.join(ROOT, "lapack_b200", "liblapack_b200.so"))
L.lb200_fp64_peak_tflops.restype = C.c_double
L.lb200_fp64_peak_tflops.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int]
L.lb200_dgemm.argtypes = [C.c_void_p, C.c_char, C.c_char, C.c_int, C.c_int, C.c_int, C.c_double, C.c_void_p,
                          C.c_longlong, C.c_void_p, C.c_longlong, C.c_double, C.c_void_p, C.c_longlong]
L.lb200_dsyrk.argtypes = [C.c_void_p, C.c_char, C.c_char, C.c_int, C.c_int, C.c_double, C.c_void_p, C.c_longlong,
                          C.c_double, C.c_void_p, C.c_longlong]
out = {}
dev = torch.device("cuda:0")
print(torch.cuda.get_device_name(0))
out["gpu"] = torch.cuda.get_device_name(0)
print(subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,clocks.max.sm,clocks.mem,power.draw,power.limit",
                      "--format=csv"], capture_output=True, text=True).stdout)


def stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def timeit(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e-3)
    return min(ts), sorted(ts)[len(ts) // 2]


# ---------------------------------------------------------------- 1. FP64 pipe peaks
peaks = {}
for kind, name in ((0, "dmma"), (1, "dfma")):
    for w, c in ((4, 1), (8, 1), (4, 2), (16, 1), (8, 2), (16, 2), (32, 1)):
        v = max(L.lb200_fp64_peak_tflops(stream(), kind, w, c, 20000) for _ in range(3))
        peaks[f"{name}_w{w}_c{c}"] = v
        print(f"peak {name} warps/cta={w} ctas/sm={c}: {v:.2f} TFLOP/s")
out["peaks"] = peaks
print(subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active",
                      "--format=csv"], capture_output=True, text=True).stdout)

# ---------------------------------------------------------------- 2. cuBLAS DGEMM through torch
lib = {}
for (m, n, k) in ((8192, 8192, 8192), (16384, 16384, 512), (16384, 16384, 256), (16384, 16384, 128), (32768, 32768, 512)):
    a = torch.randn(k, m, device=dev, dtype=torch.float64)
    b = torch.randn(n, k, device=dev, dtype=torch.float64)
    c = torch.empty(n, m, device=dev, dtype=torch.float64)
    best, med = timeit(lambda: torch.matmul(b, a, out=c))
    lib[f"cublas_{m}x{n}x{k}"] = 2.0 * m * n * k / best * 1e-12
    print(f"cuBLAS dgemm {m}x{n}x{k}: best {2.0*m*n*k/best*1e-12:.2f} TF/s  median {2.0*m*n*k/med*1e-12:.2f}")
    del a, b, c
out["cublas"] = lib


# ---------------------------------------------------------------- 3. my GEMM: correctness
def my_gemm(ta, tb, m, n, k, alpha, A, lda, B, ldb, beta, Cm, ldc):
    r = L.lb200_dgemm(stream(), ta.encode(), tb.encode(), m, n, k, alpha, A.data_ptr(), lda, B.data_ptr(), ldb, beta,
                      Cm.data_ptr(), ldc)
    assert r == 0, r


def colmajor(rows, cols, ld=None, off=0):
    """random col-major rows x cols matrix inside a (cols, ld) torch buffer; returns (buffer, view rows x cols)"""
    ld = ld or rows
    buf = torch.randn(cols * ld + off + 8, device=dev, dtype=torch.float64)
    view = torch.as_strided(buf, (rows, cols), (1, ld), off)
    return buf, view


errs = []
torch.manual_seed(1)
cases = [(128, 128, 64), (256, 384, 128), (129, 131, 17), (1, 1, 1), (5, 300, 3), (300, 5, 33), (64, 64, 4096),
         (1000, 777, 515), (127, 255, 16), (2, 2, 2)]
for cfg in (0, 1, 2):
    L.lb200_set_gemm_config(cfg)
    for ta in "NT":
        for tb in "NT":
            for (m, n, k) in cases:
                for (pad, off) in ((0, 0), (3, 1), (2, 0)):
                    ar, ac = (m, k) if ta == "N" else (k, m)
                    br, bc = (k, n) if tb == "N" else (n, k)
                    abuf, A = colmajor(ar, ac, ar + pad, off)
                    bbuf, B = colmajor(br, bc, br + pad, off)
                    cbuf, Cm = colmajor(m, n, m + pad, off)
                    cbuf0 = cbuf.clone()
                    alpha, beta = -1.0, 1.0
                    if (m + n + k) % 3 == 0:
                        alpha, beta = 0.7, 1.3
                    if (m + n + k) % 5 == 0:
                        beta = 0.0
                        Cm.fill_(float("nan"))
                        cbuf0 = cbuf.clone()
                    C0 = torch.as_strided(cbuf0, (m, n), (1, m + pad), off)
                    opA = A if ta == "N" else A.t()
                    opB = B if tb == "N" else B.t()
                    ref = alpha * (opA @ opB) + (beta * C0 if beta != 0.0 else 0.0)
                    my_gemm(ta, tb, m, n, k, alpha, A[0:1, 0:1], ar + pad, B[0:1, 0:1], br + pad, beta, Cm[0:1, 0:1], m + pad)
                    torch.cuda.synchronize()
                    scale = (opA.abs() @ opB.abs()).max().item() + (abs(beta) * C0.abs().max().item() if beta != 0 else 0) + 1e-300
                    err = (Cm - ref).abs().max().item() / scale / 2.2e-16
                    # padding must be untouched
                    mask = torch.ones_like(cbuf, dtype=torch.bool)
                    torch.as_strided(mask, (m, n), (1, m + pad), off).fill_(False)
                    pad_ok = bool(torch.equal(cbuf[mask], cbuf0[mask]) or (cbuf[mask].isnan() == cbuf0[mask].isnan()).all())
                    if not (err < 16.0) or not pad_ok:
                        errs.append((cfg, ta, tb, m, n, k, pad, off, err, pad_ok))
print("GEMM correctness failures:", len(errs))
for e in errs[:20]:
    print("  FAIL", e)
out["gemm_failures"] = len(errs)

# syrk check (triangle only)
syrk_fail = 0
for uplo in "LU":
    for tr in "NT":
        for (n, k) in ((200, 64), (129, 33), (512, 512)):
            ar, ac = (n, k) if tr == "N" else (k, n)
            abuf, A = colmajor(ar, ac)
            cbuf, Cm = colmajor(n, n, n + 2)
            C0 = Cm.clone()
            r = L.lb200_dsyrk(stream(), uplo.encode(), tr.encode(), n, k, -1.0, A.data_ptr(), ar, 1.0, Cm.data_ptr(), n + 2)
            torch.cuda.synchronize()
            opA = A if tr == "N" else A.t()
            full = C0 - opA @ opA.t()
            tri = torch.tril if uplo == "L" else torch.triu
            other = (lambda x: torch.triu(x, 1)) if uplo == "L" else (lambda x: torch.tril(x, -1))
            e1 = (tri(Cm) - tri(full)).abs().max().item()
            e2 = (other(Cm) - other(C0)).abs().max().item()
            if e1 > 1e-11 or e2 != 0.0:
                syrk_fail += 1
                print("  SYRK FAIL", uplo, tr, n, k, e1, e2)
print("SYRK failures:", syrk_fail)
out["syrk_failures"] = syrk_fail

# ---------------------------------------------------------------- 4. my GEMM: speed
perf = {}
for (m, n, k) in ((8192, 8192, 8192), (16384, 16384, 512), (16384, 16384, 256), (16384, 16384, 128), (32768, 32768, 512),
                  (32768, 256, 256), (4096, 4096, 512)):
    for cfg in (0, 1, 2):
        if cfg == 2 and m * n > 16384 * 16384:
            continue
        L.lb200_set_gemm_config(cfg)
        for (ta, tb) in (("N", "N"), ("N", "T"), ("T", "N")):
            if (ta, tb) != ("N", "N") and (m, n, k) not in ((16384, 16384, 512), (8192, 8192, 8192)):
                continue
            ar, ac = (m, k) if ta == "N" else (k, m)
            br, bc = (k, n) if tb == "N" else (n, k)
            A = torch.randn(ac, ar, device=dev, dtype=torch.float64)
            B = torch.randn(bc, br, device=dev, dtype=torch.float64)
            Cm = torch.randn(n, m, device=dev, dtype=torch.float64)
            best, med = timeit(lambda: my_gemm(ta, tb, m, n, k, -1.0, A, ar, B, br, 1.0, Cm, m), reps=4, warm=1)
            tf = 2.0 * m * n * k / best * 1e-12
            perf[f"cfg{cfg}_{ta}{tb}_{m}x{n}x{k}"] = tf
            print(f"my dgemm cfg{cfg} {ta}{tb} {m}x{n}x{k}: best {tf:.2f} TF/s (median {2.0*m*n*k/med*1e-12:.2f})")
            del A, B, Cm
L.lb200_set_gemm_config(-1)
out["gemm_perf"] = perf

# ---------------------------------------------------------------- 5. cuSOLVER context numbers through torch
ctx = {}
for n in (8192, 16384, 32768):
    a = torch.rand(n, n, device=dev, dtype=torch.float64) * 2 - 1
    best, _ = timeit(lambda: torch.linalg.lu_factor(a), reps=2, warm=1)
    ctx[f"cusolver_getrf_{n}"] = (2.0 * n**3 / 3) / best * 1e-12
    print(f"cuSOLVER getrf n={n}: {best*1e3:.1f} ms  {(2.0*n**3/3)/best*1e-12:.2f} TF/s")
    s = (a + a.t()) * 0.5 + n * torch.eye(n, device=dev, dtype=torch.float64)
    best, _ = timeit(lambda: torch.linalg.cholesky(s), reps=2, warm=1)
    ctx[f"cusolver_potrf_{n}"] = (n**3 / 3) / best * 1e-12
    print(f"cuSOLVER potrf n={n}: {best*1e3:.1f} ms  {(n**3/3)/best*1e-12:.2f} TF/s")
    del s
    if n <= 16384:
        best, _ = timeit(lambda: torch.geqrf(a), reps=2, warm=1)
        ctx[f"cusolver_geqrf_{n}"] = (4.0 * n**3 / 3) / best * 1e-12
        print(f"cuSOLVER geqrf n={n}: {best*1e3:.1f} ms  {(4.0*n**3/3)/best*1e-12:.2f} TF/s")
    del a
    torch.cuda.empty_cache()
out["cusolver"] = ctx

# ---------------------------------------------------------------- 6. PCIe
h = torch.empty(1 << 28, dtype=torch.float64).pin_memory()       # 2 GiB
d = torch.empty(1 << 28, dtype=torch.float64, device=dev)
best, _ = timeit(lambda: d.copy_(h, non_blocking=True), reps=3, warm=1)
out["h2d_gbs"] = h.numel() * 8 / best * 1e-9
best2, _ = timeit(lambda: h.copy_(d, non_blocking=True), reps=3, warm=1)
out["d2h_gbs"] = h.numel() * 8 / best2 * 1e-9
print(f"H2D {out['h2d_gbs']:.1f} GB/s   D2H {out['d2h_gbs']:.1f} GB/s (pinned)")
print(subprocess.run(["nproc"], capture_output=True, text=True).stdout, subprocess.run(["free", "-g"], capture_output=True, text=True).stdout)

json.dump(out, open(os.path.join(ROOT, "gpurun_out", "probe1.json"), "w"), indent=1)
print("done")
