"""GPU probe #3: TMA GEMM correctness (all transposes, edges, tri mode) and speed vs the cp.async kernel."""
import ctypes as C, os, sys, torch
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

def cm(rows, cols, ld=None):
    ld = ld or rows + (rows & 1)
    buf = torch.randn(cols * ld + 16, device=dev, dtype=torch.float64)
    return buf, torch.as_strided(buf, (rows, cols), (1, ld), 0)

fails = 0
L.lb200_set_gemm_config(12)
for ta in "NT":
    for tb in "NT":
        for (m, n, k) in ((128, 128, 64), (256, 384, 128), (129, 131, 33), (1000, 777, 515), (4096, 300, 48), (5, 300, 35), (2048, 2048, 1000)):
            ar, ac = (m, k) if ta == "N" else (k, m)
            br, bc = (k, n) if tb == "N" else (n, k)
            _, A = cm(ar, ac); _, B = cm(br, bc); cb, Cm = cm(m, n)
            for (alpha, beta) in ((-1.0, 1.0), (0.7, 0.0)):
                C0 = Cm.clone(); cb0 = cb.clone()
                opA = A if ta == "N" else A.t(); opB = B if tb == "N" else B.t()
                ref = alpha * (opA @ opB) + beta * C0
                lb.dev.gemm(ta, tb, alpha, A, B, beta, Cm)
                torch.cuda.synchronize()
                scale = (opA.abs() @ opB.abs()).max().item() + C0.abs().max().item()
                err = (Cm - ref).abs().max().item() / scale / 2.2e-16
                mask = torch.ones_like(cb, dtype=torch.bool)
                torch.as_strided(mask, Cm.shape, Cm.stride(), 0).fill_(False)
                pad_ok = bool(torch.equal(cb[mask], cb0[mask]))
                if not (err < 16) or not pad_ok:
                    fails += 1; print("FAIL", ta, tb, m, n, k, alpha, beta, err, pad_ok)
                Cm.copy_(C0)
# syrk through the TMA kernel
for uplo in "LU":
    for tr in "NT":
        n, k = 1500, 256
        ar, ac = (n, k) if tr == "N" else (k, n)
        _, A = cm(ar, ac); cb, Cm = cm(n, n); C0 = Cm.clone()
        lb.dev.syrk(uplo, tr, -1.0, A, 1.0, Cm); torch.cuda.synchronize()
        opA = A if tr == "N" else A.t(); full = C0 - opA @ opA.t()
        tri = torch.tril if uplo == "L" else torch.triu
        oth = (lambda x: torch.triu(x, 1)) if uplo == "L" else (lambda x: torch.tril(x, -1))
        e1 = (tri(Cm) - tri(full)).abs().max().item(); e2 = (oth(Cm) - oth(C0)).abs().max().item()
        if e1 > 1e-10 or e2 != 0: fails += 1; print("SYRK FAIL", uplo, tr, e1, e2)
print("TMA gemm failures:", fails)

for (m, n, k) in ((8192, 8192, 8192), (16384, 16384, 512), (16384, 16384, 256), (16384, 16384, 128), (32768, 32768, 512), (8192, 8192, 512), (4096, 4096, 512)):
    for cfg in (12, 8):
        L.lb200_set_gemm_config(cfg)
        for (ta, tb) in (("N", "N"), ("N", "T"), ("T", "N")):
            if (ta, tb) != ("N", "N") and (m, n, k) != (16384, 16384, 512): continue
            ar, ac = (m, k) if ta == "N" else (k, m); br, bc = (k, n) if tb == "N" else (n, k)
            _, A = cm(ar, ac); _, B = cm(br, bc); _, Cm = cm(m, n)
            t = timeit(lambda: lb.dev.gemm(ta, tb, -1.0, A, B, 1.0, Cm))
            print(f"cfg{cfg} {ta}{tb} {m}x{n}x{k}: {2.0*m*n*k/t*1e-12:.2f} TF/s", flush=True)
            del A, B, Cm
    a = torch.randn(k, m, device=dev, dtype=torch.float64); b = torch.randn(n, k, device=dev, dtype=torch.float64); c = torch.randn(n, m, device=dev, dtype=torch.float64)
    t = timeit(lambda: torch.addmm(c, b, a, alpha=-1.0, out=c))
    print(f"cublas NN {m}x{n}x{k}: {2.0*m*n*k/t*1e-12:.2f} TF/s", flush=True)
    del a, b, c
L.lb200_set_gemm_config(-1)
