"""One-call A/B of the round-2 experiments: general vs lean GEMM loader (cfg 8 / 13 / 14) and thin LU leaves.  python tools/r02_ab.py [n]
(the first-wave skew knob that the log profiles/r02_gemm_lean_ab.txt mentions was dropped from the library after this run)"""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import lapack_b200 as lb
L = lb.lib()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 32768

def t(fn, reps=5):
    fn(); torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best

if "nogemm" not in sys.argv:
    for (m, nn, k, ta, tb) in ((16384, 16384, 512, "N", "N"), (16384, 16384, 512, "N", "T"), (8192, 8192, 8192, "N", "N"), (24576, 8192, 512, "N", "N")):
        ar, ac = (m, k) if ta == "N" else (k, m); br, bc = (k, nn) if tb == "N" else (nn, k)
        A = lb.dev.colmajor(ar, ac); A.normal_(); B = lb.dev.colmajor(br, bc); B.normal_(); C0 = lb.dev.colmajor(m, nn); C0.normal_()
        ref = None
        for cfg, stg in ((8, 0), (13, 0), (14, 0)):
            L.lb200_set_gemm_config(cfg)
            C = C0.clone()
            lb.dev.gemm(ta, tb, -1.0, A, B, 1.0, C); torch.cuda.synchronize()
            if ref is None: ref = C.clone()
            same = bool(torch.equal(C, ref))
            ms = t(lambda: lb.dev.gemm(ta, tb, -1.0, A, B, 1.0, C))
            print(f"gemm {ta}{tb} {m}x{nn}x{k} cfg {cfg} stagger {stg}: {ms:.3f} ms {2*m*nn*k/ms*1e-9:.2f} TF/s  bitwise==cfg8: {same}", flush=True)
        del A, B, C, C0, ref
    L.lb200_set_gemm_config(-1)

a0 = lb.dev.larnv_matrix(n, n)
a = a0.clone()
def run(tag, reps=2):
    best = 1e9
    for _ in range(reps):
        a.copy_(a0); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); piv, info = lb.dev.getrf(a); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    print(f"LU n={n} {tag}: {best:.1f} ms  {(2*n**3/3)/best*1e-9:.2f} TFLOP/s", flush=True)
    return piv.clone(), a.clone()
p0, f0 = run("default", 3)
for cfg, stg in ((8, 0), (14, 0)):
    L.lb200_set_gemm_config(cfg)
    p, f = run(f"gemm cfg {cfg} stagger {stg}")
    print("  ipiv equal:", bool((p == p0).all()), " factors bitwise equal:", bool(torch.equal(f, f0)))
L.lb200_set_gemm_config(-1)
for mode in (1, 2):
    for mr in (16384, 12288, 8192, 4096):
        L.lb200_set_getrf_thin(mode, mr)
        p, f = run(f"thin mode {mode} min_rows {mr}")
        print("  ipiv equal:", bool((p == p0).all()), " factors bitwise equal:", bool(torch.equal(f, f0)))
L.lb200_set_getrf_thin(2, 12288); L.lb200_set_gemm_config(13)
p, f = run("thin 2/12288 + gemm cfg 13")
print("  ipiv equal:", bool((p == p0).all()))
L.lb200_set_getrf_thin(0, 16384); L.lb200_set_gemm_config(-1)
