import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import lapack_b200 as lb
L = lb.lib()
for n in (8192, 16384, 32768):
    a0 = lb.dev.larnv_matrix(n, n)
    a = a0.clone()
    for nb in (128, 256, 384, 512):
        L.lb200_set_geqrf_params(nb, 1)
        best = 1e30
        for _ in range(2):
            a.copy_(a0); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); lb.dev.geqrf(a); e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        print("geqrf", n, "nb", nb, f"{best:.2f} ms", f"{4/3*n**3/best*1e-9:.2f} TF/s", flush=True)
    del a, a0; torch.cuda.empty_cache()
