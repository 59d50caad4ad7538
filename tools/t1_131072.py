"""T1: ONE DGETRF of order n (default 131072 = BASELINE configs[4]'s matrix) on a single B200, so that the multi-GPU
parallel efficiency is like-for-like.  128 GiB matrix: generated in place (DLARNV jump-ahead), never
copied.  python tools/t1_131072.py [n] -> JSON line (also written to gpurun_out/r02_t1_dgetrf_<n>.json)"""
import json, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import lapack_b200 as lb
n = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
a = lb.dev.larnv_matrix(n, n)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); ipiv, info = lb.dev.getrf(a); e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
fl = 2.0 * n ** 3 / 3 - n ** 2 / 2 + 5.0 * n / 6
out = {"n": n, "ms": ms, "tflops": fl / (ms * 1e-3) * 1e-12, "info": int(info.item()), "runs": 1,
       "note": "single cold run (first call of the process): one factorization takes ~1 minute"}
print(json.dumps(out))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", f"r02_t1_dgetrf_{n}.json"), "w"))
