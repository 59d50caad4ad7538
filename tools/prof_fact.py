"""One factorization for profiling: python tools/prof_fact.py {getrf|potrf|geqrf} N [reps]"""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import lapack_b200 as lb
if os.environ.get("LB200_GEMM_CFG"): lb.lib().lb200_set_gemm_config(int(os.environ["LB200_GEMM_CFG"]))
which, n = sys.argv[1], int(sys.argv[2])
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 1
a0 = lb.dev.larnv_matrix(n, 512 if which == "panel" else n)
if which == "potrf":
    lb.dev.make_spd(a0, float(n))
a = a0.clone()
for _ in range(reps):
    a.copy_(a0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    if which == "getrf": lb.dev.getrf(a)
    elif which == "panel": lb.dev.getrf(a, recursive=True)
    elif which == "potrf": lb.dev.potrf("L", a)
    else: lb.dev.geqrf(a)
    e1.record(); torch.cuda.synchronize()
    print(which, n, "ms", e0.elapsed_time(e1), flush=True)
