"""torchrun tool: time distributed DPOTRF / DGEQRF (block-column cyclic) at order N.  usage: dist_bench.py which N NB"""
import os, sys, time
import numpy as np, torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import lapack_b200 as lb
from lapack_b200.dist import BlockCyclic1D, GpuOps, fill_local_random, ppotrf, pgeqrf, pgetrf

which, n, nb = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
rank, world, lrank = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lrank)
dev = torch.device("cuda", lrank)
os.environ.setdefault("NCCL_DEBUG", "WARN")
dist.init_process_group("nccl", device_id=dev)
desc = BlockCyclic1D(n, nb, world, rank)
ops = GpuOps(dev)
a0 = fill_local_random(ops, desc, device=dev)
if which == "potrf":
    # diagonally dominant SPD without forming the transpose: only the lower triangle is referenced, so
    # A_lower = random lower part with n added on the diagonal
    cols = torch.tensor([desc.global_col(c) for c in range(desc.local_cols())], device=dev, dtype=torch.long)
    a0[cols, torch.arange(len(cols), device=dev)] += float(n)
a = a0.clone()
for it in range(3):
    a.copy_(a0)
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    if which == "potrf": info = ppotrf(ops, dist, desc, a)
    elif which == "geqrf": info = pgeqrf(ops, dist, desc, a)
    else: info = pgetrf(ops, dist, desc, a)[1]
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    fl = {"potrf": n**3 / 3, "geqrf": 4 * n**3 / 3, "getrf": 2 * n**3 / 3}[which]
    if rank == 0:
        print(f"{which} n={n} nb={nb} gpus={world}: {t.item():.1f} ms {fl / t.item() * 1e-9:.1f} TFLOP/s", flush=True)
dist.destroy_process_group()
