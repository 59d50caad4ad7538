"""torchrun tool: time the P x Q distributed DGETRF (lapack_b200/dist2d.py).  usage: dist2d_bench.py P Q N NB [reps]"""
import os, sys
import torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import lapack_b200 as lb
from lapack_b200.dist2d import BlockCyclic2D, GpuOps2D, Groups, fill_local_random_2d, pgetrf2d

P, Q, n, nb = (int(v) for v in sys.argv[1:5])
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 3
rank, world, lrank = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lrank)
dev = torch.device("cuda", lrank)
os.environ.setdefault("NCCL_DEBUG", "WARN")
dist.init_process_group("nccl", device_id=dev)
desc = BlockCyclic2D(n, nb, P, Q, rank)
ops = GpuOps2D(dev)
groups = Groups(dist, desc)
a0 = fill_local_random_2d(desc, device=dev)
a = a0.clone()
for it in range(reps):
    a.copy_(a0)
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ipiv, info = pgetrf2d(ops, dist, desc, a, groups)
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    fl = 2 * n ** 3 / 3
    if rank == 0:
        print(f"getrf2d grid={P}x{Q} n={n} nb={nb}: {t.item():.1f} ms {fl / t.item() * 1e-9:.1f} TFLOP/s ({fl / t.item() * 1e-9 / world:.2f} per GPU) info={info}", flush=True)
dist.destroy_process_group()
