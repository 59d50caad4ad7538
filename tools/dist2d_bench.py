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
ops = GpuOps2D(dev, panel_stream=os.environ.get("SERIAL_PANEL", "0") != "1", panel_priority=int(os.environ.get("PANEL_PRIO", "-1")))
groups = Groups(dist, desc)
a0 = fill_local_random_2d(desc, device=dev)
a = a0.clone()
for it in range(reps):
    a.copy_(a0)
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    tr = [] if (it == reps - 1 and os.environ.get("TRACE")) else None
    ipiv, info = pgetrf2d(ops, dist, desc, a, groups, trace=tr)
    e1.record(); torch.cuda.synchronize()
    if tr is not None and rank == int(os.environ.get("TRACE_RANK", "0")):
        steps = [x for x in tr if x[0] == "step"]
        pans = {x[1]: x[2].elapsed_time(x[3]) for x in tr if x[0] == "panel"}
        nq = max(1, len(steps) // 8)
        print(f"TRACE rank {rank}: per eighth of the factorization: wait-for-panel ms | update ms | panel-stream ms (panels this rank took part in)")
        for o in range(0, len(steps), nq):
            seg = steps[o:o + nq]
            wait = sum(x[2].elapsed_time(x[3]) for x in seg)
            upd = sum(x[3].elapsed_time(x[4]) for x in seg)
            pan = sum(pans.get(x[1], 0.0) for x in seg)
            ks = {x[1] for x in seg}
            us = [x for x in tr if x[0] == "upd" and x[1] in ks]
            t_sw = sum(x[2].elapsed_time(x[3]) for x in us)
            t_tr = sum(x[3].elapsed_time(x[4]) for x in us)
            t_ge = sum(x[4].elapsed_time(x[5]) for x in us)
            fl = sum(x[6] for x in us)
            print(f"TRACE steps {seg[0][1]:4d}-{seg[-1][1]:4d}: wait {wait:8.1f} | update {upd:8.1f} (laswp {t_sw:7.1f} trsm {t_tr:7.1f} gemm {t_ge:8.1f} = "
                  f"{fl / max(t_ge, 1e-9) * 1e-9:5.1f} TFLOP/s) | panel {pan:8.1f}", flush=True)
    t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    fl = 2 * n ** 3 / 3
    if rank == 0:
        print(f"getrf2d grid={P}x{Q} n={n} nb={nb}: {t.item():.1f} ms {fl / t.item() * 1e-9:.1f} TFLOP/s ({fl / t.item() * 1e-9 / world:.2f} per GPU) info={info}", flush=True)
dist.destroy_process_group()
