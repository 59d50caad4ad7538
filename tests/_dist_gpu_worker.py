"""torchrun worker for tests/test_gpu_dist.py: distributed DGETRF on N GPUs vs the single-GPU factorization."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import lapack_b200 as lb  # noqa: E402
from lapack_b200.dist import BlockCyclic1D, GpuOps, fill_local_random, pgetrf  # noqa: E402
from lapack_b200.dist_check import randomized_residual  # noqa: E402


def main():
    n, nb = int(sys.argv[1]), int(sys.argv[2])
    rank, world, lrank = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lrank)
    dev = torch.device("cuda", lrank)
    dist.init_process_group("nccl", device_id=dev)
    desc = BlockCyclic1D(n, nb, world, rank)
    ops = GpuOps(dev)
    a0 = fill_local_random(ops, desc, device=dev)
    for la in (True, False):
        a = a0.clone()
        ipiv, info = pgetrf(ops, dist, desc, a, lookahead=la)
        torch.cuda.synchronize()
        res = randomized_residual(torch, dist, desc, a0, a, ipiv)
        # single-GPU reference on every rank (same DLARNV stream)
        full = lb.dev.larnv_matrix(n, n, device=dev)
        p1, i1 = lb.dev.getrf(full)
        cols = torch.tensor([desc.global_col(c) for c in range(desc.local_cols())], device=dev)
        same_piv = bool(np.array_equal(p1.cpu().numpy(), ipiv))
        diff = (full[:, cols] - a).abs().max().item() if len(cols) else 0.0
        ok = same_piv and info == 0 and int(i1.item()) == 0 and res < 30.0 and diff < 1e-9
        flag = torch.tensor([1 if ok else 0], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if rank == 0:
            print(f"DIST_RESULT lookahead={la} ok={int(flag.item())} resid={res:.3f} same_piv={same_piv} maxdiff={diff:.2e}", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
