"""torchrun worker for tests/test_gpu_dist.py and bench.py's pre-timing check: P x Q distributed DGETRF on N GPUs vs the
single-GPU factorization of the same DLARNV matrix (IPIV must be identical, factors equal up to summation order)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import lapack_b200 as lb  # noqa: E402
from lapack_b200.dist2d import BlockCyclic2D, GpuOps2D, Groups, fill_local_random_2d, pgetrf2d, randomized_residual_2d  # noqa: E402


def check_against_single_gpu(dev, desc, groups, ops, lookahead=True):
    """returns (ok, message); every rank factors the full matrix on its own GPU as the reference"""
    n = desc.n
    a0 = fill_local_random_2d(desc, device=dev)
    a = a0.clone()
    ipiv, info = pgetrf2d(ops, dist, desc, a, groups, lookahead=lookahead)
    torch.cuda.synchronize()
    res = randomized_residual_2d(torch, dist, desc, a0, a, ipiv)
    full = lb.dev.larnv_matrix(n, n, device=dev)
    p1, i1 = lb.dev.getrf(full)
    rows = torch.from_numpy(desc.global_rows()).to(dev)
    cols = torch.from_numpy(desc.global_cols()).to(dev)
    same_piv = bool(np.array_equal(p1.cpu().numpy(), ipiv))
    diff = (full[rows][:, cols] - a).abs().max().item() if a.numel() else 0.0
    ok = same_piv and info == 0 and int(i1.item()) == 0 and res < 30.0 and diff < 1e-9
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    return bool(flag.item()), f"resid={res:.3f} same_piv={same_piv} maxdiff={diff:.2e} info={info}"


def main():
    P, Q, n, nb = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
    rank, world, lrank = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    assert world == P * Q
    torch.cuda.set_device(lrank)
    dev = torch.device("cuda", lrank)
    dist.init_process_group("nccl", device_id=dev)
    desc = BlockCyclic2D(n, nb, P, Q, rank)
    ops = GpuOps2D(dev)
    groups = Groups(dist, desc)
    for la in (True, False):
        ok, msg = check_against_single_gpu(dev, desc, groups, ops, lookahead=la)
        if rank == 0:
            print(f"DIST2D_RESULT grid={P}x{Q} lookahead={la} ok={int(ok)} {msg}", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
