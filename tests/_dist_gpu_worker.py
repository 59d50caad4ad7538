"""torchrun worker for tests/test_gpu_dist.py: distributed DGETRF on N GPUs vs the single-GPU factorization."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import lapack_b200 as lb  # noqa: E402
from lapack_b200.dist import BlockCyclic1D, GpuOps, fill_local_random, pgetrf, ppotrf, pgeqrf  # noqa: E402
from lapack_b200.dist_check import randomized_residual  # noqa: E402


def main():
    n, nb = int(sys.argv[1]), int(sys.argv[2])
    rank, world, lrank = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lrank)
    dev = torch.device("cuda", lrank)
    dist.init_process_group("nccl", device_id=dev)
    desc = BlockCyclic1D(n, nb, world, rank)
    ops = GpuOps(dev)
    which = sys.argv[3] if len(sys.argv) > 3 else "getrf"
    cols = torch.tensor([desc.global_col(c) for c in range(desc.local_cols())], device=dev, dtype=torch.long)
    if which in ("potrf", "geqrf"):
        # single-GPU reference on every rank (same DLARNV stream), distributed run on its block columns
        full0 = lb.dev.larnv_matrix(n, n, device=dev)
        if which == "potrf":
            lb.dev.make_spd(full0, float(n))
        for la in (True, False):
            a = lb.dev.colmajor(n, len(cols), device=dev)
            a.copy_(full0[:, cols])
            full = full0.clone()
            if which == "potrf":
                upper = torch.triu(torch.ones(n, n, dtype=torch.bool, device=dev), 1)
                a[upper[:, cols]] = -1.0e10                          # the upper triangle must not be referenced
                info = ppotrf(ops, dist, desc, a, lookahead=la)
                i1 = int(lb.dev.potrf("L", full).item())
                low = ~upper[:, cols]
                diff = ((full[:, cols] - a).abs() * low).max().item() if len(cols) else 0.0
                untouched = bool((a[upper[:, cols]] == -1.0e10).all().item())
                ok = info == 0 and i1 == 0 and diff < 1e-9 and untouched
                msg = f"info={info} untouched={untouched}"
            else:
                tau = pgeqrf(ops, dist, desc, a, lookahead=la)
                tau1 = lb.dev.geqrf(full).cpu().numpy()
                scale = full.abs().max().item()
                diff = (full[:, cols] - a).abs().max().item() / scale if len(cols) else 0.0
                dt = float(np.max(np.abs(tau - tau1)))
                ok = diff < 1e-10 and dt < 1e-10
                msg = f"tau_diff={dt:.2e}"
            torch.cuda.synchronize()
            flag = torch.tensor([1 if ok else 0], device=dev)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            if rank == 0:
                print(f"DIST_RESULT {which} lookahead={la} ok={int(flag.item())} maxdiff={diff:.2e} {msg}", flush=True)
        dist.destroy_process_group()
        return
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
