"""Worker for tests/test_dist2d_cpu.py: runs lapack_b200.dist2d.pgetrf2d over gloo on CPU tensors with a MOCK local backend
(the CPU oracle stands in for the CUDA kernels -- test infrastructure only) to exercise the P x Q index maps, the panel
gather / return / row broadcast, the cross-process-row interchanges, the U12 broadcast and the look-ahead schedule."""
import contextlib
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402
from lapack_b200.dist2d import BlockCyclic2D, Groups, pgetrf2d  # noqa: E402


class MockOps2D:
    torch = torch

    def empty_vec(self, n):
        return torch.zeros(n, dtype=torch.float64)

    def panel_factor(self, panel):
        a = np.asfortranarray(panel.numpy())
        ipiv, info = O.dgetrf2(a)
        panel.copy_(torch.from_numpy(a))
        return torch.from_numpy(ipiv.copy()), torch.tensor([info], dtype=torch.int32)

    def laswp(self, a, k1, k2, ipiv):
        if a.shape[1] > 0:
            x = np.asfortranarray(a.numpy())
            O.dlaswp(x, k1, k2, np.ascontiguousarray(ipiv.numpy(), dtype=np.int32), 1)
            a.copy_(torch.from_numpy(x))

    def trsm_llnu(self, l11, b):
        if b.shape[1] > 0:
            x = np.asfortranarray(b.numpy())
            O.dtrsm("L", "L", "N", "U", b.shape[0], b.shape[1], 1.0, np.asfortranarray(l11.numpy()), x)
            b.copy_(torch.from_numpy(x))

    def gemm_update(self, l21, u12, c):
        if c.shape[0] > 0 and c.shape[1] > 0:
            x = np.asfortranarray(c.numpy())
            O.dgemm("N", "N", c.shape[0], c.shape[1], u12.shape[0], -1.0, np.asfortranarray(l21.numpy()),
                    np.asfortranarray(u12.numpy()), 1.0, x)
            c.copy_(torch.from_numpy(x))

    def copy(self, dst, src):
        dst.copy_(src)

    def compose(self, piv):
        """sequential restatement of lb200_laswp_compose"""
        pv = piv.numpy()
        np_ = len(pv)
        cur = {}
        get = lambda r: cur.get(r, r)
        for t in range(np_):
            ip = int(pv[t]) - 1
            if ip != t:
                a, b = get(t), get(ip)
                cur[t], cur[ip] = b, a
        src_top = np.array([get(t) for t in range(np_)], dtype=np.int32)
        inv_top = np.full(np_, -1, dtype=np.int32)
        for r, o in list(cur.items()) + [(t, get(t)) for t in range(np_)]:
            if o < np_:
                inv_top[o] = r
        return torch.from_numpy(src_top), torch.from_numpy(inv_top)

    def gather_rows(self, a, idx, w):
        ix = idx.numpy()
        for t in np.nonzero(ix >= 0)[0]:
            w[t, :] = a[int(ix[t]), :]

    def scatter_rows(self, w, idx, a):
        ix = idx.numpy()
        for t in np.nonzero(ix >= 0)[0]:
            a[int(ix[t]), :] = w[t, :]

    def panel_stream(self, rows=0):
        return contextlib.nullcontext()

    def fork_panel(self):
        pass

    def join_panel(self):
        pass


def main():
    P, Q, n, nb, lookahead, outdir = (int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5]),
                                      sys.argv[6])
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    assert world == P * Q
    dist.init_process_group("gloo", rank=rank, world_size=world)
    desc = BlockCyclic2D(n, nb, P, Q, rank)
    a, _ = O.random_matrix(n, n, (1988, 1989, 1990, 1991))
    if len(sys.argv) > 7:
        a[:, int(sys.argv[7]) - 1] = 0.0                                # exactly singular: INFO = that column
    rows, cols = desc.global_rows(), desc.global_cols()
    aloc = torch.zeros((len(cols), len(rows)), dtype=torch.float64).t()
    aloc.copy_(torch.from_numpy(np.ascontiguousarray(a[np.ix_(rows, cols)])))
    ops = MockOps2D()
    ipiv, info = pgetrf2d(ops, dist, desc, aloc, Groups(dist, desc), lookahead=bool(lookahead))
    np.savez(os.path.join(outdir, f"rank{rank}.npz"), lu=aloc.numpy(), rows=rows, cols=cols, ipiv=ipiv, info=info)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
