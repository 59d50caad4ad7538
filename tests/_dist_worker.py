"""Worker for tests/test_dist_cpu.py: runs lapack_b200.dist.pgetrf over gloo on CPU tensors with a MOCK local
backend (the CPU oracle stands in for the CUDA kernels -- test infrastructure only) to exercise the index maps,
the look-ahead schedule and the broadcast protocol with world_size > 1."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402
from lapack_b200.dist import BlockCyclic1D, pgetrf, ppotrf, pgeqrf  # noqa: E402


class MockOps:
    def zeros(self, m, n):
        return torch.zeros((n, m), dtype=torch.float64).t()

    def empty_vec(self, n):
        return torch.zeros(n, dtype=torch.float64)

    def panel_factor(self, panel):
        a = panel.numpy()
        ipiv, info = O.dgetrf2(a)
        return torch.from_numpy(ipiv.copy()), torch.tensor([info], dtype=torch.int32)

    def laswp(self, a, k1, k2, ipiv):
        if a.shape[1] > 0:
            O.dlaswp(a.numpy(), k1, k2, np.ascontiguousarray(ipiv.numpy(), dtype=np.int32), 1)

    def trsm_llnu(self, l11, b):
        if b.shape[1] > 0:
            O.dtrsm("L", "L", "N", "U", b.shape[0], b.shape[1], 1.0, l11.numpy(), b.numpy())

    def gemm_update(self, l21, u12, c):
        if c.shape[0] > 0 and c.shape[1] > 0:
            O.dgemm("N", "N", c.shape[0], c.shape[1], u12.shape[0], -1.0, l21.numpy(), u12.numpy(), 1.0, c.numpy())

    def copy(self, dst, src):
        dst.copy_(src)

    def potrf_panel(self, panel):
        jb = panel.shape[1]
        a = panel.numpy()
        info = O.dpotrf("L", a[:jb, :])
        if a.shape[0] > jb:
            O.dtrsm("R", "L", "T", "N", a.shape[0] - jb, jb, 1.0, a[:jb, :], a[jb:, :])
        return torch.tensor([info], dtype=torch.int32)

    def syrk_update(self, l, c):
        O.dsyrk("L", "N", c.shape[0], l.shape[1], -1.0, l.numpy(), 1.0, c.numpy())

    def gemm_nt_update(self, a, b, c):
        if c.shape[0] > 0 and c.shape[1] > 0:
            O.dgemm("N", "T", c.shape[0], c.shape[1], a.shape[1], -1.0, a.numpy(), b.numpy(), 1.0, c.numpy())

    def qr_panel(self, panel):
        a = panel.numpy()
        tau, info, _ = O.dgeqrf(a)
        assert info == 0
        t = O.dlarft(np.asfortranarray(a[:, :len(tau)]), tau)
        return torch.from_numpy(tau.copy()), torch.from_numpy(np.ascontiguousarray(t.T)).t()

    def larfb_lt(self, v, t, c):
        if c.shape[1] > 0:
            O.dlarfb("L", "T", np.asfortranarray(v.numpy()), np.asfortranarray(t.numpy()), c.numpy())

    def to_int32(self, x):
        return x.to(torch.int32)

    def to_float64(self, x):
        return x.to(torch.float64)


def main():
    n, nb, lookahead, outdir = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]
    which = sys.argv[5] if len(sys.argv) > 5 else "getrf"
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dist.init_process_group("gloo", rank=rank, world_size=world)
    desc = BlockCyclic1D(n, nb, world, rank)
    if which == "potrf":
        a, _ = O.spd_matrix(n, (1988, 1989, 1990, 1991))
        a = np.tril(a) - 1.0e10 * np.triu(np.ones((n, n)), 1)        # the upper triangle must not be referenced
        if len(sys.argv) > 6:
            iz = int(sys.argv[6])                                     # not positive definite: zero row/column iz
            a[iz - 1, :iz] = 0.0
            a[iz - 1:, iz - 1] = 0.0
    else:
        a, _ = O.random_matrix(n, n, (1988, 1989, 1990, 1991))
        if which == "getrf" and len(sys.argv) > 6:
            a[:, int(sys.argv[6]) - 1] = 0.0                          # exactly singular: INFO = that column (dchkge.f:328-347)
    cols = np.array([desc.global_col(c) for c in range(desc.local_cols())], dtype=np.int64)
    ops = MockOps()
    aloc = ops.zeros(n, len(cols))
    aloc.copy_(torch.from_numpy(np.ascontiguousarray(a[:, cols])))
    out = dict(cols=cols)
    if which == "getrf":
        ipiv, info = pgetrf(ops, dist, desc, aloc, lookahead=bool(lookahead))
        out.update(ipiv=ipiv, info=info)
    elif which == "potrf":
        out.update(info=ppotrf(ops, dist, desc, aloc, lookahead=bool(lookahead)))
    else:
        out.update(tau=pgeqrf(ops, dist, desc, aloc, lookahead=bool(lookahead)))
    np.savez(os.path.join(outdir, f"rank{rank}.npz"), lu=aloc.numpy(), **out)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
