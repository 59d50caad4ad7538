"""Multi-GPU DGETRF: one large matrix partitioned block-cyclically over the GPUs of one node (BASELINE config C5a).

Layout: 2D block-cyclic with a 1 x P process grid (SURVEY 8e allows any P x Q; with Q = #GPUs every GPU holds
whole block columns, so the panel -- pivot search included -- is factored by its owner without any cross-GPU
exchange, and the only data-path collective is the panel broadcast): global block column k (NB columns) lives
on rank k % P as local block k // P.  Per step k:

    owner(k)  : panel k already factored (look-ahead) and packed into a broadcast buffer
    all ranks : receive panel k (L11, L21, pivots, info) by NCCL broadcast over NVLink
                row interchanges on every local column (SRC/dgetrf.f:193,199), U12 = L11^-1 A12 (dgetrf.f:204),
                A22 -= L21 U12 (dgetrf.f:212) on the local trailing columns
    owner(k+1): updates its block k+1 FIRST, factors it (recursive panel, getrf.cu), packs it and starts the
                broadcast of panel k+1 while everybody (itself included) is still busy with update k.

There is no reference counterpart (ScaLAPACK is a separate project); results are identical to the single-GPU
factorization up to the summation order of the trailing GEMMs, and IPIV is the same (tested).

torch / torch.distributed are plumbing only (device memory, streams, the NCCL broadcast); all arithmetic goes
through the C ABI (`lb200_*`).  The local compute backend is a small object so that the communication schedule
and index maps can be exercised on CPU with gloo by the tests (tests/test_dist_cpu.py supplies a mock backend
there; the product ships only `GpuOps`).
"""
from __future__ import annotations

from dataclasses import dataclass


@dataclass
class BlockCyclic1D:
    """Block-column cyclic distribution of an n x n matrix over `world` ranks with block width nb."""
    n: int
    nb: int
    world: int
    rank: int

    @property
    def nblocks(self) -> int:
        return (self.n + self.nb - 1) // self.nb

    def owner(self, k: int) -> int:
        return k % self.world

    def width(self, k: int) -> int:
        return min(self.nb, self.n - k * self.nb)

    def local_blocks(self):
        return list(range(self.rank, self.nblocks, self.world))

    def local_cols(self) -> int:
        return sum(self.width(k) for k in self.local_blocks())

    def local_offset(self, k: int) -> int:
        """first local column of global block k (k must be local)"""
        assert self.owner(k) == self.rank
        return (k // self.world) * self.nb

    def first_local_col_after(self, k: int) -> int:
        """local column where the blocks with global index > k start"""
        cnt = 0 if k < self.rank else (k - self.rank) // self.world + 1
        return min(cnt * self.nb, self.local_cols())

    def local_cols_before(self, k: int) -> int:
        """number of local columns belonging to blocks with global index < k"""
        cnt = 0 if k - 1 < self.rank else (k - 1 - self.rank) // self.world + 1
        return min(cnt * self.nb, self.local_cols())

    def global_col(self, local_col: int) -> int:
        lk, off = divmod(local_col, self.nb)
        return (lk * self.world + self.rank) * self.nb + off


class GpuOps:
    """Local compute through the lapack_b200 C ABI on torch CUDA tensors (column-major views)."""

    def __init__(self, device):
        import torch
        from . import dev
        self.torch = torch
        self.dev = dev
        self.device = device

    def zeros(self, m, n):
        a = self.dev.colmajor(m, n, device=self.device)
        a.zero_()
        return a

    def empty_vec(self, n):
        return self.torch.empty(n, dtype=self.torch.float64, device=self.device)

    def panel_factor(self, panel):
        """in-place recursive LU of a tall panel; returns (ipiv int32 relative 1-based, info) as device tensors"""
        return self.dev.getrf(panel, recursive=True)

    def laswp(self, a, k1, k2, ipiv):
        if a.shape[1] > 0:
            self.dev.laswp(a, k1, k2, ipiv, 1)

    def trsm_llnu(self, l11, b):
        if b.shape[1] > 0:
            self.dev.trsm("L", "L", "N", "U", 1.0, l11, b)

    def gemm_update(self, l21, u12, c):
        if c.shape[0] > 0 and c.shape[1] > 0:
            self.dev.gemm("N", "N", -1.0, l21, u12, 1.0, c)

    def copy(self, dst, src):
        dst.copy_(src)

    def to_int32(self, x):
        return x.to(self.torch.int32)

    def to_float64(self, x):
        return x.to(self.torch.float64)

    # ---- Cholesky (lower)
    def potrf_panel(self, panel):
        """panel = A(j:n, j:j+jb): Cholesky of the jb x jb diagonal block, then L21 = A21 L11^-T; returns device INFO"""
        jb = panel.shape[1]
        info = self.dev.potrf("L", panel[:jb, :])
        if panel.shape[0] > jb:
            self.dev.trsm("R", "L", "T", "N", 1.0, panel[:jb, :], panel[jb:, :])
        return info

    def syrk_update(self, l, c):
        """c (w x w, lower triangle) -= l l^T"""
        self.dev.syrk("L", "N", -1.0, l, 1.0, c)

    def gemm_nt_update(self, a, b, c):
        """c -= a b^T"""
        if c.shape[0] > 0 and c.shape[1] > 0:
            self.dev.gemm("N", "T", -1.0, a, b, 1.0, c)

    # ---- QR
    def qr_panel(self, panel):
        """in-place DGEQRF of a tall panel; returns (tau, T) with T the jb x jb triangular factor of the block reflector"""
        tau = self.dev.geqrf(panel)
        k = tau.shape[0]
        t = self.dev.larft(panel[:, :k], tau)
        return tau, t

    def larfb_lt(self, v, t, c):
        """c := H^T c with H = I - V T V^T (V unit lower trapezoidal, stored below the diagonal of v)"""
        if c.shape[1] > 0:
            self.dev.larfb("L", "T", v, t, c)


def pgetrf(ops, dist, desc: BlockCyclic1D, aloc, lookahead: bool = True):
    """Distributed LU with partial pivoting of the block-column-cyclic matrix `aloc` (n x local_cols, in place).

    Returns (ipiv, info): ipiv = global 1-based pivot rows (length n, replicated on every rank, same convention
    as DGETRF), info = 0 or the index of the first exactly-zero pivot.
    `dist` is torch.distributed (already initialised) or None for a single rank.
    """
    n, nb, me = desc.n, desc.nb, desc.rank
    nblk = desc.nblocks
    ipiv_all = []
    info = 0
    # broadcast buffer: panel (n x nb, ld = n) + nb pivots + info, double buffered
    bufs = [ops.empty_vec(n * nb + nb + 8) for _ in range(2)]

    def panel_view(buf, rows, jb):
        return buf[: n * nb].view(nb, n).t()[:rows, :jb]      # column-major (rows x jb), ld = n

    def factor_and_pack(k, buf):
        """owner only: factor local block k (rows j..n) and pack L, pivots, info into buf"""
        j, jb = k * nb, desc.width(k)
        lo = desc.local_offset(k)
        panel = aloc[j:, lo:lo + jb]
        piv, inf = ops.panel_factor(panel)
        ops.copy(panel_view(buf, n - j, jb), panel)
        tail = buf[n * nb: n * nb + nb + 8]
        tail.zero_()
        ops.copy(tail[: piv.shape[0]], ops.to_float64(piv))
        ops.copy(tail[nb: nb + 1], ops.to_float64(inf))

    def start_bcast(k, buf):
        if dist is None or desc.world == 1:
            return None
        return dist.broadcast(buf, src=desc.owner(k), async_op=True)

    # prologue: panel 0
    if desc.owner(0) == me:
        factor_and_pack(0, bufs[0])
    work = start_bcast(0, bufs[0])

    for k in range(nblk):
        cur, nxt = bufs[k % 2], bufs[(k + 1) % 2]
        j, jb = k * nb, desc.width(k)
        jn = j + jb
        if work is not None:
            work.wait()
        tail = cur[n * nb: n * nb + nb + 8]
        piv = ops.to_int32(tail[:jb])                      # relative to row j, 1-based
        ipiv_all.append((j, piv, tail[nb: nb + 1] * 1.0))       # a copy: the buffer is reused two steps later
        pv = panel_view(cur, n - j, jb)                    # rows j..n-1 of panel k
        l11, l21 = pv[:jb, :], pv[jb:, :]

        c_after = desc.first_local_col_after(k)            # local trailing columns start here
        c_before = desc.local_cols_before(k)               # local columns left of the panel
        nloc = aloc.shape[1]
        work = None
        next_k = k + 1
        i_own_next = next_k < nblk and desc.owner(next_k) == me

        def update(c0, c1):
            if c1 <= c0:
                return
            ops.laswp(aloc[j:, c0:c1], 1, jb, piv)                         # dgetrf.f:199
            ops.trsm_llnu(l11, aloc[j:jn, c0:c1])                          # dgetrf.f:204
            if jn < n:
                ops.gemm_update(l21, aloc[j:jn, c0:c1], aloc[jn:, c0:c1])  # dgetrf.f:212

        if next_k < nblk:
            if i_own_next:
                w = desc.width(next_k)
                if lookahead:
                    update(c_after, c_after + w)            # my block k+1 first ...
                    factor_and_pack(next_k, nxt)            # ... factor it ...
                    work = start_bcast(next_k, nxt)         # ... and ship it while update k continues
                    update(c_after + w, nloc)
                else:
                    update(c_after, nloc)
                    factor_and_pack(next_k, nxt)
                    work = start_bcast(next_k, nxt)
            else:
                if lookahead:
                    work = start_bcast(next_k, nxt)         # post the receive, then compute
                    update(c_after, nloc)
                else:
                    update(c_after, nloc)
                    work = start_bcast(next_k, nxt)
        else:
            update(c_after, nloc)
        # interchanges left of the panel (dgetrf.f:193)
        if c_before > 0:
            ops.laswp(aloc[j:, :c_before], 1, jb, piv)

    # assemble the global IPIV / INFO (host side, tiny)
    import numpy as np
    ipiv = np.zeros(n, dtype=np.int32)
    for (j, piv, inf) in ipiv_all:
        p = piv.cpu().numpy()
        ipiv[j:j + len(p)] = p + j
        v = int(round(float(inf.cpu().numpy()[0])))
        if info == 0 and v > 0:
            info = v + j
    return ipiv, info


def ppotrf(ops, dist, desc: BlockCyclic1D, aloc, lookahead: bool = True):
    """Distributed Cholesky A = L L^T (UPLO = 'L') of the block-column-cyclic matrix `aloc` (n x local_cols, in place;
    only the lower triangle is referenced, like SRC/dpotrf.f:65-71).  Right-looking (SRC/VARIANTS/cholesky/RL/dpotrf.f:
    205-229): per step the owner's factored block column L(j:n, j:j+jb) is broadcast and every rank updates its local
    trailing block columns c:  A(c:n, c) -= L(c:n, k) L(c:c+w, k)^T  (DSYRK on the diagonal block, DGEMM below).
    Returns INFO (0, or the order of the first leading minor that is not positive)."""
    n, nb, me = desc.n, desc.nb, desc.rank
    nblk = desc.nblocks
    bufs = [ops.empty_vec(n * nb + 8) for _ in range(2)]
    infos = []

    def panel_view(buf, rows, jb):
        return buf[: n * nb].view(nb, n).t()[:rows, :jb]

    def factor_and_pack(k, buf):
        j, jb = k * nb, desc.width(k)
        lo = desc.local_offset(k)
        panel = aloc[j:, lo:lo + jb]
        inf = ops.potrf_panel(panel)
        ops.copy(panel_view(buf, n - j, jb), panel)
        tail = buf[n * nb: n * nb + 8]
        tail.zero_()
        ops.copy(tail[:1], ops.to_float64(inf))

    def start_bcast(k, buf):
        if dist is None or desc.world == 1:
            return None
        return dist.broadcast(buf, src=desc.owner(k), async_op=True)

    def update_block(c, pv, j):
        """apply panel (rows j.., in pv) to local block column c (global index)"""
        gc, w = c * nb, desc.width(c)
        lo = desc.local_offset(c)
        lc = pv[gc - j: gc - j + w, :]                     # L(gc:gc+w, k)
        ops.syrk_update(lc, aloc[gc:gc + w, lo:lo + w])
        if gc + w < n:
            ops.gemm_nt_update(pv[gc - j + w:, :], lc, aloc[gc + w:, lo:lo + w])

    if desc.owner(0) == me:
        factor_and_pack(0, bufs[0])
    work = start_bcast(0, bufs[0])
    for k in range(nblk):
        cur, nxt = bufs[k % 2], bufs[(k + 1) % 2]
        j, jb = k * nb, desc.width(k)
        if work is not None:
            work.wait()
        infos.append((j, cur[n * nb: n * nb + 1] * 1.0))        # a copy: the buffer is reused two steps later
        pv = panel_view(cur, n - j, jb)
        mine = [c for c in desc.local_blocks() if c > k]
        work = None
        nk = k + 1
        if nk < nblk:
            if desc.owner(nk) == me:
                if lookahead:
                    update_block(nk, pv, j)
                    factor_and_pack(nk, nxt)
                    work = start_bcast(nk, nxt)
                    for c in mine:
                        if c != nk:
                            update_block(c, pv, j)
                else:
                    for c in mine:
                        update_block(c, pv, j)
                    factor_and_pack(nk, nxt)
                    work = start_bcast(nk, nxt)
            else:
                if lookahead:
                    work = start_bcast(nk, nxt)
                for c in mine:
                    update_block(c, pv, j)
                if not lookahead:
                    work = start_bcast(nk, nxt)
    info = 0
    for (j, inf) in infos:
        v = int(round(float(inf.cpu().numpy()[0])))
        if info == 0 and v > 0:
            info = v + j
    return info


def pgeqrf(ops, dist, desc: BlockCyclic1D, aloc, lookahead: bool = True):
    """Distributed Householder QR (SRC/dgeqrf.f:244-267) of the block-column-cyclic matrix `aloc` (n x local_cols, in
    place: R on/above the diagonal, the reflectors V below).  Per step the owner factors its block column (DGEQRF
    panel + DLARFT), broadcasts V, T and tau in one message, and every rank applies H^T = I - V T^T V^T to its local
    trailing columns (DLARFB 'L','T','F','C', dgeqrf.f:262).  Returns tau (length n, replicated)."""
    n, nb, me = desc.n, desc.nb, desc.rank
    nblk = desc.nblocks
    bufs = [ops.empty_vec(n * nb + nb * nb + nb) for _ in range(2)]
    taus = []

    def panel_view(buf, rows, jb):
        return buf[: n * nb].view(nb, n).t()[:rows, :jb]

    def t_view(buf, jb):
        return buf[n * nb: n * nb + nb * nb].view(nb, nb).t()[:jb, :jb]

    def factor_and_pack(k, buf):
        j, jb = k * nb, desc.width(k)
        lo = desc.local_offset(k)
        panel = aloc[j:, lo:lo + jb]
        tau, t = ops.qr_panel(panel)
        ops.copy(panel_view(buf, n - j, jb), panel)
        kk = tau.shape[0]
        tv = t_view(buf, jb)
        tv.zero_()
        ops.copy(tv[:kk, :kk], t)
        tt = buf[n * nb + nb * nb: n * nb + nb * nb + nb]
        tt.zero_()
        ops.copy(tt[:kk], tau)

    def start_bcast(k, buf):
        if dist is None or desc.world == 1:
            return None
        return dist.broadcast(buf, src=desc.owner(k), async_op=True)

    if desc.owner(0) == me:
        factor_and_pack(0, bufs[0])
    work = start_bcast(0, bufs[0])
    for k in range(nblk):
        cur, nxt = bufs[k % 2], bufs[(k + 1) % 2]
        j, jb = k * nb, desc.width(k)
        kk = min(jb, n - j)
        if work is not None:
            work.wait()
        taus.append((j, kk, cur[n * nb + nb * nb: n * nb + nb * nb + nb] * 1.0))   # a copy (buffer reuse)
        v = panel_view(cur, n - j, jb)[:, :kk]
        t = t_view(cur, jb)[:kk, :kk]
        c_after = desc.first_local_col_after(k)
        nloc = aloc.shape[1]
        work = None
        nk = k + 1

        def apply(c0, c1):
            if c1 > c0:
                ops.larfb_lt(v, t, aloc[j:, c0:c1])

        if nk < nblk:
            if desc.owner(nk) == me:
                w = desc.width(nk)
                if lookahead:
                    apply(c_after, c_after + w)
                    factor_and_pack(nk, nxt)
                    work = start_bcast(nk, nxt)
                    apply(c_after + w, nloc)
                else:
                    apply(c_after, nloc)
                    factor_and_pack(nk, nxt)
                    work = start_bcast(nk, nxt)
            else:
                if lookahead:
                    work = start_bcast(nk, nxt)
                apply(c_after, nloc)
                if not lookahead:
                    work = start_bcast(nk, nxt)
        else:
            apply(c_after, nloc)
    import numpy as np
    tau = np.zeros(n)
    for (j, kk, tt) in taus:
        tau[j:j + kk] = tt.cpu().numpy()[:kk]
    return tau


def fill_local_random(ops_dev, desc: BlockCyclic1D, iseed=(1988, 1989, 1990, 1991), device="cuda"):
    """Local part of the global DLARNV(2) matrix (column-by-column stream): block k gets stream offset k*nb*n."""
    from . import dev
    import ctypes as C
    from . import lib
    n, nb = desc.n, desc.nb
    aloc = dev.colmajor(n, desc.local_cols(), device=device)
    seed = (C.c_int * 4)(*iseed)
    for k in desc.local_blocks():
        lo = desc.local_offset(k)
        w = desc.width(k)
        view = aloc[:, lo:lo + w]
        rc = lib().lb200_dlarnv_matrix(dev.stream(), C.byref(seed), k * nb * n, n, w, view.data_ptr(), dev.ld(view))
        assert rc == 0
    return aloc
