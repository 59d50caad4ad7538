"""Multi-GPU DGETRF on a P x Q process grid: ONE matrix, 2D block-cyclic with square NB x NB blocks (BASELINE config C5a,
SURVEY 8e: 8 = 2 x 4, 4 = 2 x 2, 2 = 1 x 2 or 2 x 1).  Distributed form of SRC/dgetrf.f:180-212.

Rank r sits at (p, q) = (r // Q, r % Q) and holds the blocks (bi, bj) with bi % P == p, bj % Q == q as one column-major
local matrix (mloc x nloc).  Per step k (panel = global block column k, diagonal block on process row pk = k % P, process
column qk = k % Q):

  panel     the P ranks of process column qk send their pieces of the panel to the diagonal owner (pk, qk); the owner
            factors the full-height panel with the single-GPU recursive kernel (DGETRF2 -- the pivot search spans the
            whole process column, IPIV is bit-identical to the single-GPU factorization), returns the factored pieces to
            the column ranks, and every (p, qk) broadcasts its piece + the pivots + INFO along its process ROW (so each
            rank receives only the (n-j)/P rows of L it multiplies with).  Runs on its own high-priority stream and its own
            communicators, one step ahead of the update (look-ahead).
  swaps     dgetrf.f:193,199 across process rows: the panel's NB interchanges are composed on the device into two maps
            (lb200_laswp_compose); every rank packs the rows it owns that move INTO block row k and sends them to process
            row pk, process row pk sends the NB original rows of block row k to the others, everybody scatters.  Fixed-size
            messages (NB x local columns), no host synchronisation, pairwise inside each process column.
  U12       process row pk solves with L11 (dgetrf.f:204) and broadcasts U12 along the process COLUMN.
  update    local DGEMM A22 -= L21 U12 (dgetrf.f:212).

With P = 1 the column-group traffic disappears and the scheme is the 1 x Q block-column layout of lapack_b200/dist.py
(but with the panel on its own stream and messages trimmed to the live rows).  torch / torch.distributed are plumbing
(device memory, streams, NCCL groups); all arithmetic goes through the C ABI.  The local backend is an object so that
tests/test_dist2d_cpu.py can run the identical schedule over gloo with the oracle as backend.
"""
from __future__ import annotations

import contextlib
from dataclasses import dataclass


@dataclass
class BlockCyclic2D:
    n: int
    nb: int
    P: int
    Q: int
    rank: int

    @property
    def p(self) -> int:
        return self.rank // self.Q

    @property
    def q(self) -> int:
        return self.rank % self.Q

    @property
    def nblocks(self) -> int:
        return (self.n + self.nb - 1) // self.nb

    def rank_of(self, p: int, q: int) -> int:
        return p * self.Q + q

    def width(self, kb: int) -> int:
        return min(self.nb, self.n - kb * self.nb)

    @staticmethod
    def _count_before(kb: int, me: int, period: int) -> int:
        """number of indices in {me, me+period, ...} that are < kb"""
        return 0 if kb <= me else (kb - me + period - 1) // period

    def _extent(self, me: int, period: int) -> int:
        return sum(self.width(b) for b in range(me, self.nblocks, period))

    def mloc_of(self, p: int) -> int:
        return self._extent(p, self.P)

    def nloc_of(self, q: int) -> int:
        return self._extent(q, self.Q)

    @property
    def mloc(self) -> int:
        return self.mloc_of(self.p)

    @property
    def nloc(self) -> int:
        return self.nloc_of(self.q)

    def lrow0_of(self, p: int, kb: int) -> int:
        """first local row (on process row p) whose global block row is >= kb"""
        return min(self._count_before(kb, p, self.P) * self.nb, self.mloc_of(p))

    def lrow0(self, kb: int) -> int:
        return self.lrow0_of(self.p, kb)

    def lcol0(self, kb: int) -> int:
        return min(self._count_before(kb, self.q, self.Q) * self.nb, self.nloc)

    def global_rows(self, p=None):
        import numpy as np
        p = self.p if p is None else p
        out = [np.arange(b * self.nb, b * self.nb + self.width(b)) for b in range(p, self.nblocks, self.P)]
        return np.concatenate(out) if out else np.zeros(0, dtype=np.int64)

    def global_cols(self, q=None):
        import numpy as np
        q = self.q if q is None else q
        out = [np.arange(b * self.nb, b * self.nb + self.width(b)) for b in range(q, self.nblocks, self.Q)]
        return np.concatenate(out) if out else np.zeros(0, dtype=np.int64)


def default_grid(world: int):
    """8 = 2 x 4, 4 = 2 x 2, 2 = 1 x 2; otherwise the most square P <= Q"""
    p = 1
    for c in range(1, int(world ** 0.5) + 1):
        if world % c == 0:
            p = c
    return p, world // p


class Groups:
    """Communicators of one rank: its process row (panel broadcast), and TWO for its process column -- panel traffic and
    update traffic run concurrently on different streams and must not queue behind each other."""

    def __init__(self, dist, desc: BlockCyclic2D):
        self.dist = dist
        self.row = self.col_pan = self.col_upd = None
        if dist is None or desc.P * desc.Q == 1:
            return
        # every rank creates every group, in the same order (new_group is collective)
        for pp in range(desc.P):
            g = dist.new_group([desc.rank_of(pp, qq) for qq in range(desc.Q)]) if desc.Q > 1 else None
            if pp == desc.p:
                self.row = g
        for kind in ("col_pan", "col_upd"):
            for qq in range(desc.Q):
                g = dist.new_group([desc.rank_of(pp, qq) for pp in range(desc.P)]) if desc.P > 1 else None
                if qq == desc.q:
                    setattr(self, kind, g)


class GpuOps2D:
    """Local compute through the lapack_b200 C ABI on torch CUDA tensors (column-major views) + the two streams."""

    def __init__(self, device, panel_stream: bool = True, panel_priority: int = -1, concurrent_panel_max_rows: int = 16384):
        import torch
        from . import dev
        self.torch = torch
        self.dev = dev
        self.device = device
        self.main = torch.cuda.current_stream(device)
        # panel_stream=False keeps the panel on the update stream (serial schedule, for A/B measurements)
        self.panel = torch.cuda.Stream(device=device, priority=panel_priority) if panel_stream else None
        # Panels taller than this run on the UPDATE stream instead (between the update of their own columns and the rest of
        # the update).  Measured (profiles/r02_dist_panel_stream_ab.txt): a tall panel's leaf kernels hold one whole SM per
        # 1024 rows and are ~100 separate high-priority launches, each of which drains and refills the update GEMM's CTAs --
        # overlapping them costs the GEMM 3x the panel's own duration (29.5 vs 33.4 TFLOP/s in situ); panels that fit one
        # thread-block cluster (<= 16 SMs) overlap profitably.
        self.concurrent_panel_max_rows = concurrent_panel_max_rows

    def empty_vec(self, n):
        return self.torch.empty(n, dtype=self.torch.float64, device=self.device)

    def panel_factor(self, panel):
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

    def compose(self, piv):
        return self.dev.laswp_compose(piv)

    def gather_rows(self, a, idx, w):
        if a.shape[1] > 0 and idx.shape[0] > 0:
            self.dev.gather_rows(a, idx, w)

    def scatter_rows(self, w, idx, a):
        if a.shape[1] > 0 and idx.shape[0] > 0:
            self.dev.scatter_rows(w, idx, a)

    def panel_stream(self, rows=0):
        if self.panel is None or rows > self.concurrent_panel_max_rows:
            return contextlib.nullcontext()
        return self.torch.cuda.stream(self.panel)

    def fork_panel(self):
        """panel stream waits for everything queued on the main stream so far"""
        if self.panel is not None:
            self.panel.wait_stream(self.main)

    def join_panel(self):
        """main stream waits for everything queued on the panel stream so far"""
        if self.panel is not None:
            self.main.wait_stream(self.panel)


def _colmajor_view(buf, rows, cols, off=0):
    return buf[off: off + rows * cols].view(cols, rows).t()


def pgetrf2d(ops, dist, desc: BlockCyclic2D, aloc, groups: Groups = None, lookahead: bool = True, trace: list = None):
    """In-place LU with partial pivoting of the 2D block-cyclic matrix `aloc` (mloc x nloc column-major view).
    Returns (ipiv, info): global 1-based pivot rows (length n, replicated), INFO as DGETRF defines it."""
    torch = ops.torch
    n, nb, P, Q = desc.n, desc.nb, desc.P, desc.Q
    p, q = desc.p, desc.q
    nblk = desc.nblocks
    mloc, nloc = desc.mloc, desc.nloc
    groups = groups or Groups(dist, desc)
    dev = aloc.device
    i32 = torch.int32
    tail_len = nb + 8
    # ---- persistent buffers
    pbuf = [ops.empty_vec(mloc * nb + tail_len) for _ in range(2)]      # my process row's piece of panel k + pivots + INFO
    root_state = {}
    nbuf = nb * nloc
    sbuf = ops.empty_vec(nbuf) if P > 1 else None                       # rows I own that move into block row k
    obuf = ops.empty_vec(nbuf) if P > 1 else None                       # original rows of block row k
    ubuf = ops.empty_vec(nbuf) if P > 1 else None                       # U12 received from process row pk
    srecv = {}
    ipiv_all = []
    arange_nb = torch.arange(nb, device=dev, dtype=i32)

    def mark():
        """CUDA event on the current stream (tracing only)"""
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        return e

    def p2p(ops_list, group):
        """issue a batch of sends / receives [(is_send, tensor, peer_rank)], return the work handles"""
        if not ops_list:
            return []
        reqs = [dist.P2POp(dist.isend if s else dist.irecv, t, peer, group) for (s, t, peer) in ops_list]
        return dist.batch_isend_irecv(reqs)

    def wait_all(works):
        for w in works:
            w.wait()

    def rel_rows_of(pp, k):
        """relative (to row k*nb) global rows of process row pp's local rows lrow0_of(pp,k) .. mloc_of(pp), int32 device"""
        r0, r1 = desc.lrow0_of(pp, k), desc.mloc_of(pp)
        lr = torch.arange(r0, r1, device=dev, dtype=i32)
        return ((lr // nb) * P + pp) * nb + lr % nb - k * nb

    def owner_and_local(grow):
        bg = grow // nb
        return bg % P, (bg // P) * nb + grow % nb

    # ------------------------------------------------------------------------------------------ panel
    def panel_phase(k):
        """gather -> factor -> return pieces -> broadcast along process rows.  Returns the broadcast work handle(s)."""
        buf = pbuf[k % 2]
        j, jb = k * nb, desc.width(k)
        pk, qk = k % P, k % Q
        r0 = desc.lrow0(k)
        rows = mloc - r0
        piece = _colmajor_view(buf, rows, jb)
        tail = buf[rows * jb: rows * jb + tail_len]
        msg = buf[: rows * jb + tail_len]
        if q == qk:
            lc = desc.lcol0(k)
            src = aloc[r0:, lc:lc + jb]
            if P == 1:
                piv, inf = ops.panel_factor(src)
                ops.copy(piece, src)
                tail.zero_()
                ops.copy(tail[:jb], piv.to(torch.float64))
                ops.copy(tail[nb:nb + 1], inf.to(torch.float64))
            else:
                ops.copy(piece, src)
                root = desc.rank_of(pk, qk)
                if p != pk:
                    if rows > 0:
                        wait_all(p2p([(True, buf[: rows * jb], root)], groups.col_pan))
                    wait_all(p2p([(False, msg, root)], groups.col_pan))
                else:
                    if "pan" not in root_state:
                        root_state["pan"] = ops.empty_vec(n * nb)
                        root_state["rb"] = {pp: ops.empty_vec(desc.mloc_of(pp) * nb + tail_len) for pp in range(P) if pp != p}
                    pan = _colmajor_view(root_state["pan"], n - j, jb)
                    rb = root_state["rb"]
                    rows_of = {pp: desc.mloc_of(pp) - desc.lrow0_of(pp, k) for pp in range(P)}
                    wait_all(p2p([(False, rb[pp][: rows_of[pp] * jb], desc.rank_of(pp, qk)) for pp in rb if rows_of[pp] > 0],
                                 groups.col_pan))
                    idx = {pp: rel_rows_of(pp, k) for pp in range(P)}
                    for pp in range(P):
                        if rows_of[pp] > 0:
                            w = piece if pp == p else _colmajor_view(rb[pp], rows_of[pp], jb)
                            ops.scatter_rows(w, idx[pp], pan)
                    piv, inf = ops.panel_factor(pan)
                    tail.zero_()
                    ops.copy(tail[:jb], piv.to(torch.float64))
                    ops.copy(tail[nb:nb + 1], inf.to(torch.float64))
                    sends = []
                    for pp in range(P):
                        if pp == p:
                            ops.gather_rows(pan, idx[pp], piece)
                        else:
                            if rows_of[pp] > 0:
                                ops.gather_rows(pan, idx[pp], _colmajor_view(rb[pp], rows_of[pp], jb))
                            ops.copy(rb[pp][rows_of[pp] * jb: rows_of[pp] * jb + tail_len], tail)
                            sends.append((True, rb[pp][: rows_of[pp] * jb + tail_len], desc.rank_of(pp, qk)))
                    wait_all(p2p(sends, groups.col_pan))
                ops.copy(src, piece)                                   # the factored piece back into the matrix
        if Q > 1:
            return [dist.broadcast(msg, src=desc.rank_of(p, qk), group=groups.row, async_op=True)]
        return []

    # ------------------------------------------------------------------------------------------ update
    def swap_rows(k, piv, plan, c0, c1):
        if c1 <= c0:
            return
        j, jb = k * nb, desc.width(k)
        pk = k % P
        if P == 1:
            ops.laswp(aloc[j:, c0:c1], 1, jb, piv)                     # dgetrf.f:193 / :199, all rows are local
            return
        ncols = c1 - c0
        cols = aloc[:, c0:c1]
        src_top, inv_top = plan
        own_src, loc_src = owner_and_local(j + src_top)
        s_mine = _colmajor_view(sbuf, jb, ncols)
        o_rows = _colmajor_view(obuf, jb, ncols)
        ops.gather_rows(cols, torch.where(own_src == p, loc_src, -1).to(i32), s_mine)
        ltop = desc.lrow0_of(pk, k)
        if p == pk:
            ops.copy(o_rows, aloc[ltop:ltop + jb, c0:c1])
            msgs = []
            for pp in range(P):
                if pp != p:
                    if pp not in srecv:
                        srecv[pp] = ops.empty_vec(nbuf)
                    msgs.append((False, srecv[pp][: jb * ncols], desc.rank_of(pp, q)))
                    msgs.append((True, obuf[: jb * ncols], desc.rank_of(pp, q)))
            wait_all(p2p(msgs, groups.col_upd))
            t_idx = arange_nb[:jb] + ltop
            for pp in range(P):
                sx = s_mine if pp == p else _colmajor_view(srecv[pp], jb, ncols)
                ops.scatter_rows(sx, torch.where(own_src == pp, t_idx, -1).to(i32), cols)
        else:
            peer = desc.rank_of(pk, q)
            wait_all(p2p([(True, sbuf[: jb * ncols], peer), (False, obuf[: jb * ncols], peer)], groups.col_upd))
        own_dst, loc_dst = owner_and_local(j + inv_top)
        ops.scatter_rows(o_rows, torch.where((inv_top >= jb) & (own_dst == p), loc_dst, -1).to(i32), cols)

    def update_cols(k, piv, plan, lp, c0, c1):
        """interchanges, U12 and the trailing update on the local columns [c0, c1)"""
        if c1 <= c0:
            return
        j, jb = k * nb, desc.width(k)
        pk = k % P
        e_a = mark() if trace is not None else None
        swap_rows(k, piv, plan, c0, c1)
        e_b = mark() if trace is not None else None
        ncols = c1 - c0
        r0 = desc.lrow0(k)
        lr1 = desc.lrow0(k + 1)
        if p == pk:
            top = aloc[r0:r0 + jb, c0:c1]
            ops.trsm_llnu(lp[:jb, :], top)                             # dgetrf.f:204
            u = top
            if P > 1:
                uc = _colmajor_view(ubuf, jb, ncols)
                ops.copy(uc, top)
                dist.broadcast(ubuf[: jb * ncols], src=desc.rank_of(pk, q), group=groups.col_upd)
        else:
            dist.broadcast(ubuf[: jb * ncols], src=desc.rank_of(pk, q), group=groups.col_upd)
            u = _colmajor_view(ubuf, jb, ncols)
        e_c = mark() if trace is not None else None
        if lr1 < mloc:
            ops.gemm_update(lp[lr1 - r0:, :], u, aloc[lr1:, c0:c1])    # dgetrf.f:212
        if trace is not None:
            trace.append(("upd", k, e_a, e_b, e_c, mark(), 2.0 * max(0, mloc - lr1) * jb * ncols))

    # ------------------------------------------------------------------------------------------ driver
    def traced_panel(k):
        if trace is None:
            return panel_phase(k)
        e0 = mark()
        w = panel_phase(k)
        trace.append(("panel", k, e0, mark()))
        return w

    with ops.panel_stream(n):
        ops.fork_panel()
        works = traced_panel(0)
    for k in range(nblk):
        j, jb = k * nb, desc.width(k)
        buf = pbuf[k % 2]
        r0 = desc.lrow0(k)
        rows = mloc - r0
        t_begin = mark() if trace is not None else None
        wait_all(works)
        ops.join_panel()
        t_ready = mark() if trace is not None else None
        tail = buf[rows * jb: rows * jb + tail_len]
        piv = tail[:jb].to(i32)                                        # relative to row j, 1-based
        ipiv_all.append((j, piv, tail[nb:nb + 1] * 1.0))
        lp = _colmajor_view(buf, rows, jb)                             # L rows of my process row, local rows r0..
        plan = ops.compose(piv) if P > 1 else None
        c_after = desc.lcol0(k + 1)
        c_before = desc.lcol0(k)
        works = []
        nk = k + 1
        if nk < nblk:
            if lookahead:
                cn0, cn1 = (desc.lcol0(nk), desc.lcol0(nk) + desc.width(nk)) if q == nk % Q else (c_after, c_after)
                update_cols(k, piv, plan, lp, cn0, cn1)                # the next panel's columns first ...
                with ops.panel_stream(n - nk * nb):
                    ops.fork_panel()
                    works = traced_panel(nk)                           # ... factor + ship it while update k continues
                update_cols(k, piv, plan, lp, cn1, nloc)
            else:
                update_cols(k, piv, plan, lp, c_after, nloc)
                with ops.panel_stream(n - nk * nb):
                    ops.fork_panel()
                    works = traced_panel(nk)
        else:
            update_cols(k, piv, plan, lp, c_after, nloc)
        swap_rows(k, piv, plan, 0, c_before)                           # interchanges left of the panel (dgetrf.f:193)
        if trace is not None:
            trace.append(("step", k, t_begin, t_ready, mark()))

    import numpy as np
    ipiv = np.zeros(n, dtype=np.int32)
    info = 0
    for (j, piv, inf) in ipiv_all:
        pv = piv.cpu().numpy()
        ipiv[j:j + len(pv)] = pv + j
        v = int(round(float(inf.cpu().numpy()[0])))
        if info == 0 and v > 0:
            info = v + j
    return ipiv, info


def fill_local_random_2d(desc: BlockCyclic2D, iseed=(1988, 1989, 1990, 1991), device="cuda"):
    """Local part of the global n x n DLARNV(2) matrix (column-by-column stream): element (i, j) is draw j*n + i."""
    from . import dev
    n, nb = desc.n, desc.nb
    aloc = dev.colmajor(desc.mloc, desc.nloc, device=device)
    for li, bi in enumerate(range(desc.p, desc.nblocks, desc.P)):
        h = desc.width(bi)
        # all local block columns of this block row at once would need a strided stream; go block column by block column
        for lj, bj in enumerate(range(desc.q, desc.nblocks, desc.Q)):
            w = desc.width(bj)
            dev.larnv_submatrix(aloc[li * nb: li * nb + h, lj * nb: lj * nb + w], bj * nb * n + bi * nb, n, iseed)
    return aloc


def randomized_residual_2d(torch, dist, desc: BlockCyclic2D, a0loc, luloc, ipiv, chunk=2048):
    """|| P A x - L (U x) ||_inf / (n ||A||_1 ||x||_inf eps) with A, L, U 2D block-cyclic (checker: torch ops allowed)."""
    import numpy as np
    n = desc.n
    dev = a0loc.device
    g = torch.Generator(device="cpu").manual_seed(1234)
    x = (torch.rand(n, generator=g, dtype=torch.float64) * 2 - 1).to(dev)
    rows = torch.from_numpy(desc.global_rows()).to(dev)
    cols = torch.from_numpy(desc.global_cols()).to(dev)
    multi = dist is not None and desc.P * desc.Q > 1

    def allreduce(v, op=None):
        if multi:
            dist.all_reduce(v) if op is None else dist.all_reduce(v, op=op)
        return v

    y1 = torch.zeros(n, dtype=torch.float64, device=dev)
    w = torch.zeros(n, dtype=torch.float64, device=dev)
    colsum = torch.zeros(n, dtype=torch.float64, device=dev)
    zero = torch.zeros((), dtype=torch.float64, device=dev)
    for c0 in range(0, len(cols), chunk):
        c1 = min(c0 + chunk, len(cols))
        blk = a0loc[:, c0:c1]
        y1.index_add_(0, rows, blk @ x[cols[c0:c1]])
        colsum.index_add_(0, cols[c0:c1], blk.abs().sum(dim=0))
        ublk = torch.where(rows.unsqueeze(1) <= cols[c0:c1].unsqueeze(0), luloc[:, c0:c1], zero)
        w.index_add_(0, rows, ublk @ x[cols[c0:c1]])
    allreduce(y1)
    allreduce(w)
    allreduce(colsum)
    anorm = colsum.max().item()
    z = torch.zeros(n, dtype=torch.float64, device=dev)
    for c0 in range(0, len(cols), chunk):
        c1 = min(c0 + chunk, len(cols))
        lblk = torch.where(rows.unsqueeze(1) > cols[c0:c1].unsqueeze(0), luloc[:, c0:c1], zero)
        z.index_add_(0, rows, lblk @ w[cols[c0:c1]])
    allreduce(z)
    z += w                                                             # unit diagonal of L
    perm = np.arange(n)
    piv = np.asarray(ipiv) - 1
    for i in range(n):
        pi = piv[i]
        if pi != i:
            perm[i], perm[pi] = perm[pi], perm[i]
    py = y1[torch.from_numpy(perm).to(dev)]
    num = (py - z).abs().max().item()
    den = n * anorm * x.abs().max().item() * 2.0 ** -53
    return num / den
