"""Checker for the distributed LU (test/bench infrastructure, torch ops allowed): O(n^2) randomized residual
|| P A x - L (U x) ||_inf / (n * ||A||_1 * ||x||_inf * eps) with A, L, U distributed by block columns."""
from __future__ import annotations

import numpy as np


def randomized_residual(torch, dist, desc, a0loc, luloc, ipiv, chunk=2048):
    n = desc.n
    dev = a0loc.device
    g = torch.Generator(device="cpu").manual_seed(1234)
    x = (torch.rand(n, generator=g, dtype=torch.float64) * 2 - 1).to(dev)
    cols = torch.tensor([desc.global_col(c) for c in range(desc.local_cols())], device=dev, dtype=torch.long)
    rows = torch.arange(n, device=dev).unsqueeze(1)

    def allreduce(v):
        if dist is not None and desc.world > 1:
            dist.all_reduce(v)
        return v

    # y1 = A x ; anorm = max column sum
    y1 = torch.zeros(n, dtype=torch.float64, device=dev)
    anorm = torch.zeros(1, dtype=torch.float64, device=dev)
    w = torch.zeros(n, dtype=torch.float64, device=dev)
    for c0 in range(0, len(cols), chunk):
        c1 = min(c0 + chunk, len(cols))
        blk = a0loc[:, c0:c1]
        y1 += blk @ x[cols[c0:c1]]
        anorm = torch.maximum(anorm, blk.abs().sum(dim=0).max().reshape(1))
        ublk = torch.where(rows <= cols[c0:c1].unsqueeze(0), luloc[:, c0:c1], torch.zeros((), dtype=torch.float64, device=dev))
        w += ublk @ x[cols[c0:c1]]
    allreduce(y1)
    allreduce(w)
    if dist is not None and desc.world > 1:
        dist.all_reduce(anorm, op=dist.ReduceOp.MAX)
    z = torch.zeros(n, dtype=torch.float64, device=dev)
    for c0 in range(0, len(cols), chunk):
        c1 = min(c0 + chunk, len(cols))
        lblk = torch.where(rows > cols[c0:c1].unsqueeze(0), luloc[:, c0:c1], torch.zeros((), dtype=torch.float64, device=dev))
        z += lblk @ w[cols[c0:c1]]
        z.index_add_(0, cols[c0:c1], w[cols[c0:c1]])          # unit diagonal of L
    allreduce(z)
    # P y1: apply the interchanges in order
    perm = np.arange(n)
    piv = np.asarray(ipiv) - 1
    for i in range(n):
        p = piv[i]
        if p != i:
            perm[i], perm[p] = perm[p], perm[i]
    py = y1[torch.from_numpy(perm).to(dev)]
    num = (py - z).abs().max().item()
    den = n * anorm.item() * x.abs().max().item() * 2.0 ** -53
    return num / den
