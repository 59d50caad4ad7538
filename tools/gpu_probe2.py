"""GPU probe #2: device-resident speed of DGETRF / DPOTRF / DGEQRF at large n, with LAPACK-style residual
ratios computed on the GPU from an independent product (torch / cuBLAS used in the checker only).
usage: python tools/gpu_probe2.py [sizes...]   (default 4096 8192 16384 32768)"""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import lapack_b200 as lb  # noqa: E402

EPS = 2.0 ** -53
SEED = (1988, 1989, 1990, 1991)
sizes = [int(x) for x in sys.argv[1:] if x.isdigit()] or [4096, 8192, 16384, 32768]
which = [x for x in sys.argv[1:] if not x.isdigit()] or ["getrf", "potrf", "geqrf"]
out = {}


def timed(fn):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e-3


def norm1(x):
    return x.abs().sum(dim=0).max().item()


def lu_residual(a0, lu, ipiv):
    n = a0.shape[0]
    L = torch.tril(lu, -1)
    L.diagonal().fill_(1.0)
    U = torch.triu(lu)
    prod = L @ U
    del L, U
    # apply the interchanges to A: row i <-> ipiv[i]
    perm = torch.arange(n, device=a0.device)
    piv = (ipiv.long() - 1).cpu().tolist()
    pl = perm.cpu().tolist()
    for i, p in enumerate(piv):
        if p != i:
            pl[i], pl[p] = pl[p], pl[i]
    perm = torch.tensor(pl, device=a0.device)
    prod -= a0[perm]
    return norm1(prod) / (n * norm1(a0) * EPS)


for n in sizes:
    if "getrf" in which:
        a = lb.dev.larnv_matrix(n, n, SEED)
        a0 = a.clone() if n <= 32768 else None
        ipiv, info = lb.dev.getrf(a)          # warm-up run also used for the check
        torch.cuda.synchronize()
        res = lu_residual(a0, a, ipiv) if a0 is not None else float("nan")
        # cross-check IPIV against cuSOLVER (non-authoritative)
        match = None
        if n <= 32768:
            LU2, piv2 = torch.linalg.lu_factor(a0)
            match = bool(torch.equal(piv2.int(), ipiv.int()))
            del LU2, piv2
        ts = []
        for _ in range(2):
            a.copy_(a0)
            ts.append(timed(lambda: lb.dev.getrf(a)))
        t = min(ts)
        fl = 2.0 * n ** 3 / 3 - n ** 2 / 2 + 5.0 * n / 6
        print(f"DGETRF n={n}: {t*1e3:.1f} ms  {fl/t*1e-12:.2f} TF/s  resid={res:.3f} info={int(info.item())} ipiv==cusolver:{match}", flush=True)
        out[f"getrf_{n}"] = {"ms": t * 1e3, "tflops": fl / t * 1e-12, "resid": res, "ipiv_match_cusolver": match}
        del a, a0
        torch.cuda.empty_cache()
    if "potrf" in which:
        a = lb.dev.larnv_matrix(n, n, SEED)
        lb.dev.make_spd(a, float(n))
        a0 = a.clone()
        info = lb.dev.potrf("L", a)
        torch.cuda.synchronize()
        L = torch.tril(a)
        prod = L @ L.t()
        del L
        prod -= a0
        prod = torch.tril(prod)
        r1 = (prod.abs().sum(dim=0) + prod.abs().sum(dim=1) - prod.diagonal().abs()).max().item()
        an = torch.tril(a0)
        a1 = (an.abs().sum(dim=0) + an.abs().sum(dim=1) - an.diagonal().abs()).max().item()
        res = r1 / (n * a1 * EPS)
        del prod, an
        ts = []
        for _ in range(2):
            a.copy_(a0)
            ts.append(timed(lambda: lb.dev.potrf("L", a)))
        t = min(ts)
        fl = n ** 3 / 3 + n ** 2 / 2 + n / 6
        print(f"DPOTRF n={n}: {t*1e3:.1f} ms  {fl/t*1e-12:.2f} TF/s  resid={res:.3f} info={int(info.item())}", flush=True)
        out[f"potrf_{n}"] = {"ms": t * 1e3, "tflops": fl / t * 1e-12, "resid": res}
        del a, a0
        torch.cuda.empty_cache()
    if "geqrf" in which:
        a = lb.dev.larnv_matrix(n, n, SEED)
        a0 = a.clone()
        tau = lb.dev.geqrf(a)
        torch.cuda.synchronize()
        # check R^T R = A^T A (Q-less normal-equation identity), scaled like dqrt01
        R = torch.triu(a)
        g1 = R.t() @ R
        g1 -= a0.t() @ a0
        res = norm1(g1) / (n * norm1(a0) ** 2 * EPS)
        del R, g1
        ts = []
        for _ in range(2):
            a.copy_(a0)
            ts.append(timed(lambda: lb.dev.geqrf(a)))
        t = min(ts)
        fl = 4.0 * n ** 3 / 3 + 2.0 * n ** 2 + 14.0 * n / 3
        print(f"DGEQRF n={n}: {t*1e3:.1f} ms  {fl/t*1e-12:.2f} TF/s  gram-resid={res:.3f}", flush=True)
        out[f"geqrf_{n}"] = {"ms": t * 1e3, "tflops": fl / t * 1e-12, "gram_resid": res}
        del a, a0
        torch.cuda.empty_cache()

os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "probe2.json"), "w"), indent=1)
