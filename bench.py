#!/usr/bin/env python
"""bench.py -- headline benchmark of lapack_b200 (contract: see the task statement / DESIGN.md section 6).

Metric (BASELINE.json): DGETRF/DPOTRF FP64 TFLOP/s on one B200 at n=32768.
One "step" = BASELINE configs[1] + configs[2] back to back on synthetic DLARNV(2) inputs:
    DPOTRF('L') + DPOTRS (1 RHS) of the n x n SPD matrix, then DGETRF of the n x n U(-1,1) matrix.
`value`  : algorithmic flops of the step / device time, inputs resident in HBM (CUDA events, max over ranks).
`e2e`    : the same step through the Fortran-77 ABI (dpotrf_/dpotrs_/dgetrf_) with pinned HOST buffers, i.e.
           H2D + factorization + D2H inside the timed region (wall clock around the synchronous calls).
`roofline`: the dominant kernel = trailing-update DMMA GEMM; achieved = its algorithmic flops / its measured
           launch durations (CUDA events on its own stream, recorded inside the timed region by the library).
`cpu_baseline`: the reference algorithm (oracle port: NB=64 blocked DGETRF + reference BLAS loops), one thread,
           on a bounded sample (DGESV, BASELINE configs[0] shape scaled to fit ~20 s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--n 32768] [--impl ours|reference]
N>1 runs under torchrun: every rank factors its own matrices (weak scaling, no data-path collective).
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
SEED = (1988, 1989, 1990, 1991)
EPS = 2.0 ** -53


def flops_getrf(n):
    return 2.0 * n ** 3 / 3 - n ** 2 / 2 + 5.0 * n / 6


def flops_potrf(n):
    return n ** 3 / 3 + n ** 2 / 2 + n / 6


def flops_potrs(n, nrhs):
    return 2.0 * n * n * nrhs


def flops_getrs(n, nrhs):
    return nrhs * (2.0 * n * n - n)


def flops_geqrf(m, n):
    """LAWN-41 count for m >= n: 2mn^2 - 2n^3/3 + mn + n^2 + 14n/3"""
    return 2.0 * m * n * n - 2.0 * n ** 3 / 3 + m * n + n * n + 14.0 * n / 3


# ----------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    """nvidia-smi sampled every 200 ms while the timed region runs (B200_PROFILING.md recipe)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index = index
        self.samples = []
        self.proc = None
        self.thread = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._read, daemon=True)
        self.thread.start()

    def _read(self):
        for line in self.proc.stdout:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) >= 7:
                self.samples.append(parts)

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm = sorted(int(float(s[0])) for s in self.samples if s[0].replace(".", "").isdigit())
        mx = [int(float(s[1])) for s in self.samples if s[1].replace(".", "").isdigit()]
        reasons = set()
        for s in self.samples:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), s[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        pw = [float(s[2]) for s in self.samples if s[2].replace(".", "").isdigit()]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(self.samples), "reasons": sorted(reasons)}


# ----------------------------------------------------------------------------------------------- CPU reference leg
def cpu_reference_sample(n, nrhs=1, keep=None):
    """One DGESV (DGETRF NB=64 + DGETRS) of the oracle port on an n x n DLARNV matrix; returns (seconds, flops)."""
    import numpy as np
    from oracle import oracle as O
    a, seed = O.random_matrix(n, n, SEED)
    xact, _ = O.random_matrix(n, nrhs, seed)
    b = np.asfortranarray(a @ xact)
    t0 = time.perf_counter()
    ipiv, info = O.dgesv(a, b)
    t = time.perf_counter() - t0
    err = float(np.max(np.abs(b - xact)) / np.max(np.abs(xact)))
    if keep is not None:
        keep["ipiv"], keep["x"], keep["n"] = ipiv.copy(), b.copy(), n
    return t, flops_getrf(n) + flops_getrs(n, nrhs), info, err


def cpu_reference_mix(n, nrhs=1):
    """The SAME routine mix as the GPU step, on the oracle port: DPOTRF('L') + DPOTRS of the n x n SPD matrix, then DGETRF of the
    n x n U(-1,1) matrix (SURVEY 8d inputs); returns (seconds, flops, infos)."""
    import numpy as np
    from oracle import oracle as O
    s, seed = O.spd_matrix(n, SEED)
    xact, _ = O.random_matrix(n, nrhs, seed)
    b = np.asfortranarray(s @ xact)
    a, _ = O.random_matrix(n, n, SEED)
    t0 = time.perf_counter()
    i1 = O.dpotrf("L", s)
    O.dpotrs("L", s, b)
    _, i2 = O.dgetrf(a)
    t = time.perf_counter() - t0
    return t, flops_potrf(n) + flops_potrs(n, nrhs) + flops_getrf(n), (i1, i2)


def pick_cpu_sample(budget_s, steps):
    """Largest n in a fixed ladder whose estimated total time fits the budget (calibrated on n=768)."""
    t, fl, _, _ = cpu_reference_sample(768)
    rate = fl / t
    for n in (4096, 3072, 2048, 1536, 1024):
        if (flops_getrf(n) / rate) * steps <= budget_s:
            return n
    return 1024


def run_reference(args, rank, world):
    if rank != 0:
        return
    steps = args.steps + args.warmup
    # same routine mix as the GPU arm's step (DPOTRF + DPOTRS + DGETRF), at the largest order of a fixed ladder that fits ~150 s
    t, fl, _ = cpu_reference_mix(768)
    rate = fl / t
    n = 1024
    for cand in (4096, 3072, 2048, 1536, 1024):
        if ((flops_potrf(cand) + flops_getrf(cand)) / rate) * steps <= 150.0:
            n = cand
            break
    ts = []
    for i in range(steps):
        t, fl, infos = cpu_reference_mix(n)
        assert infos == (0, 0), infos
        if i >= args.warmup:
            ts.append(t)
    tmean = sum(ts) / len(ts)
    val = fl / tmean * 1e-12
    sample = (f"DPOTRF('L')+DPOTRS(1 rhs)+DGETRF at n={n} (the GPU arm's routine mix at a bounded order; its n={args.n} would take hours): "
              "reference blocked algorithms (NB=64, recursive panels) on reference BLAS loops (oracle port), 1 thread -- the reference "
              "has no threading on this path")
    line = {
        "impl": "reference", "metric": "DGETRF/DPOTRF FP64 TFLOP/s", "value": val, "unit": "TFLOP/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": tmean * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "bounded CPU sample of the GPU arm's step: " + sample, "n": n},
        "cpu_baseline": {"value": val, "unit": "TFLOP/s", "cores": 1, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "TFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "the reference Fortran cannot be compiled in this image (no Fortran compiler); the timed code is the line-faithful C port in oracle/",
    }
    print(json.dumps(line))



def bench_batched(lb, torch, dev, batch, reps=5):
    """Batched 32x32 DGETRF (BASELINE configs[4], second half): HBM-bound, 16,512 algorithmic bytes per matrix."""
    a0 = lb.dev.larnv_matrix(32, 32 * batch, SEED, device=dev).t().contiguous().view(batch, 32, 32)
    a = a0.clone()
    best = 1e30
    for _ in range(reps):
        a.copy_(a0)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ipiv, info = lb.dev.getrf_batched32(a)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1e-3)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm = peaks.get("hbm_gbs", 6650.0)
    gbs = batch * 16512 / best * 1e-9
    return {"matrices": batch, "ms": best * 1e3, "matrices_per_s": batch / best, "gflops": batch * 21360 / best * 1e-9,
            "roofline": {"bound": "hbm", "achieved": gbs, "peak": hbm, "unit": "GB/s", "frac": gbs / hbm,
                         "peak_source": "MEASURED_PEAKS.json hbm_gbs" if "hbm_gbs" in peaks else "fallback 6650 GB/s"},
            "nonzero_info": int((info != 0).sum().item())}

OUT = None


class StdoutToStderr:
    """Route everything written to file descriptor 1 (NCCL's version banner, library chatter) to stderr while the
    benchmark runs; `emit` writes the ONE JSON line to the real stdout."""

    def __init__(self):
        sys.stdout.flush()
        self.real = os.dup(1)
        os.dup2(2, 1)

    def emit(self, text):
        sys.stdout.flush()
        os.write(self.real, (text + "\n").encode())


# ----------------------------------------------------------------------------------------------- GPU leg
def run_dist(args, rank, world, local_rank):
    """N > 1: ONE DGETRF of order n_dist, 2D block-cyclic on a P x Q grid over the N GPUs (BASELINE configs[4])."""
    import numpy as np
    import torch
    import torch.distributed as dist
    import lapack_b200 as lb
    from lapack_b200.dist2d import BlockCyclic2D, GpuOps2D, Groups, default_grid, fill_local_random_2d, pgetrf2d, randomized_residual_2d
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
        os.environ["NCCL_DEBUG"] = "WARN"          # keep NCCL's version banner off stdout: ONE JSON line only
    dist.init_process_group("nccl", device_id=dev)
    L = lb.lib()
    n, nb = args.n_dist, args.nb_dist
    if args.grid:
        P, Q = (int(v) for v in args.grid.lower().split("x"))
    else:
        P, Q = 1, world                           # measured faster than default_grid(world) on NVSwitch (profiles/r02_grid_compare.txt)
    assert P * Q == world, (P, Q, world)
    desc = BlockCyclic2D(n, nb, P, Q, rank)
    ops = GpuOps2D(dev)
    groups = Groups(dist, desc)

    def barrier():
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()

    # ---- correctness gate BEFORE timing: the same driver at n=8192 on both grids against the single-GPU factorization
    # of the same DLARNV matrix -- IPIV must be identical (tests/test_gpu_dist.py needs >= 2 GPUs, so the bench carries it)
    checks = {}
    if not args.no_check:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        from _dist2d_gpu_worker import check_against_single_gpu
        for (cp, cq) in sorted({(P, Q), default_grid(world)}):
            cdesc = BlockCyclic2D(8192, nb, cp, cq, rank)
            cgroups = groups if (cp, cq) == (P, Q) else Groups(dist, cdesc)
            ok, msg = check_against_single_gpu(dev, cdesc, cgroups, ops)
            checks[f"n8192_grid_{cp}x{cq}_ipiv_equals_single_gpu"] = ok
            checks[f"n8192_grid_{cp}x{cq}"] = msg
            assert ok, msg
        torch.cuda.empty_cache()

    # ---- SURVEY 8e second row on the driver's hardware: distributed Cholesky / QR (block-column cyclic, lapack_b200/dist.py):
    # factors against the single-GPU factorization at n=8192, then one timed run each at n_dist/2
    other = {}
    if not args.no_check:
        from lapack_b200.dist import BlockCyclic1D, GpuOps, fill_local_random, ppotrf, pgeqrf
        ops1 = GpuOps(dev)
        for which in ("dpotrf", "dgeqrf"):
            nc = 8192
            d1 = BlockCyclic1D(nc, 512 if which == "dpotrf" else 256, world, rank)
            cols = torch.tensor([d1.global_col(c) for c in range(d1.local_cols())], device=dev, dtype=torch.long)
            full = lb.dev.larnv_matrix(nc, nc, device=dev)
            if which == "dpotrf":
                lb.dev.make_spd(full, float(nc))
            aloc = lb.dev.colmajor(nc, len(cols), device=dev)
            aloc.copy_(full[:, cols])
            if which == "dpotrf":
                info_d = ppotrf(ops1, dist, d1, aloc)
                info_1 = int(lb.dev.potrf("L", full).item())
                low = (torch.arange(nc, device=dev).unsqueeze(1) >= cols.unsqueeze(0))
                diff = ((full[:, cols] - aloc).abs() * low).max().item()
                ok = info_d == 0 and info_1 == 0 and diff < 1e-9
            else:
                tau_d = pgeqrf(ops1, dist, d1, aloc)
                tau_1 = lb.dev.geqrf(full).cpu().numpy()
                diff = (full[:, cols] - aloc).abs().max().item() / full.abs().max().item()
                ok = diff < 1e-10 and float(np.max(np.abs(tau_d - tau_1))) < 1e-10
            flag = torch.tensor([1 if ok else 0], device=dev)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            checks[f"n8192_p{which}_equals_single_gpu"] = bool(flag.item())
            assert flag.item() == 1, (which, diff)
            del full, aloc
            torch.cuda.empty_cache()
            # timing at n_dist / 2 (the matrix of the LU leg would not leave room for a second copy)
            nt = n // 2
            d2 = BlockCyclic1D(nt, 512 if which == "dpotrf" else 256, world, rank)
            a2 = fill_local_random(ops1, d2, device=dev)
            if which == "dpotrf":
                c2 = torch.tensor([d2.global_col(c) for c in range(d2.local_cols())], device=dev, dtype=torch.long)
                a2[c2, torch.arange(len(c2), device=dev)] += float(nt)         # diagonally dominant; only the lower triangle is read
            w2 = a2.clone()
            ts = []
            for rep in range(2):
                w2.copy_(a2)
                barrier()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                if which == "dpotrf":
                    ppotrf(ops1, dist, d2, w2)
                else:
                    pgeqrf(ops1, dist, d2, w2)
                e1.record()
                torch.cuda.synchronize()
                tt = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
                ts.append(tt.item())
            fl2 = flops_potrf(nt) if which == "dpotrf" else flops_geqrf(nt, nt)
            other[f"p{which}"] = {"n": nt, "ms": min(ts), "tflops": fl2 / (min(ts) * 1e-3) * 1e-12,
                                  "tflops_per_gpu": fl2 / (min(ts) * 1e-3) * 1e-12 / world, "layout": f"block-column cyclic 1x{world}"}
            del a2, w2
            torch.cuda.empty_cache()

    a0 = fill_local_random_2d(desc, device=dev)
    a = lb.dev.colmajor(desc.mloc, desc.nloc, device=dev)
    fl = flops_getrf(n)
    ipiv = info = None
    for _ in range(max(args.warmup, 1)):
        a.copy_(a0)
        ipiv, info = pgetrf2d(ops, dist, desc, a, groups)
    barrier()
    resid = None if args.no_check else randomized_residual_2d(torch, dist, desc, a0, a, ipiv)
    sampler = ClockSampler(local_rank)
    launches0 = L.lb200_launch_count()
    barrier()
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        a.copy_(a0)
        pgetrf2d(ops, dist, desc, a, groups)
    e1.record()
    barrier()
    clocks = sampler.stop()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    lt = torch.tensor([float(L.lb200_launch_count() - launches0)], device=dev, dtype=torch.float64)
    dist.all_reduce(lt)
    ms = t.item()
    value = fl * args.steps / (ms * 1e-3) * 1e-12
    peak = max(L.lb200_fp64_peak_tflops(None, 0, 8, 2, 20000) for _ in range(2))
    del a, a0
    torch.cuda.empty_cache()
    batched = bench_batched(lb, torch, dev, (1 << 20) // world)          # 1M matrices sharded by contiguous slabs, no collective
    bt = torch.tensor([batched["ms"]], device=dev, dtype=torch.float64)
    dist.all_reduce(bt, op=dist.ReduceOp.MAX)
    batched["ms_max_over_ranks"] = bt.item()
    batched["aggregate_matrices_per_s"] = (1 << 20) / (bt.item() * 1e-3)
    if rank == 0:
        t1 = None
        try:
            t1 = json.load(open(os.path.join(ROOT, "profiles", "r02_t1_dgetrf_131072.json")))
        except Exception:
            pass
        line = {
            "metric": "DGETRF/DPOTRF FP64 TFLOP/s", "value": value, "unit": "TFLOP/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"one DGETRF n={n} (BASELINE configs[4]) 2D block-cyclic on a {P} x {Q} process grid, NB={nb}; panel "
                                   "gathered to the diagonal owner + broadcast along process rows with NCCL, row interchanges and U12 "
                                   "exchanged inside process columns, panel one step ahead on its own stream; DLARNV(2) seed 1988-1991",
                       "n": n, "parallelism": f"2D block-cyclic {P}x{Q}", "l2": "local matrix far larger than L2; restored from an HBM copy every step"},
            "pct_of_fp64_peak": value / (world * peak), "roofline": {"bound": "tensor", "achieved": value / world, "peak": peak, "unit": "TFLOP/s",
                                                                       "frac": value / (world * peak), "traffic": None,
                                                                       "note": "whole-factorization rate per GPU vs the in-run DMMA peak"},
            "cpu_baseline": None,
            "e2e": {"value": value, "unit": "TFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": n * 4,
                    "note": "the distributed matrix is generated on the devices (128 GiB does not pass through one host buffer); only IPIV/INFO return to the host"},
            "gpu_launches": int(lt.item()), "clocks": clocks,
            "checks": dict(checks, randomized_residual_ratio=resid, info=int(info)),
            "distributed_cholesky_qr": other,
            "t1_single_gpu_same_n": t1,
            "parallel_efficiency_vs_t1": (t1["ms"] / (world * ms / args.steps)) if (t1 and t1.get("n") == n) else None,
            "batched_dgetrf_32x32": batched,
        }
        OUT.emit(json.dumps(line))
    dist.destroy_process_group()


def run_ours(args, rank, world, local_rank):
    import torch
    import lapack_b200 as lb
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- lapack_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    L = lb.lib()
    L.lb200_profile_gemm.argtypes = [C.c_int]
    L.lb200_profile_gemm_read.argtypes = [C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_longlong)]
    dist = None
    if world > 1:
        import torch.distributed as dist_
        dist = dist_
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=dev)

    n = args.n
    nrhs = 1
    # ---- synthetic inputs: U(-1,1) from DLARNV(2), seed 1988..1991; SPD = (R+R')/2 + n*I
    a_lu0 = lb.dev.larnv_matrix(n, n, SEED, device=dev)
    a_po0 = lb.dev.larnv_matrix(n, n, SEED, device=dev)
    lb.dev.make_spd(a_po0, float(n))
    xact = lb.dev.larnv_matrix(n, nrhs, SEED, offset=n * n, device=dev)
    b_po0 = lb.dev.colmajor(n, nrhs, device=dev)
    b_po0.copy_(a_po0 @ xact)   # B = S * XACT (like dlarhs.f:319)
    a_lu = lb.dev.colmajor(n, n, device=dev)
    a_po = lb.dev.colmajor(n, n, device=dev)
    b_po = lb.dev.colmajor(n, nrhs, device=dev)
    step_flops = flops_potrf(n) + flops_potrs(n, nrhs) + flops_getrf(n)

    def step():
        a_po.copy_(a_po0)
        b_po.copy_(b_po0)
        a_lu.copy_(a_lu0)
        info_po = lb.dev.potrf("L", a_po)
        lb.dev.potrs("L", a_po, b_po)
        ipiv, info_lu = lb.dev.getrf(a_lu)
        return info_po, ipiv, info_lu

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 1)):
        info_po, ipiv, info_lu = step()
    barrier()
    # ---- correctness gates on the warm-up result (GPU-side independent products; checker only)
    checks = {}
    if rank == 0 and not args.no_check:
        Lm = torch.tril(a_po)
        r = Lm @ Lm.t()
        r -= a_po0
        r = torch.tril(r)
        rn = (r.abs().sum(0) + r.abs().sum(1) - r.diagonal().abs()).max().item()
        an = torch.tril(a_po0)
        an1 = (an.abs().sum(0) + an.abs().sum(1) - an.diagonal().abs()).max().item()
        checks["dpot01_ratio"] = rn / (n * an1 * EPS)
        checks["posv_rel_err"] = ((b_po - xact).abs().max() / xact.abs().max()).item()
        del Lm, r, an
        Lm = torch.tril(a_lu, -1)
        Lm.diagonal().fill_(1.0)
        r = Lm @ torch.triu(a_lu)
        del Lm
        perm = list(range(n))
        for i, p in enumerate((ipiv.long() - 1).cpu().tolist()):
            if p != i:
                perm[i], perm[p] = perm[p], perm[i]
        r -= a_lu0[torch.tensor(perm, device=dev)]
        checks["dget01_ratio"] = r.abs().sum(0).max().item() / (n * a_lu0.abs().sum(0).max().item() * EPS)
        checks["info"] = [int(info_po.item()), int(info_lu.item())]
        del r
        torch.cuda.empty_cache()

    # ---- timed region: device-resident
    sampler = ClockSampler(local_rank)
    launches0 = L.lb200_launch_count()
    L.lb200_profile_gemm(1)
    barrier()
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    clocks = sampler.stop()
    L.lb200_profile_gemm(0)
    ms = e0.elapsed_time(e1)
    launches = int(L.lb200_launch_count() - launches0)
    gms, gfl, gcnt = C.c_double(0), C.c_double(0), C.c_longlong(0)
    L.lb200_profile_gemm_read(C.byref(gms), C.byref(gfl), C.byref(gcnt))
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = t.item()
    value = world * step_flops * args.steps / (ms * 1e-3) * 1e-12

    # ---- per-routine legs (device-resident, CUDA events): every BASELINE config gets its own number
    def time_one(fn, restore, reps=2):
        best = 1e30
        for _ in range(reps):
            restore()
            torch.cuda.synchronize()
            s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s0.record()
            fn()
            s1.record()
            torch.cuda.synchronize()
            best = min(best, s0.elapsed_time(s1) * 1e-3)
        return best
    t_lu = time_one(lambda: lb.dev.getrf(a_lu), lambda: a_lu.copy_(a_lu0))
    t_po = time_one(lambda: lb.dev.potrf("L", a_po), lambda: a_po.copy_(a_po0))
    t_ps = time_one(lambda: lb.dev.potrs("L", a_po, b_po), lambda: b_po.copy_(b_po0), reps=3)       # a_po holds the factor
    l0 = L.lb200_launch_count()
    t_gs = time_one(lambda: lb.dev.getrs("N", a_lu, ipiv, b_po), lambda: b_po.copy_(b_po0), reps=3)  # a_lu holds the LU factors
    solve_launches = int(L.lb200_launch_count() - l0) // 3
    peak = max(L.lb200_fp64_peak_tflops(None, 0, 8, 2, 20000) for _ in range(2))
    # independent FP64 denominator: cuBLAS DGEMM 8192^3 through torch.matmul (checker only, never on the product path)
    ga = torch.randn(8192, 8192, dtype=torch.float64, device=dev)
    gb = torch.randn(8192, 8192, dtype=torch.float64, device=dev)
    gc = torch.empty_like(ga)
    torch.matmul(ga, gb, out=gc)
    t_cublas = time_one(lambda: torch.matmul(ga, gb, out=gc), lambda: None, reps=3)
    peak_cublas = 2.0 * 8192 ** 3 / t_cublas * 1e-12
    del ga, gb, gc
    # BASELINE configs[3]: DGEQRF 32768 x 32768 (SRC/dgeqrf.f:244-267) on the same U(-1,1) matrix, with DQRT01's two ratios
    tau_box = [None]

    def run_qr():
        tau_box[0] = lb.dev.geqrf(a_lu)
    t_qr = time_one(run_qr, lambda: a_lu.copy_(a_lu0))
    qr_checks = {}
    if rank == 0 and not args.no_check:
        q = a_po                                                    # reuse the Cholesky work matrix for Q
        q.copy_(a_lu)
        lb.dev.orgqr(q, tau_box[0])                                 # dqrt01.f:191 (DORGQR)
        r = q.t() @ a_lu0                                           # Q^T A (cuBLAS, checker only)
        r -= torch.triu(a_lu)
        anorm = a_lu0.abs().sum(0).max().item()
        qr_checks["dqrt01_ratio_r"] = r.abs().sum(0).max().item() / (n * anorm * EPS)        # dqrt01.f:203-211
        torch.matmul(q.t(), q, out=r)
        r.diagonal().sub_(1.0)
        qr_checks["dqrt01_ratio_orth"] = r.abs().sum(0).max().item() / (n * EPS)             # dqrt01.f:215-221
        del r, q
        torch.cuda.empty_cache()
    batched = bench_batched(lb, torch, dev, (1 << 20) // world)

    # ---- e2e: Fortran-77 ABI with pinned host buffers (H2D + compute + D2H per step)
    e2e = None
    if not args.no_e2e:
        import numpy as np
        h_lu = torch.empty((n, n), dtype=torch.float64).pin_memory()     # row-major (n,n) buffer == column-major n x n, ld = n
        h_po = torch.empty((n, n), dtype=torch.float64).pin_memory()
        h_b = torch.empty((nrhs, n), dtype=torch.float64).pin_memory()
        h_ipiv = np.zeros(n, dtype=np.int32)
        ts = []
        for i in range(args.steps + 1):
            h_lu.copy_(a_lu0.t())       # a_lu0.t() is the contiguous (n,n) storage of the column-major matrix
            h_po.copy_(a_po0.t())
            h_b.copy_(b_po0.t())
            torch.cuda.synchronize()
            if dist is not None:
                dist.barrier()
            t0 = time.perf_counter()
            i1 = lb.f77.dposv("L", n, nrhs, h_po.data_ptr(), n, h_b.data_ptr(), n)   # DPOTRF + DPOTRS (dposv.f:176-183)
            i2 = 0
            i3 = lb.f77.dgetrf(n, n, h_lu.data_ptr(), n, h_ipiv)
            dt = time.perf_counter() - t0
            if i >= 1:
                ts.append(dt)
            assert i1 == 0 and i2 == 0 and i3 == 0, (i1, i2, i3)
        te = torch.tensor([sum(ts) / len(ts)], device=dev, dtype=torch.float64)
        if dist is not None:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        ipiv_same = bool(np.array_equal(h_ipiv, ipiv.cpu().numpy()))
        nb = n * n * 8
        # Cholesky moves only the 'L' triangle, as 2048-column trapezoids (fortran_abi.cu upload_triangle)
        tri = sum((n - j0) * min(2048, n - j0) * 8 for j0 in range(0, n, 2048))
        e2e = {"value": world * step_flops / te.item() * 1e-12, "unit": "TFLOP/s", "ms_per_step": te.item() * 1e3,
               "h2d_bytes_per_step": nb + tri + n * nrhs * 8, "d2h_bytes_per_step": nb + tri + n * nrhs * 8 + n * 4,
               "api": "dposv_ (= DPOTRF+DPOTRS) and dgetrf_ (Fortran-77 ABI) on pinned host buffers; transfers overlap "
                      "with the factorizations (streamed block columns / split upload)",
               "ipiv_equals_device_run": ipiv_same}
        del h_lu, h_po, h_b

    # ---- e2e with PAGEABLE host buffers (what a Fortran caller's ALLOCATEd array is): same calls, plain numpy memory
    e2e_pageable = None
    if not args.no_e2e and world == 1:
        import numpy as np
        p_lu = np.empty((n, n), dtype=np.float64)
        p_po = np.empty((n, n), dtype=np.float64)
        p_b = np.empty((nrhs, n), dtype=np.float64)
        p_ipiv = np.zeros(n, dtype=np.int32)
        tp = []
        for i in range(2):
            torch.from_numpy(p_lu).copy_(a_lu0.t())
            torch.from_numpy(p_po).copy_(a_po0.t())
            torch.from_numpy(p_b).copy_(b_po0.t())
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            i1 = lb.f77.dposv("L", n, nrhs, p_po.ctypes.data, n, p_b.ctypes.data, n)
            i3 = lb.f77.dgetrf(n, n, p_lu.ctypes.data, n, p_ipiv)
            tp.append(time.perf_counter() - t0)
            assert i1 == 0 and i3 == 0, (i1, i3)
        e2e_pageable = {"value": step_flops / min(tp) * 1e-12, "unit": "TFLOP/s", "ms_per_step": min(tp) * 1e3,
                        "ms_first_call": tp[0] * 1e3,
                        "ipiv_equals_device_run": bool(np.array_equal(p_ipiv, ipiv.cpu().numpy())),
                        "api": "dposv_ and dgetrf_ on pageable (malloc) host arrays"}
        del p_lu, p_po, p_b

    if rank != 0:
        return
    # ---- CPU baseline (oracle port, bounded sample), rank 0 only, N=1 only
    cpu = None
    keep = {}
    if world == 1 and not args.no_cpu:
        ncpu = pick_cpu_sample(25.0, 1)
        tc, fc, _, _ = cpu_reference_sample(ncpu, keep=keep)
        cpu = {"value": fc / tc * 1e-12, "unit": "TFLOP/s", "cores": 1, "kind": "port", "seconds": tc,
               "sample": f"DGESV n={ncpu} nrhs=1: reference DGETRF (NB=64) + DGETRS on reference BLAS loops (oracle port), 1 of {os.cpu_count()} host cores"}
    # ---- BASELINE configs[0] on the GPU: DGESV n=4096, 1 RHS through dgesv_ on the SAME input the CPU leg factored;
    # IPIV must be bit-identical with the oracle's (SRC/dgesv.f:165-172)
    c1 = None
    if world == 1:
        import numpy as np
        from oracle import oracle as O
        n1 = keep.get("n", 4096)
        a1, seed1 = O.random_matrix(n1, n1, SEED)
        x1, _ = O.random_matrix(n1, 1, seed1)
        b1 = np.asfortranarray(a1 @ x1)
        tg = []
        for i in range(3):
            lu1, sol1 = a1.copy(order="F"), b1.copy(order="F")
            t0 = time.perf_counter()
            ipiv1, info1 = lb.f77.gesv(lu1, sol1)
            tg.append(time.perf_counter() - t0)
        d1 = lb.dev.larnv_matrix(n1, n1, SEED, device=dev)
        db = lb.dev.colmajor(n1, 1, device=dev)
        db.copy_(torch.from_numpy(b1))
        d1w = d1.clone()
        pv = [None]

        def c1_dev():
            pv[0], _ = lb.dev.getrf(d1w)
            lb.dev.getrs("N", d1w, pv[0], db)
        t_c1 = time_one(c1_dev, lambda: (d1w.copy_(d1), db.copy_(torch.from_numpy(b1))), reps=3)
        fl1 = flops_getrf(n1) + flops_getrs(n1, 1)
        c1 = {"n": n1, "info": int(info1), "e2e_ms_pageable_host": min(tg) * 1e3, "e2e_tflops": fl1 / min(tg) * 1e-12,
              "device_ms": t_c1 * 1e3, "device_tflops": fl1 / t_c1 * 1e-12,
              "ipiv_equals_oracle": bool(np.array_equal(ipiv1, keep["ipiv"])) if "ipiv" in keep else None,
              "ipiv_device_equals_host_call": bool(np.array_equal(pv[0].cpu().numpy(), ipiv1)),
              "x_rel_err_vs_oracle": float(np.max(np.abs(sol1 - keep["x"])) / np.max(np.abs(keep["x"]))) if "x" in keep else None,
              "x_rel_err_vs_xact": float(np.max(np.abs(sol1 - x1)) / np.max(np.abs(x1)))}
    gemm_tf = (gfl.value / (gms.value * 1e-3) * 1e-12) if gms.value > 0 else None
    traffic, traffic_note = None, "no ncu capture found under profiles/"
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "r02b_gemm_traffic.json")))
        traffic = tr["dram_bytes_read"] + tr["dram_bytes_write"]
        traffic_note = (f"dram__bytes_read.sum + dram__bytes_write.sum of ONE representative trailing-update launch (m=n={tr['m']}, k={tr['k']}; "
                        f"algorithmic {tr['algorithmic_bytes']:.3e} B), read from profiles/r02b_gemm_traffic.json ({tr['source']})")
    except Exception:
        pass
    hbm = 6650.0
    try:
        hbm = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs", hbm)
    except Exception:
        pass
    line = {
        "metric": "DGETRF/DPOTRF FP64 TFLOP/s", "value": value, "unit": "TFLOP/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"per step: DPOTRF('L')+DPOTRS(1 rhs) n={n} SPD (BASELINE configs[1]) then DGETRF n={n} (configs[2]); "
                               "DLARNV(2) seed 1988-1991; in-place, column-major",
                   "n": n, "parallelism": "1 GPU" if world == 1 else f"{world} independent replicas (no data-path collective)",
                   "l2": "inputs (8 GiB per matrix) are far larger than the 126 MB L2; restored from HBM copies every step"},
        "pct_of_fp64_peak": value / world / peak if peak else None,
        "breakdown": {"dgetrf_tflops": flops_getrf(n) / t_lu * 1e-12, "dgetrf_ms": t_lu * 1e3,
                      "dgetrf_frac_of_dmma_peak": flops_getrf(n) / t_lu * 1e-12 / peak,
                      "dgetrf_frac_of_cublas_dgemm": flops_getrf(n) / t_lu * 1e-12 / peak_cublas,
                      "dpotrf_tflops": flops_potrf(n) / t_po * 1e-12, "dpotrf_ms": t_po * 1e3,
                      "dpotrf_frac_of_dmma_peak": flops_potrf(n) / t_po * 1e-12 / peak,
                      "dpotrf_frac_of_cublas_dgemm": flops_potrf(n) / t_po * 1e-12 / peak_cublas,
                      "dgeqrf_tflops": flops_geqrf(n, n) / t_qr * 1e-12, "dgeqrf_ms": t_qr * 1e3,
                      "dgeqrf_frac_of_dmma_peak": flops_geqrf(n, n) / t_qr * 1e-12 / peak,
                      "dgeqrf_frac_of_cublas_dgemm": flops_geqrf(n, n) / t_qr * 1e-12 / peak_cublas,
                      "dpotrs_1rhs_ms": t_ps * 1e3, "dpotrs_hbm_gbs": 8.0 * n * n / t_ps * 1e-9, "dpotrs_frac_of_hbm": 8.0 * n * n / t_ps * 1e-9 / hbm,
                      "dgetrs_1rhs_ms": t_gs * 1e3, "dgetrs_hbm_gbs": 8.0 * n * n / t_gs * 1e-9, "dgetrs_frac_of_hbm": 8.0 * n * n / t_gs * 1e-9 / hbm,
                      "dgetrs_launches": solve_launches},
        "dgeqrf_32768": {"workload": f"DGEQRF {n}x{n} (BASELINE configs[3]), DLARNV(2) seed 1988-1991", "ms": t_qr * 1e3,
                         "tflops": flops_geqrf(n, n) / t_qr * 1e-12, "flops": flops_geqrf(n, n), "checks": qr_checks},
        "dgesv_4096": c1,
        "roofline": {"bound": "tensor", "kernel": "gemm_f64_dmma_kernel<64,64,2,2,...,STAGES=2,BK=16,VAR=1> (trailing updates, DMMA.8x8x4; interior tiles through the lean cp.async loader)",
                     "achieved": gemm_tf, "peak": peak, "unit": "TFLOP/s", "frac": (gemm_tf / peak) if gemm_tf else None,
                     "peak_cublas": peak_cublas, "frac_of_cublas": (gemm_tf / peak_cublas) if gemm_tf else None,
                     "peak_cublas_note": "cuBLAS DGEMM 8192^3 via torch.matmul timed in this run (independent denominator, checker only)",
                     "traffic": traffic, "traffic_note": traffic_note,
                     "launches_timed": int(gcnt.value),
                     "peak_source": "FP64 DMMA.8x8x4 issue-rate peak measured in this run by lb200_fp64_peak_tflops "
                                    "(MEASURED_PEAKS.json carries HBM and bf16 only)"},
        "cpu_baseline": cpu, "e2e": e2e, "e2e_pageable": e2e_pageable, "gpu_launches": launches, "clocks": clocks, "checks": checks,
        "batched_dgetrf_32x32": batched,
    }
    OUT.emit(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--n", type=int, default=32768)
    ap.add_argument("--n-dist", type=int, default=131072, help="order of the single distributed DGETRF when --gpus > 1")
    ap.add_argument("--nb-dist", type=int, default=512)
    ap.add_argument("--grid", default="", help="process grid PxQ of the distributed DGETRF (default 1xN)")
    ap.add_argument("--replicas", action="store_true", help="N > 1: independent n=32768 replicas instead of one distributed matrix")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-check", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    global OUT
    OUT = StdoutToStderr()          # from here on only OUT.emit() reaches stdout
    if world > 1 and not args.replicas:
        run_dist(args, rank, world, local_rank)
        return
    run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
