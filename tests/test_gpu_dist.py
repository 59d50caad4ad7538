"""Multi-GPU parity (needs >= 2 GPUs; skipped otherwise): block-cyclic DGETRF over NCCL vs the single-GPU run."""
import os
import subprocess
import sys

import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("n,nb", [(4096, 512), (3000, 256)])
def test_pgetrf_nccl(n, nb):
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "tests", "_dist_gpu_worker.py"), str(n), str(nb)]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    lines = [l for l in out.stdout.splitlines() if l.startswith("DIST_RESULT")]
    assert out.returncode == 0 and len(lines) == 2, out.stdout[-3000:] + out.stderr[-3000:]
    assert all("ok=1" in l for l in lines), lines


@pytest.mark.parametrize("which,n,nb", [("potrf", 4096, 512), ("potrf", 3000, 256), ("geqrf", 4096, 256), ("geqrf", 3000, 128)])
def test_ppotrf_pgeqrf_nccl(which, n, nb):
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29534", os.path.join(ROOT, "tests", "_dist_gpu_worker.py"), str(n), str(nb), which]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    lines = [l for l in out.stdout.splitlines() if l.startswith("DIST_RESULT")]
    assert out.returncode == 0 and len(lines) == 2, out.stdout[-3000:] + out.stderr[-3000:]
    assert all("ok=1" in l for l in lines), lines


@pytest.mark.parametrize("P,Q,n,nb", [(1, 2, 4096, 512), (2, 1, 4096, 512), (2, 1, 3000, 256), (2, 2, 4096, 256), (2, 2, 5000, 512),
                                      (2, 4, 8192, 512)])
def test_pgetrf2d_nccl(P, Q, n, nb):
    """P x Q block-cyclic DGETRF (lapack_b200/dist2d.py): IPIV identical to the single-GPU factorization"""
    if not torch.cuda.is_available() or torch.cuda.device_count() < P * Q:
        pytest.skip(f"needs >= {P * Q} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={P * Q}", "--master-addr", "127.0.0.1",
           "--master-port", "29535", os.path.join(ROOT, "tests", "_dist2d_gpu_worker.py"), str(P), str(Q), str(n), str(nb)]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    lines = [l for l in out.stdout.splitlines() if l.startswith("DIST2D_RESULT")]
    assert out.returncode == 0 and len(lines) == 2, out.stdout[-3000:] + out.stderr[-3000:]
    assert all("ok=1" in l for l in lines), lines
