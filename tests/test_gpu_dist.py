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
