"""gloo tests (CPU, world_size 2..6) of the P x Q distributed LU's host logic (lapack_b200/dist2d.py): 2D block-cyclic index
maps, panel gather / return / row broadcast, cross-process-row interchanges, U12 column broadcast, look-ahead.  Local compute
is a mock backend built on the oracle (tests/_dist2d_worker.py); the result must equal the oracle's DGETRF (IPIV exactly)."""
import os
import socket
import subprocess
import sys
import tempfile

import numpy as np
import pytest

from oracle import oracle as O
from lapack_b200.dist2d import BlockCyclic2D, default_grid

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_block_cyclic_2d_maps():
    for (n, nb, P, Q) in ((100, 16, 2, 2), (64, 16, 2, 1), (50, 64, 1, 4), (257, 32, 2, 4), (130, 16, 3, 2)):
        seen = np.zeros((n, n), dtype=int)
        for r in range(P * Q):
            d = BlockCyclic2D(n, nb, P, Q, r)
            rows, cols = d.global_rows(), d.global_cols()
            assert len(rows) == d.mloc and len(cols) == d.nloc
            assert np.all(np.diff(rows) > 0) and np.all(np.diff(cols) > 0)
            seen[np.ix_(rows, cols)] += 1
            for k in range(d.nblocks + 1):
                assert np.all(rows[d.lrow0(k):] >= k * nb) and np.all(rows[:d.lrow0(k)] < k * nb)
                assert np.all(cols[d.lcol0(k):] >= k * nb) and np.all(cols[:d.lcol0(k)] < k * nb)
            for g in rows[::7]:
                bg = g // nb
                assert bg % P == d.p and rows[(bg // P) * nb + g % nb] == g
        assert np.all(seen == 1)
    assert default_grid(8) == (2, 4) and default_grid(4) == (2, 2) and default_grid(2) == (1, 2) and default_grid(1) == (1, 1)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _spawn(P, Q, n, nb, lookahead, extra=()):
    port = _free_port()
    tmp = tempfile.mkdtemp()
    procs = []
    for r in range(P * Q):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE=str(P * Q), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port),
                   OMP_NUM_THREADS="1")
        cmd = [sys.executable, os.path.join(ROOT, "tests", "_dist2d_worker.py"), str(P), str(Q), str(n), str(nb), str(lookahead), tmp]
        procs.append(subprocess.Popen(cmd + [str(x) for x in extra], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT))
    for p in procs:
        out, _ = p.communicate(timeout=300)
        assert p.returncode == 0, out.decode()[-3000:]
    return [dict(np.load(os.path.join(tmp, f"rank{r}.npz"))) for r in range(P * Q)]


@pytest.mark.parametrize("P,Q,n,nb,lookahead", [(1, 2, 96, 16, 1), (2, 1, 96, 16, 1), (2, 2, 100, 16, 1), (2, 2, 64, 16, 0),
                                                (2, 2, 40, 64, 1), (3, 2, 130, 16, 1), (2, 3, 150, 32, 1)])
def test_pgetrf2d_gloo_matches_oracle(P, Q, n, nb, lookahead):
    res = _spawn(P, Q, n, nb, lookahead)
    a, _ = O.random_matrix(n, n, (1988, 1989, 1990, 1991))
    ref = a.copy(order="F")
    ipiv_ref, info_ref = O.dgetrf(ref)
    lu = np.zeros((n, n), order="F")
    for d in res:
        lu[np.ix_(d["rows"], d["cols"])] = d["lu"]
        assert np.array_equal(d["ipiv"], ipiv_ref)                     # IPIV replicated and identical to DGETRF's
        assert int(d["info"]) == info_ref == 0
    assert np.max(np.abs(lu - ref)) < 1e-11
    assert O.dget01(a, lu, ipiv_ref) < O.THRESH


def test_pgetrf2d_gloo_singular_info():
    n, iz = 96, 37
    res = _spawn(2, 2, n, 16, 1, extra=(iz,))
    a, _ = O.random_matrix(n, n, (1988, 1989, 1990, 1991))
    a[:, iz - 1] = 0.0
    ref = a.copy(order="F")
    ipiv_ref, info_ref = O.dgetrf(ref)
    assert info_ref == iz
    for d in res:
        assert int(d["info"]) == iz                                    # dchkge.f:328-347: INFO = IZERO, factorization completed
        assert np.array_equal(d["ipiv"], ipiv_ref)
