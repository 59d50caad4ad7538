"""world_size-2/3 gloo tests (CPU) of the multi-GPU LU driver's host logic: block-cyclic index maps, look-ahead
schedule, panel/pivot broadcast.  Local compute is a mock backend built on the oracle (tests/_dist_worker.py)."""
import os
import socket
import subprocess
import sys
import tempfile

import numpy as np
import pytest

from oracle import oracle as O
from lapack_b200.dist import BlockCyclic1D

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_block_cyclic_maps():
    for (n, nb, world) in ((100, 16, 3), (64, 16, 2), (50, 64, 4), (257, 32, 8)):
        seen = []
        for r in range(world):
            d = BlockCyclic1D(n, nb, world, r)
            cols = [d.global_col(c) for c in range(d.local_cols())]
            assert cols == sorted(cols)
            seen += cols
            for k in d.local_blocks():
                assert d.owner(k) == r
                assert d.global_col(d.local_offset(k)) == k * nb
            for k in range(d.nblocks):
                after = d.first_local_col_after(k)
                before = d.local_cols_before(k)
                assert all(c >= (k + 1) * nb for c in cols[after:]) and all(c < (k + 1) * nb for c in cols[:after])
                assert all(c < k * nb for c in cols[:before]) and all(c >= k * nb for c in cols[before:])
        assert sorted(seen) == list(range(n))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("world,n,nb,lookahead", [(2, 96, 16, 1), (2, 100, 16, 0), (3, 130, 32, 1), (2, 40, 64, 1)])
def test_pgetrf_gloo_matches_oracle(world, n, nb, lookahead):
    port = _free_port()
    with tempfile.TemporaryDirectory() as tmp:
        procs = []
        for r in range(world):
            env = dict(os.environ, RANK=str(r), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
            procs.append(subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "_dist_worker.py"), str(n), str(nb),
                                           str(lookahead), tmp], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT))
        for p in procs:
            out, _ = p.communicate(timeout=180)
            assert p.returncode == 0, out.decode()[-2000:]
        a, _ = O.random_matrix(n, n, (1988, 1989, 1990, 1991))
        ref = a.copy(order="F")
        ipiv_ref, info_ref = O.dgetrf(ref)
        lu = np.zeros((n, n), order="F")
        for r in range(world):
            d = np.load(os.path.join(tmp, f"rank{r}.npz"))
            lu[:, d["cols"]] = d["lu"]
            assert np.array_equal(d["ipiv"], ipiv_ref)
            assert int(d["info"]) == info_ref == 0
        assert np.max(np.abs(lu - ref)) < 1e-11
        assert O.dget01(a, lu, ipiv_ref) < O.THRESH


def _spawn(world, n, nb, lookahead, which, extra=()):
    """returns the list of per-rank npz dicts"""
    port = _free_port()
    tmp = tempfile.mkdtemp()
    procs = []
    for r in range(world):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        cmd = [sys.executable, os.path.join(ROOT, "tests", "_dist_worker.py"), str(n), str(nb), str(lookahead), tmp, which]
        procs.append(subprocess.Popen(cmd + [str(x) for x in extra], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT))
    for p in procs:
        out, _ = p.communicate(timeout=180)
        assert p.returncode == 0, out.decode()[-2000:]
    return [dict(np.load(os.path.join(tmp, f"rank{r}.npz"))) for r in range(world)]


@pytest.mark.parametrize("world,n,nb,lookahead", [(2, 96, 16, 1), (3, 100, 16, 0), (2, 130, 32, 1)])
def test_ppotrf_gloo_matches_oracle(world, n, nb, lookahead):
    res = _spawn(world, n, nb, lookahead, "potrf")
    s, _ = O.spd_matrix(n, (1988, 1989, 1990, 1991))
    ref = s.copy(order="F")
    assert O.dpotrf("L", ref) == 0
    fac = np.zeros((n, n), order="F")
    for d in res:
        fac[:, d["cols"]] = d["lu"]
        assert int(d["info"]) == 0
    assert np.all(fac[np.triu_indices(n, 1)] == -1.0e10)              # the other triangle was never touched
    assert np.max(np.abs(np.tril(fac) - np.tril(ref))) < 1e-11
    assert O.dpot01("L", s, np.asfortranarray(np.tril(fac))) < O.THRESH


def test_ppotrf_gloo_not_positive_definite():
    n, iz = 96, 41
    res = _spawn(2, n, 16, 1, "potrf", extra=(iz,))
    for d in res:
        assert int(d["info"]) == iz                                   # dchkpo.f:313-344: INFO = IZERO


@pytest.mark.parametrize("world,n,nb,lookahead", [(2, 96, 16, 1), (3, 100, 16, 0), (2, 130, 32, 1)])
def test_pgeqrf_gloo_matches_oracle(world, n, nb, lookahead):
    res = _spawn(world, n, nb, lookahead, "geqrf")
    a, _ = O.random_matrix(n, n, (1988, 1989, 1990, 1991))
    af = np.zeros((n, n), order="F")
    for d in res:
        af[:, d["cols"]] = d["lu"]
        assert np.array_equal(d["tau"], res[0]["tau"])                # tau is replicated
    tau = res[0]["tau"]
    r1, r2 = O.dqrt01(a, af, tau)
    assert r1 < O.THRESH and r2 < O.THRESH
    ref = a.copy(order="F")
    O.set_nb(geqrf=nb, nx=1)
    try:
        tau_ref, _, _ = O.dgeqrf(ref)
    finally:
        O.set_nb()
    assert np.max(np.abs(tau - tau_ref)) < 1e-11
    assert np.max(np.abs(af - ref)) < 1e-10


def test_pgetrf_gloo_singular_info():
    n, iz = 96, 23
    res = _spawn(2, n, 16, 1, "getrf", extra=(iz,))
    a, _ = O.random_matrix(n, n, (1988, 1989, 1990, 1991))
    a[:, iz - 1] = 0.0
    ref = a.copy(order="F")
    ipiv_ref, info_ref = O.dgetrf(ref)
    assert info_ref == iz
    for d in res:
        assert int(d["info"]) == iz
        assert np.array_equal(d["ipiv"], ipiv_ref)
