"""World-size-2/3 CPU tests (gloo) of the multi-GPU host logic: row partition, diagonal / off-diagonal split, ghost
list (garray), neighbour discovery and pack lists of MatCreateMPIAIJWithArrays -- the layout of PETSc's Mat_MPIAIJ
(include/permon/private/petsc/mpiaij.h:49-83 in the reference).  The set-up-time host exchange runs over gloo
callbacks; the halo exchange itself is emulated with gloo following the plan, and must reproduce the global SpMV."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, size, port, kind, ret):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=size)
    from permon_b200 import api as P
    from permon_b200 import problems as PR

    def agi(v):
        out = [None] * size
        dist.all_gather_object(out, int(v))
        return out

    def agv(data):
        out = [None] * size
        dist.all_gather_object(out, bytes(data))
        return out

    P.initialize()
    P.comm_set_host_exchange(size, rank, agi, agv)
    if kind == "2d":
        N = 12
        n_glob = N * N
        starts = PR.row_partition(n_glob, size, align=N)
        full = PR.obstacle2d(N)
        pr = PR.obstacle2d(N, rows=(starts[rank], starts[rank + 1]))
    else:
        N = 6
        n_glob = N ** 3
        starts = PR.row_partition(n_glob, size, align=N * N)
        full = PR.varcoef3d(N)
        pr = PR.varcoef3d(N, rows=(starts[rank], starts[rank + 1]))
    r0, r1 = starts[rank], starts[rank + 1]
    A = P.MatCreateAIJ(pr.ia, pr.ja, pr.a, ncols_local=pr.n)
    info = P.MatGetHaloInfo(A)
    # 1. garray = sorted unique non-local columns
    cols = np.unique(pr.ja)
    expect = [int(c) for c in cols if c < r0 or c >= r1]
    assert info["garray"] == expect, (info["garray"], expect)
    # 2. neighbours own contiguous slices of the ghost list
    for q, (a, b) in zip(info["neigh"], zip(info["recv_off"][:-1], info["recv_off"][1:])):
        assert all(starts[q] <= g < starts[q + 1] for g in info["garray"][a:b])
    # 3. boundary rows = rows with at least one ghost column
    nb = sum(1 for r in range(pr.n) if np.any((pr.ja[pr.ia[r]:pr.ia[r + 1]] < r0) | (pr.ja[pr.ia[r]:pr.ia[r + 1]] >= r1)))
    assert info["nboundary"] == nb
    # 3b. the split itself: diagonal block (local columns) + off-diagonal rows (ghost-buffer columns) put together again give back
    #     every local row exactly, entries in their original order within each block
    dia, dja, da, oia, oja, oa, orow = P.MatGetHostSplit(A, pr.n)
    garr = np.asarray(info["garray"], dtype=np.int64)
    opos = {int(r): k for k, r in enumerate(orow)}
    assert len(orow) == nb and np.all(np.diff(orow) > 0)
    for r in range(pr.n):
        cols, vals = np.asarray(pr.ja[pr.ia[r]:pr.ia[r + 1]], dtype=np.int64), np.asarray(pr.a[pr.ia[r]:pr.ia[r + 1]])
        loc = (cols >= r0) & (cols < r1)
        assert np.array_equal(dja[dia[r]:dia[r + 1]] + r0, cols[loc]) and np.array_equal(da[dia[r]:dia[r + 1]], vals[loc])
        if r in opos:
            k = opos[r]
            assert np.array_equal(garr[oja[oia[k]:oia[k + 1]]], cols[~loc]) and np.array_equal(oa[oia[k]:oia[k + 1]], vals[~loc])
        else:
            assert loc.all()
    # 4. emulate the halo exchange with gloo following the plan, then the split SpMV must equal the global one
    rng = np.random.default_rng(7)
    xg = rng.standard_normal(n_glob)
    xl = xg[r0:r1]
    sendbufs = {q: xl[info["send_idx"][a:b]].copy() for q, (a, b) in zip(info["neigh"], zip(info["send_off"][:-1], info["send_off"][1:]))}
    allsend = [None] * size
    dist.all_gather_object(allsend, sendbufs)
    ghost = np.zeros(len(info["garray"]))
    for q, (a, b) in zip(info["neigh"], zip(info["recv_off"][:-1], info["recv_off"][1:])):
        ghost[a:b] = allsend[q][rank]
    assert np.array_equal(ghost, xg[info["garray"]])
    import scipy.sparse as sp
    Aloc = sp.csr_matrix((pr.a, pr.ja, pr.ia), shape=(pr.n, n_glob))
    Afull = sp.csr_matrix((full.a, full.ja, full.ia), shape=(n_glob, n_glob))
    assert np.allclose(Aloc @ xg, (Afull @ xg)[r0:r1], rtol=1e-14, atol=1e-14)
    # 5. layouts of vectors created on the communicator
    v = P.VecFromArray(xl.copy())
    lo, hi = __import__("ctypes").c_int(), __import__("ctypes").c_int()
    P.call("VecGetOwnershipRange", v, __import__("ctypes").byref(lo), __import__("ctypes").byref(hi))
    assert (lo.value, hi.value) == (r0, r1)
    # 6. arithmetic still needs a GPU: no CPU fallback on the multi-rank path either
    if P.device_count() == 0:
        w = P.VecDuplicate(v)
        with pytest.raises(P.PermonError):
            P.MatMult(A, v, w)
    # 7. a short and wide row-partitioned matrix (equality rows B_E assembled like any MPIAIJ: rank r owns row r, global columns) is accepted
    #    without a halo plan (its pattern is not symmetric); its arithmetic needs a GPU as well
    import scipy.sparse as sp2
    Bfull = np.vstack([np.ones(n_glob), np.arange(n_glob, dtype=float)])
    mine = [rank] if rank < 2 else []
    S = sp2.csr_matrix(Bfull[mine, :]) if mine else sp2.csr_matrix((0, n_glob))
    BE = P.MatCreateAIJ(S.indptr.astype(np.int32), S.indices.astype(np.int32), S.data.astype(np.float64), ncols_local=pr.n)
    if P.device_count() == 0:
        t = P.VecFromArray(np.zeros(len(mine)))
        with pytest.raises(P.PermonError):
            P.MatMult(BE, v, t)
    P.MatDestroy(BE)
    dist.barrier()
    dist.destroy_process_group()
    ret[rank] = True


@pytest.mark.parametrize("size,kind", [(2, "2d"), (3, "2d"), (2, "3d")])
def test_halo_plan_gloo(size, kind):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    mgr = ctx.Manager()
    ret = mgr.dict()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, size, port, kind, ret)) for r in range(size)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=240)
    for p in procs:
        if p.is_alive():
            p.kill()
            pytest.fail("worker hung")
        assert p.exitcode == 0
    assert all(ret.get(r) for r in range(size))
