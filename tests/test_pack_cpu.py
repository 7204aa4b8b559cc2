"""Packed tile format (permon_b200/csrc/pack.cpp) checked on the CPU: the blob is decoded here, in Python, and must give back the CSR
matrix exactly -- same columns, same value bits, same order inside every row.  No GPU, no arithmetic in the library."""
import ctypes as C

import numpy as np
import pytest

from permon_b200 import api as P
from permon_b200 import problems as PR

TR = 256


def pack(ia, ja, a):
    ia = np.ascontiguousarray(ia, dtype=np.int32)
    ja = np.ascontiguousarray(ja, dtype=np.int32)
    a = np.ascontiguousarray(a, dtype=np.float64)
    n = len(ia) - 1
    blob, off = C.POINTER(C.c_ubyte)(), C.POINTER(C.c_uint32)()
    nt, nc = C.c_int(), C.c_int()
    rc = P.lib().PermonB200PackTiles(C.c_int(n), ia.ctypes.data_as(C.c_void_p), ja.ctypes.data_as(C.c_void_p), a.ctypes.data_as(C.c_void_p),
                                     C.byref(blob), C.byref(off), C.byref(nt), C.byref(nc))
    assert rc == 0
    if not blob:
        return None
    o = np.ctypeslib.as_array(off, shape=(nt.value + 1,)).copy()
    b = np.ctypeslib.as_array(blob, shape=(int(o[-1]) * 16,)).copy()
    P.lib().PermonB200PackFree(blob, off)
    return b, o, nc.value


def up(v, a):
    return (v + a - 1) // a * a


def decode(b, o, n):
    """-> (ia, ja, a) rebuilt from the blob"""
    ia, ja, a = [0], [], []
    for t in range(len(o) - 1):
        p = b[int(o[t]) * 16:int(o[t + 1]) * 16]
        kind, nd = p[0:4].view(np.uint16)
        nnz = int(p[4:8].view(np.uint32)[0])
        ulen, nrows = (int(v) for v in p[8:12].view(np.uint16))
        pad = int(p[12:16].view(np.uint32)[0])          # skip code + 1, 0 = rows are not padded
        assert nrows == min(TR, n - t * TR)
        q = 16
        if kind == 2:
            # stencil tile: pattern of nd (delta, value) pairs + one presence byte per row
            L = int(nd)
            assert 1 <= L <= 8 and ulen == L and pad == 0
            val = p[q:q + 8 * L].view(np.float64)
            q += up(L, 2) * 8
            dlt = p[q:q + 4 * L].view(np.int32)
            q += up(L, 4) * 4
            masks = p[q:q + nrows].astype(np.int64)
            assert masks.max() < (1 << L)
            bits = (masks[:, None] >> np.arange(L)[None, :]) & 1          # [nrows, L], storage order = pattern order
            rr, jj = np.nonzero(bits)
            assert len(rr) == nnz
            ja.append(rr + t * TR + dlt[jj])
            a.append(val[jj])
            ro = np.concatenate([[0], np.cumsum(bits.sum(axis=1))])
        elif kind == 1:
            val = p[q:q + 8 * nd].view(np.float64)
            q += up(int(nd), 2) * 8
            dlt = p[q:q + 4 * nd].view(np.int32)
            q += up(int(nd), 4) * 4
            if ulen == 0xFFFF:
                ro = p[q:q + 2 * (nrows + 1)].view(np.uint16).astype(np.int64)
                q += up(nrows + 1, 8) * 2
            else:
                ro = np.arange(nrows + 1, dtype=np.int64) * ulen
            ncodes = int(ro[-1])
            codes = p[q:q + ncodes]
            assert ncodes == 0 or codes.max() < nd
            rows = np.repeat(np.arange(nrows), np.diff(ro)) + t * TR
            if pad:
                assert pad == nd and ulen <= 8 and val[pad - 1] == 0.0 and dlt[pad - 1] == 0
                real = codes != pad - 1
                # padding only at the end of a row
                m = real.reshape(nrows, ulen)
                assert np.all(m[:, :-1] | ~m[:, 1:])
                rows, codes = rows[real], codes[real]
                ro = np.concatenate([[0], np.cumsum(m.sum(axis=1))])
            assert len(codes) == nnz
            ja.append(rows + dlt[codes])
            a.append(val[codes])
        else:
            assert nd == 0 and ulen == 0xFFFF
            a.append(p[q:q + 8 * nnz].view(np.float64))
            q += up(nnz, 2) * 8
            ja.append(p[q:q + 4 * nnz].view(np.int32).astype(np.int64))
            q += up(nnz, 4) * 4
            ro = p[q:q + 2 * (nrows + 1)].view(np.uint16).astype(np.int64)
            assert ro[-1] == nnz
        ia.extend((ia[-1] - ro[0] + ro[1:]).tolist())
    return np.array(ia), np.concatenate(ja) if ja else np.zeros(0), np.concatenate(a) if a else np.zeros(0)


def cases():
    rng = np.random.default_rng(3)
    pr = PR.obstacle2d(70)
    yield "stencil5", pr.ia, pr.ja, pr.a, "all"
    pr = PR.obstacle3d(17)
    yield "stencil7", pr.ia, pr.ja, pr.a, "all"
    pr = PR.varcoef3d(16)
    yield "varcoef", pr.ia, pr.ja, pr.a, "all"
    pr = PR.obstacle2d(64)
    yield "distinct_values", pr.ia, pr.ja, pr.a * (1 + rng.random(len(pr.a))), "none"
    a = pr.a.copy()
    a[pr.ia[1000]:pr.ia[1700]] *= 1 + rng.random(pr.ia[1700] - pr.ia[1000])
    yield "mixed", pr.ia, pr.ja, a, "some"
    import scipy.sparse as sp
    S = sp.random(3000, 3000, density=0.001, random_state=1, format="csr")
    S.data[:] = rng.integers(1, 3, size=len(S.data)).astype(float)
    S.sort_indices()
    yield "ragged_with_empty_rows", S.indptr, S.indices, S.data, "any"
    yield "signed_zero_and_nan_bits", np.array([0, 2, 4] + [4] * 62), np.array([0, 1, 0, 1]), np.array([0.0, -0.0, np.inf, 1.0]), "all"


@pytest.mark.parametrize("stencil", ["stencil_form", "coded_form"])
@pytest.mark.parametrize("case", list(cases()), ids=lambda c: c[0])
def test_blob_decodes_to_the_same_csr(case, stencil, monkeypatch):
    monkeypatch.setenv("PERMON_B200_PACK_STENCIL", "1" if stencil == "stencil_form" else "0")
    _, ia, ja, a, expect = case
    n = len(ia) - 1
    out = pack(ia, ja, a)
    assert out is not None
    b, o, ncoded = out
    ntiles = (n + TR - 1) // TR
    if expect == "all":
        assert ncoded == ntiles
    elif expect == "none":
        assert ncoded == 0
    elif expect == "some":
        assert 0 < ncoded < ntiles
    ia2, ja2, a2 = decode(b, o, n)
    assert np.array_equal(ia2, np.asarray(ia, dtype=np.int64))
    assert np.array_equal(ja2, np.asarray(ja, dtype=np.int64))
    assert np.array_equal(np.asarray(a2).view(np.uint64), np.ascontiguousarray(a, dtype=np.float64).view(np.uint64))   # bit pattern, -0.0 != 0.0


def tile_kinds(b, o):
    return [int(b[int(o[t]) * 16:int(o[t]) * 16 + 2].view(np.uint16)[0]) for t in range(len(o) - 1)]


def test_constant_coefficient_stencils_become_stencil_tiles(monkeypatch):
    """every tile of the obstacle Hessians (C1/C2/C3) is a pattern + one presence byte per row; the two-material operator of C5 and
    matrices with distinct values are not"""
    monkeypatch.setenv("PERMON_B200_PACK_STENCIL", "1")
    for pr in (PR.obstacle2d(70), PR.obstacle3d(17), PR.obstacle3d(24)):
        b, o, ncoded = pack(pr.ia, pr.ja, pr.a)
        assert set(tile_kinds(b, o)) == {2} and ncoded == len(o) - 1
        assert len(b) + 4 * len(o) < 0.04 * (12 * len(pr.a) + 4 * (pr.n + 1))
    pr = PR.varcoef3d(16)
    b, o, _ = pack(pr.ia, pr.ja, pr.a)
    assert 1 in set(tile_kinds(b, o))
    pr = PR.obstacle2d(64, scaled=True)
    b, o, ncoded = pack(pr.ia, pr.ja, pr.a)
    assert set(tile_kinds(b, o)) == {0} and ncoded == 0


def test_ragged_short_rows_are_padded_to_a_common_length(monkeypatch):
    monkeypatch.setenv("PERMON_B200_PACK_STENCIL", "0")
    pr = PR.obstacle3d(24)                       # lines of 24 rows: every tile holds boundary rows with 4..6 entries
    b, o, ncoded = pack(pr.ia, pr.ja, pr.a)
    npad = 0
    for t in range(len(o) - 1):
        p = b[int(o[t]) * 16:int(o[t]) * 16 + 16]
        npad += int(p[12:16].view(np.uint32)[0]) != 0
        assert int(p[8:10].view(np.uint16)[0]) != 0xFFFF        # no tile needs the row-offset table
    assert npad > 0 and ncoded == len(o) - 1


def test_stencil_matrix_stream_shrinks(monkeypatch):
    monkeypatch.setenv("PERMON_B200_PACK_STENCIL", "0")
    pr = PR.obstacle2d(128)
    b, o, _ = pack(pr.ia, pr.ja, pr.a)
    csr = 12 * len(pr.a) + 4 * (pr.n + 1)
    assert len(b) + 4 * len(o) < 0.12 * csr


def test_rows_too_long_for_a_tile_are_not_packed():
    n = 256
    ia = np.arange(n + 1) * 300            # 76800 non-zeros in one tile: beyond 16-bit row offsets
    ja = np.tile(np.arange(300), n)
    assert pack(ia, ja, np.ones(len(ja))) is None
