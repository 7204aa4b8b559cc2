"""The control logic of the device-driven MPGP iteration (permon_b200/csrc/mpgp_ctl.h: step selection, stopping test, direction mode --
the functions the CUDA kernels run in their prologues) is plain __host__ __device__ code.  tests/ctl_emulator.cpp compiles it with g++
and runs the fused K_A -> K_B -> K_A' -> K_C protocol with the big kernels replaced by loops over host arrays; on the reference's golden
problems the protocol must arrive at the reference's counts.  CPU only: a regression here shows up before any GPU time is spent."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle import oracle_py as O
from permon_b200 import problems as PR

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("ctl") / "libctl_emulator.so")
    subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", os.path.join(HERE, "ctl_emulator.cpp"), "-o", so])
    return C.CDLL(so)


def run(emu, pr, rtol=1e-5, atol=1e-50, max_it=10000):
    op = O.Operator(pr.ia, pr.ja, pr.a)
    maxeig, _ = O.max_eigenvalue(op)
    ia, ja, a = (np.ascontiguousarray(v, dtype=t) for v, t in ((pr.ia, np.int32), (pr.ja, np.int32), (pr.a, np.float64)))
    b = np.ascontiguousarray(pr.b, dtype=np.float64)
    lb = None if pr.lb is None else np.ascontiguousarray(pr.lb, dtype=np.float64)
    ub = None if pr.ub is None else np.ascontiguousarray(pr.ub, dtype=np.float64)
    x = np.ascontiguousarray(pr.x0, dtype=np.float64).copy()
    out = (C.c_int * 6)()
    ptr = lambda v: v.ctypes.data_as(C.c_void_p) if v is not None else None
    emu.ctl_emulate(C.c_int(pr.n), ptr(ia), ptr(ja), ptr(a), ptr(b), ptr(lb), ptr(ub), ptr(x), C.c_double(rtol), C.c_double(atol), C.c_double(1e4),
                    C.c_int(max_it), C.c_double(2.0 / maxeig), out)
    return x, tuple(out)


@pytest.mark.parametrize("name", ["ex1_1", "ex2_1_infinite-true", "ex3_1", "jbearing2_4", "jbearing2_5", "jbearing2_6"])
def test_fused_protocol_reaches_the_golden_counts(emu, golden, name):
    g = golden[name]
    kw = {}
    if g["problem"] == "ex1":
        pr = PR.tutorial_ex1(g["n"])
    elif g["problem"] == "ex2":
        pr = PR.tutorial_ex2(g["n"], g["infinite"])
    elif g["problem"] == "ex3dual":
        pr = PR.tutorial_ex3_dual(g["n"])
    else:
        pr = PR.jbearing2(g["mx"], g["my"])
        kw = dict(rtol=1e-6, atol=1e-8)
    x, counts = run(emu, pr, **kw)
    assert counts == (g["its"], g["nmv"], g["ncg"], g["nexp"], g["nprop"], g["reason"])
    xr, _ = O.mpgp_solve(O.Operator(pr.ia, pr.ja, pr.a), pr.b, O.BoxC(pr.n, pr.lb, pr.ub), pr.x0, O.mpgp_opts(**kw))
    assert np.linalg.norm(x - xr) <= 1e-12 * max(np.linalg.norm(xr), 1e-300)


def test_iteration_limit_and_two_sided_box(emu):
    pr = PR.varcoef3d(10)
    x, counts = run(emu, pr, rtol=1e-8, max_it=100000)
    xr, ro = O.mpgp_solve(O.Operator(pr.ia, pr.ja, pr.a), pr.b, O.BoxC(pr.n, pr.lb, pr.ub), pr.x0, O.mpgp_opts(rtol=1e-8, max_it=100000))
    assert counts == (ro["its"], ro["nmv"], ro["ncg"], ro["nexp"], ro["nprop"], ro["reason"])
    pr = PR.tutorial_ex1(100)
    x, counts = run(emu, pr, max_it=25)
    xr, ro = O.mpgp_solve(O.Operator(pr.ia, pr.ja, pr.a), pr.b, O.BoxC(pr.n, pr.lb, None), pr.x0, O.mpgp_opts(max_it=25))
    assert counts == (ro["its"], ro["nmv"], ro["ncg"], ro["nexp"], ro["nprop"], ro["reason"]) and counts[5] == -3
