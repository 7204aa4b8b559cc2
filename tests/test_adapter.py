"""adapters/qpsb200.c -- the PERMON plug-in that registers the B200 solvers under "mpgp" / "smalxe"  -- compiled against
the mock PETSc of adapters/mock/ and driven by tests/adapter_driver.c, "user code" that follows the reference tutorial's call sequence
(src/tutorials/ex1.c:108-157).  CPU part: it compiles warning-free, fills every slot of _QPSOps, and fails loudly without a GPU.  GPU part:
the reference's ex1 problem solved THROUGH the plug-in boundary reproduces the golden counts of src/tutorials/output/ex1_1.out."""
import os
import re
import subprocess

import numpy as np
import pytest

from permon_b200 import problems as PR

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
LIB = os.path.join(ROOT, "permon_b200", "libpermon_b200.so")
# include/permon/private/qpsimpl.h:12-24 of the reference
QPS_OPS = ["solve", "setup", "destroy", "view", "viewconvergence", "setfromoptions", "reset", "resetstatistics", "isqpcompatible", "monitor",
           "monitorcostfunction"]


@pytest.fixture(scope="module")
def driver(tmp_path_factory):
    d = tmp_path_factory.mktemp("adapter")
    cc = ["/usr/bin/gcc", "-std=c11", "-D_GNU_SOURCE", "-O1", "-Wall", "-Wextra", "-Werror"]
    subprocess.check_call(cc + ["-DPERMON_B200_MOCK_PETSC", "-I" + os.path.join(ROOT, "adapters"), "-c", os.path.join(ROOT, "adapters", "qpsb200.c"), "-o", str(d / "qpsb200.o")])
    subprocess.check_call(cc + ["-c", os.path.join(ROOT, "adapters", "mock", "permon_mock.c"), "-o", str(d / "mock.o")])
    subprocess.check_call(cc + ["-c", os.path.join(HERE, "adapter_driver.c"), "-o", str(d / "driver.o")])
    exe = str(d / "adapter_driver")
    subprocess.check_call(["/usr/bin/gcc", str(d / "driver.o"), str(d / "qpsb200.o"), str(d / "mock.o"), "-ldl", "-o", exe])
    return exe, d


def write_problem(path, pr, be=None):
    with open(path, "wb") as f:
        np.array([pr.n, len(pr.ja), pr.lb is not None, pr.ub is not None, be is not None], dtype=np.int32).tofile(f)
        np.asarray(pr.ia, dtype=np.int32).tofile(f)
        np.asarray(pr.ja, dtype=np.int32).tofile(f)
        for arr in (pr.a, pr.b, pr.x0, pr.lb, pr.ub, be):
            if arr is not None:
                np.asarray(arr, dtype=np.float64).tofile(f)


def run(driver, pr, qtype, options=(), be=None):
    exe, d = driver
    write_problem(d / "problem.bin", pr, be)
    env = dict(os.environ, PERMON_B200_LIBRARY=LIB)
    p = subprocess.run([exe, str(d / "problem.bin"), str(d / "x.bin"), qtype, *options], env=env, capture_output=True, text=True, timeout=600)
    return p


def test_adapter_fills_every_qps_ops_slot():
    src = open(os.path.join(ROOT, "adapters", "qpsb200.c")).read()
    for slot in QPS_OPS:
        assert re.search(r"qps->ops->%s\s*=\s*\w+_B200;" % slot, src), slot
    assert "QPSRegister(QPSMPGP, QPSCreate_MPGP_B200)" in src and "QPSRegister(QPSSMALXE, QPSCreate_SMALXE_B200)" in src
    assert "#if defined(PERMON_B200_HAVE_PETSC)" in src


def test_adapter_without_a_gpu_fails_loudly(driver):
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("a GPU is present")
    except ImportError:
        pass
    p = run(driver, PR.tutorial_ex1(100), "mpgp")
    assert p.returncode != 0
    assert "no CPU execution path" in p.stderr


@pytest.mark.gpu
def test_ex1_through_the_plugin_matches_the_golden_output(driver, golden):
    from oracle import oracle_py as O
    pr = PR.tutorial_ex1(100)
    p = run(driver, pr, "mpgp")
    assert p.returncode == 0, p.stderr
    g = golden["ex1_1"]
    out = p.stdout
    assert f"iterations {g['its']}" in out and "reason 2" in out
    # the lines QPSViewConvergence_MPGP prints (mpgp.c:751-770), byte for byte as in src/tutorials/output/ex1_1.out
    for line in (f"number of Hessian multiplications {g['nmv']}", f"number of CG steps {g['ncg']}", f"number of expansion steps {g['nexp']}",
                 f"number of proportioning steps {g['nprop']}"):
        assert line + "\n" in out, out
    x = np.fromfile(driver[1] / "x.bin")
    xr, ro = O.mpgp_solve(O.Operator(pr.ia, pr.ja, pr.a), pr.b, O.BoxC(pr.n, pr.lb, pr.ub), pr.x0, O.mpgp_opts())
    assert np.linalg.norm(x - xr) <= 1e-9 * np.linalg.norm(xr)


@pytest.mark.gpu
def test_options_and_equality_row_travel_through_the_plugin(driver):
    """-qps_* keys of the (mock) PETSc options database reach the library; a MATONEROW equality constraint is forwarded for SMALXE"""
    from permon_b200 import api as P
    pr = PR.tutorial_ex1(100)
    p = run(driver, pr, "mpgp", ["-qps_mpgp_expansion_type", "gf", "-qps_mpgp_expansion_length_type", "opt"])
    assert p.returncode == 0, p.stderr
    assert "iterations 184" in p.stdout and "number of Hessian multiplications 217" in p.stdout      # src/tutorials/output/ex1_opt.out
    prs = PR.svm_dual(300, d=60, nnz_per_row=6)
    import scipy.sparse as sp
    Z = sp.csr_matrix((prs.a, prs.ja, prs.ia), shape=(300, 60))
    H = (Z @ Z.T + 1e-3 * sp.identity(300)).tocsr()
    H.sort_indices()
    prs.ia, prs.ja, prs.a, prs.second = H.indptr.astype(np.int32), H.indices.astype(np.int32), H.data, None
    p = run(driver, prs, "smalxe", be=prs.B[0])
    assert p.returncode == 0, p.stderr
    x = np.fromfile(driver[1] / "x.bin")
    P.initialize()
    r = P.solve_problem(prs, "smalxe")
    assert f"iterations {r.its}" in p.stdout
    assert np.linalg.norm(x - r.x) <= 1e-12 * max(np.linalg.norm(r.x), 1e-300)
    assert abs(float(prs.B[0] @ x)) <= 1e-4
