"""The drop-in boundary exercised from plain C: tests/ex1_consumer.c is the reference tutorial (src/tutorials/ex1.c:58-157) compiled with cc
against include/permonqps.h and linked to libpermon_b200.so -- no Python, no ctypes in the loop.  Run with the arguments of the reference's
own test spec (ex1.c:161-184) its output, filtered like the spec filters it, must equal the golden files byte for byte."""
import os
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
LIBDIR = os.path.join(ROOT, "permon_b200")


@pytest.fixture(scope="module")
def exe(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("cconsumer") / "ex1_consumer")
    subprocess.check_call(["/usr/bin/gcc", "-std=c11", "-D_GNU_SOURCE", "-Wall", "-Wextra", "-Werror", "-I" + os.path.join(ROOT, "include"),
                           os.path.join(HERE, "ex1_consumer.c"), "-L" + LIBDIR, "-lpermon_b200", "-Wl,-rpath," + LIBDIR, "-lm", "-o", out])
    return out


def grep(text, *pats):
    return "".join(l for l in text.splitlines(keepends=True) if any(p in l for p in pats))


def test_c_program_builds_and_fails_loudly_without_a_gpu(exe):
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("a GPU is present")
    except ImportError:
        pass
    p = subprocess.run([exe, "-n", "100"], capture_output=True, text=True, timeout=120)
    assert p.returncode == 97 and "no CPU execution path" in p.stderr          # PETSC_ERR_GPU


SPEC = {   # src/tutorials/ex1.c:161-184
    "ex1_1": [],
    "ex1_opt": ["-qps_mpgp_expansion_type", "gf", "-qps_mpgp_expansion_length_type", "opt"],
    "ex1_optapprox": ["-qps_mpgp_expansion_type", "g", "-qps_mpgp_expansion_length_type", "optapprox"],
    "ex1_bb": ["-qps_mpgp_expansion_type", "gfgr", "-qps_mpgp_expansion_length_type", "bb"],
    "ex1_projcg": ["-qps_mpgp_expansion_type", "projcg"],
}


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(SPEC))
def test_c_program_output_equals_the_golden_file(exe, name):
    p = subprocess.run([exe, "-n", "100", "-qps_view_convergence", "-qp_chain_view_kkt", *SPEC[name]], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stderr
    assert "same_storage = 1" in p.stderr                                        # the user's x array is the solution storage (qp.c:1987-1991)
    gold = open(os.path.join(HERE, "golden", "out", name + ".out")).read()
    ours = grep(p.stdout, "CONVERGED", "number", "r =")
    # the CONVERGED line and the four counters must be identical; a KKT residual may differ in its last printed digit (summation order)
    go, oo = gold.splitlines(), ours.splitlines()
    assert len(go) == len(oo), (ours, gold)
    for a, b in zip(oo, go):
        if a.startswith("r ="):
            assert a.split("=")[1] == b.split("=")[1], (a, b)
        else:
            assert a == b, (a, b)
    print(name, "byte-identical lines:", sum(a == b for a, b in zip(oo, go)), "of", len(go))
