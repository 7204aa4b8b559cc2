"""CPU-side checks of the drop-in boundary: the shared library loads, exports every symbol include/permon_b200.h
declares, keeps the PETSc/PERMON conventions that do not need a device, and FAILS LOUDLY (no CPU fallback) when a
computation is requested without a CUDA device."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "permon_b200.h")


@pytest.fixture(scope="module")
def P():
    from permon_b200 import build
    build.build()
    from permon_b200 import api
    api.lib()
    api.initialize()
    return api


def declared_symbols():
    text = open(HEADER).read()
    names = re.findall(r"^(?:extern )?PERMON_EXTERN\s+[A-Za-z_][A-Za-z0-9_ ]*?\s*\**\s*([A-Za-z_][A-Za-z0-9_]*)\s*\(", text, flags=re.M)
    return sorted(set(names))


def test_every_declared_symbol_is_exported(P):
    lib = P.lib()
    names = declared_symbols()
    assert len(names) > 200
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    for must in ("QPCreate", "QPSetOperator", "QPSetRhs", "QPSetBox", "QPSetEq", "QPSCreate", "QPSSetType", "QPSSolve",
                 "QPSSetFromOptions", "QPCProject", "QPCGrads", "QPCGradReduced", "QPCFeas", "QPPFApplyGtG", "MatGetMaxEigenvalue",
                 "QPSSMALXEGetInnerQPS", "QPSMPGPSetAlpha", "PermonInitialize"):
        assert must in names
    C.c_void_p.in_dll(lib, "PETSC_COMM_WORLD")


def test_header_compiles_as_c(tmp_path):
    src = tmp_path / "t.c"
    src.write_text('#include "permon_b200.h"\nint main(void){ QP qp = 0; QPS qps = 0; (void)qp; (void)qps; return PETSC_SUCCESS; }\n')
    subprocess.check_call(["/usr/bin/gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), "-c", str(src), "-o", str(tmp_path / "t.o")])
    for compat in ("permonqp.h", "permonqps.h", "permonqpc.h", "permonqppf.h", "permonmat.h", "permonsys.h", "permonvec.h"):
        s2 = tmp_path / ("c_" + compat.replace(".h", ".c"))
        s2.write_text(f'#include <{compat}>\nint f(void){{ return (int)sizeof(QPS); }}\n')
        subprocess.check_call(["/usr/bin/gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), "-c", str(s2), "-o", str(tmp_path / "c.o")])


def test_host_side_object_conventions(P):
    """reference counting, getters returning borrowed pointers, defaults (qps.c:73-76, smalxe.c:1203), options DB"""
    x = P.VecFromArray(np.arange(5, dtype=np.float64))
    assert P.VecGetLocalSize(x) == 5
    assert np.array_equal(P.VecGetArray(x), np.arange(5.0))     # host-valid: no device needed
    qp = P.QPCreate()
    P.QPSetInitialVector(qp, x)
    assert P.QPGetSolutionVector(qp).value == x.value            # the user's Vec IS the solution storage (qp.c:1987-1991)
    assert not P.QPIsSolved(qp)
    qps = P.QPSCreate()
    rtol, atol, dtol, maxit = C.c_double(), C.c_double(), C.c_double(), C.c_int()
    P.call("QPSGetTolerances", qps, C.byref(rtol), C.byref(atol), C.byref(dtol), C.byref(maxit))
    assert (rtol.value, atol.value, dtol.value, maxit.value) == (1e-5, 1e-50, 1e4, 10000)
    P.QPSSetType(qps, "smalxe")
    P.call("QPSGetTolerances", qps, None, None, None, C.byref(maxit))
    assert maxit.value == 100
    P.QPSSetType(qps, "mpgp")
    a, t = P.QPSMPGPGetAlpha(qps)
    assert a == -1 and t == 0                                     # PETSC_DECIDE -> 2.0/maxeig at set-up (mpgp.c:419)
    P.options_clear()
    P.options_set("-qps_rtol", "1e-7")
    P.options_set("-qps_mpgp_gamma", "0.5")
    P.QPSSetFromOptions(qps)
    P.call("QPSGetTolerances", qps, C.byref(rtol), None, None, None)
    g = C.c_double()
    P.call("QPSMPGPGetGamma", qps, C.byref(g))
    assert rtol.value == 1e-7 and g.value == 0.5
    with pytest.raises(P.PermonError) as e:
        P.QPSSetTolerances(qps, rtol=2.0)                        # qps.c: rtol must be < 1
    assert e.value.code == 63
    with pytest.raises(P.PermonError) as e:
        P.QPSSetType(qps, "tao")                                 # out of scope on this path
    assert e.value.code == 86
    P.options_clear()
    P.QPSDestroy(qps)
    P.QPDestroy(qp)
    P.VecDestroy(x)
    assert qps.value is None and qp.value is None and x.value is None   # Destroy nulls the handle (qps.c:340-362)


def test_solver_types_and_compatibility_rules(P):
    """QPSRegister'ed types and their QPSIsQPCompatible rules are host logic: mpgp needs a box and no equality (mpgp.c:695-711), ksp
    an unconstrained QP (qpsksp.c:205-219), pcpg an equality constraint and no box (pcpg.c:13-22); QPSKSP only offers CG here"""
    def compat(qps, qp):
        f = C.c_int()
        P.call("QPSIsQPCompatible", qps, qp, C.byref(f))
        return bool(f.value)
    x, lb = P.VecFromArray(np.zeros(5)), P.VecFromArray(np.zeros(5))
    qp = P.QPCreate()
    P.QPSetInitialVector(qp, x)
    qps = P.QPSCreate()
    for t in ("ksp", "pcpg", "mpgp", "smalxe"):
        P.QPSSetType(qps, t)
        assert P.QPSGetType(qps) == t
    P.QPSSetType(qps, "ksp")
    assert compat(qps, qp)
    t = C.c_char_p()
    P.call("QPSKSPGetType", qps, C.byref(t))
    assert t.value == b"cg"
    P.call("QPSKSPSetType", qps, b"cg")
    with pytest.raises(P.PermonError) as e:
        P.call("QPSKSPSetType", qps, b"gmres")
    assert e.value.code == 56                                    # PETSC_ERR_SUP
    P.QPSSetType(qps, "pcpg")
    assert not compat(qps, qp)                                    # no equality constraint
    P.QPSetBox(qp, None, lb, None)
    assert not compat(qps, qp)
    P.QPSSetType(qps, "ksp")
    assert not compat(qps, qp)                                    # box-constrained
    P.QPSSetType(qps, "mpgp")
    assert compat(qps, qp)
    with pytest.raises(P.PermonError) as e:
        P.call("QPSKSPGetType", qps, C.byref(t))                  # "This is a QPSKSP specific routine!"
    assert e.value.code == 56
    P.QPSDestroy(qps), P.QPDestroy(qp), P.VecDestroy(x), P.VecDestroy(lb)


def test_no_cpu_fallback(P):
    if P.device_count() > 0:
        pytest.skip("a CUDA device is present")
    from permon_b200 import problems as PR
    with pytest.raises(P.PermonError) as e:
        P.solve_problem(PR.tutorial_ex1(20), "mpgp")
    assert e.value.code == 97 and "no CPU execution path" in str(e.value)
    v = P.VecFromArray(np.ones(4))
    with pytest.raises(P.PermonError):
        P.VecNorm(v)
    P.VecDestroy(v)


def test_product_does_not_touch_the_oracle():
    """the product package must never import / link / execute anything under oracle/"""
    pkg = os.path.join(ROOT, "permon_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cpp", ".cu", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle_py" not in text and "liboracle" not in text and "permon_oracle" not in text, f
    out = subprocess.run(["ldd", os.path.join(pkg, "libpermon_b200.so")], capture_output=True, text=True).stdout
    assert "oracle" not in out
