"""Edge cases of the CUDA path against the oracle: ragged sizes (n not a multiple of the 256-row tile, odd n => scalar
K_B), upper-bound-only and two-sided boxes, bounds with +-infinity entries, trivial solves (already converged, max_it 0),
tiny problems, an iterate that starts outside the box, user monitor + custom convergence test (host-sync mode),
unaligned device pointers (plain-load SpMV path)."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import oracle_py as O
from permon_b200 import problems as PR


@pytest.fixture(scope="module")
def P():
    from permon_b200 import api
    if api.device_count() == 0:
        pytest.fail("no CUDA device")
    api.initialize()
    yield api
    api.options_clear()


def oracle(pr, **kw):
    op = O.Operator(pr.ia, pr.ja, pr.a)
    x, r = O.mpgp_solve(op, pr.b, O.BoxC(pr.n, pr.lb, pr.ub, pr.is_), pr.x0, O.mpgp_opts(**kw))
    r["objective"] = O.objective(op, pr.b, x)
    return x, r


def lap1d(n, shift=0.0):
    import scipy.sparse as sp
    A = sp.diags([-np.ones(n - 1), (2.0 + shift) * np.ones(n), -np.ones(n - 1)], [-1, 0, 1], format="csr")
    A.sort_indices()
    return A.indptr.astype(np.int32), A.indices.astype(np.int32), A.data.astype(np.float64)


def make(n, lb=None, ub=None, b=None, x0=None, shift=0.05):
    ia, ja, a = lap1d(n, shift)
    return PR.QPProblem(f"lap1d_{n}", n, 0, n, ia, ja, a, b if b is not None else -np.ones(n) / n, lb, ub, x0 if x0 is not None else np.zeros(n))


def compare(P, pr, opts="-qps_rtol 1e-10", **okw):
    okw.setdefault("rtol", 1e-10)
    r = P.solve_problem(pr, "mpgp", opts)
    xr, ro = oracle(pr, **okw)
    assert r.reason == ro["reason"], (r.reason, ro["reason"])
    assert abs(r.its - ro["its"]) <= max(2, 0.03 * ro["its"]), (r.its, ro["its"])
    nx = max(np.linalg.norm(xr), 1e-300)
    assert np.linalg.norm(r.x - xr) <= 1e-7 * nx, np.linalg.norm(r.x - xr) / nx
    if abs(ro["objective"]) > 0:
        assert abs(r.objective - ro["objective"]) <= 1e-10 * abs(ro["objective"])
    return r, ro


@pytest.mark.parametrize("n", [1, 2, 3, 255, 256, 257, 511, 1001, 4099])
def test_ragged_sizes_lower_bound(P, n):
    x = np.linspace(0, 1, n)
    lb = -0.02 - 0.05 * np.sin(6 * x) ** 2
    compare(P, make(n, lb=lb))


@pytest.mark.parametrize("n", [300, 1025])
def test_upper_bound_only_and_two_sided(P, n):
    x = np.linspace(0, 1, n)
    ub = 0.02 + 0.05 * np.sin(5 * x) ** 2
    compare(P, make(n, ub=ub, b=np.ones(n) / n))                       # lb == NULL branch of QPCProject/QPCGrads (qpcbox.c:298-303)
    lb = -0.03 - 0.04 * np.cos(7 * x) ** 2
    pr = make(n, lb=lb, ub=ub, b=np.sin(9 * x) / n * 3)
    compare(P, pr)
    lb2, ub2 = lb.copy(), ub.copy()
    lb2[::3] = PR.PETSC_NINFINITY                                        # infinite entries are skipped by QPCFeas (qpcbox.c:126,132)
    ub2[1::3] = PR.PETSC_INFINITY
    compare(P, make(n, lb=lb2, ub=ub2, b=np.sin(9 * x) / n * 3))


def test_infeasible_start_is_projected(P):
    n = 700
    lb = np.full(n, -0.01)
    x0 = np.full(n, -5.0)                                                # QPCProject at mpgp.c:497
    x0[::2] = 3.0
    r, ro = compare(P, make(n, lb=lb, x0=x0))
    assert np.all(r.x >= lb - 1e-15)


def test_trivial_solves(P):
    n = 400
    pr = make(n, lb=np.full(n, -1.0), b=np.zeros(n))                     # b = 0, x0 = 0: converged at iteration 0 by atol? rnorm = 0
    r = P.solve_problem(pr, "mpgp", "")
    xr, ro = oracle(pr)
    assert (r.its, r.reason) == (ro["its"], ro["reason"]) == (0, 3)      # rnorm = 0 <= ttol and < atol -> CONVERGED_ATOL
    assert r.counts["nmv"] == 1
    pr = make(n, lb=np.full(n, -1.0))
    r = P.solve_problem(pr, "mpgp", "-qps_max_it 0")                     # i > max_it -> DIVERGED_ITS after exactly 1 iteration (qps.c:688)
    xr, ro = oracle(pr, max_it=0)
    assert (r.its, r.reason) == (ro["its"], ro["reason"]) == (1, -3)
    assert not r.solved
    assert np.linalg.norm(r.x - xr) <= 1e-12 * np.linalg.norm(xr)


def test_divergence_tolerance_and_direct_alpha(P):
    n = 500
    pr = make(n, lb=np.full(n, -1e-3))
    r = P.solve_problem(pr, "mpgp", "-qps_divtol 1e-3")                  # rnorm >= divtol*||b|| at iteration 0 -> DIVERGED_DTOL
    xr, ro = oracle(pr, divtol=1e-3)
    assert (r.its, r.reason) == (ro["its"], ro["reason"])
    assert r.reason == -4
    r = P.solve_problem(pr, "mpgp", "-qps_mpgp_alpha 0.3 -qps_mpgp_alpha_direct 1 -qps_rtol 1e-9 -qps_mpgp_gamma 0.7")
    xr, ro = oracle(pr, alpha_user=0.3, alpha_direct=1, rtol=1e-9, gamma=0.7)
    assert r.reason == ro["reason"] and abs(r.its - ro["its"]) <= max(2, 0.03 * ro["its"])
    assert np.linalg.norm(r.x - xr) <= 1e-7 * np.linalg.norm(xr)


def test_monitor_and_custom_convergence_test(P):
    """user callbacks force one host sync per iteration; they must see the same sequence as the device-driven run"""
    pr = PR.obstacle2d(64, -100.0)
    base = P.solve_problem(pr, "mpgp", "-qps_rtol 1e-7")
    seen = []
    r = P.solve_problem(pr, "mpgp", "-qps_rtol 1e-7", monitor=lambda q, it, rn: seen.append((it, rn)))
    assert r.its == base.its and r.counts == base.counts
    assert [s[0] for s in seen] == list(range(base.its + 1))
    assert seen[-1][1] == pytest.approx(base.rnorm, rel=1e-12)
    assert np.array_equal(r.x, base.x)                                    # deterministic reductions: bit-identical iterate
    # custom test: stop at iteration 17 with CONVERGED_ITS
    CONV = C.CFUNCTYPE(C.c_int, C.c_void_p, C.POINTER(C.c_int))

    def conv(qps, reason):
        it = C.c_int()
        P.lib().QPSGetIterationNumber(C.c_void_p(qps), C.byref(it))
        reason[0] = 4 if it.value >= 17 else 0
        return 0

    cb = CONV(conv)
    rr = P.solve_problem(pr, "mpgp", "", keep=True) if False else None
    A = P.MatCreateAIJ(pr.ia, pr.ja, pr.a)
    b, x, lb = P.VecFromArray(pr.b.copy()), P.VecFromArray(pr.x0.copy()), P.VecFromArray(pr.lb.copy())
    qp = P.QPCreate(); P.QPSetOperator(qp, A); P.QPSetRhs(qp, b); P.QPSetInitialVector(qp, x); P.QPSetBox(qp, None, lb, None)
    qps = P.QPSCreate(); P.QPSSetType(qps, "mpgp"); P.QPSSetQP(qps, qp)
    P.call("QPSSetConvergenceTest", qps, cb, None, None)
    P.QPSSolve(qps)
    assert P.QPSGetIterationNumber(qps) == 17 and P.QPSGetConvergedReason(qps) == 4 and P.QPIsSolved(qp)
    xr, ro = oracle(pr, max_it=16)                                        # same 17 iterations in the oracle
    assert np.linalg.norm(P.VecGetArray(x) - xr) <= 1e-9 * np.linalg.norm(xr)
    P.QPSDestroy(qps); P.QPDestroy(qp)
    for v in (b, x, lb):
        P.VecDestroy(v)
    P.MatDestroy(A)


def test_unaligned_device_arrays_use_plain_load_path(P):
    """caller-owned device arrays that are only 8-byte aligned cannot be fed to the TMA engine: same answer via plain loads"""
    torch = pytest.importorskip("torch")
    pr = PR.obstacle2d(80, -100.0)
    dev = torch.device("cuda")

    def off(arr, dtype):   # tensor view starting 8 bytes into an allocation
        t = torch.empty(arr.size + 2, dtype=dtype, device=dev)
        k = 2 if dtype == torch.int32 else 1
        v = t[k:k + arr.size]
        v.copy_(torch.from_numpy(arr))
        return t, v

    keep = []
    views = {}
    for name, arr, dt in (("ia", pr.ia, torch.int32), ("ja", pr.ja, torch.int32), ("a", pr.a, torch.float64), ("b", pr.b, torch.float64),
                          ("lb", pr.lb, torch.float64), ("x", pr.x0, torch.float64)):
        t, v = off(arr, dt)
        keep.append(t)
        views[name] = v
        assert v.data_ptr() % 16 == 8
    torch.cuda.synchronize()
    A = P.MatCreateAIJFromDevicePointers(pr.n, pr.n, views["ia"].data_ptr(), views["ja"].data_ptr(), views["a"].data_ptr())
    b, lb, x = (P.VecFromDevicePointer(views[k].data_ptr(), pr.n) for k in ("b", "lb", "x"))
    qp = P.QPCreate(); P.QPSetOperator(qp, A); P.QPSetRhs(qp, b); P.QPSetInitialVector(qp, x); P.QPSetBox(qp, None, lb, None)
    qps = P.QPSCreate(); P.QPSSetType(qps, "mpgp"); P.QPSSetQP(qps, qp); P.QPSSetTolerances(qps, rtol=1e-8)
    P.QPSSolve(qps)
    P.synchronize()
    ref = P.solve_problem(pr, "mpgp", "-qps_rtol 1e-8")
    # other kernels (plain-load SpMV, scalar K_B) => other reduction grids => last-bit differences in the dots
    assert abs(P.QPSGetIterationNumber(qps) - ref.its) <= max(2, 0.02 * ref.its)
    xu = views["x"].cpu().numpy()
    assert np.linalg.norm(xu - ref.x) <= 1e-7 * np.linalg.norm(ref.x)
    P.QPSDestroy(qps); P.QPDestroy(qp)
    for v in (b, lb, x):
        P.VecDestroy(v)
    P.MatDestroy(A)


@pytest.mark.parametrize("prob,opt,okw", [
    ("ex1", "-qps_mpgp_fallback 1 -qps_mpgp_expansion_type g -qps_mpgp_alpha 3.0", dict(fallback=1, exptype="g", alpha_user=3.0)),                      # 1 fallback taken
    ("ex1", "-qps_mpgp_fallback2 1 -qps_mpgp_expansion_type g -qps_mpgp_expansion_length_type opt", dict(fallback2=1, exptype="g", explengthtype="opt")),  # 10 cost increases
    ("ob32", "-qps_mpgp_fallback2 1 -qps_mpgp_expansion_type gf -qps_mpgp_alpha 3.5", dict(fallback2=1, exptype="gf", alpha_user=3.5)),
    ("ex1", "-qps_mpgp_expansion_type ggr -qps_mpgp_expansion_length_type bb", dict(exptype="ggr", explengthtype="bb")),                                 # the reference diverges (DTOL)
    ("ex1", "-qps_mpgp_expansion_type gfgr -qps_mpgp_expansion_length_type optapprox", dict(exptype="gfgr", explengthtype="optapprox"))])
def test_fallback_and_remaining_expansion_variants(P, prob, opt, okw):
    """mpgp.c:582-611 (fallback / fallback2) and the remaining expansion-type x length-type pairs, generic GPU driver;
    whatever the reference does on these inputs (including diverging) the GPU path must do too"""
    pr = PR.tutorial_ex1(100) if prob == "ex1" else PR.obstacle2d(32, -100.0)
    r = P.solve_problem(pr, "mpgp", opt)
    xr, ro = oracle(pr, **okw)
    assert r.reason == ro["reason"], (r.reason, ro["reason"])
    assert abs(r.its - ro["its"]) <= max(5, 0.08 * ro["its"]), (r.its, ro["its"])
    assert abs(r.counts["nmv"] - ro["nmv"]) <= max(8, 0.08 * ro["nmv"])
    if ro["reason"] > 0:
        assert np.linalg.norm(r.x - xr) <= 1e-4 * np.linalg.norm(xr)


@pytest.mark.parametrize("fused", [True, False])
def test_power_method_on_a_singular_hessian(P, fused, monkeypatch):
    """v = 1 lies in the null space of a pure-Neumann Laplacian; the reference restarts from a PETSCRAND48 vector
    (permonmatutils.c:493-502).  Fused one-kernel power step and the generic Mat/Vec path, both against the oracle."""
    import scipy.sparse as sp
    if fused:
        monkeypatch.delenv("PERMON_B200_NOFUSEDPOWER", raising=False)
    else:
        monkeypatch.setenv("PERMON_B200_NOFUSEDPOWER", "1")
    n = 5000
    L = sp.diags([-np.ones(n - 1), 2 * np.ones(n), -np.ones(n - 1)], [-1, 0, 1]).tolil()
    L[0, 0] = 1.0
    L[n - 1, n - 1] = 1.0
    L = L.tocsr()
    L.sort_indices()
    A = P.MatCreateAIJ(L.indptr.astype(np.int32), L.indices.astype(np.int32), L.data)
    lam = P.MatGetMaxEigenvalue(A, tol=1e-7, maxits=300)
    lam_ref, _ = O.max_eigenvalue(O.Operator(L.indptr, L.indices, L.data), tol=1e-7, maxits=300)
    P.MatDestroy(A)
    assert np.isfinite(lam) and 3.0 < lam <= 4.0
    assert lam == pytest.approx(lam_ref, rel=1e-11)


def test_bound_chop_tol(P):
    """-qps_mpgp_bound_chop_tol (mpgp.c:379-382): VecFilter zeroes the bound entries that are closer to 0 than the tolerance, in the user's
    vectors, before the solve.  Oracle: the same filter applied to the bounds up front."""
    pr = PR.obstacle2d(48, -100.0)
    rng = np.random.default_rng(5)
    pr.lb = np.where(rng.random(pr.n) < 0.5, -1e-4 * rng.random(pr.n), -0.02 - 0.01 * rng.random(pr.n))   # half of the bounds within 1e-3 of zero
    tol = 1e-3
    r = P.solve_problem(pr, "mpgp", f"-qps_rtol 1e-8 -qps_mpgp_bound_chop_tol {tol}")
    lbf = np.where(np.abs(pr.lb) < tol, 0.0, pr.lb)
    assert np.count_nonzero(lbf == 0.0) > pr.n // 3
    xr, ro = O.mpgp_solve(O.Operator(pr.ia, pr.ja, pr.a), pr.b, O.BoxC(pr.n, lbf, None), pr.x0, O.mpgp_opts(rtol=1e-8))
    assert r.reason == ro["reason"] == 2
    assert abs(r.its - ro["its"]) <= max(2, 0.02 * ro["its"])
    assert np.linalg.norm(r.x - xr) <= 1e-7 * np.linalg.norm(xr)
    assert np.all(r.x >= lbf - 1e-12)


def test_vec_invalidate_like_reference_tests_ex4(P):
    """src/tests/ex4.c: a vector is valid, VecInvalidate makes it invalid, any later write (VecSet here, VecCopy / VecGetArray as well)
    makes it valid again (permonvecutils.c:266-326: the invalid mark remembers the object state)"""
    import ctypes as C
    v = P.VecCreate(4)
    P.call("VecSet", v, C.c_double(1.0))
    assert not P.VecIsInvalidated(v)
    P.call("VecInvalidate", v)
    assert P.VecIsInvalidated(v)
    P.call("VecSet", v, C.c_double(1.0))
    assert not P.VecIsInvalidated(v)
    P.call("VecInvalidate", v)
    w = P.VecFromArray(np.arange(4.0))
    P.call("VecCopy", w, v)                      # a copy INTO the vector validates it
    assert not P.VecIsInvalidated(v)
    assert np.array_equal(P.VecGetArray(v), np.arange(4.0))
    P.call("VecInvalidate", w)
    P.call("VecCopy", w, v)                      # ... unless the source itself is invalid
    assert P.VecIsInvalidated(v)
    P.VecDestroy(v), P.VecDestroy(w)
