"""The two linear QPS types next to the hot path against their oracle restatements:
"ksp"  = QPSKSP  (unpreconditioned CG, default type of an unconstrained QP, qps.c:448-451; PETSc's KSPCG restated: parity unpinned)
"pcpg" = QPSPCPG (src/qps/impls/pcpg/pcpg.c:49-131), with and without a non-zero right-hand side of the equality constraint."""
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


def solve(P, pr, b, x0, qps_type=None, G=None, c=None, options="-qps_rtol 1e-9"):
    P.options_clear()
    P.call("PetscOptionsInsertString", None, options.encode())
    A = P.MatCreateAIJ(pr.ia, pr.ja, pr.a)
    vb, vx = P.VecFromArray(b.copy()), P.VecFromArray(x0.copy())
    qp = P.QPCreate()
    P.QPSetOperator(qp, A), P.QPSetRhs(qp, vb), P.QPSetInitialVector(qp, vx)
    keep = []
    if G is not None:
        import scipy.sparse as sp
        S = sp.csr_matrix(G)
        BE = P.MatCreateAIJ(S.indptr, S.indices, S.data, ncols_local=pr.n)
        cE = P.VecFromArray(np.asarray(c, dtype=np.float64).copy()) if c is not None else None
        P.QPSetEq(qp, BE, cE)
        keep += [BE, cE]
    qps = P.QPSCreate()
    if qps_type:
        P.QPSSetType(qps, qps_type)
    P.QPSSetQP(qps, qp)
    P.QPSSetFromOptions(qps)
    P.QPSSolve(qps)
    out = dict(x=P.VecGetArray(vx).copy(), its=P.QPSGetIterationNumber(qps), reason=P.QPSGetConvergedReason(qps), rnorm=P.QPSGetResidualNorm(qps),
               type=P.QPSGetType(qps), solved=P.QPIsSolved(qp))
    P.QPSDestroy(qps), P.QPDestroy(qp), P.VecDestroy(vb), P.VecDestroy(vx), P.MatDestroy(A)
    for k in keep:
        if k is not None:
            (P.MatDestroy if k is keep[0] else P.VecDestroy)(k)
    return out


@pytest.mark.parametrize("driver", ["auto", "generic"])
@pytest.mark.parametrize("N", [24, 96])
def test_unconstrained_qp_defaults_to_cg(P, N, driver):
    """driver auto: the fused CG (SpMV + p.Ap in one kernel, x / r update + r.r in one kernel); generic: one kernel per KSPCG call"""
    pr = PR.obstacle2d(N)
    rng = np.random.default_rng(N)
    b, x0 = rng.standard_normal(pr.n), rng.standard_normal(pr.n)
    r = solve(P, pr, b, x0, options=f"-qps_rtol 1e-9 -qps_ksp_b200_driver {driver}")     # no type given, no constraints -> "ksp"
    assert r["type"] == "ksp" and r["solved"]
    xo, ro = O.cg_solve(O.Operator(pr.ia, pr.ja, pr.a), b, x0, O.lin_opts(rtol=1e-9))
    assert r["reason"] == ro["reason"] == 2
    assert abs(r["its"] - ro["its"]) <= max(1, 0.02 * ro["its"])
    assert np.linalg.norm(r["x"] - xo) <= 1e-7 * np.linalg.norm(xo)
    assert r["rnorm"] == pytest.approx(ro["rnorm"], rel=0.5)


def test_cg_iteration_limit_and_trivial_solve(P):
    pr = PR.obstacle2d(32)
    b = np.ones(pr.n)
    r = solve(P, pr, b, np.zeros(pr.n), "ksp", options="-qps_rtol 1e-14 -qps_max_it 7")
    xo, ro = O.cg_solve(O.Operator(pr.ia, pr.ja, pr.a), b, None, O.lin_opts(rtol=1e-14, max_it=7))
    assert (r["its"], r["reason"]) == (ro["its"], ro["reason"]) == (7, -3)
    assert np.allclose(r["x"], xo, rtol=1e-12, atol=1e-14)
    r = solve(P, pr, np.zeros(pr.n), np.zeros(pr.n), "ksp")       # b = 0, x0 = 0: converged at iteration 0
    assert r["its"] == 0 and r["reason"] > 0


@pytest.mark.parametrize("with_c", [False, True])
def test_pcpg_against_oracle_and_saddle_point(P, with_c):
    import scipy.sparse as sp
    import scipy.sparse.linalg as sl
    pr = PR.obstacle2d(48)
    n = pr.n
    rng = np.random.default_rng(5)
    b, x0 = rng.standard_normal(n), np.zeros(n)
    G = np.vstack([np.ones(n), np.sin(np.arange(n) * 0.01), (np.arange(n) % 7 == 0).astype(float)])
    c = np.array([1.0, -2.0, 0.5]) if with_c else None
    r = solve(P, pr, b, x0, "pcpg", G=G, c=c, options="-qps_rtol 1e-10")
    xo, ro = O.pcpg_solve(O.Operator(pr.ia, pr.ja, pr.a), b, G, c, x0, O.lin_opts(rtol=1e-10))
    assert r["reason"] == ro["reason"] == 2
    assert abs(r["its"] - ro["its"]) <= max(1, 0.02 * ro["its"])
    assert np.linalg.norm(r["x"] - xo) <= 1e-7 * np.linalg.norm(xo)
    A = sp.csr_matrix((pr.a, pr.ja, pr.ia), shape=(n, n))
    K = sp.bmat([[A, sp.csr_matrix(G.T)], [sp.csr_matrix(G), None]]).tocsc()
    ref = sl.spsolve(K, np.concatenate([b, c if with_c else np.zeros(3)]))[:n]
    assert np.linalg.norm(r["x"] - ref) <= 1e-6 * np.linalg.norm(ref)
    assert np.max(np.abs(G @ r["x"] - (c if with_c else 0.0))) <= 1e-8


def test_incompatible_qp_is_rejected(P):
    pr = PR.obstacle2d(16)
    with pytest.raises(P.PermonError):
        P.solve_problem(pr, "ksp")                                 # a box-constrained QP is not a KSP problem (qpsksp.c:205-219)
    with pytest.raises(P.PermonError):
        solve(P, pr, np.ones(pr.n), np.zeros(pr.n), "pcpg")        # PCPG needs an equality constraint (pcpg.c:13-22)
