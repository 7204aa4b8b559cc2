"""QPTOrthonormalizeEq + QPTEnforceEqByProjector (the classical SMALBE form: orthonormal equality rows, Hessian
P A P, SMALXE with the injected eigenvalue estimate) and the equality-only variant (P A solved by CG), against the oracle's
restatement of the same chain.  The reference's own outputs for these transforms (ex3, FETI) need QPTDualize + MUMPS, so parity is
pinned to the oracle and to the untransformed problem's solution only."""
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


def eq_rows(n, m, seed=3):
    rng = np.random.default_rng(seed)
    B = np.zeros((m, n))
    B[0] = 1.0
    if m > 1:
        B[1, n // 3:] = rng.random(n - n // 3)
    if m > 2:
        B[2] = np.sin(np.arange(n) * 0.05)
    return B


def build_qp(P, pr, B, c, with_box=True):
    import scipy.sparse as sp
    A = P.MatCreateAIJ(pr.ia, pr.ja, pr.a)
    vb, vx = P.VecFromArray(np.asarray(pr.b, dtype=np.float64).copy()), P.VecFromArray(np.zeros(pr.n))
    qp = P.QPCreate()
    P.QPSetOperator(qp, A), P.QPSetRhs(qp, vb), P.QPSetInitialVector(qp, vx)
    keep = [A, vb, vx]
    if with_box:
        vl = P.VecFromArray(np.asarray(pr.lb, dtype=np.float64).copy())
        P.QPSetBox(qp, None, vl, None)
        keep.append(vl)
    S = sp.csr_matrix(B)
    BE = P.MatCreateAIJ(S.indptr, S.indices, S.data, ncols_local=pr.n)
    cE = P.VecFromArray(np.asarray(c, dtype=np.float64).copy()) if c is not None else None
    P.QPSetEq(qp, BE, cE)
    keep += [BE, cE]
    return qp, vx, keep


def run(P, qp, vx, options):
    P.options_clear()
    P.call("PetscOptionsInsertString", None, options.encode())
    qps = P.QPSCreate()
    P.QPSSetQP(qps, qp)
    P.QPSSetFromOptions(qps)
    P.QPSSolve(qps)
    out = dict(x=P.VecGetArray(vx).copy(), its=P.QPSGetIterationNumber(qps), reason=P.QPSGetConvergedReason(qps), type=P.QPSGetType(qps))
    if out["type"] == "smalxe":
        out["stats"] = P.QPSSMALXEGetStatistics(qps)
    P.QPSDestroy(qps)
    return out


@pytest.mark.parametrize("orth", ["gs", "cholesky"])
@pytest.mark.parametrize("with_c", [False, True])
def test_orthonormalize_project_smalxe(P, orth, with_c):
    pr = PR.obstacle2d(32)
    n = pr.n
    B = eq_rows(n, 2)
    pr.b = np.asarray(pr.b) * (1.0 + 40.0 * np.sin(np.arange(n) * 0.013) ** 2)      # not in the range of B' (else x = 0 solves it)
    c = np.array([-0.6 * n, -0.3 * B[1].sum()]) if with_c else None               # pushes 67 dofs onto the obstacle
    qp, vx, keep = build_qp(P, pr, B, c)
    P.QPTOrthonormalizeEq(qp, orth)
    P.QPTEnforceEqByProjector(qp)
    r = run(P, qp, vx, "-qps_rtol 1e-8")
    assert r["type"] == "smalxe" and r["reason"] > 0

    # the same chain with the oracle's pieces
    op = O.Operator(pr.ia, pr.ja, pr.a)
    TB, Tc, _ = O.orth_rows(B, c, orth)
    bx = O.BoxC(n, pr.lb, None)
    if with_c:
        xt, bh, lbh, _ = O.homogenize(op, pr.b, bx, TB, Tc)
    else:
        xt, bh, lbh = np.zeros(n), np.asarray(pr.b, dtype=np.float64), np.asarray(pr.lb, dtype=np.float64)
    opP = O.Operator(pr.ia, pr.ja, pr.a)
    opP.set_projector(TB, 2)
    xo, ro = O.smalxe_solve(opP, O.apply_P(TB, bh), O.BoxC(n, lbh, None), TB, None, np.zeros(n), O.smalxe_opts(rtol=1e-8))
    xo = xo + xt
    assert ro["reason"] == r["reason"]
    # iteration counts move with the summation order of the dot products: the projector adds two more
    # reductions per Hessian application, and the oracle itself spans +-10 % over its thread counts on this chain
    assert abs(r["its"] - ro["outer_its"]) <= max(2, 0.15 * ro["outer_its"])
    assert abs(r["stats"]["inner_iter_accu"] - ro["inner_its_accu"]) <= max(8, 0.15 * ro["inner_its_accu"])
    assert np.linalg.norm(r["x"] - xo) <= 1e-6 * np.linalg.norm(xo)
    # and the untransformed problem solved by plain SMALXE has the same solution (looser: two different algorithms at rtol 1e-8)
    x2, _ = O.smalxe_solve(O.Operator(pr.ia, pr.ja, pr.a), pr.b, bx, B, c, np.zeros(n), O.smalxe_opts(rtol=1e-8))
    assert np.linalg.norm(r["x"] - x2) <= 1e-5 * np.linalg.norm(x2)
    assert np.max(np.abs(B @ r["x"] - (c if with_c else 0.0))) <= 1e-6 * n
    assert np.min(r["x"] - pr.lb) >= -1e-12
    if with_c:
        assert np.sum(r["x"] - pr.lb < 1e-12) == np.sum(xo - pr.lb < 1e-12) > 10          # same (non-trivial) active set
    P.QPDestroy(qp)


def test_equality_only_project_then_cg(P):
    pr = PR.obstacle2d(40)
    n = pr.n
    rng = np.random.default_rng(9)
    pr.b = rng.standard_normal(n)
    B = eq_rows(n, 3)
    qp, vx, keep = build_qp(P, pr, B, None, with_box=False)
    P.QPTEnforceEqByProjector(qp)                      # only equality constraints: they are eliminated, child = (P A, P b)
    r = run(P, qp, vx, "-qps_rtol 1e-10")
    assert r["type"] == "ksp" and r["reason"] > 0
    op = O.Operator(pr.ia, pr.ja, pr.a)
    op.set_projector(B, 1)
    xo, ro = O.cg_solve(op, O.apply_P(B, pr.b), None, O.lin_opts(rtol=1e-10))
    assert abs(r["its"] - ro["its"]) <= max(1, 0.02 * ro["its"])
    assert np.linalg.norm(r["x"] - xo) <= 1e-7 * np.linalg.norm(xo)
    assert np.max(np.abs(B @ r["x"])) <= 1e-8 * n
    P.QPDestroy(qp)


@pytest.mark.parametrize("lag", [False, True])
@pytest.mark.parametrize("with_c", [False, True])
def test_implicit_orthonormalisation_smalxeon(P, lag, with_c):
    """QPTOrthonormalizeEq(MAT_ORTH_IMPLICIT) (qptransform.c:566-636, permonmatorth.c:176-192): B_E becomes a dummy without MatMult, the
    QPPF applies G'G as Q = G'(GG')^-1 G, and SMALXE switches to the u'B'Bu norm update (smalxe.c:265-285) -- with
    -qps_smalxe_norm_update_lag to its lagged variant (:289-370).  Against the oracle's restatement of the same chain, and against the
    explicitly (Cholesky) orthonormalised problem, which has the same solution."""
    pr = PR.obstacle2d(32)
    n = pr.n
    B = eq_rows(n, 3)
    pr.b = np.asarray(pr.b) * (1.0 + 40.0 * np.sin(np.arange(n) * 0.013) ** 2)
    c = np.array([-0.6 * n, -0.3 * B[1].sum(), 0.05 * n]) if with_c else None
    lag_opts = " -qps_smalxe_norm_update_lag -qps_smalxe_norm_update_lag_offset 3 -qps_smalxe_norm_update_lag_start 4 -qps_smalxe_norm_update_lag_step 2 -qps_smalxe_norm_update_lag_end 8" if lag else ""
    qp, vx, keep = build_qp(P, pr, B, c)
    P.QPTOrthonormalizeEq(qp, "implicit")
    r = run(P, qp, vx, "-qps_type smalxe -qps_rtol 1e-8" + lag_opts)
    assert r["type"] == "smalxe" and r["reason"] > 0
    P.QPDestroy(qp)

    op = O.Operator(pr.ia, pr.ja, pr.a)
    bx = O.BoxC(n, pr.lb, None)
    kw = dict(rtol=1e-8, implicit_orth=1)
    if lag:
        kw.update(lag_enabled=1, lag_offset=3, Jstart=4, Jstep=2, Jend=8)
    xo, ro = O.smalxe_solve(op, pr.b, bx, B, c, np.zeros(n), O.smalxe_opts(**kw))
    assert ro["reason"] == r["reason"]
    assert abs(r["its"] - ro["outer_its"]) <= max(2, 0.15 * ro["outer_its"])
    assert abs(r["stats"]["inner_iter_accu"] - ro["inner_its_accu"]) <= max(8, 0.15 * ro["inner_its_accu"])
    assert np.linalg.norm(r["x"] - xo) <= 1e-6 * np.linalg.norm(xo)
    assert np.max(np.abs(B @ r["x"] - (c if with_c else 0.0))) <= 1e-6 * n
    assert np.min(r["x"] - pr.lb) >= -1e-12
    # the explicitly orthonormalised chain (Cholesky) is the same problem
    qp2, vx2, keep2 = build_qp(P, pr, B, c)
    P.QPTOrthonormalizeEq(qp2, "cholesky")
    r2 = run(P, qp2, vx2, "-qps_type smalxe -qps_rtol 1e-8")
    assert np.linalg.norm(r["x"] - r2["x"]) <= 1e-5 * np.linalg.norm(r2["x"])
    P.QPDestroy(qp2)
