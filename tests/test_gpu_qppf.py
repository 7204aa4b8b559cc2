"""QPPF (src/qppf/interface/qppf.c): the projector factory on G = B_E -- Q, P, G^T G shell operators, the coarse problem (G G^T) and its
explicit-inverse variant (-qppf_explicit), -qppf_redundancy, alpha_tilde -- against dense numpy algebra.  m = 6 equality rows: the un-fused
route (more rows than the fused rank-m update keeps in registers)."""
import ctypes as C

import numpy as np
import pytest
import scipy.sparse as sp

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def P():
    from permon_b200 import api
    if api.device_count() == 0:
        pytest.fail("no CUDA device")
    api.initialize()
    yield api
    api.options_clear()


def make(P, n=500, m=6, seed=5):
    rng = np.random.default_rng(seed)
    G = rng.standard_normal((m, n))
    G[0] = 1.0
    S = sp.csr_matrix(G)
    Gm = P.MatCreateAIJ(S.indptr, S.indices, S.data, ncols_local=n)
    pf = C.c_void_p()
    P.call("QPPFCreate", P.world(), C.byref(pf))
    P.call("QPPFSetG", pf, Gm)
    return G, Gm, pf, rng


def apply(P, fn, pf, v, nout):
    x, y = P.VecFromArray(v.copy()), P.VecFromArray(np.zeros(nout))
    P.call(fn, pf, x, y)
    out = P.VecGetArray(y).copy()
    P.VecDestroy(x), P.VecDestroy(y)
    return out


@pytest.mark.parametrize("explicit", [False, True])
def test_qppf_operators_against_dense_algebra(P, explicit):
    G, Gm, pf, rng = make(P)
    m, n = G.shape
    P.options_clear()
    if explicit:
        P.call("PetscOptionsInsertString", None, b"-qppf_explicit -qppf_redundancy 2")
        P.call("QPPFSetFromOptions", pf)
    P.call("QPPFSetUp", pf)
    GGt = G @ G.T
    Q = G.T @ np.linalg.solve(GGt, G)
    v = rng.standard_normal(n)
    tol = 1e-11 * np.linalg.norm(v) * np.linalg.cond(GGt)
    assert np.linalg.norm(apply(P, "QPPFApplyQ", pf, v, n) - Q @ v) <= tol
    at = C.c_void_p()
    P.call("QPPFGetAlphaTilde", pf, C.byref(at))
    assert at.value, "alpha_tilde exists after an application of Q"
    assert np.linalg.norm(P.VecGetArray(at) - np.linalg.solve(GGt, G @ v)) <= tol
    assert np.linalg.norm(apply(P, "QPPFApplyP", pf, v, n) - (v - Q @ v)) <= tol
    assert np.linalg.norm(apply(P, "QPPFApplyGtG", pf, v, n) - G.T @ (G @ v)) <= 1e-12 * np.linalg.norm(G.T @ (G @ v))
    assert np.linalg.norm(apply(P, "QPPFApplyHalfQ", pf, v, m) - np.linalg.solve(GGt, G @ v)) <= tol
    r = rng.standard_normal(m)
    assert np.linalg.norm(apply(P, "QPPFApplyCP", pf, r, m) - np.linalg.solve(GGt, r)) <= 1e-11 * np.linalg.norm(r) * np.linalg.cond(GGt)
    assert np.linalg.norm(apply(P, "QPPFApplyHalfQTranspose", pf, r, n) - G.T @ np.linalg.solve(GGt, r)) <= tol
    # shell operators (QPPFCreateQ / P / GtG): MatMult == the Apply functions
    for fn, ref in (("QPPFCreateQ", Q @ v), ("QPPFCreateP", v - Q @ v), ("QPPFCreateGtG", G.T @ (G @ v))):
        M = C.c_void_p()
        P.call(fn, pf, C.byref(M))
        x, y = P.VecFromArray(v.copy()), P.VecFromArray(np.zeros(n))
        P.MatMult(M, x, y)
        assert np.linalg.norm(P.VecGetArray(y) - ref) <= tol, fn
        P.VecDestroy(x), P.VecDestroy(y), P.MatDestroy(M)
    # the coarse-problem matrix: available unless the explicit inverse replaced it (qppf.c:753)
    GG = C.c_void_p()
    P.call("QPPFGetGGt", pf, C.byref(GG))
    assert bool(GG.value) == (not explicit)
    flg = C.c_int()
    P.call("QPPFGetGHasOrthonormalRows", pf, C.byref(flg))
    assert flg.value == 0
    P.call("QPPFDestroy", C.byref(pf))
    P.MatDestroy(Gm)
    P.options_clear()


def test_explicit_inverse_equals_factor_solve(P):
    """-qppf_explicit changes how inv(G G^T) is applied (matrix product instead of two triangular solves), not the result beyond rounding"""
    outs = []
    for explicit in (False, True):
        G, Gm, pf, rng = make(P, n=300, m=5, seed=11)
        P.call("QPPFSetExplicitInv", pf, C.c_int(1 if explicit else 0))
        v = rng.standard_normal(300)
        outs.append(apply(P, "QPPFApplyQ", pf, v, 300))
        P.call("QPPFDestroy", C.byref(pf))
        P.MatDestroy(Gm)
    assert np.linalg.norm(outs[0] - outs[1]) <= 1e-11 * np.linalg.norm(outs[0])


def test_smalxe_lag_options_are_accepted_and_inert_when_BE_has_mult(P):
    """smalxe.c:878-886: the lagged / u'B'Bu norm updates are only selected when B_E has no MatMult; with an ordinary equality matrix the
    -qps_smalxe_norm_update_lag* switches are read (smalxe.c:754-762) and change nothing"""
    from permon_b200 import problems as PR
    N = 24
    res = []
    for opts in ("-qps_rtol 1e-9", "-qps_rtol 1e-9 -qps_smalxe_norm_update_lag -qps_smalxe_norm_update_lag_offset 3 -qps_smalxe_norm_update_lag_start 4 "
                 "-qps_smalxe_norm_update_lag_step 2 -qps_smalxe_norm_update_lag_end 9 -qps_smalxe_norm_update_lag_lower 0.2 -qps_smalxe_norm_update_lag_upper 1.5"):
        pr = PR.obstacle2d(N)
        n = N * N
        pr.B = np.full((1, n), 1.0 / np.sqrt(n))
        pr.c = np.array([-0.05 * np.sqrt(n)])
        res.append(P.solve_problem(pr, "smalxe", opts))
    assert res[0].reason == res[1].reason == 2
    assert res[0].its == res[1].its and res[0].stats["inner_iter_accu"] == res[1].stats["inner_iter_accu"]
    assert np.array_equal(res[0].x, res[1].x)
