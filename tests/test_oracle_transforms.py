"""CPU checks of the oracle's restatement of QPTOrthonormalizeEq / QPTHomogenizeEq / QPTEnforceEqByProjector.
No reference output exists for these transforms without QPTDualize + MUMPS, so they are pinned to their defining identities and
to the solution of the untransformed problem."""
import numpy as np
import pytest
import scipy.sparse as sp

from oracle import oracle_py as O
from permon_b200 import problems as PR


def setup(N=24, m=3, seed=0):
    pr = PR.obstacle2d(N)
    n = pr.n
    rng = np.random.default_rng(seed)
    B = np.vstack([np.ones(n), rng.random(n), np.sin(np.arange(n) * 0.1)])[:m]
    A = sp.csr_matrix((pr.a, pr.ja, pr.ia), shape=(n, n))
    return pr, n, A, B, rng


@pytest.mark.parametrize("kind", ["gs", "cholesky"])
def test_orth_rows_identities(kind):
    pr, n, A, B, rng = setup()
    c = np.array([0.5, -1.0, 0.2])
    TB, Tc, T = O.orth_rows(B, c, kind)
    assert np.abs(TB @ TB.T - np.eye(3)).max() <= 1e-13            # orthonormal rows
    assert np.abs(T @ B - TB).max() <= 1e-14 and np.abs(T @ c - Tc).max() <= 1e-15
    assert np.allclose(np.triu(T, 1), 0.0)                          # both variants are triangular (LQ factorisation)
    # same row space: the projector is unchanged
    v = rng.standard_normal(n)
    assert np.abs(O.apply_P(TB, v) - O.apply_P(B, v)).max() <= 1e-12


def test_projected_operator_and_homogenisation():
    pr, n, A, B, rng = setup()
    TB, Tc, _ = O.orth_rows(B, np.array([1.0, 2.0, -0.5]), "gs")
    Pm = np.eye(n) - TB.T @ TB
    v = rng.standard_normal(n)
    op = O.Operator(pr.ia, pr.ja, pr.a)
    op.set_projector(TB, 2)
    assert np.abs(op.apply(v) - Pm @ (A @ (Pm @ v))).max() <= 1e-12
    op.set_projector(TB, 1)
    assert np.abs(op.apply(v) - Pm @ (A @ v)).max() <= 1e-12
    op0 = O.Operator(pr.ia, pr.ja, pr.a)
    xt, bh, lbh, _ = O.homogenize(op0, pr.b, O.BoxC(n, pr.lb, None), TB, Tc)
    assert np.abs(TB @ xt - Tc).max() <= 1e-14 and np.abs(Pm @ xt).max() <= 1e-13     # xtilde is the minimum-norm solution of TB x = Tc
    assert np.abs(bh - (pr.b - A @ xt)).max() <= 1e-15 and np.abs(lbh - (pr.lb - xt)).max() == 0.0


def test_projected_chain_reaches_the_solution_of_the_original_problem():
    pr = PR.obstacle2d(32)
    n = pr.n
    rng = np.random.default_rng(3)
    B = np.zeros((2, n))
    B[0] = 1.0
    B[1, n // 3:] = rng.random(n - n // 3)
    b = np.asarray(pr.b) * (1.0 + 40.0 * np.sin(np.arange(n) * 0.013) ** 2)
    c = np.array([-0.6 * n, -0.3 * B[1].sum()])
    bx = O.BoxC(n, pr.lb, None)
    TB, Tc, _ = O.orth_rows(B, c, "cholesky")
    xt, bh, lbh, _ = O.homogenize(O.Operator(pr.ia, pr.ja, pr.a), b, bx, TB, Tc)
    opP = O.Operator(pr.ia, pr.ja, pr.a)
    opP.set_projector(TB, 2)
    x, r = O.smalxe_solve(opP, O.apply_P(TB, bh), O.BoxC(n, lbh, None), TB, None, np.zeros(n), O.smalxe_opts(rtol=1e-8))
    x = x + xt
    x2, r2 = O.smalxe_solve(O.Operator(pr.ia, pr.ja, pr.a), b, bx, B, c, np.zeros(n), O.smalxe_opts(rtol=1e-8))
    assert r["reason"] == r2["reason"] == 2
    assert np.linalg.norm(x - x2) <= 1e-6 * np.linalg.norm(x2)
    assert np.sum(x - pr.lb < 1e-12) == np.sum(x2 - pr.lb < 1e-12) > 10


def test_implicit_orthonormalisation_restatement_is_consistent():
    """QPTOrthonormalizeEq(MAT_ORTH_IMPLICIT) + SMALXE with the u'B'Bu norm update (smalxe.c:265-285) and its lagged variant (:289-370):
    penalising with Q = B'(BB')^-1 B instead of B'B changes the path, not the solution.  (Parity unpinned against the reference: its outputs
    for this chain need FETI / MUMPS; the restatement is checked against the untransformed SMALXE solve.)"""
    pr = PR.obstacle2d(24)
    n = pr.n
    rng = np.random.default_rng(3)
    B = np.zeros((3, n))
    B[0] = 1.0
    B[1, n // 3:] = rng.random(n - n // 3)
    B[2] = np.sin(np.arange(n) * 0.05)
    b = np.asarray(pr.b) * (1.0 + 40.0 * np.sin(np.arange(n) * 0.013) ** 2)
    for c in (None, np.array([-0.6 * n, -0.3 * B[1].sum(), 0.05 * n])):
        bx = O.BoxC(n, pr.lb, None)
        x0, r0 = O.smalxe_solve(O.Operator(pr.ia, pr.ja, pr.a), b, bx, B, c, np.zeros(n), O.smalxe_opts(rtol=1e-9))
        x1, r1 = O.smalxe_solve(O.Operator(pr.ia, pr.ja, pr.a), b, bx, B, c, np.zeros(n), O.smalxe_opts(rtol=1e-9, implicit_orth=1))
        x2, r2 = O.smalxe_solve(O.Operator(pr.ia, pr.ja, pr.a), b, bx, B, c, np.zeros(n),
                                O.smalxe_opts(rtol=1e-9, implicit_orth=1, lag_enabled=1, lag_offset=3, Jstart=4, Jstep=2, Jend=8))
        assert r0["reason"] == r1["reason"] == r2["reason"] == 2
        assert np.linalg.norm(x1 - x0) <= 1e-6 * np.linalg.norm(x0)
        assert np.linalg.norm(x2 - x0) <= 1e-6 * np.linalg.norm(x0)
        assert np.max(np.abs(B @ x1 - (0.0 if c is None else c))) <= 1e-5 * n
