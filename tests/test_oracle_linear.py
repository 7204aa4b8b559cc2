"""CPU checks of the oracle's QPSKSP (CG) and QPSPCPG restatements.  The reference holds no golden output for either solver
(no test runs -qps_type ksp / pcpg), so these are pinned only to the mathematics: the solution of the linear / saddle-point
system, the CG iteration count against scipy's CG on the same system, and the behaviour at the iteration limit."""
import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as sl

from oracle import oracle_py as O
from permon_b200 import problems as PR


def system(N=30, seed=0):
    pr = PR.obstacle2d(N)
    A = sp.csr_matrix((pr.a, pr.ja, pr.ia), shape=(pr.n, pr.n))
    return pr, A, np.random.default_rng(seed).standard_normal(pr.n)


def test_cg_solves_the_system_in_as_many_steps_as_textbook_cg():
    pr, A, b = system()
    x, r = O.cg_solve(O.Operator(pr.ia, pr.ja, pr.a), b, None, O.lin_opts(rtol=1e-10))
    assert r["reason"] == 2 and np.linalg.norm(A @ x - b) <= 1.01e-10 * np.linalg.norm(b)
    its = []
    sl.cg(A, b, rtol=1e-10, atol=0.0, callback=lambda xk: its.append(1))
    assert abs(len(its) - r["its"]) <= 1


def test_cg_nonzero_initial_guess_and_iteration_limit():
    pr, A, b = system(seed=1)
    xs = sl.spsolve(A.tocsc(), b)
    x, r = O.cg_solve(O.Operator(pr.ia, pr.ja, pr.a), b, xs + 1e-3, O.lin_opts(rtol=1e-12))
    assert r["reason"] == 2 and np.linalg.norm(x - xs) <= 1e-9 * np.linalg.norm(xs)
    x, r = O.cg_solve(O.Operator(pr.ia, pr.ja, pr.a), b, None, O.lin_opts(rtol=1e-14, max_it=5))
    assert (r["its"], r["reason"]) == (5, -3)
    x, r = O.cg_solve(O.Operator(pr.ia, pr.ja, pr.a), np.zeros(pr.n), None, O.lin_opts())
    assert r["its"] == 0 and r["reason"] > 0


def test_pcpg_solves_the_saddle_point_system():
    pr, A, b = system(seed=2)
    n = pr.n
    G = np.vstack([np.ones(n), np.cos(np.arange(n) * 0.02)])
    for c in (None, np.array([0.3, -1.0])):
        x, r = O.pcpg_solve(O.Operator(pr.ia, pr.ja, pr.a), b, G, c, None, O.lin_opts(rtol=1e-11))
        K = sp.bmat([[A, sp.csr_matrix(G.T)], [sp.csr_matrix(G), None]]).tocsc()
        ref = sl.spsolve(K, np.concatenate([b, c if c is not None else np.zeros(2)]))[:n]
        assert r["reason"] == 2
        assert np.linalg.norm(x - ref) <= 1e-8 * np.linalg.norm(ref)
        assert np.max(np.abs(G @ x - (c if c is not None else 0.0))) <= 1e-10
