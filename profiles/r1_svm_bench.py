"""C4-style measurement: SVM dual  min 1/2 a'(Z Z')a - 1'a,  0 <= a <= C,  y'a = 0  through SMALXE + MPGP, Hessian applied
as two SpMVs (Z' then Z) with exactly 1000 non-zeros per row of Z -> the long-row `k_spmv_vector` kernels.  Scaled-down twin of
BASELINE.json's configs[3] (10^5 x 10^5 instead of 10^6 x 10^6: 10^8 non-zeros, 1.2 GB per factor) so that host generation stays
under a minute.  Usage: python profiles/r1_svm_bench.py [n]   -> one JSON line."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch
    from permon_b200 import api as P
    from permon_b200 import problems as PR
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
    P.initialize()
    t0 = time.time()
    pr = PR.svm_dual(n, n, nnz_per_row=1000)
    t_gen = time.time() - t0
    nnz = len(pr.a)
    # Hessian application alone: t = Z' p (n x d transposed factor), Ap = Z t
    A1 = P.MatCreateAIJ(pr.ia, pr.ja, pr.a, ncols_local=pr.meta["d"])
    A2 = P.MatCreateAIJ(pr.second[0], pr.second[1], pr.second[2], ncols_local=pr.n)
    A = P.MatCreateProd([A2, A1])
    x, y = P.VecFromArray(np.random.default_rng(0).standard_normal(n)), P.VecCreate(n)
    for _ in range(5):
        P.MatMult(A, x, y)
    P.call("PermonB200Synchronize")
    reps = 50
    t0 = time.perf_counter()
    for _ in range(reps):
        P.MatMult(A, x, y)
    P.call("PermonB200Synchronize")
    ms_mult = (time.perf_counter() - t0) * 1e3 / reps
    bytes_mult = 2 * (12 * nnz + 4 * (n + 1)) + 8 * n * 4      # two CSR streams, x, t (w + r), y
    info = [P.MatStorageInfo(A1), P.MatStorageInfo(A2)]
    for o in (x, y):
        P.VecDestroy(o)
    for o in (A, A1, A2):
        P.MatDestroy(o)
    # the whole SMALXE solve through the public API (upload + set-up + outer/inner iterations + download)
    torch.cuda.synchronize()
    t0 = time.time()
    # at this size the inner MPGP needs more than its default 10000 iterations per outer step (n = d: Z Z' is badly conditioned), so the
    # window is bounded: 2 outer iterations of at most 1500 inner iterations each; what is reported is throughput, not convergence
    r = P.solve_problem(pr, "smalxe", "-qps_rtol 1e-5 -qps_max_it 2 -smalxe_qps_max_it 1500")
    torch.cuda.synchronize()
    t_solve = time.time() - t0
    inner = r.stats["inner_iter_accu"]
    print(json.dumps(dict(workload=f"C4s SVM dual {n} x {n}, 1000 nnz/row ({nnz / 1e6:.0f}M nnz per factor), SMALXE + MPGP", generate_s=round(t_gen, 1),
                          hessian_apply_ms=round(ms_mult, 4), hessian_apply_gbs=round(bytes_mult / (ms_mult * 1e-3) / 1e9, 1),
                          hessian_apply_frac_of_measured_peak=round(bytes_mult / (ms_mult * 1e-3) / 1e9 / 6550.1, 4), storage=info,
                          smalxe=dict(reason=r.reason, outer_its=r.its, inner_its=inner, counts=r.counts, seconds_end_to_end=round(t_solve, 3),
                                      inner_its_per_s_end_to_end=round(inner / t_solve, 1), Bx_residual=float(abs(pr.B @ r.x).max()),
                                      active_lower=int((r.x <= 1e-12).sum()), active_upper=int((r.x >= 1.0 - 1e-12).sum())))))


if __name__ == "__main__":
    main()
