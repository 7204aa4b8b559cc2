"""QPSKSP (unconstrained CG) on the C2 Hessian (2-D Laplacian 4096^2, 16.7 M dofs): iterations / s of the fused form
(SpMV + p.Ap in one kernel, x / r update + r.r in one kernel) against the one-kernel-per-KSPCG-call form.  usage: python profiles/r2_cg_bench.py"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import bench
    from permon_b200 import api as P
    P.initialize()
    pr = bench.generate(bench.workload_spec("c2"), 0, 1)
    b = np.random.default_rng(1).standard_normal(pr.n)
    out = {}
    for driver in ("auto", "generic"):
        P.options_clear()
        K = 300
        P.call("PetscOptionsInsertString", None, f"-qps_rtol 1e-30 -qps_atol 1e-300 -qps_max_it {K} -qps_ksp_b200_driver {driver}".encode())
        A = P.MatCreateAIJ(pr.ia, pr.ja, pr.a)
        vb, vx = P.VecFromArray(b.copy()), P.VecFromArray(np.zeros(pr.n))
        qp = P.QPCreate()
        P.QPSetOperator(qp, A), P.QPSetRhs(qp, vb), P.QPSetInitialVector(qp, vx)
        qps = P.QPSCreate()
        P.QPSSetType(qps, "ksp"); P.QPSSetQP(qps, qp); P.QPSSetFromOptions(qps)
        P.QPSSetUp(qps)
        P.QPSSolve(qps)                      # warm-up (uploads, pools)
        P.VecSetArray(vx, np.zeros(pr.n))
        P.call("PermonB200Synchronize")
        t0 = time.perf_counter()
        P.QPSSolve(qps)
        P.call("PermonB200Synchronize")
        t = time.perf_counter() - t0
        its = P.QPSGetIterationNumber(qps)
        out[driver] = dict(its=its, seconds=round(t, 4), it_per_s=round(its / t, 1), rnorm=P.QPSGetResidualNorm(qps))
        P.QPSDestroy(qps); P.QPDestroy(qp); P.VecDestroy(vb); P.VecDestroy(vx); P.MatDestroy(A)
    out["speedup"] = round(out["auto"]["it_per_s"] / out["generic"]["it_per_s"], 3)
    print(json.dumps(dict(workload="CG on the C2 Hessian (16.7M dofs)", **out)))


if __name__ == "__main__":
    main()
