"""Where the end-to-end time of a short solve goes (host buffers -> QPSSolve -> download): wall-clock phases through the
public API, for the K the driver uses (20).  With PERMON_B200_TIMING=1 the library adds its own phase lines on stderr.

Usage: python profiles/r2_e2e_breakdown.py [workload=c2] [K=20] [reps=3]  -> one JSON line per repetition."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import bench
    from permon_b200 import api as P
    wl = sys.argv[1] if len(sys.argv) > 1 else "c2"
    K = int(sys.argv[2]) if len(sys.argv) > 2 else 20
    reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
    P.initialize()
    t0 = time.perf_counter()
    pr = bench.generate(bench.workload_spec(wl), 0, 1)
    t_gen = time.perf_counter() - t0
    keys = ("ia", "ja", "a", "b", "lb") + (("ub",) if pr.ub is not None else ())
    t0 = time.perf_counter()
    host = {k: torch.from_numpy(np.ascontiguousarray(getattr(pr, k))).pin_memory() for k in keys}
    xh = torch.zeros(pr.n, dtype=torch.float64).pin_memory()
    print(json.dumps(dict(workload=wl, n=pr.n, nnz=pr.nnz, generate_s=round(t_gen, 2), pin_s=round(time.perf_counter() - t0, 2))), flush=True)
    sync = lambda: P.call("PermonB200Synchronize")
    for rep in range(reps):
        sys.stderr.write(f"---- repetition {rep}\n")
        xh.zero_()
        t = [time.perf_counter()]
        A = P.MatCreateAIJ(host["ia"].numpy(), host["ja"].numpy(), host["a"].numpy())
        sync(); t.append(time.perf_counter())
        vb, vl, vx = P.VecFromArray(host["b"].numpy()), P.VecFromArray(host["lb"].numpy()), P.VecFromArray(xh.numpy())
        vu = P.VecFromArray(host["ub"].numpy()) if "ub" in host else None
        qp = P.QPCreate()
        P.QPSetOperator(qp, A); P.QPSetRhs(qp, vb); P.QPSetInitialVector(qp, vx); P.QPSetBox(qp, None, vl, vu)
        qps = P.QPSCreate()
        P.QPSSetType(qps, "mpgp"); P.QPSSetQP(qps, qp); P.QPSSetAutoPostSolve(qps, False)
        P.QPSSetTolerances(qps, rtol=1e-30, atol=1e-300, maxits=K - 1)
        sync(); t.append(time.perf_counter())
        P.QPSSetUp(qps)
        sync(); t.append(time.perf_counter())
        P.QPSSolve(qps)
        sync(); t.append(time.perf_counter())
        P.VecSyncToHost(vx)
        sync(); t.append(time.perf_counter())
        P.QPSDestroy(qps); P.QPDestroy(qp)
        for v in (vb, vl, vx, vu):
            if v is not None:
                P.VecDestroy(v)
        P.MatDestroy(A)
        sync(); t.append(time.perf_counter())
        names = ["matrix: re-code + upload", "vectors + QP/QPS objects", "QPSSetUp (work vectors, power method)", f"QPSSolve (vector uploads + {K} iterations)",
                 "download x", "destroy (outside e2e)"]
        out = {n: round(1e3 * (b - a), 2) for n, a, b in zip(names, t[:-1], t[1:])}
        out["e2e_total_ms"] = round(1e3 * (t[-2] - t[0]), 2)
        out["rep"] = rep
        out["x_checksum"] = float(np.sum(xh.numpy()))
        print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
