"""Where the end-to-end time of a C2 solve goes (host buffers -> QPSSolve -> download): wall-clock phases through the public API.
Usage: python profiles/r1_e2e_breakdown.py [K]  -> one JSON line."""
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
    K = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
    P.initialize()
    pr = PR.obstacle2d(4096)
    host = {k: torch.from_numpy(np.ascontiguousarray(getattr(pr, k))).pin_memory() for k in ("ia", "ja", "a", "b", "lb")}
    xh = torch.zeros(pr.n, dtype=torch.float64).pin_memory()
    out = {}
    for rep in range(2):                       # second repetition: allocator and driver warmed up
        sync = lambda: P.call("PermonB200Synchronize")
        t = [time.perf_counter()]
        A = P.MatCreateAIJ(host["ia"].numpy(), host["ja"].numpy(), host["a"].numpy())
        sync(); t.append(time.perf_counter())
        vb, vl, vx = P.VecFromArray(host["b"].numpy()), P.VecFromArray(host["lb"].numpy()), P.VecFromArray(xh.numpy())
        qp = P.QPCreate()
        P.QPSetOperator(qp, A); P.QPSetRhs(qp, vb); P.QPSetInitialVector(qp, vx); P.QPSetBox(qp, None, vl, None)
        qps = P.QPSCreate()
        P.QPSSetType(qps, "mpgp"); P.QPSSetQP(qps, qp); P.QPSSetAutoPostSolve(qps, False)
        P.QPSSetTolerances(qps, rtol=1e-30, atol=1e-300, maxits=K - 1)
        sync(); t.append(time.perf_counter())
        P.QPSSetUp(qps)
        sync(); t.append(time.perf_counter())
        P.QPSSolve(qps)
        sync(); t.append(time.perf_counter())
        x = P.VecGetArray(vx)
        sync(); t.append(time.perf_counter())
        P.QPSDestroy(qps); P.QPDestroy(qp)
        for v in (vb, vl, vx):
            P.VecDestroy(v)
        P.MatDestroy(A)
        sync(); t.append(time.perf_counter())
        names = ["matrix: re-code + upload", "vectors + QP/QPS objects", "QPSSetUp (work vectors, vector uploads, power method)", f"QPSSolve ({K} iterations)",
                 "download x", "destroy"]
        out = {n: round(1e3 * (b - a), 2) for n, a, b in zip(names, t[:-1], t[1:])}
        out["total_ms"] = round(1e3 * (t[-2] - t[0]), 2)
        out["x_checksum"] = float(np.sum(x))
    print(json.dumps(out))


if __name__ == "__main__":
    main()
