"""Multi-GPU worker: launched once per GPU (torchrun or torch.multiprocessing) by tests/test_gpu_multi.py.
Each rank owns a contiguous row block of the problem, solves it through the C ABI on its GPU (NCCL halo + all-gathers),
rank 0 gathers the solution and checks it against the single-address-space CPU oracle."""
import json
import os
import sys

# torchrun exports OMP_NUM_THREADS=1; the host side of the set-up (split + re-coding of this rank's rows) is OpenMP-parallel: the ranks of
# the node share the host cores (must happen before the OpenMP runtime starts)
if os.environ.get("OMP_NUM_THREADS", "1") == "1" and not os.environ.get("PERMON_B200_KEEP_OMP"):
    os.environ["OMP_NUM_THREADS"] = str(max(1, (os.cpu_count() or 1) // int(os.environ.get("LOCAL_WORLD_SIZE", os.environ.get("WORLD_SIZE", "1")))))

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def build_problem(kind, rank, size):
    from permon_b200 import problems as PR
    if kind == "obstacle2d":
        N = 192
        starts = PR.row_partition(N * N, size, align=N)
        return PR.obstacle2d(N, -100.0), PR.obstacle2d(N, -100.0, rows=(starts[rank], starts[rank + 1])), starts, "-qps_rtol 1e-8 -qps_max_it 100000", dict(rtol=1e-8, max_it=100000)
    if kind == "obstacle3d":
        N = 40
        starts = PR.row_partition(N ** 3, size, align=N * N)
        return PR.obstacle3d(N), PR.obstacle3d(N, rows=(starts[rank], starts[rank + 1])), starts, "-qps_rtol 1e-8", dict(rtol=1e-8)
    if kind == "varcoef3d":
        N = 24
        starts = PR.row_partition(N ** 3, size, align=N * N)
        return PR.varcoef3d(N), PR.varcoef3d(N, rows=(starts[rank], starts[rank + 1])), starts, "-qps_rtol 1e-8 -qps_max_it 100000", dict(rtol=1e-8, max_it=100000)
    if kind == "varcoef3d64t":
        # C5-style two-bound case at 64^3 (262 144 dofs), truncated run: same step kinds for the first 300 iterations and the
        # BASELINE tolerances on x / objective at the cut
        N = 64
        starts = PR.row_partition(N ** 3, size, align=N * N)
        return PR.varcoef3d(N), PR.varcoef3d(N, rows=(starts[rank], starts[rank + 1])), starts, "-qps_rtol 1e-30 -qps_atol 1e-300 -qps_max_it 299", dict(rtol=1e-30, atol=1e-300, max_it=299)
    if kind == "smalxe":
        N = 64
        starts = PR.row_partition(N * N, size, align=N)
        full = PR.obstacle2d(N)
        loc = PR.obstacle2d(N, rows=(starts[rank], starts[rank + 1]))
        n = N * N
        full.B = np.full((1, n), 1.0 / np.sqrt(n))
        full.c = np.array([-0.05 * np.sqrt(n)])
        loc.B = full.B[:, starts[rank]:starts[rank + 1]].copy()
        loc.c = full.c if rank == 0 else np.zeros(0)   # rank 0 owns the single equality row (MatCreateOneRow layout)
        return full, loc, starts, "-qps_rtol 1e-9", dict(rtol=1e-9)
    if kind == "smalxe_aij2":
        # two equality rows assembled as a ROW-partitioned AIJ matrix with global columns (rank 0 owns row 0, rank 1 row 1): the
        # library re-distributes them by columns (shim.cpp: mat_eqrows_dense)
        import scipy.sparse as sp
        N = 64
        n = N * N
        starts = PR.row_partition(n, size, align=N)
        full = PR.obstacle2d(N)
        loc = PR.obstacle2d(N, rows=(starts[rank], starts[rank + 1]))
        full.B = np.vstack([np.full(n, 1.0 / np.sqrt(n)), np.sin(np.arange(n) * 0.01) / np.sqrt(n)])
        full.c = np.array([-0.05 * np.sqrt(n), 0.02])
        mine = [rank] if rank < 2 else []
        S = sp.csr_matrix(full.B[mine, :]) if mine else sp.csr_matrix((0, n))
        loc.BE_local = (S.indptr.astype(np.int32), S.indices.astype(np.int32), S.data.astype(np.float64))
        loc.B = None
        loc.c = full.c[mine].copy()
        return full, loc, starts, "-qps_rtol 1e-9", dict(rtol=1e-9)
    if kind == "projector":
        # SURVEY 8f rank 2 on several GPUs: QPTOrthonormalizeEq (Cholesky) + QPTEnforceEqByProjector, then SMALXE on (P A P, P b, box, T B):
        # same solution as plain SMALXE on the untransformed problem (checked against the oracle's solve of that)
        N = 48
        n = N * N
        starts = PR.row_partition(n, size, align=N)
        full = PR.obstacle2d(N)
        loc = PR.obstacle2d(N, rows=(starts[rank], starts[rank + 1]))
        scale = 1.0 + 40.0 * np.sin(np.arange(n) * 0.013) ** 2
        full.b = np.asarray(full.b) * scale
        loc.b = np.asarray(loc.b) * scale[starts[rank]:starts[rank + 1]]
        rng = np.random.default_rng(3)
        B = np.zeros((2, n))
        B[0] = 1.0
        B[1, n // 3:] = rng.random(n - n // 3)
        import scipy.sparse as sp
        full.B = B
        full.c = None
        mine = [rank] if rank < 2 else []           # row-partitioned AIJ equality matrix with global columns
        S = sp.csr_matrix(B[mine, :]) if mine else sp.csr_matrix((0, n))
        loc.BE_local = (S.indptr.astype(np.int32), S.indices.astype(np.int32), S.data.astype(np.float64))
        loc.B = None
        loc.c = None
        loc.transforms = ("orth_cholesky", "projector")
        return full, loc, starts, "-qps_type smalxe -qps_rtol 1e-8", dict(rtol=1e-8)
    raise ValueError(kind)


def main():
    import torch
    import torch.distributed as dist
    rank, size = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    kinds = sys.argv[1].split(",")
    from permon_b200 import api as P
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    P.call("PermonB200SetDevice", local)
    P.initialize()
    dist.init_process_group("nccl", device_id=dev)
    idt = torch.zeros(128, dtype=torch.uint8, device=dev)
    if rank == 0:
        idt.copy_(torch.frombuffer(bytearray(P.get_unique_id()), dtype=torch.uint8))
    dist.broadcast(idt, 0)
    P.comm_init_rank(size, rank, idt.cpu().numpy().tobytes())
    out = {}
    for kind in kinds:
        full, loc, starts, opts, okw = build_problem(kind, rank, size)
        qtype = "smalxe" if (kind.startswith("smalxe") or kind == "projector") else "mpgp"
        r = P.solve_problem(loc, qtype, opts)
        xs = [torch.zeros(starts[q + 1] - starts[q], dtype=torch.float64, device=dev) for q in range(size)]
        dist.all_gather(xs, torch.from_numpy(r.x).to(dev))
        x = torch.cat(xs).cpu().numpy()
        if rank == 0:
            from oracle import oracle_py as O
            op = O.Operator(full.ia, full.ja, full.a)
            bx = O.BoxC(full.n, full.lb, full.ub)
            if kind.startswith("smalxe") or kind == "projector":
                xr, ro = O.smalxe_solve(op, full.b, bx, full.B, full.c, full.x0, O.smalxe_opts(**okw))
                op.c.m = 0
                its_ref, its = ro["inner_its_accu"], r.stats["inner_iter_accu"]
            else:
                xr, ro = O.mpgp_solve(op, full.b, bx, full.x0, O.mpgp_opts(**okw))
                its_ref, its = ro["its"], r.its
                band = [its_ref]
                for t in ((2, 3, 5, 8) if kind != "varcoef3d64t" else ()):
                    _, rb = O.mpgp_solve(op, full.b, bx, full.x0, O.mpgp_opts(nthreads=t, **okw))
                    band.append(rb["its"])
                ro["band"] = band
            relx = float(np.linalg.norm(x - xr) / np.linalg.norm(xr))
            fo, fg = O.objective(op, full.b, xr), O.objective(op, full.b, x)
            out[kind] = dict(its=its, its_ref=its_ref, band=ro.get("band"), reason=r.reason, reason_ref=ro["reason"], relx=relx,
                             relf=float(abs(fg - fo) / abs(fo)), counts=r.counts,
                             counts_ref={k: ro[k] for k in ("ncg", "nexp", "nprop", "nmv")} if not (kind.startswith("smalxe") or kind == "projector") else None)
        dist.barrier()
    if rank == 0:
        print("MGPU_RESULT " + json.dumps(out), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
