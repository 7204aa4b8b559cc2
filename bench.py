#!/usr/bin/env python
"""bench.py -- MPGP iterations/second on the named configurations (BASELINE.json), one process per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c2|c2x|c3|c1]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

A "step" is ONE MPGP iteration (SpMV + fused update + direction update, src/qps/impls/mpgp/mpgp.c:511-641) of a
QPSSolve that runs through the C ABI of libpermon_b200.so.  W warm-up iterations, then exactly K timed ones:

  value  device-resident leg: CSR + vectors already in HBM, CUDA events on the library's stream, max over ranks
  e2e    same K iterations through the reference-facing calls with HOST buffers: MatCreate...WithArrays (H2D of the
         CSR), VecCreate...WithArray, QPSSetUp (power method), QPSSolve, VecGetArrayRead (D2H of x) all inside the
         timed region (wall clock around the calls, device synchronised on both sides)
  roofline      K_A (the fused SpMV) timed per launch with CUDA events inside a repeat of the timed region
  cpu_baseline  the CPU oracle (restatement of the reference's un-fused PETSc call sequence, OpenMP threads standing in
                for MPI ranks) on a bounded sample of the same workload, on this box's host cores

N = 1 runs C2 (2-D obstacle 4096^2, 16.7 M dofs); N > 1 runs C3 (3-D obstacle 512^3, 134 M dofs) row-partitioned in
z-slabs, strong scaling (the global problem is fixed).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "mpgp_iterations_per_second"
UNIT = "it/s"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def workload_spec(name, n_gpus):
    if name == "auto":
        name = "c2" if n_gpus == 1 else "c3"
    spec = {
        "c1": dict(kind="2d", N=256, bscale=-30.0, label="C1 2-D obstacle 256^2 (65 536 dofs), 5-point Laplacian, lower bound"),
        "c2": dict(kind="2d", N=4096, bscale=-30.0, label="C2 2-D obstacle 4096^2 (16.7M dofs), 5-point Laplacian, lower bound"),
        "c2x": dict(kind="2d", N=4096, bscale=-100.0, label="C2x expansion-heavy 2-D obstacle 4096^2, b=-100h^2"),
        "c3": dict(kind="3d", N=512, label="C3 3-D obstacle 512^3 (134M dofs), 7-point Laplacian, lower bound, z-slab row partition"),
        "c3s": dict(kind="3d", N=256, label="3-D obstacle 256^3 (16.7M dofs) stand-in"),
        "c5": dict(kind="var", N=256, label="C5 variable-coefficient 3-D Laplacian 256^3 (16.7M dofs, contrast 1e4), lb and ub arrays, ~50% active at the solution"),
        "c5s": dict(kind="var", N=128, label="variable-coefficient 3-D Laplacian 128^3 stand-in"),
    }[name]
    spec["name"] = name
    return spec


def generate(spec, rank, size):
    """this rank's row block, generated in ~2M-row chunks on a thread pool (numpy releases the GIL)"""
    from concurrent.futures import ThreadPoolExecutor
    from permon_b200 import problems as PR
    if spec["kind"] == "2d":
        N = spec["N"]
        starts = PR.row_partition(N * N, size, align=N)
        step = max(N, (2_000_000 // N) * N)
        make = lambda r0, r1: PR.obstacle2d(N, spec["bscale"], rows=(r0, r1))
    else:
        N = spec["N"]
        P = N * N
        starts = PR.row_partition(N ** 3, size, align=P)
        step = max(P, (2_000_000 // P) * P)
        if spec["kind"] == "var":
            make = lambda r0, r1: PR.varcoef3d(N, rows=(r0, r1))
        else:
            make = lambda r0, r1: PR.obstacle3d(N, rows=(r0, r1))
    rows = (starts[rank], starts[rank + 1])
    ranges = [(r0, min(r0 + step, rows[1])) for r0 in range(rows[0], rows[1], step)]
    workers = max(1, min(len(ranges), (os.cpu_count() or 1) // max(1, min(size, 8))))
    with ThreadPoolExecutor(workers) as ex:
        chunks = list(ex.map(lambda ab: make(*ab), ranges))
    ia = [np.zeros(1, np.int64)]
    off = 0
    for c in chunks:
        ia.append(c.ia[1:].astype(np.int64) + off)
        off += int(c.ia[-1])
    pr = chunks[0]
    pr.ia = np.concatenate(ia).astype(np.int32)
    pr.ja = np.concatenate([c.ja for c in chunks])
    pr.a = np.concatenate([c.a for c in chunks])
    pr.b = np.concatenate([c.b for c in chunks])
    pr.lb = np.concatenate([c.lb for c in chunks])
    pr.ub = np.concatenate([c.ub for c in chunks]) if chunks[0].ub is not None else None
    pr.x0 = np.zeros(rows[1] - rows[0])
    pr.r0, pr.r1 = rows
    return pr


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md clocks line).  The sampler is
    started before the warm-up so that it is already running when the (possibly very short) timed window opens; samples
    are selected by wall-clock time stamps, falling back to warm-up + timed when the window caught none."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None
        self.t0 = self.t1 = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def wait_first(self, timeout=8.0):
        """nvidia-smi needs a few hundred ms to start (longer on 8-GPU boxes): block until it delivers"""
        t = time.time()
        while self.proc and not self.rows and time.time() - t < timeout:
            time.sleep(0.01)

    def window_open(self):
        self.t0 = time.time()

    def window_close(self):
        self.t1 = time.time()

    def stop(self):
        if not self.proc:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        time.sleep(0.05)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()

        def parse(rows):
            sm, mx, reasons, pw = [], [], set(), []
            for _, r in rows:
                f = [t.strip() for t in r.split(",")]
                if len(f) < 7:
                    continue
                try:
                    sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
                except ValueError:
                    continue
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
            return sm, mx, reasons, pw

        inwin = [r for r in self.rows if self.t0 is not None and self.t1 is not None and self.t0 <= r[0] <= self.t1 + 0.03]
        window = "timed"
        if not inwin:
            inwin, window = self.rows, "set-up + warm-up + timed (timed window shorter than the sampling period)"
        sm, mx, reasons, pw = parse(inwin)
        if not sm:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["no samples"])
        return dict(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), power_w_max=float(max(pw)), samples=len(sm), window=window, reasons=sorted(reasons))


def algorithmic_bytes(n, nnz, counts, both_bounds=False, matrix_bytes=None):
    """B_cg = M + 8 n V ; B_exp = 2 M + 8 n V_e ; proportioning ~ B_cg, with M the matrix stream of one SpMV:
    12 nnz + 4(n+1) for CSR (the survey's formula), or the bytes of the packed tile format the library actually keeps in HBM."""
    V, Ve = (18, 19) if both_bounds else (16, 16)
    if matrix_bytes is not None:
        V -= 0.75          # layout actually streamed: K_B writes a byte mask instead of gf (-7/8 pass), K_C reads g + the mask (+1/8)
    M = 12 * nnz + 4 * (n + 1) if matrix_bytes is None else matrix_bytes
    b_cg = M + 8 * n * V
    b_exp = 2 * M + 8 * n * Ve
    return (counts["ncg"] + counts["nprop"]) * b_cg + counts["nexp"] * b_exp, b_cg, b_exp


# ----------------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle on the host cores
# ----------------------------------------------------------------------------------------------------------
def run_oracle(pr, warmup, steps, budget_s, maxeig=None):
    from oracle import oracle_py as O
    threads = os.cpu_count() or 1
    op = O.Operator(pr.ia, pr.ja, pr.a)
    bx = O.BoxC(pr.n, pr.lb, pr.ub)
    kw = dict(nthreads=threads)
    if maxeig:
        kw["maxeig"] = float(maxeig)
    x, r0 = O.mpgp_solve(op, pr.b, bx, pr.x0, O.mpgp_opts(max_it=max(warmup - 1, 0), **kw))
    t_it = r0["seconds"] / max(r0["its"], 1)
    n_t = int(max(5, min(steps, budget_s / max(t_it, 1e-9))))
    x2, r = O.mpgp_solve(op, pr.b, bx, x, O.mpgp_opts(max_it=n_t - 1, maxeig=r0["maxeig"], nthreads=threads))
    its = r["its"]
    return dict(value=its / r["seconds"], its=its, seconds=r["seconds"], threads=threads, counts={k: r[k] for k in ("ncg", "nexp", "nprop", "nmv")},
                sample=f"{its} MPGP iterations (after {r0['its']} warm-up iterations) of the full-size workload, {threads} OpenMP threads standing in for MPI ranks")


def short_device_leg(P, torch, dev, stream, pr, W, K):
    """device-resident MPGP window on one GPU, no profiling / e2e: used for the N = 1 point of the C3 strong-scaling series"""
    A = P.MatCreateAIJ(pr.ia, pr.ja, pr.a, ncols_local=pr.n)
    d = {k: torch.from_numpy(np.ascontiguousarray(getattr(pr, k))).to(dev) for k in ("b", "lb")}
    d["x"] = torch.zeros(pr.n, dtype=torch.float64, device=dev)
    vb, vlb, vx = (P.VecFromDevicePointer(d[k].data_ptr(), pr.n) for k in ("b", "lb", "x"))
    qp = P.QPCreate()
    P.QPSetOperator(qp, A); P.QPSetRhs(qp, vb); P.QPSetInitialVector(qp, vx); P.QPSetBox(qp, None, vlb, None)
    qps = P.QPSCreate()
    P.QPSSetType(qps, "mpgp"); P.QPSSetQP(qps, qp); P.QPSSetAutoPostSolve(qps, False)
    P.QPSSetTolerances(qps, rtol=1e-30, atol=1e-300, maxits=W - 1)
    P.QPSSetUp(qps)
    P.QPSSolve(qps)
    c0 = P.QPSMPGPGetStepCounts(qps)
    P.QPSSetTolerances(qps, maxits=K - 1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record(stream)
    P.QPSSolve(qps)
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    assert P.QPSGetIterationNumber(qps) == K
    c1 = P.QPSMPGPGetStepCounts(qps)
    P.QPSDestroy(qps); P.QPDestroy(qp)
    for v in (vb, vlb, vx):
        P.VecDestroy(v)
    P.MatDestroy(A)
    del d
    torch.cuda.empty_cache()
    return dict(value=round(K / (ms * 1e-3), 2), unit=UNIT, ms_per_step=round(ms / K, 5), steps=K, warmup=W, step_mix={k: c1[k] - c0[k] for k in c1})


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=50)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--workload", default="auto")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-scaling-base", action="store_true")
    ap.add_argument("--cpu-budget", type=float, default=20.0)
    args = ap.parse_args()
    K, W = max(1, args.steps), max(3, args.warmup)
    rank = int(os.environ.get("RANK", "0"))
    size = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    spec = workload_spec(args.workload, max(args.gpus, size))

    if args.impl == "reference":
        # the reference's own CPU implementation cannot be built here (PETSc/MPI absent): time the oracle port
        if rank != 0:
            return
        pr = generate(spec, 0, 1)
        res = run_oracle(pr, W, K, budget_s=60.0)
        line = dict(metric=METRIC, value=res["value"], unit=UNIT, n_gpus=args.gpus, steps=K, warmup=W, ms_per_step=1e3 / res["value"],
                    higher_is_better=True, scaling="strong", vs_baseline=None, dtype="f64", data="synthetic", impl="reference",
                    config=dict(workload=spec["label"], n=pr.N, nnz=pr.nnz, step_mix=res["counts"]),
                    cpu_baseline=dict(value=res["value"], unit=UNIT, cores=res["threads"], kind="port", sample=res["sample"]),
                    e2e=dict(value=res["value"], unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0), gpu_launches=0)
        print(json.dumps(line), flush=True)
        return

    import torch
    from permon_b200 import api as P

    if P.device_count() == 0:
        raise SystemExit("bench.py: no CUDA device -- the product has no CPU path")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    P.call("PermonB200SetDevice", local_rank)
    P.initialize()
    if size > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
        idt = torch.zeros(128, dtype=torch.uint8, device=dev)
        if rank == 0:
            idt.copy_(torch.frombuffer(bytearray(P.get_unique_id()), dtype=torch.uint8))
        dist.broadcast(idt, 0)
        P.comm_init_rank(size, rank, bytes(idt.cpu().numpy().tobytes()))
    stream = torch.cuda.current_stream()
    P.set_stream(stream.cuda_stream)

    def barrier():
        torch.cuda.synchronize()
        if size > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v):
        if size == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    t_gen = time.time()
    pr = generate(spec, rank, size)
    t_gen = time.time() - t_gen
    n_loc, nnz_loc = pr.n, pr.nnz
    # pinned host buffers (the e2e leg copies from these)
    vec_keys = ("b", "lb") + (("ub",) if pr.ub is not None else ())
    host = {k: torch.from_numpy(np.ascontiguousarray(getattr(pr, k))).pin_memory() for k in ("ia", "ja", "a") + vec_keys}
    both = pr.ub is not None

    def make_solver(device_resident, keep):
        """QP + QPS through the C ABI; returns handles"""
        h = {}
        if device_resident:
            # the matrix goes through the host constructor (the library re-codes it into its packed tile format and, when
            # row-partitioned, splits diagonal / off-diagonal blocks) and is resident before the timed region; vectors are
            # device arrays owned by the caller
            h["A"] = P.MatCreateAIJ(pr.ia, pr.ja, pr.a, ncols_local=n_loc)
            d = {k: host[k].to(dev) for k in vec_keys}
            d["x"] = torch.zeros(n_loc, dtype=torch.float64, device=dev)
            for k in vec_keys + ("x",):
                h[k] = P.VecFromDevicePointer(d[k].data_ptr(), n_loc)
            h["dev"] = d
        else:
            h["xh"] = torch.zeros(n_loc, dtype=torch.float64).pin_memory()
            h["A"] = P.MatCreateAIJ(host["ia"].numpy(), host["ja"].numpy(), host["a"].numpy(), ncols_local=n_loc)
            for k in vec_keys:
                h[k] = P.VecFromArray(host[k].numpy())
            h["x"] = P.VecFromArray(h["xh"].numpy())
        qp = P.QPCreate()
        P.QPSetOperator(qp, h["A"]); P.QPSetRhs(qp, h["b"]); P.QPSetInitialVector(qp, h["x"]); P.QPSetBox(qp, None, h["lb"], h.get("ub"))
        qps = P.QPSCreate()
        P.QPSSetType(qps, "mpgp")
        P.QPSSetQP(qps, qp)
        P.QPSSetAutoPostSolve(qps, False)
        h["qp"], h["qps"] = qp, qps
        keep.append(h)
        return h

    def destroy(h):
        P.QPSDestroy(h["qps"]); P.QPDestroy(h["qp"])
        for k in ("x",) + vec_keys:
            P.VecDestroy(h[k])
        P.MatDestroy(h["A"])

    keep = []
    sampler = ClockSampler(local_rank)
    sampler.start()
    # ---------------- device-resident leg -----------------------------------------------------------------
    h = make_solver(True, keep)
    P.QPSSetTolerances(h["qps"], rtol=1e-30, atol=1e-300, maxits=W - 1)     # never converge inside the window
    P.QPSSetUp(h["qps"])                                                      # upload done, power method done
    maxeig = P.QPSMPGPGetOperatorMaxEigenvalue(h["qps"])
    storage = P.MatStorageInfo(h["A"])
    sampler.wait_first()
    P.QPSSolve(h["qps"])                                                      # W warm-up iterations
    assert P.QPSGetIterationNumber(h["qps"]) == W, (P.QPSGetIterationNumber(h["qps"]), W)
    x_after_warmup = h["dev"]["x"].clone()
    c_warm = P.QPSMPGPGetStepCounts(h["qps"])
    P.QPSSetTolerances(h["qps"], maxits=K - 1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = P.launch_count()
    barrier()
    sampler.window_open()
    e0.record(stream)
    P.QPSSolve(h["qps"])                                                      # exactly K timed iterations
    e1.record(stream)
    barrier()
    sampler.window_close()
    clocks = sampler.stop()
    ms = max_over_ranks(e0.elapsed_time(e1))
    launches = P.launch_count() - launches0
    its = P.QPSGetIterationNumber(h["qps"])
    assert its == K, (its, K)
    c_all = P.QPSMPGPGetStepCounts(h["qps"])
    counts = {k: c_all[k] - c_warm[k] for k in c_all}
    value = K / (ms * 1e-3)

    # ---------------- roofline of the dominant kernel (K_A), measured in a repeat of the timed region ----------
    h["dev"]["x"].copy_(x_after_warmup)
    barrier()
    P.profile_begin()
    P.QPSSolve(h["qps"])
    prof = P.profile_end()
    if os.environ.get("PERMON_B200_TIMELINE"):
        P.call("PermonB200ProfileDump", (os.environ["PERMON_B200_TIMELINE"] + f".rank{rank}.csv").encode())
    barrier()
    peak, peak_src = peaks()
    nb = 1 + (1 if both else 0)                      # bound vectors
    M = storage["stream_bytes"]
    n_cg, n_ex = counts["ncg"] + counts["nprop"], counts["nexp"]
    # algorithmic bytes per launch of each fused kernel in the layout it streams, averaged over the
    # window's step mix where the kernel does different work per step kind
    kbytes = {
        "K_A spmv+dots+feas": M + 8 * n_loc * (4 + nb),                                                   # p g x bounds -> Ap
        "K_B update+split": 8 * n_loc * ((n_cg * (6.125 + nb) + n_ex * (5 + nb)) / max(n_cg + n_ex, 1)),  # c/p: 4+nb r, x g + byte mask w; e: 4+nb r, 1 w
        "K_A' spmv+grad+split": M + 8 * n_loc * (4 + nb),                                                 # x b bounds -> g p
        "K_C direction": 8 * n_loc * 3.125,                                                                # g, byte mask, p -> p
    }
    klaunch = {"K_A spmv+dots+feas": None, "K_B update+split": None, "K_A' spmv+grad+split": n_ex, "K_C direction": n_cg}
    per_kernel = {}
    for fam, byts in kbytes.items():
        pf = prof.get(fam)
        if not pf or not pf["launches"]:
            continue
        real = klaunch[fam] if klaunch[fam] is not None else pf["launches"]   # launches that did work (the rest exit at once)
        if fam == "K_A spmv+dots+feas" and size > 1:
            real = pf["launches"] / 2                                           # diagonal pass + ghost pass per SpMV
        if not real:
            continue
        avg = pf["total_ms"] / real
        gbs = byts / (avg * 1e-3) / 1e9
        per_kernel[fam] = dict(total_ms=round(pf["total_ms"], 3), working_launches=int(real), avg_ms=round(avg, 5), bytes_per_launch=int(byts),
                               achieved_gbs=round(gbs, 1), frac_of_measured_peak=round(gbs / peak, 4))
    fam_ms = {k: round(v["total_ms"], 3) for k, v in prof.items() if v["launches"]}
    total_prof_ms = sum(v["total_ms"] for v in prof.values())
    dom = max(per_kernel, key=lambda k: per_kernel[k]["total_ms"]) if per_kernel else None
    dk = per_kernel.get(dom, dict(achieved_gbs=0.0, avg_ms=0.0, bytes_per_launch=0, working_launches=0, total_ms=0.0))
    csr_M = 12 * nnz_loc + 4 * (n_loc + 1)
    # DRAM traffic of the dominant kernel: taken from the committed `ncu --set full` capture of the same workload (never measured
    # inside a bench run); null when no capture exists for this workload / rank count
    traffic, traffic_src = None, None
    try:
        tj = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "r1b_ncu_traffic.json")))
        if size == 1 and spec["label"].startswith(tj["workload"] + " ") and dom in tj["dram_bytes_per_launch"]:
            traffic, traffic_src = tj["dram_bytes_per_launch"][dom], tj["source"]
    except Exception:
        pass
    roofline = dict(bound="hbm", kernel=dom, achieved=dk["achieved_gbs"], peak=peak, unit="GB/s", frac=round(dk["achieved_gbs"] / peak, 4), traffic=traffic,
                    traffic_source=traffic_src,
                    peak_source=peak_src, launches=dk["working_launches"], avg_launch_ms=dk["avg_ms"], algorithmic_bytes_per_launch=dk["bytes_per_launch"],
                    kernel_share_of_step=round(dk["total_ms"] / total_prof_ms, 4) if total_prof_ms else None,
                    bytes_note="bytes of the layout the kernels stream (packed matrix tiles + fp64 vectors); with the CSR formula of SURVEY 8d K_A / K_A' "
                               "would count csr_matrix_bytes instead of matrix_bytes",
                    matrix_bytes=int(M), csr_matrix_bytes=int(csr_M), per_kernel=per_kernel, family_ms=fam_ms)
    step_bytes, b_cg, b_exp = algorithmic_bytes(n_loc, nnz_loc, counts, both_bounds=both, matrix_bytes=storage["stream_bytes"])
    step_bytes_csr, b_cg_csr, b_exp_csr = algorithmic_bytes(n_loc, nnz_loc, counts, both_bounds=both)
    whole_iter_gbs = step_bytes / (ms * 1e-3) / 1e9
    whole_iter_gbs_csr = step_bytes_csr / (ms * 1e-3) / 1e9
    destroy(h)
    keep.clear()
    del h, x_after_warmup
    torch.cuda.empty_cache()

    # ---------------- e2e leg: host buffers, every copy inside the timed region ------------------------------
    e2e = None
    if not args.no_e2e:
        keep2 = []
        host_in = sum(host[k].numel() * host[k].element_size() for k in host) + n_loc * 8      # CSR + b, bounds, x0 handed over in host memory
        d2h = n_loc * 8
        barrier()
        t0 = time.perf_counter()
        h2 = make_solver(False, keep2)
        P.QPSSetTolerances(h2["qps"], rtol=1e-30, atol=1e-300, maxits=K - 1)
        P.QPSSolve(h2["qps"])                                                 # set-up (upload, power method) + K iterations
        P.VecSyncToHost(h2["x"])                                              # D2H of the iterate into the caller's (pinned) x buffer
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        xres = h2["xh"].numpy()
        barrier()
        te = max_over_ranks(t1 - t0)
        assert P.QPSGetIterationNumber(h2["qps"]) == K
        # what crosses PCIe: the matrix as the library stores it (packed tiles) + the vectors; the CSR itself is read on the host only
        h2d = P.MatStorageInfo(h2["A"])["stream_bytes"] + sum(host[k].numel() * host[k].element_size() for k in vec_keys) + n_loc * 8
        e2e = dict(value=K / te, unit=UNIT, h2d_bytes_per_step=h2d / K, d2h_bytes_per_step=d2h / K, seconds=round(te, 4), host_input_bytes=host_in,
                   note="host CSR + vectors -> re-code + upload + power-method set-up + K iterations + one download per QPSSolve; bytes are the totals of the solve divided by K",
                   x_checksum=float(np.sum(xres)))
        destroy(h2)

    # ---------------- CPU baseline (rank 0, N = 1 only) -----------------------------------------------------
    cpu = None
    if rank == 0 and size == 1 and not args.no_cpu_baseline:
        res = run_oracle(pr, W, K, budget_s=args.cpu_budget, maxeig=maxeig)
        cpu = dict(value=round(res["value"], 3), unit=UNIT, cores=res["threads"], kind="port", sample=res["sample"],
                   note="restatement of the reference CPU path (un-fused PETSc call sequence), not PETSc itself; the reference cannot be built here")

    # ---------------- N = 1 point of the strong-scaling series --------------------------------------------------
    # BASELINE.json quotes the metric on C2 for one GPU and on C3 for 1/2/4/8 GPUs: the N = 1 headline is C2, the N > 1 lines are
    # C3, so the one-GPU C3 number that the scaling efficiency has to be computed against is measured here as well
    scaling_base = None
    if size == 1 and args.workload == "auto" and not args.no_scaling_base:
        spec3 = workload_spec("c3", 1)
        t3 = time.time()
        pr3 = generate(spec3, 0, 1)
        sb = short_device_leg(P, torch, dev, stream, pr3, 20, min(K, 300))
        sb.update(workload=spec3["label"], seconds_total=round(time.time() - t3, 1),
                  note="N = 1 point of the C3 strong-scaling series that `bench.py --gpus N` (N > 1) reports: efficiency(N) = value(N) / (N * scaling_base.value)")
        scaling_base = sb
        del pr3
    scaling_note = None
    if size > 1:
        scaling_note = ("strong scaling of C3 (fixed 134M-dof problem); its one-GPU point is the `scaling_base` object of the N = 1 line, "
                        "whose headline `value` is C2 as BASELINE.json asks")

    if rank == 0:
        line = dict(metric=METRIC, value=round(value, 2), unit=UNIT, n_gpus=size, steps=K, warmup=W, ms_per_step=round(ms / K, 5), higher_is_better=True,
                    scaling="strong", vs_baseline=None, dtype="f64", data="synthetic",
                    config=dict(workload=spec["label"], n=pr.N, n_local=n_loc, nnz_local=nnz_loc, step_mix=counts,
                                l2="inputs larger than L2 (every kernel streams >= 3 vectors of 8n bytes, n >= 8.4M per GPU, vs 126 MB L2)",
                                matrix_storage=storage, maxeig=maxeig, bytes_per_cg_step=b_cg, bytes_per_expansion_step=b_exp,
                                bytes_per_cg_step_csr_formula=b_cg_csr, bytes_per_expansion_step_csr_formula=b_exp_csr,
                                csr_equivalent_gbs_whole_iteration=round(whole_iter_gbs_csr * size, 1),
                                achieved_gbs_whole_iteration=round(whole_iter_gbs * size, 1),
                                frac_of_measured_hbm_whole_iteration=round(whole_iter_gbs / peak, 4), frac_of_8tbs_whole_iteration=round(whole_iter_gbs / 8000.0, 4),
                                generate_s=round(t_gen, 1)),
                    clocks=clocks, e2e=e2e, gpu_launches=int(launches), roofline=roofline, cpu_baseline=cpu)
        if scaling_base:
            line["scaling_base"] = scaling_base
        if scaling_note:
            line["scaling_note"] = scaling_note
        print(json.dumps(line), flush=True)
    if size > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
