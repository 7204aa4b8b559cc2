#!/bin/bash
# multi-GPU evidence on one box -- usage: profiles/r2_mgpu.sh <tag> <ngpus> [pytest -k filter | no] [c5]
# 1. tests/test_gpu_multi.py (row-partitioned MPGP / SMALXE against the oracle, peer-memory and NCCL paths)
# 2. bench.py at the driver's K = 20 / W = 5 (all legs: e2e, parity against the oracle on the full-size problem)
# 3. bench.py K = 300 / W = 50, device-resident leg only, with the per-launch timeline of the profiled repeat
tag=$1; n=$2; tests=${3:-yes}; c5=${4:-no}
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29655 "$@"; }
nvidia-smi --query-gpu=index,name,clocks.max.sm --format=csv,noheader > gpurun_out/${tag}_gpus.txt
if [ "$tests" != no ]; then
  kf=(); [ "$tests" != yes ] && kf=(-k "$tests")
  timeout 900 python -m pytest tests/test_gpu_multi.py -q -x -s "${kf[@]}" 2>&1 | grep -v "^$" | tail -40 > gpurun_out/${tag}_tests.log
  tail -3 gpurun_out/${tag}_tests.log
fi
timeout 420 bash -c "$(declare -f run); n=$n; run bench.py --gpus $n --steps 20 --warmup 5" > gpurun_out/${tag}_bench_k20.json 2> gpurun_out/${tag}_bench_k20.err
PERMON_B200_TIMELINE=gpurun_out/${tag}_tl timeout 420 bash -c "$(declare -f run); n=$n; run bench.py --gpus $n --steps 300 --warmup 50 --no-e2e --no-parity --no-cpu-baseline" > gpurun_out/${tag}_bench_k300.json 2> gpurun_out/${tag}_bench_k300.err
if [ "$c5" = c5 ]; then
  timeout 420 bash -c "$(declare -f run); n=$n; run bench.py --gpus $n --workload c5 --steps 300 --warmup 50 --no-e2e --cpu-budget 10" > gpurun_out/${tag}_bench_c5_k300.json 2> gpurun_out/${tag}_bench_c5_k300.err
fi
ls gpurun_out/${tag}_tl*rank[1-9]*.csv 2>/dev/null | xargs -r rm -f      # rank 0's timeline is enough
python - $tag <<'PY'
import json, sys
import os
for k in ("k20", "k300", "c5_k300"):
    if not os.path.exists(f"gpurun_out/{sys.argv[1]}_bench_{k}.json"):
        continue
    try:
        d = json.loads([l for l in open(f"gpurun_out/{sys.argv[1]}_bench_{k}.json") if l.startswith("{")][-1])
        e = d.get("e2e") or {}
        p = d.get("parity") or {}
        print(k, d["n_gpus"], "gpus", d["value"], "it/s", d["details"]["step_mix"], "e2e", e.get("value"), e.get("seconds"), "parity", p.get("ok"), p.get("relx"), p.get("step_mix_equal"),
              {kk.split()[0]: vv["avg_ms"] for kk, vv in d["roofline"]["per_kernel"].items()}, d["roofline"]["family_ms"])
    except Exception as ex:
        print(k, "FAILED", ex)
        print(open(f"gpurun_out/{sys.argv[1]}_bench_{k}.err").read()[-1500:])
PY
