#!/bin/bash
# A/B runs of the device-resident leg (one JSON line each) -- usage: profiles/r2_ab.sh <tag> <workload> <K> [ENV=VAL ...]
tag=$1; wl=$2; K=$3; shift 3
out=gpurun_out/${tag}.json
env "$@" timeout 600 python bench.py --workload $wl --steps $K --warmup 50 --no-e2e --no-parity --no-cpu-baseline > $out 2> gpurun_out/${tag}.err
python - "$out" "$tag" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    pk = d["roofline"]["per_kernel"]
    print(sys.argv[2], d["value"], "it/s", d["details"]["step_mix"], {k.split()[0]: (v["avg_ms"], v["frac_of_measured_peak"]) for k, v in pk.items()}, d["details"]["frac_of_measured_hbm_whole_iteration"])
except Exception as e:
    print(sys.argv[2], "FAILED", e)
PY
