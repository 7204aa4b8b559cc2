# final build: the whole GPU suite, the driver's bench configuration, the e2e phase timers
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -25 > gpurun_out/r2v_gpu_tests.log; tail -3 gpurun_out/r2v_gpu_tests.log
PERMON_B200_TIMING=1 timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2v_bench_k20.json 2> gpurun_out/r2v_bench_k20.err
python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/r2v_bench_k20.json") if l.startswith("{")][-1])
print("C3", d["value"], "e2e", round(d["e2e"]["value"], 2), d["e2e"]["seconds"], "parity", d["parity"]["ok"], "traffic", d["roofline"]["traffic"], "frac", d["roofline"]["frac"],
      "| C2", d["c2"]["value"], "e2e", round(d["c2"]["e2e"]["value"], 1), d["c2"]["e2e"]["seconds"], d["c2"]["parity"]["ok"], "traffic", d["c2"]["roofline"]["traffic"])
PY
grep "timing\|permon_b200\]" gpurun_out/r2v_bench_k20.err | tail -30
