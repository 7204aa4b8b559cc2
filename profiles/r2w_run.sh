timeout 900 python -m pytest tests/test_gpu_edge_cases.py tests/test_gpu_golden_text.py tests/test_c_consumer.py tests/test_adapter.py -q -m gpu 2>&1 | tail -3
PERMON_B200_TIMING=1 timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2w_bench_k20.json 2> gpurun_out/r2w_bench_k20.err
python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/r2w_bench_k20.json") if l.startswith("{")][-1])
print("C3", d["value"], "e2e", round(d["e2e"]["value"], 2), d["e2e"]["seconds"], "parity", d["parity"]["ok"],
      "| C2", d["c2"]["value"], "e2e", round(d["c2"]["e2e"]["value"], 1), d["c2"]["e2e"]["seconds"], d["c2"]["parity"]["ok"])
PY
grep "timing" gpurun_out/r2w_bench_k20.err | sed -n 16,26p
