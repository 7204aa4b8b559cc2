mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu --deselect tests/test_gpu_multi.py > gpurun_out/gpu_tests.log 2>&1; tail -3 gpurun_out/gpu_tests.log
timeout 900 python bench.py --steps 1000 --warmup 50 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/r1d_bench_c2.json
timeout 900 python bench.py --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/r1d_bench_c2_k2000.json
python - <<'PY'
import json
for f in ("r1d_bench_c2","r1d_bench_c2_k2000"):
    d=json.loads(open(f"gpurun_out/{f}.json").read())
    print(f, d["value"], d["e2e"]["value"], d["roofline"]["kernel"], d["roofline"]["frac"], d["roofline"]["traffic"], {k:(v["avg_ms"],v["frac_of_measured_peak"]) for k,v in d["roofline"]["per_kernel"].items()}, d["clocks"])
PY
