free -g | head -2 > gpurun_out/r2o_host.txt; nproc >> gpurun_out/r2o_host.txt
timeout 1100 python bench.py --workload c4 --steps 100 --warmup 10 > gpurun_out/r2o_c4_1m.json 2> gpurun_out/r2o_c4_1m.err
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/r2o_c4_1m.json"))
    print("c4 1M", d["value"], "it/s", d["details"], d["roofline"]["avg_launch_ms"], d["roofline"]["frac"], d["e2e"], d["cpu_baseline"], d["parity"])
except Exception as e:
    print("FAILED", e); print(open("gpurun_out/r2o_c4_1m.err").read()[-2000:])
PY
