bash profiles/r2_ncu_all.sh r2m_c2 c2 6
bash profiles/r2_ncu_all.sh r2m_c3 c3 6
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2m_bench_k20.json 2> gpurun_out/r2m_bench_k20.err
python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/r2m_bench_k20.json") if l.startswith("{")][-1])
print(d["value"], d["e2e"]["value"], d["e2e"]["seconds"], d["parity"]["ok"], d["c2"]["value"], d["c2"]["e2e"]["value"], d["c2"]["e2e"]["seconds"], d["c2"]["parity"]["ok"])
PY
