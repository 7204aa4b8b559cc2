python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_edge_cases.py -q -x -k "packed or spmv or ragged or unaligned or tutorials or obstacle3d" 2>&1 | tail -2
PERMON_B200_TIMING=1 timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2z_bench_k20.json 2> gpurun_out/r2z_bench_k20.err
python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/r2z_bench_k20.json") if l.startswith("{")][-1])
print("C3", d["value"], "e2e", round(d["e2e"]["value"], 2), d["e2e"]["seconds"], "parity", d["parity"]["ok"],
      "| C2", d["c2"]["value"], "e2e", round(d["c2"]["e2e"]["value"], 1), d["c2"]["e2e"]["seconds"], d["c2"]["parity"]["ok"])
PY
grep "MatCreateSeq\|stencil form" gpurun_out/r2z_bench_k20.err
