mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q 2>&1 | tail -3
PERMON_B200_TIMELINE=gpurun_out/tl6_c3 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 300 --warmup 20 --no-e2e 2>&1 | tail -1 > gpurun_out/pk6_c3_2gpu.json
python - <<'PY'
import json
for f in ("pk6_c3_2gpu",):
    d=json.loads(open(f"gpurun_out/{f}.json").read())
    print(f, d["value"], d["ms_per_step"], d["roofline"]["achieved"], d["roofline"]["avg_launch_ms"], d["roofline"]["family_ms"])
PY
