timeout 900 python -m pytest tests/test_gpu_transforms.py tests/test_gpu_qppf.py -x -q 2>&1 | tail -15
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "raw or packed or mixed" 2>&1 | tail -5
bash profiles/r2_ab.sh r2r_c2r c2r 1000
