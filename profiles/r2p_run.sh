bash profiles/r2_ab.sh r2p_c2r c2r 1000
bash profiles/r2_ab.sh r2p_c2r_tma c2r 1000 PERMON_B200_SPMV=tma
bash profiles/r2_ncu_all.sh r2p_c2 c2 6
bash profiles/r2_ncu_all.sh r2p_c3 c3 6
timeout 300 python -m pytest tests/test_gpu_qppf.py -x -q 2>&1 | tail -15
