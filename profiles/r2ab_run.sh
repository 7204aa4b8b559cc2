bash profiles/r2_ab.sh r2ab_c2_onewave c2 1000
bash profiles/r2_ab.sh r2ab_c2_twowaves c2 1000 PERMON_B200_KB_TWO_WAVES=1
bash profiles/r2_ab.sh r2ab_c2x_onewave c2x 1000
bash profiles/r2_ab.sh r2ab_c2x_twowaves c2x 1000 PERMON_B200_KB_TWO_WAVES=1
bash profiles/r2_ab.sh r2ab_c3_onewave c3 300
bash profiles/r2_ab.sh r2ab_c3_twowaves c3 300 PERMON_B200_KB_TWO_WAVES=1
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "tutorials or ex3 or jbearing or smalxe or midsize" 2>&1 | tail -2
