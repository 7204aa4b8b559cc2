PERMON_B200_SD_PAIRS=1 timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "packed or obstacle or golden or ex1 or spmv" 2>&1 | tail -4
for occ in 3 4 5; do bash profiles/r2_ab.sh r2s_c2_pairs_occ$occ c2 1000 PERMON_B200_SD_PAIRS=1 PERMON_B200_SD2_OCC=$occ; done
bash profiles/r2_ab.sh r2s_c2_single c2 1000
for occ in 3 4 5; do bash profiles/r2_ab.sh r2s_c3_pairs_occ$occ c3 300 PERMON_B200_SD_PAIRS=1 PERMON_B200_SD2_OCC=$occ; done
bash profiles/r2_ab.sh r2s_c3_single c3 300
bash profiles/r2_ab.sh r2s_c2x_pairs c2x 1000 PERMON_B200_SD_PAIRS=1
