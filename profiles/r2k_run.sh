timeout 600 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -3 > gpurun_out/r2k_tests.log; tail -2 gpurun_out/r2k_tests.log
for pf in 0 1 2 4; do bash profiles/r2_ab.sh r2k_c2_pf$pf c2 1000 PERMON_B200_SD_PF=$pf; done
bash profiles/r2_ab.sh r2k_c2_pf2_occ6 c2 1000 PERMON_B200_SD_PF=2 PERMON_B200_SD_OCC=6
bash profiles/r2_ab.sh r2k_c2_pf2_occ4 c2 1000 PERMON_B200_SD_PF=2 PERMON_B200_SD_OCC=4
bash profiles/r2_ab.sh r2k_c2x_pf2 c2x 1000 PERMON_B200_SD_PF=2
bash profiles/r2_ab.sh r2k_c5_pf2 c5 1000 PERMON_B200_SD_PF=2
for pf in 0 2 4; do bash profiles/r2_ab.sh r2k_c3_pf$pf c3 300 PERMON_B200_SD_PF=$pf; done
bash profiles/r2_ab.sh r2k_c3_pf2_occ6 c3 300 PERMON_B200_SD_PF=2 PERMON_B200_SD_OCC=6
timeout 300 python bench.py --workload c4 --c4-n 200000 --steps 300 --warmup 20 > gpurun_out/r2k_c4_200k.json 2> gpurun_out/r2k_c4_200k.err; python -c "
import json; d=json.load(open('gpurun_out/r2k_c4_200k.json')); print('c4', d['value'], d['roofline']['avg_launch_ms'], d['roofline']['frac'], d['e2e']['seconds'], d['parity']['ok'])"
PERMON_B200_TIMING=1 timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2k_bench_k20.json 2> gpurun_out/r2k_bench_k20.err; tail -c 600 gpurun_out/r2k_bench_k20.json | head -c 300; grep -c timing gpurun_out/r2k_bench_k20.err
