#!/bin/bash
# ncu evidence for one workload on the current build (1 GPU):
#   1. launch list (gpu__time_duration.sum, --clock-control none) of a short bench run -> gpurun_out/<tag>_launches.csv
#   2. one `ncu --set full` capture of each fused kernel (K_A, K_B, K_C: the first working launches after the warm-up), exported to CSV on the
#      box (raw page; the .ncu-rep is too big to travel) -> gpurun_out/<tag>_full.raw.csv
# usage: profiles/r2_ncu_all.sh <tag> <workload> [steps=6] [ENV=VAL ...]
tag=$1; wl=$2; K=${3:-6}; shift 3
B="python bench.py --workload $wl --steps $K --warmup 3 --no-e2e --no-parity --no-cpu-baseline --no-c2"
env "$@" timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${tag}_launches.csv $B > gpurun_out/${tag}_launches.log 2>&1
# the iteration kernels only (demangled names: the power method's k_spmv_sd<EpiPower> does not match); skip K_A' / K_C of the initial phase
env "$@" timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"EpiAT<|EpiA2T<|k_update_B|k_direction_C" -s 2 -c 8 -o /tmp/$tag -f $B > gpurun_out/${tag}_full.log 2>&1
ncu -i /tmp/$tag.ncu-rep --page raw --csv > gpurun_out/${tag}_full.raw.csv 2>/dev/null
ls -la gpurun_out/${tag}_*
