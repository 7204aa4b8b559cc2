#!/bin/bash
# one `ncu --set full` capture of the SpMV kernels inside a short bench run, exported to CSV on the box (the .ncu-rep itself is too big to travel)
# usage: profiles/r2_ncu_csv.sh <tag> <workload> <skip> <count> [ENV=VAL ...]
tag=$1; wl=$2; skip=$3; cnt=$4; shift 4
rep=/tmp/$tag
env "$@" timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_spmv -s $skip -c $cnt -o $rep -f \
   python bench.py --workload $wl --steps 6 --warmup 3 --no-e2e --no-parity --no-cpu-baseline > gpurun_out/$tag.log 2>&1
ncu -i $rep.ncu-rep --page raw --csv > gpurun_out/$tag.raw.csv 2>/dev/null
ncu -i $rep.ncu-rep --page source --csv --print-source sass > gpurun_out/$tag.source.csv 2>/dev/null
ls -la gpurun_out/$tag.*
