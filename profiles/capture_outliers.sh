#!/bin/bash
# Run on the GPU box (under gpurun): each command first without ncu, then under it.  Outputs land in gpurun_out/.
set -e
CMD="python bench.py --steps 2 --warmup 3 --sequences 16 --threads 4 --no-cpu-baseline"
$CMD > gpurun_out/plain_r1b.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 150 -c 500 --csv --log-file gpurun_out/launches_r1b.csv $CMD > gpurun_out/ncu_list_r1b.log 2>&1 || true
CMD2="python profiles/profile_outliers.py 4100 2"
$CMD2 > gpurun_out/plain_outliers.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_remove_outliers -s 2 -c 1 -o gpurun_out/prof_outliers $CMD2 > gpurun_out/ncu_full_outliers.log 2>&1 || true
ls -la gpurun_out | tail -8
