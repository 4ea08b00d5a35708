#!/bin/bash
# Run on the GPU box (under gpurun): plain run first, then the ncu launch list and one full capture of the
# fused filter+NMS kernel.  Outputs land in gpurun_out/.
set -e
CMD="python bench.py --steps 2 --warmup 3 --sequences 8 --threads 4 --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1 || true
ncu --set full --clock-control none --import-source on -k regex:k_filter_nms -s 3 -c 2 -o gpurun_out/prof_filter_nms $CMD > gpurun_out/ncu_full.log 2>&1 || true
ls -la gpurun_out
