#!/bin/bash
# final check of the built libraries, then one ncu capture of the outlier kernel (one list of 4100 matches, 1024 threads,
# warp-level top merges); the capture command runs plainly first
cd "$(dirname "$0")/.."
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -2
CMD2="python profiles/profile_outliers.py 4100 2"
timeout 120 $CMD2 > gpurun_out/r2_plain_outliers.log 2>&1; echo "plain rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_remove_outliers -s 2 -c 1 -o gpurun_out/r2_prof_outliers $CMD2 > gpurun_out/r2_ncu_full_outliers.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out/r2_prof_outliers.ncu-rep
