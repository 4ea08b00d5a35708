#!/bin/bash
# ncu --set full of k_match<0> (flow instantiation, 64 registers) inside the flow pipeline with plain launches
cd "$(dirname "$0")/.."
CMD="python bench.py --steps 4 --warmup 3 --sequences 64 --threads 4 --no-cpu-baseline --no-extra --no-roofline"
VISOCU_GRAPHS=0 timeout 200 $CMD > gpurun_out/r2b_match_plain.log 2>&1; echo "plain rc=$?"
VISOCU_GRAPHS=0 timeout 300 ncu --set full --clock-control none --import-source on --kernel-name-base function -k regex:^k_match$ -s 40 -c 2 -o gpurun_out/r2b_prof_match $CMD > gpurun_out/r2b_match_ncu.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out/r2b_prof_match.ncu-rep
