#!/bin/bash
# final launch list of the flow pipeline (plain launches, see r2_pipeline_summary.md), after the parity tests and the plain run
cd "$(dirname "$0")/.."
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
CMD="python bench.py --steps 4 --warmup 3 --sequences 64 --threads 4 --no-cpu-baseline --no-extra --no-roofline"
VISOCU_GRAPHS=0 timeout 300 $CMD > gpurun_out/r2b_plain.log 2>&1; echo "plain rc=$?"; tail -c 300 gpurun_out/r2b_plain.log
VISOCU_GRAPHS=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 1100 -c 600 --csv --log-file gpurun_out/r2b_launches.csv $CMD > gpurun_out/r2b_ncu.log 2>&1; echo "ncu rc=$?"
