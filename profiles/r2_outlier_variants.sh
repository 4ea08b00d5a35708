#!/bin/bash
# Outlier kernel variants (threads per CTA x warp-level top merges): phase times of one list, parity tests, pipeline bench.
# usage (GPU box): bash profiles/r2_outlier_variants.sh > gpurun_out/ro_variants.log 2>&1
cd "$(dirname "$0")/.."
for n in 4100 900; do
  for T in 1024 512 256; do for W in 0 1; do
    echo "== n=$n threads=$T warp=$W"
    VISOCU_RO_THREADS=$T VISOCU_RO_WARP=$W VISOCU_RO_STATS=1 timeout 120 python profiles/profile_outliers.py $n 5 2>&1 | grep -E "batch=(1|64):|outliers\]" | tail -3
  done; done
done
echo "== parity tests, defaults"; timeout 600 python -m pytest tests/test_gpu_outliers.py -q 2>&1 | tail -3
for T in 256 512 1024; do
  echo "== parity tests, threads=$T"; VISOCU_RO_THREADS=$T timeout 600 python -m pytest tests/test_gpu_outliers.py -q 2>&1 | tail -3
done
B="python bench.py --no-extra --no-roofline --no-cpu-baseline --steps 60 --warmup 3"
pick='import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d["value"], d["e2e"]["value"], d["ms_per_step"])'
echo "== bench 1024/warp0"; VISOCU_RO_THREADS=1024 VISOCU_RO_WARP=0 timeout 300 $B 2>/dev/null | python -c "$pick"
echo "== bench 1024/warp1"; VISOCU_RO_THREADS=1024 timeout 300 $B 2>/dev/null | python -c "$pick"
echo "== bench default (512 batched)"; timeout 300 $B 2>/dev/null | python -c "$pick"
echo "== bench 256"; VISOCU_RO_THREADS=256 timeout 300 $B 2>/dev/null | python -c "$pick"
echo "== bench default again"; timeout 300 $B 2>/dev/null | python -c "$pick"
