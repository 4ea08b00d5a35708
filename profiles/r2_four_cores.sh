#!/bin/bash
# one GPU fed by four host cores (the share of a rank on the 8-GPU box with 32 vCPUs): worker count and wait policy
cd "$(dirname "$0")/.."
pick='import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d["value"], d["e2e"]["value"], d["ms_per_step"])'
B="python bench.py --no-extra --no-roofline --no-cpu-baseline --steps 80 --warmup 3"
for t in 4 6 8 12; do
  echo "== 4 cores, threads=$t"; taskset -c 0-3 timeout 300 $B --threads $t 2>/dev/null | python -c "$pick"
done
echo "== 4 cores, threads=8, sleep 20us"; VISOCU_WAIT_SLEEP_US=20 taskset -c 0-3 timeout 300 $B --threads 8 2>/dev/null | python -c "$pick"
echo "== 4 cores, threads=8, depth 2"; VISOB_DEPTH=2 taskset -c 0-3 timeout 300 $B --threads 8 2>/dev/null | python -c "$pick"
echo "== 16 cores, threads=16"; timeout 300 $B 2>/dev/null | python -c "$pick"
