#!/bin/bash
# all 16 host cores, worker count sweep (64 sequences): value / e2e / ms per step
cd "$(dirname "$0")/.."
pick='import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d["value"], d["e2e"]["value"], d["ms_per_step"])'
B="python bench.py --no-extra --no-roofline --no-cpu-baseline --steps 80 --warmup 3"
for t in 4 6 8 10 12 16; do
  echo "== 16 cores, threads=$t"; timeout 300 $B --threads $t 2>/dev/null | python -c "$pick"
done
echo "== 16 cores, threads=8 again"; timeout 300 $B --threads 8 2>/dev/null | python -c "$pick"
