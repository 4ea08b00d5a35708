#!/bin/bash
cd "$(dirname "$0")/.."
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
pick='import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d["value"], d["e2e"]["value"], d["ms_per_step"], d["e2e"]["h2d_bytes_per_step"])'
B="python bench.py --no-extra --no-roofline --no-cpu-baseline --steps 80 --warmup 3"
for t in 6 8 12 16; do
  echo "== threads=$t"; timeout 300 $B --threads $t 2>/dev/null | python -c "$pick"
done
