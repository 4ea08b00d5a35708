#!/bin/bash
# Large lists (3840x2160): tree nodes built on the device vs all-host triangulation.  Parity tests, then the 4K bench both ways.
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests/test_gpu_outliers.py tests/test_gpu_fullsize.py tests/test_gpu_configs.py -q -x 2>&1 | tail -5
B="python bench.py --workload flow4k --sequences 16 --steps 6 --warmup 3 --no-extra --no-roofline --no-cpu-baseline"
pick='import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d["value"], d["e2e"]["value"], d["ms_per_step"])'
echo "== 4K bench, device nodes"; timeout 400 $B 2>gpurun_out/b4k_nodes.err | python -c "$pick"
echo "== 4K bench, all-host triangulation of the large lists"; VISOB_DEVICE_NODES=0 timeout 400 $B 2>gpurun_out/b4k_host.err | python -c "$pick"
tail -3 gpurun_out/b4k_nodes.err
