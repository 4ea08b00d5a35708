cd /root/repo
pick='import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d["value"], d["e2e"]["value"], d["ms_per_step"], d["config"].get("host_ms_per_call"), d["config"].get("device_outlier_removal"))'
for cfg in "16 16" "32 16" "32 32" "24 24"; do set -- $cfg
  echo "== 4K sequences=$1 threads=$2"; timeout 400 python bench.py --workload flow4k --sequences $1 --threads $2 --steps 6 --warmup 3 --no-extra --no-roofline --no-cpu-baseline 2>/dev/null | python -c "$pick"
done
