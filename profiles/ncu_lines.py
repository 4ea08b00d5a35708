"""Aggregate an `ncu --page source --csv --print-source cuda,sass` dump per source line and per function range.
usage: python profiles/ncu_lines.py dump.csv source.cu [top=25]"""
import csv
import re
import sys

path, src = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
lines = open(src).read().splitlines()
# function ranges: a line that starts a __device__ / __global__ function opens a bucket until the next one
starts = []
for i, l in enumerate(lines, 1):
    m = re.match(r'^(?:template.*)?\s*(?:__device__|__global__|__host__ __device__).*?\b(\w+)\s*\(', l)
    if m and not l.startswith(' '):
        starts.append((i, m.group(1)))
    elif re.match(r'^k_\w+\(', l):
        starts.append((i, l.split('(')[0]))
def bucket(n):
    name = 'other'
    for s, nm in starts:
        if n >= s:
            name = nm
    return name
rows = list(csv.reader(open(path)))
cur_file = None
hdr = None
per_line = {}
for r in rows:
    if not r:
        continue
    if r[0] == 'File Path':
        cur_file = r[1]; continue
    if r[0] == 'Line No':
        hdr = r; continue
    if hdr is None or cur_file is None or not r[0].isdigit() or r[2] != '-':
        continue
    d = dict(zip(hdr[4:], r[4:]))
    key = (cur_file.split('/')[-1], int(r[0]), r[1][:90])
    inst = int(d['Instructions Executed'] or 0); th = int(d['Thread Instructions Executed'] or 0)
    smp = int(d['# Samples'] or 0)
    exc = int(d.get('L1 Wavefronts Shared Excessive') or 0)
    per_line[key] = (inst, th, smp, exc, {k: int(v or 0) for k, v in d.items() if k.startswith('stall_') and 'Not Issued' not in k})
tot_i = sum(v[0] for v in per_line.values()); tot_t = sum(v[1] for v in per_line.values()); tot_s = sum(v[2] for v in per_line.values())
print('total warp inst %d, thread inst %d (%.1f lanes), samples %d' % (tot_i, tot_t, tot_t / max(tot_i, 1), tot_s))
bk = {}
for (f, n, s), v in per_line.items():
    b = bucket(n) if f == src.split('/')[-1] else f
    a = bk.setdefault(b, [0, 0, 0, 0, {}])
    a[0] += v[0]; a[1] += v[1]; a[2] += v[2]; a[3] += v[3]
    for k, x in v[4].items():
        a[4][k] = a[4].get(k, 0) + x
print('\n%-28s %8s %6s %6s %8s %10s  top stalls' % ('function', 'inst%', 'lanes', 'smp%', 'instM', 'bankconfM'))
for b, a in sorted(bk.items(), key=lambda kv: -kv[1][2]):
    st = sorted(a[4].items(), key=lambda kv: -kv[1])[:4]
    print('%-28s %7.1f%% %6.1f %5.1f%% %8.2f %10.2f  %s' % (b, 100.0 * a[0] / tot_i, a[1] / max(a[0], 1), 100.0 * a[2] / max(tot_s, 1), a[0] / 1e6, a[3] / 1e6,
                                                ' '.join('%s=%d' % (k[6:], x) for k, x in st)))
print('\ntop lines by samples')
for (f, n, s), v in sorted(per_line.items(), key=lambda kv: -kv[1][2])[:top]:
    st = sorted(v[4].items(), key=lambda kv: -kv[1])[:3]
    print('%5.1f%% smp %5.1f%% inst %4.1f lanes  %s:%d  %s   [%s]' % (100.0 * v[2] / max(tot_s, 1), 100.0 * v[0] / tot_i, v[1] / max(v[0], 1), f[:12], n, s.strip()[:70],
                                                              ' '.join('%s=%d' % (k[6:], x) for k, x in st)))
