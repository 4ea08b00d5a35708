"""SAD matcher on its own (for ncu): flow matching pass 1 (no prior, full +-radius window) on a dense synthetic pair.
usage: python profiles/profile_match.py [width height half_resolution reps]"""
import os
import sys
import time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'opencl-structure-from-motion_b200')]
import numpy as np
import synth
import visocu_py as V

W = int(sys.argv[1]) if len(sys.argv) > 1 else 3840
H = int(sys.argv[2]) if len(sys.argv) > 2 else 2160
half = int(sys.argv[3]) if len(sys.argv) > 3 else 0
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
ctx = V.Context(0)
vp = V.Params(half_resolution=half)
if half:
    vp.match_radius //= 2
ctx.configure(vp, W, H, 2)
a, b = synth.blob_pair(W, H, seed=404, n_blobs=int(6000 * W * H / (1241 * 376.0) * 0.5))
ns, nd = ctx.push_frames([0, 1], [a, b])
quad = [(0, -1, 1, -1)]
for p in (0, 1):
    ctx.match(quad, 0, p)
    ctx.match_stats()
    ctx.sync()
    t0 = time.perf_counter()
    for _ in range(reps):
        m = ctx.match(quad, 0, p)
    dt = (time.perf_counter() - t0) / reps
    cand, scanned = ctx.match_stats()
    print('pass %d: %d x %d features, %d matches, %.0f candidates (%.0f scanned) per call, %.3f ms per call incl. copies' % (
        p, (ns, nd)[p][1], (ns, nd)[p][0], len(m[0]), cand / reps, scanned / reps, dt * 1e3))
