"""Config 4 of BASELINE.json: flow matching on high-density 3840x2160 frames (Matcher only; the reference's odometry does
not survive this density).  Prints one JSON line: frame pairs/s of this library (sequence runner, frames
resident in HBM) next to the unmodified reference on the same host cores.
usage: python profiles/bench_4k.py [sequences=8] [steps=4]"""
import json
import os
import sys
import threading
import time
os.environ.setdefault('CUDA_DEVICE_MAX_CONNECTIONS', '32')
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'opencl-structure-from-motion_b200'), os.path.join(ROOT, 'oracle')]
import numpy as np
import synth
import host_py as H
import visocu_py as V

S = int(sys.argv[1]) if len(sys.argv) > 1 else 8
K = int(sys.argv[2]) if len(sys.argv) > 2 else 4
W, Hh = 3840, 2160
NF = 4
pool = [synth.blob_sequence(NF, W, Hh, seed=1234 + s, n_blobs=120000) for s in range(2)]
flat = np.ascontiguousarray(np.stack([np.stack(p) for p in pool]))
ctx = V.Context(0)
dev = ctx.device_alloc(flat.nbytes)
ctx.memcpy_h2d(dev, flat)
fb = W * Hh
ptr = lambda s, k: dev + ((s % 2) * NF + k % NF) * fb
threads = min(S, os.cpu_count() or 1)
runner = H.Runner(0, S, threads, 0, 0, H.MonoParams(match=V.Params()))
dims = np.array([W, Hh, W], np.int32)
for k in range(2):
    runner.step([ptr(s, k) for s in range(S)], dims, on_device=True)
t0 = time.perf_counter()
secs, nm, ok = runner.run([[ptr(s, 2 + k) for s in range(S)] for k in range(K)], dims, on_device=True)
ours = S * K / (time.perf_counter() - t0)
stats = runner.outlier_stats()
runner.close()

import pyref
ref = pyref.RefLib()
cores = os.cpu_count() or 1
out = {}


def worker(t):
    imgs = np.ascontiguousarray(np.stack([pool[t % 2][k % NF] for k in range(3)]))
    tot, per, n = pyref.time_matcher_sequence(ref, pyref.MatcherParams(), 0, imgs)
    out[t] = tot


th = [threading.Thread(target=worker, args=(t,)) for t in range(cores)]
t0 = time.perf_counter()
for t in th:
    t.start()
for t in th:
    t.join()
refrate = cores * 2 / (time.perf_counter() - t0)
print(json.dumps({'metric': 'matched frame-pairs/sec at 3840x2160 (flow, Matcher defaults, 120 k blobs)', 'value': round(ours, 2),
                  'unit': 'frame-pairs/s', 'sequences': S, 'host_threads': threads, 'matches_per_pair': float(nm[-1].mean()),
                  'outlier_removal': stats,
                  'cpu_baseline': {'value': round(refrate, 2), 'cores': cores, 'kind': 'reference',
                                   'sample': '%d threads x 2 frame pairs' % cores}}))
