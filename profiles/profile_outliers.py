"""Device-side outlier removal on its own: wall time per call (host lists in, survivors out) for a few batch sizes.
usage: python profiles/profile_outliers.py [n_points=4100] [reps=20]"""
import os
import sys
import time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'opencl-structure-from-motion_b200'), os.path.join(ROOT, 'oracle')]
import numpy as np
import visocu_py as V

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4100
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
rng = np.random.default_rng(3)


def make(n):
    m = np.zeros(n, V.P_MATCH)
    cells = rng.choice(620 * 188, size=n, replace=False)
    m['u1c'] = (cells % 620) * 2 + 6; m['v1c'] = (cells // 620) * 2 + 6
    flow = rng.integers(-3, 4, (n, 2))
    m['u1p'] = m['u1c'] - flow[:, 0]; m['v1p'] = m['v1c'] - flow[:, 1]
    return m


ctx = V.Context(0)
ctx.configure(V.Params(), 1241, 376, 4)
for batch in (1, 4, 16, 64):
    lists = [make(n) for _ in range(batch)]
    ctx.remove_outliers(lists, 0)
    t0 = time.perf_counter()
    for _ in range(reps):
        out, status = ctx.remove_outliers(lists, 0)
    dt = (time.perf_counter() - t0) / reps
    print('n=%d batch=%d: %.3f ms per call, %.3f ms per list, status %s, survivors %d' % (
        n, batch, dt * 1e3, dt * 1e3 / batch, sorted(set(status.tolist())), len(out[0])))
