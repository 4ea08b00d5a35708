"""Roofline leg of bench.py on its own (for ncu): 128 frames of 1241x376 per fused filter+NMS launch.
usage: python profiles/profile_filter.py [half_resolution=0] [launches=4] [width height]"""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'opencl-structure-from-motion_b200')]
import numpy as np
import synth
import visocu_py as V

half = int(sys.argv[1]) if len(sys.argv) > 1 else 0
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
W = int(sys.argv[3]) if len(sys.argv) > 3 else 1241
H = int(sys.argv[4]) if len(sys.argv) > 4 else 376
nb = 128 if W * H < 2000000 else 8
ctx = V.Context(0)
vp = V.Params(half_resolution=half)
if half:
    vp.match_radius //= 2
ctx.configure(vp, W, H, nb)
seq = synth.blob_sequence(8, W, H, seed=1234)
dev = ctx.device_alloc(seq.nbytes)
ctx.memcpy_h2d(dev, seq)
ptrs = [dev + (i % 8) * W * H for i in range(nb)]
frames = list(range(nb))
for _ in range(2):
    ctx.push_frames(frames, ptrs=ptrs, bpl_in=W, on_device=True)
ctx.profile(True)
for _ in range(reps):
    ns, nd = ctx.push_frames(frames, ptrs=ptrs, bpl_in=W, on_device=True)
ms, nl, nf = ctx.profile_read()
bpl = W + 15 - (W - 1) % 16
# algorithmic bytes per frame: 3 B per pixel of the padded image in full-resolution mode; in half-resolution
# mode 1 B read + 2 B (du_full, dv_full) + 0.5 B (half-resolution du, dv) = 3.5 B per full-resolution pixel (the half image
# is never written by the fused kernel; with VISOCU_FUSED_HALF=0 three kernels run and only the last one is timed here)
alg = (3.5 if half else 3.0) * bpl * H
print('frames/launch %d  launch %.4f ms  %.2f us/frame  algorithmic %.1f GB/s  (features per frame: %d sparse, %d dense)' % (
    nb, ms / nl, 1e3 * ms / nf, alg * nf / (ms * 1e-3) / 1e9, ns[0], nd[0]))
