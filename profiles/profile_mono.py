"""Mono odometry with the CLI's bucketing (bucket.max_features = 1000, about 4 k matches into RANSAC) on its own, for
ncu: exercises k_hypotheses / k_score / k_finish (csrc/ransac.cu), k_triangulate / k_plane_sums (csrc/pose.cu) and the
feature kernels on 1241x376 corridor frames.   usage: python profiles/profile_mono.py [frames=4] [bucket_max=1000]"""
import os
import sys
import time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'opencl-structure-from-motion_b200')]
import synth
import visocu_py as V
import host_py as H

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4
bucket = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
seq = synth.corridor_sequence(n, seed=1234)
hp = H.MonoParams(match=V.Params(), f=synth.KITTI_F, cu=synth.KITTI_CU, cv=synth.KITTI_CV, height=1.6, pitch=-0.08,
                  bucket_max_features=bucket)
vo = H.Mono(hp)
for k in range(n):
    t0 = time.perf_counter()
    ok = vo.process(seq[k])
    print('frame %d ok=%d matches=%d inliers=%d  %.2f ms' % (k, ok, len(vo.matches()), len(vo.inliers()), 1e3 * (time.perf_counter() - t0)))
