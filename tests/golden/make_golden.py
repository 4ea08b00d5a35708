"""Generates tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref/libvisoref.so, built from /root/reference by
oracle/Makefile).  Run in the container that has /root/reference; the fixtures travel, the reference does not.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (os.path.join(ROOT, 'oracle'), os.path.join(ROOT, 'opencl-structure-from-motion_b200')):
    sys.path.insert(0, p)
import pyref      # noqa: E402
import synth      # noqa: E402


def flow_case(name, w, h, seed, **kw):
    ref = pyref.RefLib()
    a, b = synth.blob_pair(w, h, seed=seed, n_blobs=int(6000 * w * h / (1241 * 376.0)))
    rm = ref.matcher(pyref.MatcherParams(**kw))
    rm.push(a); rm.push(b)
    out = dict(img_p=a, img_c=b, params=np.array(sorted(kw.items()), dtype=object) if kw else np.zeros((0, 2), object))
    for tag in ('1p1', '1c1', '1p2', '1c2'):
        out['rec_' + tag] = rm.maxima(tag)
    p = pyref.MatcherParams(**kw)
    if p.multi_stage:
        raw1 = rm.matching(0, 0, False)
        out['raw1'] = raw1
        kept1 = rm.remove_outliers(raw1, 0)
        out['kept1'] = kept1
        out['ranges'] = rm.prior(kept1, 0)
        raw2 = rm.matching(1, 0, True)
    else:
        raw2 = rm.matching(1, 0, False)
    out['raw2'] = raw2
    out['refined2'] = rm.refinement(raw2, 0)
    rm.match_features(0)
    out['final'] = rm.matches(2)
    du, dv, dims = rm.sobel('1c')
    out['du_c'] = du; out['dv_c'] = dv; out['dims_m'] = np.array(dims)
    np.savez_compressed(os.path.join(HERE, name + '.npz'), **out)
    print(name, {k: (v.shape if hasattr(v, 'shape') else v) for k, v in out.items() if k.startswith(('rec', 'raw', 'final'))})


def quad_case(name, w, h, seed, **kw):
    ref = pyref.RefLib()
    imgs = synth.blob_quad(w, h, seed=seed, n_blobs=int(6000 * w * h / (1241 * 376.0)))
    rm = ref.matcher(pyref.MatcherParams(**kw))
    rm.push(imgs[0], imgs[1]); rm.push(imgs[2], imgs[3])
    out = dict(img_1p=imgs[0], img_2p=imgs[1], img_1c=imgs[2], img_2c=imgs[3],
               params=np.array(sorted(kw.items()), dtype=object))
    out['raw1'] = rm.matching(0, 2, False)
    rm.match_features(2)
    out['final'] = rm.matches(2)
    np.savez_compressed(os.path.join(HERE, name + '.npz'), **out)
    print(name, len(out['raw1']), len(out['final']))


def ransac_case(name):
    ref = pyref.RefLib('nofma')
    seq = synth.corridor_sequence(2, 640, 200, seed=77)
    mp = pyref.MonoParams(match=pyref.MatcherParams(), f=synth.KITTI_F * 640 / 1241.0, cu=synth.KITTI_CU * 640 / 1241.0,
                          cv=synth.KITTI_CV * 640 / 1241.0, height=1.6, pitch=-0.08, bucket_max_features=5)
    vo = ref.mono(mp)
    vo.process(seq[0]); ok = vo.process(seq[1])
    pm = vo.matches()
    okn, pmn, Tp, Tc = vo.normalize(pm)
    rng = np.random.default_rng(11)
    samples = np.stack([rng.choice(len(pmn), 8, replace=False) for _ in range(400)]).astype(np.int32)
    res = vo.ransac_with_samples(pmn, samples)
    np.savez_compressed(os.path.join(HERE, name + '.npz'), matches=pm, normalized=pmn, Tp=Tp, Tc=Tc, samples=samples,
                        counts=res['counts'], F=res['F'], F_all=res['F_all'], inliers=res['inliers'], best_iter=res['best_iter'])
    print(name, len(pm), res['n_inliers'], res['best_iter'], ok)


if __name__ == '__main__':
    flow_case('flow_320x200_defaults', 320, 200, 101)
    flow_case('flow_322x160_fullres', 322, 160, 102, half_resolution=0)
    flow_case('flow_250x130_single_nms2', 250, 130, 103, half_resolution=0, multi_stage=0, nms_n=2)
    quad_case('quad_322x160_nms2', 322, 160, 104, half_resolution=0, nms_n=2)
    ransac_case('ransac_corridor_640x200')
