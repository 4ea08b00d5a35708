"""CPU tests: the plain-C restatement against the committed golden fixtures (generated from the unmodified reference by
tests/golden/make_golden.py).  Runs anywhere -- neither /root/reference nor oracle/_ref nor a GPU is needed."""
import os

import numpy as np
import pytest

import pyoracle as O

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def load(name):
    z = np.load(os.path.join(GOLD, name + '.npz'), allow_pickle=True)
    kw = {k: int(v) for k, v in z['params']} if len(z['params']) else {}
    return z, kw


@pytest.mark.parametrize('name', ['flow_320x200_defaults', 'flow_322x160_fullres', 'flow_250x130_single_nms2'])
def test_flow_fixture(name):
    z, kw = load(name)
    p = O.Params(**kw)
    fa = O.compute_features(z['img_p'], p); fb = O.compute_features(z['img_c'], p)
    if p.multi_stage:
        assert np.array_equal(fa['rec1'], z['rec_1p1']) and np.array_equal(fb['rec1'], z['rec_1c1'])
    assert np.array_equal(fa['rec2'], z['rec_1p2']) and np.array_equal(fb['rec2'], z['rec_1c2'])
    wm = int(z['dims_m'][0])
    assert np.array_equal(fb['du'][2:-2, 2:wm - 2], z['du_c'][2:-2, 2:wm - 2])
    assert np.array_equal(fb['dv'][2:-2, 2:wm - 2], z['dv_c'][2:-2, 2:wm - 2])
    eff = p.effective()
    ranges = None
    if p.multi_stage:
        raw1 = O.matching(0, fa['rec1'], None, fb['rec1'], None, fa['dims'], eff)
        assert raw1.tobytes() == z['raw1'].tobytes()
        ranges = O.prior_statistics(z['kept1'], 0, fa['dims'], eff)
        assert np.array_equal(ranges.reshape(-1, 4, 4)[:, :, :2], z['ranges'].reshape(-1, 4, 4)[:, :, :2])
    raw2 = O.matching(0, fa['rec2'], None, fb['rec2'], None, fa['dims'], eff, ranges=ranges)
    assert raw2.tobytes() == z['raw2'].tobytes()
    full = bool(p.half_resolution)
    planes = dict(du1p=fa['du_full' if full else 'du'], dv1p=fa['dv_full' if full else 'dv'],
                  du1c=fb['du_full' if full else 'du'], dv1c=fb['dv_full' if full else 'dv'])
    assert O.refine_pixel(raw2, 0, fa['dims'], planes).tobytes() == z['refined2'].tobytes()


def test_quad_fixture():
    z, kw = load('quad_322x160_nms2')
    p = O.Params(**kw)
    f = [O.compute_features(z[k], p) for k in ('img_1p', 'img_2p', 'img_1c', 'img_2c')]
    raw1 = O.matching(2, f[0]['rec1'], f[1]['rec1'], f[2]['rec1'], f[3]['rec1'], f[0]['dims'], p.effective())
    assert raw1.tobytes() == z['raw1'].tobytes()


def _unit(F):
    F = F / np.linalg.norm(F)
    return F * np.sign(F.flat[np.argmax(np.abs(F))])


def test_ransac_fixture():
    z = np.load(os.path.join(GOLD, 'ransac_corridor_640x200.npz'))
    ok, mn, Tp, Tc = O.normalize(z['matches'])
    assert ok and mn.tobytes() == z['normalized'].tobytes()
    got = O.ransac(z['normalized'], z['samples'])
    assert (got['counts'] != z['counts']).mean() <= 0.01
    assert got['best_iter'] == int(z['best_iter']) and np.array_equal(got['inliers'], z['inliers'])
    assert np.abs(_unit(got['F']) - _unit(z['F'])).max() < 1e-8          # tolerance: SURVEY.md 8d P8
