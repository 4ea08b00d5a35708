"""GPU parity against the committed golden fixtures (outputs of the unmodified reference): the CUDA path, through the
C-ABI and through the C++ Matcher, must reproduce them without the reference being present."""
import os

import numpy as np
import pytest

import visocu_py as V
import host_py as H

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def load(name):
    z = np.load(os.path.join(GOLD, name + '.npz'), allow_pickle=True)
    kw = {k: int(v) for k, v in z['params']} if len(z['params']) else {}
    return z, kw


@pytest.mark.parametrize('name', ['flow_320x200_defaults', 'flow_322x160_fullres', 'flow_250x130_single_nms2'])
def test_flow_fixture(ctx, name):
    z, kw = load(name)
    vp = V.Params(**kw)
    if vp.half_resolution:
        vp.match_radius //= 2
    h, w = z['img_p'].shape
    ctx.configure(vp, w, h, 2)
    ctx.push_frames([0, 1], [z['img_p'], z['img_c']])
    if vp.multi_stage:
        assert np.array_equal(ctx.features(0, 0), z['rec_1p1']) and np.array_equal(ctx.features(1, 0), z['rec_1c1'])
    assert np.array_equal(ctx.features(0, 1), z['rec_1p2']) and np.array_equal(ctx.features(1, 1), z['rec_1c2'])
    quad = (0, -1, 1, -1)
    ranges = None
    if vp.multi_stage:
        assert ctx.match([quad], 0, 0)[0].tobytes() == z['raw1'].tobytes()
        ranges = [z['ranges']]
    assert ctx.match([quad], 0, 1, ranges=ranges)[0].tobytes() == z['raw2'].tobytes()
    assert ctx.match([quad], 0, 1, ranges=ranges, refine=True)[0].tobytes() == z['refined2'].tobytes()
    hm = H.Matcher(V.Params(**kw))
    hm.push(z['img_p']); hm.push(z['img_c']); hm.match_features(0)
    assert hm.matches(2).tobytes() == z['final'].tobytes()


def test_quad_fixture():
    z, kw = load('quad_322x160_nms2')
    hm = H.Matcher(V.Params(**kw))
    hm.push(z['img_1p'], z['img_2p']); hm.push(z['img_1c'], z['img_2c']); hm.match_features(2)
    assert hm.matches(2).tobytes() == z['final'].tobytes()


def _unit(F):
    F = F / np.linalg.norm(F)
    return F * np.sign(F.flat[np.argmax(np.abs(F))])


def test_ransac_fixture(ctx):
    z = np.load(os.path.join(GOLD, 'ransac_corridor_640x200.npz'))
    mn = z['normalized']
    uv = np.stack([mn['u1p'], mn['v1p'], mn['u1c'], mn['v1c']], axis=1)
    got = ctx.ransac([uv], [z['samples']], 1e-5, want_all=True)[0]
    assert (got['counts'] != z['counts']).mean() <= 0.01
    assert got['best_iter'] == int(z['best_iter']) and np.array_equal(got['inliers'], z['inliers'])
    assert np.abs(_unit(got['F']) - _unit(z['F'])).max() < 1e-8
