"""CPU tests: pin the plain-C restatement (oracle/viso_oracle.c) against the unmodified reference compiled into
oracle/_ref (the reference holds no golden vectors for this path, so parity is pinned by executing
it).  Integer stages bit-exact, FP64 stages to the stated tolerance.  Needs no GPU."""
import numpy as np
import pytest

import synth
import pyref
import pyoracle as O


def valid(p, w, m=2):
    return p[m:-m, m:w - m]


@pytest.fixture(scope='module')
def img():
    return synth.blob_pair(400, 240, seed=17)


def test_simd_known_answers(ref):
    """The reference's own unit test (test/simd.cpp:100-131) pins the 16- and 32-byte SAD: same check here."""
    rng = np.random.default_rng(0)
    for _ in range(50):
        a = rng.integers(0, 256, 32, dtype=np.uint8); b = rng.integers(0, 256, 32, dtype=np.uint8)
        assert ref.sad32(a, b) == O.sad(a, b) == int(np.abs(a.astype(int) - b.astype(int)).sum())
        assert ref.sad16(a[:16], b[:16]) == O.sad(a[:16], b[:16])
    assert O.sad(np.zeros(32, np.uint8), np.full(32, 255, np.uint8)) == 8160


def test_filters(ref, img):
    I = O.pad(img[0])
    w = img[0].shape[1]
    du, dv = O.sobel5x5(I); rdu, rdv = ref.sobel5x5(I)
    assert np.array_equal(valid(du, w), valid(rdu, w)) and np.array_equal(valid(dv, w), valid(rdv, w))
    du, dv = O.sobel3x3(I); rdu, rdv = ref.sobel3x3(I)
    assert np.array_equal(valid(du, w, 1), valid(rdu, w, 1)) and np.array_equal(valid(dv, w, 1), valid(rdv, w, 1))
    assert np.array_equal(O.blob5x5(I)[3:-2, 3:w - 2], ref.blob5x5(I)[3:-2, 3:w - 2])
    assert np.array_equal(valid(O.checkerboard5x5(I), w)[:-1], valid(ref.checkerboard5x5(I), w)[:-1])


def test_half_image_and_nms(ref, img):
    rm = ref.matcher(pyref.MatcherParams())
    I = O.pad(img[0])
    w = img[0].shape[1]
    h1, d1 = O.half_image(I, w); h2, d2 = rm.half_image(I, w)
    assert d1 == d2 and np.array_equal(h1[:, :d1[0]], h2[:, :d2[0]])
    f1 = ref.blob5x5(I); f2 = ref.checkerboard5x5(I)
    for n in (2, 3, 9, 10):
        assert np.array_equal(O.nms(f1, f2, w, n, 50), rm.nms(f1, f2, w, n))
    assert [O.lib().vo_sparse_nms_n(k) for k in (1, 2, 3, 4, 8, 12)] == [3, 6, 9, 10, 10, 12]


@pytest.mark.parametrize('kw', [dict(), dict(half_resolution=0), dict(half_resolution=0, nms_n=2), dict(multi_stage=0)],
                         ids=['defaults', 'fullres', 'nms2', 'singlestage'])
def test_features_matching_refinement(ref, img, kw):
    """computeFeatures records, flow matching pass 1, priors, pass 2 and pixel refinement, all bit-exact."""
    rp = pyref.MatcherParams(**kw); op = O.Params(**kw)
    a, b = img
    rm = ref.matcher(rp)
    rm.push(a); rm.push(b)
    fa = O.compute_features(a, op); fb = O.compute_features(b, op)
    if op.multi_stage:
        assert np.array_equal(fa['rec1'], rm.maxima('1p1')) and np.array_equal(fb['rec1'], rm.maxima('1c1'))
    assert np.array_equal(fa['rec2'], rm.maxima('1p2')) and np.array_equal(fb['rec2'], rm.maxima('1c2'))
    eff = op.effective()
    dims = fa['dims']
    planes = dict(du1p=fa['du_full'] if op.half_resolution else fa['du'], dv1p=fa['dv_full'] if op.half_resolution else fa['dv'],
                  du1c=fb['du_full'] if op.half_resolution else fb['du'], dv1c=fb['dv_full'] if op.half_resolution else fb['dv'])
    if op.multi_stage:
        want1 = rm.matching(0, 0, False)
        got1 = O.matching(0, fa['rec1'], None, fb['rec1'], None, dims, eff)
        assert len(want1) > 20 and got1.tobytes() == want1.tobytes()
        kept = rm.remove_outliers(want1, 0)
        want_r = rm.prior(kept, 0).reshape(-1, 4, 4)[:, :, :2]
        got_r = O.prior_statistics(kept, 0, dims, eff)
        assert np.array_equal(got_r.reshape(-1, 4, 4)[:, :, :2], want_r)
        want2 = rm.matching(1, 0, True)
        got2 = O.matching(0, fa['rec2'], None, fb['rec2'], None, dims, eff, ranges=got_r)
    else:
        want2 = rm.matching(1, 0, False)
        got2 = O.matching(0, fa['rec2'], None, fb['rec2'], None, dims, eff)
    assert len(want2) > 100 and got2.tobytes() == want2.tobytes()
    want3 = rm.refinement(want2, 0)
    assert O.refine_pixel(want2, 0, dims, planes).tobytes() == want3.tobytes()


def test_quad_matching(ref):
    lp, rpv, lc, rc = synth.blob_quad(401, 240, seed=19)
    kw = dict(nms_n=2, half_resolution=0)
    rm = ref.matcher(pyref.MatcherParams(**kw)); op = O.Params(**kw)
    rm.push(lp, rpv); rm.push(lc, rc)
    f = [O.compute_features(i, op) for i in (lp, rpv, lc, rc)]
    want1 = rm.matching(0, 2, False)
    got1 = O.matching(2, f[0]['rec1'], f[1]['rec1'], f[2]['rec1'], f[3]['rec1'], f[0]['dims'], op.effective())
    assert len(want1) > 20 and got1.tobytes() == want1.tobytes()
    kept = rm.remove_outliers(want1, 2)
    ranges = O.prior_statistics(kept, 2, f[0]['dims'], op.effective())
    assert np.array_equal(ranges, rm.prior(kept, 2))
    want2 = rm.matching(1, 2, True)
    got2 = O.matching(2, f[0]['rec2'], f[1]['rec2'], f[2]['rec2'], f[3]['rec2'], f[0]['dims'], op.effective(), ranges=ranges)
    assert len(want2) > 100 and got2.tobytes() == want2.tobytes()
    planes = dict(du1p=f[0]['du'], dv1p=f[0]['dv'], du2p=f[1]['du'], dv2p=f[1]['dv'], du1c=f[2]['du'], dv1c=f[2]['dv'],
                  du2c=f[3]['du'], dv2c=f[3]['dv'])
    assert O.refine_pixel(want2, 2, f[0]['dims'], planes).tobytes() == rm.refinement(want2, 2).tobytes()


def _unit(F):
    F = F / np.linalg.norm(F)
    return F * np.sign(F.flat[np.argmax(np.abs(F))])


def test_ransac_pieces(ref_nofma):
    """FP64: F from the 8-point fit to 1e-9 (unit norm, fixed sign), inlier sets identical for the same F, RANSAC winner
    and inlier set identical, per-hypothesis counts equal for >= 99% of the hypotheses."""
    rng = np.random.default_rng(3)
    n = 300
    # synthetic two-view geometry: points in front of two cameras, 30% gross outliers, already Hartley-scaled
    X = np.stack([rng.uniform(-2, 2, n), rng.uniform(-1, 1, n), rng.uniform(4, 12, n)], 1)
    R = np.array([[0.9998, 0.0, 0.02], [0, 1, 0], [-0.02, 0, 0.9998]]); t = np.array([0.05, 0.01, -0.8])
    x1 = X[:, :2] / X[:, 2:]; X2 = X @ R.T + t; x2 = X2[:, :2] / X2[:, 2:]
    bad = rng.random(n) < 0.3
    x2[bad] += rng.normal(0, 0.2, (bad.sum(), 2))
    m = np.zeros(n, pyref.P_MATCH)
    m['u1p'], m['v1p'], m['u1c'], m['v1c'] = 500 * x1[:, 0] + 600, 500 * x1[:, 1] + 180, 500 * x2[:, 0] + 600, 500 * x2[:, 1] + 180
    vo = ref_nofma.mono(pyref.MonoParams())
    ok, mn, Tp, Tc = vo.normalize(m)
    ok2, mn2, Tp2, Tc2 = O.normalize(m)
    assert ok and ok2 and mn.tobytes() == mn2.tobytes() and np.allclose(Tp, Tp2, rtol=0, atol=1e-12) and np.allclose(Tc, Tc2, rtol=0, atol=1e-12)
    for _ in range(20):
        act = rng.choice(n, 8, replace=False).astype(np.int32)
        Fr = vo.fundamental(mn, act); Fo = O.fundamental(mn, act)
        assert np.abs(_unit(Fr) - _unit(Fo)).max() < 1e-9
        assert np.array_equal(vo.get_inlier(mn, Fr), O.get_inlier(mn, Fr))
    samples = np.stack([rng.choice(n, 8, replace=False) for _ in range(500)]).astype(np.int32)
    want = vo.ransac_with_samples(mn, samples); got = O.ransac(mn, samples)
    assert (want['counts'] != got['counts']).mean() <= 0.01
    assert want['best_iter'] == got['best_iter'] and np.array_equal(want['inliers'], got['inliers'])
    assert np.abs(_unit(want['F']) - _unit(got['F'])).max() < 1e-8
