"""GPU parity of the drop-in C++ classes (libviso_b200.so) against the reference classes: final getMatches()
after outlier removal and bucketing (P7), and VisualOdometryMono::process end to end (P8, toleranced)."""
import numpy as np
import pytest

import synth
import pyref
import visocu_py as V
import host_py as H

pytestmark = pytest.mark.gpu


def pad_safe_width(w, nms_n, half):
    """True if no maximum can sit at u = w-7, where the reference's descriptor depends on uninitialised pad bytes
    (SURVEY.md 8c caveat 3): (w - 2n - 13) mod (n+1) != 0 for the sparse and the dense neighbourhood size."""
    wm = w // 2 if half else w
    ns = nms_n * 3
    if ns > 10:
        ns = max(nms_n, 10)
    return all((wm - 2 * n - 13) % (n + 1) != 0 for n in (ns, nms_n))


@pytest.mark.parametrize('kw', [dict(), dict(half_resolution=0), dict(multi_stage=0), dict(nms_n=4, half_resolution=0),
                                dict(refinement=0)],
                         ids=['defaults', 'fullres', 'singlestage', 'nms4', 'norefine'])
def test_matcher_flow_final_list(ref, kw):
    p = V.Params(**kw)
    width = 1241 if pad_safe_width(1241, p.nms_n, p.half_resolution) else 1242
    assert pad_safe_width(width, p.nms_n, p.half_resolution)
    a, b = synth.blob_pair(width, 376, seed=41)
    rm = ref.matcher(pyref.MatcherParams(**kw)); hm = H.Matcher(V.Params(**kw))
    for m in (rm, hm):
        m.push(a); m.push(b); m.match_features(0)
    assert rm.counts() == hm.counts()
    assert hm.matches(1).tobytes() == rm.matches(1).tobytes()
    want = rm.matches(2)
    assert len(want) > 1000
    assert hm.matches(2).tobytes() == want.tobytes()
    # prior statistics of the host layer against the reference's (P5), on the stages that are defined for flow
    if V.Params(**kw).multi_stage:
        r1 = rm.ranges().reshape(-1, 4, 4)[:, :, :2]
        r2 = hm.prior(hm.matches(1), 0).reshape(-1, 4, 4)[:, :, :2]
        assert np.array_equal(r1, r2)
    assert abs(hm.gain(np.arange(50)) - rm.gain(np.arange(50))) < 1e-6


def test_matcher_sequence_ring_buffer_and_replace(ref):
    """Three frames, the third pushed with replace=true, dims given with a caller stride larger than the width."""
    seq = synth.blob_sequence(3, 800, 300, seed=43)
    rm = ref.matcher(pyref.MatcherParams()); hm = H.Matcher(V.Params())
    for m in (rm, hm):
        m.push(seq[0]); m.push(seq[1]); m.match_features(0)
    assert hm.matches(2).tobytes() == rm.matches(2).tobytes()
    for m in (rm, hm):
        m.push(seq[2], replace=True); m.match_features(0)
    assert len(rm.matches(2)) > 300
    assert hm.matches(2).tobytes() == rm.matches(2).tobytes()
    for m in (rm, hm):
        m.push(seq[1]); m.match_features(0)
    assert hm.matches(2).tobytes() == rm.matches(2).tobytes()


def test_matcher_quad_final_list_and_bucketing(ref):
    assert pad_safe_width(1242, 2, 0)
    lp, rpv, lc, rc = synth.blob_quad(1242, 376, seed=45)
    kw = dict(nms_n=2, half_resolution=0)
    rm = ref.matcher(pyref.MatcherParams(**kw)); hm = H.Matcher(V.Params(**kw))
    for m in (rm, hm):
        m.push(lp, rpv); m.push(lc, rc); m.match_features(2)
    want = rm.matches(2)
    assert len(want) > 1000
    assert hm.matches(2).tobytes() == want.tobytes()
    # bucketing draws from the process-wide rand() stream: compare the kept SET per bucket size, not the order
    for m in (rm, hm):
        m.bucket(1000, 50.0, 50.0)
    a, b = rm.matches(2), hm.matches(2)
    assert len(a) == len(b)
    assert sorted(a.tolist()) == sorted(b.tolist())


def test_matcher_error_behaviour(ref, capfd):
    hm = H.Matcher(V.Params())
    hm.match_features(0)                      # nothing pushed: silently no matches (matcher.cpp:190-212)
    assert len(hm.matches(2)) == 0
    img = synth.blob_pair(200, 100, seed=1)[0]
    hm.push(img)
    hm.match_features(0)                      # only one frame: still silent
    assert len(hm.matches(2)) == 0
    H.lib().visob_matcher_push(hm.h, None, None, np.array([200, 100, 100], np.int32).ctypes.data_as(__import__('ctypes').c_void_p), 0)
    assert 'ERROR: Image dimension mismatch!' in capfd.readouterr().err


def test_filter_namespace(ref):
    img = synth.blob_pair(256, 128, seed=9)[0]
    du, dv = H.filter_call(0, img); rdu, rdv = ref.sobel5x5(img)
    assert np.array_equal(du[2:-2, 2:-2], rdu[2:-2, 2:-2]) and np.array_equal(dv[2:-2, 2:-2], rdv[2:-2, 2:-2])
    assert np.array_equal(H.filter_call(2, img)[3:-2, 3:-2], ref.blob5x5(img)[3:-2, 3:-2])


def _mono_params(lib, bucket_max):
    kw = dict(f=synth.KITTI_F, cu=synth.KITTI_CU, cv=synth.KITTI_CV, height=1.6, pitch=-0.08, bucket_max_features=bucket_max)
    return pyref.MonoParams(match=pyref.MatcherParams(), **kw), H.MonoParams(match=V.Params(), **kw)


def test_mono_odometry_sequence():
    """VisualOdometryMono::process over a short corridor drive.  Both sides draw the RANSAC samples from a fresh
    std::default_random_engine(71) and bucket with rand() after srand(0), so the sample tables coincide as long as
    the match lists do.  Tolerances: rotation entries 1e-6 absolute, translation 1e-6 relative."""
    ref = pyref.RefLib(fresh=True)                    # own copy: the sample generator is process-wide state (viso.cpp:88)
    seq = synth.corridor_sequence(4, seed=1234)
    rp, hp = _mono_params(ref, 2)
    # rand() is process-wide state: run the two implementations one after the other, each from its own srand(0)
    rv = ref.mono(rp)
    want = []
    for k in range(len(seq)):
        ok = rv.process(seq[k])
        want.append((ok, rv.matches(), rv.inliers(), rv.motion()))
    del rv
    hv = H.Mono(hp)
    for k in range(len(seq)):
        ok_h = hv.process(seq[k])
        ok_r, a, inl, Tr = want[k]
        assert ok_r == ok_h
        if k == 0:
            continue
        assert ok_r
        b = hv.matches()
        assert len(a) > 100
        assert a.tobytes() == b.tobytes()                       # same bucketed list, same order
        assert np.array_equal(inl, hv.inliers())
        Th = hv.motion()
        assert np.abs(Tr[:3, :3] - Th[:3, :3]).max() < 1e-6
        assert np.abs(Tr[:3, 3] - Th[:3, 3]).max() < 1e-6 * max(1.0, np.abs(Tr[:3, 3]).max())
        assert abs(Th[2, 3]) > 0.3                               # the drive moves 0.8 m forward per frame


def test_matcher_stereo_method1(ref):
    """Matcher method 1 (stereo, matcher.cpp:1045-1084): one stereo pair, left-right-left circle, positive disparity."""
    lp, rpv, lc, rc = synth.blob_quad(1244, 376, seed=47)
    for kw in (dict(half_resolution=0), dict()):
        assert pad_safe_width(1244, 3, kw.get('half_resolution', 1))
        rm = ref.matcher(pyref.MatcherParams(**kw)); hm = H.Matcher(V.Params(**kw))
        for m in (rm, hm):
            m.push(lc, rc); m.match_features(1)
        want = rm.matches(2)
        assert len(want) > 1000
        assert np.all(want['u1c'] >= want['u2c']) and np.all(want['i1p'] == -1)
        assert hm.matches(1).tobytes() == rm.matches(1).tobytes()
        assert hm.matches(2).tobytes() == want.tobytes()


def test_matcher_subpixel_refinement(ref):
    """refinement = 2 (parabolicFitting, matcher.cpp:1379-1454) with the flow demo's parameters
    (matlab/demo_matching_flow.m:12-21).  Integer stages identical; the sub-pixel coordinates come from a
    double-precision least-squares solve done differently here (constant pseudo-inverse instead of a Gauss-Jordan solve
    per match): tolerance 1e-4 pixel."""
    kw = dict(nms_n=4, refinement=2, half_resolution=0)
    assert pad_safe_width(1242, 4, 0)
    a, b = synth.blob_pair(1242, 376, seed=49)
    rm = ref.matcher(pyref.MatcherParams(**kw)); hm = H.Matcher(V.Params(**kw))
    for m in (rm, hm):
        m.push(a); m.push(b); m.match_features(0)
    want, got = rm.matches(2), hm.matches(2)
    assert len(want) > 1000 and len(got) == len(want)
    for f in ('i1p', 'i1c', 'u1c', 'v1c'):
        assert np.array_equal(got[f], want[f])
    assert np.abs(got['u1p'] - want['u1p']).max() < 1e-4 and np.abs(got['v1p'] - want['v1p']).max() < 1e-4
    frac = np.abs(want['u1p'] - np.round(want['u1p']))
    assert (frac > 1e-3).mean() > 0.5                                   # it really is sub-pixel
    # quad matching with sub-pixel refinement: three relocations per match
    lp, rpv, lc, rc = synth.blob_quad(1242, 376, seed=51)
    kw = dict(nms_n=2, refinement=2, half_resolution=0)
    rm = ref.matcher(pyref.MatcherParams(**kw)); hm = H.Matcher(V.Params(**kw))
    for m in (rm, hm):
        m.push(lp, rpv); m.push(lc, rc); m.match_features(2)
    want, got = rm.matches(2), hm.matches(2)
    assert len(want) > 500 and len(got) == len(want)
    for f in ('i1p', 'i2p', 'i1c', 'i2c', 'u1c', 'v1c'):
        assert np.array_equal(got[f], want[f])
    for f in ('u1p', 'v1p', 'u2p', 'v2p', 'u2c', 'v2c'):
        assert np.abs(got[f] - want[f]).max() < 1e-4


def test_sequence_runner_equals_single_matchers(ref):
    """Sharded runner (MatcherBatch per worker: one batched launch per GPU stage for all of a worker's sequences) gives,
    for every sequence and frame, exactly the list a stand-alone Matcher / the reference gives."""
    import ctypes as C
    S, T = 5, 3
    seqs = [synth.blob_sequence(T, 500, 260, seed=60 + s) for s in range(S)]
    mp = H.MonoParams(match=V.Params())
    for threads in (1, 2):
        runner = H.Runner(0, S, threads, 0, 0, mp)
        dims = np.array([500, 260, 500], np.int32)
        for k in range(T):
            imgs = [np.ascontiguousarray(seqs[s][k]) for s in range(S)]
            secs, nm, ok = runner.step([i.ctypes.data for i in imgs], dims)
            if k == 0:
                assert nm.tolist() == [0] * S
        got = [runner.matches(s) for s in range(S)]
        runner.close()
        for s in range(S):
            rm = ref.matcher(pyref.MatcherParams())
            for k in range(T):
                rm.push(seqs[s][k])
                if k:
                    rm.match_features(0)
            want = rm.matches(2)
            assert len(want) > 300 and got[s].tobytes() == want.tobytes(), (threads, s)


def test_mono_runner_equals_single_objects():
    """Mono odometry through the sharded runner (shared MatcherBatch + per-sequence odometry objects) equals stand-alone
    VisualOdometryMono objects: same matches, same inliers, same pose."""
    S, T = 3, 3
    seqs = [synth.corridor_sequence(T, seed=1234 + s) for s in range(S)]
    _, hp = _mono_params(None, 2)
    runner = H.Runner(0, S, 2, 1, 0, hp)
    dims = np.array([1241, 376, 1241], np.int32)
    poses = []
    for k in range(T):
        imgs = [np.ascontiguousarray(seqs[s][k]) for s in range(S)]
        secs, nm, ok = runner.step([i.ctypes.data for i in imgs], dims)
        assert k == 0 or ok.tolist() == [1] * S
    got_m = [runner.matches(s) for s in range(S)]
    got_T = [runner.motion(s) for s in range(S)]
    runner.close()
    for s in range(S):
        hv = H.Mono(hp)
        for k in range(T):
            hv.process(seqs[s][k])
        assert hv.matches().tobytes() == got_m[s].tobytes()
        assert np.abs(hv.motion() - got_T[s]).max() < 1e-9


def test_quad_matching_with_motion_prediction(ref_nofma):
    """Tr_delta-guided quad matching (matcher.cpp:1112-1138): double-precision cost SAD + 4 * distance to the predicted
    position.  The GPU uses explicit IEEE operations without FMA contraction, so the list is compared bit-exactly against
    the reference built with -ffp-contract=off."""
    assert pad_safe_width(1242, 2, 0)
    lp, rpv, lc, rc = synth.blob_quad(1242, 376, seed=53, shift=(3, 1), disparity=12)
    kw = dict(nms_n=2, half_resolution=0)
    Tr = np.eye(4); Tr[0, 3] = 0.02; Tr[2, 3] = -0.3; Tr[0, 2] = 0.01; Tr[2, 0] = -0.01
    rm = ref_nofma.matcher(pyref.MatcherParams(**kw)); hm = H.Matcher(V.Params(**kw))
    for m in (rm, hm):
        m.set_intrinsics(645.2, 635.9, 194.1, 0.54)
        m.push(lp, rpv); m.push(lc, rc); m.match_features(2, tr_delta=Tr)
    want, got = rm.matches(2), hm.matches(2)
    assert len(want) > 500
    assert got.tobytes() == want.tobytes()


def test_stereo_odometry_sequence():
    """VisualOdometryStereo::process over a short stereo corridor drive: quad matching with motion prediction from the
    second pair on, 3-point RANSAC + Gauss-Newton on the host.  Same sample stream (fresh engine(71) on both sides),
    same matches -> same inliers; pose within 1e-6 (rotation entries absolute, translation relative)."""
    ref_nofma = pyref.RefLib('nofma', fresh=True)     # own copy: the sample generator is process-wide state (viso.cpp:88)
    frames = [synth.corridor_stereo_frame(k, seed=1234) for k in range(4)]
    kw = dict(f=synth.KITTI_F, cu=synth.KITTI_CU, cv=synth.KITTI_CV, base=0.54)
    rp = pyref.StereoParams(match=pyref.MatcherParams(), **kw); hp = H.StereoParams(match=V.Params(), **kw)
    rv = ref_nofma.stereo(rp)
    want = []
    for l, r in frames:
        ok = rv.process(l, r)
        want.append((ok, rv.matches(), rv.inliers(), rv.motion()))
    del rv
    hv = H.Stereo(hp)
    for k, (l, r) in enumerate(frames):
        ok_h = hv.process(l, r)
        ok_r, a, inl, Tr = want[k]
        assert ok_h == ok_r
        if k == 0:
            continue
        assert ok_r and len(a) > 100
        assert hv.matches().tobytes() == a.tobytes()
        assert np.array_equal(hv.inliers(), inl)
        Th = hv.motion()
        assert np.abs(Tr[:3, :3] - Th[:3, :3]).max() < 1e-6
        assert np.abs(Tr[:3, 3] - Th[:3, 3]).max() < 1e-6 * max(1.0, np.abs(Tr[:3, 3]).max())
        assert abs(abs(Th[2, 3]) - 0.8) < 0.05                      # 0.8 m forward per frame, metric thanks to the baseline


def test_structure_from_motion_facade():
    """StructureFromMotion::update (sfm.hh:46-77): odometry, pose accumulation, replace-on-failure and track-based
    reconstruction.  The reference facade itself needs the OpenCL headers, so its few lines are restated here on top of
    the reference's own VisualOdometryMono and Reconstruction objects (a private copy of the library: the sample
    generator of the shared one has been advanced by the odometry test)."""
    ref = pyref.RefLib(fresh=True)
    seq = synth.corridor_sequence(7, seed=1234)
    rp, hp = _mono_params(ref, 2)
    rv = ref.mono(rp); rr = ref.reconstruction()
    rr.set_calibration(rp.f, rp.cu, rp.cv)
    pose, replace, first = np.eye(4), False, True
    for k in range(len(seq)):
        ok = rv.process(seq[k], replace)
        if first:
            first = False
        elif ok:
            pose = pose @ np.linalg.inv(rv.motion())
            rr.update(rv.matches(), rv.motion(), 0, 2, 30.0, 3.0)
            replace = False
        else:
            replace = True
    want_pts = rr.points()
    del rv
    sfm = H.Sfm(hp, 1241, 376)
    for k in range(len(seq)):
        sfm.update(seq[k])
    got_pts = sfm.points()
    assert len(want_pts) > 20
    assert len(got_pts) == len(want_pts)
    assert np.abs(got_pts - want_pts).max() <= 1e-3 * max(1.0, np.abs(want_pts).max())
    assert np.abs(sfm.pose() - pose).max() < 1e-5


def test_pipelined_runner_equals_stepwise(ref):
    """runner.run keeps several steps of every sequence in flight (MatcherBatch::stepSubmit / stepCollect: consecutive steps
    on different lanes of the context, a ring of four frames per sequence, steps replayed as CUDA graphs once they repeat):
    per step and sequence the same match counts as stepping synchronously, the same final lists as the reference, and for
    the odometry mode the same poses.  Twelve steps, so that every lane runs plain, captured and replayed."""
    S, T = 5, 12
    dims = np.array([500, 260, 500], np.int32)
    seqs = [synth.blob_sequence(T, 500, 260, seed=80 + s) for s in range(S)]
    imgs = [[np.ascontiguousarray(seqs[s][k]) for s in range(S)] for k in range(T)]
    ptrs = [[i.ctypes.data for i in row] for row in imgs]
    mp = H.MonoParams(match=V.Params())
    a = H.Runner(0, S, 2, 0, 0, mp)
    counts_step = np.array([a.step(ptrs[k], dims)[1] for k in range(T)])
    want = [a.matches(s) for s in range(S)]
    a.close()
    for s in range(S):
        rm = ref.matcher(pyref.MatcherParams())
        rm.push(seqs[s][T - 2]); rm.push(seqs[s][T - 1]); rm.match_features(0)
        assert want[s].tobytes() == rm.matches(2).tobytes()
    for depth in (1, 2, 3):
        H.set_pipeline_depth(depth)
        b = H.Runner(0, S, 2, 0, 0, mp)
        secs, nm, ok = b.run(ptrs, dims)
        got = [b.matches(s) for s in range(S)]
        b.close()
        assert np.array_equal(nm, counts_step) and nm[1:].min() > 300, depth
        for s in range(S):
            assert got[s].tobytes() == want[s].tobytes(), (depth, s)
    H.set_pipeline_depth(3)
    # the same frames in pinned host memory: the runner copies them asynchronously (visocu_push_frames, on_device = 2)
    import torch
    pinned = torch.empty((T, S, 260, 500), dtype=torch.uint8).pin_memory()
    pinned.numpy()[:] = np.stack([np.stack(row) for row in imgs])
    pptrs = [[pinned[k, s].data_ptr() for s in range(S)] for k in range(T)]
    b = H.Runner(0, S, 2, 0, 0, mp)
    secs, nm, ok = b.run(pptrs, dims)
    got = [b.matches(s) for s in range(S)]
    b.close()
    assert np.array_equal(nm, counts_step)
    for s in range(S):
        assert got[s].tobytes() == want[s].tobytes(), ('pinned', s)
    # the drop-in pushBack consumes a pinned image before it returns: the caller may overwrite its buffer at once
    hm = H.Matcher(V.Params())
    buf = torch.empty((260, 500), dtype=torch.uint8).pin_memory()
    for k in (T - 2, T - 1):
        buf.numpy()[:] = seqs[0][k]
        hm.push(buf.numpy())                                 # a view of the pinned buffer: same address
        buf.numpy()[:] = 0
    hm.match_features(0)
    assert hm.matches(2).tobytes() == want[0].tobytes()
    # odometry mode
    S, T = 3, 7
    seqs = [synth.corridor_sequence(T, seed=1234 + s) for s in range(S)]
    imgs = [[np.ascontiguousarray(seqs[s][k]) for s in range(S)] for k in range(T)]
    ptrs = [[i.ctypes.data for i in row] for row in imgs]
    dims = np.array([1241, 376, 1241], np.int32)
    _, hp = _mono_params(None, 2)
    a = H.Runner(0, S, 2, 1, 0, hp)
    oks = np.array([a.step(ptrs[k], dims)[2] for k in range(T)])
    want_T = [a.motion(s) for s in range(S)]; want_m = [a.matches(s) for s in range(S)]
    a.close()
    b = H.Runner(0, S, 2, 1, 0, hp)
    secs, nm, ok = b.run(ptrs, dims)
    assert np.array_equal(ok, oks) and ok[1:].min() == 1
    for s in range(S):
        assert b.matches(s).tobytes() == want_m[s].tobytes()
        assert np.abs(b.motion(s) - want_T[s]).max() < 1e-9
    b.close()


def test_device_prior_ranges_equal_reference(ref):
    """Multi-stage flow matching computes Matcher::computePriorStatistics (matcher.cpp:734-868) on the device between the
    two passes (k_prior_ranges): the ranges must be bit-identical to the reference's statistics of the same first-pass
    list, in half- and full-resolution mode."""
    a, b = synth.blob_pair(1244, 376, seed=91, shift=(4, -2))
    for kw in (dict(), dict(half_resolution=0)):
        assert pad_safe_width(1244, 3, kw.get('half_resolution', 1))
        rm = ref.matcher(pyref.MatcherParams(**kw)); hm = H.Matcher(V.Params(**kw))
        for m in (rm, hm):
            m.push(a); m.push(b); m.match_features(0)
        first = hm.matches(1)
        assert len(first) > 200 and first.tobytes() == rm.matches(1).tobytes()
        # flow matching uses stages 0 and 1 of the four a range holds (matcher.cpp:1020-1026); the others are never read
        want = rm.prior(first, 0).reshape(-1, 4, 4)[:, :, :2]
        got = hm.ranges().reshape(-1, 4, 4)[:, :, :2]
        assert got.shape == want.shape and got.tobytes() == want.tobytes()
