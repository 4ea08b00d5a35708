"""TEST INFRASTRUCTURE ONLY: ctypes bindings for oracle/_ref/libvisoref*.so (the unmodified reference
CPU path compiled by oracle/Makefile) and oracle/_build/libvisooracle.so (this repo's plain-C
restatement, viso_oracle.c).  May be imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs only -- never by the product package.
"""
import ctypes as C
import os
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))

P_MATCH = np.dtype([('u1p', 'f4'), ('v1p', 'f4'), ('i1p', 'i4'), ('u2p', 'f4'), ('v2p', 'f4'), ('i2p', 'i4'),
                    ('u1c', 'f4'), ('v1c', 'f4'), ('i1c', 'i4'), ('u2c', 'f4'), ('v2c', 'f4'), ('i2c', 'i4')])
assert P_MATCH.itemsize == 48


class MatcherParams(C.Structure):
    """Same field order and defaults as Matcher::parameters (reference matcher.h:42-69)."""
    _fields_ = [(n, C.c_int32) for n in ('nms_n', 'nms_tau', 'match_binsize', 'match_radius', 'match_disp_tolerance',
                                         'outlier_disp_tolerance', 'outlier_flow_tolerance', 'multi_stage',
                                         'half_resolution', 'refinement')] + \
               [(n, C.c_double) for n in ('f', 'cu', 'cv', 'base')]

    def __init__(self, **kw):
        super().__init__()
        d = dict(nms_n=3, nms_tau=50, match_binsize=50, match_radius=200, match_disp_tolerance=2,
                 outlier_disp_tolerance=5, outlier_flow_tolerance=5, multi_stage=1, half_resolution=1,
                 refinement=1, f=1.0, cu=0.0, cv=0.0, base=1.0)
        d.update(kw)
        for k, v in d.items():
            setattr(self, k, v)


class MonoParams(C.Structure):
    """VisualOdometryMono::parameters flattened (viso.h:33-62, viso_mono.h:32-45)."""
    _fields_ = [('match', MatcherParams), ('bucket_max_features', C.c_int32), ('bucket_width', C.c_double),
                ('bucket_height', C.c_double), ('f', C.c_double), ('cu', C.c_double), ('cv', C.c_double),
                ('height', C.c_double), ('pitch', C.c_double), ('ransac_iters', C.c_int32),
                ('inlier_threshold', C.c_double), ('motion_threshold', C.c_double)]

    def __init__(self, match=None, **kw):
        super().__init__()
        self.match = match if match is not None else MatcherParams()
        d = dict(bucket_max_features=2, bucket_width=50.0, bucket_height=50.0, f=1.0, cu=0.0, cv=0.0,
                 height=1.0, pitch=0.0, ransac_iters=2000, inlier_threshold=0.00001, motion_threshold=100.0)
        d.update(kw)
        for k, v in d.items():
            setattr(self, k, v)


class StereoParams(C.Structure):
    """VisualOdometryStereo::parameters flattened (viso.h:33-62, viso_stereo.h:32-46)."""
    _fields_ = [('match', MatcherParams), ('bucket_max_features', C.c_int32), ('bucket_width', C.c_double),
                ('bucket_height', C.c_double), ('f', C.c_double), ('cu', C.c_double), ('cv', C.c_double),
                ('base', C.c_double), ('ransac_iters', C.c_int32), ('inlier_threshold', C.c_double), ('reweighting', C.c_int32)]

    def __init__(self, match=None, **kw):
        super().__init__()
        self.match = match if match is not None else MatcherParams()
        d = dict(bucket_max_features=2, bucket_width=50.0, bucket_height=50.0, f=1.0, cu=0.0, cv=0.0,
                 base=1.0, ransac_iters=200, inlier_threshold=2.0, reweighting=1)
        d.update(kw)
        for k, v in d.items():
            setattr(self, k, v)


def _p(a, t=None):
    if a is None:
        return None
    return a.ctypes.data_as(C.c_void_p)


def padded(img):
    """Copy an (h,w) uint8 image into the 16-byte-multiple stride the reference uses internally
    (matcher.cpp:158-175), pad columns zero."""
    h, w = img.shape
    bpl = w + 15 - (w - 1) % 16
    out = np.zeros((h, bpl), dtype=np.uint8)
    out[:, :w] = img
    return out


class RefLib:
    """The reference itself.  variant: '' (g++ default FMA contraction) or 'nofma' (-ffp-contract=off)."""

    def __init__(self, variant='', fresh=False):
        name = 'libvisoref%s.so' % ('_' + variant if variant else '')
        path = os.path.join(HERE, '_ref', name)
        if not os.path.exists(path):
            raise FileNotFoundError(path + ' (run `make -C oracle ref` where /root/reference exists)')
        if fresh:
            # The reference keeps process-wide state (the function-static sample generator of viso.cpp:88): a test
            # that needs it in its initial state loads a private copy of the library, which has its own statics.
            import shutil
            import tempfile
            self._tmp = tempfile.NamedTemporaryFile(suffix='.so', delete=False)
            self._tmp.close()
            shutil.copyfile(path, self._tmp.name)
            path = self._tmp.name
        self.lib = L = C.CDLL(path)
        L.ref_build_info.restype = C.c_char_p
        L.ref_matcher_create.restype = C.c_void_p
        L.ref_mono_create.restype = C.c_void_p
        L.ref_stereo_create.restype = C.c_void_p
        L.ref_recon_create.restype = C.c_void_p
        L.ref_mono_matcher.restype = C.c_void_p
        L.ref_matcher_gain.restype = C.c_float
        L.ref_time_matcher_sequence.restype = C.c_double
        L.ref_time_mono_sequence.restype = C.c_double
        L.ref_find_best_plane.restype = C.c_double
        L.ref_time_parallel.restype = C.c_double
        self.info = L.ref_build_info().decode()

    # ---- filters (w must be a multiple of 16; returns planes of the same shape)
    def sobel5x5(self, img):
        h, w = img.shape
        slack = 64
        du = np.zeros(h * w + slack, np.uint8); dv = np.zeros(h * w + slack, np.uint8)
        src = np.zeros(h * w + slack, np.uint8); src[:h * w] = img.ravel()
        self.lib.ref_sobel5x5(_p(src), _p(du), _p(dv), w, h)
        return du[:h * w].reshape(h, w), dv[:h * w].reshape(h, w)

    def sobel3x3(self, img):
        h, w = img.shape
        slack = 64
        du = np.zeros(h * w + slack, np.uint8); dv = np.zeros(h * w + slack, np.uint8)
        src = np.zeros(h * w + slack, np.uint8); src[:h * w] = img.ravel()
        self.lib.ref_sobel3x3(_p(src), _p(du), _p(dv), w, h)
        return du[:h * w].reshape(h, w), dv[:h * w].reshape(h, w)

    def blob5x5(self, img):
        h, w = img.shape
        out = np.zeros(h * w + 64, np.int16)
        src = np.zeros(h * w + 64, np.uint8); src[:h * w] = img.ravel()
        self.lib.ref_blob5x5(_p(src), _p(out), w, h)
        return out[:h * w].reshape(h, w)

    def checkerboard5x5(self, img):
        h, w = img.shape
        out = np.zeros(h * w + 64, np.int16)
        src = np.zeros(h * w + 64, np.uint8); src[:h * w] = img.ravel()
        self.lib.ref_checkerboard5x5(_p(src), _p(out), w, h)
        return out[:h * w].reshape(h, w)

    def sad32(self, a, b):
        return self.lib.ref_sad32(_p(np.ascontiguousarray(a, np.uint8)), _p(np.ascontiguousarray(b, np.uint8)))

    def sad16(self, a, b):
        return self.lib.ref_sad16(_p(np.ascontiguousarray(a, np.uint8)), _p(np.ascontiguousarray(b, np.uint8)))

    def svd(self, A):
        A = np.ascontiguousarray(A, np.float64)
        m, n = A.shape
        U = np.zeros((m, m)); W = np.zeros(min(m, n)); V = np.zeros((n, n))
        self.lib.ref_svd(_p(A), m, n, _p(U), _p(W), _p(V))
        return U, W, V

    def matcher(self, params):
        return RefMatcher(self, params)

    def mono(self, params):
        return RefMono(self, params)

    def reconstruction(self):
        return Recon(self.lib, 'ref_recon')

    def stereo(self, params):
        return RefStereo(self, params)


class RefMatcher:
    WHICH = {'1p1': 0, '2p1': 1, '1c1': 2, '2c1': 3, '1p2': 4, '2p2': 5, '1c2': 6, '2c2': 7}

    def __init__(self, ref, params, handle=None):
        self.ref, self.lib = ref, ref.lib
        self.params = params
        self.owned = handle is None
        self.h = C.c_void_p(self.lib.ref_matcher_create(C.byref(params))) if handle is None else C.c_void_p(handle)

    def __del__(self):
        if getattr(self, 'owned', False) and self.h:
            self.lib.ref_matcher_destroy(self.h)
            self.h = None

    def push(self, I1, I2=None, replace=False):
        I1 = np.ascontiguousarray(I1, np.uint8)
        h, w = I1.shape
        dims = np.array([w, h, w], np.int32)
        if I2 is not None:
            I2 = np.ascontiguousarray(I2, np.uint8)
        self.lib.ref_matcher_push(self.h, _p(I1), _p(I2), _p(dims), int(replace))

    def counts(self):
        out = np.zeros(8, np.int32)
        self.lib.ref_matcher_counts(self.h, _p(out))
        return dict(zip(self.WHICH.keys(), out.tolist()))

    def maxima(self, which):
        k = self.WHICH[which] if isinstance(which, str) else which
        n = self.lib.ref_matcher_get_maxima(self.h, k, None)
        out = np.zeros((n, 12), np.int32)
        if n:
            self.lib.ref_matcher_get_maxima(self.h, k, _p(out))
        return out

    def sobel(self, which, full=False):
        k = {'1p': 0, '2p': 1, '1c': 2, '2c': 3}[which]
        dims = np.zeros(3, np.int32)
        if not self.lib.ref_matcher_get_sobel(self.h, k, int(full), None, None, _p(dims)):
            return None
        w, h, bpl = dims.tolist()
        du = np.zeros((h, bpl), np.uint8); dv = np.zeros((h, bpl), np.uint8)
        self.lib.ref_matcher_get_sobel(self.h, k, int(full), _p(du), _p(dv), _p(dims))
        return du, dv, (w, h, bpl)

    def half_image(self, Ipad, w):
        h, bpl = Ipad.shape
        dims = np.array([w, h, bpl], np.int32); dh = np.zeros(3, np.int32)
        out = np.zeros((h // 2) * (bpl), np.uint8)
        self.lib.ref_half_image(self.h, _p(np.ascontiguousarray(Ipad)), _p(dims), _p(out), _p(dh))
        wh, hh, bh = dh.tolist()
        return out[:hh * bh].reshape(hh, bh), (wh, hh, bh)

    def nms(self, f1, f2, w, n):
        h, bpl = f1.shape
        dims = np.array([w, h, bpl], np.int32)
        cap = 4 * (w // (n + 1) + 1) * (h // (n + 1) + 1)
        out = np.zeros((cap, 4), np.int32)
        cnt = self.lib.ref_nms(self.h, _p(np.ascontiguousarray(f1, np.int16)), _p(np.ascontiguousarray(f2, np.int16)),
                               _p(dims), n, _p(out), cap)
        assert cnt <= cap
        return out[:cnt]

    def descriptor(self, du, dv, u, v):
        out = np.zeros(32, np.uint8)
        self.lib.ref_descriptor(self.h, _p(du), _p(dv), du.shape[1], u, v, _p(out))
        return out

    def matching(self, pass_, method, use_prior):
        n = self.lib.ref_matcher_matching(self.h, pass_, method, int(use_prior), None, 0)
        out = np.zeros(n, P_MATCH)
        if n:
            self.lib.ref_matcher_matching(self.h, pass_, method, int(use_prior), _p(out), n)
        return out

    def remove_outliers(self, matches, method):
        m = np.array(matches, dtype=P_MATCH, copy=True)
        n = self.lib.ref_matcher_remove_outliers(self.h, _p(m), len(m), method)
        return m[:n].copy()

    def refinement(self, matches, method):
        m = np.array(matches, dtype=P_MATCH, copy=True)
        n = self.lib.ref_matcher_refinement(self.h, _p(m), len(m), method)
        return m[:n].copy()

    def prior(self, matches, method):
        m = np.ascontiguousarray(matches, dtype=P_MATCH)
        nb = self.lib.ref_matcher_prior(self.h, _p(m), len(m), method, None, 0)
        out = np.zeros((nb, 16), np.float32)
        self.lib.ref_matcher_get_ranges(self.h, _p(out), nb)
        return out

    def ranges(self):
        nb = self.lib.ref_matcher_get_ranges(self.h, None, 0)
        out = np.zeros((nb, 16), np.float32)
        if nb:
            self.lib.ref_matcher_get_ranges(self.h, _p(out), nb)
        return out

    def match_features(self, method, tr_delta=None):
        if tr_delta is None:
            self.lib.ref_matcher_match_features(self.h, method)
        else:
            self.lib.ref_matcher_match_features_tr(self.h, method, _p(np.ascontiguousarray(tr_delta, np.float64)))

    def set_intrinsics(self, f, cu, cv, base):
        self.lib.ref_matcher_set_intrinsics(self.h, C.c_double(f), C.c_double(cu), C.c_double(cv), C.c_double(base))

    def bucket(self, max_features, bw, bh):
        self.lib.ref_matcher_bucket(self.h, max_features, C.c_float(bw), C.c_float(bh))

    def matches(self, stage=2):
        n = self.lib.ref_matcher_get_matches(self.h, stage, None, 0)
        out = np.zeros(n, P_MATCH)
        if n:
            self.lib.ref_matcher_get_matches(self.h, stage, _p(out), n)
        return out

    def gain(self, inliers):
        a = np.ascontiguousarray(inliers, np.int32)
        return float(self.lib.ref_matcher_gain(self.h, _p(a), len(a)))


class Recon:
    """Reconstruction driven through a C shim; `prefix` selects the library's entry points (the reference shim and the
    B200 host library export the same five functions under different prefixes)."""

    def __init__(self, lib, prefix):
        self.lib, self.prefix = lib, prefix
        self.h = C.c_void_p(self._f('create')())

    def _f(self, name):
        return getattr(self.lib, '%s_%s' % (self.prefix, name))

    def __del__(self):
        if getattr(self, 'h', None):
            self._f('destroy')(self.h)
            self.h = None

    def set_calibration(self, f, cu, cv):
        self._f('set_calibration')(self.h, C.c_double(f), C.c_double(cu), C.c_double(cv))

    def update(self, matches, tr, point_type=1, min_track_length=2, max_dist=30.0, min_angle=2.0):
        matches = np.ascontiguousarray(matches, P_MATCH)
        tr = np.ascontiguousarray(tr, np.float64)
        self._f('update')(self.h, _p(matches), len(matches), _p(tr), int(point_type), int(min_track_length),
                          C.c_double(max_dist), C.c_double(min_angle))

    def points(self):
        n = self._f('get_points')(self.h, None, 0)
        out = np.zeros((n, 3), np.float32)
        if n:
            self._f('get_points')(self.h, _p(out), n)
        return out


class RefStereo:
    def __init__(self, ref, params):
        self.ref, self.lib, self.params = ref, ref.lib, params
        self.h = C.c_void_p(self.lib.ref_stereo_create(C.byref(params)))

    def __del__(self):
        if getattr(self, 'h', None):
            self.lib.ref_stereo_destroy(self.h)
            self.h = None

    def process(self, I1, I2, replace=False):
        I1 = np.ascontiguousarray(I1, np.uint8); I2 = np.ascontiguousarray(I2, np.uint8)
        h, w = I1.shape
        dims = np.array([w, h, w], np.int32)
        return bool(self.lib.ref_stereo_process(self.h, _p(I1), _p(I2), _p(dims), int(replace)))

    def process_matches(self, matches):
        m = np.ascontiguousarray(matches, dtype=P_MATCH)
        return bool(self.lib.ref_stereo_process_matches(self.h, _p(m), len(m)))

    def motion(self):
        out = np.zeros((4, 4))
        self.lib.ref_stereo_get_motion(self.h, _p(out))
        return out

    def matches(self):
        n = self.lib.ref_stereo_get_matches(self.h, None, 0)
        out = np.zeros(n, P_MATCH)
        if n:
            self.lib.ref_stereo_get_matches(self.h, _p(out), n)
        return out

    def inliers(self):
        n = self.lib.ref_stereo_get_inliers(self.h, None, 0)
        out = np.zeros(n, np.int32)
        if n:
            self.lib.ref_stereo_get_inliers(self.h, _p(out), n)
        return out


class RefMono:
    def __init__(self, ref, params):
        self.ref, self.lib, self.params = ref, ref.lib, params
        self.h = C.c_void_p(self.lib.ref_mono_create(C.byref(params)))
        self.matcher = RefMatcher(ref, params.match, handle=self.lib.ref_mono_matcher(self.h))

    def __del__(self):
        if getattr(self, 'h', None):
            self.lib.ref_mono_destroy(self.h)
            self.h = None

    def process(self, I, replace=False):
        I = np.ascontiguousarray(I, np.uint8)
        h, w = I.shape
        dims = np.array([w, h, w], np.int32)
        return bool(self.lib.ref_mono_process(self.h, _p(I), _p(dims), int(replace)))

    def process_matches(self, matches):
        m = np.ascontiguousarray(matches, dtype=P_MATCH)
        return bool(self.lib.ref_mono_process_matches(self.h, _p(m), len(m)))

    def motion(self):
        out = np.zeros((4, 4))
        self.lib.ref_mono_get_motion(self.h, _p(out))
        return out

    def matches(self):
        n = self.lib.ref_mono_get_matches(self.h, None, 0)
        out = np.zeros(n, P_MATCH)
        if n:
            self.lib.ref_mono_get_matches(self.h, _p(out), n)
        return out

    def inliers(self):
        n = self.lib.ref_mono_get_inliers(self.h, None, 0)
        out = np.zeros(n, np.int32)
        if n:
            self.lib.ref_mono_get_inliers(self.h, _p(out), n)
        return out

    def random_sample(self, N, num=8):
        out = np.zeros(num, np.int32)
        self.lib.ref_random_sample(self.h, N, num, _p(out))
        return out

    def normalize(self, matches):
        m = np.array(matches, dtype=P_MATCH, copy=True)
        Tp = np.zeros((3, 3)); Tc = np.zeros((3, 3))
        ok = self.lib.ref_normalize(self.h, _p(m), len(m), _p(Tp), _p(Tc))
        return bool(ok), m, Tp, Tc

    def fundamental(self, matches, active):
        m = np.ascontiguousarray(matches, dtype=P_MATCH)
        a = np.ascontiguousarray(active, np.int32)
        F = np.zeros((3, 3))
        self.lib.ref_fundamental(self.h, _p(m), len(m), _p(a), len(a), _p(F))
        return F

    def get_inlier(self, matches, F):
        m = np.ascontiguousarray(matches, dtype=P_MATCH)
        out = np.zeros(len(m), np.int32)
        n = self.lib.ref_get_inlier(self.h, _p(m), len(m), _p(np.ascontiguousarray(F, np.float64)), _p(out))
        return out[:n].copy()

    def ransac_with_samples(self, matches, samples, want_all=True):
        m = np.ascontiguousarray(matches, dtype=P_MATCH)
        s = np.ascontiguousarray(samples, np.int32)
        iters = len(s)
        F = np.zeros((3, 3)); inl = np.zeros(len(m), np.int32)
        counts = np.zeros(iters, np.int32); Fall = np.zeros((iters, 3, 3)); best = C.c_int32(-1)
        n = self.lib.ref_ransac_with_samples(self.h, _p(m), len(m), _p(s), iters, _p(F), _p(inl),
                                             _p(counts) if want_all else None, _p(Fall) if want_all else None, C.byref(best))
        return dict(n_inliers=n, F=F, inliers=inl[:max(n, 0)].copy(), counts=counts, F_all=Fall, best_iter=best.value)

    def triangulate_chieral(self, matches, K, R, t):
        """VisualOdometryMono::triangulateChieral (viso_mono.cpp:394-431): (X 4 x n, points in front of both cameras)."""
        m = np.ascontiguousarray(matches, dtype=P_MATCH)
        X = np.zeros((4, len(m)))
        num = self.lib.ref_triangulate_chieral(self.h, _p(m), len(m), _p(np.ascontiguousarray(K, np.float64)),
                                               _p(np.ascontiguousarray(R, np.float64)), _p(np.ascontiguousarray(t, np.float64).ravel()), _p(X))
        return X, num

    def find_best_plane(self, x_plane, threshold, weight):
        """VisualOdometryMono::findBestPlane (viso_mono.cpp:74-98); x_plane 2 x n.  Returns the winning distance."""
        xp = np.ascontiguousarray(x_plane, np.float64)
        return float(self.lib.ref_find_best_plane(self.h, _p(xp), xp.shape[1], C.c_double(threshold), C.c_double(weight)))

    def estimate_motion(self, matches):
        m = np.ascontiguousarray(matches, dtype=P_MATCH)
        tr = np.zeros(6)
        ok = self.lib.ref_estimate_motion(self.h, _p(m), len(m), _p(tr))
        return bool(ok), tr


def time_matcher_sequence(ref, params, method, imgs, imgs2=None, bucket=None):
    """Reference CPU timing of pushBack+matchFeatures[+bucketFeatures] over a sequence (first frame untimed)."""
    imgs = np.ascontiguousarray(imgs, np.uint8)
    n, h, w = imgs.shape
    dims = np.array([w, h, w], np.int32)
    per = np.zeros(n - 1); nm = np.zeros(n - 1, np.int32)
    if imgs2 is not None:
        imgs2 = np.ascontiguousarray(imgs2, np.uint8)
    bm, bw, bh = bucket if bucket else (0, 50.0, 50.0)
    tot = ref.lib.ref_time_matcher_sequence(C.byref(params), method, _p(imgs), _p(imgs2), C.c_size_t(h * w), _p(dims), n,
                                            bm, C.c_float(bw), C.c_float(bh), _p(per), _p(nm))
    return tot, per, nm


def time_mono_sequence(ref, params, imgs):
    imgs = np.ascontiguousarray(imgs, np.uint8)
    n, h, w = imgs.shape
    dims = np.array([w, h, w], np.int32)
    per = np.zeros(n - 1); ok = np.zeros(n - 1, np.int32); mot = np.zeros((n - 1, 4, 4))
    tot = ref.lib.ref_time_mono_sequence(C.byref(params), _p(imgs), C.c_size_t(h * w), _p(dims), n, _p(per), _p(ok), _p(mot))
    return tot, per, ok, mot


def time_parallel(ref, mono_params, workload, imgs, imgs2, nthreads, warm_pairs, timed_pairs, bucket=None):
    """The reference on nthreads persistent host threads (one sequence each); returns (wall seconds of the timed pairs,
    pairs per thread).  workload: 0 flow, 1 stereo quad, 2 mono odometry."""
    imgs = np.ascontiguousarray(imgs, np.uint8)
    n, h, w = imgs.shape
    dims = np.array([w, h, w], np.int32)
    if imgs2 is not None:
        imgs2 = np.ascontiguousarray(imgs2, np.uint8)
    bm, bw, bh = bucket if bucket else (0, 50.0, 50.0)
    done = np.zeros(nthreads, np.int32)
    wall = ref.lib.ref_time_parallel(C.byref(mono_params), int(workload), _p(imgs), _p(imgs2), C.c_size_t(h * w), n, _p(dims), int(nthreads),
                                     int(warm_pairs), int(timed_pairs), int(bm), C.c_float(bw), C.c_float(bh), _p(done))
    return wall, done
