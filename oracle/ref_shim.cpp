// TEST INFRASTRUCTURE ONLY -- never linked into the product.
//
// C dumping shim over the UNMODIFIED reference CPU sources (libviso2 fork at
// $VISO_REF_DIR, default /root/reference).  oracle/Makefile compiles
// viso/{filter,matcher,matrix,triangle,viso,viso_mono}.cpp where they lie and
// links them with this file into oracle/_ref/libvisoref*.so; nothing from the
// reference is copied into this repository.  Only tests/, __graft_entry__.smoke()
// and bench.py's cpu_baseline / --impl reference legs may load the result.
//
// The reference keeps every stage output private (matcher.h:138-245,
// viso_mono.h:66-86); the shim reads them by including the reference headers
// with the access keywords neutralised (standard headers are pulled in first so
// only the reference's own classes are affected).

#include <stdint.h>
#include <cstdint>
#include <cstring>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <array>
#include <iostream>
#include <sstream>
#include <algorithm>
#include <random>
#include <iterator>
#include <limits>
#include <chrono>
#include <mutex>
#include <thread>
#include <atomic>
#include <streambuf>
#include <mm_malloc.h>

#define private public
#define protected public
#include "matcher.h"
#include "reconstruction.h"
#include "filter.h"
#include "viso_mono.h"
#include "viso_stereo.h"
#undef private
#undef protected

#define REF_API extern "C" __attribute__((visibility("default")))

namespace {

struct RefMatcherParams {  // mirrors Matcher::parameters (matcher.h:42-69)
  int32_t nms_n, nms_tau, match_binsize, match_radius, match_disp_tolerance;
  int32_t outlier_disp_tolerance, outlier_flow_tolerance, multi_stage, half_resolution, refinement;
  double f, cu, cv, base;
};

Matcher::parameters to_ref(const RefMatcherParams* p) {
  Matcher::parameters q;
  q.nms_n = p->nms_n; q.nms_tau = p->nms_tau; q.match_binsize = p->match_binsize;
  q.match_radius = p->match_radius; q.match_disp_tolerance = p->match_disp_tolerance;
  q.outlier_disp_tolerance = p->outlier_disp_tolerance; q.outlier_flow_tolerance = p->outlier_flow_tolerance;
  q.multi_stage = p->multi_stage; q.half_resolution = p->half_resolution; q.refinement = p->refinement;
  q.f = p->f; q.cu = p->cu; q.cv = p->cv; q.base = p->base;
  return q;
}

struct RefMonoParams {
  RefMatcherParams match;
  int32_t bucket_max_features; double bucket_width, bucket_height;
  double f, cu, cv;
  double height, pitch; int32_t ransac_iters; double inlier_threshold, motion_threshold;
};

// subclass only to reach the private virtuals/non-virtuals through one object
struct MonoProbe : public VisualOdometryMono {
  explicit MonoProbe(VisualOdometryMono::parameters p) : VisualOdometryMono(p) {}
};

VisualOdometryMono::parameters to_ref(const RefMonoParams* p) {
  VisualOdometryMono::parameters q;
  q.match = to_ref(&p->match);
  q.bucket.max_features = p->bucket_max_features;
  q.bucket.bucket_width = p->bucket_width; q.bucket.bucket_height = p->bucket_height;
  q.calib.f = p->f; q.calib.cu = p->cu; q.calib.cv = p->cv;
  q.height = p->height; q.pitch = p->pitch; q.ransac_iters = p->ransac_iters;
  q.inlier_threshold = p->inlier_threshold; q.motion_threshold = p->motion_threshold;
  return q;
}

// StartTimer prints "Estimate F time" / "Best plane time" on every estimateMotion call
// (timer.hh:9-34, viso_mono.cpp:117-173); keep the test/bench stdout clean.  Several threads may
// run reference objects at once (bench.py times one sequence per host core), so the silencer is
// reference counted and the sink is a stateless discard buffer.
struct DiscardBuf : public std::streambuf {
  int overflow(int c) override { return c; }
  std::streamsize xsputn(const char*, std::streamsize n) override { return n; }
};
struct CoutSilencer {
  static std::mutex& mtx() { static std::mutex m; return m; }
  static int& users() { static int n = 0; return n; }
  static std::streambuf*& saved() { static std::streambuf* p = nullptr; return p; }
  static DiscardBuf& sink() { static DiscardBuf b; return b; }
  CoutSilencer() {
    std::lock_guard<std::mutex> lock(mtx());
    if (users()++ == 0) saved() = std::cout.rdbuf(&sink());
  }
  ~CoutSilencer() {
    std::lock_guard<std::mutex> lock(mtx());
    if (--users() == 0) std::cout.rdbuf(saved());
  }
};

int copy_matches(const std::vector<Matcher::p_match>& v, void* out, int cap) {
  int n = (int)v.size();
  if (out && n > 0) memcpy(out, v.data(), sizeof(Matcher::p_match) * (size_t)std::min(n, cap));
  return n;
}

}  // namespace

static_assert(sizeof(Matcher::p_match) == 48, "p_match layout");
static_assert(sizeof(Matcher::maximum) == 48, "maximum layout");

// ---------------------------------------------------------------- build info
REF_API const char* ref_build_info() {
#ifdef __FMA__
  #ifdef REF_NO_CONTRACT
  return "reference r30, g++ " __VERSION__ ", USE_SIMD, -ffp-contract=off";
  #else
  return "reference r30, g++ " __VERSION__ ", USE_SIMD, fma contraction allowed";
  #endif
#else
  return "reference r30, g++ " __VERSION__ ", USE_SIMD, no fma isa";
#endif
}

// ------------------------------------------------------------------- filters
// The reference row passes read a few int16 past the temp planes and write 16-B
// chunks past out+2; callers pass buffers with >= 64 bytes of slack.
REF_API void ref_sobel5x5(const uint8_t* in, uint8_t* out_v, uint8_t* out_h, int w, int h) { filter::sobel5x5(in, out_v, out_h, w, h); }
REF_API void ref_sobel3x3(const uint8_t* in, uint8_t* out_v, uint8_t* out_h, int w, int h) { filter::sobel3x3(in, out_v, out_h, w, h); }
REF_API void ref_blob5x5(const uint8_t* in, int16_t* out, int w, int h) { filter::blob5x5(in, out, w, h); }
REF_API void ref_checkerboard5x5(const uint8_t* in, int16_t* out, int w, int h) { filter::checkerboard5x5(in, out, w, h); }

// simd.hh known-answer probes (test/simd.cpp covers the same ops)
REF_API int32_t ref_sad32(const uint8_t* a, const uint8_t* b) {
  using namespace simd;
  alignas(16) uint8_t ta[32], tb[32];
  memcpy(ta, a, 32); memcpy(tb, b, 32);
  return sad_array(load_aligned_array((array_16xuint8_t*)ta), load_aligned_array((array_16xuint8_t*)(ta + 16)),
                   load_aligned_array((array_16xuint8_t*)tb), load_aligned_array((array_16xuint8_t*)(tb + 16)));
}
REF_API int32_t ref_sad16(const uint8_t* a, const uint8_t* b) {
  using namespace simd;
  alignas(16) uint8_t ta[16], tb[16];
  memcpy(ta, a, 16); memcpy(tb, b, 16);
  return sad_array(load_aligned_array((array_16xuint8_t*)ta), load_aligned_array((array_16xuint8_t*)tb));
}

// ------------------------------------------------------------------- Matcher
REF_API void* ref_matcher_create(const RefMatcherParams* p) { return new Matcher(to_ref(p)); }
REF_API void ref_matcher_destroy(void* m) { delete (Matcher*)m; }
REF_API void ref_matcher_push(void* m, uint8_t* I1, uint8_t* I2, const int32_t* dims, int replace) {
  uint32_t d[3] = {(uint32_t)dims[0], (uint32_t)dims[1], (uint32_t)dims[2]};
  ((Matcher*)m)->pushBack(I1, I2, d, replace != 0);
}
// order: 1p1 2p1 1c1 2c1 1p2 2p2 1c2 2c2   (suffix 1 = sparse pass, 2 = dense pass)
REF_API void ref_matcher_counts(void* m_, int32_t* out) {
  Matcher* m = (Matcher*)m_;
  int32_t v[8] = {m->n1p1, m->n2p1, m->n1c1, m->n2c1, m->n1p2, m->n2p2, m->n1c2, m->n2c2};
  memcpy(out, v, sizeof v);
}
REF_API int ref_matcher_get_maxima(void* m_, int which, int32_t* out) {
  Matcher* m = (Matcher*)m_;
  int32_t* ptr[8] = {m->m1p1, m->m2p1, m->m1c1, m->m2c1, m->m1p2, m->m2p2, m->m1c2, m->m2c2};
  int32_t n[8] = {m->n1p1, m->n2p1, m->n1c1, m->n2c1, m->n1p2, m->n2p2, m->n1c2, m->n2c2};
  if (out && ptr[which] && n[which] > 0) memcpy(out, ptr[which], (size_t)n[which] * 48);
  return ptr[which] ? n[which] : 0;
}
// which: 0=1p 1=2p 2=1c 3=2c ; full: 0 = matching-resolution planes, 1 = *_full planes.
// dims_out = {w,h,bpl} of the plane.  Returns 0 if the plane does not exist.
REF_API int ref_matcher_get_sobel(void* m_, int which, int full, uint8_t* du, uint8_t* dv, int32_t* dims_out) {
  Matcher* m = (Matcher*)m_;
  uint8_t* pu[8] = {m->I1p_du, m->I2p_du, m->I1c_du, m->I2c_du, m->I1p_du_full, m->I2p_du_full, m->I1c_du_full, m->I2c_du_full};
  uint8_t* pv[8] = {m->I1p_dv, m->I2p_dv, m->I1c_dv, m->I2c_dv, m->I1p_dv_full, m->I2p_dv_full, m->I1c_dv_full, m->I2c_dv_full};
  const int32_t* dsrc = (which == 0 || which == 1) ? m->dims_p : m->dims_c;
  int32_t d[3] = {dsrc[0], dsrc[1], dsrc[2]};
  if (!full && m->param.half_resolution) m->getHalfResolutionDimensions(dsrc, d);
  if (full && !m->param.half_resolution) return 0;
  int k = which + (full ? 4 : 0);
  if (!pu[k] || !pv[k]) return 0;
  if (dims_out) memcpy(dims_out, d, sizeof d);
  if (du) memcpy(du, pu[k], (size_t)d[1] * d[2]);
  if (dv) memcpy(dv, pv[k], (size_t)d[1] * d[2]);
  return 1;
}
REF_API void ref_half_image(void* m_, uint8_t* I, const int32_t* dims, uint8_t* out, int32_t* dims_half) {
  Matcher* m = (Matcher*)m_;
  m->getHalfResolutionDimensions(dims, dims_half);
  uint8_t* h = m->createHalfResolutionImage(I, dims);
  // pad columns are never written by the reference (matcher.cpp:639-645); copy the valid part only
  for (int v = 0; v < dims_half[1]; v++) memcpy(out + (size_t)v * dims_half[2], h + (size_t)v * dims_half[2], dims_half[0]);
  _mm_free(h);
}
// raw nonMaximumSuppression on caller-provided response maps; out = (u,v,val,c) per maximum
REF_API int ref_nms(void* m_, int16_t* f1, int16_t* f2, const int32_t* dims, int nms_n, int32_t* out, int cap) {
  Matcher* m = (Matcher*)m_;
  std::vector<Matcher::maximum> mx;
  m->nonMaximumSuppression(f1, f2, dims, mx, nms_n);
  int n = (int)mx.size();
  for (int i = 0; i < std::min(n, cap); i++) {
    out[4 * i + 0] = mx[i].u; out[4 * i + 1] = mx[i].v; out[4 * i + 2] = mx[i].val; out[4 * i + 3] = mx[i].c;
  }
  return n;
}
// computeDescriptor/computeSmallDescriptor are `inline` members defined in matcher.cpp (no symbol is
// emitted for other TUs), so the 32-byte descriptor is obtained through computeDescriptors on one maximum.
REF_API void ref_descriptor(void* m_, uint8_t* du, uint8_t* dv, int bpl, int u, int v, uint8_t* out32) {
  std::vector<Matcher::maximum> mx(1, Matcher::maximum(u, v, 0, 0));
  ((Matcher*)m_)->computeDescriptors(du, dv, bpl, mx);
  memcpy(out32, &mx[0].d1, 32);
}
// private matching() on the ring buffer: pass 0 = sparse sets, 1 = dense sets (matcher.cpp:222,229)
REF_API int ref_matcher_matching(void* m_, int pass, int method, int use_prior, void* out, int cap) {
  Matcher* m = (Matcher*)m_;
  std::vector<Matcher::p_match> pm;
  if (pass == 0) m->matching(m->m1p1, m->m2p1, m->m1c1, m->m2c1, m->n1p1, m->n2p1, m->n1c1, m->n2c1, pm, method, use_prior != 0, 0);
  else           m->matching(m->m1p2, m->m2p2, m->m1c2, m->m2c2, m->n1p2, m->n2p2, m->n1c2, m->n2c2, pm, method, use_prior != 0, 0);
  return copy_matches(pm, out, cap);
}
REF_API int ref_matcher_remove_outliers(void* m_, void* inout, int n, int method) {
  std::vector<Matcher::p_match> pm((Matcher::p_match*)inout, (Matcher::p_match*)inout + n);
  ((Matcher*)m_)->removeOutliers(pm, method);
  return copy_matches(pm, inout, n);
}
REF_API int ref_matcher_refinement(void* m_, void* inout, int n, int method) {
  std::vector<Matcher::p_match> pm((Matcher::p_match*)inout, (Matcher::p_match*)inout + n);
  ((Matcher*)m_)->refinement(pm, method);
  return copy_matches(pm, inout, n);
}
// computePriorStatistics; ranges_out = bin_num x 16 floats {u_min[4],u_max[4],v_min[4],v_max[4]} (matcher.h:152-157)
REF_API int ref_matcher_prior(void* m_, void* matches, int n, int method, float* ranges_out, int cap_bins) {
  Matcher* m = (Matcher*)m_;
  std::vector<Matcher::p_match> pm((Matcher::p_match*)matches, (Matcher::p_match*)matches + n);
  m->computePriorStatistics(pm, method);
  int nb = (int)m->ranges.size();
  if (ranges_out) memcpy(ranges_out, m->ranges.data(), sizeof(Matcher::range) * (size_t)std::min(nb, cap_bins));
  return nb;
}
REF_API int ref_matcher_get_ranges(void* m_, float* ranges_out, int cap_bins) {
  Matcher* m = (Matcher*)m_;
  int nb = (int)m->ranges.size();
  if (ranges_out) memcpy(ranges_out, m->ranges.data(), sizeof(Matcher::range) * (size_t)std::min(nb, cap_bins));
  return nb;
}
REF_API void ref_matcher_match_features(void* m_, int method) { ((Matcher*)m_)->matchFeatures(method, 0); }
REF_API void ref_matcher_bucket(void* m_, int max_features, float bw, float bh) { ((Matcher*)m_)->bucketFeatures(max_features, bw, bh); }
// stage 1 = p_matched_1 (after removeOutliers of pass 1), 2 = p_matched_2 (== getMatches())
REF_API int ref_matcher_get_matches(void* m_, int stage, void* out, int cap) {
  Matcher* m = (Matcher*)m_;
  return copy_matches(stage == 1 ? m->p_matched_1 : m->p_matched_2, out, cap);
}
REF_API float ref_matcher_gain(void* m_, const int32_t* inl, int n) {
  return ((Matcher*)m_)->getGain(std::vector<int32_t>(inl, inl + n));
}

// --------------------------------------------------------------- mono odometry
REF_API void* ref_mono_create(const RefMonoParams* p) { return new MonoProbe(to_ref(p)); }
REF_API void ref_mono_destroy(void* v) { delete (MonoProbe*)v; }
REF_API int ref_mono_process(void* v, uint8_t* I, const int32_t* dims, int replace) {
  uint32_t d[3] = {(uint32_t)dims[0], (uint32_t)dims[1], (uint32_t)dims[2]};
  CoutSilencer quiet;
  return ((MonoProbe*)v)->process(I, d, replace != 0) ? 1 : 0;
}
REF_API int ref_mono_process_matches(void* v, const void* matches, int n) {
  std::vector<Matcher::p_match> pm((const Matcher::p_match*)matches, (const Matcher::p_match*)matches + n);
  CoutSilencer quiet;
  return ((VisualOdometry*)(MonoProbe*)v)->process(pm) ? 1 : 0;
}
REF_API void ref_mono_get_motion(void* v, double* out16) {
  Matrix T = ((MonoProbe*)v)->getMotion();
  for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) out16[4 * i + j] = T.val[i][j];
}
REF_API int ref_mono_get_matches(void* v, void* out, int cap) { return copy_matches(((MonoProbe*)v)->p_matched, out, cap); }
REF_API int ref_mono_get_inliers(void* v, int32_t* out, int cap) {
  std::vector<int32_t> in = ((MonoProbe*)v)->getInlierIndices();
  int n = (int)in.size();
  if (out) memcpy(out, in.data(), sizeof(int32_t) * (size_t)std::min(n, cap));
  return n;
}
REF_API void* ref_mono_matcher(void* v) { return ((MonoProbe*)v)->matcher; }
// draws from the reference's function-static generator (viso.cpp:86-102): process-wide state!
REF_API void ref_random_sample(void* v, int N, int num, int32_t* out) {
  std::vector<int> s = ((MonoProbe*)v)->getRandomSample((unsigned)N, (unsigned)num);
  for (int i = 0; i < num; i++) out[i] = s[i];
}
REF_API int ref_normalize(void* v, void* inout, int n, double* Tp9, double* Tc9) {
  std::vector<Matcher::p_match> pm((Matcher::p_match*)inout, (Matcher::p_match*)inout + n);
  Matrix Tp, Tc;
  bool ok = ((MonoProbe*)v)->normalizeFeaturePoints(pm, Tp, Tc);
  memcpy(inout, pm.data(), sizeof(Matcher::p_match) * (size_t)n);
  if (ok) for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) { Tp9[3 * i + j] = Tp.val[i][j]; Tc9[3 * i + j] = Tc.val[i][j]; }
  return ok ? 1 : 0;
}
REF_API void ref_fundamental(void* v, const void* matches, int n, const int32_t* active, int nactive, double* F9) {
  std::vector<Matcher::p_match> pm((const Matcher::p_match*)matches, (const Matcher::p_match*)matches + n);
  std::vector<int32_t> act(active, active + nactive);
  Matrix F;
  ((MonoProbe*)v)->fundamentalMatrix(pm, act, F);
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) F9[3 * i + j] = F.val[i][j];
}
REF_API int ref_get_inlier(void* v, const void* matches, int n, const double* F9, int32_t* out) {
  std::vector<Matcher::p_match> pm((const Matcher::p_match*)matches, (const Matcher::p_match*)matches + n);
  Matrix F(3, 3, F9);
  std::vector<int32_t> in = ((MonoProbe*)v)->getInlier(pm, F);
  if (out) memcpy(out, in.data(), sizeof(int32_t) * in.size());
  return (int)in.size();
}
// RANSAC driven by an explicit sample table (iters x 8) instead of the static generator, but
// otherwise the loop of viso_mono.cpp:41-72.  counts_out (iters) / F_all (iters x 9) may be null.
REF_API int ref_ransac_with_samples(void* v, const void* matches, int n, const int32_t* samples, int iters,
                                    double* F9, int32_t* inliers_out, int32_t* counts_out, double* F_all, int32_t* best_iter) {
  MonoProbe* vo = (MonoProbe*)v;
  std::vector<Matcher::p_match> pm((const Matcher::p_match*)matches, (const Matcher::p_match*)matches + n);
  std::vector<int32_t> best; Matrix F; int bi = -1;
  for (int k = 0; k < iters; k++) {
    std::vector<int32_t> act(samples + 8 * k, samples + 8 * k + 8);
    vo->fundamentalMatrix(pm, act, F);
    if (F_all) for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) F_all[9 * k + 3 * i + j] = F.val[i][j];
    std::vector<int32_t> cur = vo->getInlier(pm, F);
    if (counts_out) counts_out[k] = (int32_t)cur.size();
    if (cur.size() > best.size()) { best = cur; bi = k; }
  }
  if (best_iter) *best_iter = bi;
  if (best.size() < 10) { for (int i = 0; i < 9; i++) F9[i] = 0; return -(int)best.size() - 1; }
  vo->fundamentalMatrix(pm, best, F);
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) F9[3 * i + j] = F.val[i][j];
  if (inliers_out) memcpy(inliers_out, best.data(), sizeof(int32_t) * best.size());
  return (int)best.size();
}
// the virtual hook itself (uses the static generator)
REF_API int ref_ransac(void* v, const void* matches, int n, double* F9, int32_t* inliers_out) {
  MonoProbe* vo = (MonoProbe*)v;
  std::vector<Matcher::p_match> pm((const Matcher::p_match*)matches, (const Matcher::p_match*)matches + n);
  Matrix F = vo->ransacEstimateF(pm);
  if (F.val == nullptr) return -1;
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) F9[3 * i + j] = F.val[i][j];
  if (inliers_out) memcpy(inliers_out, vo->inliers.data(), sizeof(int32_t) * vo->inliers.size());
  return (int)vo->inliers.size();
}
REF_API int ref_estimate_motion(void* v, const void* matches, int n, double* tr6) {
  std::vector<Matcher::p_match> pm((const Matcher::p_match*)matches, (const Matcher::p_match*)matches + n);
  CoutSilencer quiet;
  std::vector<double> tr = ((MonoProbe*)v)->estimateMotion(pm);
  if (tr.size() != 6) return 0;
  for (int i = 0; i < 6; i++) tr6[i] = tr[i];
  return 1;
}
// VisualOdometryMono::triangulateChieral (viso_mono.cpp:394-431) for one (R|t) candidate: K, R 3x3 row-major, t 3;
// X4n receives the 4 x n homogeneous points (row-major), the return value is the number of points in front of both cameras
REF_API int ref_triangulate_chieral(void* v, const void* matches, int n, const double* K9, const double* R9, const double* t3, double* X4n) {
  std::vector<Matcher::p_match> pm((const Matcher::p_match*)matches, (const Matcher::p_match*)matches + n);
  Matrix K(3, 3, K9), R(3, 3, R9), t(3, 1, t3), X;
  const int num = ((MonoProbe*)v)->triangulateChieral(pm, K, R, t, X);
  if (X4n) for (int i = 0; i < 4; i++) for (int j = 0; j < n; j++) X4n[(size_t)i * n + j] = X.val[i][j];
  return num;
}
// VisualOdometryMono::findBestPlane (viso_mono.cpp:74-98): x_plane is 2 x n row-major; returns the winning plane distance
REF_API double ref_find_best_plane(void* v, const double* x_plane2n, int n, double threshold, double weight) {
  Matrix xp(2, n, x_plane2n);
  return ((MonoProbe*)v)->findBestPlane(xp, threshold, weight);
}
// Matrix::svd (matrix.cpp:586-814): A is m x n row-major; U m x m, W min(m,n), V n x n
REF_API void ref_svd(const double* A, int m, int n, double* U, double* W, double* V) {
  Matrix M(m, n, A), Um, Wm, Vm;
  M.svd(Um, Wm, Vm);
  for (int i = 0; i < m; i++) for (int j = 0; j < m; j++) U[i * m + j] = Um.val[i][j];
  for (int i = 0; i < std::min(m, n); i++) W[i] = Wm.val[i][0];
  for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) V[i * n + j] = Vm.val[i][j];
}

// ------------------------------------------------------------- stereo odometry
namespace {
struct RefStereoParams {
  RefMatcherParams match;
  int32_t bucket_max_features; double bucket_width, bucket_height;
  double f, cu, cv;
  double base; int32_t ransac_iters; double inlier_threshold; int32_t reweighting;
};
VisualOdometryStereo::parameters to_ref(const RefStereoParams* p) {
  VisualOdometryStereo::parameters q;
  q.match = to_ref(&p->match);
  q.bucket.max_features = p->bucket_max_features;
  q.bucket.bucket_width = p->bucket_width; q.bucket.bucket_height = p->bucket_height;
  q.calib.f = p->f; q.calib.cu = p->cu; q.calib.cv = p->cv;
  q.base = p->base; q.ransac_iters = p->ransac_iters; q.inlier_threshold = p->inlier_threshold; q.reweighting = p->reweighting != 0;
  return q;
}
}  // namespace
// ---- Reconstruction (reconstruction.h:40-67), driven with explicit match lists and motions
REF_API void* ref_recon_create() { return new Reconstruction(); }
REF_API void ref_recon_destroy(void* r) { delete (Reconstruction*)r; }
REF_API void ref_recon_set_calibration(void* r, double f, double cu, double cv) { ((Reconstruction*)r)->setCalibration(f, cu, cv); }
REF_API void ref_recon_update(void* r, const void* matches, int n, const double* tr16, int point_type, int min_track_length,
                              double max_dist, double min_angle) {
  const Matcher::p_match* m = (const Matcher::p_match*)matches;
  ((Reconstruction*)r)->update(std::vector<Matcher::p_match>(m, m + n), Matrix(4, 4, tr16), point_type, min_track_length, max_dist, min_angle);
}
REF_API int ref_recon_get_points(void* r, float* out3, int cap) {
  const std::vector<Point3d>& pts = ((Reconstruction*)r)->getPoints();
  for (int i = 0; i < (int)pts.size() && i < cap; i++) { out3[3 * i] = pts[i].x; out3[3 * i + 1] = pts[i].y; out3[3 * i + 2] = pts[i].z; }
  return (int)pts.size();
}

REF_API void* ref_stereo_create(const RefStereoParams* p) { return new VisualOdometryStereo(to_ref(p)); }
REF_API void ref_stereo_destroy(void* v) { delete (VisualOdometryStereo*)v; }
REF_API int ref_stereo_process(void* v, uint8_t* I1, uint8_t* I2, const int32_t* dims, int replace) {
  uint32_t d[3] = {(uint32_t)dims[0], (uint32_t)dims[1], (uint32_t)dims[2]};
  return ((VisualOdometryStereo*)v)->process(I1, I2, d, replace != 0) ? 1 : 0;
}
REF_API int ref_stereo_process_matches(void* v, const void* matches, int n) {
  std::vector<Matcher::p_match> pm((const Matcher::p_match*)matches, (const Matcher::p_match*)matches + n);
  CoutSilencer quiet;
  return ((VisualOdometry*)(VisualOdometryStereo*)v)->process(pm) ? 1 : 0;
}
REF_API void ref_stereo_get_motion(void* v, double* out16) {
  Matrix T = ((VisualOdometryStereo*)v)->getMotion();
  for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) out16[4 * i + j] = T.val[i][j];
}
REF_API int ref_stereo_get_matches(void* v, void* out, int cap) { return copy_matches(((VisualOdometryStereo*)v)->p_matched, out, cap); }
REF_API int ref_stereo_get_inliers(void* v, int32_t* out, int cap) {
  std::vector<int32_t> in = ((VisualOdometryStereo*)v)->getInlierIndices();
  if (out) memcpy(out, in.data(), sizeof(int32_t) * (size_t)std::min((int)in.size(), cap));
  return (int)in.size();
}
// quad matching with the motion-predicted search window (matcher.cpp:1112-1138); tr16 = 4x4 row-major
REF_API void ref_matcher_match_features_tr(void* m_, int method, const double* tr16) {
  Matrix T(4, 4, tr16);
  ((Matcher*)m_)->matchFeatures(method, &T);
}
REF_API void ref_matcher_set_intrinsics(void* m_, double f, double cu, double cv, double base) { ((Matcher*)m_)->setIntrinsics(f, cu, cv, base); }

// ------------------------------------------------------------ CPU baseline timing
// Times pushBack + matchFeatures(method) [+ bucketFeatures] exactly as main.cpp drives them, on
// nframes images (image k = imgs + k*stride; right images from imgs2 or null).  The first frame only
// fills the ring buffer and is not timed.  Returns seconds for (nframes-1) pairs;
// per_pair_s (nframes-1) optional.
REF_API double ref_time_matcher_sequence(const RefMatcherParams* p, int method, uint8_t* imgs, uint8_t* imgs2, size_t stride,
                                         const int32_t* dims, int nframes, int bucket_max, float bw, float bh,
                                         double* per_pair_s, int32_t* n_matches) {
  Matcher m(to_ref(p));
  uint32_t d[3] = {(uint32_t)dims[0], (uint32_t)dims[1], (uint32_t)dims[2]};
  m.pushBack(imgs, imgs2, d, false);
  double total = 0;
  for (int k = 1; k < nframes; k++) {
    auto t0 = std::chrono::steady_clock::now();
    m.pushBack(imgs + k * stride, imgs2 ? imgs2 + k * stride : 0, d, false);
    m.matchFeatures(method, 0);
    if (bucket_max > 0) m.bucketFeatures(bucket_max, bw, bh);
    std::vector<Matcher::p_match> r = m.getMatches();
    auto t1 = std::chrono::steady_clock::now();
    double s = std::chrono::duration<double>(t1 - t0).count();
    total += s;
    if (per_pair_s) per_pair_s[k - 1] = s;
    if (n_matches) n_matches[k - 1] = (int32_t)r.size();
  }
  return total;
}
REF_API double ref_time_mono_sequence(const RefMonoParams* p, uint8_t* imgs, size_t stride, const int32_t* dims, int nframes,
                                      double* per_pair_s, int32_t* ok_out, double* motions16) {
  MonoProbe vo(to_ref(p));
  uint32_t d[3] = {(uint32_t)dims[0], (uint32_t)dims[1], (uint32_t)dims[2]};
  CoutSilencer quiet;
  vo.process(imgs, d, false);
  double total = 0;
  for (int k = 1; k < nframes; k++) {
    auto t0 = std::chrono::steady_clock::now();
    bool ok = vo.process(imgs + k * stride, d, false);
    auto t1 = std::chrono::steady_clock::now();
    double s = std::chrono::duration<double>(t1 - t0).count();
    total += s;
    if (per_pair_s) per_pair_s[k - 1] = s;
    if (ok_out) ok_out[k - 1] = ok ? 1 : 0;
    if (motions16) { Matrix T = vo.getMotion(); for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) motions16[16 * (k - 1) + 4 * i + j] = T.val[i][j]; }
  }
  return total;
}

// The reference on all host cores, for bench.py's reference arm and cpu_baseline: nthreads persistent threads, one
// independent sequence each (the reference is single-threaded per sequence).  Every thread builds its own object, pushes
// its first frame, runs warm_pairs frame pairs, meets the others at a barrier and then runs timed_pairs pairs; only those
// are timed (thread creation, construction and the first frame are not).  workload 0: Matcher flow (pushBack +
// matchFeatures(0) + getMatches), 1: stereo quad (+ bucketFeatures), 2: VisualOdometryMono::process.  Frames come from a
// pool of nframes consecutive frames (imgs + k * stride; right images imgs2), walked back and forth so that consecutive
// frames are always neighbours; thread t starts at frame 3 t.  Returns the wall-clock seconds from the barrier to the
// last thread's finish; pairs_done (optional) receives the timed pairs of every thread.
REF_API double ref_time_parallel(const RefMonoParams* mp, int workload, uint8_t* imgs, uint8_t* imgs2, size_t stride, int nframes,
                                 const int32_t* dims, int nthreads, int warm_pairs, int timed_pairs, int bucket_max, float bw, float bh,
                                 int32_t* pairs_done) {
  std::atomic<int> ready(0), go(0);
  std::vector<double> finish(nthreads, 0.0);
  std::chrono::steady_clock::time_point t0;
  auto frame_of = [&](int t, int k) { const int period = 2 * (nframes - 1); int i = (3 * t + k) % period; return i < nframes ? i : period - i; };
  auto work = [&](int t) {
    uint32_t d[3] = {(uint32_t)dims[0], (uint32_t)dims[1], (uint32_t)dims[2]};
    CoutSilencer* quiet = nullptr; (void)quiet;
    Matcher* m = nullptr; MonoProbe* vo = nullptr;
    if (workload == 2) vo = new MonoProbe(to_ref(mp)); else m = new Matcher(to_ref(&mp->match));
    int done = 0;
    auto step = [&](int k) {
      const int f = frame_of(t, k);
      uint8_t* I1 = imgs + (size_t)f * stride;
      uint8_t* I2 = (workload == 1 && imgs2) ? imgs2 + (size_t)f * stride : 0;
      if (vo) { vo->process(I1, d, false); return; }
      m->pushBack(I1, I2, d, false);
      if (k == 0) return;
      m->matchFeatures(workload == 1 ? 2 : 0, 0);
      if (bucket_max > 0) m->bucketFeatures(bucket_max, bw, bh);
      std::vector<Matcher::p_match> r = m->getMatches();
      (void)r;
    };
    for (int k = 0; k <= warm_pairs; k++) step(k);
    ready.fetch_add(1);
    while (go.load() == 0) std::this_thread::yield();
    for (int k = warm_pairs + 1; k <= warm_pairs + timed_pairs; k++) { step(k); done++; }
    finish[t] = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (pairs_done) pairs_done[t] = done;
    delete m; delete vo;
  };
  CoutSilencer quiet;                                    // "Estimate F time" etc. of estimateMotion
  std::vector<std::thread> pool;
  for (int t = 0; t < nthreads; t++) pool.emplace_back(work, t);
  while (ready.load() < nthreads) std::this_thread::yield();
  t0 = std::chrono::steady_clock::now();
  go.store(1);
  for (std::thread& th : pool) th.join();
  double wall = 0;
  for (double f : finish) wall = std::max(wall, f);
  return wall;
}
