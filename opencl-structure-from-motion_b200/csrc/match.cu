// SAD circle matching and pixel refinement.
//
// What it reproduces (reference paths relative to /root/reference/viso):
//   Matcher::findMatch ................. matcher.cpp:892-963   (32-byte SAD, simd.hh:426-432)
//   Matcher::matching, method 0 and 2 .. matcher.cpp:965-1041, 1088-1153
//   Matcher::relocateMinimum ........... matcher.cpp:1456-1496 (+ computeSmallDescriptor 479-506)
//   Matcher::refinement (pixel mode) ... matcher.cpp:1498-1585
//
// Design: the reference walks candidate bins in a fixed order and keeps the first minimum.  Here a group of
// 8 lanes scans the candidates of one query in parallel and reduces complete keys
//   (SAD cost, candidate u_bin, candidate v_bin, candidate index)
// whose minimum is exactly the candidate the reference's sequential loop would have kept, so the bin lists
// need no internal order.  All hops of a circle run inside one kernel because each hop's query is the previous
// hop's answer.  Accepted circles are compacted in ascending query order with a two-kernel scan; the flow
// "pixel already matched" rule (matcher.cpp:1036-1039) only ever involves the up-to-three preceding records
// because two features share a pixel only when they come from the same NMS cell.
#include "visocu_internal.cuh"
#include "outliers.cuh"
#include <mutex>
#include <utility>
#include <cmath>
#include <cstring>

namespace {

constexpr int G = 8;                 // lanes per query
constexpr int MATCH_THREADS = 256;   // 32 queries per block
constexpr int CHUNK = 1024;          // queries per compaction block

struct SetDev { const int32_t* rec; const int32_t* bin_start; const int2* bin_ent; int n; int pad; };
struct MatchJob {
  SetDev s[4];                       // 0 = 1p, 1 = 2p, 2 = 1c, 3 = 2c
  const visocu_range* ranges;
  int4* res;                         // per query: the other three indices of the circle + accepted flag
  int32_t* blk;                      // accepted circles per CHUNK
  visocu_pmatch* out;
  uint8_t* keep;                     // sub-pixel refinement: 0 = match dropped
  int32_t* n_out;
  double tr[12];                     // rows 0..2 of Tr_delta (previous -> current), used if has_tr
  double f, cu, cv, base;            // calibration for the prediction (Matcher::parameters)
  int has_tr, pad2;
  const uint8_t* du[4];              // planes used by the refinement (full resolution)
  const uint8_t* dv[4];
  int nq, dyn;                       // dyn: nq is only the capacity, the real count comes from cnt[] (device memory)
  const int32_t* cnt[4];             // record counters of the four sets for this pass (null = set not used)
  const int32_t* gate;               // second pass of a fused call: status word of the first pass's outlier removal.  Non-zero
                                     // (list declined: the host votes, computes the ranges and repeats this pass) = no queries
};

__device__ __forceinline__ int bin_index(const Geometry& g, int u, int v) {
  const float bs = (float)g.binsize;
  int ub = min((int)floorf((float)u / bs), g.ub - 1);
  int vb = min((int)floorf((float)v / bs), g.vb - 1);
  return vb * g.ub + ub;
}

// number of queries of a job.  Lazy mode (the host has not read the record counts back): taken from the frames' counters,
// with the reference's rule that nothing is matched if a needed set is empty (matcher.cpp:190-212).
__device__ __forceinline__ int job_nq(const MatchJob& J, int method) {
  if (!J.dyn) return J.nq;
  if (J.gate && *J.gate != 0) return 0;
  const int a = J.cnt[0] ? *J.cnt[0] : 1, b = J.cnt[1] ? *J.cnt[1] : 1, c = J.cnt[2] ? *J.cnt[2] : 1, d = J.cnt[3] ? *J.cnt[3] : 1;
  if (a == 0 || b == 0 || c == 0 || d == 0) return 0;
  return min(method == 2 ? a : c, J.nq);
}

// one hop: best candidate in set B for feature i1 of set A (all G lanes of the group call this together)
__device__ __forceinline__ int find_match(const Geometry& g, const SetDev& A, int i1, const SetDev& B, int stat_bin, int stage,
                                          bool flow, bool use_prior, const visocu_range* __restrict__ ranges, bool active, int sub,
                                          unsigned& n_cand, unsigned& n_scan) {
  unsigned long long best = ~0ull;
  if (active) {
    const int32_t* q = A.rec + (size_t)i1 * 12;
    const int4 hdr = *(const int4*)q;
    const uint4 qa = *(const uint4*)(q + 4), qb = *(const uint4*)(q + 8);
    const int u1 = hdr.x, v1 = hdr.y, c = hdr.w;
    float u_min, u_max, v_min, v_max;
    if (use_prior) {
      const visocu_range* r = ranges + stat_bin;
      u_min = (float)u1 + r->u_min[stage]; u_max = (float)u1 + r->u_max[stage];
      v_min = (float)v1 + r->v_min[stage]; v_max = (float)v1 + r->v_max[stage];
    } else {
      u_min = (float)(u1 - g.radius); u_max = (float)(u1 + g.radius);
      v_min = (float)(v1 - g.radius); v_max = (float)(v1 + g.radius);
    }
    if (!flow) { v_min = (float)(v1 - g.disp_tol); v_max = (float)(v1 + g.disp_tol); }
    const float bs = (float)g.binsize;
    const int ubmin = min(max((int)floorf(u_min / bs), 0), g.ub - 1), ubmax = min(max((int)floorf(u_max / bs), 0), g.ub - 1);
    const int vbmin = min(max((int)floorf(v_min / bs), 0), g.vb - 1), vbmax = min(max((int)floorf(v_max / bs), 0), g.vb - 1);
    // The window test of matcher.cpp:943 compares integer positions with float bounds: for an integer p, (float)p >= b is
    // p >= ceil(b) and (float)p <= b is p <= floor(b), so the bounds are converted once instead of every scanned position
    // (positions lie in [0, 16382], visocu_configure: bounds outside [-1, 65534] change nothing, and the 0xFFFF of a padding
    // entry never passes; a NaN bound admits nothing, as it does there)
    int iu_min = (int)fmaxf(ceilf(u_min), -1.f), iu_max = (int)fminf(floorf(u_max), 65534.f);
    int iv_min = (int)fmaxf(ceilf(v_min), -1.f), iv_max = (int)fminf(floorf(v_max), 65534.f);
    if (u_min != u_min || u_max != u_max || v_min != v_min || v_max != v_max) { iu_min = 1; iu_max = 0; }
    for (int vbin = vbmin; vbin <= vbmax; vbin++) {
      const int row = (c * g.vb + vbin) * g.ub;
      const int e0 = B.bin_start[row + ubmin], e1 = ubmax >= ubmin ? B.bin_start[row + ubmax + 1] : e0;
      // Most scanned entries fail the window test, so the scan is a chain of entry loads: four entries per lane are in
      // flight at a time (a position of 0xFFFF, 0xFFFF stands for "past the end": it fails every window)
      for (int e = e0 + sub; e < e1; e += 4 * G) {
        int2 ent4[4];
#pragma unroll
        for (int k = 0; k < 4; k++) ent4[k] = e + k * G < e1 ? B.bin_ent[e + k * G] : make_int2(-1, 0);
#pragma unroll
        for (int k = 0; k < 4; k++) {
          const int2 ent = ent4[k];
          const int u2 = ent.x & 0xFFFF, v2 = (int)((unsigned)ent.x >> 16);
          n_scan += e + k * G < e1 ? 1 : 0;
          if (u2 >= iu_min && u2 <= iu_max && v2 >= iv_min && v2 <= iv_max) {
            const int32_t* t = B.rec + (size_t)ent.y * 12;
            const uint4 ta = *(const uint4*)(t + 4), tb = *(const uint4*)(t + 8);
            unsigned sad = __vsadu4(qa.x, ta.x) + __vsadu4(qa.y, ta.y) + __vsadu4(qa.z, ta.z) + __vsadu4(qa.w, ta.w) +
                           __vsadu4(qb.x, tb.x) + __vsadu4(qb.y, tb.y) + __vsadu4(qb.z, tb.z) + __vsadu4(qb.w, tb.w);
            const int ub2 = min((int)floorf((float)u2 / bs), g.ub - 1);
            unsigned long long key = ((unsigned long long)sad << 48) | ((unsigned long long)ub2 << 36) |
                                     ((unsigned long long)vbin << 24) | (unsigned long long)ent.y;
            best = key < best ? key : best;
            n_cand++;
          }
        }
      }
    }
  }
  __syncwarp();
#pragma unroll
  for (int o = G / 2; o; o >>= 1) {
    unsigned long long other = __shfl_xor_sync(0xFFFFFFFFu, best, o);
    best = other < best ? other : best;
  }
  return best == ~0ull ? 0 : (int)(best & 0xFFFFFFull);    // min_ind defaults to 0 (matcher.cpp:898)
}

// the same hop with a predicted position (up, vp): cost = SAD + 4 * distance to the prediction, in double precision
// (matcher.cpp:948-953).  The arithmetic uses explicit IEEE operations (no FMA contraction), so it equals the reference
// built with -ffp-contract=off bit for bit.  Keys: (cost as ordered double bits, then u_bin, v_bin, index).
__device__ __forceinline__ int find_match_predicted(const Geometry& g, const SetDev& A, int i1, const SetDev& B, int stat_bin, int stage,
                                                    bool flow, bool use_prior, const visocu_range* __restrict__ ranges, bool active,
                                                    int sub, double up, double vp, unsigned& n_cand, unsigned& n_scan) {
  unsigned long long best_cost = ~0ull, best_ord = ~0ull;
  if (active) {
    const int32_t* q = A.rec + (size_t)i1 * 12;
    const int4 hdr = *(const int4*)q;
    const uint4 qa = *(const uint4*)(q + 4), qb = *(const uint4*)(q + 8);
    const int u1 = hdr.x, v1 = hdr.y, c = hdr.w;
    float u_min, u_max, v_min, v_max;
    if (use_prior) {
      const visocu_range* r = ranges + stat_bin;
      u_min = (float)u1 + r->u_min[stage]; u_max = (float)u1 + r->u_max[stage];
      v_min = (float)v1 + r->v_min[stage]; v_max = (float)v1 + r->v_max[stage];
    } else {
      u_min = (float)(u1 - g.radius); u_max = (float)(u1 + g.radius);
      v_min = (float)(v1 - g.radius); v_max = (float)(v1 + g.radius);
    }
    if (!flow) { v_min = (float)(v1 - g.disp_tol); v_max = (float)(v1 + g.disp_tol); }
    const float bs = (float)g.binsize;
    const int ubmin = min(max((int)floorf(u_min / bs), 0), g.ub - 1), ubmax = min(max((int)floorf(u_max / bs), 0), g.ub - 1);
    const int vbmin = min(max((int)floorf(v_min / bs), 0), g.vb - 1), vbmax = min(max((int)floorf(v_max / bs), 0), g.vb - 1);
    const bool predicted = up >= 0 && vp >= 0;
    for (int vbin = vbmin; vbin <= vbmax; vbin++) {
      const int row = (c * g.vb + vbin) * g.ub;
      const int e0 = B.bin_start[row + ubmin], e1 = ubmax >= ubmin ? B.bin_start[row + ubmax + 1] : e0;
      for (int e = e0 + sub; e < e1; e += G) {
        const int2 ent = B.bin_ent[e];
        const int u2 = ent.x & 0xFFFF, v2 = (int)((unsigned)ent.x >> 16);
        n_scan++;
        if ((float)u2 >= u_min && (float)u2 <= u_max && (float)v2 >= v_min && (float)v2 <= v_max) {
          const int32_t* t = B.rec + (size_t)ent.y * 12;
          const uint4 ta = *(const uint4*)(t + 4), tb = *(const uint4*)(t + 8);
          unsigned sad = __vsadu4(qa.x, ta.x) + __vsadu4(qa.y, ta.y) + __vsadu4(qa.z, ta.z) + __vsadu4(qa.w, ta.w) +
                         __vsadu4(qb.x, tb.x) + __vsadu4(qb.y, tb.y) + __vsadu4(qb.z, tb.z) + __vsadu4(qb.w, tb.w);
          double cost = (double)sad;
          if (predicted) {
            const double du = __dsub_rn((double)u2, up), dv = __dsub_rn((double)v2, vp);
            const double dist = __dsqrt_rn(__dadd_rn(__dmul_rn(du, du), __dmul_rn(dv, dv)));
            cost = __dadd_rn(cost, __dmul_rn(4.0, dist));
          }
          n_cand++;
          // the reference starts from min_cost = 1e7 and keeps a candidate only if cost < min_cost (matcher.cpp:899,955):
          // NaN, +inf and costs >= 1e7 (a degenerate projection: z2c near 0) never win and min_ind stays 0
          if (!(cost < 10000000.0)) continue;
          const int ub2 = min((int)floorf((float)u2 / bs), g.ub - 1);
          const unsigned long long cb = (unsigned long long)__double_as_longlong(cost);      // cost >= 0: bit order = value order
          const unsigned long long ord = ((unsigned long long)ub2 << 36) | ((unsigned long long)vbin << 24) | (unsigned long long)ent.y;
          if (cb < best_cost || (cb == best_cost && ord < best_ord)) { best_cost = cb; best_ord = ord; }
        }
      }
    }
  }
  __syncwarp();
#pragma unroll
  for (int o = G / 2; o; o >>= 1) {
    const unsigned long long oc = __shfl_xor_sync(0xFFFFFFFFu, best_cost, o), oo = __shfl_xor_sync(0xFFFFFFFFu, best_ord, o);
    if (oc < best_cost || (oc == best_cost && oo < best_ord)) { best_cost = oc; best_ord = oo; }
  }
  return best_cost == ~0ull ? 0 : (int)(best_ord & 0xFFFFFFull);
}

// one instantiation per matching method: flow and stereo matching stay clear of the registers the double-precision
// prediction of quad matching needs
template <int method>
__global__ void __launch_bounds__(MATCH_THREADS, method == 2 ? 3 : 4) k_match(Geometry g, const MatchJob* jobs, int use_prior, uint64_t* stats) {
  const MatchJob& J = jobs[blockIdx.y];
  const int sub = threadIdx.x % G;
  const int nq = job_nq(J, method);
  unsigned n_cand = 0, n_scan = 0;
  // a block takes 32 queries at a time; the grid covers all of them unless it was sized by the capacity (lazy mode)
  for (int base = blockIdx.x * (MATCH_THREADS / G); base < nq; base += gridDim.x * (MATCH_THREADS / G)) {
  const int i = base + threadIdx.x / G;
  const bool active = i < nq;
  int4 res = make_int4(0, 0, 0, 0);
  if (method == 0) {
    int stat_bin = 0;
    if (active) { const int2 uv = *(const int2*)(J.s[2].rec + (size_t)i * 12); stat_bin = bin_index(g, uv.x, uv.y); }
    const int i1p = find_match(g, J.s[2], i, J.s[0], stat_bin, 0, true, use_prior, J.ranges, active, sub, n_cand, n_scan);
    const int i1c2 = find_match(g, J.s[0], i1p, J.s[2], stat_bin, 1, true, use_prior, J.ranges, active, sub, n_cand, n_scan);
    res = make_int4(i1p, 0, 0, i1c2 == i);
  } else if (method == 1) {
    // stereo (matcher.cpp:1045-1084): current left -> current right -> back, positive disparity
    int stat_bin = 0;
    if (active) { const int2 uv = *(const int2*)(J.s[2].rec + (size_t)i * 12); stat_bin = bin_index(g, uv.x, uv.y); }
    const int i2c = find_match(g, J.s[2], i, J.s[3], stat_bin, 0, false, use_prior, J.ranges, active, sub, n_cand, n_scan);
    const int i1c2 = find_match(g, J.s[3], i2c, J.s[2], stat_bin, 1, false, use_prior, J.ranges, active, sub, n_cand, n_scan);
    int ok = 0;
    if (active && i1c2 == i) ok = J.s[2].rec[(size_t)i * 12] >= J.s[3].rec[(size_t)i2c * 12];
    res = make_int4(i2c, 0, 0, ok);
  } else {
    int stat_bin = 0;
    if (active) { const int2 uv = *(const int2*)(J.s[0].rec + (size_t)i * 12); stat_bin = bin_index(g, uv.x, uv.y); }
    const int i2p = find_match(g, J.s[0], i, J.s[1], stat_bin, 0, false, use_prior, J.ranges, active, sub, n_cand, n_scan);
    int i2c, i1c, i1p2;
    if (J.has_tr) {
      // motion-predicted search (matcher.cpp:1112-1138): the previous stereo match is triangulated, moved by Tr_delta and
      // projected into the current right image; candidates are penalised by their distance to that prediction.
      double up = -1, vp = -1, u1pd = -1, v1pd = -1;
      if (active) {
        const int2 a = *(const int2*)(J.s[0].rec + (size_t)i * 12);
        const int u2p = J.s[1].rec[(size_t)i2p * 12];
        u1pd = (double)a.x; v1pd = (double)a.y;
        const double d = fmax(__dsub_rn(u1pd, (double)u2p), 1.0);
        const double x1p = __ddiv_rn(__dmul_rn(__dsub_rn(u1pd, J.cu), J.base), d);
        const double y1p = __ddiv_rn(__dmul_rn(__dsub_rn(v1pd, J.cv), J.base), d);
        const double z1p = __ddiv_rn(__dmul_rn(J.f, J.base), d);
        const double* t = J.tr;
        const double x2c = __dsub_rn(__dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(t[0], x1p), __dmul_rn(t[1], y1p)), __dmul_rn(t[2], z1p)), t[3]), J.base);
        const double y2c = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(t[4], x1p), __dmul_rn(t[5], y1p)), __dmul_rn(t[6], z1p)), t[7]);
        const double z2c = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(t[8], x1p), __dmul_rn(t[9], y1p)), __dmul_rn(t[10], z1p)), t[11]);
        up = __dadd_rn(__ddiv_rn(__dmul_rn(J.f, x2c), z2c), J.cu);
        vp = __dadd_rn(__ddiv_rn(__dmul_rn(J.f, y2c), z2c), J.cv);
      }
      i2c = find_match_predicted(g, J.s[1], i2p, J.s[3], stat_bin, 1, true, use_prior, J.ranges, active, sub, up, vp, n_cand, n_scan);
      i1c = find_match(g, J.s[3], i2c, J.s[2], stat_bin, 2, false, use_prior, J.ranges, active, sub, n_cand, n_scan);
      i1p2 = find_match_predicted(g, J.s[2], i1c, J.s[0], stat_bin, 3, true, use_prior, J.ranges, active, sub, u1pd, v1pd, n_cand, n_scan);
    } else {
      i2c = find_match(g, J.s[1], i2p, J.s[3], stat_bin, 1, true, use_prior, J.ranges, active, sub, n_cand, n_scan);
      i1c = find_match(g, J.s[3], i2c, J.s[2], stat_bin, 2, false, use_prior, J.ranges, active, sub, n_cand, n_scan);
      i1p2 = find_match(g, J.s[2], i1c, J.s[0], stat_bin, 3, true, use_prior, J.ranges, active, sub, n_cand, n_scan);
    }
    int ok = 0;
    if (active && i1p2 == i) {
      const int u1p = J.s[0].rec[(size_t)i * 12], u2p = J.s[1].rec[(size_t)i2p * 12];
      const int u1c = J.s[2].rec[(size_t)i1c * 12], u2c = J.s[3].rec[(size_t)i2c * 12];
      ok = (u1p >= u2p && u1c >= u2c);
    }
    res = make_int4(i2p, i2c, i1c, ok);
  }
  if (active && sub == 0) J.res[i] = res;
  }
  // work counters: one atomic per warp
  for (int o = 16; o; o >>= 1) { n_cand += __shfl_xor_sync(0xFFFFFFFFu, n_cand, o); n_scan += __shfl_xor_sync(0xFFFFFFFFu, n_scan, o); }
  if ((threadIdx.x & 31) == 0 && (n_cand | n_scan)) {
    atomicAdd((unsigned long long*)&stats[0], (unsigned long long)n_cand);
    atomicAdd((unsigned long long*)&stats[1], (unsigned long long)n_scan);
  }
}

// accepted-and-kept flag of query i (identical in the count and the emit kernel)
__device__ __forceinline__ int keep_flag(const MatchJob& J, int method, int i, int nq) {
  if (i >= nq) return 0;
  if (!J.res[i].w) return 0;
  if (method == 2) return 1;                       // flow and stereo keep one match per current-left pixel
  const int2 uv = *(const int2*)(J.s[2].rec + (size_t)i * 12);
  for (int j = i - 1; j >= 0 && j >= i - 3; j--) {
    if (!J.res[j].w) continue;
    const int2 o = *(const int2*)(J.s[2].rec + (size_t)j * 12);
    if (o.x == uv.x && o.y == uv.y) return 0;      // an earlier accepted circle already owns this pixel
  }
  return 1;
}

__global__ void __launch_bounds__(CHUNK) k_match_count(const MatchJob* jobs, int method) {
  const MatchJob& J = jobs[blockIdx.y];
  const int nq = job_nq(J, method);
  if ((int)blockIdx.x * CHUNK >= nq && blockIdx.x > 0) return;
  int f = keep_flag(J, method, blockIdx.x * CHUNK + threadIdx.x, nq);
  __shared__ int wsum[32];
  for (int o = 16; o; o >>= 1) f += __shfl_xor_sync(0xFFFFFFFFu, f, o);
  if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = f;
  __syncthreads();
  if (threadIdx.x < 32) {
    int v = wsum[threadIdx.x];
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
    if (threadIdx.x == 0) J.blk[blockIdx.x] = v;
  }
}

__global__ void __launch_bounds__(CHUNK) k_match_emit(const MatchJob* jobs, int method) {
  const MatchJob& J = jobs[blockIdx.y];
  const int nq = job_nq(J, method);
  const int nchunk = (nq + CHUNK - 1) / CHUNK;
  if ((int)blockIdx.x >= nchunk) {
    if (nchunk == 0 && blockIdx.x == 0 && threadIdx.x == 0) *J.n_out = 0;
    return;
  }
  __shared__ int wsum[32];
  __shared__ int s_base;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  int acc = 0;
  for (int b = tid; b < (int)blockIdx.x; b += CHUNK) acc += J.blk[b];
  for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xFFFFFFFFu, acc, o);
  if (lane == 0) wsum[wid] = acc;
  __syncthreads();
  if (tid < 32) {
    int v = wsum[tid];
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
    if (tid == 0) s_base = v;
  }
  __syncthreads();
  const int base = s_base;
  __syncthreads();
  const int i = blockIdx.x * CHUNK + tid;
  const int f = keep_flag(J, method, i, nq);
  int incl = f;
  for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xFFFFFFFFu, incl, o); if (lane >= o) incl += t; }
  if (lane == 31) wsum[wid] = incl;
  __syncthreads();
  if (tid < 32) {
    int v = wsum[tid], s = v;
    for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xFFFFFFFFu, s, o); if (lane >= o) s += t; }
    wsum[tid] = s - v;
    if (tid == 31 && (int)blockIdx.x == nchunk - 1) *J.n_out = base + s;
  }
  __syncthreads();
  if (!f) return;
  const int pos = base + wsum[wid] + incl - 1;
  const int4 r = J.res[i];
  visocu_pmatch m;
  if (method == 0) {
    const int2 c = *(const int2*)(J.s[2].rec + (size_t)i * 12);
    const int2 p = *(const int2*)(J.s[0].rec + (size_t)r.x * 12);
    m.u1p = (float)p.x; m.v1p = (float)p.y; m.i1p = r.x;
    m.u2p = -1.f; m.v2p = -1.f; m.i2p = -1;
    m.u1c = (float)c.x; m.v1c = (float)c.y; m.i1c = i;
    m.u2c = -1.f; m.v2c = -1.f; m.i2c = -1;
  } else if (method == 1) {
    const int2 c = *(const int2*)(J.s[2].rec + (size_t)i * 12);
    const int2 d = *(const int2*)(J.s[3].rec + (size_t)r.x * 12);
    m.u1p = -1.f; m.v1p = -1.f; m.i1p = -1;
    m.u2p = -1.f; m.v2p = -1.f; m.i2p = -1;
    m.u1c = (float)c.x; m.v1c = (float)c.y; m.i1c = i;
    m.u2c = (float)d.x; m.v2c = (float)d.y; m.i2c = r.x;
  } else {
    const int2 a = *(const int2*)(J.s[0].rec + (size_t)i * 12);
    const int2 b = *(const int2*)(J.s[1].rec + (size_t)r.x * 12);
    const int2 d = *(const int2*)(J.s[3].rec + (size_t)r.y * 12);
    const int2 c = *(const int2*)(J.s[2].rec + (size_t)r.z * 12);
    m.u1p = (float)a.x; m.v1p = (float)a.y; m.i1p = i;
    m.u2p = (float)b.x; m.v2p = (float)b.y; m.i2p = r.x;
    m.u1c = (float)c.x; m.v1c = (float)c.y; m.i1c = r.z;
    m.u2c = (float)d.x; m.v2c = (float)d.y; m.i2c = r.y;
  }
  J.out[pos] = m;
}

// ---------------------------------------------------------------------------------------------------------
// pixel refinement.  16-byte descriptor of computeSmallDescriptor (matcher.cpp:479-506): (plane, dx, dy)
__constant__ int8_t c_sd_plane[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1};
__constant__ int8_t c_sd_dx[16] = {0, -2, 0, 2, -1, 0, 0, 1, -2, 0, 2, 0, 0, -1, 1, 0};
__constant__ int8_t c_sd_dy[16] = {-2, -1, -1, -1, 0, 0, 0, 0, 1, 1, 1, 2, -1, 0, 0, 1};

// SAD of the 16-byte descriptors at a (image 1) and b (image 2); the pointers address the centre pixel in the du and dv
// planes.  The sample list is spelled out so that every load is [row pointer + immediate]: (plane, dx, dy) as in the
// tables above.
#define VISO_SD_LIST(X) X(0, 0, -2) X(0, -2, -1) X(0, 0, -1) X(0, 2, -1) X(0, -1, 0) X(0, 0, 0) X(0, 0, 0) X(0, 1, 0) \
                        X(0, -2, 1) X(0, 0, 1) X(0, 2, 1) X(0, 0, 2) X(1, 0, -1) X(1, -1, 0) X(1, 1, 0) X(1, 0, 1)
__device__ __forceinline__ int sd_cost(const uint8_t* __restrict__ au, const uint8_t* __restrict__ av, const uint8_t* __restrict__ bu,
                                       const uint8_t* __restrict__ bv, int bpl) {
  unsigned sad = 0;
#define X(P, DX, DY) sad = __usad((unsigned)(P ? av : au)[(DY) * bpl + (DX)], (unsigned)(P ? bv : bu)[(DY) * bpl + (DX)], sad);
  VISO_SD_LIST(X)
#undef X
  return (int)sad;
}

// one warp relocates (u2,v2) in image 2 to the best of the 5x5 positions around it; lane = candidate
__device__ __forceinline__ void relocate(const uint8_t* du1, const uint8_t* dv1, const uint8_t* du2, const uint8_t* dv2,
                                         int bpl, int w, int h, float u1, float v1, float& u2, float& v2, int lane) {
  if (u2 - 2 < VISO_MARGIN || u2 + 2 > w - 1 - VISO_MARGIN || v2 - 2 < VISO_MARGIN || v2 + 2 > h - 1 - VISO_MARGIN) return;
  const int iu1 = (int)u1, iv1 = (int)v1;
  int key = 0x7FFFFFFF;
  if (lane < 25) {
    const int cu = (int)u2 + lane % 5 - 2, cv = (int)v2 + lane / 5 - 2;
    const size_t o1 = (size_t)iv1 * bpl + iu1, o2 = (size_t)cv * bpl + cu;
    key = sd_cost(du1 + o1, dv1 + o1, du2 + o2, dv2 + o2, bpl) * 32 + lane;
  }
  key = __reduce_min_sync(0xFFFFFFFFu, key);
  const int best = key & 31;
  u2 += (float)(best % 5) - 2.0f;
  v2 += (float)(best / 5) - 2.0f;
}

// (AtA)^-1 At of the 9 x 6 design matrix [j^2, i^2, i*j, j, i, 1], i, j = -1..1 (matcher.cpp:1509-1521): the
// least-squares paraboloid through the 3x3 costs around the integer minimum is b = c_pinv * costs.
__constant__ double c_pinv[54];

// sub-pixel variant (parabolicFitting, matcher.cpp:1379-1454): 7x7 integer costs, first minimum, reject minima on the
// border of the search area, paraboloid fit, reject offsets of a pixel or more.  Returns false if the match is dropped.
__device__ __forceinline__ bool parabolic(const uint8_t* du1, const uint8_t* dv1, const uint8_t* du2, const uint8_t* dv2,
                                          int bpl, int w, int h, float u1, float v1, float& u2, float& v2, int lane) {
  if (u2 - 3 < VISO_MARGIN || u2 + 3 > w - 1 - VISO_MARGIN || v2 - 3 < VISO_MARGIN || v2 + 3 > h - 1 - VISO_MARGIN) return false;
  const int iu1 = (int)u1, iv1 = (int)v1, iu2 = (int)u2, iv2 = (int)v2;
  int cost[2];
#pragma unroll
  for (int r = 0; r < 2; r++) {
    const int idx = lane + 32 * r;
    int sad = 0x7FFFFF;
    if (idx < 49) {
      const int cu = iu2 + idx % 7 - 3, cv = iv2 + idx / 7 - 3;
      const size_t o1 = (size_t)iv1 * bpl + iu1, o2 = (size_t)cv * bpl + cu;
      sad = sd_cost(du1 + o1, dv1 + o1, du2 + o2, dv2 + o2, bpl);
    }
    cost[r] = sad;
  }
  int key = min(cost[0] * 64 + lane, cost[1] * 64 + lane + 32);
  key = __reduce_min_sync(0xFFFFFFFFu, key);
  const int best = key & 63, bu = best % 7, bv = best / 7;
  if (bu == 0 || bu == 6 || bv == 0 || bv == 6) return false;
  double b[6] = {0, 0, 0, 0, 0, 0};
#pragma unroll
  for (int q = 0; q < 9; q++) {
    const int idx = (bv + q / 3 - 1) * 7 + (bu + q % 3 - 1);
    const double c = (double)__shfl_sync(0xFFFFFFFFu, idx < 32 ? cost[0] : cost[1], idx & 31);
#pragma unroll
    for (int r = 0; r < 6; r++) b[r] = fma(c_pinv[r * 9 + q], c, b[r]);
  }
  const float divisor = (float)(b[2] * b[2] - 4.0 * b[0] * b[1]);
  if (fabsf(divisor) < 1e-8 || fabs(b[2]) < 1e-8) return false;
  const float ddv = (float)((2.0 * b[0] * b[4] - b[2] * b[3]) / (double)divisor);
  const float ddu = (float)(-(b[4] + 2.0 * b[1] * (double)ddv) / b[2]);
  if (fabsf(ddu) >= 1.0f || fabsf(ddv) >= 1.0f) return false;
  u2 = (float)((double)u2 + ((double)(float)bu - 3.0 + (double)ddu));
  v2 = (float)((double)v2 + ((double)(float)bv - 3.0 + (double)ddv));
  return true;
}

// mode 1 = relocateMinimum (pixel), 2 = parabolicFitting (sub-pixel, may drop the match: keep[i] = 0)
__global__ void __launch_bounds__(256) k_refine(Geometry g, const MatchJob* jobs, int method, int mode, visocu_pmatch* direct,
                                                int n_direct, uint8_t* keep_direct) {
  const MatchJob& J = jobs[blockIdx.y];
  const int n = direct ? n_direct : *J.n_out;
  visocu_pmatch* list = direct ? direct : J.out;
  uint8_t* keep = direct ? keep_direct : J.keep;
  const int lane = threadIdx.x & 31;
  for (int i = blockIdx.x * 8 + (threadIdx.x >> 5); i < n; i += gridDim.x * 8) {
    visocu_pmatch m = list[i];
    bool ok = true;
    // the reference descriptor is always taken at (u1c,v1c) of the current left image (matcher.cpp:1544-1577);
    // hops: previous left (flow, quad), current right (stereo, quad), previous right (quad)
    if (method == 0 || method == 2) {
      if (mode == 2) ok = parabolic(J.du[2], J.dv[2], J.du[0], J.dv[0], g.bpl, g.w, g.h, m.u1c, m.v1c, m.u1p, m.v1p, lane);
      else relocate(J.du[2], J.dv[2], J.du[0], J.dv[0], g.bpl, g.w, g.h, m.u1c, m.v1c, m.u1p, m.v1p, lane);
    }
    if (ok && (method == 1 || method == 2)) {
      if (mode == 2) ok = parabolic(J.du[2], J.dv[2], J.du[3], J.dv[3], g.bpl, g.w, g.h, m.u1c, m.v1c, m.u2c, m.v2c, lane);
      else relocate(J.du[2], J.dv[2], J.du[3], J.dv[3], g.bpl, g.w, g.h, m.u1c, m.v1c, m.u2c, m.v2c, lane);
    }
    if (ok && method == 2) {
      if (mode == 2) ok = parabolic(J.du[2], J.dv[2], J.du[1], J.dv[1], g.bpl, g.w, g.h, m.u1c, m.v1c, m.u2p, m.v2p, lane);
      else relocate(J.du[2], J.dv[2], J.du[1], J.dv[1], g.bpl, g.w, g.h, m.u1c, m.v1c, m.u2p, m.v2p, lane);
    }
    if (lane == 0) {
      list[i] = m;
      if (keep) keep[i] = ok ? 1 : 0;
    }
  }
}

// host: pseudo-inverse of the paraboloid design matrix by Gauss-Jordan elimination with partial pivoting
void make_pinv(double* out54) {
  double A[9][6], M[6][12];
  for (int q = 0; q < 9; q++) {
    const double i = q / 3 - 1, j = q % 3 - 1;
    const double row[6] = {j * j, i * i, i * j, j, i, 1};
    for (int c = 0; c < 6; c++) A[q][c] = row[c];
  }
  for (int r = 0; r < 6; r++)
    for (int c = 0; c < 6; c++) {
      double s = 0;
      for (int q = 0; q < 9; q++) s += A[q][r] * A[q][c];
      M[r][c] = s; M[r][6 + c] = r == c;
    }
  for (int c = 0; c < 6; c++) {
    int piv = c;
    for (int r = c + 1; r < 6; r++) if (fabs(M[r][c]) > fabs(M[piv][c])) piv = r;
    for (int k = 0; k < 12; k++) { double t = M[c][k]; M[c][k] = M[piv][k]; M[piv][k] = t; }
    const double d = M[c][c];
    for (int k = 0; k < 12; k++) M[c][k] /= d;
    for (int r = 0; r < 6; r++) {
      if (r == c) continue;
      const double f = M[r][c];
      for (int k = 0; k < 12; k++) M[r][k] -= f * M[c][k];
    }
  }
  for (int r = 0; r < 6; r++)
    for (int q = 0; q < 9; q++) {
      double s = 0;
      for (int c = 0; c < 6; c++) s += M[r][6 + c] * A[q][c];
      out54[r * 9 + q] = s;
    }
}

int upload_pinv(visocu_ctx* ctx) {
  if (ctx->pinv_ready) return VISOCU_OK;
  double p[54];
  make_pinv(p);
  CU_TRY(ctx, cudaMemcpyToSymbolAsync(c_pinv, p, sizeof p, 0, cudaMemcpyHostToDevice, ctx->stream));
  CU_TRY(ctx, visocu_stream_wait(ctx));
  ctx->pinv_ready = 1;
  return VISOCU_OK;
}

int fill_job(visocu_ctx* ctx, const visocu_quad& q, int method, int pass, MatchJob& J, bool dyn = false) {
  const int ids[4] = {q.f1p, q.f2p, q.f1c, q.f2c};
  const bool need[4] = {method != 1, method == 2, true, method != 0};
  memset(&J, 0, sizeof J);
  for (int k = 0; k < 4; k++) {
    if (!need[k]) continue;
    int f = ids[k];
    if (f < 0 || f >= ctx->n_frames) return visocu_set_error(ctx, VISOCU_EINVAL, "match job references frame %d", f);
    if (!ctx->frame_valid[f]) return visocu_set_error(ctx, VISOCU_ESTATE, "frame %d holds no features", f);
    const FrameDev& F = ctx->frames_h[f];
    J.s[k].rec = F.rec[pass]; J.s[k].bin_start = F.bin_start[pass]; J.s[k].bin_ent = F.bin_ent[pass];
    J.s[k].n = dyn ? ctx->g.cap[pass] : ctx->h_counts[2 * (size_t)f + pass];
    if (dyn) J.cnt[k] = F.counts + pass;
    J.du[k] = ctx->g.half ? F.du_full : F.du;
    J.dv[k] = ctx->g.half ? F.dv_full : F.dv;
  }
  return VISOCU_OK;
}

}  // namespace

// the second pass of a fused call works in a second set of scratch and pinned staging memory: the first set still holds
// the first pass's lists and job descriptors
struct ScratchSwap {
  visocu_ctx* ctx; bool on;
  ScratchSwap(visocu_ctx* c, bool enable) : ctx(c), on(enable) { swap(); }
  ~ScratchSwap() { swap(); }
  void swap() {
    if (!on) return;
    std::swap(ctx->scratch, ctx->scratch2); std::swap(ctx->scratch_bytes, ctx->scratch2_bytes);
    std::swap(ctx->pinned, ctx->pinned2); std::swap(ctx->pinned_bytes, ctx->pinned2_bytes);
  }
};

// mode 0: complete call.  2 / 3: first / second pass of a fused call (visocu_match_fused): everything enqueued on the main stream, nothing
// waited for, state in ctx->part[]; the second pass works in the second scratch set and takes its prior ranges from
// device memory (dev_ranges, one block of dev_ranges_stride bytes per job).
static int match_impl(visocu_ctx* ctx, int32_t n_jobs, const visocu_quad* jobs, int32_t method, int32_t pass,
                      int32_t use_prior, const visocu_range* const* ranges, const double* const* tr_delta, int32_t refine,
                      visocu_pmatch* const* out, const int32_t* cap, int32_t* n_out, int32_t* outliers, int mode,
                      const uint8_t* dev_ranges = nullptr, size_t dev_ranges_stride = 0, bool dyn = false) {
  const bool deferred = mode != 0;                   // no result is delivered by this call
  const bool second_set = mode == 3;
  if (!ctx) return VISOCU_EINVAL;
  if (!ctx->configured) return visocu_set_error(ctx, VISOCU_ESTATE, "context not configured");
  if (n_jobs <= 0 || !jobs || (!deferred && (!out || !cap || !n_out))) return visocu_set_error(ctx, VISOCU_EINVAL, "bad match arguments");
  if (deferred && (method != 0 || n_jobs > VISO_MAX_BATCH || refine == 2))
    return visocu_set_error(ctx, VISOCU_EINVAL, "deferred matching: flow method, pixel refinement, at most %d jobs", VISO_MAX_BATCH);
  ScratchSwap swap_guard(ctx, second_set);
  if (method < 0 || method > 2) return visocu_set_error(ctx, VISOCU_EINVAL, "method %d not supported (0 = flow, 1 = stereo, 2 = quad)", method);
  if (refine < 0 || refine > 2) return visocu_set_error(ctx, VISOCU_EINVAL, "refine must be 0, 1 or 2");
  if (refine == 2) { int rc0 = upload_pinv(ctx); if (rc0) return rc0; }
  if (pass < ctx->g.first_pass || pass > 1) return visocu_set_error(ctx, VISOCU_EINVAL, "pass %d not available", pass);
  if (use_prior && !ranges && !dev_ranges) return visocu_set_error(ctx, VISOCU_EINVAL, "use_prior needs ranges");
  CU_TRY(ctx, cudaSetDevice(ctx->device));
  const Geometry& g = ctx->g;
  const int nstat = g.ub * g.vb;
  if (!dyn) {
    // record counts the host has not read back yet (frames pushed without count pointers): fetch them now
    std::vector<int32_t> unknown;
    for (int j = 0; j < n_jobs; j++) {
      const int ids[4] = {jobs[j].f1p, jobs[j].f2p, jobs[j].f1c, jobs[j].f2c};
      for (int k = 0; k < 4; k++)
        if (ids[k] >= 0 && ids[k] < ctx->n_frames && ctx->h_counts[2 * (size_t)ids[k]] < 0) unknown.push_back(ids[k]);
    }
    if (!unknown.empty()) { int rc0 = visocu_frame_counts(ctx, (int32_t)unknown.size(), unknown.data(), nullptr, nullptr); if (rc0) return rc0; }
  }
  for (int start = 0; start < n_jobs; start += VISO_MAX_BATCH) {
    const int nb = n_jobs - start < VISO_MAX_BATCH ? n_jobs - start : VISO_MAX_BATCH;
    std::vector<MatchJob> hj(nb);
    const bool ro = outliers != nullptr || deferred;    // outlier removal on the device
    std::vector<RoJob> rj(ro ? nb : 0);
    int maxq = 0;
    for (int j = 0; j < nb; j++) {
      int rc = fill_job(ctx, jobs[start + j], method, pass, hj[j], dyn);
      if (rc) return rc;
      bool empty = false;
      const bool need[4] = {method != 1, method == 2, true, method != 0};
      for (int k = 0; k < 4; k++) if (need[k] && hj[j].s[k].n == 0) empty = true;
      const int nq = empty ? 0 : (method == 2 ? hj[j].s[0].n : hj[j].s[2].n);
      hj[j].nq = nq; hj[j].dyn = dyn ? 1 : 0;
      if (mode == 3 && dyn && n_jobs <= VISO_MAX_BATCH && ctx->part[0].pending) hj[j].gate = ctx->part[0].dev_words + 16 * j + 1;
      if (nq > maxq) maxq = nq;
      if (use_prior && !dev_ranges && !ranges[start + j]) return visocu_set_error(ctx, VISOCU_EINVAL, "job %d has no ranges", start + j);
    }
    // Device scratch.  A header block (job descriptors of both kernels and the prior ranges of every job) goes up in ONE
    // copy; the per-job result words come back in ONE copy; the match lists have a uniform stride so that ONE 2-D copy
    // into pinned memory fetches them all.  (Every API call costs microseconds under the driver's lock, and with many
    // worker threads sharing a GPU that lock is what bounds the throughput.)
    const size_t rb = (use_prior && !dev_ranges) ? align_up((size_t)nstat * sizeof(visocu_range), 256) : 0;
    const size_t h_mj = 0, h_rj = align_up(sizeof(MatchJob) * nb, 256);
    const size_t h_rng = h_rj + (ro ? align_up(sizeof(RoJob) * nb, 256) : 0);
    const size_t hdr_bytes = h_rng + rb * nb;
    const size_t o_words = hdr_bytes, words_bytes = (size_t)nb * 64;                  // 16 int32 per job
    const size_t ostride = align_up((size_t)(maxq + 1) * 48, 256);
    const size_t o_list = align_up(o_words + words_bytes, 256);                      // lists of the matching kernels
    const size_t o_list2 = o_list + ostride * nb;                                    // survivors of the outlier removal
    size_t off = o_list2 + (ro ? ostride * nb : 0);
    std::vector<size_t> o_res(nb), o_blk(nb), o_keep(nb), o_idx(nb), o_vert(nb), o_hnd(nb), o_rep(nb);
    const bool dedupe = ro && method != 0;       // flow matching keeps one match per pixel (matcher.cpp:1036-1039)
    const size_t kstride = (size_t)maxq + 1;             // words per job in the position buffer
    size_t o_keys = 0;
    for (int j = 0; j < nb; j++) {
      const int nq = hj[j].nq;
      o_res[j] = off; off += align_up((size_t)(nq + 1) * 16, 256);
      o_blk[j] = off; off += align_up((size_t)(nq / CHUNK + 2) * 4, 256);
      o_keep[j] = off; off += align_up((size_t)nq + 1, 256);
      if (ro) {
        o_idx[j] = off; off += align_up((size_t)(nq + 1) * 4, 256);
        o_vert[j] = off; off += align_up((size_t)(nq + 1) * 4, 256);
        o_hnd[j] = off; off += align_up((size_t)(nq / 2 + 2) * 8, 256);
        o_rep[j] = off; off += align_up((size_t)nq + 1, 256);
      }
    }
    if (dedupe) { o_keys = off; off += align_up(kstride * 4 * nb, 256); }
    int rc = visocu_ensure_scratch(ctx, off);
    if (rc) return rc;
    const bool stage_lists = !dyn && ostride * nb <= ((size_t)64 << 20);             // else copy list by list to the caller (lazy mode: the deliver kernel)
    const size_t p_words = align_up(hdr_bytes, 256), p_lists = p_words + align_up(words_bytes, 256);
    const size_t p_keys = p_lists + (stage_lists ? ostride * nb : 0);
    if ((rc = visocu_ensure_pinned(ctx, p_keys + (dedupe ? kstride * 5 * nb : 0)))) return rc;
    uint8_t* sb = (uint8_t*)ctx->scratch;
    uint8_t* pin = (uint8_t*)ctx->pinned;
    for (int j = 0; j < nb; j++) {
      int32_t* words = (int32_t*)(sb + o_words) + 16 * j;
      hj[j].res = (int4*)(sb + o_res[j]); hj[j].blk = (int32_t*)(sb + o_blk[j]);
      hj[j].out = (visocu_pmatch*)(sb + o_list + ostride * j); hj[j].n_out = words + 8;
      hj[j].keep = sb + o_keep[j];
      hj[j].f = ctx->param.f; hj[j].cu = ctx->param.cu; hj[j].cv = ctx->param.cv; hj[j].base = ctx->param.base;
      hj[j].has_tr = (method == 2 && tr_delta && tr_delta[start + j]) ? 1 : 0;
      if (hj[j].has_tr) memcpy(hj[j].tr, tr_delta[start + j], sizeof hj[j].tr);
      if (use_prior && dev_ranges) {
        hj[j].ranges = (const visocu_range*)(dev_ranges + dev_ranges_stride * j);
      } else if (use_prior) {
        memcpy(pin + h_rng + rb * j, ranges[start + j], (size_t)nstat * sizeof(visocu_range));
        hj[j].ranges = (const visocu_range*)(sb + h_rng + rb * j);
      }
      if (ro) {
        rj[j].in = hj[j].out; rj[j].keep_in = (refine == 2 && maxq > 0) ? hj[j].keep : nullptr; rj[j].n_in = hj[j].n_out;
        rj[j].out = (visocu_pmatch*)(sb + o_list2 + ostride * j);
        rj[j].result = words; rj[j].idx = (int32_t*)(sb + o_idx[j]); rj[j].vert = (int32_t*)(sb + o_vert[j]);
        rj[j].hnd = (uint16_t*)(sb + o_hnd[j]); rj[j].rep = nullptr;
      }
    }
    memcpy(pin + h_mj, hj.data(), sizeof(MatchJob) * nb);
    if (ro) memcpy(pin + h_rj, rj.data(), sizeof(RoJob) * nb);
    CU_COPY(ctx, sb, pin, hdr_bytes, cudaMemcpyHostToDevice);
    const MatchJob* dj = (const MatchJob*)(sb + h_mj);
    if (maxq > 0) {
      int gmx = (maxq + MATCH_THREADS / G - 1) / (MATCH_THREADS / G);
      if (dyn && gmx > 160) gmx = 160;                    // sized by the capacity: the blocks stride over the real queries
      dim3 gm(gmx, nb);
      if (method == 0) k_match<0><<<gm, MATCH_THREADS, 0, ctx->stream>>>(g, dj, use_prior, ctx->d_stats);
      else if (method == 1) k_match<1><<<gm, MATCH_THREADS, 0, ctx->stream>>>(g, dj, use_prior, ctx->d_stats);
      else k_match<2><<<gm, MATCH_THREADS, 0, ctx->stream>>>(g, dj, use_prior, ctx->d_stats);
      CU_LAUNCH_CHECK(ctx);
    }
    dim3 gc((maxq + CHUNK - 1) / CHUNK > 0 ? (maxq + CHUNK - 1) / CHUNK : 1, nb);
    k_match_count<<<gc, CHUNK, 0, ctx->stream>>>(dj, method);
    CU_LAUNCH_CHECK(ctx);
    k_match_emit<<<gc, CHUNK, 0, ctx->stream>>>(dj, method);
    CU_LAUNCH_CHECK(ctx);
    if (refine && maxq > 0) {
      int gx = (maxq + 7) / 8; if (gx > 4096) gx = 4096;
      if (dyn && gx > 592) gx = 592;                      // sized by the capacity: the warps stride over the real list
      dim3 gr(gx, nb);
      k_refine<<<gr, 256, 0, ctx->stream>>>(g, dj, method, refine, nullptr, 0, nullptr);
      CU_LAUNCH_CHECK(ctx);
    }
    // Matcher::removeOutliers on the device: one CTA per list, reads the list (and the sub-pixel keep flags) where the
    // kernels above left them and writes the survivors to the second list area
    if (dedupe && maxq > 0) {
      // Quad and stereo matching can put several matches on one pixel.  Which of them Triangle triangulates is decided
      // by replaying its quicksort on the host (ro_resolve_duplicates): positions down, flags up, one round trip.
      if ((rc = visocu_launch_outlier_keys(ctx, (const RoJob*)(sb + h_rj), nb, (uint32_t*)(sb + o_keys), (int)kstride))) return rc;
      uint32_t* pk = (uint32_t*)(pin + p_keys);
      uint8_t* prep = pin + p_keys + kstride * 4 * nb;
      int32_t* pw0 = (int32_t*)(pin + p_words);
      CU_COPY(ctx, pw0, sb + o_words, words_bytes, cudaMemcpyDeviceToHost);
      CU_COPY(ctx, pk, sb + o_keys, kstride * 4 * nb, cudaMemcpyDeviceToHost);
      CU_TRY(ctx, visocu_stream_wait(ctx));
      bool any = false;
      for (int j = 0; j < nb; j++) {
        const int n = pw0[16 * j + 8];
        if (n > 3 && n <= (int)kstride && ro_resolve_duplicates(pk + kstride * j, n, prep + kstride * j)) { rj[j].rep = sb + o_rep[j]; any = true; }
      }
      if (any) {
        for (int j = 0; j < nb; j++)
          if (rj[j].rep) CU_COPY(ctx, sb + o_rep[j], prep + kstride * j, (size_t)pw0[16 * j + 8], cudaMemcpyHostToDevice);
        memcpy(pin + h_rj, rj.data(), sizeof(RoJob) * nb);
        CU_COPY(ctx, sb + h_rj, pin + h_rj, sizeof(RoJob) * nb, cudaMemcpyHostToDevice);
      }
    }
    if (mode >= 2) {
      // one pass of a fused call: outlier removal and the read-back of the result words queued on the main stream
      // Shared memory of the outlier kernel follows the longest list.  Lazy mode: the host does not know the lengths when it
      // launches; the context keeps the longest list seen so far per pass plus a margin (ro_bound, -1 = nothing seen yet:
      // room for a complete feature list).  A longer list is declined by the kernel and voted on by the host, and the
      // bound grows.
      int max_list = maxq;
      if (dyn && ctx->ro_bound[pass] >= 0 && ctx->ro_bound[pass] < max_list) max_list = ctx->ro_bound[pass];
      if ((rc = visocu_launch_remove_outliers(ctx, (const RoJob*)(sb + h_rj), nb, method, max_list, ctx->stream))) return rc;
      visocu_deferred& st = ctx->part[mode - 2];
      st.pending = true; st.nb = nb; st.pin_words = pin + p_words; st.pin_lists = pin + p_lists;
      st.dev_lists = sb + o_list2; st.ostride = ostride; st.dev_words = (const int32_t*)(sb + o_words); st.dev_jobs = dj;
      return VISOCU_OK;
    }
    if (ro && (rc = visocu_launch_remove_outliers(ctx, (const RoJob*)(sb + h_rj), nb, method, maxq, ctx->stream))) return rc;
    int32_t* pw = (int32_t*)(pin + p_words);
    CU_COPY(ctx, pw, sb + o_words, words_bytes, cudaMemcpyDeviceToHost);
    CU_TRY(ctx, visocu_stream_wait(ctx));
    int maxn = 0;
    for (int j = 0; j < nb; j++) {
      const int n = outliers ? pw[16 * j] : pw[16 * j + 8];
      n_out[start + j] = n;
      if (n > maxn) maxn = n;
      if (n > cap[start + j]) return visocu_set_error(ctx, VISOCU_ECAPACITY, "job %d produced %d matches, room for %d", start + j, n, cap[start + j]);
      if (outliers) {
        const int status = pw[16 * j + 1];
        outliers[start + j] = status == 0 ? 1 : 0;
        if (status == 0 && n > 3) { for (int k = 0; k < 4; k++) ctx->ro_ns[k] += (uint64_t)pw[16 * j + 4 + k]; ctx->ro_jobs++; }
        else if (status != 0) { ctx->ro_declined++; ctx->ro_reason[status & 3]++; ctx->ro_declined_n += (uint64_t)pw[16 * j + 3]; }
      }
    }
    const uint8_t* lists = sb + (ro ? o_list2 : o_list);
    if (maxn > 0 && stage_lists) {
      const size_t wbytes = (size_t)maxn * 48;
      ctx->d2h_bytes += (uint64_t)wbytes * nb;
      CU_TRY(ctx, cudaMemcpy2DAsync(pin + p_lists, wbytes, lists, ostride, wbytes, nb, cudaMemcpyDeviceToHost, ctx->stream));
    } else if (maxn > 0) {
      for (int j = 0; j < nb; j++)
        if (n_out[start + j] > 0) CU_COPY(ctx, out[start + j], lists + ostride * j, (size_t)n_out[start + j] * 48, cudaMemcpyDeviceToHost);
    }
    const bool host_keep = refine == 2 && !outliers;       // the device outlier pass applies the keep flags itself
    std::vector<std::vector<uint8_t> > keep(host_keep ? nb : 0);
    for (int j = 0; j < (int)keep.size(); j++) {
      keep[j].resize((size_t)n_out[start + j] + 1);
      if (n_out[start + j] > 0) CU_COPY(ctx, keep[j].data(), hj[j].keep, (size_t)n_out[start + j], cudaMemcpyDeviceToHost);
    }
    if (maxn > 0) CU_TRY(ctx, visocu_stream_wait(ctx));
    if (maxn > 0 && stage_lists)
      for (int j = 0; j < nb; j++)
        if (n_out[start + j] > 0) memcpy(out[start + j], pin + p_lists + (size_t)maxn * 48 * j, (size_t)n_out[start + j] * 48);
    // sub-pixel refinement drops matches (matcher.cpp:1546-1577 `continue`): order-preserving compaction
    for (int j = 0; j < (int)keep.size(); j++) {
      visocu_pmatch* list = out[start + j];
      int kept = 0;
      for (int i = 0; i < n_out[start + j]; i++)
        if (keep[j][i]) list[kept++] = list[i];
      n_out[start + j] = kept;
    }
  }
  return VISOCU_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
// Matcher::computePriorStatistics (matcher.cpp:734-868) for the flow method, on the survivors of the first pass where
// the outlier kernel left them: per bin, minimum and maximum of the displacements of all matches in the 3x3 bin
// neighbourhood, widened to at least 20 pixels, or +-match_radius for bins nobody touched.  One CTA per job.  Minima
// and maxima of a set do not depend on the order, so float atomics (as ordered integers) reproduce the host loop exactly.
__device__ __forceinline__ void atomic_min_float(float* addr, float v) {
  if (v >= 0.f) atomicMin((int*)addr, __float_as_int(v)); else atomicMax((unsigned int*)addr, __float_as_uint(v));
}
__device__ __forceinline__ void atomic_max_float(float* addr, float v) {
  if (v >= 0.f) atomicMax((int*)addr, __float_as_int(v)); else atomicMin((unsigned int*)addr, __float_as_uint(v));
}

__global__ void __launch_bounds__(1024) k_prior_ranges(Geometry g, const uint8_t* lists, size_t ostride, const int32_t* words,
                                                      uint8_t* ranges, size_t rstride, float* tmp_global, int use_smem) {
  extern __shared__ float s_tmp[];
  const int j = blockIdx.x, tid = threadIdx.x;
  const int nbin = g.ub * g.vb;
  float* lo = use_smem ? s_tmp : tmp_global + (size_t)j * 9 * nbin;         // 4 values per bin each, then the seen flags
  float* hi = lo + 4 * nbin;
  float* seen = hi + 4 * nbin;
  for (int i = tid; i < 4 * nbin; i += 1024) { lo[i] = 1000000.f; hi[i] = -1000000.f; }
  for (int i = tid; i < nbin; i += 1024) seen[i] = 0.f;
  __syncthreads();
  const int n = words[16 * j + 1] == 0 ? words[16 * j] : 0;                 // a declined list is redone by the caller
  const visocu_pmatch* list = (const visocu_pmatch*)(lists + ostride * j);
  const float bs = (float)g.binsize;
  for (int i = tid; i < n; i += 1024) {
    const visocu_pmatch m = list[i];
    const float d[4] = {__fsub_rn(m.u1p, m.u1c), __fsub_rn(m.v1p, m.v1c), __fsub_rn(m.u1c, m.u1p), __fsub_rn(m.v1c, m.v1p)};
    const int cu = (int)floorf(__fdiv_rn(m.u1c, bs)), cv = (int)floorf(__fdiv_rn(m.v1c, bs));
    const int u0 = min(max(cu - 1, 0), g.ub - 1), u1 = min(max(cu + 1, 0), g.ub - 1);
    const int v0 = min(max(cv - 1, 0), g.vb - 1), v1 = min(max(cv + 1, 0), g.vb - 1);
    for (int v = v0; v <= v1; v++)
      for (int u = u0; u <= u1; u++) {
        const int b = v * g.ub + u;
        seen[b] = 1.f;
#pragma unroll
        for (int k = 0; k < 4; k++) { atomic_min_float(&lo[4 * b + k], d[k]); atomic_max_float(&hi[4 * b + k], d[k]); }
      }
  }
  __syncthreads();
  visocu_range* out = (visocu_range*)(ranges + rstride * j);
  const float radius = (float)g.radius;
  for (int b = tid; b < nbin; b += 1024) {
    visocu_range r;
    memset(&r, 0, sizeof r);
    const bool s = seen[b] != 0.f;
#pragma unroll
    for (int st = 0; st < 2; st++) {
      float dmin[2], dmax[2];
#pragma unroll
      for (int a = 0; a < 2; a++) {
        dmin[a] = s ? lo[4 * b + 2 * st + a] : -radius;
        dmax[a] = s ? hi[4 * b + 2 * st + a] : radius;
        const float span = __fsub_rn(dmax[a], dmin[a]);
        if (span < 20.f) {                                                   // matcher.cpp:845-854
          const float grow = ceilf(__fdiv_rn(__fsub_rn(20.f, span), 2.f));
          dmin[a] = __fsub_rn(dmin[a], grow); dmax[a] = __fadd_rn(dmax[a], grow);
        }
      }
      r.u_min[st] = dmin[0]; r.u_max[st] = dmax[0]; r.v_min[st] = dmin[1]; r.v_max[st] = dmax[1];
    }
    out[b] = r;
  }
}

extern "C" int visocu_match(visocu_ctx* ctx, int32_t n_jobs, const visocu_quad* jobs, int32_t method, int32_t pass,
                            int32_t use_prior, const visocu_range* const* ranges, const double* const* tr_delta, int32_t refine,
                            visocu_pmatch* const* out, const int32_t* cap, int32_t* n_out, int32_t* outliers) {
  return match_impl(ctx, n_jobs, jobs, method, pass, use_prior, ranges, tr_delta, refine, out, cap, n_out, outliers, 0);
}

// Results of a fused call, written by the GPU itself into mapped pinned host memory: per job a header (the 16 result words
// of each pass and the record counts of the two frames) and the two survivor lists.  One kernel, one CTA per job; the
// host needs no second round trip to learn how many records to copy.
struct DeliverArgs {
  const MatchJob* jobs2;                      // second-pass jobs: cnt[0] / cnt[2] point at the dense counters of f1p / f1c
  const int32_t* words1; const int32_t* words2;
  const uint8_t* lists1; const uint8_t* lists2;
  size_t ostride1, ostride2;
  uint8_t* dst; size_t dstride, off1, off2;
  int with_lists;
  int packed16;                               // second list as 12 bytes per match (see below)
};
__global__ void __launch_bounds__(256) k_deliver(DeliverArgs a) {
  const int j = blockIdx.x, tid = threadIdx.x;
  uint8_t* out = a.dst + a.dstride * j;
  int32_t* hdr = (int32_t*)out;
  if (tid < 16) { hdr[tid] = a.words1[16 * j + tid]; hdr[16 + tid] = a.words2[16 * j + tid]; }
  if (tid >= 32 && tid < 38) {
    // sparse, dense, overflow flag of the previous and of the current frame (FrameDev::counts)
    const int32_t* c = (tid < 35 ? a.jobs2[j].cnt[0] : a.jobs2[j].cnt[2]) - 1;
    hdr[tid] = c[(tid - 32) % 3];
  }
  if (!a.with_lists) return;
  const int n1 = a.words1[16 * j], n2 = a.words2[16 * j];
  if (a.with_lists & 2) {
    const uint4* s1 = (const uint4*)(a.lists1 + a.ostride1 * j); uint4* d1 = (uint4*)(out + a.off1);
    for (int i = tid; i < 3 * n1; i += 256) d1[i] = s1[i];
  }
  // Flow matches carry -1 in the six fields of the right images (matcher.cpp:1037): only the 24 bytes that say something
  // cross PCIe - (u1p, v1p, i1p) and (u1c, v1c, i1c) as six words per record - and the host puts the constants back.
  const uint32_t* s2 = (const uint32_t*)(a.lists2 + a.ostride2 * j); uint32_t* d2 = (uint32_t*)(out + a.off2);
  if (!a.packed16) {
    if (tid == 38) hdr[38] = 1;                 // format of list 2: 1 = six words per match
    for (int i = tid; i < 6 * n2; i += 256) { const int r = i / 6, w = i - 6 * r; d2[i] = s2[12 * r + (w < 3 ? w : w + 3)]; }
    return;
  }
  // Without sub-pixel refinement the coordinates of a match are whole pixels (matcher.cpp:1013-1036 copies the integer
  // feature positions, relocateMinimum moves them by whole pixels), and they and the feature indices fit 16 bits: three
  // words per match, (u1p | v1p << 16), (u1c | v1c << 16), (i1p | i1c << 16).  A value that would not survive is reported
  // (format 0x102: the host refuses the list instead of delivering something else than the reference).
  int bad = 0;
  for (int i = tid; i < 3 * n2; i += 256) {
    const int r = i / 3, w = i - 3 * r;
    uint32_t lo, hi;
    if (w < 2) {
      const float fu = __uint_as_float(s2[12 * r + 6 * w]), fv = __uint_as_float(s2[12 * r + 6 * w + 1]);
      const int u = (int)fu, v = (int)fv;
      if ((float)u != fu || (float)v != fv || (unsigned)u > 0xFFFFu || (unsigned)v > 0xFFFFu) bad = 1;
      lo = (uint32_t)u; hi = (uint32_t)v;
    } else {
      lo = s2[12 * r + 2]; hi = s2[12 * r + 8];
      if (lo > 0xFFFFu || hi > 0xFFFFu) bad = 1;
    }
    d2[i] = (lo & 0xFFFFu) | (hi << 16);
  }
  bad = __syncthreads_or(bad);
  if (tid == 38) hdr[38] = bad ? 0x102 : 2;     // format of list 2: 2 = three words per match
}

namespace {
struct FusedLayout { size_t rstride, off1, off2, dstride; bool zero_copy; };
FusedLayout fused_layout(const Geometry& g) {
  FusedLayout L;
  L.rstride = align_up((size_t)g.ub * g.vb * sizeof(visocu_range), 256);
  // delivery area: header + both lists at their capacity, per job.  Small geometries: the lists travel with the header
  // (zero-copy writes of the last kernel, one wait).  Large ones (4K: over a hundred megabytes of capacity per job): only
  // the header does, and the lists are copied once their lengths are known.
  L.off1 = 256; L.off2 = L.off1 + align_up((size_t)(g.cap[0] + 1) * 48, 256);
  const size_t full = L.off2 + align_up((size_t)(g.cap[1] + 1) * 48, 256);
  static const bool allow = [] { const char* e = getenv("VISOCU_ZEROCOPY"); return !(e && e[0] == '0'); }();   // 0: always copy the lists after the header
  L.zero_copy = allow && full <= ((size_t)8 << 20);
  L.dstride = L.zero_copy ? full : 256;
  return L;
}
}  // namespace

// everything of a fused call that is enqueued on the lane's stream (replayable as a graph: depends on the job list, the
// refinement mode, the ranges flag and the outlier bound only)
static int fused_enqueue(visocu_ctx* ctx, int32_t n_jobs, const visocu_quad* jobs, int32_t refine, bool want_ranges, bool want_list1) {
  const Geometry& g = ctx->g;
  const int nbin = g.ub * g.vb;
  const FusedLayout L = fused_layout(g);
  uint8_t* d_rng = (uint8_t*)ctx->d_ranges;
  // first pass (sparse features, no prior) and its outlier removal; the record counts are taken from device memory
  int rc = match_impl(ctx, n_jobs, jobs, 0, 0, 0, nullptr, nullptr, 0, nullptr, nullptr, nullptr, nullptr, 2, nullptr, 0, true);
  if (rc) return rc;
  // prior ranges from its survivors, on the device
  const visocu_deferred& A = ctx->part[0];
  const size_t smem = (size_t)9 * nbin * sizeof(float);
  const int use_smem = smem <= 40 * 1024 ? 1 : 0;
  k_prior_ranges<<<n_jobs, 1024, use_smem ? smem : 0, ctx->stream>>>(g, A.dev_lists, A.ostride, A.dev_words, d_rng, L.rstride,
                                                                   (float*)(d_rng + L.rstride * n_jobs), use_smem);
  CU_LAUNCH_CHECK(ctx);
  // second pass (dense features, the ranges as prior), refinement, outlier removal
  rc = match_impl(ctx, n_jobs, jobs, 0, 1, 1, nullptr, nullptr, refine, nullptr, nullptr, nullptr, nullptr, 3, d_rng, L.rstride, true);
  if (rc) return rc;
  const visocu_deferred& B = ctx->part[1];
  DeliverArgs da;
  da.jobs2 = (const MatchJob*)B.dev_jobs; da.words1 = A.dev_words; da.words2 = B.dev_words;
  da.lists1 = A.dev_lists; da.lists2 = B.dev_lists; da.ostride1 = A.ostride; da.ostride2 = B.ostride;
  da.dst = (uint8_t*)ctx->deliver_dev; da.dstride = L.dstride; da.off1 = L.off1; da.off2 = L.off2;
  da.with_lists = L.zero_copy ? (want_list1 ? 3 : 1) : 0;
  da.packed16 = (refine != 2 && g.cap[0] <= 0xFFFF && g.cap[1] <= 0xFFFF) ? 1 : 0;
  k_deliver<<<n_jobs, 256, 0, ctx->stream>>>(da);
  CU_LAUNCH_CHECK(ctx);
  if (want_ranges) CU_TRY(ctx, cudaMemcpyAsync(ctx->pin_ranges, d_rng, L.rstride * n_jobs, cudaMemcpyDeviceToHost, ctx->stream));
  return VISOCU_OK;
}

extern "C" int visocu_match_fused_submit(visocu_ctx* ctx, int32_t n_jobs, const visocu_quad* jobs, int32_t refine, int32_t flags,
                                         int32_t after_lane) {
  const int32_t want_ranges = flags;
  if (!ctx) return VISOCU_EINVAL;
  if (!ctx->configured) return visocu_set_error(ctx, VISOCU_ESTATE, "context not configured");
  if (n_jobs <= 0 || n_jobs > VISO_MAX_BATCH || !jobs) return visocu_set_error(ctx, VISOCU_EINVAL, "bad fused match arguments (at most %d jobs)", VISO_MAX_BATCH);
  if (ctx->g.first_pass != 0) return visocu_set_error(ctx, VISOCU_EINVAL, "fused matching needs multi_stage");
  if (refine < 0 || refine > 1) return visocu_set_error(ctx, VISOCU_EINVAL, "fused matching: refine must be 0 or 1");
  if (ctx->fused_pending) return visocu_set_error(ctx, VISOCU_ESTATE, "the lane's previous fused call has not been collected");
  CU_TRY(ctx, cudaSetDevice(ctx->device));
  const Geometry& g = ctx->g;
  const int nbin = g.ub * g.vb;
  const FusedLayout L = fused_layout(g);
  const size_t need = L.rstride * n_jobs + (size_t)n_jobs * 9 * nbin * sizeof(float);
  if (need > ctx->d_ranges_bytes) {
    CU_TRY(ctx, visocu_stream_wait(ctx));
    visocu_drop_lane_graphs(ctx);
    if (ctx->d_ranges) cudaFree(ctx->d_ranges);
    ctx->d_ranges = nullptr; ctx->d_ranges_bytes = 0;
    CU_TRY(ctx, cudaMalloc(&ctx->d_ranges, need));
    ctx->d_ranges_bytes = need;
    if (ctx->pin_ranges) cudaFreeHost(ctx->pin_ranges);
    ctx->pin_ranges = nullptr;
    CU_TRY(ctx, cudaMallocHost(&ctx->pin_ranges, L.rstride * VISO_MAX_BATCH));
  }
  if (L.dstride * n_jobs > ctx->deliver_bytes) {
    CU_TRY(ctx, visocu_stream_wait(ctx));
    visocu_drop_lane_graphs(ctx);
    if (ctx->deliver) cudaFreeHost(ctx->deliver);
    ctx->deliver = nullptr; ctx->deliver_bytes = 0;
    CU_TRY(ctx, cudaHostAlloc(&ctx->deliver, L.dstride * n_jobs, cudaHostAllocMapped));
    CU_TRY(ctx, cudaHostGetDevicePointer(&ctx->deliver_dev, ctx->deliver, 0));
    ctx->deliver_bytes = L.dstride * n_jobs;
  }
  for (int j = 0; j < n_jobs; j++) {
    const int f[2] = {jobs[j].f1p, jobs[j].f1c};
    for (int k = 0; k < 2; k++)
      if (f[k] < 0 || f[k] >= ctx->n_frames || !ctx->frame_valid[f[k]]) return visocu_set_error(ctx, VISOCU_ESTATE, "frame %d holds no features", f[k]);
  }
  // the previous frames may have been pushed on another lane: its feature kernels must be done before ours read them
  if (after_lane >= 0 && after_lane < VISO_LANES && after_lane != ctx->lane && ctx->lanes[after_lane].ev_push)
    CU_TRY(ctx, cudaStreamWaitEvent(ctx->stream, ctx->lanes[after_lane].ev_push, 0));
  uint64_t key = 1469598103934665603ull;
  auto mix = [&](const void* p, size_t n) { const uint8_t* b = (const uint8_t*)p; for (size_t i = 0; i < n; i++) { key ^= b[i]; key *= 1099511628211ull; } };
  mix(&n_jobs, sizeof n_jobs); mix(jobs, sizeof(visocu_quad) * (size_t)n_jobs); mix(&refine, sizeof refine); mix(&want_ranges, sizeof want_ranges);
  mix(ctx->ro_bound, sizeof ctx->ro_bound);
  ctx->in_step++;
  const int rc = visocu_run_or_replay(ctx, ctx->g_match, key, [&]() -> int { return fused_enqueue(ctx, n_jobs, jobs, refine, (want_ranges & 1) != 0, (want_ranges & 2) != 0); });
  ctx->in_step--;
  if (rc) return rc;
  CU_TRY(ctx, visocu_stream_signal(ctx, &ctx->fused_seq));
  ctx->fused_pending = true; ctx->fused_n = n_jobs; ctx->fused_ranges = (want_ranges & 1) != 0; ctx->fused_list1 = (want_ranges & 2) != 0;
  ctx->fused_jobs.assign(jobs, jobs + n_jobs);
  return VISOCU_OK;
}

extern "C" int visocu_match_fused_collect(visocu_ctx* ctx, const visocu_pmatch** list1, int32_t* n1, int32_t* done1,
                                          const visocu_pmatch** list2, int32_t* n2, int32_t* done2,
                                          visocu_range* const* ranges_out, int32_t* counts, int32_t* list2_compact) {
  if (!ctx || !list1 || !n1 || !done1 || !list2 || !n2 || !done2 || !list2_compact) return ctx ? visocu_set_error(ctx, VISOCU_EINVAL, "bad collect arguments") : VISOCU_EINVAL;
  if (!ctx->fused_pending) return visocu_set_error(ctx, VISOCU_ESTATE, "no fused matching call to collect on this lane");
  CU_TRY(ctx, cudaSetDevice(ctx->device));
  ctx->fused_pending = false;
  CU_TRY(ctx, visocu_stream_wait_seq(ctx, ctx->fused_seq));          // the only wait of a push + match step
  ctx->part[0].pending = ctx->part[1].pending = false;
  const Geometry& g = ctx->g;
  const int nbin = g.ub * g.vb, n_jobs = ctx->fused_n;
  const FusedLayout L = fused_layout(g);
  *list2_compact = L.zero_copy ? ((const int32_t*)ctx->deliver)[38] : 0;        // written by k_deliver: 1 = 24, 2 = 12 bytes per match
  const visocu_quad* jobs = ctx->fused_jobs.data();
  const visocu_deferred& A = ctx->part[0];
  const visocu_deferred& B = ctx->part[1];
  int overflow = -1;
  for (int j = 0; j < n_jobs; j++) {
    const uint8_t* base = (const uint8_t*)ctx->deliver + L.dstride * j;
    const int32_t* hdr = (const int32_t*)base;
    if (L.zero_copy && hdr[38] != *list2_compact)
      return visocu_set_error(ctx, VISOCU_ECAPACITY, "job %d: a match does not fit the packed delivery format", j);
    for (int p = 0; p < 2; p++) {
      const int32_t* w = hdr + 16 * p;
      const int n = w[0], status = w[1];
      (p ? n2 : n1)[j] = n; (p ? done2 : done1)[j] = status == 0 ? 1 : 0;
      if (status == 0 && n > 3) { for (int k = 0; k < 4; k++) ctx->ro_ns[k] += (uint64_t)w[4 + k]; ctx->ro_jobs++; }
      else if (status != 0) { ctx->ro_declined++; ctx->ro_reason[status & 3]++; ctx->ro_declined_n += (uint64_t)w[3]; }
      ctx->d2h_bytes += 64 + (L.zero_copy ? (p ? (uint64_t)n * (*list2_compact == 2 ? 12 : 24) : (ctx->fused_list1 ? (uint64_t)n * 48 : 0)) : (uint64_t)n * 48);
    }
    if (L.zero_copy) { list1[j] = ctx->fused_list1 ? (const visocu_pmatch*)(base + L.off1) : nullptr; list2[j] = (const visocu_pmatch*)(base + L.off2); }
    const int fr[2] = {jobs[j].f1p, jobs[j].f1c};
    for (int k = 0; k < 2; k++) {
      ctx->h_counts[2 * (size_t)fr[k] + 0] = hdr[32 + 3 * k]; ctx->h_counts[2 * (size_t)fr[k] + 1] = hdr[33 + 3 * k];
      if (hdr[34 + 3 * k]) overflow = fr[k];
      if (counts) { counts[4 * j + 2 * k] = hdr[32 + 3 * k]; counts[4 * j + 2 * k + 1] = hdr[33 + 3 * k]; }
    }
    // shared memory of the next outlier launches: the longest match list seen per pass (declined ones included), with a
    // margin.  Tight on purpose: what the outlier CTA does not take is what kernels of other streams can use on its SM
    for (int p = 0; p < 2; p++) {
      const int c = hdr[16 * p + 3], want = (c + c / 8 + 128 + 255) & ~255;          // in steps of 256: the bound is part of the graph key
      if (c > ctx->ro_bound_seen[p]) { ctx->ro_bound_seen[p] = c; if (ctx->ro_bound[p] < want) ctx->ro_bound[p] = want < g.cap[p] ? want : g.cap[p]; }
    }
    ctx->d2h_bytes += 24;
  }
  if (overflow >= 0) return visocu_set_error(ctx, VISOCU_ECAPACITY, "feature list of frame %d overflowed", overflow);
  if (L.zero_copy && !ctx->fused_list1) {
    // the first list stayed on the device; a job whose first list the outlier kernel declined needs it on the host
    size_t total = 0;
    for (int j = 0; j < n_jobs; j++) if (!done1[j]) total += align_up((size_t)n1[j] * 48 + 48, 256);
    if (total > 0) {
      if (total > ctx->deliver2_bytes) {
        if (ctx->deliver2) cudaFreeHost(ctx->deliver2);
        ctx->deliver2 = nullptr; ctx->deliver2_bytes = 0;
        CU_TRY(ctx, cudaMallocHost(&ctx->deliver2, total));
        ctx->deliver2_bytes = total;
      }
      uint8_t* dst = (uint8_t*)ctx->deliver2;
      for (int j = 0; j < n_jobs; j++) {
        if (done1[j]) continue;
        list1[j] = (const visocu_pmatch*)dst;
        if (n1[j] > 0) CU_TRY(ctx, cudaMemcpyAsync(dst, A.dev_lists + A.ostride * j, (size_t)n1[j] * 48, cudaMemcpyDeviceToHost, ctx->stream));
        dst += align_up((size_t)n1[j] * 48 + 48, 256);
      }
      CU_TRY(ctx, visocu_stream_wait(ctx));
    }
  }
  if (!L.zero_copy) {
    size_t total = 0;
    for (int j = 0; j < n_jobs; j++) total += align_up((size_t)n1[j] * 48 + 48, 256) + align_up((size_t)n2[j] * 48 + 48, 256);
    if (total > ctx->deliver2_bytes) {
      if (ctx->deliver2) cudaFreeHost(ctx->deliver2);
      ctx->deliver2 = nullptr; ctx->deliver2_bytes = 0;
      CU_TRY(ctx, cudaMallocHost(&ctx->deliver2, total + total / 4));
      ctx->deliver2_bytes = total + total / 4;
    }
    uint8_t* dst = (uint8_t*)ctx->deliver2;
    for (int j = 0; j < n_jobs; j++) {
      list1[j] = (const visocu_pmatch*)dst;
      if (n1[j] > 0) CU_TRY(ctx, cudaMemcpyAsync(dst, A.dev_lists + A.ostride * j, (size_t)n1[j] * 48, cudaMemcpyDeviceToHost, ctx->stream));
      dst += align_up((size_t)n1[j] * 48 + 48, 256);
      list2[j] = (const visocu_pmatch*)dst;
      if (n2[j] > 0) CU_TRY(ctx, cudaMemcpyAsync(dst, B.dev_lists + B.ostride * j, (size_t)n2[j] * 48, cudaMemcpyDeviceToHost, ctx->stream));
      dst += align_up((size_t)n2[j] * 48 + 48, 256);
    }
    CU_TRY(ctx, visocu_stream_wait(ctx));
  }
  if (ctx->fused_ranges) ctx->d2h_bytes += L.rstride * n_jobs;
  if (ranges_out && ctx->fused_ranges)
    for (int j = 0; j < n_jobs; j++)
      if (ranges_out[j]) memcpy(ranges_out[j], (const uint8_t*)ctx->pin_ranges + L.rstride * j, (size_t)nbin * sizeof(visocu_range));
  return VISOCU_OK;
}

extern "C" int visocu_match_fused(visocu_ctx* ctx, int32_t n_jobs, const visocu_quad* jobs, int32_t refine,
                                  const visocu_pmatch** list1, int32_t* n1, int32_t* done1,
                                  const visocu_pmatch** list2, int32_t* n2, int32_t* done2,
                                  visocu_range* const* ranges_out, int32_t* counts, int32_t* list2_compact) {
  const int rc = visocu_match_fused_submit(ctx, n_jobs, jobs, refine, (ranges_out ? 1 : 0) | 2, -1);
  if (rc) return rc;
  return visocu_match_fused_collect(ctx, list1, n1, done1, list2, n2, done2, ranges_out, counts, list2_compact);
}

extern "C" int visocu_refine(visocu_ctx* ctx, const visocu_quad* job, int32_t method, int32_t mode, visocu_pmatch* inout, int32_t n,
                             int32_t* n_out) {
  if (!ctx || !job || (!inout && n > 0) || !n_out) return VISOCU_EINVAL;
  if (!ctx->configured) return visocu_set_error(ctx, VISOCU_ESTATE, "context not configured");
  if (method < 0 || method > 2 || mode < 1 || mode > 2) return visocu_set_error(ctx, VISOCU_EINVAL, "bad method / mode");
  *n_out = n;
  if (n <= 0) return VISOCU_OK;
  CU_TRY(ctx, cudaSetDevice(ctx->device));
  if (mode == 2) { int rc0 = upload_pinv(ctx); if (rc0) return rc0; }
  MatchJob J;
  int rc = fill_job(ctx, *job, method, 1, J);
  if (rc) return rc;
  size_t o_list = align_up(sizeof(MatchJob), 256), o_keep = o_list + align_up((size_t)n * 48, 256);
  if ((rc = visocu_ensure_scratch(ctx, o_keep + (size_t)n + 256))) return rc;
  uint8_t* sb = (uint8_t*)ctx->scratch;
  visocu_pmatch* d_list = (visocu_pmatch*)(sb + o_list);
  CU_COPY(ctx, sb, &J, sizeof J, cudaMemcpyHostToDevice);
  CU_COPY(ctx, d_list, inout, (size_t)n * 48, cudaMemcpyHostToDevice);
  int gx = (n + 7) / 8; if (gx > 4096) gx = 4096;
  k_refine<<<dim3(gx, 1), 256, 0, ctx->stream>>>(ctx->g, (const MatchJob*)sb, method, mode, d_list, n, sb + o_keep);
  CU_LAUNCH_CHECK(ctx);
  std::vector<uint8_t> keep(n);
  CU_COPY(ctx, inout, d_list, (size_t)n * 48, cudaMemcpyDeviceToHost);
  CU_COPY(ctx, keep.data(), sb + o_keep, (size_t)n, cudaMemcpyDeviceToHost);
  CU_TRY(ctx, visocu_stream_wait(ctx));
  if (mode == 2) {
    int kept = 0;
    for (int i = 0; i < n; i++)
      if (keep[i]) inout[kept++] = inout[i];
    *n_out = kept;
  }
  return VISOCU_OK;
}

extern "C" int visocu_match_stats(visocu_ctx* ctx, uint64_t* sad_candidates, uint64_t* entries_scanned) {
  if (!ctx) return VISOCU_EINVAL;
  CU_TRY(ctx, cudaSetDevice(ctx->device));
  uint64_t now[2];
  CU_COPY(ctx, now, ctx->d_stats, sizeof now, cudaMemcpyDeviceToHost);
  CU_TRY(ctx, visocu_stream_wait(ctx));
  if (sad_candidates) *sad_candidates = now[0] - ctx->h_stats[0];
  if (entries_scanned) *entries_scanned = now[1] - ctx->h_stats[1];
  ctx->h_stats[0] = now[0]; ctx->h_stats[1] = now[1];
  return VISOCU_OK;
}
