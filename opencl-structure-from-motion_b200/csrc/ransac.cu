// Normalised 8-point fundamental-matrix RANSAC.
//
// What it reproduces (reference paths relative to /root/reference/viso):
//   VisualOdometryMono::ransacEstimateF ... viso_mono.cpp:41-72   (virtual hook, viso_mono.h:74)
//   VisualOdometryMono::fundamentalMatrix . viso_mono.cpp:265-296 (constraint rows 272-283, rank-2 step 291-295)
//   VisualOdometryMono::getInlier ......... viso_mono.cpp:298-345 (FP64 Sampson distance on float inputs)
// and the reference's OpenCL offload of the scoring (viso/kernels/plane_and_inliers.cl:12-139, float32 + ushort
// counters, 3 launches per 16 hypotheses) which this replaces with FP64 kernels over all hypotheses at once.
//
// Design (not the reference's Numerical-Recipes svdcmp, matrix.cpp:586-814):
//   k_hypotheses  three hypotheses per warp.  Lane 9h + c owns column c of the 8x9 constraint matrix of hypothesis h and of V;
//                 a one-sided (Hestenes) Jacobi SVD orthogonalises the columns with a 9-round round-robin
//                 pairing (partner of column i in round r is (r - i) mod 9), partner columns travel by
//                 warp shuffle.  The column that ends with the smallest norm marks the null vector; the
//                 rank-2 step is a 3x3 Jacobi done redundantly by every lane.
//   k_score       batched pass: one thread per match, hypotheses streamed through shared memory, per-
//                 hypothesis counts by warp ballot.
//   k_finish      one CTA per job: arg-max with the earliest-wins tie rule, inlier mask of the winner,
//                 Householder QR of the inlier constraint matrix down to a 9x9 R, then the same warp Jacobi.
// Floating-point contract: results agree with the reference to rounding (tests state the tolerance); they are
// not bit-identical because the SVD algorithm differs and nvcc contracts to FMA.
#include "visocu_internal.cuh"
#include <cstring>

namespace {

constexpr double JACOBI_TOL = 1e-15;
constexpr int JACOBI_MAX_SWEEPS = 40;
constexpr int HYP_PER_TILE = 64;       // hypotheses staged in shared memory per scoring step
constexpr int SCORE_THREADS = 256;
constexpr int FINISH_THREADS = 512;

struct RansacJob {
  const float4* uv;        // N x (u1p, v1p, u1c, v1c), normalised
  const int32_t* samples;  // iters x 8
  double* F_all;           // iters x 9
  int32_t* counts;         // iters
  uint8_t* mask;           // N
  int32_t* inl;            // N (compacted inlier indices)
  double* A;               // 9 columns x lda (column-major constraint matrix of the inliers)
  double* F9;              // 9
  int32_t* n_inl;          // [0] = inlier count, [1] = winning hypothesis
  int N, lda;
};

// constraint row of one match (viso_mono.cpp:272-283): the products are float products, as in the reference
__device__ __forceinline__ void constraint_row(const float4 m, double* a) {
  const float u1p = m.x, v1p = m.y, u1c = m.z, v1c = m.w;
  a[0] = (double)__fmul_rn(u1c, u1p); a[1] = (double)__fmul_rn(u1c, v1p); a[2] = (double)u1c;
  a[3] = (double)__fmul_rn(v1c, u1p); a[4] = (double)__fmul_rn(v1c, v1p); a[5] = (double)v1c;
  a[6] = (double)u1p; a[7] = (double)v1p; a[8] = 1.0;
}

// One-sided Jacobi on the columns of R x 9 matrices held one column per lane, THREE matrices per warp: lane = 9 h + c
// owns column c of matrix h (lanes 27..31 idle but shuffling).  On return nullv[0..8] on every lane of group h = the right
// singular vector of the smallest singular value of matrix h.
template <int R>
__device__ __forceinline__ void warp_jacobi_null(double (&g)[R], int lane, double (&nullv)[9]) {
  const int c = lane % 9, h = lane / 9;
  const bool col = lane < 27;
  const int base = col ? 9 * h : 0;
  double v[9];
#pragma unroll
  for (int k = 0; k < 9; k++) v[k] = (col && k == c) ? 1.0 : 0.0;
  // A column whose norm has fallen 14 orders of magnitude below the matrix norm is numerically null (the 8 x 9
  // constraint matrix always has one): its entries are rounding noise that no rotation can orthogonalise any further,
  // so it is left alone instead of being rotated until the sweep limit.
  double own = 0;
#pragma unroll
  for (int k = 0; k < R; k++) own = fma(g[k], g[k], own);
  double total = 0;
#pragma unroll
  for (int k = 0; k < 9; k++) total += __shfl_sync(0xFFFFFFFFu, own, base + k);
  const double tiny = 1e-28 * total;
  for (int sweep = 0; sweep < JACOBI_MAX_SWEEPS; sweep++) {
    bool rotated = false;
    for (int r = 0; r < 9; r++) {
      int j = r - c; if (j < 0) j += 9;
      const int src = col ? base + j : lane;
      double og[R], ov[9];
#pragma unroll
      for (int k = 0; k < R; k++) og[k] = __shfl_sync(0xFFFFFFFFu, g[k], src);
#pragma unroll
      for (int k = 0; k < 9; k++) ov[k] = __shfl_sync(0xFFFFFFFFu, v[k], src);
      if (col && j != c) {
        const bool first = c < j;          // this lane holds the lower-numbered column of the pair
        double alpha = 0, beta = 0, gamma = 0;
#pragma unroll
        for (int k = 0; k < R; k++) {
          const double a = first ? g[k] : og[k], b = first ? og[k] : g[k];
          alpha = fma(a, a, alpha); beta = fma(b, b, beta); gamma = fma(a, b, gamma);
        }
        if (alpha > tiny && beta > tiny && fabs(gamma) > JACOBI_TOL * sqrt(alpha * beta)) {
          const double zeta = (beta - alpha) / (2.0 * gamma);
          const double t = copysign(1.0, zeta) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
          const double cs = 1.0 / sqrt(1.0 + t * t), sn = cs * t;
          // a' = c a - s b ; b' = s a + c b
          const double mo = first ? -sn : sn;
#pragma unroll
          for (int k = 0; k < R; k++) g[k] = cs * g[k] + mo * og[k];
#pragma unroll
          for (int k = 0; k < 9; k++) v[k] = cs * v[k] + mo * ov[k];
          rotated = true;
        }
      }
    }
    if (!__any_sync(0xFFFFFFFFu, rotated)) break;
  }
  double nrm = 0;
#pragma unroll
  for (int k = 0; k < R; k++) nrm = fma(g[k], g[k], nrm);
  int best = 0;
  double bn = __shfl_sync(0xFFFFFFFFu, nrm, base);
#pragma unroll
  for (int k = 1; k < 9; k++) {
    const double on = __shfl_sync(0xFFFFFFFFu, nrm, base + k);
    if (on < bn) { bn = on; best = k; }
  }
#pragma unroll
  for (int k = 0; k < 9; k++) nullv[k] = __shfl_sync(0xFFFFFFFFu, v[k], base + best);
}

// rank-2 projection of a 3x3 matrix (row-major): F <- F - sigma3 u3 v3^T  (viso_mono.cpp:291-295)
__device__ __forceinline__ void rank2(double (&F)[9]) {
  double g[3][3], v[3][3];     // columns: g[c][r]
#pragma unroll
  for (int c = 0; c < 3; c++)
#pragma unroll
    for (int r = 0; r < 3; r++) { g[c][r] = F[3 * r + c]; v[c][r] = (r == c) ? 1.0 : 0.0; }
  for (int sweep = 0; sweep < JACOBI_MAX_SWEEPS; sweep++) {
    bool rotated = false;
#pragma unroll
    for (int pr = 0; pr < 3; pr++) {
      const int p = pr == 2 ? 1 : 0, q = pr == 0 ? 1 : 2;
      double alpha = 0, beta = 0, gamma = 0;
#pragma unroll
      for (int k = 0; k < 3; k++) { alpha = fma(g[p][k], g[p][k], alpha); beta = fma(g[q][k], g[q][k], beta); gamma = fma(g[p][k], g[q][k], gamma); }
      if (fabs(gamma) > JACOBI_TOL * sqrt(alpha * beta)) {
        const double zeta = (beta - alpha) / (2.0 * gamma);
        const double t = copysign(1.0, zeta) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
        const double c = 1.0 / sqrt(1.0 + t * t), s = c * t;
#pragma unroll
        for (int k = 0; k < 3; k++) {
          const double a = g[p][k], b = g[q][k];
          g[p][k] = c * a - s * b; g[q][k] = s * a + c * b;
          const double va = v[p][k], vb = v[q][k];
          v[p][k] = c * va - s * vb; v[q][k] = s * va + c * vb;
        }
        rotated = true;
      }
    }
    if (!rotated) break;
  }
  double n0 = 0, n1 = 0, n2 = 0;
#pragma unroll
  for (int k = 0; k < 3; k++) { n0 = fma(g[0][k], g[0][k], n0); n1 = fma(g[1][k], g[1][k], n1); n2 = fma(g[2][k], g[2][k], n2); }
  const int js = (n0 <= n1 && n0 <= n2) ? 0 : (n1 <= n2 ? 1 : 2);
#pragma unroll
  for (int r = 0; r < 3; r++)
#pragma unroll
    for (int c = 0; c < 3; c++) {
      const double gr = js == 0 ? g[0][r] : (js == 1 ? g[1][r] : g[2][r]);
      const double vc = js == 0 ? v[0][c] : (js == 1 ? v[1][c] : v[2][c]);
      F[3 * r + c] -= gr * vc;
    }
}

__global__ void __launch_bounds__(128) k_hypotheses(const RansacJob* jobs, int iters) {
  const RansacJob& J = jobs[blockIdx.y];
  const int lane = threadIdx.x & 31, c = lane % 9, h = lane / 9;
  const int hyp = (blockIdx.x * 4 + (threadIdx.x >> 5)) * 3 + h;       // three hypotheses per warp
  const bool live = lane < 27 && hyp < iters;
  if ((blockIdx.x * 4 + (threadIdx.x >> 5)) * 3 >= iters) return;      // whole warp beyond the table
  // lane 9 h + c builds column c of the 8 x 9 constraint matrix of hypothesis h
  double g[8];
#pragma unroll
  for (int r = 0; r < 8; r++) {
    double x = 0;
    if (live) {
      const int idx = J.samples[hyp * 8 + r];
      double row[9];
      constraint_row(J.uv[idx], row);
#pragma unroll
      for (int k = 0; k < 9; k++) x = (k == c) ? row[k] : x;
    }
    g[r] = x;
  }
  double f[9];
  warp_jacobi_null<8>(g, lane, f);
  rank2(f);
  if (live) {
    double x = 0;
#pragma unroll
    for (int k = 0; k < 9; k++) x = (k == c) ? f[k] : x;
    J.F_all[(size_t)hyp * 9 + c] = x;
  }
}

// |Sampson distance| < thresh, exactly the expression order of viso_mono.cpp:322-341
__device__ __forceinline__ bool is_inlier(const double* f, double u1, double v1, double u2, double v2, double thresh) {
  const double Fx1u = f[0] * u1 + f[1] * v1 + f[2];
  const double Fx1v = f[3] * u1 + f[4] * v1 + f[5];
  const double Fx1w = f[6] * u1 + f[7] * v1 + f[8];
  const double Ftx2u = f[0] * u2 + f[3] * v2 + f[6];
  const double Ftx2v = f[1] * u2 + f[4] * v2 + f[7];
  const double x2tFx1 = u2 * Fx1u + v2 * Fx1v + Fx1w;
  const double d = x2tFx1 * x2tFx1 / (Fx1u * Fx1u + Fx1v * Fx1v + Ftx2u * Ftx2u + Ftx2v * Ftx2v);
  return fabs(d) < thresh;
}

__global__ void __launch_bounds__(SCORE_THREADS) k_score(const RansacJob* jobs, int iters, double thresh) {
  const RansacJob& J = jobs[blockIdx.z];
  if ((int)blockIdx.x * SCORE_THREADS >= J.N) return;
  __shared__ double sF[HYP_PER_TILE * 9];
  __shared__ int sCnt[HYP_PER_TILE];
  const int i = blockIdx.x * SCORE_THREADS + threadIdx.x;
  const bool live = i < J.N;
  double u1 = 0, v1 = 0, u2 = 0, v2 = 0;
  if (live) { const float4 m = J.uv[i]; u1 = m.x; v1 = m.y; u2 = m.z; v2 = m.w; }
  // blockIdx.y strides over hypothesis tiles
  for (int h0 = blockIdx.y * HYP_PER_TILE; h0 < iters; h0 += gridDim.y * HYP_PER_TILE) {
    const int nh = min(HYP_PER_TILE, iters - h0);
    __syncthreads();
    for (int k = threadIdx.x; k < nh * 9; k += SCORE_THREADS) sF[k] = J.F_all[(size_t)h0 * 9 + k];
    if (threadIdx.x < HYP_PER_TILE) sCnt[threadIdx.x] = 0;
    __syncthreads();
    for (int h = 0; h < nh; h++) {
      const bool in = live && is_inlier(sF + 9 * h, u1, v1, u2, v2, thresh);
      const unsigned b = __ballot_sync(0xFFFFFFFFu, in);
      if ((threadIdx.x & 31) == 0 && b) atomicAdd(&sCnt[h], __popc(b));
    }
    __syncthreads();
    if (threadIdx.x < nh && sCnt[threadIdx.x]) atomicAdd(&J.counts[h0 + threadIdx.x], sCnt[threadIdx.x]);
  }
}

__device__ __forceinline__ double block_sum(double x, double* red, int tid) {
  for (int o = 16; o; o >>= 1) x += __shfl_xor_sync(0xFFFFFFFFu, x, o);
  __syncthreads();
  if ((tid & 31) == 0) red[tid >> 5] = x;
  __syncthreads();
  double s = 0;
  for (int k = 0; k < FINISH_THREADS / 32; k++) s += red[k];   // same order on every thread: deterministic
  return s;
}

__global__ void __launch_bounds__(FINISH_THREADS) k_finish(const RansacJob* jobs, int iters, double thresh) {
  const RansacJob& J = jobs[blockIdx.x];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  __shared__ long long s_best[FINISH_THREADS / 32];
  __shared__ double s_red[FINISH_THREADS / 32];
  __shared__ double s_F[9];
  __shared__ int s_wsum[FINISH_THREADS / 32];
  __shared__ int s_carry, s_bestk;
  __shared__ double s_R[9][9];
  __shared__ double s_dots[9];

  // ---- arg-max: largest count, earliest hypothesis on ties (viso_mono.cpp:56-57 uses a strict '>')
  long long key = -1;
  for (int k = tid; k < iters; k += FINISH_THREADS) {
    long long cand = ((long long)J.counts[k] << 32) | (unsigned)(0x7FFFFFFF - k);
    key = cand > key ? cand : key;
  }
  for (int o = 16; o; o >>= 1) { long long other = __shfl_xor_sync(0xFFFFFFFFu, key, o); key = other > key ? other : key; }
  if (lane == 0) s_best[wid] = key;
  __syncthreads();
  if (tid == 0) {
    long long b = -1;
    for (int k = 0; k < FINISH_THREADS / 32; k++) b = s_best[k] > b ? s_best[k] : b;
    const int bestk = b < 0 ? 0 : 0x7FFFFFFF - (int)(b & 0xFFFFFFFF);
    s_bestk = bestk; s_carry = 0;
    for (int k = 0; k < 9; k++) s_F[k] = iters > 0 ? J.F_all[(size_t)bestk * 9 + k] : 0.0;
  }
  __syncthreads();

  // ---- inlier mask of the winner + ordered compaction of the inlier indices
  for (int i0 = 0; i0 < J.N; i0 += FINISH_THREADS) {
    const int i = i0 + tid;
    int in = 0;
    if (i < J.N) {
      const float4 m = J.uv[i];
      in = is_inlier(s_F, m.x, m.y, m.z, m.w, thresh) ? 1 : 0;
      J.mask[i] = (uint8_t)in;
    }
    int incl = in;
    for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xFFFFFFFFu, incl, o); if (lane >= o) incl += t; }
    if (lane == 31) s_wsum[wid] = incl;
    __syncthreads();
    if (tid < 32) {
      int v = tid < FINISH_THREADS / 32 ? s_wsum[tid] : 0, s = v;
      for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xFFFFFFFFu, s, o); if (lane >= o) s += t; }
      if (tid < FINISH_THREADS / 32) s_wsum[tid] = s - v;
    }
    __syncthreads();
    const int pos = s_carry + s_wsum[wid] + incl - in;
    if (in) J.inl[pos] = i;
    __syncthreads();
    if (tid == FINISH_THREADS - 1) s_carry = pos + in;
    __syncthreads();
  }
  const int n = s_carry;
  if (tid == 0) { J.n_inl[0] = n; J.n_inl[1] = s_bestk; }
  if (n < 10) {                                   // viso_mono.cpp:59-60: F = Matrix()
    if (tid < 9) J.F9[tid] = 0.0;
    return;
  }

  // ---- refit on all inliers (viso_mono.cpp:61-69): A = Q R by Householder, then SVD of the 9x9 R
  const int lda = J.lda;
  for (int r = tid; r < n; r += FINISH_THREADS) {
    double row[9];
    constraint_row(J.uv[J.inl[r]], row);
#pragma unroll
    for (int c = 0; c < 9; c++) J.A[(size_t)c * lda + r] = row[c];
  }
  __syncthreads();
  for (int j = 0; j < 9; j++) {
    double* x = J.A + (size_t)j * lda;
    double part = 0;
    for (int r = j + tid; r < n; r += FINISH_THREADS) part = fma(x[r], x[r], part);
    const double sigma = block_sum(part, s_red, tid);
    const double x0 = x[j];
    __syncthreads();
    if (sigma == 0.0) {
      if (tid < 9) s_R[j][tid] = 0.0;
      __syncthreads();
      continue;
    }
    const double alpha = x0 >= 0 ? -sqrt(sigma) : sqrt(sigma);
    const double vnorm2 = 2.0 * (sigma - x0 * alpha);          // ||v||^2 with v = x - alpha e1
    if (tid == 0) x[j] = x0 - alpha;                            // store v in place
    __syncthreads();
    for (int c = j + 1; c < 9; c++) {
      const double* y = J.A + (size_t)c * lda;
      double p = 0;
      for (int r = j + tid; r < n; r += FINISH_THREADS) p = fma(x[r], y[r], p);
      const double d = block_sum(p, s_red, tid);
      if (tid == 0) s_dots[c] = d;
    }
    __syncthreads();
    for (int c = j + 1; c < 9; c++) {
      double* y = J.A + (size_t)c * lda;
      const double f = 2.0 * s_dots[c] / vnorm2;
      for (int r = j + tid; r < n; r += FINISH_THREADS) y[r] -= f * x[r];
    }
    __syncthreads();
    if (tid < 9) s_R[j][tid] = tid < j ? 0.0 : (tid == j ? alpha : J.A[(size_t)tid * lda + j]);
    __syncthreads();
  }
  if (wid == 0) {
    double g[9];
#pragma unroll
    for (int r = 0; r < 9; r++) g[r] = lane < 9 ? s_R[r][lane] : 0.0;
    double f[9];
    warp_jacobi_null<9>(g, lane, f);
    rank2(f);
    if (lane < 9) {
      double x = 0;
#pragma unroll
      for (int c = 0; c < 9; c++) x = (c == lane) ? f[c] : x;
      J.F9[lane] = x;
    }
  }
}

}  // namespace

extern "C" int visocu_ransac_F(visocu_ctx* ctx, int32_t n_jobs, const float* const* uv, const int32_t* N,
                               const int32_t* const* samples, int32_t iters, double thresh,
                               double* F9, uint8_t* const* inlier_mask, int32_t* n_inliers, int32_t* best_iter,
                               int32_t* const* counts, double* const* F_all) {
  if (!ctx) return VISOCU_EINVAL;
  if (n_jobs <= 0 || !uv || !N || !samples || iters <= 0 || !F9 || !n_inliers)
    return visocu_set_error(ctx, VISOCU_EINVAL, "bad ransac arguments");
  CU_TRY(ctx, cudaSetDevice(ctx->device));
  for (int start = 0; start < n_jobs; start += VISO_MAX_BATCH) {
    const int nb = n_jobs - start < VISO_MAX_BATCH ? n_jobs - start : VISO_MAX_BATCH;
    std::vector<RansacJob> hj(nb);
    // Device scratch mirrors the pinned staging area for everything that goes up (job descriptors, match coordinates,
    // sample tables: ONE upload) and keeps everything that comes back in two contiguous blocks (one 128-byte result
    // slot and one mask row per job: TWO read-backs) - API calls, not bytes, are what this call costs.
    std::vector<size_t> o_uv(nb), o_smp(nb), o_F(nb), o_inl(nb), o_A(nb);
    size_t off = align_up(sizeof(RansacJob) * nb, 256);
    int maxN = 0;
    for (int j = 0; j < nb; j++) {
      const int n = N[start + j];
      if (n < 8 || !uv[start + j] || !samples[start + j]) return visocu_set_error(ctx, VISOCU_EINVAL, "ransac job %d: need >= 8 matches", start + j);
      for (int k = 0; k < iters * 8; k++)
        if (samples[start + j][k] < 0 || samples[start + j][k] >= n) return visocu_set_error(ctx, VISOCU_EINVAL, "ransac job %d: sample index out of range", start + j);
      if (n > maxN) maxN = n;
      hj[j].N = n; hj[j].lda = (n + 3) & ~3;
      o_uv[j] = off; off += align_up((size_t)n * 16, 256);
      o_smp[j] = off; off += align_up((size_t)iters * 32, 256);
    }
    const size_t up_bytes = off;                                        // uploaded in one copy
    const size_t cstride = align_up((size_t)iters * 4, 256), mstride = align_up((size_t)maxN, 256);
    const size_t o_cnt = off; off += cstride * nb;
    const size_t o_slot = off; off += (size_t)128 * nb;                 // F9 (72 B) + n_inl, best (8 B) per job
    const size_t o_mask = off; off += mstride * nb;
    for (int j = 0; j < nb; j++) {
      o_F[j] = off; off += align_up((size_t)iters * 72, 256);
      o_inl[j] = off; off += align_up((size_t)hj[j].N * 4, 256);
      o_A[j] = off; off += align_up((size_t)hj[j].lda * 72, 256);
    }
    int rc = visocu_ensure_scratch(ctx, off);
    if (rc) return rc;
    const size_t p_slot = align_up(up_bytes, 256), p_mask = p_slot + align_up((size_t)128 * nb, 256);
    if ((rc = visocu_ensure_pinned(ctx, p_mask + mstride * nb))) return rc;
    uint8_t* sb = (uint8_t*)ctx->scratch;
    uint8_t* pin = (uint8_t*)ctx->pinned;
    for (int j = 0; j < nb; j++) {
      const int n = hj[j].N;
      hj[j].uv = (const float4*)(sb + o_uv[j]); hj[j].samples = (const int32_t*)(sb + o_smp[j]);
      hj[j].F_all = (double*)(sb + o_F[j]); hj[j].counts = (int32_t*)(sb + o_cnt + cstride * j);
      hj[j].mask = sb + o_mask + mstride * j; hj[j].inl = (int32_t*)(sb + o_inl[j]); hj[j].A = (double*)(sb + o_A[j]);
      hj[j].F9 = (double*)(sb + o_slot + 128 * j); hj[j].n_inl = (int32_t*)(sb + o_slot + 128 * j + 72);
      memcpy(pin + o_uv[j], uv[start + j], (size_t)n * 16);
      memcpy(pin + o_smp[j], samples[start + j], (size_t)iters * 32);
    }
    memcpy(pin, hj.data(), sizeof(RansacJob) * nb);
    CU_COPY(ctx, sb, pin, up_bytes, cudaMemcpyHostToDevice);
    CU_TRY(ctx, cudaMemsetAsync(sb + o_cnt, 0, cstride * nb, ctx->stream));
    const RansacJob* dj = (const RansacJob*)sb;
    k_hypotheses<<<dim3((iters + 11) / 12, nb), 128, 0, ctx->stream>>>(dj, iters);
    CU_LAUNCH_CHECK(ctx);
    int ty = (iters + HYP_PER_TILE - 1) / HYP_PER_TILE;
    k_score<<<dim3((maxN + SCORE_THREADS - 1) / SCORE_THREADS, ty, nb), SCORE_THREADS, 0, ctx->stream>>>(dj, iters, thresh);
    CU_LAUNCH_CHECK(ctx);
    k_finish<<<nb, FINISH_THREADS, 0, ctx->stream>>>(dj, iters, thresh);
    CU_LAUNCH_CHECK(ctx);
    CU_COPY(ctx, pin + p_slot, sb + o_slot, (size_t)128 * nb, cudaMemcpyDeviceToHost);
    if (inlier_mask) CU_COPY(ctx, pin + p_mask, sb + o_mask, mstride * nb, cudaMemcpyDeviceToHost);
    for (int j = 0; j < nb; j++) {                  // optional per-hypothesis tables (tests): straight to the caller
      if (counts && counts[start + j])
        CU_COPY(ctx, counts[start + j], hj[j].counts, (size_t)iters * 4, cudaMemcpyDeviceToHost);
      if (F_all && F_all[start + j])
        CU_COPY(ctx, F_all[start + j], hj[j].F_all, (size_t)iters * 72, cudaMemcpyDeviceToHost);
    }
    CU_TRY(ctx, visocu_stream_wait(ctx));
    for (int j = 0; j < nb; j++) {
      memcpy(F9 + 9 * (size_t)(start + j), pin + p_slot + 128 * j, 72);
      const int32_t* nn = (const int32_t*)(pin + p_slot + 128 * j + 72);
      n_inliers[start + j] = nn[0];
      if (best_iter) best_iter[start + j] = nn[1];
      if (inlier_mask && inlier_mask[start + j]) memcpy(inlier_mask[start + j], pin + p_mask + mstride * j, (size_t)hj[j].N);
    }
  }
  return VISOCU_OK;
}
