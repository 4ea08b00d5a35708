// Feature front end: fused 5x5 filters + both non-maximum-suppression passes, ordered compaction with the
// 32-byte descriptor gather, and the (class, v_bin, u_bin) bin index.
//
// What it reproduces (reference paths relative to /root/reference/viso):
//   filter::sobel5x5 / blob5x5 / checkerboard5x5 .......... filter.cpp:316-365 (semantics restated in SURVEY.md 9.3)
//   Matcher::createHalfResolutionImage ..................... matcher.cpp:630-647
//   Matcher::nonMaximumSuppression, sparse + dense pass .... matcher.cpp:330-431, 684-694
//   Matcher::computeDescriptors + record packing ........... matcher.cpp:433-477, 707-731
//   Matcher::createIndexVector ............................. matcher.cpp:870-890
// None of it is a translation of the reference's SSE row/column passes or its integral image: the blob and
// checkerboard responses never leave the SM (they live in shared memory only), the NMS neighbourhood test is
// reduced to "window minimum == cell minimum", and the output order (cell-column-major, then class) is
// rebuilt with a deterministic scan instead of push_back.
#include "visocu_internal.cuh"
#include <cstdlib>
#include <mutex>

namespace {

constexpr int TW = 128;          // tile core width  (pixels whose du/dv this CTA writes, whose NMS cells it owns)
constexpr int TH = 64;           // tile core height
constexpr int SEG = 16;          // rows per column walker
constexpr int FILTER_THREADS = 256;
constexpr int CELLS_PER_BLOCK = 512;

// ----------------------------------------------------------------------------------------------------------
// half-resolution image: 2x2 box mean, truncating (matcher.cpp:636-647).  One thread = 4 output pixels.
__global__ void k_half_image(Geometry g, const FrameDev* frames, SlotList sl) {
  const FrameDev F = frames[sl.s[blockIdx.z]];
  int wx = blockIdx.x * blockDim.x + threadIdx.x;        // output word index along the row
  int y = blockIdx.y * blockDim.y + threadIdx.y;
  if (y >= g.hm || wx * 4 >= g.bplm) return;
  const uint8_t* r0 = F.img + (size_t)(2 * y) * g.bpl + 8 * wx;
  const uint8_t* r1 = r0 + g.bpl;
  uint32_t out = 0;
  if (8 * wx + 7 < g.bpl) {
    uint2 a = *(const uint2*)r0, b = *(const uint2*)r1;
    uint32_t aw[2] = {a.x, a.y}, bw[2] = {b.x, b.y};
#pragma unroll
    for (int k = 0; k < 4; k++) {
      uint32_t wa = aw[k >> 1] >> (16 * (k & 1)), wb = bw[k >> 1] >> (16 * (k & 1));
      uint32_t s = (wa & 255) + ((wa >> 8) & 255) + (wb & 255) + ((wb >> 8) & 255);
      uint32_t px = (4 * wx + k < g.wm) ? (s >> 2) : 0;    // pad columns stay zero
      out |= px << (8 * k);
    }
  } else {
#pragma unroll
    for (int k = 0; k < 4; k++) {
      int x = 4 * wx + k;
      uint32_t px = 0;
      if (x < g.wm) px = ((uint32_t)r0[2 * k] + r0[2 * k + 1] + r1[2 * k] + r1[2 * k + 1]) >> 2;
      out |= px << (8 * k);
    }
  }
  *(uint32_t*)(F.half + (size_t)y * g.bplm + 4 * wx) = out;
}

// ----------------------------------------------------------------------------------------------------------
// horizontal 5-tap sums of one image row at columns x-2..x+2 (p[0..4])
struct HRow { int hd, ha, h1, h3, hc, pc; };
__device__ __forceinline__ HRow hrow(const uint8_t* p) {
  int p0 = p[0], p1 = p[1], p2 = p[2], p3 = p[3], p4 = p[4];
  HRow r;
  r.hd = (p0 - p4) + 2 * (p1 - p3);          // (1,2,0,-2,-1) left -> right
  r.ha = (p0 + p4) + 4 * (p1 + p3) + 6 * p2; // (1,4,6,4,1)
  r.h1 = p0 + p1 + p2 + p3 + p4;
  r.h3 = p1 + p2 + p3;
  r.hc = (p0 + p1) - (p3 + p4);              // (1,1,0,-1,-1)
  r.pc = p2;
  return r;
}

// full-resolution Sobel planes only (the *_full planes of matcher.cpp:676 used by the refinement).
// One thread walks SEG rows of one column with a 5-row register window.
__global__ void __launch_bounds__(256) k_sobel_full(Geometry g, const FrameDev* frames, SlotList sl) {
  const FrameDev F = frames[sl.s[blockIdx.z]];
  int x = blockIdx.x * 256 + threadIdx.x;
  int y0 = blockIdx.y * 32;
  if (x >= g.bpl) return;
  int hd[5], ha[5];
  auto load = [&](int y, int slot) {
    int a = 0, d = 0;
    if (y >= 0 && y < g.h && x >= 2 && x <= g.w - 3) {
      HRow r = hrow(F.img + (size_t)y * g.bpl + x - 2);
      a = r.ha; d = r.hd;
    }
    hd[slot] = d; ha[slot] = a;
  };
#pragma unroll
  for (int k = 0; k < 4; k++) load(y0 - 2 + k, k);
  for (int r = 0; r < 32; r += 5) {
#pragma unroll
    for (int ph = 0; ph < 5; ph++) {
      int y = y0 + r + ph;
      if (r + ph < 32 && y < g.h) {
        load(y + 2, (ph + 4) % 5);
        int du = 128, dv = 128;
        if (x >= 2 && x <= g.w - 3 && y >= 2 && y <= g.h - 3) {
          int s0 = ph % 5, s1 = (ph + 1) % 5, s2 = (ph + 2) % 5, s3 = (ph + 3) % 5, s4 = (ph + 4) % 5;
          du = (((hd[s0] + hd[s4]) + 4 * (hd[s1] + hd[s3]) + 6 * hd[s2]) >> 7) + 128;
          dv = (((ha[s0] - ha[s4]) + 2 * (ha[s1] - ha[s3])) >> 7) + 128;
          du = min(max(du, 0), 255); dv = min(max(dv, 0), 255);
        }
        F.du_full[(size_t)y * g.bpl + x] = (uint8_t)du;
        F.dv_full[(size_t)y * g.bpl + x] = (uint8_t)dv;
      }
    }
  }
}

// ----------------------------------------------------------------------------------------------------------
// Fused kernel.  One CTA = one TW x TH tile of one image.
//   phase 1  one elected thread issues a TMA tile load (cp.async.bulk.tensor.3d, zero fill outside the image) of the
//            image tile plus halo into shared memory; everybody waits on the mbarrier.
//   phase 2  register-window filters.  A thread owns one 32-bit word (4 pixels) of a row segment and walks down the
//            rows.  All arithmetic is done on two 16-bit lanes per register ((b0,b2) and (b1,b3) of the word): every
//            intermediate is kept non-negative by a bias (765 on the horizontal (1,2,0,-2,-1) sum, 510 on the
//            (1,1,0,-1,-1) sum, ...), so plain 32-bit integer adds and subtracts never carry between the lanes.
//            du and dv leave as one coalesced 32-bit store per word and row; the blob response f1 (+6375) and the
//            checkerboard response f2 (+2040) go to shared memory only.
//   phase 3  NMS on the biased responses.  Dense cells: one thread per cell; sparse cells: four lanes per cell, each
//            scanning every fourth column, combined with warp shuffles.  Extrema are reduced as keys
//            (value << 8 | position) so that "first in column-major order" falls out of a min / max.  The four
//            candidates of a cell that pass the tau test get a window scan for a strictly better value.
// A cell is owned by the tile that contains its origin; the halo is nmax to the left/top and 2*nmax to the
// right/bottom because a cell spans [i,i+n] and its extremum's neighbourhood [i-n,i+2n] (matcher.cpp:356,383).
constexpr int BIAS_F1 = 6375;    // 25 * 255: f1 + BIAS_F1 >= 0
constexpr int BIAS_F2 = 2040;    //  8 * 255
#define K2(v) ((uint32_t)(v) | ((uint32_t)(v) << 16))   /* the same constant in both 16-bit lanes */

struct TileShape { int nmax, FW, FH, NWo, NW, IS, IH, FS, iw0; size_t img_bytes, smem; };
__host__ __device__ inline TileShape tile_shape(int nmax) {
  TileShape t;
  t.nmax = nmax;
  t.FW = TW + 3 * nmax; t.FH = TH + 3 * nmax;
  // response columns start at fxs = fx0 rounded down to a multiple of 4 (fx0 = x0 - nmax, x0 multiple of TW)
  const int lead = ((-nmax) % 4 + 4) % 4;                 // fx0 - fxs
  t.NWo = (lead + t.FW + 3) / 4;                           // response words per row
  // the image tile starts one word left of the response columns, moved further left to a 16-byte boundary because
  // the TMA needs a 16-byte aligned global address for the first element of the box
  t.iw0 = ((((-(nmax + lead + 4)) % 16) + 16) % 16) / 4;   // words between the tile origin and response word -1
  t.NW = (t.iw0 + t.NWo + 2 + 3) & ~3;                     // image words per row, 16 B multiple
  t.IS = t.NW * 4; t.IH = t.FH + 4;
  t.FS = t.NWo * 4;
  t.img_bytes = ((size_t)t.IH * t.IS + 127) & ~(size_t)127;
  t.smem = t.img_bytes + 2 * (size_t)t.FH * t.FS * sizeof(int16_t) + 16;   // + the mbarrier
  return t;
}

struct Win {   // 5-row windows of the horizontal sums of one word: [row slot][lane pair A=(b0,b2), B=(b1,b3)]
  uint32_t hd[5][2], ha[5][2], h1[5][2], h3[5][2], hc[5][2], pc[5][2];
};

// horizontal sums of the 4 pixels of word C (neighbours L, R) into window slot S
template <int S>
__device__ __forceinline__ void hrow4(Win& w, uint32_t L, uint32_t C, uint32_t R, bool core) {
  const uint32_t s2a = __funnelshift_r(L, C, 16);          // bytes l2 l3 b0 b1
  const uint32_t s2b = __funnelshift_r(C, R, 16);          // bytes b2 b3 r0 r1
  uint32_t t[6];
  t[0] = s2a & 0x00FF00FFu;                                // (l2, b0)
  t[1] = __byte_perm(s2a, 0, 0x4341);                      // (l3, b1)
  t[2] = C & 0x00FF00FFu;                                  // (b0, b2)
  t[3] = __byte_perm(C, 0, 0x4341);                        // (b1, b3)
  t[4] = s2b & 0x00FF00FFu;                                // (b2, r0)
  t[5] = __byte_perm(s2b, 0, 0x4341);                      // (b3, r1)
#pragma unroll
  for (int k = 0; k < 2; k++) {                            // k = 0: pixels (b0,b2), k = 1: pixels (b1,b3)
    const uint32_t t0 = t[k], t1 = t[k + 1], t2 = t[k + 2], t3 = t[k + 3], t4 = t[k + 4];
    if (core) {
      w.hd[S][k] = (t0 + 2 * t1 + K2(765)) - (t4 + 2 * t3);          // (1,2,0,-2,-1) + 765
      w.ha[S][k] = (t0 + t4) + 4 * (t1 + t3) + 6 * t2;               // (1,4,6,4,1)
    }
    w.h1[S][k] = t0 + t1 + t2 + t3 + t4;
    w.h3[S][k] = t1 + t2 + t3;
    w.hc[S][k] = (t0 + t1 + K2(510)) - (t3 + t4);                    // (1,1,0,-1,-1) + 510
    w.pc[S][k] = t2;
  }
}

__device__ __forceinline__ bool mbar_wait_or_trap(uint32_t bar, uint32_t phase) {
  const long long t0 = clock64();
  for (;;) {
    uint32_t done;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(bar), "r"(phase) : "memory");
    if (done) return true;
    if (clock64() - t0 > 2000000000LL) __trap();           // about one second: never hang the GPU on a TMA mistake
  }
}

__global__ void __launch_bounds__(FILTER_THREADS, 2)
k_filter_nms(Geometry g, const FrameDev* frames, SlotList sl, int nmax, const __grid_constant__ CUtensorMap tmap, int use_tma, int max_cells) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int slot = sl.s[blockIdx.z];
  const FrameDev F = frames[slot];
  const int tid = threadIdx.x;
  const TileShape ts = tile_shape(nmax);
  const int x0 = blockIdx.x * TW, y0 = blockIdx.y * TH;
  const int fx0 = x0 - nmax, fy0 = y0 - nmax;              // origin of the response region that NMS reads
  const int fxs = fx0 & ~3;                                // response columns in shared memory start here
  const int rx0 = fxs - 4 - 4 * ts.iw0, iy0 = fy0 - 2;      // image tile origin (16-byte aligned in x)
  const int FH = ts.FH, FS = ts.FS, IS = ts.IS, IH = ts.IH, NWo = ts.NWo;
  uint8_t* simg = smem;
  int16_t* sf1 = (int16_t*)(smem + ts.img_bytes);
  int16_t* sf2 = sf1 + (size_t)FH * FS;
  uint64_t* s_bar = (uint64_t*)(sf2 + (size_t)FH * FS);

  // ---- phase 1
  if (use_tma) {
    const uint32_t bar = (uint32_t)__cvta_generic_to_shared(s_bar);
    const uint32_t dst = (uint32_t)__cvta_generic_to_shared(simg);
    if (tid == 0) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(bar));
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0) {
      const uint32_t bytes = (uint32_t)(IH * IS);
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
      asm volatile(
          "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
          :: "r"(dst), "l"(&tmap), "r"(rx0), "r"(iy0), "r"(slot), "r"(bar) : "memory");
    }
    mbar_wait_or_trap(bar, 0);
  } else {
    const uint8_t* __restrict__ I = g.half ? F.half : F.img;
    const int words = IS >> 2;
    for (int idx = tid; idx < IH * words; idx += FILTER_THREADS) {
      const int ly = idx / words, lw = idx - ly * words;
      const int gx = rx0 + 4 * lw, gy = iy0 + ly;
      uint32_t v = 0;
      if (gy >= 0 && gy < g.hm && gx >= 0 && gx < g.bplm) v = __ldg((const uint32_t*)(I + (size_t)gy * g.bplm + gx));
      *(uint32_t*)(simg + ly * IS + 4 * lw) = v;
    }
    __syncthreads();
  }

  // ---- phase 2
  {
    const int nseg = (FH + SEG - 1) / SEG;
    for (int item = tid; item < NWo * nseg; item += FILTER_THREADS) {
      const int seg = item / NWo, j = item - seg * NWo;
      const int ly0 = seg * SEG;
      const int nrow = min(SEG, FH - ly0);
      const int gx = fxs + 4 * j;                                          // first pixel of this word
      const uint32_t* col = (const uint32_t*)simg + ts.iw0 + j;            // words L, C, R of an image-tile row
      const int IW = IS >> 2;
      const bool core = gx >= x0 && gx < x0 + TW && gx < g.bplm;           // du/dv are written for core words only
      Win w;
#define VISO_HROW(S, r) { const uint32_t* q = col + (r) * IW; hrow4<S>(w, q[0], q[1], q[2], core); }
      VISO_HROW(0, ly0) VISO_HROW(1, ly0 + 1) VISO_HROW(2, ly0 + 2) VISO_HROW(3, ly0 + 3)
      for (int r5 = 0; r5 < nrow; r5 += 5) {
#define VISO_STEP(PH)                                                                                                  \
        if (r5 + PH < nrow) {                                                                                          \
          constexpr int s0 = PH % 5, s1 = (PH + 1) % 5, s2 = (PH + 2) % 5, s3 = (PH + 3) % 5, s4 = (PH + 4) % 5;       \
          const int ly = ly0 + r5 + PH, gy = fy0 + ly;                                                                 \
          VISO_HROW(s4, ly + 4)                                                                                        \
          uint32_t f1[2], f2[2], du[2], dv[2];                                                                         \
          _Pragma("unroll") for (int k = 0; k < 2; k++) {                                                              \
            const uint32_t b3 = w.h3[s1][k] + w.h3[s2][k] + w.h3[s3][k];                                               \
            const uint32_t b5 = w.h1[s0][k] + w.h1[s1][k] + w.h1[s2][k] + w.h1[s3][k] + w.h1[s4][k];                   \
            f1[k] = (7 * w.pc[s2][k] + 2 * b3 + K2(BIAS_F1)) - b5;                                                     \
            f2[k] = (w.hc[s0][k] + w.hc[s1][k] + K2(BIAS_F2)) - (w.hc[s3][k] + w.hc[s4][k]);                           \
            if (core) {                                                                                                \
              const uint32_t a = (w.hd[s0][k] + w.hd[s4][k]) + 4 * (w.hd[s1][k] + w.hd[s3][k]) + 6 * w.hd[s2][k] + K2(4144); \
              du[k] = (a >> 7) & 0x01FF01FFu;                                                                          \
              const uint32_t b = (w.ha[s0][k] + 2 * w.ha[s1][k] + K2(12240 + 4144)) - (w.ha[s4][k] + 2 * w.ha[s3][k]); \
              dv[k] = (b >> 7) & 0x01FF01FFu;                                                                          \
            }                                                                                                          \
          }                                                                                                            \
          *(uint2*)(sf1 + ly * FS + 4 * j) = make_uint2(__byte_perm(f1[0], f1[1], 0x5410), __byte_perm(f1[0], f1[1], 0x7632)); \
          *(uint2*)(sf2 + ly * FS + 4 * j) = make_uint2(__byte_perm(f2[0], f2[1], 0x5410), __byte_perm(f2[0], f2[1], 0x7632)); \
          if (core && gy >= y0 && gy < y0 + TH && gy < g.hm) {                                                         \
            uint32_t wu = du[0] | (du[1] << 8), wv = dv[0] | (dv[1] << 8);                                             \
            if (gy < 2 || gy > g.hm - 3) { wu = 0x80808080u; wv = 0x80808080u; }                                       \
            if (gx == 0) { wu = (wu & 0xFFFF0000u) | 0x8080u; wv = (wv & 0xFFFF0000u) | 0x8080u; }                     \
            if (gx == g.bplm - 4) { wu = (wu & 0x0000FFFFu) | 0x80800000u; wv = (wv & 0x0000FFFFu) | 0x80800000u; }    \
            *(uint32_t*)(F.du + (size_t)gy * g.bplm + gx) = wu;                                                        \
            *(uint32_t*)(F.dv + (size_t)gy * g.bplm + gx) = wv;                                                        \
          }                                                                                                            \
        }
        VISO_STEP(0) VISO_STEP(1) VISO_STEP(2) VISO_STEP(3) VISO_STEP(4)
#undef VISO_STEP
      }
#undef VISO_HROW
    }
  }
  __syncthreads();

  // ---- phase 3
  {
    uint32_t* s_code = (uint32_t*)(s_bar + 2);                             // one code word per owned cell (both passes)
    uint32_t* s_queue = s_code + max_cells;                                // candidates that passed the tau test
    int* s_qn = (int*)(s_bar + 1);
    const int xhi = g.wm - 1 - VISO_MARGIN - fxs, yhi = g.hm - 1 - VISO_MARGIN - fy0;   // window clamps, smem coordinates
    int klo[2], nk[2], llo[2], nl[2], cbase[2];
    int ncell_total = 0;
#pragma unroll
    for (int p = 0; p < 2; p++) {
      klo[p] = llo[p] = nk[p] = nl[p] = 0; cbase[p] = ncell_total;
      if (p < g.first_pass) continue;
      const int step = g.n[p] + 1, org = g.n[p] + VISO_MARGIN;
      int a = x0 - org; klo[p] = a > 0 ? (a + step - 1) / step : 0;
      int b = x0 + TW - org; const int khi = b > 0 ? min((b + step - 1) / step, g.ncx[p]) : 0;
      a = y0 - org; llo[p] = a > 0 ? (a + step - 1) / step : 0;
      b = y0 + TH - org; const int lhi = b > 0 ? min((b + step - 1) / step, g.ncy[p]) : 0;
      nk[p] = max(khi - klo[p], 0); nl[p] = max(lhi - llo[p], 0);
      ncell_total += nk[p] * nl[p];
    }
    if (tid == 0) *s_qn = 0;
    __syncthreads();

    // 3a: cell extrema.  Dense cells: one thread per cell; sparse cells: four lanes per cell, each scanning every
    // fourth column, combined with warp shuffles.  Candidates that pass the tau test are queued.
#pragma unroll
    for (int p = 0; p < 2; p++) {
      if (p < g.first_pass) continue;
      const int n = g.n[p], step = n + 1, org = n + VISO_MARGIN;
      const int ncell = nk[p] * nl[p];
      const int G = p == 0 ? 4 : 1;                                        // lanes per cell
      const int nitem = ((ncell * G + 31) & ~31);                          // whole warps take part in the shuffles
      for (int item = tid; item < nitem; item += FILTER_THREADS) {
        const int cell = item / G, t = item - cell * G;
        const bool live = cell < ncell;
        const int kk = live ? cell % nk[p] : 0, ll = live ? cell / nk[p] : 0;
        const int lx = org + (klo[p] + kk) * step - fxs, ly = org + (llo[p] + ll) * step - fy0;
        int k1min = 0x7FFFFFFF, k1max = -1, k2min = 0x7FFFFFFF, k2max = -1;
        if (live) {
          for (int di = t; di <= n; di += G) {
            const int16_t* q1 = sf1 + ly * FS + lx + di;
            const int16_t* q2 = sf2 + ly * FS + lx + di;
            for (int dj = 0; dj <= n; dj++) {
              const int pos = (di << 4) | dj;
              const int a1 = q1[dj * FS] * 256 + pos, a2 = q2[dj * FS] * 256 + pos;      // (value << 8) | position
              k1min = min(k1min, a1); k1max = max(k1max, a1 ^ 255);                    // ^255: position -> 255 - position
              k2min = min(k2min, a2); k2max = max(k2max, a2 ^ 255);
            }
          }
        }
        if (G == 4) {
#pragma unroll
          for (int o = 1; o < 4; o <<= 1) {
            k1min = min(k1min, __shfl_xor_sync(0xFFFFFFFFu, k1min, o)); k1max = max(k1max, __shfl_xor_sync(0xFFFFFFFFu, k1max, o));
            k2min = min(k2min, __shfl_xor_sync(0xFFFFFFFFu, k2min, o)); k2max = max(k2max, __shfl_xor_sync(0xFFFFFFFFu, k2max, o));
          }
        }
        if (live) {
          if (t == 0) s_code[cbase[p] + cell] = 0xFFFFFFFFu;
          // class c: 0 = f1 min, 1 = f1 max, 2 = f2 min, 3 = f2 max
#pragma unroll
          for (int c = 0; c < 4; c++) {
            if (G == 4 && c != t) continue;
            const int key = c == 0 ? k1min : (c == 1 ? k1max : (c == 2 ? k2min : k2max));
            const bool is_min = (c & 1) == 0;
            const int val = key >> 8, pos = is_min ? (key & 255) : 255 - (key & 255);
            const int bias = c < 2 ? BIAS_F1 : BIAS_F2;
            if (is_min ? (val <= bias - g.tau) : (val >= bias + g.tau))
              s_queue[atomicAdd(s_qn, 1)] = (uint32_t)(cbase[p] + cell) | ((uint32_t)c << 16) | ((uint32_t)pos << 18) | ((uint32_t)p << 26);
          }
        }
      }
    }
    __syncthreads();

    // 3b: one thread per queued candidate.  Keep it iff nothing in its clamped (2n+1)^2 window is strictly better;
    // positions inside the cell can never be strictly better than the cell extremum, so the reference's cell
    // exclusion is implied.  Row-wise packed scan: two responses per 32-bit word, VIMNMX.U16x2 reduction, one exit
    // test per row; maxima are handled as minima of the complemented values; lanes outside the window read as 0xFFFF.
    const int nq = *s_qn;
    for (int q = tid; q < nq; q += FILTER_THREADS) {
      const uint32_t ent = s_queue[q];
      const int p = ent >> 26, c = (ent >> 16) & 3, pos = (ent >> 18) & 255, gcell = ent & 0xFFFF;
      const int cell = gcell - cbase[p];
      const int n = g.n[p], step = n + 1, org = n + VISO_MARGIN;
      const int kk = cell % nk[p], ll = cell / nk[p];
      const int lx = org + (klo[p] + kk) * step - fxs, ly = org + (llo[p] + ll) * step - fy0;
      const int16_t* sf = c < 2 ? sf1 : sf2;
      const bool is_min = (c & 1) == 0;
      const int ex = lx + (pos >> 4), ey = ly + (pos & 15);
      const int val = sf[ey * FS + ex];
      const int xs = ex - n, xe = min(ex + n, xhi), ye = min(ey + n, yhi);
      const uint32_t flip = is_min ? 0u : 0xFFFFFFFFu;
      const uint32_t target = is_min ? (uint32_t)val : (uint32_t)(0xFFFF - val);
      const int w0 = xs >> 1, w1 = xe >> 1;
      const uint32_t m0 = (xs & 1) ? 0x0000FFFFu : 0u, m1 = (xe & 1) ? 0u : 0xFFFF0000u;
      bool keep = true;
      for (int j2 = ey - n; j2 <= ye && keep; j2++) {
        const uint32_t* row = (const uint32_t*)(sf + j2 * FS);
        uint32_t acc = (row[w0] ^ flip) | m0;
        if (w1 > w0) {
#pragma unroll 4
          for (int wq = w0 + 1; wq < w1; wq++) acc = __vminu2(acc, row[wq] ^ flip);
          acc = __vminu2(acc, (row[w1] ^ flip) | m1);
        } else {
          acc |= m1;
        }
        if (min(acc & 0xFFFFu, acc >> 16) < target) keep = false;
      }
      if (keep) ((uint8_t*)s_code)[4 * gcell + c] = (uint8_t)pos;
    }
    __syncthreads();

    // 3c: code words to global memory, cell-column-major
    for (int idx = tid; idx < ncell_total; idx += FILTER_THREADS) {
      const int p = (idx >= cbase[1] && nk[1] * nl[1] > 0) ? 1 : 0;
      const int cell = idx - cbase[p];
      const int kk = cell % nk[p], ll = cell / nk[p];
      F.codes[p][(size_t)(klo[p] + kk) * g.ncy[p] + (llo[p] + ll)] = s_code[idx];
    }
  }
}

// ----------------------------------------------------------------------------------------------------------
// ordered compaction, step 1: records emitted per chunk of CELLS_PER_BLOCK cells (cells in column-major order)
__device__ __forceinline__ int code_count(uint32_t code) {
  return ((code & 0xFFu) != 0xFFu) + ((code & 0xFF00u) != 0xFF00u) + ((code & 0xFF0000u) != 0xFF0000u) + ((code >> 24) != 0xFFu);
}

__global__ void __launch_bounds__(CELLS_PER_BLOCK) k_cell_count(Geometry g, const FrameDev* frames, SlotList sl) {
  const int p = blockIdx.y;
  if (p < g.first_pass) return;
  const FrameDev F = frames[sl.s[blockIdx.z]];
  const int ncells = g.ncx[p] * g.ncy[p];
  const int nchunk = (ncells + CELLS_PER_BLOCK - 1) / CELLS_PER_BLOCK;
  if ((int)blockIdx.x >= nchunk) return;
  const int c = blockIdx.x * CELLS_PER_BLOCK + threadIdx.x;
  int cnt = c < ncells ? code_count(F.codes[p][c]) : 0;
  __shared__ int wsum[CELLS_PER_BLOCK / 32];
  for (int o = 16; o; o >>= 1) cnt += __shfl_xor_sync(0xFFFFFFFFu, cnt, o);
  if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = cnt;
  __syncthreads();
  if (threadIdx.x < 32) {
    int v = threadIdx.x < CELLS_PER_BLOCK / 32 ? wsum[threadIdx.x] : 0;
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
    if (threadIdx.x == 0) F.blk[p][blockIdx.x] = v;
  }
}

// descriptor sample offsets (matcher.cpp:445-476): 16 (du,dv) pairs
__constant__ int8_t c_desc_ox[16] = {-3, -3, -1, -1, 3, 3, 1, 1, -1, -1, 1, 1, -5, -5, 5, 5};
__constant__ int8_t c_desc_oy[16] = {-1, 1, -1, 1, -1, 1, -1, 1, -5, 5, -5, 5, -3, 3, -3, 3};

// step 2: chunk base = sum of earlier chunks, block scan, then one warp per record gathers the 32 descriptor
// bytes (one byte per lane), packs them to words with shuffles and stores the 48-byte record with 12 lanes.
__global__ void __launch_bounds__(CELLS_PER_BLOCK) k_emit_records(Geometry g, const FrameDev* frames, SlotList sl) {
  const int p = blockIdx.y;
  if (p < g.first_pass) return;
  const FrameDev F = frames[sl.s[blockIdx.z]];
  const int ncells = g.ncx[p] * g.ncy[p];
  const int nchunk = (ncells + CELLS_PER_BLOCK - 1) / CELLS_PER_BLOCK;
  if ((int)blockIdx.x >= nchunk) {
    if (nchunk == 0 && blockIdx.x == 0 && threadIdx.x == 0) F.counts[p] = 0;
    return;
  }
  __shared__ int wsum[CELLS_PER_BLOCK / 32];
  __shared__ int s_base, s_total;
  __shared__ uint32_t s_list[4 * CELLS_PER_BLOCK];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;

  // base of this chunk
  int acc = 0;
  for (int b = tid; b < (int)blockIdx.x; b += CELLS_PER_BLOCK) acc += F.blk[p][b];
  for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xFFFFFFFFu, acc, o);
  if (lane == 0) wsum[wid] = acc;
  __syncthreads();
  if (tid < 32) {
    int v = tid < CELLS_PER_BLOCK / 32 ? wsum[tid] : 0;
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
    if (tid == 0) s_base = v;
  }
  __syncthreads();

  // exclusive scan of the per-cell counts
  const int c = blockIdx.x * CELLS_PER_BLOCK + tid;
  const uint32_t code = c < ncells ? F.codes[p][c] : 0xFFFFFFFFu;
  const int cnt = code_count(code);
  int incl = cnt;
  for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xFFFFFFFFu, incl, o); if (lane >= o) incl += t; }
  if (lane == 31) wsum[wid] = incl;
  __syncthreads();
  if (tid < 32) {
    int v = tid < CELLS_PER_BLOCK / 32 ? wsum[tid] : 0, s = v;
    for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xFFFFFFFFu, s, o); if (lane >= o) s += t; }
    if (tid < CELLS_PER_BLOCK / 32) wsum[tid] = s - v;
    if (tid == CELLS_PER_BLOCK / 32 - 1) s_total = s;
  }
  __syncthreads();
  int off = wsum[wid] + incl - cnt;
  if (cnt) {
    const int n = g.n[p];
    const int k = c / g.ncy[p], l = c - k * g.ncy[p];
    const int i = n + VISO_MARGIN + k * (n + 1), j = n + VISO_MARGIN + l * (n + 1);
#pragma unroll
    for (int cls = 0; cls < 4; cls++) {
      uint32_t b = (code >> (8 * cls)) & 0xFFu;
      if (b != 0xFFu) s_list[off++] = (uint32_t)(i + (b >> 4)) | ((uint32_t)(j + (b & 15)) << 13) | ((uint32_t)cls << 26);
    }
  }
  __syncthreads();
  const int total = s_total, base = s_base;
  if (blockIdx.x == nchunk - 1 && tid == 0) {
    int all = base + total;
    if (all > g.cap[p]) { F.counts[2] = 1; all = g.cap[p]; }
    F.counts[p] = all;
  }
  for (int e = wid; e < total; e += CELLS_PER_BLOCK / 32) {
    if (base + e >= g.cap[p]) break;
    const uint32_t ent = s_list[e];
    const int u = ent & 0x1FFF, v = (ent >> 13) & 0x1FFF, cls = ent >> 26;
    const int s = lane >> 1;
    const uint8_t* plane = (lane & 1) ? F.dv : F.du;
    uint32_t byte = plane[(size_t)(v + c_desc_oy[s]) * g.bplm + u + c_desc_ox[s]];
    uint32_t w = byte << (8 * (lane & 3));
    w |= __shfl_xor_sync(0xFFFFFFFFu, w, 1);
    w |= __shfl_xor_sync(0xFFFFFFFFu, w, 2);
    uint32_t dw = __shfl_sync(0xFFFFFFFFu, w, ((lane - 4) & 7) * 4);
    if (lane < 12) {
      int32_t out;
      if (lane == 0) out = u * g.scale; else if (lane == 1) out = v * g.scale; else if (lane == 2) out = 0;
      else if (lane == 3) out = cls; else out = (int32_t)dw;
      F.rec[p][(size_t)(base + e) * 12 + lane] = out;
    }
  }
}

// ----------------------------------------------------------------------------------------------------------
// bin index (createIndexVector, matcher.cpp:870-890): counting sort of the records by (class, v_bin, u_bin).
// The order inside a bin is not the reference's (ascending index); the matcher compares complete
// (cost, u_bin, v_bin, index) keys instead, which reproduces the reference's first-minimum rule exactly.
__global__ void __launch_bounds__(1024) k_build_bins(Geometry g, const FrameDev* frames, SlotList sl) {
  const int p = blockIdx.y;
  if (p < g.first_pass) return;
  const FrameDev F = frames[sl.s[blockIdx.z]];
  const int tid = threadIdx.x;
  const int n = F.counts[p];
  int32_t* start = F.bin_start[p];
  int32_t* cursor = F.bin_cursor[p];
  const int32_t* rec = F.rec[p];
  for (int b = tid; b <= g.nbins; b += 1024) cursor[b] = 0;
  __syncthreads();
  const float bs = (float)g.binsize;
  auto bin_of = [&](int u, int v, int c) {
    int ubin = min((int)floorf((float)u / bs), g.ub - 1);
    int vbin = min((int)floorf((float)v / bs), g.vb - 1);
    return (c * g.vb + vbin) * g.ub + ubin;
  };
  for (int i = tid; i < n; i += 1024) {
    int4 hdr = *(const int4*)(rec + (size_t)i * 12);
    atomicAdd(&cursor[bin_of(hdr.x, hdr.y, hdr.w)], 1);
  }
  __syncthreads();
  // exclusive scan of the histogram, 1024 bins per round with a running carry
  __shared__ int wsum[32];
  __shared__ int s_carry;
  if (tid == 0) s_carry = 0;
  __syncthreads();
  for (int b0 = 0; b0 <= g.nbins; b0 += 1024) {
    const int b = b0 + tid;
    const int v = b < g.nbins ? cursor[b] : 0;
    int incl = v;
    const int lane = tid & 31, wid = tid >> 5;
    for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xFFFFFFFFu, incl, o); if (lane >= o) incl += t; }
    if (lane == 31) wsum[wid] = incl;
    __syncthreads();
    if (tid < 32) {
      int w = wsum[tid], s = w;
      for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xFFFFFFFFu, s, o); if (lane >= o) s += t; }
      wsum[tid] = s - w;
    }
    __syncthreads();
    const int excl = s_carry + wsum[wid] + incl - v;
    if (b <= g.nbins) { start[b] = excl; cursor[b] = excl; }
    __syncthreads();
    if (tid == 1023) s_carry = excl + v;
    __syncthreads();
  }
  for (int i = tid; i < n; i += 1024) {
    int4 hdr = *(const int4*)(rec + (size_t)i * 12);
    int pos = atomicAdd(&cursor[bin_of(hdr.x, hdr.y, hdr.w)], 1);
    F.bin_ent[p][pos] = make_int2(hdr.x | (hdr.y << 16), i);
  }
}

// record counts (and the overflow flag) of the frames of one launch, gathered for a single read-back
__global__ void k_gather_counts(const FrameDev* frames, SlotList sl, int32_t* out) {
  const int i = threadIdx.x;
  if (i < sl.n) {
    const int32_t* c = frames[sl.s[i]].counts;
    out[4 * i] = c[0]; out[4 * i + 1] = c[1]; out[4 * i + 2] = c[2]; out[4 * i + 3] = 0;
  }
}

}  // namespace

int visocu_launch_features(visocu_ctx* ctx, const SlotList& sl) {
  const Geometry& g = ctx->g;
  cudaStream_t st = ctx->stream;
  if (g.half) {
    dim3 bh(32, 8), gh((g.bplm / 4 + 31) / 32, (g.hm + 7) / 8, sl.n);
    k_half_image<<<gh, bh, 0, st>>>(g, ctx->frames_d, sl);
    CU_LAUNCH_CHECK(ctx);
    dim3 gs((g.bpl + 255) / 256, (g.h + 31) / 32, sl.n);
    k_sobel_full<<<gs, 256, 0, st>>>(g, ctx->frames_d, sl);
    CU_LAUNCH_CHECK(ctx);
  }
  const int nmax = g.first_pass == 0 ? (g.n[0] > g.n[1] ? g.n[0] : g.n[1]) : g.n[1];
  {
    const TileShape ts = tile_shape(nmax);
    // owned NMS cells of one tile (both passes): code words + a queue of up to four candidates per cell
    int max_cells = 0;
    for (int p = g.first_pass; p < 2; p++) max_cells += ((TW + g.n[p]) / (g.n[p] + 1) + 1) * ((TH + g.n[p]) / (g.n[p] + 1) + 1);
    const size_t smem_bytes = ts.smem + (size_t)max_cells * 20;
    // opt in to large dynamic shared memory once per device (the attribute belongs to the function on a device, not
    // to a context: several contexts with different tile shapes share it)
    {
      static std::mutex mtx;
      static bool done[64] = {false};
      std::lock_guard<std::mutex> lock(mtx);
      if (!done[ctx->device & 63]) {
        CU_TRY(ctx, cudaFuncSetAttribute(k_filter_nms, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        done[ctx->device & 63] = true;
      }
    }
    if (smem_bytes > 227 * 1024) return visocu_set_error(ctx, VISOCU_EINVAL, "tile needs %zu bytes of shared memory", smem_bytes);
    dim3 grid((g.bplm + TW - 1) / TW, (g.hm + TH - 1) / TH, sl.n);
    if (ctx->profile) CU_TRY(ctx, cudaEventRecord(ctx->pev0, st));
    k_filter_nms<<<grid, FILTER_THREADS, smem_bytes, st>>>(g, ctx->frames_d, sl, nmax, ctx->tmap_img, ctx->use_tma, max_cells);
    CU_LAUNCH_CHECK(ctx);
    if (ctx->profile) {
      // profiling mode only: this synchronises the stream after every fused launch
      float ms = 0;
      CU_TRY(ctx, cudaEventRecord(ctx->pev1, st));
      CU_TRY(ctx, cudaEventSynchronize(ctx->pev1));
      CU_TRY(ctx, cudaEventElapsedTime(&ms, ctx->pev0, ctx->pev1));
      ctx->filter_ms += ms; ctx->filter_launches++; ctx->filter_frames += sl.n;
    }
  }
  int maxchunk = 1;
  for (int p = g.first_pass; p < 2; p++) {
    int c = (g.ncx[p] * g.ncy[p] + CELLS_PER_BLOCK - 1) / CELLS_PER_BLOCK;
    if (c > maxchunk) maxchunk = c;
  }
  dim3 gc(maxchunk, 2, sl.n);
  k_cell_count<<<gc, CELLS_PER_BLOCK, 0, st>>>(g, ctx->frames_d, sl);
  CU_LAUNCH_CHECK(ctx);
  k_emit_records<<<gc, CELLS_PER_BLOCK, 0, st>>>(g, ctx->frames_d, sl);
  CU_LAUNCH_CHECK(ctx);
  dim3 gb(1, 2, sl.n);
  k_build_bins<<<gb, 1024, 0, st>>>(g, ctx->frames_d, sl);
  CU_LAUNCH_CHECK(ctx);
  if (ctx->counts_stage) {
    k_gather_counts<<<1, VISO_MAX_BATCH, 0, st>>>(ctx->frames_d, sl, ctx->counts_stage);
    CU_LAUNCH_CHECK(ctx);
  }
  return VISOCU_OK;
}

// TMA descriptor of the matching-resolution image plane of every frame slot: a 3-D tensor (bytes per line, rows,
// frame slots) with the tile box of tile_shape(nmax); out-of-bounds elements are filled with zeros, which is what
// the filters expect outside the image.
int visocu_make_tensor_map(visocu_ctx* ctx, size_t frame_stride_bytes) {
  const Geometry& g = ctx->g;
  const int nmax = g.first_pass == 0 ? (g.n[0] > g.n[1] ? g.n[0] : g.n[1]) : g.n[1];
  const TileShape ts = tile_shape(nmax);
  ctx->use_tma = 0;
  const char* env = getenv("VISOCU_TMA");
  if (env && env[0] == '0') return VISOCU_OK;             // debugging switch: stage the tile with plain loads
  typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  CU_TRY(ctx, cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
  if (!fn || qres != cudaDriverEntryPointSuccess) return visocu_set_error(ctx, VISOCU_ECUDA, "cuTensorMapEncodeTiled is not available");
  const FrameDev& F0 = ctx->frames_h[0];
  void* base = g.half ? (void*)F0.half : (void*)F0.img;
  const cuuint64_t dims[3] = {(cuuint64_t)g.bplm, (cuuint64_t)g.hm, (cuuint64_t)ctx->n_frames};
  const cuuint64_t strides[2] = {(cuuint64_t)g.bplm, (cuuint64_t)frame_stride_bytes};
  const cuuint32_t box[3] = {(cuuint32_t)ts.IS, (cuuint32_t)ts.IH, 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  if (ts.IS > 256 || ts.IH > 256) return visocu_set_error(ctx, VISOCU_EINVAL, "tile box %dx%d exceeds the TMA limit", ts.IS, ts.IH);
  CUresult r = ((EncodeFn)fn)(&ctx->tmap_img, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, base, dims, strides, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return visocu_set_error(ctx, VISOCU_ECUDA, "cuTensorMapEncodeTiled failed with %d", (int)r);
  ctx->use_tma = 1;
  return VISOCU_OK;
}
