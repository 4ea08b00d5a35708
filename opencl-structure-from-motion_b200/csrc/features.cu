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
//   phase 1  stage the image tile plus halo in shared memory (word loads, zero outside the image)
//   phase 2  column walkers: 5-row register windows of the five horizontal sums -> du, dv (global, core only),
//            blob f1 and checkerboard f2 (shared memory only)
//   phase 3  one thread per owned NMS cell and pass: cell extrema (first in column-major order wins), then for
//            the extrema that pass the tau test a window scan for a strictly better value
// A cell is owned by the tile that contains its origin; the halo is nmax to the left/top and 2*nmax to the
// right/bottom because a cell spans [i,i+n] and its extremum's neighbourhood [i-n,i+2n] (matcher.cpp:356,383).
__global__ void __launch_bounds__(FILTER_THREADS) k_filter_nms(Geometry g, const FrameDev* frames, SlotList sl, int nmax) {
  extern __shared__ __align__(16) uint8_t smem[];
  const FrameDev F = frames[sl.s[blockIdx.z]];
  const uint8_t* __restrict__ I = g.half ? F.half : F.img;
  const int tid = threadIdx.x;
  const int x0 = blockIdx.x * TW, y0 = blockIdx.y * TH;
  const int fx0 = x0 - nmax, fy0 = y0 - nmax;              // origin of the response region
  const int FW = TW + 3 * nmax, FH = TH + 3 * nmax;
  const int FS = (FW + 1) & ~1;                            // int16 row stride of the response planes
  const int ix0 = (fx0 - 2) & ~3;                          // image tile origin, rounded down to a word
  const int iy0 = fy0 - 2;
  const int IH = FH + 4;
  const int IS = ((fx0 + FW + 2 - ix0) + 15) & ~15;        // byte row stride of the image tile
  uint8_t* simg = smem;
  int16_t* sf1 = (int16_t*)(smem + (((size_t)IH * IS + 15) & ~(size_t)15));
  int16_t* sf2 = sf1 + (size_t)FH * FS;

  // ---- phase 1
  {
    const int words = IS >> 2;
    for (int idx = tid; idx < IH * words; idx += FILTER_THREADS) {
      int ly = idx / words, lw = idx - ly * words;
      int gx = ix0 + 4 * lw, gy = iy0 + ly;
      uint32_t v = 0;
      if (gy >= 0 && gy < g.hm && gx >= 0 && gx < g.bplm) v = __ldg((const uint32_t*)(I + (size_t)gy * g.bplm + gx));
      *(uint32_t*)(simg + ly * IS + 4 * lw) = v;
    }
  }
  __syncthreads();

  // ---- phase 2
  {
    const int nseg = (FH + SEG - 1) / SEG;
    for (int item = tid; item < FW * nseg; item += FILTER_THREADS) {
      const int seg = item / FW, cx = item - seg * FW;
      const int ly0 = seg * SEG;
      const int nrow = min(SEG, FH - ly0);
      const int gx = fx0 + cx;
      const uint8_t* col = simg + (gx - 2 - ix0);          // p[0] of hrow for image-tile row r is col[r*IS]
      const bool core_x = gx >= x0 && gx < x0 + TW && gx < g.bplm;
      // columns w-2 .. bpl-3 are computed too, from the zero pad: the reference filters its whole 16-byte-stride
      // buffer and a descriptor of a maximum at u = w-7 samples column w-2 (its pad bytes are uninitialised heap
      // memory there, zero in the common case of a fresh allocation)
      const bool valid_x = gx >= 2 && gx <= g.bplm - 3;
      int hd[5], ha[5], h1[5], h3[5], hc[5], pc[5];
#pragma unroll
      for (int k = 0; k < 4; k++) {
        HRow r = hrow(col + (ly0 + k) * IS);
        hd[k] = r.hd; ha[k] = r.ha; h1[k] = r.h1; h3[k] = r.h3; hc[k] = r.hc; pc[k] = r.pc;
      }
      for (int r5 = 0; r5 < nrow; r5 += 5) {
#pragma unroll
        for (int ph = 0; ph < 5; ph++) {
          const int lr = r5 + ph;
          if (lr < nrow) {
            const int s0 = ph % 5, s1 = (ph + 1) % 5, s2 = (ph + 2) % 5, s3 = (ph + 3) % 5, s4 = (ph + 4) % 5;
            {
              HRow r = hrow(col + (ly0 + lr + 4) * IS);
              hd[s4] = r.hd; ha[s4] = r.ha; h1[s4] = r.h1; h3[s4] = r.h3; hc[s4] = r.hc; pc[s4] = r.pc;
            }
            const int ly = ly0 + lr, gy = fy0 + ly;
            int f1 = -(h1[s0] + h1[s1] + h1[s2] + h1[s3] + h1[s4]) + 2 * (h3[s1] + h3[s2] + h3[s3]) + 7 * pc[s2];
            int f2 = (hc[s0] + hc[s1]) - (hc[s3] + hc[s4]);
            sf1[ly * FS + cx] = (int16_t)f1;
            sf2[ly * FS + cx] = (int16_t)f2;
            if (core_x && gy >= y0 && gy < y0 + TH && gy < g.hm) {
              int du = 128, dv = 128;
              if (valid_x && gy >= 2 && gy <= g.hm - 3) {
                du = (((hd[s0] + hd[s4]) + 4 * (hd[s1] + hd[s3]) + 6 * hd[s2]) >> 7) + 128;   // |sum| <= 12240: no saturation
                dv = (((ha[s0] - ha[s4]) + 2 * (ha[s1] - ha[s3])) >> 7) + 128;
              }
              F.du[(size_t)gy * g.bplm + gx] = (uint8_t)du;
              F.dv[(size_t)gy * g.bplm + gx] = (uint8_t)dv;
            }
          }
        }
      }
    }
  }
  __syncthreads();

  // ---- phase 3
  {
    int klo[2], nk[2], llo[2], nl[2], total = 0, first[2];
    for (int p = 0; p < 2; p++) {
      nk[p] = nl[p] = klo[p] = llo[p] = 0; first[p] = total;
      if (p < g.first_pass) continue;
      const int n = g.n[p], step = n + 1, org = n + VISO_MARGIN;
      int a = x0 - org; klo[p] = a > 0 ? (a + step - 1) / step : 0;
      int b = x0 + TW - org; int khi = b > 0 ? min((b + step - 1) / step, g.ncx[p]) : 0;
      a = y0 - org; llo[p] = a > 0 ? (a + step - 1) / step : 0;
      b = y0 + TH - org; int lhi = b > 0 ? min((b + step - 1) / step, g.ncy[p]) : 0;
      nk[p] = max(khi - klo[p], 0); nl[p] = max(lhi - llo[p], 0);
      total += nk[p] * nl[p];
    }
    const int xhi = g.wm - 1 - VISO_MARGIN - fx0, yhi = g.hm - 1 - VISO_MARGIN - fy0;   // window clamps, region-local
    for (int item = tid; item < total; item += FILTER_THREADS) {
      const int p = (item >= first[1] && nk[1] * nl[1] > 0) ? 1 : 0;
      const int loc = item - first[p];
      const int kk = loc % nk[p], ll = loc / nk[p];
      const int k = klo[p] + kk, l = llo[p] + ll;
      const int n = g.n[p];
      const int lx = n + VISO_MARGIN + k * (n + 1) - fx0, ly = n + VISO_MARGIN + l * (n + 1) - fy0;
      const int16_t* q1 = sf1 + ly * FS + lx;
      const int16_t* q2 = sf2 + ly * FS + lx;
      int f1min = q1[0], f1max = f1min, f2min = q2[0], f2max = f2min;
      int p1min = 0, p1max = 0, p2min = 0, p2max = 0;      // (di<<4 | dj)
      for (int di = 0; di <= n; di++) {
        for (int dj = 0; dj <= n; dj++) {
          int v = q1[dj * FS + di], pos = (di << 4) | dj;
          if (v < f1min) { f1min = v; p1min = pos; } else if (v > f1max) { f1max = v; p1max = pos; }
          v = q2[dj * FS + di];
          if (v < f2min) { f2min = v; p2min = pos; } else if (v > f2max) { f2max = v; p2max = pos; }
        }
      }
      // keep an extremum iff nothing in its clamped (2n+1)^2 window is strictly better; positions inside the
      // cell can never be strictly better than the cell extremum, so the reference's cell exclusion is implied
      auto keep = [&](const int16_t* sf, int pos, int val, bool is_min) -> bool {
        const int ex = lx + (pos >> 4), ey = ly + (pos & 15);
        const int xe = min(ex + n, xhi), ye = min(ey + n, yhi);
        for (int i2 = ex - n; i2 <= xe; i2++)
          for (int j2 = ey - n; j2 <= ye; j2++) {
            int v = sf[j2 * FS + i2];
            if (is_min ? (v < val) : (v > val)) return false;
          }
        return true;
      };
      uint32_t code = 0xFFFFFFFFu;
      if (f1min <= -g.tau && keep(sf1, p1min, f1min, true))  code = (code & 0xFFFFFF00u) | (uint32_t)p1min;
      if (f1max >= g.tau  && keep(sf1, p1max, f1max, false)) code = (code & 0xFFFF00FFu) | ((uint32_t)p1max << 8);
      if (f2min <= -g.tau && keep(sf2, p2min, f2min, true))  code = (code & 0xFF00FFFFu) | ((uint32_t)p2min << 16);
      if (f2max >= g.tau  && keep(sf2, p2max, f2max, false)) code = (code & 0x00FFFFFFu) | ((uint32_t)p2max << 24);
      F.codes[p][(size_t)k * g.ncy[p] + l] = code;
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
    const int FW = TW + 3 * nmax, FH = TH + 3 * nmax, FS = (FW + 1) & ~1;
    const int IS = (FW + 4 + 3 + 15) & ~15, IH = FH + 4;
    size_t smem = (((size_t)IH * IS + 15) & ~(size_t)15) + 2 * (size_t)FH * FS * sizeof(int16_t);
    if (smem > ctx->filter_smem_attr) {
      CU_TRY(ctx, cudaFuncSetAttribute(k_filter_nms, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      ctx->filter_smem_attr = smem;
    }
    dim3 grid((g.bplm + TW - 1) / TW, (g.hm + TH - 1) / TH, sl.n);
    if (ctx->profile) CU_TRY(ctx, cudaEventRecord(ctx->pev0, st));
    k_filter_nms<<<grid, FILTER_THREADS, smem, st>>>(g, ctx->frames_d, sl, nmax);
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
  return VISOCU_OK;
}
