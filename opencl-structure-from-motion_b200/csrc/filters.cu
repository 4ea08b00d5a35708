// Stand-alone filter:: entry points (reference viso/filter.h:78-94) and Matcher::nonMaximumSuppression on
// caller-supplied response maps.  These keep the reference's signatures for callers that use the filters on
// their own; the Matcher path never goes through here (it uses the fused kernel in features.cu).
//
// Output contract: identical to the reference wherever the reference's result does not depend on its
// flat-array row pass wrapping across line ends or on reads past the buffer:
//   sobel:        2 <= x <= w-3, 2 <= y <= h-3; everything else is 128
//   blob:         3 <= x <= w-3, 3 <= y <= h-3; everything else is 0
//   checkerboard: 2 <= x <= w-3, 2 <= y <= h-3; everything else is 0
#include "visocu_internal.cuh"

namespace {

__device__ __forceinline__ int px(const uint8_t* I, int w, int x, int y) { return I[(size_t)y * w + x]; }

// mode 0: sobel5x5 -> (o8a, o8b); 1: sobel3x3 -> (o8a, o8b); 2: blob5x5 -> o16; 3: checkerboard5x5 -> o16
__global__ void k_filter_plain(const uint8_t* __restrict__ I, uint8_t* o8a, uint8_t* o8b, int16_t* o16, int w, int h, int mode) {
  int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= w || y >= h) return;
  size_t o = (size_t)y * w + x;
  if (mode == 0) {
    int du = 128, dv = 128;
    if (x >= 2 && x <= w - 3 && y >= 2 && y <= h - 3) {
      const int a[5] = {1, 4, 6, 4, 1}, d[5] = {1, 2, 0, -2, -1};
      int su = 0, sv = 0;
      for (int j = 0; j < 5; j++)
        for (int i = 0; i < 5; i++) {
          int p = px(I, w, x - 2 + i, y - 2 + j);
          su += a[j] * d[i] * p;
          sv += d[j] * a[i] * p;
        }
      du = min(max((su >> 7) + 128, 0), 255);
      dv = min(max((sv >> 7) + 128, 0), 255);
    }
    o8a[o] = (uint8_t)du; o8b[o] = (uint8_t)dv;
  } else if (mode == 1) {
    int du = 128, dv = 128;
    if (x >= 1 && x <= w - 2 && y >= 1 && y <= h - 2) {
      const int a[3] = {1, 2, 1}, d[3] = {1, 0, -1};
      int su = 0, sv = 0;
      for (int j = 0; j < 3; j++)
        for (int i = 0; i < 3; i++) {
          int p = px(I, w, x - 1 + i, y - 1 + j);
          su += a[j] * d[i] * p;
          sv += d[j] * a[i] * p;
        }
      du = min(max((su >> 2) + 128, 0), 255);
      dv = min(max((sv >> 2) + 128, 0), 255);
    }
    o8a[o] = (uint8_t)du; o8b[o] = (uint8_t)dv;
  } else if (mode == 2) {
    int f = 0;
    if (x >= 3 && x <= w - 3 && y >= 3 && y <= h - 3) {
      for (int j = -2; j <= 2; j++)
        for (int i = -2; i <= 2; i++) {
          int p = px(I, w, x + i, y + j);
          int ring = max(abs(i), abs(j));
          f += ring == 2 ? -p : (ring == 1 ? p : 8 * p);
        }
    }
    o16[o] = (int16_t)f;
  } else {
    int f = 0;
    if (x >= 2 && x <= w - 3 && y >= 2 && y <= h - 3) {
      const int c[5] = {1, 1, 0, -1, -1};
      for (int j = 0; j < 5; j++)
        for (int i = 0; i < 5; i++) f += c[j] * c[i] * px(I, w, x - 2 + i, y - 2 + j);
    }
    o16[o] = (int16_t)f;
  }
}

// one thread per NMS cell on global response maps; same code word as the fused kernel
__global__ void k_nms_plain(const int16_t* __restrict__ f1, const int16_t* __restrict__ f2, int w, int h, int bpl, int n, int tau,
                            int ncx, int ncy, uint32_t* codes) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= ncx * ncy) return;
  int k = c / ncy, l = c - k * ncy;
  int i = n + VISO_MARGIN + k * (n + 1), j = n + VISO_MARGIN + l * (n + 1);
  auto at = [&](const int16_t* f, int x, int y) { return (int)f[(size_t)y * bpl + x]; };
  int f1min = at(f1, i, j), f1max = f1min, f2min = at(f2, i, j), f2max = f2min, p1min = 0, p1max = 0, p2min = 0, p2max = 0;
  for (int di = 0; di <= n; di++)
    for (int dj = 0; dj <= n; dj++) {
      int v = at(f1, i + di, j + dj), pos = (di << 4) | dj;
      if (v < f1min) { f1min = v; p1min = pos; } else if (v > f1max) { f1max = v; p1max = pos; }
      v = at(f2, i + di, j + dj);
      if (v < f2min) { f2min = v; p2min = pos; } else if (v > f2max) { f2max = v; p2max = pos; }
    }
  auto keep = [&](const int16_t* f, int pos, int val, bool is_min) {
    int ex = i + (pos >> 4), ey = j + (pos & 15);
    int xe = min(ex + n, w - 1 - VISO_MARGIN), ye = min(ey + n, h - 1 - VISO_MARGIN);
    for (int x = ex - n; x <= xe; x++)
      for (int y = ey - n; y <= ye; y++) {
        int v = at(f, x, y);
        if (is_min ? v < val : v > val) return false;
      }
    return true;
  };
  uint32_t code = 0xFFFFFFFFu;
  if (f1min <= -tau && keep(f1, p1min, f1min, true))  code = (code & 0xFFFFFF00u) | (uint32_t)p1min;
  if (f1max >= tau  && keep(f1, p1max, f1max, false)) code = (code & 0xFFFF00FFu) | ((uint32_t)p1max << 8);
  if (f2min <= -tau && keep(f2, p2min, f2min, true))  code = (code & 0xFF00FFFFu) | ((uint32_t)p2min << 16);
  if (f2max >= tau  && keep(f2, p2max, f2max, false)) code = (code & 0x00FFFFFFu) | ((uint32_t)p2max << 24);
  codes[c] = code;
}

int run_filter(visocu_ctx* ctx, const uint8_t* in, uint8_t* oa, uint8_t* ob, int16_t* o16, int w, int h, int mode) {
  if (!ctx || !in || w <= 0 || h <= 0) return ctx ? visocu_set_error(ctx, VISOCU_EINVAL, "bad filter arguments") : VISOCU_EINVAL;
  CU_TRY(ctx, cudaSetDevice(ctx->device));
  size_t n = (size_t)w * h, na = align_up(n, 256);
  int rc = visocu_ensure_scratch(ctx, na * 5);
  if (rc) return rc;
  uint8_t* d_in = (uint8_t*)ctx->scratch;
  uint8_t* d_a = d_in + na;
  uint8_t* d_b = d_a + na;
  int16_t* d_16 = (int16_t*)(d_b + na);
  CU_COPY(ctx, d_in, in, n, cudaMemcpyHostToDevice);
  dim3 b(32, 8), g((w + 31) / 32, (h + 7) / 8);
  k_filter_plain<<<g, b, 0, ctx->stream>>>(d_in, d_a, d_b, d_16, w, h, mode);
  CU_LAUNCH_CHECK(ctx);
  if (mode <= 1) {
    CU_COPY(ctx, oa, d_a, n, cudaMemcpyDeviceToHost);
    CU_COPY(ctx, ob, d_b, n, cudaMemcpyDeviceToHost);
  } else {
    CU_COPY(ctx, o16, d_16, n * 2, cudaMemcpyDeviceToHost);
  }
  CU_TRY(ctx, visocu_stream_wait(ctx));
  return VISOCU_OK;
}

}  // namespace

extern "C" int visocu_sobel5x5(visocu_ctx* ctx, const uint8_t* in, uint8_t* out_v, uint8_t* out_h, int32_t w, int32_t h) {
  if (!out_v || !out_h) return VISOCU_EINVAL;
  return run_filter(ctx, in, out_v, out_h, nullptr, w, h, 0);
}
extern "C" int visocu_sobel3x3(visocu_ctx* ctx, const uint8_t* in, uint8_t* out_v, uint8_t* out_h, int32_t w, int32_t h) {
  if (!out_v || !out_h) return VISOCU_EINVAL;
  return run_filter(ctx, in, out_v, out_h, nullptr, w, h, 1);
}
extern "C" int visocu_blob5x5(visocu_ctx* ctx, const uint8_t* in, int16_t* out, int32_t w, int32_t h) {
  if (!out) return VISOCU_EINVAL;
  return run_filter(ctx, in, nullptr, nullptr, out, w, h, 2);
}
extern "C" int visocu_checkerboard5x5(visocu_ctx* ctx, const uint8_t* in, int16_t* out, int32_t w, int32_t h) {
  if (!out) return VISOCU_EINVAL;
  return run_filter(ctx, in, nullptr, nullptr, out, w, h, 3);
}

extern "C" int visocu_nms(visocu_ctx* ctx, const int16_t* f1, const int16_t* f2, int32_t w, int32_t h, int32_t bpl,
                          int32_t nms_n, int32_t tau, int32_t* out4, int32_t cap, int32_t* n_out) {
  if (!ctx || !f1 || !f2 || !n_out) return VISOCU_EINVAL;
  if (w <= 0 || h <= 0 || bpl < w || nms_n < 1 || nms_n > 14) return visocu_set_error(ctx, VISOCU_EINVAL, "bad nms arguments");
  CU_TRY(ctx, cudaSetDevice(ctx->device));
  const int ncx = viso_cell_count(w, nms_n), ncy = viso_cell_count(h, nms_n);
  const size_t cells = (size_t)ncx * ncy;
  *n_out = 0;
  if (cells == 0) return VISOCU_OK;
  size_t plane = align_up((size_t)bpl * h * 2, 256);
  int rc = visocu_ensure_scratch(ctx, 2 * plane + cells * 4);
  if (rc) return rc;
  int16_t* d1 = (int16_t*)ctx->scratch;
  int16_t* d2 = (int16_t*)((uint8_t*)ctx->scratch + plane);
  uint32_t* dc = (uint32_t*)((uint8_t*)ctx->scratch + 2 * plane);
  CU_COPY(ctx, d1, f1, (size_t)bpl * h * 2, cudaMemcpyHostToDevice);
  CU_COPY(ctx, d2, f2, (size_t)bpl * h * 2, cudaMemcpyHostToDevice);
  k_nms_plain<<<(unsigned)((cells + 127) / 128), 128, 0, ctx->stream>>>(d1, d2, w, h, bpl, nms_n, tau, ncx, ncy, dc);
  CU_LAUNCH_CHECK(ctx);
  std::vector<uint32_t> codes(cells);
  CU_COPY(ctx, codes.data(), dc, cells * 4, cudaMemcpyDeviceToHost);
  CU_TRY(ctx, visocu_stream_wait(ctx));
  // expand the per-cell code words into the reference's (u, v, val, class) list, cell-column-major
  int n = 0;
  for (size_t c = 0; c < cells; c++) {
    uint32_t code = codes[c];
    if (code == 0xFFFFFFFFu) continue;
    int k = (int)(c / ncy), l = (int)(c % ncy);
    int i = nms_n + VISO_MARGIN + k * (nms_n + 1), j = nms_n + VISO_MARGIN + l * (nms_n + 1);
    for (int cls = 0; cls < 4; cls++) {
      uint32_t b = (code >> (8 * cls)) & 0xFFu;
      if (b == 0xFFu) continue;
      int u = i + (int)(b >> 4), v = j + (int)(b & 15);
      if (out4 && n < cap) {
        out4[4 * n + 0] = u; out4[4 * n + 1] = v;
        out4[4 * n + 2] = (cls < 2 ? f1 : f2)[(size_t)v * bpl + u];
        out4[4 * n + 3] = cls;
      }
      n++;
    }
  }
  *n_out = n;
  if (out4 && n > cap) return visocu_set_error(ctx, VISOCU_ECAPACITY, "need room for %d maxima", n);
  return VISOCU_OK;
}
