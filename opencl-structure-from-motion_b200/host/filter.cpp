#include "filter.h"

#include <iostream>
#include <stdexcept>

#include "visocu.h"
#include "device.h"

namespace {
struct ThreadCtx {
  visocu_ctx* ctx = nullptr;
  int device = -1;
  ~ThreadCtx() { if (ctx) visocu_destroy(ctx); }
  visocu_ctx* get() {
    const int want = visob::current_device();
    if (ctx && device != want) { visocu_destroy(ctx); ctx = nullptr; }
    if (!ctx) {
      if (visocu_create(want, &ctx) != VISOCU_OK) throw std::runtime_error(std::string("filter: ") + visocu_last_error(nullptr));
      device = want;
    }
    return ctx;
  }
};
thread_local ThreadCtx t_ctx;

void check(int rc, visocu_ctx* c, const char* what) {
  if (rc != VISOCU_OK) throw std::runtime_error(std::string(what) + ": " + visocu_last_error(c));
}
}  // namespace

namespace filter {
void sobel3x3(const uint8_t* in, uint8_t* out_v, uint8_t* out_h, int w, int h) {
  visocu_ctx* c = t_ctx.get(); check(visocu_sobel3x3(c, in, out_v, out_h, w, h), c, "filter::sobel3x3");
}
void sobel5x5(const uint8_t* in, uint8_t* out_v, uint8_t* out_h, int w, int h) {
  visocu_ctx* c = t_ctx.get(); check(visocu_sobel5x5(c, in, out_v, out_h, w, h), c, "filter::sobel5x5");
}
void checkerboard5x5(const uint8_t* in, int16_t* out, int w, int h) {
  visocu_ctx* c = t_ctx.get(); check(visocu_checkerboard5x5(c, in, out, w, h), c, "filter::checkerboard5x5");
}
void blob5x5(const uint8_t* in, int16_t* out, int w, int h) {
  visocu_ctx* c = t_ctx.get(); check(visocu_blob5x5(c, in, out, w, h), c, "filter::blob5x5");
}
}  // namespace filter
