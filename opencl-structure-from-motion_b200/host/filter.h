// filter:: functions with the reference's signatures (viso/filter.h:78-94).  They run on the GPU through
// visocu_sobel5x5 / ... (include/visocu.h) on a lazily created per-thread context; there is no CPU path.
// Output contract: see csrc/filters.cu (identical to the reference on the region that does not depend on the
// reference's flat-array wrap-around; borders hold 128 / 0).
#ifndef VISOB_FILTER_H
#define VISOB_FILTER_H
#include <stdint.h>

namespace filter {
void sobel3x3(const uint8_t* in, uint8_t* out_v, uint8_t* out_h, int w, int h);
void sobel5x5(const uint8_t* in, uint8_t* out_v, uint8_t* out_h, int w, int h);
void checkerboard5x5(const uint8_t* in, int16_t* out, int w, int h);
void blob5x5(const uint8_t* in, int16_t* out, int w, int h);
}
#endif
