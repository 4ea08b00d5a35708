// Device-side outlier removal (outliers.cu): job descriptor shared with the matcher's launch code.
#pragma once
#include "visocu_internal.cuh"

struct RoJob {
  const visocu_pmatch* in;     // match list as the matching / refinement kernels left it
  const uint8_t* keep_in;      // flags of the sub-pixel refinement (0 = dropped), or null
  const int32_t* n_in;         // number of records in `in` (device counter written by the emit kernel)
  visocu_pmatch* out;          // survivors, order preserved
  int32_t* result;             // [0] records in out, [1] status: 0 = outliers removed, 1 = list unchanged, host must vote
                               // [2] edges allocated (diagnostics)
  int32_t* idx;                // scratch, n_in entries: list position -> record index
  int32_t* vert;               // scratch, n_in entries: mesh vertex -> list position
};

// launches one CTA per job on the context's stream; jobs is a device array
int visocu_launch_remove_outliers(visocu_ctx* ctx, const RoJob* jobs_dev, int n_jobs, int method, int max_records);
