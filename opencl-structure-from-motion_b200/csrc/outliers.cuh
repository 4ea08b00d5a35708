// Device-side outlier removal (outliers.cu): job descriptor shared with the matcher's launch code.
#pragma once
#include "visocu_internal.cuh"

struct RoJob {
  const visocu_pmatch* in;     // match list as the matching / refinement kernels left it
  const uint8_t* keep_in;      // flags of the sub-pixel refinement (0 = dropped), or null
  const int32_t* n_in;         // number of records in `in` (device counter written by the emit kernel)
  visocu_pmatch* out;          // survivors, order preserved
  int32_t* result;             // [0] records in out, [1] status: 0 = outliers removed, else list unchanged and the host must
                               //     vote (1 = too long for shared memory, 2 = duplicate positions, 3 = internal guard)
                               // [2] edges allocated (diagnostics)
  int32_t* idx;                // scratch, n_in entries: list position -> record index
  int32_t* vert;               // scratch, n_in entries: mesh vertex -> list position
  const uint8_t* rep;          // optional, per record: 0 = duplicate position that Triangle ignores (ro_resolve_duplicates)
  uint16_t* hnd;               // scratch, 4 * (n_in / 2 + 2) entries: hull handles and free lists of the subtrees
  // Triangulation only (visocu_delaunay_subtrees): the vertices are given as points, x | y << 16, all distinct, instead of
  // match records; the job is a node of a larger divide-and-conquer tree whose cut is along axis0 (0 = vertical), and the
  // mesh is written out instead of being voted on, in the numbering of the whole problem: 2 * ro_edge_capacity(n_in)
  // half-edge records (onext, oprev, origin, x | y << 16 of the origin) starting at half-edge he_base, vertices numbered
  // from v_base in the order of the node's partition tree; deleted and unused entries have origin -1.  vert = mesh vertex
  // (local) -> index into pts_in; result[0] vertices, [1] status, [2] edges allocated, [4] and [5] the hull handles (ldo, rdo)
  const uint32_t* pts_in;
  int4* mesh_out;
  int32_t axis0, he_base, v_base;
};


// Positions of all records of a list as x << 16 | y (0xFFFFFFFF for records dropped by the sub-pixel refinement), for the
// host-side duplicate resolution below.  keys_dev: n_jobs rows of `stride` words.
int visocu_launch_outlier_keys(visocu_ctx* ctx, const RoJob* jobs_dev, int n_jobs, uint32_t* keys_dev, int stride);
// Several matches on one pixel (quad and stereo matching): Triangle triangulates one point per position, the one its
// randomised quicksort leaves first among the equal ones.  Replays that sort on the host (microseconds) and marks the
// records that take part: rep[i] = 1.  Returns false, leaving rep untouched, if all positions are distinct.
bool ro_resolve_duplicates(const uint32_t* keys, int n_records, uint8_t* rep);
// launches one CTA per job on the context's stream; jobs is a device array
int visocu_launch_remove_outliers(visocu_ctx* ctx, const RoJob* jobs_dev, int n_jobs, int method, int max_records, cudaStream_t stream);
