// Exact Delaunay triangulation of integer points (see delaunay.cpp).
#ifndef VISOB_DELAUNAY_H
#define VISOB_DELAUNAY_H
#include <stdint.h>
#include <vector>

struct visocu_ctx;

namespace visob {
// tri receives 3 indices (into x/y) per triangle, counter-clockwise.  Coordinates must satisfy |x|,|y| < 2^15.
// Of several points with identical coordinates only one is triangulated (the one the reference would keep).
void delaunay_triangles(const int32_t* x, const int32_t* y, int n, std::vector<int32_t>& tri);
// The same triangulation as an edge list: (a, b, t) per undirected edge, t = number of triangles (1 or 2) it bounds.
// This is all the support vote of removeOutliers needs and skips the face enumeration.
void delaunay_edges(const int32_t* x, const int32_t* y, int n, std::vector<int32_t>& edges);
// Calls of this thread from now on may build the lower part of large triangulations (more than 6000 distinct points) on
// the device behind ctx (visocu_delaunay_subtrees, on the context's current lane): the result is the same, the sorting,
// the top merges and the edge list stay here.  ctx = null: host only.
void delaunay_use_device(::visocu_ctx* ctx);
// nodes this thread has taken over from the device so far (tests: the device path was really used)
long delaunay_device_nodes();
}
#endif
