// Delaunay triangulation of integer pixel coordinates for Matcher::removeOutliers (reference matcher.cpp:1207-1377,
// which calls the vendored Triangle library with switches "zQB", matcher.cpp:1255-1256).
//
// This is an independent implementation: Guibas-Stolfi divide and conquer on a primal-only quad-edge structure
// (onext/oprev rings), with the alternating vertical/horizontal cuts of Dwyer, and exact 64-bit integer orientation
// and in-circle predicates (the inputs are pixel coordinates < 2^15, so no adaptive floating-point arithmetic is
// needed).  Pixel grids are full of co-circular point quadruples, for which the Delaunay triangulation is not
// unique; the support vote of removeOutliers depends on which diagonal is chosen.  To stay bit-compatible with the
// reference's match lists, every tie is decided the way Triangle's divide-and-conquer code decides it: same split
// sizes (n/2 by alternating axis, leaves of 2-3 vertices sorted by x), both tangent tests per iteration of the
// lower-tangent search, strict `> 0` in-circle tests, candidate validity evaluated before the candidate-removal
// loops, and the same walks that re-seat the hull handles on the y-extremes before, and on the x-extremes after, a
// merge across a horizontal cut.  tests/test_host_delaunay.py checks the resulting outlier decisions against the
// reference on grid-heavy random inputs.
#include "delaunay.h"

#include <algorithm>
#include <cstdint>

namespace visob {
namespace {

struct Mesh {
  const int32_t* px;
  const int32_t* py;
  std::vector<int32_t> onext, oprev, org;   // per directed edge; sym(e) = e ^ 1
  std::vector<uint8_t> dead;

  int make_edge(int a, int b) {
    int e = (int)org.size();
    org.push_back(a); org.push_back(b);
    onext.push_back(e); onext.push_back(e + 1);
    oprev.push_back(e); oprev.push_back(e + 1);
    dead.push_back(0); dead.push_back(0);
    return e;
  }
  static int sym(int e) { return e ^ 1; }
  int dest(int e) const { return org[e ^ 1]; }
  int lnext(int e) const { return oprev[e ^ 1]; }
  int rprev(int e) const { return onext[e ^ 1]; }
  // put the isolated half-edge e right after x in the counter-clockwise ring around x's origin
  void insert_after(int x, int e) {
    int n = onext[x];
    onext[e] = n; oprev[e] = x;
    oprev[n] = e; onext[x] = e;
  }
  void unlink(int e) {
    int n = onext[e], p = oprev[e];
    onext[p] = n; oprev[n] = p;
    onext[e] = oprev[e] = e;
  }
  // new edge from dest(a) to org(b) so that a, the new edge and b share their left face
  int connect(int a, int b) {
    int e = make_edge(dest(a), org[b]);
    insert_after(lnext(a), e);
    insert_after(b, sym(e));
    return e;
  }
  void remove(int e) {
    unlink(e); unlink(sym(e));
    dead[e] = dead[e ^ 1] = 1;
  }
  // > 0 iff a, b, c make a left turn
  int64_t ccw(int a, int b, int c) const {
    return (int64_t)(px[a] - px[c]) * (py[b] - py[c]) - (int64_t)(py[a] - py[c]) * (px[b] - px[c]);
  }
  // > 0 iff d lies strictly inside the circle through a, b, c (a, b, c counter-clockwise)
  int64_t incircle(int a, int b, int c, int d) const {
    const int64_t adx = px[a] - px[d], ady = py[a] - py[d];
    const int64_t bdx = px[b] - px[d], bdy = py[b] - py[d];
    const int64_t cdx = px[c] - px[d], cdy = py[c] - py[d];
    const int64_t al = adx * adx + ady * ady, bl = bdx * bdx + bdy * bdy, cl = cdx * cdx + cdy * cdy;
    return al * (bdx * cdy - cdx * bdy) + bl * (cdx * ady - adx * cdy) + cl * (adx * bdy - bdx * ady);
  }
};

// x(n+1) = (1366 x(n) + 150889) mod 714025, scaled to [0, choices)
struct PivotSource {
  unsigned long long seed = 1;
  int next(unsigned choices) {
    seed = (seed * 1366ull + 150889ull) % 714025ull;
    return (int)(seed / (714025ull / choices + 1));
  }
};

bool before_xy(const Mesh& m, int a, int px, int py) { return m.px[a] < px || (m.px[a] == px && m.py[a] < py); }
bool after_xy(const Mesh& m, int a, int px, int py) { return m.px[a] > px || (m.px[a] == px && m.py[a] > py); }

void pivot_quicksort(const Mesh& m, int* a, int n, PivotSource& rnd) {
  if (n < 2) return;
  if (n == 2) {
    if (after_xy(m, a[0], m.px[a[1]], m.py[a[1]])) std::swap(a[0], a[1]);
    return;
  }
  const int pv = a[rnd.next((unsigned)n)];
  const int px = m.px[pv], py = m.py[pv];
  int left = -1, right = n;
  while (left < right) {
    do { left++; } while (left <= right && before_xy(m, a[left], px, py));
    do { right--; } while (left <= right && after_xy(m, a[right], px, py));
    if (left < right) std::swap(a[left], a[right]);
  }
  if (left > 1) pivot_quicksort(m, a, left, rnd);
  if (right < n - 2) pivot_quicksort(m, a + right + 1, n - right - 1, rnd);
}

struct Handles { int ldo, rdo; };   // ccw hull edge out of the leftmost vertex, cw hull edge out of the rightmost

bool less_axis(const Mesh& m, int a, int b, int axis) {
  if (axis == 0) return m.px[a] != m.px[b] ? m.px[a] < m.px[b] : m.py[a] < m.py[b];
  return m.py[a] != m.py[b] ? m.py[a] < m.py[b] : m.px[a] < m.px[b];
}

// The partition tree: halves by the node's axis, children by the other axis, leaves (<= 3 vertices) sorted by x.
void partition(const Mesh& m, int* v, int n, int axis) {
  if (n <= 3) axis = 0;
  const int divider = n >> 1;
  if (n <= 3) {
    std::sort(v, v + n, [&](int a, int b) { return less_axis(m, a, b, 0); });
    return;
  }
  std::nth_element(v, v + divider, v + n, [&](int a, int b) { return less_axis(m, a, b, axis); });
  partition(m, v, divider, 1 - axis);
  partition(m, v + divider, n - divider, 1 - axis);
}

Handles merge(Mesh& m, Handles L, Handles R, int axis) {
  int ldo = L.ldo, ldi = L.rdo, rdi = R.ldo, rdo = R.rdo;
  if (axis == 1) {
    // the two sets are separated by a horizontal line: seat the handles on the y-extremes
    while (m.py[m.dest(ldo)] < m.py[m.org[ldo]]) ldo = m.rprev(ldo);
    while (m.py[m.dest(m.onext[ldi])] > m.py[m.org[ldi]]) ldi = Mesh::sym(m.onext[ldi]);
    while (m.py[m.dest(rdi)] < m.py[m.org[rdi]]) rdi = m.rprev(rdi);
    while (m.py[m.dest(m.onext[rdo])] > m.py[m.org[rdo]]) rdo = Mesh::sym(m.onext[rdo]);
  }
  // lower common tangent
  bool changed;
  do {
    changed = false;
    if (m.ccw(m.org[ldi], m.dest(ldi), m.org[rdi]) > 0) { ldi = m.lnext(ldi); changed = true; }
    if (m.ccw(m.dest(rdi), m.org[rdi], m.org[ldi]) > 0) { rdi = m.rprev(rdi); changed = true; }
  } while (changed);
  int basel = m.connect(Mesh::sym(rdi), ldi);       // from org(rdi) = lower right to org(ldi) = lower left
  if (m.org[ldi] == m.org[ldo]) ldo = Mesh::sym(basel);
  if (m.org[rdi] == m.org[rdo]) rdo = basel;
  // knit the seam upwards
  for (;;) {
    const int lowerright = m.org[basel], lowerleft = m.dest(basel);
    int lcand = m.onext[Mesh::sym(basel)], rcand = m.oprev[basel];
    int upperleft = m.dest(lcand), upperright = m.dest(rcand);
    const bool leftfinished = m.ccw(upperleft, lowerleft, lowerright) <= 0;
    const bool rightfinished = m.ccw(upperright, lowerleft, lowerright) <= 0;
    if (leftfinished && rightfinished) break;
    if (!leftfinished) {
      for (;;) {
        const int nx = m.onext[lcand];
        if (nx == Mesh::sym(basel)) break;
        const int apex = m.dest(nx);
        if (m.ccw(lowerleft, upperleft, apex) <= 0) break;                 // no real triangle beyond lcand
        if (m.incircle(lowerleft, lowerright, upperleft, apex) <= 0) break;
        m.remove(lcand);
        lcand = nx; upperleft = apex;
      }
    }
    if (!rightfinished) {
      for (;;) {
        const int nx = m.oprev[rcand];
        if (nx == basel) break;
        const int apex = m.dest(nx);
        if (m.ccw(lowerright, apex, upperright) <= 0) break;
        if (m.incircle(lowerleft, lowerright, upperright, apex) <= 0) break;
        m.remove(rcand);
        rcand = nx; upperright = apex;
      }
    }
    if (leftfinished || (!rightfinished && m.incircle(upperleft, lowerleft, lowerright, upperright) > 0))
      basel = m.connect(rcand, Mesh::sym(basel));      // new base: upper right -> lower left
    else
      basel = m.connect(Mesh::sym(basel), Mesh::sym(lcand));   // new base: lower right -> upper left
  }
  if (axis == 1) {
    // back to the x-extremes expected by the parent (vertical cut) and by the leaves
    while (m.px[m.dest(m.oprev[ldo])] < m.px[m.org[ldo]]) ldo = Mesh::sym(m.oprev[ldo]);
    while (m.px[m.dest(rdo)] > m.px[m.org[rdo]]) rdo = m.lnext(rdo);
  }
  return Handles{ldo, rdo};
}

Handles build(Mesh& m, const int* v, int n, int axis) {
  if (n == 2) {
    int a = m.make_edge(v[0], v[1]);
    return Handles{a, Mesh::sym(a)};
  }
  if (n == 3) {
    int a = m.make_edge(v[0], v[1]);
    int b = m.make_edge(v[1], v[2]);
    m.insert_after(b, Mesh::sym(a));                 // ring around v[1]: b, sym(a)
    const int64_t area = m.ccw(v[0], v[1], v[2]);
    if (area == 0) return Handles{a, Mesh::sym(b)};
    int c = m.connect(b, a);                         // v[2] -> v[0]
    if (area > 0) return Handles{a, Mesh::sym(b)};
    return Handles{Mesh::sym(c), c};
  }
  const int divider = n >> 1;
  Handles L = build(m, v, divider, 1 - axis);
  Handles R = build(m, v + divider, n - divider, 1 - axis);
  return merge(m, L, R, axis);
}

}  // namespace

void delaunay_triangles(const int32_t* x, const int32_t* y, int n, std::vector<int32_t>& tri) {
  tri.clear();
  if (n < 3) return;
  Mesh m;
  m.px = x; m.py = y;
  m.org.reserve(8 * (size_t)n); m.onext.reserve(8 * (size_t)n); m.oprev.reserve(8 * (size_t)n); m.dead.reserve(8 * (size_t)n);
  // sort by (x, y); of several input points on the same pixel only one takes part (Triangle ignores duplicates).
  // Which of the duplicates survives depends on the order an unstable sort leaves them in, and removeOutliers
  // keeps exactly the survivor, so the sort is the same randomised quicksort (Hoare partition, pivots from the
  // generator seeded with 1) that the reference's triangulation uses.
  std::vector<int> v(n);
  for (int i = 0; i < n; i++) v[i] = i;
  PivotSource pivots;
  pivot_quicksort(m, v.data(), n, pivots);
  int k = 0;
  for (int i = 1; i < n; i++)
    if (x[v[i]] != x[v[k]] || y[v[i]] != y[v[k]]) v[++k] = v[i];
  const int nu = k + 1;
  if (nu < 3) return;
  // root: vertical cut of the x-sorted list, children by alternating axes
  const int divider = nu >> 1;
  if (nu - divider >= 2) {
    if (divider >= 2) partition(m, v.data(), divider, 1);
    partition(m, v.data() + divider, nu - divider, 1);
  }
  build(m, v.data(), nu, 0);
  // every bounded face is a counter-clockwise triangle; report each once, from its lowest-numbered half-edge
  const int ne = (int)m.org.size();
  for (int e = 0; e < ne; e++) {
    if (m.dead[e]) continue;
    const int e2 = m.lnext(e), e3 = m.lnext(e2);
    if (m.lnext(e3) != e || e2 < e || e3 < e) continue;
    const int a = m.org[e], b = m.org[e2], c = m.org[e3];
    if (m.ccw(a, b, c) > 0) { tri.push_back(a); tri.push_back(b); tri.push_back(c); }
  }
}

}  // namespace visob
