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
// merge across a horizontal cut.  tests/test_host_cpu.py::test_remove_outliers_matches_reference checks the resulting
// outlier decisions against the reference on grid-heavy random inputs (duplicates and lists beyond the device limit included).
#include "delaunay.h"
#include "visocu.h"

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

namespace visob {
namespace {

// xy: coordinates of the origin (x | y << 16, relative to the bounding box), so that the predicates need no second,
// dependent load through the vertex number
struct HalfEdge { int32_t onext, oprev, org; uint32_t xy; };

struct Pt { int32_t x, y; };

struct Mesh {
  const Pt* pt;
  HalfEdge* he;                             // per directed edge; sym(e) = e ^ 1 (storage owned by the caller)
  int nhe, cap;
  std::vector<HalfEdge> store;
  std::vector<uint8_t> flag;                // per half-edge: 1 = removed, 2 = left face is the unbounded face (records taken over
                                            // from the device mark removed and unused entries with org = -1 instead)
  void set_points(const Pt* p) { pt = p; }
  void reset(size_t want) {
    if (store.size() < want) store.resize(want);
    he = store.data(); cap = (int)store.size(); nhe = 0;
    flag.assign(store.size(), 0);
  }
  uint32_t packed(int v) const { return (uint32_t)pt[v].x | ((uint32_t)pt[v].y << 16); }

  // records that live elsewhere (the mesh the device delivered): used of them are taken, up to capacity may be written
  void adopt(HalfEdge* records, int used, int capacity) {
    he = records; nhe = used; cap = capacity;
    flag.assign((size_t)capacity, 0);
  }
  void grow() {                             // cannot happen with the 12 n reservation (3 n live + deleted edges); grow anyway
    std::vector<HalfEdge> bigger((size_t)cap * 2 + 64);
    std::copy(he, he + nhe, bigger.begin());
    store.swap(bigger);
    he = store.data(); cap = (int)store.size();
    flag.resize(store.size(), 0);
  }
  int make_edge(int a, int b) {
    if (nhe + 2 > cap) grow();
    const int e = nhe;
    he[e] = HalfEdge{e, e, a, packed(a)};
    he[e + 1] = HalfEdge{e + 1, e + 1, b, packed(b)};
    nhe += 2;
    return e;
  }
  // edge between the origins of half-edges ea and eb (numbers and coordinates copied from them)
  int make_edge_like(int ea, int eb) {
    if (nhe + 2 > cap) grow();
    const int e = nhe;
    he[e] = HalfEdge{e, e, he[ea].org, he[ea].xy};
    he[e + 1] = HalfEdge{e + 1, e + 1, he[eb].org, he[eb].xy};
    nhe += 2;
    return e;
  }
  uint32_t xy(int e) const { return he[e].xy; }          // coordinates of the origin of e
  uint32_t dxy(int e) const { return he[e ^ 1].xy; }     // ... of its destination
  static int sym(int e) { return e ^ 1; }
  int org(int e) const { return he[e].org; }
  int dest(int e) const { return he[e ^ 1].org; }
  int onext(int e) const { return he[e].onext; }
  int oprev(int e) const { return he[e].oprev; }
  int lnext(int e) const { return he[e ^ 1].oprev; }
  int rprev(int e) const { return he[e ^ 1].onext; }
  // put the isolated half-edge e right after x in the counter-clockwise ring around x's origin
  void insert_after(int x, int e) {
    const int n = he[x].onext;
    he[e].onext = n; he[e].oprev = x;
    he[n].oprev = e; he[x].onext = e;
  }
  void unlink(int e) {
    const int n = he[e].onext, p = he[e].oprev;
    he[p].onext = n; he[n].oprev = p;
    he[e].onext = he[e].oprev = e;
  }
  // new edge from dest(a) to org(b) so that a, the new edge and b share their left face
  int connect(int a, int b) {
    int e = make_edge_like(a ^ 1, b);
    insert_after(lnext(a), e);
    insert_after(b, sym(e));
    return e;
  }
  void remove(int e) {
    unlink(e); unlink(sym(e));
    flag[e] = flag[e ^ 1] = 1;
  }
  // > 0 iff a, b, c make a left turn (vertex numbers)
  int64_t ccw(int a, int b, int c) const {
    const Pt A = pt[a], B = pt[b], C = pt[c];
    return (int64_t)(A.x - C.x) * (B.y - C.y) - (int64_t)(A.y - C.y) * (B.x - C.x);
  }
  // the same on packed coordinates
  static int64_t ccw_xy(uint32_t A, uint32_t B, uint32_t C) {
    const int cx = (int)(C & 0xFFFF), cy = (int)(C >> 16);
    const int64_t ax = (int)(A & 0xFFFF) - cx, ay = (int)(A >> 16) - cy;
    const int64_t bx = (int)(B & 0xFFFF) - cx, by = (int)(B >> 16) - cy;
    return ax * by - ay * bx;
  }
  static int64_t incircle_xy(uint32_t A, uint32_t B, uint32_t C, uint32_t D) {
    const int dx = (int)(D & 0xFFFF), dy = (int)(D >> 16);
    const int64_t adx = (int)(A & 0xFFFF) - dx, ady = (int)(A >> 16) - dy;
    const int64_t bdx = (int)(B & 0xFFFF) - dx, bdy = (int)(B >> 16) - dy;
    const int64_t cdx = (int)(C & 0xFFFF) - dx, cdy = (int)(C >> 16) - dy;
    const int64_t al = adx * adx + ady * ady, bl = bdx * bdx + bdy * bdy, cl = cdx * cdx + cdy * cdy;
    return al * (bdx * cdy - cdx * bdy) + bl * (cdx * ady - adx * cdy) + cl * (adx * bdy - bdx * ady);
  }
  static int px(uint32_t v) { return (int)(v & 0xFFFF); }
  static int py(uint32_t v) { return (int)(v >> 16); }
  // > 0 iff d lies strictly inside the circle through a, b, c (a, b, c counter-clockwise)
  int64_t incircle(int a, int b, int c, int d) const {
    const Pt A = pt[a], B = pt[b], C = pt[c], D = pt[d];
    const int64_t adx = A.x - D.x, ady = A.y - D.y;
    const int64_t bdx = B.x - D.x, bdy = B.y - D.y;
    const int64_t cdx = C.x - D.x, cdy = C.y - D.y;
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

bool before_xy(const Mesh& m, int a, int px, int py) { return m.pt[a].x < px || (m.pt[a].x == px && m.pt[a].y < py); }
bool after_xy(const Mesh& m, int a, int px, int py) { return m.pt[a].x > px || (m.pt[a].x == px && m.pt[a].y > py); }

void pivot_quicksort(const Mesh& m, int* a, int n, PivotSource& rnd) {
  if (n < 2) return;
  if (n == 2) {
    if (after_xy(m, a[0], m.pt[a[1]].x, m.pt[a[1]].y)) std::swap(a[0], a[1]);
    return;
  }
  const int pv = a[rnd.next((unsigned)n)];
  const int px = m.pt[pv].x, py = m.pt[pv].y;
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

// The partition tree: halves by the node's axis, children by the other axis, leaves (<= 3 vertices) sorted by x.
// Vertices are numbered by their rank in (x, y) order.  Every node keeps its vertex set twice, once in x order (xl)
// and once in (y, x) order (yl); a split takes the first half of the list of its axis and distributes the other list
// stably with one linear pass, so no comparison sort is needed below the root.  On return xl holds the final order.
int split(int* xl, int* yl, int n, int axis, uint8_t* side, int* tmp) {
  const int divider = n >> 1;
  int* from = axis == 0 ? xl : yl;          // list that defines the halves
  int* other = axis == 0 ? yl : xl;
  for (int i = 0; i < divider; i++) side[from[i]] = 0;
  for (int i = divider; i < n; i++) side[from[i]] = 1;
  int a = 0, b = 0;
  for (int i = 0; i < n; i++) {            // branch-free stable distribution (the side bits are unpredictable)
    const int id = other[i];
    const int s = side[id];
    tmp[b] = id; other[a] = id;
    b += s; a += 1 - s;
  }
  for (int i = 0; i < b; i++) other[divider + i] = tmp[i];
  return divider;
}

void partition(int* xl, int* yl, int n, int axis, uint8_t* side, int* tmp) {
  if (n <= 3) return;                       // xl is already the x-sorted leaf
  const int divider = split(xl, yl, n, axis, side, tmp);
  partition(xl, yl, divider, 1 - axis, side, tmp);
  partition(xl + divider, yl + divider, n - divider, 1 - axis, side, tmp);
}

// the upper part of the same tree only: nodes of at most `stop` vertices are left as they are (x order in xl, (y, x)
// order in yl) and listed, in order of their first vertex
struct Node { int lo, n, axis; };
void partition_nodes(int* xl, int* yl, int lo, int n, int axis, uint8_t* side, int* tmp, int stop, std::vector<Node>& nodes) {
  if (n <= stop) { nodes.push_back(Node{lo, n, axis}); return; }
  const int divider = split(xl + lo, yl + lo, n, axis, side, tmp);
  partition_nodes(xl, yl, lo, divider, 1 - axis, side, tmp, stop, nodes);
  partition_nodes(xl, yl, lo + divider, n - divider, 1 - axis, side, tmp, stop, nodes);
}

Handles merge(Mesh& m, Handles L, Handles R, int axis) {
  int ldo = L.ldo, ldi = L.rdo, rdi = R.ldo, rdo = R.rdo;
  if (axis == 1) {
    // the two sets are separated by a horizontal line: seat the handles on the y-extremes
    while (Mesh::py(m.dxy(ldo)) < Mesh::py(m.xy(ldo))) ldo = m.rprev(ldo);
    while (Mesh::py(m.dxy(m.onext(ldi))) > Mesh::py(m.xy(ldi))) ldi = Mesh::sym(m.onext(ldi));
    while (Mesh::py(m.dxy(rdi)) < Mesh::py(m.xy(rdi))) rdi = m.rprev(rdi);
    while (Mesh::py(m.dxy(m.onext(rdo))) > Mesh::py(m.xy(rdo))) rdo = Mesh::sym(m.onext(rdo));
  }
  // lower common tangent
  bool changed;
  do {
    changed = false;
    if (Mesh::ccw_xy(m.xy(ldi), m.dxy(ldi), m.xy(rdi)) > 0) { ldi = m.lnext(ldi); changed = true; }
    if (Mesh::ccw_xy(m.dxy(rdi), m.xy(rdi), m.xy(ldi)) > 0) { rdi = m.rprev(rdi); changed = true; }
  } while (changed);
  int basel = m.connect(Mesh::sym(rdi), ldi);       // from org(rdi) = lower right to org(ldi) = lower left
  if (m.org(ldi) == m.org(ldo)) ldo = Mesh::sym(basel);
  if (m.org(rdi) == m.org(rdo)) rdo = basel;
  // knit the seam upwards
  for (;;) {
    const uint32_t lowerright = m.xy(basel), lowerleft = m.dxy(basel);
    int lcand = m.onext(Mesh::sym(basel)), rcand = m.oprev(basel);
    uint32_t upperleft = m.dxy(lcand), upperright = m.dxy(rcand);
    const bool leftfinished = Mesh::ccw_xy(upperleft, lowerleft, lowerright) <= 0;
    const bool rightfinished = Mesh::ccw_xy(upperright, lowerleft, lowerright) <= 0;
    if (leftfinished && rightfinished) break;
    if (!leftfinished) {
      for (;;) {
        const int nx = m.onext(lcand);
        if (nx == Mesh::sym(basel)) break;
        const uint32_t apex = m.dxy(nx);
        if (Mesh::ccw_xy(lowerleft, upperleft, apex) <= 0) break;                 // no real triangle beyond lcand
        if (Mesh::incircle_xy(lowerleft, lowerright, upperleft, apex) <= 0) break;
        m.remove(lcand);
        lcand = nx; upperleft = apex;
      }
    }
    if (!rightfinished) {
      for (;;) {
        const int nx = m.oprev(rcand);
        if (nx == basel) break;
        const uint32_t apex = m.dxy(nx);
        if (Mesh::ccw_xy(lowerright, apex, upperright) <= 0) break;
        if (Mesh::incircle_xy(lowerleft, lowerright, upperright, apex) <= 0) break;
        m.remove(rcand);
        rcand = nx; upperright = apex;
      }
    }
    if (leftfinished || (!rightfinished && Mesh::incircle_xy(upperleft, lowerleft, lowerright, upperright) > 0))
      basel = m.connect(rcand, Mesh::sym(basel));      // new base: upper right -> lower left
    else
      basel = m.connect(Mesh::sym(basel), Mesh::sym(lcand));   // new base: lower right -> upper left
  }
  if (axis == 1) {
    // back to the x-extremes expected by the parent (vertical cut) and by the leaves
    while (Mesh::px(m.dxy(m.oprev(ldo))) < Mesh::px(m.xy(ldo))) ldo = Mesh::sym(m.oprev(ldo));
    while (Mesh::px(m.dxy(rdo)) > Mesh::px(m.xy(rdo))) rdo = m.lnext(rdo);
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

namespace {

struct Scratch {
  std::vector<int32_t> sx, sy, orig, cnt;
  std::vector<Pt> pts, pts_in;
  std::vector<int> v, yl, tmp, a, b;
  std::vector<uint8_t> side;
  std::vector<uint64_t> ka, kb;
  Mesh m;
  // device-built nodes
  std::vector<Node> nodes;
  std::vector<uint32_t> packed;
  std::vector<int32_t> first, count, axis;
  std::vector<int32_t> mesh_first;
  std::vector<Handles> node_h;
};

thread_local visocu_ctx* g_device = nullptr;
thread_local long g_device_nodes = 0;
const int kDeviceMin = 6000;          // fewer distinct points: the list fits the device path of removeOutliers as a whole
const int kDeviceNode = 4096;         // largest node handed to the device (shared memory of one CTA)

// stable counting sort of ids by key[id] (keys in [lo, lo + range))
void counting_pass(const std::vector<int>& in, std::vector<int>& out, const int32_t* key, int lo, int range, std::vector<int32_t>& cnt) {
  cnt.assign((size_t)range + 1, 0);
  for (int id : in) cnt[key[id] - lo + 1]++;
  for (int k = 0; k < range; k++) cnt[k + 1] += cnt[k];
  out.resize(in.size());
  for (int id : in) out[cnt[key[id] - lo]++] = id;
}

int build_on_device(Scratch& S, int nu, int xlo, int ylo, Handles& out);

// stable counting sort of records by the 16-bit digit at `shift` (values in [0, range))
void radix_pass(const std::vector<uint64_t>& in, std::vector<uint64_t>& out, int shift, int range, std::vector<int32_t>& cnt) {
  cnt.assign((size_t)range + 1, 0);
  for (uint64_t k : in) cnt[((k >> shift) & 0xFFFF) + 1]++;
  for (int k = 0; k < range; k++) cnt[k + 1] += cnt[k];
  for (uint64_t k : in) out[cnt[(k >> shift) & 0xFFFF]++] = k;
}

// Builds the triangulation in S.m; S.orig maps mesh vertex numbers to input indices.  Returns the final hull handles
// ({-1,-1} if fewer than two distinct points).
Handles triangulate(Scratch& S, const int32_t* x, const int32_t* y, int n) {
  Handles none{-1, -1};
  if (n < 2) return none;
  int xlo = x[0], xhi = x[0], ylo = y[0], yhi = y[0];
  for (int i = 1; i < n; i++) {
    xlo = std::min(xlo, x[i]); xhi = std::max(xhi, x[i]);
    ylo = std::min(ylo, y[i]); yhi = std::max(yhi, y[i]);
  }
  if (xhi - xlo > 0xFFFF || yhi - ylo > 0xFFFF) {
    fprintf(stderr, "ERROR: delaunay: coordinate range %d x %d exceeds 65535\n", xhi - xlo, yhi - ylo);
    return none;
  }
  // sort by (x, y, input index): two stable counting passes (pixel coordinates span a few thousand values) over records
  // that carry their keys along, x | y | index, so that every pass reads and writes one array front to back
  S.ka.resize(n); S.kb.resize(n);
  for (int i = 0; i < n; i++) S.ka[i] = ((uint64_t)(uint32_t)(x[i] - xlo) << 48) | ((uint64_t)(uint32_t)(y[i] - ylo) << 32) | (uint32_t)i;
  radix_pass(S.ka, S.kb, 32, yhi - ylo + 1, S.cnt);
  radix_pass(S.kb, S.ka, 48, xhi - xlo + 1, S.cnt);
  S.a.resize(n);
  for (int i = 0; i < n; i++) S.a[i] = (int)(uint32_t)S.ka[i];
  bool duplicates = false;
  for (int i = 1; i < n && !duplicates; i++) duplicates = (S.ka[i] >> 32) == (S.ka[i - 1] >> 32);
  S.orig.clear();
  if (!duplicates) {
    S.orig.assign(S.a.begin(), S.a.end());
  } else {
    // Of several input points on the same pixel only one takes part (Triangle ignores duplicates).  Which one
    // survives depends on the order an unstable sort leaves them in, and removeOutliers keeps exactly the survivor,
    // so in this (rare, quad matching only) case the sort is redone with the same randomised quicksort (Hoare
    // partition, pivots from the generator seeded with 1) the reference's triangulation uses, and the first of
    // each run of equal points is kept.
    for (int i = 0; i < n; i++) S.a[i] = i;
    Mesh tmpm;
    S.pts_in.resize(n);
    for (int i = 0; i < n; i++) S.pts_in[i] = Pt{x[i], y[i]};
    tmpm.set_points(S.pts_in.data());
    PivotSource pivots;
    pivot_quicksort(tmpm, S.a.data(), n, pivots);
    S.orig.push_back(S.a[0]);
    for (int i = 1; i < n; i++)
      if (x[S.a[i]] != x[S.orig.back()] || y[S.a[i]] != y[S.orig.back()]) S.orig.push_back(S.a[i]);
  }
  const int nu = (int)S.orig.size();
  if (nu < 2) return none;
  // vertices renumbered by (x, y) rank, coordinates stored in that order
  S.sx.resize(nu); S.sy.resize(nu);
  if (!duplicates) for (int i = 0; i < nu; i++) { S.sx[i] = xlo + (int32_t)(S.ka[i] >> 48); S.sy[i] = ylo + (int32_t)((S.ka[i] >> 32) & 0xFFFF); }
  else for (int i = 0; i < nu; i++) { S.sx[i] = x[S.orig[i]]; S.sy[i] = y[S.orig[i]]; }
  S.v.resize(nu);
  for (int i = 0; i < nu; i++) S.v[i] = i;
  counting_pass(S.v, S.yl, S.sy.data(), ylo, yhi - ylo + 1, S.cnt);     // ids in (y, x) order
  S.side.resize(nu); S.tmp.resize(nu);
  if (g_device && nu > kDeviceMin && xhi - xlo <= 8191 && yhi - ylo <= 8191) {
    Handles h;
    if (build_on_device(S, nu, xlo, ylo, h)) return h;
  } else {
    // root: vertical cut of the x-sorted list, children by alternating axes
    partition(S.v.data(), S.yl.data(), nu, 0, S.side.data(), S.tmp.data());
  }
  S.m.reset(12 * (size_t)nu + 64);
  // renumber once more, in partition order: every subtree of the divide-and-conquer then works on a contiguous range
  // of vertices (and of the edges it creates), which keeps the merge loops in cache
  S.pts.resize(nu);
  // coordinates relative to the bounding box: they are packed into 16 + 16 bits inside the half-edge records
  for (int i = 0; i < nu; i++) { S.pts[i] = Pt{S.sx[S.v[i]] - xlo, S.sy[S.v[i]] - ylo}; S.tmp[i] = S.orig[S.v[i]]; }
  for (int i = 0; i < nu; i++) { S.orig[i] = S.tmp[i]; S.v[i] = i; }
  S.m.set_points(S.pts.data());
  return build(S.m, S.v.data(), nu, 0);
}

// Large triangulations: the tree is cut where its nodes fit the device kernel (depth 5 for the 87 k matches of a
// 3840x2160 frame pair: 32 nodes of 2.7 k vertices), the device triangulates the nodes (one CTA each, in parallel), and
// the few merges above them - whose seams are all that is left of the sequential work - run here, in place on the records the device delivered.
// Returns 1 = done (handles in out), 0 = the device did not deliver; S.v / S.yl are then partitioned completely, exactly
// as partition() from the root would have left them, and the caller builds on the host.
int build_on_device(Scratch& S, int nu, int xlo, int ylo, Handles& out) {
  S.nodes.clear();
  partition_nodes(S.v.data(), S.yl.data(), 0, nu, 0, S.side.data(), S.tmp.data(), kDeviceNode, S.nodes);
  const int nj = (int)S.nodes.size();
  auto finish_on_host = [&]() {
    for (const Node& nd : S.nodes) partition(S.v.data() + nd.lo, S.yl.data() + nd.lo, nd.n, nd.axis, S.side.data(), S.tmp.data());
    return 0;
  };
  S.packed.resize(nu); S.first.resize(nj); S.count.resize(nj); S.axis.resize(nj); S.mesh_first.resize(nj);
  for (int i = 0; i < nu; i++) S.packed[i] = (uint32_t)(S.sx[S.v[i]] - xlo) | ((uint32_t)(S.sy[S.v[i]] - ylo) << 16);
  for (int j = 0; j < nj; j++) { S.first[j] = S.nodes[j].lo; S.count[j] = S.nodes[j].n; S.axis[j] = S.nodes[j].axis; }
  int32_t* mesh = nullptr; const int32_t* vert = nullptr; const int32_t* res = nullptr;
  int32_t n_he = 0;
  const int extra = 65536 + nu / 2;        // half-edges the merges above the nodes may add (a few per seam vertex)
  if (visocu_delaunay_subtrees(g_device, S.packed.data(), nu, nj, S.first.data(), S.count.data(), S.axis.data(), extra, &mesh, S.mesh_first.data(),
                               &n_he, &vert, &res) != VISOCU_OK) {
    fprintf(stderr, "WARNING: delaunay: device nodes failed (%s), triangulating on the host\n", visocu_last_error(g_device));
    return finish_on_host();
  }
  // status, sizes, handles and the vertex maps are checked before use (the records themselves are taken as they are,
  // like every other list a kernel delivers)
  static_assert(sizeof(HalfEdge) == 16, "the device writes half-edge records of four words");
  HalfEdge* he = reinterpret_cast<HalfEdge*>(mesh);
  S.node_h.resize(nj);
  for (int j = 0; j < nj; j++) {
    const int32_t* r = res + 16 * j;
    const int lo = S.nodes[j].lo, n = S.nodes[j].n, E = r[2], C2 = 2 * visocu_delaunay_edge_capacity(n), b0 = S.mesh_first[j];
    if (r[1] != 0 || r[0] != n || E <= 0 || 2 * E > C2 || b0 < 0 || b0 + C2 > n_he || r[4] < b0 || r[4] >= b0 + 2 * E || r[5] < b0 || r[5] >= b0 + 2 * E)
      return finish_on_host();
    if (he[r[4]].org < 0 || he[r[5]].org < 0) return finish_on_host();
    std::fill(S.side.begin() + lo, S.side.begin() + lo + n, 0);
    for (int i = 0; i < n; i++) {
      const int32_t k = vert[lo + i];
      if (k < 0 || k >= n || S.side[lo + k]) return finish_on_host();
      S.side[lo + k] = 1;
      S.tmp[lo + i] = S.v[lo + k];            // vertex order of the complete partition tree (what partition() computes on the host)
    }
    S.node_h[j] = Handles{r[4], r[5]};
  }
  std::copy(S.tmp.begin(), S.tmp.begin() + nu, S.v.begin());
  S.pts.resize(nu);
  for (int i = 0; i < nu; i++) { S.pts[i] = Pt{S.sx[S.v[i]] - xlo, S.sy[S.v[i]] - ylo}; S.tmp[i] = S.orig[S.v[i]]; }
  for (int i = 0; i < nu; i++) { S.orig[i] = S.tmp[i]; S.v[i] = i; }
  Mesh& m = S.m;
  m.set_points(S.pts.data());
  m.adopt(he, n_he, n_he + extra);           // the merges work on the delivered records in place
  // the merges above the nodes, in the order of the recursion
  struct Top {
    Scratch& S; int next;
    Handles run(int n, int axis) {
      if (n <= kDeviceNode) return S.node_h[next++];
      const int divider = n >> 1;
      const Handles L = run(divider, 1 - axis);
      const Handles R = run(n - divider, 1 - axis);
      return merge(S.m, L, R, axis);
    }
  } top{S, 0};
  out = top.run(nu, 0);
  g_device_nodes += nj;
  return 1;
}

Scratch& scratch() {
  static thread_local Scratch S;     // reused from call to call (one ~5 k point triangulation per frame pair and worker)
  return S;
}

}  // namespace

void delaunay_use_device(visocu_ctx* ctx) { g_device = ctx; }
long delaunay_device_nodes() { return g_device_nodes; }

void delaunay_triangles(const int32_t* x, const int32_t* y, int n, std::vector<int32_t>& tri) {
  tri.clear();
  if (n < 3) return;
  Scratch& S = scratch();
  if (triangulate(S, x, y, n).ldo < 0) return;
  const Mesh& m = S.m;
  // every bounded face is a counter-clockwise triangle; report each once, from its lowest-numbered half-edge
  const int ne = m.nhe;
  tri.reserve(6 * S.orig.size());
  for (int e = 0; e < ne; e++) {
    if (m.flag[e] || m.he[e].org < 0) continue;
    const int e2 = m.lnext(e);
    if (e2 < e) continue;
    const int e3 = m.lnext(e2);
    if (e3 < e || m.lnext(e3) != e) continue;
    const int a = m.org(e), b = m.org(e2), c = m.org(e3);
    if (m.ccw(a, b, c) > 0) { tri.push_back(S.orig[a]); tri.push_back(S.orig[b]); tri.push_back(S.orig[c]); }
  }
}

void delaunay_edges(const int32_t* x, const int32_t* y, int n, std::vector<int32_t>& edges) {
  edges.clear();
  if (n < 3) return;
  Scratch& S = scratch();
  Handles h = triangulate(S, x, y, n);
  if (h.ldo < 0) return;
  Mesh& m = S.m;
  // mark the half-edges whose left face is the unbounded face: one walk around it, starting at the clockwise
  // hull edge out of the rightmost vertex.  Every other face of a Delaunay triangulation is a triangle.
  int e = h.rdo;
  do { m.flag[e] |= 2; e = m.lnext(e); } while (e != h.rdo);
  const int ne = m.nhe;
  edges.resize(3 * ((size_t)ne / 2));
  int32_t* out = edges.data();
  const HalfEdge* he = m.he;
  const uint8_t* flag = m.flag.data();
  const int32_t* orig = S.orig.data();
  for (int k = 0; k < ne; k += 2) {
    const int o0 = he[k].org, o1 = he[k + 1].org;
    if (o0 < 0) continue;                    // deleted or unused record of a device node
    const int d0 = flag[k], d1 = flag[k + 1];
    if ((d0 | d1) & 1) continue;
    const int t = (d0 & 2 ? 0 : 1) + (d1 & 2 ? 0 : 1);
    if (t == 0) continue;
    out[0] = orig[o0]; out[1] = orig[o1]; out[2] = t;
    out += 3;
  }
  edges.resize((size_t)(out - edges.data()));
}

}  // namespace visob
