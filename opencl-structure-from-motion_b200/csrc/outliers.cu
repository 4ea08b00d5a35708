// Outlier removal on the device (Matcher::removeOutliers, reference matcher.cpp:1207-1377).
//
// The reference triangulates the current-image positions of all matches (Triangle, "zQB") and keeps a match if at
// least four of its triangle edges agree in flow / disparity.  The Delaunay triangulation of pixel coordinates is full
// of co-circular point sets, so WHICH triangulation comes out is a property of the algorithm: divide and conquer with
// alternating cuts and Triangle's tie-breaking.  host/delaunay.cpp restates that algorithm with exact integer
// predicates; this file runs the same algorithm on the GPU, one CTA per match list, everything in shared memory:
//   1. bitonic sort of (x, y) keys, duplicate detection, second sort for the (y, x) order;
//   2. the partition tree (halves by rank, alternating axes) level by level with block-wide stable partitions;
//   3. the divide-and-conquer build bottom-up: all subtrees of one depth are merged concurrently, one thread per
//      merge, on a 16-bit quad-edge structure with per-subtree free lists; the top levels are single threads walking
//      the seam at shared-memory latency while the other CTAs of the batch keep the SMs busy;
//   4. support vote per triangle edge with shared-memory atomics, order-preserving compaction of the survivors.
// Duplicate positions (Triangle triangulates one point per position, the one its randomised quicksort leaves first) are
// resolved before the kernel runs by replaying that sort on the host (ro_resolve_duplicates below); the kernel gets one
// flag per match.  A list that does not fit (more than about 5700 points in 227 KB), has coordinates above 8191 (the
// 32-bit predicates), fewer than four distinct positions, or trips an internal guard is handed back unchanged with a
// non-zero status, and the host runs the identical algorithm (host/delaunay.cpp) on it.
#include "visocu_internal.cuh"
#include "outliers.cuh"
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <utility>
#include <vector>

namespace {

constexpr int RO_THREADS = 1024;        // with 64 registers each: the CTA owns every register of its SM, so no other kernel's
                                        // warps compete with the single-threaded top merges for issue slots
constexpr int NIL = 0xFFFF;
constexpr uint16_t DEAD = 0xFFFF;

__host__ __device__ inline int ro_edge_capacity(int n) { return (int)(3.04f * (float)n) + 64; }
__host__ __device__ inline size_t ro_smem_bytes(int n) {
  return (((size_t)12 * ro_edge_capacity(n) + 15) & ~(size_t)15) + (size_t)4 * n + 16;
}

struct Mesh {
  // per half-edge: next / previous around the origin (counter-clockwise), origin vertex.  Three separate arrays that never
  // alias: with __restrict__ the compiler may hoist the loads of one above the stores to another
  uint16_t* __restrict__ nx; uint16_t* __restrict__ pv; uint16_t* __restrict__ og;
  const uint32_t* __restrict__ pt;   // x | y << 16 per vertex
  int* bump; int ecap; volatile int* fail;
  int guard;                   // remaining loop iterations of this thread before it gives up

  __device__ __forceinline__ int org(int e) const { return og[e]; }
  __device__ __forceinline__ int dest(int e) const { return og[e ^ 1]; }
  __device__ __forceinline__ int onext(int e) const { return nx[e]; }
  __device__ __forceinline__ int oprev(int e) const { return pv[e]; }
  __device__ __forceinline__ int lnext(int e) const { return pv[e ^ 1]; }
  __device__ __forceinline__ int rprev(int e) const { return nx[e ^ 1]; }
  __device__ __forceinline__ int px(int v) const { return (int)(pt[v] & 0xFFFFu); }
  __device__ __forceinline__ int py(int v) const { return (int)(pt[v] >> 16); }
  __device__ __forceinline__ bool tick() { return --guard >= 0; }    // false once this thread's loop budget is used up
  // Predicates in 32-bit arithmetic with 64-bit products only where needed: exact for coordinates below 8192 (checked
  // when the keys are built; larger images go to the host).  > 0 iff a, b, c make a left turn
  __device__ __forceinline__ int ccw(int a, int b, int c) const {
    const uint32_t A = pt[a], B = pt[b], C = pt[c];
    const int cx = (int)(C & 0xFFFF), cy = (int)(C >> 16);
    const int ax = (int)(A & 0xFFFF) - cx, ay = (int)(A >> 16) - cy;
    const int bx = (int)(B & 0xFFFF) - cx, by = (int)(B >> 16) - cy;
    return ax * by - ay * bx;                                        // |.| < 2^27
  }
  // > 0 iff d lies strictly inside the circle through a, b, c (counter-clockwise)
  __device__ __forceinline__ long long incircle(int a, int b, int c, int d) const {
    const uint32_t A = pt[a], B = pt[b], C = pt[c], D = pt[d];
    const int dx = (int)(D & 0xFFFF), dy = (int)(D >> 16);
    const int adx = (int)(A & 0xFFFF) - dx, ady = (int)(A >> 16) - dy;
    const int bdx = (int)(B & 0xFFFF) - dx, bdy = (int)(B >> 16) - dy;
    const int cdx = (int)(C & 0xFFFF) - dx, cdy = (int)(C >> 16) - dy;
    const int al = adx * adx + ady * ady, bl = bdx * bdx + bdy * bdy, cl = cdx * cdx + cdy * cdy;     // < 2^27
    return (long long)al * (bdx * cdy - cdx * bdy) + (long long)bl * (cdx * ady - adx * cdy) + (long long)cl * (adx * bdy - bdx * ady);
  }
};

// free edges of a subtree: singly linked through nx[] of the even half-edge
struct FreeList { int head, tail; };

__device__ __forceinline__ int make_edge(Mesh& m, FreeList& fl, int a, int b) {
  int e;
  if (fl.head != NIL) {
    e = fl.head;
    fl.head = (e == fl.tail) ? NIL : m.nx[e];
  } else {
    const int k = atomicAdd(m.bump, 1);
    if (k >= m.ecap) { *m.fail = 1; e = 0; } else e = 2 * k;
  }
  m.nx[e] = (uint16_t)e; m.pv[e] = (uint16_t)e; m.og[e] = (uint16_t)a;
  m.nx[e + 1] = (uint16_t)(e + 1); m.pv[e + 1] = (uint16_t)(e + 1); m.og[e + 1] = (uint16_t)b;
  return e;
}
// put the isolated half-edge e right after x in the ring around x's origin
__device__ __forceinline__ void insert_after(Mesh& m, int x, int e) {
  const int n = m.nx[x];
  m.nx[e] = (uint16_t)n; m.pv[e] = (uint16_t)x;
  m.pv[n] = (uint16_t)e; m.nx[x] = (uint16_t)e;
}
__device__ __forceinline__ void unlink(Mesh& m, int e) {
  const int n = m.nx[e], p = m.pv[e];
  m.nx[p] = (uint16_t)n; m.pv[n] = (uint16_t)p;
}
// new edge from dest(a) to org(b) so that a, the new edge and b share their left face
__device__ __forceinline__ int connect(Mesh& m, FreeList& fl, int a, int b) {
  const int e = make_edge(m, fl, m.dest(a), m.org(b));
  insert_after(m, m.lnext(a), e);
  insert_after(m, b, e ^ 1);
  return e;
}
__device__ __forceinline__ void remove_edge(Mesh& m, FreeList& fl, int e) {
  unlink(m, e); unlink(m, e ^ 1);
  e &= ~1;
  m.og[e] = DEAD; m.og[e + 1] = DEAD;
  if (fl.head == NIL) { fl.head = fl.tail = e; } else { m.nx[e] = (uint16_t)fl.head; fl.head = e; }
}

struct Handles { int ldo, rdo; };   // ccw hull edge out of the leftmost vertex, cw hull edge out of the rightmost

// Merge of two triangulations separated by a vertical (axis 0) or horizontal (axis 1) line; the same steps, tests and
// tie-breaking as merge() in host/delaunay.cpp.
__device__ Handles merge(Mesh& m, FreeList& fl, Handles L, Handles R, int axis) {
  int ldo = L.ldo, ldi = L.rdo, rdi = R.ldo, rdo = R.rdo;
  if (axis == 1) {
    while (m.py(m.dest(ldo)) < m.py(m.org(ldo)) && m.tick()) ldo = m.rprev(ldo);
    while (m.py(m.dest(m.onext(ldi))) > m.py(m.org(ldi)) && m.tick()) ldi = m.onext(ldi) ^ 1;
    while (m.py(m.dest(rdi)) < m.py(m.org(rdi)) && m.tick()) rdi = m.rprev(rdi);
    while (m.py(m.dest(m.onext(rdo))) > m.py(m.org(rdo)) && m.tick()) rdo = m.onext(rdo) ^ 1;
  }
  bool changed;
  do {
    changed = false;
    if (m.ccw(m.org(ldi), m.dest(ldi), m.org(rdi)) > 0) { ldi = m.lnext(ldi); changed = true; }
    if (m.ccw(m.dest(rdi), m.org(rdi), m.org(ldi)) > 0) { rdi = m.rprev(rdi); changed = true; }
  } while (changed && m.tick());
  int basel = connect(m, fl, rdi ^ 1, ldi);
  if (m.org(ldi) == m.org(ldo)) ldo = basel ^ 1;
  if (m.org(rdi) == m.org(rdo)) rdo = basel;
  while (m.tick()) {
    const int lowerright = m.org(basel), lowerleft = m.dest(basel);
    int lcand = m.onext(basel ^ 1), rcand = m.oprev(basel);
    int upperleft = m.dest(lcand), upperright = m.dest(rcand);
    // The first test of the left and of the right candidate loop read disjoint parts of the mesh (the left loop only
    // touches edges of the left triangulation and the ring of lowerleft, the right loop the mirror image), and most
    // steps delete nothing.  Both first tests are therefore evaluated up front, before any store, so that their load
    // chains overlap; the loops below continue from the second candidate on.
    const int lnx0 = m.onext(lcand), rnx0 = m.oprev(rcand);
    const int lapex0 = m.dest(lnx0), rapex0 = m.dest(rnx0);
    const bool leftfinished = m.ccw(upperleft, lowerleft, lowerright) <= 0;
    const bool rightfinished = m.ccw(upperright, lowerleft, lowerright) <= 0;
    if (leftfinished && rightfinished) break;
    const bool ldel0 = !leftfinished && lnx0 != (basel ^ 1) && m.ccw(lowerleft, upperleft, lapex0) > 0 &&
                       m.incircle(lowerleft, lowerright, upperleft, lapex0) > 0;
    const bool rdel0 = !rightfinished && rnx0 != basel && m.ccw(lowerright, rapex0, upperright) > 0 &&
                       m.incircle(lowerleft, lowerright, upperright, rapex0) > 0;
    if (ldel0) {
      remove_edge(m, fl, lcand);
      lcand = lnx0; upperleft = lapex0;
      while (m.tick()) {
        const int nx = m.onext(lcand);
        if (nx == (basel ^ 1)) break;
        const int apex = m.dest(nx);
        if (m.ccw(lowerleft, upperleft, apex) <= 0) break;
        if (m.incircle(lowerleft, lowerright, upperleft, apex) <= 0) break;
        remove_edge(m, fl, lcand);
        lcand = nx; upperleft = apex;
      }
    }
    if (rdel0) {
      remove_edge(m, fl, rcand);
      rcand = rnx0; upperright = rapex0;
      while (m.tick()) {
        const int nx = m.oprev(rcand);
        if (nx == basel) break;
        const int apex = m.dest(nx);
        if (m.ccw(lowerright, apex, upperright) <= 0) break;
        if (m.incircle(lowerleft, lowerright, upperright, apex) <= 0) break;
        remove_edge(m, fl, rcand);
        rcand = nx; upperright = apex;
      }
    }
    if (leftfinished || (!rightfinished && m.incircle(upperleft, lowerleft, lowerright, upperright) > 0))
      basel = connect(m, fl, rcand, basel ^ 1);
    else
      basel = connect(m, fl, basel ^ 1, lcand ^ 1);
  }
  if (axis == 1) {
    while (m.px(m.dest(m.oprev(ldo))) < m.px(m.org(ldo)) && m.tick()) ldo = m.oprev(ldo) ^ 1;
    while (m.px(m.dest(rdo)) > m.px(m.org(rdo)) && m.tick()) rdo = m.lnext(rdo);
  }
  return Handles{ldo, rdo};
}

// two or three vertices, consecutive numbers starting at v, sorted by x
__device__ Handles leaf(Mesh& m, FreeList& fl, int v, int n) {
  if (n == 2) {
    const int a = make_edge(m, fl, v, v + 1);
    return Handles{a, a ^ 1};
  }
  const int a = make_edge(m, fl, v, v + 1);
  const int b = make_edge(m, fl, v + 1, v + 2);
  insert_after(m, b, a ^ 1);
  const int area = m.ccw(v, v + 1, v + 2);
  if (area == 0) return Handles{a, b ^ 1};
  const int c = connect(m, fl, b, a);
  if (area > 0) return Handles{a, b ^ 1};
  return Handles{c ^ 1, c};
}

// exclusive prefix sum over the block of one value per thread; total returned to every thread
__device__ __forceinline__ int block_scan(int v, int* s_warp, int& total) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  int incl = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xFFFFFFFFu, incl, o); if (lane >= o) incl += t; }
  __syncthreads();                                   // s_warp may still be read from the previous call
  if (lane == 31) s_warp[wid] = incl;
  __syncthreads();
  int base = 0, tot = 0;
#pragma unroll
  for (int w = 0; w < RO_THREADS / 32; w++) { const int s = s_warp[w]; if (w < wid) base += s; tot += s; }
  total = tot;
  return base + incl - v;
}

__device__ __forceinline__ bool edge_agrees(const visocu_pmatch& a, const visocu_pmatch& b, int method, float flow_tol, float disp_tol) {
  // float arithmetic, operation by operation as matcher.cpp:1267-1349 (no contraction possible: no multiplications)
  if (method == 0)
    return __fadd_rn(fabsf(__fsub_rn(__fsub_rn(a.u1c, a.u1p), __fsub_rn(b.u1c, b.u1p))),
                     fabsf(__fsub_rn(__fsub_rn(a.v1c, a.v1p), __fsub_rn(b.v1c, b.v1p)))) < flow_tol;
  if (method == 1) return fabsf(__fsub_rn(__fsub_rn(a.u1c, a.u2c), __fsub_rn(b.u1c, b.u2c))) < disp_tol;
  return fabsf(__fsub_rn(__fsub_rn(a.u1p, a.u2p), __fsub_rn(b.u1p, b.u2p))) < disp_tol &&
         __fadd_rn(fabsf(__fsub_rn(__fsub_rn(a.u1c, a.u1p), __fsub_rn(b.u1c, b.u1p))),
                   fabsf(__fsub_rn(__fsub_rn(a.v1c, a.v1p), __fsub_rn(b.v1c, b.v1p)))) < flow_tol;
}

// copy 48-byte records src[list[i]] -> dst[i], 12 lanes per record
__device__ __forceinline__ void copy_records(visocu_pmatch* dst, const visocu_pmatch* src, const int32_t* list, int n) {
  const int32_t* s = (const int32_t*)src;
  int32_t* d = (int32_t*)dst;
  for (int idx = threadIdx.x; idx < 12 * n; idx += RO_THREADS) {
    const int i = idx / 12, w = idx - 12 * i;
    d[idx] = s[12 * (size_t)list[i] + w];
  }
}

__global__ void __launch_bounds__(RO_THREADS, 1)
k_remove_outliers(const RoJob* __restrict__ jobs, int method, float flow_tol, float disp_tol, int smem_bytes) {
  extern __shared__ __align__(16) uint8_t smem[];
  __shared__ int s_warp[RO_THREADS / 32];
  __shared__ int s_bump, s_fail, s_flag, s_more;
  const RoJob J = jobs[blockIdx.x];
  const int tid = threadIdx.x;
  auto now = []() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; };
  const unsigned long long t_start = now();
  const int n_in = *J.n_in;
  int32_t* src = J.idx;                              // global scratch: list position -> record index (after the keep flags)

  // ---- 0. records that survived the sub-pixel refinement (order preserved)
  int n = 0;
  {
    const int per = (n_in + RO_THREADS - 1) / RO_THREADS;
    const int i0 = min(tid * per, n_in), i1 = min(i0 + per, n_in);
    int cnt = 0;
    for (int i = i0; i < i1; i++) cnt += (!J.keep_in || J.keep_in[i]) ? 1 : 0;
    int total;
    int off = block_scan(cnt, s_warp, total);
    n = total;
    for (int i = i0; i < i1; i++) if (!J.keep_in || J.keep_in[i]) src[off++] = i;
  }
  if (tid == 0) { s_bump = 0; s_fail = 0; s_flag = 0; }
  __syncthreads();

  // capacity of this launch's shared memory: 12 E bytes of half-edges and 4 n of points.  A triangulation of n points
  // has fewer than 3 n edges and the free lists recycle every deleted one, so E = 3.04 n + 64 is enough in practice
  // (measured peak 2.98 n); running out trips the guard and hands the list to the host.
  const int ecap = ro_edge_capacity(n);
  const size_t need = ro_smem_bytes(n);
  const int nl = n;                                  // list length; n becomes the number of mesh vertices (distinct positions)
  if (n <= 3 || n > 0x7FF0 || need > (size_t)smem_bytes) {
    // nothing to vote on (matcher.cpp:1210-1211), or too large for the device path: hand the list over unchanged
    copy_records(J.out, J.in, src, n);
    if (tid == 0) { J.result[0] = n; J.result[1] = n <= 3 ? 0 : 1; J.result[2] = 0; J.result[3] = n; }
    return;
  }

  // ---- shared-memory layout
  uint16_t* he_nx = (uint16_t*)smem;
  uint16_t* he_pv = he_nx + 2 * ecap;
  uint16_t* he_og = he_pv + 2 * ecap;
  uint32_t* pts = (uint32_t*)(smem + (((size_t)12 * ecap + 15) & ~(size_t)15));
  // hull handles and free lists of the subtrees (indexed by first vertex / 2) are touched twice per merge: global scratch
  uint16_t* h_l = J.hnd;
  uint16_t* h_r = h_l + (n / 2 + 2);
  uint16_t* f_h = h_r + (n / 2 + 2);
  uint16_t* f_t = f_h + (n / 2 + 2);
  // temporaries of steps 1 and 2 live in the half-edge area
  int npad = 1;
  while (npad < n) npad <<= 1;
  unsigned long long* keys = (unsigned long long*)smem;                    // npad keys
  uint16_t* ax = (uint16_t*)(keys + npad);                                  // by x-rank: x, y, list position
  uint16_t* ay = ax + n;
  uint16_t* ain = ay + n;
  uint16_t* la = ain + n;                                                   // three list buffers (x order, y order, spare)
  uint16_t* lb = la + n;
  uint16_t* lc = lb + n;
  uint16_t* nlo = lc + n;                                                   // partition node of every list position
  uint16_t* nsz = nlo + n;
  uint16_t* pre = nsz + n;                                                  // prefix sums of the side bits
  uint8_t* side = (uint8_t*)(pre + n + 1);
  // (8 npad + 19 n + 2 <= 12 ecap because npad < 2 n)

  // Bitonic merge sort of keys[0, m), ascending.  Every compare-exchange puts the smaller key at the lower index (the
  // first step of each merge pairs i with its mirror image in the block), so the elements beyond m can stay virtual
  // "+infinity" padding: pairs that reach past m are skipped, and the work follows m, not the next power of two.
  auto bitonic = [&](int m) {
    for (int k = 2; k <= npad; k <<= 1)
      for (int j = k >> 1; j > 0; j >>= 1) {
        for (int t = tid; t < (npad >> 1); t += RO_THREADS) {
          const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));              // element with bit j clear
          const int p = (j == (k >> 1)) ? (i ^ (k - 1)) : (i | j);
          if (p < m) {
            const unsigned long long a = keys[i], b = keys[p];
            if (a > b) { keys[i] = b; keys[p] = a; }
          }
        }
        __syncthreads();
      }
  };

  // ---- 1. order by (x, y); positions are the truncated coordinates in the current left image (matcher.cpp:1230-1233)
  // (records the duplicate resolution excluded get the padding key: they are no vertices, collect no votes and drop out)
  if (tid == 0) s_more = 0;
  for (int i = tid; i < npad; i += RO_THREADS) {
    unsigned long long key = ~0ull;
    if (i < nl && (!J.rep || J.rep[src[i]])) {
      const visocu_pmatch& r = J.in[src[i]];
      const int x = (int)r.u1c, y = (int)r.v1c;
      if (x < 0 || y < 0 || x > 8191 || y > 8191) s_fail = 1;        // range of the 32-bit predicates
      key = ((unsigned long long)(unsigned)x << 40) | ((unsigned long long)(unsigned)y << 16) | (unsigned)i;
    }
    keys[i] = key;
  }
  __syncthreads();
  bitonic(nl);
  for (int i = tid; i < npad; i += RO_THREADS)
    if (keys[i] != ~0ull && (i + 1 == npad || keys[i + 1] == ~0ull)) s_more = i + 1;     // number of vertices
  __syncthreads();
  for (int i = tid; i < s_more; i += RO_THREADS) {
    const unsigned long long key = keys[i];
    if (i > 0 && (keys[i - 1] >> 16) == (key >> 16)) s_flag = 1;           // two vertices on one pixel
    ax[i] = (uint16_t)(key >> 40); ay[i] = (uint16_t)(key >> 16);
    ain[i] = (uint16_t)key;
  }
  __syncthreads();
  if (s_fail) {
    copy_records(J.out, J.in, src, nl);
    if (tid == 0) { J.result[0] = nl; J.result[1] = 1; J.result[2] = 0; J.result[3] = nl; }
    return;
  }
  if (s_flag) {
    // Two vertices on one pixel: the caller did not resolve the duplicates (ro_resolve_duplicates) - Triangle's choice
    // among them depends on its randomised quicksort, which is replayed on the host, not here.
    copy_records(J.out, J.in, src, nl);
    if (tid == 0) { J.result[0] = nl; J.result[1] = 2; J.result[2] = 0; J.result[3] = nl; }
    return;
  }
  n = s_more;                                        // number of vertices (records that take part)
  if (n <= 3) {
    copy_records(J.out, J.in, src, nl);
    if (tid == 0) { J.result[0] = nl; J.result[1] = 3; J.result[2] = 0; J.result[3] = nl; }
    return;
  }
  for (int i = tid; i < npad; i += RO_THREADS)
    keys[i] = i < n ? ((unsigned long long)ay[i] << 40) | ((unsigned long long)ax[i] << 16) | (unsigned)i : ~0ull;
  __syncthreads();
  bitonic(n);
  const unsigned long long t_sorted = now();
  uint16_t* xl = la; uint16_t* yl = lb; uint16_t* spare = lc;
  for (int i = tid; i < n; i += RO_THREADS) { xl[i] = (uint16_t)i; yl[i] = (uint16_t)keys[i]; nlo[i] = 0; nsz[i] = (uint16_t)n; }
  __syncthreads();

  // ---- 2. partition tree: halves by rank along the node's axis, the other list distributed stably
  int depth = 0;
  {
    const int per = (n + RO_THREADS - 1) / RO_THREADS;
    const int i0 = min(tid * per, n), i1 = min(i0 + per, n);
    for (;; depth++) {
      const int axis = depth & 1;
      uint16_t* from = axis == 0 ? xl : yl;
      uint16_t* other = axis == 0 ? yl : xl;
      if (tid == 0) s_more = 0;
      for (int i = i0; i < i1; i++) {
        const int sz = nsz[i];
        if (sz > 3) side[from[i]] = (uint8_t)((i - nlo[i]) >= (sz >> 1));
      }
      __syncthreads();
      int cnt = 0;
      for (int i = i0; i < i1; i++) cnt += (nsz[i] > 3) ? side[other[i]] : 0;
      int total;
      int run = block_scan(cnt, s_warp, total);
      for (int i = i0; i < i1; i++) { pre[i] = (uint16_t)run; run += (nsz[i] > 3) ? side[other[i]] : 0; }
      if (i1 == n && i0 < n) pre[n] = (uint16_t)run;
      __syncthreads();
      bool more = false;
      for (int i = i0; i < i1; i++) {
        const int sz = nsz[i], lo = nlo[i];
        const int id = other[i];
        if (sz > 3) {
          const int div = sz >> 1;
          const int ones = pre[i] - pre[lo];
          const int s = side[id];
          spare[s ? lo + div + ones : lo + (i - lo - ones)] = (uint16_t)id;
        } else {
          spare[i] = (uint16_t)id;
        }
      }
      __syncthreads();                               // every pre[lo] has been read
      for (int i = i0; i < i1; i++) {
        const int sz = nsz[i], lo = nlo[i];
        if (sz > 3) {
          const int div = sz >> 1;
          if (i - lo < div) { nsz[i] = (uint16_t)div; } else { nlo[i] = (uint16_t)(lo + div); nsz[i] = (uint16_t)(sz - div); }
          if (nsz[i] > 3) more = true;
        }
      }
      if (more) s_more = 1;
      // the distributed list replaces `other`
      if (axis == 0) { uint16_t* t = yl; yl = spare; spare = t; } else { uint16_t* t = xl; xl = spare; spare = t; }
      __syncthreads();
      if (!s_more) break;
      __syncthreads();                               // s_more is reset at the top of the next round
    }
  }
  const int maxdepth = depth + 1;                    // nodes of depth maxdepth are all leaves
  // vertices renumbered in partition order: every subtree owns a contiguous range.  The point array lies behind the
  // half-edge area, which still holds the temporaries read here.
  for (int i = tid; i < n; i += RO_THREADS) {
    const int id = xl[i];
    pts[i] = (uint32_t)ax[id] | ((uint32_t)ay[id] << 16);
    J.vert[i] = ain[id];                             // vertex -> list position
  }
  for (int i = tid; i < n / 2 + 2; i += RO_THREADS) { f_h[i] = NIL; f_t[i] = NIL; }
  __syncthreads();

  // ---- 3. build, deepest level first; one thread per subtree of the level
  const unsigned long long t_part = now();
  Mesh m;
  m.nx = he_nx; m.pv = he_pv; m.og = he_og; m.pt = pts; m.bump = &s_bump; m.ecap = ecap; m.fail = &s_fail;
  m.guard = 64 * ecap;
  unsigned long long t_level = now();
  for (int d = maxdepth; d >= 0; d--) {
    {
      // Threads of a warp that run different merges take turns (the merges diverge completely), so the subtrees of a
      // level are dealt to the warps first: 32 or fewer subtrees run on 32 different warps, one lane each.
      for (int j = (tid & 31) * (RO_THREADS / 32) + (tid >> 5); j < (1 << d); j += RO_THREADS) {
        int lo = 0, sz = n;
        bool exists = true;
        for (int t = d - 1; t >= 0; t--) {
          if (sz <= 3) { exists = false; break; }
          const int div = sz >> 1;
          if ((j >> t) & 1) { lo += div; sz -= div; } else { sz = div; }
        }
        if (!exists) continue;
        FreeList fl{NIL, NIL};
        Handles h;
        if (sz <= 3) {
          h = leaf(m, fl, lo, sz);
        } else {
          const int div = sz >> 1;
          const int sl = lo >> 1, sr = (lo + div) >> 1;
          fl.head = f_h[sl]; fl.tail = f_t[sl];
          if (f_h[sr] != NIL) {
            if (fl.head == NIL) { fl.head = f_h[sr]; fl.tail = f_t[sr]; }
            else { m.nx[fl.tail] = f_h[sr]; fl.tail = f_t[sr]; }
          }
          h = merge(m, fl, Handles{h_l[sl], h_r[sl]}, Handles{h_l[sr], h_r[sr]}, d & 1);
        }
        if (m.guard < 0) s_fail = 1;
        h_l[lo >> 1] = (uint16_t)h.ldo; h_r[lo >> 1] = (uint16_t)h.rdo;
        f_h[lo >> 1] = (uint16_t)fl.head; f_t[lo >> 1] = (uint16_t)fl.tail;
      }
    }
    __syncthreads();
    if (tid == 0 && d < 7) { const unsigned long long t = now(); J.result[9 + d] = (int32_t)(t - t_level); t_level = t; }
    else if (tid == 0) t_level = now();
    if (s_fail) break;
  }
  if (s_fail) {
    copy_records(J.out, J.in, src, nl);
    if (tid == 0) { J.result[0] = nl; J.result[1] = 3; J.result[2] = s_bump; J.result[3] = nl; }
    return;
  }

  // ---- 4. half-edges of the unbounded face, support vote, compaction
  const unsigned long long t_built = now();
  if (tid == 0) {
    const int start = h_r[0];
    int e = start, guard = 4 * ecap;
    do { he_og[e] |= 0x8000; e = m.lnext(e); } while (e != start && --guard > 0);
    if (guard <= 0) s_fail = 1;
  }
  __syncthreads();
  if (s_fail) {
    copy_records(J.out, J.in, src, nl);
    if (tid == 0) { J.result[0] = nl; J.result[1] = 3; J.result[2] = s_bump; J.result[3] = nl; }
    return;
  }
  // The ring pointers of the mesh are no longer needed.  Their arrays now hold, per list position, the vote counter and
  // the quantities the vote compares (flow / disparity of the match, computed once with the reference's float
  // operations, matcher.cpp:1267-1349), and per vertex its list position: the vote itself touches shared memory only.
  unsigned int* support = (unsigned int*)he_nx;                       // nl words
  float* q_d = (float*)he_nx + nl;                                    // disparity (methods 1 and 2), nl words <= rest of he_nx
  float* q_u = (float*)he_pv;                                         // flow u, flow v (methods 0 and 2)
  float* q_v = q_u + nl;
  uint16_t* vpos = (uint16_t*)(q_v + nl);                             // 8 nl + 2 n bytes <= 12 ecap / ... of he_pv
  const bool vote_in_smem = (size_t)8 * nl <= (size_t)4 * ecap && (size_t)8 * nl + (size_t)2 * n <= (size_t)4 * ecap;
  for (int i = tid; i < nl; i += RO_THREADS) {
    support[i] = 0;
    if (vote_in_smem) {
      const visocu_pmatch& r = J.in[src[i]];
      q_u[i] = __fsub_rn(r.u1c, r.u1p); q_v[i] = __fsub_rn(r.v1c, r.v1p);
      q_d[i] = method == 1 ? __fsub_rn(r.u1c, r.u2c) : __fsub_rn(r.u1p, r.u2p);
    }
  }
  if (vote_in_smem)
    for (int i = tid; i < n; i += RO_THREADS) vpos[i] = (uint16_t)J.vert[i];
  __syncthreads();
  const int nedge = min(s_bump, ecap);
  for (int k = tid; k < nedge; k += RO_THREADS) {
    const int o0 = he_og[2 * k], o1 = he_og[2 * k + 1];
    if (o0 == DEAD || o1 == DEAD) continue;
    // an edge with a triangle on both sides votes twice (the reference votes per triangle, matcher.cpp:1259-1362)
    const int t = ((o0 & 0x8000) ? 0 : 1) + ((o1 & 0x8000) ? 0 : 1);
    if (t == 0) continue;
    int pa, pb;
    bool agree;
    if (vote_in_smem) {
      pa = vpos[o0 & 0x7FFF]; pb = vpos[o1 & 0x7FFF];
      const bool flow_ok = __fadd_rn(fabsf(__fsub_rn(q_u[pa], q_u[pb])), fabsf(__fsub_rn(q_v[pa], q_v[pb]))) < flow_tol;
      const bool disp_ok = fabsf(__fsub_rn(q_d[pa], q_d[pb])) < disp_tol;
      agree = method == 0 ? flow_ok : (method == 1 ? disp_ok : (disp_ok && flow_ok));
    } else {
      pa = J.vert[o0 & 0x7FFF]; pb = J.vert[o1 & 0x7FFF];
      agree = edge_agrees(J.in[src[pa]], J.in[src[pb]], method, flow_tol, disp_tol);
    }
    if (agree) {
      atomicAdd(&support[pa], (unsigned)t);
      atomicAdd(&support[pb], (unsigned)t);
    }
  }
  __syncthreads();
  {
    const int per = (nl + RO_THREADS - 1) / RO_THREADS;
    const int i0 = min(tid * per, nl), i1 = min(i0 + per, nl);
    int cnt = 0;
    for (int i = i0; i < i1; i++) cnt += support[i] >= 4u ? 1 : 0;
    int total;
    int off = block_scan(cnt, s_warp, total);
    // compacted index list in place of src (entries only move towards the front, chunk by chunk behind a barrier)
    int32_t* kept = J.vert;                          // the vertex map is no longer needed
    __syncthreads();
    for (int i = i0; i < i1; i++) if (support[i] >= 4u) kept[off++] = src[i];
    __syncthreads();
    copy_records(J.out, J.in, kept, total);
    if (tid == 0) {
      J.result[0] = total; J.result[1] = 0; J.result[2] = s_bump; J.result[3] = nl;
      // phase times in nanoseconds (sort, partition, build, vote + compaction): read by profiles/profile_outliers.py
      const unsigned long long t_end = now();
      J.result[4] = (int32_t)(t_sorted - t_start); J.result[5] = (int32_t)(t_part - t_sorted);
      J.result[6] = (int32_t)(t_built - t_part); J.result[7] = (int32_t)(t_end - t_built);
    }
  }
}

// positions of all records of a list for the host-side duplicate resolution
__global__ void k_outlier_keys(const RoJob* __restrict__ jobs, uint32_t* keys, int stride) {
  const RoJob J = jobs[blockIdx.y];
  const int n = *J.n_in;
  uint32_t* out = keys + (size_t)blockIdx.y * stride;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    uint32_t k = 0xFFFFFFFFu;
    if (!J.keep_in || J.keep_in[i]) {
      const int x = (int)J.in[i].u1c, y = (int)J.in[i].v1c;
      if (x >= 0 && y >= 0 && x <= 0xFFFF && y <= 0xFFFF) k = ((uint32_t)x << 16) | (uint32_t)y;
    }
    out[i] = k;
  }
}

}  // namespace

int visocu_launch_remove_outliers(visocu_ctx* ctx, const RoJob* jobs_dev, int n_jobs, int method, int max_records, cudaStream_t stream) {
  // shared memory for the largest list of the launch, at most the 227 KB a CTA can have
  size_t smem = ro_smem_bytes(max_records);
  const size_t smem_max = 227 * 1024 - 256;          // the kernel also has a few static shared variables
  if (smem > smem_max) smem = smem_max;
  if (smem < 16 * 1024) smem = 16 * 1024;
  {
    static std::mutex mtx;
    static bool done[64] = {false};
    std::lock_guard<std::mutex> lock(mtx);
    if (!done[ctx->device & 63]) {
      CU_TRY(ctx, cudaFuncSetAttribute(k_remove_outliers, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max));
      done[ctx->device & 63] = true;
    }
  }
  k_remove_outliers<<<n_jobs, RO_THREADS, smem, stream>>>(jobs_dev, method, (float)ctx->param.outlier_flow_tolerance,
                                                              (float)ctx->param.outlier_disp_tolerance, (int)smem);
  CU_LAUNCH_CHECK(ctx);
  return VISOCU_OK;
}

extern "C" int visocu_remove_outliers(visocu_ctx* ctx, int32_t n_jobs, int32_t method, visocu_pmatch* const* inout, const int32_t* n,
                                      int32_t* n_out, int32_t* status) {
  if (!ctx) return VISOCU_EINVAL;
  if (!ctx->configured) return visocu_set_error(ctx, VISOCU_ESTATE, "context not configured");
  if (n_jobs <= 0 || !inout || !n || !n_out || !status) return visocu_set_error(ctx, VISOCU_EINVAL, "bad outlier removal arguments");
  if (method < 0 || method > 2) return visocu_set_error(ctx, VISOCU_EINVAL, "method %d not supported", method);
  CU_TRY(ctx, cudaSetDevice(ctx->device));
  std::vector<RoJob> hj(n_jobs);
  std::vector<size_t> o_in(n_jobs), o_out(n_jobs), o_idx(n_jobs), o_vert(n_jobs), o_res(n_jobs), o_hnd(n_jobs), o_rep(n_jobs);
  std::vector<std::vector<uint8_t> > reps(n_jobs);
  size_t off = align_up(sizeof(RoJob) * n_jobs, 256);
  int maxn = 0;
  for (int j = 0; j < n_jobs; j++) {
    if (n[j] < 0) return visocu_set_error(ctx, VISOCU_EINVAL, "job %d has a negative count", j);
    const size_t cnt = (size_t)n[j] + 1;
    if (n[j] > maxn) maxn = n[j];
    o_in[j] = off; off += align_up(cnt * 48, 256);
    o_out[j] = off; off += align_up(cnt * 48, 256);
    o_idx[j] = off; off += align_up(cnt * 4, 256);
    o_vert[j] = off; off += align_up(cnt * 4, 256);
    o_hnd[j] = off; off += align_up((cnt / 2 + 2) * 8, 256);
    o_rep[j] = off; off += align_up(cnt, 256);
    o_res[j] = off; off += 256;
  }
  int rc = visocu_ensure_scratch(ctx, off);
  if (rc) return rc;
  if ((rc = visocu_ensure_pinned(ctx, align_up(sizeof(RoJob) * n_jobs, 256) + (size_t)n_jobs * 64))) return rc;
  uint8_t* sb = (uint8_t*)ctx->scratch;
  uint8_t* pin = (uint8_t*)ctx->pinned;
  for (int j = 0; j < n_jobs; j++) {
    RoJob& J = hj[j];
    J.in = (const visocu_pmatch*)(sb + o_in[j]); J.keep_in = nullptr; J.out = (visocu_pmatch*)(sb + o_out[j]);
    J.result = (int32_t*)(sb + o_res[j]); J.n_in = J.result + 8;   // [9..15]: times of the top merge levels
    J.idx = (int32_t*)(sb + o_idx[j]); J.vert = (int32_t*)(sb + o_vert[j]); J.hnd = (uint16_t*)(sb + o_hnd[j]);
    if (n[j] > 0) CU_COPY(ctx, sb + o_in[j], inout[j], (size_t)n[j] * 48, cudaMemcpyHostToDevice);
    J.rep = nullptr;
    if (n[j] > 3) {
      std::vector<uint32_t> keys((size_t)n[j]);
      for (int i = 0; i < n[j]; i++) {
        const int x = (int)inout[j][i].u1c, y = (int)inout[j][i].v1c;
        keys[i] = (x >= 0 && y >= 0 && x <= 0xFFFF && y <= 0xFFFF) ? ((uint32_t)x << 16) | (uint32_t)y : 0xFFFFFFFFu;
      }
      reps[j].resize((size_t)n[j]);
      if (ro_resolve_duplicates(keys.data(), n[j], reps[j].data())) {
        CU_COPY(ctx, sb + o_rep[j], reps[j].data(), (size_t)n[j], cudaMemcpyHostToDevice);
        J.rep = sb + o_rep[j];
      }
    }
    CU_COPY(ctx, (void*)J.n_in, &n[j], 4, cudaMemcpyHostToDevice);
  }
  memcpy(pin, hj.data(), sizeof(RoJob) * n_jobs);
  CU_COPY(ctx, sb, pin, sizeof(RoJob) * n_jobs, cudaMemcpyHostToDevice);
  if ((rc = visocu_launch_remove_outliers(ctx, (const RoJob*)sb, n_jobs, method, maxn, ctx->stream))) return rc;
  int32_t* pin_res = (int32_t*)(pin + align_up(sizeof(RoJob) * n_jobs, 256));
  for (int j = 0; j < n_jobs; j++) CU_COPY(ctx, pin_res + 16 * j, hj[j].result, 64, cudaMemcpyDeviceToHost);
  CU_TRY(ctx, visocu_stream_wait(ctx));
  for (int j = 0; j < n_jobs; j++) {
    n_out[j] = pin_res[16 * j]; status[j] = pin_res[16 * j + 1];
    if (getenv("VISOCU_RO_STATS") && j == 0)
      fprintf(stderr, "[outliers] n=%d kept=%d status=%d edges=%d (%.2f n) ns: sort %d partition %d build %d vote %d\n", n[j], pin_res[0], pin_res[1],
              pin_res[2], n[j] ? (double)pin_res[2] / n[j] : 0.0, pin_res[4], pin_res[5], pin_res[6], pin_res[7]),
      fprintf(stderr, "[outliers] merge levels 0..6 (ns): %d %d %d %d %d %d %d\n", pin_res[9], pin_res[10], pin_res[11], pin_res[12], pin_res[13], pin_res[14], pin_res[15]);
    if (status[j] == 0 && n_out[j] > 0) CU_COPY(ctx, inout[j], hj[j].out, (size_t)n_out[j] * 48, cudaMemcpyDeviceToHost);
  }
  CU_TRY(ctx, visocu_stream_wait(ctx));
  return VISOCU_OK;
}

int visocu_launch_outlier_keys(visocu_ctx* ctx, const RoJob* jobs_dev, int n_jobs, uint32_t* keys_dev, int stride) {
  int gx = (stride + 255) / 256; if (gx > 64) gx = 64;
  k_outlier_keys<<<dim3(gx, n_jobs), 256, 0, ctx->stream>>>(jobs_dev, keys_dev, stride);
  CU_LAUNCH_CHECK(ctx);
  return VISOCU_OK;
}

// Triangle's vertexsort (a randomised quicksort: Hoare partition, pivot index from x' = (1366 x + 150889) mod 714025
// starting at 1, scaled to the subarray length) replayed on the positions in list order; the first of every run of equal
// positions is the point Triangle triangulates, the others it ignores (they collect no votes and are removed).
bool ro_resolve_duplicates(const uint32_t* keys, int n_records, uint8_t* rep) {
  static thread_local std::vector<uint32_t> kk, seen_words;
  static thread_local std::vector<int32_t> rec;
  kk.clear(); rec.clear();
  for (int i = 0; i < n_records; i++)
    if (keys[i] != 0xFFFFFFFFu) { kk.push_back(keys[i]); rec.push_back(i); }
  const int n = (int)kk.size();
  // any position taken twice?  (open-addressing hash set, a few n entries: stays in the cache)
  size_t cap = 64;
  while (cap < (size_t)4 * n) cap <<= 1;
  seen_words.assign(cap, 0xFFFFFFFFu);
  bool dup = false;
  for (int i = 0; i < n && !dup; i++) {
    const uint32_t k = kk[i];
    size_t h = ((size_t)k * 2654435761u) & (cap - 1);
    while (seen_words[h] != 0xFFFFFFFFu) {
      if (seen_words[h] == k) { dup = true; break; }
      h = (h + 1) & (cap - 1);
    }
    seen_words[h] = k;
  }
  if (!dup) return false;
  // elements: position in the high word (what the sort compares), index into rec[] in the low word
  static thread_local std::vector<unsigned long long> el;
  el.resize((size_t)n + 1);
  for (int i = 0; i < n; i++) el[i] = ((unsigned long long)kk[i] << 32) | (unsigned)i;
  el[n] = ~0ull;
  unsigned seed = 1;
  static thread_local std::vector<std::pair<int, int> > stack;
  stack.clear();
  stack.push_back(std::make_pair(0, n));
  while (!stack.empty()) {
    const int lo = stack.back().first, len = stack.back().second;
    stack.pop_back();
    if (len < 2) continue;
    unsigned long long* v = el.data() + lo;
    if (len == 2) {
      if ((v[0] >> 32) > (v[1] >> 32)) std::swap(v[0], v[1]);
      continue;
    }
    seed = (seed * 1366u + 150889u) % 714025u;
    const unsigned long long pkey = v[seed / (714025u / (unsigned)len + 1u)] >> 32;
    int left = -1, right = len;
    while (left < right) {
      do { left++; } while (left <= right && (v[left] >> 32) < pkey);
      do { right--; } while (left <= right && (v[right] >> 32) > pkey);
      if (left < right) std::swap(v[left], v[right]);
    }
    if (right < len - 2) stack.push_back(std::make_pair(lo + right + 1, len - right - 1));    // taken second
    if (left > 1) stack.push_back(std::make_pair(lo, left));                                   // taken first
  }
  for (int i = 0; i < n_records; i++) rep[i] = 0;
  for (int i = 0; i < n; i++)
    if (i == 0 || (el[i] >> 32) != (el[i - 1] >> 32)) rep[rec[(size_t)(el[i] & 0xFFFFFFFFu)]] = 1;
  return true;
}
