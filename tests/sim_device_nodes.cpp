// Test infrastructure (CPU): host/delaunay.cpp compiled together with a stand-in for the device entry point
// visocu_delaunay_subtrees, so that the host half of the large-list path (tree cut into nodes, import of the node meshes,
// merges above them in place on the delivered records) runs without a GPU.  The stand-in triangulates every node with the
// host algorithm and hands the result back in the device's format.  Built and driven by tests/test_host_cpu.py.
#include "../opencl-structure-from-motion_b200/host/delaunay.cpp"

#include <array>
#include <cstring>
#include <random>
#include <set>

static std::vector<int32_t> g_mesh, g_vert, g_res;
static int g_calls = 0, g_fail_mode = 0;

extern "C" const char* visocu_last_error(const visocu_ctx*) { return "stand-in"; }
extern "C" int32_t visocu_delaunay_edge_capacity(int32_t n) { return 6 * n + 64; }     // the host algorithm does not recycle deleted edges
extern "C" int visocu_delaunay_subtrees(visocu_ctx*, const uint32_t* pts, int32_t n_pts, int32_t n_jobs, const int32_t* first, const int32_t* count,
                                        const int32_t* axis, int32_t extra_halfedges, int32_t** mesh, int32_t* mesh_first, int32_t* n_halfedges,
                                        const int32_t** vert, const int32_t** result) {
  using namespace visob;
  g_calls++;
  if (g_fail_mode == 1) return -1;
  size_t n_he = 0;
  for (int j = 0; j < n_jobs; j++) { mesh_first[j] = (int32_t)n_he; n_he += 2 * (size_t)visocu_delaunay_edge_capacity(count[j]); }
  g_mesh.assign(4 * (n_he + (size_t)extra_halfedges), 0); g_vert.assign(n_pts, 0); g_res.assign(16 * (size_t)n_jobs, 0);
  for (int j = 0; j < n_jobs; j++) {
    const int n = count[j];
    const uint32_t* p = pts + first[j];
    std::vector<int> xl(n), yl(n), tmp(n);
    std::vector<uint8_t> side(n);
    for (int i = 0; i < n; i++) xl[i] = yl[i] = i;
    std::sort(yl.begin(), yl.end(), [&](int a, int b) { return (p[a] >> 16) != (p[b] >> 16) ? (p[a] >> 16) < (p[b] >> 16) : (p[a] & 0xFFFF) < (p[b] & 0xFFFF); });
    partition(xl.data(), yl.data(), n, axis[j], side.data(), tmp.data());
    std::vector<Pt> P(n);
    for (int i = 0; i < n; i++) { P[i] = Pt{(int)(p[xl[i]] & 0xFFFF), (int)(p[xl[i]] >> 16)}; g_vert[first[j] + i] = xl[i]; }
    Mesh m;
    m.reset(12 * (size_t)n + 64);
    m.set_points(P.data());
    std::vector<int> v(n);
    for (int i = 0; i < n; i++) v[i] = i;
    const Handles h = build(m, v.data(), n, axis[j]);
    const int C2 = 2 * visocu_delaunay_edge_capacity(n), b0 = mesh_first[j];
    if (m.nhe > C2) { g_res[16 * j + 1] = 1; continue; }
    for (int e = 0; e < C2; e++) {
      int32_t* r = g_mesh.data() + 4 * ((size_t)b0 + e);
      if (e < m.nhe && !(m.flag[e] & 1)) { r[0] = b0 + m.he[e].onext; r[1] = b0 + m.he[e].oprev; r[2] = first[j] + m.he[e].org; r[3] = (int32_t)m.he[e].xy; }
      else { r[0] = r[1] = b0 + e; r[2] = -1; r[3] = 0; }
    }
    g_res[16 * j + 0] = n; g_res[16 * j + 1] = g_fail_mode == 2 && j == n_jobs / 2 ? 3 : 0; g_res[16 * j + 2] = m.nhe / 2;
    g_res[16 * j + 4] = b0 + h.ldo; g_res[16 * j + 5] = b0 + h.rdo;
    if (g_fail_mode == 3 && j == 1) g_mesh[4 * ((size_t)b0 + 5)] = 0x7FFFFFF0;          // a corrupt ring pointer
  }
  *mesh = g_mesh.data(); *n_halfedges = (int32_t)n_he; *vert = g_vert.data(); *result = g_res.data();
  return 0;
}

static std::vector<std::array<int32_t, 3> > normalised(const std::vector<int32_t>& e) {
  std::vector<std::array<int32_t, 3> > out;
  for (size_t k = 0; k + 2 < e.size(); k += 3) out.push_back({std::min(e[k], e[k + 1]), std::max(e[k], e[k + 1]), e[k + 2]});
  std::sort(out.begin(), out.end());
  return out;
}

// n distinct points on a grid of pitch `grid` inside w x h; fail_mode: 0 = healthy stand-in, 1 = call fails, 2 = one node
// declined, 3 = corrupt mesh.  Returns 0 if the edge list with the stand-in equals the host-only one, and *calls tells
// whether the device path was taken at all.
extern "C" int sim_check(int n, int w, int h, unsigned seed, int grid, int fail_mode, int* calls, int* n_edges) {
  std::mt19937 g(seed);
  std::vector<int32_t> x, y;
  std::set<long long> seen;
  const int gw = w / grid, gh = h / grid;
  if ((long long)gw * gh < n) return -2;
  while ((int)x.size() < n) {
    const int a = (int)(g() % gw) * grid + 3, b = (int)(g() % gh) * grid + 5;
    if (seen.insert(a * 100000LL + b).second) { x.push_back(a); y.push_back(b); }
  }
  std::vector<int32_t> e_host, e_dev;
  visob::delaunay_use_device(nullptr);
  visob::delaunay_edges(x.data(), y.data(), n, e_host);
  g_calls = 0; g_fail_mode = fail_mode;
  visob::delaunay_use_device(reinterpret_cast<visocu_ctx*>(&g_calls));
  visob::delaunay_edges(x.data(), y.data(), n, e_dev);
  visob::delaunay_use_device(nullptr);
  *calls = g_calls; *n_edges = (int)e_host.size() / 3;
  return normalised(e_host) == normalised(e_dev) ? 0 : 1;
}
