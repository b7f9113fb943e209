// kb_engine.cu -- host side of the C ABI declared in include/klampt_b200.h: world description, BVH build and
// flattening (kb_finalize), per-batch orchestration of the kernels in kb_kernels.cu.
//
// There is no CPU fallback: every query entry point runs the CUDA kernels or returns KB_ERR_CUDA.
#include "../../include/klampt_b200.h"
#include "kb_kernels.h"
#include "kb_lbvh.h"
#include "kb_types.h"
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <functional>
#include <map>
#include <numeric>
#include <string>
#include <thread>
#include <vector>

namespace {

thread_local std::string g_err;
int fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
  g_err = buf;
  return code;
}
#define CK(call) do { cudaError_t _e = (call); if (_e != cudaSuccess) return fail(KB_ERR_CUDA, "%s: %s (%s:%d)", #call, cudaGetErrorString(_e), __FILE__, __LINE__); } while (0)

struct Xf { double R[9]; double t[3]; };
inline void xf_from12(const double* a, Xf& T) { memcpy(T.R, a, 72); memcpy(T.t, a + 9, 24); }
// same operation order as a scalar fp64 R*p+t so baked world-frame vertices match a query-time transform bit for bit
inline void xf_apply(const Xf& T, const double* p, double* o) {
  double x = p[0], y = p[1], z = p[2];
  o[0] = T.R[0] * x + T.R[1] * y + T.R[2] * z + T.t[0];
  o[1] = T.R[3] * x + T.R[4] * y + T.R[5] * z + T.t[1];
  o[2] = T.R[6] * x + T.R[7] * y + T.R[8] * z + T.t[2];
}

enum { G_EMPTY = 0, G_MESH = 1, G_CLOUD = 2, G_PRIM = 3, G_BOX = 4 };   // G_BOX: only as the element kind of solid-box hierarchies

struct Geom {
  int kind = G_EMPTY;
  double margin = 0;
  std::vector<double> tri;     // 9 per triangle (expanded), local frame
  std::vector<double> sph;     // 4 per sphere (x,y,z,r), local frame
  bool solid = false;          // box primitive: `tri` holds its 12 surface triangles and the interior counts as well
  double box[15];              // solid box: centre(3), axes = columns of a row-major 3x3 (9), half dimensions(3)
  int nelem() const { return kind == G_MESH ? (int)(tri.size() / 9) : (int)(sph.size() / 4); }
  int dyn_cap = 0;             // > 0: a point cloud whose points are replaced between batches (kb_update_pointcloud); starts empty
  double dyn_radius = 0;
  bool empty() const { return kind == G_EMPTY || (nelem() == 0 && dyn_cap == 0); }
};

struct Driver { std::vector<int32_t> links; std::vector<double> scale, offset; double dmin, dmax; };

// ------------------------------------------------------------------------------------------------- BVH build
struct BNode { double lo[3], hi[3]; int left; int first, count; };   // left >= 0: children left,left+1 ; left < 0: leaf

struct Bvh {
  std::vector<BNode> nodes;
  std::vector<int> perm;       // leaf order -> input element index
  int depth = 0;
};

// Binned-SAH top-down build (16 bins, 3 axes), siblings adjacent, median split once a branch gets deeper than 48.
void build_bvh(const std::vector<double>& elo, const std::vector<double>& ehi, int n, int leaf_size, Bvh& out) {
  out.nodes.clear(); out.perm.resize(n); out.depth = 0;
  if (n <= 0) return;
  std::iota(out.perm.begin(), out.perm.end(), 0);
  std::vector<double> cen(3 * (size_t)n);
  for (int i = 0; i < n; i++) for (int k = 0; k < 3; k++) cen[3 * (size_t)i + k] = 0.5 * (elo[3 * (size_t)i + k] + ehi[3 * (size_t)i + k]);
  out.nodes.reserve(2 * (size_t)n);
  struct Task { int node, first, count, depth; };
  std::vector<Task> todo;
  out.nodes.push_back(BNode());
  { BNode pad; memset(&pad, 0, sizeof pad); pad.left = -1; pad.first = 0; pad.count = 0; out.nodes.push_back(pad); }   // keeps sibling pairs even-aligned
  todo.push_back({0, 0, n, 0});
  const int NB = 16;
  while (!todo.empty()) {
    Task t = todo.back(); todo.pop_back();
    if (t.depth > out.depth) out.depth = t.depth;
    double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300}, clo[3] = {1e300, 1e300, 1e300}, chi[3] = {-1e300, -1e300, -1e300};
    for (int i = t.first; i < t.first + t.count; i++) {
      int e = out.perm[i];
      for (int k = 0; k < 3; k++) {
        lo[k] = std::min(lo[k], elo[3 * (size_t)e + k]); hi[k] = std::max(hi[k], ehi[3 * (size_t)e + k]);
        clo[k] = std::min(clo[k], cen[3 * (size_t)e + k]); chi[k] = std::max(chi[k], cen[3 * (size_t)e + k]);
      }
    }
    {
      BNode& nd = out.nodes[t.node];
      memcpy(nd.lo, lo, 24); memcpy(nd.hi, hi, 24); nd.first = t.first; nd.count = t.count; nd.left = -1;
    }
    if (t.count <= leaf_size) continue;
    int mid = -1;
    if (t.depth < 48) {
      double bestc = 1e300; int bax = -1, bsplit = -1;
      for (int ax = 0; ax < 3; ax++) {
        double ext = chi[ax] - clo[ax];
        if (!(ext > 0)) continue;
        int cnt[NB]; double blo[NB][3], bhi[NB][3];
        for (int b = 0; b < NB; b++) { cnt[b] = 0; for (int k = 0; k < 3; k++) { blo[b][k] = 1e300; bhi[b][k] = -1e300; } }
        double sc = NB / ext;
        for (int i = t.first; i < t.first + t.count; i++) {
          int e = out.perm[i];
          int b = (int)((cen[3 * (size_t)e + ax] - clo[ax]) * sc); if (b >= NB) b = NB - 1; if (b < 0) b = 0;
          cnt[b]++;
          for (int k = 0; k < 3; k++) { blo[b][k] = std::min(blo[b][k], elo[3 * (size_t)e + k]); bhi[b][k] = std::max(bhi[b][k], ehi[3 * (size_t)e + k]); }
        }
        double ra[NB]; int rc[NB];
        double l3[3] = {1e300, 1e300, 1e300}, h3[3] = {-1e300, -1e300, -1e300}; int c = 0;
        for (int b = NB - 1; b > 0; b--) {
          c += cnt[b];
          for (int k = 0; k < 3; k++) { l3[k] = std::min(l3[k], blo[b][k]); h3[k] = std::max(h3[k], bhi[b][k]); }
          double dx = h3[0] - l3[0], dy = h3[1] - l3[1], dz = h3[2] - l3[2];
          ra[b] = c ? 2 * (dx * dy + dy * dz + dz * dx) : 0; rc[b] = c;
        }
        for (int k = 0; k < 3; k++) { l3[k] = 1e300; h3[k] = -1e300; } c = 0;
        for (int b = 0; b < NB - 1; b++) {
          c += cnt[b];
          for (int k = 0; k < 3; k++) { l3[k] = std::min(l3[k], blo[b][k]); h3[k] = std::max(h3[k], bhi[b][k]); }
          if (c == 0 || rc[b + 1] == 0) continue;
          double dx = h3[0] - l3[0], dy = h3[1] - l3[1], dz = h3[2] - l3[2];
          double cost = 2 * (dx * dy + dy * dz + dz * dx) * c + ra[b + 1] * rc[b + 1];
          if (cost < bestc) { bestc = cost; bax = ax; bsplit = b; }
        }
      }
      if (bax >= 0) {
        double ext = chi[bax] - clo[bax], sc = NB / ext;
        auto it = std::partition(out.perm.begin() + t.first, out.perm.begin() + t.first + t.count, [&](int e) {
          int b = (int)((cen[3 * (size_t)e + bax] - clo[bax]) * sc); if (b >= NB) b = NB - 1; if (b < 0) b = 0; return b <= bsplit; });
        mid = (int)(it - out.perm.begin());
        if (mid == t.first || mid == t.first + t.count) mid = -1;
      }
    }
    if (mid < 0) {   // median split on the longest axis (deterministic tie break on the element index)
      int ax = 0; double ext = hi[0] - lo[0];
      for (int k = 1; k < 3; k++) if (hi[k] - lo[k] > ext) { ext = hi[k] - lo[k]; ax = k; }
      mid = t.first + t.count / 2;
      std::nth_element(out.perm.begin() + t.first, out.perm.begin() + mid, out.perm.begin() + t.first + t.count, [&](int a, int b) {
        double ca = cen[3 * (size_t)a + ax], cb = cen[3 * (size_t)b + ax]; return ca < cb || (ca == cb && a < b); });
    }
    int l = (int)out.nodes.size();
    out.nodes.push_back(BNode()); out.nodes.push_back(BNode());
    out.nodes[t.node].left = l;
    todo.push_back({l + 1, mid, t.first + t.count - mid, t.depth + 1});
    todo.push_back({l, t.first, mid - t.first, t.depth + 1});
  }
}


inline float f_up(double x) { float f = (float)x; if ((double)f < x) f = nextafterf(f, INFINITY); return f; }
inline float i2f(int32_t i) { float f; memcpy(&f, &i, 4); return f; }

// device-side geometry: a flattened BVH + elements in leaf order
struct DevGeom {
  int node_base = 0, nnodes = 0, elem_base = 0, nelem = 0, kind = KB_ELEM_TRI, depth = 0; double margin = 0, rmax = 0; bool empty = true; double lo[3], hi[3];
  int wide_base = -1, wide_slots = 0, wdepth = 0;   // 4-wide form of the same hierarchy (slot index into the wide array), -1 = none
  int ncover = 0; double cover[KB_COVER_MAX][4];   // covering spheres (local frame) for the clearance-grid broad phase
};

#define KB_GPU_CLOUD_MIN 4096      // clouds up to this size keep the host SAH build even with cloud_builder = 1
#define KB_GPU_MESH_MIN 16384      // meshes up to this size keep the host SAH build even with mesh_builder = 1 (link meshes: quality matters most)

struct ItemSet {
  std::vector<KbItem> items; KbItem* d_items = nullptr; int nxf = 0; int maxdepth = 0; bool all_wide = true; int wdepth = 0;
  std::vector<KbProbe> probes; std::vector<uint32_t> always_on; KbProbe* d_probes = nullptr; uint32_t* d_always_on = nullptr;
};

// ------------------------------------------------------------------------------------------------- clearance grid
// Conservative distance field of one static environment group on a uniform grid.  Voxels touched by an element are
// "occupied"; an exact Euclidean distance transform in index space gives D(v) = distance (in voxels) from v to the nearest
// occupied voxel.  For a point p in voxel a and a surface point s in occupied voxel b, |p - s| >= h * |max(|a - b| - 1, 0)|
// >= h * (|a - b| - sqrt 3) >= h * (D(a) - sqrt 3): the stored value floor(4 * (D - sqrt 3 - 0.02)) is a lower bound on
// the clearance of every point of the voxel, in quarter voxels.
struct HostGrid { double o[3]; double h = 0; int dims[3] = {0, 0, 0}; std::vector<uint8_t> q; };

void parallel_for(int n, const std::function<void(int, int)>& fn) {
  int nt = (int)std::min<unsigned>(std::max(1u, std::thread::hardware_concurrency()), 16u);
  if (nt > n) nt = n < 1 ? 1 : n;
  if (nt <= 1) { fn(0, n); return; }
  std::vector<std::thread> th;
  for (int t = 0; t < nt; t++) { int a = (int)((int64_t)n * t / nt), b = (int)((int64_t)n * (t + 1) / nt); th.emplace_back([=, &fn] { fn(a, b); }); }
  for (auto& x : th) x.join();
}

// 1-D squared distance transform of f (lower envelope of parabolas); v / z are scratch of size n and n + 1
void edt_1d(const int32_t* f, int n, int32_t* d, int* v, double* z) {
  const int32_t INF = 1000000000;
  int k = 0; v[0] = 0; z[0] = -1e30; z[1] = 1e30;
  for (int q = 1; q < n; q++) {
    double s;
    for (;;) {                                   // z[0] = -1e30 is never reached: all values are finite
      const int r = v[k];
      s = (((double)f[q] + (double)q * q) - ((double)f[r] + (double)r * r)) / (2.0 * q - 2.0 * r);
      if (s <= z[k]) k--; else break;
    }
    k++; v[k] = q; z[k] = s; z[k + 1] = 1e30;
  }
  k = 0;
  for (int q = 0; q < n; q++) {
    while (z[k + 1] < q) k++;
    const int r = v[k];
    const double val = (double)(q - r) * (q - r) + (double)f[r];
    d[q] = val >= INF ? INF : (int32_t)val;
  }
}

void build_clear_grid(int kind, const std::vector<double>& elems, const double glo[3], const double ghi[3], double pad, int res, HostGrid& G) {
  double ext = 0;
  for (int k = 0; k < 3; k++) ext = std::max(ext, ghi[k] - glo[k] + 2 * pad);
  if (!(ext > 0)) ext = 1;
  G.h = ext / res;
  for (int k = 0; k < 3; k++) {
    G.o[k] = glo[k] - pad;
    G.dims[k] = std::max(1, std::min(res, (int)std::ceil((ghi[k] - glo[k] + 2 * pad) / G.h)));
  }
  const int nx = G.dims[0], ny = G.dims[1], nz = G.dims[2];
  const size_t nvox = (size_t)nx * ny * nz;
  const int32_t INF = 1000000000;
  std::vector<int32_t> D(nvox, INF);
  const double ih = 1.0 / G.h;
  auto mark_box = [&](const double* lo, const double* hi) {
    int a[3], b[3];
    for (int k = 0; k < 3; k++) {
      a[k] = (int)std::floor((lo[k] - G.o[k]) * ih - 1e-6); b[k] = (int)std::floor((hi[k] - G.o[k]) * ih + 1e-6);
      a[k] = std::max(0, std::min(G.dims[k] - 1, a[k])); b[k] = std::max(0, std::min(G.dims[k] - 1, b[k]));
    }
    for (int z = a[2]; z <= b[2]; z++) for (int y = a[1]; y <= b[1]; y++) for (int x = a[0]; x <= b[0]; x++) D[((size_t)z * ny + y) * nx + x] = 0;
  };
  if (kind == G_MESH) {
    // triangles larger than a few voxels are split (longest edge) until their boxes are small, so a big tilted triangle
    // does not occupy its whole bounding box
    struct Tri { double p[9]; int depth; };
    std::vector<Tri> st;
    const size_t nt = elems.size() / 9;
    for (size_t t = 0; t < nt; t++) {
      Tri r; memcpy(r.p, &elems[9 * t], 72); r.depth = 0; st.push_back(r);
      while (!st.empty()) {
        Tri c = st.back(); st.pop_back();
        double lo[3], hi[3]; double cells = 1;
        for (int k = 0; k < 3; k++) {
          lo[k] = std::min(c.p[k], std::min(c.p[3 + k], c.p[6 + k])); hi[k] = std::max(c.p[k], std::max(c.p[3 + k], c.p[6 + k]));
          cells *= std::floor((hi[k] - lo[k]) * ih) + 2;
        }
        if (cells <= 27 || c.depth >= 24) { mark_box(lo, hi); continue; }
        int le = 0; double best = -1;
        for (int e2 = 0; e2 < 3; e2++) {
          const double* a = c.p + 3 * e2; const double* b = c.p + 3 * ((e2 + 1) % 3);
          double l2 = (a[0] - b[0]) * (a[0] - b[0]) + (a[1] - b[1]) * (a[1] - b[1]) + (a[2] - b[2]) * (a[2] - b[2]);
          if (l2 > best) { best = l2; le = e2; }
        }
        const int i0 = le, i1 = (le + 1) % 3, i2 = (le + 2) % 3;
        double m[3]; for (int k = 0; k < 3; k++) m[k] = 0.5 * (c.p[3 * i0 + k] + c.p[3 * i1 + k]);
        Tri x, y2; x.depth = y2.depth = c.depth + 1;
        memcpy(x.p, c.p + 3 * i0, 24); memcpy(x.p + 3, m, 24); memcpy(x.p + 6, c.p + 3 * i2, 24);
        memcpy(y2.p, m, 24); memcpy(y2.p + 3, c.p + 3 * i1, 24); memcpy(y2.p + 6, c.p + 3 * i2, 24);
        st.push_back(x); st.push_back(y2);
      }
    }
  } else {
    const size_t np = elems.size() / 4;
    for (size_t i = 0; i < np; i++) {
      const double* s = &elems[4 * i];
      double lo[3] = {s[0] - s[3], s[1] - s[3], s[2] - s[3]}, hi[3] = {s[0] + s[3], s[1] + s[3], s[2] + s[3]};
      mark_box(lo, hi);
    }
  }
  // pass x: two scans per row
  parallel_for(ny * nz, [&](int r0, int r1) {
    for (int r = r0; r < r1; r++) {
      int32_t* row = &D[(size_t)r * nx];
      int last = -1000000;
      for (int x = 0; x < nx; x++) { if (row[x] == 0) last = x; else { const int64_t d = x - last; row[x] = d * d >= INF ? INF : (int32_t)(d * d); } }
      last = 1000000 + nx;
      for (int x = nx - 1; x >= 0; x--) { if (row[x] == 0) last = x; else { const int64_t d = last - x; if (d * d < row[x]) row[x] = (int32_t)(d * d); } }
    }
  });
  // pass y (stride nx) and pass z (stride nx*ny): lower envelopes
  parallel_for(nz, [&](int z0, int z1) {
    std::vector<int32_t> f(ny), d(ny); std::vector<int> v(ny); std::vector<double> zz(ny + 1);
    for (int z = z0; z < z1; z++) for (int x = 0; x < nx; x++) {
      int32_t* col = &D[(size_t)z * ny * nx + x];
      for (int y = 0; y < ny; y++) f[y] = col[(size_t)y * nx];
      edt_1d(f.data(), ny, d.data(), v.data(), zz.data());
      for (int y = 0; y < ny; y++) col[(size_t)y * nx] = d[y];
    }
  });
  parallel_for(ny, [&](int y0, int y1) {
    std::vector<int32_t> f(nz), d(nz); std::vector<int> v(nz); std::vector<double> zz(nz + 1);
    const size_t sz = (size_t)nx * ny;
    for (int y = y0; y < y1; y++) for (int x = 0; x < nx; x++) {
      int32_t* col = &D[(size_t)y * nx + x];
      for (int z = 0; z < nz; z++) f[z] = col[(size_t)z * sz];
      edt_1d(f.data(), nz, d.data(), v.data(), zz.data());
      for (int z = 0; z < nz; z++) col[(size_t)z * sz] = d[z];
    }
  });
  G.q.resize(nvox);
  parallel_for(nz, [&](int z0, int z1) {
    for (size_t i = (size_t)z0 * nx * ny; i < (size_t)z1 * nx * ny; i++) {
      const double c = std::sqrt((double)D[i]) - 1.7320508075688772 - 0.02;
      G.q[i] = (uint8_t)(c <= 0 ? 0 : std::min(255.0, std::floor(4.0 * c)));
    }
  });
}

}  // namespace

#define KB_RAY_TLAS_DEPTH 36   // depth bound of the top-level ray hierarchy (kb_raycast.cu keeps a 40-entry stack for it)
#define KB_STAGES 8          // pieces a large host batch is uploaded / checked in (feasible_batch_host_one)
struct kb_engine {
  // ---- description
  std::vector<Geom> geoms;
  std::vector<int> terrains;
  std::vector<int> objects; std::vector<Xf> objT;
  int L = 0;
  std::vector<int32_t> parents; std::vector<uint8_t> linktype; std::vector<double> axis, T0, qmin, qmax;
  std::vector<int> linkgeom;
  std::vector<uint8_t> jtype; std::vector<int32_t> jlink; std::vector<int16_t> jidx;
  std::vector<Driver> drivers;
  std::vector<uint8_t> selfcol; bool selfcol_default = true;
  std::vector<uint8_t> mask; int nids = 0; bool mask_user = false;
  bool finalized = false;
  struct StaticAlloc { size_t member_offset, bytes; };
  std::vector<StaticAlloc> statics;          // device allocations of the read-only data, for replication (kb_finalize_multi)
  std::vector<kb_engine*> replicas;          // further devices of a multi-device handle: full copies of the static data, own streams and scratch
  int64_t multi_min = 8192;                  // host-buffer batches below this stay on the first device
  // ---- device
  int device = -1, num_sms = 148;
  cudaStream_t own_stream = nullptr, stream = nullptr, copy_stream = nullptr, aux_stream = nullptr;
  cudaEvent_t ev_stage[KB_STAGES] = {};
  cudaEvent_t ev_copy[4] = {nullptr, nullptr, nullptr, nullptr};
  std::vector<float> h_nodes;               // 8 floats per node
  std::vector<float> h_wide;                // 8 floats per slot, 4 slots per wide node
  float4* d_wide = nullptr; int wide = 1;     // option wide: 1 = the boolean query runs on the 4-wide hierarchies when every item has them
  std::vector<float> h_tris32; std::vector<double> h_tris64; std::vector<int32_t> h_triown, h_triorig;
  std::vector<float> h_sph32; std::vector<double> h_sph64; std::vector<int32_t> h_sphown, h_sphorig;
  std::vector<float> h_box32; std::vector<double> h_box64; std::vector<int32_t> h_boxown;
  std::vector<DevGeom> dsolid;              // per registered geometry: the solid of a box primitive (empty otherwise)
  std::vector<DevGeom> dgeoms;              // per registered geometry (local frame)
  std::vector<DevGeom> groups;              // merged environment groups (world frame)
  KbScene scene{};
  float4* d_nodes = nullptr; float4* d_tris32 = nullptr; double* d_tris64 = nullptr; float4* d_sph32 = nullptr; double* d_sph64 = nullptr;
  int32_t* d_triown = nullptr; int32_t* d_sphown = nullptr; int32_t* d_triorig = nullptr; int32_t* d_sphorig = nullptr;
  float4* d_box32 = nullptr; double* d_box64 = nullptr; int32_t* d_boxown = nullptr;
  KbRobotDev* d_robot = nullptr; KbDriverDev* d_drv = nullptr; int32_t* d_drv_link = nullptr; double* d_drv_scale = nullptr; double* d_drv_off = nullptr;
  ItemSet feas_items, env_items;            // env + self ; env only (distance without self)
  struct DynCloud { int geom, group, owner, cap; double radius; double T[12]; };
  std::vector<DynCloud> dyn;                // replaceable point clouds: their own environment groups, rebuilt on the GPU (kb_lbvh.cu)
  double* d_dyn_pts = nullptr; double* d_dyn_T = nullptr; void* d_dyn_scratch = nullptr; int64_t dyn_pts_cap = 0; size_t dyn_scratch_bytes = 0;
  double eps_extent = 0, eps_reach = 0, eps_lmax = 0;   // the parts of the fp32 error bound, kept so a new cloud can widen it
  std::vector<HostGrid> hgrids; uint8_t* d_grid[KB_MAX_GRIDS] = {nullptr, nullptr, nullptr, nullptr};
  int cloud_builder = 0;                     // 0: point-cloud hierarchies by binned SAH on the host; 1: linear BVH on the GPU (kb_lbvh.cu)
  int mesh_builder = 0;                      // the same choice for triangle meshes above KB_GPU_MESH_MIN triangles
  struct PendingCloud { DevGeom* dg; std::vector<double> elems; std::vector<int32_t> owners; bool mesh = false; std::vector<int32_t> origs; };
  std::vector<PendingCloud> pending_clouds;   // clouds whose hierarchy is built on the GPU once the arrays are uploaded
  // small host-buffer batches (N <= graph_max): pinned staging + one CUDA graph per batch size (copy in, FK, traversal, finish, copy out)
  struct SmallGraph { int64_t n = 0; int seen = 0; cudaGraphExec_t exec = nullptr; const void* key[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr}; };
  int64_t zero_copy_max = 64;                // small batches up to this size read / write the pinned staging buffers directly (option zero_copy_max)
  std::vector<SmallGraph> graphs; int64_t graph_max = 16384; double* h_pin_in = nullptr; uint8_t* h_pin_out = nullptr; double* g_dQ = nullptr; uint8_t* g_dout = nullptr;
  int cloud_leaf = 8;                        // points per leaf of a host-built point-cloud hierarchy (option cloud_leaf, 1..32)
  int both_limit = 0;                        // experiment: frontier size up to which comparable inner pairs descend both trees at once
  int grid_res = 256; bool use_grids = false; // clearance-grid broad phase of the boolean query (options grid_res, clear_grid)
  int64_t static_bytes = 0;
  // ---- per-batch scratch (grown on demand)
  int64_t chunk = 65536;
  double* d_xf = nullptr; int64_t xf_cap = 0;
  uint8_t* d_state = nullptr; int32_t* d_hit = nullptr; int32_t* d_hit_elem = nullptr; int64_t cfg_cap = 0;
  uint4* d_leaf_list = nullptr; int64_t leaf_cap = 0; uint8_t* d_flagged = nullptr; uint8_t* d_state2 = nullptr; int64_t split_cap = 0;
  int pipeline = 0;                        // 0 = fused kernel (default, faster on every measured workload); 1 = split pipeline (node kernel ->
                                           // global leaf-pair list -> leaf kernel -> fused kernel on the requeued configurations)
  int leaf_budget = 64; int leaf_slots = 40;
  uint32_t* d_work = nullptr; unsigned long long* d_counters = nullptr;   // counters: [0] recheck [1] node [2] leaf [3] feasible [4] visible
  double* d_Q = nullptr; int64_t q_cap = 0;             // staging for host entry points
  float* d_Qf = nullptr; int64_t qf_cap = 0;            // fp32 configurations of kb_feasible_batch_f32, widened into d_Q
  uint8_t* d_out = nullptr; int64_t out_cap = 0;
  uint32_t* d_bits = nullptr; int64_t bits_cap = 0;   // packed result bitmask of the *_bits entry points
  int32_t* d_pair = nullptr; int64_t pair_cap = 0;
  double* d_dist = nullptr; int64_t dist_cap = 0;
  double* d_cp = nullptr; int64_t cp_cap = 0;             // closest points of the *_ex distance entry points
  // edges
  double* d_A = nullptr; double* d_B = nullptr; int64_t ab_cap = 0;
  int32_t* d_nlev = nullptr; uint8_t* d_alive = nullptr; int32_t* d_nchecks = nullptr; int32_t* d_firstbad = nullptr; int32_t* d_list = nullptr; int64_t edge_cap = 0;
  double* d_eQ = nullptr; uint8_t* d_efeas = nullptr; int64_t eq_cap = 0;
  uint8_t* d_eslot = nullptr; int64_t eslot_cap = 0; int64_t edge_flat_max = 1 << 18;   // small edge batches: all midpoints in one launch (option edge_flat_max, 0 = never)
  int32_t* d_scalars = nullptr;             // [0] maxlev, [1] list count
  double* d_weights = nullptr; int64_t w_cap = 0;
  // generic transform-pair queries
  double* d_T = nullptr; int64_t t_cap = 0;
  // ray casting (kb_raycast.cu): bodies + top-level hierarchy are static data, the rest per-call scratch
  std::vector<KbRayBody> ray_bodies; int ray_nlink = 0, ray_nstatic = 0; std::vector<float> h_tlas; float tlas_ext = 0; double ray_max_margin = 0;
  KbRayBody* d_raybodies = nullptr; float4* d_tlas = nullptr;
  double* d_rays = nullptr; int32_t* d_rid = nullptr; double* d_rdist = nullptr; int32_t* d_relem = nullptr; int64_t ray_cap = 0;
  uint8_t* d_ignore = nullptr; double* d_rayq = nullptr; KbRayBody* d_onebody = nullptr;
  int ray_variant = 0, ray_tile = 1;      // experiment knobs of the ray kernel (options ray_variant, ray_tile)
  // ---- stats
  kb_stats stats{}; int64_t edge_cfg_host = 0;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  bool collect_stats = false, time_kernels = false;
  std::vector<cudaEvent_t> tev;             // event pairs around traversal launches (time_kernels)
  size_t tev_used = 0;
};

namespace {

int robot_id(const kb_engine* e) { return (int)e->terrains.size() + (int)e->objects.size(); }
int link_id(const kb_engine* e, int j) { return robot_id(e) + 1 + j; }
int num_ids(const kb_engine* e) { return (int)e->terrains.size() + (int)e->objects.size() + (e->L ? 1 + e->L : 0); }
bool geom_empty(const kb_engine* e, int g) { return g < 0 || g >= (int)e->geoms.size() || e->geoms[g].empty(); }

void init_all_self_collisions(kb_engine* e) {
  int L = e->L;
  for (int i = 0; i < L; i++) for (int j = i + 1; j < L; j++)
    e->selfcol[i * L + j] = !geom_empty(e, e->linkgeom[i]) && !geom_empty(e, e->linkgeom[j]) && e->parents[j] != i && e->parents[i] != j;
}

// WorldPlannerSettings::InitializeDefault (reference Cpp/Planning/PlannerSettings.cpp:16-41)
void init_default_mask(kb_engine* e) {
  int n = num_ids(e); e->nids = n;
  e->mask.assign((size_t)n * n, 1);
  for (int i = 0; i < n; i++) e->mask[(size_t)i * n + i] = 0;
  if (e->L) {
    int k = robot_id(e), base = k + 1, L = e->L, T = (int)e->terrains.size();
    e->mask[(size_t)k * n + k] = 1;
    for (int j = 0; j < L; j++) { e->mask[(size_t)(base + j) * n + k] = 0; e->mask[(size_t)k * n + base + j] = 0; }
    for (int j = 0; j < L; j++) for (int m = 0; m < L; m++) e->mask[(size_t)(base + j) * n + base + m] = (j < m) ? e->selfcol[j * L + m] : 0;
    for (int j = 0; j < L; j++) if (e->parents[j] == -1)
      for (int t = 0; t < T; t++) { e->mask[(size_t)(base + j) * n + t] = 0; e->mask[(size_t)t * n + base + j] = 0; }
  }
}
inline bool mask_en(const kb_engine* e, int a, int b) { return e->mask[(size_t)a * e->nids + b] != 0; }

// appends one geometry (elements given in some frame) to the host arrays: builds its BVH, writes nodes + elements
int append_geom(kb_engine* e, int kind, const std::vector<double>& elems, const std::vector<int32_t>& owners, double margin, DevGeom& dg, bool want_cover,
                const std::vector<int32_t>& origs = std::vector<int32_t>()) {
  const int stride = kind == G_MESH ? 9 : (kind == G_BOX ? 15 : 4);
  const int n = (int)(elems.size() / stride);
  dg = DevGeom(); dg.margin = margin; dg.kind = kind == G_MESH ? KB_ELEM_TRI : (kind == G_BOX ? KB_ELEM_BOX : KB_ELEM_SPHERE); dg.nelem = n; dg.empty = n == 0;
  if (n == 0) return KB_OK;
  std::vector<double> elo(3 * (size_t)n), ehi(3 * (size_t)n);
  for (int i = 0; i < n; i++) {
    const double* p = &elems[(size_t)stride * i];
    for (int k = 0; k < 3; k++) {
      if (kind == G_MESH) { elo[3 * (size_t)i + k] = std::min(p[k], std::min(p[3 + k], p[6 + k])); ehi[3 * (size_t)i + k] = std::max(p[k], std::max(p[3 + k], p[6 + k])); }
      else if (kind == G_BOX) {      // extent of the oriented box along axis k: sum |R[k][j]| h_j (a hair up: the node box must contain it)
        const double ext = (std::fabs(p[3 + 3 * k]) * p[12] + std::fabs(p[3 + 3 * k + 1]) * p[13] + std::fabs(p[3 + 3 * k + 2]) * p[14]) * (1 + 1e-12);
        elo[3 * (size_t)i + k] = p[k] - ext; ehi[3 * (size_t)i + k] = p[k] + ext;
      }
      else { elo[3 * (size_t)i + k] = p[k] - p[3]; ehi[3 * (size_t)i + k] = p[k] + p[3]; }
    }
    if (kind != G_MESH && kind != G_BOX) dg.rmax = std::max(dg.rmax, p[3]);
  }
  Bvh bvh; build_bvh(elo, ehi, n, kind == G_MESH ? 1 : (kind == G_BOX ? 8 : e->cloud_leaf), bvh);
  if ((e->h_nodes.size() / 8) & 1) e->h_nodes.insert(e->h_nodes.end(), 8, 0.f);                // even node base: sibling pairs share a 64 B line
  dg.node_base = (int)(e->h_nodes.size() / 8); dg.nnodes = (int)bvh.nodes.size(); dg.depth = bvh.depth;
  dg.elem_base = kind == G_MESH ? (int)(e->h_tris64.size() / 9) : (kind == G_BOX ? (int)(e->h_box64.size() / 16) : (int)(e->h_sph64.size() / 4));
  memcpy(dg.lo, bvh.nodes[0].lo, 24); memcpy(dg.hi, bvh.nodes[0].hi, 24);
  for (const BNode& nd : bvh.nodes) {
    float v[8];
    for (int k = 0; k < 3; k++) {   // centre / half extent; the fp32 box must contain the fp64 one
      float c = (float)(0.5 * (nd.lo[k] + nd.hi[k]));
      v[k] = c; v[4 + k] = f_up(std::max(nd.hi[k] - (double)c, (double)c - nd.lo[k]));
    }
    v[3] = v[7] = 0.f;
    if (nd.left >= 0) { v[3] = i2f(nd.left); v[7] = i2f(0); }
    else { v[3] = i2f(~nd.first); v[7] = i2f(nd.count); }
    e->h_nodes.insert(e->h_nodes.end(), v, v + 8);
  }
  if (kind != G_BOX) {
    // ---- the same hierarchy 4 wide (kb_traverse_wide_kernel).  A wide node = 4 slots = 128 B = one cache line; a slot is a node's box
    // plus a reference: >= 0 the wide node (index relative to this geometry, in wide nodes) that holds its grandchildren -- or children
    // that are leaves --, < 0 a leaf (~first element, count in the second word).  Wide node 0 is a super root whose slot 0 is the root.
    // Four lanes read the four slots of one node with one line fetch: the traversal touches 2 lines per 4 box tests instead of 2 per 1-2.
    if ((e->h_wide.size() / 8) & 3) e->h_wide.insert(e->h_wide.end(), (4 - ((e->h_wide.size() / 8) & 3)) * 8, 0.f);
    dg.wide_base = (int)(e->h_wide.size() / 8);
    const size_t base = e->h_wide.size();
    auto empty_slot = [&](size_t at) { float* v = &e->h_wide[at]; v[0] = v[1] = v[2] = 0.f; v[3] = i2f(~0); v[4] = v[5] = v[6] = -1e30f; v[7] = i2f(0); };
    auto alloc_wide = [&]() { const size_t w = (e->h_wide.size() - base) / 32; e->h_wide.insert(e->h_wide.end(), 32, 0.f); for (int k = 0; k < 4; k++) empty_slot(base + w * 32 + 8 * k); return (int)w; };
    struct Todo { int bnode; int wnode; int depth; };      // fill wide node `wnode` with the (grand)children of binary node `bnode`
    std::vector<Todo> todo;
    auto write_slot = [&](int wnode, int k, int bn, int depth) {
      const BNode& nd = bvh.nodes[bn];
      float v[8];
      for (int a = 0; a < 3; a++) { const float c = (float)(0.5 * (nd.lo[a] + nd.hi[a])); v[a] = c; v[4 + a] = f_up(std::max(nd.hi[a] - (double)c, (double)c - nd.lo[a])); }
      if (nd.left >= 0) { const int w = alloc_wide(); v[3] = i2f(w); v[7] = i2f(0); todo.push_back({bn, w, depth + 1}); }
      else { v[3] = i2f(~nd.first); v[7] = i2f(nd.count); }
      memcpy(&e->h_wide[base + (size_t)wnode * 32 + 8 * k], v, 32);       // (after alloc_wide: the vector may have moved)
    };
    const int w0 = alloc_wide();
    write_slot(w0, 0, 0, 0);
    dg.wdepth = 1;
    while (!todo.empty()) {
      const Todo t = todo.back(); todo.pop_back();
      dg.wdepth = std::max(dg.wdepth, t.depth + 1);
      int k = 0;
      const int l = bvh.nodes[t.bnode].left;
      for (int c = l; c <= l + 1; c++) {
        if (bvh.nodes[c].left < 0) write_slot(t.wnode, k++, c, t.depth);
        else { write_slot(t.wnode, k++, bvh.nodes[c].left, t.depth); write_slot(t.wnode, k++, bvh.nodes[c].left + 1, t.depth); }
      }
    }
    dg.wide_slots = (int)((e->h_wide.size() - base) / 8);
  }
  if (want_cover) {
    // up to KB_COVER_MAX spheres that together contain the geometry: a cut of the BVH grown by always splitting the
    // node with the largest sphere; sphere = box centre + the farthest element point
    struct CS { int node; double c[3], r; };
    auto sphere_of = [&](int node) {
      const BNode& nd = bvh.nodes[node];
      CS s; s.node = node; s.r = 0;
      for (int k = 0; k < 3; k++) s.c[k] = 0.5 * (nd.lo[k] + nd.hi[k]);
      for (int i = nd.first; i < nd.first + nd.count; i++) {
        const double* p = &elems[(size_t)stride * bvh.perm[i]];
        if (kind == G_MESH) {
          for (int v = 0; v < 3; v++) { double d2 = 0; for (int k = 0; k < 3; k++) d2 += (p[3 * v + k] - s.c[k]) * (p[3 * v + k] - s.c[k]); s.r = std::max(s.r, std::sqrt(d2)); }
        } else { double d2 = 0; for (int k = 0; k < 3; k++) d2 += (p[k] - s.c[k]) * (p[k] - s.c[k]); s.r = std::max(s.r, std::sqrt(d2) + p[3]); }
      }
      return s;
    };
    std::vector<CS> cut; cut.push_back(sphere_of(0));
    while ((int)cut.size() < KB_COVER_MAX) {
      int best = -1;
      for (int i = 0; i < (int)cut.size(); i++) if (bvh.nodes[cut[i].node].left >= 0 && (best < 0 || cut[i].r > cut[best].r)) best = i;
      if (best < 0) break;
      // stop once the largest sphere is itself a leaf-sized one
      bool any_larger_leaf = false;
      for (const CS& c : cut) if (bvh.nodes[c.node].left < 0 && c.r >= cut[best].r) any_larger_leaf = true;
      if (any_larger_leaf) break;
      const int l = bvh.nodes[cut[best].node].left;
      cut[best] = sphere_of(l); cut.push_back(sphere_of(l + 1));
    }
    dg.ncover = (int)cut.size();
    for (int i = 0; i < dg.ncover; i++) { memcpy(dg.cover[i], cut[i].c, 24); dg.cover[i][3] = cut[i].r * (1 + 1e-12) + 1e-300; }
  }
  for (int i = 0; i < n; i++) {
    int src = bvh.perm[i];
    const double* p = &elems[(size_t)stride * src];
    int32_t own = owners.empty() ? -1 : owners[src];
    if (kind == G_MESH) {
      e->h_tris64.insert(e->h_tris64.end(), p, p + 9);
      for (int v = 0; v < 3; v++) { float f[4] = {(float)p[3 * v], (float)p[3 * v + 1], (float)p[3 * v + 2], v == 0 ? i2f(own) : 0.f}; e->h_tris32.insert(e->h_tris32.end(), f, f + 4); }
      e->h_triown.push_back(own); e->h_triorig.push_back(origs.empty() ? src : origs[src]);
    } else if (kind == G_BOX) {      // {centre, hx} {axis0, hy} {axis1, hz} {axis2, 0}: axis j = column j of the row-major 3x3
      double b[16] = {p[0], p[1], p[2], p[12], p[3], p[6], p[9], p[13], p[4], p[7], p[10], p[14], p[5], p[8], p[11], 0.0};
      e->h_box64.insert(e->h_box64.end(), b, b + 16);
      for (int k = 0; k < 16; k++) e->h_box32.push_back((float)b[k]);
      e->h_boxown.push_back(own);
    } else {
      e->h_sph64.insert(e->h_sph64.end(), p, p + 4);
      float f[4] = {(float)p[0], (float)p[1], (float)p[2], (float)p[3]}; e->h_sph32.insert(e->h_sph32.end(), f, f + 4);
      e->h_sphown.push_back(own); e->h_sphorig.push_back(origs.empty() ? src : origs[src]);
    }
  }
  return KB_OK;
}

// reserves nodes / elements of a point cloud whose hierarchy the GPU builds after the upload (option cloud_builder = 1)
void reserve_gpu_cloud(kb_engine* e, const std::vector<double>& elems, const std::vector<int32_t>& owners, double margin, DevGeom& dg,
                       const std::vector<int32_t>& origs = std::vector<int32_t>()) {
  const int n = (int)(elems.size() / 4);
  dg = DevGeom(); dg.margin = margin; dg.kind = KB_ELEM_SPHERE; dg.nelem = n; dg.empty = n == 0; dg.depth = 64;
  for (int k = 0; k < 3; k++) { dg.lo[k] = 1e300; dg.hi[k] = -1e300; }
  for (int i = 0; i < n; i++) {
    const double* p = &elems[4 * (size_t)i];
    dg.rmax = std::max(dg.rmax, p[3]);
    for (int k = 0; k < 3; k++) { dg.lo[k] = std::min(dg.lo[k], p[k] - p[3]); dg.hi[k] = std::max(dg.hi[k], p[k] + p[3]); }
  }
  if ((e->h_nodes.size() / 8) & 1) e->h_nodes.insert(e->h_nodes.end(), 8, 0.f);
  dg.node_base = (int)(e->h_nodes.size() / 8); dg.nnodes = (int)kb_lbvh_nodes_for(n);
  const float v[8] = {0.f, 0.f, 0.f, i2f(~0), -1e30f, -1e30f, -1e30f, i2f(0)};
  e->h_nodes.insert(e->h_nodes.end(), v, v + 8);
  e->h_nodes.insert(e->h_nodes.end(), (size_t)(dg.nnodes - 1) * 8, 0.f);
  dg.elem_base = (int)(e->h_sph64.size() / 4);
  e->h_sph64.insert(e->h_sph64.end(), (size_t)n * 4, 0.0); e->h_sph32.insert(e->h_sph32.end(), (size_t)n * 4, 0.f);
  e->h_sphown.insert(e->h_sphown.end(), (size_t)n, -1); e->h_sphorig.insert(e->h_sphorig.end(), (size_t)n, -1);
  // covering spheres of the clearance-grid broad phase: one sphere round the bounds is enough for a cloud used as a link geometry
  dg.ncover = 1;
  double r2 = 0; for (int k = 0; k < 3; k++) { dg.cover[0][k] = 0.5 * (dg.lo[k] + dg.hi[k]); r2 += 0.25 * (dg.hi[k] - dg.lo[k]) * (dg.hi[k] - dg.lo[k]); }
  dg.cover[0][3] = std::sqrt(r2) * (1 + 1e-12);
  e->pending_clouds.push_back({&dg, elems, owners, false, origs});
}

// the same for a large triangle mesh (option mesh_builder = 1)
void reserve_gpu_mesh(kb_engine* e, const std::vector<double>& elems, const std::vector<int32_t>& owners, double margin, DevGeom& dg,
                      const std::vector<int32_t>& origs = std::vector<int32_t>()) {
  const int n = (int)(elems.size() / 9);
  dg = DevGeom(); dg.margin = margin; dg.kind = KB_ELEM_TRI; dg.nelem = n; dg.empty = n == 0; dg.depth = 64;
  for (int k = 0; k < 3; k++) { dg.lo[k] = 1e300; dg.hi[k] = -1e300; }
  for (size_t i = 0; i < elems.size(); i++) { const int k = (int)(i % 3); dg.lo[k] = std::min(dg.lo[k], elems[i]); dg.hi[k] = std::max(dg.hi[k], elems[i]); }
  if ((e->h_nodes.size() / 8) & 1) e->h_nodes.insert(e->h_nodes.end(), 8, 0.f);
  dg.node_base = (int)(e->h_nodes.size() / 8); dg.nnodes = (int)kb_lbvh_nodes_for(n);
  const float v[8] = {0.f, 0.f, 0.f, i2f(~0), -1e30f, -1e30f, -1e30f, i2f(0)};
  e->h_nodes.insert(e->h_nodes.end(), v, v + 8);
  e->h_nodes.insert(e->h_nodes.end(), (size_t)(dg.nnodes - 1) * 8, 0.f);
  dg.elem_base = (int)(e->h_tris64.size() / 9);
  e->h_tris64.insert(e->h_tris64.end(), (size_t)n * 9, 0.0); e->h_tris32.insert(e->h_tris32.end(), (size_t)n * 12, 0.f);
  e->h_triown.insert(e->h_triown.end(), (size_t)n, -1); e->h_triorig.insert(e->h_triorig.end(), (size_t)n, -1);
  dg.ncover = 1;
  double r2 = 0; for (int k = 0; k < 3; k++) { dg.cover[0][k] = 0.5 * (dg.lo[k] + dg.hi[k]); r2 += 0.25 * (dg.hi[k] - dg.lo[k]) * (dg.hi[k] - dg.lo[k]); }
  dg.cover[0][3] = std::sqrt(r2) * (1 + 1e-12);
  e->pending_clouds.push_back({&dg, elems, owners, true, origs});
}

KbItem make_item(const DevGeom& A, int xfA, int idA, const DevGeom& B, int xfB, int idB, bool self) {
  KbItem it; memset(&it, 0, sizeof it);
  it.nodeA = A.node_base; it.nodeB = B.node_base; it.elemA = A.elem_base; it.elemB = B.elem_base;
  it.xfA = (int16_t)xfA; it.xfB = (int16_t)xfB; it.kindA = (uint8_t)A.kind; it.kindB = (uint8_t)B.kind; it.flags = self ? 1 : 0;
  it.idA = idA; it.idB = idB; it.thr = A.margin + B.margin; it.marg = A.margin + B.margin; it.rsum = A.rmax + B.rmax;
  it.margA = A.margin; it.margB = B.margin;
  it.wideA = A.wide_base; it.wideB = B.wide_base;
  return it;
}
// the A side of an item is limited to 2^20 nodes (stack entry packing); put the smaller hierarchy there
int add_item(ItemSet& set, const DevGeom& A, int xfA, int idA, const DevGeom& B, int xfB, int idB, bool self) {
  if (A.empty || B.empty) return KB_OK;
  const bool swap = A.nnodes > B.nnodes;
  const DevGeom& a = swap ? B : A; const DevGeom& b = swap ? A : B;
  if (a.nnodes >= KB_MAX_NODES_A) return fail(KB_ERR_UNSUPPORTED, "both geometries of a pair have more than %d BVH nodes", KB_MAX_NODES_A);
  if ((int)set.items.size() >= KB_MAX_ITEMS) return fail(KB_ERR_UNSUPPORTED, "more than %d geometry pairs per configuration", KB_MAX_ITEMS);
  set.items.push_back(swap ? make_item(B, xfB, idB, A, xfA, idA, self) : make_item(A, xfA, idA, B, xfB, idB, self));
  set.maxdepth = std::max(set.maxdepth, a.depth + b.depth);
  if (a.wide_base < 0 || b.wide_base < 0 || a.wide_slots >= KB_MAX_NODES_A) set.all_wide = false;
  set.wdepth = std::max(set.wdepth, a.wdepth + b.wdepth);
  return KB_OK;
}

// rec: the engine whose member `dptr` is -- the allocation is remembered (member offset, size) so that kb_finalize_multi can replicate
// the static data on further devices with peer copies instead of rebuilding it
template <class T> int upload(T*& dptr, const void* src, size_t bytes, int64_t* total, kb_engine* rec = nullptr) {
  dptr = nullptr;
  if (rec) rec->statics.push_back({(size_t)((char*)&dptr - (char*)rec), bytes ? bytes : 16});
  if (bytes == 0) { CK(cudaMalloc((void**)&dptr, 16)); return KB_OK; }
  CK(cudaMalloc((void**)&dptr, bytes));
  CK(cudaMemcpy(dptr, src, bytes, cudaMemcpyHostToDevice));
  if (total) *total += (int64_t)bytes;
  return KB_OK;
}
template <class T> int grow(T*& p, int64_t& cap, int64_t need) {
  if (need <= cap) return KB_OK;
  if (p) cudaFree(p);
  p = nullptr; cap = 0;
  CK(cudaMalloc((void**)&p, (size_t)need * sizeof(T)));
  cap = need;
  return KB_OK;
}
template <class T> int grow_plain(T*& p, int64_t need_elems, int64_t& cap_elems) { return grow(p, cap_elems, need_elems); }

// per-configuration scratch for `n` configurations per launch (grown on demand, never shrunk)
int ensure_cfg_scratch(kb_engine* e, int nxf, int64_t n) {
  int64_t ch = std::max<int64_t>(1024, std::min(e->chunk, n));
  int64_t need_xf = ch * nxf * 12;
  if (need_xf > e->xf_cap) { if (e->d_xf) cudaFree(e->d_xf); e->d_xf = nullptr; e->xf_cap = 0; CK(cudaMalloc((void**)&e->d_xf, (size_t)need_xf * 8)); e->xf_cap = need_xf; }
  if (ch > e->cfg_cap) {
    if (e->d_state) cudaFree(e->d_state); if (e->d_hit) cudaFree(e->d_hit); if (e->d_hit_elem) cudaFree(e->d_hit_elem);
    e->cfg_cap = 0;
    CK(cudaMalloc((void**)&e->d_state, (size_t)ch)); CK(cudaMalloc((void**)&e->d_hit, (size_t)ch * 4)); CK(cudaMalloc((void**)&e->d_hit_elem, (size_t)ch * 8));
    e->cfg_cap = ch;
  }
  return KB_OK;
}

int ensure_split_scratch(kb_engine* e, int64_t n) {
  int64_t ch = std::max<int64_t>(1024, std::min(e->chunk, n));
  if (ch > e->split_cap) {
    if (e->d_flagged) cudaFree(e->d_flagged); if (e->d_state2) cudaFree(e->d_state2); if (e->d_leaf_list) cudaFree(e->d_leaf_list);
    e->d_flagged = e->d_state2 = nullptr; e->d_leaf_list = nullptr; e->split_cap = 0;
    CK(cudaMalloc((void**)&e->d_flagged, (size_t)ch)); CK(cudaMalloc((void**)&e->d_state2, (size_t)ch));
    e->leaf_cap = std::min<int64_t>(ch * e->leaf_slots, 0x7fffffff);
    CK(cudaMalloc((void**)&e->d_leaf_list, (size_t)e->leaf_cap * 16));
    e->split_cap = ch;
  }
  return KB_OK;
}

void begin_timing(kb_engine* e) { cudaEventRecord(e->ev0, e->stream); }
void end_timing(kb_engine* e, bool sync) {
  cudaEventRecord(e->ev1, e->stream);
  if (sync) { cudaEventSynchronize(e->ev1); float ms = 0; if (cudaEventElapsedTime(&ms, e->ev0, e->ev1) == cudaSuccess) e->stats.gpu_ms += ms; }
}

// resolves the pending traversal event pairs into stats.traverse_ms (needs the stream to have drained)
void fold_kernel_times(kb_engine* e) {
  for (size_t i = 0; i + 1 < e->tev_used; i += 2) {
    float ms = 0;
    if (cudaEventSynchronize(e->tev[i + 1]) == cudaSuccess && cudaEventElapsedTime(&ms, e->tev[i], e->tev[i + 1]) == cudaSuccess) {
      e->stats.traverse_ms += ms; e->stats.traverse_launches++;
    }
  }
  e->tev_used = 0;
}
cudaError_t timed_traverse(kb_engine* e, const KbTraverseParams& p, int mode, double* out_dist, double ub, float rel_err = 0.f, float abs_err = 0.f) {
  if (!e->time_kernels) return kb_launch_traverse(p, mode, out_dist, ub, e->num_sms, e->stream, nullptr, rel_err, abs_err);
  if (e->tev_used + 2 > 8192) fold_kernel_times(e);
  while (e->tev.size() < e->tev_used + 2) { cudaEvent_t ev; cudaError_t ce = cudaEventCreate(&ev); if (ce != cudaSuccess) return ce; e->tev.push_back(ev); }
  // the work-counter memset is part of kb_launch_traverse: keep it outside the bracket by issuing a no-op ordering point first
  cudaEventRecord(e->tev[e->tev_used], e->stream);
  cudaError_t ce = kb_launch_traverse(p, mode, out_dist, ub, e->num_sms, e->stream, nullptr, rel_err, abs_err);
  cudaEventRecord(e->tev[e->tev_used + 1], e->stream);
  e->tev_used += 2;
  return ce;
}

// split pipeline + fused fallback on the requeued configurations, bracketed like one traversal launch
cudaError_t timed_split(kb_engine* e, const KbTraverseParams& p, const KbSplitParams& q) {
  const bool timed = e->time_kernels;
  if (timed) {
    if (e->tev_used + 2 > 8192) fold_kernel_times(e);
    while (e->tev.size() < e->tev_used + 2) { cudaEvent_t ev; cudaError_t ce = cudaEventCreate(&ev); if (ce != cudaSuccess) return ce; e->tev.push_back(ev); }
    cudaEventRecord(e->tev[e->tev_used], e->stream);
  }
  cudaError_t ce = kb_launch_split(p, q, e->num_sms, e->stream);
  if (ce == cudaSuccess) {
    KbTraverseParams p2 = p; p2.state = q.state2;
    ce = kb_launch_traverse(p2, 0, nullptr, 0.0, e->num_sms, e->stream);
  }
  if (timed) { cudaEventRecord(e->tev[e->tev_used + 1], e->stream); e->tev_used += 2; }
  return ce;
}

KbTraverseParams make_params(kb_engine* e, const ItemSet& set, const double* xf, int64_t n, const uint8_t* state) {
  KbTraverseParams p; memset(&p, 0, sizeof p);
  p.scene = e->scene; p.items = set.d_items; p.nitems = (int)set.items.size(); p.nxf = set.nxf; p.xf64 = xf; p.N = n; p.state = state;
  p.hit = e->d_hit; p.hit_elem = e->d_hit_elem; p.work_counter = e->d_work; p.counters = e->d_counters;
  // Stack models.  Distance / all-pairs kernels: a 32-wide pop pushes at most 2 per entry, allowed while sp <= wide_limit; above it
  // one entry is popped at a time, which grows the stack by at most the BVH depth sum.  Boolean kernel: an entry pushes at most 4
  // (pairs of comparable boxes push all four child pairs), so m entries may be popped while sp + 3 m <= pop_room; single-entry
  // pops grow the stack by at most 3 entries per two tree levels (1.5 x the depth sum).
  p.wide_limit = KB_STACK_CAP - 32 - set.maxdepth - 2;
  p.pop_room = KB_STACK_CAP - (3 * set.maxdepth + 1) / 2 - 4;
  p.collect_stats = e->collect_stats ? 1 : 0;
  p.both_limit = e->both_limit;
  // 4-wide traversal: 8 entries (32 lanes) per pop while the stack has room for their <= 32 children plus a depth-first tail of
  // <= 3 entries per wide level; one entry per pop above that
  p.use_wide = (e->wide && set.all_wide && !e->use_grids && 3 * set.wdepth + 48 <= KB_STACK_CAP) ? 1 : 0;
  p.wide_room = KB_STACK_CAP - 32 - 3 * set.wdepth - 4;
  p.both_ratio = 16.f;
  for (const KbItem& it : set.items) if (!(it.flags & 1)) { p.both_ratio = 4.f; break; }      // any link-vs-environment item: 4
  for (const KbItem& it : set.items) if (it.kindA == KB_ELEM_BOX || it.kindB == KB_ELEM_BOX) { p.has_boxes = 1; break; }
  if (e->use_grids && set.d_probes && !set.probes.empty()) { p.probes = set.d_probes; p.nprobes = (int)set.probes.size(); p.always_on = set.d_always_on; }
  return p;
}

// feasibility of n configurations resident on the device: FK -> traversal -> finish, chunk by chunk
int run_feasible_small(kb_engine* e, const double* dQ, int64_t n, uint8_t* d_out, unsigned long long* d_nfeas, const uint8_t* d_alive);
#define KB_SMALL_MAX 16384      // batches up to this size are checked one warp per configuration (run_feasible_small)

int run_feasible_device(kb_engine* e, const double* dQ, int64_t N, uint8_t* d_out, int32_t* d_first_pair, unsigned long long* d_nfeas, const uint8_t* d_alive = nullptr) {
  if (N <= KB_SMALL_MAX && !d_first_pair && e->pipeline == 0 && !e->collect_stats) return run_feasible_small(e, dQ, N, d_out, d_nfeas, d_alive);
  int rc = ensure_cfg_scratch(e, e->feas_items.nxf, N); if (rc) return rc;
  for (int64_t off = 0; off < N; off += e->chunk) {
    int64_t n = std::min(e->chunk, N - off);
    CK(kb_launch_fk(e->d_robot, e->d_drv, e->d_drv_link, e->d_drv_scale, e->d_drv_off, dQ + off * e->L, n, e->d_xf, e->feas_items.nxf, e->d_state, d_alive ? d_alive + off : nullptr, e->d_hit, e->stream));
    e->stats.kernel_launches++;
    if (!e->feas_items.items.empty()) {
      KbTraverseParams p = make_params(e, e->feas_items, e->d_xf, n, e->d_state);
      if (e->pipeline == 1) {
        if ((rc = ensure_split_scratch(e, N))) return rc;
        KbSplitParams q; memset(&q, 0, sizeof q);
        q.leaf_list = e->d_leaf_list; q.leaf_count = e->d_counters + 6; q.leaf_cap = (unsigned)e->leaf_cap; q.flagged = e->d_flagged; q.state2 = e->d_state2;
        q.requeued = e->d_counters + 5; q.leaf_budget = e->leaf_budget;
        q.stack_cap = e->feas_items.maxdepth <= 96 ? 256 : KB_STACK_CAP; q.wide_limit = q.stack_cap - 32 - e->feas_items.maxdepth - 2;
        CK(timed_split(e, p, q));
        e->stats.kernel_launches += 4;
      } else {
        CK(timed_traverse(e, p, 0, nullptr, 0.0));
        e->stats.kernel_launches++;
      }
    }
    CK(kb_launch_finish(e->d_state, e->d_hit, e->d_hit_elem, e->feas_items.d_items, e->d_triown, e->d_sphown, e->d_boxown, n, d_out + off,
                        d_first_pair ? d_first_pair + 2 * off : nullptr, d_nfeas, e->stream));
    e->stats.kernel_launches++;
  }
  return KB_OK;
}

// Small batches: FK, then the traversal with one warp per configuration, which writes the result bytes itself (no counter reset, no
// finish kernel): two kernels per call.  Split-pipeline / statistics runs do not come here (feasible_small's caller checks).
int run_feasible_small(kb_engine* e, const double* dQ, int64_t n, uint8_t* d_out, unsigned long long* d_nfeas, const uint8_t* d_alive) {
  const int nxf = e->feas_items.nxf;
  int rc = ensure_cfg_scratch(e, nxf, n); if (rc) return rc;
  CK(kb_launch_fk(e->d_robot, e->d_drv, e->d_drv_link, e->d_drv_scale, e->d_drv_off, dQ, n, e->d_xf, nxf, e->d_state, d_alive, e->d_hit, e->stream));
  e->stats.kernel_launches++;
  if (e->feas_items.items.empty()) {
    CK(kb_launch_finish(e->d_state, e->d_hit, e->d_hit_elem, e->feas_items.d_items, e->d_triown, e->d_sphown, e->d_boxown, n, d_out, nullptr, d_nfeas, e->stream));
    e->stats.kernel_launches++;
    return KB_OK;
  }
  KbTraverseParams p = make_params(e, e->feas_items, e->d_xf, n, e->d_state);
  p.static_sched = 1; p.out_bytes = d_out; p.nfeasible = d_nfeas;
  CK(timed_traverse(e, p, 0, nullptr, 0.0));
  e->stats.kernel_launches++;
  return KB_OK;
}

// One PIECE of a batch whose per-configuration scratch (transforms, limit bytes, hit records) was already sized for the whole batch:
// rows [soff, soff + n) of the scratch, its own work counter (wslot) and the stream it runs on.  Pieces of one batch on two alternating
// streams overlap where it matters: the next piece's CTAs fill the SMs the previous launch's tail (its few hardest configurations) has
// left idle -- every launch of the persistent traversal kernel ends with such a tail, ~0.1 ms on C2.
int run_feasible_piece(kb_engine* e, const double* dQ, int64_t soff, int64_t n, uint8_t* d_out, int32_t* d_first_pair, unsigned long long* d_nfeas, cudaStream_t st, int wslot) {
  const int nxf = e->feas_items.nxf;
  double* xf = e->d_xf + soff * nxf * 12;
  CK(kb_launch_fk(e->d_robot, e->d_drv, e->d_drv_link, e->d_drv_scale, e->d_drv_off, dQ, n, xf, nxf, e->d_state + soff, nullptr, e->d_hit + soff, st));
  e->stats.kernel_launches++;
  if (!e->feas_items.items.empty()) {
    KbTraverseParams p = make_params(e, e->feas_items, xf, n, e->d_state + soff);
    p.hit = e->d_hit + soff; p.hit_elem = e->d_hit_elem + 2 * soff; p.work_counter = e->d_work + wslot;
    CK(kb_launch_traverse(p, 0, nullptr, 0.0, e->num_sms, st));
    e->stats.kernel_launches++;
  }
  CK(kb_launch_finish(e->d_state + soff, e->d_hit + soff, e->d_hit_elem + 2 * soff, e->feas_items.d_items, e->d_triown, e->d_sphown, e->d_boxown, n, d_out,
                      d_first_pair, d_nfeas, st));
  e->stats.kernel_launches++;
  return KB_OK;
}

int upload_itemset(ItemSet& s, int64_t* total, kb_engine* rec = nullptr) {
  if (s.d_items) { cudaFree(s.d_items); s.d_items = nullptr; }
  if (s.d_probes) { cudaFree(s.d_probes); s.d_probes = nullptr; }
  if (s.d_always_on) { cudaFree(s.d_always_on); s.d_always_on = nullptr; }
  if (!s.probes.empty()) {
    int rc = upload(s.d_probes, s.probes.data(), s.probes.size() * sizeof(KbProbe), total, rec); if (rc) return rc;
    if ((rc = upload(s.d_always_on, s.always_on.data(), s.always_on.size() * 4, total, rec))) return rc;
  }
  return upload(s.d_items, s.items.data(), s.items.size() * sizeof(KbItem), total, rec);
}

// Multi-device handles (kb_finalize_multi): a host-buffer batch is cut into contiguous shards, one per device, each run by its own
// host thread on that device's replica (own stream, own scratch, static data replicated) and written straight into the caller's
// buffers -- SURVEY 8e: configurations shard naturally, no exchange step besides the results landing in one array.
template <class F> static int run_sharded(kb_engine* e, int64_t N, int64_t align, F fn) {
  const int nd = 1 + (int)e->replicas.size();
  if (nd == 1 || N < e->multi_min) return fn(e, (int64_t)0, N);
  int64_t per = (N + nd - 1) / nd; per = ((per + align - 1) / align) * align;
  std::vector<int> rcs((size_t)nd, KB_OK); std::vector<std::string> errs((size_t)nd);
  std::vector<std::thread> th;
  for (int k = 0; k < nd; k++) {
    const int64_t off = std::min(N, (int64_t)k * per), n = std::min(per, N - off);
    if (n <= 0) break;
    kb_engine* r = k == 0 ? e : e->replicas[(size_t)k - 1];
    th.emplace_back([&, k, r, off, n]() { rcs[(size_t)k] = fn(r, off, n); if (rcs[(size_t)k]) errs[(size_t)k] = g_err; });
  }
  for (auto& t : th) t.join();
  for (int k = 0; k < nd; k++) if (rcs[(size_t)k]) return fail(rcs[(size_t)k], "device %d: %s", k == 0 ? e->device : e->replicas[(size_t)k - 1]->device, errs[(size_t)k].c_str());
  return KB_OK;
}


}  // namespace

// =================================================================================================== C ABI
extern "C" {

const char* kb_last_error(void) { return g_err.c_str(); }
const char* kb_version(void) { return "klampt_b200 0.1 (sm_100a)"; }

int kb_engine_create(kb_engine** out) {
  if (!out) return fail(KB_ERR_INVALID, "null out pointer");
  *out = new kb_engine();
  return KB_OK;
}

void kb_engine_destroy(kb_engine* e) {
  if (!e) return;
  for (kb_engine* r : e->replicas) kb_engine_destroy(r);
  e->replicas.clear();
  if (e->device >= 0) {
    cudaSetDevice(e->device);
    void* ptrs[] = {e->d_nodes, e->d_tris32, e->d_tris64, e->d_sph32, e->d_sph64, e->d_triown, e->d_sphown, e->d_robot, e->d_drv, e->d_drv_link,
                    e->d_drv_scale, e->d_drv_off, e->feas_items.d_items, e->env_items.d_items, e->d_xf, e->d_state, e->d_hit, e->d_hit_elem, e->d_leaf_list, e->d_flagged, e->d_state2, e->d_work,
                    e->d_counters, e->d_Q, e->d_out, e->d_pair, e->d_dist, e->d_A, e->d_B, e->d_nlev, e->d_alive, e->d_nchecks, e->d_firstbad, e->d_list,
                    e->d_eQ, e->d_efeas, e->d_scalars, e->d_weights, e->d_T, e->feas_items.d_probes, e->feas_items.d_always_on,
                    e->d_grid[0], e->d_grid[1], e->d_grid[2], e->d_grid[3], e->d_box32, e->d_box64, e->d_boxown, e->d_dyn_pts, e->d_dyn_T, e->d_dyn_scratch, e->d_Qf, e->d_bits, e->d_triorig, e->d_sphorig, e->d_cp, e->d_eslot, e->d_wide,
                    e->d_raybodies, e->d_tlas, e->d_rays, e->d_rid, e->d_rdist, e->d_relem, e->d_ignore, e->d_rayq, e->d_onebody};
    for (void* p : ptrs) if (p) cudaFree(p);
    if (e->ev0) cudaEventDestroy(e->ev0);
    if (e->ev1) cudaEventDestroy(e->ev1);
    for (auto& x : e->graphs) if (x.exec) cudaGraphExecDestroy(x.exec);
    if (e->h_pin_in) cudaFreeHost(e->h_pin_in); if (e->h_pin_out) cudaFreeHost(e->h_pin_out);
    if (e->g_dQ) cudaFree(e->g_dQ); if (e->g_dout) cudaFree(e->g_dout);
    for (cudaEvent_t ev : e->tev) cudaEventDestroy(ev);
    for (int k = 0; k < 4; k++) if (e->ev_copy[k]) cudaEventDestroy(e->ev_copy[k]);
    for (int k = 0; k < KB_STAGES; k++) if (e->ev_stage[k]) cudaEventDestroy(e->ev_stage[k]);
    if (e->copy_stream) cudaStreamDestroy(e->copy_stream);
    if (e->aux_stream) cudaStreamDestroy(e->aux_stream);
    if (e->own_stream) cudaStreamDestroy(e->own_stream);
  }
  delete e;
}

// every coordinate that enters the scene must be a finite number: one NaN would poison the SAH build and every box above it
static bool all_finite(const double* p, size_t n) {
  for (size_t i = 0; i < n; i++) if (!std::isfinite(p[i])) return false;
  return true;
}

int kb_add_trimesh(kb_engine* e, const double* verts, int nv, const int32_t* tris, int nt, double margin) {
  if (!e || e->finalized) return fail(KB_ERR_STATE, "engine is null or already finalized");
  if (nt < 0 || nv < 0 || !(margin >= 0) || !std::isfinite(margin)) return fail(KB_ERR_INVALID, "negative size or margin");
  if (nt > 0 && (!verts || !tris)) return fail(KB_ERR_INVALID, "null vertex or triangle array");
  if (nt > 0 && !all_finite(verts, 3 * (size_t)nv)) return fail(KB_ERR_INVALID, "mesh has a non-finite vertex coordinate");
  Geom g; g.kind = nt > 0 ? G_MESH : G_EMPTY; g.margin = margin; g.tri.resize(9 * (size_t)nt);
  for (int t = 0; t < nt; t++) for (int v = 0; v < 3; v++) {
    int idx = tris[3 * t + v];
    if (idx < 0 || idx >= nv) return fail(KB_ERR_INVALID, "triangle %d references vertex %d of %d", t, idx, nv);
    for (int k = 0; k < 3; k++) g.tri[9 * (size_t)t + 3 * v + k] = verts[3 * (size_t)idx + k];
  }
  // A triangle whose area is below 1e-12 of its longest edge squared cannot be told from a segment in fp64 (the signs of the
  // orientation tests against its "plane" are rounding noise): it becomes the segment between its two farthest vertices, written as
  // the exactly degenerate triangle (p, q, q), which the predicates treat as a segment (degenerate_tri_tri).  The oracle applies
  // the same rule.
  for (int t = 0; t < nt; t++) {
    double* T = &g.tri[9 * (size_t)t];
    double e0[3], e1[3], e2[3];
    for (int k = 0; k < 3; k++) { e0[k] = T[3 + k] - T[k]; e1[k] = T[6 + k] - T[k]; e2[k] = T[6 + k] - T[3 + k]; }
    const double n[3] = {e0[1] * e1[2] - e0[2] * e1[1], e0[2] * e1[0] - e0[0] * e1[2], e0[0] * e1[1] - e0[1] * e1[0]};
    const double l0 = e0[0] * e0[0] + e0[1] * e0[1] + e0[2] * e0[2], l1 = e1[0] * e1[0] + e1[1] * e1[1] + e1[2] * e1[2], l2 = e2[0] * e2[0] + e2[1] * e2[1] + e2[2] * e2[2];
    const double L = std::max(l0, std::max(l1, l2));
    if (n[0] * n[0] + n[1] * n[1] + n[2] * n[2] > 1e-24 * L * L) continue;
    double p[3], q[3];
    if (L == l0) { memcpy(p, T, 24); memcpy(q, T + 3, 24); } else if (L == l1) { memcpy(p, T, 24); memcpy(q, T + 6, 24); } else { memcpy(p, T + 3, 24); memcpy(q, T + 6, 24); }
    memcpy(T, p, 24); memcpy(T + 3, q, 24); memcpy(T + 6, q, 24);
  }
  e->geoms.push_back(std::move(g));
  return (int)e->geoms.size() - 1;
}

int kb_add_pointcloud(kb_engine* e, const double* pts, int n, const double* radius, double margin) {
  if (!e || e->finalized) return fail(KB_ERR_STATE, "engine is null or already finalized");
  if (n < 0 || !(margin >= 0) || !std::isfinite(margin)) return fail(KB_ERR_INVALID, "negative size or margin");
  if (n > 0 && !pts) return fail(KB_ERR_INVALID, "null point array");
  if (n > 0 && (!all_finite(pts, 3 * (size_t)n) || (radius && !all_finite(radius, (size_t)n)))) return fail(KB_ERR_INVALID, "point cloud has a non-finite coordinate or radius");
  if (radius) for (int i = 0; i < n; i++) if (radius[i] < 0) return fail(KB_ERR_INVALID, "point %d has a negative radius", i);
  Geom g; g.kind = n > 0 ? G_CLOUD : G_EMPTY; g.margin = margin; g.sph.resize(4 * (size_t)n);
  for (int i = 0; i < n; i++) { for (int k = 0; k < 3; k++) g.sph[4 * (size_t)i + k] = pts[3 * (size_t)i + k]; g.sph[4 * (size_t)i + 3] = radius ? radius[i] : 0.0; }
  e->geoms.push_back(std::move(g));
  return (int)e->geoms.size() - 1;
}

int kb_add_primitive(kb_engine* e, int type, const double* params, double margin) {
  if (!e || e->finalized) return fail(KB_ERR_STATE, "engine is null or already finalized");
  if (!(margin >= 0) || !std::isfinite(margin) || !params) return fail(KB_ERR_INVALID, "negative margin or null parameters");
  {
    static const int nparams[6] = {3, 4, 9, 15, 6, 6};      // point, sphere, triangle, box, aabb, segment
    if (type >= 0 && type < 6 && !all_finite(params, nparams[type])) return fail(KB_ERR_INVALID, "primitive has a non-finite parameter");
    if (type == KB_PRIM_SPHERE && params[3] < 0) return fail(KB_ERR_INVALID, "sphere has a negative radius");
  }
  if (type == KB_PRIM_TRIANGLE) {       // a triangle primitive is a one-triangle mesh for every query of this path
    const int32_t idx[3] = {0, 1, 2};
    return kb_add_trimesh(e, params, 3, idx, 1, margin);
  }
  if (type == KB_PRIM_SEGMENT) {        // a segment is the zero-area triangle (a, b, b): the predicates give it segment semantics
    if (params[0] == params[3] && params[1] == params[4] && params[2] == params[5]) return fail(KB_ERR_INVALID, "segment of zero length: use a point");
    const int32_t idx[3] = {0, 1, 1};
    return kb_add_trimesh(e, params, 2, idx, 1, margin);
  }
  if (type == KB_PRIM_BOX || type == KB_PRIM_AABB) {
    // solid box: its surface as 12 triangles (vertex = R l + c, the same arithmetic and the same triangle list as the oracle's) plus
    // the solid descriptor.  Vertex index = (x > 0) + 2 (y > 0) + 4 (z > 0).
    double c[3], R[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, h[3];
    if (type == KB_PRIM_AABB) { for (int k = 0; k < 3; k++) { c[k] = 0.5 * (params[k] + params[3 + k]); h[k] = 0.5 * (params[3 + k] - params[k]); } }
    else { memcpy(c, params, 24); memcpy(R, params + 3, 72); memcpy(h, params + 12, 24); }
    for (int k = 0; k < 3; k++) if (!(h[k] >= 0)) return fail(KB_ERR_INVALID, "box half dimensions must be >= 0");
    const int nzero = (h[0] == 0) + (h[1] == 0) + (h[2] == 0);
    if (nzero >= 2) return fail(KB_ERR_UNSUPPORTED, "a box with two zero dimensions is a segment or a point: use those primitives");
    double v[24]; int nv = 0;
    for (int sz = -1; sz <= 1; sz += 2) for (int sy = -1; sy <= 1; sy += 2) for (int sx = -1; sx <= 1; sx += 2) {
      const double l[3] = {sx * h[0], sy * h[1], sz * h[2]};
      for (int k = 0; k < 3; k++) v[3 * nv + k] = R[3 * k] * l[0] + R[3 * k + 1] * l[1] + R[3 * k + 2] * l[2] + c[k];
      nv++;
    }
    static const int T[36] = {0, 2, 3, 0, 3, 1, 4, 5, 7, 4, 7, 6, 0, 1, 5, 0, 5, 4, 2, 6, 7, 2, 7, 3, 0, 4, 6, 0, 6, 2, 1, 3, 7, 1, 7, 5};
    // face order of T: z-, z+, y-, y+, x-, x+ (two triangles each).  A flat box keeps only the two faces across its zero dimension:
    // the other four have no area, and zero-area triangles have no place in the exact predicates.
    Geom g; g.kind = G_MESH; g.margin = margin; g.solid = true;
    for (int f = 0; f < 6; f++) {
      const int axis = 2 - f / 2;
      if (nzero == 1 && h[axis] != 0) continue;
      for (int t = 6 * f; t < 6 * f + 6; t++) g.tri.insert(g.tri.end(), v + 3 * T[t], v + 3 * T[t] + 3);
    }
    memcpy(g.box, c, 24); memcpy(g.box + 3, R, 72); memcpy(g.box + 12, h, 24);
    e->geoms.push_back(std::move(g));
    return (int)e->geoms.size() - 1;
  }
  if (type != KB_PRIM_POINT && type != KB_PRIM_SPHERE) return fail(KB_ERR_UNSUPPORTED, "primitive type %d is not supported (point, sphere, segment, triangle, box and aabb are)", type);
  Geom g; g.kind = G_PRIM; g.margin = margin; g.sph = {params[0], params[1], params[2], type == KB_PRIM_SPHERE ? params[3] : 0.0};
  e->geoms.push_back(std::move(g));
  return (int)e->geoms.size() - 1;
}

int kb_add_dynamic_pointcloud(kb_engine* e, int capacity, double radius, double margin) {
  if (!e || e->finalized) return fail(KB_ERR_STATE, "engine is null or already finalized");
#ifdef KB_QNODES
  return fail(KB_ERR_UNSUPPORTED, "the experimental quantised-node build has no GPU hierarchy builder");
#endif
  if (capacity < 1 || capacity > (1 << 27) || radius < 0 || margin < 0) return fail(KB_ERR_INVALID, "capacity must be in [1, 2^27], radius and margin >= 0");
  Geom g; g.kind = G_CLOUD; g.margin = margin; g.dyn_cap = capacity; g.dyn_radius = radius;
  e->geoms.push_back(std::move(g));
  return (int)e->geoms.size() - 1;
}

int kb_update_pointcloud(kb_engine* e, int geom, const double* pts, int n) {
  if (!e || !e->finalized) return fail(KB_ERR_STATE, "engine is not finalized");
  for (kb_engine* r : e->replicas) { int rc = kb_update_pointcloud(r, geom, pts, n); if (rc) return rc; }      // every device rebuilds its own copy
  const kb_engine::DynCloud* dc = nullptr;
  for (const auto& d : e->dyn) if (d.geom == geom) dc = &d;
  if (!dc) return fail(KB_ERR_INVALID, "geometry %d is not a dynamic point cloud attached to a terrain / rigid object with an enabled link pair", geom);
  if (n < 0 || n > dc->cap || (n > 0 && !pts)) return fail(KB_ERR_INVALID, "n = %d outside [0, capacity %d]", n, dc->cap);
  CK(cudaSetDevice(e->device));
  int maxcap = 0; for (const auto& d : e->dyn) maxcap = std::max(maxcap, d.cap);
  if (!e->d_dyn_scratch) {
    e->dyn_scratch_bytes = kb_lbvh_scratch_bytes(maxcap);
    CK(cudaMalloc(&e->d_dyn_scratch, e->dyn_scratch_bytes)); CK(cudaMalloc((void**)&e->d_dyn_pts, (size_t)maxcap * 24)); CK(cudaMalloc((void**)&e->d_dyn_T, 96));
  }
  const DevGeom& G = e->groups[dc->group];
  CK(cudaMemcpyAsync(e->d_dyn_T, dc->T, 96, cudaMemcpyHostToDevice, e->stream));
  if (n > 0) CK(cudaMemcpyAsync(e->d_dyn_pts, pts, (size_t)n * 24, cudaMemcpyHostToDevice, e->stream));
  float maxabs = 0.f;
  CK(kb_lbvh_build(e->d_dyn_pts, nullptr, dc->radius, n, e->d_dyn_T, dc->owner, nullptr, e->d_sph64 + 4 * (size_t)G.elem_base, e->d_sph32 + G.elem_base,
                   e->d_sphown + G.elem_base, e->d_nodes + 2 * (size_t)G.node_base, e->d_dyn_scratch, e->dyn_scratch_bytes, maxcap, &maxabs, e->stream, nullptr, e->d_sphorig + G.elem_base));
  e->stats.kernel_launches += 7;
  // the fp32 error bound follows the scene extent: a cloud that reaches further out widens it (never narrows)
  e->eps_extent = std::max(e->eps_extent, (double)maxabs + dc->radius);
  double S = std::max(e->eps_extent, e->eps_reach + e->eps_lmax) + e->eps_lmax; if (!(S > 0)) S = 1;
  e->scene.eps_abs = std::max(e->scene.eps_abs, (float)(8.0 * 5.9604645e-8 * S));
  return KB_OK;
}

int kb_add_terrain(kb_engine* e, int geom) {
  if (!e || e->finalized) return fail(KB_ERR_STATE, "engine is null or already finalized");
  if (geom < 0 || geom >= (int)e->geoms.size()) return fail(KB_ERR_INVALID, "unknown geometry %d", geom);
  e->terrains.push_back(geom); return (int)e->terrains.size() - 1;
}

int kb_add_rigid_object(kb_engine* e, int geom, const double T[12]) {
  if (!e || e->finalized) return fail(KB_ERR_STATE, "engine is null or already finalized");
  if (geom < 0 || geom >= (int)e->geoms.size()) return fail(KB_ERR_INVALID, "unknown geometry %d", geom);
  if (!T || !all_finite(T, 12)) return fail(KB_ERR_INVALID, "rigid object needs a finite transform");
  Xf x; xf_from12(T, x); e->objects.push_back(geom); e->objT.push_back(x); return (int)e->objects.size() - 1;
}

int kb_robot_create(kb_engine* e, int L, const int32_t* parents, const uint8_t* linktype, const double* axis, const double* T0,
                    const double* qmin, const double* qmax) {
  if (!e || e->finalized) return fail(KB_ERR_STATE, "engine is null or already finalized");
  if (e->L) return fail(KB_ERR_STATE, "the engine already has its active robot");
  if (L <= 0 || L > KB_MAX_LINKS) return fail(KB_ERR_UNSUPPORTED, "robot has %d links; supported: 1..%d", L, KB_MAX_LINKS);
  if (!parents || !linktype || !axis || !T0 || !qmin || !qmax) return fail(KB_ERR_INVALID, "null robot array");
  if (!all_finite(axis, 3 * (size_t)L) || !all_finite(T0, 12 * (size_t)L)) return fail(KB_ERR_INVALID, "robot has a non-finite axis or parent transform");
  for (int i = 0; i < L; i++) {
    if (linktype[i] != KB_REVOLUTE && linktype[i] != KB_PRISMATIC) return fail(KB_ERR_INVALID, "link %d has type %d (0 = revolute, 1 = prismatic)", i, (int)linktype[i]);
    if (std::isnan(qmin[i]) || std::isnan(qmax[i])) return fail(KB_ERR_INVALID, "link %d has a NaN joint limit (use +-inf for no limit)", i);
  }
  for (int i = 0; i < L; i++) if (parents[i] >= i || parents[i] < -1) return fail(KB_ERR_INVALID, "parents[%d]=%d must be -1 or < %d", i, parents[i], i);
  e->L = L;
  e->parents.assign(parents, parents + L); e->linktype.assign(linktype, linktype + L);
  e->axis.assign(axis, axis + 3 * L); e->T0.assign(T0, T0 + 12 * L); e->qmin.assign(qmin, qmin + L); e->qmax.assign(qmax, qmax + L);
  e->linkgeom.assign(L, -1);
  e->jtype.assign(L, KB_JOINT_NORMAL); e->jlink.resize(L); std::iota(e->jlink.begin(), e->jlink.end(), 0);
  e->selfcol.assign((size_t)L * L, 0);
  return KB_OK;
}

int kb_robot_set_link_geometry(kb_engine* e, int link, int geom) {
  if (!e || e->finalized) return fail(KB_ERR_STATE, "engine is null or already finalized");
  if (link < 0 || link >= e->L || geom >= (int)e->geoms.size()) return fail(KB_ERR_INVALID, "bad link %d or geometry %d", link, geom);
  if (geom >= 0 && e->geoms[geom].dyn_cap > 0) return fail(KB_ERR_UNSUPPORTED, "a dynamic point cloud can only be a terrain or a rigid object");
  e->linkgeom[link] = geom; return KB_OK;
}

// links a joint drives, root to tip: the chain from its base (exclusive) to its link (inclusive) -- RobotModel::GetJointIndices,
// reference Cpp/Modeling/Robot.cpp:2120-2144
static int joint_chain(const kb_engine* e, int link, int base, int idx[6]) {
  int n = 0, tmp[8];
  while (link != base) { if (link < 0 || n >= 6) return -1; tmp[n++] = link; link = e->parents[link]; }
  for (int i = 0; i < n; i++) idx[i] = tmp[n - 1 - i];
  return n;
}

int kb_robot_set_joints(kb_engine* e, int nj, const uint8_t* jtype, const int32_t* jlink, const int32_t* jbase) {
  if (!e || e->finalized) return fail(KB_ERR_STATE, "engine is null or already finalized");
  if (nj < 0 || nj > KB_MAX_LINKS) return fail(KB_ERR_INVALID, "bad joint count %d", nj);
  std::vector<int16_t> jidx((size_t)nj * 6, 0);
  for (int i = 0; i < nj; i++) {
    if (jlink[i] < 0 || jlink[i] >= e->L) return fail(KB_ERR_INVALID, "joint %d references link %d", i, jlink[i]);
    const int t = jtype[i];
    if (t == KB_JOINT_CLOSED || t > KB_JOINT_CLOSED) return fail(KB_ERR_UNSUPPORTED, "joint %d: closed-chain joints are not supported", i);
    if (t == KB_JOINT_FLOATING || t == KB_JOINT_FLOATINGPLANAR || t == KB_JOINT_BALLANDSOCKET) {
      // the link layout the reference asserts (Cpp/Modeling/Interpolate.cpp:24-26,231-236): translations, then rotations about z, y, x
      int idx[6]; const int base = jbase ? jbase[i] : -2;
      if (base < -1) return fail(KB_ERR_INVALID, "joint %d spans several links and needs its base link (jbase)", i);
      const int n = joint_chain(e, jlink[i], base, idx);
      const int want = t == KB_JOINT_FLOATING ? 6 : 3;
      if (n != want) return fail(KB_ERR_INVALID, "joint %d drives %d links from base %d; its type needs %d", i, n, base, want);
      auto rev = [&](int l) { return e->linktype[l] == KB_REVOLUTE; };
      auto ax = [&](int l, int k) { return e->axis[3 * l + k] == 1.0; };
      bool ok = true;
      if (t == KB_JOINT_FLOATING) ok = !rev(idx[0]) && !rev(idx[1]) && !rev(idx[2]) && rev(idx[3]) && rev(idx[4]) && rev(idx[5]) && ax(idx[3], 2) && ax(idx[4], 1) && ax(idx[5], 0);
      else if (t == KB_JOINT_BALLANDSOCKET) ok = rev(idx[0]) && rev(idx[1]) && rev(idx[2]) && ax(idx[0], 2) && ax(idx[1], 1) && ax(idx[2], 0);
      else ok = rev(idx[2]);
      if (!ok) return fail(KB_ERR_INVALID, "joint %d: link types / axes do not match the layout its type requires (translations, then rotations about z, y, x)", i);
      for (int k = 0; k < n; k++) jidx[(size_t)i * 6 + k] = (int16_t)idx[k];
    }
  }
  e->jtype.assign(jtype, jtype + nj); e->jlink.assign(jlink, jlink + nj); e->jidx = jidx; return KB_OK;
}

int kb_robot_add_driver(kb_engine* e, int n, const int32_t* links, const double* scale, const double* offset, double dmin, double dmax) {
  if (!e || e->finalized) return fail(KB_ERR_STATE, "engine is null or already finalized");
  if (n <= 0) return fail(KB_ERR_INVALID, "driver needs at least one link");
  Driver d; d.dmin = dmin; d.dmax = dmax;
  for (int i = 0; i < n; i++) {
    if (links[i] < 0 || links[i] >= e->L) return fail(KB_ERR_INVALID, "driver references link %d", links[i]);
    d.links.push_back(links[i]); d.scale.push_back(scale ? scale[i] : 1.0); d.offset.push_back(offset ? offset[i] : 0.0);
  }
  e->drivers.push_back(d); return (int)e->drivers.size() - 1;
}

int kb_robot_set_self_collision(kb_engine* e, int i, int j, int enabled) {
  if (!e || e->finalized) return fail(KB_ERR_STATE, "engine is null or already finalized");
  if (i > j) std::swap(i, j);
  if (i == j || i < 0 || j >= e->L) return fail(KB_ERR_INVALID, "bad self-collision pair %d,%d", i, j);
  if (e->selfcol_default) { init_all_self_collisions(e); e->selfcol_default = false; }
  if (enabled && (geom_empty(e, e->linkgeom[i]) || geom_empty(e, e->linkgeom[j]))) enabled = 0;
  e->selfcol[i * e->L + j] = enabled != 0; return KB_OK;
}

int kb_set_pair_mask(kb_engine* e, const uint8_t* mask, int n_ids) {
  if (!e || e->finalized) return fail(KB_ERR_STATE, "engine is null or already finalized");
  if (n_ids != num_ids(e)) return fail(KB_ERR_INVALID, "mask is %d x %d but the world has %d ids", n_ids, n_ids, num_ids(e));
  e->mask.assign(mask, mask + (size_t)n_ids * n_ids); e->nids = n_ids; e->mask_user = true; return KB_OK;
}

int kb_num_ids(const kb_engine* e) { return e ? num_ids(e) : 0; }

int kb_get_pair_mask(const kb_engine* e, uint8_t* out) {
  if (!e || !e->finalized) return fail(KB_ERR_STATE, "engine is not finalized");
  memcpy(out, e->mask.data(), e->mask.size()); return e->nids;
}

int kb_finalize(kb_engine* e, int device) {
  if (!e || e->finalized) return fail(KB_ERR_STATE, "engine is null or already finalized");
  if (!e->L) return fail(KB_ERR_STATE, "no robot: call kb_robot_create first");
  if (e->selfcol_default) { init_all_self_collisions(e); e->selfcol_default = false; }
  if (!e->mask_user) init_default_mask(e);
  int ndev = 0;
  cudaError_t ce = cudaGetDeviceCount(&ndev);
  if (ce != cudaSuccess || ndev == 0) return fail(KB_ERR_CUDA, "no CUDA device available (%s); this engine has no CPU fallback", cudaGetErrorString(ce));
  if (device < 0 || device >= ndev) return fail(KB_ERR_INVALID, "device %d out of range (0..%d)", device, ndev - 1);
  CK(cudaSetDevice(device));
  e->device = device;
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, device));
  e->num_sms = prop.multiProcessorCount;
  CK(cudaStreamCreateWithFlags(&e->own_stream, cudaStreamNonBlocking));
  e->stream = e->own_stream;
  CK(cudaEventCreate(&e->ev0)); CK(cudaEventCreate(&e->ev1));
  CK(cudaStreamCreateWithFlags(&e->copy_stream, cudaStreamNonBlocking)); CK(cudaStreamCreateWithFlags(&e->aux_stream, cudaStreamNonBlocking));
  for (int k = 0; k < 4; k++) CK(cudaEventCreateWithFlags(&e->ev_copy[k], cudaEventDisableTiming));
  for (int k = 0; k < KB_STAGES; k++) CK(cudaEventCreateWithFlags(&e->ev_stage[k], cudaEventDisableTiming));

  // ---- 1. per-geometry local-frame BVHs (links, and every geometry for explicit pair queries)
  e->dgeoms.resize(e->geoms.size());
  std::vector<int32_t> none;
  for (size_t g = 0; g < e->geoms.size(); g++) {
    const Geom& G = e->geoms[g];
    if (G.dyn_cap > 0) { e->dgeoms[g] = DevGeom(); continue; }   // replaceable clouds live in their environment group only
    if (e->cloud_builder == 1 && G.kind == G_CLOUD && G.nelem() > KB_GPU_CLOUD_MIN) { reserve_gpu_cloud(e, G.sph, none, G.margin, e->dgeoms[g]); continue; }
    if (e->mesh_builder == 1 && G.kind == G_MESH && G.nelem() > KB_GPU_MESH_MIN) { reserve_gpu_mesh(e, G.tri, none, G.margin, e->dgeoms[g]); continue; }
    int rc = append_geom(e, G.kind == G_MESH ? G_MESH : G_CLOUD, G.kind == G_MESH ? G.tri : G.sph, none, G.margin, e->dgeoms[g], true);
    if (rc) return rc;
    if (G.kind == G_EMPTY) e->dgeoms[g].empty = true;
  }
  e->dsolid.assign(e->geoms.size(), DevGeom());
  for (size_t g = 0; g < e->geoms.size(); g++) if (e->geoms[g].solid) {
    std::vector<double> b(e->geoms[g].box, e->geoms[g].box + 15);
    int rc = append_geom(e, G_BOX, b, none, e->geoms[g].margin, e->dsolid[g], false);
    if (rc) return rc;
  }
  // ---- 2. merged world-frame environment groups: static objects with the same (element kind, margin, link mask)
  const int L = e->L, T = (int)e->terrains.size(), O = (int)e->objects.size();
  struct Grp { int kind; double margin; std::string sig; std::vector<double> elems; std::vector<int32_t> owners; int dyn_geom = -1; int dyn_owner = -1; std::vector<int32_t> origs; };
  std::vector<Grp> grp;
  std::map<std::string, int> grp_index;
  std::vector<int> static_grp((size_t)(T + O), -1);     // merged group that holds a static body's elements in the world frame (ray casting reuses it)
  double extent = 0;
  for (int s = 0; s < T + O; s++) {
    int gi = s < T ? e->terrains[s] : e->objects[s - T];
    if (geom_empty(e, gi)) continue;
    const Geom& G = e->geoms[gi];
    std::string sig((size_t)L, '0'); bool any = false;
    for (int j = 0; j < L; j++) if (!geom_empty(e, e->linkgeom[j]) && (mask_en(e, link_id(e, j), s) || mask_en(e, s, link_id(e, j)))) { sig[j] = '1'; any = true; }
    if (!any) continue;
    int kind = G.kind == G_MESH ? G_MESH : G_CLOUD;
    if (G.dyn_cap > 0) {        // a replaceable cloud is a group of its own: its storage is reserved, its hierarchy built on the GPU
      for (const Grp& other : grp) if (other.dyn_geom == gi) return fail(KB_ERR_UNSUPPORTED, "a dynamic point cloud can be attached to one terrain / rigid object only");
      Grp dg; dg.kind = G_CLOUD; dg.margin = G.margin; dg.sig = sig; dg.dyn_geom = gi; dg.dyn_owner = s;
      grp.push_back(dg);
      continue;
    }
    char key[64]; snprintf(key, sizeof key, "%d:%.17g:", kind, G.margin);
    std::string k = std::string(key) + sig;
    auto it = grp_index.find(k);
    if (it == grp_index.end()) { grp_index[k] = (int)grp.size(); grp.push_back({kind, G.margin, sig, {}, {}}); it = grp_index.find(k); }
    Xf X; if (s < T) { memset(&X, 0, sizeof X); X.R[0] = X.R[4] = X.R[8] = 1; } else X = e->objT[s - T];
    if (G.solid) {      // the solid of a static box primitive, baked into the world frame: centre X(c), axes X.R R
      char bkey[64]; snprintf(bkey, sizeof bkey, "%d:%.17g:", (int)G_BOX, G.margin);
      std::string bk = std::string(bkey) + sig;
      auto bit = grp_index.find(bk);
      if (bit == grp_index.end()) { grp_index[bk] = (int)grp.size(); grp.push_back({G_BOX, G.margin, sig, {}, {}}); bit = grp_index.find(bk); }
      double w[15];
      xf_apply(X, G.box, w);
      for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) w[3 + 3 * i + j] = X.R[3 * i] * G.box[3 + j] + X.R[3 * i + 1] * G.box[6 + j] + X.R[3 * i + 2] * G.box[9 + j];
      memcpy(w + 12, G.box + 12, 24);
      grp[bit->second].elems.insert(grp[bit->second].elems.end(), w, w + 15); grp[bit->second].owners.push_back(s);
    }
    Grp& gr = grp[it->second];       // taken after the push_back above: a reference into `grp` would not survive it
    static_grp[s] = it->second;
    if (kind == G_MESH) {
      int nt = (int)(G.tri.size() / 9);
      for (int t = 0; t < nt; t++) {
        double w[9];
        for (int v = 0; v < 3; v++) { if (s < T) memcpy(w + 3 * v, &G.tri[9 * (size_t)t + 3 * v], 24); else xf_apply(X, &G.tri[9 * (size_t)t + 3 * v], w + 3 * v); }
        gr.elems.insert(gr.elems.end(), w, w + 9); gr.owners.push_back(s); gr.origs.push_back(t);
        for (int k2 = 0; k2 < 9; k2++) extent = std::max(extent, std::fabs(w[k2]));
      }
    } else {
      int np = (int)(G.sph.size() / 4);
      for (int i = 0; i < np; i++) {
        double w[4];
        if (s < T) memcpy(w, &G.sph[4 * (size_t)i], 24); else xf_apply(X, &G.sph[4 * (size_t)i], w);
        w[3] = G.sph[4 * (size_t)i + 3];
        gr.elems.insert(gr.elems.end(), w, w + 4); gr.owners.push_back(s); gr.origs.push_back(i);
        for (int k2 = 0; k2 < 3; k2++) extent = std::max(extent, std::fabs(w[k2]) + w[3]);
      }
    }
  }
  e->groups.resize(grp.size());
  e->hgrids.assign(std::min<size_t>(grp.size(), KB_MAX_GRIDS), HostGrid());
  e->dyn.clear();
  for (size_t g = 0; g < grp.size(); g++) {
    if (grp[g].dyn_geom >= 0) {
      const Geom& G = e->geoms[grp[g].dyn_geom];
      DevGeom& dg = e->groups[g];
      dg = DevGeom(); dg.margin = G.margin; dg.kind = KB_ELEM_SPHERE; dg.nelem = G.dyn_cap; dg.empty = false; dg.rmax = G.dyn_radius;
      dg.depth = 64;                                              // bound for the stack model: 30 key bits + index bits of a linear BVH
      for (int k = 0; k < 3; k++) dg.lo[k] = dg.hi[k] = 0;
      if ((e->h_nodes.size() / 8) & 1) e->h_nodes.insert(e->h_nodes.end(), 8, 0.f);
      dg.node_base = (int)(e->h_nodes.size() / 8); dg.nnodes = (int)kb_lbvh_nodes_for(G.dyn_cap);
      {   // until the first update: a root leaf without elements whose box overlaps nothing
        float v[8] = {0.f, 0.f, 0.f, i2f(~0), -1e30f, -1e30f, -1e30f, i2f(0)};
        e->h_nodes.insert(e->h_nodes.end(), v, v + 8);
        e->h_nodes.insert(e->h_nodes.end(), (size_t)(dg.nnodes - 1) * 8, 0.f);
      }
      dg.elem_base = (int)(e->h_sph64.size() / 4);
      e->h_sph64.insert(e->h_sph64.end(), (size_t)G.dyn_cap * 4, 0.0); e->h_sph32.insert(e->h_sph32.end(), (size_t)G.dyn_cap * 4, 0.f);
      e->h_sphown.insert(e->h_sphown.end(), (size_t)G.dyn_cap, grp[g].dyn_owner); e->h_sphorig.insert(e->h_sphorig.end(), (size_t)G.dyn_cap, -1);
      kb_engine::DynCloud dc; dc.geom = grp[g].dyn_geom; dc.group = (int)g; dc.owner = grp[g].dyn_owner; dc.cap = G.dyn_cap; dc.radius = G.dyn_radius;
      const int so = grp[g].dyn_owner;
      Xf X; if (so < T) { memset(&X, 0, sizeof X); X.R[0] = X.R[4] = X.R[8] = 1; } else X = e->objT[so - T];
      memcpy(dc.T, X.R, 72); memcpy(dc.T + 9, X.t, 24);
      e->dyn.push_back(dc);
      continue;
    }
    if (e->cloud_builder == 1 && grp[g].kind == G_CLOUD && (int)(grp[g].elems.size() / 4) > KB_GPU_CLOUD_MIN) {
      reserve_gpu_cloud(e, grp[g].elems, grp[g].owners, grp[g].margin, e->groups[g], grp[g].origs);
    } else if (e->mesh_builder == 1 && grp[g].kind == G_MESH && (int)(grp[g].elems.size() / 9) > KB_GPU_MESH_MIN) {
      reserve_gpu_mesh(e, grp[g].elems, grp[g].owners, grp[g].margin, e->groups[g], grp[g].origs);
    } else {
      int rc = append_geom(e, grp[g].kind, grp[g].elems, grp[g].owners, grp[g].margin, e->groups[g], false, grp[g].kind == G_BOX ? std::vector<int32_t>() : grp[g].origs);
      if (rc) return rc;
    }
    if (g < KB_MAX_GRIDS && e->grid_res >= 8 && !e->groups[g].empty && grp[g].kind != G_BOX && grp[g].dyn_geom < 0) {
      // pad so that every covering sphere of the links that meet this group can be cleared outside the group's bounds
      double need = 0;
      for (int j = 0; j < L; j++) if (grp[g].sig[j] == '1') {
        const DevGeom& lg = e->dgeoms[e->linkgeom[j]];
        for (int c = 0; c < lg.ncover; c++) need = std::max(need, lg.cover[c][3] + lg.margin + grp[g].margin);
      }
      double ext = 0; for (int k = 0; k < 3; k++) ext = std::max(ext, e->groups[g].hi[k] - e->groups[g].lo[k]);
      const double pad = 1.25 * need + 6.0 * (ext + 2.5 * need) / e->grid_res;
      build_clear_grid(grp[g].kind, grp[g].elems, e->groups[g].lo, e->groups[g].hi, pad, e->grid_res, e->hgrids[g]);
    }
    std::vector<double>().swap(grp[g].elems);
  }
  // ---- 3. work items per configuration: links vs groups (environment first, as CheckCollisionFree does), then self pairs
  e->feas_items = ItemSet(); e->env_items = ItemSet();
  e->feas_items.nxf = e->env_items.nxf = L;
  struct ItemSrc { int item, link, group; };
  std::vector<ItemSrc> item_src;
  for (size_t g = 0; g < grp.size(); g++)
    for (int j = 0; j < L; j++) if (grp[g].sig[j] == '1') {
      const DevGeom& lg = e->dgeoms[e->linkgeom[j]];
      const size_t before = e->feas_items.items.size();
      int rc = add_item(e->feas_items, lg, j, link_id(e, j), e->groups[g], -1, -1, false); if (rc) return rc;
      if (e->feas_items.items.size() > before) item_src.push_back({(int)before, j, (int)g});
      rc = add_item(e->env_items, lg, j, link_id(e, j), e->groups[g], -1, -1, false); if (rc) return rc;
      // a link that is a box primitive is solid: its interior against the group's elements (not against other solids --
      // their surfaces are in the mesh groups)
      const DevGeom& ls = e->dsolid[e->linkgeom[j]];
      if (!ls.empty && grp[g].kind != G_BOX) {
        rc = add_item(e->feas_items, ls, j, link_id(e, j), e->groups[g], -1, -1, false); if (rc) return rc;
        rc = add_item(e->env_items, ls, j, link_id(e, j), e->groups[g], -1, -1, false); if (rc) return rc;
      }
    }
  for (int i = 0; i < L; i++) for (int j = i + 1; j < L; j++) {
    if (geom_empty(e, e->linkgeom[i]) || geom_empty(e, e->linkgeom[j])) continue;
    // WorldPlannerSettings::CheckCollision(world, ids): enabled(i,j) || enabled(i,i); the second term is the diagonal (always false
    // for links under InitializeDefault, but honoured if a caller's mask sets it)
    if (!(mask_en(e, link_id(e, i), link_id(e, j)) || mask_en(e, link_id(e, i), link_id(e, i)))) continue;
    int rc = add_item(e->feas_items, e->dgeoms[e->linkgeom[i]], i, link_id(e, i), e->dgeoms[e->linkgeom[j]], j, link_id(e, j), true); if (rc) return rc;
    if (!e->dsolid[e->linkgeom[i]].empty) { rc = add_item(e->feas_items, e->dsolid[e->linkgeom[i]], i, link_id(e, i), e->dgeoms[e->linkgeom[j]], j, link_id(e, j), true); if (rc) return rc; }
    if (!e->dsolid[e->linkgeom[j]].empty) { rc = add_item(e->feas_items, e->dgeoms[e->linkgeom[i]], i, link_id(e, i), e->dsolid[e->linkgeom[j]], j, link_id(e, j), true); if (rc) return rc; }
  }
  if (KB_STACK_CAP - (3 * e->feas_items.maxdepth + 1) / 2 - 4 < 160) return fail(KB_ERR_UNSUPPORTED, "BVH depth sum %d exceeds the traversal stack model", e->feas_items.maxdepth);
  // ---- 4. fp32 coordinate error bound: scene extent + robot reach
  double reach = 0;
  for (int j = 0; j < L; j++) {
    double t = std::sqrt(e->T0[12 * j + 9] * e->T0[12 * j + 9] + e->T0[12 * j + 10] * e->T0[12 * j + 10] + e->T0[12 * j + 11] * e->T0[12 * j + 11]);
    if (e->linktype[j] == KB_PRISMATIC) t += std::max(std::fabs(e->qmin[j]), std::fabs(e->qmax[j]));
    reach += t;
  }
  double lmax = 0;
  for (const DevGeom& dg : e->dgeoms) if (!dg.empty) for (int k = 0; k < 3; k++) lmax = std::max(lmax, std::max(std::fabs(dg.lo[k]), std::fabs(dg.hi[k])));
  double S = std::max(extent, reach + lmax) + lmax;
  if (!(S > 0)) S = 1;
  e->scene.eps_abs = (float)(8.0 * 5.9604645e-8 * S);
  e->eps_extent = extent; e->eps_reach = reach; e->eps_lmax = lmax;
  // ---- 4b. clearance probes of the (link, static group) items
  {
    ItemSet& fs = e->feas_items;
    fs.probes.clear(); fs.always_on.assign((fs.items.size() + 31) / 32 + 1, 0u);
    std::vector<char> probed(fs.items.size(), 0);
    for (const ItemSrc& is : item_src) {
      if (is.group >= (int)e->hgrids.size() || !(e->hgrids[is.group].h > 0)) continue;
      const HostGrid& G = e->hgrids[is.group];
      const DevGeom& lg = e->dgeoms[e->linkgeom[is.link]];
      if (lg.ncover <= 0) continue;
      probed[is.item] = 1;
      for (int c = 0; c < lg.ncover; c++) {
        KbProbe pr; memset(&pr, 0, sizeof pr);
        for (int k = 0; k < 3; k++) pr.c[k] = (float)lg.cover[c][k];
        // radius + threshold + fp32 slack (sphere centre rounded to fp32 and moved by fp32 transforms), in quarter voxels, rounded up
        double lmaxc = std::fabs(lg.cover[c][0]) + std::fabs(lg.cover[c][1]) + std::fabs(lg.cover[c][2]);
        const double reach_m = lg.cover[c][3] + fs.items[is.item].thr + 16.0 * (double)e->scene.eps_abs + 4e-7 * lmaxc;
        const double qv = std::ceil(4.0 * reach_m / G.h) + 1.0;
        pr.need = qv > 255.0 ? 256u : (uint32_t)qv;          // 256: can never be cleared by a u8 value
        pr.item = is.item; pr.xf = is.link; pr.grid = is.group;
        fs.probes.push_back(pr);
      }
    }
    for (size_t i = 0; i < fs.items.size(); i++) if (!probed[i]) fs.always_on[i >> 5] |= 1u << (i & 31);
  }
  // ---- 4c. ray casting: one body per link / rigid object / terrain that has a geometry, whatever the collision mask says (a ray
  // sees every body: WorldModel::RayCast, World.cpp:465-516), static ones under a top-level hierarchy of their world boxes
  {
    e->ray_bodies.clear(); e->h_tlas.clear(); e->ray_max_margin = 0; e->tlas_ext = 0;
    auto body_of = [&](const DevGeom& dg, int id, int rank, int xf, const Xf* X) {
      KbRayBody b; memset(&b, 0, sizeof b);
      b.margin = dg.margin; b.node_base = dg.node_base; b.elem_base = dg.elem_base; b.kind = dg.kind; b.id = id; b.rank = rank; b.xf = xf;
      b.T[0] = b.T[4] = b.T[8] = 1;
      if (X) { memcpy(b.T, X->R, 72); memcpy(b.T + 9, X->t, 24); b.has_T = 1; }
      double ext = 0; for (int k = 0; k < 3; k++) ext = std::max(ext, std::max(std::fabs(dg.lo[k]), std::fabs(dg.hi[k])));
      b.ext = (float)(3.0 * ext * (1 + 1e-6));
      return b;
    };
    for (int j = 0; j < L; j++) {
      if (geom_empty(e, e->linkgeom[j])) continue;
      const DevGeom& dg = e->dgeoms[e->linkgeom[j]];
      if (dg.empty || dg.depth >= 90) continue;
      e->ray_bodies.push_back(body_of(dg, link_id(e, j), j, j, nullptr));
    }
    for (const auto& dc : e->dyn) {      // replaceable clouds: world-frame groups whose boxes change with every update -> not under the hierarchy
      KbRayBody b = body_of(e->groups[dc.group], dc.owner, dc.owner < T ? L + O + dc.owner : L + (dc.owner - T), -1, nullptr);
      b.ext = -1.f;                      // extent unknown ahead of the updates: the kernel derives it from the scene's error bound, which every update widens
      e->ray_bodies.push_back(b);
    }
    e->ray_nlink = (int)e->ray_bodies.size();
    std::vector<KbRayBody> st; std::vector<double> blo, bhi;
    auto add_static = [&](const KbRayBody& b, const double* lo, const double* hi, double margin, bool mesh) {
      st.push_back(b);
      const double grow = margin + 1e-9 * (1 + std::fabs(lo[0]) + std::fabs(lo[1]) + std::fabs(lo[2]) + std::fabs(hi[0]) + std::fabs(hi[1]) + std::fabs(hi[2]));
      for (int k = 0; k < 3; k++) { blo.push_back(lo[k] - grow); e->tlas_ext = std::max(e->tlas_ext, (float)std::fabs(lo[k] - grow)); }
      for (int k = 0; k < 3; k++) { bhi.push_back(hi[k] + grow); e->tlas_ext = std::max(e->tlas_ext, (float)std::fabs(hi[k] + grow)); }
      if (mesh) e->ray_max_margin = std::max(e->ray_max_margin, margin);
    };
    // static bodies whose elements already sit in a merged world-frame group are cast through that group (one hierarchy for the
    // whole environment, no per-body transform; owner id and rank come from the element), the others one by one in their own frames
    std::vector<char> group_used(e->groups.size(), 0);
    for (int s2 = 0; s2 < T + O; s2++) {
      const int gi = s2 < T ? e->terrains[s2] : e->objects[s2 - T];
      if (geom_empty(e, gi) || e->geoms[gi].dyn_cap > 0) continue;
      const int gidx = static_grp[s2];
      if (gidx >= 0 && !e->groups[gidx].empty && e->groups[gidx].depth < 90 && e->groups[gidx].kind != KB_ELEM_BOX) {
        if (!group_used[gidx]) {
          group_used[gidx] = 1;
          const DevGeom& G = e->groups[gidx];
          KbRayBody b = body_of(G, -1, -1, -1, nullptr);
          add_static(b, G.lo, G.hi, G.margin, G.kind == KB_ELEM_TRI);
        }
        continue;
      }
      const DevGeom& dg = e->dgeoms[gi];
      if (dg.empty || dg.depth >= 90) continue;
      const Xf* X = s2 < T ? nullptr : &e->objT[s2 - T];
      double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
      for (int c = 0; c < 8; c++) {
        double pl[3] = {(c & 1) ? dg.hi[0] : dg.lo[0], (c & 2) ? dg.hi[1] : dg.lo[1], (c & 4) ? dg.hi[2] : dg.lo[2]}, pw[3];
        if (X) xf_apply(*X, pl, pw); else memcpy(pw, pl, 24);
        for (int k = 0; k < 3; k++) { lo[k] = std::min(lo[k], pw[k]); hi[k] = std::max(hi[k], pw[k]); }
      }
      add_static(body_of(dg, s2, s2 < T ? L + O + s2 : L + (s2 - T), -1, X), lo, hi, dg.margin, dg.kind == KB_ELEM_TRI);
    }
    e->tlas_ext *= 3.f;
    e->ray_nstatic = (int)st.size();
    if (!st.empty()) {
      Bvh tb; build_bvh(blo, bhi, (int)st.size(), 2, tb);
      if (tb.depth >= KB_RAY_TLAS_DEPTH) return fail(KB_ERR_UNSUPPORTED, "top-level ray hierarchy is %d deep", tb.depth);
      for (int i = 0; i < (int)st.size(); i++) e->ray_bodies.push_back(st[tb.perm[i]]);
      for (const BNode& nd : tb.nodes) {
        float v[8];
        for (int k = 0; k < 3; k++) { const float c = (float)(0.5 * (nd.lo[k] + nd.hi[k])); v[k] = c; v[4 + k] = f_up(std::max(nd.hi[k] - (double)c, (double)c - nd.lo[k])); }
        if (nd.left >= 0) { v[3] = i2f(nd.left); v[7] = i2f(0); } else { v[3] = i2f(~nd.first); v[7] = i2f(nd.count); }
        e->h_tlas.insert(e->h_tlas.end(), v, v + 8);
      }
    }
  }
  // ---- 5. upload
  e->static_bytes = 0;
  int rc;
#ifdef KB_QNODES
  {   // experimental: re-encode the fp32 nodes as 16-byte quantised nodes on one scene-wide grid (conservative: boxes only grow)
    const size_t nn = e->h_nodes.size() / 8;
    double qlo[3] = {1e300, 1e300, 1e300}, qhi[3] = {-1e300, -1e300, -1e300};
    for (size_t i = 0; i < nn; i++) for (int k = 0; k < 3; k++) {
      const double c = e->h_nodes[8 * i + k], h = e->h_nodes[8 * i + 4 + k];
      qlo[k] = std::min(qlo[k], c - h); qhi[k] = std::max(qhi[k], c + h);
    }
    for (int k = 0; k < 3; k++) { if (!(qhi[k] > qlo[k])) qhi[k] = qlo[k] + 1.0; e->scene.qo[k] = (float)qlo[k]; e->scene.qs[k] = (float)((qhi[k] - qlo[k]) / 65000.0); }
    std::vector<uint32_t> q(4 * nn);
    for (size_t i = 0; i < nn; i++) {
      uint32_t cq[3], hq[3];
      for (int k = 0; k < 3; k++) {
        const double c = e->h_nodes[8 * i + k], h = e->h_nodes[8 * i + 4 + k], o = e->scene.qo[k], st = e->scene.qs[k];
        double cc = std::floor((c - o) / st + 0.5); cc = std::min(65535.0, std::max(0.0, cc));
        const double cdeq = o + st * cc;
        double hh = std::ceil((h + std::fabs(c - cdeq)) / st) + 1.0; hh = std::min(65535.0, std::max(0.0, hh));
        cq[k] = (uint32_t)cc; hq[k] = (uint32_t)hh;
      }
      int32_t left, count; memcpy(&left, &e->h_nodes[8 * i + 3], 4); memcpy(&count, &e->h_nodes[8 * i + 7], 4);
      int32_t ref = left;
      if (left < 0) { const int32_t first = ~left; if (count < 1) count = 1; if (count > 8 || first >= (1 << 28)) return fail(KB_ERR_UNSUPPORTED, "leaf not encodable in a quantised node"); ref = -1 - (first * 8 + (count - 1)); }
      q[4 * i] = cq[0] | (cq[1] << 16); q[4 * i + 1] = cq[2] | (hq[0] << 16); q[4 * i + 2] = hq[1] | (hq[2] << 16); memcpy(&q[4 * i + 3], &ref, 4);
    }
    if ((rc = upload(e->d_nodes, q.data(), q.size() * 4, &e->static_bytes, e))) return rc;
  }
#else
  if ((rc = upload(e->d_nodes, e->h_nodes.data(), e->h_nodes.size() * 4, &e->static_bytes, e))) return rc;
#endif
  if ((rc = upload(e->d_wide, e->h_wide.data(), e->h_wide.size() * 4, &e->static_bytes, e))) return rc;
  if ((rc = upload(e->d_tris32, e->h_tris32.data(), e->h_tris32.size() * 4, &e->static_bytes, e))) return rc;
  if ((rc = upload(e->d_tris64, e->h_tris64.data(), e->h_tris64.size() * 8, &e->static_bytes, e))) return rc;
  if ((rc = upload(e->d_sph32, e->h_sph32.data(), e->h_sph32.size() * 4, &e->static_bytes, e))) return rc;
  if ((rc = upload(e->d_sph64, e->h_sph64.data(), e->h_sph64.size() * 8, &e->static_bytes, e))) return rc;
  if ((rc = upload(e->d_box32, e->h_box32.data(), e->h_box32.size() * 4, &e->static_bytes, e))) return rc;
  if ((rc = upload(e->d_box64, e->h_box64.data(), e->h_box64.size() * 8, &e->static_bytes, e))) return rc;
  if ((rc = upload(e->d_boxown, e->h_boxown.data(), e->h_boxown.size() * 4, &e->static_bytes, e))) return rc;
  if ((rc = upload(e->d_triown, e->h_triown.data(), e->h_triown.size() * 4, &e->static_bytes, e))) return rc;
  if ((rc = upload(e->d_sphown, e->h_sphown.data(), e->h_sphown.size() * 4, &e->static_bytes, e))) return rc;
  if ((rc = upload(e->d_triorig, e->h_triorig.data(), e->h_triorig.size() * 4, &e->static_bytes, e))) return rc;
  if ((rc = upload(e->d_sphorig, e->h_sphorig.data(), e->h_sphorig.size() * 4, &e->static_bytes, e))) return rc;
  if ((rc = upload(e->d_raybodies, e->ray_bodies.data(), e->ray_bodies.size() * sizeof(KbRayBody), &e->static_bytes, e))) return rc;
  if ((rc = upload(e->d_tlas, e->h_tlas.data(), e->h_tlas.size() * 4, &e->static_bytes, e))) return rc;
  e->scene.nodes = e->d_nodes; e->scene.wide = e->d_wide; e->scene.tris32 = e->d_tris32; e->scene.tris64 = e->d_tris64; e->scene.sph32 = e->d_sph32; e->scene.sph64 = e->d_sph64;
  e->scene.triown = e->d_triown; e->scene.sphown = e->d_sphown; e->scene.triorig = e->d_triorig; e->scene.sphorig = e->d_sphorig;
  e->scene.box32 = e->d_box32; e->scene.box64 = e->d_box64; e->scene.boxown = e->d_boxown;
  if ((rc = upload_itemset(e->feas_items, &e->static_bytes, e))) return rc;
  if ((rc = upload_itemset(e->env_items, &e->static_bytes, e))) return rc;
  for (size_t g = 0; g < e->hgrids.size(); g++) {
    const HostGrid& G = e->hgrids[g];
    if (!(G.h > 0)) continue;
    if ((rc = upload(e->d_grid[g], G.q.data(), G.q.size(), &e->static_bytes, e))) return rc;
    KbClearGrid& D = e->scene.grids[g];
    D.data = e->d_grid[g]; D.inv_h = (float)(1.0 / G.h);
    for (int k = 0; k < 3; k++) { D.o[k] = (float)G.o[k]; D.dims[k] = G.dims[k]; }
    std::vector<uint8_t>().swap(e->hgrids[g].q);
  }
  KbRobotDev* R = new KbRobotDev(); memset(R, 0, sizeof(KbRobotDev));
  R->L = L; R->nj = (int)e->jtype.size(); R->ndrv = (int)e->drivers.size();
  for (int i = 0; i < L; i++) { R->parents[i] = e->parents[i]; R->linktype[i] = e->linktype[i]; R->qmin[i] = e->qmin[i]; R->qmax[i] = e->qmax[i]; }
  memcpy(R->axis, e->axis.data(), sizeof(double) * 3 * L); memcpy(R->T0, e->T0.data(), sizeof(double) * 12 * L);
  for (int i = 0; i < R->nj; i++) { R->jtype[i] = e->jtype[i]; R->jlink[i] = e->jlink[i]; for (int k = 0; k < 6; k++) R->jidx[i][k] = (size_t)i * 6 + k < e->jidx.size() ? e->jidx[(size_t)i * 6 + k] : 0; }
  std::vector<KbDriverDev> dd; std::vector<int32_t> dl; std::vector<double> ds, dofs;
  for (const Driver& d : e->drivers) {
    KbDriverDev x; x.first = (int)dl.size(); x.n = (int)d.links.size(); x.dmin = d.dmin; x.dmax = d.dmax; dd.push_back(x);
    dl.insert(dl.end(), d.links.begin(), d.links.end()); ds.insert(ds.end(), d.scale.begin(), d.scale.end()); dofs.insert(dofs.end(), d.offset.begin(), d.offset.end());
  }
  R->ndrv_terms = (int)dl.size();
  rc = upload(e->d_robot, R, sizeof(KbRobotDev), &e->static_bytes, e); delete R; if (rc) return rc;
  if ((rc = upload(e->d_drv, dd.data(), dd.size() * sizeof(KbDriverDev), &e->static_bytes, e))) return rc;
  if ((rc = upload(e->d_drv_link, dl.data(), dl.size() * 4, &e->static_bytes, e))) return rc;
  if ((rc = upload(e->d_drv_scale, ds.data(), ds.size() * 8, &e->static_bytes, e))) return rc;
  if ((rc = upload(e->d_drv_off, dofs.data(), dofs.size() * 8, &e->static_bytes, e))) return rc;
  if (!e->pending_clouds.empty()) {      // hierarchies of the large point clouds on the GPU (Morton order + Karras, kb_lbvh.cu)
    size_t maxn = 0; for (const auto& pc : e->pending_clouds) maxn = std::max(maxn, pc.elems.size() / (pc.mesh ? 9 : 4));
    double* d_p = nullptr; double* d_r = nullptr; int32_t* d_o = nullptr; int32_t* d_g = nullptr; double* d_T = nullptr; void* d_s = nullptr;
    const size_t sb = kb_lbvh_scratch_bytes((int)maxn);
    CK(cudaMalloc((void**)&d_p, maxn * 72)); CK(cudaMalloc((void**)&d_r, maxn * 8)); CK(cudaMalloc((void**)&d_o, maxn * 4)); CK(cudaMalloc((void**)&d_g, maxn * 4)); CK(cudaMalloc((void**)&d_T, 96)); CK(cudaMalloc(&d_s, sb));
    const double I12[12] = {1, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0};
    CK(cudaMemcpy(d_T, I12, 96, cudaMemcpyHostToDevice));
    std::vector<double> hp, hr;
    for (auto& pc : e->pending_clouds) {
      if (pc.mesh) {
        const size_t nt = pc.elems.size() / 9;
        CK(cudaMemcpy(d_p, pc.elems.data(), nt * 72, cudaMemcpyHostToDevice));
        if (!pc.owners.empty()) CK(cudaMemcpy(d_o, pc.owners.data(), nt * 4, cudaMemcpyHostToDevice));
        if (!pc.origs.empty()) CK(cudaMemcpy(d_g, pc.origs.data(), nt * 4, cudaMemcpyHostToDevice));
        const DevGeom& G = *pc.dg;
        CK(kb_lbvh_build_tris(d_p, (int)nt, -1, pc.owners.empty() ? nullptr : d_o, e->d_tris64 + 9 * (size_t)G.elem_base, e->d_tris32 + 3 * (size_t)G.elem_base,
                              e->d_triown + G.elem_base, e->d_nodes + 2 * (size_t)G.node_base, d_s, sb, (int)maxn, e->stream, pc.origs.empty() ? nullptr : d_g, e->d_triorig + G.elem_base));
        CK(cudaStreamSynchronize(e->stream));
        continue;
      }
      const size_t n = pc.elems.size() / 4;
      hp.resize(3 * n); hr.resize(n);
      for (size_t i = 0; i < n; i++) { hp[3 * i] = pc.elems[4 * i]; hp[3 * i + 1] = pc.elems[4 * i + 1]; hp[3 * i + 2] = pc.elems[4 * i + 2]; hr[i] = pc.elems[4 * i + 3]; }
      CK(cudaMemcpy(d_p, hp.data(), n * 24, cudaMemcpyHostToDevice)); CK(cudaMemcpy(d_r, hr.data(), n * 8, cudaMemcpyHostToDevice));
      if (!pc.owners.empty()) CK(cudaMemcpy(d_o, pc.owners.data(), n * 4, cudaMemcpyHostToDevice));
      if (!pc.origs.empty()) CK(cudaMemcpy(d_g, pc.origs.data(), n * 4, cudaMemcpyHostToDevice));
      const DevGeom& G = *pc.dg;
      CK(kb_lbvh_build(d_p, d_r, 0.0, (int)n, d_T, -1, pc.owners.empty() ? nullptr : d_o, e->d_sph64 + 4 * (size_t)G.elem_base, e->d_sph32 + G.elem_base,
                       e->d_sphown + G.elem_base, e->d_nodes + 2 * (size_t)G.node_base, d_s, sb, (int)maxn, nullptr, e->stream, pc.origs.empty() ? nullptr : d_g, e->d_sphorig + G.elem_base));
      CK(cudaStreamSynchronize(e->stream));
    }
    cudaFree(d_p); cudaFree(d_r); cudaFree(d_o); cudaFree(d_g); cudaFree(d_T); cudaFree(d_s);
    e->pending_clouds.clear();
  }
  CK(cudaMalloc((void**)&e->d_work, 64)); CK(cudaMalloc((void**)&e->d_counters, 128)); CK(cudaMemset(e->d_counters, 0, 128));
  CK(cudaMalloc((void**)&e->d_scalars, 64));
  // the host copies of the big arrays are no longer needed
  std::vector<float>().swap(e->h_tris32); std::vector<double>().swap(e->h_tris64); std::vector<float>().swap(e->h_sph32); std::vector<double>().swap(e->h_sph64);
  std::vector<float>().swap(e->h_nodes); std::vector<float>().swap(e->h_wide);
  // Configurations per launch.  Configuration cost varies by two orders of magnitude, so every launch ends with a tail of
  // idle SMs; measured on C2 a 71 k chunk (transforms L2-resident) runs 26 % slower than a 1 M chunk (transforms through
  // HBM: 2 x 96 L bytes per configuration, ~2 % of the step).  Use up to 1 M per launch within a 2 GB scratch budget.
  int64_t per_cfg = (int64_t)L * 96 + 16;
  int64_t ch = (2048ll << 20) / per_cfg; ch = std::max<int64_t>(8192, std::min<int64_t>(ch, 1 << 20)); e->chunk = (ch / 1024) * 1024;
  e->finalized = true;
  return KB_OK;
}

// a replica of a finalized engine on another device: host-side description copied, every static device array copied peer to peer
static int clone_to_device(const kb_engine* src, int device, kb_engine** out) {
  kb_engine* r = new kb_engine(*src);
  r->replicas.clear(); r->tev.clear(); r->tev_used = 0; r->graphs.clear(); r->h_pin_in = nullptr; r->h_pin_out = nullptr; r->g_dQ = nullptr; r->g_dout = nullptr;
  r->own_stream = r->stream = r->copy_stream = r->aux_stream = nullptr; r->ev0 = r->ev1 = nullptr;
  for (int k = 0; k < KB_STAGES; k++) r->ev_stage[k] = nullptr;
  for (int k = 0; k < 4; k++) r->ev_copy[k] = nullptr;
  // per-batch scratch starts empty on the new device
  r->d_xf = nullptr; r->xf_cap = 0; r->d_state = nullptr; r->d_hit = nullptr; r->d_hit_elem = nullptr; r->cfg_cap = 0;
  r->d_leaf_list = nullptr; r->leaf_cap = 0; r->d_flagged = nullptr; r->d_state2 = nullptr; r->split_cap = 0;
  r->d_work = nullptr; r->d_counters = nullptr; r->d_Q = nullptr; r->q_cap = 0; r->d_Qf = nullptr; r->qf_cap = 0; r->d_out = nullptr; r->out_cap = 0;
  r->d_bits = nullptr; r->bits_cap = 0; r->d_pair = nullptr; r->pair_cap = 0; r->d_dist = nullptr; r->dist_cap = 0; r->d_cp = nullptr; r->cp_cap = 0;
  r->d_A = r->d_B = nullptr; r->ab_cap = 0; r->d_nlev = r->d_nchecks = r->d_firstbad = r->d_list = nullptr; r->d_alive = nullptr; r->edge_cap = 0;
  r->d_eQ = nullptr; r->d_efeas = nullptr; r->eq_cap = 0; r->d_eslot = nullptr; r->eslot_cap = 0; r->d_scalars = nullptr; r->d_weights = nullptr; r->w_cap = 0; r->d_T = nullptr; r->t_cap = 0;
  r->d_dyn_pts = nullptr; r->d_dyn_T = nullptr; r->d_dyn_scratch = nullptr; r->dyn_pts_cap = 0; r->dyn_scratch_bytes = 0;
  r->d_rays = nullptr; r->d_rid = nullptr; r->d_rdist = nullptr; r->d_relem = nullptr; r->ray_cap = 0; r->d_ignore = nullptr; r->d_rayq = nullptr; r->d_onebody = nullptr;
  memset(&r->stats, 0, sizeof r->stats);
  // static arrays: null first so that a failure half way destroys cleanly
  for (const auto& a : src->statics) *(void**)((char*)r + a.member_offset) = nullptr;
  r->device = device;
  *out = r;
  CK(cudaSetDevice(device));
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, device));
  r->num_sms = prop.multiProcessorCount;
  CK(cudaStreamCreateWithFlags(&r->own_stream, cudaStreamNonBlocking)); r->stream = r->own_stream;
  CK(cudaEventCreate(&r->ev0)); CK(cudaEventCreate(&r->ev1));
  CK(cudaStreamCreateWithFlags(&r->copy_stream, cudaStreamNonBlocking)); CK(cudaStreamCreateWithFlags(&r->aux_stream, cudaStreamNonBlocking));
  for (int k = 0; k < 4; k++) CK(cudaEventCreateWithFlags(&r->ev_copy[k], cudaEventDisableTiming));
  for (int k = 0; k < KB_STAGES; k++) CK(cudaEventCreateWithFlags(&r->ev_stage[k], cudaEventDisableTiming));
  for (const auto& a : src->statics) {
    void* d = nullptr;
    CK(cudaMalloc(&d, a.bytes));
    *(void**)((char*)r + a.member_offset) = d;
    CK(cudaMemcpyPeer(d, device, *(void* const*)((const char*)src + a.member_offset), src->device, a.bytes));
  }
  // the scene block holds device pointers: point it at this device's copies
  r->scene.nodes = r->d_nodes; r->scene.wide = r->d_wide; r->scene.tris32 = r->d_tris32; r->scene.tris64 = r->d_tris64; r->scene.sph32 = r->d_sph32; r->scene.sph64 = r->d_sph64;
  r->scene.triown = r->d_triown; r->scene.sphown = r->d_sphown; r->scene.triorig = r->d_triorig; r->scene.sphorig = r->d_sphorig;
  r->scene.box32 = r->d_box32; r->scene.box64 = r->d_box64; r->scene.boxown = r->d_boxown;
  for (int g = 0; g < KB_MAX_GRIDS; g++) if (src->scene.grids[g].data) r->scene.grids[g].data = r->d_grid[g];
  CK(cudaMalloc((void**)&r->d_work, 64)); CK(cudaMalloc((void**)&r->d_counters, 128)); CK(cudaMemset(r->d_counters, 0, 128));
  CK(cudaMalloc((void**)&r->d_scalars, 64));
  CK(cudaDeviceSynchronize());
  return KB_OK;
}

int kb_finalize_multi(kb_engine* e, const int* devices, int n_devices) {
  if (!e || e->finalized) return fail(KB_ERR_STATE, "engine is null or already finalized");
  if (!devices || n_devices < 1 || n_devices > 64) return fail(KB_ERR_INVALID, "need 1..64 devices");
  for (int i = 0; i < n_devices; i++) for (int j = 0; j < i; j++) if (devices[i] == devices[j]) return fail(KB_ERR_INVALID, "device %d listed twice", devices[i]);
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail(KB_ERR_CUDA, "no CUDA device available; this engine has no CPU fallback");
  for (int i = 0; i < n_devices; i++) if (devices[i] < 0 || devices[i] >= ndev) return fail(KB_ERR_INVALID, "device %d out of range (0..%d)", devices[i], ndev - 1);
  int rc = kb_finalize(e, devices[0]); if (rc) return rc;
  CK(cudaSetDevice(e->device)); CK(cudaDeviceSynchronize());
  for (int i = 1; i < n_devices; i++) {
    kb_engine* r = nullptr;
    rc = clone_to_device(e, devices[i], &r);
    if (r) e->replicas.push_back(r);
    if (rc) { cudaSetDevice(e->device); return rc; }
  }
  CK(cudaSetDevice(e->device));
  return KB_OK;
}
int kb_num_devices(const kb_engine* e) { return e && e->finalized ? 1 + (int)e->replicas.size() : 0; }

int kb_set_stream(kb_engine* e, void* s) {
  if (!e || !e->finalized) return fail(KB_ERR_STATE, "engine is not finalized");
  // NULL = the engine's own non-blocking stream; cudaStreamLegacy (0x1) / cudaStreamPerThread (0x2) select the default streams
  e->stream = s ? (cudaStream_t)s : e->own_stream; return KB_OK;
}

int kb_set_option(kb_engine* e, const char* name, int64_t value) {
  if (!e || !name) return fail(KB_ERR_INVALID, "null argument");
  for (kb_engine* r : e->replicas) { int rc = kb_set_option(r, name, value); if (rc) return rc; }
  if (!strcmp(name, "multi_min")) { if (value < 1) return fail(KB_ERR_INVALID, "multi_min must be >= 1"); e->multi_min = value; return KB_OK; }
  if (!strcmp(name, "collect_stats")) { e->collect_stats = value != 0; return KB_OK; }
  if (!strcmp(name, "time_kernels")) { e->time_kernels = value != 0; return KB_OK; }
  if (!strcmp(name, "pipeline")) { if (value != 0 && value != 1) return fail(KB_ERR_INVALID, "pipeline must be 0 (fused) or 1 (split)"); e->pipeline = (int)value; return KB_OK; }
  if (!strcmp(name, "leaf_budget")) { if (value < 1 || value > 100000) return fail(KB_ERR_INVALID, "leaf_budget out of range"); e->leaf_budget = (int)value; return KB_OK; }
  if (!strcmp(name, "clear_grid")) { e->use_grids = value != 0; return KB_OK; }
  if (!strcmp(name, "ray_variant")) { e->ray_variant = (int)value; for (kb_engine* r : e->replicas) r->ray_variant = (int)value; return KB_OK; }
  if (!strcmp(name, "ray_tile")) { e->ray_tile = value != 0; for (kb_engine* r : e->replicas) r->ray_tile = value != 0; return KB_OK; }
  if (!strcmp(name, "mesh_builder")) {
    if (e->finalized) return fail(KB_ERR_STATE, "mesh_builder must be set before kb_finalize");
    if (value != 0 && value != 1) return fail(KB_ERR_INVALID, "mesh_builder: 0 = binned SAH on the host (default), 1 = linear BVH on the GPU for large meshes");
    e->mesh_builder = (int)value; return KB_OK;
  }
  if (!strcmp(name, "cloud_builder")) {
    if (e->finalized) return fail(KB_ERR_STATE, "cloud_builder must be set before kb_finalize");
    if (value != 0 && value != 1) return fail(KB_ERR_INVALID, "cloud_builder: 0 = binned SAH on the host (default), 1 = linear BVH on the GPU");
    e->cloud_builder = (int)value; return KB_OK;
  }
  if (!strcmp(name, "both_limit")) { e->both_limit = (int)value; return KB_OK; }
  if (!strcmp(name, "wide")) { e->wide = value != 0; return KB_OK; }
  if (!strcmp(name, "edge_flat_max")) { if (value < 0) return fail(KB_ERR_INVALID, "edge_flat_max must be >= 0"); e->edge_flat_max = value; return KB_OK; }
  if (!strcmp(name, "zero_copy_max")) { if (value < 0) return fail(KB_ERR_INVALID, "zero_copy_max must be >= 0"); e->zero_copy_max = value; return KB_OK; }
  if (!strcmp(name, "graph_max")) { if (value < 0 || value > 65536) return fail(KB_ERR_INVALID, "graph_max must be in [0, 65536]"); if (e->h_pin_in && value > e->graph_max) return fail(KB_ERR_STATE, "graph_max can only grow before the first small batch"); e->graph_max = value; return KB_OK; }
  if (!strcmp(name, "cloud_leaf")) {
    if (e->finalized) return fail(KB_ERR_STATE, "cloud_leaf must be set before kb_finalize");
    if (value < 1 || value > 32) return fail(KB_ERR_INVALID, "cloud_leaf must be in [1, 32]");
    e->cloud_leaf = (int)value; return KB_OK;
  }
  if (!strcmp(name, "grid_res")) {
    if (e->finalized) return fail(KB_ERR_STATE, "grid_res must be set before kb_finalize");
    if (value != 0 && (value < 8 || value > 512)) return fail(KB_ERR_INVALID, "grid_res must be 0 (no clearance grids) or in [8, 512]");
    e->grid_res = (int)value; return KB_OK;
  }
  if (!strcmp(name, "chunk")) {
    if (value < 256 || value > (1 << 22)) return fail(KB_ERR_INVALID, "chunk must be in [256, 4194304]");
    e->chunk = value; return KB_OK;
  }
  return fail(KB_ERR_INVALID, "unknown option '%s'", name);
}

int kb_synchronize(kb_engine* e) {
  if (!e || !e->finalized) return fail(KB_ERR_STATE, "engine is not finalized");
  for (kb_engine* r : e->replicas) { CK(cudaSetDevice(r->device)); CK(cudaStreamSynchronize(r->stream)); }
  CK(cudaSetDevice(e->device)); CK(cudaStreamSynchronize(e->stream)); return KB_OK;
}

int kb_fk_batch(kb_engine* e, const double* Q, int64_t N, double* T_out) {
  if (!e || !e->finalized) return fail(KB_ERR_STATE, "engine is not finalized");
  if (N < 0 || (N > 0 && (!Q || !T_out))) return fail(KB_ERR_INVALID, "bad arguments");
  CK(cudaSetDevice(e->device));
  int rc = ensure_cfg_scratch(e, e->L, N); if (rc) return rc;
  if ((rc = grow(e->d_Q, e->q_cap, std::min(N, e->chunk) * e->L))) return rc;
  for (int64_t off = 0; off < N; off += e->chunk) {
    int64_t n = std::min(e->chunk, N - off);
    CK(cudaMemcpyAsync(e->d_Q, Q + off * e->L, (size_t)n * e->L * 8, cudaMemcpyHostToDevice, e->stream));
    CK(kb_launch_fk(e->d_robot, e->d_drv, e->d_drv_link, e->d_drv_scale, e->d_drv_off, e->d_Q, n, e->d_xf, e->L, nullptr, nullptr, nullptr, e->stream));
    e->stats.kernel_launches++;
    CK(cudaMemcpyAsync(T_out + off * e->L * 12, e->d_xf, (size_t)n * e->L * 96, cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
  }
  return KB_OK;
}

int kb_feasible_batch_device(kb_engine* e, const double* dQ, int64_t N, uint8_t* d_out, int32_t* d_first_pair) {
  if (!e || !e->finalized) return fail(KB_ERR_STATE, "engine is not finalized");
  if (N < 0 || (N > 0 && (!dQ || !d_out))) return fail(KB_ERR_INVALID, "bad arguments");
  CK(cudaSetDevice(e->device));
  int rc = run_feasible_device(e, dQ, N, d_out, d_first_pair, e->d_counters + 3);
  if (rc) return rc;
  e->stats.configs_checked += N;
  return KB_OK;
}

// Small batches -- the calls a planner makes between its own decisions (one configuration from RRT's extend step, a few hundred from a
// batched PRM).  Their cost is fixed overhead, not kernels: N = 1 on C2 took 64 us of which the traversal was 15.  So the caller's
// rows go through pinned staging and the whole sequence (copy in, FK, counter reset, traversal, finish, copy out) is ONE CUDA graph,
// captured once per batch size and replayed: one launch call and one synchronisation per query.  A graph remembers the scratch pointers
// it was captured with and is re-captured if a larger batch made the engine reallocate them.
static int feasible_small(kb_engine* e, const double* Q, int64_t N, uint8_t* out) {
  int rc;
  if (!e->h_pin_in) {
    const size_t cap = (size_t)std::max<int64_t>(1, e->graph_max);
    CK(cudaMallocHost((void**)&e->h_pin_in, cap * e->L * 8)); CK(cudaMallocHost((void**)&e->h_pin_out, cap));
    CK(cudaMalloc((void**)&e->g_dQ, cap * e->L * 8)); CK(cudaMalloc((void**)&e->g_dout, cap));
  }
  if ((rc = ensure_cfg_scratch(e, e->feas_items.nxf, N))) return rc;
  // The smallest batches skip both copies: FK reads the rows from the pinned staging buffer and the finish kernel writes the result
  // bytes into the pinned result buffer directly (pinned host memory is device-addressable under unified addressing) -- two graph
  // nodes fewer on a path that is all fixed overhead.
  const bool zc = N <= e->zero_copy_max;
  const double* dQ = zc ? e->h_pin_in : e->g_dQ;
  uint8_t* dout = zc ? e->h_pin_out : e->g_dout;
  const void* key[6] = {e->d_xf, e->d_state, e->d_hit, e->d_hit_elem, (const void*)e->stream, (const void*)(intptr_t)((e->use_grids ? 1 : 0) + 2 * e->chunk + ((int64_t)e->both_limit << 40) + ((int64_t)(zc ? 1 : 0) << 50))};
  kb_engine::SmallGraph* g = nullptr;
  for (auto& x : e->graphs) if (x.n == N) g = &x;
  if (g && g->exec && memcmp(g->key, key, sizeof key) != 0) { cudaGraphExecDestroy(g->exec); g->exec = nullptr; }
  if (!g) {
    if (e->graphs.size() >= 32) { for (auto& x : e->graphs) if (x.exec) cudaGraphExecDestroy(x.exec); e->graphs.clear(); }
    e->graphs.emplace_back(); g = &e->graphs.back(); g->n = N;
  }
  memcpy(e->h_pin_in, Q, (size_t)N * e->L * 8);
  if (!g->exec) {
    // one plain run first: it sets the kernels' attributes (not capturable) and leaves the answer for this very call
    if (!zc) CK(cudaMemcpyAsync(e->g_dQ, e->h_pin_in, (size_t)N * e->L * 8, cudaMemcpyHostToDevice, e->stream));
    if ((rc = run_feasible_small(e, dQ, N, dout, e->d_counters + 3, nullptr))) return rc;
    if (!zc) CK(cudaMemcpyAsync(e->h_pin_out, e->g_dout, (size_t)N, cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    memcpy(out, e->h_pin_out, (size_t)N);
    e->stats.configs_checked += N;
    // a batch size is captured the second time it is seen: a planner whose batch sizes never repeat must not pay a capture and an
    // instantiation (hundreds of microseconds) on every call
    if (g->seen++ == 0) return KB_OK;
    cudaGraph_t graph = nullptr;
    const int64_t launches_before = e->stats.kernel_launches;
    CK(cudaStreamBeginCapture(e->stream, cudaStreamCaptureModeThreadLocal));
    cudaError_t ce = zc ? cudaSuccess : cudaMemcpyAsync(e->g_dQ, e->h_pin_in, (size_t)N * e->L * 8, cudaMemcpyHostToDevice, e->stream);
    rc = ce == cudaSuccess ? run_feasible_small(e, dQ, N, dout, e->d_counters + 3, nullptr) : KB_ERR_CUDA;
    if (rc == KB_OK && !zc) ce = cudaMemcpyAsync(e->h_pin_out, e->g_dout, (size_t)N, cudaMemcpyDeviceToHost, e->stream);
    cudaError_t ce2 = cudaStreamEndCapture(e->stream, &graph);
    e->stats.kernel_launches = launches_before;                 // nothing ran during the capture
    if (rc != KB_OK || ce != cudaSuccess || ce2 != cudaSuccess || !graph) { if (graph) cudaGraphDestroy(graph); cudaGetLastError(); return KB_OK; }   // no graph: plain runs keep working
    ce = cudaGraphInstantiate(&g->exec, graph, 0);
    cudaGraphDestroy(graph);
    if (ce != cudaSuccess) { g->exec = nullptr; cudaGetLastError(); return KB_OK; }
    memcpy(g->key, key, sizeof key);
    return KB_OK;
  }
  CK(cudaGraphLaunch(g->exec, e->stream));
  CK(cudaStreamSynchronize(e->stream));
  memcpy(out, e->h_pin_out, (size_t)N);
  e->stats.configs_checked += N;
  e->stats.kernel_launches += 2;
  return KB_OK;
}

// host-buffer feasibility for configurations given as doubles (esz 8) or floats (esz 4; widened to fp64 on the device)
static int feasible_batch_host_one(kb_engine* e, const void* Qv, int esz, int64_t N, uint8_t* out, int32_t* first_pair, bool bits) {
  if (!e || !e->finalized) return fail(KB_ERR_STATE, "engine is not finalized");
  if (N < 0 || (N > 0 && (!Qv || !out))) return fail(KB_ERR_INVALID, "bad arguments");
  if (N == 0) return KB_OK;
  CK(cudaSetDevice(e->device));
  int rc;
  if (N <= e->graph_max && esz == 8 && !first_pair && !bits && e->pipeline == 0 && !e->time_kernels && !e->collect_stats) return feasible_small(e, (const double*)Qv, N, out);
  const char* Q = (const char*)Qv;
  if (esz == 4 && (rc = grow(e->d_Qf, e->qf_cap, N * e->L))) return rc;
  char* d_in = esz == 4 ? (char*)e->d_Qf : nullptr;
  if ((rc = grow(e->d_Q, e->q_cap, N * e->L))) return rc;
  if (esz == 8) d_in = (char*)e->d_Q;
  if ((rc = grow(e->d_out, e->out_cap, N))) return rc;
  if (bits && (rc = grow(e->d_bits, e->bits_cap, (N + 31) / 32))) return rc;
  if (first_pair && (rc = grow(e->d_pair, e->pair_cap, 2 * N))) return rc;
  begin_timing(e);
  // Staged upload.  A large batch crosses PCIe in KB_STAGES pieces on the copy stream; piece k is checked while piece k + 1 is in
  // flight, and consecutive pieces run on two alternating streams with their own scratch rows and work counters, so a piece starts
  // filling the SMs that the previous launch's tail has left idle.  Cost ~ copy(N / KB_STAGES) + max(copy, compute): one GPU on its own
  // x16 link is compute bound (56 MB in ~1.1 ms against 5 ms of checking); eight ranks sharing one host's memory are close to copy bound,
  // where two pieces (round 1) left the GPU waiting for the second one.  Small batches stay single-stage.
  const size_t row = (size_t)e->L * esz;
  const bool staged = N >= (1 << 17) && N <= e->chunk && e->pipeline == 0 && !e->time_kernels && !e->collect_stats;
  if (!staged) {
    CK(cudaMemcpyAsync(d_in, Q, (size_t)N * row, cudaMemcpyHostToDevice, e->stream));
    if (esz == 4) { CK(kb_launch_widen_f32(e->d_Qf, e->d_Q, N * e->L, e->stream)); e->stats.kernel_launches++; }
    if ((rc = run_feasible_device(e, e->d_Q, N, e->d_out, first_pair ? e->d_pair : nullptr, e->d_counters + 3))) return rc;
  } else {
    if ((rc = ensure_cfg_scratch(e, e->feas_items.nxf, N))) return rc;
    const int K = KB_STAGES;
    int64_t cut[KB_STAGES + 1];
    // a small first piece (1/32 of the batch: the only upload nothing overlaps with), then eighths, a quarter at the end
    static const int frac32[KB_STAGES + 1] = {0, 1, 4, 8, 12, 16, 20, 24, 32};
    for (int k = 0; k <= K; k++) cut[k] = k == K ? N : ((N * frac32[k] / 32) / 1024) * 1024;
    CK(cudaEventRecord(e->ev_copy[0], e->stream));                  // neither the copies nor the second stream may run ahead of earlier work
    CK(cudaStreamWaitEvent(e->copy_stream, e->ev_copy[0], 0));
    CK(cudaStreamWaitEvent(e->aux_stream, e->ev_copy[0], 0));
    for (int k = 0; k < K; k++) {
      CK(cudaMemcpyAsync(d_in + cut[k] * row, Q + cut[k] * row, (size_t)(cut[k + 1] - cut[k]) * row, cudaMemcpyHostToDevice, e->copy_stream));
      CK(cudaEventRecord(e->ev_stage[k], e->copy_stream));
    }
    for (int k = 0; k < K; k++) {
      cudaStream_t st = (k & 1) ? e->aux_stream : e->stream;
      const int64_t n = cut[k + 1] - cut[k];
      CK(cudaStreamWaitEvent(st, e->ev_stage[k], 0));
      if (esz == 4) { CK(kb_launch_widen_f32(e->d_Qf + cut[k] * e->L, e->d_Q + cut[k] * e->L, n * e->L, st)); e->stats.kernel_launches++; }
      if ((rc = run_feasible_piece(e, e->d_Q + cut[k] * e->L, cut[k], n, e->d_out + cut[k], first_pair ? e->d_pair + 2 * cut[k] : nullptr, e->d_counters + 3, st, k & 1))) return rc;
    }
    CK(cudaEventRecord(e->ev_copy[1], e->aux_stream));                // results are read back on e->stream: wait for the other stream's pieces
    CK(cudaStreamWaitEvent(e->stream, e->ev_copy[1], 0));
  }
  if (bits) {
    CK(kb_launch_pack_bits(e->d_out, N, e->d_bits, e->stream)); e->stats.kernel_launches++;
    CK(cudaMemcpyAsync(out, e->d_bits, (size_t)((N + 7) / 8), cudaMemcpyDeviceToHost, e->stream));
  } else CK(cudaMemcpyAsync(out, e->d_out, (size_t)N, cudaMemcpyDeviceToHost, e->stream));
  if (first_pair) CK(cudaMemcpyAsync(first_pair, e->d_pair, (size_t)N * 8, cudaMemcpyDeviceToHost, e->stream));
  end_timing(e, true);
  CK(cudaStreamSynchronize(e->stream));
  e->stats.configs_checked += N;
  return KB_OK;
}

static int feasible_batch_host(kb_engine* e, const void* Qv, int esz, int64_t N, uint8_t* out, int32_t* first_pair, bool bits = false) {
  if (!e || !e->finalized) return fail(KB_ERR_STATE, "engine is not finalized");
  if (N < 0 || (N > 0 && (!Qv || !out))) return fail(KB_ERR_INVALID, "bad arguments");
  const size_t row = (size_t)e->L * esz;
  return run_sharded(e, N, 32, [&](kb_engine* r, int64_t off, int64_t n) {
    return feasible_batch_host_one(r, (const char*)Qv + off * row, esz, n, out + (bits ? off / 8 : off), first_pair ? first_pair + 2 * off : nullptr, bits);
  });
}
int kb_feasible_batch(kb_engine* e, const double* Q, int64_t N, uint8_t* out, int32_t* first_pair) { return feasible_batch_host(e, Q, 8, N, out, first_pair); }
int kb_feasible_batch_f32(kb_engine* e, const float* Q, int64_t N, uint8_t* out, int32_t* first_pair) { return feasible_batch_host(e, Q, 4, N, out, first_pair); }
int kb_feasible_batch_bits(kb_engine* e, const double* Q, int64_t N, uint8_t* out_bits) { return feasible_batch_host(e, Q, 8, N, out_bits, nullptr, true); }
int kb_feasible_batch_bits_device(kb_engine* e, const double* dQ, int64_t N, uint8_t* d_out_bits) {
  if (!e || !e->finalized) return fail(KB_ERR_STATE, "engine is not finalized");
  if (N < 0 || (N > 0 && (!dQ || !d_out_bits))) return fail(KB_ERR_INVALID, "bad arguments");
  if (((uintptr_t)d_out_bits & 3) != 0) return fail(KB_ERR_INVALID, "d_out_bits must be 4-byte aligned (written as 32-bit words, (N + 31) / 32 of them)");
  if (N == 0) return KB_OK;
  CK(cudaSetDevice(e->device));
  int rc;
  if ((rc = grow(e->d_out, e->out_cap, N))) return rc;
  if ((rc = run_feasible_device(e, dQ, N, e->d_out, nullptr, e->d_counters + 3))) return rc;
  CK(kb_launch_pack_bits(e->d_out, N, (uint32_t*)d_out_bits, e->stream)); e->stats.kernel_launches++;
  e->stats.configs_checked += N;
  return KB_OK;
}

int kb_edges_visible_batch_device(kb_engine* e, const double* dA, const double* dB, int64_t N, double eps, const double* weights_host,
                                  uint8_t* d_out, int32_t* d_nchecks) {
  if (!e || !e->finalized) return fail(KB_ERR_STATE, "engine is not finalized");
  if (N < 0 || (N > 0 && (!dA || !dB || !d_out)) || !(eps > 0)) return fail(KB_ERR_INVALID, "bad arguments (eps must be > 0)");
  if (N == 0) return KB_OK;
  if (N > 0x7fffffff) return fail(KB_ERR_UNSUPPORTED, "more than 2^31 edges per call");
  CK(cudaSetDevice(e->device));
  int rc;
  if (N > e->edge_cap) {
    void* olds[] = {e->d_nlev, e->d_alive, e->d_nchecks, e->d_firstbad, e->d_list};
    for (void* p : olds) if (p) cudaFree(p);
    e->d_nlev = e->d_nchecks = e->d_firstbad = e->d_list = nullptr; e->d_alive = nullptr; e->edge_cap = 0;
    CK(cudaMalloc((void**)&e->d_nlev, (size_t)N * 4)); CK(cudaMalloc((void**)&e->d_alive, (size_t)N)); CK(cudaMalloc((void**)&e->d_nchecks, (size_t)N * 4));
    CK(cudaMalloc((void**)&e->d_firstbad, (size_t)N * 4)); CK(cudaMalloc((void**)&e->d_list, (size_t)N * 4));
    e->edge_cap = N;
  }
  if (e->chunk > e->eq_cap) {
    if (e->d_eQ) cudaFree(e->d_eQ); if (e->d_efeas) cudaFree(e->d_efeas); e->d_eQ = nullptr; e->d_efeas = nullptr; e->eq_cap = 0;
    CK(cudaMalloc((void**)&e->d_eQ, (size_t)e->chunk * e->L * 8)); CK(cudaMalloc((void**)&e->d_efeas, (size_t)e->chunk));
    e->eq_cap = e->chunk;
  }
  const double* d_w = nullptr;
  if (weights_host) {
    if ((rc = grow(e->d_weights, e->w_cap, (int64_t)e->jtype.size()))) return rc;
    CK(cudaMemcpyAsync(e->d_weights, weights_host, e->jtype.size() * 8, cudaMemcpyHostToDevice, e->stream));
    d_w = e->d_weights;
  }
  CK(cudaMemsetAsync(e->d_scalars, 0, 64, e->stream));
  CK(kb_launch_edge_setup(e->d_robot, dA, dB, d_w, N, eps, e->d_nlev, e->d_alive, e->d_nchecks, e->d_scalars, e->stream)); e->stats.kernel_launches++;
  CK(kb_launch_fill_i32(e->d_firstbad, N, 0x7fffffff, e->stream)); e->stats.kernel_launches++;
  int32_t head[3] = {0, 0, 0};
  CK(cudaMemcpyAsync(head, e->d_scalars, 12, cudaMemcpyDeviceToHost, e->stream));
  CK(cudaStreamSynchronize(e->stream));
  const int32_t maxlev = head[0];
  if (head[2]) return fail(KB_ERR_UNSUPPORTED, "an edge is longer than 2^24 eps: it would need more than 16 M feasibility checks (eps = %g)", eps);
  int64_t cfg_checks = 0;
  const int64_t per_max = maxlev >= 1 && maxlev <= 20 ? ((int64_t)1 << maxlev) - 1 : 0;
  if (per_max > 0 && N * per_max <= std::min<int64_t>(e->edge_flat_max, e->chunk)) {
    // small batch: every midpoint of every level at once (see kb_edge_flat_expand_kernel): one launch sequence, no read-back per level
    const int64_t nslots = N * per_max;
    if (!e->d_eslot || e->eslot_cap < e->chunk) { if (e->d_eslot) cudaFree(e->d_eslot); e->d_eslot = nullptr; CK(cudaMalloc((void**)&e->d_eslot, (size_t)e->chunk)); e->eslot_cap = e->chunk; }
    CK(kb_launch_edge_flat_expand(e->d_robot, dA, dB, e->d_nlev, e->d_alive, nslots, (int)per_max, e->d_eQ, e->d_eslot, e->d_counters + 9, e->stream)); e->stats.kernel_launches++;
    // (the midpoints of a small edge batch are a small configuration batch: run_feasible_device checks them one warp per midpoint)
    if ((rc = run_feasible_device(e, e->d_eQ, nslots, e->d_efeas, nullptr, nullptr, e->d_eslot))) return rc;
    CK(kb_launch_edge_flat_finish(e->d_efeas, e->d_eslot, nslots, (int)per_max, e->d_nlev, e->d_firstbad, N, e->d_alive, e->d_nchecks, e->stream)); e->stats.kernel_launches += 2;
  } else
  for (int lev = 1; lev <= maxlev; lev++) {
    CK(cudaMemsetAsync(e->d_scalars + 1, 0, 4, e->stream));
    CK(kb_launch_edge_count(e->d_nlev, e->d_alive, N, lev, e->d_list, (unsigned int*)(e->d_scalars + 1), e->stream)); e->stats.kernel_launches++;
    uint32_t nlist = 0;
    CK(cudaMemcpyAsync(&nlist, e->d_scalars + 1, 4, cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    if (nlist == 0) break;
    const int64_t per = (int64_t)1 << (lev - 1), total = (int64_t)nlist * per;
    for (int64_t off = 0; off < total; off += e->chunk) {
      int64_t n = std::min(e->chunk, total - off);
      CK(kb_launch_edge_expand(e->d_robot, dA, dB, e->d_list, off, n, lev, e->d_eQ, e->stream)); e->stats.kernel_launches++;
      if ((rc = run_feasible_device(e, e->d_eQ, n, e->d_efeas, nullptr, nullptr))) return rc;
      CK(kb_launch_edge_reduce(e->d_efeas, e->d_list, off, n, lev, e->d_firstbad, e->stream)); e->stats.kernel_launches++;
    }
    CK(kb_launch_edge_level_end(e->d_list, nlist, lev, e->d_firstbad, e->d_alive, e->d_nchecks, e->stream)); e->stats.kernel_launches++;
    cfg_checks += total;
  }
  CK(kb_launch_copy_u8(e->d_alive, d_out, N, e->d_counters + 4, e->stream)); e->stats.kernel_launches++;
  if (d_nchecks) CK(cudaMemcpyAsync(d_nchecks, e->d_nchecks, (size_t)N * 4, cudaMemcpyDeviceToDevice, e->stream));
  e->stats.edges_checked += N; e->edge_cfg_host += cfg_checks; e->stats.edge_config_checks += cfg_checks;
  return KB_OK;
}

static int edges_visible_host(kb_engine* e, const double* A, const double* B, int64_t N, double eps, const double* weights, uint8_t* out, int32_t* nchecks, bool bits);
static int edges_visible_host_one(kb_engine* e, const double* A, const double* B, int64_t N, double eps, const double* weights, uint8_t* out, int32_t* nchecks, bool bits);
int kb_edges_visible_batch(kb_engine* e, const double* A, const double* B, int64_t N, double eps, const double* weights, uint8_t* out, int32_t* nchecks) {
  return edges_visible_host(e, A, B, N, eps, weights, out, nchecks, false);
}
int kb_edges_visible_batch_bits(kb_engine* e, const double* A, const double* B, int64_t N, double eps, const double* weights, uint8_t* out_bits, int32_t* nchecks) {
  return edges_visible_host(e, A, B, N, eps, weights, out_bits, nchecks, true);
}
static int edges_visible_host(kb_engine* e, const double* A, const double* B, int64_t N, double eps, const double* weights, uint8_t* out, int32_t* nchecks, bool bits) {
  if (!e || !e->finalized) return fail(KB_ERR_STATE, "engine is not finalized");
  if (N < 0 || (N > 0 && (!A || !B || !out))) return fail(KB_ERR_INVALID, "bad arguments");
  const int64_t L = e->L;
  return run_sharded(e, N, 32, [&](kb_engine* r, int64_t off, int64_t n) {
    return edges_visible_host_one(r, A + off * L, B + off * L, n, eps, weights, out + (bits ? off / 8 : off), nchecks ? nchecks + off : nullptr, bits);
  });
}
static int edges_visible_host_one(kb_engine* e, const double* A, const double* B, int64_t N, double eps, const double* weights, uint8_t* out, int32_t* nchecks, bool bits) {
  if (!e || !e->finalized) return fail(KB_ERR_STATE, "engine is not finalized");
  if (N < 0 || (N > 0 && (!A || !B || !out))) return fail(KB_ERR_INVALID, "bad arguments");
  if (N == 0) return KB_OK;
  CK(cudaSetDevice(e->device));
  int rc;
  if (N * e->L > e->ab_cap) {
    if (e->d_A) cudaFree(e->d_A); if (e->d_B) cudaFree(e->d_B); e->d_A = e->d_B = nullptr; e->ab_cap = 0;
    CK(cudaMalloc((void**)&e->d_A, (size_t)N * e->L * 8)); CK(cudaMalloc((void**)&e->d_B, (size_t)N * e->L * 8)); e->ab_cap = N * e->L;
  }
  if ((rc = grow(e->d_out, e->out_cap, N))) return rc;
  if ((rc = grow(e->d_pair, e->pair_cap, 2 * N))) return rc;
  begin_timing(e);
  CK(cudaMemcpyAsync(e->d_A, A, (size_t)N * e->L * 8, cudaMemcpyHostToDevice, e->stream));
  CK(cudaMemcpyAsync(e->d_B, B, (size_t)N * e->L * 8, cudaMemcpyHostToDevice, e->stream));
  if ((rc = kb_edges_visible_batch_device(e, e->d_A, e->d_B, N, eps, weights, e->d_out, nchecks ? e->d_pair : nullptr))) return rc;
  if (bits) {
    if ((rc = grow(e->d_bits, e->bits_cap, (N + 31) / 32))) return rc;
    CK(kb_launch_pack_bits(e->d_out, N, e->d_bits, e->stream)); e->stats.kernel_launches++;
    CK(cudaMemcpyAsync(out, e->d_bits, (size_t)((N + 7) / 8), cudaMemcpyDeviceToHost, e->stream));
  } else
  CK(cudaMemcpyAsync(out, e->d_out, (size_t)N, cudaMemcpyDeviceToHost, e->stream));
  if (nchecks) CK(cudaMemcpyAsync(nchecks, e->d_pair, (size_t)N * 4, cudaMemcpyDeviceToHost, e->stream));
  end_timing(e, true);
  CK(cudaStreamSynchronize(e->stream));
  return KB_OK;
}

int kb_colliding_pairs_batch(kb_engine* e, const double* Q, int64_t N, int max_pairs, int32_t* out_pairs, int32_t* out_count) {
  if (!e || !e->finalized) return fail(KB_ERR_STATE, "engine is not finalized");
  if (N < 0 || (N > 0 && (!Q || !out_pairs || !out_count)) || max_pairs < 1 || max_pairs > 32) return fail(KB_ERR_INVALID, "bad arguments (1 <= max_pairs <= 32)");
  if (N == 0) return KB_OK;
  CK(cudaSetDevice(e->device));
  int rc;
  if ((rc = grow(e->d_Q, e->q_cap, N * e->L))) return rc;
  if ((rc = grow(e->d_pair, e->pair_cap, std::max<int64_t>(2 * N, N * max_pairs * 2 + N)))) return rc;
  if ((rc = ensure_cfg_scratch(e, e->feas_items.nxf, N))) return rc;
  int32_t* d_pairs = e->d_pair; int32_t* d_count = e->d_pair + N * max_pairs * 2;
  CK(cudaMemcpyAsync(e->d_Q, Q, (size_t)N * e->L * 8, cudaMemcpyHostToDevice, e->stream));
  for (int64_t off = 0; off < N; off += e->chunk) {
    int64_t n = std::min(e->chunk, N - off);
    CK(kb_launch_fk(e->d_robot, e->d_drv, e->d_drv_link, e->d_drv_scale, e->d_drv_off, e->d_Q + off * e->L, n, e->d_xf, e->feas_items.nxf, e->d_state, nullptr, e->d_hit, e->stream));
    KbTraverseParams p = make_params(e, e->feas_items, e->d_xf, n, e->d_state);
    if (p.nitems > 0) CK(kb_launch_allpairs(p, max_pairs, d_pairs + off * max_pairs * 2, d_count + off, e->num_sms, e->stream));
    else { CK(cudaMemsetAsync(d_pairs + off * max_pairs * 2, 0xff, (size_t)n * max_pairs * 8, e->stream)); CK(cudaMemsetAsync(d_count + off, 0, (size_t)n * 4, e->stream)); }
    e->stats.kernel_launches += 2;
  }
  CK(cudaMemcpyAsync(out_pairs, d_pairs, (size_t)N * max_pairs * 8, cudaMemcpyDeviceToHost, e->stream));
  CK(cudaMemcpyAsync(out_count, d_count, (size_t)N * 4, cudaMemcpyDeviceToHost, e->stream));
  CK(cudaStreamSynchronize(e->stream));
  return KB_OK;
}

static int distance_device(kb_engine* e, const double* dQ, int64_t N, double abs_err, double rel_err, double upper_bound, int include_self,
                           double* d_out_d, int32_t* d_out_pair, double* d_out_cp, int32_t* d_out_elem) {
  if (!e || !e->finalized) return fail(KB_ERR_STATE, "engine is not finalized");
  if (N < 0 || (N > 0 && (!dQ || !d_out_d))) return fail(KB_ERR_INVALID, "bad arguments");
  if (!(abs_err >= 0) || !(rel_err >= 0)) return fail(KB_ERR_INVALID, "absErr and relErr must be >= 0");
  CK(cudaSetDevice(e->device));
  const ItemSet& set = include_self ? e->feas_items : e->env_items;
  int rc = ensure_cfg_scratch(e, set.nxf, N); if (rc) return rc;
  if (std::isinf(upper_bound) || upper_bound > 1e300) upper_bound = 1e300;
  for (int64_t off = 0; off < N; off += e->chunk) {
    int64_t n = std::min(e->chunk, N - off);
    CK(kb_launch_fk(e->d_robot, e->d_drv, e->d_drv_link, e->d_drv_scale, e->d_drv_off, dQ + off * e->L, n, e->d_xf, set.nxf, nullptr, nullptr, e->d_hit, e->stream));
    e->stats.kernel_launches++;
    KbTraverseParams p = make_params(e, set, e->d_xf, n, nullptr);
    if (p.nitems > 0) { CK(timed_traverse(e, p, 1, d_out_d + off, upper_bound, (float)rel_err, (float)abs_err)); e->stats.kernel_launches++; }
    else return fail(KB_ERR_STATE, "no enabled geometry pairs to measure");
    if (d_out_pair) { CK(kb_launch_pair_ids(e->d_hit, e->d_hit_elem, set.d_items, e->d_triown, e->d_sphown, e->d_boxown, n, d_out_pair + 2 * off, e->stream)); e->stats.kernel_launches++; }
    if (d_out_cp || d_out_elem) {
      CK(kb_launch_closest_points(e->scene, set.d_items, e->d_xf, set.nxf, e->d_hit, e->d_hit_elem, n, d_out_cp ? d_out_cp + 6 * off : nullptr,
                                  d_out_elem ? d_out_elem + 2 * off : nullptr, e->stream));
      e->stats.kernel_launches++;
    }
  }
  return KB_OK;
}

int kb_distance_batch_device(kb_engine* e, const double* dQ, int64_t N, double upper_bound, int include_self, double* d_out_d, int32_t* d_out_pair) {
  return distance_device(e, dQ, N, 0.0, 0.0, upper_bound, include_self, d_out_d, d_out_pair, nullptr, nullptr);
}

static int distance_host_one(kb_engine* e, const double* Q, int64_t N, double abs_err, double rel_err, double upper_bound, int include_self,
                             double* out_d, int32_t* out_pair, double* out_cp, int32_t* out_elem);
static int distance_host(kb_engine* e, const double* Q, int64_t N, double abs_err, double rel_err, double upper_bound, int include_self,
                         double* out_d, int32_t* out_pair, double* out_cp, int32_t* out_elem) {
  if (!e || !e->finalized) return fail(KB_ERR_STATE, "engine is not finalized");
  if (N < 0 || (N > 0 && (!Q || !out_d))) return fail(KB_ERR_INVALID, "bad arguments");
  const int64_t L = e->L;
  return run_sharded(e, N, 1, [&](kb_engine* r, int64_t off, int64_t n) {
    return distance_host_one(r, Q + off * L, n, abs_err, rel_err, upper_bound, include_self, out_d + off, out_pair ? out_pair + 2 * off : nullptr,
                             out_cp ? out_cp + 6 * off : nullptr, out_elem ? out_elem + 2 * off : nullptr);
  });
}
static int distance_host_one(kb_engine* e, const double* Q, int64_t N, double abs_err, double rel_err, double upper_bound, int include_self,
                             double* out_d, int32_t* out_pair, double* out_cp, int32_t* out_elem) {
  if (!e || !e->finalized) return fail(KB_ERR_STATE, "engine is not finalized");
  if (N < 0 || (N > 0 && (!Q || !out_d))) return fail(KB_ERR_INVALID, "bad arguments");
  if (N == 0) return KB_OK;
  CK(cudaSetDevice(e->device));
  int rc;
  if ((rc = grow(e->d_Q, e->q_cap, N * e->L))) return rc;
  if ((rc = grow(e->d_dist, e->dist_cap, N))) return rc;
  if ((out_pair || out_elem) && (rc = grow(e->d_pair, e->pair_cap, 4 * N))) return rc;      // pair ids, then element ids
  if (out_cp && (rc = grow(e->d_cp, e->cp_cap, 6 * N))) return rc;
  begin_timing(e);
  CK(cudaMemcpyAsync(e->d_Q, Q, (size_t)N * e->L * 8, cudaMemcpyHostToDevice, e->stream));
  if ((rc = distance_device(e, e->d_Q, N, abs_err, rel_err, upper_bound, include_self, e->d_dist, out_pair ? e->d_pair : nullptr, out_cp ? e->d_cp : nullptr,
                            out_elem ? e->d_pair + 2 * N : nullptr))) return rc;
  CK(cudaMemcpyAsync(out_d, e->d_dist, (size_t)N * 8, cudaMemcpyDeviceToHost, e->stream));
  if (out_pair) CK(cudaMemcpyAsync(out_pair, e->d_pair, (size_t)N * 8, cudaMemcpyDeviceToHost, e->stream));
  if (out_elem) CK(cudaMemcpyAsync(out_elem, e->d_pair + 2 * N, (size_t)N * 8, cudaMemcpyDeviceToHost, e->stream));
  if (out_cp) CK(cudaMemcpyAsync(out_cp, e->d_cp, (size_t)N * 48, cudaMemcpyDeviceToHost, e->stream));
  end_timing(e, true);
  CK(cudaStreamSynchronize(e->stream));
  if (std::isinf(upper_bound)) for (int64_t i = 0; i < N; i++) if (out_d[i] >= 1e300) out_d[i] = INFINITY;
  return KB_OK;
}

int kb_distance_batch(kb_engine* e, const double* Q, int64_t N, double upper_bound, int include_self, double* out_d, int32_t* out_pair) {
  return distance_host(e, Q, N, 0.0, 0.0, upper_bound, include_self, out_d, out_pair, nullptr, nullptr);
}
int kb_distance_batch_ex(kb_engine* e, const double* Q, int64_t N, double abs_err, double rel_err, double upper_bound, int include_self,
                         double* out_d, int32_t* out_pair, double* out_cp, int32_t* out_elem) {
  return distance_host(e, Q, N, abs_err, rel_err, upper_bound, include_self, out_d, out_pair, out_cp, out_elem);
}

// explicit geometry pair at N transform pairs: a 1-item work list over a 2-slot transform table
static int geom_pair_query(kb_engine* e, int ga, const double* Ta, int gb, const double* Tb, int64_t N, double tol, int mode, double upper_bound,
                           uint8_t* out, double* out_d, double abs_err = 0.0, double rel_err = 0.0, double* out_cp = nullptr, int32_t* out_elem = nullptr) {
  if (!e || !e->finalized) return fail(KB_ERR_STATE, "engine is not finalized");
  if (ga < 0 || gb < 0 || ga >= (int)e->dgeoms.size() || gb >= (int)e->dgeoms.size()) return fail(KB_ERR_INVALID, "unknown geometry");
  if (N < 0 || (N > 0 && (!Ta || !Tb))) return fail(KB_ERR_INVALID, "bad arguments");
  if (N == 0) return KB_OK;
  CK(cudaSetDevice(e->device));
  const DevGeom& A = e->dgeoms[ga]; const DevGeom& B = e->dgeoms[gb];
  if (A.empty || B.empty) {   // null geometry: no collision / infinite distance (PlannerSettings.cpp:96-115)
    for (int64_t i = 0; i < N; i++) { if (out) out[i] = 0; if (out_d) out_d[i] = INFINITY; }
    return KB_OK;
  }
  ItemSet set; set.nxf = 2;
  int rc = add_item(set, A, 0, ga, B, 1, gb, false); if (rc) return rc;
  // box primitives are solid: each solid against the other geometry's elements
  if (!e->dsolid[ga].empty) { if ((rc = add_item(set, e->dsolid[ga], 0, ga, B, 1, gb, false))) return rc; }
  if (!e->dsolid[gb].empty) { if ((rc = add_item(set, A, 0, ga, e->dsolid[gb], 1, gb, false))) return rc; }
  for (KbItem& it : set.items) it.thr += tol;
  if ((rc = upload_itemset(set, nullptr))) return rc;
  std::vector<double> host((size_t)N * 24);
  for (int64_t i = 0; i < N; i++) { memcpy(&host[24 * (size_t)i], Ta + 12 * i, 96); memcpy(&host[24 * (size_t)i + 12], Tb + 12 * i, 96); }
  if ((rc = grow(e->d_T, e->t_cap, N * 24))) { cudaFree(set.d_items); return rc; }
  if ((rc = grow(e->d_dist, e->dist_cap, N))) { cudaFree(set.d_items); return rc; }
  int64_t save_chunk = e->chunk;
  e->chunk = std::max(e->chunk, N);            // the explicit-pair query runs as one launch
  rc = ensure_cfg_scratch(e, 2, N);
  e->chunk = save_chunk;
  if (rc) { cudaFree(set.d_items); return rc; }
  cudaError_t ce = cudaMemcpyAsync(e->d_T, host.data(), host.size() * 8, cudaMemcpyHostToDevice, e->stream);
  if (ce == cudaSuccess) ce = kb_launch_fill_i32(e->d_hit, N, -1, e->stream);
  KbTraverseParams p = make_params(e, set, e->d_T, N, nullptr);
  if (std::isinf(upper_bound) || upper_bound > 1e300) upper_bound = 1e300;
  if (ce == cudaSuccess) ce = kb_launch_traverse(p, mode, e->d_dist, upper_bound, e->num_sms, e->stream, nullptr, (float)rel_err, (float)abs_err);
  e->stats.kernel_launches += 2;
  if (ce == cudaSuccess && mode == 1 && (out_cp || out_elem)) {
    // geometry ids of an explicit pair are (ga, gb): the first reported point belongs to ga
    if (out_cp && grow(e->d_cp, e->cp_cap, 6 * N)) { cudaFree(set.d_items); return KB_ERR_CUDA; }
    if (out_elem && grow(e->d_pair, e->pair_cap, 2 * N)) { cudaFree(set.d_items); return KB_ERR_CUDA; }
    ce = kb_launch_closest_points(e->scene, set.d_items, e->d_T, 2, e->d_hit, e->d_hit_elem, N, out_cp ? e->d_cp : nullptr, out_elem ? e->d_pair : nullptr, e->stream);
    if (ce == cudaSuccess && out_cp) ce = cudaMemcpyAsync(out_cp, e->d_cp, (size_t)N * 48, cudaMemcpyDeviceToHost, e->stream);
    if (ce == cudaSuccess && out_elem) ce = cudaMemcpyAsync(out_elem, e->d_pair, (size_t)N * 8, cudaMemcpyDeviceToHost, e->stream);
    e->stats.kernel_launches++;
  }
  std::vector<int32_t> hit((size_t)N);
  if (ce == cudaSuccess) ce = cudaMemcpyAsync(hit.data(), e->d_hit, (size_t)N * 4, cudaMemcpyDeviceToHost, e->stream);
  if (ce == cudaSuccess && out_d) ce = cudaMemcpyAsync(out_d, e->d_dist, (size_t)N * 8, cudaMemcpyDeviceToHost, e->stream);
  if (ce == cudaSuccess) ce = cudaStreamSynchronize(e->stream);
  cudaFree(set.d_items);
  if (ce != cudaSuccess) return fail(KB_ERR_CUDA, "geometry pair query: %s", cudaGetErrorString(ce));
  if (out) for (int64_t i = 0; i < N; i++) out[i] = hit[i] >= 0;
  if (out_d) for (int64_t i = 0; i < N; i++) if (out_d[i] >= 1e300) out_d[i] = INFINITY;
  if (ga < gb) {   // the kernel orders a pair as (larger id, smaller id), the robot-vs-environment convention; here the first point belongs to ga
    if (out_cp) for (int64_t i = 0; i < N; i++) for (int k = 0; k < 3; k++) std::swap(out_cp[6 * i + k], out_cp[6 * i + 3 + k]);
    if (out_elem) for (int64_t i = 0; i < N; i++) std::swap(out_elem[2 * i], out_elem[2 * i + 1]);
  }
  return KB_OK;
}

int kb_geom_collides_batch(kb_engine* e, int ga, const double* Ta, int gb, const double* Tb, int64_t N, double tol, uint8_t* out) {
  if (!out || tol < 0) return fail(KB_ERR_INVALID, "bad arguments");
  return geom_pair_query(e, ga, Ta, gb, Tb, N, tol, 0, 0.0, out, nullptr);
}

int kb_geom_distance_batch(kb_engine* e, int ga, const double* Ta, int gb, const double* Tb, int64_t N, double upper_bound, double* out_d) {
  if (!out_d) return fail(KB_ERR_INVALID, "bad arguments");
  return geom_pair_query(e, ga, Ta, gb, Tb, N, 0.0, 1, upper_bound, nullptr, out_d);
}

int kb_geom_distance_batch_ex(kb_engine* e, int ga, const double* Ta, int gb, const double* Tb, int64_t N, double abs_err, double rel_err, double upper_bound,
                              double* out_d, double* out_cp, int32_t* out_elem) {
  if (!out_d || !(abs_err >= 0) || !(rel_err >= 0)) return fail(KB_ERR_INVALID, "bad arguments");
  return geom_pair_query(e, ga, Ta, gb, Tb, N, 0.0, 1, upper_bound, nullptr, out_d, abs_err, rel_err, out_cp, out_elem);
}

// ---- ray casting --------------------------------------------------------------------------------------------------------------
static int raycast_run(kb_engine* e, const double* q_host, const double* d_rays, int64_t N, const uint8_t* ignore_host, const KbRayBody* one_body_host,
                       int32_t* d_id, double* d_dist, int32_t* d_elem, const KbRayParams* cam = nullptr, float* d_depth = nullptr) {
  KbRayParams p; memset(&p, 0, sizeof p);
  if (cam) p = *cam;
  p.scene = e->scene; p.rays = d_rays; p.N = N; p.out_id = d_id; p.out_dist = d_dist; p.out_elem = d_elem; p.out_depth = d_depth;
  if (one_body_host) {          // Geometry3D::rayCast: one geometry at an explicit transform
    if (!e->d_onebody) CK(cudaMalloc((void**)&e->d_onebody, sizeof(KbRayBody)));
    CK(cudaMemcpyAsync(e->d_onebody, one_body_host, sizeof(KbRayBody), cudaMemcpyHostToDevice, e->stream));
    p.bodies = e->d_onebody; p.nlinkbodies = 1; p.nstatic = 0;
  } else {
    p.bodies = e->d_raybodies; p.nstatic = e->ray_nstatic; p.tlas = e->d_tlas; p.tlas_ext = e->tlas_ext; p.max_margin = e->ray_max_margin;
    p.nterr = (int)e->terrains.size(); p.nobj = (int)e->objects.size(); p.nlinks = e->L;
    const int nl = e->ray_nlink;
    if (q_host) {
      int rc = ensure_cfg_scratch(e, e->L, 1); if (rc) return rc;
      if (!e->d_rayq) CK(cudaMalloc((void**)&e->d_rayq, (size_t)KB_MAX_LINKS * 8));
      CK(cudaMemcpyAsync(e->d_rayq, q_host, (size_t)e->L * 8, cudaMemcpyHostToDevice, e->stream));
      CK(kb_launch_fk(e->d_robot, e->d_drv, e->d_drv_link, e->d_drv_scale, e->d_drv_off, e->d_rayq, 1, e->d_xf, e->L, nullptr, nullptr, nullptr, e->stream));
      e->stats.kernel_launches++;
      p.xf64 = e->d_xf; p.nlinkbodies = nl;
    } else {
      // no configuration: the robot is left out; only the replaceable clouds of the first section remain (they follow the links there)
      int nrobot = 0; for (int b = 0; b < nl; b++) if (e->ray_bodies[b].xf >= 0) nrobot++;
      p.bodies = e->d_raybodies + nrobot; p.nlinkbodies = nl - nrobot;
    }
    if (ignore_host) {
      if (!e->d_ignore) CK(cudaMalloc((void**)&e->d_ignore, (size_t)std::max(16, e->nids)));
      CK(cudaMemcpyAsync(e->d_ignore, ignore_host, (size_t)e->nids, cudaMemcpyHostToDevice, e->stream));
      p.ignore = e->d_ignore;
    }
  }
  if (p.cam_on) p.tile = e->ray_tile;
  CK(kb_launch_raycast(p, e->stream, e->ray_variant));
  e->stats.kernel_launches++;
  return KB_OK;
}

int kb_raycast_batch_device(kb_engine* e, const double* q, const double* d_rays, int64_t N, const uint8_t* ignore_ids, int32_t* d_out_id, double* d_out_dist, int32_t* d_out_elem) {
  if (!e || !e->finalized) return fail(KB_ERR_STATE, "engine is not finalized");
  if (N < 0 || (N > 0 && (!d_rays || !d_out_id || !d_out_dist))) return fail(KB_ERR_INVALID, "bad arguments");
  if (q && !all_finite(q, (size_t)e->L)) return fail(KB_ERR_INVALID, "non-finite configuration");
  if (N == 0) return KB_OK;
  CK(cudaSetDevice(e->device));
  return raycast_run(e, q, d_rays, N, ignore_ids, nullptr, d_out_id, d_out_dist, d_out_elem);
}

static int raycast_host(kb_engine* e, const double* q, const void* rays_v, int esz, int64_t N, const uint8_t* ignore_ids, const KbRayBody* one, int32_t* out_id, double* out_dist, int32_t* out_elem) {
  const char* rays = (const char*)rays_v;
  CK(cudaSetDevice(e->device));
  int rc;
  const int64_t chunk = 1 << 22;
  if (std::min(N, chunk) > e->ray_cap) {
    void* olds[] = {e->d_rays, e->d_rid, e->d_rdist, e->d_relem};
    for (void* o : olds) if (o) cudaFree(o);
    e->d_rays = nullptr; e->d_rid = nullptr; e->d_rdist = nullptr; e->d_relem = nullptr; e->ray_cap = 0;
    const int64_t cap = std::min(N, chunk);
    // 48 B per fp64 ray + 24 B per fp32 ray behind them (kb_raycast_batch_f32)
    CK(cudaMalloc((void**)&e->d_rays, (size_t)cap * 72)); CK(cudaMalloc((void**)&e->d_rid, (size_t)cap * 4));
    CK(cudaMalloc((void**)&e->d_rdist, (size_t)cap * 8)); CK(cudaMalloc((void**)&e->d_relem, (size_t)cap * 4));
    e->ray_cap = cap;
  }
  begin_timing(e);
  for (int64_t off = 0; off < N; off += chunk) {
    const int64_t n = std::min(chunk, N - off);
    if (esz == 8) CK(cudaMemcpyAsync(e->d_rays, rays + (size_t)off * 48, (size_t)n * 48, cudaMemcpyHostToDevice, e->stream));
    else {      // fp32 rays: half the upload, widened on the device (they land behind the fp64 rows of the ray scratch)
      float* d_rf = (float*)((char*)e->d_rays + (size_t)e->ray_cap * 48);
      CK(cudaMemcpyAsync(d_rf, rays + (size_t)off * 24, (size_t)n * 24, cudaMemcpyHostToDevice, e->stream));
      CK(kb_launch_widen_f32(d_rf, e->d_rays, n * 6, e->stream)); e->stats.kernel_launches++;
    }
    if ((rc = raycast_run(e, q, e->d_rays, n, ignore_ids, one, e->d_rid, e->d_rdist, e->d_relem))) return rc;
    if (out_id) CK(cudaMemcpyAsync(out_id + off, e->d_rid, (size_t)n * 4, cudaMemcpyDeviceToHost, e->stream));
    CK(cudaMemcpyAsync(out_dist + off, e->d_rdist, (size_t)n * 8, cudaMemcpyDeviceToHost, e->stream));
    if (out_elem) CK(cudaMemcpyAsync(out_elem + off, e->d_relem, (size_t)n * 4, cudaMemcpyDeviceToHost, e->stream));
    if (off + chunk < N) CK(cudaStreamSynchronize(e->stream));      // the scratch is reused by the next piece
  }
  end_timing(e, true);
  CK(cudaStreamSynchronize(e->stream));
  e->stats.rays_cast += N;
  return KB_OK;
}

int kb_raycast_batch(kb_engine* e, const double* q, const double* rays, int64_t N, const uint8_t* ignore_ids, int32_t* out_id, double* out_dist, int32_t* out_elem) {
  if (!e || !e->finalized) return fail(KB_ERR_STATE, "engine is not finalized");
  if (N < 0 || (N > 0 && (!rays || !out_id || !out_dist))) return fail(KB_ERR_INVALID, "bad arguments");
  if (q && !all_finite(q, (size_t)e->L)) return fail(KB_ERR_INVALID, "non-finite configuration");
  if (N == 0) return KB_OK;
  // a multi-device handle casts contiguous blocks of the rays on its devices (every replica runs FK for q itself)
  return run_sharded(e, N, 1, [&](kb_engine* r, int64_t off, int64_t n) {
    return raycast_host(r, q, rays + 6 * off, 8, n, ignore_ids, nullptr, out_id + off, out_dist + off, out_elem ? out_elem + off : nullptr);
  });
}

int kb_raycast_batch_f32(kb_engine* e, const double* q, const float* rays, int64_t N, const uint8_t* ignore_ids, int32_t* out_id, double* out_dist, int32_t* out_elem) {
  if (!e || !e->finalized) return fail(KB_ERR_STATE, "engine is not finalized");
  if (N < 0 || (N > 0 && (!rays || !out_id || !out_dist))) return fail(KB_ERR_INVALID, "bad arguments");
  if (q && !all_finite(q, (size_t)e->L)) return fail(KB_ERR_INVALID, "non-finite configuration");
  if (N == 0) return KB_OK;
  return run_sharded(e, N, 1, [&](kb_engine* r, int64_t off, int64_t n) {
    return raycast_host(r, q, rays + 6 * off, 4, n, ignore_ids, nullptr, out_id + off, out_dist + off, out_elem ? out_elem + off : nullptr);
  });
}

int kb_camera_depth(kb_engine* e, const double* q, const kb_camera* cam, const uint8_t* ignore_ids, float* out_depth, int32_t* out_id) {
  if (!e || !e->finalized) return fail(KB_ERR_STATE, "engine is not finalized");
  if (!cam || (!out_depth && !out_id)) return fail(KB_ERR_INVALID, "bad arguments");
  if (cam->xres <= 0 || cam->yres <= 0 || (int64_t)cam->xres * cam->yres > (1ll << 28)) return fail(KB_ERR_INVALID, "image of %d x %d pixels", cam->xres, cam->yres);
  if (!all_finite(cam->pose, 12) || !(cam->fx > 0) || !(cam->fy > 0) || !std::isfinite(cam->cx) || !std::isfinite(cam->cy) || !(cam->zmin >= 0) || !(cam->zmax >= cam->zmin))
    return fail(KB_ERR_INVALID, "camera needs a finite pose, fx, fy > 0 and 0 <= zmin <= zmax");
  if (q && !all_finite(q, (size_t)e->L)) return fail(KB_ERR_INVALID, "non-finite configuration");
  CK(cudaSetDevice(e->device));
  const int64_t N = (int64_t)cam->xres * cam->yres;
  if (N > e->ray_cap) {
    void* olds[] = {e->d_rays, e->d_rid, e->d_rdist, e->d_relem};
    for (void* o : olds) if (o) cudaFree(o);
    e->d_rays = nullptr; e->d_rid = nullptr; e->d_rdist = nullptr; e->d_relem = nullptr; e->ray_cap = 0;
    CK(cudaMalloc((void**)&e->d_rays, (size_t)N * 72)); CK(cudaMalloc((void**)&e->d_rid, (size_t)N * 4));
    CK(cudaMalloc((void**)&e->d_rdist, (size_t)N * 8)); CK(cudaMalloc((void**)&e->d_relem, (size_t)N * 4));
    e->ray_cap = N;
  }
  // camera frame: x right, y down, z forward (the sensor convention; CameraSensor::GetViewport flips y and z into OpenGL's, :890-897)
  KbRayParams cp; memset(&cp, 0, sizeof cp);
  cp.cam_on = 1; cp.xres = cam->xres; cp.yres = cam->yres; cp.cx = cam->cx; cp.cy = cam->cy; cp.zmin = cam->zmin; cp.zmax = cam->zmax;
  const double ifx = 1.0 / cam->fx, ify = 1.0 / cam->fy;
  for (int k = 0; k < 3; k++) {
    cp.eye[k] = cam->pose[9 + k]; cp.fwd[k] = cam->pose[3 * k + 2];
    cp.dx[k] = cam->pose[3 * k] * ifx; cp.dy[k] = -cam->pose[3 * k + 1] * ify;
  }
  begin_timing(e);
  float* d_depth = (float*)e->d_rdist;             // the distance scratch doubles as the float image
  int rc = raycast_run(e, q, nullptr, N, ignore_ids, nullptr, out_id ? e->d_rid : nullptr, nullptr, nullptr, &cp, out_depth ? d_depth : nullptr);
  if (rc) return rc;
  if (out_depth) CK(cudaMemcpyAsync(out_depth, d_depth, (size_t)N * 4, cudaMemcpyDeviceToHost, e->stream));
  if (out_id) CK(cudaMemcpyAsync(out_id, e->d_rid, (size_t)N * 4, cudaMemcpyDeviceToHost, e->stream));
  end_timing(e, true);
  CK(cudaStreamSynchronize(e->stream));
  e->stats.rays_cast += N;
  return KB_OK;
}

int kb_geom_raycast_batch(kb_engine* e, int geom, const double* T, const double* rays, int64_t N, int32_t* out_elem, double* out_dist) {
  if (!e || !e->finalized) return fail(KB_ERR_STATE, "engine is not finalized");
  if (geom < 0 || geom >= (int)e->geoms.size()) return fail(KB_ERR_INVALID, "unknown geometry %d", geom);
  if (N < 0 || (N > 0 && (!rays || !out_elem || !out_dist)) || (T && !all_finite(T, 12))) return fail(KB_ERR_INVALID, "bad arguments");
  if (N == 0) return KB_OK;
  const DevGeom& dg = e->dgeoms[geom];
  if (e->geoms[geom].dyn_cap > 0) return fail(KB_ERR_UNSUPPORTED, "a replaceable point cloud is cast through kb_raycast_batch (it lives in the world frame only)");
  if (dg.empty) {
    for (int64_t i = 0; i < N; i++) { out_elem[i] = -1; out_dist[i] = std::numeric_limits<double>::infinity(); }
    return KB_OK;
  }
  if (dg.depth >= 90) return fail(KB_ERR_UNSUPPORTED, "hierarchy of geometry %d is %d deep; the ray kernel's stack holds 96", geom, dg.depth);
  KbRayBody b; memset(&b, 0, sizeof b);
  b.margin = dg.margin; b.node_base = dg.node_base; b.elem_base = dg.elem_base; b.kind = dg.kind; b.id = 0; b.rank = 0; b.xf = -1; b.has_T = T ? 1 : 0;
  b.T[0] = b.T[4] = b.T[8] = 1; if (T) memcpy(b.T, T, 96);
  double ext = 0; for (int k = 0; k < 3; k++) ext = std::max(ext, std::max(std::fabs(dg.lo[k]), std::fabs(dg.hi[k])));
  b.ext = (float)(3.0 * ext * (1 + 1e-6));
  return raycast_host(e, nullptr, rays, 8, N, nullptr, &b, nullptr, out_dist, out_elem);
}

int kb_get_stats(kb_engine* e, kb_stats* out) {
  if (!e || !out) return fail(KB_ERR_INVALID, "null argument");
  if (e->finalized) {
    CK(cudaSetDevice(e->device)); CK(cudaStreamSynchronize(e->stream));
    unsigned long long c[16]; CK(cudaMemcpy(c, e->d_counters, 128, cudaMemcpyDeviceToHost));
    e->stats.recheck_pairs = (int64_t)c[0]; e->stats.node_tests = (int64_t)c[1]; e->stats.elem_tests = (int64_t)c[2];
    e->stats.configs_feasible = (int64_t)c[3]; e->stats.edges_visible = (int64_t)c[4]; e->stats.items_dropped = (int64_t)c[7]; e->stats.node_iterations = (int64_t)c[8];
    e->stats.edge_config_checks = e->edge_cfg_host + (int64_t)c[9];
    fold_kernel_times(e);
  }
  *out = e->stats;
  for (kb_engine* r : e->replicas) {     // a multi-device handle reports the sum over its devices
    kb_stats t; int rc = kb_get_stats(r, &t); if (rc) return rc;
    out->configs_checked += t.configs_checked; out->configs_feasible += t.configs_feasible; out->edges_checked += t.edges_checked;
    out->edges_visible += t.edges_visible; out->edge_config_checks += t.edge_config_checks; out->node_tests += t.node_tests;
    out->elem_tests += t.elem_tests; out->recheck_pairs += t.recheck_pairs; out->kernel_launches += t.kernel_launches;
    out->traverse_launches += t.traverse_launches; out->traverse_ms += t.traverse_ms; out->gpu_ms = std::max(out->gpu_ms, t.gpu_ms);
    out->items_dropped += t.items_dropped; out->node_iterations += t.node_iterations; out->rays_cast += t.rays_cast;
  }
  if (!e->replicas.empty()) CK(cudaSetDevice(e->device));
  return KB_OK;
}

int kb_reset_stats(kb_engine* e) {
  if (!e) return fail(KB_ERR_INVALID, "null argument");
  for (kb_engine* r : e->replicas) { int rc = kb_reset_stats(r); if (rc) return rc; }
  memset(&e->stats, 0, sizeof e->stats); e->edge_cfg_host = 0;
  e->tev_used = 0;
  if (e->finalized) { CK(cudaSetDevice(e->device)); CK(cudaMemsetAsync(e->d_counters, 0, 128, e->stream)); }
  return KB_OK;
}

int kb_get_layout(const kb_engine* e, int64_t out[6]) {
  if (!e || !e->finalized) return fail(KB_ERR_STATE, "engine is not finalized");
  int64_t nodes = 0, elems = 0;
  for (const DevGeom& g : e->dgeoms) { nodes += g.nnodes; elems += g.nelem; }
  for (const DevGeom& g : e->groups) { nodes += g.nnodes; elems += g.nelem; }
  out[0] = nodes; out[1] = elems; out[2] = e->static_bytes; out[3] = (int64_t)e->feas_items.items.size(); out[4] = (int64_t)e->groups.size(); out[5] = e->feas_items.maxdepth;
  return KB_OK;
}

}  // extern "C"
