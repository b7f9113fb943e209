// kb_closest.cu -- closest points and element indices of the pair a distance query reported.
//
// AnyCollisionQuery::Distance gives more than a number: the reference's DistanceQueryResult carries cp1 / cp2 (world frame) and
// elem1 / elem2 (Python/klampt/src/geometry.h:631-694; filled by Geometry3D.distance_ext, src/robotsim.cpp:1765-1819, and read by
// the constraint code, Cpp/Planning/NumericalConstraint.cpp:298-310).  The branch-and-bound kernel ends with the element pair that
// realises the minimum; this kernel evaluates THAT pair once more in plain fp64 and keeps the arg-min points: one thread per
// configuration, nothing hot.  Points are reported on the margin-inflated surfaces (each point moves by its geometry's margin and
// its sphere radius towards the other), so |cp2 - cp1| = d whenever d > 0.
#include "kb_types.h"
#include "kb_kernels.h"
#include <cuda_runtime.h>
#include <math.h>

namespace {

struct D3 { double x, y, z; };
__device__ __forceinline__ D3 mk(double x, double y, double z) { D3 r; r.x = x; r.y = y; r.z = z; return r; }
__device__ __forceinline__ D3 operator-(const D3& a, const D3& b) { return mk(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ D3 operator+(const D3& a, const D3& b) { return mk(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ D3 operator*(const D3& a, double s) { return mk(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ double dot(const D3& a, const D3& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ D3 cross(const D3& a, const D3& b) { return mk(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }

__device__ __forceinline__ D3 xf_point(const double* __restrict__ xf, int slot, const double* __restrict__ p) {
  if (slot < 0) return mk(p[0], p[1], p[2]);
  const double* T = xf + 12 * slot;
  return mk(T[0] * p[0] + T[1] * p[1] + T[2] * p[2] + T[9], T[3] * p[0] + T[4] * p[1] + T[5] * p[2] + T[10], T[6] * p[0] + T[7] * p[1] + T[8] * p[2] + T[11]);
}

// closest point of triangle abc to p (Voronoi regions); returns the squared distance
__device__ double closest_pt_tri(const D3& p, const D3& a, const D3& b, const D3& c, D3& q) {
  const D3 ab = b - a, ac = c - a, ap = p - a;
  const double d1 = dot(ab, ap), d2 = dot(ac, ap);
  if (d1 <= 0 && d2 <= 0) { q = a; return dot(ap, ap); }
  const D3 bp = p - b;
  const double d3 = dot(ab, bp), d4 = dot(ac, bp);
  if (d3 >= 0 && d4 <= d3) { q = b; return dot(bp, bp); }
  const double vc = d1 * d4 - d3 * d2;
  if (vc <= 0 && d1 >= 0 && d3 <= 0) { q = a + ab * (d1 / (d1 - d3)); const D3 d = p - q; return dot(d, d); }
  const D3 cp = p - c;
  const double d5 = dot(ab, cp), d6 = dot(ac, cp);
  if (d6 >= 0 && d5 <= d6) { q = c; return dot(cp, cp); }
  const double vb = d5 * d2 - d1 * d6;
  if (vb <= 0 && d2 >= 0 && d6 <= 0) { q = a + ac * (d2 / (d2 - d6)); const D3 d = p - q; return dot(d, d); }
  const double va = d3 * d6 - d5 * d4;
  if (va <= 0 && (d4 - d3) >= 0 && (d5 - d6) >= 0) { q = b + (c - b) * ((d4 - d3) / ((d4 - d3) + (d5 - d6))); const D3 d = p - q; return dot(d, d); }
  const double den = va + vb + vc;
  if (!(den != 0.0)) {            // zero-area triangle: the nearest vertex (its edges were covered by the regions above)
    const double da = dot(ap, ap), db = dot(bp, bp), dc = dot(cp, cp);
    if (da <= db && da <= dc) { q = a; return da; }
    if (db <= dc) { q = b; return db; }
    q = c; return dc;
  }
  const double v = vb / den, w = vc / den;
  q = a + ab * v + ac * w;
  const D3 d = p - q; return dot(d, d);
}

// closest points of two segments (clamped); returns the squared distance
__device__ double closest_seg_seg(const D3& p1, const D3& q1, const D3& p2, const D3& q2, D3& c1, D3& c2) {
  const D3 d1 = q1 - p1, d2 = q2 - p2, r = p1 - p2;
  const double a = dot(d1, d1), e = dot(d2, d2), f = dot(d2, r);
  double s, t;
  if (a == 0 && e == 0) { s = t = 0; }
  else if (a == 0) { s = 0; t = fmin(fmax(f / e, 0.0), 1.0); }
  else {
    const double c = dot(d1, r);
    if (e == 0) { t = 0; s = fmin(fmax(-c / a, 0.0), 1.0); }
    else {
      const double b = dot(d1, d2), denom = a * e - b * b;
      s = denom > 0 ? fmin(fmax((b * f - c * e) / denom, 0.0), 1.0) : 0.0;
      t = (b * s + f) / e;
      if (t < 0) { t = 0; s = fmin(fmax(-c / a, 0.0), 1.0); }
      else if (t > 1) { t = 1; s = fmin(fmax((b - c) / a, 0.0), 1.0); }
    }
  }
  c1 = p1 + d1 * s; c2 = p2 + d2 * t;
  const D3 d = c1 - c2; return dot(d, d);
}

// a point the closed segment pq shares with the closed triangle abc, if it crosses the triangle's plane inside it
__device__ bool seg_tri_point(const D3& p, const D3& q, const D3& a, const D3& b, const D3& c, D3& x) {
  const D3 n = cross(b - a, c - a);
  const double sp = dot(n, p - a), sq = dot(n, q - a);
  if ((sp > 0 && sq > 0) || (sp < 0 && sq < 0) || sp == sq) return false;
  x = p + (q - p) * (sp / (sp - sq));
  const double nn = dot(n, n);
  const double u = dot(cross(b - a, x - a), n), v = dot(cross(x - a, c - a), n), tol = 1e-12 * nn;
  return u >= -tol && v >= -tol && u + v <= nn + tol;
}

// closest points of two triangles: a common point if they intersect, else the best of 9 edge pairs and 6 vertex-face pairs
__device__ double closest_tri_tri(const D3* A, const D3* B, D3& ca, D3& cb) {
  for (int i = 0; i < 3; i++) {
    D3 x;
    if (seg_tri_point(A[i], A[(i + 1) % 3], B[0], B[1], B[2], x)) { ca = cb = x; return 0.0; }
    if (seg_tri_point(B[i], B[(i + 1) % 3], A[0], A[1], A[2], x)) { ca = cb = x; return 0.0; }
  }
  double best = 1e300;
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) {
    D3 c1, c2; const double d = closest_seg_seg(A[i], A[(i + 1) % 3], B[j], B[(j + 1) % 3], c1, c2);
    if (d < best) { best = d; ca = c1; cb = c2; }
  }
  for (int i = 0; i < 3; i++) {
    D3 q; double d = closest_pt_tri(A[i], B[0], B[1], B[2], q);
    if (d < best) { best = d; ca = A[i]; cb = q; }
    d = closest_pt_tri(B[i], A[0], A[1], A[2], q);
    if (d < best) { best = d; ca = q; cb = B[i]; }
  }
  return best;
}

// closest point of the solid box `bi` (frame = transform slot) to the world point pw
__device__ D3 closest_on_box(const KbScene& sc, const double* __restrict__ xf, int slot, int bi, const D3& pw) {
  const double* b = sc.box64 + 16 * (size_t)bi;
  D3 pl = pw;
  if (slot >= 0) { const double* T = xf + 12 * slot; const D3 d = mk(pw.x - T[9], pw.y - T[10], pw.z - T[11]);
    pl = mk(T[0] * d.x + T[3] * d.y + T[6] * d.z, T[1] * d.x + T[4] * d.y + T[7] * d.z, T[2] * d.x + T[5] * d.y + T[8] * d.z); }
  const D3 d = mk(pl.x - b[0], pl.y - b[1], pl.z - b[2]);
  D3 ql = mk(b[0], b[1], b[2]);
  for (int k = 0; k < 3; k++) {
    const D3 ax = mk(b[4 + 4 * k], b[5 + 4 * k], b[6 + 4 * k]);
    const double h = b[3 + 4 * k], t = fmin(fmax(dot(ax, d), -h), h);
    ql = ql + ax * t;
  }
  const double q3[3] = {ql.x, ql.y, ql.z};
  return xf_point(xf, slot, q3);
}

__device__ __forceinline__ void order_pair(unsigned flags, int& a, int& b, bool& swapped) {
  const bool self = (flags & 1u) != 0;
  swapped = self ? (a > b) : (a < b);
  if (swapped) { const int t = a; a = b; b = t; }
}

__global__ void kb_closest_points_kernel(const KbScene sc, const KbItem* __restrict__ items, const double* __restrict__ xf64, int nxf,
                                         const int32_t* __restrict__ hit, const int32_t* __restrict__ hit_elem, int64_t N,
                                         double* __restrict__ out_cp, int32_t* __restrict__ out_elem) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= N) return;
  const int h = hit[c];
  const double nan = __longlong_as_double(0x7ff8000000000000ll);
  if (h < 0) {
    if (out_cp) for (int k = 0; k < 6; k++) out_cp[6 * c + k] = nan;
    if (out_elem) { out_elem[2 * c] = -1; out_elem[2 * c + 1] = -1; }
    return;
  }
  const KbItem it = items[h];
  const int ea = hit_elem[2 * c], eb = hit_elem[2 * c + 1];
  const double* xf = xf64 + c * (int64_t)nxf * 12;
  D3 pa, pb; double ra = 0.0, rb = 0.0;
  if (it.kindA == KB_ELEM_BOX || it.kindB == KB_ELEM_BOX) {
    const bool aBox = it.kindA == KB_ELEM_BOX;
    const int ko = aBox ? it.kindB : it.kindA, so = aBox ? it.xfB : it.xfA, eo = aBox ? eb : ea;
    double r = 0.0; D3 p;
    if (ko == KB_ELEM_TRI) p = xf_point(xf, so, sc.tris64 + 9 * (size_t)eo);
    else { p = xf_point(xf, so, sc.sph64 + 4 * (size_t)eo); r = sc.sph64[4 * (size_t)eo + 3]; }
    const D3 q = closest_on_box(sc, xf, aBox ? it.xfA : it.xfB, aBox ? ea : eb, p);
    if (aBox) { pa = q; pb = p; rb = r; } else { pa = p; pb = q; ra = r; }
  } else if (it.kindA == KB_ELEM_TRI && it.kindB == KB_ELEM_TRI) {
    D3 A[3], B[3];
    for (int v = 0; v < 3; v++) { A[v] = xf_point(xf, it.xfA, sc.tris64 + 9 * (size_t)ea + 3 * v); B[v] = xf_point(xf, it.xfB, sc.tris64 + 9 * (size_t)eb + 3 * v); }
    closest_tri_tri(A, B, pa, pb);
  } else if (it.kindA == KB_ELEM_TRI) {
    D3 A[3]; for (int v = 0; v < 3; v++) A[v] = xf_point(xf, it.xfA, sc.tris64 + 9 * (size_t)ea + 3 * v);
    pb = xf_point(xf, it.xfB, sc.sph64 + 4 * (size_t)eb); rb = sc.sph64[4 * (size_t)eb + 3];
    closest_pt_tri(pb, A[0], A[1], A[2], pa);
  } else if (it.kindB == KB_ELEM_TRI) {
    D3 B[3]; for (int v = 0; v < 3; v++) B[v] = xf_point(xf, it.xfB, sc.tris64 + 9 * (size_t)eb + 3 * v);
    pa = xf_point(xf, it.xfA, sc.sph64 + 4 * (size_t)ea); ra = sc.sph64[4 * (size_t)ea + 3];
    closest_pt_tri(pa, B[0], B[1], B[2], pb);
  } else {
    pa = xf_point(xf, it.xfA, sc.sph64 + 4 * (size_t)ea); ra = sc.sph64[4 * (size_t)ea + 3];
    pb = xf_point(xf, it.xfB, sc.sph64 + 4 * (size_t)eb); rb = sc.sph64[4 * (size_t)eb + 3];
  }
  // onto the inflated surfaces: each point moves towards the other by its radius + margin
  const D3 d = pb - pa;
  const double len = sqrt(dot(d, d));
  if (len > 0.0) {
    const D3 n = d * (1.0 / len);
    pa = pa + n * (ra + it.margA); pb = pb - n * (rb + it.margB);
  }
  int ia = it.idA, ib = it.idB;
  if (ia < 0) ia = (it.kindA == KB_ELEM_TRI ? sc.triown : (it.kindA == KB_ELEM_BOX ? sc.boxown : sc.sphown))[ea];
  if (ib < 0) ib = (it.kindB == KB_ELEM_TRI ? sc.triown : (it.kindB == KB_ELEM_BOX ? sc.boxown : sc.sphown))[eb];
  bool swapped; order_pair(it.flags, ia, ib, swapped);
  int oa = it.kindA == KB_ELEM_TRI ? sc.triorig[ea] : (it.kindA == KB_ELEM_SPHERE ? sc.sphorig[ea] : 0);
  int ob = it.kindB == KB_ELEM_TRI ? sc.triorig[eb] : (it.kindB == KB_ELEM_SPHERE ? sc.sphorig[eb] : 0);
  if (swapped) { const D3 t = pa; pa = pb; pb = t; const int o = oa; oa = ob; ob = o; }
  if (out_cp) { double* o = out_cp + 6 * c; o[0] = pa.x; o[1] = pa.y; o[2] = pa.z; o[3] = pb.x; o[4] = pb.y; o[5] = pb.z; }
  if (out_elem) { out_elem[2 * c] = oa; out_elem[2 * c + 1] = ob; }
}

}  // namespace

cudaError_t kb_launch_closest_points(const KbScene& sc, const KbItem* items, const double* xf64, int nxf, const int32_t* hit, const int32_t* hit_elem,
                                     int64_t N, double* out_cp, int32_t* out_elem, cudaStream_t s) {
  if (N <= 0 || (!out_cp && !out_elem)) return cudaSuccess;
  kb_closest_points_kernel<<<(unsigned)((N + 127) / 128), 128, 0, s>>>(sc, items, xf64, nxf, hit, hit_elem, N, out_cp, out_elem);
  return cudaGetLastError();
}
