// kb_geom.cuh -- element predicates, written once and instantiated twice:
//   T = float   : the fast path.  Every sign decision carries a static error bound; a decision whose
//                 magnitude is inside the bound returns KB_UNCERTAIN and the caller re-runs the pair in fp64.
//   T = ExactD  : fp64 with explicitly rounded (never FMA-contracted) operations, so the recheck follows the same
//                 arithmetic as a plain-C fp64 evaluation.
// Replaces (CPU): PQP TriContact / TriDist leaf tests behind AnyCollisionQuery::Collide / WithinDistance / Distance
// (call sites: reference Cpp/Planning/PlannerSettings.cpp:96-115).  The triangle-triangle overlap test is an
// orientation-predicate formulation (interval overlap on the line where the two supporting planes meet), not PQP's
// projection test: any exact test yields the same boolean because it is a geometric fact.
#pragma once
#include <cuda_runtime.h>
#include <math.h>

#define KB_NO 0
#define KB_YES 1
#define KB_UNCERTAIN 2

struct ExactD {
  double v;
  __host__ __device__ ExactD() {}
  __host__ __device__ ExactD(double x) : v(x) {}
};
#ifdef __CUDA_ARCH__
__device__ __forceinline__ ExactD operator+(ExactD a, ExactD b) { return ExactD(__dadd_rn(a.v, b.v)); }
__device__ __forceinline__ ExactD operator-(ExactD a, ExactD b) { return ExactD(__dsub_rn(a.v, b.v)); }
__device__ __forceinline__ ExactD operator*(ExactD a, ExactD b) { return ExactD(__dmul_rn(a.v, b.v)); }
__device__ __forceinline__ ExactD operator/(ExactD a, ExactD b) { return ExactD(__ddiv_rn(a.v, b.v)); }
#else
inline ExactD operator+(ExactD a, ExactD b) { return ExactD(a.v + b.v); }
inline ExactD operator-(ExactD a, ExactD b) { return ExactD(a.v - b.v); }
inline ExactD operator*(ExactD a, ExactD b) { return ExactD(a.v * b.v); }
inline ExactD operator/(ExactD a, ExactD b) { return ExactD(a.v / b.v); }
#endif
__host__ __device__ __forceinline__ ExactD operator-(ExactD a) { return ExactD(-a.v); }
__host__ __device__ __forceinline__ bool operator<(ExactD a, ExactD b) { return a.v < b.v; }
__host__ __device__ __forceinline__ bool operator>(ExactD a, ExactD b) { return a.v > b.v; }
__host__ __device__ __forceinline__ bool operator<=(ExactD a, ExactD b) { return a.v <= b.v; }
__host__ __device__ __forceinline__ bool operator>=(ExactD a, ExactD b) { return a.v >= b.v; }
__host__ __device__ __forceinline__ bool operator==(ExactD a, ExactD b) { return a.v == b.v; }

__host__ __device__ __forceinline__ float kb_abs(float a) { return fabsf(a); }
__host__ __device__ __forceinline__ ExactD kb_abs(ExactD a) { return ExactD(fabs(a.v)); }
__host__ __device__ __forceinline__ float kb_max(float a, float b) { return fmaxf(a, b); }
__host__ __device__ __forceinline__ ExactD kb_max(ExactD a, ExactD b) { return a.v > b.v ? a : b; }
__host__ __device__ __forceinline__ float kb_min(float a, float b) { return fminf(a, b); }
__host__ __device__ __forceinline__ ExactD kb_min(ExactD a, ExactD b) { return a.v < b.v ? a : b; }
__device__ __forceinline__ float kb_sqrt(float a) { return sqrtf(a); }
__device__ __forceinline__ ExactD kb_sqrt(ExactD a) { return ExactD(__dsqrt_rn(a.v)); }
__host__ __device__ __forceinline__ double kb_val(float a) { return (double)a; }
__host__ __device__ __forceinline__ double kb_val(ExactD a) { return a.v; }

template <class T> struct V3 { T x, y, z; };
template <class T> __host__ __device__ __forceinline__ V3<T> mk3(T x, T y, T z) { V3<T> r; r.x = x; r.y = y; r.z = z; return r; }
template <class T> __device__ __forceinline__ V3<T> operator-(const V3<T>& a, const V3<T>& b) { return mk3<T>(a.x - b.x, a.y - b.y, a.z - b.z); }
template <class T> __device__ __forceinline__ V3<T> operator+(const V3<T>& a, const V3<T>& b) { return mk3<T>(a.x + b.x, a.y + b.y, a.z + b.z); }
template <class T> __device__ __forceinline__ T dot(const V3<T>& a, const V3<T>& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
template <class T> __device__ __forceinline__ V3<T> cross(const V3<T>& a, const V3<T>& b) {
  return mk3<T>(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
template <class T> __device__ __forceinline__ V3<T> madd(const V3<T>& a, const V3<T>& d, T s) { return mk3<T>(a.x + s * d.x, a.y + s * d.y, a.z + s * d.z); }
template <class T> __device__ __forceinline__ T maxabs(const V3<T>& a) { return kb_max(kb_abs(a.x), kb_max(kb_abs(a.y), kb_abs(a.z))); }

// ---------------------------------------------------------------------------------------------------------------
// Filtered sign.  side(a,b,c,d) = (d-a).((b-a)x(c-a)) > 0  iff d lies on the side of plane abc its normal points to.
// fp32 error model: every coordinate of the six vertices carries an absolute error <= delta (transform rounding,
// kb_finalize derives delta from the scene extent), so each difference carries 2*delta (+ half an ulp).  With
// Mu, Mv, Mw the largest |component| of the three difference vectors u, v, w, every cofactor of an entry of row u is
// bounded by 2*Mv*Mw (and cyclically), so perturbing the 9 entries by 2*delta moves the determinant by at most
// 3*2*delta*2*(MvMw + MuMw + MuMv) = 12 delta (MuMv + MvMw + MwMu); rounding of the 3x3 expansion adds
// < 4e-6 Mu Mv Mw.  filt = 24*delta leaves a 2x safety factor.
struct FiltF { float filt; };   // fp32: filt = 24 * delta
struct FiltE {};                // exact: sign of the fp64 value, zero is zero

__device__ __forceinline__ int kb_sign(float det, float Mu, float Mv, float Mw, const FiltF& f) {
  float thr = f.filt * (Mu * Mv + Mv * Mw + Mw * Mu) + 4e-6f * Mu * Mv * Mw;
  return det > thr ? 1 : (det < -thr ? -1 : KB_UNCERTAIN);
}
__device__ __forceinline__ int kb_sign(ExactD det, ExactD, ExactD, ExactD, const FiltE&) { return det.v > 0.0 ? 1 : (det.v < 0.0 ? -1 : 0); }

template <class T, class F>
__device__ __forceinline__ int side_sign(const V3<T>& a, const V3<T>& b, const V3<T>& c, const V3<T>& d, const F& f) {
  V3<T> u = b - a, v = c - a, w = d - a;
  T det = dot(w, cross(u, v));
  return kb_sign(det, maxabs(u), maxabs(v), maxabs(w), f);
}

// exact 2D helpers for the coplanar case (fp64 only)
__device__ __forceinline__ ExactD orient2(const ExactD* a, const ExactD* b, const ExactD* c) {
  return (b[0] - a[0]) * (c[1] - a[1]) - (b[1] - a[1]) * (c[0] - a[0]); }
__device__ __forceinline__ bool onseg2(const ExactD* a, const ExactD* b, const ExactD* p) {
  return p[0].v >= fmin(a[0].v, b[0].v) && p[0].v <= fmax(a[0].v, b[0].v) && p[1].v >= fmin(a[1].v, b[1].v) && p[1].v <= fmax(a[1].v, b[1].v); }
__device__ inline bool segseg2(const ExactD* a, const ExactD* b, const ExactD* c, const ExactD* d) {
  double o1 = orient2(a, b, c).v, o2 = orient2(a, b, d).v, o3 = orient2(c, d, a).v, o4 = orient2(c, d, b).v;
  if (((o1 > 0 && o2 < 0) || (o1 < 0 && o2 > 0)) && ((o3 > 0 && o4 < 0) || (o3 < 0 && o4 > 0))) return true;
  if (o1 == 0 && onseg2(a, b, c)) return true;
  if (o2 == 0 && onseg2(a, b, d)) return true;
  if (o3 == 0 && onseg2(c, d, a)) return true;
  if (o4 == 0 && onseg2(c, d, b)) return true;
  return false;
}
__device__ inline bool pintri2(const ExactD* p, const ExactD* a, const ExactD* b, const ExactD* c) {
  double o1 = orient2(a, b, p).v, o2 = orient2(b, c, p).v, o3 = orient2(c, a, p).v;
  return (o1 >= 0 && o2 >= 0 && o3 >= 0) || (o1 <= 0 && o2 <= 0 && o3 <= 0);
}
template <class T> __device__ __forceinline__ T seg_seg_dist2(const V3<T>& p1, const V3<T>& q1, const V3<T>& p2, const V3<T>& q2);
__device__ __noinline__ int coplanar_tri_tri(const V3<ExactD>* A, const V3<ExactD>* B) {
  V3<ExactD> n = cross(B[1] - B[0], B[2] - B[0]);
  if (n.x.v == 0 && n.y.v == 0 && n.z.v == 0) n = cross(A[1] - A[0], A[2] - A[0]);
  if (n.x.v == 0 && n.y.v == 0 && n.z.v == 0) {   // two zero-area triangles: segments in space, which meet only if some pair of edges touches
#pragma unroll 1
    for (int i = 0; i < 3; i++)
#pragma unroll 1
      for (int j = 0; j < 3; j++)
        if (seg_seg_dist2<ExactD>(A[i], A[(i + 1) % 3], B[j], B[(j + 1) % 3]).v == 0.0) return KB_YES;
    return KB_NO;
  }
  int ax = 0; double m = fabs(n.x.v);
  if (fabs(n.y.v) > m) { ax = 1; m = fabs(n.y.v); }
  if (fabs(n.z.v) > m) ax = 2;
  ExactD a[3][2], b[3][2];
  for (int k = 0; k < 3; k++) {
    ExactD ca[3] = {A[k].x, A[k].y, A[k].z}, cb[3] = {B[k].x, B[k].y, B[k].z};
    a[k][0] = ca[(ax + 1) % 3]; a[k][1] = ca[(ax + 2) % 3]; b[k][0] = cb[(ax + 1) % 3]; b[k][1] = cb[(ax + 2) % 3];
  }
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++)
    if (segseg2(a[i], a[(i + 1) % 3], b[j], b[(j + 1) % 3])) return KB_YES;
  if (pintri2(a[0], b[0], b[1], b[2])) return KB_YES;
  if (pintri2(b[0], a[0], a[1], a[2])) return KB_YES;
  return KB_NO;
}
__device__ __forceinline__ int coplanar_case(const V3<float>*, const V3<float>*) { return KB_UNCERTAIN; }
__device__ __forceinline__ int coplanar_case(const V3<ExactD>* A, const V3<ExactD>* B) { return coplanar_tri_tri(A, B); }

// One triangle has no area (repeated or collinear vertices): every orientation test against ITS plane vanishes although the other
// triangle may be anywhere.  It is a segment: its edges are tested against the proper triangle P (closed segment vs closed
// triangle; sd[] = signs of D's vertices against plane(P); an edge lying in the plane is covered by the coplanar case, which
// needs all of D in the plane and was taken before this one).
__device__ __noinline__ int degenerate_tri_tri(const V3<ExactD>* D, const V3<ExactD>* P, int sd0, int sd1, int sd2) {
  const int sd[3] = {sd0, sd1, sd2};
  const FiltE f;
  for (int i = 0; i < 3; i++) {
    const int j = (i + 1) % 3, sp = sd[i], sq = sd[j];
    if ((sp > 0 && sq > 0) || (sp < 0 && sq < 0) || (sp == 0 && sq == 0)) continue;
    const int s1 = side_sign(D[i], D[j], P[0], P[1], f), s2 = side_sign(D[i], D[j], P[1], P[2], f), s3 = side_sign(D[i], D[j], P[2], P[0], f);
    if ((s1 >= 0 && s2 >= 0 && s3 >= 0) || (s1 <= 0 && s2 <= 0 && s3 <= 0)) return KB_YES;
  }
  return KB_NO;
}
__device__ __forceinline__ int degenerate_case(const V3<float>*, const V3<float>*, int, int, int) { return KB_UNCERTAIN; }
__device__ __forceinline__ int degenerate_case(const V3<ExactD>* D, const V3<ExactD>* P, int s0, int s1, int s2) { return degenerate_tri_tri(D, P, s0, s1, s2); }

// index of the vertex that is alone on its side of the other triangle's plane, and the side it is on.
// s[i] in {-1,0,+1}; not all equal-nonzero, not all zero.
__device__ __forceinline__ void lone_vertex(int s0, int s1, int s2, int& k, int& sigma) {
  int P = (s0 > 0) + (s1 > 0) + (s2 > 0), N = (s0 < 0) + (s1 < 0) + (s2 < 0);
  int want = (P == 1) ? 1 : ((N == 1) ? -1 : 0);
  k = (s0 == want) ? 0 : ((s1 == want) ? 1 : 2);
  if (want != 0) sigma = want;
  else { int o = (k == 0) ? s1 : s0; if (o == 0) o = (k == 2) ? s1 : s2; sigma = -o; }
}
template <class T> __device__ __forceinline__ void rot3(V3<T>* t, int k) {   // bring vertex k to position 0, cyclic
  if (k == 1) { V3<T> x = t[0]; t[0] = t[1]; t[1] = t[2]; t[2] = x; }
  else if (k == 2) { V3<T> x = t[2]; t[2] = t[1]; t[1] = t[0]; t[0] = x; }
}

// Do two closed triangles share a point?  KB_NO / KB_YES / KB_UNCERTAIN (fp32 only).
// Canonical form: A[0] alone on the positive side of plane(B), B[0] alone on the positive side of plane(A).  Then along
// d = nA x nB the triangles cut the common line in [j,i] (edges A0A2, A0A1) and [k,l] (edges B0B1, B0B2), and they
// overlap iff k <= i and j <= l, i.e. side(A0,A1,B0,B1) <= 0 and side(A0,A2,B0,B2) >= 0.
template <class T, class F>
__device__ __forceinline__ int tri_tri_intersect(V3<T>* A, V3<T>* B, const F& f) {
  int sa0 = side_sign(B[0], B[1], B[2], A[0], f), sa1 = side_sign(B[0], B[1], B[2], A[1], f), sa2 = side_sign(B[0], B[1], B[2], A[2], f);
  if (sa0 == sa1 && sa1 == sa2 && (sa0 == 1 || sa0 == -1)) return KB_NO;
  int sb0 = side_sign(A[0], A[1], A[2], B[0], f), sb1 = side_sign(A[0], A[1], A[2], B[1], f), sb2 = side_sign(A[0], A[1], A[2], B[2], f);
  if (sb0 == sb1 && sb1 == sb2 && (sb0 == 1 || sb0 == -1)) return KB_NO;
  if (sa0 == KB_UNCERTAIN || sa1 == KB_UNCERTAIN || sa2 == KB_UNCERTAIN || sb0 == KB_UNCERTAIN || sb1 == KB_UNCERTAIN || sb2 == KB_UNCERTAIN)
    return KB_UNCERTAIN;
  {
    const bool azero = (sa0 | sa1 | sa2) == 0, bzero = (sb0 | sb1 | sb2) == 0;
    if (azero && bzero) return coplanar_case(A, B);        // each triangle lies in the other's plane (or both are segments)
    if (azero) return degenerate_case(B, A, sb0, sb1, sb2);  // B has no area: its edges against A (sb = B's vertices against plane(A))
    if (bzero) return degenerate_case(A, B, sa0, sa1, sa2);
  }
  int ka, sga, kb, sgb;
  lone_vertex(sa0, sa1, sa2, ka, sga);
  lone_vertex(sb0, sb1, sb2, kb, sgb);
  rot3(A, ka); rot3(B, kb);
  if (sga < 0) { V3<T> x = B[1]; B[1] = B[2]; B[2] = x; }
  if (sgb < 0) { V3<T> x = A[1]; A[1] = A[2]; A[2] = x; }
  int p1 = side_sign(A[0], A[1], B[0], B[1], f);
  if (p1 == 1) return KB_NO;
  int p2 = side_sign(A[0], A[2], B[0], B[2], f);
  if (p2 == -1) return KB_NO;
  if (p1 == KB_UNCERTAIN || p2 == KB_UNCERTAIN) return KB_UNCERTAIN;
  return KB_YES;
}

// ---------------------------------------------------------------------------------------------------------------
// Is the face normal n = ab x ac (nn = |n|^2) usable?  fp64 (every product rounded separately, so a zero-area triangle
// gives exactly n = 0 and its edge regions catch every point): any non-zero normal.  fp32 (products are contracted into
// FMAs, so a zero-area triangle gives rounding noise instead of 0, and the direction of a thin triangle's normal is off by
// ~1e-7 / sin(angle)): only when sin(angle at a) >= 0.01; otherwise the fp32 result is NaN, which every caller's
// comparisons turn into "uncertain" / "evaluate in fp64".
__device__ __forceinline__ bool kb_face_ok(float nn, float ab2, float ac2) { return nn >= 1e-4f * ab2 * ac2; }
__device__ __forceinline__ bool kb_face_ok(ExactD nn, ExactD, ExactD) { return nn.v != 0.0; }
__device__ __forceinline__ float kb_no_face(float, float) { return __int_as_float(0x7fc00000); }
__device__ __forceinline__ ExactD kb_no_face(ExactD a, ExactD b) { return kb_min(a, b); }

// squared distance point - triangle (Voronoi-region walk)
template <class T>
__device__ __forceinline__ T point_tri_dist2(const V3<T>& p, const V3<T>& a, const V3<T>& b, const V3<T>& c) {
  const T zero = T(0.0f);
  V3<T> ab = b - a, ac = c - a, ap = p - a;
  T d1 = dot(ab, ap), d2 = dot(ac, ap);
  if (d1 <= zero && d2 <= zero) return dot(ap, ap);
  V3<T> bp = p - b;
  T d3 = dot(ab, bp), d4 = dot(ac, bp);
  if (d3 >= zero && d4 <= d3) return dot(bp, bp);
  T vc = d1 * d4 - d3 * d2;
  if (vc <= zero && d1 >= zero && d3 <= zero) { T v = d1 / (d1 - d3); V3<T> q = madd(a, ab, v) - p; return dot(q, q); }
  V3<T> cp = p - c;
  T d5 = dot(ab, cp), d6 = dot(ac, cp);
  if (d6 >= zero && d5 <= d6) return dot(cp, cp);
  T vb = d5 * d2 - d1 * d6;
  if (vb <= zero && d2 >= zero && d6 <= zero) { T w = d2 / (d2 - d6); V3<T> q = madd(a, ac, w) - p; return dot(q, q); }
  T va = d3 * d6 - d5 * d4;
  if (va <= zero && (d4 - d3) >= zero && (d5 - d6) >= zero) {
    T w = (d4 - d3) / ((d4 - d3) + (d5 - d6)); V3<T> q = madd(b, c - b, w) - p; return dot(q, q); }
  V3<T> n = cross(ab, ac);
  T nn = dot(n, n);
  if (!kb_face_ok(nn, dot(ab, ab), dot(ac, ac))) return kb_no_face(dot(ap, ap), kb_min(dot(bp, bp), dot(cp, cp)));
  T h = dot(ap, n);
  return h * h / nn;
}

// squared distance segment - segment (clamped closest points)
template <class T>
__device__ __forceinline__ T seg_seg_dist2(const V3<T>& p1, const V3<T>& q1, const V3<T>& p2, const V3<T>& q2) {
  const T zero = T(0.0f), one = T(1.0f);
  V3<T> d1 = q1 - p1, d2 = q2 - p2, r = p1 - p2;
  T a = dot(d1, d1), e = dot(d2, d2), f = dot(d2, r);
  T s, t;
  if (a == zero && e == zero) return dot(r, r);
  if (a == zero) { s = zero; t = f / e; t = kb_min(kb_max(t, zero), one); }
  else {
    T c = dot(d1, r);
    if (e == zero) { t = zero; s = kb_min(kb_max(-c / a, zero), one); }
    else {
      T b = dot(d1, d2), denom = a * e - b * b;
      if (denom > zero) s = kb_min(kb_max((b * f - c * e) / denom, zero), one); else s = zero;
      t = (b * s + f) / e;
      if (t < zero) { t = zero; s = kb_min(kb_max(-c / a, zero), one); }
      else if (t > one) { t = one; s = kb_min(kb_max((b - c) / a, zero), one); }
    }
  }
  V3<T> c1 = madd(p1, d1, s), c2 = madd(p2, d2, t), d = c1 - c2;
  return dot(d, d);
}

// squared distance between two closed, non-intersecting triangles: min over 9 edge pairs and 6 vertex-face pairs
template <class T>
__device__ __noinline__ T tri_tri_dist2_disjoint(const V3<T>* A, const V3<T>* B) {
  T m = seg_seg_dist2(A[0], A[1], B[0], B[1]);
#pragma unroll 1
  for (int i = 0; i < 3; i++)
#pragma unroll 1
    for (int j = 0; j < 3; j++) {
      if (i == 0 && j == 0) continue;
      m = kb_min(m, seg_seg_dist2(A[i], A[(i + 1) % 3], B[j], B[(j + 1) % 3]));
    }
#pragma unroll 1
  for (int i = 0; i < 3; i++) {
    m = kb_min(m, point_tri_dist2(A[i], B[0], B[1], B[2]));
    m = kb_min(m, point_tri_dist2(B[i], A[0], A[1], A[2]));
  }
  return m;
}
