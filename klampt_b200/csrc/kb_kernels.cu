// kb_kernels.cu -- hand-written sm_100a kernels of the batched configuration-feasibility path.
//
//   kb_fk_kernel          K1+K2  batched forward kinematics over SoA joint arrays + joint / driver limits
//                                 (replaces RobotKinematics3D::UpdateFrames + RobotWithGeometry::UpdateGeometry and
//                                 SingleRobotCSpace::CheckJointLimits; call sites Cpp/Planning/RobotCSpace.cpp:610-637)
//   kb_traverse_kernel    K3-K7  warp-cooperative dual-BVH traversal for every enabled geometry pair of one
//                                 configuration: fp32 OBB tests, fp32 filtered element tests, inline fp64 recheck
//                                 (replaces WorldPlannerSettings::CheckCollision -> AnyCollisionQuery::Collide /
//                                 WithinDistance / Distance; Cpp/Planning/PlannerSettings.cpp:96-115,241-331,570-620)
//   kb_edge_* kernels     K8     edge expansion in bisection order, level by level, with per-edge early exit
//                                 (replaces EpsilonEdgeChecker::IsVisible; Cpp/Planning/RobotCSpace.cpp:835-838)
//
// No tensor cores: no stage of this path is a dense contraction.  The bound is L2/HBM latency and bandwidth on
// BVH nodes, so the design keeps all 32 lanes of a warp on one configuration's frontier of node pairs.
#include "kb_types.h"
#include "kb_geom.cuh"
#include "kb_kernels.h"
#include <cuda_runtime.h>
#include <stdio.h>

#define FULL 0xffffffffu

// =============================================================================================== FK (fp64)
// One thread per configuration.  All products use explicitly rounded fp64 operations (no FMA contraction) so the
// result follows the same arithmetic as the reference's scalar fp64 code up to the last ulp of sin/cos.
struct Xf64 { ExactD r[9]; ExactD t[3]; };

__device__ __forceinline__ void xf_mul(const Xf64& A, const Xf64& B, Xf64& C) {
#pragma unroll
  for (int i = 0; i < 3; i++) {
#pragma unroll
    for (int j = 0; j < 3; j++) C.r[3 * i + j] = A.r[3 * i] * B.r[j] + A.r[3 * i + 1] * B.r[3 + j] + A.r[3 * i + 2] * B.r[6 + j];
    C.t[i] = A.r[3 * i] * B.t[0] + A.r[3 * i + 1] * B.t[1] + A.r[3 * i + 2] * B.t[2] + A.t[i];
  }
}

#define KB_FK_THREADS 128
__global__ void __launch_bounds__(KB_FK_THREADS)
kb_fk_kernel(const KbRobotDev* __restrict__ robot, const KbDriverDev* __restrict__ drv, const int32_t* __restrict__ drv_link,
             const double* __restrict__ drv_scale, const double* __restrict__ drv_off,
             const double* __restrict__ Q, int64_t N, double* __restrict__ xf64, int nxf,
             uint8_t* __restrict__ state, const uint8_t* __restrict__ alive, int32_t* __restrict__ hit) {
  __shared__ KbRobotDev R;
  // per warp a staging tile of 32 configurations x 12 doubles (row stride 13: conflict-free): the transform of one link leaves
  // as 96-byte runs (three full sectors per configuration) instead of 32 partial sectors per store instruction
  __shared__ double stage_all[KB_FK_THREADS / 32][32 * 13];
  {
    const int* src = (const int*)robot; int* dst = (int*)&R;
    for (int i = threadIdx.x; i < (int)(sizeof(KbRobotDev) / 4); i += blockDim.x) dst[i] = src[i];
  }
  __syncthreads();
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  const int64_t c_warp = c - lane;
  if (c_warp >= N) return;
  double* stage = stage_all[threadIdx.x >> 5];
  const int L = R.L;
  const double* q = Q + (c < N ? c : 0) * L;
  bool want_xf = false;
  if (c < N) {
    if (hit) hit[c] = -1;
    if (alive && !alive[c]) { if (state) state[c] = 0; }
    else {
      // K2: joint limits (closed interval, Normal / Weld joints only) and driver limits
      bool ok = true;
      for (int j = 0; j < R.nj; j++) {
        int t = R.jtype[j];
        if (t == 1 || t == 0) { int k = R.jlink[j]; double v = q[k]; if (v < R.qmin[k] || v > R.qmax[k]) ok = false; }
      }
      for (int d = 0; d < R.ndrv; d++) {
        ExactD v(0.0);
        for (int k = drv[d].first; k < drv[d].first + drv[d].n; k++) v = v + (ExactD(q[drv_link[k]]) - ExactD(drv_off[k])) / ExactD(drv_scale[k]);
        v = v / ExactD((double)drv[d].n);
        if (v.v < drv[d].dmin || v.v > drv[d].dmax) ok = false;
      }
      if (state) state[c] = ok ? 1 : 0;
      want_xf = ok || !state;          // infeasible by limits: the traversal skips it, no transforms needed
    }
  }
  const unsigned wm = __ballot_sync(FULL, want_xf);
  if (!wm) return;
  // K1: T_World[i] = T_World[parent] * (T0_Parent[i] * T_loc(q_i))
  double* out = xf64 + (c < N ? c : 0) * (int64_t)nxf * 12;
  double* out_warp = xf64 + c_warp * (int64_t)nxf * 12;
  Xf64 prev;                          // transform of link i-1 stays in registers (chains)
  for (int i = 0; i < L; i++) {
    if (want_xf) {
      Xf64 T0, loc, rel, W;
#pragma unroll
      for (int k = 0; k < 9; k++) T0.r[k] = ExactD(R.T0[12 * i + k]);
#pragma unroll
      for (int k = 0; k < 3; k++) T0.t[k] = ExactD(R.T0[12 * i + 9 + k]);
      ExactD wx(R.axis[3 * i]), wy(R.axis[3 * i + 1]), wz(R.axis[3 * i + 2]), qi(q[i]);
      if (R.linktype[i] == 1) {
        loc.r[0] = loc.r[4] = loc.r[8] = ExactD(1.0);
        loc.r[1] = loc.r[2] = loc.r[3] = loc.r[5] = loc.r[6] = loc.r[7] = ExactD(0.0);
        loc.t[0] = qi * wx; loc.t[1] = qi * wy; loc.t[2] = qi * wz;
      } else {
        double sn, cs; sincos(qi.v, &sn, &cs);
        ExactD s(sn), co(cs), v = ExactD(1.0) - co;
        loc.r[0] = co + v * wx * wx;      loc.r[1] = v * wx * wy - s * wz; loc.r[2] = v * wx * wz + s * wy;
        loc.r[3] = v * wy * wx + s * wz;  loc.r[4] = co + v * wy * wy;     loc.r[5] = v * wy * wz - s * wx;
        loc.r[6] = v * wz * wx - s * wy;  loc.r[7] = v * wz * wy + s * wx; loc.r[8] = co + v * wz * wz;
        loc.t[0] = loc.t[1] = loc.t[2] = ExactD(0.0);
      }
      xf_mul(T0, loc, rel);
      int par = R.parents[i];
      if (par < 0) W = rel;
      else if (par == i - 1) xf_mul(prev, rel, W);
      else {                            // branch: the parent's transform was written by this warp earlier in the loop
        Xf64 P;
#pragma unroll
        for (int k = 0; k < 9; k++) P.r[k] = ExactD(__ldcg(out + 12 * par + k));
#pragma unroll
        for (int k = 0; k < 3; k++) P.t[k] = ExactD(__ldcg(out + 12 * par + 9 + k));
        xf_mul(P, rel, W);
      }
#pragma unroll
      for (int k = 0; k < 9; k++) stage[lane * 13 + k] = W.r[k].v;
#pragma unroll
      for (int k = 0; k < 3; k++) stage[lane * 13 + 9 + k] = W.t[k].v;
      prev = W;
    }
    __syncwarp();
    for (int idx = lane; idx < 32 * 12; idx += 32) {
      const int r = idx / 12, k2 = idx - 12 * r;
      if ((wm >> r) & 1u) out_warp[(size_t)r * nxf * 12 + 12 * i + k2] = stage[r * 13 + k2];
    }
    __syncwarp();
  }
}

// =============================================================================================== traversal helpers
// reported id pairs follow the reference's order: (link, environment id) for robot-vs-environment pairs (ids1 = {robot}, ids2 =
// {terrains, objects} in CheckCollisionFree), (lower link, higher link) for self pairs -- whichever side the engine put first
__device__ __forceinline__ void kb_order_pair(unsigned flags, int& a, int& b) {
  const bool self = (flags & 1u) != 0;
  if (self ? (a > b) : (a < b)) { const int t = a; a = b; b = t; }
}

struct XfF { float r[9]; float t[3]; };

// T = A^-1 * B for two slots of the per-warp fp32 transform table; slot < 0 = identity
__device__ __forceinline__ void rel_xf(const float* __restrict__ xfw, int sa, int sb, XfF& T) {
  if (sa < 0 && sb < 0) {
    T.r[0] = T.r[4] = T.r[8] = 1.f; T.r[1] = T.r[2] = T.r[3] = T.r[5] = T.r[6] = T.r[7] = 0.f; T.t[0] = T.t[1] = T.t[2] = 0.f;
    return;
  }
  if (sa < 0) {
    const float4* b = (const float4*)(xfw + 12 * sb);
    float4 b0 = b[0], b1 = b[1], b2 = b[2];
    T.r[0] = b0.x; T.r[1] = b0.y; T.r[2] = b0.z; T.r[3] = b0.w; T.r[4] = b1.x; T.r[5] = b1.y; T.r[6] = b1.z; T.r[7] = b1.w; T.r[8] = b2.x;
    T.t[0] = b2.y; T.t[1] = b2.z; T.t[2] = b2.w;
    return;
  }
  const float4* a = (const float4*)(xfw + 12 * sa);
  float4 a0 = a[0], a1 = a[1], a2 = a[2];
  float A[9] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w, a2.x};
  float ta[3] = {a2.y, a2.z, a2.w};
  if (sb < 0) {
#pragma unroll
    for (int i = 0; i < 3; i++) {
#pragma unroll
      for (int j = 0; j < 3; j++) T.r[3 * i + j] = A[3 * j + i];
      T.t[i] = -(A[i] * ta[0] + A[3 + i] * ta[1] + A[6 + i] * ta[2]);
    }
    return;
  }
  const float4* b = (const float4*)(xfw + 12 * sb);
  float4 b0 = b[0], b1 = b[1], b2 = b[2];
  float B[9] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w, b2.x};
  float d[3] = {b2.y - ta[0], b2.z - ta[1], b2.w - ta[2]};
#pragma unroll
  for (int i = 0; i < 3; i++) {
#pragma unroll
    for (int j = 0; j < 3; j++) T.r[3 * i + j] = A[i] * B[j] + A[3 + i] * B[3 + j] + A[6 + i] * B[6 + j];
    T.t[i] = A[i] * d[0] + A[3 + i] * d[1] + A[6 + i] * d[2];
  }
}

__device__ __forceinline__ V3<float> xform(const XfF& T, const float4& p) {
  return mk3<float>(T.r[0] * p.x + T.r[1] * p.y + T.r[2] * p.z + T.t[0],
                    T.r[3] * p.x + T.r[4] * p.y + T.r[5] * p.z + T.t[1],
                    T.r[6] * p.x + T.r[7] * p.y + T.r[8] * p.z + T.t[2]);
}

// fp64 world-frame element fetch for the exact recheck; same operation order as a scalar fp64 R*p+t
__device__ __forceinline__ V3<ExactD> xform64(const double* __restrict__ xf, int slot, const double* __restrict__ p) {
  ExactD x(p[0]), y(p[1]), z(p[2]);
  if (slot < 0) return mk3<ExactD>(x, y, z);
  const double* T = xf + 12 * slot;
  return mk3<ExactD>(ExactD(T[0]) * x + ExactD(T[1]) * y + ExactD(T[2]) * z + ExactD(T[9]),
                     ExactD(T[3]) * x + ExactD(T[4]) * y + ExactD(T[5]) * z + ExactD(T[10]),
                     ExactD(T[6]) * x + ExactD(T[7]) * y + ExactD(T[8]) * z + ExactD(T[11]));
}

// ---- solid boxes (the interior of a box primitive; its surface is 12 ordinary triangles).  Every element of the other
// geometry is measured against the solid through one reference point -- a triangle's first vertex, a sphere's centre:
// d = dist(reference point, solid box) - radius, 0 for a point inside.  For a triangle outside the box the surface distance is
// the true one and never larger than this term; for a triangle inside, the surfaces do not meet and this term is 0.
__device__ __forceinline__ float point_box_dist32(const float4* __restrict__ bx, const V3<float>& p) {
  const float4 c = __ldg(bx), u0 = __ldg(bx + 1), u1 = __ldg(bx + 2), u2 = __ldg(bx + 3);
  const float dx = p.x - c.x, dy = p.y - c.y, dz = p.z - c.z;
  const float g0 = fmaxf(fabsf(dx * u0.x + dy * u0.y + dz * u0.z) - c.w, 0.f);
  const float g1 = fmaxf(fabsf(dx * u1.x + dy * u1.y + dz * u1.z) - u0.w, 0.f);
  const float g2 = fmaxf(fabsf(dx * u2.x + dy * u2.y + dz * u2.z) - u1.w, 0.f);
  return sqrtf(g0 * g0 + g1 * g1 + g2 * g2);
}
// squared fp64 distance from a world-frame point to the solid box `bi` whose frame is transform slot `slot`
__device__ __noinline__ ExactD point_box_dist2_64(const KbScene& sc, const double* __restrict__ xf, int slot, int bi, const V3<ExactD>& pw) {
  V3<ExactD> pl = pw;
  if (slot >= 0) {
    const double* T = xf + 12 * slot;
    const V3<ExactD> d = mk3<ExactD>(pw.x - ExactD(T[9]), pw.y - ExactD(T[10]), pw.z - ExactD(T[11]));
    pl = mk3<ExactD>(ExactD(T[0]) * d.x + ExactD(T[3]) * d.y + ExactD(T[6]) * d.z, ExactD(T[1]) * d.x + ExactD(T[4]) * d.y + ExactD(T[7]) * d.z,
                     ExactD(T[2]) * d.x + ExactD(T[5]) * d.y + ExactD(T[8]) * d.z);
  }
  const double* b = sc.box64 + 16 * (size_t)bi;
  const V3<ExactD> d = mk3<ExactD>(pl.x - ExactD(b[0]), pl.y - ExactD(b[1]), pl.z - ExactD(b[2]));
  ExactD s(0.0);
#pragma unroll
  for (int k = 0; k < 3; k++) {
    const ExactD q = ExactD(b[4 + 4 * k]) * d.x + ExactD(b[5 + 4 * k]) * d.y + ExactD(b[6 + 4 * k]) * d.z;
    const ExactD g = kb_abs(q) - ExactD(b[3 + 4 * k]);
    if (g.v > 0.0) s = s + g * g;
  }
  return s;
}
// reference point (world frame) and radius of element `el` of a TRI / SPHERE side
__device__ __forceinline__ V3<ExactD> elem_ref_point64(const KbScene& sc, const double* __restrict__ xf, int kind, int slot, int el, double& radius) {
  if (kind == KB_ELEM_TRI) { radius = 0.0; return xform64(xf, slot, sc.tris64 + 9 * (size_t)el); }
  radius = sc.sph64[4 * (size_t)el + 3];
  return xform64(xf, slot, sc.sph64 + 4 * (size_t)el);
}

// exact (fp64) distance between two elements in the world frame, minus sphere radii; 0 when triangles intersect
template <bool BOXES>
__device__ __noinline__ double exact_elem_distance(const KbScene& sc, const KbItem& it, const double* __restrict__ xf,
                                                   int ea, int eb) {
  if (BOXES && (it.kindA == KB_ELEM_BOX || it.kindB == KB_ELEM_BOX)) {
    const bool aBox = it.kindA == KB_ELEM_BOX;
    double r;
    const V3<ExactD> p = aBox ? elem_ref_point64(sc, xf, it.kindB, it.xfB, eb, r) : elem_ref_point64(sc, xf, it.kindA, it.xfA, ea, r);
    return (kb_sqrt(point_box_dist2_64(sc, xf, aBox ? it.xfA : it.xfB, aBox ? ea : eb, p)) - ExactD(r)).v;
  }
  if (it.kindA == KB_ELEM_TRI && it.kindB == KB_ELEM_TRI) {
    V3<ExactD> A[3], B[3];
#pragma unroll
    for (int v = 0; v < 3; v++) { A[v] = xform64(xf, it.xfA, sc.tris64 + 9 * (size_t)ea + 3 * v); B[v] = xform64(xf, it.xfB, sc.tris64 + 9 * (size_t)eb + 3 * v); }
    V3<ExactD> A2[3] = {A[0], A[1], A[2]}, B2[3] = {B[0], B[1], B[2]};
    if (tri_tri_intersect<ExactD, FiltE>(A2, B2, FiltE()) == KB_YES) return 0.0;
    return kb_sqrt(tri_tri_dist2_disjoint<ExactD>(A, B)).v;
  }
  if (it.kindA == KB_ELEM_TRI) {
    V3<ExactD> A[3];
#pragma unroll
    for (int v = 0; v < 3; v++) A[v] = xform64(xf, it.xfA, sc.tris64 + 9 * (size_t)ea + 3 * v);
    V3<ExactD> p = xform64(xf, it.xfB, sc.sph64 + 4 * (size_t)eb);
    return (kb_sqrt(point_tri_dist2<ExactD>(p, A[0], A[1], A[2])) - ExactD(sc.sph64[4 * (size_t)eb + 3])).v;
  }
  if (it.kindB == KB_ELEM_TRI) {
    V3<ExactD> B[3];
#pragma unroll
    for (int v = 0; v < 3; v++) B[v] = xform64(xf, it.xfB, sc.tris64 + 9 * (size_t)eb + 3 * v);
    V3<ExactD> p = xform64(xf, it.xfA, sc.sph64 + 4 * (size_t)ea);
    return (kb_sqrt(point_tri_dist2<ExactD>(p, B[0], B[1], B[2])) - ExactD(sc.sph64[4 * (size_t)ea + 3])).v;
  }
  V3<ExactD> p = xform64(xf, it.xfA, sc.sph64 + 4 * (size_t)ea), s = xform64(xf, it.xfB, sc.sph64 + 4 * (size_t)eb), d = p - s;
  return (kb_sqrt(dot(d, d)) - ExactD(sc.sph64[4 * (size_t)ea + 3]) - ExactD(sc.sph64[4 * (size_t)eb + 3])).v;
}

// exact boolean: elements within thr of each other (thr == 0 and two triangles: surfaces intersect)
template <bool BOXES>
__device__ __noinline__ bool exact_elem_collide(const KbScene& sc, const KbItem& it, const double* __restrict__ xf, int ea, int eb) {
  if (BOXES && (it.kindA == KB_ELEM_BOX || it.kindB == KB_ELEM_BOX)) {
    const bool aBox = it.kindA == KB_ELEM_BOX;
    double r;
    const V3<ExactD> p = aBox ? elem_ref_point64(sc, xf, it.kindB, it.xfB, eb, r) : elem_ref_point64(sc, xf, it.kindA, it.xfA, ea, r);
    return (kb_sqrt(point_box_dist2_64(sc, xf, aBox ? it.xfA : it.xfB, aBox ? ea : eb, p)) - ExactD(r)).v <= it.thr;
  }
  if (it.kindA == KB_ELEM_TRI && it.kindB == KB_ELEM_TRI) {
    V3<ExactD> A[3], B[3];
#pragma unroll
    for (int v = 0; v < 3; v++) { A[v] = xform64(xf, it.xfA, sc.tris64 + 9 * (size_t)ea + 3 * v); B[v] = xform64(xf, it.xfB, sc.tris64 + 9 * (size_t)eb + 3 * v); }
    V3<ExactD> A2[3] = {A[0], A[1], A[2]}, B2[3] = {B[0], B[1], B[2]};
    if (tri_tri_intersect<ExactD, FiltE>(A2, B2, FiltE()) == KB_YES) return true;
    if (it.thr == 0.0) return false;
    ExactD d2 = tri_tri_dist2_disjoint<ExactD>(A, B);
    return d2.v <= (ExactD(it.thr) * ExactD(it.thr)).v;
  }
  // sphere cases: compare squared centre distance with (radius + thr)^2 like the scalar fp64 code does
  if (it.kindA == KB_ELEM_TRI || it.kindB == KB_ELEM_TRI) {
    bool aTri = it.kindA == KB_ELEM_TRI;
    int et = aTri ? ea : eb, es = aTri ? eb : ea, st = aTri ? it.xfA : it.xfB, ss = aTri ? it.xfB : it.xfA;
    V3<ExactD> Tt[3];
#pragma unroll
    for (int v = 0; v < 3; v++) Tt[v] = xform64(xf, st, sc.tris64 + 9 * (size_t)et + 3 * v);
    V3<ExactD> p = xform64(xf, ss, sc.sph64 + 4 * (size_t)es);
    ExactD r = ExactD(sc.sph64[4 * (size_t)es + 3]) + ExactD(it.thr);
    return point_tri_dist2<ExactD>(p, Tt[0], Tt[1], Tt[2]).v <= (r * r).v;
  }
  V3<ExactD> p = xform64(xf, it.xfA, sc.sph64 + 4 * (size_t)ea), s = xform64(xf, it.xfB, sc.sph64 + 4 * (size_t)eb), d = p - s;
  ExactD r = ExactD(sc.sph64[4 * (size_t)ea + 3]) + ExactD(sc.sph64[4 * (size_t)eb + 3]) + ExactD(it.thr);
  return dot(d, d).v <= (r * r).v;
}

// fp32 distance (radius subtracted) between the solid box on one side of an item and the reference point of the element on the
// other side; T maps B's frame into A's.  (Inlined: as a separate function its XfF argument forced the relative transform of every
// element phase through local memory -- C2 6.23 -> 6.99 ms.)
__device__ __forceinline__ float fast_box_elem_distance(const KbScene& sc, const KbItem& it, const XfF& T, int ea, int eb) {
  if (it.kindA == KB_ELEM_BOX) {           // box in A's frame, reference point of B's element mapped into it
    float r = 0.f; V3<float> p;
    if (it.kindB == KB_ELEM_TRI) p = xform(T, __ldg(sc.tris32 + 3 * (size_t)eb));
    else { const float4 sp = __ldg(sc.sph32 + eb); r = sp.w; p = xform(T, sp); }
    return point_box_dist32(sc.box32 + 4 * (size_t)ea, p) - r;
  }
  float r = 0.f; float4 q;                 // box in B's frame: A's reference point through the inverse of T
  if (it.kindA == KB_ELEM_TRI) q = __ldg(sc.tris32 + 3 * (size_t)ea); else { q = __ldg(sc.sph32 + ea); r = q.w; }
  const float dx = q.x - T.t[0], dy = q.y - T.t[1], dz = q.z - T.t[2];
  const V3<float> p = mk3<float>(T.r[0] * dx + T.r[3] * dy + T.r[6] * dz, T.r[1] * dx + T.r[4] * dy + T.r[7] * dz, T.r[2] * dx + T.r[5] * dy + T.r[8] * dz);
  return point_box_dist32(sc.box32 + 4 * (size_t)eb, p) - r;
}

// fp32 filtered boolean for one element pair in A's frame.  Returns KB_NO / KB_YES / KB_UNCERTAIN.
template <bool BOXES>
__device__ __forceinline__ int fast_elem_collide(const KbScene& sc, const KbItem& it, const XfF& T, int ea, int eb, float thr) {
  const float delta = sc.eps_abs;
  const float band = 16.f * delta;
  if (BOXES && (it.kindA == KB_ELEM_BOX || it.kindB == KB_ELEM_BOX)) {
    const float d = fast_box_elem_distance(sc, it, T, ea, eb);
    return d < thr - band ? KB_YES : (d > thr + band ? KB_NO : KB_UNCERTAIN);
  }
  if (it.kindA == KB_ELEM_TRI && it.kindB == KB_ELEM_TRI) {
    const float4* ta = sc.tris32 + 3 * (size_t)ea; const float4* tb = sc.tris32 + 3 * (size_t)eb;
    float4 a0 = __ldg(ta), a1 = __ldg(ta + 1), a2 = __ldg(ta + 2), b0 = __ldg(tb), b1 = __ldg(tb + 1), b2 = __ldg(tb + 2);
    V3<float> A[3] = {mk3<float>(a0.x, a0.y, a0.z), mk3<float>(a1.x, a1.y, a1.z), mk3<float>(a2.x, a2.y, a2.z)};
    V3<float> B[3] = {xform(T, b0), xform(T, b1), xform(T, b2)};
    FiltF f; f.filt = 24.f * delta;
    if (thr == 0.f) return tri_tri_intersect<float, FiltF>(A, B, f);
    V3<float> A2[3] = {A[0], A[1], A[2]}, B2[3] = {B[0], B[1], B[2]};
    int r = tri_tri_intersect<float, FiltF>(A2, B2, f);
    if (r != KB_NO) return r;
    {   // a thin triangle's fp32 face normal is not accurate enough for the vertex-face distances: leave the pair to fp64
      const V3<float> ea1 = A[1] - A[0], ea2 = A[2] - A[0], eb1 = B[1] - B[0], eb2 = B[2] - B[0];
      const V3<float> nA = cross(ea1, ea2), nB = cross(eb1, eb2);
      if (!kb_face_ok(dot(nA, nA), dot(ea1, ea1), dot(ea2, ea2)) || !kb_face_ok(dot(nB, nB), dot(eb1, eb1), dot(eb2, eb2))) return KB_UNCERTAIN;
    }
    float d = sqrtf(tri_tri_dist2_disjoint<float>(A, B));
    return d < thr - band ? KB_YES : (d > thr + band ? KB_NO : KB_UNCERTAIN);
  }
  if (it.kindA == KB_ELEM_TRI) {          // triangle (A frame) vs sphere of B
    const float4* ta = sc.tris32 + 3 * (size_t)ea;
    float4 a0 = __ldg(ta), a1 = __ldg(ta + 1), a2 = __ldg(ta + 2), s = __ldg(sc.sph32 + eb);
    V3<float> p = xform(T, s);
    float d = sqrtf(point_tri_dist2<float>(p, mk3<float>(a0.x, a0.y, a0.z), mk3<float>(a1.x, a1.y, a1.z), mk3<float>(a2.x, a2.y, a2.z)));
    float r = s.w + thr;
    return d < r - band ? KB_YES : (d > r + band ? KB_NO : KB_UNCERTAIN);
  }
  if (it.kindB == KB_ELEM_TRI) {          // sphere of A vs triangle of B (mapped into A's frame)
    const float4* tb = sc.tris32 + 3 * (size_t)eb;
    float4 b0 = __ldg(tb), b1 = __ldg(tb + 1), b2 = __ldg(tb + 2), s = __ldg(sc.sph32 + ea);
    float d = sqrtf(point_tri_dist2<float>(mk3<float>(s.x, s.y, s.z), xform(T, b0), xform(T, b1), xform(T, b2)));
    float r = s.w + thr;
    return d < r - band ? KB_YES : (d > r + band ? KB_NO : KB_UNCERTAIN);
  }
  float4 sa = __ldg(sc.sph32 + ea), sb = __ldg(sc.sph32 + eb);
  V3<float> p = xform(T, sb), q = mk3<float>(sa.x, sa.y, sa.z), dv = p - q;
  float d = sqrtf(dot(dv, dv)), r = sa.w + sb.w + thr;
  return d < r - band ? KB_YES : (d > r + band ? KB_NO : KB_UNCERTAIN);
}


// fp32 lower estimate of the element distance (radii subtracted) with absolute error <= 16*delta, used by the distance
// kernel to skip the fp64 evaluation of pairs that cannot improve the running minimum.  Point / sphere pairs: the distance
// itself; triangle pairs: a plane-separation lower bound (80 flops instead of the 15-feature distance).
template <bool BOXES>
__device__ __forceinline__ float fast_elem_distance(const KbScene& sc, const KbItem& it, const XfF& T, int ea, int eb) {
  if (BOXES && (it.kindA == KB_ELEM_BOX || it.kindB == KB_ELEM_BOX)) return fast_box_elem_distance(sc, it, T, ea, eb);
  if (it.kindA == KB_ELEM_TRI && it.kindB == KB_ELEM_TRI) {
    const float4* ta = sc.tris32 + 3 * (size_t)ea; const float4* tb = sc.tris32 + 3 * (size_t)eb;
    float4 a0 = __ldg(ta), a1 = __ldg(ta + 1), a2 = __ldg(ta + 2), b0 = __ldg(tb), b1 = __ldg(tb + 1), b2 = __ldg(tb + 2);
    V3<float> A[3] = {mk3<float>(a0.x, a0.y, a0.z), mk3<float>(a1.x, a1.y, a1.z), mk3<float>(a2.x, a2.y, a2.z)};
    V3<float> B[3] = {xform(T, b0), xform(T, b1), xform(T, b2)};
    // cheap lower bound instead of the full 15-feature distance: if one triangle lies entirely on one side of the other's
    // plane, the distance is at least its smallest vertex-to-plane distance
    // The fp32 normal of a thin triangle points in a direction that is off by ~1e-7 / sin(angle), and that of a zero-area
    // triangle is pure rounding noise (the products are contracted into FMAs, so cross(u, u) != 0): the signed values
    // below carry an error of up to 4e-7 |e1| |e2| L (L = largest vertex offset), so 1e-6 |e1| |e2| L is taken off
    // before they count as a separation.  A well-shaped triangle loses < 1e-6 L of its bound, a degenerate one gets none.
    const V3<float> ea1 = A[1] - A[0], ea2 = A[2] - A[0], eb1 = B[1] - B[0], eb2 = B[2] - B[0];
    const V3<float> nA = cross(ea1, ea2), nB = cross(eb1, eb2);
    const V3<float> w0 = B[0] - A[0], w1 = B[1] - A[0], w2 = B[2] - A[0], u1 = A[1] - B[0], u2 = A[2] - B[0];
    const float b0s = dot(nA, w0), b1s = dot(nA, w1), b2s = dot(nA, w2);
    const float a0s = -dot(nB, w0), a1s = dot(nB, u1), a2s = dot(nB, u2);
    const float slackA = 1e-6f * sqrtf(dot(ea1, ea1) * dot(ea2, ea2) * fmaxf(dot(w0, w0), fmaxf(dot(w1, w1), dot(w2, w2))));
    const float slackB = 1e-6f * sqrtf(dot(eb1, eb1) * dot(eb2, eb2) * fmaxf(dot(w0, w0), fmaxf(dot(u1, u1), dot(u2, u2))));
    float lbA = 0.f, lbB = 0.f;
    if ((b0s > 0.f && b1s > 0.f && b2s > 0.f) || (b0s < 0.f && b1s < 0.f && b2s < 0.f))
      lbA = fmaxf(fminf(fabsf(b0s), fminf(fabsf(b1s), fabsf(b2s))) - slackA, 0.f) * rsqrtf(fmaxf(dot(nA, nA), 1e-30f));
    if ((a0s > 0.f && a1s > 0.f && a2s > 0.f) || (a0s < 0.f && a1s < 0.f && a2s < 0.f))
      lbB = fmaxf(fminf(fabsf(a0s), fminf(fabsf(a1s), fabsf(a2s))) - slackB, 0.f) * rsqrtf(fmaxf(dot(nB, nB), 1e-30f));
    return fmaxf(lbA, lbB) * (1.f - 1e-5f);
  }
  if (it.kindA == KB_ELEM_TRI) {
    const float4* ta = sc.tris32 + 3 * (size_t)ea;
    float4 a0 = __ldg(ta), a1 = __ldg(ta + 1), a2 = __ldg(ta + 2), s = __ldg(sc.sph32 + eb);
    return sqrtf(point_tri_dist2<float>(xform(T, s), mk3<float>(a0.x, a0.y, a0.z), mk3<float>(a1.x, a1.y, a1.z), mk3<float>(a2.x, a2.y, a2.z))) - s.w;
  }
  if (it.kindB == KB_ELEM_TRI) {
    const float4* tb = sc.tris32 + 3 * (size_t)eb;
    float4 b0 = __ldg(tb), b1 = __ldg(tb + 1), b2 = __ldg(tb + 2), s = __ldg(sc.sph32 + ea);
    return sqrtf(point_tri_dist2<float>(mk3<float>(s.x, s.y, s.z), xform(T, b0), xform(T, b1), xform(T, b2))) - s.w;
  }
  float4 sa = __ldg(sc.sph32 + ea), sb = __ldg(sc.sph32 + eb);
  V3<float> p = xform(T, sb), q = mk3<float>(sa.x, sa.y, sa.z), dv = p - q;
  return sqrtf(dot(dv, dv)) - sa.w - sb.w;
}

// Deferred fp64 rechecks.  Element pairs whose fp32 result is inside the error band are parked in a per-warp queue
// (configuration, item, elemA, elemB) that survives from one configuration to the next, and are re-run in fp64 32 at a
// time so the slow path also runs on full warps.  A configuration whose traversal ended without a certain hit is
// written as "no hit" and upgraded here if one of its parked pairs turns out to collide.
#ifndef KB_RQ_CAP
#define KB_RQ_CAP 96
#endif
#ifndef KB_LEAF_FIRST
#define KB_LEAF_FIRST 16   // leaf pairs that trigger the FIRST element phase of a configuration (a colliding configuration's first leaf pairs are
                           // likely hits); later phases wait for KB_LEAF_TRIGGER.  Measured: 16 -> C2 6.23 ms / C3 7.83 ms, 32 -> 6.30 / 7.94, 8 -> 6.26 / 7.87
#endif
#ifndef KB_BOOL_LEAFQ_CAP
#define KB_BOOL_LEAFQ_CAP KB_LEAFQ_CAP   // leaf-pair queue of the boolean kernel (>= KB_LEAF_TRIGGER + 32)
#endif
template <bool BOXES>
__device__ __forceinline__ void drain_rechecks(const KbTraverseParams& p, uint4* rq, int* rq_count, int lane, int64_t cur_c,
                                               int& found, int& found_ea, int& found_eb, bool all) {
  __syncwarp();
  int rqn = *rq_count;
  if (rqn > KB_RQ_CAP) rqn = KB_RQ_CAP;
  while (rqn >= (all ? 1 : 32)) {
    const int m = rqn < 32 ? rqn : 32;
    bool yes = false;
    uint4 e = make_uint4(0u, 0u, 0u, 0u);
    if (lane < m) {
      e = rq[rqn - 1 - lane];
      const int64_t c = (int64_t)e.x;
      const bool moot = (c == cur_c) ? (found >= 0) : (p.hit[c] >= 0);
      if (!moot) {
        const KbItem& it = p.items[e.y];
        yes = exact_elem_collide<BOXES>(p.scene, it, p.xf64 + c * (int64_t)p.nxf * 12, (int)e.z, (int)e.w);
      }
    }
    rqn -= m;
    unsigned ym = __ballot_sync(FULL, yes);
    while (ym) {
      const int src = __ffs(ym) - 1; ym &= ym - 1;
      const int64_t c = (int64_t)__shfl_sync(FULL, e.x, src);
      const int item = (int)__shfl_sync(FULL, e.y, src), ea = (int)__shfl_sync(FULL, e.z, src), eb = (int)__shfl_sync(FULL, e.w, src);
      if (c == cur_c) { if (found < 0) { found = item; found_ea = ea; found_eb = eb; } }
      else if (lane == 0 && p.hit[c] < 0) { p.hit[c] = item; if (p.hit_elem) { p.hit_elem[2 * c] = ea; p.hit_elem[2 * c + 1] = eb; } }
    }
    __syncwarp();
  }
  __syncwarp();                       // every lane has read the counter before lane 0 rewrites it
  if (lane == 0) *rq_count = rqn;
  __syncwarp();
}

// =============================================================================================== traversal (boolean)
// kb_traverse_kernel -- collide / within-threshold for every enabled geometry pair of one configuration.
// One warp per configuration (persistent CTAs, configurations handed out by an atomic counter with guided chunk sizes).
// The warp keeps a LIFO frontier of (item, nodeA, nodeB) pairs in shared memory; every iteration of the tight inner
// loop the 32 lanes pop up to 32 pairs, load both nodes (one 256-bit load each), run a 6-axis separating-axis test
// (the face normals of both boxes; conservative, ~3x cheaper than the 15-axis test for +4 % node visits on C2) and push
// the two children of every overlapping pair (descend the larger box) with ballot/popc compaction.  Leaf pairs go to a
// second queue drained 32 at a time so the element tests also run on full warps; any certain hit ends the
// configuration for all lanes (__ballot_sync early exit).  Per configuration the relative transform of every work item
// is computed once into shared memory (ITC) and a 16 B static record per item is cached per CTA, so a node test is
// LDS.64 + 4 x LDS.128 + 2 x LDG.256 + ~56 FP instructions (165 SASS instructions per 32-pair iteration in total).

// one 32-byte node = one 256-bit load (LDG.E.ENL2.256 on sm_100) instead of two 128-bit ones
#ifdef KB_QNODES
// experimental 16-byte quantised node: u16 centre[3], u16 half[3] on one scene-wide grid, i32 ref (inner: left child; leaf:
// -1 - (first*8 + count-1)).  Needs the scene for the grid, so it is a macro-selected overload.
#define load_node(nodes, idx, n0, n1) load_node_q(sc, nodes, idx, n0, n1)
__device__ __forceinline__ void load_node_q(const KbScene& sc, const float4* __restrict__ nodes, size_t idx, float4& n0, float4& n1) {
  const uint4 w = __ldg((const uint4*)nodes + idx);
  n0.x = fmaf(__uint2float_rn(w.x & 0xffffu), sc.qs[0], sc.qo[0]); n0.y = fmaf(__uint2float_rn(w.x >> 16), sc.qs[1], sc.qo[1]);
  n0.z = fmaf(__uint2float_rn(w.y & 0xffffu), sc.qs[2], sc.qo[2]);
  n1.x = __uint2float_rn(w.y >> 16) * sc.qs[0]; n1.y = __uint2float_rn(w.z & 0xffffu) * sc.qs[1]; n1.z = __uint2float_rn(w.z >> 16) * sc.qs[2];
  const int ref = (int)w.w;
  if (ref >= 0) { n0.w = __int_as_float(ref); n1.w = __int_as_float(0); }
  else { const int code = -1 - ref; n0.w = __int_as_float(~(code >> 3)); n1.w = __int_as_float((code & 7) + 1); }
}
#else
__device__ __forceinline__ void load_node(const float4* __restrict__ nodes, size_t idx, float4& n0, float4& n1) {
  unsigned long long x0, x1, x2, x3;
  asm volatile("ld.global.nc.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(x0), "=l"(x1), "=l"(x2), "=l"(x3) : "l"(nodes + 2 * idx));
  n0.x = __uint_as_float((unsigned)x0); n0.y = __uint_as_float((unsigned)(x0 >> 32)); n0.z = __uint_as_float((unsigned)x1); n0.w = __uint_as_float((unsigned)(x1 >> 32));
  n1.x = __uint_as_float((unsigned)x2); n1.y = __uint_as_float((unsigned)(x2 >> 32)); n1.z = __uint_as_float((unsigned)x3); n1.w = __uint_as_float((unsigned)(x3 >> 32));
}
#endif

struct ItemS { int32_t nodeA, nodeB; float infl; int32_t xf; };   // 16 B static per-item record cached per block

__device__ __forceinline__ bool sat6_overlap(const float4& ac, const float4& ah, const float4& bc, const float4& bh, const XfF& T, float infl) {
  const float hax = ah.x + infl, hay = ah.y + infl, haz = ah.z + infl;
  const float tx = T.r[0] * bc.x + T.r[1] * bc.y + T.r[2] * bc.z + T.t[0] - ac.x;
  const float ty = T.r[3] * bc.x + T.r[4] * bc.y + T.r[5] * bc.z + T.t[1] - ac.y;
  const float tz = T.r[6] * bc.x + T.r[7] * bc.y + T.r[8] * bc.z + T.t[2] - ac.z;
  bool sep = fabsf(tx) > hax + fabsf(T.r[0]) * bh.x + fabsf(T.r[1]) * bh.y + fabsf(T.r[2]) * bh.z;
  sep |= fabsf(ty) > hay + fabsf(T.r[3]) * bh.x + fabsf(T.r[4]) * bh.y + fabsf(T.r[5]) * bh.z;
  sep |= fabsf(tz) > haz + fabsf(T.r[6]) * bh.x + fabsf(T.r[7]) * bh.y + fabsf(T.r[8]) * bh.z;
  sep |= fabsf(tx * T.r[0] + ty * T.r[3] + tz * T.r[6]) > bh.x + hax * fabsf(T.r[0]) + hay * fabsf(T.r[3]) + haz * fabsf(T.r[6]);
  sep |= fabsf(tx * T.r[1] + ty * T.r[4] + tz * T.r[7]) > bh.y + hax * fabsf(T.r[1]) + hay * fabsf(T.r[4]) + haz * fabsf(T.r[7]);
  sep |= fabsf(tx * T.r[2] + ty * T.r[5] + tz * T.r[8]) > bh.z + hax * fabsf(T.r[2]) + hay * fabsf(T.r[5]) + haz * fabsf(T.r[8]);
  return !sep;
}

// lower bound on the distance between the two boxes: per-axis gaps in A's frame and in B's frame
__device__ __forceinline__ float box_dist_lb(const float4& ac, const float4& ah, const float4& bc, const float4& bh, const XfF& T) {
  const float tx = T.r[0] * bc.x + T.r[1] * bc.y + T.r[2] * bc.z + T.t[0] - ac.x;
  const float ty = T.r[3] * bc.x + T.r[4] * bc.y + T.r[5] * bc.z + T.t[1] - ac.y;
  const float tz = T.r[6] * bc.x + T.r[7] * bc.y + T.r[8] * bc.z + T.t[2] - ac.z;
  float g, g2a = 0.f, g2b = 0.f;
  g = fmaxf(fabsf(tx) - ah.x - (fabsf(T.r[0]) * bh.x + fabsf(T.r[1]) * bh.y + fabsf(T.r[2]) * bh.z), 0.f); g2a += g * g;
  g = fmaxf(fabsf(ty) - ah.y - (fabsf(T.r[3]) * bh.x + fabsf(T.r[4]) * bh.y + fabsf(T.r[5]) * bh.z), 0.f); g2a += g * g;
  g = fmaxf(fabsf(tz) - ah.z - (fabsf(T.r[6]) * bh.x + fabsf(T.r[7]) * bh.y + fabsf(T.r[8]) * bh.z), 0.f); g2a += g * g;
  g = fmaxf(fabsf(tx * T.r[0] + ty * T.r[3] + tz * T.r[6]) - bh.x - (ah.x * fabsf(T.r[0]) + ah.y * fabsf(T.r[3]) + ah.z * fabsf(T.r[6])), 0.f); g2b += g * g;
  g = fmaxf(fabsf(tx * T.r[1] + ty * T.r[4] + tz * T.r[7]) - bh.y - (ah.x * fabsf(T.r[1]) + ah.y * fabsf(T.r[4]) + ah.z * fabsf(T.r[7])), 0.f); g2b += g * g;
  g = fmaxf(fabsf(tx * T.r[2] + ty * T.r[5] + tz * T.r[8]) - bh.z - (ah.x * fabsf(T.r[2]) + ah.y * fabsf(T.r[5]) + ah.z * fabsf(T.r[8])), 0.f); g2b += g * g;
  return sqrtf(fmaxf(g2a, g2b));
}

__device__ __forceinline__ void load_itc(const float* __restrict__ itc, int item, XfF& T) {
  const float4* q = (const float4*)(itc + 12 * item);
  const float4 q0 = q[0], q1 = q[1], q2 = q[2];
  T.r[0] = q0.x; T.r[1] = q0.y; T.r[2] = q0.z; T.r[3] = q0.w; T.r[4] = q1.x; T.r[5] = q1.y; T.r[6] = q1.z; T.r[7] = q1.w; T.r[8] = q2.x;
  T.t[0] = q2.y; T.t[1] = q2.z; T.t[2] = q2.w;
}

// one popped frontier entry: fetch its item record + relative transform + both nodes, 6-axis SAT, and on overlap either
// mark a leaf pair or produce the two child entries (descend the larger box)
// KB_BOTH_MODE: 0 = a surviving inner pair always splits its larger box; 1 = pairs of comparable boxes split both boxes at
// once (four child pairs: one level of each tree per iteration); 2 = like 1 but only while the frontier is at most option
// both_limit entries (experiment knob).  KB_BOTH_RATIO bounds the ratio of the squared box diagonals that counts as comparable.
#ifndef KB_BOTH_MODE
#define KB_BOTH_MODE 1
#endif
#ifndef KB_BOTH_RATIO
#define KB_BOTH_RATIO 4.f
#endif
template <bool ITC>
__device__ __forceinline__ void node_test(const KbTraverseParams& p, const ItemS* __restrict__ s_items, const float* __restrict__ itc,
                                          const float* __restrict__ xfw, float slack, const uint2 e, const bool allow_both,
                                          bool& push2, bool& push4, bool& leafpair, uint2& c0e, uint2& c1e) {
  const KbScene& sc = p.scene;
  const int item = (int)(e.x >> KB_NODEA_BITS);
  int nodeA, nodeB; float infl; XfF T;
  if (ITC) {
    const ItemS s = s_items[item];
    nodeA = s.nodeA; nodeB = s.nodeB; infl = s.infl;
    load_itc(itc, item, T);
  } else {
    const KbItem* itp = p.items + item;
    nodeA = itp->nodeA; nodeB = itp->nodeB; infl = (float)itp->thr + slack;
    rel_xf(xfw, itp->xfA, itp->xfB, T);
  }
  const int na = (int)(e.x & (KB_MAX_NODES_A - 1)), nb = (int)e.y;
  float4 a0, a1, b0, b1;
  load_node(sc.nodes, (size_t)(nodeA + na), a0, a1);
  load_node(sc.nodes, (size_t)(nodeB + nb), b0, b1);
  if (sat6_overlap(a0, a1, b0, b1, T, infl)) {
    const int la = __float_as_int(a0.w), lb = __float_as_int(b0.w);
    if (la < 0 && lb < 0) leafpair = true;
    else {
      push2 = true;
      const float sa2 = a1.x * a1.x + a1.y * a1.y + a1.z * a1.z, sb2 = b1.x * b1.x + b1.y * b1.y + b1.z * b1.z;
      if (KB_BOTH_MODE && allow_both && la >= 0 && lb >= 0 && sa2 < KB_BOTH_RATIO * sb2 && sb2 < KB_BOTH_RATIO * sa2) {
        // comparable boxes and a narrow frontier: descend both trees at once (c0e = (a0, b0); the other three are derived at the push)
        push4 = true;
        c0e = make_uint2((e.x & ~(unsigned)(KB_MAX_NODES_A - 1)) | (unsigned)la, (unsigned)lb);
        c1e = make_uint2(c0e.x + 1u, (unsigned)lb);
      } else if (lb < 0 || (la >= 0 && sa2 >= sb2)) {
        c0e = make_uint2((e.x & ~(unsigned)(KB_MAX_NODES_A - 1)) | (unsigned)la, e.y);
        c1e = make_uint2(c0e.x + 1u, e.y);
      } else {
        c0e = make_uint2(e.x, (unsigned)lb);
        c1e = make_uint2(e.x, (unsigned)lb + 1u);
      }
    }
  }
}

// BOXES: the work list holds solid-box items.  The box predicates -- fp32 and the fp64 recheck -- are compiled only into that
// instantiation so the common kernel keeps its register allocation: with the box branch inside the shared, not inlined
// exact_elem_collide, its larger clobber set cost the calling kernel spills on the hot path (C2 6.23 -> 6.99 ms, C3 7.83 -> 9.59).
template <bool ITC, bool STATS, int BPS, bool BOXES, bool STATIC>
__global__ void __launch_bounds__(KB_WARPS_PER_BLOCK * 32, BPS)
kb_traverse_kernel(const KbTraverseParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int xf_floats = (p.nxf * 12 + 3) & ~3;
  const int nit_c = ITC ? p.nitems : 0;
  ItemS* s_items = (ItemS*)smem_raw;
  const int mask_words = p.nprobes > 0 ? (((p.nitems + 31) >> 5) + 3) & ~3 : 0;
  const int nprobes_s = p.nprobes <= KB_PROBES_SMEM_MAX ? p.nprobes : 0;      // probes cached per CTA
  const size_t per_warp = (size_t)KB_STACK_CAP * 8 + (size_t)KB_BOOL_LEAFQ_CAP * 8 + (size_t)KB_RQ_CAP * 16 + 16 + (size_t)xf_floats * 4 + (size_t)nit_c * 48 + (size_t)mask_words * 4;
  KbProbe* s_probes = (KbProbe*)(smem_raw + (size_t)nit_c * 16);
  unsigned char* base = smem_raw + (size_t)nit_c * 16 + (size_t)nprobes_s * 32 + warp * per_warp;
  uint2* stack = (uint2*)base;
  uint2* leafq = (uint2*)(base + (size_t)KB_STACK_CAP * 8);
  uint4* rq = (uint4*)(base + (size_t)KB_STACK_CAP * 8 + (size_t)KB_BOOL_LEAFQ_CAP * 8);
  int* rq_count = (int*)(rq + KB_RQ_CAP);
  float* xfw = (float*)(base + (size_t)KB_STACK_CAP * 8 + (size_t)KB_BOOL_LEAFQ_CAP * 8 + (size_t)KB_RQ_CAP * 16 + 16);
  float* itc = xfw + xf_floats;
  unsigned* amask = (unsigned*)(itc + (size_t)nit_c * 12);   // per configuration: bit i = item i must be traversed
  const KbScene& sc = p.scene;
  const float slack = 4.f * sc.eps_abs;
  if (ITC) {
    for (int i = threadIdx.x; i < p.nitems; i += blockDim.x) {
      const KbItem* it = p.items + i;
      ItemS s; s.nodeA = it->nodeA; s.nodeB = it->nodeB; s.infl = (float)it->thr + slack;
      s.xf = (int)((unsigned)(unsigned short)it->xfA | ((unsigned)(unsigned short)it->xfB << 16));
      s_items[i] = s;
    }
  }
  for (int i = threadIdx.x; i < nprobes_s * 2; i += blockDim.x) ((uint4*)s_probes)[i] = __ldg((const uint4*)p.probes + i);
  if (lane == 0) *rq_count = 0;
  __syncthreads();
  unsigned lt_mask;
  asm("mov.u32 %0, %%lanemask_lt;" : "=r"(lt_mask));
  unsigned st_node = 0, st_leaf = 0, st_re = 0, st_drop = 0, st_iter = 0;   // only maintained when STATS
  const bool use_probes = p.nprobes > 0;
  const KbProbe* probes = nprobes_s ? s_probes : p.probes;

  // guided self-scheduling: 8 configurations per grab while work is plentiful, down to 1 near the end of the launch, so
  // the tail is one configuration long (configuration cost varies by two orders of magnitude)
  const unsigned total_warps = gridDim.x * KB_WARPS_PER_BLOCK;
  unsigned grab;                      // first grab by the same rule as the later ones: a mid-size batch must not sit on an eighth of the warps
  { const unsigned g0 = (unsigned)p.N / (4u * total_warps); grab = g0 >= 8u ? 8u : (g0 < 1u ? 1u : g0); }
  bool static_done = false;
  for (;;) {
    unsigned int c0 = 0;
    if (STATIC) {                     // one configuration per warp, by warp index (small batches; a compile-time variant: the
                                      // batch kernel's register allocation stays what it was)
      if (static_done) break;
      static_done = true; grab = 1;
      c0 = blockIdx.x * KB_WARPS_PER_BLOCK + warp;
    } else {
      if (lane == 0) c0 = atomicAdd(p.work_counter, grab);
      c0 = __shfl_sync(FULL, c0, 0);
    }
    if ((int64_t)c0 >= p.N) break;
    const unsigned nN = (unsigned)p.N;
    const unsigned cend = (c0 + grab < nN) ? c0 + grab : nN;
    {
      const unsigned g = (nN - cend) / (4u * total_warps);
      grab = g >= 8u ? 8u : (g < 1u ? 1u : g);
    }
    for (unsigned c = c0; c < cend; c++) {
      if (p.state && p.state[c] == 0) continue;
      const double* xf = p.xf64 + (size_t)c * (size_t)p.nxf * 12;
      __syncwarp();
      // the transforms of the next configuration of this grab come from DRAM (FK wrote 96 L bytes x 1 M configurations): start them
      // towards L2 now, one 128-byte line per lane.  Pays for long rows only (19 links: 7.83 -> 7.59 ms; 7 links: 6.22 -> 6.29).
      if (p.nxf >= 12 && c + 1 < cend && lane * 16 < p.nxf * 12) asm volatile("prefetch.global.L2 [%0];" :: "l"(xf + (size_t)p.nxf * 12 + lane * 16));
      for (int i = lane; i < p.nxf * 12; i += 32) xfw[i] = (float)xf[i];
      __syncwarp();
      if (ITC) {
        for (int i = lane; i < p.nitems; i += 32) {
          const int xfp = s_items[i].xf;
          XfF T; rel_xf(xfw, (int)(short)(xfp & 0xffff), (int)(short)(xfp >> 16), T);
          float4* q = (float4*)(itc + 12 * i);
          q[0] = make_float4(T.r[0], T.r[1], T.r[2], T.r[3]); q[1] = make_float4(T.r[4], T.r[5], T.r[6], T.r[7]); q[2] = make_float4(T.r[8], T.t[0], T.t[1], T.t[2]);
        }
        __syncwarp();
      }
      if (use_probes) {
        // clearance-grid broad phase: an item whose covering spheres all have more clearance from the static group than
        // radius + threshold cannot collide and is never fed to the traversal
        for (int w = lane; w < ((p.nitems + 31) >> 5); w += 32) amask[w] = __ldg(p.always_on + w);
        __syncwarp();
        // four probes per lane per round: all grid bytes of a round are in flight together (one L2 round trip per 128 probes)
        for (int s0 = 0; s0 < p.nprobes; s0 += 128) {
          unsigned q[4], need[4]; int item[4];
#pragma unroll
          for (int u = 0; u < 4; u++) {
            const int s = s0 + 32 * u + lane;
            q[u] = 0xffffffffu; need[u] = 0u; item[u] = 0;
            if (s < p.nprobes) {
              const float4 pc = *(const float4*)(probes + s);
              const int4 pi = *((const int4*)(probes + s) + 1);
              const float* T = xfw + 12 * pi.y;
              const float wx = T[0] * pc.x + T[1] * pc.y + T[2] * pc.z + T[9];
              const float wy = T[3] * pc.x + T[4] * pc.y + T[5] * pc.z + T[10];
              const float wz = T[6] * pc.x + T[7] * pc.y + T[8] * pc.z + T[11];
              const KbClearGrid& g = sc.grids[pi.z];
              // points outside the grid are clamped onto it: the projection onto a convex box never increases the distance
              // to anything inside the box, and everything the grid measures lies inside
              int ix = __float2int_rd((wx - g.o[0]) * g.inv_h), iy = __float2int_rd((wy - g.o[1]) * g.inv_h), iz = __float2int_rd((wz - g.o[2]) * g.inv_h);
              ix = min(max(ix, 0), g.dims[0] - 1); iy = min(max(iy, 0), g.dims[1] - 1); iz = min(max(iz, 0), g.dims[2] - 1);
              q[u] = __ldg(g.data + ((size_t)iz * g.dims[1] + iy) * g.dims[0] + ix);
              need[u] = __float_as_uint(pc.w); item[u] = pi.x;
            }
          }
#pragma unroll
          for (int u = 0; u < 4; u++) if (q[u] < need[u]) atomicOr(amask + (item[u] >> 5), 1u << (item[u] & 31));
        }
        __syncwarp();
      }
      int sp = 0, nleaf = 0, cursor = 0;
      int leaf_trig = KB_LEAF_FIRST;                 // the first element phase of a configuration runs on a smaller batch
      int found = -1, found_ea = -1, found_eb = -1;
      for (;;) {
        if (sp < 32 && cursor < p.nitems) {          // feed root pairs of the next work items
          int k = p.nitems - cursor; if (k > 32) k = 32;
          if (use_probes) {
            const bool on = lane < k && ((amask[(cursor + lane) >> 5] >> ((cursor + lane) & 31)) & 1u);
            const unsigned om = __ballot_sync(FULL, on);
            if (on) stack[sp + __popc(om & lt_mask)] = make_uint2(((unsigned)(cursor + lane) << KB_NODEA_BITS), 0u);
            sp += __popc(om);
            if (STATS) st_drop += (lane == 0) ? (unsigned)(k - __popc(om)) : 0u;
          } else {
            if (lane < k) stack[sp + lane] = make_uint2(((unsigned)(cursor + lane) << KB_NODEA_BITS), 0u);
            sp += k;
          }
          cursor += k;
          __syncwarp();
        }
        if (sp == 0 && nleaf == 0) break;
        if (nleaf >= leaf_trig || sp == 0) {
          leaf_trig = KB_LEAF_TRIGGER;
          // ---------------------------------------------------------------- element phase
          int m = nleaf < 32 ? nleaf : 32;
          int res = KB_NO, ea = -1, eb = -1, item = 0;
          if (lane < m) {
            uint2 e = leafq[nleaf - 1 - lane];
            item = (int)(e.x >> KB_NODEA_BITS);
            const KbItem& it = p.items[item];
            int na = (int)(e.x & (KB_MAX_NODES_A - 1)), nb = (int)e.y;
            float4 a0, a1, b0, b1;
            load_node(sc.nodes, (size_t)(it.nodeA + na), a0, a1);
            load_node(sc.nodes, (size_t)(it.nodeB + nb), b0, b1);
            int fa = it.elemA + ~__float_as_int(a0.w), ca = __float_as_int(a1.w);
            int fb = it.elemB + ~__float_as_int(b0.w), cb = __float_as_int(b1.w);
            {
              XfF T;
              if (ITC) load_itc(itc, item, T); else rel_xf(xfw, it.xfA, it.xfB, T);
              const float thr = (float)it.thr;
              for (int i = 0; i < ca && res != KB_YES; i++)
                for (int j = 0; j < cb && res != KB_YES; j++) {
                  int r = fast_elem_collide<BOXES>(sc, it, T, fa + i, fb + j, thr);
                  if (STATS) st_leaf++;
                  if (r == KB_UNCERTAIN) {
                    if (STATS) st_re++;
                    const int slot = atomicAdd(rq_count, 1);
                    if (slot < KB_RQ_CAP) { rq[slot] = make_uint4((unsigned)c, (unsigned)item, (unsigned)(fa + i), (unsigned)(fb + j)); r = KB_NO; }
                    else r = exact_elem_collide<BOXES>(sc, it, xf, fa + i, fb + j) ? KB_YES : KB_NO;   // queue full: recheck in place
                  }
                  if (r == KB_YES) { res = KB_YES; ea = fa + i; eb = fb + j; }
                }
            }
          }
          nleaf -= m;
          {
            unsigned hm = __ballot_sync(FULL, res == KB_YES);
            if (hm) {
              int src = __ffs(hm) - 1;
              found = __shfl_sync(FULL, item, src); found_ea = __shfl_sync(FULL, ea, src); found_eb = __shfl_sync(FULL, eb, src);
              break;
            }
            drain_rechecks<BOXES>(p, rq, rq_count, lane, (int64_t)c, found, found_ea, found_eb, false);
            if (found >= 0) break;
          }
          __syncwarp();
          continue;
        }
        // ------------------------------------------------------------------ node phase: a tight inner loop that runs until leaf pairs
        // are due, the stack runs dry or new root pairs must be fed (keeps the loop state in registers)
        {
        // loop state in fresh variables: the allocator keeps them in registers for the loop instead of in the spill slots the
        // element phase forces on the outer copies
        int sp_l = sp, nleaf_l = nleaf;
        const float* itc_l = itc;
        const bool more_items = cursor < p.nitems;
        const int pop_room = p.pop_room;
        const int leaf_trig_l = leaf_trig;
        const int lane_l = lane;
        const float both_ratio = p.both_ratio;
        do {
        int m = sp_l < KB_POP_WIDTH ? sp_l : KB_POP_WIDTH;
        { const int room = (pop_room - sp_l) / 3; m = m < room ? m : (room > 1 ? room : 1); }     // narrower pops as the stack fills up
        const bool act = lane_l < m;
        uint2 e = make_uint2(0u, 0u);
        if (act) e = stack[sp_l - 1 - lane_l];
        sp_l -= m;
        __syncwarp();
#if KB_BOTH_MODE == 1 && !defined(KB_BRANCHY_NODE_TEST)
        // Branch-free form: inactive lanes test entry (item 0, root, root) and discard the result, and the child entries are
        // selected arithmetically -- no divergent regions inside the loop body.
        const int item = (int)(e.x >> KB_NODEA_BITS);
        int nodeA, nodeB; float infl; XfF T;
        if (ITC) { const ItemS si = s_items[item]; nodeA = si.nodeA; nodeB = si.nodeB; infl = si.infl; load_itc(itc_l, item, T); }
        else { const KbItem* itp = p.items + item; nodeA = itp->nodeA; nodeB = itp->nodeB; infl = (float)itp->thr + slack; rel_xf(xfw, itp->xfA, itp->xfB, T); }
        float4 a0, a1, b0, b1;
        load_node(sc.nodes, (size_t)(nodeA + (int)(e.x & (KB_MAX_NODES_A - 1))), a0, a1);
        load_node(sc.nodes, (size_t)(nodeB + (int)e.y), b0, b1);
        const bool ov = act & sat6_overlap(a0, a1, b0, b1, T, infl);
        if (STATS) st_node += act;
        if (STATS) st_iter += (lane_l == 0);
        const int la = __float_as_int(a0.w), lb = __float_as_int(b0.w);
        const bool leafpair = ov & ((la & lb) < 0);
        const bool inner = ov & ((la & lb) >= 0);
        const float sa2 = a1.x * a1.x + a1.y * a1.y + a1.z * a1.z, sb2 = b1.x * b1.x + b1.y * b1.y + b1.z * b1.z;
        const bool both = inner & ((la | lb) >= 0) & (sa2 < both_ratio * sb2) & (sb2 < both_ratio * sa2);
        const bool splitA = (lb < 0) | ((la >= 0) & (sa2 >= sb2));          // only meaningful when inner && !both
        const bool useA = both | splitA, useB = both | !splitA;
        const unsigned ax = useA ? ((e.x & ~(unsigned)(KB_MAX_NODES_A - 1)) | (unsigned)la) : e.x;
        const unsigned by = useB ? (unsigned)lb : e.y;
        const unsigned pm = __ballot_sync(FULL, inner), lm = __ballot_sync(FULL, leafpair), pm4 = __ballot_sync(FULL, both);
        if (inner) {
          const int off = sp_l + 2 * __popc(pm & lt_mask) + 2 * __popc(pm4 & lt_mask);
          stack[off] = useA ? make_uint2(ax + 1u, by) : make_uint2(ax, by + 1u);
          stack[off + 1] = make_uint2(ax, by);
          if (both) { stack[off + 2] = make_uint2(ax + 1u, by + 1u); stack[off + 3] = make_uint2(ax, by + 1u); }
        }
        if (leafpair) leafq[nleaf_l + __popc(lm & lt_mask)] = e;
        sp_l += 2 * __popc(pm) + 2 * __popc(pm4); nleaf_l += __popc(lm);
#else
        bool push2 = false, push4 = false, leafpair = false;
        uint2 c0e = e, c1e = e;
        if (act) {
          node_test<ITC>(p, s_items, itc_l, xfw, slack, e, KB_BOTH_MODE == 1 ? true : sp_l + m <= p.both_limit, push2, push4, leafpair, c0e, c1e);
          if (STATS) st_node++;
        }
        if (STATS) st_iter += (lane_l == 0);
#if KB_BOTH_MODE
        const unsigned pm = __ballot_sync(FULL, push2), lm = __ballot_sync(FULL, leafpair), pm4 = __ballot_sync(FULL, push4);
        if (push2) {
          int off = sp_l + 2 * __popc(pm & lt_mask) + 2 * __popc(pm4 & lt_mask); stack[off] = c1e; stack[off + 1] = c0e;
          if (push4) { stack[off + 2] = make_uint2(c1e.x, c1e.y + 1u); stack[off + 3] = make_uint2(c0e.x, c0e.y + 1u); }
        }
        if (leafpair) leafq[nleaf_l + __popc(lm & lt_mask)] = e;
        sp_l += 2 * __popc(pm) + 2 * __popc(pm4); nleaf_l += __popc(lm);
#else
        const unsigned pm = __ballot_sync(FULL, push2), lm = __ballot_sync(FULL, leafpair);
        if (push2) { int off = sp_l + 2 * __popc(pm & lt_mask); stack[off] = c1e; stack[off + 1] = c0e; }
        if (leafpair) leafq[nleaf_l + __popc(lm & lt_mask)] = e;
        sp_l += 2 * __popc(pm); nleaf_l += __popc(lm);
#endif
#endif
        __syncwarp();
        } while (sp_l > 0 && nleaf_l < leaf_trig_l && !(sp_l < 32 && more_items));
        sp = sp_l; nleaf = nleaf_l;
        }
      }
      if (STATS) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          st_node += __shfl_xor_sync(FULL, st_node, o); st_leaf += __shfl_xor_sync(FULL, st_leaf, o); st_re += __shfl_xor_sync(FULL, st_re, o);
        }
        if (lane == 0 && p.counters) {
          atomicAdd(p.counters + 0, (unsigned long long)st_re); atomicAdd(p.counters + 1, (unsigned long long)st_node); atomicAdd(p.counters + 2, (unsigned long long)st_leaf);
          atomicAdd(p.counters + 7, (unsigned long long)st_drop); atomicAdd(p.counters + 8, (unsigned long long)st_iter);
        }
        st_node = st_leaf = st_re = st_drop = st_iter = 0;
      }
      if (lane == 0) {
        p.hit[c] = found;
        if (p.hit_elem) { p.hit_elem[2 * (size_t)c] = found_ea; p.hit_elem[2 * (size_t)c + 1] = found_eb; }
      }
    }
  }
  // pairs still parked for the fp64 recheck belong to configurations already written as "no hit": resolve them now
  { int f = 0, fa = 0, fb = 0; drain_rechecks<BOXES>(p, rq, rq_count, lane, (int64_t)-1, f, fa, fb, true); }
  if (STATIC) {
    // the result byte of this warp's configuration, after every pair it parked for the fp64 recheck was resolved
    const int64_t c = (int64_t)blockIdx.x * KB_WARPS_PER_BLOCK + warp;
    if (c < p.N && lane == 0) {
      const bool feas = (!p.state || p.state[c] != 0) && p.hit[c] < 0;
      p.out_bytes[c] = feas ? 1 : 0;
      if (feas && p.nfeasible) atomicAdd(p.nfeasible, 1ull);
    }
  }
}

// =============================================================================================== traversal (boolean, 4-wide hierarchies)
// kb_traverse_wide_kernel -- kb_traverse_kernel on the 4-wide form of the hierarchies (KbScene::wide).  Same warp-per-configuration
// frontier, same element phase, same deferred fp64 rechecks; what changes is the node step.  The binary kernel is bound by the number
// of distinct 128-byte lines its two node loads touch per iteration (profiles/r02_experiments.md): 32 lanes test 32 pairs and fetch
// 8-16+ lines per load.  Here a popped entry is (item, wide node of the side to expand, slot of the other side's box) and FOUR lanes
// test the node's four child slots against that box: one line on the expanded side, one 32-byte slot on the other, per four tests.
// The side to expand is always the larger box (decided when the entry is pushed, where both boxes are in registers), i.e. the
// descend-larger rule two levels at a time -- no descend-both step, whose extra tests cost C2 more than they saved.
template <bool ITC, bool STATS, int BPS, bool BOXES, bool STATIC>
__global__ void __launch_bounds__(KB_WARPS_PER_BLOCK * 32, BPS)
kb_traverse_wide_kernel(const KbTraverseParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int xf_floats = (p.nxf * 12 + 3) & ~3;
  const int nit_c = ITC ? p.nitems : 0;
  ItemS* s_items = (ItemS*)smem_raw;
  const int mask_words = p.nprobes > 0 ? (((p.nitems + 31) >> 5) + 3) & ~3 : 0;
  const int nprobes_s = p.nprobes <= KB_PROBES_SMEM_MAX ? p.nprobes : 0;      // probes cached per CTA
  const size_t per_warp = (size_t)KB_STACK_CAP * 8 + (size_t)KB_BOOL_LEAFQ_CAP * 8 + (size_t)KB_RQ_CAP * 16 + 16 + (size_t)xf_floats * 4 + (size_t)nit_c * 48 + (size_t)mask_words * 4;
  KbProbe* s_probes = (KbProbe*)(smem_raw + (size_t)nit_c * 16);
  unsigned char* base = smem_raw + (size_t)nit_c * 16 + (size_t)nprobes_s * 32 + warp * per_warp;
  uint2* stack = (uint2*)base;
  uint2* leafq = (uint2*)(base + (size_t)KB_STACK_CAP * 8);
  uint4* rq = (uint4*)(base + (size_t)KB_STACK_CAP * 8 + (size_t)KB_BOOL_LEAFQ_CAP * 8);
  int* rq_count = (int*)(rq + KB_RQ_CAP);
  float* xfw = (float*)(base + (size_t)KB_STACK_CAP * 8 + (size_t)KB_BOOL_LEAFQ_CAP * 8 + (size_t)KB_RQ_CAP * 16 + 16);
  float* itc = xfw + xf_floats;
  unsigned* amask = (unsigned*)(itc + (size_t)nit_c * 12);   // per configuration: bit i = item i must be traversed
  const KbScene& sc = p.scene;
  const float slack = 4.f * sc.eps_abs;
  if (ITC) {
    for (int i = threadIdx.x; i < p.nitems; i += blockDim.x) {
      const KbItem* it = p.items + i;
      ItemS s; s.nodeA = it->wideA; s.nodeB = it->wideB; s.infl = (float)it->thr + slack;      // slot bases of the two 4-wide hierarchies
      s.xf = (int)((unsigned)(unsigned short)it->xfA | ((unsigned)(unsigned short)it->xfB << 16));
      s_items[i] = s;
    }
  }
  for (int i = threadIdx.x; i < nprobes_s * 2; i += blockDim.x) ((uint4*)s_probes)[i] = __ldg((const uint4*)p.probes + i);
  if (lane == 0) *rq_count = 0;
  __syncthreads();
  unsigned lt_mask;
  asm("mov.u32 %0, %%lanemask_lt;" : "=r"(lt_mask));
  unsigned st_node = 0, st_leaf = 0, st_re = 0, st_drop = 0, st_iter = 0;   // only maintained when STATS
  const bool use_probes = p.nprobes > 0;
  const KbProbe* probes = nprobes_s ? s_probes : p.probes;

  // guided self-scheduling: 8 configurations per grab while work is plentiful, down to 1 near the end of the launch, so
  // the tail is one configuration long (configuration cost varies by two orders of magnitude)
  const unsigned total_warps = gridDim.x * KB_WARPS_PER_BLOCK;
  unsigned grab;                      // first grab by the same rule as the later ones: a mid-size batch must not sit on an eighth of the warps
  { const unsigned g0 = (unsigned)p.N / (4u * total_warps); grab = g0 >= 8u ? 8u : (g0 < 1u ? 1u : g0); }
  bool static_done = false;
  for (;;) {
    unsigned int c0 = 0;
    if (STATIC) {                     // one configuration per warp, by warp index (small batches; a compile-time variant: the
                                      // batch kernel's register allocation stays what it was)
      if (static_done) break;
      static_done = true; grab = 1;
      c0 = blockIdx.x * KB_WARPS_PER_BLOCK + warp;
    } else {
      if (lane == 0) c0 = atomicAdd(p.work_counter, grab);
      c0 = __shfl_sync(FULL, c0, 0);
    }
    if ((int64_t)c0 >= p.N) break;
    const unsigned nN = (unsigned)p.N;
    const unsigned cend = (c0 + grab < nN) ? c0 + grab : nN;
    {
      const unsigned g = (nN - cend) / (4u * total_warps);
      grab = g >= 8u ? 8u : (g < 1u ? 1u : g);
    }
    for (unsigned c = c0; c < cend; c++) {
      if (p.state && p.state[c] == 0) continue;
      const double* xf = p.xf64 + (size_t)c * (size_t)p.nxf * 12;
      __syncwarp();
      // the transforms of the next configuration of this grab come from DRAM (FK wrote 96 L bytes x 1 M configurations): start them
      // towards L2 now, one 128-byte line per lane.  Pays for long rows only (19 links: 7.83 -> 7.59 ms; 7 links: 6.22 -> 6.29).
      if (p.nxf >= 12 && c + 1 < cend && lane * 16 < p.nxf * 12) asm volatile("prefetch.global.L2 [%0];" :: "l"(xf + (size_t)p.nxf * 12 + lane * 16));
      for (int i = lane; i < p.nxf * 12; i += 32) xfw[i] = (float)xf[i];
      __syncwarp();
      if (ITC) {
        for (int i = lane; i < p.nitems; i += 32) {
          const int xfp = s_items[i].xf;
          XfF T; rel_xf(xfw, (int)(short)(xfp & 0xffff), (int)(short)(xfp >> 16), T);
          float4* q = (float4*)(itc + 12 * i);
          q[0] = make_float4(T.r[0], T.r[1], T.r[2], T.r[3]); q[1] = make_float4(T.r[4], T.r[5], T.r[6], T.r[7]); q[2] = make_float4(T.r[8], T.t[0], T.t[1], T.t[2]);
        }
        __syncwarp();
      }
      if (use_probes) {
        // clearance-grid broad phase: an item whose covering spheres all have more clearance from the static group than
        // radius + threshold cannot collide and is never fed to the traversal
        for (int w = lane; w < ((p.nitems + 31) >> 5); w += 32) amask[w] = __ldg(p.always_on + w);
        __syncwarp();
        // four probes per lane per round: all grid bytes of a round are in flight together (one L2 round trip per 128 probes)
        for (int s0 = 0; s0 < p.nprobes; s0 += 128) {
          unsigned q[4], need[4]; int item[4];
#pragma unroll
          for (int u = 0; u < 4; u++) {
            const int s = s0 + 32 * u + lane;
            q[u] = 0xffffffffu; need[u] = 0u; item[u] = 0;
            if (s < p.nprobes) {
              const float4 pc = *(const float4*)(probes + s);
              const int4 pi = *((const int4*)(probes + s) + 1);
              const float* T = xfw + 12 * pi.y;
              const float wx = T[0] * pc.x + T[1] * pc.y + T[2] * pc.z + T[9];
              const float wy = T[3] * pc.x + T[4] * pc.y + T[5] * pc.z + T[10];
              const float wz = T[6] * pc.x + T[7] * pc.y + T[8] * pc.z + T[11];
              const KbClearGrid& g = sc.grids[pi.z];
              // points outside the grid are clamped onto it: the projection onto a convex box never increases the distance
              // to anything inside the box, and everything the grid measures lies inside
              int ix = __float2int_rd((wx - g.o[0]) * g.inv_h), iy = __float2int_rd((wy - g.o[1]) * g.inv_h), iz = __float2int_rd((wz - g.o[2]) * g.inv_h);
              ix = min(max(ix, 0), g.dims[0] - 1); iy = min(max(iy, 0), g.dims[1] - 1); iz = min(max(iz, 0), g.dims[2] - 1);
              q[u] = __ldg(g.data + ((size_t)iz * g.dims[1] + iy) * g.dims[0] + ix);
              need[u] = __float_as_uint(pc.w); item[u] = pi.x;
            }
          }
#pragma unroll
          for (int u = 0; u < 4; u++) if (q[u] < need[u]) atomicOr(amask + (item[u] >> 5), 1u << (item[u] & 31));
        }
        __syncwarp();
      }
      int sp = 0, nleaf = 0, cursor = 0;
      int leaf_trig = KB_LEAF_FIRST;                 // the first element phase of a configuration runs on a smaller batch
      int found = -1, found_ea = -1, found_eb = -1;
      for (;;) {
        if (sp < 32 && nleaf <= KB_BOOL_LEAFQ_CAP - 32 && cursor < p.nitems) {
          // root pairs of the next work items, one lane each: the two root boxes (slot 0 of each super root) are tested right here and
          // only the pairs that overlap enter the frontier, already knowing which side to expand
          int kf = p.nitems - cursor; if (kf > 32) kf = 32;
          const bool act = lane < kf;
          const int item = act ? cursor + lane : 0;
          int baseA, baseB; float infl; XfF T;
          if (ITC) { const ItemS si = s_items[item]; baseA = si.nodeA; baseB = si.nodeB; infl = si.infl; load_itc(itc, item, T); }
          else { const KbItem* itp = p.items + item; baseA = itp->wideA; baseB = itp->wideB; infl = (float)itp->thr + slack; rel_xf(xfw, itp->xfA, itp->xfB, T); }
          float4 a0, a1, b0, b1;
          load_node(sc.wide, (size_t)baseA, a0, a1);
          load_node(sc.wide, (size_t)baseB, b0, b1);
          const bool ov = act & sat6_overlap(a0, a1, b0, b1, T, infl);
          if (STATS) st_node += act;
          if (STATS) st_iter += (lane == 0);
          const int la = __float_as_int(a0.w), lb = __float_as_int(b0.w);
          const bool leafpair = ov & ((la & lb) < 0);
          const bool inner = ov & ((la & lb) >= 0);
          const float sa2 = a1.x * a1.x + a1.y * a1.y + a1.z * a1.z, sb2 = b1.x * b1.x + b1.y * b1.y + b1.z * b1.z;
          const bool nextA = (lb < 0) | ((la >= 0) & (sa2 >= sb2));
          const unsigned itembits = (unsigned)item << KB_NODEA_BITS;
          const unsigned pm = __ballot_sync(FULL, inner), lm = __ballot_sync(FULL, leafpair);
          if (inner) stack[sp + __popc(pm & lt_mask)] = make_uint2(itembits | (unsigned)(nextA ? 4 * la : 0), nextA ? 0u : (0x80000000u | (unsigned)(4 * lb)));
          if (leafpair) leafq[nleaf + __popc(lm & lt_mask)] = make_uint2(itembits, 0u);
          sp += __popc(pm); nleaf += __popc(lm);
          cursor += kf;
          __syncwarp();
        }
        if (sp == 0 && nleaf == 0) { if (cursor < p.nitems) continue; break; }
        if (nleaf >= leaf_trig || sp == 0) {
          leaf_trig = KB_LEAF_TRIGGER;
          // ---------------------------------------------------------------- element phase
          int m = nleaf < 32 ? nleaf : 32;
          int res = KB_NO, ea = -1, eb = -1, item = 0;
          if (lane < m) {
            uint2 e = leafq[nleaf - 1 - lane];
            item = (int)(e.x >> KB_NODEA_BITS);
            const KbItem& it = p.items[item];
            int na = (int)(e.x & (KB_MAX_NODES_A - 1)), nb = (int)e.y;
            float4 a0, a1, b0, b1;
            load_node(sc.wide, (size_t)(it.wideA + na), a0, a1);
            load_node(sc.wide, (size_t)(it.wideB + nb), b0, b1);
            int fa = it.elemA + ~__float_as_int(a0.w), ca = __float_as_int(a1.w);
            int fb = it.elemB + ~__float_as_int(b0.w), cb = __float_as_int(b1.w);
            {
              XfF T;
              if (ITC) load_itc(itc, item, T); else rel_xf(xfw, it.xfA, it.xfB, T);
              const float thr = (float)it.thr;
              for (int i = 0; i < ca && res != KB_YES; i++)
                for (int j = 0; j < cb && res != KB_YES; j++) {
                  int r = fast_elem_collide<BOXES>(sc, it, T, fa + i, fb + j, thr);
                  if (STATS) st_leaf++;
                  if (r == KB_UNCERTAIN) {
                    if (STATS) st_re++;
                    const int slot = atomicAdd(rq_count, 1);
                    if (slot < KB_RQ_CAP) { rq[slot] = make_uint4((unsigned)c, (unsigned)item, (unsigned)(fa + i), (unsigned)(fb + j)); r = KB_NO; }
                    else r = exact_elem_collide<BOXES>(sc, it, xf, fa + i, fb + j) ? KB_YES : KB_NO;   // queue full: recheck in place
                  }
                  if (r == KB_YES) { res = KB_YES; ea = fa + i; eb = fb + j; }
                }
            }
          }
          nleaf -= m;
          {
            unsigned hm = __ballot_sync(FULL, res == KB_YES);
            if (hm) {
              int src = __ffs(hm) - 1;
              found = __shfl_sync(FULL, item, src); found_ea = __shfl_sync(FULL, ea, src); found_eb = __shfl_sync(FULL, eb, src);
              break;
            }
            drain_rechecks<BOXES>(p, rq, rq_count, lane, (int64_t)c, found, found_ea, found_eb, false);
            if (found >= 0) break;
          }
          __syncwarp();
          continue;
        }
        // ------------------------------------------------------------------ node phase: a tight inner loop that runs until leaf pairs
        // are due, the stack runs dry or new root pairs must be fed (keeps the loop state in registers)
        {
        // loop state in fresh variables: the allocator keeps them in registers for the loop instead of in the spill slots the
        // element phase forces on the outer copies
        int sp_l = sp, nleaf_l = nleaf;
        const float* itc_l = itc;
        const bool more_items = cursor < p.nitems;
        const int leaf_trig_l = leaf_trig;
        const int lane_l = lane;
        const int wide_room = p.wide_room;
        do {
        // Eight entries per iteration, four lanes each.  An entry = (item, side to expand, the wide node of that side, the slot of the
        // other side's box); lane t of a group tests child slot t of the wide node against that box: the four lanes read ONE 128-byte
        // line on the expanded side and the same 32 bytes on the other.
        int m = sp_l < 8 ? sp_l : 8;
        if (sp_l > wide_room) m = 1;                                  // nearly full: depth-first, one entry at a time
        const int g = lane_l >> 2, t = lane_l & 3;
        const bool act = g < m;
        uint2 e = make_uint2(0u, 0u);
        if (act) e = stack[sp_l - 1 - g];
        sp_l -= m;
        __syncwarp();
        const int item = (int)(e.x >> KB_NODEA_BITS);
        int baseA, baseB; float infl; XfF T;
        if (ITC) { const ItemS si = s_items[item]; baseA = si.nodeA; baseB = si.nodeB; infl = si.infl; load_itc(itc_l, item, T); }
        else { const KbItem* itp = p.items + item; baseA = itp->wideA; baseB = itp->wideB; infl = (float)itp->thr + slack; rel_xf(xfw, itp->xfA, itp->xfB, T); }
        const bool expB = (e.y >> 31) != 0;
        const int sa = (int)(e.x & (KB_MAX_NODES_A - 1)) + (expB ? 0 : t), sb = (int)(e.y & 0x7fffffffu) + (expB ? t : 0);
        float4 a0, a1, b0, b1;
        load_node(sc.wide, (size_t)(baseA + sa), a0, a1);
        load_node(sc.wide, (size_t)(baseB + sb), b0, b1);
        const bool ov = act & sat6_overlap(a0, a1, b0, b1, T, infl);
        if (STATS) st_node += act;
        if (STATS) st_iter += (lane_l == 0);
        const int la = __float_as_int(a0.w), lb = __float_as_int(b0.w);
        const bool leafpair = ov & ((la & lb) < 0);
        const bool inner = ov & ((la & lb) >= 0);
        const float sa2 = a1.x * a1.x + a1.y * a1.y + a1.z * a1.z, sb2 = b1.x * b1.x + b1.y * b1.y + b1.z * b1.z;
        const bool nextA = (lb < 0) | ((la >= 0) & (sa2 >= sb2));          // expand the larger box next (the only inner one if the other is a leaf)
        const unsigned ex = (e.x & ~(unsigned)(KB_MAX_NODES_A - 1)) | (unsigned)(nextA ? 4 * la : sa);
        const unsigned ey = nextA ? (unsigned)sb : (0x80000000u | (unsigned)(4 * lb));
        const unsigned pm = __ballot_sync(FULL, inner), lm = __ballot_sync(FULL, leafpair);
        if (inner) stack[sp_l + __popc(pm & lt_mask)] = make_uint2(ex, ey);
        if (leafpair) leafq[nleaf_l + __popc(lm & lt_mask)] = make_uint2((e.x & ~(unsigned)(KB_MAX_NODES_A - 1)) | (unsigned)sa, (unsigned)sb);
        sp_l += __popc(pm); nleaf_l += __popc(lm);
        __syncwarp();
        } while (sp_l > 0 && nleaf_l < leaf_trig_l && !(sp_l < 32 && more_items));
        sp = sp_l; nleaf = nleaf_l;
        }
      }
      if (STATS) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          st_node += __shfl_xor_sync(FULL, st_node, o); st_leaf += __shfl_xor_sync(FULL, st_leaf, o); st_re += __shfl_xor_sync(FULL, st_re, o);
        }
        if (lane == 0 && p.counters) {
          atomicAdd(p.counters + 0, (unsigned long long)st_re); atomicAdd(p.counters + 1, (unsigned long long)st_node); atomicAdd(p.counters + 2, (unsigned long long)st_leaf);
          atomicAdd(p.counters + 7, (unsigned long long)st_drop); atomicAdd(p.counters + 8, (unsigned long long)st_iter);
        }
        st_node = st_leaf = st_re = st_drop = st_iter = 0;
      }
      if (lane == 0) {
        p.hit[c] = found;
        if (p.hit_elem) { p.hit_elem[2 * (size_t)c] = found_ea; p.hit_elem[2 * (size_t)c + 1] = found_eb; }
      }
    }
  }
  // pairs still parked for the fp64 recheck belong to configurations already written as "no hit": resolve them now
  { int f = 0, fa = 0, fb = 0; drain_rechecks<BOXES>(p, rq, rq_count, lane, (int64_t)-1, f, fa, fb, true); }
  if (STATIC) {
    // the result byte of this warp's configuration, after every pair it parked for the fp64 recheck was resolved
    const int64_t c = (int64_t)blockIdx.x * KB_WARPS_PER_BLOCK + warp;
    if (c < p.N && lane == 0) {
      const bool feas = (!p.state || p.state[c] != 0) && p.hit[c] < 0;
      p.out_bytes[c] = feas ? 1 : 0;
      if (feas && p.nfeasible) atomicAdd(p.nfeasible, 1ull);
    }
  }
}

// =============================================================================================== all colliding pairs
// kb_allpairs_kernel -- no early exit: lists every colliding (idA, idB) world-id pair of a configuration, up to max_pairs.
// Replaces evaluating SingleRobotCSpace's per-pair CollisionFreeSet constraints one by one (reference
// Cpp/Planning/RobotCSpace.cpp:697-747; CSpaceInterface::feasibilityFailures, Python/klampt/src/motionplanning.h:122-171).
// Same warp-cooperative node traversal; leaf pairs whose id pair is already listed are skipped, and items that map to a
// single id pair (link vs link, link vs a one-object group) stop traversing once found.  Uncertain fp32 results are
// re-run in fp64 in place (this is a diagnostic query, not the hot loop).
#define KB_AP_MAX 32
template <bool ITC>
__global__ void __launch_bounds__(KB_WARPS_PER_BLOCK * 32, 3)
kb_allpairs_kernel(const KbTraverseParams p, int max_pairs, int32_t* __restrict__ out_pairs, int32_t* __restrict__ out_count) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int xf_floats = (p.nxf * 12 + 3) & ~3;
  const int nit_c = ITC ? p.nitems : 0;
  ItemS* s_items = (ItemS*)smem_raw;
  const size_t per_warp = (size_t)KB_STACK_CAP * 8 + (size_t)KB_LEAFQ_CAP * 8 + (size_t)KB_AP_MAX * 8 + (size_t)xf_floats * 4 + (size_t)nit_c * 48;
  unsigned char* base = smem_raw + (size_t)nit_c * 16 + warp * per_warp;
  uint2* stack = (uint2*)base;
  uint2* leafq = (uint2*)(base + (size_t)KB_STACK_CAP * 8);
  int2* foundp = (int2*)(base + (size_t)KB_STACK_CAP * 8 + (size_t)KB_LEAFQ_CAP * 8);
  float* xfw = (float*)(base + (size_t)KB_STACK_CAP * 8 + (size_t)KB_LEAFQ_CAP * 8 + (size_t)KB_AP_MAX * 8);
  float* itc = xfw + xf_floats;
  const KbScene& sc = p.scene;
  const float slack = 4.f * sc.eps_abs;
  if (ITC) {
    for (int i = threadIdx.x; i < p.nitems; i += blockDim.x) {
      const KbItem* it = p.items + i;
      ItemS s; s.nodeA = it->nodeA; s.nodeB = it->nodeB; s.infl = (float)it->thr + slack;
      s.xf = (int)((unsigned)(unsigned short)it->xfA | ((unsigned)(unsigned short)it->xfB << 16));
      s_items[i] = s;
    }
  }
  __syncthreads();
  unsigned lt_mask;
  asm("mov.u32 %0, %%lanemask_lt;" : "=r"(lt_mask));
  for (;;) {
    unsigned int c = 0;
    if (lane == 0) c = atomicAdd(p.work_counter, 1u);
    c = __shfl_sync(FULL, c, 0);
    if ((int64_t)c >= p.N) break;
    if (p.state && p.state[c] == 0) {
      if (lane == 0) out_count[c] = -1;
      for (int k = lane; k < 2 * max_pairs; k += 32) out_pairs[(size_t)c * max_pairs * 2 + k] = -1;
      continue;
    }
    const double* xf = p.xf64 + (size_t)c * (size_t)p.nxf * 12;
    __syncwarp();
    for (int i = lane; i < p.nxf * 12; i += 32) xfw[i] = (float)xf[i];
    __syncwarp();
    if (ITC) {
      for (int i = lane; i < p.nitems; i += 32) {
        const int xfp = s_items[i].xf;
        XfF T; rel_xf(xfw, (int)(short)(xfp & 0xffff), (int)(short)(xfp >> 16), T);
        float4* qv = (float4*)(itc + 12 * i);
        qv[0] = make_float4(T.r[0], T.r[1], T.r[2], T.r[3]); qv[1] = make_float4(T.r[4], T.r[5], T.r[6], T.r[7]); qv[2] = make_float4(T.r[8], T.t[0], T.t[1], T.t[2]);
      }
      __syncwarp();
    }
    int sp = 0, nleaf = 0, cursor = 0, nfound = 0;
    for (;;) {
      if (sp < 32 && cursor < p.nitems) {
        int k = p.nitems - cursor; if (k > 32) k = 32;
        if (lane < k) stack[sp + lane] = make_uint2(((unsigned)(cursor + lane) << KB_NODEA_BITS), 0u);
        sp += k; cursor += k;
        __syncwarp();
      }
      if (sp == 0 && nleaf == 0) break;
      if (nleaf >= 32 || sp == 0) {
        const int m = nleaf < 32 ? nleaf : 32;
        bool hit = false; int ia = -1, ib = -1;
        if (lane < m) {
          const uint2 e = leafq[nleaf - 1 - lane];
          const int item = (int)(e.x >> KB_NODEA_BITS);
          const KbItem& it = p.items[item];
          const int na = (int)(e.x & (KB_MAX_NODES_A - 1)), nb = (int)e.y;
          float4 a0, a1, b0, b1;
          load_node(sc.nodes, (size_t)(it.nodeA + na), a0, a1);
          load_node(sc.nodes, (size_t)(it.nodeB + nb), b0, b1);
          const int fa = it.elemA + ~__float_as_int(a0.w), ca = __float_as_int(a1.w);
          const int fb = it.elemB + ~__float_as_int(b0.w), cb = __float_as_int(b1.w);
          XfF T;
          if (ITC) load_itc(itc, item, T); else rel_xf(xfw, it.xfA, it.xfB, T);
          const float thr = (float)it.thr;
          for (int i = 0; i < ca && !hit; i++)
            for (int j = 0; j < cb && !hit; j++) {
              int pa = it.idA, pb = it.idB;
              if (pa < 0) pa = (it.kindA == KB_ELEM_TRI ? sc.triown : (it.kindA == KB_ELEM_BOX ? sc.boxown : sc.sphown))[fa + i];
              if (pb < 0) pb = (it.kindB == KB_ELEM_TRI ? sc.triown : (it.kindB == KB_ELEM_BOX ? sc.boxown : sc.sphown))[fb + j];
              kb_order_pair(it.flags, pa, pb);
              bool known = false;
              for (int k = 0; k < nfound && k < KB_AP_MAX; k++) known |= (foundp[k].x == pa && foundp[k].y == pb);
              if (known) continue;
              int r = fast_elem_collide<true>(sc, it, T, fa + i, fb + j, thr);
              if (r == KB_UNCERTAIN) r = exact_elem_collide<true>(sc, it, xf, fa + i, fb + j) ? KB_YES : KB_NO;
              if (r == KB_YES) { hit = true; ia = pa; ib = pb; }
            }
        }
        nleaf -= m;
        unsigned hm = __ballot_sync(FULL, hit);
        while (hm) {                                   // append the new id pairs one by one (warp-uniform)
          const int src = __ffs(hm) - 1; hm &= hm - 1;
          const int pa = __shfl_sync(FULL, ia, src), pb = __shfl_sync(FULL, ib, src);
          bool known = false;
          for (int k = 0; k < nfound && k < KB_AP_MAX; k++) known |= (foundp[k].x == pa && foundp[k].y == pb);
          if (!known) { if (lane == 0 && nfound < KB_AP_MAX) foundp[nfound] = make_int2(pa, pb); nfound++; __syncwarp(); }
        }
        __syncwarp();
        continue;
      }
      const int m = (sp <= p.wide_limit) ? (sp < 32 ? sp : 32) : 1;
      const bool act = lane < m;
      uint2 e = make_uint2(0u, 0u);
      if (act) e = stack[sp - 1 - lane];
      sp -= m;
      __syncwarp();
      bool push2 = false, leafpair = false;
      uint2 c0e = e, c1e = e;
      if (act) {
        // an item that names one id pair is finished once that pair is listed
        const KbItem* itp = p.items + (e.x >> KB_NODEA_BITS);
        bool skip = false;
        if (itp->idA >= 0 && itp->idB >= 0) {
          int qa = itp->idA, qb = itp->idB;
          kb_order_pair(itp->flags, qa, qb);
          for (int k = 0; k < nfound && k < KB_AP_MAX; k++) skip |= (foundp[k].x == qa && foundp[k].y == qb);
        }
        bool push4 = false;
        if (!skip) node_test<ITC>(p, s_items, itc, xfw, slack, e, false, push2, push4, leafpair, c0e, c1e);
      }
      const unsigned pm = __ballot_sync(FULL, push2), lm = __ballot_sync(FULL, leafpair);
      if (push2) { const int off = sp + 2 * __popc(pm & lt_mask); stack[off] = c1e; stack[off + 1] = c0e; }
      if (leafpair) leafq[nleaf + __popc(lm & lt_mask)] = e;
      sp += 2 * __popc(pm); nleaf += __popc(lm);
      __syncwarp();
    }
    if (lane == 0) out_count[c] = nfound;
    for (int k = lane; k < max_pairs; k += 32) {
      const int2 v = (k < nfound && k < KB_AP_MAX) ? foundp[k] : make_int2(-1, -1);
      out_pairs[((size_t)c * max_pairs + k) * 2] = v.x; out_pairs[((size_t)c * max_pairs + k) * 2 + 1] = v.y;
    }
    __syncwarp();
  }
}

// =============================================================================================== split pipeline
// The fused kernel above carries the element tests (fp32 filtered tri-tri + fp64 recheck) in the same kernel as the
// node loop; their register demand (166 natural) caps the whole kernel at 16 warps per SM although the node loop itself
// lives in ~50 registers.  The split pipeline re-queues between phases instead:
//   kb_nodes_kernel    lean (<= 64 registers, 32 warps / SM): the same warp-cooperative node traversal, but leaf pairs
//                      are appended to a global list (config, item, nodeA, nodeB) instead of being tested in place.
//                      A configuration that has emitted `leaf_budget` pairs stops traversing (it is almost surely in
//                      collision) and is flagged; so is one whose pairs did not fit the list.
//   kb_leaves_kernel   fat: one thread per listed leaf pair, fp32 filtered element tests + inline fp64 recheck; the first
//                      hit of a configuration is recorded with an atomicCAS.
//   kb_requeue_kernel  flagged configurations without a hit are marked for the fused kernel, which re-runs them from
//                      scratch (rare: a configuration needs >= leaf_budget candidate pairs none of which intersects).
// Early exit inside a configuration is given up; measured on C2 it only matters for configurations with more than 32
// candidate pairs, which the budget covers.
#define KB_NSTAGE_CAP 64

template <bool ITC, bool STATS>
__global__ void __launch_bounds__(KB_WARPS_PER_BLOCK * 32, 8)
kb_nodes_kernel(const KbTraverseParams p, const KbSplitParams q) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int xf_floats = (p.nxf * 12 + 3) & ~3;
  const int nit_c = ITC ? p.nitems : 0;
  ItemS* s_items = (ItemS*)smem_raw;
  const size_t per_warp = (size_t)q.stack_cap * 8 + (size_t)KB_NSTAGE_CAP * 8 + (size_t)xf_floats * 4 + (size_t)nit_c * 48;
  unsigned char* base = smem_raw + (size_t)nit_c * 16 + warp * per_warp;
  uint2* stack = (uint2*)base;
  uint2* stage = (uint2*)(base + (size_t)q.stack_cap * 8);
  float* xfw = (float*)(base + (size_t)q.stack_cap * 8 + (size_t)KB_NSTAGE_CAP * 8);
  float* itc = xfw + xf_floats;
  const KbScene& sc = p.scene;
  const float slack = 4.f * sc.eps_abs;
  if (ITC) {
    for (int i = threadIdx.x; i < p.nitems; i += blockDim.x) {
      const KbItem* it = p.items + i;
      ItemS s; s.nodeA = it->nodeA; s.nodeB = it->nodeB; s.infl = (float)it->thr + slack;
      s.xf = (int)((unsigned)(unsigned short)it->xfA | ((unsigned)(unsigned short)it->xfB << 16));
      s_items[i] = s;
    }
  }
  __syncthreads();
  unsigned lt_mask;
  asm("mov.u32 %0, %%lanemask_lt;" : "=r"(lt_mask));
  unsigned st_node = 0;
  const unsigned total_warps = gridDim.x * KB_WARPS_PER_BLOCK;
  unsigned grab = 8;
  for (;;) {
    unsigned int c0 = 0;
    if (lane == 0) c0 = atomicAdd(p.work_counter, grab);
    c0 = __shfl_sync(FULL, c0, 0);
    if ((int64_t)c0 >= p.N) break;
    const unsigned nN = (unsigned)p.N;
    const unsigned cend = (c0 + grab < nN) ? c0 + grab : nN;
    { const unsigned g = (nN - cend) / (4u * total_warps); grab = g >= 8u ? 8u : (g < 1u ? 1u : g); }
    for (unsigned c = c0; c < cend; c++) {
      if (p.state && p.state[c] == 0) continue;
      const double* xf = p.xf64 + (size_t)c * (size_t)p.nxf * 12;
      __syncwarp();
      for (int i = lane; i < p.nxf * 12; i += 32) xfw[i] = (float)xf[i];
      __syncwarp();
      if (ITC) {
        for (int i = lane; i < p.nitems; i += 32) {
          const int xfp = s_items[i].xf;
          XfF T; rel_xf(xfw, (int)(short)(xfp & 0xffff), (int)(short)(xfp >> 16), T);
          float4* qv = (float4*)(itc + 12 * i);
          qv[0] = make_float4(T.r[0], T.r[1], T.r[2], T.r[3]); qv[1] = make_float4(T.r[4], T.r[5], T.r[6], T.r[7]); qv[2] = make_float4(T.r[8], T.t[0], T.t[1], T.t[2]);
        }
        __syncwarp();
      }
      int sp = 0, nstage = 0, cursor = 0, emitted = 0;
      bool flagged = false;
      for (;;) {
        if (sp < 32 && cursor < p.nitems && emitted < q.leaf_budget) {   // feed root pairs of the next work items
          int k = p.nitems - cursor; if (k > 32) k = 32;
          if (lane < k) stack[sp + lane] = make_uint2(((unsigned)(cursor + lane) << KB_NODEA_BITS), 0u);
          sp += k; cursor += k;
          __syncwarp();
        }
        const bool done = sp == 0 || emitted >= q.leaf_budget;
        if (nstage >= 32 || (done && nstage > 0)) {  // append up to 32 staged leaf pairs to the global list
          const int n = nstage < 32 ? nstage : 32;
          unsigned long long b = 0;
          if (lane == 0) b = atomicAdd(q.leaf_count, (unsigned long long)n);
          b = __shfl_sync(FULL, b, 0);
          if (b + (unsigned long long)n <= (unsigned long long)q.leaf_cap) {
            if (lane < n) { const uint2 e = stage[nstage - 1 - lane]; q.leaf_list[b + lane] = make_uint4(c, e.x, e.y, 0u); }
          } else flagged = true;                     // list full: this configuration is re-run by the fused kernel
          nstage -= n;
          __syncwarp();
          continue;
        }
        if (done) { if (sp > 0 || cursor < p.nitems) flagged = true; break; }
        do {
          const int m = (sp <= q.wide_limit) ? (sp < 32 ? sp : 32) : 1;
          const bool act = lane < m;
          uint2 e = make_uint2(0u, 0u);
          if (act) e = stack[sp - 1 - lane];
          sp -= m;
          __syncwarp();
          bool push2 = false, leafpair = false;
          uint2 c0e = e, c1e = e;
          if (act) {
            const int item = (int)(e.x >> KB_NODEA_BITS);
            int nodeA, nodeB; float infl; XfF T;
            if (ITC) {
              const ItemS s = s_items[item];
              nodeA = s.nodeA; nodeB = s.nodeB; infl = s.infl;
              load_itc(itc, item, T);
            } else {
              const KbItem* itp = p.items + item;
              nodeA = itp->nodeA; nodeB = itp->nodeB; infl = (float)itp->thr + slack;
              rel_xf(xfw, itp->xfA, itp->xfB, T);
            }
            const int na = (int)(e.x & (KB_MAX_NODES_A - 1)), nb = (int)e.y;
            float4 a0, a1, b0, b1;
            load_node(sc.nodes, (size_t)(nodeA + na), a0, a1);
            load_node(sc.nodes, (size_t)(nodeB + nb), b0, b1);
            if (STATS) st_node++;
            if (sat6_overlap(a0, a1, b0, b1, T, infl)) {
              const int la = __float_as_int(a0.w), lb = __float_as_int(b0.w);
              if (la < 0 && lb < 0) leafpair = true;
              else {
                push2 = true;
                const float sa2 = a1.x * a1.x + a1.y * a1.y + a1.z * a1.z, sb2 = b1.x * b1.x + b1.y * b1.y + b1.z * b1.z;
                if (lb < 0 || (la >= 0 && sa2 >= sb2)) {
                  c0e = make_uint2((e.x & ~(unsigned)(KB_MAX_NODES_A - 1)) | (unsigned)la, e.y);
                  c1e = make_uint2(c0e.x + 1u, e.y);
                } else {
                  c0e = make_uint2(e.x, (unsigned)lb);
                  c1e = make_uint2(e.x, (unsigned)lb + 1u);
                }
              }
            }
          }
          const unsigned pm = __ballot_sync(FULL, push2), lm = __ballot_sync(FULL, leafpair);
          if (push2) { const int off = sp + 2 * __popc(pm & lt_mask); stack[off] = c1e; stack[off + 1] = c0e; }
          if (leafpair) stage[nstage + __popc(lm & lt_mask)] = e;
          sp += 2 * __popc(pm); nstage += __popc(lm); emitted += __popc(lm);
          __syncwarp();
        } while (sp > 0 && nstage < 32 && emitted < q.leaf_budget && !(sp < 32 && cursor < p.nitems));
      }
      if (lane == 0) q.flagged[c] = flagged ? 1 : 0;
      if (STATS) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) st_node += __shfl_xor_sync(FULL, st_node, o);
        if (lane == 0 && p.counters) atomicAdd(p.counters + 1, (unsigned long long)st_node);
        st_node = 0;
      }
    }
  }
}

// one thread per listed leaf pair
template <bool STATS>
__global__ void __launch_bounds__(256)
kb_leaves_kernel(const KbTraverseParams p, const KbSplitParams q) {
  const KbScene& sc = p.scene;
  unsigned long long n = *q.leaf_count;
  if (n > (unsigned long long)q.leaf_cap) n = q.leaf_cap;
  unsigned n_leaf = 0, n_re = 0;
  for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (unsigned long long)gridDim.x * blockDim.x) {
    const uint4 e = q.leaf_list[i];
    const unsigned c = e.x;
    if (p.hit[c] >= 0) continue;                       // already decided by another pair
    const int item = (int)(e.y >> KB_NODEA_BITS);
    const KbItem& it = p.items[item];
    const int na = (int)(e.y & (KB_MAX_NODES_A - 1)), nb = (int)e.z;
    float4 a0, a1, b0, b1;
    load_node(sc.nodes, (size_t)(it.nodeA + na), a0, a1);
    load_node(sc.nodes, (size_t)(it.nodeB + nb), b0, b1);
    const int fa = it.elemA + ~__float_as_int(a0.w), ca = __float_as_int(a1.w);
    const int fb = it.elemB + ~__float_as_int(b0.w), cb = __float_as_int(b1.w);
    const double* xf = p.xf64 + (size_t)c * (size_t)p.nxf * 12;
    // relative transform from the fp64 table, rounded once (the same values the node kernel used)
    __align__(16) float loc[24];
    const int sa = it.xfA, sb = it.xfB;
#pragma unroll
    for (int k = 0; k < 12; k++) { loc[k] = sa >= 0 ? (float)xf[12 * sa + k] : 0.f; loc[12 + k] = sb >= 0 ? (float)xf[12 * sb + k] : 0.f; }
    XfF T; rel_xf(loc, sa >= 0 ? 0 : -1, sb >= 0 ? 1 : -1, T);
    const float thr = (float)it.thr;
    bool hit = false; int ea = -1, eb = -1;
    for (int ii = 0; ii < ca && !hit; ii++)
      for (int jj = 0; jj < cb && !hit; jj++) {
        int r = fast_elem_collide<true>(sc, it, T, fa + ii, fb + jj, thr);
        if (STATS) n_leaf++;
        if (r == KB_UNCERTAIN) { if (STATS) n_re++; r = exact_elem_collide<true>(sc, it, xf, fa + ii, fb + jj) ? KB_YES : KB_NO; }
        if (r == KB_YES) { hit = true; ea = fa + ii; eb = fb + jj; }
      }
    if (hit && atomicCAS(p.hit + c, -1, item) == -1 && p.hit_elem) { p.hit_elem[2 * (size_t)c] = ea; p.hit_elem[2 * (size_t)c + 1] = eb; }
  }
  if (STATS && p.counters) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { n_leaf += __shfl_xor_sync(FULL, n_leaf, o); n_re += __shfl_xor_sync(FULL, n_re, o); }
    if ((threadIdx.x & 31) == 0) { atomicAdd(p.counters + 2, (unsigned long long)n_leaf); atomicAdd(p.counters + 0, (unsigned long long)n_re); }
  }
}

// state2[c] = 1 for the configurations the fused kernel has to redo: flagged by the node kernel and still without a hit
__global__ void kb_requeue_kernel(const uint8_t* __restrict__ state, const uint8_t* __restrict__ flagged, const int32_t* __restrict__ hit,
                                  int64_t N, uint8_t* __restrict__ state2, unsigned long long* requeued) {
  int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  bool r = false;
  if (c < N) { r = (!state || state[c] != 0) && flagged[c] != 0 && hit[c] < 0; state2[c] = r ? 1 : 0; }
  unsigned m = __ballot_sync(FULL, r);
  if (requeued && (threadIdx.x & 31) == 0 && m) atomicAdd(requeued, (unsigned long long)__popc(m));
}

// =============================================================================================== distance traversal
// Branch and bound over the same flattened BVHs (replaces AnyCollisionQuery::Distance behind
// WorldPlannerSettings::DistanceLowerBound, reference Cpp/Planning/PlannerSettings.cpp:109-115,570-620).  One warp per
// configuration, LIFO frontier in shared memory like the boolean kernel, but organised for pruning instead of early exit:
//   * test-before-push: a lane pops one pair, loads the two children of the side chosen for splitting (one 64 B line)
//     and the other side's node, computes both box-distance lower bounds and pushes the survivors FARTHER FIRST, so the
//     nearer child is on top of the stack -- a greedy descent that reaches a small distance quickly;
//   * every entry carries its lower bound and is re-checked against the running minimum when popped;
//   * narrow pops (8) until the first element distance has tightened the bound, then 32-wide;
//   * leaf pairs are evaluated eagerly while no bound is known; element distances are fp64 throughout.
#define KB_SPLIT_B 0x80000000u
#ifndef KB_DIST_SORT_FEED
#define KB_DIST_SORT_FEED 1
#endif

__device__ __forceinline__ void classify_pair(unsigned itembits, int na, int nb, const float4& a0, const float4& a1, const float4& b0, const float4& b1,
                                              bool& leaf, uint2& e) {
  const int la = __float_as_int(a0.w), lb = __float_as_int(b0.w);
  if (la < 0 && lb < 0) { leaf = true; e = make_uint2(itembits | (unsigned)na, (unsigned)nb); return; }
  leaf = false;
  const float sa2 = a1.x * a1.x + a1.y * a1.y + a1.z * a1.z, sb2 = b1.x * b1.x + b1.y * b1.y + b1.z * b1.z;
  if (lb < 0 || (la >= 0 && sa2 >= sb2)) e = make_uint2(itembits | (unsigned)la, (unsigned)nb);          // split A: slot A holds A's first child
  else e = make_uint2(itembits | (unsigned)na, (unsigned)lb | KB_SPLIT_B);                                // split B: slot B holds B's first child
}

// ---- fp32 element distances with explicit bounds: lo <= (true element distance - radii) <= hi, hi = +inf when the fp32 value
// cannot be trusted (thin / zero-area triangle, uncertain intersection).  `cut`: a pair whose cheap lower bound is already >= cut
// is not worth the full evaluation (returns lo >= cut, hi = +inf).  lb_fallback: a lower bound the caller already has (leaf boxes).
// squared distance point - triangle without the Voronoi-region walk: min over the three clamped edge projections, or the plane
// distance when the projection falls inside (branch-free: the lanes of a warp hold unrelated pairs).  NaN for a triangle whose
// fp32 normal cannot be trusted (kb_face_ok).
__device__ __forceinline__ float point_tri_dist2_bf(const V3<float>& p, const V3<float>& a, const V3<float>& b, const V3<float>& c) {
  const V3<float> ab = b - a, ac = c - a, bc = c - b, ap = p - a, bp = p - b;
  const float ab2 = dot(ab, ab), ac2 = dot(ac, ac), bc2 = dot(bc, bc);
  const float t1 = ab2 > 0.f ? __saturatef(__fdividef(dot(ab, ap), ab2)) : 0.f;
  const float t2 = ac2 > 0.f ? __saturatef(__fdividef(dot(ac, ap), ac2)) : 0.f;
  const float t3 = bc2 > 0.f ? __saturatef(__fdividef(dot(bc, bp), bc2)) : 0.f;
  const V3<float> q1 = madd(ap, ab, -t1), q2 = madd(ap, ac, -t2), q3 = madd(bp, bc, -t3);
  const float em = fminf(dot(q1, q1), fminf(dot(q2, q2), dot(q3, q3)));
  const V3<float> n = cross(ab, ac);
  const float nn = dot(n, n);
  if (!kb_face_ok(nn, ab2, ac2)) return __int_as_float(0x7fc00000);
  const float u = dot(cross(ab, ap), n), v = dot(cross(ap, ac), n), h = dot(ap, n);
  const bool inside = (u >= 0.f) & (v >= 0.f) & (u + v <= nn);
  return inside ? __fdividef(h * h, nn) : em;
}

template <bool BOXES>
__device__ __forceinline__ void elem_bounds(const KbScene& sc, const KbItem& it, const XfF& T, int ea, int eb, float band, float cut, float lb_fallback,
                                            float& lo, float& hi) {
  const float INF = __int_as_float(0x7f800000);
  if (BOXES && (it.kindA == KB_ELEM_BOX || it.kindB == KB_ELEM_BOX)) { const float d = fast_box_elem_distance(sc, it, T, ea, eb); lo = d - band; hi = d + band; return; }
  if (it.kindA == KB_ELEM_TRI && it.kindB == KB_ELEM_TRI) {
    const float4* ta = sc.tris32 + 3 * (size_t)ea; const float4* tb = sc.tris32 + 3 * (size_t)eb;
    const float4 a0 = __ldg(ta), a1 = __ldg(ta + 1), a2 = __ldg(ta + 2), b0 = __ldg(tb), b1 = __ldg(tb + 1), b2 = __ldg(tb + 2);
    V3<float> A[3] = {mk3<float>(a0.x, a0.y, a0.z), mk3<float>(a1.x, a1.y, a1.z), mk3<float>(a2.x, a2.y, a2.z)};
    V3<float> B[3] = {xform(T, b0), xform(T, b1), xform(T, b2)};
    // cheap plane-separation lower bound first (see fast_elem_distance): most pairs of a leaf phase cannot beat the running minimum
    const V3<float> ea1 = A[1] - A[0], ea2 = A[2] - A[0], eb1 = B[1] - B[0], eb2 = B[2] - B[0];
    const V3<float> nA = cross(ea1, ea2), nB = cross(eb1, eb2);
    const V3<float> w0 = B[0] - A[0], w1 = B[1] - A[0], w2 = B[2] - A[0], u1 = A[1] - B[0], u2 = A[2] - B[0];
    const float b0s = dot(nA, w0), b1s = dot(nA, w1), b2s = dot(nA, w2);
    const float a0s = -dot(nB, w0), a1s = dot(nB, u1), a2s = dot(nB, u2);
    const float slackA = 1e-6f * sqrtf(dot(ea1, ea1) * dot(ea2, ea2) * fmaxf(dot(w0, w0), fmaxf(dot(w1, w1), dot(w2, w2))));
    const float slackB = 1e-6f * sqrtf(dot(eb1, eb1) * dot(eb2, eb2) * fmaxf(dot(w0, w0), fmaxf(dot(u1, u1), dot(u2, u2))));
    float lbA = 0.f, lbB = 0.f;
    if ((b0s > 0.f && b1s > 0.f && b2s > 0.f) || (b0s < 0.f && b1s < 0.f && b2s < 0.f))
      lbA = fmaxf(fminf(fabsf(b0s), fminf(fabsf(b1s), fabsf(b2s))) - slackA, 0.f) * rsqrtf(fmaxf(dot(nA, nA), 1e-30f));
    if ((a0s > 0.f && a1s > 0.f && a2s > 0.f) || (a0s < 0.f && a1s < 0.f && a2s < 0.f))
      lbB = fmaxf(fminf(fabsf(a0s), fminf(fabsf(a1s), fabsf(a2s))) - slackB, 0.f) * rsqrtf(fmaxf(dot(nB, nB), 1e-30f));
    const float lbt = fmaxf(fmaxf(lbA, lbB) * (1.f - 1e-5f) - band, 0.f);
    lo = lbt; hi = INF;
    if (lbt >= cut) return;
    FiltF f; f.filt = 24.f * sc.eps_abs;
    V3<float> A2[3] = {A[0], A[1], A[2]}, B2[3] = {B[0], B[1], B[2]};
    const int r = tri_tri_intersect<float, FiltF>(A2, B2, f);
    if (r == KB_YES) { lo = 0.f; hi = 0.f; return; }
    if (r == KB_UNCERTAIN) { lo = 0.f; return; }
    if (!kb_face_ok(dot(nA, nA), dot(ea1, ea1), dot(ea2, ea2)) || !kb_face_ok(dot(nB, nB), dot(eb1, eb1), dot(eb2, eb2))) return;
    const float d = sqrtf(tri_tri_dist2_disjoint<float>(A, B));
    lo = fmaxf(lbt, d - band); hi = d + band;
    return;
  }
  float d;
  if (it.kindA == KB_ELEM_TRI) {
    const float4* ta = sc.tris32 + 3 * (size_t)ea;
    const float4 a0 = __ldg(ta), a1 = __ldg(ta + 1), a2 = __ldg(ta + 2), s = __ldg(sc.sph32 + eb);
    d = sqrtf(point_tri_dist2_bf(xform(T, s), mk3<float>(a0.x, a0.y, a0.z), mk3<float>(a1.x, a1.y, a1.z), mk3<float>(a2.x, a2.y, a2.z))) - s.w;
  } else if (it.kindB == KB_ELEM_TRI) {
    const float4* tb = sc.tris32 + 3 * (size_t)eb;
    const float4 b0 = __ldg(tb), b1 = __ldg(tb + 1), b2 = __ldg(tb + 2), s = __ldg(sc.sph32 + ea);
    d = sqrtf(point_tri_dist2_bf(mk3<float>(s.x, s.y, s.z), xform(T, b0), xform(T, b1), xform(T, b2))) - s.w;
  } else {
    const float4 sa = __ldg(sc.sph32 + ea), sb = __ldg(sc.sph32 + eb);
    const V3<float> dv = xform(T, sb) - mk3<float>(sa.x, sa.y, sa.z);
    d = sqrtf(dot(dv, dv)) - sa.w - sb.w;
  }
  if (d != d) { lo = lb_fallback; hi = INF; return; }      // NaN: a face normal fp32 cannot be trusted with -> the pair goes to fp64
  lo = d - band; hi = d + band;
}

// kb_distance_kernel (v2).  One warp per configuration, branch and bound as before, re-organised after the round-1 profile
// (166 registers, 13.9 active lanes, eager fp64 leaf distances inside the node loop):
//   * the hot path is fp32 only.  Every element pair gets bounds lo <= d <= hi (elem_bounds); `bound` = the smallest hi seen is a
//     valid upper bound of the answer and drives the pruning; a pair with lo < thr MAY be the minimum and is parked in a per-warp
//     candidate list.  The list is evaluated in fp64 32 candidates at a time (exact_elem_distance, the same arithmetic as before, in
//     one not-inlined cold path), which yields the exact minimum: the true closest pair p* always satisfies lo(p*) <= d(p*) <= bound.
//   * node phase: a lane pops one pair, loads the two children of the split side (one 64-byte line) and the other node, and
//     computes both box-gap bounds; the split side is chosen with selects, not branches.  Survivors are pushed farther first.
//   * element phase: a lane owns one leaf pair and walks its (<= 8 x 8) element pairs with the branch-free point / triangle code.
//   * relErr / absErr of AnyCollisionQuery::Distance: a pair is pruned when lb + absErr >= bound or lb (1 + relErr) >= bound,
//     so the reported value is within that tolerance above the true minimum (0 / 0 = exact).
struct KbDistArgs { double* out_dist; double* out_cp; double upper_bound; float rel_err, abs_err; };
#ifndef KB_CAND_CAP
#define KB_CAND_CAP 64
#endif
#ifndef KB_DIST_BPS
#define KB_DIST_BPS 4
#endif

// pruning threshold below the current bound: a pair whose lower bound lb satisfies lb + absErr >= bound or lb (1 + relErr) >= bound
// cannot improve the answer by more than the tolerance (PQP's rule), so value <= exact + absErr and value <= exact (1 + relErr)
__device__ __forceinline__ float kb_thr_from(float bound, float rel_err, float abs_err) {
  const float r = bound > 0.f ? __fdividef(bound, 1.f + rel_err) * (1.f - 2e-7f) : bound * (1.f + rel_err);
  return fminf(bound - abs_err, rel_err > 0.f ? r : bound);
}

template <bool BOXES>
__device__ __noinline__ void drain_candidates(const KbTraverseParams& p, const double* __restrict__ xf, uint4* cand, const float* cand_hi, int* cand_count, int lane,
                                              float rel_err, float abs_err, double& best64, int& best_item, int& best_ea, int& best_eb, float& bound, float& thr) {
  __syncwarp();
  int n = *cand_count; if (n > KB_CAND_CAP) n = KB_CAND_CAP;
  for (int base = 0; base < n; base += 32) {
    const int k = base + lane;
    double d = 1e300; int item = -1, ea = -1, eb = -1;
    if (k < n) {
      const uint4 c = cand[k];
      // still able to be the answer (lo <= thr), or the pair whose fp32 upper bound IS the current bound: that one is always evaluated, so
      // the bound the pruning used is backed by an exact value
      if (__uint_as_float(c.w) <= thr || cand_hi[k] <= bound) {
        item = (int)c.x; ea = (int)c.y; eb = (int)c.z;
        const KbItem& it = p.items[item];
        d = exact_elem_distance<BOXES>(p.scene, it, xf, ea, eb) - it.marg;
      }
    }
    double wmin = d;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { const double x = __shfl_xor_sync(FULL, wmin, o); wmin = x < wmin ? x : wmin; }
    if (wmin < best64) {
      const unsigned who = __ballot_sync(FULL, d == wmin);
      const int src = __ffs(who) - 1;
      best64 = wmin; best_item = __shfl_sync(FULL, item, src); best_ea = __shfl_sync(FULL, ea, src); best_eb = __shfl_sync(FULL, eb, src);
      float bf = (float)best64; if ((double)bf < best64) bf = nextafterf(bf, INFINITY);
      if (bf < bound) { bound = bf; thr = kb_thr_from(bound, rel_err, abs_err); }
    }
  }
  __syncwarp();
  if (lane == 0) *cand_count = 0;
  __syncwarp();
}

template <bool ITC, bool STATS, bool BOXES>
__global__ void __launch_bounds__(KB_WARPS_PER_BLOCK * 32, KB_DIST_BPS)
kb_distance_kernel(const KbTraverseParams p, const KbDistArgs da) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int xf_floats = (p.nxf * 12 + 3) & ~3;
  const int nit_c = ITC ? p.nitems : 0;
  ItemS* s_items = (ItemS*)smem_raw;
  const size_t per_warp = (size_t)KB_STACK_CAP * 12 + (size_t)KB_LEAFQ_CAP * 12 + (size_t)KB_CAND_CAP * 20 + 16 + (size_t)xf_floats * 4 + (size_t)nit_c * 48;
  unsigned char* base = smem_raw + (size_t)nit_c * 16 + warp * per_warp;
  uint2* stack = (uint2*)base;
  uint2* leafq = (uint2*)(base + (size_t)KB_STACK_CAP * 8);
  float* stack_lb = (float*)(base + (size_t)KB_STACK_CAP * 8 + (size_t)KB_LEAFQ_CAP * 8);
  float* leaf_lb = stack_lb + KB_STACK_CAP;
  uint4* cand = (uint4*)(leaf_lb + KB_LEAFQ_CAP);
  float* cand_hi = (float*)(cand + KB_CAND_CAP);
  int* cand_count = (int*)(cand_hi + KB_CAND_CAP);
  float* xfw = (float*)(cand_count + 4);
  float* itc = xfw + xf_floats;
  const KbScene& sc = p.scene;
  const float slack = 4.f * sc.eps_abs;
  const float INF = __int_as_float(0x7f800000);
  if (ITC) {
    for (int i = threadIdx.x; i < p.nitems; i += blockDim.x) {
      const KbItem* it = p.items + i;
      ItemS s; s.nodeA = it->nodeA; s.nodeB = it->nodeB; s.infl = (float)(it->marg);      // infl slot reused: margin sum
      s.xf = (int)((unsigned)(unsigned short)it->xfA | ((unsigned)(unsigned short)it->xfB << 16));
      s_items[i] = s;
    }
  }
  if (lane == 0) *cand_count = 0;
  __syncthreads();
  unsigned lt_mask;
  asm("mov.u32 %0, %%lanemask_lt;" : "=r"(lt_mask));
  unsigned st_node = 0, st_leaf = 0, st_exact = 0, st_iter = 0;
  const unsigned total_warps = gridDim.x * KB_WARPS_PER_BLOCK;
  unsigned grab;
  { const unsigned g0 = (unsigned)p.N / (4u * total_warps); grab = g0 >= 4u ? 4u : (g0 < 1u ? 1u : g0); }
  for (;;) {
    unsigned int c0 = 0;
    if (lane == 0) c0 = atomicAdd(p.work_counter, grab);
    c0 = __shfl_sync(FULL, c0, 0);
    if ((int64_t)c0 >= p.N) break;
    const unsigned nN = (unsigned)p.N;
    const unsigned cend = (c0 + grab < nN) ? c0 + grab : nN;
    { const unsigned g = (nN - cend) / (4u * total_warps); grab = g >= 4u ? 4u : (g < 1u ? 1u : g); }
    for (unsigned c = c0; c < cend; c++) {
      const double* xf = p.xf64 + (size_t)c * (size_t)p.nxf * 12;
      __syncwarp();
      for (int i = lane; i < p.nxf * 12; i += 32) xfw[i] = (float)xf[i];
      __syncwarp();
      if (ITC) {
        for (int i = lane; i < p.nitems; i += 32) {
          const int xfp = s_items[i].xf;
          XfF T; rel_xf(xfw, (int)(short)(xfp & 0xffff), (int)(short)(xfp >> 16), T);
          float4* q = (float4*)(itc + 12 * i);
          q[0] = make_float4(T.r[0], T.r[1], T.r[2], T.r[3]); q[1] = make_float4(T.r[4], T.r[5], T.r[6], T.r[7]); q[2] = make_float4(T.r[8], T.t[0], T.t[1], T.t[2]);
        }
        __syncwarp();
      }
      int sp = 0, nleaf = 0, cursor = 0;
      double best64 = da.upper_bound;                     // exact running minimum over the evaluated candidates, margins subtracted
      int best_item = -1, best_ea = -1, best_eb = -1;
      float bound = (float)fmin(da.upper_bound, 3.0e38);  // fp32 upper bound of the answer (rounded up), drives the pruning
      if ((double)bound < da.upper_bound) bound = nextafterf(bound, INFINITY);
      const float bound0 = bound;
      float thr = kb_thr_from(bound, da.rel_err, da.abs_err);
      const float band = 16.f * sc.eps_abs;
      for (;;) {
        const bool feed = sp < 32 && nleaf < 32 && cursor < p.nitems;
        if (!feed) {
          if (sp == 0 && nleaf == 0) break;
          if (nleaf >= 32 || sp == 0 || (nleaf > 0 && bound >= bound0)) {
            // -------------------------------------------------------------- element phase (fp32 bounds, candidates parked)
            const int m = nleaf < 32 ? nleaf : 32;
            float bound_l = bound, thr_l = thr;
            double dl = 1e300; int il = -1, eal = -1, ebl = -1;        // exact results of candidates that did not fit the list
            if (lane < m && leaf_lb[nleaf - 1 - lane] <= thr) {
              const uint2 e = leafq[nleaf - 1 - lane];
              const float lbq = leaf_lb[nleaf - 1 - lane];
              const int item = (int)(e.x >> KB_NODEA_BITS);
              const KbItem& it = p.items[item];
              const int na = (int)(e.x & (KB_MAX_NODES_A - 1)), nb = (int)e.y;
              float4 a0, a1, b0, b1;
              load_node(sc.nodes, (size_t)(it.nodeA + na), a0, a1);
              load_node(sc.nodes, (size_t)(it.nodeB + nb), b0, b1);
              const int fa = it.elemA + ~__float_as_int(a0.w), ca = __float_as_int(a1.w);
              const int fb = it.elemB + ~__float_as_int(b0.w), cb = __float_as_int(b1.w);
              XfF T;
              if (ITC) load_itc(itc, item, T); else rel_xf(xfw, it.xfA, it.xfB, T);
              const float margf = (float)it.marg, bandm = band + 2e-7f * fabsf(margf);
              for (int i = 0; i < ca; i++) {
                for (int j = 0; j < cb; j++) {
                  if (STATS) st_leaf++;
                  float lo, hi;
                  elem_bounds<BOXES>(sc, it, T, fa + i, fb + j, bandm, thr_l + margf + bandm, lbq + margf, lo, hi);
                  lo -= margf + bandm - band; hi -= margf - (bandm - band);
                  if (lo <= thr_l || hi < bound_l) {
                    const int slot = atomicAdd(cand_count, 1);
                    if (slot < KB_CAND_CAP) { cand[slot] = make_uint4((unsigned)item, (unsigned)(fa + i), (unsigned)(fb + j), __float_as_uint(lo)); cand_hi[slot] = hi; }
                    else {                                    // list full: evaluate in place (rare)
                      if (STATS) st_exact++;
                      const double d = exact_elem_distance<BOXES>(sc, it, xf, fa + i, fb + j) - it.marg;
                      if (d < dl) { dl = d; il = item; eal = fa + i; ebl = fb + j; }
                      float df = (float)d; if ((double)df < d) df = nextafterf(df, INFINITY);
                      hi = fminf(hi, df);
                    }
                    if (hi < bound_l) { bound_l = hi; thr_l = kb_thr_from(bound_l, da.rel_err, da.abs_err); }
                  }
                }
              }
            }
            nleaf -= m;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) bound_l = fminf(bound_l, __shfl_xor_sync(FULL, bound_l, o));
            if (bound_l < bound) { bound = bound_l; thr = kb_thr_from(bound, da.rel_err, da.abs_err); }
            if (__any_sync(FULL, dl < 1e300)) {
              double wmin = dl;
#pragma unroll
              for (int o = 16; o > 0; o >>= 1) { const double x = __shfl_xor_sync(FULL, wmin, o); wmin = x < wmin ? x : wmin; }
              if (wmin < best64) {
                const int src = __ffs(__ballot_sync(FULL, dl == wmin)) - 1;
                best64 = wmin; best_item = __shfl_sync(FULL, il, src); best_ea = __shfl_sync(FULL, eal, src); best_eb = __shfl_sync(FULL, ebl, src);
              }
            }
            __syncwarp();
            if (*cand_count >= KB_CAND_CAP / 2) {
              if (STATS) st_exact += (lane == 0) ? (unsigned)min(*cand_count, KB_CAND_CAP) : 0u;
              drain_candidates<BOXES>(p, xf, cand, cand_hi, cand_count, lane, da.rel_err, da.abs_err, best64, best_item, best_ea, best_eb, bound, thr);
            }
            continue;
          }
        }
        // ------------------------------------------------------------------ node phase (or root feed)
        int m;
        uint2 e = make_uint2(0u, 0u);
        bool live = false;
        if (feed) { m = p.nitems - cursor; if (m > 32) m = 32; live = lane < m; }
        else {
          const int width = bound < bound0 ? 32 : 8;
          m = (sp <= p.wide_limit) ? (sp < width ? sp : width) : 1;
          if (lane < m) { e = stack[sp - 1 - lane]; live = stack_lb[sp - 1 - lane] < thr; }   // re-check against the current bound
          sp -= m;
          __syncwarp();
        }
        if (STATS) st_iter += (lane == 0);
        bool ok0 = false, ok1 = false, leaf0 = false, leaf1 = false;
        uint2 e0 = e, e1 = e;
        float lb0 = 0.f, lb1 = 0.f;
        if (live) {
          const int item = feed ? cursor + lane : (int)(e.x >> KB_NODEA_BITS);
          const unsigned itembits = (unsigned)item << KB_NODEA_BITS;
          int nodeA, nodeB; float marg, rsum; XfF T;
          if (ITC) {
            const ItemS s = s_items[item];
            nodeA = s.nodeA; nodeB = s.nodeB; marg = s.infl;
            load_itc(itc, item, T);
            rsum = (float)p.items[item].rsum;
          } else {
            const KbItem* itp = p.items + item;
            nodeA = itp->nodeA; nodeB = itp->nodeB; marg = (float)itp->marg; rsum = (float)itp->rsum;
            rel_xf(xfw, itp->xfA, itp->xfB, T);
          }
          // conservative fp32 bound on (element distance - margins): box gap minus slack; touching boxes only bound by -radii
          const float margu = marg > 0.f ? marg * (1.f + 2.4e-7f) + 1e-30f : 0.f;
          const float touch = -rsum * (1.f + 2.4e-7f) - margu;     // floor of the item: nothing is closer than -(radii + margins); == 0 for bare meshes,
                                                                   // so once two triangles intersect (bound 0) every remaining pair is pruned
          if (feed) {
            float4 a0, a1, b0, b1;
            load_node(sc.nodes, (size_t)nodeA, a0, a1);
            load_node(sc.nodes, (size_t)nodeB, b0, b1);
            if (STATS) st_node++;
            const float g = box_dist_lb(a0, a1, b0, b1, T) - slack;
            lb0 = g > 0.f ? g * (1.f - 4e-7f) - margu : touch;
            ok0 = lb0 < thr;
            if (ok0) classify_pair(itembits, 0, 0, a0, a1, b0, b1, leaf0, e0);
          } else {
            const bool splitB = (e.y & KB_SPLIT_B) != 0;
            const int ra = (int)(e.x & (KB_MAX_NODES_A - 1)), rb = (int)(e.y & ~KB_SPLIT_B);
            // children of the split side (siblings: one 64-byte line) and the node of the other side; which is which by selects
            float4 c00, c01, c10, c11, o0, o1;
            const size_t ic = (size_t)(splitB ? nodeB + rb : nodeA + ra), io = (size_t)(splitB ? nodeA + ra : nodeB + rb);
            load_node(sc.nodes, ic, c00, c01); load_node(sc.nodes, ic + 1, c10, c11); load_node(sc.nodes, io, o0, o1);
            if (STATS) st_node += 2;
            const float4 A00 = splitB ? o0 : c00, A01 = splitB ? o1 : c01, B00 = splitB ? c00 : o0, B01 = splitB ? c01 : o1;
            const float4 A10 = splitB ? o0 : c10, A11 = splitB ? o1 : c11, B10 = splitB ? c10 : o0, B11 = splitB ? c11 : o1;
            const float g0 = box_dist_lb(A00, A01, B00, B01, T) - slack, g1 = box_dist_lb(A10, A11, B10, B11, T) - slack;
            lb0 = g0 > 0.f ? g0 * (1.f - 4e-7f) - margu : touch;
            lb1 = g1 > 0.f ? g1 * (1.f - 4e-7f) - margu : touch;
            ok0 = lb0 < thr; ok1 = lb1 < thr;
            if (ok0) classify_pair(itembits, ra, rb, A00, A01, B00, B01, leaf0, e0);
            if (ok1) classify_pair(itembits, splitB ? ra : ra + 1, splitB ? rb + 1 : rb, A10, A11, B10, B11, leaf1, e1);
            if (ok0 && ok1 && lb0 < lb1) {              // candidate 1 is pushed last = on top: make it the nearer one
              const uint2 te = e0; e0 = e1; e1 = te; const bool tl = leaf0; leaf0 = leaf1; leaf1 = tl; const float tf = lb0; lb0 = lb1; lb1 = tf;
            }
          }
        }
        if (feed) cursor += m;
        const unsigned pm0 = __ballot_sync(FULL, ok0 && !leaf0), pm1 = __ballot_sync(FULL, ok1 && !leaf1);
        const unsigned lm0 = __ballot_sync(FULL, ok0 && leaf0), lm1 = __ballot_sync(FULL, ok1 && leaf1);
        int slot0 = __popc(pm0 & lt_mask);
        if (feed && KB_DIST_SORT_FEED) {
          // root pairs go onto the stack farthest first, so the nearest work items are on top and are descended first: the
          // running minimum is tight before the far items are looked at (they are then rejected when popped)
          slot0 = 0;
          for (int j = 0; j < 32; j++) {
            const float lj = __shfl_sync(FULL, lb0, j);
            if (((pm0 >> j) & 1u) && (lj > lb0 || (lj == lb0 && j < lane))) slot0++;
          }
        }
        if (ok0 && !leaf0) { const int o = sp + slot0; stack[o] = e0; stack_lb[o] = lb0; }
        if (ok1 && !leaf1) { const int o = sp + __popc(pm0) + __popc(pm1 & lt_mask); stack[o] = e1; stack_lb[o] = lb1; }
        if (ok0 && leaf0) { const int o = nleaf + __popc(lm0 & lt_mask); leafq[o] = e0; leaf_lb[o] = lb0; }
        if (ok1 && leaf1) { const int o = nleaf + __popc(lm0) + __popc(lm1 & lt_mask); leafq[o] = e1; leaf_lb[o] = lb1; }
        sp += __popc(pm0) + __popc(pm1); nleaf += __popc(lm0) + __popc(lm1);
        __syncwarp();
      }
      if (STATS) st_exact += (lane == 0) ? (unsigned)min(*cand_count, KB_CAND_CAP) : 0u;
      drain_candidates<BOXES>(p, xf, cand, cand_hi, cand_count, lane, da.rel_err, da.abs_err, best64, best_item, best_ea, best_eb, bound, thr);
      if (STATS) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { st_node += __shfl_xor_sync(FULL, st_node, o); st_leaf += __shfl_xor_sync(FULL, st_leaf, o); st_exact += __shfl_xor_sync(FULL, st_exact, o); }
        if (lane == 0 && p.counters) {
          atomicAdd(p.counters + 1, (unsigned long long)st_node); atomicAdd(p.counters + 2, (unsigned long long)st_leaf);
          atomicAdd(p.counters + 0, (unsigned long long)st_exact); atomicAdd(p.counters + 8, (unsigned long long)st_iter);
        }
        st_node = st_leaf = st_exact = st_iter = 0;
      }
      if (lane == 0) {
        da.out_dist[c] = best64;
        p.hit[c] = best_item;
        if (p.hit_elem) { p.hit_elem[2 * (size_t)c] = best_ea; p.hit_elem[2 * (size_t)c + 1] = best_eb; }
      }
    }
  }
}

// =============================================================================================== result kernels
// feasible[c] = limits ok && no hit ; first_pair = world ids of the reported pair
__global__ void kb_finish_kernel(const uint8_t* __restrict__ state, const int32_t* __restrict__ hit, const int32_t* __restrict__ hit_elem,
                                 const KbItem* __restrict__ items, const int32_t* __restrict__ triown, const int32_t* __restrict__ sphown, const int32_t* __restrict__ boxown,
                                 int64_t N, uint8_t* __restrict__ out, int32_t* __restrict__ first_pair, unsigned long long* nfeasible) {
  int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  bool feas = false;
  if (c < N) {
    int h = hit[c];
    feas = state[c] != 0 && h < 0;
    if (out) out[c] = feas ? 1 : 0;
    if (first_pair) {
      int ia = -1, ib = -1;
      if (state[c] != 0 && h >= 0) {
        const KbItem it = items[h];
        ia = it.idA; ib = it.idB;
        if (ia < 0) ia = (it.kindA == KB_ELEM_TRI ? triown : (it.kindA == KB_ELEM_BOX ? boxown : sphown))[hit_elem[2 * c]];
        if (ib < 0) ib = (it.kindB == KB_ELEM_TRI ? triown : (it.kindB == KB_ELEM_BOX ? boxown : sphown))[hit_elem[2 * c + 1]];
    kb_order_pair(it.flags, ia, ib);
        kb_order_pair(it.flags, ia, ib);
      }
      first_pair[2 * c] = ia; first_pair[2 * c + 1] = ib;
    }
  }
  if (nfeasible) {
    unsigned m = __ballot_sync(FULL, feas);
    if ((threadIdx.x & 31) == 0 && m) atomicAdd(nfeasible, (unsigned long long)__popc(m));
  }
}

__global__ void kb_pair_ids_kernel(const int32_t* __restrict__ hit, const int32_t* __restrict__ hit_elem, const KbItem* __restrict__ items,
                                   const int32_t* __restrict__ triown, const int32_t* __restrict__ sphown, const int32_t* __restrict__ boxown, int64_t N, int32_t* __restrict__ pair) {
  int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= N) return;
  int h = hit[c], ia = -1, ib = -1;
  if (h >= 0) {
    const KbItem it = items[h];
    ia = it.idA; ib = it.idB;
    if (ia < 0) ia = (it.kindA == KB_ELEM_TRI ? triown : (it.kindA == KB_ELEM_BOX ? boxown : sphown))[hit_elem[2 * c]];
    if (ib < 0) ib = (it.kindB == KB_ELEM_TRI ? triown : (it.kindB == KB_ELEM_BOX ? boxown : sphown))[hit_elem[2 * c + 1]];
    kb_order_pair(it.flags, ia, ib);
  }
  pair[2 * c] = ia; pair[2 * c + 1] = ib;
}

// =============================================================================================== SO(3) for Floating / BallAndSocket joints
// Klampt::Interpolate / Klampt::Distance (Cpp/Modeling/Interpolate.cpp:16-52,229-278) treat the last three links of a Floating
// joint (and the three links of a BallAndSocket joint) as Euler angles about z, y, x: R = Rz(a) Ry(b) Rx(c); interpolation
// follows the SO(3) geodesic Ra exp(u log(Ra^T Rb)) and the metric collects the geodesic angle acos((tr(Ra Rb^T) - 1) / 2).
__device__ __forceinline__ void kb_euler_zyx_to_matrix(double a, double b, double c, double* R) {
  double sa, ca, sb, cb, sc, cc; sincos(a, &sa, &ca); sincos(b, &sb, &cb); sincos(c, &sc, &cc);
  R[0] = ca * cb; R[1] = ca * sb * sc - sa * cc; R[2] = ca * sb * cc + sa * sc;
  R[3] = sa * cb; R[4] = sa * sb * sc + ca * cc; R[5] = sa * sb * cc - ca * sc;
  R[6] = -sb;     R[7] = cb * sc;                R[8] = cb * cc;
}
__device__ __forceinline__ double kb_so3_angle(const double* R) { double c = 0.5 * (R[0] + R[4] + R[8] - 1.0); c = fmin(1.0, fmax(-1.0, c)); return acos(c); }
__device__ double kb_euler_zyx_angle_between(const double* ea, const double* eb) {
  double Ra[9], Rb[9], D[9]; kb_euler_zyx_to_matrix(ea[0], ea[1], ea[2], Ra); kb_euler_zyx_to_matrix(eb[0], eb[1], eb[2], Rb);
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) D[3 * i + j] = Ra[3 * i] * Rb[3 * j] + Ra[3 * i + 1] * Rb[3 * j + 1] + Ra[3 * i + 2] * Rb[3 * j + 2];
  return kb_so3_angle(D);
}
__device__ void kb_euler_zyx_interp(const double* ea, const double* eb, double u, double* out) {
  const double PI = 3.14159265358979323846;
  double Ra[9], Rb[9], D[9], w[3], E[9], Ru[9];
  kb_euler_zyx_to_matrix(ea[0], ea[1], ea[2], Ra); kb_euler_zyx_to_matrix(eb[0], eb[1], eb[2], Rb);
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) D[3 * i + j] = Ra[i] * Rb[j] + Ra[3 + i] * Rb[3 + j] + Ra[6 + i] * Rb[6 + j];      // Ra^T Rb
  {   // log
    const double th = kb_so3_angle(D);
    const double v[3] = {D[7] - D[5], D[2] - D[6], D[3] - D[1]};
    if (th < 1e-9) { for (int k = 0; k < 3; k++) w[k] = 0.5 * v[k]; }
    else if (PI - th < 1e-6) {   // near a half turn: axis from the diagonal of (R + I) / 2, signs from the symmetric and skew parts
      double ax[3]; for (int k = 0; k < 3; k++) { double d = 0.5 * (D[4 * k] + 1.0); ax[k] = d > 0 ? sqrt(d) : 0.0; }
      int m = 0; for (int k = 1; k < 3; k++) if (ax[k] > ax[m]) m = k;
      for (int k = 0; k < 3; k++) if (k != m) { double s2 = D[3 * m + k] + D[3 * k + m]; if (s2 < 0) ax[k] = -ax[k]; }
      double sg = ax[0] * v[0] + ax[1] * v[1] + ax[2] * v[2]; if (sg < 0) for (int k = 0; k < 3; k++) ax[k] = -ax[k];
      double n = sqrt(ax[0] * ax[0] + ax[1] * ax[1] + ax[2] * ax[2]); for (int k = 0; k < 3; k++) w[k] = th * ax[k] / n;
    } else { const double f = th / (2.0 * sin(th)); for (int k = 0; k < 3; k++) w[k] = f * v[k]; }
  }
  for (int k = 0; k < 3; k++) w[k] *= u;
  {   // exp (Rodrigues)
    const double th = sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
    if (th < 1e-12) { E[0] = 1; E[1] = -w[2]; E[2] = w[1]; E[3] = w[2]; E[4] = 1; E[5] = -w[0]; E[6] = -w[1]; E[7] = w[0]; E[8] = 1; }
    else {
      const double x = w[0] / th, y = w[1] / th, z = w[2] / th; double s, c; sincos(th, &s, &c); const double vv = 1.0 - c;
      E[0] = c + vv * x * x;     E[1] = vv * x * y - s * z; E[2] = vv * x * z + s * y;
      E[3] = vv * y * x + s * z; E[4] = c + vv * y * y;     E[5] = vv * y * z - s * x;
      E[6] = vv * z * x - s * y; E[7] = vv * z * y + s * x; E[8] = c + vv * z * z;
    }
  }
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) Ru[3 * i + j] = Ra[3 * i] * E[j] + Ra[3 * i + 1] * E[3 + j] + Ra[3 * i + 2] * E[6 + j];
  double sb = fmin(1.0, fmax(-1.0, -Ru[6]));
  out[1] = asin(sb);
  if (fabs(Ru[6]) < 1.0 - 1e-12) { out[0] = atan2(Ru[3], Ru[0]); out[2] = atan2(Ru[7], Ru[8]); }
  else { out[2] = 0.0; out[0] = atan2(-Ru[1], Ru[4]); }      // gimbal lock: the whole turn goes into the z angle
}

// =============================================================================================== edges (K8)
// Edge e of C-space length len needs nlev = number of halvings until len/2^nlev <= eps.  Level l (1-based) holds the
// 2^(l-1) odd multiples k/2^l.  All alive edges are expanded level by level; an edge dies at the first level that
// holds an infeasible midpoint, and the sequential checker's check count is recovered from the lowest failing k.
#ifndef KB_EDGE_MAX_LEVELS
#define KB_EDGE_MAX_LEVELS 24    // an edge is bisected into at most 2^24 pieces (eps = 1e-7 of its length)
#endif
__global__ void kb_edge_setup_kernel(const KbRobotDev* __restrict__ robot, const double* __restrict__ A, const double* __restrict__ B,
                                     const double* __restrict__ weights, int64_t N, double eps, int32_t* __restrict__ nlev,
                                     uint8_t* __restrict__ alive, int32_t* __restrict__ nchecks, int32_t* maxlev) {
  int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= N) return;
  const int L = robot->L;
  ExactD s(0.0);
  for (int j = 0; j < robot->nj; j++) {
    int t = robot->jtype[j]; int k = robot->jlink[j];
    ExactD w(weights ? weights[j] : 1.0), d(0.0);
    if (t == 1) d = ExactD(A[e * L + k]) - ExactD(B[e * L + k]);
    else if (t == 2) {
      double x = fmod(A[e * L + k], 6.283185307179586476925286766559); if (x < 0) x += 6.283185307179586476925286766559;
      double y = fmod(B[e * L + k], 6.283185307179586476925286766559); if (y < 0) y += 6.283185307179586476925286766559;
      double dd = x - y; if (dd > 3.14159265358979323846) dd -= 6.283185307179586476925286766559; else if (dd < -3.14159265358979323846) dd += 6.283185307179586476925286766559;
      d = ExactD(dd);
    } else if (t == 3) {          // Floating: three translations + the geodesic angle (floatingRotationWeight = 1)
      const int16_t* ix = robot->jidx[j];
      for (int q3 = 0; q3 < 3; q3++) { ExactD dt = ExactD(A[e * L + ix[q3]]) - ExactD(B[e * L + ix[q3]]); s = s + w * dt * dt; }
      const double ea[3] = {A[e * L + ix[3]], A[e * L + ix[4]], A[e * L + ix[5]]}, eb[3] = {B[e * L + ix[3]], B[e * L + ix[4]], B[e * L + ix[5]]};
      d = ExactD(kb_euler_zyx_angle_between(ea, eb));
    } else if (t == 5) {          // BallAndSocket: the geodesic angle
      const int16_t* ix = robot->jidx[j];
      const double ea[3] = {A[e * L + ix[0]], A[e * L + ix[1]], A[e * L + ix[2]]}, eb[3] = {B[e * L + ix[0]], B[e * L + ix[1]], B[e * L + ix[2]]};
      d = ExactD(kb_euler_zyx_angle_between(ea, eb));
    } else continue;              // Weld; FloatingPlanar contributes nothing (Interpolate.cpp:338-340)
    s = s + w * d * d;
  }
  double len = kb_sqrt(s).v;
  if (!(len == len) || len > 1.7e308) {       // NaN / inf coordinates: no path, and no 2^k midpoints to enumerate
    nlev[e] = 0; alive[e] = 0; nchecks[e] = 0;
    return;
  }
  int n = 0;
  while (len > eps && n < KB_EDGE_MAX_LEVELS) { len *= 0.5; n++; }
  if (len > eps) atomicMax(maxlev + 2, 1);    // still longer than eps after 2^KB_EDGE_MAX_LEVELS pieces: reported as an error, never checked coarser than asked
  nlev[e] = n; alive[e] = 1; nchecks[e] = 0;
  if (n > 0) atomicMax(maxlev, n);
}

// compacts the edges that are alive and reach level `lev`, and gives each its slot offset in this level's batch
__global__ void kb_edge_count_kernel(const int32_t* __restrict__ nlev, const uint8_t* __restrict__ alive, int64_t N, int lev,
                                     int32_t* __restrict__ list, unsigned int* __restrict__ count) {
  int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  bool on = e < N && alive[e] && nlev[e] >= lev;
  unsigned m = __ballot_sync(FULL, on);
  int lane = threadIdx.x & 31;
  unsigned base = 0;
  if (lane == 0 && m) base = atomicAdd(count, (unsigned)__popc(m));
  base = __shfl_sync(FULL, base, 0);
  if (on) list[base + __popc(m & ((1u << lane) - 1u))] = (int32_t)e;
}

// writes the midpoints of level `lev` for the listed edges: slot = i * per + j, k = 2j+1, u = k / 2^lev.
// Klampt::Interpolate (Cpp/Modeling/Interpolate.cpp:10-71): out = x*(1-u); out += y*u; Spin joints take the short arc.
// Klampt::Interpolate(a, b, u) of edge e into q (Cpp/Modeling/Interpolate.cpp:10-71): out = x*(1-u); out += y*u; Spin joints and the
// angle of FloatingPlanar joints take the short arc, Euler-ZYX triplets of Floating / BallAndSocket joints the SO(3) geodesic
__device__ __forceinline__ void kb_edge_midpoint(const KbRobotDev* __restrict__ robot, const double* __restrict__ A, const double* __restrict__ B,
                                                 int64_t e, ExactD u, double* __restrict__ q) {
  const int L = robot->L;
  ExactD um = ExactD(1.0) - u;
  for (int k = 0; k < L; k++) q[k] = (ExactD(A[e * L + k]) * um + ExactD(B[e * L + k]) * u).v;
  for (int jn = 0; jn < robot->nj; jn++) {
    const int jt = robot->jtype[jn];
    if (jt == 3 || jt == 5) {     // Floating / BallAndSocket: Euler-ZYX triplet along the SO(3) geodesic
      const int16_t* ix = robot->jidx[jn] + (jt == 3 ? 3 : 0);
      const double ea[3] = {A[e * L + ix[0]], A[e * L + ix[1]], A[e * L + ix[2]]}, eb[3] = {B[e * L + ix[0]], B[e * L + ix[1]], B[e * L + ix[2]]};
      double eu[3]; kb_euler_zyx_interp(ea, eb, u.v, eu);
      q[ix[0]] = eu[0]; q[ix[1]] = eu[1]; q[ix[2]] = eu[2];
      continue;
    }
    if (jt != 2 && jt != 4) continue;
    int k = jt == 2 ? robot->jlink[jn] : robot->jidx[jn][2];      // Spin, or the angle of a FloatingPlanar joint: short arc
    const double tp = 6.283185307179586476925286766559;
    double x = fmod(A[e * L + k], tp); if (x < 0) x += tp;
    double y = fmod(B[e * L + k], tp); if (y < 0) y += tp;
    double d = y - x; if (d > 3.14159265358979323846) d -= tp; else if (d < -3.14159265358979323846) d += tp;
    double r = fmod((ExactD(x) + u * ExactD(d)).v, tp); if (r < 0) r += tp;
    q[k] = r;
  }
}

__global__ void kb_edge_expand_kernel(const KbRobotDev* __restrict__ robot, const double* __restrict__ A, const double* __restrict__ B,
                                      const int32_t* __restrict__ list, int64_t first_slot, int64_t nslots, int lev, double* __restrict__ Q) {
  int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= nslots) return;
  const int64_t per = (int64_t)1 << (lev - 1);
  int64_t g = first_slot + s;
  int64_t i = g / per, j = g % per;
  kb_edge_midpoint(robot, A, B, (int64_t)list[i], ExactD((double)(2 * j + 1) / (double)((int64_t)1 << lev)), Q + s * robot->L);
}

// ---- small edge batches: every midpoint of every level in ONE batch.  The level-by-level form stops an edge at the first level with an
// infeasible midpoint, at the price of a launch sequence and a host read-back per level (0.7-1.3 ms for one edge at eps = 0.01).  When
// N (2^maxlev - 1) midpoints fit one launch the machine is idle anyway: slot s = edge (s / per_max), midpoint j = s % per_max in the
// sequential checker's order (level l = 1 + floor(log2(j + 1)), k = j + 1 - 2^(l-1), u = (2k + 1) / 2^l), all checked at once, and the
// lowest failing j per edge gives the same visibility and the same nchecks (= j + 1) as the sequential early exit.
__global__ void kb_edge_flat_expand_kernel(const KbRobotDev* __restrict__ robot, const double* __restrict__ A, const double* __restrict__ B,
                                           const int32_t* __restrict__ nlev, const uint8_t* __restrict__ alive, int64_t nslots, int per_max,
                                           double* __restrict__ Q, uint8_t* __restrict__ slot_on, unsigned long long* __restrict__ nactive) {
  const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  bool on = false;
  if (s < nslots) {
    const int64_t e = s / per_max; const int j = (int)(s % per_max);
    on = alive[e] && j < (1 << nlev[e]) - 1;
    slot_on[s] = on ? 1 : 0;
    if (on) {
      const int l = 32 - __clz(j + 1), k = j + 1 - (1 << (l - 1));
      kb_edge_midpoint(robot, A, B, e, ExactD((double)(2 * k + 1) / (double)(1 << l)), Q + s * robot->L);
    } else for (int k = 0; k < robot->L; k++) Q[s * robot->L + k] = 0.0;
  }
  const unsigned m = __ballot_sync(FULL, on);
  if (nactive && (threadIdx.x & 31) == 0 && m) atomicAdd(nactive, (unsigned long long)__popc(m));
}
__global__ void kb_edge_flat_reduce_kernel(const uint8_t* __restrict__ feas, const uint8_t* __restrict__ slot_on, int64_t nslots, int per_max, int32_t* __restrict__ firstbad) {
  const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= nslots || !slot_on[s] || feas[s]) return;
  atomicMin(firstbad + s / per_max, (int32_t)(s % per_max));
}
__global__ void kb_edge_flat_end_kernel(const int32_t* __restrict__ nlev, const int32_t* __restrict__ firstbad, int64_t N, uint8_t* __restrict__ alive, int32_t* __restrict__ nchecks) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= N || !alive[e]) return;
  const int count = (1 << nlev[e]) - 1, fb = firstbad[e];
  if (fb < count) { alive[e] = 0; nchecks[e] = fb + 1; } else nchecks[e] = count;
}

// folds the feasibility bytes of one level chunk back into the edges: lowest infeasible j per edge
__global__ void kb_edge_reduce_kernel(const uint8_t* __restrict__ feas, const int32_t* __restrict__ list, int64_t first_slot, int64_t nslots,
                                      int lev, int32_t* __restrict__ firstbad) {
  int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= nslots) return;
  if (feas[s]) return;
  const int64_t per = (int64_t)1 << (lev - 1);
  int64_t g = first_slot + s;
  atomicMin(firstbad + list[g / per], (int32_t)(g % per));
}

__global__ void kb_edge_level_end_kernel(const int32_t* __restrict__ list, unsigned int nlist, int lev, int32_t* __restrict__ firstbad,
                                         uint8_t* __restrict__ alive, int32_t* __restrict__ nchecks) {
  unsigned int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nlist) return;
  int e = list[i];
  int fb = firstbad[e];
  int per = 1 << (lev - 1);
  if (fb < per) { alive[e] = 0; nchecks[e] += fb + 1; firstbad[e] = 0x7fffffff; }
  else nchecks[e] += per;
}

__global__ void kb_fill_i32_kernel(int32_t* p, int64_t n, int32_t v) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; if (i < n) p[i] = v; }

__global__ void kb_copy_u8_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, int64_t n, unsigned long long* ones) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  bool v = false;
  if (i < n) { v = src[i] != 0; dst[i] = src[i]; }
  if (ones) { unsigned m = __ballot_sync(FULL, v); if ((threadIdx.x & 31) == 0 && m) atomicAdd(ones, (unsigned long long)__popc(m)); }
}

// =============================================================================================== launchers
static inline unsigned int nblocks(int64_t n, int bs) { return (unsigned int)((n + bs - 1) / bs); }

#define KB_ITC_MAX_ITEMS 256
size_t kb_distance_smem_bytes(int nxf, int nitems) {
  size_t xf_floats = ((size_t)nxf * 12 + 3) & ~(size_t)3;
  size_t nit = nitems <= KB_ITC_MAX_ITEMS ? (size_t)nitems : 0;
  return nit * 16 + (size_t)KB_WARPS_PER_BLOCK * ((size_t)KB_STACK_CAP * 12 + (size_t)KB_LEAFQ_CAP * 12 + (size_t)KB_CAND_CAP * 20 + 16 + xf_floats * 4 + nit * 48);
}

template <bool ITC, bool STATS, bool BOXES>
static cudaError_t launch_distance_t(const KbTraverseParams& p, const KbDistArgs& da, int num_sms, size_t smem, cudaStream_t s) {
  static bool attr_set[64] = {false};      // the attribute is per device
  int dev = 0; cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64 || !attr_set[dev]) {
    cudaError_t e = cudaFuncSetAttribute(kb_distance_kernel<ITC, STATS, BOXES>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) return e;
    if (dev >= 0 && dev < 64) attr_set[dev] = true;
  }
  int per_sm = (int)((224 * 1024) / (smem + 1024)); if (per_sm < 1) per_sm = 1; if (per_sm > KB_DIST_BPS) per_sm = KB_DIST_BPS;
  int64_t want = (p.N + KB_WARPS_PER_BLOCK - 1) / KB_WARPS_PER_BLOCK;
  int64_t grid = (int64_t)num_sms * per_sm; if (grid > want) grid = want; if (grid < 1) grid = 1;
  kb_distance_kernel<ITC, STATS, BOXES><<<(unsigned)grid, KB_WARPS_PER_BLOCK * 32, smem, s>>>(p, da);
  return cudaGetLastError();
}

size_t kb_traverse_smem_bytes(int nxf, int nitems, int nprobes) {
  size_t xf_floats = ((size_t)nxf * 12 + 3) & ~(size_t)3;
  size_t nit = nitems <= KB_ITC_MAX_ITEMS ? (size_t)nitems : 0;
  size_t mask_words = nprobes > 0 ? ((((size_t)nitems + 31) >> 5) + 3) & ~(size_t)3 : 0;
  size_t nps = nprobes <= KB_PROBES_SMEM_MAX ? (size_t)(nprobes > 0 ? nprobes : 0) : 0;
  return nit * 16 + nps * 32 + (size_t)KB_WARPS_PER_BLOCK * ((size_t)KB_STACK_CAP * 8 + (size_t)KB_BOOL_LEAFQ_CAP * 8 + (size_t)KB_RQ_CAP * 16 + 16 + xf_floats * 4 + nit * 48 + mask_words * 4);
}

cudaError_t kb_launch_fk(const KbRobotDev* robot, const KbDriverDev* drv, const int32_t* drv_link, const double* drv_scale,
                         const double* drv_off, const double* Q, int64_t N, double* xf64, int nxf, uint8_t* state,
                         const uint8_t* alive, int32_t* hit, cudaStream_t s) {
  if (N <= 0) return cudaSuccess;
  kb_fk_kernel<<<nblocks(N, KB_FK_THREADS), KB_FK_THREADS, 0, s>>>(robot, drv, drv_link, drv_scale, drv_off, Q, N, xf64, nxf, state, alive, hit);
  return cudaGetLastError();
}

template <bool ITC, bool STATS, int BPS, bool BOXES>
static cudaError_t launch_traverse_t(const KbTraverseParams& p, int num_sms, size_t smem, cudaStream_t s) {
  static bool attr_set[64] = {false};      // the attribute is per device
  int dev = 0; cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64 || !attr_set[dev]) {
    cudaError_t e = cudaFuncSetAttribute(kb_traverse_kernel<ITC, STATS, BPS, BOXES, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e == cudaSuccess && !BOXES) e = cudaFuncSetAttribute(kb_traverse_wide_kernel<ITC, STATS, BPS, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e == cudaSuccess && !STATS) e = cudaFuncSetAttribute(kb_traverse_kernel<ITC, false, 3, BOXES, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e == cudaSuccess && !STATS && !BOXES) e = cudaFuncSetAttribute(kb_traverse_wide_kernel<ITC, false, 3, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) return e;
    if (dev >= 0 && dev < 64) attr_set[dev] = true;
  }
  int per_sm = (int)((224 * 1024) / (smem + 1024)); if (per_sm < 1) per_sm = 1; if (per_sm > BPS) per_sm = BPS;
  int64_t want = (p.N + KB_WARPS_PER_BLOCK - 1) / KB_WARPS_PER_BLOCK;      // at most one warp per configuration
  int64_t grid = (int64_t)num_sms * per_sm; if (grid > want) grid = want; if (grid < 1) grid = 1;
  if (p.static_sched) grid = (p.N + KB_WARPS_PER_BLOCK - 1) / KB_WARPS_PER_BLOCK;      // one warp per configuration
  if (p.static_sched && !STATS) {      // small batches: the 168-register build (occupancy is irrelevant at one warp per configuration)
    if (p.use_wide && !BOXES) kb_traverse_wide_kernel<ITC, false, 3, false, true><<<(unsigned)grid, KB_WARPS_PER_BLOCK * 32, smem, s>>>(p);
    else kb_traverse_kernel<ITC, false, 3, BOXES, true><<<(unsigned)grid, KB_WARPS_PER_BLOCK * 32, smem, s>>>(p);
  }
  else if (p.use_wide && !BOXES) kb_traverse_wide_kernel<ITC, STATS, BPS, false, false><<<(unsigned)grid, KB_WARPS_PER_BLOCK * 32, smem, s>>>(p);
  else kb_traverse_kernel<ITC, STATS, BPS, BOXES, false><<<(unsigned)grid, KB_WARPS_PER_BLOCK * 32, smem, s>>>(p);
  return cudaGetLastError();
}

cudaError_t kb_launch_traverse(const KbTraverseParams& p, int mode, double* out_dist, double upper_bound, int num_sms, cudaStream_t s, double* out_cp, float rel_err, float abs_err) {
  if (p.N <= 0) return cudaSuccess;
  const size_t smem = mode == 0 ? kb_traverse_smem_bytes(p.nxf, p.nitems, p.nprobes) : kb_distance_smem_bytes(p.nxf, p.nitems);
  if (smem > 200 * 1024) return cudaErrorInvalidConfiguration;
  cudaError_t e = (mode == 0 && p.static_sched) ? cudaSuccess : cudaMemsetAsync(p.work_counter, 0, sizeof(uint32_t), s);
  if (e != cudaSuccess) return e;
  const bool itc = p.nitems <= KB_ITC_MAX_ITEMS;
  if (p.N > 0xfffffff0ll) return cudaErrorInvalidValue;
  if (mode == 1) {
    KbDistArgs da; da.out_dist = out_dist; da.out_cp = out_cp; da.upper_bound = upper_bound; da.rel_err = rel_err; da.abs_err = abs_err;
#define KB_LD(I, S) (p.has_boxes ? launch_distance_t<I, S, true>(p, da, num_sms, smem, s) : launch_distance_t<I, S, false>(p, da, num_sms, smem, s))
    if (p.collect_stats) return itc ? KB_LD(true, true) : KB_LD(false, true);
    return itc ? KB_LD(true, false) : KB_LD(false, false);
#undef KB_LD
  }
  if (p.N > 0xfffffff0ll) return cudaErrorInvalidValue;
  // two register budgets are compiled: KB_BPS_HI = 5 CTAs/SM (20 warps, 102 registers, some spills in the element phase) wins
  // when the per-CTA shared memory is small (few work items per configuration; C2: 7.33 ms at 5, 7.68 at 4, 8.07 at 6),
  // 3 CTAs/SM (168 registers, no spills) wins when the item cache is large (C3, 107 pairs: 10.5 vs 11.8 ms at 4).
  // Measured on B200, profiles/r01_experiments.md.
#ifndef KB_HI_SMEM_LIMIT
#define KB_HI_SMEM_LIMIT (40 * 1024)
#endif
  const bool four = smem <= KB_HI_SMEM_LIMIT;
#ifndef KB_BPS_HI
#define KB_BPS_HI 5
#endif
#define KB_LT2(I, S, B) (four ? launch_traverse_t<I, S, KB_BPS_HI, B>(p, num_sms, smem, s) : launch_traverse_t<I, S, 3, B>(p, num_sms, smem, s))
#define KB_LT(I, S) (p.has_boxes ? KB_LT2(I, S, true) : KB_LT2(I, S, false))
  if (p.collect_stats) return itc ? KB_LT(true, true) : KB_LT(false, true);
  return itc ? KB_LT(true, false) : KB_LT(false, false);
#undef KB_LT
#undef KB_LT2
}

size_t kb_nodes_smem_bytes(int nxf, int nitems, int stack_cap) {
  size_t xf_floats = ((size_t)nxf * 12 + 3) & ~(size_t)3;
  size_t nit = nitems <= KB_ITC_MAX_ITEMS ? (size_t)nitems : 0;
  return nit * 16 + (size_t)KB_WARPS_PER_BLOCK * ((size_t)stack_cap * 8 + (size_t)KB_NSTAGE_CAP * 8 + xf_floats * 4 + nit * 48);
}

template <bool ITC, bool STATS>
static cudaError_t launch_nodes_t(const KbTraverseParams& p, const KbSplitParams& q, int num_sms, size_t smem, cudaStream_t s) {
  static bool attr_set[64] = {false};
  int dev = 0; cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64 || !attr_set[dev]) {
    cudaError_t e = cudaFuncSetAttribute(kb_nodes_kernel<ITC, STATS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) return e;
    if (dev >= 0 && dev < 64) attr_set[dev] = true;
  }
  int per_sm = (int)((224 * 1024) / (smem + 1024)); if (per_sm < 1) per_sm = 1; if (per_sm > 8) per_sm = 8;
  int64_t want = (p.N + 8 * KB_WARPS_PER_BLOCK - 1) / (8 * KB_WARPS_PER_BLOCK);
  int64_t grid = (int64_t)num_sms * per_sm; if (grid > want) grid = want; if (grid < 1) grid = 1;
  kb_nodes_kernel<ITC, STATS><<<(unsigned)grid, KB_WARPS_PER_BLOCK * 32, smem, s>>>(p, q);
  return cudaGetLastError();
}

// split pipeline, stages 1-3 (node traversal -> leaf tests -> requeue mask); the caller then runs the fused kernel on q.state2
cudaError_t kb_launch_split(const KbTraverseParams& p, const KbSplitParams& q, int num_sms, cudaStream_t s) {
  if (p.N <= 0) return cudaSuccess;
  if (p.N > 0xfffffff0ll) return cudaErrorInvalidValue;
  const size_t smem = kb_nodes_smem_bytes(p.nxf, p.nitems, q.stack_cap);
  if (smem > 200 * 1024) return cudaErrorInvalidConfiguration;
  cudaError_t e = cudaMemsetAsync(p.work_counter, 0, sizeof(uint32_t), s);
  if (e != cudaSuccess) return e;
  e = cudaMemsetAsync(q.leaf_count, 0, sizeof(unsigned long long), s);
  if (e != cudaSuccess) return e;
  const bool itc = p.nitems <= KB_ITC_MAX_ITEMS;
  if (p.collect_stats) e = itc ? launch_nodes_t<true, true>(p, q, num_sms, smem, s) : launch_nodes_t<false, true>(p, q, num_sms, smem, s);
  else e = itc ? launch_nodes_t<true, false>(p, q, num_sms, smem, s) : launch_nodes_t<false, false>(p, q, num_sms, smem, s);
  if (e != cudaSuccess) return e;
  const unsigned lgrid = (unsigned)num_sms * 8;
  if (p.collect_stats) kb_leaves_kernel<true><<<lgrid, 256, 0, s>>>(p, q); else kb_leaves_kernel<false><<<lgrid, 256, 0, s>>>(p, q);
  e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  kb_requeue_kernel<<<nblocks(p.N, 256), 256, 0, s>>>(p.state, q.flagged, p.hit, p.N, q.state2, q.requeued);
  return cudaGetLastError();
}

cudaError_t kb_launch_allpairs(const KbTraverseParams& p, int max_pairs, int32_t* out_pairs, int32_t* out_count, int num_sms, cudaStream_t s) {
  if (p.N <= 0) return cudaSuccess;
  if (p.N > 0xfffffff0ll || max_pairs < 1 || max_pairs > KB_AP_MAX) return cudaErrorInvalidValue;
  const bool itc = p.nitems <= KB_ITC_MAX_ITEMS;
  const size_t xf_floats = ((size_t)p.nxf * 12 + 3) & ~(size_t)3, nit = itc ? (size_t)p.nitems : 0;
  const size_t smem = nit * 16 + (size_t)KB_WARPS_PER_BLOCK * ((size_t)KB_STACK_CAP * 8 + (size_t)KB_LEAFQ_CAP * 8 + (size_t)KB_AP_MAX * 8 + xf_floats * 4 + nit * 48);
  if (smem > 200 * 1024) return cudaErrorInvalidConfiguration;
  cudaError_t e = cudaMemsetAsync(p.work_counter, 0, sizeof(uint32_t), s);
  if (e != cudaSuccess) return e;
  e = itc ? cudaFuncSetAttribute(kb_allpairs_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)
          : cudaFuncSetAttribute(kb_allpairs_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  if (e != cudaSuccess) return e;
  int per_sm = (int)((224 * 1024) / (smem + 1024)); if (per_sm < 1) per_sm = 1; if (per_sm > 3) per_sm = 3;
  int64_t want = (p.N + KB_WARPS_PER_BLOCK - 1) / KB_WARPS_PER_BLOCK;
  int64_t grid = (int64_t)num_sms * per_sm; if (grid > want) grid = want; if (grid < 1) grid = 1;
  if (itc) kb_allpairs_kernel<true><<<(unsigned)grid, KB_WARPS_PER_BLOCK * 32, smem, s>>>(p, max_pairs, out_pairs, out_count);
  else kb_allpairs_kernel<false><<<(unsigned)grid, KB_WARPS_PER_BLOCK * 32, smem, s>>>(p, max_pairs, out_pairs, out_count);
  return cudaGetLastError();
}

cudaError_t kb_launch_finish(const uint8_t* state, const int32_t* hit, const int32_t* hit_elem, const KbItem* items, const int32_t* triown,
                             const int32_t* sphown, const int32_t* boxown, int64_t N, uint8_t* out, int32_t* first_pair, unsigned long long* nfeasible, cudaStream_t s) {
  if (N <= 0) return cudaSuccess;
  kb_finish_kernel<<<nblocks(N, 256), 256, 0, s>>>(state, hit, hit_elem, items, triown, sphown, boxown, N, out, first_pair, nfeasible);
  return cudaGetLastError();
}

cudaError_t kb_launch_pair_ids(const int32_t* hit, const int32_t* hit_elem, const KbItem* items, const int32_t* triown, const int32_t* sphown, const int32_t* boxown,
                               int64_t N, int32_t* pair, cudaStream_t s) {
  if (N <= 0) return cudaSuccess;
  kb_pair_ids_kernel<<<nblocks(N, 256), 256, 0, s>>>(hit, hit_elem, items, triown, sphown, boxown, N, pair);
  return cudaGetLastError();
}

cudaError_t kb_launch_edge_setup(const KbRobotDev* robot, const double* A, const double* B, const double* w, int64_t N, double eps,
                                 int32_t* nlev, uint8_t* alive, int32_t* nchecks, int32_t* maxlev, cudaStream_t s) {
  if (N <= 0) return cudaSuccess;
  kb_edge_setup_kernel<<<nblocks(N, 256), 256, 0, s>>>(robot, A, B, w, N, eps, nlev, alive, nchecks, maxlev);
  return cudaGetLastError();
}
cudaError_t kb_launch_edge_count(const int32_t* nlev, const uint8_t* alive, int64_t N, int lev, int32_t* list, unsigned int* count, cudaStream_t s) {
  if (N <= 0) return cudaSuccess;
  kb_edge_count_kernel<<<nblocks(N, 256), 256, 0, s>>>(nlev, alive, N, lev, list, count);
  return cudaGetLastError();
}
cudaError_t kb_launch_edge_expand(const KbRobotDev* robot, const double* A, const double* B, const int32_t* list, int64_t first_slot,
                                  int64_t nslots, int lev, double* Q, cudaStream_t s) {
  if (nslots <= 0) return cudaSuccess;
  kb_edge_expand_kernel<<<nblocks(nslots, 256), 256, 0, s>>>(robot, A, B, list, first_slot, nslots, lev, Q);
  return cudaGetLastError();
}
cudaError_t kb_launch_edge_flat_expand(const KbRobotDev* robot, const double* A, const double* B, const int32_t* nlev, const uint8_t* alive, int64_t nslots, int per_max,
                                       double* Q, uint8_t* slot_on, unsigned long long* nactive, cudaStream_t s) {
  if (nslots <= 0) return cudaSuccess;
  kb_edge_flat_expand_kernel<<<nblocks(nslots, 256), 256, 0, s>>>(robot, A, B, nlev, alive, nslots, per_max, Q, slot_on, nactive);
  return cudaGetLastError();
}
cudaError_t kb_launch_edge_flat_finish(const uint8_t* feas, const uint8_t* slot_on, int64_t nslots, int per_max, const int32_t* nlev, int32_t* firstbad, int64_t N,
                                       uint8_t* alive, int32_t* nchecks, cudaStream_t s) {
  if (nslots <= 0 || N <= 0) return cudaSuccess;
  kb_edge_flat_reduce_kernel<<<nblocks(nslots, 256), 256, 0, s>>>(feas, slot_on, nslots, per_max, firstbad);
  kb_edge_flat_end_kernel<<<nblocks(N, 256), 256, 0, s>>>(nlev, firstbad, N, alive, nchecks);
  return cudaGetLastError();
}
cudaError_t kb_launch_edge_reduce(const uint8_t* feas, const int32_t* list, int64_t first_slot, int64_t nslots, int lev, int32_t* firstbad, cudaStream_t s) {
  if (nslots <= 0) return cudaSuccess;
  kb_edge_reduce_kernel<<<nblocks(nslots, 256), 256, 0, s>>>(feas, list, first_slot, nslots, lev, firstbad);
  return cudaGetLastError();
}
cudaError_t kb_launch_edge_level_end(const int32_t* list, unsigned int nlist, int lev, int32_t* firstbad, uint8_t* alive, int32_t* nchecks, cudaStream_t s) {
  if (nlist == 0) return cudaSuccess;
  kb_edge_level_end_kernel<<<nblocks(nlist, 256), 256, 0, s>>>(list, nlist, lev, firstbad, alive, nchecks);
  return cudaGetLastError();
}
__global__ void kb_widen_f32_kernel(const float* __restrict__ src, double* __restrict__ dst, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = (double)src[i];
}
cudaError_t kb_launch_widen_f32(const float* src, double* dst, int64_t n, cudaStream_t s) {
  if (n <= 0) return cudaSuccess;
  kb_widen_f32_kernel<<<nblocks(n, 256), 256, 0, s>>>(src, dst, n);
  return cudaGetLastError();
}

// result bytes -> packed bitmask: bit (c & 7) of byte (c >> 3) = configuration c (the interface SURVEY 8b names: N / 8 bytes gathered)
__global__ void kb_pack_bits_kernel(const uint8_t* __restrict__ src, int64_t n, uint32_t* __restrict__ dst) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned m = __ballot_sync(FULL, i < n && src[i] != 0);
  if ((threadIdx.x & 31) == 0 && (i - (i & 31)) < n) dst[i >> 5] = m;
}
cudaError_t kb_launch_pack_bits(const uint8_t* src, int64_t n, uint32_t* dst, cudaStream_t s) {
  if (n <= 0) return cudaSuccess;
  kb_pack_bits_kernel<<<nblocks(n, 256), 256, 0, s>>>(src, n, dst);
  return cudaGetLastError();
}

cudaError_t kb_launch_fill_i32(int32_t* p, int64_t n, int32_t v, cudaStream_t s) {
  if (n <= 0) return cudaSuccess;
  kb_fill_i32_kernel<<<nblocks(n, 256), 256, 0, s>>>(p, n, v);
  return cudaGetLastError();
}
cudaError_t kb_launch_copy_u8(const uint8_t* src, uint8_t* dst, int64_t n, unsigned long long* ones, cudaStream_t s) {
  if (n <= 0) return cudaSuccess;
  kb_copy_u8_kernel<<<nblocks(n, 256), 256, 0, s>>>(src, dst, n, ones);
  return cudaGetLastError();
}
