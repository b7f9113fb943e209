// kb_raycast.cu -- batched ray casting against the world.
//
// Replaces WorldModel::RayCast / RayCastIgnore called once per ray (reference Cpp/Modeling/World.cpp:465-588): the camera sensor's
// fallback renders a depth image with one call per pixel (Cpp/Sensing/VisualSensors.cpp:430-475), the laser sensor one per
// measurement (:113-134), and Python's WorldCollider.rayCast / collide.ray_cast loop over Geometry3D.rayCast
// (Python/klampt/model/collide.py:225-243,700-748; Python/klampt/src/geometry.cpp:1821-1852).  Semantics kept: the closest hit over
// the robot's links at one configuration, the rigid objects and the terrains; a tie keeps the body the reference visits first
// (links in order, then objects, then terrains); a triangle mesh reports its nearest two-sided ray / triangle intersection
// (watertight: no ray slips between two triangles that share an edge) minus its collision margin; a point cloud is the union of spheres of radius (point radius + margin).
//
// One thread per ray (neighbouring pixels of an image share most of their path, so a warp's node loads fall into the same lines).
// Bodies are the per-geometry local-frame hierarchies every registered geometry already has; static bodies sit under a small
// top-level hierarchy of their world boxes, links are tested one by one in the frames FK gives.  Node boxes are tested in fp32 with
// the box padded by a bound on the rounding of the ray, elements in fp64: the reported distance is the fp64 one.
#include "kb_types.h"
#include "kb_kernels.h"
#include <cuda_runtime.h>
#include <math.h>
#include <float.h>

namespace {

#define KB_RAY_STACK 96
#define KB_RAY_TSTACK 40

__device__ __forceinline__ void ld_node(const float4* __restrict__ nodes, size_t idx, float4& n0, float4& n1) {
  const float4* p = nodes + 2 * idx;
  n0 = __ldg(p); n1 = __ldg(p + 1);
}

struct RayF { float ox, oy, oz, ix, iy, iz; };

// slab test of a padded box; tn = entry parameter (<= 0 when the source is inside).  Conservative: pad covers the rounding of the
// fp32 ray and box, the relative slack the rounding of the products.
__device__ __forceinline__ bool slab(const float4& n0, const float4& n1, const RayF& r, float pad, float tlim, float& tn) {
  const float cx = n0.x - r.ox, cy = n0.y - r.oy, cz = n0.z - r.oz;
  const float hx = n1.x + pad, hy = n1.y + pad, hz = n1.z + pad;
  const float ax = (cx - hx) * r.ix, bx = (cx + hx) * r.ix;
  const float ay = (cy - hy) * r.iy, by = (cy + hy) * r.iy;
  const float az = (cz - hz) * r.iz, bz = (cz + hz) * r.iz;
  const float t0 = fmaxf(fmaxf(fminf(ax, bx), fminf(ay, by)), fminf(az, bz));
  const float t1 = fminf(fminf(fmaxf(ax, bx), fmaxf(ay, by)), fmaxf(az, bz));
  tn = t0;
  return (t0 - 2e-6f * fabsf(t0) <= t1 + 2e-6f * fabsf(t1)) && (t1 >= 0.f) && (t0 - 2e-6f * fabsf(t0) <= tlim);
}

// Two-sided, watertight ray / triangle intersection in fp64 -- the oracle's statement operation by operation (no contraction: every
// product and sum is rounded on its own, so a static body answers bit for bit like the strict CPU build).  The triangle is sheared
// into the ray's frame and tested with three 2-D edge functions whose sign is exact (difference of two products evaluated with
// its rounding error) and antisymmetric in the edge's vertices: a ray through an edge or vertex shared by two triangles hits at
// least one of them (Woop, Benthin, Wald 2013).
struct RayShear { int kx, ky, kz; double Sx, Sy, Sz; };
__device__ __forceinline__ void ray_shear(const double* d, RayShear& r) {
  int kz = 0; if (fabs(d[1]) > fabs(d[kz])) kz = 1; if (fabs(d[2]) > fabs(d[kz])) kz = 2;
  int kx = kz == 2 ? 0 : kz + 1, ky = kx == 2 ? 0 : kx + 1;
  const double dz = kz == 0 ? d[0] : (kz == 1 ? d[1] : d[2]);
  if (dz < 0.0) { const int t = kx; kx = ky; ky = t; }
  const double dx = kx == 0 ? d[0] : (kx == 1 ? d[1] : d[2]), dy = ky == 0 ? d[0] : (ky == 1 ? d[1] : d[2]);
  r.kx = kx; r.ky = ky; r.kz = kz; r.Sx = __ddiv_rn(dx, dz); r.Sy = __ddiv_rn(dy, dz); r.Sz = __ddiv_rn(1.0, dz);
}
__device__ __forceinline__ double pick(const double* __restrict__ v, int k) { return k == 0 ? v[0] : (k == 1 ? v[1] : v[2]); }
__device__ __forceinline__ double diff_of_products(double a, double b, double c, double d) {
  const double p1 = __dmul_rn(a, b), e1 = __fma_rn(a, b, -p1), p2 = __dmul_rn(c, d), e2 = __fma_rn(c, d, -p2);
  return __dadd_rn(__dsub_rn(p1, p2), __dsub_rn(e1, e2));
}
__device__ __forceinline__ bool ray_tri(const double* s, const RayShear& r, const double* __restrict__ T, double& t) {
  const double sx = pick(s, r.kx), sy = pick(s, r.ky), sz = pick(s, r.kz);
  const double Az = __dsub_rn(pick(T, r.kz), sz), Bz = __dsub_rn(pick(T + 3, r.kz), sz), Cz = __dsub_rn(pick(T + 6, r.kz), sz);
  const double Ax = __dsub_rn(__dsub_rn(pick(T, r.kx), sx), __dmul_rn(r.Sx, Az)), Ay = __dsub_rn(__dsub_rn(pick(T, r.ky), sy), __dmul_rn(r.Sy, Az));
  const double Bx = __dsub_rn(__dsub_rn(pick(T + 3, r.kx), sx), __dmul_rn(r.Sx, Bz)), By = __dsub_rn(__dsub_rn(pick(T + 3, r.ky), sy), __dmul_rn(r.Sy, Bz));
  const double Cx = __dsub_rn(__dsub_rn(pick(T + 6, r.kx), sx), __dmul_rn(r.Sx, Cz)), Cy = __dsub_rn(__dsub_rn(pick(T + 6, r.ky), sy), __dmul_rn(r.Sy, Cz));
  const double U = diff_of_products(Cx, By, Cy, Bx), V = diff_of_products(Ax, Cy, Ay, Cx), W = diff_of_products(Bx, Ay, By, Ax);
  if ((U < 0.0 || V < 0.0 || W < 0.0) && (U > 0.0 || V > 0.0 || W > 0.0)) return false;
  const double det = __dadd_rn(__dadd_rn(U, V), W);
  if (!(det != 0.0)) return false;
  const double Tn = __dadd_rn(__dadd_rn(__dmul_rn(U, __dmul_rn(r.Sz, Az)), __dmul_rn(V, __dmul_rn(r.Sz, Bz))), __dmul_rn(W, __dmul_rn(r.Sz, Cz)));
  if ((det < 0.0 && Tn > 0.0) || (det > 0.0 && Tn < 0.0)) return false;
  const double tt = __ddiv_rn(Tn, det);
  if (!(tt >= 0.0)) return false;
  t = tt; return true;
}

__device__ __forceinline__ bool ray_sphere(const double* s, const double* d, const double* __restrict__ c, double r, double& t) {
  if (!(r > 0.0)) return false;
  const double mx = s[0] - c[0], my = s[1] - c[1], mz = s[2] - c[2];
  const double b = mx * d[0] + my * d[1] + mz * d[2], cc = mx * mx + my * my + mz * mz - r * r;
  if (cc <= 0.0) { t = 0.0; return true; }
  if (b > 0.0) return false;
  const double disc = b * b - cc;
  if (disc < 0.0) return false;
  t = -b - sqrt(disc); if (t < 0.0) t = 0.0;
  return true;
}

struct Best { double d; int rank, id, elem; };

struct WorldCounts { int nterr, nobj, nlinks; };
// order in which WorldModel::RayCast visits a static body with world id `own`: the links come first, then the rigid objects, then the terrains
__device__ __forceinline__ int rank_of_owner(int own, const WorldCounts& wc) { return own < wc.nterr ? wc.nlinks + wc.nobj + own : wc.nlinks + (own - wc.nterr); }

// nearest hit of the world-frame ray (s, d) with one body whose frame is T (world <- local, 12 doubles, null = world frame).
// Two nested loops (Aila & Laine's while-while form): descend through inner nodes until a leaf is reached, then test its elements --
// the lanes of a warp stay together in the cheap fp32 node loop instead of waiting for the one lane that is in an fp64 element test.
__device__ void cast_body(const KbScene& sc, const KbRayBody& B, const double* __restrict__ T, const double* s, const double* d, Best& best,
                          const uint8_t* __restrict__ ignore, const WorldCounts& wc) {
  double sl[3], dl[3];
  if (T) {
    // R^T (s - t) and R^T d, every product and sum rounded on its own (the oracle's order: bit-identical local rays)
    const double mx = __dsub_rn(s[0], T[9]), my = __dsub_rn(s[1], T[10]), mz = __dsub_rn(s[2], T[11]);
#pragma unroll
    for (int k = 0; k < 3; k++) {
      sl[k] = __dadd_rn(__dadd_rn(__dmul_rn(T[k], mx), __dmul_rn(T[3 + k], my)), __dmul_rn(T[6 + k], mz));
      dl[k] = __dadd_rn(__dadd_rn(__dmul_rn(T[k], d[0]), __dmul_rn(T[3 + k], d[1])), __dmul_rn(T[6 + k], d[2]));
    }
  } else { sl[0] = s[0]; sl[1] = s[1]; sl[2] = s[2]; dl[0] = d[0]; dl[1] = d[1]; dl[2] = d[2]; }
  const bool mesh = B.kind == KB_ELEM_TRI;
  const bool grouped = B.id < 0;                              // merged environment group: owner id (and rank) per element
  RayShear sh; bool have_shear = false;         // three fp64 divisions: only once a leaf is reached (most bodies are missed at their root box)
  const double shift = mesh ? B.margin : 0.0;                 // a mesh reports t - margin; a cloud's margin is part of its spheres
  double tbest = best.d + shift;                              // raw parameter this body has to beat (ties resolved by rank)
  if (!(tbest >= 0.0)) return;
  RayF r;
  r.ox = (float)sl[0]; r.oy = (float)sl[1]; r.oz = (float)sl[2];
  {
    const float fx = (float)dl[0], fy = (float)dl[1], fz = (float)dl[2];
    r.ix = 1.f / (fabsf(fx) > 1e-30f ? fx : copysignf(1e-30f, fx));
    r.iy = 1.f / (fabsf(fy) > 1e-30f ? fy : copysignf(1e-30f, fy));
    r.iz = 1.f / (fabsf(fz) > 1e-30f ? fz : copysignf(1e-30f, fz));
  }
  // (a replaceable cloud has no fixed extent: eps_abs = 8 * 2^-24 * scene extent, and every update of the cloud widens it)
  const float ext = B.ext >= 0.f ? B.ext : 3.f * sc.eps_abs * 2.1e6f;
  const float pad = (mesh ? 0.f : (float)B.margin * 1.000001f) + 8.f * sc.eps_abs + 4e-6f * (fabsf(r.ox) + fabsf(r.oy) + fabsf(r.oz) + ext);
  float tlim = tbest < 3e38 ? __double2float_ru(tbest) * 1.00001f + 1e-30f : FLT_MAX;
  const float4* __restrict__ nodes = sc.nodes + 2 * (size_t)B.node_base;
  int stack_n[KB_RAY_STACK]; float stack_t[KB_RAY_STACK];
  int sp = 0, elem = -1, rank = best.rank, id = -1;
  bool hit = false;
  float4 n0, n1; float tn;
  ld_node(nodes, 0, n0, n1);
  if (!slab(n0, n1, r, pad, tlim, tn)) return;
  for (;;) {
    // ---- node loop: nearer child first, the farther one deferred with its entry distance
    bool alive = true;
    int ref;
    while ((ref = __float_as_int(n0.w)) >= 0) {
      float4 a0, a1, b0, b1; float ta, tb;
      ld_node(nodes, (size_t)ref, a0, a1);
      ld_node(nodes, (size_t)ref + 1, b0, b1);
      const bool ha = slab(a0, a1, r, pad, tlim, ta), hb = slab(b0, b1, r, pad, tlim, tb);
      if (ha | hb) {
        const bool afirst = ha && (!hb || ta <= tb);
        if (ha & hb && sp < KB_RAY_STACK) { stack_n[sp] = afirst ? ref + 1 : ref; stack_t[sp] = afirst ? tb : ta; sp++; }
        if (afirst) { n0 = a0; n1 = a1; } else { n0 = b0; n1 = b1; }
        continue;
      }
      // both children missed: the nearest deferred subtree that can still beat the best hit
      alive = false;
      while (sp > 0) {
        sp--;
        if (stack_t[sp] - 2e-6f * fabsf(stack_t[sp]) <= tlim) { ld_node(nodes, (size_t)stack_n[sp], n0, n1); alive = true; break; }
      }
      if (!alive) break;
    }
    if (!alive) break;
    // ---- element loop of the leaf in n0 / n1
    {
      const int first = ~ref, cnt = __float_as_int(n1.w);
      if (mesh && !have_shear) { ray_shear(dl, sh); have_shear = true; }
      for (int i = 0; i < cnt; i++) {
        const int e = B.elem_base + first + i;
        double t; bool h;
        if (mesh) h = ray_tri(sl, sh, sc.tris64 + 9 * (size_t)e, t);
        else h = ray_sphere(sl, dl, sc.sph64 + 4 * (size_t)e, sc.sph64[4 * (size_t)e + 3] + B.margin, t);
        if (h && t <= tbest) {
          const int orig = mesh ? sc.triorig[e] : sc.sphorig[e];
          int own = B.id, rk = B.rank;
          if (grouped) {
            own = mesh ? sc.triown[e] : sc.sphown[e];
            if (ignore && ignore[own]) continue;
            rk = rank_of_owner(own, wc);
          }
          // nearer wins; at equal distance the body the reference visits first, then the lower element index
          if (t < tbest || rk < rank || (rk == rank && (!hit || orig < elem))) {
            tbest = t; elem = orig; rank = rk; id = own; hit = true;
            tlim = __double2float_ru(tbest) * 1.00001f + 1e-30f;
          }
        }
      }
    }
    bool got = false;
    while (sp > 0) {
      sp--;
      if (stack_t[sp] - 2e-6f * fabsf(stack_t[sp]) <= tlim) { ld_node(nodes, (size_t)stack_n[sp], n0, n1); got = true; break; }
    }
    if (!got) break;
  }
  if (!hit) return;
  best.d = tbest - shift; best.rank = rank; best.id = id; best.elem = elem;
}

template <int BLOCK, int MINB>
__global__ void __launch_bounds__(BLOCK, MINB)
kb_raycast_kernel(const KbRayParams p) {
  int64_t i = (int64_t)blockIdx.x * BLOCK + threadIdx.x;
  if (p.cam_on && p.tile) {
    // camera mode: a warp renders an 8 x 4 pixel tile instead of 32 pixels of one row (its rays share more of their path)
    const int64_t w = i >> 5; const int l = (int)(i & 31);
    const int tiles_x = (p.xres + 7) >> 3;
    const int64_t ty = w / tiles_x; const int tx = (int)(w - ty * tiles_x);
    const int px = tx * 8 + (l & 7); const int64_t py = ty * 4 + (l >> 3);
    if (px >= p.xres || py >= p.yres) return;
    i = py * p.xres + px;
  }
  if (i >= p.N) return;
  double s[3], d[3];
  if (p.cam_on) {
    // pixel (ii, jj): direction fwd + (ii - cx) dx + (cy - jj) dy, source zmin along that unnormalised vector (the reference's order)
    const int jj = (int)(i / p.xres), ii = (int)(i - (int64_t)jj * p.xres);
    const double u = (double)ii - p.cx, v = p.cy - (double)jj;
#pragma unroll
    for (int k = 0; k < 3; k++) {
      d[k] = __dadd_rn(__dadd_rn(p.fwd[k], __dmul_rn(u, p.dx[k])), __dmul_rn(v, p.dy[k]));
      s[k] = __dadd_rn(p.eye[k], __dmul_rn(d[k], p.zmin));
    }
  } else {
    const double* __restrict__ ray = p.rays + 6 * i;
    s[0] = ray[0]; s[1] = ray[1]; s[2] = ray[2]; d[0] = ray[3]; d[1] = ray[4]; d[2] = ray[5];
  }
  Best best; best.d = INFINITY; best.rank = 0x7fffffff; best.id = -1; best.elem = -1;
  WorldCounts wc; wc.nterr = p.nterr; wc.nobj = p.nobj; wc.nlinks = p.nlinks;
  const double n2 = __dadd_rn(__dadd_rn(__dmul_rn(d[0], d[0]), __dmul_rn(d[1], d[1])), __dmul_rn(d[2], d[2]));
  const bool ok = n2 > 0.0 && isfinite(n2) && isfinite(s[0]) && isfinite(s[1]) && isfinite(s[2]);
  if (ok) {
    const double n = __dsqrt_rn(n2);
    d[0] = __ddiv_rn(d[0], n); d[1] = __ddiv_rn(d[1], n); d[2] = __ddiv_rn(d[2], n);
    // ---- the robot's links in the frames of this call's configuration
    for (int b = 0; b < p.nlinkbodies; b++) {
      const KbRayBody& B = p.bodies[b];
      if (B.id >= 0 && p.ignore && p.ignore[B.id]) continue;
      cast_body(p.scene, B, B.xf >= 0 ? p.xf64 + 12 * (size_t)B.xf : (B.has_T ? B.T : nullptr), s, d, best, p.ignore, wc);
    }
    // ---- static bodies under the top-level hierarchy (world boxes, already grown by each body's margin)
    if (p.nstatic > 0) {
      RayF r;
      r.ox = (float)s[0]; r.oy = (float)s[1]; r.oz = (float)s[2];
      const float fx = (float)d[0], fy = (float)d[1], fz = (float)d[2];
      r.ix = 1.f / (fabsf(fx) > 1e-30f ? fx : copysignf(1e-30f, fx));
      r.iy = 1.f / (fabsf(fy) > 1e-30f ? fy : copysignf(1e-30f, fy));
      r.iz = 1.f / (fabsf(fz) > 1e-30f ? fz : copysignf(1e-30f, fz));
      const float pad = 8.f * p.scene.eps_abs + 4e-6f * (fabsf(r.ox) + fabsf(r.oy) + fabsf(r.oz) + p.tlas_ext);
      int tstack[KB_RAY_TSTACK]; int tsp = 0;
      tstack[tsp++] = 0;
      while (tsp > 0) {
        const int node = tstack[--tsp];
        float4 n0, n1; float tn;
        ld_node(p.tlas, (size_t)node, n0, n1);
        const double lim = best.d + p.max_margin;
        const float tlim = lim < 3e38 ? __double2float_ru(lim) * 1.00001f + 1e-30f : FLT_MAX;
        if (!slab(n0, n1, r, pad, tlim, tn)) continue;
        const int ref = __float_as_int(n0.w);
        if (ref < 0) {
          const int first = ~ref, cnt = __float_as_int(n1.w);
          for (int k = 0; k < cnt; k++) {
            const KbRayBody& B = p.bodies[p.nlinkbodies + first + k];
            if (B.id >= 0 && p.ignore && p.ignore[B.id]) continue;
            cast_body(p.scene, B, B.has_T ? B.T : nullptr, s, d, best, p.ignore, wc);
          }
        } else if (tsp + 2 <= KB_RAY_TSTACK) {
          // nearer child on top: its hits shorten the ray before the farther one is looked at
          float4 a0, a1, b0, b1; float ta, tb;
          ld_node(p.tlas, (size_t)ref, a0, a1); ld_node(p.tlas, (size_t)ref + 1, b0, b1);
          slab(a0, a1, r, pad, tlim, ta); slab(b0, b1, r, pad, tlim, tb);
          if (ta <= tb) { tstack[tsp++] = ref + 1; tstack[tsp++] = ref; } else { tstack[tsp++] = ref; tstack[tsp++] = ref + 1; }
        }
      }
    }
  }
  if (p.out_id) p.out_id[i] = best.id;
  if (p.out_dist) p.out_dist[i] = best.d;
  if (p.out_elem) p.out_elem[i] = best.elem;
  if (p.out_depth) {
    // depth along the viewing direction: fwd . (pt - eye); below zmin = not sensed, beyond zmax = zmax (VisualSensors.cpp:461-466)
    double z = p.zmax;
    if (best.id >= 0) {
      z = 0.0;
#pragma unroll
      for (int k = 0; k < 3; k++) z += p.fwd[k] * (s[k] + best.d * d[k] - p.eye[k]);
      z = z < p.zmin ? p.zmax : fmin(z, p.zmax);
    }
    p.out_depth[i] = (float)z;
  }
}

}  // namespace

// 64-thread blocks, 8 per SM: measured best of {128 x 4, 128 x 5 / 6 (spills), 64 x 8, 64 x 10 (spills), 32 x 16} on 640 x 480 images of the C2
// world (profiles/r02_experiments.md); variant 1 keeps the 128 x 4 shape for comparison
cudaError_t kb_launch_raycast(const KbRayParams& p, cudaStream_t s, int variant) {
  if (p.N <= 0) return cudaSuccess;
  int64_t threads = p.N;
  if (p.cam_on && p.tile) threads = (int64_t)((p.xres + 7) / 8) * ((p.yres + 3) / 4) * 32;
  if (variant == 1) kb_raycast_kernel<128, 4><<<(unsigned)((threads + 127) / 128), 128, 0, s>>>(p);
  else kb_raycast_kernel<64, 8><<<(unsigned)((threads + 63) / 64), 64, 0, s>>>(p);
  return cudaGetLastError();
}
