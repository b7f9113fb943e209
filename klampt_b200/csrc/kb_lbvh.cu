// kb_lbvh.cu -- GPU construction of a point-cloud BVH (linear BVH: Morton order + Karras' parallel hierarchy), for
// environment point clouds that are replaced between batches (sensor streams; the reference keeps such geometries in
// ManagedGeometry as "dynamic" geometries, Cpp/Modeling/ManagedGeometry.h:49-52, and rebuilds their collision data on the
// CPU -- Cpp/docs/Manual-Geometry.md:157-162 quotes 504 ms for a 70 k-triangle mesh).
//
// Output = the engine's node format (kb_types.h): 32-byte nodes {centre.xyz, left}{half.xyz, count}, siblings adjacent and
// even-aligned.  The Karras tree is built over the single points; its internal nodes with more than KB_LBVH_LEAF points are
// kept, smaller subtrees collapse into leaves.  Kept node number r (dense, by prefix sum) owns the pair slot (2 r + 2, 2 r + 3)
// for its two children; the root sits at 0 and slot 1 is padding.
//
//   1. points local -> world (fp64, same operation order as a scalar R p + t), bounds by warp + atomic reduction
//   2. 30-bit Morton keys, cub::DeviceRadixSort::SortPairs, gather into BVH order (sph64 / sph32 / owner id)
//   3. point boxes
//   4. Karras 2012: one thread per internal node finds its range and split from the common-prefix lengths of the keys
//   5. bottom-up boxes: each point climbs, the second thread to reach a node merges its children's boxes
//   6. cub::DeviceScan numbers the kept nodes; every kept node writes the two records of its pair slot
#include "kb_lbvh.h"
#include <cub/cub.cuh>
#include <float.h>
#include <limits.h>
#include <string.h>

#define KB_LBVH_LEAF 8

namespace {

__device__ __forceinline__ unsigned expand10(unsigned v) {       // 10 bits -> every third bit
  v = (v * 0x00010001u) & 0xFF0000FFu; v = (v * 0x00000101u) & 0x0F00F00Fu; v = (v * 0x00000011u) & 0xC30C30C3u; v = (v * 0x00000005u) & 0x49249249u;
  return v;
}
__device__ __forceinline__ int f2ord(float f) { int i = __float_as_int(f); return i >= 0 ? i : i ^ 0x7fffffff; }   // order-preserving
__device__ __forceinline__ float ord2f(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); }

// 1. transform + bounds.  bounds[0..2] = min, [3..5] = max as order-preserving ints
__global__ void lbvh_transform_kernel(const double* __restrict__ pts, const double* __restrict__ radius, double uniform_r, int n, const double* __restrict__ T12,
                                      double* __restrict__ w64, int* __restrict__ bounds) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
  if (i < n) {
    const double x = pts[3 * (size_t)i], y = pts[3 * (size_t)i + 1], z = pts[3 * (size_t)i + 2];
    double w[3];
#pragma unroll
    for (int k = 0; k < 3; k++)
      w[k] = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(T12[3 * k], x), __dmul_rn(T12[3 * k + 1], y)), __dmul_rn(T12[3 * k + 2], z)), T12[9 + k]);
    const double r = radius ? radius[i] : uniform_r;
    w64[4 * (size_t)i] = w[0]; w64[4 * (size_t)i + 1] = w[1]; w64[4 * (size_t)i + 2] = w[2]; w64[4 * (size_t)i + 3] = r;
#pragma unroll
    for (int k = 0; k < 3; k++) { lo[k] = __double2float_rd(w[k]); hi[k] = __double2float_ru(w[k]); }
  }
#pragma unroll
  for (int k = 0; k < 3; k++) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { lo[k] = fminf(lo[k], __shfl_xor_sync(0xffffffffu, lo[k], o)); hi[k] = fmaxf(hi[k], __shfl_xor_sync(0xffffffffu, hi[k], o)); }
  }
  if ((threadIdx.x & 31) == 0) {
#pragma unroll
    for (int k = 0; k < 3; k++) { atomicMin(bounds + k, f2ord(lo[k])); atomicMax(bounds + 3 + k, f2ord(hi[k])); }
  }
}

// 2a. Morton keys of the world points inside the bounds
__global__ void lbvh_morton_kernel(const double* __restrict__ w64, int n, const int* __restrict__ bounds, unsigned* __restrict__ keys, int* __restrict__ idx) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  unsigned code = 0;
#pragma unroll
  for (int k = 0; k < 3; k++) {
    const float lo = ord2f(bounds[k]), hi = ord2f(bounds[3 + k]);
    const float ext = fmaxf(hi - lo, 1e-30f);
    float u = ((float)w64[4 * (size_t)i + k] - lo) / ext * 1024.f;
    u = fminf(fmaxf(u, 0.f), 1023.f);
    code |= expand10((unsigned)u) << (2 - k);
  }
  keys[i] = code; idx[i] = i;
}

// 2b. gather into BVH order
__global__ void lbvh_gather_kernel(const double* __restrict__ w64, const int* __restrict__ sorted_idx, int n, int owner, const int32_t* __restrict__ owner_in,
                                   double* __restrict__ sph64, float4* __restrict__ sph32, int* __restrict__ sphown,
                                   const int32_t* __restrict__ orig_in, int32_t* __restrict__ orig_out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int s = sorted_idx[i];
  const double x = w64[4 * (size_t)s], y = w64[4 * (size_t)s + 1], z = w64[4 * (size_t)s + 2], r = w64[4 * (size_t)s + 3];
  sph64[4 * (size_t)i] = x; sph64[4 * (size_t)i + 1] = y; sph64[4 * (size_t)i + 2] = z; sph64[4 * (size_t)i + 3] = r;
  sph32[i] = make_float4((float)x, (float)y, (float)z, (float)r);
  sphown[i] = owner_in ? owner_in[s] : owner;
  if (orig_out) orig_out[i] = orig_in ? orig_in[s] : s;      // index of the point in the caller's array
}

// 3. point boxes (fp64 extents rounded outwards to fp32): the leaves of the Karras tree are the single points
__global__ void lbvh_pointbox_kernel(const double* __restrict__ sph64, int n, float* __restrict__ blo, float* __restrict__ bhi) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double r = sph64[4 * (size_t)i + 3];
#pragma unroll
  for (int k = 0; k < 3; k++) { const double c = sph64[4 * (size_t)i + k]; blo[3 * (size_t)i + k] = __double2float_rd(c - r); bhi[3 * (size_t)i + k] = __double2float_ru(c + r); }
}

// common-prefix length of points i and j of the sorted order; equal keys are told apart by their index (Karras 2012, section 4)
__device__ __forceinline__ int lbvh_delta(const unsigned* __restrict__ k, int n, int i, int j) {
  if (j < 0 || j >= n) return -1;
  const unsigned a = k[i], b = k[j];
  if (a == b) return 32 + __clz((unsigned)i ^ (unsigned)j);
  return __clz(a ^ b);
}

// 4. hierarchy over the n points.  child >= 0: internal node index; child < 0: ~point index.  An internal node whose range holds
// more than KB_LBVH_LEAF points is KEPT as a node of the emitted tree; smaller ones become its leaves.  Because the split is always
// at the highest differing key bit, a leaf never straddles a jump of the Morton curve (consecutive-run clustering does: its
// worst leaves span metres and the traversal pays for them -- measured 20-45x slower queries).
__global__ void lbvh_hierarchy_kernel(const unsigned* __restrict__ keys, int n, int* __restrict__ childL, int* __restrict__ childR,
                                      int* __restrict__ parentI, int* __restrict__ parentLeaf, int* __restrict__ first, int* __restrict__ count,
                                      int* __restrict__ keep, int leaf_max) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n - 1) return;
  const int d = (lbvh_delta(keys, n, i, i + 1) - lbvh_delta(keys, n, i, i - 1)) >= 0 ? 1 : -1;
  const int dmin = lbvh_delta(keys, n, i, i - d);
  int lmax = 2;
  while (lbvh_delta(keys, n, i, i + lmax * d) > dmin) lmax <<= 1;
  int l = 0;
  for (int t = lmax >> 1; t >= 1; t >>= 1) if (lbvh_delta(keys, n, i, i + (l + t) * d) > dmin) l += t;
  const int j = i + l * d;
  const int dnode = lbvh_delta(keys, n, i, j);
  int s = 0;
  for (int t = (l + 1) >> 1; ; t = (t + 1) >> 1) {
    if (lbvh_delta(keys, n, i, i + (s + t) * d) > dnode) s += t;
    if (t == 1) break;
  }
  const int gamma = i + s * d + min(d, 0);
  const int lo = min(i, j), hi = max(i, j);
  const int cl = (lo == gamma) ? ~gamma : gamma;
  const int cr = (hi == gamma + 1) ? ~(gamma + 1) : gamma + 1;
  childL[i] = cl; childR[i] = cr;
  if (cl >= 0) parentI[cl] = i; else parentLeaf[~cl] = i;
  if (cr >= 0) parentI[cr] = i; else parentLeaf[~cr] = i;
  if (i == 0) parentI[0] = -1;
  first[i] = lo; count[i] = hi - lo + 1;
  keep[i] = (hi - lo + 1) > leaf_max ? 1 : 0;
}

// 5. bottom-up boxes of the internal nodes
__global__ void lbvh_refit_kernel(int n, const int* __restrict__ childL, const int* __restrict__ childR, const int* __restrict__ parentI,
                                  const int* __restrict__ parentLeaf, const float* __restrict__ llo, const float* __restrict__ lhi,
                                  float* __restrict__ ilo, float* __restrict__ ihi, int* __restrict__ flags) {
  const int l = blockIdx.x * blockDim.x + threadIdx.x;
  if (l >= n) return;
  int node = parentLeaf[l];
  while (node >= 0) {
    if (atomicAdd(flags + node, 1) == 0) return;          // the first child to arrive leaves; the second one merges
    __threadfence();
    const int a = childL[node], b = childR[node];
#pragma unroll
    for (int k = 0; k < 3; k++) {
      const float alo = a >= 0 ? __ldcg(ilo + 3 * (size_t)a + k) : llo[3 * (size_t)(~a) + k], blo = b >= 0 ? __ldcg(ilo + 3 * (size_t)b + k) : llo[3 * (size_t)(~b) + k];
      const float ahi = a >= 0 ? __ldcg(ihi + 3 * (size_t)a + k) : lhi[3 * (size_t)(~a) + k], bhi = b >= 0 ? __ldcg(ihi + 3 * (size_t)b + k) : lhi[3 * (size_t)(~b) + k];
      ilo[3 * (size_t)node + k] = fminf(alo, blo); ihi[3 * (size_t)node + k] = fmaxf(ahi, bhi);
    }
    __threadfence();
    node = parentI[node];
  }
}

__device__ __forceinline__ void lbvh_write_node(float4* __restrict__ nodes, int pos, const float* lo, const float* hi, int left, int count) {
  float c[3], h[3];
#pragma unroll
  for (int k = 0; k < 3; k++) {
    c[k] = 0.5f * (lo[k] + hi[k]);
    h[k] = fmaxf(__fsub_ru(hi[k], c[k]), __fsub_ru(c[k], lo[k]));      // the centre/half box contains [lo, hi]
  }
  nodes[2 * (size_t)pos] = make_float4(c[0], c[1], c[2], __int_as_float(left));
  nodes[2 * (size_t)pos + 1] = make_float4(h[0], h[1], h[2], __int_as_float(count));
}

// 6. emit: kept node i (dense number rank[i]) writes its children at 2 rank + 2 and 2 rank + 3; thread 0 also writes the root and
// the padding node.  A child that is a single point or a dropped (small) internal node becomes a leaf record.
__global__ void lbvh_emit_kernel(int n, const int* __restrict__ childL, const int* __restrict__ childR, const int* __restrict__ first,
                                 const int* __restrict__ count, const int* __restrict__ keep, const int* __restrict__ rank, const float* __restrict__ llo,
                                 const float* __restrict__ lhi, const float* __restrict__ ilo, const float* __restrict__ ihi, float4* __restrict__ nodes) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0) {
    if (n == 0) {      // empty cloud: a leaf without elements whose box overlaps nothing
      nodes[0] = make_float4(0.f, 0.f, 0.f, __int_as_float(~0)); nodes[1] = make_float4(-1e30f, -1e30f, -1e30f, __int_as_float(0));
    } else if (n == 1) lbvh_write_node(nodes, 0, llo, lhi, ~0, 1);
    else if (!keep[0]) lbvh_write_node(nodes, 0, ilo, ihi, ~0, n);           // the whole cloud fits one leaf
    else lbvh_write_node(nodes, 0, ilo, ihi, 2, 0);
    nodes[2] = make_float4(0.f, 0.f, 0.f, __int_as_float(~0)); nodes[3] = make_float4(-1e30f, -1e30f, -1e30f, __int_as_float(0));
  }
  if (i >= n - 1 || !keep[i]) return;
  const int slot = 2 * rank[i] + 2;
#pragma unroll
  for (int side = 0; side < 2; side++) {
    const int c = side ? childR[i] : childL[i];
    const int pos = slot + side;
    if (c < 0) lbvh_write_node(nodes, pos, llo + 3 * (size_t)(~c), lhi + 3 * (size_t)(~c), ~(~c), 1);              // one point: first = its index
    else if (keep[c]) lbvh_write_node(nodes, pos, ilo + 3 * (size_t)c, ihi + 3 * (size_t)c, 2 * rank[c] + 2, 0);
    else lbvh_write_node(nodes, pos, ilo + 3 * (size_t)c, ihi + 3 * (size_t)c, ~first[c], count[c]);
  }
}

// ---- triangle meshes: keys from the centroids, one triangle per leaf
__global__ void lbvh_tri_bounds_kernel(const double* __restrict__ tris, int n, int* __restrict__ bounds) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
  if (i < n) {
#pragma unroll
    for (int k = 0; k < 3; k++) {
      const double c = (tris[9 * (size_t)i + k] + tris[9 * (size_t)i + 3 + k] + tris[9 * (size_t)i + 6 + k]) * (1.0 / 3.0);
      lo[k] = __double2float_rd(c); hi[k] = __double2float_ru(c);
    }
  }
#pragma unroll
  for (int k = 0; k < 3; k++) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { lo[k] = fminf(lo[k], __shfl_xor_sync(0xffffffffu, lo[k], o)); hi[k] = fmaxf(hi[k], __shfl_xor_sync(0xffffffffu, hi[k], o)); }
  }
  if ((threadIdx.x & 31) == 0) {
#pragma unroll
    for (int k = 0; k < 3; k++) { atomicMin(bounds + k, f2ord(lo[k])); atomicMax(bounds + 3 + k, f2ord(hi[k])); }
  }
}
__global__ void lbvh_tri_morton_kernel(const double* __restrict__ tris, int n, const int* __restrict__ bounds, unsigned* __restrict__ keys, int* __restrict__ idx) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  unsigned code = 0;
#pragma unroll
  for (int k = 0; k < 3; k++) {
    const float lo = ord2f(bounds[k]), hi = ord2f(bounds[3 + k]);
    const float ext = fmaxf(hi - lo, 1e-30f);
    const double c = (tris[9 * (size_t)i + k] + tris[9 * (size_t)i + 3 + k] + tris[9 * (size_t)i + 6 + k]) * (1.0 / 3.0);
    float u = ((float)c - lo) / ext * 1024.f;
    u = fminf(fmaxf(u, 0.f), 1023.f);
    code |= expand10((unsigned)u) << (2 - k);
  }
  keys[i] = code; idx[i] = i;
}
// gather into BVH order (tris64, tris32 with the owner id in .w of vertex 0, triown) and the triangle boxes
__global__ void lbvh_tri_gather_kernel(const double* __restrict__ tris, const int* __restrict__ sorted_idx, int n, int owner, const int32_t* __restrict__ owner_in,
                                       double* __restrict__ tris64, float4* __restrict__ tris32, int* __restrict__ triown,
                                       float* __restrict__ blo, float* __restrict__ bhi, const int32_t* __restrict__ orig_in, int32_t* __restrict__ orig_out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int s = sorted_idx[i];
  const int own = owner_in ? owner_in[s] : owner;
  if (orig_out) orig_out[i] = orig_in ? orig_in[s] : s;
  double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
#pragma unroll
  for (int v = 0; v < 3; v++) {
    const double x = tris[9 * (size_t)s + 3 * v], y = tris[9 * (size_t)s + 3 * v + 1], z = tris[9 * (size_t)s + 3 * v + 2];
    tris64[9 * (size_t)i + 3 * v] = x; tris64[9 * (size_t)i + 3 * v + 1] = y; tris64[9 * (size_t)i + 3 * v + 2] = z;
    tris32[3 * (size_t)i + v] = make_float4((float)x, (float)y, (float)z, v == 0 ? __int_as_float(own) : 0.f);
    lo[0] = fmin(lo[0], x); lo[1] = fmin(lo[1], y); lo[2] = fmin(lo[2], z); hi[0] = fmax(hi[0], x); hi[1] = fmax(hi[1], y); hi[2] = fmax(hi[2], z);
  }
  triown[i] = own;
#pragma unroll
  for (int k = 0; k < 3; k++) { blo[3 * (size_t)i + k] = __double2float_rd(lo[k]); bhi[3 * (size_t)i + k] = __double2float_ru(hi[k]); }
}

inline unsigned nb(int n, int bs) { return (unsigned)((n + bs - 1) / bs); }

}  // namespace

// worst case: every kept node has a one-point child (points on a line with geometric spacing): n - KB_LBVH_LEAF kept nodes
size_t kb_lbvh_nodes_for(int capacity) { return 2 * (size_t)capacity + 4; }

static size_t lbvh_sort_tmp(int capacity) {
  size_t a = 0, b = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, a, (const unsigned*)nullptr, (unsigned*)nullptr, (const int*)nullptr, (int*)nullptr, capacity, 0, 30);
  cub::DeviceScan::ExclusiveSum(nullptr, b, (const int*)nullptr, (int*)nullptr, capacity);
  return a > b ? a : b;
}

size_t kb_lbvh_scratch_bytes(int capacity) {
  const size_t n = (size_t)capacity + 1;
  size_t b = 256;                      // bounds
  b += n * 32 + 256;                   // world points
  b += 4 * (n * 4 + 256);              // keys, idx, sorted keys, sorted idx
  b += 4 * (n * 12 + 256);             // point / internal boxes
  b += 9 * (n * 4 + 256);              // children, parents, flags, first, count, keep, rank
  b += lbvh_sort_tmp(capacity) + 4096;
  return b;
}

cudaError_t kb_lbvh_build(const double* d_pts_local, const double* d_radius, double uniform_radius, int n, const double* d_T12, int owner,
                          const int32_t* d_owner_in, double* sph64, float4* sph32, int32_t* sphown, float4* nodes, void* scratch, size_t scratch_bytes, int capacity, float* h_maxabs,
                          cudaStream_t s, const int32_t* d_orig_in, int32_t* orig_out) {
  if (n < 0 || n > capacity) return cudaErrorInvalidValue;
  if (scratch_bytes < kb_lbvh_scratch_bytes(capacity)) return cudaErrorInvalidValue;
  const size_t cap = (size_t)capacity + 1;
  unsigned char* p = (unsigned char*)scratch;
  auto take = [&](size_t bytes) { void* r = p; p += (bytes + 255) & ~(size_t)255; return r; };
  int* bounds = (int*)take(64);
  double* w64 = (double*)take(cap * 32);
  unsigned* keys = (unsigned*)take(cap * 4); int* idx = (int*)take(cap * 4);
  unsigned* skeys = (unsigned*)take(cap * 4); int* sidx = (int*)take(cap * 4);
  float* llo = (float*)take(cap * 12); float* lhi = (float*)take(cap * 12); float* ilo = (float*)take(cap * 12); float* ihi = (float*)take(cap * 12);
  int* childL = (int*)take(cap * 4); int* childR = (int*)take(cap * 4); int* parentI = (int*)take(cap * 4); int* parentLeaf = (int*)take(cap * 4); int* flags = (int*)take(cap * 4);
  int* first = (int*)take(cap * 4); int* count = (int*)take(cap * 4); int* keep = (int*)take(cap * 4); int* rank = (int*)take(cap * 4);
  size_t tmp_bytes = lbvh_sort_tmp(capacity);
  void* tmp = take(tmp_bytes);
  if ((size_t)(p - (unsigned char*)scratch) > scratch_bytes) return cudaErrorInvalidValue;

  const int init[6] = {0x7f7fffff, 0x7f7fffff, 0x7f7fffff, INT_MIN, INT_MIN, INT_MIN};   // +FLT_MAX / -FLT_MAX in the order-preserving int form
  cudaError_t e = cudaMemcpyAsync(bounds, init, sizeof init, cudaMemcpyHostToDevice, s);
  if (e != cudaSuccess) return e;
  if (n > 0) {
    lbvh_transform_kernel<<<nb(n, 256), 256, 0, s>>>(d_pts_local, d_radius, uniform_radius, n, d_T12, w64, bounds);
    lbvh_morton_kernel<<<nb(n, 256), 256, 0, s>>>(w64, n, bounds, keys, idx);
    size_t tb = tmp_bytes;
    e = cub::DeviceRadixSort::SortPairs(tmp, tb, keys, skeys, idx, sidx, n, 0, 30, s);
    if (e != cudaSuccess) return e;
    lbvh_gather_kernel<<<nb(n, 256), 256, 0, s>>>(w64, sidx, n, owner, d_owner_in, sph64, sph32, sphown, d_orig_in, orig_out);
    lbvh_pointbox_kernel<<<nb(n, 256), 256, 0, s>>>(sph64, n, llo, lhi);
    if (n > 1) {
      e = cudaMemsetAsync(flags, 0, (size_t)n * 4, s);
      if (e != cudaSuccess) return e;
      lbvh_hierarchy_kernel<<<nb(n - 1, 256), 256, 0, s>>>(skeys, n, childL, childR, parentI, parentLeaf, first, count, keep, KB_LBVH_LEAF);
      lbvh_refit_kernel<<<nb(n, 256), 256, 0, s>>>(n, childL, childR, parentI, parentLeaf, llo, lhi, ilo, ihi, flags);
      tb = tmp_bytes;
      e = cub::DeviceScan::ExclusiveSum(tmp, tb, keep, rank, n - 1, s);      // dense numbering of the kept nodes (the root is 0)
      if (e != cudaSuccess) return e;
    }
  }
  lbvh_emit_kernel<<<nb(n > 1 ? n - 1 : 1, 256), 256, 0, s>>>(n, childL, childR, first, count, keep, rank, llo, lhi, ilo, ihi, nodes);
  e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  if (h_maxabs) {       // largest |coordinate| of the cloud: the caller folds it into the scene's fp32 error bound
    int hb[6];
    e = cudaMemcpyAsync(hb, bounds, sizeof hb, cudaMemcpyDeviceToHost, s);
    if (e != cudaSuccess) return e;
    e = cudaStreamSynchronize(s);
    if (e != cudaSuccess) return e;
    float m = 0.f;
    if (n > 0) for (int k = 0; k < 6; k++) { const int o = hb[k]; float f; const int bits = o >= 0 ? o : o ^ 0x7fffffff; memcpy(&f, &bits, 4); m = fmaxf(m, fabsf(f)); }
    *h_maxabs = m;
  }
  return cudaSuccess;
}

// the same for a triangle mesh (triangles already in the frame of the hierarchy, 9 doubles each): one triangle per leaf
cudaError_t kb_lbvh_build_tris(const double* d_tris_in, int n, int owner, const int32_t* d_owner_in, double* tris64, float4* tris32, int32_t* triown,
                               float4* nodes, void* scratch, size_t scratch_bytes, int capacity, cudaStream_t s, const int32_t* d_orig_in, int32_t* orig_out) {
  if (n < 0 || n > capacity) return cudaErrorInvalidValue;
  if (scratch_bytes < kb_lbvh_scratch_bytes(capacity)) return cudaErrorInvalidValue;
  const size_t cap = (size_t)capacity + 1;
  unsigned char* p = (unsigned char*)scratch;
  auto take = [&](size_t bytes) { void* r = p; p += (bytes + 255) & ~(size_t)255; return r; };
  int* bounds = (int*)take(64);
  (void)take(cap * 32);
  unsigned* keys = (unsigned*)take(cap * 4); int* idx = (int*)take(cap * 4);
  unsigned* skeys = (unsigned*)take(cap * 4); int* sidx = (int*)take(cap * 4);
  float* llo = (float*)take(cap * 12); float* lhi = (float*)take(cap * 12); float* ilo = (float*)take(cap * 12); float* ihi = (float*)take(cap * 12);
  int* childL = (int*)take(cap * 4); int* childR = (int*)take(cap * 4); int* parentI = (int*)take(cap * 4); int* parentLeaf = (int*)take(cap * 4); int* flags = (int*)take(cap * 4);
  int* first = (int*)take(cap * 4); int* count = (int*)take(cap * 4); int* keep = (int*)take(cap * 4); int* rank = (int*)take(cap * 4);
  size_t tmp_bytes = lbvh_sort_tmp(capacity);
  void* tmp = take(tmp_bytes);
  if ((size_t)(p - (unsigned char*)scratch) > scratch_bytes) return cudaErrorInvalidValue;
  const int init[6] = {0x7f7fffff, 0x7f7fffff, 0x7f7fffff, INT_MIN, INT_MIN, INT_MIN};
  cudaError_t e = cudaMemcpyAsync(bounds, init, sizeof init, cudaMemcpyHostToDevice, s);
  if (e != cudaSuccess) return e;
  if (n > 0) {
    lbvh_tri_bounds_kernel<<<nb(n, 256), 256, 0, s>>>(d_tris_in, n, bounds);
    lbvh_tri_morton_kernel<<<nb(n, 256), 256, 0, s>>>(d_tris_in, n, bounds, keys, idx);
    size_t tb = tmp_bytes;
    e = cub::DeviceRadixSort::SortPairs(tmp, tb, keys, skeys, idx, sidx, n, 0, 30, s);
    if (e != cudaSuccess) return e;
    lbvh_tri_gather_kernel<<<nb(n, 256), 256, 0, s>>>(d_tris_in, sidx, n, owner, d_owner_in, tris64, tris32, triown, llo, lhi, d_orig_in, orig_out);
    if (n > 1) {
      e = cudaMemsetAsync(flags, 0, (size_t)n * 4, s);
      if (e != cudaSuccess) return e;
      lbvh_hierarchy_kernel<<<nb(n - 1, 256), 256, 0, s>>>(skeys, n, childL, childR, parentI, parentLeaf, first, count, keep, 1);
      lbvh_refit_kernel<<<nb(n, 256), 256, 0, s>>>(n, childL, childR, parentI, parentLeaf, llo, lhi, ilo, ihi, flags);
      tb = tmp_bytes;
      e = cub::DeviceScan::ExclusiveSum(tmp, tb, keep, rank, n - 1, s);
      if (e != cudaSuccess) return e;
    }
  }
  lbvh_emit_kernel<<<nb(n > 1 ? n - 1 : 1, 256), 256, 0, s>>>(n, childL, childR, first, count, keep, rank, llo, lhi, ilo, ihi, nodes);
  return cudaGetLastError();
}
