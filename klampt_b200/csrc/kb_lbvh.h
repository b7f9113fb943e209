// kb_lbvh.h -- GPU linear-BVH builder for replaceable environment point clouds (kb_lbvh.cu)
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

// nodes (32 B each) a cloud of up to `capacity` points can need, and the scratch the builder wants for it
size_t kb_lbvh_nodes_for(int capacity);
size_t kb_lbvh_scratch_bytes(int capacity);
// Builds the hierarchy of n points given in the geometry's local frame (device pointers; radius may be null = uniform_radius):
// transforms them by d_T12 (row-major 3x3 + translation, device), writes them in BVH order to sph64 / sph32 / sphown (= owner) and the
// nodes to `nodes`.  If h_maxabs is not null the stream is synchronised and the largest |world coordinate| is returned.
// d_owner_in (may be null = `owner` for every point): owner id per input point, carried through the sort.
cudaError_t kb_lbvh_build(const double* d_pts_local, const double* d_radius, double uniform_radius, int n, const double* d_T12, int owner,
                          const int32_t* d_owner_in, double* sph64, float4* sph32, int32_t* sphown, float4* nodes, void* scratch, size_t scratch_bytes, int capacity, float* h_maxabs,
                          cudaStream_t s, const int32_t* d_orig_in = nullptr, int32_t* orig_out = nullptr);   // orig_out[i] = caller's index of element i (via d_orig_in if given)
// triangle meshes (9 doubles per triangle, already in the hierarchy's frame): keys from the centroids, one triangle per leaf
cudaError_t kb_lbvh_build_tris(const double* d_tris_in, int n, int owner, const int32_t* d_owner_in, double* tris64, float4* tris32, int32_t* triown,
                               float4* nodes, void* scratch, size_t scratch_bytes, int capacity, cudaStream_t s, const int32_t* d_orig_in = nullptr, int32_t* orig_out = nullptr);
