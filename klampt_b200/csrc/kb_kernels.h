// kb_kernels.h -- host-side launch wrappers of the sm_100a kernels in kb_kernels.cu
#pragma once
#include "kb_types.h"
#include <cuda_runtime.h>

size_t kb_traverse_smem_bytes(int nxf, int nitems, int nprobes);

cudaError_t kb_launch_fk(const KbRobotDev* robot, const KbDriverDev* drv, const int32_t* drv_link, const double* drv_scale,
                         const double* drv_off, const double* Q, int64_t N, double* xf64, int nxf, uint8_t* state,
                         const uint8_t* alive, int32_t* hit, cudaStream_t s);
// mode 0: boolean collide, mode 1: branch-and-bound distance (out_dist, upper_bound)
// out_cp (mode 1, optional): closest points, 6 doubles per configuration; rel_err / abs_err: tolerances of AnyCollisionQuery::Distance
cudaError_t kb_launch_traverse(const KbTraverseParams& p, int mode, double* out_dist, double upper_bound, int num_sms, cudaStream_t s,
                               double* out_cp = nullptr, float rel_err = 0.f, float abs_err = 0.f);
cudaError_t kb_launch_finish(const uint8_t* state, const int32_t* hit, const int32_t* hit_elem, const KbItem* items, const int32_t* triown,
                             const int32_t* sphown, const int32_t* boxown, int64_t N, uint8_t* out, int32_t* first_pair, unsigned long long* nfeasible, cudaStream_t s);
cudaError_t kb_launch_pair_ids(const int32_t* hit, const int32_t* hit_elem, const KbItem* items, const int32_t* triown, const int32_t* sphown, const int32_t* boxown,
                               int64_t N, int32_t* pair, cudaStream_t s);
cudaError_t kb_launch_edge_setup(const KbRobotDev* robot, const double* A, const double* B, const double* w, int64_t N, double eps,
                                 int32_t* nlev, uint8_t* alive, int32_t* nchecks, int32_t* maxlev, cudaStream_t s);
cudaError_t kb_launch_edge_count(const int32_t* nlev, const uint8_t* alive, int64_t N, int lev, int32_t* list, unsigned int* count, cudaStream_t s);
cudaError_t kb_launch_edge_expand(const KbRobotDev* robot, const double* A, const double* B, const int32_t* list, int64_t first_slot,
                                  int64_t nslots, int lev, double* Q, cudaStream_t s);
cudaError_t kb_launch_edge_reduce(const uint8_t* feas, const int32_t* list, int64_t first_slot, int64_t nslots, int lev, int32_t* firstbad, cudaStream_t s);
cudaError_t kb_launch_edge_level_end(const int32_t* list, unsigned int nlist, int lev, int32_t* firstbad, uint8_t* alive, int32_t* nchecks, cudaStream_t s);
cudaError_t kb_launch_pack_bits(const uint8_t* src, int64_t n, uint32_t* dst, cudaStream_t s);   // dst: (n + 31) / 32 words
cudaError_t kb_launch_fill_i32(int32_t* p, int64_t n, int32_t v, cudaStream_t s);
cudaError_t kb_launch_widen_f32(const float* src, double* dst, int64_t n, cudaStream_t s);
cudaError_t kb_launch_copy_u8(const uint8_t* src, uint8_t* dst, int64_t n, unsigned long long* ones, cudaStream_t s);

size_t kb_nodes_smem_bytes(int nxf, int nitems, int stack_cap);
cudaError_t kb_launch_split(const KbTraverseParams& p, const KbSplitParams& q, int num_sms, cudaStream_t s);
// every colliding world-id pair per configuration, up to max_pairs (<= 32); out_count = -1 where state == 0
cudaError_t kb_launch_allpairs(const KbTraverseParams& p, int max_pairs, int32_t* out_pairs, int32_t* out_count, int num_sms, cudaStream_t s);

// closest points (world frame, on the margin-inflated surfaces; 6 doubles: the point of the first reported id, then the second) and
// element indices (in the geometries' own element order) of the pair each configuration's distance query ended with (kb_closest.cu)
cudaError_t kb_launch_closest_points(const KbScene& sc, const KbItem* items, const double* xf64, int nxf, const int32_t* hit, const int32_t* hit_elem,
                                     int64_t N, double* out_cp, int32_t* out_elem, cudaStream_t s);


// small edge batches: all midpoints of all levels in one batch (slot = edge * per_max + sequential midpoint index)
cudaError_t kb_launch_edge_flat_expand(const KbRobotDev* robot, const double* A, const double* B, const int32_t* nlev, const uint8_t* alive, int64_t nslots, int per_max,
                                       double* Q, uint8_t* slot_on, unsigned long long* nactive, cudaStream_t s);
cudaError_t kb_launch_edge_flat_finish(const uint8_t* feas, const uint8_t* slot_on, int64_t nslots, int per_max, const int32_t* nlev, int32_t* firstbad, int64_t N,
                                       uint8_t* alive, int32_t* nchecks, cudaStream_t s);

// kb_raycast.cu: nearest hit of N rays with the world (links at one configuration + static bodies)
cudaError_t kb_launch_raycast(const KbRayParams& p, cudaStream_t s, int variant = 0);
