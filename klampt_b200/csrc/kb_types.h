// kb_types.h -- data layout shared by the host engine and the sm_100a kernels.
//
// HBM layout (all static data is uploaded once at kb_finalize and then read-only):
//   nodes    float4[2*n]   flattened BVHs of every geometry, one after the other.  Node i of a geometry =
//                          { centre.xyz, as_float(left) } { half_extent.xyz, as_float(count) }.  left >= 0: inner node,
//                          children at left and left+1 (siblings adjacent and even-aligned: one 64 B line; indices relative to the geometry's
//                          node base).  left < 0: leaf, first element = ~left (relative to the geometry's element
//                          base), count elements.  Boxes are fp32; half extents are rounded up so the fp32 box contains the fp64 one.
//   tris32   float4[3*n]   triangle vertices (fp32, local frame; merged environment groups: world frame);
//                          .w of vertex 0 = as_float(owner world id)
//   tris64   double[9*n]   the same triangles in fp64 for the exact recheck
//   sph32    float4[n]     point-cloud points / sphere primitives: xyz, radius ; sph64 double[4*n]
//   sphown   int[n]        owner world id per sphere element
//   box32/64 float4[4n] / double[16n]  solid boxes of box primitives (centre, axes, half dimensions); boxown int[n]
//   items    KbItem[]      the per-configuration work list: one entry per enabled geometry pair
//                          (link vs merged environment group, link vs link)
//   robot    KbRobotDev    SoA joint arrays for FK
// Per batch: Q (N x L f64, caller), xf64 (chunk x nxf x 12 f64, written by FK, read by the traversal),
// state/hit arrays (1-4 B per configuration).
#pragma once
#include <stdint.h>

#define KB_MAX_LINKS 128          // links per robot supported by the FK kernel's shared-memory model
#ifndef KB_STACK_CAP
#define KB_STACK_CAP 448          // node-pair stack entries per warp (448: five CTAs of the boolean kernel fit the 164 KB shared-memory carve-out, leaving 92 KB of L1)
#endif
#define KB_LEAFQ_CAP 96           // leaf-pair queue entries per warp
#define KB_ITEM_BITS 12
#define KB_NODEA_BITS 20
#define KB_MAX_ITEMS (1 << KB_ITEM_BITS)
#define KB_MAX_NODES_A (1 << KB_NODEA_BITS)
#define KB_WARPS_PER_BLOCK 4
#ifndef KB_POP_WIDTH
#define KB_POP_WIDTH 32           // frontier entries popped per iteration of the boolean kernel (<= 32)
#endif
#ifndef KB_LEAF_TRIGGER
#define KB_LEAF_TRIGGER 32        // leaf pairs queued before the element phase runs (it also runs whenever the node stack is empty)
#endif
#ifndef KB_BLOCKS_PER_SM
#define KB_BLOCKS_PER_SM 4         // resident CTAs per SM the traversal kernel is compiled for (register cap = 65536 / (128 * this))
#endif

enum { KB_ELEM_TRI = 0, KB_ELEM_SPHERE = 1, KB_ELEM_BOX = 2 };   // BOX: the solid of a box primitive (its surface is 12 TRI elements)

// Clearance grids (broad phase of the boolean query).  For every merged static environment group a uniform voxel grid
// over the group's bounds (+ a pad) stores, per voxel, a conservative lower bound on the distance from ANY point of the
// voxel to the group's surface, in quarter-voxel units (u8).  A link whose covering spheres all have more clearance than
// radius + threshold cannot touch the group, so its (link, group) work item is dropped before the BVH descent -- the
// reference's per-object AABB pre-reject (Cpp/Planning/PlannerSettings.cpp:290-317) taken to an O(1) lookup.
#define KB_MAX_GRIDS 4
#define KB_COVER_MAX 8            // covering spheres per link geometry
#define KB_PROBES_SMEM_MAX 512    // probe lists up to this size are cached in shared memory per CTA (32 B each)
struct KbClearGrid {
  const uint8_t* data;            // dims[0]*dims[1]*dims[2] bytes, x fastest
  float o[3];                     // world position of voxel (0,0,0)'s lower corner
  float inv_h;                    // 1 / voxel edge
  int32_t dims[3];
  int32_t pad_;
};
struct KbProbe {                  // one covering sphere of the moving side of a (link, static group) item: 32 bytes
  float c[3];                     // centre in the link frame
  uint32_t need;                  // quarter-voxels of clearance that rule a contact out: radius + threshold + slack, rounded up
  int32_t item;                   // work item this probe belongs to
  int32_t xf;                     // transform slot of the link
  int32_t grid;                   // clearance grid of the static side
  int32_t pad_;
};

struct KbItem {                   // 80 bytes
  int32_t nodeA, nodeB;           // global node index of the two roots
  int32_t elemA, elemB;           // global element base (into tris* or sph* according to kind)
  int16_t xfA, xfB;               // transform slot in the per-configuration table, -1 = identity (static world frame)
  uint8_t kindA, kindB;           // KB_ELEM_*
  uint16_t flags;                 // bit 0: self pair (link vs link)
  int32_t idA, idB;               // world ids; -1 = look the owner up per element (merged environment group)
  double thr;                     // collision threshold: margin_A + margin_B (+ tolerance); 0 = surfaces must intersect
  double marg;                    // margin_A + margin_B, subtracted from reported distances
  double rsum;                    // largest sphere radius of A + of B: bound on how far two touching boxes' elements can interpenetrate
  double margA, margB;            // the two margins separately: closest points are reported on the margin-inflated surfaces
  int32_t wideA, wideB;           // slot base of each side's 4-wide hierarchy in KbScene::wide, -1 = that geometry has none (GPU-built trees)
};

struct KbRobotDev {
  int32_t L;
  int32_t nj, ndrv_terms, ndrv;
  int32_t parents[KB_MAX_LINKS];
  uint8_t linktype[KB_MAX_LINKS];
  double axis[KB_MAX_LINKS * 3];
  double T0[KB_MAX_LINKS * 12];
  double qmin[KB_MAX_LINKS], qmax[KB_MAX_LINKS];
  uint8_t jtype[KB_MAX_LINKS];
  int32_t jlink[KB_MAX_LINKS];
  int16_t jidx[KB_MAX_LINKS][6];  // links driven by a multi-link joint, root to tip (RobotModel::GetJointIndices): Floating 6, FloatingPlanar / BallAndSocket 3
};

struct KbDriverDev {              // flattened affine drivers: driver d covers terms [first, first+n)
  int32_t first, n;
  double dmin, dmax;
};

struct KbScene {                  // device pointers to the static data
  const float4* nodes;
  const float4* wide;             // 4-wide hierarchies (kb_traverse_wide_kernel): slot = {centre.xyz, ref} {half.xyz, count}, 4 slots = one 128 B node
  const float4* tris32;
  const double* tris64;
  const float4* sph32;
  const double* sph64;
  const int32_t* triown;
  const int32_t* sphown;
  const float4* box32;            // solid boxes: 4 x float4 per box = {centre.xyz, hx} {axis0.xyz, hy} {axis1.xyz, hz} {axis2.xyz, -}
  const double* box64;            // 16 doubles per box in the same order
  const int32_t* boxown;
  const int32_t* triorig;         // per triangle / sphere element: its index in the geometry it came from (elements are stored in BVH leaf order)
  const int32_t* sphorig;
  float eps_abs;                  // absolute fp32 coordinate error bound for this scene (metres)
  float qo[3], qs[3];             // KB_QNODES builds only: origin and step of the 16-bit node quantisation grid
  KbClearGrid grids[KB_MAX_GRIDS];
};

struct KbTraverseParams {
  KbScene scene;
  const KbItem* items;
  int32_t nitems;
  int32_t nxf;                    // transform slots per configuration
  const double* xf64;             // N x nxf x 12
  int64_t N;
  const uint8_t* state;           // per configuration: 1 = run, 0 = skip (limits failed / edge already dead); may be null
  int32_t* hit;                   // per configuration: -1 none, else item index of a colliding pair
  int32_t* hit_elem;              // per configuration: elemA / elemB (2 ints) of that pair; may be null
  uint32_t* work_counter;         // dynamic work distribution
  unsigned long long* counters;   // [0] rechecks, [1] node tests, [2] leaf tests, [7] items dropped by the clearance grids (optional statistics)
  int32_t wide_limit;             // stack size up to which 32-wide pops are allowed
  int32_t collect_stats;
  int32_t has_boxes;              // the work list holds solid-box items (selects the kernel instantiation with the box predicates)
  int32_t use_wide;               // every item has both 4-wide hierarchies and the option is on: kb_traverse_wide_kernel
  int32_t wide_room;              // wide kernel: entries may be popped 8 at a time while sp <= wide_room, one at a time above
  int32_t pop_room;               // boolean kernel: m entries may be popped while sp + 3 m <= pop_room (see make_params)
  int32_t both_limit;             // frontier size up to which comparable inner pairs push all four child pairs (0 = never)
  float both_ratio;               // two inner nodes descend both trees at once while their squared diagonals are within this ratio (4: link vs
                                  // environment, where the boxes differ in size; 16: link vs link only -- measured, profiles/r02_experiments.md)
  const KbProbe* probes;          // clearance probes of this item set (boolean kernel only); null / 0 = no pre-filter
  int32_t nprobes;
  const uint32_t* always_on;      // bit i set: item i has no probes and is always traversed ((nitems + 31) / 32 words)
  // small batches (kb_feasible_batch, N <= graph_max): warp w checks configuration w and nothing else -- no work counter to reset, no
  // finish kernel: the warp writes the result byte itself once its parked fp64 rechecks are resolved
  int32_t static_sched;
  uint8_t* out_bytes;             // static_sched: 1 = feasible, 0 = not (limits or collision)
  unsigned long long* nfeasible;  // static_sched: feasible configurations are counted here (statistics)
};

// ray casting (kb_raycast.cu): one entry per body a ray can hit -- a robot link, a rigid object, a terrain
struct KbRayBody {                // 144 bytes
  double T[12];                   // world <- local of a static body (has_T); links take theirs from the FK table (xf >= 0)
  double margin;
  int32_t node_base, elem_base;   // the geometry's local-frame hierarchy / elements
  int32_t kind;                   // KB_ELEM_TRI / KB_ELEM_SPHERE
  int32_t id;                     // world id reported for a hit; -1 = a merged environment group: the owner id comes with the element
  int32_t rank;                   // order in which WorldModel::RayCast visits the bodies: a tie in distance keeps the lower rank (groups: from the owner)
  int32_t xf;                     // transform slot of a link, -1 = static
  int32_t has_T;                  // static body with a transform other than the identity
  float ext;                      // largest |coordinate| of the geometry's local box (bounds the fp32 rounding of its node tests); < 0: from eps_abs
};
struct KbRayParams {
  KbScene scene;
  const KbRayBody* bodies;        // [0, nlinkbodies): tested one by one; then nstatic bodies in the leaf order of the top-level hierarchy
  int32_t nlinkbodies, nstatic;
  int32_t nterr, nobj, nlinks;    // world counts: rank of an owner id inside a merged group (links, then objects, then terrains)
  const float4* tlas;             // top-level hierarchy over the static bodies' world boxes (same node format; leaf = body range)
  float tlas_ext;                 // largest |coordinate| of its root box
  double max_margin;              // largest mesh margin among the static bodies (a body reports t - margin)
  const double* xf64;             // link transforms of this call's configuration (nxf x 12), null when there is no link body
  const double* rays;             // N x 6: source, direction (any length > 0)
  int64_t N;
  const uint8_t* ignore;          // per world id: 1 = rays pass through (RayCastIgnore); may be null
  int32_t* out_id;                // world id of the nearest hit, -1 = none
  double* out_dist;               // distance along the normalised direction, +inf = none
  int32_t* out_elem;              // element index within the body's geometry; may be null
  // camera mode (kb_camera_depth): rays == null, ray k = pixel (k % xres, k / xres) of a pinhole camera, built the way the camera
  // sensor's ray-cast path builds them (VisualSensors.cpp:424-449)
  int32_t cam_on, xres, yres, tile;     // tile: a warp renders 8 x 4 pixels instead of 32 pixels of one row
  double eye[3], fwd[3], dx[3], dy[3];   // dx = right / fx, dy = up / fy
  double cx, cy, zmin, zmax;
  float* out_depth;               // forward depth per pixel, zmax where nothing is seen; may be null
};

// split pipeline (node traversal kernel -> global leaf-pair list -> leaf kernel -> requeue for the fused kernel)
struct KbSplitParams {
  uint4* leaf_list;               // (configuration, item | nodeA, nodeB, 0)
  unsigned long long* leaf_count; // entries appended (may exceed leaf_cap: the surplus was not written)
  unsigned int leaf_cap;
  uint8_t* flagged;               // per configuration: stopped on the leaf budget with work left, or list full
  uint8_t* state2;                // per configuration: re-run by the fused kernel
  unsigned long long* requeued;   // statistics: configurations re-run
  int32_t leaf_budget;            // leaf pairs after which a configuration stops traversing
  int32_t stack_cap;              // node-pair stack entries per warp of the node kernel
  int32_t wide_limit;
};
