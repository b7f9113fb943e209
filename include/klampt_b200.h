/*
 * klampt_b200.h -- C ABI of the B200 batched configuration-feasibility engine.
 *
 * This is the drop-in boundary for ONE hot path of Klamp't: forward kinematics -> self collision ->
 * robot-vs-environment collision / distance -> discretised edge checks, for N configurations (or N
 * edges) at once.  Plain C: opaque handle, int status codes (0 = ok, <0 = error; text from
 * kb_last_error()), caller-owned buffers, no exceptions across the boundary, no torch types.
 * One CUDA device and one CUDA stream per engine handle; calls on one handle are not re-entrant
 * (the reference path is not thread-safe either: Cpp/Planning/RobotCSpace.cpp:632-637).
 *
 * Every entry point names the reference interface it replaces (paths relative to the Klamp't tree).
 * The arithmetic below Klamp't's call sites lives in KrisLibrary (absent from the reference tree);
 * those rows cite the Klamp't call site.
 *
 * Conventions
 *   - rigid transforms: 12 doubles, row-major 3x3 rotation then translation
 *     (file-format convention, Cpp/docs/Manual-FileTypes.md:35-37);
 *   - configurations: row-major N x L doubles, L = number of robot links (Config = Vector of L Reals);
 *   - world IDs: terrains [0,T), rigid objects [T,T+O), robot id T+O, links T+O+1+j
 *     (Cpp/Modeling/World.cpp:47-53,110-180);
 *   - "host" entry points take host pointers and include the H2D / D2H copies;
 *     "_device" entry points take device pointers resident on the engine's device and enqueue on the
 *     engine's stream (kb_set_stream) without synchronising the host.
 */
#ifndef KLAMPT_B200_H
#define KLAMPT_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct kb_engine kb_engine;

/* link types: .rob "jointtype r|p" (Cpp/Modeling/Robot.cpp:326-336) */
enum { KB_REVOLUTE = 0, KB_PRISMATIC = 1 };
/* RobotModelJoint::Type (Cpp/Modeling/Robot.h:28) */
enum { KB_JOINT_WELD = 0, KB_JOINT_NORMAL = 1, KB_JOINT_SPIN = 2, KB_JOINT_FLOATING = 3,
       KB_JOINT_FLOATINGPLANAR = 4, KB_JOINT_BALLANDSOCKET = 5, KB_JOINT_CLOSED = 6 };
/* geometric primitives supported so far (GeometricPrimitive3D: point, sphere, segment, triangle, AABB, oriented box;
 * Cpp/docs/Manual-Geometry.md:22,241-250) */
/* segment: params = 6 doubles a,b (a != b; stored as the zero-area triangle a,b,b, which every predicate treats as the
 * segment).  triangle: params = 9 doubles a,b,c.  BOX (solid oriented box): centre(3), axes as the columns of a row-major 3x3 (9), half
 * dimensions(3).  AABB (solid): lo(3), hi(3).  A box is solid: an element of the other geometry that lies inside it collides. */
enum { KB_PRIM_POINT = 0, KB_PRIM_SPHERE = 1, KB_PRIM_TRIANGLE = 2, KB_PRIM_BOX = 3, KB_PRIM_AABB = 4, KB_PRIM_SEGMENT = 5 };

enum { KB_OK = 0, KB_ERR_INVALID = -1, KB_ERR_STATE = -2, KB_ERR_CUDA = -3, KB_ERR_UNSUPPORTED = -4,
       KB_ERR_NOMEM = -5 };

/* counters since kb_finalize / kb_reset_stats; the getStats-style dictionary of the Python adapter is built
 * from these (Python/klampt/src/motionplanning.cpp:1031-1055) */
typedef struct {
  int64_t configs_checked;      /* configurations through the feasibility kernels                     */
  int64_t configs_feasible;
  int64_t edges_checked;
  int64_t edges_visible;
  int64_t edge_config_checks;   /* configurations checked on behalf of edges                           */
  int64_t node_tests;           /* BV-pair tests        } counted only while option "collect_stats" = 1 */
  int64_t elem_tests;           /* element-pair tests   }                                              */
  int64_t recheck_pairs;        /* element pairs re-run in fp64 because the fp32 filter was uncertain  */
  int64_t kernel_launches;      /* launches of this library's kernels                                  */
  int64_t traverse_launches;    /* launches of the traversal kernel timed while option "time_kernels" = 1 */
  double  traverse_ms;          /* their summed device time (CUDA events on the engine's stream)        */
  double  gpu_ms;               /* device time of the host-buffer entry points, copies included         */
  int64_t items_dropped;        /* (link, static group) pairs ruled out by the clearance grids ("collect_stats") */
  int64_t node_iterations;      /* warp-wide iterations of the node loop ("collect_stats")              */
  int64_t rays_cast;            /* rays through kb_raycast_batch / kb_geom_raycast_batch                */
} kb_stats;

/* ---- lifetime ---------------------------------------------------------------------------------------- */
int  kb_engine_create(kb_engine** out);
void kb_engine_destroy(kb_engine* e);
/* thread-local text of the last error raised by any call on this thread */
const char* kb_last_error(void);
/* library version string and the CUDA arch it was built for ("sm_100a") */
const char* kb_version(void);

/* ---- geometry (AnyCollisionGeometry3D: local data + margin; Python/klampt/src/geometry.h:17-19,140-148) */
/* replaces Geometry3D.setTriangleMesh + setCollisionMargin; returns the geometry index (>=0) */
int kb_add_trimesh(kb_engine* e, const double* verts, int nv, const int32_t* tris, int nt, double margin);
/* replaces Geometry3D.setPointCloud; radius may be NULL (all zero) */
int kb_add_pointcloud(kb_engine* e, const double* pts, int n, const double* radius, double margin);
/* replaces Geometry3D.setGeometricPrimitive; params: point x,y,z / sphere cx,cy,cz,r / segment, triangle, box, aabb: see KB_PRIM_* */
int kb_add_primitive(kb_engine* e, int type, const double* params, double margin);

/* A point cloud whose points are replaced between batches -- sensor streams; the reference keeps such geometries as dynamic
 * geometries (Cpp/Modeling/ManagedGeometry.h:49-52) and rebuilds their collision data on the CPU after Geometry3D.setPointCloud.
 * Reserves room for `capacity` points of one `radius`; attach it with kb_add_terrain / kb_add_rigid_object (once).  It starts empty
 * and forms its own environment group.  kb_update_pointcloud (after kb_finalize) uploads n <= capacity points given in the
 * geometry's local frame and rebuilds the hierarchy on the GPU (linear BVH: Morton order, radix sort, Karras' parallel hierarchy).
 * The scene-building calls (kb_add_trimesh / _pointcloud / _primitive / _rigid_object, kb_robot_create) reject non-finite
 * coordinates, null arrays and negative radii with KB_ERR_INVALID; kb_update_pointcloud is the per-batch path and does NOT scan its
 * input: the caller drops invalid sensor returns (NaN / inf) before the call, as klampt_b200.io.parse_pcd does. */
int kb_add_dynamic_pointcloud(kb_engine* e, int capacity, double radius, double margin);
int kb_update_pointcloud(kb_engine* e, int geom, const double* pts, int n);

/* ---- world entities (WorldModel terrains / rigidObjects / robots, Cpp/Modeling/World.cpp:47-196) ------ */
int kb_add_terrain(kb_engine* e, int geom);                         /* geom = -1: empty geometry */
int kb_add_rigid_object(kb_engine* e, int geom, const double T[12]);
/* RobotKinematics3D data model: parents[i] < i, T0_Parent with the base transform already multiplied into
 * root links (Cpp/Modeling/Robot.cpp:971-975), unit axes, inclusive joint limits */
int kb_robot_create(kb_engine* e, int L, const int32_t* parents, const uint8_t* linktype,
                    const double* axis, const double* T0_parent, const double* qmin, const double* qmax);
int kb_robot_set_link_geometry(kb_engine* e, int link, int geom);
/* RobotModel::joints (type, linkIndex, baseIndex; Cpp/Modeling/Robot.h RobotModelJoint); default = one Normal joint per link.
 * jbase (may be NULL when every joint drives one link) = the link a Floating / FloatingPlanar / BallAndSocket joint hangs
 * from, -1 = world: the joint drives the chain base -> jlink (RobotModel::GetJointIndices, Cpp/Modeling/Robot.cpp:2120-2144),
 * which must be 3 prismatic + 3 revolute (z, y, x) links for Floating, 3 revolute (z, y, x) for BallAndSocket, and end in a
 * revolute link for FloatingPlanar -- the layout Klampt::Interpolate / Distance assert (Cpp/Modeling/Interpolate.cpp:24-26,231-236) */
int kb_robot_set_joints(kb_engine* e, int nj, const uint8_t* jtype, const int32_t* jlink, const int32_t* jbase);
/* RobotModelDriver limits as read by CheckJointLimits (Cpp/Modeling/Robot.cpp:2166-2187):
 * value = mean_k (q[links[k]] - offset[k]) / scale[k];  scale/offset may be NULL (1 / 0) */
int kb_robot_add_driver(kb_engine* e, int n, const int32_t* links, const double* scale, const double* offset,
                        double dmin, double dmax);
/* RobotWithGeometry::InitSelfCollisionPair / delete (Cpp/Modeling/Robot.cpp:1274-1313; robotsim.cpp:5504-5527).
 * Default when never called = InitAllSelfCollisions: all i<j, both non-empty, neither the other's parent. */
int kb_robot_set_self_collision(kb_engine* e, int i, int j, int enabled);
/* WorldPlannerSettings::collisionEnabled (n_ids x n_ids, row-major).  Default when never called =
 * WorldPlannerSettings::InitializeDefault (Cpp/Planning/PlannerSettings.cpp:16-41). */
int kb_set_pair_mask(kb_engine* e, const uint8_t* mask, int n_ids);
int kb_num_ids(const kb_engine* e);
/* writes the n_ids x n_ids mask in effect after kb_finalize */
int kb_get_pair_mask(const kb_engine* e, uint8_t* mask_out);

/* WorldModel::InitCollisions + SingleRobotCSpace::Init (Cpp/Modeling/World.cpp:266-274,
 * Cpp/Planning/RobotCSpace.cpp:668-754): builds the BVHs, flattens them and uploads everything to `device`. */
int kb_finalize(kb_engine* e, int device);
/* The same on several devices of one node (SURVEY 8b "kb_finalize(device_list)"): the hierarchies are built once, on devices[0], and
 * replicated on the others by peer copies.  Afterwards every HOST-buffer entry point (kb_feasible_batch*, kb_edges_visible_batch*,
 * kb_distance_batch*) cuts its batch into contiguous shards, one per device, runs them concurrently (one host thread, stream and
 * scratch per device) and writes the results straight into the caller's buffers; batches below option "multi_min" (default 8192)
 * stay on devices[0].  The *_device entry points, kb_fk_batch, kb_colliding_pairs_batch and the kb_geom_* queries use devices[0].
 * A C++ planner thus drives 8 GPUs through one BatchSingleRobotCSpace without a process per GPU. */
int kb_finalize_multi(kb_engine* e, const int* devices, int n_devices);
int kb_num_devices(const kb_engine* e);
/* use an existing CUDA stream (cudaStream_t) for all later work; NULL = the engine's own non-blocking stream.
 * The legacy default stream must be named explicitly as cudaStreamLegacy ((void*)0x1). */
int kb_set_stream(kb_engine* e, void* cuda_stream);
int kb_synchronize(kb_engine* e);
/* tuning / instrumentation knobs: "collect_stats" (0|1: count node / element tests and fp64 rechecks in the kernels),
 * "time_kernels" (0|1: bracket every traversal launch with CUDA events on the engine's stream, summed into kb_stats),
 * "chunk" (configurations per kernel launch, default up to 1 M within a 2 GB scratch budget),
 * "pipeline" (0 = fused traversal kernel, default; 1 = split pipeline: lean node kernel -> global leaf-pair list -> leaf kernel ->
 * fused kernel on requeued configurations; same results, measured slower on C2/C3), "leaf_budget" (split pipeline only),
 * "clear_grid" (0|1, default 0: clearance-grid broad phase that drops (link, static group) pairs before the BVH descent; exact),
 * "both_limit" (experiment knob of builds with KB_BOTH_MODE=2), "wide" (0|1, default 1: the boolean query runs on the 4-wide
 * hierarchies when every work item has them), "graph_max" (default 16384: host-buffer batches up to this size take the small-batch
 * path -- pinned staging, one CUDA graph per batch size, one warp per configuration; can only grow before the first small batch),
 * "zero_copy_max" (default 64: small batches up to this size are read from / written to pinned host memory by the kernels
 * themselves), "edge_flat_max" (default 262144: edge batches with at most this many midpoints are checked all levels at once),
 * "multi_min" (default 8192: smaller host batches of a multi-device handle stay on its first device), "ray_tile" (0|1, default 1:
 * kb_camera_depth renders 8 x 4 pixel tiles per warp), "ray_variant" (launch shape of the ray kernel, experiment knob).
 * Before kb_finalize only: "cloud_leaf" (points per leaf of a host-built point-cloud hierarchy, 1..32, default 8), "grid_res" (voxels along the longest axis of a clearance grid, 0 = none, default 256),
 * "cloud_builder" / "mesh_builder" (0 = binned SAH on the host, default; 1 = linear BVH built on the GPU for point clouds above
 * 4096 points / meshes above 16384 triangles: kb_finalize far faster, queries 7-13 % slower, identical answers). */
int kb_set_option(kb_engine* e, const char* name, int64_t value);

/* ---- the hot path ------------------------------------------------------------------------------------- */
/* RobotKinematics3D::UpdateConfig/UpdateFrames for N configurations (call sites Cpp/Planning/RobotCSpace.cpp:612,634).
 * T_out: N x L x 12 doubles. */
int kb_fk_batch(kb_engine* e, const double* Q, int64_t N, double* T_out);

/* SingleRobotCSpace::IsFeasible for N configurations (Cpp/Planning/RobotCSpace.cpp:786-823):
 * out[c] = 1 iff joint/driver limits hold and no enabled pair collides.
 * first_pair (optional, N x 2 int32): world ids of one colliding pair, or -1,-1 (which pair is reported first is
 * traversal-order dependent in the reference as well). */
int kb_feasible_batch(kb_engine* e, const double* Q, int64_t N, uint8_t* out, int32_t* first_pair);
int kb_feasible_batch_device(kb_engine* e, const double* dQ, int64_t N, uint8_t* d_out, int32_t* d_first_pair);
/* the same for configurations stored as floats (half the host-to-device bytes): every value is widened to fp64 on the device
 * and checked exactly as kb_feasible_batch checks (double)q -- the answer is for the rounded configuration */
int kb_feasible_batch_f32(kb_engine* e, const float* Q, int64_t N, uint8_t* out, int32_t* first_pair);
/* the same with the result as a packed bitmask -- bit (c & 7) of byte (c >> 3) is configuration c, (N + 7) / 8 bytes: what a
 * multi-GPU caller gathers ("only result bitmasks and distances are gathered").  The device form writes whole
 * 32-bit words: d_out_bits must be 4-byte aligned and hold (N + 31) / 32 words. */
int kb_feasible_batch_bits(kb_engine* e, const double* Q, int64_t N, uint8_t* out_bits);
int kb_feasible_batch_bits_device(kb_engine* e, const double* dQ, int64_t N, uint8_t* d_out_bits);

/* SingleRobotCSpace::PathChecker(a,b)->IsVisible() = EpsilonEdgeChecker with RobotCSpace::Distance / Interpolate
 * (Cpp/Planning/RobotCSpace.cpp:835-838, Cpp/Modeling/Interpolate.cpp:10-71,208-343) for N edges.
 * weights: per-joint metric weights or NULL.  out[e] = 1 iff every bisection midpoint is feasible (endpoints are not
 * re-checked).  nchecks (optional): number of feasibility checks the sequential early-exit checker performs. */
int kb_edges_visible_batch(kb_engine* e, const double* A, const double* B, int64_t N, double eps,
                           const double* weights, uint8_t* out, int32_t* nchecks);
int kb_edges_visible_batch_device(kb_engine* e, const double* dA, const double* dB, int64_t N, double eps,
                                  const double* weights_host, uint8_t* d_out, int32_t* d_nchecks);
/* visibility as a packed bitmask, (N + 7) / 8 bytes, bit order as kb_feasible_batch_bits */
int kb_edges_visible_batch_bits(kb_engine* e, const double* A, const double* B, int64_t N, double eps,
                                const double* weights, uint8_t* out_bits, int32_t* nchecks);

/* Every colliding pair of each configuration: SingleRobotCSpace::Init's per-pair CollisionFreeSet constraints
 * (Cpp/Planning/RobotCSpace.cpp:697-747) evaluated together, as CSpaceInterface::feasibilityFailures needs them
 * (Python/klampt/src/motionplanning.h:122-171).  out_pairs: N x max_pairs x 2 world ids, -1 padded (1 <= max_pairs <= 32);
 * out_count: pairs found (may exceed max_pairs; the surplus is not stored), -1 where the joint / driver limits fail. */
int kb_colliding_pairs_batch(kb_engine* e, const double* Q, int64_t N, int max_pairs, int32_t* out_pairs, int32_t* out_count);

/* WorldPlannerSettings::DistanceLowerBound(world, {robot}, {environment}, eps=0, bound) for N configurations
 * (Cpp/Planning/PlannerSettings.cpp:570-620): min over enabled pairs of (geometric distance - margins), capped at
 * upper_bound; include_self adds the enabled self pairs.  out_pair optional (N x 2 world ids, -1 if capped). */
int kb_distance_batch(kb_engine* e, const double* Q, int64_t N, double upper_bound, int include_self,
                      double* out_d, int32_t* out_pair);
int kb_distance_batch_device(kb_engine* e, const double* dQ, int64_t N, double upper_bound, int include_self,
                             double* d_out_d, int32_t* d_out_pair);
/* The full AnyCollisionQuery::Distance(absErr, relErr, bound) (Cpp/Planning/PlannerSettings.cpp:109-115 passes eps as both, the
 * Python binding DistanceQuerySettings, Python/klampt/src/geometry.h:605-629) with the rest of DistanceQueryResult
 * (src/geometry.h:631-694).  abs_err / rel_err >= 0: a branch is abandoned once it cannot improve the running minimum by more
 * than abs_err or rel_err * |minimum|, so the value is within that tolerance above the exact minimum (0, 0 = exact).
 * out_cp (optional, N x 6): closest points in the world frame, first the point on the geometry of out_pair[0] then the one on
 * out_pair[1], on the margin- and radius-inflated surfaces (|cp2 - cp1| = d when d > 0); NaN where the bound was returned.
 * out_elem (optional, N x 2): index of the closest triangle / point in each geometry's own element order (0 for primitives), -1
 * where the bound was returned. */
int kb_distance_batch_ex(kb_engine* e, const double* Q, int64_t N, double abs_err, double rel_err, double upper_bound, int include_self,
                         double* out_d, int32_t* out_pair, double* out_cp, int32_t* out_elem);

/* AnyCollisionQuery between two registered geometries at explicit transforms, N transform pairs at once
 * (Geometry3D.collides / withinDistance / distance, Python/klampt/src/robotsim.cpp:1656-1819).
 * Ta, Tb: N x 12.  tol = 0 -> Collide(), tol > 0 -> WithinDistance(tol).  */
int kb_geom_collides_batch(kb_engine* e, int ga, const double* Ta, int gb, const double* Tb, int64_t N, double tol,
                           uint8_t* out);
int kb_geom_distance_batch(kb_engine* e, int ga, const double* Ta, int gb, const double* Tb, int64_t N,
                           double upper_bound, double* out_d);
/* Geometry3D.distance_ext(other, settings) -> DistanceQueryResult (Python/klampt/src/robotsim.cpp:1765-1819): distance with
 * tolerances, closest points (cp1 on ga, cp2 on gb) and element indices, as kb_distance_batch_ex */
int kb_geom_distance_batch_ex(kb_engine* e, int ga, const double* Ta, int gb, const double* Tb, int64_t N, double abs_err, double rel_err,
                              double upper_bound, double* out_d, double* out_cp, int32_t* out_elem);

/* ---- ray casting --------------------------------------------------------------------- */
/* WorldModel::RayCast / RayCastIgnore (Cpp/Modeling/World.cpp:465-588) for N rays at once -- what the camera sensor's fallback
 * calls per pixel (Cpp/Sensing/VisualSensors.cpp:430-475), the laser sensor per measurement (:113-134) and WorldCollider.rayCast
 * per query (Python/klampt/model/collide.py:700-748).  q: the robot's configuration (L doubles, host), NULL = the robot is left
 * out.  rays: N x 6 = source xyz, direction xyz (any length > 0; normalised internally, as the sensors do).  ignore_ids: one byte
 * per world id (kb_num_ids), 1 = rays pass through that body; NULL = none.  out_id: world id of the nearest body hit or -1;
 * out_dist: distance from the source along the direction (+inf when nothing is hit; the hit point is source + dist * unit
 * direction); out_elem (optional): triangle / point index within the body's geometry.  Every link, rigid object and terrain with a
 * geometry is seen, whatever the collision mask says.  A triangle mesh reports its nearest two-sided intersection minus its
 * collision margin, a point cloud where the ray enters the first sphere of radius (point radius + margin); ties in distance keep
 * the body the reference visits first (links in order, rigid objects, terrains). */
int kb_raycast_batch(kb_engine* e, const double* q, const double* rays, int64_t N, const uint8_t* ignore_ids,
                     int32_t* out_id, double* out_dist, int32_t* out_elem);
/* rays given as floats (N x 6): half the upload, widened to fp64 on the device -- the answer is the fp64 answer for the rounded rays */
int kb_raycast_batch_f32(kb_engine* e, const double* q, const float* rays, int64_t N, const uint8_t* ignore_ids,
                         int32_t* out_id, double* out_dist, int32_t* out_elem);
/* the same with the rays and the results device-resident on the engine's stream (q and ignore_ids stay host pointers) */
int kb_raycast_batch_device(kb_engine* e, const double* q, const double* d_rays, int64_t N, const uint8_t* ignore_ids,
                            int32_t* d_out_id, double* d_out_dist, int32_t* d_out_elem);
/* CameraSensor's ray-cast rendering (Cpp/Sensing/VisualSensors.cpp:413-475) in one call: the rays of an xres x yres pinhole image are
 * built on the device exactly as the reference builds them per pixel -- pixel (i, j) looks along fwd + (i - cx) right / fx +
 * (cy - j) up / fy, starts zmin along that vector -- and cast as kb_raycast_batch does.  pose: the camera's world pose, 12 doubles
 * (row-major R then t), camera frame x right, y down, z forward (Klamp't's sensor convention; for a camera on a link: the link's
 * transform from kb_fk_batch times the sensor's Tsensor).  out_depth (optional, yres x xres floats, row j = image row j): depth along
 * the viewing direction, fwd . (pt - eye); readings below zmin and pixels that see nothing report zmax.  out_id (optional): world
 * id seen by each pixel, -1 = background (the reference colours the pixel by that body's appearance). */
typedef struct {
  double pose[12];
  double fx, fy, cx, cy;          /* CameraSensor::GetViewport (:865-889): fx = xres / 2 / tan(xfov / 2), cx = xres / 2, ... */
  double zmin, zmax;
  int32_t xres, yres;
} kb_camera;
int kb_camera_depth(kb_engine* e, const double* q, const kb_camera* cam, const uint8_t* ignore_ids, float* out_depth, int32_t* out_id);
/* Geometry3D::rayCast_ext (Python/klampt/src/geometry.cpp:1837-1852) of one registered geometry at transform T (12 doubles, NULL =
 * identity): out_elem = element hit or -1, out_dist as above */
int kb_geom_raycast_batch(kb_engine* e, int geom, const double* T, const double* rays, int64_t N, int32_t* out_elem, double* out_dist);

/* ---- introspection ------------------------------------------------------------------------------------ */
int kb_get_stats(kb_engine* e, kb_stats* out);
int kb_reset_stats(kb_engine* e);
/* sizes of the device-resident static data: [0] bvh nodes, [1] elements, [2] bytes, [3] work items per config,
 * [4] merged environment groups, [5] max BVH depth sum */
int kb_get_layout(const kb_engine* e, int64_t out[6]);

#ifdef __cplusplus
}
#endif
#endif
