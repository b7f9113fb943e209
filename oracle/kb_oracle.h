/*
 * kb_oracle.h -- CPU oracle for the batched configuration-feasibility path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in klampt_b200/ (the product) may include,
 * link or call this.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs use it, and only as the checker / CPU
 * baseline.
 *
 * PARITY UNPINNED: the arithmetic of this path lives in KrisLibrary
 * (github.com/krishauser/KrisLibrary @ master, unpinned: reference
 * Cpp/Dependencies/Makefile:16), which is absent from /root/reference, and the
 * reference's own tests hold no golden vectors for FK / collision / distance /
 * edge visibility.  This file is a plain fp64 restatement
 * of the semantics visible at Klamp't's call sites; every function cites the
 * reference file:line it follows.  It is pinned only by geometric definition
 * and by the analytic known-answer tests in tests/test_oracle_*.py.
 *
 * Conventions: rigid transforms are 12 doubles = row-major 3x3 R then t
 * (file-format convention, reference Cpp/docs/Manual-FileTypes.md:35-37).
 */
#ifndef KB_ORACLE_H
#define KB_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct ko_world ko_world;

/* link types: reference Cpp/Modeling/Robot.cpp:326-336 (jointtype r / p) */
enum { KO_REVOLUTE = 0, KO_PRISMATIC = 1 };
/* joint types: RobotModelJoint, reference Cpp/Modeling/Robot.h:26-39 */
enum { KO_JOINT_WELD = 0, KO_JOINT_NORMAL = 1, KO_JOINT_SPIN = 2, KO_JOINT_FLOATING = 3,
       KO_JOINT_FLOATINGPLANAR = 4, KO_JOINT_BALLANDSOCKET = 5, KO_JOINT_CLOSED = 6 };
/* primitive types (subset of GeometricPrimitive3D, Manual-Geometry.md:22) */
enum { KO_PRIM_POINT = 0, KO_PRIM_SPHERE = 1, KO_PRIM_TRIANGLE = 2, KO_PRIM_BOX = 3, KO_PRIM_AABB = 4, KO_PRIM_SEGMENT = 5 };

/* per-config traversal counters of the canonical traversal */
typedef struct {
  int64_t n_box;   /* AABB-pair tests in the broad phase (a8 / a9)          */
  int64_t n_node;  /* BV-pair overlap tests in BVH descents                 */
  int64_t n_tri;   /* triangle-pair tests                                   */
  int64_t n_pt;    /* point(-sphere) element tests                          */
} ko_counts;

ko_world* ko_create(void);
void ko_destroy(ko_world* w);

/* geometries (local frame data); return geometry index >= 0 */
int ko_add_trimesh(ko_world* w, const double* verts, int nv, const int32_t* tris, int nt, double margin);
int ko_add_pointcloud(ko_world* w, const double* pts, int n, const double* radius, double margin);
int ko_add_primitive(ko_world* w, int type, const double* params, double margin);

/* world entities; ID scheme of reference Cpp/Modeling/World.cpp:47-53,110-180:
 * terrains [0,T), rigid objects [T,T+O), robot id, then L link ids */
int ko_add_terrain(ko_world* w, int geom);
int ko_add_rigid_object(ko_world* w, int geom, const double T[12]);
int ko_robot_create(ko_world* w, int L, const int32_t* parents, const uint8_t* linktype,
                    const double* axis, const double* T0, const double* qmin, const double* qmax);
int ko_robot_set_link_geometry(ko_world* w, int link, int geom);
/* jbase: per joint the link the joint hangs from (-1 = world; RobotModelJoint::baseIndex); NULL = the parent of jlink */
int ko_robot_set_joints(ko_world* w, int nj, const uint8_t* jtype, const int32_t* jlink, const int32_t* jbase);
int ko_robot_add_affine_driver(ko_world* w, int n, const int32_t* links, const double* scale,
                               const double* offset, double dmin, double dmax);
int ko_robot_set_self_collision(ko_world* w, int i, int j, int enabled);
int ko_set_pair_mask(ko_world* w, const uint8_t* mask, int n_ids);
int ko_finalize(ko_world* w);

int ko_num_ids(const ko_world* w);
int ko_get_pair_mask(const ko_world* w, uint8_t* mask_out);

/* a3: FK; T_out = L x 12 */
void ko_fk(const ko_world* w, const double* q, double* T_out);
/* a2 */
int ko_check_joint_limits(const ko_world* w, const double* q);
/* a1: returns 1 feasible / 0 infeasible; pair_out[2] (world ids) or -1,-1; counts accumulates */
int ko_feasible(const ko_world* w, const double* q, int32_t* pair_out, ko_counts* counts);
/* brute force version (all triangle pairs of all enabled pairs, no BVH, no AABB reject) */
int ko_feasible_brute(const ko_world* w, const double* q);
void ko_feasible_batch(const ko_world* w, const double* Q, int64_t N, uint8_t* out,
                       int32_t* first_pair, ko_counts* counts_per_config, int nthreads);

/* test support for the two-sided 1e-6 m band: deepest contact of q over the enabled pairs (threshold - signed distance, with
 * intersecting triangles measured by their separating-axis penetration depth); -1 when nothing is in contact */
double ko_penetration(const ko_world* w, const double* q, int include_self);
double ko_tri_tri_depth(const double a[9], const double b[9]);
double ko_geom_penetration(const ko_world* w, int ga, const double Ta[12], int gb, const double Tb[12], double tol);

/* a17/a18: EpsilonEdgeChecker with RobotCSpace::Distance / Interpolate */
double ko_cspace_distance(const ko_world* w, const double* a, const double* b, const double* weights);
void ko_interpolate(const ko_world* w, const double* a, const double* b, double u, double* out);
int ko_edge_visible(const ko_world* w, const double* a, const double* b, double eps,
                    const double* weights, int32_t* nchecks, ko_counts* counts);
void ko_edges_visible_batch(const ko_world* w, const double* A, const double* B, int64_t N, double eps,
                            const double* weights, uint8_t* out, int32_t* nchecks, int nthreads);

/* a19: min distance robot vs environment (+ self if include_self) with upper bound */
double ko_distance(const ko_world* w, const double* q, double upper_bound, int include_self,
                   int32_t* pair_out, ko_counts* counts);
void ko_distance_batch(const ko_world* w, const double* Q, int64_t N, double upper_bound, int include_self,
                       double* out_d, int32_t* out_pair, int nthreads);

/* geometry-pair queries at explicit transforms (a12/a13) */
int ko_geom_collides(const ko_world* w, int ga, const double Ta[12], int gb, const double Tb[12]);
int ko_geom_within_distance(const ko_world* w, int ga, const double Ta[12], int gb, const double Tb[12], double tol);
double ko_geom_distance(const ko_world* w, int ga, const double Ta[12], int gb, const double Tb[12], double upper_bound);
double ko_geom_distance_brute(const ko_world* w, int ga, const double Ta[12], int gb, const double Tb[12]);
void ko_geom_aabb(const ko_world* w, int g, const double T[12], double bmin[3], double bmax[3]);

/* element predicates, exposed for known-answer tests */
int ko_tri_tri_intersect(const double a[9], const double b[9]);
double ko_tri_tri_distance(const double a[9], const double b[9]);
double ko_point_tri_distance(const double p[3], const double t[9]);
double ko_seg_seg_distance(const double p0[3], const double p1[3], const double q0[3], const double q1[3]);

/* f4: WorldModel::RayCast / RayCastIgnore (World.cpp:465-588) with the robot at q (NULL: robot left out); returns the world id hit
 * or -1, *dist = length along the (normalised) direction or +inf, *elem = element index within the body's geometry.
 * ko_geom_raycast: Geometry3D::rayCast_ext (Python/klampt/src/geometry.cpp:1837-1852) of geometry g at transform T; brute = 1
 * tests every element without the hierarchy (cross-check). */
int ko_raycast(const ko_world* w, const double* q, const double s[3], const double d[3], const uint8_t* ignore_ids, double* dist, int32_t* elem);
void ko_raycast_batch(const ko_world* w, const double* q, const double* rays, int64_t N, const uint8_t* ignore_ids,
                      int32_t* ids, double* dist, int32_t* elem, int nthreads);
void ko_raycast_counts(const ko_world* w, const double* q, const double* rays, int64_t N, ko_counts* total);
int ko_geom_raycast(const ko_world* w, int g, const double T[12], const double s[3], const double d[3], double* dist, int32_t* elem, int brute);

int ko_max_threads(void);

#ifdef __cplusplus
}
#endif
#endif
