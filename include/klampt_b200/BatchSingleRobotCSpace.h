// BatchSingleRobotCSpace.h -- header-only C++17 adapter over the C ABI (include/klampt_b200.h).
//
// Mirrors the public face of Klampt::SingleRobotCSpace (reference Cpp/Planning/RobotCSpace.h:103-141, .cpp:592-840) for
// the feasibility / visibility path, so a KrisLibrary planner that holds a CSpace* keeps calling
//   IsFeasible(x), PathChecker(a,b)->IsVisible(), Distance, Interpolate, Sample, SampleNeighborhood, Properties
// and gains the two batch entry points the north star asks for:
//   IsFeasibleBatch(Q, N, out)      IsVisibleBatch(A, B, N, eps, out)
// In a Klamp't build this class derives from CSpace and is filled from a WorldModel (INTEGRATION.md shows the ~40 lines
// of glue); here it is self-contained (Config = std::vector<double>) because KrisLibrary's headers are not available.
//
// Side-effect contract: the reference leaves the RobotModel at configuration x after IsFeasible(x); the batch calls
// do not touch any host-side robot model (stateless), the single-configuration calls record `lastConfig`.
// Errors: the reference aborts (Assert / FatalError); this adapter throws std::runtime_error with kb_last_error().
#pragma once
#include "../klampt_b200.h"

#include <algorithm>
#include <cmath>
#include <map>
#include <memory>
#include <random>
#include <stdexcept>
#include <string>
#include <vector>

namespace klampt_b200 {

typedef std::vector<double> Config;

inline void kbCheck(int rc) { if (rc < 0) throw std::runtime_error(std::string("klampt_b200: ") + kb_last_error()); }

// Plain description of the single-robot world SingleRobotCSpace works on (world + robot index + planner settings).
struct RobotDescription {
  std::vector<int32_t> parents; std::vector<uint8_t> linkType; std::vector<double> axis, T0Parent, qMin, qMax;
  std::vector<int> linkGeometry;                   // geometry index per link, -1 = none
  std::vector<uint8_t> jointType; std::vector<int32_t> jointLink;   // empty = one Normal joint per link
  std::vector<int32_t> jointBase;                                     // RobotModelJoint::baseIndex per joint; only needed for Floating / FloatingPlanar / BallAndSocket joints
};

class WorldBuilder {
 public:
  WorldBuilder() { kbCheck(kb_engine_create(&e_)); }
  ~WorldBuilder() { if (e_) kb_engine_destroy(e_); }
  int AddTriMesh(const std::vector<double>& verts, const std::vector<int32_t>& tris, double margin = 0) {
    int g = kb_add_trimesh(e_, verts.data(), (int)(verts.size() / 3), tris.data(), (int)(tris.size() / 3), margin); kbCheck(g); return g; }
  int AddPointCloud(const std::vector<double>& pts, const std::vector<double>* radius = nullptr, double margin = 0) {
    int g = kb_add_pointcloud(e_, pts.data(), (int)(pts.size() / 3), radius ? radius->data() : nullptr, margin); kbCheck(g); return g; }
  int AddSphere(const double c[3], double r, double margin = 0) { double p[4] = {c[0], c[1], c[2], r}; int g = kb_add_primitive(e_, KB_PRIM_SPHERE, p, margin); kbCheck(g); return g; }
  int AddPoint(const double c[3], double margin = 0) { int g = kb_add_primitive(e_, KB_PRIM_POINT, c, margin); kbCheck(g); return g; }
  int AddSegment(const double a[3], const double b[3], double margin = 0) { double p[6] = {a[0], a[1], a[2], b[0], b[1], b[2]}; int g = kb_add_primitive(e_, KB_PRIM_SEGMENT, p, margin); kbCheck(g); return g; }
  int AddTriangle(const double abc[9], double margin = 0) { int g = kb_add_primitive(e_, KB_PRIM_TRIANGLE, abc, margin); kbCheck(g); return g; }
  // solid boxes (GeometricPrimitive3D Box3D / AABB3D): centre, axes as the columns of a row-major 3x3, half dimensions / lo, hi
  int AddBox(const double center[3], const double axes[9], const double half[3], double margin = 0) {
    double p[15]; for (int k = 0; k < 3; k++) { p[k] = center[k]; p[12 + k] = half[k]; } for (int k = 0; k < 9; k++) p[3 + k] = axes[k];
    int g = kb_add_primitive(e_, KB_PRIM_BOX, p, margin); kbCheck(g); return g; }
  int AddAABB(const double lo[3], const double hi[3], double margin = 0) {
    double p[6] = {lo[0], lo[1], lo[2], hi[0], hi[1], hi[2]}; int g = kb_add_primitive(e_, KB_PRIM_AABB, p, margin); kbCheck(g); return g; }
  // a point cloud replaced between batches (kb_update_pointcloud on the finished engine rebuilds its hierarchy on the GPU)
  int AddDynamicPointCloud(int capacity, double radius = 0, double margin = 0) { int g = kb_add_dynamic_pointcloud(e_, capacity, radius, margin); kbCheck(g); return g; }
  int AddTerrain(int geom) { int i = kb_add_terrain(e_, geom); kbCheck(i); return i; }
  int AddRigidObject(int geom, const double T[12]) { int i = kb_add_rigid_object(e_, geom, T); kbCheck(i); return i; }
  void SetRobot(const RobotDescription& r) {
    const int L = (int)r.parents.size();
    kbCheck(kb_robot_create(e_, L, r.parents.data(), r.linkType.data(), r.axis.data(), r.T0Parent.data(), r.qMin.data(), r.qMax.data()));
    for (int i = 0; i < L; i++) kbCheck(kb_robot_set_link_geometry(e_, i, r.linkGeometry[i]));
    if (!r.jointType.empty()) kbCheck(kb_robot_set_joints(e_, (int)r.jointType.size(), r.jointType.data(), r.jointLink.data(), r.jointBase.empty() ? nullptr : r.jointBase.data()));
    robot_ = r;
  }
  // RobotModelDriver limits read by CheckJointLimits: value = mean_k (q[links[k]] - offset[k]) / scale[k]
  void AddDriver(const std::vector<int32_t>& links, const std::vector<double>* scale, const std::vector<double>* offset, double dmin, double dmax) {
    kbCheck(kb_robot_add_driver(e_, (int)links.size(), links.data(), scale ? scale->data() : nullptr, offset ? offset->data() : nullptr, dmin, dmax)); }
  void EnableSelfCollision(int i, int j, bool on) { kbCheck(kb_robot_set_self_collision(e_, i, j, on ? 1 : 0)); }
  // WorldPlannerSettings::collisionEnabled over world ids (row-major n x n)
  void SetCollisionEnabled(const std::vector<uint8_t>& mask, int n) { kbCheck(kb_set_pair_mask(e_, mask.data(), n)); }
  kb_engine* Finalize(int device = 0) { kbCheck(kb_finalize(e_, device)); kb_engine* r = e_; e_ = nullptr; return r; }
  // several GPUs behind one handle: IsFeasibleBatch / IsVisibleBatch shard their batches over `devices` (kb_finalize_multi)
  kb_engine* Finalize(const std::vector<int>& devices) { kbCheck(kb_finalize_multi(e_, devices.data(), (int)devices.size())); kb_engine* r = e_; e_ = nullptr; return r; }
  const RobotDescription& robot() const { return robot_; }
 private:
  kb_engine* e_ = nullptr;
  RobotDescription robot_;
};

class BatchSingleRobotCSpace;

// EdgePlanner face of EpsilonEdgeChecker (constructed by PathChecker; reference RobotCSpace.cpp:835-838)
class BatchEdgeChecker {
 public:
  BatchEdgeChecker(BatchSingleRobotCSpace* space, const Config& a, const Config& b, double eps) : space_(space), a_(a), b_(b), eps_(eps) {}
  bool IsVisible();
  double Length() const;
  const Config& Start() const { return a_; }
  const Config& End() const { return b_; }
  int numChecks = 0;
 private:
  BatchSingleRobotCSpace* space_; Config a_, b_; double eps_;
};
typedef std::shared_ptr<BatchEdgeChecker> EdgePlannerPtr;

class BatchSingleRobotCSpace {
 public:
  // takes ownership of a finalized engine
  BatchSingleRobotCSpace(kb_engine* engine, const RobotDescription& robot, double collisionEpsilon = 0.01)
      : collisionEpsilon(collisionEpsilon), engine_(engine), robot_(robot), rng_(12345) {
    if (robot_.jointType.empty()) { robot_.jointType.assign(robot_.parents.size(), KB_JOINT_NORMAL); robot_.jointLink.resize(robot_.parents.size()); for (size_t i = 0; i < robot_.jointLink.size(); i++) robot_.jointLink[i] = (int32_t)i; }
  }
  ~BatchSingleRobotCSpace() { if (engine_) kb_engine_destroy(engine_); }
  BatchSingleRobotCSpace(const BatchSingleRobotCSpace&) = delete;

  // ---- CSpace interface -----------------------------------------------------------------------------------------
  int NumDimensions() const { return (int)robot_.parents.size(); }
  bool IsFeasible(const Config& x) { uint8_t r = 0; IsFeasibleBatch(x.data(), 1, &r); lastConfig = x; return r != 0; }
  EdgePlannerPtr PathChecker(const Config& a, const Config& b) { return std::make_shared<BatchEdgeChecker>(this, a, b, collisionEpsilon); }
  // links a joint drives, root to tip (RobotModel::GetJointIndices, Cpp/Modeling/Robot.cpp:2120-2144)
  int JointIndices(size_t j, int idx[6]) const {
    int link = robot_.jointLink[j]; const int t = robot_.jointType[j];
    if (t == KB_JOINT_WELD || t == KB_JOINT_NORMAL || t == KB_JOINT_SPIN || robot_.jointBase.empty()) { idx[0] = link; return 1; }
    int n = 0, tmp[8];
    while (link != robot_.jointBase[j] && link >= 0 && n < 6) { tmp[n++] = link; link = robot_.parents[link]; }
    for (int i = 0; i < n; i++) idx[i] = tmp[n - 1 - i];
    return n;
  }
  // Euler ZYX triplet <-> matrix, SO(3) log / exp: what EulerAngleRotation::getMatrixZYX / setMatrixZYX, interpolateRotation and
  // AngleAxisRotation::angle compute for Floating / BallAndSocket joints (Interpolate.cpp:16-52,229-278)
  static void EulerZYXToMatrix(const double e[3], double R[9]) {
    const double ca = std::cos(e[0]), sa = std::sin(e[0]), cb = std::cos(e[1]), sb = std::sin(e[1]), cc = std::cos(e[2]), sc = std::sin(e[2]);
    R[0] = ca * cb; R[1] = ca * sb * sc - sa * cc; R[2] = ca * sb * cc + sa * sc;
    R[3] = sa * cb; R[4] = sa * sb * sc + ca * cc; R[5] = sa * sb * cc - ca * sc;
    R[6] = -sb; R[7] = cb * sc; R[8] = cb * cc;
  }
  static double RotationAngle(const double R[9]) { return std::acos(std::min(1.0, std::max(-1.0, 0.5 * (R[0] + R[4] + R[8] - 1.0)))); }
  static double EulerZYXAngleBetween(const double a[3], const double b[3]) {
    double Ra[9], Rb[9], D[9]; EulerZYXToMatrix(a, Ra); EulerZYXToMatrix(b, Rb);
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) D[3 * i + j] = Ra[3 * i] * Rb[3 * j] + Ra[3 * i + 1] * Rb[3 * j + 1] + Ra[3 * i + 2] * Rb[3 * j + 2];
    return RotationAngle(D);
  }
  static void EulerZYXInterp(const double a[3], const double b[3], double u, double out[3]) {
    double Ra[9], Rb[9], D[9], w[3], E[9], Ru[9];
    EulerZYXToMatrix(a, Ra); EulerZYXToMatrix(b, Rb);
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) D[3 * i + j] = Ra[i] * Rb[j] + Ra[3 + i] * Rb[3 + j] + Ra[6 + i] * Rb[6 + j];
    const double th = RotationAngle(D), v[3] = {D[7] - D[5], D[2] - D[6], D[3] - D[1]};
    if (th < 1e-9) for (int k = 0; k < 3; k++) w[k] = 0.5 * v[k];
    else if (M_PI - th < 1e-6) {
      double ax[3]; for (int k = 0; k < 3; k++) { const double d = 0.5 * (D[4 * k] + 1.0); ax[k] = d > 0 ? std::sqrt(d) : 0.0; }
      int m = 0; for (int k = 1; k < 3; k++) if (ax[k] > ax[m]) m = k;
      for (int k = 0; k < 3; k++) if (k != m && D[3 * m + k] + D[3 * k + m] < 0) ax[k] = -ax[k];
      if (ax[0] * v[0] + ax[1] * v[1] + ax[2] * v[2] < 0) for (int k = 0; k < 3; k++) ax[k] = -ax[k];
      const double n = std::sqrt(ax[0] * ax[0] + ax[1] * ax[1] + ax[2] * ax[2]); for (int k = 0; k < 3; k++) w[k] = th * ax[k] / n;
    } else { const double f = th / (2.0 * std::sin(th)); for (int k = 0; k < 3; k++) w[k] = f * v[k]; }
    for (int k = 0; k < 3; k++) w[k] *= u;
    const double t2 = std::sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
    if (t2 < 1e-12) { E[0] = 1; E[1] = -w[2]; E[2] = w[1]; E[3] = w[2]; E[4] = 1; E[5] = -w[0]; E[6] = -w[1]; E[7] = w[0]; E[8] = 1; }
    else {
      const double x = w[0] / t2, y = w[1] / t2, z = w[2] / t2, c = std::cos(t2), s = std::sin(t2), vv = 1.0 - c;
      E[0] = c + vv * x * x; E[1] = vv * x * y - s * z; E[2] = vv * x * z + s * y;
      E[3] = vv * y * x + s * z; E[4] = c + vv * y * y; E[5] = vv * y * z - s * x;
      E[6] = vv * z * x - s * y; E[7] = vv * z * y + s * x; E[8] = c + vv * z * z;
    }
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) Ru[3 * i + j] = Ra[3 * i] * E[j] + Ra[3 * i + 1] * E[3 + j] + Ra[3 * i + 2] * E[6 + j];
    out[1] = std::asin(std::min(1.0, std::max(-1.0, -Ru[6])));
    if (std::fabs(Ru[6]) < 1.0 - 1e-12) { out[0] = std::atan2(Ru[3], Ru[0]); out[2] = std::atan2(Ru[7], Ru[8]); }
    else { out[2] = 0.0; out[0] = std::atan2(-Ru[1], Ru[4]); }
  }
  // RobotCSpace::Distance -> Klampt::Distance, L2 over joints (Interpolate.cpp:208-343, norm = 2, floatingRotationWeight = 1)
  double Distance(const Config& x, const Config& y) const {
    double s = 0;
    for (size_t j = 0; j < robot_.jointType.size(); j++) {
      const int k = robot_.jointLink[j]; const double w = jointWeights.empty() ? 1.0 : jointWeights[j]; double d; int ix[6];
      const int t = robot_.jointType[j];
      if (t == KB_JOINT_NORMAL) d = x[k] - y[k];
      else if (t == KB_JOINT_SPIN) d = AngleDiff(AngleNormalize(x[k]), AngleNormalize(y[k]));
      else if (t == KB_JOINT_FLOATING || t == KB_JOINT_BALLANDSOCKET) {
        const int n = JointIndices(j, ix);
        if (t == KB_JOINT_FLOATING) for (int q = 0; q < 3; q++) { const double dt = x[ix[q]] - y[ix[q]]; s += w * dt * dt; }
        const double ea[3] = {x[ix[n - 3]], x[ix[n - 2]], x[ix[n - 1]]}, eb[3] = {y[ix[n - 3]], y[ix[n - 2]], y[ix[n - 1]]};
        d = EulerZYXAngleBetween(ea, eb);
      } else continue;
      s += w * d * d;
    }
    return std::sqrt(s);
  }
  // RobotCSpace::Interpolate -> Klampt::Interpolate (Interpolate.cpp:10-71)
  void Interpolate(const Config& x, const Config& y, double u, Config& out) const {
    out.resize(x.size());
    for (size_t k = 0; k < x.size(); k++) { out[k] = x[k] * (1.0 - u); out[k] += y[k] * u; }
    for (size_t j = 0; j < robot_.jointType.size(); j++) {
      const int t = robot_.jointType[j]; int ix[6];
      if (t == KB_JOINT_SPIN || t == KB_JOINT_FLOATINGPLANAR) {
        int k = robot_.jointLink[j]; if (t == KB_JOINT_FLOATINGPLANAR) { JointIndices(j, ix); k = ix[2]; }
        const double a = AngleNormalize(x[k]), b = AngleNormalize(y[k]);
        out[k] = AngleNormalize(a + u * AngleDiff(b, a));
      } else if (t == KB_JOINT_FLOATING || t == KB_JOINT_BALLANDSOCKET) {
        const int n = JointIndices(j, ix);
        const double ea[3] = {x[ix[n - 3]], x[ix[n - 2]], x[ix[n - 1]]}, eb[3] = {y[ix[n - 3]], y[ix[n - 2]], y[ix[n - 1]]}; double eu[3];
        EulerZYXInterp(ea, eb, u, eu);
        out[ix[n - 3]] = eu[0]; out[ix[n - 2]] = eu[1]; out[ix[n - 1]] = eu[2];
      }
    }
  }
  void Sample(Config& x) {   // uniform in [qMin,qMax] (RobotCSpace.cpp:85-87)
    x.resize(NumDimensions());
    for (int i = 0; i < NumDimensions(); i++) x[i] = std::uniform_real_distribution<double>(robot_.qMin[i], std::nextafter(robot_.qMax[i], 1e300))(rng_);
    for (size_t j = 0; j < fixedDofs.size(); j++) x[fixedDofs[j]] = fixedValues[j];
  }
  void SampleNeighborhood(const Config& c, double r, Config& x) {
    x.resize(c.size());
    for (size_t i = 0; i < c.size(); i++) {
      double lo = std::max(robot_.qMin[i], c[i] - r), hi = std::min(robot_.qMax[i], c[i] + r);
      x[i] = lo < hi ? std::uniform_real_distribution<double>(lo, hi)(rng_) : lo;
    }
    for (size_t j = 0; j < fixedDofs.size(); j++) x[fixedDofs[j]] = fixedValues[j];
  }
  void Properties(std::map<std::string, std::string>& props) const {   // RobotCSpace.cpp:220-279
    props["euclidean"] = "0"; props["geodesic"] = "1"; props["metric"] = jointWeights.empty() ? "euclidean" : "weighted euclidean";
    std::string lo, hi; double vol = 1;
    for (int i = 0; i < NumDimensions(); i++) { lo += std::to_string(robot_.qMin[i]) + " "; hi += std::to_string(robot_.qMax[i]) + " "; if (robot_.qMax[i] > robot_.qMin[i]) vol *= robot_.qMax[i] - robot_.qMin[i]; }
    props["minimum"] = lo; props["maximum"] = hi; props["volume"] = std::to_string(vol);
  }

  // ---- SingleRobotCSpace extras ---------------------------------------------------------------------------------
  bool CheckJointLimits(const Config& x) const {   // RobotCSpace.cpp:610-630 (drivers are checked by the engine)
    for (size_t j = 0; j < robot_.jointType.size(); j++) if (robot_.jointType[j] == KB_JOINT_NORMAL || robot_.jointType[j] == KB_JOINT_WELD) {
      const int k = robot_.jointLink[j]; if (x[k] < robot_.qMin[k] || x[k] > robot_.qMax[k]) return false; }
    return true;
  }
  bool CheckCollisionFree(const Config& x) { return !CheckJointLimits(x) ? DistanceToObstacles(x, 0.0) > 0.0 : IsFeasible(x); }
  void GetJointLimits(Config& bmin, Config& bmax) const { bmin = robot_.qMin; bmax = robot_.qMax; }
  void FixDof(int dof, double value) { fixedDofs.push_back(dof); fixedValues.push_back(value); }
  // min over enabled robot-environment (+ self) pairs, capped (WorldPlannerSettings::DistanceLowerBound)
  double DistanceToObstacles(const Config& x, double upperBound, bool includeSelf = true) {
    double d = 0; kbCheck(kb_distance_batch(engine_, x.data(), 1, upperBound > 0 ? upperBound : 1e-12, includeSelf ? 1 : 0, &d, nullptr)); return d; }

  // ---- named constraints (CSpace::constraints / constraintNames and IsFeasible(x, constraint); SingleRobotCSpace::Init,
  // RobotCSpace.cpp:668-754): "<link>_joint_limit" for every link with a finite limit, "update_geometry", then one
  // "coll[idA,idB]" per enabled (robot link, environment) pair and self pair, in world-id order.  The collision constraints of one
  // configuration are all answered by ONE all-pairs launch (the reference evaluates its CollisionFreeSets one query at a time).
  struct Constraint { int kind; int a, b; };           // kind 0: joint limit of link a; 1: update_geometry (always true); 2: collision pair of world ids a < b
  std::vector<std::string> constraintNames;
  void InitConstraints(const std::vector<std::string>* worldNames = nullptr) {
    constraints_.clear(); constraintNames.clear();
    const int L = NumDimensions();
    auto idName = [&](int id) { return worldNames && id < (int)worldNames->size() ? (*worldNames)[id] : "id" + std::to_string(id); };
    const int n = kb_num_ids(engine_), base = n - L;
    for (int i = 0; i < L; i++) if (std::isfinite(robot_.qMin[i]) || std::isfinite(robot_.qMax[i])) {
      constraints_.push_back({0, i, -1}); constraintNames.push_back(idName(base + i) + "_joint_limit"); }
    constraints_.push_back({1, -1, -1}); constraintNames.push_back("update_geometry");
    std::vector<uint8_t> mask((size_t)n * n); kbCheck(kb_get_pair_mask(engine_, mask.data()));
    for (int i = base; i < n; i++) for (int j = 0; j < n; j++) {
      if (j == base - 1 || (j >= base && j <= i)) continue;                    // the robot id itself; self pairs once, lower link first
      if (robot_.linkGeometry[i - base] < 0 || (j >= base && robot_.linkGeometry[j - base] < 0)) continue;
      if (!(mask[(size_t)i * n + j] || mask[(size_t)j * n + i])) continue;
      constraints_.push_back({2, std::min(i, j), std::max(i, j)});
      constraintNames.push_back("coll[" + idName(i) + "," + idName(j) + "]");
    }
  }
  int NumConstraints() const { return (int)constraints_.size(); }
  // indices of the constraints that fail at x (CSpace::FeasibilityFailures / CSpaceInterface::feasibilityFailures)
  void FeasibilityFailures(const Config& x, std::vector<int>& failed) {
    failed.clear();
    int32_t pairs[64], count = 0;
    bool limitsOk = true;
    for (size_t c = 0; c < constraints_.size(); c++) if (constraints_[c].kind == 0) {
      const int k = constraints_[c].a; if (x[k] < robot_.qMin[k] || x[k] > robot_.qMax[k]) { failed.push_back((int)c); limitsOk = false; } }
    (void)limitsOk;
    kbCheck(kb_colliding_pairs_batch(engine_, x.data(), 1, 32, pairs, &count));
    // a configuration outside its limits is not traversed by the engine (count = -1): ask again with the limits ignored is not
    // possible through this handle, so its collision constraints are reported as unknown = not failed
    for (int k = 0; k < (count > 0 ? count : 0); k++) {
      const int a = std::min(pairs[2 * k], pairs[2 * k + 1]), b = std::max(pairs[2 * k], pairs[2 * k + 1]);
      for (size_t c = 0; c < constraints_.size(); c++) if (constraints_[c].kind == 2 && constraints_[c].a == a && constraints_[c].b == b) failed.push_back((int)c);
    }
    lastConfig = x;
  }
  bool IsFeasible(const Config& x, int constraint) {
    if (constraint < 0 || constraint >= NumConstraints()) throw std::out_of_range("constraint index");
    const Constraint& c = constraints_[constraint];
    if (c.kind == 0) return x[c.a] >= robot_.qMin[c.a] && x[c.a] <= robot_.qMax[c.a];
    if (c.kind == 1) return true;
    std::vector<int> failed; FeasibilityFailures(x, failed);
    return std::find(failed.begin(), failed.end(), constraint) == failed.end();
  }

  // ---- batch entry points ---------------------------------------------------------------------------------------
  void IsFeasibleBatch(const double* Q, int64_t N, uint8_t* out, int32_t* firstPair = nullptr) { kbCheck(kb_feasible_batch(engine_, Q, N, out, firstPair)); }
  void IsVisibleBatch(const double* A, const double* B, int64_t N, double eps, uint8_t* out, int32_t* nchecks = nullptr) {
    kbCheck(kb_edges_visible_batch(engine_, A, B, N, eps, jointWeights.empty() ? nullptr : jointWeights.data(), out, nchecks)); }
  // every colliding (idA, idB) pair per configuration: the per-pair CollisionFreeSet constraints of Init() evaluated together
  void CollidingPairsBatch(const double* Q, int64_t N, int maxPairs, int32_t* outPairs, int32_t* outCount) { kbCheck(kb_colliding_pairs_batch(engine_, Q, N, maxPairs, outPairs, outCount)); }
  void DistanceBatch(const double* Q, int64_t N, double upperBound, bool includeSelf, double* out) { kbCheck(kb_distance_batch(engine_, Q, N, upperBound, includeSelf ? 1 : 0, out, nullptr)); }
  // WorldModel::RayCast / RayCastIgnore (Cpp/Modeling/World.cpp:465-588) for N rays (source xyz, direction xyz per row) with the robot
  // at x: ids[i] = world id hit or -1, dist[i] = distance along the unit direction (inf = nothing); ignoreIDs as RayCastIgnore's list
  void RayCastBatch(const Config& x, const double* rays, int64_t N, int32_t* ids, double* dist, const std::vector<int>* ignoreIDs = nullptr, int32_t* elems = nullptr) {
    std::vector<uint8_t> ig;
    if (ignoreIDs) { ig.assign((size_t)kb_num_ids(engine_), 0); for (int id : *ignoreIDs) if (id >= 0 && (size_t)id < ig.size()) ig[(size_t)id] = 1; }
    kbCheck(kb_raycast_batch(engine_, x.data(), rays, N, ignoreIDs ? ig.data() : nullptr, ids, dist, elems));
  }
  // the camera sensor's ray-cast rendering (Cpp/Sensing/VisualSensors.cpp:413-475) in one call: depth along the viewing direction per
  // pixel (zmax where nothing is seen) and the world id per pixel
  void CameraDepth(const Config& x, const kb_camera& cam, float* depth, int32_t* ids = nullptr) { kbCheck(kb_camera_depth(engine_, x.data(), &cam, nullptr, depth, ids)); }
  kb_stats GetStats() { kb_stats s; kbCheck(kb_get_stats(engine_, &s)); return s; }
  kb_engine* engine() { return engine_; }

  double collisionEpsilon;                 // settings->robotSettings[index].collisionEpsilon (PlannerSettings.cpp:86-87)
  std::vector<double> jointWeights;        // RobotCSpace::jointWeights
  std::vector<int> fixedDofs; std::vector<double> fixedValues;
  Config lastConfig;

  static double AngleNormalize(double a) { a = std::fmod(a, 2 * M_PI); return a < 0 ? a + 2 * M_PI : a; }
  static double AngleDiff(double a, double b) { double d = a - b; return d > M_PI ? d - 2 * M_PI : (d < -M_PI ? d + 2 * M_PI : d); }

 private:
  kb_engine* engine_;
  RobotDescription robot_;
  std::vector<Constraint> constraints_;
  std::mt19937_64 rng_;
};

inline bool BatchEdgeChecker::IsVisible() {
  uint8_t v = 0; int32_t n = 0;
  space_->IsVisibleBatch(a_.data(), b_.data(), 1, eps_, &v, &n);
  numChecks = n;
  return v != 0;
}
inline double BatchEdgeChecker::Length() const { return space_->Distance(a_, b_); }

}  // namespace klampt_b200
