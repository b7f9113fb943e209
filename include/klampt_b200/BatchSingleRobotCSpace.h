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
  int AddTerrain(int geom) { int i = kb_add_terrain(e_, geom); kbCheck(i); return i; }
  int AddRigidObject(int geom, const double T[12]) { int i = kb_add_rigid_object(e_, geom, T); kbCheck(i); return i; }
  void SetRobot(const RobotDescription& r) {
    const int L = (int)r.parents.size();
    kbCheck(kb_robot_create(e_, L, r.parents.data(), r.linkType.data(), r.axis.data(), r.T0Parent.data(), r.qMin.data(), r.qMax.data()));
    for (int i = 0; i < L; i++) kbCheck(kb_robot_set_link_geometry(e_, i, r.linkGeometry[i]));
    if (!r.jointType.empty()) kbCheck(kb_robot_set_joints(e_, (int)r.jointType.size(), r.jointType.data(), r.jointLink.data()));
    robot_ = r;
  }
  // RobotModelDriver limits read by CheckJointLimits: value = mean_k (q[links[k]] - offset[k]) / scale[k]
  void AddDriver(const std::vector<int32_t>& links, const std::vector<double>* scale, const std::vector<double>* offset, double dmin, double dmax) {
    kbCheck(kb_robot_add_driver(e_, (int)links.size(), links.data(), scale ? scale->data() : nullptr, offset ? offset->data() : nullptr, dmin, dmax)); }
  void EnableSelfCollision(int i, int j, bool on) { kbCheck(kb_robot_set_self_collision(e_, i, j, on ? 1 : 0)); }
  // WorldPlannerSettings::collisionEnabled over world ids (row-major n x n)
  void SetCollisionEnabled(const std::vector<uint8_t>& mask, int n) { kbCheck(kb_set_pair_mask(e_, mask.data(), n)); }
  kb_engine* Finalize(int device = 0) { kbCheck(kb_finalize(e_, device)); kb_engine* r = e_; e_ = nullptr; return r; }
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
  // RobotCSpace::Distance -> Klampt::Distance, L2 over joints (Interpolate.cpp:208-343, norm = 2)
  double Distance(const Config& x, const Config& y) const {
    double s = 0;
    for (size_t j = 0; j < robot_.jointType.size(); j++) {
      const int k = robot_.jointLink[j]; const double w = jointWeights.empty() ? 1.0 : jointWeights[j]; double d;
      if (robot_.jointType[j] == KB_JOINT_NORMAL) d = x[k] - y[k];
      else if (robot_.jointType[j] == KB_JOINT_SPIN) d = AngleDiff(AngleNormalize(x[k]), AngleNormalize(y[k]));
      else continue;
      s += w * d * d;
    }
    return std::sqrt(s);
  }
  // RobotCSpace::Interpolate -> Klampt::Interpolate (Interpolate.cpp:10-71)
  void Interpolate(const Config& x, const Config& y, double u, Config& out) const {
    out.resize(x.size());
    for (size_t k = 0; k < x.size(); k++) { out[k] = x[k] * (1.0 - u); out[k] += y[k] * u; }
    for (size_t j = 0; j < robot_.jointType.size(); j++) if (robot_.jointType[j] == KB_JOINT_SPIN) {
      const int k = robot_.jointLink[j]; const double a = AngleNormalize(x[k]), b = AngleNormalize(y[k]);
      out[k] = AngleNormalize(a + u * AngleDiff(b, a));
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

  // ---- batch entry points ---------------------------------------------------------------------------------------
  void IsFeasibleBatch(const double* Q, int64_t N, uint8_t* out, int32_t* firstPair = nullptr) { kbCheck(kb_feasible_batch(engine_, Q, N, out, firstPair)); }
  void IsVisibleBatch(const double* A, const double* B, int64_t N, double eps, uint8_t* out, int32_t* nchecks = nullptr) {
    kbCheck(kb_edges_visible_batch(engine_, A, B, N, eps, jointWeights.empty() ? nullptr : jointWeights.data(), out, nchecks)); }
  // every colliding (idA, idB) pair per configuration: the per-pair CollisionFreeSet constraints of Init() evaluated together
  void CollidingPairsBatch(const double* Q, int64_t N, int maxPairs, int32_t* outPairs, int32_t* outCount) { kbCheck(kb_colliding_pairs_batch(engine_, Q, N, maxPairs, outPairs, outCount)); }
  void DistanceBatch(const double* Q, int64_t N, double upperBound, bool includeSelf, double* out) { kbCheck(kb_distance_batch(engine_, Q, N, upperBound, includeSelf ? 1 : 0, out, nullptr)); }
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
