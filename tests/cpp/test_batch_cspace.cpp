// Exercises the C++ adapter end to end: reads a small world dumped by the Python test (binary), runs IsFeasibleBatch /
// IsVisibleBatch / IsFeasible / PathChecker on the GPU and writes the results back for comparison with the oracle.
#include "klampt_b200/BatchSingleRobotCSpace.h"
#include <cstdio>
#include <cstdlib>
#include <cmath>
using namespace klampt_b200;

template <class T> static std::vector<T> rd(FILE* f) { int64_t n = 0; if (fread(&n, 8, 1, f) != 1) exit(3); std::vector<T> v((size_t)n); if (n && fread(v.data(), sizeof(T), (size_t)n, f) != (size_t)n) exit(3); return v; }

int main(int argc, char** argv) {
  if (argc < 3) return 2;
  FILE* f = fopen(argv[1], "rb"); if (!f) return 2;
  WorldBuilder wb;
  RobotDescription r;
  int64_t ngeom = rd<int64_t>(f)[0];
  for (int64_t g = 0; g < ngeom; g++) { auto v = rd<double>(f); auto t = rd<int32_t>(f); wb.AddTriMesh(v, t); }
  auto terr = rd<int32_t>(f); for (int g : terr) wb.AddTerrain(g);
  auto objg = rd<int32_t>(f); auto objT = rd<double>(f);
  for (size_t i = 0; i < objg.size(); i++) wb.AddRigidObject(objg[i], &objT[12 * i]);
  r.parents = rd<int32_t>(f); r.linkType = rd<uint8_t>(f); r.axis = rd<double>(f); r.T0Parent = rd<double>(f); r.qMin = rd<double>(f); r.qMax = rd<double>(f);
  auto lg = rd<int32_t>(f); r.linkGeometry.assign(lg.begin(), lg.end());
  r.jointType = rd<uint8_t>(f); r.jointLink = rd<int32_t>(f);
  auto Q = rd<double>(f); auto A = rd<double>(f); auto B = rd<double>(f);
  fclose(f);
  wb.SetRobot(r);
  BatchSingleRobotCSpace space(wb.Finalize(0), r, 0.01);
  const int L = space.NumDimensions();
  const int64_t N = (int64_t)Q.size() / L, E = (int64_t)A.size() / L;
  std::vector<uint8_t> feas((size_t)N), vis((size_t)E); std::vector<int32_t> nchk((size_t)E);
  space.IsFeasibleBatch(Q.data(), N, feas.data());
  space.IsVisibleBatch(A.data(), B.data(), E, space.collisionEpsilon, vis.data(), nchk.data());
  // single-configuration face must agree with the batch
  for (int64_t i = 0; i < 20 && i < N; i++) { Config x(Q.begin() + i * L, Q.begin() + (i + 1) * L); if (space.IsFeasible(x) != (feas[i] != 0)) return 4; }
  for (int64_t i = 0; i < 10 && i < E; i++) {
    Config a(A.begin() + i * L, A.begin() + (i + 1) * L), b(B.begin() + i * L, B.begin() + (i + 1) * L);
    auto e = space.PathChecker(a, b);
    if (e->IsVisible() != (vis[i] != 0) || e->numChecks != nchk[i]) return 5;
  }
  // named constraints: a configuration is feasible iff none of its constraints fails; IsFeasible(x, c) agrees with the list
  space.InitConstraints();
  if (space.NumConstraints() < 3 || space.constraintNames.size() != (size_t)space.NumConstraints()) return 8;
  for (int64_t i = 0; i < 30 && i < N; i++) {
    Config x(Q.begin() + i * L, Q.begin() + (i + 1) * L);
    std::vector<int> failed; space.FeasibilityFailures(x, failed);
    if (failed.empty() != (feas[i] != 0)) return 9;
    for (int c : failed) if (space.IsFeasible(x, c)) return 10;
    if (!failed.empty() && space.constraintNames[failed[0]].rfind("coll[", 0) != 0 && space.constraintNames[failed[0]].find("_joint_limit") == std::string::npos) return 11;
  }
  Config s; space.Sample(s); if (!space.CheckJointLimits(s)) return 6;
  std::map<std::string, std::string> props; space.Properties(props); if (props["geodesic"] != "1") return 7;
  // ray casts: a fan of rays from above the scene, with the robot at the first configuration; the second pass ignores whatever the first ray hit
  const int R = 400;
  std::vector<double> rays((size_t)R * 6);
  for (int i = 0; i < R; i++) {
    const double a = 2.0 * M_PI * i / R, rad = 0.2 + 1.3 * ((i * 37) % R) / (double)R;
    rays[6 * i] = 0.3; rays[6 * i + 1] = -0.2; rays[6 * i + 2] = 3.0;
    rays[6 * i + 3] = rad * std::cos(a) - 0.3; rays[6 * i + 4] = rad * std::sin(a) + 0.2; rays[6 * i + 5] = -3.0;
  }
  Config x0(Q.begin(), Q.begin() + L);
  std::vector<int32_t> rid((size_t)R), rid2((size_t)R); std::vector<double> rdist((size_t)R), rdist2((size_t)R);
  space.RayCastBatch(x0, rays.data(), R, rid.data(), rdist.data());
  std::vector<int> ign; if (rid[0] >= 0) ign.push_back(rid[0]);
  space.RayCastBatch(x0, rays.data(), R, rid2.data(), rdist2.data(), &ign);
  if (!ign.empty()) for (int i = 0; i < R; i++) if (rid2[i] == ign[0]) return 12;
  FILE* o = fopen(argv[2], "wb"); fwrite(feas.data(), 1, feas.size(), o); fwrite(vis.data(), 1, vis.size(), o); fwrite(nchk.data(), 4, nchk.size(), o);
  fwrite(rays.data(), 8, rays.size(), o); fwrite(rid.data(), 4, rid.size(), o); fwrite(rdist.data(), 8, rdist.size(), o); fclose(o);
  kb_stats st = space.GetStats();
  printf("ok N=%lld E=%lld launches=%lld\n", (long long)N, (long long)E, (long long)st.kernel_launches);
  return 0;
}
