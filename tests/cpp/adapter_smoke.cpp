// Compiles the adapter header with plain g++ (no OpenCV) and drives every adapter once through the C ABI.
// Build: g++ -std=c++17 -I include tests/cpp/adapter_smoke.cpp -L omni-swarm_b200/csrc -lomniswarm_b200 -o adapter_smoke
// Without a GPU it only checks that create() reports OSB_ERR_NO_DEVICE; with one it runs a tiny end-to-end pass.
#include <cmath>
#include <cstring>
#include <random>
#include "omniswarm_b200_adapters.hpp"

int main() {
  std::printf("%s, devices: %d\n", osb_version(), osb_device_count());
  if (osb_device_count() == 0) {
    osb_db* h = nullptr;
    return osb_db_create(&h, 64, 16) == OSB_ERR_NO_DEVICE ? 0 : 1;
  }
  std::mt19937 rng(0);
  std::normal_distribution<float> N(0.f, 1.f);
  // --- database + matcher
  osb::IndexFlatIPB200 index(64, 128);
  std::vector<float> rows(100 * 64);
  for (auto& v : rows) v = N(rng);
  index.add(100, rows.data());
  float D[5]; osb::IndexFlatIPB200::idx_t I[5];
  index.search(1, rows.data() + 17 * 64, 5, D, I);
  if (index.ntotal != 100 || I[0] != 17) { std::printf("db search wrong: %ld\n", (long)I[0]); return 2; }
  osb::BFMatcherB200 bf;
  std::vector<osb::DMatchB200> m;
  bf.match(rows.data(), 40, rows.data(), 40, m);
  if (m.size() != 40 || m[7].trainIdx != 7 || m[7].distance != 0.f) { std::printf("matcher wrong\n"); return 3; }
  // --- pose graph: two poses, one odometry edge, first pose constant (solver.cpp:1196-1199)
  osb_solver* solver = nullptr;
  osb::check(osb_solver_create(&solver, 16, 16), "osb_solver_create");
  double a[4] = {0, 0, 0, 0}, b[4] = {0.7, 0.1, -0.2, 0.05};
  const double meas[4] = {1.0, 0.0, 0.0, 0.1};
  double S[16] = {0}; for (int i = 0; i < 4; ++i) S[i * 5] = 10.0;
  osb::FlatPoseGraph g;
  g.add_relative_pose(a, b, meas, S, false);
  g.set_constant(a);
  osb_solve_summary s = g.solve(solver);
  {
    // resident variant: two poses first, a third pose and its edge appended after the first solve
    double c0[4] = {0, 0, 0, 0}, c1[4] = {0.9, 0.1, 0, 0.05}, c2[4] = {2.2, 0, 0, 0.25};
    osb::ResidentPoseGraph rg(solver);
    rg.add_relative_pose(c0, c1, meas, S, false);
    rg.set_constant(c0);
    rg.solve();
    rg.add_relative_pose(c1, c2, meas, S, false);
    osb_solve_summary rs = rg.solve();
    const double ex = 1.0 + std::cos(0.1), ey = std::sin(0.1);
    if (std::fabs(c1[0] - 1.0) > 1e-6 || std::fabs(c2[0] - ex) > 1e-6 || std::fabs(c2[1] - ey) > 1e-6 || std::fabs(c2[3] - 0.2) > 1e-6) {
      std::printf("resident solve wrong: %f %f | %f %f %f cost %g\n", c1[0], c1[3], c2[0], c2[1], c2[3], rs.final_cost);
      return 5;
    }
  }
  {
    // bearing detection (DroneDetection4dFactor through the adapter, solver.cpp:1088-1094): drone B is seen from A along +x at
    // 2 m; with the odometry edge pulling it elsewhere in y the detection's tangent residual drags it back towards y = 0
    double d0[4] = {0, 0, 0, 0}, d1[4] = {1.8, 0.6, 0.1, 0.0};
    const double dir[3] = {1, 0, 0}, tan_base[6] = {0, 1, 0, 0, 0, 1};
    osb::FlatPoseGraph dg;
    dg.add_detection(d0, d1, dir, tan_base, /*inv_dep=*/0.5, /*enable_depth=*/true, /*extrinsic_z=*/0.0, nullptr, nullptr,
                     /*sphere_std=*/0.01, /*inv_dep_std=*/0.01, false);
    dg.set_constant(d0);
    osb_solve_summary ds = dg.solve(solver);
    if (dg.num_factors() != 1 || ds.n_residuals != 3 || std::fabs(d1[0] - 2.0) > 1e-4 || std::fabs(d1[1]) > 1e-4 || std::fabs(d1[2]) > 1e-4) {
      std::printf("detection solve wrong: %f %f %f (%d residuals, cost %g)\n", d1[0], d1[1], d1[2], ds.n_residuals, ds.final_cost);
      return 6;
    }
  }
  osb_solver_destroy(solver);
  if (std::fabs(b[0] - 1.0) > 1e-6 || std::fabs(b[3] - 0.1) > 1e-6 || a[0] != 0.0) {
    std::printf("solve wrong: %f %f %f %f cost %g\n", b[0], b[1], b[2], b[3], s.final_cost);
    return 4;
  }
  std::printf("adapters ok: solve %d iterations, final cost %.3g\n", s.iterations, s.final_cost);
  return 0;
}
