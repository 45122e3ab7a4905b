// ceres_crosscheck — pins the parity boundary where Ceres exists (it does not in the build container).
//
// NOT part of the product and not built by __graft_entry__.build(): this is the optional off-box tool SURVEY.md §8(c)
// asks for.  It solves a scene exported by tools/export_scene_text.py with the real libceres, configured the way
// GL-SLAM configures it (slam_core.cpp:799-849: AutoDiff<2,6,3> residual, CauchyLoss(1.0), cameras 0/1 constant,
// SPARSE_SCHUR, 30 iterations), and prints the per-iteration trajectory as JSON.  Commit that JSON as
// tests/golden/ceres/<name>.json; tests/test_ceres_pin.py then checks the oracle (and, under -m gpu, the CUDA path)
// against it with the north-star tolerances, and the "parity unpinned" notes can be dropped.
//
//   g++ -O2 -std=c++17 tools/ceres_crosscheck.cpp -o ceres_crosscheck $(pkg-config --cflags --libs ceres eigen3) -lglog
//   python tools/export_scene_text.py            # writes tests/golden/text/<name>.txt
//   ./ceres_crosscheck tests/golden/text/window10.txt [--loss cauchy|huber|none] [--threads 1] [--dense] > tests/golden/ceres/window10.json
//
// Scene text format (one token stream): n_cam n_pt n_obs fx fy cx cy loss | n_cam x (fixed w0 w1 w2 c0 c1 c2) |
// n_pt x (X Y Z) | n_obs x (cam pt u v).  Values are printed with 17 significant digits, so they round-trip exactly.
#include <ceres/ceres.h>
#include <ceres/rotation.h>

#include <cstdio>
#include <cstring>
#include <fstream>
#include <iostream>
#include <string>
#include <vector>

namespace {

// camera block = [angle-axis of R_wc | camera centre], point in world coordinates; residual = projection - observation
struct PinholeResidual {
  PinholeResidual(double u, double v, const double* k) : u_(u), v_(v), fx_(k[0]), fy_(k[1]), cx_(k[2]), cy_(k[3]) {}
  template <typename T>
  bool operator()(const T* const cam, const T* const pt, T* res) const {
    const T d[3] = {pt[0] - cam[3], pt[1] - cam[4], pt[2] - cam[5]};
    const T minus_w[3] = {-cam[0], -cam[1], -cam[2]};
    T pc[3];
    ceres::AngleAxisRotatePoint(minus_w, d, pc);
    res[0] = T(fx_) * pc[0] / pc[2] + T(cx_) - T(u_);
    res[1] = T(fy_) * pc[1] / pc[2] + T(cy_) - T(v_);
    return true;
  }
  double u_, v_, fx_, fy_, cx_, cy_;
};

}  // namespace

int main(int argc, char** argv) {
  if (argc < 2) { std::fprintf(stderr, "usage: %s scene.txt [--loss cauchy|huber|none] [--threads N] [--dense]\n", argv[0]); return 2; }
  std::string loss_name;
  int threads = 1;
  bool dense = false;
  for (int i = 2; i < argc; ++i) {
    if (!std::strcmp(argv[i], "--loss") && i + 1 < argc) loss_name = argv[++i];
    else if (!std::strcmp(argv[i], "--threads") && i + 1 < argc) threads = std::atoi(argv[++i]);
    else if (!std::strcmp(argv[i], "--dense")) dense = true;
  }
  std::ifstream in(argv[1]);
  long n_cam = 0, n_pt = 0, n_obs = 0;
  double K[4];
  int loss_code = 2;
  if (!(in >> n_cam >> n_pt >> n_obs >> K[0] >> K[1] >> K[2] >> K[3] >> loss_code)) { std::fprintf(stderr, "bad header\n"); return 2; }
  if (loss_name.empty()) loss_name = loss_code == 2 ? "cauchy" : loss_code == 1 ? "huber" : "none";
  std::vector<double> cam(6 * n_cam), pt(3 * n_pt);
  std::vector<int> fixed(n_cam);
  for (long i = 0; i < n_cam; ++i) { in >> fixed[i]; for (int k = 0; k < 6; ++k) in >> cam[6 * i + k]; }
  for (long j = 0; j < 3 * n_pt; ++j) in >> pt[j];
  std::vector<long> oc(n_obs), op(n_obs);
  std::vector<double> ou(n_obs), ov(n_obs);
  for (long k = 0; k < n_obs; ++k) in >> oc[k] >> op[k] >> ou[k] >> ov[k];
  if (!in) { std::fprintf(stderr, "truncated scene\n"); return 2; }

  ceres::Problem problem;
  for (long k = 0; k < n_obs; ++k) {
    ceres::CostFunction* cost = new ceres::AutoDiffCostFunction<PinholeResidual, 2, 6, 3>(new PinholeResidual(ou[k], ov[k], K));
    ceres::LossFunction* loss = nullptr;
    if (loss_name == "cauchy") loss = new ceres::CauchyLoss(1.0);
    else if (loss_name == "huber") loss = new ceres::HuberLoss(1.0);
    problem.AddResidualBlock(cost, loss, &cam[6 * oc[k]], &pt[3 * op[k]]);
  }
  for (long i = 0; i < n_cam; ++i)
    if (fixed[i] && problem.HasParameterBlock(&cam[6 * i])) problem.SetParameterBlockConstant(&cam[6 * i]);

  ceres::Solver::Options options;
  options.linear_solver_type = dense ? ceres::DENSE_SCHUR : ceres::SPARSE_SCHUR;
  options.max_num_iterations = 30;
  options.num_threads = threads;
  options.minimizer_progress_to_stdout = false;
  options.logging_type = ceres::SILENT;
  ceres::Solver::Summary summary;
  ceres::Solve(options, &problem, &summary);

  std::printf("{\"ceres_version\": \"%s\", \"loss\": \"%s\", \"linear_solver\": \"%s\", \"threads\": %d,\n", CERES_VERSION_STRING, loss_name.c_str(),
              dense ? "DENSE_SCHUR" : "SPARSE_SCHUR", threads);
  std::printf(" \"termination_type\": %d, \"usable\": %d, \"initial_cost\": %.17g, \"final_cost\": %.17g,\n", (int)summary.termination_type,
              summary.IsSolutionUsable() ? 1 : 0, summary.initial_cost, summary.final_cost);
  std::printf(" \"message\": \"%s\",\n \"iterations\": [\n", summary.message.c_str());
  for (size_t i = 0; i < summary.iterations.size(); ++i) {
    const ceres::IterationSummary& it = summary.iterations[i];
    std::printf("  {\"iteration\": %d, \"cost\": %.17g, \"cost_change\": %.17g, \"gradient_max_norm\": %.17g, \"step_norm\": %.17g, "
                "\"relative_decrease\": %.17g, \"trust_region_radius\": %.17g, \"step_is_valid\": %d, \"step_is_successful\": %d}%s\n",
                it.iteration, it.cost, it.cost_change, it.gradient_max_norm, it.step_norm, it.relative_decrease, it.trust_region_radius,
                it.step_is_valid ? 1 : 0, it.step_is_successful ? 1 : 0, i + 1 < summary.iterations.size() ? "," : "");
  }
  std::printf(" ],\n \"cam\": [");
  for (long i = 0; i < 6 * n_cam; ++i) std::printf("%s%.17g", i ? ", " : "", cam[i]);
  std::printf("],\n \"pt\": [");
  for (long j = 0; j < 3 * n_pt; ++j) std::printf("%s%.17g", j ? ", " : "", pt[j]);
  std::printf("]}\n");
  return 0;
}
