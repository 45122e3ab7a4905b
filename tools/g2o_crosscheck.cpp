// g2o_crosscheck — pins GLBA_MODE_G2O where g2o exists (it does not in the build container).
//
// NOT part of the product and not built by __graft_entry__.build().  Solves a scene exported by
// tools/export_scene_text.py with the real g2o, set up the way the reference's archived BA sets it up
// (Old/mult_img_recoverpose_single_ba:251-326, docs/old_unorganized/4image_pnp_ba.txt:321-430: VertexSE3Expmap poses,
// marginalised point vertices, EdgeProjectXYZ2UV with CameraParameters(fx, (cx, cy), 0), unit information,
// OptimizationAlgorithmLevenberg over a dense 6x3 block solver, camera 0 fixed) and prints the per-iteration statistics
// as JSON.  Commit that as tests/golden/g2o/<name>.json; tests/test_ceres_pin.py then holds the oracle's g2o mode (and
// the CUDA path under -m gpu) to it.
//
//   g++ -O2 -std=c++17 tools/g2o_crosscheck.cpp -o g2o_crosscheck -I<g2o>/include -I/usr/include/eigen3 \
//       -L<g2o>/lib -lg2o_core -lg2o_stuff -lg2o_types_sba -lg2o_types_slam3d -lg2o_solver_dense
//   ./g2o_crosscheck tests/golden/text/window10.txt [--iterations 12] [--huber 3.0] > tests/golden/g2o/window10.json
//
// The scene file holds camera-to-world poses [angle-axis(R_wc) | centre] (the live path's convention); they are
// converted to world-to-camera here, as scene.as_g2o does.  The file's fixed flags are ignored: camera 0 is fixed, as in
// the archived code.  fy is ignored (one focal length, CameraParameters).
#include <g2o/core/batch_stats.h>
#include <g2o/core/block_solver.h>
#include <g2o/core/optimization_algorithm_levenberg.h>
#include <g2o/core/robust_kernel_impl.h>
#include <g2o/core/sparse_optimizer.h>
#include <g2o/solvers/dense/linear_solver_dense.h>
#include <g2o/types/sba/types_six_dof_expmap.h>

#include <Eigen/Geometry>

#include <cstdio>
#include <cstring>
#include <fstream>
#include <memory>
#include <vector>

int main(int argc, char** argv) {
  if (argc < 2) { std::fprintf(stderr, "usage: %s scene.txt [--iterations N] [--huber delta]\n", argv[0]); return 2; }
  int iterations = 12;
  double huber = 0.0;
  for (int i = 2; i < argc; ++i) {
    if (!std::strcmp(argv[i], "--iterations") && i + 1 < argc) iterations = std::atoi(argv[++i]);
    else if (!std::strcmp(argv[i], "--huber") && i + 1 < argc) huber = std::atof(argv[++i]);
  }
  std::ifstream in(argv[1]);
  long n_cam = 0, n_pt = 0, n_obs = 0;
  double K[4];
  int loss_code = 0;
  if (!(in >> n_cam >> n_pt >> n_obs >> K[0] >> K[1] >> K[2] >> K[3] >> loss_code)) { std::fprintf(stderr, "bad header\n"); return 2; }
  std::vector<double> cam(6 * n_cam), pt(3 * n_pt);
  for (long i = 0; i < n_cam; ++i) { int fixed_ignored; in >> fixed_ignored; for (int k = 0; k < 6; ++k) in >> cam[6 * i + k]; }
  for (long j = 0; j < 3 * n_pt; ++j) in >> pt[j];
  std::vector<long> oc(n_obs), op(n_obs);
  std::vector<double> ou(n_obs), ov(n_obs);
  for (long k = 0; k < n_obs; ++k) in >> oc[k] >> op[k] >> ou[k] >> ov[k];
  if (!in) { std::fprintf(stderr, "truncated scene\n"); return 2; }

  typedef g2o::BlockSolver<g2o::BlockSolverTraits<6, 3>> BlockSolverType;
  typedef g2o::LinearSolverDense<BlockSolverType::PoseMatrixType> LinearSolverType;
  g2o::SparseOptimizer optimizer;
  optimizer.setVerbose(false);
  optimizer.setAlgorithm(new g2o::OptimizationAlgorithmLevenberg(std::make_unique<BlockSolverType>(std::make_unique<LinearSolverType>())));
  optimizer.setComputeBatchStatistics(true);

  std::vector<g2o::VertexSE3Expmap*> poses(n_cam);
  for (long i = 0; i < n_cam; ++i) {
    const Eigen::Vector3d w(cam[6 * i], cam[6 * i + 1], cam[6 * i + 2]), c(cam[6 * i + 3], cam[6 * i + 4], cam[6 * i + 5]);
    const double th = w.norm();
    const Eigen::Matrix3d R_wc = th > 0.0 ? Eigen::AngleAxisd(th, w / th).toRotationMatrix() : Eigen::Matrix3d::Identity();
    const Eigen::Matrix3d R_cw = R_wc.transpose();
    auto* v = new g2o::VertexSE3Expmap();
    v->setEstimate(g2o::SE3Quat(R_cw, -R_cw * c));
    v->setId((int)i);
    v->setFixed(i == 0);
    optimizer.addVertex(v);
    poses[i] = v;
  }
  std::vector<g2o::VertexPointXYZ*> points(n_pt);
  for (long j = 0; j < n_pt; ++j) {
    auto* v = new g2o::VertexPointXYZ();
    v->setEstimate(Eigen::Vector3d(pt[3 * j], pt[3 * j + 1], pt[3 * j + 2]));
    v->setId((int)(n_cam + j));
    v->setMarginalized(true);
    optimizer.addVertex(v);
    points[j] = v;
  }
  auto* cam_params = new g2o::CameraParameters(K[0], Eigen::Vector2d(K[2], K[3]), 0.0);
  cam_params->setId(0);
  optimizer.addParameter(cam_params);
  for (long k = 0; k < n_obs; ++k) {
    auto* e = new g2o::EdgeProjectXYZ2UV();
    e->setVertex(0, points[op[k]]);
    e->setVertex(1, poses[oc[k]]);
    e->setMeasurement(Eigen::Vector2d(ou[k], ov[k]));
    e->setInformation(Eigen::Matrix2d::Identity());
    e->setParameterId(0, 0);
    if (huber > 0.0) { auto* rk = new g2o::RobustKernelHuber(); rk->setDelta(huber); e->setRobustKernel(rk); }
    optimizer.addEdge(e);
  }
  optimizer.initializeOptimization();
  optimizer.computeActiveErrors();
  const double chi2_initial = optimizer.activeRobustChi2();
  const int done = optimizer.optimize(iterations);

  std::printf("{\"huber\": %.17g, \"iterations_requested\": %d, \"iterations_done\": %d, \"chi2_initial\": %.17g, \"chi2_final\": %.17g,\n",
              huber, iterations, done, chi2_initial, optimizer.activeRobustChi2());
  std::printf(" \"iterations\": [\n");
  const g2o::BatchStatisticsContainer& stats = optimizer.batchStatistics();
  for (size_t i = 0; i < stats.size(); ++i)
    std::printf("  {\"iteration\": %d, \"chi2\": %.17g, \"levenberg_trials\": %d}%s\n", stats[i].iteration, stats[i].chi2,
                stats[i].levenbergIterations, i + 1 < stats.size() ? "," : "");
  std::printf(" ],\n \"cam\": [");               // world-to-camera [angle-axis(R_cw) | t], the GLBA_MODE_G2O convention
  for (long i = 0; i < n_cam; ++i) {
    const g2o::SE3Quat T = poses[i]->estimate();
    const Eigen::AngleAxisd aa(T.rotation());
    const Eigen::Vector3d w = aa.axis() * aa.angle(), t = T.translation();
    std::printf("%s%.17g, %.17g, %.17g, %.17g, %.17g, %.17g", i ? ", " : "", w[0], w[1], w[2], t[0], t[1], t[2]);
  }
  std::printf("],\n \"pt\": [");
  for (long j = 0; j < n_pt; ++j) {
    const Eigen::Vector3d X = points[j]->estimate();
    std::printf("%s%.17g, %.17g, %.17g", j ? ", " : "", X[0], X[1], X[2]);
  }
  std::printf("]}\n");
  return 0;
}
