// Exercises include/glba_slam.hpp the way GL-SLAM's threads would: build a Map, call full_ba / pose_only_ba.
//   adaptor_test pack  <scene.txt>            -> prints the packed problem (no GPU needed)
//   adaptor_test solve <scene.txt>            -> runs full_ba on the GPU, prints refined keyframes / points
//   adaptor_test pose  <pose.txt>             -> runs pose_only_ba, prints R, t
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iostream>
#include <string>

#include "glba_slam.hpp"

using namespace glslam;

static bool load_scene(const char* path, Map& map, CameraMatrix& K, int& window, int& run_window) {
  std::ifstream f(path);
  int n_cam, n_pt, n_obs, first;
  if (!(f >> n_cam >> n_pt >> n_obs >> K.fx >> K.fy >> K.cx >> K.cy >> window >> run_window >> first)) return false;
  for (int i = 0; i < n_cam; ++i) {
    Frame fr; fr.id = first + i; fr.is_keyframe = true;
    for (int q = 0; q < 9; ++q) f >> fr.R.m[q];
    for (int q = 0; q < 3; ++q) f >> fr.t.v[q];
    map.keyframes[fr.id] = fr;
  }
  for (int j = 0; j < n_pt; ++j) {
    MapPoint mp; mp.id = 1000 + j;
    int bad;
    f >> mp.position.x >> mp.position.y >> mp.position.z >> bad;
    mp.is_bad = bad != 0;
    map.map_points[mp.id] = mp;
  }
  for (int k = 0; k < n_obs; ++k) {
    int c, j; Observation o;
    f >> c >> j >> o.point2D.x >> o.point2D.y;
    o.keyframe_id = first + c; o.kp_index = k;
    map.map_points[1000 + j].obs.push_back(o);
    map.keyframes[o.keyframe_id].map_point_ids.push_back(1000 + j);   // one push per observation, like update_map_and_keyframe_data
  }
  map.next_keyframe_id = first + n_cam; map.next_point_id = 1000 + n_pt;
  return (bool)f;
}

int main(int argc, char** argv) {
  if (argc < 3) return 2;
  const std::string mode = argv[1];
  std::cout.precision(17);
  if (mode == "pose") {
    std::ifstream f(argv[2]);
    int n; CameraMatrix K; Mat33 R; Vec3 t;
    f >> n >> K.fx >> K.fy >> K.cx >> K.cy;
    for (int q = 0; q < 9; ++q) f >> R.m[q];
    for (int q = 0; q < 3; ++q) f >> t.v[q];
    std::vector<Point3d> p3(n); std::vector<Point2d> p2(n);
    for (int i = 0; i < n; ++i) f >> p3[i].x >> p3[i].y >> p3[i].z >> p2[i].x >> p2[i].y;
    Backend be(0);
    // size mismatch and empty input return false without touching R, t (slam_core.cpp:1096)
    std::vector<Point2d> shorter(p2.begin(), p2.end() - 1);
    Mat33 R0 = R;
    if (pose_only_ba(be, R, t, p3, shorter, K) || std::memcmp(&R0, &R, sizeof(R)) != 0) { std::cout << "FAIL mismatch\n"; return 1; }
    glba_summary s;
    const bool ok = pose_only_ba(be, R, t, p3, p2, K, nullptr, &s);
    std::cout << (ok ? "ok " : "false ") << s.n_iters << " " << s.final_cost << "\n";
    for (int q = 0; q < 9; ++q) std::cout << R.m[q] << " ";
    for (int q = 0; q < 3; ++q) std::cout << t.v[q] << " ";
    std::cout << "\n";
    return ok ? 0 : 1;
  }
  Map map; CameraMatrix K; int window = 0, run_window = 0;
  if (!load_scene(argv[2], map, K, window, run_window)) { std::cerr << "bad scene file\n"; return 2; }
  if (mode == "pack") {
    PackedWindow pw;
    // guards of slam_core.cpp:746-749
    PackedWindow tmp;
    if (pack_window(map, 1, run_window, tmp) || pack_window(map, (int)map.keyframes.size() + 1, run_window, tmp)) { std::cout << "FAIL guard\n"; return 1; }
    if (!pack_window(map, window, run_window, pw)) { std::cout << "FAIL pack\n"; return 1; }
    std::cout << pw.camera_params.size() / 6 << " " << pw.point_params.size() / 3 << " " << pw.obs_cam.size() << " " << pw.first_frame_idx << "\n";
    for (double v : pw.camera_params) std::cout << v << " ";
    std::cout << "\n";
    for (int id : pw.point_ids) std::cout << id << " ";
    std::cout << "\n";
    for (size_t k = 0; k < pw.obs_cam.size(); ++k) std::cout << pw.obs_cam[k] << " " << pw.obs_pt[k] << " " << pw.obs_u[k] << " " << pw.obs_v[k] << " ";
    std::cout << "\n";
    for (uint8_t v : pw.cam_fixed) std::cout << (int)v << " ";
    std::cout << "\n";
    // Rodrigues round trip on every keyframe
    double worst = 0;
    for (auto& kv : map.keyframes) { double w[3]; Mat33 R2; rodrigues(kv.second.R, w); rodrigues(w, R2);
      for (int q = 0; q < 9; ++q) worst = std::max(worst, std::fabs(R2.m[q] - kv.second.R.m[q])); }
    std::cout << worst << "\n";
    // post_ba_map_update_for_new_keyframes: move the last window keyframe by a known rigid motion (with a little
    // non-orthogonal drift on R, which ProjectToSO3 must remove) and check that a late keyframe / point follow it
    {
      Map m2 = map;
      Frame& last = m2.keyframes[run_window];
      const Mat33 R_before = last.R; const Vec3 t_before = last.t;
      const double w[3] = {0.01, -0.02, 0.015};
      Mat33 D; rodrigues(w, D);
      const Vec3 d{{0.3, -0.1, 0.2}};
      last.R = mul(D, R_before);
      for (int r = 0; r < 3; ++r) last.t.v[r] = D.m[r * 3] * t_before.v[0] + D.m[r * 3 + 1] * t_before.v[1] + D.m[r * 3 + 2] * t_before.v[2] + d.v[r];
      Mat33 noisy = R_before; noisy.m[1] += 1e-9; noisy.m[5] -= 2e-9;
      Frame extra; extra.id = 999; extra.R = R_before; extra.t = t_before; m2.keyframes[999] = extra;
      MapPoint mp; mp.id = 5000; mp.position = Point3d{1.0, 2.0, 3.0}; m2.map_points[5000] = mp;
      std::vector<int> mpids{5000}, kpids{999};
      post_ba_map_update_for_new_keyframes(m2, noisy, t_before, run_window, mpids, kpids);
      double err = 0;
      for (int q = 0; q < 9; ++q) err = std::max(err, std::fabs(m2.keyframes[999].R.m[q] - last.R.m[q]));
      for (int q = 0; q < 3; ++q) err = std::max(err, std::fabs(m2.keyframes[999].t.v[q] - last.t.v[q]));
      const Point3d& X = m2.map_points[5000].position;
      const double ex = D.m[0] * 1 + D.m[1] * 2 + D.m[2] * 3 + d.v[0], ey = D.m[3] * 1 + D.m[4] * 2 + D.m[5] * 3 + d.v[1], ez = D.m[6] * 1 + D.m[7] * 2 + D.m[8] * 3 + d.v[2];
      err = std::max(err, std::max(std::fabs(X.x - ex), std::max(std::fabs(X.y - ey), std::fabs(X.z - ez))));
      std::cout << err << " " << (mpids.empty() && kpids.empty() ? 1 : 0) << "\n";
    }
    return 0;
  }
  if (mode == "g2o") {
    // the archived entry point: world-to-camera poses of every keyframe in the file, all points, camera 0 fixed
    Backend be(0);
    if (!be.ok()) { std::cout << "nodevice " << be.status() << "\n"; return 3; }
    std::vector<int> kfids;
    for (const auto& kv : map.keyframes) kfids.push_back(kv.first);
    std::sort(kfids.begin(), kfids.end());
    std::unordered_map<int, int> cam_idx;
    std::vector<Mat33> Rs; std::vector<Vec3> Ts;
    for (int kfid : kfids) {
      cam_idx[kfid] = (int)Rs.size();
      const Frame& fr = map.keyframes[kfid];
      const Mat33 Rcw = transpose(fr.R);
      Vec3 t;
      for (int r = 0; r < 3; ++r) t.v[r] = -(Rcw.m[r * 3] * fr.t.v[0] + Rcw.m[r * 3 + 1] * fr.t.v[1] + Rcw.m[r * 3 + 2] * fr.t.v[2]);
      Rs.push_back(Rcw); Ts.push_back(t);
    }
    std::vector<Point3D> pts(map.map_points.size());
    for (int j = 0; j < (int)pts.size(); ++j) {
      const MapPoint& mp = map.map_points[1000 + j];
      pts[j].position = mp.position;
      for (const Observation& o : mp.obs) { Observation2D q; q.camera_idx = cam_idx[o.keyframe_id]; q.point2D = o.point2D; pts[j].observations.push_back(q); }
    }
    glba_summary s;
    const bool ok = bundleAdjustment(be, Rs, Ts, pts, K, 12, GLBA_LOSS_NONE, 1.0, &s);
    std::cout << (ok ? "ok " : "false ") << s.n_iters << " " << s.n_successful << " " << s.initial_cost << " " << s.final_cost << "\n";
    for (size_t i = 0; i < Rs.size(); ++i) {
      for (int q = 0; q < 9; ++q) std::cout << Rs[i].m[q] << " ";
      for (int q = 0; q < 3; ++q) std::cout << Ts[i].v[q] << " ";
      std::cout << "\n";
    }
    for (const Point3D& P : pts) std::cout << P.position.x << " " << P.position.y << " " << P.position.z << " ";
    std::cout << "\n";
    return ok ? 0 : 1;
  }
  if (mode == "solve" || mode == "resident") {
    Backend be(0);
    if (!be.ok()) { std::cout << "nodevice " << be.status() << "\n"; return 3; }
    std::mutex map_mutex, tracking_mutex;
    glba_summary s;
    bool ok = false;
    // what the tracking thread created while BA ran (slam_types::kpid_to_correct / mpid_to_correct): keyframe run_window+1 starts
    // at keyframe run_window's pose, one new point seen only by it
    const int n_orig = (int)map.map_points.size();
    const int kf_new = run_window + 1, mp_new = 1000 + n_orig;
    {
      Frame extra = map.keyframes[run_window]; extra.id = kf_new; extra.map_point_ids.clear(); extra.map_point_ids.push_back(mp_new);
      map.keyframes[kf_new] = extra;
      MapPoint mp; mp.id = mp_new; mp.position = Point3d{1.0, 2.0, 30.0};
      Observation o; o.keyframe_id = kf_new; o.point2D = Point2d{600.0, 180.0}; mp.obs.push_back(o);
      map.map_points[mp_new] = mp;
    }
    const Mat33 R16_before = map.keyframes[run_window].R; const Vec3 t16_before = map.keyframes[run_window].t;
    std::vector<int> mpids{mp_new}, kpids{kf_new};
    PostBa post; post.mpid_to_correct = &mpids; post.kpid_to_correct = &kpids; post.cull_map_points = true; post.local_ba_window = window;
    double mirror_err = 0.0;
    if (mode == "resident") {
      // grow the mirror the way the mapping thread would: keyframe by keyframe, a sync after each
      ResidentMap rm(be, K);
      Map grown;
      std::vector<int> kfids;
      for (const auto& kv : map.keyframes) kfids.push_back(kv.first);
      std::sort(kfids.begin(), kfids.end());
      for (int kfid : kfids) {
        grown.keyframes[kfid] = map.keyframes[kfid];
        for (int mpid : map.keyframes[kfid].map_point_ids) {
          MapPoint& g = grown.map_points[mpid];
          const MapPoint& src = map.map_points[mpid];
          g.id = src.id; g.position = src.position; g.is_bad = src.is_bad;
          for (const Observation& o : src.obs) if (o.keyframe_id == kfid) g.obs.push_back(o);
        }
        if (!rm.sync(grown)) { std::cout << "FAIL sync\n"; return 1; }
      }
      ok = full_ba_resident(be, rm, map_mutex, map, window, run_window, &tracking_mutex, nullptr, &s, &post);
      // the device mirror followed the host: keyframe kf_new and point mp_new after glba_map_propagate
      double cam[6], xyz[3];
      if (glba_map_read_keyframes(rm.handle(), kf_new - rm.first_keyframe_id(), 1, cam) != GLBA_OK ||
          glba_map_read_points(rm.handle(), rm.device_point(mp_new), 1, xyz, nullptr) != GLBA_OK) { std::cout << "FAIL mirror read\n"; return 1; }
      Mat33 Rm; rodrigues(cam, Rm);
      const Frame& h = map.keyframes[kf_new];
      for (int q = 0; q < 9; ++q) mirror_err = std::max(mirror_err, std::fabs(Rm.m[q] - h.R.m[q]));
      for (int q = 0; q < 3; ++q) mirror_err = std::max(mirror_err, std::fabs(cam[3 + q] - h.t.v[q]));
      const Point3d& X = map.map_points[mp_new].position;
      mirror_err = std::max(mirror_err, std::max(std::fabs(xyz[0] - X.x), std::max(std::fabs(xyz[1] - X.y), std::fabs(xyz[2] - X.z))));
      // host edits reach the mirror through the push wrappers
      map.keyframes[kf_new].t.v[0] += 0.5; map.map_points[mp_new].position.x += 0.25;
      if (!rm.push_keyframes(map, {kf_new}) || !rm.push_points(map, {mp_new})) { std::cout << "FAIL push\n"; return 1; }
      glba_map_read_keyframes(rm.handle(), kf_new - rm.first_keyframe_id(), 1, cam);
      glba_map_read_points(rm.handle(), rm.device_point(mp_new), 1, xyz, nullptr);
      mirror_err = std::max(mirror_err, std::max(std::fabs(cam[3] - map.keyframes[kf_new].t.v[0]), std::fabs(xyz[0] - map.map_points[mp_new].position.x)));
      map.keyframes[kf_new].t.v[0] -= 0.5; map.map_points[mp_new].position.x -= 0.25;
    } else {
      ok = full_ba(be, map_mutex, map, K, window, run_window, &tracking_mutex, nullptr, &s, &post);
    }
    std::cout << (ok ? "ok " : "false ") << s.n_iters << " " << s.initial_cost << " " << s.final_cost << "\n";
    const int first = run_window + 1 - window;
    for (int i = first; i < first + window; ++i) {
      const Frame& fr = map.keyframes[i];
      for (int q = 0; q < 9; ++q) std::cout << fr.R.m[q] << " ";
      for (int q = 0; q < 3; ++q) std::cout << fr.t.v[q] << " ";
      std::cout << "\n";
    }
    for (int j = 0; j < n_orig; ++j) { const MapPoint& mp = map.map_points[1000 + j]; std::cout << mp.position.x << " " << mp.position.y << " " << mp.position.z << " "; }
    std::cout << "\n";
    std::cout << post.n_culled << "\n";              // culling ran inside the critical section (PostBa)
    for (int j = 0; j < n_orig; ++j) std::cout << (map.map_points[1000 + j].is_bad ? 1 : 0) << " ";
    std::cout << "\n";
    // propagation (slam_core.cpp:916-973): the new keyframe started at keyframe run_window's pose, so it must end at its refined
    // pose; the new point moved by the same rigid delta; both lists consumed
    double perr = 0.0;
    const Frame& a = map.keyframes[run_window]; const Frame& b = map.keyframes[kf_new];
    for (int q = 0; q < 9; ++q) perr = std::max(perr, std::fabs(a.R.m[q] - b.R.m[q]));
    for (int q = 0; q < 3; ++q) perr = std::max(perr, std::fabs(a.t.v[q] - b.t.v[q]));
    Mat33 dR; Vec3 dt;
    compute_delta_pose_so3(R16_before, t16_before, a.R, a.t, dR, dt);
    const Point3d& X = map.map_points[mp_new].position;
    const double e0 = dR.m[0] * 1.0 + dR.m[1] * 2.0 + dR.m[2] * 30.0 + dt.v[0], e1 = dR.m[3] * 1.0 + dR.m[4] * 2.0 + dR.m[5] * 30.0 + dt.v[1],
                 e2 = dR.m[6] * 1.0 + dR.m[7] * 2.0 + dR.m[8] * 30.0 + dt.v[2];
    perr = std::max(perr, std::max(std::fabs(X.x - e0), std::max(std::fabs(X.y - e1), std::fabs(X.z - e2))));
    std::cout << perr << " " << (mpids.empty() && kpids.empty() ? 1 : 0) << " " << mirror_err << "\n";
    return ok ? 0 : 1;
  }
  return 2;
}
