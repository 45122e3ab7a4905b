// glba_slam.hpp — header-only C++17 host adaptor: GL-SLAM's BA entry points over the C ABI of glba.h.
//
// Mirrors, without OpenCV (absent from this image), the structures and the two functions the reference's
// tracking and mapping threads call:
//   struct Observation / MapPoint / Frame / Map         include/core/slam_types.h:13-61
//   bool slam_core::full_ba(map_mutex, map, K, window)  src/core/slam_core.cpp:744-883
//   bool slam_core::pose_only_ba(R, t, p3d, p2d, K)     src/core/slam_core.cpp:1092-1140
// Same names, argument meaning and error behaviour (false = "did nothing", inputs untouched); cv::Mat
// 3x3 / 3x1 CV_64F become Mat33 / Vec3, cv::Point2d/3d become Point2d/3d.  The hidden global inputs of the
// reference (slam_types::run_window, the two write-back mutexes, slam_core.cpp:758, 857-858) are explicit
// arguments here.  A maintainer swaps the ceres::Problem/ceres::Solve block for these calls: INTEGRATION.md.
#pragma once
#include <algorithm>
#include <cmath>
#include <mutex>
#include <unordered_map>
#include <unordered_set>
#include <vector>

#include "glba.h"
#include "glba_so3.hpp"

namespace glslam {

struct Point2d { double x = 0, y = 0; };
struct Point3d { double x = 0, y = 0, z = 0; };
struct Mat33 {                 // row-major, stands in for a 3x3 CV_64F cv::Mat
  double m[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  double& at(int r, int c) { return m[r * 3 + c]; }
  double at(int r, int c) const { return m[r * 3 + c]; }
};
struct Vec3 { double v[3] = {0, 0, 0}; };
struct CameraMatrix { double fx = 0, fy = 0, cx = 0, cy = 0; };   // K(0,0), K(1,1), K(0,2), K(1,2) — slam_core.cpp:720-723

// slam_types.h:13-19
struct Observation {
  int keyframe_id = 0;
  Point2d point2D;
  int kp_index = 0;
};
// slam_types.h:21-26
struct MapPoint {
  int id = 0;
  Point3d position;
  std::vector<Observation> obs;
  bool is_bad = false;
};
// slam_types.h:28-54 (image, descriptors and the covisibility list are front-end state the BA never reads)
struct Frame {
  int id = 0;
  Mat33 R;                        // camera-to-world rotation
  Vec3 t;                         // camera centre
  std::vector<int> map_point_ids; // one push per observation: may hold duplicates (slam_core.cpp:386-387, 405)
  std::vector<int> kp_to_mpid;
  bool is_keyframe = false;
};
// slam_types.h:56-61
struct Map {
  std::unordered_map<int, MapPoint> map_points;
  std::unordered_map<int, Frame> keyframes;
  int next_point_id = 0;
  int next_keyframe_id = 0;
};

// ---- cv::Rodrigues, both directions (slam_core.cpp:769, 862, 1101, 1136) ---------------------------------------
inline void rodrigues(const Mat33& R, double w[3]) {
  const double rx = 0.5 * (R.at(2, 1) - R.at(1, 2)), ry = 0.5 * (R.at(0, 2) - R.at(2, 0)), rz = 0.5 * (R.at(1, 0) - R.at(0, 1));
  const double s = std::sqrt(rx * rx + ry * ry + rz * rz);
  const double c = std::min(1.0, std::max(-1.0, 0.5 * (R.at(0, 0) + R.at(1, 1) + R.at(2, 2) - 1.0)));
  const double theta = std::atan2(s, c);
  if (s < 1e-8) {
    if (c > 0) { w[0] = rx; w[1] = ry; w[2] = rz; return; }          // theta ~ 0: w = skew part
    // theta ~ pi: axis from the symmetric part
    double ax = std::sqrt(std::max(0.0, 0.5 * (R.at(0, 0) + 1.0))), ay = std::sqrt(std::max(0.0, 0.5 * (R.at(1, 1) + 1.0))),
           az = std::sqrt(std::max(0.0, 0.5 * (R.at(2, 2) + 1.0)));
    if (R.at(0, 1) < 0) ay = -ay;
    if (R.at(0, 2) < 0) az = -az;
    if (std::fabs(ax) < std::fabs(ay) && std::fabs(ax) < std::fabs(az) && (R.at(1, 2) > 0) != (ay * az > 0)) az = -az;
    const double n = std::sqrt(ax * ax + ay * ay + az * az);
    w[0] = theta * ax / n; w[1] = theta * ay / n; w[2] = theta * az / n;
    return;
  }
  const double k = theta / s;
  w[0] = k * rx; w[1] = k * ry; w[2] = k * rz;
}
inline void rodrigues(const double w[3], Mat33& R) {
  const double th2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2];
  const double th = std::sqrt(th2);
  if (th < 1e-300) { R = Mat33(); return; }
  const double kx = w[0] / th, ky = w[1] / th, kz = w[2] / th;
  const double c = std::cos(th), s = std::sin(th), oc = 1.0 - c;
  R.m[0] = c + oc * kx * kx;      R.m[1] = oc * kx * ky - s * kz; R.m[2] = oc * kx * kz + s * ky;
  R.m[3] = oc * ky * kx + s * kz; R.m[4] = c + oc * ky * ky;      R.m[5] = oc * ky * kz - s * kx;
  R.m[6] = oc * kz * kx - s * ky; R.m[7] = oc * kz * ky + s * kx; R.m[8] = c + oc * kz * kz;
}

// ---- one glba context per calling thread (tracking thread / mapping thread) -------------------------------------
class Backend {
 public:
  explicit Backend(int device = 0) {
    glba_device_cfg cfg{};
    cfg.device = device; cfg.rank = 0; cfg.world = 1; cfg.nccl_unique_id = nullptr; cfg.stream = nullptr;
    status_ = glba_create(&cfg, &ctx_);
  }
  ~Backend() { if (ctx_) glba_destroy(ctx_); }
  Backend(const Backend&) = delete;
  Backend& operator=(const Backend&) = delete;
  bool ok() const { return ctx_ != nullptr; }
  int status() const { return status_; }
  glba_ctx* ctx() const { return ctx_; }
 private:
  glba_ctx* ctx_ = nullptr;
  int status_ = GLBA_OK;
};

// The flat problem full_ba hands to the solver (slam_core.cpp:750-819), kept for inspection / tests.
struct PackedWindow {
  std::vector<double> camera_params;              // 6 per keyframe: angle-axis(R), t
  std::vector<double> point_params;               // 3 per map point
  std::unordered_map<int, int> kf_to_param_idx;
  std::unordered_map<int, int> point_to_param_idx;
  std::vector<int> point_ids;                     // param idx -> map point id
  std::vector<int32_t> obs_cam, obs_pt;
  std::vector<double> obs_u, obs_v;
  std::vector<uint8_t> cam_fixed;
  int first_frame_idx = 0;
};

// Restates the packing of slam_core.cpp:750-838.  Returns false where the reference returns false (:746-749)
// or would throw std::out_of_range from .at() (:760, 783, 790).  Deliberate, documented differences:
//   * points are visited in increasing id (the reference iterates an unordered_set: arbitrary order, which only
//     changes floating-point summation order);
//   * the window test is kfid >= first+window (the reference's `>` lets an observation of keyframe first+window
//     through and silently binds it to camera 0 via operator[], slam_core.cpp:808-810).
inline bool pack_window(const Map& map, int window, int run_window, PackedWindow& out) {
  if ((int)map.keyframes.size() < window || window <= 1) return false;
  out = PackedWindow();
  const int first = run_window + 1 - window;
  out.first_frame_idx = first;
  for (int i = first; i < first + window; ++i) {
    auto it = map.keyframes.find(i);
    if (it == map.keyframes.end()) return false;
    out.kf_to_param_idx[i] = (int)out.camera_params.size() / 6;
    double w[3];
    rodrigues(it->second.R, w);
    out.camera_params.insert(out.camera_params.end(), {w[0], w[1], w[2], it->second.t.v[0], it->second.t.v[1], it->second.t.v[2]});
  }
  std::vector<int> ids;
  {
    std::unordered_set<int> seen;
    for (int i = first; i < first + window; ++i)
      for (int mpid : map.keyframes.at(i).map_point_ids)
        if (seen.insert(mpid).second) ids.push_back(mpid);
    std::sort(ids.begin(), ids.end());
  }
  for (int mpid : ids) {
    auto it = map.map_points.find(mpid);
    if (it == map.map_points.end()) return false;
    const MapPoint& mp = it->second;
    if (mp.is_bad || mp.obs.empty()) continue;
    const int pidx = (int)out.point_params.size() / 3;
    out.point_to_param_idx[mpid] = pidx;
    out.point_ids.push_back(mpid);
    out.point_params.insert(out.point_params.end(), {mp.position.x, mp.position.y, mp.position.z});
    for (const Observation& o : mp.obs) {
      if (o.keyframe_id < first || o.keyframe_id >= first + window) continue;
      out.obs_cam.push_back(out.kf_to_param_idx.at(o.keyframe_id));
      out.obs_pt.push_back(pidx);
      out.obs_u.push_back(o.point2D.x);
      out.obs_v.push_back(o.point2D.y);
    }
  }
  out.cam_fixed.assign(window, 0);
  out.cam_fixed[0] = 1;           // SetParameterBlockConstant(camera 0) and (camera 1), slam_core.cpp:831-833
  out.cam_fixed[1] = 1;
  return true;
}

// What the reference does INSIDE the critical section of the write-back, right after it (slam_core.cpp:853-879): propagate the
// pose change of keyframe run_window to the keyframes / points the tracking thread created while BA ran
// (post_ba_map_update_for_new_keyframes, :873) and cull old points (post_ba_map_point_culling, :875-879).  Passing a PostBa to
// full_ba / full_ba_resident runs both while tracking_mutex and map_mutex are still held, as the reference does: released in
// between, the tracking thread could read refined window poses next to uncorrected new keyframes, or push ids that then get
// the delta applied twice.
struct PostBa {
  std::vector<int>* mpid_to_correct = nullptr;   // slam_types::mpid_to_correct (consumed)
  std::vector<int>* kpid_to_correct = nullptr;   // slam_types::kpid_to_correct (consumed)
  bool cull_map_points = false;                  // slam_types::cull_map_points
  int local_ba_window = 0;                       // slam_types::local_ba_window: culling covers keyframes [run_window - it, run_window - 4]
  int n_culled = 0;                              // out: points flagged is_bad (-1: culling failed)
};
inline void post_ba_map_update_for_new_keyframes(Map& map, const Mat33& R_before, const Vec3& t_before, int run_window,
                                                 std::vector<int>& mpid_to_correct, std::vector<int>& kpid_to_correct);
inline int post_ba_map_point_culling(Backend& be, Map& map, const CameraMatrix& K, int run_window, int local_ba_window, double max_err = 1.0,
                                     int min_obs = 3);

// Drop-in for slam_core::full_ba.  `run_window` = slam_types::run_window; tracking_mutex may be null.
// On solver failure nothing is written back (the reference ignores ceres::Solver::Summary, slam_core.cpp:848-850;
// leaving the map untouched is the safe reading of "inputs untouched on failure").
inline bool full_ba(Backend& be, std::mutex& map_mutex, Map& map, const CameraMatrix& K, int window, int run_window,
                    std::mutex* tracking_mutex = nullptr, const glba_options* options = nullptr, glba_summary* summary = nullptr,
                    PostBa* post = nullptr) {
  if (!be.ok()) return false;
  PackedWindow pw;
  if (!pack_window(map, window, run_window, pw)) return false;
  glba_options opt;
  if (options) opt = *options; else glba_default_options(&opt);      // CauchyLoss(1.0), 30 iterations, Ceres LM defaults
  glba_problem p{};
  p.n_cam = window; p.n_pt = (int32_t)pw.point_ids.size(); p.n_obs = (int64_t)pw.obs_cam.size();
  p.cam = pw.camera_params.data(); p.pt = pw.point_params.data();
  p.obs_cam = pw.obs_cam.data(); p.obs_pt = pw.obs_pt.data(); p.obs_u = pw.obs_u.data(); p.obs_v = pw.obs_v.data();
  p.cam_fixed = pw.cam_fixed.data(); p.pt_fixed = nullptr;
  p.fx = K.fx; p.fy = K.fy; p.cx = K.cx; p.cy = K.cy; p.memspace = GLBA_MEM_HOST;
  glba_summary local;
  glba_summary* s = summary ? summary : &local;
  if (glba_solve(be.ctx(), &p, &opt, s) != GLBA_OK || s->termination == GLBA_TERM_FAILURE) return false;
  // write-back, slam_core.cpp:856-871 (tracking lock first, then the map lock)
  std::unique_lock<std::mutex> tl;
  if (tracking_mutex) tl = std::unique_lock<std::mutex>(*tracking_mutex);
  std::lock_guard<std::mutex> lk(map_mutex);
  const Mat33 R_before = map.keyframes[run_window].R;       // slam_core.cpp:853-854 (read under the locks here)
  const Vec3 t_before = map.keyframes[run_window].t;
  for (const auto& kv : pw.kf_to_param_idx) {
    const double* cam = &pw.camera_params[(size_t)kv.second * 6];
    Frame& kf = map.keyframes[kv.first];
    rodrigues(cam, kf.R);
    kf.t.v[0] = cam[3]; kf.t.v[1] = cam[4]; kf.t.v[2] = cam[5];
  }
  for (const auto& kv : pw.point_to_param_idx) {
    const double* pt = &pw.point_params[(size_t)kv.second * 3];
    map.map_points[kv.first].position = Point3d{pt[0], pt[1], pt[2]};
  }
  if (post) {                                               // still holding both locks, slam_core.cpp:873-879
    std::vector<int> none;
    post_ba_map_update_for_new_keyframes(map, R_before, t_before, run_window, post->mpid_to_correct ? *post->mpid_to_correct : none,
                                         post->kpid_to_correct ? *post->kpid_to_correct : none);
    if (post->cull_map_points) post->n_culled = post_ba_map_point_culling(be, map, K, run_window, post->local_ba_window, 1.0, 3);
  }
  return true;
}

// ---- persistent device-resident map (glba_map_*) ---------------------------------------------------------
// Mirrors a Map into HBM incrementally, so that full_ba_resident() below does not repeat the packing walk of
// slam_core.cpp:750-819 on every local BA.  Call sync() wherever the reference has just grown the map
// (update_map_and_keyframe_data, slam_core.cpp:287-426).  Keyframe ids must be consecutive (they are: GL-SLAM numbers
// keyframes with run_window); point ids are arbitrary.
class ResidentMap {
 public:
  ResidentMap(Backend& be, const CameraMatrix& K) : be_(be) { if (be.ok()) glba_map_create(be.ctx(), K.fx, K.fy, K.cx, K.cy, &map_); }
  ~ResidentMap() { if (map_) glba_map_destroy(map_); }
  ResidentMap(const ResidentMap&) = delete;
  ResidentMap& operator=(const ResidentMap&) = delete;
  bool ok() const { return map_ != nullptr; }
  glba_map* handle() const { return map_; }
  int first_keyframe_id() const { return kf_base_; }
  int device_point(int mpid) const { auto it = pt_index_.find(mpid); return it == pt_index_.end() ? -1 : it->second; }
  const std::vector<int>& point_ids() const { return pt_ids_; }        // device point index -> map point id

  // Appends the keyframes the device has not seen (ascending id), the points they observe that are new, every
  // observation of the touched points not pushed before, and the touched points' is_bad flags.
  bool sync(const Map& map) {
    if (!map_) return false;
    int kf_base = kf_base_, n_kf = n_kf_;
    if (n_kf == 0) {
      if (map.keyframes.empty()) return true;
      kf_base = map.keyframes.begin()->first;
      for (const auto& kv : map.keyframes) kf_base = std::min(kf_base, kv.first);
    }
    std::vector<int> touched;
    const int first_new = kf_base + n_kf;
    if (n_kf > 0) collect(map, first_new - 1, touched);     // a new keyframe also adds observations of the one before it (:386-405)
    std::vector<double> cams;
    int next = first_new;
    for (auto it = map.keyframes.find(next); it != map.keyframes.end(); it = map.keyframes.find(++next)) {
      double w[3];
      rodrigues(it->second.R, w);
      cams.insert(cams.end(), {w[0], w[1], w[2], it->second.t.v[0], it->second.t.v[1], it->second.t.v[2]});
      collect(map, next, touched);
    }
    n_kf += (int)(cams.size() / 6);
    std::sort(touched.begin(), touched.end());
    touched.erase(std::unique(touched.begin(), touched.end()), touched.end());
    // everything is STAGED first: the host-side index (pt_index_, pt_ids_, pushed_, bad_, n_kf_) is committed only after every
    // device append has succeeded, so a failure cannot leave it ahead of the device
    std::vector<double> xyz, uv;
    std::vector<int32_t> okf, opt, bad_ids, good_ids;
    std::vector<int> new_ids;                                  // map point ids appended by this call, in device order
    std::vector<std::pair<int, size_t>> new_pushed;            // (device point, observations pushed so far)
    std::vector<std::pair<int, uint8_t>> new_bad;
    const int n_known = (int)pt_ids_.size();
    std::unordered_map<int, int> staged_index;
    for (int mpid : touched) {
      auto mp_it = map.map_points.find(mpid);
      if (mp_it == map.map_points.end()) return false;
      const MapPoint& mp = mp_it->second;
      int d;
      auto known = pt_index_.find(mpid);
      if (known != pt_index_.end()) d = known->second;
      else {
        d = n_known + (int)new_ids.size();
        staged_index.emplace(mpid, d);
        new_ids.push_back(mpid);
        xyz.insert(xyz.end(), {mp.position.x, mp.position.y, mp.position.z});
      }
      const size_t done = d < n_known ? pushed_[d] : 0;
      for (size_t k = done; k < mp.obs.size(); ++k) {
        const Observation& o = mp.obs[k];
        if (o.keyframe_id < kf_base || o.keyframe_id >= kf_base + n_kf) return false;     // observation of an unknown keyframe
        okf.push_back(o.keyframe_id - kf_base); opt.push_back(d); uv.push_back(o.point2D.x); uv.push_back(o.point2D.y);
      }
      new_pushed.emplace_back(d, mp.obs.size());
      const uint8_t was = d < n_known ? bad_[d] : 0;
      if ((uint8_t)mp.is_bad != was) { (mp.is_bad ? bad_ids : good_ids).push_back(d); new_bad.emplace_back(d, mp.is_bad ? 1 : 0); }
    }
    if (!cams.empty() && glba_map_add_keyframes(map_, (int32_t)(cams.size() / 6), cams.data(), nullptr) != GLBA_OK) return false;
    kf_base_ = kf_base; n_kf_ = n_kf;                          // the device has them: commit step by step from here on
    if (!xyz.empty() && glba_map_add_points(map_, (int32_t)(xyz.size() / 3), xyz.data(), nullptr) != GLBA_OK) return false;
    for (int mpid : new_ids) { pt_index_.emplace(mpid, (int)pt_ids_.size()); pt_ids_.push_back(mpid); pushed_.push_back(0); bad_.push_back(0); }
    if (!okf.empty() && glba_map_add_observations(map_, (int32_t)okf.size(), okf.data(), opt.data(), uv.data()) != GLBA_OK) return false;
    for (const auto& pr : new_pushed) pushed_[pr.first] = pr.second;
    if (!bad_ids.empty() && glba_map_set_bad(map_, (int32_t)bad_ids.size(), bad_ids.data(), 1) != GLBA_OK) return false;
    if (!good_ids.empty() && glba_map_set_bad(map_, (int32_t)good_ids.size(), good_ids.data(), 0) != GLBA_OK) return false;
    for (const auto& pr : new_bad) bad_[pr.first] = pr.second;
    return true;
  }

  // post_ba_map_point_culling (slam_core.cpp:977-1038) entirely on the resident map: candidates first seen by keyframes
  // [run_window - local_ba_window, run_window - 4]; sets MapPoint::is_bad in `map` for the culled points.  Returns the
  // number of points culled, -1 on error.
  int cull(Map& map, int run_window, int local_ba_window, double max_err = 1.0, int min_obs = 3) {
    if (!map_) return -1;
    std::vector<int32_t> ids(pt_ids_.size());
    int32_t n_cand = 0, n_culled = 0;
    if (glba_map_cull_points(map_, run_window - local_ba_window - kf_base_, run_window - 4 - kf_base_, min_obs, max_err, &n_cand, &n_culled,
                             ids.data(), (int32_t)ids.size()) != GLBA_OK) return -1;
    for (int32_t q = 0; q < n_culled; ++q) { bad_[ids[q]] = 1; map.map_points[pt_ids_[ids[q]]].is_bad = true; }
    return n_culled;
  }

  // Host edits of poses / positions since the last BA (the tracking thread initialising a pose, a host-side propagation):
  // push them, or the next window starts from stale device values and a later write-back restores them.
  bool push_keyframes(const Map& map, const std::vector<int>& kfids) {
    for (int id : kfids) {
      auto it = map.keyframes.find(id);
      const int d = id - kf_base_;
      if (it == map.keyframes.end() || d < 0 || d >= n_kf_) return false;
      double cam[6];
      rodrigues(it->second.R, cam);
      cam[3] = it->second.t.v[0]; cam[4] = it->second.t.v[1]; cam[5] = it->second.t.v[2];
      if (glba_map_write_keyframes(map_, d, 1, cam) != GLBA_OK) return false;
    }
    return true;
  }
  bool push_points(const Map& map, const std::vector<int>& mpids) {
    for (int id : mpids) {
      auto it = map.map_points.find(id);
      const int d = device_point(id);
      if (it == map.map_points.end() || d < 0) return false;
      const double xyz[3] = {it->second.position.x, it->second.position.y, it->second.position.z};
      if (glba_map_write_points(map_, d, 1, xyz) != GLBA_OK) return false;
    }
    return true;
  }
  // post_ba_map_update_for_new_keyframes (slam_core.cpp:916-973) on the device mirror (glba_map_propagate): the listed
  // keyframes / points must have been sync()ed.  The host map is updated by glslam::post_ba_map_update_for_new_keyframes
  // with the same arithmetic (include/glba_so3.hpp); ids the device does not know yet are skipped (they arrive, already
  // corrected, with the next sync()).
  bool propagate(const Mat33& R_before, const Vec3& t_before, int run_window, const std::vector<int>& mpids, const std::vector<int>& kfids) {
    if (!map_) return false;
    std::vector<int32_t> dk, dp;
    for (int id : kfids) { const int d = id - kf_base_; if (d >= 0 && d < n_kf_) dk.push_back(d); }
    for (int id : mpids) { const int d = device_point(id); if (d >= 0) dp.push_back(d); }
    const int last = run_window - kf_base_;
    if (last < 0 || last >= n_kf_) return false;
    return glba_map_propagate(map_, R_before.m, t_before.v, last, (int32_t)dk.size(), dk.data(), (int32_t)dp.size(), dp.data(), nullptr, nullptr) == GLBA_OK;
  }

  // is_bad flags set by the host since the last sync (post_ba_map_point_culling, slam_core.cpp:977-1038)
  bool mark_bad(const std::vector<int>& mpids) {
    std::vector<int32_t> ids;
    for (int mpid : mpids) { const int d = device_point(mpid); if (d >= 0 && !bad_[d]) { ids.push_back(d); bad_[d] = 1; } }
    return ids.empty() || glba_map_set_bad(map_, (int32_t)ids.size(), ids.data(), 1) == GLBA_OK;
  }

 private:
  static void collect(const Map& map, int kfid, std::vector<int>& out) {
    auto it = map.keyframes.find(kfid);
    if (it != map.keyframes.end()) out.insert(out.end(), it->second.map_point_ids.begin(), it->second.map_point_ids.end());
  }
  Backend& be_;
  glba_map* map_ = nullptr;
  int kf_base_ = 0, n_kf_ = 0;
  std::unordered_map<int, int> pt_index_;
  std::vector<int> pt_ids_;
  std::vector<size_t> pushed_;
  std::vector<uint8_t> bad_;
};

// full_ba on the resident map: same window rule, fixed cameras, options and write-back as full_ba() above, but the
// window is selected and packed on the device; the host only copies the refined window back into `map`.
inline bool full_ba_resident(Backend& be, ResidentMap& rm, std::mutex& map_mutex, Map& map, int window, int run_window,
                             std::mutex* tracking_mutex = nullptr, const glba_options* options = nullptr, glba_summary* summary = nullptr,
                             PostBa* post = nullptr) {
  if (!be.ok() || !rm.ok()) return false;
  if ((int)map.keyframes.size() < window || window <= 1) return false;            // slam_core.cpp:746-749
  const int first = run_window + 1 - window;
  const int first_dev = first - rm.first_keyframe_id();
  glba_options opt;
  if (options) opt = *options; else glba_default_options(&opt);
  glba_summary local;
  glba_summary* s = summary ? summary : &local;
  if (glba_map_solve_window(rm.handle(), first_dev, window, 2, 1, &opt, s, nullptr, nullptr) != GLBA_OK || s->termination == GLBA_TERM_FAILURE)
    return false;
  std::vector<double> cams(6 * (size_t)window);
  if (glba_map_read_keyframes(rm.handle(), first_dev, window, cams.data()) != GLBA_OK) return false;
  int lo = (int)rm.point_ids().size(), hi = -1;          // device range covering the window's points (ids are creation-ordered)
  std::vector<uint8_t> in_window;                        // ... of which only the points the window observes are written back
  for (int i = first; i < first + window; ++i)
    for (int mpid : map.keyframes.at(i).map_point_ids) { const int d = rm.device_point(mpid); if (d >= 0) { lo = std::min(lo, d); hi = std::max(hi, d); } }
  if (hi >= lo) {
    in_window.assign((size_t)(hi - lo + 1), 0);
    for (int i = first; i < first + window; ++i)
      for (int mpid : map.keyframes.at(i).map_point_ids) { const int d = rm.device_point(mpid); if (d >= 0) in_window[d - lo] = 1; }
  }
  std::vector<double> xyz;
  std::vector<uint8_t> bad;
  if (hi >= lo) {
    xyz.resize(3 * (size_t)(hi - lo + 1)); bad.resize((size_t)(hi - lo + 1));
    if (glba_map_read_points(rm.handle(), lo, hi - lo + 1, xyz.data(), bad.data()) != GLBA_OK) return false;
  }
  std::unique_lock<std::mutex> tl;
  if (tracking_mutex) tl = std::unique_lock<std::mutex>(*tracking_mutex);
  std::lock_guard<std::mutex> lk(map_mutex);
  const Mat33 R_before = map.keyframes[run_window].R;       // slam_core.cpp:853-854
  const Vec3 t_before = map.keyframes[run_window].t;
  for (int i = 0; i < window; ++i) {
    Frame& kf = map.keyframes[first + i];
    rodrigues(&cams[6 * (size_t)i], kf.R);
    kf.t.v[0] = cams[6 * i + 3]; kf.t.v[1] = cams[6 * i + 4]; kf.t.v[2] = cams[6 * i + 5];
  }
  for (int d = lo; d <= hi; ++d) {
    if (bad[d - lo] || !in_window[d - lo]) continue;       // points outside the window keep their host values
    map.map_points[rm.point_ids()[d]].position = Point3d{xyz[3 * (size_t)(d - lo)], xyz[3 * (size_t)(d - lo) + 1], xyz[3 * (size_t)(d - lo) + 2]};
  }
  if (post) {                                               // still holding both locks, slam_core.cpp:873-879
    std::vector<int> none;
    std::vector<int>& mp = post->mpid_to_correct ? *post->mpid_to_correct : none;
    std::vector<int>& kp = post->kpid_to_correct ? *post->kpid_to_correct : none;
    // the device applies the delta to its mirror with the pose it holds (the refined one), the host to the host map
    rm.propagate(R_before, t_before, run_window, mp, kp);
    post_ba_map_update_for_new_keyframes(map, R_before, t_before, run_window, mp, kp);
    if (post->cull_map_points) post->n_culled = rm.cull(map, run_window, post->local_ba_window, 1.0, 3);
  }
  return true;
}

// ---- the archived g2o bundle adjustment (Old/mult_img_recoverpose_single_ba:251-326) --------------------------------
// Same signature and semantics: world-to-camera poses (Rs_est[i], Ts_est[i]) as recoverPose chains them, camera 0 fixed,
// every point free, unit information, no robust kernel, CameraParameters(fx, (cx, cy), 0) — one focal length —,
// OptimizationAlgorithmLevenberg for `iterations` iterations (100 there; 50 in docs/old_unorganized/4image_pnp_ba.txt:420,
// which also sets a Huber kernel: pass loss = GLBA_LOSS_HUBER, loss_scale = delta).  Runs GLBA_MODE_G2O.
struct Observation2D { int camera_idx = 0; Point2d point2D; };
struct Point3D { Point3d position; std::vector<Observation2D> observations; };

inline bool bundleAdjustment(Backend& be, std::vector<Mat33>& Rs_est, std::vector<Vec3>& Ts_est, std::vector<Point3D>& points3D,
                             const CameraMatrix& K, int iterations = 100, int loss = GLBA_LOSS_NONE, double loss_scale = 1.0,
                             glba_summary* summary = nullptr) {
  if (!be.ok() || Rs_est.size() != Ts_est.size() || Rs_est.empty()) return false;
  const int n_cam = (int)Rs_est.size(), n_pt = (int)points3D.size();
  std::vector<double> cams(6 * (size_t)n_cam), pts(3 * (size_t)n_pt), ou, ov;
  std::vector<int32_t> oc, op;
  for (int i = 0; i < n_cam; ++i) {
    rodrigues(Rs_est[i], &cams[6 * (size_t)i]);
    for (int r = 0; r < 3; ++r) cams[6 * (size_t)i + 3 + r] = Ts_est[i].v[r];
  }
  for (int j = 0; j < n_pt; ++j) {
    pts[3 * (size_t)j] = points3D[j].position.x; pts[3 * (size_t)j + 1] = points3D[j].position.y; pts[3 * (size_t)j + 2] = points3D[j].position.z;
    for (const Observation2D& o : points3D[j].observations) {
      if (o.camera_idx < 0 || o.camera_idx >= n_cam) return false;
      oc.push_back(o.camera_idx); op.push_back(j); ou.push_back(o.point2D.x); ov.push_back(o.point2D.y);
    }
  }
  std::vector<uint8_t> fixed(n_cam, 0);
  fixed[0] = 1;                                                      // "Fix first camera", :277
  glba_problem p{};
  p.n_cam = n_cam; p.n_pt = n_pt; p.n_obs = (int64_t)oc.size();
  p.cam = cams.data(); p.pt = pts.data(); p.obs_cam = oc.data(); p.obs_pt = op.data(); p.obs_u = ou.data(); p.obs_v = ov.data();
  p.cam_fixed = fixed.data(); p.pt_fixed = nullptr;
  p.fx = K.fx; p.fy = K.fx; p.cx = K.cx; p.cy = K.cy; p.memspace = GLBA_MEM_HOST;      // single focal length, :294-295
  glba_options opt;
  glba_default_options(&opt);
  opt.mode = GLBA_MODE_G2O; opt.max_iters = iterations; opt.loss = loss; opt.loss_scale = loss_scale;
  glba_summary local;
  glba_summary* s = summary ? summary : &local;
  if (glba_solve(be.ctx(), &p, &opt, s) != GLBA_OK || s->termination == GLBA_TERM_FAILURE) return false;
  for (int i = 0; i < n_cam; ++i) {
    rodrigues(&cams[6 * (size_t)i], Rs_est[i]);
    for (int r = 0; r < 3; ++r) Ts_est[i].v[r] = cams[6 * (size_t)i + 3 + r];
  }
  for (int j = 0; j < n_pt; ++j) points3D[j].position = Point3d{pts[3 * (size_t)j], pts[3 * (size_t)j + 1], pts[3 * (size_t)j + 2]};
  return true;
}

// Drop-in for slam_core::pose_only_ba (slam_core.cpp:1092-1140): R, t updated in place only on success.
inline bool pose_only_ba(Backend& be, Mat33& R, Vec3& t, const std::vector<Point3d>& p3d, const std::vector<Point2d>& p2d,
                         const CameraMatrix& K, const glba_options* options = nullptr, glba_summary* summary = nullptr) {
  if (p3d.size() != p2d.size() || p3d.empty()) return false;          // slam_core.cpp:1096
  if (!be.ok()) return false;
  double cam[6];
  rodrigues(R, cam);
  cam[3] = t.v[0]; cam[4] = t.v[1]; cam[5] = t.v[2];
  std::vector<double> X(3 * p3d.size()), uv(2 * p2d.size());
  for (size_t i = 0; i < p3d.size(); ++i) {
    X[3 * i] = p3d[i].x; X[3 * i + 1] = p3d[i].y; X[3 * i + 2] = p3d[i].z;
    uv[2 * i] = p2d[i].x; uv[2 * i + 1] = p2d[i].y;
  }
  glba_options opt;
  if (options) opt = *options; else glba_default_options(&opt);
  glba_summary local;
  glba_summary* s = summary ? summary : &local;
  if (glba_pose_only(be.ctx(), cam, (int32_t)p3d.size(), X.data(), uv.data(), K.fx, K.fy, K.cx, K.cy, &opt, s) != GLBA_OK) return false;
  if (s->termination == GLBA_TERM_FAILURE) return false;               // !summary.IsSolutionUsable(), slam_core.cpp:1132
  rodrigues(cam, R);
  t.v[0] = cam[3]; t.v[1] = cam[4]; t.v[2] = cam[5];
  return true;
}

// slam_types.h:64-69
struct Match2D2D {
  int idx0 = 0, idx1 = 0;
  Point2d p0, p1;
};

// Drop-in for slam_core::triangulate_and_filter_3d_points (slam_core.cpp:173-256): [R|t] are the world-to-camera
// poses the tracking thread passes; returns the kept 3-D points and their matches in input order.  The DLT
// (cv::triangulatePoints, :194) and every filter (:213-241) run on the GPU, one thread per match.
inline bool triangulate_and_filter_3d_points(Backend& be, const Mat33& R1, const Vec3& t1, const Mat33& R2, const Vec3& t2,
                                             const CameraMatrix& K, const std::vector<Match2D2D>& matches, float distance_threshold,
                                             float reprojection_threshold, std::vector<Point3d>& points3d,
                                             std::vector<Match2D2D>& filteredPairs) {
  points3d.clear(); filteredPairs.clear();
  if (!be.ok()) return false;
  const int32_t n = (int32_t)matches.size();
  if (n == 0) return true;
  std::vector<double> p0(2 * (size_t)n), p1(2 * (size_t)n), X(3 * (size_t)n);
  std::vector<uint8_t> keep(n);
  for (int32_t i = 0; i < n; ++i) { p0[2 * i] = matches[i].p0.x; p0[2 * i + 1] = matches[i].p0.y; p1[2 * i] = matches[i].p1.x; p1[2 * i + 1] = matches[i].p1.y; }
  if (glba_triangulate_filter(be.ctx(), R1.m, t1.v, R2.m, t2.v, K.fx, K.fy, K.cx, K.cy, n, p0.data(), p1.data(), distance_threshold,
                              reprojection_threshold, X.data(), keep.data()) != GLBA_OK) return false;
  for (int32_t i = 0; i < n; ++i)
    if (keep[i]) { points3d.push_back(Point3d{X[3 * i], X[3 * i + 1], X[3 * i + 2]}); filteredPairs.push_back(matches[i]); }
  return true;
}

// ---- post-BA propagation to keyframes / points created while BA ran (slam_core.cpp:885-973) ----------------------------
inline Mat33 mul(const Mat33& A, const Mat33& B) {
  Mat33 C;
  for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) C.m[r * 3 + c] = A.m[r * 3] * B.m[c] + A.m[r * 3 + 1] * B.m[3 + c] + A.m[r * 3 + 2] * B.m[6 + c];
  return C;
}
inline Mat33 transpose(const Mat33& A) { Mat33 T; for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) T.m[r * 3 + c] = A.m[c * 3 + r]; return T; }
inline double det(const Mat33& A) {
  return A.m[0] * (A.m[4] * A.m[8] - A.m[5] * A.m[7]) - A.m[1] * (A.m[3] * A.m[8] - A.m[5] * A.m[6]) + A.m[2] * (A.m[3] * A.m[7] - A.m[4] * A.m[6]);
}
// ProjectToSO3 (slam_core.cpp:885-897): U V' of the SVD, last column of U flipped when det(U V') < 0.  The arithmetic is
// include/glba_so3.hpp, shared with the device kernel behind glba_map_propagate.
inline Mat33 project_to_so3(const Mat33& R_in) {
  Mat33 R;
  glba_so3::project_to_so3(R_in.m, R.m);
  return R;
}
// ComputeDeltaPose_SO3 (slam_core.cpp:899-912)
inline void compute_delta_pose_so3(const Mat33& Rb_in, const Vec3& tb, const Mat33& Ra_in, const Vec3& ta, Mat33& dR, Vec3& dt) {
  glba_so3::compute_delta_pose_so3(Rb_in.m, tb.v, Ra_in.m, ta.v, dR.m, dt.v);
}
// post_ba_map_update_for_new_keyframes (slam_core.cpp:916-973): the last window camera's pose change is applied to the
// points / keyframes the tracking thread created while BA was running (slam_types::mpid_to_correct / kpid_to_correct,
// both consumed).  R_before / t_before = pose of keyframe run_window before the write-back.
inline void post_ba_map_update_for_new_keyframes(Map& map, const Mat33& R_before, const Vec3& t_before, int run_window,
                                                 std::vector<int>& mpid_to_correct, std::vector<int>& kpid_to_correct) {
  const Frame& last = map.keyframes[run_window];
  Mat33 dR; Vec3 dt;
  compute_delta_pose_so3(R_before, t_before, last.R, last.t, dR, dt);
  while (!mpid_to_correct.empty()) {
    Point3d& X = map.map_points[mpid_to_correct.back()].position;
    const double x = X.x, y = X.y, z = X.z;
    X.x = dR.m[0] * x + dR.m[1] * y + dR.m[2] * z + dt.v[0];
    X.y = dR.m[3] * x + dR.m[4] * y + dR.m[5] * z + dt.v[1];
    X.z = dR.m[6] * x + dR.m[7] * y + dR.m[8] * z + dt.v[2];
    mpid_to_correct.pop_back();
  }
  while (!kpid_to_correct.empty()) {
    Frame& kf = map.keyframes[kpid_to_correct.back()];
    const Vec3 t = kf.t;
    kf.R = mul(dR, kf.R);
    for (int r = 0; r < 3; ++r) kf.t.v[r] = dR.m[r * 3] * t.v[0] + dR.m[r * 3 + 1] * t.v[1] + dR.m[r * 3 + 2] * t.v[2] + dt.v[r];
    kpid_to_correct.pop_back();
  }
}

// post_ba_map_point_culling (slam_core.cpp:977-1038): candidates are the points first seen by keyframes
// [run_window - local_ba_window, run_window - 4]; a point is flagged is_bad when it lies behind one of its
// cameras, has fewer than `min_obs` observations or a mean reprojection error above `max_err` pixels.  The
// per-point arithmetic over ALL of the point's observations runs on the GPU (glba_cull_points).
inline int post_ba_map_point_culling(Backend& be, Map& map, const CameraMatrix& K, int run_window, int local_ba_window,
                                     double max_err, int min_obs) {
  if (!be.ok()) return -1;
  std::vector<int> ids;
  {
    std::unordered_set<int> seen;
    for (int i = run_window - local_ba_window; i <= run_window - 4; ++i) {
      if (i < 0) continue;
      auto kf = map.keyframes.find(i);
      if (kf == map.keyframes.end()) continue;
      for (int mp : kf->second.map_point_ids) {
        auto it = map.map_points.find(mp);
        if (it == map.map_points.end() || it->second.obs.empty() || it->second.is_bad) continue;
        if (it->second.obs.front().keyframe_id == i && seen.insert(mp).second) ids.push_back(mp);
      }
    }
    std::sort(ids.begin(), ids.end());
  }
  if (ids.empty()) return 0;
  std::unordered_map<int, int> cam_idx;
  std::vector<double> cams, pts;
  std::vector<int32_t> oc, op;
  std::vector<double> ou, ov;
  for (size_t j = 0; j < ids.size(); ++j) {
    const MapPoint& mp = map.map_points.at(ids[j]);
    pts.insert(pts.end(), {mp.position.x, mp.position.y, mp.position.z});
    for (const Observation& o : mp.obs) {
      auto kf = map.keyframes.find(o.keyframe_id);
      if (kf == map.keyframes.end()) return -1;                       // .at() would throw in the reference (:1004)
      auto ins = cam_idx.emplace(o.keyframe_id, (int)cam_idx.size());
      if (ins.second) {
        double w[3];
        rodrigues(kf->second.R, w);
        cams.insert(cams.end(), {w[0], w[1], w[2], kf->second.t.v[0], kf->second.t.v[1], kf->second.t.v[2]});
      }
      oc.push_back(ins.first->second); op.push_back((int32_t)j); ou.push_back(o.point2D.x); ov.push_back(o.point2D.y);
    }
  }
  glba_problem p{};
  p.n_cam = (int32_t)cam_idx.size(); p.n_pt = (int32_t)ids.size(); p.n_obs = (int64_t)oc.size();
  p.cam = cams.data(); p.pt = pts.data(); p.obs_cam = oc.data(); p.obs_pt = op.data(); p.obs_u = ou.data(); p.obs_v = ov.data();
  p.fx = K.fx; p.fy = K.fy; p.cx = K.cx; p.cy = K.cy; p.memspace = GLBA_MEM_HOST;
  std::vector<uint8_t> bad(ids.size());
  if (glba_cull_points(be.ctx(), &p, min_obs, max_err, bad.data(), nullptr) != GLBA_OK) return -1;
  int culled = 0;
  for (size_t j = 0; j < ids.size(); ++j)
    if (bad[j]) { map.map_points[ids[j]].is_bad = true; ++culled; }
  return culled;
}

}  // namespace glslam
