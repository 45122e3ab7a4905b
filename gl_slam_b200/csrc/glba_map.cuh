// Persistent device-resident map.
//
// GL-SLAM re-walks its hash maps and re-allocates the parameter arrays on every local BA (slam_core.cpp:750-819).  Here
// the map lives in HBM as an append-only SoA mirror of slam_types.h:13-61 — keyframe poses [6], points [3] + is_bad,
// and an observation log (keyframe, point, u, v) in insertion order — fed by the three places the reference grows its
// map (update_map_and_keyframe_data, slam_core.cpp:287-426).  A local-BA window is then a device-side selection:
// observations whose keyframe lies in [first, first+window) and whose point is not bad, points compacted in id
// (= creation) order, handed to the same solver as glba_solve with device pointers, results scattered back into the
// map.  Nothing crosses PCIe but two counters and the summary.
//
// Included at the end of glba.cu (one translation unit: it uses load_problem / run_lm / write_back).
#pragma once

#include <cub/device/device_scan.cuh>

#include "../../include/glba_so3.hpp"

namespace {

__global__ void k_map_count(const long n, const int* __restrict__ o_kf, const int* __restrict__ o_pt, const uint8_t* __restrict__ bad,
                            const int first, const int window, int* __restrict__ cnt) {
  pdl_grid_sync();
  const long k = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const int f = o_kf[k] - first;
  if (f < 0 || f >= window) return;
  const int j = o_pt[k];
  if (!bad[j]) atomicAdd(cnt + j, 1);
}
__global__ void k_map_keep(const int n_pt, const int* __restrict__ cnt, const int min_obs, int* __restrict__ keep) {
  pdl_grid_sync();
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j <= n_pt) keep[j] = (j < n_pt && cnt[j] >= min_obs) ? 1 : 0;       // one past the end: the scan's total
}
__global__ void k_map_sel(const long n, const int* __restrict__ o_kf, const int* __restrict__ o_pt, const int* __restrict__ cnt,
                          const int min_obs, const int first, const int window, int* __restrict__ sel) {
  pdl_grid_sync();
  const long k = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (k > n) return;
  int v = 0;
  if (k < n) { const int f = o_kf[k] - first; v = (f >= 0 && f < window && cnt[o_pt[k]] >= min_obs) ? 1 : 0; }   // cnt > 0 implies not bad
  sel[k] = v;
}
__global__ void k_map_gather_obs(const long n, const int* __restrict__ o_kf, const int* __restrict__ o_pt, const double2* __restrict__ o_uv,
                                 const int* __restrict__ sel, const int* __restrict__ pos, const int* __restrict__ local, const int first,
                                 int* __restrict__ w_cam, int* __restrict__ w_pt, double* __restrict__ w_u, double* __restrict__ w_v) {
  pdl_grid_sync();
  const long k = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n || !sel[k]) return;
  const int d = pos[k];
  const double2 uv = o_uv[k];
  w_cam[d] = o_kf[k] - first; w_pt[d] = local[o_pt[k]]; w_u[d] = uv.x; w_v[d] = uv.y;
}
__global__ void k_map_gather_pt(const int n_pt, const int* __restrict__ keep, const int* __restrict__ local, const double* __restrict__ pt,
                                double* __restrict__ w_pt, int* __restrict__ w_id) {
  pdl_grid_sync();
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n_pt || !keep[j]) return;
  const int l = local[j];
  w_id[l] = j;
  w_pt[3 * (size_t)l] = pt[3 * (size_t)j]; w_pt[3 * (size_t)l + 1] = pt[3 * (size_t)j + 1]; w_pt[3 * (size_t)l + 2] = pt[3 * (size_t)j + 2];
}
__global__ void k_map_scatter_pt(const int n, const int* __restrict__ w_id, const double* __restrict__ w_pt, double* __restrict__ pt) {
  pdl_grid_sync();
  const int l = blockIdx.x * blockDim.x + threadIdx.x;
  if (l >= n) return;
  const size_t j = w_id[l];
  pt[3 * j] = w_pt[3 * (size_t)l]; pt[3 * j + 1] = w_pt[3 * (size_t)l + 1]; pt[3 * j + 2] = w_pt[3 * (size_t)l + 2];
}
// culling candidates: non-bad points whose earliest observation lies in keyframes [a, b]
__global__ void k_map_first_kf(const long n, const int* __restrict__ o_kf, const int* __restrict__ o_pt, int* __restrict__ first) {
  pdl_grid_sync();
  const long k = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (k < n) atomicMin(first + o_pt[k], o_kf[k]);
}
__global__ void k_map_keep_first(const int n_pt, const int* __restrict__ first, const uint8_t* __restrict__ bad, const int a, const int b,
                                 int* __restrict__ keep) {
  pdl_grid_sync();
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j <= n_pt) keep[j] = (j < n_pt && !bad[j] && first[j] >= a && first[j] <= b) ? 1 : 0;
}
__global__ void k_map_sel_kept(const long n, const int* __restrict__ o_pt, const int* __restrict__ keep, int* __restrict__ sel) {
  pdl_grid_sync();
  const long k = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (k <= n) sel[k] = (k < n && keep[o_pt[k]]) ? 1 : 0;
}
__global__ void k_map_set_u8(const int n, const int* __restrict__ ids, const uint8_t v, uint8_t* __restrict__ a) {
  pdl_grid_sync();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) a[ids[i]] = v;
}

// post_ba_map_update_for_new_keyframes (slam_core.cpp:916-973) on the resident map.
// k_map_delta: one thread computes the clean pose change of keyframe kf_last, ComputeDeltaPose_SO3 (:899-912), from the
// pose the caller saved before the write-back and the pose now in the map.  out[0..8] = dR (row-major), out[9..11] = dt.
__global__ void k_map_delta(const double* __restrict__ cam, const int kf_last, const double* __restrict__ before /* R[9] t[3] */,
                            double* __restrict__ out) {
  pdl_grid_sync();
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  const double* c = cam + 6 * (size_t)kf_last;
  double Ra[9];
  glba::so3_exp(c, Ra, nullptr);                   // keyframe rotation R_wc = exp([w]x), as cv::Rodrigues builds it
  glba_so3::compute_delta_pose_so3(before, before + 9, Ra, c + 3, out, out + 9);
}
// points X <- dR X + dt (:931-944); keyframes R <- dR R, t <- dR t + dt (:945-968), stored back as (angle-axis, centre)
__global__ void k_map_apply_delta(const double* __restrict__ d, const int n_kf, const int* __restrict__ kf_ids, double* __restrict__ cam,
                                  const int n_pt, const int* __restrict__ pt_ids, double* __restrict__ pt) {
  pdl_grid_sync();
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n_pt) {
    double* X = pt + 3 * (size_t)pt_ids[t];
    const double x = X[0], y = X[1], z = X[2];
    X[0] = d[0] * x + d[1] * y + d[2] * z + d[9];
    X[1] = d[3] * x + d[4] * y + d[5] * z + d[10];
    X[2] = d[6] * x + d[7] * y + d[8] * z + d[11];
  } else if (t < n_pt + n_kf) {
    double* c = cam + 6 * (size_t)kf_ids[t - n_pt];
    double R[9], Rn[9], w[3];
    glba::so3_exp(c, R, nullptr);
    glba_so3::mul3(d, R, Rn);
    glba::so3_log(Rn, w);
    const double x = c[3], y = c[4], z = c[5];
    c[0] = w[0]; c[1] = w[1]; c[2] = w[2];
    c[3] = d[0] * x + d[1] * y + d[2] * z + d[9];
    c[4] = d[3] * x + d[4] * y + d[5] * z + d[10];
    c[5] = d[6] * x + d[7] * y + d[8] * z + d[11];
  }
}

// append-only device array: capacity doubles, contents survive growth
int grow(glba_ctx* ctx, Buf& b, size_t used_bytes, size_t need_bytes) {
  if (b.cap >= need_bytes) return GLBA_OK;
  const size_t want = std::max(need_bytes + need_bytes / 2, (size_t)4096);
  void* q = nullptr;
  CU(cudaMalloc(&q, want));
  if (b.p && used_bytes) {
    if (cudaMemcpyAsync(q, b.p, used_bytes, cudaMemcpyDeviceToDevice, ctx->stream) != cudaSuccess || cudaStreamSynchronize(ctx->stream) != cudaSuccess) {
      cudaFree(q);
      return fail(ctx, GLBA_E_CUDA, "map: copy on growth failed");
    }
  }
  if (b.p) cudaFree(b.p);
  b.p = q; b.cap = want;
  return GLBA_OK;
}

}  // namespace

struct glba_map {
  glba_ctx* ctx = nullptr;
  double K[4] = {0, 0, 0, 0};
  int n_kf = 0, n_pt = 0;
  long n_obs = 0;
  Buf cam, pt, bad, o_kf, o_pt, o_uv;                                   // the map
  Buf cnt, keep, local, sel, pos, scan_tmp, stage;                      // window selection scratch
  Buf w_cam, w_pt, w_id, w_ocam, w_opt, w_u, w_v, w_fixed;              // the packed window problem
};

extern "C" {

int glba_map_create(glba_ctx* ctx, double fx, double fy, double cx, double cy, glba_map** out) {
  if (!ctx || !out || !(fx > 0.0) || !(fy > 0.0)) return fail(ctx, GLBA_E_INVALID_ARG, "map_create: bad argument");
  if (ctx->world > 1) return fail(ctx, GLBA_E_UNSUPPORTED, "map: local-BA windows stay on one GPU (SURVEY 8e)");
  glba_map* m = new (std::nothrow) glba_map();
  if (!m) return GLBA_E_OOM;
  m->ctx = ctx; m->K[0] = fx; m->K[1] = fy; m->K[2] = cx; m->K[3] = cy;
  *out = m;
  return GLBA_OK;
}

void glba_map_destroy(glba_map* m) {
  if (!m) return;
  cudaSetDevice(m->ctx->device);
  cudaStreamSynchronize(m->ctx->stream);
  for (Buf* b : {&m->cam, &m->pt, &m->bad, &m->o_kf, &m->o_pt, &m->o_uv, &m->cnt, &m->keep, &m->local, &m->sel, &m->pos, &m->scan_tmp, &m->stage,
                 &m->w_cam, &m->w_pt, &m->w_id, &m->w_ocam, &m->w_opt, &m->w_u, &m->w_v, &m->w_fixed})
    release(*b);
  delete m;
}

int glba_map_size(const glba_map* m, int32_t* n_kf, int32_t* n_pt, int64_t* n_obs) {
  if (!m) return GLBA_E_INVALID_ARG;
  if (n_kf) *n_kf = m->n_kf;
  if (n_pt) *n_pt = m->n_pt;
  if (n_obs) *n_obs = m->n_obs;
  return GLBA_OK;
}

int glba_map_add_keyframes(glba_map* m, int32_t n, const double* cam, int32_t* first_id) {
  if (!m) return GLBA_E_INVALID_ARG;
  glba_ctx* ctx = m->ctx;
  if (n < 0 || (n > 0 && !cam)) return fail(ctx, GLBA_E_INVALID_ARG, "map_add_keyframes: bad argument");
  CU(cudaSetDevice(ctx->device));
  if (first_id) *first_id = m->n_kf;
  if (n == 0) return GLBA_OK;
  int st = grow(ctx, m->cam, sizeof(double) * 6 * (size_t)m->n_kf, sizeof(double) * 6 * ((size_t)m->n_kf + n));
  if (st) return st;
  CU(cudaMemcpyAsync(m->cam.as<double>() + 6 * (size_t)m->n_kf, cam, sizeof(double) * 6 * (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  m->n_kf += n;
  return GLBA_OK;
}

int glba_map_add_points(glba_map* m, int32_t n, const double* xyz, int32_t* first_id) {
  if (!m) return GLBA_E_INVALID_ARG;
  glba_ctx* ctx = m->ctx;
  if (n < 0 || (n > 0 && !xyz)) return fail(ctx, GLBA_E_INVALID_ARG, "map_add_points: bad argument");
  CU(cudaSetDevice(ctx->device));
  if (first_id) *first_id = m->n_pt;
  if (n == 0) return GLBA_OK;
  int st = grow(ctx, m->pt, sizeof(double) * 3 * (size_t)m->n_pt, sizeof(double) * 3 * ((size_t)m->n_pt + n));
  if (st) return st;
  if ((st = grow(ctx, m->bad, (size_t)m->n_pt, (size_t)m->n_pt + n))) return st;
  CU(cudaMemcpyAsync(m->pt.as<double>() + 3 * (size_t)m->n_pt, xyz, sizeof(double) * 3 * (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
  CU(cudaMemsetAsync(m->bad.as<uint8_t>() + m->n_pt, 0, (size_t)n, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  m->n_pt += n;
  return GLBA_OK;
}

int glba_map_add_observations(glba_map* m, int32_t n, const int32_t* kf, const int32_t* pt, const double* uv) {
  if (!m) return GLBA_E_INVALID_ARG;
  glba_ctx* ctx = m->ctx;
  if (n < 0 || (n > 0 && (!kf || !pt || !uv))) return fail(ctx, GLBA_E_INVALID_ARG, "map_add_observations: bad argument");
  for (int32_t k = 0; k < n; ++k)
    if (kf[k] < 0 || kf[k] >= m->n_kf || pt[k] < 0 || pt[k] >= m->n_pt)
      return fail(ctx, GLBA_E_INVALID_ARG, "map_add_observations: observation %d refers to keyframe %d / point %d not in the map", k, kf[k], pt[k]);
  CU(cudaSetDevice(ctx->device));
  if (n == 0) return GLBA_OK;
  const size_t used = (size_t)m->n_obs, need = used + n;
  int st;
  if ((st = grow(ctx, m->o_kf, sizeof(int) * used, sizeof(int) * need))) return st;
  if ((st = grow(ctx, m->o_pt, sizeof(int) * used, sizeof(int) * need))) return st;
  if ((st = grow(ctx, m->o_uv, sizeof(double2) * used, sizeof(double2) * need))) return st;
  CU(cudaMemcpyAsync(m->o_kf.as<int>() + used, kf, sizeof(int) * (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
  CU(cudaMemcpyAsync(m->o_pt.as<int>() + used, pt, sizeof(int) * (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
  CU(cudaMemcpyAsync(m->o_uv.as<double2>() + used, uv, sizeof(double2) * (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  m->n_obs += n;
  return GLBA_OK;
}

int glba_map_set_bad(glba_map* m, int32_t n, const int32_t* pt_ids, uint8_t value) {
  if (!m) return GLBA_E_INVALID_ARG;
  glba_ctx* ctx = m->ctx;
  if (n < 0 || (n > 0 && !pt_ids)) return fail(ctx, GLBA_E_INVALID_ARG, "map_set_bad: bad argument");
  for (int32_t i = 0; i < n; ++i) if (pt_ids[i] < 0 || pt_ids[i] >= m->n_pt) return fail(ctx, GLBA_E_INVALID_ARG, "map_set_bad: point %d not in the map", pt_ids[i]);
  if (n == 0) return GLBA_OK;
  CU(cudaSetDevice(ctx->device));
  ENSURE(int, m->stage, n);
  CU(cudaMemcpyAsync(m->stage.p, pt_ids, sizeof(int) * (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
  LAUNCH(k_map_set_u8, cdiv(n, 256), 256, n, (const int*)m->stage.as<int>(), (uint8_t)(value ? 1 : 0), m->bad.as<uint8_t>());
  CU(cudaStreamSynchronize(ctx->stream));
  return GLBA_OK;
}

int glba_map_write_keyframes(glba_map* m, int32_t first, int32_t n, const double* cam) {
  if (!m) return GLBA_E_INVALID_ARG;
  glba_ctx* ctx = m->ctx;
  if (first < 0 || n < 0 || (long)first + n > m->n_kf || (n > 0 && !cam)) return fail(ctx, GLBA_E_INVALID_ARG, "map_write_keyframes: range outside the map");
  if (n == 0) return GLBA_OK;
  CU(cudaSetDevice(ctx->device));
  CU(cudaMemcpyAsync(m->cam.as<double>() + 6 * (size_t)first, cam, sizeof(double) * 6 * (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  return GLBA_OK;
}

int glba_map_read_keyframes(glba_map* m, int32_t first, int32_t n, double* cam) {
  if (!m) return GLBA_E_INVALID_ARG;
  glba_ctx* ctx = m->ctx;
  if (first < 0 || n < 0 || (long)first + n > m->n_kf || (n > 0 && !cam)) return fail(ctx, GLBA_E_INVALID_ARG, "map_read_keyframes: range outside the map");
  if (n == 0) return GLBA_OK;
  CU(cudaSetDevice(ctx->device));
  CU(cudaMemcpyAsync(cam, m->cam.as<double>() + 6 * (size_t)first, sizeof(double) * 6 * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  return GLBA_OK;
}

int glba_map_write_points(glba_map* m, int32_t first, int32_t n, const double* xyz) {
  if (!m) return GLBA_E_INVALID_ARG;
  glba_ctx* ctx = m->ctx;
  if (first < 0 || n < 0 || (long)first + n > m->n_pt || (n > 0 && !xyz)) return fail(ctx, GLBA_E_INVALID_ARG, "map_write_points: range outside the map");
  if (n == 0) return GLBA_OK;
  CU(cudaSetDevice(ctx->device));
  CU(cudaMemcpyAsync(m->pt.as<double>() + 3 * (size_t)first, xyz, sizeof(double) * 3 * (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  return GLBA_OK;
}

int glba_map_read_points(glba_map* m, int32_t first, int32_t n, double* xyz, uint8_t* bad) {
  if (!m) return GLBA_E_INVALID_ARG;
  glba_ctx* ctx = m->ctx;
  if (first < 0 || n < 0 || (long)first + n > m->n_pt || (n > 0 && !xyz)) return fail(ctx, GLBA_E_INVALID_ARG, "map_read_points: range outside the map");
  if (n == 0) return GLBA_OK;
  CU(cudaSetDevice(ctx->device));
  CU(cudaMemcpyAsync(xyz, m->pt.as<double>() + 3 * (size_t)first, sizeof(double) * 3 * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
  if (bad) CU(cudaMemcpyAsync(bad, m->bad.as<uint8_t>() + first, (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  return GLBA_OK;
}

// Local BA over keyframes [first_kf, first_kf + window): the first n_fixed of them constant (2 in slam_core.cpp:831-833),
// every non-bad point with >= min_obs observations inside the window free (the reference takes them all: min_obs = 1).
int glba_map_solve_window(glba_map* m, int32_t first_kf, int32_t window, int32_t n_fixed, int32_t min_obs, const glba_options* opt,
                          glba_summary* summary, int32_t* n_pt_used, int64_t* n_obs_used) {
  if (!m || !summary) return GLBA_E_INVALID_ARG;
  glba_ctx* ctx = m->ctx;
  std::memset(summary, 0, sizeof(*summary));
  if (first_kf < 0 || window <= 0 || (long)first_kf + window > m->n_kf || n_fixed < 0 || min_obs < 1)
    return summary->status = fail(ctx, GLBA_E_INVALID_ARG, "map_solve_window: window [%d, %d) outside the map (%d keyframes)", first_kf, first_kf + window, m->n_kf);
  CU(cudaSetDevice(ctx->device));
  int st = validate_options(ctx, opt);
  if (st) return summary->status = st;
  cudaStream_t s = ctx->stream;
  EventPair ev;
  cudaEventRecord(ev.a, s);
  const int n_pt = m->n_pt;
  const long n = m->n_obs;
  ENSURE(int, m->cnt, (size_t)n_pt + 1); ENSURE(int, m->keep, (size_t)n_pt + 1); ENSURE(int, m->local, (size_t)n_pt + 1);
  ENSURE(int, m->sel, (size_t)n + 1); ENSURE(int, m->pos, (size_t)n + 1);
  CU(cudaMemsetAsync(m->cnt.p, 0, sizeof(int) * ((size_t)n_pt + 1), s));
  if (n) LAUNCH(k_map_count, cdiv(n, 256), 256, n, (const int*)m->o_kf.as<int>(), (const int*)m->o_pt.as<int>(), (const uint8_t*)m->bad.as<uint8_t>(),
                first_kf, window, m->cnt.as<int>());
  LAUNCH(k_map_keep, cdiv(n_pt + 1, 256), 256, n_pt, (const int*)m->cnt.as<int>(), min_obs, m->keep.as<int>());
  LAUNCH(k_map_sel, cdiv(n + 1, 256), 256, n, (const int*)m->o_kf.as<int>(), (const int*)m->o_pt.as<int>(), (const int*)m->cnt.as<int>(), min_obs,
         first_kf, window, m->sel.as<int>());
  size_t tb1 = 0, tb2 = 0;
  CU(cub::DeviceScan::ExclusiveSum(nullptr, tb1, m->keep.as<int>(), m->local.as<int>(), n_pt + 1, s));
  CU(cub::DeviceScan::ExclusiveSum(nullptr, tb2, m->sel.as<int>(), m->pos.as<int>(), (int)(n + 1), s));
  ENSURE(char, m->scan_tmp, std::max(tb1, tb2));
  tb1 = tb2 = m->scan_tmp.cap;
  CU(cub::DeviceScan::ExclusiveSum(m->scan_tmp.p, tb1, m->keep.as<int>(), m->local.as<int>(), n_pt + 1, s));
  CU(cub::DeviceScan::ExclusiveSum(m->scan_tmp.p, tb2, m->sel.as<int>(), m->pos.as<int>(), (int)(n + 1), s));
  g_launches.fetch_add(2);
  CU(cudaMemcpyAsync(ctx->h_flags, m->local.as<int>() + n_pt, sizeof(int), cudaMemcpyDeviceToHost, s));
  CU(cudaMemcpyAsync(ctx->h_flags + 1, m->pos.as<int>() + n, sizeof(int), cudaMemcpyDeviceToHost, s));
  CU(cudaStreamSynchronize(s));
  const int wp = ctx->h_flags[0];
  const long wn = ctx->h_flags[1];
  if (n_pt_used) *n_pt_used = wp;
  if (n_obs_used) *n_obs_used = wn;
  ENSURE(double, m->w_cam, 6 * (size_t)window); ENSURE(double, m->w_pt, 3 * (size_t)wp); ENSURE(int, m->w_id, wp);
  ENSURE(int, m->w_ocam, wn); ENSURE(int, m->w_opt, wn); ENSURE(double, m->w_u, wn); ENSURE(double, m->w_v, wn); ENSURE(uint8_t, m->w_fixed, window);
  CU(cudaMemcpyAsync(m->w_cam.p, m->cam.as<double>() + 6 * (size_t)first_kf, sizeof(double) * 6 * (size_t)window, cudaMemcpyDeviceToDevice, s));
  CU(cudaMemsetAsync(m->w_fixed.p, 0, (size_t)window, s));
  if (n_fixed > 0) CU(cudaMemsetAsync(m->w_fixed.p, 1, (size_t)std::min(n_fixed, window), s));
  if (wp) LAUNCH(k_map_gather_pt, cdiv(n_pt, 256), 256, n_pt, (const int*)m->keep.as<int>(), (const int*)m->local.as<int>(), (const double*)m->pt.as<double>(),
                 m->w_pt.as<double>(), m->w_id.as<int>());
  if (wn) LAUNCH(k_map_gather_obs, cdiv(n, 256), 256, n, (const int*)m->o_kf.as<int>(), (const int*)m->o_pt.as<int>(), (const double2*)m->o_uv.as<double2>(),
                 (const int*)m->sel.as<int>(), (const int*)m->pos.as<int>(), (const int*)m->local.as<int>(), first_kf, m->w_ocam.as<int>(),
                 m->w_opt.as<int>(), m->w_u.as<double>(), m->w_v.as<double>());
  glba_problem p{};
  p.n_cam = window; p.n_pt = wp; p.n_obs = wn;
  p.cam = m->w_cam.as<double>(); p.pt = m->w_pt.as<double>();
  p.obs_cam = m->w_ocam.as<int>(); p.obs_pt = m->w_opt.as<int>(); p.obs_u = m->w_u.as<double>(); p.obs_v = m->w_v.as<double>();
  p.cam_fixed = m->w_fixed.as<uint8_t>(); p.pt_fixed = nullptr;
  p.fx = m->K[0]; p.fy = m->K[1]; p.cx = m->K[2]; p.cy = m->K[3];
  p.memspace = GLBA_MEM_DEVICE;
  cudaEventRecord(ev.b, s);
  ctx->t_phase[PH_SETUP] = 0.0;
  ctx->mode = opt->mode;
  if ((st = load_problem(ctx, &p))) return summary->status = st;
  st = run_lm(ctx, opt, summary);
  summary->status = st;
  if (st) return st;
  float pack_ms = 0.f;
  if (cudaEventElapsedTime(&pack_ms, ev.a, ev.b) == cudaSuccess) summary->t_setup_ms += pack_ms;
  if (summary->termination == GLBA_TERM_FAILURE) return GLBA_OK;           // the map keeps its values
  if ((st = write_back(ctx, p.cam, p.pt, GLBA_MEM_DEVICE))) return summary->status = st;
  CU(cudaMemcpyAsync(m->cam.as<double>() + 6 * (size_t)first_kf, m->w_cam.p, sizeof(double) * 6 * (size_t)window, cudaMemcpyDeviceToDevice, s));
  if (wp) LAUNCH(k_map_scatter_pt, cdiv(wp, 256), 256, wp, (const int*)m->w_id.as<int>(), (const double*)m->w_pt.as<double>(), m->pt.as<double>());
  CU(cudaStreamSynchronize(s));
  return GLBA_OK;
}

// post_ba_map_point_culling (slam_core.cpp:977-1038) on the resident map: candidates are the non-bad points first observed
// by keyframes [first_kf, last_kf] (the reference: [run_window - local_ba_window, run_window - 4], :981-992); a candidate
// is flagged bad when it lies behind one of its cameras, has fewer than min_obs observations or a mean reprojection
// error above max_mean_err over ALL its observations (:993-1035).  Flags are set in the map; the culled ids come back.
int glba_map_cull_points(glba_map* m, int32_t first_kf, int32_t last_kf, int32_t min_obs, double max_mean_err, int32_t* n_candidates,
                         int32_t* n_culled, int32_t* culled_ids, int32_t cap) {
  if (!m) return GLBA_E_INVALID_ARG;
  glba_ctx* ctx = m->ctx;
  if (n_candidates) *n_candidates = 0;
  if (n_culled) *n_culled = 0;
  if (cap < 0 || (cap > 0 && !culled_ids)) return fail(ctx, GLBA_E_INVALID_ARG, "map_cull_points: bad argument");
  first_kf = std::max(first_kf, 0); last_kf = std::min(last_kf, m->n_kf - 1);
  if (first_kf > last_kf || m->n_pt == 0 || m->n_obs == 0) return GLBA_OK;
  CU(cudaSetDevice(ctx->device));
  cudaStream_t s = ctx->stream;
  const int n_pt = m->n_pt;
  const long n = m->n_obs;
  ENSURE(int, m->cnt, (size_t)n_pt + 1); ENSURE(int, m->keep, (size_t)n_pt + 1); ENSURE(int, m->local, (size_t)n_pt + 1);
  ENSURE(int, m->sel, (size_t)n + 1); ENSURE(int, m->pos, (size_t)n + 1);
  CU(cudaMemsetAsync(m->cnt.p, 0x7f, sizeof(int) * ((size_t)n_pt + 1), s));            // "first keyframe" = +large
  LAUNCH(k_map_first_kf, cdiv(n, 256), 256, n, (const int*)m->o_kf.as<int>(), (const int*)m->o_pt.as<int>(), m->cnt.as<int>());
  LAUNCH(k_map_keep_first, cdiv(n_pt + 1, 256), 256, n_pt, (const int*)m->cnt.as<int>(), (const uint8_t*)m->bad.as<uint8_t>(), first_kf, last_kf,
         m->keep.as<int>());
  LAUNCH(k_map_sel_kept, cdiv(n + 1, 256), 256, n, (const int*)m->o_pt.as<int>(), (const int*)m->keep.as<int>(), m->sel.as<int>());
  size_t tb1 = 0, tb2 = 0;
  CU(cub::DeviceScan::ExclusiveSum(nullptr, tb1, m->keep.as<int>(), m->local.as<int>(), n_pt + 1, s));
  CU(cub::DeviceScan::ExclusiveSum(nullptr, tb2, m->sel.as<int>(), m->pos.as<int>(), (int)(n + 1), s));
  ENSURE(char, m->scan_tmp, std::max(tb1, tb2));
  tb1 = tb2 = m->scan_tmp.cap;
  CU(cub::DeviceScan::ExclusiveSum(m->scan_tmp.p, tb1, m->keep.as<int>(), m->local.as<int>(), n_pt + 1, s));
  CU(cub::DeviceScan::ExclusiveSum(m->scan_tmp.p, tb2, m->sel.as<int>(), m->pos.as<int>(), (int)(n + 1), s));
  g_launches.fetch_add(2);
  CU(cudaMemcpyAsync(ctx->h_flags, m->local.as<int>() + n_pt, sizeof(int), cudaMemcpyDeviceToHost, s));
  CU(cudaMemcpyAsync(ctx->h_flags + 1, m->pos.as<int>() + n, sizeof(int), cudaMemcpyDeviceToHost, s));
  CU(cudaStreamSynchronize(s));
  const int wp = ctx->h_flags[0];
  const long wn = ctx->h_flags[1];
  if (n_candidates) *n_candidates = wp;
  if (wp == 0) return GLBA_OK;
  ENSURE(double, m->w_pt, 3 * (size_t)wp); ENSURE(int, m->w_id, wp);
  ENSURE(int, m->w_ocam, wn); ENSURE(int, m->w_opt, wn); ENSURE(double, m->w_u, wn); ENSURE(double, m->w_v, wn);
  LAUNCH(k_map_gather_pt, cdiv(n_pt, 256), 256, n_pt, (const int*)m->keep.as<int>(), (const int*)m->local.as<int>(), (const double*)m->pt.as<double>(),
         m->w_pt.as<double>(), m->w_id.as<int>());
  LAUNCH(k_map_gather_obs, cdiv(n, 256), 256, n, (const int*)m->o_kf.as<int>(), (const int*)m->o_pt.as<int>(), (const double2*)m->o_uv.as<double2>(),
         (const int*)m->sel.as<int>(), (const int*)m->pos.as<int>(), (const int*)m->local.as<int>(), 0, m->w_ocam.as<int>(), m->w_opt.as<int>(),
         m->w_u.as<double>(), m->w_v.as<double>());
  glba_problem p{};
  p.n_cam = m->n_kf; p.n_pt = wp; p.n_obs = wn;            // every keyframe: a candidate's observations may lie anywhere
  p.cam = m->cam.as<double>(); p.pt = m->w_pt.as<double>();
  p.obs_cam = m->w_ocam.as<int>(); p.obs_pt = m->w_opt.as<int>(); p.obs_u = m->w_u.as<double>(); p.obs_v = m->w_v.as<double>();
  p.fx = m->K[0]; p.fy = m->K[1]; p.cx = m->K[2]; p.cy = m->K[3];
  p.memspace = GLBA_MEM_DEVICE;
  std::vector<uint8_t> flags((size_t)wp);
  std::vector<int> ids((size_t)wp);
  int st = glba_cull_points(ctx, &p, min_obs, max_mean_err, flags.data(), nullptr);
  if (st) return st;
  CU(cudaMemcpyAsync(ids.data(), m->w_id.p, sizeof(int) * (size_t)wp, cudaMemcpyDeviceToHost, s));
  CU(cudaStreamSynchronize(s));
  std::vector<int32_t> culled;
  for (int l = 0; l < wp; ++l) if (flags[l]) culled.push_back(ids[l]);
  if (n_culled) *n_culled = (int32_t)culled.size();
  for (size_t q = 0; q < culled.size() && (int32_t)q < cap; ++q) culled_ids[q] = culled[q];
  if (culled.empty()) return GLBA_OK;
  return glba_map_set_bad(m, (int32_t)culled.size(), culled.data(), 1);
}

// post_ba_map_update_for_new_keyframes (slam_core.cpp:916-973) on the resident map: the pose change of keyframe kf_last
// between (R_before, t_before) — what the caller saved before the BA write-back, :853-854 — and its pose now in the map is
// projected to SO(3) as ComputeDeltaPose_SO3 does (:885-912, including the reflection flip) and applied to the listed
// keyframes (kpid_to_correct) and points (mpid_to_correct).  dR_out[9] / dt_out[3] (may be NULL) receive the delta.
int glba_map_propagate(glba_map* m, const double* R_before, const double* t_before, int32_t kf_last, int32_t n_kf, const int32_t* kf_ids,
                       int32_t n_pt, const int32_t* pt_ids, double* dR_out, double* dt_out) {
  if (!m) return GLBA_E_INVALID_ARG;
  glba_ctx* ctx = m->ctx;
  if (!R_before || !t_before || kf_last < 0 || kf_last >= m->n_kf || n_kf < 0 || n_pt < 0 || (n_kf > 0 && !kf_ids) || (n_pt > 0 && !pt_ids))
    return fail(ctx, GLBA_E_INVALID_ARG, "map_propagate: bad argument");
  for (int32_t i = 0; i < n_kf; ++i) if (kf_ids[i] < 0 || kf_ids[i] >= m->n_kf) return fail(ctx, GLBA_E_INVALID_ARG, "map_propagate: keyframe %d not in the map", kf_ids[i]);
  for (int32_t i = 0; i < n_pt; ++i) if (pt_ids[i] < 0 || pt_ids[i] >= m->n_pt) return fail(ctx, GLBA_E_INVALID_ARG, "map_propagate: point %d not in the map", pt_ids[i]);
  CU(cudaSetDevice(ctx->device));
  cudaStream_t s = ctx->stream;
  // staging: [before 12 doubles | delta 12 doubles | kf ids | pt ids]
  const size_t bytes = 24 * sizeof(double) + sizeof(int) * ((size_t)n_kf + n_pt);
  ENSURE(char, m->stage, bytes);
  double* d_before = m->stage.as<double>();
  double* d_delta = d_before + 12;
  int* d_kf = reinterpret_cast<int*>(d_delta + 12);
  int* d_pt = d_kf + n_kf;
  double hb[12];
  for (int q = 0; q < 9; ++q) hb[q] = R_before[q];
  for (int q = 0; q < 3; ++q) hb[9 + q] = t_before[q];
  CU(cudaMemcpyAsync(d_before, hb, sizeof(hb), cudaMemcpyHostToDevice, s));
  if (n_kf) CU(cudaMemcpyAsync(d_kf, kf_ids, sizeof(int) * (size_t)n_kf, cudaMemcpyHostToDevice, s));
  if (n_pt) CU(cudaMemcpyAsync(d_pt, pt_ids, sizeof(int) * (size_t)n_pt, cudaMemcpyHostToDevice, s));
  LAUNCH(k_map_delta, 1, 32, (const double*)m->cam.as<double>(), kf_last, (const double*)d_before, d_delta);
  if (n_kf + n_pt > 0)
    LAUNCH(k_map_apply_delta, cdiv((long)n_kf + n_pt, 128), 128, (const double*)d_delta, n_kf, (const int*)d_kf, m->cam.as<double>(), n_pt, (const int*)d_pt,
           m->pt.as<double>());
  double hd[12];
  CU(cudaMemcpyAsync(hd, d_delta, sizeof(hd), cudaMemcpyDeviceToHost, s));
  CU(cudaStreamSynchronize(s));          // hb / the id arrays are the caller's: the copies have completed
  CU(cudaGetLastError());
  if (dR_out) for (int q = 0; q < 9; ++q) dR_out[q] = hd[q];
  if (dt_out) for (int q = 0; q < 3; ++q) dt_out[q] = hd[9 + q];
  return GLBA_OK;
}

}  // extern "C"
