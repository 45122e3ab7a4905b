// glba.cu — host side of the C ABI in include/glba.h: problem upload + index construction,
// the Levenberg-Marquardt driver (Ceres trust-region semantics), PCG on the implicit
// Schur complement, NCCL plumbing for sharded maps.  All arithmetic runs in the kernels of
// glba_kernels.cuh / glba_pose.cuh; there is no CPU fallback: without a CUDA device every entry
// point returns GLBA_E_NO_DEVICE.
//
// Reference boundary: slam_core::full_ba (src/core/slam_core.cpp:744-883) and
// slam_core::pose_only_ba (:1092-1140) of GL-SLAM.
#include "../../include/glba.h"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <dlfcn.h>
#include <string>
#include <vector>

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_run_length_encode.cuh>
#include <cub/device/device_scan.cuh>
#include <cuda_runtime.h>

#include "glba_kernels.cuh"
#include "glba_pose.cuh"
#include "glba_dense.cuh"
#include "glba_tiles.cuh"
#include "glba_pipe.cuh"
#include "glba_campipe.cuh"
#include "glba_cam.cuh"
#include "glba_sparse.cuh"
#include "glba_triang.cuh"

using namespace glba;

// ---- minimal NCCL surface, resolved with dlopen so single-GPU use needs no NCCL at all ----------
extern "C" {
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
}
namespace {
constexpr int kNcclFloat64 = 8, kNcclInt32 = 2, kNcclInt64 = 4, kNcclSum = 0, kNcclMax = 2;   // ncclDataType_t / ncclRedOp_t values
struct NcclApi {
  void* h = nullptr;
  int (*GetUniqueId)(ncclUniqueId*) = nullptr;
  int (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*CommDestroy)(ncclComm_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  bool load() {
    if (h) return true;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) { h = dlopen(n, RTLD_NOW | RTLD_GLOBAL); if (h) break; }
    if (!h) return false;
    GetUniqueId = (int (*)(ncclUniqueId*))dlsym(h, "ncclGetUniqueId");
    CommInitRank = (int (*)(ncclComm_t*, int, ncclUniqueId, int))dlsym(h, "ncclCommInitRank");
    AllReduce = (int (*)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t))dlsym(h, "ncclAllReduce");
    AllGather = (int (*)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t))dlsym(h, "ncclAllGather");
    CommDestroy = (int (*)(ncclComm_t))dlsym(h, "ncclCommDestroy");
    GetErrorString = (const char* (*)(int))dlsym(h, "ncclGetErrorString");
    return GetUniqueId && CommInitRank && AllReduce && CommDestroy;
  }
};
NcclApi g_nccl;
std::atomic<long long> g_launches{0};

struct Buf {
  void* p = nullptr;
  size_t cap = 0;
  template <typename T> T* as() const { return reinterpret_cast<T*>(p); }
};

enum Phase { PH_SETUP = 0, PH_LIN, PH_SCHUR, PH_SOLVE, PH_UPDATE, PH_COMM, PH_COUNT };

struct EventPair {     // two CUDA events, destroyed on every exit path
  cudaEvent_t a = nullptr, b = nullptr;
  EventPair() { cudaEventCreate(&a); cudaEventCreate(&b); }
  ~EventPair() { if (a) cudaEventDestroy(a); if (b) cudaEventDestroy(b); }
  EventPair(const EventPair&) = delete;
  EventPair& operator=(const EventPair&) = delete;
};
}  // namespace

struct glba_ctx {
  int device = 0, rank = 0, world = 1;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  cudaStream_t side_stream = nullptr;          // camera finalisation of a linearisation, concurrent with the Schur pass (single GPU)
  cudaEvent_t ev_side[2] = {nullptr, nullptr};
  cudaStream_t copy_stream = nullptr;          // host uploads of load_problem, overlapped with the index construction
  cudaEvent_t ev_copy[4] = {nullptr, nullptr, nullptr, nullptr};
  bool copies_in_flight = false;
  ncclComm_t comm = nullptr;
  // peer exchange (k_xch_pack_p2p / k_xch_unpack_p2p): a buffer every rank of the node maps through CUDA IPC
  bool env_p2p = false;                        // GLBA_P2P=1: the compact exchange over peer memory instead of ncclAllReduce (measured on par at 8 GPUs, slower at 2)
  bool p2p_ok = false;
  void* p2p_local = nullptr;                   // [slot 0 | slot 1 | flags | error]
  void* p2p_peer[16] = {nullptr};              // the same allocation of every rank, in this process' address space
  size_t p2p_cap = 0;                          // doubles per slot
  unsigned p2p_seq = 0;                        // exchanges issued so far (same on every rank)
  std::string err;
  // problem
  bool loaded = false;
  int n_cam = 0, n_pt = 0, n_chunks = 0, n_free_cam = 0;
  long n_obs = 0;
  Intr K{};
  bool sorted_input = true;
  // device buffers
  Buf in_cam, in_pt, in_ocam, in_opt, in_u, in_v, in_cfix, in_pfix, in_info;               // staging of host input
  Buf pm_cam, pm_pt, pm_uv, pm2orig, pm2cm, pt_start, cm_pt, cm_uv, cm2pm, cam_start; // index
  Buf chunk_cam, chunk_begin, chunk_end, cam_chunk_start, cam_free, pt_free, sort_tmp, keys_tmp, flags;
  Buf cam[2], camtab[2], pt4[2], cam0, pt40;                                         // state (double-buffered)
  Buf rec_pm, rec_cm, Craw, sp4, lam4, cinv, u0p, u4;
  Buf part_pm, part_cm, acc27, yhat, Bc, gc, sc, lamc, Md, Minv, rhs, cg_x, cg_r, cg_p, cg_q, pg, yg, cgst;
  double *d_accA = nullptr, *d_accB = nullptr;   // per-camera partial sums of the linearise / Schur passes
  bool schur_fresh = false;                      // Schur pieces already built for the current linearisation and radius
  double* d_scal = nullptr;   // NSCAL doubles at the tail of acc27 (one all-reduce carries camera sums + scalars)
  Buf out_a, out_b, out_c;                                                           // glba_linearize outputs
  Buf first_cam, new2old, old2new, opt_relab;                                         // locality relabelling of the points
  bool relabelled = false, allow_relabel = true, env_relabel = true, env_timing = false;
  int mode = GLBA_MODE_CERES;                 // formulation the loaded problem's poses are in (glba_mode)
  Buf hmax;                                   // GLBA_MODE_G2O: max Hessian diagonal (bit pattern of a double)
  Buf tile_cmin, tile_pt, xtab, partA, partB, partc, counters, cam_cnt, part_cm2, part_pm2;                                  // tiles, PCG gather table, camera-kernel partials
  bool use_tiles = false;
  // sharded runs, "owner-computes" layout (see load_problem): this rank's problem holds only the cameras its own tracks observe
  bool owner = false, want_owner = false;
  int n_cam_g = 0, n_shared = 0, n_free_cam_g = 0;            // cameras of the whole map; cameras observed by more than one rank
  Buf act_mask, g2l, l2g, cam_owned, cam_shared, sh_scan, xsend, xrecv, xsend6, ocam_loc, cam_loc, cfix_loc, late, sh2loc;
  Buf lmctl, dsum;              // device-resident LM control (LmCtl) and summary trace (glba_summary) of the on-device loop
  bool env_host_lm = false;     // diagnostic: GLBA_HOST_LM=1 keeps the decisions on the host for small windows too
  bool use_pipe = false;        // large maps: persistent TMA-fed tile kernels (glba_pipe.cuh)
  bool env_pipe = true, env_force_large = false;
  int env_cp_occ = 2;           // diagnostic: GLBA_CP_OCC=3 = the 85-register build of k_cam_pipe (3 CTAs/SM)
  bool env_campipe = true;      // diagnostic: GLBA_CAMPIPE=0 runs the two camera-major passes as separate kernels on large maps too
  Buf tile_desc, tile_cams, pm_slot;
  // fused product (k_pt_pipe<2>): slot lists per tile, camera -> (tile, slot) list, per-tile sums, identity CSR
  bool use_fused = false, env_fused = false;      // measured slower than the two-kernel product (see k_pt_pipe<2>): opt-in with GLBA_FUSED=1
  Buf tile_sobs, tile_sstart, tp_key, tp_val, tp_key2, cam_tp, cam_tp_start, tpart, cam_iota;
  Buf ovf_raw, ovf_k, ovf_c, ovf_key, ovf_key2, ovf_val, cam_ov, cam_ov_start;      // observations without a camera slot in their tile
  int n_ovf = 0, ovf_cap = 0;
  int occ_pt2 = 0;
  // explicit block-sparse reduced camera matrix (glba_sparse.cuh): structure built lazily at the first PCG solve of a loaded problem
  bool env_explicit = true;      // diagnostic: GLBA_EXPLICIT=0 keeps the implicit (matrix-free) product
  bool cg_reg = false;           // the loaded map runs the one-row-per-warp PCG kernel (k_cg_bsr<true>) on cg_grid_reg CTAs
  int cg_grid_reg = 0;
  bool env_cg_prof = false;      // diagnostic (GLBA_CG_PROF=1): per-phase cycle counts of the PCG kernel on stderr
  bool env_cg_reg = true;        // diagnostic: GLBA_CG_REG=0 runs the general PCG kernel (several rows per warp, vectors in global memory) on small maps too
  bool sp_tried = false, use_explicit = false;
  bool cg_pending = false;       // h_cg of the last cooperative PCG is in flight: valid after the next stream synchronisation
  int n_pairs = 0;
  long long n_inst = 0;
  Buf sp_cnt, sp_off, sp_key, sp_key2, sp_val, sp_inst, sp_ukey, sp_ucnt, sp_nruns, sp_pair_a, sp_pair_b, sp_pair_start;
  Buf sp_ekey, sp_ekey2, sp_eval, sp_eval2, sp_erow, sp_ent, sp_row_start, sp_blocks, sp_part, sp_bar, sp_pres, sp_gscan, sp_gkey, sp_gid, sp_vec, sp_prof, sp_cta_rows;
  int n_pairs_g = 0;             // blocks of the whole map (== n_pairs on one GPU)
  size_t sp_xch_len = 0;         // doubles of [blocks | sharded: Md Minv rhs of the whole map]
  int cg_grid = 0;               // CTAs of the cooperative PCG kernel (all resident)
  int occ_lin = 0, occ_pt0 = 0, occ_pt1 = 0;   // resident CTAs per SM of the pipelined kernels (occupancy API, per context)
  int opt = OPT_LARGE;          // observations per thread of the tile kernels (tile capacity = NT_T * opt)
  int n_tiles = 0, max_track = 0, grid_c = 0;
  Buf dn_part, dn_red, dn_full;                                                               // dense path: per-CTA S copies, reduced S
  bool has_dup = false;                                                              // some point is observed twice by one camera
  int dn_grid = 0, dn_ppc = 0;
  int cur = 0;
  int n_sm = 148;              // multiprocessors of ctx->device (queried in glba_create)
  bool attr_done = false;      // cudaFuncSetAttribute is per DEVICE: done once per context, after cudaSetDevice
  double* h_scal = nullptr;    // pinned
  CgState* h_cg = nullptr;     // pinned
  int* h_flags = nullptr;      // pinned
  // timing
  std::vector<cudaEvent_t> ev;
  std::vector<int> ev_phase;
  int ev_used = 0;
  bool timing = true;          // per-phase CUDA events (off for small problems: the event API calls rival the kernels)
  double t_phase[PH_COUNT] = {0, 0, 0, 0, 0, 0};
};

namespace {

int fail(glba_ctx* c, int code, const char* fmt, ...) {
  if (c) {
    char buf[512];
    va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof(buf), fmt, ap); va_end(ap);
    c->err = buf;
  }
  return code;
}

#define CU(call)                                                                                         \
  do {                                                                                                   \
    cudaError_t e__ = (call);                                                                            \
    if (e__ != cudaSuccess)                                                                              \
      return fail(ctx, e__ == cudaErrorMemoryAllocation ? GLBA_E_OOM : GLBA_E_CUDA, "%s:%d %s: %s", __FILE__, __LINE__, #call, \
                  cudaGetErrorString(e__));                                                              \
  } while (0)

// Every kernel is launched with programmatic stream serialisation allowed (PDL, see pdl_grid_sync in glba_kernels.cuh): the
// launch latency of kernel N+1 hides behind kernel N, which is what the small-window solves and the PCG iteration are made
// of (8-13 dependent launches of 5-35 us each).  GLBA_PDL=0 launches plainly (diagnostic).
// Measured on C4 (tools/ab_pdl.sh): a kernel allowed to start early behind a multi-wave or persistent kernel makes that
// transition 10-18 us SLOWER (step 0.389 -> 0.425 ms, LM 165 -> 127 it/s with every launch flagged), while chains of small
// kernels gain ~2 us per launch (C2 window 3.5 -> 3.0 ms).  So a launch takes the attribute only when it AND the kernel
// launched before it are small (at most two CTAs per SM): that is the small-window solve; large maps launch plainly.
bool g_pdl = true;
unsigned g_pdl_max_grid = 296;
struct PrevLaunch { cudaStream_t stream; bool small; };
thread_local PrevLaunch g_prev[4] = {{nullptr, false}, {nullptr, false}, {nullptr, false}, {nullptr, false}};     // per stream (a context uses two)
inline bool& prev_small(cudaStream_t s) {
  for (auto& e : g_prev) if (e.stream == s) return e.small;
  static thread_local int next = 0;
  PrevLaunch& e = g_prev[next++ & 3];
  e.stream = s; e.small = false;
  return e.small;
}
template <typename... KArgs, typename... Args>
inline void launch_kernel(cudaStream_t stream, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  const bool small = grid.x * grid.y * grid.z <= g_pdl_max_grid;
  bool& prev = prev_small(stream);
  cfg.attrs = attr; cfg.numAttrs = (g_pdl && small && prev) ? 1 : 0;
  prev = small;
  (void)cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);       // launch errors are sticky: check_launches() per phase
  g_launches.fetch_add(1, std::memory_order_relaxed);
}
#define LAUNCH(kernel, grid, block, ...) launch_kernel(ctx->stream, kernel, dim3(grid), dim3(block), 0, __VA_ARGS__)
#define LAUNCH_SMEM(kernel, grid, block, smem, ...) launch_kernel(ctx->stream, kernel, dim3(grid), dim3(block), (size_t)(smem), __VA_ARGS__)

template <typename T>
int ensure(glba_ctx* ctx, Buf& b, size_t n) {
  const size_t bytes = std::max<size_t>(n, 1) * sizeof(T);
  if (b.cap >= bytes) return GLBA_OK;
  if (b.p) { cudaFree(b.p); b.p = nullptr; b.cap = 0; }
  const size_t want = bytes + bytes / 8 + 256;
  CU(cudaMalloc(&b.p, want));
  b.cap = want;
  return GLBA_OK;
}
#define ENSURE(T, buf, n) do { int s__ = ensure<T>(ctx, buf, n); if (s__) return s__; } while (0)

void release(Buf& b) { if (b.p) cudaFree(b.p); b.p = nullptr; b.cap = 0; }

inline int cdiv(long a, int b) { return (int)((a + b - 1) / b); }

void mark(glba_ctx* ctx, int phase) {
  if (!ctx->timing) return;
  if (ctx->ev_used == (int)ctx->ev.size()) {
    cudaEvent_t e; cudaEventCreate(&e); ctx->ev.push_back(e); ctx->ev_phase.push_back(0);
  }
  cudaEventRecord(ctx->ev[ctx->ev_used], ctx->stream);
  ctx->ev_phase[ctx->ev_used] = phase;
  ctx->ev_used++;
}
// after a stream sync: attribute the time between consecutive markers to the earlier marker's phase
void collect(glba_ctx* ctx) {
  for (int i = 0; i + 1 < ctx->ev_used; ++i) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, ctx->ev[i], ctx->ev[i + 1]) == cudaSuccess && ctx->ev_phase[i] >= 0) ctx->t_phase[ctx->ev_phase[i]] += ms;
  }
  ctx->ev_used = 0;
}

int allreduce(glba_ctx* ctx, void* p, size_t n, int op, int dtype = kNcclFloat64) {
  if (ctx->world <= 1) return GLBA_OK;
  const int r = g_nccl.AllReduce(p, p, n, dtype, op, ctx->comm, ctx->stream);
  if (r != 0) return fail(ctx, GLBA_E_NCCL, "ncclAllReduce: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "error");
  return GLBA_OK;
}
#define AR(p, n, op) do { int s__ = allreduce(ctx, p, n, op); if (s__) return s__; } while (0)

// launches are asynchronous and unchecked one by one; every phase ends with this (sticky launch-configuration errors)
int check_launches(glba_ctx* ctx) {
  CU(cudaGetLastError());
  return GLBA_OK;
}
#define CHECK_LAUNCHES() do { int s__ = check_launches(ctx); if (s__) return s__; } while (0)

int validate_problem(glba_ctx* ctx, const glba_problem* p) {
  if (!p) return fail(ctx, GLBA_E_INVALID_ARG, "problem is NULL");
  if (p->n_cam < 0 || p->n_pt < 0 || p->n_obs < 0) return fail(ctx, GLBA_E_INVALID_ARG, "negative size");
  if (p->n_obs > 0x7fffffffL) return fail(ctx, GLBA_E_UNSUPPORTED, "n_obs exceeds int32 indexing");
  if (p->n_cam > 0 && !p->cam) return fail(ctx, GLBA_E_INVALID_ARG, "cam is NULL");
  if (p->n_pt > 0 && !p->pt) return fail(ctx, GLBA_E_INVALID_ARG, "pt is NULL");
  if (p->n_obs > 0 && (!p->obs_cam || !p->obs_pt || !p->obs_u || !p->obs_v)) return fail(ctx, GLBA_E_INVALID_ARG, "observation array is NULL");
  if (p->n_obs > 0 && (p->n_cam == 0 || p->n_pt == 0)) return fail(ctx, GLBA_E_INVALID_ARG, "observations without cameras/points");
  if (p->memspace != GLBA_MEM_HOST && p->memspace != GLBA_MEM_DEVICE) return fail(ctx, GLBA_E_INVALID_ARG, "bad memspace");
  return GLBA_OK;
}

int validate_options(glba_ctx* ctx, const glba_options* o) {
  if (!o) return fail(ctx, GLBA_E_INVALID_ARG, "options is NULL");
  if (o->max_iters < 0 || o->max_iters > GLBA_MAX_ITERS) return fail(ctx, GLBA_E_INVALID_ARG, "max_iters out of [0,%d]", GLBA_MAX_ITERS);
  if (o->loss < GLBA_LOSS_NONE || o->loss > GLBA_LOSS_CAUCHY) return fail(ctx, GLBA_E_INVALID_ARG, "unknown loss");
  if (!(o->loss_scale > 0.0) || !(o->initial_radius > 0.0)) return fail(ctx, GLBA_E_INVALID_ARG, "loss_scale and initial_radius must be > 0");
  if (o->mode != GLBA_MODE_CERES && o->mode != GLBA_MODE_G2O) return fail(ctx, GLBA_E_INVALID_ARG, "unknown mode");
  if (o->mode == GLBA_MODE_G2O && (!(o->g2o_tau > 0.0) || o->g2o_max_trials < 1)) return fail(ctx, GLBA_E_INVALID_ARG, "g2o_tau must be > 0 and g2o_max_trials >= 1");
  return GLBA_OK;
}

// Opt-in shared-memory sizes are a per-DEVICE property of a kernel: set them for the context's device when the
// context is created (a process may hold contexts on several devices, and on several threads).
int set_func_attributes(glba_ctx* ctx) {
  CU(cudaFuncSetAttribute(k_linearize_tile<OPT_LARGE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((size_t)8 * NT_T * OPT_LARGE * sizeof(double))));
  CU(cudaFuncSetAttribute(k_dense_schur, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(((size_t)DN_TP * DN_MAXCAM * 24 + (size_t)(DN_MAXCAM * (DN_MAXCAM + 1) / 2) * 36) * 8 + DN_TP * 4 + 64)));
  CU(cudaFuncSetAttribute(k_dense_solve, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(((size_t)(6 * DN_MAXCAM) * (6 * DN_MAXCAM + 1) + 12 * DN_MAXCAM) * 8 + 1024)));
  CU(cudaFuncSetAttribute(k_cg_bsr<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CG_SMEM_BYTES));
  CU(cudaFuncSetAttribute(k_cam_pipe<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(CamSmem)));
  CU(cudaFuncSetAttribute(k_cam_pipe<2>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
  CU(cudaFuncSetAttribute(k_cam_pipe<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(CamSmem)));
  CU(cudaFuncSetAttribute(k_cam_pipe<3>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
  CU(cudaFuncSetAttribute(k_lin_pipe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(LinSmem)));
  CU(cudaFuncSetAttribute(k_pt_pipe<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(PtSmem<0>)));
  CU(cudaFuncSetAttribute(k_pt_pipe<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(PtSmem<1>)));
  CU(cudaFuncSetAttribute(k_lin_pipe, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
  CU(cudaFuncSetAttribute(k_pt_pipe<0>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
  CU(cudaFuncSetAttribute(k_pt_pipe<1>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
  CU(cudaFuncSetAttribute(k_pt_pipe<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(PtSmem<2>)));
  CU(cudaFuncSetAttribute(k_pt_pipe<2>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
  CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctx->occ_pt2, k_pt_pipe<2>, P_NT, sizeof(PtSmem<2>)));
  CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctx->occ_lin, k_lin_pipe, P_NT, sizeof(LinSmem)));
  CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctx->occ_pt0, k_pt_pipe<0>, P_NT, sizeof(PtSmem<0>)));
  CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctx->occ_pt1, k_pt_pipe<1>, P_NT, sizeof(PtSmem<1>)));
  if (ctx->occ_lin < 1 || ctx->occ_pt0 < 1 || ctx->occ_pt1 < 1) return fail(ctx, GLBA_E_CUDA, "pipelined tile kernels do not fit an SM");
  if (const char* e = std::getenv("GLBA_OCC")) {          // diagnostic: cap the resident CTAs per SM of the pipelined kernels
    const int cap = std::max(1, std::atoi(e));
    ctx->occ_lin = std::min(ctx->occ_lin, cap); ctx->occ_pt0 = std::min(ctx->occ_pt0, cap); ctx->occ_pt1 = std::min(ctx->occ_pt1, cap);
    ctx->occ_pt2 = std::min(ctx->occ_pt2, cap);
  }
  ctx->attr_done = true;
  return GLBA_OK;
}

// ---------------------------------------------------------------------------------------------
// Upload + index construction (the device-side restatement of the packing loop, slam_core.cpp:750-819)
// ---------------------------------------------------------------------------------------------
int load_problem_impl(glba_ctx* ctx, const glba_problem* p) {
  int st = validate_problem(ctx, p);
  if (st) return st;
  ctx->loaded = false;
  ctx->ev_used = 0;
  mark(ctx, PH_SETUP);
  int n_cam = p->n_cam;
  const int n_pt = p->n_pt;
  const long n = p->n_obs;
  ctx->n_cam = n_cam; ctx->n_cam_g = n_cam; ctx->n_pt = n_pt; ctx->n_obs = n;
  ctx->owner = false; ctx->n_shared = 0;
  ctx->K = Intr{p->fx, p->fy, p->cx, p->cy};
  cudaStream_t s = ctx->stream;
  const double *d_cam, *d_pt, *d_u, *d_v;
  const int *d_ocam, *d_opt;
  const uint8_t *d_cfix = nullptr, *d_pfix = nullptr;
  const double* d_info = nullptr;
  const bool staged = (p->memspace == GLBA_MEM_HOST);
  if (staged) {
    ENSURE(double, ctx->in_cam, 6 * (size_t)n_cam); ENSURE(double, ctx->in_pt, 3 * (size_t)n_pt);
    ENSURE(int, ctx->in_ocam, n); ENSURE(int, ctx->in_opt, n); ENSURE(double, ctx->in_u, n); ENSURE(double, ctx->in_v, n);
    if (p->cam_fixed) { ENSURE(uint8_t, ctx->in_cfix, n_cam); d_cfix = ctx->in_cfix.as<uint8_t>(); }
    if (p->pt_fixed) { ENSURE(uint8_t, ctx->in_pfix, n_pt); d_pfix = ctx->in_pfix.as<uint8_t>(); }
    if (p->pt_info) { ENSURE(double, ctx->in_info, n_pt); d_info = ctx->in_info.as<double>(); }
    // uploads on their own stream, in the order the index construction needs them: indices | parameters | measurements
    cudaStream_t cs = ctx->copy_stream;
    CU(cudaEventRecord(ctx->ev_copy[3], s));
    CU(cudaStreamWaitEvent(cs, ctx->ev_copy[3], 0));       // earlier work on the compute stream may still read the staging buffers
    ctx->copies_in_flight = true;
    CU(cudaMemcpyAsync(ctx->in_opt.p, p->obs_pt, sizeof(int) * n, cudaMemcpyHostToDevice, cs));
    CU(cudaMemcpyAsync(ctx->in_ocam.p, p->obs_cam, sizeof(int) * n, cudaMemcpyHostToDevice, cs));
    CU(cudaEventRecord(ctx->ev_copy[0], cs));
    if (p->cam_fixed) CU(cudaMemcpyAsync(ctx->in_cfix.p, p->cam_fixed, n_cam, cudaMemcpyHostToDevice, cs));
    if (p->pt_fixed) CU(cudaMemcpyAsync(ctx->in_pfix.p, p->pt_fixed, n_pt, cudaMemcpyHostToDevice, cs));
    if (p->pt_info) CU(cudaMemcpyAsync(ctx->in_info.p, p->pt_info, sizeof(double) * n_pt, cudaMemcpyHostToDevice, cs));
    CU(cudaMemcpyAsync(ctx->in_cam.p, p->cam, sizeof(double) * 6 * n_cam, cudaMemcpyHostToDevice, cs));
    CU(cudaMemcpyAsync(ctx->in_pt.p, p->pt, sizeof(double) * 3 * n_pt, cudaMemcpyHostToDevice, cs));
    CU(cudaEventRecord(ctx->ev_copy[1], cs));
    CU(cudaMemcpyAsync(ctx->in_u.p, p->obs_u, sizeof(double) * n, cudaMemcpyHostToDevice, cs));
    CU(cudaMemcpyAsync(ctx->in_v.p, p->obs_v, sizeof(double) * n, cudaMemcpyHostToDevice, cs));
    CU(cudaEventRecord(ctx->ev_copy[2], cs));
    CU(cudaStreamWaitEvent(s, ctx->ev_copy[0], 0));        // the index kernels below need obs_pt / obs_cam only
    d_cam = ctx->in_cam.as<double>(); d_pt = ctx->in_pt.as<double>(); d_ocam = ctx->in_ocam.as<int>(); d_opt = ctx->in_opt.as<int>();
    d_u = ctx->in_u.as<double>(); d_v = ctx->in_v.as<double>();
  } else {
    d_cam = p->cam; d_pt = p->pt; d_ocam = p->obs_cam; d_opt = p->obs_pt; d_u = p->obs_u; d_v = p->obs_v;
    d_cfix = p->cam_fixed; d_pfix = p->pt_fixed; d_info = p->pt_info;
  }
  // state
  for (int b = 0; b < 2; ++b) { ENSURE(double, ctx->cam[b], 6 * (size_t)n_cam); ENSURE(double, ctx->camtab[b], (size_t)CAMTAB * n_cam); ENSURE(double4, ctx->pt4[b], n_pt); }
  ENSURE(double, ctx->cam0, 6 * (size_t)n_cam); ENSURE(double4, ctx->pt40, n_pt);
  ctx->cur = 0;
  // index buffers
  ENSURE(int, ctx->pm_cam, n); ENSURE(int, ctx->pm_pt, n); ENSURE(double2, ctx->pm_uv, n); ENSURE(int, ctx->pm2cm, n);
  ENSURE(int, ctx->pt_start, (size_t)n_pt + 1); ENSURE(int, ctx->cm_pt, n); ENSURE(double2, ctx->cm_uv, n); ENSURE(int, ctx->cm2pm, n);
  ENSURE(int, ctx->cam_start, (size_t)n_cam + 1); ENSURE(uint8_t, ctx->cam_free, n_cam); ENSURE(uint8_t, ctx->pt_free, n_pt);
  ENSURE(int, ctx->keys_tmp, 2 * (size_t)n + 2); ENSURE(int, ctx->flags, 16);
  ENSURE(int, ctx->pm2orig, n);
  CU(cudaMemsetAsync(ctx->flags.p, 0, 16 * sizeof(int), s));
  const int gb = cdiv(n, 256);
  if (n > 0) {
    LAUNCH(k_check_sorted, gb, 256, n, d_opt, n_pt, ctx->flags.as<int>());
    LAUNCH(k_check_range, gb, 256, n, d_ocam, n_cam, ctx->flags.as<int>() + 2);
  }
  CU(cudaMemcpyAsync(ctx->h_flags, ctx->flags.p, 4 * sizeof(int), cudaMemcpyDeviceToHost, s));
  CU(cudaStreamSynchronize(s));
  if (ctx->h_flags[1] || ctx->h_flags[2]) return fail(ctx, GLBA_E_INVALID_ARG, "observation index out of range");
  ctx->sorted_input = (ctx->h_flags[0] == 0);
  // ---- sharded maps, "owner-computes" layout ---------------------------------------------------------------------------
  // Round 1 replicated every camera-sized array on every rank and all-reduced all cameras' partial sums (6.2 MB per step at
  // 8 x 1 800 cameras, of which a rank touched 1 800).  Here a rank's problem holds only the cameras ITS tracks observe
  // (local camera ids), so every camera kernel and vector is local-sized.  Cameras observed by several ranks ("shared":
  // shard boundaries, revisited streets, loop closures) have their partial sums exchanged through a compact buffer — one
  // all-reduce per linearisation and one per PCG iteration, of n_shared rows —; every scalar counts a camera once, on its
  // OWNER (the lowest rank that observes it).  Executable specification: tests/test_owner_computes_spec.py.
  if (ctx->world > 1 && ctx->want_owner) {
    const int n_cam_g = n_cam;
    ENSURE(int, ctx->act_mask, (size_t)n_cam_g + 1); ENSURE(int, ctx->g2l, 2 * ((size_t)n_cam_g + 1)); ENSURE(int, ctx->sh_scan, 2 * ((size_t)n_cam_g + 1));
    CU(cudaMemsetAsync(ctx->act_mask.p, 0, sizeof(int) * ((size_t)n_cam_g + 1), s));
    if (n > 0) LAUNCH(k_act_mask, gb, 256, n, d_ocam, 1 << ctx->rank, ctx->act_mask.as<int>());
    { int s__ = allreduce(ctx, ctx->act_mask.p, (size_t)n_cam_g, kNcclSum, kNcclInt32); if (s__) return s__; }
    int* f_act = ctx->g2l.as<int>() + n_cam_g + 1;        // flags; the scans land in g2l[0..n_cam_g] / sh_scan[0..n_cam_g]
    int* f_sh = ctx->sh_scan.as<int>() + n_cam_g + 1;
    LAUNCH(k_own_flags, cdiv(n_cam_g + 1, 256), 256, n_cam_g, (const int*)ctx->act_mask.as<int>(), ctx->rank, f_act, f_sh);
    size_t tb = 0;
    CU(cub::DeviceScan::ExclusiveSum(nullptr, tb, f_act, ctx->g2l.as<int>(), n_cam_g + 1, s));
    ENSURE(char, ctx->sort_tmp, tb);
    tb = ctx->sort_tmp.cap;
    CU(cub::DeviceScan::ExclusiveSum(ctx->sort_tmp.p, tb, f_act, ctx->g2l.as<int>(), n_cam_g + 1, s));
    tb = ctx->sort_tmp.cap;
    CU(cub::DeviceScan::ExclusiveSum(ctx->sort_tmp.p, tb, f_sh, ctx->sh_scan.as<int>(), n_cam_g + 1, s));
    g_launches.fetch_add(2);
    CU(cudaMemcpyAsync(ctx->h_flags, ctx->g2l.as<int>() + n_cam_g, sizeof(int), cudaMemcpyDeviceToHost, s));
    CU(cudaMemcpyAsync(ctx->h_flags + 1, ctx->sh_scan.as<int>() + n_cam_g, sizeof(int), cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    const int n_act = ctx->h_flags[0];
    ctx->n_shared = ctx->h_flags[1];
    ENSURE(int, ctx->l2g, n_act); ENSURE(uint8_t, ctx->cam_owned, n_act); ENSURE(int, ctx->cam_shared, n_act);
    ENSURE(int, ctx->ocam_loc, n); ENSURE(double, ctx->cam_loc, 6 * (size_t)n_act); ENSURE(uint8_t, ctx->cfix_loc, n_act);
    if (staged) CU(cudaStreamWaitEvent(s, ctx->ev_copy[1], 0));      // camera parameters and fixed flags have arrived
    CU(cudaMemsetAsync(ctx->flags.as<int>() + 7, 0, sizeof(int), s));
    LAUNCH(k_own_build, cdiv(n_cam_g, 256), 256, n_cam_g, (const int*)ctx->act_mask.as<int>(), ctx->rank, (const int*)ctx->g2l.as<int>(),
           (const int*)ctx->sh_scan.as<int>(), d_cfix, ctx->l2g.as<int>(), ctx->cam_owned.as<uint8_t>(), ctx->cam_shared.as<int>(), ctx->flags.as<int>() + 7);
    // free cameras of the WHOLE map (every rank needs the same number: iteration caps, "nothing free")
    { int s__ = allreduce(ctx, ctx->flags.as<int>() + 7, 1, kNcclSum, kNcclInt32); if (s__) return s__; }
    CU(cudaMemcpyAsync(ctx->h_flags + 7, ctx->flags.as<int>() + 7, sizeof(int), cudaMemcpyDeviceToHost, s));
    // shared slot -> local camera (or -1): the dense payload of the peer exchange
    ENSURE(int, ctx->sh2loc, (size_t)ctx->n_shared + 1);
    CU(cudaMemsetAsync(ctx->sh2loc.p, 0xff, sizeof(int) * ((size_t)ctx->n_shared + 1), s));
    if (n_act) LAUNCH(k_xch_inverse, cdiv(n_act, 256), 256, n_act, (const int*)ctx->cam_shared.as<int>(), ctx->sh2loc.as<int>());
    if (n > 0) LAUNCH(k_relabel_cam, gb, 256, n, d_ocam, (const int*)ctx->g2l.as<int>(), ctx->ocam_loc.as<int>());
    if (n_act) LAUNCH(k_gather_cam, cdiv(n_act, 256), 256, n_act, (const int*)ctx->l2g.as<int>(), d_cam, d_cfix, ctx->cam_loc.as<double>(), ctx->cfix_loc.as<uint8_t>());
    CU(cudaStreamSynchronize(s));
    ctx->n_free_cam_g = ctx->h_flags[7];
    d_ocam = ctx->ocam_loc.as<int>(); d_cam = ctx->cam_loc.as<double>(); d_cfix = ctx->cfix_loc.as<uint8_t>();
    n_cam = n_act; ctx->n_cam = n_act; ctx->owner = true;
    const size_t xlen = 54 * (size_t)ctx->n_shared + NSCAL + 8;
    ENSURE(double, ctx->xsend, xlen); ENSURE(double, ctx->xrecv, xlen); ENSURE(double, ctx->late, 2 * NLATE);
    // two send buffers (54-wide rows of a linearisation, 6-wide rows of a PCG iteration): in each, the rows of cameras this rank
    // does not observe are zeroed here once and never written again
    ENSURE(double, ctx->xsend6, 6 * (size_t)ctx->n_shared + 8);
    CU(cudaMemsetAsync(ctx->xsend.p, 0, sizeof(double) * xlen, s));
    CU(cudaMemsetAsync(ctx->xsend6.p, 0, sizeof(double) * (6 * (size_t)ctx->n_shared + 8), s));
  }
  // ---- locality relabelling: the tile kernels stage a narrow window of cameras per tile and the camera-major gathers
  // want neighbouring observations to touch neighbouring points, both of which hold when point ids are ordered by their
  // first-observing camera (how GL-SLAM numbers map points).  If the caller's numbering is not like that (measured: 13 %
  // slower steps, 1.9x slower point-major products on C4), renumber internally; results are mapped back on exit.
  ctx->relabelled = false;
  if (n > 0 && n_pt > 1 && ctx->allow_relabel && ctx->env_relabel) {
    ENSURE(int, ctx->first_cam, n_pt);
    LAUNCH(k_fill_int, cdiv(n_pt, 256), 256, n_pt, ctx->first_cam.as<int>(), 0x7fffffff);
    LAUNCH(k_first_cam, gb, 256, n, d_ocam, d_opt, ctx->first_cam.as<int>());
    CU(cudaMemsetAsync(ctx->flags.as<int>() + 7, 0, sizeof(int), s));
    LAUNCH(k_count_descents, cdiv(n_pt, 256), 256, n_pt, (const int*)ctx->first_cam.as<int>(), ctx->flags.as<int>() + 7);
    CU(cudaMemcpyAsync(ctx->h_flags + 7, ctx->flags.as<int>() + 7, sizeof(int), cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    if (ctx->h_flags[7] > n_pt / 20) {
      ENSURE(int, ctx->new2old, n_pt); ENSURE(int, ctx->old2new, n_pt); ENSURE(int, ctx->opt_relab, n);
      int* iota_p = ctx->keys_tmp.as<int>();
      int* keys_sorted = ctx->keys_tmp.as<int>() + n + 1;
      LAUNCH(k_iota, cdiv(n_pt, 256), 256, (long)n_pt, iota_p);
      size_t tmp_bytes = 0;
      CU(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, ctx->first_cam.as<int>(), keys_sorted, iota_p, ctx->new2old.as<int>(), n_pt, 0, 32, s));
      ENSURE(char, ctx->sort_tmp, tmp_bytes);
      CU(cub::DeviceRadixSort::SortPairs(ctx->sort_tmp.p, tmp_bytes, ctx->first_cam.as<int>(), keys_sorted, iota_p, ctx->new2old.as<int>(), n_pt, 0, 32, s));
      g_launches.fetch_add(1);
      LAUNCH(k_invert_perm, cdiv(n_pt, 256), 256, n_pt, (const int*)ctx->new2old.as<int>(), ctx->old2new.as<int>());
      LAUNCH(k_relabel, gb, 256, n, d_opt, (const int*)ctx->old2new.as<int>(), ctx->opt_relab.as<int>());
      d_opt = ctx->opt_relab.as<int>();
      ctx->relabelled = true;
      ctx->sorted_input = false;       // the relabelled observation list is re-sorted by (new) point below
    }
  }
  const int* n2o = ctx->relabelled ? ctx->new2old.as<int>() : nullptr;
  int bits_pt = 1; while ((1L << bits_pt) < (long)n_pt + 1 && bits_pt < 31) ++bits_pt;
  int bits_cam = 1; while ((1L << bits_cam) < (long)n_cam + 1 && bits_cam < 31) ++bits_cam;
  int* iota = ctx->keys_tmp.as<int>();          // [0,n): iota / sorted keys scratch; [n,2n): keys out
  int* keys_out = ctx->keys_tmp.as<int>() + n + 1;
  if (n > 0) {
    if (ctx->sorted_input) {
      LAUNCH(k_gather_obs, gb, 256, n, (const int*)nullptr, d_ocam, d_opt, d_u, d_v, ctx->pm_cam.as<int>(), ctx->pm_pt.as<int>(), (double2*)nullptr);
    } else {
      LAUNCH(k_iota, gb, 256, n, iota);
      size_t tmp_bytes = 0;
      CU(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, d_opt, keys_out, iota, ctx->pm2orig.as<int>(), (int)n, 0, bits_pt, s));
      ENSURE(char, ctx->sort_tmp, tmp_bytes);
      CU(cub::DeviceRadixSort::SortPairs(ctx->sort_tmp.p, tmp_bytes, d_opt, keys_out, iota, ctx->pm2orig.as<int>(), (int)n, 0, bits_pt, s));
      g_launches.fetch_add(1);
      LAUNCH(k_gather_obs, gb, 256, n, (const int*)ctx->pm2orig.as<int>(), d_ocam, d_opt, d_u, d_v, ctx->pm_cam.as<int>(), ctx->pm_pt.as<int>(), (double2*)nullptr);
    }
    LAUNCH(k_segment_starts, cdiv(n_pt + 1, 256), 256, n, (const int*)ctx->pm_pt.as<int>(), n_pt, ctx->pt_start.as<int>());
    // camera-major order: stable sort of point-major positions by camera
    LAUNCH(k_iota, gb, 256, n, iota);
    size_t tmp_bytes = 0;
    CU(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, ctx->pm_cam.as<int>(), keys_out, iota, ctx->cm2pm.as<int>(), (int)n, 0, bits_cam, s));
    ENSURE(char, ctx->sort_tmp, tmp_bytes);
    CU(cub::DeviceRadixSort::SortPairs(ctx->sort_tmp.p, tmp_bytes, ctx->pm_cam.as<int>(), keys_out, iota, ctx->cm2pm.as<int>(), (int)n, 0, bits_cam, s));
    g_launches.fetch_add(1);
    LAUNCH(k_segment_starts, cdiv(n_cam + 1, 256), 256, n, (const int*)keys_out, n_cam, ctx->cam_start.as<int>());
    LAUNCH(k_build_cm, gb, 256, n, (const int*)ctx->cm2pm.as<int>(), (const int*)ctx->pm_pt.as<int>(), (const double2*)ctx->pm_uv.as<double2>(),
           ctx->cm_pt.as<int>(), (double2*)nullptr, ctx->pm2cm.as<int>());
  } else {
    CU(cudaMemsetAsync(ctx->pt_start.p, 0, sizeof(int) * ((size_t)n_pt + 1), s));
    CU(cudaMemsetAsync(ctx->cam_start.p, 0, sizeof(int) * ((size_t)n_cam + 1), s));
  }
  // a camera is part of the problem if ANY rank observes it: all-reduce the per-camera observation counts
  // (and the duplicate flag) so every rank takes the same code path and issues the same collectives
  ENSURE(int, ctx->cam_cnt, (size_t)n_cam + 2);
  if (n > 0 && n_cam <= DN_MAXCAM)
    LAUNCH(k_check_dup, cdiv(n_pt, 256), 256, n_pt, (const int*)ctx->pt_start.as<int>(), (const int*)ctx->pm_cam.as<int>(), ctx->flags.as<int>() + 3);
  LAUNCH(k_counts, cdiv(n_cam + 2, 256), 256, n_cam, (const int*)ctx->cam_start.as<int>(), (const int*)ctx->flags.as<int>() + 3, n > 0 ? 0 : 1, ctx->cam_cnt.as<int>());
  if (ctx->world > 1) {
    // owner layout: every local camera is observed here, only the duplicate / empty-shard flags are global
    int s__ = ctx->owner ? allreduce(ctx, ctx->cam_cnt.as<int>() + n_cam, 2, kNcclSum, kNcclInt32)
                         : allreduce(ctx, ctx->cam_cnt.p, (size_t)n_cam + 2, kNcclSum, kNcclInt32);
    if (s__) return s__;
  }
  if (staged) CU(cudaStreamWaitEvent(s, ctx->ev_copy[1], 0));   // cam_fixed / pt_fixed / cam / pt have arrived
  if (n_cam) LAUNCH(k_free_flags_cnt, cdiv(n_cam, 256), 256, n_cam, (const int*)ctx->cam_cnt.as<int>(), d_cfix, ctx->cam_free.as<uint8_t>());
  if (n_pt) LAUNCH(k_free_flags, cdiv(n_pt, 256), 256, n_pt, (const int*)ctx->pt_start.as<int>(), d_pfix, n2o, ctx->pt_free.as<uint8_t>());
  ctx->has_dup = false;
  ctx->sp_tried = false; ctx->use_explicit = false; ctx->n_pairs = 0; ctx->n_inst = 0;
  if (n > 0) LAUNCH(k_max_track, cdiv(n_pt, 256), 256, n_pt, (const int*)ctx->pt_start.as<int>(), ctx->flags.as<int>() + 4);
  CU(cudaMemcpyAsync(ctx->h_flags, ctx->flags.p, 8 * sizeof(int), cudaMemcpyDeviceToHost, s));
  CU(cudaMemcpyAsync(ctx->h_flags + 5, ctx->cam_cnt.as<int>() + n_cam, 2 * sizeof(int), cudaMemcpyDeviceToHost, s));   // global duplicate flag, #empty shards
  // chunk list for the camera-major kernels (host, from cam_start)
  std::vector<int> h_start((size_t)n_cam + 1);
  std::vector<uint8_t> h_free(std::max(n_cam, 1));
  CU(cudaMemcpyAsync(h_start.data(), ctx->cam_start.p, sizeof(int) * ((size_t)n_cam + 1), cudaMemcpyDeviceToHost, s));
  if (n_cam) CU(cudaMemcpyAsync(h_free.data(), ctx->cam_free.p, n_cam, cudaMemcpyDeviceToHost, s));
  CU(cudaStreamSynchronize(s));
  ctx->has_dup = (ctx->h_flags[5] != 0);
  if (ctx->world > 1 && ctx->h_flags[6] != 0) return fail(ctx, GLBA_E_INVALID_ARG, "a rank holds an empty shard (%d of %d): every rank needs at least one observation", ctx->h_flags[6], ctx->world);
  ctx->max_track = ctx->h_flags[4];
  ctx->opt = (n >= 400000 || ctx->env_force_large) ? OPT_LARGE : OPT_SMALL;
  if (ctx->max_track > NT_T * ctx->opt / 2) ctx->opt = OPT_LARGE;
  ctx->use_tiles = (n > 0 && ctx->max_track <= NT_T * ctx->opt / 2);
  ctx->n_tiles = 0;
  if (ctx->use_tiles) {
    const int B = NT_T * ctx->opt - ctx->max_track;
    ctx->n_tiles = (int)((n + B - 1) / B);
    ENSURE(int, ctx->tile_pt, (size_t)ctx->n_tiles + 1);
    LAUNCH(k_tile_starts, cdiv(ctx->n_tiles + 1, 256), 256, ctx->n_tiles, B, n_pt, (const int*)ctx->pt_start.as<int>(), ctx->tile_pt.as<int>());
    ENSURE(int, ctx->tile_cmin, (size_t)ctx->n_tiles);
    LAUNCH(k_tile_cmin, ctx->n_tiles, NT_T, (const int*)ctx->tile_pt.as<int>(), (const int*)ctx->pt_start.as<int>(), (const int*)ctx->pm_cam.as<int>(), ctx->tile_cmin.as<int>());
  }
  ctx->use_pipe = ctx->use_tiles && ctx->opt == OPT_LARGE && ctx->env_pipe;
  if (ctx->use_pipe) {      // per-tile descriptor, distinct-camera list and per-observation camera slot of the pipelined kernels
    ENSURE(int4, ctx->tile_desc, (size_t)ctx->n_tiles); ENSURE(int, ctx->tile_cams, (size_t)TSLOTS * ctx->n_tiles); ENSURE(uint8_t, ctx->pm_slot, n);
    CU(cudaMemsetAsync(ctx->tile_cams.p, 0xff, sizeof(int) * (size_t)TSLOTS * ctx->n_tiles, s));
    ENSURE(uint16_t, ctx->tile_sobs, n); ENSURE(uint16_t, ctx->tile_sstart, (size_t)SSTART * ctx->n_tiles);
    ctx->ovf_cap = (int)(n / 32 + 1024);
    ENSURE(int, ctx->ovf_raw, ctx->ovf_cap);
    CU(cudaMemsetAsync(ctx->flags.as<int>() + 8, 0, sizeof(int), s));
    LAUNCH(k_tile_meta, ctx->n_tiles, NT_T, (const int*)ctx->tile_pt.as<int>(), (const int*)ctx->pt_start.as<int>(), (const int*)ctx->pm_cam.as<int>(), n_cam,
           ctx->tile_desc.as<int4>(), ctx->tile_cams.as<int>(), ctx->pm_slot.as<uint8_t>(), ctx->tile_sobs.as<uint16_t>(),
           ctx->tile_sstart.as<uint16_t>(), ctx->flags.as<int>() + 8, ctx->ovf_raw.as<int>(), ctx->ovf_cap);
    // camera -> (tile, slot) list of the fused product: stable sort of the tiles' camera lists by camera (unused slots last)
    const size_t ntp = (size_t)TSLOTS * ctx->n_tiles;
    ENSURE(int, ctx->tp_key, ntp); ENSURE(int, ctx->tp_val, ntp); ENSURE(int, ctx->tp_key2, ntp); ENSURE(int, ctx->cam_tp, ntp);
    ENSURE(int, ctx->cam_tp_start, (size_t)n_cam + 2); ENSURE(double, ctx->tpart, 6 * ntp); ENSURE(int, ctx->cam_iota, (size_t)n_cam + 2);
    LAUNCH(k_tp_keys, cdiv((long)ntp, 256), 256, (long)ntp, (const int*)ctx->tile_cams.as<int>(), n_cam, ctx->tp_key.as<int>(), ctx->tp_val.as<int>());
    int bits_k = 1; while ((1L << bits_k) < (long)n_cam + 2 && bits_k < 31) ++bits_k;
    size_t tb = 0;
    CU(cub::DeviceRadixSort::SortPairs(nullptr, tb, ctx->tp_key.as<int>(), ctx->tp_key2.as<int>(), ctx->tp_val.as<int>(), ctx->cam_tp.as<int>(), (int)ntp, 0, bits_k, s));
    ENSURE(char, ctx->sort_tmp, tb);
    tb = ctx->sort_tmp.cap;
    CU(cub::DeviceRadixSort::SortPairs(ctx->sort_tmp.p, tb, ctx->tp_key.as<int>(), ctx->tp_key2.as<int>(), ctx->tp_val.as<int>(), ctx->cam_tp.as<int>(), (int)ntp, 0, bits_k, s));
    g_launches.fetch_add(1);
    LAUNCH(k_segment_starts, cdiv(n_cam + 1, 256), 256, (long)ntp, (const int*)ctx->tp_key2.as<int>(), n_cam, ctx->cam_tp_start.as<int>());
    LAUNCH(k_iota, cdiv(n_cam + 1, 256), 256, (long)n_cam + 1, ctx->cam_iota.as<int>());
    CU(cudaMemcpyAsync(ctx->h_flags + 8, ctx->flags.as<int>() + 8, sizeof(int), cudaMemcpyDeviceToHost, s));      // read after the sync below
  }
  ctx->grid_c = std::max(1, cdiv(n_cam, NT_C));
  long per = (n + (long)ctx->n_sm * 8 - 1) / ((long)ctx->n_sm * 8);
  long chunk_cap = 4096;
  if (const char* e = std::getenv("GLBA_CHUNK")) chunk_cap = std::max(256L, std::atol(e));    // diagnostic
  int chunk = (int)std::min<long>(chunk_cap, std::max<long>(NT_CM, ((per + NT_CM - 1) / NT_CM) * NT_CM));
  std::vector<int> cc, cb, ce, ccs((size_t)n_cam + 1, 0);
  ctx->n_free_cam = 0;
  for (int i = 0; i < n_cam; ++i) {
    ccs[i] = (int)cc.size();
    if (h_free[i]) ctx->n_free_cam++;
    for (int b = h_start[i]; b < h_start[i + 1]; b += chunk) { cc.push_back(i); cb.push_back(b); ce.push_back(std::min(h_start[i + 1], b + chunk)); }
  }
  ccs[n_cam] = (int)cc.size();
  if (ctx->owner) ctx->n_free_cam = ctx->n_free_cam_g;      // the same on every rank
  ctx->n_chunks = (int)cc.size();
  ENSURE(int, ctx->chunk_cam, cc.size()); ENSURE(int, ctx->chunk_begin, cc.size()); ENSURE(int, ctx->chunk_end, cc.size());
  ENSURE(int, ctx->cam_chunk_start, ccs.size());
  if (!cc.empty()) {
    CU(cudaMemcpyAsync(ctx->chunk_cam.p, cc.data(), sizeof(int) * cc.size(), cudaMemcpyHostToDevice, s));
    CU(cudaMemcpyAsync(ctx->chunk_begin.p, cb.data(), sizeof(int) * cb.size(), cudaMemcpyHostToDevice, s));
    CU(cudaMemcpyAsync(ctx->chunk_end.p, ce.data(), sizeof(int) * ce.size(), cudaMemcpyHostToDevice, s));
  }
  CU(cudaMemcpyAsync(ctx->cam_chunk_start.p, ccs.data(), sizeof(int) * ccs.size(), cudaMemcpyHostToDevice, s));
  CU(cudaStreamSynchronize(s));   // the host vectors go out of scope
  // fused product only when every observation of the map found a camera slot in its tile (sharded: on every rank alike is
  // not required, the two forms compute the same sums)
  ctx->n_ovf = ctx->use_pipe ? ctx->h_flags[8] : 0;
  ctx->use_fused = ctx->use_pipe && ctx->env_fused && ctx->n_ovf <= ctx->ovf_cap;
  if (ctx->use_fused && ctx->n_ovf > 0) {
    // observations whose camera found no slot in their tile: rows of ovf_c in ascending observation order, and per camera the
    // list of its rows (both by radix sort: the order the tiles appended them in is irrelevant)
    const int m = ctx->n_ovf;
    ENSURE(int, ctx->ovf_k, m); ENSURE(double, ctx->ovf_c, 6 * (size_t)m); ENSURE(int, ctx->ovf_key, m); ENSURE(int, ctx->ovf_key2, m);
    ENSURE(int, ctx->ovf_val, m); ENSURE(int, ctx->cam_ov, m); ENSURE(int, ctx->cam_ov_start, (size_t)n_cam + 2);
    size_t tb = 0;
    CU(cub::DeviceRadixSort::SortKeys(nullptr, tb, ctx->ovf_raw.as<int>(), ctx->ovf_k.as<int>(), m, 0, 32, s));
    ENSURE(char, ctx->sort_tmp, tb);
    tb = ctx->sort_tmp.cap;
    CU(cub::DeviceRadixSort::SortKeys(ctx->sort_tmp.p, tb, ctx->ovf_raw.as<int>(), ctx->ovf_k.as<int>(), m, 0, 32, s));
    LAUNCH(k_ovf_keys, cdiv(m, 256), 256, m, (const int*)ctx->ovf_k.as<int>(), (const int*)ctx->pm_cam.as<int>(), ctx->ovf_key.as<int>(), ctx->ovf_val.as<int>());
    tb = 0;
    CU(cub::DeviceRadixSort::SortPairs(nullptr, tb, ctx->ovf_key.as<int>(), ctx->ovf_key2.as<int>(), ctx->ovf_val.as<int>(), ctx->cam_ov.as<int>(), m, 0, 32, s));
    ENSURE(char, ctx->sort_tmp, tb);
    tb = ctx->sort_tmp.cap;
    CU(cub::DeviceRadixSort::SortPairs(ctx->sort_tmp.p, tb, ctx->ovf_key.as<int>(), ctx->ovf_key2.as<int>(), ctx->ovf_val.as<int>(), ctx->cam_ov.as<int>(), m, 0, 32, s));
    g_launches.fetch_add(2);
    LAUNCH(k_segment_starts, cdiv(n_cam + 1, 256), 256, (long)m, (const int*)ctx->ovf_key2.as<int>(), n_cam, ctx->cam_ov_start.as<int>());
  }
  // work buffers
  const int grid_pm = cdiv(n_pt, NT_PM);
  ENSURE(double4, ctx->rec_pm, n); ENSURE(double4, ctx->rec_cm, n);
  ENSURE(double, ctx->Craw, 9 * (size_t)n_pt); ENSURE(double4, ctx->sp4, n_pt); ENSURE(double4, ctx->lam4, n_pt);
  ENSURE(double, ctx->cinv, (size_t)PBLK * n_pt); ENSURE(double4, ctx->u4, n_pt);
  ENSURE(double, ctx->part_pm, 5 * (size_t)std::max(std::max(grid_pm, ctx->n_tiles), 1));
  ENSURE(double, ctx->xtab, (size_t)XTAB * n_cam); ENSURE(double, ctx->partA, ctx->grid_c); ENSURE(double, ctx->partB, ctx->grid_c);
  ENSURE(double, ctx->partc, 8 * (size_t)ctx->grid_c); ENSURE(unsigned, ctx->counters, 8); ENSURE(double, ctx->part_pm2, 5 * 64);
  CU(cudaMemsetAsync(ctx->counters.p, 0, 8 * sizeof(unsigned), s));
  ctx->timing = (n >= 200000) || ctx->env_timing; ENSURE(double, ctx->part_cm, 27 * (size_t)std::max(ctx->n_chunks, 1)); ENSURE(double, ctx->part_cm2, 27 * (size_t)std::max(ctx->n_chunks, 1));
  ENSURE(double, ctx->acc27, 54 * (size_t)n_cam + NSCAL);     // [Schur sums 27C | Hessian sums 27C | scalars]: contiguous for one all-reduce
  ctx->d_accB = ctx->acc27.as<double>(); ctx->d_accA = ctx->d_accB + 27 * (size_t)n_cam; ctx->d_scal = ctx->d_accA + 27 * (size_t)n_cam; ENSURE(double, ctx->yhat, 6 * (size_t)n_cam);
  ENSURE(double, ctx->Bc, 36 * (size_t)n_cam); ENSURE(double, ctx->gc, 6 * (size_t)n_cam); ENSURE(double, ctx->sc, 6 * (size_t)n_cam);
  ENSURE(double, ctx->lamc, 6 * (size_t)n_cam); ENSURE(double, ctx->Md, 36 * (size_t)n_cam); ENSURE(double, ctx->Minv, 36 * (size_t)n_cam);
  ENSURE(double, ctx->rhs, 6 * (size_t)n_cam); ENSURE(double, ctx->cg_x, 6 * (size_t)n_cam); ENSURE(double, ctx->cg_r, 6 * (size_t)n_cam);
  ENSURE(double, ctx->cg_p, 6 * (size_t)n_cam); ENSURE(double, ctx->cg_q, 6 * (size_t)n_cam); ENSURE(double, ctx->pg, 6 * (size_t)n_cam);
  ENSURE(double, ctx->yg, 6 * (size_t)n_cam); ENSURE(CgState, ctx->cgst, 1);
  if (n_cam <= DN_MAXCAM && n_cam > 0) {
    ctx->dn_grid = std::max(1, std::min(ctx->n_sm, cdiv(n_pt, 16)));   // spread over all SMs; a CTA stages up to DN_TP points per barrier round
    ctx->dn_ppc = cdiv(n_pt, ctx->dn_grid);
    const size_t len = (size_t)(n_cam * (n_cam + 1) / 2) * 36 + 6 * (size_t)n_cam;
    ENSURE(double, ctx->dn_part, len * ctx->dn_grid); ENSURE(double, ctx->dn_red, len); ENSURE(double, ctx->dn_full, (size_t)36 * n_cam * n_cam + 6 * (size_t)n_cam);
  }
  CU(cudaMemsetAsync(ctx->yhat.p, 0, sizeof(double) * 6 * n_cam, s));
  CU(cudaMemsetAsync(ctx->acc27.p, 0, sizeof(double) * (54 * (size_t)n_cam + NSCAL), s));
  // state (after the parameter upload) and measurements (after theirs, the last to arrive)
  CU(cudaMemcpyAsync(ctx->cam[0].p, d_cam, sizeof(double) * 6 * n_cam, cudaMemcpyDeviceToDevice, s));
  CU(cudaMemcpyAsync(ctx->cam0.p, ctx->cam[0].p, sizeof(double) * 6 * n_cam, cudaMemcpyDeviceToDevice, s));
  if (n_pt) LAUNCH(k_pack_pt, cdiv(n_pt, 256), 256, n_pt, d_pt, d_info, n2o, ctx->pt4[0].as<double4>());
  CU(cudaMemcpyAsync(ctx->pt40.p, ctx->pt4[0].p, sizeof(double4) * n_pt, cudaMemcpyDeviceToDevice, s));
  if (n_cam) LAUNCH(k_cam_prep, cdiv(n_cam, 128), 128, n_cam, (const double*)ctx->cam[0].as<double>(), ctx->camtab[0].as<double>(), ctx->mode);
  if (staged) CU(cudaStreamWaitEvent(s, ctx->ev_copy[2], 0));
  if (n > 0) LAUNCH(k_gather_uv, gb, 256, n, ctx->sorted_input ? (const int*)nullptr : (const int*)ctx->pm2orig.as<int>(), (const int*)ctx->cm2pm.as<int>(),
                    d_u, d_v, ctx->pm_uv.as<double2>(), ctx->cm_uv.as<double2>());
  mark(ctx, -1);
  ctx->loaded = true;
  return GLBA_OK;
}

// The uploads read the caller's buffers: whatever happened above, they have finished when this returns.
int load_problem(glba_ctx* ctx, const glba_problem* p) {
  const int st = load_problem_impl(ctx, p);
  if (ctx->copies_in_flight) {
    ctx->copies_in_flight = false;
    const cudaError_t e = cudaStreamSynchronize(ctx->copy_stream);
    if (e != cudaSuccess && st == GLBA_OK) return fail(ctx, GLBA_E_CUDA, "upload: %s", cudaGetErrorString(e));
  }
  return st;
}

PmArgs pm_args(glba_ctx* ctx, const glba_options* o) {
  PmArgs A;
  A.n_pt = ctx->n_pt; A.pt_start = ctx->pt_start.as<int>(); A.pm_cam = ctx->pm_cam.as<int>(); A.pm_uv = ctx->pm_uv.as<double2>();
  A.pm2cm = ctx->pm2cm.as<int>(); A.pt_free = ctx->pt_free.as<uint8_t>(); A.K = ctx->K;
  A.loss = LossP{o->loss, o->loss_scale};
  return A;
}
CmArgs cm_args(glba_ctx* ctx) {
  CmArgs A;
  A.chunk_cam = ctx->chunk_cam.as<int>(); A.chunk_begin = ctx->chunk_begin.as<int>(); A.chunk_end = ctx->chunk_end.as<int>();
  A.cm_pt = ctx->cm_pt.as<int>(); A.cm_uv = ctx->cm_uv.as<double2>(); A.cam_free = ctx->cam_free.as<uint8_t>(); A.K = ctx->K;
  return A;
}

TileMeta tile_meta(glba_ctx* ctx) {
  return TileMeta{ctx->tile_desc.as<int4>(), ctx->tile_cams.as<int>(), ctx->pm_slot.as<uint8_t>(), ctx->pm_cam.as<int>(), ctx->pm_pt.as<int>(), ctx->n_tiles,
                  ctx->tile_sobs.as<uint16_t>(), ctx->tile_sstart.as<uint16_t>(), ctx->ovf_k.as<int>(), ctx->ovf_c.as<double>(), ctx->use_fused ? ctx->n_ovf : 0};
}
int pipe_grid(const glba_ctx* ctx, int occ) { return std::max(1, std::min(ctx->n_tiles, ctx->n_sm * occ)); }
TileArgs tile_args(glba_ctx* ctx) { return TileArgs{ctx->tile_pt.as<int>(), ctx->pm_pt.as<int>(), ctx->tile_cmin.as<int>(), ctx->n_cam}; }

int reduce_pm_partials(glba_ctx* ctx, int rows, const int* slots, int max_col) {
  ReduceMap M{}; M.n = 5;
  for (int q = 0; q < 5; ++q) { M.slot[q] = slots[q]; M.is_max[q] = (q == max_col); }
  LAUNCH(k_reduce_partials, 1, NT_CAM, rows, 5, (const double*)ctx->part_pm.as<double>(), M, ctx->d_scal, (const LmCtl*)nullptr, (int)GATE_ALWAYS);
  return GLBA_OK;
}

constexpr int kInKernelReduceMaxTiles = 256;   // above this the per-tile partials are folded by k_reduce_rows (64 CTAs)
RedArgs red_args(glba_ctx* ctx, int counter, const int* slots, bool in_kernel = true, const LmCtl* ctl = nullptr, int gate = GATE_ALWAYS,
                 const LmHook* hook = nullptr) {
  RedArgs R{}; R.counter = in_kernel ? ctx->counters.as<unsigned>() + counter : nullptr; R.scal = ctx->d_scal;
  for (int q = 0; q < 5; ++q) R.slots[q] = slots[q];
  R.ctl = ctl; R.gate = gate;
  if (hook) R.hook = *hook;
  return R;
}
const int kLinSlots[5] = {S_COST, S_XN2_P, S_BAD, S_NOTPD_P, S_GMAX_P};
const int kStepSlots[5] = {S_COST_C, S_YN2_P, S_YG_P, S_YLY_P, S_BAD_C};

// point-major half of a linearisation, INCLUDING the reduction of its scalars into d_scal
int launch_linearize_points(glba_ctx* ctx, const glba_options* o, int first, double radius, const LmCtl* ctl = nullptr) {
  const int c = ctx->cur;
  if (ctx->use_pipe) {
    LAUNCH_SMEM(k_lin_pipe, pipe_grid(ctx, ctx->occ_lin), P_NT, sizeof(LinSmem), pm_args(ctx, o), tile_meta(ctx), (const double4*)ctx->pt4[c].as<double4>(),
        (const double*)ctx->camtab[c].as<double>(), ctx->rec_pm.as<double4>(), ctx->rec_cm.as<double4>(), ctx->Craw.as<double>(), ctx->sp4.as<double4>(),
        ctx->lam4.as<double4>(), ctx->cinv.as<double>(), ctx->u0p.as<double4>(), first, o->jacobi_scaling, o->min_lm_diagonal, o->max_lm_diagonal,
        1.0 / radius, ctx->part_pm.as<double>(), red_args(ctx, 4, kLinSlots, true, ctl, GATE_ACCEPTED));
  } else if (ctx->use_tiles) {
    const size_t smem = (size_t)8 * NT_T * ctx->opt * sizeof(double);
    const RedArgs RA = red_args(ctx, 4, kLinSlots, ctx->n_tiles <= kInKernelReduceMaxTiles, ctl, GATE_ACCEPTED);
#define LIN_TILE_ARGS pm_args(ctx, o), tile_args(ctx), (const double4*)ctx->pt4[c].as<double4>(), (const double*)ctx->camtab[c].as<double>(), \
    ctx->rec_pm.as<double4>(), ctx->rec_cm.as<double4>(), ctx->Craw.as<double>(), ctx->sp4.as<double4>(), ctx->lam4.as<double4>(), \
    ctx->cinv.as<double>(), ctx->u0p.as<double4>(), first, o->jacobi_scaling, o->min_lm_diagonal, o->max_lm_diagonal, 1.0 / radius, ctx->part_pm.as<double>(), RA
    if (ctx->opt == OPT_LARGE) LAUNCH_SMEM(k_linearize_tile<OPT_LARGE>, ctx->n_tiles, NT_T, smem, LIN_TILE_ARGS);
    else LAUNCH_SMEM(k_linearize_tile<OPT_SMALL>, ctx->n_tiles, NT_T, smem, LIN_TILE_ARGS);
    if (ctx->n_tiles > kInKernelReduceMaxTiles)
      LAUNCH(k_reduce_rows<4>, 64, NT_T, (const double*)ctx->part_pm.as<double>(), ctx->n_tiles, ctx->part_pm2.as<double>(), red_args(ctx, 4, kLinSlots, true, ctl, GATE_ACCEPTED));
  } else {
    LAUNCH(k_linearize_pm, cdiv(ctx->n_pt, NT_PM), NT_PM, pm_args(ctx, o), (const double4*)ctx->pt4[c].as<double4>(),
           (const double*)ctx->camtab[c].as<double>(), ctx->rec_pm.as<double4>(), ctx->rec_cm.as<double4>(), ctx->Craw.as<double>(),
           ctx->sp4.as<double4>(), ctx->lam4.as<double4>(), ctx->cinv.as<double>(), ctx->u0p.as<double4>(), first, o->jacobi_scaling, o->min_lm_diagonal,
           o->max_lm_diagonal, 1.0 / radius, ctx->part_pm.as<double>());
    reduce_pm_partials(ctx, cdiv(ctx->n_pt, NT_PM), kLinSlots, 4);
  }
  return GLBA_OK;
}

void launch_point_pass0(glba_ctx* ctx, const glba_options* o, const CgState* cg, int li) {
  const int c = ctx->cur;
  if (ctx->use_pipe) {
    LAUNCH_SMEM(k_pt_pipe<0>, pipe_grid(ctx, ctx->occ_pt0), P_NT, sizeof(PtSmem<0>), pm_args(ctx, o), tile_meta(ctx), (const double4*)ctx->rec_pm.as<double4>(),
        (const double*)ctx->camtab[c].as<double>(), (const double*)ctx->xtab.as<double>(), (const double*)ctx->cinv.as<double>(),
        (const double4*)ctx->u0p.as<double4>(), ctx->u4.as<double4>(), cg, li, (const double4*)nullptr, (double4*)nullptr, (const double*)nullptr,
        (const double*)nullptr, (const double4*)nullptr, 0.0, (double*)nullptr, RedArgs{});
  } else if (ctx->use_tiles) {
#define PT0_ARGS pm_args(ctx, o), tile_args(ctx), (const double4*)ctx->rec_pm.as<double4>(), (const double*)ctx->camtab[c].as<double>(), \
    (const double*)ctx->xtab.as<double>(), (const double*)ctx->cinv.as<double>(), (const double4*)ctx->u0p.as<double4>(), ctx->u4.as<double4>(), cg, li, (const double4*)nullptr, \
    (double4*)nullptr, (const double*)nullptr, (const double*)nullptr, (const double4*)nullptr, 0.0, (double*)nullptr, RedArgs{}
    if (ctx->opt == OPT_LARGE) LAUNCH((k_point_tile<0, OPT_LARGE>), ctx->n_tiles, NT_T, PT0_ARGS);
    else LAUNCH((k_point_tile<0, OPT_SMALL>), ctx->n_tiles, NT_T, PT0_ARGS);
  }
  else
    LAUNCH(k_point_pass<0>, cdiv(ctx->n_pt, NT_PM), NT_PM, pm_args(ctx, o), (const double4*)ctx->rec_pm.as<double4>(),
           (const double*)ctx->camtab[c].as<double>(), (const double*)ctx->xtab.as<double>(), (const double*)ctx->cinv.as<double>(), (const double4*)ctx->u0p.as<double4>(),
           ctx->u4.as<double4>(), cg, li, (const double4*)nullptr, (double4*)nullptr, (const double*)nullptr, (const double*)nullptr,
           (const double4*)nullptr, 0.0, (double*)nullptr);
}

// both halves of the implicit product in one pass over the point-major records (k_pt_pipe<2>): per-tile, per-camera-slot sums
void launch_spmv_fused(glba_ctx* ctx, const glba_options* o, const CgState* cg, int li) {
  const int c = ctx->cur;
  LAUNCH_SMEM(k_pt_pipe<2>, pipe_grid(ctx, ctx->occ_pt2), P_NT, sizeof(PtSmem<2>), pm_args(ctx, o), tile_meta(ctx), (const double4*)ctx->rec_pm.as<double4>(),
      (const double*)ctx->camtab[c].as<double>(), (const double*)ctx->xtab.as<double>(), (const double*)ctx->cinv.as<double>(),
      (const double4*)ctx->u0p.as<double4>(), ctx->u4.as<double4>(), cg, li, (const double4*)nullptr, (double4*)nullptr, (const double*)nullptr,
      (const double*)nullptr, (const double4*)nullptr, 0.0, ctx->tpart.as<double>(), RedArgs{});
}

void launch_point_pass1(glba_ctx* ctx, const glba_options* o, double radius, const LmCtl* ctl = nullptr, const LmHook* hook = nullptr) {
  const int c = ctx->cur, d = c ^ 1;
  if (ctx->use_pipe) {
    LAUNCH_SMEM(k_pt_pipe<1>, pipe_grid(ctx, ctx->occ_pt1), P_NT, sizeof(PtSmem<1>), pm_args(ctx, o), tile_meta(ctx), (const double4*)ctx->rec_pm.as<double4>(),
        (const double*)ctx->camtab[c].as<double>(), (const double*)ctx->xtab.as<double>(), (const double*)ctx->cinv.as<double>(),
        (const double4*)ctx->u0p.as<double4>(), (double4*)nullptr, (const CgState*)nullptr, 0, (const double4*)ctx->pt4[c].as<double4>(),
        ctx->pt4[d].as<double4>(), (const double*)ctx->camtab[d].as<double>(), (const double*)ctx->Craw.as<double>(),
        (const double4*)ctx->lam4.as<double4>(), 1.0 / radius, ctx->part_pm.as<double>(), red_args(ctx, 5, kStepSlots, true, ctl, GATE_ALWAYS, hook));
  } else if (ctx->use_tiles) {
    const RedArgs RA = red_args(ctx, 5, kStepSlots, ctx->n_tiles <= kInKernelReduceMaxTiles, ctl, GATE_ALWAYS, hook);
#define PT1_ARGS pm_args(ctx, o), tile_args(ctx), (const double4*)ctx->rec_pm.as<double4>(), (const double*)ctx->camtab[c].as<double>(), \
    (const double*)ctx->xtab.as<double>(), (const double*)ctx->cinv.as<double>(), (const double4*)ctx->u0p.as<double4>(), (double4*)nullptr, (const CgState*)nullptr, 0, \
    (const double4*)ctx->pt4[c].as<double4>(), ctx->pt4[d].as<double4>(), (const double*)ctx->camtab[d].as<double>(), \
    (const double*)ctx->Craw.as<double>(), (const double4*)ctx->lam4.as<double4>(), 1.0 / radius, ctx->part_pm.as<double>(), RA
    if (ctx->opt == OPT_LARGE) LAUNCH((k_point_tile<1, OPT_LARGE>), ctx->n_tiles, NT_T, PT1_ARGS);
    else LAUNCH((k_point_tile<1, OPT_SMALL>), ctx->n_tiles, NT_T, PT1_ARGS);
    if (ctx->n_tiles > kInKernelReduceMaxTiles)
      LAUNCH(k_reduce_rows<-1>, 64, NT_T, (const double*)ctx->part_pm.as<double>(), ctx->n_tiles, ctx->part_pm2.as<double>(), red_args(ctx, 5, kStepSlots, true, ctl, GATE_ALWAYS, hook));
  } else {
    LAUNCH(k_point_pass<1>, cdiv(ctx->n_pt, NT_PM), NT_PM, pm_args(ctx, o), (const double4*)ctx->rec_pm.as<double4>(),
           (const double*)ctx->camtab[c].as<double>(), (const double*)ctx->xtab.as<double>(), (const double*)ctx->cinv.as<double>(), (const double4*)ctx->u0p.as<double4>(),
           (double4*)nullptr, (const CgState*)nullptr, 0, (const double4*)ctx->pt4[c].as<double4>(), ctx->pt4[d].as<double4>(),
           (const double*)ctx->camtab[d].as<double>(), (const double*)ctx->Craw.as<double>(), (const double4*)ctx->lam4.as<double4>(),
           1.0 / radius, ctx->part_pm.as<double>());
    reduce_pm_partials(ctx, cdiv(ctx->n_pt, NT_PM), kStepSlots, -1);
  }
}

#define CAM_LIN_FIN_ARGS n_cam, (const uint8_t*)ctx->cam_free.as<uint8_t>(), (const double*)ctx->cam[c].as<double>(), \
    (const double*)ctx->camtab[c].as<double>(), (const double*)ctx->d_accA, (const int*)ctx->cam_chunk_start.as<int>(), (const double*)ctx->part_cm.as<double>(), \
    ctx->Bc.as<double>(), ctx->gc.as<double>(), ctx->sc.as<double>(), ctx->lamc.as<double>(), first, o->jacobi_scaling, o->min_lm_diagonal, \
    o->max_lm_diagonal, ctx->partc.as<double>(), ctx->counters.as<unsigned>() + 0, ctx->d_scal, ctl, hk, ctx->owner ? (const uint8_t*)ctx->cam_owned.as<uint8_t>() : (const uint8_t*)nullptr
void launch_cam_lin_fin(glba_ctx* ctx, const glba_options* o, int first, const LmCtl* ctl = nullptr, const LmHook* hook = nullptr) {
  const LmHook hk = hook ? *hook : LmHook{};
  const int c = ctx->cur, n_cam = ctx->n_cam;
  if (ctx->world > 1) LAUNCH(k_cam_lin_fin<false>, ctx->grid_c, NT_C, CAM_LIN_FIN_ARGS);
  else LAUNCH(k_cam_lin_fin<true>, ctx->grid_c, NT_C, CAM_LIN_FIN_ARGS);
}
#define CAM_SCHUR_FIN_ARGS n_cam, (const uint8_t*)ctx->cam_free.as<uint8_t>(), (const double*)ctx->camtab[c].as<double>(), (const double*)ctx->d_accB, \
    (const int*)ctx->cam_chunk_start.as<int>(), (const double*)part27, (const double*)ctx->Bc.as<double>(), (const double*)ctx->gc.as<double>(), \
    (const double*)ctx->lamc.as<double>(), 1.0 / radius, ctx->Md.as<double>(), ctx->Minv.as<double>(), ctx->rhs.as<double>(), ctx->partc.as<double>(), \
    ctx->counters.as<unsigned>() + 1, ctx->d_scal, ctx->owner ? (const uint8_t*)ctx->cam_owned.as<uint8_t>() : (const uint8_t*)nullptr
void launch_cam_schur_fin(glba_ctx* ctx, double radius, const double* part27) {
  const int c = ctx->cur, n_cam = ctx->n_cam;
  if (ctx->world > 1) LAUNCH(k_cam_schur_fin<false>, ctx->grid_c, NT_C, CAM_SCHUR_FIN_ARGS);
  else LAUNCH(k_cam_schur_fin<true>, ctx->grid_c, NT_C, CAM_SCHUR_FIN_ARGS);
}

// Linearise (K_A + K_B blocks) and, if with_schur, the Schur pieces for `radius` in the same pass.  The Schur kernel
// needs only point-local data, so when the map is sharded both sets of per-camera partial sums and the scalars travel
// in ONE all-reduce, issued after both heavy kernels.
// ---- peer exchange (opt-in, GLBA_P2P=1) --------------------------------------------------------
// Measured on C4 (profiles/r02_p2p_ab.log): 2 GPUs 22 us per exchange against 15 us for ncclAllReduce (step 0.243 vs 0.238 ms);
// 8 GPUs step 0.1352 vs 0.1370 ms.  Two extra launches and two system-scope fences cost what the posted NVLink stores save,
// so NCCL stays the default; the form that would win pushes from the chunk-sum kernel and pops inside the finalisation.
// Best effort at glba_create: any failure (no IPC, no peer access, more ranks than the table holds) leaves p2p_ok = false on
// EVERY rank (agreed through one NCCL all-reduce) and the compact exchange keeps using ncclAllReduce.
constexpr size_t kP2pCap = (size_t)1 << 17;       // doubles per (slot, source rank) area: 1 MB, 2 400 shared cameras
// one allocation per rank: areas [slot 2][source rank 16][kP2pCap doubles] | flags [slot 2][source rank 16] | error word
double* p2p_area(void* base, unsigned slot, int src) { return reinterpret_cast<double*>(base) + ((size_t)slot * P2P_MAX_WORLD + src) * kP2pCap; }
unsigned* p2p_flags(void* base, unsigned slot) { return reinterpret_cast<unsigned*>(reinterpret_cast<double*>(base) + 2 * P2P_MAX_WORLD * kP2pCap) + slot * P2P_MAX_WORLD; }
int* p2p_err(void* base) { return reinterpret_cast<int*>(p2p_flags(base, 0) + 2 * P2P_MAX_WORLD); }
int setup_p2p(glba_ctx* ctx) {
  ctx->p2p_ok = false;
  if (ctx->world <= 1) return GLBA_OK;
  int ok = (ctx->env_p2p && ctx->world <= P2P_MAX_WORLD && g_nccl.AllGather != nullptr) ? 1 : 0;
  const size_t bytes = 2 * P2P_MAX_WORLD * kP2pCap * sizeof(double) + 1024;
  cudaIpcMemHandle_t mine;
  std::memset(&mine, 0, sizeof(mine));
  if (ok && cudaMalloc(&ctx->p2p_local, bytes) != cudaSuccess) { ctx->p2p_local = nullptr; ok = 0; (void)cudaGetLastError(); }
  if (ok && (cudaMemset(ctx->p2p_local, 0, bytes) != cudaSuccess || cudaIpcGetMemHandle(&mine, ctx->p2p_local) != cudaSuccess)) { ok = 0; (void)cudaGetLastError(); }
  // handles of all ranks (every rank takes part in the collectives whatever its own outcome)
  Buf hb;
  ENSURE(char, hb, sizeof(mine) * ((size_t)ctx->world + 1) + 64);
  char* d_mine = hb.as<char>();
  char* d_all = d_mine + sizeof(mine);
  std::vector<cudaIpcMemHandle_t> all(ctx->world);
  int rc = GLBA_OK;
  if (g_nccl.AllGather != nullptr) {
    if (cudaMemcpyAsync(d_mine, &mine, sizeof(mine), cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess ||
        g_nccl.AllGather(d_mine, d_all, sizeof(mine), 1 /* ncclUint8 */, ctx->comm, ctx->stream) != 0 ||
        cudaMemcpyAsync(all.data(), d_all, sizeof(mine) * ctx->world, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess ||
        cudaStreamSynchronize(ctx->stream) != cudaSuccess) rc = GLBA_E_NCCL;
  }
  if (rc == GLBA_OK && ok) {
    for (int r = 0; r < ctx->world && ok; ++r) {
      if (r == ctx->rank) { ctx->p2p_peer[r] = ctx->p2p_local; continue; }
      if (cudaIpcOpenMemHandle(&ctx->p2p_peer[r], all[r], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { ctx->p2p_peer[r] = nullptr; ok = 0; (void)cudaGetLastError(); }
    }
  }
  // one rank's failure is everybody's
  if (rc == GLBA_OK) {
    int* d_ok = reinterpret_cast<int*>(d_mine);
    if (cudaMemcpyAsync(d_ok, &ok, sizeof(int), cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess ||
        g_nccl.AllReduce(d_ok, d_ok, 1, kNcclInt32, 3 /* ncclMin */, ctx->comm, ctx->stream) != 0 ||
        cudaMemcpyAsync(&ok, d_ok, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess ||
        cudaStreamSynchronize(ctx->stream) != cudaSuccess) rc = GLBA_E_NCCL;
  }
  release(hb);
  if (rc != GLBA_OK) return fail(ctx, rc, "peer-exchange setup: collective failed");
  ctx->p2p_ok = ok != 0;
  ctx->p2p_cap = kP2pCap;
  ctx->p2p_seq = 0;
  return GLBA_OK;
}
void teardown_p2p(glba_ctx* ctx) {
  for (int r = 0; r < 16; ++r)
    if (ctx->p2p_peer[r] && ctx->p2p_peer[r] != ctx->p2p_local) cudaIpcCloseMemHandle(ctx->p2p_peer[r]);
  if (ctx->p2p_local) cudaFree(ctx->p2p_local);
  ctx->p2p_local = nullptr; ctx->p2p_ok = false;
}

// The compact exchange of the owner-computes layout: rows of the shared cameras (A half, optionally B half) + n_tail scalars,
// summed over the ranks, back into accA / accB / scal.  Peer memory when it is set up and the payload fits, NCCL otherwise.
int owner_exchange(glba_ctx* ctx, bool with_a_out, bool with_b, int n_tail) {
  const int n_cam = ctx->n_cam;
  const int gx = std::max(1, cdiv((long)std::max(n_cam, 1) * 54, 256));
  const size_t len = 54 * (size_t)ctx->n_shared + (size_t)n_tail;
  const double* accB_in = with_b ? (const double*)ctx->d_accB : (const double*)nullptr;
  double* accA_out = with_a_out ? ctx->d_accA : (double*)nullptr;
  double* accB_out = with_b ? ctx->d_accB : (double*)nullptr;
  if (ctx->p2p_ok && len <= ctx->p2p_cap) {
    const unsigned seq = ++ctx->p2p_seq;
    const unsigned slot = seq & 1u;
    PeerTable P;
    for (int r = 0; r < P2P_MAX_WORLD; ++r) {
      void* base = r < ctx->world ? ctx->p2p_peer[r] : ctx->p2p_local;
      P.area[r] = p2p_area(base, slot, ctx->rank);
      P.flag[r] = p2p_flags(base, slot) + ctx->rank;
    }
    const int gd = std::max(1, cdiv((long)len, 256));
    LAUNCH(k_xch_push, gd, 256, (const int*)ctx->sh2loc.as<int>(), (const double*)ctx->d_accA, accB_in, ctx->n_shared, (const double*)ctx->d_scal, n_tail, P, ctx->world,
           ctx->counters.as<unsigned>() + 4, seq);
    LAUNCH(k_xch_pop, gd, 256, (const int*)ctx->sh2loc.as<int>(), (const double*)p2p_area(ctx->p2p_local, slot, 0), (const unsigned*)p2p_flags(ctx->p2p_local, slot),
           kP2pCap, ctx->world, seq, ctx->n_shared, accA_out, accB_out, ctx->d_scal, n_tail, p2p_err(ctx->p2p_local));
    return GLBA_OK;
  }
  LAUNCH(k_xch_pack, gx, 256, n_cam, (const int*)ctx->cam_shared.as<int>(), (const double*)ctx->d_accA, accB_in, ctx->n_shared, ctx->xsend.as<double>(),
         (const double*)ctx->d_scal, n_tail);
  const int r = g_nccl.AllReduce(ctx->xsend.p, ctx->xrecv.p, len, kNcclFloat64, kNcclSum, ctx->comm, ctx->stream);
  if (r != 0) return fail(ctx, GLBA_E_NCCL, "ncclAllReduce (compact exchange): %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "error");
  LAUNCH(k_xch_unpack, gx, 256, n_cam, (const int*)ctx->cam_shared.as<int>(), (const double*)ctx->xrecv.as<double>(), ctx->n_shared, accA_out, accB_out, ctx->d_scal, n_tail);
  return GLBA_OK;
}

int do_linearize_impl(glba_ctx* ctx, const glba_options* o, int first, double radius, bool with_schur) {
  const int c = ctx->cur;
  const int n_cam = ctx->n_cam, n_pt = ctx->n_pt;
  with_schur = with_schur && ctx->n_free_cam > 0;
  mark(ctx, PH_LIN);
  const bool sharded = ctx->world > 1;
  if (n_pt) { const int s__ = launch_linearize_points(ctx, o, first, radius); if (s__) return s__; }
  // Large maps, linearisation WITH Schur pieces: both camera-major passes in one TMA-fed kernel that reads the records once
  // (glba_campipe.cuh); the two finalisations follow.
  const bool fused_cm = with_schur && ctx->use_pipe && ctx->env_campipe && ctx->n_chunks > 0;
  if (fused_cm) {
    LAUNCH_SMEM(ctx->env_cp_occ == 3 ? k_cam_pipe<3> : k_cam_pipe<2>, ctx->n_chunks, CP_NT, sizeof(CamSmem), cm_args(ctx), (const double4*)ctx->rec_cm.as<double4>(),
                (const double*)ctx->camtab[c].as<double>(), (const double*)ctx->cinv.as<double>(), (const double4*)ctx->u0p.as<double4>(),
                ctx->part_cm.as<double>(), ctx->part_cm2.as<double>());
    mark(ctx, PH_SCHUR);
  }
  else if (ctx->n_chunks) LAUNCH(k_linearize_cm, ctx->n_chunks, NT_HCM, cm_args(ctx), (const double4*)ctx->rec_cm.as<double4>(),
                            (const double*)ctx->camtab[c].as<double>(), ctx->part_cm.as<double>(), (const LmCtl*)nullptr);
  // Single GPU, large maps: the camera finalisation of the linearisation (B_i, g_i, scaling, LM diagonal: ~11 us of dependent
  // arithmetic on 29 CTAs) needs only k_linearize_cm's sums, so it runs on a side stream WHILE the Schur pass streams the
  // records; the Schur finalisation waits for both.
  const bool side = with_schur && !fused_cm && !sharded && n_cam > 0 && ctx->n_chunks > 0 && ctx->n_obs >= 200000;
  if (side) {
    CU(cudaEventRecord(ctx->ev_side[0], ctx->stream));
    CU(cudaStreamWaitEvent(ctx->side_stream, ctx->ev_side[0], 0));
    cudaStream_t main_stream = ctx->stream;
    ctx->stream = ctx->side_stream;
    launch_cam_lin_fin(ctx, o, first);
    ctx->stream = main_stream;
    CU(cudaEventRecord(ctx->ev_side[1], ctx->side_stream));
  }
  if (with_schur && !fused_cm) {
    mark(ctx, PH_SCHUR);
    if (ctx->n_chunks) LAUNCH(k_schur_cm, ctx->n_chunks, NT_HCM, cm_args(ctx), (const double4*)ctx->rec_cm.as<double4>(),
                              (const double*)ctx->camtab[c].as<double>(), (const double*)ctx->cinv.as<double>(), (const double4*)ctx->u0p.as<double4>(), ctx->part_cm2.as<double>());
  }
  if (sharded) {     // one payload: per-camera sums | cost, |x_p|^2, bad, notpd | per-rank gradient max slots
    mark(ctx, PH_COMM);          // in situ: includes waiting for the slowest rank
    LAUNCH(k_chunk_sum_lin, cdiv((long)n_cam * (with_schur ? 54 : 27), 256) + (n_cam ? 0 : 1), 256, n_cam, (const int*)ctx->cam_chunk_start.as<int>(),
           (const double*)ctx->part_cm.as<double>(), with_schur ? (const double*)ctx->part_cm2.as<double>() : (const double*)nullptr,
           ctx->d_accA, ctx->d_accB, ctx->rank, ctx->d_scal);
    if (ctx->owner) {
      // compact exchange: the rows of the cameras several ranks observe + the point scalars, nothing else crosses NVLink
      const int s__ = owner_exchange(ctx, true, with_schur, S_GSLOT0 + MAX_WORLD); if (s__) return s__;
    }
    else if (with_schur) AR(ctx->d_accB, 54 * (size_t)n_cam + S_GSLOT0 + MAX_WORLD, kNcclSum);
    else AR(ctx->d_accA, 27 * (size_t)n_cam + S_GSLOT0 + MAX_WORLD, kNcclSum);
    mark(ctx, with_schur ? PH_SCHUR : PH_LIN);
  }
  if (fused_cm && !sharded && n_cam) {
    // both sets of chunk sums arrived together: one launch finishes a camera twice
    LAUNCH(k_cam_fin_both<true>, ctx->grid_c, NT_C, n_cam, (const uint8_t*)ctx->cam_free.as<uint8_t>(), (const double*)ctx->cam[c].as<double>(),
           (const double*)ctx->camtab[c].as<double>(), (const double*)ctx->d_accA, (const double*)ctx->d_accB, (const int*)ctx->cam_chunk_start.as<int>(),
           (const double*)ctx->part_cm.as<double>(), (const double*)ctx->part_cm2.as<double>(), ctx->Bc.as<double>(), ctx->gc.as<double>(), ctx->sc.as<double>(),
           ctx->lamc.as<double>(), first, o->jacobi_scaling, o->min_lm_diagonal, o->max_lm_diagonal, 1.0 / radius, ctx->Md.as<double>(), ctx->Minv.as<double>(),
           ctx->rhs.as<double>(), ctx->partc.as<double>(), ctx->counters.as<unsigned>() + 0, ctx->counters.as<unsigned>() + 1, ctx->d_scal, (const uint8_t*)nullptr);
  } else {
    if (side) CU(cudaStreamWaitEvent(ctx->stream, ctx->ev_side[1], 0));
    else if (n_cam) launch_cam_lin_fin(ctx, o, first);
    if (with_schur && n_cam) launch_cam_schur_fin(ctx, radius, ctx->part_cm2.as<double>());
  }
  ctx->schur_fresh = with_schur;
  mark(ctx, -1);
  CHECK_LAUNCHES();
  return GLBA_OK;
}
int do_linearize_schur(glba_ctx* ctx, const glba_options* o, int first, double radius) { return do_linearize_impl(ctx, o, first, radius, true); }

// re-damp point blocks for a new radius (after a rejected / invalid step)
int do_redamp(glba_ctx* ctx, double radius, const LmCtl* ctl = nullptr) {
  const int n_pt = ctx->n_pt;
  if (!n_pt) return GLBA_OK;
  const int grid_pm = cdiv(n_pt, NT_PM);
  mark(ctx, PH_SCHUR);
  LAUNCH(k_point_damp, grid_pm, NT_PM, n_pt, (const uint8_t*)ctx->pt_free.as<uint8_t>(), (const double*)ctx->Craw.as<double>(),
         (const double4*)ctx->lam4.as<double4>(), ctx->cinv.as<double>(), ctx->u0p.as<double4>(), 1.0 / radius, ctx->part_pm.as<double>(), ctl);
  ReduceMap M{}; M.n = 1; M.slot[0] = S_NOTPD_P; M.is_max[0] = 0;
  LAUNCH(k_reduce_partials, 1, NT_CAM, grid_pm, 1, (const double*)ctx->part_pm.as<double>(), M, ctx->d_scal, ctl, (int)GATE_REDAMP);
  if (ctx->world > 1) AR(ctx->d_scal + S_NOTPD_P, 1, kNcclSum);
  mark(ctx, -1);
  CHECK_LAUNCHES();
  return GLBA_OK;
}

// Schur complement pieces alone (after the radius changed): preconditioner blocks (= diagonal of S), reduced rhs
int do_schur(glba_ctx* ctx, double radius) {
  const int c = ctx->cur;
  const int n_cam = ctx->n_cam;
  mark(ctx, PH_SCHUR);
  if (ctx->n_chunks) LAUNCH(k_schur_cm, ctx->n_chunks, NT_HCM, cm_args(ctx), (const double4*)ctx->rec_cm.as<double4>(),
                            (const double*)ctx->camtab[c].as<double>(), (const double*)ctx->cinv.as<double>(), (const double4*)ctx->u0p.as<double4>(), ctx->part_cm2.as<double>());
  if (ctx->world > 1) {
    LAUNCH(k_chunk_sum<27>, cdiv((long)n_cam * 27, 256), 256, n_cam, (const int*)ctx->cam_chunk_start.as<int>(),
           (const double*)ctx->part_cm2.as<double>(), ctx->d_accB, (const CgState*)nullptr, 0);
    if (ctx->owner) {
      // same compact buffer as a linearisation (the A half travels too: it is unchanged and lands where it came from)
      const int s__ = owner_exchange(ctx, false, true, 0); if (s__) return s__;
    } else AR(ctx->d_accB, 27 * (size_t)n_cam, kNcclSum);
  }
  if (n_cam) launch_cam_schur_fin(ctx, radius, ctx->part_cm2.as<double>());
  ctx->schur_fresh = true;
  mark(ctx, -1);
  CHECK_LAUNCHES();
  return GLBA_OK;
}

// ---------------------------------------------------------------------------------------------
// Explicit block-sparse reduced camera matrix (glba_sparse.cuh).  ensure_explicit() builds the structure of the loaded
// problem once (device sorts; two small read-backs for the counts); launch_schur_pairs() fills the blocks for the current
// linearisation and damping.  Used on a single GPU when the structure is moderate: no duplicate (point, camera)
// observations, at most 16 instances per observation (very long tracks make the assembly dearer than the products it
// saves) and at most 1 GiB of blocks.  Everything else keeps the matrix-free product.
// ---------------------------------------------------------------------------------------------
int ensure_explicit(glba_ctx* ctx, const glba_options* o) {
  (void)o;
  if (ctx->sp_tried) return GLBA_OK;
  ctx->sp_tried = true; ctx->use_explicit = false;
  const int n_cam = ctx->n_cam, n_pt = ctx->n_pt;
  const bool sharded = ctx->world > 1;
  // every condition up to the veto below is the same on every rank of a sharded run (collectives follow)
  if (!ctx->env_explicit || (sharded && !ctx->owner) || ctx->has_dup || ctx->n_free_cam < 2 || n_pt == 0 || ctx->n_obs == 0) return GLBA_OK;
  const int n_cam_g = sharded ? ctx->n_cam_g : n_cam;            // cameras of the whole map: keys and rows are global
  const long long K2 = (long long)n_cam_g * n_cam_g;
  if (sharded && K2 > (1LL << 28)) return GLBA_OK;               // the union of the ranks' blocks goes through a presence table of n_cam^2 ints
  int coop = 0, occ = 0;
  CU(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, ctx->device));
  CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_cg_bsr<false>, NT_CGP, 0));
  if (!coop || occ < 1) return GLBA_OK;                          // the PCG runs in one cooperative launch: every CTA must be resident
  cudaStream_t s = ctx->stream;
  // diagnostic (GLBA_CG_PROF=1): wall time of each stage of the structure build on stderr
  auto t_last = std::chrono::steady_clock::now();
  auto lap = [&](const char* what) {
    if (!ctx->env_cg_prof) return;
    cudaStreamSynchronize(s);
    const auto t = std::chrono::steady_clock::now();
    fprintf(stderr, "[glba] pair structure: %-28s %.3f ms\n", what, std::chrono::duration<double, std::milli>(t - t_last).count());
    t_last = t;
  };
  const int* l2g = sharded ? (const int*)ctx->l2g.as<int>() : (const int*)nullptr;
  const int* g2l = sharded ? (const int*)ctx->g2l.as<int>() : (const int*)nullptr;
  mark(ctx, PH_SETUP);
  ENSURE(long long, ctx->sp_cnt, (size_t)n_pt + 1); ENSURE(long long, ctx->sp_off, (size_t)n_pt + 1);
  LAUNCH(k_pair_count, cdiv((long)n_pt + 1, 256), 256, n_pt, (const int*)ctx->pt_start.as<int>(), (const int*)ctx->pm_cam.as<int>(),
         (const uint8_t*)ctx->cam_free.as<uint8_t>(), (const uint8_t*)ctx->pt_free.as<uint8_t>(), ctx->sp_cnt.as<long long>());
  size_t tb = 0;
  CU(cub::DeviceScan::ExclusiveSum(nullptr, tb, ctx->sp_cnt.as<long long>(), ctx->sp_off.as<long long>(), n_pt + 1, s));
  ENSURE(char, ctx->sort_tmp, tb);
  tb = ctx->sort_tmp.cap;
  CU(cub::DeviceScan::ExclusiveSum(ctx->sort_tmp.p, tb, ctx->sp_cnt.as<long long>(), ctx->sp_off.as<long long>(), n_pt + 1, s));
  g_launches.fetch_add(1);
  long long n_inst = 0;
  CU(cudaMemcpyAsync(&n_inst, ctx->sp_off.as<long long>() + n_pt, sizeof(long long), cudaMemcpyDeviceToHost, s));
  CU(cudaStreamSynchronize(s));
  int veto = (n_inst > 16 * (long long)ctx->n_obs || n_inst > 0x7fffffffLL - 64 || (!sharded && n_inst <= 0)) ? 1 : 0;
  if (sharded) {             // one rank's veto is everybody's
    ENSURE(int, ctx->sp_nruns, 2);
    CU(cudaMemcpyAsync(ctx->sp_nruns.as<int>() + 1, &veto, sizeof(int), cudaMemcpyHostToDevice, s));
    { int s__ = allreduce(ctx, ctx->sp_nruns.as<int>() + 1, 1, kNcclMax, kNcclInt32); if (s__) return s__; }
    CU(cudaMemcpyAsync(&veto, ctx->sp_nruns.as<int>() + 1, sizeof(int), cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
  }
  lap("count + scan");
  if (veto) { mark(ctx, -1); return GLBA_OK; }
  int bits = 1; while ((1ULL << bits) < (unsigned long long)K2 && bits < 63) ++bits;
  int n_loc = 0;             // distinct blocks among this rank's tracks
  ENSURE(unsigned long long, ctx->sp_ukey, (size_t)std::max<long long>(n_inst, 1)); ENSURE(int, ctx->sp_ucnt, (size_t)n_inst + 2); ENSURE(int, ctx->sp_nruns, 2);
  if (n_inst > 0) {
    ENSURE(unsigned long long, ctx->sp_key, (size_t)n_inst); ENSURE(unsigned long long, ctx->sp_key2, (size_t)n_inst);
    ENSURE(int4, ctx->sp_val, (size_t)n_inst); ENSURE(int4, ctx->sp_inst, (size_t)n_inst);
    LAUNCH(k_pair_emit, cdiv(n_pt, 256), 256, n_pt, (const int*)ctx->pt_start.as<int>(), (const int*)ctx->pm_cam.as<int>(),
           (const uint8_t*)ctx->cam_free.as<uint8_t>(), (const uint8_t*)ctx->pt_free.as<uint8_t>(), (const long long*)ctx->sp_off.as<long long>(), n_cam_g, l2g,
           ctx->sp_key.as<unsigned long long>(), ctx->sp_val.as<int4>());
    lap("emit instances");
    tb = 0;
    CU(cub::DeviceRadixSort::SortPairs(nullptr, tb, ctx->sp_key.as<unsigned long long>(), ctx->sp_key2.as<unsigned long long>(), ctx->sp_val.as<int4>(),
                                       ctx->sp_inst.as<int4>(), (int)n_inst, 0, bits, s));
    ENSURE(char, ctx->sort_tmp, tb);
    tb = ctx->sort_tmp.cap;
    CU(cub::DeviceRadixSort::SortPairs(ctx->sort_tmp.p, tb, ctx->sp_key.as<unsigned long long>(), ctx->sp_key2.as<unsigned long long>(), ctx->sp_val.as<int4>(),
                                       ctx->sp_inst.as<int4>(), (int)n_inst, 0, bits, s));
    lap("sort instances");
    // distinct keys = blocks; run lengths = instances per block
    tb = 0;
    CU(cub::DeviceRunLengthEncode::Encode(nullptr, tb, ctx->sp_key2.as<unsigned long long>(), ctx->sp_ukey.as<unsigned long long>(), ctx->sp_ucnt.as<int>(),
                                          ctx->sp_nruns.as<int>(), (int)n_inst, s));
    ENSURE(char, ctx->sort_tmp, tb);
    tb = ctx->sort_tmp.cap;
    CU(cub::DeviceRunLengthEncode::Encode(ctx->sort_tmp.p, tb, ctx->sp_key2.as<unsigned long long>(), ctx->sp_ukey.as<unsigned long long>(), ctx->sp_ucnt.as<int>(),
                                          ctx->sp_nruns.as<int>(), (int)n_inst, s));
    g_launches.fetch_add(2);
    CU(cudaMemcpyAsync(&n_loc, ctx->sp_nruns.p, sizeof(int), cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
  }
  ENSURE(int, ctx->sp_pair_start, (size_t)n_loc + 1); ENSURE(int, ctx->sp_pair_a, std::max(n_loc, 1)); ENSURE(int, ctx->sp_pair_b, std::max(n_loc, 1));
  CU(cudaMemsetAsync(ctx->sp_ucnt.as<int>() + n_loc, 0, sizeof(int), s));
  tb = 0;
  CU(cub::DeviceScan::ExclusiveSum(nullptr, tb, ctx->sp_ucnt.as<int>(), ctx->sp_pair_start.as<int>(), n_loc + 1, s));
  ENSURE(char, ctx->sort_tmp, tb);
  tb = ctx->sort_tmp.cap;
  CU(cub::DeviceScan::ExclusiveSum(ctx->sort_tmp.p, tb, ctx->sp_ucnt.as<int>(), ctx->sp_pair_start.as<int>(), n_loc + 1, s));
  g_launches.fetch_add(1);
  if (n_loc) LAUNCH(k_pair_cams, cdiv(n_loc, 256), 256, n_loc, (const unsigned long long*)ctx->sp_ukey.as<unsigned long long>(), n_cam_g, g2l,
                    ctx->sp_pair_a.as<int>(), ctx->sp_pair_b.as<int>());
  lap("blocks (run lengths, scan)");
  // blocks of the whole map and this rank's place in them
  int n_glob = n_loc;
  const unsigned long long* gkey = ctx->sp_ukey.as<unsigned long long>();
  if (sharded) {
    ENSURE(int, ctx->sp_pres, (size_t)K2 + 1); ENSURE(int, ctx->sp_gscan, (size_t)K2 + 1); ENSURE(int, ctx->sp_gid, std::max(n_loc, 1));
    CU(cudaMemsetAsync(ctx->sp_pres.p, 0, sizeof(int) * ((size_t)K2 + 1), s));
    if (n_loc) LAUNCH(k_pair_mark, cdiv(n_loc, 256), 256, n_loc, (const unsigned long long*)ctx->sp_ukey.as<unsigned long long>(), ctx->sp_pres.as<int>());
    { int s__ = allreduce(ctx, ctx->sp_pres.p, (size_t)K2, kNcclSum, kNcclInt32); if (s__) return s__; }
    LAUNCH(k_pair_flag01, cdiv(K2 + 1, 256), 256, (long)K2, ctx->sp_pres.as<int>());
    tb = 0;
    CU(cub::DeviceScan::ExclusiveSum(nullptr, tb, ctx->sp_pres.as<int>(), ctx->sp_gscan.as<int>(), (int)(K2 + 1), s));
    ENSURE(char, ctx->sort_tmp, tb);
    tb = ctx->sort_tmp.cap;
    CU(cub::DeviceScan::ExclusiveSum(ctx->sort_tmp.p, tb, ctx->sp_pres.as<int>(), ctx->sp_gscan.as<int>(), (int)(K2 + 1), s));
    g_launches.fetch_add(1);
    CU(cudaMemcpyAsync(&n_glob, ctx->sp_gscan.as<int>() + K2, sizeof(int), cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    ENSURE(unsigned long long, ctx->sp_gkey, std::max(n_glob, 1));
    LAUNCH(k_pair_compact, cdiv(K2, 256), 256, (long)K2, (const int*)ctx->sp_pres.as<int>(), (const int*)ctx->sp_gscan.as<int>(), ctx->sp_gkey.as<unsigned long long>());
    if (n_loc) LAUNCH(k_pair_gid, cdiv(n_loc, 256), 256, n_loc, (const unsigned long long*)ctx->sp_ukey.as<unsigned long long>(), (const int*)ctx->sp_gscan.as<int>(),
                      ctx->sp_gid.as<int>());
    gkey = ctx->sp_gkey.as<unsigned long long>();
  }
  if (n_glob <= 0 || (size_t)n_glob * 288 > ((size_t)1 << 30)) { mark(ctx, -1); return GLBA_OK; }      // n_glob is the same on every rank
  // row lists of the whole map: both triangles, sorted by (row, column)
  const int n_ent = 2 * n_glob;
  ENSURE(unsigned long long, ctx->sp_ekey, n_ent); ENSURE(unsigned long long, ctx->sp_ekey2, n_ent); ENSURE(int, ctx->sp_eval, n_ent); ENSURE(int, ctx->sp_eval2, n_ent);
  ENSURE(int, ctx->sp_erow, n_ent); ENSURE(int2, ctx->sp_ent, n_ent); ENSURE(int, ctx->sp_row_start, (size_t)n_cam_g + 2);
  LAUNCH(k_pair_rows, cdiv(n_glob, 256), 256, n_glob, gkey, n_cam_g, ctx->sp_ekey.as<unsigned long long>(), ctx->sp_eval.as<int>());
  tb = 0;
  CU(cub::DeviceRadixSort::SortPairs(nullptr, tb, ctx->sp_ekey.as<unsigned long long>(), ctx->sp_ekey2.as<unsigned long long>(), ctx->sp_eval.as<int>(),
                                     ctx->sp_eval2.as<int>(), n_ent, 0, bits, s));
  ENSURE(char, ctx->sort_tmp, tb);
  tb = ctx->sort_tmp.cap;
  CU(cub::DeviceRadixSort::SortPairs(ctx->sort_tmp.p, tb, ctx->sp_ekey.as<unsigned long long>(), ctx->sp_ekey2.as<unsigned long long>(), ctx->sp_eval.as<int>(),
                                     ctx->sp_eval2.as<int>(), n_ent, 0, bits, s));
  g_launches.fetch_add(3);
  LAUNCH(k_pair_entries, cdiv(n_ent, 256), 256, n_ent, (const unsigned long long*)ctx->sp_ekey2.as<unsigned long long>(), (const int*)ctx->sp_eval2.as<int>(), n_cam_g,
         ctx->sp_erow.as<int>(), ctx->sp_ent.as<int2>());
  LAUNCH(k_segment_starts, cdiv(n_cam_g + 1, 256), 256, (long)n_ent, (const int*)ctx->sp_erow.as<int>(), n_cam_g, ctx->sp_row_start.as<int>());
  lap("row lists");
  // blocks; sharded: followed by the whole map's Md | Minv | rhs rows (one all-reduce carries all four), and whole-map PCG vectors
  ctx->sp_xch_len = (size_t)36 * n_glob + (sharded ? (size_t)78 * n_cam_g : 0);
  ENSURE(double, ctx->sp_blocks, ctx->sp_xch_len);
  if (sharded) ENSURE(double, ctx->sp_vec, (size_t)36 * n_cam_g);
  const int max_grid = std::min(ctx->n_sm, 256);            // (a lane of k_cg_bsr adds the partial sums of up to 8 CTAs)
  ctx->cg_grid = std::max(1, std::min(max_grid, cdiv(n_cam_g, NT_CGP / 32)));
  // Small maps (at most one row per warp of a full grid): the blocks of a CTA's rows stay in shared memory for the whole solve,
  // so the rows are dealt out in contiguous ranges of (nearly) equal ENTRY count, at most NT_CGP/32 rows and CG_SMEM_CAP entries
  // each.  With 16 consecutive rows per CTA the CTAs of revisited streets held 2.5x the average and every iteration waited for
  // them (measured with GLBA_CG_PROF=1: 9 of 13 us).  Smallest feasible bound by bisection over a greedy split.
  ctx->cg_reg = false;
  if (ctx->env_cg_reg && n_cam_g <= max_grid * (NT_CGP / 32)) {
    std::vector<int> rs((size_t)n_cam_g + 1);
    CU(cudaMemcpyAsync(rs.data(), ctx->sp_row_start.p, sizeof(int) * rs.size(), cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    const int rows_per = NT_CGP / 32;
    auto split = [&](int bound, std::vector<int>* out) -> int {
      int ctas = 0, r = 0;
      if (out) { out->clear(); out->push_back(0); }
      while (r < n_cam_g) {
        int cnt = 0, ent = 0;
        while (r < n_cam_g && cnt < rows_per && (cnt == 0 || ent + (rs[r + 1] - rs[r]) <= bound)) { ent += rs[r + 1] - rs[r]; ++cnt; ++r; }
        ++ctas;
        if (out) out->push_back(r);
      }
      return ctas;
    };
    int lo = 1, hi = std::max(1, rs[n_cam_g]);
    for (int r = 0; r < n_cam_g; ++r) lo = std::max(lo, rs[r + 1] - rs[r]);
    while (lo < hi) { const int mid = lo + (hi - lo) / 2; if (split(mid, nullptr) <= max_grid) hi = mid; else lo = mid + 1; }
    if (lo <= (int)CG_SMEM_CAP && split(lo, nullptr) <= max_grid) {
      std::vector<int> cr;
      ctx->cg_grid_reg = split(lo, &cr);
      ENSURE(int, ctx->sp_cta_rows, cr.size());
      CU(cudaMemcpyAsync(ctx->sp_cta_rows.p, cr.data(), sizeof(int) * cr.size(), cudaMemcpyHostToDevice, s));
      CU(cudaStreamSynchronize(s));
      ctx->cg_reg = true;
    }
  }
  ENSURE(double, ctx->sp_part, 2 * (size_t)max_grid + 2); ENSURE(unsigned, ctx->sp_bar, 4); ENSURE(long long, ctx->sp_prof, 8);
  lap("buffers + row split");
  CHECK_LAUNCHES();
  // (the sort scratch stays allocated: a context that solves map after map would pay cudaMalloc / cudaFree of ~50 B per
  // instance at every load: measured 190 ms against 7 ms on C4)
  ctx->n_pairs = n_loc; ctx->n_pairs_g = n_glob; ctx->n_inst = n_inst; ctx->use_explicit = true;
  mark(ctx, -1);
  return GLBA_OK;
}

// blocks of the current linearisation and damping.  Sharded: this rank's partial blocks land in the whole map's numbering and
// are summed over the ranks together with the owners' Md | Minv | rhs rows: one all-reduce per LM iteration, after which every
// rank holds the complete reduced system and solves it redundantly (no collective inside the PCG).
int launch_schur_pairs(glba_ctx* ctx) {
  const int c = ctx->cur;
  const bool sharded = ctx->world > 1;
  if (sharded) CU(cudaMemsetAsync(ctx->sp_blocks.p, 0, sizeof(double) * ctx->sp_xch_len, ctx->stream));
  if (ctx->n_pairs > 0)
    LAUNCH(k_schur_pairs, cdiv((long)ctx->n_pairs * 32, NT_SP), NT_SP, ctx->n_pairs, (const int*)ctx->sp_pair_a.as<int>(), (const int*)ctx->sp_pair_b.as<int>(),
           (const int*)ctx->sp_pair_start.as<int>(), (const int4*)ctx->sp_inst.as<int4>(), (const double4*)ctx->rec_pm.as<double4>(),
           (const double*)ctx->camtab[c].as<double>(), (const double*)ctx->cinv.as<double>(), ctx->K, sharded ? (const int*)ctx->sp_gid.as<int>() : (const int*)nullptr,
           ctx->sp_blocks.as<double>());
  if (sharded) {
    const int n_cam = ctx->n_cam, n_cam_g = ctx->n_cam_g;
    double* base = ctx->sp_blocks.as<double>() + (size_t)36 * ctx->n_pairs_g;
    const int* l2g = ctx->l2g.as<int>();
    const uint8_t* owned = ctx->cam_owned.as<uint8_t>();
    LAUNCH(k_scatter_rows, cdiv((long)n_cam * 36, 256), 256, n_cam, l2g, owned, 1, (const double*)ctx->Md.as<double>(), 36, base);
    LAUNCH(k_scatter_rows, cdiv((long)n_cam * 36, 256), 256, n_cam, l2g, owned, 1, (const double*)ctx->Minv.as<double>(), 36, base + (size_t)36 * n_cam_g);
    LAUNCH(k_scatter_rows, cdiv((long)n_cam * 6, 256), 256, n_cam, l2g, owned, 1, (const double*)ctx->rhs.as<double>(), 6, base + (size_t)72 * n_cam_g);
    AR(ctx->sp_blocks.p, ctx->sp_xch_len, kNcclSum);
  }
  return GLBA_OK;
}
// the whole PCG in one cooperative launch (k_cg_bsr); cg receives iterations and stop reason
int launch_cg_bsr(glba_ctx* ctx, const glba_options* o, int max_it) {
  const bool sharded = ctx->world > 1;
  const int n_rows = sharded ? ctx->n_cam_g : ctx->n_cam;
  const double* base = ctx->sp_blocks.as<double>() + (size_t)36 * ctx->n_pairs_g;
  const double* Md = sharded ? base : (const double*)ctx->Md.as<double>();
  const double* Minv = sharded ? base + (size_t)36 * n_rows : (const double*)ctx->Minv.as<double>();
  const double* rhs = sharded ? base + (size_t)72 * n_rows : (const double*)ctx->rhs.as<double>();
  double* v = ctx->sp_vec.as<double>();
  const size_t L = (size_t)6 * n_rows;
  double* x = sharded ? v : ctx->cg_x.as<double>();
  double* r = sharded ? v + L : ctx->cg_r.as<double>();
  double* u = sharded ? v + 2 * L : ctx->cg_q.as<double>();
  double* p = sharded ? v + 3 * L : ctx->cg_p.as<double>();
  double* sv = sharded ? v + 4 * L : ctx->pg.as<double>();
  double* w = sharded ? v + 5 * L : ctx->yg.as<double>();
  CU(cudaMemsetAsync(ctx->sp_bar.p, 0, sizeof(unsigned), ctx->stream));
  const bool reg = ctx->cg_reg;          // one row per warp, vectors in registers, the CTA's blocks in shared memory
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(reg ? ctx->cg_grid_reg : ctx->cg_grid); cfg.blockDim = dim3(NT_CGP); cfg.dynamicSmemBytes = reg ? CG_SMEM_BYTES : 0; cfg.stream = ctx->stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeCooperative;
  attr[0].val.cooperative = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  prev_small(ctx->stream) = false;
  const cudaError_t le = cudaLaunchKernelEx(&cfg, reg ? k_cg_bsr<true> : k_cg_bsr<false>, n_rows, (const int*)ctx->sp_row_start.as<int>(),
                        (const int2*)ctx->sp_ent.as<int2>(), (const double*)ctx->sp_blocks.as<double>(), Md, Minv, rhs, x, r, u, p, sv, w,
                        ctx->sp_part.as<double>(), ctx->sp_bar.as<unsigned>(), ctx->cgst.as<CgState>(), o->cg_rel_tol, max_it,
                        reg ? (int)CG_SMEM_CAP : 0, reg ? (const int*)ctx->sp_cta_rows.as<int>() : (const int*)nullptr,
                        ctx->env_cg_prof ? ctx->sp_prof.as<long long>() : (long long*)nullptr);
  if (le == cudaErrorCooperativeLaunchTooLarge || le == cudaErrorLaunchOutOfResources) {
    // the grid cannot be made resident here (SMs reserved by another client of the device): the caller falls back to the
    // matrix-free product where it can
    (void)cudaGetLastError();
    return GLBA_E_UNSUPPORTED;
  }
  CU(le);
  if (ctx->env_cg_prof) {
    long long h[8] = {0};
    CU(cudaMemcpyAsync(h, ctx->sp_prof.p, sizeof(long long) * 7, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    const double it = (double)std::max<long long>(h[6], 1);
    fprintf(stderr, "[glba] k_cg_bsr<%d> CTA 0 thread 0, cycles per iteration over %lld iterations: product %.0f | wait CTA %.0f | barrier 1 %.0f | totals %.0f | update %.0f | barrier 2 %.0f\n",
            reg ? 1 : 0, h[6], h[0] / it, h[1] / it, h[2] / it, h[3] / it, h[4] / it, h[5] / it);
  }
  g_launches.fetch_add(1, std::memory_order_relaxed);
  if (sharded)       // this rank's cameras of the solution
    LAUNCH(k_gather_rows, cdiv((long)ctx->n_cam * 6, 256), 256, ctx->n_cam, (const int*)ctx->l2g.as<int>(), (const double*)x, 6, ctx->cg_x.as<double>());
  return GLBA_OK;
}

// vectors of the single-reduction PCG (glba_cam.cuh): x = cg_x, r = cg_r, u = M^-1 r = cg_q, p = cg_p, s = S p = pg, w = S u = yg
int launch_cg_iteration(glba_ctx* ctx, const glba_options* o, double radius, CgState* cg, int li) {
  const int c = ctx->cur;
  const int n_cam = ctx->n_cam;
  // yhat partials: either per chunk of the camera-major order (two-kernel product) or per camera from the fused tile kernel
  const int* ysum_start = ctx->cam_chunk_start.as<int>();
  const double* ysum_part = ctx->part_cm.as<double>();
  if (ctx->use_fused) {
    launch_spmv_fused(ctx, o, cg, li);
    LAUNCH(k_cam_combine, cdiv((long)n_cam * 32, 256), 256, n_cam, (const int*)ctx->cam_tp_start.as<int>(), (const int*)ctx->cam_tp.as<int>(),
           (const double*)ctx->tpart.as<double>(), ctx->n_ovf > 0 ? (const int*)ctx->cam_ov_start.as<int>() : (const int*)nullptr,
           (const int*)ctx->cam_ov.as<int>(), (const double*)ctx->ovf_c.as<double>(), ctx->yhat.as<double>(), (const CgState*)cg, li);
    ysum_start = ctx->cam_iota.as<int>(); ysum_part = ctx->yhat.as<double>();
  } else {
    launch_point_pass0(ctx, o, cg, li);
    LAUNCH(k_spmv_cm, ctx->n_chunks, NT_CM, cm_args(ctx), (const double4*)ctx->rec_cm.as<double4>(), (const double*)ctx->camtab[c].as<double>(),
           (const double4*)ctx->u4.as<double4>(), (const CgState*)cg, li, ctx->part_cm.as<double>());
  }
#define CG_W_ARGS n_cam, (const uint8_t*)ctx->cam_free.as<uint8_t>(), (const double*)ctx->camtab[c].as<double>(), (const double*)ctx->Bc.as<double>(), \
    (const double*)ctx->lamc.as<double>(), 1.0 / radius, ysum_start, ysum_part, \
    (const double*)ctx->cg_r.as<double>(), (const double*)ctx->cg_q.as<double>(), ctx->yg.as<double>(), cg, li
#define CG_UPD_ARGS n_cam, (const double*)ctx->camtab[c].as<double>(), (const double*)ctx->Minv.as<double>(), ctx->cg_q.as<double>(), ctx->yg.as<double>(), \
    ctx->cg_p.as<double>(), ctx->pg.as<double>(), ctx->cg_x.as<double>(), ctx->cg_r.as<double>(), ctx->xtab.as<double>(), cg, li
  if (ctx->owner) {
    const int ns = ctx->n_shared;
    LAUNCH(k_cg_w<true>, ctx->grid_c, NT_C, CG_W_ARGS, (const uint8_t*)ctx->cam_owned.as<uint8_t>(), (const int*)ctx->cam_shared.as<int>(),
           ctx->xsend6.as<double>(), ns, ctx->partc.as<double>(), ctx->counters.as<unsigned>() + 2);
    const int r = g_nccl.AllReduce(ctx->xsend6.p, ctx->xrecv.p, 6 * (size_t)ns + 2, kNcclFloat64, kNcclSum, ctx->comm, ctx->stream);
    if (r != 0) return fail(ctx, GLBA_E_NCCL, "ncclAllReduce (PCG): %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "error");
    LAUNCH(k_cg_update<true>, ctx->grid_c, NT_C, CG_UPD_ARGS, (const int*)ctx->cam_shared.as<int>(), (const double*)ctx->xrecv.as<double>(), ns);
  } else {
    LAUNCH(k_cg_w<false>, ctx->grid_c, NT_C, CG_W_ARGS, (const uint8_t*)nullptr, (const int*)nullptr, (double*)nullptr, 0,
           ctx->partc.as<double>(), ctx->counters.as<unsigned>() + 2);
    LAUNCH(k_cg_update<false>, ctx->grid_c, NT_C, CG_UPD_ARGS, (const int*)nullptr, (const double*)nullptr, 0);
  }
  return GLBA_OK;
}

// Block-Jacobi PCG on the implicit Schur complement; solution in cg_x.  Returns iterations in *iters.
int do_pcg(glba_ctx* ctx, const glba_options* o, double radius, int* iters) {
  const int c = ctx->cur;
  const int n_cam = ctx->n_cam;
  *iters = 0;
  mark(ctx, PH_SOLVE);
  if (ctx->n_free_cam == 0 || n_cam == 0) {
    CU(cudaMemsetAsync(ctx->cg_x.p, 0, sizeof(double) * 6 * std::max(n_cam, 1), ctx->stream));
    mark(ctx, -1);
    return GLBA_OK;
  }
  if (ctx->world > 1 && !ctx->owner)
    return fail(ctx, GLBA_E_UNSUPPORTED, "sharded PCG needs the owner-computes layout (load the problem with linsolve = PCG or > %d cameras)", DN_MAXCAM);
  const int dim = 6 * ctx->n_free_cam;
  const int max_it = o->cg_max_iters > 0 ? o->cg_max_iters : std::min(4000, 4 * dim);
  CgState* cg = ctx->cgst.as<CgState>();
  if (ctx->use_explicit) {
    // assembled reduced camera matrix: blocks for this linearisation and damping, then the whole PCG in one launch
    mark(ctx, PH_SCHUR);
    { const int s__ = launch_schur_pairs(ctx); if (s__) return s__; }
    mark(ctx, PH_SOLVE);
    const int s__ = launch_cg_bsr(ctx, o, max_it);
    if (s__ == GLBA_E_UNSUPPORTED && ctx->world == 1) {
      ctx->use_explicit = false;             // this map keeps the matrix-free product from here on
      return do_pcg(ctx, o, radius, iters);
    }
    if (s__) return s__;
    CHECK_LAUNCHES();
    // the iteration count and stop reason travel with the step's scalars: no synchronisation of its own (run_lm reads h_cg
    // after the read-back that follows the step)
    CU(cudaMemcpyAsync(ctx->h_cg, cg, sizeof(CgState), cudaMemcpyDeviceToHost, ctx->stream));
    ctx->cg_pending = true;
    *iters = -1;
    mark(ctx, -1);
    return GLBA_OK;
  }
  LAUNCH(k_cg_start, ctx->grid_c, NT_C, n_cam, (const double*)ctx->camtab[c].as<double>(), (const double*)ctx->Minv.as<double>(),
         (const double*)ctx->rhs.as<double>(), ctx->cg_x.as<double>(), ctx->cg_r.as<double>(), ctx->cg_q.as<double>(), ctx->cg_p.as<double>(),
         ctx->pg.as<double>(), ctx->xtab.as<double>(), cg, o->cg_rel_tol, max_it);
  const int poll = 8;
  int launched = 0;
  // the stop test lags one product behind the update (single-reduction recurrence): max_it updates need max_it + 1 launches
  for (;;) {
    for (int b = 0; b < poll && launched < max_it + 1; ++b, ++launched) { const int s__ = launch_cg_iteration(ctx, o, radius, cg, launched); if (s__) return s__; }
    CHECK_LAUNCHES();
    CU(cudaMemcpyAsync(ctx->h_cg, cg, sizeof(CgState), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    if (ctx->h_cg->done_at <= launched || launched >= max_it + 1) break;
  }
  *iters = ctx->h_cg->iters;
  mark(ctx, -1);
  return GLBA_OK;
}

// Exact solve of the reduced camera system for small windows: explicit S, Cholesky (glba_dense.cuh).
int do_dense(glba_ctx* ctx, const glba_options* o, double radius, const LmCtl* ctl = nullptr) {
  const int c = ctx->cur;
  const int n_cam = ctx->n_cam;
  const int n = 6 * n_cam;
  const size_t sm_schur = ((size_t)DN_TP * n_cam * 24 + (size_t)(n_cam * (n_cam + 1) / 2) * 36) * sizeof(double) + DN_TP * sizeof(unsigned) + 16;
  const size_t sm_solve = ((size_t)n * (n | 1) + 2 * (size_t)n) * sizeof(double);
  const int len = (n_cam * (n_cam + 1) / 2) * 36 + n;
  mark(ctx, PH_SCHUR);
  LAUNCH_SMEM(k_dense_schur, ctx->dn_grid, DN_NT, sm_schur, pm_args(ctx, o), n_cam, (const uint8_t*)ctx->cam_free.as<uint8_t>(),
      (const double4*)ctx->rec_pm.as<double4>(), (const double*)ctx->camtab[c].as<double>(), (const double*)ctx->cinv.as<double>(), (const double4*)ctx->u0p.as<double4>(), ctx->dn_ppc,
      ctx->dn_part.as<double>(), ctl);
#define DN_RED_ARGS n_cam, (const double*)ctx->dn_part.as<double>(), (const uint8_t*)ctx->cam_free.as<uint8_t>(), (const double*)ctx->Bc.as<double>(), \
    (const double*)ctx->gc.as<double>(), (const double*)ctx->lamc.as<double>(), 1.0 / radius, ctx->dn_red.as<double>(), ctx->dn_full.as<double>()
  if (ctx->world > 1) {    // sum over this rank's CTAs, all-reduce the pair sums, then assemble
    LAUNCH(k_dense_reduce, cdiv(len * 8, 256), 256, ctx->dn_grid, DN_RED_ARGS, 0, ctl);
    AR(ctx->dn_red.as<double>(), (size_t)len, kNcclSum);
    LAUNCH(k_dense_reduce, cdiv(len * 8, 256), 256, 0, DN_RED_ARGS, 1, ctl);
  } else {
    LAUNCH(k_dense_reduce, cdiv(len * 8, 256), 256, ctx->dn_grid, DN_RED_ARGS, 1, ctl);
  }
  mark(ctx, PH_SOLVE);
  LAUNCH_SMEM(k_dense_solve, 1, DN_NS, sm_solve, n_cam, (const uint8_t*)ctx->cam_free.as<uint8_t>(), (const double*)ctx->dn_full.as<double>(),
      ctx->cg_x.as<double>(), ctx->Md.as<double>(), ctx->rhs.as<double>(), ctx->d_scal, ctl);
  mark(ctx, -1);
  CHECK_LAUNCHES();
  return GLBA_OK;
}

bool want_dense(const glba_ctx* ctx, const glba_options* o) {
  if (ctx->owner) return false;          // owner-computes layout (sharded, PCG): n_cam is this rank's share, every rank must decide alike
  if (ctx->n_free_cam == 0 || ctx->n_cam > DN_MAXCAM || ctx->has_dup) return false;
  if (o->linsolve == GLBA_LINSOLVE_PCG) return false;
  if (o->linsolve == GLBA_LINSOLVE_DENSE) return true;
  return 6 * ctx->n_free_cam <= o->dense_max_dim;
}

// Sharded runs: d_scal[S_GMAX_P] holds only this rank's max |g_point|; the global value is the max over the per-rank
// slots of the all-reduced payload.  EVERY read-back of the scalars goes through here, so all ranks take the same
// gradient-tolerance decision (a rank leaving the loop alone would hang the others in their next collective).
int read_scalars(glba_ctx* ctx) {
  CHECK_LAUNCHES();
  double h_late[NLATE];
  if (ctx->owner) {
    // owner layout: the camera-derived scalars are per-rank partial sums (a camera counts on its owner) and the candidate-step
    // sums cover this rank's points only: one small all-reduce completes both, out of place (the partials stay partial)
    double* ls = ctx->late.as<double>();
    LAUNCH(k_late_pack, 1, 32, (const double*)ctx->d_scal, ctx->rank, ls);
    const int r = g_nccl.AllReduce(ls, ls + NLATE, (size_t)NLATE, kNcclFloat64, kNcclSum, ctx->comm, ctx->stream);
    if (r != 0) return fail(ctx, GLBA_E_NCCL, "ncclAllReduce (scalars): %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "error");
    CU(cudaMemcpyAsync(h_late, ls + NLATE, sizeof(h_late), cudaMemcpyDeviceToHost, ctx->stream));
  }
  CU(cudaMemcpyAsync(ctx->h_scal, ctx->d_scal, sizeof(double) * NSCAL, cudaMemcpyDeviceToHost, ctx->stream));
  int p2p_failed = 0;
  if (ctx->p2p_ok) CU(cudaMemcpyAsync(&p2p_failed, p2p_err(ctx->p2p_local), sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  if (p2p_failed) return fail(ctx, GLBA_E_NCCL, "peer exchange: a rank did not publish its buffer within two minutes");
  if (ctx->world > 1) {
    double m = 0.0;
    for (int r = 0; r < ctx->world; ++r) m = std::max(m, ctx->h_scal[S_GSLOT0 + r]);
    ctx->h_scal[S_GMAX_P] = m;
  }
  if (ctx->owner) {
    const int dst[10] = {S_COST_C, S_YN2_P, S_YG_P, S_YLY_P, S_BAD_C, S_XN2_C, S_YN2_C, S_YG_C, S_YLY_C, S_NOTPD_C};
    for (int q = 0; q < 10; ++q) ctx->h_scal[dst[q]] = h_late[q];
    double m = 0.0;
    for (int r = 0; r < ctx->world; ++r) m = std::max(m, h_late[10 + r]);
    ctx->h_scal[S_GMAX_C] = m;
  }
  collect(ctx);
  return GLBA_OK;
}

// candidate state, back-substitution, candidate cost; leaves the scalars in h_scal (synchronises)
int do_step(glba_ctx* ctx, const glba_options* o, double radius) {
  const int c = ctx->cur, d = c ^ 1;
  const int n_cam = ctx->n_cam, n_pt = ctx->n_pt;
  mark(ctx, PH_UPDATE);
  if (n_cam) LAUNCH(k_cam_step2, ctx->grid_c, NT_C, n_cam, (const uint8_t*)ctx->cam_free.as<uint8_t>(), (const double*)ctx->cam[c].as<double>(),
                    (const double*)ctx->camtab[c].as<double>(), (const double*)ctx->cg_x.as<double>(), (const double*)ctx->gc.as<double>(),
                    (const double*)ctx->lamc.as<double>(), 1.0 / radius, ctx->cam[d].as<double>(), ctx->camtab[d].as<double>(), ctx->xtab.as<double>(),
                    ctx->partc.as<double>(), ctx->counters.as<unsigned>() + 3, ctx->d_scal, ctx->mode, (const LmCtl*)nullptr, ctx->owner ? (const uint8_t*)ctx->cam_owned.as<uint8_t>() : (const uint8_t*)nullptr);
  if (n_pt) launch_point_pass1(ctx, o, radius);
  if (ctx->world > 1 && !ctx->owner) AR(ctx->d_scal + S_COST_C, 5, kNcclSum);       // owner layout: read_scalars exchanges them
  mark(ctx, -1);
  return read_scalars(ctx);
}

int fetch_scal(glba_ctx* ctx) { return read_scalars(ctx); }

// Small windows (exact dense reduced solve, one GPU, Ceres semantics): the whole trust-region loop runs on the device.
// The host enqueues LM iterations WITHOUT knowing their outcome — every kernel is gated by the LmCtl the decision kernels
// maintain (glba_lm.cuh) — and synchronises once per kPollIters iterations to read one flag.  Entered after the initial
// linearisation (cost, |g|, |x| already on the host, sum[0] filled, loop-top tests of iteration 1 passed).
constexpr int kPollIters = 4;
bool device_lm_eligible(const glba_ctx* ctx, const glba_options* o, bool dense, bool g2o) {
  return dense && !g2o && ctx->world == 1 && ctx->use_tiles && !o->verbose && !ctx->env_host_lm && ctx->n_obs > 0;
}
int enqueue_lm_iteration(glba_ctx* ctx, const glba_options* o, const LmCtl* ctl, const LmParams& P) {
  const int c = ctx->cur, d = c ^ 1;
  const int n_cam = ctx->n_cam, n_pt = ctx->n_pt;
  const double radius = 1.0;        // by-value radii are ignored: the kernels read ctl->inv_radius
  int st;
  if ((st = do_redamp(ctx, radius, ctl))) return st;                                    // runs only after a rejected / invalid step
  if ((st = do_dense(ctx, o, radius, ctl))) return st;
  if (n_cam) LAUNCH(k_cam_step2, ctx->grid_c, NT_C, n_cam, (const uint8_t*)ctx->cam_free.as<uint8_t>(), (const double*)ctx->cam[c].as<double>(),
                    (const double*)ctx->camtab[c].as<double>(), (const double*)ctx->cg_x.as<double>(), (const double*)ctx->gc.as<double>(),
                    (const double*)ctx->lamc.as<double>(), 1.0 / radius, ctx->cam[d].as<double>(), ctx->camtab[d].as<double>(), ctx->xtab.as<double>(),
                    ctx->partc.as<double>(), ctx->counters.as<unsigned>() + 3, ctx->d_scal, ctx->mode, ctl, ctx->owner ? (const uint8_t*)ctx->cam_owned.as<uint8_t>() : (const uint8_t*)nullptr);
  LmHook hook; hook.ctl = ctx->lmctl.as<LmCtl>(); hook.sum = ctx->dsum.as<glba_summary>(); hook.P = P;
  if (n_pt) launch_point_pass1(ctx, o, radius, ctl, &hook);      // its last CTA takes the accept / reject decision (lm_decide)
  // accepted: candidate -> current, re-linearise (all three exit at once otherwise)
  LAUNCH(k_accept_copy, std::max(1, std::min(4 * ctx->n_sm, cdiv((long)30 * n_cam + n_pt, 256))), 256, ctl, n_cam, n_pt,
         (const double*)ctx->cam[d].as<double>(), (const double*)ctx->camtab[d].as<double>(), (const double4*)ctx->pt4[d].as<double4>(),
         ctx->cam[c].as<double>(), ctx->camtab[c].as<double>(), ctx->pt4[c].as<double4>());
  if (n_pt) { if ((st = launch_linearize_points(ctx, o, 0, radius, ctl))) return st; }
  if (ctx->n_chunks) LAUNCH(k_linearize_cm, ctx->n_chunks, NT_HCM, cm_args(ctx), (const double4*)ctx->rec_cm.as<double4>(),
                            (const double*)ctx->camtab[c].as<double>(), ctx->part_cm.as<double>(), ctl);
  if (n_cam) launch_cam_lin_fin(ctx, o, 0, ctl, &hook);          // its last CTA absorbs the re-linearisation (lm_absorb)
  return GLBA_OK;
}
int run_lm_device(glba_ctx* ctx, const glba_options* o, glba_summary* sum, double radius, double cost, double gmax, double x_norm) {
  ENSURE(LmCtl, ctx->lmctl, 1); ENSURE(glba_summary, ctx->dsum, 1);
  LmCtl h{};
  h.inv_radius = 1.0 / radius; h.done = 0; h.accepted = 0; h.need_redamp = 0;
  h.radius = radius; h.decrease_factor = 2.0; h.cost = cost; h.gmax = gmax; h.x_norm = x_norm;
  h.it = 0; h.n_invalid = 0; h.n_rejected = 0;
  LmParams P;
  P.max_iters = o->max_iters; P.max_invalid = o->max_consecutive_invalid_steps; P.n_free_cam = ctx->n_free_cam;
  P.function_tol = o->function_tol; P.gradient_tol = o->gradient_tol; P.parameter_tol = o->parameter_tol;
  P.max_radius = o->max_radius; P.min_radius = o->min_radius; P.min_rel = o->min_relative_decrease;
  cudaStream_t s = ctx->stream;
  CU(cudaMemcpyAsync(ctx->lmctl.p, &h, sizeof(h), cudaMemcpyHostToDevice, s));
  CU(cudaMemcpyAsync(ctx->dsum.p, sum, sizeof(*sum), cudaMemcpyHostToDevice, s));       // trace so far (index 0), counters
  const LmCtl* ctl = ctx->lmctl.as<LmCtl>();
  int enq = 0, st;
  for (;;) {
    for (int b = 0; b < kPollIters && enq < o->max_iters; ++b, ++enq)
      if ((st = enqueue_lm_iteration(ctx, o, ctl, P))) return st;
    CU(cudaMemcpyAsync(ctx->h_flags, &ctx->lmctl.as<LmCtl>()->done, sizeof(int), cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    if (ctx->h_flags[0] != 0) break;
    if (enq >= o->max_iters) return fail(ctx, GLBA_E_CUDA, "device LM loop did not terminate within max_iters");
  }
  CU(cudaMemcpyAsync(sum, ctx->dsum.p, sizeof(*sum), cudaMemcpyDeviceToHost, s));
  CU(cudaStreamSynchronize(s));
  collect(ctx);
  return GLBA_OK;
}

// The trust-region loop (Ceres TrustRegionMinimizer semantics; see oracle/glba_oracle.cpp for the
// statement-by-statement restatement this mirrors).  One host synchronisation per LM iteration: the scalars of the
// re-linearisation that follows an accepted step (cost, |g|_inf, |x|) are read together with the NEXT step's
// scalars; the loop-top tests that need them are applied retroactively (a step computed past a gradient-tolerance
// stop is simply discarded), which is observationally identical to Ceres' order of tests.
int run_lm(glba_ctx* ctx, const glba_options* o, glba_summary* sum) {
  int st;
  for (int q = 0; q < PH_COUNT; ++q) if (q != PH_SETUP) ctx->t_phase[q] = 0.0;
  sum->termination = GLBA_TERM_NO_CONVERGENCE; sum->stop_reason = GLBA_STOP_NONE;
  // GLBA_MODE_G2O runs the same loop: with no Jacobi scaling and the LM diagonal clamped to [1, 1] the damping is
  // lambda I with lambda = 1 / radius, and g2o's lambda update (x clamp(1-(2 rho-1)^3, 1/3, .) on accept; x nu, nu x 2 on
  // reject) is the radius update below.  What differs: lambda0 from the Hessian diagonal, the +1e-3 guard in rho,
  // acceptance on rho > 0, no tolerance / invalid-step tests, trials-per-iteration and iteration counting.
  const bool g2o = (o->mode == GLBA_MODE_G2O);
  if (g2o != (ctx->mode == GLBA_MODE_G2O)) return fail(ctx, GLBA_E_INVALID_ARG, "options.mode differs from the mode the problem was loaded with");
  glba_options og;
  if (g2o) {
    og = *o;
    og.jacobi_scaling = 0; og.min_lm_diagonal = 1.0; og.max_lm_diagonal = 1.0;
    og.function_tol = 0.0; og.parameter_tol = 0.0; og.gradient_tol = 0.0; og.min_relative_decrease = 0.0;
    og.min_radius = 0.0; og.max_radius = std::numeric_limits<double>::infinity();
    o = &og;
  }
  double radius = g2o ? 1.0 : o->initial_radius, decrease_factor = 2.0;
  int n_invalid = 0, n_rejected = 0;
  const bool dense = want_dense(ctx, o);
  if (o->linsolve == GLBA_LINSOLVE_DENSE && !dense && ctx->n_free_cam > 0)
    return fail(ctx, GLBA_E_UNSUPPORTED, "dense solve needs <= %d cameras and no duplicate (point,camera) observations", DN_MAXCAM);
  ctx->cg_pending = false;
  if (!dense && ctx->n_free_cam > 0) { if ((st = ensure_explicit(ctx, o))) return st; }
  if ((st = do_linearize_impl(ctx, o, 1, radius, !dense && !g2o))) return st;
  bool fresh = true;        // point blocks are damped for the current radius
  if (g2o) {                // computeLambdaInit: lambda0 = tau * max diag H over the free vertices
    ENSURE(unsigned long long, ctx->hmax, 1);
    CU(cudaMemsetAsync(ctx->hmax.p, 0, sizeof(unsigned long long), ctx->stream));
    if (ctx->n_pt + ctx->n_cam > 0)
      LAUNCH(k_hmax, cdiv((long)ctx->n_pt + ctx->n_cam, 256), 256, ctx->n_pt, (const uint8_t*)ctx->pt_free.as<uint8_t>(), (const double*)ctx->Craw.as<double>(),
             ctx->n_cam, (const uint8_t*)ctx->cam_free.as<uint8_t>(), (const double*)ctx->Bc.as<double>(), (const double*)ctx->camtab[ctx->cur].as<double>(),
             ctx->hmax.as<unsigned long long>());
    if (ctx->world > 1) AR(ctx->hmax.p, 1, kNcclMax);
    double hm = 0.0;
    CU(cudaMemcpyAsync(ctx->h_flags, ctx->hmax.p, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    std::memcpy(&hm, ctx->h_flags, sizeof(double));
    radius = (hm > 0.0) ? 1.0 / (o->g2o_tau * hm) : 0.0;
    fresh = false; ctx->schur_fresh = false;
  }
  if ((st = fetch_scal(ctx))) return st;
  const double* S = ctx->h_scal;
  if (S[S_BAD] > 0.0 || !std::isfinite(S[S_COST])) {
    sum->termination = GLBA_TERM_FAILURE; sum->stop_reason = GLBA_STOP_NUMERIC; sum->status = GLBA_E_NUMERIC;
    return fail(ctx, GLBA_E_NUMERIC, "non-finite residual at the initial point");
  }
  sum->n_linearizations = 1;
  double cost = S[S_COST];
  double gmax = std::max(S[S_GMAX_P], S[S_GMAX_C]);
  double x_norm = std::sqrt(S[S_XN2_P] + S[S_XN2_C]);
  sum->initial_cost = cost; sum->cost[0] = cost; sum->cost_candidate[0] = cost; sum->radius[0] = radius; sum->gradient_max_norm[0] = gmax;
  int it = 0;
  bool pending = false;     // a re-linearisation was enqueued whose scalars have not been read yet
  auto absorb_pending = [&]() {          // returns false on a non-finite re-linearisation
    pending = false;
    if (S[S_BAD] > 0.0 || !std::isfinite(S[S_COST])) return false;
    cost = S[S_COST]; gmax = std::max(S[S_GMAX_P], S[S_GMAX_C]); x_norm = std::sqrt(S[S_XN2_P] + S[S_XN2_C]);
    return true;
  };
  bool on_device = false;
  if (ctx->n_obs == 0 && ctx->world == 1) { sum->termination = GLBA_TERM_CONVERGENCE; sum->stop_reason = GLBA_STOP_GRADIENT_TOL; }
  else if (g2o && !(radius > 0.0)) { sum->termination = GLBA_TERM_CONVERGENCE; sum->stop_reason = GLBA_STOP_GRADIENT_TOL; }   // nothing free
  else if (device_lm_eligible(ctx, o, dense, g2o) && o->max_iters > 0 && gmax > o->gradient_tol && radius > o->min_radius) {
    // small window: decisions on the device, no host synchronisation inside the iteration (run_lm_device)
    if ((st = run_lm_device(ctx, o, sum, radius, cost, gmax, x_norm))) return st;
    on_device = true;
  }
  else for (;;) {
    if (g2o) {
      if (sum->n_successful >= o->max_iters) { sum->stop_reason = GLBA_STOP_MAX_ITERS; break; }
      if (it >= GLBA_MAX_ITERS || !(radius > 0.0)) { sum->stop_reason = GLBA_STOP_TRIALS; break; }
    } else {
      if (it >= o->max_iters) { sum->termination = GLBA_TERM_NO_CONVERGENCE; sum->stop_reason = GLBA_STOP_MAX_ITERS; break; }
      if (!pending && gmax <= o->gradient_tol) { sum->termination = GLBA_TERM_CONVERGENCE; sum->stop_reason = GLBA_STOP_GRADIENT_TOL; break; }
      if (radius <= o->min_radius) { sum->termination = GLBA_TERM_CONVERGENCE; sum->stop_reason = GLBA_STOP_MIN_RADIUS; break; }
    }
    ++it;
    if (!fresh) { if ((st = do_redamp(ctx, radius))) return st; ctx->schur_fresh = false; }
    fresh = true;
    int cg_it = 0;
    if (dense) { if ((st = do_dense(ctx, o, radius))) return st; }
    else {
      if (ctx->n_free_cam > 0 && !ctx->schur_fresh) { if ((st = do_schur(ctx, radius))) return st; }
      ctx->schur_fresh = false;
      if ((st = do_pcg(ctx, o, radius, &cg_it))) return st;
    }
    sum->cg_iters[it] = cg_it;
    if ((st = do_step(ctx, o, radius))) return st;
    if (ctx->cg_pending) {       // the PCG's CgState arrived with the step's scalars
      ctx->cg_pending = false;
      if (ctx->h_cg->reason == 4) return fail(ctx, GLBA_E_CUDA, "PCG grid barrier timed out");
      cg_it = ctx->h_cg->iters;
      sum->cg_iters[it] = cg_it;
    }
    if (pending) {
      // scalars of the linearisation that followed the previous accepted step arrived with this read-back
      if (!absorb_pending()) { --it; sum->termination = GLBA_TERM_FAILURE; sum->stop_reason = GLBA_STOP_NUMERIC; break; }
      sum->cost[it - 1] = cost; sum->gradient_max_norm[it - 1] = gmax;
      if (!g2o && gmax <= o->gradient_tol) { --it; sum->termination = GLBA_TERM_CONVERGENCE; sum->stop_reason = GLBA_STOP_GRADIENT_TOL; break; }
    }
    const bool solver_ok = (S[S_NOTPD_P] + (ctx->n_free_cam > 0 ? S[S_NOTPD_C] : 0.0)) == 0.0;
    const double model_cost_change = 0.5 * ((S[S_YG_P] + S[S_YG_C]) + (S[S_YLY_P] + S[S_YLY_C]));
    const bool valid = g2o || (solver_ok && (model_cost_change > 0.0));       // g2o: a failed solve is just a rejected trial
    if (o->verbose) fprintf(stderr, "[glba] it %d cost %.9e cand %.9e model %.3e radius %.3e cg %d\n", it, cost, S[S_COST_C], model_cost_change, radius, cg_it);
    if (!valid) {
      ++n_invalid;
      sum->cost[it] = cost; sum->cost_candidate[it] = cost; sum->step_norm[it] = 0; sum->relative_decrease[it] = 0;
      sum->gradient_max_norm[it] = gmax; sum->accepted[it] = 0;
      if (n_invalid >= o->max_consecutive_invalid_steps) { sum->radius[it] = radius; sum->termination = GLBA_TERM_FAILURE; sum->stop_reason = GLBA_STOP_INVALID_STEPS; break; }
      radius = radius / decrease_factor; decrease_factor *= 2.0; fresh = false;
      sum->radius[it] = radius;
      continue;
    }
    n_invalid = 0;
    double cand = S[S_COST_C];
    if (S[S_BAD_C] > 0.0 || !std::isfinite(cand) || !solver_ok) cand = std::numeric_limits<double>::max();
    const double step_norm = std::sqrt(S[S_YN2_P] + S[S_YN2_C]);
    sum->cost_candidate[it] = cand; sum->step_norm[it] = step_norm; sum->cost[it] = cost; sum->radius[it] = radius; sum->gradient_max_norm[it] = gmax;
    if (!g2o && step_norm <= o->parameter_tol * (x_norm + o->parameter_tol)) { sum->termination = GLBA_TERM_CONVERGENCE; sum->stop_reason = GLBA_STOP_PARAMETER_TOL; break; }
    const double cost_change = cost - cand;
    if (!g2o && std::fabs(cost_change) <= o->function_tol * cost) { sum->termination = GLBA_TERM_CONVERGENCE; sum->stop_reason = GLBA_STOP_FUNCTION_TOL; break; }
    // g2o: rho = (chi2 - chi2') / (d'(lambda d + b) + 1e-3), here in half-chi2 units
    const double rel = (cand >= std::numeric_limits<double>::max()) ? (g2o ? -1.0 : std::numeric_limits<double>::lowest())
                                                                     : cost_change / (model_cost_change + (g2o ? 0.5e-3 : 0.0));
    sum->relative_decrease[it] = rel;
    if (rel > o->min_relative_decrease) {
      ctx->cur ^= 1;     // the candidate buffers (state + camera table) become current
      const double tq = 2.0 * rel - 1.0;
      double shrink = 1.0 - tq * tq * tq;          // (the device-resident loop, glba_lm.cuh, cubes the same way)
      if (g2o) shrink = std::min(shrink, 2.0 / 3.0);                 // _goodStepUpperScale
      radius = radius / std::max(1.0 / 3.0, shrink);
      radius = std::min(o->max_radius, radius);
      decrease_factor = 2.0; n_rejected = 0;
      if ((st = do_linearize_impl(ctx, o, 0, radius, !dense))) return st;    // enqueued only: read back with the next step
      pending = true;
      cost = cand;       // provisional (the re-evaluated value replaces it at the next read-back)
      sum->n_linearizations++; sum->n_successful++; sum->accepted[it] = 1;
      fresh = true;
    } else {
      radius = radius / decrease_factor; decrease_factor *= 2.0; fresh = false;
      sum->accepted[it] = 0;
      ++n_rejected;
    }
    sum->cost[it] = cost; sum->radius[it] = radius; sum->gradient_max_norm[it] = gmax;
    // g2o "Terminate": trials exhausted, a trial with rho == 0, or lambda no longer finite
    if (g2o && !sum->accepted[it] && (n_rejected >= o->g2o_max_trials || rel == 0.0 || !(radius > 0.0))) { sum->stop_reason = GLBA_STOP_TRIALS; break; }
  }
  if (pending) {           // the loop ended right after an accepted step: fetch the scalars of its re-linearisation
    if ((st = fetch_scal(ctx))) return st;
    if (!absorb_pending()) { sum->termination = GLBA_TERM_FAILURE; sum->stop_reason = GLBA_STOP_NUMERIC; }
    else { sum->cost[it] = cost; sum->gradient_max_norm[it] = gmax; }
  }
  if (!on_device) { sum->n_iters = it; sum->final_cost = cost; }
  sum->t_setup_ms = ctx->t_phase[PH_SETUP]; sum->t_linearize_ms = ctx->t_phase[PH_LIN]; sum->t_schur_ms = ctx->t_phase[PH_SCHUR];
  sum->t_solve_ms = ctx->t_phase[PH_SOLVE]; sum->t_update_ms = ctx->t_phase[PH_UPDATE];
  sum->t_comm_ms = ctx->t_phase[PH_COMM];
  sum->t_total_ms = sum->t_setup_ms + sum->t_linearize_ms + sum->t_schur_ms + sum->t_solve_ms + sum->t_update_ms + sum->t_comm_ms;
  sum->status = GLBA_OK;
  return GLBA_OK;
}

// owner layout: every camera of the whole map from its owner rank (bit patterns summed as integers: x + 0 = x exactly)
int gather_cameras(glba_ctx* ctx, const double* cam_local, Buf& out) {
  ENSURE(double, out, 6 * (size_t)ctx->n_cam_g);
  CU(cudaMemsetAsync(out.p, 0, sizeof(double) * 6 * (size_t)ctx->n_cam_g, ctx->stream));
  if (ctx->n_cam) LAUNCH(k_scatter_rows, cdiv((long)ctx->n_cam * 6, 256), 256, ctx->n_cam, (const int*)ctx->l2g.as<int>(),
                         (const uint8_t*)ctx->cam_owned.as<uint8_t>(), 1, cam_local, 6, out.as<double>());
  return allreduce(ctx, out.p, 6 * (size_t)ctx->n_cam_g, kNcclSum, kNcclInt64);
}

int write_back(glba_ctx* ctx, double* cam, double* pt, int memspace) {
  const int c = ctx->cur;
  cudaStream_t s = ctx->stream;
  const double* cam_src = ctx->cam[c].as<double>();
  if (ctx->owner && cam) {
    const int st = gather_cameras(ctx, cam_src, ctx->out_c);
    if (st) return st;
    cam_src = ctx->out_c.as<double>();
  }
  const size_t n_cam_out = ctx->owner ? (size_t)ctx->n_cam_g : (size_t)ctx->n_cam;
  if (memspace == GLBA_MEM_HOST) {
    ENSURE(double, ctx->out_a, 3 * (size_t)ctx->n_pt);
    if (ctx->n_pt) LAUNCH(k_unpack_pt, cdiv(ctx->n_pt, 256), 256, ctx->n_pt, (const double4*)ctx->pt4[c].as<double4>(), ctx->relabelled ? (const int*)ctx->new2old.as<int>() : nullptr, ctx->out_a.as<double>());
    if (cam) CU(cudaMemcpyAsync(cam, cam_src, sizeof(double) * 6 * n_cam_out, cudaMemcpyDeviceToHost, s));
    if (pt) CU(cudaMemcpyAsync(pt, ctx->out_a.p, sizeof(double) * 3 * ctx->n_pt, cudaMemcpyDeviceToHost, s));
  } else {
    if (cam) CU(cudaMemcpyAsync(cam, cam_src, sizeof(double) * 6 * n_cam_out, cudaMemcpyDeviceToDevice, s));
    if (pt && ctx->n_pt) LAUNCH(k_unpack_pt, cdiv(ctx->n_pt, 256), 256, ctx->n_pt, (const double4*)ctx->pt4[c].as<double4>(), ctx->relabelled ? (const int*)ctx->new2old.as<int>() : nullptr, pt);
  }
  CU(cudaStreamSynchronize(s));
  return GLBA_OK;
}

// Sharded maps that will be solved by PCG (more cameras than the dense path takes, or PCG requested) use the owner-computes
// layout; small sharded windows keep the replicated layout of the exact dense solve.  Every rank decides alike.
bool wants_owner_layout(const glba_ctx* ctx, const glba_problem* p, const glba_options* o) {
  return ctx->world > 1 && p && (p->n_cam > DN_MAXCAM || o->linsolve == GLBA_LINSOLVE_PCG);
}

PoseOpts pose_opts(const glba_options* o) {
  PoseOpts P;
  P.loss = LossP{o->loss, o->loss_scale}; P.max_iters = o->max_iters;
  P.function_tol = o->function_tol; P.gradient_tol = o->gradient_tol; P.parameter_tol = o->parameter_tol;
  P.initial_radius = o->initial_radius; P.max_radius = o->max_radius; P.min_radius = o->min_radius;
  P.min_relative_decrease = o->min_relative_decrease; P.min_diag = o->min_lm_diagonal; P.max_diag = o->max_lm_diagonal;
  P.jacobi = o->jacobi_scaling; P.max_invalid = o->max_consecutive_invalid_steps;
  return P;
}

}  // namespace

// ---------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------
extern "C" {

void glba_default_options(glba_options* o) {
  if (!o) return;
  std::memset(o, 0, sizeof(*o));
  o->loss = GLBA_LOSS_CAUCHY; o->loss_scale = 1.0; o->max_iters = 30;
  o->function_tol = 1e-6; o->gradient_tol = 1e-10; o->parameter_tol = 1e-8;
  o->initial_radius = 1e4; o->max_radius = 1e16; o->min_radius = 1e-32;
  o->min_relative_decrease = 1e-3; o->min_lm_diagonal = 1e-6; o->max_lm_diagonal = 1e32;
  o->jacobi_scaling = 1; o->max_consecutive_invalid_steps = 5;
  o->linsolve = GLBA_LINSOLVE_AUTO; o->dense_max_dim = 6 * DN_MAXCAM; o->cg_rel_tol = 1e-13; o->cg_max_iters = 0; o->verbose = 0;
  o->mode = GLBA_MODE_CERES; o->g2o_tau = 1e-5; o->g2o_max_trials = 10;
}

const char* glba_strerror(int status) {
  switch (status) {
    case GLBA_OK: return "ok";
    case GLBA_E_INVALID_ARG: return "invalid argument";
    case GLBA_E_CUDA: return "CUDA runtime error";
    case GLBA_E_NO_DEVICE: return "no CUDA device (this backend has no CPU fallback)";
    case GLBA_E_OOM: return "out of device memory";
    case GLBA_E_NCCL: return "NCCL error";
    case GLBA_E_NUMERIC: return "non-finite residuals at the initial point";
    case GLBA_E_UNSUPPORTED: return "unsupported";
    default: return "unknown status";
  }
}

const char* glba_last_error(const glba_ctx* ctx) { return ctx ? ctx->err.c_str() : ""; }
int glba_version(void) { return GLBA_VERSION; }
int64_t glba_kernel_launch_count(void) { return (int64_t)g_launches.load(); }

int glba_nccl_unique_id(void* out_id) {
  if (!out_id) return GLBA_E_INVALID_ARG;
  if (!g_nccl.load()) return GLBA_E_NCCL;
  ncclUniqueId id;
  if (g_nccl.GetUniqueId(&id) != 0) return GLBA_E_NCCL;
  static_assert(sizeof(id) == GLBA_NCCL_ID_BYTES, "ncclUniqueId size");
  std::memcpy(out_id, &id, sizeof(id));
  return GLBA_OK;
}

int glba_create(const glba_device_cfg* cfg, glba_ctx** out) {
  if (!cfg || !out) return GLBA_E_INVALID_ARG;
  *out = nullptr;
  if (cfg->world < 1 || cfg->world > MAX_WORLD || cfg->rank < 0 || cfg->rank >= cfg->world) return GLBA_E_INVALID_ARG;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); return GLBA_E_NO_DEVICE; }
  if (cfg->device < 0 || cfg->device >= ndev) return GLBA_E_INVALID_ARG;
  if (cudaSetDevice(cfg->device) != cudaSuccess) return GLBA_E_CUDA;
  glba_ctx* ctx = new glba_ctx();
  ctx->device = cfg->device; ctx->rank = cfg->rank; ctx->world = cfg->world;
  { int v = 0; if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, cfg->device) == cudaSuccess && v > 0) ctx->n_sm = v; }
  if (set_func_attributes(ctx) != GLBA_OK) { delete ctx; return GLBA_E_CUDA; }
  if (const char* e = std::getenv("GLBA_TIMING")) ctx->env_timing = (e[0] == '1');          // diagnostic: phase timings for small problems too
  if (const char* e = std::getenv("GLBA_RELABEL")) ctx->env_relabel = (e[0] != '0');   // diagnostic: GLBA_RELABEL=0 keeps the caller's point order
  g_pdl_max_grid = 2u * (unsigned)ctx->n_sm;
  if (const char* e = std::getenv("GLBA_PDL")) { g_pdl = (e[0] != '0'); if (e[0] == '2') g_pdl_max_grid = 0x7fffffffu; }   // diagnostic: 0 = plain launches, 2 = every launch
  if (const char* e = std::getenv("GLBA_HOST_LM")) ctx->env_host_lm = (e[0] == '1');   // diagnostic: host-side accept/reject for small windows
  if (const char* e = std::getenv("GLBA_EXPLICIT")) ctx->env_explicit = (e[0] != '0'); // diagnostic: GLBA_EXPLICIT=0 = matrix-free product on every map
  if (const char* e = std::getenv("GLBA_CG_PROF")) ctx->env_cg_prof = (e[0] == '1');
  if (const char* e = std::getenv("GLBA_CG_REG")) ctx->env_cg_reg = (e[0] != '0');     // diagnostic: general PCG kernel on every map
  if (const char* e = std::getenv("GLBA_FUSED")) ctx->env_fused = (e[0] == '1');       // experiment: GLBA_FUSED=1 = both halves of the implicit product in one tile kernel
  if (const char* e = std::getenv("GLBA_CP_OCC")) ctx->env_cp_occ = std::atoi(e);
  if (const char* e = std::getenv("GLBA_P2P")) ctx->env_p2p = (e[0] == '1');           // opt-in: peer-memory form of the compact exchange
  if (const char* e = std::getenv("GLBA_CAMPIPE")) ctx->env_campipe = (e[0] != '0');   // diagnostic: separate k_linearize_cm / k_schur_cm on large maps
  if (const char* e = std::getenv("GLBA_PIPE")) ctx->env_pipe = (e[0] != '0');         // diagnostic: GLBA_PIPE=0 runs the round-1 tile kernels on large maps
  if (const char* e = std::getenv("GLBA_TILE")) ctx->env_force_large = (e[0] == 'l');  // diagnostic: GLBA_TILE=large = large-map tiles for any size
  if (cfg->stream) { ctx->stream = (cudaStream_t)cfg->stream; ctx->own_stream = false; }
  else { if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) { delete ctx; return GLBA_E_CUDA; } ctx->own_stream = true; }
  if (cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking) != cudaSuccess) { glba_destroy(ctx); return GLBA_E_CUDA; }
  if (cudaStreamCreateWithFlags(&ctx->side_stream, cudaStreamNonBlocking) != cudaSuccess) { glba_destroy(ctx); return GLBA_E_CUDA; }
  for (auto& e : ctx->ev_side) if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) { glba_destroy(ctx); return GLBA_E_CUDA; }
  for (auto& e : ctx->ev_copy) if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) { glba_destroy(ctx); return GLBA_E_CUDA; }
  if (cudaMallocHost((void**)&ctx->h_scal, sizeof(double) * NSCAL) != cudaSuccess || cudaMallocHost((void**)&ctx->h_cg, sizeof(CgState)) != cudaSuccess ||
      cudaMallocHost((void**)&ctx->h_flags, 16 * sizeof(int)) != cudaSuccess) { glba_destroy(ctx); return GLBA_E_CUDA; }
  if (cfg->world > 1) {
    if (!cfg->nccl_unique_id || !g_nccl.load()) { glba_destroy(ctx); return GLBA_E_NCCL; }
    ncclUniqueId id; std::memcpy(&id, cfg->nccl_unique_id, sizeof(id));
    if (g_nccl.CommInitRank(&ctx->comm, cfg->world, id, cfg->rank) != 0) { ctx->comm = nullptr; glba_destroy(ctx); return GLBA_E_NCCL; }
    if (setup_p2p(ctx) != GLBA_OK) { glba_destroy(ctx); return GLBA_E_NCCL; }
  }
  *out = ctx;
  return GLBA_OK;
}

void glba_destroy(glba_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  if (ctx->stream) cudaStreamSynchronize(ctx->stream);
  if (ctx->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(ctx->comm);
  Buf* all[] = {&ctx->in_cam, &ctx->in_pt, &ctx->in_ocam, &ctx->in_opt, &ctx->in_u, &ctx->in_v, &ctx->in_cfix, &ctx->in_pfix, &ctx->in_info, &ctx->pm_cam, &ctx->pm_pt,
                &ctx->pm_uv, &ctx->pm2orig, &ctx->pm2cm, &ctx->pt_start, &ctx->cm_pt, &ctx->cm_uv, &ctx->cm2pm, &ctx->cam_start, &ctx->chunk_cam,
                &ctx->chunk_begin, &ctx->chunk_end, &ctx->cam_chunk_start, &ctx->cam_free, &ctx->pt_free, &ctx->sort_tmp, &ctx->keys_tmp, &ctx->flags,
                &ctx->cam[0], &ctx->cam[1], &ctx->camtab[0], &ctx->camtab[1], &ctx->pt4[0], &ctx->pt4[1], &ctx->cam0, &ctx->pt40, &ctx->rec_pm, &ctx->rec_cm,
                &ctx->Craw, &ctx->sp4, &ctx->lam4, &ctx->cinv, &ctx->u0p, &ctx->u4, &ctx->part_pm, &ctx->part_cm, &ctx->acc27, &ctx->yhat, &ctx->Bc, &ctx->gc, &ctx->sc,
                &ctx->lamc, &ctx->Md, &ctx->Minv, &ctx->rhs, &ctx->cg_x, &ctx->cg_r, &ctx->cg_p, &ctx->cg_q, &ctx->pg, &ctx->yg, &ctx->cgst,
                &ctx->out_a, &ctx->out_b, &ctx->out_c, &ctx->dn_part, &ctx->dn_red, &ctx->dn_full, &ctx->tile_pt, &ctx->xtab, &ctx->partA, &ctx->partB, &ctx->partc, &ctx->counters, &ctx->cam_cnt, &ctx->part_cm2, &ctx->part_pm2, &ctx->act_mask, &ctx->g2l, &ctx->l2g, &ctx->cam_owned, &ctx->cam_shared, &ctx->sh_scan, &ctx->xsend, &ctx->xrecv, &ctx->xsend6, &ctx->ocam_loc, &ctx->cam_loc, &ctx->cfix_loc, &ctx->late, &ctx->sh2loc, &ctx->lmctl, &ctx->dsum, &ctx->tile_cmin, &ctx->tile_desc, &ctx->tile_cams, &ctx->pm_slot, &ctx->tile_sobs, &ctx->tile_sstart, &ctx->tp_key, &ctx->tp_val, &ctx->tp_key2, &ctx->cam_tp, &ctx->cam_tp_start, &ctx->tpart, &ctx->cam_iota, &ctx->ovf_raw, &ctx->ovf_k, &ctx->ovf_c, &ctx->ovf_key, &ctx->ovf_key2, &ctx->ovf_val, &ctx->cam_ov, &ctx->cam_ov_start, &ctx->first_cam, &ctx->new2old, &ctx->old2new, &ctx->opt_relab, &ctx->hmax,
                &ctx->sp_cnt, &ctx->sp_off, &ctx->sp_key, &ctx->sp_key2, &ctx->sp_val, &ctx->sp_inst, &ctx->sp_ukey, &ctx->sp_ucnt, &ctx->sp_nruns, &ctx->sp_pair_a,
                &ctx->sp_pair_b, &ctx->sp_pair_start, &ctx->sp_ekey, &ctx->sp_ekey2, &ctx->sp_eval, &ctx->sp_eval2, &ctx->sp_erow, &ctx->sp_ent, &ctx->sp_row_start,
                &ctx->sp_blocks, &ctx->sp_part, &ctx->sp_bar, &ctx->sp_pres, &ctx->sp_gscan, &ctx->sp_gkey, &ctx->sp_gid, &ctx->sp_vec, &ctx->sp_prof, &ctx->sp_cta_rows};
  for (Buf* b : all) release(*b);
  for (cudaEvent_t e : ctx->ev) cudaEventDestroy(e);
  if (ctx->h_scal) cudaFreeHost(ctx->h_scal);
  if (ctx->h_cg) cudaFreeHost(ctx->h_cg);
  if (ctx->h_flags) cudaFreeHost(ctx->h_flags);
  teardown_p2p(ctx);
  if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
  if (ctx->side_stream) { cudaStreamSynchronize(ctx->side_stream); cudaStreamDestroy(ctx->side_stream); }
  for (auto& e : ctx->ev_side) if (e) cudaEventDestroy(e);
  for (auto& e : ctx->ev_copy) if (e) cudaEventDestroy(e);
  if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
}

void* glba_stream(glba_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }

int glba_synchronize(glba_ctx* ctx) {
  if (!ctx) return GLBA_E_INVALID_ARG;
  CU(cudaStreamSynchronize(ctx->stream));
  return GLBA_OK;
}

int glba_load(glba_ctx* ctx, const glba_problem* prob, const glba_options* opt) {
  if (!ctx) return GLBA_E_INVALID_ARG;
  CU(cudaSetDevice(ctx->device));
  int st = validate_options(ctx, opt);
  if (st) return st;
  ctx->t_phase[PH_SETUP] = 0.0;
  ctx->mode = opt->mode;
  ctx->want_owner = wants_owner_layout(ctx, prob, opt);
  st = load_problem(ctx, prob);
  if (st) return st;
  CU(cudaStreamSynchronize(ctx->stream));
  collect(ctx);
  return GLBA_OK;
}

int glba_reset_resident(glba_ctx* ctx) {
  if (!ctx || !ctx->loaded) return GLBA_E_INVALID_ARG;
  CU(cudaSetDevice(ctx->device));
  ctx->cur = 0;
  CU(cudaMemcpyAsync(ctx->cam[0].p, ctx->cam0.p, sizeof(double) * 6 * ctx->n_cam, cudaMemcpyDeviceToDevice, ctx->stream));
  CU(cudaMemcpyAsync(ctx->pt4[0].p, ctx->pt40.p, sizeof(double4) * ctx->n_pt, cudaMemcpyDeviceToDevice, ctx->stream));
  if (ctx->n_cam) LAUNCH(k_cam_prep, cdiv(ctx->n_cam, 128), 128, ctx->n_cam, (const double*)ctx->cam[0].as<double>(), ctx->camtab[0].as<double>(), ctx->mode);
  return GLBA_OK;
}

int glba_read_resident(glba_ctx* ctx, double* cam, double* pt) {
  if (!ctx || !ctx->loaded) return GLBA_E_INVALID_ARG;
  CU(cudaSetDevice(ctx->device));
  return write_back(ctx, cam, pt, GLBA_MEM_HOST);
}

int glba_linearize_resident(glba_ctx* ctx, const glba_options* opt, double radius, double* cost) {
  if (!ctx || !ctx->loaded || !(radius > 0.0)) return GLBA_E_INVALID_ARG;
  CU(cudaSetDevice(ctx->device));
  int st = validate_options(ctx, opt);
  if (st) return st;
  if ((st = do_linearize_schur(ctx, opt, 1, radius))) return st;
  if (cost) {
    if ((st = fetch_scal(ctx))) return st;
    *cost = ctx->h_scal[S_COST];
  } else {
    ctx->ev_used = 0;   // no sync requested: drop the phase markers
  }
  return GLBA_OK;
}

int glba_time_kernels(glba_ctx* ctx, const glba_options* opt, double radius, int32_t reps, glba_kernel_times* out) {
  if (!ctx || !ctx->loaded || !out || reps <= 0 || !(radius > 0.0)) return GLBA_E_INVALID_ARG;
  CU(cudaSetDevice(ctx->device));
  int st = validate_options(ctx, opt);
  if (st) return st;
  std::memset(out, 0, sizeof(*out));
  const int c = ctx->cur;
  const int n_cam = ctx->n_cam, n_pt = ctx->n_pt;
  if (n_pt == 0 || ctx->n_chunks == 0) return GLBA_OK;
  // a complete pass first so every buffer the kernels read is valid
  if ((st = do_linearize_schur(ctx, opt, 1, radius))) return st;
  CgState* cg = ctx->cgst.as<CgState>();
  LAUNCH(k_cg_start, ctx->grid_c, NT_C, n_cam, (const double*)ctx->camtab[c].as<double>(), (const double*)ctx->Minv.as<double>(),
         (const double*)ctx->rhs.as<double>(), ctx->cg_x.as<double>(), ctx->cg_r.as<double>(), ctx->cg_q.as<double>(), ctx->cg_p.as<double>(),
         ctx->pg.as<double>(), ctx->xtab.as<double>(), cg, 0.0, 1 << 30);
  CU(cudaStreamSynchronize(ctx->stream));
  ctx->ev_used = 0;
  EventPair ev;
  cudaEvent_t e0 = ev.a, e1 = ev.b;
  const CmArgs CA = cm_args(ctx);
  auto timed = [&](auto&& launch, double* ms_out) -> int {
    launch();                                   // warm
    cudaEventRecord(e0, ctx->stream);
    for (int r = 0; r < reps; ++r) launch();
    cudaEventRecord(e1, ctx->stream);
    if (cudaEventSynchronize(e1) != cudaSuccess) return GLBA_E_CUDA;
    float ms = 0.f; cudaEventElapsedTime(&ms, e0, e1);
    *ms_out = ms / reps;
    return GLBA_OK;
  };
  if ((st = timed([&] { (void)launch_linearize_points(ctx, opt, 0, radius); }, &out->linearize_pm_ms))) return st;
  if ((st = timed([&] { LAUNCH(k_linearize_cm, ctx->n_chunks, NT_HCM, CA, (const double4*)ctx->rec_cm.as<double4>(), (const double*)ctx->camtab[c].as<double>(),
           ctx->part_cm.as<double>(), (const LmCtl*)nullptr); }, &out->linearize_cm_ms))) return st;
  if ((st = timed([&] { LAUNCH(k_schur_cm, ctx->n_chunks, NT_HCM, CA, (const double4*)ctx->rec_cm.as<double4>(), (const double*)ctx->camtab[c].as<double>(),
           (const double*)ctx->cinv.as<double>(), (const double4*)ctx->u0p.as<double4>(), ctx->part_cm.as<double>()); }, &out->schur_cm_ms))) return st;
  if (ctx->use_pipe && ctx->env_campipe)
    if ((st = timed([&] { LAUNCH_SMEM(ctx->env_cp_occ == 3 ? k_cam_pipe<3> : k_cam_pipe<2>, ctx->n_chunks, CP_NT, sizeof(CamSmem), CA, (const double4*)ctx->rec_cm.as<double4>(),
             (const double*)ctx->camtab[c].as<double>(), (const double*)ctx->cinv.as<double>(), (const double4*)ctx->u0p.as<double4>(),
             ctx->part_cm.as<double>(), ctx->part_cm2.as<double>()); }, &out->cam_pipe_ms))) return st;
  if ((st = timed([&] { launch_point_pass0(ctx, opt, (const CgState*)nullptr, 0); }, &out->spmv_pm_ms))) return st;
  if ((st = timed([&] { LAUNCH(k_spmv_cm, ctx->n_chunks, NT_CM, CA, (const double4*)ctx->rec_cm.as<double4>(), (const double*)ctx->camtab[c].as<double>(),
           (const double4*)ctx->u4.as<double4>(), (const CgState*)nullptr, 0, ctx->part_cm.as<double>()); }, &out->spmv_cm_ms))) return st;
  // candidate = current state + PCG start vector (any finite step exercises the same code)
  const int d = c ^ 1;
  LAUNCH(k_cam_step2, ctx->grid_c, NT_C, n_cam, (const uint8_t*)ctx->cam_free.as<uint8_t>(), (const double*)ctx->cam[c].as<double>(),
         (const double*)ctx->camtab[c].as<double>(), (const double*)ctx->cg_x.as<double>(), (const double*)ctx->gc.as<double>(),
         (const double*)ctx->lamc.as<double>(), 1.0 / radius, ctx->cam[d].as<double>(), ctx->camtab[d].as<double>(), ctx->xtab.as<double>(),
         ctx->partc.as<double>(), ctx->counters.as<unsigned>() + 3, ctx->d_scal, ctx->mode, (const LmCtl*)nullptr, ctx->owner ? (const uint8_t*)ctx->cam_owned.as<uint8_t>() : (const uint8_t*)nullptr);
  if ((st = timed([&] { launch_point_pass1(ctx, opt, radius); }, &out->backsub_cost_ms))) return st;
  if ((st = timed([&] { LAUNCH(k_point_damp, cdiv(n_pt, NT_PM), NT_PM, n_pt, (const uint8_t*)ctx->pt_free.as<uint8_t>(), (const double*)ctx->Craw.as<double>(),
           (const double4*)ctx->lam4.as<double4>(), ctx->cinv.as<double>(), ctx->u0p.as<double4>(), 1.0 / radius, ctx->part_pm.as<double>(), (const LmCtl*)nullptr); }, &out->point_damp_ms))) return st;
  // every camera-sized kernel of one linearise + Schur pass (the scalar reductions now run inside the tile kernels)
  st = timed([&] {
    launch_cam_lin_fin(ctx, opt, 0);
    launch_cam_schur_fin(ctx, radius, ctx->part_cm2.as<double>()); }, &out->small_kernels_ms);
  if (st) return st;
  {
    // assembled reduced camera matrix: structure (once per load), assembly (sharded: incl. its all-reduce), one PCG iteration.
    // Collective on a sharded map, like everything below.
    if (!ctx->sp_tried) {
      cudaEventRecord(e0, ctx->stream);
      if ((st = ensure_explicit(ctx, opt))) return st;
      cudaEventRecord(e1, ctx->stream);
      if (cudaEventSynchronize(e1) != cudaSuccess) return GLBA_E_CUDA;
      float ms = 0.f; cudaEventElapsedTime(&ms, e0, e1);
      out->pair_setup_ms = ms;
    }
    if (ctx->use_explicit) {
      // the timed launches above left Schur sums where the linearisation's belong: a consistent pass first
      if ((st = do_linearize_schur(ctx, opt, 0, radius))) return st;
      CU(cudaStreamSynchronize(ctx->stream));
      ctx->ev_used = 0;
      if ((st = timed([&] { (void)launch_schur_pairs(ctx); }, &out->schur_pairs_ms))) return st;
      // one PCG iteration on the blocks: (time of 1 + 32 iterations - time of 1 iteration) / 32, tolerance 0 so that none stops early
      glba_options ot = *opt; ot.cg_rel_tol = 0.0;
      double t1 = 0.0, t33 = 0.0;
      if ((st = timed([&] { (void)launch_cg_bsr(ctx, &ot, 1); }, &t1))) return st;
      if ((st = timed([&] { (void)launch_cg_bsr(ctx, &ot, 33); }, &t33))) return st;
      out->bsr_spmv_ms = (t33 - t1) / 32.0;
      out->n_pair_instances = ctx->n_inst; out->n_pair_blocks = ctx->n_pairs_g;
    }
  }
  if (ctx->world == 1) return GLBA_OK;
  // collective: every rank calls glba_time_kernels with the same reps
  out->n_local_cams = n_cam; out->n_shared_cams = ctx->owner ? ctx->n_shared : ctx->n_cam;
  out->exchange_bytes = 8.0 * ((ctx->owner ? 54.0 * ctx->n_shared : 54.0 * n_cam) + S_GSLOT0 + MAX_WORLD);
  if (ctx->owner) {
    st = timed([&] { (void)owner_exchange(ctx, true, true, S_GSLOT0 + MAX_WORLD); }, &out->allreduce_ms);
    if (st) return st;
  } else if ((st = timed([&] { (void)allreduce(ctx, ctx->d_accB, 54 * (size_t)n_cam + S_GSLOT0 + MAX_WORLD, kNcclSum); }, &out->allreduce_ms))) return st;
  st = timed([&] {
    LAUNCH(k_chunk_sum_lin, cdiv((long)n_cam * 54, 256), 256, n_cam, (const int*)ctx->cam_chunk_start.as<int>(), (const double*)ctx->part_cm.as<double>(),
           (const double*)ctx->part_cm2.as<double>(), ctx->d_accA, ctx->d_accB, ctx->rank, ctx->d_scal); }, &out->chunk_sum_ms);
  return st;
}

int glba_solve_resident(glba_ctx* ctx, const glba_options* opt, glba_summary* summary) {
  if (!ctx || !ctx->loaded || !summary) return GLBA_E_INVALID_ARG;
  CU(cudaSetDevice(ctx->device));
  int st = validate_options(ctx, opt);
  if (st) return st;
  std::memset(summary, 0, sizeof(*summary));
  st = run_lm(ctx, opt, summary);
  summary->status = st;
  return st;
}

int glba_solve(glba_ctx* ctx, const glba_problem* prob, const glba_options* opt, glba_summary* summary) {
  if (!ctx || !summary) return GLBA_E_INVALID_ARG;
  std::memset(summary, 0, sizeof(*summary));
  CU(cudaSetDevice(ctx->device));
  int st = validate_options(ctx, opt);
  if (st) { summary->status = st; return st; }
  ctx->t_phase[PH_SETUP] = 0.0;
  ctx->mode = opt->mode;
  ctx->want_owner = wants_owner_layout(ctx, prob, opt);
  st = load_problem(ctx, prob);
  if (st) { summary->status = st; return st; }
  st = run_lm(ctx, opt, summary);
  summary->status = st;
  if (st) return st;
  if (summary->termination != GLBA_TERM_FAILURE) {
    st = write_back(ctx, prob->cam, prob->pt, prob->memspace);
    if (st) { summary->status = st; return st; }
  }
  return GLBA_OK;
}

int glba_linearize(glba_ctx* ctx, const glba_problem* prob, const glba_options* opt, double radius, glba_linearization* out) {
  if (!ctx || !out || !(radius > 0.0)) return GLBA_E_INVALID_ARG;
  CU(cudaSetDevice(ctx->device));
  int st = validate_options(ctx, opt);
  if (st) return st;
  if (opt->mode != GLBA_MODE_CERES) return fail(ctx, GLBA_E_UNSUPPORTED, "glba_linearize reports blocks in the CERES formulation only");
  ctx->mode = GLBA_MODE_CERES;
  ctx->want_owner = wants_owner_layout(ctx, prob, opt);
  if ((st = load_problem(ctx, prob))) return st;
  for (int q = 0; q < PH_COUNT; ++q) ctx->t_phase[q] = 0.0;
  if ((st = do_linearize_schur(ctx, opt, 1, radius))) return st;
  if ((st = fetch_scal(ctx))) return st;
  out->cost = ctx->h_scal[S_COST];
  out->t_linearize_ms = ctx->t_phase[PH_LIN]; out->t_schur_ms = ctx->t_phase[PH_SCHUR];
  const int n_cam = ctx->n_cam, n_pt = ctx->n_pt; const long n = ctx->n_obs;
  cudaStream_t s = ctx->stream;
  const int c = ctx->cur;
  if ((out->residuals || out->jac_cam || out->jac_pt) && n > 0) {
    ENSURE(double, ctx->out_a, 2 * (size_t)n); ENSURE(double, ctx->out_b, 12 * (size_t)n); ENSURE(double, ctx->out_c, 6 * (size_t)n);
    LAUNCH(k_expand, cdiv(n, 256), 256, n, (const int*)ctx->pm_cam.as<int>(), (const double2*)ctx->pm_uv.as<double2>(),
           ctx->sorted_input ? (const int*)nullptr : (const int*)ctx->pm2orig.as<int>(), (const double4*)ctx->rec_pm.as<double4>(),
           (const double*)ctx->camtab[c].as<double>(), ctx->K, ctx->out_a.as<double>(), ctx->out_b.as<double>(), ctx->out_c.as<double>());
    if (out->residuals) CU(cudaMemcpyAsync(out->residuals, ctx->out_a.p, sizeof(double) * 2 * n, cudaMemcpyDeviceToHost, s));
    if (out->jac_cam) CU(cudaMemcpyAsync(out->jac_cam, ctx->out_b.p, sizeof(double) * 12 * n, cudaMemcpyDeviceToHost, s));
    if (out->jac_pt) CU(cudaMemcpyAsync(out->jac_pt, ctx->out_c.p, sizeof(double) * 6 * n, cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
  }
  if ((out->hess_pt || out->grad_pt) && n_pt > 0) {
    ENSURE(double, ctx->out_b, 9 * (size_t)n_pt); ENSURE(double, ctx->out_c, 3 * (size_t)n_pt);
    LAUNCH(k_unpack_pointblocks, cdiv(n_pt, 256), 256, n_pt, (const uint8_t*)ctx->pt_free.as<uint8_t>(), (const double*)ctx->Craw.as<double>(),
           ctx->relabelled ? (const int*)ctx->new2old.as<int>() : nullptr, ctx->out_b.as<double>(), ctx->out_c.as<double>());
    if (out->hess_pt) CU(cudaMemcpyAsync(out->hess_pt, ctx->out_b.p, sizeof(double) * 9 * n_pt, cudaMemcpyDeviceToHost, s));
    if (out->grad_pt) CU(cudaMemcpyAsync(out->grad_pt, ctx->out_c.p, sizeof(double) * 3 * n_pt, cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
  }
  if (ctx->owner) {
    // owner layout: the caller's per-camera arrays cover the whole map; this rank fills the rows of the cameras it observes
    // (complete values: shared cameras were exchanged), the others are zero
    const int ng = ctx->n_cam_g;
    struct { double* host; const Buf* src; int width; } outs[4] = {{out->grad_cam, &ctx->gc, 6}, {out->hess_cam, &ctx->Bc, 36},
                                                                  {out->schur_diag, &ctx->Md, 36}, {out->schur_rhs, &ctx->rhs, 6}};
    for (auto& o4 : outs) {
      if (!o4.host) continue;
      ENSURE(double, ctx->out_b, (size_t)o4.width * ng);
      CU(cudaMemsetAsync(ctx->out_b.p, 0, sizeof(double) * (size_t)o4.width * ng, s));
      const bool have = (o4.width == 6 && o4.src == &ctx->gc) || (o4.src == &ctx->Bc) || ctx->n_free_cam > 0;
      if (n_cam && have) LAUNCH(k_scatter_rows, cdiv((long)n_cam * o4.width, 256), 256, n_cam, (const int*)ctx->l2g.as<int>(),
                                (const uint8_t*)ctx->cam_owned.as<uint8_t>(), 0, (const double*)o4.src->as<double>(), o4.width, ctx->out_b.as<double>());
      CU(cudaMemcpyAsync(o4.host, ctx->out_b.p, sizeof(double) * (size_t)o4.width * ng, cudaMemcpyDeviceToHost, s));
      CU(cudaStreamSynchronize(s));
    }
  } else if (n_cam > 0) {
    if (out->grad_cam) CU(cudaMemcpyAsync(out->grad_cam, ctx->gc.p, sizeof(double) * 6 * n_cam, cudaMemcpyDeviceToHost, s));
    if (out->hess_cam) CU(cudaMemcpyAsync(out->hess_cam, ctx->Bc.p, sizeof(double) * 36 * n_cam, cudaMemcpyDeviceToHost, s));
    if (ctx->n_free_cam > 0) {
      if (out->schur_diag) CU(cudaMemcpyAsync(out->schur_diag, ctx->Md.p, sizeof(double) * 36 * n_cam, cudaMemcpyDeviceToHost, s));
      if (out->schur_rhs) CU(cudaMemcpyAsync(out->schur_rhs, ctx->rhs.p, sizeof(double) * 6 * n_cam, cudaMemcpyDeviceToHost, s));
    } else {
      if (out->schur_diag) std::memset(out->schur_diag, 0, sizeof(double) * 36 * n_cam);
      if (out->schur_rhs) std::memset(out->schur_rhs, 0, sizeof(double) * 6 * n_cam);
    }
    CU(cudaStreamSynchronize(s));
  }
  return GLBA_OK;
}

int glba_pose_only_batch(glba_ctx* ctx, int32_t batch, double* cams, const int32_t* offset, const double* X, const double* uv,
                         double fx, double fy, double cx, double cy, const glba_options* opt, uint8_t* usable, int32_t* n_iters,
                         double* final_cost) {
  if (!ctx || batch <= 0 || !cams || !offset || !X || !uv) return fail(ctx, GLBA_E_INVALID_ARG, "pose_only_batch: bad argument");
  CU(cudaSetDevice(ctx->device));
  int st = validate_options(ctx, opt);
  if (st) return st;
  const long n = offset[batch];
  for (int b = 0; b < batch; ++b) if (offset[b + 1] <= offset[b]) return fail(ctx, GLBA_E_INVALID_ARG, "pose_only_batch: empty problem %d", b);
  cudaStream_t s = ctx->stream;
  ENSURE(double, ctx->in_cam, 6 * (size_t)batch); ENSURE(int, ctx->in_ocam, (size_t)batch + 1); ENSURE(double, ctx->in_pt, 3 * (size_t)n);
  ENSURE(double, ctx->in_u, 2 * (size_t)n); ENSURE(uint8_t, ctx->in_cfix, batch); ENSURE(int, ctx->in_opt, 3 * (size_t)batch); ENSURE(double, ctx->in_v, batch);
  CU(cudaMemcpyAsync(ctx->in_cam.p, cams, sizeof(double) * 6 * batch, cudaMemcpyHostToDevice, s));
  CU(cudaMemcpyAsync(ctx->in_ocam.p, offset, sizeof(int) * ((size_t)batch + 1), cudaMemcpyHostToDevice, s));
  CU(cudaMemcpyAsync(ctx->in_pt.p, X, sizeof(double) * 3 * n, cudaMemcpyHostToDevice, s));
  CU(cudaMemcpyAsync(ctx->in_u.p, uv, sizeof(double) * 2 * n, cudaMemcpyHostToDevice, s));
  PoseTrace tr{}; tr.cost = nullptr; tr.stride = 0;
  LAUNCH(k_pose_only, batch, NT_POSE, batch, ctx->in_cam.as<double>(), (const int*)ctx->in_ocam.as<int>(), (const double*)ctx->in_pt.as<double>(),
         (const double*)ctx->in_u.as<double>(), Intr{fx, fy, cx, cy}, pose_opts(opt), ctx->in_cfix.as<uint8_t>(), ctx->in_opt.as<int>(),
         ctx->in_v.as<double>(), ctx->in_opt.as<int>() + batch, tr);
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(cams, ctx->in_cam.p, sizeof(double) * 6 * batch, cudaMemcpyDeviceToHost, s));
  if (usable) CU(cudaMemcpyAsync(usable, ctx->in_cfix.p, batch, cudaMemcpyDeviceToHost, s));
  if (n_iters) CU(cudaMemcpyAsync(n_iters, ctx->in_opt.p, sizeof(int) * batch, cudaMemcpyDeviceToHost, s));
  if (final_cost) CU(cudaMemcpyAsync(final_cost, ctx->in_v.p, sizeof(double) * batch, cudaMemcpyDeviceToHost, s));
  CU(cudaStreamSynchronize(s));
  return GLBA_OK;
}

int glba_pose_only(glba_ctx* ctx, double* cam, int32_t n, const double* X, const double* uv, double fx, double fy, double cx, double cy,
                   const glba_options* opt, glba_summary* summary) {
  if (!ctx || !cam || n <= 0 || !X || !uv) return fail(ctx, GLBA_E_INVALID_ARG, "pose_only: bad argument");
  CU(cudaSetDevice(ctx->device));
  int st = validate_options(ctx, opt);
  if (st) return st;
  if (summary) std::memset(summary, 0, sizeof(*summary));
  cudaStream_t s = ctx->stream;
  EventPair ev;
  cudaEvent_t e0 = ev.a, e1 = ev.b;
  const int NI = GLBA_MAX_ITERS + 1;
  ENSURE(double, ctx->in_cam, 6); ENSURE(int, ctx->in_ocam, 2); ENSURE(double, ctx->in_pt, 3 * (size_t)n); ENSURE(double, ctx->in_u, 2 * (size_t)n);
  ENSURE(uint8_t, ctx->in_cfix, 1); ENSURE(int, ctx->in_opt, 3); ENSURE(double, ctx->in_v, 1);
  ENSURE(double, ctx->out_a, 6 * (size_t)NI); ENSURE(uint8_t, ctx->out_b, NI);
  const int off[2] = {0, n};
  CU(cudaMemcpyAsync(ctx->in_cam.p, cam, sizeof(double) * 6, cudaMemcpyHostToDevice, s));
  CU(cudaMemcpyAsync(ctx->in_ocam.p, off, sizeof(off), cudaMemcpyHostToDevice, s));
  CU(cudaMemcpyAsync(ctx->in_pt.p, X, sizeof(double) * 3 * n, cudaMemcpyHostToDevice, s));
  CU(cudaMemcpyAsync(ctx->in_u.p, uv, sizeof(double) * 2 * n, cudaMemcpyHostToDevice, s));
  CU(cudaMemsetAsync(ctx->out_a.p, 0, sizeof(double) * 6 * NI, s));
  CU(cudaMemsetAsync(ctx->out_b.p, 0, NI, s));
  double* t = ctx->out_a.as<double>();
  PoseTrace tr{t, t + NI, t + 2 * NI, t + 3 * NI, t + 4 * NI, t + 5 * NI, ctx->out_b.as<uint8_t>(), NI};
  cudaEventRecord(e0, s);
  LAUNCH(k_pose_only, 1, NT_POSE, 1, ctx->in_cam.as<double>(), (const int*)ctx->in_ocam.as<int>(), (const double*)ctx->in_pt.as<double>(),
         (const double*)ctx->in_u.as<double>(), Intr{fx, fy, cx, cy}, pose_opts(opt), ctx->in_cfix.as<uint8_t>(), ctx->in_opt.as<int>(),
         ctx->in_v.as<double>(), ctx->in_opt.as<int>() + 1, tr);
  cudaEventRecord(e1, s);
  CU(cudaGetLastError());
  double out_cam[6]; uint8_t usable = 0; int meta[3] = {0, 0, 0}; double fcost = 0;
  CU(cudaMemcpyAsync(out_cam, ctx->in_cam.p, sizeof(out_cam), cudaMemcpyDeviceToHost, s));
  CU(cudaMemcpyAsync(&usable, ctx->in_cfix.p, 1, cudaMemcpyDeviceToHost, s));
  CU(cudaMemcpyAsync(meta, ctx->in_opt.p, sizeof(meta), cudaMemcpyDeviceToHost, s));
  CU(cudaMemcpyAsync(&fcost, ctx->in_v.p, sizeof(double), cudaMemcpyDeviceToHost, s));
  std::vector<double> trace(6 * (size_t)NI); std::vector<uint8_t> acc(NI);
  if (summary) {
    CU(cudaMemcpyAsync(trace.data(), ctx->out_a.p, sizeof(double) * 6 * NI, cudaMemcpyDeviceToHost, s));
    CU(cudaMemcpyAsync(acc.data(), ctx->out_b.p, NI, cudaMemcpyDeviceToHost, s));
  }
  CU(cudaStreamSynchronize(s));
  float ms = 0.f; cudaEventElapsedTime(&ms, e0, e1);
  if (summary) {
    summary->n_iters = meta[0]; summary->termination = meta[1]; summary->stop_reason = meta[2];
    std::memcpy(summary->cost, &trace[0], sizeof(double) * NI); std::memcpy(summary->cost_candidate, &trace[NI], sizeof(double) * NI);
    std::memcpy(summary->radius, &trace[2 * (size_t)NI], sizeof(double) * NI); std::memcpy(summary->step_norm, &trace[3 * (size_t)NI], sizeof(double) * NI);
    std::memcpy(summary->relative_decrease, &trace[4 * (size_t)NI], sizeof(double) * NI); std::memcpy(summary->gradient_max_norm, &trace[5 * (size_t)NI], sizeof(double) * NI);
    std::memcpy(summary->accepted, acc.data(), NI);
    for (int i = 1; i <= meta[0] && i < NI; ++i) summary->n_successful += acc[i];
    summary->n_linearizations = 1 + summary->n_successful;
    summary->initial_cost = trace[0]; summary->final_cost = fcost; summary->t_total_ms = ms; summary->t_solve_ms = ms;
    summary->status = (meta[2] == GLBA_STOP_NUMERIC) ? GLBA_E_NUMERIC : GLBA_OK;
  }
  if (meta[2] == GLBA_STOP_NUMERIC) return fail(ctx, GLBA_E_NUMERIC, "non-finite residual at the initial pose");
  if (usable) std::memcpy(cam, out_cam, sizeof(out_cam));
  return GLBA_OK;
}

int glba_cull_points(glba_ctx* ctx, const glba_problem* prob, int32_t min_obs, double max_mean_err, uint8_t* bad, double* mean_err) {
  if (!ctx || !bad) return GLBA_E_INVALID_ARG;
  CU(cudaSetDevice(ctx->device));
  ctx->mode = GLBA_MODE_CERES;
  ctx->want_owner = false;
  ctx->allow_relabel = false;            // one pass over the tracks: renumbering would cost more than it saves
  int st = load_problem(ctx, prob);
  ctx->allow_relabel = true;
  if (st) return st;
  glba_options o; glba_default_options(&o);
  const int n_pt = ctx->n_pt;
  if (n_pt == 0) return GLBA_OK;
  ENSURE(uint8_t, ctx->out_b, n_pt); ENSURE(double, ctx->out_c, n_pt);
  LAUNCH(k_cull, cdiv(n_pt, NT_PM), NT_PM, pm_args(ctx, &o), (const double4*)ctx->pt4[0].as<double4>(), (const double*)ctx->camtab[0].as<double>(),
         min_obs, max_mean_err, ctx->out_b.as<uint8_t>(), ctx->out_c.as<double>());
  CU(cudaMemcpyAsync(bad, ctx->out_b.p, n_pt, cudaMemcpyDeviceToHost, ctx->stream));
  if (mean_err) CU(cudaMemcpyAsync(mean_err, ctx->out_c.p, sizeof(double) * n_pt, cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  ctx->ev_used = 0;
  return GLBA_OK;
}

int glba_triangulate_filter(glba_ctx* ctx, const double* R1, const double* t1, const double* R2, const double* t2, double fx, double fy,
                            double cx, double cy, int32_t n, const double* p0, const double* p1, double distance_threshold,
                            double reprojection_threshold, double* X, uint8_t* keep) {
  if (!ctx || !R1 || !t1 || !R2 || !t2 || n < 0 || (n > 0 && (!p0 || !p1 || !X || !keep))) return fail(ctx, GLBA_E_INVALID_ARG, "triangulate: bad argument");
  if (n == 0) return GLBA_OK;
  CU(cudaSetDevice(ctx->device));
  TriArgs T;
  const double Kmat[9] = {fx, 0, cx, 0, fy, cy, 0, 0, 1};
  for (int r = 0; r < 3; ++r) {
    for (int c = 0; c < 3; ++c) { T.T0[4 * r + c] = R1[3 * r + c]; T.T1[4 * r + c] = R2[3 * r + c]; }
    T.T0[4 * r + 3] = t1[r]; T.T1[4 * r + 3] = t2[r];
  }
  for (int r = 0; r < 3; ++r) for (int c = 0; c < 4; ++c) {      // P = K [R|t]  (slam_core.cpp:182-183)
    double a = 0, b = 0;
    for (int k = 0; k < 3; ++k) { a += Kmat[3 * r + k] * T.T0[4 * k + c]; b += Kmat[3 * r + k] * T.T1[4 * k + c]; }
    T.P0[4 * r + c] = a; T.P1[4 * r + c] = b;
  }
  T.K = Intr{fx, fy, cx, cy}; T.dist_thr = distance_threshold; T.reproj_thr = reprojection_threshold;
  cudaStream_t s = ctx->stream;
  ENSURE(double, ctx->in_u, 2 * (size_t)n); ENSURE(double, ctx->in_v, 2 * (size_t)n); ENSURE(double, ctx->out_a, 3 * (size_t)n); ENSURE(uint8_t, ctx->out_b, n);
  CU(cudaMemcpyAsync(ctx->in_u.p, p0, sizeof(double) * 2 * n, cudaMemcpyHostToDevice, s));
  CU(cudaMemcpyAsync(ctx->in_v.p, p1, sizeof(double) * 2 * n, cudaMemcpyHostToDevice, s));
  LAUNCH(k_triangulate, cdiv(n, 128), 128, n, T, (const double*)ctx->in_u.as<double>(), (const double*)ctx->in_v.as<double>(), ctx->out_a.as<double>(),
         ctx->out_b.as<uint8_t>());
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(X, ctx->out_a.p, sizeof(double) * 3 * n, cudaMemcpyDeviceToHost, s));
  CU(cudaMemcpyAsync(keep, ctx->out_b.p, n, cudaMemcpyDeviceToHost, s));
  CU(cudaStreamSynchronize(s));
  return GLBA_OK;
}

}  // extern "C"

#include "glba_map.cuh"
