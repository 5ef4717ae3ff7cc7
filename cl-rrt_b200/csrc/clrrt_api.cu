// clrrt_api.cu — the C ABI of include/clrrt.h: context, device buffers, kernel launches (sm_100a).
// One translation unit (the kernels live in the .cuh files) so that the __constant__ parameter block is
// shared by all kernels.  No CPU fallback: every entry point needs the CUDA device of its context.
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "common.cuh"
#include "nearest.cuh"
#include "rollout.cuh"
#include "simulate.cuh"
#include "tree.cuh"
#include "exchange.cuh"
#include "sequential.cuh"

namespace {

struct RolloutScratch {
  int cap = 0;
  int32_t* d_parent = nullptr;
  uint8_t* d_gb = nullptr;
  double* d_samples = nullptr;
  clrrt_rollout* d_out = nullptr;
  double* d_traj = nullptr;
  size_t traj_cap = 0;
  double* d_ref = nullptr;
  size_t ref_cap = 0;
};

}  // namespace

struct clrrt_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  clrrt_params prm;
  DevParams dprm;
  int cap = 0, max_round = 0, n_tree = 0;
  NodeSoA tree{}, stage{};
  void *tree_mem = nullptr, *stage_mem = nullptr;
  ObsHot* d_hot = nullptr;
  ObsBound* d_bnd = nullptr;
  int32_t* d_cell_start = nullptr;   // broad-phase grid (CSR): [cells + 1]
  uint16_t* d_cell_items = nullptr;
  size_t cell_start_cap = 0, cell_items_cap = 0;
  uint4* d_pose_cells = nullptr;     // pose grid, built on the device by build_pose_grid_kernel
  size_t pose_cap = 0;
  double grid_cell = 1.0;            // requested cell size in metres (clrrt_set_grid_cell)
  bool pose_enabled = true;          // clrrt_set_grid_cell with a negative size disables the pose grid (tests)
  int nn_mode = 0;                   // candidate search: 0 = choose by size, 1 = always the sorted search, 2 = never
  int pose_sub_max = 2;              // pose cells per position cell and axis (measured on C3: 1 -> 5.84, 2 -> 4.98, 4 -> 4.88 ms per round at 4x the table)
  ObsCold* d_cold = nullptr;
  ObsMoving* d_mov = nullptr;
  int obs_cap = 0;
  unsigned char* h_obs_stage = nullptr;  // pinned staging of clrrt_set_obstacles
  size_t obs_stage_cap = 0;
  cudaEvent_t obs_stage_ev = nullptr;
  // round scratch
  double* d_samples = nullptr;
  uint8_t* d_heur = nullptr;
  int32_t *d_cand = nullptr, *d_count = nullptr, *d_valid = nullptr;
  int32_t *d_order = nullptr, *d_hist = nullptr;  // launch order of the (sample, rank) pairs; (rank, length bucket) histogram
  uint32_t* d_done = nullptr;
  // candidate search: nodes and samples sorted along the goal bearing (nearest.cuh)
  void* nn_mem = nullptr;
  NNSortArgs nn{};
  float *d_tile_ulo = nullptr, *d_tile_uhi = nullptr, *d_tile_vlo = nullptr, *d_tile_vhi = nullptr, *d_tile_ce = nullptr, *d_tile_proj = nullptr, *d_tile_feas = nullptr;
  NodeRecord* d_export = nullptr;    // staging of clrrt_tree_download_range
  size_t export_cap = 0;
  cudaStream_t copy_stream = nullptr;   // clrrt_tree_download_range_async: device-to-host copies beside the next round
  cudaEvent_t ev_export = nullptr, ev_copied = nullptr;
  bool copy_pending = false;
  void* d_init = nullptr;            // prepared rollouts of a round (setup_kernel) + per-thread scratch records
  size_t init_stride = 0;
  uint8_t* d_bucket = nullptr;
  int32_t* d_slot = nullptr;
  uint8_t* d_res_code = nullptr;
  uint16_t* d_res_steps = nullptr;
  double* d_ref_end = nullptr;
  float* d_key = nullptr;
  int32_t* d_ints = nullptr;  // [0] head main, [1] head gb, [2] gb_count, [3] total records, [4] best id
  int32_t* d_block_sums = nullptr;
  NodeRecord* d_records = nullptr;
  unsigned long long* d_counters = nullptr;  // 8
  // K = 1: the reference's order of equal keys (nearest_reference_ties)
  float* d_all_key = nullptr; uint8_t* d_all_feas = nullptr;
  int nn_lateral_log2 = -1;  // >= 0: fixed number of lateral bins of the spatial order (tests); -1: by tree size
  int tie_mode = 1;          // 1: std::sort's order of equal keys for single-sample searches; 0: lower node id everywhere
  long long tie_sorts = 0;   // searches that had to repeat the reference's sort on the host
  unsigned long long* d_timeline = nullptr;  // CLRRT_PHASE_CLOCKS builds: 3 words per staging slot (rollout.cuh)
  int32_t* h_ints = nullptr;                 // pinned
  unsigned long long* h_counters = nullptr;  // pinned
  RolloutScratch batch;
  // multi-GPU exchange (exchange.cuh): one context per rank, tree replicated, samples sharded
  ncclComm_t comm = nullptr;
  bool own_comm = false;
  int rank = 0, world = 1;
  int32_t *d_counts = nullptr, *h_counts = nullptr;  // per-rank record counts of the round (device / pinned host)
  NodeRecord* d_gather = nullptr;                    // [world][2 * max_round] gathered records
  cudaEvent_t ev[7]{};
  int num_sms = 0, blocks_per_sm_main = 1, blocks_per_sm_gb = 1;
  size_t smem_bytes = 0;
  bool nn_sorted_last = false;  // the last search sorted the samples (nn.sample_id holds their spatial order)
  // idle lanes a warp accumulates before fetching work.  With the launch order following the samples' positions a larger
  // batch keeps the lanes of a warp together (C3 kernel: 1: 4.22, 2: 4.01, 4: 3.77, 8: 3.60, 12: 3.61, 16: 3.67, 24: 3.74 ms;
  // with the round-1 order, unrelated neighbours, 4 was best)
  int refill_min = 8;
  int blocks_override = 0;
  bool defer_append = false;
  int last_records = 0;
  bool have_tree = false;
  std::string err;
};

namespace {

// The parameter block c_prm is ONE __constant__ symbol per process and device.  Every entry point that launches kernels
// re-uploads it when another context used it last (ensure_params) and synchronises its stream before returning, so
// contexts may be interleaved freely from one thread; concurrent calls from several threads (even on distinct
// contexts of one device) must be serialised by the caller — the reference is single-threaded too (include/clrrt.h).
const clrrt_ctx* g_const_owner = nullptr;  // which context last uploaded c_prm on this process

#define NCK(call)                                                                                         \
  do {                                                                                                    \
    ncclResult_t r_ = (call);                                                                             \
    if (r_ != ncclSuccess) {                                                                              \
      ctx->err = std::string(#call) + ": " + nc->GetErrorString(r_);                                      \
      return CLRRT_ERR_CUDA;                                                                              \
    }                                                                                                     \
  } while (0)
#define CK(call)                                                                                          \
  do {                                                                                                    \
    cudaError_t e_ = (call);                                                                              \
    if (e_ != cudaSuccess) {                                                                              \
      ctx->err = std::string(#call) + ": " + cudaGetErrorString(e_);                                      \
      return CLRRT_ERR_CUDA;                                                                              \
    }                                                                                                     \
  } while (0)

int alloc_soa(clrrt_ctx* ctx, NodeSoA& s, void** mem, int n) {
  const size_t nd = NODE_SOA_DOUBLE_FIELDS, nf = NODE_SOA_FLOAT_FIELDS, ni = NODE_SOA_INT_FIELDS;
  const size_t n8 = ((size_t)n + 7) & ~(size_t)7;
  const size_t bytes = n8 * (nd * 8 + nf * 4 + ni * 4);
  CK(cudaMalloc(mem, bytes));
  CK(cudaMemsetAsync(*mem, 0, bytes, ctx->stream));
  double* d = reinterpret_cast<double*>(*mem);
  double** df[] = {&s.x, &s.y, &s.th, &s.de, &s.v, &s.a, &s.t, &s.s7, &s.s8, &s.s9, &s.rfx, &s.rfy, &s.rbx, &s.rby, &s.vback, &s.angPar, &s.smx, &s.smy};
  for (size_t i = 0; i < nd; i++) *df[i] = d + i * n8;
  float* f = reinterpret_cast<float*>(d + nd * n8);
  float** ff[] = {&s.costE, &s.costS, &s.ca, &s.sa};
  for (size_t i = 0; i < nf; i++) *ff[i] = f + i * n8;
  int32_t* q = reinterpret_cast<int32_t*>(f + nf * n8);
  int32_t** qf[] = {&s.parent, &s.goal, &s.nref, &s.kind};
  for (size_t i = 0; i < ni; i++) *qf[i] = q + i * n8;
  return CLRRT_OK;
}

size_t obstacle_table_bytes(int n_static) {
  // vertices / axes / cell lists stay in global memory (read through L1)
  return ((size_t)n_static + 1) * sizeof(ObsBound);  // + the sentinel record that pads the cell lists
}

// host-side mirror of std::max semantics used by the reference's lookahead formulas
inline double hmax(double a, double b) { return (a < b) ? b : a; }

void fill_dev_params(clrrt_ctx* ctx) {
  const clrrt_params& p = ctx->prm;
  DevParams& d = ctx->dprm;
  const DevParams keep = d;  // obstacle-derived fields survive a parameter update
  memset(&d, 0, sizeof d);
  d.n_static = keep.n_static; d.n_moving = keep.n_moving; d.static_in_smem = keep.static_in_smem;
  d.grid_nx = keep.grid_nx; d.grid_ny = keep.grid_ny; d.grid_inv_cell = keep.grid_inv_cell;
  d.grid_ox = keep.grid_ox; d.grid_oy = keep.grid_oy; d.pose_sub = keep.pose_sub; d.pose_nh = keep.pose_nh;
  d.fine_margin = keep.fine_margin; d.deep_margin = keep.deep_margin;  // computed by clrrt_set_obstacles from the scene
  d.dmax = p.veh.dmax; d.ddmax = p.veh.ddmax; d.inv_Td = 1 / p.veh.Td; d.inv_Ta = 1 / p.veh.Ta;
  d.amin = p.veh.amin; d.amax = p.veh.amax; d.L = p.veh.L; d.Vch = p.veh.Vch; d.Kus = p.veh.Kus;
  d.sim_dt = p.sim_dt; d.mindla = p.ctrl_mindla; d.tla = p.ctrl_tla;
  d.dla_c = p.ctrl_mindla - p.ctrl_tla * p.ctrl_dlavmin;  // controller.cpp:14
  d.Kp = p.ctrl_Kp; d.Ki = p.ctrl_Ki; d.ref_res = p.ref_res; d.vmax = p.vmax; d.ay_road_max = p.ay_road_max;
  for (int i = 0; i < 5; i++) d.W[i] = p.Wcost[i];
  for (int i = 0; i < 4; i++) d.goal[i] = p.goal[i];
  d.feas_len = 2.1 * p.ref_res;  // rrtplanner.cpp:283
  // feasibleGoalBias, rrtplanner.cpp:294-299 (both centre coordinates use cos, as upstream)
  const double R1 = 4.77, R2 = R1 - 0.3;
  d.gb_clx = p.goal[0] + R1 * cos(p.goal[2] - M_PI_2);
  d.gb_cly = p.goal[1] + R1 * cos(p.goal[2] - M_PI_2);
  d.gb_crx = p.goal[0] + R1 * cos(p.goal[2] + M_PI_2);
  d.gb_cry = p.goal[1] + R1 * cos(p.goal[2] + M_PI_2);
  d.gb_R2 = R2;
  // getGoalReference, reference.cpp:27-50
  const double dla_end = hmax(p.ctrl_mindla, d.dla_c + p.ctrl_tla * std::fabs(p.goal[3]));
  const double Dextend = dla_end, Dalign = 1;
  d.gb_P1x = p.goal[0] + Dalign * cos(p.goal[2]); d.gb_P1y = p.goal[1] + Dalign * sin(p.goal[2]);
  d.gb_P2x = p.goal[0] - Dalign * cos(p.goal[2]); d.gb_P2y = p.goal[1] - Dalign * sin(p.goal[2]);
  d.gb_ext_x = (Dextend + Dalign) * cos(p.goal[2]);
  d.gb_ext_y = (Dextend + Dalign) * sin(p.goal[2]);
  // OBB vOBB(vPos, 2, 4.848, theta): float w, h (collision.h:24); setVertices uses w/2, h/2
  const float vw = 2, vh = 4.848;
  d.veh_hw = vw / 2; d.veh_hh = vh / 2;
  d.veh_reach = std::sqrt(d.veh_hw * d.veh_hw + d.veh_hh * d.veh_hh);
  d.exact_dist = (p.Wcost[2] != 0.0) ? 1 : 0;  // the distance only enters the cost through W2*exp(-W3*Dobs)
  int ms = 0;
  while (ms < (20 / p.sim_dt) && ms < CLRRT_MAX_STEPS_CAP * 64) ms++;  // simulation.cpp:58
  d.max_steps = ms;
  d.obs_use_pred = p.obs_use_pred;
  d.bend = p.bend ? 1 : 0; d.lane_S = p.lane_shift; d.Cxy1 = p.Cxy[1]; d.Cxy2 = p.Cxy[2];
}

int upload_params(clrrt_ctx* ctx) {
  CK(cudaMemcpyToSymbolAsync(c_prm, &ctx->dprm, sizeof(DevParams), 0, cudaMemcpyHostToDevice, ctx->stream));
  g_const_owner = ctx;
  return CLRRT_OK;
}
int ensure_params(clrrt_ctx* ctx) {
  if (g_const_owner != ctx) return upload_params(ctx);
  return CLRRT_OK;
}

template <typename R> int configure_launch_t(clrrt_ctx* ctx) {
  const int sm = (int)ctx->smem_bytes;
  CK(cudaFuncSetAttribute(rollout_kernel<R, 2, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm));
  CK(cudaFuncSetAttribute(rollout_kernel<R, 2, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm));
  CK(cudaFuncSetAttribute(rollout_kernel<R, 2, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm));
  CK(cudaFuncSetAttribute(rollout_kernel<R, 2, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm));
  int b0 = 1, b1 = 1;
  if (ctx->dprm.exact_dist) {
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b0, rollout_kernel<R, 2, true, true>, ROLLOUT_THREADS, ctx->smem_bytes));
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b1, rollout_kernel<R, 2, true, false>, ROLLOUT_THREADS, ctx->smem_bytes));
  } else {
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b0, rollout_kernel<R, 2, false, true>, ROLLOUT_THREADS, ctx->smem_bytes));
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b1, rollout_kernel<R, 2, false, false>, ROLLOUT_THREADS, ctx->smem_bytes));
  }
  // the scratch records of goal-biased continuations (d_init, one per thread of the persistent grid) are sized for
  // ROLLOUT_MAX_THREADS_PER_SM resident threads per SM
  const int cap = std::max(1, ROLLOUT_MAX_THREADS_PER_SM / ROLLOUT_THREADS);
  ctx->blocks_per_sm_main = std::min(cap, std::max(1, b0));
  ctx->blocks_per_sm_gb = std::min(cap, std::max(1, b1));
  return CLRRT_OK;
}

int configure_launch(clrrt_ctx* ctx) {
  ctx->smem_bytes = ROLLOUT_SMEM_GB_BYTES + ROLLOUT_SMEM_VB_BYTES + (ctx->dprm.static_in_smem ? obstacle_table_bytes(ctx->dprm.n_static) : 0);
  return ctx->prm.fp32 ? configure_launch_t<float>(ctx) : configure_launch_t<double>(ctx);
}

// One kernel per mode (round / batch): plain and goal-biased rollouts are told apart per lane at run time (GBM = 2).
template <typename R, bool ROUND> int launch_rollout_t(clrrt_ctx* ctx, const RolloutJob& job, int blocks) {
  if (ctx->dprm.exact_dist)
    rollout_kernel<R, 2, true, ROUND><<<blocks, ROLLOUT_THREADS, ctx->smem_bytes, ctx->stream>>>(job, ctx->d_bnd, ctx->d_hot, ctx->d_cold, ctx->d_mov, ctx->d_cell_start, ctx->d_cell_items, ctx->d_pose_cells);
  else
    rollout_kernel<R, 2, false, ROUND><<<blocks, ROLLOUT_THREADS, ctx->smem_bytes, ctx->stream>>>(job, ctx->d_bnd, ctx->d_hot, ctx->d_cold, ctx->d_mov, ctx->d_cell_start, ctx->d_cell_items, ctx->d_pose_cells);
  CK(cudaGetLastError());
  return CLRRT_OK;
}

int launch_rollout(clrrt_ctx* ctx, const RolloutJob& job_in, int n_items_hint) {
  RolloutJob job = job_in;
  const bool round = job.sample_word != nullptr;
  int per_sm = round ? ctx->blocks_per_sm_main : ctx->blocks_per_sm_gb;
  if (ctx->blocks_override > 0) per_sm = std::min(per_sm, ctx->blocks_override);
  int blocks = ctx->num_sms * per_sm;  // persistent grid: a multiple of the SM count
  // Small launches (a sequential window, a batch of a few thousand rollouts) are latency-bound: a rollout alone in its warp
  // takes 1.7 us per step, one of 32 diverging lanes several times that.  So the items are spread over ALL the warps of the
  // persistent grid, `take_cap` lanes per warp at a time, instead of filling the first warps to the brim.
  const int warps_per_block = ROLLOUT_THREADS / 32, warps = blocks * warps_per_block;
  job.take_cap = std::max(1, std::min(32, (n_items_hint + warps - 1) / warps));
  const int lanes_per_block = job.take_cap * warps_per_block;
  const int needed = (n_items_hint + lanes_per_block - 1) / lanes_per_block;
  if (needed < blocks) blocks = std::max(1, needed);
  if (round) return ctx->prm.fp32 ? launch_rollout_t<float, true>(ctx, job, blocks) : launch_rollout_t<double, true>(ctx, job, blocks);
  return ctx->prm.fp32 ? launch_rollout_t<float, false>(ctx, job, blocks) : launch_rollout_t<double, false>(ctx, job, blocks);
}

}  // namespace

extern "C" {

int clrrt_default_params(clrrt_params* p) {
  if (!p) return CLRRT_ERR_ARG;
  memset(p, 0, sizeof *p);
  // Vehicle::setPrius(), rrt/include/rrt/vehicle.h:39-60
  clrrt_vehicle& v = p->veh;
  v.dmax = 0.52; v.ddmax = 0.3294; v.Td = 0.3; v.Ta = 0.3; v.amin = -6; v.amax = 2; v.L = 2.7;
  const double lf = 1.0868, lr = 1.6132;
  v.Lrear = 1; v.Lfront = 2.7 + 0.5; v.w = 2; v.b = lr; v.rho = 5.95;
  const double Cf = 22201, Cr = 22201, m = 950 + 640;
  v.Kus = (m / v.L) * (lr / Cf - lf / Cr);
  v.Vch = 20;
  // rrt/launch/parameters.launch:3-20
  p->ctrl_tla = 1.4; p->ctrl_mindla = 3.2; p->ctrl_dlavmin = 3; p->ref_int = 0.02; p->ref_mindist = 0.2;
  p->sim_dt = 0.04; p->ctrl_Kp = 8; p->ctrl_Ki = 0.05;
  p->Wcost[0] = 10; p->Wcost[1] = 5; p->Wcost[2] = 0; p->Wcost[3] = 4; p->Wcost[4] = 1;
  p->vmax = 5; p->ay_road_max = 0;
  p->ref_res = hmax(std::fabs(0.0) * p->ref_int, p->ref_mindist);
  p->goal[0] = 50; p->goal[1] = 0; p->goal[2] = 0; p->goal[3] = 0;
  p->obs_use_pred = 1; p->fp32 = 0;
  return CLRRT_OK;
}

int clrrt_create(const clrrt_params* p, int device, int tree_capacity, int max_round, void* stream, clrrt_ctx** out) {
  if (!p || !out || tree_capacity < 1 || max_round < 1) return CLRRT_ERR_ARG;
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0 || device < 0 || device >= ndev) return CLRRT_ERR_CUDA;
  clrrt_ctx* ctx = new clrrt_ctx();
  ctx->device = device;
  ctx->prm = *p;
  ctx->cap = tree_capacity;
  ctx->max_round = max_round;
  memset(&ctx->dprm, 0, sizeof ctx->dprm);
  auto fail = [&](int code) { *out = ctx; return code; };  // caller can read clrrt_last_error, then destroy
  if (cudaSetDevice(device) != cudaSuccess) { ctx->err = "cudaSetDevice failed"; return fail(CLRRT_ERR_CUDA); }
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { ctx->err = "cudaGetDeviceProperties failed"; return fail(CLRRT_ERR_CUDA); }
  ctx->num_sms = prop.multiProcessorCount;
  if (stream) ctx->stream = (cudaStream_t)stream;
  else {
    if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) { ctx->err = "cudaStreamCreate failed"; return fail(CLRRT_ERR_CUDA); }
    ctx->own_stream = true;
  }
  int rc;
  if ((rc = alloc_soa(ctx, ctx->tree, &ctx->tree_mem, tree_capacity)) != CLRRT_OK) return fail(rc);
  if ((rc = alloc_soa(ctx, ctx->stage, &ctx->stage_mem, (CLRRT_SORT_LIMIT + 1) * max_round)) != CLRRT_OK) return fail(rc);
  auto mal = [&](void** ptr, size_t bytes) { return cudaMalloc(ptr, bytes) == cudaSuccess; };
  const size_t K = (size_t)max_round;
  bool ok = true;
  ok &= mal((void**)&ctx->d_samples, K * 2 * sizeof(double));
  ok &= mal((void**)&ctx->d_heur, K);
  ok &= mal((void**)&ctx->d_cand, K * CLRRT_SORT_LIMIT * sizeof(int32_t));
  ok &= mal((void**)&ctx->d_key, K * CLRRT_SORT_LIMIT * sizeof(float));
  ok &= mal((void**)&ctx->d_count, K * sizeof(int32_t));
  ok &= mal((void**)&ctx->d_valid, 2 * K * sizeof(int32_t));
  ok &= mal((void**)&ctx->d_order, K * CLRRT_SORT_LIMIT * sizeof(int32_t));
  ok &= mal((void**)&ctx->d_hist, 1024 * sizeof(int32_t));
  ok &= mal((void**)&ctx->d_done, K * sizeof(uint32_t));
  ok &= mal((void**)&ctx->d_bucket, K * CLRRT_SORT_LIMIT);
  {
    const size_t n8 = ((size_t)tree_capacity + 7) & ~(size_t)7, k8 = (K + 7) & ~(size_t)7;
    const size_t bytes = n8 * (7 * 8 + 9 * 4 + 4 + 4) + (n8 + k8) * 4 + k8 * 4 + (NN_HIST_INTS + 4) * 4 + (8 + NN_DIRS + NN_FCLS) * (n8 / NEAREST_TILE + 8) * 4 + 2048;
    ok &= mal(&ctx->nn_mem, bytes);
    if (ok) {
      unsigned char* p = reinterpret_cast<unsigned char*>(ctx->nn_mem);
      auto take = [&](size_t nbytes) { unsigned char* q = p; p += (nbytes + 15) & ~(size_t)15; return q; };
      NNSortArgs& s = ctx->nn;
      double** dd[] = {&s.nx, &s.ny, &s.rbx, &s.rby, &s.dpx, &s.dpy, &s.ang};
      for (auto d : dd) *d = reinterpret_cast<double*>(take(n8 * 8));
      float** ff[] = {&s.ca, &s.sa, &s.ce, &s.fx, &s.fy, &s.frx, &s.fry, &s.fdx, &s.fdy};
      for (auto f : ff) *f = reinterpret_cast<float*>(take(n8 * 4));
      s.node_id = reinterpret_cast<int32_t*>(take(n8 * 4));
      s.sbin = reinterpret_cast<int32_t*>(take(n8 * 4));
      s.bin = reinterpret_cast<int32_t*>(take((n8 + k8) * 4));
      s.sample_id = reinterpret_cast<int32_t*>(take(k8 * 4));
      s.hist = reinterpret_cast<int32_t*>(take((NN_HIST_INTS + 4) * 4));
      s.lead_sum = reinterpret_cast<float*>(s.hist + NN_HIST_INTS);
      ctx->d_tile_ce = reinterpret_cast<float*>(take((n8 / NEAREST_TILE + 8) * 4));
      ctx->d_tile_ulo = reinterpret_cast<float*>(take((n8 / NEAREST_TILE + 8) * 4));
      ctx->d_tile_uhi = reinterpret_cast<float*>(take((n8 / NEAREST_TILE + 8) * 4));
      ctx->d_tile_vlo = reinterpret_cast<float*>(take((n8 / NEAREST_TILE + 8) * 4));
      ctx->d_tile_vhi = reinterpret_cast<float*>(take((n8 / NEAREST_TILE + 8) * 4));
      ctx->d_tile_proj = reinterpret_cast<float*>(take((n8 / NEAREST_TILE + 8) * NN_DIRS * 4));
      ctx->d_tile_feas = reinterpret_cast<float*>(take((n8 / NEAREST_TILE + 8) * NN_FCLS * 4));
    }
  }
  ctx->init_stride = ((K * CLRRT_SORT_LIMIT + (size_t)ctx->num_sms * ROLLOUT_MAX_THREADS_PER_SM) + 31) & ~(size_t)31;
  ok &= mal(&ctx->d_init, ctx->init_stride * LANE_INIT_BYTES_PER_RECORD);
  ok &= mal((void**)&ctx->d_slot, K * sizeof(int32_t));
  ok &= mal((void**)&ctx->d_res_code, K * (CLRRT_SORT_LIMIT + 1));  // + the goal-biased rollout of every sample
  ok &= mal((void**)&ctx->d_res_steps, K * (CLRRT_SORT_LIMIT + 1) * sizeof(uint16_t));
  ok &= mal((void**)&ctx->d_ref_end, K * CLRRT_SORT_LIMIT * 2 * sizeof(double));
  ok &= mal((void**)&ctx->d_ints, 128 * sizeof(int32_t));
  ok &= mal((void**)&ctx->d_block_sums, ((K + SCAN_THREADS - 1) / SCAN_THREADS + 1) * sizeof(int32_t));
  ok &= mal((void**)&ctx->d_records, 2 * K * sizeof(NodeRecord));
  ok &= mal((void**)&ctx->d_counters, 32 * sizeof(unsigned long long));
  ok &= mal((void**)&ctx->d_all_key, (size_t)tree_capacity * sizeof(float));
  ok &= mal((void**)&ctx->d_all_feas, (size_t)tree_capacity);
#ifdef CLRRT_PHASE_CLOCKS
  ok &= mal((void**)&ctx->d_timeline, (K * CLRRT_SORT_LIMIT + K) * 3 * sizeof(unsigned long long));
#endif
  ok &= cudaMallocHost((void**)&ctx->h_ints, 16 * sizeof(int32_t)) == cudaSuccess;
  ok &= cudaMallocHost((void**)&ctx->h_counters, 16 * sizeof(unsigned long long)) == cudaSuccess;
  if (!ok) { ctx->err = std::string("device allocation failed: ") + cudaGetErrorString(cudaGetLastError()); return fail(CLRRT_ERR_CUDA); }
  cudaMemsetAsync(ctx->d_counters, 0, 32 * sizeof(unsigned long long), ctx->stream);
  cudaMemsetAsync(ctx->d_ints, 0, 128 * sizeof(int32_t), ctx->stream);
  for (auto& e : ctx->ev) cudaEventCreate(&e);
  fill_dev_params(ctx);
  if ((rc = configure_launch(ctx)) != CLRRT_OK) return fail(rc);
  if ((rc = upload_params(ctx)) != CLRRT_OK) return fail(rc);
  if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) { ctx->err = "initial sync failed"; return fail(CLRRT_ERR_CUDA); }
  *out = ctx;
  return CLRRT_OK;
}

int clrrt_destroy(clrrt_ctx* ctx) {
  if (!ctx) return CLRRT_ERR_ARG;
  cudaSetDevice(ctx->device);
  if (ctx->stream) cudaStreamSynchronize(ctx->stream);
  void* ptrs[] = {ctx->tree_mem, ctx->stage_mem, ctx->d_cell_start, ctx->d_cell_items, ctx->d_pose_cells, ctx->d_bnd, ctx->d_hot, ctx->d_res_code, ctx->d_res_steps, ctx->d_slot, ctx->d_ref_end, ctx->d_cold, ctx->d_mov, ctx->d_samples, ctx->d_heur,
                  ctx->d_cand, ctx->d_key, ctx->d_count, ctx->d_valid, ctx->d_order, ctx->d_hist, ctx->d_done, ctx->d_bucket, ctx->d_init, ctx->nn_mem, ctx->d_export, ctx->d_ints, ctx->d_block_sums,
                  ctx->d_records, ctx->d_counters, ctx->d_timeline, ctx->d_all_key, ctx->d_all_feas, ctx->batch.d_parent, ctx->batch.d_gb,
                  ctx->batch.d_samples, ctx->batch.d_out, ctx->batch.d_traj, ctx->batch.d_ref};
  for (void* p : ptrs) if (p) cudaFree(p);
  if (ctx->comm && ctx->own_comm) { const NcclApi* nc = nccl_api(nullptr); if (nc) nc->CommDestroy(ctx->comm); }
  if (ctx->d_counts) cudaFree(ctx->d_counts);
  if (ctx->d_gather) cudaFree(ctx->d_gather);
  if (ctx->h_counts) cudaFreeHost(ctx->h_counts);
  if (ctx->copy_stream) { cudaStreamSynchronize(ctx->copy_stream); cudaStreamDestroy(ctx->copy_stream); }
  if (ctx->ev_export) cudaEventDestroy(ctx->ev_export);
  if (ctx->ev_copied) cudaEventDestroy(ctx->ev_copied);
  if (ctx->h_obs_stage) cudaFreeHost(ctx->h_obs_stage);
  if (ctx->obs_stage_ev) cudaEventDestroy(ctx->obs_stage_ev);
  if (ctx->h_ints) cudaFreeHost(ctx->h_ints);
  if (ctx->h_counters) cudaFreeHost(ctx->h_counters);
  for (auto& e : ctx->ev) if (e) cudaEventDestroy(e);
  if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
  if (g_const_owner == ctx) g_const_owner = nullptr;
  delete ctx;
  return CLRRT_OK;
}

const char* clrrt_last_error(const clrrt_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }
int clrrt_get_device(const clrrt_ctx* ctx) { return ctx ? ctx->device : CLRRT_ERR_ARG; }

int clrrt_set_params(clrrt_ctx* ctx, const clrrt_params* p) {
  if (!ctx || !p) return CLRRT_ERR_ARG;
  CK(cudaSetDevice(ctx->device));
  const bool mode_changed = (ctx->prm.fp32 != 0) != (p->fp32 != 0) || (ctx->prm.Wcost[2] != 0.0) != (p->Wcost[2] != 0.0);
  ctx->prm = *p;
  fill_dev_params(ctx);
  if (mode_changed) {
    int rc = configure_launch(ctx);
    if (rc != CLRRT_OK) return rc;
  }
  return upload_params(ctx);
}

int clrrt_set_obstacles(clrrt_ctx* ctx, const clrrt_obstacle* host, int n) {
  if (!ctx || n < 0 || (n > 0 && !host)) return CLRRT_ERR_ARG;
  CK(cudaSetDevice(ctx->device));
  struct StaticObs { ObsHot hot; ObsCold cold; float cx, cy, reach, ohh, ohw, oc, os; double cxd, cyd; uint32_t key; };
  std::vector<StaticObs> st;
  std::vector<ObsMoving> mov;
  const float margin = 0.1f;  // rollout.cuh: circles farther apart than this cannot collide in the reference's SAT
  const float vreach = ctx->dprm.veh_reach;
  for (int i = 0; i < n; i++) {
    const clrrt_obstacle& o = host[i];
    // getOBBvector, old_collisioncheck.cpp:14-16: OBB(centre, size_x/2, size_y/2, theta) with float w, h, o
    const float w = (float)(o.size_x / 2), h = (float)(o.size_y / 2), th = (float)o.theta;
    const float co = cosf(th), so = sinf(th);  // the host libm: bit-identical to the reference's own calls
    const float ch = co * (h / 2), sw = so * (w / 2), sh = so * (h / 2), cw = co * (w / 2);
    const float reach = std::sqrt((h / 2) * (h / 2) + (w / 2) * (w / 2));
    if (o.vx == 0.0 && o.vy == 0.0) {
      StaticObs s;
      ObsHot& a = s.hot;
      ObsCold& c = s.cold;
      const double px = o.cx, py = o.cy;  // centre + 0*t
      a.vx[0] = (float)((px + (double)ch) - (double)sw); a.vy[0] = (float)((py + (double)sh) + (double)cw);
      a.vx[1] = (float)((px + (double)ch) + (double)sw); a.vy[1] = (float)((py + (double)sh) - (double)cw);
      a.vx[2] = (float)((px - (double)ch) + (double)sw); a.vy[2] = (float)((py - (double)sh) - (double)cw);
      a.vx[3] = (float)((px - (double)ch) - (double)sw); a.vy[3] = (float)((py - (double)sh) + (double)cw);
      for (int k = 0; k < 3; k++) { c.nx[k] = a.vy[k + 1] - a.vy[k]; c.ny[k] = -(a.vx[k + 1] - a.vx[k]); }
      c.nx[3] = -(a.vx[0] - a.vx[3]); c.ny[3] = 0.0f;
      for (int k = 0; k < 4; k++) {
        float mx = 0, mn = 0;
        for (int q = 0; q < 4; q++) {
          const float m1 = a.vx[q] * c.nx[k];
          const float m2 = a.vy[q] * c.ny[k];
          const float pr = m1 + m2;
          if (q == 0) { mx = pr; mn = pr; }
          else if (pr > mx) mx = pr;
          else if (pr < mn) mn = pr;
        }
        c.pmax[k] = mx; c.pmin[k] = mn;
      }
      s.cx = (float)o.cx; s.cy = (float)o.cy; s.cxd = o.cx; s.cyd = o.cy; s.reach = reach; s.key = 0;
      s.ohh = h / 2; s.ohw = w / 2; s.oc = co; s.os = so;
      st.push_back(s);
    } else {
      ObsMoving m;
      m.cx = o.cx; m.cy = o.cy; m.vx = o.vx; m.vy = o.vy; m.ch = ch; m.sw = sw; m.sh = sh; m.cw = cw;
      m.rr = reach + margin + 0.05f;  // + slack for the float rounding of the predicted centre
      m.ohh = h / 2; m.ohw = w / 2; m.oc = co; m.os = so; m.pad0 = m.pad1 = m.pad2 = 0;
      mov.push_back(m);
    }
  }
  // margins of the verdict-only check (rollout.cuh, box_class): proportional to the largest |coordinate| an obstacle of
  // this scene is tested at (moving obstacles: rollout.cuh widens the margin with the lane's own position)
  {
    double maxabs = 0;
    for (int i = 0; i < n; i++)
      maxabs = std::max(maxabs, std::max(std::fabs(host[i].cx), std::fabs(host[i].cy)) + std::fabs(host[i].size_x) + std::fabs(host[i].size_y) + 2.0 * vreach + 1.0);
    ctx->dprm.fine_margin = std::max(FINE_MARGIN_MIN, FINE_MARGIN_REL * (float)(2.0 * maxabs));
    ctx->dprm.deep_margin = DEEP_MARGIN_FACTOR * ctx->dprm.fine_margin;
  }
  if (st.size() > 32000 || mov.size() > 32000) { ctx->err = "more than 32000 static or moving obstacles"; return CLRRT_ERR_CAPACITY; }
  // Z-order sort of the static obstacles so that 32 consecutive ones form a compact group (the verdict does not
  // depend on the order in which obstacles are tested; neither does the minimum distance of the exact mode)
  if (!st.empty()) {
    float x0 = st[0].cx, x1 = st[0].cx, y0 = st[0].cy, y1 = st[0].cy;
    for (auto& s : st) { x0 = std::min(x0, s.cx); x1 = std::max(x1, s.cx); y0 = std::min(y0, s.cy); y1 = std::max(y1, s.cy); }
    const float ext = std::max(std::max(x1 - x0, y1 - y0), 1e-3f);
    auto spread = [](uint32_t v) { v &= 0xffff; v = (v | (v << 8)) & 0x00ff00ff; v = (v | (v << 4)) & 0x0f0f0f0f; v = (v | (v << 2)) & 0x33333333; v = (v | (v << 1)) & 0x55555555; return v; };
    for (auto& s : st) {
      const uint32_t qx = (uint32_t)(65535.0f * (s.cx - x0) / ext), qy = (uint32_t)(65535.0f * (s.cy - y0) / ext);
      s.key = spread(qx) | (spread(qy) << 1);
    }
    std::stable_sort(st.begin(), st.end(), [](const StaticObs& a, const StaticObs& b) { return a.key < b.key; });
  }
  const int ns = (int)st.size();
  std::vector<ObsHot> hot((size_t)std::max(ns, 1));
  std::vector<ObsCold> cold((size_t)std::max(ns, 1));
  std::vector<ObsBound> bnd((size_t)ns + 1);
  // ---- broad-phase grid over vehicle-box-centre positions (rollout.cuh, warp_collide) ---------------------------
  // Obstacle i can touch a vehicle centred at p only if |p - c_i| <= reach_i + vreach; the list of a cell holds
  // every obstacle whose circle of radius R_i = reach_i + vreach + margin (+ 1 cm for the float cell lookup) meets
  // the cell's square.  The grid covers the bounding box of the centres enlarged by max R_i: a vehicle outside it
  // is out of reach of every static obstacle.
  double gox = 0, goy = 0, cell = ctx->grid_cell;
  int gnx = 1, gny = 1;
  std::vector<int32_t> cell_start(2, 0);
  std::vector<uint16_t> cell_items;
  if (ns > 0) {
    double x0 = st[0].cxd, x1 = st[0].cxd, y0 = st[0].cyd, y1 = st[0].cyd, Rmax = 0;
    for (auto& s : st) {
      x0 = std::min(x0, s.cxd); x1 = std::max(x1, s.cxd); y0 = std::min(y0, s.cyd); y1 = std::max(y1, s.cyd);
      Rmax = std::max(Rmax, (double)s.reach + vreach + margin + 0.01);
    }
    gox = x0 - Rmax; goy = y0 - Rmax;
    const double wx = (x1 + Rmax) - gox, wy = (y1 + Rmax) - goy;
    while (std::ceil(wx / cell) * std::ceil(wy / cell) > 262144.0) cell *= 1.25;
    gnx = std::max(1, (int)std::ceil(wx / cell)); gny = std::max(1, (int)std::ceil(wy / cell));
    cell_start.assign((size_t)gnx * gny + 1, 0);
    int32_t* fill_cursor = nullptr;
    auto visit = [&](bool fill) {
      for (int i = 0; i < ns; i++) {
        const double rx = st[i].cxd - gox, ry = st[i].cyd - goy, R = (double)st[i].reach + vreach + margin + 0.01;
        const int ix0 = std::max(0, (int)std::floor((rx - R) / cell)), ix1 = std::min(gnx - 1, (int)std::floor((rx + R) / cell));
        const int iy0 = std::max(0, (int)std::floor((ry - R) / cell)), iy1 = std::min(gny - 1, (int)std::floor((ry + R) / cell));
        for (int iy = iy0; iy <= iy1; iy++)
          for (int ix = ix0; ix <= ix1; ix++) {
            const double dx = std::max(std::max(ix * cell - rx, rx - (ix + 1) * cell), 0.0);
            const double dy = std::max(std::max(iy * cell - ry, ry - (iy + 1) * cell), 0.0);
            if (dx * dx + dy * dy > R * R) continue;
            const size_t c = (size_t)iy * gnx + ix;
            if (fill) cell_items[(size_t)fill_cursor[c]++] = (uint16_t)i;
            else cell_start[c + 1]++;
          }
      }
    };
    visit(false);
    // lists are padded to whole blocks of 8 ids; cell_start is kept in blocks
    std::vector<int32_t> cnt(cell_start.begin() + 1, cell_start.end());
    cell_start[0] = 0;
    for (size_t c = 0; c < cnt.size(); c++) cell_start[c + 1] = cell_start[c] + (cnt[c] + 7) / 8;
    cell_items.assign((size_t)cell_start.back() * 8, (uint16_t)ns);
    std::vector<int32_t> fillpos(cnt.size());
    for (size_t c = 0; c < cnt.size(); c++) fillpos[c] = cell_start[c] * 8;
    fill_cursor = fillpos.data();
    visit(true);
  }
  for (int i = 0; i < ns; i++) {
    hot[i] = st[i].hot; cold[i] = st[i].cold;
    bnd[i].cx = (float)(st[i].cxd - gox); bnd[i].cy = (float)(st[i].cyd - goy);
    bnd[i].rr = st[i].reach + margin; bnd[i].ohh = st[i].ohh; bnd[i].ohw = st[i].ohw;
    bnd[i].oc = st[i].oc; bnd[i].os = st[i].os; bnd[i].pad = 0.0f;
  }
  bnd[ns].cx = 0; bnd[ns].cy = 0; bnd[ns].rr = -1.0e30f; bnd[ns].ohh = 0; bnd[ns].oc = 1; bnd[ns].os = 0; bnd[ns].ohw = 0; bnd[ns].pad = 0;
  const int total = std::max<int>(32, n + 32);
  if (total > ctx->obs_cap) {
    void* old[] = {ctx->d_hot, ctx->d_cold, ctx->d_mov, ctx->d_bnd};
    for (void* q : old) if (q) cudaFree(q);
    ctx->d_hot = nullptr; ctx->d_cold = nullptr; ctx->d_mov = nullptr; ctx->d_bnd = nullptr;
    ctx->obs_cap = 0;  // a failed allocation below must not leave the old capacity with null tables
    CK(cudaMalloc((void**)&ctx->d_hot, total * sizeof(ObsHot)));
    CK(cudaMalloc((void**)&ctx->d_cold, total * sizeof(ObsCold)));
    CK(cudaMalloc((void**)&ctx->d_mov, total * sizeof(ObsMoving)));
    CK(cudaMalloc((void**)&ctx->d_bnd, (size_t)(total + 1) * sizeof(ObsBound)));
    ctx->obs_cap = total;
  }
  if (cell_start.size() > ctx->cell_start_cap) {
    if (ctx->d_cell_start) cudaFree(ctx->d_cell_start);
    ctx->d_cell_start = nullptr; ctx->cell_start_cap = 0;
    CK(cudaMalloc((void**)&ctx->d_cell_start, cell_start.size() * sizeof(int32_t)));
    ctx->cell_start_cap = cell_start.size();
  }
  if (cell_items.size() + 8 > ctx->cell_items_cap) {
    if (ctx->d_cell_items) cudaFree(ctx->d_cell_items);
    ctx->d_cell_items = nullptr; ctx->cell_items_cap = 0;
    CK(cudaMalloc((void**)&ctx->d_cell_items, (cell_items.size() + 8) * sizeof(uint16_t)));
    ctx->cell_items_cap = cell_items.size() + 8;
  }
  // one pinned staging buffer, asynchronous copies on the context's stream: nothing here blocks the host (a query with
  // moving obstacles calls this every time, rrt/src/motionplanner.cpp:18).  The buffer is reused only after the copies of
  // the previous call have completed (event).
  {
    auto al = [](size_t b) { return (b + 255) & ~(size_t)255; };
    const size_t b_hot = hot.size() * sizeof(ObsHot), b_cold = cold.size() * sizeof(ObsCold), b_bnd = bnd.size() * sizeof(ObsBound),
                 b_cs = cell_start.size() * sizeof(int32_t), b_ci = cell_items.size() * sizeof(uint16_t), b_mov = mov.size() * sizeof(ObsMoving);
    const size_t total = al(b_hot) + al(b_cold) + al(b_bnd) + al(b_cs) + al(b_ci) + al(b_mov);
    if (ctx->obs_stage_ev) CK(cudaEventSynchronize(ctx->obs_stage_ev));
    else CK(cudaEventCreateWithFlags(&ctx->obs_stage_ev, cudaEventDisableTiming));
    if (total > ctx->obs_stage_cap) {
      if (ctx->copy_stream) { cudaStreamSynchronize(ctx->copy_stream); cudaStreamDestroy(ctx->copy_stream); }
  if (ctx->ev_export) cudaEventDestroy(ctx->ev_export);
  if (ctx->ev_copied) cudaEventDestroy(ctx->ev_copied);
  if (ctx->h_obs_stage) cudaFreeHost(ctx->h_obs_stage);
      ctx->h_obs_stage = nullptr; ctx->obs_stage_cap = 0;
      CK(cudaMallocHost((void**)&ctx->h_obs_stage, 2 * total));
      ctx->obs_stage_cap = 2 * total;
    }
    unsigned char* q = ctx->h_obs_stage;
    auto put = [&](void* dst, const void* src, size_t b) -> cudaError_t {
      if (b == 0) return cudaSuccess;
      memcpy(q, src, b);
      const cudaError_t e = cudaMemcpyAsync(dst, q, b, cudaMemcpyHostToDevice, ctx->stream);
      q += al(b);
      return e;
    };
    CK(put(ctx->d_hot, hot.data(), b_hot));
    CK(put(ctx->d_cold, cold.data(), b_cold));
    CK(put(ctx->d_bnd, bnd.data(), b_bnd));
    CK(put(ctx->d_cell_start, cell_start.data(), b_cs));
    CK(put(ctx->d_cell_items, cell_items.data(), b_ci));
    CK(put(ctx->d_mov, mov.data(), b_mov));
    CK(cudaEventRecord(ctx->obs_stage_ev, ctx->stream));
  }
  ctx->dprm.n_static = ns;
  ctx->dprm.n_moving = (int)mov.size();
  ctx->dprm.grid_nx = gnx; ctx->dprm.grid_ny = gny; ctx->dprm.grid_inv_cell = (float)(1.0 / cell);
  ctx->dprm.grid_ox = gox; ctx->dprm.grid_oy = goy;
  // ---- pose grid: 2 x 2 sub-cells per position cell, 32 heading bins over pi (16 B per pose cell) ------------------
  ctx->dprm.pose_sub = 1; ctx->dprm.pose_nh = 0;
  if (ns > 0 && ctx->pose_enabled) {
    const int nh = 32;
    int sub = ctx->pose_sub_max;   // finest pose grid within 2 M cells (32 MB)
    while (sub > 1 && (double)gnx * gny * sub * sub * nh > 2.0e6) sub >>= 1;
    const size_t cells = (size_t)gnx * gny * sub * sub * nh;
    if ((double)cells <= 2.0e6) {
      if (cells > ctx->pose_cap) {
        if (ctx->d_pose_cells) cudaFree(ctx->d_pose_cells);
        ctx->d_pose_cells = nullptr; ctx->pose_cap = 0;
        CK(cudaMalloc((void**)&ctx->d_pose_cells, cells * sizeof(uint4)));
        ctx->pose_cap = cells;
      }
      const double cell_f = cell / sub, dth = M_PI / nh;
      // farthest displacement of a point of the vehicle box when the pose moves inside a cell: half diagonal of the
      // position cell + chord of the half heading bin at the box's half diagonal (+ 1 cm / 1 mrad for the float lookup)
      const float infl = (float)(cell_f * M_SQRT1_2 + 0.01 + 2.0 * vreach * std::sin((dth / 2 + 1e-3) / 2));
      build_pose_grid_kernel<<<(unsigned)((cells + 255) / 256), 256, 0, ctx->stream>>>(
          ctx->d_bnd, ctx->d_cell_start, ctx->d_cell_items, ctx->d_pose_cells, ns, gnx, gny, sub, nh, (float)cell_f, infl,
          ctx->dprm.veh_hh, ctx->dprm.veh_hw, ctx->dprm.fine_margin);
      CK(cudaGetLastError());  // (stream-ordered with the rounds that follow: no host synchronisation)
      ctx->dprm.pose_sub = sub; ctx->dprm.pose_nh = nh;
    }
  }
  // the broad-phase table is staged in shared memory when it leaves room for the other resident blocks
  ctx->dprm.static_in_smem = (ns > 0 && obstacle_table_bytes(ns) <= 48 * 1024) ? 1 : 0;
  const size_t want_smem = ROLLOUT_SMEM_GB_BYTES + ROLLOUT_SMEM_VB_BYTES + (ctx->dprm.static_in_smem ? obstacle_table_bytes(ns) : 0);
  if (want_smem != ctx->smem_bytes) {  // function attributes / occupancy only change with the staged table's size
    int rc = configure_launch(ctx);
    if (rc != CLRRT_OK) return rc;
  }
  return upload_params(ctx);
}

int clrrt_tree_reset(clrrt_ctx* ctx, const clrrt_node* host, int n) {
  if (!ctx || !host || n < 1) return CLRRT_ERR_ARG;
  if (n > ctx->cap) return CLRRT_ERR_CAPACITY;
  // parents precede their children (initializeTree builds a chain, rrtplanner.cpp:90-93): -1 <= parent < i, so that
  // clrrt_best_path's walk over the parent array terminates inside the tree
  for (int i = 0; i < n; i++)
    if (host[i].parent < -1 || host[i].parent >= i) { ctx->err = "clrrt_tree_reset: node " + std::to_string(i) + " has parent " + std::to_string(host[i].parent); return CLRRT_ERR_ARG; }
  CK(cudaSetDevice(ctx->device));
  // route through the record path so that derived fields are produced by the same kernel as for appended nodes
  std::vector<NodeRecord> rec((size_t)n);
  for (int i = 0; i < n; i++) {
    NodeRecord& r = rec[i];
    memset(&r, 0, sizeof r);
    memcpy(r.state, host[i].state, sizeof r.state);
    r.rf[0] = host[i].ref_front[0]; r.rf[1] = host[i].ref_front[1];
    r.rb[0] = host[i].ref_back[0]; r.rb[1] = host[i].ref_back[1];
    r.vback = host[i].ref_vback; r.costE = host[i].costE; r.costS = host[i].costS;
    r.parent = host[i].parent; r.goal = host[i].goal_reached; r.nref = host[i].n_ref; r.kind = host[i].kind;
    r.smp[0] = host[i].sample[0]; r.smp[1] = host[i].sample[1];
  }
  NodeRecord* d_tmp = nullptr;
  CK(cudaMalloc((void**)&d_tmp, (size_t)n * sizeof(NodeRecord)));
  CK(cudaMemcpyAsync(d_tmp, rec.data(), (size_t)n * sizeof(NodeRecord), cudaMemcpyHostToDevice, ctx->stream));
  append_records_kernel<<<(n + 255) / 256, 256, 0, ctx->stream>>>(ctx->tree, 0, d_tmp, n, ctx->cap);
  CK(cudaGetLastError());
  CK(cudaMemsetAsync(ctx->d_counters, 0, 8 * sizeof(unsigned long long), ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  CK(cudaFree(d_tmp));
  ctx->n_tree = n;
  ctx->have_tree = true;
  return CLRRT_OK;
}

int clrrt_tree_size(const clrrt_ctx* ctx) { return ctx ? ctx->n_tree : CLRRT_ERR_ARG; }

int clrrt_tree_truncate(clrrt_ctx* ctx, int n) {
  if (!ctx || n < 1 || n > ctx->n_tree) return CLRRT_ERR_ARG;
  ctx->n_tree = n;
  return CLRRT_OK;
}

int clrrt_tree_download_range(clrrt_ctx* ctx, int first, int n, clrrt_node* host) {
  if (!ctx || !host || first < 0 || n < 0 || first + n > ctx->n_tree) return CLRRT_ERR_ARG;
  CK(cudaSetDevice(ctx->device));
  if (n == 0) return CLRRT_OK;
  if (ctx->copy_pending) { CK(cudaEventSynchronize(ctx->ev_copied)); ctx->copy_pending = false; }
  static_assert(sizeof(clrrt_node) == sizeof(NodeRecord), "clrrt_node and NodeRecord share one layout");
  if ((size_t)n > ctx->export_cap) {
    if (ctx->d_export) cudaFree(ctx->d_export);
    ctx->d_export = nullptr;
    ctx->export_cap = 0;
    const size_t cap = std::max<size_t>((size_t)n, 1024) * 2;
    CK(cudaMalloc((void**)&ctx->d_export, cap * sizeof(NodeRecord)));
    ctx->export_cap = cap;
  }
  export_nodes_kernel<<<(n + 255) / 256, 256, 0, ctx->stream>>>(ctx->tree, first, n, ctx->d_export);
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(host, ctx->d_export, (size_t)n * sizeof(NodeRecord), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return CLRRT_OK;
}

// Asynchronous form: the export kernel runs on the context's stream (ordered before whatever the caller launches next, so
// the exported range may be truncated and re-grown right away), the device-to-host copy on a second stream, so that it
// overlaps the next round.  `host` should be pinned memory and must stay valid until clrrt_download_wait; at most one
// download is in flight (a second call waits for the first).
int clrrt_tree_download_range_async(clrrt_ctx* ctx, int first, int n, clrrt_node* host) {
  if (!ctx || !host || first < 0 || n < 0 || first + n > ctx->n_tree) return CLRRT_ERR_ARG;
  CK(cudaSetDevice(ctx->device));
  if (!ctx->copy_stream) {
    CK(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
    CK(cudaEventCreateWithFlags(&ctx->ev_export, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&ctx->ev_copied, cudaEventDisableTiming));
  }
  if (ctx->copy_pending) { CK(cudaEventSynchronize(ctx->ev_copied)); ctx->copy_pending = false; }
  if (n == 0) return CLRRT_OK;
  if ((size_t)n > ctx->export_cap) {
    if (ctx->d_export) cudaFree(ctx->d_export);
    ctx->d_export = nullptr;
    ctx->export_cap = 0;
    const size_t cap = std::max<size_t>((size_t)n, 1024) * 2;
    CK(cudaMalloc((void**)&ctx->d_export, cap * sizeof(NodeRecord)));
    ctx->export_cap = cap;
  }
  export_nodes_kernel<<<(n + 255) / 256, 256, 0, ctx->stream>>>(ctx->tree, first, n, ctx->d_export);
  CK(cudaGetLastError());
  CK(cudaEventRecord(ctx->ev_export, ctx->stream));
  CK(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_export, 0));
  CK(cudaMemcpyAsync(host, ctx->d_export, (size_t)n * sizeof(NodeRecord), cudaMemcpyDeviceToHost, ctx->copy_stream));
  CK(cudaEventRecord(ctx->ev_copied, ctx->copy_stream));
  ctx->copy_pending = true;
  return CLRRT_OK;
}
int clrrt_download_wait(clrrt_ctx* ctx) {
  if (!ctx) return CLRRT_ERR_ARG;
  if (ctx->copy_pending) { CK(cudaEventSynchronize(ctx->ev_copied)); ctx->copy_pending = false; }
  return CLRRT_OK;
}

int clrrt_tree_download(clrrt_ctx* ctx, clrrt_node* host, int cap, int* n_out) {
  if (!ctx || !host || cap < 0) return CLRRT_ERR_ARG;
  const int n = std::min(cap, ctx->n_tree);
  if (n_out) *n_out = n;
  return clrrt_tree_download_range(ctx, 0, n, host);
}

int clrrt_draw_samples(const double goal[4], int K, double* sample_xy, uint8_t* heuristic) {
  if (!goal || !sample_xy || !heuristic || K < 0) return CLRRT_ERR_ARG;
  const double pi = M_PI;
  for (int j = 0; j < K; j++) {
    // sampleAroundVehicle, rrtplanner.cpp:187-201
    const double dGoal = sqrt(goal[0] * goal[0] + goal[1] * goal[1]);
    const double goalHeading = atan2(goal[1], goal[0]);
    const double latMin = -7, latMax = 7;
    const double rLong = static_cast<float>(rand()) / (static_cast<float>(RAND_MAX / (dGoal + 10)));
    const double rLat = latMin + static_cast<float>(rand()) / (static_cast<float>(RAND_MAX / (latMax - latMin)));
    sample_xy[2 * j] = rLong * cos(goalHeading) + rLat * cos(goalHeading + pi / 2);
    sample_xy[2 * j + 1] = rLong * sin(goalHeading) + rLat * sin(goalHeading + pi / 2);
    // heuristic draw, rrtplanner.cpp:142-143 (RRT.goalReached is never set: threshold stays 0.7)
    const double r = static_cast<double>(rand()) / (static_cast<double>(RAND_MAX / (1)));
    heuristic[j] = (r <= 0.7) ? 0 : 1;
  }
  return CLRRT_OK;
}

// K = 1 with the reference's order of equal keys (nearest.cuh, tie_check_kernel): when the list found on the device
// depends on how equal keys are ordered, repeat the reference's std::sort call (rrtplanner.cpp:233, :256) — the same
// libstdc++ routine on the same (node id, key) pairs in the same initial order — and walk the result as upstream does.
static int nearest_reference_ties(clrrt_ctx* ctx, const double* d_samples, const uint8_t* d_heur, int32_t* d_cand, float* d_key,
                                  int32_t* d_count) {
  cudaStream_t st = ctx->stream;
  const int n = ctx->n_tree;
  TieArgs t;
  t.tree = ctx->tree; t.n_nodes = n; t.sample_xy = d_samples; t.heuristic = d_heur; t.feas_len = ctx->dprm.feas_len;
  t.cand = d_cand; t.key = d_key; t.count = d_count; t.all_key = ctx->d_all_key; t.all_feas = ctx->d_all_feas; t.flag = ctx->d_ints + 6;
  CK(cudaMemsetAsync(ctx->d_ints + 6, 0, sizeof(int32_t), st));
  tie_check_kernel<<<(n + 127) / 128, 128, 0, st>>>(t);
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(ctx->h_ints + 6, ctx->d_ints + 6, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  if (ctx->h_ints[6] == 0) return CLRRT_OK;
  ctx->tie_sorts++;
  std::vector<float> key((size_t)n);
  std::vector<uint8_t> feas((size_t)n);
  CK(cudaMemcpyAsync(key.data(), ctx->d_all_key, (size_t)n * sizeof(float), cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(feas.data(), ctx->d_all_feas, (size_t)n, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  std::vector<std::pair<int, float>> dVector;
  for (int nodeid = 0; nodeid != n; nodeid++) dVector.push_back(std::make_pair(nodeid, key[(size_t)nodeid]));
  std::sort(dVector.begin(), dVector.end(), [](const std::pair<int, float>& a, const std::pair<int, float>& b) { return a.second < b.second; });
  int32_t cand[CLRRT_SORT_LIMIT];
  float ckey[CLRRT_SORT_LIMIT];
  int32_t cnt = 0;
  for (int r = 0; r < CLRRT_SORT_LIMIT; r++) { cand[r] = -1; ckey[r] = 0.0f; }
  for (const auto& e : dVector) {
    if (feas[(size_t)e.first]) { cand[cnt] = e.first; ckey[cnt] = e.second; cnt++; }
    if (cnt == CLRRT_SORT_LIMIT) break;
  }
  CK(cudaMemcpyAsync(d_cand, cand, sizeof cand, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(d_key, ckey, sizeof ckey, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(d_count, &cnt, sizeof cnt, cudaMemcpyHostToDevice, st));
  CK(cudaStreamSynchronize(st));  // the sources are on this stack frame
  return CLRRT_OK;
}

// Sequential windows: sample 0 of a window has a candidate whose key equals the winner's (or the last place's).  The
// device orders equal keys by node id, the reference by whatever its std::sort leaves.  This repeats the reference's sort
// for that one sample (all keys and feasibility flags from tie_check_kernel, the very std::sort call of
// nearest_reference_ties) and reports whether the OUTCOME is the same: the same candidates in the same order up to and
// including the winner — or, when no candidate succeeded, the same set of candidates.  Then the window's rollouts stand
// and only the commit is repeated; otherwise the K = 1 path runs the sample with the reference's list.
static int window_tie_same_outcome(clrrt_ctx* ctx, const double* d_s, const uint8_t* d_h, int j, bool* same) {
  cudaStream_t st = ctx->stream;
  const int n = ctx->n_tree;
  *same = false;
  int32_t cand[CLRRT_SORT_LIMIT], cnt = 0;
  uint32_t word = 0;
  TieArgs t;
  // (sample j of the window: its list was made against the tree at the start of the window; the nodes appended since are
  // known not to come before its winner — seq_commit_kernel's conflict test — and the sort below runs over all of them)
  t.tree = ctx->tree; t.n_nodes = n; t.sample_xy = d_s + 2 * (size_t)j; t.heuristic = d_h + j; t.feas_len = ctx->dprm.feas_len;
  t.cand = ctx->d_cand + (size_t)j * CLRRT_SORT_LIMIT; t.key = ctx->d_key + (size_t)j * CLRRT_SORT_LIMIT; t.count = ctx->d_count + j;
  t.all_key = ctx->d_all_key; t.all_feas = ctx->d_all_feas; t.flag = ctx->d_ints + 6;
  tie_check_kernel<<<(n + 127) / 128, 128, 0, st>>>(t);
  CK(cudaGetLastError());
  std::vector<float> key((size_t)n);
  std::vector<uint8_t> feas((size_t)n);
  CK(cudaMemcpyAsync(key.data(), ctx->d_all_key, (size_t)n * sizeof(float), cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(feas.data(), ctx->d_all_feas, (size_t)n, cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(cand, ctx->d_cand + (size_t)j * CLRRT_SORT_LIMIT, sizeof cand, cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(&cnt, ctx->d_count + j, sizeof cnt, cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(&word, ctx->d_done + j, sizeof word, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  ctx->tie_sorts++;
  std::vector<std::pair<int, float>> dVector;
  for (int nodeid = 0; nodeid != n; nodeid++) dVector.push_back(std::make_pair(nodeid, key[(size_t)nodeid]));
  std::sort(dVector.begin(), dVector.end(), [](const std::pair<int, float>& a, const std::pair<int, float>& b) { return a.second < b.second; });
  int32_t ref[CLRRT_SORT_LIMIT], rcnt = 0;
  for (const auto& e : dVector) {
    if (feas[(size_t)e.first]) ref[rcnt++] = e.first;
    if (rcnt == CLRRT_SORT_LIMIT) break;
  }
  if (rcnt != cnt) return CLRRT_OK;
  int sb = -1;
  for (int r = 0; r < cnt; r++)
    if ((word >> (16 + r)) & 1u) { sb = r; break; }
  if (sb >= 0) {
    for (int r = 0; r <= sb; r++)
      if (ref[r] != cand[r]) return CLRRT_OK;
  } else {
    std::sort(ref, ref + rcnt);
    std::sort(cand, cand + cnt);
    for (int r = 0; r < cnt; r++)
      if (ref[r] != cand[r]) return CLRRT_OK;
  }
  *same = true;
  return CLRRT_OK;
}

static int nearest_dev(clrrt_ctx* ctx, const double* d_samples, const uint8_t* d_heur, int K, int32_t* d_cand,
                       float* d_key, int32_t* d_count, bool window = false) {
  const bool sorted = ctx->n_tree > 0 && (ctx->nn_mode == 1 || (ctx->nn_mode == 0 && (double)K * (double)ctx->n_tree >= NN_SORT_MIN_PAIRS));
  const bool ref_ties = K == 1 && ctx->tie_mode == 1 && !window;  // windows flag ties themselves (tie_window_kernel)
  ctx->nn_sorted_last = sorted;
  if (ref_ties && !d_key) d_key = ctx->d_key;
  if (!sorted) {
    NearestArgs a;
    memset(&a, 0, sizeof a);
    a.sample_xy = d_samples; a.heuristic = d_heur; a.K = K; a.n_nodes = ctx->n_tree; a.tree = ctx->tree;
    a.cand = d_cand; a.key = d_key; a.count = d_count; a.feas_len = ctx->dprm.feas_len;
    nearest_topk_kernel<<<(K + NEAREST_WARPS - 1) / NEAREST_WARPS, NEAREST_THREADS, 0, ctx->stream>>>(a);
    CK(cudaGetLastError());
    if (ref_ties) return nearest_reference_ties(ctx, d_samples, d_heur, d_cand, d_key, d_count);
    return CLRRT_OK;
  }
  // 1. sort nodes and samples along the axis of the sampling box (goal bearing, rrtplanner.cpp:188-197): counting sort
  //    over NN_BINS bins covering [-15 m, dGoal + 25 m] (everything outside lands in the end bins)
  cudaStream_t st = ctx->stream;
  NNSortArgs& s = ctx->nn;
  const double gx = ctx->prm.goal[0], gy = ctx->prm.goal[1];
  const double beta = std::atan2(gy, gx), dgoal = std::sqrt(gx * gx + gy * gy);
  s.tree = ctx->tree; s.n_nodes = ctx->n_tree; s.K = K; s.sample_xy = d_samples; s.heuristic = d_heur;
  s.cb = (float)std::cos(beta); s.sb = (float)std::sin(beta);
  s.u0 = -15.0f;
  // large trees: the bins become (axis slab, lateral bin) pairs, so that tiles are compact in both directions and a
  // block can skip the tiles of a slab that lie to the side of its samples (lateral range: the +-7 m sampling band)
  s.nl_log2 = ctx->nn_lateral_log2 >= 0 ? ctx->nn_lateral_log2 : (ctx->n_tree >= 98304 ? 4 : ctx->n_tree >= 24576 ? 2 : 0);
  const float bin_w = (float)((dgoal + 40.0) / (NN_BINS >> s.nl_log2));
  s.inv_bin = 1.0f / bin_w;
  s.v0 = -8.0f;
  const float vbin_w = 16.0f / (float)(1 << s.nl_log2);
  s.inv_vbin = 1.0f / vbin_w;
  s.inv_sbin = (float)((NN_BINS >> NN_SAMPLE_LAT_LOG2) / (dgoal + 40.0));
  s.inv_svbin = (float)(1 << NN_SAMPLE_LAT_LOG2) / 16.0f;
  const int n_el = ctx->n_tree + K, n_tiles = (ctx->n_tree + NEAREST_TILE - 1) / NEAREST_TILE;
  CK(cudaMemsetAsync(s.hist, 0, (NN_HIST_INTS + 4) * sizeof(int32_t), st));
  nn_bin_kernel<<<(n_el + 255) / 256, 256, 0, st>>>(s);
  nn_scan_kernel<<<2, NN_BINS, 0, st>>>(s.hist);
  nn_scatter_kernel<<<(n_el + 255) / 256, 256, 0, st>>>(s);
  nn_tile_kernel<<<n_tiles, NEAREST_TILE, 0, st>>>(s.sbin, s.ce, s.fx, s.fy, s.frx, s.fry, s.fdx, s.fdy, s.cb, s.sb, ctx->n_tree, s.u0, bin_w, s.v0, vbin_w, s.nl_log2, ctx->d_tile_ulo,
                                                   ctx->d_tile_uhi, ctx->d_tile_vlo, ctx->d_tile_vhi, ctx->d_tile_ce, ctx->d_tile_proj, ctx->d_tile_feas, s.hist + NN_HIST_INTS + 1);
  CK(cudaGetLastError());
  // 2. the search
  NearestArgs a;
  a.sample_xy = d_samples; a.heuristic = d_heur; a.K = K; a.n_nodes = ctx->n_tree; a.tree = ctx->tree;
  a.cand = d_cand; a.key = d_key; a.count = d_count; a.feas_len = ctx->dprm.feas_len;
  a.so.n_nodes = ctx->n_tree; a.so.n_tiles = n_tiles; a.so.node_id = s.node_id;
  a.so.nx = s.nx; a.so.ny = s.ny; a.so.rbx = s.rbx; a.so.rby = s.rby; a.so.dpx = s.dpx; a.so.dpy = s.dpy; a.so.ang = s.ang;
  a.so.ca = s.ca; a.so.sa = s.sa; a.so.ce = s.ce;
  a.so.fx = s.fx; a.so.fy = s.fy; a.so.frx = s.frx; a.so.fry = s.fry; a.so.fdx = s.fdx; a.so.fdy = s.fdy;
  a.so.tile_ulo = ctx->d_tile_ulo; a.so.tile_uhi = ctx->d_tile_uhi; a.so.tile_vlo = ctx->d_tile_vlo; a.so.tile_vhi = ctx->d_tile_vhi; a.so.tile_ce = ctx->d_tile_ce; a.so.tile_proj = ctx->d_tile_proj; a.so.tile_feas = ctx->d_tile_feas;
  a.so.sample_id = s.sample_id; a.so.cb = s.cb; a.so.sb = s.sb;
  a.so.bin_end = s.hist; a.so.lead_sum = s.lead_sum; a.so.ce_floor = s.hist + NN_HIST_INTS + 1; a.so.u0 = s.u0; a.so.inv_bin = s.inv_bin; a.so.v0 = s.v0; a.so.inv_vbin = s.inv_vbin;
  a.so.nl_log2 = s.nl_log2;
  const int blocks = (K + NN_SORTED_WARPS - 1) / NN_SORTED_WARPS;
  nearest_sorted_kernel<<<blocks, NN_SORTED_WARPS * 32, 0, ctx->stream>>>(a);
  CK(cudaGetLastError());
  if (ref_ties) return nearest_reference_ties(ctx, d_samples, d_heur, d_cand, d_key, d_count);
  return CLRRT_OK;
}

int clrrt_nearest_batch(clrrt_ctx* ctx, const double* sample_xy, const uint8_t* heuristic, int K, int32_t* cand,
                        float* key, int32_t* count) {
  if (!ctx || !sample_xy || !heuristic || !cand || !count || K < 1) return CLRRT_ERR_ARG;
  if (!ctx->have_tree) return CLRRT_ERR_STATE;
  if (K > ctx->max_round) return CLRRT_ERR_CAPACITY;
  CK(cudaSetDevice(ctx->device));
  int rc = ensure_params(ctx);
  if (rc) return rc;
  CK(cudaMemcpyAsync(ctx->d_samples, sample_xy, (size_t)K * 16, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(ctx->d_heur, heuristic, (size_t)K, cudaMemcpyHostToDevice, ctx->stream));
  rc = nearest_dev(ctx, ctx->d_samples, ctx->d_heur, K, ctx->d_cand, ctx->d_key, ctx->d_count);
  if (rc) return rc;
  CK(cudaMemcpyAsync(cand, ctx->d_cand, (size_t)K * CLRRT_SORT_LIMIT * 4, cudaMemcpyDeviceToHost, ctx->stream));
  if (key) CK(cudaMemcpyAsync(key, ctx->d_key, (size_t)K * CLRRT_SORT_LIMIT * 4, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaMemcpyAsync(count, ctx->d_count, (size_t)K * 4, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return CLRRT_OK;
}

int clrrt_propagate_batch(clrrt_ctx* ctx, const int32_t* parent, const double* sample_xy, const uint8_t* goal_biased,
                          int M, clrrt_rollout* out, double* traj, int traj_stride) {
  return clrrt_propagate_batch_ex(ctx, parent, sample_xy, goal_biased, M, out, traj, traj_stride, nullptr, 0);
}

int clrrt_propagate_batch_ex(clrrt_ctx* ctx, const int32_t* parent, const double* sample_xy, const uint8_t* goal_biased,
                             int M, clrrt_rollout* out, double* traj, int traj_stride, double* ref_xyv, int ref_stride) {
  if (!ctx || !parent || !sample_xy || !out || M < 1) return CLRRT_ERR_ARG;
  if (traj && traj_stride < 2) return CLRRT_ERR_ARG;
  if (ref_xyv && ref_stride < 1) return CLRRT_ERR_ARG;
  if (!ctx->have_tree) return CLRRT_ERR_STATE;
  for (int i = 0; i < M; i++)
    if (parent[i] < 0 || parent[i] >= ctx->n_tree) return CLRRT_ERR_ARG;
  CK(cudaSetDevice(ctx->device));
  int rc = ensure_params(ctx);
  if (rc) return rc;
  RolloutScratch& b = ctx->batch;
  if (M > b.cap) {
    void* ptrs[] = {b.d_parent, b.d_gb, b.d_samples, b.d_out};
    for (void* p : ptrs) if (p) cudaFree(p);
    CK(cudaMalloc((void**)&b.d_parent, (size_t)M * 4));
    CK(cudaMalloc((void**)&b.d_gb, (size_t)M));
    CK(cudaMalloc((void**)&b.d_samples, (size_t)M * 16));
    CK(cudaMalloc((void**)&b.d_out, (size_t)M * sizeof(clrrt_rollout)));
    b.cap = M;
  }
  const size_t traj_elems = traj ? (size_t)M * traj_stride * 10 : 0;
  if (traj_elems > b.traj_cap) {
    if (b.d_traj) cudaFree(b.d_traj);
    CK(cudaMalloc((void**)&b.d_traj, traj_elems * 8));
    b.traj_cap = traj_elems;
  }
  const size_t ref_elems = ref_xyv ? (size_t)M * ref_stride * 3 : 0;
  if (ref_elems > b.ref_cap) {
    if (b.d_ref) cudaFree(b.d_ref);
    CK(cudaMalloc((void**)&b.d_ref, ref_elems * 8));
    b.ref_cap = ref_elems;
  }
  std::vector<uint8_t> gbf((size_t)M, 0);
  if (goal_biased) for (int i = 0; i < M; i++) gbf[i] = goal_biased[i] ? 1 : 0;
  CK(cudaMemcpyAsync(b.d_parent, parent, (size_t)M * 4, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(b.d_samples, sample_xy, (size_t)M * 16, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(b.d_gb, gbf.data(), (size_t)M, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemsetAsync(ctx->d_ints, 0, 4 * sizeof(int32_t), ctx->stream));
  if (traj) CK(cudaMemsetAsync(b.d_traj, 0, traj_elems * 8, ctx->stream));
  if (ref_xyv) CK(cudaMemsetAsync(b.d_ref, 0, ref_elems * 8, ctx->stream));
  RolloutJob job;
  memset(&job, 0, sizeof job);
  job.n_samples = M; job.n_ranks = 1;
  job.cand = b.d_parent; job.count = nullptr; job.cand_stride = 1; job.sample_xy = b.d_samples;
  job.parents = ctx->tree; job.out_records = b.d_out; job.traj = traj ? b.d_traj : nullptr; job.traj_stride = traj_stride;
  job.ref_out = ref_xyv ? b.d_ref : nullptr; job.ref_stride = ref_stride;
  job.counters = ctx->d_counters; job.refill_min = ctx->refill_min; job.phase_clk = ctx->d_counters + 8;
  job.n_items = M; job.gb_flags = b.d_gb; job.head = ctx->d_ints + 0;
  if ((rc = launch_rollout(ctx, job, M))) return rc;
  CK(cudaMemcpyAsync(out, b.d_out, (size_t)M * sizeof(clrrt_rollout), cudaMemcpyDeviceToHost, ctx->stream));
  if (traj) CK(cudaMemcpyAsync(traj, b.d_traj, traj_elems * 8, cudaMemcpyDeviceToHost, ctx->stream));
  if (ref_xyv) CK(cudaMemcpyAsync(ref_xyv, b.d_ref, ref_elems * 8, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return CLRRT_OK;
}

// == Simulation::Simulation(RRT, state, ref, veh, GoalBiased, genProfile, Vstart), rrt/include/rrt/simulation.h:18-19
int clrrt_simulate_batch(clrrt_ctx* ctx, int M, const double* state10, const int32_t* ref_offset, const double* ref_x,
                         const double* ref_y, double* ref_v, const uint8_t* goal_biased, const uint8_t* gen_profile,
                         const double* Vstart, const int32_t* ref_dir, clrrt_rollout* out, double* traj, int traj_stride) {
  if (!ctx || M < 1 || !state10 || !ref_offset || !ref_x || !ref_y || !ref_v || !gen_profile || !Vstart || !out) return CLRRT_ERR_ARG;
  if (traj && traj_stride < 2) return CLRRT_ERR_ARG;
  if (ref_offset[0] != 0) return CLRRT_ERR_ARG;
  for (int i = 0; i < M; i++) {
    // upstream asserts at least three reference points (reference.cpp:19, :67) and reads three in getLateralError
    if (ref_offset[i + 1] - ref_offset[i] < 3) { ctx->err = "clrrt_simulate: a reference needs at least 3 points"; return CLRRT_ERR_ARG; }
    if (ref_dir && ref_dir[i] != 1 && ref_dir[i] != -1) { ctx->err = "clrrt_simulate: ref.dir must be +1 or -1"; return CLRRT_ERR_ARG; }
  }
  CK(cudaSetDevice(ctx->device));
  int rc = ensure_params(ctx);
  if (rc) return rc;
  const size_t npts = (size_t)ref_offset[M];
  const size_t traj_elems = traj ? (size_t)M * traj_stride * 10 : 0;
  // one scratch allocation per call: this is the fidelity entry point, not the throughput path
  const size_t b_state = (size_t)M * 80, b_off = ((size_t)M + 1) * 4, b_pts = npts * 8, b_flag = (size_t)M, b_vs = (size_t)M * 8,
               b_dir = (size_t)M * 4, b_out = (size_t)M * sizeof(clrrt_rollout);
  auto al = [](size_t b) { return (b + 255) & ~(size_t)255; };
  const size_t total = al(b_state) + al(b_off) + 3 * al(b_pts) + 2 * al(b_flag) + al(b_vs) + al(b_dir) + al(b_out) + al(traj_elems * 8);
  unsigned char* base = nullptr;
  CK(cudaMalloc((void**)&base, total));
  unsigned char* q = base;
  auto take = [&](size_t b) { unsigned char* r = q; q += al(b); return r; };
  double* d_state = (double*)take(b_state); int32_t* d_off = (int32_t*)take(b_off);
  double* d_x = (double*)take(b_pts); double* d_y = (double*)take(b_pts); double* d_v = (double*)take(b_pts);
  uint8_t* d_gb = (uint8_t*)take(b_flag); uint8_t* d_gp = (uint8_t*)take(b_flag);
  double* d_vs = (double*)take(b_vs); int32_t* d_dir = (int32_t*)take(b_dir);
  clrrt_rollout* d_out = (clrrt_rollout*)take(b_out); double* d_traj = (double*)take(traj_elems * 8);
  std::vector<uint8_t> gbf((size_t)M, 0);
  if (goal_biased) for (int i = 0; i < M; i++) gbf[i] = goal_biased[i] ? 1 : 0;
  cudaStream_t st = ctx->stream;
  auto fail = [&](cudaError_t e, const char* what) { ctx->err = std::string(what) + ": " + cudaGetErrorString(e); cudaFree(base); return CLRRT_ERR_CUDA; };
  cudaError_t e;
#define SIM_CK(call) if ((e = (call)) != cudaSuccess) return fail(e, #call)
  SIM_CK(cudaMemcpyAsync(d_state, state10, b_state, cudaMemcpyHostToDevice, st));
  SIM_CK(cudaMemcpyAsync(d_off, ref_offset, b_off, cudaMemcpyHostToDevice, st));
  SIM_CK(cudaMemcpyAsync(d_x, ref_x, b_pts, cudaMemcpyHostToDevice, st));
  SIM_CK(cudaMemcpyAsync(d_y, ref_y, b_pts, cudaMemcpyHostToDevice, st));
  SIM_CK(cudaMemcpyAsync(d_v, ref_v, b_pts, cudaMemcpyHostToDevice, st));  // used as is when genProfile is false
  SIM_CK(cudaMemcpyAsync(d_gb, gbf.data(), b_flag, cudaMemcpyHostToDevice, st));
  SIM_CK(cudaMemcpyAsync(d_gp, gen_profile, b_flag, cudaMemcpyHostToDevice, st));
  SIM_CK(cudaMemcpyAsync(d_vs, Vstart, b_vs, cudaMemcpyHostToDevice, st));
  if (ref_dir) SIM_CK(cudaMemcpyAsync(d_dir, ref_dir, b_dir, cudaMemcpyHostToDevice, st));
  if (traj) SIM_CK(cudaMemsetAsync(d_traj, 0, traj_elems * 8, st));
  SimJob job;
  job.M = M; job.state10 = d_state; job.ref_off = d_off; job.rx = d_x; job.ry = d_y; job.rv = d_v; job.gb = d_gb; job.genp = d_gp;
  job.vstart = d_vs; job.dir = ref_dir ? d_dir : nullptr; job.out = d_out; job.traj = traj ? d_traj : nullptr; job.traj_stride = traj_stride;
  job.counters = ctx->d_counters;
  const int blocks = (M + 127) / 128;
  if (ctx->dprm.exact_dist)
    simulate_kernel<true><<<blocks, 128, 0, st>>>(job, ctx->d_bnd, ctx->d_hot, ctx->d_cold, ctx->d_mov, ctx->d_cell_start, ctx->d_cell_items, ctx->d_pose_cells);
  else
    simulate_kernel<false><<<blocks, 128, 0, st>>>(job, ctx->d_bnd, ctx->d_hot, ctx->d_cold, ctx->d_mov, ctx->d_cell_start, ctx->d_cell_items, ctx->d_pose_cells);
  SIM_CK(cudaGetLastError());
  SIM_CK(cudaMemcpyAsync(out, d_out, b_out, cudaMemcpyDeviceToHost, st));
  SIM_CK(cudaMemcpyAsync(ref_v, d_v, b_pts, cudaMemcpyDeviceToHost, st));
  if (traj) SIM_CK(cudaMemcpyAsync(traj, d_traj, traj_elems * 8, cudaMemcpyDeviceToHost, st));
  SIM_CK(cudaStreamSynchronize(st));
#undef SIM_CK
  cudaFree(base);
  return CLRRT_OK;
}

int clrrt_simulate(clrrt_ctx* ctx, const double* state10, const double* ref_x, const double* ref_y, double* ref_v, int n_ref,
                   int ref_dir, int goal_biased, int gen_profile, double Vstart, clrrt_rollout* out, double* traj, int traj_stride) {
  if (n_ref < 3) return CLRRT_ERR_ARG;
  const int32_t off[2] = {0, n_ref};
  const uint8_t gb = goal_biased ? 1 : 0, gp = gen_profile ? 1 : 0;
  const int32_t dir = ref_dir;
  return clrrt_simulate_batch(ctx, 1, state10, off, ref_x, ref_y, ref_v, &gb, &gp, &Vstart, &dir, out, traj, traj_stride);
}

// == checkObsDistance(states, det, carState), rrt/src/old_collisioncheck.cpp:24-51, for n poses
int clrrt_collide_batch(clrrt_ctx* ctx, const double* pose_xytht, int n, int32_t* verdict, double* dobs) {
  if (!ctx || !pose_xytht || !verdict || n < 1) return CLRRT_ERR_ARG;
  CK(cudaSetDevice(ctx->device));
  int rc = ensure_params(ctx);
  if (rc) return rc;
  double* d_pose = nullptr; int32_t* d_v = nullptr; double* d_d = nullptr;
  CK(cudaMalloc((void**)&d_pose, (size_t)n * 32));
  cudaError_t e = cudaMalloc((void**)&d_v, (size_t)n * 4);
  if (e == cudaSuccess && dobs) e = cudaMalloc((void**)&d_d, (size_t)n * 8);
  auto done = [&](int code) { cudaFree(d_pose); if (d_v) cudaFree(d_v); if (d_d) cudaFree(d_d); return code; };
  if (e != cudaSuccess) { ctx->err = std::string("clrrt_collide_batch: ") + cudaGetErrorString(e); return done(CLRRT_ERR_CUDA); }
  cudaStream_t st = ctx->stream;
  e = cudaMemcpyAsync(d_pose, pose_xytht, (size_t)n * 32, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) {
    collide_batch_kernel<<<(n + 127) / 128, 128, 0, st>>>(d_pose, n, d_v, d_d, ctx->d_bnd, ctx->d_hot, ctx->d_cold, ctx->d_mov,
                                                         ctx->d_cell_start, ctx->d_cell_items, ctx->d_pose_cells);
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaMemcpyAsync(verdict, d_v, (size_t)n * 4, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess && dobs) e = cudaMemcpyAsync(dobs, d_d, (size_t)n * 8, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  if (e != cudaSuccess) { ctx->err = std::string("clrrt_collide_batch: ") + cudaGetErrorString(e); return done(CLRRT_ERR_CUDA); }
  return done(CLRRT_OK);
}

// Records are in sample order, so when the tree is full the prefix that fits is appended (what the sequential
// reference would have added first) and CLRRT_ERR_CAPACITY tells the caller to stop expanding.
static int append_local(clrrt_ctx* ctx, const NodeRecord* d_rec, int n) {
  if (n <= 0) return CLRRT_OK;
  const int fit = std::min(n, ctx->cap - ctx->n_tree);
  if (fit > 0) {
    append_records_kernel<<<(fit + 255) / 256, 256, 0, ctx->stream>>>(ctx->tree, ctx->n_tree, d_rec, fit, ctx->cap);
    CK(cudaGetLastError());
    ctx->n_tree += fit;
  }
  if (fit < n) { ctx->err = "tree capacity exceeded: " + std::to_string(n - fit) + " accepted nodes dropped"; return CLRRT_ERR_CAPACITY; }
  return CLRRT_OK;
}

// Candidate search, rollouts of all candidates with first-success semantics and goal-biased continuations, winner per
// sample: everything of a round up to the staging SoA.  `window`: a speculative window of the sequential mode
// (sequential.cuh) — the candidate keys are kept, the goal-biased rollouts' results are recorded per sample and nothing
// is added to the failure counters (seq_commit_kernel counts the committed samples only).
static int round_core(clrrt_ctx* ctx, const double* d_sample_xy, const uint8_t* d_heuristic, int K, bool window) {
  int rc;
  cudaStream_t st = ctx->stream;
  CK(cudaEventRecord(ctx->ev[0], st));
  // 1. candidate parents
  if ((rc = nearest_dev(ctx, d_sample_xy, d_heuristic, K, ctx->d_cand, window ? ctx->d_key : nullptr, ctx->d_count, window))) return rc;
  CK(cudaEventRecord(ctx->ev[1], st));
  // 2. rollouts of all candidates, rank-major (longest references first within a rank), with early skip: equivalent to
  //    trying them in order until the first success; the goal-biased rollout of a sample's winner continues on the lane
  //    that resolved the sample (rollout.cuh)
  const int n_pairs = K * CLRRT_SORT_LIMIT;
  CK(cudaMemsetAsync(ctx->d_ints, 0, 4 * sizeof(int32_t), st));
  CK(cudaMemsetAsync(ctx->d_done, 0, (size_t)K * sizeof(uint32_t), st));
  CK(cudaMemsetAsync(ctx->d_valid + K, 0, (size_t)K * sizeof(int32_t), st));
  CK(cudaMemsetAsync(ctx->d_hist, 0, 1024 * sizeof(int32_t), st));
  if (window) CK(cudaMemsetAsync(ctx->d_res_code + n_pairs, 0, (size_t)K, st));
  CK(cudaMemcpyAsync(ctx->h_counters + 8, ctx->d_counters, 5 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
  ref_end_kernel<<<(n_pairs + 255) / 256, 256, 0, st>>>(K, CLRRT_SORT_LIMIT, ctx->d_cand, CLRRT_SORT_LIMIT, ctx->d_count, d_sample_xy,
                                                       n_pairs, ctx->tree, ctx->d_ref_end, ctx->d_bucket, ctx->d_hist);
  order_scan_kernel<<<1, 1024, 0, st>>>(ctx->d_hist, CLRRT_SORT_LIMIT * ORDER_BUCKETS, ctx->d_ints + 2);
  order_scatter_kernel<<<(n_pairs + 255) / 256, 256, 0, st>>>(K, CLRRT_SORT_LIMIT, ctx->d_count, ctx->d_bucket, ctx->d_hist, ctx->d_order,
                                                             ctx->nn_sorted_last ? ctx->nn.sample_id : nullptr);
  CK(cudaGetLastError());
  RolloutJob job;
  memset(&job, 0, sizeof job);
  job.n_samples = K; job.n_ranks = CLRRT_SORT_LIMIT; job.n_items = n_pairs; job.n_items_dev = ctx->d_ints + 2;
  job.head = ctx->d_ints + 0; job.cand = ctx->d_cand; job.count = ctx->d_count; job.order = ctx->d_order;
  job.cand_stride = CLRRT_SORT_LIMIT; job.sample_xy = d_sample_xy; job.parents = ctx->tree; job.ref_end = ctx->d_ref_end;
  job.sample_word = ctx->d_done; job.res_code = ctx->d_res_code; job.res_steps = ctx->d_res_steps;
  job.gb_results = window ? 1 : 0;
  job.out_nodes = ctx->stage; job.out_valid = ctx->d_valid;
  job.counters = window ? nullptr : ctx->d_counters; job.refill_min = ctx->refill_min; job.phase_clk = ctx->d_counters + 8;
  job.init = ctx->d_init; job.init_stride = ctx->init_stride;
#ifdef CLRRT_PHASE_CLOCKS
  job.timeline = ctx->d_timeline;
  CK(cudaMemsetAsync(ctx->d_timeline, 0, ((size_t)ctx->max_round * (CLRRT_SORT_LIMIT + 1)) * 3 * sizeof(unsigned long long), st));
#endif
  if (ctx->prm.fp32) setup_kernel<float><<<(n_pairs + 127) / 128, 128, 0, st>>>(job);
  else setup_kernel<double><<<(n_pairs + 127) / 128, 128, 0, st>>>(job);
  CK(cudaGetLastError());
  CK(cudaEventRecord(ctx->ev[5], st));
  if ((rc = launch_rollout(ctx, job, n_pairs))) return rc;
  CK(cudaEventRecord(ctx->ev[2], st));
  SelectArgs sa;
  sa.K = K; sa.n_ranks = CLRRT_SORT_LIMIT; sa.count = ctx->d_count; sa.sample_word = ctx->d_done;
  sa.res_code = ctx->d_res_code; sa.res_steps = ctx->d_res_steps; sa.valid = ctx->d_valid;
  sa.slot = ctx->d_slot; sa.counters = window ? nullptr : ctx->d_counters;
  select_kernel<<<(K + 255) / 256, 256, 0, st>>>(sa);
  CK(cudaGetLastError());
  return CLRRT_OK;
}

int clrrt_expand_round_dev(clrrt_ctx* ctx, const double* d_sample_xy, const uint8_t* d_heuristic, int K,
                           clrrt_round_stats* stats) {
  if (!ctx || !d_sample_xy || !d_heuristic || K < 1) return CLRRT_ERR_ARG;
  if (!ctx->have_tree) return CLRRT_ERR_STATE;
  if (K > ctx->max_round) return CLRRT_ERR_CAPACITY;
  CK(cudaSetDevice(ctx->device));
  int rc = ensure_params(ctx);
  if (rc) return rc;
  cudaStream_t st = ctx->stream;
  if ((rc = round_core(ctx, d_sample_xy, d_heuristic, K, false))) return rc;
  CK(cudaEventRecord(ctx->ev[3], st));
  // 4. compaction in sample order -> records -> append
  const int nblocks = (K + SCAN_THREADS - 1) / SCAN_THREADS;
  scan_block_sums_kernel<<<nblocks, SCAN_THREADS, 0, st>>>(ctx->d_valid, K, ctx->d_block_sums);
  scan_sums_kernel<<<1, SCAN_THREADS, 0, st>>>(ctx->d_block_sums, nblocks, ctx->d_ints + 3);
  pack_records_kernel<<<nblocks, SCAN_THREADS, 0, st>>>(ctx->stage, ctx->d_valid, K, ctx->d_block_sums, ctx->d_records,
                                                         ctx->d_slot, K * CLRRT_SORT_LIMIT);
  CK(cudaGetLastError());
  CK(cudaEventRecord(ctx->ev[6], st));
  const bool exchange = ctx->world > 1 && !ctx->defer_append;
  const NcclApi* nc = nullptr;
  if (exchange) {
    // per-round node all-gather (exchange.cuh): the ranks' record counts first — 4 bytes each, read back together with
    // this rank's counters in the round's one host synchronisation
    if (!(nc = nccl_api(&ctx->err))) return CLRRT_ERR_STATE;
    NCK(nc->AllGather(ctx->d_ints + 3, ctx->d_counts, 1, ncclInt32, ctx->comm, st));
    CK(cudaMemcpyAsync(ctx->h_counts, ctx->d_counts, (size_t)ctx->world * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  }
  CK(cudaMemcpyAsync(ctx->h_ints + 3, ctx->d_ints + 3, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(ctx->h_counters, ctx->d_counters, 5 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  const int n_new = ctx->h_ints[3];
  ctx->last_records = n_new;
  int n_added = n_new;
  int rc_append = CLRRT_OK;  // a full tree still appends the prefix that fits and reports the round (see append_local)
  if (exchange) {
    int stride = 0;
    long long total = 0;
    for (int r = 0; r < ctx->world; r++) { stride = std::max(stride, ctx->h_counts[r]); total += ctx->h_counts[r]; }
    n_added = (int)total;
    if (total > 0) {
      // the records at the stride of the largest chunk, then ONE append launch over all ranks' chunks in rank order
      // (= global sample order, rrtplanner.cpp:150-173), chunk offsets from the gathered counts on the device
      NCK(nc->AllGather(ctx->d_records, ctx->d_gather, (size_t)stride * sizeof(NodeRecord), ncclInt8, ctx->comm, st));
      const int fit = (int)std::min<long long>(total, ctx->cap - ctx->n_tree);
      if (fit > 0) {
        append_gathered_kernel<<<(fit + 255) / 256, 256, 0, st>>>(ctx->tree, ctx->n_tree, ctx->d_gather, stride, ctx->d_counts,
                                                                 ctx->world, fit, ctx->cap);
        CK(cudaGetLastError());
        ctx->n_tree += fit;
      }
      if (fit < total) { ctx->err = "tree capacity exceeded: " + std::to_string(total - fit) + " accepted nodes dropped"; rc_append = CLRRT_ERR_CAPACITY; }
    }
  } else if (!ctx->defer_append) {
    rc_append = append_local(ctx, ctx->d_records, n_new);
  }
  if (rc_append != CLRRT_OK && rc_append != CLRRT_ERR_CAPACITY) return rc_append;
  CK(cudaEventRecord(ctx->ev[4], st));
  CK(cudaEventSynchronize(ctx->ev[4]));
  if (stats) {
    memset(stats, 0, sizeof *stats);
    stats->samples = K;
    stats->rollouts = (int32_t)(ctx->h_counters[4] - ctx->h_counters[12]);
    stats->sim_steps = (int64_t)(ctx->h_counters[3] - ctx->h_counters[11]);
    stats->nodes_added = n_added;
    stats->nodes_local = n_new;
    stats->tree_size = ctx->n_tree;
    cudaEventElapsedTime(&stats->ms_nearest, ctx->ev[0], ctx->ev[1]);
    cudaEventElapsedTime(&stats->ms_prepare, ctx->ev[1], ctx->ev[5]);
    cudaEventElapsedTime(&stats->ms_rollout, ctx->ev[5], ctx->ev[2]);
    cudaEventElapsedTime(&stats->ms_append, ctx->ev[2], ctx->ev[4]);
    cudaEventElapsedTime(&stats->ms_exchange, ctx->ev[6], ctx->ev[4]);
  }
  return rc_append;
}

// == n consecutive calls of expandTree (rrt/src/rrtplanner.cpp:123-174), each seeing the nodes of the calls before it —
// the reference's own sequential algorithm — executed as speculative windows (sequential.cuh)
int clrrt_expand_sequential(clrrt_ctx* ctx, const double* sample_xy, const uint8_t* heuristic, int n, int window,
                            clrrt_seq_stats* stats) {
  if (!ctx || !sample_xy || !heuristic || n < 1 || window < 0) return CLRRT_ERR_ARG;
  if (!ctx->have_tree) return CLRRT_ERR_STATE;
  if (ctx->world > 1) { ctx->err = "clrrt_expand_sequential does not shard (replicas only)"; return CLRRT_ERR_STATE; }
  CK(cudaSetDevice(ctx->device));
  int rc = ensure_params(ctx);
  if (rc) return rc;
  cudaStream_t st = ctx->stream;
  const int wmax = std::min(std::min(window > 0 ? window : 16, SEQ_MAX_WINDOW), ctx->max_round);
  clrrt_seq_stats acc;
  memset(&acc, 0, sizeof acc);
  const auto t_begin = std::chrono::steady_clock::now();
  CK(cudaMemcpyAsync(ctx->h_counters + 8, ctx->d_counters, 5 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  const unsigned long long steps0 = ctx->h_counters[11], roll0 = ctx->h_counters[12];
  int pos = 0, w_next = window > 0 ? wmax : std::min(8, wmax);
  int rc_out = CLRRT_OK;
  while (pos < n) {
    // samples are uploaded in chunks of at most max_round
    const int chunk0 = pos, chunk = std::min(n - pos, ctx->max_round);
    CK(cudaMemcpyAsync(ctx->d_samples, sample_xy + 2 * (size_t)chunk0, (size_t)chunk * 16, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(ctx->d_heur, heuristic + chunk0, (size_t)chunk, cudaMemcpyHostToDevice, st));
    while (pos < chunk0 + chunk) {
      const int w = std::min(w_next, chunk0 + chunk - pos);
      const double* d_s = ctx->d_samples + 2 * (size_t)(pos - chunk0);
      const uint8_t* d_h = ctx->d_heur + (pos - chunk0);
      if ((rc = round_core(ctx, d_s, d_h, w, true))) return rc;
      CK(cudaMemsetAsync(ctx->d_ints + 8, 0, (3 + SEQ_MAX_WINDOW) * sizeof(int32_t), st));
      int32_t* d_tie = ctx->d_ints + 11;
      if (ctx->tie_mode == 1) {
        TieWindowArgs t;
        t.tree = ctx->tree; t.n_nodes = ctx->n_tree; t.sample_xy = d_s; t.heuristic = d_h; t.cand = ctx->d_cand; t.key = ctx->d_key;
        t.count = ctx->d_count; t.feas_len = ctx->dprm.feas_len; t.flag = d_tie;
        tie_window_kernel<<<dim3((ctx->n_tree + 127) / 128, w), 128, 0, st>>>(t);
        CK(cudaGetLastError());
      }
      SeqCommitArgs c;
      c.tree = ctx->tree; c.stage = ctx->stage; c.n_tree = ctx->n_tree; c.capacity = ctx->cap; c.w = w; c.n_ranks = CLRRT_SORT_LIMIT;
      c.j0 = 0; c.n_new0 = 0; c.skip_tie_first = 0;
      c.sample_xy = d_s; c.heuristic = d_h; c.key = ctx->d_key; c.count = ctx->d_count; c.sample_word = ctx->d_done;
      c.valid = ctx->d_valid; c.slot = ctx->d_slot; c.res_code = ctx->d_res_code; c.res_steps = ctx->d_res_steps; c.tie_flag = d_tie;
      c.feas_len = ctx->dprm.feas_len; c.counters = ctx->d_counters; c.out = ctx->d_ints + 8;
      seq_commit_kernel<<<1, SEQ_THREADS, 0, st>>>(c);
      CK(cudaGetLastError());
      CK(cudaEventRecord(ctx->ev[4], st));
      CK(cudaMemcpyAsync(ctx->h_ints + 8, ctx->d_ints + 8, 3 * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
      CK(cudaStreamSynchronize(st));
      {
        float a = 0, b = 0, c2 = 0, d = 0;  // device time of the window's phases (CUDA events of round_core)
        cudaEventElapsedTime(&a, ctx->ev[0], ctx->ev[1]); cudaEventElapsedTime(&b, ctx->ev[1], ctx->ev[5]);
        cudaEventElapsedTime(&c2, ctx->ev[5], ctx->ev[2]); cudaEventElapsedTime(&d, ctx->ev[2], ctx->ev[4]);
        acc.ms_search += a; acc.ms_prepare += b; acc.ms_rollout += c2; acc.ms_commit += d;
      }
      const int n_tree0 = ctx->n_tree;
      int done = ctx->h_ints[8], n_new = ctx->h_ints[9], stop = ctx->h_ints[10];
      ctx->n_tree = n_tree0 + n_new;
      while (stop == 1) {
        // sample `done` of the window: its outcome may hang on the order of equal keys.  If the reference's own sort gives
        // the same outcome, the window's rollouts stand and the commit resumes there with that check waived.
        bool same = false;
        if ((rc = window_tie_same_outcome(ctx, d_s, d_h, done, &same))) return rc;
        if (!same) break;
        CK(cudaMemsetAsync(ctx->d_ints + 8, 0, 3 * sizeof(int32_t), st));
        c.j0 = done; c.n_new0 = n_new; c.skip_tie_first = 1;
        seq_commit_kernel<<<1, SEQ_THREADS, 0, st>>>(c);
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(ctx->h_ints + 8, ctx->d_ints + 8, 3 * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        done += ctx->h_ints[8]; n_new = ctx->h_ints[9]; stop = ctx->h_ints[10];
        ctx->n_tree = n_tree0 + n_new;
        acc.tie_checks_same++;
      }
      acc.windows++; acc.nodes_added += n_new; acc.speculated += w;
      pos += done;
      if (stop == 3) { ctx->err = "tree capacity exceeded"; rc_out = CLRRT_ERR_CAPACITY; pos = n; break; }
      if (stop == 1) {
        // the reference's order of equal keys gives this sample another outcome: the K = 1 path runs it with the reference's list
        clrrt_round_stats st1;
        rc = clrrt_expand_round_dev(ctx, d_s + 2 * (size_t)done, d_h + done, 1, &st1);
        if (rc == CLRRT_ERR_CAPACITY) { rc_out = rc; pos = n; break; }
        if (rc) return rc;
        acc.exact_fallbacks++; acc.nodes_added += st1.nodes_added;
        pos += 1;
      }
      const int committed = done;
      // window size: follow the length of the committed runs (a conflict wastes the rest of the window)
      if (window == 0) w_next = std::max(4, std::min(wmax, 2 * std::max(committed, 1) + 2));
    }
  }
  CK(cudaMemcpyAsync(ctx->h_counters, ctx->d_counters, 5 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  acc.iterations = n;
  acc.sim_steps = (int64_t)(ctx->h_counters[3] - steps0);
  acc.rollouts = (int64_t)(ctx->h_counters[4] - roll0);
  acc.tree_size = ctx->n_tree;
  acc.ms_total = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t_begin).count();
  if (stats) *stats = acc;
  return rc_out;
}

int clrrt_expand_round(clrrt_ctx* ctx, const double* sample_xy, const uint8_t* heuristic, int K,
                       clrrt_round_stats* stats) {
  if (!ctx || !sample_xy || !heuristic || K < 1) return CLRRT_ERR_ARG;
  if (K > ctx->max_round) return CLRRT_ERR_CAPACITY;
  CK(cudaSetDevice(ctx->device));
  CK(cudaMemcpyAsync(ctx->d_samples, sample_xy, (size_t)K * 16, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(ctx->d_heur, heuristic, (size_t)K, cudaMemcpyHostToDevice, ctx->stream));
  return clrrt_expand_round_dev(ctx, ctx->d_samples, ctx->d_heur, K, stats);
}

int clrrt_best_path(clrrt_ctx* ctx, int32_t* ids, int cap, int* n_out) {
  if (!ctx || !ids || !n_out || cap < 1) return CLRRT_ERR_ARG;
  if (!ctx->have_tree) return CLRRT_ERR_STATE;
  CK(cudaSetDevice(ctx->device));
  best_goal_kernel<<<1, 1024, 0, ctx->stream>>>(ctx->tree, ctx->n_tree, ctx->d_ints + 4);
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(ctx->h_ints + 4, ctx->d_ints + 4, sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  const int best = ctx->h_ints[4];
  *n_out = 0;
  if (best < 0) return CLRRT_OK;
  // back-tracking (rrtplanner.cpp:351-358) over the parent array
  std::vector<int32_t> par((size_t)ctx->n_tree);
  CK(cudaMemcpy(par.data(), ctx->tree.parent, (size_t)ctx->n_tree * 4, cudaMemcpyDeviceToHost));
  std::vector<int32_t> chain;
  for (int id = best; id >= 0 && id < ctx->n_tree && (int)chain.size() <= ctx->n_tree; id = par[id]) chain.push_back(id);
  std::reverse(chain.begin(), chain.end());
  *n_out = (int)chain.size();
  for (int i = 0; i < std::min<int>(cap, (int)chain.size()); i++) ids[i] = chain[i];
  return CLRRT_OK;
}

int clrrt_counters_get(clrrt_ctx* ctx, clrrt_counters* out) {
  if (!ctx || !out) return CLRRT_ERR_ARG;
  CK(cudaSetDevice(ctx->device));
  CK(cudaMemcpyAsync(ctx->h_counters, ctx->d_counters, 5 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  out->fail_collision = (int64_t)ctx->h_counters[0];
  out->fail_acclimit = (int64_t)ctx->h_counters[1];
  out->fail_iterlimit = (int64_t)ctx->h_counters[2];
  out->sim_count = (int64_t)ctx->h_counters[3];
  out->rollouts = (int64_t)ctx->h_counters[4];
  return CLRRT_OK;
}

int clrrt_set_tuning(clrrt_ctx* ctx, int refill_min, int blocks_per_sm) {
  if (!ctx || refill_min < 1 || refill_min > 32 || blocks_per_sm < 0) return CLRRT_ERR_ARG;
  ctx->refill_min = refill_min;
  ctx->blocks_override = blocks_per_sm;
  return CLRRT_OK;
}

int clrrt_set_tie_mode(clrrt_ctx* ctx, int mode) {
  if (!ctx || mode < 0 || mode > 1) return CLRRT_ERR_ARG;
  ctx->tie_mode = mode;
  return CLRRT_OK;
}
long long clrrt_tie_sorts(const clrrt_ctx* ctx) { return ctx ? ctx->tie_sorts : -1; }

int clrrt_set_nearest_mode(clrrt_ctx* ctx, int mode) {
  // 0 auto, 1 spatial order always, 2 never; 16 + L: spatial order always with 2^L lateral bins (L = 0..5), for tests
  if (ctx && mode >= 16 && mode <= 21) { ctx->nn_mode = 1; ctx->nn_lateral_log2 = mode - 16; return CLRRT_OK; }
  if (!ctx || mode < 0 || mode > 2) return CLRRT_ERR_ARG;
  ctx->nn_lateral_log2 = -1;
  ctx->nn_mode = mode;
  return CLRRT_OK;
}

int clrrt_set_grid_cell(clrrt_ctx* ctx, double metres) {
  if (!ctx || !(std::fabs(metres) >= 0.05) || std::fabs(metres) > 1000.0) return CLRRT_ERR_ARG;
  ctx->grid_cell = std::fabs(metres);  // takes effect at the next clrrt_set_obstacles
  ctx->pose_enabled = metres > 0;      // negative: position grid only (the fallback path), for tests
  return CLRRT_OK;
}

int clrrt_debug_phase_clocks(clrrt_ctx* ctx, unsigned long long out[16], int reset) {
  if (!ctx || !out) return CLRRT_ERR_ARG;
  CK(cudaSetDevice(ctx->device));
  CK(cudaStreamSynchronize(ctx->stream));
  CK(cudaMemcpy(out, ctx->d_counters + 8, 16 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
  if (reset) CK(cudaMemset(ctx->d_counters + 8, 0, 16 * sizeof(unsigned long long)));
  return CLRRT_OK;
}

// self-check of the branch-free division (rollout.cuh, div_nb) against the operator: see div_check_kernel
int clrrt_debug_div_check(clrrt_ctx* ctx, unsigned long long seed, int pairs_per_thread, unsigned long long out[3]) {
  if (!ctx || !out || pairs_per_thread < 1) return CLRRT_ERR_ARG;
  CK(cudaSetDevice(ctx->device));
  unsigned long long* d = ctx->d_counters + 24;
  CK(cudaMemsetAsync(d, 0, 3 * sizeof(unsigned long long), ctx->stream));
  div_check_kernel<<<ctx->num_sms * 8, 256, 0, ctx->stream>>>(seed, pairs_per_thread, d);
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(out, d, 3 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return CLRRT_OK;
}

// diagnostic builds (-DCLRRT_PHASE_CLOCKS) only: start/end/steps of every rollout of the last round, 3 words per staging slot
int clrrt_debug_timeline(clrrt_ctx* ctx, unsigned long long* out, int K) {
  if (!ctx || !out || K < 1 || K > ctx->max_round) return CLRRT_ERR_ARG;
  if (!ctx->d_timeline) return CLRRT_ERR_STATE;
  CK(cudaSetDevice(ctx->device));
  CK(cudaStreamSynchronize(ctx->stream));
  CK(cudaMemcpy(out, ctx->d_timeline, (size_t)K * (CLRRT_SORT_LIMIT + 1) * 3 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
  return CLRRT_OK;
}

int clrrt_set_defer_append(clrrt_ctx* ctx, int defer) {
  if (!ctx) return CLRRT_ERR_ARG;
  ctx->defer_append = defer != 0;
  return CLRRT_OK;
}
int clrrt_round_records(clrrt_ctx* ctx, void** d_records, int* n_records) {
  if (!ctx || !d_records || !n_records) return CLRRT_ERR_ARG;
  *d_records = ctx->d_records;
  *n_records = ctx->last_records;
  return CLRRT_OK;
}
int clrrt_append_records(clrrt_ctx* ctx, const void* d_records, const int32_t* counts, int world, int stride_records) {
  if (!ctx || !d_records || !counts || world < 1 || stride_records < 0) return CLRRT_ERR_ARG;
  CK(cudaSetDevice(ctx->device));
  const NodeRecord* base = reinterpret_cast<const NodeRecord*>(d_records);
  for (int r = 0; r < world; r++) {
    int rc = append_local(ctx, base + (size_t)r * stride_records, counts[r]);
    if (rc) return rc;
  }
  CK(cudaStreamSynchronize(ctx->stream));
  return CLRRT_OK;
}


// ---- multi-GPU inside the library (exchange.cuh) -------------------------------------------------------------------
int clrrt_comm_unique_id(void* id, int bytes) {
  if (!id || bytes < (int)sizeof(ncclUniqueId)) return CLRRT_ERR_ARG;
  const NcclApi* nc = nccl_api(nullptr);
  if (!nc) return CLRRT_ERR_STATE;
  ncclUniqueId u;
  if (nc->GetUniqueId(&u) != ncclSuccess) return CLRRT_ERR_CUDA;
  memcpy(id, &u, sizeof u);
  return CLRRT_OK;
}

static int comm_buffers(clrrt_ctx* ctx) {
  if (!ctx->d_counts) CK(cudaMalloc((void**)&ctx->d_counts, CLRRT_MAX_WORLD * sizeof(int32_t)));
  if (!ctx->h_counts) CK(cudaMallocHost((void**)&ctx->h_counts, CLRRT_MAX_WORLD * sizeof(int32_t)));
  if (ctx->d_gather) { cudaFree(ctx->d_gather); ctx->d_gather = nullptr; }
  CK(cudaMalloc((void**)&ctx->d_gather, (size_t)ctx->world * 2 * (size_t)ctx->max_round * sizeof(NodeRecord)));
  return CLRRT_OK;
}

int clrrt_comm_attach(clrrt_ctx* ctx, void* nccl_comm, int rank, int world) {
  if (!ctx || !nccl_comm || world < 1 || world > CLRRT_MAX_WORLD || rank < 0 || rank >= world) return CLRRT_ERR_ARG;
  if (!nccl_api(&ctx->err)) return CLRRT_ERR_STATE;
  CK(cudaSetDevice(ctx->device));
  if (ctx->comm && ctx->own_comm) { const NcclApi* nc = nccl_api(nullptr); if (nc) nc->CommDestroy(ctx->comm); }
  ctx->comm = (ncclComm_t)nccl_comm; ctx->own_comm = false; ctx->rank = rank; ctx->world = world;
  return comm_buffers(ctx);
}

int clrrt_comm_init(clrrt_ctx* ctx, const void* id, int bytes, int rank, int world) {
  if (!ctx || !id || bytes < (int)sizeof(ncclUniqueId) || world < 1 || world > CLRRT_MAX_WORLD || rank < 0 || rank >= world) return CLRRT_ERR_ARG;
  const NcclApi* nc = nccl_api(&ctx->err);
  if (!nc) return CLRRT_ERR_STATE;
  CK(cudaSetDevice(ctx->device));
  ncclUniqueId u;
  memcpy(&u, id, sizeof u);
  ncclComm_t c = nullptr;
  NCK(nc->CommInitRank(&c, world, u, rank));
  if (ctx->comm && ctx->own_comm) nc->CommDestroy(ctx->comm);  // a second init replaces the first communicator
  ctx->comm = c; ctx->own_comm = true; ctx->rank = rank; ctx->world = world;
  return comm_buffers(ctx);
}

int clrrt_comm_info(const clrrt_ctx* ctx, int* rank, int* world) {
  if (!ctx) return CLRRT_ERR_ARG;
  if (rank) *rank = ctx->rank;
  if (world) *world = ctx->world;
  return CLRRT_OK;
}

int clrrt_counters_get_global(clrrt_ctx* ctx, clrrt_counters* out) {
  if (!ctx || !out) return CLRRT_ERR_ARG;
  if (ctx->world == 1) return clrrt_counters_get(ctx, out);
  const NcclApi* nc = nccl_api(&ctx->err);
  if (!nc) return CLRRT_ERR_STATE;
  CK(cudaSetDevice(ctx->device));
  unsigned long long* tmp = ctx->d_counters + 24;  // scratch words of the counter block
  NCK(nc->AllReduce(ctx->d_counters, tmp, 5, ncclUint64, ncclSum, ctx->comm, ctx->stream));
  CK(cudaMemcpyAsync(ctx->h_counters, tmp, 5 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  out->fail_collision = (int64_t)ctx->h_counters[0]; out->fail_acclimit = (int64_t)ctx->h_counters[1];
  out->fail_iterlimit = (int64_t)ctx->h_counters[2]; out->sim_count = (int64_t)ctx->h_counters[3];
  out->rollouts = (int64_t)ctx->h_counters[4];
  return CLRRT_OK;
}

int clrrt_tree_digest(clrrt_ctx* ctx, int first, int count, uint64_t out2[2]) {
  if (!ctx || !out2 || first < 0 || count < 0 || first + count > ctx->n_tree) return CLRRT_ERR_ARG;
  CK(cudaSetDevice(ctx->device));
  unsigned long long* d = ctx->d_counters + 30;
  CK(cudaMemsetAsync(d, 0, 2 * sizeof(unsigned long long), ctx->stream));
  if (count > 0) {
    tree_digest_kernel<<<(count + 255) / 256, 256, 0, ctx->stream>>>(ctx->tree, first, count, d);
    CK(cudaGetLastError());
  }
  CK(cudaMemcpyAsync(ctx->h_counters + 14, d, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  out2[0] = ctx->h_counters[14]; out2[1] = ctx->h_counters[15];
  return CLRRT_OK;
}

}  // extern "C"

#ifdef CLRRT_NN_STATS
extern "C" int clrrt_debug_nn_stats(unsigned long long* out8, int reset) {
  cudaDeviceSynchronize();
  if (cudaMemcpyFromSymbol(out8, g_nn_stats, sizeof(unsigned long long) * 8) != cudaSuccess) return CLRRT_ERR_CUDA;
  if (reset) { unsigned long long z[8] = {0}; cudaMemcpyToSymbol(g_nn_stats, z, sizeof z); }
  return CLRRT_OK;
}
#endif
