// simulate.cuh — the reference's Simulation with its own parameter list, and single-pose collision verdicts (sm_100a).
//
//  * simulate_kernel: Simulation::Simulation(RRT, state, MyReference& ref, veh, GoalBiased, genProfile, Vstart)
//    (rrt/include/rrt/simulation.h:18-19, rrt/src/simulation.cpp:36-143) for an ARBITRARY start state and a caller-owned
//    reference path given as arrays (ref.x, ref.y; ref.v is filled when genProfile is set, exactly what the constructor
//    does to the caller's MyReference).  The path is read point by point as upstream does: findClosestPoint is the
//    reference's full scan over [IDwp, N) (controller.cpp:96-113), because nothing is known about the shape of a caller's
//    path; everything after the waypoint search is the same code the round kernel runs (step_core / step_finish).
//    One thread per rollout; this is the boundary-fidelity entry point, not the throughput path.
//  * collide_batch_kernel: checkObsDistance(states, det, carState) (rrt/src/old_collisioncheck.cpp:24-51) for a batch of
//    poses: the verdict of the product's verdict-only check (pose grid, three-way classification, narrow phase) and,
//    on request, the reference's pseudo-distance from the exact path.  Lets a test place poses within millimetres of
//    touching and compare verdict by verdict.
#pragma once
#include "rollout.cuh"

struct SimJob {
  int32_t M;
  const double* state10;   // [M][10]
  const int32_t* ref_off;  // [M + 1] offsets into ref_x / ref_y / ref_v
  const double* rx;
  const double* ry;
  double* rv;
  const uint8_t* gb;       // [M] GoalBiased
  const uint8_t* genp;     // [M] genProfile
  const double* vstart;    // [M]
  const int32_t* dir;      // [M] MyReference::dir (+1 / -1), or nullptr => 1
  clrrt_rollout* out;      // [M]
  double* traj;            // optional [M][traj_stride][10]
  int32_t traj_stride;
  unsigned long long* counters;
};

// Controller::updateWaypoint (controller.cpp:53-68) on an explicit reference; returns dla
__device__ __forceinline__ double waypoint_arr(LaneT<double>& L, const double* X, const double* Y, int N, int dir, double& px, double& py) {
  const double dla = std_max(c_prm.mindla, c_prm.dla_c + c_prm.tla * fabs(L.v));
  px = L.x + dla * (double)dir * L.cth;
  py = L.y + dla * (double)dir * L.sth;
  // findClosestPoint, controller.cpp:96-113: first minimum of the squared distance over [IDwp, N); idmin starts at 0
  double dmin = INFINITY;
  int idmin = 0;
  for (int i = L.c; i < N; i++) {
    const double di = dist2(X[i], Y[i], px, py);
    if (di < dmin) { dmin = di; idmin = i; }
  }
  L.c = idmin;
  if ((size_t)L.c >= (size_t)N - 1 - 2) L.endreached = true;            // LAlong = 2, :62
  if ((X[L.c] == X[N - 1]) && (Y[L.c] == Y[N - 1])) L.endreached = true;  // :65
  return dla;
}

template <bool EXACT>
__global__ void __launch_bounds__(128)
simulate_kernel(const SimJob job, const ObsBound* __restrict__ g_bnd, const ObsHot* __restrict__ g_hot,
                const ObsCold* __restrict__ g_cold, const ObsMoving* __restrict__ g_mov,
                const int32_t* __restrict__ g_cell_start, const uint16_t* __restrict__ g_cell_items,
                const uint4* __restrict__ g_pose_cells) {
  __shared__ double s_t[128];
  __shared__ float s_vb[4 * VB_FLOATS * 32];
  __shared__ uint32_t s_pairs[4 * PAIR_CAP];
  __shared__ uint32_t s_hit[4];
  __shared__ double s_gb[GBF_COUNT * 128];  // LaneT's goal-bias column (unused here: every reference is an explicit array)
  ObsTables T;
  T.bnd = g_bnd; T.hot = g_hot; T.cold = g_cold; T.mov = g_mov; T.cell_start = g_cell_start; T.cell_items = g_cell_items; T.pose_cells = g_pose_cells;
  const int warp = threadIdx.x >> 5;
  float* vbw = s_vb + warp * VB_FLOATS * 32;
  double* tw = s_t + warp * 32;
  uint32_t* pairs = s_pairs + warp * PAIR_CAP;
  uint32_t* hitword = s_hit + warp;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  bool running = i < job.M;
  LaneT<double> L;
  L.gbx = s_gb + threadIdx.x; L.gb = false; L.item = i; L.rank = 0; L.cnt = 0; L.parent = -1;
  L.N = 3; L.N1 = 3; L.c = 0; L.step = 0; L.idwp0 = 0; L.endreached = false; L.tainted = false;
  L.x = L.y = L.th = L.de = L.v = L.a = L.t = 0; L.cth = 1; L.sth = 0; L.tde = 0; L.vback = 0;
  L.costE = L.costS = L.iE = 0; L.trace = 0; L.vref_log = L.dc_log = 0; L.xb = L.yb = L.ax = L.ay = 0;
  const double* X = nullptr;
  const double* Y = nullptr;
  double* V = nullptr;
  int N = 3, dir = 1;
  double s7_0 = 0;
  if (running) {
    const double* st = job.state10 + (size_t)i * 10;
    L.x = st[0]; L.y = st[1]; L.th = st[2]; L.de = st[3]; L.v = st[4]; L.a = st[5]; L.t = st[6];
    L.vref_log = st[8]; L.dc_log = st[9];
    const int o0 = job.ref_off[i];
    N = job.ref_off[i + 1] - o0;
    X = job.rx + o0; Y = job.ry + o0; V = job.rv + o0;
    dir = job.dir ? job.dir[i] : 1;
    L.N = N; L.N1 = N;
    L.ax = X[0]; L.ay = Y[0]; L.xb = X[N - 1]; L.yb = Y[N - 1];
    L.sx = L.xb; L.sy = L.yb;
    r_sincos(L.th, &L.sth, &L.cth);
    L.tde = r_tan(L.de);
    // Controller control(ref, state): IDwp = 0, updateWaypoint (controller.cpp:23-28); stateArray.back()[7] = IDwp (:41)
    double px, py;
    waypoint_arr(L, X, Y, N, dir, px, py);
    L.idwp0 = L.c;
    s7_0 = (double)L.c;
    if (job.genp[i]) {  // generateVelocityProfile(ref, 0, IDwp, Vstart, vmax, goalPose, GoalBiased), :43
      vprofile_setup<double>(L, job.vstart[i], job.gb[i] != 0);
      for (int k = 0; k < N; k++) V[k] = vprofile(L, k);
    }
    L.vback = V[N - 1];
    if (job.traj) {
      double* row = job.traj + (size_t)i * job.traj_stride * 10;
      row[0] = L.x; row[1] = L.y; row[2] = L.th; row[3] = L.de; row[4] = L.v; row[5] = L.a; row[6] = L.t;
      row[7] = s7_0; row[8] = st[8]; row[9] = st[9];
    }
  }
  int code = 0;
#ifdef CLRRT_PHASE_CLOCKS
  unsigned long long pc_[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  long long pc_t_ = clock64();
#endif
  while (__any_sync(FULL_MASK, running)) {
    StepTmpT<double> tmp;
    tmp.dx2 = tmp.vref = tmp.dcmd = 0;
    if (running) {
      double px, py;
      const double dla = waypoint_arr(L, X, Y, N, dir, px, py);
      // getLateralError, controller.cpp:70-93 ("defined" variant of the :76 test)
      int idmin, idmax;
      if (L.c == 0) { idmin = 0; idmax = 2; }
      else if (L.c == N - 1) { idmin = L.c - 2; idmax = L.c; }
      else { idmin = L.c - 1; idmax = L.c + 1; }
      const double xv[3] = {X[idmin], X[idmin + 1], X[idmax]}, yv[3] = {Y[idmin], Y[idmin + 1], Y[idmax]};
      const double ym = lateral_error_pts<double>(xv, yv, L.cth, L.sth, px, py);
      const double vref = V[min(L.c + 2, N - 1)];
      step_core<double>(L, tmp, dla, ym, vref);
      if (job.traj && L.step < job.traj_stride) {
        double* row = job.traj + ((size_t)i * job.traj_stride + L.step) * 10;
        row[0] = L.x; row[1] = L.y; row[2] = L.th; row[3] = L.de; row[4] = L.v; row[5] = L.a; row[6] = L.t;
        row[7] = (double)L.c; row[8] = tmp.vref; row[9] = tmp.dcmd;
      }
    }
    double Dobs = 100.0;
    if (c_prm.n_static + c_prm.n_moving > 0) {
      if (EXACT) {
        Dobs = obstacle_distance(running, L.x, L.y, L.th, L.cth, L.sth, L.t, T.hot, T.cold, T.mov);
      } else {
        const bool finite = (L.x - L.x) == 0.0 && (L.y - L.y) == 0.0 && (L.th - L.th) == 0.0;
        const bool hit = warp_collide(running && finite, L.x + 1.424 * L.cth, L.y + 1.424 * L.sth, L.th, (float)L.cth, (float)L.sth,
                                      L.t, T, vbw, tw, pairs, hitword
#ifdef CLRRT_PHASE_CLOCKS
                                      , pc_, pc_t_
#endif
                                      );
        if (hit || !finite) Dobs = 0.0;
      }
    }
    if (running) {
      code = step_finish<EXACT>(L, tmp, Dobs);
      if (code != 0) running = false;
    }
  }
  if (i < job.M) {
    const bool success = (code == 4) || (code == 5);
    clrrt_rollout& r = job.out[i];
    r.state[0] = L.x; r.state[1] = L.y; r.state[2] = L.th; r.state[3] = L.de; r.state[4] = L.v; r.state[5] = L.a;
    r.state[6] = L.t; r.state[7] = (double)L.c; r.state[8] = L.vref_log; r.state[9] = L.dc_log;
    r.costE = L.costE; r.costS = L.costS; r.ref_back[0] = L.xb; r.ref_back[1] = L.yb; r.ref_vback = L.vback;
    r.trace = L.trace; r.end_reached = (code == 4); r.goal_reached = (code == 5); r.n_steps = L.step;
    r.fail = success ? 0 : code; r.n_ref = N; r.idwp0 = L.idwp0; r.tainted = L.tainted ? 1 : 0; r.reserved = 0;
    if (job.counters) {  // fail_collision, fail_acclimit, fail_iterlimit, sim_count, rollouts (rrt/src/rrt_node.cpp:21-24)
      if (code >= 1 && code <= 3) atomicAdd(&job.counters[code - 1], 1ull);
      atomicAdd(&job.counters[3], (unsigned long long)L.step);
      atomicAdd(&job.counters[4], 1ull);
    }
  }
}

// verdict[i] = 1 when checkObsDistance would return 0 for the rear-axle pose (x, y, theta) at time t (x[6]); dobs (optional)
// = the value the reference returns (exact path: every obstacle's first separating axis, minimum tracked)
__global__ void __launch_bounds__(128)
collide_batch_kernel(const double* __restrict__ pose4, int n, int32_t* __restrict__ verdict, double* __restrict__ dobs,
                     const ObsBound* __restrict__ g_bnd, const ObsHot* __restrict__ g_hot, const ObsCold* __restrict__ g_cold,
                     const ObsMoving* __restrict__ g_mov, const int32_t* __restrict__ g_cell_start,
                     const uint16_t* __restrict__ g_cell_items, const uint4* __restrict__ g_pose_cells) {
  __shared__ double s_t[128];
  __shared__ float s_vb[4 * VB_FLOATS * 32];
  __shared__ uint32_t s_pairs[4 * PAIR_CAP];
  __shared__ uint32_t s_hit[4];
  ObsTables T;
  T.bnd = g_bnd; T.hot = g_hot; T.cold = g_cold; T.mov = g_mov; T.cell_start = g_cell_start; T.cell_items = g_cell_items; T.pose_cells = g_pose_cells;
  const int warp = threadIdx.x >> 5;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const bool active = i < n;
  double x = 0, y = 0, th = 0, t = 0, sth = 0, cth = 1;
  if (active) {
    x = pose4[4 * (size_t)i]; y = pose4[4 * (size_t)i + 1]; th = pose4[4 * (size_t)i + 2]; t = pose4[4 * (size_t)i + 3];
    r_sincos(th, &sth, &cth);
  }
  int v = 0;
  double d = 100.0;  // shipped stub, rrt/src/collisioncheck.cpp:6-8
  if (c_prm.n_static + c_prm.n_moving > 0) {
    const bool finite = (x - x) == 0.0 && (y - y) == 0.0 && (th - th) == 0.0;
#ifdef CLRRT_PHASE_CLOCKS
    unsigned long long pc_[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    long long pc_t_ = clock64();
#endif
    const bool hit = warp_collide(active && finite, x + 1.424 * cth, y + 1.424 * sth, th, (float)cth, (float)sth, t, T,
                                  s_vb + warp * VB_FLOATS * 32, s_t + warp * 32, s_pairs + warp * PAIR_CAP, s_hit + warp
#ifdef CLRRT_PHASE_CLOCKS
                                  , pc_, pc_t_
#endif
                                  );
    v = (hit || !finite) ? 1 : 0;
    if (dobs) d = obstacle_distance(active, x, y, th, cth, sth, t, T.hot, T.cold, T.mov);
  }
  if (active) {
    verdict[i] = v;
    if (dobs) dobs[i] = d;
  }
}

// div_nb against the division operator on n operand pairs per class (tests/test_gpu_rollout.py): class 0 = the magnitudes
// the rollout divides (1e-6 .. 1e4, both signs), 1 = random bit patterns (every exponent, NaN, Inf, subnormals), 2 = dividend
// or divisor at the edges of the fast path's range.  out[0] = pairs whose fast-path result was accepted, out[1] = accepted
// results that differ from a / b (must be 0), out[2] = pairs sent to the operator.
__device__ __forceinline__ unsigned long long splitmix(unsigned long long& s) {
  unsigned long long z = (s += 0x9e3779b97f4a7c15ull);
  z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
  z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
  return z ^ (z >> 31);
}
__global__ void __launch_bounds__(256) div_check_kernel(unsigned long long seed, int per_thread, unsigned long long* out) {
  unsigned long long s = seed + 0x1234567ull * (blockIdx.x * blockDim.x + threadIdx.x);
  unsigned long long acc = 0, diff = 0, slow = 0;
  for (int it = 0; it < per_thread; it++) {
    const int cls = it % 3;
    double a, b;
    const unsigned long long r0 = splitmix(s), r1 = splitmix(s);
    if (cls == 0) {
      const double ea = -6.0 + 10.0 * (double)(r0 >> 40) / 16777216.0, eb = -6.0 + 10.0 * (double)(r1 >> 40) / 16777216.0;
      a = exp10(ea) * (1.0 + (double)(r0 & 0xffffff) / 16777216.0) * ((r0 >> 30) & 1 ? -1.0 : 1.0);
      b = exp10(eb) * (1.0 + (double)(r1 & 0xffffff) / 16777216.0) * ((r1 >> 30) & 1 ? -1.0 : 1.0);
    } else if (cls == 1) {
      a = __longlong_as_double((long long)r0); b = __longlong_as_double((long long)r1);
    } else {
      const int ka = (int)(r0 % 7), kb = (int)(r1 % 5);
      const double edge[7] = {0.0, 1e-300, 1e-38, 6.6e-37, 1e300, 1.0, 3.0};
      const double edgb[5] = {1e-300, 1e300, 1.0, 7.0, 1e-38};
      a = edge[ka] * (1.0 + (double)(r0 >> 44) / 1048576.0); b = edgb[kb] * (1.0 + (double)(r1 >> 44) / 1048576.0);
    }
    bool bad = false;
    const double q = div_nb(a, b, bad), w = a / b;
    if (bad) slow++;
    else { acc++; if (__double_as_longlong(q) != __double_as_longlong(w)) diff++; }
  }
  atomicAdd(&out[0], acc); atomicAdd(&out[1], diff); atomicAdd(&out[2], slow);
}
