// sequential.cuh — the reference's sequential algorithm at more than one sample in flight (sm_100a).
//
// MotionPlanner::planMotion calls expandTree once per sample (rrt/src/motionplanner.cpp:39-43) and every call sees the
// nodes the previous calls appended (rrt/src/rrtplanner.cpp:150-173): K = 1 is the only formulation that equals the
// reference per QUERY.  One sample at a time leaves the GPU waiting on one chain of rollouts, so the
// library runs a WINDOW of the next W samples speculatively against the current tree — candidate search, all candidate
// rollouts, goal-biased continuation: one snapshot round — and then commits the samples in order for as long as the
// speculation provably equals what the sequential loop would have done:
//
//   sample j of the window (tree at window start: n0 nodes; nodes appended by samples 0..j-1 of the window: "new nodes")
//   would have seen the same candidate list up to and including its winner unless a new node is feasible for it
//   (feasibleNode, rrtplanner.cpp:271-289) with a key (dubinsDistance [+ costE], :227-268) at or below the winner's key —
//   or, when no candidate succeeded, at or below the key of the 10th candidate (any feasible new node when the list
//   holds fewer than 10).  Candidates after the winner are never tried (:150-160), so new nodes that sort after it do
//   not matter.  The first sample for which that fails (or whose outcome hangs on the order of two equal keys, which is
//   libstdc++'s std::sort's: tie_window_kernel, seq_commit_kernel) ends the window: it and the samples after it are run
//   again in the next window, against the tree that now contains the committed nodes.  Sample 0 of a window is never
//   speculative.
//
// seq_commit_kernel does this on the device (one block): for each sample in order, every new node is tested in
// parallel (one thread per new node), then the sample's node and its goal-biased child are appended to the tree SoA
// in place (with the derived fields append_records_kernel computes) and the failure counters of exactly the rollouts
// the sequential loop would have run are added.  The host reads back two integers per window.
#pragma once
#include "nearest.cuh"
#include "rollout.cuh"
#include "tree.cuh"

#define SEQ_MAX_WINDOW 64
#define SEQ_THREADS 128   // >= 2 * SEQ_MAX_WINDOW: one thread per node the window may have appended

// flag[j]: which entries of window sample j's candidate list share their key with another feasible node (cf.
// tie_check_kernel, K = 1) — bit r: some feasible node other than entry r has entry r's key; bit 10: a feasible node
// OUTSIDE the list has the key of the last entry (it could have taken that place); bit 11: a NaN key.  The reference's
// order of equal keys is libstdc++'s std::sort's; whether that order can change the OUTCOME is decided in
// seq_commit_kernel, where the winner is known.  grid = (ceil(n_nodes / 128), w).
#define SEQ_TIE_OUTSIDE (1 << CLRRT_SORT_LIMIT)
#define SEQ_TIE_NAN (2 << CLRRT_SORT_LIMIT)
struct TieWindowArgs {
  NodeSoA tree;
  int32_t n_nodes;
  const double* sample_xy;
  const uint8_t* heuristic;
  const int32_t* cand;   // [w][10]
  const float* key;      // [w][10]
  const int32_t* count;  // [w]
  double feas_len;
  int32_t* flag;         // [w], zeroed
};
__global__ void __launch_bounds__(128) tie_window_kernel(const TieWindowArgs a) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y;
  if (i >= a.n_nodes) return;
  const int cnt = a.count[j];
  if (cnt == 0) return;
  const double sx = a.sample_xy[2 * j], sy = a.sample_xy[2 * j + 1];
  float key = dubins_key(sx, sy, a.tree.x[i], a.tree.y[i], a.tree.ca[i], a.tree.sa[i]);
  if (a.heuristic[j]) key = a.tree.costE[i] + key;  // rrtplanner.cpp:254
  const float* kj = a.key + (size_t)j * CLRRT_SORT_LIMIT;
  const int32_t* cj = a.cand + (size_t)j * CLRRT_SORT_LIMIT;
  int bits = !(key == key) ? SEQ_TIE_NAN : 0;
  bool listed = false;
  for (int r = 0; r < cnt; r++) {
    if (cj[r] == i) listed = true;
    else if (key == kj[r]) bits |= 1 << r;
  }
  if (!listed && (bits & (1 << (CLRRT_SORT_LIMIT - 1)))) bits |= SEQ_TIE_OUTSIDE;
  if (!bits) return;
  const double rbx = a.tree.rbx[i], rby = a.tree.rby[i];
  if (feasible_node(sx, sy, rbx, rby, rbx - a.tree.rfx[i], rby - a.tree.rfy[i], a.tree.angPar[i], a.feas_len)) atomicOr(&a.flag[j], bits);
}

struct SeqCommitArgs {
  NodeSoA tree, stage;
  int32_t n_tree, capacity, w, n_ranks;   // n_tree: size of the tree the window was speculated against
  int32_t j0, n_new0;          // resume: first sample to commit, nodes the window has appended so far (0, 0 at first)
  int32_t skip_tie_first;      // 1: the host has checked sample j0's equal keys against the reference's order (same outcome)
  const double* sample_xy;
  const uint8_t* heuristic;
  const float* key;            // [w][10]
  const int32_t* count;        // [w]
  const uint32_t* sample_word; // [w]: bit 16 + r = candidate r succeeded
  const int32_t* valid;        // [2 w]: main node / goal-biased child of sample j produced
  const int32_t* slot;         // [w]: staging slot of the winner
  const uint8_t* res_code;     // [w * n_ranks + w]
  const uint16_t* res_steps;
  const int32_t* tie_flag;     // [w]
  double feas_len;
  unsigned long long* counters;  // fail_collision, fail_acclimit, fail_iterlimit, sim_count, rollouts
  int32_t* out;                // [0] samples committed, [1] nodes appended, [2] 1 = stopped by a tie, 2 = by a conflict, 3 = tree full
};

__device__ __forceinline__ void seq_append(const NodeSoA& t, int k, const NodeSoA& S, int s, int parent) {
  t.x[k] = S.x[s]; t.y[k] = S.y[s]; t.th[k] = S.th[s]; t.de[k] = S.de[s]; t.v[k] = S.v[s]; t.a[k] = S.a[s]; t.t[k] = S.t[s];
  t.s7[k] = S.s7[s]; t.s8[k] = S.s8[s]; t.s9[k] = S.s9[s];
  t.rfx[k] = S.rfx[s]; t.rfy[k] = S.rfy[s]; t.rbx[k] = S.rbx[s]; t.rby[k] = S.rby[s]; t.vback[k] = S.vback[s];
  t.costE[k] = S.costE[s]; t.costS[k] = S.costS[s]; t.parent[k] = parent; t.goal[k] = S.goal[s]; t.nref[k] = S.nref[s];
  t.kind[k] = S.kind[s]; t.smx[k] = S.smx[s]; t.smy[k] = S.smy[s];
  const float ang = (float)(-S.th[s] - M_PI * 0.0);  // rrtplanner.cpp:378
  float sn, cs;
  ref_sincosf(ang, &sn, &cs);
  t.ca[k] = cs; t.sa[k] = sn;
  t.angPar[k] = atan2(S.rby[s] - S.rfy[s], S.rbx[s] - S.rfx[s]);  // rrtplanner.cpp:273
}

__global__ void __launch_bounds__(SEQ_THREADS) seq_commit_kernel(const SeqCommitArgs a) {
  __shared__ int s_conflict;
  __shared__ int s_new;                       // nodes appended so far
  __shared__ unsigned long long s_cnt[5];
  const int tid = threadIdx.x;
  if (tid == 0) { s_conflict = 0; s_new = a.n_new0; }
  if (tid < 5) s_cnt[tid] = 0;
  __syncthreads();
  int committed = 0, stop = 0;
  for (int j = a.j0; j < a.w; j++) {
    const int n_new = s_new;
    // ---- does the speculation of sample j stand? ---------------------------------------------------------------------
    const int cnt = a.count[j];
    const int sb = __ffs(a.sample_word[j] >> 16) - 1;
    const bool won = sb >= 0 && sb < cnt;
    if (j > 0 && n_new > 0) {
      float T;  // a feasible new node with key <= T would have been tried before the speculative outcome was reached
      if (won) T = a.key[(size_t)j * CLRRT_SORT_LIMIT + sb];
      else if (cnt < CLRRT_SORT_LIMIT) T = INFINITY;
      else T = a.key[(size_t)j * CLRRT_SORT_LIMIT + CLRRT_SORT_LIMIT - 1];
      if (tid < n_new) {
        const int k = a.n_tree + tid;
        const double sx = a.sample_xy[2 * j], sy = a.sample_xy[2 * j + 1];
        float key = dubins_key(sx, sy, a.tree.x[k], a.tree.y[k], a.tree.ca[k], a.tree.sa[k]);
        if (a.heuristic[j]) key = a.tree.costE[k] + key;
        if (!(key > T)) {  // also NaN keys
          const double rbx = a.tree.rbx[k], rby = a.tree.rby[k];
          if (feasible_node(sx, sy, rbx, rby, rbx - a.tree.rfx[k], rby - a.tree.rfy[k], a.tree.angPar[k], a.feas_len)) s_conflict = 1;
        }
      }
      __syncthreads();
      if (s_conflict) { stop = 2; break; }
    }
    // Equal keys: the reference's order is std::sort's (host path, K = 1).  The order can only change the outcome when it
    // involves the WINNER (another feasible node with the winner's key may be tried first and succeed too) or, when no
    // candidate succeeded, the last place of a full list (a node outside it with the same key may have taken that place
    // and succeed).  Equal keys among candidates that fail anyway — tried in either order, counted either way — and
    // among candidates after the winner, which are never tried (rrtplanner.cpp:150-160), leave the tree and the counters
    // as they are.
    {
      const int tb = a.tie_flag[j];
      const bool matters = (tb & SEQ_TIE_NAN) || (won ? ((tb >> sb) & 1) : (cnt == CLRRT_SORT_LIMIT && (tb & SEQ_TIE_OUTSIDE)));
      if (matters && !(j == a.j0 && a.skip_tie_first)) { stop = 1; break; }
    }
    // ---- commit: MyRRT::addNode for the winner and its goal-biased child (rrtplanner.cpp:156-158, :169-172) ------------
    const bool vm = a.valid[j] != 0, vg = a.valid[a.w + j] != 0;
    const int add = (int)vm + (int)vg;
    if (a.n_tree + n_new + add > a.capacity) { stop = 3; break; }
    if (tid == 0) {
      int k = a.n_tree + n_new;
      if (vm) { const int s = a.slot[j]; seq_append(a.tree, k, a.stage, s, a.stage.parent[s]); k++; }
      if (vg) { seq_append(a.tree, k, a.stage, a.w * a.n_ranks + j, k - 1); k++; }
      s_new = n_new + add;
      // the counters of exactly the rollouts the sequential loop runs: candidates up to the winner (all of them when none
      // succeeded), and the goal-biased rollout when feasibleGoalBias held
      const int last = won ? sb : cnt - 1;
      for (int r = 0; r <= last; r++) {
        const int code = a.res_code[j * a.n_ranks + r];
        s_cnt[3] += a.res_steps[j * a.n_ranks + r];
        s_cnt[4]++;
        if (code >= 1 && code <= 3) s_cnt[code - 1]++;
      }
      const int gcode = a.res_code[a.w * a.n_ranks + j];
      if (gcode != 0) {
        s_cnt[3] += a.res_steps[a.w * a.n_ranks + j];
        s_cnt[4]++;
        if (gcode >= 1 && gcode <= 3) s_cnt[gcode - 1]++;
      }
    }
    __threadfence_block();
    __syncthreads();
    committed++;
  }
  __syncthreads();
  if (tid == 0) {
    a.out[0] = committed; a.out[1] = s_new; a.out[2] = stop;
    for (int k = 0; k < 5; k++)
      if (s_cnt[k]) atomicAdd(&a.counters[k], s_cnt[k]);
  }
}
