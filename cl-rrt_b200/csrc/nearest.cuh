// nearest.cuh — candidate-parent search: top-10 feasible tree nodes per sample by Dubins key (sm_100a).
//
// Replaces sortNodesExplore / sortNodesOptimize (rrt/src/rrtplanner.cpp:227-268), dubinsDistance (:371-406) and
// feasibleNode (:271-289).  The reference computes a key for EVERY node, fully sorts, then walks the sorted
// list testing feasibility until 10 nodes are found.  The same list is produced here without a sort:
//   block = 8 warps = 8 samples; the node fields are staged tile by tile in shared memory and shared by the
//   8 samples; lanes stride over the tile; nodes that cannot enter the list (distance bound) and infeasible nodes are
//   dropped BEFORE their key is computed (the result is the same: the reference also skips infeasible nodes, after
//   sorting); the warp keeps one sorted top-10 in registers.  Ties between equal keys go to the lower node id.
// Arithmetic types follow dubinsDistance exactly: the whole metric is float (double enters only through
// S - N.state and the M_PI terms); feasibility is double.
#pragma once
#include "common.cuh"
#include "refmath.cuh"

#define NEAREST_WARPS 8
#define NEAREST_THREADS (NEAREST_WARPS * 32)
#define NEAREST_TILE 256

struct NearestArgs {
  const double* sample_xy;
  const uint8_t* heuristic;
  int32_t K, n_nodes;
  NodeSoA tree;
  int32_t* cand;   // [K][10]
  float* key;      // [K][10] (may be nullptr)
  int32_t* count;  // [K]
  double feas_len; // 2.1*ref_res
};

// dubinsDistance(S, N, dir=1), rrtplanner.cpp:371-406, with cos(ang)/sin(ang) of the node precomputed
__device__ __forceinline__ float dubins_key(double sx, double sy, double nx, double ny, float ca, float sa) {
  const float rho = 4.77f;
  float qw_x = (float)(sx - nx);
  float qw_y = (float)(sy - ny);
  const float tmp = ca * qw_x - sa * qw_y;
  qw_y = fabsf(sa * qw_x + ca * qw_y);
  qw_x = tmp;
  const float inner = qw_x * qw_x + (qw_y - rho) * (qw_y - rho);
  const float outer = qw_x * qw_x + (qw_y + rho) * (qw_y + rho);
  const bool q_in_Dp = (outer <= rho * rho) | (inner <= rho * rho);
  if (!q_in_Dp) {
    const float dc = sqrtf(inner);
    float thetac = ref_atan2f(qw_x, rho - qw_y);
    while (thetac < 0) thetac = (float)((double)thetac + 2 * M_PI);
    return sqrtf(dc * dc - rho * rho) + rho * (thetac - ref_acosf(rho / dc));
  } else {
    const float df = sqrtf(outer);
    const float alpha = (float)(2 * M_PI - (double)ref_acosf((5 * rho * rho - df * df) / (4 * rho * rho)));
    return rho * (alpha + ref_asinf(qw_x / df) - ref_asinf(rho * ref_sinf(alpha) / df));
  }
}

__device__ __forceinline__ double nn_angle_diff(double a, double b) {  // functions.h:49-56
  double dif = fmod(b - a + M_PI, 2 * M_PI);
  if (dif < 0) dif += 2 * M_PI;
  return dif - M_PI;
}

// feasibleNode, rrtplanner.cpp:271-289.  The heading test |angleDiff(angNew, angPar)| <= pi/4 is decided from
// dot/cross products when it is not close to the threshold (relative band 1e-9, seven orders above the rounding
// of the reference's atan2 route); inside the band, and for the length test near its threshold, the reference's
// own expressions are evaluated.
__device__ __forceinline__ bool feasible_node(double sx, double sy, double rbx, double rby, double dpx, double dpy,
                                              double angPar, double feas_len) {
  const double dnx = sx - rbx, dny = sy - rby;
  const double dot = dnx * dpx + dny * dpy;
  const double crs = dnx * dpy - dny * dpx;
  const double s = dot - fabs(crs);
  const double band = 1e-9 * (fabs(dot) + fabs(crs));
  if (s < -band) return false;
  if (!(s > band)) {
    const double angNew = atan2(sy - rby, sx - rbx);
    if (fabs(nn_angle_diff(angNew, angPar)) > (M_PI / 4)) return false;
  }
  const double l2 = sq(rbx - sx) + sq(rby - sy);
  const double t2 = feas_len * feas_len;
  if (l2 < t2 * (1 - 1e-12)) return false;
  if (l2 > t2 * (1 + 1e-12)) return true;
  return !(sqrt(l2) < feas_len);
}

// The running top-10 of a sample lives in registers of lanes 0..9 of its warp (lane r = r-th best so far), ordered by
// (key, node id).  Its last entry T bounds the search: the Dubins key is never below the Euclidean distance from
// the node to the sample (checked over the whole float domain of dubinsDistance: key >= (1 - 3e-6) d outside the turning
// circles, key >= 1.58 d inside), so a node with 0.999 d > T (0.999 d + costE > T for the optimise key) cannot enter
// the list and is dropped before its feasibility or its key is evaluated.  Three compaction stages keep the expensive
// parts on full warps: distance bound (all nodes of a tile) -> feasibility (survivors) -> key (feasible survivors).
__global__ void __launch_bounds__(NEAREST_THREADS) nearest_topk_kernel(const NearestArgs a) {
  __shared__ double s_nx[NEAREST_TILE], s_ny[NEAREST_TILE], s_rbx[NEAREST_TILE], s_rby[NEAREST_TILE];
  __shared__ double s_dpx[NEAREST_TILE], s_dpy[NEAREST_TILE], s_ang[NEAREST_TILE];
  __shared__ float s_ca[NEAREST_TILE], s_sa[NEAREST_TILE], s_ce[NEAREST_TILE];
  __shared__ uint16_t s_idx[NEAREST_WARPS][NEAREST_TILE];
  __shared__ uint16_t s_idx2[NEAREST_WARPS][NEAREST_TILE];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const unsigned lt = (1u << lane) - 1u;
  const int j = blockIdx.x * NEAREST_WARPS + warp;
  const bool live = j < a.K;
  double sx = 0, sy = 0;
  bool optimize = false;
  if (live) { sx = a.sample_xy[2 * j]; sy = a.sample_xy[2 * j + 1]; optimize = a.heuristic[j] != 0; }
  float lk = INFINITY;  // entry `lane` of the list (lanes >= 10 stay at +inf / INT_MAX and never take part)
  int lid = INT_MAX;
  float T = INFINITY;   // key and id of the 10th entry (warp-uniform)
  int Tid = INT_MAX;

  for (int base = 0; base < a.n_nodes; base += NEAREST_TILE) {
    const int n = min(NEAREST_TILE, a.n_nodes - base);
    __syncthreads();
    if ((int)threadIdx.x < n) {
      const int g = base + threadIdx.x;
      s_nx[threadIdx.x] = a.tree.x[g]; s_ny[threadIdx.x] = a.tree.y[g];
      const double rbx = a.tree.rbx[g], rby = a.tree.rby[g];
      s_rbx[threadIdx.x] = rbx; s_rby[threadIdx.x] = rby;
      s_dpx[threadIdx.x] = rbx - a.tree.rfx[g]; s_dpy[threadIdx.x] = rby - a.tree.rfy[g];
      s_ang[threadIdx.x] = a.tree.angPar[g];
      s_ca[threadIdx.x] = a.tree.ca[g]; s_sa[threadIdx.x] = a.tree.sa[g]; s_ce[threadIdx.x] = a.tree.costE[g];
    }
    __syncthreads();
    if (live) {
      // stage 1: distance bound against T
      int c1 = 0;
      for (int i0 = 0; i0 < n; i0 += 32) {
        const int i = i0 + lane;
        bool keep = false;
        if (i < n) {
          const float ex = (float)(sx - s_nx[i]), ey = (float)(sy - s_ny[i]);
          const float lb = 0.999f * sqrtf(ex * ex + ey * ey) + (optimize ? s_ce[i] : 0.0f);
          keep = !(lb > T);  // NaN bounds are kept
        }
        const unsigned m = __ballot_sync(FULL_MASK, keep);
        if (keep) s_idx[warp][c1 + __popc(m & lt)] = (uint16_t)i;
        c1 += __popc(m);
      }
      __syncwarp();
      // stage 2: feasibility of the survivors
      int c2 = 0;
      for (int q0 = 0; q0 < c1; q0 += 32) {
        const int q = q0 + lane;
        int i = 0;
        bool f = false;
        if (q < c1) {
          i = s_idx[warp][q];
          f = feasible_node(sx, sy, s_rbx[i], s_rby[i], s_dpx[i], s_dpy[i], s_ang[i], a.feas_len);
        }
        const unsigned m = __ballot_sync(FULL_MASK, f);
        if (f) s_idx2[warp][c2 + __popc(m & lt)] = (uint16_t)i;
        c2 += __popc(m);
      }
      __syncwarp();
      // stage 3: Dubins keys of the feasible survivors (increasing node id) and insertion into the warp's list
      for (int q0 = 0; q0 < c2; q0 += 32) {
        const int q = q0 + lane;
        float key = INFINITY;
        int idx = INT_MAX;
        if (q < c2) {
          const int i = s_idx2[warp][q];
          key = dubins_key(sx, sy, s_nx[i], s_ny[i], s_ca[i], s_sa[i]);
          if (optimize) key = s_ce[i] + key;  // rrtplanner.cpp:254
          idx = base + i;
        }
        unsigned want = __ballot_sync(FULL_MASK, key < T || (key == T && idx < Tid));
        while (want) {
          const int src = __ffs(want) - 1;
          want &= want - 1;
          const float nk = __shfl_sync(FULL_MASK, key, src);
          const int nid = __shfl_sync(FULL_MASK, idx, src);
          if (!(nk < T || (nk == T && nid < Tid))) continue;  // T moved since the ballot
          // position = number of entries ordered before the new one; entries from there on move down one lane
          const bool before = lk < nk || (lk == nk && lid < nid);
          const int pos = __popc(__ballot_sync(FULL_MASK, before && lane < CLRRT_SORT_LIMIT));
          const float uk = __shfl_up_sync(FULL_MASK, lk, 1);
          const int uid = __shfl_up_sync(FULL_MASK, lid, 1);
          if (lane < CLRRT_SORT_LIMIT) {
            if (lane == pos) { lk = nk; lid = nid; }
            else if (lane > pos) { lk = uk; lid = uid; }
          }
          T = __shfl_sync(FULL_MASK, lk, CLRRT_SORT_LIMIT - 1);
          Tid = __shfl_sync(FULL_MASK, lid, CLRRT_SORT_LIMIT - 1);
        }
      }
      __syncwarp();
    }
  }
  if (!live) return;
  const int cnt = __popc(__ballot_sync(FULL_MASK, lane < CLRRT_SORT_LIMIT && lid != INT_MAX));
  if (lane < CLRRT_SORT_LIMIT) {
    const bool valid = lid != INT_MAX;
    a.cand[(size_t)j * CLRRT_SORT_LIMIT + lane] = valid ? lid : -1;
    if (a.key) a.key[(size_t)j * CLRRT_SORT_LIMIT + lane] = valid ? lk : 0.0f;
  }
  if (lane == 0) a.count[j] = cnt;
}

// Per-node quantities the search reads: cosf/sinf of ang = (float)(-theta) (rrtplanner.cpp:378-380) and the heading
// of the node's own reference (rrtplanner.cpp:273).
__global__ void derive_nodes_kernel(NodeSoA t, int first, int n) {
  const int i = first + blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= first + n) return;
  const float ang = (float)(-t.th[i] - M_PI * 0.0);
  float s, c;
  ref_sincosf(ang, &s, &c);
  t.ca[i] = c;
  t.sa[i] = s;
  t.angPar[i] = atan2(t.rby[i] - t.rfy[i], t.rbx[i] - t.rfx[i]);
}
