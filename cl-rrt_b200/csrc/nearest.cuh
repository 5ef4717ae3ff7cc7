// nearest.cuh — candidate-parent search: top-10 feasible tree nodes per sample by Dubins key (sm_100a).
//
// Replaces sortNodesExplore / sortNodesOptimize (rrt/src/rrtplanner.cpp:227-268), dubinsDistance (:371-406) and
// feasibleNode (:271-289).  The reference computes a key for EVERY node, fully sorts, then walks the sorted
// list testing feasibility until 10 nodes are found.  The same list is produced here without a sort:
//   block = 8 warps = 8 samples; the node fields are staged tile by tile in shared memory and shared by the
//   8 samples; lanes stride over the tile; infeasible nodes are dropped BEFORE their key is computed (the
//   result is the same: the reference also skips them, after sorting); each lane keeps its own sorted top-10
//   in registers; the 32 lists are merged with warp shuffles.  Ties between equal keys go to the lower node id.
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

__global__ void __launch_bounds__(NEAREST_THREADS) nearest_topk_kernel(const NearestArgs a) {
  __shared__ double s_nx[NEAREST_TILE], s_ny[NEAREST_TILE], s_rbx[NEAREST_TILE], s_rby[NEAREST_TILE];
  __shared__ double s_dpx[NEAREST_TILE], s_dpy[NEAREST_TILE], s_ang[NEAREST_TILE];
  __shared__ float s_ca[NEAREST_TILE], s_sa[NEAREST_TILE], s_ce[NEAREST_TILE];
  __shared__ uint16_t s_idx[NEAREST_WARPS][NEAREST_TILE];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int j = blockIdx.x * NEAREST_WARPS + warp;
  const bool live = j < a.K;
  double sx = 0, sy = 0;
  bool optimize = false;
  if (live) { sx = a.sample_xy[2 * j]; sy = a.sample_xy[2 * j + 1]; optimize = a.heuristic[j] != 0; }
  float k[CLRRT_SORT_LIMIT];
  int id[CLRRT_SORT_LIMIT];
#pragma unroll
  for (int r = 0; r < CLRRT_SORT_LIMIT; r++) { k[r] = INFINITY; id[r] = INT_MAX; }

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
      // pass 1: feasibility of the tile's nodes for this sample; feasible ones are compacted into a per-warp index
      // list with ballot/popc, so that the (much more expensive) key evaluation below runs on full warps
      int cnt = 0;
      for (int i0 = 0; i0 < n; i0 += 32) {
        const int i = i0 + lane;
        const bool f = i < n && feasible_node(sx, sy, s_rbx[i], s_rby[i], s_dpx[i], s_dpy[i], s_ang[i], a.feas_len);
        const unsigned m = __ballot_sync(FULL_MASK, f);
        if (f) s_idx[warp][cnt + __popc(m & ((1u << lane) - 1u))] = (uint16_t)i;
        cnt += __popc(m);
      }
      __syncwarp();
      // pass 2: Dubins keys of the feasible nodes, per-lane sorted top-10 (a lane sees its nodes in increasing id order)
      for (int q = lane; q < cnt; q += 32) {
        const int i = s_idx[warp][q];
        float key = dubins_key(sx, sy, s_nx[i], s_ny[i], s_ca[i], s_sa[i]);
        if (optimize) key = s_ce[i] + key;  // rrtplanner.cpp:254
        if (key < k[CLRRT_SORT_LIMIT - 1]) {
          const int idx = base + i;
          bool placed = false;
#pragma unroll
          for (int r = CLRRT_SORT_LIMIT - 1; r >= 0; r--) {
            if (!placed) {
              if (r > 0 && key < k[r - 1]) { k[r] = k[r - 1]; id[r] = id[r - 1]; }
              else { k[r] = key; id[r] = idx; placed = true; }
            }
          }
        }
      }
      __syncwarp();
    }
  }
  if (!live) return;
  // merge the 32 sorted lists: 10 rounds of warp arg-min on (key, id)
  int cnt = 0;
#pragma unroll 1
  for (int r = 0; r < CLRRT_SORT_LIMIT; r++) {
    float hk = k[0];
    int hid = id[0];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ok = __shfl_xor_sync(FULL_MASK, hk, o);
      const int oid = __shfl_xor_sync(FULL_MASK, hid, o);
      if (ok < hk || (ok == hk && oid < hid)) { hk = ok; hid = oid; }
    }
    if (id[0] == hid && hid != INT_MAX) {
#pragma unroll
      for (int q = 0; q < CLRRT_SORT_LIMIT - 1; q++) { k[q] = k[q + 1]; id[q] = id[q + 1]; }
      k[CLRRT_SORT_LIMIT - 1] = INFINITY; id[CLRRT_SORT_LIMIT - 1] = INT_MAX;
    }
    const bool valid = hid != INT_MAX;
    if (valid) cnt++;
    if (lane == 0) {
      a.cand[(size_t)j * CLRRT_SORT_LIMIT + r] = valid ? hid : -1;
      if (a.key) a.key[(size_t)j * CLRRT_SORT_LIMIT + r] = valid ? hk : 0.0f;
    }
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
