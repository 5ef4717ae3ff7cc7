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

#ifndef NEAREST_WARPS
#define NEAREST_WARPS 8   // samples per block (16 measured: 1.12 against 1.10 ms on C3)
#endif
#define NEAREST_THREADS (NEAREST_WARPS * 32)
#ifndef NN_SORTED_WARPS
#define NN_SORTED_WARPS 8   // samples per block of the sorted search (C3: 2 / 4 / 8 / 16 -> 0.64 / 0.62 / 0.57 / 0.61 ms)
#endif
#define NEAREST_TILE 256

#define NN_BINS 1024  // spatial bins of the sort (axis slab x lateral bin; the axis is the goal bearing)
// Inside a spatial bin the nodes are ordered by the DIRECTION CLASS of their own reference (angle of dp = rb - rf from the
// axis, NN_FCLS classes of 22.5 degrees), so that a tile of a large tree holds one or a few classes: that is what makes
// the feasibility bound of a tile (nn_tile_kernel) effective.  Node bin = spatial bin * NN_FCLS + class.
#define NN_FCLS 16
#ifndef NN_SAMPLE_LAT_LOG2
#define NN_SAMPLE_LAT_LOG2 4
#endif
#define NN_NODE_BINS (NN_BINS * NN_FCLS)
#define NN_HIST_INTS (NN_NODE_BINS + 2 * NN_BINS)   // node bins; samples with the explore key; samples with the optimise key
__constant__ float2 c_nn_fdir[NN_FCLS] = {   // centre direction of class c: angle -pi + (c + 0.5) 2 pi / 16
    {-0.98078528f, -0.195090322f}, {-0.831469612f, -0.555570233f}, {-0.555570233f, -0.831469612f}, {-0.195090322f, -0.98078528f},
    {0.195090322f, -0.98078528f},  {0.555570233f, -0.831469612f},  {0.831469612f, -0.555570233f},  {0.98078528f, -0.195090322f},
    {0.98078528f, 0.195090322f},   {0.831469612f, 0.555570233f},   {0.555570233f, 0.831469612f},   {0.195090322f, 0.98078528f},
    {-0.195090322f, 0.98078528f},  {-0.555570233f, 0.831469612f},  {-0.831469612f, 0.555570233f},  {-0.98078528f, 0.195090322f}};
__device__ __forceinline__ int nn_dir_class(float dpu, float dpv) {
  const float th = atan2f(dpv, dpu);
  return th == th ? min(max((int)((th + 3.14159265f) * (NN_FCLS / 6.28318531f)), 0), NN_FCLS - 1) : 0;
}
// the sort pays off from about 1e8 (sample, node) pairs per call (measured: C3 2.7e8 pairs 1.59 -> 1.15 ms; 4096 x 4096
// 0.15 -> 0.26 ms); below, all tiles are searched in storage order
#define NN_SORT_MIN_PAIRS 1.0e8

// Spatial order.  Nodes and samples are both sorted along the axis of the sampling box (the goal bearing,
// rrt/src/rrtplanner.cpp:188-197) by a counting sort over NN_BINS bins; the node fields the search reads are copied
// into that order once per round.  A block then takes 8 samples that lie next to each other on the axis (and use the
// same heuristic) and visits the node tiles from their own position outwards; a tile whose axis interval is farther
// than the running 10th key from all 8 samples ends that direction (the axis distance bounds the Euclidean distance,
// which bounds the Dubins key, see nearest_sorted_kernel).  (Measured and rejected: one warp per sample reading the
// sorted arrays directly, without the shared tile: 1.22 instead of 1.13 ms on C3, 7.9 instead of 5.7 ms at 270 k nodes.)  Results do not depend on any of this: the list is the 10 smallest (key, node id).
struct NearestSorted {
  int32_t n_nodes, n_tiles;
  const int32_t* node_id;   // [n_nodes] original id of the node at a sorted position
  const double *nx, *ny, *rbx, *rby, *dpx, *dpy, *ang;   // sorted copies
  const float *ca, *sa, *ce;
  const float *fx, *fy, *frx, *fry, *fdx, *fdy;         // float copies of nx, ny, rbx, rby, dpx, dpy: all a tile stages
  const float* tile_ulo;    // [n_tiles] axis interval of every tile of NEAREST_TILE sorted nodes
  const float* tile_uhi;
  const float* tile_vlo;    // [n_tiles] lateral interval (infinite for tiles that span more than one axis slab)
  const float* tile_vhi;
  const float* tile_ce;     // [n_tiles] smallest costE in the tile (bound for the optimise key costE + Dubins)
  const float* tile_proj;   // [n_tiles][NN_DIRS] projected cost bound of the tile (nn_tile_kernel)
  const float* tile_feas;   // [n_tiles][NN_FCLS] feasibility bound of the tile (nn_tile_kernel)
  const int32_t* sample_id; // [K] original index of the sample at a sorted position
  float cb, sb;             // axis direction
  // where the search of an explore sample starts (see nearest_sorted_kernel): the bin grid of the sort and the mean axis
  // offset from a node to the end of its own reference
  const int32_t* bin_end;   // [NN_NODE_BINS] sorted position after the last node of every node bin
  const float* lead_sum;    // sum over the nodes of (reference end - position) . axis
  const int32_t* ce_floor;  // min(0, smallest costE of the tree) as an order-preserving int (nn_tile_kernel)
  float u0, inv_bin, v0, inv_vbin;
  int32_t nl_log2;
};

struct NearestArgs {
  const double* sample_xy;
  const uint8_t* heuristic;
  int32_t K, n_nodes;
  NodeSoA tree;
  int32_t* cand;   // [K][10]
  float* key;      // [K][10] (may be nullptr)
  int32_t* count;  // [K]
  double feas_len; // 2.1*ref_res
  NearestSorted so;
};

// ---- counting sort of nodes and samples along the axis ------------------------------------------------------------
struct NNSortArgs {
  NodeSoA tree;
  int32_t n_nodes, K;
  const double* sample_xy;
  const uint8_t* heuristic;
  float cb, sb, u0, inv_bin;   // axis: slab = (u - u0) * inv_bin, NN_BINS >> nl_log2 slabs
  float v0, inv_vbin;          // lateral coordinate v = -x sb + y cb: 1 << nl_log2 bins from v0
  float inv_sbin, inv_svbin;   // the same for the SAMPLE bins: NN_BINS >> NN_SAMPLE_LAT_LOG2 slabs x (1 << NN_SAMPLE_LAT_LOG2) lateral bins
  int32_t nl_log2;             // bin = slab << nl_log2 | lateral bin (0: axis only)
  int32_t* bin;        // [n_nodes + K] bin of every element (nodes first)
  int32_t* hist;       // [NN_HIST_INTS]: nodes; samples with the explore key; samples with the optimise key (so that the 8
                       // samples of a block share a heuristic: the two keys prune very differently)
  // outputs of the scatter
  int32_t* node_id;
  double *nx, *ny, *rbx, *rby, *dpx, *dpy, *ang;
  float *ca, *sa, *ce;
  float *fx, *fy, *frx, *fry, *fdx, *fdy;
  int32_t* sbin;       // [n_nodes] bin of the node at a sorted position
  int32_t* sample_id;
  float* lead_sum;     // hist + NN_HIST_INTS (zeroed with it)
};

// float <-> int with the same order (for atomicMin on a float)
__device__ __forceinline__ int nn_ordered(float f) { const int i = __float_as_int(f); return i >= 0 ? i : i ^ 0x7fffffff; }
__device__ __forceinline__ float nn_unordered(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); }

__device__ __forceinline__ int nn_bin_of(float u, float u0, float inv_bin, int nbins) {
  const float b = (u - u0) * inv_bin;
  return b >= 0.0f ? min((int)b, nbins - 1) : 0;  // NaN and out-of-range elements land in the end bins
}

__global__ void __launch_bounds__(256) nn_bin_kernel(const NNSortArgs a) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  int key = -1;
  float lead = 0.0f;
  if (i < a.n_nodes + a.K) {
    float x, y;
    int cls = 0;
    if (i < a.n_nodes) {
      x = (float)a.tree.x[i]; y = (float)a.tree.y[i];
      const double rbx = a.tree.rbx[i], rby = a.tree.rby[i];
      lead = ((float)rbx - x) * a.cb + ((float)rby - y) * a.sb;
      const float dx = (float)(rbx - a.tree.rfx[i]), dy = (float)(rby - a.tree.rfy[i]);
      cls = nn_dir_class(dx * a.cb + dy * a.sb, dy * a.cb - dx * a.sb);
    }
    else { x = (float)a.sample_xy[2 * (i - a.n_nodes)]; y = (float)a.sample_xy[2 * (i - a.n_nodes) + 1]; }
    const float u = x * a.cb + y * a.sb, v = y * a.cb - x * a.sb;
    int b;
    if (i < a.n_nodes) {
      b = (nn_bin_of(u, a.u0, a.inv_bin, NN_BINS >> a.nl_log2) << a.nl_log2) | nn_bin_of(v, a.v0, a.inv_vbin, 1 << a.nl_log2);
      b = b * NN_FCLS + cls;
    } else {
      // samples: always (axis slab, lateral bin) cells, whatever the layout of the node bins — neighbours in the sorted order
      // are neighbours in the plane: the 8 samples of a search block want the same tiles, and the launch order of the
      // rollouts (order_scatter_kernel) follows this order so that the lanes of a warp meet the same obstacles
      b = (nn_bin_of(u, a.u0, a.inv_sbin, NN_BINS >> NN_SAMPLE_LAT_LOG2) << NN_SAMPLE_LAT_LOG2) |
          nn_bin_of(v, a.v0, a.inv_svbin, 1 << NN_SAMPLE_LAT_LOG2);
      if (a.heuristic[i - a.n_nodes]) b += NN_BINS;
    }
    a.bin[i] = b;
    key = (i < a.n_nodes ? 0 : NN_NODE_BINS) + b;
  }
  const unsigned peers = __match_any_sync(FULL_MASK, key);
  if (key >= 0 && (int)(threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(&a.hist[key], __popc(peers));
  if (blockIdx.x * blockDim.x < a.n_nodes) {  // (block-uniform) a hint only: the order of the additions does not matter
    for (int o = 16; o > 0; o >>= 1) lead += __shfl_xor_sync(FULL_MASK, lead, o);
    if ((threadIdx.x & 31) == 0 && lead != 0.0f) atomicAdd(a.lead_sum, lead);
  }
}

// exclusive scans of the histograms (block 0: the NN_NODE_BINS node bins, 16 consecutive bins per thread; block 1: the
// 2 * NN_BINS sample bins, two per thread)
__global__ void __launch_bounds__(NN_BINS) nn_scan_kernel(int32_t* __restrict__ hist) {
  __shared__ int32_t s[NN_BINS];
  int32_t* h = hist + (blockIdx.x == 0 ? NN_FCLS * threadIdx.x : NN_NODE_BINS + 2 * threadIdx.x);
  const int t = threadIdx.x;
  int32_t v[NN_FCLS];
  int sum = 0;
  if (blockIdx.x == 0) {   // 64 bytes per thread: four 16-byte loads
#pragma unroll
    for (int k = 0; k < NN_FCLS / 4; k++) {
      const int4 q = reinterpret_cast<const int4*>(h)[k];
      v[4 * k] = q.x; v[4 * k + 1] = q.y; v[4 * k + 2] = q.z; v[4 * k + 3] = q.w;
    }
  } else {
#pragma unroll
    for (int k = 0; k < NN_FCLS; k++) v[k] = k < 2 ? h[k] : 0;
  }
#pragma unroll
  for (int k = 0; k < NN_FCLS; k++) sum += v[k];
  s[t] = sum;
  __syncthreads();
  for (int o = 1; o < NN_BINS; o <<= 1) {
    const int x = t >= o ? s[t - o] : 0;
    __syncthreads();
    s[t] += x;
    __syncthreads();
  }
  int excl = s[t] - sum;
#pragma unroll
  for (int k = 0; k < NN_FCLS; k++) { const int c = v[k]; v[k] = excl; excl += c; }
  if (blockIdx.x == 0) {
#pragma unroll
    for (int k = 0; k < NN_FCLS / 4; k++) reinterpret_cast<int4*>(h)[k] = make_int4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
  } else {
    h[0] = v[0]; h[1] = v[1];
  }
}

__global__ void __launch_bounds__(256) nn_scatter_kernel(const NNSortArgs a) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  int key = -1;
  if (i < a.n_nodes + a.K) key = (i < a.n_nodes ? 0 : NN_NODE_BINS) + a.bin[i];
  const unsigned peers = __match_any_sync(FULL_MASK, key);
  const int leader = __ffs(peers) - 1;
  int base = 0;
  if (key >= 0 && (int)(threadIdx.x & 31) == leader) base = atomicAdd(&a.hist[key], __popc(peers));
  base = __shfl_sync(FULL_MASK, base, leader);
  if (key < 0) return;
  const int pos = base + __popc(peers & ((1u << (threadIdx.x & 31)) - 1u));
  if (i < a.n_nodes) {
    a.node_id[pos] = i;
    const double x = a.tree.x[i], y = a.tree.y[i], rbx = a.tree.rbx[i], rby = a.tree.rby[i];
    a.nx[pos] = x; a.ny[pos] = y; a.rbx[pos] = rbx; a.rby[pos] = rby;
    const double dpx = rbx - a.tree.rfx[i], dpy = rby - a.tree.rfy[i];
    a.dpx[pos] = dpx; a.dpy[pos] = dpy;
    a.fx[pos] = (float)x; a.fy[pos] = (float)y; a.frx[pos] = (float)rbx; a.fry[pos] = (float)rby;
    a.fdx[pos] = (float)dpx; a.fdy[pos] = (float)dpy;
    a.ang[pos] = a.tree.angPar[i];
    a.ca[pos] = a.tree.ca[i]; a.sa[pos] = a.tree.sa[i]; a.ce[pos] = a.tree.costE[i];
    a.sbin[pos] = a.bin[i];
  } else {
    a.sample_id[pos] = i - a.n_nodes;
  }
}

// Per tile: the axis interval, from the BIN edges of its first and last node (the nodes are sorted by bin, not by
// coordinate: monotone along the tile order by construction; the end bins also hold everything outside the binned
// range, so their outer edges are infinite; 1 cm of slack covers the float rounding of the bin assignment), the
// smallest costE (NaN-safe: a tile with a NaN cost reports -inf and is never skipped), and the PROJECTED bound of the
// optimise key: for any unit vector e,  costE + Dubins(N -> S) >= costE + 0.999 |S - N| >= (costE - 0.999 N.e) + 0.999 S.e,
// so  proj[t][k] = min over the tile of (costE - 0.999 N.e_k)  bounds every key of the tile from below for a sample S by
// proj[t][k] + 0.999 S.e_k — for each direction e_k of a fan of NN_DIRS around the axis, the best of which is the one
// closest to the bearing of S from the tile.  costE is the length driven from the root, so on a dense tree the 10th best
// optimise key of a sample is its straight distance from the root to within millimetres, and the box bound (distance to
// the tile's box + the tile's smallest costE) is loose by the size of the box in every tile near that line; the projected
// bound keeps the correlation between cost and position inside the tile (2.3e5 nodes: 321 -> 47 tiles per sample).
// And the FEASIBILITY bound.  feasibleNode (rrtplanner.cpp:271-289) accepts a node for a sample S only when w = S - rb (rb:
// the end of the node's own reference) lies within 45 degrees of the reference's direction dp.  Every dp of direction class
// c lies within 11.25 degrees of the class centre f_c, so a feasible w lies within 56.25 degrees of f_c and  S . f_c > rb . f_c.
// With  feas[t][c] = min over the tile's nodes of class c of rb . f_c  (+inf for a class the tile does not hold), no node of
// the tile can be feasible for S when  S . f_c < feas[t][c]  for every class, and the tile is skipped.  On a dense tree most
// nodes around a sample are infeasible for it — a node stops some 5 m short of the end of its reference, so the references
// of the nodes near S end beyond S — and the feasible ones sit on the far rim of the disc the distance bound leaves; the
// order by class inside a bin keeps the few nodes whose references point sideways or backwards (which can be feasible from
// anywhere) out of most tiles.  A node with a zero-length or non-finite reference direction makes its tile unskippable.
#define NN_DIRS 9  // -40 .. 40 degrees from the axis
__device__ __forceinline__ void nn_dir(int k, float* c, float* s) {
  const float a = (float)(k - NN_DIRS / 2) * 0.17453293f;
  *c = cosf(a); *s = sinf(a);
}
__global__ void __launch_bounds__(NEAREST_TILE) nn_tile_kernel(const int32_t* __restrict__ sbin, const float* __restrict__ ce,
                                                                const float* __restrict__ fx, const float* __restrict__ fy,
                                                                const float* __restrict__ frx, const float* __restrict__ fry,
                                                                const float* __restrict__ fdx, const float* __restrict__ fdy,
                                                                float cb, float sb,
                                                                int n_nodes, float u0, float bin_w, float v0, float vbin_w,
                                                                int nl_log2, float* __restrict__ ulo, float* __restrict__ uhi,
                                                                float* __restrict__ vlo, float* __restrict__ vhi,
                                                                float* __restrict__ cemin, float* __restrict__ proj,
                                                                float* __restrict__ feas, int32_t* __restrict__ ce_floor) {
  __shared__ float smn[NEAREST_TILE / 32];
  __shared__ float spr[NEAREST_TILE / 32][NN_DIRS];
  __shared__ float sfe[NEAREST_TILE / 32][NN_FCLS];
  __shared__ int s_bad;
  if (threadIdx.x == 0) s_bad = 0;
  __syncthreads();
  const int t = blockIdx.x, i = t * NEAREST_TILE + threadIdx.x;
  float mn = INFINITY;
  float c = 0.0f, u = 0.0f, v = 0.0f;
  if (i < n_nodes) {
    c = ce[i]; mn = c == c ? c : -INFINITY;
    const float x = fx[i], y = fy[i];
    u = x * cb + y * sb; v = y * cb - x * sb;
  }
  for (int o = 16; o > 0; o >>= 1) mn = fminf(mn, __shfl_xor_sync(FULL_MASK, mn, o));
  if ((threadIdx.x & 31) == 0) smn[threadIdx.x >> 5] = mn;
  for (int k = 0; k < NN_DIRS; k++) {
    float dc, ds;
    nn_dir(k, &dc, &ds);
    float g = INFINITY;
    if (i < n_nodes) { g = c - 0.999f * (u * dc + v * ds); if (!(g == g)) g = -INFINITY; }
    for (int o = 16; o > 0; o >>= 1) g = fminf(g, __shfl_xor_sync(FULL_MASK, g, o));
    if ((threadIdx.x & 31) == 0) spr[threadIdx.x >> 5][k] = g;
  }
  {
    int cls = -1;
    float ru = 0.0f, rv = 0.0f;
    if (i < n_nodes) {
      const float rx = frx[i], ry = fry[i], dx = fdx[i], dy = fdy[i];
      ru = rx * cb + ry * sb; rv = ry * cb - rx * sb;
      cls = sbin[i] & (NN_FCLS - 1);   // (the class the sort used: nn_bin_kernel)
      if (!(fabsf(ru) < 1.0e15f) || !(fabsf(rv) < 1.0e15f) || !(fabsf(dx) < 1.0e15f) || !(fabsf(dy) < 1.0e15f) ||
          (dx == 0.0f && dy == 0.0f)) s_bad = 1;
    }
    for (int k = 0; k < NN_FCLS; k++) {
      float m = cls == k ? ru * c_nn_fdir[k].x + rv * c_nn_fdir[k].y : INFINITY;
      if (!(m == m)) m = -INFINITY;
      for (int o = 16; o > 0; o >>= 1) m = fminf(m, __shfl_xor_sync(FULL_MASK, m, o));
      if ((threadIdx.x & 31) == 0) sfe[threadIdx.x >> 5][k] = m;
    }
  }
  __syncthreads();
  if (threadIdx.x >= 32 && threadIdx.x < 32 + NN_FCLS) {
    const int k = threadIdx.x - 32;
    float m = sfe[0][k];
    for (int w = 1; w < NEAREST_TILE / 32; w++) m = fminf(m, sfe[w][k]);
    // (less the float rounding of rb . f_c; an empty class stays at +inf)
    feas[(size_t)t * NN_FCLS + k] = s_bad ? -INFINITY : m == INFINITY ? INFINITY : m - (1.0e-3f + 1.0e-5f * fabsf(m));
  }
  if (threadIdx.x < NN_DIRS) {
    float g = spr[0][threadIdx.x];
    for (int w = 1; w < NEAREST_TILE / 32; w++) g = fminf(g, spr[w][threadIdx.x]);
    // the float rounding of the node's side of the bound (products of magnitudes up to |g| + |N|) is taken off here
    proj[(size_t)t * NN_DIRS + threadIdx.x] = g - (1.0e-3f + 1.0e-5f * fabsf(g));
  }
  if (threadIdx.x == 0) {
    for (int w = 1; w < NEAREST_TILE / 32; w++) mn = fminf(mn, smn[w]);
    const int b0 = sbin[t * NEAREST_TILE], b1 = sbin[min(t * NEAREST_TILE + NEAREST_TILE - 1, n_nodes - 1)];
    const int ns = NN_BINS >> nl_log2, nl = 1 << nl_log2;
    const int p0 = b0 / NN_FCLS, p1 = b1 / NN_FCLS;   // spatial bins of the first and the last node
    const int s0 = p0 >> nl_log2, s1 = p1 >> nl_log2, l0 = p0 & (nl - 1), l1 = p1 & (nl - 1);
    ulo[t] = s0 <= 0 ? -INFINITY : u0 + (float)s0 * bin_w - 0.01f;
    uhi[t] = s1 >= ns - 1 ? INFINITY : u0 + (float)(s1 + 1) * bin_w + 0.01f;
    // within one slab the nodes are ordered by lateral bin; a tile that spans several slabs covers every lateral position
    const bool one = s0 == s1 && nl > 1;
    vlo[t] = (!one || l0 <= 0) ? -INFINITY : v0 + (float)l0 * vbin_w - 0.01f;
    vhi[t] = (!one || l1 >= nl - 1) ? INFINITY : v0 + (float)(l1 + 1) * vbin_w + 0.01f;
    cemin[t] = mn;
    // min(0, smallest costE of the tree), order-preserving int (zeroed = 0.0f before the launch): what ends a direction of
    // the optimise-key search must hold for costs below zero too (an uploaded tree may carry any cost)
    if (mn < 0.0f) atomicMin(ce_floor, nn_ordered(mn));
  }
}

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

// Stage 1 of both searches: can the node enter the list at all?  key >= 0.999 d (+ costE), so a node with
// 0.999 d + costE > T is out.  Evaluated on float copies of the positions, squared (no square root, no double arithmetic:
// this test runs for every node of every visited tile): d > (T - costE) / 0.999 + slack, with `slack` = 1 mm + the float
// rounding of the two positions, and one part in 10^6 for the rounding of the squares.  NaNs are kept.
__device__ __forceinline__ bool stage1_keep(float ex, float ey, float ce, float T, float slack) {
  const float R = (T - ce) * (1.0f / 0.999f) + slack;
  const float d2 = __fmaf_rn(ex, ex, ey * ey);
  return !(R < 0.0f) && !(d2 * 0.999999f > R * R);
}

// Float pre-test of feasibleNode (rrtplanner.cpp:271-289) on float copies of the node's reference end (frx, fry) and reference
// direction (fdx, fdy): false = CERTAINLY infeasible in the reference's double arithmetic, true = undecided (the exact test
// runs).  The heading test |angleDiff| <= pi/4 is dot >= |cross| of (sample - reference end) and the reference direction; in
// float the difference dot - |cross| carries at most  P (4 d (|s| + |rb|) + 10 d |dn|)  of error, d = 2^-24, P = |dpx| + |dpy|,
// |s| + |rb| the 1-norms of the two points, |dn| of their difference (rounding of the four inputs, of the subtraction, of two
// products and a sum each) — the test below uses 6.7 d and 33 d.  The length test Lref >= 2.1 ref_res is decided in float 2 %
// below the threshold when the coordinates are small enough for that to be safe.  NaNs and overflow fall to `true`.  Most
// nodes near a sample fail feasibility (the parent's reference must point at the sample within 45 degrees): this keeps the
// double-precision test, and the double node fields, off 95 % of the distance-bound survivors.
__device__ __forceinline__ bool feasible_maybe(float fsx, float fsy, float frx, float fry, float fdx, float fdy, float feas_len2) {
  const float dnx = fsx - frx, dny = fsy - fry;
  const float dot = dnx * fdx + dny * fdy, crs = dnx * fdy - dny * fdx;
  const float P = fabsf(fdx) + fabsf(fdy);
  const float mag = fabsf(fsx) + fabsf(fsy) + fabsf(frx) + fabsf(fry);
  const float E = P * (4.0e-7f * mag + 2.0e-6f * (fabsf(dnx) + fabsf(dny)));
  if (dot - fabsf(crs) < -E) return false;
  if (mag < 4096.0f && dnx * dnx + dny * dny < 0.98f * feas_len2) return false;
  return true;
}

// The running top-10 of a sample lives in registers of lanes 0..9 of its warp (lane r = r-th best so far), ordered by
// (key, node id).  Its last entry T bounds the search: the Dubins key is never below the Euclidean distance from
// the node to the sample (checked over the whole float domain of dubinsDistance: key >= (1 - 3e-6) d outside the turning
// circles, key >= 1.58 d inside), so a node with 0.999 d > T (0.999 d + costE > T for the optimise key) cannot enter
// the list and is dropped before its feasibility or its key is evaluated — and a whole tile is dropped, before it is
// even loaded, when its axis interval is that far from all 8 samples of the block (tiles are visited from the samples'
// own position outwards, so T is tight after the first tile or two).  Three compaction stages keep the expensive
// parts on full warps: distance bound (all nodes of a tile) -> feasibility (survivors) -> key (feasible survivors).
#ifdef CLRRT_NN_STATS  // diagnostic build (scripts/build_variant.sh): what the search spends its time on
__device__ unsigned long long g_nn_stats[8];  // tile steps, tiles loaded, warp-tiles searched, stage-1 survivors, feasible, inserted, explore/optimise samples
#define NN_STAT(k, v) atomicAdd(&g_nn_stats[k], (unsigned long long)(v))
#else
#define NN_STAT(k, v)
#endif
__global__ void __launch_bounds__(NN_SORTED_WARPS * 32) nearest_sorted_kernel(const NearestArgs a) {
  // a tile stages seven floats per node (position, reference end, reference direction, costE): all that the distance bound
  // and the feasibility pre-test read; the double fields of the few survivors come straight from the sorted arrays
  __shared__ float s_fx[NEAREST_TILE], s_fy[NEAREST_TILE], s_ce[NEAREST_TILE];
  __shared__ float s_frx[NEAREST_TILE], s_fry[NEAREST_TILE], s_fdx[NEAREST_TILE], s_fdy[NEAREST_TILE];
  __shared__ uint16_t s_idx[NN_SORTED_WARPS][NEAREST_TILE];
  __shared__ uint16_t s_idx2[NN_SORTED_WARPS][NEAREST_TILE];
  __shared__ int s_start;
  __shared__ float s_sp[NN_SORTED_WARPS][NN_FCLS];   // sample . f_c (+ rounding) for the feasibility bound of a tile
  __shared__ unsigned s_mask[3][2];   // per chunk of 32 tiles: wanted by any sample / axis-open for any sample (3 in rotation)
  const NearestSorted& so = a.so;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const unsigned lt = (1u << lane) - 1u;
  // position in the sorted sample order; last blocks first: the optimise-key samples are sorted after the explore-key ones
  // and take about three times as long each, so they are started first and the short blocks fill in behind them
  const int js = (gridDim.x - 1 - blockIdx.x) * NN_SORTED_WARPS + warp;
  const bool live = js < a.K;
  const int j = live ? so.sample_id[js] : 0;
  double sx = 0, sy = 0;
  float su = 0.0f, sv = 0.0f;
  bool optimize = false;
  if (live) {
    sx = a.sample_xy[2 * j]; sy = a.sample_xy[2 * j + 1]; optimize = a.heuristic[j] != 0;
    su = (float)sx * so.cb + (float)sy * so.sb;
    sv = (float)sy * so.cb - (float)sx * so.sb;
  }
  const float fsx = (float)sx, fsy = (float)sy;
  const float slack = 1.0e-3f + 4.0e-7f * (fabsf(fsx) + fabsf(fsy));
  const float feas_len2 = (float)(a.feas_len * a.feas_len);
  // optimise key: 0.999 (sample . e_k) for direction k = lane of the fan, and the float rounding of the whole bound
  float proj_s = 0.0f;
  if (lane < NN_DIRS) { float c, sn; nn_dir(lane, &c, &sn); proj_s = 0.999f * (su * c + sv * sn); }
  const float proj_tol = 2.0e-3f + 1.0e-5f * (fabsf(su) + fabsf(sv));
  // the least any key can be at axis distance da: 0.999 da for the explore key, 0.999 da + the smallest cost of the tree
  // (0 on any tree the planner grows: the root) for the optimise key
  const float ce_floor = optimize ? nn_unordered(*so.ce_floor) : 0.0f;
  if (lane < NN_FCLS) s_sp[warp][lane] = su * c_nn_fdir[lane].x + sv * c_nn_fdir[lane].y + proj_tol;
  __syncwarp();
  float lk = INFINITY;  // entry `lane` of the list (lanes >= 10 stay at +inf / INT_MAX and never take part)
  int lid = INT_MAX;
  float T = INFINITY;   // key and id of the 10th entry (warp-uniform)
  int Tid = INT_MAX;

  // First tile.  Explore key (Dubins length): a feasible parent is a node whose own reference ends just before the sample
  // (feasibleNode, rrtplanner.cpp:271-289) — and a node stops about one look-ahead distance short of the end of its
  // reference, so the best parents lie that far BEHIND the sample, not around it.  The search starts at the tile of the bin
  // (axis position - mean lead of the tree's nodes, lateral position of the sample) and goes outwards in both directions;
  // started at the sample's own position it scanned some 200 tiles of infeasible nodes with T still infinite (2.3e5 nodes).
  // Optimise key (costE + Dubins length): tile 0, then upwards — the best parents are the nodes near the root, whose cost so
  // far is small (costE + length is about the straight distance from the root for them, and more for every node off that
  // line), and the root is the origin of the axis.  The sort keeps the samples of the two keys apart, so a block holds one
  // kind (the one block at the seam takes the order of its first sample).  Any start gives the same lists: the order never
  // changes a list, only how early T becomes tight.
  if (threadIdx.x == 0) {
    int start = -1;
    if (!optimize) {
      const float lead = *so.lead_sum / (float)so.n_nodes;
      const float us = su - (fabsf(lead) < 1.0e6f ? lead : 0.0f);
      const int b = (nn_bin_of(us, so.u0, so.inv_bin, NN_BINS >> so.nl_log2) << so.nl_log2) |
                    nn_bin_of(sv, so.v0, so.inv_vbin, 1 << so.nl_log2);
      const int first = b > 0 ? so.bin_end[b * NN_FCLS - 1] : 0;
      start = min(max(first, 0) / NEAREST_TILE, so.n_tiles - 1);
    }
    s_start = start;
    for (int k = 0; k < 3; k++) { s_mask[k][0] = 0; s_mask[k][1] = 0; }
  }
  __syncthreads();
  const bool from_root = s_start < 0;
  const int t0 = from_root ? 0 : s_start;
  // Tiles are voted on 32 at a time: lane i of every warp computes the lower bound `lb` of its sample's keys over tile i of
  // the chunk (box bound; for the optimise key also the projected bound), one ballot per warp and one OR over the block
  // give the tiles that ANY sample of the block still needs, and only those are loaded.  (One tile per vote cost two block
  // barriers and some 100 instructions per warp and tile — 740 of them per block at 2.3e5 nodes, nine in ten for a tile
  // nobody needed.)  Chunks alternate between the two directions, nearest chunk first; a direction is finished when the
  // farthest tile of its chunk is beyond T along the axis for every sample (axis intervals are monotone along the tile order
  // and T only shrinks) or the tiles run out.
  float proj_all[NN_DIRS];
#pragma unroll
  for (int k = 0; k < NN_DIRS; k++) proj_all[k] = __shfl_sync(FULL_MASK, proj_s, k);
  bool open_up = true, open_dn = !from_root;   // block-uniform: directions that may still hold candidates
  int cu = 0, cd = 0;                          // chunks done in each direction
  for (int chunk = 0; open_up || open_dn; chunk++) {
    const bool up = open_up && (!open_dn || cu <= cd);
    const int first = up ? t0 + 32 * cu : t0 - 1 - 32 * cd;   // tile of lane 0; lanes go outwards
    if (up) cu++; else cd++;
    const int tl = up ? first + lane : first - lane;
    float lb = INFINITY;
    bool axis_open = false;
    if (live && tl >= 0 && tl < so.n_tiles) {
      // axis distance <= Euclidean distance <= key / 0.999 (+ the tile's smallest costE for the optimise key)
      const float ulo = so.tile_ulo[tl], uhi = so.tile_uhi[tl];
      const float du = fmaxf(fmaxf(ulo - su, su - uhi), 0.0f);
      const float dv = fmaxf(fmaxf(so.tile_vlo[tl] - sv, sv - so.tile_vhi[tl]), 0.0f);
      // (dv == 0 for tiles without a lateral interval)
      lb = 0.999f * sqrtf(du * du + dv * dv);
      if (optimize) {
        lb += so.tile_ce[tl];
        // projected bound of the optimise key (nn_tile_kernel), the best direction of the fan
        const float* pr = so.tile_proj + (size_t)tl * NN_DIRS;
        float bnd = -INFINITY;
#pragma unroll
        for (int k = 0; k < NN_DIRS; k++) bnd = fmaxf(bnd, pr[k] + proj_all[k]);
#ifndef NN_NO_PROJ  // (diagnostic builds switch the bound off)
        lb = fmaxf(lb, bnd - proj_tol);   // (fmaxf drops a NaN operand: the other bound stands)
#endif
      }
      // feasibility bound (nn_tile_kernel): a class whose reference ends are not all beyond the sample?
      {
        const float4* fe = reinterpret_cast<const float4*>(so.tile_feas + (size_t)tl * NN_FCLS);
        bool any = false;
#pragma unroll
        for (int k = 0; k < NN_FCLS / 4; k++) {
          const float4 f = fe[k];
          any |= !(s_sp[warp][4 * k] < f.x) | !(s_sp[warp][4 * k + 1] < f.y) | !(s_sp[warp][4 * k + 2] < f.z) |
                 !(s_sp[warp][4 * k + 3] < f.w);
        }
#ifndef NN_NO_FEAS  // (diagnostic builds switch the bound off)
        if (!any) lb = INFINITY;
#endif
      }
      // how far the tile lies BEYOND the sample in the direction of travel: monotone along that direction whatever the
      // sample's own position is.  (The lateral interval and the costs are not monotone along the order — the root, 30 m
      // behind a sample, is the best parent by the optimise key — so they only skip tiles, they never end a direction.)
      const float da = fmaxf(up ? ulo - su : su - uhi, 0.0f);
      axis_open = !(0.999f * da + ce_floor > T);
    }
    const unsigned wm = __ballot_sync(FULL_MASK, live && tl >= 0 && tl < so.n_tiles && !(lb > T));
    const unsigned om = __ballot_sync(FULL_MASK, axis_open);
    unsigned* mk = s_mask[chunk % 3];
    if (lane == 0) { if (wm) atomicOr(&mk[0], wm); if (om) atomicOr(&mk[1], om); }
    __syncthreads();
    unsigned want_any = mk[0];
    const unsigned open_any = mk[1];
    if (threadIdx.x == 0) { s_mask[(chunk + 2) % 3][0] = 0; s_mask[(chunk + 2) % 3][1] = 0; NN_STAT(0, 1); }
    if (!(open_any >> 31)) { if (up) open_up = false; else open_dn = false; }
    while (want_any) {
      const int bit = __ffs(want_any) - 1;
      want_any &= want_any - 1;
      const int t = up ? first + bit : first - bit;
      const bool want = live && !(__shfl_sync(FULL_MASK, lb, bit) > T) && ((wm >> bit) & 1u);
      // T has moved since the chunk was voted on: is the tile still needed?  (This barrier also separates the reads of the
      // previous tile from the loads of this one.)
      if (!__syncthreads_or(want ? 1 : 0)) continue;
      const int base = t * NEAREST_TILE;
      const int n = min(NEAREST_TILE, so.n_nodes - base);
      if (threadIdx.x == 0) NN_STAT(1, 1);
      for (int q = threadIdx.x; q < n; q += NN_SORTED_WARPS * 32) {
        const int g = base + q;
        s_fx[q] = so.fx[g]; s_fy[q] = so.fy[g]; s_ce[q] = so.ce[g];
        s_frx[q] = so.frx[g]; s_fry[q] = so.fry[g];
        s_fdx[q] = so.fdx[g]; s_fdy[q] = so.fdy[g];
      }
      __syncthreads();
      if (want) {
        // stage 1: distance bound against T, then the float pre-test of feasibility for the nodes that pass it
        // The nodes of a tile are stored in axis order.  Explore key: they are taken from the sample's own axis position
        // outwards (tiles before the sample: last node first; the sample's own tile: the part before the sample backwards, then
        // the rest forwards).  Optimise key: in storage order, root side first.  The list does not depend on the order — it is
        // ordered by (key, node id) — but its cost does: best first, T is tight after a few insertions and stage 1 drops almost
        // everything that follows; worst first, every feasible node of the tile is inserted in turn (C3, 4096 nodes, per sample:
        // 61 -> 13 insertions with the explore key, 104 -> 10 with the optimise key).
        int split = 0;
        if (!optimize) {
          if (so.tile_uhi[t] < su) split = n;
          else if (!(so.tile_ulo[t] > su)) {
            for (int i0 = 0; i0 < n; i0 += 32) {
              const int i = i0 + lane;
              split += __popc(__ballot_sync(FULL_MASK, i < n && s_fx[i] * so.cb + s_fy[i] * so.sb < su));
            }
          }
        }
        int c1 = 0;
        for (int i0 = 0; i0 < n; i0 += 32) {
          const int ii = i0 + lane;
          const int i = ii < split ? split - 1 - ii : ii;
          bool keep = false;
          if (ii < n) keep = stage1_keep(fsx - s_fx[i], fsy - s_fy[i], optimize ? s_ce[i] : 0.0f, T, slack);
          if (keep) keep = feasible_maybe(fsx, fsy, s_frx[i], s_fry[i], s_fdx[i], s_fdy[i], feas_len2);
          const unsigned m = __ballot_sync(FULL_MASK, keep);
          if (keep) s_idx[warp][c1 + __popc(m & lt)] = (uint16_t)i;
          c1 += __popc(m);
        }
        __syncwarp();
        if (lane == 0) { NN_STAT(2, 1); NN_STAT(3, c1); }
        // stage 2: feasibility of the survivors, in the reference's double arithmetic
        int c2 = 0;
        for (int q0 = 0; q0 < c1; q0 += 32) {
          const int q = q0 + lane;
          int i = 0;
          bool f = false;
          if (q < c1) {
            i = s_idx[warp][q];
            const int g = base + i;
            f = feasible_node(sx, sy, so.rbx[g], so.rby[g], so.dpx[g], so.dpy[g], so.ang[g], a.feas_len);
          }
          const unsigned m = __ballot_sync(FULL_MASK, f);
          if (f) s_idx2[warp][c2 + __popc(m & lt)] = (uint16_t)i;
          c2 += __popc(m);
        }
        __syncwarp();
        if (lane == 0) NN_STAT(4, c2);
        // stage 3: Dubins keys of the feasible survivors and insertion into the warp's list
        for (int q0 = 0; q0 < c2; q0 += 32) {
          const int q = q0 + lane;
          float key = INFINITY;
          int idx = INT_MAX;
          if (q < c2) {
            const int i = s_idx2[warp][q];
            const int g = base + i;
            key = dubins_key(sx, sy, so.nx[g], so.ny[g], so.ca[g], so.sa[g]);
            if (optimize) key = s_ce[i] + key;  // rrtplanner.cpp:254
            idx = so.node_id[g];
          }
          unsigned wantm = __ballot_sync(FULL_MASK, key < T || (key == T && idx < Tid));
          while (wantm) {
            const int src = __ffs(wantm) - 1;
            wantm &= wantm - 1;
            const float nk = __shfl_sync(FULL_MASK, key, src);
            const int nid = __shfl_sync(FULL_MASK, idx, src);
            if (!(nk < T || (nk == T && nid < Tid))) continue;  // T moved since the ballot
            if (lane == 0) NN_STAT(5, 1);
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
  }
  if (!live) return;
  if (lane == 0) NN_STAT(optimize ? 7 : 6, 1);
  const int cnt = __popc(__ballot_sync(FULL_MASK, lane < CLRRT_SORT_LIMIT && lid != INT_MAX));
  if (lane < CLRRT_SORT_LIMIT) {
    const bool valid = lid != INT_MAX;
    a.cand[(size_t)j * CLRRT_SORT_LIMIT + lane] = valid ? lid : -1;
    if (a.key) a.key[(size_t)j * CLRRT_SORT_LIMIT + lane] = valid ? lk : 0.0f;
  }
  if (lane == 0) a.count[j] = cnt;
}

// Small searches (below NN_SORT_MIN_PAIRS sample-node pairs): all tiles in storage order, no sort (the sort's five
// launches and the per-tile block votes cost more than they save when there are only a few tiles).
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
  __shared__ float s_fx[NEAREST_TILE], s_fy[NEAREST_TILE];  // node positions rounded to float: stage 1 only
  __shared__ uint16_t s_idx[NEAREST_WARPS][NEAREST_TILE];
  __shared__ uint16_t s_idx2[NEAREST_WARPS][NEAREST_TILE];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const unsigned lt = (1u << lane) - 1u;
  const int j = blockIdx.x * NEAREST_WARPS + warp;
  const bool live = j < a.K;
  double sx = 0, sy = 0;
  bool optimize = false;
  if (live) { sx = a.sample_xy[2 * j]; sy = a.sample_xy[2 * j + 1]; optimize = a.heuristic[j] != 0; }
  const float fsx = (float)sx, fsy = (float)sy;
  const float slack = 1.0e-3f + 4.0e-7f * (fabsf(fsx) + fabsf(fsy));
  const float feas_len2 = (float)(a.feas_len * a.feas_len);
  float lk = INFINITY;  // entry `lane` of the list (lanes >= 10 stay at +inf / INT_MAX and never take part)
  int lid = INT_MAX;
  float T = INFINITY;   // key and id of the 10th entry (warp-uniform)
  int Tid = INT_MAX;

  for (int base = 0; base < a.n_nodes; base += NEAREST_TILE) {
    const int n = min(NEAREST_TILE, a.n_nodes - base);
    __syncthreads();
    if ((int)threadIdx.x < n) {
      const int g = base + threadIdx.x;
      const double nxg = a.tree.x[g], nyg = a.tree.y[g];
      s_nx[threadIdx.x] = nxg; s_ny[threadIdx.x] = nyg; s_fx[threadIdx.x] = (float)nxg; s_fy[threadIdx.x] = (float)nyg;
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
        if (i < n) keep = stage1_keep(fsx - s_fx[i], fsy - s_fy[i], optimize ? s_ce[i] : 0.0f, T, slack);
        if (keep) keep = feasible_maybe(fsx, fsy, (float)s_rbx[i], (float)s_rby[i], (float)s_dpx[i], (float)s_dpy[i], feas_len2);
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

// ---- K = 1, the reference's own sequential algorithm: the order of EQUAL keys ----------------------------------------
// The reference sorts (node id, key) pairs with std::sort on the key alone (rrtplanner.cpp:233, :256): the order of equal
// keys is whatever libstdc++'s unstable introsort leaves, and it decides which of two equally ranked parents is tried
// first (seen in the receding-horizon loop, where carried-over nodes produce exactly equal keys).  The kernels above order
// equal keys by node id.  For a single sample this kernel evaluates key and feasibility of EVERY node and raises a flag
// when a feasible node outside the list position it would take shares its key with a list entry (or a feasible key is
// NaN); the host then repeats the reference's very std::sort call on the downloaded keys (clrrt_api.cu, nearest_dev).
struct TieArgs {
  NodeSoA tree;
  int32_t n_nodes;
  const double* sample_xy;
  const uint8_t* heuristic;
  double feas_len;
  const int32_t* cand;   // [10] the list found with the node-id tie rule
  const float* key;      // [10]
  const int32_t* count;  // [1]
  float* all_key;        // [n_nodes]
  uint8_t* all_feas;     // [n_nodes]
  int32_t* flag;
};
__global__ void __launch_bounds__(128) tie_check_kernel(const TieArgs a) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.n_nodes) return;
  const double sx = a.sample_xy[0], sy = a.sample_xy[1];
  const bool optimize = a.heuristic[0] != 0;
  float key = dubins_key(sx, sy, a.tree.x[i], a.tree.y[i], a.tree.ca[i], a.tree.sa[i]);
  if (optimize) key = a.tree.costE[i] + key;  // rrtplanner.cpp:254
  const double rbx = a.tree.rbx[i], rby = a.tree.rby[i];
  const bool feas = feasible_node(sx, sy, rbx, rby, rbx - a.tree.rfx[i], rby - a.tree.rfy[i], a.tree.angPar[i], a.feas_len);
  a.all_key[i] = key;
  a.all_feas[i] = feas ? 1 : 0;
  if (!feas) return;
  bool tie = !(key == key);
  const int cnt = a.count[0];
  for (int r = 0; r < cnt; r++)
    if (key == a.key[r] && a.cand[r] != i) tie = true;
  if (tie) atomicOr(a.flag, 1);
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
