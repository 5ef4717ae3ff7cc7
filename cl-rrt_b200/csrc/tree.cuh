// tree.cuh — appending accepted rollouts to the tree in deterministic (sample) order (sm_100a).
//
// Replaces MyRRT::addNode (rrt/include/rrt/rrtplanner.h:111-113) for a whole round: the rollout kernels leave
// accepted nodes in a staging SoA (slot j = node from sample j, slot K+j = its goal-biased child), this file
// (slot[j] = winning candidate of sample j among its n_ranks staging slots, gb_base + j = its goal-biased child)
// compacts them in the order the sequential reference would have appended them (sample 0's node, its goal
// child, sample 1's node, ...) into fixed-stride records, and appends records to the tree SoA.  Records are
// also the unit of the multi-GPU exchange: every rank all-gathers its records and appends all ranks' chunks
// in rank order, which is global sample order, so the tree is identical for any world size.
#pragma once
#include "common.cuh"
#include "refmath.cuh"

struct __align__(16) NodeRecord {  // CLRRT_RECORD_BYTES
  double state[10];
  double rf[2], rb[2];
  double vback;
  float costE, costS;
  int32_t parent;   // tree index, or -2: "the record just before this one" (goal-biased child of the node before)
  int32_t goal, nref, kind;
  double smp[2];    // sample the reference was aimed at
};
static_assert(sizeof(NodeRecord) == CLRRT_RECORD_BYTES, "record stride");

#define SCAN_THREADS 1024

// flags are read in append order: entry 2j = valid[j] (main node of sample j), 2j+1 = valid[K+j] (goal child)
__global__ void __launch_bounds__(SCAN_THREADS) scan_block_sums_kernel(const int32_t* valid, int K, int32_t* block_sums) {
  __shared__ int s_w[32];
  const int j = blockIdx.x * SCAN_THREADS + threadIdx.x;
  int c = 0;
  if (j < K) c = (valid[j] != 0) + (valid[K + j] != 0);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) c += __shfl_down_sync(FULL_MASK, c, o);
  if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x < 32) {
    int v = s_w[threadIdx.x];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(FULL_MASK, v, o);
    if (threadIdx.x == 0) block_sums[blockIdx.x] = v;
  }
}

// exclusive scan of up to SCAN_THREADS block sums, in place; total to *total
__global__ void __launch_bounds__(SCAN_THREADS) scan_sums_kernel(int32_t* block_sums, int nblocks, int32_t* total) {
  __shared__ int s[SCAN_THREADS];
  const int t = threadIdx.x;
  int carry = 0;
  for (int base = 0; base < nblocks; base += SCAN_THREADS) {
    const int i = base + t;
    const int v = i < nblocks ? block_sums[i] : 0;
    s[t] = v;
    __syncthreads();
    for (int o = 1; o < SCAN_THREADS; o <<= 1) {
      const int add = t >= o ? s[t - o] : 0;
      __syncthreads();
      s[t] += add;
      __syncthreads();
    }
    if (i < nblocks) block_sums[i] = carry + s[t] - v;
    const int chunk_total = s[SCAN_THREADS - 1];
    __syncthreads();
    carry += chunk_total;
  }
  if (t == 0) *total = carry;
}

__device__ __forceinline__ void fill_record(NodeRecord& r, const NodeSoA& S, int k, int parent) {
  r.state[0] = S.x[k]; r.state[1] = S.y[k]; r.state[2] = S.th[k]; r.state[3] = S.de[k]; r.state[4] = S.v[k];
  r.state[5] = S.a[k]; r.state[6] = S.t[k]; r.state[7] = S.s7[k]; r.state[8] = S.s8[k]; r.state[9] = S.s9[k];
  r.rf[0] = S.rfx[k]; r.rf[1] = S.rfy[k]; r.rb[0] = S.rbx[k]; r.rb[1] = S.rby[k]; r.vback = S.vback[k];
  r.costE = S.costE[k]; r.costS = S.costS[k]; r.parent = parent; r.goal = S.goal[k]; r.nref = S.nref[k];
  r.kind = S.kind[k]; r.smp[0] = S.smx[k]; r.smp[1] = S.smy[k];
}

__global__ void __launch_bounds__(SCAN_THREADS)
pack_records_kernel(NodeSoA stage, const int32_t* valid, int K, const int32_t* block_offsets, NodeRecord* records,
                    const int32_t* slot, int gb_base) {
  __shared__ int s_w[32];
  const int j = blockIdx.x * SCAN_THREADS + threadIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const bool vm = j < K && valid[j] != 0;
  const bool vg = j < K && valid[K + j] != 0;
  const int c = (int)vm + (int)vg;
  int incl = c;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int n = __shfl_up_sync(FULL_MASK, incl, o);
    if (lane >= o) incl += n;
  }
  if (lane == 31) s_w[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    int v = s_w[lane];
    int iv = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int n = __shfl_up_sync(FULL_MASK, iv, o);
      if (lane >= o) iv += n;
    }
    s_w[lane] = iv - v;
  }
  __syncthreads();
  int pos = block_offsets[blockIdx.x] + s_w[warp] + incl - c;
  if (vm) { const int sidx = slot[j]; fill_record(records[pos], stage, sidx, stage.parent[sidx]); pos++; }
  if (vg) fill_record(records[pos], stage, gb_base + j, -2);
}

// records -> tree SoA at [first, first+n), plus the derived per-node quantities of the nearest-node search
__global__ void append_records_kernel(NodeSoA t, int first, const NodeRecord* __restrict__ records, int n, int capacity) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int k = first + i;
  if (k >= capacity) return;
  const NodeRecord r = records[i];
  t.x[k] = r.state[0]; t.y[k] = r.state[1]; t.th[k] = r.state[2]; t.de[k] = r.state[3]; t.v[k] = r.state[4];
  t.a[k] = r.state[5]; t.t[k] = r.state[6]; t.s7[k] = r.state[7]; t.s8[k] = r.state[8]; t.s9[k] = r.state[9];
  t.rfx[k] = r.rf[0]; t.rfy[k] = r.rf[1]; t.rbx[k] = r.rb[0]; t.rby[k] = r.rb[1]; t.vback[k] = r.vback;
  t.costE[k] = r.costE; t.costS[k] = r.costS;
  t.parent[k] = r.parent == -2 ? k - 1 : r.parent;
  t.goal[k] = r.goal; t.nref[k] = r.nref; t.kind[k] = r.kind; t.smx[k] = r.smp[0]; t.smy[k] = r.smp[1];
  const float ang = (float)(-r.state[2] - M_PI * 0.0);  // rrtplanner.cpp:378
  float s, c;
  ref_sincosf(ang, &s, &c);
  t.ca[k] = c; t.sa[k] = s;
  t.angPar[k] = atan2(r.rb[1] - r.rf[1], r.rb[0] - r.rf[0]);  // rrtplanner.cpp:273
}

// extractBestPath, rrtplanner.cpp:318-368: arg-min of float costS over goal-flagged nodes (first minimum wins)
// tree nodes [first, first + n) -> clrrt_node records (same 160-byte layout as NodeRecord, parents as tree indices):
// one coalesced device-to-host copy instead of one per field
__global__ void __launch_bounds__(256) export_nodes_kernel(NodeSoA t, int first, int n, NodeRecord* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int g = first + i;
  NodeRecord r;
  r.state[0] = t.x[g]; r.state[1] = t.y[g]; r.state[2] = t.th[g]; r.state[3] = t.de[g]; r.state[4] = t.v[g];
  r.state[5] = t.a[g]; r.state[6] = t.t[g]; r.state[7] = t.s7[g]; r.state[8] = t.s8[g]; r.state[9] = t.s9[g];
  r.rf[0] = t.rfx[g]; r.rf[1] = t.rfy[g]; r.rb[0] = t.rbx[g]; r.rb[1] = t.rby[g];
  r.vback = t.vback[g]; r.costE = t.costE[g]; r.costS = t.costS[g];
  r.parent = t.parent[g]; r.goal = t.goal[g]; r.nref = t.nref[g]; r.kind = t.kind[g];
  r.smp[0] = t.smx[g]; r.smp[1] = t.smy[g];
  out[i] = r;
}

__global__ void __launch_bounds__(1024) best_goal_kernel(NodeSoA t, int n, int32_t* best_id) {
  __shared__ float s_c[32];
  __shared__ int s_i[32];
  float bc = INFINITY;
  int bi = INT_MAX;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    if (t.goal[i]) {
      const float c = t.costS[i];
      if (bi == INT_MAX || c < bc) { bc = c; bi = i; }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float oc = __shfl_xor_sync(FULL_MASK, bc, o);
    const int oi = __shfl_xor_sync(FULL_MASK, bi, o);
    if (oi != INT_MAX && (bi == INT_MAX || oc < bc || (oc == bc && oi < bi))) { bc = oc; bi = oi; }
  }
  if ((threadIdx.x & 31) == 0) { s_c[threadIdx.x >> 5] = bc; s_i[threadIdx.x >> 5] = bi; }
  __syncthreads();
  if (threadIdx.x < 32) {
    bc = s_c[threadIdx.x]; bi = s_i[threadIdx.x];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float oc = __shfl_xor_sync(FULL_MASK, bc, o);
      const int oi = __shfl_xor_sync(FULL_MASK, bi, o);
      if (oi != INT_MAX && (bi == INT_MAX || oc < bc || (oc == bc && oi < bi))) { bc = oc; bi = oi; }
    }
    if (threadIdx.x == 0) *best_id = bi == INT_MAX ? -1 : bi;
  }
}
