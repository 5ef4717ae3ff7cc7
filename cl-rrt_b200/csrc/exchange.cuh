// exchange.cuh — the per-round node exchange of the multi-GPU path, inside the library (sm_100a + NCCL over NVLink).
//
// The reference is one process (expandTree, rrt/include/rrt/rrtplanner.h:87); its append order is the order of its samples
// (rrt/src/rrtplanner.cpp:150-173).  With one context per GPU every rank expands its contiguous shard of a round's
// samples against the same replicated tree, and the accepted nodes of all ranks — fixed-stride records, csrc/tree.cuh —
// are all-gathered and appended in rank order, which is global sample order: the tree is identical for any world size.
//
// NCCL is bound at run time (dlopen of libnccl.so.2, the copy the process already has when a framework loaded one), so
// the single-GPU library has no NCCL dependency.  Per round: ncclAllGather of the ranks' record counts (4 bytes each),
// one device-to-host copy of those counts together with the round's counters (the only host synchronisation, the one a
// single-GPU round needs too), ncclAllGather of the records at the stride of the largest count, ONE append launch over
// all ranks' chunks with the chunk offsets taken from the gathered counts on the device.
#pragma once
#include <dlfcn.h>
#include <nccl.h>

#include <string>

#include "tree.cuh"

struct NcclApi {
  void* lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

inline const NcclApi* nccl_api(std::string* err) {
  static NcclApi api;
  static bool tried = false;
  if (!tried) {
    tried = true;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
      api.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
      if (api.lib) break;
    }
    if (api.lib) {
      api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(dlsym(api.lib, "ncclGetUniqueId"));
      api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(dlsym(api.lib, "ncclCommInitRank"));
      api.AllGather = reinterpret_cast<decltype(api.AllGather)>(dlsym(api.lib, "ncclAllGather"));
      api.AllReduce = reinterpret_cast<decltype(api.AllReduce)>(dlsym(api.lib, "ncclAllReduce"));
      api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(dlsym(api.lib, "ncclCommDestroy"));
      api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(dlsym(api.lib, "ncclGetErrorString"));
      if (!api.GetUniqueId || !api.CommInitRank || !api.AllGather || !api.AllReduce || !api.CommDestroy || !api.GetErrorString) {
        dlclose(api.lib);
        api.lib = nullptr;
      }
    }
  }
  if (!api.lib) {
    if (err) *err = "NCCL not available: dlopen(libnccl.so.2) failed";
    return nullptr;
  }
  return &api;
}

// Gathered records -> tree SoA.  Rank r's chunk starts at record r * stride of `gathered` and holds counts[r] records;
// output position i of the concatenation goes to tree slot first + i.  At most `limit` records are appended (a full tree
// takes the prefix that fits).  One launch for all ranks; the offsets come from the device copy of the counts.
#define CLRRT_MAX_WORLD 64
__global__ void __launch_bounds__(256)
append_gathered_kernel(NodeSoA t, int first, const NodeRecord* __restrict__ gathered, int stride, const int32_t* __restrict__ counts,
                       int world, int limit, int capacity) {
  __shared__ int s_off[CLRRT_MAX_WORLD + 1];
  if (threadIdx.x == 0) {
    int acc = 0;
    for (int r = 0; r < world; r++) { s_off[r] = acc; acc += counts[r]; }
    s_off[world] = acc;
  }
  __syncthreads();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= s_off[world] || i >= limit) return;
  int r = 0;
  while (i >= s_off[r + 1]) r++;
  const int k = first + i;
  if (k >= capacity) return;
  const NodeRecord rec = gathered[(size_t)r * stride + (i - s_off[r])];
  t.x[k] = rec.state[0]; t.y[k] = rec.state[1]; t.th[k] = rec.state[2]; t.de[k] = rec.state[3]; t.v[k] = rec.state[4];
  t.a[k] = rec.state[5]; t.t[k] = rec.state[6]; t.s7[k] = rec.state[7]; t.s8[k] = rec.state[8]; t.s9[k] = rec.state[9];
  t.rfx[k] = rec.rf[0]; t.rfy[k] = rec.rf[1]; t.rbx[k] = rec.rb[0]; t.rby[k] = rec.rb[1]; t.vback[k] = rec.vback;
  t.costE[k] = rec.costE; t.costS[k] = rec.costS;
  t.parent[k] = rec.parent == -2 ? k - 1 : rec.parent;  // a goal-biased child follows its parent inside one rank's chunk
  t.goal[k] = rec.goal; t.nref[k] = rec.nref; t.kind[k] = rec.kind; t.smx[k] = rec.smp[0]; t.smy[k] = rec.smp[1];
  const float ang = (float)(-rec.state[2] - M_PI * 0.0);  // rrtplanner.cpp:378
  float s, c;
  ref_sincosf(ang, &s, &c);
  t.ca[k] = c; t.sa[k] = s;
  t.angPar[k] = atan2(rec.rb[1] - rec.rf[1], rec.rb[0] - rec.rf[0]);  // rrtplanner.cpp:273
}

// Order-sensitive 128-bit digest of tree nodes [first, first + n): every node's 160-byte record (the clrrt_node layout,
// parents as tree indices) is hashed together with its index; the per-node hashes are combined by sum and by xor.
// Equal trees give equal digests on any rank and for any world size; used to prove that in the bench and the tests.
__device__ __forceinline__ unsigned long long mix64(unsigned long long z) {
  z ^= z >> 30; z *= 0xbf58476d1ce4e5b9ull; z ^= z >> 27; z *= 0x94d049bb133111ebull; z ^= z >> 31;
  return z;
}
__global__ void __launch_bounds__(256) tree_digest_kernel(NodeSoA t, int first, int n, unsigned long long* __restrict__ out2) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  unsigned long long h = 0;
  if (i < n) {
    const int g = first + i;
    h = mix64(0x9e3779b97f4a7c15ull + (unsigned long long)g);
    const double d[17] = {t.x[g], t.y[g], t.th[g], t.de[g], t.v[g], t.a[g], t.t[g], t.s7[g], t.s8[g], t.s9[g],
                          t.rfx[g], t.rfy[g], t.rbx[g], t.rby[g], t.vback[g], t.smx[g], t.smy[g]};
#pragma unroll
    for (int k = 0; k < 17; k++) h = mix64(h ^ (unsigned long long)__double_as_longlong(d[k]));
    h = mix64(h ^ (((unsigned long long)__float_as_uint(t.costE[g]) << 32) | __float_as_uint(t.costS[g])));
    h = mix64(h ^ (((unsigned long long)(unsigned)t.parent[g] << 32) | (unsigned)t.goal[g]));
    h = mix64(h ^ (((unsigned long long)(unsigned)t.nref[g] << 32) | (unsigned)t.kind[g]));
  }
  unsigned long long sum = h, x = h;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    sum += __shfl_down_sync(FULL_MASK, sum, o);
    x ^= __shfl_down_sync(FULL_MASK, x, o);
  }
  if ((threadIdx.x & 31) == 0 && (sum | x)) {
    atomicAdd(&out2[0], sum);
    atomicXor(&out2[1], x);
  }
}
