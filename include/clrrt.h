/* clrrt.h — C ABI of the B200-native CL-RRT tree-expansion path (libclrrt_b200.so).
 *
 * This is the drop-in boundary for ONE hot path of vdBerg93/cl-rrt:
 *   MotionPlanner::planMotion -> expandTree -> Simulation (ctor + propagate)
 * with nearest-node selection (sortNodesExplore / sortNodesOptimize) and the per-step OBB/SAT
 * collision check.  The reference has no FFI of its own (it is one C++ translation unit); each entry
 * point below names the reference interface it replaces (paths relative to the reference checkout).
 * INTEGRATION.md shows the few lines a maintainer adds to rrt/src to call these instead.
 *
 * Conventions: plain pointers and sizes only; every call returns 0 or a negative clrrt_status; nothing
 * is thrown across the boundary; the caller owns all host buffers, the context owns all device memory
 * and its stream; calls on one context are not thread-safe (the reference is single-threaded and keeps
 * its state in file-scope globals).  Distinct contexts may be used alternately from one thread; calls from
 * several threads of one process must be serialised by the caller even across contexts of the same GPU (the kernels'
 * parameter block is one __constant__ symbol per device).  One context per GPU and process is the intended use.
 * There is NO CPU fallback: without a CUDA device clrrt_create fails with CLRRT_ERR_CUDA.
 */
#ifndef CLRRT_H
#define CLRRT_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CLRRT_SORT_LIMIT 10   /* MyRRT::sortLimit, rrt/src/rrtplanner.cpp:13 */
#define CLRRT_MAX_STEPS_CAP 1024

typedef enum clrrt_status {
  CLRRT_OK = 0,
  CLRRT_ERR_ARG = -1,      /* null pointer / negative size / index out of range */
  CLRRT_ERR_CUDA = -2,     /* CUDA runtime error; text in clrrt_last_error() */
  CLRRT_ERR_CAPACITY = -3, /* tree or scratch capacity exceeded */
  CLRRT_ERR_STATE = -4     /* call order (e.g. expand before tree_reset) */
} clrrt_status;

/* Vehicle, rrt/include/rrt/vehicle.h:5-21 (same field order). */
typedef struct clrrt_vehicle {
  double dmax, ddmax, Td, Ta, amin, amax, L, w, Lrear, Lfront, b, Vch, rho, Kus;
} clrrt_vehicle;

/* Everything the reference keeps in file-scope globals (rrt/src/rrt_node.cpp:2-24), ROS parameters
 * (rrt/launch/parameters.launch:3-20) and MyRRT members (rrt/include/rrt/rrtplanner.h:51-67). */
typedef struct clrrt_params {
  clrrt_vehicle veh;
  double sim_dt;        /* ctrl/sampleTime */
  double ctrl_tla, ctrl_mindla, ctrl_dlavmin, ctrl_Kp, ctrl_Ki;
  double ref_int, ref_mindist;
  double ref_res;       /* updateReferenceResolution(v), rrt/src/controller.cpp:18-21; set per query */
  double vmax;          /* MotionRequest.vmax, rrt/src/motionplanner.cpp:17 */
  double ay_road_max;   /* rrt/src/rrt_node.cpp:18 (never assigned upstream: 0) */
  double Wcost[5];      /* MyRRT::Wcost, rrt/src/rrtplanner.cpp:14-18 */
  double goal[4];       /* MyRRT::goalPose = MotionRequest.goal: x, y, heading, v */
  int32_t obs_use_pred; /* rrt/src/rrt_node.cpp:11 */
  int32_t fp32;         /* 0: reference arithmetic (double rollout, float geometry); 1: float rollout */
  /* curved-road mode (MotionRequest.bend, rrt/src/motionplanner.cpp:23): every sim step adds Wcost[4] x the distance to
   * the goal lane, getDistToLane(x, y, laneShifts[0], Cxy) (rrt/src/simulation.cpp:49-53, :92-95) */
  double Cxy[3];        /* MotionRequest.Cxy: road centre line y = Cxy[0] x^2 + Cxy[1] x + Cxy[2] (only [1], [2] are read) */
  double lane_shift;    /* MotionRequest.laneShifts[0] */
  int32_t bend;
  int32_t reserved;
} clrrt_params;

/* car_msgs/msg/Obstacle2D.msg: vision_msgs/BoundingBox2D obb {center{x,y,theta}, size_x, size_y} + vel.linear */
typedef struct clrrt_obstacle {
  double cx, cy, theta, size_x, size_y, vx, vy;
} clrrt_obstacle;

/* One tree node: the fields of struct Node (rrt/include/rrt/rrtplanner.h:35-47) the hot path reads.
 * ref_front/ref_back are ref.{x,y}.front()/back(), ref_vback is ref.v.back().  Trajectories (Node::tra)
 * are not kept on the device; clrrt_propagate_batch re-materialises them on request. */
typedef struct clrrt_node {
  double state[10];   /* x y theta delta v a t IDwp vref dcmd, rrt/src/motionplanner.cpp:90, simulation.cpp:64-67 */
  double ref_front[2];
  double ref_back[2];
  double ref_vback;
  float costE, costS; /* stored as float upstream, rrtplanner.h:40-41 */
  int32_t parent;     /* parentID, -1 for the root */
  int32_t goal_reached;
  int32_t n_ref;      /* ref.x.size() */
  int32_t kind;       /* 0 root / uploaded, 1 expansion towards `sample` (rrtplanner.cpp:151-158), 2 goal-biased (:165-171) */
  double sample[2];   /* the sample the node's reference was aimed at: with parent and kind it re-creates the rollout */
} clrrt_node;

/* Result of one Simulation (rrt/include/rrt/simulation.h:11-17) plus what expandTree derives from it. */
typedef struct clrrt_rollout {
  double state[10];   /* stateArray.back() */
  double costE, costS;
  double ref_back[2]; /* ref.x.back(), ref.y.back() of the generated reference */
  double ref_vback;   /* ref.v.back() after generateVelocityProfile */
  double trace;       /* sum over steps i=1.. of i*IDwp_i (checksum of the waypoint trace) */
  int32_t end_reached, goal_reached;
  int32_t n_steps;    /* iterations of the loop at rrt/src/simulation.cpp:58 */
  int32_t fail;       /* 0 none, 1 collision (:85), 2 lateral acceleration (:102), 3 iteration limit (:142) */
  int32_t n_ref;      /* reference points */
  int32_t idwp0;      /* waypoint chosen by the Controller ctor (stateArray[0][7]) */
  int32_t tainted;    /* some step used IDwp >= n_ref-2: the unmodified reference reads out of bounds there */
  int32_t reserved;
} clrrt_rollout;

/* Failure counters of rrt/src/rrt_node.cpp:21-24, accumulated since the last clrrt_tree_reset. */
typedef struct clrrt_counters {
  int64_t fail_collision, fail_acclimit, fail_iterlimit, sim_count;
  int64_t rollouts;      /* Simulation constructions */
} clrrt_counters;

typedef struct clrrt_round_stats {
  int32_t samples, rollouts, nodes_added, goal_nodes_added, tree_size;
  int32_t reserved;
  int64_t sim_steps;
  /* device time per phase (CUDA events on the ctx stream): candidate search; the rollout kernel alone; reference ends +
   * launch order before it; winner selection + compaction + append after it */
  float ms_nearest, ms_rollout, ms_prepare, ms_append;
  /* multi-GPU (clrrt_comm_*): the exchange's share of ms_append (count + record all-gather and the append of all ranks'
   * chunks); nodes_added then counts the nodes of ALL ranks, nodes_local this rank's, samples / rollouts / sim_steps this
   * rank's shard */
  float ms_exchange;
  int32_t nodes_local;
} clrrt_round_stats;

typedef struct clrrt_ctx clrrt_ctx;

/* Launch-file parameters + Vehicle::setPrius() (rrt/src/motionplanner.cpp:13, rrt/include/rrt/vehicle.h:39-60),
 * goal (50,0,0,0), vmax 5. */
int clrrt_default_params(clrrt_params* p);

/* Context: device buffers for a tree of up to tree_capacity nodes and rounds of up to max_round samples.
 * `stream` is a cudaStream_t to launch on (NULL: the context creates its own). */
int clrrt_create(const clrrt_params* p, int device, int tree_capacity, int max_round, void* stream, clrrt_ctx** out);
int clrrt_destroy(clrrt_ctx* ctx);
const char* clrrt_last_error(const clrrt_ctx* ctx);

/* Start-of-query updates: vmax/goal/ref_res/lookahead of rrt/src/motionplanner.cpp:16-17, weights of
 * rrt/src/rrtplanner.cpp:14-18. */
int clrrt_set_params(clrrt_ctx* ctx, const clrrt_params* p);

/* == MotionPlanner::updateObstacles (rrt/src/motionplanner.cpp:81-86): replaces the obstacle set `det` that
 * checkObsDistance (rrt/src/old_collisioncheck.cpp:24-51) tests after every sim step.  n == 0 gives the shipped
 * stub behaviour (rrt/src/collisioncheck.cpp:6-8: distance 100, never a collision). */
int clrrt_set_obstacles(clrrt_ctx* ctx, const clrrt_obstacle* host, int n);

/* == the result of initializeTree (rrt/src/rrtplanner.cpp:39-95): upload the initial tree (root node, or the
 * carried-over chain) and zero the counters. */
int clrrt_tree_reset(clrrt_ctx* ctx, const clrrt_node* host, int n);
int clrrt_tree_size(const clrrt_ctx* ctx);
int clrrt_tree_truncate(clrrt_ctx* ctx, int n); /* drop nodes >= n (bench: same snapshot every round) */
int clrrt_tree_download(clrrt_ctx* ctx, clrrt_node* host, int cap, int* n);
int clrrt_tree_download_range(clrrt_ctx* ctx, int first, int count, clrrt_node* host);
/* Same without waiting: the nodes are staged on the context's stream (the range may be truncated and re-grown by the very
 * next call) and copied to `host` (pinned memory, valid until clrrt_download_wait) on a second stream, beside the next
 * round.  One download in flight per context. */
int clrrt_tree_download_range_async(clrrt_ctx* ctx, int first, int count, clrrt_node* host);
int clrrt_download_wait(clrrt_ctx* ctx);

/* == sortNodesExplore (heuristic[j]==0) / sortNodesOptimize (==1), rrt/src/rrtplanner.cpp:227-268, for K samples
 * against the current tree: up to CLRRT_SORT_LIMIT feasible node ids in increasing key order (ties: lower id),
 * padded with -1; key = float Dubins distance (+ float costE); count = list length. */
int clrrt_nearest_batch(clrrt_ctx* ctx, const double* sample_xy, const uint8_t* heuristic, int K,
                        int32_t* cand, float* key, int32_t* count);

/* == M times { getReference / getGoalReference + Simulation::Simulation } (rrt/src/rrtplanner.cpp:151-152 when
 * goal_biased[j]==0, :165-166 when 1; rrt/src/simulation.cpp:36-143) from tree node parent[j] towards sample j.
 * traj (optional, may be NULL) receives stateArray: M x traj_stride x 10 doubles, rows 0..n_steps. */
int clrrt_propagate_batch(clrrt_ctx* ctx, const int32_t* parent, const double* sample_xy,
                          const uint8_t* goal_biased, int M, clrrt_rollout* out, double* traj, int traj_stride);
/* Same, and additionally the generated references themselves (MyReference::x, y, v after generateVelocityProfile):
 * ref_xyv (optional) receives M x ref_stride x 3 doubles, rows 0..n_ref-1.  This is how Node::tra and Node::ref of
 * an extracted best path are re-materialised (the device tree keeps only end points). */
int clrrt_propagate_batch_ex(clrrt_ctx* ctx, const int32_t* parent, const double* sample_xy,
                             const uint8_t* goal_biased, int M, clrrt_rollout* out, double* traj, int traj_stride,
                             double* ref_xyv, int ref_stride);

/* == Simulation::Simulation(const MyRRT& RRT, const vector<double>& state, MyReference& ref, const Vehicle& veh,
 *                           const bool& GoalBiased, const bool& genProfile, const double& Vstart)
 * (rrt/include/rrt/simulation.h:18-19, rrt/src/simulation.cpp:36-143) with the reference's own parameter list: an ARBITRARY
 * start state (10 entries, rrt/src/motionplanner.cpp:90 + the four logging slots) and a caller-owned reference path
 * ref.x / ref.y of n_ref >= 3 points of any shape, ref_dir = MyReference::dir (+1 / -1).  gen_profile != 0 runs
 * generateVelocityProfile and FILLS ref_v[0..n_ref) — the constructor mutates the caller's MyReference the same way
 * (expandTree then stores that ref in the new Node, rrt/src/rrtplanner.cpp:156); gen_profile == 0 uses ref_v as given.
 * RRT (goal, obstacles, weights) and veh are the context's.  out->state = stateArray.back(), out->fail / end_reached /
 * goal_reached as clrrt_propagate_batch; traj (optional) receives stateArray, traj_stride x 10.  The failure counters
 * advance as upstream (rrt/src/simulation.cpp:59, :85, :102, :142).  Out-of-bounds reads of the unmodified reference
 * (ref.v[IDwp+2], ref.x[IDwp+1] near the end of the path) follow the index-clamped "defined" variant (out->tainted). */
int clrrt_simulate(clrrt_ctx* ctx, const double* state10, const double* ref_x, const double* ref_y, double* ref_v, int n_ref,
                   int ref_dir, int goal_biased, int gen_profile, double Vstart, clrrt_rollout* out, double* traj, int traj_stride);
/* M of them in one launch: ragged references, rollout i uses points [ref_offset[i], ref_offset[i+1]) (ref_offset[0] == 0);
 * goal_biased and ref_dir may be NULL (all 0 / all +1). */
int clrrt_simulate_batch(clrrt_ctx* ctx, int M, const double* state10, const int32_t* ref_offset, const double* ref_x,
                         const double* ref_y, double* ref_v, const uint8_t* goal_biased, const uint8_t* gen_profile,
                         const double* Vstart, const int32_t* ref_dir, clrrt_rollout* out, double* traj, int traj_stride);

/* == checkObsDistance(states, det, carState) (rrt/src/old_collisioncheck.cpp:24-51; the stub of rrt/src/collisioncheck.cpp
 * when no obstacles are set) for n poses: pose_xytht[4 i ..] = states[0], states[1], states[2] (rear-axle x, y, heading) and
 * states[6] (time, for moving obstacles).  verdict[i] = 1 when the reference returns 0 (collision) — computed by the
 * product's verdict-only path (pose grid, three-way classification, float SAT only in the band around touching), the one
 * every rollout step takes when Wcost[2] == 0.  dobs (optional): the value the reference returns, from the exact path
 * (every obstacle's first separating axis in float, minimum tracked; 100 without obstacles). */
int clrrt_collide_batch(clrrt_ctx* ctx, const double* pose_xytht, int n, int32_t* verdict, double* dobs);

/* == K iterations of expandTree (rrt/src/rrtplanner.cpp:123-174) against ONE tree snapshot: candidate search,
 * rollouts in candidate order until the first success, goal-biased rollout from the node just added, append in
 * sample order.  K == 1 is the reference's sequential algorithm exactly. */
int clrrt_expand_round(clrrt_ctx* ctx, const double* sample_xy, const uint8_t* heuristic, int K,
                       clrrt_round_stats* stats);
/* == n CONSECUTIVE calls of expandTree (the loop of rrt/src/motionplanner.cpp:39-43): sample i sees the nodes appended by
 * samples 0..i-1 — the reference's sequential algorithm, the only formulation that equals it per query (K > 1 rounds are
 * a different algorithm, see clrrt_expand_round).  The result (tree, parents, counters) is exactly that of n calls of
 * clrrt_expand_round with K = 1, including the reference's order of equal keys (clrrt_set_tie_mode), but the device works
 * on a WINDOW of upcoming samples at a time: their candidate searches and rollouts run speculatively against the current
 * tree and are committed in order for as long as no node appended earlier in the window could have changed a sample's
 * candidate list up to its winner (cl-rrt_b200/csrc/sequential.cuh); the first sample that fails the test starts the next
 * window.  window: samples in flight, 1..64 (0 = adaptive).  Does not shard across ranks. */
typedef struct clrrt_seq_stats {
  int64_t iterations, sim_steps, rollouts;
  int32_t windows;          /* speculative windows run */
  int32_t speculated;       /* samples speculated in total (>= iterations; the excess was run twice) */
  int32_t nodes_added, tree_size;
  int32_t exact_fallbacks;  /* samples whose OUTCOME depended on std::sort's order of equal keys: run by the K = 1 host route */
  float ms_total;           /* wall clock of the call */
  /* device time summed over the windows: candidate search; launch order + set-up; the rollout kernel; winner selection,
   * equal-key check and in-order commit */
  float ms_search, ms_prepare, ms_rollout, ms_commit;
  int32_t tie_checks_same;  /* samples whose equal keys were checked against the reference's sort and gave the same outcome */
} clrrt_seq_stats;
int clrrt_expand_sequential(clrrt_ctx* ctx, const double* sample_xy, const uint8_t* heuristic, int n, int window,
                            clrrt_seq_stats* stats);
/* Same, samples already resident in device memory (bench: inputs in HBM before the timed region). */
int clrrt_expand_round_dev(clrrt_ctx* ctx, const double* d_sample_xy, const uint8_t* d_heuristic, int K,
                           clrrt_round_stats* stats);

/* == extractBestPath (rrt/src/rrtplanner.cpp:318-368): ids root..leaf of the cheapest (float costS) goal-reaching
 * branch; *n = 0 when no node reached the goal. */
int clrrt_best_path(clrrt_ctx* ctx, int32_t* ids, int cap, int* n);

int clrrt_counters_get(clrrt_ctx* ctx, clrrt_counters* out);

/* == sampleAroundVehicle (rrt/src/rrtplanner.cpp:187-201) + the heuristic draw (:142-143), K times, on the host
 * C library's rand() (the reference never seeds it: srand(1) stream; the caller may srand()).  Three draws per
 * sample in the reference's order: longitudinal, lateral, heuristic.  heuristic[j] = 0 explore (r <= 0.7), 1 optimize. */
int clrrt_draw_samples(const double goal[4], int K, double* sample_xy, uint8_t* heuristic);

/* Multi-GPU inside the library.  The reference is one process and one thread (expandTree,
 * rrt/include/rrt/rrtplanner.h:87); here one context per GPU holds the full tree, clrrt_expand_round receives THIS RANK's
 * contiguous shard of the round's samples (rank r: samples [r K/world, (r+1) K/world) of the global round), and before it
 * returns the nodes accepted by all ranks are exchanged with ncclAllGather (record counts, then fixed-stride records, on the
 * context's stream over NVLink) and appended in rank order = global sample order = the order of rrt/src/rrtplanner.cpp
 * :150-173, so every rank ends each round with the same tree as a single-GPU round over all K samples.
 *   clrrt_comm_unique_id: ncclGetUniqueId (rank 0 calls it and hands the 128 bytes to the other ranks by any channel);
 *   clrrt_comm_init:      ncclCommInitRank on the context's device (collective: every rank calls it);
 *   clrrt_comm_attach:    use a communicator the caller owns (ncclComm_t) instead.
 * NCCL is bound at run time (libnccl.so.2); without it these return CLRRT_ERR_STATE and single-GPU use is unaffected.
 * clrrt_counters_get_global: the failure counters summed over ranks (collective).  K = 1 sequential mode does not shard. */
#define CLRRT_COMM_ID_BYTES 128
int clrrt_comm_unique_id(void* id, int bytes);
int clrrt_comm_init(clrrt_ctx* ctx, const void* id, int bytes, int rank, int world);
int clrrt_comm_attach(clrrt_ctx* ctx, void* nccl_comm, int rank, int world);
int clrrt_comm_info(const clrrt_ctx* ctx, int* rank, int* world);
int clrrt_counters_get_global(clrrt_ctx* ctx, clrrt_counters* out);
/* Order-sensitive 128-bit digest (sum and xor of per-node hashes of the 160-byte node records and their indices) of tree
 * nodes [first, first + count), computed on the device: equal on every rank and for every world size when the trees are
 * equal — the bench and the multi-rank tests compare it. */
int clrrt_tree_digest(clrrt_ctx* ctx, int first, int count, uint64_t out2[2]);

/* Lower-level form of the same exchange, for callers that move the records themselves: after a round run with append deferred, the nodes accepted by this rank are
 * exposed as fixed-stride records for an all-gather (NCCL), and the gathered records of all ranks are appended
 * in rank order == global sample order, so every rank ends with the same tree for any world size. */
#define CLRRT_RECORD_BYTES 160
int clrrt_set_defer_append(clrrt_ctx* ctx, int defer);
/* Kernel-tuning aid: per-phase clock totals of the rollout kernels (refill, dynamics, collision, finish, warp steps),
 * non-zero only in builds compiled with -DCLRRT_PHASE_CLOCKS. */
int clrrt_debug_phase_clocks(clrrt_ctx* ctx, unsigned long long out[16], int reset);
/* Self-check of the rollout kernel's branch-free double division (csrc/rollout.cuh, div_nb) against the division operator
 * on pairs_per_thread x (SMs x 2048) operand pairs: out = {fast-path results accepted, accepted results that differ (0),
 * pairs left to the operator}. */
int clrrt_debug_div_check(clrrt_ctx* ctx, unsigned long long seed, int pairs_per_thread, unsigned long long out[3]);
int clrrt_round_records(clrrt_ctx* ctx, void** d_records, int* n_records);
int clrrt_append_records(clrrt_ctx* ctx, const void* d_records, const int32_t* counts, int world, int stride_records);

int clrrt_get_device(const clrrt_ctx* ctx);
/* Launch tuning of the rollout kernel: refill_min = idle lanes a warp accumulates before it fetches new work
 * (default 8; 1 = refill immediately); blocks_per_sm = resident blocks of the persistent grid (0 = occupancy maximum). */
int clrrt_set_tuning(clrrt_ctx* ctx, int refill_min, int blocks_per_sm);
/* Order of EQUAL keys in the candidate list.  The reference sorts (node id, key) pairs with std::sort on the key alone
 * (rrt/src/rrtplanner.cpp:233, :256), so equal keys end up in whatever order libstdc++'s introsort leaves them.
 * mode 1 (default): single-sample searches (K = 1, the reference's own sequential algorithm) reproduce that order — when,
 * and only when, a feasible node shares its key with a list entry the search repeats the reference's std::sort call on
 * the host; mode 0: lower node id first everywhere (what batched searches, K > 1, always do).
 * clrrt_tie_sorts: how many searches of this context took the host route. */
int clrrt_set_tie_mode(clrrt_ctx* ctx, int mode);
long long clrrt_tie_sorts(const clrrt_ctx* ctx);
/* Broad-phase grids of the collision check: cell size of the position grid in metres (default 1; the pose grid uses
 * half of it), applied at the next clrrt_set_obstacles; a negative size disables the pose grid, leaving the position
 * grid path only.  Results do not depend on it. */
int clrrt_set_grid_cell(clrrt_ctx* ctx, double metres);
/* Candidate search: 0 = pick by problem size (default), 1 = always sort nodes and samples along the goal bearing first,
 * 2 = never; 16 + L (L = 0..5): always sort, into (axis slab, lateral bin) cells with 2^L lateral bins — the order trees
 * of more than ~25 000 nodes get by default (L = 2, from ~100 000 nodes L = 4).  Results do not depend on it. */
int clrrt_set_nearest_mode(clrrt_ctx* ctx, int mode);

#ifdef __cplusplus
}
#endif
#endif /* CLRRT_H */
