/* clrrt_host.h — C view of the ROS-free host facade (cl-rrt_b200/host, libclrrt_host.so).
 *
 * The reference's query-level surface is C++ (MotionPlanner::planMotion / updateObstacles / updateState,
 * rrt/include/rrt/motionplanner.h:20-42) fed by ROS messages.  The C++ mirror of it lives in
 * cl-rrt_b200/host/clrrt_planner.hpp; the functions below expose the same calls over plain pointers for bindings and
 * tests, together with the ROS 1 wire format of the messages involved (car_msgs/msg/*.msg), so that a node can hand
 * raw message buffers to the planner.  All planning runs on the GPU through include/clrrt.h; there is no CPU path.
 */
#ifndef CLRRT_HOST_H
#define CLRRT_HOST_H

#include "clrrt.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Error convention: every call returns CLRRT_OK / a count, or a negative clrrt_status; nothing is thrown across the
 * boundary and nothing is printed.  clrrt_host_last_error() is the text of the last failure on the calling thread (the
 * reference itself has no error channel: failures are flags and log lines, rrt/src/motionplanner.cpp:56-58). */
const char* clrrt_host_last_error(void);

/* One MotionPlanner::planMotion query from an empty tree (rrt/src/motionplanner.cpp:8-77, commit_path = false).
 * max_iterations >= 0 replaces Timer(200) by a fixed number of expandTree calls (deterministic); otherwise budget_ms
 * of wall clock.  seed: srand(seed) before the query (the reference never seeds rand(): 1 reproduces it).
 * traj8: the published car_msgs/Trajectory after filterMPCmessage, rows x y theta delta v a a_cmd d_cmd. */
int clrrt_host_plan_motion(const double* car_state6, const double* goal4, double vmax, const clrrt_obstacle* obs,
                           int n_obs, int samples_per_round, int max_iterations, double budget_ms, unsigned seed,
                           int device, int* tree_size, int* iterations, clrrt_counters* counters, double* traj8,
                           int traj_cap, int* traj_len, int32_t* best_ids, int best_cap, int* best_len, double* remat_err);

/* clrrt::Simulation with the reference's own parameter list (rrt/include/rrt/simulation.h:18-19) through the C++ facade:
 * Prius vehicle, launch-file parameters, goal / vmax / obstacles as given; state has n_state (6 or 10) entries; the
 * reference path is caller-owned (n_ref >= 3 points) and ref_v is FILLED when gen_profile != 0.  Returns
 * stateArray.size() (or a negative status); costs2 = {costE, costS}; flags3 = {endReached, goalReached, failCode};
 * last_state10 = stateArray.back(). */
int clrrt_host_simulate(const double* goal4, double vmax, const clrrt_obstacle* obs, int n_obs, int device, const double* state,
                        int n_state, const double* ref_x, const double* ref_y, double* ref_v, int n_ref, int goal_biased,
                        int gen_profile, double Vstart, double* costs2, int32_t* flags3, double* last_state10);

/* Curved-road mode, host side (cl-rrt_b200/host/clrrt_road.hpp; rrt/src/transformations.cpp:20-287, rrt/src/rrtplanner.cpp
 * :204-224 — present upstream but not reachable from the shipped planMotion).  clrrt_host_road_transform applies one of the
 * road-frame transforms in place to n records (x, y, heading, delta): what = 0 point car->road, 1 point road->car, 2 pose
 * car->road, 3 pose road->car, 4 state car->road, 5 state road->car (Prius), 6 closest point on the road's arc.
 * clrrt_host_sample_on_lane: K x { sampleOnLane, heuristic draw } on rand(), v = the car's speed (look-ahead distance). */
int clrrt_host_road_transform(int what, const double* Cxy3, const double* Cxs3, double* xyhd, int n);
int clrrt_host_sample_on_lane(const double* Cxy3, const double* lane_shifts, int n_lanes, double Lmax, double v, int K,
                              double* sample_xy, uint8_t* heuristic);

/* Persistent planner: consecutive queries on one device context with MotionPlanner::bestNodes kept between them
 * (commit_path != 0: the next tree is initialised from the previous best path, rrt/src/rrtplanner.cpp:50-94). */
void* clrrt_host_planner_create(int device, int samples_per_round, int commit_path, int tree_capacity);
void clrrt_host_planner_destroy(void* h);
/* samples_per_round == 1 (the reference's sequential loop): iterations per device call and samples in flight inside a call
 * (clrrt_expand_sequential).  chunk = 1 issues one clrrt_expand_round per iteration; default chunk 32, window 0 = adaptive.
 * The tree is the same for every setting. */
int clrrt_host_planner_set_sequential(void* h, int chunk, int window);
/* wall-clock milliseconds of the last query by phase: parameters, obstacles, initial tree, expansion, best path, messages;
 * returns the number of speculative windows the expansion used (sequential mode) */
int clrrt_host_planner_timings(void* h, double* ms6);
/* goal4 and obs in the car frame; sizes4 = {initial tree size, final tree size, best path nodes, expandTree calls} */
int clrrt_host_planner_query(void* h, const double* world_state6, const double* goal4, double vmax, const clrrt_obstacle* obs,
                             int n_obs, int max_iterations, double budget_ms, int32_t* sizes4, double* best_cost,
                             clrrt_counters* counters);
/* bestNodes of the last query, world frame: {state[10], ref front xy, ref back xy, ref.v.back(), costE, costS, parentID,
 * goalReached, ref.x.size()} per node; returns the node count */
int clrrt_host_planner_best_nodes(void* h, double* rec20, int cap);
/* their trajectories, concatenated rows of 10 state entries; rows_per_node[i] = Node::tra.size(); returns the row count */
int clrrt_host_planner_best_traj(void* h, double* traj10, int cap_rows, int32_t* rows_per_node, int cap_nodes);

/* ROS 1 wire format (cl-rrt_b200/host/clrrt_wire.hpp): message bytes -> planner inputs, planner outputs -> message bytes.
 * Parsers return CLRRT_OK (or the element count) and a negative code on truncated / malformed buffers; writers return
 * the byte count or CLRRT_ERR_CAPACITY. */
int clrrt_wire_parse_request(const uint8_t* buf, int n, double* goal4, double* vmax, int* bend, int* n_lane_shifts);
int clrrt_wire_parse_state(const uint8_t* buf, int n, double* state6);
int clrrt_wire_parse_obstacles(const uint8_t* buf, int n, clrrt_obstacle* out, int cap);
int clrrt_wire_trajectory(const double* rows8, int n_rows, uint8_t* out, int cap);
int clrrt_wire_obstacles(const clrrt_obstacle* obs, int n, uint8_t* out, int cap);
int clrrt_wire_response_roundtrip(const double* rows10, const int32_t* rows_per_segment, int n_segments, uint8_t* out, int cap);

#ifdef __cplusplus
}
#endif
#endif
