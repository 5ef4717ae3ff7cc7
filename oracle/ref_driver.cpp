// oracle/ref_driver.cpp — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// Builds the *unmodified* reference planner (vdBerg93/cl-rrt, rrt/src/*.cpp, compiled
// from where it lies under /root/reference by oracle/build_ref.sh) into a shared
// library with a flat C interface, so that tests/golden/make_golden.py and the
// bench's cpu_baseline leg can run the reference's own expandTree / Simulation /
// sortNodes* / getOBBdist on chosen inputs.  No reference source is copied into this
// repository: the files are #include'd by name and resolved through -I at build time.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
// legs may load the resulting oracle/_ref/*.so.
//
// The reference is one translation unit (rrt/src/rrt_node.cpp:26 -> rrt/src/include.cpp);
// this file plays the role of rrt_node.cpp: it defines the same file-scope globals
// (rrt/src/rrt_node.cpp:2-24) and then includes the planner sources in the order of
// rrt/src/include.cpp:38-44, with ONE difference: the shipped collision stub
// (rrt/src/collisioncheck.cpp:6-8, "return 100") is replaced by a run-time switch
// between that stub and the uncompiled-upstream OBB/SAT code of
// rrt/src/old_collisioncheck.cpp:24-51.
//
// Variant REF_DEFINED (second .so) is compiled from sed-patched temporaries of three
// reference files (see build_ref.sh) that clamp the reference's out-of-bounds reads
// (SURVEY.md §8c "UB-tainted rollouts"); the patched copies live in a temp dir only.

#include <ros/ros.h>
#include <iostream>
#include <vector>
#include <array>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <ctime>
#include <chrono>
#include <geometry_msgs/Point.h>
#include <visualization_msgs/Marker.h>
#include <visualization_msgs/MarkerArray.h>
#include <std_msgs/Float64MultiArray.h>
#include <std_msgs/MultiArrayDimension.h>
#include <vision_msgs/Detection2DArray.h>
#include "car_msgs/getobstacles.h"
#include "car_msgs/MotionRequest.h"
#include "car_msgs/MotionResponse.h"
#include "car_msgs/State.h"
#include "car_msgs/Trajectory.h"
#include "car_msgs/MotionPlan.h"
#include "car_msgs/Obstacle2D.h"
#include "car_msgs/resetplanner.h"

// ---- globals of rrt/src/rrt_node.cpp:2-24 (same names: the planner sources use them) ----
bool draw_tree = 0;  // upstream 1; only gates rviz marker building in extractBestPath
bool draw_obs = 0;
bool draw_final_path = 0;
bool debug_mode = 0;
bool debug_reference = 0;
bool debug_velocity = 0;
bool draw_states = 0;
bool debug_sim = 0;
bool commit_path = false;
bool obs_use_pred = true;
double Tcommit{0.25};
double sim_dt;
double ctrl_tla, ctrl_dla, ctrl_mindla, ctrl_dlavmin, ctrl_Kp, ctrl_Ki;
double ref_res, ref_int, ref_mindist, vmax, vgoal;
double ay_road_max;
int fail_iterlimit{0};
int fail_collision{0};
int fail_acclimit{0};
int sim_count{0};

#include "rrt/functions.h"
#include "rrt/vehicle.h"
#include "rrt/rrtplanner.h"
#include "rrt/simulation.h"
#include "rrt/collision.h"
#include "rrt/controller.h"
#include "rrt/datatypes.h"
#include "rrt/motionplanner.h"
#include "car_msgs/Reference.h"

// ---- collision hook (the only glue between the live sources and old_collisioncheck.cpp) ----
double checkObsDistance(const vector<double>& states, const vector<car_msgs::Obstacle2D>& det,
                        const vector<double>& carState);
static const vector<car_msgs::Obstacle2D>* g_det = nullptr;  // null => shipped stub behaviour
static vector<double> g_carState(6, 0.0);
static long g_sat_calls = 0;
#include "old_collisioncheck.cpp"
double checkObsDistance(const vector<double>& x) {
  // rrt/src/simulation.cpp:83 calls this 1-argument form after every sim step.
  if (g_det == nullptr || x.size() < 7) return 100;  // == rrt/src/collisioncheck.cpp:6-8
  g_sat_calls += (long)g_det->size();
  return checkObsDistance(x, *g_det, g_carState);  // rrt/src/old_collisioncheck.cpp:24-51
}

#include "reference.cpp"
#include "rrtplanner.cpp"
#include "controller.cpp"
#include "simulation.cpp"
#include "motionplanner.cpp"

// ------------------------------------------------------------------------------------------
namespace {
Vehicle g_veh;
MyRRT* g_rrt = nullptr;
vector<car_msgs::Obstacle2D> g_obstacles;
vector<double> g_goal{50, 0, 0, 0};
bool g_bend = false;                       // MotionRequest.bend / laneShifts / Cxy (rrt/src/motionplanner.cpp:23)
vector<double> g_laneShifts{0}, g_Cxy;

void set_launch_params() {
  // rrt/launch/parameters.launch:3-20
  ros::param::set("ctrl/tla", 1.4);
  ros::param::set("ctrl/mindla", 3.2);
  ros::param::set("ctrl/dlavmin", 3);
  ros::param::set("ctrl/refint", 0.02);
  ros::param::set("ctrl/refmindist", 0.2);
  ros::param::set("ctrl/sampleTime", 0.04);
  ros::param::set("ctrl/Kp", 8);
  ros::param::set("ctrl/Ki", 0.05);
  ros::param::set("motionplanner/weight_distance", 10);
  ros::param::set("motionplanner/weight_curvature", 5);
  ros::param::set("motionplanner/weight_obstacle_gain", 0);
  ros::param::set("motionplanner/weight_obstacle_slope", 4);
  ros::param::set("motionplanner/weight_lanedeviation", 1);
}
void read_params() {
  // rrt/src/rrt_node.cpp:28-38 (updateParameters)
  ros::param::get("ctrl/tla", ctrl_tla);
  ros::param::get("ctrl/mindla", ctrl_mindla);
  ros::param::get("ctrl/dlavmin", ctrl_dlavmin);
  ros::param::get("ctrl/refint", ref_int);
  ros::param::get("ctrl/refmindist", ref_mindist);
  ros::param::get("ctrl/sampleTime", sim_dt);
  ros::param::get("ctrl/Kp", ctrl_Kp);
  ros::param::get("ctrl/Ki", ctrl_Ki);
}
enum { NODE_STRIDE = 20, OUT_STRIDE = 24 };

void fill_out(double* o, const Simulation& sim, const MyReference& ref, int c0, int a0, int i0) {
  const vector<double>& xf = sim.stateArray.back();
  for (int k = 0; k < 10; k++) o[k] = xf[k];
  o[10] = sim.costE;
  o[11] = sim.costS;
  o[12] = sim.endReached;
  o[13] = sim.goalReached;
  o[14] = (double)sim.stateArray.size() - 1;  // executed sim steps
  int code = 0;
  if (fail_collision != c0) code = 1;
  else if (fail_acclimit != a0) code = 2;
  else if (fail_iterlimit != i0) code = 3;
  o[15] = code;
  const int N = (int)ref.x.size();
  o[16] = N;
  o[17] = ref.x.back();
  o[18] = ref.y.back();
  o[19] = ref.v.empty() ? 0.0 : ref.v.back();
  // taint: any step used a waypoint index >= N-2 (then ref.v[IDwp+2] / ref.x[IDwp+1] were read
  // past the end in the unmodified reference).  x[7] logs the IDwp used.
  int tainted = 0;
  double trace = 0;
  for (size_t i = 1; i < sim.stateArray.size(); i++) {
    const int id = (int)sim.stateArray[i][7];
    if (id >= N - 2) tainted = 1;
    trace += (double)i * (double)id;
  }
  o[20] = tainted;
  o[21] = trace;  // checksum of the per-step waypoint-index trace
  o[22] = sim.stateArray[0][7];  // waypoint index chosen by the Controller ctor
  o[23] = 0;
}
}  // namespace

extern "C" {

int ref_is_defined_variant() {
#ifdef REF_DEFINED
  return 1;
#else
  return 0;
#endif
}

// Launch-file parameters + Prius vehicle (rrt/src/motionplanner.cpp:13).
void ref_init(void) {
  // The planner prints through std::cout (e.g. rrt/src/rrtplanner.cpp:345).  Inside a Python process that has
  // loaded extension modules with a statically linked libstdc++ (numpy), the iostream locale facets resolve to
  // foreign GNU-unique symbols and `cout << size_t` crashes; muting the stream skips the facet path.
  std::cout.setstate(std::ios_base::failbit);
  set_launch_params();
  read_params();
  g_veh.setPrius();
  ay_road_max = 0;
  vmax = 5;
  vgoal = 0;
}
void ref_set_weights(const double* w5) {
  ros::param::set("motionplanner/weight_distance", w5[0]);
  ros::param::set("motionplanner/weight_curvature", w5[1]);
  ros::param::set("motionplanner/weight_obstacle_gain", w5[2]);
  ros::param::set("motionplanner/weight_obstacle_slope", w5[3]);
  ros::param::set("motionplanner/weight_lanedeviation", w5[4]);
}
void ref_set_road(int bend, const double* Cxy3, double lane_shift) {
  g_bend = bend != 0;
  g_laneShifts.assign(1, lane_shift);
  g_Cxy.assign(Cxy3, Cxy3 + 3);
}
void ref_get_vehicle(double* v14) {
  const double a[14] = {g_veh.dmax, g_veh.ddmax, g_veh.Td, g_veh.Ta, g_veh.amin, g_veh.amax, g_veh.L,
                        g_veh.w, g_veh.Lrear, g_veh.Lfront, g_veh.b, g_veh.Vch, g_veh.rho, g_veh.Kus};
  memcpy(v14, a, sizeof a);
}
void ref_srand(unsigned seed) { srand(seed); }

// obstacles: n x {cx, cy, theta, size_x, size_y, vx, vy} (car_msgs/msg/Obstacle2D.msg).
// n == 0 selects the shipped stub (rrt/src/collisioncheck.cpp:6-8).
void ref_set_obstacles(const double* o7, int n) {
  g_obstacles.clear();
  for (int i = 0; i < n; i++) {
    car_msgs::Obstacle2D o;
    o.obb.center.x = o7[7 * i + 0];
    o.obb.center.y = o7[7 * i + 1];
    o.obb.center.theta = o7[7 * i + 2];
    o.obb.size_x = o7[7 * i + 3];
    o.obb.size_y = o7[7 * i + 4];
    o.vel.linear.x = o7[7 * i + 5];
    o.vel.linear.y = o7[7 * i + 6];
    g_obstacles.push_back(o);
  }
  g_det = n > 0 ? &g_obstacles : nullptr;
  if (g_rrt) g_rrt->det = g_obstacles;
}

// Start of MotionPlanner::planMotion (rrt/src/motionplanner.cpp:9-32) with commit_path=false:
// counters reset, lookahead / reference resolution from the car speed, new MyRRT, root node.
void ref_tree_init(const double* car_state6, const double* goal4, double vmax_) {
  fail_acclimit = 0; fail_collision = 0; fail_iterlimit = 0; sim_count = 0;
  vector<double> worldState(car_state6, car_state6 + 6);
  vector<double> carPose = transformStateToLocal(worldState);
  updateLookahead(carPose[4]);
  updateReferenceResolution(carPose[4]);
  g_goal.assign(goal4, goal4 + 4);
  vmax = vmax_;
  vgoal = goal4[3];
  delete g_rrt;
  g_rrt = new MyRRT(g_goal, g_laneShifts, g_Cxy, g_bend);
  g_rrt->det = g_obstacles;
  g_rrt->carState = carPose;
  g_carState = carPose;
  vector<Node> none;
  initializeTree(*g_rrt, g_veh, none, carPose);
}
double ref_get_ref_res(void) { return ref_res; }

// iters calls of expandTree (rrt/src/rrtplanner.cpp:123-174); returns the tree size.
int ref_expand(int iters) {
  vector<double> Cxy;
  for (int i = 0; i < iters; i++) expandTree(g_veh, *g_rrt, nullptr, g_obstacles, Cxy);
  return (int)g_rrt->tree.size();
}
// The 200 ms loop of rrt/src/motionplanner.cpp:39-43 with the reference's own Timer (CPU time).
int ref_expand_timed(double budget_ms, int* iters_out) {
  vector<double> Cxy;
  Timer timer(budget_ms);
  int iter = 0;
  for (; timer.Get(); iter++) expandTree(g_veh, *g_rrt, nullptr, g_obstacles, Cxy);
  if (iters_out) *iters_out = iter;
  return (int)g_rrt->tree.size();
}
int ref_tree_size(void) { return g_rrt ? (int)g_rrt->tree.size() : 0; }
void ref_counters(int* c4) {
  c4[0] = fail_collision; c4[1] = fail_acclimit; c4[2] = fail_iterlimit; c4[3] = sim_count;
}
long ref_sat_calls(void) { return g_sat_calls; }

// node record: state[10], ref front x,y, ref back x,y, ref.v.back(), costE, costS, parent, goal, nref
void ref_tree_export(double* out, int cap) {
  const int n = std::min<int>(cap, (int)g_rrt->tree.size());
  for (int i = 0; i < n; i++) {
    const Node& nd = g_rrt->tree[i];
    double* o = out + (size_t)NODE_STRIDE * i;
    for (int k = 0; k < 10; k++) o[k] = nd.state[k];
    o[10] = nd.ref.x.front(); o[11] = nd.ref.y.front();
    o[12] = nd.ref.x.back();  o[13] = nd.ref.y.back();
    o[14] = nd.ref.v.back();
    o[15] = nd.costE; o[16] = nd.costS;
    o[17] = nd.parentID; o[18] = nd.goalReached; o[19] = (double)nd.ref.x.size();
  }
}
// Replace the tree by records in the export format.  Only the fields the batched primitives read are
// rebuilt (state, ref front/back, ref.v.back(), costs, parent, goal flag): the reference paths become
// 2-point stubs, which is all getReference / getGoalReference / feasibleNode / dubinsDistance look at.
void ref_tree_import(const double* in, int n) {
  g_rrt->tree.clear();
  for (int i = 0; i < n; i++) {
    const double* o = in + (size_t)NODE_STRIDE * i;
    vector<double> st(o, o + 10);
    MyReference r;
    r.x = {o[10], o[12]}; r.y = {o[11], o[13]}; r.v = {o[14], o[14]}; r.dir = 1;
    vector<state_type> T{st};
    Node nd(st, (int)o[17], r, T, o[15], o[16], o[18] != 0);
    g_rrt->tree.push_back(nd);
  }
}

// The rollout of rrt/src/rrtplanner.cpp:151-152 (gb=0) or :165-166 (gb=1) for M (parent, sample) pairs.
// out: M x 24 doubles, see fill_out().  Returns elapsed seconds (steady_clock).
double ref_rollout_batch(const int* parent, const double* sample_xy, const unsigned char* gb, int M,
                         double* out) {
  auto t0 = std::chrono::steady_clock::now();
  for (int j = 0; j < M; j++) {
    const int c0 = fail_collision, a0 = fail_acclimit, i0 = fail_iterlimit;
    const Node& p = g_rrt->tree[parent[j]];
    if (gb && gb[j]) {
      MyReference ref = getGoalReference(g_veh, p, g_rrt->goalPose);
      // getGoalReference never sets ref.dir (rrt/src/reference.cpp:34-69): indeterminate upstream.  Inside
      // expandTree the stack slot still holds the 1 written by the preceding getReference (verified: trees
      // grown by the -O3 build match dir=1 bit-for-bit); called from here it would be garbage, so pin it.
      ref.dir = 1;
      Simulation sim(*g_rrt, p.state, ref, g_veh, true, true, p.ref.v.back());
      fill_out(out + (size_t)OUT_STRIDE * j, sim, ref, c0, a0, i0);
    } else {
      geometry_msgs::Point s; s.x = sample_xy[2 * j]; s.y = sample_xy[2 * j + 1];
      MyReference ref = getReference(s, p, 1);
      Simulation sim(*g_rrt, p.state, ref, g_veh, false, true, p.ref.v.back());
      fill_out(out + (size_t)OUT_STRIDE * j, sim, ref, c0, a0, i0);
    }
  }
  return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}
// Full trajectory of one rollout: traj = up to cap x 10 doubles (stateArray); returns its length.
int ref_rollout_traj(int parent, const double* sample_xy, int gb, double* traj, int cap, double* refv, int vcap) {
  const Node& p = g_rrt->tree[parent];
  MyReference ref;
  if (gb) { ref = getGoalReference(g_veh, p, g_rrt->goalPose); ref.dir = 1; }
  else { geometry_msgs::Point s; s.x = sample_xy[0]; s.y = sample_xy[1]; ref = getReference(s, p, 1); }
  Simulation sim(*g_rrt, p.state, ref, g_veh, gb != 0, true, p.ref.v.back());
  const int n = std::min<int>(cap, (int)sim.stateArray.size());
  for (int i = 0; i < n; i++) for (int k = 0; k < 10; k++) traj[10 * i + k] = sim.stateArray[i][k];
  if (refv) for (int i = 0; i < std::min<int>(vcap, (int)ref.v.size()); i++) refv[i] = ref.v[i];
  return (int)sim.stateArray.size();
}

// sortNodesExplore (heuristic 0) / sortNodesOptimize (1), rrt/src/rrtplanner.cpp:227-268.
// cand: K x 10 (padded with -1), key: K x 10 float keys of the chosen nodes, count: K.
double ref_nearest_batch(const double* sample_xy, const unsigned char* heuristic, int K, int* cand,
                         float* key, int* count) {
  auto t0 = std::chrono::steady_clock::now();
  for (int j = 0; j < K; j++) {
    geometry_msgs::Point s; s.x = sample_xy[2 * j]; s.y = sample_xy[2 * j + 1];
    vector<int> v = heuristic[j] ? sortNodesOptimize(*g_rrt, s) : sortNodesExplore(*g_rrt, s);
    count[j] = (int)v.size();
    for (int r = 0; r < 10; r++) {
      const bool ok = r < (int)v.size();
      cand[10 * j + r] = ok ? v[r] : -1;
      float k = 0;
      if (ok) {
        k = dubinsDistance(s, g_rrt->tree[v[r]], g_rrt->direction);
        if (heuristic[j]) k = g_rrt->tree[v[r]].costE + k;
      }
      key[10 * j + r] = k;
    }
  }
  return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}
// All keys of one sample (for near-tie audits): key[n_nodes], feasible[n_nodes].
void ref_keys(const double* sample_xy, int heuristic, float* key, unsigned char* feas) {
  geometry_msgs::Point s; s.x = sample_xy[0]; s.y = sample_xy[1];
  for (size_t i = 0; i < g_rrt->tree.size(); i++) {
    float k = dubinsDistance(s, g_rrt->tree[i], g_rrt->direction);
    if (heuristic) k = g_rrt->tree[i].costE + k;
    key[i] = k;
    feas[i] = feasibleNode(*g_rrt, g_rrt->tree[i], s);
  }
}
float ref_dubins(double sx, double sy, double nx, double ny, double nth, int dir) {
  geometry_msgs::Point s; s.x = sx; s.y = sy;
  Node n; n.state = {nx, ny, nth, 0, 0, 0, 0, 0, 0, 0};
  return dubinsDistance(s, n, dir);
}
// getOBBdist(OBB a, OBB b), rrt/src/old_collisioncheck.cpp:98-148; each box = {x, y, w, h, o}.
double ref_obb_dist(const double* a5, const double* b5) {
  OBB a(Vector2D(a5[0], a5[1]), a5[2], a5[3], a5[4]);
  OBB b(Vector2D(b5[0], b5[1]), b5[2], b5[3], b5[4]);
  return getOBBdist(a, b);
}
// checkObsDistance(states, det, carState), rrt/src/old_collisioncheck.cpp:24-51, on one 10-state.
double ref_obs_distance(const double* x10) {
  vector<double> x(x10, x10 + 10);
  return checkObsDistance(x, g_obstacles, g_carState);
}
void ref_obs_distance_batch(const double* pose4, int n, double* out) {
  for (int i = 0; i < n; i++) {
    vector<double> x = {pose4[4 * i], pose4[4 * i + 1], pose4[4 * i + 2], 0, 0, 0, pose4[4 * i + 3], 0, 0, 0};
    out[i] = checkObsDistance(x);  // the hook at rrt/src/simulation.cpp:83 (stub without obstacles)
  }
}
// Simulation::Simulation(RRT, state, ref, veh, GoalBiased, genProfile, Vstart) (rrt/include/rrt/simulation.h:18-19) on a
// caller-supplied reference; rv is filled when genProfile (the constructor mutates the caller's MyReference).
int ref_simulate(const double* state10, const double* rx, const double* ry, double* rv, int n, int dir, int gb, int genProfile,
                 double Vstart, double* out, double* traj, int cap) {
  const int c0 = fail_collision, a0 = fail_acclimit, i0 = fail_iterlimit;
  MyReference ref;
  ref.x.assign(rx, rx + n); ref.y.assign(ry, ry + n); ref.dir = dir;
  if (!genProfile) ref.v.assign(rv, rv + n);
  vector<double> st(state10, state10 + 10);
  Simulation sim(*g_rrt, st, ref, g_veh, gb != 0, genProfile != 0, Vstart);
  fill_out(out, sim, ref, c0, a0, i0);
  for (int i = 0; i < n && i < (int)ref.v.size(); i++) rv[i] = ref.v[i];
  const int rows = std::min<int>(cap, (int)sim.stateArray.size());
  if (traj) for (int i = 0; i < rows; i++) for (int k = 0; k < 10; k++) traj[10 * i + k] = sim.stateArray[i][k];
  return (int)sim.stateArray.size();
}
// Road-frame transforms of rrt/src/transformations.cpp:20-202 on n records (x, y, heading, delta), in place; `what` as in
// clrrt_host_road_transform (include/clrrt_host.h).
int ref_road_transform(int what, const double* Cxy3, const double* Cxs3, double* xyhd, int n) {
  vector<double> Cxy(Cxy3, Cxy3 + 3), Cxs(Cxs3, Cxs3 + 3);
  for (int i = 0; i < n; i++) {
    double* p = xyhd + 4 * (size_t)i;
    if (what == 0) transformPointCarToRoad(p[0], p[1], Cxy, Cxs);
    else if (what == 1) transformPointRoadToCar(p[0], p[1], Cxy, Cxs);
    else if (what == 2) transformPoseCarToRoad(p[0], p[1], p[2], Cxy, Cxs);
    else if (what == 3) transformPoseRoadToCar(p[0], p[1], p[2], Cxy, Cxs);
    else if (what == 6) { vector<double> a = findClosestPointOnArc(p[0], p[1], Cxy); p[0] = a[0]; p[1] = a[1]; }
    else {
      state_type s = {p[0], p[1], p[2], p[3], 0, 0};
      if (what == 4) transformStateCarToRoad(s, Cxy, Cxs, g_veh); else transformStateRoadToCar(s, Cxy, Cxs, g_veh);
      for (int k = 0; k < 4; k++) p[k] = s[k];
    }
  }
  return 0;
}
// sampleOnLane (rrt/src/rrtplanner.cpp:204-224) + the heuristic draw, K times; v sets the look-ahead global ctrl_dla.
int ref_sample_on_lane(const double* Cxy3, const double* lane_shifts, int n_lanes, double Lmax, double v, int K,
                       double* sample_xy, unsigned char* heuristic) {
  vector<double> Cxy(Cxy3, Cxy3 + 3), ls(lane_shifts, lane_shifts + n_lanes);
  updateLookahead(v);
  for (int j = 0; j < K; j++) {
    geometry_msgs::Point s = sampleOnLane(Cxy, ls, Lmax);
    sample_xy[2 * j] = s.x; sample_xy[2 * j + 1] = s.y;
    double r = static_cast<double>(rand()) / (static_cast<double>(RAND_MAX / (1)));
    heuristic[j] = !(r <= 0.7);
  }
  return 0;
}
// sampleAroundVehicle + the heuristic draw, in the order of rrt/src/rrtplanner.cpp:133-143.
void ref_draw_samples(int K, double* sample_xy, unsigned char* heuristic, double* r_out) {
  for (int j = 0; j < K; j++) {
    geometry_msgs::Point s = sampleAroundVehicle(g_goal);
    double r = static_cast<double>(rand()) / (static_cast<double>(RAND_MAX / (1)));
    sample_xy[2 * j] = s.x; sample_xy[2 * j + 1] = s.y;
    heuristic[j] = !(r <= 0.7);
    if (r_out) r_out[j] = r;
  }
}
int ref_feasible_goal_bias(void) { return feasibleGoalBias(*g_rrt); }
// extractBestPath (rrt/src/rrtplanner.cpp:318-368): node ids root..leaf of the cheapest goal-reaching branch.
int ref_best_path(int* ids, int cap) {
  vector<Node> best = extractBestPath(g_rrt->tree, nullptr);
  // recover ids by walking parents from the cheapest goal node (same rule as :347-358)
  int bestid = -1; float bestc = 0;
  for (size_t i = 0; i < g_rrt->tree.size(); i++)
    if (g_rrt->tree[i].goalReached && (bestid < 0 || (double)g_rrt->tree[i].costS < (double)bestc)) {
      bestid = (int)i; bestc = g_rrt->tree[i].costS;
    }
  vector<int> chain;
  for (int id = bestid; id != -1; id = g_rrt->tree[id].parentID) chain.insert(chain.begin(), id);
  const int n = std::min<int>(cap, (int)chain.size());
  for (int i = 0; i < n; i++) ids[i] = chain[i];
  return (int)best.size();
}
// ---- receding-horizon queries with a carried-over tree (config C5) ------------------------------
// MotionPlanner::planMotion (rrt/src/motionplanner.cpp:8-54) with commit_path = true, restated as the same sequence
// of calls into the reference's own functions — transformNodesWorldToCar, MyRRT, initializeTree (non-empty branch,
// rrt/src/rrtplanner.cpp:50-94), expandTree, extractBestPath, transformNodesCarToworld — with `iters` expandTree calls
// in place of Timer(200) so that the result is deterministic.  Goal and obstacles are given in the car frame, as the
// planner receives them from the mission planner / detection node.  MotionPlanner::bestNodes lives in g_bestNodes.
static vector<Node> g_bestNodes;
void ref_commit_reset(void) { g_bestNodes.clear(); }
int ref_query_commit(const double* world_state6, const double* goal4, double vmax_, int iters, int* sizes4, double* best_cost) {
  commit_path = true;
  fail_acclimit = 0; fail_collision = 0; fail_iterlimit = 0; sim_count = 0;
  vector<double> worldState(world_state6, world_state6 + 6);
  vector<double> carPose = transformStateToLocal(worldState);
  updateLookahead(carPose[4]);
  updateReferenceResolution(carPose[4]);
  vmax = vmax_;
  vgoal = goal4[3];
  transformNodesWorldToCar(g_bestNodes, worldState);
  g_goal.assign(goal4, goal4 + 4);
  delete g_rrt;
  vector<double> laneShifts{0}, Cxy;
  g_rrt = new MyRRT(g_goal, laneShifts, Cxy, false);
  g_rrt->det = g_obstacles;
  g_rrt->carState = carPose;
  g_carState = carPose;
  initializeTree(*g_rrt, g_veh, g_bestNodes, carPose);
  sizes4[0] = (int)g_rrt->tree.size();
  for (int i = 0; i < iters; i++) expandTree(g_veh, *g_rrt, nullptr, g_obstacles, Cxy);
  sizes4[1] = (int)g_rrt->tree.size();
  g_bestNodes = extractBestPath(g_rrt->tree, nullptr);
  sizes4[2] = (int)g_bestNodes.size();
  sizes4[3] = sim_count;
  if (best_cost) *best_cost = g_bestNodes.empty() ? -1.0 : (double)g_bestNodes.back().costS;
  transformNodesCarToworld(g_bestNodes, worldState);
  commit_path = false;
  return sizes4[1];
}
// bestNodes after the last query (world frame): per node {state[10], ref front xy, ref back xy, ref.v.back(), costE,
// costS, parentID, goalReached, ref.x.size()}; returns the node count
int ref_best_nodes(double* rec20, int cap) {
  int n = 0;
  for (const Node& nd : g_bestNodes) {
    if (n >= cap) break;
    double* o = rec20 + (size_t)NODE_STRIDE * n++;
    for (int k = 0; k < 10; k++) o[k] = nd.state[k];
    o[10] = nd.ref.x.front(); o[11] = nd.ref.y.front(); o[12] = nd.ref.x.back(); o[13] = nd.ref.y.back();
    o[14] = nd.ref.v.empty() ? 0.0 : nd.ref.v.back();
    o[15] = nd.costE; o[16] = nd.costS; o[17] = nd.parentID; o[18] = nd.goalReached; o[19] = (double)nd.ref.x.size();
  }
  return (int)g_bestNodes.size();
}
// concatenated trajectories of bestNodes (rows of 10 doubles, world x/y, headings as stored); rows_per_node[i] = tra.size()
int ref_best_traj(double* traj10, int cap_rows, int* rows_per_node, int cap_nodes) {
  int rows = 0, n = 0;
  for (const Node& nd : g_bestNodes) {
    if (n < cap_nodes) rows_per_node[n] = (int)nd.tra.size();
    n++;
    for (const auto& x : nd.tra) {
      if (rows < cap_rows) for (int k = 0; k < 10; k++) traj10[(size_t)10 * rows + k] = x[k];
      rows++;
    }
  }
  return rows;
}
}  // extern "C"
