// clrrt_planner.hpp — ROS-free C++ host facade over the C ABI (include/clrrt.h).
//
// Keeps the shapes of the reference's own C++ surface for the tree-expansion path so that its call sites read
// the same (names, argument meaning, public result fields):
//   Vehicle                         rrt/include/rrt/vehicle.h:5-61
//   MyReference, Node, MyRRT        rrt/include/rrt/rrtplanner.h:27-80
//   Simulation                      rrt/include/rrt/simulation.h:7-26   (all work in the constructor)
//   MotionPlanner::planMotion / updateObstacles / updateState / resetPlanner
//                                   rrt/include/rrt/motionplanner.h:20-42, rrt/src/motionplanner.cpp:8-100
//   MotionRequest, Obstacle2D, Trajectory   car_msgs/msg/*.msg (plain structs, same field names)
// The ROS pieces (publishers, the getobstacles service client) become std::function hooks.  All numerics run in
// the CUDA library; this layer only marshals.  There is no CPU fallback: constructing a MyRRT without a CUDA
// device throws.
#pragma once
#include <cstdint>
#include <functional>
#include <stdexcept>
#include <string>
#include <vector>

#include "clrrt.h"

namespace clrrt {

typedef std::vector<double> state_type;               // rrt/include/rrt/datatypes.h:11
typedef std::vector<std::vector<double>> StateArray;  // rrt/include/rrt/datatypes.h:12

struct Point { double x = 0, y = 0, z = 0; };  // geometry_msgs/Point

// rrt/include/rrt/vehicle.h
class Vehicle {
 public:
  double dmax, ddmax, Td, Ta, amin, amax, L, w, Lrear, Lfront, b, Vch, rho, Kus;
  void setTalos();
  void setPrius();
};

// car_msgs/msg/Obstacle2D.msg (vision_msgs/BoundingBox2D obb + geometry_msgs/Twist vel)
struct Obstacle2D {
  struct { struct { double x = 0, y = 0, theta = 0; } center; double size_x = 0, size_y = 0; } obb;
  struct { struct { double x = 0, y = 0, z = 0; } linear, angular; } vel;
};
// car_msgs/msg/MotionRequest.msg
struct MotionRequest {
  std::vector<double> goal;
  double vmax = 0;
  bool bend = false;
  std::vector<double> Cxy, Cxs, laneShifts;
};
// car_msgs/msg/Trajectory.msg
struct Trajectory { std::vector<double> x, y, theta, delta, v, a, a_cmd, d_cmd; };

// rrt/include/rrt/rrtplanner.h:27-33
struct MyReference {
  std::vector<double> x, y, v;
  signed int dir = 1;
  double aend = 0;
};
// rrt/include/rrt/rrtplanner.h:35-47.  ref holds front/back points (the device keeps no full paths); tra is filled
// for the nodes of an extracted best path.
struct Node {
  std::vector<double> state;
  signed int parentID = -1;
  MyReference ref;
  float costE = 0, costS = 0;
  bool goalReached = false;
  std::vector<state_type> tra;
};
struct Path {  // rrt/include/rrt/motionplanner.h:4-7
  MyReference ref;
  std::vector<state_type> tra;
};

struct PlannerParams {  // rrt/launch/parameters.launch:3-20 and rrt/src/rrt_node.cpp:2-24
  double sim_dt = 0.04, ctrl_tla = 1.4, ctrl_mindla = 3.2, ctrl_dlavmin = 3, ctrl_Kp = 8, ctrl_Ki = 0.05;
  double ref_int = 0.02, ref_mindist = 0.2;
  double Wcost[5] = {10, 5, 0, 4, 1};
  bool commit_path = false, obs_use_pred = true;
};

class Error : public std::runtime_error {
 public:
  using std::runtime_error::runtime_error;
};

// rrt/include/rrt/rrtplanner.h:51-80.  Owns one device context; `tree` lives on the GPU.
class MyRRT {
 public:
  int sortLimit = CLRRT_SORT_LIMIT;
  bool goalReached = false, bend = false;
  std::vector<double> goalPose, laneShifts, Cxy;
  signed int direction = 1;
  std::vector<Obstacle2D> det;
  std::vector<double> carState;
  double Wcost[5];
  bool treeFull = false;  // set by expandTree when a round hit the tree capacity (the fitting prefix was appended)

  MyRRT(const std::vector<double>& goalPose, const std::vector<double>& laneShifts, const std::vector<double>& Cxy,
        const bool& bend, const Vehicle& veh, const PlannerParams& prm, double vmax, double car_speed, int device = 0,
        int tree_capacity = 1 << 18, int max_round = 1 << 14);
  ~MyRRT();
  MyRRT(const MyRRT&) = delete;
  MyRRT& operator=(const MyRRT&) = delete;

  void addInitialNode(const std::vector<double>& state);  // rrt/src/rrtplanner.cpp:21-37
  // the carried-over chain of initializeTree's non-empty branch (rrt/src/rrtplanner.cpp:89-93): uploaded as the initial
  // tree; the host copies keep the full references and trajectories for extractBestPath
  void setCarriedTree(const std::vector<Node>& nodes);
  const std::vector<Node>& carried() const { return carried_; }
  // next query on the same device context: goal, vmax and the reference resolution change (motionplanner.cpp:16-23)
  void reconfigure(const std::vector<double>& goalPose, double vmax, double car_speed, const PlannerParams& prm);
  void setObstacles(const std::vector<Obstacle2D>& det);  // RRT.det = det, rrt/src/motionplanner.cpp:24
  int treeSize() const;
  std::vector<Node> tree() const;                          // download (ref = front/back points)
  clrrt_ctx* ctx() const { return ctx_; }
  const clrrt_params& params() const { return prm_; }

 private:
  void setRoad();
  clrrt_ctx* ctx_ = nullptr;
  clrrt_params prm_;
  std::vector<Node> carried_;
};

// expandTree (rrt/src/rrtplanner.cpp:123-174), K samples per call against one tree snapshot.  K == 1 is the
// reference's algorithm; samples come from sampleAroundVehicle on the C library's rand(), as upstream.
clrrt_round_stats expandTree(Vehicle& veh, MyRRT& RRT, int K = 1);
// n consecutive expandTree calls with K = 1 — same draws, tree and counters — with `window` samples in flight on the device
// (clrrt_expand_sequential; 0 = adaptive)
clrrt_seq_stats expandTreeSequential(Vehicle& veh, MyRRT& RRT, int n, int window = 0);
// extractBestPath (rrt/src/rrtplanner.cpp:318-368) with the trajectories of the returned nodes re-materialised.
std::vector<Node> extractBestPath(MyRRT& RRT, std::vector<int32_t>* ids_out = nullptr);
// initializeTree (rrt/src/rrtplanner.cpp:39-95): empty committed path -> single root node; otherwise the previous best
// path (already in the car frame) becomes the initial tree: nodes ending behind the car are dropped, goal flags and the
// costS chain are recomputed from the stored trajectories, the nodes are re-parented as a chain.
void initializeTree(MyRRT& RRT, const Vehicle& veh, std::vector<Node>& nodes, std::vector<double>& carState);
// getNodeCost (rrt/src/rrtplanner.cpp:105-119)
double getNodeCost(const MyRRT& RRT, const Vehicle& veh, const double& parentCost, const Node& node, double sim_dt);

// rrt/include/rrt/simulation.h:7-26.  As upstream, everything happens in the constructor.
class Simulation {
 public:
  StateArray stateArray;
  double costS = 0, costE = 0;
  bool goalReached = false, endReached = false;
  MyReference ref;  // x, y re-materialised on the host by LinearSpacedVector; v holds ref.v.back() only
  int failCode = 0; // 0 none, 1 collision, 2 lateral acceleration, 3 iteration limit
  // the reference's own parameter list (rrt/include/rrt/simulation.h:18-19): any start state (6 or 10 entries), a caller-owned
  // reference of any shape; with genProfile the constructor FILLS ref.v, as upstream (the caller stores that ref in the new
  // Node, rrt/src/rrtplanner.cpp:156).  veh: the planner's vehicle (the context was created with it).
  Simulation(const MyRRT& RRT, const std::vector<double>& state, MyReference& ref, const Vehicle& veh, const bool& GoalBiased,
             const bool& genProfile, const double& Vstart);
  // shorthand for the rollouts expandTree itself starts: from tree node `parent` towards `sample` (rrtplanner.cpp:151-152, :165-166)
  Simulation(const MyRRT& RRT, int parent, const Point& sample, const Vehicle& veh, const bool& GoalBiased);
  bool isvalid() const { return endReached; }
};

// rrt/include/rrt/motionplanner.h:20-42
struct MotionPlanner {
  std::vector<Path> motionplan;
  std::vector<Node> bestNodes;
  state_type state{0, 0, 0, 0, 0, 0};
  std::vector<Obstacle2D> det;
  // hooks replacing the ROS plumbing
  std::function<std::vector<Obstacle2D>()> getobstacles;   // the "getobstacles" service, rrt/src/rrt_node.cpp:73
  std::function<void(const Trajectory&)> pubMPC;           // "/path_publisher/path", rrt/src/rrt_node.cpp:63
  PlannerParams params;
  int device = 0;
  int samplesPerRound = 1;        // 1 = the reference's sequential expandTree; >1 = snapshot rounds
  int sequentialChunk = 32;       // samplesPerRound == 1: iterations per device call (1 = one clrrt_expand_round per iteration)
  int sequentialWindow = 0;       // samples in flight inside a call (0 = adaptive), see clrrt_expand_sequential
  double budget_ms = 200;         // Timer(200), rrt/src/motionplanner.cpp:39 (wall clock here, CPU time upstream)
  int maxIterations = -1;         // >= 0: deterministic iteration budget instead of the timer (tests)
  // results of the last query
  Trajectory lastTrajectory;
  int lastTreeSize = 0, lastIterations = 0;
  std::vector<int32_t> lastBestIds;   // tree indices of bestNodes
  double lastRematError = 0;          // max |tra.back() - node.state| over the best path (0: re-materialisation is exact)
  clrrt_counters lastCounters{};
  int lastCarried = 0;                // nodes of the previous best path the tree was initialised with (commit_path)
  int lastInitialTree = 0;            // tree size after initializeTree (1 = root only)
  int treeCapacity = 1 << 18;
  int lastSeqWindows = 0, lastSeqSpeculated = 0;  // sequential mode: windows run / samples speculated in the last query
  double lastMs[6] = {0, 0, 0, 0, 0, 0};  // wall clock of the last query: parameters, obstacles, initial tree, expansion, best path, messages
  ~MotionPlanner();

  void planMotion(MotionRequest req);                       // rrt/src/motionplanner.cpp:8-77
  bool updateObstacles();                                   // rrt/src/motionplanner.cpp:81-86
  void updateState(const std::vector<double>& msg_state);   // rrt/src/motionplanner.cpp:89-94
  bool resetPlanner();                                      // rrt/src/motionplanner.cpp:98-100

 private:
  MyRRT* rrt_ = nullptr;              // one device context for all queries of this planner
  int rrt_round_ = 0;
};

std::vector<Path> convertNodesToPath(const std::vector<Node>& path);   // rrt/src/motionplanner.cpp:264-275
Trajectory generateMPCmessage(const std::vector<Path>& path);          // rrt/src/motionplanner.cpp:103-128
void filterMPCmessage(Trajectory& msg);                                // rrt/src/motionplanner.cpp:130-150
void transformNodesCarToworld(std::vector<Node>& nodes, const std::vector<double> carState);  // transformations.cpp:289-301
void transformNodesWorldToCar(std::vector<Node>& nodes, const std::vector<double> carState);  // transformations.cpp:303-315

}  // namespace clrrt

// flat C view of the facade for bindings/tests (same library): include/clrrt_host.h
#include "clrrt_host.h"
