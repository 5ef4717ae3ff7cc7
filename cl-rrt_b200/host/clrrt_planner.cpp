// clrrt_planner.cpp — implementation of the ROS-free host facade (see clrrt_planner.hpp).
#include "clrrt_planner.hpp"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <cstring>

namespace clrrt {

namespace {
void ck(clrrt_ctx* ctx, int rc, const char* what) {
  if (rc != CLRRT_OK) throw Error(std::string(what) + " failed (" + std::to_string(rc) + "): " + (ctx ? clrrt_last_error(ctx) : ""));
}
// LinearSpacedVector, rrt/include/rrt/functions.h:11-21
std::vector<double> LinearSpacedVector(double a, double b, std::size_t N) {
  double h = (b - a) / static_cast<double>(N - 1);
  std::vector<double> xs(N);
  double val = a;
  for (auto it = xs.begin(); it != xs.end(); ++it, val += h) *it = val;
  return xs;
}
Node node_from_c(const clrrt_node& c) {
  Node n;
  n.state.assign(c.state, c.state + 10);
  n.parentID = c.parent;
  n.ref.x = {c.ref_front[0], c.ref_back[0]};
  n.ref.y = {c.ref_front[1], c.ref_back[1]};
  n.ref.v = {c.ref_vback};
  n.costE = c.costE; n.costS = c.costS; n.goalReached = c.goal_reached != 0;
  return n;
}
}  // namespace

void Vehicle::setTalos() {  // rrt/include/rrt/vehicle.h:23-38
  dmax = 0.5435; ddmax = 0.3294; Td = 0.3; Ta = 0.3; amin = -6; amax = 2; L = 2.885; Lrear = 1;
  Lfront = 4.848 - Lrear; w = 2; b = 1.8; Vch = 20; rho = 4.77; Kus = 0.018;
}
void Vehicle::setPrius() {  // rrt/include/rrt/vehicle.h:39-60
  dmax = 0.52; ddmax = 0.3294; Td = 0.3; Ta = 0.3; amin = -6; amax = 2; L = 2.7;
  double lf = 1.0868, lr = 1.6132;
  Lrear = 1; Lfront = 2.7 + 0.5; w = 2; b = lr; rho = 5.95;
  double Cf = 22201, Cr = 22201, m = 950 + 640;
  Kus = (m / L) * (lr / Cf - lf / Cr);
  Vch = 20;
}

MyRRT::MyRRT(const std::vector<double>& _goalPose, const std::vector<double>& _laneShifts, const std::vector<double>& _Cxy,
             const bool& _bend, const Vehicle& veh, const PlannerParams& prm, double vmax, double car_speed, int device,
             int tree_capacity, int max_round)
    : bend(_bend), goalPose(_goalPose), laneShifts(_laneShifts), Cxy(_Cxy) {
  if (goalPose.size() < 4) throw Error("goalPose needs 4 entries");
  clrrt_default_params(&prm_);
  setRoad();
  clrrt_vehicle& v = prm_.veh;
  v.dmax = veh.dmax; v.ddmax = veh.ddmax; v.Td = veh.Td; v.Ta = veh.Ta; v.amin = veh.amin; v.amax = veh.amax; v.L = veh.L;
  v.w = veh.w; v.Lrear = veh.Lrear; v.Lfront = veh.Lfront; v.b = veh.b; v.Vch = veh.Vch; v.rho = veh.rho; v.Kus = veh.Kus;
  prm_.sim_dt = prm.sim_dt; prm_.ctrl_tla = prm.ctrl_tla; prm_.ctrl_mindla = prm.ctrl_mindla; prm_.ctrl_dlavmin = prm.ctrl_dlavmin;
  prm_.ctrl_Kp = prm.ctrl_Kp; prm_.ctrl_Ki = prm.ctrl_Ki; prm_.ref_int = prm.ref_int; prm_.ref_mindist = prm.ref_mindist;
  // updateReferenceResolution(v), rrt/src/controller.cpp:18-21 (rrt/src/motionplanner.cpp:16)
  prm_.ref_res = std::max(std::abs(car_speed) * prm.ref_int, prm.ref_mindist);
  prm_.vmax = vmax;  // rrt/src/motionplanner.cpp:17
  for (int i = 0; i < 5; i++) { prm_.Wcost[i] = prm.Wcost[i]; Wcost[i] = prm.Wcost[i]; }  // rrt/src/rrtplanner.cpp:14-18
  for (int i = 0; i < 4; i++) prm_.goal[i] = goalPose[i];
  prm_.obs_use_pred = prm.obs_use_pred ? 1 : 0;
  int rc = clrrt_create(&prm_, device, tree_capacity, max_round, nullptr, &ctx_);
  if (rc != CLRRT_OK) {
    std::string msg = ctx_ ? clrrt_last_error(ctx_) : "no CUDA device";
    if (ctx_) clrrt_destroy(ctx_);
    ctx_ = nullptr;
    throw Error("clrrt_create failed (" + std::to_string(rc) + "): " + msg + " — there is no CPU fallback");
  }
}
MyRRT::~MyRRT() {
  if (ctx_) clrrt_destroy(ctx_);
}
void MyRRT::addInitialNode(const std::vector<double>& state) {
  // rrt/src/rrtplanner.cpp:21-37: reference (0,0) -> (1,0) with N = floor(1/0.1) points, v = state[4]
  clrrt_node n;
  memset(&n, 0, sizeof n);
  for (int k = 0; k < 10 && k < (int)state.size(); k++) n.state[k] = state[k];
  double xend{1}, yend{0}, res{0.1};
  int N = floor(sqrt(pow(xend, 2) + pow(yend, 2)) / res);
  std::vector<double> rx = LinearSpacedVector(0, xend, N), ry = LinearSpacedVector(0, yend, N);
  n.ref_front[0] = rx.front(); n.ref_front[1] = ry.front(); n.ref_back[0] = rx.back(); n.ref_back[1] = ry.back();
  n.ref_vback = state[4];
  n.parent = -1; n.n_ref = N;
  ck(ctx_, clrrt_tree_reset(ctx_, &n, 1), "clrrt_tree_reset");
}
void MyRRT::setRoad() {
  // MotionRequest.bend / Cxy / laneShifts (rrt/src/motionplanner.cpp:23): the lane-deviation cost of simulation.cpp:92-95
  prm_.bend = bend ? 1 : 0;
  prm_.lane_shift = 0; prm_.Cxy[0] = prm_.Cxy[1] = prm_.Cxy[2] = 0;
  if (bend) {
    if (laneShifts.empty() || Cxy.size() < 3) throw Error("bend=true needs laneShifts[0] and Cxy[0..2]");
    prm_.lane_shift = laneShifts[0];
    for (int i = 0; i < 3; i++) prm_.Cxy[i] = Cxy[i];
  }
}
void MyRRT::reconfigure(const std::vector<double>& goal, double vmax, double car_speed, const PlannerParams& prm) {
  if (goal.size() < 4) throw Error("goalPose needs 4 entries");
  goalPose = goal;
  setRoad();
  prm_.ref_res = std::max(std::abs(car_speed) * prm.ref_int, prm.ref_mindist);  // controller.cpp:18-21
  prm_.vmax = vmax;
  for (int i = 0; i < 5; i++) { prm_.Wcost[i] = prm.Wcost[i]; Wcost[i] = prm.Wcost[i]; }
  for (int i = 0; i < 4; i++) prm_.goal[i] = goalPose[i];
  prm_.obs_use_pred = prm.obs_use_pred ? 1 : 0;
  carried_.clear();
  ck(ctx_, clrrt_set_params(ctx_, &prm_), "clrrt_set_params");
}
void MyRRT::setCarriedTree(const std::vector<Node>& nodes) {
  std::vector<clrrt_node> c(nodes.size());
  for (size_t i = 0; i < nodes.size(); i++) {
    const Node& nd = nodes[i];
    clrrt_node& n = c[i];
    memset(&n, 0, sizeof n);
    for (int k = 0; k < 10 && k < (int)nd.state.size(); k++) n.state[k] = nd.state[k];
    if (nd.ref.x.empty() || nd.ref.v.empty()) throw Error("carried node without a reference");
    n.ref_front[0] = nd.ref.x.front(); n.ref_front[1] = nd.ref.y.front();
    n.ref_back[0] = nd.ref.x.back(); n.ref_back[1] = nd.ref.y.back();
    n.ref_vback = nd.ref.v.back();
    n.costE = nd.costE; n.costS = nd.costS; n.parent = nd.parentID; n.goal_reached = nd.goalReached ? 1 : 0;
    n.n_ref = (int)nd.ref.x.size(); n.kind = 0;
  }
  ck(ctx_, clrrt_tree_reset(ctx_, c.data(), (int)c.size()), "clrrt_tree_reset");
  carried_ = nodes;
}
void MyRRT::setObstacles(const std::vector<Obstacle2D>& d) {
  det = d;
  std::vector<clrrt_obstacle> o(d.size());
  for (size_t i = 0; i < d.size(); i++)
    o[i] = {d[i].obb.center.x, d[i].obb.center.y, d[i].obb.center.theta, d[i].obb.size_x, d[i].obb.size_y,
            d[i].vel.linear.x, d[i].vel.linear.y};
  ck(ctx_, clrrt_set_obstacles(ctx_, o.empty() ? nullptr : o.data(), (int)o.size()), "clrrt_set_obstacles");
}
int MyRRT::treeSize() const { return clrrt_tree_size(ctx_); }
std::vector<Node> MyRRT::tree() const {
  const int n = treeSize();
  std::vector<clrrt_node> c((size_t)n);
  int got = 0;
  ck(ctx_, clrrt_tree_download(ctx_, c.data(), n, &got), "clrrt_tree_download");
  std::vector<Node> out;
  out.reserve(got);
  for (int i = 0; i < got; i++) out.push_back(node_from_c(c[i]));
  return out;
}

double getNodeCost(const MyRRT& RRT, const Vehicle& veh, const double& parentCost, const Node& node, double sim_dt) {
  // rrt/src/rrtplanner.cpp:105-119.  checkObsDistance(RRT.carState) is the shipped stub there (collisioncheck.cpp:6-8),
  // also in a build with obstacles: carState has 6 entries at that point, which the obstacle overload rejects
  double cost = parentCost;
  for (auto it = node.tra.begin(); it != node.tra.end(); it++) {
    double Dobs = 100;
    double kappa = tan((*it)[3]) / veh.L;
    cost += RRT.Wcost[0] * (*it)[4] * sim_dt + RRT.Wcost[1] * std::abs(kappa) + RRT.Wcost[2] * exp(-RRT.Wcost[3] * Dobs);
    if (RRT.bend) {  // d2L, rrt/src/rrtplanner.cpp:98-102
      const double x = (*it)[0], y = (*it)[1], S = RRT.laneShifts[0];
      const std::vector<double>& Cxy = RRT.Cxy;
      double Lx = (x - S * Cxy[1] + y * Cxy[1] - Cxy[1] * Cxy[2]) / (pow(Cxy[1], 2) + 1);
      double Ly = S + Cxy[2] + (Cxy[1] * (x - S * Cxy[1] + y * Cxy[1] - Cxy[1] * Cxy[2])) / (pow(Cxy[1], 2) + 1);
      cost += RRT.Wcost[4] * sqrt(pow(Lx - x, 2) + pow(Ly - y, 2));
    }
  }
  return cost;
}

void initializeTree(MyRRT& RRT, const Vehicle& veh, std::vector<Node>& nodes, std::vector<double>& carState) {
  // rrt/src/rrtplanner.cpp:39-48: four logging slots appended, empty committed path -> root node
  carState.push_back(0); carState.push_back(0); carState.push_back(0); carState.push_back(0);
  if (nodes.size() == 0) {
    RRT.addInitialNode(carState);
    return;
  }
  // :51-57 erase nodes that end behind the vehicle (upstream erases through `nodes.erase(it--)`: same survivors)
  for (auto it = nodes.begin(); it != nodes.end();) {
    it->goalReached = 0;
    if (it->tra.empty()) throw Error("carried node without a trajectory");
    if ((it->tra.back()[0]) < 0) it = nodes.erase(it);
    else ++it;
  }
  if (nodes.empty()) {  // upstream reads nodes.front() of an empty vector here (undefined); defined as the empty tree
    RRT.addInitialNode(carState);
    return;
  }
  // :59-70 goal flags from the trajectories (the y distance ignores the goal's y, as upstream)
  for (auto it = nodes.begin(); it != nodes.end(); ++it) {
    for (size_t i = 0; i != it->tra.size(); i++) {
      double Dgoal = sqrt(pow(it->tra[i][0] - RRT.goalPose[0], 2) + pow(it->tra[i][1], 2));
      double Hgoal = std::abs(it->tra[i][2] - RRT.goalPose[2]);
      double dVgoal = std::abs(it->tra[i][4] - RRT.goalPose[3]);
      if ((Dgoal <= 1) && (Hgoal <= 0.05) && (dVgoal <= 0.1)) it->goalReached = 1;
    }
  }
  // :72-82 the collision re-check calls the shipped stub (Dobs = 100): it never fires
  // :84-88 cost estimates along the chain (Node::costS is float)
  const double dt = RRT.params().sim_dt;
  nodes.front().costS = getNodeCost(RRT, veh, 0, nodes.front(), dt);
  for (size_t i = 1; i != nodes.size(); i++) nodes[i].costS = getNodeCost(RRT, veh, nodes[i - 1].costS, nodes[i], dt);
  // :90-93 re-parent as a chain and make it the tree
  for (size_t i = 0; i != nodes.size(); i++) nodes[i].parentID = (int)i - 1;
  RRT.setCarriedTree(nodes);
}

clrrt_round_stats expandTree(Vehicle&, MyRRT& RRT, int K) {
  std::vector<double> s(2 * (size_t)K);
  std::vector<uint8_t> h((size_t)K);
  clrrt_draw_samples(RRT.goalPose.data(), K, s.data(), h.data());  // rrtplanner.cpp:133-143 on rand()
  clrrt_round_stats st;
  const int rc = clrrt_expand_round(RRT.ctx(), s.data(), h.data(), K, &st);
  // a full tree is not an error of the query: the round's prefix that fits was appended; the caller stops expanding
  if (rc == CLRRT_ERR_CAPACITY) { RRT.treeFull = true; return st; }
  ck(RRT.ctx(), rc, "clrrt_expand_round");
  return st;
}

// n consecutive expandTree calls (the loop body of rrt/src/motionplanner.cpp:39-43 n times): same draws, same tree, same
// counters as n calls of expandTree(veh, RRT, 1), with several samples in flight on the device (clrrt_expand_sequential)
clrrt_seq_stats expandTreeSequential(Vehicle&, MyRRT& RRT, int n, int window) {
  std::vector<double> s(2 * (size_t)n);
  std::vector<uint8_t> h((size_t)n);
  clrrt_draw_samples(RRT.goalPose.data(), n, s.data(), h.data());  // 3 rand() per iteration, in the reference's order
  clrrt_seq_stats st;
  const int rc = clrrt_expand_sequential(RRT.ctx(), s.data(), h.data(), n, window, &st);
  if (rc == CLRRT_ERR_CAPACITY) { RRT.treeFull = true; return st; }
  ck(RRT.ctx(), rc, "clrrt_expand_sequential");
  return st;
}

namespace {
// one rollout with trajectory and reference dumps -> stateArray + MyReference
void rematerialise(const MyRRT& RRT, int parent, const double sxy[2], bool gb, clrrt_rollout& r, StateArray& states,
                   MyReference& ref) {
  const int tstride = CLRRT_MAX_STEPS_CAP, rstride = 2048;
  std::vector<double> traj((size_t)tstride * 10), rxyv((size_t)rstride * 3);
  const uint8_t g = gb ? 1 : 0;
  const int32_t p = parent;
  ck(RRT.ctx(), clrrt_propagate_batch_ex(RRT.ctx(), &p, sxy, &g, 1, &r, traj.data(), tstride, rxyv.data(), rstride),
     "clrrt_propagate_batch_ex");
  states.clear();
  for (int i = 0; i <= r.n_steps && i < tstride; i++) states.emplace_back(traj.begin() + 10 * i, traj.begin() + 10 * (i + 1));
  ref = MyReference();
  for (int i = 0; i < r.n_ref && i < rstride; i++) {
    ref.x.push_back(rxyv[3 * i]); ref.y.push_back(rxyv[3 * i + 1]); ref.v.push_back(rxyv[3 * i + 2]);
  }
  ref.dir = 1;
}
}  // namespace

Simulation::Simulation(const MyRRT& RRT, int parent, const Point& sample, const Vehicle&, const bool& GoalBiased) {
  clrrt_rollout r;
  const double sxy[2] = {sample.x, sample.y};
  rematerialise(RRT, parent, sxy, GoalBiased, r, stateArray, ref);
  costE = r.costE; costS = r.costS; goalReached = r.goal_reached != 0; endReached = r.end_reached != 0; failCode = r.fail;
}

// the reference's own parameter list (rrt/include/rrt/simulation.h:18-19, rrt/src/simulation.cpp:36-47)
Simulation::Simulation(const MyRRT& RRT, const std::vector<double>& state, MyReference& ref, const Vehicle&, const bool& GoalBiased,
                       const bool& genProfile, const double& Vstart) {
  if (state.size() < 6) throw Error("Simulation: state needs at least [x, y, theta, delta, v, a]");
  if (ref.x.size() < 3 || ref.y.size() != ref.x.size()) throw Error("Simulation: the reference needs at least 3 points (reference.cpp:19)");
  double st[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  for (size_t k = 0; k < 10 && k < state.size(); k++) st[k] = state[k];
  const int N = (int)ref.x.size(), tstride = CLRRT_MAX_STEPS_CAP;
  std::vector<double> v((size_t)N, 0.0), traj((size_t)tstride * 10);
  clrrt_rollout r;
  bool gen = genProfile;
  if (genProfile && !ref.v.empty()) {
    // upstream generateVelocityProfile push_backs onto ref.v (reference.cpp:134-146): with a non-empty ref.v the rollout
    // reads the OLD leading entries.  No call site does this; reproduce it anyway: profile first, then simulate on the
    // concatenation's first N entries.
    ck(RRT.ctx(), clrrt_simulate(RRT.ctx(), st, ref.x.data(), ref.y.data(), v.data(), N, ref.dir, GoalBiased ? 1 : 0, 1, Vstart, &r, nullptr, 0),
       "clrrt_simulate");
    ref.v.insert(ref.v.end(), v.begin(), v.end());
    for (int i = 0; i < N; i++) v[(size_t)i] = ref.v[(size_t)i];
    gen = false;
  } else if (!genProfile) {
    if ((int)ref.v.size() < N) throw Error("Simulation: genProfile == false needs ref.v of the reference's length");
    for (int i = 0; i < N; i++) v[(size_t)i] = ref.v[(size_t)i];
  }
  ck(RRT.ctx(), clrrt_simulate(RRT.ctx(), st, ref.x.data(), ref.y.data(), v.data(), N, ref.dir, GoalBiased ? 1 : 0, gen ? 1 : 0, Vstart, &r,
                               traj.data(), tstride), "clrrt_simulate");
  if (gen) ref.v.assign(v.begin(), v.end());  // the constructor fills the caller's reference (simulation.cpp:43)
  stateArray.clear();
  for (int i = 0; i <= r.n_steps && i < tstride; i++) stateArray.emplace_back(traj.begin() + 10 * i, traj.begin() + 10 * (i + 1));
  this->ref = ref;
  costE = r.costE; costS = r.costS; goalReached = r.goal_reached != 0; endReached = r.end_reached != 0; failCode = r.fail;
}

std::vector<Node> extractBestPath(MyRRT& RRT, std::vector<int32_t>* ids_out) {
  const int n = RRT.treeSize();
  std::vector<int32_t> ids((size_t)n);
  int len = 0;
  ck(RRT.ctx(), clrrt_best_path(RRT.ctx(), ids.data(), n, &len), "clrrt_best_path");
  std::vector<Node> best;
  if (ids_out) ids_out->assign(ids.begin(), ids.begin() + len);
  if (len == 0) return best;  // "No solution was found!", rrtplanner.cpp:342-344
  std::vector<clrrt_node> c((size_t)len);
  for (int i = 0; i < len; i++) ck(RRT.ctx(), clrrt_tree_download_range(RRT.ctx(), ids[i], 1, &c[i]), "download");
  for (int i = 0; i < len; i++) best.push_back(node_from_c(c[i]));
  // Node::tra / Node::ref: the device tree keeps end points only.  Every node records the sample its reference was
  // aimed at and whether it was a goal-biased expansion, which with its parent re-creates the rollout exactly.
  const std::vector<Node>& carried = RRT.carried();
  for (int i = 0; i < len; i++) {
    if (ids[i] < (int)carried.size()) { best[i] = carried[ids[i]]; continue; }  // carried node: the host copy is complete
    if (c[i].kind == 0) {
      // root: addInitialNode stores the single start state and the 10-point reference (0,0)->(1,0), rrtplanner.cpp:21-37
      best[i].tra = {best[i].state};
      const int N = std::max(c[i].n_ref, 2);
      best[i].ref.x = LinearSpacedVector(0, 1, N);
      best[i].ref.y = LinearSpacedVector(0, 0, N);
      best[i].ref.v.assign((size_t)N, best[i].state[4]);
      best[i].ref.dir = 1;
      continue;
    }
    clrrt_rollout r;
    rematerialise(RRT, ids[i - 1], c[i].sample, c[i].kind == 2, r, best[i].tra, best[i].ref);
  }
  return best;
}

std::vector<Path> convertNodesToPath(const std::vector<Node>& path) {
  std::vector<Path> result;
  for (auto it = path.begin(); it != path.end(); ++it) {
    Path segment;
    segment.ref = it->ref;
    segment.tra = it->tra;
    result.push_back(segment);
  }
  return result;
}
Trajectory generateMPCmessage(const std::vector<Path>& path) {
  Trajectory tra;
  for (auto it = path.begin(); it != path.end(); ++it) {
    for (size_t i = 1; i < it->tra.size(); i++) {
      tra.x.push_back(it->tra[i][0]); tra.y.push_back(it->tra[i][1]); tra.theta.push_back(it->tra[i][2]);
      tra.delta.push_back(it->tra[i][3]); tra.v.push_back(it->tra[i][4]); tra.a.push_back(it->tra[i][5]);
      tra.a_cmd.push_back(it->tra[i][8]); tra.d_cmd.push_back(it->tra[i][9]);
    }
  }
  return tra;
}
void filterMPCmessage(Trajectory& msg) {
  Trajectory f;
  double interval = 5, d = 0;
  for (size_t i = 1; i < msg.x.size(); i++) {
    if (d == 0) {
      f.x.push_back(msg.x[i]); f.y.push_back(msg.y[i]); f.theta.push_back(msg.theta[i]); f.v.push_back(msg.v[i]);
      f.a.push_back(msg.a[i]); f.a_cmd.push_back(msg.a_cmd[i]); f.d_cmd.push_back(msg.d_cmd[i]);
    }
    d += sqrt(pow(msg.x[i] - msg.x[i - 1], 2) + pow(msg.y[i] - msg.y[i - 1], 2));
    if (d >= interval) d = 0;
  }
  msg = f;
}
void transformNodesWorldToCar(std::vector<Node>& nodes, const std::vector<double> carPose) {
  auto pt = [&](double& Xw, double& Yw) {  // transformPointWorldToCar, rrt/src/transformations.cpp:6-10
    double Xc = Xw * cos(carPose[2]) - carPose[0] * cos(carPose[2]) - carPose[1] * sin(carPose[2]) + Yw * sin(carPose[2]);
    double Yc = Yw * cos(carPose[2]) - carPose[1] * cos(carPose[2]) + carPose[0] * sin(carPose[2]) - Xw * sin(carPose[2]);
    Xw = Xc; Yw = Yc;
  };
  for (auto& n : nodes) {
    pt(n.state[0], n.state[1]);
    n.state[2] -= carPose[2];
    for (size_t i = 0; i < n.ref.x.size(); i++) pt(n.ref.x[i], n.ref.y[i]);
    for (auto& s : n.tra) pt(s[0], s[1]);  // trajectory headings are left as they are, as upstream (:310-313)
  }
}
void transformNodesCarToworld(std::vector<Node>& nodes, const std::vector<double> carPose) {
  auto pt = [&](double& Xc, double& Yc) {  // transformPointCarToWorld, rrt/src/transformations.cpp:13-17
    double Xw = cos(carPose[2]) * Xc - sin(carPose[2]) * Yc + carPose[0];
    double Yw = sin(carPose[2]) * Xc + cos(carPose[2]) * Yc + carPose[1];
    Xc = Xw; Yc = Yw;
  };
  for (auto& n : nodes) {
    pt(n.state[0], n.state[1]);
    n.state[2] += carPose[2];
    for (size_t i = 0; i < n.ref.x.size(); i++) pt(n.ref.x[i], n.ref.y[i]);
    for (auto& s : n.tra) pt(s[0], s[1]);  // headings of the trajectory are left as they are, as upstream (:310-313)
  }
}

bool MotionPlanner::updateObstacles() {
  if (getobstacles) det = getobstacles();
  return true;  // upstream falls off the end of a bool function (rrt/src/motionplanner.cpp:81-86)
}
void MotionPlanner::updateState(const std::vector<double>& msg_state) {
  state.clear();
  state.insert(state.begin(), msg_state.begin(), msg_state.end());
  if (state.size() != 6) throw Error("state must have 6 entries [x,y,theta,delta,v,a]");
}
bool MotionPlanner::resetPlanner() {
  motionplan.clear();
  return true;
}
MotionPlanner::~MotionPlanner() { delete rrt_; }
void MotionPlanner::planMotion(MotionRequest req) {
  auto tic = std::chrono::steady_clock::now();
  auto lap = [&](int k) { const auto now = std::chrono::steady_clock::now(); lastMs[k] = std::chrono::duration<double, std::milli>(now - tic).count(); tic = now; };
  Vehicle veh; veh.setPrius();                                  // :13
  std::vector<double> worldState = state;                       // :14
  std::vector<double> carPose = worldState;                     // transformStateToLocal, transformations.cpp:143-147
  carPose[0] = 0; carPose[1] = 0; carPose[2] = 0;
  updateObstacles();                                            // :18
  transformNodesWorldToCar(bestNodes, worldState);              // :22
  // :16-17, :23 — the device context is kept across queries; a query only changes its parameters
  const int round = std::max(samplesPerRound, 1);
  if (rrt_ && rrt_round_ != round) { delete rrt_; rrt_ = nullptr; }
  if (!rrt_) {
    // (the sequential mode keeps up to 64 samples in flight: its scratch is sized for that, not for one sample)
    rrt_ = new MyRRT(req.goal, req.laneShifts, req.Cxy, req.bend, veh, params, req.vmax, carPose[4], device, treeCapacity, std::max(round, 64));
    rrt_round_ = round;
  } else {
    rrt_->bend = req.bend; rrt_->laneShifts = req.laneShifts; rrt_->Cxy = req.Cxy;
    rrt_->reconfigure(req.goal, req.vmax, carPose[4], params);
  }
  MyRRT& RRT = *rrt_;
  lap(0);
  RRT.setObstacles(det);                                        // :24
  lap(1);
  RRT.carState = carPose;
  if (!params.commit_path) bestNodes.clear();                   // :28-30
  RRT.treeFull = false;
  lastSeqWindows = lastSeqSpeculated = 0;
  initializeTree(RRT, veh, bestNodes, carPose);                 // :32
  lastCarried = (int)RRT.carried().size();
  lastInitialTree = RRT.treeSize();
  lap(2);
  int iter = 0;                                                 // :39-43
  auto t0 = std::chrono::steady_clock::now();
  if (round == 1 && sequentialChunk > 1) {
    // the reference's sequential loop, `sequentialChunk` iterations per call with several of them in flight on the device;
    // the timer is polled between calls (upstream: between iterations)
    for (;;) {
      int n = sequentialChunk;
      if (maxIterations >= 0) { n = std::min(n, maxIterations - iter); if (n <= 0) break; }
      else if (std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count() >= budget_ms) break;
      const clrrt_seq_stats ss = expandTreeSequential(veh, RRT, n, sequentialWindow);
      lastSeqWindows += ss.windows; lastSeqSpeculated += ss.speculated;
      iter += n;
      if (RRT.treeFull) break;
    }
  } else {
    for (;; iter++) {
      if (maxIterations >= 0) { if (iter >= maxIterations) break; }
      else if (std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count() >= budget_ms) break;
      expandTree(veh, RRT, round);
      if (RRT.treeFull) { iter++; break; }  // capacity reached: keep what was grown, extract the best path below
    }
  }
  lastIterations = iter;
  lastTreeSize = RRT.treeSize();
  lap(3);
  clrrt_counters_get(RRT.ctx(), &lastCounters);                 // :45
  bestNodes = extractBestPath(RRT, &lastBestIds);               // :51
  lap(4);
  lastRematError = 0;
  for (size_t i = 0; i < bestNodes.size(); i++) {
    const Node& n = bestNodes[i];
    if (lastBestIds[i] < lastCarried || n.tra.empty()) continue;  // carried nodes were not re-created
    for (int k = 0; k < 10; k++) lastRematError = std::max(lastRematError, std::abs(n.tra.back()[k] - n.state[k]));
  }
  transformNodesCarToworld(bestNodes, worldState);              // :54
  lastTrajectory = Trajectory();
  if (bestNodes.size() == 0) return;                            // :56-58
  std::vector<Path> plan = convertNodesToPath(bestNodes);       // :66
  Trajectory msg = generateMPCmessage(plan);                    // :69
  filterMPCmessage(msg);                                        // :70
  lastTrajectory = msg;
  if (msg.x.size() >= 3 && pubMPC) pubMPC(msg);                 // :71-74
  lap(5);
}

}  // namespace clrrt

// text of the last failure of a clrrt_host_* call on this thread (the facade throws clrrt::Error; the C view catches it)
static thread_local std::string g_host_error;
extern "C" const char* clrrt_host_last_error(void) { return g_host_error.c_str(); }
static int host_fail(const char* where, const std::exception& e) {
  g_host_error = std::string(where) + ": " + e.what();
  return CLRRT_ERR_STATE;
}

extern "C" int clrrt_host_plan_motion(const double* car_state6, const double* goal4, double vmax, const clrrt_obstacle* obs,
                                      int n_obs, int samples_per_round, int max_iterations, double budget_ms, unsigned seed,
                                      int device, int* tree_size, int* iterations, clrrt_counters* counters, double* traj8,
                                      int traj_cap, int* traj_len, int32_t* best_ids, int best_cap, int* best_len, double* remat_err) {
  if (!car_state6 || !goal4 || n_obs < 0 || (n_obs > 0 && !obs)) { g_host_error = "clrrt_host_plan_motion: null argument"; return CLRRT_ERR_ARG; }
  try {
    g_host_error.clear();
    clrrt::MotionPlanner mp;
    mp.device = device;
    mp.samplesPerRound = samples_per_round;
    mp.maxIterations = max_iterations;
    mp.budget_ms = budget_ms;
    std::vector<clrrt::Obstacle2D> det((size_t)n_obs);
    for (int i = 0; i < n_obs; i++) {
      det[i].obb.center.x = obs[i].cx; det[i].obb.center.y = obs[i].cy; det[i].obb.center.theta = obs[i].theta;
      det[i].obb.size_x = obs[i].size_x; det[i].obb.size_y = obs[i].size_y;
      det[i].vel.linear.x = obs[i].vx; det[i].vel.linear.y = obs[i].vy;
    }
    mp.getobstacles = [&]() { return det; };
    mp.updateState(std::vector<double>(car_state6, car_state6 + 6));
    clrrt::MotionRequest req;
    req.goal.assign(goal4, goal4 + 4);
    req.vmax = vmax;
    req.laneShifts = {0};
    srand(seed);  // the reference never seeds rand(): seed 1 reproduces it
    mp.planMotion(req);
    if (tree_size) *tree_size = mp.lastTreeSize;
    if (iterations) *iterations = mp.lastIterations;
    if (counters) *counters = mp.lastCounters;
    const clrrt::Trajectory& t = mp.lastTrajectory;
    const int n = (int)t.x.size();
    if (traj_len) *traj_len = n;
    for (int i = 0; i < n && i < traj_cap && traj8; i++) {
      double* o = traj8 + 8 * i;
      o[0] = t.x[i]; o[1] = t.y[i]; o[2] = t.theta[i]; o[3] = 0; o[4] = t.v[i]; o[5] = t.a[i]; o[6] = t.a_cmd[i]; o[7] = t.d_cmd[i];
    }
    if (best_len) *best_len = (int)mp.lastBestIds.size();
    for (int i = 0; i < (int)mp.lastBestIds.size() && i < best_cap && best_ids; i++) best_ids[i] = mp.lastBestIds[i];
    if (remat_err) *remat_err = mp.lastRematError;
    return CLRRT_OK;
  } catch (const std::exception& e) {
    return host_fail("clrrt_host_plan_motion", e);
  }
}

// The facade's Simulation with the reference's parameter list, for bindings/tests: returns stateArray.size(); ref_v is
// filled when gen_profile (the constructor mutates the caller's MyReference); flags3 = {endReached, goalReached, failCode}
extern "C" int clrrt_host_simulate(const double* goal4, double vmax, const clrrt_obstacle* obs, int n_obs, int device,
                                   const double* state, int n_state, const double* ref_x, const double* ref_y, double* ref_v, int n_ref,
                                   int goal_biased, int gen_profile, double Vstart, double* costs2, int32_t* flags3, double* last_state10) {
  if (!goal4 || !state || !ref_x || !ref_y || !ref_v || n_ref < 0 || n_state < 0) { g_host_error = "clrrt_host_simulate: null argument"; return CLRRT_ERR_ARG; }
  try {
    g_host_error.clear();
    clrrt::Vehicle veh; veh.setPrius();
    clrrt::PlannerParams prm;
    std::vector<double> goal(goal4, goal4 + 4);
    clrrt::MyRRT RRT(goal, {0}, {}, false, veh, prm, vmax, n_state > 4 ? state[4] : 0.0, device, 64, 64);
    std::vector<clrrt::Obstacle2D> det((size_t)n_obs);
    for (int i = 0; i < n_obs; i++) {
      det[i].obb.center.x = obs[i].cx; det[i].obb.center.y = obs[i].cy; det[i].obb.center.theta = obs[i].theta;
      det[i].obb.size_x = obs[i].size_x; det[i].obb.size_y = obs[i].size_y;
      det[i].vel.linear.x = obs[i].vx; det[i].vel.linear.y = obs[i].vy;
    }
    RRT.setObstacles(det);
    clrrt::MyReference ref;
    ref.x.assign(ref_x, ref_x + n_ref); ref.y.assign(ref_y, ref_y + n_ref);
    if (!gen_profile) ref.v.assign(ref_v, ref_v + n_ref);
    clrrt::Simulation sim(RRT, std::vector<double>(state, state + n_state), ref, veh, goal_biased != 0, gen_profile != 0, Vstart);
    for (int i = 0; i < n_ref && i < (int)ref.v.size(); i++) ref_v[i] = ref.v[(size_t)i];
    if (costs2) { costs2[0] = sim.costE; costs2[1] = sim.costS; }
    if (flags3) { flags3[0] = sim.endReached; flags3[1] = sim.goalReached; flags3[2] = sim.failCode; }
    if (last_state10) for (int k = 0; k < 10; k++) last_state10[k] = sim.stateArray.back()[(size_t)k];
    return (int)sim.stateArray.size();
  } catch (const std::exception& e) {
    return host_fail("clrrt_host_simulate", e);
  }
}

// ---- persistent planner handle: consecutive queries (receding-horizon loop, config C5) ----------------------------
// wall-clock milliseconds of the last query's phases: parameters, obstacles, initial tree, expansion, best path, messages
extern "C" int clrrt_host_planner_timings(void* h, double* ms6) {
  if (!h || !ms6) return CLRRT_ERR_ARG;
  for (int k = 0; k < 6; k++) ms6[k] = static_cast<clrrt::MotionPlanner*>(h)->lastMs[k];
  return static_cast<clrrt::MotionPlanner*>(h)->lastSeqWindows;  // >= 0: speculative windows the last query's expansion used
}
extern "C" int clrrt_host_planner_set_sequential(void* h, int chunk, int window) {
  if (!h || chunk < 1 || window < 0 || window > 64) return CLRRT_ERR_ARG;
  clrrt::MotionPlanner* mp = static_cast<clrrt::MotionPlanner*>(h);
  mp->sequentialChunk = chunk; mp->sequentialWindow = window;
  return CLRRT_OK;
}
extern "C" void* clrrt_host_planner_create(int device, int samples_per_round, int commit_path, int tree_capacity) {
  clrrt::MotionPlanner* mp = new clrrt::MotionPlanner();
  mp->device = device;
  mp->samplesPerRound = samples_per_round;
  mp->params.commit_path = commit_path != 0;
  if (tree_capacity > 0) mp->treeCapacity = tree_capacity;
  return mp;
}
extern "C" void clrrt_host_planner_destroy(void* h) { delete static_cast<clrrt::MotionPlanner*>(h); }
// One planMotion call.  goal4 and obs are in the car frame, as the planner receives them.  sizes4 = {nodes of
// the initial tree (1 = root only), final tree size, nodes of the best path, expandTree calls}.
extern "C" int clrrt_host_planner_query(void* h, const double* world_state6, const double* goal4, double vmax,
                                        const clrrt_obstacle* obs, int n_obs, int max_iterations, double budget_ms,
                                        int32_t* sizes4, double* best_cost, clrrt_counters* counters) {
  if (!h || !world_state6 || !goal4 || n_obs < 0 || (n_obs > 0 && !obs)) { g_host_error = "clrrt_host_planner_query: null argument"; return CLRRT_ERR_ARG; }
  try {
    g_host_error.clear();
    clrrt::MotionPlanner& mp = *static_cast<clrrt::MotionPlanner*>(h);
    mp.maxIterations = max_iterations;
    mp.budget_ms = budget_ms;
    std::vector<clrrt::Obstacle2D> det((size_t)n_obs);
    for (int i = 0; i < n_obs; i++) {
      det[i].obb.center.x = obs[i].cx; det[i].obb.center.y = obs[i].cy; det[i].obb.center.theta = obs[i].theta;
      det[i].obb.size_x = obs[i].size_x; det[i].obb.size_y = obs[i].size_y;
      det[i].vel.linear.x = obs[i].vx; det[i].vel.linear.y = obs[i].vy;
    }
    mp.getobstacles = [det]() { return det; };
    mp.updateState(std::vector<double>(world_state6, world_state6 + 6));
    clrrt::MotionRequest req;
    req.goal.assign(goal4, goal4 + 4);
    req.vmax = vmax;
    req.laneShifts = {0};
    mp.planMotion(req);
    if (sizes4) {
      sizes4[0] = mp.lastInitialTree; sizes4[1] = mp.lastTreeSize; sizes4[2] = (int)mp.bestNodes.size(); sizes4[3] = mp.lastIterations;
    }
    if (best_cost) *best_cost = mp.bestNodes.empty() ? -1.0 : (double)mp.bestNodes.back().costS;
    if (counters) *counters = mp.lastCounters;
    return CLRRT_OK;
  } catch (const std::exception& e) {
    return host_fail("clrrt_host_planner_query", e);
  }
}
// bestNodes of the last query (world frame): {state[10], ref front xy, ref back xy, ref.v.back(), costE, costS,
// parentID, goalReached, ref.x.size()} per node; returns the node count
extern "C" int clrrt_host_planner_best_nodes(void* h, double* rec20, int cap) {
  const clrrt::MotionPlanner& mp = *static_cast<clrrt::MotionPlanner*>(h);
  int n = 0;
  for (const clrrt::Node& nd : mp.bestNodes) {
    if (n >= cap) break;
    double* o = rec20 + (size_t)20 * n++;
    for (int k = 0; k < 10; k++) o[k] = nd.state[k];
    o[10] = nd.ref.x.front(); o[11] = nd.ref.y.front(); o[12] = nd.ref.x.back(); o[13] = nd.ref.y.back();
    o[14] = nd.ref.v.empty() ? 0.0 : nd.ref.v.back();
    o[15] = nd.costE; o[16] = nd.costS; o[17] = nd.parentID; o[18] = nd.goalReached ? 1 : 0; o[19] = (double)nd.ref.x.size();
  }
  return (int)mp.bestNodes.size();
}
extern "C" int clrrt_host_planner_best_traj(void* h, double* traj10, int cap_rows, int32_t* rows_per_node, int cap_nodes) {
  const clrrt::MotionPlanner& mp = *static_cast<clrrt::MotionPlanner*>(h);
  int rows = 0, n = 0;
  for (const clrrt::Node& nd : mp.bestNodes) {
    if (n < cap_nodes) rows_per_node[n] = (int)nd.tra.size();
    n++;
    for (const auto& x : nd.tra) {
      if (rows < cap_rows) for (int k = 0; k < 10; k++) traj10[(size_t)10 * rows + k] = x[k];
      rows++;
    }
  }
  return rows;
}
