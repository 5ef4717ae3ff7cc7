// clrrt_road.cpp — see clrrt_road.hpp.
#include "clrrt_road.hpp"

#include <cmath>
#include <cstdlib>

namespace clrrt {

namespace {
const double kPi = M_PI;
// wrapTo2Pi, rrt/include/rrt/functions.h:35-40
double wrapTo2Pi(double x) {
  x = fmod(x, 2 * kPi);
  if (x < 0) x += 2 * kPi;
  return x;
}
// the straightened road: the line through (0, c0) with slope c1, as a rotation by atan2(c1, 1) and a shift
struct StraightRoad {
  double c, s, c0;
  StraightRoad(const std::vector<double>& Cxy) : c0(Cxy[2]) { const double th = atan2(Cxy[1], 1); c = cos(th); s = sin(th); }
  void place(double S, double rho, double& x, double& y) const { x = c * S - s * rho; y = s * S + c * rho + c0; }
};
}  // namespace

RoadFrame::RoadFrame(const std::vector<double>& Cxy, const std::vector<double>& Cxs) : Cxy_(Cxy), Cxs_(Cxs) {
  if (Cxy_.size() < 3 || Cxs_.size() < 3) throw Error("RoadFrame needs Cxy[0..2] and Cxs[0..2]");
}

void RoadFrame::closestPointOnArc(double Xcar, double Ycar, double& Xarc, double& Yarc) const {
  const std::vector<double>& Cxy = Cxy_;
  // closed-form root of the cubic d/dx |P - road(x)|^2 = 0 (upstream: "solved symbolically in MATLAB"); the temporaries are
  // float upstream and stay float here
  float t2 = std::abs(Cxy[0]);
  float t3 = pow(Cxy[0], 3);
  float t4 = pow(Cxy[1], 2);
  float t5 = Xcar * Cxy[0] * 2.0;
  float t6 = Ycar * Cxy[0] * 4.0;
  float t8 = Cxy[0] * Cxy[2] * 4.0;
  float t11 = sqrt(3.0);
  float t7 = pow(t2, 3);
  float t9 = 1.0 / t3;
  float t10 = -t8;
  float t12 = Cxy[1] + t5;
  float t13 = pow(t12, 2);
  float t14 = Cxy[1] * t7 * 9.0;
  float t15 = Xcar * Cxy[0] * t7 * 18;
  float t17 = t4 + t6 + t10 - 2.0;
  float t16 = t13 * 27;
  float t18 = pow(t17, 3);
  float t19 = -t18;
  float t20 = t16 + t19;
  float t21 = sqrt(t20);
  float t22 = t3 * t11 * t21;
  float t23 = t14 + t15 + t22;
  float t24 = t9 * t23;
  Xarc = (pow(t24, 1.0 / 3.0) * 0.2403749283845681) / t2 - Cxy[1] / (Cxy[0] * 2.0) +
         1.0 / pow(Cxy[0], 2) * t2 * t17 * 1.0 / pow(t24, 1.0 / 3.0) * 0.3466806371753173;
  Yarc = Cxy[0] * pow(Xarc, 2) + Cxy[1] * Xarc + Cxy[2];
}

void RoadFrame::pointCarToRoad(double& Xcar, double& Ycar) const {
  double Xarc, Yarc;
  closestPointOnArc(Xcar, Ycar, Xarc, Yarc);
  // (S, rho): arc length of the foot point, signed distance from it (sign by the tangent's half-plane)
  const double S = Cxs_[0] * pow(Xarc, 2) + Cxs_[1] * Xarc + Cxs_[2];
  double rho = sqrt(pow(Xarc - Xcar, 2) + pow(Yarc - Ycar, 2));
  if (Ycar < (Yarc - Xarc * (Cxy_[1] + 2 * Xarc * Cxy_[0]) + Xcar * (Cxy_[1] + 2 * Xarc * Cxy_[0]))) rho = -rho;
  const double theta = atan2(Cxy_[1], 1);
  const double Xstraight = cos(theta) * S - sin(theta) * rho;
  const double Ystraight = sin(theta) * S + cos(theta) * rho + Cxy_[2];
  Xcar = Xstraight; Ycar = Ystraight;
}

void RoadFrame::pointRoadToCar(double& Xstraight, double& Ystraight) const {
  const std::vector<double>& Cxy = Cxy_;
  const std::vector<double>& Cxs = Cxs_;
  // foot point on the straightened road, (S, rho) from it
  const double Xroads = (Xstraight - Cxy[2] * Cxy[1] + Cxy[1] * Ystraight) / (pow(Cxy[1], 2) + 1);
  const double Yroads = Cxy[2] + (Cxy[1] * (Xstraight - Cxy[2] * Cxy[1] + Cxy[1] * Ystraight)) / (pow(Cxy[1], 2) + 1);
  const double S = sqrt(pow(Xroads, 2) + pow(Yroads - Cxy[2], 2));
  double rho = sqrt(pow(Xroads - Xstraight, 2) + pow(Yroads - Ystraight, 2));
  if (Ystraight < (Cxy[1] * Xstraight + Cxy[2])) rho = -rho;
  // back onto the arc: invert the arc-length polynomial, then along the normal
  const double Xarc = -(Cxs[1] - sqrt(pow(Cxs[1], 2) - 4 * Cxs[0] * Cxs[2] + 4 * Cxs[0] * S)) / (2 * Cxs[0]);
  const double Yarc = Cxy[0] * pow(Xarc, 2) + Cxy[1] * Xarc + Cxy[2];
  const double dydx = 2 * Cxy[0] * Xarc + Cxy[1];
  const double vx = 1, vy = dydx;
  const double L = sqrt(pow(vx, 2) + pow(vy, 2));
  const double nx = -(1 / L) * dydx, ny = (1 / L);
  Xstraight = Xarc + nx * rho; Ystraight = Yarc + ny * rho;
}

void RoadFrame::poseCarToRoad(double& Xcar, double& Ycar, double& Hcar) const {
  double Xarc, Yarc;
  closestPointOnArc(Xcar, Ycar, Xarc, Yarc);
  const double S = Cxs_[0] * pow(Xarc, 2) + Cxs_[1] * Xarc + Cxs_[2];
  double rho = sqrt(pow(Xarc - Xcar, 2) + pow(Yarc - Ycar, 2));
  const double dydx = 2 * Cxy_[0] * Xarc + Cxy_[1];
  if (Ycar < (dydx * Xcar + Yarc - dydx * Xarc)) rho = -rho;
  const double theta = atan2(Cxy_[1], 1);   // heading of the straightened road
  const double Hroad = atan2(dydx, 1);      // heading of the road at the foot point
  const double Hstraight = wrapTo2Pi((Hcar - Hroad) + theta);
  const double Xstraight = cos(theta) * S - sin(theta) * rho;
  const double Ystraight = sin(theta) * S + cos(theta) * rho + Cxy_[2];
  Xcar = Xstraight; Ycar = Ystraight; Hcar = Hstraight;
}

void RoadFrame::poseRoadToCar(double& Xstraight, double& Ystraight, double& Hstraight) const {
  const std::vector<double>& Cxy = Cxy_;
  const std::vector<double>& Cxs = Cxs_;
  const double Xroads = (Xstraight - Cxy[2] * Cxy[1] + Cxy[1] * Ystraight) / (pow(Cxy[1], 2) + 1);
  const double Yroads = Cxy[2] + (Cxy[1] * (Xstraight - Cxy[2] * Cxy[1] + Cxy[1] * Ystraight)) / (pow(Cxy[1], 2) + 1);
  const double S = sqrt(pow(Xroads, 2) + pow(Yroads - Cxy[2], 2));
  double rho = sqrt(pow(Xroads - Xstraight, 2) + pow(Yroads - Ystraight, 2));
  if (Ystraight < (Cxy[1] * Xstraight + Cxy[2])) rho = -rho;
  const double Xarc = -(Cxs[1] - sqrt(pow(Cxs[1], 2) - 4 * Cxs[0] * Cxs[2] + 4 * Cxs[0] * S)) / (2 * Cxs[0]);
  const double Yarc = Cxy[0] * pow(Xarc, 2) + Cxy[1] * Xarc + Cxy[2];
  const double dydx = 2 * Cxy[0] * Xarc + Cxy[1];
  const double HroadC = atan2(dydx, 1), HroadS = atan2(Cxy[1], 1);
  const double Hcar = wrapTo2Pi((Hstraight - HroadS) + HroadC);
  const double vx = 1, vy = dydx;
  const double L = sqrt(pow(vx, 2) + pow(vy, 2));
  const double nx = -(1 / L) * dydx, ny = (1 / L);
  Xstraight = Xarc + nx * rho; Ystraight = Yarc + ny * rho; Hstraight = Hcar;
}

double RoadFrame::steerOfRoadCurvature(double x, double y, const Vehicle& veh) const {
  double Xarc, Yarc;
  closestPointOnArc(x, y, Xarc, Yarc);
  // upstream writes pow(..., (3/2)): integer division, exponent 1
  const double curvature = (2 * Cxy_[0]) / pow((pow(Cxy_[1] + 2 * Cxy_[0] * Xarc, 2) + 1), (3 / 2));
  return atan(curvature * veh.L);
}
void RoadFrame::stateCarToRoad(state_type& state, const Vehicle& veh) const {
  const double delta = steerOfRoadCurvature(state[0], state[1], veh);  // foot point of the car-frame position
  poseCarToRoad(state[0], state[1], state[2]);
  state[3] -= delta;
}
void RoadFrame::stateRoadToCar(state_type& state, const Vehicle& veh) const {
  poseRoadToCar(state[0], state[1], state[2]);
  state[3] += steerOfRoadCurvature(state[0], state[1], veh);           // foot point of the transformed position
}

namespace {
template <typename Seq, typename PointFn, typename StateFn> void each(Seq& seq, PointFn pf, StateFn sf) {
  for (auto& e : seq) {
    for (size_t i = 0; i != e.ref.x.size(); i++) pf(e.ref.x[i], e.ref.y[i]);
    for (size_t i = 0; i != e.tra.size(); i++) sf(e.tra[i]);
  }
}
}  // namespace
void RoadFrame::pathCarToRoad(std::vector<Path>& path, const Vehicle& veh) const {
  each(path, [&](double& x, double& y) { pointCarToRoad(x, y); }, [&](state_type& s) { stateCarToRoad(s, veh); });
}
void RoadFrame::pathRoadToCar(std::vector<Path>& path, const Vehicle& veh) const {
  each(path, [&](double& x, double& y) { pointRoadToCar(x, y); }, [&](state_type& s) { stateRoadToCar(s, veh); });
}
void RoadFrame::nodesCarToRoad(std::vector<Node>& nodes, const Vehicle& veh) const {
  for (auto& n : nodes) stateCarToRoad(n.state, veh);
  each(nodes, [&](double& x, double& y) { pointCarToRoad(x, y); }, [&](state_type& s) { stateCarToRoad(s, veh); });
}
void RoadFrame::nodesRoadToCar(std::vector<Node>& nodes, const Vehicle& veh) const {
  for (auto& n : nodes) stateRoadToCar(n.state, veh);
  each(nodes, [&](double& x, double& y) { pointRoadToCar(x, y); }, [&](state_type& s) { stateRoadToCar(s, veh); });
}

void transformPointWorldToCar(double& Xw, double& Yw, const std::vector<double>& carPose) {
  const double Xc = Xw * cos(carPose[2]) - carPose[0] * cos(carPose[2]) - carPose[1] * sin(carPose[2]) + Yw * sin(carPose[2]);
  const double Yc = Yw * cos(carPose[2]) - carPose[1] * cos(carPose[2]) + carPose[0] * sin(carPose[2]) - Xw * sin(carPose[2]);
  Xw = Xc; Yw = Yc;
}
void transformPointCarToWorld(double& Xc, double& Yc, const std::vector<double>& carPose) {
  const double Xw = cos(carPose[2]) * Xc - sin(carPose[2]) * Yc + carPose[0];
  const double Yw = sin(carPose[2]) * Xc + cos(carPose[2]) * Yc + carPose[1];
  Xc = Xw; Yc = Yw;
}
void transformPathWorldToCar(std::vector<Path>& path, const std::vector<double>& carPose) {
  each(path, [&](double& x, double& y) { transformPointWorldToCar(x, y, carPose); },
       [&](state_type& s) { transformPointWorldToCar(s[0], s[1], carPose); s[2] -= carPose[2]; });
}
void transformPathCarToWorld(std::vector<Path>& path, const std::vector<double>& worldState) {
  each(path, [&](double& x, double& y) { transformPointCarToWorld(x, y, worldState); },
       [&](state_type& s) { transformPointCarToWorld(s[0], s[1], worldState); s[2] += worldState[2]; });
}

double lookaheadDistance(double v, const PlannerParams& prm) {  // updateLookahead, controller.cpp:13-16
  const double dla_c = prm.ctrl_mindla - prm.ctrl_tla * prm.ctrl_dlavmin;
  return std::max(prm.ctrl_mindla, dla_c + prm.ctrl_tla * std::abs(v));
}
Point sampleOnLane(const std::vector<double>& Cxy, const std::vector<double>& laneShifts, double Lmax, double ctrl_dla) {
  if (Cxy.size() < 3 || laneShifts.empty()) throw Error("sampleOnLane needs Cxy[0..2] and at least one lane shift");
  // arc length uniform in [dla, Lmax] (float division of rand(), as upstream), then one of the lane shifts
  const double S = ctrl_dla + static_cast<float>(rand()) / (static_cast<float>(RAND_MAX / (Lmax - ctrl_dla)));
  const double r = static_cast<double>(rand()) / (static_cast<double>(RAND_MAX / (((laneShifts.size() - 1)))));
  const int laneIndex = (int)floor(r + 0.5);
  const double rho = laneShifts[(size_t)laneIndex];
  Point sample;
  StraightRoad(Cxy).place(S, rho, sample.x, sample.y);
  return sample;
}

}  // namespace clrrt

// ---- flat C view (include/clrrt_host.h) ------------------------------------------------------------------------------
// what: 0 point car->road, 1 point road->car, 2 pose car->road, 3 pose road->car, 4 state car->road, 5 state road->car
// (Prius vehicle), 6 closest point on the arc (xyh[0..1] in, out).  xyh: n x 4 doubles (x, y, heading, delta), in place.
extern "C" int clrrt_host_road_transform(int what, const double* Cxy3, const double* Cxs3, double* xyhd, int n) {
  if (!Cxy3 || !Cxs3 || !xyhd || n < 0 || what < 0 || what > 6) return CLRRT_ERR_ARG;
  try {
    clrrt::RoadFrame rf(std::vector<double>(Cxy3, Cxy3 + 3), std::vector<double>(Cxs3, Cxs3 + 3));
    clrrt::Vehicle veh; veh.setPrius();
    for (int i = 0; i < n; i++) {
      double* p = xyhd + 4 * (size_t)i;
      if (what == 0) rf.pointCarToRoad(p[0], p[1]);
      else if (what == 1) rf.pointRoadToCar(p[0], p[1]);
      else if (what == 2) rf.poseCarToRoad(p[0], p[1], p[2]);
      else if (what == 3) rf.poseRoadToCar(p[0], p[1], p[2]);
      else if (what == 6) { double xa, ya; rf.closestPointOnArc(p[0], p[1], xa, ya); p[0] = xa; p[1] = ya; }
      else {
        clrrt::state_type s = {p[0], p[1], p[2], p[3], 0, 0};
        if (what == 4) rf.stateCarToRoad(s, veh); else rf.stateRoadToCar(s, veh);
        for (int k = 0; k < 4; k++) p[k] = s[(size_t)k];
      }
    }
    return CLRRT_OK;
  } catch (const std::exception&) {
    return CLRRT_ERR_ARG;
  }
}
// K samples of sampleOnLane followed each by the heuristic draw of expandTree (rrtplanner.cpp:142-143): three rand() per
// sample in upstream's order.  v: the car's speed (look-ahead distance of updateLookahead).
extern "C" int clrrt_host_sample_on_lane(const double* Cxy3, const double* lane_shifts, int n_lanes, double Lmax, double v, int K,
                                         double* sample_xy, uint8_t* heuristic) {
  if (!Cxy3 || !lane_shifts || n_lanes < 1 || K < 0 || !sample_xy || !heuristic) return CLRRT_ERR_ARG;
  const std::vector<double> Cxy(Cxy3, Cxy3 + 3), ls(lane_shifts, lane_shifts + n_lanes);
  const double dla = clrrt::lookaheadDistance(v, clrrt::PlannerParams());
  for (int j = 0; j < K; j++) {
    const clrrt::Point s = clrrt::sampleOnLane(Cxy, ls, Lmax, dla);
    sample_xy[2 * j] = s.x; sample_xy[2 * j + 1] = s.y;
    const double r = static_cast<double>(rand()) / (static_cast<double>(RAND_MAX / (1)));
    heuristic[j] = (r <= 0.7) ? 0 : 1;
  }
  return CLRRT_OK;
}
