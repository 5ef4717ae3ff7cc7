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
  // Foot point of P on the parabola: the real root of d/dx |P - road(x)|^2 = 0, a depressed cubic solved in closed form
  // (Cardano).  Upstream's generated code keeps every intermediate in float; the same roundings are kept here, one named
  // float per intermediate, so the root — and everything derived from it — is bit-identical.
  const double a = Cxy_[0], b = Cxy_[1], c = Cxy_[2];
  const float abs_a = std::abs(a);
  const float abs_a3 = pow(abs_a, 3);
  const float a3 = pow(a, 3);
  const float inv_a3 = 1.0 / a3;
  const float b2 = pow(b, 2);
  const float two_ax = Xcar * a * 2.0;
  const float four_ay = Ycar * a * 4.0;
  const float four_ac = a * c * 4.0;
  const float neg_four_ac = -four_ac;
  const float root3 = sqrt(3.0);
  const float lin = b + two_ax;                       // slope-like term of the cubic
  const float lin2 = pow(lin, 2);
  const float term_b = b * abs_a3 * 9.0;
  const float term_x = Xcar * a * abs_a3 * 18;
  const float q = b2 + four_ay + neg_four_ac - 2.0;   // the cubic's linear coefficient (scaled)
  const float lin2_27 = lin2 * 27;
  const float q3 = pow(q, 3);
  const float neg_q3 = -q3;
  const float disc = lin2_27 + neg_q3;                // discriminant
  const float sqrt_disc = sqrt(disc);
  const float rad = a3 * root3 * sqrt_disc;
  const float num = term_b + term_x + rad;
  const float w = inv_a3 * num;                       // the number whose cube root carries the solution
  const double cbrt_w = pow(w, 1.0 / 3.0);
  Xarc = (cbrt_w * 0.2403749283845681) / abs_a - b / (a * 2.0) + 1.0 / pow(a, 2) * abs_a * q * 1.0 / pow(w, 1.0 / 3.0) * 0.3466806371753173;
  Yarc = a * pow(Xarc, 2) + b * Xarc + c;
}

namespace {
// (S, rho) of a car-frame point from its foot point on the arc; `tangent_form` selects which of upstream's two (algebraically
// equal, differently rounded) half-plane expressions decides the sign of rho: the point transform expands the slope, the
// pose transform uses it as one number
struct RoadCoords { double S, rho, dydx; };
RoadCoords roadCoordsOfCarPoint(const std::vector<double>& Cxy, const std::vector<double>& Cxs, double Xcar, double Ycar, double Xarc,
                                double Yarc, bool tangent_form) {
  RoadCoords r;
  r.S = Cxs[0] * pow(Xarc, 2) + Cxs[1] * Xarc + Cxs[2];
  r.rho = sqrt(pow(Xarc - Xcar, 2) + pow(Yarc - Ycar, 2));
  r.dydx = 2 * Cxy[0] * Xarc + Cxy[1];
  const bool below = tangent_form ? (Ycar < (r.dydx * Xcar + Yarc - r.dydx * Xarc))
                                  : (Ycar < (Yarc - Xarc * (Cxy[1] + 2 * Xarc * Cxy[0]) + Xcar * (Cxy[1] + 2 * Xarc * Cxy[0])));
  if (below) r.rho = -r.rho;
  return r;
}
// the arc point, slope and signed offset that a point of the straightened road maps back to
struct ArcCoords { double Xarc, Yarc, dydx, rho; };
ArcCoords arcCoordsOfRoadPoint(const std::vector<double>& Cxy, const std::vector<double>& Cxs, double Xs, double Ys) {
  const double Xroads = (Xs - Cxy[2] * Cxy[1] + Cxy[1] * Ys) / (pow(Cxy[1], 2) + 1);
  const double Yroads = Cxy[2] + (Cxy[1] * (Xs - Cxy[2] * Cxy[1] + Cxy[1] * Ys)) / (pow(Cxy[1], 2) + 1);
  const double S = sqrt(pow(Xroads, 2) + pow(Yroads - Cxy[2], 2));
  ArcCoords r;
  r.rho = sqrt(pow(Xroads - Xs, 2) + pow(Yroads - Ys, 2));
  if (Ys < (Cxy[1] * Xs + Cxy[2])) r.rho = -r.rho;
  r.Xarc = -(Cxs[1] - sqrt(pow(Cxs[1], 2) - 4 * Cxs[0] * Cxs[2] + 4 * Cxs[0] * S)) / (2 * Cxs[0]);  // inverse of the arc-length fit
  r.Yarc = Cxy[0] * pow(r.Xarc, 2) + Cxy[1] * r.Xarc + Cxy[2];
  r.dydx = 2 * Cxy[0] * r.Xarc + Cxy[1];
  return r;
}
// from the arc point along the road's normal
void offsetAlongNormal(const ArcCoords& r, double& x, double& y) {
  const double vx = 1, vy = r.dydx;
  const double L = sqrt(pow(vx, 2) + pow(vy, 2));
  const double nx = -(1 / L) * r.dydx, ny = (1 / L);
  x = r.Xarc + nx * r.rho; y = r.Yarc + ny * r.rho;
}
}  // namespace

void RoadFrame::pointCarToRoad(double& Xcar, double& Ycar) const {
  double Xarc, Yarc;
  closestPointOnArc(Xcar, Ycar, Xarc, Yarc);
  const RoadCoords r = roadCoordsOfCarPoint(Cxy_, Cxs_, Xcar, Ycar, Xarc, Yarc, false);
  const double theta = atan2(Cxy_[1], 1);
  const double Xstraight = cos(theta) * r.S - sin(theta) * r.rho;
  const double Ystraight = sin(theta) * r.S + cos(theta) * r.rho + Cxy_[2];
  Xcar = Xstraight; Ycar = Ystraight;
}

void RoadFrame::pointRoadToCar(double& Xstraight, double& Ystraight) const {
  offsetAlongNormal(arcCoordsOfRoadPoint(Cxy_, Cxs_, Xstraight, Ystraight), Xstraight, Ystraight);
}

void RoadFrame::poseCarToRoad(double& Xcar, double& Ycar, double& Hcar) const {
  double Xarc, Yarc;
  closestPointOnArc(Xcar, Ycar, Xarc, Yarc);
  const RoadCoords r = roadCoordsOfCarPoint(Cxy_, Cxs_, Xcar, Ycar, Xarc, Yarc, true);
  const double theta = atan2(Cxy_[1], 1);    // heading of the straightened road
  const double Hroad = atan2(r.dydx, 1);     // heading of the road at the foot point
  const double Hstraight = wrapTo2Pi((Hcar - Hroad) + theta);
  const double Xstraight = cos(theta) * r.S - sin(theta) * r.rho;
  const double Ystraight = sin(theta) * r.S + cos(theta) * r.rho + Cxy_[2];
  Xcar = Xstraight; Ycar = Ystraight; Hcar = Hstraight;
}

void RoadFrame::poseRoadToCar(double& Xstraight, double& Ystraight, double& Hstraight) const {
  const ArcCoords r = arcCoordsOfRoadPoint(Cxy_, Cxs_, Xstraight, Ystraight);
  const double HroadC = atan2(r.dydx, 1), HroadS = atan2(Cxy_[1], 1);
  const double Hcar = wrapTo2Pi((Hstraight - HroadS) + HroadC);
  offsetAlongNormal(r, Xstraight, Ystraight);
  Hstraight = Hcar;
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
