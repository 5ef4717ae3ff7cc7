// clrrt_road.hpp — curved-road mode of the planner's host side: lane sampling and the road-frame
// ("straightened road") transforms.
//
// Upstream these live in rrt/src/rrtplanner.cpp:204-224 (sampleOnLane) and rrt/src/transformations.cpp:20-287 (the arc
// projection and the car <-> road transforms of points, poses, states, paths and nodes).  None of them is reachable from
// the shipped planMotion (expandTree samples around the vehicle in both modes, rrtplanner.cpp:130-135; bend is false in every
// request the mission planner sends), so nothing here is on the GPU path: it completes the host-side mirror for a caller
// that plans in the road frame.  The road is y = Cxy[0] x^2 + Cxy[1] x + Cxy[2] in the car frame, its arc length
// S = Cxs[0] x^2 + Cxs[1] x + Cxs[2]; the straightened road is the line through (0, Cxy[2]) with slope Cxy[1].
// Arithmetic follows the reference expression by expression (including the float temporaries of the arc projection), so
// results equal the reference's bit for bit (tests/test_road_frame.py against oracle/_ref).
#pragma once
#include <vector>

#include "clrrt_planner.hpp"

namespace clrrt {

class RoadFrame {
 public:
  RoadFrame(const std::vector<double>& Cxy, const std::vector<double>& Cxs);
  // findClosestPointOnArc, transformations.cpp:21-52: foot point of (x, y) on the road's parabola (closed form)
  void closestPointOnArc(double x, double y, double& xarc, double& yarc) const;
  // transformPointCarToRoad :55-80 / transformPointRoadToCar :83-111
  void pointCarToRoad(double& x, double& y) const;
  void pointRoadToCar(double& x, double& y) const;
  // transformPoseCarToRoad :149-173 / transformPoseRoadToCar :175-202 (heading wrapped to [0, 2 pi))
  void poseCarToRoad(double& x, double& y, double& heading) const;
  void poseRoadToCar(double& x, double& y, double& heading) const;
  // transformStateCarToRoad :126-133 / transformStateRoadToCar :135-142: pose + the steer angle that follows the road's
  // curvature at the foot point (upstream's exponent (3/2) is integer division: 1)
  void stateCarToRoad(state_type& state, const Vehicle& veh) const;
  void stateRoadToCar(state_type& state, const Vehicle& veh) const;
  // transformPathCarToRoad / RoadToCar :204-216, :242-253 and transformNodesCarToRoad / RoadToCar :256-287: every reference
  // point, every trajectory state (and the node's own state)
  void pathCarToRoad(std::vector<Path>& path, const Vehicle& veh) const;
  void pathRoadToCar(std::vector<Path>& path, const Vehicle& veh) const;
  void nodesCarToRoad(std::vector<Node>& nodes, const Vehicle& veh) const;
  void nodesRoadToCar(std::vector<Node>& nodes, const Vehicle& veh) const;

 private:
  double steerOfRoadCurvature(double x, double y, const Vehicle& veh) const;
  std::vector<double> Cxy_, Cxs_;
};

// transformPointWorldToCar / CarToWorld, transformStateWorldToCar / CarToWorld, transformPathWorldToCar / CarToWorld
// (transformations.cpp:6-17, :117-124, :217-240)
void transformPointWorldToCar(double& Xw, double& Yw, const std::vector<double>& carPose);
void transformPointCarToWorld(double& Xc, double& Yc, const std::vector<double>& carPose);
void transformPathWorldToCar(std::vector<Path>& path, const std::vector<double>& carPose);
void transformPathCarToWorld(std::vector<Path>& path, const std::vector<double>& worldState);

// sampleOnLane, rrtplanner.cpp:204-224: a point on one of the lane centre lines of the straightened road, arc length
// uniform in [ctrl_dla, Lmax], lane picked at random; two rand() draws in upstream's order.  ctrl_dla is the look-ahead
// distance global (updateLookahead(v), controller.cpp:13-16).
Point sampleOnLane(const std::vector<double>& Cxy, const std::vector<double>& laneShifts, double Lmax, double ctrl_dla);
double lookaheadDistance(double v, const PlannerParams& prm);

}  // namespace clrrt
