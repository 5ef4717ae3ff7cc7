// clrrt_wire.hpp — ROS 1 wire format of the planner's messages, without ROS.
//
// A ROS node wrapping this library exchanges car_msgs with the mission planner, the MPC and the detection node.  The
// byte layout of those messages is fixed by their .msg definitions (little endian; float64[] = uint32 count + doubles;
// T[] of messages = uint32 count + each message; bool = 1 byte; int32 = 4 bytes):
//   car_msgs/MotionRequest   goal[], vmax, bend, Cxy[], Cxs[], laneShifts[]          car_msgs/msg/MotionRequest.msg
//   car_msgs/State           state[]                                                 car_msgs/msg/State.msg
//   car_msgs/Trajectory      x[], y[], theta[], delta[], v[], a[], a_cmd[], d_cmd[]  car_msgs/msg/Trajectory.msg
//   car_msgs/Reference       x[], y[], v[], int32 dir                                car_msgs/msg/Reference.msg
//   car_msgs/MotionResponse  Reference[] ref, Trajectory[] tra                       car_msgs/msg/MotionResponse.msg
//   car_msgs/Obstacle2D      vision_msgs/BoundingBox2D obb (Pose2D center {x, y, theta}, size_x, size_y),
//                            geometry_msgs/Twist vel (Vector3 linear, Vector3 angular)   car_msgs/msg/Obstacle2D.msg
//   car_msgs/getobstacles    response: Obstacle2D[] obstacles                        car_msgs/srv/getobstacles.srv
// so the planner can be fed from, and answer into, raw message buffers (e.g. a rosbag or a TCPROS socket).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "clrrt_planner.hpp"

namespace clrrt {
namespace wire {

typedef std::vector<uint8_t> Bytes;

// car_msgs/msg/Reference.msg and MotionResponse.msg as plain structs (the facade's MyReference has the same fields)
struct MotionResponse {
  std::vector<MyReference> ref;
  std::vector<Trajectory> tra;
};

Bytes serialize(const MotionRequest& m);
Bytes serialize(const Trajectory& m);
Bytes serialize(const MyReference& m);                 // car_msgs/Reference
Bytes serialize(const MotionResponse& m);
Bytes serialize(const std::vector<Obstacle2D>& obstacles);  // getobstacles response
Bytes serializeState(const std::vector<double>& state);     // car_msgs/State

// each returns false on a truncated or oversized buffer; `used` = bytes consumed
bool deserialize(const uint8_t* p, size_t n, MotionRequest& m, size_t* used = nullptr);
bool deserialize(const uint8_t* p, size_t n, Trajectory& m, size_t* used = nullptr);
bool deserialize(const uint8_t* p, size_t n, MyReference& m, size_t* used = nullptr);
bool deserialize(const uint8_t* p, size_t n, MotionResponse& m, size_t* used = nullptr);
bool deserialize(const uint8_t* p, size_t n, std::vector<Obstacle2D>& obstacles, size_t* used = nullptr);
bool deserializeState(const uint8_t* p, size_t n, std::vector<double>& state, size_t* used = nullptr);

}  // namespace wire

// preparePathMessage (rrt/src/motionplanner.cpp:235-261): references and trajectories of a plan (first state of each
// segment skipped, as upstream)
wire::MotionResponse preparePathMessage(const std::vector<Path>& path);
// getCommittedPath (rrt/src/motionplanner.cpp:201-221): the part of the best path covered within Tcommit
std::vector<Path> getCommittedPath(std::vector<Node> bestPath, double& Tp, double sim_dt, double Tcommit);

}  // namespace clrrt

// flat entry points for bindings and tests: include/clrrt_host.h
