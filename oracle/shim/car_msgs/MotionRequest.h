#pragma once
// car_msgs/msg/MotionRequest.msg:1-6
#include <vector>
namespace car_msgs { struct MotionRequest { std::vector<double> goal; double vmax = 0; bool bend = false; std::vector<double> Cxy, Cxs, laneShifts; }; }
