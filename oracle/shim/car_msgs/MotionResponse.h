#pragma once
#include <car_msgs/Reference.h>
#include <car_msgs/Trajectory.h>
namespace car_msgs { struct MotionResponse { std::vector<Reference> ref; std::vector<Trajectory> tra; }; }
