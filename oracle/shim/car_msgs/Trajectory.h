#pragma once
#include <vector>
namespace car_msgs { struct Trajectory { std::vector<double> x, y, theta, delta, v, a, a_cmd, d_cmd; }; }
