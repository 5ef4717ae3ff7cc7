#pragma once
#include <vector>
namespace car_msgs { struct State { std::vector<double> state; }; }
