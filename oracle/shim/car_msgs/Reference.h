#pragma once
#include <vector>
namespace car_msgs { struct Reference { std::vector<double> x, y, v; int dir = 0; }; }
