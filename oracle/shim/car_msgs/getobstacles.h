#pragma once
#include <vector>
#include <car_msgs/Obstacle2D.h>
namespace car_msgs { struct getobstacles { struct {} request; struct { std::vector<Obstacle2D> obstacles; } response; }; }
