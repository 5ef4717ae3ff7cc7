#pragma once
// car_msgs/msg/Obstacle2D.msg:1-2
#include <geometry_msgs/Point.h>
#include <vision_msgs/Detection2DArray.h>
namespace car_msgs { struct Obstacle2D { vision_msgs::BoundingBox2D obb; geometry_msgs::Twist vel; }; }
