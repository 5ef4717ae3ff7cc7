#pragma once
#include <car_msgs/Reference.h>
namespace car_msgs { struct MotionPlan { std::vector<Reference> refArray; }; }
