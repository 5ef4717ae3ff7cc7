#pragma once
namespace vision_msgs {
struct Pose2D { double x = 0, y = 0, theta = 0; };
struct BoundingBox2D { Pose2D center; double size_x = 0, size_y = 0; };
}
