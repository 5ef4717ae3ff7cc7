#pragma once
#include <string>
#include <vector>
#include <ros/ros.h>
#include <geometry_msgs/Point.h>
namespace visualization_msgs {
struct Marker {
  enum { ARROW = 0, LINE_STRIP = 4, POINTS = 8 };
  enum { ADD = 0, DELETE = 2, DELETEALL = 3 };
  struct { std::string frame_id; ros::Time stamp; } header;
  std::string ns;
  int id = 0, type = 0, action = 0;
  geometry_msgs::Pose pose;
  geometry_msgs::Vector3 scale;
  struct { float r = 0, g = 0, b = 0, a = 0; } color;
  ros::Duration lifetime;
  std::vector<geometry_msgs::Point> points;
};
}
