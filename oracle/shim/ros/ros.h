// Minimal stand-in for <ros/ros.h>: just enough surface for the reference's
// planner sources (rrt/src/*.cpp) to compile without ROS.  TEST INFRASTRUCTURE
// ONLY (used by oracle/build_ref.sh); nothing in the product includes this.
// Logging macros compile to nothing (a real ROS build would print inside the
// per-step loop, rrt/src/simulation.cpp:70, and be far slower).
#pragma once
#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstdlib>
#include <iostream>
#include <limits>
#include <map>
#include <string>
#include <vector>

namespace ros {
namespace param {
inline std::map<std::string, double>& store() {
  static std::map<std::string, double> s;
  return s;
}
inline bool get(const std::string& key, double& out) {
  auto it = store().find(key);
  if (it == store().end()) return false;
  out = it->second;
  return true;
}
inline void set(const std::string& key, double v) { store()[key] = v; }
}  // namespace param

struct Time {
  double sec = 0;
  static Time now() { return Time(); }
};
struct Duration {
  double sec;
  Duration() : sec(0) {}
  explicit Duration(double s) : sec(s) {}
};
struct Publisher {
  template <class M> void publish(const M&) const {}
};
// The only service the planner calls is getobstacles (rrt/src/motionplanner.cpp:83-85):
// the driver installs a hook that fills the response.
struct ServiceClient {
  void (*hook)(void* srv) = nullptr;
  template <class S> bool call(S& srv) {
    if (hook) hook(&srv);
    return true;
  }
};
}  // namespace ros

#define ROS_INFO_STREAM(x) do {} while (0)
#define ROS_WARN_STREAM(x) do {} while (0)
#define ROS_ERROR_STREAM(x) do {} while (0)
#define ROS_DEBUG_STREAM(x) do {} while (0)
#define ROS_INFO_STREAM_THROTTLE(p, x) do {} while (0)
#define ROS_WARN_STREAM_THROTTLE(p, x) do {} while (0)
#define ROS_WARN(...) do {} while (0)
#define ROS_INFO(...) do {} while (0)
#define ROS_ERROR(...) do {} while (0)
