// angles/angles.h — stand-in (TEST INFRASTRUCTURE ONLY, see oracle/refshim/README.md): the two functions of ROS
// `angles` (noetic, 1.9.13) that dpose_goal_tolerance.cpp calls.
#ifndef DPOSE_REFSHIM_ANGLES
#define DPOSE_REFSHIM_ANGLES
#include <cmath>
namespace angles {
static inline double
normalize_angle(double angle) {
  const double result = std::fmod(angle + M_PI, 2.0 * M_PI);
  if (result <= 0.0) return result + M_PI;
  return result - M_PI;
}
static inline double
shortest_angular_distance(double from, double to) {
  return normalize_angle(to - from);
}
}  // namespace angles
#endif
