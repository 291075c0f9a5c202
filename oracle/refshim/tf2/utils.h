// tf2/utils.h — stand-in (TEST INFRASTRUCTURE ONLY, see oracle/refshim/README.md): yaw of a quaternion message and the
// yaw-only quaternion the plugin's to_msg builds.
#ifndef DPOSE_REFSHIM_TF2
#define DPOSE_REFSHIM_TF2
#include <geometry_msgs/Point.h>

#include <cmath>
namespace tf2 {
struct Quaternion {
  double x = 0, y = 0, z = 0, w = 1;
  void setRPY(double roll, double pitch, double yaw) {
    const double cr = std::cos(roll / 2), sr = std::sin(roll / 2), cp = std::cos(pitch / 2), sp = std::sin(pitch / 2);
    const double cy = std::cos(yaw / 2), sy = std::sin(yaw / 2);
    x = sr * cp * cy - cr * sp * sy;
    y = cr * sp * cy + sr * cp * sy;
    z = cr * cp * sy - sr * sp * cy;
    w = cr * cp * cy + sr * sp * sy;
  }
};
inline geometry_msgs::Quaternion
toMsg(const Quaternion& q) {
  geometry_msgs::Quaternion m;
  m.x = q.x, m.y = q.y, m.z = q.z, m.w = q.w;
  return m;
}
inline double
getYaw(const geometry_msgs::Quaternion& q) {
  return std::atan2(2.0 * (q.w * q.z + q.x * q.y), 1.0 - 2.0 * (q.y * q.y + q.z * q.z));
}
}  // namespace tf2
#endif
