// geometry_msgs — stand-in messages (TEST INFRASTRUCTURE ONLY, see oracle/refshim/README.md)
#ifndef DPOSE_REFSHIM_GEOMETRY_MSGS
#define DPOSE_REFSHIM_GEOMETRY_MSGS
#include <string>
namespace geometry_msgs {
struct Point {
  double x = 0, y = 0, z = 0;
};
struct Quaternion {
  double x = 0, y = 0, z = 0, w = 1;
};
struct Pose {
  Point position;
  Quaternion orientation;
};
struct Header {
  std::string frame_id;
};
struct PoseStamped {
  Header header;
  Pose pose;
};
}  // namespace geometry_msgs
#endif
