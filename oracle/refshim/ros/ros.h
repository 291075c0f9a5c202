// ros/ros.h — stand-in (TEST INFRASTRUCTURE ONLY, see oracle/refshim/README.md): logging macros that evaluate nothing,
// a NodeHandle whose param() returns the default, a Publisher that drops what it is given.
#ifndef DPOSE_REFSHIM_ROS
#define DPOSE_REFSHIM_ROS
#include <string>
#define ROS_DEBUG_STREAM(args) do { } while (0)
#define ROS_INFO_STREAM(args) do { } while (0)
#define ROS_WARN_STREAM(args) do { } while (0)
#define ROS_ERROR_STREAM(args) do { } while (0)
#define ROS_FATAL_STREAM(args) do { } while (0)
namespace ros {
inline void init(int&, char**, const std::string&) {}
struct Publisher {
  template <typename M>
  void publish(const M&) const {}
};
struct NodeHandle {
  NodeHandle() = default;
  explicit NodeHandle(const std::string&) {}
  template <typename T>
  T param(const std::string&, const T& fallback) const {
    return fallback;
  }
  std::string param(const std::string&, const char* fallback) const { return fallback; }
  template <typename M>
  Publisher advertise(const std::string&, unsigned) {
    return Publisher();
  }
};
}  // namespace ros
#endif
