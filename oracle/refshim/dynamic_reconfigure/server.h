// dynamic_reconfigure/server.h — stand-in (TEST INFRASTRUCTURE ONLY): setCallback calls back once with the defaults,
// as the real server does
#ifndef DPOSE_REFSHIM_DYNAMIC_RECONFIGURE
#define DPOSE_REFSHIM_DYNAMIC_RECONFIGURE
#include <ros/ros.h>

#include <cstdint>
#include <functional>
namespace dynamic_reconfigure {
template <typename Config>
class Server {
public:
  explicit Server(const ros::NodeHandle&) {}
  void setCallback(const std::function<void(Config&, uint32_t)>& cb) {
    Config cfg;
    cb(cfg, ~0u);
  }
};
}  // namespace dynamic_reconfigure
#endif
