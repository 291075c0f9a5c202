// gpp_interface/pre_planning_interface.hpp — stand-in (TEST INFRASTRUCTURE ONLY, see oracle/refshim/README.md)
#ifndef DPOSE_REFSHIM_GPP_INTERFACE
#define DPOSE_REFSHIM_GPP_INTERFACE
#include <costmap_2d/costmap_2d_ros.h>
#include <geometry_msgs/Point.h>

#include <string>
namespace gpp_interface {
struct PrePlanningInterface {
  using Pose = geometry_msgs::PoseStamped;
  using Map = costmap_2d::Costmap2DROS;
  virtual ~PrePlanningInterface() = default;
  virtual bool preProcess(Pose& _start, Pose& _goal) = 0;
  virtual void initialize(const std::string& _name, Map* _map) = 0;
};
}  // namespace gpp_interface
#endif
