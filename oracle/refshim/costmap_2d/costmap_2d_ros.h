// costmap_2d/costmap_2d_ros.h — stand-in (TEST INFRASTRUCTURE ONLY, see oracle/refshim/README.md)
#ifndef DPOSE_REFSHIM_COSTMAP_2D_ROS
#define DPOSE_REFSHIM_COSTMAP_2D_ROS
#include <costmap_2d/layered_costmap.h>

#include <string>
namespace costmap_2d {
class Costmap2DROS {
public:
  Costmap2DROS(LayeredCostmap* layered, std::string frame) : layered_(layered), frame_(std::move(frame)) {}
  LayeredCostmap* getLayeredCostmap() const { return layered_; }
  Costmap2D* getCostmap() const { return layered_->getCostmap(); }
  const std::string& getGlobalFrameID() const { return frame_; }

private:
  LayeredCostmap* layered_;
  std::string frame_;
};
}  // namespace costmap_2d
#endif
