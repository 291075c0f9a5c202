// costmap_2d/layered_costmap.h — stand-in (TEST INFRASTRUCTURE ONLY, see oracle/refshim/README.md): the two
// LayeredCostmap accessors that make_footprint uses (dpose_costmap.cpp:40-55).
#ifndef DPOSE_REFSHIM_LAYERED_COSTMAP
#define DPOSE_REFSHIM_LAYERED_COSTMAP

#include <costmap_2d/costmap_2d.h>
#include <geometry_msgs/Point.h>

#include <memory>
#include <string>

#include <utility>
#include <vector>

namespace costmap_2d {

class LayeredCostmap {
public:
  /// views a costmap owned by the caller (bridge.cpp)
  LayeredCostmap(Costmap2D* map, std::vector<geometry_msgs::Point> footprint) : map_(map), footprint_(std::move(footprint)) {}
  /// the ROS constructor (global_frame, rolling_window, track_unknown): an empty costmap until resizeMap
  LayeredCostmap(std::string, bool, bool) {}

  void resizeMap(unsigned int size_x, unsigned int size_y, double resolution, double origin_x, double origin_y, bool = false) {
    own_.reset(new Costmap2D(size_x, size_y, resolution, origin_x, origin_y));
    map_ = own_.get();
  }
  void setFootprint(const std::vector<geometry_msgs::Point>& footprint) { footprint_ = footprint; }
  Costmap2D* getCostmap() { return map_; }
  const std::vector<geometry_msgs::Point>& getFootprint() { return footprint_; }

private:
  std::unique_ptr<Costmap2D> own_;
  Costmap2D* map_ = nullptr;
  std::vector<geometry_msgs::Point> footprint_;
};

}  // namespace costmap_2d

#endif  // DPOSE_REFSHIM_LAYERED_COSTMAP
