// costmap_2d/layered_costmap.h — stand-in (TEST INFRASTRUCTURE ONLY, see oracle/refshim/README.md): the two
// LayeredCostmap accessors that make_footprint uses (dpose_costmap.cpp:40-55).
#ifndef DPOSE_REFSHIM_LAYERED_COSTMAP
#define DPOSE_REFSHIM_LAYERED_COSTMAP

#include <costmap_2d/costmap_2d.h>

#include <utility>
#include <vector>

namespace geometry_msgs {
struct Point {
  double x = 0, y = 0, z = 0;
};
}  // namespace geometry_msgs

namespace costmap_2d {

class LayeredCostmap {
public:
  LayeredCostmap(Costmap2D* map, std::vector<geometry_msgs::Point> footprint) : map_(map), footprint_(std::move(footprint)) {}
  Costmap2D* getCostmap() { return map_; }
  const std::vector<geometry_msgs::Point>& getFootprint() { return footprint_; }

private:
  Costmap2D* map_;
  std::vector<geometry_msgs::Point> footprint_;
};

}  // namespace costmap_2d

#endif  // DPOSE_REFSHIM_LAYERED_COSTMAP
