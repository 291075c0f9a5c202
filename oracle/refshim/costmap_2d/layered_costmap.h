// costmap_2d/layered_costmap.h — stand-in (TEST INFRASTRUCTURE ONLY, see oracle/refshim/README.md) for the
// members of costmap_2d::Costmap2D / LayeredCostmap that dpose_costmap.cpp calls (ROS noetic signatures:
// unsigned sizes, row-major unsigned char grid, index = my * size_x + mx, recursive mutex).
#ifndef DPOSE_REFSHIM_COSTMAP_2D
#define DPOSE_REFSHIM_COSTMAP_2D

#include <boost/thread/lock_types.hpp>

#include <utility>
#include <vector>

namespace geometry_msgs {
struct Point {
  double x = 0, y = 0, z = 0;
};
}  // namespace geometry_msgs

namespace costmap_2d {

static const unsigned char NO_INFORMATION = 255;
static const unsigned char LETHAL_OBSTACLE = 254;
static const unsigned char INSCRIBED_INFLATED_OBSTACLE = 253;
static const unsigned char FREE_SPACE = 0;

class Costmap2D {
public:
  typedef boost::recursive_mutex mutex_t;

  /// views a caller-owned grid (the bridge keeps it alive); nothing is copied
  Costmap2D(unsigned char* grid, unsigned int size_x, unsigned int size_y, double resolution = 1.0) :
      grid_(grid), size_x_(size_x), size_y_(size_y), resolution_(resolution) {}

  unsigned int getSizeInCellsX() const { return size_x_; }
  unsigned int getSizeInCellsY() const { return size_y_; }
  double getResolution() const { return resolution_; }
  unsigned char* getCharMap() const { return grid_; }
  unsigned int getIndex(unsigned int mx, unsigned int my) const { return my * size_x_ + mx; }
  unsigned char getCost(unsigned int mx, unsigned int my) const { return grid_[getIndex(mx, my)]; }
  mutex_t* getMutex() { return &mutex_; }

private:
  unsigned char* grid_;
  unsigned int size_x_, size_y_;
  double resolution_;
  mutex_t mutex_;
};

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

#endif  // DPOSE_REFSHIM_COSTMAP_2D
