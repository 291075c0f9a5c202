// costmap_2d/costmap_2d.h — stand-in (TEST INFRASTRUCTURE ONLY, see oracle/refshim/README.md) for the members of
// costmap_2d::Costmap2D that dpose_core and its unit tests call (ROS noetic signatures and behaviour: unsigned
// sizes, row-major unsigned char grid, index = my * size_x + mx, recursive mutex, raytraceLine that marks the end
// cell, polygonOutlineCells that closes the polygon, convexFillCells that fills every x column between the lowest
// and the highest outline cell).
#ifndef DPOSE_REFSHIM_COSTMAP_2D_H
#define DPOSE_REFSHIM_COSTMAP_2D_H

#include <boost/thread/lock_types.hpp>
#include <costmap_2d/cost_values.h>

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <vector>

namespace costmap_2d {

struct MapLocation {
  unsigned int x;
  unsigned int y;
};

class Costmap2D {
public:
  typedef boost::recursive_mutex mutex_t;

  /// owns its grid, as the ROS constructor (cells_size_x, cells_size_y, resolution, origin_x, origin_y, default)
  Costmap2D(unsigned int size_x, unsigned int size_y, double resolution, double origin_x, double origin_y,
            unsigned char default_value = 0) :
      own_((size_t)size_x * size_y, default_value),
      grid_(own_.data()),
      size_x_(size_x),
      size_y_(size_y),
      resolution_(resolution),
      origin_x_(origin_x),
      origin_y_(origin_y) {}

  /// views a caller-owned grid (bridge.cpp keeps it alive); nothing is copied
  Costmap2D(unsigned char* grid, unsigned int size_x, unsigned int size_y, double resolution = 1.0) :
      grid_(grid), size_x_(size_x), size_y_(size_y), resolution_(resolution) {}

  Costmap2D(const Costmap2D&) = delete;
  Costmap2D& operator=(const Costmap2D&) = delete;

  unsigned int getSizeInCellsX() const { return size_x_; }
  unsigned int getSizeInCellsY() const { return size_y_; }
  double getResolution() const { return resolution_; }
  double getOriginX() const { return origin_x_; }
  double getOriginY() const { return origin_y_; }
  unsigned char* getCharMap() const { return grid_; }
  unsigned int getIndex(unsigned int mx, unsigned int my) const { return my * size_x_ + mx; }
  void indexToCells(unsigned int index, unsigned int& mx, unsigned int& my) const {
    my = index / size_x_;
    mx = index - (my * size_x_);
  }
  unsigned char getCost(unsigned int mx, unsigned int my) const { return grid_[getIndex(mx, my)]; }
  void setCost(unsigned int mx, unsigned int my, unsigned char cost) { grid_[getIndex(mx, my)] = cost; }
  mutex_t* getMutex() { return &mutex_; }

  void mapToWorld(unsigned int mx, unsigned int my, double& wx, double& wy) const {
    wx = origin_x_ + (mx + 0.5) * resolution_;
    wy = origin_y_ + (my + 0.5) * resolution_;
  }
  bool worldToMap(double wx, double wy, unsigned int& mx, unsigned int& my) const {
    if (wx < origin_x_ || wy < origin_y_) return false;
    mx = (int)((wx - origin_x_) / resolution_);
    my = (int)((wy - origin_y_) / resolution_);
    return mx < size_x_ && my < size_y_;
  }

  /// Bresenham from (x0, y0) to (x1, y1), both end cells included (raytraceLine + bresenham2D)
  template <typename Action>
  void raytraceLine(Action at, unsigned int x0, unsigned int y0, unsigned int x1, unsigned int y1) const {
    const int dx = (int)x1 - (int)x0, dy = (int)y1 - (int)y0;
    const unsigned int abs_dx = std::abs(dx), abs_dy = std::abs(dy);
    const int offset_dx = dx > 0 ? 1 : -1, offset_dy = (dy > 0 ? 1 : -1) * (int)size_x_;
    unsigned int offset = y0 * size_x_ + x0;
    if (abs_dx >= abs_dy)
      bresenham2D(at, abs_dx, abs_dy, (int)(abs_dx / 2), offset_dx, offset_dy, offset);
    else
      bresenham2D(at, abs_dy, abs_dx, (int)(abs_dy / 2), offset_dy, offset_dx, offset);
  }

  void polygonOutlineCells(const std::vector<MapLocation>& polygon, std::vector<MapLocation>& cells) const {
    auto gather = [&](unsigned int offset) {
      MapLocation loc;
      indexToCells(offset, loc.x, loc.y);
      cells.push_back(loc);
    };
    for (size_t i = 0; i + 1 < polygon.size(); ++i)
      raytraceLine(gather, polygon[i].x, polygon[i].y, polygon[i + 1].x, polygon[i + 1].y);
    if (!polygon.empty())  // close the polygon: last vertex back to the first
      raytraceLine(gather, polygon.back().x, polygon.back().y, polygon[0].x, polygon[0].y);
  }

  void convexFillCells(const std::vector<MapLocation>& polygon, std::vector<MapLocation>& cells) const {
    if (polygon.size() < 3) return;
    polygonOutlineCells(polygon, cells);
    // ROS bubble-sorts by x; a stable sort by x gives the same per-column groups
    std::stable_sort(cells.begin(), cells.end(), [](const MapLocation& a, const MapLocation& b) { return a.x < b.x; });
    size_t i = 0;
    const size_t n_outline = cells.size();
    const unsigned int min_x = cells.front().x, max_x = cells[n_outline - 1].x;
    for (unsigned int x = min_x; x <= max_x; ++x) {
      if (i + 1 >= n_outline) break;
      MapLocation min_pt, max_pt;
      if (cells[i].y < cells[i + 1].y)
        min_pt = cells[i], max_pt = cells[i + 1];
      else
        min_pt = cells[i + 1], max_pt = cells[i];
      i += 2;
      while (i < n_outline && cells[i].x == x) {
        if (cells[i].y < min_pt.y)
          min_pt = cells[i];
        else if (cells[i].y > max_pt.y)
          max_pt = cells[i];
        ++i;
      }
      for (unsigned int y = min_pt.y; y <= max_pt.y; ++y) cells.push_back(MapLocation{x, y});
    }
  }

private:
  template <typename Action>
  void bresenham2D(Action at, unsigned int abs_da, unsigned int abs_db, int error_b, int offset_a, int offset_b,
                   unsigned int offset) const {
    for (unsigned int i = 0; i < abs_da; ++i) {
      at(offset);
      offset += offset_a;
      error_b += abs_db;
      if ((unsigned int)error_b >= abs_da) {
        offset += offset_b;
        error_b -= abs_da;
      }
    }
    at(offset);
  }

  std::vector<unsigned char> own_;
  unsigned char* grid_;
  unsigned int size_x_, size_y_;
  double resolution_;
  double origin_x_ = 0, origin_y_ = 0;
  mutex_t mutex_;
};

}  // namespace costmap_2d

#endif  // DPOSE_REFSHIM_COSTMAP_2D_H
