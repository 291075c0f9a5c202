// dpose_costmap.hpp — costmap-side helpers of dpose_core on top of the B200 C ABI.
//
// Mirrors upstream dpose_core/src/dpose_core/dpose_costmap.hpp:36-85 (make_footprint, raytrace,
// interval, cell_rays, to_rays, lethal_cells_within) with the same names, argument meaning and
// exceptions.  lethal_cells_within runs the stream-compaction kernel; raytrace / to_rays /
// make_footprint are integer host arithmetic (a few dozen cells).
//
// is_inside, x_bresenham and check_footprint (upstream dpose_costmap.hpp:87-196, dpose_costmap.cpp:186-582)
// validate a solved pose; check_footprint runs on the GPU (one thread per candidate pose, same scan-line
// algorithm) with a batched overload for multi-start.  The four bresenham::* iterator classes are folded
// into x_bresenham (same get_x / get_dx / advance_x behaviour).
#ifndef DPOSE_CORE__DPOSE_COSTMAP__HPP
#define DPOSE_CORE__DPOSE_COSTMAP__HPP

#include "dpose_core.hpp"

#if DPOSE_HAVE_COSTMAP_2D
#include <costmap_2d/layered_costmap.h>
#endif

#include <algorithm>
#include <cmath>
#include <limits>
#include <memory>
#include <mutex>
#include <unordered_map>

namespace dpose_core {

/// metric footprint -> cells, round(p / resolution) per coordinate, order kept, not closed
/// (upstream dpose_costmap.cpp:40-55).  @throws std::runtime_error if the resolution is <= 0
inline polygon
make_footprint(const std::vector<geometry_msgs::Point>& _fp, double _resolution) {
  if (!(_resolution > 0)) throw std::runtime_error("resolution must be positive");
  std::vector<double> xy(2 * _fp.size());
  for (size_t i = 0; i < _fp.size(); ++i) {
    xy[2 * i] = _fp[i].x;
    xy[2 * i + 1] = _fp[i].y;
  }
  polygon out(2, (int)_fp.size());
  detail::check(dpose_make_footprint(xy.data(), (int)_fp.size(), _resolution, out.data()));
  return out;
}

inline polygon
make_footprint(costmap_2d::LayeredCostmap& _cm) {
  return make_footprint(_cm.getFootprint(), _cm.getCostmap()->getResolution());
}

/// Bresenham ray from _begin to _end (exclusive), ROS raytraceLine compatible (upstream :57-104)
inline cell_vector
raytrace(const cell& _begin, const cell& _end) noexcept {
  const int32_t b[2] = {_begin.x(), _begin.y()}, e[2] = {_end.x(), _end.y()};
  const size_t n = (size_t)std::max(std::abs(e[0] - b[0]), std::abs(e[1] - b[1]));
  cell_vector out(n);
  size_t got = 0;
  if (n) dpose_raytrace(b, e, reinterpret_cast<int32_t*>(out.data()), n, &got);
  return out;
}

/// closed interval [min, max], grown with extend() (upstream dpose_costmap.hpp:57-70)
template <typename _T>
struct interval {
  interval() : min{std::numeric_limits<_T>::max()}, max{std::numeric_limits<_T>::lowest()} {}

  inline void
  extend(const _T& _v) noexcept {
    min = std::min(min, _v);
    max = std::max(max, _v);
  }

  _T min, max;
};

using cell_interval = interval<size_t>;
using cell_rays = std::unordered_map<int, cell_interval>;
using cell_rectangle = rectangle<int>;

/// per-row [min x, max x] of the ring's four edges (upstream dpose_costmap.cpp:106-115)
inline cell_rays
to_rays(const cell_rectangle& _rect) noexcept {
  cell_rays rays;
  for (int i = 0; i < 4; ++i) {
    const cell_vector ray = raytrace(cell(_rect(0, i), _rect(1, i)), cell(_rect(0, i + 1), _rect(1, i + 1)));
    for (const auto& c : ray) rays[c.y()].extend((size_t)c.x());
  }
  return rays;
}

/// lethal (== 254) cells inside _bounds, on a costmap that already lives on the device
inline cell_vector
lethal_cells_within(const device_costmap& _map, const cell_rectangle& _bounds) {
  int32_t ring[10];
  for (int i = 0; i < 5; ++i) {
    ring[2 * i] = _bounds(0, i);
    ring[2 * i + 1] = _bounds(1, i);
  }
  int32_t* xy = nullptr;
  size_t n = 0;
  detail::check(dpose_lethal_cells_within(_map.handle(), ring, &xy, &n));
  cell_vector out(n);
  if (n) std::memcpy(static_cast<void*>(out.data()), xy, sizeof(int32_t) * 2 * n);
  dpose_free(xy);
  return out;
}

namespace detail {

/// Device mirror of the host costmap this thread used last.  The upstream signatures take a host Costmap2D whose
/// bytes may have changed since the previous call, so nothing cached is ever trusted: every call re-uploads exactly
/// the window it is about to read (dpose_map_update) into the mirror and then runs on the device.  What the mirror
/// saves is the allocation, stream and event set-up of a device costmap per call.  One mirror per thread; a costmap
/// with another buffer or size replaces it.
inline device_costmap&
mirror_of(const costmap_2d::Costmap2D& _map) {
  struct slot_t {
    const unsigned char* data = nullptr;
    unsigned int sx = 0, sy = 0;
    std::unique_ptr<device_costmap> dev;
  };
  thread_local slot_t slot;
  const unsigned int sx = _map.getSizeInCellsX(), sy = _map.getSizeInCellsY();
  if (!slot.dev || slot.data != _map.getCharMap() || slot.sx != sx || slot.sy != sy) {
    slot.dev.reset(new device_costmap(_map.getCharMap(), sx, sy, (size_t)sx));
    slot.data = _map.getCharMap();
    slot.sx = sx;
    slot.sy = sy;
  }
  return *slot.dev;
}

/// uploads the cells [x0, x1] x [y0, y1] of _map into its mirror and returns the mirror
inline device_costmap&
mirror_window(const costmap_2d::Costmap2D& _map, int x0, int y0, int x1, int y1) {
  device_costmap& dev = mirror_of(_map);
  dev.update(_map.getCharMap() + _map.getIndex((unsigned)x0, (unsigned)y0), (size_t)_map.getSizeInCellsX(), (uint32_t)x0,
             (uint32_t)y0, (uint32_t)(x1 - x0 + 1), (uint32_t)(y1 - y0 + 1));
  return dev;
}

}  // namespace detail

/// upstream signature (dpose_costmap.hpp:84-85, dpose_costmap.cpp:117-184): takes the costmap's
/// mutex (:138), uploads the bounding rows/columns of the clamped box and compacts on the GPU.
/// Rows come back y ascending, x ascending (the reference's row order is unspecified).
inline cell_vector
lethal_cells_within(costmap_2d::Costmap2D& _map, const cell_rectangle& _bounds) {
  const int sx = (int)_map.getSizeInCellsX(), sy = (int)_map.getSizeInCellsY();
  if (sx == 0 || sy == 0) return {};  // :124-125
  // clamp into the map (:121,:128-129); the bounding window of the clamped ring is what gets uploaded
  cell_rectangle ring = _bounds;
  int x0 = sx - 1, y0 = sy - 1, x1 = 0, y1 = 0;
  for (int i = 0; i < 5; ++i) {
    ring(0, i) = std::min(std::max(ring(0, i), 0), sx - 1);
    ring(1, i) = std::min(std::max(ring(1, i), 0), sy - 1);
    x0 = std::min(x0, ring(0, i));
    x1 = std::max(x1, ring(0, i));
    y0 = std::min(y0, ring(1, i));
    y1 = std::max(y1, ring(1, i));
  }
  std::lock_guard<costmap_2d::Costmap2D::mutex_t> lock(*_map.getMutex());
  return lethal_cells_within(detail::mirror_window(_map, x0, y0, x1, y1), ring);
}

/// true if every vertex of _footprint lies inside _map (upstream dpose_costmap.cpp:186-199)
inline bool
is_inside(const costmap_2d::Costmap2D& _map, const polygon& _footprint) noexcept {
  const int sx = (int)_map.getSizeInCellsX(), sy = (int)_map.getSizeInCellsY();
  for (int c = 0; c < (int)_footprint.cols(); ++c)
    if (_footprint(0, c) < 0 || _footprint(1, c) < 0 || _footprint(0, c) >= sx || _footprint(1, c) >= sy) return false;
  return true;
}

/// x-value of a line per scan-line, starting at the end point with the lower y (upstream :201-317: the
/// x_minor / x_major, ascending / descending Bresenham iterators behind one interface)
struct x_bresenham {
  x_bresenham(const cell& _c0, const cell& _c1) noexcept {
    const int ddx = _c1.x() - _c0.x(), ddy = _c1.y() - _c0.y();
    const int ax = std::abs(ddx), ay = std::abs(ddy);
    x_sign = (ddx > 0) - (ddx < 0);
    den = std::max(ax, ay);
    add = std::min(ax, ay);
    num = den / 2;
    dx = static_cast<double>(ddx) / ddy;
    const bool up = _c1.y() > _c0.y();
    kind = ax > ay ? (up ? 2 : 3) : (up ? 0 : 1);
    x_curr = up ? _c0.x() : _c1.x();
  }

  inline void
  advance_x() noexcept {
    switch (kind) {
      case 0:  // x-minor, ascending (:226-237)
        num += add;
        if (num >= den) num -= den, x_curr += x_sign;
        break;
      case 1:  // x-minor, descending (:244-255)
        num -= add;
        if (num < 0) num += den, x_curr -= x_sign;
        break;
      case 2: {  // x-major, ascending (:264-279)
        const int step = (den - num) / add;
        num += step * add, x_curr += step * x_sign;
        if (num < den) num += add, x_curr += x_sign;
        num -= den;
        break;
      }
      default: {  // x-major, descending (:286-301)
        const int step = num / add;
        num -= step * add, x_curr -= step * x_sign;
        if (num >= 0) num -= add, x_curr -= x_sign;
        num += den;
        break;
      }
    }
  }

  inline const int&
  get_x() const noexcept {
    return x_curr;
  }

  inline const double&
  get_dx() const noexcept {
    return dx;
  }

private:
  int x_curr, x_sign, den, add, num, kind;
  double dx;
};

/// true iff _footprint lies inside the map and covers no cell equal to _cost, on a device-resident costmap
inline bool
check_footprint(const device_costmap& _map, const polygon& _footprint, uint8_t _cost) {
  int ok = 0;
  detail::check(dpose_check_footprint(_map.handle(), _footprint.data(), (int)_footprint.cols(), _cost, &ok));
  return ok != 0;
}

/// upstream signature (dpose_costmap.hpp:194-196, dpose_costmap.cpp:460-582); uploads the bounding window of
/// the footprint into this thread's device mirror of the costmap.  The caller is expected to hold the costmap's
/// mutex, as with the reference.
inline bool
check_footprint(const costmap_2d::Costmap2D& _map, const polygon& _footprint, uint8_t _cost) {
  if (!is_inside(_map, _footprint)) return false;
  const int n = (int)_footprint.cols();
  if (n == 0) return true;
  int x0 = _footprint(0, 0), x1 = x0, y0 = _footprint(1, 0), y1 = y0;
  for (int c = 1; c < n; ++c) {
    x0 = std::min(x0, _footprint(0, c)), x1 = std::max(x1, _footprint(0, c));
    y0 = std::min(y0, _footprint(1, c)), y1 = std::max(y1, _footprint(1, c));
  }
  return check_footprint(detail::mirror_window(_map, x0, y0, x1, y1), _footprint, _cost);
}

/// NEW: the plugin's validation step (dpose_goal_tolerance.cpp:386-399) for many candidate poses at once:
/// _footprint (cells, robot frame) is transformed by every pose (x, y, theta), rounded and checked.
inline std::vector<uint8_t>
check_footprint(const device_costmap& _map, const polygon& _footprint, const double* _poses, size_t _n, uint8_t _cost) {
  std::vector<uint8_t> ok(_n);
  detail::check(dpose_check_footprint_batch(_map.handle(), _footprint.data(), (int)_footprint.cols(), _poses, _n, _cost,
                                            ok.data()));
  return ok;
}

}  // namespace dpose_core

#endif  // DPOSE_CORE__DPOSE_COSTMAP__HPP
