// dpose_core.hpp — the dpose_core C++ surface on top of the B200 C ABI (include/dpose_b200.h).
//
// Same namespace, type names, member names, argument meaning and exceptions as upstream
// dpose_core/src/dpose_core/dpose_core.hpp:36-167, so that dpose_goal_tolerance (its `problem`
// class and Ipopt loop, dpose_goal_tolerance.cpp:90-110) and the reference's tests recompile against
// this header unchanged.  Every member forwards to libdpose_b200.so; nothing is computed on the CPU
// here and there is no fallback: without a CUDA device the constructors throw std::runtime_error
// carrying the library's message.
//
// With Eigen / OpenCV installed the real Eigen::Matrix / cv::Mat types are used (define nothing);
// without them (this image) the minimal stand-ins of detail/standin.hpp take their place.
//
// Additions over upstream (the batched entry point the north star asks for):
//   device_costmap                     device-resident copy of a costmap (dpose_map_*)
//   pose_gradient::get_cost(map, poses, costs, jacobians)   one call for n poses on one GPU
//   pose_gradient_pool                 the same, sharded over the GPUs of one box
#ifndef DPOSE_CORE__DPOSE_CORE__HPP
#define DPOSE_CORE__DPOSE_CORE__HPP

#if defined(__has_include)
#if __has_include(<Eigen/Dense>)
#define DPOSE_HAVE_EIGEN 1
#endif
#if __has_include(<opencv2/opencv.hpp>)
#define DPOSE_HAVE_OPENCV 1
#endif
#if __has_include(<costmap_2d/layered_costmap.h>)
#define DPOSE_HAVE_COSTMAP_2D 1
#endif
#endif
#ifndef DPOSE_HAVE_EIGEN
#define DPOSE_HAVE_EIGEN 0
#endif
#ifndef DPOSE_HAVE_OPENCV
#define DPOSE_HAVE_OPENCV 0
#endif
#ifndef DPOSE_HAVE_COSTMAP_2D
#define DPOSE_HAVE_COSTMAP_2D 0
#endif

#if DPOSE_HAVE_OPENCV
#include <opencv2/opencv.hpp>
#endif
#if DPOSE_HAVE_EIGEN
#include <Eigen/Dense>
#endif
#include "detail/standin.hpp"

#include <array>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../dpose_b200.h"

namespace dpose_core {

/// closed rectangle ring: 2 x 5, first row x, second row y (upstream dpose_core.hpp:39-40)
template <typename _T>
using rectangle = Eigen::Matrix<_T, 2, 5>;

/// ring (0,0),(w,0),(w,h),(0,h),(0,0) (upstream :45-54)
template <typename _T>
inline rectangle<_T>
to_rectangle(const _T& w, const _T& h) noexcept {
  rectangle<_T> box;
  const _T xs[5] = {0, w, w, 0, 0}, ys[5] = {0, 0, h, h, 0};
  for (int i = 0; i < 5; ++i) {
    box(0, i) = xs[i];
    box(1, i) = ys[i];
  }
  return box;
}

/// ring through (_min, _max) (upstream :59-69)
template <typename _T>
inline rectangle<_T>
to_rectangle(const Eigen::Matrix<_T, 2, 1>& _min, const Eigen::Matrix<_T, 2, 1>& _max) noexcept {
  rectangle<_T> box;
  const _T xs[5] = {_min.x(), _max.x(), _max.x(), _min.x(), _min.x()};
  const _T ys[5] = {_min.y(), _min.y(), _max.y(), _max.y(), _min.y()};
  for (int i = 0; i < 5; ++i) {
    box(0, i) = xs[i];
    box(1, i) = ys[i];
  }
  return box;
}

/// polygon: first row x, second row y; column-major, i.e. x0,y0,x1,y1,... in memory (upstream :72)
using polygon = Eigen::Matrix<int, 2, Eigen::Dynamic>;
using cell = Eigen::Vector2i;  // 8 contiguous bytes (upstream :75)
using cell_vector = std::vector<cell>;

namespace detail {

inline void
check(int rc) {
  if (rc != DPOSE_OK) throw std::runtime_error(dpose_last_error());
}

struct pg_deleter {
  void operator()(dpose_pg* p) const noexcept { dpose_pg_destroy(p); }
};
struct map_deleter {
  void operator()(dpose_map* p) const noexcept { dpose_map_destroy(p); }
};

static_assert(sizeof(cell) == 2 * sizeof(int32_t), "cell must be two packed int32 (x, y)");

}  // namespace detail

/// device-resident costmap (row-major uint8, index = y * size_x + x; only 254 is lethal)
struct device_costmap {
  device_costmap() = default;
  /// uploads _size_y rows of _size_x bytes; _pitch is the host row stride in bytes
  device_costmap(const uint8_t* _data, uint32_t _size_x, uint32_t _size_y, size_t _pitch, int _device = 0) {
    dpose_map* m = nullptr;
    detail::check(dpose_map_create(_data, _size_x, _size_y, _pitch, _device, &m));
    map_.reset(m, detail::map_deleter());
  }
  /// re-uploads a dirty window; the caller holds the costmap's mutex (upstream dpose_costmap.cpp:138)
  void
  update(const uint8_t* _window, size_t _pitch, uint32_t _x0, uint32_t _y0, uint32_t _w, uint32_t _h) {
    detail::check(dpose_map_update(map_.get(), _window, _pitch, _x0, _y0, _w, _h));
  }
  dpose_map*
  handle() const noexcept {
    return map_.get();
  }

private:
  std::shared_ptr<dpose_map> map_;
};

/// the cost kernel of a footprint: float32 table, center cell, bounding box (upstream :79-102)
struct cost_data {
  cost_data() = default;

  /// rasterises the footprint, exact EDT, scale / shift / blur — all on the GPU
  /// (replaces upstream dpose_core.cpp:55-139)
  cost_data(const polygon& _footprint, size_t _padding) : cost_data(_footprint, _padding, 0) {}

  cost_data(const polygon& _footprint, size_t _padding, int _device) {
    dpose_pg* pg = nullptr;
    detail::check(dpose_pg_create(_footprint.data(), (int)_footprint.cols(), (uint32_t)_padding, _device, &pg));
    pg_.reset(pg, detail::pg_deleter());
    const float* host = nullptr;
    int rows = 0, cols = 0;
    int32_t c[2], b[10];
    detail::check(dpose_pg_table(pg, &host, &rows, &cols, c, b));
    cost = cv::Mat(rows, cols, CV_32F);
    for (int r = 0; r < rows; ++r) std::memcpy(cost.ptr<float>(r), host + (size_t)r * cols, sizeof(float) * cols);
    center = cell(c[0], c[1]);
    for (int i = 0; i < 5; ++i) {
      box(0, i) = b[2 * i];
      box(1, i) = b[2 * i + 1];
    }
  }

  inline const cv::Mat&
  get_data() const noexcept {
    return cost;
  }

  inline const cell&
  get_center() const noexcept {
    return center;
  }

  inline const rectangle<int>&
  get_box() const noexcept {
    return box;
  }

  /// the device-side object (null for a default-constructed cost_data)
  inline dpose_pg*
  handle() const noexcept {
    return pg_.get();
  }

private:
  cv::Mat cost;        ///< cost matrix (host copy of the device table)
  cell center;         ///< center cell
  rectangle<int> box;  ///< the bounding box
  std::shared_ptr<dpose_pg> pg_;
};

/// cost and Jacobian of SE2 poses against lethal cells (upstream :133-167); cell domain, not metric
struct pose_gradient {
  struct parameter {
    unsigned int padding = 2;  ///< padding of the given footprint (in cells)
  };

  using jacobian = Eigen::Vector3d;
  using pose = Eigen::Vector3d;

  /// @throws std::runtime_error if the _footprint has fewer than three points (upstream dpose_core.cpp:152-153)
  pose_gradient(const polygon& _footprint, const parameter& _param) : pose_gradient(_footprint, _param, 0) {}

  pose_gradient(const polygon& _footprint, const parameter& _param, int _device) {
    if (_footprint.cols() < 3) throw std::runtime_error("footprint must contain at least three points");
    data_ = cost_data(_footprint, _param.padding, _device);
  }
  pose_gradient() = default;

  /// cost of _se2 against the cells [_begin, _end), optional Jacobian (d/dx, d/dy, d/dtheta);
  /// replaces upstream dpose_core.cpp:200-260
  double
  get_cost(const pose& _se2, cell_vector::const_iterator _begin, cell_vector::const_iterator _end,
           jacobian* _J) const {
    if (!data_.handle()) throw std::runtime_error("pose_gradient is not initialised");
    const double se2[3] = {_se2(0), _se2(1), _se2(2)};
    const size_t n = (size_t)(_end - _begin);
    double cost = 0.0, J[3] = {0.0, 0.0, 0.0};
    const int32_t* cells = n ? reinterpret_cast<const int32_t*>(&*_begin) : nullptr;
    detail::check(dpose_pg_get_cost_cells(data_.handle(), se2, cells, n, &cost, _J ? J : nullptr));
    if (_J) *_J = jacobian(J[0], J[1], J[2]);
    return cost;
  }

  /// NEW: n poses against every lethal cell of _map in one call.  _poses: n x (x, y, theta);
  /// _out: n x (cost, Jx, Jy, Jtheta).  Host buffers; chunks stream through the GPU.
  void
  get_cost(const device_costmap& _map, const double* _poses, size_t _n, double* _out, bool _jacobian = true) const {
    if (!data_.handle()) throw std::runtime_error("pose_gradient is not initialised");
    detail::check(dpose_pg_get_cost_batch(data_.handle(), _map.handle(), _poses, _n, _out, _jacobian ? 1 : 0));
  }

  /// NEW: the same on DEVICE buffers, asynchronous on _stream (a cudaStream_t), restricted to the map rows
  /// [_row_lo, _row_hi) — only their lethal cells count, only they are re-read of a volatile map (the batched form of
  /// the search box, dpose_goal_tolerance.cpp:119-141).  _row_hi == 0: the whole map.  See get_reach().
  void
  get_cost_device(const device_costmap& _map, const double* _d_poses, size_t _n, double* _d_out, bool _jacobian = true,
                  void* _stream = nullptr, uint32_t _row_lo = 0, uint32_t _row_hi = 0) const {
    if (!data_.handle()) throw std::runtime_error("pose_gradient is not initialised");
    if (_row_hi == 0)
      detail::check(dpose_pg_get_cost_batch_device(data_.handle(), _map.handle(), _d_poses, _n, _d_out, _jacobian ? 1 : 0, _stream));
    else
      detail::check(dpose_pg_get_cost_batch_device_rows(data_.handle(), _map.handle(), _d_poses, _n, _d_out, _jacobian ? 1 : 0,
                                                        _stream, _row_lo, _row_hi));
  }

  /// rows / columns around a pose beyond which no cell contributes to its cost (dpose_pg_reach)
  uint32_t
  get_reach() const {
    if (!data_.handle()) throw std::runtime_error("pose_gradient is not initialised");
    uint32_t r = 0;
    detail::check(dpose_pg_reach(data_.handle(), &r));
    return r;
  }

  /// convenience overload of the batched call on std::vector
  void
  get_cost(const device_costmap& _map, const std::vector<pose>& _poses, std::vector<double>& _costs,
           std::vector<jacobian>* _J) const {
    const size_t n = _poses.size();
    std::vector<double> in(3 * n), out(4 * n);
    for (size_t i = 0; i < n; ++i)
      for (int k = 0; k < 3; ++k) in[3 * i + k] = _poses[i](k);
    get_cost(_map, in.data(), n, out.data(), _J != nullptr);
    _costs.resize(n);
    if (_J) _J->resize(n);
    for (size_t i = 0; i < n; ++i) {
      _costs[i] = out[4 * i];
      if (_J) (*_J)[i] = jacobian(out[4 * i + 1], out[4 * i + 2], out[4 * i + 3]);
    }
  }

  inline const cost_data&
  get_data() const noexcept {
    return data_;
  }

private:
  cost_data data_;
};

/// NEW: one pose_gradient + costmap replica per GPU of the box; pose batches are sharded contiguously
/// (no collective, results land in the caller's buffer) — dpose_pg_get_cost_batch_multi
struct pose_gradient_pool {
  pose_gradient_pool(const polygon& _footprint, const pose_gradient::parameter& _param, const uint8_t* _map,
                     uint32_t _size_x, uint32_t _size_y, size_t _pitch, int _n_devices) {
    for (int d = 0; d < _n_devices; ++d) {
      pgs_.emplace_back(_footprint, _param, d);
      maps_.emplace_back(_map, _size_x, _size_y, _pitch, d);
    }
  }
  void
  get_cost(const double* _poses, size_t _n, double* _out, bool _jacobian = true) const {
    std::vector<dpose_pg*> p;
    std::vector<dpose_map*> m;
    for (size_t d = 0; d < pgs_.size(); ++d) {
      p.push_back(pgs_[d].get_data().handle());
      m.push_back(maps_[d].handle());
    }
    detail::check(dpose_pg_get_cost_batch_multi(p.data(), m.data(), (int)p.size(), _poses, _n, _out, _jacobian ? 1 : 0));
  }

private:
  std::vector<pose_gradient> pgs_;
  std::vector<device_costmap> maps_;
};

}  // namespace dpose_core

#endif  // DPOSE_CORE__DPOSE_CORE__HPP
