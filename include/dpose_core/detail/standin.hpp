// standin.hpp — minimal stand-ins for the third-party types that appear in dpose_core's public
// signatures (Eigen::Matrix, cv::Mat, costmap_2d::Costmap2D / LayeredCostmap).
//
// They are compiled ONLY when the real headers are absent (this image has no Eigen / OpenCV / ROS);
// with the real libraries installed include/dpose_core/dpose_core.hpp uses those instead and the
// dpose_goal_tolerance plugin recompiles unchanged.  Only the members dpose_core's callers touch
// are provided (upstream: dpose_core/src/dpose_core/dpose_core.hpp:27-76, dpose_costmap.hpp:29,
// dpose_costmap.cpp:124-177; dpose_goal_tolerance.cpp:90-141).
#ifndef DPOSE_CORE__DETAIL__STANDIN__HPP
#define DPOSE_CORE__DETAIL__STANDIN__HPP

#include <cstddef>
#include <cstdint>
#include <memory>
#include <mutex>
#include <vector>

#if !DPOSE_HAVE_EIGEN
namespace Eigen {

constexpr int Dynamic = -1;

// Column-major dense matrix, fixed or dynamic number of columns; just enough for 2 x N polygons,
// 2-vectors (cells), 3-vectors (poses / jacobians) and the 2 x 5 rectangle ring.
template <typename T, int R, int C>
class Matrix {
  static_assert(R > 0, "only the column count may be dynamic");

public:
  Matrix() : cols_(C == Dynamic ? 0 : C), d_((size_t)R * (C == Dynamic ? 0 : C), T()) {}
  // two arguments mean (rows, cols) for matrices and (x, y) for 2-vectors, as in Eigen
  template <typename A, typename B>
  Matrix(A a, B b) : cols_(C == 1 ? 1 : (int)b), d_((size_t)R * (C == 1 ? 1 : (int)b), T()) {
    if (C == 1) {
      d_[0] = (T)a;
      d_[1] = (T)b;
    }
  }
  Matrix(T x, T y, T z) : cols_(1), d_{x, y, z} { static_assert(R == 3 && C == 1, "3-vector only"); }

  int rows() const { return R; }
  int cols() const { return cols_; }
  size_t size() const { return d_.size(); }
  T *data() { return d_.data(); }
  const T *data() const { return d_.data(); }
  T &operator()(int r, int c) { return d_[(size_t)c * R + r]; }
  const T &operator()(int r, int c) const { return d_[(size_t)c * R + r]; }
  T &operator()(int i) { return d_[i]; }
  const T &operator()(int i) const { return d_[i]; }
  T &operator[](int i) { return d_[i]; }
  const T &operator[](int i) const { return d_[i]; }
  T &x() { return d_[0]; }
  const T &x() const { return d_[0]; }
  T &y() { return d_[1]; }
  const T &y() const { return d_[1]; }
  T &z() { return d_[2]; }
  const T &z() const { return d_[2]; }
  Matrix<T, R, 1> col(int c) const {
    Matrix<T, R, 1> v;
    for (int r = 0; r < R; ++r) v(r) = (*this)(r, c);
    return v;
  }
  void setZero() {
    for (auto &v : d_) v = T();
  }
  bool operator==(const Matrix &o) const { return cols_ == o.cols_ && d_ == o.d_; }
  bool operator!=(const Matrix &o) const { return !(*this == o); }

private:
  int cols_;
  std::vector<T> d_;
};

// fixed-size 2-vectors must be 8 contiguous bytes (dpose_core passes cell_vector storage as int32 pairs)
template <>
class Matrix<int, 2, 1> {
public:
  Matrix() : d_{0, 0} {}
  Matrix(int x, int y) : d_{x, y} {}
  int rows() const { return 2; }
  int cols() const { return 1; }
  int *data() { return d_; }
  const int *data() const { return d_; }
  int &operator()(int i) { return d_[i]; }
  const int &operator()(int i) const { return d_[i]; }
  int &operator[](int i) { return d_[i]; }
  const int &operator[](int i) const { return d_[i]; }
  int &x() { return d_[0]; }
  const int &x() const { return d_[0]; }
  int &y() { return d_[1]; }
  const int &y() const { return d_[1]; }
  bool operator==(const Matrix &o) const { return d_[0] == o.d_[0] && d_[1] == o.d_[1]; }
  bool operator!=(const Matrix &o) const { return !(*this == o); }

private:
  int d_[2];
};

using Vector2i = Matrix<int, 2, 1>;
using Vector2d = Matrix<double, 2, 1>;
using Vector3d = Matrix<double, 3, 1>;

}  // namespace Eigen
#endif  // !DPOSE_HAVE_EIGEN

#if !DPOSE_HAVE_OPENCV
#ifndef CV_32F
#define CV_32F 5
#endif
namespace cv {

// float32 single-channel image, reference-counted like cv::Mat
class Mat {
public:
  int rows = 0, cols = 0;
  Mat() = default;
  Mat(int r, int c, int type) : rows(r), cols(c), buf_(std::make_shared<std::vector<float>>((size_t)r * c)) {
    (void)type;
  }
  bool empty() const { return rows == 0 || cols == 0; }
  int type() const { return CV_32F; }
  template <typename T>
  T &at(int r, int c) {
    return reinterpret_cast<T *>(buf_->data())[(size_t)r * cols + c];
  }
  template <typename T>
  const T &at(int r, int c) const {
    return reinterpret_cast<const T *>(buf_->data())[(size_t)r * cols + c];
  }
  template <typename T>
  T *ptr(int r = 0) {
    return reinterpret_cast<T *>(buf_->data()) + (size_t)r * cols;
  }
  template <typename T>
  const T *ptr(int r = 0) const {
    return reinterpret_cast<const T *>(buf_->data()) + (size_t)r * cols;
  }

private:
  std::shared_ptr<std::vector<float>> buf_;
};

}  // namespace cv
#endif  // !DPOSE_HAVE_OPENCV

#if !DPOSE_HAVE_COSTMAP_2D
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

// row-major uint8 grid, index = my * size_x + mx, guarded by a recursive mutex
class Costmap2D {
public:
  typedef std::recursive_mutex mutex_t;
  Costmap2D(unsigned int size_x, unsigned int size_y, double resolution, double origin_x, double origin_y,
            unsigned char default_value = 0)
      : size_x_(size_x), size_y_(size_y), resolution_(resolution), origin_x_(origin_x), origin_y_(origin_y),
        map_((size_t)size_x * size_y, default_value), mutex_(new mutex_t()) {}
  unsigned char *getCharMap() const { return const_cast<unsigned char *>(map_.data()); }
  unsigned int getSizeInCellsX() const { return size_x_; }
  unsigned int getSizeInCellsY() const { return size_y_; }
  double getResolution() const { return resolution_; }
  double getOriginX() const { return origin_x_; }
  double getOriginY() const { return origin_y_; }
  unsigned int getIndex(unsigned int mx, unsigned int my) const { return my * size_x_ + mx; }
  unsigned char getCost(unsigned int mx, unsigned int my) const { return map_[getIndex(mx, my)]; }
  void setCost(unsigned int mx, unsigned int my, unsigned char cost) { map_[getIndex(mx, my)] = cost; }
  mutex_t *getMutex() { return mutex_.get(); }

private:
  unsigned int size_x_, size_y_;
  double resolution_, origin_x_, origin_y_;
  std::vector<unsigned char> map_;
  std::shared_ptr<mutex_t> mutex_;
};

class LayeredCostmap {
public:
  LayeredCostmap(Costmap2D *map, std::vector<geometry_msgs::Point> footprint)
      : map_(map), footprint_(std::move(footprint)) {}
  Costmap2D *getCostmap() { return map_; }
  const std::vector<geometry_msgs::Point> &getFootprint() { return footprint_; }

private:
  Costmap2D *map_;
  std::vector<geometry_msgs::Point> footprint_;
};

}  // namespace costmap_2d
#endif  // !DPOSE_HAVE_COSTMAP_2D

#endif  // DPOSE_CORE__DETAIL__STANDIN__HPP
