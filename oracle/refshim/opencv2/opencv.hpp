// opencv2/opencv.hpp — minimal stand-in for the slice of OpenCV that the reference's dpose_core uses.
//
// TEST INFRASTRUCTURE ONLY (see oracle/refshim/README.md).  OpenCV's C++ headers are not installed in this
// image.  The image arithmetic is delegated to the oracle's restatements of the same OpenCV functions
// (oracle/dpose_oracle.c: cv::polylines, cv::fillPoly, cv::distanceTransform; pinned bit-exact against the
// cv2 4.13.0 wheel by tests/golden), so what oracle/_ref adds over the oracle for cost_data is the
// reference's own control flow around them (dpose_core.cpp:55-139), not a second source of pixels.
#ifndef DPOSE_REFSHIM_OPENCV
#define DPOSE_REFSHIM_OPENCV

#include <cmath>
#include <cstdint>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "../../dpose_oracle.h"

namespace cv {

enum { CV_8U_ = 0, CV_32F_ = 5 };
#ifndef CV_8U
#define CV_8U 0
#define CV_32F 5
#endif
enum { DIST_L2 = 2, DIST_MASK_PRECISE = 0 };

template <typename T>
struct DataType;
template <>
struct DataType<uint8_t> {
  enum { type = CV_8U_ };
};
template <>
struct DataType<float> {
  enum { type = CV_32F_ };
};

struct Point2i {
  Point2i() = default;
  Point2i(int _x, int _y) : x(_x), y(_y) {}
  int x = 0, y = 0;
};

struct Size {
  Size() = default;
  Size(int w, int h) : width(w), height(h) {}
  int width = 0, height = 0;
  Size operator+(const Size& o) const { return Size(width + o.width, height + o.height); }
  bool operator==(const Size& o) const { return width == o.width && height == o.height; }
  bool operator!=(const Size& o) const { return !(*this == o); }
};

/// cv::Mat::size: the dimensions, comparable
struct MatSize {
  int rows = 0, cols = 0;
  bool operator==(const MatSize& o) const { return rows == o.rows && cols == o.cols; }
  bool operator!=(const MatSize& o) const { return !(*this == o); }
};

struct Rect {
  int x = 0, y = 0, width = 0, height = 0;
  bool empty() const { return width <= 0 || height <= 0; }
  Size size() const { return Size(width, height); }
};

struct Scalar {
  Scalar(double v = 0) : v0(v) {}
  double v0;
};

/// reference-counted row-major image of uint8 or float, like cv::Mat (copies share the pixels)
class Mat {
public:
  Mat() = default;
  Mat(Size s, int type) : Mat(s.height, s.width, type) {}
  Mat(int r, int c, int type) : rows(r), cols(c), type_(type) {
    size.rows = r, size.cols = c;
    buf_ = std::make_shared<std::vector<uint8_t>>((size_t)rows * cols * elem());
  }
  Mat(Size s, int type, const Scalar& v) : Mat(s, type) { setTo(v); }

  int rows = 0, cols = 0;
  MatSize size;
  int type() const { return type_; }
  size_t elem() const { return type_ == CV_32F_ ? sizeof(float) : 1; }
  size_t total() const { return (size_t)rows * cols; }
  bool empty() const { return total() == 0; }

  template <typename T>
  T* ptr() {
    return reinterpret_cast<T*>(buf_->data());
  }
  template <typename T>
  const T* ptr() const {
    return reinterpret_cast<const T*>(buf_->data());
  }
  template <typename T>
  T* ptr(int r) {
    return ptr<T>() + (size_t)r * cols;
  }
  template <typename T>
  const T* ptr(int r) const {
    return ptr<T>() + (size_t)r * cols;
  }
  template <typename T>
  const T& at(int r, int c) const {
    return ptr<T>()[(size_t)r * cols + c];
  }
  template <typename T>
  T& at(int r, int c) {
    return ptr<T>()[(size_t)r * cols + c];
  }

  Mat& setTo(const Scalar& v) {
    if (type_ == CV_32F_)
      for (size_t i = 0; i < total(); ++i) ptr<float>()[i] = (float)v.v0;
    else
      std::memset(buf_->data(), (int)v.v0, total());
    return *this;
  }
  /// float image times a double: cv::Mat::operator*= converts the factor to the image depth
  Mat& operator*=(double a) {
    const float f = (float)a;
    for (size_t i = 0; i < total(); ++i) ptr<float>()[i] = ptr<float>()[i] * f;
    return *this;
  }
  Mat& operator-=(double a) {
    const float f = (float)a;
    for (size_t i = 0; i < total(); ++i) ptr<float>()[i] = ptr<float>()[i] - f;
    return *this;
  }

private:
  int type_ = CV_8U_;
  std::shared_ptr<std::vector<uint8_t>> buf_;
};

inline Mat
operator*(const Mat& m, double) {
  return m;  // only inside assert(cv::imwrite(...)) of the reference
}

inline bool
imwrite(const std::string&, const Mat&) {
  return true;
}

inline int
countNonZero(const Mat& m) {
  int n = 0;
  for (size_t i = 0; i < m.total(); ++i) n += m.type() == CV_32F_ ? (m.ptr<float>()[i] != 0.f) : (m.ptr<uint8_t>()[i] != 0);
  return n;
}

inline Rect
boundingRect(const std::vector<Point2i>& pts) {
  Rect r;
  if (pts.empty()) return r;
  int x0 = pts[0].x, x1 = x0, y0 = pts[0].y, y1 = y0;
  for (const auto& p : pts) {
    x0 = p.x < x0 ? p.x : x0, x1 = p.x > x1 ? p.x : x1;
    y0 = p.y < y0 ? p.y : y0, y1 = p.y > y1 ? p.y : y1;
  }
  r.x = x0, r.y = y0, r.width = x1 - x0 + 1, r.height = y1 - y0 + 1;  // inclusive pixel bounds
  return r;
}

namespace detail {
inline std::vector<int32_t>
flat(const std::vector<Point2i>& pts) {
  std::vector<int32_t> xy;
  xy.reserve(2 * pts.size());
  for (const auto& p : pts) xy.push_back(p.x), xy.push_back(p.y);
  return xy;
}
inline int
reflect101(int i, int n) {
  if (n == 1) return 0;
  while (i < 0 || i >= n) i = i < 0 ? -i : 2 * (n - 1) - i;
  return i;
}
}  // namespace detail

inline void
polylines(Mat& img, const std::vector<Point2i>& pts, bool /*closed*/, const Scalar& color) {
  const auto xy = detail::flat(pts);
  oracle_cv_polylines_closed(img.ptr<uint8_t>(), img.cols, img.rows, xy.data(), (int)pts.size(), (uint8_t)color.v0);
}

inline void
fillPoly(Mat& img, const std::vector<std::vector<Point2i>>& contours, const Scalar& color) {
  for (const auto& c : contours) {
    const auto xy = detail::flat(c);
    oracle_cv_fillpoly(img.ptr<uint8_t>(), img.cols, img.rows, xy.data(), (int)c.size(), (uint8_t)color.v0);
  }
}

inline void
distanceTransform(const Mat& src, Mat& dst, int /*DIST_L2*/, int /*DIST_MASK_PRECISE*/, int /*CV_32F*/) {
  std::vector<int32_t> d2(src.total());
  oracle_cv_edt_sq(src.ptr<uint8_t>(), src.cols, src.rows, d2.data());
  dst = Mat(Size(src.cols, src.rows), CV_32F_);
  for (size_t i = 0; i < d2.size(); ++i) dst.ptr<float>()[i] = std::sqrt((float)d2[i]);
}

inline void
copyTo(const Mat& src, Mat& dst, const Mat& mask) {
  for (size_t i = 0; i < src.total(); ++i)
    if (mask.ptr<uint8_t>()[i] != 0) dst.ptr<float>()[i] = src.ptr<float>()[i];
}

inline void
minMaxIdx(const Mat& m, double* min_val) {
  float mn = m.ptr<float>()[0];
  for (size_t i = 1; i < m.total(); ++i) mn = m.ptr<float>()[i] < mn ? m.ptr<float>()[i] : mn;
  *min_val = mn;
}

/// 3x3, sigma 0: separable [1 2 1]/4, row pass then column pass, BORDER_REFLECT_101
inline void
GaussianBlur(const Mat& src, Mat& dst, Size /*3x3*/, double /*sigma*/) {
  const int rows = src.rows, cols = src.cols;
  std::vector<float> in(src.ptr<float>(), src.ptr<float>() + src.total()), tmp(src.total());
  for (int y = 0; y < rows; ++y)
    for (int x = 0; x < cols; ++x) {
      const float s = in[(size_t)y * cols + detail::reflect101(x - 1, cols)] + in[(size_t)y * cols + detail::reflect101(x + 1, cols)];
      tmp[(size_t)y * cols + x] = s * 0.25f + in[(size_t)y * cols + x] * 0.5f;
    }
  if (dst.total() != src.total()) dst = Mat(Size(cols, rows), CV_32F_);
  for (int y = 0; y < rows; ++y)
    for (int x = 0; x < cols; ++x) {
      const float s = tmp[(size_t)detail::reflect101(y - 1, rows) * cols + x] + tmp[(size_t)detail::reflect101(y + 1, rows) * cols + x];
      dst.ptr<float>()[(size_t)y * cols + x] = s * 0.25f + tmp[(size_t)y * cols + x] * 0.5f;
    }
}

}  // namespace cv

#endif  // DPOSE_REFSHIM_OPENCV
