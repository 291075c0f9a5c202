// bridge.cpp — extern "C" access to the reference's own dpose_core classes, for oracle/_ref.
//
// TEST INFRASTRUCTURE ONLY (see README.md in this directory).  This file is compiled together with the
// reference's dpose_core.cpp and dpose_costmap.cpp (taken from /root/reference where they lie, never
// copied) against the stand-in headers next to it.  Every function below only marshals plain buffers into
// the reference's types and calls the reference function named in its comment.
#include <dpose_core/dpose_costmap.hpp>

#include <cmath>
#include <cstdlib>
#include <cstring>
#include <stdexcept>

using dpose_core::cell;
using dpose_core::cell_vector;
using dpose_core::polygon;
using dpose_core::pose_gradient;

namespace {

polygon
to_polygon(const int32_t* xy, int n) {
  polygon p(2, n);
  for (int i = 0; i < n; ++i) {
    p(0, i) = xy[2 * i];
    p(1, i) = xy[2 * i + 1];
  }
  return p;
}

dpose_core::rectangle<int>
to_ring(const int32_t* xy) {
  dpose_core::rectangle<int> r;
  for (int i = 0; i < 5; ++i) {
    r(0, i) = xy[2 * i];
    r(1, i) = xy[2 * i + 1];
  }
  return r;
}

costmap_2d::Costmap2D
view(const uint8_t* map, uint32_t sx, uint32_t sy) {
  return costmap_2d::Costmap2D(const_cast<unsigned char*>(map), sx, sy);
}

}  // namespace

extern "C" {

/// pose_gradient::pose_gradient (dpose_core.cpp:148-156); NULL if the reference throws
void*
ref_pg_create(const int32_t* fp_xy, int n, unsigned padding) {
  try {
    pose_gradient::parameter param;
    param.padding = padding;
    return new pose_gradient(to_polygon(fp_xy, n), param);
  } catch (const std::exception&) {
    return nullptr;
  }
}

void
ref_pg_free(void* pg) {
  delete static_cast<pose_gradient*>(pg);
}

/// cost_data::get_data / get_center / get_box (dpose_core.hpp:83-96)
void
ref_pg_info(const void* pg, int* rows, int* cols, int32_t center[2], int32_t box[10]) {
  const auto& d = static_cast<const pose_gradient*>(pg)->get_data();
  *rows = d.get_data().rows;
  *cols = d.get_data().cols;
  center[0] = d.get_center().x();
  center[1] = d.get_center().y();
  for (int i = 0; i < 5; ++i) {
    box[2 * i] = d.get_box()(0, i);
    box[2 * i + 1] = d.get_box()(1, i);
  }
}

void
ref_pg_table(const void* pg, float* out) {
  const cv::Mat& m = static_cast<const pose_gradient*>(pg)->get_data().get_data();
  for (int r = 0; r < m.rows; ++r)
    for (int c = 0; c < m.cols; ++c) out[(size_t)r * m.cols + c] = m.at<float>(r, c);
}

/// pose_gradient::get_cost (dpose_core.cpp:200-260)
double
ref_get_cost(const void* pg, const double se2[3], const int32_t* cells_xy, size_t n, double* J) {
  cell_vector cells(n);
  for (size_t i = 0; i < n; ++i) cells[i] = cell(cells_xy[2 * i], cells_xy[2 * i + 1]);
  pose_gradient::jacobian jac;
  const double cost = static_cast<const pose_gradient*>(pg)->get_cost(pose_gradient::pose(se2[0], se2[1], se2[2]),
                                                                      cells.cbegin(), cells.cend(), J ? &jac : nullptr);
  if (J) J[0] = jac.x(), J[1] = jac.y(), J[2] = jac.z();
  return cost;
}

/// make_footprint (dpose_costmap.cpp:40-55); -1 if the reference throws
int
ref_make_footprint(const double* xy_m, int n, double res, int32_t* xy_cells) {
  std::vector<geometry_msgs::Point> fp((size_t)n);
  for (int i = 0; i < n; ++i) fp[i].x = xy_m[2 * i], fp[i].y = xy_m[2 * i + 1];
  unsigned char none = 0;
  costmap_2d::Costmap2D map(&none, 1, 1, res);
  costmap_2d::LayeredCostmap lc(&map, fp);
  try {
    const polygon p = dpose_core::make_footprint(lc);
    for (int i = 0; i < n; ++i) xy_cells[2 * i] = p(0, i), xy_cells[2 * i + 1] = p(1, i);
    return 0;
  } catch (const std::runtime_error&) {
    return -1;
  }
}

/// raytrace (dpose_costmap.cpp:57-104); writes min(cap, size) cells, returns size
size_t
ref_raytrace(const int32_t b[2], const int32_t e[2], int32_t* out_xy, size_t cap) {
  const cell_vector ray = dpose_core::raytrace(cell(b[0], b[1]), cell(e[0], e[1]));
  for (size_t i = 0; i < ray.size() && i < cap; ++i) out_xy[2 * i] = ray[i].x(), out_xy[2 * i + 1] = ray[i].y();
  return ray.size();
}

/// lethal_cells_within (dpose_costmap.cpp:117-184); *out is malloc'ed (ref_free), order as the reference returns it
size_t
ref_lethal_cells_within(const uint8_t* map, uint32_t sx, uint32_t sy, const int32_t box[10], int32_t** out) {
  costmap_2d::Costmap2D cm = view(map, sx, sy);
  const cell_vector cells = dpose_core::lethal_cells_within(cm, to_ring(box));
  *out = static_cast<int32_t*>(std::malloc(sizeof(int32_t) * 2 * (cells.size() ? cells.size() : 1)));
  for (size_t i = 0; i < cells.size(); ++i) (*out)[2 * i] = cells[i].x(), (*out)[2 * i + 1] = cells[i].y();
  return cells.size();
}

void
ref_free(void* p) {
  std::free(p);
}

/// is_inside (dpose_costmap.cpp:186-199)
int
ref_is_inside(uint32_t sx, uint32_t sy, const int32_t* fp_xy, int n) {
  unsigned char none = 0;
  costmap_2d::Costmap2D cm(&none, sx, sy);
  return dpose_core::is_inside(cm, to_polygon(fp_xy, n)) ? 1 : 0;
}

/// check_footprint (dpose_costmap.cpp:460-582)
int
ref_check_footprint(const uint8_t* map, uint32_t sx, uint32_t sy, const int32_t* fp_xy, int n, uint8_t cost) {
  costmap_2d::Costmap2D cm = view(map, sx, sy);
  return dpose_core::check_footprint(cm, to_polygon(fp_xy, n), cost) ? 1 : 0;
}

/// x_bresenham (dpose_costmap.cpp:201-317): get_x() before every one of `steps` advance_x() calls, and get_dx()
void
ref_xb_trace(const int32_t c0[2], const int32_t c1[2], int steps, int32_t* xs, double* dx) {
  dpose_core::x_bresenham xb(cell(c0[0], c0[1]), cell(c1[0], c1[1]));
  *dx = xb.get_dx();
  for (int i = 0; i < steps; ++i) {
    xs[i] = xb.get_x();
    xb.advance_x();
  }
}

/// The BASELINE batch workload on the reference's functions: per pose, the kernel box rotated into the map
/// (test/lethal_cells_within.cpp:82-86), lethal_cells_within, get_cost.  out = n x (cost, Jx, Jy, Jtheta).
void
ref_get_cost_batch(const void* pg_, const uint8_t* map, uint32_t sx, uint32_t sy, const double* poses, size_t n,
                   double* out, int want_jacobian, int nthreads) {
  const pose_gradient* pg = static_cast<const pose_gradient*>(pg_);
  const auto box = pg->get_data().get_box();
#ifdef _OPENMP
#pragma omp parallel num_threads(nthreads > 1 ? nthreads : 1)
#endif
  {
    costmap_2d::Costmap2D cm = view(map, sx, sy);  // one per thread: its mutex is taken per call
#ifdef _OPENMP
#pragma omp for schedule(dynamic, 256)
#endif
    for (long long i = 0; i < (long long)n; ++i) {
      const double* p = poses + 3 * i;
      const Eigen::Isometry2d t = Eigen::Translation2d(p[0], p[1]) * Eigen::Rotation2Dd(p[2]);
      dpose_core::rectangle<int> ring;
      for (int c = 0; c < 5; ++c) {
        const Eigen::Vector2d v = t * Eigen::Vector2d(box(0, c), box(1, c));
        ring(0, c) = (int)std::round(v.x());
        ring(1, c) = (int)std::round(v.y());
      }
      const cell_vector cells = dpose_core::lethal_cells_within(cm, ring);
      pose_gradient::jacobian J = pose_gradient::jacobian::Zero();
      out[4 * i] = pg->get_cost(pose_gradient::pose(p[0], p[1], p[2]), cells.cbegin(), cells.cend(),
                                want_jacobian ? &J : nullptr);
      out[4 * i + 1] = J.x(), out[4 * i + 2] = J.y(), out[4 * i + 3] = J.z();
    }
  }
  (void)nthreads;
}

}  // extern "C"
