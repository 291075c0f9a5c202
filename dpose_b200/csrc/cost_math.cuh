// cost_math.cuh — geometry shared by the get_cost kernels (host + device).
#pragma once
#include "common.cuh"

namespace dpose {

// ---- shared geometry (host + device) ---------------------------------------------------
struct TableGeom {
  int rows, cols;
  double cx, cy;  // center
};

// axis-aligned bounding box (inclusive integer cells) of the valid kernel region
// [1.5, cols-1.5) x [1.5, rows-1.5) mapped into the map by p = t + R(theta)(k - C)
__host__ __device__ inline WindowRect pose_window(const TableGeom &g, double tx, double ty, double c,
                                                  double s, int size_x, int size_y) {
  const double au0 = 1.5 - g.cx, au1 = (double)g.cols - 1.5 - g.cx;
  const double av0 = 1.5 - g.cy, av1 = (double)g.rows - 1.5 - g.cy;
  const double xa = c * au0, xb = c * au1, xc = -s * av0, xd = -s * av1;
  const double ya = s * au0, yb = s * au1, yc = c * av0, yd = c * av1;
  const double eps = 1e-7;
  const double xmin = tx + fmin(xa, xb) + fmin(xc, xd) - eps, xmax = tx + fmax(xa, xb) + fmax(xc, xd) + eps;
  const double ymin = ty + fmin(ya, yb) + fmin(yc, yd) - eps, ymax = ty + fmax(ya, yb) + fmax(yc, yd) + eps;
  WindowRect w;
  w.x0 = w.y0 = 0;
  w.x1 = w.y1 = -1;  // empty; also the answer for NaN poses and windows off the map
  if (!(xmax >= 0.0) || !(ymax >= 0.0) || !(xmin <= (double)size_x - 1.0) || !(ymin <= (double)size_y - 1.0))
    return w;
  // clamped in floating point, so the int casts cannot overflow
  w.x0 = (int)ceil(fmax(xmin, 0.0));
  w.y0 = (int)ceil(fmax(ymin, 0.0));
  w.x1 = (int)floor(fmin(xmax, (double)size_x - 1.0));
  w.y1 = (int)floor(fmin(ymax, (double)size_y - 1.0));
  return w;
}

}  // namespace dpose
