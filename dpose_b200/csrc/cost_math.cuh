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
// [1.5, cols-1.5) x [1.5, rows-1.5) mapped into the map by p = t + R(theta)(k - C), widened by 1e-7 cell.
// Centre / half-extent form: the box of a rotated rectangle is its rotated centre -+ (|c| hu + |s| hv, |s| hu + |c| hv) —
// 16 FP64 operations and no floating-point min / max (a double fmin is a compare, two selects and NaN fix-ups on the
// GPU; the corner form needed twelve of them per pose).  Conversions saturate, the clamps are integer.
// Rows: the window is clamped to [y_lo, size_y) — callers that restrict a launch to a band of map rows (BatchSpec::row_lo /
// row_hi) pass the band's end as size_y and its start as y_lo; the clamp constants change, the instruction count does not.
__host__ __device__ inline WindowRect pose_window(const TableGeom &g, double tx, double ty, double c,
                                                  double s, int size_x, int size_y, int y_lo = 0) {
  const double hu = 0.5 * ((double)g.cols - 3.0), hv = 0.5 * ((double)g.rows - 3.0);  // half extents of the valid region
  const double mu = 0.5 * (double)g.cols - g.cx, mv = 0.5 * (double)g.rows - g.cy;    // its centre relative to C
  const double eps = 1e-7;
  const double px = tx + fma(c, mu, -s * mv), py = ty + fma(s, mu, c * mv);
  const double hx = fma(fabs(c), hu, fabs(s) * hv) + eps, hy = fma(fabs(s), hu, fabs(c) * hv) + eps;
  const double xmin = px - hx, xmax = px + hx, ymin = py - hy, ymax = py + hy;
  WindowRect w;
  w.x0 = w.y0 = 0;
  w.x1 = w.y1 = -1;  // empty; also the answer for NaN poses and windows off the map
  const double t = xmin + ymin;
  if (t != t) return w;  // a NaN anywhere in the pose (or opposite infinities) ends up here
#ifdef __CUDA_ARCH__
  // cvt.rpi / cvt.rmi saturate to the int range, so windows far off the map come out empty after the clamps
  w.x0 = max(__double2int_ru(xmin), 0);
  w.y0 = max(__double2int_ru(ymin), y_lo);
  w.x1 = min(__double2int_rd(xmax), size_x - 1);
  w.y1 = min(__double2int_rd(ymax), size_y - 1);
#else
  // the same values on the host (clamped in floating point, so the int casts cannot overflow)
  w.x0 = (int)ceil(fmin(fmax(xmin, 0.0), 2147483647.0));
  w.y0 = (int)ceil(fmin(fmax(ymin, (double)y_lo), 2147483647.0));
  w.x1 = (int)floor(fmax(fmin(xmax, (double)size_x - 1.0), -2147483648.0));
  w.y1 = (int)floor(fmax(fmin(ymax, (double)size_y - 1.0), -2147483648.0));
#endif
  return w;
}

#ifdef __CUDACC__
// ---- bit walking (device) --------------------------------------------------------------------------------
// index of the most significant set bit (-1 for 0): one FLO, no bit reversal
__device__ __forceinline__ int top_bit(unsigned x) {
  int r;
  asm("bfind.u32 %0, %1;" : "=r"(r) : "r"(x));
  return r;
}
// m without its most significant set bit, b = top_bit(m) (b = -1 for m == 0: m stays 0): AND with a mask of the b low bits
// (BMSK) — no materialised 1 to shift
__device__ __forceinline__ unsigned clear_top(unsigned m, int b) {
  unsigned r;
  asm("bmsk.clamp.b32 %0, 0, %1;" : "=r"(r) : "r"(b));
  return m & r;
}
// 1 + x / 32 for 0 <= x < 32 as a double without the conversion pipe: x is planted into the top five mantissa bits of 1.0 by
// one integer instruction on the high word.  `zero` must be a 0 the compiler cannot see through (it becomes the low word;
// a literal is re-materialised in every loop iteration).  A cell walk that needs c * x + s * y uses
//   (32 c) X + (32 s) Y - 32 c - 32 s,  X = unit_plus_32nd(x), Y = unit_plus_32nd(y)
// with the constants folded into its origin (32 c, 32 s exact).  I2F.F64 runs on the 4-lane conversion pipe, 8 issue
// cycles per warp instruction — as many as four DFMAs.
__device__ __forceinline__ double unit_plus_32nd(int x, int zero) { return __hiloint2double(0x3FF00000 + (x << 15), zero); }
#endif

}  // namespace dpose
