/*
 * dpose_oracle.c — CPU restatement of dpose_core's footprint-cost path (plain C11).
 *
 * TEST INFRASTRUCTURE ONLY (see dpose_oracle.h).  Parity pin: the image pipeline is
 * checked bit-for-bit against OpenCV (cv2 4.13.0, IPP off) by tests/test_oracle.py
 * and frozen in tests/golden/; get_cost is checked against the known-answer vectors of
 * SURVEY.md §8(c) and the reference's own Jacobian self-consistency test.
 *
 * The arithmetic of the image pipeline lives in OpenCV imgproc, which is not vendored
 * by the reference (package.xml: libopencv-dev of ROS noetic).  The rules restated here
 * (LineIterator, FillEdgeCollection, trueDistTrans, cvtScale, GaussianBlur) are OpenCV's
 * published algorithms, anchored on the reference's call sites dpose_core.cpp:61-111.
 */
#include "dpose_oracle.h"

#include <limits.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

int oracle_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* ------------------------------------------------------------------------- */
/* make_footprint — dpose_costmap.cpp:40-55                                   */
/* ------------------------------------------------------------------------- */
int oracle_make_footprint(const double *xy_m, int n, double res, int32_t *xy_cells) {
  if (res <= 0) return -1; /* :44-45 throws "resolution must be positive" */
  for (int i = 0; i < n; ++i) {
    xy_cells[2 * i + 0] = (int32_t)round(xy_m[2 * i + 0] / res); /* :51 std::round */
    xy_cells[2 * i + 1] = (int32_t)round(xy_m[2 * i + 1] / res); /* :52 */
  }
  return 0;
}

/* ------------------------------------------------------------------------- */
/* cv::polylines(img, pts, closed, 0) — call site dpose_core.cpp:73           */
/* OpenCV drawing.cpp Line(): 8-connected LineIterator with leftToRight=true.  */
/* ------------------------------------------------------------------------- */
static void cv_line8(uint8_t *img, int cols, int rows, int x1, int y1, int x2, int y2,
                     uint8_t color) {
  int dx = x2 - x1, dy = y2 - y1;
  int step_x = 1, step_y = 1;
  if (dx < 0) { /* leftToRight: start from the left end point */
    dx = -dx;
    dy = -dy;
    x1 = x2;
    y1 = y2;
  }
  if (dy < 0) {
    dy = -dy;
    step_y = -1;
  }
  const int vert = dy > dx;
  int major = vert ? dy : dx, minor = vert ? dx : dy;
  int err = major - 2 * minor;
  int x = x1, y = y1;
  for (int i = 0; i <= major; ++i) {
    if (x >= 0 && x < cols && y >= 0 && y < rows) img[(size_t)y * cols + x] = color;
    /* err < 0 -> take the diagonal step, else the axis step */
    const int diag = err < 0;
    err += -2 * minor + (diag ? 2 * major : 0);
    if (vert) {
      y += step_y;
      if (diag) x += step_x;
    } else {
      x += step_x;
      if (diag) y += step_y;
    }
  }
}

static void cv_polylines_closed(uint8_t *img, int cols, int rows, const int32_t *xy, int n,
                                uint8_t color) {
  /* PolyLine(): p0 = v[n-1]; for each i: line(p0, v[i]); p0 = v[i] */
  int px = xy[2 * (n - 1)], py = xy[2 * (n - 1) + 1];
  for (int i = 0; i < n; ++i) {
    cv_line8(img, cols, rows, px, py, xy[2 * i], xy[2 * i + 1], color);
    px = xy[2 * i];
    py = xy[2 * i + 1];
  }
}

/* ------------------------------------------------------------------------- */
/* cv::fillPoly(img, {pts}, 0) — call site dpose_core.cpp:94                   */
/* OpenCV drawing.cpp CollectPolyEdges + FillEdgeCollection (non-AA, shift 0): */
/* outline via Line(), then even-odd scanline spans from 16.16 edge walking.   */
/* ------------------------------------------------------------------------- */
typedef struct {
  int y0, y1;
  int64_t x, dx;
} poly_edge;

static int cmp_edges(const void *a, const void *b) {
  const poly_edge *e1 = (const poly_edge *)a, *e2 = (const poly_edge *)b;
  if (e1->y0 != e2->y0) return e1->y0 < e2->y0 ? -1 : 1;
  if (e1->x != e2->x) return e1->x < e2->x ? -1 : 1;
  if (e1->dx != e2->dx) return e1->dx < e2->dx ? -1 : 1;
  return 0;
}

static int cmp_i64(const void *a, const void *b) {
  const int64_t x = *(const int64_t *)a, y = *(const int64_t *)b;
  return x < y ? -1 : (x > y ? 1 : 0);
}

static void cv_fillpoly(uint8_t *img, int cols, int rows, const int32_t *xy, int n,
                        uint8_t color) {
  enum { XY_SHIFT = 16 };
  poly_edge *edges = (poly_edge *)malloc(sizeof(poly_edge) * (size_t)(n > 0 ? n : 1));
  int64_t *xs = (int64_t *)malloc(sizeof(int64_t) * (size_t)(n > 0 ? n : 1));
  int total = 0;
  int y_min = INT_MAX, y_max = INT_MIN;
  int px = xy[2 * (n - 1)], py = xy[2 * (n - 1) + 1];
  for (int i = 0; i < n; ++i) {
    const int qx = xy[2 * i], qy = xy[2 * i + 1];
    cv_line8(img, cols, rows, px, py, qx, qy, color);
    if (py != qy) {
      poly_edge e;
      e.dx = (((int64_t)qx - px) * (1 << XY_SHIFT)) / ((int64_t)qy - py); /* trunc toward 0 */
      if (py < qy) {
        e.y0 = py;
        e.y1 = qy;
        e.x = (int64_t)px * (1 << XY_SHIFT);
      } else {
        e.y0 = qy;
        e.y1 = py;
        e.x = (int64_t)qx * (1 << XY_SHIFT);
      }
      edges[total++] = e;
      if (e.y0 < y_min) y_min = e.y0;
      if (e.y1 > y_max) y_max = e.y1;
    }
    px = qx;
    py = qy;
  }
  if (total >= 2) {
    qsort(edges, (size_t)total, sizeof(poly_edge), cmp_edges);
    if (y_max > rows) y_max = rows;
    for (int y = y_min; y < y_max; ++y) {
      int na = 0;
      for (int i = 0; i < total; ++i)
        if (edges[i].y0 <= y && y < edges[i].y1)
          xs[na++] = edges[i].x + (int64_t)(y - edges[i].y0) * edges[i].dx;
      qsort(xs, (size_t)na, sizeof(int64_t), cmp_i64);
      if (y < 0) continue;
      for (int i = 0; i + 1 < na; i += 2) {
        /* left end rounded to nearest, right end floored (cv2 4.13 behaviour,
         * probed 0 mismatches on random polygons; the outline drawn above hides
         * the difference to a ceil'ed left end) */
        int x1 = (int)((xs[i] + (1 << (XY_SHIFT - 1))) >> XY_SHIFT);
        int x2 = (int)(xs[i + 1] >> XY_SHIFT);
        if (x1 < cols && x2 >= 0) {
          if (x1 < 0) x1 = 0;
          if (x2 >= cols) x2 = cols - 1;
          for (int x = x1; x <= x2; ++x) img[(size_t)y * cols + x] = color;
        }
      }
    }
  }
  free(xs);
  free(edges);
}

/* ------------------------------------------------------------------------- */
/* cv::distanceTransform(DIST_L2, DIST_MASK_PRECISE, CV_32F) — :82-83          */
/* exact squared Euclidean distance to the nearest zero pixel (definition).    */
/* ------------------------------------------------------------------------- */
static void exact_edt_sq(const uint8_t *img, int cols, int rows, int32_t *d2) {
  /* separable exact EDT (column pass then exhaustive row pass): O(rows*cols*cols) */
  const int32_t INF = 1 << 29;
  int32_t *g = (int32_t *)malloc(sizeof(int32_t) * (size_t)rows * cols);
  for (int x = 0; x < cols; ++x) {
    int32_t last = -1;
    for (int y = 0; y < rows; ++y) {
      if (img[(size_t)y * cols + x] == 0) last = y;
      g[(size_t)y * cols + x] = last < 0 ? INF : (y - last);
    }
    last = -1;
    for (int y = rows - 1; y >= 0; --y) {
      if (img[(size_t)y * cols + x] == 0) last = y;
      if (last >= 0 && (last - y) < g[(size_t)y * cols + x]) g[(size_t)y * cols + x] = last - y;
    }
  }
  for (int y = 0; y < rows; ++y) {
    const int32_t *gr = g + (size_t)y * cols;
    for (int x = 0; x < cols; ++x) {
      int64_t best = (int64_t)INF * 4;
      for (int i = 0; i < cols; ++i) {
        if (gr[i] >= INF) continue;
        const int64_t v = (int64_t)(x - i) * (x - i) + (int64_t)gr[i] * gr[i];
        if (v < best) best = v;
      }
      d2[(size_t)y * cols + x] = best > INT_MAX ? INT_MAX : (int32_t)best;
    }
  }
  free(g);
}

/* The three OpenCV image primitives above, exported for oracle/refshim/opencv2 (the stand-in
 * cv:: layer the reference's own dpose_core.cpp is compiled against in oracle/_ref). */
void oracle_cv_polylines_closed(uint8_t *img, int cols, int rows, const int32_t *xy, int n, uint8_t color) {
  cv_polylines_closed(img, cols, rows, xy, n, color);
}
void oracle_cv_fillpoly(uint8_t *img, int cols, int rows, const int32_t *xy, int n, uint8_t color) {
  cv_fillpoly(img, cols, rows, xy, n, color);
}
void oracle_cv_edt_sq(const uint8_t *img, int cols, int rows, int32_t *d2) { exact_edt_sq(img, cols, rows, d2); }

/* ------------------------------------------------------------------------- */
/* _get_cost + cost_data::cost_data — dpose_core.cpp:55-139                    */
/* ------------------------------------------------------------------------- */
static inline int reflect101(int i, int n) {
  if (n == 1) return 0;
  while (i < 0 || i >= n) i = i < 0 ? -i : 2 * (n - 1) - i;
  return i;
}

int oracle_cost_data_create(const int32_t *fp_xy, int n, unsigned padding,
                            oracle_cost_data *out) {
  memset(out, 0, sizeof(*out));
  if (n < 1) return -1;
  const int pad = (int)padding;
  /* :124-125 center = padding - rowwise min */
  int minx = INT_MAX, miny = INT_MAX, maxx = INT_MIN, maxy = INT_MIN;
  for (int i = 0; i < n; ++i) {
    const int x = fp_xy[2 * i], y = fp_xy[2 * i + 1];
    if (x < minx) minx = x;
    if (x > maxx) maxx = x;
    if (y < miny) miny = y;
    if (y > maxy) maxy = y;
  }
  out->center[0] = pad - minx;
  out->center[1] = pad - miny;
  /* :128 shifted footprint */
  int32_t *fp = (int32_t *)malloc(sizeof(int32_t) * 2 * (size_t)n);
  for (int i = 0; i < n; ++i) {
    fp[2 * i] = fp_xy[2 * i] + out->center[0];
    fp[2 * i + 1] = fp_xy[2 * i + 1] + out->center[1];
  }
  /* :61-69 boundingRect (w = dx+1, h = dy+1) + 2*padding */
  const int cols = (maxx - minx + 1) + 2 * pad;
  const int rows = (maxy - miny + 1) + 2 * pad;
  const size_t np = (size_t)rows * cols;
  out->rows = rows;
  out->cols = cols;
  out->outline = (uint8_t *)malloc(np);
  out->mask = (uint8_t *)malloc(np);
  out->edt_sq = (int32_t *)malloc(sizeof(int32_t) * np);
  out->pre_blur = (float *)malloc(sizeof(float) * np);
  out->cost = (float *)malloc(sizeof(float) * np);
  /* :72-73 white image, outline drawn with 0 */
  memset(out->outline, 255, np);
  cv_polylines_closed(out->outline, cols, rows, fp, n, 0);
  /* :81-83 exact EDT */
  exact_edt_sq(out->outline, cols, rows, out->edt_sq);
  /* :92-94 mask: 0 inside/on the polygon */
  memset(out->mask, 255, np);
  cv_fillpoly(out->mask, cols, rows, fp, n, 0);
  free(fp);
  /* :100-106 outer = edt where mask != 0, scaled by (float)-0.01; min over image */
  const float scale = (float)-0.01;
  float mn = 0.0f; /* inside pixels contribute 0 */
  int any = 0;
  for (size_t i = 0; i < np; ++i) {
    const float d = sqrtf((float)out->edt_sq[i]);
    float v;
    if (out->mask[i] != 0) {
      volatile float o = d * scale; /* volatile: forbid contraction with the subtract */
      v = o;
    } else {
      v = 0.0f;
    }
    if (!any || v < mn) {
      mn = v;
      any = 1;
    }
  }
  out->min_outer = mn;
  /* :109-110 edt = (outside ? outer : edt) - min */
  for (size_t i = 0; i < np; ++i) {
    const float d = sqrtf((float)out->edt_sq[i]);
    volatile float o = d * scale;
    const float v = out->mask[i] != 0 ? (float)o : d;
    volatile float r = v - mn;
    out->pre_blur[i] = r;
  }
  /* :111 GaussianBlur 3x3, sigma 0 -> [0.25 0.5 0.25] separable, row pass then column
   * pass, BORDER_REFLECT_101 */
  float *tmp = (float *)malloc(sizeof(float) * np);
  for (int y = 0; y < rows; ++y)
    for (int x = 0; x < cols; ++x) {
      const float a = out->pre_blur[(size_t)y * cols + reflect101(x - 1, cols)];
      const float b = out->pre_blur[(size_t)y * cols + x];
      const float c = out->pre_blur[(size_t)y * cols + reflect101(x + 1, cols)];
      volatile float s = a + c;
      volatile float r = s * 0.25f + b * 0.5f; /* both products exact */
      tmp[(size_t)y * cols + x] = r;
    }
  for (int y = 0; y < rows; ++y)
    for (int x = 0; x < cols; ++x) {
      const float a = tmp[(size_t)reflect101(y - 1, rows) * cols + x];
      const float b = tmp[(size_t)y * cols + x];
      const float c = tmp[(size_t)reflect101(y + 1, rows) * cols + x];
      volatile float s = a + c;
      volatile float r = s * 0.25f + b * 0.5f;
      out->cost[(size_t)y * cols + x] = r;
    }
  free(tmp);
  /* :138 box = ring(cols, rows) - center */
  const int bx[5] = {0, cols, cols, 0, 0};
  const int by[5] = {0, 0, rows, rows, 0};
  for (int i = 0; i < 5; ++i) {
    out->box[2 * i] = bx[i] - out->center[0];
    out->box[2 * i + 1] = by[i] - out->center[1];
  }
  return 0;
}

void oracle_cost_data_free(oracle_cost_data *cd) {
  free(cd->cost);
  free(cd->pre_blur);
  free(cd->edt_sq);
  free(cd->outline);
  free(cd->mask);
  memset(cd, 0, sizeof(*cd));
}

/* ------------------------------------------------------------------------- */
/* Eigen::Isometry2d restated — dpose_core.cpp:141-146, 206-219, 227          */
/* ------------------------------------------------------------------------- */
typedef struct {
  double r00, r01, r10, r11, t0, t1;
} iso2;

/* to_eigen: Translation2d(x, y) * Rotation2Dd(yaw) */
static inline iso2 to_eigen(double x, double y, double yaw) {
  iso2 m;
  const double c = cos(yaw), s = sin(yaw);
  m.r00 = c;
  m.r01 = -s;
  m.r10 = s;
  m.r11 = c;
  m.t0 = x;
  m.t1 = y;
  return m;
}

/* Transform * Transform: 3x3 coefficient-wise product, terms summed left to right */
static inline iso2 iso_mul(const iso2 *a, const iso2 *b) {
  iso2 m;
  m.r00 = a->r00 * b->r00 + a->r01 * b->r10;
  m.r01 = a->r00 * b->r01 + a->r01 * b->r11;
  m.r10 = a->r10 * b->r00 + a->r11 * b->r10;
  m.r11 = a->r10 * b->r01 + a->r11 * b->r11;
  m.t0 = (a->r00 * b->t0 + a->r01 * b->t1) + a->t0;
  m.t1 = (a->r10 * b->t0 + a->r11 * b->t1) + a->t1;
  return m;
}

/* Transform<.., Isometry>::inverse(): linear^T, -linear^T * translation */
static inline iso2 iso_inv(const iso2 *a) {
  iso2 m;
  m.r00 = a->r00;
  m.r01 = a->r10;
  m.r10 = a->r01;
  m.r11 = a->r11;
  m.t0 = -(m.r00 * a->t0 + m.r01 * a->t1);
  m.t1 = -(m.r10 * a->t0 + m.r11 * a->t1);
  return m;
}

static inline void iso_apply(const iso2 *a, double x, double y, double *ox, double *oy) {
  *ox = (a->r00 * x + a->r01 * y) + a->t0;
  *oy = (a->r10 * x + a->r11 * y) + a->t1;
}

/* interpolator — dpose_core.hpp:111-126, dpose_core.cpp:158-198 */
typedef struct {
  const float *img;
  int cols, rows;
  int upx, upy;
  double m00, m01, m10, m11; /* m(i,j) = f(x_i, y_j) */
} interp;

static inline int interp_init(interp *ip, double kx, double ky) {
  /* :177 round() half away from zero; :178 lower = upper - 1 */
  const int ux = (int)round(kx), uy = (int)round(ky);
  const int lx = ux - 1, ly = uy - 1;
  /* :181 bounds = (cols-1, rows-1) */
  if (lx < 1 || ly < 1 || ux >= ip->cols - 1 || uy >= ip->rows - 1) return 0;
  ip->upx = ux;
  ip->upy = uy;
  /* :162-165 */
  ip->m00 = ip->img[(size_t)ly * ip->cols + lx];
  ip->m01 = ip->img[(size_t)uy * ip->cols + lx];
  ip->m10 = ip->img[(size_t)ly * ip->cols + ux];
  ip->m11 = ip->img[(size_t)uy * ip->cols + ux];
  return 1;
}

static inline double interp_get(const interp *ip, double kx, double ky) {
  /* :193-197 c_rel = k - k_upper + 0.5; [1-cx, cx] * m * [1-cy, cy]^T */
  const double cx = kx - (double)ip->upx + 0.5;
  const double cy = ky - (double)ip->upy + 0.5;
  const double x0 = 1 - cx, x1 = cx, y0 = 1 - cy, y1 = cy;
  const double r0 = x0 * ip->m00 + x1 * ip->m10;
  const double r1 = x0 * ip->m01 + x1 * ip->m11;
  return r0 * y0 + r1 * y1;
}

double oracle_get_cost(const oracle_cost_data *cd, const double se2[3],
                       const int32_t *cells_xy, size_t n, double J[3]) {
  const double offset = 1e-6; /* :211 */
  const iso2 b_to_k = to_eigen(-(double)cd->center[0], -(double)cd->center[1], 0.0);
  iso2 tr[7];
  const double dx[7] = {0, -offset, offset, 0, 0, 0, 0};
  const double dy[7] = {0, 0, 0, -offset, offset, 0, 0};
  const double dz[7] = {0, 0, 0, 0, 0, -offset, offset};
  for (int i = 0; i < 7; ++i) {
    const iso2 m_to_b = to_eigen(se2[0] + dx[i], se2[1] + dy[i], se2[2] + dz[i]);
    const iso2 m_to_k = iso_mul(&m_to_b, &b_to_k);
    tr[i] = iso_inv(&m_to_k);
  }
  double sum = 0;
  if (J) J[0] = J[1] = J[2] = 0; /* :224-225 */
  interp ip;
  ip.img = cd->cost;
  ip.cols = cd->cols;
  ip.rows = cd->rows;
  for (size_t i = 0; i < n; ++i) {
    const double px = (double)cells_xy[2 * i], py = (double)cells_xy[2 * i + 1];
    double kx, ky;
    iso_apply(&tr[0], px, py, &kx, &ky); /* :234 */
    if (!interp_init(&ip, kx, ky)) continue; /* :237-238 */
    sum += interp_get(&ip, kx, ky); /* :241 */
    if (J) { /* :244-256 */
      double lo, hi;
      for (int a = 0; a < 3; ++a) {
        iso_apply(&tr[1 + 2 * a], px, py, &kx, &ky);
        lo = interp_get(&ip, kx, ky);
        iso_apply(&tr[2 + 2 * a], px, py, &kx, &ky);
        hi = interp_get(&ip, kx, ky);
        J[a] += (hi - lo) * 5e5;
      }
    }
  }
  return sum;
}

long double oracle_get_cost_truth(const oracle_cost_data *cd, const double se2[3],
                                  const int32_t *cells_xy, size_t n, long double J[3]) {
  const long double tx = se2[0], ty = se2[1];
  const long double c = cosl((long double)se2[2]), s = sinl((long double)se2[2]);
  const long double cx = cd->center[0], cy = cd->center[1];
  long double sum = 0, jx = 0, jy = 0, jt = 0;
  for (size_t i = 0; i < n; ++i) {
    const long double dx = (long double)cells_xy[2 * i] - tx;
    const long double dy = (long double)cells_xy[2 * i + 1] - ty;
    /* k = c + R(-theta) (p - t) */
    const long double ku = cx + c * dx + s * dy;
    const long double kv = cy - s * dx + c * dy;
    const int ux = (int)roundl(ku), uy = (int)roundl(kv);
    const int lx = ux - 1, ly = uy - 1;
    if (lx < 1 || ly < 1 || ux >= cd->cols - 1 || uy >= cd->rows - 1) continue;
    const long double m00 = cd->cost[(size_t)ly * cd->cols + lx];
    const long double m01 = cd->cost[(size_t)uy * cd->cols + lx];
    const long double m10 = cd->cost[(size_t)ly * cd->cols + ux];
    const long double m11 = cd->cost[(size_t)uy * cd->cols + ux];
    const long double rx = ku - ux + 0.5L, ry = kv - uy + 0.5L;
    sum += (1 - rx) * (1 - ry) * m00 + rx * (1 - ry) * m10 + (1 - rx) * ry * m01 +
           rx * ry * m11;
    const long double fu = (m10 - m00) * (1 - ry) + (m11 - m01) * ry;
    const long double fv = (m01 - m00) * (1 - rx) + (m11 - m10) * rx;
    jx += -c * fu + s * fv;
    jy += -s * fu - c * fv;
    jt += fu * (-s * dx + c * dy) + fv * (-c * dx - s * dy);
  }
  if (J) {
    J[0] = jx;
    J[1] = jy;
    J[2] = jt;
  }
  return sum;
}

/* ------------------------------------------------------------------------- */
/* raytrace / to_rays / lethal_cells_within — dpose_costmap.cpp:57-184         */
/* ------------------------------------------------------------------------- */
static inline int sgn(int v) { return (v > 0) - (v < 0); }

size_t oracle_raytrace(const int32_t begin[2], const int32_t end[2], int32_t *out_xy) {
  const int draw[2] = {end[0] - begin[0], end[1] - begin[1]};
  const int dabs[2] = {abs(draw[0]), abs(draw[1])};
  /* :64-71 Eigen maxCoeff/minCoeff return the FIRST extremum; equal -> minor = 1-major */
  int major = dabs[1] > dabs[0] ? 1 : 0;
  int minor = dabs[1] < dabs[0] ? 1 : 0;
  if (major == minor) minor = 1 - major;
  const int den = dabs[major], add = dabs[minor];
  int num = den / 2; /* :76 */
  int inc_minor[2] = {sgn(draw[0]), sgn(draw[1])};
  int inc_major[2] = {inc_minor[0], inc_minor[1]};
  inc_minor[major] = 0; /* :84-85 */
  inc_major[minor] = 0;
  int cur[2] = {begin[0], begin[1]};
  for (int i = 0; i < den; ++i) { /* :92-101 */
    out_xy[2 * i] = cur[0];
    out_xy[2 * i + 1] = cur[1];
    num += add;
    if (num >= den) {
      num -= den;
      cur[0] += inc_minor[0];
      cur[1] += inc_minor[1];
    }
    cur[0] += inc_major[0];
    cur[1] += inc_major[1];
  }
  return (size_t)den;
}

int oracle_to_rays(const int32_t box[10], int *y_lo, int *y_hi, int32_t **xmin_out,
                   int32_t **xmax_out) {
  int lo = INT_MAX, hi = INT_MIN, maxlen = 0;
  for (int i = 0; i < 5; ++i) {
    if (box[2 * i + 1] < lo) lo = box[2 * i + 1];
    if (box[2 * i + 1] > hi) hi = box[2 * i + 1];
  }
  for (int i = 0; i < 4; ++i) {
    const int a = abs(box[2 * i + 2] - box[2 * i]), b = abs(box[2 * i + 3] - box[2 * i + 1]);
    const int m = a > b ? a : b;
    if (m > maxlen) maxlen = m;
  }
  const int nrows = hi - lo + 1;
  int32_t *xmin = (int32_t *)malloc(sizeof(int32_t) * (size_t)nrows);
  int32_t *xmax = (int32_t *)malloc(sizeof(int32_t) * (size_t)nrows);
  for (int i = 0; i < nrows; ++i) {
    xmin[i] = INT_MAX;
    xmax[i] = INT_MIN;
  }
  int32_t *ray = (int32_t *)malloc(sizeof(int32_t) * 2 * (size_t)(maxlen + 1));
  for (int e = 0; e < 4; ++e) { /* :108-113 */
    const size_t len = oracle_raytrace(&box[2 * e], &box[2 * e + 2], ray);
    for (size_t i = 0; i < len; ++i) {
      const int x = ray[2 * i], y = ray[2 * i + 1];
      if (x < xmin[y - lo]) xmin[y - lo] = x;
      if (x > xmax[y - lo]) xmax[y - lo] = x;
    }
  }
  free(ray);
  *y_lo = lo;
  *y_hi = hi;
  *xmin_out = xmin;
  *xmax_out = xmax;
  return nrows;
}

static inline int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

size_t oracle_lethal_cells_within(const uint8_t *map, uint32_t size_x, uint32_t size_y,
                                  const int32_t box[10], int32_t **cells_xy) {
  *cells_xy = NULL;
  if (size_x == 0 || size_y == 0) return 0; /* :124-125 */
  int32_t sb[10];
  for (int i = 0; i < 5; ++i) { /* :121, :128-129 */
    sb[2 * i] = clampi(box[2 * i], 0, (int)size_x - 1);
    sb[2 * i + 1] = clampi(box[2 * i + 1], 0, (int)size_y - 1);
  }
  int y_lo, y_hi;
  int32_t *xmin, *xmax;
  const int nrows = oracle_to_rays(sb, &y_lo, &y_hi, &xmin, &xmax);
  size_t count = 0;
  for (int r = 0; r < nrows; ++r) { /* :143-155 count pass */
    if (xmin[r] > xmax[r]) continue;
    const uint8_t *row = map + (size_t)(y_lo + r) * size_x;
    for (int x = xmin[r]; x <= xmax[r]; ++x) count += (row[x] == ORACLE_LETHAL);
  }
  if (count) { /* :157-179 write pass */
    int32_t *cells = (int32_t *)malloc(sizeof(int32_t) * 2 * count);
    size_t dd = 0;
    for (int r = 0; r < nrows; ++r) {
      if (xmin[r] > xmax[r]) continue;
      const uint8_t *row = map + (size_t)(y_lo + r) * size_x;
      for (int x = xmin[r]; x <= xmax[r]; ++x)
        if (row[x] == ORACLE_LETHAL) {
          cells[2 * dd] = x;
          cells[2 * dd + 1] = y_lo + r;
          ++dd;
        }
    }
    *cells_xy = cells;
  }
  free(xmin);
  free(xmax);
  return count;
}

/* ------------------------------------------------------------------------- */
/* batch driver: per pose lethal_cells_within(rotated box) +  */
/* get_cost.  Box rotation as in test/lethal_cells_within.cpp:82-86.           */
/* ------------------------------------------------------------------------- */
static void rotated_box(const oracle_cost_data *cd, const double *pose, int32_t out[10]) {
  const iso2 rot = to_eigen(pose[0], pose[1], pose[2]);
  for (int i = 0; i < 5; ++i) {
    double x, y;
    iso_apply(&rot, (double)cd->box[2 * i], (double)cd->box[2 * i + 1], &x, &y);
    out[2 * i] = (int32_t)round(x);
    out[2 * i + 1] = (int32_t)round(y);
  }
}

void oracle_get_cost_batch(const oracle_cost_data *cd, const uint8_t *map, uint32_t size_x,
                           uint32_t size_y, const double *poses, size_t n, double *out,
                           int want_jacobian, int shift, int nthreads) {
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 256) num_threads(nthreads > 1 ? nthreads : 1)
#endif
  for (long long i = 0; i < (long long)n; ++i) {
    const double *p = poses + 3 * i;
    int32_t box[10];
    rotated_box(cd, p, box);
    int32_t *cells = NULL;
    const size_t nc = oracle_lethal_cells_within(map, size_x, size_y, box, &cells);
    double se2[3] = {p[0], p[1], p[2]};
    if (shift) {
      const int sx = (int)floor(p[0]), sy = (int)floor(p[1]);
      se2[0] -= sx;
      se2[1] -= sy;
      for (size_t k = 0; k < nc; ++k) {
        cells[2 * k] -= sx;
        cells[2 * k + 1] -= sy;
      }
    }
    double J[3] = {0, 0, 0};
    out[4 * i] = oracle_get_cost(cd, se2, cells, nc, want_jacobian ? J : NULL);
    out[4 * i + 1] = J[0];
    out[4 * i + 2] = J[1];
    out[4 * i + 3] = J[2];
    free(cells);
  }
  (void)nthreads;
}

void oracle_get_cost_batch_truth(const oracle_cost_data *cd, const uint8_t *map,
                                 uint32_t size_x, uint32_t size_y, const double *poses,
                                 size_t n, double *out, int nthreads) {
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 256) num_threads(nthreads > 1 ? nthreads : 1)
#endif
  for (long long i = 0; i < (long long)n; ++i) {
    const double *p = poses + 3 * i;
    int32_t box[10];
    rotated_box(cd, p, box);
    int32_t *cells = NULL;
    const size_t nc = oracle_lethal_cells_within(map, size_x, size_y, box, &cells);
    long double J[3] = {0, 0, 0};
    out[4 * i] = (double)oracle_get_cost_truth(cd, p, cells, nc, J);
    out[4 * i + 1] = (double)J[0];
    out[4 * i + 2] = (double)J[1];
    out[4 * i + 3] = (double)J[2];
    free(cells);
  }
  (void)nthreads;
}


/* ------------------------------------------------------------------------- */
/* dpose_goal_tolerance::problem — dpose_goal_tolerance.cpp:54-292            */
/* ------------------------------------------------------------------------- */
#define ORACLE_PI 3.14159265358979323846 /* M_PI (not in strict C11) */
double oracle_normalize_angle(double a) {
  /* angles::normalize_angle (ROS `angles`, noetic): */
  const double r = fmod(a + ORACLE_PI, 2.0 * ORACLE_PI);
  return r <= 0.0 ? r + ORACLE_PI : r - ORACLE_PI;
}

static double shortest_angular_distance(double from, double to) { return oracle_normalize_angle(to - from); }

/* pose_regularization::init (:62-74) */
static void reg_init(oracle_reg *r, double w_lin_unscaled, double w_rot_unscaled, const double start[3],
                     const double goal[3]) {
  double w_lin = (start[0] - goal[0]) * (start[0] - goal[0]) + (start[1] - goal[1]) * (start[1] - goal[1]);
  double w_rot = pow(shortest_angular_distance(start[2], goal[2]), 2);
  w_lin = w_lin_unscaled / fmax(w_lin, 1e-3); /* :70-71 avoid division by zero */
  w_rot = w_rot_unscaled / fmax(w_rot, 1e-3);
  for (int i = 0; i < 3; ++i) r->goal[i] = goal[i];
  r->norm[0] = w_lin, r->norm[1] = w_lin, r->norm[2] = w_rot;
}

/* pose_regularization::get_cost (:76-88): cost, and J = 2 * norm .* diff */
static double reg_cost(const oracle_reg *r, const double pose[3], double J[3]) {
  double diff[3] = {pose[0] - r->goal[0], pose[1] - r->goal[1], pose[2] - r->goal[2]};
  diff[2] = oracle_normalize_angle(diff[2]);
  double cost = 0.0;
  for (int i = 0; i < 3; ++i) {
    J[i] = 2 * (r->norm[i] * diff[i]);
    cost += r->norm[i] * (diff[i] * diff[i]);
  }
  return cost;
}

int oracle_problem_init(const oracle_cost_data *cd, const uint8_t *map, uint32_t size_x, uint32_t size_y,
                        const double start[3], const double goal[3], const oracle_gt_config *cfg,
                        oracle_problem *out) {
  memset(out, 0, sizeof(*out));
  for (int i = 0; i < 3; ++i) out->pose[i] = goal[i]; /* :114 */
  out->lin_tol_sq = pow(cfg->lin_tolerance, 2);       /* :115-116 */
  out->rot_tol_sq = pow(cfg->rot_tolerance, 2);
  int max_coeff = 0; /* :122-123 */
  for (int i = 0; i < 10; ++i) max_coeff = abs(cd->box[i]) > max_coeff ? abs(cd->box[i]) : max_coeff;
  /* :126-133: h is a double, the ring is a rectangle<int>: Eigen's comma initialiser truncates each coefficient
   * toward zero before the goal is added and the sum rounded (:134-138) */
  const double h = (double)(int)(max_coeff + fabs(cfg->lin_tolerance));
  const double bx[5] = {-h, h, h, -h, -h}, by[5] = {-h, -h, h, h, -h}; /* :129-133 */
  for (int i = 0; i < 5; ++i) {                                        /* :134-138, Eigen round = std::round */
    out->search_box[2 * i] = (int32_t)round(bx[i] + goal[0]);
    out->search_box[2 * i + 1] = (int32_t)round(by[i] + goal[1]);
  }
  out->n_cells = oracle_lethal_cells_within(map, size_x, size_y, out->search_box, &out->cells); /* :141 */
  out->n_regs = 0;
  if (cfg->weight_goal_lin != 0 || cfg->weight_goal_rot != 0) { /* :145-150 */
    const double s[3] = {goal[0] + cfg->lin_tolerance, goal[1] + 0, goal[2] + cfg->rot_tolerance};
    reg_init(&out->regs[out->n_regs++], cfg->weight_goal_lin, cfg->weight_goal_rot, s, goal);
  }
  if (cfg->weight_start_lin != 0 || cfg->weight_start_rot != 0) { /* :152-180 */
    double s[3] = {start[0], start[1], start[2]};
    double diff[3] = {start[0] - goal[0], start[1] - goal[1], start[2] - goal[2]};
    diff[2] = oracle_normalize_angle(diff[2]);
    const double lin_norm = sqrt(diff[0] * diff[0] + diff[1] * diff[1]);
    const double rot_norm = fabs(diff[2]);
    if (lin_norm != 0 && lin_norm > cfg->lin_tolerance) /* :170-171 (theta corrected below) */
      for (int i = 0; i < 3; ++i) s[i] = goal[i] + diff[i] * cfg->lin_tolerance / lin_norm;
    if (rot_norm != 0 && rot_norm > cfg->rot_tolerance) /* :174-177 */
      s[2] = goal[2] + diff[2] * cfg->rot_tolerance / rot_norm;
    else
      s[2] = start[2];
    reg_init(&out->regs[out->n_regs++], cfg->weight_start_lin, cfg->weight_start_rot, goal, s); /* :179 */
  }
  return 0;
}

void oracle_problem_free(oracle_problem *p) {
  free(p->cells);
  p->cells = NULL;
  p->n_cells = 0;
}

void oracle_problem_eval(const oracle_cost_data *cd, const oracle_problem *p, const double x[3], double *f,
                         double grad_f[3], double g[2], double jac_g[3]) {
  const double pose[3] = {x[0] + p->pose[0], x[1] + p->pose[1], x[2] + p->pose[2]}; /* :98-100 */
  double J[3];
  double cost = oracle_get_cost(cd, pose, p->cells, p->n_cells, J); /* :103 */
  for (int r = 0; r < p->n_regs; ++r) {                             /* :106-109 */
    double Jr[3];
    cost += reg_cost(&p->regs[r], pose, Jr);
    for (int i = 0; i < 3; ++i) J[i] += Jr[i];
  }
  if (f) *f = cost;
  if (grad_f)
    for (int i = 0; i < 3; ++i) grad_f[i] = J[i];
  if (g) { /* :238-247 */
    g[0] = x[0] * x[0] + x[1] * x[1];
    g[1] = x[2] * x[2];
  }
  if (jac_g) /* :270-279 */
    for (int i = 0; i < 3; ++i) jac_g[i] = x[i] * 2;
}

void oracle_problem_eval_batch(const oracle_cost_data *cd, const oracle_problem *p, const double *x, size_t n,
                               double *f, double *grad_f, double *g, double *jac_g, int nthreads) {
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 64) num_threads(nthreads > 1 ? nthreads : 1)
#endif
  for (long long i = 0; i < (long long)n; ++i)
    oracle_problem_eval(cd, p, x + 3 * i, f ? f + i : NULL, grad_f ? grad_f + 3 * i : NULL, g ? g + 2 * i : NULL,
                        jac_g ? jac_g + 3 * i : NULL);
  (void)nthreads;
}


/* ------------------------------------------------------------------------- */
/* the multi-start solve: step rule restated (see dpose_oracle.h)            */
/* ------------------------------------------------------------------------- */
/* Euclidean projection onto {x^2 + y^2 <= lin_tol^2} x {|theta| <= rot_tol}: the feasible set of get_bounds_info
 * (dpose_goal_tolerance.cpp:196-207) with g0, g1 of eval_g (:238-247) */
static void solver_project(double p[3], const oracle_solver_cfg *c) {
  const double r2 = p[0] * p[0] + p[1] * p[1];
  if (r2 > c->lin_tol * c->lin_tol) {
    const double sc = c->lin_tol / sqrt(r2);
    p[0] = p[0] * sc;
    p[1] = p[1] * sc;
  }
  p[2] = fmin(fmax(p[2], -c->rot_tol), c->rot_tol);
}

/* H = gamma * diag(1, 1, 1 / rho^2) */
static void solver_h0(oracle_solver_state *s, double gamma, const oracle_solver_cfg *c) {
  s->H[0] = gamma, s->H[1] = 0.0, s->H[2] = 0.0, s->H[3] = gamma, s->H[4] = 0.0, s->H[5] = gamma / c->rho2;
}

static void solver_direction(oracle_solver_state *s) { /* d = -(H g) */
  const double *H = s->H, *g = s->g;
  s->d[0] = -(H[0] * g[0] + H[1] * g[1] + H[2] * g[2]);
  s->d[1] = -(H[1] * g[0] + H[3] * g[1] + H[4] * g[2]);
  s->d[2] = -(H[2] * g[0] + H[4] * g[1] + H[5] * g[2]);
}

/* first step of a (re)started descent: at most one cell in the metric dx^2 + dy^2 + rho^2 dtheta^2 */
static double solver_first_alpha(const oracle_solver_state *s, const oracle_solver_cfg *c) {
  const double nd = sqrt(s->d[0] * s->d[0] + s->d[1] * s->d[1] + c->rho2 * (s->d[2] * s->d[2]));
  return nd > 1.0 ? 1.0 / nd : 1.0;
}

/* trial point xt = P(x + alpha d); usable iff it is a descent step of non-negligible length */
static int solver_trial(oracle_solver_state *s, const oracle_solver_cfg *c) {
  for (int k = 0; k < 3; ++k) s->xt[k] = s->x[k] + s->alpha * s->d[k];
  solver_project(s->xt, c);
  const double e0 = s->xt[0] - s->x[0], e1 = s->xt[1] - s->x[1], e2 = s->xt[2] - s->x[2];
  s->gd = s->g[0] * e0 + s->g[1] * e1 + s->g[2] * e2;
  const double step = fmax(fmax(fabs(e0), fabs(e1)), fabs(e2));
  return s->gd < 0.0 && step > 1e-12;
}

void oracle_solver_start(oracle_solver_state *s, const oracle_solver_cfg *c, const double x0[3]) {
  memset(s, 0, sizeof(*s));
  for (int k = 0; k < 3; ++k) s->xt[k] = x0[k];
  solver_project(s->xt, c);
  for (int k = 0; k < 3; ++k) s->x[k] = s->xt[k];
  s->alpha = 1.0;
  solver_h0(s, 1.0, c);
}

void oracle_solver_advance(oracle_solver_state *s, const oracle_solver_cfg *c, double ft, const double gt[3]) {
  s->evals += 1;
  int fresh = 1;
  if (s->evals == 1) { /* the starting point is taken as it is */
    for (int k = 0; k < 3; ++k) s->x[k] = s->xt[k], s->g[k] = gt[k];
    s->f = ft;
    solver_h0(s, 1.0, c);
    s->n_acc = 0;
  } else if (ft <= s->f + c->c1 * s->gd) { /* Armijo on the projected arc: accept, BFGS update */
    const double sx[3] = {s->xt[0] - s->x[0], s->xt[1] - s->x[1], s->xt[2] - s->x[2]};
    const double y[3] = {gt[0] - s->g[0], gt[1] - s->g[1], gt[2] - s->g[2]};
    const double sy = sx[0] * y[0] + sx[1] * y[1] + sx[2] * y[2];
    const double ss = sx[0] * sx[0] + sx[1] * sx[1] + sx[2] * sx[2];
    const double yy = y[0] * y[0] + y[1] * y[1] + y[2] * y[2];
    if (sy > 1e-10 * (sqrt(ss) * sqrt(yy))) {
      if (s->n_acc == 0) solver_h0(s, sy / (y[0] * y[0] + y[1] * y[1] + (y[2] * y[2]) / c->rho2), c);
      double *H = s->H;
      const double rho = 1.0 / sy;
      const double Hy[3] = {H[0] * y[0] + H[1] * y[1] + H[2] * y[2], H[1] * y[0] + H[3] * y[1] + H[4] * y[2],
                            H[2] * y[0] + H[4] * y[1] + H[5] * y[2]};
      const double yHy = y[0] * Hy[0] + y[1] * Hy[1] + y[2] * Hy[2];
      const double cf = (1.0 + rho * yHy) * rho;
      /* H <- H + cf s s' - rho (Hy s' + s Hy') */
      static const int ri[6] = {0, 0, 0, 1, 1, 2}, ci[6] = {0, 1, 2, 1, 2, 2};
      for (int e = 0; e < 6; ++e) {
        const int i = ri[e], j = ci[e];
        H[e] = H[e] + cf * (sx[i] * sx[j]) - rho * (Hy[i] * sx[j] + sx[i] * Hy[j]);
      }
      s->n_acc += 1;
    }
    for (int k = 0; k < 3; ++k) s->x[k] = s->xt[k], s->g[k] = gt[k];
    s->f = ft;
  } else { /* shorter step: minimiser of the parabola through f, gd and ft, kept inside [0.1, 0.5] */
    double t = -s->gd / (2.0 * ((ft - s->f) - s->gd));
    if (!(t >= 0.1)) t = 0.1;
    if (t > 0.5) t = 0.5;
    s->alpha = s->alpha * t;
    fresh = 0;
  }
  if (fresh) {
    double p[3] = {s->x[0] - s->g[0], s->x[1] - s->g[1], s->x[2] - s->g[2]};
    solver_project(p, c);
    const double kkt = fmax(fmax(fabs(s->x[0] - p[0]), fabs(s->x[1] - p[1])), fabs(s->x[2] - p[2]));
    if (kkt <= c->kkt_tol) {
      s->status = 1;
      return;
    }
    solver_direction(s);
    s->alpha = s->n_acc == 0 ? solver_first_alpha(s, c) : 1.0;
  }
  if (s->evals > c->max_iter) {
    s->status = 2;
    return;
  }
  if (solver_trial(s, c)) return;
  if (fresh && s->n_acc > 0) { /* quasi-Newton direction bent by the projection: restart from the scaled gradient */
    solver_h0(s, 1.0, c);
    s->n_acc = 0;
    solver_direction(s);
    s->alpha = solver_first_alpha(s, c);
    if (solver_trial(s, c)) return;
  }
  s->status = 3;
}

int oracle_solve(const oracle_solver_cfg *cfg, oracle_eval_fn eval, void *ctx, const double x0[3], double x[3], double *f,
                 int *evals, double *trace) {
  oracle_solver_state s;
  oracle_solver_start(&s, cfg, x0);
  while (s.status == 0) {
    double ft, gt[3], g[2], jac[3];
    eval(ctx, s.xt, &ft, gt, g, jac);
    const int slot = s.evals;
    const double alpha_used = s.alpha, xt[3] = {s.xt[0], s.xt[1], s.xt[2]};
    oracle_solver_advance(&s, cfg, ft, gt);
    if (trace) {
      double *t = trace + 12 * (size_t)slot;
      t[0] = xt[0], t[1] = xt[1], t[2] = xt[2], t[3] = ft, t[4] = gt[0], t[5] = gt[1], t[6] = gt[2];
      t[7] = g[0], t[8] = g[1], t[9] = alpha_used, t[10] = (double)s.status, t[11] = 0.0;
    }
  }
  for (int k = 0; k < 3; ++k) x[k] = s.x[k];
  *f = s.f;
  if (evals) *evals = s.evals;
  return s.status;
}

void oracle_problem_eval_cb(void *ctx, const double x[3], double *f, double grad_f[3], double g[2], double jac_g[3]) {
  const oracle_problem_ctx *c = (const oracle_problem_ctx *)ctx;
  oracle_problem_eval(c->cd, c->p, x, f, grad_f, g, jac_g);
}

void oracle_solve_batch(const oracle_cost_data *cd, const oracle_problem *p, const oracle_solver_cfg *cfg,
                        const double *x0, size_t n, double *x, double *f, uint8_t *status, int32_t *evals, double *trace,
                        int nthreads) {
  oracle_problem_ctx ctx = {cd, p};
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 16) num_threads(nthreads > 1 ? nthreads : 1)
#endif
  for (long long i = 0; i < (long long)n; ++i) {
    int ev = 0;
    const int st = oracle_solve(cfg, oracle_problem_eval_cb, &ctx, x0 + 3 * i, x + 3 * i, f + i, &ev,
                                trace ? trace + (size_t)i * 12 * (size_t)(cfg->max_iter + 1) : NULL);
    if (status) status[i] = (uint8_t)st;
    if (evals) evals[i] = ev;
  }
  (void)nthreads;
}

/* std::mt19937 (32-bit Mersenne twister, the standard's parameters) */
typedef struct {
  uint32_t mt[624];
  int idx;
} oracle_mt19937;
static void mt_seed(oracle_mt19937 *m, uint32_t seed) {
  m->mt[0] = seed;
  for (int i = 1; i < 624; ++i) m->mt[i] = 1812433253u * (m->mt[i - 1] ^ (m->mt[i - 1] >> 30)) + (uint32_t)i;
  m->idx = 624;
}
static uint32_t mt_next(oracle_mt19937 *m) {
  if (m->idx >= 624) {
    for (int i = 0; i < 624; ++i) {
      const uint32_t y = (m->mt[i] & 0x80000000u) | (m->mt[(i + 1) % 624] & 0x7fffffffu);
      m->mt[i] = m->mt[(i + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
    }
    m->idx = 0;
  }
  uint32_t y = m->mt[m->idx++];
  y ^= y >> 11;
  y ^= (y << 7) & 0x9d2c5680u;
  y ^= (y << 15) & 0xefc60000u;
  y ^= y >> 18;
  return y;
}
/* libstdc++ std::uniform_real_distribution<double>(a, b)(mt19937): generate_canonical<double, 53> takes two draws,
 * sum = r1 + r2 * 2^32 (in double), / 2^64, clamped below 1; then u * (b - a) + a */
static double mt_uniform(oracle_mt19937 *m, double a, double b) {
  const double r1 = (double)mt_next(m), r2 = (double)mt_next(m);
  double u = (r1 + r2 * 4294967296.0) / 18446744073709551616.0;
  if (u >= 1.0) u = nextafter(1.0, 0.0);
  return u * (b - a) + a;
}

void oracle_sample_offsets(uint32_t seed, double lin_tol, double rot_tol, size_t n, int first_zero, double *x0) {
  oracle_mt19937 m;
  mt_seed(&m, seed);
  for (size_t i = 0; i < n; ++i) {
    if (i == 0 && first_zero) { /* attempt 0 starts at the goal (dpose_goal_tolerance.cpp:431) */
      x0[0] = x0[1] = x0[2] = 0.0;
      continue;
    }
    const double radius = mt_uniform(&m, 0.0, lin_tol); /* :346-348: radius, angle, yaw in this order */
    const double pi = 3.14159265358979323846; /* M_PI */
    const double angle = mt_uniform(&m, -pi, pi);
    const double yaw = mt_uniform(&m, 0.0, rot_tol);
    x0[3 * i] = cos(angle) * radius; /* :351-352 */
    x0[3 * i + 1] = sin(angle) * radius;
    x0[3 * i + 2] = yaw;
  }
}

/* ------------------------------------------------------------------------- */
/* check_footprint — dpose_costmap.cpp:186-582                                */
/* ------------------------------------------------------------------------- */
int oracle_is_inside(uint32_t size_x, uint32_t size_y, const int32_t *fp, int n) { /* :186-199 */
  for (int i = 0; i < n; ++i)
    if (fp[2 * i] < 0 || fp[2 * i + 1] < 0 || fp[2 * i] >= (int)size_x || fp[2 * i + 1] >= (int)size_y) return 0;
  return 1;
}

/* base_iterator (:203-219) */
static void xb_base(oracle_xb *b, const int32_t begin[2], const int32_t end[2]) {
  const int ddx = end[0] - begin[0], ddy = end[1] - begin[1];
  const int ax = abs(ddx), ay = abs(ddy);
  b->x_sign = (ddx > 0) - (ddx < 0);
  b->den = ax > ay ? ax : ay;
  b->add = ax < ay ? ax : ay;
  b->num = b->den / 2;
  b->dx = (double)ddx / ddy;
}

void oracle_xb_init(oracle_xb *b, const int32_t c0[2], const int32_t c1[2]) { /* :303-317 */
  const int ax = abs(c1[0] - c0[0]), ay = abs(c1[1] - c0[1]);
  if (ax > ay) {
    if (c1[1] > c0[1]) { /* x_major_ascending(c0, c1) :259-262 */
      b->kind = 2;
      xb_base(b, c0, c1);
      b->x_curr = c0[0];
    } else { /* x_major_descending(begin = c1, end = c0): base_iterator(end, begin) :281-284 */
      b->kind = 3;
      xb_base(b, c0, c1);
      b->x_curr = c1[0];
    }
  } else {
    if (c1[1] > c0[1]) { /* x_minor_ascending(c0, c1) :221-224 */
      b->kind = 0;
      xb_base(b, c0, c1);
      b->x_curr = c0[0];
    } else { /* x_minor_descending(begin = c1, end = c0): base_iterator(end, begin) :239-242 */
      b->kind = 1;
      xb_base(b, c0, c1);
      b->x_curr = c1[0];
    }
  }
}

void oracle_xb_advance(oracle_xb *b) {
  switch (b->kind) {
    case 0: /* :226-237 */
      b->num += b->add;
      if (b->num >= b->den) {
        b->num -= b->den;
        b->x_curr += b->x_sign;
      }
      break;
    case 1: /* :244-255 */
      b->num -= b->add;
      if (b->num < 0) {
        b->num += b->den;
        b->x_curr -= b->x_sign;
      }
      break;
    case 2: { /* :264-279 */
      const int diff = b->den - b->num, step = diff / b->add;
      b->num += step * b->add;
      b->x_curr += step * b->x_sign;
      if (b->num < b->den) {
        b->num += b->add;
        b->x_curr += b->x_sign;
      }
      b->num -= b->den;
      break;
    }
    default: { /* :286-301 */
      const int step = b->num / b->add;
      b->num -= step * b->add;
      b->x_curr -= step * b->x_sign;
      if (b->num >= 0) {
        b->num -= b->add;
        b->x_curr -= b->x_sign;
      }
      b->num += b->den;
      break;
    }
  }
}

typedef struct {
  oracle_xb xb;
  int lower_y; /* line::lower.y (:320-325): y of the edge's FIRST vertex as given */
  int active;
} cf_line;

static int ray_is_good(const uint8_t *map, uint32_t size_x, const int32_t a[2], const int32_t b[2], uint8_t cost) {
  const int dx = abs(b[0] - a[0]), dy = abs(b[1] - a[1]);
  const size_t cap = (size_t)(dx > dy ? dx : dy);
  if (cap == 0) return 1;
  int32_t *cells = (int32_t *)malloc(sizeof(int32_t) * 2 * cap);
  const size_t k = oracle_raytrace(a, b, cells);
  int good = 1;
  for (size_t i = 0; i < k && good; ++i)
    if (map[(size_t)cells[2 * i + 1] * size_x + cells[2 * i]] == cost) good = 0;
  free(cells);
  return good;
}

int oracle_check_footprint(const uint8_t *map, uint32_t size_x, uint32_t size_y, const int32_t *fp, int n,
                           uint8_t cost) {
  if (!oracle_is_inside(size_x, size_y, fp, n)) return 0; /* :464-465 */
  if (n < 3) {                                            /* check_without_area :392-422 */
    if (n == 0) return 1;
    if (n == 1) return map[(size_t)fp[1] * size_x + fp[0]] != cost;
    return ray_is_good(map, size_x, fp, fp + 2, cost);
  }
  /* check_outline :432-458 */
  for (int i = 1; i < n; ++i)
    if (!ray_is_good(map, size_x, fp + 2 * (i - 1), fp + 2 * i, cost)) return 0;
  const int is_closed = fp[2 * (n - 1)] == fp[0] && fp[2 * (n - 1) + 1] == fp[1];
  if (!is_closed && !ray_is_good(map, size_x, fp + 2 * (n - 1), fp, cost)) return 0;

  /* edges, horizontal ones skipped (:487-496) */
  cf_line *lines = (cf_line *)malloc(sizeof(cf_line) * (size_t)(n + 1));
  int nl = 0;
  for (int c = 1; c < n; ++c)
    if (fp[2 * (c - 1) + 1] != fp[2 * c + 1]) {
      oracle_xb_init(&lines[nl].xb, fp + 2 * (c - 1), fp + 2 * c);
      lines[nl].lower_y = fp[2 * (c - 1) + 1];
      lines[nl].active = 0;
      ++nl;
    }
  if (!is_closed && fp[2 * (n - 1) + 1] != fp[1]) {
    oracle_xb_init(&lines[nl].xb, fp + 2 * (n - 1), fp);
    lines[nl].lower_y = fp[2 * (n - 1) + 1];
    lines[nl].active = 0;
    ++nl;
  }
  if (nl == 0) { /* degenerate: every edge horizontal; upstream would dereference lines.front() — treat as outline only */
    free(lines);
    return 1;
  }
  /* vertex_set (:503-510): multimap keyed by y, equal keys in insertion order */
  int *vy = (int *)malloc(sizeof(int) * (size_t)nl), *vl = (int *)malloc(sizeof(int) * (size_t)nl),
      *vr = (int *)malloc(sizeof(int) * (size_t)nl), *order = (int *)malloc(sizeof(int) * (size_t)nl);
  int nv = 0;
  for (int l = 0; l + 1 < nl; ++l) {
    vy[nv] = lines[l + 1].lower_y;
    vl[nv] = l;
    vr[nv] = l + 1;
    ++nv;
  }
  vy[nv] = lines[0].lower_y;
  vl[nv] = nl - 1;
  vr[nv] = 0;
  ++nv;
  for (int i = 0; i < nv; ++i) order[i] = i;
  for (int i = 1; i < nv; ++i) { /* stable insertion sort by y */
    const int o = order[i];
    int j = i - 1;
    while (j >= 0 && vy[order[j]] > vy[o]) {
      order[j + 1] = order[j];
      --j;
    }
    order[j + 1] = o;
  }
  int *active = (int *)malloc(sizeof(int) * (size_t)nl), na = 0;
  int ymin = fp[1];
  for (int i = 1; i < n; ++i) ymin = fp[2 * i + 1] < ymin ? fp[2 * i + 1] : ymin;
  int yy = ymin - 1; /* :519 */
  int result = 1;
  for (int v = 0; v < nv && result;) {
    for (; yy < vy[order[v]] && result; ++yy) { /* :549-556 */
      /* check_line (:349-378): pairwise sweep between the active lines */
      for (int a = 0; a + 1 < na && result; a += 2) {
        const int lx = lines[active[a]].xb.x_curr + 1, rx = lines[active[a + 1]].xb.x_curr;
        for (int x = lx; x < rx; ++x)
          if (map[(size_t)yy * size_x + x] == cost) {
            result = 0;
            break;
          }
      }
      for (int a = 0; a < na; ++a) oracle_xb_advance(&lines[active[a]].xb);
    }
    if (!result) break;
    for (; v < nv && vy[order[v]] == yy; ++v) { /* :562-581 */
      const int pair[2] = {vl[order[v]], vr[order[v]]};
      for (int k = 0; k < 2; ++k) {
        const int l = pair[k];
        if (lines[l].active) {
          int pos = 0;
          while (pos < na && active[pos] != l) ++pos;
          if (pos < na) { /* swap and pop */
            active[pos] = active[na - 1];
            --na;
          }
        } else {
          lines[l].active = 1;
          active[na++] = l;
        }
      }
    }
    /* std::sort(active, lines_compare) :583 — by x, then dx; insertion sort (stable) for determinism */
    for (int i = 1; i < na; ++i) {
      const int o = active[i];
      int j = i - 1;
      while (j >= 0 && (lines[active[j]].xb.x_curr > lines[o].xb.x_curr ||
                        (lines[active[j]].xb.x_curr == lines[o].xb.x_curr && lines[active[j]].xb.dx > lines[o].xb.dx))) {
        active[j + 1] = active[j];
        --j;
      }
      active[j + 1] = o;
    }
  }
  free(lines);
  free(vy);
  free(vl);
  free(vr);
  free(order);
  free(active);
  return result;
}

void oracle_check_footprint_batch(const uint8_t *map, uint32_t size_x, uint32_t size_y, const int32_t *fp, int n,
                                  const double *poses, size_t n_poses, uint8_t cost, uint8_t *ok, int nthreads) {
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 64) num_threads(nthreads > 1 ? nthreads : 1)
#endif
  for (long long i = 0; i < (long long)n_poses; ++i) {
    int32_t *t = (int32_t *)malloc(sizeof(int32_t) * 2 * (size_t)(n > 0 ? n : 1));
    const double c = cos(poses[3 * i + 2]), s = sin(poses[3 * i + 2]);
    for (int k = 0; k < n; ++k) { /* dpose_goal_tolerance.cpp:388-392 */
      const double x = fp[2 * k], y = fp[2 * k + 1];
      t[2 * k] = (int32_t)round((c * x + -s * y) + poses[3 * i]);
      t[2 * k + 1] = (int32_t)round((s * x + c * y) + poses[3 * i + 1]);
    }
    ok[i] = (uint8_t)oracle_check_footprint(map, size_x, size_y, t, n, cost);
    free(t);
  }
  (void)nthreads;
}
