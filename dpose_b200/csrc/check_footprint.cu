// check_footprint.cu — batched footprint collision check.
//
// Upstream check_footprint (dpose_costmap.cpp:460-582) validates ONE solved pose per Ipopt solve
// (dpose_goal_tolerance.cpp:386-399): the footprint is transformed by the pose, rounded to cells, and the
// outline (Bresenham rays) plus the scan-line interior between pairs of active edges (custom x-major /
// x-minor Bresenham iterators, :201-317) is searched for a cell equal to `cost`.  With multi-start in
// lock-step there are thousands of candidates per iteration, so the check runs here with one thread per
// candidate pose.  Each thread executes the reference's algorithm step for step (same iterators, same
// half-open vertex rule, same pairwise sweep l_x + 1 .. r_x - 1), so the boolean is identical by
// construction; the work per pose is a few hundred byte reads, all L2 hits on the map.
#include <cmath>

#include "common.cuh"

namespace dpose {
namespace {

constexpr int kMaxVertices = 64;

struct XB {  // x_bresenham state (base_iterator, :203-219)
  int x_curr, x_sign, den, add, num;
  double dx;
  int kind;  // 0 x_minor_ascending, 1 x_minor_descending, 2 x_major_ascending, 3 x_major_descending
};

__device__ inline void xb_init(XB &b, int x0, int y0, int x1, int y1) {  // :303-317
  const int ddx = x1 - x0, ddy = y1 - y0;
  const int ax = abs(ddx), ay = abs(ddy);
  b.x_sign = (ddx > 0) - (ddx < 0);
  b.den = max(ax, ay);
  b.add = min(ax, ay);
  b.num = b.den / 2;
  b.dx = (double)ddx / (double)ddy;
  const bool up = y1 > y0;
  b.kind = ax > ay ? (up ? 2 : 3) : (up ? 0 : 1);
  b.x_curr = up ? x0 : x1;  // iteration starts at the vertex with the lower y
}

__device__ inline void xb_advance(XB &b) {
  switch (b.kind) {
    case 0:  // :226-237
      b.num += b.add;
      if (b.num >= b.den) {
        b.num -= b.den;
        b.x_curr += b.x_sign;
      }
      break;
    case 1:  // :244-255
      b.num -= b.add;
      if (b.num < 0) {
        b.num += b.den;
        b.x_curr -= b.x_sign;
      }
      break;
    case 2: {  // :264-279
      const int step = (b.den - b.num) / b.add;
      b.num += step * b.add;
      b.x_curr += step * b.x_sign;
      if (b.num < b.den) {
        b.num += b.add;
        b.x_curr += b.x_sign;
      }
      b.num -= b.den;
      break;
    }
    default: {  // :286-301
      const int step = b.num / b.add;
      b.num -= step * b.add;
      b.x_curr -= step * b.x_sign;
      if (b.num >= 0) {
        b.num -= b.add;
        b.x_curr -= b.x_sign;
      }
      b.num += b.den;
      break;
    }
  }
}

// raytrace (dpose_costmap.cpp:57-104), end exclusive; true if no visited cell equals cost
__device__ inline bool ray_is_good(const uint8_t *__restrict__ map, size_t pitch, int bx, int by, int ex, int ey,
                                   uint8_t cost) {
  const int ddx = ex - bx, ddy = ey - by;
  const int ax = abs(ddx), ay = abs(ddy);
  const bool y_major = ay > ax;
  const int den = y_major ? ay : ax, add = y_major ? ax : ay;
  const int sx = (ddx > 0) - (ddx < 0), sy = (ddy > 0) - (ddy < 0);
  int err = den / 2, x = bx, y = by;
  for (int i = 0; i < den; ++i) {
    if (__ldg(map + (size_t)y * pitch + x) == cost) return false;
    err += add;
    const bool step_minor = err >= den;
    if (step_minor) err -= den;
    if (y_major) {
      y += sy;
      if (step_minor) x += sx;
    } else {
      x += sx;
      if (step_minor) y += sy;
    }
  }
  return true;
}

struct CheckArgs {
  const uint8_t *map;
  size_t pitch;
  int size_x, size_y;
  int n;                      // vertices
  int fp[2 * kMaxVertices];   // footprint in cells (x0, y0, x1, y1, ...)
  const double *poses;        // n_poses x 3, or null: the footprint is used as given (n_poses == 1)
  size_t n_poses;
  uint8_t cost;
  uint8_t *ok;
};

struct MapView {
  const uint8_t *map;
  size_t pitch;
  int size_x, size_y;
  uint8_t cost;
};

__device__ bool check_one(const MapView a, const int n, const int2 *pt) {
  for (int i = 0; i < n; ++i)  // is_inside (:186-199)
    if (pt[i].x < 0 || pt[i].y < 0 || pt[i].x >= a.size_x || pt[i].y >= a.size_y) return false;
  if (n == 0) return true;  // check_without_area (:392-422)
  if (n == 1) return __ldg(a.map + ((size_t)pt[0].y * a.pitch + (size_t)pt[0].x)) != a.cost;
  if (n == 2) return ray_is_good(a.map, a.pitch, pt[0].x, pt[0].y, pt[1].x, pt[1].y, a.cost);
  // An open footprint gets its first vertex appended, which turns upstream's special cases for the closing
  // edge (:450-456 outline, :495-496 scan-line edges) into the last iteration of the same loops.
  const bool is_closed = pt[n - 1].x == pt[0].x && pt[n - 1].y == pt[0].y;
  const int m = is_closed ? n : n + 1;  // vertices of the closed ring
  auto vx = [&](int i) { return pt[i == n ? 0 : i].x; };
  auto vyy = [&](int i) { return pt[i == n ? 0 : i].y; };
  // check_outline (:432-458)
  for (int i = 1; i < m; ++i)
    if (!ray_is_good(a.map, a.pitch, vx(i - 1), vyy(i - 1), vx(i), vyy(i), a.cost)) return false;

  // edges, horizontal ones skipped (:487-496)
  XB xb[kMaxVertices];
  int lower_y[kMaxVertices];
  bool was_active[kMaxVertices];
  int nl = 0;
  for (int c = 1; c < m; ++c)
    if (vyy(c - 1) != vyy(c)) {
      xb_init(xb[nl], vx(c - 1), vyy(c - 1), vx(c), vyy(c));
      lower_y[nl] = vyy(c - 1);
      was_active[nl] = false;
      ++nl;
    }
  if (nl == 0) return true;
  // vertex_set (:503-510): (y, left line, right line), ordered by y, equal keys in insertion order
  int vy[kMaxVertices], vl[kMaxVertices], vr[kMaxVertices];
  int nv = 0;
  auto insert_vertex = [&](int y, int l, int r) {  // stable insertion
    int j = nv - 1;
    while (j >= 0 && vy[j] > y) {
      vy[j + 1] = vy[j], vl[j + 1] = vl[j], vr[j + 1] = vr[j];
      --j;
    }
    vy[j + 1] = y, vl[j + 1] = l, vr[j + 1] = r;
    ++nv;
  };
  for (int l = 0; l + 1 < nl; ++l) insert_vertex(lower_y[l + 1], l, l + 1);
  insert_vertex(lower_y[0], nl - 1, 0);

  int active[kMaxVertices], na = 0;
  int ymin = pt[0].y;
  for (int i = 1; i < n; ++i) ymin = min(ymin, pt[i].y);
  int yy = ymin - 1;  // :519
  for (int v = 0; v < nv;) {
    for (; yy < vy[v]; ++yy) {  // :549-556
      for (int p = 0; p + 1 < na; p += 2) {  // check_line (:349-378)
        const int lx = xb[active[p]].x_curr + 1, rx = xb[active[p + 1]].x_curr;
        const uint8_t *row = a.map + (size_t)yy * a.pitch;
        for (int x = lx; x < rx; ++x)
          if (__ldg(row + x) == a.cost) return false;
      }
      for (int p = 0; p < na; ++p) xb_advance(xb[active[p]]);
    }
    for (; v < nv && vy[v] == yy; ++v) {  // :562-581
      const int pair[2] = {vl[v], vr[v]};
      for (int k = 0; k < 2; ++k) {
        const int l = pair[k];
        if (was_active[l]) {
          int pos = 0;
          while (pos < na && active[pos] != l) ++pos;
          if (pos < na) active[pos] = active[--na];  // swap and pop
        } else {
          was_active[l] = true;
          active[na++] = l;
        }
      }
    }
    for (int i = 1; i < na; ++i) {  // sort by (x, dx) (:583), stable insertion sort
      const int o = active[i];
      int j = i - 1;
      while (j >= 0 && (xb[active[j]].x_curr > xb[o].x_curr ||
                        (xb[active[j]].x_curr == xb[o].x_curr && xb[active[j]].dx > xb[o].dx))) {
        active[j + 1] = active[j];
        --j;
      }
      active[j + 1] = o;
    }
  }
  return true;
}

__global__ void __launch_bounds__(128) k_check_footprint(const CheckArgs a) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.n_poses) return;
  int2 pt[kMaxVertices];
  if (a.poses) {  // dpose_goal_tolerance.cpp:388-392: round(T(x, y) R(theta) footprint)
    const double tx = a.poses[3 * i], ty = a.poses[3 * i + 1];
    double s, c;
    sincos(a.poses[3 * i + 2], &s, &c);
    for (int k = 0; k < a.n; ++k) {
      const double x = (double)a.fp[2 * k], y = (double)a.fp[2 * k + 1];
      const double rx = round(__dadd_rn(__dadd_rn(__dmul_rn(c, x), __dmul_rn(-s, y)), tx));
      const double ry = round(__dadd_rn(__dadd_rn(__dmul_rn(s, x), __dmul_rn(c, y)), ty));
      // poses far off the map: any vertex outside fails is_inside; clamp so the int cast is defined
      pt[k] = make_int2((int)fmin(fmax(rx, -2.0), (double)a.size_x + 1.0), (int)fmin(fmax(ry, -2.0), (double)a.size_y + 1.0));
    }
  } else {
    for (int k = 0; k < a.n; ++k) pt[k] = make_int2(a.fp[2 * k], a.fp[2 * k + 1]);
  }
  const MapView view = {a.map, a.pitch, a.size_x, a.size_y, a.cost};
  a.ok[i] = check_one(view, a.n, pt) ? 1 : 0;
}

}  // namespace
}  // namespace dpose

using namespace dpose;

static int check_impl(const dpose_map *map, const int32_t *fp, int n, const double *poses, size_t n_poses, bool dev_poses,
                      uint8_t cost, uint8_t *ok, bool dev_ok, cudaStream_t stream) {
  if (!map || !ok || (n && !fp)) return fail(DPOSE_EINVAL, "NULL argument");
  if (n < 0 || n > kMaxVertices) return fail(DPOSE_EINVAL, "footprint has %d vertices; at most %d supported", n, kMaxVertices);
  if (n_poses == 0) return DPOSE_OK;
  DeviceGuard guard(map->device);
  if (!guard.ok) return fail(DPOSE_ECUDA, "cudaSetDevice(%d) failed", map->device);
  CheckArgs a;
  a.map = map->d_data;
  a.pitch = map->pitch;
  a.size_x = (int)map->size_x;
  a.size_y = (int)map->size_y;
  a.n = n;
  for (int i = 0; i < 2 * n; ++i) a.fp[i] = fp[i];
  a.n_poses = n_poses;
  a.cost = cost;
  double *d_poses = nullptr;
  uint8_t *d_ok = nullptr;
  auto body = [&]() -> int {
    if (poses && !dev_poses) {
      DPOSE_CUDA(cudaMallocAsync(&d_poses, sizeof(double) * 3 * n_poses, stream));
      DPOSE_CUDA(cudaMemcpyAsync(d_poses, poses, sizeof(double) * 3 * n_poses, cudaMemcpyHostToDevice, stream));
    }
    if (!dev_ok) DPOSE_CUDA(cudaMallocAsync(&d_ok, n_poses, stream));
    a.poses = poses ? (dev_poses ? poses : d_poses) : nullptr;
    a.ok = dev_ok ? ok : d_ok;
    k_check_footprint<<<(unsigned)((n_poses + 127) / 128), 128, 0, stream>>>(a);
    DPOSE_LAUNCHED();
    if (!dev_ok) {
      DPOSE_CUDA(cudaMemcpyAsync(ok, d_ok, n_poses, cudaMemcpyDeviceToHost, stream));
      DPOSE_CUDA(cudaStreamSynchronize(stream));
    }
    return DPOSE_OK;
  };
  const int rc = body();
  if (d_poses) cudaFreeAsync(d_poses, stream);
  if (d_ok) cudaFreeAsync(d_ok, stream);
  return rc;
}

extern "C" {

int dpose_check_footprint(const dpose_map *map, const int32_t *footprint_xy, int n, uint8_t cost, int *ok) {
  if (!ok) return fail(DPOSE_EINVAL, "ok is NULL");
  uint8_t r = 0;
  DPOSE_TRY(check_impl(map, footprint_xy, n, nullptr, 1, false, cost, &r, false, nullptr));
  *ok = r;
  return DPOSE_OK;
}

int dpose_check_footprint_batch(const dpose_map *map, const int32_t *footprint_xy, int n, const double *poses,
                                size_t n_poses, uint8_t cost, uint8_t *ok) {
  if (n_poses && !poses) return fail(DPOSE_EINVAL, "poses is NULL");
  return check_impl(map, footprint_xy, n, poses, n_poses, false, cost, ok, false, nullptr);
}

int dpose_check_footprint_batch_device(const dpose_map *map, const int32_t *footprint_xy, int n, const double *d_poses,
                                       size_t n_poses, uint8_t cost, uint8_t *d_ok, void *stream) {
  if (n_poses && !d_poses) return fail(DPOSE_EINVAL, "d_poses is NULL");
  return check_impl(map, footprint_xy, n, d_poses, n_poses, true, cost, d_ok, true, (cudaStream_t)stream);
}

}  // extern "C"
