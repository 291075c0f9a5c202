// check_footprint.cu — batched footprint collision check.
//
// Upstream check_footprint (dpose_costmap.cpp:460-582) validates ONE solved pose per Ipopt solve
// (dpose_goal_tolerance.cpp:386-399): the footprint is transformed by the pose, rounded to cells, and the
// outline (Bresenham rays) plus the scan-line interior between pairs of active edges (custom x-major /
// x-minor Bresenham iterators, :201-317) is searched for a cell equal to `cost`.  With multi-start in
// lock-step there are thousands of candidates per iteration.  Two kernels, chosen by the batch size:
//   * k_check_footprint / k_check_footprint_large — one thread per pose, the reference's algorithm step for step
//     (same iterators, same half-open vertex rule, same pairwise sweep l_x + 1 .. r_x - 1): large batches.
//     Up to kMaxVertices (64) vertices the per-pose state is thread-local; larger footprints (the reference's own
//     benchmark checks 100- and 1000-vertex circles, perf/dpose_costmap.cpp:103-150) keep it in a global workspace,
//     one slice per pose, poses processed in chunks that bound that workspace.
//   * k_check_footprint_coop — one CTA per pose, for up to 32768 poses: only the sweep's vertex events are serial,
//     the outline cells and the (scan line, edge pair) spans are evaluated by all threads with closed forms of the
//     Bresenham iterators (see the section further down).
// Either way the set of map cells compared with `cost` is exactly the reference's, so the boolean is identical.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>

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

// per-pose working arrays of the scan-line sweep, each with room for n + 1 entries
struct Scratch {
  XB *xb;
  int2 *pt;
  int *lower_y, *vy, *vl, *vr, *active;
  int *xk;  // x of every line on the current event's scan line (parallel form only; may be null otherwise)
  uint8_t *was_active;
};

__host__ __device__ inline size_t scratch_bytes(int cap) {  // one pose, 8-byte aligned
  const size_t raw = (size_t)cap * (sizeof(XB) + sizeof(int2) + 6 * sizeof(int) + 1);
  return (raw + 7) & ~(size_t)7;
}

__device__ inline Scratch scratch_at(unsigned char *base, int cap) {
  Scratch s;
  s.xb = reinterpret_cast<XB *>(base);
  s.pt = reinterpret_cast<int2 *>(s.xb + cap);
  s.lower_y = reinterpret_cast<int *>(s.pt + cap);
  s.vy = s.lower_y + cap;
  s.vl = s.vy + cap;
  s.vr = s.vl + cap;
  s.active = s.vr + cap;
  s.xk = s.active + cap;
  s.was_active = reinterpret_cast<uint8_t *>(s.xk + cap);
  return s;
}

// ---- the scan-line sweep of upstream check_footprint (:487-583), in pieces shared by the serial and the parallel form
struct SweepSetup {
  int nl, nv, ymin;
};

// edges (horizontal ones skipped, :487-496) and the vertex set (:503-510): (y, left line, right line), ordered by y,
// equal keys in insertion order.  An open footprint gets its first vertex appended, which turns upstream's special
// cases for the closing edge (:450-456 outline, :495-496 scan-line edges) into the last iteration of the same loops.
__device__ inline SweepSetup sweep_setup(const int n, const Scratch sc) {
  const int2 *pt = sc.pt;
  XB *xb = sc.xb;
  int *lower_y = sc.lower_y, *vy = sc.vy, *vl = sc.vl, *vr = sc.vr;
  const bool is_closed = pt[n - 1].x == pt[0].x && pt[n - 1].y == pt[0].y;
  const int m = is_closed ? n : n + 1;  // vertices of the closed ring
  auto vx = [&](int i) { return pt[i == n ? 0 : i].x; };
  auto vyy = [&](int i) { return pt[i == n ? 0 : i].y; };
  int nl = 0;
  for (int c = 1; c < m; ++c)
    if (vyy(c - 1) != vyy(c)) {
      xb_init(xb[nl], vx(c - 1), vyy(c - 1), vx(c), vyy(c));
      lower_y[nl] = vyy(c - 1);
      sc.was_active[nl] = 0;
      ++nl;
    }
  int nv = 0;
  if (nl) {
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
  }
  int ymin = pt[0].y;
  for (int i = 1; i < n; ++i) ymin = min(ymin, pt[i].y);
  return SweepSetup{nl, nv, ymin};
}

// vertex event v (:562-581): each of its two lines is taken out (swap with the last, pop) if it was in, else appended
__device__ inline void sweep_toggle(const Scratch sc, int v, int &na) {
  const int pair[2] = {sc.vl[v], sc.vr[v]};
  for (int k = 0; k < 2; ++k) {
    const int l = pair[k];
    if (sc.was_active[l]) {
      int pos = 0;
      while (pos < na && sc.active[pos] != l) ++pos;
      if (pos < na) sc.active[pos] = sc.active[--na];
    } else {
      sc.was_active[l] = 1;
      sc.active[na++] = l;
    }
  }
}

// the active lines by (x on the current scan line, dx) (:583), stable insertion sort
template <typename KeyFn>
__device__ inline void sweep_sort(const Scratch sc, int na, KeyFn xkey) {
  int *active = sc.active;
  for (int i = 1; i < na; ++i) {
    const int o = active[i];
    const int xo = xkey(o);
    const double dxo = sc.xb[o].dx;
    int j = i - 1;
    while (j >= 0) {
      const int xj = xkey(active[j]);
      if (!(xj > xo || (xj == xo && sc.xb[active[j]].dx > dxo))) break;
      active[j + 1] = active[j];
      --j;
    }
    active[j + 1] = o;
  }
}

// Serial form, for a ring of n >= 3 vertices: calls span(yy, lx, rx) for every pair of active edges on every scan line
// with lx < rx (the cells lx .. rx - 1 of row yy are the interior to search) and stops with false as soon as span()
// returns false.
template <typename SpanFn>
__device__ bool sweep_spans(const int n, const Scratch sc, SpanFn span) {
  XB *xb = sc.xb;
  const int *active = sc.active;
  const SweepSetup su = sweep_setup(n, sc);
  if (su.nl == 0) return true;
  int na = 0;
  int yy = su.ymin - 1;  // :519
  for (int v = 0; v < su.nv;) {
    for (; yy < sc.vy[v]; ++yy) {  // :549-556
      for (int p = 0; p + 1 < na; p += 2) {  // check_line (:349-378)
        const int lx = xb[active[p]].x_curr + 1, rx = xb[active[p + 1]].x_curr;
        if (lx < rx && !span(yy, lx, rx)) return false;
      }
      for (int p = 0; p < na; ++p) xb_advance(xb[active[p]]);
    }
    for (; v < su.nv && sc.vy[v] == yy; ++v) sweep_toggle(sc, v, na);
    sweep_sort(sc, na, [&](int l) { return xb[l].x_curr; });
  }
  return true;
}

// x of an edge after k advances from its initial state, in closed form: the four iterators (:201-317) add / subtract
// `add` until a multiple of `den` is crossed, so the number of wraps (x-minor) or of x steps (x-major) after k scan
// lines is a floor / ceiling of k (checked against the stepping iterators for 6.6e5 (edge, k) pairs).
__device__ inline int xb_at(const XB &b, int k) {
  if (k <= 0) return b.x_curr;
  if (k < 32768 && b.den < 32768) {  // every product below fits 31 bits: 32-bit divisions (the usual case)
    const unsigned h = (unsigned)b.den / 2u, kk = (unsigned)k, den = (unsigned)b.den, add = (unsigned)b.add;
    switch (b.kind) {
      case 0:
        return b.x_curr + b.x_sign * (int)((h + kk * add) / den);
      case 1:
        return b.x_curr - b.x_sign * (kk * add > h ? (int)((kk * add - h + den - 1u) / den) : 0);
      case 2:
        return b.x_curr + b.x_sign * (int)((kk * den - h + add - 1u) / add);
      default:
        return b.x_curr - b.x_sign * (int)((h + (kk - 1u) * den) / add + 1u);
    }
  }
  const long long h = b.den / 2, kk = k;
  switch (b.kind) {
    case 0:
      return b.x_curr + b.x_sign * (int)((h + kk * b.add) / b.den);
    case 1: {
      const long long t = kk * b.add - h;
      return b.x_curr - b.x_sign * (t > 0 ? (int)((t + b.den - 1) / b.den) : 0);
    }
    case 2:
      return b.x_curr + b.x_sign * (int)((kk * b.den - h + b.add - 1) / b.add);
    default:
      return b.x_curr - b.x_sign * (int)((h + (kk - 1) * b.den) / b.add + 1);
  }
}

// Parallel form, part 1 (one thread): only the vertex events are processed.  Between two events the order of the
// active edges is fixed, so an "epoch" (rows [y0, y1), the active list as sorted at y0) describes every span of those
// rows; part 2 lets all threads evaluate (row, pair) items with xb_at.  lower_y[] is reused for each line's lowest row.
struct Epochs {
  int *y0, *y1, *off, *np, *ipre;  // per epoch: rows, offset of its active list in alist, pairs, items before it
  int *alist;
  int alist_cap;
};

__device__ inline Epochs epochs_at(int *region, int region_ints, int cap) {
  Epochs e;
  e.y0 = region;
  e.y1 = e.y0 + cap;
  e.off = e.y1 + cap;
  e.np = e.off + cap;
  e.ipre = e.np + cap;  // cap + 1 entries
  e.alist = e.ipre + cap + 1;
  e.alist_cap = region_ints - (5 * cap + 1);
  return e;
}

// The event loop on a prepared vertex set (lines in sc.xb, their lowest rows in sc.lower_y, events in sc.vy / vl / vr,
// sc.was_active cleared).  Returns the number of epochs, or -1 if the active lists do not fit (the caller then sweeps
// serially).
__device__ int epochs_from_events(const Scratch sc, const Epochs ep, int nl, int nv, int ymin) {
  if (nl == 0) {
    ep.ipre[0] = 0;
    return 0;
  }
  int na = 0, ne = 0, used = 0, items = 0;
  int yy = ymin - 1;
  for (int v = 0; v < nv;) {
    const int ye = sc.vy[v];
    if (ye > yy && na >= 2) {
      if (used + na > ep.alist_cap) return -1;
      ep.y0[ne] = yy, ep.y1[ne] = ye, ep.off[ne] = used, ep.np[ne] = na / 2, ep.ipre[ne] = items;
      for (int p = 0; p < na; ++p) ep.alist[used + p] = sc.active[p];
      used += na;
      items += (ye - yy) * (na / 2);
      ++ne;
    }
    yy = ye;
    for (; v < nv && sc.vy[v] == yy; ++v) sweep_toggle(sc, v, na);
    for (int p = 0; p < na; ++p) {  // one closed-form evaluation per active line, not one per comparison
      const int l = sc.active[p];
      sc.xk[l] = xb_at(sc.xb[l], yy - sc.lower_y[l]);
    }
    sweep_sort(sc, na, [&](int l) { return sc.xk[l]; });
  }
  ep.ipre[ne] = items;
  return ne;
}

__device__ inline bool span_is_good(const MapView &a, int yy, int lx, int rx) {
  const uint8_t *row = a.map + (size_t)yy * a.pitch;
  for (int x = lx; x < rx; ++x)
    if (__ldg(row + x) == a.cost) return false;
  return true;
}

// one thread runs the whole check (the kernel for large batches)
__device__ bool check_one(const MapView a, const int n, const Scratch sc) {
  const int2 *pt = sc.pt;
  for (int i = 0; i < n; ++i)  // is_inside (:186-199)
    if (pt[i].x < 0 || pt[i].y < 0 || pt[i].x >= a.size_x || pt[i].y >= a.size_y) return false;
  if (n == 0) return true;  // check_without_area (:392-422)
  if (n == 1) return __ldg(a.map + ((size_t)pt[0].y * a.pitch + (size_t)pt[0].x)) != a.cost;
  if (n == 2) return ray_is_good(a.map, a.pitch, pt[0].x, pt[0].y, pt[1].x, pt[1].y, a.cost);
  const bool is_closed = pt[n - 1].x == pt[0].x && pt[n - 1].y == pt[0].y;
  const int m = is_closed ? n : n + 1;
  for (int i = 1; i < m; ++i) {  // check_outline (:432-458)
    const int2 p = pt[i - 1], q = pt[i == n ? 0 : i];
    if (!ray_is_good(a.map, a.pitch, p.x, p.y, q.x, q.y, a.cost)) return false;
  }
  return sweep_spans(n, sc, [&](int yy, int lx, int rx) { return span_is_good(a, yy, lx, rx); });
}

// footprint vertex k of pose i: round(T(x, y) R(theta) footprint) (dpose_goal_tolerance.cpp:388-392), or as given
__device__ inline int2 vertex_of(const CheckArgs &a, const int *fp, size_t i, int k, double tx, double ty, double s, double c) {
  if (!a.poses) return make_int2(fp[2 * k], fp[2 * k + 1]);
  const double x = (double)fp[2 * k], y = (double)fp[2 * k + 1];
  const double rx = round(__dadd_rn(__dadd_rn(__dmul_rn(c, x), __dmul_rn(-s, y)), tx));
  const double ry = round(__dadd_rn(__dadd_rn(__dmul_rn(s, x), __dmul_rn(c, y)), ty));
  // poses far off the map: any vertex outside fails is_inside; clamp so the int cast is defined
  return make_int2((int)fmin(fmax(rx, -2.0), (double)a.size_x + 1.0), (int)fmin(fmax(ry, -2.0), (double)a.size_y + 1.0));
}

__global__ void __launch_bounds__(128) k_check_footprint(const CheckArgs a) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.n_poses) return;
  XB xb[kMaxVertices];
  int2 pt[kMaxVertices];
  int lower_y[kMaxVertices], vy[kMaxVertices], vl[kMaxVertices], vr[kMaxVertices], active[kMaxVertices];
  uint8_t was_active[kMaxVertices];
  double tx = 0, ty = 0, s = 0, c = 1;
  if (a.poses) {
    tx = a.poses[3 * i], ty = a.poses[3 * i + 1];
    sincos(a.poses[3 * i + 2], &s, &c);
  }
  for (int k = 0; k < a.n; ++k) pt[k] = vertex_of(a, a.fp, i, k, tx, ty, s, c);
  const MapView view = {a.map, a.pitch, a.size_x, a.size_y, a.cost};
  const Scratch sc = {xb, pt, lower_y, vy, vl, vr, active, nullptr, was_active};
  a.ok[i] = check_one(view, a.n, sc) ? 1 : 0;
}

// more than kMaxVertices vertices: footprint and per-pose state in global memory; poses [first, first + count)
__global__ void __launch_bounds__(128) k_check_footprint_large(const CheckArgs a, const int *__restrict__ d_fp,
                                                               unsigned char *__restrict__ ws, size_t first, size_t count) {
  const size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= count) return;
  const size_t i = first + j;
  const int cap = a.n + 1;
  const Scratch sc = scratch_at(ws + j * scratch_bytes(cap), cap);
  double tx = 0, ty = 0, s = 0, c = 1;
  if (a.poses) {
    tx = a.poses[3 * i], ty = a.poses[3 * i + 1];
    sincos(a.poses[3 * i + 2], &s, &c);
  }
  for (int k = 0; k < a.n; ++k) sc.pt[k] = vertex_of(a, d_fp, i, k, tx, ty, s, c);
  const MapView view = {a.map, a.pitch, a.size_x, a.size_y, a.cost};
  a.ok[i] = check_one(view, a.n, sc) ? 1 : 0;
}

// ---- few poses: one CTA per pose ---------------------------------------------------------------------------------
// A single thread per pose leaves the GPU idle and pays one memory round trip per cell when there are only a
// handful of candidates (the reference's own call: one pose per solve).  Here thread 0 only processes the sweep's
// vertex events (build_epochs); meanwhile warps 1.. test the outline cells, every cell of every ray computed in closed
// form (cell k of a Bresenham ray: k steps on the major axis, floor((den / 2 + k * add) / den) on the minor one); then
// all warps take (scan line, edge pair) items of the epochs, place the two edges with xb_at and search the span between
// them, lanes along x.  The set of cells compared with `cost` is exactly the reference's, so the boolean is too.
__host__ __device__ inline size_t coop_bytes(int cap, int epoch_ints) {  // one pose: Scratch | ray prefix | epochs
  const size_t prefix = (((size_t)cap + 1) * sizeof(int) + 7) & ~(size_t)7;
  return scratch_bytes(cap) + prefix + (((size_t)epoch_ints * sizeof(int) + 7) & ~(size_t)7);
}

template <int kCoopThreads>
__global__ void __launch_bounds__(kCoopThreads) k_check_footprint_coop(const CheckArgs a, const int *__restrict__ d_fp,
                                                                      unsigned char *__restrict__ ws, int epoch_ints,
                                                                      int state_in_smem) {
  const size_t i = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int kWarps = kCoopThreads / 32;
  // the threads that search the outline while thread 0 walks the vertex events: the other warps, or (one-warp CTAs)
  // the other lanes
  constexpr int kOutlineFirst = kCoopThreads > 32 ? 32 : 1;
  const int n = a.n, cap = n + 1;
  unsigned char *base = ws + i * coop_bytes(cap, epoch_ints);
  // the per-pose state (vertices, lines, vertex set, active list) sits in dynamic shared memory when it fits there
  // (up to ~3000 vertices), else in this pose's slice of the workspace
  extern __shared__ __align__(8) unsigned char s_state[];
  const Scratch sc = scratch_at(state_in_smem ? s_state : base, cap);
  int *prefix = reinterpret_cast<int *>(base + scratch_bytes(cap));
  const Epochs ep = epochs_at(
      reinterpret_cast<int *>(base + scratch_bytes(cap) + ((((size_t)cap + 1) * sizeof(int) + 7) & ~(size_t)7)), epoch_ints, cap);
  const int *fp = d_fp ? d_fp : a.fp;
  const MapView view = {a.map, a.pitch, a.size_x, a.size_y, a.cost};
  __shared__ int s_hit, s_nepochs, s_ymin, s_warp_sum[kWarps], s_warp_sum2[kWarps];
  if (tid == 0) s_hit = 0, s_nepochs = 0, s_ymin = 0x7fffffff;

  double tx = 0, ty = 0, s = 0, c = 1;
  if (a.poses) {
    tx = a.poses[3 * i], ty = a.poses[3 * i + 1];
    sincos(a.poses[3 * i + 2], &s, &c);
  }
  bool inside = true;
  for (int k = tid; k < n; k += kCoopThreads) {
    const int2 p = vertex_of(a, fp, i, k, tx, ty, s, c);
    sc.pt[k] = p;
    inside &= p.x >= 0 && p.y >= 0 && p.x < a.size_x && p.y < a.size_y;  // is_inside (:186-199)
  }
  if (!__syncthreads_and(inside)) {
    if (tid == 0) a.ok[i] = 0;
    return;
  }
  if (n < 2) {  // check_without_area (:392-422): nothing, or one cell
    if (tid == 0) a.ok[i] = n == 0 || __ldg(a.map + ((size_t)sc.pt[0].y * a.pitch + (size_t)sc.pt[0].x)) != a.cost;
    return;
  }
  // the edges: the single segment of a 2-vertex footprint, else the ring (closing edge included).  Edge e runs from
  // vertex e to vertex e + 1 (the first vertex again after the last one of an open footprint).
  const bool is_closed = n >= 3 && sc.pt[n - 1].x == sc.pt[0].x && sc.pt[n - 1].y == sc.pt[0].y;
  const int nedges = n == 2 ? 1 : (is_closed ? n : n + 1) - 1;
  auto edge_from = [&](int e) { return sc.pt[e]; };
  auto edge_to = [&](int e) { return sc.pt[e + 1 == n ? 0 : e + 1]; };
  // two exclusive prefix sums over the edges in one pass (thread-local chunk, then a block scan of the chunk sums):
  // the ray lengths (outline cells before edge e) and the non-horizontal edges (the scan-line `line` index of edge e)
  const int chunk = (nedges + kCoopThreads - 1) / kCoopThreads;
  const int e0 = min(nedges, tid * chunk), e1 = min(nedges, e0 + chunk);
  int local = 0, local2 = 0, my_ymin = 0x7fffffff;
  for (int e = e0; e < e1; ++e) {
    const int2 p = edge_from(e), q = edge_to(e);
    local += max(abs(q.x - p.x), abs(q.y - p.y));
    local2 += p.y != q.y;
    my_ymin = min(my_ymin, min(p.y, q.y));
  }
  int incl = local, incl2 = local2;
  for (int d = 1; d < 32; d <<= 1) {
    const int v = __shfl_up_sync(0xffffffffu, incl, d), v2 = __shfl_up_sync(0xffffffffu, incl2, d);
    if (lane >= d) incl += v, incl2 += v2;
  }
  if (lane == 31) s_warp_sum[warp] = incl, s_warp_sum2[warp] = incl2;
  if (my_ymin != 0x7fffffff) atomicMin(&s_ymin, my_ymin);
  __syncthreads();
  int before = incl - local, line = incl2 - local2, nl = 0;
  for (int w = 0; w < kWarps; ++w) {
    if (w < warp) before += s_warp_sum[w], line += s_warp_sum2[w];
    nl += s_warp_sum2[w];
  }
  int *begin_y = sc.active;  // y of each line's first vertex, until the vertex set is built (then the active list)
  for (int e = e0; e < e1; ++e) {
    const int2 p = edge_from(e), q = edge_to(e);
    prefix[e] = before;
    before += max(abs(q.x - p.x), abs(q.y - p.y));
    if (n >= 3 && p.y != q.y) {  // a scan-line edge (:487-496)
      xb_init(sc.xb[line], p.x, p.y, q.x, q.y);
      begin_y[line] = p.y;
      sc.lower_y[line] = min(p.y, q.y);
      sc.was_active[line] = 0;
      ++line;
    }
  }
  if (e1 == nedges && e0 < nedges) prefix[nedges] = before;
  __syncthreads();
  const int total = prefix[nedges];
  if (n >= 3) {
    // the vertex set (:503-510): event k joins lines k and k + 1 at the first vertex of line k + 1, the last one joins
    // the last and the first line; ordered by y, equal keys in insertion order: each event's rank is counted directly
    for (int k = tid; k < nl; k += kCoopThreads) {
      const int yk = begin_y[k + 1 == nl ? 0 : k + 1];
      int rank = 0;
      for (int j = 0; j < nl; ++j) {
        const int yj = begin_y[j + 1 == nl ? 0 : j + 1];
        rank += yj < yk || (yj == yk && j < k);
      }
      sc.vy[rank] = yk, sc.vl[rank] = k, sc.vr[rank] = k + 1 == nl ? 0 : k + 1;
    }
    __syncthreads();  // begin_y is dead from here on: sc.active becomes the active list
  }

  if (tid < kOutlineFirst) {
    if (tid == 0 && n >= 3) {  // the sweep's vertex events; if the epochs do not fit, the whole sweep right here
      int ne = epochs_from_events(sc, ep, nl, nl, s_ymin);
      if (ne < 0) {
        ne = 0;
        if (!sweep_spans(n, sc, [&](int yy, int lx, int rx) { return span_is_good(view, yy, lx, rx); })) s_hit = 1;
      }
      s_nepochs = ne;
    }
  } else {  // check_outline (:432-458) / the 2-vertex ray: cell t of the concatenated rays
    bool hit = false;
    for (int t = tid - kOutlineFirst; t < total && !hit; t += kCoopThreads - kOutlineFirst) {
      int lo = 0, hi = nedges - 1;  // last edge with prefix[e] <= t
      while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (prefix[mid] <= t) lo = mid; else hi = mid - 1;
      }
      const int2 p = edge_from(lo), q = edge_to(lo);
      const int k = t - prefix[lo];
      const int ddx = q.x - p.x, ddy = q.y - p.y, ax = abs(ddx), ay = abs(ddy);
      const bool y_major = ay > ax;
      const int den = y_major ? ay : ax, add = y_major ? ax : ay;
      const int minor = (int)(((long long)(den / 2) + (long long)k * add) / den);
      const int sx = (ddx > 0) - (ddx < 0), sy = (ddy > 0) - (ddy < 0);
      const int x = p.x + sx * (y_major ? minor : k), y = p.y + sy * (y_major ? k : minor);
      hit = __ldg(a.map + (size_t)y * a.pitch + x) == a.cost;
    }
    if (hit) s_hit = 1;
  }
  __syncthreads();
  const int ne = s_nepochs;
  if (!s_hit && ne > 0) {  // check_line (:349-378) for every (scan line, pair of active edges) item
    // each lane places one item's two edges (the divisions run 32 wide), then the warp searches the 32 spans one
    // after the other, lanes along x
    const int items = ep.ipre[ne];
    bool hit = false;
    for (int t0 = warp * 32; t0 < items && !hit; t0 += kWarps * 32) {
      const int t = t0 + lane;
      int yy = 0, lx = 0, rx = 0;
      if (t < items) {
        int lo = 0, hi = ne - 1;  // last epoch with ipre[e] <= t
        while (lo < hi) {
          const int mid = (lo + hi + 1) >> 1;
          if (ep.ipre[mid] <= t) lo = mid; else hi = mid - 1;
        }
        const int local_t = t - ep.ipre[lo], np = ep.np[lo], r = local_t / np, p = 2 * (local_t - r * np);
        const int l_line = ep.alist[ep.off[lo] + p], r_line = ep.alist[ep.off[lo] + p + 1];
        yy = ep.y0[lo] + r;
        lx = xb_at(sc.xb[l_line], yy - sc.lower_y[l_line]) + 1;
        rx = xb_at(sc.xb[r_line], yy - sc.lower_y[r_line]);
      }
      unsigned todo = __ballot_sync(0xffffffffu, lx < rx);
      while (todo) {
        const int j = __ffs(todo) - 1;
        todo &= todo - 1;
        const int y_j = __shfl_sync(0xffffffffu, yy, j), l_j = __shfl_sync(0xffffffffu, lx, j), r_j = __shfl_sync(0xffffffffu, rx, j);
        const uint8_t *row = a.map + (size_t)y_j * a.pitch;
        for (int x = l_j + lane; x < r_j; x += 32) hit |= __ldg(row + x) == a.cost;
      }
      hit = __any_sync(0xffffffffu, hit);
    }
    if (hit) s_hit = 1;
  }
  __syncthreads();
  if (tid == 0) a.ok[i] = s_hit ? 0 : 1;
}

}  // namespace
}  // namespace dpose

using namespace dpose;

// batches up to this many poses use the CTA-per-pose kernel (test hook: the environment variable DPOSE_CHECK_COOP_MAX
// overrides, 0 disables it, so that tests can reach the thread-per-pose kernel with small batches).
// Measured on a B200 (tools/ab_check_footprint.py, profiles/r1_v26_check_footprint_ab.jsonl), microseconds per
// host-buffer call, thread-per-pose against CTA-per-pose: 20x12-cell rectangle 128 poses 162 / 38, 4096 poses 178 / 57,
// 16384 poses 208 / 127, 65536 poses 300 / 358; 40-vertex star 128 poses 1084 / 116, 4096 poses 1161 / 207, 16384 poses
// 1242 / 623.  Beyond ~32k poses one thread per pose has enough poses in flight and wins.
static int coop_max_poses() {
  static const int v = [] {
    const char *e = getenv("DPOSE_CHECK_COOP_MAX");
    return e ? atoi(e) : 32768;
  }();
  return v;
}

// per host thread: a page of mapped pinned memory the kernels of a small host-buffer call write their flags into
constexpr size_t kPinnedOk = 4096;
static uint8_t *pinned_ok() {
  thread_local uint8_t *p = [] {
    void *q = nullptr;
    if (cudaHostAlloc(&q, kPinnedOk, cudaHostAllocPortable | cudaHostAllocMapped) != cudaSuccess) {
      cudaGetLastError();
      q = nullptr;
    }
    return static_cast<uint8_t *>(q);
  }();
  return p;
}

// device memory the per-pose working state of one call may take (DPOSE_CHECK_WS_MB overrides; tests use it to force
// the chunked path)
static size_t workspace_budget() {
  static const size_t v = [] {
    const char *e = getenv("DPOSE_CHECK_WS_MB");
    return (size_t)(e ? std::max(1, atoi(e)) : 512) << 20;
  }();
  return v;
}

static int coop_threads(size_t n_poses) {
  return n_poses <= 1024 ? 128 : 32;  // one pose: 128 threads 100 us / 32 threads 114 us (star); 4096 poses: 351 / 207
}

static int check_impl(const dpose_map *map, const int32_t *fp, int n, const double *poses, size_t n_poses, bool dev_poses,
                      uint8_t cost, uint8_t *ok, bool dev_ok, cudaStream_t stream) {
  if (!map || !ok || (n && !fp)) return fail(DPOSE_EINVAL, "NULL argument");
  if (n < 0) return fail(DPOSE_EINVAL, "negative vertex count");
  if (n_poses == 0) return DPOSE_OK;
  DeviceGuard guard(map->device);
  if (!guard.ok) return fail(DPOSE_ECUDA, "cudaSetDevice(%d) failed", map->device);
  CheckArgs a;
  a.map = map->d_data;
  a.pitch = map->pitch;
  a.size_x = (int)map->size_x;
  a.size_y = (int)map->size_y;
  a.n = n;
  const bool large = n > kMaxVertices;
  if (!large)
    for (int i = 0; i < 2 * n; ++i) a.fp[i] = fp[i];
  a.n_poses = n_poses;
  a.cost = cost;
  double *d_poses = nullptr;
  uint8_t *d_ok = nullptr;
  int *d_fp = nullptr;
  unsigned char *d_ws = nullptr;
  auto body = [&]() -> int {
    if (poses && !dev_poses) {
      DPOSE_CUDA(cudaMallocAsync(&d_poses, sizeof(double) * 3 * n_poses, stream));
      DPOSE_CUDA(cudaMemcpyAsync(d_poses, poses, sizeof(double) * 3 * n_poses, cudaMemcpyHostToDevice, stream));
    }
    // results of a small host-buffer call go straight into mapped pinned memory (no device buffer, no copy back)
    uint8_t *h_ok = !dev_ok && n_poses <= kPinnedOk ? pinned_ok() : nullptr;
    if (!dev_ok && !h_ok) DPOSE_CUDA(cudaMallocAsync(&d_ok, n_poses, stream));
    a.poses = poses ? (dev_poses ? poses : d_poses) : nullptr;
    a.ok = dev_ok ? ok : (h_ok ? h_ok : d_ok);
    // the epochs of one pose: five ints per vertex event plus the active lists (at most every line at every event;
    // beyond 2^20 entries the kernel sweeps serially instead)
    const long long cap_ll = (long long)n + 1;
    const int epoch_ints = (int)(5 * cap_ll + 1 + std::min<long long>(cap_ll * cap_ll, 1 << 20));
    const size_t coop_per_pose = coop_bytes(n + 1, epoch_ints);
    const size_t kWorkspaceBudget = workspace_budget();  // bytes of per-pose state alive at once
    if (n_poses <= (size_t)coop_max_poses() && n_poses * coop_per_pose <= kWorkspaceBudget) {
      if (large) {
        DPOSE_CUDA(cudaMallocAsync(&d_fp, sizeof(int) * 2 * (size_t)n, stream));
        DPOSE_CUDA(cudaMemcpyAsync(d_fp, fp, sizeof(int) * 2 * (size_t)n, cudaMemcpyHostToDevice, stream));
      }
      DPOSE_CUDA(cudaMallocAsync(&d_ws, n_poses * coop_per_pose, stream));
      // few poses: wide CTAs finish each pose sooner; many: narrow ones keep more sweeps in flight per SM
      const int threads = coop_threads(n_poses);
      // the per-pose state in shared memory if it fits (beyond 48 KB the kernel attribute is raised first)
      const size_t state_bytes = scratch_bytes(n + 1);
      const bool in_smem = state_bytes <= (size_t)200 << 10;
      const size_t dyn = in_smem ? state_bytes : 0;
      auto launch = [&](auto kern) -> int {
        if (dyn > ((size_t)48 << 10)) DPOSE_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
        kern<<<(unsigned)n_poses, threads, dyn, stream>>>(a, d_fp, d_ws, epoch_ints, in_smem ? 1 : 0);
        return DPOSE_OK;
      };
      if (threads == 128)
        DPOSE_TRY(launch(k_check_footprint_coop<128>));
      else if (threads == 64)
        DPOSE_TRY(launch(k_check_footprint_coop<64>));
      else
        DPOSE_TRY(launch(k_check_footprint_coop<32>));
      DPOSE_LAUNCHED();
    } else if (!large) {
      k_check_footprint<<<(unsigned)((n_poses + 127) / 128), 128, 0, stream>>>(a);
      DPOSE_LAUNCHED();
    } else {
      const size_t per_pose = scratch_bytes(n + 1);
      const size_t chunk = std::max<size_t>(1, std::min(n_poses, kWorkspaceBudget / per_pose));
      DPOSE_CUDA(cudaMallocAsync(&d_fp, sizeof(int) * 2 * (size_t)n, stream));
      // the footprint is pageable host memory owned by the caller: the copy is staged before the call returns
      DPOSE_CUDA(cudaMemcpyAsync(d_fp, fp, sizeof(int) * 2 * (size_t)n, cudaMemcpyHostToDevice, stream));
      DPOSE_CUDA(cudaMallocAsync(&d_ws, chunk * per_pose, stream));
      for (size_t first = 0; first < n_poses; first += chunk) {
        const size_t count = std::min(chunk, n_poses - first);
        k_check_footprint_large<<<(unsigned)((count + 127) / 128), 128, 0, stream>>>(a, d_fp, d_ws, first, count);
        DPOSE_LAUNCHED();
      }
    }
    if (h_ok) {
      DPOSE_CUDA(cudaStreamSynchronize(stream));
      std::memcpy(ok, h_ok, n_poses);
    } else if (!dev_ok) {
      DPOSE_CUDA(cudaMemcpyAsync(ok, d_ok, n_poses, cudaMemcpyDeviceToHost, stream));
      DPOSE_CUDA(cudaStreamSynchronize(stream));
    }
    return DPOSE_OK;
  };
  const int rc = body();
  if (d_poses) cudaFreeAsync(d_poses, stream);
  if (d_ok) cudaFreeAsync(d_ok, stream);
  if (d_fp) cudaFreeAsync(d_fp, stream);
  if (d_ws) cudaFreeAsync(d_ws, stream);
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
