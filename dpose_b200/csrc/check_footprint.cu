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
//
// Footprints of up to kMaxVertices (64) vertices keep the per-pose state in thread-local arrays.  Larger ones (the
// reference's own benchmark checks 100- and 1000-vertex circles, perf/dpose_costmap.cpp:103-150) run the same code
// with the state in a global workspace, one slice per pose, poses processed in chunks that bound that workspace.
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
  uint8_t *was_active;
};

__host__ __device__ inline size_t scratch_bytes(int cap) {  // one pose, 8-byte aligned
  const size_t raw = (size_t)cap * (sizeof(XB) + sizeof(int2) + 5 * sizeof(int) + 1);
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
  s.was_active = reinterpret_cast<uint8_t *>(s.active + cap);
  return s;
}

// The scan-line sweep of upstream check_footprint (:487-583) for a ring of n >= 3 vertices: calls
// span(yy, lx, rx) for every pair of active edges on every scan line with lx < rx (the cells lx .. rx - 1 of row yy
// are the interior to search) and stops with false as soon as span() returns false.
template <typename SpanFn>
__device__ bool sweep_spans(const int n, const Scratch sc, SpanFn span) {
  const int2 *pt = sc.pt;
  XB *xb = sc.xb;
  int *lower_y = sc.lower_y, *vy = sc.vy, *vl = sc.vl, *vr = sc.vr, *active = sc.active;
  uint8_t *was_active = sc.was_active;
  // An open footprint gets its first vertex appended, which turns upstream's special cases for the closing
  // edge (:450-456 outline, :495-496 scan-line edges) into the last iteration of the same loops.
  const bool is_closed = pt[n - 1].x == pt[0].x && pt[n - 1].y == pt[0].y;
  const int m = is_closed ? n : n + 1;  // vertices of the closed ring
  auto vx = [&](int i) { return pt[i == n ? 0 : i].x; };
  auto vyy = [&](int i) { return pt[i == n ? 0 : i].y; };
  // edges, horizontal ones skipped (:487-496)
  int nl = 0;
  for (int c = 1; c < m; ++c)
    if (vyy(c - 1) != vyy(c)) {
      xb_init(xb[nl], vx(c - 1), vyy(c - 1), vx(c), vyy(c));
      lower_y[nl] = vyy(c - 1);
      was_active[nl] = 0;
      ++nl;
    }
  if (nl == 0) return true;
  // vertex_set (:503-510): (y, left line, right line), ordered by y, equal keys in insertion order
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

  int na = 0;
  int ymin = pt[0].y;
  for (int i = 1; i < n; ++i) ymin = min(ymin, pt[i].y);
  int yy = ymin - 1;  // :519
  for (int v = 0; v < nv;) {
    for (; yy < vy[v]; ++yy) {  // :549-556
      for (int p = 0; p + 1 < na; p += 2) {  // check_line (:349-378)
        const int lx = xb[active[p]].x_curr + 1, rx = xb[active[p + 1]].x_curr;
        if (lx < rx && !span(yy, lx, rx)) return false;
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
          was_active[l] = 1;
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
  const Scratch sc = {xb, pt, lower_y, vy, vl, vr, active, was_active};
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
// handful of candidates (the reference's own call: one pose per solve).  Here thread 0 runs the sweep WITHOUT
// touching the map and only records the spans it would search; meanwhile warps 1.. test the outline cells, every
// cell of every ray computed in closed form (cell k of a Bresenham ray: k steps on the major axis,
// floor((den / 2 + k * add) / den) on the minor one); then all warps search the recorded spans, lanes along x.
// The set of cells compared with `cost` is exactly the reference's, so the boolean is too.
struct Span {
  int y, lx, rx;
};

__host__ __device__ inline size_t coop_bytes(int cap, int span_cap) {  // one pose: Scratch | ray prefix | spans
  const size_t prefix = (((size_t)cap + 1) * sizeof(int) + 7) & ~(size_t)7;
  return scratch_bytes(cap) + prefix + (((size_t)span_cap * sizeof(Span) + 7) & ~(size_t)7);
}

template <int kCoopThreads>
__global__ void __launch_bounds__(kCoopThreads) k_check_footprint_coop(const CheckArgs a, const int *__restrict__ d_fp,
                                                                      unsigned char *__restrict__ ws, int span_cap) {
  const size_t i = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int kWarps = kCoopThreads / 32;
  // the threads that search the outline while thread 0 sweeps: the other warps, or (one-warp CTAs) the other lanes
  constexpr int kOutlineFirst = kCoopThreads > 32 ? 32 : 1;
  const int n = a.n, cap = n + 1;
  unsigned char *base = ws + i * coop_bytes(cap, span_cap);
  // the sweep's state is read and written a few times per scan line by one thread: up to kMaxVertices vertices it
  // sits in shared memory (latency), beyond that in this pose's slice of the workspace
  __shared__ __align__(8) unsigned char s_state[((kMaxVertices + 1) * (sizeof(XB) + sizeof(int2) + 5 * sizeof(int) + 1) + 7) & ~7];
  const Scratch sc = scratch_at(n <= kMaxVertices ? s_state : base, cap);
  int *prefix = reinterpret_cast<int *>(base + scratch_bytes(cap));
  Span *spans = reinterpret_cast<Span *>(base + scratch_bytes(cap) + ((((size_t)cap + 1) * sizeof(int) + 7) & ~(size_t)7));
  const int *fp = d_fp ? d_fp : a.fp;
  const MapView view = {a.map, a.pitch, a.size_x, a.size_y, a.cost};
  __shared__ int s_hit, s_nspans, s_warp_sum[kWarps];
  if (tid == 0) s_hit = 0, s_nspans = 0;

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
  // the rays to search: the single segment of a 2-vertex footprint, else the ring edges (closing edge included)
  const bool is_closed = n >= 3 && sc.pt[n - 1].x == sc.pt[0].x && sc.pt[n - 1].y == sc.pt[0].y;
  const int nedges = n == 2 ? 1 : (is_closed ? n : n + 1) - 1;
  auto edge_den = [&](int e) {
    const int2 p = sc.pt[e], q = sc.pt[e + 1 == n ? 0 : e + 1];
    return max(abs(q.x - p.x), abs(q.y - p.y));
  };
  // exclusive prefix sum of the ray lengths over the edges: thread-local chunk, then a block scan of the chunk sums
  const int chunk = (nedges + kCoopThreads - 1) / kCoopThreads;
  const int e0 = min(nedges, tid * chunk), e1 = min(nedges, e0 + chunk);
  int local = 0;
  for (int e = e0; e < e1; ++e) local += edge_den(e);
  int incl = local;
  for (int d = 1; d < 32; d <<= 1) {
    const int v = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= d) incl += v;
  }
  if (lane == 31) s_warp_sum[warp] = incl;
  __syncthreads();
  int before = incl - local;
  for (int w = 0; w < warp; ++w) before += s_warp_sum[w];
  for (int e = e0; e < e1; ++e) {
    prefix[e] = before;
    before += edge_den(e);
  }
  if (e1 == nedges && e0 < nedges) prefix[nedges] = before;
  __syncthreads();
  const int total = prefix[nedges];

  if (tid < kOutlineFirst) {
    if (tid == 0 && n >= 3) {  // the sweep: record spans; past the capacity, search them right here
      int count = 0;
      const bool good = sweep_spans(n, sc, [&](int yy, int lx, int rx) {
        if (count < span_cap) {
          spans[count++] = Span{yy, lx, rx};
          return true;
        }
        return span_is_good(view, yy, lx, rx);
      });
      s_nspans = count;
      if (!good) s_hit = 1;
    }
  } else {  // check_outline (:432-458) / the 2-vertex ray: cell t of the concatenated rays
    bool hit = false;
    for (int t = tid - kOutlineFirst; t < total && !hit; t += kCoopThreads - kOutlineFirst) {
      int lo = 0, hi = nedges - 1;  // last edge with prefix[e] <= t
      while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (prefix[mid] <= t) lo = mid; else hi = mid - 1;
      }
      const int2 p = sc.pt[lo], q = sc.pt[lo + 1 == n ? 0 : lo + 1];
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
  if (!s_hit) {
    const int nspans = s_nspans;
    bool hit = false;
    for (int sp = warp; sp < nspans && !hit; sp += kWarps) {
      const Span v = spans[sp];
      const uint8_t *row = a.map + (size_t)v.y * a.pitch;
      for (int x = v.lx + lane; x < v.rx; x += 32) hit |= __ldg(row + x) == a.cost;
    }
    if (hit) s_hit = 1;
  }
  __syncthreads();
  if (tid == 0) a.ok[i] = s_hit ? 0 : 1;
}

}  // namespace
}  // namespace dpose

using namespace dpose;

// batches up to this many poses use the CTA-per-pose kernel (DPOSE_CHECK_COOP_MAX overrides; 0 disables it).
// Measured on a B200 (tools/ab_check_footprint.py, profiles/r1_v21_check_footprint_ab.jsonl): 20x12-cell rectangle,
// 16384 poses 193 us against 212 us thread-per-pose (4096: 93 vs 187; 128: 56 vs 169); beyond that one thread per
// pose has more poses in flight and wins (65536: 311 vs 593 us).
static int coop_max_poses() {
  static const int v = [] {
    const char *e = getenv("DPOSE_CHECK_COOP_MAX");
    return e ? atoi(e) : 16384;
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
    return (size_t)(e ? std::max(1, atoi(e)) : 256) << 20;
  }();
  return v;
}

static int coop_threads(size_t n_poses) {
  static const int forced = [] {
    const char *e = getenv("DPOSE_CHECK_COOP_THREADS");
    return e ? atoi(e) : 0;
  }();
  if (forced == 32 || forced == 64 || forced == 128) return forced;
  return n_poses <= 1024 ? 128 : 32;
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
    // spans one pose can produce: at most one per pair of edges on every scan line of its bounding box
    long long height = 0;
    if (poses) {
      long long r2 = 0;
      for (int k = 0; k < n; ++k) r2 = std::max(r2, (long long)fp[2 * k] * fp[2 * k] + (long long)fp[2 * k + 1] * fp[2 * k + 1]);
      height = 2 * (long long)std::ceil(std::sqrt((double)r2)) + 4;
    } else if (n > 0) {
      int lo = fp[1], hi = fp[1];
      for (int k = 1; k < n; ++k) lo = std::min(lo, fp[2 * k + 1]), hi = std::max(hi, fp[2 * k + 1]);
      height = (long long)hi - lo + 3;
    }
    const int span_cap = (int)std::min<long long>(1 << 16, std::max<long long>(1, height * ((n + 1) / 2 + 1)));
    const size_t coop_per_pose = coop_bytes(n + 1, span_cap);
    const size_t kWorkspaceBudget = workspace_budget();  // bytes of per-pose state alive at once
    if (n_poses <= (size_t)coop_max_poses() && n_poses * coop_per_pose <= kWorkspaceBudget) {
      if (large) {
        DPOSE_CUDA(cudaMallocAsync(&d_fp, sizeof(int) * 2 * (size_t)n, stream));
        DPOSE_CUDA(cudaMemcpyAsync(d_fp, fp, sizeof(int) * 2 * (size_t)n, cudaMemcpyHostToDevice, stream));
      }
      DPOSE_CUDA(cudaMallocAsync(&d_ws, n_poses * coop_per_pose, stream));
      // few poses: wide CTAs finish each pose sooner; many: narrow ones keep more sweeps in flight per SM
      const int threads = coop_threads(n_poses);
      if (threads == 128)
        k_check_footprint_coop<128><<<(unsigned)n_poses, 128, 0, stream>>>(a, d_fp, d_ws, span_cap);
      else if (threads == 64)
        k_check_footprint_coop<64><<<(unsigned)n_poses, 64, 0, stream>>>(a, d_fp, d_ws, span_cap);
      else
        k_check_footprint_coop<32><<<(unsigned)n_poses, 32, 0, stream>>>(a, d_fp, d_ws, span_cap);
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
