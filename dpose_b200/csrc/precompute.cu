// precompute.cu — cost_data precompute on the GPU (replaces _get_cost, dpose_core.cpp:55-114).
//
// One fused, single-CTA kernel (the images are a few KB: 17x25 ... ~200x200 pixels, so the
// work is latency-bound; one launch beats a chain of ten).  Phases, separated by
// __syncthreads():
//   1. outline raster  : one thread per polygon edge, 8-connected Bresenham drawn left to
//                        right (the rule of cv::polylines / cv::fillPoly's outline)
//   2. even-odd fill   : one warp per scanline, 16.16 fixed-point edge crossings,
//                        ballot-compacted and rank-sorted inside the warp
//   3. exact EDT       : column pass (one thread per column, two sweeps), then the row pass
//                        d2(x) = min_i g(i)^2 + (x - i)^2 in one of two exact integer forms:
//                        tables up to kScanCols columns — each warp owns 32 output pixels and scans the
//                        row's parabolas tile by tile, broadcasting candidates with __shfl_sync and
//                        skipping tiles that cannot beat the current best (a handful of tiles: lowest
//                        latency); wider tables — Meijster's lower envelope, one thread per row: a
//                        forward scan keeps the stack of parabolas that own a piece of the row (their
//                        indices s and the first column t each one wins), a backward scan reads the
//                        winners off, O(cols) per row instead of O(cols^2 / 32).
//                        All integer -> squared distances are exact (bit-exact gate).
//   4. cost shaping    : sqrt, outside *= -0.01f, global min (warp shuffles + one smem
//                        atomic), shift, separable [1 2 1]/4 blur with REFLECT_101 borders.
//                        Every float op uses an explicit round-to-nearest intrinsic so the
//                        compiler cannot contract it; the table is bit-identical to OpenCV's.
//   5. patch tables    : FP64 bilinear coefficients per interpolation patch (the
//                        reference-consistent form of "gradient tables").
#include <algorithm>
#include <cstring>
#include <new>

#include "common.cuh"

namespace dpose {
namespace {

constexpr int kThreads = 1024;
constexpr int kWarps = kThreads / 32;
constexpr int kInfG = 1 << 20;   // "no outline pixel in this column"
constexpr int kBigD2 = 1 << 29;  // its square stand-in; (x-i)^2 < 2^30 keeps sums in int32

constexpr int kScanCols = 160;   // widest table whose row pass is the warp-shuffle scan (B200: crossover ~150-200)
constexpr int kParamVerts = 64;  // footprints up to this many vertices travel in the kernel parameters

struct PrecomputeArgs {
  const int2 *fp;  // shifted footprint vertices (null: fp_param)
  int2 fp_param[kParamVerts];
  int n;
  int rows, cols;
  uint8_t *outline, *mask;
  int *g;              // column-pass distances
  int *d2;             // squared EDT
  float *pre, *tmp, *cost;
  float *cost_host;  // mapped pinned copy of the table for the caller (or null)
  int want_stages;   // kSmem only: copy outline / mask / d2 / pre-blur table out to their global images at the end
  PatchCoef *coef;
  long long *scratch;  // kWarps * 2 * n crossings
};

__device__ __forceinline__ int reflect101(int i, int n) {
  if (n == 1) return 0;
  if (i < 0) return -i;
  if (i >= n) return 2 * (n - 1) - i;
  return i;
}

// cv::Line, 8-connected, left-to-right LineIterator: err = major - 2*minor, diagonal step when err < 0
__device__ void draw_edge(uint8_t *a, uint8_t *b, int cols, int rows, int x1, int y1, int x2,
                          int y2) {
  int dx = x2 - x1, dy = y2 - y1, sy = 1;
  if (dx < 0) {
    dx = -dx;
    dy = -dy;
    x1 = x2;
    y1 = y2;
  }
  if (dy < 0) {
    dy = -dy;
    sy = -1;
  }
  const bool vert = dy > dx;
  const int major = vert ? dy : dx, minor = vert ? dx : dy;
  int err = major - 2 * minor;
  int x = x1, y = y1;
  for (int i = 0; i <= major; ++i) {
    if (x >= 0 && x < cols && y >= 0 && y < rows) {
      a[y * cols + x] = 0;
      b[y * cols + x] = 0;
    }
    const bool diag = err < 0;
    err += -2 * minor + (diag ? 2 * major : 0);
    if (vert) {
      y += sy;
      if (diag) x += 1;
    } else {
      x += 1;
      if (diag) y += sy;
    }
  }
}

// kSmem: every intermediate image lives in shared memory (18 bytes per pixel + the fill's scratch: tables up to ~11 000
// pixels, i.e. all of the reference's fixtures).  The phases are separated by block barriers and chained through these
// images, so with the images in global memory every phase starts with an L2 round trip per access (the CTA's own stores
// invalidate its L1 lines); in shared memory the same accesses cost ~30 cycles.  Only the table, the coefficients and
// — on request — the debug stages leave the SM.
template <bool kSmem>
__global__ void __launch_bounds__(kThreads, 1) k_precompute(PrecomputeArgs p) {
  extern __shared__ __align__(16) unsigned char s_dyn[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int rows = p.rows, cols = p.cols, np = rows * cols, n = p.n;
  __shared__ unsigned s_maxmag;
  uint8_t *outline = p.outline, *mask = p.mask;
  int *g = p.g, *d2 = p.d2;
  float *pre = p.pre, *tmp = p.tmp;
  long long *scratch = p.scratch;
  if (kSmem) {  // same order as smem_bytes() below
    const size_t np4 = ((size_t)np + 3) & ~(size_t)3;
    g = reinterpret_cast<int *>(s_dyn);
    d2 = g + np4;
    pre = reinterpret_cast<float *>(d2 + np4);
    tmp = pre + np4;
    scratch = reinterpret_cast<long long *>(tmp + np4);
    outline = reinterpret_cast<uint8_t *>(scratch + (size_t)kWarps * 2 * n);
    mask = outline + np4;
  }
  // vertices that travelled in the kernel parameters go to shared memory first: the parameters live in the constant
  // bank, which serves one address per warp access — lanes reading different vertices (the fill takes one edge per lane)
  // serialise 32-fold (the 40-vertex footprint lost 7 us that way)
  __shared__ int2 s_fp[kParamVerts];
  if (!p.fp)
    for (int e = tid; e < n; e += kThreads) s_fp[e] = p.fp_param[e];
  const int2 *verts = p.fp ? p.fp : s_fp;
  auto vertex = [&](int e) { return verts[e]; };

  // ---- phase 0: white images ---------------------------------------------------------
  for (int i = tid; i < np; i += kThreads) {
    outline[i] = 255;
    mask[i] = 255;
  }
  if (tid == 0) s_maxmag = 0u;
  __syncthreads();

  // ---- phase 1: outline (polylines == fillPoly's outline) ------------------------------
  for (int e = tid; e < n; e += kThreads) {
    const int2 a = vertex(e == 0 ? n - 1 : e - 1), b = vertex(e);
    draw_edge(outline, mask, cols, rows, a.x, a.y, b.x, b.y);
  }
  __syncthreads();

  // ---- phase 2: even-odd scanline fill into mask ---------------------------------------
  {
    long long *xs = scratch + (size_t)warp * 2 * n, *sorted = xs + n;
    for (int y = warp; y < rows; y += kWarps) {
      int na = 0;
      for (int e0 = 0; e0 < n; e0 += 32) {
        const int e = e0 + lane;
        bool active = false;
        long long x = 0;
        if (e < n) {
          const int2 a = vertex(e == 0 ? n - 1 : e - 1), b = vertex(e);
          if (a.y != b.y) {
            const long long dx = ((long long)(b.x - a.x) * 65536) / (long long)(b.y - a.y);
            const int y0 = a.y < b.y ? a.y : b.y, y1 = a.y < b.y ? b.y : a.y;
            const long long xstart = (long long)(a.y < b.y ? a.x : b.x) * 65536;
            active = y0 <= y && y < y1;
            x = xstart + (long long)(y - y0) * dx;
          }
        }
        const unsigned m = __ballot_sync(0xffffffffu, active);
        if (active) xs[na + __popc(m & ((1u << lane) - 1u))] = x;
        na += __popc(m);
      }
      __syncwarp();
      for (int i = lane; i < na; i += 32) {
        const long long v = xs[i];
        int rank = 0;
        for (int j = 0; j < na; ++j) {
          const long long w = xs[j];
          rank += (w < v) || (w == v && j < i);
        }
        sorted[rank] = v;
      }
      __syncwarp();
      for (int k = 0; k + 1 < na; k += 2) {
        int x1 = (int)((sorted[k] + 32768) >> 16);  // left end rounded, right end floored
        int x2 = (int)(sorted[k + 1] >> 16);
        if (x1 < cols && x2 >= 0) {
          x1 = x1 < 0 ? 0 : x1;
          x2 = x2 >= cols ? cols - 1 : x2;
          for (int x = x1 + lane; x <= x2; x += 32) mask[y * cols + x] = 0;
        }
      }
      __syncwarp();
    }
  }
  __syncthreads();

  // ---- phase 3a: EDT column pass ---------------------------------------------------------
  for (int x = tid; x < cols; x += kThreads) {
    int last = -1;
    for (int y = 0; y < rows; ++y) {
      if (outline[y * cols + x] == 0) last = y;
      g[y * cols + x] = last < 0 ? kInfG : y - last;
    }
    last = -1;
    for (int y = rows - 1; y >= 0; --y) {
      if (outline[y * cols + x] == 0) last = y;
      if (last >= 0 && last - y < g[y * cols + x]) g[y * cols + x] = last - y;
    }
  }
  __syncthreads();

  // ---- phase 3b: EDT row pass ----------------------------------------------------------------------
  if (cols <= kScanCols) {  // warp-shuffle scan over the row's parabolas
    const int segs = (cols + 31) / 32;
    for (int item = warp; item < rows * segs; item += kWarps) {
      const int y = item / segs, s = item - y * segs;
      const int x = s * 32 + lane;
      const int *gr = g + y * cols;
      int best = 0x7fffffff;
      for (int t = 0; t < segs; ++t) {
        // smallest |x - i| between this segment and tile t
        const int sep = t > s ? t - s : s - t;
        const int gap = sep > 0 ? (sep - 1) * 32 + 1 : 0;
        if (__all_sync(0xffffffffu, best <= gap * gap)) continue;
        const int i = t * 32 + lane;
        const int gi = i < cols ? gr[i] : kInfG;
        const int g2 = gi >= kInfG ? kBigD2 : gi * gi;
#pragma unroll 8
        for (int j = 0; j < 32; ++j) {
          const int cand = __shfl_sync(0xffffffffu, g2, j);
          const int dxi = x - (t * 32 + j);
          best = min(best, dxi * dxi + cand);
        }
      }
      if (x < cols) d2[y * cols + x] = best;
    }
  } else {  // Meijster's lower envelope, one thread per row
    // the stacks live in the (still unused) float scratch images, [q][row] so that the threads of a warp, whose stack
    // depths stay close, touch neighbouring words; the top of the stack is carried in registers
    int *st_s = reinterpret_cast<int *>(pre), *st_t = reinterpret_cast<int *>(tmp);
    for (int y = tid; y < rows; y += kThreads) {
      const int *gr = g + y * cols;
      auto g2_of = [&](int i) {
        const int gi = gr[i];
        return gi >= kInfG ? kBigD2 : gi * gi;
      };
      int q = 0, s_top = 0, t_top = 0, g2_top = g2_of(0);
      st_s[y] = 0, st_t[y] = 0;
      for (int u = 1; u < cols; ++u) {
        const int g2u = g2_of(u);
        // pop every parabola that u beats already at the column where it starts to own the row
        while (q >= 0) {
          const int a = t_top - s_top, b = t_top - u;
          if (a * a + g2_top <= b * b + g2u) break;
          --q;
          if (q >= 0) {
            s_top = st_s[q * rows + y];
            t_top = st_t[q * rows + y];
            g2_top = g2_of(s_top);
          }
        }
        if (q < 0) {
          q = 0, s_top = u, t_top = 0, g2_top = g2u;
          st_s[y] = u, st_t[y] = 0;
        } else {
          // first column where u is strictly better than s_top: 1 + floor((u^2 - s^2 + g2u - g2s) / (2 (u - s)));
          // the numerator is >= 0 here (s_top still wins at t_top >= 0) and < 2^31 (|g2| <= 2^29, u < 2^9)
          const int w = 1 + (u * u - s_top * s_top + g2u - g2_top) / (2 * (u - s_top));
          if (w < cols) {
            ++q;
            s_top = u, t_top = w, g2_top = g2u;
            st_s[q * rows + y] = u, st_t[q * rows + y] = w;
          }
        }
      }
      for (int u = cols - 1; u >= 0; --u) {
        const int dxi = u - s_top;
        d2[y * cols + u] = dxi * dxi + g2_top;
        if (u == t_top && q > 0) {
          --q;
          s_top = st_s[q * rows + y];
          t_top = st_t[q * rows + y];
          g2_top = g2_of(s_top);
        }
      }
    }
  }
  __syncthreads();

  // ---- phase 4a: min over the scaled outside part (all values <= 0) ------------------------
  {
    unsigned mag = 0u;
    for (int i = tid; i < np; i += kThreads) {
      const float d = __fsqrt_rn(__int2float_rn(d2[i]));
      const float o = mask[i] != 0 ? __fmul_rn(d, -0.01f) : 0.0f;
      mag = max(mag, __float_as_uint(o) & 0x7fffffffu);
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) mag = max(mag, __shfl_xor_sync(0xffffffffu, mag, off));
    if (lane == 0) atomicMax(&s_maxmag, mag);
  }
  __syncthreads();
  const float mn = -__uint_as_float(s_maxmag);

  // ---- phase 4b: shift ----------------------------------------------------------------------
  for (int i = tid; i < np; i += kThreads) {
    const float d = __fsqrt_rn(__int2float_rn(d2[i]));
    const float v = mask[i] != 0 ? __fmul_rn(d, -0.01f) : d;
    pre[i] = __fsub_rn(v, mn);
  }
  __syncthreads();

  // ---- phase 4c: blur rows then columns ([0.25 0.5 0.25], BORDER_REFLECT_101) -----------------
  for (int i = tid; i < np; i += kThreads) {
    const int y = i / cols, x = i - y * cols;
    const float a = pre[y * cols + reflect101(x - 1, cols)];
    const float b = pre[i];
    const float c = pre[y * cols + reflect101(x + 1, cols)];
    tmp[i] = __fadd_rn(__fmul_rn(__fadd_rn(a, c), 0.25f), __fmul_rn(b, 0.5f));
  }
  __syncthreads();
  for (int i = tid; i < np; i += kThreads) {
    const int y = i / cols, x = i - y * cols;
    const float a = tmp[reflect101(y - 1, rows) * cols + x];
    const float b = tmp[i];
    const float c = tmp[reflect101(y + 1, rows) * cols + x];
    const float v = __fadd_rn(__fmul_rn(__fadd_rn(a, c), 0.25f), __fmul_rn(b, 0.5f));
    p.cost[i] = v;
    if (p.cost_host) p.cost_host[i] = v;  // straight into the caller's (mapped pinned) copy: no copy back
  }
  __syncthreads();

  // ---- phase 5: per-patch FP64 coefficients, indexed by k_upper ---------------------------------
  for (int i = tid; i < np; i += kThreads) {
    const int uy = i / cols, ux = i - uy * cols;
    PatchCoef c = {0.0, 0.0, 0.0, 0.0};
    if (ux >= 2 && uy >= 2 && ux <= cols - 2 && uy <= rows - 2) {
      const double m00 = p.cost[(uy - 1) * cols + ux - 1], m10 = p.cost[(uy - 1) * cols + ux];
      const double m01 = p.cost[uy * cols + ux - 1], m11 = p.cost[uy * cols + ux];
      c.a00 = m00;
      c.a10 = m10 - m00;
      c.a01 = m01 - m00;
      c.a11 = (m11 - m10) - (m01 - m00);
    }
    p.coef[i] = c;
  }
  if (kSmem && p.want_stages) {  // debug / parity aid: the intermediate images leave the SM only on request
    for (int i = tid; i < np; i += kThreads) {
      p.outline[i] = outline[i];
      p.mask[i] = mask[i];
      p.d2[i] = d2[i];
      p.pre[i] = pre[i];
    }
  }
}

// shared-memory bytes of the kSmem variant (layout in the kernel)
size_t smem_bytes(size_t np, int n) {
  const size_t np4 = (np + 3) & ~(size_t)3;
  return 16 * np4 + sizeof(long long) * kWarps * 2 * (size_t)n + 2 * np4;
}

}  // namespace

// Device memory of one construction: the persistent part (cost table + patch coefficients) and a scratch arena with
// the footprint and every intermediate image, both from the stream-ordered pool (no cudaMalloc after the first call
// of a process).  The scratch is freed as soon as the kernel has run; dpose_pg_stages — a debugging / parity aid, like
// the /tmp/*.jpg dumps of the reference's assert-enabled build (dpose_core.cpp:76,85,135) — re-runs the kernel on a
// scratch arena of its own and downloads the intermediates from there.
namespace {

struct Arena {
  size_t o_fp, o_outline, o_mask, o_g, o_d2, o_pre, o_tmp, o_scratch, total;
  Arena(size_t np, int n) {
    auto up = [](size_t b) { return (b + 255) & ~(size_t)255; };
    o_fp = 0;
    o_outline = o_fp + up(sizeof(int2) * (size_t)n);
    o_mask = o_outline + up(np);
    o_g = o_mask + up(np);
    o_d2 = o_g + up(sizeof(int) * np);
    o_pre = o_d2 + up(sizeof(int) * np);
    o_tmp = o_pre + up(sizeof(float) * np);
    o_scratch = o_tmp + up(sizeof(float) * np);
    total = o_scratch + up(sizeof(long long) * kWarps * 2 * (size_t)n);
  }
};

// per host thread: timing events and a mapped pinned page the kernel writes small tables into (grow-only)
struct ThreadCtx {
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  float *h_table = nullptr;
  size_t cap = 0;
  int device = -1;
};

int launch_precompute(const dpose_pg *pg, const int32_t *fp_xy, int n, unsigned char *base, const Arena &ar, float *d_cost,
                      PatchCoef *d_coef, float *cost_host, bool want_stages, cudaEvent_t e0, cudaEvent_t e1) {
  PrecomputeArgs a;
  a.n = n;
  if (n <= kParamVerts) {
    a.fp = nullptr;
    for (int i = 0; i < n; ++i) a.fp_param[i] = make_int2(fp_xy[2 * i], fp_xy[2 * i + 1]);
  } else {
    DPOSE_CUDA(cudaMemcpyAsync(base + ar.o_fp, fp_xy, sizeof(int2) * n, cudaMemcpyHostToDevice, nullptr));
    a.fp = reinterpret_cast<const int2 *>(base + ar.o_fp);
  }
  a.rows = pg->rows;
  a.cols = pg->cols;
  a.outline = base + ar.o_outline;
  a.mask = base + ar.o_mask;
  a.g = reinterpret_cast<int *>(base + ar.o_g);
  a.d2 = reinterpret_cast<int *>(base + ar.o_d2);
  a.pre = reinterpret_cast<float *>(base + ar.o_pre);
  a.tmp = reinterpret_cast<float *>(base + ar.o_tmp);
  a.cost = d_cost;
  a.cost_host = cost_host;
  a.coef = d_coef;
  a.scratch = reinterpret_cast<long long *>(base + ar.o_scratch);
  a.want_stages = want_stages ? 1 : 0;
  const int smem_optin = device_info(pg->device).smem_optin;
  const size_t dyn = smem_bytes((size_t)pg->rows * pg->cols, n);
  if (e0) DPOSE_CUDA(cudaEventRecord(e0, nullptr));
  if (dyn + 2048 <= (size_t)smem_optin) {
    if (dyn > ((size_t)40 << 10))
      DPOSE_CUDA(cudaFuncSetAttribute(k_precompute<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
    k_precompute<true><<<1, kThreads, dyn>>>(a);
  } else {
    k_precompute<false><<<1, kThreads>>>(a);
  }
  DPOSE_LAUNCHED();
  if (e1) DPOSE_CUDA(cudaEventRecord(e1, nullptr));
  return DPOSE_OK;
}

}  // namespace

int precompute_run(dpose_pg *pg, const int32_t *fp_xy, int n) {
  const size_t np = (size_t)pg->rows * pg->cols;
  const Arena ar(np, n);
  thread_local ThreadCtx tc;
  if (tc.device != pg->device) {  // a thread that moves between devices rebuilds its context (rare)
    if (tc.e0) cudaEventDestroy(tc.e0);
    if (tc.e1) cudaEventDestroy(tc.e1);
    if (tc.h_table) cudaFreeHost(tc.h_table);
    tc = ThreadCtx();
    DPOSE_CUDA(cudaEventCreate(&tc.e0));
    DPOSE_CUDA(cudaEventCreate(&tc.e1));
    tc.device = pg->device;
  }
  constexpr size_t kZeroCopyTable = (size_t)1 << 16;  // tables up to 64 K cells land in mapped pinned memory
  if (np <= kZeroCopyTable && tc.cap < np) {
    if (tc.h_table) cudaFreeHost(tc.h_table);
    tc.h_table = nullptr;
    tc.cap = 0;
    const size_t cap = std::max<size_t>(np, 4096);
    DPOSE_CUDA(cudaHostAlloc(&tc.h_table, sizeof(float) * cap, cudaHostAllocMapped));
    tc.cap = cap;
  }
  try {
    pg->h_cost.resize(np);
  } catch (const std::bad_alloc &) {
    return fail(DPOSE_ENOMEM, "out of host memory");
  }
  unsigned char *scratch = nullptr;
  auto body = [&]() -> int {
    // persistent: [cost | pad | coef] in one stream-ordered allocation
    const size_t o_coef = (sizeof(float) * np + 255) & ~(size_t)255;
    unsigned char *keep = nullptr;
    DPOSE_CUDA(cudaMallocAsync(&keep, o_coef + sizeof(PatchCoef) * np, nullptr));
    pg->d_cost = reinterpret_cast<float *>(keep);
    pg->d_coef = reinterpret_cast<PatchCoef *>(keep + o_coef);
    DPOSE_CUDA(cudaMallocAsync(&scratch, ar.total, nullptr));
    float *cost_host = np <= kZeroCopyTable ? tc.h_table : nullptr;
    DPOSE_TRY(launch_precompute(pg, fp_xy, n, scratch, ar, pg->d_cost, pg->d_coef, cost_host, false, tc.e0, tc.e1));
    if (!cost_host)
      DPOSE_CUDA(cudaMemcpyAsync(pg->h_cost.data(), pg->d_cost, sizeof(float) * np, cudaMemcpyDeviceToHost, nullptr));
    DPOSE_CUDA(cudaFreeAsync(scratch, nullptr));
    scratch = nullptr;
    DPOSE_CUDA(cudaStreamSynchronize(nullptr));
    if (cost_host) std::memcpy(pg->h_cost.data(), cost_host, sizeof(float) * np);
    float ms = 0.f;
    DPOSE_CUDA(cudaEventElapsedTime(&ms, tc.e0, tc.e1));
    pg->precompute_us = ms * 1000.f;
    return DPOSE_OK;
  };
  const int rc = body();
  if (scratch) cudaFreeAsync(scratch, nullptr);
  return rc;
}

// the intermediate images: the kernel once more, on a scratch arena of its own, downloaded on first request
int precompute_fetch_stages(dpose_pg *pg) {
  if (!pg->h_edt_sq.empty()) return DPOSE_OK;
  const size_t np = (size_t)pg->rows * pg->cols;
  const int n = (int)(pg->footprint.size() / 2);
  DeviceGuard guard(pg->device);
  if (!guard.ok) return fail(DPOSE_ECUDA, "cudaSetDevice(%d) failed", pg->device);
  std::vector<int32_t> shifted(pg->footprint);  // dpose_core.cpp:128
  for (int i = 0; i < n; ++i) {
    shifted[2 * i] += pg->center[0];
    shifted[2 * i + 1] += pg->center[1];
  }
  const Arena ar(np, n);
  unsigned char *scratch = nullptr;
  std::vector<int32_t> d2;
  auto body = [&]() -> int {
    try {
      pg->h_pre_blur.resize(np);
      pg->h_outline.resize(np);
      pg->h_mask.resize(np);
      d2.resize(np);
    } catch (const std::bad_alloc &) {
      return fail(DPOSE_ENOMEM, "out of host memory");
    }
    const size_t o_cost = ar.total, o_coef = (o_cost + sizeof(float) * np + 255) & ~(size_t)255;
    DPOSE_CUDA(cudaMallocAsync(&scratch, o_coef + sizeof(PatchCoef) * np, nullptr));
    DPOSE_TRY(launch_precompute(pg, shifted.data(), n, scratch, ar, reinterpret_cast<float *>(scratch + o_cost),
                                reinterpret_cast<PatchCoef *>(scratch + o_coef), nullptr, true, nullptr, nullptr));
    DPOSE_CUDA(cudaMemcpyAsync(pg->h_pre_blur.data(), scratch + ar.o_pre, sizeof(float) * np, cudaMemcpyDeviceToHost, nullptr));
    DPOSE_CUDA(cudaMemcpyAsync(d2.data(), scratch + ar.o_d2, sizeof(int) * np, cudaMemcpyDeviceToHost, nullptr));
    DPOSE_CUDA(cudaMemcpyAsync(pg->h_outline.data(), scratch + ar.o_outline, np, cudaMemcpyDeviceToHost, nullptr));
    DPOSE_CUDA(cudaMemcpyAsync(pg->h_mask.data(), scratch + ar.o_mask, np, cudaMemcpyDeviceToHost, nullptr));
    DPOSE_CUDA(cudaStreamSynchronize(nullptr));
    return DPOSE_OK;
  };
  const int rc = body();
  if (scratch) cudaFreeAsync(scratch, nullptr);
  if (rc == DPOSE_OK) pg->h_edt_sq.swap(d2);  // last: h_edt_sq non-empty marks "fetched"
  return rc;
}

}  // namespace dpose
