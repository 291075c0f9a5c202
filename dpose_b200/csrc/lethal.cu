// lethal.cu — lethal-cell extraction (replaces lethal_cells_within, dpose_costmap.cpp:117-184).
//
// Host: clamp the ring into the map, trace its four edges (integer Bresenham, ROS
// raytraceLine-compatible) into per-row [xmin, xmax] intervals (to_rays, :106-115).
// Device: stream compaction over the uint8 grid, two sweeps like the reference
// (:143-155 count, :167-179 write) with an exclusive scan between them:
//   k_count : one warp per 2 KB row tile, four aligned 16-byte loads per lane (all in flight before the
//             first use), byte compare against 254 -> 16-bit lane masks -> popc -> per-tile count
//   k_scan  : two-level exclusive scan of the tile counts
//   k_write : same tiling; one warp prefix per 512-byte piece, each lane emits its (x, y)
// Output order is row-major (y ascending, x ascending) and deterministic.
// HBM traffic: 2 x box bytes read + 8 B per lethal cell written.
// Boxes up to 4 M cells (the plugin's search box is ~1e4) size the result by the box and run both sweeps back
// to back with one synchronisation; larger boxes read the total back before allocating.
#include <algorithm>
#include <climits>
#include <cstdlib>
#include <mutex>
#include <new>

#include "common.cuh"

namespace dpose {

// ---- host geometry ---------------------------------------------------------------------
// Integer Bresenham from begin to end (end exclusive), max(|dx|,|dy|) cells; error term
// starts at den/2 and the minor axis steps when it reaches den — the stepping rule of
// ROS' raytraceLine that dpose_costmap.cpp:57-104 follows.
template <typename Visit>
static size_t trace_line(int bx, int by, int ex, int ey, Visit &&visit) {
  const int ddx = ex - bx, ddy = ey - by;
  const int ax = std::abs(ddx), ay = std::abs(ddy);
  const bool y_major = ay > ax;  // ties: x is the major axis (first maximum)
  const int den = y_major ? ay : ax, add = y_major ? ax : ay;
  const int sx = (ddx > 0) - (ddx < 0), sy = (ddy > 0) - (ddy < 0);
  int err = den / 2, x = bx, y = by;
  for (int i = 0; i < den; ++i) {
    visit(x, y);
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
  return (size_t)den;
}

size_t host_raytrace(const int32_t begin[2], const int32_t end[2], int32_t *out_xy, size_t cap) {
  size_t k = 0;
  return trace_line(begin[0], begin[1], end[0], end[1], [&](int x, int y) {
    if (k < cap) {
      out_xy[2 * k] = x;
      out_xy[2 * k + 1] = y;
    }
    ++k;
  });
}

namespace {

constexpr int kSub = 4;                  // 512-byte pieces per warp tile: 4 independent 16-byte loads per lane in flight
constexpr int kTileBytes = 512 * kSub;   // one warp x 4 x 16 B
constexpr int kWarpsPerCta = 8;
constexpr int kScanChunk = 2048;  // tiles per scan CTA (512 threads x 4)

struct RowSpan {
  int xmin, xmax;
};

struct LethalArgs {
  const uint8_t *map;
  size_t pitch;
  const RowSpan *spans;  // nrows
  int y_lo, nrows;
  int x_base;   // 16-byte aligned
  int tiles_x;  // tiles per row
  long long n_tiles;
};

// 16 bytes -> bit i set iff byte i == 254 and x_first + i in [xmin, xmax]
__device__ __forceinline__ unsigned lethal_mask16(const uint4 v, int x_first, int xmin, int xmax) {
  const unsigned k = 0xFEFEFEFEu;
  unsigned m = 0;
  const unsigned w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const unsigned eq = __vcmpeq4(w[i], k) & 0x01010101u;  // bit 0/8/16/24
    m |= (((eq * 0x00204081u) >> 21) & 0xFu) << (4 * i);
  }
  // clip to the row interval
  const int lo = xmin - x_first, hi = xmax - x_first;  // valid bit range [lo, hi]
  const unsigned lo_mask = lo <= 0 ? 0xFFFFu : (0xFFFFu << lo) & 0xFFFFu;
  const unsigned hi_mask = hi >= 15 ? 0xFFFFu : (0xFFFFu >> (15 - hi));
  return m & lo_mask & hi_mask;
}

// the lane's kSub 16-byte pieces of a tile: piece c covers x = x_first + 512 c ... + 15 of row y; all loads are
// issued before the first use
__device__ __forceinline__ void tile_lane_masks(const LethalArgs &a, long long tile, int lane, unsigned m[kSub],
                                                int *x_first, int *y) {
  const int r = (int)(tile / a.tiles_x), j = (int)(tile - (long long)r * a.tiles_x);
  const RowSpan s = a.spans[r];
  const int x0 = a.x_base + j * kTileBytes + lane * 16;
  *x_first = x0;
  *y = a.y_lo + r;
  const uint8_t *row = a.map + (size_t)(a.y_lo + r) * a.pitch;
  uint4 v[kSub];
  bool live[kSub];
#pragma unroll
  for (int c = 0; c < kSub; ++c) {
    const int xc = x0 + 512 * c;
    // xc is 16-byte aligned; a piece that overlaps the span starts below size_x <= pitch (pitch % 16 == 0)
    live[c] = s.xmin <= s.xmax && xc <= s.xmax && xc + 15 >= s.xmin;
    v[c] = make_uint4(0u, 0u, 0u, 0u);
    if (live[c]) v[c] = __ldg(reinterpret_cast<const uint4 *>(row + xc));
  }
#pragma unroll
  for (int c = 0; c < kSub; ++c) m[c] = live[c] ? lethal_mask16(v[c], x0 + 512 * c, s.xmin, s.xmax) : 0u;
}

__global__ void __launch_bounds__(kWarpsPerCta * 32) k_count(LethalArgs a, unsigned *tile_count) {
  const int lane = threadIdx.x & 31;
  const long long tile = (long long)blockIdx.x * kWarpsPerCta + (threadIdx.x >> 5);
  if (tile >= a.n_tiles) return;
  int xf, y;
  unsigned m[kSub];
  tile_lane_masks(a, tile, lane, m, &xf, &y);
  unsigned c = 0;
#pragma unroll
  for (int k = 0; k < kSub; ++k) c += __popc(m[k]);
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) c += __shfl_xor_sync(0xffffffffu, c, off);
  if (lane == 0) tile_count[tile] = c;
}

// level 1: each CTA turns kScanChunk counts into exclusive prefixes + one chunk total
__global__ void __launch_bounds__(512) k_scan_chunks(const unsigned *tile_count, unsigned *tile_prefix,
                                                     unsigned long long *chunk_total, long long n_tiles) {
  __shared__ unsigned warp_sum[16];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const long long base = (long long)blockIdx.x * kScanChunk + tid * 4;
  unsigned v[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) v[i] = base + i < n_tiles ? tile_count[base + i] : 0u;
  const unsigned mine = v[0] + v[1] + v[2] + v[3];
  unsigned inc = mine;
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    const unsigned t = __shfl_up_sync(0xffffffffu, inc, off);
    if (lane >= off) inc += t;
  }
  if (lane == 31) warp_sum[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    unsigned w = lane < 16 ? warp_sum[lane] : 0u;
    unsigned winc = w;
#pragma unroll
    for (int off = 1; off < 16; off <<= 1) {
      const unsigned t = __shfl_up_sync(0xffffffffu, winc, off);
      if (lane >= off) winc += t;
    }
    if (lane < 16) warp_sum[lane] = winc - w;  // exclusive
    if (lane == 15) chunk_total[blockIdx.x] = winc;
  }
  __syncthreads();
  unsigned ex = warp_sum[warp] + inc - mine;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    if (base + i < n_tiles) tile_prefix[base + i] = ex;
    ex += v[i];
  }
}

// level 2: exclusive scan of the chunk totals (one CTA, sequential over 1024-wide slabs)
__global__ void __launch_bounds__(1024) k_scan_totals(unsigned long long *chunk_total, int n_chunks,
                                                      unsigned long long *grand_total) {
  __shared__ unsigned long long warp_sum[32];
  __shared__ unsigned long long carry;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) carry = 0ull;
  __syncthreads();
  for (int base = 0; base < n_chunks; base += 1024) {
    const int i = base + tid;
    const unsigned long long mine = i < n_chunks ? chunk_total[i] : 0ull;
    unsigned long long inc = mine;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const unsigned long long t = __shfl_up_sync(0xffffffffu, inc, off);
      if (lane >= off) inc += t;
    }
    if (lane == 31) warp_sum[warp] = inc;
    __syncthreads();
    if (warp == 0) {
      const unsigned long long w = warp_sum[lane];
      unsigned long long winc = w;
#pragma unroll
      for (int off = 1; off < 32; off <<= 1) {
        const unsigned long long t = __shfl_up_sync(0xffffffffu, winc, off);
        if (lane >= off) winc += t;
      }
      warp_sum[lane] = winc - w;
    }
    __syncthreads();
    const unsigned long long ex = carry + warp_sum[warp] + inc - mine;
    if (i < n_chunks) chunk_total[i] = ex;
    __syncthreads();
    if (tid == 1023) carry = ex + mine;
    __syncthreads();
  }
  if (tid == 0) *grand_total = carry;
}

__global__ void __launch_bounds__(kWarpsPerCta * 32)
    k_write(LethalArgs a, const unsigned *tile_prefix, const unsigned long long *chunk_offset, int2 *out) {
  const int lane = threadIdx.x & 31;
  const long long tile = (long long)blockIdx.x * kWarpsPerCta + (threadIdx.x >> 5);
  if (tile >= a.n_tiles) return;
  int xf, y;
  unsigned m[kSub];
  tile_lane_masks(a, tile, lane, m, &xf, &y);
  // x ascending inside the tile = piece-major, then lane, then bit: one warp prefix per piece
  unsigned long long base = chunk_offset[tile / kScanChunk] + tile_prefix[tile];
#pragma unroll
  for (int k = 0; k < kSub; ++k) {
    const unsigned c = __popc(m[k]);
    unsigned inc = c;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const unsigned t = __shfl_up_sync(0xffffffffu, inc, off);
      if (lane >= off) inc += t;
    }
    unsigned long long dst = base + (inc - c);
    unsigned mk = m[k];
    while (mk) {
      const int b = __ffs(mk) - 1;
      mk &= mk - 1;
      out[dst++] = make_int2(xf + 512 * k + b, y);
    }
    base += __shfl_sync(0xffffffffu, inc, 31);
  }
}

// Per-map workspace: scratch buffers, pinned staging and a private stream live in the map handle
// (grow-only), so a call costs no cudaMalloc / cudaFree besides the stream-ordered allocation of its
// result.  Calls on one map serialise on the workspace mutex (the reference holds the costmap mutex
// for the whole extraction anyway, dpose_costmap.cpp:138).
struct LethalWs {
  std::mutex mu;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  RowSpan *d_spans = nullptr, *h_spans = nullptr;  // h_spans pinned
  size_t rows_cap = 0;
  unsigned *d_count = nullptr, *d_prefix = nullptr;
  size_t tiles_cap = 0;
  unsigned long long *d_chunk = nullptr;  // n_chunks + 1 (last = grand total)
  size_t chunks_cap = 0;
  unsigned long long *h_total = nullptr;  // pinned
  float last_kernel_us = 0.f;

  int init(int device) {
    DPOSE_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    DPOSE_CUDA(cudaEventCreate(&ev0));
    DPOSE_CUDA(cudaEventCreate(&ev1));
    DPOSE_CUDA(cudaHostAlloc(&h_total, sizeof(*h_total), cudaHostAllocDefault));
    // keep freed result buffers in the device's default pool instead of returning them to the OS
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
      unsigned long long keep = ~0ull;
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
    return DPOSE_OK;
  }
  int reserve(size_t rows, size_t tiles, size_t chunks) {
    if (rows > rows_cap) {
      cudaFree(d_spans);
      cudaFreeHost(h_spans);
      d_spans = h_spans = nullptr;
      rows_cap = 0;
      const size_t cap = rows + rows / 2 + 64;
      DPOSE_CUDA(cudaMalloc(&d_spans, sizeof(RowSpan) * cap));
      DPOSE_CUDA(cudaHostAlloc(&h_spans, sizeof(RowSpan) * cap, cudaHostAllocDefault));
      rows_cap = cap;
    }
    if (tiles > tiles_cap) {
      cudaFree(d_count);
      cudaFree(d_prefix);
      d_count = d_prefix = nullptr;
      tiles_cap = 0;
      const size_t cap = tiles + tiles / 2 + 1024;
      DPOSE_CUDA(cudaMalloc(&d_count, sizeof(unsigned) * cap));
      DPOSE_CUDA(cudaMalloc(&d_prefix, sizeof(unsigned) * cap));
      tiles_cap = cap;
    }
    if (chunks + 1 > chunks_cap) {
      cudaFree(d_chunk);
      d_chunk = nullptr;
      chunks_cap = 0;
      const size_t cap = chunks + chunks / 2 + 16;
      DPOSE_CUDA(cudaMalloc(&d_chunk, sizeof(unsigned long long) * cap));
      chunks_cap = cap;
    }
    return DPOSE_OK;
  }
  void release() {
    cudaFree(d_spans);
    cudaFreeHost(h_spans);
    cudaFree(d_count);
    cudaFree(d_prefix);
    cudaFree(d_chunk);
    cudaFreeHost(h_total);
    if (ev0) cudaEventDestroy(ev0);
    if (ev1) cudaEventDestroy(ev1);
    if (stream) cudaStreamDestroy(stream);
  }
};

std::mutex g_ws_mu;  // guards creation of the per-map workspace
constexpr unsigned long long kSingleSyncCells = 1ull << 22;  // boxes up to 4 M cells: result sized by the box (32 MB)

}  // namespace

void map_release_lethal_ws(dpose_map *map) {
  LethalWs *ws = static_cast<LethalWs *>(map->lethal_ws);
  if (!ws) return;
  ws->release();
  delete ws;
  map->lethal_ws = nullptr;
}

float lethal_last_kernel_us(const dpose_map *map) {
  const LethalWs *ws = static_cast<const LethalWs *>(map->lethal_ws);
  return ws ? ws->last_kernel_us : 0.f;
}

int lethal_extract(const dpose_map *map, const int32_t box[10], int2 **d_cells, size_t *n) {
  *d_cells = nullptr;
  *n = 0;
  if (map->size_x == 0 || map->size_y == 0) return DPOSE_OK;  // dpose_costmap.cpp:124-125
  // clamp the ring into the map (:121, :128-129)
  int32_t ring[10];
  for (int i = 0; i < 5; ++i) {
    ring[2 * i] = std::min(std::max(box[2 * i], 0), (int)map->size_x - 1);
    ring[2 * i + 1] = std::min(std::max(box[2 * i + 1], 0), (int)map->size_y - 1);
  }
  int y_lo = INT_MAX, y_hi = INT_MIN;
  for (int i = 0; i < 5; ++i) {
    y_lo = std::min(y_lo, ring[2 * i + 1]);
    y_hi = std::max(y_hi, ring[2 * i + 1]);
  }
  const int nrows = y_hi - y_lo + 1;

  DeviceGuard guard(map->device);
  if (!guard.ok) return fail(DPOSE_ECUDA, "cudaSetDevice(%d) failed", map->device);
  LethalWs *ws;
  {
    std::lock_guard<std::mutex> lock(g_ws_mu);
    dpose_map *m = const_cast<dpose_map *>(map);
    if (!m->lethal_ws) {
      LethalWs *fresh = new (std::nothrow) LethalWs();
      if (!fresh) return fail(DPOSE_ENOMEM, "out of host memory");
      const int rc = fresh->init(map->device);
      if (rc != DPOSE_OK) {
        fresh->release();
        delete fresh;
        return rc;
      }
      m->lethal_ws = fresh;
    }
    ws = static_cast<LethalWs *>(m->lethal_ws);
  }
  std::lock_guard<std::mutex> call_lock(ws->mu);
  DPOSE_TRY(ws->reserve((size_t)nrows, 0, 0));
  RowSpan *spans = ws->h_spans;
  for (int r = 0; r < nrows; ++r) spans[r] = RowSpan{INT_MAX, INT_MIN};
  for (int e = 0; e < 4; ++e)  // to_rays (:106-115)
    trace_line(ring[2 * e], ring[2 * e + 1], ring[2 * e + 2], ring[2 * e + 3], [&](int x, int y) {
      RowSpan &s = spans[(size_t)(y - y_lo)];
      s.xmin = std::min(s.xmin, x);
      s.xmax = std::max(s.xmax, x);
    });
  int gx_min = INT_MAX, gx_max = INT_MIN;
  unsigned long long box_cells = 0;  // cells inside the row intervals: an upper bound of the result size
  for (int r = 0; r < nrows; ++r)
    if (spans[r].xmin <= spans[r].xmax) {
      gx_min = std::min(gx_min, spans[r].xmin);
      gx_max = std::max(gx_max, spans[r].xmax);
      box_cells += (unsigned long long)(spans[r].xmax - spans[r].xmin + 1);
    }
  if (gx_min > gx_max) return DPOSE_OK;  // degenerate ring (all four edges empty)

  LethalArgs a;
  a.map = map->d_data;
  a.pitch = map->pitch;
  a.y_lo = y_lo;
  a.nrows = nrows;
  a.x_base = gx_min & ~15;
  a.tiles_x = (gx_max - a.x_base) / kTileBytes + 1;
  a.n_tiles = (long long)nrows * a.tiles_x;
  const int n_chunks = (int)((a.n_tiles + kScanChunk - 1) / kScanChunk);
  DPOSE_TRY(ws->reserve((size_t)nrows, (size_t)a.n_tiles, (size_t)n_chunks));
  a.spans = ws->d_spans;

  cudaStream_t st = ws->stream;
  int2 *d_out = nullptr;
  auto body = [&]() -> int {
    DPOSE_CUDA(cudaMemcpyAsync(ws->d_spans, ws->h_spans, sizeof(RowSpan) * nrows, cudaMemcpyHostToDevice, st));
    const unsigned grid = (unsigned)((a.n_tiles + kWarpsPerCta - 1) / kWarpsPerCta);
    DPOSE_CUDA(cudaEventRecord(ws->ev0, st));
    k_count<<<grid, kWarpsPerCta * 32, 0, st>>>(a, ws->d_count);
    DPOSE_LAUNCHED();
    k_scan_chunks<<<n_chunks, 512, 0, st>>>(ws->d_count, ws->d_prefix, ws->d_chunk, a.n_tiles);
    DPOSE_LAUNCHED();
    k_scan_totals<<<1, 1024, 0, st>>>(ws->d_chunk, n_chunks, ws->d_chunk + n_chunks);
    DPOSE_LAUNCHED();
    DPOSE_CUDA(cudaMemcpyAsync(ws->h_total, ws->d_chunk + n_chunks, sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
    unsigned long long total;
    if (box_cells <= kSingleSyncCells) {
      // small boxes (the plugin's search box): size the result by the box and run both sweeps back to back —
      // one synchronisation per call instead of a host round trip between the sweeps
      DPOSE_CUDA(cudaMallocAsync(&d_out, sizeof(int2) * box_cells, st));
      k_write<<<grid, kWarpsPerCta * 32, 0, st>>>(a, ws->d_prefix, ws->d_chunk, d_out);
      DPOSE_LAUNCHED();
      DPOSE_CUDA(cudaEventRecord(ws->ev1, st));
      DPOSE_CUDA(cudaStreamSynchronize(st));
      total = *ws->h_total;
      if (total == 0) {  // :157-158
        cudaFreeAsync(d_out, st);
        d_out = nullptr;
        return DPOSE_OK;
      }
    } else {
      DPOSE_CUDA(cudaStreamSynchronize(st));
      total = *ws->h_total;
      if (total == 0) return DPOSE_OK;  // :157-158
      DPOSE_CUDA(cudaMallocAsync(&d_out, sizeof(int2) * total, st));
      k_write<<<grid, kWarpsPerCta * 32, 0, st>>>(a, ws->d_prefix, ws->d_chunk, d_out);
      DPOSE_LAUNCHED();
      DPOSE_CUDA(cudaEventRecord(ws->ev1, st));
      DPOSE_CUDA(cudaStreamSynchronize(st));
    }
    cudaEventElapsedTime(&ws->last_kernel_us, ws->ev0, ws->ev1);  // ms, includes the host round trip for the total
    ws->last_kernel_us *= 1000.f;
    *d_cells = d_out;
    *n = (size_t)total;
    d_out = nullptr;
    return DPOSE_OK;
  };
  const int rc = body();
  if (d_out) cudaFree(d_out);
  return rc;
}

}  // namespace dpose
