// lethal.cu — lethal-cell extraction (replaces lethal_cells_within, dpose_costmap.cpp:117-184).
//
// Host: clamp the ring into the map, trace its four edges (integer Bresenham, ROS
// raytraceLine-compatible) into per-row [xmin, xmax] intervals (to_rays, :106-115).
// Device: stream compaction.  The reference sweeps the uint8 grid twice (:143-155 count, :167-179 write).  Here the
// uint8 map is read ONCE — by k_pack_tiles (get_cost_batch_tile.cu: 16-byte coalesced loads, byte compare against 254),
// and only for the rows that changed since the lethal tile plane was last built (all rows of a volatile map, none of a
// clean one) — and the compaction itself runs over that 1-bit-per-cell plane (1/8 of the bytes, L2 resident):
//   k_count : one warp per (tile row, chunk of 32 tiles): one 256-bit load per lane = 1 KB of contiguous plane per warp,
//             each row word clipped to the row's interval (interior units skip the clip), popc, packed butterflies ->
//             the counts of the 8 (row, 1024-cell chunk) units
//   k_scan  : exclusive scan of the unit counts in row-major order, both levels in one launch (last-CTA pattern)
//   k_write : same load (L2 hot); row by row every lane expands its word into x offsets in shared memory, then the warp
//             writes the unit's (x, y) pairs as consecutive 8-byte stores (full sectors)
//   k_extract_small : boxes of up to 4096 units (the plugin's search box, a footprint's box) in one CTA, no look-back
// Output order is row-major (y ascending, x ascending) and deterministic.
// HBM traffic, 10000 x 10000 map at 10 % lethal: 100 MB read + 12.6 MB written by the pack (if the plane is stale),
// 2 x 12.6 MB of plane read (L2 hits right after the pack), 80 MB of cells written — §8(d)'s 180 MB plus the plane.
// The result buffer is sized by the box while that is small (<= 4 M cells), otherwise by what the previous extraction on
// the same map returned (+25 %): one synchronisation per call; a first call on a large box, or one whose estimate
// was too small, reads the total back before the write pass.
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

// One unit of work = one map row x one chunk of 32 tile columns (1024 cells): a warp, one plane word per lane.
// A CTA of 8 warps takes the 8 rows of one tile row of one chunk: its 32 tiles are 1 KB of contiguous plane.
constexpr int kChunkTiles = 32;
constexpr int kWarpsPerCta = 8;

struct RowSpan {
  int xmin, xmax;
};

struct LethalArgs {
  const uint32_t *tiles;  // lethal tile plane (common.cuh)
  uint32_t ntx;
  const RowSpan *spans;  // nrows, indexed by y - y_lo
  int y_lo, nrows;
  int in_x0, in_x1;  // columns every row's interval contains (max of xmin, min of xmax; empty if in_x0 > in_x1)
  int ty0;        // first tile row (y_lo >> 3)
  int tx0;        // first tile column of chunk 0
  int chunks_x;   // chunks per row
  int n_tile_rows;
  long long n_units;  // nrows * chunks_x, unit = (y - y_lo) * chunks_x + chunk
};

// One warp per (tile row, chunk of 32 tile columns): lane l fetches tile tx0 + 32 chunk + l with ONE 256-bit load — the
// warp's 32 tiles are 1 KB of contiguous plane — and holds the 8 row words of that tile.  Rows outside the box and
// cells outside a row's interval are masked out.  No shared-memory staging, no block barrier: every warp has its own
// kilobyte in flight (8 warps per CTA, 8 CTAs per SM: 64 KB of loads in flight per SM).
struct Tile8 {
  unsigned w[8];
};
__device__ __forceinline__ void warp_tile(const LethalArgs &a, int tr, int chunk, int lane, Tile8 *t, int *ty_out) {
  const int ty = a.ty0 + tr, tx = a.tx0 + chunk * kChunkTiles + lane;
  *ty_out = ty;
#pragma unroll
  for (int r = 0; r < 8; ++r) t->w[r] = 0u;
  if ((uint32_t)tx < a.ntx) {  // tiles past the plane's width read as zero
    const uint32_t *p = a.tiles + ((size_t)ty * a.ntx + (size_t)tx) * 8;
    asm volatile("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(t->w[0]), "=r"(t->w[1]), "=r"(t->w[2]), "=r"(t->w[3]), "=r"(t->w[4]), "=r"(t->w[5]), "=r"(t->w[6]),
                   "=r"(t->w[7])
                 : "l"(p));
  }
  // interior of the box (warp-uniform test): all 8 rows belong to it and every row's interval covers the chunk's 1024
  // columns — nothing to clip.  Nearly every unit of a large box takes this path.
  const int xc = (a.tx0 + chunk * kChunkTiles) * 32;
  if (ty * 8 >= a.y_lo && ty * 8 + 7 < a.y_lo + a.nrows && xc >= a.in_x0 && xc + kChunkTiles * 32 - 1 <= a.in_x1) return;
  const int x0 = tx * 32;
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    const int yr = ty * 8 + r - a.y_lo;
    unsigned m = 0u;
    if (yr >= 0 && yr < a.nrows) {
      const RowSpan sp = a.spans[yr];
      const int lo = sp.xmin - x0, hi = sp.xmax - x0;  // valid bit range [lo, hi]; an empty interval has xmin > xmax
      if (!(hi < 0 || lo > 31 || sp.xmin > sp.xmax)) {
        m = t->w[r];
        if (lo > 0) m &= 0xffffffffu << lo;
        if (hi < 31) m &= 0xffffffffu >> (31 - hi);
      }
    }
    t->w[r] = m;
  }
}

// The compaction over the tile plane: count, scan, write.  A warp owns one (tile row, chunk of 1024 cells) — 8 units, one
// per map row — and never synchronises with another warp:
//   k_count : 256-bit loads (one tile per lane = 1 KB of contiguous plane per warp), every row word clipped to the row's
//             interval (skipped for interior units), popc, packed butterflies -> unit_count[(row, chunk)]
//   k_scan  : exclusive scan of the unit counts in row-major order (the output order): 2048 units per CTA; the CTA that
//             finishes last (ticket) scans the CTA totals, so one launch does both levels
//   k_write : the same load (L2 hot); the 8 units' offsets are fetched up front by 8 lanes; then, row by row, every lane
//             expands its word into x offsets in shared memory and the warp writes the row's (x, y) pairs as consecutive
//             8-byte stores (full sectors)
// (Two other shapes were built and measured on the full 10000 x 10000 map, clean plane: a single-pass chained scan with
// decoupled look-back — 77 us, the look-back over ~1000 resident CTAs costs more than a second L2-resident read of
// 12.6 MB — and a CTA-per-tile-row totals pass + write pass with the offset summed on the fly — 61 us, two block
// barriers and a block scan per 8 KB of input.)
// `cap`: cells the result buffer holds; a unit that would run past it writes nothing (the host sees total > cap and
// repeats the write pass with an exact allocation).
constexpr int kScanUnits = 2048;  // units per scan CTA (512 threads x 4)

__global__ void __launch_bounds__(kWarpsPerCta * 32) k_count(LethalArgs a, unsigned *unit_count) {
  DPOSE_PDL_WAIT();     // behind k_pack_tiles (or the copy of the row intervals)
  DPOSE_PDL_TRIGGER();  // k_scan may queue up
  const int lane = threadIdx.x & 31;
  const long long wid = (long long)blockIdx.x * kWarpsPerCta + (threadIdx.x >> 5);
  if (wid >= (long long)a.n_tile_rows * a.chunks_x) return;
  const int tr = (int)(wid / a.chunks_x), chunk = (int)(wid - (long long)tr * a.chunks_x);
  Tile8 t;
  int ty;
  warp_tile(a, tr, chunk, lane, &t, &ty);
  unsigned c[4];  // eight counts per lane (each <= 32, warp sums <= 1024): two per register, four butterflies
#pragma unroll
  for (int q = 0; q < 4; ++q) c[q] = __popc(t.w[2 * q]) | (__popc(t.w[2 * q + 1]) << 16);
#pragma unroll
  for (int off = 16; off > 0; off >>= 1)
#pragma unroll
    for (int q = 0; q < 4; ++q) c[q] += __shfl_xor_sync(0xffffffffu, c[q], off);
  if (lane < 8) {
    const int yr = ty * 8 + lane - a.y_lo;
    if (yr >= 0 && yr < a.nrows) unit_count[(long long)yr * a.chunks_x + chunk] = (c[lane >> 1] >> ((lane & 1) * 16)) & 0xffffu;
  }
}

// both scan levels in one launch: every CTA turns kScanUnits counts into exclusive prefixes and a CTA total; the CTA
// whose ticket is the last one then scans the totals in place (exclusive) and publishes the grand total
__global__ void __launch_bounds__(512) k_scan(const unsigned *unit_count, unsigned *unit_prefix, unsigned long long *cta_total,
                                              unsigned *ticket, long long n_units, unsigned long long *grand_total) {
  __shared__ unsigned warp_sum[16];
  __shared__ bool s_last;
  DPOSE_PDL_WAIT();
  DPOSE_PDL_TRIGGER();
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const long long base = (long long)blockIdx.x * kScanUnits + tid * 4;
  unsigned v[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) v[i] = base + i < n_units ? unit_count[base + i] : 0u;
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
    if (lane == 15) {
      cta_total[blockIdx.x] = winc;
      __threadfence();
      s_last = atomicAdd(ticket, 1u) == gridDim.x - 1;
    }
  }
  __syncthreads();
  unsigned ex = warp_sum[warp] + inc - mine;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    if (base + i < n_units) unit_prefix[base + i] = ex;
    ex += v[i];
  }
  if (!s_last) return;
  // last CTA: exclusive scan of the gridDim.x totals (a 10000 x 10000 map has 49 of them), 512 at a time
  __threadfence();
  __shared__ unsigned long long s_w[16];
  __shared__ unsigned long long s_carry;
  if (tid == 0) s_carry = 0ull;
  __syncthreads();
  volatile unsigned long long *tot = cta_total;
  for (int b0 = 0; b0 < (int)gridDim.x; b0 += 512) {
    const int i = b0 + tid;
    const unsigned long long m = i < (int)gridDim.x ? tot[i] : 0ull;
    unsigned long long s = m;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const unsigned long long t = __shfl_up_sync(0xffffffffu, s, off);
      if (lane >= off) s += t;
    }
    if (lane == 31) s_w[warp] = s;
    __syncthreads();
    if (warp == 0) {
      const unsigned long long w = lane < 16 ? s_w[lane] : 0ull;
      unsigned long long ws = w;
#pragma unroll
      for (int off = 1; off < 16; off <<= 1) {
        const unsigned long long t = __shfl_up_sync(0xffffffffu, ws, off);
        if (lane >= off) ws += t;
      }
      if (lane < 16) s_w[lane] = ws - w;
    }
    __syncthreads();
    const unsigned long long exc = s_carry + s_w[warp] + s - m;
    if (i < (int)gridDim.x) tot[i] = exc;
    __syncthreads();
    if (tid == 511) s_carry = exc + m;
    __syncthreads();
  }
  if (tid == 0) {
    *grand_total = s_carry;
    *ticket = 0u;  // ready for the next call
  }
}

__global__ void __launch_bounds__(kWarpsPerCta * 32)
    k_write(LethalArgs a, const unsigned *unit_prefix, const unsigned long long *cta_offset, int2 *out, unsigned long long cap) {
  __shared__ unsigned short s_stage[kWarpsPerCta][kChunkTiles * 32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const long long wid = (long long)blockIdx.x * kWarpsPerCta + w;
  if (wid >= (long long)a.n_tile_rows * a.chunks_x) return;
  const int tr = (int)(wid / a.chunks_x), chunk = (int)(wid - (long long)tr * a.chunks_x);
  Tile8 t;
  int ty;
  warp_tile(a, tr, chunk, lane, &t, &ty);  // the plane and the intervals are older than k_scan: fetched before the wait
  DPOSE_PDL_WAIT();
  // offsets of the 8 units, fetched by lanes 0..7 before the row loop (two dependent loads once, not once per row)
  unsigned long long my_base = 0;
  {
    const int yr = ty * 8 + (lane & 7) - a.y_lo;
    if (lane < 8 && yr >= 0 && yr < a.nrows) {
      const long long unit = (long long)yr * a.chunks_x + chunk;
      my_base = cta_offset[unit / kScanUnits] + unit_prefix[unit];
    }
  }
  const int xc = (a.tx0 + chunk * kChunkTiles) * 32;
  const unsigned xl = (unsigned)(lane * 32);
  unsigned short *stage = s_stage[w];
  const uint32_t stage_u32 = (uint32_t)__cvta_generic_to_shared(stage);
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    unsigned m = t.w[r];
    const unsigned c = __popc(m);
    unsigned inc = c;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const unsigned v = __shfl_up_sync(0xffffffffu, inc, off);
      if (lane >= off) inc += v;
    }
    const unsigned total = __shfl_sync(0xffffffffu, inc, 31);
    const unsigned long long base = __shfl_sync(0xffffffffu, my_base, r);
    if (total == 0) continue;  // warp-uniform (also for rows outside the box)
    uint32_t sp = stage_u32 + 2u * (inc - c);
    m = __brev(m);  // lowest x first with one FLO per bit
    while (m) {     // ~3 set bits per lane at 10 % density
      const unsigned b = (unsigned)__clz(m);
      m ^= 0x80000000u >> b;
      asm volatile("st.shared.u16 [%0], %1;" ::"r"(sp), "h"((unsigned short)(xl + b)) : "memory");
      sp += 2u;
    }
    __syncwarp();
    if (base + total <= cap) {
      const int y = ty * 8 + r;
      int2 *o = out + base + lane;
      const unsigned short *q = stage + lane;
      for (unsigned k = lane; k < total; k += 32, o += 32, q += 32) *o = make_int2(xc + (int)*q, y);
    }
    __syncwarp();  // the staging row is reused by the next row
  }
}

// Small boxes (the plugin's search box, a footprint's rotated box): count, scan and emit in ONE launch of one CTA —
// the multi-kernel pipeline above costs four launches, which is all a 100 x 100 box is.
constexpr int kSmallUnits = 4096;
__global__ void __launch_bounds__(1024) k_extract_small(LethalArgs a, int2 *out, unsigned long long cap,
                                                        unsigned long long *grand_total) {
  __shared__ unsigned s_count[kSmallUnits];
  __shared__ unsigned s_warp[32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  auto word_of = [&](int u, int *y, int *x0) -> unsigned {
    const int yr = u / a.chunks_x, chunk = u - yr * a.chunks_x;
    *y = a.y_lo + yr;
    const int tx = a.tx0 + chunk * kChunkTiles + lane;
    *x0 = tx * 32;
    unsigned m = 0u;
    if ((uint32_t)tx < a.ntx) m = __ldg(a.tiles + ((size_t)(*y >> 3) * a.ntx + (size_t)tx) * 8 + (*y & 7));
    const RowSpan sp = a.spans[yr];
    const int lo = sp.xmin - *x0, hi = sp.xmax - *x0;
    if (hi < 0 || lo > 31 || sp.xmin > sp.xmax) return 0u;
    if (lo > 0) m &= 0xffffffffu << lo;
    if (hi < 31) m &= 0xffffffffu >> (31 - hi);
    return m;
  };
  const int n_units = (int)a.n_units;
  for (int u = warp; u < n_units; u += 32) {
    int y, x0;
    unsigned c = __popc(word_of(u, &y, &x0));
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) c += __shfl_xor_sync(0xffffffffu, c, off);
    if (lane == 0) s_count[u] = c;
  }
  __syncthreads();
  {  // exclusive scan of s_count[0 .. n_units) in place: 4 per thread, warp scan, scan of the warp sums
    unsigned v[4];
    unsigned mine = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      v[i] = tid * 4 + i < n_units ? s_count[tid * 4 + i] : 0u;
      mine += v[i];
    }
    unsigned inc = mine;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const unsigned t = __shfl_up_sync(0xffffffffu, inc, off);
      if (lane >= off) inc += t;
    }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    if (warp == 0) {
      const unsigned wv = s_warp[lane];
      unsigned winc = wv;
#pragma unroll
      for (int off = 1; off < 32; off <<= 1) {
        const unsigned t = __shfl_up_sync(0xffffffffu, winc, off);
        if (lane >= off) winc += t;
      }
      s_warp[lane] = winc - wv;
      if (lane == 31) *grand_total = winc;
    }
    __syncthreads();
    unsigned ex = s_warp[warp] + inc - mine;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (tid * 4 + i < n_units) s_count[tid * 4 + i] = ex;
      ex += v[i];
    }
  }
  __syncthreads();
  for (int u = warp; u < n_units; u += 32) {
    int y, x0;
    unsigned m = word_of(u, &y, &x0);
    const unsigned c = __popc(m);
    unsigned inc = c;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const unsigned t = __shfl_up_sync(0xffffffffu, inc, off);
      if (lane >= off) inc += t;
    }
    unsigned long long dst = (unsigned long long)s_count[u] + (inc - c);
    while (m) {  // ~3 cells per lane at 10 % density; a staged emit does not pay at this size
      const int b = __ffs(m) - 1;
      m &= m - 1;
      if (dst < cap) out[dst] = make_int2(x0 + b, y);
      ++dst;
    }
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
  unsigned *d_count = nullptr, *d_prefix = nullptr;  // per (row, chunk) unit
  size_t units_cap = 0;
  unsigned long long *d_cta = nullptr;  // per scan CTA: total, then exclusive offset; [n] = grand total, [n + 1] = ticket
  size_t cta_cap = 0;
  unsigned long long *h_total = nullptr;  // pinned
  float last_kernel_us = 0.f;
  unsigned long long hint = 0;  // cells the last extraction on this map returned

  int init(int device) {
    DPOSE_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    DPOSE_CUDA(cudaEventCreate(&ev0));
    DPOSE_CUDA(cudaEventCreate(&ev1));
    DPOSE_CUDA(cudaHostAlloc(&h_total, sizeof(*h_total), cudaHostAllocDefault));
    (void)device;  // the default pool's release threshold is raised once per device in api.cu (check_device)
    return DPOSE_OK;
  }
  int reserve(size_t rows, size_t units) {
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
    if (units > units_cap) {
      cudaFree(d_count);
      cudaFree(d_prefix);
      d_count = d_prefix = nullptr;
      units_cap = 0;
      const size_t cap = units + units / 2 + 1024;
      DPOSE_CUDA(cudaMalloc(&d_count, sizeof(unsigned) * cap));
      DPOSE_CUDA(cudaMalloc(&d_prefix, sizeof(unsigned) * cap));
      units_cap = cap;
    }
    const size_t ctas = (units + kScanUnits - 1) / kScanUnits;
    if (ctas + 2 > cta_cap) {
      cudaFree(d_cta);
      d_cta = nullptr;
      cta_cap = 0;
      const size_t cap = ctas + ctas / 2 + 16;
      DPOSE_CUDA(cudaMalloc(&d_cta, sizeof(unsigned long long) * cap));
      DPOSE_CUDA(cudaMemset(d_cta, 0, sizeof(unsigned long long) * cap));  // the ticket starts at 0
      DPOSE_CUDA(cudaStreamSynchronize(nullptr));  // the memset is asynchronous; its users run on a non-blocking stream
      cta_cap = cap;
    }
    return DPOSE_OK;
  }
  void release() {
    cudaFree(d_spans);
    cudaFreeHost(h_spans);
    cudaFree(d_count);
    cudaFree(d_prefix);
    cudaFree(d_cta);
    cudaFreeHost(h_total);
    if (ev0) cudaEventDestroy(ev0);
    if (ev1) cudaEventDestroy(ev1);
    if (stream) cudaStreamDestroy(stream);
  }
};

std::mutex g_ws_mu;  // guards creation of the per-map workspace
constexpr unsigned long long kBoxSizedCells = 1ull << 22;  // boxes up to 4 M cells: result sized by the box (32 MB)

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
  if (map->size_x == 0 || map->size_y == 0 || !map->d_tiles) return DPOSE_OK;  // dpose_costmap.cpp:124-125
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
  DPOSE_TRY(ws->reserve((size_t)nrows, 0));
  RowSpan *spans = ws->h_spans;
  for (int r = 0; r < nrows; ++r) spans[r] = RowSpan{INT_MAX, INT_MIN};
  for (int e = 0; e < 4; ++e)  // to_rays (:106-115)
    trace_line(ring[2 * e], ring[2 * e + 1], ring[2 * e + 2], ring[2 * e + 3], [&](int x, int y) {
      RowSpan &s = spans[(size_t)(y - y_lo)];
      s.xmin = std::min(s.xmin, x);
      s.xmax = std::max(s.xmax, x);
    });
  int gx_min = INT_MAX, gx_max = INT_MIN;
  int in_x0 = INT_MIN, in_x1 = INT_MAX;  // the columns every row's interval contains
  unsigned long long box_cells = 0;  // cells inside the row intervals: an upper bound of the result size
  for (int r = 0; r < nrows; ++r) {
    if (spans[r].xmin <= spans[r].xmax) {
      gx_min = std::min(gx_min, spans[r].xmin);
      gx_max = std::max(gx_max, spans[r].xmax);
      box_cells += (unsigned long long)(spans[r].xmax - spans[r].xmin + 1);
    }
    in_x0 = std::max(in_x0, spans[r].xmin);
    in_x1 = std::min(in_x1, spans[r].xmax);
  }
  if (gx_min > gx_max) return DPOSE_OK;  // degenerate ring (all four edges empty)

  LethalArgs a;
  a.tiles = map->d_tiles;
  a.ntx = map->tiles_ntx;
  a.y_lo = y_lo;
  a.nrows = nrows;
  a.in_x0 = in_x0;
  a.in_x1 = in_x1;
  a.ty0 = y_lo >> 3;
  a.n_tile_rows = (y_hi >> 3) - a.ty0 + 1;
  a.tx0 = gx_min >> 5;
  a.chunks_x = ((gx_max >> 5) - a.tx0) / kChunkTiles + 1;
  a.n_units = (long long)nrows * a.chunks_x;
  if (a.chunks_x > 4096) return fail(DPOSE_EINVAL, "box wider than 4 M cells");  // 32 B of shared memory per chunk
  DPOSE_TRY(ws->reserve((size_t)nrows, (size_t)a.n_units));
  a.spans = ws->d_spans;
  // the ticket lives at a fixed slot (the end of the array) so that it survives calls with different CTA counts
  const int n_scan = (int)((a.n_units + kScanUnits - 1) / kScanUnits);
  unsigned long long *d_total = ws->d_cta + (ws->cta_cap - 2);
  unsigned *d_ticket = reinterpret_cast<unsigned *>(ws->d_cta + (ws->cta_cap - 1));

  cudaStream_t st = ws->stream;
  int2 *d_out = nullptr;
  auto body = [&]() -> int {
    DPOSE_CUDA(cudaMemcpyAsync(ws->d_spans, ws->h_spans, sizeof(RowSpan) * nrows, cudaMemcpyHostToDevice, st));
    DPOSE_CUDA(cudaEventRecord(ws->ev0, st));
    // the ONE read of the uint8 map: the lethal tile plane, rebuilt for the rows that changed (of a volatile map: the
    // rows of this box — a search box on a 10000^2 map does not pay for the whole map's pass); a clean plane costs nothing
    // here.  The compaction below reads the 1-bit plane (1/8 of the map's bytes).
    DPOSE_TRY(tileplane_ensure(map, st, false, (uint32_t)y_lo, (uint32_t)y_hi + 1u));
    // capacity of the result buffer: the box itself while that is small, else what the last call on this map needed
    // (+25 %); a first call on a large box reads the total back before the write pass (one more synchronisation)
    unsigned long long cap = box_cells <= kBoxSizedCells ? box_cells : (ws->hint ? ws->hint + ws->hint / 4 + 4096 : 0);
    cap = std::min(cap, box_cells);
    unsigned long long total = 0;
    if (a.n_units <= kSmallUnits) {  // one CTA: count, scan and emit; <= 4 M cells, sized by the box
      cap = box_cells;
      DPOSE_CUDA(cudaMallocAsync(&d_out, sizeof(int2) * cap, st));
      k_extract_small<<<1, 1024, 0, st>>>(a, d_out, cap, d_total);
      DPOSE_LAUNCHED();
      DPOSE_CUDA(cudaMemcpyAsync(ws->h_total, d_total, sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
      DPOSE_CUDA(cudaEventRecord(ws->ev1, st));
      DPOSE_CUDA(cudaStreamSynchronize(st));
      total = *ws->h_total;
    } else {
      const unsigned grid = (unsigned)(((size_t)a.n_tile_rows * a.chunks_x + kWarpsPerCta - 1) / kWarpsPerCta);
      DPOSE_CUDA(launch_pdl(k_count, dim3(grid), dim3(kWarpsPerCta * 32), 0, st, a, ws->d_count));
      DPOSE_LAUNCHED();
      DPOSE_CUDA(launch_pdl(k_scan, dim3((unsigned)n_scan), dim3(512), 0, st, (const unsigned *)ws->d_count, ws->d_prefix,
                            ws->d_cta, d_ticket, a.n_units, d_total));
      DPOSE_LAUNCHED();
      for (int attempt = 0; attempt < 2; ++attempt) {
        if (!cap) {  // no estimate (or it was too small): the exact total first
          DPOSE_CUDA(cudaMemcpyAsync(ws->h_total, d_total, sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
          DPOSE_CUDA(cudaStreamSynchronize(st));
          total = *ws->h_total;
          if (total == 0) break;
          cap = total;
        }
        DPOSE_CUDA(cudaMallocAsync(&d_out, sizeof(int2) * cap, st));
        DPOSE_CUDA(launch_pdl(k_write, dim3(grid), dim3(kWarpsPerCta * 32), 0, st, a, (const unsigned *)ws->d_prefix,
                              (const unsigned long long *)ws->d_cta, d_out, cap));
        DPOSE_LAUNCHED();
        DPOSE_CUDA(cudaEventRecord(ws->ev1, st));
        DPOSE_CUDA(cudaMemcpyAsync(ws->h_total, d_total, sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
        DPOSE_CUDA(cudaStreamSynchronize(st));
        total = *ws->h_total;
        if (total <= cap) break;
        cudaFreeAsync(d_out, st);
        d_out = nullptr;
        cap = 0;
      }
    }
    ws->hint = total;
    if (total == 0) {  // :157-158
      if (d_out) cudaFreeAsync(d_out, st);
      d_out = nullptr;
      return DPOSE_OK;
    }
    cudaEventElapsedTime(&ws->last_kernel_us, ws->ev0, ws->ev1);  // ms: plane rebuild (if any) + the compaction pass(es)
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
