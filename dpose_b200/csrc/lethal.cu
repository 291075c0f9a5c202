// lethal.cu — lethal-cell extraction (replaces lethal_cells_within, dpose_costmap.cpp:117-184).
//
// Host: clamp the ring into the map, trace its four edges (integer Bresenham, ROS
// raytraceLine-compatible) into per-row [xmin, xmax] intervals (to_rays, :106-115).
// Device: stream compaction.  The reference sweeps the uint8 grid twice (:143-155 count, :167-179 write).  Here the
// uint8 map is read ONCE — by k_pack_tiles (get_cost_batch_tile.cu: 16-byte coalesced loads, byte compare against 254),
// and only for the rows that changed since the lethal tile plane was last built (all rows of a volatile map, none of a
// clean one) — and the compaction itself is ONE pass over that 1-bit-per-cell plane (1/8 of the bytes, L2 resident):
//   k_extract<false> : a CTA owns one tile row (8 map rows).  Its warps fetch the row's tiles with 256-bit loads (one
//               tile per lane = 1 KB of contiguous plane per warp), clip every row word to the row's interval and count;
//               the tile row's total goes to totals[], to its group of 64 tile rows and to the grand total (2 atomics)
//   k_extract<true>  : the counts again (L2 hot), scanned inside the CTA in row-major order; the CTA's global offset is
//               summed on the fly from the group totals and the tile-row totals before it (two short reads: no scan
//               kernel, no look-back chain); then, row by row, every lane expands its word into x offsets in shared
//               memory and the warp writes the row's (x, y) pairs as consecutive 8-byte stores.
//               (A single-pass chained scan with decoupled look-back was built and measured first: 77 us against 58 us
//               for count / scan / write — with 1 KB of input per warp the look-back over ~1000 resident CTAs costs more
//               than the second, L2-resident read of the plane.)
//   k_extract_small : boxes of up to 4096 units (the plugin's search box, a footprint's box) in one CTA, no look-back
// Output order is row-major (y ascending, x ascending) and deterministic.
// HBM traffic, 10000 x 10000 map at 10 % lethal: 100 MB read + 12.6 MB written by the pack (if the plane is stale),
// 12.6 MB of plane read (L2 hits right after the pack), 80 MB of cells written — §8(d)'s 180 MB plus the plane.
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

// The compaction over the tile plane, two launches of the same kernel body.  A CTA owns one tile row of the box (8 map
// rows x chunks_x chunks of 1024 cells); its warps fetch the row's tiles with 256-bit loads (one tile per lane = 1 KB
// of contiguous plane per warp), clip every row word to the row's interval and count the (row, chunk) units.
//   kWrite = false : the tile row's total goes to totals[tr], to its group of 64 tile rows (one atomic) and to the grand
//                    total (one atomic): 2 atomics per CTA instead of a scan kernel
//   kWrite = true  : the counts again (the plane is L2 / L1 hot), an in-CTA scan of the units in row-major order, the
//                    CTA's global offset summed on the fly from the group totals before its group and the tile-row
//                    totals before it inside the group (two short coalesced reads — no look-back chain, no scan
//                    kernel), then the emit: row by row every lane expands its word into x offsets in shared memory and
//                    the warp writes the row's (x, y) pairs as consecutive 8-byte stores (full sectors).
// `cap`: cells the result buffer holds; a unit that would run past it writes nothing (the host sees total > cap and
// repeats the write pass with an exact allocation).
constexpr int kGroup = 64;  // tile rows per group total

struct ExtractArgs {
  LethalArgs box;
  unsigned *totals;              // one per tile row of the box
  unsigned long long *groups;    // one per kGroup tile rows; [n_groups] = grand total
  int n_groups;
  int2 *out;
  unsigned long long cap;
};

template <bool kWrite>
__global__ void __launch_bounds__(kWarpsPerCta * 32) k_extract(ExtractArgs e) {
  // dynamic: [8][chunks_x] unit counts (then their exclusive prefixes, row-major) | per warp 1024 staged x offsets
  extern __shared__ unsigned s_cnt[];
  __shared__ unsigned s_part[32];
  __shared__ unsigned long long s_red[kWarpsPerCta];
  const LethalArgs &a = e.box;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int n_warps = blockDim.x >> 5, n_thr = blockDim.x;  // 1 .. kWarpsPerCta warps, chosen by the host
  const int cx = a.chunks_x, n_units = 8 * cx;
  const int tr = blockIdx.x;
  // ---- counts of this tile row's units -----------------------------------------------------------------------------
  for (int chunk = w; chunk < cx; chunk += n_warps) {
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
    if (lane < 8) s_cnt[lane * cx + chunk] = (c[lane >> 1] >> ((lane & 1) * 16)) & 0xffffu;
  }
  __syncthreads();
  // ---- exclusive scan of the n_units counts (row-major = output order); total of the tile row -----------------------------
  {
    const int per = (n_units + n_thr - 1) / n_thr;
    const int lo = min(tid * per, n_units), hi = min(lo + per, n_units);
    unsigned mine = 0;
    for (int k = lo; k < hi; ++k) mine += s_cnt[k];
    unsigned inc = mine;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const unsigned v = __shfl_up_sync(0xffffffffu, inc, off);
      if (lane >= off) inc += v;
    }
    if (lane == 31) s_part[w] = inc;
    __syncthreads();
    if (w == 0) {
      const unsigned pv = lane < n_warps ? s_part[lane] : 0u;
      unsigned pinc = pv;
#pragma unroll
      for (int off = 1; off < 32; off <<= 1) {
        const unsigned v = __shfl_up_sync(0xffffffffu, pinc, off);
        if (lane >= off) pinc += v;
      }
      if (lane < n_warps) s_part[lane] = pinc - pv;
      if (lane == 31) s_part[31] = pinc;  // total of the tile row
    }
    __syncthreads();
    if (kWrite) {
      unsigned ex = s_part[w] + inc - mine;
      for (int k = lo; k < hi; ++k) {
        const unsigned v = s_cnt[k];
        s_cnt[k] = ex;
        ex += v;
      }
    }
  }
  if (!kWrite) {
    if (tid == 0) {
      const unsigned total = s_part[31];
      e.totals[tr] = total;
      if (total) {
        atomicAdd(&e.groups[tr / kGroup], (unsigned long long)total);
        atomicAdd(&e.groups[e.n_groups], (unsigned long long)total);
      }
    }
    return;
  }
  // ---- global offset of this tile row: group totals before its group + tile-row totals before it inside the group --------
  {
    const int g = tr / kGroup;
    unsigned long long v = 0;
    for (int k = tid; k < g; k += n_thr) v += e.groups[k];
    for (int k2 = g * kGroup + tid; k2 < tr; k2 += n_thr) v += e.totals[k2];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    if (lane == 0) s_red[w] = v;
  }
  __syncthreads();
  unsigned long long base_tr = 0;
  for (int k = 0; k < n_warps; ++k) base_tr += s_red[k];
  // ---- emit, row by row: x offsets staged in shared memory, then consecutive stores ------------------------------------
  unsigned short *stage = reinterpret_cast<unsigned short *>(s_cnt + n_units) + w * (kChunkTiles * 32);
  const uint32_t stage_u32 = (uint32_t)__cvta_generic_to_shared(stage);
  for (int chunk = w; chunk < cx; chunk += n_warps) {
    Tile8 t;
    int ty;
    warp_tile(a, tr, chunk, lane, &t, &ty);  // the same kilobyte again: an L1 / L2 hit
    const int xc = (a.tx0 + chunk * kChunkTiles) * 32;
    const unsigned xl = (unsigned)(lane * 32);
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
      const unsigned long long base = base_tr + s_cnt[r * cx + chunk];
      if (base + total <= e.cap) {
        const int y = ty * 8 + r;
        int2 *o = e.out + base + lane;
        const unsigned short *q = stage + lane;
        for (unsigned k = lane; k < total; k += 32, o += 32, q += 32) *o = make_int2(xc + (int)*q, y);
      }
      __syncwarp();  // the staging row is reused by the next row
    }
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
  unsigned *d_totals = nullptr;            // one per tile row of the box
  unsigned long long *d_groups = nullptr;  // one per 64 tile rows, + the grand total
  size_t totals_cap = 0;
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
  int reserve(size_t rows, size_t tile_rows) {
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
    if (tile_rows > totals_cap) {
      cudaFree(d_totals);
      cudaFree(d_groups);
      d_totals = nullptr, d_groups = nullptr;
      totals_cap = 0;
      const size_t cap = tile_rows + tile_rows / 2 + 64;
      DPOSE_CUDA(cudaMalloc(&d_totals, sizeof(unsigned) * cap));
      DPOSE_CUDA(cudaMalloc(&d_groups, sizeof(unsigned long long) * (cap / kGroup + 2)));
      totals_cap = cap;
    }
    return DPOSE_OK;
  }
  void release() {
    cudaFree(d_spans);
    cudaFreeHost(h_spans);
    cudaFree(d_totals);
    cudaFree(d_groups);
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
  DPOSE_TRY(ws->reserve((size_t)nrows, (size_t)a.n_tile_rows));
  a.spans = ws->d_spans;
  const int n_groups = (a.n_tile_rows + kGroup - 1) / kGroup;
  unsigned long long *d_total = ws->d_groups + n_groups;

  cudaStream_t st = ws->stream;
  int2 *d_out = nullptr;
  auto body = [&]() -> int {
    DPOSE_CUDA(cudaMemcpyAsync(ws->d_spans, ws->h_spans, sizeof(RowSpan) * nrows, cudaMemcpyHostToDevice, st));
    DPOSE_CUDA(cudaEventRecord(ws->ev0, st));
    // the ONE read of the uint8 map: the lethal tile plane, rebuilt for the rows that changed (every row of a volatile
    // map); a clean plane costs nothing here.  The compaction below reads the 1-bit plane (1/8 of the map's bytes).
    DPOSE_TRY(tileplane_ensure(map, st, false));
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
      // warps per CTA: every warp one or two of the tile row's chunks, and CTAs small enough that the whole grid is
      // resident at once when it can be (1250 tile rows of a 10000-row map: 5 warps -> 12 CTAs per SM, one wave)
      const int warps = std::min(kWarpsPerCta, std::max(1, (a.chunks_x + 1) / 2));
      const size_t dyn_cnt = sizeof(unsigned) * 8 * (size_t)a.chunks_x;
      const size_t dyn_wr = dyn_cnt + (size_t)warps * kChunkTiles * 32 * sizeof(unsigned short);
      if (dyn_wr > ((size_t)40 << 10)) {
        DPOSE_CUDA(cudaFuncSetAttribute(k_extract<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn_cnt));
        DPOSE_CUDA(cudaFuncSetAttribute(k_extract<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn_wr));
      }
      ExtractArgs e;
      e.box = a;
      e.totals = ws->d_totals;
      e.groups = ws->d_groups;
      e.n_groups = n_groups;
      e.out = nullptr;
      e.cap = 0;
      DPOSE_CUDA(cudaMemsetAsync(ws->d_groups, 0, sizeof(unsigned long long) * (n_groups + 1), st));
      k_extract<false><<<(unsigned)a.n_tile_rows, warps * 32, dyn_cnt, st>>>(e);
      DPOSE_LAUNCHED();
      DPOSE_CUDA(cudaMemcpyAsync(ws->h_total, d_total, sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
      for (int attempt = 0; attempt < 2; ++attempt) {
        if (!cap) {  // no estimate (or it was too small): the exact total first
          DPOSE_CUDA(cudaStreamSynchronize(st));
          total = *ws->h_total;
          if (total == 0) break;
          cap = total;
        }
        DPOSE_CUDA(cudaMallocAsync(&d_out, sizeof(int2) * cap, st));
        e.out = d_out;
        e.cap = cap;
        k_extract<true><<<(unsigned)a.n_tile_rows, warps * 32, dyn_wr, st>>>(e);
        DPOSE_LAUNCHED();
        DPOSE_CUDA(cudaEventRecord(ws->ev1, st));
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
