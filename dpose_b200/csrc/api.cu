// api.cu — the extern "C" surface of libdpose_b200.so (include/dpose_b200.h): argument
// validation, handle lifetime, host<->device staging and the multi-GPU shard driver.
#include <algorithm>
#include <atomic>
#include <climits>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <thread>

#include "common.cuh"

namespace dpose {

std::atomic<uint64_t> g_launches{0};

std::string &last_error() {
  thread_local std::string e;
  return e;
}

const DeviceInfo &device_info(int device) {
  static DeviceInfo info[64];
  static std::once_flag once[64];
  const int d = device >= 0 && device < 64 ? device : 0;
  std::call_once(once[d], [d] {
    cudaDeviceGetAttribute(&info[d].sms, cudaDevAttrMultiProcessorCount, d);
    cudaDeviceGetAttribute(&info[d].smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, d);
    cudaGetLastError();
  });
  return info[d];
}

int fail(int code, const char *fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  last_error() = buf;
  return code;
}

static int check_device(int device) {
  int n = 0;
  const cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0)
    return fail(DPOSE_ENODEV, "no CUDA device available (%s); libdpose_b200 has no CPU fallback",
                e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
  if (device < 0 || device >= n) return fail(DPOSE_EINVAL, "device %d out of range [0, %d)", device, n);
  // once per device: keep freed stream-ordered allocations (results of lethal_cells_within, the per-call workspaces of
  // check_footprint and the solve) in the device's default pool instead of handing them back to the OS at every
  // synchronisation — with the default threshold of 0 each call pays a fresh device allocation (~0.4 ms)
  static std::once_flag once[64];
  if (device < 64)
    std::call_once(once[device], [device] {
      cudaMemPool_t pool;
      if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
        unsigned long long keep = ~0ull;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
      }
      cudaGetLastError();
    });
  return DPOSE_OK;
}

}  // namespace dpose

using namespace dpose;

namespace {
constexpr int kBuf = 3;
constexpr int kMaxTableSide = 512;
// poses per staging chunk of the host-buffer call: 3 MB in / 4 MB out, short pipeline fill and drain.  Measured on
// B200 (cfg5, 1e7 poses): 2^16 .. 2^21 poses give 7.50 / 6.68 / 6.54 / 6.58 / 6.79 / 7.27 ms per call on one box,
// 7.54 / 6.68 / 7.03 on another (profiles/r2_ab_tile_walk.txt, run v52): with both copy directions busy 2^17 is the
// robust choice.  Calls that keep ONE copy direction busy — cost only (8 B per pose back), a lattice sweep (nothing in) —
// take chunks of twice that: cost-only 1.95e9 -> 2.13e9 poses/s, sweep 1.68e9 -> 1.71e9.
#ifndef DPOSE_BATCH_CHUNK_LOG2
#define DPOSE_BATCH_CHUNK_LOG2 17
#endif
constexpr size_t kChunkPoses = (size_t)1 << DPOSE_BATCH_CHUNK_LOG2;

// Calls of up to this many poses (the multi-start pattern: a few thousand candidates per solver iteration) skip the
// device staging altogether: poses and results live in mapped pinned memory the kernel reads and writes over PCIe
// (no staged copies of pageable buffers on either side).  4096 poses: 51 -> 42 us per call.
constexpr size_t kZeroCopyPoses = 8192;

struct BatchCtx {
  std::mutex mu;
  size_t cap = 0;  // poses per staging buffer
  double *h_zc = nullptr;  // mapped pinned: kZeroCopyPoses x 3 poses | kZeroCopyPoses x 4 results
  double *zc_in() const { return h_zc; }
  double *zc_out() const { return h_zc + 3 * kZeroCopyPoses; }
  int ensure_zc() {
    if (h_zc) return DPOSE_OK;
    DPOSE_CUDA(cudaHostAlloc(&h_zc, sizeof(double) * 7 * kZeroCopyPoses, cudaHostAllocMapped));
    return DPOSE_OK;
  }
  cudaStream_t s_in = nullptr, s_k = nullptr, s_out = nullptr;
  cudaEvent_t in_done[kBuf] = {}, k_done[kBuf] = {}, out_done[kBuf] = {};
  double *d_p[kBuf] = {}, *d_o[kBuf] = {};

  int init() {
    DPOSE_CUDA(cudaStreamCreateWithFlags(&s_in, cudaStreamNonBlocking));
    DPOSE_CUDA(cudaStreamCreateWithFlags(&s_k, cudaStreamNonBlocking));
    DPOSE_CUDA(cudaStreamCreateWithFlags(&s_out, cudaStreamNonBlocking));
    for (int b = 0; b < kBuf; ++b) {
      DPOSE_CUDA(cudaEventCreateWithFlags(&in_done[b], cudaEventDisableTiming));
      DPOSE_CUDA(cudaEventCreateWithFlags(&k_done[b], cudaEventDisableTiming));
      DPOSE_CUDA(cudaEventCreateWithFlags(&out_done[b], cudaEventDisableTiming));
    }
    return DPOSE_OK;
  }
  int reserve(size_t poses) {
    if (poses <= cap) return DPOSE_OK;
    for (int b = 0; b < kBuf; ++b) {
      cudaFree(d_p[b]);
      cudaFree(d_o[b]);
      d_p[b] = d_o[b] = nullptr;
    }
    cap = 0;
    for (int b = 0; b < kBuf; ++b) {
      DPOSE_CUDA(cudaMalloc(&d_p[b], sizeof(double) * 3 * poses));
      DPOSE_CUDA(cudaMalloc(&d_o[b], sizeof(double) * 4 * poses));
    }
    cap = poses;
    return DPOSE_OK;
  }
  void release() {
    for (int b = 0; b < kBuf; ++b) {
      cudaFree(d_p[b]);
      cudaFree(d_o[b]);
      if (in_done[b]) cudaEventDestroy(in_done[b]);
      if (k_done[b]) cudaEventDestroy(k_done[b]);
      if (out_done[b]) cudaEventDestroy(out_done[b]);
    }
    if (s_in) cudaStreamDestroy(s_in);
    if (s_k) cudaStreamDestroy(s_k);
    if (s_out) cudaStreamDestroy(s_out);
    if (h_zc) cudaFreeHost(h_zc);
  }
};

std::mutex g_ctx_mu;  // guards creation of the per-handle context
}  // namespace

namespace dpose {
void pg_release_batch_ctx(dpose_pg *pg) {
  BatchCtx *ctx = static_cast<BatchCtx *>(pg->batch_ctx);
  if (!ctx) return;
  ctx->release();
  delete ctx;
  pg->batch_ctx = nullptr;
}
}  // namespace dpose


extern "C" {

const char *dpose_last_error(void) { return last_error().c_str(); }

int dpose_device_count(int *n) {
  if (!n) return fail(DPOSE_EINVAL, "n is NULL");
  *n = 0;
  const cudaError_t e = cudaGetDeviceCount(n);
  if (e != cudaSuccess) {
    *n = 0;
    return fail(DPOSE_ENODEV, "cudaGetDeviceCount: %s", cudaGetErrorString(e));
  }
  return DPOSE_OK;
}

uint64_t dpose_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

int dpose_host_alloc(void **ptr, size_t bytes) {
  if (!ptr) return fail(DPOSE_EINVAL, "ptr is NULL");
  DPOSE_CUDA(cudaHostAlloc(ptr, bytes ? bytes : 1, cudaHostAllocPortable));
  return DPOSE_OK;
}

void dpose_host_free(void *ptr) {
  if (ptr) cudaFreeHost(ptr);
}

void dpose_free(void *ptr) { std::free(ptr); }

// ---- footprint ----------------------------------------------------------------------------
int dpose_make_footprint(const double *xy_m, int n, double resolution, int32_t *xy_cells) {
  if (!(resolution > 0)) return fail(DPOSE_EINVAL, "resolution must be positive");
  if (n < 0 || (n > 0 && (!xy_m || !xy_cells))) return fail(DPOSE_EINVAL, "bad footprint buffer");
  for (int i = 0; i < 2 * n; ++i) xy_cells[i] = (int32_t)std::round(xy_m[i] / resolution);
  return DPOSE_OK;
}

// ---- world <-> cell poses (dpose_goal_tolerance.cpp:294-330) -----------------------------------------------
int dpose_world_to_cell(const double *world, size_t n, double origin_x, double origin_y, double resolution,
                        uint32_t size_x, uint32_t size_y, double *cells, uint8_t *ok) {
  if (n && (!world || !cells)) return fail(DPOSE_EINVAL, "NULL argument");
  if (!(resolution > 0)) return fail(DPOSE_EINVAL, "resolution must be positive");
  for (size_t i = 0; i < n; ++i) {
    const double wx = world[3 * i], wy = world[3 * i + 1];
    bool good = !(wx < origin_x || wy < origin_y) && std::isfinite(wx) && std::isfinite(wy);  // Costmap2D::worldToMap
    double mx = 0, my = 0;
    if (good) {
      mx = std::floor((wx - origin_x) / resolution);  // (int) cast of a non-negative value
      my = std::floor((wy - origin_y) / resolution);
      good = mx < (double)size_x && my < (double)size_y;
    }
    cells[3 * i] = good ? mx : 0.0;
    cells[3 * i + 1] = good ? my : 0.0;
    cells[3 * i + 2] = world[3 * i + 2];
    if (ok) ok[i] = good ? 1 : 0;
  }
  return DPOSE_OK;
}

int dpose_cell_to_world(const double *cells, size_t n, double origin_x, double origin_y, double resolution,
                        double *world, uint8_t *ok) {
  if (n && (!cells || !world)) return fail(DPOSE_EINVAL, "NULL argument");
  for (size_t i = 0; i < n; ++i) {
    const double cx = cells[3 * i], cy = cells[3 * i + 1];
    const bool good = !(cx < 0 || cy < 0) && std::isfinite(cx) && std::isfinite(cy);  // :314-315 "negative pose"
    const double mx = good ? std::round(cx) : 0.0, my = good ? std::round(cy) : 0.0;   // :318-319
    world[3 * i] = origin_x + (mx + 0.5) * resolution;  // Costmap2D::mapToWorld
    world[3 * i + 1] = origin_y + (my + 0.5) * resolution;
    world[3 * i + 2] = cells[3 * i + 2];
    if (ok) ok[i] = good ? 1 : 0;
  }
  return DPOSE_OK;
}

// ---- pose_gradient / cost_data ----------------------------------------------------------------
static int pg_create_impl(const int32_t *fp, int n, uint32_t padding, int device, dpose_pg **out);

int dpose_pg_create(const int32_t *fp, int n, uint32_t padding, int device, dpose_pg **out) {
  try {  // no exception crosses the C ABI
    return pg_create_impl(fp, n, padding, device, out);
  } catch (const std::bad_alloc &) {
    return fail(DPOSE_ENOMEM, "out of host memory");
  } catch (const std::exception &e) {
    return fail(DPOSE_ECUDA, "dpose_pg_create: %s", e.what());
  }
}

static int pg_create_impl(const int32_t *fp, int n, uint32_t padding, int device, dpose_pg **out) {
  if (!out) return fail(DPOSE_EINVAL, "out is NULL");
  *out = nullptr;
  if (n < 3 || !fp) return fail(DPOSE_EINVAL, "footprint must contain at least three points");
  if (n > 4096) return fail(DPOSE_EINVAL, "footprint has %d vertices; at most 4096 supported", n);
  if (padding > 4096) return fail(DPOSE_EINVAL, "padding %u too large", padding);
  DPOSE_TRY(check_device(device));
  long long minx = LLONG_MAX, miny = LLONG_MAX, maxx = LLONG_MIN, maxy = LLONG_MIN;
  for (int i = 0; i < n; ++i) {
    minx = std::min<long long>(minx, fp[2 * i]);
    maxx = std::max<long long>(maxx, fp[2 * i]);
    miny = std::min<long long>(miny, fp[2 * i + 1]);
    maxy = std::max<long long>(maxy, fp[2 * i + 1]);
  }
  const long long cols = maxx - minx + 1 + 2ll * padding, rows = maxy - miny + 1 + 2ll * padding;
  // 512 x 512 cells is a 25 m footprint at 0.05 m: beyond that neither the single-CTA precompute (< 10 ms up to here)
  // nor the fp32 window clip of the batched kernel (get_cost_batch_tile.cu, clip_margin_for) is specified
  if (cols > kMaxTableSide || rows > kMaxTableSide)
    return fail(DPOSE_EINVAL, "cost table %lld x %lld too large (at most %d x %d cells)", rows, cols, kMaxTableSide,
                kMaxTableSide);
  dpose_pg *pg = new (std::nothrow) dpose_pg();
  if (!pg) return fail(DPOSE_ENOMEM, "out of host memory");
  pg->device = device;
  pg->padding = padding;
  pg->footprint.assign(fp, fp + 2 * (size_t)n);
  pg->rows = (int)rows;
  pg->cols = (int)cols;
  // dpose_core.cpp:124-125 center = padding - rowwise min; :138 box = ring(cols, rows) - center
  pg->center[0] = (int32_t)((long long)padding - minx);
  pg->center[1] = (int32_t)((long long)padding - miny);
  const int bx[5] = {0, pg->cols, pg->cols, 0, 0}, by[5] = {0, 0, pg->rows, pg->rows, 0};
  for (int i = 0; i < 5; ++i) {
    pg->box[2 * i] = bx[i] - pg->center[0];
    pg->box[2 * i + 1] = by[i] - pg->center[1];
  }
  const double a = pg->cols - 3, b = pg->rows - 3;
  pg->win_max = (int)std::floor(std::sqrt(a * a + b * b)) + 2;
  std::vector<int32_t> shifted(2 * (size_t)n);  // :128
  for (int i = 0; i < n; ++i) {
    shifted[2 * i] = fp[2 * i] + pg->center[0];
    shifted[2 * i + 1] = fp[2 * i + 1] + pg->center[1];
  }
  int rc;
  {
    DeviceGuard guard(device);
    rc = guard.ok ? precompute_run(pg, shifted.data(), n) : fail(DPOSE_ECUDA, "cudaSetDevice(%d) failed", device);
  }
  if (rc != DPOSE_OK) {
    dpose_pg_destroy(pg);
    return rc;
  }
  *out = pg;
  return DPOSE_OK;
}

void dpose_pg_destroy(dpose_pg *pg) {
  if (!pg) return;
  {
    DeviceGuard guard(pg->device);
    if (pg->d_cost) cudaFreeAsync(pg->d_cost, nullptr);  // one stream-ordered allocation: [cost | coef] (precompute.cu)
    cudaFree(pg->d_coef_image);
    if (pg->tex_hi) cudaDestroyTextureObject((cudaTextureObject_t)pg->tex_hi);
    cudaFree(pg->d_coef_hi);
    cudaFree(pg->d_coef_abs);
    pg_release_batch_ctx(pg);
  }
  delete pg;
}

int dpose_pg_table(const dpose_pg *pg, const float **cost, int *rows, int *cols, int32_t center[2],
                   int32_t box[10]) {
  if (!pg) return fail(DPOSE_EINVAL, "pg is NULL");
  if (cost) *cost = pg->h_cost.data();
  if (rows) *rows = pg->rows;
  if (cols) *cols = pg->cols;
  if (center) std::memcpy(center, pg->center, sizeof(pg->center));
  if (box) std::memcpy(box, pg->box, sizeof(pg->box));
  return DPOSE_OK;
}

int dpose_pg_stages(const dpose_pg *pg, const int32_t **edt_sq, const uint8_t **outline,
                    const uint8_t **mask, const float **pre_blur) {
  if (!pg) return fail(DPOSE_EINVAL, "pg is NULL");
  {
    static std::mutex mu;  // the lazy download mutates the handle
    std::lock_guard<std::mutex> lock(mu);
    DPOSE_TRY(precompute_fetch_stages(const_cast<dpose_pg *>(pg)));
  }
  if (edt_sq) *edt_sq = pg->h_edt_sq.data();
  if (outline) *outline = pg->h_outline.data();
  if (mask) *mask = pg->h_mask.data();
  if (pre_blur) *pre_blur = pg->h_pre_blur.data();
  return DPOSE_OK;
}

int dpose_pg_precompute_us(const dpose_pg *pg, float *us) {
  if (!pg || !us) return fail(DPOSE_EINVAL, "NULL argument");
  *us = pg->precompute_us;
  return DPOSE_OK;
}

// ---- costmap ------------------------------------------------------------------------------------
static int map_create_impl(const uint8_t *data, uint32_t size_x, uint32_t size_y, size_t pitch,
                           int device, cudaMemcpyKind kind, dpose_map **out) {
  if (!out) return fail(DPOSE_EINVAL, "out is NULL");
  *out = nullptr;
  if ((size_x && size_y && !data) || pitch < size_x) return fail(DPOSE_EINVAL, "bad costmap buffer / pitch");
  if (size_x > (1u << 30) || size_y > (1u << 30)) return fail(DPOSE_EINVAL, "costmap too large");
  DPOSE_TRY(check_device(device));
  dpose_map *m = new (std::nothrow) dpose_map();
  if (!m) return fail(DPOSE_ENOMEM, "out of host memory");
  m->device = device;
  m->size_x = size_x;
  m->size_y = size_y;
  m->pitch = ((size_t)size_x + 31) & ~(size_t)31;  // 32-byte rows: k_pack_tiles reads a tile row with one 256-bit load
  auto body = [&]() -> int {
    DeviceGuard guard(device);
    if (!guard.ok) return fail(DPOSE_ECUDA, "cudaSetDevice(%d) failed", device);
    if (size_x == 0 || size_y == 0) return DPOSE_OK;
    DPOSE_CUDA(cudaMalloc(&m->d_data, m->pitch * size_y));
    if (m->pitch != size_x) DPOSE_CUDA(cudaMemset(m->d_data, 0, m->pitch * size_y));
    DPOSE_CUDA(cudaMemcpy2D(m->d_data, m->pitch, data, pitch, size_x, size_y, kind));
    // A "synchronous" copy out of pageable host memory returns once the bytes are STAGED; the DMA may still be in
    // flight, and only work on the legacy stream is ordered behind it.  The kernels that read this map run on
    // non-blocking streams, so wait for the copy (and the memset) to land.
    DPOSE_CUDA(cudaStreamSynchronize(nullptr));
    m->tiles_ntx = (size_x + 31) / 32 + 1;
    m->tiles_nty = (size_y + 7) / 8;
    DPOSE_CUDA(cudaMalloc(&m->d_tiles, (size_t)m->tiles_ntx * m->tiles_nty * 32));
    m->tiles_dirty = true;
    return DPOSE_OK;
  };
  const int rc = body();
  if (rc != DPOSE_OK) {
    dpose_map_destroy(m);
    return rc;
  }
  *out = m;
  return DPOSE_OK;
}

int dpose_map_create(const uint8_t *data, uint32_t size_x, uint32_t size_y, size_t pitch, int device,
                     dpose_map **out) {
  return map_create_impl(data, size_x, size_y, pitch, device, cudaMemcpyHostToDevice, out);
}

int dpose_map_create_from_device(const uint8_t *d_data, uint32_t size_x, uint32_t size_y, size_t pitch,
                                 int device, dpose_map **out) {
  return map_create_impl(d_data, size_x, size_y, pitch, device, cudaMemcpyDeviceToDevice, out);
}

int dpose_map_update(dpose_map *map, const uint8_t *data, size_t pitch, uint32_t x0, uint32_t y0,
                     uint32_t w, uint32_t h) {
  if (!map || !data) return fail(DPOSE_EINVAL, "NULL argument");
  if ((uint64_t)x0 + w > map->size_x || (uint64_t)y0 + h > map->size_y || pitch < w)
    return fail(DPOSE_EINVAL, "update window outside the map");
  if (w == 0 || h == 0) return DPOSE_OK;
  DeviceGuard guard(map->device);
  if (!guard.ok) return fail(DPOSE_ECUDA, "cudaSetDevice(%d) failed", map->device);
  constexpr size_t kStageMax = (size_t)1 << 20;
  const size_t bytes = (size_t)w * h;
  if (bytes <= kStageMax && h > 1) {
    if (map->update_stage_cap < bytes) {
      if (map->h_update_stage) cudaFreeHost(map->h_update_stage);
      map->h_update_stage = nullptr;
      map->update_stage_cap = 0;
      const size_t cap = std::max<size_t>(bytes * 2, (size_t)64 << 10);
      DPOSE_CUDA(cudaHostAlloc(&map->h_update_stage, std::min(cap, kStageMax), cudaHostAllocDefault));
      map->update_stage_cap = std::min(cap, kStageMax);
    }
    for (uint32_t r = 0; r < h; ++r) std::memcpy(map->h_update_stage + (size_t)r * w, data + (size_t)r * pitch, w);
    DPOSE_CUDA(cudaMemcpy2DAsync(map->d_data + (size_t)y0 * map->pitch + x0, map->pitch, map->h_update_stage, w, w, h,
                                 cudaMemcpyHostToDevice, nullptr));
    DPOSE_CUDA(cudaStreamSynchronize(nullptr));  // the staging buffer is reused by the next call
  } else {
    DPOSE_CUDA(cudaMemcpy2D(map->d_data + (size_t)y0 * map->pitch + x0, map->pitch, data, pitch, w, h,
                            cudaMemcpyHostToDevice));
    DPOSE_CUDA(cudaStreamSynchronize(nullptr));  // staged copy: see dpose_map_create
  }
  // narrow the lazy plane rebuild to the rows of this window (union with what is already pending)
  std::lock_guard<std::mutex> lock(map_state_mutex());
  if (!map->tiles_dirty) {
    map->tiles_dirty_lo = y0;
    map->tiles_dirty_hi = y0 + h;
  } else {
    map->tiles_dirty_lo = std::min(map->tiles_dirty_lo, y0);
    map->tiles_dirty_hi = std::max(map->tiles_dirty_hi, y0 + h);
  }
  map->tiles_dirty = true;
  return DPOSE_OK;
}

int dpose_map_set_volatile(dpose_map *map, int is_volatile) {
  if (!map) return fail(DPOSE_EINVAL, "map is NULL");
  std::lock_guard<std::mutex> lock(map_state_mutex());
  map->is_volatile = is_volatile != 0;
  map->tiles_dirty = true;
  map->tiles_dirty_lo = 0;
  map->tiles_dirty_hi = 0xffffffffu;
  return DPOSE_OK;
}

int dpose_map_device_ptr(const dpose_map *map, uint8_t **d_data, size_t *pitch) {
  if (!map || !d_data || !pitch) return fail(DPOSE_EINVAL, "NULL argument");
  *d_data = map->d_data;
  *pitch = map->pitch;
  return DPOSE_OK;
}

int dpose_map_size(const dpose_map *map, uint32_t *size_x, uint32_t *size_y) {
  if (!map) return fail(DPOSE_EINVAL, "map is NULL");
  if (size_x) *size_x = map->size_x;
  if (size_y) *size_y = map->size_y;
  return DPOSE_OK;
}

void dpose_map_destroy(dpose_map *map) {
  if (!map) return;
  {
    DeviceGuard guard(map->device);
    map_release_lethal_ws(map);
    cudaFree(map->d_data);
    if (map->h_update_stage) cudaFreeHost(map->h_update_stage);
    cudaFree(map->d_tiles);
    if (map->tiles_ready) cudaEventDestroy(map->tiles_ready);
  }
  delete map;
}

// ---- lethal cells -----------------------------------------------------------------------------------
int dpose_raytrace(const int32_t begin[2], const int32_t end[2], int32_t *out_xy, size_t cap, size_t *n) {
  if (!begin || !end || !n || (cap && !out_xy)) return fail(DPOSE_EINVAL, "NULL argument");
  *n = host_raytrace(begin, end, out_xy, cap);
  return DPOSE_OK;
}

int dpose_lethal_cells_within_device(const dpose_map *map, const int32_t box[10], dpose_cells **out) {
  if (!map || !box || !out) return fail(DPOSE_EINVAL, "NULL argument");
  *out = nullptr;
  DPOSE_TRY(check_device(map->device));
  dpose_cells *c = new (std::nothrow) dpose_cells();
  if (!c) return fail(DPOSE_ENOMEM, "out of host memory");
  c->device = map->device;
  c->pooled = true;
  const int rc = lethal_extract(map, box, &c->d_xy, &c->n);
  if (rc != DPOSE_OK) {
    delete c;
    return rc;
  }
  *out = c;
  return DPOSE_OK;
}

int dpose_map_lethal_us(const dpose_map *map, float *us) {
  if (!map || !us) return fail(DPOSE_EINVAL, "NULL argument");
  *us = lethal_last_kernel_us(map);
  return DPOSE_OK;
}

int dpose_lethal_cells_within(const dpose_map *map, const int32_t box[10], int32_t **cells_xy, size_t *n) {
  if (!cells_xy || !n) return fail(DPOSE_EINVAL, "NULL argument");
  *cells_xy = nullptr;
  *n = 0;
  dpose_cells *c = nullptr;
  DPOSE_TRY(dpose_lethal_cells_within_device(map, box, &c));
  int rc = DPOSE_OK;
  if (c->n) {
    int32_t *h = (int32_t *)std::malloc(sizeof(int32_t) * 2 * c->n);
    if (!h)
      rc = fail(DPOSE_ENOMEM, "out of host memory");
    else if ((rc = dpose_cells_download(c, h)) == DPOSE_OK) {
      *cells_xy = h;
      *n = c->n;
    } else
      std::free(h);
  }
  dpose_cells_destroy(c);
  return rc;
}

int dpose_cells_upload(const int32_t *cells_xy, size_t n, int device, dpose_cells **out) {
  if (!out || (n && !cells_xy)) return fail(DPOSE_EINVAL, "NULL argument");
  *out = nullptr;
  DPOSE_TRY(check_device(device));
  dpose_cells *c = new (std::nothrow) dpose_cells();
  if (!c) return fail(DPOSE_ENOMEM, "out of host memory");
  c->device = device;
  c->n = n;
  auto body = [&]() -> int {
    DeviceGuard guard(device);
    if (!guard.ok) return fail(DPOSE_ECUDA, "cudaSetDevice(%d) failed", device);
    if (!n) return DPOSE_OK;
    DPOSE_CUDA(cudaMalloc(&c->d_xy, sizeof(int2) * n));
    DPOSE_CUDA(cudaMemcpy(c->d_xy, cells_xy, sizeof(int2) * n, cudaMemcpyHostToDevice));
    DPOSE_CUDA(cudaStreamSynchronize(nullptr));  // staged copy: see dpose_map_create
    return DPOSE_OK;
  };
  const int rc = body();
  if (rc != DPOSE_OK) {
    dpose_cells_destroy(c);
    return rc;
  }
  *out = c;
  return DPOSE_OK;
}

int dpose_cells_size(const dpose_cells *cells, size_t *n) {
  if (!cells || !n) return fail(DPOSE_EINVAL, "NULL argument");
  *n = cells->n;
  return DPOSE_OK;
}

int dpose_cells_download(const dpose_cells *cells, int32_t *cells_xy) {
  if (!cells || (cells->n && !cells_xy)) return fail(DPOSE_EINVAL, "NULL argument");
  if (!cells->n) return DPOSE_OK;
  DeviceGuard guard(cells->device);
  if (!guard.ok) return fail(DPOSE_ECUDA, "cudaSetDevice(%d) failed", cells->device);
  DPOSE_CUDA(cudaMemcpy(cells_xy, cells->d_xy, sizeof(int2) * cells->n, cudaMemcpyDeviceToHost));
  return DPOSE_OK;
}

void dpose_cells_destroy(dpose_cells *cells) {
  if (!cells) return;
  {
    DeviceGuard guard(cells->device);
    if (cells->pooled && cells->d_xy)
      cudaFreeAsync(cells->d_xy, nullptr);  // back to the pool without a device-wide synchronisation
    else
      cudaFree(cells->d_xy);
  }
  delete cells;
}

// ---- get_cost -----------------------------------------------------------------------------------------
int dpose_pg_get_cost_cells(const dpose_pg *pg, const double se2[3], const int32_t *cells_xy, size_t n,
                            double *cost, double *J) {
  if (!pg || !se2 || !cost || (n && !cells_xy)) return fail(DPOSE_EINVAL, "NULL argument");
  return get_cost_cells_host(pg, se2, cells_xy, n, cost, J);
}

int dpose_pg_get_cost_cells_device(const dpose_pg *pg, const double se2[3], const dpose_cells *cells,
                                   double *cost, double *J) {
  if (!pg || !se2 || !cost || !cells) return fail(DPOSE_EINVAL, "NULL argument");
  if (cells->device != pg->device) return fail(DPOSE_EINVAL, "cells and pose_gradient live on different devices");
  return get_cost_cells(pg, se2, cells->d_xy, cells->n, cost, J);
}

int dpose_pg_get_cost_batch_device(const dpose_pg *pg, const dpose_map *map, const double *d_poses,
                                   size_t n, double *d_out, int want_jacobian, void *stream) {
  if (!pg || !map || (n && (!d_poses || !d_out))) return fail(DPOSE_EINVAL, "NULL argument");
  return get_cost_batch_device(pg, map, d_poses, n, d_out, want_jacobian, (cudaStream_t)stream);
}

int dpose_order_poses_by_tile(const double *poses, size_t n, uint32_t size_x, uint32_t size_y, uint64_t *perm) {
  if (n && (!poses || !perm)) return fail(DPOSE_EINVAL, "NULL argument");
  if (size_x == 0 || size_y == 0) return fail(DPOSE_EINVAL, "empty map");
  try {
    // a stable counting sort over the tiles of the lethal plane (32 x 8 cells, row-major) + one bucket for NaN positions
    const uint64_t ntx = ((uint64_t)size_x + 31) / 32, nty = ((uint64_t)size_y + 7) / 8, nb = ntx * nty + 1;
    auto bucket = [&](size_t i) -> uint64_t {
      const double x = poses[3 * i], y = poses[3 * i + 1];
      if (x != x || y != y) return nb - 1;
      const double cx = std::min(std::max(x, 0.0), (double)size_x - 1.0), cy = std::min(std::max(y, 0.0), (double)size_y - 1.0);
      return ((uint64_t)cy / 8) * ntx + (uint64_t)cx / 32;
    };
    std::vector<uint64_t> start(nb + 1, 0);
    std::vector<uint32_t> key32;
    std::vector<uint64_t> key64;
    const bool narrow = nb <= 0xffffffffull;
    if (narrow) key32.resize(n); else key64.resize(n);
    for (size_t i = 0; i < n; ++i) {
      const uint64_t b = bucket(i);
      if (narrow) key32[i] = (uint32_t)b; else key64[i] = b;
      ++start[b + 1];
    }
    for (uint64_t b = 0; b < nb; ++b) start[b + 1] += start[b];
    for (size_t i = 0; i < n; ++i) perm[start[narrow ? key32[i] : key64[i]]++] = i;
    return DPOSE_OK;
  } catch (const std::exception &e) {
    return fail(DPOSE_ENOMEM, "dpose_order_poses_by_tile: %s", e.what());
  }
}

int dpose_pg_reach(const dpose_pg *pg, uint32_t *reach) {
  if (!pg || !reach) return fail(DPOSE_EINVAL, "NULL argument");
  double r2 = 0.0;  // farthest corner of the table from its centre (the valid region, :181-182, lies 1.5 cells inside)
  for (int i = 0; i < 4; ++i) r2 = std::max(r2, (double)pg->box[2 * i] * pg->box[2 * i] + (double)pg->box[2 * i + 1] * pg->box[2 * i + 1]);
  *reach = (uint32_t)std::ceil(std::sqrt(r2)) + 1u;
  return DPOSE_OK;
}

int dpose_pg_get_cost_batch_device_rows(const dpose_pg *pg, const dpose_map *map, const double *d_poses, size_t n,
                                        double *d_out, int want_jacobian, void *stream, uint32_t row_lo, uint32_t row_hi) {
  if (!pg || !map || (n && (!d_poses || !d_out))) return fail(DPOSE_EINVAL, "NULL argument");
  if (row_lo >= row_hi || row_hi > map->size_y)
    return fail(DPOSE_EINVAL, "row band [%u, %u) is empty or outside the map (%u rows)", row_lo, row_hi, map->size_y);
  return get_cost_batch_device(pg, map, d_poses, n, d_out, want_jacobian, (cudaStream_t)stream, false, row_lo, row_hi);
}

// host buffers: H2D / kernel / D2H chunks on three streams, triple-buffered device staging.  The
// streams, events and staging buffers live in the pose_gradient handle (created on first use, grown
// on demand), so a call costs no cudaMalloc / cudaFree; concurrent calls on one handle serialise.
// One body for the three host-buffer entry points: pose list or lattice in, n x 4 results or n costs out.
struct HostBatch {
  const double *poses = nullptr;      // n x 3, or null with `sweep`
  const dpose_sweep *sweep = nullptr;
  size_t n = 0;
  double *out = nullptr;
  int want_jacobian = 1, cost_only = 0;
};

static int batch_host(const dpose_pg *pg, const dpose_map *map, const HostBatch &hb) {
  const size_t n = hb.n;
  if (n == 0) return DPOSE_OK;
  DPOSE_TRY(check_device(pg->device));
  DeviceGuard guard(pg->device);
  if (!guard.ok) return fail(DPOSE_ECUDA, "cudaSetDevice(%d) failed", pg->device);
  BatchCtx *ctx;
  {
    std::lock_guard<std::mutex> lock(g_ctx_mu);
    dpose_pg *p = const_cast<dpose_pg *>(pg);
    if (!p->batch_ctx) {
      BatchCtx *fresh = new (std::nothrow) BatchCtx();
      if (!fresh) return fail(DPOSE_ENOMEM, "out of host memory");
      const int rc = fresh->init();
      if (rc != DPOSE_OK) {
        fresh->release();
        delete fresh;
        return rc;
      }
      p->batch_ctx = fresh;
    }
    ctx = static_cast<BatchCtx *>(p->batch_ctx);
  }
  std::lock_guard<std::mutex> call_lock(ctx->mu);
  const size_t ow = hb.cost_only ? 1 : 4;  // doubles per pose that travel back
  BatchSpec spec;
  if (hb.sweep) spec.sweep = *hb.sweep;
  spec.want_jacobian = hb.want_jacobian;
  spec.cost_only = hb.cost_only;
  spec.call_n = n;
  if (n <= kZeroCopyPoses) {  // latency path, see kZeroCopyPoses
    DPOSE_TRY(ctx->ensure_zc());
    if (hb.poses) std::memcpy(ctx->zc_in(), hb.poses, sizeof(double) * 3 * n);
    spec.d_poses = hb.poses ? ctx->zc_in() : nullptr;
    spec.n = n;
    spec.d_out = ctx->zc_out();
    int rc = get_cost_batch_spec(pg, map, spec, ctx->s_k);
    if (rc == DPOSE_OK) {
      // A stream synchronisation, not a flag polled in the result page: the results are written by every thread of
      // the kernel, and only a synchronisation point guarantees that all of them have reached host memory.
      const cudaError_t e = cudaStreamSynchronize(ctx->s_k);
      if (e != cudaSuccess) rc = fail(DPOSE_ECUDA, "batched get_cost failed: %s", cudaGetErrorString(e));
      if (rc == DPOSE_OK) std::memcpy(hb.out, ctx->zc_out(), sizeof(double) * ow * n);
    }
    if (rc != DPOSE_OK) cudaDeviceSynchronize();
    return rc;
  }
  const size_t chunk = std::min<size_t>(n, hb.cost_only || !hb.poses ? 2 * kChunkPoses : kChunkPoses);
  DPOSE_TRY(ctx->reserve(chunk));
  auto body = [&]() -> int {
    size_t it = 0;
    for (size_t off = 0; off < n; off += chunk, ++it) {
      const int b = (int)(it % kBuf);
      const size_t m = std::min(chunk, n - off);
      if (it >= kBuf) {  // buffer reuse: its previous D2H must be complete
        DPOSE_CUDA(cudaStreamWaitEvent(hb.poses ? ctx->s_in : ctx->s_k, ctx->out_done[b], 0));
      }
      if (hb.poses) {
        DPOSE_CUDA(cudaMemcpyAsync(ctx->d_p[b], hb.poses + 3 * off, sizeof(double) * 3 * m, cudaMemcpyHostToDevice, ctx->s_in));
        DPOSE_CUDA(cudaEventRecord(ctx->in_done[b], ctx->s_in));
        DPOSE_CUDA(cudaStreamWaitEvent(ctx->s_k, ctx->in_done[b], 0));
      }
      spec.d_poses = hb.poses ? ctx->d_p[b] : nullptr;
      spec.sweep_base = off;
      spec.n = m;
      spec.d_out = ctx->d_o[b];
      spec.hold_planes = it > 0;  // volatile maps: the lethal plane is rebuilt for the first chunk only (once per call)
      DPOSE_TRY(get_cost_batch_spec(pg, map, spec, ctx->s_k));
      DPOSE_CUDA(cudaEventRecord(ctx->k_done[b], ctx->s_k));
      DPOSE_CUDA(cudaStreamWaitEvent(ctx->s_out, ctx->k_done[b], 0));
      DPOSE_CUDA(cudaMemcpyAsync(hb.out + ow * off, ctx->d_o[b], sizeof(double) * ow * m, cudaMemcpyDeviceToHost, ctx->s_out));
      DPOSE_CUDA(cudaEventRecord(ctx->out_done[b], ctx->s_out));
    }
    DPOSE_CUDA(cudaStreamSynchronize(ctx->s_out));
    DPOSE_CUDA(cudaStreamSynchronize(ctx->s_k));
    DPOSE_CUDA(cudaStreamSynchronize(ctx->s_in));
    return DPOSE_OK;
  };
  const int rc = body();
  if (rc != DPOSE_OK) cudaDeviceSynchronize();
  return rc;
}

static int sweep_count(const dpose_sweep *sw, size_t *n) {
  if (!sw) return fail(DPOSE_EINVAL, "sweep is NULL");
  const unsigned __int128 total = (unsigned __int128)sw->nx * sw->ny * sw->ntheta;
  if (total > ((unsigned __int128)1 << 40)) return fail(DPOSE_EINVAL, "sweep lattice too large");
  *n = (size_t)total;
  return DPOSE_OK;
}

int dpose_pg_get_cost_batch(const dpose_pg *pg, const dpose_map *map, const double *poses, size_t n,
                            double *out, int want_jacobian) {
  if (!pg || !map || (n && (!poses || !out))) return fail(DPOSE_EINVAL, "NULL argument");
  HostBatch hb;
  hb.poses = poses, hb.n = n, hb.out = out, hb.want_jacobian = want_jacobian;
  return batch_host(pg, map, hb);
}

int dpose_pg_get_cost_batch_costs(const dpose_pg *pg, const dpose_map *map, const double *poses, size_t n,
                                  double *cost) {
  if (!pg || !map || (n && (!poses || !cost))) return fail(DPOSE_EINVAL, "NULL argument");
  HostBatch hb;
  hb.poses = poses, hb.n = n, hb.out = cost, hb.want_jacobian = 0, hb.cost_only = 1;
  return batch_host(pg, map, hb);
}

int dpose_pg_get_cost_sweep(const dpose_pg *pg, const dpose_map *map, const dpose_sweep *sweep, double *out,
                            int want_jacobian, int cost_only) {
  if (!pg || !map) return fail(DPOSE_EINVAL, "NULL argument");
  size_t n = 0;
  DPOSE_TRY(sweep_count(sweep, &n));
  if (n && !out) return fail(DPOSE_EINVAL, "out is NULL");
  HostBatch hb;
  hb.sweep = sweep, hb.n = n, hb.out = out, hb.want_jacobian = want_jacobian, hb.cost_only = cost_only != 0;
  return batch_host(pg, map, hb);
}

int dpose_pg_get_cost_sweep_device(const dpose_pg *pg, const dpose_map *map, const dpose_sweep *sweep, double *d_out,
                                   int want_jacobian, int cost_only, void *stream) {
  if (!pg || !map) return fail(DPOSE_EINVAL, "NULL argument");
  size_t n = 0;
  DPOSE_TRY(sweep_count(sweep, &n));
  if (n && !d_out) return fail(DPOSE_EINVAL, "d_out is NULL");
  BatchSpec spec;
  spec.sweep = *sweep;
  spec.n = n;
  spec.d_out = d_out;
  spec.want_jacobian = want_jacobian;
  spec.cost_only = cost_only != 0;
  return get_cost_batch_spec(pg, map, spec, (cudaStream_t)stream);
}

int dpose_pg_get_cost_batch_multi(dpose_pg *const *pgs, dpose_map *const *maps, int ndev,
                                  const double *poses, size_t n, double *out, int want_jacobian) {
  if (!pgs || !maps || ndev < 1 || (n && (!poses || !out))) return fail(DPOSE_EINVAL, "bad argument");
  for (int d = 0; d < ndev; ++d)
    if (!pgs[d] || !maps[d]) return fail(DPOSE_EINVAL, "NULL handle for shard %d", d);
  if (ndev == 1) return dpose_pg_get_cost_batch(pgs[0], maps[0], poses, n, out, want_jacobian);
  // contiguous shards, one host thread per device, no collective
  std::vector<std::thread> th;
  try {
    std::vector<int> rcs((size_t)ndev, DPOSE_OK);
    std::vector<std::string> errs((size_t)ndev);
    const size_t per = (n + ndev - 1) / ndev;
    for (int d = 0; d < ndev; ++d) {
      const size_t lo = std::min(n, per * d), hi = std::min(n, per * (d + 1));
      th.emplace_back([&, d, lo, hi]() {
        rcs[d] = dpose_pg_get_cost_batch(pgs[d], maps[d], poses + 3 * lo, hi - lo, out + 4 * lo, want_jacobian);
        if (rcs[d] != DPOSE_OK) errs[d] = last_error();
      });
    }
    for (auto &t : th) t.join();
    th.clear();
    for (int d = 0; d < ndev; ++d)
      if (rcs[d] != DPOSE_OK) return fail(rcs[d], "shard %d: %s", d, errs[d].c_str());
    return DPOSE_OK;
  } catch (const std::exception &e) {  // thread creation / allocation failed: no exception crosses the C ABI
    for (auto &t : th)
      if (t.joinable()) t.join();
    return fail(DPOSE_ENOMEM, "dpose_pg_get_cost_batch_multi: %s", e.what());
  }
}

int dpose_pg_window_bytes(const dpose_pg *pg, const dpose_map *map, const double *poses, size_t n,
                          uint64_t *bytes) {
  if (!pg || !map || !bytes || (n && !poses)) return fail(DPOSE_EINVAL, "NULL argument");
  *bytes = window_bytes_host(pg, map, poses, n);
  return DPOSE_OK;
}

}  // extern "C"
