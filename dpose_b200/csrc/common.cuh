// common.cuh — handle layouts, error plumbing and launch accounting shared by the
// translation units of libdpose_b200.so.  Not part of the public surface.
#pragma once

#include <cuda_runtime.h>

#include <atomic>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/dpose_b200.h"

namespace dpose {

constexpr uint8_t kLethal = 254;  // costmap_2d::LETHAL_OBSTACLE

// ---- error plumbing -----------------------------------------------------------------
std::string &last_error();
int fail(int code, const char *fmt, ...);
extern std::atomic<uint64_t> g_launches;

#define DPOSE_CUDA(expr)                                                              \
  do {                                                                                \
    cudaError_t _e = (expr);                                                          \
    if (_e != cudaSuccess)                                                            \
      return ::dpose::fail(_e == cudaErrorNoDevice || _e == cudaErrorInsufficientDriver \
                               ? DPOSE_ENODEV                                         \
                               : DPOSE_ECUDA,                                         \
                           "%s:%d: %s -> %s", __FILE__, __LINE__, #expr,              \
                           cudaGetErrorString(_e));                                   \
  } while (0)

#define DPOSE_TRY(expr)          \
  do {                           \
    int _rc = (expr);            \
    if (_rc != DPOSE_OK) return _rc; \
  } while (0)

// counts the launch and reports launch-configuration errors
#define DPOSE_LAUNCHED()                          \
  do {                                            \
    ::dpose::g_launches.fetch_add(1, std::memory_order_relaxed); \
    DPOSE_CUDA(cudaGetLastError());               \
  } while (0)

// Launch with programmatic stream serialisation allowed: the kernel may start while its predecessor in the stream is
// still draining (the launch latency and the prologue overlap it).  The kernel must execute griddepcontrol.wait before
// it touches anything the predecessor wrote; predecessors call griddepcontrol.launch_dependents as early as they like.
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                              Args &&...args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
#define DPOSE_PDL_WAIT() asm volatile("griddepcontrol.wait;" ::: "memory")
#define DPOSE_PDL_TRIGGER() asm volatile("griddepcontrol.launch_dependents;" ::: "memory")

// device attributes the launch code needs on every call, queried once per device
struct DeviceInfo {
  int sms = 148, smem_optin = 0;
};
const DeviceInfo &device_info(int device);  // api.cu

struct DeviceGuard {
  int prev = -1;
  bool ok = false;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    ok = cudaSetDevice(dev) == cudaSuccess;
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};

// per-patch bilinear coefficients, indexed by k_upper = (ux, uy):
//   f(cx, cy) = a00 + a10*cx + a01*cy + a11*cx*cy,   c_rel in [0,1)^2
// a00 = T(uy-1,ux-1), a10 = T(uy-1,ux)-a00, a01 = T(uy,ux-1)-a00,
// a11 = T(uy,ux) - T(uy-1,ux) - T(uy,ux-1) + a00  (all exact in FP64)
struct __align__(32) PatchCoef {
  double a00, a10, a01, a11;
};

}  // namespace dpose

// ---- handles --------------------------------------------------------------------------
struct dpose_pg {
  int device = 0;
  int rows = 0, cols = 0;
  int32_t center[2] = {0, 0};
  int32_t box[10] = {0};
  uint32_t padding = 0;
  // device
  float *d_cost = nullptr;          // rows*cols
  dpose::PatchCoef *d_coef = nullptr;  // rows*cols, valid for 2<=ux<=cols-2, 2<=uy<=rows-2
  // host copies
  std::vector<float> h_cost, h_pre_blur;
  std::vector<int32_t> h_edt_sq;
  std::vector<uint8_t> h_outline, h_mask;
  float precompute_us = 0.f;
  // batched kernel geometry: side of the largest window (cells) any pose can touch
  int win_max = 0;
  // replicated, bank-interleaved copy of d_coef for the thread-per-pose kernel (get_cost_batch_tile.cu),
  // built on first use: R = 1 << coef_image_shift replicas (0: first halves twice, -1: nothing in shared memory;
  // kNoCoefImage: not built yet)
  void *d_coef_image = nullptr;
  static constexpr int kNoCoefImage = -100;
  int coef_image_shift = kNoCoefImage;
  // texture modes (coef_image_shift <= 0): the patch polynomials as a plain global array behind a texture object
  void *d_coef_hi = nullptr;
  unsigned long long tex_hi = 0;
  // one 32-byte record (b00, b10, b01, b11) per patch, absolute k' coordinates, zeros outside the valid range: what the
  // cooperative kernels gather from global memory (coef_abs_ensure, built on first use)
  double *d_coef_abs = nullptr;
  // the footprint as handed to dpose_pg_create (cells, robot frame): check_footprint of solved poses (solve.cu)
  std::vector<int32_t> footprint;
  // streams, events and staging buffers of the host-buffer batch entry point (api.cu), built on first use
  void *batch_ctx = nullptr;
};

struct dpose_map {
  int device = 0;
  uint32_t size_x = 0, size_y = 0;
  size_t pitch = 0;  // bytes, multiple of 32 (pad bytes are zero)
  uint8_t *d_data = nullptr;
  // lethal tile plane (get_cost_batch_tile.cu): bit = (cell == 254), in 32 x 8 cell tiles of 32 bytes: word
  // ((ty * tiles_ntx + tx) * 8 + (y & 7)), bit x & 31; tiles_ntx includes one zero pad tile per tile row.
  // Rebuilt from d_data when `tiles_dirty` (after create / update; only the rows [lo, hi) that changed) or inside
  // every batched call when the map is marked volatile (its bytes may change behind the library's back).
  // tiles_dirty / the dirty rows / is_volatile are guarded by map_state_mutex() (get_cost_batch_tile.cu).
  uint32_t *d_tiles = nullptr;
  uint32_t tiles_ntx = 0, tiles_nty = 0;
  bool tiles_dirty = true;
  uint32_t tiles_dirty_lo = 0, tiles_dirty_hi = 0xffffffffu;
  bool is_volatile = false;
  cudaEvent_t tiles_ready = nullptr;  // recorded after a rebuild; readers on other streams wait on it
  // scratch buffers, pinned staging and stream of lethal_cells_within (lethal.cu), built on first use
  void *lethal_ws = nullptr;
  // pinned staging of dpose_map_update for small dirty windows (a 2-D copy out of pageable memory is staged row
  // by row by the driver: 40 us for 92 x 92 cells; packed into pinned memory first it is one DMA), built on first use
  uint8_t *h_update_stage = nullptr;
  size_t update_stage_cap = 0;
};

struct dpose_cells {
  int device = 0;
  size_t n = 0;
  int2 *d_xy = nullptr;
  bool pooled = false;  // d_xy comes from the stream-ordered allocator (lethal extraction): freed with cudaFreeAsync
};

namespace dpose {
// api.cu
void pg_release_batch_ctx(dpose_pg *pg);
// precompute.cu
int precompute_run(dpose_pg *pg, const int32_t *shifted_fp_xy, int n);
int precompute_fetch_stages(dpose_pg *pg);
// lethal.cu
int lethal_extract(const dpose_map *map, const int32_t box[10], int2 **d_cells, size_t *n);
void map_release_lethal_ws(dpose_map *map);
float lethal_last_kernel_us(const dpose_map *map);
size_t host_raytrace(const int32_t begin[2], const int32_t end[2], int32_t *out_xy, size_t cap);
// get_cost.cu
int get_cost_cells(const dpose_pg *pg, const double se2[3], const int2 *d_cells, size_t n,
                   double *cost, double *J);
int get_cost_cells_host(const dpose_pg *pg, const double se2[3], const int32_t *cells_xy, size_t n,
                        double *cost, double *J);
uint64_t window_bytes_host(const dpose_pg *pg, const dpose_map *map, const double *poses, size_t n);
// What a batched launch evaluates and where the results go.
struct BatchSpec {
  const double *d_poses = nullptr;  // n x 3 (x, y, theta), or null: the poses are the lattice `sweep`
  dpose_sweep sweep = {};           // pose i = (x0 + dx ix, y0 + dy iy, theta0 + dtheta it), i = (it ny + iy) nx + ix
  size_t sweep_base = 0;            // lattice index of this launch's first pose (chunked host-buffer calls)
  size_t n = 0;
  double *d_out = nullptr;          // n x 4 (cost, Jx, Jy, Jtheta), or n costs if cost_only
  int want_jacobian = 1;
  int cost_only = 0;
  // do not rebuild the lethal plane of a volatile map (the host-buffer entry point streams the chunks of ONE call
  // through this function and rebuilds for the first chunk only)
  bool hold_planes = false;
  // poses of the whole CALL this launch is a part of (0: the launch is the call).  The kernel is chosen by it, not by
  // the size of a staging chunk / lattice slab: the cooperative kernel sums a pose's cells in another order, and the
  // bits of a pose must not depend on where a chunk boundary falls (a 131 073-pose call used to run its last pose there)
  size_t call_n = 0;
  // band of map rows [row_lo, row_hi) the launch is restricted to (row_hi == 0: the whole map): only lethal cells of
  // these rows count, and only these rows of a volatile map's tile plane are rebuilt in front of the launch — a caller
  // that shards its poses spatially (a rank's slice of a sweep sorted by map row) shards the map pass with them
  uint32_t row_lo = 0, row_hi = 0;
};
int get_cost_batch_spec(const dpose_pg *pg, const dpose_map *map, const BatchSpec &spec, cudaStream_t stream);
int get_cost_batch_device(const dpose_pg *pg, const dpose_map *map, const double *d_poses,
                          size_t n, double *d_out, int want_jacobian, cudaStream_t stream,
                          bool hold_planes = false, uint32_t row_lo = 0, uint32_t row_hi = 0);
// lattice pose of index gi, the same bits as numpy's x0 + dx * ix (one rounded product, one rounded sum)
__host__ __device__ inline void sweep_pose(const dpose_sweep &sw, size_t gi, double *x, double *y, double *th) {
  const size_t ix = gi % sw.nx, r = gi / sw.nx;
  const size_t iy = r % sw.ny, it = r / sw.ny;
#ifdef __CUDA_ARCH__
  *x = __dadd_rn(sw.x0, __dmul_rn(sw.dx, (double)ix));
  *y = __dadd_rn(sw.y0, __dmul_rn(sw.dy, (double)iy));
  *th = __dadd_rn(sw.theta0, __dmul_rn(sw.dtheta, (double)it));
#else
  *x = sw.x0 + sw.dx * (double)ix;
  *y = sw.y0 + sw.dy * (double)iy;
  *th = sw.theta0 + sw.dtheta * (double)it;
#endif
}
// get_cost_batch_tile.cu: the lethal tile plane, the thread-per-pose batch kernel (large batches) and the dispatcher
std::mutex &map_state_mutex();  // guards the lazily built per-handle state (coefficient images, tile-plane dirty flags)
int tileplane_ensure(const dpose_map *map, cudaStream_t stream, bool hold_planes, uint32_t row_lo = 0, uint32_t row_hi = 0);
int coef_abs_ensure(const dpose_pg *pg, cudaStream_t stream);
int get_cost_batch_tile(const dpose_pg *pg, const dpose_map *map, const BatchSpec &spec, cudaStream_t stream);
// get_cost_batch_coop.cu: G lanes per pose, for batches too small to fill the GPU with one thread per pose; used when
// the batch would get at least DPOSE_COOP_MIN_GROUP lanes per pose.
// B200, 17 x 25 table, clean plane, device time per launch (profiles/r2_ab_tile_walk.txt, runs v28 / v29):
//   poses              1000   4096   9000   18000
//   G lanes per pose   6.1    7.1    8.4    10.3  us   (G = 32, 32, 16, 8)
//   one thread/pose    6.5    6.7    6.7    6.7   us
// Since the thread-per-pose kernel lost its staging wait and its conversion-pipe traffic it wins from a few thousand poses
// on (round 2 started with the crossover at 18 944 poses): the cooperative BATCH kernel is kept for G = 32 only
// (<= 4736 poses); coop_eval itself is what k_solve runs on.
#ifndef DPOSE_COOP_MIN_GROUP
#define DPOSE_COOP_MIN_GROUP 32
#endif
bool coop_batch_wanted(const dpose_pg *pg, size_t n);
int get_cost_batch_coop(const dpose_pg *pg, const dpose_map *map, const BatchSpec &spec, cudaStream_t stream);
// host window geometry shared by the kernel launch code and dpose_pg_window_bytes
struct WindowRect {
  int x0, y0, x1, y1;  // inclusive cell range, already clipped; empty if x1 < x0
};
}  // namespace dpose
