// get_cost.cu — pose_gradient::get_cost on the GPU (replaces dpose_core.cpp:158-260).
//
// Math (all FP64; table values are float32 promoted exactly, as in the reference):
//   k = C + R(-theta) (p - t)                      map cell p -> kernel frame   (:206-209,:227,:234)
//   k_upper = round(k) (half away), valid iff 2 <= k_upper <= (cols-2, rows-2)   (:177-182)
//   c_rel = k - k_upper + 0.5; f = a00 + a10 cx + a01 cy + a11 cx cy              (:193-197)
// The reference's Jacobian is a central difference (h = 1e-6) of f along x, y, theta on the
// SAME patch (:244-256); f is bilinear, so that difference equals the analytic patch
// derivative up to round-off / O(h^2).  With a = k_u - C_x, b = k_v - C_y,
// f_u = a10 + a11 cy, f_v = a01 + a11 cx:
//   J_x = -c Sum f_u + s Sum f_v,  J_y = -s Sum f_u - c Sum f_v,  J_theta = Sum (f_u b - f_v a)
//
// k_cost_cells: one pose, explicit cell list (the reference's call shape); grid-strided, fixed-order block
// reduction, last CTA folds the CTA partials in index order.  The batched kernels (many poses against the whole
// costmap) live in get_cost_batch_tile.cu / get_cost_batch_coop.cu.
#include <cmath>
#include <atomic>
#include <cstring>
#include <mutex>

#include "cost_math.cuh"

namespace dpose {

namespace {

struct Accum {
  double f, su, sv, st;
};

// one lethal cell with offset (dx, dy) = p - t; adds its contribution
template <bool kJac>
__device__ __forceinline__ void eval_cell(const PatchCoef *__restrict__ coef, int rows, int cols,
                                          double cx, double cy, double c, double s, double dx,
                                          double dy, Accum &acc) {
  const double a = fma(c, dx, s * dy);   // k_u - C_x
  const double b = fma(c, dy, -s * dx);  // k_v - C_y
  const double ku = cx + a, kv = cy + b;
  const double ru = round(ku), rv = round(kv);  // half away from zero == std::round (:177)
  // |k| beyond any table -> reject before the int cast can saturate oddly
  if (!(fabs(ru) < 1e9) || !(fabs(rv) < 1e9)) return;
  const int ux = (int)ru, uy = (int)rv;
  if (ux < 2 || uy < 2 || ux > cols - 2 || uy > rows - 2) return;  // :181-182
  const double rx = (ku - ru) + 0.5, ry = (kv - rv) + 0.5;         // :193
  const PatchCoef p = coef[uy * cols + ux];
  const double fv = fma(p.a11, rx, p.a01);
  acc.f += fma(ry, fv, fma(p.a10, rx, p.a00));
  if (kJac) {
    const double fu = fma(p.a11, ry, p.a10);
    acc.su += fu;
    acc.sv += fv;
    acc.st += fma(fu, b, -fv * a);
  }
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
  return v;
}

__device__ __forceinline__ void finish(const Accum &t, double c, double s, double *out) {
  out[0] = t.f;
  out[1] = -c * t.su + s * t.sv;
  out[2] = -s * t.su - c * t.sv;
  out[3] = t.st;
}

// ---- single pose, explicit cell list ------------------------------------------------------
constexpr int kCellThreads = 256;

template <bool kJac>
__global__ void __launch_bounds__(kCellThreads)
    k_cost_cells(const PatchCoef *__restrict__ coef, TableGeom g, double tx, double ty, double th,
                 const int2 *__restrict__ cells, size_t n, double *partials, unsigned *ticket,
                 double *out, unsigned long long seq) {
  __shared__ double s_part[kCellThreads / 32][4];
  __shared__ bool s_last;
  double s, c;
  sincos(th, &s, &c);
  Accum acc = {0.0, 0.0, 0.0, 0.0};
  const size_t stride = (size_t)gridDim.x * kCellThreads;
  for (size_t i = (size_t)blockIdx.x * kCellThreads + threadIdx.x; i < n; i += stride) {
    const int2 p = cells[i];
    eval_cell<kJac>(coef, g.rows, g.cols, g.cx, g.cy, c, s, (double)p.x - tx, (double)p.y - ty, acc);
  }
  acc.f = warp_sum(acc.f);
  acc.su = warp_sum(acc.su);
  acc.sv = warp_sum(acc.sv);
  acc.st = warp_sum(acc.st);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) {
    s_part[warp][0] = acc.f;
    s_part[warp][1] = acc.su;
    s_part[warp][2] = acc.sv;
    s_part[warp][3] = acc.st;
  }
  __syncthreads();
  if (gridDim.x == 1) {  // the usual case (up to 1024 cells): no partials, no ticket, no trip through global memory
    if (threadIdx.x == 0) {
      Accum t = {0.0, 0.0, 0.0, 0.0};
      for (int w = 0; w < kCellThreads / 32; ++w) {  // warp index order, as the multi-CTA path below
        t.f += s_part[w][0];
        t.su += s_part[w][1];
        t.sv += s_part[w][2];
        t.st += s_part[w][3];
      }
      finish(t, c, s, out);
      __threadfence_system();
      *reinterpret_cast<volatile unsigned long long *>(out + 4) = seq;
    }
    return;
  }
  if (threadIdx.x < 4) {
    double v = 0.0;
    for (int w = 0; w < kCellThreads / 32; ++w) v += s_part[w][threadIdx.x];
    partials[(size_t)blockIdx.x * 4 + threadIdx.x] = v;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = atomicAdd(ticket, 1u) == gridDim.x - 1;
  __syncthreads();
  if (s_last && threadIdx.x == 0) {
    __threadfence();
    Accum t = {0.0, 0.0, 0.0, 0.0};
    const volatile double *vp = partials;
    for (unsigned b = 0; b < gridDim.x; ++b) {  // CTA index order: deterministic
      t.f += vp[4 * b];
      t.su += vp[4 * b + 1];
      t.sv += vp[4 * b + 2];
      t.st += vp[4 * b + 3];
    }
    finish(t, c, s, out);
    *ticket = 0u;
    // out is mapped pinned host memory: publish the four doubles, then the call's sequence number the host spins on
    __threadfence_system();
    *reinterpret_cast<volatile unsigned long long *>(out + 4) = seq;
  }
}

struct CellScratch {
  std::mutex mu;
  int device = -1;
  int2 *d_cells = nullptr;
  size_t cap = 0;
  double *d_partials = nullptr;  // kMaxCellCtas * 4
  unsigned *d_ticket = nullptr;
  double *h_out = nullptr;  // pinned host memory the kernel writes into (UVA): 4 doubles + the call's sequence number
  unsigned long long seq = 0;
  // Small cell lists (the reference's call shape: tens to hundreds of cells per call) are not copied to the device
  // at all: the caller's cells are memcpy'd into this mapped pinned page and the kernel reads them over PCIe — one
  // round trip all threads share, instead of a staged H2D copy and its API call (~4 us of a 14 us call).
  int2 *h_cells = nullptr;
  cudaStream_t stream = nullptr;  // private non-blocking stream: no implicit sync with the caller's streams
};
constexpr int kMaxCellCtas = 592;  // 4 per SM
constexpr size_t kZeroCopyCells = 4096;  // 32 KB of mapped pinned memory

}  // namespace

// per-device scratch for the single-pose entry points (grow-only, mutex-protected)
static CellScratch g_scratch[64];

static int ensure_scratch(CellScratch &sc, int device, size_t ncells) {
  if (!sc.d_partials) {
    DPOSE_CUDA(cudaMalloc(&sc.d_partials, sizeof(double) * 4 * kMaxCellCtas));
    DPOSE_CUDA(cudaMalloc(&sc.d_ticket, sizeof(unsigned)));
    DPOSE_CUDA(cudaMemset(sc.d_ticket, 0, sizeof(unsigned)));
    DPOSE_CUDA(cudaStreamSynchronize(nullptr));  // the memset is asynchronous; its users run on a non-blocking stream
    DPOSE_CUDA(cudaHostAlloc(&sc.h_out, sizeof(double) * 8, cudaHostAllocMapped));
    std::memset(sc.h_out, 0, sizeof(double) * 8);
    DPOSE_CUDA(cudaHostAlloc(&sc.h_cells, sizeof(int2) * kZeroCopyCells, cudaHostAllocMapped));
    DPOSE_CUDA(cudaStreamCreateWithFlags(&sc.stream, cudaStreamNonBlocking));
    sc.device = device;
  }
  if (ncells > sc.cap) {
    cudaFree(sc.d_cells);
    sc.d_cells = nullptr;
    sc.cap = 0;
    const size_t cap = ncells + ncells / 2 + 1024;
    DPOSE_CUDA(cudaMalloc(&sc.d_cells, sizeof(int2) * cap));
    sc.cap = cap;
  }
  return DPOSE_OK;
}

// d_cells == nullptr && h_cells != nullptr: upload first
static int cost_cells_impl(const dpose_pg *pg, const double se2[3], const int2 *d_cells,
                           const int32_t *h_cells, size_t n, double *cost, double *J) {
  if (pg->device < 0 || pg->device >= 64) return fail(DPOSE_EINVAL, "device index out of range");
  DeviceGuard guard(pg->device);
  if (!guard.ok) return fail(DPOSE_ECUDA, "cudaSetDevice(%d) failed", pg->device);
  CellScratch &sc = g_scratch[pg->device];
  std::lock_guard<std::mutex> lock(sc.mu);
  const bool zero_copy = h_cells && n <= kZeroCopyCells;
  DPOSE_TRY(ensure_scratch(sc, pg->device, h_cells && !zero_copy ? n : 0));
  if (zero_copy) {
    std::memcpy(sc.h_cells, h_cells, sizeof(int2) * n);
    d_cells = sc.h_cells;  // UVA: the mapped host page is addressable from the device as is
  } else if (h_cells && n) {
    DPOSE_CUDA(cudaMemcpyAsync(sc.d_cells, h_cells, sizeof(int2) * n, cudaMemcpyHostToDevice, sc.stream));
    d_cells = sc.d_cells;
  }
  const TableGeom g = {pg->rows, pg->cols, (double)pg->center[0], (double)pg->center[1]};
  int grid = (int)((n + 4 * kCellThreads - 1) / (4 * kCellThreads));
  grid = grid < 1 ? 1 : (grid > kMaxCellCtas ? kMaxCellCtas : grid);
  const unsigned long long seq = ++sc.seq;
  if (J)
    k_cost_cells<true><<<grid, kCellThreads, 0, sc.stream>>>(pg->d_coef, g, se2[0], se2[1], se2[2], d_cells, n,
                                                             sc.d_partials, sc.d_ticket, sc.h_out, seq);
  else
    k_cost_cells<false><<<grid, kCellThreads, 0, sc.stream>>>(pg->d_coef, g, se2[0], se2[1], se2[2], d_cells, n,
                                                              sc.d_partials, sc.d_ticket, sc.h_out, seq);
  DPOSE_LAUNCHED();
  // The result lands in pinned host memory; spinning on its sequence number returns a few microseconds before
  // cudaStreamSynchronize would.  The stream is queried now and then so that a failed launch cannot hang the caller,
  // and synchronised at the end of a long wait (a debugger, a preempted process) to stay correct either way.
  {
    const volatile unsigned long long *flag = reinterpret_cast<const volatile unsigned long long *>(sc.h_out + 4);
    bool done = false;
    for (unsigned spin = 0; spin < (1u << 22) && !done; ++spin) {
      done = *flag == seq;
      if (!done && (spin & 0x3ff) == 0x3ff) {
        const cudaError_t q = cudaStreamQuery(sc.stream);
        if (q != cudaErrorNotReady) break;  // finished (flag visible on the next read) or failed
      }
    }
    if (!done) DPOSE_CUDA(cudaStreamSynchronize(sc.stream));
    std::atomic_thread_fence(std::memory_order_acquire);
  }
  const double *h = sc.h_out;
  *cost = h[0];
  if (J) {
    J[0] = h[1];
    J[1] = h[2];
    J[2] = h[3];
  }
  return DPOSE_OK;
}

int get_cost_cells(const dpose_pg *pg, const double se2[3], const int2 *d_cells, size_t n,
                   double *cost, double *J) {
  return cost_cells_impl(pg, se2, d_cells, nullptr, n, cost, J);
}

int get_cost_cells_host(const dpose_pg *pg, const double se2[3], const int32_t *cells_xy, size_t n,
                        double *cost, double *J) {
  return cost_cells_impl(pg, se2, nullptr, cells_xy, n, cost, J);
}

uint64_t window_bytes_host(const dpose_pg *pg, const dpose_map *map, const double *poses, size_t n) {
  const TableGeom g = {pg->rows, pg->cols, (double)pg->center[0], (double)pg->center[1]};
  uint64_t total = 0;
  for (size_t i = 0; i < n; ++i) {
    const double c = std::cos(poses[3 * i + 2]), s = std::sin(poses[3 * i + 2]);
    const WindowRect w = pose_window(g, poses[3 * i], poses[3 * i + 1], c, s, (int)map->size_x, (int)map->size_y);
    if (w.x1 >= w.x0 && w.y1 >= w.y0) total += (uint64_t)(w.x1 - w.x0 + 1) * (uint64_t)(w.y1 - w.y0 + 1);
  }
  return total;
}

}  // namespace dpose
