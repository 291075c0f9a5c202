// get_cost_batch_tile.cu — batched get_cost(poses[]) for B200 (sm_100a), one THREAD per pose.
//
// Why this shape.  The warp-per-pose kernel (get_cost_batch.cu) spends ~355 warp instructions and 78
// shared-memory wavefronts per pose (ncu, profiles/r1_v2_*): a pose has only ~31 valid lethal cells, so
// warp-wide scans, prefix sums, compaction queues and shuffle reductions cost more than the
// arithmetic.  Here every lane owns a pose: nothing is exchanged between lanes, there is no barrier
// in the main loop and no reduction at all — a pose's sum is accumulated by one thread in a fixed
// (row-major) order, so the result is deterministic and independent of the batch position.
//
// Data layout.
//   * lethal plane in 32 x 8 cell tiles (tileplane, below): one tile = 8 rows x 32 bits = 32 bytes =
//     one L2 sector, tiles row-major with one zero pad tile per tile row.  A 26 x 26 window touches
//     <= 2 x 5 tiles (6.7 on average) instead of 26 separate sectors of a row-major plane; a thread
//     fetches a tile with ONE 256-bit load (LDG.E.256, new on sm_100).
//   * patch coefficients (4 doubles per interpolation patch) live in shared memory, replicated R
//     times and interleaved at 16-byte granularity: lane l reads replica l % R, so for R = 8 the
//     eight lanes of an LDS.128 phase hit eight different 16-byte bank groups whatever patches
//     they address — random gathers become conflict-free.  The replicated image is built once per
//     pose_gradient in global memory and staged by a TMA bulk copy (cp.async.bulk -> UBLKCP) per CTA.
//   * per-thread row masks in shared memory, [row][thread] (a thread only touches its own column:
//     conflict-free, no synchronisation).
//
// Per pose (thread), per 32 x 32 chunk of its window (one chunk for tables whose windows fit):
//   phase 1  load the <= 5 x 2 tiles, funnel-shift each row's 32 window bits into place, AND them
//            with the row's x-interval of the rotated kernel rectangle (fp32, conservative by 0.02
//            cell; linear in the row index, so 4 FMAs per row), park the row mask, note non-empty
//            rows in a 32-bit row mask.
//   phase 2  walk the set bits (find-first-set over row mask, then mask): k + 0.5 = k0 + x (c,-s) +
//            y (s, c) in FP64; floor by adding 1.5 * 2^52 rounding down; exact bounds test
//            (dpose_core.cpp:181-182); two LDS.128 for the patch; f, f_u, f_v by FMA; accumulate.
// Tensor cores are not used: there is no dense contraction on this path.
//
// Algorithmic bytes per pose: 24 (pose) + 32 (result) + |window| map bytes.
#include <algorithm>
#include <cstdlib>
#include <mutex>

#include "cost_math.cuh"

namespace dpose {
namespace {

// ---- tile plane ---------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned lethal_bits16(const uint4 v) {
  const unsigned k = 0xFEFEFEFEu;
  const unsigned e0 = __vcmpeq4(v.x, k) & 0x01010101u, e1 = __vcmpeq4(v.y, k) & 0x01010101u;
  const unsigned e2 = __vcmpeq4(v.z, k) & 0x01010101u, e3 = __vcmpeq4(v.w, k) & 0x01010101u;
  const unsigned n0 = (e0 * 0x00204081u) >> 21 & 0xFu, n1 = (e1 * 0x00204081u) >> 21 & 0xFu;
  const unsigned n2 = (e2 * 0x00204081u) >> 21 & 0xFu, n3 = (e3 * 0x00204081u) >> 21 & 0xFu;
  return n0 | (n1 << 4) | (n2 << 8) | (n3 << 12);
}

// One thread per plane word.  Lane l of a warp handles row (l >> 2) of tile column 4 * warp + (l & 3):
// the warp reads 8 map rows x 128 contiguous bytes and writes 4 tiles = 128 contiguous bytes.
__global__ void __launch_bounds__(256) k_pack_tiles(const uint8_t *__restrict__ map, size_t pitch, uint32_t size_y,
                                                    uint32_t *__restrict__ tiles, uint32_t ntx, uint32_t nty) {
  const uint32_t quads = (ntx + 3) / 4;  // groups of 4 tile columns
  const size_t n_warps = (size_t)quads * nty;
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t r = lane >> 2, j = lane & 3;
  for (size_t w = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; w < n_warps;
       w += ((size_t)gridDim.x * blockDim.x) >> 5) {
    const uint32_t by = (uint32_t)(w / quads), bx = (uint32_t)(w - (size_t)by * quads) * 4 + j;
    if (bx >= ntx) continue;
    const uint32_t y = by * 8 + r;
    const size_t x = (size_t)bx * 32;
    unsigned out = 0u;
    if (y < size_y) {  // pitch % 16 == 0 and the pad bytes of the device copy are zero
      const uint8_t *row = map + (size_t)y * pitch;
      if (x < pitch) out = lethal_bits16(__ldg(reinterpret_cast<const uint4 *>(row + x)));
      if (x + 16 < pitch) out |= lethal_bits16(__ldg(reinterpret_cast<const uint4 *>(row + x + 16))) << 16;
    }
    tiles[((size_t)by * ntx + bx) * 8 + r] = out;
  }
}

// ---- replicated coefficient image -----------------------------------------------------------------------
// image[((patch * 2 + half) * R + rep)] (16 bytes each) = half `half` of PatchCoef[patch]
__global__ void k_replicate_coef(const uint4 *__restrict__ coef, uint4 *__restrict__ image, int n_halves, int rep_shift) {
  const int total = n_halves << rep_shift;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x)
    image[i] = __ldg(coef + (i >> rep_shift));
}

// ---- PTX wrappers ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

struct Tile8 {
  unsigned w[8];
};
__device__ __forceinline__ Tile8 ldg_tile(const uint32_t *p) {  // one 32-byte sector, read-only path
  Tile8 t;
  asm volatile("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(t.w[0]), "=r"(t.w[1]), "=r"(t.w[2]), "=r"(t.w[3]), "=r"(t.w[4]), "=r"(t.w[5]), "=r"(t.w[6]),
                 "=r"(t.w[7])
               : "l"(p));
  return t;
}
// PTX shifts clamp the amount to 32 (C++ shifts by >= 32 are undefined)
__device__ __forceinline__ unsigned shl_clamp(unsigned v, unsigned n) {
  unsigned r;
  asm("shl.b32 %0, %1, %2;" : "=r"(r) : "r"(v), "r"(n));
  return r;
}
__device__ __forceinline__ unsigned shr_clamp(unsigned v, unsigned n) {
  unsigned r;
  asm("shr.b32 %0, %1, %2;" : "=r"(r) : "r"(v), "r"(n));
  return r;
}
__device__ __forceinline__ float rcp_approx(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ void stg_result(double *p, double a, double b, double c, double d) {  // 32 B, one sector
  asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(p), "d"(a), "d"(b), "d"(c), "d"(d) : "memory");
}
// small non-negative int -> double without the conversion pipe: 2^52 + x has x in its low mantissa bits
__device__ __forceinline__ double small_int_to_double(int x) {
  return __hiloint2double(0x43300000, x) - 4503599627370496.0;
}

struct TileArgs {
  const uint4 *coef_image;  // replicated coefficient image (global)
  uint32_t coef_bytes;      // its size in bytes (multiple of 128)
  int rep_shift;            // log2(R)
  TableGeom g;
  int size_x, size_y;
  const uint32_t *tiles;
  uint32_t ntx;
  const double *poses;
  double *out;
  size_t n;
};

struct Acc4 {
  double f, su, sv, st;
};

template <bool kJac, int kThreads>
__global__ void __launch_bounds__(kThreads, 1) k_cost_batch_tile(TileArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) unsigned long long s_bar;
  const int tid = threadIdx.x, lane = tid & 31;

  // ---- stage the replicated coefficient image with TMA bulk copies -----------------------------------------
  const uint32_t coef_u32 = smem_u32(smem);
  const uint32_t bar = smem_u32(&s_bar);
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(a.coef_bytes) : "memory");
    const unsigned char *src = reinterpret_cast<const unsigned char *>(a.coef_image);
    for (uint32_t off = 0; off < a.coef_bytes; off += 16384) {
      const uint32_t sz = min(16384u, a.coef_bytes - off);
      asm volatile(
          "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(coef_u32 + off),
          "l"(src + off), "r"(sz), "r"(bar)
          : "memory");
    }
  }
  __syncthreads();  // barrier initialised before anyone polls it
  {
    uint32_t ok = 0;
    for (uint32_t spin = 0; spin < (1u << 24) && !ok; ++spin)  // bounded: a lost copy must not hang the GPU
      asm volatile(
          "{\n\t.reg .pred p;\n\t"
          "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\t"
          "selp.u32 %0, 1, 0, p;\n\t}"
          : "=r"(ok)
          : "r"(bar)
          : "memory");
    if (!ok) __trap();
  }

  uint32_t *s_mask = reinterpret_cast<uint32_t *>(smem + a.coef_bytes) + tid;  // [row * kThreads]
  const uint32_t my_coef = coef_u32 + ((uint32_t)lane & ((1u << a.rep_shift) - 1u)) * 16u;
  const int patch_shift = a.rep_shift + 5;  // patch stride = 2 halves * R * 16 bytes
  const uint32_t half_off = 16u << a.rep_shift;
  const int cols = a.g.cols;
  const unsigned cols_m4 = (unsigned)(a.g.cols - 4), rows_m4 = (unsigned)(a.g.rows - 4);
  const float hi_u = (float)a.g.cols - 1.0f, hi_v = (float)a.g.rows - 1.0f;  // k + 0.5 < dim - 1

  for (size_t base = (size_t)blockIdx.x * kThreads; base < a.n; base += (size_t)gridDim.x * kThreads) {
    const size_t idx = base + tid;
    if (idx >= a.n) break;  // no barrier below: threads past the end just leave
    const double *pp = a.poses + 3 * idx;
    const double tx = __ldg(pp), ty = __ldg(pp + 1), th = __ldg(pp + 2);
    double s, c;
    sincos(th, &s, &c);
    const WindowRect w = pose_window(a.g, tx, ty, c, s, a.size_x, a.size_y);
    const int W = w.x1 - w.x0 + 1, H = w.y1 - w.y0 + 1;  // <= 0 when empty
    double ku0 = 0.0, kv0 = 0.0;
    {
      const double dx = (double)w.x0 - tx, dy = (double)w.y0 - ty;
      ku0 = (a.g.cx + 0.5) + fma(c, dx, s * dy);  // k + 0.5 at the window origin
      kv0 = (a.g.cy + 0.5) + fma(c, dy, -s * dx);
    }
    // conservative fp32 x-interval of the kernel rectangle per window row y (on k + 0.5):
    //   2 <= ku0 + s y + c x < cols - 1   and   2 <= kv0 + c y - s x < rows - 1, each linear in y
    float l1 = -1e30f, h1 = 1e30f, d1 = 0.f, l2 = -1e30f, h2 = 1e30f, d2 = 0.f;
    {
      const float cf = (float)c, sf = (float)s, ku0f = (float)ku0, kv0f = (float)kv0;
      if (fabsf(cf) > 0.02f) {
        const float inv = rcp_approx(cf);
        const float t0 = (2.0f - ku0f) * inv, t1 = (hi_u - ku0f) * inv;
        l1 = fminf(t0, t1) - 0.02f, h1 = fmaxf(t0, t1) + 0.02f, d1 = -sf * inv;
      }
      if (fabsf(sf) > 0.02f) {
        const float inv = rcp_approx(sf);
        const float t0 = (kv0f - 2.0f) * inv, t1 = (kv0f - hi_v) * inv;
        l2 = fminf(t0, t1) - 0.02f, h2 = fmaxf(t0, t1) + 0.02f, d2 = cf * inv;
      }
    }

    Acc4 acc = {0.0, 0.0, 0.0, 0.0};
    for (int rb = 0; rb < H; rb += 32) {
      const int hc = min(32, H - rb), ys = w.y0 + rb;
      for (int cb = 0; cb < W; cb += 32) {
        const int wc = min(32, W - cb), xs = w.x0 + cb;
        // ---- phase 1: row masks of this chunk --------------------------------------------------------------
        unsigned rowmask = 0u;
        {
          const unsigned sh = (unsigned)xs & 31u;
          const bool need_b = (int)sh + wc > 32;
          const int tb0 = ys >> 3, tb_last = (ys + hc - 1) >> 3;
          const uint32_t *tp = a.tiles + ((size_t)tb0 * a.ntx + (size_t)(xs >> 5)) * 8;
          const float fcb = (float)cb, frb = (float)rb;
          const float l1c = fmaf(frb, d1, l1) - fcb, h1c = fmaf(frb, d1, h1) - fcb;
          const float l2c = fmaf(frb, d2, l2) - fcb, h2c = fmaf(frb, d2, h2) - fcb;
          const float wmax = (float)(wc - 1);
          const int off = ys & 7;  // chunk row i sits at tile row (off + i) & 7 of tile row tb0 + ((off + i) >> 3)
#pragma unroll
          for (int t = 0; t < 5; ++t) {
            if (tb0 + t <= tb_last) {
              const Tile8 A = ldg_tile(tp + (size_t)t * a.ntx * 8);
              Tile8 B;
#pragma unroll
              for (int r = 0; r < 8; ++r) B.w[r] = 0u;
              if (need_b) B = ldg_tile(tp + (size_t)t * a.ntx * 8 + 8);
#pragma unroll
              for (int r = 0; r < 8; ++r) {
                const int i = t * 8 + r - off;
                if ((unsigned)i < (unsigned)hc) {
                  const float fi = (float)i;
                  const float lo = fmaxf(fmaxf(fmaf(fi, d1, l1c), fmaf(fi, d2, l2c)), 0.0f);
                  const float hi = fminf(fminf(fmaf(fi, d1, h1c), fmaf(fi, d2, h2c)), wmax);
                  // ceil / floor through the 1.5 * 2^23 trick: the integer lands in the low mantissa bits
                  const int xlo = __float_as_int(__fadd_ru(fminf(lo, 40.0f), 12582912.0f)) - 0x4B400000;
                  const int xhi = __float_as_int(__fadd_rd(fmaxf(hi, -1.0f), 12582912.0f)) - 0x4B400000;
                  unsigned m = __funnelshift_r(A.w[r], B.w[r], sh);
                  m &= shl_clamp(0xffffffffu, (unsigned)xlo) & shr_clamp(0xffffffffu, (unsigned)(31 - xhi));
                  s_mask[i * kThreads] = m;
                  if (m) rowmask |= 1u << i;
                }
              }
            }
          }
        }
        // ---- phase 2: walk the set bits ----------------------------------------------------------------------
        unsigned mask = 0u;
        int row = 0;
        while (true) {
          if (mask == 0u) {
            if (rowmask == 0u) break;
            row = __ffs((int)rowmask) - 1;
            rowmask &= rowmask - 1u;
            mask = s_mask[row * kThreads];
          }
          const int b = __ffs((int)mask) - 1;
          mask &= mask - 1u;
          const double X = small_int_to_double(cb + b), Y = small_int_to_double(rb + row);
          const double ku = fma(c, X, fma(s, Y, ku0));
          const double kv = fma(c, Y, fma(-s, X, kv0));
          const double M = 6755399441055744.0;  // 1.5 * 2^52: adding it (rounding down) floors to an integer
          const double tu = __dadd_rd(ku, M), tv = __dadd_rd(kv, M);
          const int ux = __double2loint(tu), uy = __double2loint(tv);
          if ((unsigned)(ux - 2) <= cols_m4 && (unsigned)(uy - 2) <= rows_m4) {  // dpose_core.cpp:181-182
            const double rx = ku - (tu - M), ry = kv - (tv - M);                 // c_rel in [0, 1)  (:193)
            const uint32_t addr = my_coef + ((uint32_t)(uy * cols + ux) << patch_shift);
            double a00, a10, a01, a11;
            asm("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(a00), "=d"(a10) : "r"(addr));
            asm("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(a01), "=d"(a11) : "r"(addr + half_off));
            const double fv = fma(a11, rx, a01);
            acc.f += fma(ry, fv, fma(a10, rx, a00));
            if (kJac) {
              const double fu = fma(a11, ry, a10);
              acc.su += fu;
              acc.sv += fv;
              acc.st = fma(fu, kv, acc.st);  // Sum(fu*(kv+.5) - fv*(ku+.5)); center terms folded in below
              acc.st = fma(-fv, ku, acc.st);
            }
          }
        }
      }
    }
    double jx = 0.0, jy = 0.0, jt = 0.0;
    if (kJac) {
      jx = -c * acc.su + s * acc.sv;
      jy = -s * acc.su - c * acc.sv;
      jt = acc.st - (a.g.cy + 0.5) * acc.su + (a.g.cx + 0.5) * acc.sv;
    }
    stg_result(a.out + 4 * idx, acc.f, jx, jy, jt);
  }
}

std::mutex g_tile_mu;  // guards the lazily built per-handle state below

template <bool kJac, int kThreads>
int launch_tile(const TileArgs &a, int grid, size_t smem_bytes, cudaStream_t stream) {
  auto kern = k_cost_batch_tile<kJac, kThreads>;
  DPOSE_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
  kern<<<grid, kThreads, smem_bytes, stream>>>(a);
  DPOSE_LAUNCHED();
  return DPOSE_OK;
}

}  // namespace

int tileplane_build(const dpose_map *map, cudaStream_t stream) {
  if (!map->d_tiles || map->size_x == 0 || map->size_y == 0) return DPOSE_OK;
  const uint32_t quads = (map->tiles_ntx + 3) / 4;
  const size_t n_warps = (size_t)quads * map->tiles_nty;
  const unsigned grid = (unsigned)std::min<size_t>((n_warps + 7) / 8, (size_t)148 * 32);
  k_pack_tiles<<<grid, 256, 0, stream>>>(map->d_data, map->pitch, map->size_y, map->d_tiles, map->tiles_ntx,
                                         map->tiles_nty);
  DPOSE_LAUNCHED();
  return DPOSE_OK;
}

// Returns DPOSE_OK and sets *handled = false if this kernel cannot take the table (shared memory).
int get_cost_batch_tile(const dpose_pg *pg, const dpose_map *map, const double *d_poses, size_t n, double *d_out,
                        int want_jacobian, cudaStream_t stream, bool *handled) {
  *handled = false;
  if (!map->d_tiles || pg->rows < 4 || pg->cols < 4) return DPOSE_OK;
  int sms = 148, smem_optin = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, pg->device);
  cudaDeviceGetAttribute(&smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, pg->device);

  static const int env_threads = []() {
    const char *e = std::getenv("DPOSE_TILE_THREADS");
    return e ? std::atoi(e) : 0;
  }();
  static const int env_rep = []() {
    const char *e = std::getenv("DPOSE_TILE_REPLICAS");
    return e ? std::atoi(e) : 0;
  }();
  int threads = env_threads == 512 || env_threads == 768 || env_threads == 1024 ? env_threads : 768;
  const size_t table_bytes = sizeof(PatchCoef) * (size_t)pg->rows * pg->cols;
  const size_t mask_bytes = (size_t)32 * 4 * threads;
  const size_t budget = (size_t)smem_optin - 64;
  if (table_bytes + mask_bytes > budget) {
    threads = 512;
    if (table_bytes + (size_t)32 * 4 * threads > budget) return DPOSE_OK;  // table too big for shared memory
  }
  int rep_shift = 3;
  if (env_rep == 1 || env_rep == 2 || env_rep == 4 || env_rep == 8) rep_shift = env_rep == 1 ? 0 : env_rep == 2 ? 1 : env_rep == 4 ? 2 : 3;
  while (rep_shift > 0 && (table_bytes << rep_shift) + (size_t)32 * 4 * threads > budget) --rep_shift;

  dpose_pg *p = const_cast<dpose_pg *>(pg);
  dpose_map *m = const_cast<dpose_map *>(map);
  {
    std::lock_guard<std::mutex> lock(g_tile_mu);
    if (!p->d_coef_image || p->coef_image_shift != rep_shift) {
      if (p->d_coef_image) {
        DPOSE_CUDA(cudaStreamSynchronize(stream));
        DPOSE_CUDA(cudaFree(p->d_coef_image));
        p->d_coef_image = nullptr;
      }
      DPOSE_CUDA(cudaMalloc(&p->d_coef_image, table_bytes << rep_shift));
      const int n_halves = pg->rows * pg->cols * 2;
      k_replicate_coef<<<64, 256, 0, stream>>>(reinterpret_cast<const uint4 *>(pg->d_coef),
                                               reinterpret_cast<uint4 *>(p->d_coef_image), n_halves, rep_shift);
      DPOSE_LAUNCHED();
      DPOSE_CUDA(cudaStreamSynchronize(stream));  // other streams may use the image right after
      p->coef_image_shift = rep_shift;
    }
    if (!m->tiles_ready) DPOSE_CUDA(cudaEventCreateWithFlags(&m->tiles_ready, cudaEventDisableTiming));
    if (m->tiles_dirty || m->is_volatile) {
      DPOSE_TRY(tileplane_build(map, stream));
      DPOSE_CUDA(cudaEventRecord(m->tiles_ready, stream));
      m->tiles_dirty = false;
    } else {
      DPOSE_CUDA(cudaStreamWaitEvent(stream, m->tiles_ready, 0));  // built on another stream?
    }
  }

  TileArgs a;
  a.coef_image = reinterpret_cast<const uint4 *>(p->d_coef_image);
  a.coef_bytes = (uint32_t)(table_bytes << rep_shift);
  a.rep_shift = rep_shift;
  a.g = TableGeom{pg->rows, pg->cols, (double)pg->center[0], (double)pg->center[1]};
  a.size_x = (int)map->size_x;
  a.size_y = (int)map->size_y;
  a.tiles = map->d_tiles;
  a.ntx = map->tiles_ntx;
  a.poses = d_poses;
  a.out = d_out;
  a.n = n;
  const size_t smem_bytes = (size_t)a.coef_bytes + (size_t)32 * 4 * threads;
  const size_t want = (n + threads - 1) / threads;
  const int grid = (int)std::min<size_t>(want, (size_t)sms);
  *handled = true;
#define DPOSE_TILE_DISPATCH(T)                                                      \
  return want_jacobian ? launch_tile<true, T>(a, grid, smem_bytes, stream)          \
                       : launch_tile<false, T>(a, grid, smem_bytes, stream)
  if (threads == 512) DPOSE_TILE_DISPATCH(512);
  if (threads == 1024) DPOSE_TILE_DISPATCH(1024);
  DPOSE_TILE_DISPATCH(768);
#undef DPOSE_TILE_DISPATCH
}

}  // namespace dpose
