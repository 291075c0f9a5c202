// get_cost_batch.cu — the batched get_cost(poses[]) kernel for B200 (sm_100a).
//
// Data layout: the costmap is read through its 1-bit-per-cell lethal plane (bitplane.cu), rows
// padded to 16 bytes.  A pose's window (axis-aligned bounding box of the valid kernel region,
// clipped to the map) is W <= win_max bits wide and H <= win_max rows high.
//
// Work decomposition (one persistent CTA grid, every warp independent — no CTA-wide barrier inside
// the main loop):
//   * a warp takes a group of 32 poses.  Phase A is lane-parallel (lane = pose): load the pose,
//     sincos, window, and park the per-pose constants in shared memory.
//   * Phase B walks the 32 poses warp-wide.  Lane 0 keeps a ring of TMA tile loads
//     (cp.async.bulk.tensor.2d over the bit plane, one box per pose window, completion on a
//     per-stage mbarrier) a few poses ahead: window bits land in shared memory without touching
//     registers or L1.
//       - scan   : lane = window row; two LDS.32 + a funnel shift give the row's 32 window bits,
//                  clipped to the row's x-interval of the rotated kernel rectangle (fp32,
//                  conservative; exact test happens in eval)
//       - compact: warp prefix over the popcounts; lanes push (row, x) of their set bits into a
//                  per-warp queue in shared memory
//       - eval   : lanes take queue entries 32 at a time: k + 0.5 = k0 + x*(c,-s) + y*(s,c) (FP64
//                  FMA); floor via a round-down add of 1.5*2^52 gives k_upper (std::round for the
//                  positive k that can be valid) and c_rel in one go; bounds test; one 32-byte
//                  patch-coefficient read from the shared-memory copy of the table; f, f_u, f_v by
//                  FMA; per-lane accumulation
//       - reduce : cost, Su, Sv, Stheta are reduced across the warp with __shfl_xor_sync in a fixed,
//                  transposed tree (payload halves at xor 16 and xor 8), so the result is
//                  deterministic and independent of where the pose sits in the batch; 4 lanes write
//                  the pose's 32-byte result (one full sector).
// Tensor cores are not used: there is no dense contraction anywhere on this path.
//
// Algorithmic bytes per pose: 24 (pose) + 32 (result) + |window| map bytes.
#include <cuda.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <mutex>

#include "cost_math.cuh"

namespace dpose {
namespace {

constexpr int kWarps = 8;        // warps per CTA
constexpr int kQueueCap = 256;   // queue entries per warp (u16 each); denser windows take the slow path
constexpr int kMaxStages = 8;
constexpr int kMinCtasPerSm = 4; // 32 resident warps per SM: the kernel is issue-latency bound

struct __align__(16) PoseParam {
  double c, s, ku0, kv0;  // rotation; kernel coordinates (+0.5) of the window origin (x0, y0)
  int bx, y0;             // TMA box origin: byte column in the bit plane (16-byte aligned), row
  int o_w;                // o | w << 16: bit offset of the window inside the box row, window width
  int h;                  // window height (0 if the window is empty)
};

struct FastArgs {
  const PatchCoef *coef;
  TableGeom g;
  int size_x, size_y;
  const double *poses;
  double *out;
  size_t n;
  int box_wb, box_h;    // TMA box: bytes per row (multiple of 16) x rows, both <= 256
  int stage_bytes;      // box_wb * box_h rounded up to 128
  int stages;           // ring depth per warp
  int per_warp_bytes;   // shared memory per warp
  int coef_bytes;       // shared copy of the coefficient table (0: read from global)
};

// ---- PTX wrappers -------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  // bounded: a lost TMA must become a launch failure, never a hung GPU
  for (uint32_t spin = 0; spin < (1u << 24); ++spin)
    if (mbar_try_wait(bar, parity)) return;
  __trap();
}
// The box's first byte must be 16-byte aligned in the innermost dimension (measured on B200: a
// misaligned start raises "illegal instruction"), hence the 16-byte aligned bx.
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *tmap, int x, int y, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
      ::"r"(dst), "l"(tmap), "r"(x), "r"(y), "r"(bar)
      : "memory");
}
__device__ __forceinline__ float rcp_approx(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

struct Acc4 {
  double f, su, sv, st;
};

// One candidate cell at window-relative (xr, yr).  P.ku0 / P.kv0 already contain the +0.5 of
// c_rel = k - round(k) + 0.5 (dpose_core.cpp:193), so floor(k + 0.5) is k_upper (:177, std::round
// for every k that can pass the bounds test) and (k + 0.5) - floor(k + 0.5) is c_rel.
template <bool kJac, bool kSmemCoef>
__device__ __forceinline__ void eval_entry(const PatchCoef *__restrict__ coef, uint32_t coef_u32, int cols,
                                           unsigned cols_m4, unsigned rows_m4, const PoseParam &P, int xr, int yr,
                                           Acc4 &acc) {
  const double X = (double)xr, Y = (double)yr;
  const double ku = fma(P.c, X, fma(P.s, Y, P.ku0));
  const double kv = fma(P.c, Y, fma(-P.s, X, P.kv0));
  const double M = 6755399441055744.0;  // 1.5 * 2^52: adding it (rounding down) floors to an integer
  const double tu = __dadd_rd(ku, M), tv = __dadd_rd(kv, M);
  const int ux = __double2loint(tu), uy = __double2loint(tv);
  if ((unsigned)(ux - 2) <= cols_m4 && (unsigned)(uy - 2) <= rows_m4) {  // :181-182
    const double rx = ku - (tu - M), ry = kv - (tv - M);                 // c_rel in [0, 1)
    PatchCoef p;
    if (kSmemCoef) {  // 32-bit shared address: no generic -> shared conversion per load
      const uint32_t addr = coef_u32 + (uint32_t)(uy * cols + ux) * 32u;
      asm("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(p.a00), "=d"(p.a10) : "r"(addr));
      asm("ld.shared.v2.f64 {%0, %1}, [%2+16];" : "=d"(p.a01), "=d"(p.a11) : "r"(addr));
    } else {
      const double2 *g = reinterpret_cast<const double2 *>(coef + uy * cols + ux);
      const double2 lo = __ldg(g), hi = __ldg(g + 1);
      p.a00 = lo.x, p.a10 = lo.y, p.a01 = hi.x, p.a11 = hi.y;
    }
    const double fv = fma(p.a11, rx, p.a01);
    acc.f += fma(ry, fv, fma(p.a10, rx, p.a00));
    if (kJac) {
      const double fu = fma(p.a11, ry, p.a10);
      acc.su += fu;
      acc.sv += fv;
      acc.st = fma(fu, kv, acc.st);  // Sum(fu*(kv+.5) - fv*(ku+.5)); center terms folded in at the end
      acc.st = fma(-fv, ku, acc.st);
    }
  }
}

// kSmall: every window fits 32 x 32 (no row / column chunk loops)
template <bool kJac, bool kSmemCoef, bool kSmall>
__global__ void __launch_bounds__(kWarps * 32, kMinCtasPerSm) k_cost_batch_tma(const __grid_constant__ CUtensorMap tmap, FastArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

  // ---- shared memory carve-up -----------------------------------------------------------------
  PatchCoef *s_coef = reinterpret_cast<PatchCoef *>(smem);
  unsigned char *wbase = smem + a.coef_bytes + (size_t)warp * a.per_warp_bytes;
  unsigned char *s_win = wbase;                                                            // stages * stage_bytes
  PoseParam *s_param = reinterpret_cast<PoseParam *>(s_win + a.stages * a.stage_bytes);     // 32
  unsigned short *s_queue = reinterpret_cast<unsigned short *>(s_param + 32);               // kQueueCap
  unsigned long long *s_bar = reinterpret_cast<unsigned long long *>(s_queue + kQueueCap);  // stages

  if (kSmemCoef) {
    const int n4 = a.g.rows * a.g.cols * 2;  // uint4 per table
    const uint4 *src = reinterpret_cast<const uint4 *>(a.coef);
    uint4 *dst = reinterpret_cast<uint4 *>(s_coef);
    for (int i = threadIdx.x; i < n4; i += kWarps * 32) dst[i] = __ldg(src + i);
  }
  if (lane == 0) {
    for (int s = 0; s < a.stages; ++s) mbar_init(smem_u32(s_bar + s), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();

  const PatchCoef *coef = kSmemCoef ? s_coef : a.coef;
  const uint32_t coef_u32 = smem_u32(s_coef);
  const int cols = a.g.cols;
  const unsigned cols_m4 = (unsigned)(a.g.cols - 4), rows_m4 = (unsigned)(a.g.rows - 4);
  const float hi_u = (float)a.g.cols - 1.0f, hi_v = (float)a.g.rows - 1.0f;  // k + 0.5 < dim - 1
  const uint32_t win_u32 = smem_u32(s_win), bar_u32 = smem_u32(s_bar);
  const uint32_t tx_bytes = (uint32_t)(a.box_wb * a.box_h);
  const int box_wb = a.box_wb, stages = a.stages, stage_bytes = a.stage_bytes;
  int ist = 0;            // next stage to issue into (lane 0 only)
  int cst = 0, cpar = 0;  // stage / phase parity to consume next (warp-uniform)

  const size_t n_groups = (a.n + 31) / 32;
  const size_t warps_total = (size_t)gridDim.x * kWarps;
  for (size_t group = (size_t)blockIdx.x * kWarps + warp; group < n_groups; group += warps_total) {
    const size_t pose0 = group * 32;
    const int cnt = (int)min((size_t)32, a.n - pose0);

    // ---- phase A: lane = pose --------------------------------------------------------------------
    {
      PoseParam P;
      P.c = 1.0, P.s = 0.0, P.ku0 = 0.0, P.kv0 = 0.0;
      P.bx = P.y0 = P.o_w = P.h = 0;
      if (lane < cnt) {
        const double *pp = a.poses + 3 * (pose0 + lane);
        const double tx = pp[0], ty = pp[1], th = pp[2];
        double s, c;
        sincos(th, &s, &c);
        const WindowRect w = pose_window(a.g, tx, ty, c, s, a.size_x, a.size_y);
        P.c = c;
        P.s = s;
        if (w.x1 >= w.x0 && w.y1 >= w.y0) {
          const double dx = (double)w.x0 - tx, dy = (double)w.y0 - ty;
          P.ku0 = (a.g.cx + 0.5) + fma(c, dx, s * dy);
          P.kv0 = (a.g.cy + 0.5) + fma(c, dy, -s * dx);
          const int xa = w.x0 & ~127;  // first bit of the 16-byte aligned box
          P.bx = xa >> 3;
          P.y0 = w.y0;
          const int o = w.x0 - xa;
          const int ww = min(min(w.x1 - w.x0 + 1, box_wb * 8 - 32 - o), kSmall ? 32 : 256);
          P.o_w = o | (ww << 16);
          P.h = min(w.y1 - w.y0 + 1, kSmall ? 32 : a.box_h);
        }
      }
      s_param[lane] = P;
    }
    __syncwarp();

    // ---- TMA prologue -------------------------------------------------------------------------------
    if (lane == 0) {
      const int pre = min(stages, cnt);
      for (int q = 0; q < pre; ++q) {
        mbar_expect_tx(bar_u32 + 8 * ist, tx_bytes);
        tma_load_2d(win_u32 + ist * stage_bytes, &tmap, s_param[q].bx, s_param[q].y0, bar_u32 + 8 * ist);
        if (++ist == stages) ist = 0;
      }
    }

    // ---- phase B: one pose at a time, warp-wide --------------------------------------------------------
    for (int q = 0; q < cnt; ++q) {
      const PoseParam P = s_param[q];
      mbar_wait(bar_u32 + 8 * cst, cpar);
      const unsigned char *win = s_win + cst * stage_bytes;
      if (++cst == stages) {
        cst = 0;
        cpar ^= 1;
      }
      const int o = P.o_w & 0xffff, W = P.o_w >> 16;

      Acc4 acc = {0.0, 0.0, 0.0, 0.0};
      // conservative fp32 description of the kernel rectangle for row clipping (on k + 0.5)
      const float cf = (float)P.c, sf = (float)P.s, ku0f = (float)P.ku0, kv0f = (float)P.kv0;
      const bool use_c = fabsf(cf) > 0.02f, use_s = fabsf(sf) > 0.02f;
      const float inv_c = use_c ? rcp_approx(cf) : 0.f, inv_ms = use_s ? -rcp_approx(sf) : 0.f;

      for (int rb = 0; rb < (kSmall ? 1 : P.h); rb += 32) {
        const int row = rb + lane;
        int xlo = 0, xhi = -1;
        if (row < P.h) {
          const float rf = (float)row;
          const float A = fmaf(sf, rf, ku0f), B = fmaf(cf, rf, kv0f);  // k + 0.5 at window column 0
          float lo = 0.f, hi = (float)(W - 1);
          if (use_c) {  // 2 <= A + c x < cols - 1
            const float t0 = (2.0f - A) * inv_c, t1 = (hi_u - A) * inv_c;
            lo = fmaxf(lo, fminf(t0, t1) - 0.02f);
            hi = fminf(hi, fmaxf(t0, t1) + 0.02f);
          }
          if (use_s) {  // 2 <= B - s x < rows - 1
            const float t0 = (2.0f - B) * inv_ms, t1 = (hi_v - B) * inv_ms;
            lo = fmaxf(lo, fminf(t0, t1) - 0.02f);
            hi = fminf(hi, fmaxf(t0, t1) + 0.02f);
          }
          xlo = (int)ceilf(lo);
          xhi = (int)floorf(hi);
        }
        const unsigned char *rowp = win + row * box_wb;
        for (int cb = 0; cb < (kSmall ? 1 : W); cb += 32) {
          // -- scan: the 32 window bits [cb, cb + 32) of this lane's row -----------------------------------
          // one LDS.128 of the 16-byte piece holding the field (conflict-free across the quarter-warp
          // for row pitches of 16 * odd bytes, 2-way for 32) instead of two same-column LDS.32 (8-way)
          unsigned mask = 0u;
          const int lo = max(xlo - cb, 0), hi = min(xhi - cb, 31);
          if (lo <= hi) {
            const int bo = o + cb, wsel = (bo >> 5) & 3;  // warp-uniform
            const uint4 v = *reinterpret_cast<const uint4 *>(rowp + ((bo >> 7) << 4));
            unsigned nx = 0u;  // first word of the next piece, only needed when the field straddles it
            if (wsel == 3 && (bo & 31) + hi >= 32) nx = *reinterpret_cast<const unsigned *>(rowp + ((bo >> 7) << 4) + 16);
            const unsigned a0 = (wsel & 2) ? v.z : v.x, a1 = (wsel & 2) ? v.w : v.y, a2 = (wsel & 2) ? nx : v.z;
            const unsigned w0 = (wsel & 1) ? a1 : a0, w1 = (wsel & 1) ? a2 : a1;
            mask = __funnelshift_r(w0, w1, bo & 31);
            mask &= (0xffffffffu << lo) & (0xffffffffu >> (31 - hi));
          }
          // -- compact: warp prefix over the counts -------------------------------------------------------
          const int mine = __popc(mask);
          int inc = mine;
#pragma unroll
          for (int off = 1; off < 32; off <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, inc, off);
            if (lane >= off) inc += t;
          }
          const int total = __shfl_sync(0xffffffffu, inc, 31);
          if (total == 0) continue;
          if (total <= kQueueCap) {
            unsigned short *dst = s_queue + (inc - mine);
            const unsigned base = (unsigned)((row << 8) | cb);
#pragma unroll 1
            while (mask) {
              const int b = __ffs((int)mask) - 1;
              mask &= mask - 1;
              *dst++ = (unsigned short)(base + b);
            }
            __syncwarp();
#pragma unroll 1
            for (int e = lane; e < total; e += 32) {
              const unsigned en = s_queue[e];
              eval_entry<kJac, kSmemCoef>(coef, coef_u32, cols, cols_m4, rows_m4, P, (int)(en & 255u), (int)(en >> 8),
                                          acc);
            }
            __syncwarp();
          } else {  // dense maps: each lane walks its own bits (correct, not fast)
            while (mask) {
              const int b = __ffs((int)mask) - 1;
              mask &= mask - 1;
              eval_entry<kJac, kSmemCoef>(coef, coef_u32, cols, cols_m4, rows_m4, P, cb + b, row, acc);
            }
          }
        }
      }

      // ---- window consumed: refill this stage -------------------------------------------------------------
      __syncwarp();
      if (lane == 0 && q + stages < cnt) {
        mbar_expect_tx(bar_u32 + 8 * ist, tx_bytes);
        tma_load_2d(win_u32 + ist * stage_bytes, &tmap, s_param[q + stages].bx, s_param[q + stages].y0,
                    bar_u32 + 8 * ist);
        if (++ist == stages) ist = 0;
      }

      // ---- reduce the 4 per-lane sums with a fixed xor-shuffle tree, transposed so that every level
      //      halves the payload: after xor 16 a lane carries 2 of the 4 sums, after xor 8 one; lanes
      //      0-7 end with cost, 8-15 with Su, 16-23 with Sv, 24-31 with Stheta -----------------------------
      if (kJac) {
        const bool b4 = lane & 16, b3 = lane & 8;
        double k0 = b4 ? acc.sv : acc.f, k1 = b4 ? acc.st : acc.su;  // kept
        const double g0 = b4 ? acc.f : acc.sv, g1 = b4 ? acc.su : acc.st;  // given to lane ^ 16
        k0 += __shfl_xor_sync(0xffffffffu, g0, 16);
        k1 += __shfl_xor_sync(0xffffffffu, g1, 16);
        double k = b3 ? k1 : k0;
        k += __shfl_xor_sync(0xffffffffu, b3 ? k0 : k1, 8);
        k += __shfl_xor_sync(0xffffffffu, k, 4);
        k += __shfl_xor_sync(0xffffffffu, k, 2);
        k += __shfl_xor_sync(0xffffffffu, k, 1);
        const double su = __shfl_sync(0xffffffffu, k, 8), sv = __shfl_sync(0xffffffffu, k, 16);
        if ((lane & 7) == 0) {
          const int comp = lane >> 3;
          double o_val = k;  // comp 0: cost
          if (comp == 1) o_val = -P.c * su + P.s * sv;
          if (comp == 2) o_val = -P.s * su - P.c * sv;
          if (comp == 3) o_val = k - (a.g.cy + 0.5) * su + (a.g.cx + 0.5) * sv;
          a.out[4 * (pose0 + q) + comp] = o_val;
        }
      } else {
        double k = acc.f;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) k += __shfl_xor_sync(0xffffffffu, k, off);
        if (lane < 4) a.out[4 * (pose0 + q) + lane] = lane == 0 ? k : 0.0;
      }
    }
    __syncwarp();  // s_param is rewritten by the next group's phase A
  }
}

// ---- host side: tensor-map cache and dispatch ----------------------------------------------------------
struct TmaEntry {
  int box_wb, box_h;
  CUtensorMap map;
};
struct TmaCache {
  std::mutex mu;
  std::vector<TmaEntry> entries;
};

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

std::mutex g_map_mu;  // guards lazy per-map state (tensor maps, bit plane rebuild flag)

int get_tensor_map(const dpose_map *map, int box_wb, int box_h, CUtensorMap *out) {
  static EncodeTiledFn encode = nullptr;
  dpose_map *m = const_cast<dpose_map *>(map);
  {
    std::lock_guard<std::mutex> lock(g_map_mu);
    if (!encode) {
      void *fn = nullptr;
      cudaDriverEntryPointQueryResult qres;
      DPOSE_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
      if (qres != cudaDriverEntryPointSuccess || !fn)
        return fail(DPOSE_ECUDA, "cuTensorMapEncodeTiled is not available in this driver");
      encode = (EncodeTiledFn)fn;
    }
    if (!m->tma_cache) m->tma_cache = new TmaCache();
  }
  TmaCache *cache = static_cast<TmaCache *>(m->tma_cache);
  std::lock_guard<std::mutex> lock(cache->mu);
  for (const TmaEntry &e : cache->entries)
    if (e.box_wb == box_wb && e.box_h == box_h) {
      *out = e.map;
      return DPOSE_OK;
    }
  TmaEntry e;
  e.box_wb = box_wb;
  e.box_h = box_h;
  // rank-2 uint8 tensor over the bit plane: dim0 = byte column, dim1 = row; box elements outside the
  // tensor are zero-filled, i.e. "not lethal" — this is what clips windows at the map border
  const cuuint64_t gdim[2] = {map->bits_pitch, map->size_y};
  const cuuint64_t gstride[1] = {map->bits_pitch};
  const cuuint32_t box[2] = {(cuuint32_t)box_wb, (cuuint32_t)box_h};
  const cuuint32_t estride[2] = {1, 1};
  const CUresult r = encode(&e.map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, map->d_bits, gdim, gstride, box, estride,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(DPOSE_ECUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  cache->entries.push_back(e);
  *out = e.map;
  return DPOSE_OK;
}

template <bool kJac, bool kSmemCoef, bool kSmall>
int launch_tma(const CUtensorMap &tmap, const FastArgs &a, int grid, size_t smem_bytes, cudaStream_t stream) {
  auto kern = k_cost_batch_tma<kJac, kSmemCoef, kSmall>;
  DPOSE_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
  kern<<<grid, kWarps * 32, smem_bytes, stream>>>(tmap, a);
  DPOSE_LAUNCHED();
  return DPOSE_OK;
}

template <bool kJac>
int launch_tma_jac(const CUtensorMap &tmap, const FastArgs &a, bool small, int grid, size_t smem_bytes,
                   cudaStream_t stream) {
  if (a.coef_bytes)
    return small ? launch_tma<kJac, true, true>(tmap, a, grid, smem_bytes, stream)
                 : launch_tma<kJac, true, false>(tmap, a, grid, smem_bytes, stream);
  return small ? launch_tma<kJac, false, true>(tmap, a, grid, smem_bytes, stream)
               : launch_tma<kJac, false, false>(tmap, a, grid, smem_bytes, stream);
}

}  // namespace

void map_release_cache(dpose_map *map) {
  delete static_cast<TmaCache *>(map->tma_cache);
  map->tma_cache = nullptr;
}

int get_cost_batch_device(const dpose_pg *pg, const dpose_map *map, const double *d_poses, size_t n,
                          double *d_out, int want_jacobian, cudaStream_t stream, bool hold_planes) {
  if (n == 0) return DPOSE_OK;
  if (pg->device != map->device) return fail(DPOSE_EINVAL, "pose_gradient and map live on different devices");
  DeviceGuard guard(pg->device);
  if (!guard.ok) return fail(DPOSE_ECUDA, "cudaSetDevice(%d) failed", pg->device);
  const TableGeom g = {pg->rows, pg->cols, (double)pg->center[0], (double)pg->center[1]};

  // kernel choice: the thread-per-pose tile kernel unless DPOSE_BATCH_KERNEL asks for the warp-per-pose
  // TMA kernel ("tma") or the generic window scan ("generic") — kept for A/B measurements and as the
  // fallbacks for tables that do not fit the tile kernel's shared-memory layout
  static const int forced = []() {
    const char *e = std::getenv("DPOSE_BATCH_KERNEL");
    if (!e) return 0;
    return !std::strcmp(e, "tma") ? 1 : !std::strcmp(e, "generic") ? 2 : 0;
  }();
  if (forced == 0 && map->size_x > 0 && map->size_y > 0) {
    bool handled = false;
    DPOSE_TRY(get_cost_batch_tile(pg, map, d_poses, n, d_out, want_jacobian, stream, hold_planes, &handled));
    if (handled) return DPOSE_OK;
  }

  // TMA box over the bit plane: the window's W <= win_max bits start at bit offset o in [0, 128) of
  // the 16-byte aligned box row; one spare word keeps the funnel shift's second load inside the row
  const int win = pg->win_max;
  const int box_wb = ((127 + win + 32 + 7) / 8 + 15) & ~15;
  const int box_h = win;
  int sms = 148, smem_optin = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, pg->device);
  cudaDeviceGetAttribute(&smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, pg->device);

  bool fast = forced != 2 && map->size_x > 0 && map->size_y > 0 && win <= 256 && box_wb <= 256 && pg->rows >= 4 && pg->cols >= 4;
  FastArgs a;
  size_t smem_bytes = 0;
  int ctas_per_sm = 2;
  if (fast) {
    a.coef = pg->d_coef;
    a.g = g;
    a.size_x = (int)map->size_x;
    a.size_y = (int)map->size_y;
    a.poses = d_poses;
    a.out = d_out;
    a.n = n;
    a.box_wb = box_wb;
    a.box_h = box_h;
    a.stage_bytes = (box_wb * box_h + 127) & ~127;
    const size_t table_bytes = sizeof(PatchCoef) * (size_t)pg->rows * pg->cols;
    a.coef_bytes = table_bytes <= 64 * 1024 ? (int)((table_bytes + 127) & ~(size_t)127) : 0;
    const int fixed = (int)(sizeof(PoseParam) * 32 + sizeof(unsigned short) * kQueueCap + 8 * kMaxStages);
    // ring depth: aim for kMinCtasPerSm resident CTAs per SM with >= 3 windows in flight per warp;
    // bigger tables get fewer CTAs per SM, and the generic kernel if even 2 stages do not fit
    auto stages_for = [&](size_t cta_budget) {
      const long long per_warp = ((long long)cta_budget - a.coef_bytes) / kWarps - fixed - 128;
      const long long st = per_warp / a.stage_bytes;
      return (int)(st > kMaxStages ? kMaxStages : st);
    };
    const size_t sm_smem = 227 * 1024;
    int stages = 0;
    for (ctas_per_sm = kMinCtasPerSm; ctas_per_sm >= 1; --ctas_per_sm) {
      const size_t budget = std::min<size_t>((size_t)smem_optin, sm_smem / ctas_per_sm - 1024);
      stages = stages_for(budget);
      if (stages >= 3 || ctas_per_sm == 1) break;
    }
    if (stages < 2) fast = false;
    a.stages = stages;
    a.per_warp_bytes = (stages * a.stage_bytes + fixed + 127) & ~127;
    smem_bytes = (size_t)a.coef_bytes + (size_t)kWarps * a.per_warp_bytes;
  }
  if (!fast) {
    BatchArgs b;
    b.coef = pg->d_coef;
    b.g = g;
    b.map = map->d_data;
    b.pitch = map->pitch;
    b.size_x = (int)map->size_x;
    b.size_y = (int)map->size_y;
    b.poses = d_poses;
    b.out = d_out;
    b.n = n;
    return launch_batch_generic(b, want_jacobian, pg->device, stream);
  }
  // (re)build the lethal bit plane in front of the batch kernel when the map bytes changed
  {
    dpose_map *m = const_cast<dpose_map *>(map);
    std::lock_guard<std::mutex> lock(g_map_mu);
    if (!m->bits_ready) DPOSE_CUDA(cudaEventCreateWithFlags(&m->bits_ready, cudaEventDisableTiming));
    if (m->bits_dirty || (m->is_volatile && !hold_planes)) {
      DPOSE_TRY(bitplane_build(map, stream));
      DPOSE_CUDA(cudaEventRecord(m->bits_ready, stream));
      m->bits_dirty = false;
    } else {
      DPOSE_CUDA(cudaStreamWaitEvent(stream, m->bits_ready, 0));  // built on another stream?
    }
  }
  CUtensorMap tmap;
  DPOSE_TRY(get_tensor_map(map, box_wb, box_h, &tmap));
  const size_t n_groups = (n + 31) / 32;
  const size_t want = (n_groups + kWarps - 1) / kWarps;
  const size_t cap = (size_t)sms * ctas_per_sm;
  const int grid = (int)(want < cap ? want : cap);
  const bool small = win <= 32;
  return want_jacobian ? launch_tma_jac<true>(tmap, a, small, grid, smem_bytes, stream)
                       : launch_tma_jac<false>(tmap, a, small, grid, smem_bytes, stream);
}

}  // namespace dpose
