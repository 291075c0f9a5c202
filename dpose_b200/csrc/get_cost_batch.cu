// get_cost_batch.cu — the batched get_cost(poses[]) kernel for B200 (sm_100a).
//
// Work decomposition (one persistent CTA grid, every warp independent — no CTA-wide barrier
// inside the main loop):
//   * a warp takes a group of 32 poses.  Phase A is lane-parallel (lane = pose): load the pose,
//     sincos, clip the pose's window (AABB of the valid kernel region) to the map and park the
//     per-pose constants (c, s, k at the window origin, window size) in shared memory.
//   * Phase B walks the 32 poses warp-wide.  Lane 0 keeps a ring of TMA tile loads
//     (cp.async.bulk.tensor.2d, one box per pose window, completion on a per-stage mbarrier) a few
//     poses ahead, so window bytes arrive in shared memory without touching registers or L1.
//       - scan   : lane = window row; 16-byte LDS pieces -> byte==254 bit mask, clipped to the
//                  row's x-interval of the rotated kernel rectangle (fp32, conservative)
//       - compact: warp prefix over the popcounts, lanes push (row, x) of their set bits into a
//                  per-warp queue in shared memory
//       - eval   : lanes take queue entries 32 at a time: k = k0 + x*(c,-s) + y*(s,c) (FP64 FMA),
//                  round-to-nearest via the 2^52 magic constant (+ explicit half-away tie fix),
//                  bounds test, one 32-byte patch-coefficient read from the shared-memory copy
//                  of the table, f / f_u / f_v by FMA, per-lane accumulation
//       - reduce : lanes park their 4 partial sums in shared memory; every 4 poses the warp sums
//                  them column-wise in a fixed order (2 lanes per output, 16 adds each + one
//                  __shfl_xor_sync) and writes 4 poses x 4 doubles = 128 contiguous bytes.
// Tensor cores are not used: there is no dense contraction anywhere on this path.
//
// Algorithmic bytes per pose: 24 (pose) + 32 (result) + |window| map bytes.
#include <cuda.h>

#include <algorithm>
#include <mutex>

#include "cost_math.cuh"

namespace dpose {
namespace {

constexpr int kWarps = 8;          // warps per CTA
constexpr int kQueueCap = 512;     // queue entries per warp (u16 each)
constexpr int kPartPitch = 33;     // doubles per (pose, component) row of the partial-sum block
constexpr int kPartPoses = 4;      // poses reduced together
constexpr int kMaxStages = 8;

struct __align__(16) PoseParam {
  double c, s, ku0, kv0;  // rotation and kernel coordinates of the TMA box origin (xa, y0)
  int xa, y0;             // box origin in the map; xa is 16-byte aligned (TMA requirement, see below)
  int xoff_w;             // xoff | w << 16: first window column inside the box, window width
  int h;                  // window height (0 if the window is empty)
};

struct FastArgs {
  const PatchCoef *coef;
  TableGeom g;
  int size_x, size_y;
  const double *poses;
  double *out;
  size_t n;
  int box_w, box_h;      // TMA box (bytes x rows); box_w = 16 * odd where possible, both <= 256
  int stage_bytes;       // box_w * box_h rounded up to 128
  int stages;            // ring depth per warp
  int per_warp_bytes;    // shared memory per warp
  int coef_bytes;        // shared copy of the coefficient table (0: read from global)
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
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *tmap, int x, int y, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
      ::"r"(dst), "l"(tmap), "r"(x), "r"(y), "r"(bar)
      : "memory");
}

// 16 bytes -> bit i set iff byte i == 254
__device__ __forceinline__ unsigned lethal_bits16(const uint4 v) {
  const unsigned k = 0xFEFEFEFEu;
  const unsigned e0 = __vcmpeq4(v.x, k) & 0x01010101u, e1 = __vcmpeq4(v.y, k) & 0x01010101u;
  const unsigned e2 = __vcmpeq4(v.z, k) & 0x01010101u, e3 = __vcmpeq4(v.w, k) & 0x01010101u;
  // gather bits 0/8/16/24 of each word into a nibble (multiply-shift), then pack the 4 nibbles
  const unsigned n0 = (e0 * 0x00204081u) >> 21 & 0xFu, n1 = (e1 * 0x00204081u) >> 21 & 0xFu;
  const unsigned n2 = (e2 * 0x00204081u) >> 21 & 0xFu, n3 = (e3 * 0x00204081u) >> 21 & 0xFu;
  return n0 | (n1 << 4) | (n2 << 8) | (n3 << 12);
}

struct Acc4 {
  double f, su, sv, st;
};

template <bool kJac, bool kSmemCoef>
__device__ __forceinline__ void eval_entry(const PatchCoef *__restrict__ coef, int rows, int cols,
                                           const PoseParam &P, int xr, int yr, Acc4 &acc) {
  const double X = (double)xr, Y = (double)yr;
  const double ku = fma(P.c, X, fma(P.s, Y, P.ku0));
  const double kv = fma(P.c, Y, fma(-P.s, X, P.kv0));
  // round to nearest integer with the 1.5*2^52 constant: low word = integer, tu - M = rint(ku)
  const double M = 6755399441055744.0;
  const double tu = ku + M, tv = kv + M;
  int ux = __double2loint(tu), uy = __double2loint(tv);
  double du = ku - (tu - M), dv = kv - (tv - M);  // in [-0.5, 0.5], exact
  // rint rounds ties to even; std::round (dpose_core.cpp:177) rounds them away from zero
  if (du == 0.5) {
    ux += 1;
    du = -0.5;
  }
  if (dv == 0.5) {
    uy += 1;
    dv = -0.5;
  }
  if ((unsigned)(ux - 2) > (unsigned)(cols - 4) || (unsigned)(uy - 2) > (unsigned)(rows - 4)) return;  // :181-182
  const double rx = du + 0.5, ry = dv + 0.5;  // c_rel (:193)
  PatchCoef p;
  if (kSmemCoef) {
    p = coef[uy * cols + ux];
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
    acc.st = fma(fu, kv, acc.st);  // Sum(fu*kv - fv*ku); the center terms are folded in at the end
    acc.st = fma(-fv, ku, acc.st);
  }
}

template <bool kJac, bool kSmemCoef>
__global__ void __launch_bounds__(kWarps * 32, 2) k_cost_batch_tma(const __grid_constant__ CUtensorMap tmap, FastArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

  // ---- shared memory carve-up -----------------------------------------------------------------
  PatchCoef *s_coef = reinterpret_cast<PatchCoef *>(smem);
  unsigned char *wbase = smem + a.coef_bytes + (size_t)warp * a.per_warp_bytes;
  unsigned char *s_win = wbase;                                               // stages * stage_bytes
  double *s_part = reinterpret_cast<double *>(s_win + a.stages * a.stage_bytes);  // 4*4*33 doubles
  PoseParam *s_param = reinterpret_cast<PoseParam *>(s_part + kPartPoses * 4 * kPartPitch);  // 32
  unsigned short *s_queue = reinterpret_cast<unsigned short *>(s_param + 32);     // kQueueCap
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
  const int rows = a.g.rows, cols = a.g.cols;
  const uint32_t win_u32 = smem_u32(s_win), bar_u32 = smem_u32(s_bar);
  const uint32_t tx_bytes = (uint32_t)(a.box_w * a.box_h);
  const int box_w = a.box_w;
  // conflict-free order of the 16-byte pieces of a row (bank analysis)
  // Bank conflicts of the 16-byte row pieces: lane l reads byte l*box_w + 16*piece, i.e. 16-byte slot
  // (l*m + piece) mod 8 of the 128-byte bank line with m = box_w/16.  For odd m the 8 lanes of a
  // quarter-warp hit 8 different slots: conflict-free, and the host picks odd m whenever it can.
  // Only box_w = 256 keeps an even m; there the piece order is rotated per lane (2-way conflict).
  const int rot_mask = ((box_w >> 4) & 1) ? 0 : 3;
  const int rot = (lane >> 1) & 3;
  uint32_t issue_seq = 0;  // TMA loads issued so far (meaningful in lane 0 only)
  uint32_t cons_seq = 0;   // windows consumed so far (warp-uniform)

  const size_t n_groups = (a.n + 31) / 32;
  const size_t warps_total = (size_t)gridDim.x * kWarps;
  for (size_t group = (size_t)blockIdx.x * kWarps + warp; group < n_groups; group += warps_total) {
    const size_t pose0 = group * 32;
    const int cnt = (int)min((size_t)32, a.n - pose0);

    // ---- phase A: lane = pose --------------------------------------------------------------------
    {
      PoseParam P;
      P.c = 1.0, P.s = 0.0, P.ku0 = 0.0, P.kv0 = 0.0;
      P.xa = P.y0 = P.xoff_w = P.h = 0;
      if (lane < cnt) {
        const double *pp = a.poses + 3 * (pose0 + lane);
        const double tx = pp[0], ty = pp[1], th = pp[2];
        double s, c;
        sincos(th, &s, &c);
        const WindowRect w = pose_window(a.g, tx, ty, c, s, a.size_x, a.size_y);
        P.c = c;
        P.s = s;
        if (w.x1 >= w.x0 && w.y1 >= w.y0) {
          // TMA tile loads need the box's first byte 16-byte aligned in the innermost dimension
          // (measured on B200: a misaligned start raises "illegal instruction"), so the box starts
          // at xa = x0 & ~15 and the window occupies columns [xoff, xoff + w) of it
          const int xa = w.x0 & ~15;
          const double dx = (double)xa - tx, dy = (double)w.y0 - ty;
          P.ku0 = a.g.cx + fma(c, dx, s * dy);
          P.kv0 = a.g.cy + fma(c, dy, -s * dx);
          P.xa = xa;
          P.y0 = w.y0;
          const int xoff = w.x0 - xa;
          const int ww = min(w.x1 - w.x0 + 1, box_w - xoff);
          P.xoff_w = xoff | (ww << 16);
          P.h = min(w.y1 - w.y0 + 1, a.box_h);
        }
      }
      s_param[lane] = P;
    }
    __syncwarp();

    // ---- TMA prologue -------------------------------------------------------------------------------
    if (lane == 0) {
      const int pre = min(a.stages, cnt);
      for (int q = 0; q < pre; ++q) {
        const uint32_t st = issue_seq % a.stages;
        mbar_expect_tx(bar_u32 + 8 * st, tx_bytes);
        tma_load_2d(win_u32 + st * a.stage_bytes, &tmap, s_param[q].xa, s_param[q].y0, bar_u32 + 8 * st);
        ++issue_seq;
      }
    }

    // ---- phase B: one pose at a time, warp-wide --------------------------------------------------------
    for (int q = 0; q < cnt; ++q) {
      const PoseParam P = s_param[q];
      const uint32_t st = cons_seq % a.stages;
      mbar_wait(bar_u32 + 8 * st, (cons_seq / a.stages) & 1);
      ++cons_seq;
      const unsigned char *win = s_win + st * a.stage_bytes;
      const int xoff = P.xoff_w & 0xffff, xend = xoff + (P.xoff_w >> 16);  // window columns [xoff, xend)

      Acc4 acc = {0.0, 0.0, 0.0, 0.0};
      // conservative fp32 description of the kernel rectangle for row clipping
      const float cf = (float)P.c, sf = (float)P.s, ku0f = (float)P.ku0, kv0f = (float)P.kv0;
      const bool use_c = fabsf(cf) > 0.02f, use_s = fabsf(sf) > 0.02f;
      const float inv_c = use_c ? __frcp_rn(cf) : 0.f, inv_ms = use_s ? -__frcp_rn(sf) : 0.f;
      const float hi_u = (float)cols - 1.5f, hi_v = (float)rows - 1.5f;

      for (int rb = 0; rb < P.h; rb += 32) {
        const int row = rb + lane;
        int xlo = 0, xhi = -1;
        if (row < P.h) {
          const float rf = (float)row;
          const float A = fmaf(sf, rf, ku0f), B = fmaf(cf, rf, kv0f);  // k at box column 0 of this row
          float lo = (float)xoff, hi = (float)(xend - 1);
          if (use_c) {  // 1.5 <= A + c x < cols - 1.5
            const float t0 = (1.5f - A) * inv_c, t1 = (hi_u - A) * inv_c;
            lo = fmaxf(lo, fminf(t0, t1) - 0.02f);
            hi = fminf(hi, fmaxf(t0, t1) + 0.02f);
          }
          if (use_s) {  // 1.5 <= B - s x < rows - 1.5
            const float t0 = (1.5f - B) * inv_ms, t1 = (hi_v - B) * inv_ms;
            lo = fmaxf(lo, fminf(t0, t1) - 0.02f);
            hi = fminf(hi, fmaxf(t0, t1) + 0.02f);
          }
          xlo = (int)ceilf(lo);
          xhi = (int)floorf(hi);
        }
        for (int cb = (xoff >> 6) << 6; cb < xend; cb += 64) {
          // -- scan: up to four 16-byte pieces of this lane's row ---------------------------------------
          unsigned long long mask = 0ull;
          const int lo = max(xlo - cb, 0), hi = min(xhi - cb, 63);
          if (lo <= hi) {
            const int np = min(4, (box_w - cb) >> 4);
            const unsigned char *rowp = win + row * box_w + cb;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int piece = rot_mask ? ((i + rot) & rot_mask) : i;
              if (piece < np && piece * 16 <= hi && piece * 16 + 15 >= lo) {
                const uint4 v = *reinterpret_cast<const uint4 *>(rowp + piece * 16);
                mask |= (unsigned long long)lethal_bits16(v) << (piece * 16);
              }
            }
            mask &= (~0ull << lo) & (~0ull >> (63 - hi));
          }
          // -- compact: warp prefix over the counts -------------------------------------------------------
          const int mine = __popcll(mask);
          int inc = mine;
#pragma unroll
          for (int off = 1; off < 32; off <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, inc, off);
            if (lane >= off) inc += t;
          }
          const int total = __shfl_sync(0xffffffffu, inc, 31);
          if (total == 0) continue;
          if (total <= kQueueCap) {
            int dst = inc - mine;
            while (mask) {
              const int b = __ffsll((long long)mask) - 1;
              mask &= mask - 1;
              s_queue[dst++] = (unsigned short)((row << 8) | (cb + b));
            }
            __syncwarp();
            for (int e = lane; e < total; e += 32) {
              const unsigned en = s_queue[e];
              eval_entry<kJac, kSmemCoef>(coef, rows, cols, P, (int)(en & 255u), (int)(en >> 8), acc);
            }
            __syncwarp();
          } else {  // dense maps: each lane walks its own bits (correct, not fast)
            while (mask) {
              const int b = __ffsll((long long)mask) - 1;
              mask &= mask - 1;
              eval_entry<kJac, kSmemCoef>(coef, rows, cols, P, cb + b, row, acc);
            }
          }
        }
      }

      // ---- window consumed: refill this stage -------------------------------------------------------------
      __syncwarp();
      if (lane == 0 && q + a.stages < cnt) {
        const uint32_t st2 = issue_seq % a.stages;
        mbar_expect_tx(bar_u32 + 8 * st2, tx_bytes);
        tma_load_2d(win_u32 + st2 * a.stage_bytes, &tmap, s_param[q + a.stages].xa, s_param[q + a.stages].y0,
                    bar_u32 + 8 * st2);
        ++issue_seq;
      }

      // ---- park the partial sums; every 4 poses reduce them in a fixed order ---------------------------------
      {
        double *pp = s_part + (q & (kPartPoses - 1)) * 4 * kPartPitch + lane;
        pp[0] = acc.f;
        if (kJac) {
          pp[kPartPitch] = acc.su;
          pp[2 * kPartPitch] = acc.sv;
          pp[3 * kPartPitch] = acc.st;
        }
      }
      if ((q & (kPartPoses - 1)) == kPartPoses - 1 || q == cnt - 1) {
        __syncwarp();
        const int qb = q & ~(kPartPoses - 1);  // first pose of this block
        // lane = half * 16 + r: the 16 lanes of a half-warp read 16 different rows of pitch 33
        // doubles -> 16 different bank pairs, no conflict
        const int r = lane & 15, half = lane >> 4;  // r = pose_in_block * 4 + component
        const int pj = r >> 2, comp = r & 3;
        double v = 0.0;
        if (kJac || comp == 0) {
          const double *src = s_part + r * kPartPitch + half * 16;
#pragma unroll
          for (int i = 0; i < 16; ++i) v += src[i];
        }
        v += __shfl_xor_sync(0xffffffffu, v, 16);
        // gather the pose's Su / Sv (components 1, 2) for the rotation into map axes
        const double su = __shfl_sync(0xffffffffu, v, pj * 4 + 1);
        const double sv = __shfl_sync(0xffffffffu, v, pj * 4 + 2);
        if (half == 0 && qb + pj <= q) {
          const PoseParam Q = s_param[qb + pj];
          double o = v;  // comp 0: cost
          if (kJac) {
            if (comp == 1) o = -Q.c * su + Q.s * sv;
            if (comp == 2) o = -Q.s * su - Q.c * sv;
            if (comp == 3) o = v - a.g.cy * su + a.g.cx * sv;
          } else if (comp != 0) {
            o = 0.0;
          }
          a.out[4 * (pose0 + qb + pj) + comp] = o;
        }
        __syncwarp();
      }
    }
    __syncwarp();  // s_param is rewritten by the next group's phase A
  }
}

// ---- host side: tensor-map cache and dispatch ----------------------------------------------------------
struct TmaEntry {
  int box_w, box_h;
  CUtensorMap map;
};
struct TmaCache {
  std::mutex mu;
  std::vector<TmaEntry> entries;
};

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int get_tensor_map(const dpose_map *map, int box_w, int box_h, CUtensorMap *out) {
  static std::mutex init_mu;
  static EncodeTiledFn encode = nullptr;
  {
    std::lock_guard<std::mutex> lock(init_mu);
    if (!encode) {
      void *fn = nullptr;
      cudaDriverEntryPointQueryResult qres;
      DPOSE_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
      if (qres != cudaDriverEntryPointSuccess || !fn)
        return fail(DPOSE_ECUDA, "cuTensorMapEncodeTiled is not available in this driver");
      encode = (EncodeTiledFn)fn;
    }
    dpose_map *m = const_cast<dpose_map *>(map);
    if (!m->tma_cache) m->tma_cache = new TmaCache();
  }
  TmaCache *cache = static_cast<TmaCache *>(map->tma_cache);
  std::lock_guard<std::mutex> lock(cache->mu);
  for (const TmaEntry &e : cache->entries)
    if (e.box_w == box_w && e.box_h == box_h) {
      *out = e.map;
      return DPOSE_OK;
    }
  TmaEntry e;
  e.box_w = box_w;
  e.box_h = box_h;
  // rank-2 uint8 tensor: dim0 = x (size_x valid bytes, pitch-strided rows), dim1 = y; out-of-bounds
  // box elements are zero-filled (0 != 254, i.e. never lethal)
  const cuuint64_t gdim[2] = {map->size_x, map->size_y};
  const cuuint64_t gstride[1] = {map->pitch};
  const cuuint32_t box[2] = {(cuuint32_t)box_w, (cuuint32_t)box_h};
  const cuuint32_t estride[2] = {1, 1};
  const CUresult r = encode(&e.map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, map->d_data, gdim, gstride, box, estride,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(DPOSE_ECUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  cache->entries.push_back(e);
  *out = e.map;
  return DPOSE_OK;
}

template <bool kJac, bool kSmemCoef>
int launch_tma(const CUtensorMap &tmap, const FastArgs &a, int grid, size_t smem_bytes, cudaStream_t stream) {
  auto kern = k_cost_batch_tma<kJac, kSmemCoef>;
  DPOSE_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
  kern<<<grid, kWarps * 32, smem_bytes, stream>>>(tmap, a);
  DPOSE_LAUNCHED();
  return DPOSE_OK;
}

}  // namespace

void map_release_cache(dpose_map *map) {
  delete static_cast<TmaCache *>(map->tma_cache);
  map->tma_cache = nullptr;
}

int get_cost_batch_device(const dpose_pg *pg, const dpose_map *map, const double *d_poses, size_t n,
                          double *d_out, int want_jacobian, cudaStream_t stream) {
  if (n == 0) return DPOSE_OK;
  if (pg->device != map->device) return fail(DPOSE_EINVAL, "pose_gradient and map live on different devices");
  DeviceGuard guard(pg->device);
  if (!guard.ok) return fail(DPOSE_ECUDA, "cudaSetDevice(%d) failed", pg->device);
  const TableGeom g = {pg->rows, pg->cols, (double)pg->center[0], (double)pg->center[1]};

  // box: window (<= win_max cells) + up to 15 bytes of alignment slack, as 16 * m bytes with m odd
  // (conflict-free shared-memory rows) unless that would exceed the 256-element TMA limit
  int box_w = (pg->win_max + 15 + 15) & ~15;
  if (((box_w >> 4) & 1) == 0 && box_w + 16 <= 256) box_w += 16;
  const int box_h = pg->win_max;
  int sms = 148, smem_optin = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, pg->device);
  cudaDeviceGetAttribute(&smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, pg->device);

  bool fast = map->size_x > 0 && map->size_y > 0 && box_w <= 256 && box_h <= 256 && pg->rows >= 4 && pg->cols >= 4;
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
    a.box_w = box_w;
    a.box_h = box_h;
    a.stage_bytes = (box_w * box_h + 127) & ~127;
    const size_t table_bytes = sizeof(PatchCoef) * (size_t)pg->rows * pg->cols;
    a.coef_bytes = table_bytes <= 64 * 1024 ? (int)((table_bytes + 127) & ~(size_t)127) : 0;
    const int fixed = (int)(sizeof(double) * kPartPoses * 4 * kPartPitch + sizeof(PoseParam) * 32 +
                            sizeof(unsigned short) * kQueueCap + 8 * kMaxStages);
    // ring depth: aim for 2 resident CTAs per SM (16 warps) with >= 3 windows in flight per warp;
    // big tables fall back to 1 CTA per SM, and to the generic kernel if even 2 stages do not fit
    auto stages_for = [&](size_t cta_budget) {
      const long long per_warp = ((long long)cta_budget - a.coef_bytes) / kWarps - fixed - 128;
      const long long st = per_warp / a.stage_bytes;
      return (int)(st > kMaxStages ? kMaxStages : st);
    };
    const size_t sm_smem = 227 * 1024;
    int stages = stages_for(sm_smem / 2 - 1024);
    ctas_per_sm = 2;
    if (stages < 3) {
      stages = stages_for(std::min<size_t>((size_t)smem_optin, sm_smem - 1024));
      ctas_per_sm = 1;
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
  CUtensorMap tmap;
  DPOSE_TRY(get_tensor_map(map, box_w, box_h, &tmap));
  const size_t n_groups = (n + 31) / 32;
  const size_t want = (n_groups + kWarps - 1) / kWarps;
  const size_t cap = (size_t)sms * ctas_per_sm;
  const int grid = (int)(want < cap ? want : cap);
  if (want_jacobian)
    return a.coef_bytes ? launch_tma<true, true>(tmap, a, grid, smem_bytes, stream)
                        : launch_tma<true, false>(tmap, a, grid, smem_bytes, stream);
  return a.coef_bytes ? launch_tma<false, true>(tmap, a, grid, smem_bytes, stream)
                      : launch_tma<false, false>(tmap, a, grid, smem_bytes, stream);
}

}  // namespace dpose
