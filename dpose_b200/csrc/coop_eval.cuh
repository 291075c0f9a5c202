// coop_eval.cuh — pose_gradient::get_cost (dpose_core.cpp:200-260) of ONE pose against the lethal tile plane by G
// cooperating lanes of a warp (G = 2 .. 32, a power of two).
//
// The thread-per-pose kernel (get_cost_batch_tile.cu) is built for throughput: a pose is one thread's serial work
// (~2600 instructions), which is the right shape when 150 000+ poses are in flight.  A multi-start solve or a
// solver-iteration-sized batch has a few thousand poses: there the GPU is empty and what counts is the LATENCY of one
// evaluation.  Here the rows of the pose's window are dealt round-robin to the G lanes of a group:
//   phase 1  lane g takes the row slots j = g, g + G, ...: one 32-bit load of the tile word (8 lanes share a 32-byte
//            sector), funnel shift, clip to the row's x-interval of the rotated kernel rectangle — in FP64 here (each
//            lane has 2 .. 10 slots, the interval arithmetic is not the bottleneck), so there is no table-size limit
//   phase 2  every lane walks the set bits of ITS rows (same arithmetic as the tile kernel: absolute-coordinate patch
//            polynomial, no bounds test thanks to the clip invariant), coefficients gathered from global memory / L1 as
//            one 32-byte record per patch (no shared-memory image to stage: a latency-bound launch cannot afford it)
//   reduce   cost, Sum f_u, Sum f_v, Sum theta terms by __shfl_xor_sync over the G lanes in a fixed order: every lane of the group
//            ends up with the same four sums, bit for bit, independent of the batch position
// Tensor cores are not used: there is no dense contraction on this path.
//
// Every floating-point operation that could be contracted is written as an explicit fma / __dmul_rn, so the function
// gives the same bits under -fmad=true and -fmad=false (solve.cu is compiled without contraction, see build.py).
#pragma once
#include "cost_math.cuh"

#ifndef DPOSE_TILE_CHECK
#define DPOSE_TILE_CHECK 0  // 1: debug build that traps on any gather outside the coefficient records
#endif

namespace dpose {

struct CoopCtx {
  TableGeom g;
  int size_x, size_y;     // size_y: end of the band of map rows that count (BatchSpec::row_hi; the map's height otherwise)
  int y_lo = 0;           // its start
  const uint32_t *tiles;  // lethal tile plane (common.cuh)
  uint32_t ntx, nty;
  const double *coef_abs;  // 4 doubles per patch: b00, b10, b01, b11 (coef_abs_ensure)
  double clip_thr;         // a strip of the kernel rectangle is clipped per row only if |cos| (|sin|) exceeds this
};

// get_cost_batch_coop.cu: the context of a (pose_gradient, map) pair; coef_abs_ensure / tileplane_ensure come first
CoopCtx coop_ctx(const dpose_pg *pg, const dpose_map *map);

// slots a 32-row chunk can touch: its rows start at slot (y & 7) of the first tile row
constexpr int kCoopSlots = 32 + 7;
template <int G>
struct CoopK {
  static constexpr int value = (kCoopSlots + G - 1) / G;  // row slots per lane
};

namespace coop_detail {
__device__ __forceinline__ unsigned shl_clamp(unsigned v, unsigned n) {  // PTX shifts clamp the amount to 32
  unsigned r;
  asm("shl.b32 %0, %1, %2;" : "=r"(r) : "r"(v), "r"(n));
  return r;
}
__device__ __forceinline__ unsigned shr_clamp(unsigned v, unsigned n) {
  unsigned r;
  asm("shr.b32 %0, %1, %2;" : "=r"(r) : "r"(v), "r"(n));
  return r;
}
__device__ __forceinline__ void ldg_patch(const double *p, double &b00, double &b10, double &b01, double &b11) {
  unsigned w0, w1, w2, w3, w4, w5, w6, w7;  // one 32-byte sector, read-only path (LDG.E.256, new on sm_100)
  asm volatile("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(w0), "=r"(w1), "=r"(w2), "=r"(w3), "=r"(w4), "=r"(w5), "=r"(w6), "=r"(w7)
               : "l"(p));
  b00 = __hiloint2double((int)w1, (int)w0);
  b10 = __hiloint2double((int)w3, (int)w2);
  b01 = __hiloint2double((int)w5, (int)w4);
  b11 = __hiloint2double((int)w7, (int)w6);
}
}  // namespace coop_detail

// out: cost, J_x, J_y, J_theta (the last three 0 unless kJac).  g: this lane's index in its group; gmask: the lanes of
// the group (all of them must call); s_mask: this thread's column of a [CoopK<G>::value][stride] block of shared words.
template <int G, bool kJac>
__device__ __forceinline__ void coop_eval(const CoopCtx &a, double tx, double ty, double th, int g, unsigned gmask,
                                          uint32_t *s_mask, int stride, double out[4]) {
  using namespace coop_detail;
  constexpr int K = CoopK<G>::value;
  double s, c;
  sincos(th, &s, &c);
  const WindowRect w = pose_window(a.g, tx, ty, c, s, a.size_x, a.size_y, a.y_lo);
  const int W = w.x1 - w.x0 + 1, H = w.y1 - w.y0 + 1;  // <= 0 when empty (also for NaN poses)
  const int cols = a.g.cols;
  const double cx15 = a.g.cx - 1.5, cy15 = a.g.cy - 1.5;
  // Kernel coordinates are carried as k' = k - 1.5: floor(k') = k_upper - 2 (k_upper = round(k), dpose_core.cpp:177),
  // and the bounds test (:181-182) becomes 0 <= floor(k') <= dim - 4, i.e. 0 <= k' < dim - 3.
  const double hi_u = (double)a.g.cols - 3.0, hi_v = (double)a.g.rows - 3.0;
  double ku0, kv0;  // k' at the window origin
  {
    const double dx = (double)w.x0 - tx, dy = (double)w.y0 - ty;
    ku0 = cx15 + fma(c, dx, __dmul_rn(s, dy));
    kv0 = cy15 + fma(c, dy, -__dmul_rn(s, dx));
  }
  // x-interval of the kernel rectangle per window row y: 0 <= ku0 + s y + c x < cols - 3 and
  // 0 <= kv0 + c y - s x < rows - 3, each linear in y.  CLIP INVARIANT (what lets the walk drop the bounds test):
  // every cell that survives has k' within one patch of the valid range, so floor(k') is in [-1, dim - 3] and
  // addresses a record of coef_abs (zeros outside the valid range).  A clipped strip overshoots by 1e-6 |c| cell; a
  // strip left unclipped because |c| <= clip_thr = 0.5 / win_max varies by at most 0.5 cell along a window row, and
  // every window row crosses the valid rectangle (the window is its bounding box).
  const double kMargin = 1e-6;
  double l1 = -1e9, h1 = 1e9, d1 = 0.0, l2 = -1e9, h2 = 1e9, d2 = 0.0;
  if (fabs(c) > a.clip_thr) {
    const double inv = 1.0 / c;
    const double t0 = __dmul_rn(-ku0, inv), t1 = __dmul_rn(hi_u - ku0, inv);
    l1 = fmin(t0, t1) - kMargin, h1 = fmax(t0, t1) + kMargin, d1 = __dmul_rn(-s, inv);
  }
  if (fabs(s) > a.clip_thr) {
    const double inv = 1.0 / s;
    const double t0 = __dmul_rn(kv0, inv), t1 = __dmul_rn(kv0 - hi_v, inv);
    l2 = fmin(t0, t1) - kMargin, h2 = fmax(t0, t1) + kMargin, d2 = __dmul_rn(c, inv);
  }

  double f = 0.0, su = 0.0, sv = 0.0, st = 0.0;
  const double cw = __dmul_rn(32.0, c), sw = __dmul_rn(32.0, s);  // exact
  int zero_x = (int)(a.ntx >> 31), zero_y = (int)(a.nty >> 31);   // 0, but not to the compiler (see unit_plus_32nd)
  asm volatile("" : "+r"(zero_x), "+r"(zero_y));
  for (int rb = 0; rb < H; rb += 32) {
    const int hc = min(32, H - rb), ys = w.y0 + rb;
    for (int cb = 0; cb < W; cb += 32) {
      const int wc = min(32, W - cb), xs = w.x0 + cb;
      const int off = ys & 7;  // chunk row i sits in row slot i + off of the chunk's tile rows
      // ---- phase 1: this lane's row masks ----------------------------------------------------------------------------
      unsigned local = 0u;  // non-empty local slots
      {
        const unsigned sh = (unsigned)xs & 31u;
        const bool need_b = (int)sh + wc > 32;
        const uint32_t *tp = a.tiles + ((size_t)(ys >> 3) * a.ntx + (size_t)(xs >> 5)) * 8;
        const size_t tstride = (size_t)a.ntx * 8;
        const double dcb = (double)cb, wmax = (double)(wc - 1);
#pragma unroll
        for (int k = 0; k < K; ++k) {
          const int j = g + k * G;
          unsigned m = 0u;
          if (j >= off && j < off + hc) {  // (j >> 3) <= (off + hc - 1) >> 3: inside the window's tile rows
            const uint32_t *wp = tp + (size_t)(j >> 3) * tstride + (j & 7);
            const unsigned A = __ldg(wp), B = need_b ? __ldg(wp + 8) : 0u;
            const double yrel = (double)(rb + j - off);
            const double lo = fmax(fmax(fma(yrel, d1, l1), fma(yrel, d2, l2)) - dcb, 0.0);
            const double hi = fmin(fmin(fma(yrel, d1, h1), fma(yrel, d2, h2)) - dcb, wmax);
            const int xlo = __double2int_ru(lo), xhi = __double2int_rd(hi);  // xlo >= 0, xhi <= 31; empty if xlo > xhi
            m = __funnelshift_r(A, B, sh) & shl_clamp(0xffffffffu, (unsigned)xlo) &
                shr_clamp(0xffffffffu, (unsigned)(31 - xhi));
          }
          s_mask[k * stride] = m;
          if (m) local |= 1u << k;
        }
      }
      // ---- phase 2: walk the set bits of this lane's rows, highest slot / highest bit first ------------------------------
      // the bit / row index enters as X = 1 + b / 32, Y = 1 + row / 32 (unit_plus_32nd, cost_math.cuh: no int -> double
      // conversion): c b + s row = cw X + sw Y - cw - sw, the constants ride on the chunk origin (cb / 32 - 1 and
      // rb / 32 - 1 are exact integers)
      const double dcb = (double)((cb >> 5) - 1), drb = (double)((rb >> 5) - 1);
      const double kuc = fma(cw, dcb, fma(sw, drb, ku0)), kvc = fma(cw, drb, fma(-sw, dcb, kv0));
      unsigned mask = 0u;
      int row = 0;
      while (true) {
        if ((mask | local) == 0u) break;
        if (mask == 0u) {
          const int k = top_bit(local);
          local = clear_top(local, k);
          mask = s_mask[k * stride];
          row = g + k * G - off;
        }
        const int b = top_bit(mask);
        mask = clear_top(mask, b);
        const double X = unit_plus_32nd(b, zero_x), Y = unit_plus_32nd(row, zero_y);
        const double ku = fma(cw, X, fma(sw, Y, kuc));
        const double kv = fma(cw, Y, fma(-sw, X, kvc));
        const double M = 6755399441055744.0;  // 1.5 * 2^52: adding it (rounding down) floors to an integer
        const int ux = __double2loint(__dadd_rd(ku, M)), uy = __double2loint(__dadd_rd(kv, M));  // k_upper - 2
        // no bounds test (dpose_core.cpp:181-182): records outside the valid range are zero, see the clip invariant
        const int pidx = (uy + 2) * cols + ux + 2;
#if DPOSE_TILE_CHECK
        if (pidx < 0 || pidx >= a.g.rows * cols || ux < -2 || ux > cols - 3 || uy < -2 || uy > a.g.rows - 3) __trap();
#endif
        double b00, b10, b01, b11;
        ldg_patch(a.coef_abs + 4 * (size_t)pidx, b00, b10, b01, b11);
        const double fv = fma(b11, ku, b01);  // df/dkv
        f += fma(kv, fv, fma(b10, ku, b00));
        if (kJac) {
          const double fu = fma(b11, kv, b10);  // df/dku
          su += fu;
          sv += fv;
          st = fma(fu, kv, st);  // Sum(fu * kv' - fv * ku'); center terms folded in below
          st = fma(-fv, ku, st);
        }
      }
    }
  }
  // fixed-order butterfly over the group's lanes: every lane ends with the same bits
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) {
    f += __shfl_xor_sync(gmask, f, o);
    if (kJac) {
      su += __shfl_xor_sync(gmask, su, o);
      sv += __shfl_xor_sync(gmask, sv, o);
      st += __shfl_xor_sync(gmask, st, o);
    }
  }
  out[0] = f;
  out[1] = out[2] = out[3] = 0.0;
  if (kJac) {
    out[1] = fma(s, sv, -__dmul_rn(c, su));
    out[2] = fma(-c, sv, -__dmul_rn(s, su));
    out[3] = fma(cx15, sv, fma(-cy15, su, st));
  }
}

}  // namespace dpose
