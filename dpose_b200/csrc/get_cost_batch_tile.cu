// get_cost_batch_tile.cu — batched get_cost(poses[]) for B200 (sm_100a), one THREAD per pose.
//
// Why this shape.  A warp-per-pose kernel (round 1's first design) spent ~355 warp instructions and 78
// shared-memory wavefronts per pose (ncu, profiles/r1_v2_*): a pose has only ~31 valid lethal cells, so
// warp-wide scans, prefix sums, compaction queues and shuffle reductions cost more than the
// arithmetic.  Here every lane owns a pose: nothing is exchanged between lanes, there is no barrier
// in the main loop and no reduction at all — a pose's sum is accumulated by one thread in a fixed
// (row-major) order, so the result is deterministic and independent of the batch position.  A CTA owns a contiguous
// share of the batch; its warps draw 32-pose groups from a shared-memory counter (which warp evaluates a pose changes
// nothing in its result, and the kernel no longer ends with the warp that was dealt the heaviest poses).
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
//     Tables that fit only twice keep two replicas with the halves in two planes (kPlanes), tables that fit once
//     fetch the second halves through the texture unit, larger ones both halves.
//   * per-thread row masks in shared memory, [row][thread] (a thread only touches its own column:
//     conflict-free, no synchronisation).
//
// Per pose (thread), per 32 x 32 chunk of its window (one chunk for tables whose windows fit):
//   phase 1  load the <= 5 x 2 tiles, funnel-shift each row's 32 window bits into place, AND them
//            with the row's x-interval of the rotated kernel rectangle (fp32, conservative by 0.02
//            cell; linear in the row index, so 4 FMAs per row), park the row mask, note non-empty
//            rows in a 32-bit row mask.
//   phase 2  walk the set bits (find-leading-one over the row mask, then the mask): k' = k - 1.5 =
//            k0 + x (c,-s) + y (s, c) in FP64; floor(k') by adding 1.5 * 2^52 rounding down selects the
//            patch; two LDS.128 fetch its bilinear polynomial in absolute k' coordinates (so no per-cell
//            fraction); f, f_u, f_v by FMA; accumulate.  There is no bounds test (dpose_core.cpp:181-182):
//            patches outside the valid range are zero and the clip keeps every cell within one patch of
//            it (CLIP INVARIANT in the kernel) — the loop body is branch-free, 33 instructions, 15 FP64.
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
// 4 bytes -> 4 bits (bit j set iff byte j == 254).  The SIMD-in-a-word byte compare of older architectures
// (__vcmpeq4) is emulated on sm_100 and made k_pack_tiles instruction bound (ncu: issue slots 66 % busy, "math pipe
// throttle" the top stall); exact zero-byte detection needs three logic ops: z = v ^ 0xFEFEFEFE is zero exactly in the
// lethal bytes, (z & 0x7F..) + 0x7F.. carries into bit 7 of every byte whose low 7 bits are non-zero without ever
// crossing a byte, so ~(that | z) & 0x80.. flags the zero bytes.  One multiply gathers the four flags (bits 7, 15, 23,
// 31 land on bits 28 .. 31; no other partial product reaches them).
__device__ __forceinline__ unsigned lethal_nibble(unsigned v) {
  const unsigned z = v ^ 0xFEFEFEFEu;
  const unsigned t = ((z & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | z;
  const unsigned m = ~t & 0x80808080u;
  return (m * 0x00204081u) >> 28;
}
__device__ __forceinline__ unsigned lethal_bits32(const uint32_t v[8]) {
  unsigned out = 0u;
#pragma unroll
  for (int k = 0; k < 8; ++k) out |= lethal_nibble(v[k]) << (4 * k);
  return out;
}

// A CTA of 8 warps packs blocks of 8 map rows x 1024 columns (32 tiles = 1 KB of plane): warp w reads row w of the
// block as ONE contiguous kilobyte (a 256-bit load per lane — long DRAM bursts instead of eight 128-byte pieces of eight
// different rows), every lane turns its 32 bytes into one plane word, the words are transposed through shared memory
// and leave as one coalesced kilobyte of tiles.  Two blocks per iteration, both loads issued before the first compare:
// 64 bytes in flight per thread.
__global__ void __launch_bounds__(256) k_pack_tiles(const uint8_t *__restrict__ map, size_t pitch, uint32_t size_y,
                                                    uint32_t *__restrict__ tiles, uint32_t ntx, uint32_t ty0,
                                                    uint32_t ty1) {
  // Programmatic dependent launch, both ways.  This grid may have been set up while its predecessor in the stream (the
  // batch kernel of the previous call, still reading the plane this grid overwrites, or the caller's kernel that wrote the
  // map) was running: wait for that grid's completion before touching anything.  The batch kernel queued behind this one
  // may start its prologue (staging the coefficient image) now; it waits for this grid's completion before it reads a tile.
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  __shared__ uint32_t s_words[2][32][9];  // [block of the pair][tile][row], padded: conflict-free both ways
  const uint32_t lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const uint32_t chunks = (ntx + 31) / 32;  // blocks per tile row
  const uint32_t n_blocks = chunks * (ty1 - ty0);  // < 2^32: the host checks (32-bit index arithmetic: a 64-bit
                                                   // division per block was a third of this kernel's instructions)
  auto fetch = [&](uint32_t blk, uint32_t v[8]) {  // pitch % 32 == 0 and the pad bytes of the device copy are zero
    const uint32_t tr = blk / chunks, c = blk - tr * chunks;
    const uint32_t y = (ty0 + tr) * 8 + w;
    const size_t x = ((size_t)c * 32 + lane) * 32;
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = 0u;
    if (y < size_y && x < pitch)
      asm volatile("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                   : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                   : "l"(map + (size_t)y * pitch + x));
  };
  auto store = [&](uint32_t blk, int which) {  // thread t writes word t of the block's kilobyte: tile t >> 3, row t & 7
    const uint32_t tr = blk / chunks, c = blk - tr * chunks;
    const uint32_t tile = c * 32 + (threadIdx.x >> 3);
    if (tile < ntx) tiles[((size_t)(ty0 + tr) * ntx + tile) * 8 + (threadIdx.x & 7)] = s_words[which][threadIdx.x >> 3][threadIdx.x & 7];
  };
  for (uint32_t blk = blockIdx.x; blk < n_blocks; blk += 2 * gridDim.x) {
    const uint32_t blk2 = blk + gridDim.x;
    const bool two = blk2 < n_blocks;
    uint32_t v0[8], v1[8];
    fetch(blk, v0);
    if (two) fetch(blk2, v1);
    s_words[0][lane][w] = lethal_bits32(v0);
    if (two) s_words[1][lane][w] = lethal_bits32(v1);
    __syncthreads();
    store(blk, 0);
    if (two) store(blk2, 1);
    __syncthreads();
  }
}

// ---- replicated coefficient image -----------------------------------------------------------------------
// The kernel evaluates a patch in ABSOLUTE shifted kernel coordinates k' = k - 1.5 (no per-cell fraction):
//   f(k') = b00 + b10 ku + b01 kv + b11 ku kv,   with  (U, V) = floor(k') = k_upper - 2  and
//   b11 = a11, b10 = a10 - a11 V, b01 = a01 - a11 U, b00 = a00 - a10 U - a01 V + a11 U V
// (the bilinear patch of dpose_core.cpp:189-198 expanded around the patch origin; |U|, |V| < 2^8 keeps
// the cancellation error below 1e-13, eight orders under the parity tolerance).  Patches that fail the
// reference's bounds test (:181-182) hold zeros and contribute exactly 0: the walk needs no bounds test
// as long as every visited cell lands within one patch of the valid range — see the clip invariant
// in the kernel.
// image[((patch * 2 + half) * R + rep)] (16 bytes each), patch = k_upper.y * cols + k_upper.x
__global__ void k_build_coef_image(const PatchCoef *__restrict__ coef, double2 *__restrict__ image,
                                   double2 *__restrict__ hi_only, int rows, int cols, int rep_shift, int planes) {
  const int n_patches = rows * cols;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_patches; i += gridDim.x * blockDim.x) {
    const int uy = i / cols, ux = i - uy * cols;
    double2 lo = make_double2(0.0, 0.0), hi = make_double2(0.0, 0.0);  // (b00, b10), (b01, b11)
    if (ux >= 2 && uy >= 2 && ux <= cols - 2 && uy <= rows - 2) {
      const PatchCoef a = coef[i];
      const double U = (double)(ux - 2), V = (double)(uy - 2);
      hi.y = a.a11;
      lo.y = a.a10 - a.a11 * V;
      hi.x = a.a01 - a.a11 * U;
      lo.x = (a.a00 - a.a10 * U) - (a.a01 * V - (a.a11 * U) * V);
    }
    if (hi_only) {  // texture modes: shared image = first halves only (if any), both halves in a plain global array
      // two planes, second halves first: the single-replica mode fetches only those, and packed 16 bytes apart they
      // take half the texture-cache footprint of an interleaved layout (cfg3: 1.24e9 against 1.16e9 poses/s)
      hi_only[(size_t)i] = hi;
      hi_only[(size_t)n_patches + i] = lo;
      if (image)
        for (int r = 0; r < (1 << rep_shift); ++r) image[((size_t)i << rep_shift) + r] = lo;
    } else {
      for (int r = 0; r < (1 << rep_shift); ++r) {
        if (planes) {  // two planes of 16-byte halves (the two-replica mode, see kPlanes in the kernel)
          image[((size_t)i << rep_shift) + r] = lo;
          image[(((size_t)n_patches + i) << rep_shift) + r] = hi;
        } else {
          image[(((size_t)i * 2 + 0) << rep_shift) + r] = lo;
          image[(((size_t)i * 2 + 1) << rep_shift) + r] = hi;
        }
      }
    }
  }
}

// ---- PTX wrappers ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ unsigned lds_u32(uint32_t addr) {
  unsigned v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}

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
  TableGeom g;
  int size_x, size_y;       // size_y: end of the band of map rows that count (the map's height unless BatchSpec::row_hi)
  int y_lo;                 // its start
  const uint32_t *tiles;
  uint32_t ntx, nty;
  const double *poses;
  double *out;
  int cost_only;  // out holds n costs instead of n x 4
  size_t n;
  cudaTextureObject_t tex_hi;  // texture mode: second halves of the patch polynomials (int4 texels)
  const int4 *planes;          // the same buffer as a plain pointer (DPOSE_TILE_LDG experiment)
  int out_aligned;  // out is 32-byte aligned: one 256-bit store per result
  unsigned long long per_cta;  // CTA c owns the poses [c per_cta, (c + 1) per_cta) (a multiple of 32; equal shares)
  float clip_thr;   // a strip of the kernel rectangle is clipped per row only if |cos| (|sin|) exceeds this
  float clip_margin;  // the clipped x-interval is widened by this many cells on both sides (fp32 round-off, see the launch code)
};

struct Acc4 {
  double f, su, sv, st;
};

// lattice poses of a sweep, gi = base .. base + n (see sweep_pose, common.cuh)
__global__ void __launch_bounds__(256) k_sweep_poses(dpose_sweep sw, size_t base, size_t n, double *__restrict__ poses) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double x, y, th;
  sweep_pose(sw, base + i, &x, &y, &th);
  poses[3 * i] = x, poses[3 * i + 1] = y, poses[3 * i + 2] = th;
}

// DPOSE_TILE_CVT: how the walk turns a bit / row index into a double.
//   0: 2^52 trick (one DADD on the FP64 pipe + register-pair set-up), 1: I2F.F64 (conversion pipe: with the two FLO of
//   an iteration that is 4 x 8 issue cycles on the 4-lane XU, as busy as the FP64 pipe), 2: no conversion at all — the
//   index x < 32 is planted into the top five mantissa bits of 1.0 by an integer multiply-add on the high word
//   (1 + x / 32, exact), and the walk's FMAs run on 32 c, 32 s with the constant terms folded into the chunk origin.
// DPOSE_TILE_COMPACT 1 stores only the rows that can belong to a chunk: less shared memory, so that 8
// coefficient replicas (conflict-free LDS.128 gathers) fit next to 1024 threads' row masks.  Measured on B200
// (cfg5): 9.26e9 poses/s against 8.89e9 for "every slot stored, 4 replicas"; the L1/shared data pipe drops from
// 85 % to ~60 % busy.  0 keeps every scanned slot (no store predicates).
#ifndef DPOSE_TILE_COMPACT
#define DPOSE_TILE_COMPACT 1
#endif
#ifndef DPOSE_TILE_THREADS
#define DPOSE_TILE_THREADS 1024  // largest CTA (A/B builds: 512, 768)
#endif
#ifndef DPOSE_TILE_REPLICAS
#define DPOSE_TILE_REPLICAS 8  // coefficient replicas wanted in shared memory (A/B builds: 1, 4)
#endif
// Big tables (the coefficient image fits shared memory only once: kRepShift == 0) run in "texture mode": shared
// memory holds the (b00, b10) halves twice (same bytes as one full copy, half the bank conflicts) and the
// (b01, b11) halves are fetched through the texture unit from a plain global array, which takes those gathers
// off the LSU data pipe.  Measured on B200, 35 x 55 table (cfg3): +7 %; on the small rect table, where 8 full
// replicas fit, it is 5 % SLOWER than shared memory only — hence tied to the single-replica case.
// Tables that do not fit shared memory at all (kRepShift == -1) fetch both halves through the texture unit: the
// kernel then has no table-size limit (L1 / L2 serve the gathers).
// DPOSE_TILE_CHECK 1: debug build that traps on any access outside the coefficient image, the tile plane or the
// row-mask block (compute-sanitizer is not available on the GPU pool): run the GPU tests and the bench with it
// after touching the clip logic.
#ifndef DPOSE_TILE_LDG
#define DPOSE_TILE_LDG 0  // 1: fetch the texture-mode planes with LDG.128.CONSTANT instead of TEX (A/B experiment)
#endif
#ifndef DPOSE_TILE_CHECK
#define DPOSE_TILE_CHECK 0
#endif
#ifndef DPOSE_TILE_CVT
#define DPOSE_TILE_CVT 2
#endif
__device__ __forceinline__ double index_to_double(int x, int zero) {  // zero: an opaque 0 (CVT 2: the low word)
#if DPOSE_TILE_CVT == 0
  return small_int_to_double(x);
#elif DPOSE_TILE_CVT == 2
  return unit_plus_32nd(x, zero);
#else
  return (double)x;
#endif
}

// kRepShift: log2 of the coefficient replicas; kRows: rows of a chunk (>= min(32, largest window height)).  A
// chunk's rows start at slot (y & 7) of its first tile row, so kRows + 7 row slots are scanned; only the
// kRows that can belong to the chunk are stored.
template <bool kJac, int kThreads, int kRepShift, int kRows>
__global__ void __launch_bounds__(kThreads, 1) k_cost_batch_tile(TileArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) unsigned long long s_bar;
  __shared__ unsigned s_next;  // next 32-pose group of this CTA's share
  const int tid = threadIdx.x, lane = tid & 31;
  if (tid == 0) s_next = 0u;  // published by the barrier below

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
  // The wait for the image is deferred to the first cell walk: phase 1 of the first pose needs no coefficients, so the
  // ~109 KB of bulk copies (all 148 CTAs pull the same image out of L2 at once) land behind ~600 instructions of useful
  // work instead of in front of them — 2-3 us of a 15 us launch on a solver-sized batch.
  // The barrier is never re-armed, so once phase 0 has completed every later wait returns at its first poll: the wait is
  // simply issued in front of every walk (3 instructions of ~2600 per pose, no flag register).
  auto wait_staged = [&]() {
    // try_wait suspends for a hardware time slice and returns false on time-out: loop until the bytes have landed.
    // No trap after a bounded number of polls — a CTA delayed by time-slicing / MPS / a debugger must not take the
    // whole context down; the copies were issued by this CTA, so completion is guaranteed.
    uint32_t ok = 0;
    while (!ok)
      asm volatile(
          "{\n\t.reg .pred p;\n\t"
          "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\t"
          "selp.u32 %0, 1, 0, p;\n\t}"
          : "=r"(ok)
          : "r"(bar)
          : "memory");
  };
  // Programmatic dependent launch: this kernel may have been started while its predecessor in the stream (k_pack_tiles
  // rebuilding the tile plane, or the kernel that wrote the poses) was still running, so that the launch latency and the
  // staging above overlap it.  Nothing of the predecessor's output is touched before this point.
  asm volatile("griddepcontrol.wait;" ::: "memory");
  // (This kernel does NOT trigger its own dependents early.  Measured: with griddepcontrol.launch_dependents here, the
  // plane rebuild of the next call is set up while this grid runs and its CTAs then sit on freed SMs waiting — a volatile-map
  // step got 2 us (1e5 poses) to 7 us (1.25e6 poses) SLOWER, profiles/r2_ab_tile_walk.txt run v39.)

  constexpr int kSlots = kRows + 7;
  // DPOSE_TILE_COMPACT 1: only the kRows rows that can belong to a chunk are stored (less shared memory, room
  // for 8 coefficient replicas at 1024 threads); 0: every scanned slot is stored (no store predicates)
  constexpr int kStored = DPOSE_TILE_COMPACT ? kRows : kSlots;
  (void)kStored;
  // [row * kThreads], -1 <= row < kStored: row -1 is a guard the rotated cell loop may read (never use) when it runs out
  // of rows — its last row advance is not predicated on "a row is left"
  uint32_t *s_mask = reinterpret_cast<uint32_t *>(smem + a.coef_bytes) + kThreads + tid;
  constexpr bool kTex = kRepShift <= 0;             // texture mode, see above
  constexpr bool kGlobal = kRepShift < 0;           // nothing of the table in shared memory
  constexpr int kLoShift = kTex ? 1 : kRepShift;    // log2 replicas of what lives in shared memory
  // Two replicas (kRepShift == 1): the halves live in two PLANES, plane[(patch * 2 + rep)] — the 16-byte bank group of a
  // gather is then (patch & 3) * 2 + rep, eight classes for the eight lanes of an LDS.128 phase.  Interleaved halves
  // would leave four ((patch & 1) * 4 + half * 2 + rep with half fixed per instruction): measured on the 35 x 55 table,
  // 84.7 M bank conflicts for 165 M wavefronts, the LSU data pipe 89 % busy.  With 4 / 8 replicas the interleaved form
  // is conflict-free per replica and its second half sits at an immediate offset.
  constexpr bool kPlanes = kRepShift == 1;
  constexpr uint32_t kPatchStride = (kTex || kPlanes ? 16u : 32u) << kLoShift, kHalfOff = 16u << kLoShift;
  (void)kHalfOff;
  const uint32_t plane_off = a.coef_bytes / 2;  // kPlanes: bytes between the two planes
  (void)plane_off;
  const int cols = a.g.cols;
  // Kernel coordinates are carried as k' = k + 0.5 - 2 = k - 1.5: floor(k') = k_upper - 2 (k_upper = round(k),
  // dpose_core.cpp:177), frac(k') = c_rel (:193), and the bounds test (:181-182) becomes
  // 0 <= floor(k') <= dim - 4.  The patch of k_upper = (2, 2) therefore sits at the base address.
  uint32_t my_coef =
      coef_u32 + ((uint32_t)lane & ((1u << kLoShift) - 1u)) * 16u + (uint32_t)(2 * cols + 2) * kPatchStride;
  uint32_t colstride = (uint32_t)cols * kPatchStride;  // bytes between the images of two table rows
  asm volatile("" : "+r"(my_coef), "+r"(colstride));  // two registers: nothing of either is re-derived inside the cell walk
  const float hi_u = (float)a.g.cols - 3.0f, hi_v = (float)a.g.rows - 3.0f;  // 0 <= k' < dim - 3
  const double cx15 = a.g.cx - 1.5, cy15 = a.g.cy - 1.5;

  // The CTA owns a contiguous share of the batch and its WARPS draw 32-pose groups from it through a shared-memory
  // counter: a warp's poses differ by +-12 % in cost, so with a static deal (the same lanes sweep after sweep) the warp
  // that drew the heaviest poses finished ~10 us after the SM's average (tools/sweep_timing.py: 15.6 us of fixed cost per
  // launch with random poses, 5.2 us with identical ones) — a tenth of an 8-GPU shard's step.  Which warp evaluates a
  // pose changes nothing in its result.
  const size_t cta_base = (size_t)blockIdx.x * (size_t)a.per_cta;
  const size_t cta_end = min(a.n, cta_base + (size_t)a.per_cta);
  // The draw sits at the top of the loop, where the warp is converged anyway (at the end of a pose the lanes leave the
  // cell walk one by one: a shuffle there costs a second reconvergence), and it is made when the previous group is DONE:
  // a group drawn ahead is a group no idle warp can take (measured on the 35 x 55 table, where a group is 80 us of a
  // warp's time: -2.2 %).  Dealing a single-wave batch statically instead (no draw) gains 1.7 % at 1e5 poses and costs
  // the large batches 0.5 - 2 % through the extra loop state (profiles/r2_ab_tile_walk.txt, runs v33 - v37): one form.
  for (;;) {
    unsigned grp = 0u;
    if (lane == 0) grp = atomicAdd(&s_next, 1u);
    grp = __shfl_sync(0xffffffffu, grp, 0);
    const size_t gbase = cta_base + (size_t)grp * 32u;
    if (gbase >= cta_end) break;  // warp-uniform
    const size_t idx = gbase + (size_t)lane;
    if (idx >= cta_end) continue;  // the ragged last group (the next draw ends the loop)
    const double *pp = a.poses + 3 * idx;
    const double tx = __ldg(pp), ty = __ldg(pp + 1), th = __ldg(pp + 2);
    double s, c;
    sincos(th, &s, &c);
    const WindowRect w = pose_window(a.g, tx, ty, c, s, a.size_x, a.size_y, a.y_lo);
    const int W = w.x1 - w.x0 + 1, H = w.y1 - w.y0 + 1;  // <= 0 when empty
    double ku0, kv0;  // k' at the window origin
    {
      const double dx = (double)w.x0 - tx, dy = (double)w.y0 - ty;
      ku0 = cx15 + fma(c, dx, s * dy);
      kv0 = cy15 + fma(c, dy, -s * dx);
    }
    // conservative fp32 x-interval of the kernel rectangle per window row y:
    //   0 <= ku0 + s y + c x < cols - 3   and   0 <= kv0 + c y - s x < rows - 3, each linear in y.
    // All values stay below 2^22 (|t| <= 2 win_max^2 and win_max <= 1024, see dpose_pg_create), which the integer
    // extraction in phase 1 relies on.
    // CLIP INVARIANT (what lets phase 2 drop the bounds test): every cell that survives has k' within one
    // patch of the valid range, so floor(k') is in [-1, dim - 3] and addresses a patch of the table (zeros
    // outside the valid range).  A clipped strip overshoots by clip_margin |c| cell, where clip_margin (< 0.5 for
    // every table dpose_pg_create admits) covers the fp32 round-off of the interval ends, which grows with
    // win_max^2; a strip left unclipped because |c| <= clip_thr = 0.5 / win_max varies by at most
    // |c| * window width <= 0.5 cell along a window row, and every window row crosses the valid rectangle (the
    // window is its bounding box).
    float l1 = -1000.f, h1 = 1000.f, d1 = 0.f, l2 = -1000.f, h2 = 1000.f, d2 = 0.f;
    {
      const float cf = (float)c, sf = (float)s, ku0f = (float)ku0, kv0f = (float)kv0;
      if (fabsf(cf) > a.clip_thr) {
        const float inv = rcp_approx(cf);
        const float t0 = -ku0f * inv, t1 = (hi_u - ku0f) * inv;
        l1 = fminf(t0, t1) - a.clip_margin, h1 = fmaxf(t0, t1) + a.clip_margin, d1 = -sf * inv;
      }
      if (fabsf(sf) > a.clip_thr) {
        const float inv = rcp_approx(sf);
        const float t0 = kv0f * inv, t1 = (kv0f - hi_v) * inv;
        l2 = fminf(t0, t1) - a.clip_margin, h2 = fmaxf(t0, t1) + a.clip_margin, d2 = cf * inv;
      }
    }

    Acc4 acc = {0.0, 0.0, 0.0, 0.0};
#if DPOSE_TILE_CVT == 2
    const double cw = 32.0 * c, sw = 32.0 * s;
#endif
    // low words of the walk's index doubles: opaque zeros that stay in the low halves of two register pairs, so that
    // an index costs one integer instruction on the high half (a literal 0 is re-materialised in every iteration)
    int zero_x = (int)(a.per_cta >> 48), zero_y = (int)(a.ntx >> 31);  // both 0 (per_cta <= n < 2^48, ntx < 2^31)
    asm volatile("" : "+r"(zero_x), "+r"(zero_y));
    // kRows < 32: every window of the table fits one 32 x 32 chunk — the chunk loops and what they keep alive across
    // the cell walk (the clip intervals, the window, its origin in kernel coordinates) drop out at compile time
    constexpr bool kSingle = kRows < 32;
    int rb = 0;
    if (H > 0 && W > 0) do {
      const int hc = min(32, H - rb), ys = w.y0 + rb;
      int cb = 0;
      do {
        const int wc = min(32, W - cb), xs = w.x0 + cb;
        const int off = ys & 7;  // chunk row i sits in row slot i + off of the chunk's tile rows
        // ---- phase 1: row masks of this chunk; branch-free per slot, every slot is stored ---------------------
        unsigned rowmask;
        {
          const unsigned sh = (unsigned)xs & 31u;
          const bool need_b = (int)sh + wc > 32;
          const int tb0 = ys >> 3, n_tb = ((ys + hc - 1) >> 3) - tb0;  // last tile row index, relative
#if DPOSE_TILE_CHECK
          if (tb0 < 0 || tb0 + n_tb >= (int)a.nty || (xs >> 5) < 0 || (xs >> 5) + (need_b ? 1 : 0) >= (int)a.ntx || hc > kRows ||
              wc > 32 || n_tb > 4)
            __trap();
#endif
          const uint32_t *tp = a.tiles + ((size_t)tb0 * a.ntx + (size_t)(xs >> 5)) * 8;
          const size_t tstride = (size_t)a.ntx * 8;
          const float fcb = (float)cb, fr0 = (float)(rb - off);
          const float l1c = fmaf(fr0, d1, l1) - fcb, h1c = fmaf(fr0, d1, h1) - fcb;
          const float l2c = fmaf(fr0, d2, l2) - fcb, h2c = fmaf(fr0, d2, h2) - fcb;
          const float wmax = (float)(wc - 1);
          unsigned sm_lo = 0u, sm_hi = 0u;  // non-empty slots
          // slot j holds chunk row j - off; rows outside [0, kRows) are not stored (slots 7 .. kRows - 1 always are)
          uint32_t *s_slot = DPOSE_TILE_COMPACT ? s_mask - off * kThreads : s_mask;
          // with 1024 threads (64 registers) the next tile row is not prefetched: the extra warps hide the latency
          constexpr bool kPrefetch = kThreads < 1024;
          Tile8 A, B, An, Bn;
#pragma unroll
          for (int r = 0; r < 8; ++r) A.w[r] = B.w[r] = An.w[r] = Bn.w[r] = 0u;
          if (kPrefetch) {
            A = ldg_tile(tp);
            if (need_b) B = ldg_tile(tp + 8);
          }
#pragma unroll
          for (int t = 0; t < 5; ++t) {
            if (t * 8 < kSlots) {
              if (kPrefetch) {
                if ((t + 1) * 8 < kSlots && t + 1 <= n_tb) {  // prefetch the next tile row
                  An = ldg_tile(tp + (size_t)(t + 1) * tstride);
                  if (need_b) Bn = ldg_tile(tp + (size_t)(t + 1) * tstride + 8);
                }
              } else if (t <= n_tb) {
                A = ldg_tile(tp + (size_t)t * tstride);
                if (need_b) B = ldg_tile(tp + (size_t)t * tstride + 8);
              }
#pragma unroll
              for (int r = 0; r < 8; ++r) {
                const int j = t * 8 + r;
                if (j < kSlots) {
                  const float fj = (float)j;
                  const float lo = fmaxf(fmaxf(fmaf(fj, d1, l1c), fmaf(fj, d2, l2c)), 0.0f);
                  const float hi = fminf(fminf(fmaf(fj, d1, h1c), fmaf(fj, d2, h2c)), wmax);
                  // ceil / floor through the 1.5 * 2^23 trick: the integer lands in the low mantissa bits
                  const int xlo = __float_as_int(__fadd_ru(lo, 12582912.0f)) - 0x4B400000;  // >= 0
                  const int xhi = __float_as_int(__fadd_rd(hi, 12582912.0f)) - 0x4B400000;  // <= 31
                  unsigned m = __funnelshift_r(A.w[r], B.w[r], sh);
                  m &= shl_clamp(0xffffffffu, (unsigned)xlo) & shr_clamp(0xffffffffu, (unsigned)(31 - xhi));
                  if (!DPOSE_TILE_COMPACT || (j >= 7 && j < kRows))
                    s_slot[j * kThreads] = m;
                  else if ((unsigned)(j - off) < (unsigned)kRows)
                    s_slot[j * kThreads] = m;
                  if (m) {
                    if (j < 32)
                      sm_lo |= 1u << (j & 31);
                    else
                      sm_hi |= 1u << (j & 31);
                  }
                }
              }
              if (kPrefetch) {
                A = An;
                B = Bn;
              }
            }
          }
          // slots [off, off + hc) are the chunk's rows; everything else (stale tiles, rows past the window) is dropped
          rowmask = __funnelshift_r(sm_lo, sm_hi, (unsigned)off) & shr_clamp(0xffffffffu, (unsigned)(32 - hc));
        }
        // ---- phase 2: walk the set bits, highest row / highest bit first ---------------------------------------
        // 32-bit shared address of chunk row 0, pinned in a register: without the opaque asm the compiler
        // re-derives it from threadIdx and the kernel arguments inside the walk (measured: +20 % instructions)
        uint32_t srow = smem_u32(DPOSE_TILE_COMPACT ? s_mask : s_mask + off * kThreads);
        asm volatile("" : "+r"(srow));
#if DPOSE_TILE_CVT == 2
        // X = 1 + b / 32, Y = 1 + row / 32 (index_to_double): c b + s row = cw X + sw Y - cw - sw with cw = 32 c, sw = 32 s
        // (exact scalings); the constants ride on the chunk origin: cb / 32 - 1 and rb / 32 - 1 are exact integers
        const double dcb = (double)((cb >> 5) - 1), drb = (double)((rb >> 5) - 1);
        const double kuc = fma(cw, dcb, fma(sw, drb, ku0)), kvc = fma(cw, drb, fma(-sw, dcb, kv0));
#else
        const double dcb = (double)cb, drb = (double)rb;
        const double kuc = fma(c, dcb, fma(s, drb, ku0)), kvc = fma(c, drb, fma(-s, dcb, kv0));
#endif
        wait_staged();
        // One cell = bit b of window row `row`.  gather: k' of the cell, its patch, the two 16-byte halves of the patch
        // polynomial; accumulate: value and derivatives.  Split so that the loop below can put the selection of the next
        // cell between the two.
        struct Cell {
          double ku, kv, b00, b10, b01, b11;
        };
        auto gather = [&](int b, int row) {
          Cell q;
          const double X = index_to_double(b, zero_x), Y = index_to_double(row, zero_y);
#if DPOSE_TILE_CVT == 2
          q.ku = fma(cw, X, fma(sw, Y, kuc));
          q.kv = fma(cw, Y, fma(-sw, X, kvc));
#else
          q.ku = fma(c, X, fma(s, Y, kuc));
          q.kv = fma(c, Y, fma(-s, X, kvc));
#endif
          const double M = 6755399441055744.0;  // 1.5 * 2^52: adding it (rounding down) floors to an integer
          const int ux = __double2loint(__dadd_rd(q.ku, M)), uy = __double2loint(__dadd_rd(q.kv, M));  // k_upper - 2
          // no bounds test (dpose_core.cpp:181-182): out-of-range patches are zero, see the clip invariant
          const int pidx = uy * cols + ux;
#if DPOSE_TILE_CHECK
          if (pidx + 2 * cols + 2 < 0 || pidx + 2 * cols + 2 >= a.g.rows * cols || row < 0 || row >= kRows || ux < -2 ||
              ux > cols - 3 || uy < -2 || uy > a.g.rows - 3)
            __trap();
#endif
          // shared-memory modes: two multiply-adds on registers (base + uy * row stride + ux * patch stride)
          const uint32_t addr = kTex ? my_coef + (uint32_t)pidx * kPatchStride
                                     : my_coef + (uint32_t)uy * colstride + (uint32_t)ux * kPatchStride;
          if (kGlobal) {
            const int4 t = DPOSE_TILE_LDG ? __ldg(a.planes + (a.g.rows * cols + pidx + 2 * cols + 2))
                                          : tex1Dfetch<int4>(a.tex_hi, a.g.rows * cols + pidx + 2 * cols + 2);
            q.b00 = __hiloint2double(t.y, t.x);
            q.b10 = __hiloint2double(t.w, t.z);
          } else {
            asm("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(q.b00), "=d"(q.b10) : "r"(addr));
          }
          if (kTex) {
            const int4 t = DPOSE_TILE_LDG ? __ldg(a.planes + (pidx + 2 * cols + 2)) : tex1Dfetch<int4>(a.tex_hi, pidx + 2 * cols + 2);
            q.b01 = __hiloint2double(t.y, t.x);
            q.b11 = __hiloint2double(t.w, t.z);
          } else {
            if (kPlanes)
              asm("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(q.b01), "=d"(q.b11) : "r"(addr + plane_off));
            else
              asm("ld.shared.v2.f64 {%0, %1}, [%2+%3];" : "=d"(q.b01), "=d"(q.b11) : "r"(addr), "n"(kHalfOff));
          }
          return q;
        };
        auto accumulate = [&](const Cell &q) {
          const double fv = fma(q.b11, q.ku, q.b01);  // df/dkv
          acc.f += fma(q.kv, fv, fma(q.b10, q.ku, q.b00));
          if (kJac) {
            const double fu = fma(q.b11, q.kv, q.b10);  // df/dku
            acc.su += fu;
            acc.sv += fv;
            acc.st = fma(fu, q.kv, acc.st);  // Sum(fu * kv' - fv * ku'); center terms folded in below
            acc.st = fma(-fv, q.ku, acc.st);
          }
        };
        auto row_mask = [&](int row) { return lds_u32(srow + (uint32_t)row * (uint32_t)(kThreads * 4)); };
        // The row advance is branch-free in both forms: selects and one predicated load (a nested branch measured slower).
        if (!kTex) {
          // Rotated loop: the selection of the NEXT cell (row advance = an LDS of the next row's mask, then a
          // find-leading-one) sits between the gathers of this cell's polynomial and the arithmetic that consumes them, so
          // that the two latency chains of an iteration overlap inside one warp instead of following each other
          // (+1 % at cfg5; the texture modes lose 3 % with it — their gathers are long enough to cover everything).
          if (rowmask != 0u) {
            int row = top_bit(rowmask);
            rowmask = clear_top(rowmask, row);
            unsigned mask = row_mask(row);
            int b = top_bit(mask);
            while (true) {
              mask = clear_top(mask, b);
              const Cell q = gather(b, row);
              // (when nothing is left, adv still holds: r2 = -1, the guard row is read and the loop ends)
              const bool adv = mask == 0u;
              const bool more = !adv || rowmask != 0u;
              const int r2 = top_bit(rowmask);
              row = adv ? r2 : row;
              rowmask = adv ? clear_top(rowmask, r2) : rowmask;
              if (adv) mask = row_mask(row);
              b = top_bit(mask);
              accumulate(q);
              if (!more) break;
            }
          }
        } else {
          unsigned mask = 0u;
          int row = 0;
          while ((mask | rowmask) != 0u) {
            const bool adv = mask == 0u;
            const int r2 = top_bit(rowmask);  // -1 when no row is left; then adv is false and nothing below uses it
            row = adv ? r2 : row;
            rowmask = adv ? clear_top(rowmask, r2) : rowmask;
            if (adv) mask = row_mask(row);
            const int b = top_bit(mask);
            mask = clear_top(mask, b);
            accumulate(gather(b, row));
          }
        }
        cb += 32;
      } while (!kSingle && cb < W);
      rb += 32;
    } while (!kSingle && rb < H);
    double jx = 0.0, jy = 0.0, jt = 0.0;
    if (kJac) {
      jx = -c * acc.su + s * acc.sv;
      jy = -s * acc.su - c * acc.sv;
      jt = acc.st - cy15 * acc.su + cx15 * acc.sv;
    }
    if (a.cost_only) {
      a.out[idx] = acc.f;
    } else if (a.out_aligned) {
      stg_result(a.out + 4 * idx, acc.f, jx, jy, jt);
    } else {
      double *o = a.out + 4 * idx;
      o[0] = acc.f, o[1] = jx, o[2] = jy, o[3] = jt;
    }
  }
  wait_staged();  // a thread that met only empty windows: still no exit before the copies have landed
}


template <bool kJac, int kThreads, int kRepShift, int kRows>
int launch_tile4(const TileArgs &a, int grid, size_t smem_bytes, cudaStream_t stream) {
  auto kern = k_cost_batch_tile<kJac, kThreads, kRepShift, kRows>;
  // the opt-in shared-memory size is a per-device property of the function: raised once per (instantiation, device)
  static std::atomic<size_t> raised[64];
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64 || raised[dev].load(std::memory_order_relaxed) < smem_bytes) {
    DPOSE_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
    if (dev >= 0 && dev < 64) raised[dev].store(smem_bytes, std::memory_order_relaxed);
  }
  DPOSE_CUDA(launch_pdl(kern, dim3((unsigned)grid), dim3(kThreads), smem_bytes, stream, a));  // see griddepcontrol.wait
  DPOSE_LAUNCHED();
  return DPOSE_OK;
}
template <bool kJac, int kThreads, int kRepShift>
int launch_tile3(const TileArgs &a, int rows, int grid, size_t smem_bytes, cudaStream_t stream) {
  if (rows == 20) return launch_tile4<kJac, kThreads, kRepShift, 20>(a, grid, smem_bytes, stream);
  if (rows == 28) return launch_tile4<kJac, kThreads, kRepShift, 28>(a, grid, smem_bytes, stream);
  return launch_tile4<kJac, kThreads, kRepShift, 32>(a, grid, smem_bytes, stream);
}
template <bool kJac, int kThreads>
int launch_tile2(const TileArgs &a, int rep_shift, int rows, int grid, size_t smem_bytes, cudaStream_t stream) {
  if (rep_shift < 0) return launch_tile3<kJac, kThreads, -1>(a, rows, grid, smem_bytes, stream);
  if (rep_shift == 0) return launch_tile3<kJac, kThreads, 0>(a, rows, grid, smem_bytes, stream);
  if (rep_shift == 1) return launch_tile3<kJac, kThreads, 1>(a, rows, grid, smem_bytes, stream);
  if (rep_shift == 2) return launch_tile3<kJac, kThreads, 2>(a, rows, grid, smem_bytes, stream);
  return launch_tile3<kJac, kThreads, 3>(a, rows, grid, smem_bytes, stream);
}
template <bool kJac>
int launch_tile1(const TileArgs &a, int threads, int rep_shift, int rows, int grid, size_t smem_bytes,
                 cudaStream_t stream) {
  if (threads == 128) return launch_tile2<kJac, 128>(a, rep_shift, rows, grid, smem_bytes, stream);
  if (threads == 256) return launch_tile2<kJac, 256>(a, rep_shift, rows, grid, smem_bytes, stream);
  if (threads == 512) return launch_tile2<kJac, 512>(a, rep_shift, rows, grid, smem_bytes, stream);
  if (threads == 768) return launch_tile2<kJac, 768>(a, rep_shift, rows, grid, smem_bytes, stream);
  return launch_tile2<kJac, 1024>(a, rep_shift, rows, grid, smem_bytes, stream);
}

}  // namespace

std::mutex &map_state_mutex() {
  static std::mutex mu;  // guards the lazily built per-handle state: coefficient images, tile-plane dirty flags / events
  return mu;
}

// caller holds map_state_mutex()
static int tileplane_build_locked(const dpose_map *map, cudaStream_t stream, uint32_t row_lo, uint32_t row_hi) {
  if (!map->d_tiles || map->size_x == 0 || map->size_y == 0) return DPOSE_OK;
  // only the tile rows that cover the dirty map rows (all of them after create); a volatile map: every row, or the band
  // of rows the launch behind this rebuild is restricted to (BatchSpec::row_lo / row_hi) — whatever reads the plane of a
  // volatile map next rebuilds what it needs itself, so the rows outside the band may stay stale
  dpose_map *m = const_cast<dpose_map *>(map);
  const bool band = map->is_volatile && row_hi != 0;
  const uint32_t lo = band ? std::min(row_lo, map->size_y) : map->is_volatile ? 0u : std::min(map->tiles_dirty_lo, map->size_y);
  const uint32_t hi = band ? std::min(row_hi, map->size_y) : map->is_volatile ? map->size_y : std::min(map->tiles_dirty_hi, map->size_y);
  m->tiles_dirty_lo = 0xffffffffu;
  m->tiles_dirty_hi = 0;
  if (hi <= lo) return DPOSE_OK;
  const uint32_t ty0 = lo / 8, ty1 = (hi + 7) / 8;
  const size_t n_blocks = (size_t)((map->tiles_ntx + 31) / 32) * (ty1 - ty0);
  if (n_blocks >= ((size_t)1 << 31)) return fail(DPOSE_EINVAL, "costmap too large for the tile plane (more than 2^44 cells)");
  const unsigned grid = (unsigned)std::min<size_t>((n_blocks + 1) / 2, (size_t)148 * 8);
  DPOSE_CUDA(launch_pdl(k_pack_tiles, dim3(grid), dim3(256), 0, stream, map->d_data, map->pitch, map->size_y, map->d_tiles,
                        map->tiles_ntx, ty0, ty1));  // see griddepcontrol.wait in the kernel
  DPOSE_LAUNCHED();
  return DPOSE_OK;
}

// Makes the lethal tile plane of `map` current for work queued on `stream` afterwards: rebuilds the dirty rows (every
// row of a volatile map unless hold_planes) in front of it, or orders `stream` behind the rebuild another stream did.
int tileplane_ensure(const dpose_map *map, cudaStream_t stream, bool hold_planes, uint32_t row_lo, uint32_t row_hi) {
  if (!map->d_tiles) return DPOSE_OK;
  dpose_map *m = const_cast<dpose_map *>(map);
  std::lock_guard<std::mutex> lock(map_state_mutex());
  if (!m->tiles_ready) DPOSE_CUDA(cudaEventCreateWithFlags(&m->tiles_ready, cudaEventDisableTiming));
  if (m->tiles_dirty || (m->is_volatile && !hold_planes)) {
    DPOSE_TRY(tileplane_build_locked(map, stream, row_lo, row_hi));
    DPOSE_CUDA(cudaEventRecord(m->tiles_ready, stream));
    m->tiles_dirty = false;  // (a volatile map is rebuilt in front of every use whatever this flag says)
  } else {
    DPOSE_CUDA(cudaStreamWaitEvent(stream, m->tiles_ready, 0));  // built on another stream?
  }
  return DPOSE_OK;
}

// The patch polynomials in absolute k' coordinates, one 32-byte record (b00, b10, b01, b11) per patch, zeros outside the
// valid range: what the cooperative kernels (coop_eval.cuh) gather straight from global memory / L1.  Built on first use.
int coef_abs_ensure(const dpose_pg *pg, cudaStream_t stream) {
  dpose_pg *p = const_cast<dpose_pg *>(pg);
  std::lock_guard<std::mutex> lock(map_state_mutex());
  if (p->d_coef_abs) return DPOSE_OK;
  double2 *buf = nullptr;
  DPOSE_CUDA(cudaMalloc(&buf, sizeof(PatchCoef) * (size_t)pg->rows * pg->cols));
  k_build_coef_image<<<64, 256, 0, stream>>>(pg->d_coef, buf, nullptr, pg->rows, pg->cols, 0, 0);
  DPOSE_LAUNCHED();
  DPOSE_CUDA(cudaStreamSynchronize(stream));  // other streams may use it right after
  p->d_coef_abs = reinterpret_cast<double *>(buf);
  return DPOSE_OK;
}

// fp32 round-off of the clip interval ends grows with win_max^2 (|t| <= 2 win_max^2 before the cancellation): measured
// worst case 3.5 win_max^2 2^-23 cells (float32 emulation of the exact operation order at angles just above clip_thr);
// the margin carries twice that plus the 0.02 cell the small tables were tuned with.  dpose_pg_create caps tables at
// 512 x 512 (win_max <= 724), where margin = 0.52 and margin + error < 1 patch: the CLIP INVARIANT holds.
static float clip_margin_for(int win_max) { return 0.02f + 8.0f * (float)win_max * (float)win_max * 1.1920929e-7f; }

// Tables with fewer than 4 rows or columns have no valid patch (dpose_core.cpp:181-182 needs 2 <= k_upper <= dim - 2):
// every cost is 0, like an empty map.  Everything else runs k_cost_batch_tile.
int get_cost_batch_tile(const dpose_pg *pg, const dpose_map *map, const BatchSpec &spec, cudaStream_t stream) {
  const size_t n = spec.n;
  double *d_out = spec.d_out;
  const int want_jacobian = spec.want_jacobian && !spec.cost_only;
  const bool hold_planes = spec.hold_planes;
  const int sms = device_info(pg->device).sms, smem_optin = device_info(pg->device).smem_optin;

  // rows kept per chunk
  const int need = std::min(32, pg->win_max);
  const int rows = need <= 20 ? 20 : need <= 28 ? 28 : 32;
  const int stored = DPOSE_TILE_COMPACT ? rows : rows + 7;
  // bytes of one replica of the shared-memory image
  const size_t table_bytes = sizeof(PatchCoef) * (size_t)pg->rows * pg->cols;
  const size_t budget = (size_t)smem_optin - 64;
  // preference: many threads first (the kernel is issue bound), then replicas (bank conflicts).
  // A/B builds: -DDPOSE_TILE_THREADS=512|768|1024, -DDPOSE_TILE_REPLICAS=1|4|8
  int threads = DPOSE_TILE_THREADS;
  int rep_shift = DPOSE_TILE_REPLICAS == 1 ? 0 : DPOSE_TILE_REPLICAS == 4 ? 2 : 3;  // 8 replicas if they fit (conflict-free LDS.128)
  auto fits = [&](int th, int rs) { return (rs < 0 ? 0 : table_bytes << rs) + (size_t)(stored + 1) * 4 * th <= budget; };  // + guard row
  // candidates in order of preference: replicas first at full width, then two full replicas on a narrower CTA (shared
  // memory only: the texture path's latency is what bounds the single-replica mode), then the texture modes
  {
    const int want_shift = rep_shift, want_threads = threads;
    const int cand[][2] = {{want_threads, want_shift}, {want_threads, 2}, {want_threads, 1}, {768, 1}, {want_threads, 0}, {768, 0}, {512, 0}};
    bool found = false;
    for (const auto &cd : cand) {
      if (cd[0] > want_threads || cd[1] > want_shift) continue;
#ifdef DPOSE_TILE_NO_REP2
      if (cd[1] == 1) continue;
#endif
      if (fits(cd[0], cd[1])) {
        threads = cd[0], rep_shift = cd[1], found = true;
        break;
      }
    }
    if (!found) rep_shift = -1, threads = want_threads;  // not even once: both halves through the texture unit
  }
  // small batches are latency bound: spread them over all SMs with smaller CTAs (the replica count, and with it
  // the cached coefficient image, stays the one chosen for the largest CTA)
  {
    const size_t per_sm = (n + (size_t)sms - 1) / (size_t)sms;
    const int choices[5] = {128, 256, 512, 768, 1024};
    for (int c : choices)
      if ((size_t)c >= per_sm || c >= threads) {
        threads = std::min(threads, c);
        break;
      }
  }

  dpose_pg *p = const_cast<dpose_pg *>(pg);
  {
    std::lock_guard<std::mutex> lock(map_state_mutex());
    if (p->coef_image_shift != rep_shift || (rep_shift >= 0 && !p->d_coef_image)) {
      if (p->d_coef_image) {
        DPOSE_CUDA(cudaDeviceSynchronize());
        DPOSE_CUDA(cudaFree(p->d_coef_image));
        p->d_coef_image = nullptr;
      }
      if (rep_shift >= 0) DPOSE_CUDA(cudaMalloc(&p->d_coef_image, table_bytes << rep_shift));
      const bool tex_mode = rep_shift <= 0;  // first halves twice in the image (or nowhere), both halves behind a texture
      if (tex_mode && !p->d_coef_hi) {
        const size_t hi_bytes = 2 * sizeof(double2) * (size_t)pg->rows * pg->cols;
        DPOSE_CUDA(cudaMalloc(&p->d_coef_hi, hi_bytes));
        cudaResourceDesc res = {};
        res.resType = cudaResourceTypeLinear;
        res.res.linear.devPtr = p->d_coef_hi;
        res.res.linear.desc = cudaCreateChannelDesc<int4>();
        res.res.linear.sizeInBytes = hi_bytes;
        cudaTextureDesc td = {};
        td.readMode = cudaReadModeElementType;
        cudaTextureObject_t tex = 0;
        DPOSE_CUDA(cudaCreateTextureObject(&tex, &res, &td, nullptr));
        p->tex_hi = (unsigned long long)tex;
      }
      k_build_coef_image<<<64, 256, 0, stream>>>(pg->d_coef, reinterpret_cast<double2 *>(p->d_coef_image),
                                                 tex_mode ? reinterpret_cast<double2 *>(p->d_coef_hi) : nullptr, pg->rows,
                                                 pg->cols, tex_mode ? 1 : rep_shift, rep_shift == 1);
      DPOSE_LAUNCHED();
      DPOSE_CUDA(cudaStreamSynchronize(stream));  // other streams may use the image right after
      p->coef_image_shift = rep_shift;
    }
  }
  DPOSE_TRY(tileplane_ensure(map, stream, hold_planes, spec.row_lo, spec.row_hi));

  TileArgs a;
  a.coef_image = reinterpret_cast<const uint4 *>(p->d_coef_image);
  a.coef_bytes = rep_shift < 0 ? 0u : (uint32_t)(table_bytes << rep_shift);
  a.g = TableGeom{pg->rows, pg->cols, (double)pg->center[0], (double)pg->center[1]};
  a.size_x = (int)map->size_x;
  a.size_y = (int)(spec.row_hi ? std::min(spec.row_hi, map->size_y) : map->size_y);  // the band of rows that count
  a.y_lo = (int)spec.row_lo;
  a.tiles = map->d_tiles;
  a.ntx = map->tiles_ntx;
  a.nty = map->tiles_nty;
  a.poses = spec.d_poses;
  a.out = d_out;
  a.cost_only = spec.cost_only;
  a.n = n;
  a.tex_hi = (cudaTextureObject_t)p->tex_hi;
  a.planes = reinterpret_cast<const int4 *>(p->d_coef_hi);
  a.out_aligned = (reinterpret_cast<uintptr_t>(d_out) & 31u) == 0;
  a.clip_thr = 0.5f / (float)std::max(pg->win_max, 25);
  a.clip_margin = clip_margin_for(pg->win_max);
  const size_t smem_bytes = (size_t)a.coef_bytes + (size_t)(stored + 1) * 4 * threads;
  const size_t want = (n + threads - 1) / threads;
  int grid = (int)std::min<size_t>(want, (size_t)sms);
  if (n >= (size_t)sms * 32) grid = sms;  // every SM gets a CTA
  a.per_cta = ((n + (size_t)grid - 1) / (size_t)grid + 31) / 32 * 32;  // equal contiguous shares, whole groups
  return want_jacobian ? launch_tile1<true>(a, threads, rep_shift, rows, grid, smem_bytes, stream)
                       : launch_tile1<false>(a, threads, rep_shift, rows, grid, smem_bytes, stream);
}

int get_cost_batch_spec(const dpose_pg *pg, const dpose_map *map, const BatchSpec &spec, cudaStream_t stream) {
  if (spec.n == 0) return DPOSE_OK;
  if (pg->device != map->device) return fail(DPOSE_EINVAL, "pose_gradient and map live on different devices");
  if (spec.row_hi != 0 && (spec.row_lo >= spec.row_hi || spec.row_hi > map->size_y))
    return fail(DPOSE_EINVAL, "row band [%u, %u) is empty or outside the map (%u rows)", spec.row_lo, spec.row_hi, map->size_y);
  DeviceGuard guard(pg->device);
  if (!guard.ok) return fail(DPOSE_ECUDA, "cudaSetDevice(%d) failed", pg->device);
  if (!map->d_tiles || pg->rows < 4 || pg->cols < 4) {  // empty map / no valid patch: every sum is empty
    DPOSE_CUDA(cudaMemsetAsync(spec.d_out, 0, sizeof(double) * (spec.cost_only ? 1 : 4) * spec.n, stream));
    return DPOSE_OK;
  }
  const size_t call_n = spec.call_n ? spec.call_n : spec.n;  // the kernel follows the size of the whole call
  if (!spec.d_poses) {
    // a pose lattice: k_sweep_poses writes a slab of poses (<= 2^20, 24 MB, L2 resident) that the batch kernel launched
    // right behind it consumes — the generation stays out of the hot kernels (measured: generating in-kernel cost the
    // thread-per-pose kernel 4 % through register pressure), nothing crosses PCIe, and the slab never reaches HBM hot
    constexpr size_t kSlab = (size_t)1 << 20;
    double *d_slab = nullptr;
    DPOSE_CUDA(cudaMallocAsync(&d_slab, sizeof(double) * 3 * std::min(spec.n, kSlab), stream));
    int rc = DPOSE_OK;
    for (size_t off = 0; off < spec.n && rc == DPOSE_OK; off += kSlab) {
      BatchSpec part = spec;
      part.n = std::min(kSlab, spec.n - off);
      part.sweep_base = spec.sweep_base + off;
      part.d_poses = d_slab;
      part.d_out = spec.d_out + (spec.cost_only ? 1 : 4) * off;
      part.hold_planes = spec.hold_planes || off > 0;
      k_sweep_poses<<<(unsigned)((part.n + 255) / 256), 256, 0, stream>>>(spec.sweep, part.sweep_base, part.n, d_slab);
      g_launches.fetch_add(1, std::memory_order_relaxed);
      rc = coop_batch_wanted(pg, call_n) ? get_cost_batch_coop(pg, map, part, stream) : get_cost_batch_tile(pg, map, part, stream);
    }
    cudaFreeAsync(d_slab, stream);
    return rc;
  }
  if (coop_batch_wanted(pg, call_n)) return get_cost_batch_coop(pg, map, spec, stream);
  return get_cost_batch_tile(pg, map, spec, stream);
}

int get_cost_batch_device(const dpose_pg *pg, const dpose_map *map, const double *d_poses, size_t n, double *d_out,
                          int want_jacobian, cudaStream_t stream, bool hold_planes, uint32_t row_lo, uint32_t row_hi) {
  BatchSpec spec;
  spec.row_lo = row_lo;
  spec.row_hi = row_hi;
  spec.d_poses = d_poses;
  spec.n = n;
  spec.d_out = d_out;
  spec.want_jacobian = want_jacobian;
  spec.hold_planes = hold_planes;
  return get_cost_batch_spec(pg, map, spec, stream);
}

}  // namespace dpose
