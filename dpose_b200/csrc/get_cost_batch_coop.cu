// get_cost_batch_coop.cu — batched get_cost(poses[]) for batches too small to fill a B200 with one thread per pose:
// G lanes per pose (coop_eval.cuh), G chosen per launch so that the batch covers the GPU's lanes once.
//
// The thread-per-pose kernel needs >= 148 x 1024 poses to reach its throughput; a solver iteration's 4096 candidates
// would occupy one warp per SM and wait ~40 serial cell evaluations per warp.  With G = 32 the same batch runs as
// 4096 warps, each evaluation ~1/20 of that latency.  The results differ from the thread-per-pose kernel's only by the
// summation order (<= 1e-13 relative); each kernel is deterministic and independent of the batch position.
#include <algorithm>

#include "coop_eval.cuh"

namespace dpose {
namespace {

constexpr int kCoopThreads = 128;

struct CoopBatchArgs {
  CoopCtx ctx;
  const double *poses;
  double *out;
  size_t n;
  int out_aligned;
  int cost_only;  // out holds n costs instead of n x 4
};

template <int G, bool kJac>
__global__ void __launch_bounds__(kCoopThreads) k_cost_batch_coop(CoopBatchArgs a) {
  constexpr int K = CoopK<G>::value;
  __shared__ uint32_t s_mask[K * kCoopThreads];
  const int tid = threadIdx.x, lane = tid & 31, g = lane & (G - 1);
  const unsigned gmask = G == 32 ? 0xffffffffu : (((1u << G) - 1u) << (lane & ~(G - 1)));
  constexpr int kPerCta = kCoopThreads / G;
  // programmatic dependent launch: started while the predecessor in the stream (k_pack_tiles, or the kernel that wrote
  // the poses) may still be running; nothing of its output is touched before this point
  asm volatile("griddepcontrol.wait;" ::: "memory");
  for (size_t i = (size_t)blockIdx.x * kPerCta + tid / G; i < a.n; i += (size_t)gridDim.x * kPerCta) {
    const double *pp = a.poses + 3 * i;
    const double tx = __ldg(pp), ty = __ldg(pp + 1), th = __ldg(pp + 2);
    double r[4];
    coop_eval<G, kJac>(a.ctx, tx, ty, th, g, gmask, s_mask + tid, kCoopThreads, r);
    if (g == 0 && a.cost_only) {
      a.out[i] = r[0];
    } else if (g == 0) {
      double *o = a.out + 4 * i;
      if (a.out_aligned)
        asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(o), "d"(r[0]), "d"(r[1]), "d"(r[2]), "d"(r[3]) : "memory");
      else
        o[0] = r[0], o[1] = r[1], o[2] = r[2], o[3] = r[3];
    }
  }
}

template <int G>
int launch_coop(const CoopBatchArgs &a, int want_jacobian, int sms, cudaStream_t stream) {
  const size_t per_cta = kCoopThreads / G;
  const size_t want = (a.n + per_cta - 1) / per_cta;
  const unsigned grid = (unsigned)std::min<size_t>(want, (size_t)sms * 16);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kCoopThreads);
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;  // see griddepcontrol.wait in the kernel
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (want_jacobian)
    DPOSE_CUDA(cudaLaunchKernelEx(&cfg, k_cost_batch_coop<G, true>, a));
  else
    DPOSE_CUDA(cudaLaunchKernelEx(&cfg, k_cost_batch_coop<G, false>, a));
  DPOSE_LAUNCHED();
  return DPOSE_OK;
}

// lanes one launch should cover: the SMs' resident threads at this kernel's register footprint (1024 per SM)
#ifndef DPOSE_COOP_LANES_PER_SM
#define DPOSE_COOP_LANES_PER_SM 1024
#endif
int coop_group(size_t n, int sms) {
  const size_t lanes = (size_t)sms * DPOSE_COOP_LANES_PER_SM;
  int G = 32;
  while (G > 1 && n * (size_t)G > lanes) G >>= 1;
  return G;
}

}  // namespace

CoopCtx coop_ctx(const dpose_pg *pg, const dpose_map *map) {
  CoopCtx c;
  c.g = TableGeom{pg->rows, pg->cols, (double)pg->center[0], (double)pg->center[1]};
  c.size_x = (int)map->size_x;
  c.size_y = (int)map->size_y;
  c.tiles = map->d_tiles;
  c.ntx = map->tiles_ntx;
  c.nty = map->tiles_nty;
  c.coef_abs = pg->d_coef_abs;
  c.clip_thr = 0.5 / (double)std::max(pg->win_max, 25);
  return c;
}

bool coop_batch_wanted(const dpose_pg *pg, size_t n) {
  return coop_group(n, device_info(pg->device).sms) >= DPOSE_COOP_MIN_GROUP;
}

int get_cost_batch_coop(const dpose_pg *pg, const dpose_map *map, const BatchSpec &spec, cudaStream_t stream) {
  const size_t n = spec.n;
  const int want_jacobian = spec.want_jacobian && !spec.cost_only;
  const int sms = device_info(pg->device).sms;
  DPOSE_TRY(coef_abs_ensure(pg, stream));
  DPOSE_TRY(tileplane_ensure(map, stream, spec.hold_planes, spec.row_lo, spec.row_hi));
  CoopBatchArgs a;
  a.ctx = coop_ctx(pg, map);
  if (spec.row_hi) a.ctx.size_y = (int)std::min(spec.row_hi, map->size_y), a.ctx.y_lo = (int)spec.row_lo;  // the band of rows that count
  a.poses = spec.d_poses;
  a.out = spec.d_out;
  a.n = n;
  a.out_aligned = (reinterpret_cast<uintptr_t>(spec.d_out) & 31u) == 0;
  a.cost_only = spec.cost_only;
  // only G = 32 is launched (DPOSE_COOP_MIN_GROUP, common.cuh); an A/B build with a smaller minimum gets 16 / 8 too
  switch (coop_group(n, sms)) {
#if DPOSE_COOP_MIN_GROUP <= 16
    case 16: return launch_coop<16>(a, want_jacobian, sms, stream);
    case 8: return launch_coop<8>(a, want_jacobian, sms, stream);
#endif
    default: return launch_coop<32>(a, want_jacobian, sms, stream);
  }
}

}  // namespace dpose
