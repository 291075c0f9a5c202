// solve.cu — the goal-tolerance solve, batched over multi-start displacements and resident on the device
// (upstream: DposeGoalTolerance::preProcess / preProcessImpl, dpose_goal_tolerance.cpp:376-459).
//
// Upstream runs, per attempt (at most `attempts_`, :430-456): Ipopt on the NLP `problem` from one starting offset
// (:378), then check_footprint of the solved pose (:386-399), and takes the first attempt that passes.  Ipopt calls
// the four TNLP callbacks once per iterate, one pose at a time (:96-110, :215-281).  Here ALL starts run at once:
//
//   k_solve<G>          one launch for the whole solve.  G lanes per start (G = 32 for a few thousand starts): every
//                       evaluation of f / grad f is a cooperative get_cost over the lethal tile plane of the search-box
//                       crop (coop_eval.cuh) plus the pose regularisers (:76-88); between evaluations every lane of the
//                       group advances the same scalar state machine (below).  The iterates never leave the registers:
//                       no PCIe, no launch and no grid-wide synchronisation per iteration — starts are independent.
//   k_check_footprint*  the plugin's validation (:386-399) of every final pose against the FULL costmap
//   k_solve_select      accepted = converged and footprint free; the first accepted start (upstream's attempt order) and
//                       the accepted start of lowest cost
// and the results come back with ONE device-to-host copy (one host-to-device copy brings the starting offsets).
//
// The solver.  Ipopt is not reproduced (an interior-point method with a filter line search is not a per-thread
// algorithm, and upstream runs it with tol = 5 / max_iter = 20, :504-507, i.e. as a coarse descent).  The constraint
// set of the NLP — g0 = x^2 + y^2 <= lin_tol^2, g1 = theta^2 <= rot_tol^2 (:196-207, :238-247) — is a disc times an
// interval, so its Euclidean projection is closed-form and the method is a projected quasi-Newton descent:
//   direction   d = -H grad f, H a dense 3 x 3 BFGS inverse-Hessian approximation (Ipopt is configured with
//               hessian_approximation = limited-memory, :513: it never sees second derivatives either); H0 =
//               diag(1, 1, 1 / rho^2) with rho the footprint radius in cells, so a step in theta is measured by the
//               distance the footprint's rim moves
//   trial       x_t = P(x + alpha d), accepted if f(x_t) <= f(x) + c1 grad f . (x_t - x)  (Armijo on the projected arc);
//               otherwise alpha shrinks by safeguarded quadratic interpolation.  EVERY trial costs one evaluation and the
//               budget counts evaluations: at most max_iter after the first ("one batched evaluation per iteration")
//   stop        converged when || x - P(x - grad f) ||_inf <= min(tol, dual_inf_tol) (the KKT residual of the projected
//               problem; Ipopt's own test needs the unscaled dual infeasibility under dual_inf_tol = 1 as well)
// The evaluated functions are the reference's (parity-tested per iterate against the compiled plugin's callbacks); the
// step rule is restated independently in the test infrastructure and must be reproduced bit for bit
// (tests/test_solver.py): this file is compiled with -fmad=false and every product-then-sum below is evaluated as
// written — do not reorder the arithmetic of project / advance without changing the restatement.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <new>
#include <random>

#include "coop_eval.cuh"
#include "problem.cuh"

namespace dpose {
namespace {

constexpr int kSolveThreads = 128;
constexpr int kTraceStride = 12;  // doubles per evaluation: x_t[3], f, grad[3], g0, g1, alpha, status after the step, 0

enum : int { kRunning = 0, kConverged = 1, kMaxIter = 2, kStalled = 3 };

struct SolveArgs {
  CoopCtx ctx;
  int has_map;
  const double *x0;  // n x 3
  size_t n;
  double pose[3];  // goal (:114)
  double ox, oy;   // goal - crop origin: the crop's coordinates of the goal
  int n_regs;
  Reg regs[2];
  double lin_tol, rot_tol, rho2;
  int max_iter;
  double kkt_tol, c1;
  // results
  double *x, *f, *final_pose;  // n x 3, n, n x 3
  int32_t *evals;
  uint8_t *status;
  double *trace;  // n x (max_iter + 1) x kTraceStride or null
};

struct State {
  double x[3], f, g[3];
  double H[6];  // symmetric: 00 01 02 11 12 22
  double d[3], xt[3];
  double alpha, gd;
  int n_acc, evals, status;
};

__device__ __forceinline__ void project(double p[3], double lin_tol, double rot_tol) {
  const double r2 = p[0] * p[0] + p[1] * p[1];
  if (r2 > lin_tol * lin_tol) {
    const double sc = lin_tol / sqrt(r2);
    p[0] = p[0] * sc;
    p[1] = p[1] * sc;
  }
  p[2] = fmin(fmax(p[2], -rot_tol), rot_tol);
}

__device__ __forceinline__ void set_h0(State &s, double gamma, double rho2) {
  s.H[0] = gamma, s.H[1] = 0.0, s.H[2] = 0.0, s.H[3] = gamma, s.H[4] = 0.0, s.H[5] = gamma / rho2;
}

__device__ __forceinline__ void direction(State &s) {  // d = -(H g)
  s.d[0] = -(s.H[0] * s.g[0] + s.H[1] * s.g[1] + s.H[2] * s.g[2]);
  s.d[1] = -(s.H[1] * s.g[0] + s.H[3] * s.g[1] + s.H[4] * s.g[2]);
  s.d[2] = -(s.H[2] * s.g[0] + s.H[4] * s.g[1] + s.H[5] * s.g[2]);
}

__device__ __forceinline__ double first_alpha(const State &s, double rho2) {  // first move <= one cell at the rim
  const double nd = sqrt(s.d[0] * s.d[0] + s.d[1] * s.d[1] + rho2 * (s.d[2] * s.d[2]));
  return nd > 1.0 ? 1.0 / nd : 1.0;
}

// the trial point of the current direction / step; false if it is not a usable descent step
__device__ __forceinline__ bool make_trial(State &s, const SolveArgs &a) {
  for (int k = 0; k < 3; ++k) s.xt[k] = s.x[k] + s.alpha * s.d[k];
  project(s.xt, a.lin_tol, a.rot_tol);
  const double e0 = s.xt[0] - s.x[0], e1 = s.xt[1] - s.x[1], e2 = s.xt[2] - s.x[2];
  s.gd = s.g[0] * e0 + s.g[1] * e1 + s.g[2] * e2;
  const double step = fmax(fmax(fabs(e0), fabs(e1)), fabs(e2));
  return s.gd < 0.0 && step > 1e-12;
}

// after the evaluation of s.xt gave (ft, gt): accept or backtrack, then set up the next trial (or stop)
__device__ __forceinline__ void advance(State &s, const SolveArgs &a, double ft, const double gt[3]) {
  s.evals += 1;
  bool fresh = true;  // a new direction (as opposed to a shorter step along the old one)
  if (s.evals == 1) {
    for (int k = 0; k < 3; ++k) s.x[k] = s.xt[k], s.g[k] = gt[k];
    s.f = ft;
    set_h0(s, 1.0, a.rho2);
    s.n_acc = 0;
  } else if (ft <= s.f + a.c1 * s.gd) {
    const double sx[3] = {s.xt[0] - s.x[0], s.xt[1] - s.x[1], s.xt[2] - s.x[2]};
    const double y[3] = {gt[0] - s.g[0], gt[1] - s.g[1], gt[2] - s.g[2]};
    const double sy = sx[0] * y[0] + sx[1] * y[1] + sx[2] * y[2];
    const double ss = sx[0] * sx[0] + sx[1] * sx[1] + sx[2] * sx[2];
    const double yy = y[0] * y[0] + y[1] * y[1] + y[2] * y[2];
    if (sy > 1e-10 * (sqrt(ss) * sqrt(yy))) {  // curvature condition; otherwise H is kept
      if (s.n_acc == 0) set_h0(s, sy / (y[0] * y[0] + y[1] * y[1] + (y[2] * y[2]) / a.rho2), a.rho2);
      const double rho = 1.0 / sy;
      const double Hy[3] = {s.H[0] * y[0] + s.H[1] * y[1] + s.H[2] * y[2], s.H[1] * y[0] + s.H[3] * y[1] + s.H[4] * y[2],
                            s.H[2] * y[0] + s.H[4] * y[1] + s.H[5] * y[2]};
      const double yHy = y[0] * Hy[0] + y[1] * Hy[1] + y[2] * Hy[2];
      const double cf = (1.0 + rho * yHy) * rho;
      s.H[0] = s.H[0] + cf * (sx[0] * sx[0]) - rho * (Hy[0] * sx[0] + sx[0] * Hy[0]);
      s.H[1] = s.H[1] + cf * (sx[0] * sx[1]) - rho * (Hy[0] * sx[1] + sx[0] * Hy[1]);
      s.H[2] = s.H[2] + cf * (sx[0] * sx[2]) - rho * (Hy[0] * sx[2] + sx[0] * Hy[2]);
      s.H[3] = s.H[3] + cf * (sx[1] * sx[1]) - rho * (Hy[1] * sx[1] + sx[1] * Hy[1]);
      s.H[4] = s.H[4] + cf * (sx[1] * sx[2]) - rho * (Hy[1] * sx[2] + sx[1] * Hy[2]);
      s.H[5] = s.H[5] + cf * (sx[2] * sx[2]) - rho * (Hy[2] * sx[2] + sx[2] * Hy[2]);
      s.n_acc += 1;
    }
    for (int k = 0; k < 3; ++k) s.x[k] = s.xt[k], s.g[k] = gt[k];
    s.f = ft;
  } else {  // Armijo failed (also for a NaN cost): shorter step, minimiser of the quadratic through f, gd, ft
    double t = -s.gd / (2.0 * ((ft - s.f) - s.gd));
    if (!(t >= 0.1)) t = 0.1;
    if (t > 0.5) t = 0.5;
    s.alpha = s.alpha * t;
    fresh = false;
  }
  if (fresh) {
    double p[3] = {s.x[0] - s.g[0], s.x[1] - s.g[1], s.x[2] - s.g[2]};
    project(p, a.lin_tol, a.rot_tol);
    const double kkt = fmax(fmax(fabs(s.x[0] - p[0]), fabs(s.x[1] - p[1])), fabs(s.x[2] - p[2]));
    if (kkt <= a.kkt_tol) {
      s.status = kConverged;
      return;
    }
    direction(s);
    s.alpha = s.n_acc == 0 ? first_alpha(s, a.rho2) : 1.0;
  }
  if (s.evals > a.max_iter) {
    s.status = kMaxIter;
    return;
  }
  if (make_trial(s, a)) return;
  // not a descent step: fall back to the scaled gradient once (a quasi-Newton direction bent by the projection)
  if (fresh && s.n_acc > 0) {
    set_h0(s, 1.0, a.rho2);
    s.n_acc = 0;
    direction(s);
    s.alpha = first_alpha(s, a.rho2);
    if (make_trial(s, a)) return;
  }
  s.status = kStalled;
}

// Occupancy: uncapped the kernel takes 126 registers (4 CTAs = 512 lanes per SM), and a group size chosen for 1024 lanes
// per SM put 4096 starts x 32 lanes into 1.7 waves.  Capped at 96 registers (5 CTAs = 640 lanes per SM; the spilled
// words are solver state touched once per evaluation) with G chosen for the lanes that are really resident, 4096 starts
// run as ONE wave of 16-lane groups: 156 -> 119 us of device time per solve (profiles/r2_ab_tile_walk.txt, runs v30 / v31).
#ifndef DPOSE_SOLVE_MIN_CTAS
#define DPOSE_SOLVE_MIN_CTAS 5
#endif
#ifndef DPOSE_SOLVE_LANES_PER_SM
#define DPOSE_SOLVE_LANES_PER_SM (DPOSE_SOLVE_MIN_CTAS * 128)
#endif
template <int G>
__global__ void __launch_bounds__(kSolveThreads, DPOSE_SOLVE_MIN_CTAS) k_solve(SolveArgs a) {
  constexpr int K = CoopK<G>::value;
  __shared__ uint32_t s_mask[K * kSolveThreads];
  const int tid = threadIdx.x, lane = tid & 31, g = lane & (G - 1);
  const unsigned gmask = G == 32 ? 0xffffffffu : (((1u << G) - 1u) << (lane & ~(G - 1)));
  const size_t i = ((size_t)blockIdx.x * kSolveThreads + tid) / G;
  if (i >= a.n) return;  // whole groups leave together
  State s;
  for (int k = 0; k < 3; ++k) s.xt[k] = __ldg(a.x0 + 3 * i + k);
  project(s.xt, a.lin_tol, a.rot_tol);
  for (int k = 0; k < 3; ++k) s.x[k] = s.xt[k], s.g[k] = 0.0, s.d[k] = 0.0;
  s.f = 0.0, s.alpha = 1.0, s.gd = 0.0, s.n_acc = 0, s.evals = 0, s.status = kRunning;
  set_h0(s, 1.0, a.rho2);
  while (s.status == kRunning) {
    // eval_f / eval_grad_f at the displacement s.xt (dpose_goal_tolerance.cpp:96-110)
    double r[4] = {0.0, 0.0, 0.0, 0.0};
    if (a.has_map)
      coop_eval<G, true>(a.ctx, s.xt[0] + a.ox, s.xt[1] + a.oy, s.xt[2] + a.pose[2], g, gmask, s_mask + tid, kSolveThreads, r);
    const double p[3] = {s.xt[0] + a.pose[0], s.xt[1] + a.pose[1], s.xt[2] + a.pose[2]};  // :98-100
    double cost = r[0];
    double J[3] = {r[1], r[2], r[3]};
    for (int q = 0; q < a.n_regs; ++q) {  // pose_regularization::get_cost (:76-88), summed as in :106-109
      double diff[3] = {p[0] - a.regs[q].goal[0], p[1] - a.regs[q].goal[1], p[2] - a.regs[q].goal[2]};
      diff[2] = normalize_angle(diff[2]);
      double rc = 0.0;
      for (int k = 0; k < 3; ++k) {
        J[k] += 2 * (a.regs[q].norm[k] * diff[k]);
        rc += a.regs[q].norm[k] * (diff[k] * diff[k]);
      }
      cost += rc;
    }
    const double alpha_used = s.alpha;
    const double xt[3] = {s.xt[0], s.xt[1], s.xt[2]};
    const int slot = s.evals;
    advance(s, a, cost, J);
    if (a.trace && g == 0) {
      double *t = a.trace + ((size_t)i * (size_t)(a.max_iter + 1) + (size_t)slot) * kTraceStride;
      t[0] = xt[0], t[1] = xt[1], t[2] = xt[2], t[3] = cost, t[4] = J[0], t[5] = J[1], t[6] = J[2];
      t[7] = xt[0] * xt[0] + xt[1] * xt[1];  // eval_g (:238-247)
      t[8] = xt[2] * xt[2];
      t[9] = alpha_used, t[10] = (double)s.status, t[11] = 0.0;
    }
  }
  if (g == 0) {
    for (int k = 0; k < 3; ++k) {
      a.x[3 * i + k] = s.x[k];
      a.final_pose[3 * i + k] = s.x[k] + a.pose[k];  // finalize_solution: pose_ += x (:283-291)
    }
    a.f[i] = s.f;
    a.evals[i] = s.evals;
    a.status[i] = (uint8_t)s.status;
  }
}

// accepted = converged and footprint free (dpose_goal_tolerance.cpp:381-399); one CTA
__global__ void __launch_bounds__(1024) k_solve_select(const double *__restrict__ f, const double *__restrict__ x,
                                                       const double *__restrict__ final_pose, uint8_t *status,
                                                       const uint8_t *__restrict__ ok, size_t n,
                                                       dpose_solve_summary *summary) {
  __shared__ unsigned long long s_first[32], s_conv[32], s_acc[32];
  __shared__ double s_bf[32];
  __shared__ unsigned long long s_bi[32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const unsigned long long kNone = ~0ull;
  unsigned long long first = kNone, bi = kNone, n_conv = 0, n_acc = 0;
  double bf = 0.0;
  for (size_t i = tid; i < n; i += 1024) {
    const bool conv = (status[i] & DPOSE_SOLVE_STATUS_MASK) == kConverged;
    const bool free_ = ok[i] != 0;
    status[i] = (uint8_t)((status[i] & DPOSE_SOLVE_STATUS_MASK) | (free_ ? DPOSE_SOLVE_FOOTPRINT_FREE : 0) |
                          (conv && free_ ? DPOSE_SOLVE_ACCEPTED : 0));
    n_conv += conv;
    if (conv && free_) {
      n_acc += 1;
      if (first == kNone) first = i;  // i ascends within a thread
      const double fi = f[i];
      if (bi == kNone || fi < bf) bf = fi, bi = i;  // ties keep the lower index
    }
  }
  auto better = [](double fa, unsigned long long ia, double fb, unsigned long long ib) {
    return ib == ~0ull ? true : ia == ~0ull ? false : (fa < fb || (fa == fb && ia < ib));
  };
  for (int o = 16; o > 0; o >>= 1) {
    const unsigned long long of = __shfl_xor_sync(0xffffffffu, first, o);
    first = of < first ? of : first;
    n_conv += __shfl_xor_sync(0xffffffffu, n_conv, o);
    n_acc += __shfl_xor_sync(0xffffffffu, n_acc, o);
    const double obf = __shfl_xor_sync(0xffffffffu, bf, o);
    const unsigned long long obi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (!better(bf, bi, obf, obi)) bf = obf, bi = obi;
  }
  if (lane == 0) s_first[warp] = first, s_conv[warp] = n_conv, s_acc[warp] = n_acc, s_bf[warp] = bf, s_bi[warp] = bi;
  __syncthreads();
  if (tid == 0) {
    for (int w = 1; w < 32; ++w) {
      first = s_first[w] < first ? s_first[w] : first;
      n_conv += s_conv[w];
      n_acc += s_acc[w];
      if (!better(bf, bi, s_bf[w], s_bi[w])) bf = s_bf[w], bi = s_bi[w];
    }
    summary->first = first == kNone ? -1 : (long long)first;
    summary->best = bi == kNone ? -1 : (long long)bi;
    summary->n_converged = n_conv;
    summary->n_accepted = n_acc;
    summary->best_f = bi == kNone ? 0.0 : bf;
    for (int k = 0; k < 3; ++k) {
      summary->best_pose[k] = bi == kNone ? 0.0 : final_pose[3 * bi + k];
      summary->first_pose[k] = first == kNone ? 0.0 : final_pose[3 * first + k];
    }
  }
}

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// device / pinned result block of one solve: [x 3n | f n | final_pose 3n] doubles, evals n int32, status n, ok n,
// summary — one contiguous D2H
struct Layout {
  size_t x, f, pose, evals, status, ok, summary, total;
  explicit Layout(size_t n) {
    x = 0;
    f = x + sizeof(double) * 3 * n;
    pose = f + sizeof(double) * n;
    evals = pose + sizeof(double) * 3 * n;
    status = evals + sizeof(int32_t) * n;
    ok = status + n;
    summary = align_up(ok + n, 16);
    total = summary + sizeof(dpose_solve_summary);
  }
};

int solve_group(size_t n, int sms) {
  int G = 32;
  while (G > 2 && n * (size_t)G > (size_t)sms * DPOSE_SOLVE_LANES_PER_SM) G >>= 1;
  return G;
}

template <int G>
void launch_solve(const SolveArgs &a, cudaStream_t st) {
  const size_t threads = a.n * G;
  k_solve<G><<<(unsigned)((threads + kSolveThreads - 1) / kSolveThreads), kSolveThreads, 0, st>>>(a);
}

}  // namespace
}  // namespace dpose

using namespace dpose;

extern "C" {

void dpose_solver_default_options(dpose_solver_options *opt) {
  if (!opt) return;
  opt->max_iter = 20;       // dpose_goal_tolerance.cpp:507
  opt->tol = 5.0;           // :504
  opt->dual_inf_tol = 1.0;  // Ipopt's default
  opt->armijo_c1 = 1e-4;
  opt->lethal = kLethal;    // :395-396
}

// uniform_pose_generator (dpose_goal_tolerance.cpp:333-373): radius ~ U[0, lin_tol), angle ~ U[-pi, pi), yaw ~ U[0, rot_tol)
// drawn in this order from one std::mt19937 through std::uniform_real_distribution<double> — the standard library's
// own classes, so the stream equals the plugin's for the same seed (upstream seeds from std::random_device).
int dpose_sample_offsets(uint32_t seed, double lin_tolerance, double rot_tolerance, size_t n, int first_zero, double *x0) {
  if (n && !x0) return fail(DPOSE_EINVAL, "x0 is NULL");
  if (!(lin_tolerance >= 0) || !(rot_tolerance >= 0) || !std::isfinite(lin_tolerance) || !std::isfinite(rot_tolerance))
    return fail(DPOSE_EINVAL, "tolerances must be finite and non-negative");
  std::mt19937 gen(seed);
  std::uniform_real_distribution<double> dis_rad(0, lin_tolerance), dis_ang(-M_PI, M_PI), dis_yaw(0, rot_tolerance);
  for (size_t i = 0; i < n; ++i) {
    if (i == 0 && first_zero) {  // attempt 0 starts at the goal itself (:431)
      x0[0] = x0[1] = x0[2] = 0.0;
      continue;
    }
    const double radius = dis_rad(gen);
    const double angle = dis_ang(gen);
    const double yaw = dis_yaw(gen);
    x0[3 * i] = std::cos(angle) * radius;
    x0[3 * i + 1] = std::sin(angle) * radius;
    x0[3 * i + 2] = yaw;
  }
  return DPOSE_OK;
}

int dpose_problem_solve_us(const dpose_problem *p, float *us) {
  if (!p || !us) return fail(DPOSE_EINVAL, "NULL argument");
  *us = p->last_solve_us;
  return DPOSE_OK;
}

int dpose_problem_solve_batch(dpose_problem *p, const double *x0, size_t n, const dpose_solver_options *opt_in, double *x,
                              double *f, uint8_t *status, int32_t *evals, dpose_solve_summary *summary, double *trace) {
  if (!p || (n && !x0)) return fail(DPOSE_EINVAL, "NULL argument");
  if (!p->ready) return fail(DPOSE_EINVAL, "dpose_problem_init has not been called");
  dpose_solver_options opt;
  if (opt_in)
    opt = *opt_in;
  else
    dpose_solver_default_options(&opt);
  if (opt.max_iter < 0 || opt.max_iter > 10000 || !(opt.tol > 0) || !(opt.dual_inf_tol > 0) || !(opt.armijo_c1 > 0) ||
      !(opt.armijo_c1 < 1))
    return fail(DPOSE_EINVAL, "bad solver options");
  if (n > ((size_t)1 << 26)) return fail(DPOSE_EINVAL, "too many starts");
  if (summary) {
    std::memset(summary, 0, sizeof(*summary));
    summary->first = summary->best = -1;
  }
  if (n == 0) return DPOSE_OK;
  const dpose_pg *pg = p->pg;
  DeviceGuard guard(pg->device);
  if (!guard.ok) return fail(DPOSE_ECUDA, "cudaSetDevice(%d) failed", pg->device);
  std::lock_guard<std::mutex> lock(p->mu);
  if (!p->stream) DPOSE_CUDA(cudaStreamCreateWithFlags(&p->stream, cudaStreamNonBlocking));
  if (!p->solve_ev0) {
    DPOSE_CUDA(cudaEventCreate(&p->solve_ev0));
    DPOSE_CUDA(cudaEventCreate(&p->solve_ev1));
  }
  cudaStream_t st = p->stream;
  const Layout lay(n);
  const size_t x0_bytes = sizeof(double) * 3 * n;
  if (n > p->solve_cap) {
    cudaFree(p->d_solve);
    cudaFreeHost(p->h_solve);
    p->d_solve = p->h_solve = nullptr;
    p->solve_cap = 0;
    const size_t cap = n + n / 2 + 256;
    const Layout lc(cap);
    DPOSE_CUDA(cudaMalloc(&p->d_solve, lc.total + sizeof(double) * 3 * cap));
    DPOSE_CUDA(cudaHostAlloc(&p->h_solve, lc.total + sizeof(double) * 3 * cap, cudaHostAllocDefault));
    p->solve_cap = cap;
  }
  const size_t trace_doubles = trace ? n * (size_t)(opt.max_iter + 1) * kTraceStride : 0;
  if (trace_doubles > p->solve_trace_cap) {
    cudaFree(p->d_trace);
    p->d_trace = nullptr;
    p->solve_trace_cap = 0;
    DPOSE_CUDA(cudaMalloc(&p->d_trace, sizeof(double) * trace_doubles));
    p->solve_trace_cap = trace_doubles;
  }
  // the block holds the results first (the part that travels back) and the starting offsets behind them
  unsigned char *d_base = p->d_solve, *h_base = p->h_solve;
  double *d_x0 = reinterpret_cast<double *>(d_base + align_up(lay.total, 16));
  double *h_x0 = reinterpret_cast<double *>(h_base + align_up(lay.total, 16));

  const int sms = device_info(pg->device).sms;
  SolveArgs a;
  a.has_map = p->crop && p->crop->d_tiles && pg->rows >= 4 && pg->cols >= 4;
  auto body = [&]() -> int {
    std::memcpy(h_x0, x0, x0_bytes);
    DPOSE_CUDA(cudaMemcpyAsync(d_x0, h_x0, x0_bytes, cudaMemcpyHostToDevice, st));  // the one H2D of the solve
    if (a.has_map) {
      DPOSE_TRY(coef_abs_ensure(pg, st));
      DPOSE_TRY(tileplane_ensure(p->crop, st, false));
      a.ctx = coop_ctx(pg, p->crop);
    } else {
      std::memset(&a.ctx, 0, sizeof(a.ctx));
    }
    DPOSE_CUDA(cudaEventRecord(p->solve_ev0, st));
    a.x0 = d_x0;
    a.n = n;
    for (int k = 0; k < 3; ++k) a.pose[k] = p->pose[k];
    a.ox = p->pose[0] - p->origin[0];
    a.oy = p->pose[1] - p->origin[1];
    a.n_regs = p->n_regs;
    a.regs[0] = p->regs[0];
    a.regs[1] = p->regs[1];
    a.lin_tol = p->lin_tol;
    a.rot_tol = p->rot_tol;
    // footprint radius in cells: the largest distance of a box corner from the centre cell
    double rho2 = 1.0;
    for (int k = 0; k < 5; ++k)
      rho2 = std::max(rho2, (double)pg->box[2 * k] * pg->box[2 * k] + (double)pg->box[2 * k + 1] * pg->box[2 * k + 1]);
    a.rho2 = rho2;
    a.max_iter = opt.max_iter;
    a.kkt_tol = std::min(opt.tol, opt.dual_inf_tol);
    a.c1 = opt.armijo_c1;
    a.x = reinterpret_cast<double *>(d_base + lay.x);
    a.f = reinterpret_cast<double *>(d_base + lay.f);
    a.final_pose = reinterpret_cast<double *>(d_base + lay.pose);
    a.evals = reinterpret_cast<int32_t *>(d_base + lay.evals);
    a.status = d_base + lay.status;
    a.trace = trace ? p->d_trace : nullptr;
    switch (solve_group(n, sms)) {
      case 32: launch_solve<32>(a, st); break;
      case 16: launch_solve<16>(a, st); break;
      case 8: launch_solve<8>(a, st); break;
      case 4: launch_solve<4>(a, st); break;
      default: launch_solve<2>(a, st); break;
    }
    DPOSE_LAUNCHED();
    // the plugin's validation of the solved pose (:386-399) against the full costmap, every start
    uint8_t *d_ok = d_base + lay.ok;
    DPOSE_TRY(dpose_check_footprint_batch_device(p->map, pg->footprint.data(), (int)(pg->footprint.size() / 2), a.final_pose,
                                                 n, opt.lethal, d_ok, st));
    k_solve_select<<<1, 1024, 0, st>>>(a.f, a.x, a.final_pose, a.status, d_ok, n,
                                       reinterpret_cast<dpose_solve_summary *>(d_base + lay.summary));
    DPOSE_LAUNCHED();
    DPOSE_CUDA(cudaEventRecord(p->solve_ev1, st));
    DPOSE_CUDA(cudaMemcpyAsync(h_base, d_base, lay.total, cudaMemcpyDeviceToHost, st));  // the one D2H of the solve
    if (trace) DPOSE_CUDA(cudaMemcpyAsync(trace, p->d_trace, sizeof(double) * trace_doubles, cudaMemcpyDeviceToHost, st));
    DPOSE_CUDA(cudaStreamSynchronize(st));
    cudaEventElapsedTime(&p->last_solve_us, p->solve_ev0, p->solve_ev1);
    p->last_solve_us *= 1000.f;
    if (x) std::memcpy(x, h_base + lay.x, sizeof(double) * 3 * n);
    if (f) std::memcpy(f, h_base + lay.f, sizeof(double) * n);
    if (evals) std::memcpy(evals, h_base + lay.evals, sizeof(int32_t) * n);
    if (status) std::memcpy(status, h_base + lay.status, n);
    if (summary) std::memcpy(summary, h_base + lay.summary, sizeof(*summary));
    return DPOSE_OK;
  };
  const int rc = body();
  if (rc != DPOSE_OK) cudaStreamSynchronize(st);
  return rc;
}

}  // extern "C"
