// problem.cu — batched evaluation of dpose_goal_tolerance's NLP.
//
// Upstream `problem` (dpose_goal_tolerance.cpp:90-292) evaluates, for ONE displacement x per Ipopt
// callback: f = pose_gradient::get_cost(goal + x, lethal_cells_in_search_box) + pose regularisers,
// grad f, the constraints g = (x^2 + y^2, theta^2) and their Jacobian.  Here the same quantities are
// evaluated for n displacements per call (multi-start in lock-step) with everything on the device:
//   init       : the search box (:119-138) is cropped out of the device costmap into a small map of its own —
//                "the lethal cells inside the box" (:141) become "all lethal cells of the crop", which is
//                exactly what the batched get_cost kernel sums (it clips windows at the map border) —
//                and the regularisers (:144-180) are set up on the host (a dozen flops).
//   eval_batch : k_problem_poses (x -> goal + x in crop coordinates), the batched get_cost kernel,
//                k_problem_finish (regularisers, g, jac_g).
#include <algorithm>
#include <cmath>
#include <cstring>
#include <mutex>
#include <new>

#include "problem.cuh"

namespace dpose {
namespace {

struct FinishArgs {
  const double *x;     // n x 3 displacements
  const double *cost4; // n x 4 from the batched get_cost
  size_t n;
  double pose[3];
  int n_regs;
  Reg regs[2];
  double *f, *grad, *g, *jac;  // any may be null
};

__global__ void k_problem_poses(const double *__restrict__ x, size_t n, double ox, double oy, double ot,
                                double *__restrict__ poses) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  poses[3 * i] = x[3 * i] + ox;  // (goal - crop origin) + x: cost depends on cell - pose only
  poses[3 * i + 1] = x[3 * i + 1] + oy;
  poses[3 * i + 2] = x[3 * i + 2] + ot;
}

__global__ void k_problem_finish(FinishArgs a) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.n) return;
  const double x[3] = {a.x[3 * i], a.x[3 * i + 1], a.x[3 * i + 2]};
  const double p[3] = {x[0] + a.pose[0], x[1] + a.pose[1], x[2] + a.pose[2]};  // :98-100
  double cost = a.cost4[4 * i];
  double J[3] = {a.cost4[4 * i + 1], a.cost4[4 * i + 2], a.cost4[4 * i + 3]};
  for (int r = 0; r < a.n_regs; ++r) {  // pose_regularization::get_cost (:76-88), summed as in :106-109
    double diff[3] = {p[0] - a.regs[r].goal[0], p[1] - a.regs[r].goal[1], p[2] - a.regs[r].goal[2]};
    diff[2] = normalize_angle(diff[2]);
    double rc = 0.0;
    for (int k = 0; k < 3; ++k) {
      J[k] += 2 * (a.regs[r].norm[k] * diff[k]);
      rc += a.regs[r].norm[k] * (diff[k] * diff[k]);
    }
    cost += rc;
  }
  if (a.f) a.f[i] = cost;
  if (a.grad)
    for (int k = 0; k < 3; ++k) a.grad[3 * i + k] = J[k];
  if (a.g) {  // :238-247
    a.g[2 * i] = x[0] * x[0] + x[1] * x[1];
    a.g[2 * i + 1] = x[2] * x[2];
  }
  if (a.jac)  // :270-279
    for (int k = 0; k < 3; ++k) a.jac[3 * i + k] = x[k] * 2;
}

// pose_regularization::init (:62-74)
void reg_init(Reg *r, double w_lin_unscaled, double w_rot_unscaled, const double start[3], const double goal[3]) {
  double w_lin = (start[0] - goal[0]) * (start[0] - goal[0]) + (start[1] - goal[1]) * (start[1] - goal[1]);
  double w_rot = std::pow(normalize_angle(goal[2] - start[2]), 2);  // shortest_angular_distance(start, goal)
  w_lin = w_lin_unscaled / std::max(w_lin, 1e-3);
  w_rot = w_rot_unscaled / std::max(w_rot, 1e-3);
  for (int i = 0; i < 3; ++i) r->goal[i] = goal[i];
  r->norm[0] = w_lin, r->norm[1] = w_lin, r->norm[2] = w_rot;
}

}  // namespace
}  // namespace dpose

using namespace dpose;

static int problem_reserve(dpose_problem *p, size_t n) {
  if (!p->stream) DPOSE_CUDA(cudaStreamCreateWithFlags(&p->stream, cudaStreamNonBlocking));
  if (n <= p->cap) return DPOSE_OK;
  cudaFree(p->d_x);
  cudaFreeHost(p->h_stage);
  p->d_x = p->h_stage = nullptr;
  p->cap = 0;
  const size_t cap = n + n / 2 + 256;
  DPOSE_CUDA(cudaMalloc(&p->d_x, sizeof(double) * 19 * cap));
  DPOSE_CUDA(cudaHostAlloc(&p->h_stage, sizeof(double) * 12 * cap, cudaHostAllocMapped));
  p->d_work = p->d_x + 3 * cap;  // 7 n: cost4 | poses
  p->d_res = p->d_work + 7 * cap;
  p->cap = cap;
  return DPOSE_OK;
}

extern "C" {

int dpose_problem_create(const dpose_pg *pg, const dpose_map *map, dpose_problem **out) {
  if (!out) return fail(DPOSE_EINVAL, "out is NULL");
  *out = nullptr;
  if (!pg || !map) return fail(DPOSE_EINVAL, "NULL argument");
  if (pg->device != map->device) return fail(DPOSE_EINVAL, "pose_gradient and map live on different devices");
  dpose_problem *p = new (std::nothrow) dpose_problem();
  if (!p) return fail(DPOSE_ENOMEM, "out of host memory");
  p->pg = pg;
  p->map = map;
  *out = p;
  return DPOSE_OK;
}

void dpose_problem_destroy(dpose_problem *p) {
  if (!p) return;
  {
    DeviceGuard guard(p->pg->device);
    cudaFree(p->d_x);
    cudaFreeHost(p->h_stage);
    cudaFree(p->d_solve);
    cudaFreeHost(p->h_solve);
    cudaFree(p->d_trace);
    if (p->solve_ev0) cudaEventDestroy(p->solve_ev0);
    if (p->solve_ev1) cudaEventDestroy(p->solve_ev1);
    if (p->stream) cudaStreamDestroy(p->stream);
  }
  dpose_map_destroy(p->crop);
  delete p;
}

int dpose_problem_init(dpose_problem *p, const double start[3], const double goal[3], const dpose_gt_config *cfg) {
  if (!p || !start || !goal || !cfg) return fail(DPOSE_EINVAL, "NULL argument");
  std::lock_guard<std::mutex> lock(p->mu);
  p->ready = false;
  for (int i = 0; i < 3; ++i) p->pose[i] = goal[i];  // :114
  p->lin_tol_sq = std::pow(cfg->lin_tolerance, 2);   // :115-116
  p->rot_tol_sq = std::pow(cfg->rot_tolerance, 2);
  p->lin_tol = std::abs(cfg->lin_tolerance);
  p->rot_tol = std::abs(cfg->rot_tolerance);
  int max_coeff = 0;  // :122-123
  for (int i = 0; i < 10; ++i) max_coeff = std::max(max_coeff, std::abs(p->pg->box[i]));
  // :126-133: h is a double, but the ring it is written into is a rectangle<int>: Eigen's comma initialiser converts
  // every coefficient to int (truncation toward zero) BEFORE the goal is added and the sum rounded (:134-138)
  const double h_real = max_coeff + std::abs(cfg->lin_tolerance);
  if (!(h_real < 1e9)) return fail(DPOSE_EINVAL, "lin_tolerance %g out of range", cfg->lin_tolerance);
  const double h = (double)(int)h_real;
  const double bx[5] = {-h, h, h, -h, -h}, by[5] = {-h, -h, h, h, -h};
  for (int i = 0; i < 5; ++i) {  // :129-138 (Eigen round = std::round)
    p->search_box[2 * i] = (int32_t)std::round(bx[i] + goal[0]);
    p->search_box[2 * i + 1] = (int32_t)std::round(by[i] + goal[1]);
  }
  // lethal_cells_within(search_box) (:141) clamps the ring into the map (dpose_costmap.cpp:121-129); the cells
  // it returns are all lethal cells of the clamped rectangle = all lethal cells of this crop
  dpose_map_destroy(p->crop);
  p->crop = nullptr;
  if (p->map->size_x && p->map->size_y) {
    int x0 = INT32_MAX, y0 = INT32_MAX, x1 = INT32_MIN, y1 = INT32_MIN;
    for (int i = 0; i < 5; ++i) {
      const int x = std::min(std::max(p->search_box[2 * i], 0), (int)p->map->size_x - 1);
      const int y = std::min(std::max(p->search_box[2 * i + 1], 0), (int)p->map->size_y - 1);
      x0 = std::min(x0, x), x1 = std::max(x1, x), y0 = std::min(y0, y), y1 = std::max(y1, y);
    }
    p->origin[0] = x0;
    p->origin[1] = y0;
    DPOSE_TRY(dpose_map_create_from_device(p->map->d_data + (size_t)y0 * p->map->pitch + x0, (uint32_t)(x1 - x0 + 1),
                                           (uint32_t)(y1 - y0 + 1), p->map->pitch, p->map->device, &p->crop));
  }
  // regularisers (:144-180)
  p->n_regs = 0;
  if (cfg->weight_goal_lin != 0 || cfg->weight_goal_rot != 0) {  // :145-150
    const double s[3] = {goal[0] + cfg->lin_tolerance, goal[1], goal[2] + cfg->rot_tolerance};
    reg_init(&p->regs[p->n_regs++], cfg->weight_goal_lin, cfg->weight_goal_rot, s, goal);
  }
  if (cfg->weight_start_lin != 0 || cfg->weight_start_rot != 0) {  // :152-180
    double s[3] = {start[0], start[1], start[2]};
    double diff[3] = {start[0] - goal[0], start[1] - goal[1], start[2] - goal[2]};
    diff[2] = normalize_angle(diff[2]);
    const double lin_norm = std::sqrt(diff[0] * diff[0] + diff[1] * diff[1]), rot_norm = std::abs(diff[2]);
    if (lin_norm != 0 && lin_norm > cfg->lin_tolerance)  // :170-171
      for (int i = 0; i < 3; ++i) s[i] = goal[i] + diff[i] * cfg->lin_tolerance / lin_norm;
    if (rot_norm != 0 && rot_norm > cfg->rot_tolerance)  // :174-177
      s[2] = goal[2] + diff[2] * cfg->rot_tolerance / rot_norm;
    else
      s[2] = start[2];
    reg_init(&p->regs[p->n_regs++], cfg->weight_start_lin, cfg->weight_start_rot, goal, s);  // :179
  }
  p->ready = true;
  return DPOSE_OK;
}

int dpose_problem_info(const dpose_problem *p, int32_t search_box[10], double g_upper[2]) {
  if (!p) return fail(DPOSE_EINVAL, "problem is NULL");
  if (search_box) std::memcpy(search_box, p->search_box, sizeof(p->search_box));
  if (g_upper) {  // get_bounds_info (:196-207): 0 <= g <= (lin_tol^2, rot_tol^2)
    g_upper[0] = p->lin_tol_sq;
    g_upper[1] = p->rot_tol_sq;
  }
  return DPOSE_OK;
}

// device pointers in, device pointers out; asynchronous on `stream`.  d_work: 7 n doubles of scratch.
static int problem_eval_device(const dpose_problem *p, const double *d_x, size_t n, double *d_work, double *d_f,
                               double *d_grad, double *d_g, double *d_jac, cudaStream_t stream) {
  double *d_cost4 = d_work, *d_poses = d_work + 4 * n;
  const unsigned grid = (unsigned)((n + 255) / 256);
  if (p->crop) {
    k_problem_poses<<<grid, 256, 0, stream>>>(d_x, n, p->pose[0] - p->origin[0], p->pose[1] - p->origin[1], p->pose[2],
                                              d_poses);
    DPOSE_LAUNCHED();
    DPOSE_TRY(get_cost_batch_device(p->pg, p->crop, d_poses, n, d_cost4, 1, stream));
  } else {
    DPOSE_CUDA(cudaMemsetAsync(d_cost4, 0, sizeof(double) * 4 * n, stream));
  }
  FinishArgs a;
  a.x = d_x;
  a.cost4 = d_cost4;
  a.n = n;
  for (int i = 0; i < 3; ++i) a.pose[i] = p->pose[i];
  a.n_regs = p->n_regs;
  a.regs[0] = p->regs[0];
  a.regs[1] = p->regs[1];
  a.f = d_f, a.grad = d_grad, a.g = d_g, a.jac = d_jac;
  k_problem_finish<<<grid, 256, 0, stream>>>(a);
  DPOSE_LAUNCHED();
  return DPOSE_OK;
}

constexpr size_t kZeroCopyDisplacements = 8192;

int dpose_problem_eval_batch(dpose_problem *p, const double *x, size_t n, double *f, double *grad_f, double *g,
                             double *jac_g) {
  if (!p || (n && !x)) return fail(DPOSE_EINVAL, "NULL argument");
  if (!p->ready) return fail(DPOSE_EINVAL, "dpose_problem_init has not been called");
  if (n == 0) return DPOSE_OK;
  DeviceGuard guard(p->pg->device);
  if (!guard.ok) return fail(DPOSE_ECUDA, "cudaSetDevice(%d) failed", p->pg->device);
  std::lock_guard<std::mutex> lock(p->mu);
  DPOSE_TRY(problem_reserve(p, n));
  cudaStream_t st = p->stream;
  double *d_f = p->d_res, *d_grad = d_f + n, *d_g = d_grad + 3 * n, *d_jac = d_g + 2 * n;
  auto body = [&]() -> int {
    // pinned staging: one H2D of x, one D2H of [f | grad | g | jac]
    double *h_x = p->h_stage, *h_res = p->h_stage + 3 * n;
    std::memcpy(h_x, x, sizeof(double) * 3 * n);
    if (n <= kZeroCopyDisplacements) {
      // a solver iteration's worth of displacements: the kernels read x and write [f | grad | g | jac] in the mapped
      // pinned staging itself, no copies in either direction (the writes overlap the evaluation)
      DPOSE_TRY(problem_eval_device(p, h_x, n, p->d_work, f ? h_res : nullptr, grad_f ? h_res + n : nullptr,
                                    g ? h_res + 4 * n : nullptr, jac_g ? h_res + 6 * n : nullptr, st));
    } else {
      DPOSE_CUDA(cudaMemcpyAsync(p->d_x, h_x, sizeof(double) * 3 * n, cudaMemcpyHostToDevice, st));
      DPOSE_TRY(problem_eval_device(p, p->d_x, n, p->d_work, f ? d_f : nullptr, grad_f ? d_grad : nullptr,
                                    g ? d_g : nullptr, jac_g ? d_jac : nullptr, st));
      DPOSE_CUDA(cudaMemcpyAsync(h_res, p->d_res, sizeof(double) * 9 * n, cudaMemcpyDeviceToHost, st));
    }
    DPOSE_CUDA(cudaStreamSynchronize(st));
    if (f) std::memcpy(f, h_res, sizeof(double) * n);
    if (grad_f) std::memcpy(grad_f, h_res + n, sizeof(double) * 3 * n);
    if (g) std::memcpy(g, h_res + 4 * n, sizeof(double) * 2 * n);
    if (jac_g) std::memcpy(jac_g, h_res + 6 * n, sizeof(double) * 3 * n);
    return DPOSE_OK;
  };
  const int rc = body();
  if (rc != DPOSE_OK) cudaStreamSynchronize(st);
  return rc;
}

int dpose_problem_eval_batch_device(const dpose_problem *p, const double *d_x, size_t n, double *d_work, double *d_f,
                                    double *d_grad_f, double *d_g, double *d_jac_g, void *stream) {
  if (!p || (n && (!d_x || !d_work))) return fail(DPOSE_EINVAL, "NULL argument");
  if (!p->ready) return fail(DPOSE_EINVAL, "dpose_problem_init has not been called");
  if (n == 0) return DPOSE_OK;
  DeviceGuard guard(p->pg->device);
  if (!guard.ok) return fail(DPOSE_ECUDA, "cudaSetDevice(%d) failed", p->pg->device);
  return problem_eval_device(p, d_x, n, d_work, d_f, d_grad_f, d_g, d_jac_g, (cudaStream_t)stream);
}

}  // extern "C"
