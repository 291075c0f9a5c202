// problem.cuh — dpose_goal_tolerance::problem state shared by problem.cu (batched NLP evaluation) and solve.cu (the
// device-resident multi-start solve).  Not part of the public surface.
#pragma once
#include <mutex>

#include "common.cuh"

namespace dpose {

// pose_regularization after init (dpose_goal_tolerance.cpp:62-74)
struct Reg {
  double goal[3], norm[3];
};

constexpr double kPi = 3.14159265358979323846;

// angles::normalize_angle (ROS `angles`): fmod(a + pi, 2 pi), result <= 0 ? + pi : - pi
__host__ __device__ inline double normalize_angle(double a) {
  const double r = fmod(a + kPi, 2.0 * kPi);
  return r <= 0.0 ? r + kPi : r - kPi;
}

}  // namespace dpose

struct dpose_problem {
  const dpose_pg *pg = nullptr;    // borrowed
  const dpose_map *map = nullptr;  // borrowed
  dpose_map *crop = nullptr;       // owned: the clamped search box as a map of its own
  bool ready = false;
  int32_t search_box[10] = {0};
  int32_t origin[2] = {0, 0};  // crop origin in map cells
  double pose[3] = {0, 0, 0};
  double lin_tol = 1, rot_tol = 1;  // |lin_tolerance|, |rot_tolerance| (the constraint set of :196-207)
  double lin_tol_sq = 1, rot_tol_sq = 1;
  int n_regs = 0;
  dpose::Reg regs[2];
  // workspace (grow-only), one call at a time per problem
  std::mutex mu;
  cudaStream_t stream = nullptr;
  size_t cap = 0;
  double *d_x = nullptr, *d_work = nullptr, *d_res = nullptr;  // 3n, 7n, 9n
  double *h_stage = nullptr;  // pinned, 12n: x in, then [f | grad | g | jac] out — one copy each way
  // solve.cu: device results + pinned staging of dpose_problem_solve_batch (grow-only)
  size_t solve_cap = 0, solve_trace_cap = 0;
  unsigned char *d_solve = nullptr, *h_solve = nullptr;
  double *d_trace = nullptr;
  cudaEvent_t solve_ev0 = nullptr, solve_ev1 = nullptr;
  float last_solve_us = 0.f;  // device time of the last solve: k_solve + check_footprint + k_solve_select
};
