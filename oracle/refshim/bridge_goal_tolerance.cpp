// bridge_goal_tolerance.cpp — extern "C" access to the reference's dpose_goal_tolerance::problem, for oracle/_ref.
//
// TEST INFRASTRUCTURE ONLY (see README.md in this directory).  This translation unit INCLUDES the reference's
// dpose_goal_tolerance.cpp from /root/reference where it lies (nothing is copied; -I points there) and is compiled
// against the stand-in ROS / Ipopt headers next to it — included rather than linked so that the helpers the plugin
// defines only inside its .cpp (uniform_pose_distribution, :333-357) can be driven too.  There is no Ipopt here: what
// is exposed is the NLP the plugin hands to Ipopt — problem::init and the TNLP callbacks eval_f / eval_grad_f /
// eval_g / eval_jac_g / get_bounds_info (dpose_goal_tolerance.cpp:90-292) —, the plugin's two pose conversions
// (to_cell / to_msg, :293-330), its sampler of starting offsets, and a CPU driver that runs the product's step rule
// (restated in oracle/dpose_oracle.c) on those callbacks followed by the reference's check_footprint: the reference arm
// of bench.py --workload cfg4.
#include <dpose_goal_tolerance.cpp>  // the reference's own source file, unmodified

#include <dpose_core/dpose_costmap.hpp>

#include "../dpose_oracle.h"

#include <tf2/utils.h>

#include <cmath>
#include <memory>
#include <random>
#ifdef _OPENMP
#include <omp.h>
#endif
#include <stdexcept>
#include <vector>

namespace dpose_goal_tolerance {
// defined with external linkage in the reference's dpose_goal_tolerance.cpp:293-330
pose to_cell(const DposeGoalTolerance::Pose& _pose, const DposeGoalTolerance::Map& _map);
DposeGoalTolerance::Pose to_msg(const pose& _pose, const DposeGoalTolerance::Map& _map);
}  // namespace dpose_goal_tolerance

namespace {

struct holder {
  holder(const uint8_t* map, uint32_t sx, uint32_t sy, const int32_t* fp, int n, unsigned padding) :
      cm(const_cast<unsigned char*>(map), sx, sy, 1.0), lc(&cm, points(fp, n)), footprint(2, n) {
    for (int i = 0; i < n; ++i) footprint(0, i) = fp[2 * i], footprint(1, i) = fp[2 * i + 1];
    dpose_core::pose_gradient::parameter param;
    param.padding = padding;
    p.reset(new dpose_goal_tolerance::problem(lc, param));
  }
  // resolution 1: make_footprint (dpose_costmap.cpp:40-55) rounds x / 1.0, so the cells go through unchanged
  static std::vector<geometry_msgs::Point> points(const int32_t* fp, int n) {
    std::vector<geometry_msgs::Point> out((size_t)n);
    for (int i = 0; i < n; ++i) out[i].x = fp[2 * i], out[i].y = fp[2 * i + 1];
    return out;
  }
  costmap_2d::Costmap2D cm;
  costmap_2d::LayeredCostmap lc;
  dpose_core::polygon footprint;
  std::unique_ptr<dpose_goal_tolerance::problem> p;
};

}  // namespace

extern "C" {

void*
ref_problem_create(const uint8_t* map, uint32_t sx, uint32_t sy, const int32_t* fp_xy, int n, unsigned padding) {
  try {
    return new holder(map, sx, sy, fp_xy, n, padding);
  } catch (const std::exception&) {
    return nullptr;
  }
}

void
ref_problem_free(void* h) {
  delete static_cast<holder*>(h);
}

/// problem::init (dpose_goal_tolerance.cpp:112-181); cfg = lin_tolerance, rot_tolerance, weight_goal_lin,
/// weight_goal_rot, weight_start_lin, weight_start_rot.  -1 if the reference throws (both weights of a regulariser 0)
int
ref_problem_init(void* h, const double start[3], const double goal[3], const double cfg[6]) {
  dpose_goal_tolerance::DposeGoalToleranceConfig c;
  c.lin_tolerance = cfg[0], c.rot_tolerance = cfg[1];
  c.weight_goal_lin = cfg[2], c.weight_goal_rot = cfg[3], c.weight_start_lin = cfg[4], c.weight_start_rot = cfg[5];
  try {
    static_cast<holder*>(h)->p->init(dpose_goal_tolerance::pose(start[0], start[1], start[2]),
                                     dpose_goal_tolerance::pose(goal[0], goal[1], goal[2]), c);
    return 0;
  } catch (const std::exception&) {
    return -1;
  }
}

/// one Ipopt iterate: eval_f, eval_grad_f, eval_g, eval_jac_g (values) at x, the first with new_x = true as Ipopt does
void
ref_problem_eval(void* h, const double x[3], double* f, double grad_f[3], double g[2], double jac_g[3]) {
  auto& p = *static_cast<holder*>(h)->p;
  p.eval_f(3, x, true, *f);
  p.eval_grad_f(3, x, false, grad_f);
  p.eval_g(3, x, false, 2, g);
  p.eval_jac_g(3, x, false, 2, 3, nullptr, nullptr, jac_g);
}

void
ref_problem_eval_batch(void* h, const double* x, size_t n, double* f, double* grad_f, double* g, double* jac_g) {
  for (size_t i = 0; i < n; ++i) ref_problem_eval(h, x + 3 * i, f + i, grad_f + 3 * i, g + 2 * i, jac_g + 3 * i);
}

/// uniform_pose_distribution (:333-357) driven by std::mt19937(seed): n offsets (radius, angle, yaw drawn in the
/// reference's order)
void
ref_sample_offsets(uint32_t seed, double lin_tol, double rot_tol, size_t n, double* out) {
  dpose_goal_tolerance::DposeGoalToleranceConfig c;
  c.lin_tolerance = lin_tol, c.rot_tolerance = rot_tol;
  dpose_goal_tolerance::uniform_pose_distribution dis(c);
  std::mt19937 gen(seed);
  for (size_t i = 0; i < n; ++i) {
    const dpose_goal_tolerance::pose p = dis(gen);
    out[3 * i] = p.x(), out[3 * i + 1] = p.y(), out[3 * i + 2] = p.z();
  }
}

/// preProcessImpl's validation (:386-399) of a solved displacement: footprint moved by T(goal + x) R(theta), rounded,
/// then the reference's check_footprint against the whole costmap
static bool
footprint_free(holder& h, const double goal[3], const double x[3], uint8_t lethal) {
  const dpose_goal_tolerance::pose g(goal[0] + x[0], goal[1] + x[1], goal[2] + x[2]);
  Eigen::Isometry2d trans(Eigen::Translation2d(g.x(), g.y()) * Eigen::Rotation2Dd(g.z()));
  const dpose_core::polygon moved = (trans * h.footprint.cast<double>()).array().round().cast<int>().matrix();
  return dpose_core::check_footprint(h.cm, moved, lethal);
}

/// CPU arm of the multi-start solve: the step rule of oracle_solve on the compiled plugin's TNLP callbacks, OpenMP over
/// starts with one `problem` per thread (the callbacks cache cost_ / J_, :96-110), then the validation above.
/// status: the product's encoding (low 2 bits status, 0x4 footprint free, 0x8 accepted).
void
ref_solve_batch(void** holders, int n_holders, const double goal[3], const oracle_solver_cfg* cfg, uint8_t lethal,
                const double* x0, size_t n, double* x, double* f, uint8_t* status, int32_t* evals, double* trace) {
#pragma omp parallel for schedule(dynamic, 16) num_threads(n_holders)
  for (long long i = 0; i < (long long)n; ++i) {
#ifdef _OPENMP
    holder* h = static_cast<holder*>(holders[omp_get_thread_num()]);
#else
    holder* h = static_cast<holder*>(holders[0]);
#endif
    int ev = 0;
    int st = oracle_solve(cfg, ref_problem_eval, h, x0 + 3 * i, x + 3 * i, f + i, &ev,
                          trace ? trace + (size_t)i * 12 * (size_t)(cfg->max_iter + 1) : nullptr);
    const bool free_ = footprint_free(*h, goal, x + 3 * i, lethal);
    status[i] = (uint8_t)(st | (free_ ? 0x4 : 0) | (st == 1 && free_ ? 0x8 : 0));
    if (evals) evals[i] = ev;
  }
}

/// get_nlp_info + get_bounds_info + the sparsity pattern of eval_jac_g (:183-207, :249-268)
void
ref_problem_info(void* h, int dims[4], double x_l[3], double x_u[3], double g_l[2], double g_u[2], int rows[3], int cols[3]) {
  auto& p = *static_cast<holder*>(h)->p;
  Ipopt::Index n, m, nnz_jac, nnz_h = -1;
  Ipopt::TNLP::IndexStyleEnum style;
  p.get_nlp_info(n, m, nnz_jac, nnz_h, style);
  dims[0] = n, dims[1] = m, dims[2] = nnz_jac, dims[3] = (int)style;
  p.get_bounds_info(n, x_l, x_u, m, g_l, g_u);
  const double x[3] = {0, 0, 0};
  p.eval_jac_g(n, x, false, m, nnz_jac, rows, cols, nullptr);
}

/// to_cell (:293-309) on a costmap of the given geometry; 0, or -1 if the reference throws
int
ref_to_cell(uint32_t sx, uint32_t sy, double resolution, double origin_x, double origin_y, const double world[3], double cell[3]) {
  costmap_2d::LayeredCostmap lc("map", false, false);
  lc.resizeMap(sx, sy, resolution, origin_x, origin_y);
  costmap_2d::Costmap2DROS ros_map(&lc, "map");
  geometry_msgs::PoseStamped msg;
  msg.header.frame_id = "map";
  msg.pose.position.x = world[0], msg.pose.position.y = world[1];
  tf2::Quaternion q;
  q.setRPY(0, 0, world[2]);
  msg.pose.orientation = tf2::toMsg(q);
  try {
    const auto c = dpose_goal_tolerance::to_cell(msg, ros_map);
    cell[0] = c.x(), cell[1] = c.y(), cell[2] = c.z();
    return 0;
  } catch (const std::runtime_error&) {
    return -1;
  }
}

/// to_msg (:311-330): cell pose -> metric position and yaw; 0, or -1 if the reference throws
int
ref_to_msg(uint32_t sx, uint32_t sy, double resolution, double origin_x, double origin_y, const double cell[3], double world[3]) {
  costmap_2d::LayeredCostmap lc("map", false, false);
  lc.resizeMap(sx, sy, resolution, origin_x, origin_y);
  costmap_2d::Costmap2DROS ros_map(&lc, "map");
  try {
    const auto msg = dpose_goal_tolerance::to_msg(dpose_goal_tolerance::pose(cell[0], cell[1], cell[2]), ros_map);
    world[0] = msg.pose.position.x, world[1] = msg.pose.position.y, world[2] = tf2::getYaw(msg.pose.orientation);
    return 0;
  } catch (const std::runtime_error&) {
    return -1;
  }
}

}  // extern "C"
