// dpose_solve.hpp — optional C++ surface over the batched goal-tolerance solve (dpose_b200.h: dpose_problem_*,
// dpose_problem_solve_batch, dpose_sample_offsets).  Not an upstream header: upstream keeps `problem` and the attempt
// loop inside the plugin (dpose_goal_tolerance/src/dpose_goal_tolerance.cpp:90-292, :376-459); this is what a
// maintainer calls from DposeGoalTolerance::preProcess to try all starting offsets at once.
#ifndef DPOSE_CORE__DPOSE_SOLVE__HPP
#define DPOSE_CORE__DPOSE_SOLVE__HPP

#include <dpose_core/dpose_core.hpp>

#include <cstdint>
#include <vector>

namespace dpose_core {

/// dpose_goal_tolerance::problem (dpose_goal_tolerance.cpp:90-292) for many displacements at once, on the device
class multi_start_problem {
public:
  /// problem::problem (:90-94); both arguments must outlive this object
  multi_start_problem(const pose_gradient& _pg, const device_costmap& _map) {
    dpose_problem* p = nullptr;
    detail::check(dpose_problem_create(_pg.get_data().handle(), _map.handle(), &p));
    p_.reset(p, [](dpose_problem* q) { dpose_problem_destroy(q); });
  }

  /// problem::init (:112-181); poses in cells / radians, lin_tolerance in cells (:530-535)
  void
  init(const pose_gradient::pose& _start, const pose_gradient::pose& _goal, const dpose_gt_config& _cfg) {
    const double s[3] = {_start(0), _start(1), _start(2)}, g[3] = {_goal(0), _goal(1), _goal(2)};
    detail::check(dpose_problem_init(p_.get(), s, g, &_cfg));
    cfg_ = _cfg;
  }

  struct result {
    std::vector<double> x, f;        ///< final displacement (n x 3) and cost per start
    std::vector<uint8_t> status;     ///< DPOSE_SOLVE_* flags per start
    std::vector<int32_t> evals;
    dpose_solve_summary summary;     ///< first / best accepted start and its pose (problem::get_pose, :283-291)
    bool accepted() const { return summary.first >= 0; }
  };

  /// the attempt loop of preProcess (:430-456) for _n_starts offsets from the plugin's sampler (:333-373) at once:
  /// offset 0 is the goal itself (:431), the others are drawn from std::mt19937(_seed)
  result
  solve(size_t _n_starts, uint32_t _seed, const dpose_solver_options* _opt = nullptr) {
    std::vector<double> x0(3 * _n_starts);
    detail::check(dpose_sample_offsets(_seed, cfg_.lin_tolerance, cfg_.rot_tolerance, _n_starts, 1, x0.data()));
    return solve(x0, _opt);
  }

  /// the same from caller-chosen starting displacements (n x 3)
  result
  solve(const std::vector<double>& _x0, const dpose_solver_options* _opt = nullptr) {
    const size_t n = _x0.size() / 3;
    result r;
    r.x.resize(3 * n), r.f.resize(n), r.status.resize(n), r.evals.resize(n);
    detail::check(dpose_problem_solve_batch(p_.get(), _x0.data(), n, _opt, r.x.data(), r.f.data(), r.status.data(),
                                            r.evals.data(), &r.summary, nullptr));
    return r;
  }

private:
  std::shared_ptr<dpose_problem> p_;
  dpose_gt_config cfg_{1, 1, 1, 1, 1, 1};
};

}  // namespace dpose_core

#endif  // DPOSE_CORE__DPOSE_SOLVE__HPP
