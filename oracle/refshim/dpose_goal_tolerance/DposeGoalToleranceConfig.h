// dpose_goal_tolerance/DposeGoalToleranceConfig.h — stand-in (TEST INFRASTRUCTURE ONLY) for the header catkin generates
// from dpose_goal_tolerance/cfg/DposeGoalTolerance.cfg: six doubles, default 1.
#ifndef DPOSE_REFSHIM_GOAL_TOLERANCE_CONFIG
#define DPOSE_REFSHIM_GOAL_TOLERANCE_CONFIG
namespace dpose_goal_tolerance {
struct DposeGoalToleranceConfig {
  double lin_tolerance = 1, rot_tolerance = 1;
  double weight_goal_lin = 1, weight_goal_rot = 1, weight_start_lin = 1, weight_start_rot = 1;
};
}  // namespace dpose_goal_tolerance
#endif
