"""dpose_b200 — B200-native footprint-cost path of dpose_core.

Python mirror of the reference's dpose_core interface (same names, argument meaning and
error behaviour) on top of the C ABI in include/dpose_b200.h.  The compute runs in
hand-written CUDA kernels (dpose_b200/csrc); there is no CPU fallback.
"""
from ._lib import DposeError, LIB_PATH  # noqa: F401
from .core import (Costmap, check_footprint, check_footprint_batch, cell_vector_device, cost_data, device_count, launch_count,  # noqa: F401
                   lethal_cells_within, make_footprint, pose_gradient, problem, raytrace, sample_offsets, solver_options,
                   to_cell, to_msg, to_rectangle, get_cost_batch_multi, sweep_poses, order_poses_by_tile,
                   SOLVE_ACCEPTED, SOLVE_CONVERGED, SOLVE_FOOTPRINT_FREE, SOLVE_MAXITER, SOLVE_STALLED, SOLVE_STATUS_MASK)

LETHAL_OBSTACLE = 254
