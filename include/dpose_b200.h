/*
 * dpose_b200.h — C ABI of the B200-native footprint-cost path of dpose_core.
 *
 * Plain C, opaque handles, caller-owned buffers, int status codes; no exceptions, no
 * torch / Eigen / OpenCV types.  This is the boundary a maintainer of dorezyuk/dpose
 * binds instead of the CPU bodies cited per entry point (paths relative to the upstream
 * tree).  the headers under include/dpose_core/ are the reference-compatible C++ surface on top of it.
 *
 * Conventions
 *  - cells and polygons are int32 pairs "x0,y0,x1,y1,..." (Eigen::Matrix<int,2,N>
 *    column-major, dpose_core.hpp:72-76); rectangles are the closed 5-point ring
 *    (dpose_core.hpp:39-69), 10 int32.
 *  - poses are (x, y, theta) float64 in CELL units / radians (dpose_core.hpp:141-142).
 *  - costmaps are uint8, row-major, index = y * pitch + x (costmap_2d::Costmap2D::getIndex);
 *    only the value 254 (costmap_2d::LETHAL_OBSTACLE) is lethal.
 *  - every function returns DPOSE_OK (0) or an error code; dpose_last_error() gives the
 *    thread-local message.  There is no CPU fallback: without a CUDA device every
 *    compute entry point fails with DPOSE_ENODEV / DPOSE_ECUDA.
 */
#ifndef DPOSE_B200_H
#define DPOSE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DPOSE_OK 0
#define DPOSE_EINVAL 1 /* bad argument (e.g. footprint with < 3 vertices) */
#define DPOSE_ECUDA 2  /* CUDA runtime / driver error */
#define DPOSE_ENOMEM 3
#define DPOSE_ENODEV 4 /* no usable CUDA device */

typedef struct dpose_pg dpose_pg;       /* pose_gradient + its cost_data on one device */
typedef struct dpose_map dpose_map;     /* device-resident costmap */
typedef struct dpose_cells dpose_cells; /* device-resident cell_vector */
typedef struct dpose_problem dpose_problem; /* dpose_goal_tolerance::problem, batched */

const char *dpose_last_error(void);
int dpose_device_count(int *n);
/* number of CUDA kernels this library has launched in the calling process */
uint64_t dpose_launch_count(void);

/* pinned host memory for the host-buffer entry points (optional, pageable works) */
int dpose_host_alloc(void **ptr, size_t bytes);
void dpose_host_free(void *ptr);
/* frees buffers this library returned (dpose_lethal_cells_within) */
void dpose_free(void *ptr);

/* ---- footprint --------------------------------------------------------------------
 * replaces make_footprint (dpose_costmap.cpp:40-55): cells = round(metres / resolution).
 * DPOSE_EINVAL if resolution <= 0 (reference: std::runtime_error). Host arithmetic. */
int dpose_make_footprint(const double *xy_m, int n, double resolution, int32_t *xy_cells);

/* ---- cost_data / pose_gradient ------------------------------------------------------
 * replaces pose_gradient::pose_gradient (dpose_core.cpp:148-156) and cost_data::cost_data
 * + _get_cost (dpose_core.cpp:55-139): outline raster, even-odd fill, exact EDT,
 * scale/shift, 3x3 Gaussian blur and the per-patch bilinear coefficient tables, all on
 * `device`.  DPOSE_EINVAL if n < 3 (reference: "footprint must contain at least three
 * points").  */
int dpose_pg_create(const int32_t *footprint_xy, int n, uint32_t padding, int device,
                    dpose_pg **out);
void dpose_pg_destroy(dpose_pg *pg);
/* replaces cost_data::get_data / get_center / get_box (dpose_core.hpp:83-96).
 * *cost points at a host copy (rows*cols float32, row-major) owned by the handle. */
int dpose_pg_table(const dpose_pg *pg, const float **cost, int *rows, int *cols,
                   int32_t center[2], int32_t box[10]);
/* intermediate images for bit-exact checks (host copies owned by the handle):
 * squared EDT (int32), polylines outline and fillPoly mask (uint8, 0 = drawn),
 * table before the blur (float32).  Any out pointer may be NULL. */
int dpose_pg_stages(const dpose_pg *pg, const int32_t **edt_sq, const uint8_t **outline,
                    const uint8_t **mask, const float **pre_blur);
/* device time of the last precompute (all kernels), microseconds */
int dpose_pg_precompute_us(const dpose_pg *pg, float *us);

/* ---- costmap ------------------------------------------------------------------------
 * device copy of a costmap_2d::Costmap2D char map (getCharMap(), getSizeInCellsX/Y()).
 * The copy is pitch-padded to 16 B for TMA.  `pitch` is the host row stride in bytes. */
int dpose_map_create(const uint8_t *data, uint32_t size_x, uint32_t size_y, size_t pitch,
                     int device, dpose_map **out);
/* same, from a buffer already in device memory on `device` (device-to-device copy) */
int dpose_map_create_from_device(const uint8_t *d_data, uint32_t size_x, uint32_t size_y,
                                 size_t pitch, int device, dpose_map **out);
/* re-uploads a dirty window [x0,x0+w) x [y0,y0+h); data points at the window's first
 * byte in a host buffer with row stride `pitch`.  The caller serialises it against
 * concurrent use of the map (the reference holds Costmap2D::getMutex(),
 * dpose_costmap.cpp:138). */
int dpose_map_update(dpose_map *map, const uint8_t *data, size_t pitch, uint32_t x0,
                     uint32_t y0, uint32_t w, uint32_t h);
/* The batched kernel reads the map through a derived 1-bit-per-cell lethal plane that the
 * library rebuilds (one streaming pass on the call's stream) after create / update.  Mark the map
 * volatile if its device bytes can change without dpose_map_update (see dpose_map_device_ptr):
 * the plane is then rebuilt inside every batched call. */
int dpose_map_set_volatile(dpose_map *map, int is_volatile);
/* the device copy itself (row stride *pitch), e.g. to let a GPU costmap layer write into it */
int dpose_map_device_ptr(const dpose_map *map, uint8_t **d_data, size_t *pitch);
int dpose_map_size(const dpose_map *map, uint32_t *size_x, uint32_t *size_y);
void dpose_map_destroy(dpose_map *map);

/* ---- lethal cells --------------------------------------------------------------------
 * replaces raytrace (dpose_costmap.cpp:57-104); host arithmetic; writes at most `cap`
 * cells and reports max(|dx|,|dy|) in *n. */
int dpose_raytrace(const int32_t begin[2], const int32_t end[2], int32_t *out_xy, size_t cap,
                   size_t *n);
/* replaces lethal_cells_within (dpose_costmap.cpp:117-184): clamps the ring into the map,
 * builds the per-row intervals (to_rays, :106-115) and stream-compacts the cells equal to
 * 254 on the GPU.  Order: y ascending, x ascending.  *cells_xy is allocated by the
 * library (dpose_free). */
int dpose_lethal_cells_within(const dpose_map *map, const int32_t box[10], int32_t **cells_xy,
                              size_t *n);
/* device time of the last dpose_lethal_cells_within* call on this map: both sweeps and the scans,
 * including the host round trip that sizes the result; microseconds */
int dpose_map_lethal_us(const dpose_map *map, float *us);
/* same, result stays on the device */
int dpose_lethal_cells_within_device(const dpose_map *map, const int32_t box[10],
                                     dpose_cells **out);
int dpose_cells_upload(const int32_t *cells_xy, size_t n, int device, dpose_cells **out);
int dpose_cells_size(const dpose_cells *cells, size_t *n);
int dpose_cells_download(const dpose_cells *cells, int32_t *cells_xy);
void dpose_cells_destroy(dpose_cells *cells);

/* ---- get_cost ------------------------------------------------------------------------
 * replaces pose_gradient::get_cost(se2, begin, end, J) (dpose_core.cpp:200-260).
 * J (3 doubles: d/dx, d/dy, d/dtheta) may be NULL.  Cells outside the kernel are
 * skipped (:237-238); n == 0 gives cost 0, J 0. */
int dpose_pg_get_cost_cells(const dpose_pg *pg, const double se2[3], const int32_t *cells_xy,
                            size_t n, double *cost, double *J);
int dpose_pg_get_cost_cells_device(const dpose_pg *pg, const double se2[3],
                                   const dpose_cells *cells, double *cost, double *J);

/* NEW batched entry point: for every pose i the cost over ALL lethal cells of `map`
 * (equivalently: the cells of lethal_cells_within(map, rotated get_box()) for poses whose
 * box lies inside the map) and its Jacobian.  out: n x 4 doubles (cost, Jx, Jy, Jtheta);
 * with want_jacobian == 0 the three J entries are written as 0.
 * Host-buffer version: chunks H2D / kernel / D2H on three streams. */
int dpose_pg_get_cost_batch(const dpose_pg *pg, const dpose_map *map, const double *poses,
                            size_t n, double *out, int want_jacobian);
/* device-buffer version, asynchronous on `stream` (a cudaStream_t, NULL = default) */
int dpose_pg_get_cost_batch_device(const dpose_pg *pg, const dpose_map *map,
                                   const double *d_poses, size_t n, double *d_out,
                                   int want_jacobian, void *stream);
/* The same, restricted to the band of map rows [row_lo, row_hi): only lethal cells of these rows count (the batched
 * form of the reference's search box, dpose_goal_tolerance.cpp:119-141 — there, too, a solve sees only the cells of
 * lethal_cells_within(map, box)), and only these rows of a VOLATILE map are re-read in front of the launch.  For callers
 * that shard a pose set spatially ("pre-sorted by map tile so each GPU's slice touches a compact
 * region"): a rank whose slice of the poses lives in a stripe of the map passes the stripe grown by
 * dpose_pg_reach() rows and pays for 1/N of the map pass; results are bit-identical to the unrestricted call for every
 * pose whose kernel window lies inside the band.  DPOSE_EINVAL if the band is empty or exceeds the map. */
int dpose_pg_get_cost_batch_device_rows(const dpose_pg *pg, const dpose_map *map,
                                        const double *d_poses, size_t n, double *d_out,
                                        int want_jacobian, void *stream, uint32_t row_lo,
                                        uint32_t row_hi);
/* rows (= columns) around a pose beyond which no cell can contribute to its cost: the largest distance of a corner of
 * the cost table from its centre, rounded up, + 1.  A pose at y sees lethal cells of rows [y - reach, y + reach] only. */
int dpose_pg_reach(const dpose_pg *pg, uint32_t *reach);
/* Set-up helper for callers that evaluate one pose set against many versions of a map (a sweep lattice, a planner's
 * candidate set): the order of the poses by the 32 x 8-cell tile of the lethal plane their position falls into (tiles
 * row-major; stable; positions outside the map go to the nearest border tile, NaN last).  perm[i] = index of the pose that
 * comes i-th.  In this order a 32-pose group of the batch kernel reads the same few tiles (L1 hits instead of L2
 * round trips: 1e7 poses on the 10000^2 map 0.948 -> 0.898 ms), and the contiguous slices of
 * dpose_pg_get_cost_batch_multi / of one rank per GPU are stripes of the map (see
 * dpose_pg_get_cost_batch_device_rows).  Host arithmetic, O(n + tiles), deterministic. */
int dpose_order_poses_by_tile(const double *poses /* n x 3 */, size_t n, uint32_t size_x, uint32_t size_y,
                              uint64_t *perm /* n */);
/* pose batch sharded contiguously over ndev devices (pgs[i] / maps[i] live on device i's
 * GPU; map and table replicated; no collective), one host thread per device. */
int dpose_pg_get_cost_batch_multi(dpose_pg *const *pgs, dpose_map *const *maps, int ndev,
                                  const double *poses, size_t n, double *out,
                                  int want_jacobian);
/* cost only: cost[i] for n poses, 8 bytes per pose back instead of 32 (no Jacobian is formed).  Host buffers. */
int dpose_pg_get_cost_batch_costs(const dpose_pg *pg, const dpose_map *map, const double *poses, size_t n,
                                  double *cost);
/* A pose LATTICE instead of a pose list (BASELINE's "pose sweep"): pose i = (x0 + dx ix, y0 + dy iy,
 * theta0 + dtheta it) with i = (it ny + iy) nx + ix, generated inside the kernel — nothing travels to the device.
 * out: nx ny ntheta x 4 doubles (or nx ny ntheta costs if cost_only != 0).  Host buffers. */
typedef struct {
  double x0, dx;
  uint64_t nx;
  double y0, dy;
  uint64_t ny;
  double theta0, dtheta;
  uint64_t ntheta;
} dpose_sweep;
int dpose_pg_get_cost_sweep(const dpose_pg *pg, const dpose_map *map, const dpose_sweep *sweep, double *out,
                            int want_jacobian, int cost_only);
/* device output buffer, asynchronous on `stream` */
int dpose_pg_get_cost_sweep_device(const dpose_pg *pg, const dpose_map *map, const dpose_sweep *sweep, double *d_out,
                                   int want_jacobian, int cost_only, void *stream);
/* algorithmic window bytes of a pose batch: sum over poses of the
 * number of map cells inside the axis-aligned bounding box of the valid kernel region,
 * clipped to the map.  Host arithmetic. */
int dpose_pg_window_bytes(const dpose_pg *pg, const dpose_map *map, const double *poses,
                          size_t n, uint64_t *bytes);

/* ---- dpose_goal_tolerance::problem, batched (the caller of the hot path) ---------------------
 * The NLP the gpp plugin hands to Ipopt (dpose_goal_tolerance.cpp:90-292): displacement x of the goal,
 *   f(x) = get_cost(goal + x, lethal cells inside the search box) + pose regularisers (:54-88, :96-110)
 *   g(x) = (x^2 + y^2, theta^2) <= (lin_tol^2, rot_tol^2)                            (:196-207, :238-247)
 * evaluated for n displacements per call (multi-start in lock-step) instead of one per callback. */
typedef struct {
  /* DposeGoalToleranceConfig (cfg/DposeGoalTolerance.cfg:7-12); lin_tolerance in CELLS
   * (dpose_goal_tolerance.cpp:530-535 converts metres to cells before problem::init sees it) */
  double lin_tolerance, rot_tolerance;
  double weight_goal_lin, weight_goal_rot, weight_start_lin, weight_start_rot;
} dpose_gt_config;

/* problem::problem (:90-94).  pg and map are borrowed and must outlive the problem. */
int dpose_problem_create(const dpose_pg *pg, const dpose_map *map, dpose_problem **out);
void dpose_problem_destroy(dpose_problem *problem);
/* problem::init(start, goal, config) (:112-181): search box, its lethal cells (kept on the device as a
 * cropped costmap), regularisers.  Poses in cells / radians. */
int dpose_problem_init(dpose_problem *problem, const double start[3], const double goal[3],
                       const dpose_gt_config *cfg);
/* the search box ring (:129-138) and get_bounds_info's constraint upper bounds (:203-206); either may be NULL */
int dpose_problem_info(const dpose_problem *problem, int32_t search_box[10], double g_upper[2]);
/* eval_f / eval_grad_f / eval_g / eval_jac_g (:215-281) for n displacements x (n x 3).
 * f: n, grad_f: n x 3, g: n x 2, jac_g: n x 3 (the three structural non-zeros dg0/dx, dg0/dy, dg1/dtheta);
 * any output may be NULL.  Host buffers. */
int dpose_problem_eval_batch(dpose_problem *problem, const double *x, size_t n, double *f, double *grad_f,
                             double *g, double *jac_g);
/* device pointers, asynchronous on `stream`; d_work: 7 n doubles of scratch */
int dpose_problem_eval_batch_device(const dpose_problem *problem, const double *d_x, size_t n, double *d_work,
                                    double *d_f, double *d_grad_f, double *d_g, double *d_jac_g, void *stream);

/* ---- the goal-tolerance solve, multi-start and device resident --------------------------------------------------
 * replaces the attempt loop of DposeGoalTolerance::preProcess / preProcessImpl (dpose_goal_tolerance.cpp:376-459):
 * upstream runs Ipopt from one starting offset per attempt (:378, :430-456), validates the solved pose with
 * check_footprint (:386-399) and takes the first attempt that passes.  Here all n starting offsets are solved at once
 * in ONE kernel launch (iterates stay in registers: no host copy, no launch per iteration), validated in a second
 * one and ranked in a third; one H2D (x0) and one D2H (results) per call.  The evaluated functions are the
 * reference's f / grad f (:96-110); the constraints (:196-207) are kept by projection; the step rule is a projected
 * BFGS descent (csrc/solve.cu), since Ipopt itself is not a per-thread algorithm. */
typedef struct {
  int max_iter;        /* Ipopt "max_iter" (plugin default 20, dpose_goal_tolerance.cpp:507): evaluations after the first */
  double tol;          /* Ipopt "tol" (plugin default 5, :504) */
  double dual_inf_tol; /* Ipopt's default 1: converged when the KKT residual ||x - P(x - grad f)||_inf <= min(tol, dual_inf_tol) */
  double armijo_c1;    /* sufficient-decrease constant of the line search, default 1e-4 */
  uint8_t lethal;      /* cost value check_footprint tests, costmap_2d::LETHAL_OBSTACLE (:395-396) */
} dpose_solver_options;
void dpose_solver_default_options(dpose_solver_options *opt);

/* per-start status byte */
#define DPOSE_SOLVE_STATUS_MASK 0x3    /* 1 converged, 2 evaluation budget exhausted, 3 stalled (no descent step left) */
#define DPOSE_SOLVE_FOOTPRINT_FREE 0x4 /* check_footprint(map, T(goal + x) footprint, lethal) passed (:388-399) */
#define DPOSE_SOLVE_ACCEPTED 0x8       /* converged and footprint free: what preProcessImpl returns true for */

typedef struct {
  long long first; /* lowest accepted start index: upstream's choice (attempts are tried in order, :430-447); -1 if none */
  long long best;  /* accepted start of lowest cost f (ties: lowest index); -1 if none */
  unsigned long long n_converged, n_accepted;
  double best_f;
  double best_pose[3];  /* goal + x of `best`, cells / radians (problem::get_pose after finalize_solution, :283-291) */
  double first_pose[3]; /* the same for `first` */
} dpose_solve_summary;

/* uniform_pose_generator (:333-373): n starting offsets, radius ~ U[0, lin_tol), angle ~ U[-pi, pi), yaw ~ U[0, rot_tol)
 * from one std::mt19937(seed) (upstream seeds from std::random_device); first_zero != 0 makes offset 0 the zero pose
 * like upstream's attempt 0 (:431).  Host arithmetic. */
int dpose_sample_offsets(uint32_t seed, double lin_tolerance, double rot_tolerance, size_t n, int first_zero,
                         double *x0 /* n x 3 */);
/* x0: n x 3 starting displacements (host).  Outputs (host, any may be NULL): x n x 3 final displacements, f n costs,
 * status n bytes (flags above), evals n evaluation counts, summary; trace (debug / parity tests): n x (max_iter + 1) x 12
 * doubles, per evaluation [x_trial(3), f, grad_f(3), g0, g1, alpha, status after the step, 0]. */
int dpose_problem_solve_batch(dpose_problem *problem, const double *x0, size_t n, const dpose_solver_options *opt,
                              double *x, double *f, uint8_t *status, int32_t *evals, dpose_solve_summary *summary,
                              double *trace);

/* device time of the last dpose_problem_solve_batch on this problem (the solve, validation and ranking kernels,
 * without the two copies), microseconds */
int dpose_problem_solve_us(const dpose_problem *problem, float *us);

/* ---- world <-> cell poses ------------------------------------------------------------------------
 * replaces to_cell / to_msg (dpose_goal_tolerance.cpp:294-330) for n poses; host arithmetic.  The costmap
 * geometry is costmap_2d::Costmap2D's (getOriginX/Y, getResolution, getSizeInCellsX/Y); yaw <-> quaternion
 * stays with the caller (tf2).
 * to_cell: Costmap2D::worldToMap — cell = (int)((w - origin) / resolution), valid iff w >= origin and
 * cell < size; out pose = (cell x, cell y, yaw).  ok[i] = 0 marks "pose outside of the map" (:303-304). */
int dpose_world_to_cell(const double *world /* n x (x, y, yaw) */, size_t n, double origin_x, double origin_y,
                        double resolution, uint32_t size_x, uint32_t size_y, double *cells /* n x 3 */,
                        uint8_t *ok /* n, may be NULL */);
/* to_msg: cell = round(pose) (std::round), Costmap2D::mapToWorld — w = origin + (cell + 0.5) * resolution.
 * ok[i] = 0 marks "negative pose" (:314-315). */
int dpose_cell_to_world(const double *cells /* n x (x, y, theta) */, size_t n, double origin_x, double origin_y,
                        double resolution, double *world /* n x 3 */, uint8_t *ok /* n, may be NULL */);

/* ---- check_footprint (validation of solved poses) -----------------------------------------------
 * replaces check_footprint(map, footprint, cost) (dpose_costmap.cpp:460-582, with is_inside :186-199 and the
 * x_bresenham iterators :201-317): *ok = 1 iff every vertex is inside the map and no cell covered by the polygon
 * (outline rays + scan-line interior) equals `cost`.  footprint_xy: n vertices in cells, closed or open; any n (up to 64
 * vertices the per-pose state stays in registers / local memory, beyond that it lives in a device workspace). */
int dpose_check_footprint(const dpose_map *map, const int32_t *footprint_xy, int n, uint8_t cost, int *ok);
/* the plugin's use (dpose_goal_tolerance.cpp:386-399) for n_poses candidates at once: the footprint is
 * transformed by T(x, y) R(theta) and rounded to cells per pose, then checked.  ok: n_poses bytes (1 = free). */
int dpose_check_footprint_batch(const dpose_map *map, const int32_t *footprint_xy, int n, const double *poses,
                                size_t n_poses, uint8_t cost, uint8_t *ok);
/* device pointers, asynchronous on `stream` */
int dpose_check_footprint_batch_device(const dpose_map *map, const int32_t *footprint_xy, int n,
                                       const double *d_poses, size_t n_poses, uint8_t cost, uint8_t *d_ok,
                                       void *stream);

#ifdef __cplusplus
}
#endif
#endif /* DPOSE_B200_H */
