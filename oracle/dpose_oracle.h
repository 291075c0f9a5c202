/*
 * dpose_oracle.h — CPU restatement of dpose_core's footprint-cost path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is linked into, imported by or
 * executed from the product (dpose_b200/, include/).  Only tests/, the smoke check in
 * __graft_entry__.py and bench.py's cpu_baseline / --impl reference legs may use it,
 * and there only as the checker / the timed CPU baseline.
 *
 * Every function cites the reference file:line it restates (paths relative to the
 * upstream dorezyuk/dpose tree).  Pins, strongest first:
 *   1. oracle/_ref — the reference's own dpose_core.cpp / dpose_costmap.cpp compiled from
 *      /root/reference against the stand-in headers of oracle/refshim (the real Eigen,
 *      OpenCV, costmap_2d and boost are absent from this image); every function below is
 *      bit-identical to it (tests/test_reference_build.py);
 *   2. the image arithmetic lives in OpenCV (un-vendored; ROS noetic's libopencv-dev) and is
 *      pinned against the same OpenCV functions through the cv2 4.13.0 Python wheel
 *      (oracle/cv_pipeline.py, tests/golden/make_golden.py);
 *   3. the known-answer vectors of SURVEY.md §8(c) and the reference's own unit tests restated.
 */
#ifndef DPOSE_ORACLE_H
#define DPOSE_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* costmap_2d/cost_values.h (external): LETHAL_OBSTACLE = 254 */
#define ORACLE_LETHAL 254

/* cost_data (dpose_core.hpp:79-102): cost image + center + box, plus the
 * intermediate images kept for bit-exact checks. */
typedef struct oracle_cost_data {
  int rows, cols;       /* cv::Mat size of `cost` */
  int center[2];        /* cost_data::center (x, y) */
  int box[10];          /* cost_data::box, Eigen 2x5 column-major: x0,y0,x1,y1,... */
  float *cost;          /* rows*cols, row-major, final blurred table */
  float *pre_blur;      /* rows*cols, table before GaussianBlur */
  int32_t *edt_sq;      /* rows*cols, exact squared distance to the outline */
  uint8_t *outline;     /* rows*cols, 0 on the polylines outline, 255 else */
  uint8_t *mask;        /* rows*cols, 0 inside (fillPoly), 255 outside */
  float min_outer;      /* minMaxIdx(outer) minimum */
} oracle_cost_data;

/* make_footprint (dpose_costmap.cpp:40-55): metres -> cells, round half away.
 * returns 0, or -1 if res <= 0 (the reference throws std::runtime_error). */
int oracle_make_footprint(const double *xy_m, int n, double res, int32_t *xy_cells);

/* cost_data::cost_data + _get_cost (dpose_core.cpp:55-139).
 * fp_xy: x0,y0,x1,y1... (n vertices).  returns 0 on success, -1 on bad input. */
int oracle_cost_data_create(const int32_t *fp_xy, int n, unsigned padding,
                            oracle_cost_data *out);
void oracle_cost_data_free(oracle_cost_data *cd);

/* OpenCV primitives behind _get_cost, one by one (cv::polylines closed / cv::fillPoly single contour /
 * cv::distanceTransform DIST_L2 PRECISE as squared integers); used by oracle/refshim/opencv2. */
void oracle_cv_polylines_closed(uint8_t *img, int cols, int rows, const int32_t *xy, int n, uint8_t color);
void oracle_cv_fillpoly(uint8_t *img, int cols, int rows, const int32_t *xy, int n, uint8_t color);
void oracle_cv_edt_sq(const uint8_t *img, int cols, int rows, int32_t *d2);

/* pose_gradient::get_cost (dpose_core.cpp:200-260) incl. interpolator
 * (dpose_core.cpp:158-198): finite-difference Jacobian exactly as the reference.
 * J may be NULL.  cells_xy: x0,y0,x1,y1,... */
double oracle_get_cost(const oracle_cost_data *cd, const double se2[3],
                       const int32_t *cells_xy, size_t n, double J[3]);

/* Same cost, Jacobian by the analytic chain rule on the bilinear patch, evaluated in
 * long double (adjudication "truth", SURVEY.md §7 hard part 2).  Not in the reference. */
long double oracle_get_cost_truth(const oracle_cost_data *cd, const double se2[3],
                                  const int32_t *cells_xy, size_t n, long double J[3]);

/* raytrace (dpose_costmap.cpp:57-104): writes max(|dx|,|dy|) cells (end exclusive)
 * into out_xy (caller sized) and returns that count. */
size_t oracle_raytrace(const int32_t begin[2], const int32_t end[2], int32_t *out_xy);

/* to_rays (dpose_costmap.cpp:106-115) on a closed 5-point ring: per-row [min,max] x.
 * rows are reported for y in [*y_lo, *y_hi]; xmin/xmax arrays are caller sized
 * (y_hi-y_lo+1 <= span of the ring) and rows never touched keep xmin>xmax.
 * returns number of rows spanned. */
int oracle_to_rays(const int32_t box[10], int *y_lo, int *y_hi, int32_t **xmin,
                   int32_t **xmax);

/* lethal_cells_within (dpose_costmap.cpp:117-184).  Output order is canonical
 * (y ascending, x ascending); the reference's order is unordered_map iteration order.
 * *cells_xy is malloc'ed (free with free()); returns count. */
size_t oracle_lethal_cells_within(const uint8_t *map, uint32_t size_x, uint32_t size_y,
                                  const int32_t box[10], int32_t **cells_xy);

/* The natural per-pose use of dpose_core for arbitrary poses:
 * box = round(R(theta) * get_box() + t); cells = lethal_cells_within(map, box);
 * get_cost(pose, cells, &J).  out: n x 4 (cost, Jx, Jy, Jtheta).
 * shift != 0 evaluates on translated-equivalent inputs (pose and cells shifted by
 * floor(tx), floor(ty)) — the parity protocol of SURVEY.md §7 hard part 2.
 * nthreads <= 1: single thread (the reference is single-threaded); else OpenMP. */
void oracle_get_cost_batch(const oracle_cost_data *cd, const uint8_t *map, uint32_t size_x,
                           uint32_t size_y, const double *poses, size_t n, double *out,
                           int want_jacobian, int shift, int nthreads);

/* long-double analytic evaluation of the same batch (adjudication). out: n x 4 doubles. */
void oracle_get_cost_batch_truth(const oracle_cost_data *cd, const uint8_t *map,
                                 uint32_t size_x, uint32_t size_y, const double *poses,
                                 size_t n, double *out, int nthreads);

/* ---- dpose_goal_tolerance::problem: the NLP the Ipopt loop evaluates ------------- */
/* DposeGoalToleranceConfig (cfg/DposeGoalTolerance.cfg:7-12); lin_tolerance already in CELLS
 * (dpose_goal_tolerance.cpp:530-535 converts it in the reconfigure callback). */
typedef struct {
  double lin_tolerance, rot_tolerance;
  double weight_goal_lin, weight_goal_rot, weight_start_lin, weight_start_rot;
} oracle_gt_config;

/* pose_regularization after init() (dpose_goal_tolerance.cpp:54-74) */
typedef struct {
  double goal[3], norm[3];
} oracle_reg;

/* problem after init() (dpose_goal_tolerance.cpp:112-181) */
typedef struct {
  double pose[3];            /* pose_ = goal */
  double lin_tol_sq, rot_tol_sq;
  int32_t search_box[10];    /* :129-138 */
  int n_regs;
  oracle_reg regs[2];        /* :144-180: goal regulariser first, then start */
  int32_t *cells;            /* lethal_cells_ (:141), malloc'ed */
  size_t n_cells;
} oracle_problem;

/* angles::normalize_angle / shortest_angular_distance (ROS package `angles`, un-vendored; restated
 * from its published header: fmod(angle + pi, 2 pi), result <= 0 ? + pi : - pi) */
double oracle_normalize_angle(double a);

int oracle_problem_init(const oracle_cost_data *cd, const uint8_t *map, uint32_t size_x, uint32_t size_y,
                        const double start[3], const double goal[3], const oracle_gt_config *cfg,
                        oracle_problem *out);
void oracle_problem_free(oracle_problem *p);
/* on_new_x + eval_f / eval_grad_f / eval_g / eval_jac_g (dpose_goal_tolerance.cpp:96-110, 215-281)
 * for one displacement x; any output may be NULL. */
void oracle_problem_eval(const oracle_cost_data *cd, const oracle_problem *p, const double x[3], double *f,
                         double grad_f[3], double g[2], double jac_g[3]);
/* the same for n displacements (OpenMP over x when nthreads > 1) */
void oracle_problem_eval_batch(const oracle_cost_data *cd, const oracle_problem *p, const double *x, size_t n,
                               double *f, double *grad_f, double *g, double *jac_g, int nthreads);

/* ---- check_footprint: dpose_costmap.cpp:186-582 ------------------------------------ */
/* is_inside (:186-199): 1 if every vertex lies inside the map */
int oracle_is_inside(uint32_t size_x, uint32_t size_y, const int32_t *fp_xy, int n);

/* x_bresenham (:201-317): x of a line per scan-line, starting at the vertex with the lower y */
typedef struct {
  int x_curr, x_sign, den, add, num;
  double dx;
  int kind; /* 0 x_minor_ascending, 1 x_minor_descending, 2 x_major_ascending, 3 x_major_descending */
} oracle_xb;
void oracle_xb_init(oracle_xb *b, const int32_t c0[2], const int32_t c1[2]);
void oracle_xb_advance(oracle_xb *b);

/* check_footprint (:460-582): 1 if no cell covered by the polygon (outline + scan-line interior) holds
 * `cost`; 0 otherwise or if a vertex is outside the map.  fp_xy: vertices in cells (closed or open). */
int oracle_check_footprint(const uint8_t *map, uint32_t size_x, uint32_t size_y, const int32_t *fp_xy, int n,
                           uint8_t cost);
/* the plugin's use (dpose_goal_tolerance.cpp:386-395): footprint transformed by T(x, y) R(theta), rounded
 * (Eigen round = std::round), then check_footprint.  ok: n_poses bytes. */
void oracle_check_footprint_batch(const uint8_t *map, uint32_t size_x, uint32_t size_y, const int32_t *fp_xy, int n,
                                  const double *poses, size_t n_poses, uint8_t cost, uint8_t *ok, int nthreads);

/* ---- the multi-start solve ----------------------------------------------------------------
 * The reference solves its NLP with Ipopt (dpose_goal_tolerance.cpp:378), which is absent from this image and is not
 * a per-thread algorithm.  The product's solver (dpose_b200/csrc/solve.cu) is a projected BFGS descent of its own; what
 * IS the reference's are the functions it evaluates (oracle_problem_eval / the compiled plugin's callbacks), the
 * constraint set (:196-207), the validation of the solved pose (:386-399) and the sampler of the starting offsets
 * (:333-373).  Below: an independent C restatement of that step rule, written from the rule's description in
 * solve.cu's header (not from its code), so that the tests can check (a) every iterate's f / grad f against the
 * reference's callbacks and (b) every step against this state machine, bit for bit; and a driver that runs the whole
 * solve on the CPU with ANY evaluation callback — the compiled plugin's for bench.py's reference arm. */
typedef struct {
  int max_iter;               /* evaluations after the first */
  double kkt_tol, c1, rho2;   /* stop tolerance, Armijo constant, squared footprint radius (theta metric) */
  double lin_tol, rot_tol;    /* the constraint set: x^2 + y^2 <= lin_tol^2, |theta| <= rot_tol */
} oracle_solver_cfg;
typedef struct {
  double x[3], f, g[3];  /* accepted iterate */
  double H[6];           /* BFGS inverse Hessian approximation, symmetric: 00 01 02 11 12 22 */
  double d[3], xt[3];    /* search direction at x, trial point to evaluate next */
  double alpha, gd;      /* step along d, grad f . (xt - x) */
  int n_acc, evals, status; /* status: 0 running, 1 converged, 2 evaluation budget exhausted, 3 stalled */
} oracle_solver_state;
void oracle_solver_start(oracle_solver_state *s, const oracle_solver_cfg *cfg, const double x0[3]);
/* feeds f / grad f evaluated at s->xt; on return s->xt is the next point to evaluate unless s->status != 0 */
void oracle_solver_advance(oracle_solver_state *s, const oracle_solver_cfg *cfg, double ft, const double gt[3]);
/* eval_f + eval_grad_f (+ eval_g, eval_jac_g) at a displacement: the signature of oracle/refshim's ref_problem_eval */
typedef void (*oracle_eval_fn)(void *ctx, const double x[3], double *f, double grad_f[3], double g[2], double jac_g[3]);
/* one start, to the end; trace (may be NULL): (max_iter + 1) x 12 doubles, per evaluation
 * [x_trial(3), f, grad_f(3), g0, g1, alpha, status after the step, 0].  Returns the status. */
int oracle_solve(const oracle_solver_cfg *cfg, oracle_eval_fn eval, void *ctx, const double x0[3], double x[3], double *f,
                 int *evals, double *trace);
/* adapter: ctx for oracle_problem_eval as an oracle_eval_fn */
typedef struct {
  const oracle_cost_data *cd;
  const oracle_problem *p;
} oracle_problem_ctx;
void oracle_problem_eval_cb(void *ctx, const double x[3], double *f, double grad_f[3], double g[2], double jac_g[3]);
/* n starts with the oracle's own problem evaluation (OpenMP over starts) */
void oracle_solve_batch(const oracle_cost_data *cd, const oracle_problem *p, const oracle_solver_cfg *cfg,
                        const double *x0, size_t n, double *x, double *f, uint8_t *status, int32_t *evals, double *trace,
                        int nthreads);
/* uniform_pose_generator (dpose_goal_tolerance.cpp:333-373) with std::mt19937(seed) and libstdc++'s
 * std::uniform_real_distribution<double> restated (generate_canonical<double, 53>: two 32-bit draws, low word first);
 * pinned against the plugin's own uniform_pose_distribution compiled into oracle/_ref */
void oracle_sample_offsets(uint32_t seed, double lin_tol, double rot_tol, size_t n, int first_zero, double *x0);

/* number of OpenMP threads available (1 if built without OpenMP) */
int oracle_max_threads(void);

#ifdef __cplusplus
}
#endif
#endif
