"""ctypes binding of oracle/_build/libdpose_oracle.so (the C restatement).

TEST INFRASTRUCTURE ONLY — see oracle/dpose_oracle.h.  `build()` runs oracle/Makefile.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libdpose_oracle.so")


class _CostData(C.Structure):
    _fields_ = [
        ("rows", C.c_int), ("cols", C.c_int),
        ("center", C.c_int * 2), ("box", C.c_int * 10),
        ("cost", C.POINTER(C.c_float)), ("pre_blur", C.POINTER(C.c_float)),
        ("edt_sq", C.POINTER(C.c_int32)), ("outline", C.POINTER(C.c_uint8)),
        ("mask", C.POINTER(C.c_uint8)), ("min_outer", C.c_float),
    ]


class _XB(C.Structure):
    _fields_ = [("x_curr", C.c_int), ("x_sign", C.c_int), ("den", C.c_int), ("add", C.c_int), ("num", C.c_int),
                ("dx", C.c_double), ("kind", C.c_int)]


class _GtConfig(C.Structure):
    _fields_ = [(k, C.c_double) for k in ("lin_tolerance", "rot_tolerance", "weight_goal_lin", "weight_goal_rot",
                                          "weight_start_lin", "weight_start_rot")]


class _Reg(C.Structure):
    _fields_ = [("goal", C.c_double * 3), ("norm", C.c_double * 3)]


class _Problem(C.Structure):
    _fields_ = [("pose", C.c_double * 3), ("lin_tol_sq", C.c_double), ("rot_tol_sq", C.c_double),
                ("search_box", C.c_int32 * 10), ("n_regs", C.c_int), ("regs", _Reg * 2),
                ("cells", C.POINTER(C.c_int32)), ("n_cells", C.c_size_t)]


class SolverCfg(C.Structure):
    """oracle_solver_cfg"""
    _fields_ = [("max_iter", C.c_int), ("kkt_tol", C.c_double), ("c1", C.c_double), ("rho2", C.c_double),
                ("lin_tol", C.c_double), ("rot_tol", C.c_double)]


class SolverState(C.Structure):
    """oracle_solver_state"""
    _fields_ = [("x", C.c_double * 3), ("f", C.c_double), ("g", C.c_double * 3), ("H", C.c_double * 6),
                ("d", C.c_double * 3), ("xt", C.c_double * 3), ("alpha", C.c_double), ("gd", C.c_double),
                ("n_acc", C.c_int), ("evals", C.c_int), ("status", C.c_int)]


TRACE_STRIDE = 12  # doubles per evaluation: x_trial(3), f, grad_f(3), g0, g1, alpha, status after the step, 0


def solver_cfg(box, lin_tolerance, rot_tolerance, max_iter=20, tol=5.0, dual_inf_tol=1.0, armijo_c1=1e-4):
    """the step rule's constants for a footprint whose cost_data box is `box` ((5, 2) ring): rho^2 = the largest
    squared distance of a box corner from the centre cell (at least 1)"""
    b = np.asarray(box, np.float64).reshape(5, 2)
    return SolverCfg(int(max_iter), float(min(tol, dual_inf_tol)), float(armijo_c1),
                     float(max(1.0, (b ** 2).sum(axis=1).max())), abs(float(lin_tolerance)), abs(float(rot_tolerance)))


def solver_replay(cfg, x0, trace_f_grad):
    """feeds a recorded sequence of (f, grad f) into the restated step rule starting from x0; returns the trial
    points it asks for (one per evaluation), the status after each step and the final state"""
    s = SolverState()
    x0 = np.ascontiguousarray(np.asarray(x0, np.float64).reshape(3))
    lib().oracle_solver_start(C.byref(s), C.byref(cfg), x0.ctypes.data)
    trials, statuses = [], []
    for ft, gt in trace_f_grad:
        if s.status != 0:
            break
        trials.append(np.array(s.xt[:]))
        g = np.ascontiguousarray(np.asarray(gt, np.float64).reshape(3))
        lib().oracle_solver_advance(C.byref(s), C.byref(cfg), float(ft), g.ctypes.data)
        statuses.append(s.status)
    return np.array(trials).reshape(-1, 3), np.array(statuses, np.int64), s


def sample_offsets(seed, lin_tolerance, rot_tolerance, n, first_zero=True):
    """uniform_pose_generator (dpose_goal_tolerance.cpp:333-373) with std::mt19937(seed), restated"""
    out = np.zeros((int(n), 3))
    lib().oracle_sample_offsets(int(seed), float(lin_tolerance), float(rot_tolerance), int(n), int(bool(first_zero)),
                                out.ctypes.data)
    return out


def build(force=False):
    src_newer = (not os.path.exists(_SO)) or any(
        os.path.getmtime(os.path.join(_HERE, f)) > os.path.getmtime(_SO)
        for f in ("dpose_oracle.c", "dpose_oracle.h"))
    if force or src_newer:
        subprocess.run(["make", "-C", _HERE, "-s"], check=True)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        L.oracle_make_footprint.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_void_p]
        L.oracle_cost_data_create.argtypes = [C.c_void_p, C.c_int, C.c_uint, C.POINTER(_CostData)]
        L.oracle_cost_data_free.argtypes = [C.POINTER(_CostData)]
        L.oracle_get_cost.restype = C.c_double
        L.oracle_get_cost.argtypes = [C.POINTER(_CostData), C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
        L.oracle_get_cost_truth.restype = C.c_longdouble
        L.oracle_get_cost_truth.argtypes = [C.POINTER(_CostData), C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
        L.oracle_raytrace.restype = C.c_size_t
        L.oracle_raytrace.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.oracle_lethal_cells_within.restype = C.c_size_t
        L.oracle_lethal_cells_within.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.POINTER(C.c_void_p)]
        L.oracle_get_cost_batch.argtypes = [C.POINTER(_CostData), C.c_void_p, C.c_uint32, C.c_uint32,
                                            C.c_void_p, C.c_size_t, C.c_void_p, C.c_int, C.c_int, C.c_int]
        L.oracle_get_cost_batch_truth.argtypes = [C.POINTER(_CostData), C.c_void_p, C.c_uint32, C.c_uint32,
                                                  C.c_void_p, C.c_size_t, C.c_void_p, C.c_int]
        L.oracle_max_threads.restype = C.c_int
        L.oracle_is_inside.argtypes = [C.c_uint32, C.c_uint32, C.c_void_p, C.c_int]
        L.oracle_xb_init.argtypes = [C.POINTER(_XB), C.c_void_p, C.c_void_p]
        L.oracle_xb_advance.argtypes = [C.POINTER(_XB)]
        L.oracle_check_footprint.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_int, C.c_uint8]
        L.oracle_check_footprint_batch.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_int, C.c_void_p,
                                                   C.c_size_t, C.c_uint8, C.c_void_p, C.c_int]
        L.oracle_normalize_angle.restype = C.c_double
        L.oracle_normalize_angle.argtypes = [C.c_double]
        L.oracle_problem_init.argtypes = [C.POINTER(_CostData), C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p,
                                          C.c_void_p, C.POINTER(_GtConfig), C.POINTER(_Problem)]
        L.oracle_problem_free.argtypes = [C.POINTER(_Problem)]
        L.oracle_problem_eval_batch.argtypes = [C.POINTER(_CostData), C.POINTER(_Problem), C.c_void_p, C.c_size_t,
                                                C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.oracle_solver_start.argtypes = [C.POINTER(SolverState), C.POINTER(SolverCfg), C.c_void_p]
        L.oracle_solver_advance.argtypes = [C.POINTER(SolverState), C.POINTER(SolverCfg), C.c_double, C.c_void_p]
        L.oracle_solve_batch.argtypes = [C.POINTER(_CostData), C.POINTER(_Problem), C.POINTER(SolverCfg), C.c_void_p,
                                         C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.oracle_sample_offsets.argtypes = [C.c_uint32, C.c_double, C.c_double, C.c_size_t, C.c_int, C.c_void_p]
        _lib = L
    return _lib


_libc = C.CDLL(None)
_libc.free.argtypes = [C.c_void_p]


def _i32(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.int32))


def make_footprint(xy_m, res):
    """dpose_costmap.cpp:40-55; xy_m: (n,2) metres -> (n,2) int32 cells."""
    xy = np.ascontiguousarray(np.asarray(xy_m, dtype=np.float64).reshape(-1, 2))
    out = np.empty(xy.shape, dtype=np.int32)
    if lib().oracle_make_footprint(xy.ctypes.data, xy.shape[0], float(res), out.ctypes.data) != 0:
        raise RuntimeError("resolution must be positive")
    return out


class CostData:
    """cost_data (dpose_core.cpp:122-139). footprint: (n,2) int cells (x,y)."""

    def __init__(self, footprint, padding):
        fp = _i32(footprint).reshape(-1, 2)
        self._cd = _CostData()
        if lib().oracle_cost_data_create(fp.ctypes.data, fp.shape[0], int(padding), C.byref(self._cd)) != 0:
            raise RuntimeError("bad footprint")
        cd = self._cd
        shape = (cd.rows, cd.cols)
        n = cd.rows * cd.cols
        self.rows, self.cols = shape
        self.cost = np.ctypeslib.as_array(cd.cost, (n,)).reshape(shape).copy()
        self.pre_blur = np.ctypeslib.as_array(cd.pre_blur, (n,)).reshape(shape).copy()
        self.edt_sq = np.ctypeslib.as_array(cd.edt_sq, (n,)).reshape(shape).copy()
        self.outline = np.ctypeslib.as_array(cd.outline, (n,)).reshape(shape).copy()
        self.mask = np.ctypeslib.as_array(cd.mask, (n,)).reshape(shape).copy()
        self.center = np.array(list(cd.center), dtype=np.int32)
        self.box = np.array(list(cd.box), dtype=np.int32).reshape(5, 2)  # rows = ring points (x,y)
        self.min_outer = float(cd.min_outer)

    def __del__(self):
        try:
            lib().oracle_cost_data_free(C.byref(self._cd))
        except Exception:
            pass

    def get_cost(self, se2, cells, want_jacobian=True):
        """pose_gradient::get_cost (dpose_core.cpp:200-260), FD Jacobian."""
        se2 = np.ascontiguousarray(np.asarray(se2, dtype=np.float64))
        cells = _i32(cells).reshape(-1, 2)
        J = np.zeros(3, dtype=np.float64)
        cost = lib().oracle_get_cost(C.byref(self._cd), se2.ctypes.data, cells.ctypes.data,
                                     cells.shape[0], J.ctypes.data if want_jacobian else None)
        return cost, J

    def get_cost_truth(self, se2, cells):
        se2 = np.ascontiguousarray(np.asarray(se2, dtype=np.float64))
        cells = _i32(cells).reshape(-1, 2)
        J = np.zeros(3, dtype=np.longdouble)
        cost = lib().oracle_get_cost_truth(C.byref(self._cd), se2.ctypes.data, cells.ctypes.data,
                                           cells.shape[0], J.ctypes.data)
        return float(cost), J.astype(np.float64)

    def get_cost_batch(self, costmap, poses, want_jacobian=True, shift=False, nthreads=1):
        m = np.ascontiguousarray(np.asarray(costmap, dtype=np.uint8))
        p = np.ascontiguousarray(np.asarray(poses, dtype=np.float64).reshape(-1, 3))
        out = np.zeros((p.shape[0], 4), dtype=np.float64)
        lib().oracle_get_cost_batch(C.byref(self._cd), m.ctypes.data, m.shape[1], m.shape[0],
                                    p.ctypes.data, p.shape[0], out.ctypes.data,
                                    int(want_jacobian), int(shift), int(nthreads))
        return out

    def get_cost_batch_truth(self, costmap, poses, nthreads=1):
        m = np.ascontiguousarray(np.asarray(costmap, dtype=np.uint8))
        p = np.ascontiguousarray(np.asarray(poses, dtype=np.float64).reshape(-1, 3))
        out = np.zeros((p.shape[0], 4), dtype=np.float64)
        lib().oracle_get_cost_batch_truth(C.byref(self._cd), m.ctypes.data, m.shape[1], m.shape[0],
                                          p.ctypes.data, p.shape[0], out.ctypes.data, int(nthreads))
        return out


def raytrace(begin, end):
    """dpose_costmap.cpp:57-104."""
    b, e = _i32(begin), _i32(end)
    n = int(max(abs(int(e[0]) - int(b[0])), abs(int(e[1]) - int(b[1]))))
    out = np.empty((max(n, 1), 2), dtype=np.int32)
    k = lib().oracle_raytrace(b.ctypes.data, e.ctypes.data, out.ctypes.data)
    return out[:k].copy()


def lethal_cells_within(costmap, box):
    """dpose_costmap.cpp:117-184; box: (5,2) ring (x,y). Returns (n,2) int32, (y,x)-sorted."""
    m = np.ascontiguousarray(np.asarray(costmap, dtype=np.uint8))
    b = _i32(box).reshape(10)
    ptr = C.c_void_p()
    sy, sx = (m.shape if m.ndim == 2 else (0, 0))
    n = lib().oracle_lethal_cells_within(m.ctypes.data, sx, sy, b.ctypes.data, C.byref(ptr))
    if n == 0:
        return np.empty((0, 2), dtype=np.int32)
    arr = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_int32)), (n * 2,)).reshape(n, 2).copy()
    _libc.free(ptr)
    return arr


def max_threads():
    """host threads the CPU baseline may use: the cores this process is allowed on.  Deliberately NOT
    omp_get_max_threads(): torchrun exports OMP_NUM_THREADS=1 to its workers, which would silently turn
    the all-cores baseline into a single-thread one (num_threads() in the C code overrides the env)."""
    import os
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def normalize_angle(a):
    return lib().oracle_normalize_angle(float(a))


class Problem:
    """dpose_goal_tolerance::problem restated (dpose_goal_tolerance.cpp:90-292): init() + the TNLP evaluations."""

    def __init__(self, cost_data, costmap):
        self.cd = cost_data
        self.map = np.ascontiguousarray(np.asarray(costmap, dtype=np.uint8))
        self._p = _Problem()
        self._live = False

    def init(self, start, goal, **cfg):
        if self._live:
            lib().oracle_problem_free(C.byref(self._p))
        c = _GtConfig(**{k: float(cfg.get(k, 1.0)) for k, _ in _GtConfig._fields_})
        s = np.ascontiguousarray(np.asarray(start, np.float64).reshape(3))
        g = np.ascontiguousarray(np.asarray(goal, np.float64).reshape(3))
        lib().oracle_problem_init(C.byref(self.cd._cd), self.map.ctypes.data, self.map.shape[1], self.map.shape[0],
                                  s.ctypes.data, g.ctypes.data, C.byref(c), C.byref(self._p))
        self._live = True

    @property
    def search_box(self):
        return np.array(self._p.search_box[:], np.int32).reshape(5, 2)

    @property
    def n_cells(self):
        return int(self._p.n_cells)

    @property
    def bounds(self):
        """get_bounds_info (:196-207): g upper bounds (lin_tol^2, rot_tol^2)"""
        return self._p.lin_tol_sq, self._p.rot_tol_sq

    def eval_batch(self, x, nthreads=1):
        x = np.ascontiguousarray(np.asarray(x, np.float64).reshape(-1, 3))
        n = x.shape[0]
        f, grad, g, jac = np.empty(n), np.empty((n, 3)), np.empty((n, 2)), np.empty((n, 3))
        lib().oracle_problem_eval_batch(C.byref(self.cd._cd), C.byref(self._p), x.ctypes.data, n, f.ctypes.data,
                                        grad.ctypes.data, g.ctypes.data, jac.ctypes.data, int(nthreads))
        return f, grad, g, jac

    def solve_batch(self, cfg, x0, nthreads=1, trace=False):
        """the restated step rule on the oracle's own evaluation, n starts: (x, f, status, evals[, trace])"""
        x0 = np.ascontiguousarray(np.asarray(x0, np.float64).reshape(-1, 3))
        n = x0.shape[0]
        x, f = np.empty((n, 3)), np.empty(n)
        status, evals = np.zeros(n, np.uint8), np.zeros(n, np.int32)
        tr = np.zeros((n, cfg.max_iter + 1, TRACE_STRIDE)) if trace else None
        lib().oracle_solve_batch(C.byref(self.cd._cd), C.byref(self._p), C.byref(cfg), x0.ctypes.data, n, x.ctypes.data,
                                 f.ctypes.data, status.ctypes.data, evals.ctypes.data,
                                 tr.ctypes.data if trace else None, int(nthreads))
        return (x, f, status, evals, tr) if trace else (x, f, status, evals)

    def __del__(self):
        try:
            if self._live:
                lib().oracle_problem_free(C.byref(self._p))
        except Exception:
            pass


def is_inside(size_x, size_y, footprint):
    fp = _i32(footprint).reshape(-1, 2)
    return bool(lib().oracle_is_inside(int(size_x), int(size_y), fp.ctypes.data, fp.shape[0]))


class XBresenham:
    """x_bresenham (dpose_costmap.cpp:303-317)"""

    def __init__(self, c0, c1):
        self._b = _XB()
        a, b = _i32(c0), _i32(c1)
        lib().oracle_xb_init(C.byref(self._b), a.ctypes.data, b.ctypes.data)

    def get_x(self):
        return int(self._b.x_curr)

    def get_dx(self):
        return float(self._b.dx)

    def advance_x(self):
        lib().oracle_xb_advance(C.byref(self._b))


def check_footprint(costmap, footprint, cost=254):
    m = np.ascontiguousarray(np.asarray(costmap, dtype=np.uint8))
    fp = _i32(footprint).reshape(-1, 2)
    return bool(lib().oracle_check_footprint(m.ctypes.data, m.shape[1], m.shape[0], fp.ctypes.data, fp.shape[0], int(cost)))


def check_footprint_batch(costmap, footprint, poses, cost=254, nthreads=1):
    m = np.ascontiguousarray(np.asarray(costmap, dtype=np.uint8))
    fp = _i32(footprint).reshape(-1, 2)
    p = np.ascontiguousarray(np.asarray(poses, np.float64).reshape(-1, 3))
    ok = np.zeros(p.shape[0], np.uint8)
    lib().oracle_check_footprint_batch(m.ctypes.data, m.shape[1], m.shape[0], fp.ctypes.data, fp.shape[0], p.ctypes.data,
                                       p.shape[0], int(cost), ok.ctypes.data, int(nthreads))
    return ok.astype(bool)
