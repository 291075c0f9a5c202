"""Python mirror of dpose_core's public interface over the C ABI.

reference surface                         | here
------------------------------------------|---------------------------------------------
cost_data(footprint, padding)             | cost_data            (dpose_core.hpp:79-102)
pose_gradient(footprint, parameter)       | pose_gradient        (dpose_core.hpp:133-167)
pose_gradient::get_cost(se2, b, e, &J)    | pose_gradient.get_cost(se2, cells, jacobian=True)
make_footprint(layered_costmap)           | make_footprint(points_m, resolution)
lethal_cells_within(costmap, box)         | lethal_cells_within(costmap, box)
raytrace(begin, end)                      | raytrace(begin, end)
to_rectangle(w, h) / (min, max)           | to_rectangle
NEW get_cost(poses[])                     | pose_gradient.get_cost_batch(costmap, poses)

Polygons / cell vectors are (n, 2) int32 arrays of (x, y); rectangles are (5, 2) rings.
"""
import ctypes as C

import numpy as np

from ._lib import DposeError, check, lib, EINVAL


def device_count():
    n = C.c_int(0)
    lib().dpose_device_count(C.byref(n))
    return n.value


def launch_count():
    return int(lib().dpose_launch_count())


def _i32(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.int32))


def to_rectangle(a, b):
    """dpose_core.hpp:45-69"""
    if np.ndim(a) == 0:
        lo, hi = (0, 0), (int(a), int(b))
    else:
        lo, hi = a, b
    return np.array([[lo[0], lo[1]], [hi[0], lo[1]], [hi[0], hi[1]], [lo[0], hi[1]], [lo[0], lo[1]]],
                    dtype=np.int32)


def make_footprint(points_m, resolution):
    """dpose_costmap.cpp:40-55.  Raises RuntimeError("resolution must be positive")."""
    xy = np.ascontiguousarray(np.asarray(points_m, dtype=np.float64).reshape(-1, 2))
    out = np.empty(xy.shape, dtype=np.int32)
    rc = lib().dpose_make_footprint(xy.ctypes.data, xy.shape[0], float(resolution), out.ctypes.data)
    if rc == EINVAL:
        raise RuntimeError(lib().dpose_last_error().decode())
    check(rc)
    return out


def raytrace(begin, end):
    """dpose_costmap.cpp:57-104 (end exclusive)."""
    b, e = _i32(begin), _i32(end)
    cap = int(max(abs(int(e[0]) - int(b[0])), abs(int(e[1]) - int(b[1]))))
    out = np.empty((max(cap, 1), 2), dtype=np.int32)
    n = C.c_size_t(0)
    check(lib().dpose_raytrace(b.ctypes.data, e.ctypes.data, out.ctypes.data, cap, C.byref(n)))
    return out[: n.value].copy()


class Costmap:
    """Device-resident copy of a costmap_2d::Costmap2D char map (uint8, [y, x])."""

    def __init__(self, data=None, device=0, *, device_ptr=None, size_x=None, size_y=None, pitch=None):
        self._h = C.c_void_p()
        self.device = int(device)
        if device_ptr is not None:
            check(lib().dpose_map_create_from_device(C.c_void_p(int(device_ptr)), size_x, size_y,
                                                     pitch or size_x, self.device, C.byref(self._h)))
            self.size_x, self.size_y = int(size_x), int(size_y)
        else:
            m = np.ascontiguousarray(np.asarray(data, dtype=np.uint8))
            if m.ndim != 2:
                m = m.reshape(0, 0)
            self.size_y, self.size_x = m.shape
            check(lib().dpose_map_create(m.ctypes.data, self.size_x, self.size_y, m.strides[0] if m.size else 0,
                                         self.device, C.byref(self._h)))

    def set_volatile(self, flag=True):
        """rebuild the derived lethal bit plane inside every batched call (the device bytes may
        change behind the library's back); bench.py times with this ON"""
        check(lib().dpose_map_set_volatile(self._h, int(bool(flag))))

    def device_ptr(self):
        """(device pointer, pitch) of the library's copy of the map: a producer on the device (a costmap layer kernel)
        writes here; mark the map volatile, or the lethal plane will not follow (dpose_map_device_ptr)"""
        p, pitch = C.c_void_p(), C.c_size_t(0)
        check(lib().dpose_map_device_ptr(self._h, C.byref(p), C.byref(pitch)))
        return int(p.value), int(pitch.value)

    def update(self, window, x0, y0):
        """re-upload a dirty window (rows [y0, y0+h), columns [x0, x0+w))"""
        w = np.ascontiguousarray(np.asarray(window, dtype=np.uint8))
        check(lib().dpose_map_update(self._h, w.ctypes.data, w.strides[0], int(x0), int(y0), w.shape[1], w.shape[0]))

    def lethal_us(self):
        """device time of the last lethal_cells_within on this map, microseconds"""
        us = C.c_float(0)
        check(lib().dpose_map_lethal_us(self._h, C.byref(us)))
        return us.value

    def close(self):
        if self._h:
            lib().dpose_map_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class cell_vector_device:
    """A cell_vector kept in device memory (result of lethal_cells_within(..., on_device=True))."""

    def __init__(self, handle, device):
        self._h = handle
        self.device = device

    @classmethod
    def upload(cls, cells, device=0):
        c = _i32(cells).reshape(-1, 2)
        h = C.c_void_p()
        check(lib().dpose_cells_upload(c.ctypes.data, c.shape[0], int(device), C.byref(h)))
        return cls(h, int(device))

    def __len__(self):
        n = C.c_size_t(0)
        check(lib().dpose_cells_size(self._h, C.byref(n)))
        return n.value

    def numpy(self):
        out = np.empty((len(self), 2), dtype=np.int32)
        check(lib().dpose_cells_download(self._h, out.ctypes.data))
        return out

    def __del__(self):
        try:
            if self._h:
                lib().dpose_cells_destroy(self._h)
        except Exception:
            pass


def lethal_cells_within(costmap, box, on_device=False):
    """dpose_costmap.cpp:117-184.  costmap: Costmap (or a uint8 array, uploaded on the fly).
    Returns (n, 2) int32 cells, y ascending then x ascending."""
    own = None
    if not isinstance(costmap, Costmap):
        own = costmap = Costmap(costmap)
    b = _i32(box).reshape(10)
    try:
        if on_device:
            h = C.c_void_p()
            check(lib().dpose_lethal_cells_within_device(costmap._h, b.ctypes.data, C.byref(h)))
            return cell_vector_device(h, costmap.device)
        ptr, n = C.c_void_p(), C.c_size_t(0)
        check(lib().dpose_lethal_cells_within(costmap._h, b.ctypes.data, C.byref(ptr), C.byref(n)))
        if n.value == 0:
            return np.empty((0, 2), dtype=np.int32)
        arr = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_int32)), (n.value * 2,)).reshape(-1, 2).copy()
        lib().dpose_free(ptr)
        return arr
    finally:
        if own is not None:
            own.close()


class cost_data:
    """dpose_core.hpp:79-102 — cost matrix, center cell and bounding box of a footprint."""

    def __init__(self, footprint, padding, device=0):
        fp = _i32(footprint).reshape(-1, 2)
        self._h = C.c_void_p()
        self.device = int(device)
        # the reference's cost_data accepts any polygon; the <3-vertex check lives in
        # pose_gradient (dpose_core.cpp:152-153).  The C ABI enforces it for both.
        rc = lib().dpose_pg_create(fp.ctypes.data, fp.shape[0], int(padding), self.device, C.byref(self._h))
        if rc == EINVAL:
            raise RuntimeError(lib().dpose_last_error().decode())
        check(rc)
        cost, rows, cols = C.c_void_p(), C.c_int(0), C.c_int(0)
        center = np.zeros(2, np.int32)
        box = np.zeros(10, np.int32)
        check(lib().dpose_pg_table(self._h, C.byref(cost), C.byref(rows), C.byref(cols), center.ctypes.data,
                                   box.ctypes.data))
        self.rows, self.cols = rows.value, cols.value
        n = self.rows * self.cols
        self._cost = np.ctypeslib.as_array(C.cast(cost, C.POINTER(C.c_float)), (n,)).reshape(self.rows, self.cols).copy()
        self._center = center
        self._box = box.reshape(5, 2)

    def get_data(self):
        return self._cost

    def get_center(self):
        return self._center

    def get_box(self):
        return self._box

    def stages(self):
        """(edt_sq int32, outline uint8, mask uint8, pre_blur float32) for bit-exact checks"""
        ptrs = [C.c_void_p() for _ in range(4)]
        check(lib().dpose_pg_stages(self._h, *[C.byref(p) for p in ptrs]))
        n, shape = self.rows * self.cols, (self.rows, self.cols)
        types = (C.c_int32, C.c_uint8, C.c_uint8, C.c_float)
        return tuple(np.ctypeslib.as_array(C.cast(p, C.POINTER(t)), (n,)).reshape(shape).copy()
                     for p, t in zip(ptrs, types))

    def precompute_us(self):
        us = C.c_float(0)
        check(lib().dpose_pg_precompute_us(self._h, C.byref(us)))
        return us.value

    def __del__(self):
        try:
            if self._h:
                lib().dpose_pg_destroy(self._h)
        except Exception:
            pass


class _Sweep(C.Structure):
    """dpose_sweep"""
    _fields_ = [("x0", C.c_double), ("dx", C.c_double), ("nx", C.c_uint64), ("y0", C.c_double), ("dy", C.c_double),
                ("ny", C.c_uint64), ("theta0", C.c_double), ("dtheta", C.c_double), ("ntheta", C.c_uint64)]


def sweep_poses(x, y, theta):
    """the lattice of pose_gradient.get_cost_sweep as an (n, 3) array, same bits as the kernel generates"""
    xs = x[0] + x[1] * np.arange(int(x[2]), dtype=np.float64)
    ys = y[0] + y[1] * np.arange(int(y[2]), dtype=np.float64)
    ts = theta[0] + theta[1] * np.arange(int(theta[2]), dtype=np.float64)
    T, Y, X = np.meshgrid(ts, ys, xs, indexing="ij")
    return np.stack([X.ravel(), Y.ravel(), T.ravel()], axis=1)


class pose_gradient:
    """dpose_core.hpp:133-167."""

    class parameter:
        def __init__(self, padding=2):
            self.padding = int(padding)

    def __init__(self, footprint, param=None, device=0):
        param = param or pose_gradient.parameter()
        if isinstance(param, int):
            param = pose_gradient.parameter(param)
        fp = _i32(footprint).reshape(-1, 2)
        if fp.shape[0] < 3:  # dpose_core.cpp:152-153
            raise RuntimeError("footprint must contain at least three points")
        self._data = cost_data(fp, param.padding, device)
        self.device = int(device)

    def get_data(self):
        return self._data

    def get_cost(self, se2, cells, jacobian=True):
        """cost (and J if jacobian) of one pose over an explicit cell list (dpose_core.cpp:200-260).
        cells: (n, 2) int32 array or a cell_vector_device.  Returns (cost, J) with J = None when
        jacobian is False."""
        p = np.ascontiguousarray(np.asarray(se2, dtype=np.float64).reshape(3))
        cost = C.c_double(0.0)
        J = np.zeros(3, dtype=np.float64)
        Jp = J.ctypes.data if jacobian else None
        if isinstance(cells, cell_vector_device):
            check(lib().dpose_pg_get_cost_cells_device(self._data._h, p.ctypes.data, cells._h, C.byref(cost), Jp))
        else:
            c = _i32(cells).reshape(-1, 2)
            check(lib().dpose_pg_get_cost_cells(self._data._h, p.ctypes.data, c.ctypes.data, c.shape[0],
                                                C.byref(cost), Jp))
        return cost.value, (J if jacobian else None)

    def get_cost_batch(self, costmap, poses, jacobian=True, out=None):
        """NEW: (n, 3) float64 poses -> (n, 4) float64 [cost, Jx, Jy, Jtheta] over all lethal cells
        of `costmap`.  Host buffers; H2D and D2H are part of the call."""
        p = np.ascontiguousarray(np.asarray(poses, dtype=np.float64).reshape(-1, 3))
        if out is None:
            out = np.empty((p.shape[0], 4), dtype=np.float64)
        assert out.flags.c_contiguous and out.dtype == np.float64 and out.shape == (p.shape[0], 4)
        check(lib().dpose_pg_get_cost_batch(self._data._h, costmap._h, p.ctypes.data, p.shape[0],
                                            out.ctypes.data, int(bool(jacobian))))
        return out

    def get_cost_batch_costs(self, costmap, poses, out=None):
        """cost only: (n, 3) poses -> (n,) costs; 8 bytes per pose travel back instead of 32"""
        p = np.ascontiguousarray(np.asarray(poses, dtype=np.float64).reshape(-1, 3))
        if out is None:
            out = np.empty(p.shape[0], dtype=np.float64)
        assert out.flags.c_contiguous and out.dtype == np.float64 and out.shape == (p.shape[0],)
        check(lib().dpose_pg_get_cost_batch_costs(self._data._h, costmap._h, p.ctypes.data, p.shape[0], out.ctypes.data))
        return out

    def get_cost_sweep(self, costmap, x, y, theta, jacobian=True, cost_only=False, out=None, d_out=None, stream=0):
        """a pose LATTICE generated inside the kernel (nothing travels to the device): x, y, theta are (start, step,
        count) triples; pose i = (x0 + dx ix, y0 + dy iy, theta0 + dtheta it), i = (it ny + iy) nx + ix.
        Returns (ntheta, ny, nx, 4) results, or (ntheta, ny, nx) costs if cost_only.  d_out: device pointer instead of
        a host result (asynchronous on `stream`)."""
        sw = _Sweep(float(x[0]), float(x[1]), int(x[2]), float(y[0]), float(y[1]), int(y[2]), float(theta[0]),
                    float(theta[1]), int(theta[2]))
        shape = (sw.ntheta, sw.ny, sw.nx) + (() if cost_only else (4,))
        if d_out is not None:
            check(lib().dpose_pg_get_cost_sweep_device(self._data._h, costmap._h, C.byref(sw), C.c_void_p(int(d_out)),
                                                       int(bool(jacobian)), int(bool(cost_only)), C.c_void_p(int(stream))))
            return None
        if out is None:
            out = np.empty(shape, dtype=np.float64)
        assert out.flags.c_contiguous and out.dtype == np.float64 and out.size == int(np.prod(shape))
        check(lib().dpose_pg_get_cost_sweep(self._data._h, costmap._h, C.byref(sw), out.ctypes.data, int(bool(jacobian)),
                                            int(bool(cost_only))))
        return out.reshape(shape)

    def get_cost_batch_device(self, costmap, d_poses, n, d_out, jacobian=True, stream=0, rows=None):
        """device pointers (ints), asynchronous on `stream` (a cudaStream_t as int).
        rows = (row_lo, row_hi): only lethal cells of these map rows count, and only these rows of a volatile map are
        re-read in front of the launch — for callers that shard their poses by map row (see `reach`, dist.row_band)."""
        if rows is None:
            check(lib().dpose_pg_get_cost_batch_device(self._data._h, costmap._h, C.c_void_p(int(d_poses)), int(n),
                                                       C.c_void_p(int(d_out)), int(bool(jacobian)),
                                                       C.c_void_p(int(stream))))
        else:
            check(lib().dpose_pg_get_cost_batch_device_rows(self._data._h, costmap._h, C.c_void_p(int(d_poses)), int(n),
                                                            C.c_void_p(int(d_out)), int(bool(jacobian)),
                                                            C.c_void_p(int(stream)), int(rows[0]), int(rows[1])))

    def reach(self):
        """rows / columns around a pose beyond which no cell contributes to its cost (dpose_pg_reach)"""
        r = C.c_uint32(0)
        check(lib().dpose_pg_reach(self._data._h, C.byref(r)))
        return int(r.value)

    def window_bytes(self, costmap, poses):
        p = np.ascontiguousarray(np.asarray(poses, dtype=np.float64).reshape(-1, 3))
        b = C.c_uint64(0)
        check(lib().dpose_pg_window_bytes(self._data._h, costmap._h, p.ctypes.data, p.shape[0], C.byref(b)))
        return int(b.value)


def order_poses_by_tile(poses, size_x, size_y):
    """permutation that orders (n, 3) poses by the 32 x 8-cell map tile of their position (dpose_order_poses_by_tile): a
    set-up step for pose sets evaluated against many versions of a map — `poses[perm]` runs faster through the batch
    kernel, and its contiguous slices are stripes of the map (Costmap rows for get_cost_batch_device(rows=...))"""
    p = np.ascontiguousarray(np.asarray(poses, dtype=np.float64).reshape(-1, 3))
    perm = np.empty(p.shape[0], dtype=np.uint64)
    check(lib().dpose_order_poses_by_tile(p.ctypes.data, p.shape[0], int(size_x), int(size_y), perm.ctypes.data))
    return perm.astype(np.int64)


def get_cost_batch_multi(pgs, maps, poses, jacobian=True, out=None):
    """pose batch sharded contiguously over len(pgs) GPUs (one pose_gradient + Costmap per device)."""
    p = np.ascontiguousarray(np.asarray(poses, dtype=np.float64).reshape(-1, 3))
    if out is None:
        out = np.empty((p.shape[0], 4), dtype=np.float64)
    n = len(pgs)
    pg_arr = (C.c_void_p * n)(*[pg._data._h for pg in pgs])
    map_arr = (C.c_void_p * n)(*[m._h for m in maps])
    check(lib().dpose_pg_get_cost_batch_multi(pg_arr, map_arr, n, p.ctypes.data, p.shape[0], out.ctypes.data,
                                              int(bool(jacobian))))
    return out


SOLVE_STATUS_MASK, SOLVE_FOOTPRINT_FREE, SOLVE_ACCEPTED = 0x3, 0x4, 0x8
SOLVE_CONVERGED, SOLVE_MAXITER, SOLVE_STALLED = 1, 2, 3


class _SolverOptions(C.Structure):
    _fields_ = [("max_iter", C.c_int), ("tol", C.c_double), ("dual_inf_tol", C.c_double), ("armijo_c1", C.c_double),
                ("lethal", C.c_uint8)]


class _SolveSummary(C.Structure):
    _fields_ = [("first", C.c_longlong), ("best", C.c_longlong), ("n_converged", C.c_ulonglong),
                ("n_accepted", C.c_ulonglong), ("best_f", C.c_double), ("best_pose", C.c_double * 3),
                ("first_pose", C.c_double * 3)]


def solver_options():
    """dpose_solver_options with the plugin's defaults (max_iter 20, tol 5: dpose_goal_tolerance.cpp:504-507)"""
    o = _SolverOptions()
    lib().dpose_solver_default_options(C.byref(o))
    return o


def sample_offsets(seed, lin_tolerance, rot_tolerance, n, first_zero=True):
    """uniform_pose_generator (dpose_goal_tolerance.cpp:333-373) seeded with `seed`: (n, 3) starting offsets"""
    out = np.zeros((int(n), 3))
    check(lib().dpose_sample_offsets(int(seed), float(lin_tolerance), float(rot_tolerance), int(n),
                                     int(bool(first_zero)), out.ctypes.data))
    return out


class problem:
    """dpose_goal_tolerance::problem (dpose_goal_tolerance.hpp:115-200, dpose_goal_tolerance.cpp:90-292) with the
    Ipopt TNLP evaluations batched over displacements.  Same member names and argument meaning; poses in cells."""

    CONFIG_KEYS = ("lin_tolerance", "rot_tolerance", "weight_goal_lin", "weight_goal_rot", "weight_start_lin",
                   "weight_start_rot")

    def __init__(self, pg, costmap):
        self._pg, self._map = pg, costmap  # keep the borrowed handles alive
        self._h = C.c_void_p()
        check(lib().dpose_problem_create(pg._data._h, costmap._h, C.byref(self._h)))

    def init(self, start, goal, config=None, **kw):
        """config: mapping with the DposeGoalToleranceConfig keys (cfg/DposeGoalTolerance.cfg:7-12, defaults 1);
        lin_tolerance in cells"""
        cfg = dict(config or {})
        cfg.update(kw)
        c = np.array([float(cfg.get(k, 1.0)) for k in self.CONFIG_KEYS], dtype=np.float64)
        s = np.ascontiguousarray(np.asarray(start, np.float64).reshape(3))
        g = np.ascontiguousarray(np.asarray(goal, np.float64).reshape(3))
        check(lib().dpose_problem_init(self._h, s.ctypes.data, g.ctypes.data, c.ctypes.data))

    def get_nlp_info(self):
        """(n, m, nnz_jac_g) — :183-191"""
        return 3, 2, 3

    def get_bounds_info(self):
        """x in [-1e6, 1e6]^3, g in [0, (lin_tol^2, rot_tol^2)] — :193-207"""
        gu = np.zeros(2)
        check(lib().dpose_problem_info(self._h, None, gu.ctypes.data))
        return np.full(3, -1e6), np.full(3, 1e6), np.zeros(2), gu

    @property
    def search_box(self):
        b = np.zeros(10, np.int32)
        check(lib().dpose_problem_info(self._h, b.ctypes.data, None))
        return b.reshape(5, 2)

    def eval_batch(self, x, want=("f", "grad_f", "g", "jac_g")):
        """n displacements -> dict of f (n), grad_f (n,3), g (n,2), jac_g (n,3)"""
        x = np.ascontiguousarray(np.asarray(x, np.float64).reshape(-1, 3))
        n = x.shape[0]
        shapes = {"f": (n,), "grad_f": (n, 3), "g": (n, 2), "jac_g": (n, 3)}
        out = {k: np.empty(shapes[k]) for k in want}
        ptr = lambda k: out[k].ctypes.data if k in out else None  # noqa: E731
        check(lib().dpose_problem_eval_batch(self._h, x.ctypes.data, n, ptr("f"), ptr("grad_f"), ptr("g"), ptr("jac_g")))
        return out

    def solve_batch(self, x0, options=None, trace=False, **kw):
        """the attempt loop of DposeGoalTolerance::preProcess (dpose_goal_tolerance.cpp:376-459) for all starting
        offsets x0 (n, 3) at once, device resident: projected BFGS descent on f, check_footprint of every final pose,
        ranking.  options / kw: solver_options fields (max_iter, tol, dual_inf_tol, armijo_c1, lethal).
        Returns a dict: x (n, 3), f (n), status (n, uint8 flags SOLVE_*), evals (n), summary (dict)[, trace]."""
        x0 = np.ascontiguousarray(np.asarray(x0, np.float64).reshape(-1, 3))
        n = x0.shape[0]
        opt = solver_options()
        for k, v in dict(options or {}, **kw).items():
            setattr(opt, k, v)
        x, f = np.empty((n, 3)), np.empty(n)
        status, evals = np.zeros(n, np.uint8), np.zeros(n, np.int32)
        summ = _SolveSummary()
        tr = np.zeros((n, opt.max_iter + 1, 12)) if trace else None
        check(lib().dpose_problem_solve_batch(self._h, x0.ctypes.data, n, C.byref(opt), x.ctypes.data, f.ctypes.data,
                                              status.ctypes.data, evals.ctypes.data, C.byref(summ),
                                              tr.ctypes.data if trace else None))
        out = {"x": x, "f": f, "status": status, "evals": evals,
               "summary": {"first": int(summ.first), "best": int(summ.best), "n_converged": int(summ.n_converged),
                           "n_accepted": int(summ.n_accepted), "best_f": float(summ.best_f),
                           "best_pose": np.array(summ.best_pose[:]), "first_pose": np.array(summ.first_pose[:])}}
        if trace:
            out["trace"] = tr
        return out

    def solve_us(self):
        """device time of the last solve_batch (kernels only), microseconds"""
        us = C.c_float(0)
        check(lib().dpose_problem_solve_us(self._h, C.byref(us)))
        return us.value

    # the single-x TNLP callbacks, for callers that keep Ipopt's one-iterate-at-a-time shape
    def eval_f(self, x):
        return float(self.eval_batch([x], want=("f",))["f"][0])

    def eval_grad_f(self, x):
        return self.eval_batch([x], want=("grad_f",))["grad_f"][0]

    def eval_g(self, x):
        return self.eval_batch([x], want=("g",))["g"][0]

    def eval_jac_g(self, x):
        return self.eval_batch([x], want=("jac_g",))["jac_g"][0]

    def __del__(self):
        try:
            if self._h:
                lib().dpose_problem_destroy(self._h)
        except Exception:
            pass


def check_footprint(costmap, footprint, cost=254):
    """dpose_costmap.cpp:460-582: True iff the polygon (cells, closed or open) lies inside the map and covers no
    cell equal to `cost`.  costmap: Costmap (or a uint8 array, uploaded on the fly)."""
    own = None
    if not isinstance(costmap, Costmap):
        own = costmap = Costmap(costmap)
    try:
        fp = _i32(footprint).reshape(-1, 2)
        ok = C.c_int(0)
        check(lib().dpose_check_footprint(costmap._h, fp.ctypes.data, fp.shape[0], int(cost), C.byref(ok)))
        return bool(ok.value)
    finally:
        if own is not None:
            own.close()


def check_footprint_batch(costmap, footprint, poses, cost=254):
    """the plugin's validation step (dpose_goal_tolerance.cpp:386-399) for many candidate poses: footprint (cells,
    robot frame) transformed by each pose, rounded, checked.  Returns a bool array."""
    fp = _i32(footprint).reshape(-1, 2)
    p = np.ascontiguousarray(np.asarray(poses, np.float64).reshape(-1, 3))
    ok = np.zeros(p.shape[0], np.uint8)
    check(lib().dpose_check_footprint_batch(costmap._h, fp.ctypes.data, fp.shape[0], p.ctypes.data, p.shape[0],
                                            int(cost), ok.ctypes.data))
    return ok.astype(bool)


def to_cell(world_poses, origin, resolution, size):
    """to_cell (dpose_goal_tolerance.cpp:294-309) for n metric poses (x, y, yaw): Costmap2D::worldToMap + yaw.
    Returns (cells (n, 3), ok (n,) bool); a False entry is the reference's "pose outside of the map"."""
    w = np.ascontiguousarray(np.asarray(world_poses, np.float64).reshape(-1, 3))
    out, ok = np.empty_like(w), np.zeros(w.shape[0], np.uint8)
    check(lib().dpose_world_to_cell(w.ctypes.data, w.shape[0], float(origin[0]), float(origin[1]), float(resolution),
                                    int(size[0]), int(size[1]), out.ctypes.data, ok.ctypes.data))
    return out, ok.astype(bool)


def to_msg(cell_poses, origin, resolution):
    """to_msg (dpose_goal_tolerance.cpp:311-330) for n cell poses: round to a cell, Costmap2D::mapToWorld (cell
    centre).  Returns (world (n, 3), ok (n,) bool); a False entry is the reference's "negative pose"."""
    c = np.ascontiguousarray(np.asarray(cell_poses, np.float64).reshape(-1, 3))
    out, ok = np.empty_like(c), np.zeros(c.shape[0], np.uint8)
    check(lib().dpose_cell_to_world(c.ctypes.data, c.shape[0], float(origin[0]), float(origin[1]), float(resolution),
                                    out.ctypes.data, ok.ctypes.data))
    return out, ok.astype(bool)
