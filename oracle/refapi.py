"""ctypes binding of oracle/_ref/libdpose_ref.so — the reference's own dpose_core sources, compiled.

TEST INFRASTRUCTURE ONLY.  oracle/Makefile (`make ref`) compiles /root/reference/dpose_core/src/
{dpose_core,dpose_costmap}.cpp from where they lie against the stand-in headers of oracle/refshim/ and
links refshim/bridge.cpp in front of them.  The library is git-ignored, built in the container that has
/root/reference and travels prebuilt to the GPU box.  `available()` says whether it can be used; callers
(tests, bench.py --impl reference) skip or fall back to the port when it cannot.

Same call shapes as oracle/capi.py so the tests can run one body against both.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_ref", "libdpose_ref.so")
REFERENCE_ROOT = os.environ.get("DPOSE_REFERENCE_ROOT", "/root/reference")
_SOURCES = ("dpose_core/src/dpose_core.cpp", "dpose_core/src/dpose_costmap.cpp",
            "dpose_core/src/dpose_core/dpose_core.hpp", "dpose_core/src/dpose_core/dpose_costmap.hpp",
            "dpose_goal_tolerance/src/dpose_goal_tolerance.cpp",
            "dpose_goal_tolerance/src/dpose_goal_tolerance/dpose_goal_tolerance.hpp")


def reference_present():
    return all(os.path.exists(os.path.join(REFERENCE_ROOT, s)) for s in _SOURCES)


def build(force=False):
    """(re)build oracle/_ref when the reference checkout is here; otherwise keep whatever travelled."""
    if reference_present():
        cmd = ["make", "-C", _HERE, "-s", "ref", "REF=" + REFERENCE_ROOT]
        if force:
            cmd.insert(1, "-B")
        subprocess.run(cmd, check=True)
    return _SO if os.path.exists(_SO) else None


def available():
    try:
        return lib() is not None
    except Exception:
        return False


_lib = None


def lib():
    global _lib
    if _lib is None:
        if build() is None:
            return None
        L = C.CDLL(_SO)
        vp, i32, u32, sz, dbl = C.c_void_p, C.c_int, C.c_uint32, C.c_size_t, C.c_double
        L.ref_pg_create.restype = vp
        L.ref_pg_create.argtypes = [vp, i32, C.c_uint]
        L.ref_pg_free.argtypes = [vp]
        L.ref_pg_info.argtypes = [vp, vp, vp, vp, vp]
        L.ref_pg_table.argtypes = [vp, vp]
        L.ref_get_cost.restype = dbl
        L.ref_get_cost.argtypes = [vp, vp, vp, sz, vp]
        L.ref_make_footprint.restype = i32
        L.ref_make_footprint.argtypes = [vp, i32, dbl, vp]
        L.ref_raytrace.restype = sz
        L.ref_raytrace.argtypes = [vp, vp, vp, sz]
        L.ref_lethal_cells_within.restype = sz
        L.ref_lethal_cells_within.argtypes = [vp, u32, u32, vp, vp]
        L.ref_free.argtypes = [vp]
        L.ref_is_inside.restype = i32
        L.ref_is_inside.argtypes = [u32, u32, vp, i32]
        L.ref_check_footprint.restype = i32
        L.ref_check_footprint.argtypes = [vp, u32, u32, vp, i32, C.c_uint8]
        L.ref_xb_trace.argtypes = [vp, vp, i32, vp, vp]
        L.ref_get_cost_batch.argtypes = [vp, vp, u32, u32, vp, sz, vp, i32, i32]
        L.ref_problem_create.restype = vp
        L.ref_problem_create.argtypes = [vp, u32, u32, vp, i32, C.c_uint]
        L.ref_problem_free.argtypes = [vp]
        L.ref_problem_init.restype = i32
        L.ref_problem_init.argtypes = [vp, vp, vp, vp]
        L.ref_problem_eval_batch.argtypes = [vp, vp, sz, vp, vp, vp, vp]
        L.ref_problem_info.argtypes = [vp] * 8
        L.ref_sample_offsets.argtypes = [u32, dbl, dbl, sz, vp]
        L.ref_solve_batch.argtypes = [vp, i32, vp, vp, C.c_uint8, vp, sz, vp, vp, vp, vp, vp]
        L.ref_to_cell.restype = i32
        L.ref_to_cell.argtypes = [u32, u32, dbl, dbl, dbl, vp, vp]
        L.ref_to_msg.restype = i32
        L.ref_to_msg.argtypes = [u32, u32, dbl, dbl, dbl, vp, vp]
        _lib = L
    return _lib


def _i32(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.int32))


def make_footprint(xy_m, res):
    """make_footprint (dpose_costmap.cpp:40-55)."""
    xy = np.ascontiguousarray(np.asarray(xy_m, dtype=np.float64).reshape(-1, 2))
    out = np.empty(xy.shape, dtype=np.int32)
    if lib().ref_make_footprint(xy.ctypes.data, xy.shape[0], float(res), out.ctypes.data) != 0:
        raise RuntimeError("resolution must be positive")
    return out


class PoseGradient:
    """pose_gradient + its cost_data (dpose_core.cpp:122-156, 200-260)."""

    def __init__(self, footprint, padding):
        fp = _i32(footprint).reshape(-1, 2)
        self._pg = lib().ref_pg_create(fp.ctypes.data, fp.shape[0], int(padding))
        if not self._pg:
            raise RuntimeError("footprint must contain at least three points")
        rows, cols = C.c_int(), C.c_int()
        center = np.zeros(2, dtype=np.int32)
        box = np.zeros(10, dtype=np.int32)
        lib().ref_pg_info(self._pg, C.byref(rows), C.byref(cols), center.ctypes.data, box.ctypes.data)
        self.rows, self.cols = rows.value, cols.value
        self.center, self.box = center, box.reshape(5, 2)
        self.cost = np.empty((self.rows, self.cols), dtype=np.float32)
        lib().ref_pg_table(self._pg, self.cost.ctypes.data)

    def __del__(self):
        try:
            if self._pg:
                lib().ref_pg_free(self._pg)
        except Exception:
            pass

    def get_cost(self, se2, cells, want_jacobian=True):
        se2 = np.ascontiguousarray(np.asarray(se2, dtype=np.float64))
        cells = _i32(cells).reshape(-1, 2)
        J = np.zeros(3, dtype=np.float64)
        cost = lib().ref_get_cost(self._pg, se2.ctypes.data, cells.ctypes.data, cells.shape[0],
                                  J.ctypes.data if want_jacobian else None)
        return cost, J

    def get_cost_batch(self, costmap, poses, want_jacobian=True, nthreads=1):
        m = np.ascontiguousarray(np.asarray(costmap, dtype=np.uint8))
        p = np.ascontiguousarray(np.asarray(poses, dtype=np.float64).reshape(-1, 3))
        out = np.zeros((p.shape[0], 4), dtype=np.float64)
        lib().ref_get_cost_batch(self._pg, m.ctypes.data, m.shape[1], m.shape[0], p.ctypes.data, p.shape[0],
                                 out.ctypes.data, int(want_jacobian), int(nthreads))
        return out


def raytrace(begin, end):
    """raytrace (dpose_costmap.cpp:57-104)."""
    b, e = _i32(begin), _i32(end)
    n = int(max(abs(int(e[0]) - int(b[0])), abs(int(e[1]) - int(b[1]))))
    out = np.empty((max(n, 1), 2), dtype=np.int32)
    k = lib().ref_raytrace(b.ctypes.data, e.ctypes.data, out.ctypes.data, out.shape[0])
    return out[:k].copy()


def lethal_cells_within(costmap, box, sort=True):
    """lethal_cells_within (dpose_costmap.cpp:117-184).  The reference walks a std::unordered_map of rows, so
    its row order is unspecified; sort=True returns the cells (y, x)-sorted like oracle/capi.py."""
    m = np.ascontiguousarray(np.asarray(costmap, dtype=np.uint8))
    b = _i32(box).reshape(10)
    sy, sx = (m.shape if m.ndim == 2 else (0, 0))
    ptr = C.c_void_p()
    n = lib().ref_lethal_cells_within(m.ctypes.data, sx, sy, b.ctypes.data, C.byref(ptr))
    arr = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_int32)), (max(n, 1) * 2,))[:2 * n].reshape(n, 2).copy()
    lib().ref_free(ptr)
    if sort and n:
        arr = arr[np.lexsort((arr[:, 0], arr[:, 1]))]
    return arr


def is_inside(size_x, size_y, footprint):
    fp = _i32(footprint).reshape(-1, 2)
    return bool(lib().ref_is_inside(int(size_x), int(size_y), fp.ctypes.data, fp.shape[0]))


def check_footprint(costmap, footprint, cost=254):
    """check_footprint (dpose_costmap.cpp:460-582)."""
    m = np.ascontiguousarray(np.asarray(costmap, dtype=np.uint8))
    fp = _i32(footprint).reshape(-1, 2)
    return bool(lib().ref_check_footprint(m.ctypes.data, m.shape[1], m.shape[0], fp.ctypes.data, fp.shape[0], int(cost)))


def xb_trace(c0, c1, steps):
    """x_bresenham(c0, c1): get_x() before each of `steps` advance_x() calls, and get_dx()."""
    a, b = _i32(c0), _i32(c1)
    xs = np.zeros(max(int(steps), 1), dtype=np.int32)
    dx = C.c_double()
    lib().ref_xb_trace(a.ctypes.data, b.ctypes.data, int(steps), xs.ctypes.data, C.byref(dx))
    return xs[:int(steps)].copy(), dx.value


CONFIG_KEYS = ("lin_tolerance", "rot_tolerance", "weight_goal_lin", "weight_goal_rot", "weight_start_lin", "weight_start_rot")


class Problem:
    """dpose_goal_tolerance::problem (dpose_goal_tolerance.cpp:90-292) of the compiled reference: the NLP the plugin
    hands to Ipopt.  footprint in cells (the bridge feeds make_footprint a resolution of 1)."""

    def __init__(self, footprint, padding, costmap):
        self.map = np.ascontiguousarray(np.asarray(costmap, dtype=np.uint8))  # viewed, not copied: keep it alive
        fp = _i32(footprint).reshape(-1, 2)
        self._fp, self._padding, self._clones = fp, int(padding), []
        self._h = lib().ref_problem_create(self.map.ctypes.data, self.map.shape[1], self.map.shape[0], fp.ctypes.data,
                                           fp.shape[0], int(padding))
        if not self._h:
            raise RuntimeError("footprint must contain at least three points")

    def init(self, start, goal, **cfg):
        self._start, self._goal, self._cfg = list(start), list(goal), dict(cfg)
        for c in self._clones:
            c.init(start, goal, **cfg)
        c = np.array([float(cfg.get(k, 1.0)) for k in CONFIG_KEYS])
        s = np.ascontiguousarray(np.asarray(start, np.float64).reshape(3))
        g = np.ascontiguousarray(np.asarray(goal, np.float64).reshape(3))
        if lib().ref_problem_init(self._h, s.ctypes.data, g.ctypes.data, c.ctypes.data) != 0:
            raise ValueError("both weights cannot be zero")

    def eval_batch(self, x):
        x = np.ascontiguousarray(np.asarray(x, np.float64).reshape(-1, 3))
        n = x.shape[0]
        f, grad, g, jac = np.empty(n), np.empty((n, 3)), np.empty((n, 2)), np.empty((n, 3))
        lib().ref_problem_eval_batch(self._h, x.ctypes.data, n, f.ctypes.data, grad.ctypes.data, g.ctypes.data, jac.ctypes.data)
        return f, grad, g, jac

    def solve_batch(self, cfg, x0, nthreads=1, lethal=254, trace=False):
        """CPU arm of the multi-start solve: the restated step rule (oracle_solve) on THIS compiled plugin's TNLP callbacks,
        then preProcessImpl's footprint validation with the reference's check_footprint; one `problem` per thread.
        cfg: oracle.capi.SolverCfg.  Returns (x, f, status, evals[, trace]) with the product's status encoding."""
        x0 = np.ascontiguousarray(np.asarray(x0, np.float64).reshape(-1, 3))
        n = x0.shape[0]
        nthreads = max(1, int(nthreads))
        while len(self._clones) < nthreads - 1:  # the callbacks cache cost_ / J_: every thread needs its own problem
            c = Problem(self._fp, self._padding, self.map)
            c.init(self._start, self._goal, **self._cfg)
            self._clones.append(c)
        hs = (C.c_void_p * nthreads)(*([self._h] + [c._h for c in self._clones[:nthreads - 1]]))
        x, f = np.empty((n, 3)), np.empty(n)
        status, evals = np.zeros(n, np.uint8), np.zeros(n, np.int32)
        tr = np.zeros((n, cfg.max_iter + 1, 12)) if trace else None
        goal = np.ascontiguousarray(np.asarray(self._goal, np.float64).reshape(3))
        lib().ref_solve_batch(hs, nthreads, goal.ctypes.data, C.byref(cfg), int(lethal), x0.ctypes.data, n, x.ctypes.data,
                              f.ctypes.data, status.ctypes.data, evals.ctypes.data, tr.ctypes.data if trace else None)
        return (x, f, status, evals, tr) if trace else (x, f, status, evals)

    def info(self):
        """(n, m, nnz_jac_g, index_style), x bounds, g bounds, Jacobian pattern (rows, cols)"""
        dims = np.zeros(4, np.int32)
        xl, xu, gl, gu = np.zeros(3), np.zeros(3), np.zeros(2), np.zeros(2)
        rows, cols = np.zeros(3, np.int32), np.zeros(3, np.int32)
        lib().ref_problem_info(self._h, dims.ctypes.data, xl.ctypes.data, xu.ctypes.data, gl.ctypes.data, gu.ctypes.data,
                               rows.ctypes.data, cols.ctypes.data)
        return dims, (xl, xu), (gl, gu), (rows, cols)

    def __del__(self):
        try:
            if self._h:
                lib().ref_problem_free(self._h)
        except Exception:
            pass


def sample_offsets(seed, lin_tolerance, rot_tolerance, n):
    """the plugin's uniform_pose_distribution (dpose_goal_tolerance.cpp:333-357) driven by std::mt19937(seed)"""
    out = np.zeros((int(n), 3))
    lib().ref_sample_offsets(int(seed), float(lin_tolerance), float(rot_tolerance), int(n), out.ctypes.data)
    return out


def to_cell(world, origin, resolution, size):
    """the plugin's to_cell (dpose_goal_tolerance.cpp:293-309): (x, y, theta) metres -> cells, None if it throws"""
    w = np.ascontiguousarray(np.asarray(world, np.float64).reshape(3))
    out = np.zeros(3)
    rc = lib().ref_to_cell(int(size[0]), int(size[1]), float(resolution), float(origin[0]), float(origin[1]), w.ctypes.data,
                           out.ctypes.data)
    return out if rc == 0 else None


def to_msg(cell, origin, resolution, size):
    """the plugin's to_msg (dpose_goal_tolerance.cpp:311-330): cell pose -> (x, y, yaw) metres, None if it throws"""
    c = np.ascontiguousarray(np.asarray(cell, np.float64).reshape(3))
    out = np.zeros(3)
    rc = lib().ref_to_msg(int(size[0]), int(size[1]), float(resolution), float(origin[0]), float(origin[1]), c.ctypes.data,
                          out.ctypes.data)
    return out if rc == 0 else None
