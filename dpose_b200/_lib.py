"""ctypes loader of libdpose_b200.so (the C ABI of include/dpose_b200.h).

There is no fallback: if the shared library is missing this raises, and on a box without
a CUDA device every compute entry point returns DPOSE_ENODEV, surfaced as DposeError.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "lib", "libdpose_b200.so")

OK, EINVAL, ECUDA, ENOMEM, ENODEV = 0, 1, 2, 3, 4

# every symbol include/dpose_b200.h declares: name -> (restype, argtypes)
_vp, _sz, _i, _u32, _u64, _d = C.c_void_p, C.c_size_t, C.c_int, C.c_uint32, C.c_uint64, C.c_double
_pp = C.POINTER(C.c_void_p)
SYMBOLS = {
    "dpose_last_error": (C.c_char_p, []),
    "dpose_device_count": (_i, [C.POINTER(_i)]),
    "dpose_launch_count": (_u64, []),
    "dpose_host_alloc": (_i, [_pp, _sz]),
    "dpose_host_free": (None, [_vp]),
    "dpose_free": (None, [_vp]),
    "dpose_make_footprint": (_i, [_vp, _i, _d, _vp]),
    "dpose_pg_create": (_i, [_vp, _i, _u32, _i, _pp]),
    "dpose_pg_destroy": (None, [_vp]),
    "dpose_pg_table": (_i, [_vp, _pp, C.POINTER(_i), C.POINTER(_i), _vp, _vp]),
    "dpose_pg_stages": (_i, [_vp, _pp, _pp, _pp, _pp]),
    "dpose_pg_precompute_us": (_i, [_vp, C.POINTER(C.c_float)]),
    "dpose_map_create": (_i, [_vp, _u32, _u32, _sz, _i, _pp]),
    "dpose_map_create_from_device": (_i, [_vp, _u32, _u32, _sz, _i, _pp]),
    "dpose_map_update": (_i, [_vp, _vp, _sz, _u32, _u32, _u32, _u32]),
    "dpose_map_set_volatile": (_i, [_vp, _i]),
    "dpose_map_device_ptr": (_i, [_vp, _pp, C.POINTER(_sz)]),
    "dpose_map_size": (_i, [_vp, C.POINTER(_u32), C.POINTER(_u32)]),
    "dpose_map_destroy": (None, [_vp]),
    "dpose_raytrace": (_i, [_vp, _vp, _vp, _sz, C.POINTER(_sz)]),
    "dpose_lethal_cells_within": (_i, [_vp, _vp, _pp, C.POINTER(_sz)]),
    "dpose_map_lethal_us": (_i, [_vp, C.POINTER(C.c_float)]),
    "dpose_lethal_cells_within_device": (_i, [_vp, _vp, _pp]),
    "dpose_cells_upload": (_i, [_vp, _sz, _i, _pp]),
    "dpose_cells_size": (_i, [_vp, C.POINTER(_sz)]),
    "dpose_cells_download": (_i, [_vp, _vp]),
    "dpose_cells_destroy": (None, [_vp]),
    "dpose_pg_get_cost_cells": (_i, [_vp, _vp, _vp, _sz, C.POINTER(_d), _vp]),
    "dpose_pg_get_cost_cells_device": (_i, [_vp, _vp, _vp, C.POINTER(_d), _vp]),
    "dpose_pg_get_cost_batch": (_i, [_vp, _vp, _vp, _sz, _vp, _i]),
    "dpose_pg_get_cost_batch_device": (_i, [_vp, _vp, _vp, _sz, _vp, _i, _vp]),
    "dpose_pg_get_cost_batch_device_rows": (_i, [_vp, _vp, _vp, _sz, _vp, _i, _vp, C.c_uint32, C.c_uint32]),
    "dpose_order_poses_by_tile": (_i, [_vp, _sz, C.c_uint32, C.c_uint32, _vp]),
    "dpose_pg_reach": (_i, [_vp, C.POINTER(C.c_uint32)]),
    "dpose_pg_get_cost_batch_costs": (_i, [_vp, _vp, _vp, _sz, _vp]),
    "dpose_pg_get_cost_sweep": (_i, [_vp, _vp, _vp, _vp, _i, _i]),
    "dpose_pg_get_cost_sweep_device": (_i, [_vp, _vp, _vp, _vp, _i, _i, _vp]),
    "dpose_pg_get_cost_batch_multi": (_i, [_vp, _vp, _i, _vp, _sz, _vp, _i]),
    "dpose_problem_create": (_i, [_vp, _vp, _pp]),
    "dpose_problem_destroy": (None, [_vp]),
    "dpose_problem_init": (_i, [_vp, _vp, _vp, _vp]),
    "dpose_problem_info": (_i, [_vp, _vp, _vp]),
    "dpose_problem_eval_batch": (_i, [_vp, _vp, _sz, _vp, _vp, _vp, _vp]),
    "dpose_problem_eval_batch_device": (_i, [_vp, _vp, _sz, _vp, _vp, _vp, _vp, _vp, _vp]),
    "dpose_solver_default_options": (None, [_vp]),
    "dpose_sample_offsets": (_i, [_u32, _d, _d, _sz, _i, _vp]),
    "dpose_problem_solve_batch": (_i, [_vp, _vp, _sz, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "dpose_problem_solve_us": (_i, [_vp, C.POINTER(C.c_float)]),
    "dpose_world_to_cell": (_i, [_vp, _sz, _d, _d, _d, _u32, _u32, _vp, _vp]),
    "dpose_cell_to_world": (_i, [_vp, _sz, _d, _d, _d, _vp, _vp]),
    "dpose_check_footprint": (_i, [_vp, _vp, _i, C.c_uint8, C.POINTER(_i)]),
    "dpose_check_footprint_batch": (_i, [_vp, _vp, _i, _vp, _sz, C.c_uint8, _vp]),
    "dpose_check_footprint_batch_device": (_i, [_vp, _vp, _i, _vp, _sz, C.c_uint8, _vp, _vp]),
    "dpose_pg_window_bytes": (_i, [_vp, _vp, _vp, _sz, C.POINTER(_u64)]),
}


class DposeError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"dpose_b200 error {code}: {msg}")
        self.code = code


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -m dpose_b200.build` "
                "(dpose_b200 has no CPU fallback)")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)  # AttributeError if the .so does not export it
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc):
    if rc != OK:
        raise DposeError(rc, lib().dpose_last_error().decode("utf-8", "replace"))
