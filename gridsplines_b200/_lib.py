"""ctypes binding of libgs_b200.so (include/gs_b200.h).

There is no fallback: if the shared library is missing, or no B200 is visible when a context
is created, the call raises.  Nothing in this package imports `oracle/`.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libgs_b200.so")

# enums of include/gs_b200.h
CONST, LINE, CUBIC = 0, 1, 2
ORTHO, RADIAL = 0, 1
TEMP, FLOW = 0, 1
SPARSE, DENSE = 0, 1
UPP, LOW, RIGHT, LEFT = 0, 1, 2, 3
SEL_APPLY, SEL_THERM, SEL_CAP, SEL_TEMPCOND = 0, 1, 2, 3
MAP_SINGLE, MAP_PER_ROW, MAP_PER_COL, MAP_DENSE = 0, 1, 2, 3
GRID, PREDICT, RESULT = 0, 1, 2
LINE_MAJOR, INTERLEAVED = 0, 1
SPLINE_SINGLE, SPLINE_PER_NODE = 0, 1
SOLVER_AUTO, SOLVER_THOMAS, SOLVER_PARTITIONED = 0, 1, 2
FLAG_SPLINE_UNDEFINED, FLAG_NOT_DOMINANT, FLAG_NONFINITE = 1, 2, 4

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
_vp = C.c_void_p
_vpp = C.POINTER(C.c_void_p)

# name -> (argtypes); every function returns int status unless listed in _SPECIAL
SIGNATURES = {
    "gs_ctx_create": (C.c_int, _vp, _vpp),
    "gs_ctx_destroy": (_vp,),
    "gs_ctx_synchronize": (_vp,),
    "gs_ctx_launch_count": (_vp, C.POINTER(C.c_uint64)),
    "gs_ctx_set_option": (_vp, C.c_char_p, C.c_int64),
    "gs_spline_create": (_vp, C.c_int, C.c_int, _dp, _dp, _dp, C.c_double, C.c_double, C.c_double, C.c_double, _ip),
    "gs_spline_eval_apply": (_vp, C.c_int, C.c_int, _dp, _dp),
    "gs_spline_eval_average": (_vp, C.c_int, C.c_int, _dp, _dp, _dp),
    "gs_spline_eval_status": (_vp, C.POINTER(C.c_uint32)),
    "gs_predef": (_vp, C.c_int, C.c_double, _dp, C.c_int, C.c_double, C.c_double, _dp),
    "gs_heatflow_geometry": (_vp, C.c_int, C.c_double, _dp, C.c_int, C.c_double, _dp),
    "gs_grid1d_create": (_vp, C.c_int64, C.c_int, C.c_int, C.c_double, _dp, C.c_double, C.c_double, C.c_int, _ip, _dp, _vpp),
    "gs_grid1d_destroy": (_vp,),
    "gs_grid1d_upload": (_vp, C.c_int, C.c_int, _vp),
    "gs_grid1d_download": (_vp, C.c_int, C.c_int, _vp),
    "gs_grid1d_fill": (_vp, C.c_int, C.c_double),
    "gs_grid1d_set_bc": (_vp, _dp, _dp, C.c_int),
    "gs_grid1d_step": (_vp, C.c_double, C.c_int),
    "gs_grid1d_step_host": (_vp, C.c_double, C.c_int, _vp, _dp, _dp, C.c_int),
    "gs_grid1d_set_solver": (_vp, C.c_int),
    "gs_grid1d_status": (_vp, C.POINTER(C.c_uint32)),
    "gs_grid1d_clear_status": (_vp,),
    "gs_grid1d_device_ptr": (_vp, C.c_int, _vpp, C.POINTER(C.c_int64)),
    "gs_grid1d_bytes_per_step": (_vp, _dp),
    "gs_sharded2d_create": (C.c_int, _ip, C.c_int, C.c_int, C.c_int, C.c_int, _dp, _dp, _vpp),
    "gs_sharded2d_destroy": (_vp,),
    "gs_sharded2d_device_count": (_vp, _ip),
    "gs_sharded2d_spline_create": (_vp, C.c_int, C.c_int, _dp, _dp, _dp, C.c_double, C.c_double, C.c_double, C.c_double, _ip),
    "gs_sharded2d_set_map": (_vp, C.c_int, C.c_int, _ip),
    "gs_sharded2d_set_bound": (_vp, C.c_int, C.c_int, C.c_int, _dp, C.c_int),
    "gs_sharded2d_fill": (_vp, C.c_double),
    "gs_sharded2d_upload": (_vp, C.c_int, _vp),
    "gs_sharded2d_download": (_vp, C.c_int, _vp),
    "gs_sharded2d_iteration": (_vp, C.c_double, C.c_int),
    "gs_sharded2d_iteration_timed": (_vp, C.c_double, C.c_int, _dp),
    "gs_sharded2d_synchronize": (_vp,),
    "gs_sharded2d_status": (_vp, C.POINTER(C.c_uint32)),
    "gs_sharded2d_launch_count": (_vp, C.POINTER(C.c_uint64)),
    "gs_sharded2d_set_option": (_vp, C.c_char_p, C.c_int64),
    "gs_grid2d_create": (_vp, C.c_int, C.c_int, C.c_int, C.c_int, _dp, _dp, _vpp),
    "gs_grid2d_destroy": (_vp,),
    "gs_grid2d_set_map": (_vp, C.c_int, C.c_int, _ip),
    "gs_grid2d_set_bound": (_vp, C.c_int, C.c_int, C.c_int, _dp, C.c_int),
    "gs_grid2d_get_bound": (_vp, C.c_int, _dp, C.c_int),
    "gs_grid2d_fill": (_vp, C.c_double),
    "gs_grid2d_upload": (_vp, C.c_int, _vp),
    "gs_grid2d_download": (_vp, C.c_int, _vp),
    "gs_grid2d_get_row": (_vp, C.c_int, C.c_int, _dp),
    "gs_grid2d_get_col": (_vp, C.c_int, C.c_int, _dp),
    "gs_grid2d_set_row": (_vp, C.c_int, _dp),
    "gs_grid2d_set_col": (_vp, C.c_int, _dp),
    "gs_grid2d_update_x": (_vp, _dp),
    "gs_grid2d_update_y": (_vp, _dp),
    "gs_grid2d_xiter": (_vp, C.c_double),
    "gs_grid2d_yiter": (_vp, C.c_double),
    "gs_grid2d_update": (_vp,),
    "gs_grid2d_yiter_update": (_vp, C.c_double),
    "gs_grid2d_yiter_update_begin": (_vp, C.c_double, _vp, _vp, _vp),
    "gs_grid2d_yiter_update_end": (_vp, C.c_double, _vp, C.c_int, C.c_int),
    "gs_grid2d_copy_from_device": (_vp, C.c_int, _vp),
    "gs_grid2d_copy_to_device": (_vp, C.c_int, _vp),
    "gs_grid2d_bind_device": (_vp, C.c_int, _vp),
    "gs_grid2d_iteration": (_vp, C.c_double, C.c_int),
    "gs_grid2d_update_predicted": (_vp,),
    "gs_grid2d_no_heat_flow": (_vp, C.c_int),
    "gs_grid2d_edge_average": (_vp, C.c_int, _dp),
    "gs_grid2d_av_value": (_vp, _dp),
    "gs_grid2d_set_solver": (_vp, C.c_int),
    "gs_grid2d_status": (_vp, C.POINTER(C.c_uint32)),
    "gs_grid2d_clear_status": (_vp,),
    "gs_grid2d_device_ptr": (_vp, C.c_int, _vpp),
    "gs_grid2d_bytes_per_step": (_vp, _dp),
    "gs_ragged_create": (_vp, C.c_int64, C.c_int, _ip, _ip, _dp, C.c_int, _ip, _vpp),
    "gs_ragged_destroy": (_vp,),
    "gs_ragged_upload": (_vp, C.c_int, _dp),
    "gs_ragged_download": (_vp, C.c_int, _dp),
    "gs_ragged_iteration": (_vp, C.c_double, _dp, _dp),
    "gs_ragged_status": (_vp, C.POINTER(C.c_uint32)),
}
_SPECIAL = {"gs_version": (C.c_int, ()), "gs_last_error": (C.c_char_p, ())}


class GsError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__(f"libgs_b200 status {status}: {message}")
        self.status = status


def load() -> C.CDLL:
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(make -C gridsplines_b200/csrc). gridsplines_b200 has no CPU fallback."
        )
    lib = C.CDLL(LIB_PATH)
    for name, args in SIGNATURES.items():
        f = getattr(lib, name)
        f.restype = C.c_int
        f.argtypes = list(args)
    for name, (res, args) in _SPECIAL.items():
        f = getattr(lib, name)
        f.restype = res
        f.argtypes = list(args)
    return lib


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        _lib = load()
    return _lib


def check(status: int) -> None:
    if status != 0:
        raise GsError(status, lib().gs_last_error().decode("utf-8", "replace"))
