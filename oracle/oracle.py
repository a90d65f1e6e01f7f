"""ctypes face of the C parity oracle (oracle/gs_oracle.c).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs.  The product package
(gridsplines_b200) never imports this module.

The classes mirror the reference's Scala names (Spline, AlwaysDefinedSpline,
OneDGrid, TwoDGrid, XDim/YDim ...) so parity tests read like the reference's own
tests (approximation/src/test/scala/approximation/*.scala).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Iterable, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libgs_oracle.so")

CONST, LINE, CUBIC = 0, 1, 2
ORTHO, RADIAL = 0, 1
TEMP, FLOW = 0, 1
UPP, LOW, RIGHT, LEFT = 0, 1, 2, 3
SPARSE, DENSE = 0, 1
SEL_APPLY, SEL_THERM, SEL_CAP, SEL_TEMPCOND = 0, 1, 2, 3
MAP_SINGLE, MAP_PER_ROW, MAP_PER_COL, MAP_DENSE = 0, 1, 2, 3

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
_vp = C.c_void_p


def build(force: bool = False) -> str:
    """Compile the oracle with gcc (strict IEEE: -ffp-contract=off)."""
    src = os.path.join(_HERE, "gs_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < max(
        os.path.getmtime(src), os.path.getmtime(os.path.join(_HERE, "gs_oracle.h"))
    ):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B"])
    return _LIB_PATH


def _load():
    build()
    lib = C.CDLL(_LIB_PATH)

    def sig(name, res, *args):
        f = getattr(lib, name)
        f.restype = res
        f.argtypes = list(args)

    d, i = C.c_double, C.c_int
    sig("gso_error", i)
    sig("gso_clear_error", None)
    sig("gso_spline_from_pieces", _vp, i, _ip, _dp, _dp, _dp, _dp)
    sig("gso_spline_const", _vp, d)
    sig("gso_spline_const_range", _vp, d, d, d)
    sig("gso_spline_line", _vp, d, d, d, d)
    sig("gso_spline_lines", _vp, i, _dp, _dp)
    sig("gso_spline_m1hermite3", _vp, i, _dp, _dp)
    sig("gso_spline_fhermite3", _vp, i, _dp, _dp)
    sig("gso_spline_div", _vp, _vp, _vp, i, i)
    sig("gso_spline_free", None, _vp)
    sig("gso_spline_size", i, _vp)
    sig("gso_spline_piece", None, _vp, i, _ip, _dp, _dp, _dp, _dp)
    for n in ("apply", "der", "antider"):
        sig("gso_spline_" + n, d, _vp, d)
    sig("gso_spline_area", d, _vp, d, d)
    sig("gso_spline_area_ascending", d, _vp, d, d)
    sig("gso_spline_average", d, _vp, d, d)
    sig("gso_spline_lower_bound", d, _vp)
    sig("gso_spline_upper_bound", d, _vp)
    sig("gso_spline_true_upper_bound", d, _vp)
    sig("gso_ads_create", _vp, _vp, i)
    sig("gso_ads_create_explicit", _vp, _vp, d, d, d, d)
    sig("gso_ads_free", None, _vp)
    sig("gso_ads_clamps", None, _vp, _dp)
    sig("gso_ads_apply", d, _vp, d)
    sig("gso_ads_area", d, _vp, d, d)
    sig("gso_ads_average", d, _vp, d, d)
    sig("gso_set_ascending_fold", None, i)
    sig("gso_set_use_libm_pow", None, i)
    sig("gso_ah", d, d, d, d)
    sig("gso_dh", d, d, d)
    sig("gso_r_at_half", d, d, d)
    sig("gso_vol", d, d, d, d)
    sig("gso_predef", None, i, d, _dp, i, d, d, _dp)
    sig("gso_heatflow_geometry", None, i, d, _dp, i, d, _dp)
    sig("gso_aval_1d", d, d, d)
    sig("gso_aval_2d", d, d, d)
    sig("gso_passion_iteration", None, d, d, _dp, i, d, _dp, _dp, C.POINTER(_vp), _dp, _dp)
    sig("gso_passion_iteration_single", None, d, d, _dp, i, d, _dp, _dp, _vp, _dp, _dp)
    sig("gso_grid1d_create", _vp, i, d, _dp, i, d, C.POINTER(_vp), d)
    sig("gso_grid1d_create_single", _vp, i, d, _dp, i, d, _vp, d, i)
    sig("gso_grid1d_free", None, _vp)
    sig("gso_grid1d_grid", _dp, _vp)
    sig("gso_grid1d_predict", _dp, _vp)
    sig("gso_grid1d_step", None, _vp, d, d, d)
    sig("gso_batch1d_step", None, i, d, _dp, i, d, _vp, d, i, _dp, _dp, _dp, _dp, d, i, i)
    sig("gso_grid2d_create", _vp, i, i, i, i, _dp, _dp)
    sig("gso_grid2d_free", None, _vp)
    sig("gso_grid2d_set_splines", None, _vp, i, C.POINTER(_vp))
    sig("gso_grid2d_set_map", None, _vp, i, i, _ip)
    sig("gso_grid2d_set_bound", None, _vp, i, i, i, _dp, i)
    sig("gso_grid2d_get_bound", None, _vp, i, _dp, i)
    sig("gso_grid2d_array", _dp, _vp, i)
    sig("gso_grid2d_fill", None, _vp, d)
    sig("gso_grid2d_xiter", None, _vp, d)
    sig("gso_grid2d_yiter", None, _vp, d)
    sig("gso_grid2d_update", None, _vp)
    sig("gso_grid2d_iteration", None, _vp, d)
    sig("gso_grid2d_no_heat_flow", None, _vp, i)
    sig("gso_grid2d_update_predicted", None, _vp)
    sig("gso_grid2d_edge_average", d, _vp, i)
    sig("gso_grid2d_av_value", d, _vp)
    return lib


lib = _load()


def _d(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float64)


def _p(a: np.ndarray):
    return a.ctypes.data_as(_dp)


def error() -> int:
    return lib.gso_error()


def clear_error() -> None:
    lib.gso_clear_error()


def set_ascending_fold(on: bool) -> None:
    lib.gso_set_ascending_fold(int(on))


def set_use_libm_pow(on: bool) -> None:
    lib.gso_set_use_libm_pow(int(on))


class Spline:
    """piecewise.Spline (P/Spline.scala)."""

    def __init__(self, handle):
        if not handle:
            raise ValueError("spline construction failed (None in the reference)")
        self._h = handle

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h and lib is not None:
            lib.gso_spline_free(h)

    # -- factories (P/Spline.scala:355-393)
    @staticmethod
    def const(value: float, lo: float | None = None, hi: float | None = None) -> "Spline":
        if lo is None:
            return Spline(lib.gso_spline_const(value))
        return Spline(lib.gso_spline_const_range(lo, hi, value))

    @staticmethod
    def line(xl, yl, xu, yu) -> "Spline":
        return Spline(lib.gso_spline_line(xl, yl, xu, yu))

    @staticmethod
    def lines(points: Sequence[tuple[float, float]]) -> "Spline":
        x, y = _d([p[0] for p in points]), _d([p[1] for p in points])
        return Spline(lib.gso_spline_lines(len(x), _p(x), _p(y)))

    @staticmethod
    def m1Hermite3(points: Sequence[tuple[float, float]]) -> "Spline":
        x, y = _d([p[0] for p in points]), _d([p[1] for p in points])
        return Spline(lib.gso_spline_m1hermite3(len(x), _p(x), _p(y)))

    @staticmethod
    def fHermite3(points: Sequence[tuple[float, float]]) -> "Spline":
        x, y = _d([p[0] for p in points]), _d([p[1] for p in points])
        return Spline(lib.gso_spline_fhermite3(len(x), _p(x), _p(y)))

    @staticmethod
    def from_pieces(kind, lo, hi, x0, coefs) -> "Spline":
        k = np.ascontiguousarray(kind, dtype=np.int32)
        lo, hi, x0, coefs = _d(lo), _d(hi), _d(x0), _d(coefs)
        return Spline(
            lib.gso_spline_from_pieces(len(k), k.ctypes.data_as(_ip), _p(lo), _p(hi), _p(x0), _p(coefs))
        )

    def div(self, other: "Spline", which: int = 2, true_upper: bool = False) -> "Spline":
        """`this / other` (P/Spline.scala:207-210); which: 0 const 1 lines 2 m1Hermite3 3 fHermite3."""
        return Spline(lib.gso_spline_div(self._h, other._h, which, int(true_upper)))

    # -- evaluation
    def __call__(self, x: float) -> float:
        return lib.gso_spline_apply(self._h, x)

    def der(self, x):
        return lib.gso_spline_der(self._h, x)

    def antider(self, x):
        return lib.gso_spline_antider(self._h, x)

    integral = antider

    def area(self, lo, hi):
        return lib.gso_spline_area(self._h, lo, hi)

    def area_ascending(self, lo, hi):
        return lib.gso_spline_area_ascending(self._h, lo, hi)

    def average(self, a, b):
        return lib.gso_spline_average(self._h, a, b)

    @property
    def size(self) -> int:
        return lib.gso_spline_size(self._h)

    @property
    def lowerBound(self):
        return lib.gso_spline_lower_bound(self._h)

    @property
    def upperBound(self):
        return lib.gso_spline_upper_bound(self._h)

    @property
    def trueUpperBound(self):
        return lib.gso_spline_true_upper_bound(self._h)

    def pieces(self):
        """[(kind, lo, hi, x0, [c0..c3])] -- what a host flattens into the C ABI's tables."""
        out = []
        k = C.c_int()
        lo, hi, x0 = C.c_double(), C.c_double(), C.c_double()
        c4 = (C.c_double * 4)()
        for p in range(self.size):
            lib.gso_spline_piece(self._h, p, C.byref(k), C.byref(lo), C.byref(hi), C.byref(x0), c4)
            out.append((k.value, lo.value, hi.value, x0.value, list(c4)))
        return out


class AlwaysDefinedSpline:
    """approximation.AlwaysDefinedSpline (A/AlwaysDefinedSpline.scala)."""

    def __init__(self, spl: Spline, true_upper: bool = False, clamps=None):
        self.spl = spl
        if clamps is None:
            self._h = lib.gso_ads_create(spl._h, int(true_upper))
        else:
            self._h = lib.gso_ads_create_explicit(spl._h, *clamps)

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h and lib is not None:
            lib.gso_ads_free(h)

    @property
    def clamps(self):
        o = (C.c_double * 4)()
        lib.gso_ads_clamps(self._h, o)
        return tuple(o)

    def __call__(self, x):
        return lib.gso_ads_apply(self._h, x)

    def area(self, lo, hi):
        return lib.gso_ads_area(self._h, lo, hi)

    def average(self, lo, hi):
        return lib.gso_ads_average(self._h, lo, hi)


def predef(direction: int, low: float, rng, upp: float, sigma: float = 1.0) -> np.ndarray:
    rng = _d(rng)
    out = np.empty((len(rng), 2))
    lib.gso_predef(direction, low, _p(rng), len(rng), upp, sigma, _p(out))
    return out


def heatflow_geometry(direction: int, low: float, rng, upp: float) -> np.ndarray:
    rng = _d(rng)
    out = np.empty(4)
    lib.gso_heatflow_geometry(direction, low, _p(rng), len(rng), upp, _p(out))
    return out


def passion_iteration(time, firstT, t, lastT, predict, preDef, conds) -> np.ndarray:
    """passion.iteration overload 1 (A/passion/package.scala:52-95) on one line: t, predict [n]; preDef [n][2];
    conds = n (left, right) AlwaysDefinedSpline pairs.  Returns result[n]."""
    t, predict, pd = _d(t), _d(predict), _d(preDef).reshape(-1)
    n = len(t)
    arr = (C.c_void_p * (2 * n))(*[x._h for pair in conds for x in pair])
    tp, res = np.zeros(2 * n), np.zeros(n)
    lib.gso_passion_iteration(time, firstT, _p(t), n, lastT, _p(predict), _p(pd), arr, _p(tp), _p(res))
    return res


class OneDGrid:
    """approximation.OneDGrid (A/OneDGrid.scala:14-34); `conductivity` a Spline (companion apply)
    or a list of n (left, right) AlwaysDefinedSpline pairs."""

    def __init__(self, leftX, rangeX, rightX, conductivity, sigma, direction=ORTHO, true_upper=False):
        self.rangeX = _d(rangeX)
        self.n = len(self.rangeX)
        self._keep = conductivity
        if isinstance(conductivity, Spline):
            self._h = lib.gso_grid1d_create_single(
                direction, leftX, _p(self.rangeX), self.n, rightX, conductivity._h, sigma, int(true_upper)
            )
        else:
            arr = (_vp * (2 * self.n))()
            for i, (a, b) in enumerate(conductivity):
                arr[2 * i], arr[2 * i + 1] = a._h, b._h
            self._h = lib.gso_grid1d_create(direction, leftX, _p(self.rangeX), self.n, rightX, arr, sigma)
        self.grid = np.ctypeslib.as_array(lib.gso_grid1d_grid(self._h), shape=(self.n,))
        self.predict = np.ctypeslib.as_array(lib.gso_grid1d_predict(self._h), shape=(self.n,))

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h and lib is not None:
            lib.gso_grid1d_free(h)

    def step(self, time, leftT, rightT):
        lib.gso_grid1d_step(self._h, time, leftT, rightT)


def batch1d_step(direction, leftX, rangeX, rightX, conductivity: Spline, sigma, grid, predict, leftT, rightT,
                 time, nsteps=1, nthreads=1):
    """nlines independent OneDGrid.step()s on line-major arrays grid[line, node] (in place)."""
    rangeX = _d(rangeX)
    assert grid.flags.c_contiguous and predict.flags.c_contiguous and grid.dtype == np.float64
    nlines, n = grid.shape
    leftT, rightT = _d(leftT), _d(rightT)
    lib.gso_batch1d_step(direction, leftX, _p(rangeX), n, rightX, conductivity._h, sigma, nlines, _p(grid),
                         _p(predict), _p(leftT), _p(rightT), time, nsteps, nthreads)


class TwoDGrid:
    """approximation.TwoDGrid (A/TwoDGrid.scala).  xvalues = low ++ range ++ upp."""

    def __init__(self, xvalues, yvalues, xdir=ORTHO, ydir=ORTHO):
        self.xvalues, self.yvalues = _d(xvalues), _d(yvalues)
        self.cols, self.rows = len(self.xvalues) - 2, len(self.yvalues) - 2
        self._h = lib.gso_grid2d_create(self.cols, self.rows, xdir, ydir, _p(self.xvalues), _p(self.yvalues))
        if not self._h:
            raise ValueError("grid needs at least 2 columns and 2 rows")
        shape = (self.rows, self.cols)
        self.grid = np.ctypeslib.as_array(lib.gso_grid2d_array(self._h, 0), shape=shape)
        self.predict = np.ctypeslib.as_array(lib.gso_grid2d_array(self._h, 1), shape=shape)
        self.result = np.ctypeslib.as_array(lib.gso_grid2d_array(self._h, 2), shape=shape)
        self._pool = []

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h and lib is not None:
            lib.gso_grid2d_free(h)

    def set_splines(self, pool: Sequence[AlwaysDefinedSpline]):
        self._pool = list(pool)
        arr = (_vp * len(pool))(*[a._h for a in pool])
        lib.gso_grid2d_set_splines(self._h, len(pool), arr)

    def set_map(self, selector: int, mode: int, ids):
        ids = np.ascontiguousarray(ids, dtype=np.int32).ravel()
        lib.gso_grid2d_set_map(self._h, selector, mode, ids.ctypes.data_as(_ip))

    def set_constant_coef(self, therm: AlwaysDefinedSpline, cap: AlwaysDefinedSpline, tempcond: AlwaysDefinedSpline):
        """ConstantCoef (A/TwoDGrid.scala:790-816): apply == tempConductivity."""
        self.set_splines([therm, cap, tempcond])
        self.set_map(SEL_APPLY, MAP_SINGLE, [2])
        self.set_map(SEL_THERM, MAP_SINGLE, [0])
        self.set_map(SEL_CAP, MAP_SINGLE, [1])
        self.set_map(SEL_TEMPCOND, MAP_SINGLE, [2])

    def set_bound(self, side: int, kind: int, rep: int, values):
        v = _d(np.atleast_1d(values))
        lib.gso_grid2d_set_bound(self._h, side, kind, rep, _p(v), len(v))

    def get_bound(self, side: int) -> np.ndarray:
        n = self.cols if side in (UPP, LOW) else self.rows
        out = np.empty(n)
        lib.gso_grid2d_get_bound(self._h, side, _p(out), n)
        return out

    def fill(self, v):  # `grid *= v`
        lib.gso_grid2d_fill(self._h, v)

    def xIter(self, time):
        lib.gso_grid2d_xiter(self._h, time)

    def yIter(self, time):
        lib.gso_grid2d_yiter(self._h, time)

    def update(self):
        lib.gso_grid2d_update(self._h)

    def iteration(self, time):
        lib.gso_grid2d_iteration(self._h, time)

    def noHeatFlow(self, side):
        lib.gso_grid2d_no_heat_flow(self._h, side)

    def updatePredicted(self):
        lib.gso_grid2d_update_predicted(self._h)

    def edgeAverage(self, side):
        return lib.gso_grid2d_edge_average(self._h, side)

    @property
    def avValue(self):
        return lib.gso_grid2d_av_value(self._h)


def iterate_range(low: float, increment, upp: float) -> np.ndarray:
    """XDim/YDim generator ctor (A/XDim.scala:10-15): iterate(inc(low))(inc).takeWhile(_ < upp)."""
    out = []
    x = increment(low)
    while x < upp:
        out.append(x)
        x = increment(x)
    return np.array(out, dtype=np.float64)
