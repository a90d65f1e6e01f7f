"""Host-side mirror of the reference's Scala surface for the time-stepper path, on top of the
C ABI (include/gs_b200.h).  Names and argument meaning follow the reference so that parity tests
read like approximation/src/test/scala/approximation/*.scala:

    Spline, AlwaysDefinedSpline           piecewise/.../Spline.scala, approximation/.../AlwaysDefinedSpline.scala
    XDim, YDim, Ortho, Radial             approximation/.../XDim.scala, YDim.scala, TypeDir.scala
    OneDGrid (+ BatchedOneDGrid)          approximation/.../OneDGrid.scala
    TwoDGrid, Temperature, HeatFlow, One, Many, Temp, Flow,
    ConstantCoef, VarXCoef, VarYCoef      approximation/.../TwoDGrid.scala

Host code here only builds tables and marshals arrays; every number the stepper produces comes
from the CUDA library.  There is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import math
import os
from typing import Callable, Iterable, Optional, Sequence

import numpy as np

from . import _lib as L
from ._lib import (CONST, CUBIC, DENSE, FLOW, GRID, INTERLEAVED, LEFT, LINE, LINE_MAJOR, LOW, MAP_DENSE,
                   MAP_PER_COL, MAP_PER_ROW, MAP_SINGLE, ORTHO, PREDICT, RADIAL, RESULT, RIGHT, SEL_APPLY,
                   SEL_CAP, SEL_TEMPCOND, SEL_THERM, SPARSE, SPLINE_PER_NODE, SPLINE_SINGLE, TEMP, UPP, GsError)

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
DBL_MAX = 1.7976931348623157e308


def _d(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float64)


def _p(a: np.ndarray):
    return a.ctypes.data_as(_dp)


# direction tags (approximation/package.scala:5-8)
Ortho, Radial = ORTHO, RADIAL


class Context:
    """One CUDA device + stream; owns the spline pool (gs_ctx)."""

    def __init__(self, device: int = 0, stream: Optional[int] = None):
        self._h = C.c_void_p()
        L.check(L.lib().gs_ctx_create(device, C.c_void_p(stream) if stream else None, C.byref(self._h)))
        self.device = device

    def close(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            L.lib().gs_ctx_destroy(h)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def synchronize(self):
        L.check(L.lib().gs_ctx_synchronize(self._h))

    @property
    def launch_count(self) -> int:
        c = C.c_uint64()
        L.check(L.lib().gs_ctx_launch_count(self._h, C.byref(c)))
        return c.value

    def set_option(self, name: str, value: int):
        L.check(L.lib().gs_ctx_set_option(self._h, name.encode(), int(value)))

    # K1
    def predef(self, direction, low, rng, upp, sigma=1.0) -> np.ndarray:
        rng = _d(rng)
        out = np.empty((len(rng), 2))
        L.check(L.lib().gs_predef(self._h, direction, low, _p(rng), len(rng), upp, sigma, _p(out)))
        return out

    def heatflow_geometry(self, direction, low, rng, upp) -> np.ndarray:
        rng = _d(rng)
        out = np.empty(4)
        L.check(L.lib().gs_heatflow_geometry(self._h, direction, low, _p(rng), len(rng), upp, _p(out)))
        return out


_default_ctx: Optional[Context] = None


def default_context() -> Context:
    global _default_ctx
    if _default_ctx is None:
        _default_ctx = Context(0)
    return _default_ctx


# ------------------------------------------------------------------------------------------
# piecewise.Spline: host-side table (kind, knots, x0, coefs).  Construction follows
# piecewise/.../Spline.scala:355-411, M1Hermite3.scala:76-126, Hermite3.scala:139-170, 256-272.
# ------------------------------------------------------------------------------------------
def _sq(x: float) -> float:
    # scala.math.pow(x, 2.0); HotSpot evaluates exponent 2 as x*x
    return x * x


def _jmin(a: float, b: float) -> float:
    if a != a:
        return a
    if a == 0.0 and b == 0.0:
        return a if math.copysign(1.0, a) < 0 else b
    return a if a <= b else b


def _fdiv(a: float, b: float) -> float:
    """IEEE-754 division (Python raises on x / 0.0; the JVM returns Inf/NaN)."""
    with np.errstate(all="ignore"):
        return float(np.float64(a) / np.float64(b))


def _signum(x: float) -> float:
    return x if (x == 0.0 or x != x) else (1.0 if x > 0.0 else -1.0)


class Spline:
    def __init__(self, kind: int, knots, x0, coefs):
        self.kind = kind
        self.knots = _d(knots)
        self.x0 = _d(x0)
        self.coefs = _d(coefs).reshape(-1, 4)
        assert len(self.knots) == len(self.x0) + 1 == len(self.coefs) + 1

    # -- factories
    @staticmethod
    def const(value: float, lo: float = -DBL_MAX, hi: float = DBL_MAX) -> "Spline":
        """Spline.const(value) = singleton over (Double.MinValue, Double.MaxValue) (Spline.scala:355-369)."""
        return Spline(CONST, [lo, hi], [0.0], [[value, 0.0, 0.0, 0.0]])

    @staticmethod
    def from_pieces(kind, knots, x0, coefs) -> "Spline":
        return Spline(kind, knots, x0, coefs)

    @staticmethod
    def _sorted(points):
        return sorted(((float(x), float(y)) for x, y in points), key=lambda p: p[0])  # stable, like sortBy

    @staticmethod
    def lines(points) -> "Spline":
        """Spline.lines (Spline.scala:384-385); Line.apply(xLow, xUpp, yLow, yUpp) Line.scala:69-73."""
        pts = Spline._sorted(points)
        knots, coefs = [pts[0][0]], []
        for (xl, yl), (xu, yu) in zip(pts[:-1], pts[1:]):
            slope = (yu - yl) / (xu - xl)
            intercept = yl if (xu - xl) == 0.0 else yl + (yu - yl) / (xu - xl) * (0.0 - xl)
            coefs.append([intercept, slope, 0.0, 0.0])
            knots.append(xu)
        return Spline(LINE, knots, [0.0] * len(coefs), coefs)

    @staticmethod
    def line(xl, yl, xu, yu) -> "Spline":
        return Spline.lines([(xl, yl), (xu, yu)])

    @staticmethod
    def lagrange2(points) -> "Spline":
        """Quadratic Lagrange pieces (P/Lagrange2.scala:62-113): every window of three consecutive points gives one
        parabola a x^2 + b x + c, coefficients by `polynomials` in its own operation order, evaluated about 0 by the
        FMA Horner rule (:20-21) -- on the device a GS_CUBIC piece with x0 = 0, c3 = 0 (same doubles).
        The reference lets consecutive windows overlap ((x_k, x_{k+2}) each); here window k owns [x_k, x_{k+1}) and
        the last one runs to the last point."""
        pts = Spline._sorted(points)
        if len(pts) < 3:
            raise ValueError("Must be 3 points")                     # IllegalArgumentException :67
        knots, coefs = [], []
        for (x0, y0), (x1, y1), (x2, y2) in zip(pts[:-2], pts[1:-1], pts[2:]):
            d = [y0 / ((x0 - x1) * (x0 - x2)), y1 / ((x1 - x0) * (x1 - x2)), y2 / ((x2 - x0) * (x2 - x1))]
            up = [[1.0, -(a + b), a * b] for a, b in ((x1, x2), (x0, x2), (x0, x1))]
            fin = [0.0, 0.0, 0.0]
            for k in range(3):
                for t in range(3):
                    fin[t] = fin[t] + up[k][t] * d[k]
            a_, b_, c_ = fin                                         # (a, b, c) -> new Lagrange2(Array(c, b, a))
            knots.append(x0)
            coefs.append([c_, b_, a_, 0.0])
        knots.append(pts[-1][0])
        return Spline(CUBIC, knots, [0.0] * len(coefs), coefs)

    @staticmethod
    def lagrange3(points) -> "Spline":
        """Natural cubic spline pieces (P/Lagrange3.scala:54-200, ListPassion.solve for the second-derivative terms):
        piece k over [x_k, x_{k+1}] is cubicRuleOfHorner(x - low, d, c, b, a) (:33-34) -- a GS_CUBIC piece with
        x0 = low.  The arithmetic follows the reference's list recursion literally."""
        v = Spline._sorted(points)
        n = len(v)
        if n < 3:
            raise ValueError("at least 3 points")
        hk = lambda x2, x1: x2 - x1
        f = lambda p1, p2: (p2[1] - p1[1]) / abs(p2[0] - p1[0]) if (p2[0] - p1[0]) != 0.0 else 0.0
        vect = lambda t1, t2, t3: 3 * f(t2, t3) - 3 * f(t1, t2)
        rows, vec = [], []
        for i in range(n - 2):                                       # cCoefs :76-100
            t1, t2, t3 = v[i], v[i + 1], v[i + 2]
            h1, h2 = hk(t2[0], t1[0]), hk(t3[0], t2[0])
            rows.append((h1, 2 * (h2 + h1), h2))
            vec.append(vect(t1, t2, t3))
        t1, t2 = v[n - 2], v[n - 1]
        h1 = hk(t2[0], t1[0])
        rows.append((h1, 2 * (0.0 + h1), 0.0))
        vec.append(vect(t1, t2, t2))
        # ListPassion.solve: forwardFirst / forwardUnit / forwardLast, backwardPassion; .head = all but the last value
        _, c0, d0 = rows[0]
        dl = [(-d0 / c0, vec[0] / c0)]
        for (b, c, d), r in zip(rows[1:-1], vec[1:-1]):
            assert abs(c) >= abs(d) + abs(b), "Passion is not correct or unstable"
            big = c + b * dl[-1][0]
            dl.append((-d / big, (r - b * dl[-1][1]) / big))
        b, c, d = rows[-1]
        big = c + b * dl[-1][0]
        dl.append((0.0, (vec[-1] - b * dl[-1][1]) / big))
        xs, prev = [], 0.0
        for delta, lam in reversed(dl):
            prev = delta * prev + lam
            xs.append(prev)
        sol = list(reversed(xs))
        cc = sol[:-1]                                                # solve(...).head: the last unknown is split off
        if len(cc) != n - 2:
            raise ValueError("list of values size minus 2 must be more than c list size")
        dd = []                                                      # dCoefs :103-125
        for i in range(n - 1):
            h = hk(v[i + 1][0], v[i][0])
            if i == 0:
                dd.append((cc[0] - 0.0) / 3 * h)
            elif i == n - 2:
                dd.append((0.0 - cc[i - 1]) / 3 * h)
            else:
                dd.append((cc[i - 1] - cc[i]) / 3 * h)
        bb = []                                                      # bCoefs :127-139 (stops when c or d run out)
        for i in range(min(n - 1, len(cc), len(dd))):
            h = hk(v[i + 1][0], v[i][0])
            bb.append(f(v[i], v[i + 1]) + cc[i] * _sq(h) - dd[i] * (h ** 3.0))
        aa = [p[1] for p in v[1:]]                                   # aCoefs :141-143
        knots, x0, coefs = [v[0][0]], [], []
        for i in range(len(cc)):                                     # while (c.nonEmpty) :61-69
            lo, hi = min(v[i][0], v[i + 1][0]), max(v[i][0], v[i + 1][0])
            coefs.append([dd[i], cc[i], bb[i], aa[i]])
            x0.append(lo)
            knots.append(hi)
        return Spline(CUBIC, knots, x0, coefs)

    @staticmethod
    def _sources(pts):
        """Hermite3.makeSources (Hermite3.scala:256-272): [yLow, yUpp, dLow, dUpp, xLow, xUpp]."""
        def deriv(a, b):
            return _fdiv(b[1] - a[1], b[0] - a[0])

        n = len(pts)
        src = []
        der1 = deriv(pts[0], pts[1])
        for k in range(n - 1):
            der2 = deriv(pts[k], pts[k + 2]) if k + 2 < n else deriv(pts[k], pts[k + 1])
            src.append([pts[k][1], pts[k + 1][1], der1, der2, pts[k][0], pts[k + 1][0]])
            der1 = der2
        return src

    @staticmethod
    def _cubic(src) -> "Spline":
        """constructSpline (M1Hermite3.scala:76-86 / Hermite3.scala:20-29)."""
        knots, x0, coefs = [src[0][4]], [], []
        for yl, yu, dl, du, xl, xu in src:
            delta = _fdiv(yu - yl, xu - xl)
            h = xu - xl
            coefs.append([yl, dl, _fdiv(-2.0 * dl - du + 3.0 * delta, h), _fdiv(dl + du - 2.0 * delta, _sq(h))])
            x0.append(xl)
            knots.append(xu)
        return Spline(CUBIC, knots, x0, coefs)

    @staticmethod
    def m1Hermite3(points) -> "Spline":
        """Spline.m1Hermite3 (Spline.scala:387-389): secant derivatives, Fritsch-Carlson
        monotonisation (Hermite3.monothone(Normal) :86-91, 139-170), neighbour smoothing
        (M1Hermite3.scala:88-121)."""
        pts = Spline._sorted(points)
        src = Spline._sources(pts)
        for s in src:
            yl, yu, dl, du, xl, xu = s
            h = xu - xl
            delta = _fdiv(yu - yl, h)
            alpha, beta = abs(_fdiv(dl, delta)), abs(_fdiv(du, delta))
            if any(math.isnan(v) or math.isinf(v) for v in (alpha, beta)):  # SmoothType.monotonize :70-73
                sa = sb = 0.0
            else:  # Normal.withType :86-91
                tau = _fdiv(3.0, math.sqrt(_sq(alpha) + _sq(beta)))
                sa, sb = _jmin(alpha, tau * alpha), _jmin(beta, tau * beta)
            s[2], s[3] = sa * delta, sb * delta
            if math.isnan(s[2]) or math.isnan(s[3]):
                # the reference throws here ("Function after monotonization must be monothone", :163-164)
                raise ValueError("m1Hermite3: degenerate piece (both end derivatives zero on a sloped piece)")
        for k in range(len(src) - 1):
            d_left, d_right = src[k][3], src[k + 1][2]
            der = _signum(d_left) * _jmin(abs(d_right), abs(d_left))
            src[k][3] = der
            src[k + 1][2] = der
        return Spline._cubic(src)

    @staticmethod
    def fHermite3(points) -> "Spline":
        return Spline._cubic(Spline._sources(Spline._sorted(points)))

    # -- bounds (Spline.scala:304-306)
    @property
    def size(self) -> int:
        return len(self.x0)

    @property
    def lowerBound(self) -> float:
        return float(self.knots[0])

    @property
    def upperBound(self) -> float:
        """AbsITree.greatestK returns the LOWER bound of the trailing UpBoundLeaf
        (AbsITree.scala:271-280): kept literally."""
        return float(self.knots[-2])

    @property
    def trueUpperBound(self) -> float:
        return float(self.knots[-1])

    # -- values and algebra.  Spline.apply on the host side of this package IS the device's (one arithmetic for
    #    setup and stepping): the table is registered unclamped in the context's pool and evaluated there.
    def apply(self, xs, ctx: Optional["Context"] = None) -> np.ndarray:
        ctx = ctx or default_context()
        xs = np.atleast_1d(_d(xs))
        sid, out = C.c_int(), np.empty_like(xs)
        L.check(L.lib().gs_spline_create(ctx._h, self.kind, self.size, _p(self.knots), _p(self.x0),
                                         _p(np.ascontiguousarray(self.coefs)), -math.inf, math.inf, math.nan, math.nan,
                                         C.byref(sid)))
        L.check(L.lib().gs_spline_eval_apply(ctx._h, sid.value, len(xs), _p(xs), _p(out)))
        return out

    def _sum_arguments(self, other: "Spline", true_upper: bool):
        """Spline.sumArguments (Spline.scala:198-205): both splines' piece bounds inside the common range, sorted,
        duplicates dropped.  `upp` follows the same literal rule as `upperBound` unless true_upper."""
        xs = [x for s in (other, self) for i in range(s.size) for x in (float(s.knots[i]), float(s.knots[i + 1]))]
        ua = self.trueUpperBound if true_upper else self.upperBound
        ub = other.trueUpperBound if true_upper else other.upperBound
        mn, mx = max(self.lowerBound, other.lowerBound), min(ua, ub)
        return sorted(set(x for x in xs if mn <= x <= mx))

    def combine(self, other: "Spline", op, factory=None, ctx=None, true_upper: bool = False) -> Optional["Spline"]:
        """Spline./ + - * (Spline.scala:207-225): resample `op(this(x), other(x))` at the merged piece bounds and
        rebuild with `factory` (the implicit PieceFunFactory; default M1Hermite3).  None where the reference returns
        None (fewer than two common points)."""
        xs = self._sum_arguments(other, true_upper)
        if len(xs) < 2:
            return None
        ya, yb = self.apply(xs, ctx), other.apply(xs, ctx)
        return (factory or Spline.m1Hermite3)([(x, op(float(a), float(b))) for x, a, b in zip(xs, ya, yb)])

    def __truediv__(self, other):
        return self.combine(other, lambda a, b: a / b)

    def __add__(self, other):
        return self.combine(other, lambda a, b: a + b)

    def __sub__(self, other):
        return self.combine(other, lambda a, b: a - b)

    def __mul__(self, other):
        return self.combine(other, lambda a, b: a * b)


class AlwaysDefinedSpline:
    """approximation.AlwaysDefinedSpline (AlwaysDefinedSpline.scala:4-46) registered in a context's
    device pool.  `clamps` = (lowerX, upperX, lowerY, upperY) overrides the literal rule."""

    def __init__(self, spl: Spline, ctx: Optional[Context] = None, true_upper: bool = False, clamps=None):
        self.ctx = ctx or default_context()
        self.spl = spl
        if clamps is None:
            lower_x = spl.lowerBound
            upper_x = spl.trueUpperBound if true_upper else spl.upperBound
            if spl.kind == CONST and spl.size == 1:
                lower_y = upper_y = float(spl.coefs[0][0])
            else:
                # lowerY = spl(lowerX), upperY = spl(upperX) (:7-8), evaluated by the device's own
                # Spline.apply through an unclamped registration of the same table
                raw = self._register(spl, -math.inf, math.inf, math.nan, math.nan)
                lower_y, upper_y = (float(v) for v in self._eval_apply(raw, [lower_x, upper_x]))
            clamps = (lower_x, upper_x, lower_y, upper_y)
        self.clamps = tuple(float(c) for c in clamps)
        self.id = self._register(spl, *self.clamps)

    def _register(self, spl, lx, ux, ly, uy) -> int:
        out = C.c_int()
        L.check(L.lib().gs_spline_create(self.ctx._h, spl.kind, spl.size, _p(spl.knots), _p(spl.x0),
                                         _p(np.ascontiguousarray(spl.coefs)), lx, ux, ly, uy, C.byref(out)))
        return out.value

    def _eval_flags(self) -> int:
        f = C.c_uint32()
        L.check(L.lib().gs_spline_eval_status(self.ctx._h, C.byref(f)))
        return f.value

    def _eval_apply(self, sid, xs) -> np.ndarray:
        xs = _d(xs)
        out = np.empty_like(xs)
        L.check(L.lib().gs_spline_eval_apply(self.ctx._h, sid, len(xs), _p(xs), _p(out)))
        if self._eval_flags() & 1:
            # "Spline is not defined at point ..." RuntimeException (P/intervaltree/AbsITree.scala:475-479)
            raise RuntimeError("spline is not defined at one of the arguments (NaN or outside the knots)")
        return out

    def __call__(self, x):
        xs = np.atleast_1d(_d(x))
        out = self._eval_apply(self.id, xs)
        return out if np.ndim(x) else float(out[0])

    def average(self, lo, hi):
        lo_a, hi_a = np.atleast_1d(_d(lo)), np.atleast_1d(_d(hi))
        out = np.empty_like(lo_a)
        L.check(L.lib().gs_spline_eval_average(self.ctx._h, self.id, len(lo_a), _p(lo_a), _p(hi_a), _p(out)))
        if self._eval_flags() & 1:
            raise RuntimeError("spline is not defined at one of the arguments (NaN or outside the knots)")
        return out if np.ndim(lo) else float(out[0])


def _as_ads(s, ctx) -> AlwaysDefinedSpline:
    if isinstance(s, AlwaysDefinedSpline):
        return s
    return AlwaysDefinedSpline(s, ctx)


# ------------------------------------------------------------------------------------------
# OneDGrid
# ------------------------------------------------------------------------------------------
class BatchedOneDGrid:
    """nlines independent OneDGrid (OneDGrid.scala:14-34) sharing geometry and conductivity.
    `conductivity`: a Spline / AlwaysDefinedSpline (companion apply, :66-79) or a list of n
    (left, right) AlwaysDefinedSpline pairs (`conductivities(z)(0..1)`)."""

    def __init__(self, leftX, rangeX, rightX, conductivity, sigma, direction=ORTHO, nlines=1,
                 ctx: Optional[Context] = None, predef=None):
        self.ctx = ctx or default_context()
        self.rangeX = _d(rangeX)
        self.n = len(self.rangeX)
        self.nlines = int(nlines)
        if isinstance(conductivity, (Spline, AlwaysDefinedSpline)):
            self._splines = [_as_ads(conductivity, self.ctx)]
            mode, ids = SPLINE_SINGLE, [self._splines[0].id]
        else:
            self._splines = [(_as_ads(a, self.ctx), _as_ads(b, self.ctx)) for a, b in conductivity]
            mode, ids = SPLINE_PER_NODE, [s.id for pair in self._splines for s in pair]
        ids = np.ascontiguousarray(ids, dtype=np.int32)
        pd = None if predef is None else _d(predef)
        self._h = C.c_void_p()
        L.check(L.lib().gs_grid1d_create(self.ctx._h, self.nlines, self.n, direction, leftX, _p(self.rangeX),
                                         rightX, sigma, mode, ids.ctypes.data_as(_ip),
                                         _p(pd) if pd is not None else None, C.byref(self._h)))

    def close(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            L.lib().gs_grid1d_destroy(h)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def upload(self, which: int, host: np.ndarray, layout: int = LINE_MAJOR):
        host = _d(host)
        assert host.size == self.nlines * self.n
        L.check(L.lib().gs_grid1d_upload(self._h, which, layout, host.ctypes.data))

    def download(self, which: int = GRID, layout: int = LINE_MAJOR, out: Optional[np.ndarray] = None) -> np.ndarray:
        shape = (self.nlines, self.n) if layout == LINE_MAJOR else (self.n, self.nlines)
        if out is None:
            out = np.empty(shape)
        L.check(L.lib().gs_grid1d_download(self._h, which, layout, out.ctypes.data))
        return out

    def upload_ptr(self, which: int, host_ptr: int, layout: int = LINE_MAJOR):
        L.check(L.lib().gs_grid1d_upload(self._h, which, layout, host_ptr))

    def download_ptr(self, which: int, host_ptr: int, layout: int = LINE_MAJOR):
        L.check(L.lib().gs_grid1d_download(self._h, which, layout, host_ptr))

    def fill(self, which: int, value: float):
        L.check(L.lib().gs_grid1d_fill(self._h, which, value))

    def set_bc(self, leftT, rightT):
        lt, rt = np.atleast_1d(_d(leftT)), np.atleast_1d(_d(rightT))
        stride = 0 if (lt.size == 1 and rt.size == 1) else 1
        if stride:
            lt = np.ascontiguousarray(np.broadcast_to(lt, (self.nlines,)))
            rt = np.ascontiguousarray(np.broadcast_to(rt, (self.nlines,)))
        L.check(L.lib().gs_grid1d_set_bc(self._h, _p(lt), _p(rt), stride))

    def step(self, time: float, leftT=None, rightT=None, nsteps: int = 1):
        if leftT is not None:
            self.set_bc(leftT, rightT)
        L.check(L.lib().gs_grid1d_step(self._h, time, nsteps))

    def step_host(self, time: float, host_grid, leftT, rightT, nsteps: int = 1):
        """OneDGrid.step with the state in a HOST array (numpy, or an int pointer to pinned memory): upload grid,
        step, download grid, pipelined over line chunks.  `host_grid` [nlines, n] is updated in place."""
        lt, rt = np.atleast_1d(_d(leftT)), np.atleast_1d(_d(rightT))
        stride = 0 if (lt.size == 1 and rt.size == 1) else 1
        if stride:
            lt = np.ascontiguousarray(np.broadcast_to(lt, (self.nlines,)))
            rt = np.ascontiguousarray(np.broadcast_to(rt, (self.nlines,)))
        ptr = host_grid if isinstance(host_grid, int) else host_grid.ctypes.data
        L.check(L.lib().gs_grid1d_step_host(self._h, time, nsteps, ptr, _p(lt), _p(rt), stride))

    def set_solver(self, solver: int):
        L.check(L.lib().gs_grid1d_set_solver(self._h, solver))

    @property
    def status(self) -> int:
        f = C.c_uint32()
        L.check(L.lib().gs_grid1d_status(self._h, C.byref(f)))
        return f.value

    def clear_status(self):
        L.check(L.lib().gs_grid1d_clear_status(self._h))

    @property
    def bytes_per_step(self) -> float:
        b = C.c_double()
        L.check(L.lib().gs_grid1d_bytes_per_step(self._h, C.byref(b)))
        return b.value


class OneDGrid(BatchedOneDGrid):
    """approximation.OneDGrid: `grid` is the public array (OneDGrid.scala:20), `step(time, leftT, rightT)`."""

    def __init__(self, leftX, rangeX, rightX, conductivity, sigma, direction=ORTHO, ctx=None):
        super().__init__(leftX, rangeX, rightX, conductivity, sigma, direction, 1, ctx)

    @property
    def grid(self) -> np.ndarray:
        return self.download(GRID)[0]

    @grid.setter
    def grid(self, values):
        self.upload(GRID, _d(values).reshape(1, -1))

    @property
    def predict(self) -> np.ndarray:
        return self.download(PREDICT)[0]


class RaggedLines:
    """A ragged batch of lines inside one flattened array: the fixed version of the reference's experimental
    passion.iteration overload 3 (A/passion/package.scala:161-227).  `lines[l]` = the positions of line l's nodes in the
    flattened matrix; `preDef[l]` = its (a, c) metric pairs (e.g. `predef(...)` of the line's coordinates);
    `conductivity` = one Spline / AlwaysDefinedSpline, or per line a list of (left, right) pairs."""

    def __init__(self, matrix_size: int, lines, preDef, conductivity, ctx: Optional[Context] = None):
        self.ctx = ctx or default_context()
        self.matrix_size = int(matrix_size)
        lengths = np.array([len(l) for l in lines], dtype=np.int32)
        indices = np.ascontiguousarray(np.concatenate([np.asarray(l, dtype=np.int32) for l in lines]))
        pd = np.ascontiguousarray(np.concatenate([_d(p).reshape(-1, 2) for p in preDef]))
        assert len(pd) == len(indices)
        if isinstance(conductivity, (Spline, AlwaysDefinedSpline)):
            self._splines = [_as_ads(conductivity, self.ctx)]
            mode, ids = SPLINE_SINGLE, np.array([self._splines[0].id], dtype=np.int32)
        else:
            pairs = [pr for line in conductivity for pr in line]
            self._splines = [(_as_ads(a, self.ctx), _as_ads(b, self.ctx)) for a, b in pairs]
            mode = SPLINE_PER_NODE
            ids = np.array([[a.id, b.id] for a, b in self._splines], dtype=np.int32).reshape(-1)
            assert len(ids) == 2 * len(indices)
        self.nlines = len(lengths)
        self._h = C.c_void_p()
        L.check(L.lib().gs_ragged_create(self.ctx._h, self.matrix_size, self.nlines, lengths.ctypes.data_as(L._ip),
                                         indices.ctypes.data_as(L._ip), _p(pd), mode, ids.ctypes.data_as(L._ip),
                                         C.byref(self._h)))

    def close(self):
        if getattr(self, "_h", None):
            L.lib().gs_ragged_destroy(self._h)
            self._h = None

    __del__ = close

    def _get(self, which) -> np.ndarray:
        out = np.empty(self.matrix_size)
        L.check(L.lib().gs_ragged_download(self._h, which, _p(out)))
        return out

    def _set(self, which, values):
        v = _d(values).reshape(-1)
        assert v.size == self.matrix_size
        L.check(L.lib().gs_ragged_upload(self._h, which, _p(v)))

    grid = property(lambda s: s._get(GRID), lambda s, v: s._set(GRID, v))
    predicted = property(lambda s: s._get(PREDICT), lambda s, v: s._set(PREDICT, v))
    result = property(lambda s: s._get(RESULT), lambda s, v: s._set(RESULT, v))

    def iteration(self, time: float, upper, lower):
        u, lo = _d(upper), _d(lower)
        assert u.size == self.nlines and lo.size == self.nlines
        L.check(L.lib().gs_ragged_iteration(self._h, time, _p(u), _p(lo)))

    @property
    def status(self) -> int:
        out = C.c_uint32()
        L.check(L.lib().gs_ragged_status(self._h, C.byref(out)))
        return out.value


# ------------------------------------------------------------------------------------------
# TwoDGrid
# ------------------------------------------------------------------------------------------
class _Dim:
    def __init__(self, low: float, rng_or_inc, upp: float, t: int = ORTHO):
        if callable(rng_or_inc):
            # generator ctor: Iterator.iterate(inc(low))(inc).takeWhile(_ < upp)  (XDim.scala:10-15)
            out, x = [], rng_or_inc(low)
            while x < upp:
                out.append(x)
                x = rng_or_inc(x)
            rng_or_inc = out
        self.low, self.upp, self.t = float(low), float(upp), t
        self.range = _d(rng_or_inc)

    @property
    def values(self) -> np.ndarray:  # Dim.values, Dim.scala:13
        return np.concatenate([[self.low], self.range, [self.upp]])

    def coord(self, i):
        return float(self.range[i])


class XDim(_Dim):
    @property
    def colsNum(self):
        return len(self.range)


class YDim(_Dim):
    @property
    def rowsNum(self):
        return len(self.range)


# bound-building tags of TwoDGrid.apply (TwoDGrid.scala:521-551)
One, Many = SPARSE, DENSE
Temp, Flow = TEMP, FLOW


class Coefficients:
    """Flattened Coefficients hierarchy (TwoDGrid.scala:782-1137): for each member a (mode, ids) pair."""

    def maps(self, cols: int, rows: int):
        raise NotImplementedError


class ConstantCoef(Coefficients):
    """ConstantCoef(thermConductivity, capacity, tempConductivity) (:790-816); apply == tempConductivity."""

    def __init__(self, therm, cap, tempcond, ctx=None):
        ctx = ctx or default_context()
        self.therm, self.cap, self.tempcond = _as_ads(therm, ctx), _as_ads(cap, ctx), _as_ads(tempcond, ctx)

    def maps(self, cols, rows):
        return {SEL_APPLY: (MAP_SINGLE, [self.tempcond.id]), SEL_THERM: (MAP_SINGLE, [self.therm.id]),
                SEL_CAP: (MAP_SINGLE, [self.cap.id]), SEL_TEMPCOND: (MAP_SINGLE, [self.tempcond.id])}


class _VarCoef(Coefficients):
    mode = MAP_PER_ROW

    def __init__(self, therm: Sequence, cap: Sequence, tempcond: Sequence, ctx=None):
        ctx = ctx or default_context()
        self.therm = [_as_ads(s, ctx) for s in therm]
        self.cap = [_as_ads(s, ctx) for s in cap]
        self.tempcond = [_as_ads(s, ctx) for s in tempcond]

    def maps(self, cols, rows):
        n = (rows if self.mode == MAP_PER_ROW else cols) + 1
        assert len(self.therm) == len(self.cap) == len(self.tempcond) == n, "one spline per index from -1"
        ids = lambda lst: [s.id for s in lst]
        return {SEL_APPLY: (self.mode, ids(self.tempcond)), SEL_THERM: (self.mode, ids(self.therm)),
                SEL_CAP: (self.mode, ids(self.cap)), SEL_TEMPCOND: (self.mode, ids(self.tempcond))}


class VarYCoef(_VarCoef):
    """per-row splines, index row+1 (:897-909)"""
    mode = MAP_PER_ROW


class VarXCoef(_VarCoef):
    """The reference's VarXCoef, literally: its arrays are built per COLUMN interval (xDim.values.sliding(2), cols + 1
    entries, :878-893) but every member indexes them by y + 1 (:862-874) -- so it behaves as a per-row
    map over the first rows + 1 entries and throws (ArrayIndexOutOfBounds -> IndexError here) when rows > cols.  Kept
    literal for parity; `VarXCoefByColumn` is the per-column reading the name suggests."""
    mode = MAP_PER_ROW

    def maps(self, cols, rows):
        n = rows + 1
        if min(len(self.therm), len(self.cap), len(self.tempcond)) < n:
            raise IndexError("VarXCoef indexes its per-column arrays by y + 1 (TwoDGrid.scala:862-874): "
                             f"{n} entries needed, {len(self.tempcond)} given")
        ids = lambda lst: [s.id for s in lst[:n]]
        return {SEL_APPLY: (MAP_PER_ROW, ids(self.tempcond)), SEL_THERM: (MAP_PER_ROW, ids(self.therm)),
                SEL_CAP: (MAP_PER_ROW, ids(self.cap)), SEL_TEMPCOND: (MAP_PER_ROW, ids(self.tempcond))}


class VarXCoefByColumn(_VarCoef):
    """per-column splines, index col + 1: what VarXCoef's constructor arguments describe (not what the reference's
    members return, see VarXCoef)."""
    mode = MAP_PER_COL


class _PatchCoef(Coefficients):
    """PatchXCoef / PatchYCoef (:895-1089): per-interval arrays that apply inside rangeX x rangeY only (`contains`,
    closed-open index intervals).  Only meaningful inside a Patched*Coef."""
    apply_by_x = False

    def __init__(self, therm: Sequence, cap: Sequence, tempcond: Sequence, rangeX, rangeY, ctx=None):
        ctx = ctx or default_context()
        self.therm = [_as_ads(s, ctx) for s in therm]
        self.cap = [_as_ads(s, ctx) for s in cap]
        self.tempcond = [_as_ads(s, ctx) for s in tempcond]
        self.rangeX, self.rangeY = tuple(rangeX), tuple(rangeY)

    def contains(self, x: int, y: int) -> bool:    # Interval.contains on CsdOpn: lower <= i < upper
        return self.rangeX[0] <= x < self.rangeX[1] and self.rangeY[0] <= y < self.rangeY[1]

    def member(self, sel: int, x: int, y: int) -> AlwaysDefinedSpline:
        if sel == SEL_THERM:
            return self.therm[y + 1]
        if sel == SEL_CAP:
            return self.cap[y + 1]
        if sel == SEL_TEMPCOND:
            return self.tempcond[y + 1]
        return self.tempcond[(x if self.apply_by_x else y) + 1]


class PatchXCoef(_PatchCoef):
    """:895-947 -- every member, `apply` included, indexes by y + 1."""


class PatchYCoef(_PatchCoef):
    """:1019-1056 -- thermConductivity / capacity / tempConductivity index by y + 1, `apply` by x + 1 (:1056)."""
    apply_by_x = True


class _PatchedCoef(Coefficients):
    """PatchedXCoef(xCoefs, path) / PatchedYCoef(yCoefs, path) (:1091-1137): the patch where it `contains` the node,
    the base map elsewhere.  Flattened to DENSE id maps (what a Scala host does by calling coefs(col, row) for every
    index from -1)."""

    def __init__(self, base: _VarCoef, path: _PatchCoef):
        self.base, self.path = base, path

    def maps(self, cols, rows):
        base = self.base.maps(cols, rows)
        out = {}
        for sel in (SEL_APPLY, SEL_THERM, SEL_CAP, SEL_TEMPCOND):
            mode, ids = base[sel]
            a = np.empty((rows + 1, cols + 1), dtype=np.int32)
            for y in range(-1, rows):
                for x in range(-1, cols):
                    if self.path.contains(x, y):
                        a[y + 1, x + 1] = self.path.member(sel, x, y).id
                    else:
                        a[y + 1, x + 1] = ids[(y if mode == MAP_PER_ROW else x) + 1]
            out[sel] = (MAP_DENSE, a.ravel())
        return out


class PatchedXCoef(_PatchedCoef):
    pass


class PatchedYCoef(_PatchedCoef):
    pass


class DenseCoef(Coefficients):
    """Arbitrary coefs(col,row): id arrays of shape (rows+1, cols+1), entry [row+1, col+1] (Patch*Coef)."""

    def __init__(self, apply_ids, therm_ids, cap_ids, tempcond_ids):
        self.ids = {SEL_APPLY: apply_ids, SEL_THERM: therm_ids, SEL_CAP: cap_ids, SEL_TEMPCOND: tempcond_ids}

    def maps(self, cols, rows):
        return {k: (MAP_DENSE, np.asarray(v, dtype=np.int32).reshape(rows + 1, cols + 1).ravel())
                for k, v in self.ids.items()}


class _Bound:
    def __init__(self, grid: "TwoDGrid", side: int, kind: int, rep: int):
        self.grid, self.side, self.kind, self.rep = grid, side, kind, rep

    def update(self, values):
        v = np.atleast_1d(_d(values))
        L.check(L.lib().gs_grid2d_set_bound(self.grid._h, self.side, self.kind, self.rep, _p(v), len(v)))

    def get(self, i: Optional[int] = None):
        n = self.grid.cols if self.side in (UPP, LOW) else self.grid.rows
        out = np.empty(n)
        L.check(L.lib().gs_grid2d_get_bound(self.grid._h, self.side, _p(out), n))
        return out if i is None else float(out[i])


class _Bounds:
    def __init__(self, upp, low, right, left):
        self.upp, self.low, self.right, self.left = upp, low, right, left


class _GridArrays:
    """TwoDGrid.Grid (:461-519): grid / predict / result as host copies on demand."""

    def __init__(self, owner: "TwoDGrid"):
        self._o = owner

    def _get(self, which):
        out = np.empty((self._o.rows, self._o.cols))
        L.check(L.lib().gs_grid2d_download(self._o._h, which, out.ctypes.data))
        return out

    def _set(self, which, values):
        v = _d(values)
        assert v.size == self._o.rows * self._o.cols
        L.check(L.lib().gs_grid2d_upload(self._o._h, which, v.ctypes.data))

    grid = property(lambda s: s._get(GRID), lambda s, v: s._set(GRID, v))
    predict = property(lambda s: s._get(PREDICT), lambda s, v: s._set(PREDICT, v))
    result = property(lambda s: s._get(RESULT), lambda s, v: s._set(RESULT, v))

    def update(self):
        L.check(L.lib().gs_grid2d_update(self._o._h))

    @property
    def avValue(self) -> float:
        out = C.c_double()
        L.check(L.lib().gs_grid2d_av_value(self._o._h, C.byref(out)))
        return out.value


class TwoDGrid:
    """approximation.TwoDGrid.  TwoDGrid(x, y)(upCount, upKind)(lowCount, lowKind)(rightCount,
    rightKind)(leftCount, leftKind)(coefs) of the reference (:553-574) becomes keyword pairs."""

    def __init__(self, x: XDim, y: YDim, upp=(One, Temp), low=(One, Temp), right=(One, Temp), left=(One, Temp),
                 coefs: Optional[Coefficients] = None, ctx: Optional[Context] = None):
        self.ctx = ctx or default_context()
        self.x, self.y = x, y
        self.cols, self.rows = x.colsNum, y.rowsNum
        xv, yv = _d(x.values), _d(y.values)
        self._h = C.c_void_p()
        L.check(L.lib().gs_grid2d_create(self.ctx._h, self.cols, self.rows, x.t, y.t, _p(xv), _p(yv), C.byref(self._h)))
        self.grid = _GridArrays(self)
        mk = lambda side, spec: _Bound(self, side, spec[1], spec[0])
        self.bounds = _Bounds(mk(UPP, upp), mk(LOW, low), mk(RIGHT, right), mk(LEFT, left))
        for b in (self.bounds.upp, self.bounds.low, self.bounds.right, self.bounds.left):
            b.update(0.0)
        self.coefs = coefs
        if coefs is not None:
            self.set_coefs(coefs)
        # the class body ends with noHeatFlow on all four sides (:446-449)
        for side in (UPP, LOW, RIGHT, LEFT):
            self.noHeatFlow(side)

    def close(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            L.lib().gs_grid2d_destroy(h)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_coefs(self, coefs: Coefficients):
        self.coefs = coefs
        for sel, (mode, ids) in coefs.maps(self.cols, self.rows).items():
            ids = np.ascontiguousarray(ids, dtype=np.int32)
            L.check(L.lib().gs_grid2d_set_map(self._h, sel, mode, ids.ctypes.data_as(_ip)))

    def fill(self, v: float):  # `grid *= v` (:84, 488-495)
        L.check(L.lib().gs_grid2d_fill(self._h, v))

    def xIter(self, time: float):
        L.check(L.lib().gs_grid2d_xiter(self._h, time))

    def yIter(self, time: float):
        L.check(L.lib().gs_grid2d_yiter(self._h, time))

    def iteration(self, time: float, nsteps: int = 1):
        L.check(L.lib().gs_grid2d_iteration(self._h, time, nsteps))

    def yIterUpdate(self, time: float):
        """yIter + Grid.update in one pass (the second half of iteration)."""
        L.check(L.lib().gs_grid2d_yiter_update(self._h, time))

    def yIterUpdateBegin(self, time: float, halo_prev: int | None, halo_next: int | None, exported: int):
        """First half of the y sweep of a row-sharded grid (device pointers; None = physical boundary on that end):
        writes this rank's 6 x cols interface values to `exported` for the all-gather."""
        L.check(L.lib().gs_grid2d_yiter_update_begin(self._h, time, halo_prev, halo_next, exported))

    def yIterUpdateEnd(self, time: float, gathered: int, nranks: int, rank: int):
        """Second half: `gathered` = [nranks][6][cols] device doubles, every rank's `exported` in rank order."""
        L.check(L.lib().gs_grid2d_yiter_update_end(self._h, time, gathered, nranks, rank))

    def upload_ptr(self, which: int, host_ptr: int):
        """gs_grid2d_upload from a raw host pointer (e.g. pinned memory): rows * cols doubles, row-major."""
        L.check(L.lib().gs_grid2d_upload(self._h, which, host_ptr))

    def download_ptr(self, which: int, host_ptr: int):
        L.check(L.lib().gs_grid2d_download(self._h, which, host_ptr))

    def copy_from_device(self, which: int, dptr: int):
        L.check(L.lib().gs_grid2d_copy_from_device(self._h, which, dptr))

    def copy_to_device(self, which: int, dptr: int):
        L.check(L.lib().gs_grid2d_copy_to_device(self._h, which, dptr))

    def bind_device(self, which: int, dptr: int):
        L.check(L.lib().gs_grid2d_bind_device(self._h, which, dptr))

    def updatePredicted(self):
        L.check(L.lib().gs_grid2d_update_predicted(self._h))

    def noHeatFlow(self, side: int):
        L.check(L.lib().gs_grid2d_no_heat_flow(self._h, side))

    # ---- slices (TwoDGrid.row / col / left / right / upper / lower, updateRow / updateColumn :260-352; they address
    #      `grid`), device-side gather / scatter + one contiguous copy
    def row(self, rowNum: int, copyTo: Optional[np.ndarray] = None, which: int = GRID) -> np.ndarray:
        out = np.empty(self.cols) if copyTo is None else copyTo
        assert out.dtype == np.float64 and out.flags.c_contiguous and out.size >= self.cols
        L.check(L.lib().gs_grid2d_get_row(self._h, which, rowNum, _p(out)))
        return out

    def col(self, colNum: int, copyTo: Optional[np.ndarray] = None, which: int = GRID) -> np.ndarray:
        out = np.empty(self.rows) if copyTo is None else copyTo
        assert out.dtype == np.float64 and out.flags.c_contiguous and out.size >= self.rows
        L.check(L.lib().gs_grid2d_get_col(self._h, which, colNum, _p(out)))
        return out

    def left(self, colNum: int = 0, copyTo=None):
        return self.col(colNum, copyTo)

    def right(self, colNumFromRight: int = 0, copyTo=None):
        return self.col(self.cols - 1 - colNumFromRight, copyTo)

    def upper(self, rowNum: int = 0, copyTo=None):
        return self.row(rowNum, copyTo)

    def lower(self, rowFromLast: int = 0, copyTo=None):
        return self.row(self.rows - 1 - rowFromLast, copyTo)

    def updateRow(self, rowNum: int, copyFrom):
        v = _d(copyFrom)
        assert v.size >= self.cols
        L.check(L.lib().gs_grid2d_set_row(self._h, rowNum, _p(v)))

    def updateColumn(self, colNum: int, copyFrom):
        v = _d(copyFrom)
        assert v.size >= self.rows
        L.check(L.lib().gs_grid2d_set_col(self._h, colNum, _p(v)))

    def updateX(self, f):
        """grid(col, row) = f(x_col) for every row (TwoDGrid.updateX :29-41; `grid` only).  The reference's
        ColumnIterator bounds its walk by rowsNum * (colsNum - 1) (A/YDim.scala:42,51), which is the last row only on
        square grids; this mirror fills every row of every column, the evident intent."""
        v = _d([f(self.x.coord(c)) for c in range(self.cols)])
        L.check(L.lib().gs_grid2d_update_x(self._h, _p(v)))

    def updateY(self, f):
        """grid(col, row) = f(y_row) for every column (TwoDGrid.updateY :47-58; `grid` only)."""
        v = _d([f(self.y.coord(r)) for r in range(self.rows)])
        L.check(L.lib().gs_grid2d_update_y(self._h, _p(v)))

    def _edge(self, side) -> float:
        out = C.c_double()
        L.check(L.lib().gs_grid2d_edge_average(self._h, side, C.byref(out)))
        return out.value

    leftAverage = property(lambda s: s._edge(LEFT))
    rightAverage = property(lambda s: s._edge(RIGHT))
    upperAverage = property(lambda s: s._edge(UPP))
    lowerAverage = property(lambda s: s._edge(LOW))

    @property
    def avValue(self) -> float:
        return self.grid.avValue

    def set_solver(self, solver: int):
        L.check(L.lib().gs_grid2d_set_solver(self._h, solver))

    @property
    def status(self) -> int:
        f = C.c_uint32()
        L.check(L.lib().gs_grid2d_status(self._h, C.byref(f)))
        return f.value

    def clear_status(self):
        L.check(L.lib().gs_grid2d_clear_status(self._h))

    @property
    def bytes_per_step(self) -> float:
        b = C.c_double()
        L.check(L.lib().gs_grid2d_bytes_per_step(self._h, C.byref(b)))
        return b.value


# ------------------------------------------------------------------------------------------
# Heat-flux helper formulas of object TwoDGrid (A/TwoDGrid.scala:1139-1181): pure host arithmetic
# on values a user loop reads through the edge accessors (AT/GridTest.scala:51-66).  Operation order as written there.
# ------------------------------------------------------------------------------------------
def getRadialMidPoint(r0, r1, r2, t0, t1, k0=None, k1=None) -> float:
    rAtHalf0 = (r0 + r1) / 2.0
    rAtHalf1 = (r1 + r2) / 2.0
    if k0 is None:
        assert r2 - r1 == r1 - r0
        return (t0 * rAtHalf0 + t1 * rAtHalf1) / (rAtHalf0 + rAtHalf1)
    a = rAtHalf0 * k0 * (r1 - r0)
    b = rAtHalf1 * k1 * (r2 - r1)
    return (a * t0 + b * t1) / (a + b)


def discHeatFlow(t0, t1, r0, r1, k) -> float:
    rAtHalf = (r0 + r1) / 2.0
    return k * rAtHalf * (t0 - t1) / (r1 - r0) * math.pi * 2.0


def radialHeatFlow(t0, t1, r0, r1, k) -> float:
    resist = math.log(r1 / r0) / (math.pi * 2.0 * k)
    return (t0 - t1) / resist


def discTWithFlow(heatFlow, t0, r0, r1, k) -> float:
    rAtHalf = (r0 + r1) / 2.0
    coef = k * rAtHalf / (r1 - r0) * math.pi * 2.0
    return t0 - heatFlow / coef


# ---- text output: TwoDGrid.write / toString (A/TwoDGrid.scala:62-133), XDim.toString (A/XDim.scala:17-23),
# OneDGrid.writeGrid (A/OneDGrid.scala:37-63).  The JVM's two number formatters are restated from their documented
# behaviour (no JVM exists in this image: tests/test_host_helpers.py pins known-answer strings).
def java_number_format(x: float, max_fraction_digits: int) -> str:
    """java.text.NumberFormat.getInstance(Locale.ROOT) with setMaximumFractionDigits(n): DecimalFormat "#,##0.###",
    RoundingMode.HALF_EVEN applied to the EXACT binary value (JDK >= 8), grouping by three, no trailing zeros, "-0" for
    a negative value that rounds to zero, "NaN" and the infinity sign for non-finite values."""
    import decimal

    x = float(x)
    if x != x:
        return "NaN"
    if x in (math.inf, -math.inf):
        return ("-" if x < 0 else "") + "∞"
    neg = math.copysign(1.0, x) < 0
    q = decimal.Decimal(abs(x)).quantize(decimal.Decimal(1).scaleb(-max_fraction_digits), rounding=decimal.ROUND_HALF_EVEN)
    digits = f"{q:f}"
    whole, _, frac = digits.partition(".")
    frac = frac.rstrip("0")
    groups = []
    while len(whole) > 3:
        groups.insert(0, whole[-3:])
        whole = whole[:-3]
    groups.insert(0, whole)
    return ("-" if neg else "") + ",".join(groups) + ("." + frac if frac else "")


def java_fixed(x: float, digits: int) -> str:
    """Scala's f"$x%.3f" = java.util.Formatter %.nf: RoundingMode.HALF_UP applied to the decimal digits of
    Double.toString (the shortest digits that round-trip, which is what repr() gives), always `digits` decimals."""
    import decimal

    x = float(x)
    if x != x:
        return "NaN"
    if x in (math.inf, -math.inf):
        return "-Infinity" if x < 0 else "Infinity"
    q = decimal.Decimal(repr(x)).quantize(decimal.Decimal(1).scaleb(-digits), rounding=decimal.ROUND_HALF_UP)
    s = f"{abs(q):f}"
    return ("-" if math.copysign(1.0, x) < 0 else "") + s


def _dim_to_string(d: "_Dim", sep: str = os.linesep) -> str:
    """XDim.toString (A/XDim.scala:17-23)"""
    return (f"X/Y\t\t|\t{java_fixed(d.low, 3)}\t|\t" + "\t".join(java_fixed(v, 3) for v in d.range)
            + f"\t|\t{java_fixed(d.upp, 3)}\t|\t..." + sep)


def twod_write(g: "TwoDGrid", writer, newline: str = os.linesep) -> None:
    """TwoDGrid.write(writer) (:62-82): per row `format(rowCoord) format(g(row, 0)) ... ` + BufferedWriter.newLine()"""
    a = g.grid.grid
    for r in range(g.rows):
        writer.write(" ".join([java_number_format(g.y.range[r], 3)] + [java_number_format(v, 3) for v in a[r]]))
        writer.write(newline)


def twod_to_string(g: "TwoDGrid", sep: str = os.linesep) -> str:
    """TwoDGrid.toString (:86-133), quirks included: the lower-bound line comes first and ends with y.upp, the right
    bound is printed with four decimals, the upper-bound line starts with y.upp."""
    f3 = lambda v: java_fixed(v, 3)
    xc = _dim_to_string(g.x, sep)
    low, upp, left, right = g.bounds.low.get(), g.bounds.upp.get(), g.bounds.left.get(), g.bounds.right.get()
    out = [xc]
    out.append(f"{f3(g.y.low)}\t|\t...\t\t|\t" + "\t".join(f3(low[i]) for i in range(g.cols))
               + f"\t|\t...\t\t|\t{f3(g.y.upp)}" + sep)
    a = g.grid.grid
    for r in range(g.rows):
        yc = g.y.range[r]
        out.append(f"{f3(yc)}\t|\t{f3(left[r])}\t|\t" + "\t".join(f3(v) for v in a[r])
                   + f"\t|\t{java_fixed(right[r], 4)}\t|\t{f3(yc)}" + sep)
    out.append(f"{f3(g.y.upp)}\t|\t...\t\t|\t" + "\t".join(f3(upp[i]) for i in range(g.cols))
               + f"\t|\t...\t\t|\t{f3(g.y.upp)}" + sep)
    out.append(xc)
    return "".join(out)


def oned_write_grid(g: "OneDGrid", at, path: str, newline: str = os.linesep) -> str:
    """OneDGrid.writeGrid(at: LocalDateTime, path) (:37-63): `grid MONTH Dd Hh YEAR.dat`, a two-line header and one
    `coord<TAB>temperature` line per node, every element followed by "\\r\\n" AND BufferedWriter.newLine() (sic).
    `at` is a datetime; the file must already exist (StandardOpenOption.TRUNCATE_EXISTING without CREATE) in the
    reference -- here it is created.  Returns the file name."""
    os.makedirs(path, exist_ok=True)
    name = os.path.join(path, f"grid {at.strftime('%B').upper()} {at.day}d {at.hour}h {at.year}.dat")
    temps = g.grid
    with open(name, "w", newline="") as f:
        f.write("# One dimension grid output results\r\n# Coords \t Temperatures\r\n" + newline)
        for x, t in zip(g.rangeX, temps):
            f.write(f"{java_number_format(x, 2)}\t{java_number_format(t, 3)}\r\n" + newline)
    return name


TwoDGrid.write = twod_write
TwoDGrid.toString = twod_to_string
TwoDGrid.__str__ = lambda self: twod_to_string(self)
_Dim.toString = _dim_to_string
OneDGrid.writeGrid = oned_write_grid


class ShardedTwoDGrid:
    """ONE TwoDGrid on several GPUs of a box, driven by this process (gs_sharded2d_*, csrc/gs_sharded2d.cu): rows are
    sharded over `devices`, the step's exchange happens inside the kernels over NVLink peer memory.  Same constructor
    shape as TwoDGrid; `coefs_factory(register)` receives a function that registers a Spline on every device and
    returns its id, and returns {selector: (mode, ids)} for the whole grid."""

    def __init__(self, x: XDim, y: YDim, devices: Sequence[int], upp=(One, Temp), low=(One, Temp), right=(One, Temp),
                 left=(One, Temp)):
        self.x, self.y = x, y
        self.cols, self.rows = x.colsNum, y.rowsNum
        dev = np.ascontiguousarray(list(devices), dtype=np.int32)
        xv, yv = x.values, y.values
        self._h = C.c_void_p()
        L.check(L.lib().gs_sharded2d_create(len(dev), dev.ctypes.data_as(_ip), self.cols, self.rows, x.t, y.t, _p(xv),
                                            _p(yv), C.byref(self._h)))
        self._kinds = {UPP: upp, LOW: low, RIGHT: right, LEFT: left}
        for side in (UPP, LOW, RIGHT, LEFT):
            self.set_bound(side, 0.0)

    def close(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            L.lib().gs_sharded2d_destroy(h)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def register(self, spl: Spline, true_upper: bool = False) -> int:
        """AlwaysDefinedSpline(spl) in every device's pool (clamps by the reference's literal rule, evaluated on the host
        mirror of the piece formulas for Const tables, else through device 0 of the default context)."""
        a = AlwaysDefinedSpline(spl, true_upper=true_upper)       # computes the clamps (default context)
        out = C.c_int()
        L.check(L.lib().gs_sharded2d_spline_create(self._h, spl.kind, spl.size, _p(spl.knots), _p(spl.x0),
                                                   _p(np.ascontiguousarray(spl.coefs)), *a.clamps, C.byref(out)))
        return out.value

    def set_constant_coef(self, therm: Spline, cap: Spline, tempcond: Spline):
        t, c, a = self.register(therm), self.register(cap), self.register(tempcond)
        for sel, sid in ((SEL_APPLY, a), (SEL_THERM, t), (SEL_CAP, c), (SEL_TEMPCOND, a)):
            self.set_map(sel, MAP_SINGLE, [sid])

    def set_map(self, selector: int, mode: int, ids):
        ids = np.ascontiguousarray(ids, dtype=np.int32)
        L.check(L.lib().gs_sharded2d_set_map(self._h, selector, mode, ids.ctypes.data_as(_ip)))

    def set_bound(self, side: int, values):
        rep, kind = self._kinds[side]
        v = np.atleast_1d(_d(values))
        L.check(L.lib().gs_sharded2d_set_bound(self._h, side, kind, rep, _p(v), len(v)))

    def fill(self, v: float):
        L.check(L.lib().gs_sharded2d_fill(self._h, v))

    def upload(self, which: int, values):
        v = _d(values)
        assert v.size == self.rows * self.cols
        L.check(L.lib().gs_sharded2d_upload(self._h, which, v.ctypes.data))

    def download(self, which: int = GRID) -> np.ndarray:
        out = np.empty((self.rows, self.cols))
        L.check(L.lib().gs_sharded2d_download(self._h, which, out.ctypes.data))
        return out

    def upload_ptr(self, which: int, host_ptr: int):
        L.check(L.lib().gs_sharded2d_upload(self._h, which, host_ptr))

    def download_ptr(self, which: int, host_ptr: int):
        L.check(L.lib().gs_sharded2d_download(self._h, which, host_ptr))

    def iteration(self, time: float, nsteps: int = 1):
        L.check(L.lib().gs_sharded2d_iteration(self._h, time, nsteps))

    def iteration_timed(self, time: float, nsteps: int = 1) -> float:
        """nsteps iterations; returns the device-measured time in ms (max over the devices)."""
        ms = C.c_double()
        L.check(L.lib().gs_sharded2d_iteration_timed(self._h, time, nsteps, C.byref(ms)))
        return ms.value

    def synchronize(self):
        L.check(L.lib().gs_sharded2d_synchronize(self._h))

    def set_option(self, name: str, value: int):
        L.check(L.lib().gs_sharded2d_set_option(self._h, name.encode(), value))

    @property
    def status(self) -> int:
        f = C.c_uint32()
        L.check(L.lib().gs_sharded2d_status(self._h, C.byref(f)))
        return f.value

    @property
    def launch_count(self) -> int:
        c = C.c_uint64()
        L.check(L.lib().gs_sharded2d_launch_count(self._h, C.byref(c)))
        return c.value
