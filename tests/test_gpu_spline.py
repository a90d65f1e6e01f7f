"""GPU parity: K1 metric coefficients and device spline evaluation against the oracle (bit-exact)."""
import numpy as np
import pytest

from helpers import README_RHO, README_T, assert_bit_equal, max_rel_err

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("direction", [0, 1])
@pytest.mark.parametrize("n", [2, 3, 200, 4097])
def test_predef_and_heatflow_geometry(O, gs, ctx, direction, n):
    r = np.random.default_rng(n)
    rng = 0.05 + np.cumsum(r.uniform(1e-3, 2e-2, n))
    low, upp = 0.05, rng[-1] + 0.013
    for sigma in (1.0, 0.37):
        assert_bit_equal(ctx.predef(direction, low, rng, upp, sigma), O.predef(direction, low, rng, upp, sigma), "preDef")
    assert_bit_equal(ctx.heatflow_geometry(direction, low, rng, upp), O.heatflow_geometry(direction, low, rng, upp), "heatflow")


def _pairs(O, gs):
    pts = list(zip(README_T, README_RHO))
    lin = [(0.0, 1.0), (10.0, 3.0), (25.0, 2.0), (40.0, 7.5)]
    return [
        (O.Spline.const(1.5), gs.Spline.const(1.5)),
        (O.Spline.const(2.0, 0.0, 50.0), gs.Spline.const(2.0, 0.0, 50.0)),
        (O.Spline.lines(lin), gs.Spline.lines(lin)),
        (O.Spline.m1Hermite3(pts), gs.Spline.m1Hermite3(pts)),
        (O.Spline.fHermite3(pts), gs.Spline.fHermite3(pts)),
    ]


@pytest.mark.parametrize("true_upper", [False, True])
def test_always_defined_spline_apply_and_average(O, gs, ctx, true_upper):
    r = np.random.default_rng(11)
    for os_, gsp in _pairs(O, gs):
        oa = O.AlwaysDefinedSpline(os_, true_upper=true_upper)
        ga = gs.AlwaysDefinedSpline(gsp, ctx, true_upper=true_upper)
        assert oa.clamps == ga.clamps
        lo_k, hi_k = max(gsp.knots[0], -50.0), min(gsp.knots[-1], 250.0)
        xs = np.concatenate([r.uniform(lo_k - 30.0, hi_k + 30.0, 4000), gsp.knots[np.isfinite(gsp.knots)],
                             [lo_k, hi_k, 0.0, -0.0]])
        xs = xs[np.abs(xs) < 1e300]
        assert_bit_equal(ga(xs), np.array([oa(x) for x in xs]), "apply")
        a = r.uniform(lo_k - 30.0, hi_k + 30.0, 4000)
        b = np.where(r.uniform(size=4000) < 0.5, a + r.uniform(0.0, 3.0, 4000), r.uniform(lo_k - 30.0, hi_k + 30.0, 4000))
        b[:50] = a[:50]  # the t_lo == t_hi branch
        # ordered (takeAverage) and raw order (Neumann calls, A/Dim.scala:55,116)
        O.set_ascending_fold(True)
        want = np.array([oa.average(x, y) for x, y in zip(a, b)])
        assert_bit_equal(ga.average(a, b), want, "average (ascending fold)")
        O.set_ascending_fold(False)
        lit = np.array([oa.average(x, y) for x, y in zip(a, b)])
        ok = np.isfinite(lit) & (lit != 0.0)
        assert max_rel_err(ga.average(a, b)[ok], lit[ok]) <= 1e-12


def test_spline_undefined_is_flagged(gs, ctx):
    g = gs.OneDGrid(0.0, [0.1, 0.2, 0.3], 0.4, gs.Spline.const(1e-6), 1.0)
    g.grid = [1.0, float("nan"), 2.0]
    g.step(900.0, 1.0, 2.0)
    assert g.status & gs.FLAG_NONFINITE
    # a NaN temperature reaches Spline.apply: RuntimeException in the reference (AbsITree.scala:475-479)
    lin = gs.Spline.lines([(-50.0, 1e-6), (50.0, 2e-6), (150.0, 3e-6)])
    g2 = gs.OneDGrid(0.0, [0.1, 0.2, 0.3], 0.4, lin, 1.0)
    g2.upload(gs.PREDICT, np.array([[1.0, float("nan"), 2.0]]))
    g2.step(900.0, 1.0, 2.0)
    assert g2.status & gs.FLAG_SPLINE_UNDEFINED
    # same with a constant spline on the general path; the hoisted path (default for constant properties) never
    # reads predict, so it cannot see that NaN (documented in DESIGN.md section 4)
    ctx.set_option("const_hoist", 0)
    try:
        g3 = gs.OneDGrid(0.0, [0.1, 0.2, 0.3], 0.4, gs.Spline.const(1e-6), 1.0)
        g3.upload(gs.PREDICT, np.array([[1.0, float("nan"), 2.0]]))
        g3.step(900.0, 1.0, 2.0)
        assert g3.status & gs.FLAG_SPLINE_UNDEFINED
    finally:
        ctx.set_option("const_hoist", 1)


def _oracle_twin(O, s):
    """The same table for the oracle, piece by piece (kind, lo, hi, x0, coefs)."""
    np_ = len(s.x0)
    return O.Spline.from_pieces([s.kind] * np_, s.knots[:-1], s.knots[1:], s.x0, s.coefs)


@pytest.mark.parametrize("factory", ["lagrange2", "lagrange3"])
def test_lagrange_tables(O, gs, factory):
    """Lagrange tables (P/Lagrange2.scala, P/Lagrange3.scala) reach the device as cubic pieces (x0 = 0, c3 = 0 for the
    parabolas; x0 = low for the cubics): point values, interval averages and a batched 1-D step with such a table as
    the conductivity are the oracle's, bit for bit."""
    pts = [(-20.0 + 10.0 * i, 1e-6 * (1.0 + 0.1 * README_RHO[i])) for i in range(11)]
    s = getattr(gs.Spline, factory)(pts)
    if factory == "lagrange2":       # every parabola interpolates its three points
        for k in range(len(s.x0)):
            c = s.coefs[k]
            for x, y in pts[k:k + 3]:
                assert abs((c[2] * x + c[1]) * x + c[0] - y) <= 1e-12 * abs(y) * 100
    o = _oracle_twin(O, s)
    ga, oa = gs.AlwaysDefinedSpline(s), O.AlwaysDefinedSpline(o)
    r = np.random.default_rng(5)
    hi_knot = float(s.knots[-1])
    xs = np.concatenate([r.uniform(-30.0, hi_knot + 10.0, 400), s.knots, [-20.0, hi_knot]])
    assert_bit_equal(ga(xs), np.array([oa(float(x)) for x in xs]), factory + " apply")
    a, b = r.uniform(-25.0, hi_knot + 5.0, 300), r.uniform(-25.0, hi_knot + 5.0, 300)
    lo, hi = np.minimum(a, b), np.maximum(a, b)
    O.set_ascending_fold(True)
    assert_bit_equal(ga.average(lo, hi), np.array([oa.average(float(p), float(q)) for p, q in zip(lo, hi)]),
                     factory + " average")
    # a batched 1-D step with the table as conductivity
    n, nlines = 64, 40
    rangeX = np.array([0.01 * (i + 1) for i in range(n)])
    grid0 = r.uniform(-10.0, 60.0, (nlines, n))
    lt, rt = r.uniform(10.0, 30.0, nlines), r.uniform(0.0, 10.0, nlines)
    g = gs.BatchedOneDGrid(0.0, rangeX, 0.01 * (n + 1), s, 1.0, nlines=nlines)
    g.upload(gs.GRID, grid0)
    g.step(900.0, lt, rt, nsteps=3)
    og, op = grid0.copy(), np.full((nlines, n), 4.0)
    O.batch1d_step(O.ORTHO, 0.0, rangeX, 0.01 * (n + 1), o, 1.0, og, op, lt, rt, 900.0, nsteps=3, nthreads=2)
    assert_bit_equal(g.download(gs.GRID), og, factory + " 1-D grid")
    assert_bit_equal(g.download(gs.PREDICT), op, factory + " 1-D predict")
