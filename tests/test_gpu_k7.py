"""K7 -- the fast bit-exact 2-D sweeps (faces pass + one thread per line fed by a cp.async.bulk ring,
csrc/gs_grid2d_k7.cuh) -- against the oracle, bit for bit: every boundary kind, ragged sizes (partial strips, partial
chunks), every coefficient-map mode and table kind, the native-division fallback, and the first-draft kernels
(k_xsweep / k_ysweep, option k7 = 0) as a second witness.  K7 needs an even number of columns (16-byte bulk copies)."""
import itertools

import numpy as np
import pytest

import cases
from helpers import assert_bit_equal

pytestmark = pytest.mark.gpu


@pytest.fixture(params=[2, 0], ids=["fused-faces", "separate-faces"])
def k7(ctx, request):
    """Both variants: k7f_sweep (faces by a producer warp inside the sweep) and k7_sweep after a k6_faces pass."""
    ctx.set_option("k7", 1)
    ctx.set_option("k7_fuse", request.param)
    ctx.k7_fuse = request.param
    yield ctx
    ctx.set_option("k7", 1)
    ctx.set_option("k7_fuse", 2)


def _snap_equal(g, o, what):
    got = cases.snapshot(g)
    for k, v in cases.snapshot(o).items():
        assert_bit_equal(got[k], v, f"{what} {k}")


@pytest.mark.parametrize("combo", ["".join(c) for c in itertools.product("TF", repeat=4)])
@pytest.mark.parametrize("dense", [True, False])
def test_k7_all_boundary_kinds(O, gs, k7, combo, dense):
    xdir = 1 if combo.count("F") % 2 else 0
    p = cases.bc_problem(combo, xdir=xdir, ydir=0, cols=40, rows=36, dense=dense)
    g, o = p.build_gs(gs), p.build_oracle(O)
    g.set_solver(gs.SOLVER_THOMAS)
    n0 = k7.launch_count
    g.xIter(900.0); o.xIter(900.0)
    assert k7.launch_count - n0 == (2 if k7.k7_fuse else 3), "K7 = k6_ends [+ k6_faces] + sweep"
    assert_bit_equal(g.grid.result, o.result, combo + " xIter result")
    g.yIter(900.0); o.yIter(900.0)
    assert_bit_equal(g.grid.result, o.result, combo + " yIter result")
    g.grid.update(); o.update()
    g.iteration(900.0, nsteps=2)
    o.iteration(900.0); o.iteration(900.0)
    _snap_equal(g, o, f"{combo} dense={dense}")
    assert g.status == 0


@pytest.mark.parametrize("cols,rows", [(4, 4), (4, 9), (16, 16), (18, 5), (32, 32), (34, 31), (64, 65), (130, 97),
                                       (258, 190), (48, 17), (16, 33)])
@pytest.mark.parametrize("combo", ["TTTT", "FFFF", "TFTF"])
def test_k7_ragged_sizes(O, gs, k7, cols, rows, combo):
    p = cases.bc_problem(combo, xdir=1, ydir=1, cols=cols, rows=rows)
    g, o = p.build_gs(gs), p.build_oracle(O)
    g.set_solver(gs.SOLVER_THOMAS)
    g.iteration(900.0, nsteps=2)
    o.iteration(900.0); o.iteration(900.0)
    _snap_equal(g, o, f"{cols}x{rows} {combo}")


@pytest.mark.parametrize("coefs", ["hermite", "per_row", "per_col", "dense", "lines_uneven", "hermite_uneven"])
@pytest.mark.parametrize("combo,cols,rows", [("TTTT", 38, 29), ("TTFT", 38, 29), ("FTTF", 130, 70)])
def test_k7_tables_and_maps(O, gs, k7, coefs, combo, cols, rows):
    p = cases.bc_problem(combo, xdir=1, cols=cols, rows=rows, coefs=coefs)
    r = np.random.default_rng(4)
    p.grid0 = r.uniform(-15.0, 70.0, p.grid0.shape)     # several table pieces: the interval average crosses knots
    p.predict0 = p.grid0 + r.uniform(-12.0, 12.0, p.grid0.shape)
    g, o = p.build_gs(gs), p.build_oracle(O, ascending=True)
    g.set_solver(gs.SOLVER_THOMAS)
    g.iteration(900.0, nsteps=3)
    for _ in range(3):
        o.iteration(900.0)
    _snap_equal(g, o, f"{coefs} {combo}")


def test_k7_native_division_fallback(O, gs, k7):
    """Exact zeros (dividend outside the shared-reciprocal divider's range) send a warp through the native-division
    walk: same bits."""
    p = cases.bc_problem("TFFT", xdir=1, cols=96, rows=80, coefs="hermite")
    p.grid0[5, :] = 0.0
    p.grid0[:, 7] = 0.0
    p.grid0[40:44, 60:66] = 0.0
    p.predict0 = p.grid0.copy()
    g, o = p.build_gs(gs), p.build_oracle(O)
    g.set_solver(gs.SOLVER_THOMAS)
    g.iteration(900.0, nsteps=2)
    o.iteration(900.0); o.iteration(900.0)
    _snap_equal(g, o, "zeros")


def test_k7_nan_is_flagged_like_the_first_draft_kernels(gs, k7):
    p = cases.bc_problem("TTTT", cols=64, rows=48, coefs="hermite")
    p.grid0[10, 20] = np.nan
    p.predict0 = p.grid0.copy()
    out = {}
    for opt in (1, 0):
        k7.set_option("k7", opt)
        g = p.build_gs(gs)
        g.set_solver(gs.SOLVER_THOMAS)
        g.iteration(900.0)
        out[opt] = (cases.snapshot(g), g.status)
        g.close()
    assert out[1][1] == out[0][1] and out[1][1] & 4      # same sticky flags; GS_FLAG_NONFINITE among them
    for k in ("grid", "predict"):
        assert np.array_equal(np.isnan(out[1][0][k]), np.isnan(out[0][0][k])), k
        m = ~np.isnan(out[0][0][k])
        assert_bit_equal(out[1][0][k][m], out[0][0][k][m], "NaN case, finite part of " + k)


@pytest.mark.parametrize("coefs,combo", [("hermite", "TTTT"), ("const", "TTFF"), ("dense", "FTTT")])
def test_k7_medium_vs_oracle_and_first_draft(O, gs, k7, coefs, combo):
    cols, rows = 1056, 650          # 33 x strips + partial ones, 66 chunks + a partial one
    p = cases.bc_problem(combo, xdir=1, cols=cols, rows=rows, coefs=coefs)
    r = np.random.default_rng(11)
    p.grid0 = r.uniform(-10.0, 60.0, p.grid0.shape)
    p.predict0 = p.grid0 + r.uniform(-3.0, 3.0, p.grid0.shape)
    o = p.build_oracle(O)
    for _ in range(2):
        o.iteration(900.0)
    for opt in (1, 0):
        k7.set_option("k7", opt)
        g = p.build_gs(gs)
        g.set_solver(gs.SOLVER_THOMAS)
        g.iteration(900.0, nsteps=2)
        _snap_equal(g, o, f"{coefs} k7={opt}")
        assert g.status == 0
        g.close()


def test_k7_c4_shape_4096_vs_oracle(O, gs, k7):
    """The C4 problem (radial x, M1Hermite3 k(T), c(T), alpha = k / c, Dirichlet 30 / 7) at 4096^2, two steps, every node
    compared with the oracle bit for bit."""
    n = 4096
    r = np.random.default_rng(2029)
    field = 7.0 + r.uniform(size=(n, n))
    xv = np.concatenate([[0.05], 0.05 * 1.0005 ** np.arange(1, n + 1), [0.05 * 1.0005 ** (n + 1)]])
    yv = np.concatenate([[0.0], 0.01 * np.arange(1, n + 1), [0.01 * (n + 1)]])
    b = {0: (0, 1, np.full(n, 7.0)), 1: (0, 1, np.full(n, 7.0)), 2: (0, 1, np.full(n, 7.0)), 3: (0, 1, np.full(n, 30.0))}
    p = cases.Problem2D(xv, yv, 1, 0, b, coefs="hermite", init=(field, field.copy()))
    g, o = p.build_gs(gs), p.build_oracle(O)
    g.set_solver(gs.SOLVER_THOMAS)
    g.iteration(900.0, nsteps=2)
    o.iteration(900.0); o.iteration(900.0)
    _snap_equal(g, o, "C4 shape 4096^2")
    assert g.status == 0


@pytest.mark.parametrize("coefs", ["hermite", "const"])
@pytest.mark.parametrize("value", [0.0, 7.0, -0.0])
def test_k7_uniform_and_zero_fields_stay_exact(O, gs, coefs, value):
    """A field that starts uniform -- and one of exactly 0.0 -- is ordinary input (AT/GridTest.scala fills 7.0): equal
    neighbours make average = apply, temperatures a few ulps apart make the antiderivative difference cancel to 0.0 (a
    face value of 0.0), zeros make -t / tau = -0.0.  All of it must stay on the straight-line tiers' bits = the oracle's
    (signs of zero included), for several steps while the boundary layer grows into the field."""
    p = cases.bc_problem("TTTT", xdir=1, ydir=0, cols=96, rows=70, coefs=coefs)
    p.grid0 = np.full(p.grid0.shape, value)
    p.predict0 = p.grid0.copy()
    g, o = p.build_gs(gs), p.build_oracle(O)
    g.set_solver(gs.SOLVER_THOMAS)
    for step in range(4):
        g.iteration(900.0)
        o.iteration(900.0)
        got, want = cases.snapshot(g), cases.snapshot(o)
        for k in ("grid", "predict", "result"):
            assert_bit_equal(got[k], want[k], f"{coefs} from {value!r}, step {step}: {k}")
    assert g.status == 0


@pytest.mark.parametrize("value,upp,low", [(76.0, 79.0, 72.0), (-24.0, -22.0, -30.0), (69.0, 75.0, 71.0)])
def test_k7_fields_beyond_the_table_clamps_stay_exact(O, gs, value, upp, low):
    """Temperatures at or beyond the same clamp of the table (hotter than upperX = the second-to-last knot under the
    reference's greatestK rule, or colder than lowerX) take the clamped face tier: ((hi - lo) * Y) / (hi - lo) as
    AlwaysDefinedSpline.area / average compute it; a field that straddles the clamp (third case) mixes every tier."""
    p = cases.bc_problem("TTTT", xdir=1, ydir=0, cols=96, rows=70, coefs="hermite")
    r = np.random.default_rng(5)
    p.grid0 = value + r.uniform(-0.5, 0.5, p.grid0.shape)
    p.grid0[:20, :30] = value          # a uniform patch beyond the clamp: apply = Y
    p.predict0 = p.grid0.copy()
    g, o = p.build_gs(gs), p.build_oracle(O)
    for side, v in ((gs.UPP, upp), (gs.LOW, low), (gs.LEFT, upp), (gs.RIGHT, low)):
        getattr(g.bounds, {gs.UPP: "upp", gs.LOW: "low", gs.LEFT: "left", gs.RIGHT: "right"}[side]).update(v)
        o.set_bound(side, O.TEMP, O.DENSE, np.full(p.cols if side in (gs.UPP, gs.LOW) else p.rows, v))
    g.set_solver(gs.SOLVER_THOMAS)
    for step in range(3):
        g.iteration(900.0)
        o.iteration(900.0)
        got, want = cases.snapshot(g), cases.snapshot(o)
        for k in ("grid", "predict", "result"):
            assert_bit_equal(got[k], want[k], f"around {value}, step {step}: {k}")
    assert g.status == 0
