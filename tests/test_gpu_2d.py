"""GPU parity of the 2-D stepper (K3 x-sweep, K4 y-sweep, Grid.update) against the oracle and the committed
fixtures: all four boundary kinds, both direction types, every coefficient-map mode.  GS_SOLVER_THOMAS keeps
the reference's operation order: bit-exact (against the ascending-fold oracle for multi-piece tables;
<= 1e-12 against the literal tree-order fold)."""
import itertools
import os

import numpy as np
import pytest

import cases
from helpers import assert_bit_equal, check_tol, elem_rel_err, max_rel_err

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def _small(coefs="unit"):
    xv, yv = cases.small2d_dims()
    b = {s: (0, 0, [1.0]) for s in range(4)}
    return cases.Problem2D(xv, yv, 1, 0, b, coefs=coefs, init=(np.ones((10, 8)), np.ones((10, 8))))


def test_kat_r3_r4_r5(O, gs):
    p = _small()
    p.bounds[3] = (0, 0, [4.0])
    g, o = p.build_gs(gs), p.build_oracle(O)
    g.xIter(900.0); o.xIter(900.0)
    g.grid.update(); o.update()
    assert list(g.grid.grid[0]) == [3.10981041911049, 2.5775720576159933, 2.1989042489843205, 1.9055689452327573,
                                    1.6664820491786831, 1.464861323850035, 1.2906054403084588, 1.137154150062293]
    assert g.leftAverage == o.edgeAverage(O.LEFT) and g.rightAverage == o.edgeAverage(O.RIGHT)
    assert g.leftAverage > g.rightAverage  # AT/TwoDGridUnit.scala:63-73
    # R4: yIter alone consumes whatever `result` holds (zeros)
    p = _small()
    p.bounds[0] = (0, 0, [4.0])
    g, o = p.build_gs(gs), p.build_oracle(O)
    g.yIter(900.0); o.yIter(900.0)
    g.grid.update(); o.update()
    for k, v in cases.snapshot(o).items():
        assert_bit_equal(cases.snapshot(g)[k], v, "R4 " + k)
    assert g.upperAverage == o.edgeAverage(O.UPP) and g.lowerAverage == o.edgeAverage(O.LOW)
    assert g.upperAverage > g.lowerAverage  # :75-85
    # R5
    p = _small()
    p.bounds[3] = (0, 0, [4.0])
    g = p.build_gs(gs)
    g.iteration(900.0)
    assert g.grid.grid[0, 0] == 1.0115798259742925
    assert g.avValue == pytest.approx(1.0110804508847722, rel=1e-13) and g.avValue > 1.0  # :87-94


@pytest.mark.parametrize("name,combo,xdir,coefs", [
    ("grid2d_TTFF_ortho_const", "TTFF", 0, "const"),
    ("grid2d_FFTT_radial_const", "FFTT", 1, "const"),
    ("grid2d_TTTT_radial_hermite", "TTTT", 1, "hermite"),
])
def test_vs_golden(gs, name, combo, xdir, coefs):
    p = cases.bc_problem(combo, xdir=xdir, coefs=coefs)
    g = p.build_gs(gs)
    g.iteration(900.0, nsteps=3)
    want = np.load(os.path.join(GOLDEN, name + ".npz"))
    got = cases.snapshot(g)
    for k in want.files:
        assert_bit_equal(got[k], want[k], f"{name}.{k}")
    assert g.status == 0


@pytest.mark.parametrize("combo", ["".join(c) for c in itertools.product("TF", repeat=4)])
@pytest.mark.parametrize("dense", [True, False])
def test_all_boundary_kinds(O, gs, combo, dense):
    xdir = 1 if combo.count("F") % 2 else 0
    p = cases.bc_problem(combo, xdir=xdir, ydir=0, cols=41, rows=35, dense=dense)
    g, o = p.build_gs(gs), p.build_oracle(O)
    # separate calls first (xIter, yIter, update), then fused iterations
    g.xIter(900.0); o.xIter(900.0)
    assert_bit_equal(g.grid.result, o.result, combo + " xIter result")
    g.yIter(900.0); o.yIter(900.0)
    assert_bit_equal(g.grid.result, o.result, combo + " yIter result")
    g.grid.update(); o.update()
    g.iteration(900.0, nsteps=2)
    o.iteration(900.0); o.iteration(900.0)
    for k, v in cases.snapshot(o).items():
        assert_bit_equal(cases.snapshot(g)[k], v, f"{combo} dense={dense} {k}")


@pytest.mark.parametrize("cols,rows", [(2, 2), (2, 9), (9, 2), (3, 3), (32, 32), (33, 31), (64, 65), (130, 97)])
@pytest.mark.parametrize("combo", ["TTTT", "FFFF", "TFTF"])
def test_ragged_sizes(O, gs, cols, rows, combo):
    p = cases.bc_problem(combo, xdir=1, ydir=1, cols=cols, rows=rows)
    g, o = p.build_gs(gs), p.build_oracle(O)
    g.iteration(900.0, nsteps=2)
    o.iteration(900.0); o.iteration(900.0)
    for k, v in cases.snapshot(o).items():
        assert_bit_equal(cases.snapshot(g)[k], v, f"{cols}x{rows} {combo} {k}")


@pytest.mark.parametrize("coefs", ["hermite", "per_row", "per_col", "dense"])
@pytest.mark.parametrize("combo", ["TTTT", "TTFT"])
def test_coefficient_maps_and_tables(O, gs, coefs, combo):
    p = cases.bc_problem(combo, xdir=1, cols=37, rows=29, coefs=coefs)
    # spread the temperatures over several table pieces so the interval average crosses knots
    r = np.random.default_rng(4)
    p.grid0 = r.uniform(-15.0, 70.0, p.grid0.shape)
    p.predict0 = p.grid0 + r.uniform(-12.0, 12.0, p.grid0.shape)
    g, o = p.build_gs(gs), p.build_oracle(O, ascending=True)
    g.iteration(900.0, nsteps=3)
    for _ in range(3):
        o.iteration(900.0)
    got = cases.snapshot(g)
    for k, v in cases.snapshot(o).items():
        assert_bit_equal(got[k], v, f"{coefs} {combo} {k}")
    lit = p.build_oracle(O, ascending=False)
    for _ in range(3):
        lit.iteration(900.0)
    O.set_ascending_fold(True)
    for k, v in cases.snapshot(lit).items():
        check_tol(got[k], v, f"exact solver vs literal tree-order fold {coefs} {combo} {k}", 1e-12)


def test_gridtest_physical_validation(O, gs):
    """AT/GridTest.scala: 137 x 45 radial x ortho, heated Neumann wall, 195 steps; bit-exact vs golden and
    the wall temperature of KAT R1."""
    p = cases.gridtest_problem()
    g = p.build_gs(gs)
    g.fill(7.0)
    got = cases.run_gridtest(g, True)
    want = np.load(os.path.join(GOLDEN, "gridtest_137x45x195.npz"))
    for k in want.files:
        assert_bit_equal(got[k], want[k], "GridTest " + k)
    assert got["grid"][0, 0] == 54.44328715390563


def test_no_heat_flow_and_bounds(O, gs):
    p = cases.bc_problem("TFTF", cols=13, rows=11, dense=True)
    g, o = p.build_gs(gs), p.build_oracle(O)
    for side in range(4):
        g.noHeatFlow(side); o.noHeatFlow(side)
    for side, b in zip((0, 1, 2, 3), (g.bounds.upp, g.bounds.low, g.bounds.right, g.bounds.left)):
        assert_bit_equal(b.get(), o.get_bound(side), f"noHeatFlow side {side}")  # AT/TwoDGridUnit.scala:104-130
    p = cases.bc_problem("TFTF", cols=13, rows=11, dense=False)
    g, o = p.build_gs(gs), p.build_oracle(O)
    for side in range(4):
        g.noHeatFlow(side); o.noHeatFlow(side)
    for side, b in zip((0, 1, 2, 3), (g.bounds.upp, g.bounds.low, g.bounds.right, g.bounds.left)):
        assert_bit_equal(b.get(), o.get_bound(side), f"sparse noHeatFlow side {side}")
    # Sparse.update(vals) stores the mean
    vals = np.linspace(1.0, 2.0, 13)
    g.bounds.upp.update(vals); o.set_bound(0, 0, 0, vals)
    assert_bit_equal(g.bounds.upp.get(), o.get_bound(0))


def test_medium_grid_properties(O, gs):
    """512 x 384 against the oracle (bit-exact), then a 2048 x 2048 run checked through properties: a field that
    is uniform along x with x-independent BCs stays uniform along x (every column solve is identical)."""
    p = cases.bc_problem("TTFF", cols=512, rows=384)
    g, o = p.build_gs(gs), p.build_oracle(O)
    g.set_solver(gs.SOLVER_THOMAS)
    g.iteration(900.0, nsteps=2)
    o.iteration(900.0); o.iteration(900.0)
    for k, v in cases.snapshot(o).items():
        assert_bit_equal(cases.snapshot(g)[k], v, "512x384 " + k)
    # 2048 x 2048 through properties: lines are independent, so a field that is uniform across the lines
    # gives bit-identical lines, and each equals the oracle's solve of a 2-line grid with the same data.
    n = 2048
    xv, yv = 0.05 + 0.01 * np.arange(n + 2), 0.01 * np.arange(n + 2)
    r = np.random.default_rng(2)
    line_p, line_g = 7.0 + r.uniform(size=n), 7.0 + r.uniform(size=n)
    bx = {0: (0, 1, np.full(n, 20.0)), 1: (0, 1, np.full(n, 4.0)), 2: (1, 0, [3.0]), 3: (0, 0, [9.0])}
    for sweep in ("x", "y"):
        tile = (lambda v: np.tile(v[None, :], (n, 1))) if sweep == "x" else (lambda v: np.tile(v[:, None], (1, n)))
        big = cases.Problem2D(xv, yv, 1, 0, bx, coefs="const", init=(tile(line_g), tile(line_p)))
        g = big.build_gs(gs)
        g.set_solver(gs.SOLVER_THOMAS)
        if sweep == "x":
            small = cases.Problem2D(xv, yv[:4], 1, 0, {k: (v[0], v[1], v[2][:2] if len(v[2]) > 1 else v[2]) for k, v in bx.items()},
                                    coefs="const", init=(tile(line_g)[:2], tile(line_p)[:2]))
            o = small.build_oracle(O)
            g.xIter(900.0); o.xIter(900.0)
            res = g.grid.result
            assert np.all(res.view(np.uint64) == res[:1].view(np.uint64))
            assert_bit_equal(res[0], o.result[0], "2048^2 xIter row")
        else:
            by = {0: (0, 0, [20.0]), 1: (0, 0, [4.0]), 2: (0, 0, [1.0]), 3: (0, 0, [1.0])}
            big.bounds = by
            g = big.build_gs(gs)
            g.set_solver(gs.SOLVER_THOMAS)
            g.grid.result = tile(line_g)
            small = cases.Problem2D(xv[:4], yv, 1, 0, by, coefs="const", init=(tile(line_g)[:, :2], tile(line_p)[:, :2]))
            o = small.build_oracle(O)
            o.result[:] = tile(line_g)[:, :2]
            g.yIter(900.0); o.yIter(900.0)
            res = g.grid.result
            assert np.all(res.view(np.uint64) == res[:, :1].view(np.uint64))
            assert_bit_equal(res[:, 0], o.result[:, 0], "2048^2 yIter column")
        assert g.status == 0


# ---------------------------------------------------------------------------------------------
# GS_SOLVER_PARTITIONED: segments + reduced system.  Operation order differs from the reference's single
# chain, so the bar is the north star's tolerance: <= 1e-12 max relative error per field.
# ---------------------------------------------------------------------------------------------
TOL = 1e-12
# Spline TABLES under a reordered solver: the reference's face property (antider(hi) - antider(lo)) / (hi - lo)
# (A/AlwaysDefinedSpline.scala:43-46) cancels for neighbouring temperatures, so a last-bit difference of `predict` moves
# the reference's OWN next step by up to ~1e-11 (measured by _oracle_sensitivity; profiles/r2_parity.json has the
# numbers at full size).  The tolerance family therefore cannot hold 1e-12 on tables after several steps; the bar for
# it is this FIXED constant (both measured errors are logged), and the within-1e-12 path for tables is the bit-exact
# solver (K7, tests/test_gpu_k7.py).
TOL_TABLES = 2e-11


@pytest.fixture
def seg_len(ctx):
    def set_(m):
        ctx.set_option("seg_len", m)

    yield set_
    ctx.set_option("seg_len", 32)


def _compare_tol(got, want, what, tol=TOL, tol_elem=None):
    for k, v in want.items():
        check_tol(got[k], v, f"{what} {k}", tol, tol_elem)


def _oracle_sensitivity(O, p, steps):
    """How much the REFERENCE algorithm itself moves when `predict` changes by one ulp.  With spline tables the
    face property is (antider(hi) - antider(lo)) / (hi - lo): for neighbouring temperatures a fraction of a degree
    apart the cancellation amplifies a 1e-16 input difference to ~1e-11 in the coefficient, so no implementation
    with a different operation order can agree with the reference better than this, whatever its own accuracy."""
    a, b = p.build_oracle(O, ascending=False), p.build_oracle(O, ascending=False)
    b.predict[:] = np.nextafter(b.predict, np.inf)
    worst = 0.0
    for _ in range(steps):
        a.iteration(900.0); b.iteration(900.0)
        worst = max(worst, max_rel_err(a.grid, b.grid), max_rel_err(a.predict, b.predict))
    O.set_ascending_fold(True)
    return worst


@pytest.mark.parametrize("combo", ["".join(c) for c in itertools.product("TF", repeat=4)])
@pytest.mark.parametrize("m", [2, 5, 8])
def test_partitioned_all_boundary_kinds(O, gs, seg_len, combo, m):
    seg_len(m)
    xdir = 1 if combo.count("F") % 2 else 0
    p = cases.bc_problem(combo, xdir=xdir, ydir=0, cols=41, rows=35, dense=True)
    g, o = p.build_gs(gs), p.build_oracle(O)
    g.set_solver(gs.SOLVER_PARTITIONED)
    g.xIter(900.0); o.xIter(900.0)
    check_tol(g.grid.result, o.result, f"partitioned xIter {combo} m={m}", TOL)
    g.yIter(900.0); o.yIter(900.0)
    check_tol(g.grid.result, o.result, f"partitioned yIter {combo} m={m}", TOL)
    g.grid.update(); o.update()
    g.iteration(900.0, nsteps=3)
    for _ in range(3):
        o.iteration(900.0)
    _compare_tol(cases.snapshot(g), cases.snapshot(o), f"{combo} m={m}")
    assert g.status == 0


@pytest.mark.parametrize("cols,rows,m", [(4, 4, 2), (5, 9, 2), (33, 31, 4), (64, 65, 32), (130, 97, 7), (300, 270, 32)])
@pytest.mark.parametrize("combo", ["TTTT", "FFFF", "TFTF"])
def test_partitioned_ragged_sizes(O, gs, seg_len, cols, rows, m, combo):
    seg_len(m)
    p = cases.bc_problem(combo, xdir=1, ydir=1, cols=cols, rows=rows)
    g, o = p.build_gs(gs), p.build_oracle(O)
    g.set_solver(gs.SOLVER_PARTITIONED)
    g.iteration(900.0, nsteps=2)
    o.iteration(900.0); o.iteration(900.0)
    _compare_tol(cases.snapshot(g), cases.snapshot(o), f"{cols}x{rows} m={m} {combo}")


@pytest.mark.parametrize("coefs", ["hermite", "per_row", "per_col", "dense"])
def test_partitioned_coefficient_maps_and_tables(O, gs, seg_len, coefs):
    seg_len(6)
    p = cases.bc_problem("TTFT", xdir=1, cols=37, rows=29, coefs=coefs)
    r = np.random.default_rng(4)
    p.grid0 = r.uniform(-15.0, 70.0, p.grid0.shape)
    p.predict0 = p.grid0 + r.uniform(-12.0, 12.0, p.grid0.shape)
    g, o = p.build_gs(gs), p.build_oracle(O, ascending=False)
    g.set_solver(gs.SOLVER_PARTITIONED)
    g.iteration(900.0, nsteps=3)
    for _ in range(3):
        o.iteration(900.0)
    O.set_ascending_fold(True)
    _compare_tol(cases.snapshot(g), cases.snapshot(o), f"partitioned {coefs}", tol=TOL_TABLES)


def test_partitioned_gridtest_and_medium(O, gs, seg_len):
    """AT/GridTest.scala with the partitioned solver: wall temperature within 1e-12 of KAT R1; then 512 x 384
    (the AUTO solver's choice at that size) against the oracle."""
    seg_len(8)
    p = cases.gridtest_problem()
    g = p.build_gs(gs)
    g.set_solver(gs.SOLVER_PARTITIONED)
    g.fill(7.0)
    got = cases.run_gridtest(g, True)
    want = np.load(os.path.join(GOLDEN, "gridtest_137x45x195.npz"))
    _compare_tol(got, {k: want[k] for k in want.files}, "GridTest")
    assert abs(got["grid"][0, 0] - 54.44328715390563) <= 1e-12 * 54.44328715390563
    seg_len(32)
    p = cases.bc_problem("TTFF", cols=512, rows=384)
    g, o = p.build_gs(gs), p.build_oracle(O)
    g.iteration(900.0, nsteps=3)   # AUTO
    for _ in range(3):
        o.iteration(900.0)
    _compare_tol(cases.snapshot(g), cases.snapshot(o), "512x384 auto")


@pytest.mark.parametrize("cols,rows", [(128, 128), (130, 257), (512, 384), (1000, 140), (129, 1025)])
@pytest.mark.parametrize("combo", ["TTTT", "FFFF", "TFFT", "FTTF"])
def test_k5_hoisted_constant_coefficients(O, gs, cols, rows, combo):
    """K5: literal Spline.const coefficients -> line-independent matrix, table-driven sweeps (tolerance solver)."""
    p = cases.bc_problem(combo, xdir=1, ydir=0, cols=cols, rows=rows, coefs="const")
    g, o = p.build_gs(gs), p.build_oracle(O)
    g.set_solver(gs.SOLVER_PARTITIONED)   # K5 for both directions at these sizes
    g.iteration(900.0, nsteps=3)
    for _ in range(3):
        o.iteration(900.0)
    _compare_tol(cases.snapshot(g), cases.snapshot(o), f"K5 {cols}x{rows} {combo}")
    # xIter alone (K5 x sweep) and yIter alone (general partitioned: K5's y sweep only exists fused)
    g.xIter(900.0); o.xIter(900.0)
    assert max_rel_err(g.grid.result, o.result) <= TOL
    g.yIter(900.0); o.yIter(900.0)
    assert max_rel_err(g.grid.result, o.result) <= TOL
    assert g.status == 0


@pytest.mark.parametrize("coefs", ["hermite", "per_row", "per_col", "dense", "const"])
@pytest.mark.parametrize("combo,cols,rows", [("TTTT", 130, 257), ("FFTF", 300, 140), ("TFFT", 128, 128), ("FTFT", 513, 129)])
def test_k6_faces_plus_elimination(O, gs, ctx, coefs, combo, cols, rows):
    """K6: every spline evaluation in an elementwise faces kernel, then K5-style fused partitioned sweeps over the
    precomputed face values (temperature-dependent tables, every coefficient-map mode, all bound kinds)."""
    ctx.set_option("k5", 0)   # constant coefficients would otherwise take K5
    try:
        p = cases.bc_problem(combo, xdir=1, ydir=0, cols=cols, rows=rows, coefs=coefs)
        r = np.random.default_rng(6)
        p.grid0 = r.uniform(-15.0, 70.0, p.grid0.shape)
        p.predict0 = p.grid0 + r.uniform(-12.0, 12.0, p.grid0.shape)
        g, o = p.build_gs(gs), p.build_oracle(O, ascending=False)
        g.set_solver(gs.SOLVER_PARTITIONED)
        g.iteration(900.0, nsteps=3)
        for _ in range(3):
            o.iteration(900.0)
        O.set_ascending_fold(True)
        tol = TOL if coefs == "const" else TOL_TABLES
        _compare_tol(cases.snapshot(g), cases.snapshot(o), f"K6 {coefs} {combo} {cols}x{rows}", tol=tol)
        g.xIter(900.0); o.xIter(900.0)
        assert max_rel_err(g.grid.result, o.result) <= tol
    finally:
        ctx.set_option("k5", 1)


@pytest.mark.parametrize("coefs,combo,cols,rows", [("hermite", "TTTT", 300, 260), ("dense", "TFFT", 257, 301),
                                                   ("hermite", "FTTF", 1024, 64)])
def test_k6_fused_faces_variant(O, gs, ctx, coefs, combo, cols, rows):
    """Option k6_fuse: phase A of the sweep evaluates the faces along its lines itself (no faces pass)."""
    ctx.set_option("k5", 0)
    ctx.set_option("k6_fuse", 1)
    try:
        p = cases.bc_problem(combo, xdir=1, ydir=0, cols=cols, rows=rows, coefs=coefs)
        r = np.random.default_rng(13)
        p.grid0 = r.uniform(-15.0, 70.0, p.grid0.shape)
        p.grid0[:30, :40] = 11.0
        p.predict0 = p.grid0 + r.uniform(-3.0, 3.0, p.grid0.shape) * (p.grid0 != 11.0)
        g, o = p.build_gs(gs), p.build_oracle(O, ascending=False)
        g.set_solver(gs.SOLVER_PARTITIONED)
        g.iteration(900.0, nsteps=3)
        for _ in range(3):
            o.iteration(900.0)
        O.set_ascending_fold(True)
        tol = TOL_TABLES
        _compare_tol(cases.snapshot(g), cases.snapshot(o), f"K6 fused faces {coefs}", tol=tol)
        assert g.status == 0
    finally:
        ctx.set_option("k5", 1)
        ctx.set_option("k6_fuse", 0)


@pytest.mark.parametrize("coefs", ["lines_uneven", "hermite_uneven"])
def test_k6_tables_outside_the_fast_tier(O, gs, ctx, coefs):
    """Piecewise-linear tables and unevenly spaced knots: every face goes through the second tier or the general walk."""
    p = cases.bc_problem("TTFT", xdir=1, ydir=0, cols=160, rows=256, coefs=coefs)
    r = np.random.default_rng(12)
    p.grid0 = r.uniform(-20.0, 80.0, p.grid0.shape)
    p.grid0[:40, :50] = 12.0
    p.predict0 = p.grid0 + r.uniform(-1.0, 1.0, p.grid0.shape) * (p.grid0 != 12.0)
    g, o = p.build_gs(gs), p.build_oracle(O, ascending=False)
    g.set_solver(gs.SOLVER_PARTITIONED)
    g.iteration(900.0, nsteps=2)
    for _ in range(2):
        o.iteration(900.0)
    O.set_ascending_fold(True)
    tol = TOL_TABLES
    _compare_tol(cases.snapshot(g), cases.snapshot(o), f"K6 {coefs}", tol=tol)
    assert g.status == 0


@pytest.mark.parametrize("field", ["uniform", "clamped", "blocks", "knots"])
def test_k6_face_tiers(O, gs, ctx, field):
    """The faces kernel answers the common face (two different temperatures inside one table piece) in straight-line
    code and hands everything else to slower tiers: equal temperatures (average = apply), one straddled knot, and the
    reference's general walk (clamps, several pieces, values on a knot).  Each field below lives in one of those."""
    ctx.set_option("k5", 0)
    try:
        p = cases.bc_problem("TTTT", xdir=0, ydir=1, cols=256, rows=288, coefs="hermite")
        r = np.random.default_rng(11)
        shape = p.grid0.shape
        if field == "uniform":
            g0 = np.full(shape, 4.0)
        elif field == "clamped":
            g0 = r.uniform(-60.0, 140.0, shape)          # the tables cover a narrower range: clamp areas on many faces
        elif field == "blocks":
            g0 = np.kron(r.integers(-2, 8, (shape[0] // 16, shape[1] // 16)) * 9.5, np.ones((16, 16)))
        else:
            g0 = r.integers(-2, 8, shape) * 10.0         # every temperature exactly on a knot
        p.grid0 = g0.astype(float)
        p.predict0 = p.grid0.copy()
        g, o = p.build_gs(gs), p.build_oracle(O, ascending=False)
        g.set_solver(gs.SOLVER_PARTITIONED)
        g.iteration(900.0, nsteps=2)
        for _ in range(2):
            o.iteration(900.0)
        O.set_ascending_fold(True)
        tol = TOL_TABLES
        _compare_tol(cases.snapshot(g), cases.snapshot(o), f"K6 tiers {field}", tol=tol)
    finally:
        ctx.set_option("k5", 1)


@pytest.mark.parametrize("cols,rows", [(1024, 64), (64, 1024), (2048, 300)])
def test_k6_clustered_strips(O, gs, ctx, cols, rows):
    """Few lines in one direction: a strip of 32 lines is eliminated by a cluster of 2 or 4 CTAs whose reduced system
    is solved through distributed shared memory.  Same answer as one CTA per strip, and as the reference."""
    ctx.set_option("k5", 0)
    try:
        p = cases.bc_problem("TFTT", xdir=1, ydir=0, cols=cols, rows=rows, coefs="hermite")
        r = np.random.default_rng(8)
        p.grid0 = r.uniform(-10.0, 60.0, p.grid0.shape)
        p.predict0 = p.grid0 + r.uniform(-2.0, 2.0, p.grid0.shape)
        g, o = p.build_gs(gs), p.build_oracle(O, ascending=False)
        g.set_solver(gs.SOLVER_PARTITIONED)
        g.iteration(900.0, nsteps=2)
        for _ in range(2):
            o.iteration(900.0)
        O.set_ascending_fold(True)
        tol = TOL_TABLES
        _compare_tol(cases.snapshot(g), cases.snapshot(o), f"K6 clusters {cols}x{rows}", tol=tol)
        ctx.set_option("k6_cluster", 0)
        h = p.build_gs(gs)
        h.set_solver(gs.SOLVER_PARTITIONED)
        h.iteration(900.0, nsteps=2)
        _compare_tol(cases.snapshot(g), cases.snapshot(h), "clusters vs single CTA", tol=tol)
    finally:
        ctx.set_option("k5", 1)
        ctx.set_option("k6_cluster", 1)


@pytest.mark.parametrize("cols,rows,combo", [(2048, 64, "TTFF"), (64, 2048, "TFTT"), (4096, 200, "TTTT")])
def test_k5_clustered_strips(O, gs, ctx, cols, rows, combo):
    """Constant coefficients, few lines in one direction: a strip's segments are spread over a cluster of up to 8
    CTAs (distributed shared memory for the reduced system).  Same answer as one CTA per strip, and as the reference."""
    p = cases.bc_problem(combo, xdir=1, ydir=0, cols=cols, rows=rows, coefs="const")
    g, o = p.build_gs(gs), p.build_oracle(O)
    g.set_solver(gs.SOLVER_PARTITIONED)
    g.iteration(900.0, nsteps=2)
    for _ in range(2):
        o.iteration(900.0)
    _compare_tol(cases.snapshot(g), cases.snapshot(o), f"K5 clusters {cols}x{rows}")
    ctx.set_option("k5_cluster", 0)
    try:
        h = p.build_gs(gs)
        h.set_solver(gs.SOLVER_PARTITIONED)
        h.iteration(900.0, nsteps=2)
        _compare_tol(cases.snapshot(g), cases.snapshot(h), "clusters vs single CTA")
    finally:
        ctx.set_option("k5_cluster", 1)


@pytest.mark.parametrize("cols,rows,combo", [(512, 384, "TTFF"), (1000, 140, "FTTF"), (130, 1025, "TFFT"), (4096, 300, "TTTT"),
                                             (2048, 64, "TTFF"), (64, 2048, "TFTT"), (8192, 40, "FFTT")])
def test_k5t_tma_rings_match_k5(O, gs, ctx, cols, rows, combo):
    """K5T (tensor-map TMA rings: swizzled / wide boxes, box stores, clusters of any size, soft groups through global
    memory) runs K5's arithmetic: same answer as k5_sweep (k5t = 0) far inside the tolerance, and as the reference."""
    p = cases.bc_problem(combo, xdir=1, ydir=0, cols=cols, rows=rows, coefs="const")
    o = p.build_oracle(O)
    for _ in range(2):
        o.iteration(900.0)
    want = cases.snapshot(o)
    runs = {}
    variants = [("k5", {"k5t": 0}), ("k5t", {}), ("wide x", {"k5t_wide": 2}), ("narrow x", {"k5t_wide": 0, "k5t_sx": 8}),
                ("no soft groups", {"k5t_soft": 0}), ("hints", {"k5t_hints": 3, "k5t_promo": 0}), ("short rings", {"k5t_nt": 2})]
    defaults = {"k5t": 1, "k5t_wide": 1, "k5t_sx": 0, "k5t_soft": 1, "k5t_hints": 0, "k5t_promo": 1, "k5t_nt": 0}
    try:
        for name, opts in variants:
            for k, v in {**defaults, "k5t": 2, **opts}.items():   # k5t = 2: K5T also below its size threshold
                ctx.set_option(k, v)
            g = p.build_gs(gs)
            g.set_solver(gs.SOLVER_PARTITIONED)
            g.iteration(900.0, nsteps=2)
            runs[name] = cases.snapshot(g)
            assert g.status == 0
            _compare_tol(runs[name], want, f"{name} {cols}x{rows} vs oracle")
    finally:
        for k, v in defaults.items():
            ctx.set_option(k, v)
    for name in runs:
        for k in ("grid", "predict"):
            assert max_rel_err(runs[name][k], runs["k5"][k]) <= 2e-13, (name, k)


def test_k5t_x_sweep_bits_equal_k5_with_the_same_segments(gs, ctx):
    """With the same segmentation (16 segments of 256 columns) K5T's x sweep and k5_sweep perform the same operations
    in the same order: `result` is bit-identical."""
    p = cases.bc_problem("TFFT", xdir=0, ydir=0, cols=4096, rows=70, coefs="const")
    out = {}
    try:
        for name, opts in (("k5", {"k5t": 0, "k5_cluster": 0}), ("k5t", {"k5t": 2, "k5_cluster": 0, "k5t_wide": 0, "k5t_sx": 16})):
            for k, v in opts.items():
                ctx.set_option(k, v)
            g = p.build_gs(gs)
            g.set_solver(gs.SOLVER_PARTITIONED)
            g.xIter(900.0)
            out[name] = np.array(g.grid.result)
    finally:
        for k, v in {"k5t": 1, "k5_cluster": 1, "k5t_wide": 1, "k5t_sx": 0}.items():
            ctx.set_option(k, v)
    assert_bit_equal(out["k5t"], out["k5"], "K5T x sweep vs k5_sweep")


def test_k5_on_caller_owned_arrays_that_are_not_16_byte_aligned(gs):
    """gs_grid2d_bind_device with buffers that start 8 bytes into an allocation: tensor maps need 16-byte alignment, so the
    constant-coefficient sweeps fall back from K5T to k5_sweep (clusters, not soft groups) -- same answer as library-owned
    arrays far inside the tolerance."""
    import torch

    p = cases.bc_problem("TFFT", xdir=1, ydir=0, cols=2048, rows=96, coefs="const")
    ref = p.build_gs(gs)
    ref.set_solver(gs.SOLVER_PARTITIONED)
    ref.iteration(900.0, nsteps=2)
    want = cases.snapshot(ref)
    g = p.build_gs(gs)
    g.set_solver(gs.SOLVER_PARTITIONED)
    n = p.grid0.size
    keep = []
    for which, init in ((gs.GRID, p.grid0), (gs.PREDICT, p.predict0 if hasattr(p, "predict0") else p.grid0), (gs.RESULT, None)):
        buf = torch.zeros(n + 1, dtype=torch.float64, device="cuda")
        view = buf[1:]
        assert view.data_ptr() % 16 == 8
        if init is not None:
            view.copy_(torch.from_numpy(np.ascontiguousarray(init, dtype=np.float64).ravel()))
        torch.cuda.synchronize()
        g.bind_device(which, view.data_ptr())
        keep.append(buf)
    g.iteration(900.0, nsteps=2)
    got = cases.snapshot(g)
    assert g.status == 0
    for k in ("grid", "predict"):
        assert max_rel_err(got[k], want[k]) <= 2e-13, k


@pytest.mark.parametrize("cols,rows,coefs,solver", [(4096, 300, "const", 2), (2048, 64, "const", 2), (64, 2048, "const", 2),
                                                    (1000, 140, "const", 2), (512, 384, "hermite", 1), (130, 257, "hermite", 1)])
def test_ring_kernels_are_deterministic(gs, cols, rows, coefs, solver):
    """compute-sanitizer is closed on this pool (profiles/r2_sanitizer_closed.txt), so the mbarrier rings, cluster /
    soft-group rendezvous and in-place shared-memory rewrites of K5T and K7F are checked the other way round: a race there
    shows up as run-to-run differences.  Five runs of three steps from the same state must agree bit for bit."""
    sol = gs.SOLVER_PARTITIONED if solver == 2 else gs.SOLVER_THOMAS
    p = cases.bc_problem("TFFT" if coefs == "const" else "TTTT", xdir=1, ydir=0, cols=cols, rows=rows, coefs=coefs)
    first = None
    gs.default_context().set_option("k5t", 2)   # K5T also below its size threshold
    try:
        for run in range(5):
            g = p.build_gs(gs)
            g.set_solver(sol)
            g.iteration(900.0, nsteps=3)
            snap = cases.snapshot(g)
            assert g.status == 0
            if first is None:
                first = snap
            else:
                for k in ("grid", "predict"):
                    assert_bit_equal(snap[k], first[k], f"run {run} {k}")
    finally:
        gs.default_context().set_option("k5t", 1)


def test_k5_triple_buffer_variant(O, gs, ctx):
    """Option k5_nb = 3: three tile buffers per warp on the y sweep (measured slower, kept as a tested variant)."""
    ctx.set_option("k5_nb", 3)
    try:
        p = cases.bc_problem("TTFF", xdir=0, ydir=0, cols=200, rows=1500, coefs="const")
        g, o = p.build_gs(gs), p.build_oracle(O)
        g.set_solver(gs.SOLVER_PARTITIONED)
        g.iteration(900.0, nsteps=2)
        for _ in range(2):
            o.iteration(900.0)
        _compare_tol(cases.snapshot(g), cases.snapshot(o), "K5 triple buffer")
    finally:
        ctx.set_option("k5_nb", 2)


def test_k5_tables_follow_coefficient_and_bound_changes(O, gs):
    """K5 caches its line-independent tables per (grid, direction, time step): changing the coefficients, the time
    step or a bound's kind must rebuild them."""
    p = cases.bc_problem("TTTT", xdir=0, ydir=0, cols=256, rows=256, coefs="const")
    g, o = p.build_gs(gs), p.build_oracle(O)
    g.set_solver(gs.SOLVER_PARTITIONED)
    g.iteration(900.0); o.iteration(900.0)
    # other constants
    g.set_coefs(gs.ConstantCoef(gs.Spline.const(2.5), gs.Spline.const(1.1e6), gs.Spline.const(2.5 / 1.1e6)))
    ads = [O.AlwaysDefinedSpline(O.Spline.const(v)) for v in (2.5, 1.1e6, 2.5 / 1.1e6)]
    o.set_constant_coef(*ads)
    g.iteration(900.0); o.iteration(900.0)
    # other time step
    g.iteration(450.0); o.iteration(450.0)
    # a bound changes kind (Temperature -> HeatFlow)
    q = np.full(256, 12.0)
    g.bounds.left.kind = gs.FLOW
    g.bounds.left.update(q)
    o.set_bound(O.LEFT, O.FLOW, O.DENSE, q)
    g.iteration(450.0); o.iteration(450.0)
    _compare_tol(cases.snapshot(g), cases.snapshot(o), "K5 cache invalidation")


def test_slices_and_coordinate_fills(gs):
    """row / col / left / right / upper / lower, updateRow / updateColumn, updateX / updateY (A/TwoDGrid.scala:29-58,
    260-352): device-side gathers and scatters; the writers touch `grid` only."""
    p = cases.bc_problem("TTTT", cols=37, rows=29, coefs="const")
    g = p.build_gs(gs)
    G, P = g.grid.grid.copy(), g.grid.predict.copy()
    for r in (0, 11, 28):
        assert_bit_equal(g.row(r), G[r])
    for c in (0, 5, 36):
        assert_bit_equal(g.col(c), G[:, c])
    assert_bit_equal(g.left(), G[:, 0]); assert_bit_equal(g.right(1), G[:, -2])
    assert_bit_equal(g.upper(2), G[2]); assert_bit_equal(g.lower(), G[-1])
    buf = np.zeros(37); g.row(3, buf); assert_bit_equal(buf, G[3])
    assert_bit_equal(g.row(4, which=gs.PREDICT), P[4])
    g.updateRow(7, np.arange(37.0)); G[7] = np.arange(37.0)
    g.updateColumn(9, -np.arange(29.0)); G[:, 9] = -np.arange(29.0)
    assert_bit_equal(g.grid.grid, G); assert_bit_equal(g.grid.predict, P)
    g.updateX(lambda x: 3.0 * x)
    want = np.tile(np.array([3.0 * g.x.coord(c) for c in range(37)]), (29, 1))
    assert_bit_equal(g.grid.grid, want); assert_bit_equal(g.grid.predict, P)
    g.updateY(lambda y: 1.0 + y)
    want = np.repeat(np.array([1.0 + g.y.coord(r) for r in range(29)])[:, None], 37, axis=1)
    assert_bit_equal(g.grid.grid, want); assert_bit_equal(g.grid.predict, P)
    with pytest.raises(gs.GsError):
        g.row(29)
    with pytest.raises(gs.GsError):
        g.updateColumn(37, np.zeros(29))


@pytest.mark.parametrize("which,factory", [(1, "lines"), (2, "m1Hermite3"), (3, "fHermite3")])
def test_spline_quotient_matches_reference_resampling(O, gs, which, factory):
    """Spline./ (P/Spline.scala:198-210): k(T) / c(T) resampled at the merged knots and rebuilt -- how the reference
    derives the diffusivity table (ConstantCoef, A/TwoDGrid.scala:790-816)."""
    rho = [0.598, 0.826, 1.121, 1.496, 1.966, 2.547, 3.258, 4.122, 5.156, 6.397, 7.862]
    T = [-20.0 + 10.0 * i for i in range(11)]
    kp = [(t, 1.0 + 0.1 * q) for t, q in zip(T, rho)]
    cp = [(t + (2.5 if i % 3 == 1 else 0.0), 2.0e6 + 1.0e5 * q) for i, (t, q) in enumerate(zip(T, rho))]   # other knots
    ok, oc = O.Spline.m1Hermite3(kp), O.Spline.m1Hermite3(cp)
    gk, gc = gs.Spline.m1Hermite3(kp), gs.Spline.m1Hermite3(cp)
    want = ok.div(oc, which=which)
    got = gk.combine(gc, lambda a, b: a / b, factory=getattr(gs.Spline, factory))
    pieces = want.pieces()
    assert got.size == len(pieces)
    xs = np.linspace(T[0], T[-2], 301)
    assert_bit_equal(got.apply(xs), np.array([want(float(x)) for x in xs]))
    # operators use the default (M1Hermite3) factory
    if which == 2:
        assert_bit_equal((gk / gc).apply(xs), got.apply(xs))
        s = gk + gc
        assert_bit_equal(s.apply(xs[:5]), gs.Spline.m1Hermite3([(x, float(a) + float(b)) for x, a, b in
                         zip(gk._sum_arguments(gc, False), gk.apply(gk._sum_arguments(gc, False)),
                             gc.apply(gk._sum_arguments(gc, False)))]).apply(xs[:5]))
    assert gs.Spline.line(0.0, 1.0, 1.0, 2.0).combine(gs.Spline.line(5.0, 1.0, 6.0, 2.0), lambda a, b: a / b) is None


def test_text_output_formats(gs, tmp_path):
    """f-4: TwoDGrid.write / toString (A/TwoDGrid.scala:62-133) and OneDGrid.writeGrid (A/OneDGrid.scala:37-63)."""
    import datetime
    import io

    x = gs.XDim(1.0, [2.0, 3.0, 4.0, 5.0], 6.0, gs.RADIAL)
    y = gs.YDim(0.0, [10.0, 20.0], 30.0, gs.ORTHO)
    g = gs.TwoDGrid(x, y, upp=(gs.Many, gs.Temp), low=(gs.One, gs.Temp), right=(gs.Many, gs.Temp), left=(gs.One, gs.Flow),
                    coefs=gs.ConstantCoef(gs.Spline.const(1.5), gs.Spline.const(2.2e6), gs.Spline.const(1.5 / 2.2e6)))
    g.grid.grid = np.array([[1234.5678, 7.0, 0.0005, -0.0004], [0.0625, 2.0, 54.44328715390563, 1e6]])
    g.bounds.upp.update([1.0, 2.0, 3.0, 4.0]); g.bounds.low.update(9.5)
    g.bounds.left.update(0.25); g.bounds.right.update([7.0, 8.12345])
    buf = io.StringIO()
    g.write(buf, newline="\n")
    assert buf.getvalue() == "10 1,234.568 7 0.001 -0\n20 0.062 2 54.443 1,000,000\n"
    s = g.toString(sep="\n").split("\n")
    assert s[0] == "X/Y\t\t|\t1.000\t|\t2.000\t3.000\t4.000\t5.000\t|\t6.000\t|\t..."
    assert s[1] == "0.000\t|\t...\t\t|\t9.500\t9.500\t9.500\t9.500\t|\t...\t\t|\t30.000"
    assert s[2] == "10.000\t|\t0.250\t|\t1234.568\t7.000\t0.001\t-0.000\t|\t7.0000\t|\t10.000"
    assert s[3] == "20.000\t|\t0.250\t|\t0.063\t2.000\t54.443\t1000000.000\t|\t8.1235\t|\t20.000"
    assert s[4] == "30.000\t|\t...\t\t|\t1.000\t2.000\t3.000\t4.000\t|\t...\t\t|\t30.000"
    assert s[5] == s[0]
    o = gs.OneDGrid(0.0, [0.125, 0.25, 0.375], 0.5, gs.Spline.const(1e-6), 1.0)
    o.grid = [4.0, 1234.5678, 0.5]
    name = o.writeGrid(datetime.datetime(2026, 10, 20, 9, 30), str(tmp_path), newline="\n")
    assert os.path.basename(name) == "grid OCTOBER 20d 9h 2026.dat"
    assert open(name, newline="").read() == ("# One dimension grid output results\r\n# Coords \t Temperatures\r\n\n"
                                             "0.12\t4\r\n\n0.25\t1,234.568\r\n\n0.38\t0.5\r\n\n")
