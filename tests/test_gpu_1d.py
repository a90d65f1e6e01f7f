"""GPU parity of the batched 1-D stepper (K2) against the oracle and the committed fixtures.
Per-thread Thomas keeps the reference's operation order: every comparison is bit-exact."""
import os

import numpy as np
import pytest

import cases
from helpers import assert_bit_equal

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def test_c1_radial_200_nodes_1000_steps(O, gs):
    got = cases.run_c1(gs, steps=1000)
    want = np.load(os.path.join(GOLDEN, "c1_radial_200x1000.npz"))
    for k in ("grid", "predict"):
        assert_bit_equal(got[k], want[k], f"C1 {k} vs golden")
    g = got["grid"]
    assert g[0] == 19.163547799904077 and g[99] == 6.303223056805375 and g[199] == 4.014707461585766  # KAT R2
    ref = cases.run_c1(O, steps=1000, oracle=True)
    for k in ("grid", "predict"):
        assert_bit_equal(got[k], ref[k], f"C1 {k} vs oracle")


def test_c1_step_by_step_equals_one_call(gs):
    a = cases.run_c1(gs, steps=7)
    s = cases.c1_setup()
    g = gs.OneDGrid(s["leftX"], s["rangeX"], s["rightX"], gs.Spline.const(1e-6), s["sigma"], direction=gs.RADIAL)
    for _ in range(7):
        g.step(s["time"], s["leftT"], s["rightT"])
    assert_bit_equal(g.grid, a["grid"])
    assert_bit_equal(g.predict, a["predict"])


@pytest.mark.parametrize("name,args", [
    ("batch_const_37x256x5", dict(nlines=37, n=256, steps=5, spline="const")),
    ("batch_hermite_19x64x4", dict(nlines=19, n=64, steps=4, spline="hermite", random_grid=True)),
])
def test_batch_vs_golden(gs, name, args):
    got = cases.run_batch(gs, **args)
    want = np.load(os.path.join(GOLDEN, name + ".npz"))
    for k in ("grid", "predict"):
        assert_bit_equal(got[k], want[k], f"{name}.{k}")


@pytest.fixture
def options(ctx):
    """Sets tuning / path-selection options on the shared context and restores the defaults afterwards."""
    defaults = {"const_hoist": 1, "pipeline": 1, "native_div": 0, "pipe_cfg": 0, "threads_per_sm_1d": 0,
                "hoist_ck": 1, "ck_block": 16, "faces1d": 0, "k2t": 1, "k2t_cps": 8, "k2t_hints": 31, "k2t_ctas": 0, "k2r": 1}

    def set_(**kw):
        for k, v in kw.items():
            ctx.set_option(k, v)

    yield set_
    for k, v in defaults.items():
        ctx.set_option(k, v)


# every code path of K2: hoisted coefficients (K2h), the general TMA pipeline, the direct-load kernel,
# plain IEEE division instead of the shared-reciprocal one, other ring shapes
PATHS = {
    "default": {},                                          # constants: K2r for these small batches (K2c beyond 2 blocks per SM)
    "k2r_native_div": {"native_div": 1},
    "k2c": {"k2r": 0},                                      # K2c: hoisted coefficients, checkpoint / recompute
    "hoist_scratch": {"hoist_ck": 0, "k2r": 0},
    "hoist_ck32": {"ck_block": 32, "k2r": 0},
    "hoist_ck_native": {"native_div": 1, "k2r": 0},
    "general": {"const_hoist": 0, "k2t": 0},                # K2: TMA ring + (delta, lambda) scratch
    "k2t": {"const_hoist": 0, "k2t": 2},                            # K2T: faces warp + chain warp, L2-resident scratch (2: any table; cubic tables on equal knots take it by default)
    "k2t_native_div": {"const_hoist": 0, "native_div": 1, "k2t": 2},
    "k2t_cps4_nohints": {"const_hoist": 0, "k2t": 2, "k2t_cps": 4, "k2t_hints": 0},   # three-stage rings, no L2 hints / discards
    "k2t_few_ctas": {"const_hoist": 0, "k2t": 2, "k2t_ctas": 3},      # persistent CTAs: every CTA walks several line blocks
    "k2f": {"const_hoist": 0, "faces1d": 1, "k2t": 0},      # K2f: faces kernel + elimination with checkpoint/recompute
    "k2f_native_div": {"const_hoist": 0, "faces1d": 1, "native_div": 1, "k2t": 0},
    "direct": {"const_hoist": 0, "pipeline": 0},
    "native_div": {"native_div": 1, "hoist_ck": 0, "k2r": 0},
    "general_native_div": {"const_hoist": 0, "native_div": 1, "k2t": 0},
    "ring1": {"pipe_cfg": 1, "hoist_ck": 0, "k2r": 0},
    "general_ring2": {"const_hoist": 0, "pipe_cfg": 2, "k2t": 0},
    "few_warps": {"threads_per_sm_1d": 32, "k2t": 0, "k2r": 0},
}


@pytest.mark.parametrize("nlines,n", [(1, 2), (3, 3), (33, 5), (1000, 256), (4099, 33), (257, 1024), (40, 4099)])
@pytest.mark.parametrize("spline", ["const", "hermite", "lines"])
@pytest.mark.parametrize("path", list(PATHS))
def test_batch_vs_oracle(O, gs, options, nlines, n, spline, path):
    options(**PATHS[path])
    kw = dict(nlines=nlines, n=n, steps=3, spline=spline, random_grid=True)
    got = cases.run_batch(gs, solver=gs.SOLVER_THOMAS, **kw)   # one chain per line: the reference's operation order
    want = cases.run_batch(O, oracle=True, **kw)
    for k in ("grid", "predict"):
        assert_bit_equal(got[k], want[k], f"{path} {spline} {nlines}x{n} {k}")


def test_zero_and_tiny_values_take_the_ieee_division_path(O, gs):
    """Zeros / denormal-range values are outside the proven range of the shared-reciprocal division; the warp
    redoes the block with plain division.  Results stay bit-exact."""
    nlines, n = 70, 64
    s = cases.batch_setup(nlines, n)
    grid0 = np.zeros((nlines, n))
    grid0[::3, ::5] = 1e-300
    grid0[1::7] = np.random.default_rng(3).uniform(0.0, 30.0, n)
    for const in (True, False):
        spl = (lambda m: m.Spline.const(1e-6)) if const else (
            lambda m: m.Spline.lines([(-20.0 + 25.0 * i, 1e-6 * (1.0 + 0.3 * i)) for i in range(5)]))
        g = gs.BatchedOneDGrid(s["leftX"], s["rangeX"], s["rightX"], spl(gs), 1.0, nlines=nlines)
        g.upload(gs.GRID, grid0)
        g.fill(gs.PREDICT, 0.0)
        lt, rt = np.zeros(nlines), np.zeros(nlines)
        lt[::2] = 5.0
        g.step(900.0, lt, rt, nsteps=3)
        og, op = grid0.copy(), np.zeros((nlines, n))
        O.batch1d_step(0, s["leftX"], s["rangeX"], s["rightX"], spl(O), 1.0, og, op, lt, rt, 900.0, nsteps=3)
        assert_bit_equal(g.download(gs.GRID), og, "zeros grid")
        assert_bit_equal(g.download(gs.PREDICT), op, "zeros predict")


@pytest.mark.parametrize("n", [32, 48, 64, 250])
def test_k2t_special_values(O, gs, n):
    """K2T (the default for a cubic table on equal knots) on data that leaves its straight-line tiers: zeros and
    denormal-range values (the block is redone with plain division), temperatures beyond the table's clamps and exactly on
    its knots (general evaluation out of line), line ends far outside the table; full and partial last chunks; padding lanes
    (70 lines = 2 blocks + 6 lanes).  Bit-exact against the oracle, and the same bits as K2."""
    nlines = 70
    s = cases.batch_setup(nlines, n)
    r = np.random.default_rng(11)
    pts = [(-20.0 + 10.0 * i, 1e-6 * (1.0 + 0.1 * cases.README_RHO[i])) for i in range(11)]
    grid0 = r.uniform(-60.0, 120.0, (nlines, n))          # well beyond the table's [-20, 80]
    grid0[::3, ::5] = 0.0
    grid0[1::7, ::4] = 1e-300
    grid0[2::9] = np.round(grid0[2::9] / 10.0) * 10.0      # exactly on knots
    grid0[5] = 7.0                                         # a uniform line
    lt, rt = r.uniform(-100.0, 150.0, nlines), r.uniform(-100.0, 150.0, nlines)
    lt[::4] = 0.0
    outs = {}
    for k2t in (1, 0):
        gs.default_context().set_option("k2t", k2t)
        try:
            g = gs.BatchedOneDGrid(s["leftX"], s["rangeX"], s["rightX"], gs.Spline.m1Hermite3(pts), 1.0, nlines=nlines)
            g.upload(gs.GRID, grid0)
            g.upload(gs.PREDICT, grid0)
            g.step(900.0, lt, rt, nsteps=4)
            outs[k2t] = (g.download(gs.GRID), g.download(gs.PREDICT), g.status)
        finally:
            gs.default_context().set_option("k2t", 1)
    og, op = grid0.copy(), grid0.copy()
    O.batch1d_step(0, s["leftX"], s["rangeX"], s["rightX"], O.Spline.m1Hermite3(pts), 1.0, og, op, lt, rt, 900.0, nsteps=4)
    for k2t in (1, 0):
        assert_bit_equal(outs[k2t][0], og, f"k2t={k2t} n={n} grid")
        assert_bit_equal(outs[k2t][1], op, f"k2t={k2t} n={n} predict")
    assert outs[1][2] == outs[0][2], "K2T and K2 raise the same status flags"


def test_k2t_nan_is_flagged(gs):
    """A NaN in predict makes the reference's tree lookup throw; K2T raises GS_FLAG_SPLINE_UNDEFINED like K2."""
    nlines, n = 40, 64
    s = cases.batch_setup(nlines, n)
    pts = [(-20.0 + 10.0 * i, 1e-6 * (1.0 + 0.1 * cases.README_RHO[i])) for i in range(11)]
    g = gs.BatchedOneDGrid(s["leftX"], s["rangeX"], s["rightX"], gs.Spline.m1Hermite3(pts), 1.0, nlines=nlines)
    f = np.full((nlines, n), 5.0)
    g.upload(gs.GRID, f)
    f[17, 33] = np.nan
    g.upload(gs.PREDICT, f)
    g.step(900.0, np.zeros(nlines), np.zeros(nlines), nsteps=1)
    assert g.status & gs.FLAG_SPLINE_UNDEFINED


@pytest.mark.parametrize("direction", [0, 1])
def test_per_node_spline_pairs(O, gs, direction):
    """passion.iteration overload 1: conductivities(z)(0) != conductivities(z)(1)."""
    kwc = dict(nlines=37, n=40, steps=3, spline="const", random_grid=True, per_node=True, direction=direction)
    got, want = cases.run_batch(gs, **kwc), cases.run_batch(O, oracle=True, **kwc)  # all-const pairs: K2h
    for k in ("grid", "predict"):
        assert_bit_equal(got[k], want[k], "const pairs " + k)
    kw = dict(nlines=5, n=40, steps=3, spline="hermite", random_grid=True, per_node=True, direction=direction)
    got, want = cases.run_batch(gs, **kw), cases.run_batch(O, oracle=True, **kw)
    for k in ("grid", "predict"):
        assert_bit_equal(got[k], want[k], k)


def test_layouts_round_trip(gs):
    r = np.random.default_rng(1)
    nlines, n = 77, 19
    g = gs.BatchedOneDGrid(0.0, 0.1 * np.arange(1, n + 1), 0.1 * (n + 1), gs.Spline.const(1e-6), 1.0, nlines=nlines)
    a = r.uniform(size=(nlines, n))
    g.upload(gs.GRID, a)
    assert_bit_equal(g.download(gs.GRID), a)
    assert_bit_equal(g.download(gs.GRID, layout=gs.INTERLEAVED), a.T)
    g.upload(gs.PREDICT, np.ascontiguousarray(a.T), layout=gs.INTERLEAVED)
    assert_bit_equal(g.download(gs.PREDICT), a)
    # reference init: grid = predict = 4.0 (A/OneDGrid.scala:20-21)
    g2 = gs.BatchedOneDGrid(0.0, [0.1, 0.2], 0.3, gs.Spline.const(1e-6), 1.0, nlines=3)
    assert np.all(g2.download(gs.GRID) == 4.0) and np.all(g2.download(gs.PREDICT) == 4.0)


def test_bad_arguments_are_reported(gs, ctx):
    with pytest.raises(gs.GsError):
        gs.BatchedOneDGrid(0.0, [0.1], 0.2, gs.Spline.const(1.0), 1.0)  # n < 2
    g = gs.BatchedOneDGrid(0.0, [0.1, 0.2], 0.3, gs.Spline.const(1.0), 1.0)
    with pytest.raises(gs.GsError):
        g.step(900.0)  # no BCs set


def test_not_dominant_is_flagged(gs):
    """assertion(b, c, d) of forwardUnit (A/passion/package.scala:279-282): negative conductivity breaks it."""
    g = gs.OneDGrid(0.0, [0.1, 0.2, 0.3, 0.4], 0.5, gs.Spline.const(-1e-3), 1.0)
    g.step(900.0, 1.0, 2.0)
    assert g.status & gs.FLAG_NOT_DOMINANT
    g.clear_status()
    assert g.status == 0


def test_large_batch_properties(gs):
    """Full-width smoke at a size the oracle would take long for: 65536 x 256.  Properties: every line with the
    same BCs and initial data evolves identically (lines are independent), and the solution stays within the
    discrete maximum principle bounds."""
    nlines, n = 65536, 256
    r = np.random.default_rng(9)
    base = r.uniform(0.0, 30.0, n)
    grid0 = np.tile(base, (nlines, 1))
    g = gs.BatchedOneDGrid(0.0, 0.01 * np.arange(1, n + 1), 0.01 * (n + 1), gs.Spline.const(1e-6), 1.0, nlines=nlines)
    g.upload(gs.GRID, grid0)
    g.step(900.0, 20.0, 4.0, nsteps=5)
    out = g.download(gs.GRID)
    assert np.all(out.view(np.uint64) == out[0].view(np.uint64)[None, :])
    assert out.min() >= 0.0 and out.max() <= 30.0
    assert g.status == 0


@pytest.mark.parametrize("chunk_mb,spline", [(64, "const"), (1, "const"), (1, "hermite")])
def test_step_host_pipeline(O, gs, ctx, chunk_mb, spline):
    """gs_grid1d_step_host = the drop-in OneDGrid.step with the public `grid` array on the host: upload, step,
    download, pipelined over line chunks.  Two calls in a row (state: host grid + device-private predict)."""
    ctx.set_option("host_chunk_mb", chunk_mb)
    try:
        nlines, n = 3000, 64
        s = cases.batch_setup(nlines, n, random_grid=True)
        mk = (lambda m: m.Spline.const(1e-6)) if spline == "const" else (
            lambda m: m.Spline.m1Hermite3([(-20.0 + 10.0 * i, 1e-6 * (1.0 + 0.1 * (i + 1))) for i in range(11)]))
        g = gs.BatchedOneDGrid(s["leftX"], s["rangeX"], s["rightX"], mk(gs), 1.0, nlines=nlines)
        host = s["grid0"].copy()
        og, op = s["grid0"].copy(), np.full((nlines, n), 4.0)
        for _ in range(2):
            g.step_host(900.0, host, s["leftT"], s["rightT"], nsteps=2)
            O.batch1d_step(0, s["leftX"], s["rangeX"], s["rightX"], mk(O), 1.0, og, op, s["leftT"], s["rightT"], 900.0,
                           nsteps=2)
            assert_bit_equal(host, og, "step_host grid")
        assert_bit_equal(g.download(gs.PREDICT), op, "device-private predict")
        assert g.status == 0
    finally:
        ctx.set_option("host_chunk_mb", 64)


@pytest.mark.parametrize("nlines,n,solver", [(70, 300, 2), (33, 1000, 2), (40, 4099, 0), (64, 8192, 0), (5, 129, 2)])
@pytest.mark.parametrize("per_node", [False, True])
def test_long_lines_partitioned(O, gs, nlines, n, solver, per_node):
    """K5 on 1-D batches (config C5's path: constant properties, long lines cut into segments + reduced system).
    AUTO switches to it from 2048 nodes; tolerance 1e-12 relative per field."""
    from helpers import max_rel_err

    kw = dict(nlines=nlines, n=n, steps=3, spline="const", random_grid=True, per_node=per_node)
    got = cases.run_batch(gs, solver=solver, **kw)
    want = cases.run_batch(O, oracle=True, **kw)
    for k in ("grid", "predict"):
        err = max_rel_err(got[k], want[k])
        assert err <= 1e-12, f"{k}: {err:.3e}"
    exact = cases.run_batch(gs, solver=gs.SOLVER_THOMAS, **kw)
    for k in ("grid", "predict"):
        assert_bit_equal(exact[k], want[k], "THOMAS stays bit-exact " + k)


@pytest.mark.parametrize("nlines,n", [(96, 16384), (1000, 4096), (33, 2048)])
def test_long_lines_k5t_variants(O, gs, ctx, nlines, n):
    """K5T on 1-D batches (TMA boxes of 32 lines x 16 nodes, clusters / soft groups when the strips are few) against
    k5_sweep and the oracle."""
    from helpers import max_rel_err

    kw = dict(nlines=nlines, n=n, steps=2, spline="const", random_grid=True)
    want = cases.run_batch(O, oracle=True, **kw)
    defaults = {"k5t": 1, "k5t_soft": 1, "k5_cluster": 1, "k5t_s1": 0}
    try:
        for opts in ({}, {"k5t": 0}, {"k5t_soft": 0}, {"k5_cluster": 0}, {"k5t_s1": 4}):
            for k, v in {**defaults, "k5t": 2, **opts}.items():   # k5t = 2: K5T also below its size threshold
                ctx.set_option(k, v)
            got = cases.run_batch(gs, solver=gs.SOLVER_PARTITIONED, **kw)
            for k in ("grid", "predict"):
                err = max_rel_err(got[k], want[k])
                assert err <= 1e-12, f"{opts} {k}: {err:.3e}"
    finally:
        for k, v in defaults.items():
            ctx.set_option(k, v)


def test_long_line_kernels_are_deterministic(gs):
    """Same idea as tests/test_gpu_2d.py::test_ring_kernels_are_deterministic for the 1-D K5T paths: soft groups (few
    strips), clusters and plain CTAs must give the same bits on every run."""
    gs.default_context().set_option("k5t", 2)   # K5T also below its size threshold
    try:
        for nlines, n in ((64, 16384), (600, 8192), (33, 2048)):
            kw = dict(nlines=nlines, n=n, steps=3, spline="const", random_grid=True)
            first = cases.run_batch(gs, solver=gs.SOLVER_PARTITIONED, **kw)
            for _ in range(4):
                again = cases.run_batch(gs, solver=gs.SOLVER_PARTITIONED, **kw)
                for k in ("grid", "predict"):
                    assert_bit_equal(again[k], first[k], f"{nlines} x {n} {k}")
    finally:
        gs.default_context().set_option("k5t", 1)
