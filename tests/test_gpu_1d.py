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


@pytest.mark.parametrize("nlines,n", [(1, 2), (3, 3), (33, 5), (1000, 256), (4099, 33), (257, 1024)])
@pytest.mark.parametrize("spline", ["const", "hermite", "lines"])
def test_batch_vs_oracle(O, gs, nlines, n, spline):
    kw = dict(nlines=nlines, n=n, steps=3, spline=spline, random_grid=True)
    got, want = cases.run_batch(gs, **kw), cases.run_batch(O, oracle=True, **kw)
    for k in ("grid", "predict"):
        assert_bit_equal(got[k], want[k], f"{spline} {nlines}x{n} {k}")


@pytest.mark.parametrize("direction", [0, 1])
def test_per_node_spline_pairs(O, gs, direction):
    """passion.iteration overload 1: conductivities(z)(0) != conductivities(z)(1)."""
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
