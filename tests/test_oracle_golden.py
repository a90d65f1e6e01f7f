"""Pins the CPU oracle (oracle/gs_oracle.c) before anything is compared with it:
 - the six published known answers of the reference (README.md:77-110), bit-exact;
 - the physical assertion of AT/GridTest.scala:99 (k_est within 0.1 of 1.5);
 - the inequalities of AT/TwoDGridUnit.scala:63-94;
 - the restatement KATs R1-R5 of SURVEY App. C (regression values of the literal restatement);
 - the committed fixtures tests/golden/*.npz (generated from this oracle; they pin drift).
No GPU needed."""
import math
import os

import numpy as np
import pytest

import cases
from helpers import README_RHO, README_T, assert_bit_equal

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def test_readme_known_answers(O):
    s = O.Spline.m1Hermite3(list(zip(README_T, README_RHO)))
    assert s(110.0) == 0.826
    assert s(150.0) == 2.547
    assert s(135.0) == 1.718125
    assert s.der(185.0) == 0.1238875000000001
    assert s.antider(185.0) == 27.242786458333335
    assert s.area(150.0, 185.0) == 139.148203125
    assert s.area_ascending(150.0, 185.0) == 139.148203125


def test_host_spline_construction_matches_oracle(O, gs):
    """gridsplines_b200.Spline factories (host logic) build the same tables as the oracle's."""
    pts = list(zip(README_T, README_RHO))
    for kind in ("m1Hermite3", "fHermite3", "lines"):
        o, g = getattr(O.Spline, kind)(pts), getattr(gs.Spline, kind)(pts)
        assert o.size == g.size
        for i, (k, lo, hi, x0, c) in enumerate(o.pieces()):
            assert k == g.kind and lo == g.knots[i] and hi == g.knots[i + 1] and x0 == g.x0[i], (kind, i)
            assert c == list(g.coefs[i]), (kind, i)
        assert o.lowerBound == g.lowerBound and o.upperBound == g.upperBound
        assert o.trueUpperBound == g.trueUpperBound
    # shuffled input is sorted first (Spline.apply sorts by x, P/Spline.scala:395-400)
    rs = np.random.default_rng(3).permutation(len(pts))
    g2 = gs.Spline.m1Hermite3([pts[i] for i in rs])
    assert np.array_equal(g2.coefs, gs.Spline.m1Hermite3(pts).coefs)


def test_kat_r2_c1(O):
    out = cases.run_c1(O, oracle=True)
    g = out["grid"]
    assert g[0] == 19.163547799904077 and g[99] == 6.303223056805375 and g[199] == 4.014707461585766
    assert g.sum() == 1489.2906440389272


def _small(O, coefs="unit"):
    xv, yv = cases.small2d_dims()
    b = {s: (0, 0, [1.0]) for s in range(4)}
    p = cases.Problem2D(xv, yv, 1, 0, b, coefs=coefs, init=(np.ones((10, 8)), np.ones((10, 8))))
    return p


def test_kat_r3_xsweep(O):
    p = _small(O)
    p.bounds[3] = (0, 0, [4.0])
    g = p.build_oracle(O, ascending=False)
    g.xIter(900.0)
    g.update()
    row0 = [3.10981041911049, 2.5775720576159933, 2.1989042489843205, 1.9055689452327573, 1.6664820491786831,
            1.464861323850035, 1.2906054403084588, 1.137154150062293]
    assert list(g.grid[0]) == row0 and list(g.grid[9]) == row0
    # the survey's averages came from a pairwise sum; the reference sums serially (A/TwoDGrid.scala:353-379)
    assert g.edgeAverage(O.LEFT) == pytest.approx(3.10981041911049, rel=1e-14)
    assert g.edgeAverage(O.RIGHT) == pytest.approx(1.137154150062293, rel=1e-14)
    assert g.edgeAverage(O.LEFT) > g.edgeAverage(O.RIGHT)  # AT/TwoDGridUnit.scala:63-73


def test_kat_r4_ysweep_alone(O):
    p = _small(O)
    p.bounds[0] = (0, 0, [4.0])
    g = p.build_oracle(O, ascending=False)
    g.result[:] = 0.0
    g.yIter(900.0)
    g.update()
    assert g.edgeAverage(O.UPP) == pytest.approx(3.711285138039772, rel=1e-14)
    assert g.edgeAverage(O.LOW) == pytest.approx(1.261272050346904, rel=1e-14)
    assert g.edgeAverage(O.UPP) > g.edgeAverage(O.LOW)  # AT/TwoDGridUnit.scala:75-85


def test_kat_r5_full_step(O):
    p = _small(O)
    p.bounds[3] = (0, 0, [4.0])
    g = p.build_oracle(O, ascending=False)
    g.iteration(900.0)
    assert g.avValue == pytest.approx(1.0110804508847722, rel=1e-14) and g.grid[0, 0] == 1.0115798259742925
    assert g.avValue > 1.0  # AT/TwoDGridUnit.scala:87-94


def test_kat_r1_gridtest(O):
    """AT/GridTest.scala: regress wall temperature on ln t; power / (4 pi 92 slope) within 0.1 of 1.5."""
    p = cases.gridtest_problem()
    assert (p.cols, p.rows) == (137, 45)
    g = p.build_oracle(O, ascending=False)
    for side in range(4):
        g.noHeatFlow(side)
    g.fill(7.0)
    heat = 4500.0 / (92.0 * 2.0 * math.pi)
    pts = []
    cur = 0
    for _ in range(195):
        g.set_bound(3, 1, 1, [heat])
        pts.append((math.log(cur) if cur > 0 else -math.inf, g.grid[0, 0]))
        for side in (2, 0, 1):
            g.noHeatFlow(side)
        g.iteration(900.0)
        cur += 900
    assert g.grid[0, 0] == 54.44328715390563 and g.grid[0, 1] == 53.94901954479779
    assert g.grid[22, 0] == 54.44368106189198
    # SimpleRegression over points with time > 20000 (AT/GridTest.scala:91-97)
    xs = np.array([x for (x, y), t in zip(pts, range(0, 195 * 900, 900)) if t > 20000])
    ys = np.array([y for (x, y), t in zip(pts, range(0, 195 * 900, 900)) if t > 20000])
    slope = np.polyfit(xs, ys, 1)[0]
    k_est = 4500.0 / (4.0 * math.pi * 92.0 * slope)
    assert abs(k_est - 1.5) < 0.1, k_est


def test_literal_vs_ascending_fold(O):
    """The tree-order fold and the ascending fold agree bit-exactly for <= 2 overlapped pieces and to
    1e-12 beyond."""
    s = O.Spline.m1Hermite3(list(zip(README_T, README_RHO)))
    r = np.random.default_rng(5)
    worst = 0.0
    for _ in range(2000):
        a, b = sorted(r.uniform(100.0, 200.0, 2))
        lit, asc = s.area(a, b), s.area_ascending(a, b)
        if b - a < 10.0:
            assert lit == asc
        worst = max(worst, abs(lit - asc) / abs(lit))
    assert worst < 1e-12


@pytest.mark.parametrize("name", sorted(f[:-4] for f in os.listdir(GOLDEN) if f.endswith(".npz")) if os.path.isdir(GOLDEN) else [])
def test_oracle_reproduces_golden(O, name):
    from golden.make_golden import CASES

    want = np.load(os.path.join(GOLDEN, name + ".npz"))
    got = CASES[name](O)
    assert sorted(want.files) == sorted(got)
    for k in want.files:
        assert_bit_equal(got[k], want[k], f"{name}.{k}")
