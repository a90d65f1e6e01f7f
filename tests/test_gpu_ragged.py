"""Ragged batch of lines in one flattened array (gs_ragged_*): the fixed version of the reference's experimental
passion.iteration overload 3 (A/passion/package.scala:161-227).  Oracle: overload 1 (:52-95) applied to every line's
gathered nodes -- the semantics overload 3 was written for."""
import numpy as np
import pytest

from helpers import README_RHO, assert_bit_equal

pytestmark = pytest.mark.gpu


def _problem(rows=23, cols=31, seed=5):
    """A staircase domain inside a rows x cols matrix: row r spans a different range of columns (row lines), and
    a few column lines run through its interior (stride `cols` between nodes)."""
    r = np.random.default_rng(seed)
    lines, coords = [], []
    for row in range(rows):
        c0, c1 = row % 5, cols - (row * 3) % 7
        lines.append(np.arange(c0, c1) + row * cols)
        coords.append(0.05 + 0.01 * np.arange(c0 - 1, c1 + 1))          # low, nodes..., upp
    for col in (6, 7, 19):
        lines.append(col + cols * np.arange(2, rows - 3))
        coords.append(0.02 * np.arange(1, rows - 2))
    grid = r.uniform(0.0, 30.0, rows * cols)
    pred = grid + r.uniform(-2.0, 2.0, rows * cols)
    upper, lower = r.uniform(10.0, 30.0, len(lines)), r.uniform(0.0, 10.0, len(lines))
    return rows * cols, lines, coords, grid, pred, upper, lower


@pytest.mark.parametrize("mode", ["single", "per_node"])
def test_ragged_lines_match_overload_1_per_line(O, gs, mode):
    size, lines, coords, grid, pred, upper, lower = _problem()
    pts = [(-20.0 + 10.0 * i, 1e-6 * (1.0 + 0.1 * README_RHO[i])) for i in range(11)]
    mk = {"a": lambda m: m.Spline.m1Hermite3(pts), "b": lambda m: m.Spline.const(2e-6)}
    predefs = [O.predef(l % 2, c[0], c[1:-1], c[-1], 1.0).reshape(-1, 2) for l, c in enumerate(coords)]
    oa, ob = O.AlwaysDefinedSpline(mk["a"](O)), O.AlwaysDefinedSpline(mk["b"](O))
    if mode == "single":
        g = gs.RaggedLines(size, lines, predefs, mk["a"](gs))
        oconds = [[(oa, oa)] * len(l) for l in lines]
    else:
        ga, gb = gs.AlwaysDefinedSpline(mk["a"](gs)), gs.AlwaysDefinedSpline(mk["b"](gs))
        gconds = [[((ga, gb) if (k + li) % 2 else (gb, ga)) for k in range(len(l))] for li, l in enumerate(lines)]
        oconds = [[((oa, ob) if (k + li) % 2 else (ob, oa)) for k in range(len(l))] for li, l in enumerate(lines)]
        g = gs.RaggedLines(size, lines, predefs, gconds)
    g.grid, g.predicted = grid, pred
    marker = np.full(size, -777.0)
    g.result = marker
    g.iteration(900.0, upper, lower)
    got = g.result
    want = marker.copy()
    # row lines first, then column lines: a later line overwrites `result` where lines cross, like a serial walk
    for li, l in enumerate(lines):
        want[l] = O.passion_iteration(900.0, upper[li], grid[l], lower[li], pred[l], predefs[li], oconds[li])
    cross = np.zeros(size, bool)
    seen = np.zeros(size, int)
    for l in lines:
        seen[l] += 1
    cross = seen > 1
    assert_bit_equal(got[~cross], want[~cross], "ragged result")
    untouched = seen == 0
    assert np.all(got[untouched] == -777.0)
    # where a row line and a column line cross, either one's value may land (the lines run concurrently)
    for i in np.nonzero(cross)[0]:
        cands = [O.passion_iteration(900.0, upper[li], grid[l], lower[li], pred[l], predefs[li], oconds[li])[list(l).index(i)]
                 for li, l in enumerate(lines) if i in l]
        assert got[i] in cands
    assert g.status == 0


def test_ragged_rejects_bad_input(gs):
    with pytest.raises(gs.GsError):
        gs.RaggedLines(10, [np.array([3])], [np.ones((1, 2))], gs.Spline.const(1e-6))          # one-node line
    with pytest.raises(gs.GsError):
        gs.RaggedLines(10, [np.array([3, 12])], [np.ones((2, 2))], gs.Spline.const(1e-6))      # index outside
