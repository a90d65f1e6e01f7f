"""The reference's Coefficients classes as the host mirror flattens them (A/TwoDGrid.scala:782-1137), literally:
VarXCoef indexes by y + 1 like VarYCoef (:862-874); PatchXCoef by y + 1, PatchYCoef.apply by x + 1 (:1056);
Patched*Coef = the patch where it `contains` the node.  The oracle gets id maps written out by hand from the Scala."""
import numpy as np
import pytest

import cases
from helpers import README_RHO, assert_bit_equal

pytestmark = pytest.mark.gpu


def _pool(mk_const, mk_herm, n):
    pts = [(-20.0 + 10.0 * i, 6.0e-7 * (1.0 + 0.1 * README_RHO[i])) for i in range(11)]
    therm = [mk_const(1.5 + 0.01 * i) for i in range(n)]
    cap = [mk_const(2.2e6 + 1.0e4 * i) for i in range(n)]
    temp = [mk_herm(pts) if i % 3 == 0 else mk_const(6.0e-7 + 2.0e-8 * i) for i in range(n)]
    return therm, cap, temp


def _run(O, gs, p, gs_coefs, omaps, opool, steps=2):
    g = p.build_gs(gs)
    g.set_coefs(gs_coefs)
    o = p.build_oracle(O)
    o.set_splines(opool)
    for sel, (mode, ids) in omaps.items():
        o.set_map(sel, mode, ids)
    g.iteration(900.0, nsteps=steps)
    for _ in range(steps):
        o.iteration(900.0)
    got = cases.snapshot(g)
    for k, v in cases.snapshot(o).items():
        assert_bit_equal(got[k], v, k)


def test_varxcoef_is_indexed_by_row_like_the_reference(O, gs):
    cols, rows = 24, 18            # cols + 1 entries per array (one per column interval), rows + 1 of them are used
    p = cases.bc_problem("TTFT", xdir=1, cols=cols, rows=rows)
    n = cols + 1
    gt, gc, ga = _pool(lambda v: gs.AlwaysDefinedSpline(gs.Spline.const(v)),
                       lambda pts: gs.AlwaysDefinedSpline(gs.Spline.m1Hermite3(pts)), n)
    ot, oc, oa = _pool(lambda v: O.AlwaysDefinedSpline(O.Spline.const(v)),
                       lambda pts: O.AlwaysDefinedSpline(O.Spline.m1Hermite3(pts)), n)
    opool = ot + oc + oa
    row_ids = lambda base: [base + (y + 1) for y in range(-1, rows)]          # member(x, y) = array(y + 1)
    omaps = {0: (1, row_ids(2 * n)), 1: (1, row_ids(0)), 2: (1, row_ids(n)), 3: (1, row_ids(2 * n))}
    _run(O, gs, p, gs.VarXCoef(gt, gc, ga), omaps, opool)
    with pytest.raises(IndexError):      # rows > cols: ArrayIndexOutOfBoundsException in the reference
        gs.VarXCoef(gt[:5], gc[:5], ga[:5]).maps(4, 9)


@pytest.mark.parametrize("kind", ["X", "Y"])
def test_patched_coefs(O, gs, kind):
    cols, rows = 22, 22
    p = cases.bc_problem("TTTT", xdir=0, cols=cols, rows=rows)
    n = cols + 1
    mkc = lambda m: (lambda v: m.AlwaysDefinedSpline(m.Spline.const(v)))
    mkh = lambda m: (lambda pts: m.AlwaysDefinedSpline(m.Spline.m1Hermite3(pts)))
    gt, gc, ga = _pool(mkc(gs), mkh(gs), n)
    ot, oc, oa = _pool(mkc(O), mkh(O), n)
    # the patch's own arrays: shifted values so that it is visible
    gpt, gpc, gpa = (list(reversed(a)) for a in (gt, gc, ga))
    rx, ry = (3, 15), (-1, 9)
    if kind == "X":
        coefs = gs.PatchedXCoef(gs.VarXCoef(gt, gc, ga), gs.PatchXCoef(gpt, gpc, gpa, rx, ry))
    else:
        coefs = gs.PatchedYCoef(gs.VarYCoef(gt, gc, ga), gs.PatchYCoef(gpt, gpc, gpa, rx, ry))
    opool = ot + oc + oa                       # ids: therm 0.., cap n.., temp 2n..
    rev = lambda i: n - 1 - i                  # index into the un-reversed pool of the patch's entry i
    dense = {s: np.empty((rows + 1, cols + 1), dtype=np.int32) for s in range(4)}
    for y in range(-1, rows):
        for x in range(-1, cols):
            inside = rx[0] <= x < rx[1] and ry[0] <= y < ry[1]
            if inside:
                ia = rev((x if kind == "Y" else y) + 1)      # PatchYCoef.apply: tempCond(x + 1); PatchXCoef: y + 1
                it = ic = im = rev(y + 1)
            else:
                ia = it = ic = im = y + 1                     # VarXCoef and VarYCoef both: y + 1
            dense[0][y + 1, x + 1] = 2 * n + ia
            dense[1][y + 1, x + 1] = it
            dense[2][y + 1, x + 1] = n + ic
            dense[3][y + 1, x + 1] = 2 * n + im
    omaps = {s: (3, dense[s].ravel().tolist()) for s in range(4)}
    _run(O, gs, p, coefs, omaps, opool)
