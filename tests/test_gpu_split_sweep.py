"""Split y sweep of a row-sharded grid (gs_grid2d_yiter_update_begin / _end), checked on ONE device: P local grids
stand in for P ranks, the two all-gathers are torch.cat.  The multi-process driver around the same calls is
gridsplines_b200.sharded2d.SpikeShardedTwoDGrid (tests/test_sharding_gloo.py, tools/run_sharded2d.py)."""
import numpy as np
import pytest

import cases
from helpers import check_tol, max_rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-12
TOL_TABLES = 2e-11
UPP, LOW, RIGHT, LEFT = 0, 1, 2, 3


def _slab(p, r0, rl):
    """The problem restricted to rows [r0, r0 + rl): neighbouring y coordinates kept as the slab's end coordinates,
    left / right bounds sliced, upper / lower bounds kept (only the first / last slab uses them)."""
    b = {}
    for side, (kind, rep, v) in p.bounds.items():
        v = np.atleast_1d(v)
        b[side] = (kind, rep, v[r0:r0 + rl] if side in (RIGHT, LEFT) and v.size > 1 else v)
    q = cases.Problem2D(p.xv, p.yv[r0:r0 + rl + 2], p.xdir, p.ydir, b, coefs=p.coefs,
                        init=(p.grid0[r0:r0 + rl].copy(), p.predict0[r0:r0 + rl].copy()))
    return q


def _split_step(gs, torch, locals_, time):
    P = len(locals_)
    for g in locals_:
        g.xIter(time)
    edges = []
    for g in locals_:
        pr = torch.from_numpy(g.grid.predict).cuda()
        edges.append((pr[0].contiguous(), pr[-1].contiguous()))
    exported = [torch.empty((6, g.cols), dtype=torch.float64, device="cuda") for g in locals_]
    torch.cuda.synchronize()
    for r, g in enumerate(locals_):
        prev = edges[r - 1][1].data_ptr() if r > 0 else None
        nxt = edges[r + 1][0].data_ptr() if r + 1 < P else None
        g.yIterUpdateBegin(time, prev, nxt, exported[r].data_ptr())
        g.ctx.synchronize()
    gathered = torch.stack(exported).contiguous()
    torch.cuda.synchronize()
    for r, g in enumerate(locals_):
        g.yIterUpdateEnd(time, gathered.data_ptr(), P, r)
        g.ctx.synchronize()


@pytest.mark.parametrize("P", [1, 2, 4])
@pytest.mark.parametrize("coefs,combo,ydir,rows,cols", [
    ("const", "TTTT", 0, 512, 96), ("hermite", "TTTT", 0, 512, 96), ("hermite", "FTFT", 1, 512, 96),
    ("dense", "TTTF", 0, 512, 96),
    ("hermite", "TTTT", 0, 4096, 40),      # two strips of columns: every strip goes to a cluster of CTAs
])
def test_split_y_sweep_equals_whole_grid(O, gs, ctx, coefs, combo, ydir, rows, cols, P):
    torch = pytest.importorskip("torch")
    ctx.set_option("k5", 0)   # literal constants would otherwise take K5 on the whole grid
    try:
        p = cases.bc_problem(combo, xdir=1, ydir=ydir, cols=cols, rows=rows, coefs=coefs)
        r = np.random.default_rng(3)
        p.grid0 = r.uniform(-10.0, 60.0, p.grid0.shape)
        p.predict0 = p.grid0 + r.uniform(-3.0, 3.0, p.grid0.shape)
        o = p.build_oracle(O, ascending=False)
        rl = rows // P
        locals_ = []
        for k in range(P):
            g = _slab(p, k * rl, rl).build_gs(gs)
            g.set_solver(gs.SOLVER_PARTITIONED)
            locals_.append(g)
        steps = 3
        for _ in range(steps):
            _split_step(gs, torch, locals_, 900.0)
            o.iteration(900.0)
        O.set_ascending_fold(True)
        got_grid = np.concatenate([g.grid.grid for g in locals_])
        got_pred = np.concatenate([g.grid.predict for g in locals_])
        tol = TOL if coefs == "const" else TOL_TABLES     # fixed constants (tests/test_gpu_2d.py explains TOL_TABLES)
        check_tol(got_grid, o.grid, f"split sweep {coefs} P={P} grid", tol)
        check_tol(got_pred, o.predict, f"split sweep {coefs} P={P} predict", tol)
        for g in locals_:
            assert g.status == 0
    finally:
        ctx.set_option("k5", 1)


def test_split_needs_matching_begin(gs):
    p = cases.bc_problem("TTTT", cols=64, rows=256, coefs="hermite")
    g = p.build_gs(gs)
    with pytest.raises(gs.GsError):
        g.yIterUpdateEnd(900.0, 1, 1, 0)
    g.set_solver(gs.SOLVER_THOMAS)
    with pytest.raises(gs.GsError):
        g.yIterUpdateBegin(900.0, None, None, 1)
