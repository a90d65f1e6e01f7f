"""Named parity cases, each built identically for the oracle (`O`) and for gridsplines_b200 (`gs`).
`run_*` returns a dict name -> ndarray; tests compare dicts (bit patterns) with each other and with the
committed fixtures in tests/golden/ (generated from the oracle by tests/golden/make_golden.py)."""
import numpy as np

from helpers import README_RHO, README_T, hermite_coefs, uneven_coefs

SEED = 20261019


# ------------------------------------------------------------------------------ 1-D
def c1_setup():
    """SURVEY 8(d) C1 / KAT R2: OneDGrid[Radial], 200 nodes, const 1e-6, step(900, 20, 4)."""
    rng = np.array([0.05 + 0.01 * (i + 1) for i in range(200)])
    return dict(leftX=0.05, rangeX=rng, rightX=rng[-1] + 0.01, sigma=1.0, time=900.0, leftT=20.0, rightT=4.0)


def run_c1(mod, steps=1000, oracle=False):
    s = c1_setup()
    if oracle:
        g = mod.OneDGrid(s["leftX"], s["rangeX"], s["rightX"], mod.Spline.const(1e-6), s["sigma"], direction=mod.RADIAL)
        for _ in range(steps):
            g.step(s["time"], s["leftT"], s["rightT"])
        return {"grid": g.grid.copy(), "predict": g.predict.copy()}
    g = mod.OneDGrid(s["leftX"], s["rangeX"], s["rightX"], mod.Spline.const(1e-6), s["sigma"], direction=mod.RADIAL)
    g.step(s["time"], s["leftT"], s["rightT"], nsteps=steps)
    return {"grid": g.grid, "predict": g.predict}


def batch_setup(nlines, n, seed=SEED, random_grid=False):
    r = np.random.default_rng(seed)
    rangeX = np.array([0.01 * (i + 1) for i in range(n)])
    d = dict(leftX=0.0, rangeX=rangeX, rightX=0.01 * (n + 1), sigma=1.0, time=900.0,
             leftT=r.uniform(10.0, 30.0, nlines), rightT=r.uniform(0.0, 10.0, nlines))
    d["grid0"] = r.uniform(0.0, 30.0, (nlines, n)) if random_grid else None
    return d


def run_batch(mod, nlines, n, steps, spline="const", random_grid=False, oracle=False, direction=0, per_node=False,
              solver=None):
    s = batch_setup(nlines, n, random_grid=random_grid)
    if spline == "const":
        mk = lambda m: m.Spline.const(1e-6)
    elif spline == "hermite":
        pts = [(-20.0 + 10.0 * i, 1e-6 * (1.0 + 0.1 * README_RHO[i])) for i in range(11)]
        mk = lambda m: m.Spline.m1Hermite3(pts)
    elif spline == "lines":
        pts = [(-20.0 + 25.0 * i, 1e-6 * (1.0 + 0.3 * i)) for i in range(5)]
        mk = lambda m: m.Spline.lines(pts)
    else:
        raise ValueError(spline)
    if oracle:
        O = mod
        grid = np.full((nlines, n), 4.0) if s["grid0"] is None else s["grid0"].copy()
        predict = np.full((nlines, n), 4.0)
        if per_node:
            # overload 1 with distinct left/right splines per node: one OneDGrid per line
            sp = [mk(O), O.Spline.const(2e-6)]
            ads = [O.AlwaysDefinedSpline(x) for x in sp]
            conds = [(ads[i % 2], ads[(i + 1) % 2]) for i in range(n)]
            for l in range(nlines):
                g = O.OneDGrid(s["leftX"], s["rangeX"], s["rightX"], conds, s["sigma"], direction=direction)
                g.grid[:] = grid[l]
                for _ in range(steps):
                    g.step(s["time"], s["leftT"][l], s["rightT"][l])
                grid[l], predict[l] = g.grid, g.predict
        else:
            O.batch1d_step(direction, s["leftX"], s["rangeX"], s["rightX"], mk(O), s["sigma"], grid, predict,
                           s["leftT"], s["rightT"], s["time"], nsteps=steps, nthreads=4)
        return {"grid": grid, "predict": predict}
    gs = mod
    if per_node:
        ads = [gs.AlwaysDefinedSpline(mk(gs)), gs.AlwaysDefinedSpline(gs.Spline.const(2e-6))]
        cond = [(ads[i % 2], ads[(i + 1) % 2]) for i in range(n)]
    else:
        cond = mk(gs)
    g = gs.BatchedOneDGrid(s["leftX"], s["rangeX"], s["rightX"], cond, s["sigma"], direction=direction, nlines=nlines)
    if solver is not None:
        g.set_solver(solver)
    if s["grid0"] is not None:
        g.upload(gs.GRID, s["grid0"])
    g.step(s["time"], s["leftT"], s["rightT"], nsteps=steps)
    out = {"grid": g.download(gs.GRID), "predict": g.download(gs.PREDICT)}
    assert g.status == 0, g.status
    return out


# ------------------------------------------------------------------------------ 2-D
def small2d_dims():
    """KAT R3-R5: XDim[Radial](1, +1, 10) -> cols 2..9; YDim[Ortho](1, +1, 12) -> rows 2..11."""
    return np.arange(1.0, 11.0), np.arange(1.0, 13.0)


class Problem2D:
    """A 2-D problem description that can be instantiated on either implementation."""

    def __init__(self, xvalues, yvalues, xdir, ydir, bounds, coefs="const", init=None, seed=SEED):
        self.xv, self.yv = np.asarray(xvalues, float), np.asarray(yvalues, float)
        self.xdir, self.ydir = xdir, ydir
        self.bounds = bounds  # side -> (kind, rep, values)
        self.coefs = coefs
        self.cols, self.rows = len(self.xv) - 2, len(self.yv) - 2
        r = np.random.default_rng(seed)
        if init is None:
            self.grid0 = 7.0 + r.uniform(0.0, 1.0, (self.rows, self.cols))
            self.grid0[: self.rows // 3, : self.cols // 3] = 7.0  # keeps the t_lo == t_hi branch covered
            self.predict0 = self.grid0.copy()
        else:
            self.grid0, self.predict0 = init

    # ---- oracle
    def build_oracle(self, O, ascending=True):
        O.set_ascending_fold(ascending)
        g = O.TwoDGrid(self.xv, self.yv, self.xdir, self.ydir)
        self._okeep = []
        if self.coefs == "const":
            ads = [O.AlwaysDefinedSpline(O.Spline.const(v)) for v in (1.5, 2.2e6, 1.5 / 2.2e6)]
            g.set_constant_coef(*ads)
        elif self.coefs == "unit":
            ads = [O.AlwaysDefinedSpline(O.Spline.const(1.0)) for _ in range(3)]
            g.set_constant_coef(*ads)
        elif self.coefs == "hermite":
            import gridsplines_b200 as gs_

            (ok, oc, oa), _ = hermite_coefs(O, gs_)
            ads = [O.AlwaysDefinedSpline(s) for s in (ok, oc, oa)]
            g.set_constant_coef(*ads)
        elif self.coefs in ("lines_uneven", "hermite_uneven"):
            import gridsplines_b200 as gs_

            (ok, oc, oa), _ = uneven_coefs(O, gs_, "lines" if self.coefs == "lines_uneven" else "m1Hermite3")
            ads = [O.AlwaysDefinedSpline(s) for s in (ok, oc, oa)]
            g.set_constant_coef(*ads)
        elif self.coefs in ("per_row", "per_col", "dense"):
            pool, maps = self._var_pool(lambda v: O.AlwaysDefinedSpline(O.Spline.const(v)),
                                        lambda pts: O.AlwaysDefinedSpline(O.Spline.m1Hermite3(pts)))
            g.set_splines(pool)
            for sel, (mode, ids) in maps.items():
                g.set_map(sel, mode, ids)
            ads = pool
        self._okeep = ads
        for side, (kind, rep, values) in self.bounds.items():
            g.set_bound(side, kind, rep, values)
        g.grid[:] = self.grid0
        g.predict[:] = self.predict0
        return g

    def _var_pool(self, const, herm):
        """4 distinct tempConductivity splines + therm/cap, assigned by row / column / (row, col)."""
        pts = [(-20.0 + 10.0 * i, 6.0e-7 * (1.0 + 0.1 * README_RHO[i])) for i in range(11)]
        pool = [const(1.5), const(2.2e6), const(6.8e-7), const(9.1e-7), herm(pts), const(4.0e-7)]
        tc = [2, 3, 4, 5]
        rows, cols = self.rows, self.cols
        if self.coefs == "per_row":
            mode, ids = 1, [tc[(i * 7) % 4] for i in range(rows + 1)]
        elif self.coefs == "per_col":
            mode, ids = 2, [tc[(i * 5) % 4] for i in range(cols + 1)]
        else:
            mode = 3
            ids = [tc[(r * 3 + c * 5) % 4] for r in range(rows + 1) for c in range(cols + 1)]
        n = len(ids)
        maps = {0: (mode, ids), 1: (mode, [0] * n), 2: (mode, [1] * n), 3: (mode, ids)}
        return pool, maps

    # ---- gridsplines_b200
    def build_gs(self, gs):
        x = gs.XDim(self.xv[0], self.xv[1:-1], self.xv[-1], self.xdir)
        y = gs.YDim(self.yv[0], self.yv[1:-1], self.yv[-1], self.ydir)
        spec = {s: (rep, kind) for s, (kind, rep, _) in self.bounds.items()}
        dflt = (gs.One, gs.Temp)
        g = gs.TwoDGrid(x, y, upp=spec.get(gs.UPP, dflt), low=spec.get(gs.LOW, dflt), right=spec.get(gs.RIGHT, dflt),
                        left=spec.get(gs.LEFT, dflt))
        if self.coefs == "const":
            g.set_coefs(gs.ConstantCoef(gs.Spline.const(1.5), gs.Spline.const(2.2e6), gs.Spline.const(1.5 / 2.2e6)))
        elif self.coefs == "unit":
            g.set_coefs(gs.ConstantCoef(gs.Spline.const(1.0), gs.Spline.const(1.0), gs.Spline.const(1.0)))
        elif self.coefs == "hermite":
            from oracle import oracle as O_

            _, (gk, gc, ga) = hermite_coefs(O_, gs)
            g.set_coefs(gs.ConstantCoef(gk, gc, ga))
        elif self.coefs in ("lines_uneven", "hermite_uneven"):
            from oracle import oracle as O_

            _, (gk, gc, ga) = uneven_coefs(O_, gs, "lines" if self.coefs == "lines_uneven" else "m1Hermite3")
            g.set_coefs(gs.ConstantCoef(gk, gc, ga))
        else:
            pool, maps = self._var_pool(lambda v: gs.AlwaysDefinedSpline(gs.Spline.const(v)),
                                        lambda pts: gs.AlwaysDefinedSpline(gs.Spline.m1Hermite3(pts)))
            ids = lambda lst: np.array([pool[i].id for i in lst], dtype=np.int32)
            if self.coefs == "dense":
                g.set_coefs(gs.DenseCoef(*[ids(maps[k][1]) for k in (0, 1, 2, 3)]))
            else:
                cls = gs.VarYCoef if self.coefs == "per_row" else gs.VarXCoefByColumn
                g.set_coefs(cls([pool[i] for i in maps[1][1]], [pool[i] for i in maps[2][1]],
                                [pool[i] for i in maps[3][1]]))
        side_attr = {gs.UPP: "upp", gs.LOW: "low", gs.RIGHT: "right", gs.LEFT: "left"}
        for side, (kind, rep, values) in self.bounds.items():
            getattr(g.bounds, side_attr[side]).update(values)
        g.grid.grid = self.grid0
        g.grid.predict = self.predict0
        return g


def snapshot(g):
    """grid / predict / result of either implementation as arrays."""
    if hasattr(g.grid, "grid"):  # gridsplines_b200.TwoDGrid
        return {"grid": g.grid.grid, "predict": g.grid.predict, "result": g.grid.result}
    return {"grid": g.grid.copy(), "predict": g.predict.copy(), "result": g.result.copy()}


def bc_problem(combo, xdir=0, ydir=0, cols=45, rows=37, coefs="const", dense=True, seed=SEED):
    """combo: 4 letters for (upp, low, right, left), T = Temperature, F = HeatFlow."""
    r = np.random.default_rng(seed + 1)
    xv = 0.05 + 0.01 * np.arange(cols + 2)
    yv = 0.01 * np.arange(rows + 2)
    if xdir == 1:
        xv = 0.05 * 1.02 ** np.arange(cols + 2)
    bounds = {}
    for side, letter in zip((0, 1, 2, 3), combo):
        n = cols if side in (0, 1) else rows
        if letter == "T":
            vals = r.uniform(4.0, 20.0, n) if dense else np.array([r.uniform(4.0, 20.0)])
            bounds[side] = (0, 1 if dense else 0, vals)
        else:
            vals = r.uniform(0.0, 50.0, n) if dense else np.array([r.uniform(0.0, 50.0)])
            bounds[side] = (1, 1 if dense else 0, vals)
    return Problem2D(xv, yv, xdir, ydir, bounds, coefs=coefs, seed=seed)


def gridtest_problem():
    """AT/GridTest.scala:17-66: XDim[Radial] 5e-5 .. 25 x1.1 (137 cols), YDim[Ortho] 0..92 step 2 (45 rows),
    k = 1.5, c = 2.2e6, (Many,Temp) x3 + (Many,Flow) left with q = 4500/(92*2*pi)."""
    import math

    xs, x = [], 5e-5 * 1.1
    while x < 25.0:
        xs.append(x)
        x = x * 1.1
    ys, y = [], 0.0 + 2.0
    while y < 92.0:
        ys.append(y)
        y = y + 2.0
    xv = np.array([5e-5] + xs + [25.0])
    yv = np.array([0.0] + ys + [92.0])
    q = 4500.0 / (92.0 * 2.0 * math.pi)
    cols, rows = len(xs), len(ys)
    bounds = {0: (0, 1, np.full(cols, 7.0)), 1: (0, 1, np.full(cols, 7.0)), 2: (0, 1, np.full(rows, 7.0)),
              3: (1, 1, np.full(rows, q))}
    init = (np.full((rows, cols), 7.0), np.full((rows, cols), 7.0))
    return Problem2D(xv, yv, 1, 0, bounds, coefs="const", init=init)


def run_gridtest(g, is_gs, steps=195):
    """The stepping loop of AT/GridTest.scala:44-66 (bounds refreshed by noHeatFlow every step)."""
    import math

    heat = 4500.0 / (92.0 * 2.0 * math.pi)
    for _ in range(steps):
        if is_gs:
            g.bounds.left.update(heat)
        else:
            g.set_bound(3, 1, 1, [heat])
        for side in (2, 0, 1):  # Right, Upper, Lower
            g.noHeatFlow(side)
        g.iteration(900.0)
    return snapshot(g)
