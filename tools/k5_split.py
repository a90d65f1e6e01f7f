"""Times the x sweep and the y sweep (+ Grid.update) of a constant-coefficient TwoDGrid separately (K5 / K5T), for a list
of tuning options:  python tools/k5_split.py --n 4096 --sweep k5t_sx=8,16 --sweep k5t_sy=4,8"""
import argparse, sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gridsplines_b200 as gs

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=4096)
ap.add_argument("--rows", type=int, default=0)
ap.add_argument("--steps", type=int, default=20)
ap.add_argument("--sweep", action="append", default=[])
a = ap.parse_args()
stream = torch.cuda.Stream()
ctx = gs.Context(0, stream=stream.cuda_stream)
n, nr = a.n, a.rows or a.n
r = np.random.default_rng(1)
x = gs.XDim(0.0, 0.01 * np.arange(1, n + 1), 0.01 * (n + 1), gs.ORTHO)
y = gs.YDim(0.0, 0.01 * np.arange(1, nr + 1), 0.01 * (nr + 1), gs.ORTHO)
coefs = gs.ConstantCoef(gs.Spline.const(1.5), gs.Spline.const(2.2e6), gs.Spline.const(1.5 / 2.2e6), ctx=ctx)
g = gs.TwoDGrid(x, y, coefs=coefs, ctx=ctx, upp=(gs.Many, gs.Temp), low=(gs.Many, gs.Temp), right=(gs.Many, gs.Flow),
                left=(gs.Many, gs.Flow))
g.bounds.upp.update(20.0); g.bounds.low.update(4.0); g.bounds.left.update(50.0); g.bounds.right.update(0.0)
f = 7.0 + r.uniform(size=(nr, n))
g.grid.grid = f; g.grid.predict = f


def timeit(fn):
    with torch.cuda.stream(stream):
        fn(); fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(a.steps):
            fn()
        e1.record(stream)
        torch.cuda.synchronize()
    return e0.elapsed_time(e1) / a.steps


def report(tag):
    tx = timeit(lambda: g.xIter(900.0))
    ty = timeit(lambda: g.yIterUpdate(900.0))
    ts = timeit(lambda: g.iteration(900.0))
    nodes = nr * n
    print(f"{tag:24s} x {tx*1e3:7.1f} us ({24*nodes/tx/1e9:6.0f} GB/s alg)  y {ty*1e3:7.1f} us ({40*nodes/ty/1e9:6.0f} GB/s alg)"
          f"  step {ts:7.4f} ms = {64*nodes/ts/1e6/6543.4*100:5.1f} %", flush=True)


report("default")
for sw in a.sweep:
    name, vals = sw.split("=")
    for v in vals.split(","):
        ctx.set_option(name, int(v))
        report(f"{name}={v}")
print("status", g.status)
