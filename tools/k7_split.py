"""Times the x sweep, the fused y sweep and the whole step of a TwoDGrid with Hermite tables (K7 / K7F) for tuning options:
   python tools/k7_split.py --n 16384 --sweep k7_nsb=3,4,8"""
import argparse, sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gridsplines_b200 as gs

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=16384)
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--sweep", action="append", default=[])
a = ap.parse_args()
stream = torch.cuda.Stream()
ctx = gs.Context(0, stream=stream.cuda_stream)
n = a.n
r = np.random.default_rng(1)
x = gs.XDim(0.0, 0.01 * np.arange(1, n + 1), 0.01 * (n + 1), gs.ORTHO)
y = gs.YDim(0.0, 0.01 * np.arange(1, n + 1), 0.01 * (n + 1), gs.ORTHO)
rho = [0.598, 0.826, 1.121, 1.496, 1.966, 2.547, 3.258, 4.122, 5.156, 6.397, 7.862]
T = [-20.0 + 10.0 * i for i in range(11)]
k = [(t, 1.0 + 0.1 * q) for t, q in zip(T, rho)]; c = [(t, 2.0e6 + 1.0e5 * q) for t, q in zip(T, rho)]
al = [(t, kk / cc) for (t, kk), (_, cc) in zip(k, c)]
coefs = gs.ConstantCoef(gs.Spline.m1Hermite3(k), gs.Spline.m1Hermite3(c), gs.Spline.m1Hermite3(al), ctx=ctx)
g = gs.TwoDGrid(x, y, coefs=coefs, ctx=ctx, upp=(gs.Many, gs.Temp), low=(gs.Many, gs.Temp), right=(gs.Many, gs.Temp),
                left=(gs.Many, gs.Temp))
g.bounds.upp.update(20.0); g.bounds.low.update(4.0); g.bounds.left.update(30.0); g.bounds.right.update(7.0)
f = 7.0 + r.uniform(size=(n, n))
g.grid.grid = f; g.grid.predict = f


def timeit(fn):
    with torch.cuda.stream(stream):
        fn()
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
    ts = timeit(lambda: g.iteration(900.0))
    print(f"{tag:24s} x {tx:7.3f} ms   step {ts:7.3f} ms (y ~ {ts - tx:6.3f}) = {64*n*n/ts/1e6/6543.4*100:5.1f} %", flush=True)


report("default")
for sw in a.sweep:
    name, vals = sw.split("=")
    for v in vals.split(","):
        ctx.set_option(name, int(v))
        report(f"{name}={v}")
print("status", g.status)
