"""Times TwoDGrid.iteration at a chosen size / solver (C3 shape by default)."""
import argparse, sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gridsplines_b200 as gs

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=4096)
ap.add_argument("--rows", type=int, default=0, help="rows when not square (a row shard of a larger grid)")
ap.add_argument("--steps", type=int, default=5)
ap.add_argument("--coefs", default="const")
ap.add_argument("--solver", action="append", type=int, default=[])
ap.add_argument("--sweep", action="append", default=[])
a = ap.parse_args()
stream = torch.cuda.Stream()
ctx = gs.Context(0, stream=stream.cuda_stream)
n = a.n
r = np.random.default_rng(1)
x = gs.XDim(0.0, 0.01 * np.arange(1, n + 1), 0.01 * (n + 1), gs.ORTHO)
nr = a.rows or n
y = gs.YDim(0.0, 0.01 * np.arange(1, nr + 1), 0.01 * (nr + 1), gs.ORTHO)
if a.coefs == "const":
    coefs = gs.ConstantCoef(gs.Spline.const(1.5), gs.Spline.const(2.2e6), gs.Spline.const(1.5 / 2.2e6), ctx=ctx)
    kinds = dict(upp=(gs.Many, gs.Temp), low=(gs.Many, gs.Temp), right=(gs.Many, gs.Flow), left=(gs.Many, gs.Flow))
else:
    rho = [0.598, 0.826, 1.121, 1.496, 1.966, 2.547, 3.258, 4.122, 5.156, 6.397, 7.862]
    T = [-20.0 + 10.0 * i for i in range(11)]
    k = [(t, 1.0 + 0.1 * q) for t, q in zip(T, rho)]; c = [(t, 2.0e6 + 1.0e5 * q) for t, q in zip(T, rho)]
    al = [(t, kk / cc) for (t, kk), (_, cc) in zip(k, c)]
    coefs = gs.ConstantCoef(gs.Spline.m1Hermite3(k), gs.Spline.m1Hermite3(c), gs.Spline.m1Hermite3(al), ctx=ctx)
    kinds = dict(upp=(gs.Many, gs.Temp), low=(gs.Many, gs.Temp), right=(gs.Many, gs.Temp), left=(gs.Many, gs.Temp))
g = gs.TwoDGrid(x, y, coefs=coefs, ctx=ctx, **kinds)
g.bounds.upp.update(20.0); g.bounds.low.update(4.0)
if a.coefs == "const":
    g.bounds.left.update(50.0); g.bounds.right.update(0.0)
else:
    g.bounds.left.update(30.0); g.bounds.right.update(7.0)
f = 7.0 + r.uniform(size=(nr, n))
g.grid.grid = f; g.grid.predict = f
def timeit():
    with torch.cuda.stream(stream):
        g.iteration(900.0, nsteps=2)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream); g.iteration(900.0, nsteps=a.steps); e1.record(stream)
        torch.cuda.synchronize()
    return e0.elapsed_time(e1) / a.steps
def report(tag, ms):
    ups = nr * n / (ms * 1e-3)
    print(f"{tag:40s} {ms:8.3f} ms/step  {ups/1e9:8.2f} Gupd/s  {64*ups/6543.4e9*100:6.1f}% of HBM roofline", flush=True)
for s in (a.solver or [0]):
    g.set_solver(s)
    report(f"solver={s}", timeit())
    for sw in a.sweep:
        name, vals = sw.split("=")
        for v in vals.split(","):
            ctx.set_option(name, int(v))
            report(f"solver={s} {name}={v}", timeit())
print("status", g.status)
