import sys; sys.path.insert(0, '/root/repo')
import numpy as np, torch
import gridsplines_b200 as gs
stream = torch.cuda.Stream()
ctx = gs.Context(0, stream=stream.cuda_stream)
n = 4096
r = np.random.default_rng(1)
x = gs.XDim(0.0, 0.01 * np.arange(1, n + 1), 0.01 * (n + 1), gs.ORTHO)
y = gs.YDim(0.0, 0.01 * np.arange(1, n + 1), 0.01 * (n + 1), gs.ORTHO)
rho = [0.598, 0.826, 1.121, 1.496, 1.966, 2.547, 3.258, 4.122, 5.156, 6.397, 7.862]
T = [-20.0 + 10.0 * i for i in range(11)]
k = [(t, 1.0 + 0.1 * q) for t, q in zip(T, rho)]; c = [(t, 2.0e6 + 1.0e5 * q) for t, q in zip(T, rho)]
al = [(t, kk / cc) for (t, kk), (_, cc) in zip(k, c)]
coefs = gs.ConstantCoef(gs.Spline.m1Hermite3(k), gs.Spline.m1Hermite3(c), gs.Spline.m1Hermite3(al), ctx=ctx)
kinds = dict(upp=(gs.Many, gs.Temp), low=(gs.Many, gs.Temp), right=(gs.Many, gs.Temp), left=(gs.Many, gs.Temp))
g = gs.TwoDGrid(x, y, coefs=coefs, ctx=ctx, **kinds)
g.bounds.upp.update(20.0); g.bounds.low.update(4.0); g.bounds.left.update(30.0); g.bounds.right.update(7.0)
f = 7.0 + r.uniform(size=(n, n))
g.grid.grid = f; g.grid.predict = f
g.iteration(900.0, nsteps=2)
g.xIter(900.0); g.xIter(900.0); g.yIterUpdate(900.0); g.yIterUpdate(900.0); g.xIter(900.0)
ctx.synchronize()
