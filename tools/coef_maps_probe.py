"""Per-step time of the exact solver (AUTO -> K7F) with a SINGLE table against per-row tables (VarYCoef: layered material)
and a patched map, same size: the Coefficients hierarchy must not fall off the fast face path."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import gridsplines_b200 as gs
ctx = gs.default_context()
rho = [0.598, 0.826, 1.121, 1.496, 1.966, 2.547, 3.258, 4.122, 5.156, 6.397, 7.862]
T = [-20.0 + 10.0 * i for i in range(11)]
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
x = gs.XDim(0.0, 0.01 * np.arange(1, n + 1), 0.01 * (n + 1), gs.ORTHO)
y = gs.YDim(0.0, 0.01 * np.arange(1, n + 1), 0.01 * (n + 1), gs.ORTHO)
mk = lambda s: gs.AlwaysDefinedSpline(gs.Spline.m1Hermite3([(t, s * (1.0 + 0.1 * q)) for t, q in zip(T, rho)]))
layers = [(mk(1.0 + 0.2 * l), mk(2.0e6), mk((1.0 + 0.2 * l) / 2.0e6)) for l in range(4)]
def per_row(i): return layers[min(3, (i * 4) // (n + 1))]
single = gs.ConstantCoef(*layers[0])
vary = gs.VarYCoef([per_row(i)[0] for i in range(n + 1)], [per_row(i)[1] for i in range(n + 1)], [per_row(i)[2] for i in range(n + 1)])
patch = gs.PatchedXCoef(vary, gs.PatchXCoef([layers[3][0]] * (n + 1), [layers[3][1]] * (n + 1), [layers[3][2]] * (n + 1), (n // 4, n // 2), (n // 4, n // 2)))
many = (gs.Many, gs.Temp)
r = np.random.default_rng(1)
f = 7.0 + r.uniform(size=(n, n))
for name, coefs in (("single table", single), ("VarYCoef (4 layers)", vary), ("PatchedXCoef (dense map)", patch)):
    g = gs.TwoDGrid(x, y, coefs=coefs, upp=many, low=many, right=many, left=many)
    g.bounds.upp.update(20.0); g.bounds.low.update(4.0); g.bounds.left.update(30.0); g.bounds.right.update(7.0)
    g.grid.grid = f; g.grid.predict = f
    g.iteration(900.0, nsteps=3); ctx.synchronize()
    t0 = time.perf_counter(); g.iteration(900.0, nsteps=10); ctx.synchronize()
    print(f"{n}^2 {name:28s} {1e3 * (time.perf_counter() - t0) / 10:8.3f} ms/step  status {g.status}", flush=True)
    g.close()
