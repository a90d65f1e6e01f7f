"""Per-step time of a 512^2 table grid (exact solver) from a uniform start, after reset, on a random field, on a nearly
uniform field and with a few option variants: uniform / zero fields must not leave the fast tiers."""
import sys, os, time
sys.path.insert(0, os.getcwd())
import numpy as np
import gridsplines_b200 as gs
ctx = gs.default_context()
rho = [0.598, 0.826, 1.121, 1.496, 1.966, 2.547, 3.258, 4.122, 5.156, 6.397, 7.862]
T = [-20.0 + 10.0 * i for i in range(11)]
n = 512
x = gs.XDim(0.0, 0.01 * np.arange(1, n + 1), 0.01 * (n + 1), gs.ORTHO)
y = gs.YDim(0.0, 0.01 * np.arange(1, n + 1), 0.01 * (n + 1), gs.ORTHO)
k = gs.Spline.m1Hermite3([(t, 1.0 + 0.1 * q) for t, q in zip(T, rho)]); cc = gs.Spline.m1Hermite3([(t, 2.0e6 + 1.0e5 * q) for t, q in zip(T, rho)])
many = (gs.Many, gs.Temp)
g = gs.TwoDGrid(x, y, coefs=gs.ConstantCoef(k, cc, k / cc), upp=many, low=many, right=many, left=many)
g.bounds.upp.update(20.0); g.bounds.low.update(4.0); g.bounds.left.update(30.0); g.bounds.right.update(7.0)
def run(tag, steps=300):
    ctx.synchronize(); t0 = time.perf_counter()
    g.iteration(900.0, nsteps=steps); ctx.synchronize()
    print(f"{tag:40s} {1e6*(time.perf_counter()-t0)/steps:7.1f} us/step  status {g.status}", flush=True)
g.fill(7.0); run("uniform 7.0, first 300 steps")
run("next 300")
run("next 300")
g.fill(7.0); run("reset to uniform 7.0, 300 steps")
r = np.random.default_rng(1); f = 7.0 + r.uniform(size=(n, n)); g.grid.grid = f; g.grid.predict = f
run("random 7 + U[0,1), 300 steps")
g.fill(7.0); run("uniform again, 30 steps", 30); run("next 30", 30); run("next 30", 30); run("next 100", 100)
f = 7.0 + 1e-9 * r.uniform(size=(n, n)); g.grid.grid = f; g.grid.predict = f
run("7 + 1e-9 U (nearly uniform), 30 steps", 30); run("next 30", 30)
ctx.set_option("native_div", 1)
f = 7.0 + r.uniform(size=(n, n)); g.grid.grid = f; g.grid.predict = f
run("random, native_div = 1, 100 steps", 100)
ctx.set_option("native_div", 0)
g.fill(7.0); ctx.set_option("k7_fuse", 0); run("uniform, separate faces pass (k7_fuse = 0), 30 steps", 30); run("next 30", 30)
