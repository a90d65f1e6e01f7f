"""Per-step cost of TwoDGrid.iteration on small and medium grids: 300 single-step calls (a user loop that also updates
bounds every step) against one call with nsteps = 300 -- host enqueue time and total time."""
import sys, os, time
sys.path.insert(0, os.getcwd())
import numpy as np
import gridsplines_b200 as gs
ctx = gs.default_context()
rho = [0.598, 0.826, 1.121, 1.496, 1.966, 2.547, 3.258, 4.122, 5.156, 6.397, 7.862]
T = [-20.0 + 10.0 * i for i in range(11)]
def mk(n, m, coefs):
    x = gs.XDim(0.0, 0.01 * np.arange(1, n + 1), 0.01 * (n + 1), gs.ORTHO)
    y = gs.YDim(0.0, 0.01 * np.arange(1, m + 1), 0.01 * (m + 1), gs.ORTHO)
    if coefs == "const":
        c = gs.ConstantCoef(gs.Spline.const(1.5), gs.Spline.const(2.2e6), gs.Spline.const(1.5 / 2.2e6))
    else:
        k = gs.Spline.m1Hermite3([(t, 1.0 + 0.1 * q) for t, q in zip(T, rho)]); cc = gs.Spline.m1Hermite3([(t, 2.0e6 + 1.0e5 * q) for t, q in zip(T, rho)])
        c = gs.ConstantCoef(k, cc, k / cc)
    many = (gs.Many, gs.Temp)
    g = gs.TwoDGrid(x, y, coefs=c, upp=many, low=many, right=many, left=many)
    g.bounds.upp.update(20.0); g.bounds.low.update(4.0); g.bounds.left.update(30.0); g.bounds.right.update(7.0)
    g.fill(7.0)
    return g
for n, m, coefs in ((136, 44, "const"), (136, 44, "hermite"), (512, 512, "hermite"), (512, 512, "const")):
    g = mk(n, m, coefs)
    g.iteration(900.0, nsteps=10); ctx.synchronize()
    t0 = time.perf_counter()
    for _ in range(300):
        g.iteration(900.0)
    t1 = time.perf_counter()          # host time to enqueue
    ctx.synchronize()
    t2 = time.perf_counter()
    g.iteration(900.0, nsteps=300); ctx.synchronize()
    t3 = time.perf_counter()
    for _ in range(300):
        g.iteration(900.0)
    ctx.synchronize()
    t4 = time.perf_counter()
    print(f"   (again, 300 single-step calls: {1e6*(t4-t3)/300:6.1f} us/step)")
    print(f"{n}x{m} {coefs:8s}: 300 single-step calls: enqueue {1e6*(t1-t0)/300:6.1f} us/step, total {1e6*(t2-t0)/300:6.1f} us/step; one call with nsteps=300: {1e6*(t3-t2)/300:6.1f} us/step")
