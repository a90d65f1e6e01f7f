"""Runs a few steps of the C2 kernel at a chosen size (for ncu)."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gridsplines_b200 as gs
nl = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
opts = [a for a in sys.argv[2:] if "=" in a and not a.startswith("spline=")]
spline = ([a.split("=")[1] for a in sys.argv[2:] if a.startswith("spline=")] or ["const"])[0]
ctx = gs.default_context()
for kv in opts:
    k, v = kv.split("="); ctx.set_option(k, int(v))
n = 256
r = np.random.default_rng(1)
RHO = [0.598, 0.826, 1.121, 1.496, 1.966, 2.547, 3.258, 4.122, 5.156, 6.397, 7.862]
cond = gs.Spline.const(1e-6) if spline == "const" else gs.Spline.m1Hermite3(
    [(-20.0 + 10.0 * i, 1e-6 * (1.0 + 0.1 * RHO[i])) for i in range(11)])      # spline=hermite: the general kernel K2
g = gs.BatchedOneDGrid(0.0, 0.01 * np.arange(1, n + 1), 0.01 * (n + 1), cond, 1.0, gs.ORTHO, nl, ctx=ctx)
g.upload(gs.GRID, r.uniform(0, 30, (nl, n)))
g.set_bc(r.uniform(10, 30, nl), r.uniform(0, 10, nl))
for _ in range(4):
    g.step(900.0)
ctx.synchronize()
print("ok", g.status)
