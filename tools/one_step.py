"""Runs a few steps of the C2 kernel at a chosen size (for ncu)."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gridsplines_b200 as gs
nl = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
opts = sys.argv[2:]
ctx = gs.default_context()
for kv in opts:
    k, v = kv.split("="); ctx.set_option(k, int(v))
n = 256
r = np.random.default_rng(1)
g = gs.BatchedOneDGrid(0.0, 0.01 * np.arange(1, n + 1), 0.01 * (n + 1), gs.Spline.const(1e-6), 1.0, gs.ORTHO, nl, ctx=ctx)
g.upload(gs.GRID, r.uniform(0, 30, (nl, n)))
g.set_bc(r.uniform(10, 30, nl), r.uniform(0, 10, nl))
for _ in range(4):
    g.step(900.0)
ctx.synchronize()
print("ok", g.status)
