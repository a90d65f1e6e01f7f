"""Times gs_grid1d_step for several tuning options in one process (C2 shape unless overridden)."""
import argparse, sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gridsplines_b200 as gs

ap = argparse.ArgumentParser()
ap.add_argument("--lines", type=int, default=1 << 20)
ap.add_argument("--nodes", type=int, default=256)
ap.add_argument("--steps", type=int, default=5)
ap.add_argument("--spline", default="const")
ap.add_argument("--sweep", action="append", default=[], help="name=v1,v2,...")
ap.add_argument("--fused", type=int, default=0, help="also time one call with nsteps=N")
a = ap.parse_args()
stream = torch.cuda.Stream()
ctx = gs.Context(0, stream=stream.cuda_stream)
n, nl = a.nodes, a.lines
r = np.random.default_rng(1)
if a.spline == "const":
    spl = gs.Spline.const(1e-6)
else:
    rho = [0.598, 0.826, 1.121, 1.496, 1.966, 2.547, 3.258, 4.122, 5.156, 6.397, 7.862]
    spl = gs.Spline.m1Hermite3([(-20.0 + 10.0 * i, 1e-6 * (1.0 + 0.1 * rho[i])) for i in range(11)])
g = gs.BatchedOneDGrid(0.0, 0.01 * np.arange(1, n + 1), 0.01 * (n + 1), spl, 1.0, gs.ORTHO, nl, ctx=ctx)
h = torch.empty((nl, n), dtype=torch.float64)
h.uniform_(0.0, 30.0)
g.upload_ptr(gs.GRID, h.data_ptr())
g.set_bc(r.uniform(10, 30, nl), r.uniform(0, 10, nl))
def timeit(nsteps_per_call=1, calls=a.steps):
    with torch.cuda.stream(stream):
        g.step(900.0, nsteps=nsteps_per_call); g.step(900.0, nsteps=nsteps_per_call)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(calls):
            g.step(900.0, nsteps=nsteps_per_call)
        e1.record(stream)
        torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (calls * nsteps_per_call)
def report(tag, ms):
    ups = nl * n / (ms * 1e-3)
    print(f"{tag:40s} {ms:8.3f} ms/step  {ups/1e9:8.2f} Gupd/s  {32*ups/6543.4e9*100:6.1f}% of HBM roofline", flush=True)
report("default", timeit())
for sw in a.sweep:
    name, vals = sw.split("=")
    for v in vals.split(","):
        ctx.set_option(name, int(v))
        report(f"{name}={v}", timeit())
        if a.fused:
            report(f"{name}={v} fused nsteps={a.fused}", timeit(a.fused, 2))
print("status", g.status)
