"""C4 (or C3) on N GPUs from ONE process through gs_sharded2d_* -- the multi-device drop-in boundary: strong scaling
against the same library call on one device, the distance between the two results, and the distance from the exact
single-GPU solver.

    python tools/run_sharded_abi.py --size 16384 --devices 8 [--steps 5] [--coefs hermite]
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import gridsplines_b200 as gs  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--size", type=int, default=16384)
ap.add_argument("--devices", type=int, default=torch.cuda.device_count())
ap.add_argument("--steps", type=int, default=5)
ap.add_argument("--coefs", default="hermite")
ap.add_argument("--no-exact", action="store_true")
a = ap.parse_args()
n = a.size
RHO = [0.598, 0.826, 1.121, 1.496, 1.966, 2.547, 3.258, 4.122, 5.156, 6.397, 7.862]
T = [-20.0 + 10.0 * i for i in range(11)]
k = gs.Spline.m1Hermite3([(t, 1.0 + 0.1 * q) for t, q in zip(T, RHO)])
c = gs.Spline.m1Hermite3([(t, 2.0e6 + 1.0e5 * q) for t, q in zip(T, RHO)])
tables = (k, c, k / c) if a.coefs == "hermite" else (gs.Spline.const(1.5), gs.Spline.const(2.2e6), gs.Spline.const(1.5 / 2.2e6))
x = gs.XDim(0.05, 0.05 * 1.0005 ** np.arange(1, n + 1), 0.05 * 1.0005 ** (n + 1), gs.RADIAL)
y = gs.YDim(0.0, 0.01 * np.arange(1, n + 1), 0.01 * (n + 1), gs.ORTHO)
many = (gs.Many, gs.Temp)
field = torch.empty((n, n), dtype=torch.float64).pin_memory()
field.copy_(7.0 + torch.rand((n, n), dtype=torch.float64, generator=torch.Generator().manual_seed(1234)))


def run(ndev):
    s = gs.ShardedTwoDGrid(x, y, list(range(ndev)), upp=many, low=many, right=many, left=many)
    s.set_constant_coef(*tables)
    for side, v in ((gs.UPP, 7.0), (gs.LOW, 7.0), (gs.RIGHT, 7.0), (gs.LEFT, 30.0)):
        s.set_bound(side, v)
    s.upload_ptr(gs.GRID, field.data_ptr())
    s.upload_ptr(gs.PREDICT, field.data_ptr())
    s.iteration(900.0, nsteps=2)
    l0 = s.launch_count
    ms = s.iteration_timed(900.0, nsteps=a.steps) / a.steps
    launches = (s.launch_count - l0) / a.steps
    out = s.download(gs.GRID)
    assert s.status == 0
    s.close()
    return ms, out, launches


ms_n, got, launches = run(a.devices)
ms_1, one, _ = run(1)
res = {"config": f"C4 shape {n}x{n} {a.coefs}", "n_gpus": a.devices, "ms_per_step": ms_n, "n1_ms_per_step": ms_1,
       "updates_per_s": n * n / (ms_n * 1e-3), "efficiency_vs_n1": ms_1 / (ms_n * a.devices),
       "kernel_launches_per_step": launches, "exchange": "peer-to-peer stores from k6_sweep<K6_BEGIN> + halo rows read in place",
       "error_vs_n1_same_solver": {"field": float(np.abs(got - one).max() / np.abs(one).max()),
                                   "elementwise": float((np.abs(got - one) / np.abs(one)).max())}}
if not a.no_exact:
    ctx = gs.default_context()
    g = gs.TwoDGrid(x, y, coefs=gs.ConstantCoef(*tables), upp=many, low=many, right=many, left=many)
    for b, v in ((g.bounds.upp, 7.0), (g.bounds.low, 7.0), (g.bounds.right, 7.0), (g.bounds.left, 30.0)):
        b.update(v)
    g.set_solver(gs.SOLVER_THOMAS)
    g.upload_ptr(gs.GRID, field.data_ptr()); g.upload_ptr(gs.PREDICT, field.data_ptr())
    g.iteration(900.0, nsteps=2 + a.steps)
    ex = g.grid.grid
    res["error_vs_exact_solver"] = {"field": float(np.abs(got - ex).max() / np.abs(ex).max()),
                                    "elementwise": float((np.abs(got - ex) / np.abs(ex)).max()), "steps": 2 + a.steps}
print(json.dumps(res), flush=True)
