"""gs_sharded2d_*: one TwoDGrid on several devices behind the C ABI, the exchange done by the kernels over peer memory
(csrc/gs_sharded2d.cu).  On a one-GPU box the shards all live on device 0 (the same code path: separate contexts,
streams and events per shard); with more GPUs visible they spread over them."""
import numpy as np
import pytest

import cases
from helpers import check_tol

pytestmark = pytest.mark.gpu


def _devices(nshards):
    import torch

    n = torch.cuda.device_count()
    return [p % n for p in range(nshards)]


@pytest.mark.parametrize("coefs", ["const", "hermite"])
@pytest.mark.parametrize("nshards", [1, 2, 4])
def test_sharded_abi_matches_the_single_device_solver(O, gs, coefs, nshards):
    cols, rows = 288, 512
    p = cases.bc_problem("TTTT", xdir=1, ydir=0, cols=cols, rows=rows, coefs=coefs)
    r = np.random.default_rng(3)
    p.grid0 = r.uniform(-10.0, 60.0, p.grid0.shape)
    p.predict0 = p.grid0 + r.uniform(-3.0, 3.0, p.grid0.shape)
    x = gs.XDim(p.xv[0], p.xv[1:-1], p.xv[-1], 1)
    y = gs.YDim(p.yv[0], p.yv[1:-1], p.yv[-1], 0)
    many = (gs.Many, gs.Temp)
    s = gs.ShardedTwoDGrid(x, y, _devices(nshards), upp=many, low=many, right=many, left=many)
    if coefs == "const":
        s.set_constant_coef(gs.Spline.const(1.5), gs.Spline.const(2.2e6), gs.Spline.const(1.5 / 2.2e6))
    else:
        from helpers import hermite_coefs

        _, (gk, gc, ga) = hermite_coefs(O, gs)
        s.set_constant_coef(gk, gc, ga)
    for side, (_, _, values) in p.bounds.items():
        s.set_bound(side, values)
    s.upload(gs.GRID, p.grid0)
    s.upload(gs.PREDICT, p.predict0)
    s.iteration(900.0, nsteps=3)
    o = p.build_oracle(O, ascending=False)
    for _ in range(3):
        o.iteration(900.0)
    O.set_ascending_fold(True)
    tol = 1e-12 if coefs == "const" else 2e-11       # tests/test_gpu_2d.py explains the table constant
    check_tol(s.download(gs.GRID), o.grid, f"sharded ABI {coefs} x{nshards} grid", tol)
    check_tol(s.download(gs.PREDICT), o.predict, f"sharded ABI {coefs} x{nshards} predict", tol)
    assert s.status == 0
    s.close()


def test_sharded_abi_argument_checks(gs):
    x = gs.XDim(0.0, 0.01 * np.arange(1, 65), 0.65, 0)
    y = gs.YDim(0.0, 0.01 * np.arange(1, 101), 1.01, 0)
    with pytest.raises(gs.GsError):
        gs.ShardedTwoDGrid(x, y, [0, 0, 0])             # 100 rows do not divide over 3 shards
    with pytest.raises(gs.GsError):
        gs.ShardedTwoDGrid(x, y, [0, 0], low=(gs.Many, gs.Flow))      # lower HeatFlow on a sharded grid
