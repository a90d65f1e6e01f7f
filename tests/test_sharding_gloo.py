"""CPU tests of the N > 1 host logic (world_size 2, gloo): the line partition covers every line exactly once,
sharded execution needs no exchange (per-rank results concatenate to the single-process result), and the timing
reduction is a max over ranks."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gridsplines_b200 import sharding


def test_block_shard_covers_exactly_once():
    for total in (0, 1, 5, 32, 33, 1 << 20, 1000003):
        for world in (1, 2, 3, 4, 8):
            seen = 0
            for rank in range(world):
                start, count = sharding.block_shard(total, world, rank)
                assert start == seen and count >= 0
                seen += count
            assert seen == total
            counts = [sharding.block_shard(total, world, r)[1] for r in range(world)]
            assert max(counts) - min(counts) <= 1
    with pytest.raises(ValueError):
        sharding.block_shard(10, 2, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, nlines, n, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    sys.path.insert(0, os.path.join(root, "tests"))
    import cases
    from oracle import oracle as O

    s = cases.batch_setup(nlines, n, random_grid=True)
    start, count = sharding.line_shard(nlines, world, rank)
    grid = s["grid0"][start:start + count].copy()
    predict = np.full((count, n), 4.0)
    # each rank steps ITS lines only (the oracle stands in for the device here); no data-path collective
    O.batch1d_step(0, s["leftX"], s["rangeX"], s["rightX"], O.Spline.const(1e-6), 1.0, grid, predict,
                   s["leftT"][start:start + count], s["rightT"][start:start + count], s["time"], nsteps=3)
    np.save(os.path.join(out_dir, f"grid_{rank}.npy"), grid)
    # timing reduction: rank r pretends to have taken (r + 1) ms
    (mx,) = sharding.max_over_ranks([float(rank + 1)])
    assert mx == float(world)
    # every rank reports the same whole-job rate
    rate = sharding.whole_job_rate(count * n, world, mx * 1e-3)
    t = torch.tensor([rate], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if count * world == nlines:
        assert t.item() == rate
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_lines_need_no_exchange(tmp_path, O):
    import cases

    world, nlines, n = 2, 37, 24
    mp.spawn(_worker, args=(world, _free_port(), nlines, n, str(tmp_path)), nprocs=world, join=True)
    parts = [np.load(tmp_path / f"grid_{r}.npy") for r in range(world)]
    whole = cases.run_batch(O, nlines, n, 3, spline="const", random_grid=True, oracle=True)["grid"]
    assert np.array_equal(np.concatenate(parts).view(np.uint64), whole.view(np.uint64))


# ---------------------------------------------------------------------------------------------------------
# row-sharded 2-D grid: the orchestration (x sweep on rows -> all-to-all -> y sweep + update on columns -> back)
# runs here with the oracle as the local solver and gloo as the transport; on GPUs the same driver uses
# libgs_b200 + nccl (tools/run_sharded2d.py).
# ---------------------------------------------------------------------------------------------------------
class _OracleLocal:
    def __init__(self, O, xv, yv, xdir, ydir):
        self.O = O
        self.g = O.TwoDGrid(xv, yv, xdir, ydir)
        ads = [O.AlwaysDefinedSpline(O.Spline.const(v)) for v in (1.5, 2.2e6, 1.5 / 2.2e6)]
        self._keep = ads
        self.g.set_constant_coef(*ads)

    def set_bound(self, side, kind, rep, values):
        v = np.atleast_1d(np.asarray(values, float))
        n = self.g.cols if side in (0, 1) else self.g.rows
        self.g.set_bound(side, kind, rep, v if v.size in (1, n) else v[:1])

    def _arr(self, which):
        return (self.g.grid, self.g.predict, self.g.result)[which]

    def get(self, which):
        return torch.from_numpy(self._arr(which).copy())

    def set(self, which, t):
        self._arr(which)[:] = t.numpy()

    def xIter(self, time):
        self.g.xIter(time)

    def yIterUpdate(self, time):
        self.g.yIter(time)
        self.g.update()


def _sharded_worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    sys.path.insert(0, os.path.join(root, "tests"))
    import cases
    from gridsplines_b200.sharded2d import RowShardedTwoDGrid
    from oracle import oracle as O

    p = cases.bc_problem("TTFF", xdir=1, ydir=0, cols=24, rows=20, coefs="const")
    sg = RowShardedTwoDGrid(lambda xv, yv: _OracleLocal(O, xv, yv, p.xdir, p.ydir), p.xv, p.yv, p.bounds, world, rank)
    sl = slice(sg.r0, sg.r0 + sg.rl)
    sg.set_local_rows(torch.from_numpy(p.grid0[sl].copy()), torch.from_numpy(p.predict0[sl].copy()))
    sg.iteration(900.0, nsteps=3)
    np.save(os.path.join(out_dir, f"rows_{rank}.npy"), sg.local_rows(0).numpy())
    np.save(os.path.join(out_dir, f"pred_{rank}.npy"), sg.local_rows(1).numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [1, 2])
def test_row_sharded_2d_matches_single_process(tmp_path, O, world):
    import cases

    mp.spawn(_sharded_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    p = cases.bc_problem("TTFF", xdir=1, ydir=0, cols=24, rows=20, coefs="const")
    o = p.build_oracle(O)
    for _ in range(3):
        o.iteration(900.0)
    grid = np.concatenate([np.load(tmp_path / f"rows_{r}.npy") for r in range(world)])
    pred = np.concatenate([np.load(tmp_path / f"pred_{r}.npy") for r in range(world)])
    assert np.array_equal(grid.view(np.uint64), o.grid.view(np.uint64))
    assert np.array_equal(pred.view(np.uint64), o.predict.view(np.uint64))


def test_transposes_round_trip_single_rank():
    from gridsplines_b200.sharded2d import cols_to_rows, rows_to_cols

    x = torch.arange(6 * 8, dtype=torch.float64).view(6, 8)
    assert torch.equal(cols_to_rows(rows_to_cols(x, 1), 1), x)
