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


# ---------------------------------------------------------------------------------------------------------
# SpikeShardedTwoDGrid: the y sweep stays row-sharded; ranks exchange two predict rows and six numbers per column.
# The local solver here is test infrastructure: the oracle for the x sweep (whole lines) and dense linear algebra
# for a column's local block, with the unknown neighbour values carried symbolically -- an independent statement
# of what gs_grid2d_yiter_update_begin / _end compute on the device (tests/test_gpu_split_sweep.py checks those).
# ---------------------------------------------------------------------------------------------------------
class _SlabLocal(_OracleLocal):
    A = 1.5 / 2.2e6   # the constant tempConductivity of _OracleLocal: every face value

    def __init__(self, O, xv, yv, xdir, ydir):
        super().__init__(O, xv, yv, xdir, ydir)
        self.cf = O.predef(ydir, yv[0], yv[1:-1], yv[-1], 1.0).reshape(-1, 2)    # Dim.coefs of the local rows
        self.bv = {}

    def set_bound(self, side, kind, rep, values):
        super().set_bound(side, kind, rep, values)
        self.bv[side] = np.broadcast_to(np.atleast_1d(np.asarray(values, float)), (self.g.cols,)) if side in (0, 1) else None

    def predict_edges(self):
        return torch.from_numpy(np.stack((self.g.predict[0], self.g.predict[-1])))

    def yIterUpdateBegin(self, time, halo_prev, halo_next):
        from scipy.linalg import solve_banded

        n, cols = self.g.rows, self.g.cols
        sub, sup = self.cf[:, 0] * self.A, self.cf[:, 1] * self.A
        diag = -(sub + sup) - 1.0 / time
        r = -self.g.result / time
        if halo_prev is None:
            r[0] -= sub[0] * self.bv[0]          # closed end: the bound's temperature is a known neighbour
        if halo_next is None:
            r[-1] -= sup[-1] * self.bv[1]
        ab = np.zeros((3, n))
        ab[0, 1:], ab[1], ab[2, :-1] = sup[:-1], diag, sub[1:]
        e0, e1 = np.zeros(n), np.zeros(n)
        e0[0], e1[-1] = -sub[0], -sup[-1]
        sol = solve_banded((1, 1), ab, np.column_stack((r, e0, e1)))
        self.u = sol[:, :cols]
        self.v = sol[:, cols] if halo_prev is not None else np.zeros(n)      # d x / d L0
        self.z = sol[:, cols + 1] if halo_next is not None else np.zeros(n)  # d x / d X
        ones = np.ones(cols)
        C, D, W = self.z[-1] * ones, self.u[-1], self.v[-1] * ones           # E = C X + D + W L0
        if halo_next is not None:
            P, Q, R = self.z[0] / C, self.u[0] - self.z[0] * D / C, self.v[0] - self.z[0] * W / C
        else:
            P, Q, R = 0.0 * ones, self.u[0], self.v[0] * ones                # x_first = P E + Q + R L0
        return torch.from_numpy(np.stack((C, D, W, P, Q, R)))

    def yIterUpdateEnd(self, time, gathered, world, rank):
        G = gathered.numpy()
        cols = self.g.cols
        c, d = np.zeros(cols), np.zeros(cols)
        cs, ds = [], []
        for p in range(world):
            de, la, om = G[p, 0], G[p, 1], G[p, 2]
            Pn, Qn, Rn = (G[p + 1, 3], G[p + 1, 4], G[p + 1, 5]) if p + 1 < world else (0.0, 0.0, 0.0)
            a, b, g, rr = -om, 1.0 - de * Rn, -(de * Pn), la + de * Qn
            inv = 1.0 / (b + a * c)
            c, d = -g * inv, (rr - a * d) * inv
            cs.append(c); ds.append(d)
        E = [None] * world
        x = np.zeros(cols)
        for p in range(world - 1, -1, -1):
            x = cs[p] * x + ds[p]
            E[p] = x
        L0 = E[rank - 1] if rank > 0 else 0.0
        X = G[rank + 1, 3] * E[rank + 1] + G[rank + 1, 4] + G[rank + 1, 5] * E[rank] if rank + 1 < world else 0.0
        xn = self.u + np.outer(self.v, L0 * np.ones(cols)) + np.outer(self.z, X * np.ones(cols))
        dl = xn - self.g.grid
        self.g.predict[:] = np.where(np.abs(dl) > 1.5, xn, xn + dl / 2.0)    # Grid.update
        self.g.grid[:] = xn


def _spike_worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    sys.path.insert(0, os.path.join(root, "tests"))
    import cases
    from gridsplines_b200.sharded2d import SpikeShardedTwoDGrid
    from oracle import oracle as O

    p = cases.bc_problem("TTFT", xdir=1, ydir=0, cols=24, rows=40, coefs="const")
    sg = SpikeShardedTwoDGrid(lambda xv, yv: _SlabLocal(O, xv, yv, p.xdir, p.ydir), p.xv, p.yv, p.bounds, world, rank)
    sl = slice(sg.r0, sg.r0 + sg.rl)
    sg.set_local_rows(torch.from_numpy(p.grid0[sl].copy()), torch.from_numpy(p.predict0[sl].copy()))
    sg.iteration(900.0, nsteps=3)
    np.save(os.path.join(out_dir, f"rows_{rank}.npy"), sg.local_rows(0).numpy())
    np.save(os.path.join(out_dir, f"pred_{rank}.npy"), sg.local_rows(1).numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [1, 2, 4])
def test_spike_sharded_2d_matches_single_process(tmp_path, O, world):
    import cases

    mp.spawn(_spike_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    p = cases.bc_problem("TTFT", xdir=1, ydir=0, cols=24, rows=40, coefs="const")
    o = p.build_oracle(O)
    for _ in range(3):
        o.iteration(900.0)
    grid = np.concatenate([np.load(tmp_path / f"rows_{r}.npy") for r in range(world)])
    pred = np.concatenate([np.load(tmp_path / f"pred_{r}.npy") for r in range(world)])
    scale = np.abs(o.grid).max()
    assert np.abs(grid - o.grid).max() / scale <= 1e-12
    assert np.abs(pred - o.predict).max() / scale <= 1e-12
