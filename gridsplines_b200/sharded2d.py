"""Row-sharded TwoDGrid over the GPUs of one box (SURVEY 8e, config C4's partitioning).

Rank p owns rows [p R/P, (p+1) R/P) of grid / predict / result.  One time step (TwoDGrid.iteration,
A/TwoDGrid.scala:254-258):

  1. x sweep on the local rows -- every row is an independent line, no communication;
  2. `result` is transposed to a column-sharded layout with ONE all-to-all (NCCL over NVLink; each rank sends
     (P-1)/P of its shard), so that every column is whole on one rank;
  3. y sweep + Grid.update on the local columns (grid and predict are kept column-sharded as well);
  4. grid' and predict' return to the row-sharded layout with two more all-to-alls for the next x sweep.

The exchange is the only collective on the data path; both sweeps are the single-GPU kernels unchanged, so with
GS_SOLVER_THOMAS the sharded run is bit-identical to the single-GPU run.  (The SPIKE alternative -- exchanging the
partitioned solver's reduced-system coefficients instead of whole arrays -- is the planned optimisation; see
DESIGN.md.)  Restrictions: rows and cols divisible by the world size; a lower HeatFlow bound is not supported when
sharded (the reference's quirk there reads the neighbouring column, A/TwoDGrid.scala:241).

The driver is backend-agnostic: `make_local(xvalues, yvalues)` returns an object with
    set_bound(side, kind, rep, values) / get(which) -> torch tensor / set(which, tensor) / xIter(t) / yIterUpdate(t)
so the same orchestration runs on CUDA (`GsLocalGrid`, nccl) and in the CPU tests (oracle-backed, gloo).
"""
from __future__ import annotations

from typing import Callable, Dict, Tuple

import numpy as np
import torch
import torch.distributed as dist

from . import _lib as L
from .sharding import block_shard

UPP, LOW, RIGHT, LEFT = 0, 1, 2, 3
GRID, PREDICT, RESULT = 0, 1, 2


def rows_to_cols(x: torch.Tensor, world: int, out: torch.Tensor = None) -> torch.Tensor:
    """[rows/P, cols] on every rank (row-sharded)  ->  [rows, cols/P] (column-sharded), one all-to-all.
    `out` (contiguous [rows, cols/P]) receives the blocks directly when given."""
    rl, cols = x.shape
    cl = cols // world
    send = x.view(rl, world, cl).permute(1, 0, 2).contiguous()      # block j = my rows x rank j's columns
    recv = send.new_empty(send.shape) if out is None else out.view(world, rl, cl)
    if world > 1:
        dist.all_to_all_single(recv, send)
    else:
        recv.copy_(send)
    return recv.view(world * rl, cl)                                 # block i = rank i's rows x my columns


def cols_to_rows(x: torch.Tensor, world: int, out: torch.Tensor = None) -> torch.Tensor:
    """[rows, cols/P] (column-sharded, contiguous)  ->  [rows/P, cols] (row-sharded), one all-to-all."""
    rows, cl = x.shape
    rl = rows // world
    send = x.view(world, rl, cl)                                     # block i = rank i's rows x my columns (a view)
    recv = torch.empty_like(send)
    if world > 1:
        dist.all_to_all_single(recv, send)
    else:
        recv.copy_(send)
    res = recv.permute(1, 0, 2)                                      # block j = my rows x rank j's columns
    if out is None:
        return res.reshape(rl, world * cl).contiguous()
    out.view(rl, world, cl).copy_(res)
    return out


class RowShardedTwoDGrid:
    def __init__(self, make_local: Callable, xvalues, yvalues, bounds: Dict[int, Tuple[int, int, np.ndarray]],
                 world: int, rank: int):
        self.world, self.rank = world, rank
        xv, yv = np.asarray(xvalues, float), np.asarray(yvalues, float)
        self.cols, self.rows = len(xv) - 2, len(yv) - 2
        if self.rows % world or self.cols % world:
            raise ValueError("rows and cols must be divisible by the world size")
        if bounds.get(LOW, (0,))[0] == 1 and world > 1:
            raise ValueError("a lower HeatFlow bound is not supported on a sharded grid")
        self.r0, self.rl = block_shard(self.rows, world, rank)
        self.c0, self.cl = block_shard(self.cols, world, rank)
        # Dim.coefs of node i needs its two neighbours' coordinates: slicing `values` keeps them (A/Dim.scala:13-15)
        self.gx = make_local(xv, yv[self.r0:self.r0 + self.rl + 2])     # my rows, all columns: x sweep
        self.gy = make_local(xv[self.c0:self.c0 + self.cl + 2], yv)     # all rows, my columns: y sweep + update
        for side, (kind, rep, values) in bounds.items():
            v = np.atleast_1d(np.asarray(values, float))
            full = v if v.size > 1 else None
            if side in (LEFT, RIGHT):
                self.gx.set_bound(side, kind, rep, full[self.r0:self.r0 + self.rl] if full is not None else v)
                self.gy.set_bound(side, kind, rep, v if full is None else full)      # unused by the y sweep
            else:
                self.gy.set_bound(side, kind, rep, full[self.c0:self.c0 + self.cl] if full is not None else v)
                self.gx.set_bound(side, kind, rep, v if full is None else full)      # unused by the x sweep

    # ---- state (row-sharded view is the public one)
    def set_local_rows(self, grid: torch.Tensor, predict: torch.Tensor):
        """grid / predict: this rank's rows [rl, cols].  Also fills the column-sharded copies (two all-to-alls)."""
        self.gx.set(GRID, grid)
        self.gx.set(PREDICT, predict)
        self.gy.set(GRID, rows_to_cols(grid, self.world))
        self.gy.set(PREDICT, rows_to_cols(predict, self.world))

    def local_rows(self, which: int = GRID) -> torch.Tensor:
        return self.gx.get(which)

    def iteration(self, time: float, nsteps: int = 1):
        zero_copy = hasattr(self.gx, "view")
        for _ in range(nsteps):
            self.gx.xIter(time)                                                  # 1. local rows
            if zero_copy:
                # the local grids work in torch-owned buffers: the collective receives straight into them
                rows_to_cols(self.gx.view(RESULT), self.world, out=self.gy.view(RESULT))           # 2. exchange
                self.gy.yIterUpdate(time)                                                          # 3. local columns
                cols_to_rows(self.gy.view(GRID), self.world, out=self.gx.view(GRID))               # 4. back to rows
                cols_to_rows(self.gy.view(PREDICT), self.world, out=self.gx.view(PREDICT))
            else:
                self.gy.set(RESULT, rows_to_cols(self.gx.get(RESULT), self.world))
                self.gy.yIterUpdate(time)
                self.gx.set(GRID, cols_to_rows(self.gy.get(GRID), self.world))
                self.gx.set(PREDICT, cols_to_rows(self.gy.get(PREDICT), self.world))


class SpikeShardedTwoDGrid:
    """Row-sharded TwoDGrid whose y sweep stays row-sharded: no transposes.

    Rank p owns rows [p R/P, (p+1) R/P) of grid / predict / result -- a contiguous SEGMENT of every column.  One step:

      1. x sweep on the local rows (whole lines, no communication);
      2. all-gather of every rank's first and last predict row (2 x cols doubles per rank): the faces across a cut
         need the neighbour's temperature;
      3. begin: faces, forward elimination of the local segment of every column, condensed to six numbers per column
         (the segment's map  x_first = P E + Q + R L0,  E = C X + D + W L0  in terms of the unknown values L0 above,
         X below and E at its own last row);
      4. all-gather of those 6 x cols doubles per rank -- the only data-path exchange, ~0.8 MB per rank at 16384
         columns against the 3 x 268 MB per rank the transposing driver moves;
      5. end: every rank solves the P-unknown reduced system of every column, back-substitutes its rows and applies
         Grid.update.

    This is the partitioned solver of the single-GPU K5/K6 kernels continued across GPUs (tolerance family, <= 1e-12
    per field against the reference); `RowShardedTwoDGrid` remains the bit-exact variant.  `make_local(xvalues,
    yvalues)` returns the local solver: set_bound / get / set / xIter / predict_edges() -> [2, cols] tensor /
    yIterUpdateBegin(time, halo_prev, halo_next) -> [6, cols] tensor / yIterUpdateEnd(time, gathered, world, rank).
    """

    def __init__(self, make_local: Callable, xvalues, yvalues, bounds: Dict[int, Tuple[int, int, np.ndarray]],
                 world: int, rank: int):
        self.world, self.rank = world, rank
        xv, yv = np.asarray(xvalues, float), np.asarray(yvalues, float)
        self.cols, self.rows = len(xv) - 2, len(yv) - 2
        if self.rows % world:
            raise ValueError("rows must be divisible by the world size")
        if bounds.get(LOW, (0,))[0] == 1 and world > 1:
            raise ValueError("a lower HeatFlow bound is not supported on a sharded grid")
        self.r0, self.rl = block_shard(self.rows, world, rank)
        self.g = make_local(xv, yv[self.r0:self.r0 + self.rl + 2])
        if hasattr(self.g, "use_partitioned"):
            self.g.use_partitioned()    # the split sweep is the partitioned solver; AUTO would pick the exact one for x
        for side, (kind, rep, values) in bounds.items():
            v = np.atleast_1d(np.asarray(values, float))
            if side in (LEFT, RIGHT) and v.size > 1:
                v = v[self.r0:self.r0 + self.rl]
            self.g.set_bound(side, kind, rep, v)
        self._edges = self._gathered = None

    def set_local_rows(self, grid: torch.Tensor, predict: torch.Tensor):
        self.g.set(GRID, grid)
        self.g.set(PREDICT, predict)

    def local_rows(self, which: int = GRID) -> torch.Tensor:
        return self.g.get(which)

    def iteration(self, time: float, nsteps: int = 1):
        P, r = self.world, self.rank
        for _ in range(nsteps):
            mine = self.g.predict_edges()          # predict is frozen until the end of the step
            if self._edges is None:
                self._edges = mine.new_empty((P,) + tuple(mine.shape))
                self._side = torch.cuda.Stream() if mine.is_cuda else None
            # the halo rows travel on a side stream while the x sweep runs (they are only needed by the y sweep)
            ready = None
            if P > 1 and self._side is not None:
                main = torch.cuda.current_stream()
                self._side.wait_stream(main)
                with torch.cuda.stream(self._side):
                    dist.all_gather_into_tensor(self._edges.view(-1, mine.shape[-1]), mine)   # rank-major concatenation
                    ready = self._side.record_event()
                mine.record_stream(self._side)
            elif P > 1:
                dist.all_gather_into_tensor(self._edges.view(-1, mine.shape[-1]), mine)
            else:
                self._edges[0].copy_(mine)
            self.g.xIter(time)
            if ready is not None:
                torch.cuda.current_stream().wait_event(ready)
            prev = self._edges[r - 1, 1] if r > 0 else None
            nxt = self._edges[r + 1, 0] if r + 1 < P else None
            exported = self.g.yIterUpdateBegin(time, prev, nxt)
            if self._gathered is None:
                self._gathered = exported.new_empty((P,) + tuple(exported.shape))
            if P > 1:
                dist.all_gather_into_tensor(self._gathered.view(-1, exported.shape[-1]), exported)
            else:
                self._gathered[0].copy_(exported)
            self.g.yIterUpdateEnd(time, self._gathered, P, r)


class GsLocalGrid:
    """libgs_b200-backed local grid for RowShardedTwoDGrid: CUDA tensors in and out (device-to-device copies on the
    context's stream, which must be torch's current stream so that NCCL and the kernels are ordered)."""

    def __init__(self, gs, ctx, xvalues, yvalues, xdir, ydir, coefs_factory, solver=None):
        x = gs.XDim(xvalues[0], xvalues[1:-1], xvalues[-1], xdir)
        y = gs.YDim(yvalues[0], yvalues[1:-1], yvalues[-1], ydir)
        many = (gs.Many, gs.Temp)
        self.gs, self.g = gs, gs.TwoDGrid(x, y, upp=many, low=many, right=many, left=many, coefs=coefs_factory(ctx),
                                          ctx=ctx)
        if solver is not None:
            self.g.set_solver(solver)
        self.shape = (self.g.rows, self.g.cols)
        # the three arrays live in torch tensors bound into the library (gs_grid2d_bind_device): the collectives
        # of the sharded driver send from / receive into them without staging copies
        self._t = [torch.zeros(self.shape, dtype=torch.float64, device="cuda") for _ in range(3)]
        torch.cuda.current_stream().synchronize()
        for which, t in enumerate(self._t):
            self.g.bind_device(which, t.data_ptr())

    def view(self, which) -> torch.Tensor:
        return self._t[which]

    def use_partitioned(self):
        self.g.set_solver(self.gs.SOLVER_PARTITIONED)

    def set_bound(self, side, kind, rep, values):
        v = np.atleast_1d(np.asarray(values, float))
        n = self.g.cols if side in (UPP, LOW) else self.g.rows
        if v.size not in (1, n):
            v = v[:1]      # a bound the sweep on this layout never reads
        L.check(L.lib().gs_grid2d_set_bound(self.g._h, side, kind, rep, v.ctypes.data_as(L._dp), len(v)))

    def get(self, which) -> torch.Tensor:
        return self._t[which].clone()

    def set(self, which, t: torch.Tensor):
        assert tuple(t.shape) == self.shape and t.dtype == torch.float64 and t.is_cuda
        self._t[which].copy_(t)

    def xIter(self, time):
        self.g.xIter(time)

    def yIterUpdate(self, time):
        self.g.yIterUpdate(time)

    # ---- split y sweep (SpikeShardedTwoDGrid)
    def predict_edges(self) -> torch.Tensor:
        p = self._t[PREDICT]
        return torch.stack((p[0], p[-1]))

    def yIterUpdateBegin(self, time, halo_prev, halo_next) -> torch.Tensor:
        if getattr(self, "_exported", None) is None:
            self._exported = torch.empty((6, self.g.cols), dtype=torch.float64, device="cuda")
        for h in (halo_prev, halo_next):
            assert h is None or (h.is_contiguous() and h.dtype == torch.float64 and h.numel() == self.g.cols)
        self.g.yIterUpdateBegin(time, None if halo_prev is None else halo_prev.data_ptr(),
                                None if halo_next is None else halo_next.data_ptr(), self._exported.data_ptr())
        return self._exported

    def yIterUpdateEnd(self, time, gathered, world, rank):
        assert gathered.is_contiguous() and tuple(gathered.shape) == (world, 6, self.g.cols)
        self.g.yIterUpdateEnd(time, gathered.data_ptr(), world, rank)
