"""Row-sharded 2-D LOD step on N GPUs (torchrun): parity against a single-GPU run and throughput.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
      tools/run_sharded2d.py --size 4096 --coefs hermite --steps 5 [--mode spike|transpose] [--check]
"""
import argparse, json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import gridsplines_b200 as gs
from gridsplines_b200.sharded2d import GsLocalGrid, RowShardedTwoDGrid, SpikeShardedTwoDGrid

ap = argparse.ArgumentParser()
ap.add_argument("--size", type=int, default=4096)
ap.add_argument("--steps", type=int, default=5)
ap.add_argument("--coefs", default="hermite")
ap.add_argument("--solver", type=int, default=0)
ap.add_argument("--check", action="store_true")
ap.add_argument("--mode", default="spike", choices=["spike", "transpose"],
                help="spike: y sweep stays row-sharded, 6 x cols doubles exchanged; transpose: three all-to-alls")
a = ap.parse_args()
world, rank, local = int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
stream = torch.cuda.Stream()
n = a.size
RHO = [0.598, 0.826, 1.121, 1.496, 1.966, 2.547, 3.258, 4.122, 5.156, 6.397, 7.862]
T = [-20.0 + 10.0 * i for i in range(11)]

def coefs_factory(ctx):
    if a.coefs == "const":
        return gs.ConstantCoef(gs.Spline.const(1.5), gs.Spline.const(2.2e6), gs.Spline.const(1.5 / 2.2e6), ctx=ctx)
    k = [(t, 1.0 + 0.1 * q) for t, q in zip(T, RHO)]
    c = [(t, 2.0e6 + 1.0e5 * q) for t, q in zip(T, RHO)]
    al = [(t, kk / cc) for (t, kk), (_, cc) in zip(k, c)]
    return gs.ConstantCoef(gs.Spline.m1Hermite3(k), gs.Spline.m1Hermite3(c), gs.Spline.m1Hermite3(al), ctx=ctx)

# radial x (r_i = 0.05 * 1.0005^(i+1)), ortho y (spacing 0.01), Dirichlet on all four sides
xv = 0.05 * 1.0005 ** np.arange(n + 2)
yv = 0.01 * np.arange(n + 2)
bounds = {0: (0, 1, np.full(n, 7.0)), 1: (0, 1, np.full(n, 7.0)), 2: (0, 1, np.full(n, 7.0)), 3: (0, 1, np.full(n, 30.0))}
with torch.cuda.stream(stream):
    ctx = gs.Context(local, stream=stream.cuda_stream)
    mk = lambda xv_, yv_: GsLocalGrid(gs, ctx, xv_, yv_, gs.RADIAL, gs.ORTHO, coefs_factory, solver=a.solver)
    Driver = SpikeShardedTwoDGrid if a.mode == "spike" else RowShardedTwoDGrid
    sg = Driver(mk, xv, yv, bounds, world, rank)
    g0 = torch.Generator(device="cuda").manual_seed(1234)
    full = 7.0 + torch.rand((n, n), dtype=torch.float64, device="cuda", generator=g0)   # same on every rank
    mine = full[sg.r0:sg.r0 + sg.rl].contiguous()
    sg.set_local_rows(mine, mine.clone())
    sg.iteration(900.0, nsteps=2)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    sg.iteration(900.0, nsteps=a.steps)
    e1.record(stream)
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / a.steps], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    out = {"n": n, "coefs": a.coefs, "solver": a.solver, "mode": a.mode, "n_gpus": world, "ms_per_step": float(ms.item()),
           "updates_per_s": n * n / (float(ms.item()) * 1e-3)}
    if a.check:
        # rank 0 repeats the run on one GPU and compares its rows with everybody's (gathered) rows
        rows = sg.local_rows(0)
        gathered = [torch.empty_like(rows) for _ in range(world)]
        if world > 1:
            dist.all_gather(gathered, rows)
        else:
            gathered = [rows]
        if rank == 0:
            ref = Driver(mk, xv, yv, bounds, 1, 0)
            ref.set_local_rows(full.clone(), full.clone())
            ref.iteration(900.0, nsteps=2 + a.steps)
            want, got = ref.local_rows(0), torch.cat(gathered)
            out["bit_identical_to_single_gpu"] = bool(torch.equal(want.view(torch.int64), got.view(torch.int64)))
            out["max_rel_err_vs_single_gpu"] = float((want - got).abs().max() / want.abs().max())
    if rank == 0:
        print(json.dumps(out), flush=True)
if world > 1:
    dist.destroy_process_group()
