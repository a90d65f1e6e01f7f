"""Config C5 (4096 lines x 65536 nodes, constant properties) with the LINES sharded over N GPUs (torchrun): strong
scaling of the long-line solver K5.  No data-path collective: NCCL only carries the barrier and the max of the times.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
      tools/run_c5_sharded.py [--lines 4096] [--nodes 65536] [--steps 10]
"""
import argparse, json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import gridsplines_b200 as gs
from gridsplines_b200.sharding import block_shard

ap = argparse.ArgumentParser()
ap.add_argument("--lines", type=int, default=4096)
ap.add_argument("--nodes", type=int, default=65536)
ap.add_argument("--steps", type=int, default=10)
a = ap.parse_args()
world, rank, local = int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
stream = torch.cuda.Stream()
start, count = block_shard(a.lines, world, rank)
n = a.nodes
r = np.random.default_rng(7 + rank)
with torch.cuda.stream(stream):
    ctx = gs.Context(local, stream=stream.cuda_stream)
    g = gs.BatchedOneDGrid(0.0, 1e-3 * np.arange(1, n + 1), 1e-3 * (n + 1), gs.Spline.const(1e-6), 1.0, nlines=count, ctx=ctx)
    g.upload(gs.GRID, r.uniform(0.0, 30.0, (count, n)))
    lt, rt = r.uniform(10.0, 30.0, count), r.uniform(0.0, 10.0, count)
    g.step(900.0, lt, rt, nsteps=3)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    g.step(900.0, lt, rt, nsteps=a.steps)
    e1.record(stream)
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / a.steps], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        t = float(ms.item())
        print(json.dumps({"config": "C5", "lines": a.lines, "nodes": n, "n_gpus": world, "lines_per_gpu": count,
                          "ms_per_step": t, "updates_per_s": a.lines * n / (t * 1e-3), "scaling": "strong",
                          "status": g.status}), flush=True)
if world > 1:
    dist.destroy_process_group()
