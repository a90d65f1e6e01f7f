"""Latency of the two all-gathers of the row-sharded C4 step (2 x cols and 6 x cols doubles per rank) under torchrun."""
import os, sys, json
import torch, torch.distributed as dist
world, rank, local = int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
cols = 16384
out = {}
for rows in (2, 6):
    mine = torch.rand((rows, cols), dtype=torch.float64, device="cuda")
    allg = torch.empty((world, rows, cols), dtype=torch.float64, device="cuda")
    for _ in range(5):
        dist.all_gather_into_tensor(allg.view(-1, cols), mine)
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50):
        dist.all_gather_into_tensor(allg.view(-1, cols), mine)
    e1.record(); torch.cuda.synchronize()
    out[f"allgather_{rows}x{cols}_us"] = e0.elapsed_time(e1) / 50 * 1e3
if rank == 0:
    print(json.dumps({"n_gpus": world, **out}), flush=True)
dist.destroy_process_group()
