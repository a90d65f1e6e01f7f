"""Reduces an `ncu --page raw --csv` dump to the handful of numbers the design argues with -> text summary.
   python tools/ncu_summary.py gpurun_out/X_raw.csv [nodes_per_launch] > profiles/X.txt"""
import csv
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum",
        "smsp__inst_executed_pipe_fp64.sum", "sm__inst_executed_pipe_fp64.sum",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__cycles_elapsed.max", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct"]
rows = list(csv.reader(open(sys.argv[1])))
nodes = float(sys.argv[2]) if len(sys.argv) > 2 else None
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    print("kernel:", r[hdr.index("Kernel Name")])
    val = {}
    for k in KEYS:
        if k in hdr:
            v, u = r[hdr.index(k)], units[hdr.index(k)]
            val[k] = (v, u)
            print(f"  {k:70s} {v} {u}")
    stalls = []
    for i, h in enumerate(hdr):
        if h.startswith("smsp__pcsamp_warps_issue_stalled_") and not h.endswith("_not_issued"):
            try:
                stalls.append((float(r[i]), h.replace("smsp__pcsamp_warps_issue_stalled_", "")))
            except ValueError:
                pass
    tot = sum(s for s, _ in stalls) or 1.0
    print("  stall samples: " + ", ".join(f"{n} {100 * s / tot:.0f}%" for s, n in sorted(stalls, reverse=True)[:6]))
    try:
        def num(k):
            v, u = val[k]
            f = float(v.replace(",", ""))
            return f * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1e-3, "us": 1e-6, "ns": 1e-9,
                        "s": 1.0}.get(u, 1.0)
        t = num("gpu__time_duration.sum")
        b = num("dram__bytes_read.sum") + num("dram__bytes_write.sum")
        print(f"  => DRAM traffic {b / 1e6:.1f} MB per launch, {b / t / 1e9:.0f} GB/s under ncu (cold, serialised)")
        if nodes:
            print(f"  => {b / nodes:.1f} B/node, {num('smsp__inst_executed.sum') * 32 / nodes / 32:.1f} warp-instructions per 32 nodes / 32"
                  f" = {num('smsp__inst_executed.sum') / (nodes / 32):.1f} warp instructions per node-row")
            for k in ("smsp__inst_executed_pipe_fp64.sum", "sm__inst_executed_pipe_fp64.sum"):
                if k in val:
                    print(f"  => FP64 pipe: {num(k) / (nodes / 32):.1f} warp instructions per node-row ({k})")
                    break
    except Exception as e:
        print("  (derived numbers unavailable:", e, ")")
    print()
