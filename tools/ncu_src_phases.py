"""Reduces an `ncu --page source --csv --print-source sass` dump (tools/ncu_capture.sh with SRC=1) to warp-sample shares per
block of SASS instructions plus the most-sampled instructions -- where a warp-specialised kernel spends its time.
   python tools/ncu_src_phases.py gpurun_out/X_sass.csv [bucket=500] [top=12] > profiles/X_phases.txt"""
import csv
import sys

csv.field_size_limit(10 ** 9)
rows = list(csv.reader(open(sys.argv[1])))
bucket = int(sys.argv[2]) if len(sys.argv) > 2 else 500
top = int(sys.argv[3]) if len(sys.argv) > 3 else 12
secs = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
seen = set()
for si, start in enumerate(secs):
    name = rows[start][1]
    if name in seen:      # ncu repeats every kernel's table (one per view)
        continue
    seen.add(name)
    end = secs[si + 1] if si + 1 < len(secs) else len(rows)
    hdr, data = rows[start + 1], rows[start + 2:end]
    ix = {h: i for i, h in enumerate(hdr)}
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    val = lambda r, c: int(r[ix[c]] or 0) if len(r) > ix[c] else 0
    smp = [val(r, "# Samples") for r in data]
    exe = [val(r, "Instructions Executed") for r in data]
    tot = sum(smp) or 1
    print(f"kernel: {name}\n  SASS instructions {len(data)}, warp samples {tot}, warp instructions executed {sum(exe) / 1e6:.1f} M")
    print("  samples per block of instructions (a warp is sampled whether it issues or stalls: share of warp-time):")
    for b in range(0, len(data), bucket):
        s, e = sum(smp[b:b + bucket]), sum(exe[b:b + bucket])
        if s:
            print(f"    [{b:5d},{b + bucket:5d})  {s:7d}  {100 * s / tot:5.1f} %   {e / 1e6:8.1f} M instructions")
    print(f"  {top} most-sampled instructions:")
    for s, n in sorted(((s, n) for n, s in enumerate(smp)), reverse=True)[:top]:
        r = data[n]
        st = sorted(((val(r, c), c[6:]) for c in stalls), reverse=True)[:2]
        print(f"    #{n:5d} {s:7d} {100 * s / tot:5.1f} %  executed {exe[n]:9d}  {r[ix['Source']].strip()[:58]:58s} {st[0][1]} {st[0][0]}, {st[1][1]} {st[1][0]}")
    print()
