import torch, time
n = 2164260864 // 8
h_in = torch.empty(n, dtype=torch.float64).pin_memory()
h_out = torch.empty(n, dtype=torch.float64).pin_memory()
d_in = torch.empty(n, dtype=torch.float64, device="cuda")
d_out = torch.empty(n, dtype=torch.float64, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(h2d, d2h, chunk=None):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3):
        if chunk is None:
            if h2d:
                with torch.cuda.stream(s1): d_in.copy_(h_in, non_blocking=True)
            if d2h:
                with torch.cuda.stream(s2): h_out.copy_(d_out, non_blocking=True)
        else:
            for o in range(0, n, chunk):
                if h2d:
                    with torch.cuda.stream(s1): d_in[o:o+chunk].copy_(h_in[o:o+chunk], non_blocking=True)
                if d2h:
                    with torch.cuda.stream(s2): h_out[o:o+chunk].copy_(d_out[o:o+chunk], non_blocking=True)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 3
    return dt
for name, a in (("H2D only", (True, False)), ("D2H only", (False, True)), ("both", (True, True))):
    dt = run(*a)
    print(f"{name:10s} {dt*1e3:7.2f} ms  {n*8/dt/1e9:6.1f} GB/s per direction")
for mb in (16, 64, 256):
    dt = run(True, True, mb * (1 << 20) // 8)
    print(f"both, {mb} MB chunks {dt*1e3:7.2f} ms  {n*8/dt/1e9:6.1f} GB/s per direction")
