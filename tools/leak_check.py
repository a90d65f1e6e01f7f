"""Creates and destroys grids (1-D long lines with soft groups, 2-D constants with K5T) sixty times and prints the device
memory in use after the sixth and the last round: equal numbers = nothing leaks."""
import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import gridsplines_b200 as gs
ctx = gs.default_context()
def used():
    free, total = torch.cuda.mem_get_info()
    return (total - free) / 1e6
n = 16384
base = None
for it in range(60):
    g = gs.BatchedOneDGrid(0.0, 1e-3 * np.arange(1, n + 1), 1e-3 * (n + 1), gs.Spline.const(1e-6), 1.0, nlines=64, ctx=ctx)
    g.set_bc(20.0, 4.0)
    g.step(900.0, nsteps=2)
    t = gs.TwoDGrid(gs.XDim(0.0, 0.01 * np.arange(1, 2049), 20.49), gs.YDim(0.0, 0.01 * np.arange(1, 65), 0.65),
                    coefs=gs.ConstantCoef(gs.Spline.const(1.5), gs.Spline.const(2.2e6), gs.Spline.const(1.5 / 2.2e6)),
                    upp=(gs.Many, gs.Temp), low=(gs.Many, gs.Temp), right=(gs.Many, gs.Temp), left=(gs.Many, gs.Temp))
    t.set_solver(gs.SOLVER_PARTITIONED)
    ctx.set_option("k5t", 2)
    t.iteration(900.0, nsteps=2)
    ctx.synchronize()
    g.close(); t.close()
    if it == 5: base = used()
print("MB used after 6 iterations", base, "after 60", used())
