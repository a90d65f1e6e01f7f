"""Small runs of the mbarrier / cluster / bulk-copy kernels for compute-sanitizer (memcheck, racecheck, synccheck):
K2c and K2 (1-D rings), K5 and K6 (partitioned sweeps, with and without clusters), K7 (exact sweeps).
  compute-sanitizer --tool memcheck python tools/sanitize_cases.py [case ...]
Each case also checks its result against the oracle so that a 'clean' log is a log of a correct run."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases  # noqa: E402
import gridsplines_b200 as gs  # noqa: E402
from helpers import assert_bit_equal, max_rel_err  # noqa: E402
from oracle import oracle as O  # noqa: E402


def k2c():
    kw = dict(nlines=96, n=256, steps=2, spline="const", random_grid=True)
    got, want = cases.run_batch(gs, **kw), cases.run_batch(O, oracle=True, **kw)
    for k in ("grid", "predict"):
        assert_bit_equal(got[k], want[k], "K2c " + k)


def k2():
    kw = dict(nlines=70, n=130, steps=2, spline="hermite", random_grid=True)
    got, want = cases.run_batch(gs, **kw), cases.run_batch(O, oracle=True, **kw)
    for k in ("grid", "predict"):
        assert_bit_equal(got[k], want[k], "K2 " + k)


def _two_d(coefs, solver, cols, rows, combo, exact):
    p = cases.bc_problem(combo, xdir=1, cols=cols, rows=rows, coefs=coefs)
    g, o = p.build_gs(gs), p.build_oracle(O)
    g.set_solver(solver)
    g.iteration(900.0, nsteps=2)
    o.iteration(900.0); o.iteration(900.0)
    got = cases.snapshot(g)
    for k, v in cases.snapshot(o).items():
        if exact:
            assert_bit_equal(got[k], v, f"{coefs} {k}")
        else:
            assert max_rel_err(got[k], v) <= 1e-11, (coefs, k, max_rel_err(got[k], v))
    g.close()


def k5():
    _two_d("const", gs.SOLVER_PARTITIONED, 288, 260, "TTFF", False)     # few strips: cluster path


def k6():
    _two_d("hermite", gs.SOLVER_PARTITIONED, 288, 260, "TTTT", False)


def k7():
    _two_d("hermite", gs.SOLVER_THOMAS, 130, 97, "TFTF", True)
    _two_d("const", gs.SOLVER_THOMAS, 66, 40, "FTFT", True)


ALL = dict(k2c=k2c, k2=k2, k5=k5, k6=k6, k7=k7)
if __name__ == "__main__":
    for name in (sys.argv[1:] or list(ALL)):
        ALL[name]()
        gs.default_context().synchronize()
        print("case", name, "ok", flush=True)
