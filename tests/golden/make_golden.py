"""Generates tests/golden/*.npz from the CPU oracle (oracle/gs_oracle.c).

The reference is Scala and cannot run in the build container (no JVM), so these fixtures are outputs of
the literal restatement; they pin the oracle against drift and give the GPU tests a second,
oracle-independent comparison at run time.  Regenerate with:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import cases  # noqa: E402


def _c1(O):
    return cases.run_c1(O, steps=1000, oracle=True)


def _batch_const(O):
    return cases.run_batch(O, 37, 256, 5, spline="const", oracle=True)


def _batch_hermite(O):
    return cases.run_batch(O, 19, 64, 4, spline="hermite", random_grid=True, oracle=True)


def _grid2d(combo, xdir, coefs, steps=3, ascending=True):
    def run(O):
        p = cases.bc_problem(combo, xdir=xdir, coefs=coefs)
        g = p.build_oracle(O, ascending=ascending)
        for _ in range(steps):
            g.iteration(900.0)
        return cases.snapshot(g)

    return run


def _gridtest(O):
    p = cases.gridtest_problem()
    g = p.build_oracle(O, ascending=False)
    for side in range(4):
        g.noHeatFlow(side)
    g.fill(7.0)
    return cases.run_gridtest(g, False)


CASES = {
    "c1_radial_200x1000": _c1,
    "batch_const_37x256x5": _batch_const,
    "batch_hermite_19x64x4": _batch_hermite,
    "grid2d_TTFF_ortho_const": _grid2d("TTFF", 0, "const"),
    "grid2d_FFTT_radial_const": _grid2d("FFTT", 1, "const"),
    "grid2d_TTTT_radial_hermite": _grid2d("TTTT", 1, "hermite"),
    "gridtest_137x45x195": _gridtest,
}

if __name__ == "__main__":
    from oracle import oracle as O

    for name, fn in CASES.items():
        out = fn(O)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        print(name, {k: v.shape for k, v in out.items()})
