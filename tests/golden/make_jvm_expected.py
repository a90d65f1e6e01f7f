"""Writes tests/golden/jvm_expected.txt: the bit patterns scala/OracleCheck.scala must print when the real reference
runs on a JVM -- produced here by the ORACLE (oracle/gs_oracle.c), which is the thing being pinned.  Same line format:
`name count hex hex ...` (java.lang.Long.toHexString of doubleToLongBits).

    python tests/golden/make_jvm_expected.py
"""
import math
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases  # noqa: E402
from oracle import oracle as O  # noqa: E402

lines = []


def bits(name, a):
    a = np.ascontiguousarray(np.asarray(a, dtype=np.float64).ravel())
    lines.append(name + " " + str(a.size) + " " + " ".join(format(int(v), "x") for v in a.view(np.uint64)))


# C1 / R2
s = cases.c1_setup()
g = O.OneDGrid(s["leftX"], s["rangeX"], s["rightX"], O.Spline.const(1e-6), s["sigma"], direction=O.RADIAL)
done = 0
for upto in (1, 10, 1000):
    while done < upto:
        g.step(s["time"], s["leftT"], s["rightT"])
        done += 1
    bits(f"c1_grid_after_{upto}", g.grid)


# R3 - R5
def small(left, upper):
    xv, yv = cases.small2d_dims()
    b = {side: (0, 0, [1.0]) for side in range(4)}
    b[3] = (0, 0, [left])
    b[0] = (0, 0, [upper])
    p = cases.Problem2D(xv, yv, 1, 0, b, coefs="unit", init=(np.ones((10, 8)), np.ones((10, 8))))
    return p.build_oracle(O)


o = small(4.0, 1.0); o.xIter(900.0); o.update(); bits("r3_grid", o.grid); bits("r3_predict", o.predict)
o = small(1.0, 4.0); o.yIter(900.0); o.update(); bits("r4_grid", o.grid); bits("r4_predict", o.predict)
o = small(4.0, 1.0); o.iteration(900.0); bits("r5_grid", o.grid); bits("r5_predict", o.predict)

# GridTest
p = cases.gridtest_problem()
o = p.build_oracle(O)
o.fill(7.0)
for side in range(4):
    o.set_bound(side, p.bounds[side][0], 1, np.full(len(p.bounds[side][2]), 7.0))
snap = cases.run_gridtest(o, False)
bits("gridtest_grid", snap["grid"])
bits("gridtest_predict", snap["predict"])

# spline evaluation
T = [100.0 + 10.0 * i for i in range(11)]
RHO = [0.598, 0.826, 1.121, 1.496, 1.966, 2.547, 3.258, 4.122, 5.156, 6.397, 7.862]
O.set_ascending_fold(False)        # the reference's own tree-order fold
a = O.AlwaysDefinedSpline(O.Spline.m1Hermite3(list(zip(T, RHO))))
xs = [95.0 + 2.75 * i for i in range(41)]
bits("hermite_apply", [a(x) for x in xs])
bits("hermite_average", [a.average(xs[i], xs[i + 1]) for i in range(40)])
O.set_ascending_fold(True)

with open(os.path.join(HERE, "jvm_expected.txt"), "w") as f:
    f.write("\n".join(lines) + "\n")
print("wrote", len(lines), "lines")
