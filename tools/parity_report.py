"""Measured parity of every solver family at the CONFIGURED sizes and step counts -> profiles/r2_parity.json.

  python tools/parity_report.py [--out profiles/r2_parity.json] [--quick]

Both metrics are reported everywhere: `field` = max|a-b| / max|b| over the field, `elementwise` = max(|a-b| / |b|).
  C3  4096^2, 100 steps : AUTO (K5, tolerance family) vs GS_SOLVER_THOMAS (K7); K7 vs the oracle, every node, 3 steps
  C4  16384^2, 20 steps : AUTO (= K7) vs the first-draft exact kernels (bit-identical?); GS_SOLVER_PARTITIONED (K6) vs
                          the exact solver after 1 / 5 / 10 / 20 steps; K7 vs the oracle at 4096^2, every node, 20 steps
  C5  4096 x 65536, 20  : AUTO (K5) vs GS_SOLVER_THOMAS; THOMAS vs the oracle on sampled lines (bit-exact?)
  sensitivity           : how far the ORACLE ITSELF moves when `predict` changes by one ulp (2048^2 reduction of C4)
The oracle is the checker here (test infrastructure), never the thing measured."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

import bench as B  # noqa: E402
import cases  # noqa: E402
import gridsplines_b200 as gs  # noqa: E402
from helpers import elem_rel_err, field_rel_err  # noqa: E402
from oracle import oracle as O  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "r2_parity.json"))
ap.add_argument("--quick", action="store_true", help="quarter sizes (smoke run)")
a = ap.parse_args()
Q = 4 if a.quick else 1
stream = torch.cuda.Stream()
ctx = gs.Context(0, stream=stream.cuda_stream)
rep = {"metrics": {"field": "max|a-b| / max|b|", "elementwise": "max(|a-b| / |b|)"}, "tolerance_stated": 1e-12}


def both(a_, b_):
    f, e = B._errs(torch, a_, b_)
    return {"field": f, "elementwise": e}


def bits_equal(a_, b_):
    return bool(torch.equal(a_.view(torch.int64), b_.view(torch.int64)))


def oracle_problem(which, n, field):
    """The same C3 / C4 problem for the oracle (tests/cases.Problem2D)."""
    if which == "C3":
        xv = np.concatenate([[0.0], 0.01 * np.arange(1, n + 1), [0.01 * (n + 1)]])
        b = {0: (0, 1, np.full(n, 20.0)), 1: (0, 1, np.full(n, 4.0)), 2: (1, 1, np.full(n, 0.0)), 3: (1, 1, np.full(n, 50.0))}
        coefs, xdir = "const", 0
    else:
        xv = np.concatenate([[0.05], 0.05 * 1.0005 ** np.arange(1, n + 1), [0.05 * 1.0005 ** (n + 1)]])
        b = {0: (0, 1, np.full(n, 7.0)), 1: (0, 1, np.full(n, 7.0)), 2: (0, 1, np.full(n, 7.0)), 3: (0, 1, np.full(n, 30.0))}
        coefs, xdir = "hermite", 1
    yv = np.concatenate([[0.0], 0.01 * np.arange(1, n + 1), [0.01 * (n + 1)]])
    return cases.Problem2D(xv, yv, xdir, 0, b, coefs=coefs, init=(field, field.copy()))


with torch.cuda.stream(stream):
    # ------------------------------------------------------------------ C3
    n = 4096 // Q
    f = B._field2d(torch, n, "C3")
    fast = B._grid2d(gs, ctx, torch, "C3", n, gs.SOLVER_AUTO, f)
    exact = B._grid2d(gs, ctx, torch, "C3", n, gs.SOLVER_THOMAS, f)
    c3 = {"size": n, "growth": {}}
    done = 0
    for upto in (1, 10, 50, 100):
        fast.iteration(900.0, nsteps=upto - done); exact.iteration(900.0, nsteps=upto - done)
        done = upto
        x, y = B._device_arrays(gs, torch, fast, n), B._device_arrays(gs, torch, exact, n)
        c3["growth"][str(upto)] = {k: both(x[k], y[k]) for k in x}
    c3["auto_K5_vs_exact_K7_after_100_steps"] = c3["growth"]["100"]
    c3["status"] = int(fast.status | exact.status)
    fast.close(); exact.close()
    t0 = time.time()
    p = oracle_problem("C3", n, f.cpu().numpy())
    g, o = p.build_gs(gs), p.build_oracle(O)
    g.set_solver(gs.SOLVER_THOMAS)
    g.iteration(900.0, nsteps=3)
    for _ in range(3):
        o.iteration(900.0)
    snap = cases.snapshot(g)
    c3["exact_K7_vs_oracle_3_steps_every_node"] = {
        k: {"bit_identical": bool(np.array_equal(snap[k].view(np.uint64), v.view(np.uint64))),
            "field": field_rel_err(snap[k], v), "elementwise": elem_rel_err(snap[k], v)} for k, v in cases.snapshot(o).items()}
    c3["oracle_seconds"] = time.time() - t0
    g.close(); del o
    rep["C3"] = c3
    print("C3 done", json.dumps(c3["auto_K5_vs_exact_K7_after_100_steps"]), flush=True)

    # ------------------------------------------------------------------ C4
    n = 16384 // Q
    f = B._field2d(torch, n, "C4")
    c4 = {"size": n}
    g = B._grid2d(gs, ctx, torch, "C4", n, gs.SOLVER_AUTO, f)
    g.iteration(900.0, nsteps=2)
    x = B._device_arrays(gs, torch, g, n)
    ctx.set_option("k7", 0)
    h = B._grid2d(gs, ctx, torch, "C4", n, gs.SOLVER_THOMAS, f)
    h.iteration(900.0, nsteps=2)
    y = B._device_arrays(gs, torch, h, n)
    ctx.set_option("k7", 1)
    c4["auto_K7_vs_first_draft_exact_kernels_2_steps"] = {k: {"bit_identical": bits_equal(x[k], y[k])} for k in x}
    h.close(); del x, y
    # K6 (tolerance family) against the exact solver along the configured 20 steps
    g.copy_from_device(gs.GRID, f.data_ptr()); g.copy_from_device(gs.PREDICT, f.data_ptr())
    part = B._grid2d(gs, ctx, torch, "C4", n, gs.SOLVER_PARTITIONED, f)
    c4["partitioned_K6_vs_exact_K7"] = {}
    done = 0
    for upto in (1, 5, 10, 20):
        g.iteration(900.0, nsteps=upto - done); part.iteration(900.0, nsteps=upto - done)
        done = upto
        x, y = B._device_arrays(gs, torch, part, n), B._device_arrays(gs, torch, g, n)
        c4["partitioned_K6_vs_exact_K7"][str(upto)] = {k: both(x[k], y[k]) for k in x}
        del x, y
    c4["status"] = int(g.status | part.status)
    g.close(); part.close()
    torch.cuda.empty_cache()
    # K7 vs the oracle, every node, the configured 20 steps, at 4096^2
    m = 4096 // Q
    fm = B._field2d(torch, m, "C4")
    t0 = time.time()
    p = oracle_problem("C4", m, fm.cpu().numpy())
    g, o = p.build_gs(gs), p.build_oracle(O)
    g.iteration(900.0, nsteps=20)          # AUTO -> K7
    for _ in range(20):
        o.iteration(900.0)
    snap = cases.snapshot(g)
    c4[f"auto_K7_vs_oracle_20_steps_every_node_{m}"] = {
        k: {"bit_identical": bool(np.array_equal(snap[k].view(np.uint64), v.view(np.uint64))),
            "field": field_rel_err(snap[k], v), "elementwise": elem_rel_err(snap[k], v)} for k, v in cases.snapshot(o).items()}
    c4["oracle_seconds"] = time.time() - t0
    g.close()
    rep["C4"] = c4
    print("C4 done", json.dumps(c4["partitioned_K6_vs_exact_K7"]["20"]), flush=True)

    # ------------------------------------------------------------------ sensitivity of the reference algorithm itself
    m = 2048 // Q
    fm = B._field2d(torch, m, "C4").cpu().numpy()
    pa, pb = oracle_problem("C4", m, fm), oracle_problem("C4", m, fm)
    oa, ob = pa.build_oracle(O), pb.build_oracle(O)
    ob.predict[:] = np.nextafter(ob.predict, np.inf)
    sens = {}
    for step in range(1, 6):
        oa.iteration(900.0); ob.iteration(900.0)
        sens[str(step)] = {"field": field_rel_err(ob.grid, oa.grid), "elementwise": elem_rel_err(ob.grid, oa.grid)}
    rep["oracle_one_ulp_sensitivity_C4_%d" % m] = {
        "what": "the oracle run twice, `predict` of the second run moved up by one ulp before step 1; distance of the two grids",
        "by_step": sens}
    del oa, ob
    print("sensitivity done", json.dumps(sens["5"]), flush=True)

    # ------------------------------------------------------------------ C5
    lines, nn = 4096 // Q, 65536 // Q
    r = np.random.default_rng(B.SEED + 5)
    grid0 = r.uniform(0.0, 30.0, (lines, nn))
    lt, rt = r.uniform(10.0, 30.0, lines), r.uniform(0.0, 10.0, lines)
    fast = B._c5_grid(gs, ctx, lines, nn, gs.SOLVER_AUTO, grid0, lt, rt)
    exact = B._c5_grid(gs, ctx, lines, nn, gs.SOLVER_THOMAS, grid0, lt, rt)
    fast.step(900.0, nsteps=20); exact.step(900.0, nsteps=20)
    c5 = {"lines": lines, "nodes": nn, "auto_K5_vs_exact_after_20_steps": {}}
    ex = {}
    for name, which in (("grid", gs.GRID), ("predict", gs.PREDICT)):
        x, y = fast.download(which), exact.download(which)
        ex[name] = y
        c5["auto_K5_vs_exact_after_20_steps"][name] = {"field": field_rel_err(x, y), "elementwise": elem_rel_err(x, y)}
        del x
    c5["status"] = int(fast.status | exact.status)
    fast.close(); exact.close()
    idx = np.unique(np.concatenate([[0, 1, 31, 32, lines - 1], np.random.default_rng(3).integers(0, lines, 11)]))
    og, op = grid0[idx].copy(), np.full((len(idx), nn), 4.0)
    O.batch1d_step(O.ORTHO, 0.0, 1e-3 * np.arange(1, nn + 1), 1e-3 * (nn + 1), O.Spline.const(1e-6), 1.0, og, op, lt[idx],
                   rt[idx], 900.0, nsteps=20, nthreads=os.cpu_count() or 1)
    c5["exact_vs_oracle_sampled_lines"] = {
        "lines": [int(i) for i in idx],
        "grid_bit_identical": bool(np.array_equal(ex["grid"][idx].view(np.uint64), og.view(np.uint64))),
        "predict_bit_identical": bool(np.array_equal(ex["predict"][idx].view(np.uint64), op.view(np.uint64)))}
    rep["C5"] = c5
    print("C5 done", json.dumps(c5["auto_K5_vs_exact_after_20_steps"]), flush=True)

os.makedirs(os.path.dirname(a.out), exist_ok=True)
with open(a.out, "w") as fjson:
    json.dump(rep, fjson, indent=1)
print("wrote", a.out)
