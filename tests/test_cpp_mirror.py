"""The C++ host mirror (gridsplines_b200/host/gridsplines.hpp) over the C ABI: compiles on the CPU box, where its
device-free parts (number formatters, spline factories) must agree with the Python mirror; on the GPU box it runs config
C1, two README known answers and the GridTest scenario and must print the oracle's bits, and its other classes (spline
algebra, per-node splines, per-line ends, ragged lines, the Coefficients hierarchy, text output) must give what the
Python mirror gives for the same inputs, bit for bit and byte for byte."""
import os
import struct
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "host_mirror_test.cpp")


def _compile(out):
    lib_dir = os.path.join(ROOT, "gridsplines_b200")
    cmd = ["g++", "-std=c++17", "-O2", "-ffp-contract=off", SRC, "-o", out, "-L" + lib_dir, "-lgs_b200",
           "-Wl,-rpath," + lib_dir]
    subprocess.check_call(cmd)


def test_cpp_mirror_compiles_and_links(tmp_path):
    _compile(str(tmp_path / "host_mirror_test"))


@pytest.mark.gpu
def test_cpp_mirror_reproduces_kats(tmp_path):
    exe = str(tmp_path / "host_mirror_test")
    _compile(exe)
    out = subprocess.check_output([exe], text=True)
    vals = {l.split()[0]: l.split()[1:] for l in out.splitlines()}
    assert float(vals["c1_grid0"][0]) == 19.163547799904077 and vals["c1_grid0"][1] == "0x403329de44c3ebb7"  # KAT R2
    assert float(vals["c1_grid99"][0]) == 6.303223056805375 and float(vals["c1_grid199"][0]) == 4.014707461585766
    assert float(vals["readme_135"][0]) == 1.718125                       # README.md:85-86
    assert abs(float(vals["readme_area_150_185"][0]) - 139.148203125) < 1e-12 * 139.148203125  # README.md:99-100
    assert vals["gridtest_dims"] == ["137", "45"]
    assert float(vals["gridtest_grid0"][0]) == 54.44328715390563          # KAT R1
    assert float(vals["gridtest_grid1"][0]) == 53.94901954479779
    assert vals["gridtest_slices_ok"] == ["1"]


RHO = [0.598, 0.826, 1.121, 1.496, 1.966, 2.547, 3.258, 4.122, 5.156, 6.397, 7.862]
XS = [0.0, -0.0, 1.0, -1.5, 2.5, 0.125, 0.0005, -0.0005, 1234567.891, 0.1 + 0.2, 1e-7, 54.44328715390563, 999.9995,
      1e15 / 3.0, -2.675, 8.345, 0.045, 1.0005, 123456.7895]


def _parse(out):
    return {l.split(" ", 1)[0]: (l.split(" ", 1)[1] if " " in l else "") for l in out.splitlines()}


def _hex(a):
    return " ".join(struct.pack(">d", float(v)).hex() for v in np.asarray(a, dtype=np.float64).ravel())


def _text(vals, key):
    return bytes.fromhex(vals[key]).decode("utf-8")


def _check_spline(vals, name, spl):
    assert int(vals[name + "_kind"]) == spl.kind, name
    assert vals[name + "_knots"] == _hex(spl.knots), name + " knots"
    assert vals[name + "_x0"] == _hex(spl.x0), name + " x0"
    assert vals[name + "_coefs"] == _hex(spl.coefs), name + " coefs"


def _check_host_only(vals):
    api = pytest.importorskip("gridsplines_b200.api")
    for i, x in enumerate(XS):
        assert vals[f"nf3_{i}"] == api.java_number_format(x, 3), (x, "NumberFormat 3")
        assert vals[f"nf2_{i}"] == api.java_number_format(x, 2), (x, "NumberFormat 2")
        assert vals[f"fx3_{i}"] == api.java_fixed(x, 3), (x, "%.3f")
        assert vals[f"fx4_{i}"] == api.java_fixed(x, 4), (x, "%.4f")
    assert vals["nf_nan"] == "NaN" and vals["fx_inf"] == "-Infinity"
    pts = [(0.0, 1.0), (1.0, 3.0), (2.5, 2.0), (4.0, 5.0), (6.0, 4.0), (7.5, 4.5)]
    rho = [(100.0 + 10.0 * k, RHO[k]) for k in range(11)]
    S = api.Spline
    for name, spl in (("lag2", S.lagrange2(pts)), ("lag3", S.lagrange3(pts)), ("m1h", S.m1Hermite3(rho)),
                      ("fh", S.fHermite3(rho)), ("lines", S.lines(rho)), ("const", S.const(2.5))):
        _check_spline(vals, name, spl)
    x = api.XDim(0.05, [0.1, 0.2004, 0.30049999, 0.4], 0.5, api.RADIAL)
    assert _text(vals, "dim_str") == api._dim_to_string(x, "\n")


def test_cpp_mirror_host_only_parts_agree_with_the_python_mirror(tmp_path):
    exe = str(tmp_path / "host_mirror_test")
    _compile(exe)
    _check_host_only(_parse(subprocess.check_output([exe, "cpu"], text=True)))


@pytest.mark.gpu
def test_cpp_mirror_device_sections_agree_with_the_python_mirror(tmp_path):
    import gridsplines_b200 as gs

    exe = str(tmp_path / "host_mirror_test")
    _compile(exe)
    vals = _parse(subprocess.check_output([exe], text=True))
    _check_host_only(vals)
    ctx = gs.default_context()
    S = gs.Spline
    T = [-20.0 + 10.0 * i for i in range(11)]
    k = S.m1Hermite3([(t, 1.0 + 0.1 * q) for t, q in zip(T, RHO)])
    cpts = [(t, 2.0e6 + 1.0e5 * q) for t, q in zip(T, RHO)]
    c = S.m1Hermite3(cpts)
    _check_spline(vals, "quot", k / c)
    _check_spline(vals, "sum", k.combine(S.lines(cpts), lambda a, b: a + b, factory=S.fHermite3))
    rng = [0.01 * (i + 1) for i in range(40)]
    a1 = gs.AlwaysDefinedSpline(S.m1Hermite3([(-20.0, 1.0e-6), (10.0, 1.3e-6), (40.0, 1.9e-6), (80.0, 2.2e-6)]))
    a2 = gs.AlwaysDefinedSpline(S.const(1.5e-6))
    a3 = gs.AlwaysDefinedSpline(S.lines([(-20.0, 1.1e-6), (30.0, 1.4e-6), (80.0, 2.0e-6)]))
    pool = [a1, a2, a3]
    # per-node pairs
    g = gs.OneDGrid(0.0, rng, 0.41, [(pool[i % 3], pool[(i + 1) % 3]) for i in range(40)], 1.0, gs.ORTHO)
    g.grid = [5.0 + 0.37 * i for i in range(40)]
    for _ in range(3):
        g.step(900.0, 20.0, 4.0)
    assert vals["pernode_grid"] == _hex(g.grid) and vals["pernode_predict"] == _hex(g.predict)
    assert int(vals["pernode_status"]) == g.status == 0
    # per-line end temperatures, both layouts
    b = gs.BatchedOneDGrid(0.005, rng, 0.41, a1, 1.0, gs.RADIAL, 5)
    b.upload(gs.GRID, np.array([3.0 + 0.11 * (i % 37) for i in range(200)]).reshape(5, 40))
    b.step(900.0, [10.0 + 3.0 * l for l in range(5)], [1.0 + l for l in range(5)], nsteps=2)
    assert vals["batched_grid"] == _hex(b.download(gs.GRID))
    assert vals["batched_interleaved"] == _hex(b.download(gs.GRID, gs.INTERLEAVED))
    # ragged lines
    lines = [[0, 1, 2, 3, 4, 5, 6], [10, 12, 14, 16, 18], [29, 28, 27, 26, 25, 24, 23, 22, 21]]
    pd = []
    for l in lines:
        out = np.empty(2 * len(l))
        r = np.array([0.02 * (i + 1) for i in range(len(l))])
        gs._lib.check(gs._lib.lib().gs_predef(ctx._h, gs.ORTHO, 0.0, r.ctypes.data_as(gs._lib._dp), len(l),
                                              0.02 * (len(l) + 1), 1.0, out.ctypes.data_as(gs._lib._dp)))
        pd.append(out)
    rg = gs.RaggedLines(30, lines, pd, a3)
    init = [2.0 + 0.5 * i for i in range(30)]
    rg.grid = init
    rg.predicted = init
    rg.iteration(900.0, [20.0, 15.0, 12.0], [4.0, 3.0, 1.0])
    assert vals["ragged_result"] == _hex(rg.result) and int(vals["ragged_status"]) == rg.status == 0
    # Coefficients hierarchy + text output
    x = gs.XDim(0.0, lambda v: v + 0.01, 0.125, gs.ORTHO)
    y = gs.YDim(0.0, lambda v: v + 0.02, 0.21, gs.ORTHO)
    cols, rows = x.colsNum, y.rowsNum
    assert vals["patched_dims"] == f"{cols} {rows}"
    mk = lambda scale: gs.AlwaysDefinedSpline(S.m1Hermite3([(-20.0 + 10.0 * i, scale * (1.0 + 0.1 * RHO[i])) for i in range(11)]))
    th, cp, tc, pth, pcp, ptc = [], [], [], [], [], []
    for r in range(rows + 1):
        th.append(mk(1.0 + 0.05 * r)); cp.append(mk(2.0e6)); tc.append(mk((1.0 + 0.05 * r) / 2.0e6))
        pth.append(mk(3.0)); pcp.append(mk(1.0e6)); ptc.append(mk(3.0e-6))
    coefs = gs.PatchedXCoef(gs.VarYCoef(th, cp, tc), gs.PatchXCoef(pth, pcp, ptc, (2, 7), (1, 6)))
    t = gs.TwoDGrid(x, y, upp=(gs.Many, gs.Temp), low=(gs.One, gs.Temp), right=(gs.Many, gs.Temp),
                    left=(gs.Many, gs.Flow), coefs=coefs)
    t.set_solver(gs.SOLVER_THOMAS)
    init = np.array([7.0 + 0.013 * float((i * 7919) % 101) for i in range(cols * rows)]).reshape(rows, cols)
    t.grid.grid = init
    t.grid.predict = init
    t.bounds.upp.update(20.0); t.bounds.low.update(4.0); t.bounds.right.update(9.5); t.bounds.left.update(50.0)
    t.iteration(900.0, nsteps=2)
    assert vals["patched_grid"] == _hex(t.grid.grid) and vals["patched_predict"] == _hex(t.grid.predict)
    assert vals["patched_left_bound"] == _hex(t.bounds.left.get()) and int(vals["patched_status"]) == t.status == 0
    assert _text(vals, "patched_tostring") == gs.api.twod_to_string(t, "\n")
    import io

    w = io.StringIO()
    gs.api.twod_write(t, w, "\n")
    assert _text(vals, "patched_write") == w.getvalue()
    # ShardedTwoDGrid on one device
    xs = gs.XDim(0.0, [0.01 * (i + 1) for i in range(64)], 0.65, gs.ORTHO)
    ys = gs.YDim(0.0, [0.02 * (i + 1) for i in range(48)], 0.98, gs.ORTHO)
    many = (gs.Many, gs.Temp)
    sh = gs.ShardedTwoDGrid(xs, ys, [0], upp=many, low=many, right=many, left=many)
    sh.set_constant_coef(S.const(1.5), S.const(2.2e6), S.const(1.5 / 2.2e6))
    for side, v in ((gs.UPP, 20.0), (gs.LOW, 4.0), (gs.RIGHT, 9.0), (gs.LEFT, 30.0)):
        sh.set_bound(side, v)
    init = np.array([7.0 + 0.01 * float((i * 131) % 97) for i in range(64 * 48)]).reshape(48, 64)
    sh.upload(gs.GRID, init)
    sh.upload(gs.PREDICT, init)
    sh.iteration(900.0, nsteps=2)
    assert vals["sharded_grid"] == _hex(sh.download(gs.GRID)) and vals["sharded_status"] == f"{sh.status} 1"
