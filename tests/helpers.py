"""Shared builders: the same problem set up in the oracle (O.*) and in gridsplines_b200 (gs.*)."""
import numpy as np

README_T = [100.0 + 10.0 * i for i in range(11)]
README_RHO = [0.598, 0.826, 1.121, 1.496, 1.966, 2.547, 3.258, 4.122, 5.156, 6.397, 7.862]


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float64).view(np.uint64)


def assert_bit_equal(a, b, what=""):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    ba, bb = bits(a), bits(b)
    if not np.array_equal(ba, bb):
        bad = np.flatnonzero(ba.ravel() != bb.ravel())
        i = bad[0]
        raise AssertionError(
            f"{what}: {bad.size} of {a.size} values differ; first at flat index {i}: "
            f"{a.ravel()[i]!r} ({ba.ravel()[i]:#x}) vs {b.ravel()[i]!r} ({bb.ravel()[i]:#x})"
        )


def field_rel_err(a, b):
    """Max relative error per FIELD (one reading of BASELINE.json's "<= 1e-12 max relative error per field"):
    max |a - b| over the field divided by max |b| over the field."""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    scale = float(np.max(np.abs(b)))
    return float(np.max(np.abs(a - b))) / max(scale, 1e-300)


def elem_rel_err(a, b):
    """Max ELEMENT-WISE relative error: max over the field of |a - b| / |b| (the stricter reading; an exact zero in
    the reference compares absolutely against the field's scale)."""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    den = np.abs(b)
    den = np.where(den == 0.0, max(float(np.max(den)), 1e-300), den)
    return float(np.max(np.abs(a - b) / den))


max_rel_err = field_rel_err   # the name the round-1 tests use

# every tolerance comparison of a test session: (what, field, elem, tol_field, tol_elem); tests/conftest.py writes the
# list to gpurun_out/parity_tests.json at the end of a GPU session
PARITY_LOG = []


def check_tol(got, want, what, tol_field, tol_elem=None):
    """Asserts BOTH metrics against FIXED tolerances, with the measured values in the message and in PARITY_LOG."""
    fe, ee = field_rel_err(got, want), elem_rel_err(got, want)
    PARITY_LOG.append(dict(what=what, field=fe, elem=ee, tol_field=tol_field, tol_elem=tol_elem))
    assert fe <= tol_field, f"{what}: field-normalised error {fe:.3e} > {tol_field:.1e} (element-wise {ee:.3e})"
    if tol_elem is not None:
        assert ee <= tol_elem, f"{what}: element-wise error {ee:.3e} > {tol_elem:.1e} (field-normalised {fe:.3e})"
    return fe, ee


def both_splines(O, gs, kind, *args):
    """(oracle spline, gs spline) built by the same factory."""
    return getattr(O.Spline, kind)(*args), getattr(gs.Spline, kind)(*args)


def table_points(lo=-20.0, hi=80.0, n=11, base=1.0, scale=0.1):
    """y_i = base + scale * rho_i over n knots in [lo, hi]."""
    xs = np.linspace(lo, hi, n)
    return [(float(x), base + scale * README_RHO[i % len(README_RHO)]) for i, x in enumerate(xs)]


def hermite_coefs(O, gs, ctx=None):
    """k(T), c(T) M1Hermite3 tables and alpha = k / c rebuilt as M1Hermite3 (Spline./, P/Spline.scala:207-210)."""
    kp = table_points(base=1.0, scale=0.1)
    cp = table_points(base=2.0e6, scale=1.0e5)
    ap = [(x, k / c) for (x, k), (_, c) in zip(kp, cp)]
    ok, oc, oa = (O.Spline.m1Hermite3(p) for p in (kp, cp, ap))
    gk, gc, ga = (gs.Spline.m1Hermite3(p) for p in (kp, cp, ap))
    return (ok, oc, oa), (gk, gc, ga)


def uneven_coefs(O, gs, factory="lines"):
    """k(T), c(T), alpha(T) on UNEVENLY spaced knots, as piecewise-linear tables (Spline.lines) or M1Hermite3: the
    tables the straight-line face path of K6 does not take."""
    T = [-25.0, -11.0, -4.0, 3.0, 9.5, 21.0, 30.0, 44.0, 61.0, 90.0]
    kp = [(t, 1.0 + 0.02 * i * i) for i, t in enumerate(T)]
    cp = [(t, 2.0e6 + 3.0e4 * i) for i, t in enumerate(T)]
    ap = [(x, k / c) for (x, k), (_, c) in zip(kp, cp)]
    mk_o, mk_g = getattr(O.Spline, factory), getattr(gs.Spline, factory)
    return tuple(mk_o(p) for p in (kp, cp, ap)), tuple(mk_g(p) for p in (kp, cp, ap))
