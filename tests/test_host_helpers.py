"""Host-side helper formulas of object TwoDGrid (A/TwoDGrid.scala:1139-1181) in the Python mirror: no GPU needed."""
import math

import pytest


def test_heat_flux_helpers_are_consistent():
    api = pytest.importorskip("gridsplines_b200.api")
    # a steady radial profile T(r) = T0 - Q ln(r / r0) / (2 pi k) carries the same flow Q through every ring
    k, q, r0, t0 = 1.5, 40.0, 0.05, 30.0
    T = lambda r: t0 - q * math.log(r / r0) / (2.0 * math.pi * k)
    for r1 in (0.06, 0.1, 2.0):
        assert api.radialHeatFlow(t0, T(r1), r0, r1, k) == pytest.approx(q, rel=1e-13)
    # discTWithFlow inverts discHeatFlow
    t1 = 21.5
    f = api.discHeatFlow(t0, t1, 0.05, 0.07, k)
    assert api.discTWithFlow(f, t0, 0.05, 0.07, k) == pytest.approx(t1, rel=1e-14)
    # mid-point weights: equal spacing, equal conductivities -> both overloads agree
    a = api.getRadialMidPoint(1.0, 2.0, 3.0, 10.0, 20.0)
    b = api.getRadialMidPoint(1.0, 2.0, 3.0, 10.0, 20.0, k0=2.0, k1=2.0)
    assert a == pytest.approx(b, rel=1e-15) and 10.0 < a < 20.0
    assert a == (10.0 * 1.5 + 20.0 * 2.5) / (1.5 + 2.5)
