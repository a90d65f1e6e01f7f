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


def test_java_number_formatters_known_answers():
    """java.text.NumberFormat (HALF_EVEN on the exact binary value, grouping, no trailing zeros) and Scala's
    f"%.3f" (java.util.Formatter: HALF_UP on Double.toString's digits): documented JDK behaviour, known answers."""
    from gridsplines_b200.api import java_fixed, java_number_format

    nf = lambda v, n=3: java_number_format(v, n)
    assert nf(1234.5678) == "1,234.568"
    assert nf(7.0) == "7"
    assert nf(0.5) == "0.5"
    assert nf(1234567.0) == "1,234,567"
    assert nf(-0.0004) == "-0"
    assert nf(0.0005) == "0.001"        # the double nearest 0.0005 lies above it
    assert nf(0.0625, 3) == "0.062"     # an exact tie: HALF_EVEN
    assert nf(0.1875, 3) == "0.188"     # an exact tie the other way
    assert nf(2.675, 2) == "2.67"       # 2.675 is stored as 2.67499999999999982236431605997495353221893310546875
    assert nf(54.44328715390563) == "54.443"
    assert nf(float("nan")) == "NaN"
    assert java_fixed(1.0005, 3) == "1.001"     # HALF_UP on the digits "1.0005" (the exact value is below the tie)
    assert java_fixed(0.0625, 3) == "0.063"
    assert java_fixed(7.0, 3) == "7.000"
    assert java_fixed(-2.5, 0) == "-3"
    assert java_fixed(12345.678949, 4) == "12345.6789"
    assert java_fixed(-0.0001, 3) == "-0.000"
