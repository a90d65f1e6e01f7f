// Flattening of the reference's object graphs at the C-ABI boundary (include/gs_b200.h).  NOT compiled in the image this
// repository was written in (no JDK / scalac); the same flattening runs in gridsplines_b200/api.py (Python mirror) and
// gridsplines_b200/host/gridsplines.hpp (C++ mirror), which the tests exercise.
// Citations: A/ = approximation/src/main/scala/approximation/, P/ = piecewise/src/main/scala/piecewise/.
package approximation.b200

import java.util.IdentityHashMap

import approximation._
import approximation.TwoDGrid._
import approximation.b200.GsB200._
import piecewise._

object Flatten {

  def check(status: Int): Unit =
    if (status != GS_OK) throw new RuntimeException(gs_last_error().getString)

  /** TypeDir -> GS_ORTHO / GS_RADIAL (A/TypeDir.scala:23-98) */
  def dir(t: TypeDir): Int = t match {
    case _: Radial => GS_RADIAL
    case _         => GS_ORTHO
  }

  /** JVM exceptions of the path, from the sticky device status word (gs_*_status) */
  def raise(flags: Int): Unit = {
    if ((flags & GS_FLAG_NOT_DOMINANT) != 0)
      throw new AssertionError("Passion is not correct or unstable.")            // A/passion/package.scala:279-282
    if ((flags & GS_FLAG_SPLINE_UNDEFINED) != 0)
      throw new RuntimeException("Spline is not defined at the given point")     // P/intervaltree/AbsITree.scala:475-479
  }

  /** One piece as (kind, x0, c0..c3): P/Const.scala:5, P/Line.scala:6, P/Hermite3.scala:18, P/M1Hermite3.scala:15,
    * P/Lagrange2.scala (quadratic about 0: a cubic with x0 = 0, c3 = 0), P/Lagrange3.scala:33 (cubic about `low`). */
  private def piece(lo: Double, p: PieceFunction): (Int, Double, Array[Double]) = p match {
    case c: Const      => (GS_CONST, 0.0, Array(c.value, 0.0, 0.0, 0.0))
    case l: Line       => (GS_LINE, 0.0, Array(l.intercept, l.slope, 0.0, 0.0))
    case h: Hermite3   => (GS_CUBIC, h.x0, h.coefs.clone())
    case m: M1Hermite3 => (GS_CUBIC, m.x0, m.coefs.clone())
    case q: Lagrange2  =>
      // `coefs` is protected: the three coefficients are recovered from values (exact for a quadratic about 0 is not
      // guaranteed), so a maintainer should expose them; until then sample c0 = q(0) and use the public derivative
      val c0 = q(0.0); val c1 = q.derivative(0.0); val c2 = (q.derivative(1.0) - c1) / 2.0
      (GS_CUBIC, 0.0, Array(c0, c1, c2, 0.0))
    case g: Lagrange3  => throw new UnsupportedOperationException(
      "Lagrange3 keeps coefs / low private[piecewise]: add an accessor, then map to (GS_CUBIC, low, coefs)")
    case other         => throw new UnsupportedOperationException(s"piece type ${other.getClass}")
  }

  /** One gs_spline_create per DISTINCT AlwaysDefinedSpline (by identity); pieces from Spline.sources
    * (P/Spline.scala:115-123), clamps exactly as AlwaysDefinedSpline computes them (A/AlwaysDefinedSpline.scala:5-8:
    * lowerX = spl.lowerBound, upperX = spl.upperBound -- the greatestK quirk included --, lowerY/upperY = spl at them). */
  final class Pool(ctx: gs_ctx) {
    private val seen = new IdentityHashMap[AnyRef, Integer]()
    def id(a: AlwaysDefinedSpline[PieceFunction]): Int = {
      val known = seen.get(a)
      if (known != null) return known.intValue
      val src = a.spl.sources.toArray                                  // ((lo, hi), piece), ascending
      val knots = src.map(_._1._1) :+ src.last._1._2
      val ps = src.map { case ((lo, _), p) => piece(lo, p) }
      val kind = ps.map(_._1).max                                      // Const / Line tables are homogeneous
      val lowerX = a.spl.lowerBound
      val upperX = a.spl.upperBound
      val out = new Array[Int](1)
      check(gs_spline_create(ctx, kind, src.length, knots, ps.map(_._2), ps.flatMap(_._3),
                             lowerX, upperX, a.spl(lowerX), a.spl(upperX), out))
      seen.put(a, out(0))
      out(0)
    }
  }

  /** Coefficients (A/TwoDGrid.scala:782-1137) -> four id maps.  Every member is asked for every index from -1, exactly
    * as Dim.first / general / last would (so VarXCoef's y + 1 indexing, PatchYCoef.apply's x + 1 and the `contains`
    * rule of Patched*Coef come out literally); the map collapses to SINGLE / PER_ROW / PER_COL when the answers allow. */
  def coefficients[P <: PieceFunction](pool: Pool, coefs: Coefficients[P], cols: Int, rows: Int)
      : Array[(Int, Array[Int])] = {
    type Member = (Int, Int) => AlwaysDefinedSpline[PieceFunction]
    val members: Array[Member] = Array(
      (x, y) => coefs(x, y), (x, y) => coefs.thermConductivity(x, y),
      (x, y) => coefs.capacity(x, y), (x, y) => coefs.tempConductivity(x, y))     // GS_SEL_APPLY .. GS_SEL_TEMPCOND
    members.map { m =>
      val dense = Array.tabulate(rows + 1, cols + 1) { (r, c) =>
        try pool.id(m(c - 1, r - 1))
        catch { case _: ArrayIndexOutOfBoundsException => -1 }        // an index the reference itself cannot serve
      }
      val fix = dense.flatten.find(_ >= 0).getOrElse(0)
      for (r <- dense.indices; c <- dense(r).indices if dense(r)(c) < 0) dense(r)(c) = fix
      val perRow = dense.forall(row => row.forall(_ == row(0)))
      val perCol = dense.forall(row => row.sameElements(dense(0)))
      if (perRow && perCol) (GS_MAP_SINGLE, Array(dense(0)(0)))
      else if (perRow) (GS_MAP_PER_ROW, dense.map(_(0)))
      else if (perCol) (GS_MAP_PER_COL, dense(0).clone())
      else (GS_MAP_DENSE, dense.flatten)
    }
  }

  /** Bound (A/TwoDGrid.scala:687-728) -> (kind, representation) */
  def bound(b: Bound): (Int, Int) = {
    val kind = b match { case _: HeatFlow => GS_FLOW; case _ => GS_TEMP }
    val rep = b.bound match { case _: Dense => GS_DENSE; case _ => GS_SPARSE }
    (kind, rep)
  }
}
