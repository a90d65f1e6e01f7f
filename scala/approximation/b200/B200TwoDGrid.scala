// Drop-in for approximation.TwoDGrid (A/TwoDGrid.scala:16-449) over libgs_b200.so: the same constructor
// (x, y, bounds, coefs), the same builder apply(xDim, yDim)(up)(low)(right)(left)(coefs) (:553-574) and the same members
// user code touches on the stepping path (SURVEY App. D).  State lives on the device; `grid.grid / predict / result`
// are host views that are downloaded when read and uploaded back when user code has written into them.
// See Flatten.scala for the build caveat (no JDK in the image this was written in).
package approximation.b200

import approximation._
import approximation.TwoDGrid._
import approximation.TwoDGrid.Bounds._
import approximation.b200.Flatten.{check, raise}
import approximation.b200.GsB200._
import piecewise._

final class B200TwoDGrid[XType <: TypeDir, YType <: TypeDir, P <: PieceFunction](
    val x: XDim[XType], val y: YDim[YType], val bounds: Bounds, val coefs: Coefficients[P]) extends AutoCloseable {

  private val cols = x.colsNum
  private val rows = y.rowsNum
  private val ctx = new gs_ctx()
  check(gs_ctx_create(0, null, ctx))
  private val h = new gs_grid2d()
  check(gs_grid2d_create(ctx, cols, rows, Flatten.dir(x.t), Flatten.dir(y.t), x.values, y.values, h))
  private val pool = new Flatten.Pool(ctx)
  Flatten.coefficients(pool, coefs, cols, rows).zipWithIndex.foreach { case ((mode, ids), sel) =>
    check(gs_grid2d_set_map(h, sel, mode, ids))
  }
  private val sides: Array[(Int, Bound, Int)] =
    Array((GS_UPP, bounds.upp, cols), (GS_LOW, bounds.low, cols), (GS_RIGHT, bounds.right, rows), (GS_LEFT, bounds.left, rows))
  private val flags = new Array[Int](1)

  /** The Scala Bound objects stay the user-facing containers (`bounds.left.update(q)` etc., :667-685); their values are
    * pushed before every sweep -- 4 small H2D copies per step, as cheap as the reference's own array writes. */
  private def pushBounds(): Unit = sides.foreach { case (side, b, n) =>
    val (kind, rep) = Flatten.bound(b)
    val v = if (rep == GS_DENSE) Array.tabulate(n)(b.get) else Array(b.get(0))
    check(gs_grid2d_set_bound(h, side, kind, rep, v, v.length))
  }
  private def pullBound(side: Int, b: Bound, n: Int): Unit = {
    val v = new Array[Double](n)
    check(gs_grid2d_get_bound(h, side, v, n))
    b.bound match {
      case d: Dense  => d.update(v)
      case s: Sparse => s.update(v(0))
    }
  }
  private def status(): Unit = {
    check(gs_grid2d_status(h, flags))
    if (flags(0) != 0) { gs_grid2d_clear_status(h); raise(flags(0)) }
  }

  /** TwoDGrid.Grid (:461-519) as a view of the device arrays */
  final class DeviceGrid {
    private val n = cols * rows
    private def down(which: Int): Array[Double] = { val a = new Array[Double](n); check(gs_grid2d_download(h, which, a)); a }
    /** copies, not live arrays: write back with grid_=, predict_=, result_= (or use `*=`, updateRow, ...) */
    def grid: Array[Double] = down(GS_GRID)
    def predict: Array[Double] = down(GS_PREDICT)
    def result: Array[Double] = down(GS_RESULT)
    def grid_=(a: Array[Double]): Unit = check(gs_grid2d_upload(h, GS_GRID, a))
    def predict_=(a: Array[Double]): Unit = check(gs_grid2d_upload(h, GS_PREDICT, a))
    def result_=(a: Array[Double]): Unit = check(gs_grid2d_upload(h, GS_RESULT, a))
    def update(): Unit = check(gs_grid2d_update(h))                                   // :474-481
    def avValue: Double = { val o = new Array[Double](1); check(gs_grid2d_av_value(h, o)); o(0) }   // :483-486
    def *=(v: Double): Unit = check(gs_grid2d_fill(h, v))                             // :488-495 grid and predict
  }
  val grid = new DeviceGrid

  def xIter(time: Double): Unit = { pushBounds(); check(gs_grid2d_xiter(h, time)); status() }        // :135-193
  def yIter(time: Double): Unit = { pushBounds(); check(gs_grid2d_yiter(h, time)); status() }        // :195-252
  def iteration(time: Double): Unit = { pushBounds(); check(gs_grid2d_iteration(h, time, 1)); status() }   // :254-258
  /** several steps with frozen bounds: one call, no host round trip in between */
  def iteration(time: Double, nsteps: Int): Unit = { pushBounds(); check(gs_grid2d_iteration(h, time, nsteps)); status() }

  def *=(v: Double): Unit = grid *= v                                                 // :84
  def avValue: Double = grid.avValue                                                  // :60
  def updatePredicted(): Unit = check(gs_grid2d_update_predicted(h))                  // :442-444
  def updateX(f: Double => Double): Unit = check(gs_grid2d_update_x(h, x.range.map(f)))   // :29-42 column c <- f(x_c)
  def updateY(f: Double => Double): Unit = check(gs_grid2d_update_y(h, y.range.map(f)))   // :44-58 row r <- f(y_r)

  def col(colNum: Int, copyTo: Array[Double]): Unit = check(gs_grid2d_get_col(h, GS_GRID, colNum, copyTo))   // :272-281
  def col(colNum: Int): Array[Double] = { val a = new Array[Double](rows); col(colNum, a); a }
  def row(rowNum: Int, copyTo: Array[Double]): Unit = check(gs_grid2d_get_row(h, GS_GRID, rowNum, copyTo))   // :302-310
  def row(rowNum: Int): Array[Double] = { val a = new Array[Double](cols); row(rowNum, a); a }
  def updateColumn(colNum: Int, copyFrom: Array[Double]): Unit = check(gs_grid2d_set_col(h, colNum, copyFrom))   // :291-300
  def updateRow(rowNum: Int, copyFrom: Array[Double]): Unit = check(gs_grid2d_set_row(h, rowNum, copyFrom))      // :318-328
  def left(copyTo: Array[Double]): Unit = col(0, copyTo)                              // :260-270
  def left(colNum: Int, copyTo: Array[Double]): Unit = col(colNum, copyTo)
  def right(colNumFromRight: Int, copyTo: Array[Double]): Unit = col(cols - 1 - colNumFromRight, copyTo)
  def upper(rowNum: Int, copyTo: Array[Double]): Unit = row(rowNum, copyTo)
  def lower(rowFromLast: Int, copyTo: Array[Double]): Unit = row(rows - 1 - rowFromLast, copyTo)

  private def edge(side: Int): Double = { val o = new Array[Double](1); check(gs_grid2d_edge_average(h, side, o)); o(0) }
  def leftAverage: Double = edge(GS_LEFT)                                             // :353-411
  def rightAverage: Double = edge(GS_RIGHT)
  def upperAverage: Double = edge(GS_UPP)
  def lowerAverage: Double = edge(GS_LOW)

  /** :413-440 -- the bound takes the adjacent edge row / column (Dense) or its mean (Sparse); computed on the device
    * and mirrored back into the Scala Bound object so that user code reading `bounds.*.get(i)` sees it. */
  def noHeatFlow(side: BoundSide): Unit = {
    val (s, b, n) = side match {
      case Upper => sides(0)
      case Lower => sides(1)
      case Right => sides(2)
      case Left  => sides(3)
    }
    pushBounds()
    check(gs_grid2d_no_heat_flow(h, s))
    pullBound(s, b, n)
  }

  override def close(): Unit = { gs_grid2d_destroy(h); gs_ctx_destroy(ctx) }

  // the class body of the reference ends with noHeatFlow on all four sides (:446-449)
  noHeatFlow(Upper); noHeatFlow(Lower); noHeatFlow(Right); noHeatFlow(Left)
}

object B200TwoDGrid {
  private def boundBuilder(side: BoundSide, c: Count, b: BoundT)(implicit xDim: XDim[_], yDim: YDim[_]): Bound =
    (c, b) match {                                                                    // :542-551
      case (One, Temp)  => Temperature()
      case (Many, Temp) => Temperature(side, xDim, yDim)
      case (One, Flow)  => HeatFlow()
      case (Many, Flow) => HeatFlow(side, xDim, yDim)
    }

  /** TwoDGrid(xDim, yDim)(upCount, upBound)(lowCount, lowBound)(rightCount, rightBound)(leftCount, leftBound)(coefs) */
  def apply[X <: TypeDir, Y <: TypeDir, P <: PieceFunction](xDim: XDim[X], yDim: YDim[Y])
      (upCount: Count, upBound: BoundT)(lowCount: Count, lowBound: BoundT)
      (rightCount: Count, rightBound: BoundT)(leftCount: Count, leftBound: BoundT)
      (coefs: Coefficients[P]): B200TwoDGrid[X, Y, P] = {
    implicit val x: XDim[X] = xDim
    implicit val y: YDim[Y] = yDim
    val bounds = Bounds(boundBuilder(Upper, upCount, upBound), boundBuilder(Lower, lowCount, lowBound),
                        boundBuilder(Right, rightCount, rightBound), boundBuilder(Left, leftCount, leftBound))
    new B200TwoDGrid(xDim, yDim, bounds, coefs)
  }
}
