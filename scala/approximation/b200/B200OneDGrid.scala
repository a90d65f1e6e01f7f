// Drop-in for approximation.OneDGrid (A/OneDGrid.scala:14-79) over libgs_b200.so: same constructor arguments, the same
// public `grid` array, the same `step(time, leftT, rightT)`.  See Flatten.scala for the build caveat.
package approximation.b200

import approximation._
import approximation.b200.Flatten.{check, raise}
import approximation.b200.GsB200._
import piecewise._

final class B200OneDGrid[Dir <: TypeDir](val leftX: Double, val rangeX: Array[Double], val rightX: Double,
    val conductivities: Array[Array[AlwaysDefinedSpline[PieceFunction]]], val sigma: Double)(implicit dir: Dir)
    extends AutoCloseable {

  /** `val grid = Array.fill(rangeX.length)(4.0)` (:20): public and mutable, user code pokes it between steps */
  val grid: Array[Double] = Array.fill(rangeX.length)(4.0)

  private val ctx = new gs_ctx()
  check(gs_ctx_create(0, null, ctx))
  private val pool = new Flatten.Pool(ctx)
  private val ids: Array[Int] = conductivities.flatMap(pair => Array(pool.id(pair(0)), pool.id(pair(1))))
  private val single = ids.distinct.length == 1
  private val h = new gs_grid1d()
  // preDef = dir.preDef(leftX, rangeX, rightX, sigma) (:24) is computed on the device (gs_predef's kernel)
  check(gs_grid1d_create(ctx, 1L, rangeX.length, Flatten.dir(dir), leftX, rangeX, rightX, sigma,
    if (single) GS_SPLINE_SINGLE else GS_SPLINE_PER_NODE, if (single) ids.take(1) else ids, null, h))
  private val flags = new Array[Int](1)

  /** (:26-34): iteration + predictor + grid <- result.  `predict` and `result` are private in the reference and live
    * on the device; `grid` travels both ways because it is public. */
  def step(time: Double, leftT: Double, rightT: Double): Unit = {
    check(gs_grid1d_step_host(h, time, 1, grid, Array(leftT), Array(rightT), 0))
    check(gs_grid1d_status(h, flags))
    if (flags(0) != 0) { gs_grid1d_clear_status(h); raise(flags(0)) }
  }

  override def close(): Unit = { gs_grid1d_destroy(h); gs_ctx_destroy(ctx) }
}

object B200OneDGrid {
  /** companion apply of the reference (:66-79): one conductivity spline for every node */
  def apply[Dir <: TypeDir](leftX: Double, rangeX: Array[Double], rightX: Double,
                            conductivity: Spline[PieceFunction], sigma: Double)(implicit dir: Dir): B200OneDGrid[Dir] = {
    val ads = new AlwaysDefinedSpline(conductivity)
    new B200OneDGrid[Dir](leftX, rangeX, rightX, Array.fill(rangeX.length)(Array(ads, ads)), sigma)
  }
}

/** Many independent OneDGrids that share geometry and conductivity (BASELINE config C2 / C5): the batched entry point
  * the reference only hints at (passion.iteration overload 3, A/passion/package.scala:141-227).  `grids` is
  * line-major: line l owns grids(l * n until (l + 1) * n). */
final class B200BatchedOneDGrid[Dir <: TypeDir](nlines: Long, leftX: Double, rangeX: Array[Double], rightX: Double,
    conductivity: Spline[PieceFunction], sigma: Double)(implicit dir: Dir) extends AutoCloseable {
  private val ctx = new gs_ctx()
  check(gs_ctx_create(0, null, ctx))
  private val id = new Flatten.Pool(ctx).id(new AlwaysDefinedSpline(conductivity))
  private val h = new gs_grid1d()
  check(gs_grid1d_create(ctx, nlines, rangeX.length, Flatten.dir(dir), leftX, rangeX, rightX, sigma,
    GS_SPLINE_SINGLE, Array(id), null, h))
  private val flags = new Array[Int](1)

  /** state on the device between calls */
  def upload(grids: Array[Double]): Unit = check(gs_grid1d_upload(h, GS_GRID, GS_LINE_MAJOR, grids))
  def download(grids: Array[Double]): Unit = check(gs_grid1d_download(h, GS_GRID, GS_LINE_MAJOR, grids))
  def step(time: Double, leftT: Array[Double], rightT: Array[Double], nsteps: Int = 1): Unit = {
    check(gs_grid1d_set_bc(h, leftT, rightT, if (leftT.length == 1) 0 else 1))
    check(gs_grid1d_step(h, time, nsteps))
    check(gs_grid1d_status(h, flags))
    if (flags(0) != 0) { gs_grid1d_clear_status(h); raise(flags(0)) }
  }
  /** state in a host array: upload, nsteps, download, pipelined over line chunks */
  def stepHost(time: Double, grids: Array[Double], leftT: Array[Double], rightT: Array[Double], nsteps: Int = 1): Unit =
    check(gs_grid1d_step_host(h, time, nsteps, grids, leftT, rightT, if (leftT.length == 1) 0 else 1))
  override def close(): Unit = { gs_grid1d_destroy(h); gs_ctx_destroy(ctx) }
}
