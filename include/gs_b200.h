/*
 * gs_b200.h -- C ABI of libgs_b200.so: the B200-native (sm_100a, FP64 CUDA)
 * implicit heat/diffusion time-stepper of gridsplines.
 *
 * This is the drop-in boundary for ONE path of the reference: everything that
 * runs under TwoDGrid.iteration / xIter / yIter / Grid.update, OneDGrid.step and
 * passion.iteration.  The reference has no FFI of its own (it is pure Scala);
 * each entry point below names the Scala member it replaces.  Citations are
 * relative to the reference tree:
 *   A/ = approximation/src/main/scala/approximation/
 *   P/ = piecewise/src/main/scala/piecewise/
 *
 * Conventions
 *   - plain C: pointers, sizes, enums; no C++/torch/CUDA types in signatures.
 *     `stream` arguments are a cudaStream_t passed as void* (NULL = the library
 *     creates its own non-blocking stream).
 *   - every function returns GS_OK (0) or a negative gs_status; gs_last_error()
 *     returns a thread-local message for the last failure.  No exception
 *     crosses the ABI.
 *   - host arrays are caller-owned and copied; handles own device memory.
 *   - a context and the handles made from it belong to one host thread at a
 *     time (the reference is single-threaded and non-reentrant per Dim,
 *     A/Dim.scala:18-19).
 *   - there is NO CPU implementation behind these symbols: without a CUDA
 *     device gs_ctx_create fails with GS_ERR_NO_DEVICE.
 *   - numeric faults that the JVM reports by exception are reported through a
 *     sticky device status word (gs_*_status): see GS_FLAG_*.
 */
#ifndef GS_B200_H
#define GS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GS_B200_VERSION 100 /* 0.1.0 */

typedef enum gs_status {
  GS_OK = 0,
  GS_ERR_BAD_ARG = -1,
  GS_ERR_NO_DEVICE = -2,
  GS_ERR_CUDA = -3,
  GS_ERR_OOM = -4,
  GS_ERR_STATE = -5
} gs_status;

/* piece kinds: P/Const.scala, P/Line.scala, P/Hermite3.scala + P/M1Hermite3.scala
 * (shifted cubic, FMA Horner, P/PieceFunction.scala:188-198) */
enum { GS_CONST = 0, GS_LINE = 1, GS_CUBIC = 2 };
/* direction kinds: A/TypeDir.scala:23-51 (Ortho), :54-98 (Radial) */
enum { GS_ORTHO = 0, GS_RADIAL = 1 };
/* bound kinds A/TwoDGrid.scala:687-728, representations :593-665, sides :744-781 */
enum { GS_TEMP = 0, GS_FLOW = 1 };
enum { GS_SPARSE = 0, GS_DENSE = 1 };
enum { GS_UPP = 0, GS_LOW = 1, GS_RIGHT = 2, GS_LEFT = 3 };
/* Coefficients members A/TwoDGrid.scala:782-787 */
enum { GS_SEL_APPLY = 0, GS_SEL_THERM = 1, GS_SEL_CAP = 2, GS_SEL_TEMPCOND = 3 };
/* how a selector map is indexed (flattened Coefficients hierarchy, A/TwoDGrid.scala:790-1137):
 * SINGLE 1 id; PER_ROW rows+1 ids at [row+1]; PER_COL cols+1 ids at [col+1];
 * DENSE (rows+1)*(cols+1) ids at [(row+1)*(cols+1)+(col+1)]  (indices start at -1) */
enum { GS_MAP_SINGLE = 0, GS_MAP_PER_ROW = 1, GS_MAP_PER_COL = 2, GS_MAP_DENSE = 3 };
/* arrays of a grid: Grid.grid / predict / result  A/TwoDGrid.scala:461-463, A/OneDGrid.scala:20-22 */
enum { GS_GRID = 0, GS_PREDICT = 1, GS_RESULT = 2 };
/* host layouts of a 1-D batch: LINE_MAJOR host[line*n + node] (one OneDGrid.grid after
 * another); INTERLEAVED host[node*nlines + line] (node-major) */
enum { GS_LINE_MAJOR = 0, GS_INTERLEAVED = 1 };
/* 1-D spline selection: SINGLE = OneDGrid.apply(leftX, rangeX, rightX, conductivity, sigma)
 * A/OneDGrid.scala:66-79 / passion.iteration overload 2 A/passion/package.scala:97-139;
 * PER_NODE = conductivities(z)(0), conductivities(z)(1), overload 1 :52-95 (2n ids) */
enum { GS_SPLINE_SINGLE = 0, GS_SPLINE_PER_NODE = 1 };
/* line-solver algorithm.  THOMAS = one thread per line in the reference's operation order
 * (bit-exact).  PARTITIONED = a line is cut into segments solved concurrently and stitched
 * through a reduced system (changes the operation order: <= 1e-12 relative, not bit-exact). */
enum { GS_SOLVER_AUTO = 0, GS_SOLVER_THOMAS = 1, GS_SOLVER_PARTITIONED = 2 };

/* sticky status flags */
#define GS_FLAG_SPLINE_UNDEFINED 1u /* RuntimeException, P/intervaltree/AbsITree.scala:475-479 (NaN argument) */
#define GS_FLAG_NOT_DOMINANT 2u     /* AssertionError, A/passion/package.scala:279-282 */
#define GS_FLAG_NONFINITE 4u        /* a solved temperature is NaN/Inf (no JVM analogue: values just propagate) */

typedef struct gs_ctx gs_ctx;
typedef struct gs_grid1d gs_grid1d;
typedef struct gs_grid2d gs_grid2d;
typedef struct gs_ragged gs_ragged;
typedef struct gs_sharded2d gs_sharded2d;

int gs_version(void);
const char *gs_last_error(void);

/* ---- context ---------------------------------------------------------- */
int gs_ctx_create(int device, void *stream, gs_ctx **out);
int gs_ctx_destroy(gs_ctx *ctx);
int gs_ctx_synchronize(gs_ctx *ctx);
/* number of kernels this context has launched so far (bench.py's gpu_launches) */
int gs_ctx_launch_count(gs_ctx *ctx, uint64_t *count);
/* tuning knobs (documented in DESIGN.md); returns GS_ERR_BAD_ARG for unknown names */
int gs_ctx_set_option(gs_ctx *ctx, const char *name, int64_t value);

/* ---- spline tables ---------------------------------------------------- */
/* One AlwaysDefinedSpline (A/AlwaysDefinedSpline.scala:4-46) over a piecewise.Spline
 * (P/Spline.scala:13-133): npieces contiguous pieces, piece p owning [knots[p], knots[p+1])
 * and the last one [knots[np-1], knots[np]] (P/intervaltree/AbsITree.scala:141-146).
 * x0[p], coefs[4p..4p+3]: CONST c0=value; LINE c0=intercept c1=slope (evaluated slope*x+intercept,
 * P/Line.scala:24); CUBIC c0..c3 about x0 (P/M1Hermite3.scala:18).  The four clamp scalars are
 * AlwaysDefinedSpline.lowerX/upperX/lowerY/upperY (:5-8) exactly as the host computed them
 * (including the reference's greatestK quirk).  *id_out indexes the context pool. */
int gs_spline_create(gs_ctx *ctx, int kind, int npieces, const double *knots, const double *x0,
                     const double *coefs, double lowerX, double upperX, double lowerY,
                     double upperY, int *id_out);
/* device evaluation of the pool (used by the parity tests; mirrors
 * AlwaysDefinedSpline.apply :10-14 and .average :43-46); n host values in/out */
int gs_spline_eval_apply(gs_ctx *ctx, int id, int n, const double *x, double *out);
int gs_spline_eval_average(gs_ctx *ctx, int id, int n, const double *lo, const double *hi,
                           double *out);
/* GS_FLAG_* raised by the last gs_spline_eval_* call on this context: an argument where the reference throws
 * (NaN -> RuntimeException, P/intervaltree/AbsITree.scala:475-479) sets GS_FLAG_SPLINE_UNDEFINED; the value returned
 * for it is NaN */
int gs_spline_eval_status(gs_ctx *ctx, uint32_t *flags);

/* ---- metric coefficients (K1) ----------------------------------------- */
/* TypeDir.preDef: arraygrid.makeOrthogonalMatrix A/arraygrid/package.scala:21-78 /
 * makeRadialMatrix :120-176; out[2i], out[2i+1] = (a_i, c_i).  Computed on the device. */
int gs_predef(gs_ctx *ctx, int dir, double low, const double *range, int n, double upp,
              double sigma, double *out_2n);
/* (firstVol, fC, lastVol, lA): TypeDir.lowerHeatFlow/upperHeatFlow A/TypeDir.scala:31-44, 65-79 */
int gs_heatflow_geometry(gs_ctx *ctx, int dir, double low, const double *range, int n, double upp,
                         double *out4);

/* ---- batched 1-D grids: nlines independent OneDGrid (A/OneDGrid.scala:14-34) ---- */
/* All lines share geometry and spline selection.  grid = predict = result = 4.0 at creation
 * (:20-22).  spline_ids: 1 id (SINGLE) or 2n ids (PER_NODE).  host_predef (2n doubles) may be
 * NULL (computed on the device from dir/leftX/rangeX/rightX/sigma, :24). */
int gs_grid1d_create(gs_ctx *ctx, int64_t nlines, int n, int dir, double leftX, const double *rangeX,
                     double rightX, double sigma, int spline_mode, const int *spline_ids,
                     const double *host_predef, gs_grid1d **out);
int gs_grid1d_destroy(gs_grid1d *g);
int gs_grid1d_upload(gs_grid1d *g, int which, int layout, const double *host);
int gs_grid1d_download(gs_grid1d *g, int which, int layout, double *host);
int gs_grid1d_fill(gs_grid1d *g, int which, double value);
/* Dirichlet end temperatures; stride 0 = one value for all lines, 1 = per line */
int gs_grid1d_set_bc(gs_grid1d *g, const double *leftT, const double *rightT, int stride);
/* nsteps x OneDGrid.step(time, leftT, rightT) (A/OneDGrid.scala:26-34) for every line */
int gs_grid1d_step(gs_grid1d *g, double time, int nsteps);
/* The drop-in OneDGrid.step with HOST state (the public `grid` array, A/OneDGrid.scala:20): upload grid, nsteps
 * steps, download grid -- chunked over line blocks and pipelined (H2D / compute / D2H overlap).  host_grid is
 * LINE_MAJOR [nlines][n], updated in place; `predict` (private in the reference) stays on the device.
 * Solver: always the per-thread Thomas kernels (bit-exact), chunk by chunk -- also for long constant-property lines,
 * where gs_grid1d_step under GS_SOLVER_AUTO takes the partitioned solver (tolerance 1e-12): a chunk of lines does not
 * change what a line computes, so results equal gs_grid1d_step with GS_SOLVER_THOMAS. */
int gs_grid1d_step_host(gs_grid1d *g, double time, int nsteps, double *host_grid, const double *leftT,
                        const double *rightT, int bc_stride);
int gs_grid1d_set_solver(gs_grid1d *g, int solver);
int gs_grid1d_status(gs_grid1d *g, uint32_t *flags);
int gs_grid1d_clear_status(gs_grid1d *g);
/* raw device pointer of the blocked array a[line / B][node][line % B]; *pitch receives B (32) */
int gs_grid1d_device_ptr(gs_grid1d *g, int which, void **dptr, int64_t *pitch);
/* algorithmic HBM bytes one step moves (32 B per node) */
int gs_grid1d_bytes_per_step(gs_grid1d *g, double *bytes);

/* ---- 2-D grids: TwoDGrid (A/TwoDGrid.scala) ---------------------------- */
/* xvalues = low ++ range ++ upp (cols+2 doubles), yvalues likewise (rows+2): Dim.values
 * A/Dim.scala:13.  grid/predict/result are zero at creation (makeArray :453-459); all four
 * bounds start as Temperature(Sparse(0.0)); all selector maps start SINGLE -> id 0. */
int gs_grid2d_create(gs_ctx *ctx, int cols, int rows, int xdir, int ydir, const double *xvalues,
                     const double *yvalues, gs_grid2d **out);
int gs_grid2d_destroy(gs_grid2d *g);
/* flattened Coefficients: ids index the context spline pool */
int gs_grid2d_set_map(gs_grid2d *g, int selector, int mode, const int *ids);
/* Bound.update(value) / update(values): n == 1 scalar, else n == side length.  SPARSE with
 * n > 1 stores the mean (Sparse.update(vals) A/TwoDGrid.scala:595-604) */
int gs_grid2d_set_bound(gs_grid2d *g, int side, int kind, int rep, const double *values, int n);
int gs_grid2d_get_bound(gs_grid2d *g, int side, double *out, int n);
int gs_grid2d_fill(gs_grid2d *g, double value);          /* Grid.*=(v): grid and predict :488-495 */
int gs_grid2d_upload(gs_grid2d *g, int which, const double *host);   /* row-major rows x cols */
int gs_grid2d_download(gs_grid2d *g, int which, double *host);
/* one row / column of an array (TwoDGrid.row / col / left / right / upper / lower(n, copyTo) :260-352 read `grid`) */
int gs_grid2d_get_row(gs_grid2d *g, int which, int row, double *host);   /* cols values */
int gs_grid2d_get_col(gs_grid2d *g, int which, int col, double *host);   /* rows values */
int gs_grid2d_set_row(gs_grid2d *g, int row, const double *host);        /* updateRow :318-328: grid only */
int gs_grid2d_set_col(gs_grid2d *g, int col, const double *host);        /* updateColumn :291-300: grid only */
/* updateX(f) / updateY(f) :29-58 with f evaluated by the host: column c <- value_per_col[c] / row r <- value_per_row[r];
 * grid only (`grid *= (i, v)` :497-499) */
int gs_grid2d_update_x(gs_grid2d *g, const double *value_per_col);
int gs_grid2d_update_y(gs_grid2d *g, const double *value_per_row);
int gs_grid2d_xiter(gs_grid2d *g, double time);           /* TwoDGrid.xIter :135-193 */
int gs_grid2d_yiter(gs_grid2d *g, double time);           /* TwoDGrid.yIter :195-252 */
int gs_grid2d_update(gs_grid2d *g);                       /* Grid.update :474-481 */
int gs_grid2d_yiter_update(gs_grid2d *g, double time);    /* yIter + Grid.update fused (second half of iteration) */
/* yIter + Grid.update of a row-sharded grid (this rank holds rows [r0, r1) of every column), split around the ranks'
 * exchange; partitioned solver.  halo_prev / halo_next: device, cols doubles = the predict row above / below the
 * local rows, NULL where this rank holds the physical boundary.  begin writes exported[6][cols] (device); the caller
 * all-gathers it in rank order into gathered[nranks][6][cols] and calls end.  Replaces the column walk of
 * TwoDGrid.yIter :195-252 when one column spans several GPUs. */
int gs_grid2d_yiter_update_begin(gs_grid2d *g, double time, const void *halo_prev, const void *halo_next, void *exported);
int gs_grid2d_yiter_update_end(gs_grid2d *g, double time, const void *gathered, int nranks, int rank);
/* device-to-device copies of a whole row-major array on the context's stream (multi-GPU glue) */
int gs_grid2d_copy_from_device(gs_grid2d *g, int which, const void *dptr);
int gs_grid2d_copy_to_device(gs_grid2d *g, int which, void *dptr);
/* use caller-owned device memory (rows * cols doubles) for an array, e.g. the buffer of a collective */
int gs_grid2d_bind_device(gs_grid2d *g, int which, void *dptr);
int gs_grid2d_iteration(gs_grid2d *g, double time, int nsteps); /* TwoDGrid.iteration :254-258 */
int gs_grid2d_update_predicted(gs_grid2d *g);             /* :442-444 */
int gs_grid2d_no_heat_flow(gs_grid2d *g, int side);       /* :413-440 */
int gs_grid2d_edge_average(gs_grid2d *g, int side, double *out); /* left/right/upper/lowerAverage :353-411 */
int gs_grid2d_av_value(gs_grid2d *g, double *out);        /* Grid.avValue :483-486 */
int gs_grid2d_set_solver(gs_grid2d *g, int solver);
int gs_grid2d_status(gs_grid2d *g, uint32_t *flags);
int gs_grid2d_clear_status(gs_grid2d *g);
int gs_grid2d_device_ptr(gs_grid2d *g, int which, void **dptr);
int gs_grid2d_bytes_per_step(gs_grid2d *g, double *bytes); /* 64 B per node */

/* ---- ragged batch of lines in one flattened array: the fixed version of the experimental
 * passion.iteration(time, indices, lengths, upper, grid, predicted, localResult, result, lower, preDef, conds,
 * toPassion) A/passion/package.scala:161-227.  Line l owns lengths[l] consecutive entries of `indices` (positions in
 * the flattened matrix), of predef[.][2] and -- in GS_SPLINE_PER_NODE mode -- of spline_ids[.][2]; its neighbours are
 * the line's previous / next entries, its ends see upper[l] / lower[l].  Every line is passion.iteration overload 1
 * (:52-95) on its gathered nodes; `result` is written at the same positions, the rest of it is left alone. ---- */
int gs_ragged_create(gs_ctx *ctx, int64_t matrix_size, int nlines, const int *lengths, const int *indices,
                     const double *predef, int spline_mode, const int *spline_ids, gs_ragged **out);
int gs_ragged_destroy(gs_ragged *g);
int gs_ragged_upload(gs_ragged *g, int which, const double *host);      /* GS_GRID / GS_PREDICT / GS_RESULT, matrix_size */
int gs_ragged_download(gs_ragged *g, int which, double *host);
int gs_ragged_iteration(gs_ragged *g, double time, const double *upper, const double *lower);   /* nlines each */
int gs_ragged_status(gs_ragged *g, uint32_t *flags);

/* ---- ONE TwoDGrid on several GPUs of a box, owned by one host process (the multi-device form of the boundary, SURVEY
 * 8(b)/(e)): rows sharded over the devices, x sweep local, y sweep = the partitioned elimination continued across the
 * devices.  The exchange of a step (six doubles per column and device + one halo row of `predict` each way) is done by
 * the kernels themselves over NVLink peer memory; the host only orders the devices' streams with events.  Tolerance
 * solver family (GS_SOLVER_PARTITIONED semantics).  devices == NULL: devices 0 .. ndev-1.  rows % ndev == 0.
 * Replaces TwoDGrid(x, y)(...)(coefs) + iteration (A/TwoDGrid.scala:254-258, 553-574) when one grid spans GPUs. ---- */
#define GS_MAX_SHARDS_ABI 16
int gs_sharded2d_create(int ndev, const int *devices, int cols, int rows, int xdir, int ydir, const double *xvalues,
                        const double *yvalues, gs_sharded2d **out);
int gs_sharded2d_destroy(gs_sharded2d *s);
int gs_sharded2d_device_count(gs_sharded2d *s, int *ndev);
/* registers the table in every device's pool; one id for all */
int gs_sharded2d_spline_create(gs_sharded2d *s, int kind, int npieces, const double *knots, const double *x0,
                               const double *coefs, double lowerX, double upperX, double lowerY, double upperY,
                               int *id_out);
int gs_sharded2d_set_map(gs_sharded2d *s, int selector, int mode, const int *ids);       /* ids of the WHOLE grid */
int gs_sharded2d_set_bound(gs_sharded2d *s, int side, int kind, int rep, const double *values, int n);
int gs_sharded2d_fill(gs_sharded2d *s, double value);
int gs_sharded2d_upload(gs_sharded2d *s, int which, const double *host);                  /* rows x cols, row-major */
int gs_sharded2d_download(gs_sharded2d *s, int which, double *host);
int gs_sharded2d_iteration(gs_sharded2d *s, double time, int nsteps);
/* the same, timed on the devices (events on every device's stream): *ms_max = the longest device interval */
int gs_sharded2d_iteration_timed(gs_sharded2d *s, double time, int nsteps, double *ms_max);
int gs_sharded2d_synchronize(gs_sharded2d *s);
int gs_sharded2d_status(gs_sharded2d *s, uint32_t *flags);
int gs_sharded2d_launch_count(gs_sharded2d *s, uint64_t *count);
int gs_sharded2d_set_option(gs_sharded2d *s, const char *name, int64_t value);

#ifdef __cplusplus
}
#endif
#endif /* GS_B200_H */
