/*
 * gs_oracle.h -- CPU restatement of the gridsplines implicit heat/diffusion
 * time-stepper (TEST INFRASTRUCTURE ONLY).
 *
 * This is the parity oracle for gridsplines_b200.  It restates, in plain C and
 * in the reference's own operation order, the functions listed in SURVEY.md
 * section 8(a).  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load it.  The product path
 * (libgs_b200.so) never links or calls anything in this directory.
 *
 * Pinning status: the spline construction/evaluation half is pinned bit-exactly
 * by the reference's published known answers (README.md:77-110, six values).
 * The solver half (assembly + Thomas) has NO floating-point golden vector in
 * the reference (its tests assert inequalities and one 7 % physical check,
 * approximation/src/test/scala/approximation/GridTest.scala:99), and the
 * reference cannot be executed in this image (no JVM): for the solver proper
 * the header says it plainly -- PARITY UNPINNED by reference outputs; the pin is
 * this literal restatement plus GridTest's physical assertion.
 *
 * All citations are relative to /root/reference/:
 *   A/ = approximation/src/main/scala/approximation/
 *   P/ = piecewise/src/main/scala/piecewise/
 *
 * Build: gcc -O2 -ffp-contract=off -fno-fast-math (see oracle/Makefile).  JVM
 * doubles are strict IEEE-754 binary64 with no contraction; fma() appears only
 * where the Scala calls Math.fma.
 */
#ifndef GS_ORACLE_H
#define GS_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

/* piece kinds (P/Const.scala, P/Line.scala, P/Hermite3.scala, P/M1Hermite3.scala) */
enum { GSO_CONST = 0, GSO_LINE = 1, GSO_CUBIC = 2 };
/* direction kinds (A/TypeDir.scala:23-98) */
enum { GSO_ORTHO = 0, GSO_RADIAL = 1 };
/* bound kinds (A/TwoDGrid.scala:687-728) and sides (:744-781) */
enum { GSO_TEMP = 0, GSO_FLOW = 1 };
enum { GSO_UPP = 0, GSO_LOW = 1, GSO_RIGHT = 2, GSO_LEFT = 3 };
enum { GSO_SPARSE = 0, GSO_DENSE = 1 };
/* coefficient selector (A/TwoDGrid.scala:782-787) */
enum { GSO_SEL_APPLY = 0, GSO_SEL_THERM = 1, GSO_SEL_CAP = 2, GSO_SEL_TEMPCOND = 3 };
/* coefficient map mode */
enum { GSO_MAP_SINGLE = 0, GSO_MAP_PER_ROW = 1, GSO_MAP_PER_COL = 2, GSO_MAP_DENSE = 3 };

typedef struct gso_spline gso_spline;
typedef struct gso_ads gso_ads;
typedef struct gso_grid1d gso_grid1d;
typedef struct gso_grid2d gso_grid2d;

/* sticky error word of the calling thread: 0 ok, 1 spline undefined at x
 * (RuntimeException, P/intervaltree/AbsITree.scala:475-479), 2 dominance assert
 * (A/passion/package.scala:279-282), 4 bad argument. */
int gso_error(void);
void gso_clear_error(void);

/* ---- piecewise.Spline ------------------------------------------------- */
/* explicit pieces: piece p covers [lo[p], hi[p]) (last: closed), kind[p], x0[p],
 * coefs[4*p..]; tree built like AbsITree.build (P/intervaltree/AbsITree.scala:82-166). */
gso_spline *gso_spline_from_pieces(int npieces, const int *kind, const double *lo,
                                   const double *hi, const double *x0,
                                   const double *coefs);
gso_spline *gso_spline_const(double value);                       /* P/Spline.scala:362-369 */
gso_spline *gso_spline_const_range(double lo, double hi, double v);/* :355-360 */
gso_spline *gso_spline_line(double xl, double yl, double xu, double yu); /* :375-382 */
gso_spline *gso_spline_lines(int n, const double *x, const double *y);   /* :384-385 */
gso_spline *gso_spline_m1hermite3(int n, const double *x, const double *y); /* :387-389 */
gso_spline *gso_spline_fhermite3(int n, const double *x, const double *y);  /* :391-393 */
/* this / that resampled at the common knots (P/Spline.scala:198-210); which:
 * 0 const-pieces 1 lines 2 m1hermite3 3 fhermite3.  true_upper selects the
 * non-literal upper bound. NULL on failure (None). */
gso_spline *gso_spline_div(const gso_spline *a, const gso_spline *b, int which, int true_upper);
void gso_spline_free(gso_spline *s);

int gso_spline_size(const gso_spline *s);
/* copies piece p out (flattening used to feed the C ABI under test) */
void gso_spline_piece(const gso_spline *s, int p, int *kind, double *lo, double *hi,
                      double *x0, double *coefs4);
double gso_spline_apply(const gso_spline *s, double x);      /* P/Spline.scala:26 */
double gso_spline_der(const gso_spline *s, double x);        /* :49 */
double gso_spline_antider(const gso_spline *s, double x);    /* :71-73 */
double gso_spline_area(const gso_spline *s, double lo, double hi); /* :104-108, literal tree fold */
double gso_spline_area_ascending(const gso_spline *s, double lo, double hi); /* same terms, ascending sum */
double gso_spline_average(const gso_spline *s, double a, double b); /* :92-96 */
double gso_spline_lower_bound(const gso_spline *s);          /* :304 */
double gso_spline_upper_bound(const gso_spline *s);          /* :306, literal greatestK quirk */
double gso_spline_true_upper_bound(const gso_spline *s);

/* ---- approximation.AlwaysDefinedSpline (A/AlwaysDefinedSpline.scala) -- */
gso_ads *gso_ads_create(const gso_spline *s, int true_upper);
gso_ads *gso_ads_create_explicit(const gso_spline *s, double lowerX, double upperX,
                                 double lowerY, double upperY);
void gso_ads_free(gso_ads *a);
void gso_ads_clamps(const gso_ads *a, double *four);
double gso_ads_apply(const gso_ads *a, double x);            /* :10-14 */
double gso_ads_area(const gso_ads *a, double lo, double hi); /* :20-41 */
double gso_ads_average(const gso_ads *a, double lo, double hi); /* :43-46 */
/* ascending-order variant of area (the order the CUDA path uses) */
void gso_set_ascending_fold(int on);

/* ---- approximation.arraygrid (A/arraygrid/package.scala) -------------- */
double gso_ah(double l1, double l2, double l3);              /* :17-19 */
double gso_dh(double l1, double l2);                         /* :98-100 */
double gso_r_at_half(double r1, double r2);                  /* :185-187 */
double gso_vol(double r1, double r2, double r3);             /* :196-198 (pow(x,2.0) as x*x) */
void gso_set_use_libm_pow(int on);                           /* vol via pow(x, 2.0) instead */
/* out[2*i], out[2*i+1] = a, c of node i */
void gso_make_orthogonal_matrix(double low, const double *range, int n, double upp,
                                double sigma, double *out);  /* :21-78 */
void gso_make_radial_matrix(double low, const double *range, int n, double upp,
                            double sigma, double *out);      /* :120-176 */
void gso_predef(int dir, double low, const double *range, int n, double upp,
                double sigma, double *out);                  /* A/TypeDir.scala:25-27,55-57 */
/* out = (firstVol, fC, lastVol, lA)  A/Dim.scala:16-17, A/TypeDir.scala:31-44,65-79 */
void gso_heatflow_geometry(int dir, double low, const double *range, int n, double upp,
                           double *out4);
double gso_aval_1d(double oldv, double newv);                /* A/arraygrid/package.scala:85-87 */
double gso_aval_2d(double oldv, double newv);                /* A/TwoDGrid.scala:508-512 */

/* ---- approximation.passion (A/passion/package.scala) ------------------ */
/* overload 1 (:52-95): conds[2*i], conds[2*i+1] per-node spline pair */
void gso_passion_iteration(double time, double firstT, const double *t, int n, double lastT,
                           const double *predict, const double *predef,
                           const gso_ads *const *conds, double *to_passion, double *result);
/* overload 2 (:97-139): one spline */
void gso_passion_iteration_single(double time, double firstT, const double *t, int n,
                                  double lastT, const double *predict, const double *predef,
                                  const gso_ads *z, double *to_passion, double *result);

/* ---- approximation.OneDGrid (A/OneDGrid.scala:14-34) ------------------ */
gso_grid1d *gso_grid1d_create(int dir, double leftX, const double *rangeX, int n,
                              double rightX, const gso_ads *const *conds /* 2n */,
                              double sigma);
gso_grid1d *gso_grid1d_create_single(int dir, double leftX, const double *rangeX, int n,
                                     double rightX, const gso_spline *conductivity,
                                     double sigma, int true_upper);
void gso_grid1d_free(gso_grid1d *g);
double *gso_grid1d_grid(gso_grid1d *g);      /* public val grid */
double *gso_grid1d_predict(gso_grid1d *g);   /* private in the reference; exposed for tests */
void gso_grid1d_step(gso_grid1d *g, double time, double leftT, double rightT);
/* nlines independent OneDGrids sharing geometry/spline, line-major arrays
 * grid[line*n+i]; used as the CPU baseline for the batched config. */
void gso_batch1d_step(int dir, double leftX, const double *rangeX, int n, double rightX,
                      const gso_spline *conductivity, double sigma, int nlines,
                      double *grid, double *predict, const double *leftT,
                      const double *rightT, double time, int nsteps, int nthreads);

/* ---- approximation.TwoDGrid (A/TwoDGrid.scala) ------------------------ */
/* xvalues = low ++ range ++ upp (cols+2), likewise yvalues (A/Dim.scala:13). */
gso_grid2d *gso_grid2d_create(int cols, int rows, int xdir, int ydir,
                              const double *xvalues, const double *yvalues);
void gso_grid2d_free(gso_grid2d *g);
/* spline pool + selector maps: ids index the pool; map layout per mode:
 * SINGLE 1; PER_ROW rows+1 (index row+1); PER_COL cols+1 (index col+1);
 * DENSE (rows+1)*(cols+1) at (row+1)*(cols+1)+(col+1). */
void gso_grid2d_set_splines(gso_grid2d *g, int nspl, const gso_ads *const *pool);
void gso_grid2d_set_map(gso_grid2d *g, int selector, int mode, const int *ids);
void gso_grid2d_set_bound(gso_grid2d *g, int side, int kind, int rep, const double *v, int n);
void gso_grid2d_get_bound(const gso_grid2d *g, int side, double *out, int n);
double *gso_grid2d_array(gso_grid2d *g, int which); /* 0 grid 1 predict 2 result */
void gso_grid2d_fill(gso_grid2d *g, double v);               /* Grid.*= :488-495 */
void gso_grid2d_xiter(gso_grid2d *g, double time);           /* :135-193 */
void gso_grid2d_yiter(gso_grid2d *g, double time);           /* :195-252 */
void gso_grid2d_update(gso_grid2d *g);                       /* :474-481 */
void gso_grid2d_iteration(gso_grid2d *g, double time);       /* :254-258 */
void gso_grid2d_no_heat_flow(gso_grid2d *g, int side);       /* :413-440 */
void gso_grid2d_update_predicted(gso_grid2d *g);             /* :442-444 */
double gso_grid2d_edge_average(const gso_grid2d *g, int side); /* :353-411 */
double gso_grid2d_av_value(const gso_grid2d *g);             /* :483-486 */

#ifdef __cplusplus
}
#endif
#endif
