/*
 * gs_oracle.c -- literal CPU restatement of the gridsplines implicit
 * time-stepper.  TEST INFRASTRUCTURE ONLY; see gs_oracle.h for the pinning
 * statement and the citation conventions (A/ and P/ prefixes).
 *
 * Every arithmetic expression below keeps the reference's operation order.
 * Build with -ffp-contract=off so that gcc never fuses a*b+c; fma() is called
 * exactly where the Scala calls Math.fma (P/PieceFunction.scala:143-231).
 */
#include "gs_oracle.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <pthread.h>

static __thread int g_err = 0;
static int g_ascending_fold = 0;
static int g_libm_pow = 0;

int gso_error(void) { return g_err; }
void gso_clear_error(void) { g_err = 0; }
void gso_set_ascending_fold(int on) { g_ascending_fold = on; }
void gso_set_use_libm_pow(int on) { g_libm_pow = on; }

/* java.lang.Math.min / max (NaN-propagating, -0.0 < +0.0); scala.math.min/max and
 * cats' Order[Double].min/max both delegate to them. */
static double jmin(double a, double b) {
  if (a != a) return a;
  if (a == 0.0 && b == 0.0) return signbit(a) ? a : b;
  return (a <= b) ? a : b;
}
static double jmax(double a, double b) {
  if (a != a) return a;
  if (a == 0.0 && b == 0.0) return signbit(a) ? b : a;
  return (a >= b) ? a : b;
}
/* java.lang.Math.signum */
static double jsignum(double x) { return (x == 0.0 || x != x) ? x : (x > 0.0 ? 1.0 : -1.0); }
/* scala.math.pow(x, 2.0): HotSpot returns x*x for exponent 2; libm pow optional (App. B Q8) */
static double sq(double x) { return g_libm_pow ? pow(x, 2.0) : x * x; }

/* ===================================================================== */
/* piece functions                                                       */
/* ===================================================================== */
typedef struct {
  int kind;
  double lo, hi; /* interval bounds as stored in the tree */
  double x0;
  double c[4]; /* CONST: c0=value; LINE: c0=intercept,c1=slope; CUBIC: c0..c3 */
} piece_t;

/* P/PieceFunction.scala:188-192 */
static double cubic_horner(double x, double a0, double a1, double a2, double a3) {
  return fma(fma(fma(a3, x, a2), x, a1), x, a0);
}
/* P/PieceFunction.scala:173-176 */
static double quad_horner(double x, double a0, double a1, double a2, double a3, double a4) {
  return fma(fma(fma(fma(a4, x, a3), x, a2), x, a1), x, a0);
}
/* P/PieceFunction.scala:213-215 */
static double quadratic_horner(double x, double a0, double a1, double a2) {
  return fma(fma(a2, x, a1), x, a0);
}

static double piece_apply(const piece_t *p, double x) {
  switch (p->kind) {
    case GSO_CONST: return p->c[0];                       /* P/Const.scala:7 */
    case GSO_LINE: return p->c[1] * x + p->c[0];          /* P/Line.scala:24 */
    default: return cubic_horner(x - p->x0, p->c[0], p->c[1], p->c[2], p->c[3]); /* P/M1Hermite3.scala:18 */
  }
}
static double piece_der(const piece_t *p, double x) {
  switch (p->kind) {
    case GSO_CONST: return 0.0;                           /* P/Const.scala:9 */
    case GSO_LINE: return p->c[1];                        /* P/Line.scala:13 */
    default: /* cubicHornerDerivative P/PieceFunction.scala:200-204 */
      return quadratic_horner(x - p->x0, p->c[1], 2.0 * p->c[2], 3.0 * p->c[3]);
  }
}
static double piece_antider(const piece_t *p, double x) {
  switch (p->kind) {
    case GSO_CONST: return x * p->c[0];                   /* P/Const.scala:11 */
    case GSO_LINE: return p->c[1] / 2.0 * x * x + p->c[0] * x; /* P/Line.scala:26 */
    default: /* cubicHornerIntegral P/PieceFunction.scala:194-198 */
      return quad_horner(x - p->x0, 0.0, p->c[0], 0.5 * p->c[1], p->c[2] / 3.0, 0.25 * p->c[3]);
  }
}
static double piece_area(const piece_t *p, double l, double u) {
  switch (p->kind) {
    case GSO_CONST: return (u - l) * p->c[0];             /* P/Const.scala:30 */
    case GSO_LINE: return (piece_apply(p, u) + piece_apply(p, l)) / 2.0 * (u - l); /* P/Line.scala:28-29 */
    default:
      /* Hermite3/M1Hermite3.area is `???` in the snapshot (P/M1Hermite3.scala:72);
       * the base-class rule P/PieceFunction.scala:126-128 reproduces README.md:99-100
       * bit-exactly. */
      return piece_antider(p, u) - piece_antider(p, l);
  }
}

/* ===================================================================== */
/* interval tree (shape only; P/intervaltree/AbsITree.scala)             */
/* ===================================================================== */
enum { N_INNODE = 0, N_LEAF = 1, N_UPLEAF = 2 };
typedef struct {
  int type;
  int piece;
  int left, right; /* -1 = EmptyInt */
} node_t;

struct gso_spline {
  int np;
  piece_t *p;
  node_t *nodes;
  int nnodes;
  int root;
};

typedef struct {
  gso_spline *s;
  int next; /* vals iterator position */
} builder_t;

static int new_node(gso_spline *s, int type, int piece, int l, int r) {
  node_t *n = &s->nodes[s->nnodes];
  n->type = type; n->piece = piece; n->left = l; n->right = r;
  return s->nnodes++;
}
static int build_right(builder_t *b, int size);
/* AbsITree.buildLeft :109-135 */
static int build_left(builder_t *b, int size) {
  gso_spline *s = b->s;
  if (size == 1) return new_node(s, N_LEAF, b->next++, -1, -1);
  if (size == 2) { int leaf = new_node(s, N_LEAF, b->next++, -1, -1); return new_node(s, N_INNODE, b->next++, leaf, -1); }
  if (size == 3) {
    int l = new_node(s, N_LEAF, b->next++, -1, -1);
    int mid = b->next++;
    int r = new_node(s, N_LEAF, b->next++, -1, -1);
    return new_node(s, N_INNODE, mid, l, r);
  }
  int left_index = (size % 2 == 0) ? (size + 1) / 2 : size / 2;
  int l = build_left(b, left_index);
  int mid = b->next++;
  int r = build_right(b, size - left_index - 1);
  return new_node(s, N_INNODE, mid, l, r);
}
/* AbsITree.buildRight :137-166 */
static int build_right(builder_t *b, int size) {
  gso_spline *s = b->s;
  if (size == 0) return -1;
  if (size == 1) {
    int p = b->next++;
    int has_next = b->next < s->np;
    return new_node(s, has_next ? N_LEAF : N_UPLEAF, p, -1, -1);
  }
  if (size == 2) { int mid = b->next++; int r = build_right(b, 1); return new_node(s, N_INNODE, mid, -1, r); }
  if (size == 3) {
    int l = new_node(s, N_LEAF, b->next++, -1, -1);
    int mid = b->next++;
    int r = build_right(b, 1);
    return new_node(s, N_INNODE, mid, l, r);
  }
  int left_index = (size % 2 == 0) ? (size - 1) / 2 : size / 2;
  int l = build_left(b, left_index);
  int mid = b->next++;
  int r = build_right(b, size - left_index - 1);
  return new_node(s, N_INNODE, mid, l, r);
}

/* Interval.relateTo: EQ contained, LT interval below x, GT interval above x
 * (P/intervaltree/AbstractInterval.scala:330-335, Bound.scala:26-98).  CsdOpn for
 * InNode/Leaf, Csd for UpBoundLeaf. */
enum { R_LT = -1, R_EQ = 0, R_GT = 1 };
static int relate(const gso_spline *s, const node_t *n, double x) {
  const piece_t *p = &s->p[n->piece];
  if (p->lo <= x) {
    if (n->type == N_UPLEAF ? (x <= p->hi) : (x < p->hi)) return R_EQ;
    return R_LT;
  }
  return R_GT;
}
/* AbsITree.find :316-333; returns piece index or -1 */
static int tree_find(const gso_spline *s, double x) {
  int cur = s->root;
  while (cur >= 0) {
    const node_t *n = &s->nodes[cur];
    int rel = relate(s, n, x);
    if (rel == R_EQ) return n->piece;
    if (n->type != N_INNODE) return -1;
    /* Interval.isUpperThan(i)(x) <=> x below the interval */
    if (rel == R_GT) cur = n->left; else cur = n->right;
  }
  return -1;
}

/* Interval.foldWithFilter (P/intervaltree/AbstractInterval.scala:368-377) */
static double fold_filter(const gso_spline *s, const node_t *n, double low, double upp) {
  const piece_t *p = &s->p[n->piece];
  double nl = jmax(low, p->lo);
  double nu = jmin(upp, p->hi);
  double diff = nu - nl;
  if (diff <= 0.0) return 0.0;
  return piece_area(p, nl, nu);
}
static double fold_values(const gso_spline *s, const node_t *n) {
  const piece_t *p = &s->p[n->piece];
  return piece_area(p, p->lo, p->hi);
}
/* AbsITree.viewFoldAll :598-611 */
static double fold_all(const gso_spline *s, int cur) {
  if (cur < 0) return 0.0;
  const node_t *n = &s->nodes[cur];
  if (n->type == N_INNODE) {
    double node = fold_values(s, n);
    double leafs = fold_all(s, n->left) + fold_all(s, n->right);
    return node + leafs;
  }
  return fold_values(s, n);
}
/* AbsITree.viewFoldLeft :534-561 */
static double fold_left(const gso_spline *s, int cur, double low) {
  if (cur < 0) return 0.0;
  const node_t *n = &s->nodes[cur];
  const piece_t *p = &s->p[n->piece];
  if (n->type == N_INNODE) {
    int rel = relate(s, n, low);
    if (rel == R_LT) return fold_left(s, n->right, low);
    if (rel == R_EQ) return fold_filter(s, n, low, p->hi) + fold_all(s, n->right);
    double after = fold_left(s, n->left, low);
    double before = fold_values(s, n) + fold_all(s, n->right);
    return before + after;
  }
  return fold_filter(s, n, low, p->hi);
}
/* AbsITree.viewFoldRight :563-595 */
static double fold_right(const gso_spline *s, int cur, double upp) {
  if (cur < 0) return 0.0;
  const node_t *n = &s->nodes[cur];
  const piece_t *p = &s->p[n->piece];
  if (n->type == N_INNODE) {
    int rel = relate(s, n, upp);
    if (rel == R_LT) {
      double before = fold_values(s, n) + fold_all(s, n->left);
      double after = fold_right(s, n->right, upp);
      return before + after;
    }
    if (rel == R_EQ) return fold_filter(s, n, p->lo, upp) + fold_all(s, n->left);
    return fold_right(s, n->left, upp);
  }
  return fold_filter(s, n, p->lo, upp);
}
/* AbsITree.viewFold :485-532 (UpBoundInNode never arises from AbsITree.build) */
static double view_fold(const gso_spline *s, int cur, double low, double upp) {
  if (cur < 0) return 0.0;
  const node_t *n = &s->nodes[cur];
  if (n->type != N_INNODE) return fold_filter(s, n, low, upp);
  int rl = relate(s, n, low);
  if (rl == R_LT) {
    int ru = relate(s, n, upp);
    if (ru == R_LT) return view_fold(s, n->right, low, upp);
    return 0.0;
  }
  if (rl == R_EQ) return fold_filter(s, n, low, upp) + fold_right(s, n->right, upp);
  int ru = relate(s, n, upp);
  if (ru == R_GT) return view_fold(s, n->left, low, upp);
  if (ru == R_EQ) return fold_left(s, n->left, low) + fold_filter(s, n, low, upp);
  return (fold_left(s, n->left, low) + fold_filter(s, n, low, upp)) + fold_right(s, n->right, upp);
}

/* ===================================================================== */
/* Spline                                                                */
/* ===================================================================== */
static gso_spline *spline_alloc(int np) {
  gso_spline *s = (gso_spline *)calloc(1, sizeof(*s));
  s->np = np;
  s->p = (piece_t *)calloc((size_t)np, sizeof(piece_t));
  s->nodes = (node_t *)calloc((size_t)np, sizeof(node_t));
  return s;
}
static void spline_build(gso_spline *s) {
  builder_t b; b.s = s; b.next = 0;
  s->nnodes = 0;
  s->root = build_right(&b, s->np); /* AbsITree.build :82-85 */
}

gso_spline *gso_spline_from_pieces(int np, const int *kind, const double *lo, const double *hi,
                                   const double *x0, const double *coefs) {
  if (np <= 0) { g_err |= 4; return NULL; }
  gso_spline *s = spline_alloc(np);
  for (int i = 0; i < np; ++i) {
    s->p[i].kind = kind[i]; s->p[i].lo = lo[i]; s->p[i].hi = hi[i]; s->p[i].x0 = x0[i];
    memcpy(s->p[i].c, coefs + 4 * i, 4 * sizeof(double));
  }
  spline_build(s);
  return s;
}
void gso_spline_free(gso_spline *s) {
  if (!s) return;
  free(s->p); free(s->nodes); free(s);
}
/* Spline.const(value): singleton(Double.MinValue, Double.MaxValue, Const) P/Spline.scala:362-369;
 * scala Double.MinValue = -Double.MaxValue */
gso_spline *gso_spline_const(double v) { return gso_spline_const_range(-DBL_MAX, DBL_MAX, v); }
gso_spline *gso_spline_const_range(double lo, double hi, double v) {
  gso_spline *s = spline_alloc(1);
  s->p[0].kind = GSO_CONST; s->p[0].lo = lo; s->p[0].hi = hi; s->p[0].x0 = 0.0; s->p[0].c[0] = v;
  spline_build(s);
  return s;
}
/* Line.apply(xLow,xUp,yLow,yUp) P/Line.scala:69-73; interpolate P/PieceFunction.scala:233-236 */
static void make_line(piece_t *p, double xl, double xu, double yl, double yu) {
  double slope = (yu - yl) / (xu - xl);
  double intercept = (xu - xl == 0.0) ? yl : yl + (yu - yl) / (xu - xl) * (0.0 - xl);
  p->kind = GSO_LINE; p->lo = xl; p->hi = xu; p->x0 = 0.0;
  p->c[0] = intercept; p->c[1] = slope; p->c[2] = 0.0; p->c[3] = 0.0;
}
gso_spline *gso_spline_line(double xl, double yl, double xu, double yu) {
  gso_spline *s = spline_alloc(1);
  make_line(&s->p[0], xl, xu, yl, yu);
  spline_build(s);
  return s;
}

typedef struct { double x, y; } pt_t;
static int pt_cmp(const void *a, const void *b) {
  double xa = ((const pt_t *)a)->x, xb = ((const pt_t *)b)->x;
  return (xa > xb) - (xa < xb);
}
/* Spline.apply(list): list.sortBy(_._1) (stable) P/Spline.scala:395-400 */
static pt_t *sorted_points(int n, const double *x, const double *y) {
  pt_t *pts = (pt_t *)malloc((size_t)n * sizeof(pt_t));
  for (int i = 0; i < n; ++i) { pts[i].x = x[i]; pts[i].y = y[i]; }
  /* insertion sort keeps stability like sortBy */
  for (int i = 1; i < n; ++i) {
    pt_t k = pts[i]; int j = i - 1;
    while (j >= 0 && pts[j].x > k.x) { pts[j + 1] = pts[j]; --j; }
    pts[j + 1] = k;
  }
  (void)pt_cmp;
  return pts;
}
gso_spline *gso_spline_lines(int n, const double *x, const double *y) {
  if (n < 2) { g_err |= 4; return NULL; }
  pt_t *pts = sorted_points(n, x, y);
  gso_spline *s = spline_alloc(n - 1);
  for (int i = 0; i + 1 < n; ++i) make_line(&s->p[i], pts[i].x, pts[i + 1].x, pts[i].y, pts[i + 1].y);
  free(pts);
  spline_build(s);
  return s;
}
static gso_spline *spline_const_pieces(int n, const pt_t *pts) {
  /* MakeConstPieceFunctions P/Spline.scala:487-497 */
  gso_spline *s = spline_alloc(n - 1);
  for (int i = 0; i + 1 < n; ++i) {
    s->p[i].kind = GSO_CONST; s->p[i].lo = pts[i].x; s->p[i].hi = pts[i + 1].x; s->p[i].c[0] = pts[i].y;
  }
  spline_build(s);
  return s;
}

/* deriv P/package.scala:93-94 */
static double deriv(pt_t a, pt_t b) { return (b.y - a.y) / (b.x - a.x); }

/* Hermite3.makeSources(list, der1) P/Hermite3.scala:256-272: src[k] = {yLow,yUpp,dLow,dUpp,xLow,xUpp} */
static void make_sources(int n, const pt_t *v, double (*src)[6]) {
  double der1 = deriv(v[0], v[1]);
  for (int k = 0; k + 1 < n; ++k) {
    double der2 = (k + 2 < n) ? deriv(v[k], v[k + 2]) : deriv(v[k], v[k + 1]);
    src[k][0] = v[k].y; src[k][1] = v[k + 1].y; src[k][2] = der1; src[k][3] = der2;
    src[k][4] = v[k].x; src[k][5] = v[k + 1].x;
    der1 = der2;
  }
}
/* Hermite3.monothone(_)(Normal) P/Hermite3.scala:86-91,139-170 */
static void monothone_normal(double *s) {
  double yLow = s[0], yUpp = s[1], dLow = s[2], dUpp = s[3], xLow = s[4], xUpp = s[5];
  double h = xUpp - xLow;
  double delta = (yUpp - yLow) / h;
  double alpha = fabs(dLow / delta);
  double beta = fabs(dUpp / delta);
  double sa, sb;
  if (isnan(alpha) || isnan(beta) || isinf(alpha) || isinf(beta)) { sa = 0.0; sb = 0.0; }
  else {
    double tau = 3.0 / sqrt(sq(alpha) + sq(beta));
    sa = jmin(alpha, tau * alpha);
    sb = jmin(beta, tau * beta);
  }
  s[2] = sa * delta;
  s[3] = sb * delta;
}
/* M1Hermite3.constructSpline P/M1Hermite3.scala:76-86 (also Hermite3 ctor P/Hermite3.scala:20-29) */
static void construct_cubic(piece_t *p, const double *s) {
  double yLow = s[0], yUpp = s[1], sdLow = s[2], sdUpp = s[3], xLow = s[4], xUpp = s[5];
  double delta_ = (yUpp - yLow) / (xUpp - xLow);
  double h_ = xUpp - xLow;
  p->kind = GSO_CUBIC; p->lo = xLow; p->hi = xUpp; p->x0 = xLow;
  p->c[0] = yLow;
  p->c[1] = sdLow;
  p->c[2] = (-2.0 * sdLow - sdUpp + 3.0 * delta_) / h_;
  p->c[3] = (sdLow + sdUpp - 2.0 * delta_) / sq(h_);
}
static gso_spline *m1hermite3_sorted(int n, const pt_t *pts) {
  double (*src)[6] = (double (*)[6])malloc((size_t)(n - 1) * sizeof(*src));
  make_sources(n, pts, src);
  for (int k = 0; k + 1 < n; ++k) monothone_normal(src[k]);
  /* M1Hermite3.smoothness P/M1Hermite3.scala:88-97, applied pairwise in order :109-121 */
  for (int k = 0; k + 2 < n; ++k) {
    double dLeft = src[k][3], dRight = src[k + 1][2];
    double der = jsignum(dLeft) * jmin(fabs(dRight), fabs(dLeft));
    src[k][3] = der; src[k + 1][2] = der;
  }
  gso_spline *s = spline_alloc(n - 1);
  for (int k = 0; k + 1 < n; ++k) construct_cubic(&s->p[k], src[k]);
  free(src);
  spline_build(s);
  return s;
}
gso_spline *gso_spline_m1hermite3(int n, const double *x, const double *y) {
  if (n < 2) { g_err |= 4; return NULL; }
  pt_t *pts = sorted_points(n, x, y);
  gso_spline *s = m1hermite3_sorted(n, pts);
  free(pts);
  return s;
}
/* Hermite3.apply(values: Iterator) P/Hermite3.scala:274-281 (the list-based source builder; the
 * incremental builder behind Spline.fHermite3 drops its last piece in this snapshot) */
static gso_spline *fhermite3_sorted(int n, const pt_t *pts) {
  double (*src)[6] = (double (*)[6])malloc((size_t)(n - 1) * sizeof(*src));
  make_sources(n, pts, src);
  gso_spline *s = spline_alloc(n - 1);
  for (int k = 0; k + 1 < n; ++k) construct_cubic(&s->p[k], src[k]);
  free(src);
  spline_build(s);
  return s;
}
gso_spline *gso_spline_fhermite3(int n, const double *x, const double *y) {
  if (n < 2) { g_err |= 4; return NULL; }
  pt_t *pts = sorted_points(n, x, y);
  gso_spline *s = fhermite3_sorted(n, pts);
  free(pts);
  return s;
}

int gso_spline_size(const gso_spline *s) { return s->np; }
void gso_spline_piece(const gso_spline *s, int i, int *kind, double *lo, double *hi, double *x0,
                      double *c4) {
  *kind = s->p[i].kind; *lo = s->p[i].lo; *hi = s->p[i].hi; *x0 = s->p[i].x0;
  memcpy(c4, s->p[i].c, 4 * sizeof(double));
}
/* AbsITree.unsafeValueAt :473-482 */
static const piece_t *value_at(const gso_spline *s, double x) {
  int i = tree_find(s, x);
  if (i < 0) { g_err |= 1; return NULL; }
  return &s->p[i];
}
double gso_spline_apply(const gso_spline *s, double x) {
  const piece_t *p = value_at(s, x);
  return p ? piece_apply(p, x) : NAN;
}
double gso_spline_der(const gso_spline *s, double x) {
  const piece_t *p = value_at(s, x);
  return p ? piece_der(p, x) : NAN;
}
double gso_spline_antider(const gso_spline *s, double x) {
  const piece_t *p = value_at(s, x);
  return p ? piece_antider(p, x) : NAN;
}
double gso_spline_area_ascending(const gso_spline *s, double lo, double hi) {
  double acc = 0.0;
  for (int i = 0; i < s->np; ++i) {
    double nl = jmax(lo, s->p[i].lo), nu = jmin(hi, s->p[i].hi);
    if (nu - nl <= 0.0) continue;
    acc = acc + piece_area(&s->p[i], nl, nu);
  }
  return acc;
}
double gso_spline_area(const gso_spline *s, double lo, double hi) {
  if (g_ascending_fold) return gso_spline_area_ascending(s, lo, hi);
  return view_fold(s, s->root, lo, hi);
}
double gso_spline_average(const gso_spline *s, double a, double b) {
  double lower = jmin(a, b), upper = jmax(a, b);
  return gso_spline_area(s, lower, upper) / (upper - lower);
}
/* AbsITree.leastK :253-268: lower bound of the leftmost node */
double gso_spline_lower_bound(const gso_spline *s) { return s->p[0].lo; }
/* AbsITree.greatestK :271-280: `case UpBoundLeaf(i, _) => Some(i.lower.bound)` -- the LOWER
 * bound of the last piece.  AbsITree.build always ends in an UpBoundLeaf. */
double gso_spline_upper_bound(const gso_spline *s) { return s->p[s->np - 1].lo; }
double gso_spline_true_upper_bound(const gso_spline *s) { return s->p[s->np - 1].hi; }

/* Spline./ P/Spline.scala:198-210 */
gso_spline *gso_spline_div(const gso_spline *a, const gso_spline *b, int which, int true_upper) {
  int cap = 2 * (a->np + b->np);
  double *xs = (double *)malloc((size_t)cap * sizeof(double));
  int m = 0;
  for (int i = 0; i < b->np; ++i) { xs[m++] = b->p[i].lo; xs[m++] = b->p[i].hi; }
  for (int i = 0; i < a->np; ++i) { xs[m++] = a->p[i].lo; xs[m++] = a->p[i].hi; }
  double ua = true_upper ? gso_spline_true_upper_bound(a) : gso_spline_upper_bound(a);
  double ub = true_upper ? gso_spline_true_upper_bound(b) : gso_spline_upper_bound(b);
  double mn = jmax(gso_spline_lower_bound(a), gso_spline_lower_bound(b));
  double mx = jmin(ua, ub);
  int k = 0;
  for (int i = 0; i < m; ++i) if (xs[i] >= mn && xs[i] <= mx) xs[k++] = xs[i];
  /* to[SortedSet] */
  for (int i = 1; i < k; ++i) { double v = xs[i]; int j = i - 1; while (j >= 0 && xs[j] > v) { xs[j + 1] = xs[j]; --j; } xs[j + 1] = v; }
  int u = 0;
  for (int i = 0; i < k; ++i) if (u == 0 || xs[i] != xs[u - 1]) xs[u++] = xs[i];
  if (u < 2) { free(xs); return NULL; } /* EmptyInt => None */
  pt_t *pts = (pt_t *)malloc((size_t)u * sizeof(pt_t));
  for (int i = 0; i < u; ++i) { pts[i].x = xs[i]; pts[i].y = gso_spline_apply(a, xs[i]) / gso_spline_apply(b, xs[i]); }
  gso_spline *r = NULL;
  switch (which) {
    case 0: r = spline_const_pieces(u, pts); break;
    case 1: { r = spline_alloc(u - 1); for (int i = 0; i + 1 < u; ++i) make_line(&r->p[i], pts[i].x, pts[i + 1].x, pts[i].y, pts[i + 1].y); spline_build(r); break; }
    case 2: r = m1hermite3_sorted(u, pts); break;
    default: r = fhermite3_sorted(u, pts); break;
  }
  free(pts); free(xs);
  return r;
}

/* ===================================================================== */
/* AlwaysDefinedSpline  (A/AlwaysDefinedSpline.scala)                    */
/* ===================================================================== */
struct gso_ads {
  const gso_spline *spl;
  double lowerX, upperX, lowerY, upperY;
};
gso_ads *gso_ads_create(const gso_spline *s, int true_upper) {
  gso_ads *a = (gso_ads *)calloc(1, sizeof(*a));
  a->spl = s;
  a->lowerX = gso_spline_lower_bound(s);                        /* :5 */
  a->upperX = true_upper ? gso_spline_true_upper_bound(s) : gso_spline_upper_bound(s); /* :6 */
  a->lowerY = gso_spline_apply(s, a->lowerX);                   /* :7 */
  a->upperY = gso_spline_apply(s, a->upperX);                   /* :8 */
  return a;
}
gso_ads *gso_ads_create_explicit(const gso_spline *s, double lx, double ux, double ly, double uy) {
  gso_ads *a = (gso_ads *)calloc(1, sizeof(*a));
  a->spl = s; a->lowerX = lx; a->upperX = ux; a->lowerY = ly; a->upperY = uy;
  return a;
}
void gso_ads_free(gso_ads *a) { free(a); }
void gso_ads_clamps(const gso_ads *a, double *o) { o[0] = a->lowerX; o[1] = a->upperX; o[2] = a->lowerY; o[3] = a->upperY; }
double gso_ads_apply(const gso_ads *a, double x) {
  if (x <= a->lowerX) return a->lowerY;
  else if (x >= a->upperX) return a->upperY;
  else return gso_spline_apply(a->spl, x);
}
double gso_ads_area(const gso_ads *a, double xLow, double xUpp) {
  double lowerArea =
      (xLow <= a->lowerX) ? ((xUpp <= a->lowerX) ? (xUpp - xLow) * a->lowerY : (a->lowerX - xLow) * a->lowerY) : 0.0;
  double upperArea =
      (xUpp >= a->upperX) ? ((xLow >= a->upperX) ? (xUpp - xLow) * a->upperY : (xUpp - a->upperX) * a->upperY) : 0.0;
  double low = jmax(xLow, a->lowerX);
  double upp = jmin(xUpp, a->upperX);
  double middleArea = gso_spline_area(a->spl, low, upp);
  return lowerArea + middleArea + upperArea;
}
double gso_ads_average(const gso_ads *a, double xLow, double xUpp) {
  if (xLow == xUpp) return gso_ads_apply(a, xLow);
  else return gso_ads_area(a, xLow, xUpp) / (xUpp - xLow);
}

/* ===================================================================== */
/* arraygrid  (A/arraygrid/package.scala)                                */
/* ===================================================================== */
double gso_ah(double l1, double l2, double l3) { return (l3 + l2) / 2.0 - (l1 + l2) / 2.0; }
double gso_dh(double l1, double l2) { return l2 - l1; }
double gso_r_at_half(double r1, double r2) { return (r1 + r2) / 2.0; }
double gso_vol(double r1, double r2, double r3) {
  return (sq(gso_r_at_half(r2, r3)) - sq(gso_r_at_half(r1, r2))) / 2.0;
}
void gso_make_orthogonal_matrix(double low, const double *range, int n, double upp, double sigma, double *out) {
  for (int i = 0; i < n; ++i) {
    double l1 = (i == 0) ? low : range[i - 1];
    double l2 = range[i];
    double l3 = (i == n - 1) ? upp : range[i + 1];
    double average = gso_ah(l1, l2, l3);
    out[2 * i] = sigma / average / gso_dh(l1, l2);
    out[2 * i + 1] = sigma / average / gso_dh(l2, l3);
  }
}
void gso_make_radial_matrix(double low, const double *range, int n, double upp, double sigma, double *out) {
  for (int i = 0; i < n; ++i) {
    double r1 = (i == 0) ? low : range[i - 1];
    double r2 = range[i];
    double r3 = (i == n - 1) ? upp : range[i + 1];
    double volume = gso_vol(r1, r2, r3);
    out[2 * i] = sigma / volume * gso_r_at_half(r1, r2) / gso_dh(r1, r2);
    out[2 * i + 1] = sigma / volume * gso_r_at_half(r2, r3) / gso_dh(r2, r3);
  }
}
void gso_predef(int dir, double low, const double *range, int n, double upp, double sigma, double *out) {
  if (dir == GSO_RADIAL) gso_make_radial_matrix(low, range, n, upp, sigma, out);
  else gso_make_orthogonal_matrix(low, range, n, upp, sigma, out);
}
void gso_heatflow_geometry(int dir, double low, const double *range, int n, double upp, double *o) {
  double x0 = range[0], x1 = range[1], xm = range[n - 2], xl = range[n - 1];
  if (dir == GSO_RADIAL) {
    o[0] = gso_vol(low, x0, x1);                     /* A/TypeDir.scala:65-71 */
    o[1] = gso_r_at_half(x0, x1) / gso_dh(x0, x1);
    o[2] = gso_vol(xm, xl, upp);                     /* :73-79 */
    o[3] = gso_r_at_half(xm, xl) / gso_dh(xm, xl);
  } else {
    o[0] = gso_ah(low, x0, x1);                      /* A/TypeDir.scala:31-37 */
    o[1] = 1.0 / gso_dh(x0, x1);
    o[2] = gso_ah(xm, xl, upp);                      /* :39-44 */
    o[3] = 1.0 / gso_dh(xm, xl);
  }
}
double gso_aval_1d(double oldv, double newv) { return newv + (newv - oldv) / 3.0; }
double gso_aval_2d(double oldv, double newv) {
  double delta = newv - oldv;
  if (fabs(delta) > 1.5) return newv;
  else return newv + delta / 2.0;
}

/* ===================================================================== */
/* passion  (A/passion/package.scala)                                    */
/* ===================================================================== */
static double cons(double t1, double t2, double c, const gso_ads *z) { /* :21-25 */
  double cond = gso_ads_apply(z, (t1 + t2) / 2.0);
  return cond * c;
}
static double take_average(double t1, double t2, const gso_ads *z) {   /* :33-36 */
  return gso_ads_average(z, jmin(t1, t2), jmax(t1, t2));
}
static double del(double c, double b, double delta) { return c + b * delta; }          /* :230 */
static double lam(double r, double b, double lambda, double d) { return (r - b * lambda) / d; } /* :234 */
static void forward_first(double c, double d, double vect, double *res) {              /* :245-248 */
  res[0] = -d / c;
  res[1] = vect / c;
}
static void forward_unit(double b, double c, double d, double vect, double delta, double lambda, double *res) { /* :257-265 */
  if (!(fabs(c) >= fabs(d) + fabs(b))) g_err |= 2; /* assertion :279-282 */
  double bigDelta = del(c, b, delta);
  res[0] = -d / bigDelta;
  res[1] = lam(vect, b, lambda, bigDelta);
}
static void forward_last(double b, double c, double vect, double delta, double lambda, double *res) { /* :269-276 */
  double bigDelta = del(c, b, delta);
  double nextLambda = lam(vect, b, lambda, bigDelta);
  res[0] = 0.0;
  res[1] = nextLambda;
}
static void backward(const double *tp, double *result, int n, long stride) { /* :303-365 */
  double inter = 0.0;
  for (int i = n - 1; i > -1; --i) {
    inter = tp[2 * i] * inter + tp[2 * i + 1];
    result[(long)i * stride] = inter;
  }
}

void gso_passion_iteration(double time, double firstT, const double *t, int n, double lastT,
                           const double *predict, const double *pd, const gso_ads *const *conds,
                           double *tp, double *result) {
  int z = 0;
  double t1 = firstT, t2 = predict[z], t3 = predict[z + 1];
  double a = cons(t1, t2, pd[2 * z], conds[2 * z]);
  double c = cons(t2, t3, pd[2 * z + 1], conds[2 * z + 1]);
  double vect = -t[z] / time - a * t1;
  forward_first(-(a + c) - 1.0 / time, c, vect, tp + 2 * z);
  z += 1;
  while (z < n - 1) {
    t1 = t2; t2 = predict[z]; t3 = predict[z + 1];
    a = cons(t1, t2, pd[2 * z], conds[2 * z]);
    c = cons(t2, t3, pd[2 * z + 1], conds[2 * z + 1]);
    vect = -t[z] / time;
    forward_unit(a, -(a + c) - 1.0 / time, c, vect, tp[2 * (z - 1)], tp[2 * (z - 1) + 1], tp + 2 * z);
    z += 1;
  }
  t1 = t2; t2 = predict[z]; t3 = lastT;
  a = cons(t1, t2, pd[2 * z], conds[2 * z]);
  c = cons(t2, t3, pd[2 * z + 1], conds[2 * z + 1]);
  vect = -t[z] / time - c * t3;
  forward_last(a, -(a + c) - 1.0 / time, vect, tp[2 * (z - 1)], tp[2 * (z - 1) + 1], tp + 2 * z);
  backward(tp, result, n, 1);
}
void gso_passion_iteration_single(double time, double firstT, const double *t, int n, double lastT,
                                  const double *predict, const double *pd, const gso_ads *zz,
                                  double *tp, double *result) {
  int r = 0;
  double t1 = firstT, t2 = predict[r], t3 = predict[r + 1];
  double a = cons(t1, t2, pd[2 * r], zz);
  double c = cons(t2, t3, pd[2 * r + 1], zz);
  double vect = -t[r] / time - a * t1;
  forward_first(-(a + c) - 1.0 / time, c, vect, tp + 2 * r);
  r += 1;
  while (r < n - 1) {
    t1 = t2; t2 = predict[r]; t3 = predict[r + 1];
    a = cons(t1, t2, pd[2 * r], zz);
    c = cons(t2, t3, pd[2 * r + 1], zz);
    vect = -t[r] / time;
    forward_unit(a, -(a + c) - 1.0 / time, c, vect, tp[2 * (r - 1)], tp[2 * (r - 1) + 1], tp + 2 * r);
    r += 1;
  }
  t1 = t2; t2 = predict[r]; t3 = lastT;
  a = cons(t1, t2, pd[2 * r], zz);
  c = cons(t2, t3, pd[2 * r + 1], zz);
  vect = -t[r] / time - c * t3;
  /* `1/time` here (:135) is Int/Double = the same double division as 1.0/time */
  forward_last(a, -(a + c) - 1 / time, vect, tp[2 * (r - 1)], tp[2 * (r - 1) + 1], tp + 2 * r);
  backward(tp, result, n, 1);
}

/* ===================================================================== */
/* OneDGrid  (A/OneDGrid.scala)                                          */
/* ===================================================================== */
struct gso_grid1d {
  int n;
  double *grid, *predict, *result, *passion, *predef;
  const gso_ads **conds; /* 2n */
  gso_ads *owned;        /* single-spline convenience ctor */
};
gso_grid1d *gso_grid1d_create(int dir, double leftX, const double *rangeX, int n, double rightX,
                              const gso_ads *const *conds, double sigma) {
  gso_grid1d *g = (gso_grid1d *)calloc(1, sizeof(*g));
  g->n = n;
  g->grid = (double *)malloc((size_t)n * sizeof(double));
  g->predict = (double *)malloc((size_t)n * sizeof(double));
  g->result = (double *)malloc((size_t)n * sizeof(double));
  g->passion = (double *)calloc((size_t)2 * n, sizeof(double));
  g->predef = (double *)malloc((size_t)2 * n * sizeof(double));
  g->conds = (const gso_ads **)malloc((size_t)2 * n * sizeof(*g->conds));
  for (int i = 0; i < n; ++i) { g->grid[i] = 4.0; g->predict[i] = 4.0; g->result[i] = 4.0; } /* :20-22 */
  for (int i = 0; i < 2 * n; ++i) g->conds[i] = conds[i];
  gso_predef(dir, leftX, rangeX, n, rightX, sigma, g->predef); /* :24 */
  return g;
}
gso_grid1d *gso_grid1d_create_single(int dir, double leftX, const double *rangeX, int n, double rightX,
                                     const gso_spline *cond, double sigma, int true_upper) {
  gso_ads *a = gso_ads_create(cond, true_upper); /* companion apply :66-79 */
  const gso_ads **cs = (const gso_ads **)malloc((size_t)2 * n * sizeof(*cs));
  for (int i = 0; i < 2 * n; ++i) cs[i] = a;
  gso_grid1d *g = gso_grid1d_create(dir, leftX, rangeX, n, rightX, cs, sigma);
  free(cs);
  g->owned = a;
  return g;
}
void gso_grid1d_free(gso_grid1d *g) {
  if (!g) return;
  free(g->grid); free(g->predict); free(g->result); free(g->passion); free(g->predef);
  free((void *)g->conds); gso_ads_free(g->owned); free(g);
}
double *gso_grid1d_grid(gso_grid1d *g) { return g->grid; }
double *gso_grid1d_predict(gso_grid1d *g) { return g->predict; }
void gso_grid1d_step(gso_grid1d *g, double time, double leftT, double rightT) { /* :26-34 */
  gso_passion_iteration(time, leftT, g->grid, g->n, rightT, g->predict, g->predef, g->conds, g->passion, g->result);
  for (int i = 0; i != g->n; ++i) g->predict[i] = gso_aval_1d(g->grid[i], g->result[i]);
  memcpy(g->grid, g->result, (size_t)g->n * sizeof(double));
}

typedef struct {
  int n, l0, l1, nsteps;
  double time;
  const double *pd, *leftT, *rightT;
  const gso_ads *z;
  double *grid, *predict;
  int err;
} batch_job_t;
static void *batch_worker(void *arg) {
  batch_job_t *j = (batch_job_t *)arg;
  int n = j->n;
  double *tp = (double *)malloc((size_t)2 * n * sizeof(double));
  double *res = (double *)malloc((size_t)n * sizeof(double));
  for (int l = j->l0; l < j->l1; ++l) {
    double *g = j->grid + (size_t)l * n, *p = j->predict + (size_t)l * n;
    for (int s = 0; s < j->nsteps; ++s) {
      gso_passion_iteration_single(j->time, j->leftT[l], g, n, j->rightT[l], p, j->pd, j->z, tp, res);
      for (int i = 0; i != n; ++i) p[i] = gso_aval_1d(g[i], res[i]);
      memcpy(g, res, (size_t)n * sizeof(double));
    }
  }
  free(tp); free(res);
  j->err = g_err;
  return NULL;
}
void gso_batch1d_step(int dir, double leftX, const double *rangeX, int n, double rightX,
                      const gso_spline *cond, double sigma, int nlines, double *grid, double *predict,
                      const double *leftT, const double *rightT, double time, int nsteps, int nthreads) {
  double *pd = (double *)malloc((size_t)2 * n * sizeof(double));
  gso_predef(dir, leftX, rangeX, n, rightX, sigma, pd);
  gso_ads *z = gso_ads_create(cond, 0);
  if (nthreads < 1) nthreads = 1;
  if (nthreads > nlines) nthreads = nlines;
  pthread_t *th = (pthread_t *)malloc((size_t)nthreads * sizeof(pthread_t));
  batch_job_t *jobs = (batch_job_t *)calloc((size_t)nthreads, sizeof(batch_job_t));
  for (int k = 0; k < nthreads; ++k) {
    batch_job_t *j = &jobs[k];
    j->n = n; j->nsteps = nsteps; j->time = time; j->pd = pd; j->leftT = leftT; j->rightT = rightT;
    j->z = z; j->grid = grid; j->predict = predict;
    j->l0 = (int)((long)nlines * k / nthreads);
    j->l1 = (int)((long)nlines * (k + 1) / nthreads);
    if (k + 1 < nthreads) pthread_create(&th[k], NULL, batch_worker, j);
  }
  batch_worker(&jobs[nthreads - 1]);
  for (int k = 0; k + 1 < nthreads; ++k) { pthread_join(th[k], NULL); g_err |= jobs[k].err; }
  free(th); free(jobs);
  gso_ads_free(z);
  free(pd);
}

/* ===================================================================== */
/* TwoDGrid  (A/TwoDGrid.scala, A/Dim.scala)                             */
/* ===================================================================== */
typedef struct {
  int dir, n;
  double low, upp;
  double *range;
  double *coefs;     /* n x 2, Dim.coefs A/Dim.scala:15 (sigma = 1.0) */
  double firstVol, fC, lastVol, lA; /* :16-17 */
  double *to_passion; /* :18-19 */
} dim_t;
typedef struct {
  int kind, rep, n;
  double *values; /* DENSE n, SPARSE 1 */
} bound_t;
struct gso_grid2d {
  int cols, rows;
  dim_t x, y;
  bound_t b[4];
  double *grid, *predict, *result;
  int nspl;
  const gso_ads **pool;
  int map_mode[4];
  int *map[4];
};

static void dim_init(dim_t *d, int dir, int n, const double *values) {
  d->dir = dir; d->n = n; d->low = values[0]; d->upp = values[n + 1];
  d->range = (double *)malloc((size_t)n * sizeof(double));
  memcpy(d->range, values + 1, (size_t)n * sizeof(double));
  d->coefs = (double *)malloc((size_t)2 * n * sizeof(double));
  gso_predef(dir, d->low, d->range, n, d->upp, 1.0, d->coefs);
  double hf[4];
  gso_heatflow_geometry(dir, d->low, d->range, n, d->upp, hf);
  d->firstVol = hf[0]; d->fC = hf[1]; d->lastVol = hf[2]; d->lA = hf[3];
  d->to_passion = (double *)calloc((size_t)2 * n, sizeof(double));
}
static void dim_free(dim_t *d) { free(d->range); free(d->coefs); free(d->to_passion); }

static int side_len(const gso_grid2d *g, int side) { return (side == GSO_UPP || side == GSO_LOW) ? g->cols : g->rows; }

gso_grid2d *gso_grid2d_create(int cols, int rows, int xdir, int ydir, const double *xv, const double *yv) {
  if (cols < 2 || rows < 2) { g_err |= 4; return NULL; } /* Dim needs range(1) A/Dim.scala:16 */
  gso_grid2d *g = (gso_grid2d *)calloc(1, sizeof(*g));
  g->cols = cols; g->rows = rows;
  dim_init(&g->x, xdir, cols, xv);
  dim_init(&g->y, ydir, rows, yv);
  size_t n = (size_t)cols * rows;
  g->grid = (double *)calloc(n, sizeof(double));    /* makeArray :453-459 */
  g->predict = (double *)calloc(n, sizeof(double));
  g->result = (double *)calloc(n, sizeof(double));
  for (int s = 0; s < 4; ++s) { /* default (One, Temp) = Temperature(Sparse(0.0)) */
    g->b[s].kind = GSO_TEMP; g->b[s].rep = GSO_SPARSE; g->b[s].n = 1;
    g->b[s].values = (double *)calloc(1, sizeof(double));
  }
  for (int s = 0; s < 4; ++s) { g->map_mode[s] = GSO_MAP_SINGLE; g->map[s] = (int *)calloc(1, sizeof(int)); }
  return g;
}
void gso_grid2d_free(gso_grid2d *g) {
  if (!g) return;
  dim_free(&g->x); dim_free(&g->y);
  free(g->grid); free(g->predict); free(g->result);
  for (int s = 0; s < 4; ++s) { free(g->b[s].values); free(g->map[s]); }
  free((void *)g->pool); free(g);
}
void gso_grid2d_set_splines(gso_grid2d *g, int nspl, const gso_ads *const *pool) {
  free((void *)g->pool);
  g->nspl = nspl;
  g->pool = (const gso_ads **)malloc((size_t)nspl * sizeof(*g->pool));
  for (int i = 0; i < nspl; ++i) g->pool[i] = pool[i];
}
static size_t map_len(const gso_grid2d *g, int mode) {
  switch (mode) {
    case GSO_MAP_SINGLE: return 1;
    case GSO_MAP_PER_ROW: return (size_t)g->rows + 1;
    case GSO_MAP_PER_COL: return (size_t)g->cols + 1;
    default: return ((size_t)g->rows + 1) * ((size_t)g->cols + 1);
  }
}
void gso_grid2d_set_map(gso_grid2d *g, int sel, int mode, const int *ids) {
  size_t n = map_len(g, mode);
  free(g->map[sel]);
  g->map_mode[sel] = mode;
  g->map[sel] = (int *)malloc(n * sizeof(int));
  memcpy(g->map[sel], ids, n * sizeof(int));
}
/* Coefficients.{apply,thermConductivity,capacity,tempConductivity}(x, y), indices from -1 */
static const gso_ads *coef(const gso_grid2d *g, int sel, int col, int row) {
  const int *m = g->map[sel];
  int id;
  switch (g->map_mode[sel]) {
    case GSO_MAP_SINGLE: id = m[0]; break;
    case GSO_MAP_PER_ROW: id = m[row + 1]; break;
    case GSO_MAP_PER_COL: id = m[col + 1]; break;
    default: id = m[(size_t)(row + 1) * (g->cols + 1) + (col + 1)]; break;
  }
  return g->pool[id];
}
/* Bound.update(value)/update(values)  A/TwoDGrid.scala:593-665 */
void gso_grid2d_set_bound(gso_grid2d *g, int side, int kind, int rep, const double *v, int n) {
  bound_t *b = &g->b[side];
  int len = side_len(g, side);
  free(b->values);
  b->kind = kind; b->rep = rep;
  if (rep == GSO_DENSE) {
    b->n = len;
    b->values = (double *)malloc((size_t)len * sizeof(double));
    if (n == 1) for (int i = 0; i < len; ++i) b->values[i] = v[0]; /* Dense.update(value) :632-638 */
    else memcpy(b->values, v, (size_t)len * sizeof(double));        /* Dense.update(vals) :628-630 */
  } else {
    b->n = 1;
    b->values = (double *)malloc(sizeof(double));
    if (n == 1) b->values[0] = v[0];                                /* Sparse.update(el) :606-608 */
    else { double sum = 0.0; for (int i = 0; i != n; ++i) sum += v[i]; b->values[0] = sum / n; } /* :595-604 */
  }
}
void gso_grid2d_get_bound(const gso_grid2d *g, int side, double *out, int n) {
  const bound_t *b = &g->b[side];
  for (int i = 0; i < n; ++i) out[i] = (b->rep == GSO_DENSE) ? b->values[i] : b->values[0];
}
static double bget(const gso_grid2d *g, int side, int i) {
  const bound_t *b = &g->b[side];
  return (b->rep == GSO_DENSE) ? b->values[i] : b->values[0];
}
double *gso_grid2d_array(gso_grid2d *g, int which) { return which == 0 ? g->grid : which == 1 ? g->predict : g->result; }
void gso_grid2d_fill(gso_grid2d *g, double v) {
  size_t n = (size_t)g->cols * g->rows;
  for (size_t i = 0; i < n; ++i) { g->grid[i] = v; g->predict[i] = v; }
}

/* Dim.first A/Dim.scala:28-43 */
static double dim_first(dim_t *d, double time, double t1, double t2, double t3, double t,
                        const gso_ads *z0, const gso_ads *z) {
  double co0 = take_average(t1, t2, z0);
  double co = take_average(t2, t3, z);
  double a = d->coefs[0] * co0;
  double c = d->coefs[1] * co;
  double vect = -t / time - a * t1;
  forward_first(-(a + c) - 1.0 / time, c, vect, d->to_passion);
  return co;
}
/* Dim.firstHeatFlow :46-64 */
static double dim_first_heatflow(dim_t *d, double time, double heatFlow, double t2, double t3, double t,
                                 const gso_ads *conductivity, const gso_ads *capacity) {
  double cap = gso_ads_apply(capacity, t);
  double cond = gso_ads_average(conductivity, t2, t3);
  double vol = d->firstVol * cap / time;
  double c = -cond * d->fC;
  double b = vol - c;
  double vect = vol + heatFlow;
  double tCond = cond / cap;
  forward_first(b, c, vect, d->to_passion);
  return tCond;
}
/* Dim.general :67-83 */
static double dim_general(dim_t *d, int pos, double time, double t2, double t3, double t, double co0,
                          const gso_ads *z) {
  double co = take_average(t2, t3, z);
  double a = d->coefs[2 * pos] * co0;
  double c = d->coefs[2 * pos + 1] * co;
  double vect = -t / time;
  forward_unit(a, -(a + c) - 1.0 / time, c, vect, d->to_passion[2 * (pos - 1)],
               d->to_passion[2 * (pos - 1) + 1], d->to_passion + 2 * pos);
  return co;
}
/* Dim.last :86-103 */
static void dim_last(dim_t *d, int pos, double time, double t2, double t3, double t, double co0,
                     const gso_ads *z) {
  double co = take_average(t2, t3, z);
  double a = d->coefs[2 * pos] * co0;
  double c = d->coefs[2 * pos + 1] * co;
  double vect = -t / time - c * t3;
  forward_last(a, -(a + c) - 1.0 / time, vect, d->to_passion[2 * (pos - 1)],
               d->to_passion[2 * (pos - 1) + 1], d->to_passion + 2 * pos);
}
/* Dim.lastHeatFlow :106-128 */
static void dim_last_heatflow(dim_t *d, int pos, double time, double heatFlow, double t1, double t2,
                              double t, const gso_ads *conductivity, const gso_ads *capacity) {
  double cap = gso_ads_apply(capacity, t);
  double cond = gso_ads_average(conductivity, t1, t2);
  double vol = cap * d->lastVol / time;
  double a = -cond * d->lA;
  double b = vol - a;
  double vect = vol - heatFlow;
  forward_last(a, b, vect, d->to_passion[2 * (pos - 1)], d->to_passion[2 * (pos - 1) + 1],
               d->to_passion + 2 * pos);
}

/* TwoDGrid.xIter :135-193.  grid(i) reads predict (:470-472), grid.get(i) reads grid (:464). */
void gso_grid2d_xiter(gso_grid2d *g, double time) {
  const int colsNum = g->cols;
  const double *P = g->predict, *G = g->grid;
  for (int row = 0; row != g->rows; ++row) {
    int col = 0;
    int i = row * colsNum;
    double t2 = P[i], t3 = P[i + 1], t = G[i];
    double c0;
    if (g->b[GSO_LEFT].kind == GSO_TEMP) {
      double t1 = bget(g, GSO_LEFT, row);
      c0 = dim_first(&g->x, time, t1, t2, t3, t, coef(g, GSO_SEL_APPLY, -1, row), coef(g, GSO_SEL_APPLY, col, row));
    } else {
      double heat = bget(g, GSO_LEFT, row);
      c0 = dim_first_heatflow(&g->x, time, heat, t2, t3, t, coef(g, GSO_SEL_THERM, col, row), coef(g, GSO_SEL_CAP, col, row));
    }
    col += 1; i += 1;
    while (col != colsNum - 1) {
      t2 = P[i]; t3 = P[i + 1]; t = G[i];
      c0 = dim_general(&g->x, col, time, t2, t3, t, c0, coef(g, GSO_SEL_APPLY, col, row));
      col += 1; i += 1;
    }
    t2 = P[i]; t = G[i];
    if (g->b[GSO_RIGHT].kind == GSO_TEMP) {
      t3 = bget(g, GSO_RIGHT, row);
      dim_last(&g->x, col, time, t2, t3, t, c0, coef(g, GSO_SEL_APPLY, col, row));
    } else {
      /* t3 is stale (= predict of this last node) :183 */
      dim_last_heatflow(&g->x, col, time, bget(g, GSO_RIGHT, row), t2, t3, t,
                        coef(g, GSO_SEL_TEMPCOND, col, row), coef(g, GSO_SEL_CAP, col, row));
    }
    backward(g->x.to_passion, g->result + (size_t)row * colsNum, colsNum, 1); /* backwardRow */
  }
}
/* TwoDGrid.yIter :195-252.  grid.res(i) reads result (:466). */
void gso_grid2d_yiter(gso_grid2d *g, double time) {
  const int colsNum = g->cols;
  const double *P = g->predict;
  double *R = g->result;
  for (int col = 0; col != colsNum; ++col) {
    int i = col;
    int row = 0;
    double t2 = P[i], t3 = P[i + colsNum], t = R[i];
    double c0;
    if (g->b[GSO_UPP].kind == GSO_TEMP) {
      double t1 = bget(g, GSO_UPP, col);
      c0 = dim_first(&g->y, time, t1, t2, t3, t, coef(g, GSO_SEL_APPLY, col, -1), coef(g, GSO_SEL_APPLY, col, row));
    } else {
      c0 = dim_first_heatflow(&g->y, time, bget(g, GSO_UPP, col), t2, t3, t,
                              coef(g, GSO_SEL_TEMPCOND, col, row), coef(g, GSO_SEL_CAP, col, row));
    }
    i += colsNum; row += 1;
    while (row != g->rows - 1) {
      t2 = P[i]; t3 = P[i + colsNum]; t = R[i];
      c0 = dim_general(&g->y, row, time, t2, t3, t, c0, coef(g, GSO_SEL_APPLY, col, row));
      i += colsNum; row += 1;
    }
    t2 = P[i]; t = R[i];
    if (g->b[GSO_LOW].kind == GSO_TEMP) {
      t3 = bget(g, GSO_LOW, col);
      dim_last(&g->y, row, time, t2, t3, t, c0, coef(g, GSO_SEL_APPLY, col, row));
    } else {
      double t1 = P[i - 1]; /* sic: the left neighbour in the same row :241 */
      dim_last_heatflow(&g->y, row, time, bget(g, GSO_LOW, col), t1, t3, t,
                        coef(g, GSO_SEL_TEMPCOND, col, row), coef(g, GSO_SEL_CAP, col, row));
    }
    backward(g->y.to_passion, R + col, g->rows, colsNum); /* backwardColumn */
  }
}
void gso_grid2d_update(gso_grid2d *g) { /* Grid.update :474-481 */
  size_t n = (size_t)g->cols * g->rows;
  for (size_t i = 0; i < n; ++i) g->predict[i] = gso_aval_2d(g->grid[i], g->result[i]);
  memcpy(g->grid, g->result, n * sizeof(double));
}
void gso_grid2d_iteration(gso_grid2d *g, double time) {
  gso_grid2d_xiter(g, time);
  gso_grid2d_yiter(g, time);
  gso_grid2d_update(g);
}
void gso_grid2d_update_predicted(gso_grid2d *g) {
  memcpy(g->predict, g->grid, (size_t)g->cols * g->rows * sizeof(double));
}
double gso_grid2d_edge_average(const gso_grid2d *g, int side) {
  const double *a = g->grid;
  double sum = 0.0;
  int cols = g->cols, rows = g->rows;
  switch (side) {
    case GSO_LEFT: for (int i = 0; i < cols * rows; i += cols) sum += a[i]; return sum / rows;       /* :353-363 */
    case GSO_RIGHT: for (int i = cols - 1; i < cols * rows; i += cols) sum += a[i]; return sum / rows; /* :369-379 */
    case GSO_UPP: for (int i = 0; i != cols; ++i) sum += a[i]; return sum / cols;                    /* :385-395 */
    default: for (int i = cols * (rows - 1); i != cols * rows; ++i) sum += a[i]; return sum / cols;  /* :401-411 */
  }
}
void gso_grid2d_no_heat_flow(gso_grid2d *g, int side) { /* :413-440 */
  bound_t *b = &g->b[side];
  int cols = g->cols, rows = g->rows;
  if (b->rep == GSO_SPARSE) { b->values[0] = gso_grid2d_edge_average(g, side); return; }
  switch (side) {
    case GSO_UPP: memcpy(b->values, g->grid, (size_t)cols * sizeof(double)); break;
    case GSO_LOW: memcpy(b->values, g->grid + (size_t)cols * (rows - 1), (size_t)cols * sizeof(double)); break;
    case GSO_RIGHT: for (int r = 0; r < rows; ++r) b->values[r] = g->grid[(size_t)r * cols + cols - 1]; break;
    default: for (int r = 0; r < rows; ++r) b->values[r] = g->grid[(size_t)r * cols]; break;
  }
}
double gso_grid2d_av_value(const gso_grid2d *g) {
  size_t n = (size_t)g->cols * g->rows;
  double sum = 0.0;
  for (size_t i = 0; i < n; ++i) sum += g->grid[i]; /* grid.sum = foldLeft(0.0)(_ + _) */
  return sum / (double)n;
}
