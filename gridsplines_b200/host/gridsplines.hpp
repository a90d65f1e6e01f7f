// gridsplines.hpp -- header-only C++17 host mirror of the reference's Scala surface for the time-stepper path,
// on top of the C ABI (include/gs_b200.h).  Class and member names follow the reference:
//   piecewise::Spline (P/Spline.scala), approximation::AlwaysDefinedSpline (A/AlwaysDefinedSpline.scala),
//   XDim / YDim (A/XDim.scala, A/YDim.scala), OneDGrid (A/OneDGrid.scala), TwoDGrid + ConstantCoef + bounds
//   (A/TwoDGrid.scala).
// Host code builds tables and marshals arrays; every number the stepper produces comes from libgs_b200.so.
// Compile host translation units with -ffp-contract=off: spline construction below must round like the JVM.
#pragma once
#include <algorithm>
#include <array>
#include <cfloat>
#include <cmath>
#include <functional>
#include <limits>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "../../include/gs_b200.h"

namespace gridsplines {

inline void check(int status) {
  if (status != GS_OK) throw std::runtime_error(std::string("libgs_b200: ") + gs_last_error());
}

class Context {
 public:
  explicit Context(int device = 0, void *stream = nullptr) { check(gs_ctx_create(device, stream, &h_)); }
  ~Context() { gs_ctx_destroy(h_); }
  Context(const Context &) = delete;
  Context &operator=(const Context &) = delete;
  gs_ctx *handle() const { return h_; }
  void synchronize() { check(gs_ctx_synchronize(h_)); }

 private:
  gs_ctx *h_ = nullptr;
};

// ---- piecewise.Spline ------------------------------------------------------------------------------------
namespace detail {
inline double sq(double x) { return x * x; }  // scala.math.pow(x, 2.0)
inline double jmin(double a, double b) {       // java.lang.Math.min
  if (a != a) return a;
  if (a == 0.0 && b == 0.0) return std::signbit(a) ? a : b;
  return a <= b ? a : b;
}
inline double signum(double x) { return (x == 0.0 || x != x) ? x : (x > 0.0 ? 1.0 : -1.0); }
}  // namespace detail

class Spline {
 public:
  int kind = GS_CONST;
  std::vector<double> knots, x0, coefs;  // coefs: 4 per piece

  int size() const { return (int)x0.size(); }
  double lowerBound() const { return knots.front(); }
  // AbsITree.greatestK returns the LOWER bound of the trailing UpBoundLeaf (AbsITree.scala:271-280, App. B Q1)
  double upperBound() const { return knots[knots.size() - 2]; }
  double trueUpperBound() const { return knots.back(); }

  // Spline.const(value): one Const piece over (Double.MinValue, Double.MaxValue)  (Spline.scala:355-369)
  static Spline constant(double value, double lo = -DBL_MAX, double hi = DBL_MAX) {
    Spline s;
    s.kind = GS_CONST;
    s.knots = {lo, hi};
    s.x0 = {0.0};
    s.coefs = {value, 0.0, 0.0, 0.0};
    return s;
  }
  using Points = std::vector<std::pair<double, double>>;
  static Points sorted(Points p) {
    std::stable_sort(p.begin(), p.end(), [](auto &a, auto &b) { return a.first < b.first; });  // sortBy(_._1)
    return p;
  }
  // Spline.lines (Spline.scala:384-385); Line.apply (Line.scala:69-73)
  static Spline lines(const Points &points) {
    Points p = sorted(points);
    Spline s;
    s.kind = GS_LINE;
    s.knots.push_back(p[0].first);
    for (size_t i = 0; i + 1 < p.size(); ++i) {
      double xl = p[i].first, yl = p[i].second, xu = p[i + 1].first, yu = p[i + 1].second;
      double slope = (yu - yl) / (xu - xl);
      double intercept = (xu - xl == 0.0) ? yl : yl + (yu - yl) / (xu - xl) * (0.0 - xl);
      s.coefs.insert(s.coefs.end(), {intercept, slope, 0.0, 0.0});
      s.x0.push_back(0.0);
      s.knots.push_back(xu);
    }
    return s;
  }
  // Spline.m1Hermite3 (Spline.scala:387-389; M1Hermite3.scala:76-126; Hermite3.scala:139-170, 256-272)
  static Spline m1Hermite3(const Points &points) { return hermite(points, true); }
  static Spline fHermite3(const Points &points) { return hermite(points, false); }

 private:
  static Spline hermite(const Points &points, bool monotone) {
    Points v = sorted(points);
    const size_t n = v.size();
    if (n < 2) throw std::invalid_argument("a spline needs at least two points");
    auto deriv = [](const std::pair<double, double> &a, const std::pair<double, double> &b) {
      return (b.second - a.second) / (b.first - a.first);
    };
    std::vector<std::array<double, 6>> src(n - 1);  // yLow, yUpp, dLow, dUpp, xLow, xUpp (makeSources)
    double der1 = deriv(v[0], v[1]);
    for (size_t k = 0; k + 1 < n; ++k) {
      double der2 = (k + 2 < n) ? deriv(v[k], v[k + 2]) : deriv(v[k], v[k + 1]);
      src[k] = {v[k].second, v[k + 1].second, der1, der2, v[k].first, v[k + 1].first};
      der1 = der2;
    }
    if (monotone) {
      for (auto &s : src) {  // monothone(_)(Normal)
        double h = s[5] - s[4], delta = (s[1] - s[0]) / h;
        double alpha = std::fabs(s[2] / delta), beta = std::fabs(s[3] / delta), sa, sb;
        if (std::isnan(alpha) || std::isnan(beta) || std::isinf(alpha) || std::isinf(beta)) {
          sa = sb = 0.0;
        } else {
          double tau = 3.0 / std::sqrt(detail::sq(alpha) + detail::sq(beta));
          sa = detail::jmin(alpha, tau * alpha);
          sb = detail::jmin(beta, tau * beta);
        }
        s[2] = sa * delta;
        s[3] = sb * delta;
        if (std::isnan(s[2]) || std::isnan(s[3]))
          throw std::runtime_error("Function after monotonization must be monothone");  // Hermite3.scala:163-164
      }
      for (size_t k = 0; k + 2 < n; ++k) {  // M1Hermite3.smoothness :88-97
        double dLeft = src[k][3], dRight = src[k + 1][2];
        double der = detail::signum(dLeft) * detail::jmin(std::fabs(dRight), std::fabs(dLeft));
        src[k][3] = der;
        src[k + 1][2] = der;
      }
    }
    Spline out;
    out.kind = GS_CUBIC;
    out.knots.push_back(src[0][4]);
    for (auto &s : src) {  // constructSpline :76-86
      double delta = (s[1] - s[0]) / (s[5] - s[4]), h = s[5] - s[4];
      out.coefs.insert(out.coefs.end(), {s[0], s[2], (-2.0 * s[2] - s[3] + 3.0 * delta) / h,
                                         (s[2] + s[3] - 2.0 * delta) / detail::sq(h)});
      out.x0.push_back(s[4]);
      out.knots.push_back(s[5]);
    }
    return out;
  }
};

// ---- approximation.AlwaysDefinedSpline -------------------------------------------------------------------
class AlwaysDefinedSpline {
 public:
  AlwaysDefinedSpline(Context &ctx, const Spline &spl, bool true_upper = false) : ctx_(ctx) {
    lowerX = spl.lowerBound();
    upperX = true_upper ? spl.trueUpperBound() : spl.upperBound();
    if (spl.kind == GS_CONST && spl.size() == 1) {
      lowerY = upperY = spl.coefs[0];
    } else {  // lowerY = spl(lowerX), upperY = spl(upperX) (:7-8) through the device's own Spline.apply
      int raw = reg(spl, -INFINITY, INFINITY, NAN, NAN);
      double xs[2] = {lowerX, upperX}, ys[2];
      check(gs_spline_eval_apply(ctx.handle(), raw, 2, xs, ys));
      lowerY = ys[0];
      upperY = ys[1];
    }
    id = reg(spl, lowerX, upperX, lowerY, upperY);
  }
  double operator()(double x) const {
    double y;
    check(gs_spline_eval_apply(ctx_.handle(), id, 1, &x, &y));
    return y;
  }
  double average(double lo, double hi) const {
    double y;
    check(gs_spline_eval_average(ctx_.handle(), id, 1, &lo, &hi, &y));
    return y;
  }
  int id = -1;
  double lowerX, upperX, lowerY, upperY;

 private:
  int reg(const Spline &s, double lx, double ux, double ly, double uy) {
    int out = -1;
    check(gs_spline_create(ctx_.handle(), s.kind, s.size(), s.knots.data(), s.x0.data(), s.coefs.data(), lx, ux, ly,
                           uy, &out));
    return out;
  }
  Context &ctx_;
};

// ---- approximation.OneDGrid (nlines independent copies when nlines > 1) ----------------------------------
class OneDGrid {
 public:
  OneDGrid(Context &ctx, int dir, double leftX, const std::vector<double> &rangeX, double rightX,
           const AlwaysDefinedSpline &conductivity, double sigma, long long nlines = 1)
      : n_((int)rangeX.size()), nlines_(nlines) {
    int id = conductivity.id;
    check(gs_grid1d_create(ctx.handle(), nlines, n_, dir, leftX, rangeX.data(), rightX, sigma, GS_SPLINE_SINGLE, &id,
                           nullptr, &h_));
  }
  ~OneDGrid() { gs_grid1d_destroy(h_); }
  OneDGrid(const OneDGrid &) = delete;
  // OneDGrid.step(time, leftT, rightT)  A/OneDGrid.scala:26-34
  void step(double time, double leftT, double rightT, int nsteps = 1) {
    check(gs_grid1d_set_bc(h_, &leftT, &rightT, 0));
    check(gs_grid1d_step(h_, time, nsteps));
  }
  std::vector<double> grid() const { return get(GS_GRID); }  // the public `grid` array (:20)
  void setGrid(const std::vector<double> &v) { check(gs_grid1d_upload(h_, GS_GRID, GS_LINE_MAJOR, v.data())); }
  std::vector<double> get(int which) const {
    std::vector<double> out((size_t)nlines_ * n_);
    check(gs_grid1d_download(h_, which, GS_LINE_MAJOR, out.data()));
    return out;
  }
  unsigned status() const {
    uint32_t f = 0;
    check(gs_grid1d_status(h_, &f));
    return f;
  }

 private:
  gs_grid1d *h_ = nullptr;
  int n_;
  long long nlines_;
};

// ---- XDim / YDim (A/XDim.scala:10-15, A/YDim.scala:7-13) ---------------------------------------------------
struct Dim {
  double low, upp;
  std::vector<double> range;
  int t;  // GS_ORTHO / GS_RADIAL
  Dim(double low_, std::vector<double> range_, double upp_, int t_) : low(low_), upp(upp_), range(std::move(range_)), t(t_) {}
  // generator constructor: Iterator.iterate(inc(low))(inc).takeWhile(_ < upp)
  Dim(double low_, const std::function<double(double)> &inc, double upp_, int t_) : low(low_), upp(upp_), t(t_) {
    for (double x = inc(low_); x < upp_; x = inc(x)) range.push_back(x);
  }
  std::vector<double> values() const {  // Dim.values  A/Dim.scala:13
    std::vector<double> v{low};
    v.insert(v.end(), range.begin(), range.end());
    v.push_back(upp);
    return v;
  }
  double coord(int i) const { return range[i]; }
};
struct XDim : Dim {
  using Dim::Dim;
  int colsNum() const { return (int)range.size(); }
};
struct YDim : Dim {
  using Dim::Dim;
  int rowsNum() const { return (int)range.size(); }
};

// ConstantCoef(thermConductivity, capacity, tempConductivity)  A/TwoDGrid.scala:790-816
struct ConstantCoef {
  const AlwaysDefinedSpline &therm, &cap, &tempcond;
};

// ---- approximation.TwoDGrid --------------------------------------------------------------------------------
class TwoDGrid {
 public:
  struct BoundSpec {
    int rep = GS_SPARSE, kind = GS_TEMP;  // (One | Many, Temp | Flow)  A/TwoDGrid.scala:521-551
  };
  TwoDGrid(Context &ctx, const XDim &x, const YDim &y, BoundSpec upp, BoundSpec low, BoundSpec right, BoundSpec left,
           const ConstantCoef &coefs)
      : cols(x.colsNum()), rows(y.rowsNum()) {
    auto xv = x.values(), yv = y.values();
    check(gs_grid2d_create(ctx.handle(), cols, rows, x.t, y.t, xv.data(), yv.data(), &h_));
    spec_[GS_UPP] = upp; spec_[GS_LOW] = low; spec_[GS_RIGHT] = right; spec_[GS_LEFT] = left;
    for (int s = 0; s < 4; ++s) updateBound(s, 0.0);
    int ids[4] = {coefs.tempcond.id, coefs.therm.id, coefs.cap.id, coefs.tempcond.id};  // apply == tempConductivity
    for (int sel = 0; sel < 4; ++sel) check(gs_grid2d_set_map(h_, sel, GS_MAP_SINGLE, &ids[sel]));
    for (int s = 0; s < 4; ++s) noHeatFlow(s);  // the class body ends with noHeatFlow on all sides (:446-449)
  }
  ~TwoDGrid() { gs_grid2d_destroy(h_); }
  TwoDGrid(const TwoDGrid &) = delete;

  void updateBound(int side, double v) { check(gs_grid2d_set_bound(h_, side, spec_[side].kind, spec_[side].rep, &v, 1)); }
  void updateBound(int side, const std::vector<double> &v) {
    check(gs_grid2d_set_bound(h_, side, spec_[side].kind, spec_[side].rep, v.data(), (int)v.size()));
  }
  void fill(double v) { check(gs_grid2d_fill(h_, v)); }                       // grid *= v
  void xIter(double time) { check(gs_grid2d_xiter(h_, time)); }
  void yIter(double time) { check(gs_grid2d_yiter(h_, time)); }
  void update() { check(gs_grid2d_update(h_)); }
  void iteration(double time, int nsteps = 1) { check(gs_grid2d_iteration(h_, time, nsteps)); }
  void noHeatFlow(int side) { check(gs_grid2d_no_heat_flow(h_, side)); }
  void updatePredicted() { check(gs_grid2d_update_predicted(h_)); }
  void setSolver(int solver) { check(gs_grid2d_set_solver(h_, solver)); }
  double edgeAverage(int side) const {
    double v;
    check(gs_grid2d_edge_average(h_, side, &v));
    return v;
  }
  double leftAverage() const { return edgeAverage(GS_LEFT); }
  double rightAverage() const { return edgeAverage(GS_RIGHT); }
  double upperAverage() const { return edgeAverage(GS_UPP); }
  double lowerAverage() const { return edgeAverage(GS_LOW); }
  double avValue() const {
    double v;
    check(gs_grid2d_av_value(h_, &v));
    return v;
  }
  std::vector<double> array(int which) const {  // Grid.grid / predict / result, row-major
    std::vector<double> out((size_t)cols * rows);
    check(gs_grid2d_download(h_, which, out.data()));
    return out;
  }
  void setArray(int which, const std::vector<double> &v) { check(gs_grid2d_upload(h_, which, v.data())); }
  // slices of `grid` and coordinate-wise fills (A/TwoDGrid.scala:29-58, 260-352)
  std::vector<double> row(int rowNum) const {
    std::vector<double> out((size_t)cols);
    check(gs_grid2d_get_row(h_, GS_GRID, rowNum, out.data()));
    return out;
  }
  std::vector<double> col(int colNum) const {
    std::vector<double> out((size_t)rows);
    check(gs_grid2d_get_col(h_, GS_GRID, colNum, out.data()));
    return out;
  }
  std::vector<double> left(int colNum = 0) const { return col(colNum); }
  std::vector<double> right(int colNumFromRight = 0) const { return col(cols - 1 - colNumFromRight); }
  std::vector<double> upper(int rowNum = 0) const { return row(rowNum); }
  std::vector<double> lower(int rowFromLast = 0) const { return row(rows - 1 - rowFromLast); }
  void updateRow(int rowNum, const std::vector<double> &v) { check(gs_grid2d_set_row(h_, rowNum, v.data())); }
  void updateColumn(int colNum, const std::vector<double> &v) { check(gs_grid2d_set_col(h_, colNum, v.data())); }
  template <class F>
  void updateX(const XDim &x, F f) {   // grid(col, row) = f(x_col)
    std::vector<double> v((size_t)cols);
    for (int c = 0; c < cols; ++c) v[c] = f(x.range[c]);
    check(gs_grid2d_update_x(h_, v.data()));
  }
  template <class F>
  void updateY(const YDim &y, F f) {   // grid(col, row) = f(y_row)
    std::vector<double> v((size_t)rows);
    for (int r = 0; r < rows; ++r) v[r] = f(y.range[r]);
    check(gs_grid2d_update_y(h_, v.data()));
  }
  const int cols, rows;

 private:
  gs_grid2d *h_ = nullptr;
  BoundSpec spec_[4];
};

// ---- heat-flux helper formulas of object TwoDGrid (A/TwoDGrid.scala:1139-1181), operation order as written there
inline double getRadialMidPoint(double r0, double r1, double r2, double t0, double t1) {
  const double h0 = (r0 + r1) / 2.0, h1 = (r1 + r2) / 2.0;
  return (t0 * h0 + t1 * h1) / (h0 + h1);
}
inline double getRadialMidPoint(double r0, double r1, double r2, double k0, double k1, double t0, double t1) {
  const double h0 = (r0 + r1) / 2.0, h1 = (r1 + r2) / 2.0;
  const double a = h0 * k0 * (r1 - r0), b = h1 * k1 * (r2 - r1);
  return (a * t0 + b * t1) / (a + b);
}
inline double discHeatFlow(double t0, double t1, double r0, double r1, double k) {
  const double h = (r0 + r1) / 2.0;
  return k * h * (t0 - t1) / (r1 - r0) * 3.141592653589793 * 2.0;
}
inline double radialHeatFlow(double t0, double t1, double r0, double r1, double k) {
  const double resist = std::log(r1 / r0) / (3.141592653589793 * 2.0 * k);
  return (t0 - t1) / resist;
}
inline double discTWithFlow(double heatFlow, double t0, double r0, double r1, double k) {
  const double h = (r0 + r1) / 2.0;
  const double coef = k * h / (r1 - r0) * 3.141592653589793 * 2.0;
  return t0 - heatFlow / coef;
}

}  // namespace gridsplines
