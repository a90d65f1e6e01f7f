// gridsplines.hpp -- header-only C++17 host mirror of the reference's Scala surface for the time-stepper path,
// on top of the C ABI (include/gs_b200.h).  Class and member names follow the reference:
//   piecewise::Spline (P/Spline.scala), approximation::AlwaysDefinedSpline (A/AlwaysDefinedSpline.scala),
//   XDim / YDim (A/XDim.scala, A/YDim.scala), OneDGrid (A/OneDGrid.scala), passion.iteration's ragged overload
//   (RaggedLines), TwoDGrid + the Coefficients hierarchy + bounds + slices + text output (A/TwoDGrid.scala), and
//   ShardedTwoDGrid (one TwoDGrid on several GPUs, gs_sharded2d_*).  It offers what gridsplines_b200/api.py offers.
// Host code builds tables and marshals arrays; every number the stepper produces comes from libgs_b200.so.
// Compile host translation units with -ffp-contract=off: spline construction below must round like the JVM.
#pragma once
#include <algorithm>
#include <array>
#include <cfloat>
#include <charconv>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <functional>
#include <limits>
#include <memory>
#include <optional>
#include <ostream>
#include <set>
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
  void setOption(const char *name, long long value) { check(gs_ctx_set_option(h_, name, value)); }
  unsigned long long launchCount() const {
    uint64_t c = 0;
    check(gs_ctx_launch_count(h_, &c));
    return c;
  }
  // the context spline algebra and stand-alone evaluations use when none is given (device 0)
  static Context &instance() {
    static Context ctx(0);
    return ctx;
  }

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
  static Spline line(double xl, double yl, double xu, double yu) { return lines({{xl, yl}, {xu, yu}}); }
  // Spline.m1Hermite3 (Spline.scala:387-389; M1Hermite3.scala:76-126; Hermite3.scala:139-170, 256-272)
  static Spline m1Hermite3(const Points &points) { return hermite(points, true); }
  static Spline fHermite3(const Points &points) { return hermite(points, false); }
  // Quadratic Lagrange pieces (P/Lagrange2.scala:62-113): every window of three consecutive points gives one parabola
  // a x^2 + b x + c (`polynomials` in its own operation order), evaluated about 0 by the FMA Horner rule (:20-21): a
  // GS_CUBIC piece with x0 = 0, c3 = 0.  Window k owns [x_k, x_{k+1}), the last one runs to the last point.
  static Spline lagrange2(const Points &points) {
    Points p = sorted(points);
    if (p.size() < 3) throw std::invalid_argument("Must be 3 points");
    Spline s;
    s.kind = GS_CUBIC;
    for (size_t k = 0; k + 2 < p.size(); ++k) {
      const double x0 = p[k].first, y0 = p[k].second, x1 = p[k + 1].first, y1 = p[k + 1].second, x2 = p[k + 2].first,
                   y2 = p[k + 2].second;
      const double d[3] = {y0 / ((x0 - x1) * (x0 - x2)), y1 / ((x1 - x0) * (x1 - x2)), y2 / ((x2 - x0) * (x2 - x1))};
      const double up[3][3] = {{1.0, -(x1 + x2), x1 * x2}, {1.0, -(x0 + x2), x0 * x2}, {1.0, -(x0 + x1), x0 * x1}};
      double fin[3] = {0.0, 0.0, 0.0};
      for (int a = 0; a < 3; ++a)
        for (int t = 0; t < 3; ++t) fin[t] = fin[t] + up[a][t] * d[a];
      s.knots.push_back(x0);
      s.x0.push_back(0.0);
      s.coefs.insert(s.coefs.end(), {fin[2], fin[1], fin[0], 0.0});   // (a, b, c) -> new Lagrange2(Array(c, b, a))
    }
    s.knots.push_back(p.back().first);
    return s;
  }
  // Natural cubic spline pieces (P/Lagrange3.scala:54-200, ListPassion.solve for the second-derivative terms): piece k
  // over [x_k, x_{k+1}] is cubicRuleOfHorner(x - low, d, c, b, a) (:33-34).  The reference's list recursion, literally.
  static Spline lagrange3(const Points &points) {
    Points v = sorted(points);
    const int n = (int)v.size();
    if (n < 3) throw std::invalid_argument("at least 3 points");
    auto f = [](const std::pair<double, double> &a, const std::pair<double, double> &b) {
      return (b.first - a.first) != 0.0 ? (b.second - a.second) / std::fabs(b.first - a.first) : 0.0;
    };
    auto vect = [&](int a, int b, int c) { return 3 * f(v[b], v[c]) - 3 * f(v[a], v[b]); };
    std::vector<std::array<double, 3>> rows;
    std::vector<double> vec;
    for (int i = 0; i + 2 < n; ++i) {   // cCoefs :76-100
      const double h1 = v[i + 1].first - v[i].first, h2 = v[i + 2].first - v[i + 1].first;
      rows.push_back({h1, 2 * (h2 + h1), h2});
      vec.push_back(vect(i, i + 1, i + 2));
    }
    {
      const double h1 = v[n - 1].first - v[n - 2].first;
      rows.push_back({h1, 2 * (0.0 + h1), 0.0});
      vec.push_back(vect(n - 2, n - 1, n - 1));
    }
    // ListPassion.solve: forwardFirst / forwardUnit / forwardLast, backwardPassion; .head = all but the last value
    std::vector<std::pair<double, double>> dl;
    dl.push_back({-rows[0][2] / rows[0][1], vec[0] / rows[0][1]});
    for (size_t r = 1; r + 1 < rows.size(); ++r) {
      const double b = rows[r][0], c = rows[r][1], d = rows[r][2];
      if (!(std::fabs(c) >= std::fabs(d) + std::fabs(b))) throw std::runtime_error("Passion is not correct or unstable");
      const double big = c + b * dl.back().first;
      dl.push_back({-d / big, (vec[r] - b * dl.back().second) / big});
    }
    {
      const double b = rows.back()[0], c = rows.back()[1];
      const double big = c + b * dl.back().first;
      dl.push_back({0.0, (vec.back() - b * dl.back().second) / big});
    }
    std::vector<double> sol(dl.size());
    double prev = 0.0;
    for (int i = (int)dl.size() - 1; i >= 0; --i) {
      prev = dl[i].first * prev + dl[i].second;
      sol[i] = prev;
    }
    std::vector<double> cc(sol.begin(), sol.end() - 1);   // solve(...).head: the last unknown is split off
    if ((int)cc.size() != n - 2) throw std::invalid_argument("list of values size minus 2 must be more than c list size");
    std::vector<double> dd, bb;
    for (int i = 0; i + 1 < n; ++i) {   // dCoefs :103-125
      const double h = v[i + 1].first - v[i].first;
      if (i == 0) dd.push_back((cc[0] - 0.0) / 3 * h);
      else if (i == n - 2) dd.push_back((0.0 - cc[i - 1]) / 3 * h);
      else dd.push_back((cc[i - 1] - cc[i]) / 3 * h);
    }
    const int nb = std::min({n - 1, (int)cc.size(), (int)dd.size()});
    for (int i = 0; i < nb; ++i) {      // bCoefs :127-139 (stops when c or d run out)
      const double h = v[i + 1].first - v[i].first;
      bb.push_back(f(v[i], v[i + 1]) + cc[i] * detail::sq(h) - dd[i] * std::pow(h, 3.0));
    }
    Spline out;
    out.kind = GS_CUBIC;
    out.knots.push_back(v[0].first);
    for (size_t i = 0; i < cc.size(); ++i) {   // while (c.nonEmpty) :61-69; aCoefs = the upper points' values :141-143
      const double lo = std::min(v[i].first, v[i + 1].first), hi = std::max(v[i].first, v[i + 1].first);
      out.coefs.insert(out.coefs.end(), {dd[i], cc[i], bb[i], v[i + 1].second});
      out.x0.push_back(lo);
      out.knots.push_back(hi);
    }
    return out;
  }

  // Spline.apply: the table registered unclamped in the context's pool and evaluated there (one arithmetic for set-up
  // and stepping)
  std::vector<double> apply(const std::vector<double> &xs, Context &ctx = Context::instance()) const {
    int id = -1;
    check(gs_spline_create(ctx.handle(), kind, size(), knots.data(), x0.data(), coefs.data(), -INFINITY, INFINITY, NAN, NAN, &id));
    std::vector<double> out(xs.size());
    check(gs_spline_eval_apply(ctx.handle(), id, (int)xs.size(), xs.data(), out.data()));
    return out;
  }
  double apply(double x, Context &ctx = Context::instance()) const { return apply(std::vector<double>{x}, ctx)[0]; }
  // Spline.sumArguments (Spline.scala:198-205): both splines' piece bounds inside the common range, sorted, distinct
  std::vector<double> sumArguments(const Spline &other, bool true_upper = false) const {
    const double ua = true_upper ? trueUpperBound() : upperBound(), ub = true_upper ? other.trueUpperBound() : other.upperBound();
    const double mn = std::max(lowerBound(), other.lowerBound()), mx = std::min(ua, ub);
    std::set<double> xs;
    for (const Spline *s : {&other, this})
      for (double k : s->knots)
        if (mn <= k && k <= mx) xs.insert(k);
    return std::vector<double>(xs.begin(), xs.end());
  }
  // Spline./ + - * (Spline.scala:207-225): resample op(this(x), other(x)) at the merged piece bounds and rebuild with
  // `factory` (the implicit PieceFunFactory; default M1Hermite3); nullopt where the reference returns None
  using Factory = std::function<Spline(const Points &)>;
  std::optional<Spline> combine(const Spline &other, const std::function<double(double, double)> &op,
                                const Factory &factory = nullptr, Context &ctx = Context::instance(),
                                bool true_upper = false) const {
    const std::vector<double> xs = sumArguments(other, true_upper);
    if (xs.size() < 2) return std::nullopt;
    const std::vector<double> ya = apply(xs, ctx), yb = other.apply(xs, ctx);
    Points pts;
    for (size_t i = 0; i < xs.size(); ++i) pts.push_back({xs[i], op(ya[i], yb[i])});
    return factory ? factory(pts) : m1Hermite3(pts);
  }
  std::optional<Spline> operator/(const Spline &o) const { return combine(o, [](double a, double b) { return a / b; }); }
  std::optional<Spline> operator+(const Spline &o) const { return combine(o, [](double a, double b) { return a + b; }); }
  std::optional<Spline> operator-(const Spline &o) const { return combine(o, [](double a, double b) { return a - b; }); }
  std::optional<Spline> operator*(const Spline &o) const { return combine(o, [](double a, double b) { return a * b; }); }

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
// (AlwaysDefinedSpline.scala:4-46) registered in a context's device pool
class AlwaysDefinedSpline {
 public:
  AlwaysDefinedSpline(Context &ctx, const Spline &spl, bool true_upper = false) : ctx_(&ctx) {
    lowerX = spl.lowerBound();
    upperX = true_upper ? spl.trueUpperBound() : spl.upperBound();
    if (spl.kind == GS_CONST && spl.size() == 1) {
      lowerY = upperY = spl.coefs[0];
    } else {  // lowerY = spl(lowerX), upperY = spl(upperX) (:7-8) through the device's own Spline.apply
      int raw = reg(spl, -INFINITY, INFINITY, NAN, NAN);
      double xs[2] = {lowerX, upperX}, ys[2];
      check(gs_spline_eval_apply(ctx.handle(), raw, 2, xs, ys));
      throwIfUndefined();
      lowerY = ys[0];
      upperY = ys[1];
    }
    id = reg(spl, lowerX, upperX, lowerY, upperY);
  }
  // explicit clamps (lowerX, upperX, lowerY, upperY) instead of the literal rule
  AlwaysDefinedSpline(Context &ctx, const Spline &spl, const std::array<double, 4> &clamps)
      : lowerX(clamps[0]), upperX(clamps[1]), lowerY(clamps[2]), upperY(clamps[3]), ctx_(&ctx) {
    id = reg(spl, lowerX, upperX, lowerY, upperY);
  }
  double operator()(double x) const { return (*this)(std::vector<double>{x})[0]; }
  std::vector<double> operator()(const std::vector<double> &xs) const {
    std::vector<double> out(xs.size());
    check(gs_spline_eval_apply(ctx_->handle(), id, (int)xs.size(), xs.data(), out.data()));
    throwIfUndefined();
    return out;
  }
  double average(double lo, double hi) const { return average(std::vector<double>{lo}, std::vector<double>{hi})[0]; }
  std::vector<double> average(const std::vector<double> &lo, const std::vector<double> &hi) const {
    std::vector<double> out(lo.size());
    check(gs_spline_eval_average(ctx_->handle(), id, (int)lo.size(), lo.data(), hi.data(), out.data()));
    throwIfUndefined();
    return out;
  }
  Context &context() const { return *ctx_; }
  int id = -1;
  double lowerX, upperX, lowerY, upperY;

 private:
  int reg(const Spline &s, double lx, double ux, double ly, double uy) {
    int out = -1;
    check(gs_spline_create(ctx_->handle(), s.kind, s.size(), s.knots.data(), s.x0.data(), s.coefs.data(), lx, ux, ly,
                           uy, &out));
    return out;
  }
  // "Spline is not defined at point ..." RuntimeException (P/intervaltree/AbsITree.scala:475-479)
  void throwIfUndefined() const {
    uint32_t f = 0;
    check(gs_spline_eval_status(ctx_->handle(), &f));
    if (f & GS_FLAG_SPLINE_UNDEFINED)
      throw std::runtime_error("spline is not defined at one of the arguments (NaN or outside the knots)");
  }
  Context *ctx_;
};

// ---- approximation.OneDGrid: nlines independent copies sharing geometry and conductivity --------------------
class BatchedOneDGrid {
 public:
  using Pair = std::pair<const AlwaysDefinedSpline *, const AlwaysDefinedSpline *>;
  // one conductivity for every node (companion apply, A/OneDGrid.scala:66-79)
  BatchedOneDGrid(Context &ctx, int dir, double leftX, const std::vector<double> &rangeX, double rightX,
                  const AlwaysDefinedSpline &conductivity, double sigma, long long nlines = 1,
                  const std::vector<double> *predef = nullptr)
      : rangeX_(rangeX), n_((int)rangeX.size()), nlines_(nlines) {
    int id = conductivity.id;
    check(gs_grid1d_create(ctx.handle(), nlines, n_, dir, leftX, rangeX.data(), rightX, sigma, GS_SPLINE_SINGLE, &id,
                           predef ? predef->data() : nullptr, &h_));
  }
  // `conductivities(z)(0..1)`: a (left, right) pair per node
  BatchedOneDGrid(Context &ctx, int dir, double leftX, const std::vector<double> &rangeX, double rightX,
                  const std::vector<Pair> &conductivities, double sigma, long long nlines = 1,
                  const std::vector<double> *predef = nullptr)
      : rangeX_(rangeX), n_((int)rangeX.size()), nlines_(nlines) {
    if ((int)conductivities.size() != n_) throw std::invalid_argument("one (left, right) spline pair per node");
    std::vector<int> ids;
    for (auto &pr : conductivities) {
      ids.push_back(pr.first->id);
      ids.push_back(pr.second->id);
    }
    check(gs_grid1d_create(ctx.handle(), nlines, n_, dir, leftX, rangeX.data(), rightX, sigma, GS_SPLINE_PER_NODE,
                           ids.data(), predef ? predef->data() : nullptr, &h_));
  }
  ~BatchedOneDGrid() { gs_grid1d_destroy(h_); }
  BatchedOneDGrid(const BatchedOneDGrid &) = delete;
  BatchedOneDGrid &operator=(const BatchedOneDGrid &) = delete;

  void upload(int which, const std::vector<double> &host, int layout = GS_LINE_MAJOR) {
    if (host.size() != (size_t)nlines_ * n_) throw std::invalid_argument("nlines * n values");
    check(gs_grid1d_upload(h_, which, layout, host.data()));
  }
  std::vector<double> download(int which = GS_GRID, int layout = GS_LINE_MAJOR) const {
    std::vector<double> out((size_t)nlines_ * n_);
    check(gs_grid1d_download(h_, which, layout, out.data()));
    return out;
  }
  std::vector<double> get(int which) const { return download(which); }
  void fill(int which, double v) { check(gs_grid1d_fill(h_, which, v)); }
  void setBc(double leftT, double rightT) { check(gs_grid1d_set_bc(h_, &leftT, &rightT, 0)); }
  void setBc(const std::vector<double> &leftT, const std::vector<double> &rightT) {   // one pair per line
    if ((long long)leftT.size() != nlines_ || (long long)rightT.size() != nlines_)
      throw std::invalid_argument("one end temperature per line");
    check(gs_grid1d_set_bc(h_, leftT.data(), rightT.data(), 1));
  }
  // OneDGrid.step(time, leftT, rightT)  A/OneDGrid.scala:26-34
  void step(double time, double leftT, double rightT, int nsteps = 1) {
    setBc(leftT, rightT);
    check(gs_grid1d_step(h_, time, nsteps));
  }
  void step(double time, const std::vector<double> &leftT, const std::vector<double> &rightT, int nsteps = 1) {
    setBc(leftT, rightT);
    check(gs_grid1d_step(h_, time, nsteps));
  }
  void step(double time, int nsteps = 1) { check(gs_grid1d_step(h_, time, nsteps)); }
  // the same with the state in a HOST array [nlines][n], updated in place (upload, step, download, pipelined)
  void stepHost(double time, double *host_grid, const std::vector<double> &leftT, const std::vector<double> &rightT,
                int nsteps = 1) {
    const int stride = (leftT.size() == 1 && rightT.size() == 1) ? 0 : 1;
    check(gs_grid1d_step_host(h_, time, nsteps, host_grid, leftT.data(), rightT.data(), stride));
  }
  void setSolver(int solver) { check(gs_grid1d_set_solver(h_, solver)); }
  unsigned status() const {
    uint32_t f = 0;
    check(gs_grid1d_status(h_, &f));
    return f;
  }
  void clearStatus() { check(gs_grid1d_clear_status(h_)); }
  double bytesPerStep() const {
    double b = 0.0;
    check(gs_grid1d_bytes_per_step(h_, &b));
    return b;
  }
  int nodes() const { return n_; }
  long long lines() const { return nlines_; }
  const std::vector<double> &rangeX() const { return rangeX_; }
  gs_grid1d *handle() const { return h_; }

 protected:
  gs_grid1d *h_ = nullptr;
  std::vector<double> rangeX_;
  int n_;
  long long nlines_;
};

class OneDGrid : public BatchedOneDGrid {
 public:
  OneDGrid(Context &ctx, int dir, double leftX, const std::vector<double> &rangeX, double rightX,
           const AlwaysDefinedSpline &conductivity, double sigma, long long nlines = 1)
      : BatchedOneDGrid(ctx, dir, leftX, rangeX, rightX, conductivity, sigma, nlines) {}
  OneDGrid(Context &ctx, int dir, double leftX, const std::vector<double> &rangeX, double rightX,
           const std::vector<Pair> &conductivities, double sigma)
      : BatchedOneDGrid(ctx, dir, leftX, rangeX, rightX, conductivities, sigma, 1) {}
  std::vector<double> grid() const { return download(GS_GRID); }  // the public `grid` array (:20)
  std::vector<double> predict() const { return download(GS_PREDICT); }
  void setGrid(const std::vector<double> &v) { upload(GS_GRID, v); }
};

// ---- a ragged batch of lines inside one flattened array: the fixed version of passion.iteration overload 3
// (A/passion/package.scala:161-227).  lines[l] = positions of line l's nodes; preDef[l] = its (a, c) pairs, flattened.
class RaggedLines {
 public:
  RaggedLines(Context &ctx, long long matrix_size, const std::vector<std::vector<int>> &lines,
              const std::vector<std::vector<double>> &preDef, const AlwaysDefinedSpline &conductivity)
      : size_(matrix_size), nlines_((int)lines.size()) {
    std::vector<int> lengths, indices;
    std::vector<double> pd;
    flatten(lines, preDef, lengths, indices, pd);
    int id = conductivity.id;
    check(gs_ragged_create(ctx.handle(), matrix_size, nlines_, lengths.data(), indices.data(), pd.data(), GS_SPLINE_SINGLE,
                           &id, &h_));
  }
  // per line a (left, right) pair for every node
  RaggedLines(Context &ctx, long long matrix_size, const std::vector<std::vector<int>> &lines,
              const std::vector<std::vector<double>> &preDef,
              const std::vector<std::vector<BatchedOneDGrid::Pair>> &conductivities)
      : size_(matrix_size), nlines_((int)lines.size()) {
    std::vector<int> lengths, indices, ids;
    std::vector<double> pd;
    flatten(lines, preDef, lengths, indices, pd);
    for (auto &line : conductivities)
      for (auto &pr : line) {
        ids.push_back(pr.first->id);
        ids.push_back(pr.second->id);
      }
    if (ids.size() != 2 * indices.size()) throw std::invalid_argument("a spline pair per node of every line");
    check(gs_ragged_create(ctx.handle(), matrix_size, nlines_, lengths.data(), indices.data(), pd.data(),
                           GS_SPLINE_PER_NODE, ids.data(), &h_));
  }
  ~RaggedLines() { gs_ragged_destroy(h_); }
  RaggedLines(const RaggedLines &) = delete;
  std::vector<double> get(int which) const {
    std::vector<double> out((size_t)size_);
    check(gs_ragged_download(h_, which, out.data()));
    return out;
  }
  void set(int which, const std::vector<double> &v) {
    if ((long long)v.size() != size_) throw std::invalid_argument("matrix_size values");
    check(gs_ragged_upload(h_, which, v.data()));
  }
  std::vector<double> grid() const { return get(GS_GRID); }
  std::vector<double> predicted() const { return get(GS_PREDICT); }
  std::vector<double> result() const { return get(GS_RESULT); }
  void iteration(double time, const std::vector<double> &upper, const std::vector<double> &lower) {
    if ((int)upper.size() != nlines_ || (int)lower.size() != nlines_) throw std::invalid_argument("one end value per line");
    check(gs_ragged_iteration(h_, time, upper.data(), lower.data()));
  }
  unsigned status() const {
    uint32_t f = 0;
    check(gs_ragged_status(h_, &f));
    return f;
  }

 private:
  static void flatten(const std::vector<std::vector<int>> &lines, const std::vector<std::vector<double>> &preDef,
                      std::vector<int> &lengths, std::vector<int> &indices, std::vector<double> &pd) {
    for (size_t l = 0; l < lines.size(); ++l) {
      lengths.push_back((int)lines[l].size());
      indices.insert(indices.end(), lines[l].begin(), lines[l].end());
      if (preDef[l].size() != 2 * lines[l].size()) throw std::invalid_argument("two metric coefficients per node");
      pd.insert(pd.end(), preDef[l].begin(), preDef[l].end());
    }
  }
  gs_ragged *h_ = nullptr;
  long long size_;
  int nlines_;
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

// ---- the Coefficients hierarchy (A/TwoDGrid.scala:782-1137), flattened: per member a (mode, ids) pair -----------
struct CoefMap {
  int mode = GS_MAP_SINGLE;
  std::vector<int> ids;
};
using CoefMaps = std::array<CoefMap, 4>;   // indexed by GS_SEL_APPLY / THERM / CAP / TEMPCOND
using Splines = std::vector<const AlwaysDefinedSpline *>;

struct Coefficients {
  virtual ~Coefficients() = default;
  virtual CoefMaps maps(int cols, int rows) const = 0;
};

// ConstantCoef(thermConductivity, capacity, tempConductivity) :790-816; apply == tempConductivity
struct ConstantCoef : Coefficients {
  const AlwaysDefinedSpline &therm, &cap, &tempcond;
  ConstantCoef(const AlwaysDefinedSpline &k, const AlwaysDefinedSpline &c, const AlwaysDefinedSpline &a)
      : therm(k), cap(c), tempcond(a) {}
  CoefMaps maps(int, int) const override {
    CoefMaps m;
    m[GS_SEL_APPLY] = {GS_MAP_SINGLE, {tempcond.id}};
    m[GS_SEL_THERM] = {GS_MAP_SINGLE, {therm.id}};
    m[GS_SEL_CAP] = {GS_MAP_SINGLE, {cap.id}};
    m[GS_SEL_TEMPCOND] = {GS_MAP_SINGLE, {tempcond.id}};
    return m;
  }
};

namespace detail {
inline std::vector<int> ids_of(const Splines &v, size_t n) {
  std::vector<int> out;
  for (size_t i = 0; i < n; ++i) out.push_back(v[i]->id);
  return out;
}
}  // namespace detail

// one spline per index from -1, by row (VarYCoef :897-909) or by column
struct VarCoefBase : Coefficients {
  Splines therm, cap, tempcond;
  VarCoefBase(Splines k, Splines c, Splines a) : therm(std::move(k)), cap(std::move(c)), tempcond(std::move(a)) {}
  virtual int mode() const = 0;
  CoefMaps maps(int cols, int rows) const override {
    const size_t n = (size_t)(mode() == GS_MAP_PER_ROW ? rows : cols) + 1;
    if (therm.size() != n || cap.size() != n || tempcond.size() != n)
      throw std::invalid_argument("one spline per index from -1");
    CoefMaps m;
    m[GS_SEL_APPLY] = {mode(), detail::ids_of(tempcond, n)};
    m[GS_SEL_THERM] = {mode(), detail::ids_of(therm, n)};
    m[GS_SEL_CAP] = {mode(), detail::ids_of(cap, n)};
    m[GS_SEL_TEMPCOND] = {mode(), detail::ids_of(tempcond, n)};
    return m;
  }
};
struct VarYCoef : VarCoefBase {
  using VarCoefBase::VarCoefBase;
  int mode() const override { return GS_MAP_PER_ROW; }
};
// per-column splines, index col + 1: what VarXCoef's constructor arguments describe
struct VarXCoefByColumn : VarCoefBase {
  using VarCoefBase::VarCoefBase;
  int mode() const override { return GS_MAP_PER_COL; }
};
// The reference's VarXCoef, literally: its arrays are built per COLUMN interval (cols + 1 entries, :878-893) but every
// member indexes them by y + 1 (:862-874): a per-row map over the first rows + 1 entries, and an
// ArrayIndexOutOfBoundsException (std::out_of_range here) when rows > cols.
struct VarXCoef : VarCoefBase {
  using VarCoefBase::VarCoefBase;
  int mode() const override { return GS_MAP_PER_ROW; }
  CoefMaps maps(int, int rows) const override {
    const size_t n = (size_t)rows + 1;
    if (std::min({therm.size(), cap.size(), tempcond.size()}) < n)
      throw std::out_of_range("VarXCoef indexes its per-column arrays by y + 1 (TwoDGrid.scala:862-874)");
    CoefMaps m;
    m[GS_SEL_APPLY] = {GS_MAP_PER_ROW, detail::ids_of(tempcond, n)};
    m[GS_SEL_THERM] = {GS_MAP_PER_ROW, detail::ids_of(therm, n)};
    m[GS_SEL_CAP] = {GS_MAP_PER_ROW, detail::ids_of(cap, n)};
    m[GS_SEL_TEMPCOND] = {GS_MAP_PER_ROW, detail::ids_of(tempcond, n)};
    return m;
  }
};

// PatchXCoef / PatchYCoef (:895-1089): per-interval arrays that apply inside rangeX x rangeY only (`contains`,
// closed-open index intervals); only meaningful inside a Patched*Coef
struct PatchCoefBase {
  Splines therm, cap, tempcond;
  std::pair<int, int> rangeX, rangeY;
  PatchCoefBase(Splines k, Splines c, Splines a, std::pair<int, int> rx, std::pair<int, int> ry)
      : therm(std::move(k)), cap(std::move(c)), tempcond(std::move(a)), rangeX(rx), rangeY(ry) {}
  virtual ~PatchCoefBase() = default;
  virtual bool applyByX() const = 0;
  bool contains(int x, int y) const { return rangeX.first <= x && x < rangeX.second && rangeY.first <= y && y < rangeY.second; }
  const AlwaysDefinedSpline &member(int sel, int x, int y) const {
    if (sel == GS_SEL_THERM) return *therm.at((size_t)y + 1);
    if (sel == GS_SEL_CAP) return *cap.at((size_t)y + 1);
    if (sel == GS_SEL_TEMPCOND) return *tempcond.at((size_t)y + 1);
    return *tempcond.at((size_t)(applyByX() ? x : y) + 1);
  }
};
struct PatchXCoef : PatchCoefBase {   // :895-947 -- every member, `apply` included, indexes by y + 1
  using PatchCoefBase::PatchCoefBase;
  bool applyByX() const override { return false; }
};
struct PatchYCoef : PatchCoefBase {   // :1019-1056 -- `apply` indexes by x + 1 (:1056)
  using PatchCoefBase::PatchCoefBase;
  bool applyByX() const override { return true; }
};
// PatchedXCoef(xCoefs, path) / PatchedYCoef(yCoefs, path) (:1091-1137): the patch where it `contains` the node, the base
// map elsewhere; flattened to DENSE id maps (a Scala host calls coefs(col, row) for every index from -1)
struct PatchedCoef : Coefficients {
  const VarCoefBase &base;
  const PatchCoefBase &path;
  PatchedCoef(const VarCoefBase &b, const PatchCoefBase &p) : base(b), path(p) {}
  CoefMaps maps(int cols, int rows) const override {
    const CoefMaps bm = base.maps(cols, rows);
    CoefMaps m;
    for (int sel = 0; sel < 4; ++sel) {
      m[sel].mode = GS_MAP_DENSE;
      m[sel].ids.resize((size_t)(rows + 1) * (cols + 1));
      for (int y = -1; y < rows; ++y)
        for (int x = -1; x < cols; ++x)
          m[sel].ids[(size_t)(y + 1) * (cols + 1) + (x + 1)] =
              path.contains(x, y) ? path.member(sel, x, y).id : bm[sel].ids[(size_t)(bm[sel].mode == GS_MAP_PER_ROW ? y : x) + 1];
    }
    return m;
  }
};
using PatchedXCoef = PatchedCoef;
using PatchedYCoef = PatchedCoef;
// arbitrary coefs(col, row): id arrays of shape (rows + 1, cols + 1), entry [row + 1][col + 1]
struct DenseCoef : Coefficients {
  CoefMaps m;
  DenseCoef(std::vector<int> apply, std::vector<int> therm, std::vector<int> cap, std::vector<int> tempcond) {
    m[GS_SEL_APPLY] = {GS_MAP_DENSE, std::move(apply)};
    m[GS_SEL_THERM] = {GS_MAP_DENSE, std::move(therm)};
    m[GS_SEL_CAP] = {GS_MAP_DENSE, std::move(cap)};
    m[GS_SEL_TEMPCOND] = {GS_MAP_DENSE, std::move(tempcond)};
  }
  CoefMaps maps(int cols, int rows) const override {
    for (auto &e : m)
      if (e.ids.size() != (size_t)(rows + 1) * (cols + 1)) throw std::invalid_argument("(rows + 1) x (cols + 1) ids");
    return m;
  }
};

// ---- approximation.TwoDGrid --------------------------------------------------------------------------------
class TwoDGrid {
 public:
  struct BoundSpec {
    int rep = GS_SPARSE, kind = GS_TEMP;  // (One | Many, Temp | Flow)  A/TwoDGrid.scala:521-551
  };
  // one side's bound (Temperature / HeatFlow x Dense / Sparse, :577-728): update(value[s]) and get
  class Bound {
   public:
    void update(double v) { check(gs_grid2d_set_bound(g_->h_, side_, kind, rep, &v, 1)); }
    void update(const std::vector<double> &v) { check(gs_grid2d_set_bound(g_->h_, side_, kind, rep, v.data(), (int)v.size())); }
    std::vector<double> get() const {
      std::vector<double> out((size_t)(side_ == GS_UPP || side_ == GS_LOW ? g_->cols : g_->rows));
      check(gs_grid2d_get_bound(g_->h_, side_, out.data(), (int)out.size()));
      return out;
    }
    double get(int i) const { return get().at((size_t)i); }
    int kind = GS_TEMP, rep = GS_SPARSE;   // assignable (a bound may change kind, e.g. Temperature -> HeatFlow)

   private:
    friend class TwoDGrid;
    TwoDGrid *g_ = nullptr;
    int side_ = 0;
  };
  struct Bounds {
    Bound upp, low, right, left;
  };

  TwoDGrid(Context &ctx, const XDim &x_, const YDim &y_, BoundSpec upp, BoundSpec low, BoundSpec right, BoundSpec left,
           const Coefficients &coefs)
      : cols(x_.colsNum()), rows(y_.rowsNum()), x(x_), y(y_) {
    auto xv = x.values(), yv = y.values();
    check(gs_grid2d_create(ctx.handle(), cols, rows, x.t, y.t, xv.data(), yv.data(), &h_));
    const BoundSpec spec[4] = {upp, low, right, left};
    Bound *b[4] = {&bounds.upp, &bounds.low, &bounds.right, &bounds.left};
    for (int s = 0; s < 4; ++s) {
      b[s]->g_ = this;
      b[s]->side_ = s;
      b[s]->kind = spec[s].kind;
      b[s]->rep = spec[s].rep;
      b[s]->update(0.0);
    }
    setCoefs(coefs);
    for (int s = 0; s < 4; ++s) noHeatFlow(s);  // the class body ends with noHeatFlow on all sides (:446-449)
  }
  ~TwoDGrid() { gs_grid2d_destroy(h_); }
  TwoDGrid(const TwoDGrid &) = delete;
  TwoDGrid &operator=(const TwoDGrid &) = delete;

  void setCoefs(const Coefficients &coefs) {
    const CoefMaps m = coefs.maps(cols, rows);
    for (int sel = 0; sel < 4; ++sel) check(gs_grid2d_set_map(h_, sel, m[sel].mode, m[sel].ids.data()));
  }
  Bound &bound(int side) { return side == GS_UPP ? bounds.upp : side == GS_LOW ? bounds.low : side == GS_RIGHT ? bounds.right : bounds.left; }
  void updateBound(int side, double v) { bound(side).update(v); }
  void updateBound(int side, const std::vector<double> &v) { bound(side).update(v); }
  void fill(double v) { check(gs_grid2d_fill(h_, v)); }                       // grid *= v
  void xIter(double time) { check(gs_grid2d_xiter(h_, time)); }
  void yIter(double time) { check(gs_grid2d_yiter(h_, time)); }
  void update() { check(gs_grid2d_update(h_)); }
  void iteration(double time, int nsteps = 1) { check(gs_grid2d_iteration(h_, time, nsteps)); }
  void yIterUpdate(double time) { check(gs_grid2d_yiter_update(h_, time)); }   // yIter + Grid.update in one pass
  // the y sweep of a row-sharded grid, split around the ranks' exchange (device pointers; nullptr = physical boundary)
  void yIterUpdateBegin(double time, const void *halo_prev, const void *halo_next, void *exported) {
    check(gs_grid2d_yiter_update_begin(h_, time, halo_prev, halo_next, exported));
  }
  void yIterUpdateEnd(double time, const void *gathered, int nranks, int rank) {
    check(gs_grid2d_yiter_update_end(h_, time, gathered, nranks, rank));
  }
  void noHeatFlow(int side) { check(gs_grid2d_no_heat_flow(h_, side)); }
  void updatePredicted() { check(gs_grid2d_update_predicted(h_)); }
  void setSolver(int solver) { check(gs_grid2d_set_solver(h_, solver)); }
  unsigned status() const {
    uint32_t f = 0;
    check(gs_grid2d_status(h_, &f));
    return f;
  }
  void clearStatus() { check(gs_grid2d_clear_status(h_)); }
  double bytesPerStep() const {
    double b = 0.0;
    check(gs_grid2d_bytes_per_step(h_, &b));
    return b;
  }
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
  void setArray(int which, const std::vector<double> &v) {
    if (v.size() != (size_t)cols * rows) throw std::invalid_argument("rows * cols values");
    check(gs_grid2d_upload(h_, which, v.data()));
  }
  // device-to-device copies and caller-owned device arrays
  void copyFromDevice(int which, const void *dptr) { check(gs_grid2d_copy_from_device(h_, which, dptr)); }
  void copyToDevice(int which, void *dptr) const { check(gs_grid2d_copy_to_device(h_, which, dptr)); }
  void bindDevice(int which, void *dptr) { check(gs_grid2d_bind_device(h_, which, dptr)); }
  void *devicePtr(int which) const {
    void *p = nullptr;
    check(gs_grid2d_device_ptr(h_, which, &p));
    return p;
  }
  // slices of `grid` (or another array) and coordinate-wise fills (A/TwoDGrid.scala:29-58, 260-352)
  std::vector<double> row(int rowNum, int which = GS_GRID) const {
    std::vector<double> out((size_t)cols);
    check(gs_grid2d_get_row(h_, which, rowNum, out.data()));
    return out;
  }
  std::vector<double> col(int colNum, int which = GS_GRID) const {
    std::vector<double> out((size_t)rows);
    check(gs_grid2d_get_col(h_, which, colNum, out.data()));
    return out;
  }
  std::vector<double> left(int colNum = 0) const { return col(colNum); }
  std::vector<double> right(int colNumFromRight = 0) const { return col(cols - 1 - colNumFromRight); }
  std::vector<double> upper(int rowNum = 0) const { return row(rowNum); }
  std::vector<double> lower(int rowFromLast = 0) const { return row(rows - 1 - rowFromLast); }
  void updateRow(int rowNum, const std::vector<double> &v) { check(gs_grid2d_set_row(h_, rowNum, v.data())); }
  void updateColumn(int colNum, const std::vector<double> &v) { check(gs_grid2d_set_col(h_, colNum, v.data())); }
  template <class F>
  void updateX(F f) {   // grid(col, row) = f(x_col)
    std::vector<double> v((size_t)cols);
    for (int c = 0; c < cols; ++c) v[c] = f(x.range[c]);
    check(gs_grid2d_update_x(h_, v.data()));
  }
  template <class F>
  void updateY(F f) {   // grid(col, row) = f(y_row)
    std::vector<double> v((size_t)rows);
    for (int r = 0; r < rows; ++r) v[r] = f(y.range[r]);
    check(gs_grid2d_update_y(h_, v.data()));
  }
  template <class F>
  void updateX(const XDim &, F f) { updateX(f); }
  template <class F>
  void updateY(const YDim &, F f) { updateY(f); }
  // text output: TwoDGrid.write (:62-82) and toString (:86-133); defined after the number formatters below
  void write(std::ostream &out, const std::string &newline = "\n") const;
  std::string toString(const std::string &sep = "\n") const;

  const int cols, rows;
  const XDim x;
  const YDim y;
  Bounds bounds;
  gs_grid2d *handle() const { return h_; }

 private:
  gs_grid2d *h_ = nullptr;
};

// ---- ONE TwoDGrid on several GPUs of a box, driven by this process (gs_sharded2d_*): rows sharded over `devices`, the
// step's exchange happens inside the kernels over NVLink peer memory ---------------------------------------------
class ShardedTwoDGrid {
 public:
  ShardedTwoDGrid(const XDim &x, const YDim &y, const std::vector<int> &devices, TwoDGrid::BoundSpec upp,
                  TwoDGrid::BoundSpec low, TwoDGrid::BoundSpec right, TwoDGrid::BoundSpec left)
      : cols(x.colsNum()), rows(y.rowsNum()) {
    auto xv = x.values(), yv = y.values();
    check(gs_sharded2d_create((int)devices.size(), devices.data(), cols, rows, x.t, y.t, xv.data(), yv.data(), &h_));
    spec_[GS_UPP] = upp; spec_[GS_LOW] = low; spec_[GS_RIGHT] = right; spec_[GS_LEFT] = left;
    for (int s = 0; s < 4; ++s) setBound(s, 0.0);
  }
  ~ShardedTwoDGrid() { gs_sharded2d_destroy(h_); }
  ShardedTwoDGrid(const ShardedTwoDGrid &) = delete;
  // AlwaysDefinedSpline(spl) in every device's pool (clamps by the reference's literal rule, computed on `ctx`)
  int registerSpline(const Spline &spl, bool true_upper = false, Context &ctx = Context::instance()) {
    AlwaysDefinedSpline a(ctx, spl, true_upper);
    int id = -1;
    check(gs_sharded2d_spline_create(h_, spl.kind, spl.size(), spl.knots.data(), spl.x0.data(), spl.coefs.data(), a.lowerX,
                                     a.upperX, a.lowerY, a.upperY, &id));
    return id;
  }
  void setConstantCoef(const Spline &therm, const Spline &cap, const Spline &tempcond) {
    const int t = registerSpline(therm), c = registerSpline(cap), a = registerSpline(tempcond);
    setMap(GS_SEL_APPLY, GS_MAP_SINGLE, {a});
    setMap(GS_SEL_THERM, GS_MAP_SINGLE, {t});
    setMap(GS_SEL_CAP, GS_MAP_SINGLE, {c});
    setMap(GS_SEL_TEMPCOND, GS_MAP_SINGLE, {a});
  }
  void setMap(int selector, int mode, const std::vector<int> &ids) { check(gs_sharded2d_set_map(h_, selector, mode, ids.data())); }
  void setBound(int side, double v) { check(gs_sharded2d_set_bound(h_, side, spec_[side].kind, spec_[side].rep, &v, 1)); }
  void setBound(int side, const std::vector<double> &v) {
    check(gs_sharded2d_set_bound(h_, side, spec_[side].kind, spec_[side].rep, v.data(), (int)v.size()));
  }
  void fill(double v) { check(gs_sharded2d_fill(h_, v)); }
  void upload(int which, const std::vector<double> &v) {
    if (v.size() != (size_t)cols * rows) throw std::invalid_argument("rows * cols values");
    check(gs_sharded2d_upload(h_, which, v.data()));
  }
  std::vector<double> download(int which = GS_GRID) const {
    std::vector<double> out((size_t)cols * rows);
    check(gs_sharded2d_download(h_, which, out.data()));
    return out;
  }
  void iteration(double time, int nsteps = 1) { check(gs_sharded2d_iteration(h_, time, nsteps)); }
  double iterationTimed(double time, int nsteps = 1) {   // device-measured milliseconds (max over the devices)
    double ms = 0.0;
    check(gs_sharded2d_iteration_timed(h_, time, nsteps, &ms));
    return ms;
  }
  void synchronize() { check(gs_sharded2d_synchronize(h_)); }
  void setOption(const char *name, long long value) { check(gs_sharded2d_set_option(h_, name, value)); }
  int deviceCount() const {
    int n = 0;
    check(gs_sharded2d_device_count(h_, &n));
    return n;
  }
  unsigned status() const {
    uint32_t f = 0;
    check(gs_sharded2d_status(h_, &f));
    return f;
  }
  unsigned long long launchCount() const {
    uint64_t c = 0;
    check(gs_sharded2d_launch_count(h_, &c));
    return c;
  }
  const int cols, rows;

 private:
  gs_sharded2d *h_ = nullptr;
  TwoDGrid::BoundSpec spec_[4];
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


// ---- text output: TwoDGrid.write / toString (A/TwoDGrid.scala:62-133), XDim.toString
// (A/XDim.scala:17-23), OneDGrid.writeGrid (A/OneDGrid.scala:37-63).  The JVM's two number formatters are restated
// from their documented behaviour, like gridsplines_b200/api.py does.
// java.text.NumberFormat.getInstance(Locale.ROOT) with setMaximumFractionDigits(n): DecimalFormat "#,##0.###",
// RoundingMode.HALF_EVEN applied to the EXACT binary value, grouping by three, no trailing zeros, "-0" for a negative
// value that rounds to zero, "NaN" and the infinity sign for non-finite values.
inline std::string java_number_format(double x, int max_fraction_digits) {
  if (x != x) return "NaN";
  if (std::isinf(x)) return std::string(x < 0 ? "-" : "") + "\xE2\x88\x9E";
  const bool neg = std::signbit(x);
  std::string digits(400, '\0');
  // glibc's printf rounds the exact binary value, ties to even (round-to-nearest mode)
  digits.resize((size_t)std::snprintf(&digits[0], digits.size(), "%.*f", max_fraction_digits, std::fabs(x)));
  std::string whole = digits, frac;
  const size_t dot = digits.find('.');
  if (dot != std::string::npos) {
    whole = digits.substr(0, dot);
    frac = digits.substr(dot + 1);
    while (!frac.empty() && frac.back() == '0') frac.pop_back();
  }
  std::string grouped;
  for (size_t i = 0; i < whole.size(); ++i) {
    if (i && (whole.size() - i) % 3 == 0) grouped += ',';
    grouped += whole[i];
  }
  return std::string(neg ? "-" : "") + grouped + (frac.empty() ? "" : "." + frac);
}
// Scala's f"$x%.3f" = java.util.Formatter %.nf: RoundingMode.HALF_UP applied to the decimal digits of Double.toString
// (the shortest digits that round-trip), always `digits` decimals
inline std::string java_fixed(double x, int digits) {
  if (x != x) return "NaN";
  if (std::isinf(x)) return x < 0 ? "-Infinity" : "Infinity";
  char buf[400];
  auto r = std::to_chars(buf, buf + sizeof buf, std::fabs(x), std::chars_format::fixed);   // shortest round-trip digits
  std::string t(buf, r.ptr);
  std::string whole = t, frac;
  const size_t dot = t.find('.');
  if (dot != std::string::npos) {
    whole = t.substr(0, dot);
    frac = t.substr(dot + 1);
  }
  bool up = (int)frac.size() > digits && frac[(size_t)digits] >= '5';
  frac.resize((size_t)digits, '0');
  std::string all = whole + frac;   // decimal point after whole.size() digits
  for (int i = (int)all.size() - 1; up && i >= 0; --i) {
    if (all[(size_t)i] == '9') all[(size_t)i] = '0';
    else { ++all[(size_t)i]; up = false; }
  }
  if (up) all.insert(all.begin(), '1');
  const size_t w = all.size() - (size_t)digits;
  return std::string(std::signbit(x) ? "-" : "") + all.substr(0, w) + (digits ? "." + all.substr(w) : "");
}
// XDim.toString (A/XDim.scala:17-23)
inline std::string dim_to_string(const Dim &d, const std::string &sep = "\n") {
  std::string out = "X/Y\t\t|\t" + java_fixed(d.low, 3) + "\t|\t";
  for (size_t i = 0; i < d.range.size(); ++i) out += (i ? "\t" : "") + java_fixed(d.range[i], 3);
  return out + "\t|\t" + java_fixed(d.upp, 3) + "\t|\t..." + sep;
}
// TwoDGrid.write(writer) (:62-82): per row `format(rowCoord) format(g(row, 0)) ...` + BufferedWriter.newLine()
inline void TwoDGrid::write(std::ostream &out, const std::string &newline) const {
  const std::vector<double> a = array(GS_GRID);
  for (int r = 0; r < rows; ++r) {
    out << java_number_format(y.range[(size_t)r], 3);
    for (int c = 0; c < cols; ++c) out << ' ' << java_number_format(a[(size_t)r * cols + c], 3);
    out << newline;
  }
}
// TwoDGrid.toString (:86-133), quirks included: the lower-bound line comes first and ends with y.upp, the right bound is
// printed with four decimals, the upper-bound line starts with y.upp
inline std::string TwoDGrid::toString(const std::string &sep) const {
  auto f3 = [](double v) { return java_fixed(v, 3); };
  auto joined = [&](const double *v, int n) {
    std::string o;
    for (int i = 0; i < n; ++i) o += (i ? "\t" : "") + f3(v[i]);
    return o;
  };
  const std::string xc = dim_to_string(x, sep);
  const std::vector<double> low = bounds.low.get(), upp = bounds.upp.get(), left = bounds.left.get(), right = bounds.right.get();
  const std::vector<double> a = array(GS_GRID);
  std::string out = xc;
  out += f3(y.low) + "\t|\t...\t\t|\t" + joined(low.data(), cols) + "\t|\t...\t\t|\t" + f3(y.upp) + sep;
  for (int r = 0; r < rows; ++r) {
    const double yc = y.range[(size_t)r];
    out += f3(yc) + "\t|\t" + f3(left[(size_t)r]) + "\t|\t" + joined(&a[(size_t)r * cols], cols) + "\t|\t" +
           java_fixed(right[(size_t)r], 4) + "\t|\t" + f3(yc) + sep;
  }
  out += f3(y.upp) + "\t|\t...\t\t|\t" + joined(upp.data(), cols) + "\t|\t...\t\t|\t" + f3(y.upp) + sep;
  return out + xc;
}
// OneDGrid.writeGrid(at: LocalDateTime, path) (:37-63): file `grid MONTH Dd Hh YEAR.dat`, a two-line header and one
// `coord<TAB>temperature` line per node, every element followed by "\r\n" AND BufferedWriter.newLine() (sic).  Writes
// the text to `out`, returns the file name the reference would use (month 1..12).
inline std::string oned_write_grid(const OneDGrid &g, std::ostream &out, int year, int month, int day, int hour,
                                   const std::string &newline = "\n") {
  static const char *MONTHS[12] = {"JANUARY", "FEBRUARY", "MARCH", "APRIL", "MAY", "JUNE", "JULY", "AUGUST", "SEPTEMBER",
                                   "OCTOBER", "NOVEMBER", "DECEMBER"};
  const std::vector<double> t = g.grid();
  out << "# One dimension grid output results\r\n# Coords \t Temperatures\r\n" << newline;
  for (size_t i = 0; i < g.rangeX().size(); ++i)
    out << java_number_format(g.rangeX()[i], 2) << '\t' << java_number_format(t[i], 3) << "\r\n" << newline;
  return std::string("grid ") + MONTHS[(month - 1) % 12] + " " + std::to_string(day) + "d " + std::to_string(hour) + "h " +
         std::to_string(year) + ".dat";
}

}  // namespace gridsplines
