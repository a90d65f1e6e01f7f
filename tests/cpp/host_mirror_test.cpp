// Runs config C1 (KAT R2) and the GridTest scenario (KAT R1, AT/GridTest.scala:17-66) through the C++ host mirror
// and prints bit patterns; tests/test_cpp_mirror.py compares them with the oracle's values.  `host_mirror_test cpu`
// prints what needs no device (number formatters, spline factories); without an argument the device sections follow
// (spline algebra, per-node splines, per-line ends, ragged lines, the Coefficients hierarchy, text output), compared
// with the Python mirror's results for the same inputs.
#include <array>
#include <cinttypes>
#include <cstdio>
#include <cstring>
#include <sstream>

#include "../../gridsplines_b200/host/gridsplines.hpp"

using namespace gridsplines;

static void print(const char *name, double v) {
  uint64_t b;
  memcpy(&b, &v, 8);
  printf("%s %.17g 0x%016" PRIx64 "\n", name, v, b);
}

static void print_array(const char *name, const std::vector<double> &v) {
  printf("%s", name);
  for (double x : v) {
    uint64_t b;
    memcpy(&b, &x, 8);
    printf(" %016" PRIx64, b);
  }
  printf("\n");
}
static void print_spline(const char *name, const Spline &s) {
  printf("%s_kind %d\n", name, s.kind);
  print_array((std::string(name) + "_knots").c_str(), s.knots);
  print_array((std::string(name) + "_x0").c_str(), s.x0);
  print_array((std::string(name) + "_coefs").c_str(), s.coefs);
}
static void print_text(const char *name, const std::string &t) {   // bytes as hex: tabs and newlines survive
  printf("%s ", name);
  for (unsigned char c : t) printf("%02x", c);
  printf("\n");
}
static const double RHO[11] = {0.598, 0.826, 1.121, 1.496, 1.966, 2.547, 3.258, 4.122, 5.156, 6.397, 7.862};

static void host_only() {
  const double xs[] = {0.0, -0.0, 1.0, -1.5, 2.5, 0.125, 0.0005, -0.0005, 1234567.891, 0.1 + 0.2, 1e-7, 54.44328715390563,
                       999.9995, 1e15 / 3.0, -2.675, 8.345, 0.045, 1.0005, 123456.7895};
  int i = 0;
  for (double x : xs) {
    printf("nf3_%d %s\n", i, java_number_format(x, 3).c_str());
    printf("nf2_%d %s\n", i, java_number_format(x, 2).c_str());
    printf("fx3_%d %s\n", i, java_fixed(x, 3).c_str());
    printf("fx4_%d %s\n", i, java_fixed(x, 4).c_str());
    ++i;
  }
  printf("nf_nan %s\n", java_number_format(NAN, 3).c_str());
  printf("fx_inf %s\n", java_fixed(-INFINITY, 3).c_str());
  Spline::Points pts = {{0.0, 1.0}, {1.0, 3.0}, {2.5, 2.0}, {4.0, 5.0}, {6.0, 4.0}, {7.5, 4.5}};
  print_spline("lag2", Spline::lagrange2(pts));
  print_spline("lag3", Spline::lagrange3(pts));
  Spline::Points rho;
  for (int k = 0; k < 11; ++k) rho.push_back({100.0 + 10.0 * k, RHO[k]});
  print_spline("m1h", Spline::m1Hermite3(rho));
  print_spline("fh", Spline::fHermite3(rho));
  print_spline("lines", Spline::lines(rho));
  print_spline("const", Spline::constant(2.5));
  XDim x(0.05, std::vector<double>{0.1, 0.2004, 0.30049999, 0.4}, 0.5, GS_RADIAL);
  print_text("dim_str", dim_to_string(x, "\n"));
}

static void device_sections(Context &ctx) {
  Spline::Points kp, cp;
  for (int i = 0; i < 11; ++i) {
    kp.push_back({-20.0 + 10.0 * i, 1.0 + 0.1 * RHO[i]});
    cp.push_back({-20.0 + 10.0 * i, 2.0e6 + 1.0e5 * RHO[i]});
  }
  const Spline k = Spline::m1Hermite3(kp), c = Spline::m1Hermite3(cp);
  {  // spline algebra (Spline./ :207-210): the quotient resampled at the merged knots
    auto q = k.combine(c, [](double a, double b) { return a / b; }, nullptr, ctx);
    print_spline("quot", *q);
    auto sum = k.combine(Spline::lines(cp), [](double a, double b) { return a + b; }, Spline::Factory(Spline::fHermite3), ctx);
    print_spline("sum", *sum);
  }
  std::vector<double> rng(40);
  for (int i = 0; i < 40; ++i) rng[i] = 0.01 * (i + 1);
  AlwaysDefinedSpline a1(ctx, Spline::m1Hermite3({{-20.0, 1.0e-6}, {10.0, 1.3e-6}, {40.0, 1.9e-6}, {80.0, 2.2e-6}})),
      a2(ctx, Spline::constant(1.5e-6)), a3(ctx, Spline::lines({{-20.0, 1.1e-6}, {30.0, 1.4e-6}, {80.0, 2.0e-6}}));
  {  // per-node (left, right) conductivity pairs (passion.iteration overload 1)
    std::vector<BatchedOneDGrid::Pair> pairs;
    const AlwaysDefinedSpline *pool[3] = {&a1, &a2, &a3};
    for (int i = 0; i < 40; ++i) pairs.push_back({pool[i % 3], pool[(i + 1) % 3]});
    OneDGrid g(ctx, GS_ORTHO, 0.0, rng, 0.41, pairs, 1.0);
    std::vector<double> init(40);
    for (int i = 0; i < 40; ++i) init[i] = 5.0 + 0.37 * i;
    g.setGrid(init);
    for (int s = 0; s < 3; ++s) g.step(900.0, 20.0, 4.0);
    print_array("pernode_grid", g.grid());
    print_array("pernode_predict", g.predict());
    printf("pernode_status %u\n", g.status());
  }
  {  // five lines, one pair of end temperatures per line
    BatchedOneDGrid g(ctx, GS_RADIAL, 0.005, rng, 0.41, a1, 1.0, 5);
    std::vector<double> init(200), lt(5), rt(5);
    for (int i = 0; i < 200; ++i) init[i] = 3.0 + 0.11 * (i % 37);
    for (int l = 0; l < 5; ++l) { lt[l] = 10.0 + 3.0 * l; rt[l] = 1.0 + l; }
    g.upload(GS_GRID, init);
    g.step(900.0, lt, rt, 2);
    print_array("batched_grid", g.download(GS_GRID));
    print_array("batched_interleaved", g.download(GS_GRID, GS_INTERLEAVED));
  }
  {  // ragged lines in one flattened array of 30 values
    std::vector<std::vector<int>> lines = {{0, 1, 2, 3, 4, 5, 6}, {10, 12, 14, 16, 18}, {29, 28, 27, 26, 25, 24, 23, 22, 21}};
    std::vector<std::vector<double>> pd;
    for (auto &l : lines) {
      std::vector<double> r(l.size()), out(2 * l.size());
      for (size_t i = 0; i < l.size(); ++i) r[i] = 0.02 * (i + 1);
      check(gs_predef(ctx.handle(), GS_ORTHO, 0.0, r.data(), (int)r.size(), 0.02 * (l.size() + 1), 1.0, out.data()));
      pd.push_back(out);
    }
    RaggedLines g(ctx, 30, lines, pd, a3);
    std::vector<double> init(30);
    for (int i = 0; i < 30; ++i) init[i] = 2.0 + 0.5 * i;
    g.set(GS_GRID, init);
    g.set(GS_PREDICT, init);
    g.iteration(900.0, {20.0, 15.0, 12.0}, {4.0, 3.0, 1.0});
    print_array("ragged_result", g.result());
    printf("ragged_status %u\n", g.status());
  }
  {  // the Coefficients hierarchy: per-row splines with a patch, and the text output of the grid
    XDim x(0.0, [](double v) { return v + 0.01; }, 0.125, GS_ORTHO);
    YDim y(0.0, [](double v) { return v + 0.02; }, 0.21, GS_ORTHO);
    const int cols = x.colsNum(), rows = y.rowsNum();
    std::vector<std::unique_ptr<AlwaysDefinedSpline>> own;
    auto mk = [&](double scale) {
      Spline::Points p;
      for (int i = 0; i < 11; ++i) p.push_back({-20.0 + 10.0 * i, scale * (1.0 + 0.1 * RHO[i])});
      own.emplace_back(new AlwaysDefinedSpline(ctx, Spline::m1Hermite3(p)));
      return own.back().get();
    };
    Splines th, cp2, tc, pth, pcp, ptc;
    for (int r = 0; r <= rows; ++r) {
      th.push_back(mk(1.0 + 0.05 * r));
      cp2.push_back(mk(2.0e6));
      tc.push_back(mk((1.0 + 0.05 * r) / 2.0e6));
      pth.push_back(mk(3.0));
      pcp.push_back(mk(1.0e6));
      ptc.push_back(mk(3.0e-6));
    }
    VarYCoef base(th, cp2, tc);
    PatchXCoef patch(pth, pcp, ptc, {2, 7}, {1, 6});
    PatchedXCoef coefs(base, patch);
    TwoDGrid::BoundSpec manyTemp{GS_DENSE, GS_TEMP}, oneTemp{GS_SPARSE, GS_TEMP}, manyFlow{GS_DENSE, GS_FLOW};
    TwoDGrid g(ctx, x, y, manyTemp, oneTemp, manyTemp, manyFlow, coefs);
    g.setSolver(GS_SOLVER_THOMAS);
    std::vector<double> init((size_t)cols * rows);
    for (size_t i = 0; i < init.size(); ++i) init[i] = 7.0 + 0.013 * (double)((i * 7919) % 101);
    g.setArray(GS_GRID, init);
    g.setArray(GS_PREDICT, init);
    g.bounds.upp.update(20.0);
    g.bounds.low.update(4.0);
    g.bounds.right.update(9.5);
    g.bounds.left.update(50.0);
    g.iteration(900.0, 2);
    printf("patched_dims %d %d\n", cols, rows);
    print_array("patched_grid", g.array(GS_GRID));
    print_array("patched_predict", g.array(GS_PREDICT));
    print_array("patched_left_bound", g.bounds.left.get());
    printf("patched_status %u\n", g.status());
    print_text("patched_tostring", g.toString("\n"));
    std::ostringstream w;
    g.write(w, "\n");
    print_text("patched_write", w.str());
  }
  {  // one TwoDGrid behind gs_sharded2d_* (here: one device), constant coefficients
    std::vector<double> xr(64), yr(48);
    for (int i = 0; i < 64; ++i) xr[i] = 0.01 * (i + 1);
    for (int i = 0; i < 48; ++i) yr[i] = 0.02 * (i + 1);
    XDim x(0.0, xr, 0.65, GS_ORTHO);
    YDim y(0.0, yr, 0.98, GS_ORTHO);
    TwoDGrid::BoundSpec manyTemp{GS_DENSE, GS_TEMP};
    ShardedTwoDGrid s(x, y, {0}, manyTemp, manyTemp, manyTemp, manyTemp);
    s.setConstantCoef(Spline::constant(1.5), Spline::constant(2.2e6), Spline::constant(1.5 / 2.2e6));
    s.setBound(GS_UPP, 20.0);
    s.setBound(GS_LOW, 4.0);
    s.setBound(GS_RIGHT, 9.0);
    s.setBound(GS_LEFT, 30.0);
    std::vector<double> init(64 * 48);
    for (size_t i = 0; i < init.size(); ++i) init[i] = 7.0 + 0.01 * (double)((i * 131) % 97);
    s.upload(GS_GRID, init);
    s.upload(GS_PREDICT, init);
    s.iteration(900.0, 2);
    print_array("sharded_grid", s.download(GS_GRID));
    printf("sharded_status %u %d\n", s.status(), s.deviceCount());
  }
}

int main(int argc, char **argv) {
  if (argc > 1 && !strcmp(argv[1], "cpu")) {
    host_only();
    return 0;
  }
  Context ctx(0);
  {  // C1: OneDGrid[Radial], 200 nodes, Spline.const(1e-6), step(900, 20, 4) x 1000
    std::vector<double> rng(200);
    for (int i = 0; i < 200; ++i) rng[i] = 0.05 + 0.01 * (i + 1);
    AlwaysDefinedSpline k(ctx, Spline::constant(1e-6));
    OneDGrid g(ctx, GS_RADIAL, 0.05, rng, rng[199] + 0.01, k, 1.0);
    for (int s = 0; s < 1000; ++s) g.step(900.0, 20.0, 4.0);
    auto v = g.grid();
    print("c1_grid0", v[0]);
    print("c1_grid99", v[99]);
    print("c1_grid199", v[199]);
  }
  {  // README known answers through the device evaluator
    Spline::Points pts;
    double rho[11] = {0.598, 0.826, 1.121, 1.496, 1.966, 2.547, 3.258, 4.122, 5.156, 6.397, 7.862};
    for (int i = 0; i < 11; ++i) pts.push_back({100.0 + 10.0 * i, rho[i]});
    AlwaysDefinedSpline s(ctx, Spline::m1Hermite3(pts), /*true_upper=*/true);
    print("readme_135", s(135.0));
    print("readme_area_150_185", s.average(150.0, 185.0) * 35.0);
  }
  {  // GridTest: XDim[Radial] 5e-5 .. 25 (x1.1), YDim[Ortho] 0 .. 92 (+2), k = 1.5, c = 2.2e6, left heat flow
    XDim x(5e-5, [](double v) { return v * 1.1; }, 25.0, GS_RADIAL);
    YDim y(0.0, [](double v) { return v + 2.0; }, 92.0, GS_ORTHO);
    AlwaysDefinedSpline k(ctx, Spline::constant(1.5)), c(ctx, Spline::constant(2.2e6)),
        a(ctx, Spline::constant(1.5 / 2.2e6));
    ConstantCoef coefs{k, c, a};
    TwoDGrid::BoundSpec manyTemp{GS_DENSE, GS_TEMP}, manyFlow{GS_DENSE, GS_FLOW};
    TwoDGrid g(ctx, x, y, manyTemp, manyTemp, manyTemp, manyFlow, coefs);
    g.setSolver(GS_SOLVER_THOMAS);
    g.fill(7.0);
    const double heat = 4500.0 / (92.0 * 2.0 * 3.141592653589793);
    for (int it = 0; it < 195; ++it) {
      g.updateBound(GS_LEFT, heat);
      g.noHeatFlow(GS_RIGHT);
      g.noHeatFlow(GS_UPP);
      g.noHeatFlow(GS_LOW);
      g.iteration(900.0);
    }
    auto v = g.array(GS_GRID);
    printf("gridtest_dims %d %d\n", g.cols, g.rows);
    print("gridtest_grid0", v[0]);
    print("gridtest_grid1", v[1]);
    // slices address the same doubles (TwoDGrid.row / col / left :260-352)
    auto r1 = g.row(1), c1 = g.col(1), lf = g.left();
    int same = r1[1] == v[(size_t)g.cols + 1] && c1[2] == v[(size_t)2 * g.cols + 1] && lf[3] == v[(size_t)3 * g.cols];
    g.updateColumn(0, std::vector<double>((size_t)g.rows, 5.0));
    g.updateY(y, [](double yy) { return yy; });
    auto w = g.array(GS_GRID);
    same = same && w[(size_t)4 * g.cols + 7] == y.range[4];
    printf("gridtest_slices_ok %d\n", same);
  }
  host_only();
  device_sections(ctx);
  return 0;
}
