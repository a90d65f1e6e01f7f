// Runs config C1 (KAT R2) and the GridTest scenario (KAT R1, AT/GridTest.scala:17-66) through the C++ host mirror
// and prints bit patterns; tests/test_gpu_cpp_mirror.py compares them with the oracle's values.
#include <array>
#include <cinttypes>
#include <cstdio>
#include <cstring>

#include "../../gridsplines_b200/host/gridsplines.hpp"

using namespace gridsplines;

static void print(const char *name, double v) {
  uint64_t b;
  memcpy(&b, &v, 8);
  printf("%s %.17g 0x%016" PRIx64 "\n", name, v, b);
}

int main() {
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
  return 0;
}
