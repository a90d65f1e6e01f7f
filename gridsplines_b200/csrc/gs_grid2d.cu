// gs_grid2d.cu -- TwoDGrid (A/TwoDGrid.scala): the locally-one-dimensional implicit step
//   xIter (:135-193)  every row:    (I - tau*Lx) u* = u^n,      RHS from `grid`,   into `result`
//   yIter (:195-252)  every column: (I - tau*Ly) u^{n+1} = u*,  RHS from `result`, into `result`
//   Grid.update (:474-481)          predict = aVal(grid, result); grid <- result
// with coefficients frozen at `predict`, assembled by Dim.first/general/last and the heat-flow
// ends (A/Dim.scala:28-128).
//
// Device layout: the reference's own row-major arrays, i = row * cols + col (A/XDim.scala:30-32).
//   K4 (y-sweep): one thread per column; adjacent threads touch adjacent addresses at every row
//       step, so the column walk is coalesced as it stands.  The fused variant applies Grid.update
//       in its back-substitution (writes grid', predict' instead of result).
//   K3 (x-sweep): one thread per row.  A warp owns 32 consecutive rows and streams 32x32 tiles of
//       predict/grid through shared memory (cp.async double buffer, 256 B coalesced row segments,
//       padded 33-column tiles = conflict-free transposed reads); results leave through the same
//       transposed tiles.
// Both keep the reference's operation order exactly (GS_SOLVER_THOMAS, bit-exact).
#include <cuda.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>

#include "gs_device.cuh"
#include "gs_internal.h"
#include "gs_grid2d_k5.cuh"

int gs_launch_predef(gs_ctx *ctx, int dir, const double *values_dev, int n, double sigma, double *coefs_dev);
int gs_launch_heatflow(gs_ctx *ctx, int dir, const double *values_dev, int n, double *out4_dev);
int gs_launch_fill(gs_ctx *ctx, double *a, long long count, double v);

struct Maps2d {
  int mode[4];
  int single_off[4];
  const int *ids[4];
};

struct Sweep2dParams {
  int cols, rows;
  const double *predict;   // coefficients' temperatures (Grid.apply/getP, :470-472)
  const double *rhs;       // x: grid (Grid.get :464); y: result (Grid.res :466)
  double *out;             // result (x, unfused y)
  double *grid_io;         // fused y: old grid in, new grid out
  double *predict_out;     // fused y: new predict
  double2 *scratch;
  const double *coefs;     // Dim.coefs of the swept direction [2n]
  double firstVol, fC, lastVol, lA;
  const double *bound_first, *bound_last;  // dense, side length
  int kind_first, kind_last;
  const double *low_snap;  // fused y with a lower HeatFlow bound: predict[(rows-1)*cols - 1 ...] before the sweep
  Maps2d maps;
  const double *pool;
  int pool_len, smem_pool;
  double time;
  unsigned *status;
};

GS_DEV int map_off(const Maps2d &m, int sel, int col, int row, int cols) {
  switch (m.mode[sel]) {
    case GS_MAP_SINGLE: return m.single_off[sel];
    case GS_MAP_PER_ROW: return m.ids[sel][row + 1];
    case GS_MAP_PER_COL: return m.ids[sel][col + 1];
    default: return m.ids[sel][(size_t)(row + 1) * (size_t)(cols + 1) + (size_t)(col + 1)];
  }
}

// Coefficients.{apply, thermConductivity, capacity, tempConductivity}(col, row), indices from -1
template <bool SINGLE>
struct CoefSel {
  typedef GsSpline Z;
  static constexpr bool single = SINGLE;
  const Maps2d &m;
  const double *pool;
  int cols;
  GsSpline z[4];
  GS_DEV CoefSel(const Maps2d &maps, const double *pool_, int cols_) : m(maps), pool(pool_), cols(cols_) {
    if (SINGLE) {
#pragma unroll
      for (int s = 0; s < 4; ++s) z[s] = gs_spline_view(pool, m.single_off[s]);
    }
  }
  GS_DEV GsSpline get(int sel, int col, int row) const {
    if (SINGLE) return z[sel];
    return gs_spline_view(pool, map_off(m, sel, col, row, cols));
  }
};
// every selector is SINGLE and a literal Spline.const (ConstantCoef over Spline.const values)
struct CoefSelLit {
  typedef GsLitConst Z;
  static constexpr bool single = true;
  GsLitConst z[4];
  GS_DEV CoefSelLit(const Maps2d &maps, const double *pool, int) {
#pragma unroll
    for (int s = 0; s < 4; ++s) z[s].v = pool[maps.single_off[s] + GS_HDR + 3];
  }
  GS_DEV GsLitConst get(int sel, int, int) const { return z[sel]; }
};

// one line's forward elimination state
struct LineState {
  double delta, lambda, carry;
};

struct SweepConsts {
  GsRcp rt;
  bool fast_allowed;
  double inv_time;
  double firstVol, fC, lastVol, lA;
};

// Dim.first  A/Dim.scala:28-43
template <bool FAST>
GS_DEV void node_first_temp(LineState &st, const SweepConsts &k, double cf0, double cf1, double t1, double t2,
                            double t3, double t, const GsSpline &z0, const GsSpline &z, unsigned &flags, bool &ok) {
  double co0 = gs_take_average<FAST>(z0, t1, t2, flags, ok);
  double co = gs_take_average<FAST>(z, t2, t3, flags, ok);
  double a = cf0 * co0;
  double c = cf1 * co;
  double vect = gs_div<FAST>(-t, k.rt, ok) - a * t1;
  gs_forward_first<FAST>(-(a + c) - k.inv_time, c, vect, st.delta, st.lambda, ok);
  st.carry = co;
}
// Dim.firstHeatFlow :46-64
template <bool FAST>
GS_DEV void node_first_flow(LineState &st, const SweepConsts &k, double heatFlow, double t2, double t3, double t,
                            const GsSpline &conductivity, const GsSpline &capacity, unsigned &flags, bool &ok) {
  double cap = gs_ads_apply(capacity, t, flags);
  double cond = gs_ads_average<FAST>(conductivity, t2, t3, flags, ok);
  double vol = gs_div<FAST>(k.firstVol * cap, k.rt, ok);
  double c = -cond * k.fC;
  double b = vol - c;
  double vect = vol + heatFlow;
  double tCond = gs_div<FAST>(cond, gs_rcp<FAST>(cap, ok), ok);
  gs_forward_first<FAST>(b, c, vect, st.delta, st.lambda, ok);
  st.carry = tCond;
}
// Dim.general :67-83
template <bool FAST>
GS_DEV void node_general(LineState &st, const SweepConsts &k, double cf0, double cf1, double t2, double t3,
                         double t, const GsSpline &z, unsigned &flags, bool &ok) {
  double co = gs_take_average<FAST>(z, t2, t3, flags, ok);
  double a = cf0 * st.carry;
  double c = cf1 * co;
  double vect = gs_div<FAST>(-t, k.rt, ok);
  gs_forward_unit<FAST>(a, -(a + c) - k.inv_time, c, vect, st.delta, st.lambda, flags, ok);
  st.carry = co;
}
// Dim.last :86-103
template <bool FAST>
GS_DEV void node_last_temp(LineState &st, const SweepConsts &k, double cf0, double cf1, double t2, double t3,
                           double t, const GsSpline &z, unsigned &flags, bool &ok) {
  double co = gs_take_average<FAST>(z, t2, t3, flags, ok);
  double a = cf0 * st.carry;
  double c = cf1 * co;
  double vect = gs_div<FAST>(-t, k.rt, ok) - c * t3;
  gs_forward_last<FAST>(a, -(a + c) - k.inv_time, vect, st.delta, st.lambda, ok);
}
// Dim.lastHeatFlow :106-128
template <bool FAST>
GS_DEV void node_last_flow(LineState &st, const SweepConsts &k, double heatFlow, double t1, double t2, double t,
                           const GsSpline &conductivity, const GsSpline &capacity, unsigned &flags, bool &ok) {
  double cap = gs_ads_apply(capacity, t, flags);
  double cond = gs_ads_average<FAST>(conductivity, t1, t2, flags, ok);
  double vol = gs_div<FAST>(cap * k.lastVol, k.rt, ok);
  double a = -cond * k.lA;
  double b = vol - a;
  double vect = vol - heatFlow;
  gs_forward_last<FAST>(a, b, vect, st.delta, st.lambda, ok);
}

// Grid.aVal  A/TwoDGrid.scala:508-512
GS_DEV double aval2d(double oldv, double newv) {
  double delta = newv - oldv;
  if (fabs(delta) > 1.5) return newv;
  else return newv + delta / 2.0;
}

GS_DEV const double *stage_pool(const Sweep2dParams &p, double *sm) {
  const double *pool = p.pool;
  if (p.smem_pool) {
    for (int i = threadIdx.x; i < p.pool_len; i += blockDim.x) sm[i] = p.pool[i];
    pool = sm;
    __syncthreads();
  }
  return pool;
}

// ---------------------------------------------------------------------------------
// K4: y-sweep, one thread per column (TwoDGrid.yIter :195-252 [+ Grid.update :474-481])
// ---------------------------------------------------------------------------------
template <bool SINGLE, bool FUSE>
__global__ void __launch_bounds__(128) k_ysweep(Sweep2dParams p) {
  extern __shared__ double sm[];
  const double *pool = stage_pool(p, sm);
  const int col = blockIdx.x * blockDim.x + threadIdx.x;
  if (col >= p.cols) return;
  const int cols = p.cols, rows = p.rows;
  const double *__restrict__ P = p.predict + col;
  const double *__restrict__ R = p.rhs + col;
  double2 *__restrict__ sc = p.scratch + col;
  CoefSel<SINGLE> sel(p.maps, pool, cols);
  SweepConsts k;
  k.fast_allowed = true;
  k.rt = gs_rcp_invariant(p.time, k.fast_allowed);
  bool ok = true;
  k.inv_time = 1.0 / p.time;
  k.firstVol = p.firstVol; k.fC = p.fC; k.lastVol = p.lastVol; k.lA = p.lA;
  unsigned flags = 0;
  LineState st;

  int row = 0;
  double t2 = P[0], t3 = P[cols], t = R[0];
  if (p.kind_first == GS_TEMP) {
    double t1 = p.bound_first[col];
    node_first_temp<false>(st, k, p.coefs[0], p.coefs[1], t1, t2, t3, t, sel.get(GS_SEL_APPLY, col, -1),
                    sel.get(GS_SEL_APPLY, col, row), flags, ok);
  } else {
    node_first_flow<false>(st, k, p.bound_first[col], t2, t3, t, sel.get(GS_SEL_TEMPCOND, col, row),
                    sel.get(GS_SEL_CAP, col, row), flags, ok);
  }
  sc[0] = make_double2(st.delta, st.lambda);
  row = 1;
#pragma unroll 4
  for (; row < rows - 1; ++row) {
    size_t i = (size_t)row * cols;
    t2 = t3;
    t3 = P[i + cols];
    t = R[i];
    node_general<false>(st, k, p.coefs[2 * row], p.coefs[2 * row + 1], t2, t3, t, sel.get(GS_SEL_APPLY, col, row), flags, ok);
    sc[i] = make_double2(st.delta, st.lambda);
  }
  {
    size_t i = (size_t)row * cols;
    t2 = t3;   // predict of the last row; t3 keeps that value too ("stale" in the reference, :231-246)
    t = R[i];
    if (p.kind_last == GS_TEMP) {
      t3 = p.bound_last[col];
      node_last_temp<false>(st, k, p.coefs[2 * row], p.coefs[2 * row + 1], t2, t3, t, sel.get(GS_SEL_APPLY, col, row), flags, ok);
    } else {
      // `val t1 = grid(i - 1)`: the left neighbour in the same row (sic, :241)
      double t1 = FUSE ? p.low_snap[col] : p.predict[i + col - 1];
      node_last_flow<false>(st, k, p.bound_last[col], t1, t3, t, sel.get(GS_SEL_TEMPCOND, col, row),
                     sel.get(GS_SEL_CAP, col, row), flags, ok);
    }
  }
  // backwardColumn :324-342
  double u = 0.0;
  double dl = st.delta, ll = st.lambda;
  for (row = rows - 1; row >= 0; --row) {
    size_t i = (size_t)row * cols;
    if (row != rows - 1) {
      double2 v = sc[i];
      dl = v.x;
      ll = v.y;
    }
    u = dl * u + ll;
    if (FUSE) {
      double *G = p.grid_io + col;
      double g = G[i];
      p.predict_out[i + col] = aval2d(g, u);
      G[i] = u;
    } else {
      p.out[i + col] = u;
    }
  }
  if (!(fabs(u) <= 1.7976931348623157e308)) flags |= GS_FLAG_NONFINITE;
  gs_raise(p.status, flags);
}

// ---------------------------------------------------------------------------------
// K3: x-sweep, one thread per row, 32x32 transposed tiles (TwoDGrid.xIter :135-193)
// ---------------------------------------------------------------------------------
#define XT 32
#define XTP 33

GS_DEV void cp_async8(double *smem_dst, const double *gsrc) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(s), "l"(gsrc) : "memory");
}
GS_DEV void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
GS_DEV void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

template <bool SINGLE>
__global__ void __launch_bounds__(32) k_xsweep(Sweep2dParams p) {
  extern __shared__ double sm[];
  // [2 buffers][2 arrays][XT][XTP] tiles, then the spline pool
  double *tiles = sm;
  const int tile_elems = XT * XTP;
  const double *pool = p.pool;
  if (p.smem_pool) {
    double *ps = sm + 4 * tile_elems;
    for (int i = threadIdx.x; i < p.pool_len; i += blockDim.x) ps[i] = p.pool[i];
    pool = ps;
    __syncwarp();
  }
  const int lane = threadIdx.x;
  const int cols = p.cols, rows = p.rows;
  const int row0 = blockIdx.x * XT;
  const int row = row0 + lane;
  const bool active = row < rows;
  const int nrows = min(XT, rows - row0);
  const int ntiles = (cols + XT - 1) / XT;

  auto issue_tile = [&](int tile, int buf) {
    int c = tile * XT + lane;
    double *tp = tiles + (buf * 2 + 0) * tile_elems;
    double *tg = tiles + (buf * 2 + 1) * tile_elems;
    if (c < cols) {
      for (int r = 0; r < nrows; ++r) {
        size_t gi = (size_t)(row0 + r) * cols + c;
        cp_async8(tp + r * XTP + lane, p.predict + gi);
        cp_async8(tg + r * XTP + lane, p.rhs + gi);
      }
    }
    cp_async_commit();
  };

  CoefSel<SINGLE> sel(p.maps, pool, cols);
  SweepConsts k;
  k.fast_allowed = true;
  k.rt = gs_rcp_invariant(p.time, k.fast_allowed);
  bool ok = true;
  k.inv_time = 1.0 / p.time;
  k.firstVol = p.firstVol; k.fC = p.fC; k.lastVol = p.lastVol; k.lA = p.lA;
  unsigned flags = 0;
  LineState st;
  st.delta = st.lambda = st.carry = 0.0;
  double2 *__restrict__ sc = p.scratch + (active ? row : 0);

  // ---- forward: node `col` is assembled when predict[col + 1] has been read ----
  double pcur = 0.0, gcur = 0.0;
  issue_tile(0, 0);
  for (int tile = 0; tile < ntiles; ++tile) {
    int buf = tile & 1;
    if (tile + 1 < ntiles) {
      issue_tile(tile + 1, buf ^ 1);
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncwarp();
    const double *tp = tiles + (buf * 2 + 0) * tile_elems + lane * XTP;
    const double *tg = tiles + (buf * 2 + 1) * tile_elems + lane * XTP;
    int len = min(XT, cols - tile * XT);
    if (active) {
      for (int m = 0; m < len; ++m) {
        int cnext = tile * XT + m;   // column whose predict is read now
        double pnext = tp[m];
        if (cnext >= 1) {
          int col = cnext - 1;
          if (col == 0) {
            if (p.kind_first == GS_TEMP) {
              node_first_temp<false>(st, k, p.coefs[0], p.coefs[1], p.bound_first[row], pcur, pnext, gcur,
                              sel.get(GS_SEL_APPLY, -1, row), sel.get(GS_SEL_APPLY, 0, row), flags, ok);
            } else {
              node_first_flow<false>(st, k, p.bound_first[row], pcur, pnext, gcur, sel.get(GS_SEL_THERM, 0, row),
                              sel.get(GS_SEL_CAP, 0, row), flags, ok);
            }
          } else {
            node_general<false>(st, k, p.coefs[2 * col], p.coefs[2 * col + 1], pcur, pnext, gcur,
                         sel.get(GS_SEL_APPLY, col, row), flags, ok);
          }
          sc[(size_t)col * rows] = make_double2(st.delta, st.lambda);
        }
        pcur = pnext;
        gcur = tg[m];
      }
    }
    __syncwarp();
  }
  if (active) {
    int col = cols - 1;
    if (p.kind_last == GS_TEMP) {
      node_last_temp<false>(st, k, p.coefs[2 * col], p.coefs[2 * col + 1], pcur, p.bound_last[row], gcur,
                     sel.get(GS_SEL_APPLY, col, row), flags, ok);
    } else {
      // (t1, t2) = (predict[last], stale t3 = predict[last])  :172-187
      node_last_flow<false>(st, k, p.bound_last[row], pcur, pcur, gcur, sel.get(GS_SEL_TEMPCOND, col, row),
                     sel.get(GS_SEL_CAP, col, row), flags, ok);
    }
  }

  // ---- backward (backwardRow :303-322) through transposed output tiles ----
  double *to = tiles;  // [XT][XTP]
  double u = 0.0;
  for (int tile = ntiles - 1; tile >= 0; --tile) {
    int len = min(XT, cols - tile * XT);
    if (active) {
      for (int m = len - 1; m >= 0; --m) {
        int col = tile * XT + m;
        if (col != cols - 1) {
          double2 v = sc[(size_t)col * rows];
          st.delta = v.x;
          st.lambda = v.y;
        }
        u = st.delta * u + st.lambda;
        to[lane * XTP + m] = u;
      }
    }
    __syncwarp();
    int c = tile * XT + lane;
    if (c < cols) {
      for (int r = 0; r < nrows; ++r) p.out[(size_t)(row0 + r) * cols + c] = to[r * XTP + lane];
    }
    __syncwarp();
  }
  if (active && !(fabs(u) <= 1.7976931348623157e308)) flags |= GS_FLAG_NONFINITE;
  gs_raise(p.status, flags);
}

// ---------------------------------------------------------------------------------
// GS_SOLVER_PARTITIONED: a line is cut into segments of m nodes solved concurrently.
//
// The serial Thomas chain costs ~140 clk per node (tools/microbench), so a 4096^2 sweep with one thread per
// line is latency-bound at a few percent of the HBM roofline.  Here thread (line, k) owns nodes
// [s, e] = [k m, k m + m) and the interface values E_k = x_e are found from a reduced system:
//   A  forward sweep of the segment with the left neighbour value L = x_{s-1} as a parameter:
//        x_i = delta_i x_{i+1} + lambda_i + omega_i L      (same assembly and recurrences as the reference,
//        plus omega_i = -(sub_i omega_{i-1}) / Delta_i), and the composed map  x_s = p x_e + q + r L.
//        Nothing per node is stored: six doubles per segment leave the kernel.
//   B  one thread per line solves the S x S tridiagonal system for E_k (Thomas):
//        (1 - delta_e^k r^{k+1}) E_k - delta_e^k p^{k+1} E_{k+1} - omega_e^k E_{k-1} = lambda_e^k + delta_e^k q^{k+1}
//   C  every segment is now a Dirichlet problem (left value E_{k-1}, own last value E_k): the reference's
//        forward step with (delta, lambda) kept in shared memory, then the back-substitution
//        (fused with Grid.update in the y direction; predict' goes to a second buffer because neighbouring
//        segments still read predict).
// Operation order differs from the reference's single chain: results agree to rounding (<= 1e-12 relative per
// field, tests/test_gpu_2d.py), not bit for bit.  HBM traffic: inputs are read by A and again by C.
// Thread mapping: y sweep lanes <-> adjacent columns (coalesced rows); x sweep lanes <-> adjacent segments of
// one row (each lane streams its own contiguous 8 m bytes; a warp covers 32 m contiguous columns).
// ---------------------------------------------------------------------------------
struct Row4 {
  double sub, diag, sup, rhs;
};

struct PartParams {
  Sweep2dParams sw;
  int m, S;            // segment length, segments per line
  double *iface;       // [7][nlines * S]: delta_e, lambda_e, omega_e, p, q, r, E
  double *predict_alt; // fused y: predict' destination
};

// Dim.general :67-83 up to the elimination: (sub, diag, sup, rhs) of an interior node
template <bool FAST, class Z>
GS_DEV Row4 asm_general(const SweepConsts &k, double cf0, double cf1, double face_left, double p0, double p1,
                        double t, const Z &z, double &face_right, unsigned &flags, bool &ok) {
  double co = gs_take_average<FAST>(z, p0, p1, flags, ok);
  Row4 r;
  r.sub = cf0 * face_left;
  r.sup = cf1 * co;
  r.rhs = gs_div<FAST>(-t, k.rt, ok);
  r.diag = -(r.sub + r.sup) - k.inv_time;
  face_right = co;
  return r;
}
// Dim.first :28-43
template <bool FAST, class Z>
GS_DEV Row4 asm_first_temp(const SweepConsts &k, double cf0, double cf1, double t1, double p0, double p1, double t,
                           const Z &z0, const Z &z, double &face_right, unsigned &flags, bool &ok) {
  double co0 = gs_take_average<FAST>(z0, t1, p0, flags, ok);
  double co = gs_take_average<FAST>(z, p0, p1, flags, ok);
  double a = cf0 * co0;
  Row4 r;
  r.sub = 0.0;
  r.sup = cf1 * co;
  r.rhs = gs_div<FAST>(-t, k.rt, ok) - a * t1;
  r.diag = -(a + r.sup) - k.inv_time;
  face_right = co;
  return r;
}
// Dim.firstHeatFlow :46-64
template <bool FAST, class Z>
GS_DEV Row4 asm_first_flow(const SweepConsts &k, double heatFlow, double p0, double p1, double t,
                           const Z &conductivity, const Z &capacity, double &face_right,
                           unsigned &flags, bool &ok) {
  double cap = gs_ads_apply(capacity, t, flags);
  double cond = gs_ads_average<FAST>(conductivity, p0, p1, flags, ok);
  double vol = gs_div<FAST>(k.firstVol * cap, k.rt, ok);
  Row4 r;
  r.sub = 0.0;
  r.sup = -cond * k.fC;
  r.diag = vol - r.sup;
  r.rhs = vol + heatFlow;
  face_right = gs_div<FAST>(cond, gs_rcp<FAST>(cap, ok), ok);
  return r;
}
// Dim.last :86-103
template <bool FAST, class Z>
GS_DEV Row4 asm_last_temp(const SweepConsts &k, double cf0, double cf1, double face_left, double p0, double t3,
                          double t, const Z &z, unsigned &flags, bool &ok) {
  double co = gs_take_average<FAST>(z, p0, t3, flags, ok);
  Row4 r;
  r.sub = cf0 * face_left;
  double c = cf1 * co;
  r.sup = 0.0;
  r.rhs = gs_div<FAST>(-t, k.rt, ok) - c * t3;
  r.diag = -(r.sub + c) - k.inv_time;
  return r;
}
// Dim.lastHeatFlow :106-128
template <bool FAST, class Z>
GS_DEV Row4 asm_last_flow(const SweepConsts &k, double heatFlow, double t1, double t2, double t,
                          const Z &conductivity, const Z &capacity, unsigned &flags, bool &ok) {
  double cap = gs_ads_apply(capacity, t, flags);
  double cond = gs_ads_average<FAST>(conductivity, t1, t2, flags, ok);
  double vol = gs_div<FAST>(cap * k.lastVol, k.rt, ok);
  Row4 r;
  r.sub = -cond * k.lA;
  r.sup = 0.0;
  r.diag = vol - r.sub;
  r.rhs = vol - heatFlow;
  return r;
}

// one line of the swept direction as seen by a segment thread
template <bool XDIR>
struct LineAcc {
  const Sweep2dParams &p;
  int line;
  size_t base, stride;
  GS_DEV LineAcc(const Sweep2dParams &p_, int line_) : p(p_), line(line_) {
    base = XDIR ? (size_t)line_ * p_.cols : (size_t)line_;
    stride = XDIR ? 1 : (size_t)p_.cols;
  }
  GS_DEV int n() const { return XDIR ? p.cols : p.rows; }
  GS_DEV size_t idx(int i) const { return base + (size_t)i * stride; }
  GS_DEV double pred(int i) const { return p.predict[idx(i)]; }
  GS_DEV double rhs(int i) const { return p.rhs[idx(i)]; }
  GS_DEV int col(int i) const { return XDIR ? i : line; }
  GS_DEV int row(int i) const { return XDIR ? line : i; }
};

// assembles node i of the line (any position) from the reference's formulas; `face` is the carried face value
// (Dim's co0), `pm`, `p0` the predict values of nodes i-1 and i (pm unused at i = 0)
template <bool XDIR, class SEL, bool FAST>
GS_DEV Row4 asm_node(const LineAcc<XDIR> &A, const SEL &sel, const SweepConsts &k, int i, double p0,
                     double pnext, double t, double &face, unsigned &flags, bool &ok) {
  const Sweep2dParams &p = A.p;
  const int n = A.n();
  if (i == 0) {
    if (p.kind_first == GS_TEMP) {
      return asm_first_temp<FAST>(k, p.coefs[0], p.coefs[1], p.bound_first[A.line], p0, pnext, t,
                                  XDIR ? sel.get(GS_SEL_APPLY, -1, A.line) : sel.get(GS_SEL_APPLY, A.line, -1),
                                  sel.get(GS_SEL_APPLY, A.col(0), A.row(0)), face, flags, ok);
    }
    // left: thermConductivity (:154); upper: tempConductivity (:212)
    return asm_first_flow<FAST>(k, p.bound_first[A.line], p0, pnext, t,
                                sel.get(XDIR ? GS_SEL_THERM : GS_SEL_TEMPCOND, A.col(0), A.row(0)),
                                sel.get(GS_SEL_CAP, A.col(0), A.row(0)), face, flags, ok);
  }
  if (i == n - 1) {
    if (p.kind_last == GS_TEMP) {
      return asm_last_temp<FAST>(k, p.coefs[2 * i], p.coefs[2 * i + 1], face, p0, p.bound_last[A.line], t,
                                 sel.get(GS_SEL_APPLY, A.col(i), A.row(i)), flags, ok);
    }
    // right: (t1, t2) = (predict[last], predict[last]) (:172-187); lower: t1 = predict[flat - 1] (:241)
    double t1 = XDIR ? p0 : (p.low_snap ? p.low_snap[A.line] : p.predict[A.idx(i) - 1]);
    return asm_last_flow<FAST>(k, p.bound_last[A.line], t1, p0, t, sel.get(GS_SEL_TEMPCOND, A.col(i), A.row(i)),
                               sel.get(GS_SEL_CAP, A.col(i), A.row(i)), flags, ok);
  }
  return asm_general<FAST>(k, p.coefs[2 * i], p.coefs[2 * i + 1], face, p0, pnext, t,
                           sel.get(GS_SEL_APPLY, A.col(i), A.row(i)), face, flags, ok);
}

// face value between nodes s-1 and s as node s-1 would have carried it (Dim.general's `co`)
template <bool XDIR, class SEL, bool FAST>
GS_DEV double face_before(const LineAcc<XDIR> &A, const SEL &sel, const SweepConsts &k, int s, double pm,
                          double p0, unsigned &flags, bool &ok) {
  const Sweep2dParams &p = A.p;
  if (s - 1 == 0 && p.kind_first == GS_FLOW) {
    // the carry out of a heat-flow first node is cond / cap (Dim.firstHeatFlow :60)
    double t = A.rhs(0);
    double cap = gs_ads_apply(sel.get(GS_SEL_CAP, A.col(0), A.row(0)), t, flags);
    double cond = gs_ads_average<FAST>(sel.get(XDIR ? GS_SEL_THERM : GS_SEL_TEMPCOND, A.col(0), A.row(0)), pm, p0,
                                       flags, ok);
    return gs_div<FAST>(cond, gs_rcp<FAST>(cap, ok), ok);
  }
  return gs_take_average<FAST>(sel.get(GS_SEL_APPLY, A.col(s - 1), A.row(s - 1)), pm, p0, flags, ok);
}

// Inputs of a segment, fetched two nodes ahead of their use so that the global-load latency overlaps the
// (long, dependent) FP64 work of the node being eliminated.
template <bool XDIR>
struct SegStream {
  const LineAcc<XDIR> &A;
  int n;
  double pA, pB, tA, tB;   // predict[i+1], predict[i+2], rhs[i], rhs[i+1] for the current i
  GS_DEV SegStream(const LineAcc<XDIR> &A_, int s) : A(A_), n(A_.n()) {
    pA = (s + 1 < n) ? A.pred(s + 1) : 0.0;
    pB = (s + 2 < n) ? A.pred(s + 2) : 0.0;
    tA = A.rhs(s);
    tB = (s + 1 < n) ? A.rhs(s + 1) : 0.0;
  }
  // returns (pnext, t) of node i and issues the loads node i + 2 will need
  GS_DEV void next(int i, double &pnext, double &t) {
    pnext = pA;
    t = tA;
    pA = pB;
    tA = tB;
    pB = (i + 3 < n) ? A.pred(i + 3) : 0.0;
    tB = (i + 2 < n) ? A.rhs(i + 2) : 0.0;
  }
};

GS_DEV void part_ids(bool xdir, int nlines, int S, int &line, int &seg) {
  if (xdir) {
    seg = blockIdx.x * blockDim.x + threadIdx.x;
    line = blockIdx.y;
  } else {
    line = blockIdx.x * blockDim.x + threadIdx.x;
    seg = blockIdx.y;
  }
}
GS_DEV size_t iface_idx(bool xdir, int nlines, int S, int line, int seg) {
  return xdir ? (size_t)line * S + seg : (size_t)seg * nlines + line;
}

template <bool XDIR, class SEL, bool FAST>
GS_DEV bool part_A_body(const PartParams &q, const SEL &sel, const SweepConsts &k, int line, int seg,
                        double *out6, unsigned &flags) {
  const Sweep2dParams &p = q.sw;
  LineAcc<XDIR> A(p, line);
  const int n = A.n();
  const int s = seg * q.m, e = min(n, s + q.m) - 1;
  bool ok = true;
  double delta = 0.0, lambda = 0.0, omega = 0.0;
  double P = 1.0, Q = 0.0, R = 0.0;
  double p0 = A.pred(s), pnext = 0.0, t = 0.0, face = 0.0;
  SegStream<XDIR> in(A, s);
  if (s > 0) face = face_before<XDIR, SEL, FAST>(A, sel, k, s, A.pred(s - 1), p0, flags, ok);
  for (int i = s; i <= e; ++i) {
    in.next(i, pnext, t);
    Row4 r = asm_node<XDIR, SEL, FAST>(A, sel, k, i, p0, pnext, t, face, flags, ok);
    if (i > s) {  // the map x_s = P x_i + Q + R L gains node i-1's relation
      Q = Q + P * lambda;
      R = R + P * omega;
      P = P * delta;
    }
    // assertion of forwardUnit (A/passion/package.scala:279-282): interior nodes of the line only
    if (i > 0 && i < n - 1 && !(fabs(r.diag) >= fabs(r.sup) + fabs(r.sub))) flags |= GS_FLAG_NOT_DOMINANT;
    if (i == s) {
      GsRcp rd = gs_rcp<FAST>(r.diag, ok);
      delta = gs_div<FAST>(-r.sup, rd, ok);
      lambda = gs_div<FAST>(r.rhs, rd, ok);
      omega = (s > 0) ? gs_div<FAST>(-r.sub, rd, ok) : 0.0;
    } else {
      double bigDelta = r.diag + r.sub * delta;
      GsRcp rd = gs_rcp<FAST>(bigDelta, ok);
      delta = gs_div<FAST>(-r.sup, rd, ok);
      lambda = gs_div<FAST>(r.rhs - r.sub * lambda, rd, ok);
      omega = (s > 0) ? gs_div<FAST>(-(r.sub * omega), rd, ok) : 0.0;
    }
    p0 = pnext;
  }
  out6[0] = delta;   // 0 on the line's last node (sup = 0)
  out6[1] = lambda;
  out6[2] = omega;
  out6[3] = P;
  out6[4] = Q;
  out6[5] = R;
  return ok;
}

template <bool XDIR, class SEL>
__global__ void __launch_bounds__(128, 5) k_part_A(PartParams q) {
  extern __shared__ double sm[];
  const Sweep2dParams &p = q.sw;
  const double *pool = stage_pool(p, sm);
  const int nlines = XDIR ? p.rows : p.cols;
  int line, seg;
  part_ids(XDIR, nlines, q.S, line, seg);
  if (line >= nlines || seg >= q.S) return;
  SEL sel(p.maps, pool, p.cols);
  SweepConsts k;
  k.fast_allowed = true;
  k.rt = gs_rcp_invariant(p.time, k.fast_allowed);
  k.inv_time = 1.0 / p.time;
  k.firstVol = p.firstVol; k.fC = p.fC; k.lastVol = p.lastVol; k.lA = p.lA;
  unsigned flags = 0;
  double o[6];
  bool ok = k.fast_allowed && part_A_body<XDIR, SEL, true>(q, sel, k, line, seg, o, flags);
  if (!ok) {
    flags = 0;
    part_A_body<XDIR, SEL, false>(q, sel, k, line, seg, o, flags);
  }
  const size_t N = (size_t)nlines * q.S, id = iface_idx(XDIR, nlines, q.S, line, seg);
#pragma unroll
  for (int j = 0; j < 6; ++j) q.iface[j * N + id] = o[j];
  gs_raise(p.status, flags);
}

// B: reduced tridiagonal system per line (Thomas); E_k overwrites slot 6, slots 3/4 hold the scratch
template <bool XDIR>
__global__ void k_part_B(PartParams q) {
  const Sweep2dParams &p = q.sw;
  const int nlines = XDIR ? p.rows : p.cols;
  const int line = blockIdx.x * blockDim.x + threadIdx.x;
  if (line >= nlines) return;
  const int S = q.S;
  const size_t N = (size_t)nlines * S;
  double *F = q.iface;
  double cp = 0.0, dp = 0.0;
  for (int k = 0; k < S; ++k) {
    size_t id = iface_idx(XDIR, nlines, S, line, k);
    double de = F[0 * N + id], la = F[1 * N + id], om = F[2 * N + id];
    double pn = 0.0, qn = 0.0, rn = 0.0;
    if (k + 1 < S) {
      size_t idn = iface_idx(XDIR, nlines, S, line, k + 1);
      pn = F[3 * N + idn];
      qn = F[4 * N + idn];
      rn = F[5 * N + idn];
    }
    double sub = -om, diag = 1.0 - de * rn, sup = -(de * pn), rhs = la + de * qn;
    double den = diag - sub * cp;     // k = 0: sub = 0
    cp = sup / den;
    dp = (rhs - sub * dp) / den;
    F[3 * N + id] = cp;               // p, q of segment k are no longer needed (k+1's were just read)
    F[4 * N + id] = dp;
  }
  double x = 0.0;
  for (int k = S - 1; k >= 0; --k) {
    size_t id = iface_idx(XDIR, nlines, S, line, k);
    x = F[4 * N + id] - F[3 * N + id] * x;
    F[6 * N + id] = x;
  }
}

template <bool XDIR, class SEL, bool FAST, bool FUSE>
GS_DEV bool part_C_body(const PartParams &q, const SEL &sel, const SweepConsts &k, int line, int seg,
                        double2 *sc, int scs, unsigned &flags) {
  const Sweep2dParams &p = q.sw;
  LineAcc<XDIR> A(p, line);
  const int n = A.n();
  const int nlines = XDIR ? p.rows : p.cols;
  const int s = seg * q.m, e = min(n, s + q.m) - 1;
  const size_t N = (size_t)nlines * q.S;
  const double xe = q.iface[6 * N + iface_idx(XDIR, nlines, q.S, line, seg)];
  const double L = seg > 0 ? q.iface[6 * N + iface_idx(XDIR, nlines, q.S, line, seg - 1)] : 0.0;
  bool ok = true;
  double delta = 0.0, lambda = 0.0;
  double p0 = A.pred(s), pnext = 0.0, t = 0.0, face = 0.0;
  SegStream<XDIR> in(A, s);
  if (s > 0) face = face_before<XDIR, SEL, FAST>(A, sel, k, s, A.pred(s - 1), p0, flags, ok);
  for (int i = s; i < e; ++i) {   // node e's value is already known
    in.next(i, pnext, t);
    Row4 r = asm_node<XDIR, SEL, FAST>(A, sel, k, i, p0, pnext, t, face, flags, ok);
    if (i == s) {
      GsRcp rd = gs_rcp<FAST>(r.diag, ok);
      delta = gs_div<FAST>(-r.sup, rd, ok);
      lambda = gs_div<FAST>(s > 0 ? r.rhs - r.sub * L : r.rhs, rd, ok);
    } else {
      double bigDelta = r.diag + r.sub * delta;
      GsRcp rd = gs_rcp<FAST>(bigDelta, ok);
      delta = gs_div<FAST>(-r.sup, rd, ok);
      lambda = gs_div<FAST>(r.rhs - r.sub * lambda, rd, ok);
    }
    sc[(i - s) * scs] = make_double2(delta, lambda);
    p0 = pnext;
  }
  if (FAST && !ok) return false;
  double x = xe;
  double gA = 0.0, gB = 0.0;
  if (FUSE) {
    gA = p.grid_io[A.idx(e)];
    gB = (e - 1 >= s) ? p.grid_io[A.idx(e - 1)] : 0.0;
  }
  for (int i = e; i >= s; --i) {
    if (i < e) {
      double2 v = sc[(i - s) * scs];
      x = v.x * x + v.y;
    }
    size_t id = A.idx(i);
    if (FUSE) {
      double g = gA;
      gA = gB;
      gB = (i - 2 >= s) ? p.grid_io[A.idx(i - 2)] : 0.0;   // two nodes ahead (own nodes only: no aliasing with the stores below)
      q.predict_alt[id] = aval2d(g, x);
      p.grid_io[id] = x;
    } else {
      p.out[id] = x;
    }
  }
  if (!(fabs(x) <= 1.7976931348623157e308)) flags |= GS_FLAG_NONFINITE;
  return true;
}

template <bool XDIR, class SEL, bool FUSE>
__global__ void __launch_bounds__(128, 5) k_part_C(PartParams q) {
  extern __shared__ __align__(16) unsigned char smc[];
  const Sweep2dParams &p = q.sw;
  double2 *scratch = reinterpret_cast<double2 *>(smc);
  double *pool_sm = reinterpret_cast<double *>(smc + (size_t)q.m * blockDim.x * sizeof(double2));
  const double *pool = stage_pool(p, pool_sm);
  const int nlines = XDIR ? p.rows : p.cols;
  int line, seg;
  part_ids(XDIR, nlines, q.S, line, seg);
  if (line >= nlines || seg >= q.S) return;
  SEL sel(p.maps, pool, p.cols);
  SweepConsts k;
  k.fast_allowed = true;
  k.rt = gs_rcp_invariant(p.time, k.fast_allowed);
  k.inv_time = 1.0 / p.time;
  k.firstVol = p.firstVol; k.fC = p.fC; k.lastVol = p.lastVol; k.lA = p.lA;
  unsigned flags = 0;
  double2 *sc = scratch + threadIdx.x;
  bool done = k.fast_allowed && part_C_body<XDIR, SEL, true, FUSE>(q, sel, k, line, seg, sc, blockDim.x, flags);
  if (!done) {
    flags = 0;
    part_C_body<XDIR, SEL, false, FUSE>(q, sel, k, line, seg, sc, blockDim.x, flags);
  }
  gs_raise(p.status, flags);
}

#include "gs_grid2d_k6.cuh"
#include "gs_grid2d_k7.cuh"
#include "gs_grid2d_k5t.cuh"

// ---------------------------------------------------------------------------------
// small kernels: Grid.update, edges, reductions
// ---------------------------------------------------------------------------------
__global__ void k_update2d(double *grid, double *predict, const double *result, long long count) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < count; i += stride) {
    double r = result[i];
    predict[i] = aval2d(grid[i], r);
    grid[i] = r;
  }
}
// out[j] = a[start + j * stride]
__global__ void k_gather(const double *a, long long start, long long stride, int count, double *out) {
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j < count) out[j] = a[start + (long long)j * stride];
}
// sequential sum in index order (the reference's while loops, :353-411), one thread
__global__ void k_serial_mean(const double *v, int count, double *out) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  double sum = 0.0;
  for (int i = 0; i < count; ++i) sum += v[i];
  out[0] = sum / (double)count;
}
__global__ void k_broadcast(const double *src1, double *dst, int count) {
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j < count) dst[j] = src1[0];
}
// pairwise block sums (Grid.avValue is a diagnostic, not on the stepping path; its foldLeft order
// is not reproduced -- documented tolerance 1e-12 relative)
__global__ void k_block_sums(const double *a, long long count, double *partial) {
  __shared__ double sh[256];
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long stride = (long long)gridDim.x * blockDim.x;
  double s = 0.0;
  for (; i < count; i += stride) s += a[i];
  sh[threadIdx.x] = s;
  __syncthreads();
  for (int w = 128; w > 0; w >>= 1) {
    if ((int)threadIdx.x < w) sh[threadIdx.x] += sh[threadIdx.x + w];
    __syncthreads();
  }
  if (threadIdx.x == 0) partial[blockIdx.x] = sh[0];
}

// ---------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------
static int side_len(const gs_grid2d *g, int side) { return (side == GS_UPP || side == GS_LOW) ? g->cols : g->rows; }

static int dim_init(gs_ctx *ctx, gs_dim2d *d, int dir, int n, const double *values) {
  d->dir = dir;
  d->n = n;
  GS_CUDA(cudaMalloc(&d->values, ((size_t)n + 2) * sizeof(double)));
  GS_CUDA(cudaMalloc(&d->coefs, 2 * (size_t)n * sizeof(double)));
  GS_CUDA(cudaMemcpyAsync(d->values, values, ((size_t)n + 2) * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  // Dim.coefs = t.preDef(low, range, upp, 1.0)  A/Dim.scala:15 ; (firstVol, fC), (lastVol, lA) :16-17
  GS_TRY(gs_launch_predef(ctx, dir, d->values, n, 1.0, d->coefs));
  GS_TRY(gs_ctx_stage(ctx, 4 * sizeof(double)));
  GS_TRY(gs_launch_heatflow(ctx, dir, d->values, n, ctx->stage));
  GS_CUDA(cudaMemcpyAsync(d->hf, ctx->stage, 4 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  GS_CUDA(cudaStreamSynchronize(ctx->stream));
  return GS_OK;
}

extern "C" int gs_grid2d_create(gs_ctx *ctx, int cols, int rows, int xdir, int ydir, const double *xvalues,
                                const double *yvalues, gs_grid2d **out) {
  GS_REQUIRE(out != nullptr, "out is NULL");
  *out = nullptr;
  GS_TRY(gs_ctx_use(ctx));
  GS_REQUIRE(cols >= 2 && rows >= 2, "a grid needs at least 2 columns and 2 rows (Dim reads range(1), A/Dim.scala:16)");
  GS_REQUIRE(xvalues && yvalues, "NULL coordinates");
  GS_REQUIRE((xdir == GS_ORTHO || xdir == GS_RADIAL) && (ydir == GS_ORTHO || ydir == GS_RADIAL), "unknown direction kind");
  gs_grid2d *g = new gs_grid2d();
  g->ctx = ctx;
  gs_ctx_retain(ctx);
  g->cols = cols;
  g->rows = rows;
  int s = GS_OK;
  auto body = [&]() -> int {
    GS_TRY(dim_init(ctx, &g->x, xdir, cols, xvalues));
    GS_TRY(dim_init(ctx, &g->y, ydir, rows, yvalues));
    size_t bytes = (size_t)cols * rows * sizeof(double);
    GS_CUDA(cudaMalloc(&g->grid, bytes));
    GS_CUDA(cudaMalloc(&g->predict, bytes));
    GS_CUDA(cudaMalloc(&g->result, bytes));
    GS_CUDA(cudaMemsetAsync(g->grid, 0, bytes, ctx->stream));     // makeArray :453-459
    GS_CUDA(cudaMemsetAsync(g->predict, 0, bytes, ctx->stream));
    GS_CUDA(cudaMemsetAsync(g->result, 0, bytes, ctx->stream));
    GS_CUDA(cudaMalloc(&g->status, sizeof(unsigned)));
    GS_CUDA(cudaMemsetAsync(g->status, 0, sizeof(unsigned), ctx->stream));
    for (int side = 0; side < 4; ++side) {  // (One, Temp) = Temperature(Sparse(0.0))
      int len = side_len(g, side);
      GS_CUDA(cudaMalloc(&g->b[side].values, (size_t)len * sizeof(double)));
      GS_CUDA(cudaMemsetAsync(g->b[side].values, 0, (size_t)len * sizeof(double), ctx->stream));
    }
    for (int sel = 0; sel < 4; ++sel) {
      g->map_mode[sel] = GS_MAP_SINGLE;
      g->map_host[sel].assign(1, 0);
    }
    GS_CUDA(cudaStreamSynchronize(ctx->stream));
    return GS_OK;
  };
  s = body();
  if (s != GS_OK) {
    gs_grid2d_destroy(g);
    return s;
  }
  *out = g;
  return GS_OK;
}

extern "C" int gs_grid2d_destroy(gs_grid2d *g) {
  if (!g) return GS_OK;
  cudaSetDevice(g->ctx->device);
  cudaStreamSynchronize(g->ctx->stream);
  cudaFree(g->x.values); cudaFree(g->x.coefs);
  cudaFree(g->y.values); cudaFree(g->y.coefs);
  if (!g->bound[GS_GRID]) cudaFree(g->grid);
  if (!g->bound[GS_PREDICT]) cudaFree(g->predict);
  if (!g->bound[GS_RESULT]) cudaFree(g->result);
  for (int s = 0; s < 4; ++s) { cudaFree(g->b[s].values); cudaFree(g->map[s]); }
  cudaFree(g->scratch);
  cudaFree(g->iface);
  cudaFree(g->predict_alt);
  cudaFree(g->k5_ckpt);
  cudaFree(g->k6_F);
  cudaFree(g->k6_ends);
  cudaFree(g->k6_ckpt);
  cudaFree(g->k6_split);
  cudaFree(g->k7_scratch);
  for (gs_k5dir *kd : {&g->k5x, &g->k5y}) {
    cudaFree(kd->tab);
    cudaFree(kd->omega_end);
    cudaFree(kd->seg);
    cudaFree(kd->ends);
    cudaFree(kd->soft_buf);
    cudaFree(kd->soft_sync);
  }
  cudaFree(g->status);
  gs_ctx *owner = g->ctx;
  delete g;
  gs_ctx_release(owner);
  return GS_OK;
}

static size_t map_len(const gs_grid2d *g, int mode) {
  switch (mode) {
    case GS_MAP_SINGLE: return 1;
    case GS_MAP_PER_ROW: return (size_t)g->rows + 1;
    case GS_MAP_PER_COL: return (size_t)g->cols + 1;
    default: return ((size_t)g->rows + 1) * ((size_t)g->cols + 1);
  }
}

extern "C" int gs_grid2d_set_map(gs_grid2d *g, int selector, int mode, const int *ids) {
  GS_REQUIRE(g && ids, "NULL argument");
  GS_REQUIRE(selector >= 0 && selector < 4, "unknown selector");
  GS_REQUIRE(mode >= GS_MAP_SINGLE && mode <= GS_MAP_DENSE, "unknown map mode");
  gs_ctx *ctx = g->ctx;
  GS_TRY(gs_ctx_use(ctx));
  size_t n = map_len(g, mode);
  int nspl = (int)ctx->spline_off.size();
  std::vector<int> offs(n);
  for (size_t i = 0; i < n; ++i) {
    GS_REQUIRE(ids[i] >= 0 && ids[i] < nspl, "spline id out of range");
    offs[i] = ctx->spline_off[ids[i]];
  }
  GS_CUDA(cudaStreamSynchronize(ctx->stream));
  cudaFree(g->map[selector]);
  g->map[selector] = nullptr;
  GS_CUDA(cudaMalloc(&g->map[selector], n * sizeof(int)));
  GS_CUDA(cudaMemcpyAsync(g->map[selector], offs.data(), n * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
  GS_CUDA(cudaStreamSynchronize(ctx->stream));
  g->map_mode[selector] = mode;
  g->map_host[selector].assign(ids, ids + n);
  g->k5x.valid = g->k5y.valid = false;   // K5's line-independent tables were built from the previous coefficients
  return GS_OK;
}

extern "C" int gs_grid2d_set_bound(gs_grid2d *g, int side, int kind, int rep, const double *values, int n) {
  GS_REQUIRE(g && values, "NULL argument");
  GS_REQUIRE(side >= 0 && side < 4, "unknown side");
  GS_REQUIRE(kind == GS_TEMP || kind == GS_FLOW, "unknown bound kind");
  GS_REQUIRE(rep == GS_SPARSE || rep == GS_DENSE, "unknown bound representation");
  int len = side_len(g, side);
  GS_REQUIRE(n == 1 || n == len, "n must be 1 or the side length");
  gs_ctx *ctx = g->ctx;
  GS_TRY(gs_ctx_use(ctx));
  gs_bound2d *b = &g->b[side];
  b->kind = kind;
  b->rep = rep;
  if (n == 1) {
    // Dense.update(value) :632-638 / Sparse.update(el) :606-608
    GS_TRY(gs_launch_fill(ctx, b->values, len, values[0]));
  } else if (rep == GS_DENSE) {
    GS_CUDA(cudaMemcpyAsync(b->values, values, (size_t)len * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  } else {
    // Sparse.update(vals) :595-604: sequential mean, on the device
    GS_TRY(gs_ctx_stage(ctx, ((size_t)len + 1) * sizeof(double)));
    GS_CUDA(cudaMemcpyAsync(ctx->stage, values, (size_t)len * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    k_serial_mean<<<1, 32, 0, ctx->stream>>>(ctx->stage, len, ctx->stage + len);
    k_broadcast<<<(len + 127) / 128, 128, 0, ctx->stream>>>(ctx->stage + len, b->values, len);
    ctx->launches += 2;
    GS_CUDA(cudaGetLastError());
  }
  GS_CUDA(cudaStreamSynchronize(ctx->stream));
  return GS_OK;
}

extern "C" int gs_grid2d_get_bound(gs_grid2d *g, int side, double *out, int n) {
  GS_REQUIRE(g && out, "NULL argument");
  GS_REQUIRE(side >= 0 && side < 4, "unknown side");
  GS_REQUIRE(n >= 0 && n <= side_len(g, side), "n exceeds the side length");
  GS_TRY(gs_ctx_use(g->ctx));
  GS_CUDA(cudaMemcpyAsync(out, g->b[side].values, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, g->ctx->stream));
  GS_CUDA(cudaStreamSynchronize(g->ctx->stream));
  return GS_OK;
}

// `result` is only materialised when somebody looks at it after a fused iteration
static int materialise_result(gs_grid2d *g) {
  if (!g->result_is_grid) return GS_OK;
  GS_CUDA(cudaMemcpyAsync(g->result, g->grid, (size_t)g->cols * g->rows * sizeof(double), cudaMemcpyDeviceToDevice,
                          g->ctx->stream));
  g->result_is_grid = false;
  return GS_OK;
}

extern "C" int gs_grid2d_fill(gs_grid2d *g, double value) {
  GS_REQUIRE(g != nullptr, "NULL argument");
  GS_TRY(gs_ctx_use(g->ctx));
  GS_TRY(materialise_result(g));   // `result` keeps the last solve (the reference never touches it here)
  long long count = (long long)g->cols * g->rows;
  GS_TRY(gs_launch_fill(g->ctx, g->grid, count, value));
  GS_TRY(gs_launch_fill(g->ctx, g->predict, count, value));
  return GS_OK;
}

static double *array2d(gs_grid2d *g, int which) {
  return which == GS_GRID ? g->grid : which == GS_PREDICT ? g->predict : g->result;
}

extern "C" int gs_grid2d_upload(gs_grid2d *g, int which, const double *host) {
  GS_REQUIRE(g && host, "NULL argument");
  GS_REQUIRE(which >= GS_GRID && which <= GS_RESULT, "unknown array");
  GS_TRY(gs_ctx_use(g->ctx));
  if (which == GS_RESULT) g->result_is_grid = false;
  if (which == GS_GRID) GS_TRY(materialise_result(g));
  GS_CUDA(cudaMemcpyAsync(array2d(g, which), host, (size_t)g->cols * g->rows * sizeof(double), cudaMemcpyHostToDevice,
                          g->ctx->stream));
  GS_CUDA(cudaStreamSynchronize(g->ctx->stream));
  return GS_OK;
}
extern "C" int gs_grid2d_download(gs_grid2d *g, int which, double *host) {
  GS_REQUIRE(g && host, "NULL argument");
  GS_REQUIRE(which >= GS_GRID && which <= GS_RESULT, "unknown array");
  GS_TRY(gs_ctx_use(g->ctx));
  if (which == GS_RESULT) GS_TRY(materialise_result(g));
  GS_CUDA(cudaMemcpyAsync(host, array2d(g, which), (size_t)g->cols * g->rows * sizeof(double), cudaMemcpyDeviceToHost,
                          g->ctx->stream));
  GS_CUDA(cudaStreamSynchronize(g->ctx->stream));
  return GS_OK;
}

static int ensure_scratch(gs_grid2d *g) {
  size_t need = (size_t)g->cols * g->rows;
  if (g->scratch_elems >= need) return GS_OK;
  GS_CUDA(cudaMalloc(&g->scratch, need * sizeof(double2)));
  g->scratch_elems = need;
  return GS_OK;
}

static void fill_common(gs_grid2d *g, Sweep2dParams &p, double time, bool &single) {
  gs_ctx *ctx = g->ctx;
  p.cols = g->cols;
  p.rows = g->rows;
  p.predict = g->predict;
  p.scratch = g->scratch;
  single = true;
  for (int s = 0; s < 4; ++s) {
    p.maps.mode[s] = g->map_mode[s];
    p.maps.ids[s] = g->map[s];
    p.maps.single_off[s] = ctx->spline_off[g->map_host[s][0]];
    single = single && g->map_mode[s] == GS_MAP_SINGLE;
  }
  p.pool = ctx->pool_dev;
  p.pool_len = (int)ctx->pool.size();
  p.smem_pool = (size_t)p.pool_len * sizeof(double) <= 16 * 1024 ? 1 : 0;
  p.time = time;
  p.status = g->status;
  p.low_snap = nullptr;
  p.out = g->result;
  p.grid_io = nullptr;
  p.predict_out = nullptr;
}

static int launch_xsweep(gs_grid2d *g, double time) {
  gs_ctx *ctx = g->ctx;
  GS_REQUIRE(!ctx->spline_off.empty(), "no spline has been created in this context");
  GS_TRY(gs_ctx_sync_pool(ctx));
  GS_TRY(ensure_scratch(g));
  Sweep2dParams p;
  bool single;
  fill_common(g, p, time, single);
  p.rhs = g->grid;
  p.coefs = g->x.coefs;
  p.firstVol = g->x.hf[0]; p.fC = g->x.hf[1]; p.lastVol = g->x.hf[2]; p.lA = g->x.hf[3];
  p.bound_first = g->b[GS_LEFT].values;
  p.bound_last = g->b[GS_RIGHT].values;
  p.kind_first = g->b[GS_LEFT].kind;
  p.kind_last = g->b[GS_RIGHT].kind;
  size_t smem = 4 * (size_t)XT * XTP * sizeof(double) + (p.smem_pool ? (size_t)p.pool_len * sizeof(double) : 0);
  int blocks = (g->rows + XT - 1) / XT;
  if (smem > 48 * 1024) {
    GS_CUDA(cudaFuncSetAttribute(k_xsweep<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    GS_CUDA(cudaFuncSetAttribute(k_xsweep<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  if (single) k_xsweep<true><<<blocks, 32, smem, ctx->stream>>>(p);
  else k_xsweep<false><<<blocks, 32, smem, ctx->stream>>>(p);
  ctx->launches++;
  GS_CUDA(cudaGetLastError());
  return GS_OK;
}

static int launch_ysweep(gs_grid2d *g, double time, bool fuse) {
  gs_ctx *ctx = g->ctx;
  GS_REQUIRE(!ctx->spline_off.empty(), "no spline has been created in this context");
  GS_TRY(gs_ctx_sync_pool(ctx));
  GS_TRY(ensure_scratch(g));
  Sweep2dParams p;
  bool single;
  fill_common(g, p, time, single);
  p.rhs = g->result;
  p.coefs = g->y.coefs;
  p.firstVol = g->y.hf[0]; p.fC = g->y.hf[1]; p.lastVol = g->y.hf[2]; p.lA = g->y.hf[3];
  p.bound_first = g->b[GS_UPP].values;
  p.bound_last = g->b[GS_LOW].values;
  p.kind_first = g->b[GS_UPP].kind;
  p.kind_last = g->b[GS_LOW].kind;
  if (fuse) {
    p.grid_io = g->grid;
    p.predict_out = g->predict;
    if (p.kind_last == GS_FLOW) {
      // the lower HeatFlow end reads predict[i - 1] of the neighbouring column (:241); the fused
      // back-substitution overwrites predict, so that row is snapshotted first
      GS_TRY(gs_ctx_stage(ctx, (size_t)g->cols * sizeof(double)));
      long long start = (long long)(g->rows - 1) * g->cols - 1;
      k_gather<<<(g->cols + 127) / 128, 128, 0, ctx->stream>>>(g->predict, start, 1, g->cols, ctx->stage);
      ctx->launches++;
      p.low_snap = ctx->stage;
    }
  }
  int block = g->cols <= 32 * 2 * ctx->sm_count ? 32 : (g->cols <= 64 * 2 * ctx->sm_count ? 64 : 128);
  int blocks = (g->cols + block - 1) / block;
  size_t smem = p.smem_pool ? (size_t)p.pool_len * sizeof(double) : 0;
  if (single) {
    if (fuse) k_ysweep<true, true><<<blocks, block, smem, ctx->stream>>>(p);
    else k_ysweep<true, false><<<blocks, block, smem, ctx->stream>>>(p);
  } else {
    if (fuse) k_ysweep<false, true><<<blocks, block, smem, ctx->stream>>>(p);
    else k_ysweep<false, false><<<blocks, block, smem, ctx->stream>>>(p);
  }
  ctx->launches++;
  GS_CUDA(cudaGetLastError());
  return GS_OK;
}

// ---- partitioned sweeps -------------------------------------------------------------------
static int seg_len(gs_grid2d *g) { return (int)std::max<int64_t>(2, std::min<int64_t>(64, g->ctx->opt_seg_len)); }

// which solver a direction of n nodes uses
static bool use_partitioned(gs_grid2d *g, int n) {
  if (g->solver == GS_SOLVER_THOMAS) return false;
  int m = seg_len(g);
  if (g->solver == GS_SOLVER_PARTITIONED) return n >= 2 * m;
  return n >= 8 * m;  // AUTO: short lines stay on the bit-exact single chain
}

static int ensure_iface(gs_grid2d *g, size_t doubles) {
  if (g->iface_len >= doubles) return GS_OK;
  GS_CUDA(cudaStreamSynchronize(g->ctx->stream));
  cudaFree(g->iface);
  g->iface = nullptr;
  g->iface_len = 0;
  GS_CUDA(cudaMalloc(&g->iface, doubles * sizeof(double)));
  g->iface_len = doubles;
  return GS_OK;
}

static int launch_part(gs_grid2d *g, double time, bool xdir, bool fuse) {
  gs_ctx *ctx = g->ctx;
  GS_REQUIRE(!ctx->spline_off.empty(), "no spline has been created in this context");
  GS_TRY(gs_ctx_sync_pool(ctx));
  PartParams q;
  bool single;
  fill_common(g, q.sw, time, single);
  Sweep2dParams &p = q.sw;
  const gs_dim2d &d = xdir ? g->x : g->y;
  const int n = xdir ? g->cols : g->rows, nlines = xdir ? g->rows : g->cols;
  p.rhs = xdir ? g->grid : g->result;
  p.coefs = d.coefs;
  p.firstVol = d.hf[0]; p.fC = d.hf[1]; p.lastVol = d.hf[2]; p.lA = d.hf[3];
  const int sf = xdir ? GS_LEFT : GS_UPP, sl = xdir ? GS_RIGHT : GS_LOW;
  p.bound_first = g->b[sf].values;
  p.bound_last = g->b[sl].values;
  p.kind_first = g->b[sf].kind;
  p.kind_last = g->b[sl].kind;
  q.m = seg_len(g);
  q.S = (n + q.m - 1) / q.m;
  GS_REQUIRE(nlines <= 65535 && q.S <= 65535, "grid too large for the partitioned launch shape");
  GS_TRY(ensure_iface(g, 7 * (size_t)nlines * q.S));
  q.iface = g->iface;
  q.predict_alt = nullptr;
  if (fuse) {
    if (!g->predict_alt) GS_CUDA(cudaMalloc(&g->predict_alt, (size_t)g->cols * g->rows * sizeof(double)));
    p.grid_io = g->grid;
    q.predict_alt = g->predict_alt;
  }
  const int bx = xdir ? std::min(128, (q.S + 31) / 32 * 32) : 128;
  dim3 grid = xdir ? dim3((q.S + bx - 1) / bx, nlines) : dim3((nlines + bx - 1) / bx, q.S);
  size_t pool_smem = p.smem_pool ? (size_t)p.pool_len * sizeof(double) : 0;
  size_t c_smem = (size_t)q.m * bx * sizeof(double2) + pool_smem;
#define GS_PART(XD, SG)                                                                         \
  do {                                                                                          \
    k_part_A<XD, SG><<<grid, bx, pool_smem, ctx->stream>>>(q);                                  \
    k_part_B<XD><<<(nlines + 63) / 64, 64, 0, ctx->stream>>>(q);                                \
    if (fuse) {                                                                                 \
      GS_CUDA(cudaFuncSetAttribute(k_part_C<XD, SG, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c_smem)); \
      k_part_C<XD, SG, true><<<grid, bx, c_smem, ctx->stream>>>(q);                             \
    } else {                                                                                    \
      GS_CUDA(cudaFuncSetAttribute(k_part_C<XD, SG, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c_smem)); \
      k_part_C<XD, SG, false><<<grid, bx, c_smem, ctx->stream>>>(q);                            \
    }                                                                                           \
  } while (0)
  bool lit = single;
  for (int sl4 = 0; sl4 < 4 && lit; ++sl4) {
    const double *h = ctx->pool.data() + p.maps.single_off[sl4];
    lit = (int)h[0] == GS_CONST && (int)h[1] == 1 && h[2] == -1.7976931348623157e308 && h[3] == h[2] &&
          h[4] == h[GS_HDR + 3] && h[5] == h[GS_HDR + 3] && h[GS_HDR] == h[2] && h[GS_HDR + 1] == -h[2];
  }
  if (xdir) {
    if (lit) GS_PART(true, CoefSelLit); else if (single) GS_PART(true, CoefSel<true>); else GS_PART(true, CoefSel<false>);
  } else {
    if (lit) GS_PART(false, CoefSelLit); else if (single) GS_PART(false, CoefSel<true>); else GS_PART(false, CoefSel<false>);
  }
#undef GS_PART
  ctx->launches += 3;
  GS_CUDA(cudaGetLastError());
  if (fuse) {
    if (g->bound[GS_PREDICT]) {   // caller-owned predict: copy back instead of swapping pointers
      GS_CUDA(cudaMemcpyAsync(g->predict, g->predict_alt, (size_t)g->cols * g->rows * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
    } else {
      std::swap(g->predict, g->predict_alt);  // predict' was written to the second buffer
    }
  }
  return GS_OK;
}

static bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
// tensor map of a row-major rows x cols array of doubles with K7's box: x sweep 16 columns x 32 rows, 128-byte swizzle
// (the transposing tile); y sweep 32 columns x 16 rows.  cuTensorMapEncodeTiled is a driver entry point: fetched through
// the runtime, libcuda is not linked.
typedef CUresult (*gs_tmap_encode_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                      const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                      CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static_assert(sizeof(CUtensorMap) == sizeof(K7Map), "CUtensorMap is a 128-byte blob");
// generic form: a row-major dim1 x dim0 array of doubles (dim0 innermost), box0 x box1 boxes
static int gs_encode_map2d(K7Map *out, const void *base, uint64_t dim0, uint64_t dim1, unsigned box0, unsigned box1,
                           bool swizzle128, int promo256 = 0) {
  static gs_tmap_encode_fn encode = nullptr;
  if (!encode) {
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    GS_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    if (qres != cudaDriverEntryPointSuccess || !fn) {
      gs_set_error("cuTensorMapEncodeTiled is not available from this driver");
      return GS_ERR_CUDA;
    }
    encode = reinterpret_cast<gs_tmap_encode_fn>(fn);
  }
  const cuuint64_t dims[2] = {(cuuint64_t)dim0, (cuuint64_t)dim1};
  const cuuint64_t strides[1] = {(cuuint64_t)dim0 * sizeof(double)};
  const cuuint32_t box[2] = {box0, box1};
  const cuuint32_t estr[2] = {1u, 1u};
  CUresult rc = encode(reinterpret_cast<CUtensorMap *>(out), CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<void *>(base),
                       dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                       swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                       promo256 ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B : CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (rc != CUDA_SUCCESS) {
    gs_set_error("cuTensorMapEncodeTiled failed with CUresult %d (%llu x %llu, box %u x %u)", (int)rc,
                 (unsigned long long)dim1, (unsigned long long)dim0, box1, box0);
    return GS_ERR_CUDA;
  }
  return GS_OK;
}
// ---- K5: hoisted sweeps for literal-constant coefficients --------------------------------------------------
static bool literal_const(gs_ctx *ctx, int off, double *value) {
  const double *h = ctx->pool.data() + off;
  bool lit = (int)h[0] == GS_CONST && (int)h[1] == 1 && h[2] == -1.7976931348623157e308 && h[3] == h[2] &&
             h[4] == h[GS_HDR + 3] && h[5] == h[GS_HDR + 3] && h[GS_HDR] == h[2] && h[GS_HDR + 1] == -h[2];
  if (lit && value) *value = h[GS_HDR + 3];
  return lit;
}

static void k5_shape(int n, int *m, int *S, int smax = 16) {
  int s = std::max(2, std::min(smax, n / 256));
  int mm = ((n + s - 1) / s + K5_CH - 1) / K5_CH * K5_CH;
  *m = mm;
  *S = (n + mm - 1) / mm;
}
static size_t k5_smem(bool xdir, int m, int S, int cs = 1, int nb = 2) {
  size_t tile = (size_t)K5_CH * (xdir ? 33 : 32);
  (void)m;
  // clustered strips: the first CTA gathers lambda_e and Q of all cs * S segments (every CTA of the launch is sized alike)
  size_t gather = cs > 1 ? (size_t)2 * S * cs * 32 : 0;
  return ((size_t)2 * S * 32 + (size_t)S * nb * (tile + 4 * K5_CH) + gather) * sizeof(double);
}
// tile buffers per warp: two.  Three (option k5_nb = 3; lane <-> line sweeps without clusters, where they fit) were
// MEASURED slower on B200: C3 4096^2 0.236 ms against 0.220 ms, C5 2.01 ms against 1.78 ms -- a second chunk in flight
// per warp is not what these sweeps lack.
static int k5_nb(gs_ctx *ctx, bool xdir, int m, int S, int cs) {
  if (ctx->opt_k5_nb != 3 || xdir || cs > 1) return 2;
  return k5_smem(xdir, m, S, cs, 3) <= ctx->smem_optin ? 3 : 2;
}
// With fewer strips than half the SMs a strip is given to a cluster of 2 / 4 / 8 CTAs (S segments each)
static int k5_cluster(gs_ctx *ctx, void (*kern)(K5Params), bool xdir, int n, int nlines, int *m, int *S) {
  k5_shape(n, m, S);
  if (!ctx->opt_k5_cluster) return 1;
  const int strips = (nlines + 31) / 32;
  int cs = 1;
  while (cs < 8 && strips * cs * 2 <= ctx->sm_count) cs *= 2;
  for (; cs > 1; cs /= 2) {
    const int st = *S * cs;
    const int mm = ((n + st - 1) / st + K5_CH - 1) / K5_CH * K5_CH;
    const size_t smem = k5_smem(xdir, mm, *S, cs);
    if (!(mm >= 2 * K5_CH && (n + mm - 1) / mm == st && smem <= ctx->smem_optin)) continue;
    // every cluster must be resident at once (a cluster needs cs free SMs in ONE GPC: with one CTA per SM, 8-CTA
    // clusters leave SMs of every GPC unused and the strips would run in two waves)
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) continue;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(strips * cs);
    cfg.blockDim = dim3(32 * *S);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cs; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int resident = 0;
    if (cudaOccupancyMaxActiveClusters(&resident, kern, &cfg) != cudaSuccess) { cudaGetLastError(); continue; }
    if (resident < strips) continue;
    *m = mm;
    return cs;
  }
  return 1;
}

// K5T (gs_grid2d_k5t.cuh): the tuning options a shape was chosen under, and the shape itself.
// kind: 0 = x sweep, 1 = y sweep (2-D), 2 = 1-D batch.  Returns false where K5T does not apply (k5_sweep then runs).
static long long k5_opt_sig(gs_ctx *ctx) {
  return ((long long)(ctx->opt_k5t & 3) << 40) | ((long long)(ctx->opt_k5t_sx & 255) << 1) | ((long long)(ctx->opt_k5t_sy & 255) << 9) |
         ((long long)(ctx->opt_k5t_s1 & 255) << 17) | ((long long)(ctx->opt_k5t_nt & 255) << 25) |
         ((long long)ctx->opt_k5_cluster << 33) | ((long long)ctx->opt_k5t_wide << 34) | ((long long)ctx->opt_k5t_soft << 36);
}
// wide x stages: 0 never, 1 for lines of 8192 nodes and more (measured: 16384^2 x sweep 1.120 against 1.153 ms; 4096^2
// 71.9 against 70.6 us), 2 always
static bool k5t_wide(gs_ctx *ctx, int kind, int n) {
  return kind == 0 && (ctx->opt_k5t_wide >= 2 || (ctx->opt_k5t_wide == 1 && n >= 8192));
}
typedef void (*k5t_kern_t)(K5Params, int, int, const K5TMaps);
static k5t_kern_t k5t_kernel(int kind, int S, bool wide) {
  if (kind == 0 && wide) return S <= 8 ? k5t_sweep<true, false, 256, true> : k5t_sweep<true, false, 512, true>;
  if (kind == 0) return S <= 8 ? k5t_sweep<true, false, 256, false> : k5t_sweep<true, false, 512, false>;
  if (kind == 1) return S <= 8 ? k5t_sweep<false, false, 256, false> : k5t_sweep<false, false, 512, false>;
  return S <= 8 ? k5t_sweep<false, true, 256, false> : k5t_sweep<false, true, 512, false>;
}
static int k5t_configure(gs_ctx *ctx, k5t_kern_t kern) {
  // function attributes are per device: remembered per context
  const void *key = reinterpret_cast<const void *>(kern);
  if (std::find(ctx->configured.begin(), ctx->configured.end(), key) == ctx->configured.end()) {
    GS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ctx->smem_optin));
    GS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    ctx->configured.push_back(key);
  }
  return GS_OK;
}
// The K5T shape of a sweep over `nlines` lines of n nodes: segments per CTA S (= warps), segment length m, tiles per
// warp ring NT, CTAs per strip cs (a thread-block cluster when the strips alone would leave half the SMs idle; every
// cluster must be resident at once).  Returns false where K5T does not apply (k5_sweep then runs).
static bool k5t_plan(gs_ctx *ctx, int kind, int n, int nlines, int *m, int *S, int *NT, int *cs_out, bool *soft_out) {
  if (!ctx->opt_k5t) return false;
  // small problems are latency-bound and k5_sweep's cp.async tiles start faster than tensor-map boxes (MEASURED ms per
  // step K5T / K5: 256^2 0.0285 / 0.0247, 512^2 0.0298 / 0.0268, 1024^2 0.0343 / 0.0360, 2048^2 0.0556 / 0.0699)
  if (ctx->opt_k5t < 2 && (long long)n * nlines < (1ll << 20)) return false;
  const bool xdir = kind == 0, wide = k5t_wide(ctx, kind, n);
  int smax = kind == 0 ? ctx->opt_k5t_sx : (kind == 1 ? ctx->opt_k5t_sy : ctx->opt_k5t_s1);
  if (smax <= 0) smax = (xdir && !wide) ? 16 : 8;
  smax = std::min(16, std::max(2, smax));
  const int need = kind == 1 ? 2 : 1;   // the y sweep's stages are pairs of tiles
  const int cap = ctx->opt_k5t_nt > 0 ? std::min(8, ctx->opt_k5t_nt) : 8;
  const int strips = (nlines + 31) / 32;
  int mm1, ss;
  k5_shape(n, &mm1, &ss, smax);
  if (ss < 2) return false;
  // segment length when a strip's line is split over cs CTAs of ss segments each (0: does not divide)
  auto seg_len = [&](int cs) {
    if (cs == 1) return mm1;
    const int st = ss * cs;
    const int mm = ((n + st - 1) / st + K5_CH - 1) / K5_CH * K5_CH;
    // (segments shorter than 4 chunks are all pipeline ramp: 4096 columns x 512 rows, x sweep: m = 32 measured 84 us, m = 64 29 us)
    return (mm >= 4 * K5_CH && (n + mm - 1) / mm == st && st <= 128) ? mm : 0;
  };
  auto tiles_for = [&](int cs_smem) {
    int nt = 0;
    for (int t = cap; t >= need; --t)
      if (k5t_smem(xdir, wide, ss, t, cs_smem) <= ctx->smem_optin) { nt = t; break; }
    if (kind == 1) nt &= ~1;
    return nt;
  };
  k5t_kern_t kern = k5t_kernel(kind, ss, wide);
  // hardware clusters: any size up to 8 (not only powers of two: GPCs of 16 - 20 SMs hold two clusters of 6 where only
  // one of 8 fits), the largest whose clusters are all resident at once
  int cs = 1;
  if (ctx->opt_k5_cluster && strips * 2 <= ctx->sm_count) cs = std::min(8, ctx->sm_count / strips);
  int hw_cs = 0, hw_nt = 0, hw_m = 0;
  for (; cs >= 1; --cs) {
    const int mm = seg_len(cs), nt = tiles_for(cs);
    if (!mm || nt < need) continue;
    if (cs > 1) {
      if (k5t_configure(ctx, kern) != GS_OK) { cudaGetLastError(); continue; }
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(strips * cs);
      cfg.blockDim = dim3(32 * ss);
      cfg.dynamicSmemBytes = k5t_smem(xdir, wide, ss, nt, cs);
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeClusterDimension;
      attr[0].val.clusterDim.x = cs; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
      cfg.attrs = attr;
      cfg.numAttrs = 1;
      int resident = 0;
      if (cudaOccupancyMaxActiveClusters(&resident, kern, &cfg) != cudaSuccess) { cudaGetLastError(); continue; }
      if (resident < strips) continue;
    }
    hw_cs = cs; hw_nt = nt; hw_m = mm;
    break;
  }
  // soft groups (global-memory gather, no GPC constraint): taken when they put at least 10 % more CTAs to work; every
  // CTA must be resident (they wait for each other)
  int sf_cs = 0, sf_nt = 0, sf_m = 0;
  if (ctx->opt_k5_cluster && ctx->opt_k5t_soft && strips * 2 <= ctx->sm_count) {
    const int nt = tiles_for(1);
    int per_sm = 0;
    if (nt >= need && k5t_configure(ctx, kern) == GS_OK &&
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 32 * ss, k5t_smem(xdir, wide, ss, nt, 1)) == cudaSuccess) {
      for (int c2 = std::min(16, per_sm * ctx->sm_count / strips); c2 >= 2; --c2) {
        const int mm = seg_len(c2);
        // the first CTA stages [2][c2 * ss][32] doubles in its output tiles (idle until phase C stores)
        if (!mm || (size_t)2 * c2 * ss * 256 > (size_t)ss * (xdir ? 1 : 2) * K5T_TILE) continue;
        sf_cs = c2; sf_nt = nt; sf_m = mm;
        break;
      }
    } else {
      cudaGetLastError();
    }
  }
  const bool soft = sf_cs > 0 && (hw_cs == 0 || sf_cs * 10 >= hw_cs * 11);
  if (!soft && hw_cs == 0) return false;
  *m = soft ? sf_m : hw_m; *S = ss; *NT = soft ? sf_nt : hw_nt; *cs_out = soft ? sf_cs : hw_cs; *soft_out = soft;
  if (getenv("GS_DEBUG"))
    fprintf(stderr, "[gs] K5T plan kind=%d n=%d lines=%d: S=%d m=%d NT=%d CTAs per strip=%d (%s) wide=%d\n", kind, n, nlines, ss,
            *m, *NT, *cs_out, soft ? "soft group" : "cluster", (int)wide);
  return true;
}
// buffers of a soft group: [3][strips][cs * S][32] doubles + [strips][2] words (zeroed once: the counters are monotonic)
static int k5t_soft_buffers(gs_ctx *ctx, gs_k5dir &kd, int strips) {
  const size_t need = (size_t)3 * strips * kd.cs * kd.S * 32;
  if (kd.soft_len < need || kd.soft_strips < (size_t)strips) {
    GS_CUDA(cudaStreamSynchronize(ctx->stream));
    cudaFree(kd.soft_buf);
    cudaFree(kd.soft_sync);
    kd.soft_buf = nullptr; kd.soft_sync = nullptr; kd.soft_len = 0; kd.soft_strips = 0;
    GS_CUDA(cudaMalloc(&kd.soft_buf, need * sizeof(double)));
    GS_CUDA(cudaMalloc(&kd.soft_sync, (size_t)2 * strips * sizeof(unsigned)));
    GS_CUDA(cudaMemsetAsync(kd.soft_sync, 0, (size_t)2 * strips * sizeof(unsigned), ctx->stream));
    kd.soft_len = need; kd.soft_strips = strips;
    kd.epoch = 0;
  }
  return GS_OK;
}
static int k5t_launch(gs_ctx *ctx, int kind, bool wide, int blocks, int S, int NT, int cs, const K5Params &q0, const K5TMaps &tm) {
  K5Params q = q0;
  const int hw = q.soft ? 1 : cs;   // soft groups: plain CTAs
  q.pdl = ctx->opt_k5t_pdl ? 1 : 0;
  k5t_kern_t kern = k5t_kernel(kind, S, wide);
  GS_TRY(k5t_configure(ctx, kern));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(blocks * cs);
  cfg.blockDim = dim3(32 * S);
  cfg.dynamicSmemBytes = k5t_smem(kind == 0, wide, S, NT, hw);
  cfg.stream = ctx->stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = hw; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  // programmatic dependent launch: the CTAs' set-up overlaps the tail of the previous kernel of the stream
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = q.pdl ? 2 : 1;
  GS_CUDA(cudaLaunchKernelEx(&cfg, kern, q, NT, (int)ctx->opt_k5t_hints, tm));
  return GS_OK;
}

static bool use_k5(gs_grid2d *g, bool xdir) {
  gs_ctx *ctx = g->ctx;
  if (g->solver == GS_SOLVER_THOMAS || !ctx->opt_k5) return false;
  const int n = xdir ? g->cols : g->rows;
  if (n < (g->solver == GS_SOLVER_PARTITIONED ? 128 : 256)) return false;  // AUTO keeps short lines bit-exact
  for (int s = 0; s < 4; ++s) {
    if (g->map_mode[s] != GS_MAP_SINGLE) return false;
    if (!literal_const(ctx, ctx->spline_off[g->map_host[s][0]], nullptr)) return false;
  }
  int m, S;
  k5_shape(n, &m, &S);
  return S >= 2 && k5_smem(xdir, m, S) <= ctx->smem_optin;
}

static int k5_prepare(gs_grid2d *g, bool xdir, double time) {
  gs_ctx *ctx = g->ctx;
  gs_k5dir &kd = xdir ? g->k5x : g->k5y;
  const gs_dim2d &d = xdir ? g->x : g->y;
  const int n = d.n;
  const int sf = xdir ? GS_LEFT : GS_UPP, sl = xdir ? GS_RIGHT : GS_LOW;
  // (caller-owned arrays bound with gs_grid2d_bind_device may be less than 16-byte aligned: no tensor maps then)
  const bool al = aligned16(g->grid) && aligned16(g->predict) && aligned16(g->result);
  if (kd.valid && kd.time == time && kd.kind_first == g->b[sf].kind && kd.kind_last == g->b[sl].kind &&
      kd.pool_gen == ctx->spline_off.size() && kd.opt_sig == k5_opt_sig(ctx) && kd.aligned == al)
    return GS_OK;
  kd.aligned = al;
  kd.cs = k5_cluster(ctx, xdir ? k5_sweep<true, false, 2> : k5_sweep<false, false, 2>, xdir, n, xdir ? g->rows : g->cols, &kd.m, &kd.S);
  // K5T: tensor-map strides are multiples of 16 bytes (an even number of columns)
  kd.soft = false;
  kd.tma = al && !(g->cols & 1) && k5t_plan(ctx, xdir ? 0 : 1, n, xdir ? g->rows : g->cols, &kd.m, &kd.S, &kd.NT, &kd.cs, &kd.soft);
  kd.opt_sig = k5_opt_sig(ctx);
  if (!kd.tab) {
    GS_CUDA(cudaMalloc(&kd.tab, (size_t)n * sizeof(double4)));
    GS_CUDA(cudaMalloc(&kd.omega_end, ((size_t)n / K5_CH + 2) * sizeof(double)));
    GS_CUDA(cudaMalloc(&kd.seg, 128 * sizeof(double4)));
    GS_CUDA(cudaMalloc(&kd.ends, 4 * sizeof(double)));
  }
  K5Setup q;
  q.n = n; q.m = kd.m; q.S = kd.S * kd.cs; q.dir_x = xdir ? 1 : 0;
  q.coefs = d.coefs;
  q.firstVol = d.hf[0]; q.fC = d.hf[1]; q.lastVol = d.hf[2]; q.lA = d.hf[3];
  q.kind_first = g->b[sf].kind; q.kind_last = g->b[sl].kind;
  double va, vth, vc, vt;
  literal_const(ctx, ctx->spline_off[g->map_host[GS_SEL_APPLY][0]], &va);
  literal_const(ctx, ctx->spline_off[g->map_host[GS_SEL_THERM][0]], &vth);
  literal_const(ctx, ctx->spline_off[g->map_host[GS_SEL_CAP][0]], &vc);
  literal_const(ctx, ctx->spline_off[g->map_host[GS_SEL_TEMPCOND][0]], &vt);
  q.v_apply = va; q.v_first_cond = xdir ? vth : vt; q.v_cap = vc; q.v_tempcond = vt;
  q.time = time;
  q.tab = kd.tab; q.omega_end = kd.omega_end; q.seg = kd.seg; q.ends = kd.ends; q.status = g->status;
  k5_setup<<<1, 32, 0, ctx->stream>>>(q);
  ctx->launches++;
  GS_CUDA(cudaGetLastError());
  GS_CUDA(cudaMemcpyAsync(kd.ends_host, kd.ends, 4 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  GS_CUDA(cudaStreamSynchronize(ctx->stream));
  kd.valid = true;
  kd.time = time;
  kd.kind_first = q.kind_first;
  kd.kind_last = q.kind_last;
  kd.pool_gen = ctx->spline_off.size();
  return GS_OK;
}

static int k5_launch(void (*kern)(K5Params), int strips, int threads, size_t smem, cudaStream_t stream, const K5Params &q) {
  GS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(strips * q.csize);
  cfg.blockDim = dim3(threads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = q.csize; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  GS_CUDA(cudaLaunchKernelEx(&cfg, kern, q));
  return GS_OK;
}

static int launch_k5(gs_grid2d *g, double time, bool xdir) {
  gs_ctx *ctx = g->ctx;
  GS_TRY(k5_prepare(g, xdir, time));
  gs_k5dir &kd = xdir ? g->k5x : g->k5y;
  K5Params q;
  q.soft = 0; q.epoch = 0; q.g_buf = nullptr; q.g_sync = nullptr; q.pdl = 0;
  q.d.n = xdir ? g->cols : g->rows;
  q.d.nlines = xdir ? g->rows : g->cols;
  q.d.m = kd.m; q.d.S = kd.S;
  q.d.tab = kd.tab; q.d.omega_end = kd.omega_end; q.d.seg = kd.seg;
  q.d.e0a = kd.ends_host[0]; q.d.e0b = kd.ends_host[1]; q.d.e1a = kd.ends_host[2]; q.d.e1b = kd.ends_host[3];
  q.d.bound_first = g->b[xdir ? GS_LEFT : GS_UPP].values;
  q.d.bound_last = g->b[xdir ? GS_RIGHT : GS_LOW].values;
  q.cols = g->cols; q.rows = g->rows;
  q.rhs = xdir ? g->grid : g->result;
  q.out = g->result;
  q.grid_io = g->grid;
  q.predict_out = g->predict;
  q.status = g->status;
  {
    size_t need = ((size_t)q.d.n / K5_CH + 2) * (((size_t)q.d.nlines + 31) / 32 * 32);
    if (g->k5_ckpt_len < need) {
      GS_CUDA(cudaStreamSynchronize(ctx->stream));
      cudaFree(g->k5_ckpt);
      g->k5_ckpt = nullptr;
      g->k5_ckpt_len = 0;
      GS_CUDA(cudaMalloc(&g->k5_ckpt, need * sizeof(double)));
      g->k5_ckpt_len = need;
    }
    q.ckpt = g->k5_ckpt;
  }
  const int nb = k5_nb(ctx, xdir, kd.m, kd.S, kd.cs);
  size_t smem = k5_smem(xdir, kd.m, kd.S, kd.cs, nb);
  int blocks = (q.d.nlines + 31) / 32, threads = 32 * kd.S;
  q.cta_stride = 32;        // y sweep: a CTA owns 32 adjacent columns
  q.pitch = g->cols;
  q.bc_stride = 1;
  q.strip = 32;
  q.csize = kd.cs;
  if (kd.cs > 1) {
    // clustered strips: full 32-line strips
  } else if (xdir && ctx->opt_k5_xstrip > 0 && ctx->opt_k5_xstrip <= 32) {
    // x sweep: rows per CTA may be fewer than 32 so that the number of CTAs matches the SM count
    q.strip = (int)ctx->opt_k5_xstrip;
    blocks = (q.d.nlines + q.strip - 1) / q.strip;
  } else if (xdir && ctx->opt_k5_xstrip == 0 && blocks < ctx->sm_count) {
    // fewer CTAs than SMs: narrow the strips (down to 24 rows) so that every SM gets one (+1-2 % at 4096 rows)
    q.strip = std::max(24, (q.d.nlines + ctx->sm_count - 1) / ctx->sm_count);
    blocks = (q.d.nlines + q.strip - 1) / q.strip;
  }
  if (kd.tma) {
    K5TMaps tm;
    const uint64_t c = (uint64_t)g->cols, r = (uint64_t)g->rows;
    // lines per CTA: one CTA per SM and a memory-bound kernel, so what counts is waves x strip width (4096 lines: 147
    // strips of 28 instead of 128 of 32).  The y sweep's box rows are `strip` doubles: an even number (16-byte units).
    if (kd.cs == 1) {
      const int64_t forced = xdir ? ctx->opt_k5_xstrip : ctx->opt_k5t_ystrip;
      int best = 32;
      long long best_cost = -1;
      // (y: measured -- 28-column strips = 224-byte row pieces that straddle 256-byte units cost more than the idle SMs:
      // 4096^2 0.198 against 0.183 ms per step, 16384^2 3.65 against 2.91)
      for (int w = 32; w >= (xdir ? 24 : 32); w -= (xdir ? 1 : 2)) {
        const long long ctas = (q.d.nlines + w - 1) / w, waves = (ctas + ctx->sm_count - 1) / ctx->sm_count;
        if (best_cost < 0 || waves * w < best_cost) { best = w; best_cost = waves * w; }
      }
      q.strip = (forced > 0 && forced <= 32 && (xdir || !(forced & 1))) ? (int)forced : best;
      blocks = (q.d.nlines + q.strip - 1) / q.strip;
    }
    if (kd.soft) {
      GS_TRY(k5t_soft_buffers(ctx, kd, blocks));
      q.soft = 1; q.epoch = ++kd.epoch; q.g_buf = kd.soft_buf; q.g_sync = kd.soft_sync;
    }
    if (xdir) {
      const bool wide = k5t_wide(ctx, 0, q.d.n);
      if (wide) GS_TRY(gs_encode_map2d(&tm.rhs, g->grid, c, r, (unsigned)K5T_WCOLS, (unsigned)q.strip, false, ctx->opt_k5t_promo));
      else GS_TRY(gs_encode_map2d(&tm.rhs, g->grid, c, r, 16u, (unsigned)q.strip, true, ctx->opt_k5t_promo));
      GS_TRY(gs_encode_map2d(&tm.out, g->result, c, r, 16u, (unsigned)q.strip, true, ctx->opt_k5t_promo));
      tm.gold = tm.rhs;
      tm.pout = tm.out;
      GS_TRY(k5t_launch(ctx, 0, wide, blocks, kd.S, kd.NT, kd.cs, q, tm));
    } else {
      q.cta_stride = q.strip;
      GS_TRY(gs_encode_map2d(&tm.rhs, g->result, c, r, (unsigned)q.strip, 16u, false));
      GS_TRY(gs_encode_map2d(&tm.out, g->grid, c, r, (unsigned)q.strip, 16u, false));
      tm.gold = tm.out;
      GS_TRY(gs_encode_map2d(&tm.pout, g->predict, c, r, (unsigned)q.strip, 16u, false));
      GS_TRY(k5t_launch(ctx, 1, false, blocks, kd.S, kd.NT, kd.cs, q, tm));
    }
    ctx->launches++;
    return GS_OK;
  }
  GS_REQUIRE(smem <= ctx->smem_optin, "K5 shape does not fit shared memory");
  GS_TRY(k5_launch(xdir ? k5_sweep<true, false, 2> : (nb == 3 ? k5_sweep<false, false, 3> : k5_sweep<false, false, 2>), blocks, threads,
                   smem, ctx->stream, q));
  ctx->launches++;
  GS_CUDA(cudaGetLastError());
  return GS_OK;
}

// ---- K5 for long 1-D lines (config C5): the blocked 1-D layout is a stack of 32-column strips with pitch 32 ----
bool gs_k5_1d_usable(gs_grid1d *g) {
  gs_ctx *ctx = g->ctx;
  if (!g->all_const || !ctx->opt_k5 || g->solver == GS_SOLVER_THOMAS) return false;
  if (g->n < (g->solver == GS_SOLVER_PARTITIONED ? 128 : 2048)) return false;   // AUTO keeps short lines bit-exact
  for (int i = 0; i < (g->spline_mode == GS_SPLINE_SINGLE ? 1 : 2 * g->n); ++i) {
    int off = g->spline_mode == GS_SPLINE_SINGLE ? g->single_off : g->node_off_host[i];
    if (!literal_const(ctx, off, nullptr)) return false;
  }
  int m, S;
  k5_shape(g->n, &m, &S);
  return S >= 2 && k5_smem(false, m, S) <= ctx->smem_optin;
}

int gs_k5_1d_step(gs_grid1d *g, double time, int nsteps) {
  gs_ctx *ctx = g->ctx;
  gs_k5dir &kd = g->k5;
  const int n = g->n;
  if (!(kd.valid && kd.time == time && kd.opt_sig == k5_opt_sig(ctx))) {
    kd.cs = k5_cluster(ctx, k5_sweep<false, true, 2>, false, n, (int)g->nlines, &kd.m, &kd.S);
    // K5T: a 16-node box must not straddle two 32-line blocks; box coordinates are ints
    kd.soft = false;
    kd.tma = n % K5_CH == 0 && (g->pitch / 32) * (int64_t)n < (int64_t)1 << 31 && aligned16(g->grid) &&
             aligned16(g->predict) && k5t_plan(ctx, 2, n, (int)g->nlines, &kd.m, &kd.S, &kd.NT, &kd.cs, &kd.soft);
    kd.opt_sig = k5_opt_sig(ctx);
    if (!kd.tab) {
      GS_CUDA(cudaMalloc(&kd.tab, (size_t)n * sizeof(double4)));
      GS_CUDA(cudaMalloc(&kd.omega_end, ((size_t)n / K5_CH + 2) * sizeof(double)));
      GS_CUDA(cudaMalloc(&kd.seg, 128 * sizeof(double4)));
      GS_CUDA(cudaMalloc(&kd.ends, 4 * sizeof(double)));
      size_t need = ((size_t)n / K5_CH + 2) * (size_t)g->pitch;
      GS_CUDA(cudaMalloc(&g->k5_ckpt, need * sizeof(double)));
    }
    K5Setup1d q;
    q.n = n; q.m = kd.m; q.S = kd.S * kd.cs; q.mode = g->spline_mode;
    q.predef = g->predef; q.pool = ctx->pool_dev; q.single_off = g->single_off; q.node_off = g->node_off;
    q.time = time;
    q.tab = kd.tab; q.omega_end = kd.omega_end; q.seg = kd.seg; q.ends = kd.ends; q.status = g->status;
    k5_setup_1d<<<1, 32, 0, ctx->stream>>>(q);
    ctx->launches++;
    GS_CUDA(cudaGetLastError());
    GS_CUDA(cudaMemcpyAsync(kd.ends_host, kd.ends, 4 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    GS_CUDA(cudaStreamSynchronize(ctx->stream));
    kd.valid = true;
    kd.time = time;
  }
  K5Params q;
  q.soft = 0; q.epoch = 0; q.g_buf = nullptr; q.g_sync = nullptr; q.pdl = 0;
  q.d.n = n;
  q.d.nlines = (int)g->nlines;
  q.d.m = kd.m; q.d.S = kd.S;
  q.d.tab = kd.tab; q.d.omega_end = kd.omega_end; q.d.seg = kd.seg;
  q.d.e0a = kd.ends_host[0]; q.d.e0b = kd.ends_host[1]; q.d.e1a = kd.ends_host[2]; q.d.e1b = kd.ends_host[3];
  q.d.bound_first = g->leftT;
  q.d.bound_last = g->rightT;
  q.cols = 32; q.rows = n;
  q.rhs = g->grid;
  q.out = nullptr;
  q.grid_io = g->grid;
  q.predict_out = g->predict;
  q.ckpt = g->k5_ckpt;
  q.cta_stride = (long long)n * 32;   // one 32-line block of the [line/32][node][line%32] layout per CTA
  q.pitch = 32;
  q.strip = 32;
  q.bc_stride = g->bc_stride;
  q.csize = kd.cs;
  q.status = g->status;
  const int nb = k5_nb(ctx, false, kd.m, kd.S, kd.cs);
  size_t smem = k5_smem(false, kd.m, kd.S, kd.cs, nb);
  int blocks = (int)(g->pitch / 32), threads = 32 * kd.S;
  if (kd.tma) {
    K5TMaps tm;
    const uint64_t rows = (uint64_t)(g->pitch / 32) * (uint64_t)n;
    GS_TRY(gs_encode_map2d(&tm.rhs, g->grid, 32u, rows, 32u, 16u, false));
    tm.out = tm.rhs;
    tm.gold = tm.rhs;
    GS_TRY(gs_encode_map2d(&tm.pout, g->predict, 32u, rows, 32u, 16u, false));
    if (kd.soft) {
      GS_TRY(k5t_soft_buffers(ctx, kd, blocks));
      q.soft = 1; q.g_buf = kd.soft_buf; q.g_sync = kd.soft_sync;
    }
    for (int s = 0; s < nsteps; ++s) {
      if (kd.soft) q.epoch = ++kd.epoch;
      GS_TRY(k5t_launch(ctx, 2, false, blocks, kd.S, kd.NT, kd.cs, q, tm));
      ctx->launches++;
    }
    return GS_OK;
  }
  for (int s = 0; s < nsteps; ++s) {
    GS_TRY(k5_launch(nb == 3 ? k5_sweep<false, true, 3> : k5_sweep<false, true, 2>, blocks, threads, smem, ctx->stream, q));
    ctx->launches++;
  }
  GS_CUDA(cudaGetLastError());
  return GS_OK;
}

// ---- K6: faces kernel + elimination sweeps for temperature-dependent tables --------------------------------
static int k6_chunk(gs_ctx *) { return 8; }   // nodes per chunk: 16 warps per CTA (16-node chunks / 8 warps measured slower)
static void k6_shape(gs_ctx *ctx, int n, int *m, int *S) {
  const int ch = k6_chunk(ctx);
  int s = std::max(2, std::min(ch == 16 ? 8 : 16, n / 64));
  int mm = ((n + s - 1) / s + ch - 1) / ch * ch;
  *m = mm;
  *S = (n + mm - 1) / mm;
}
// With fewer strips than half the SMs (a row shard of a larger grid, a small grid) a strip is given to a thread-block
// cluster of 2 or 4 CTAs, each eliminating S of the line's csize * S segments.  Keeps csize = 1 unless the segments
// divide evenly (an open last end tolerates no empty trailing segment).
static int k6_cluster(gs_ctx *ctx, int n, int nlines, int *m, int *S) {
  k6_shape(ctx, n, m, S);
  if (!ctx->opt_k6_cluster) return 1;
  const int ch = k6_chunk(ctx), strips = (nlines + 31) / 32;
  int cs = 1;
  while (cs < 4 && strips * cs * 2 <= ctx->sm_count) cs *= 2;
  for (; cs > 1; cs /= 2) {
    const int st = *S * cs;
    const int mm = ((n + st - 1) / st + ch - 1) / ch * ch;
    if (mm >= 2 * ch && (n + mm - 1) / mm == st) {
      *m = mm;
      return cs;
    }
  }
  return 1;
}
// ff: fused faces -- the second tile of a buffer holds ch + 1 rows of predict, and the spline pool is staged as well
static size_t k6_smem(gs_ctx *ctx, bool xdir, int S, bool ff = false, int pool_len = 0) {
  const int ch = k6_chunk(ctx);
  const size_t pitch = xdir ? 33 : 32;
  const size_t buf = (size_t)ch * pitch + (size_t)(ff ? ch + 1 : ch) * pitch;
  return ((size_t)6 * S * 32 + (size_t)S * (2 * buf + 4 * ch) + (ff ? (size_t)pool_len : 0)) * sizeof(double);
}
static bool use_k6(gs_grid2d *g, bool xdir) {
  gs_ctx *ctx = g->ctx;
  if (g->solver == GS_SOLVER_THOMAS || !ctx->opt_k6) return false;
  const int n = xdir ? g->cols : g->rows;
  if (n < (g->solver == GS_SOLVER_PARTITIONED ? 128 : 256)) return false;
  int m, S;
  k6_shape(ctx, n, &m, &S);
  return S >= 2 && k6_smem(ctx, xdir, S) <= ctx->smem_optin;
}

// arguments of the split y sweep (row-sharded grids): gs_grid2d_yiter_update_begin / _end
struct K6Split {
  int mode;                              // K6_BEGIN / K6_END
  const double *halo_prev, *halo_next;   // device, [cols]; NULL = that end is the physical boundary
  double *exported;                      // device, [6][cols]
  const double *gathered;                // device, [nranks][6][cols]
  int nranks, rank;
  double *const *export_peer = nullptr;  // single-process multi-GPU: one destination per device (see K6Params)
  int nexport_peer = 0;
};

// 1. faces + end rows of a direction (reads predict, rhs; every spline evaluation of the sweep happens here).  Shared by
// K6 (tolerance solver) and K7 (exact solver).  `skip_faces`: the sweep evaluates the interior faces itself (k6_fuse).
static int launch_faces(gs_grid2d *g, double time, bool xdir, const K6Split *split, K6FacesParams &fq, bool *ff_out,
                        bool ends_only = false) {
  gs_ctx *ctx = g->ctx;
  const int n = xdir ? g->cols : g->rows, nlines = xdir ? g->rows : g->cols;
  const size_t cells = (size_t)g->cols * g->rows;
  if (!g->k6_F && !ends_only) GS_CUDA(cudaMalloc(&g->k6_F, cells * sizeof(double)));
  const size_t ends_need = 7 * (size_t)std::max(g->cols, g->rows);
  if (!g->k6_ends) GS_CUDA(cudaMalloc(&g->k6_ends, ends_need * sizeof(double)));
  const int mode = split ? split->mode : K6_FUSED;
  bool single;
  fill_common(g, fq.sw, time, single);
  Sweep2dParams &p = fq.sw;
  const gs_dim2d &d = xdir ? g->x : g->y;
  p.rhs = xdir ? g->grid : g->result;
  p.coefs = d.coefs;
  p.firstVol = d.hf[0]; p.fC = d.hf[1]; p.lastVol = d.hf[2]; p.lA = d.hf[3];
  const int sf = xdir ? GS_LEFT : GS_UPP, sl = xdir ? GS_RIGHT : GS_LOW;
  p.bound_first = g->b[sf].values; p.bound_last = g->b[sl].values;
  p.kind_first = g->b[sf].kind; p.kind_last = g->b[sl].kind;
  fq.F = g->k6_F;
  fq.first3 = g->k6_ends;
  fq.last3 = g->k6_ends + 3 * (size_t)nlines;
  fq.open_first = split && split->halo_prev;
  fq.open_last = split && split->halo_next;
  fq.halo_prev = split ? split->halo_prev : nullptr;
  fq.halo_next = split ? split->halo_next : nullptr;
  fq.fm1 = g->k6_split;
  fq.face0 = ends_only ? g->k6_ends + 6 * (size_t)std::max(g->cols, g->rows) : nullptr;
  size_t pool_smem = p.smem_pool ? (size_t)p.pool_len * sizeof(double) : 0;
  // blocks of 256 columns x a stride of rows, ONE wave: as many CTAs as are resident at once, equal rows each
  // (several short waves leave the chip part-empty while the last one drains)
  const int fbx = (g->cols + 255) / 256;
  bool lit = single;
  for (int s4 = 0; s4 < 4 && lit; ++s4) lit = literal_const(ctx, p.maps.single_off[s4], nullptr);
  const bool tab = !lit && p.smem_pool;
  const int eblocks = (nlines + 127) / 128;   // (one warp per CTA measured slower: every CTA stages the pool)
  // fused faces: the sweep evaluates the interior faces itself (phase A), only the end rows need their own kernel
  const bool ff = ends_only || (ff_out && tab && ctx->opt_k6_fuse && k6_chunk(ctx) == 8);
  if (ff_out) *ff_out = ff;
  if (ff && !ends_only) {
    int m_, S_;
    k6_shape(ctx, n, &m_, &S_);
    GS_REQUIRE(k6_smem(ctx, xdir, S_, true, p.pool_len) <= ctx->smem_optin, "K6 shared memory");
  }
#define GS_FACES(XD, SG, TB)                                        \
  do {                                                              \
    k6_ends<XD, SG><<<eblocks, 128, pool_smem, ctx->stream>>>(fq);  \
    ctx->launches++;                                                \
    if (ff) break;                                                  \
    int resident = 0;                                               \
    GS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, k6_faces<XD, SG, TB>, 256, pool_smem)); \
    dim3 fblocks(fbx, std::max(1, std::min(g->rows, ctx->sm_count * std::max(resident, 1) / fbx)));          \
    if (TB) fblocks = dim3((unsigned)std::max<long long>(1, std::min<long long>((long long)ctx->sm_count * std::max(resident, 1), (cells / 32 + 7) / 8)), 1); \
    k6_faces<XD, SG, TB><<<fblocks, 256, pool_smem, ctx->stream>>>(fq); \
    ctx->launches++;                                                \
  } while (0)
  if (mode == K6_END) {
    // the faces and end rows of the begin call are still in place
  } else if (xdir) {
    if (lit) GS_FACES(true, CoefSelLit, false);
    else if (tab) { if (single) GS_FACES(true, CoefSel<true>, true); else GS_FACES(true, CoefSel<false>, true); }
    else if (single) GS_FACES(true, CoefSel<true>, false);
    else GS_FACES(true, CoefSel<false>, false);
  } else {
    if (lit) GS_FACES(false, CoefSelLit, false);
    else if (tab) { if (single) GS_FACES(false, CoefSel<true>, true); else GS_FACES(false, CoefSel<false>, true); }
    else if (single) GS_FACES(false, CoefSel<true>, false);
    else GS_FACES(false, CoefSel<false>, false);
  }
#undef GS_FACES
  GS_CUDA(cudaGetLastError());
  return GS_OK;
}

static int launch_k6(gs_grid2d *g, double time, bool xdir, const K6Split *split = nullptr) {
  gs_ctx *ctx = g->ctx;
  GS_REQUIRE(!ctx->spline_off.empty(), "no spline has been created in this context");
  GS_TRY(gs_ctx_sync_pool(ctx));
  const int n = xdir ? g->cols : g->rows, nlines = xdir ? g->rows : g->cols;
  const size_t ck_need = ((size_t)std::max(g->cols, g->rows) / 8 + 2) * (((size_t)std::max(g->cols, g->rows) + 31) / 32 * 32);
  if (!g->k6_ckpt) GS_CUDA(cudaMalloc(&g->k6_ckpt, ck_need * sizeof(double4)));
  const int mode = split ? split->mode : K6_FUSED;
  const size_t nl_pad = ((size_t)nlines + 31) / 32 * 32;
  if (split && !g->k6_split) GS_CUDA(cudaMalloc(&g->k6_split, (nl_pad + 64 * 3 * nl_pad) * sizeof(double)));
  K6FacesParams fq;
  bool ff = false;
  GS_TRY(launch_faces(g, time, xdir, split, fq, &ff));
  Sweep2dParams &p = fq.sw;
  const gs_dim2d &d = xdir ? g->x : g->y;
  // 2. elimination sweep
  K6Params q;
  q.n = n; q.nlines = nlines; q.cols = g->cols; q.rows = g->rows;
  q.csize = k6_cluster(ctx, n, nlines, &q.m, &q.S);
  q.coefs = d.coefs;
  q.F = g->k6_F; q.first3 = fq.first3; q.last3 = fq.last3;
  q.rhs = p.rhs;
  q.out = g->result;
  q.grid_io = g->grid;
  q.predict_out = g->predict;
  q.ckpt = g->k6_ckpt;
  q.time = time;
  q.status = g->status;
  q.open_first = fq.open_first; q.open_last = fq.open_last;
  q.fm1 = g->k6_split;
  q.cdw = g->k6_split ? g->k6_split + nl_pad : nullptr;
  q.exported = split ? split->exported : nullptr;
  q.nexport_peer = split ? split->nexport_peer : 0;
  for (int t = 0; t < q.nexport_peer; ++t) q.export_peer[t] = split->export_peer[t];
  q.gathered = split ? split->gathered : nullptr;
  q.nranks = split ? split->nranks : 1;
  q.rank = split ? split->rank : 0;
  q.predict = p.predict; q.Fout = g->k6_F; q.pool = p.pool; q.pool_len = p.pool_len; q.maps = p.maps;
  size_t smem = k6_smem(ctx, xdir, q.S, ff, p.pool_len);
  int blocks = (nlines + 31) / 32, threads = 32 * q.S;
  void (*kern)(K6Params);
  if (ff) {
    if (mode == K6_BEGIN) kern = k6_sweep<false, 8, K6_BEGIN, true>;
    else if (mode == K6_END) kern = k6_sweep<false, 8, K6_END, true>;
    else kern = xdir ? k6_sweep<true, 8, K6_FUSED, true> : k6_sweep<false, 8, K6_FUSED, true>;
  } else if (mode == K6_BEGIN) kern = k6_sweep<false, 8, K6_BEGIN, false>;
  else if (mode == K6_END) kern = k6_sweep<false, 8, K6_END, false>;
  else kern = xdir ? k6_sweep<true, 8, K6_FUSED, false> : k6_sweep<false, 8, K6_FUSED, false>;
  GS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(blocks * q.csize);
    cfg.blockDim = dim3(threads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = ctx->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = q.csize; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    GS_CUDA(cudaLaunchKernelEx(&cfg, kern, q));
  }
  ctx->launches++;
  GS_CUDA(cudaGetLastError());
  return GS_OK;
}

// ---- K7: exact sweeps (faces pass + one thread per line in the reference's order) -----------------------------
static bool use_k7(gs_grid2d *g, bool xdir) {
  gs_ctx *ctx = g->ctx;
  if (!ctx->opt_k7 || g->solver == GS_SOLVER_PARTITIONED) return false;
  const int n = xdir ? g->cols : g->rows;
  // bulk copies move 16-byte units: row pieces must start and end on even columns
  if (n < 4 || (g->cols & 1)) return false;
  return aligned16(g->grid) && aligned16(g->predict) && aligned16(g->result);
}
static int k7_encode(K7Map *out, const void *base, int cols, int rows, bool xbox, bool pred_box = false, int promo = 0) {
  // pred_box (fused faces): one more node along the line, 18 columns (a 16-byte multiple) without the swizzle for x
  return gs_encode_map2d(out, base, (uint64_t)cols, (uint64_t)rows, xbox ? (pred_box ? 18u : 16u) : 32u,
                         xbox ? 32u : (pred_box ? 17u : 16u), xbox && !pred_box, xbox ? promo : 0);
}

static int launch_k7(gs_grid2d *g, double time, bool xdir, bool fuse) {
  gs_ctx *ctx = g->ctx;
  GS_REQUIRE(!ctx->spline_off.empty(), "no spline has been created in this context");
  GS_TRY(gs_ctx_sync_pool(ctx));
  const int n = xdir ? g->cols : g->rows, nlines = xdir ? g->rows : g->cols;
  const int strips = (nlines + 31) / 32;
  const size_t need = (size_t)strips * 32 * (size_t)n;
  if (g->k7_scratch_elems < need) {
    GS_CUDA(cudaStreamSynchronize(ctx->stream));
    cudaFree(g->k7_scratch);
    g->k7_scratch = nullptr;
    g->k7_scratch_elems = 0;
    // enough for either direction
    const size_t ex = (size_t)((g->rows + 31) / 32) * 32 * (size_t)g->cols, ey = (size_t)((g->cols + 31) / 32) * 32 * (size_t)g->rows;
    const size_t elems = std::max(ex, ey);
    GS_CUDA(cudaMalloc(&g->k7_scratch, elems * sizeof(double2)));
    g->k7_scratch_elems = elems;
  }
  // fused faces (k7f_sweep): the spline pool must fit its 4 KB of shared memory, or all coefficients are literal constants
  bool lit = true;
  for (int s4 = 0; s4 < 4 && lit; ++s4)
    lit = g->map_mode[s4] == GS_MAP_SINGLE && literal_const(ctx, ctx->spline_off[g->map_host[s4][0]], nullptr);
  // what is staged in shared memory: the whole pool if it fits, else the one table of a SINGLE `apply` map
  int stage_off = 0, stage_len = (int)ctx->pool.size();
  if ((size_t)stage_len * sizeof(double) > K7F_POOL && g->map_mode[GS_SEL_APPLY] == GS_MAP_SINGLE) {
    stage_off = ctx->spline_off[g->map_host[GS_SEL_APPLY][0]];
    const int np = (int)ctx->pool[stage_off + 1];
    stage_len = GS_HDR + (np + 1) + np * GS_PIECE;
  }
  // k7_fuse: 0 never; 1 literal constants only; 2 (default) tables too.  MEASURED (B200, ms/step, fused against the
  // separate faces pass): 4096^2 constants 0.87 / 1.24; 4096^2 Hermite tables 1.25 / 1.29 with two faces warps (one
  // faces warp: 1.88 -- it needs ~4500 clk for the 16 x 32 table faces of a chunk, the chain warp 2400); 16384^2 Hermite
  // tables 7.14 / 8.66.
  const bool fused = lit ? ctx->opt_k7_fuse >= 1
                         : (ctx->opt_k7_fuse >= 2 && (size_t)stage_len * sizeof(double) <= K7F_POOL);
  K6FacesParams fq;
  GS_TRY(launch_faces(g, time, xdir, nullptr, fq, nullptr, fused));
  K7Params q;
  q.n = n; q.nlines = nlines; q.cols = g->cols; q.rows = g->rows;
  q.coefs = (xdir ? g->x : g->y).coefs;
  q.first3 = fq.first3; q.last3 = fq.last3;
  q.out = g->result;
  q.grid_io = g->grid;
  q.predict_out = g->predict;
  q.scratch = g->k7_scratch;
  q.time = time;
  q.status = g->status;
  q.face0 = fq.face0;
  q.pool = fq.sw.pool;
  q.pool_len = fq.sw.pool_len;
  q.stage_off = stage_off; q.stage_len = stage_len;
  q.nsb = (int)ctx->opt_k7_nsb;
  q.maps = fq.sw.maps;
  K7Maps tm;
  const int promo = (int)ctx->opt_k7_promo;   // x sweep boxes: 256-byte L2 promotion (measured neutral: off)
  GS_TRY(k7_encode(&tm.F, fused ? (const void *)g->predict : (const void *)fq.F, g->cols, g->rows, xdir, false, promo));
  GS_TRY(k7_encode(&tm.rhs, fq.sw.rhs, g->cols, g->rows, xdir, false, promo));
  GS_TRY(k7_encode(&tm.gold, g->grid, g->cols, g->rows, xdir, false, promo));
  GS_TRY(k7_encode(&tm.out, g->result, g->cols, g->rows, xdir, false, promo));
  GS_TRY(k7_encode(&tm.pred, g->predict, g->cols, g->rows, xdir, true));
  if (fused) {
    // literal constants: one faces warp keeps up with the chain warp; tables: two (each takes half of a chunk's nodes)
    void (*kern)(K7Params, const K7Maps);
    // faces warps per CTA: literal constants 1; tables: while an SM holds few strips the faces warps bound the pipeline and
    // more of them pay -- MEASURED ms per step with 2 / 3 / 4 faces warps: 4096^2 (one strip per SM) 1.229 / 1.160 / 0.934
    // (eight: 1.116), 8192^2 (two per SM) 2.770 / 2.585 / 3.536, 16384^2 (3.5 per SM: the SM itself is busy) 7.24 / 7.35 / -
    // (four warps at the 102 registers that keep four CTAs resident spill: 8.48)
    int nfw = 2;
    if (ctx->opt_k7f_nfw >= 2 && ctx->opt_k7f_nfw <= 4) nfw = (int)ctx->opt_k7f_nfw;
    else if (strips <= ctx->sm_count) nfw = 4;          // one strip per SM: four faces warps with all the registers they want
    else if (strips <= 2 * ctx->sm_count) nfw = 3;
    if (lit) nfw = 1;
#define K7F_PICK(X, F) (lit ? k7f_sweep<X, F, true, 1> : nfw == 4 ? k7f_sweep<X, F, false, 4> : nfw == 3 ? k7f_sweep<X, F, false, 3> : k7f_sweep<X, F, false, 2>)
    if (xdir) kern = K7F_PICK(true, false);
    else if (fuse) kern = K7F_PICK(false, true);
    else kern = K7F_PICK(false, false);
#undef K7F_PICK
    const int threads = 32 + 32 * nfw;
    // function attributes are per device: remembered per context
    const void *key = reinterpret_cast<const void *>(kern);
    if (std::find(ctx->configured.begin(), ctx->configured.end(), key) == ctx->configured.end()) {
      GS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, K7F_SMEM));
      GS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
      ctx->configured.push_back(key);
    }
    kern<<<strips, threads, K7F_SMEM, ctx->stream>>>(q, tm);
  } else if (xdir) k7_sweep<true, false><<<strips, 32, K7_SMEM, ctx->stream>>>(q, tm);
  else if (fuse) k7_sweep<false, true><<<strips, 32, K7_SMEM, ctx->stream>>>(q, tm);
  else k7_sweep<false, false><<<strips, 32, K7_SMEM, ctx->stream>>>(q, tm);
  ctx->launches++;
  GS_CUDA(cudaGetLastError());
  return GS_OK;
}

// Solver choice per sweep.  GS_SOLVER_THOMAS: K7 (exact), first-draft exact kernels where K7 does not apply.
// GS_SOLVER_PARTITIONED: K5 / K6 / first-generation partitioned kernels (tolerance family).  GS_SOLVER_AUTO: K5 for
// literal-constant coefficients (<= 1e-12 measured at the configured step counts, profiles/r2_parity.json); everything
// else -- spline tables, coefficient maps, short lines -- takes the exact K7, because a reordered elimination cannot
// stay within 1e-12 of the reference on tables; option auto_exact = 0 restores the round-1 choice.
static bool auto_exact(gs_grid2d *g, bool xdir) {
  return g->solver == GS_SOLVER_AUTO && g->ctx->opt_auto_exact && use_k7(g, xdir);
}
static int sweep_x(gs_grid2d *g, double time) {
  if (use_k5(g, true)) return launch_k5(g, time, true);
  if (auto_exact(g, true)) return launch_k7(g, time, true, false);
  if (use_k6(g, true)) return launch_k6(g, time, true);
  if (use_partitioned(g, g->cols)) return launch_part(g, time, true, false);
  return use_k7(g, true) ? launch_k7(g, time, true, false) : launch_xsweep(g, time);
}
static int sweep_y(gs_grid2d *g, double time, bool fuse) {
  if (fuse && use_k5(g, false)) return launch_k5(g, time, false);   // K5's / K6's y sweeps are always fused with Grid.update
  if (auto_exact(g, false)) return launch_k7(g, time, false, fuse);
  if (fuse && use_k6(g, false)) return launch_k6(g, time, false);
  if (use_partitioned(g, g->rows)) return launch_part(g, time, false, fuse);
  return use_k7(g, false) ? launch_k7(g, time, false, fuse) : launch_ysweep(g, time, fuse);
}
extern "C" int gs_grid2d_xiter(gs_grid2d *g, double time) {
  GS_REQUIRE(g != nullptr, "NULL argument");
  GS_TRY(gs_ctx_use(g->ctx));
  g->result_is_grid = false;  // xIter overwrites every element of result
  return sweep_x(g, time);
}
extern "C" int gs_grid2d_yiter(gs_grid2d *g, double time) {
  GS_REQUIRE(g != nullptr, "NULL argument");
  GS_TRY(gs_ctx_use(g->ctx));
  GS_TRY(materialise_result(g));
  return sweep_y(g, time, false);
}
// yIter followed by Grid.update in one pass (what iteration() runs after the x sweep); used by the row-sharded
// multi-GPU driver, which has to exchange `result` between the two sweeps
extern "C" int gs_grid2d_yiter_update(gs_grid2d *g, double time) {
  GS_REQUIRE(g != nullptr, "NULL argument");
  GS_TRY(gs_ctx_use(g->ctx));
  GS_TRY(materialise_result(g));
  GS_TRY(sweep_y(g, time, true));
  // (a caller-owned `result` buffer is the caller's business: it is not lazily refreshed from `grid`)
  g->result_is_grid = !g->bound[GS_RESULT];
  return GS_OK;
}
// The y sweep + Grid.update of a ROW-SHARDED grid, split around the ranks' exchange (partitioned solver only).  This
// grid holds rows [r0, r1) of every column; `halo_prev` / `halo_next` are the predict rows just above / below them
// (device, cols doubles; NULL where this rank holds the physical boundary, whose bound then applies).
//   begin: faces, forward elimination of the local rows and their condensation into ONE segment per column ->
//          exported[6][cols] (device).  The caller all-gathers `exported` of all ranks, in rank order.
//   end:   every rank solves the ranks' reduced system from gathered[nranks][6][cols], then back-substitutes, updates
//          grid and predict.
extern "C" int gs_grid2d_yiter_update_begin(gs_grid2d *g, double time, const void *halo_prev, const void *halo_next,
                                            void *exported) {
  GS_REQUIRE(g && exported, "NULL argument");
  GS_TRY(gs_ctx_use(g->ctx));
  GS_REQUIRE(g->solver != GS_SOLVER_THOMAS, "the split y sweep needs the partitioned solver (GS_SOLVER_AUTO / PARTITIONED)");
  GS_REQUIRE(g->ctx->opt_k6 && g->rows >= 16, "the split y sweep needs K6 and at least 16 local rows");
  {
    int m, S;
    k6_shape(g->ctx, g->rows, &m, &S);
    GS_REQUIRE(S >= 2 && S <= 16 && k6_smem(g->ctx, false, S) <= g->ctx->smem_optin, "local rows do not fit the K6 shape");
  }
  GS_REQUIRE(!(halo_next == nullptr && g->b[GS_LOW].kind == GS_FLOW && halo_prev != nullptr),
             "a lower HeatFlow bound is not supported on a sharded grid (it reads the neighbouring column's last row)");
  GS_TRY(materialise_result(g));
  K6Split sp{K6_BEGIN, (const double *)halo_prev, (const double *)halo_next, (double *)exported, nullptr, 1, 0};
  GS_TRY(launch_k6(g, time, false, &sp));
  g->k6_split_open = true;
  return GS_OK;
}
// begin with the export written straight into several (peer) destinations: used by gs_sharded2d
int gs_grid2d_split_begin_peers(gs_grid2d *g, double time, const double *halo_prev, const double *halo_next,
                                double *const *dsts, int ndst) {
  GS_REQUIRE(g && dsts && ndst >= 1 && ndst <= GS_MAX_SHARDS, "bad peer destinations");
  GS_TRY(gs_ctx_use(g->ctx));
  GS_REQUIRE(g->solver != GS_SOLVER_THOMAS && g->ctx->opt_k6 && g->rows >= 16, "the split y sweep needs K6 and at least 16 local rows");
  {
    int m, S;
    k6_shape(g->ctx, g->rows, &m, &S);
    GS_REQUIRE(S >= 2 && S <= 16 && k6_smem(g->ctx, false, S) <= g->ctx->smem_optin, "local rows do not fit the K6 shape");
  }
  GS_TRY(materialise_result(g));
  K6Split sp{K6_BEGIN, halo_prev, halo_next, dsts[0], nullptr, 1, 0};
  sp.export_peer = dsts;
  sp.nexport_peer = ndst;
  GS_TRY(launch_k6(g, time, false, &sp));
  g->k6_split_open = true;
  return GS_OK;
}
extern "C" int gs_grid2d_yiter_update_end(gs_grid2d *g, double time, const void *gathered, int nranks, int rank) {
  GS_REQUIRE(g && gathered, "NULL argument");
  GS_REQUIRE(nranks >= 1 && nranks <= 16 && rank >= 0 && rank < nranks, "bad rank / nranks (at most 16 ranks)");
  GS_REQUIRE(g->k6_split_open, "gs_grid2d_yiter_update_end without a matching _begin");
  GS_TRY(gs_ctx_use(g->ctx));
  K6Split sp{K6_END, rank > 0 ? (const double *)gathered : nullptr, rank + 1 < nranks ? (const double *)gathered : nullptr,
             nullptr, (const double *)gathered, nranks, rank};
  GS_TRY(launch_k6(g, time, false, &sp));
  g->k6_split_open = false;
  g->result_is_grid = !g->bound[GS_RESULT];
  return GS_OK;
}
// device-to-device copies of a whole array on the context stream (dptr: rows * cols doubles, row-major)
extern "C" int gs_grid2d_copy_from_device(gs_grid2d *g, int which, const void *dptr) {
  GS_REQUIRE(g && dptr, "NULL argument");
  GS_REQUIRE(which >= GS_GRID && which <= GS_RESULT, "unknown array");
  GS_TRY(gs_ctx_use(g->ctx));
  if (which == GS_RESULT) g->result_is_grid = false;
  if (which == GS_GRID) GS_TRY(materialise_result(g));
  GS_CUDA(cudaMemcpyAsync(array2d(g, which), dptr, (size_t)g->cols * g->rows * sizeof(double), cudaMemcpyDeviceToDevice,
                          g->ctx->stream));
  return GS_OK;
}
extern "C" int gs_grid2d_copy_to_device(gs_grid2d *g, int which, void *dptr) {
  GS_REQUIRE(g && dptr, "NULL argument");
  GS_REQUIRE(which >= GS_GRID && which <= GS_RESULT, "unknown array");
  GS_TRY(gs_ctx_use(g->ctx));
  if (which == GS_RESULT) GS_TRY(materialise_result(g));
  GS_CUDA(cudaMemcpyAsync(dptr, array2d(g, which), (size_t)g->cols * g->rows * sizeof(double), cudaMemcpyDeviceToDevice,
                          g->ctx->stream));
  return GS_OK;
}

// Makes the grid use caller-owned device memory (rows * cols doubles, row-major) for one of its arrays, e.g. a torch
// tensor that is also the send / receive buffer of a collective: no staging copies.  The previous contents are not
// copied; the caller keeps the memory alive and on the context's device.
extern "C" int gs_grid2d_bind_device(gs_grid2d *g, int which, void *dptr) {
  GS_REQUIRE(g && dptr, "NULL argument");
  GS_REQUIRE(which >= GS_GRID && which <= GS_RESULT, "unknown array");
  GS_TRY(gs_ctx_use(g->ctx));
  if (which == GS_GRID) GS_TRY(materialise_result(g));   // the old grid buffer may be the only copy of `result`
  GS_CUDA(cudaStreamSynchronize(g->ctx->stream));
  double **slot = which == GS_GRID ? &g->grid : which == GS_PREDICT ? &g->predict : &g->result;
  if (!g->bound[which]) cudaFree(*slot);
  *slot = static_cast<double *>(dptr);
  g->bound[which] = true;
  if (which == GS_RESULT) g->result_is_grid = false;
  return GS_OK;
}

extern "C" int gs_grid2d_update(gs_grid2d *g) {
  GS_REQUIRE(g != nullptr, "NULL argument");
  gs_ctx *ctx = g->ctx;
  GS_TRY(gs_ctx_use(ctx));
  GS_TRY(materialise_result(g));
  long long count = (long long)g->cols * g->rows;
  int blocks = (int)std::min<long long>((count + 255) / 256, (long long)ctx->sm_count * 16);
  k_update2d<<<blocks, 256, 0, ctx->stream>>>(g->grid, g->predict, g->result, count);
  ctx->launches++;
  GS_CUDA(cudaGetLastError());
  return GS_OK;
}
extern "C" int gs_grid2d_iteration(gs_grid2d *g, double time, int nsteps) {
  GS_REQUIRE(g != nullptr, "NULL argument");
  GS_REQUIRE(nsteps >= 0, "nsteps must be >= 0");
  GS_TRY(gs_ctx_use(g->ctx));
  for (int s = 0; s < nsteps; ++s) {
    GS_TRY(sweep_x(g, time));
    GS_TRY(sweep_y(g, time, true));
    g->result_is_grid = !g->bound[GS_RESULT];
  }
  return GS_OK;
}
extern "C" int gs_grid2d_update_predicted(gs_grid2d *g) {
  GS_REQUIRE(g != nullptr, "NULL argument");
  GS_TRY(gs_ctx_use(g->ctx));
  GS_CUDA(cudaMemcpyAsync(g->predict, g->grid, (size_t)g->cols * g->rows * sizeof(double), cudaMemcpyDeviceToDevice,
                          g->ctx->stream));
  return GS_OK;
}

// ---- row / column slices and coordinate-wise fills (TwoDGrid.row / col / updateRow / updateColumn :260-330,
// updateX / updateY :29-58): strided gathers / scatters on the device, one contiguous copy across PCIe ----
__global__ void k_scatter(double *a, long long start, long long stride, int count, const double *in) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < count) a[start + (long long)i * stride] = in[i];
}
// a[row][col] = per_col ? v[col] : v[row]
__global__ void k_fill_lines(double *a, int cols, int rows, const double *v, int per_col) {
  const int col = blockIdx.x * blockDim.x + threadIdx.x;
  if (col >= cols) return;
  for (int row = blockIdx.y; row < rows; row += gridDim.y) a[(size_t)row * cols + col] = per_col ? v[col] : v[row];
}
static int slice_geometry(gs_grid2d *g, bool is_row, int index, long long *start, long long *stride, int *len) {
  GS_REQUIRE(index >= 0 && index < (is_row ? g->rows : g->cols), "row / column index out of range");
  if (is_row) { *start = (long long)index * g->cols; *stride = 1; *len = g->cols; }
  else { *start = index; *stride = g->cols; *len = g->rows; }
  return GS_OK;
}
static int get_slice(gs_grid2d *g, int which, bool is_row, int index, double *host) {
  GS_REQUIRE(g && host, "NULL argument");
  GS_REQUIRE(which >= GS_GRID && which <= GS_RESULT, "unknown array");
  gs_ctx *ctx = g->ctx;
  GS_TRY(gs_ctx_use(ctx));
  long long start, stride;
  int len;
  GS_TRY(slice_geometry(g, is_row, index, &start, &stride, &len));
  if (which == GS_RESULT) GS_TRY(materialise_result(g));
  GS_TRY(gs_ctx_stage(ctx, (size_t)len * sizeof(double)));
  k_gather<<<(len + 127) / 128, 128, 0, ctx->stream>>>(array2d(g, which), start, stride, len, ctx->stage);
  ctx->launches++;
  GS_CUDA(cudaGetLastError());
  GS_CUDA(cudaMemcpyAsync(host, ctx->stage, (size_t)len * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  GS_CUDA(cudaStreamSynchronize(ctx->stream));
  return GS_OK;
}
static int set_slice(gs_grid2d *g, bool is_row, int index, const double *host) {
  GS_REQUIRE(g && host, "NULL argument");
  gs_ctx *ctx = g->ctx;
  GS_TRY(gs_ctx_use(ctx));
  long long start, stride;
  int len;
  GS_TRY(slice_geometry(g, is_row, index, &start, &stride, &len));
  GS_TRY(materialise_result(g));   // `result` stops mirroring `grid` once grid is edited
  g->result_is_grid = false;
  GS_TRY(gs_ctx_stage(ctx, (size_t)len * sizeof(double)));
  GS_CUDA(cudaMemcpyAsync(ctx->stage, host, (size_t)len * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  k_scatter<<<(len + 127) / 128, 128, 0, ctx->stream>>>(g->grid, start, stride, len, ctx->stage);
  ctx->launches++;
  GS_CUDA(cudaGetLastError());
  GS_CUDA(cudaStreamSynchronize(ctx->stream));   // the caller may reuse `host` right away
  return GS_OK;
}
extern "C" int gs_grid2d_get_row(gs_grid2d *g, int which, int row, double *host) { return get_slice(g, which, true, row, host); }
extern "C" int gs_grid2d_get_col(gs_grid2d *g, int which, int col, double *host) { return get_slice(g, which, false, col, host); }
extern "C" int gs_grid2d_set_row(gs_grid2d *g, int row, const double *host) { return set_slice(g, true, row, host); }
extern "C" int gs_grid2d_set_col(gs_grid2d *g, int col, const double *host) { return set_slice(g, false, col, host); }
static int fill_lines(gs_grid2d *g, const double *host, bool per_col) {
  GS_REQUIRE(g && host, "NULL argument");
  gs_ctx *ctx = g->ctx;
  GS_TRY(gs_ctx_use(ctx));
  const int len = per_col ? g->cols : g->rows;
  GS_TRY(materialise_result(g));
  g->result_is_grid = false;
  GS_TRY(gs_ctx_stage(ctx, (size_t)len * sizeof(double)));
  GS_CUDA(cudaMemcpyAsync(ctx->stage, host, (size_t)len * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  dim3 blocks((g->cols + 255) / 256, std::max(1, std::min(g->rows, 8 * ctx->sm_count)));
  k_fill_lines<<<blocks, 256, 0, ctx->stream>>>(g->grid, g->cols, g->rows, ctx->stage, per_col ? 1 : 0);
  ctx->launches++;
  GS_CUDA(cudaGetLastError());
  GS_CUDA(cudaStreamSynchronize(ctx->stream));
  return GS_OK;
}
// updateX(f): every node of column c gets f(x_c); updateY(f): every node of row r gets f(y_r).  The host evaluates f.
extern "C" int gs_grid2d_update_x(gs_grid2d *g, const double *value_per_col) { return fill_lines(g, value_per_col, true); }
extern "C" int gs_grid2d_update_y(gs_grid2d *g, const double *value_per_row) { return fill_lines(g, value_per_row, false); }

// edge of `grid` into dst (device, side length)
static int gather_edge(gs_grid2d *g, int side, double *dst) {
  gs_ctx *ctx = g->ctx;
  int len = side_len(g, side);
  long long start = 0, stride = 1;
  switch (side) {
    case GS_UPP: start = 0; stride = 1; break;
    case GS_LOW: start = (long long)g->cols * (g->rows - 1); stride = 1; break;
    case GS_RIGHT: start = g->cols - 1; stride = g->cols; break;
    default: start = 0; stride = g->cols; break;
  }
  k_gather<<<(len + 127) / 128, 128, 0, ctx->stream>>>(g->grid, start, stride, len, dst);
  ctx->launches++;
  GS_CUDA(cudaGetLastError());
  return GS_OK;
}

extern "C" int gs_grid2d_edge_average(gs_grid2d *g, int side, double *out) {
  GS_REQUIRE(g && out, "NULL argument");
  GS_REQUIRE(side >= 0 && side < 4, "unknown side");
  gs_ctx *ctx = g->ctx;
  GS_TRY(gs_ctx_use(ctx));
  int len = side_len(g, side);
  GS_TRY(gs_ctx_stage(ctx, ((size_t)len + 1) * sizeof(double)));
  GS_TRY(gather_edge(g, side, ctx->stage));
  k_serial_mean<<<1, 32, 0, ctx->stream>>>(ctx->stage, len, ctx->stage + len);
  ctx->launches++;
  GS_CUDA(cudaGetLastError());
  GS_CUDA(cudaMemcpyAsync(out, ctx->stage + len, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  GS_CUDA(cudaStreamSynchronize(ctx->stream));
  return GS_OK;
}

extern "C" int gs_grid2d_no_heat_flow(gs_grid2d *g, int side) {
  GS_REQUIRE(g != nullptr, "NULL argument");
  GS_REQUIRE(side >= 0 && side < 4, "unknown side");
  gs_ctx *ctx = g->ctx;
  GS_TRY(gs_ctx_use(ctx));
  int len = side_len(g, side);
  gs_bound2d *b = &g->b[side];
  if (b->rep == GS_DENSE) return gather_edge(g, side, b->values);
  GS_TRY(gs_ctx_stage(ctx, ((size_t)len + 1) * sizeof(double)));
  GS_TRY(gather_edge(g, side, ctx->stage));
  k_serial_mean<<<1, 32, 0, ctx->stream>>>(ctx->stage, len, ctx->stage + len);
  k_broadcast<<<(len + 127) / 128, 128, 0, ctx->stream>>>(ctx->stage + len, b->values, len);
  ctx->launches += 2;
  GS_CUDA(cudaGetLastError());
  return GS_OK;
}

extern "C" int gs_grid2d_av_value(gs_grid2d *g, double *out) {
  GS_REQUIRE(g && out, "NULL argument");
  gs_ctx *ctx = g->ctx;
  GS_TRY(gs_ctx_use(ctx));
  long long count = (long long)g->cols * g->rows;
  int blocks = (int)std::min<long long>((count + 255) / 256, 1024);
  GS_TRY(gs_ctx_stage(ctx, ((size_t)blocks + 1) * sizeof(double)));
  k_block_sums<<<blocks, 256, 0, ctx->stream>>>(g->grid, count, ctx->stage);
  ctx->launches++;
  GS_CUDA(cudaGetLastError());
  std::vector<double> partial(blocks);
  GS_CUDA(cudaMemcpyAsync(partial.data(), ctx->stage, (size_t)blocks * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  GS_CUDA(cudaStreamSynchronize(ctx->stream));
  double s = 0.0;
  for (double v : partial) s += v;
  *out = s / (double)count;
  return GS_OK;
}

extern "C" int gs_grid2d_set_solver(gs_grid2d *g, int solver) {
  GS_REQUIRE(g != nullptr, "NULL argument");
  GS_REQUIRE(solver == GS_SOLVER_AUTO || solver == GS_SOLVER_THOMAS || solver == GS_SOLVER_PARTITIONED, "unknown solver");
  g->solver = solver;
  return GS_OK;
}
extern "C" int gs_grid2d_status(gs_grid2d *g, uint32_t *flags) {
  GS_REQUIRE(g && flags, "NULL argument");
  GS_TRY(gs_ctx_use(g->ctx));
  GS_CUDA(cudaMemcpyAsync(flags, g->status, sizeof(unsigned), cudaMemcpyDeviceToHost, g->ctx->stream));
  GS_CUDA(cudaStreamSynchronize(g->ctx->stream));
  return GS_OK;
}
extern "C" int gs_grid2d_clear_status(gs_grid2d *g) {
  GS_REQUIRE(g != nullptr, "NULL argument");
  GS_TRY(gs_ctx_use(g->ctx));
  GS_CUDA(cudaMemsetAsync(g->status, 0, sizeof(unsigned), g->ctx->stream));
  return GS_OK;
}
extern "C" int gs_grid2d_device_ptr(gs_grid2d *g, int which, void **dptr) {
  GS_REQUIRE(g && dptr, "NULL argument");
  GS_REQUIRE(which >= GS_GRID && which <= GS_RESULT, "unknown array");
  GS_TRY(gs_ctx_use(g->ctx));
  if (which == GS_RESULT) GS_TRY(materialise_result(g));
  *dptr = array2d(g, which);
  return GS_OK;
}
extern "C" int gs_grid2d_bytes_per_step(gs_grid2d *g, double *bytes) {
  GS_REQUIRE(g && bytes, "NULL argument");
  *bytes = 64.0 * (double)g->cols * (double)g->rows;
  return GS_OK;
}
