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
#include <stdio.h>
#include <string.h>

#include <algorithm>

#include "gs_device.cuh"
#include "gs_internal.h"

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
  cudaFree(g->grid); cudaFree(g->predict); cudaFree(g->result);
  for (int s = 0; s < 4; ++s) { cudaFree(g->b[s].values); cudaFree(g->map[s]); }
  cudaFree(g->scratch);
  cudaFree(g->status);
  delete g;
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

extern "C" int gs_grid2d_xiter(gs_grid2d *g, double time) {
  GS_REQUIRE(g != nullptr, "NULL argument");
  GS_TRY(gs_ctx_use(g->ctx));
  g->result_is_grid = false;  // xIter overwrites every element of result
  return launch_xsweep(g, time);
}
extern "C" int gs_grid2d_yiter(gs_grid2d *g, double time) {
  GS_REQUIRE(g != nullptr, "NULL argument");
  GS_TRY(gs_ctx_use(g->ctx));
  GS_TRY(materialise_result(g));
  return launch_ysweep(g, time, false);
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
    GS_TRY(launch_xsweep(g, time));
    GS_TRY(launch_ysweep(g, time, true));
    g->result_is_grid = true;
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
