// gs_grid1d.cu -- batched 1-D grids: nlines independent OneDGrid.step() (A/OneDGrid.scala:26-34
// = passion.iteration A/passion/package.scala:52-139 + predictor A/arraygrid/package.scala:85-87).
//
// Device layout: node-major ("interleaved") arrays a[node * pitch + line], so the 32 lines of a
// warp are 32 adjacent doubles at every node step (one 256 B coalesced access per array).
// One thread owns one line and runs the reference's Thomas recurrence in the reference's
// operation order; (delta, lambda) of the forward pass go to a per-thread scratch column that is
// reused for every line the thread processes (persistent grid), laid out [node][thread] so it is
// coalesced too.
#include <stdio.h>
#include <string.h>

#include <algorithm>

#include "gs_device.cuh"
#include "gs_internal.h"

int gs_launch_predef(gs_ctx *ctx, int dir, const double *values_dev, int n, double sigma, double *coefs_dev);

struct Step1dParams {
  int n;
  long long nlines, pitch, G;
  double *grid, *predict;
  double2 *scratch;
  const double *predef;
  const double *pool;
  int pool_len;
  int single_off;
  const int *node_off;
  const double *leftT, *rightT;
  int bc_stride;
  double time;
  int nsteps;
  unsigned *status;
  int smem_predef, smem_pool;
};

// K2: one thread = one line, all nsteps of it (lines are independent and their end
// temperatures are fixed during the call, so the step loop sits inside the line loop).
template <int MODE>
__global__ void __launch_bounds__(256) k_step1d(Step1dParams p) {
  extern __shared__ double sm[];
  const int n = p.n;
  const double *pd = p.predef;
  const double *pool = p.pool;
  {
    int base = 0;
    if (p.smem_predef) {
      for (int i = threadIdx.x; i < 2 * n; i += blockDim.x) sm[i] = p.predef[i];
      pd = sm;
      base = 2 * n;
    }
    if (p.smem_pool) {
      for (int i = threadIdx.x; i < p.pool_len; i += blockDim.x) sm[base + i] = p.pool[i];
      pool = sm + base;
    }
    __syncthreads();
  }
  const long long gtid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long G = p.G, pitch = p.pitch;
  double2 *sc = p.scratch + gtid;
  const GsRcp rt = gs_rcp(p.time);
  const GsRcp r3 = gs_rcp(3.0);
  const double inv_time = 1.0 / p.time;
  unsigned flags = 0;
  GsSpline z;
  if (MODE == GS_SPLINE_SINGLE) z = gs_spline_view(pool, p.single_off);

  for (long long line = gtid; line < p.nlines; line += G) {
    double *Gp = p.grid + line;
    double *Pp = p.predict + line;
    const double firstT = p.leftT[line * p.bc_stride];
    const double lastT = p.rightT[line * p.bc_stride];
    for (int step = 0; step < p.nsteps; ++step) {
      // ---- forward: assembly + elimination (passion.iteration :52-95 / :97-139) ----
      double t1 = firstT, t2 = Pp[0], t3 = Pp[pitch];
      double condL, condR;
      if (MODE == GS_SPLINE_SINGLE) {
        condL = gs_ads_apply(z, (t1 + t2) / 2.0, flags);   // cons :21-25
        condR = gs_ads_apply(z, (t2 + t3) / 2.0, flags);
      } else {
        condL = gs_ads_apply(gs_spline_view(pool, p.node_off[0]), (t1 + t2) / 2.0, flags);
        condR = gs_ads_apply(gs_spline_view(pool, p.node_off[1]), (t2 + t3) / 2.0, flags);
      }
      double a = condL * pd[0];
      double c = condR * pd[1];
      double vect = gs_div(-Gp[0], rt) - a * t1;
      double delta, lambda;
      gs_forward_first(-(a + c) - inv_time, c, vect, delta, lambda);
      sc[0] = make_double2(delta, lambda);
      int i = 1;
#pragma unroll 4
      for (; i < n - 1; ++i) {
        t1 = t2;
        t2 = t3;
        t3 = Pp[(long long)(i + 1) * pitch];
        if (MODE == GS_SPLINE_SINGLE) {
          // node i's left face is node i-1's right face evaluated with the same spline at the
          // same argument: the same double, so it is carried instead of re-evaluated
          condL = condR;
          condR = gs_ads_apply(z, (t2 + t3) / 2.0, flags);
        } else {
          condL = gs_ads_apply(gs_spline_view(pool, p.node_off[2 * i]), (t1 + t2) / 2.0, flags);
          condR = gs_ads_apply(gs_spline_view(pool, p.node_off[2 * i + 1]), (t2 + t3) / 2.0, flags);
        }
        a = condL * pd[2 * i];
        c = condR * pd[2 * i + 1];
        vect = gs_div(-Gp[(long long)i * pitch], rt);
        gs_forward_unit(a, -(a + c) - inv_time, c, vect, delta, lambda, flags);
        sc[(long long)i * G] = make_double2(delta, lambda);
      }
      t1 = t2;
      t2 = t3;
      t3 = lastT;
      if (MODE == GS_SPLINE_SINGLE) {
        condL = condR;
        condR = gs_ads_apply(z, (t2 + t3) / 2.0, flags);
      } else {
        condL = gs_ads_apply(gs_spline_view(pool, p.node_off[2 * i]), (t1 + t2) / 2.0, flags);
        condR = gs_ads_apply(gs_spline_view(pool, p.node_off[2 * i + 1]), (t2 + t3) / 2.0, flags);
      }
      a = condL * pd[2 * i];
      c = condR * pd[2 * i + 1];
      vect = gs_div(-Gp[(long long)i * pitch], rt) - c * t3;
      gs_forward_last(a, -(a + c) - inv_time, vect, delta, lambda);

      // ---- backward (backwardPassion :345-365) fused with OneDGrid.step's predictor and
      // grid <- result (A/OneDGrid.scala:29-33) ----
      double u = 0.0;
      {
        u = delta * u + lambda;
        double g = Gp[(long long)i * pitch];
        Pp[(long long)i * pitch] = u + gs_div(u - g, r3);   // arraygrid.aVal :85-87
        Gp[(long long)i * pitch] = u;
      }
#pragma unroll 4
      for (i = n - 2; i >= 0; --i) {
        double2 dl = sc[(long long)i * G];
        u = dl.x * u + dl.y;
        double g = Gp[(long long)i * pitch];
        Pp[(long long)i * pitch] = u + gs_div(u - g, r3);
        Gp[(long long)i * pitch] = u;
      }
      if (!(fabs(u) <= 1.7976931348623157e308)) flags |= GS_FLAG_NONFINITE;
    }
  }
  gs_raise(p.status, flags);
}

// ---------------------------------------------------------------------------------
// layout helpers
// ---------------------------------------------------------------------------------
// src line-major [L][n]  ->  dst[node * pitch + line0 + l]
__global__ void k_lines_to_interleaved(const double *src, double *dst, int L, int n, long long pitch,
                                       long long line0) {
  __shared__ double tile[32][33];
  int node = blockIdx.x * 32 + threadIdx.x;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    int l = blockIdx.y * 32 + j;
    if (l < L && node < n) tile[j][threadIdx.x] = src[(long long)l * n + node];
  }
  __syncthreads();
  int l = blockIdx.y * 32 + threadIdx.x;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    int nd = blockIdx.x * 32 + j;
    if (l < L && nd < n) dst[(long long)nd * pitch + line0 + l] = tile[threadIdx.x][j];
  }
}
__global__ void k_interleaved_to_lines(const double *src, double *dst, int L, int n, long long pitch,
                                       long long line0) {
  __shared__ double tile[32][33];
  int l = blockIdx.y * 32 + threadIdx.x;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    int nd = blockIdx.x * 32 + j;
    if (l < L && nd < n) tile[j][threadIdx.x] = src[(long long)nd * pitch + line0 + l];
  }
  __syncthreads();
  int node = blockIdx.x * 32 + threadIdx.x;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    int ll = blockIdx.y * 32 + j;
    if (ll < L && node < n) dst[(long long)ll * n + node] = tile[threadIdx.x][j];
  }
}
__global__ void k_fill(double *a, long long count, double v) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < count; i += stride) a[i] = v;
}

int gs_launch_fill(gs_ctx *ctx, double *a, long long count, double v) {
  if (count <= 0) return GS_OK;
  int blocks = (int)std::min<long long>((count + 255) / 256, (long long)ctx->sm_count * 16);
  k_fill<<<blocks, 256, 0, ctx->stream>>>(a, count, v);
  ctx->launches++;
  GS_CUDA(cudaGetLastError());
  return GS_OK;
}

// ---------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------
static bool spline_is_const(gs_ctx *ctx, int id) {
  const double *h = ctx->pool.data() + ctx->spline_off[id];
  return (int)h[0] == GS_CONST && (int)h[1] == 1;
}

extern "C" int gs_grid1d_create(gs_ctx *ctx, int64_t nlines, int n, int dir, double leftX,
                                const double *rangeX, double rightX, double sigma, int spline_mode,
                                const int *spline_ids, const double *host_predef, gs_grid1d **out) {
  GS_REQUIRE(out != nullptr, "out is NULL");
  *out = nullptr;
  GS_TRY(gs_ctx_use(ctx));
  GS_REQUIRE(nlines >= 1, "nlines must be >= 1");
  GS_REQUIRE(n >= 2, "a line needs at least 2 nodes (passion.iteration reads predict(1), A/passion/package.scala:59)");
  GS_REQUIRE(dir == GS_ORTHO || dir == GS_RADIAL, "unknown direction kind");
  GS_REQUIRE(rangeX || host_predef, "rangeX is NULL and no host preDef given");
  GS_REQUIRE(spline_mode == GS_SPLINE_SINGLE || spline_mode == GS_SPLINE_PER_NODE, "unknown spline mode");
  GS_REQUIRE(spline_ids != nullptr, "spline_ids is NULL");
  int nids = spline_mode == GS_SPLINE_SINGLE ? 1 : 2 * n;
  int nspl = (int)ctx->spline_off.size();
  for (int i = 0; i < nids; ++i) GS_REQUIRE(spline_ids[i] >= 0 && spline_ids[i] < nspl, "spline id out of range");

  gs_grid1d *g = new gs_grid1d();
  g->ctx = ctx;
  g->nlines = nlines;
  g->n = n;
  g->dir = dir;
  g->pitch = (nlines + 15) / 16 * 16;
  g->spline_mode = spline_mode;
  g->all_const = true;
  for (int i = 0; i < nids; ++i) g->all_const = g->all_const && spline_is_const(ctx, spline_ids[i]);
  int status = GS_OK;
  auto fail = [&](int s) { gs_grid1d_destroy(g); return s; };
#define G1_CUDA(expr)                                                                        \
  do {                                                                                       \
    cudaError_t e__ = (expr);                                                                \
    if (e__ != cudaSuccess) {                                                                \
      gs_set_error("%s failed: %s", #expr, cudaGetErrorString(e__));                         \
      return fail(e__ == cudaErrorMemoryAllocation ? GS_ERR_OOM : GS_ERR_CUDA);              \
    }                                                                                        \
  } while (0)
  size_t elems = (size_t)g->pitch * n;
  G1_CUDA(cudaMalloc(&g->grid, elems * sizeof(double)));
  G1_CUDA(cudaMalloc(&g->predict, elems * sizeof(double)));
  G1_CUDA(cudaMalloc(&g->predef, 2 * (size_t)n * sizeof(double)));
  G1_CUDA(cudaMalloc(&g->status, sizeof(unsigned)));
  G1_CUDA(cudaMemsetAsync(g->status, 0, sizeof(unsigned), ctx->stream));
  G1_CUDA(cudaMalloc(&g->leftT, sizeof(double)));
  G1_CUDA(cudaMalloc(&g->rightT, sizeof(double)));
  if (spline_mode == GS_SPLINE_SINGLE) {
    g->single_off = ctx->spline_off[spline_ids[0]];
  } else {
    std::vector<int> offs(2 * (size_t)n);
    for (int i = 0; i < 2 * n; ++i) offs[i] = ctx->spline_off[spline_ids[i]];
    G1_CUDA(cudaMalloc(&g->node_off, offs.size() * sizeof(int)));
    G1_CUDA(cudaMemcpyAsync(g->node_off, offs.data(), offs.size() * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    G1_CUDA(cudaStreamSynchronize(ctx->stream));
  }
  if (host_predef) {
    G1_CUDA(cudaMemcpyAsync(g->predef, host_predef, 2 * (size_t)n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  } else {
    // preDef = dir.preDef(leftX, rangeX, rightX, sigma)  A/OneDGrid.scala:24, on the device (K1)
    std::vector<double> v((size_t)n + 2);
    v[0] = leftX;
    memcpy(v.data() + 1, rangeX, (size_t)n * sizeof(double));
    v[n + 1] = rightX;
    if ((status = gs_ctx_stage(ctx, v.size() * sizeof(double))) != GS_OK) return fail(status);
    G1_CUDA(cudaMemcpyAsync(ctx->stage, v.data(), v.size() * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    G1_CUDA(cudaStreamSynchronize(ctx->stream));
    if ((status = gs_launch_predef(ctx, dir, ctx->stage, n, sigma, g->predef)) != GS_OK) return fail(status);
  }
  // grid = predict = 4.0  (A/OneDGrid.scala:20-21)
  if ((status = gs_launch_fill(ctx, g->grid, (long long)elems, 4.0)) != GS_OK) return fail(status);
  if ((status = gs_launch_fill(ctx, g->predict, (long long)elems, 4.0)) != GS_OK) return fail(status);
  G1_CUDA(cudaStreamSynchronize(ctx->stream));
#undef G1_CUDA
  *out = g;
  return GS_OK;
}

extern "C" int gs_grid1d_destroy(gs_grid1d *g) {
  if (!g) return GS_OK;
  cudaSetDevice(g->ctx->device);
  cudaStreamSynchronize(g->ctx->stream);
  cudaFree(g->grid);
  cudaFree(g->predict);
  cudaFree(g->predef);
  cudaFree(g->node_off);
  cudaFree(g->leftT);
  cudaFree(g->rightT);
  cudaFree(g->scratch);
  cudaFree(g->status);
  delete g;
  return GS_OK;
}

static double *array_of(gs_grid1d *g, int which) {
  // OneDGrid keeps `result` private and equal to `grid` after every step (A/OneDGrid.scala:33)
  return which == GS_PREDICT ? g->predict : g->grid;
}

static const size_t kStageBytes = (size_t)128 << 20;

extern "C" int gs_grid1d_upload(gs_grid1d *g, int which, int layout, const double *host) {
  GS_REQUIRE(g && host, "NULL argument");
  GS_REQUIRE(which == GS_GRID || which == GS_PREDICT || which == GS_RESULT, "unknown array");
  gs_ctx *ctx = g->ctx;
  GS_TRY(gs_ctx_use(ctx));
  double *dst = array_of(g, which);
  if (layout == GS_INTERLEAVED) {
    GS_CUDA(cudaMemcpy2DAsync(dst, (size_t)g->pitch * sizeof(double), host, (size_t)g->nlines * sizeof(double),
                              (size_t)g->nlines * sizeof(double), (size_t)g->n, cudaMemcpyHostToDevice, ctx->stream));
    GS_CUDA(cudaStreamSynchronize(ctx->stream));
    return GS_OK;
  }
  GS_REQUIRE(layout == GS_LINE_MAJOR, "unknown layout");
  size_t line_bytes = (size_t)g->n * sizeof(double);
  long long chunk = (long long)std::max<size_t>(32, kStageBytes / line_bytes / 32 * 32);
  chunk = std::min<long long>(chunk, (g->nlines + 31) / 32 * 32);
  GS_TRY(gs_ctx_stage(ctx, (size_t)chunk * line_bytes));
  for (long long l0 = 0; l0 < g->nlines; l0 += chunk) {
    int L = (int)std::min<long long>(chunk, g->nlines - l0);
    GS_CUDA(cudaMemcpyAsync(ctx->stage, host + (size_t)l0 * g->n, (size_t)L * line_bytes, cudaMemcpyHostToDevice, ctx->stream));
    dim3 blk(32, 8), grd((g->n + 31) / 32, (L + 31) / 32);
    k_lines_to_interleaved<<<grd, blk, 0, ctx->stream>>>(ctx->stage, dst, L, g->n, g->pitch, l0);
    ctx->launches++;
    GS_CUDA(cudaGetLastError());
  }
  GS_CUDA(cudaStreamSynchronize(ctx->stream));
  return GS_OK;
}

extern "C" int gs_grid1d_download(gs_grid1d *g, int which, int layout, double *host) {
  GS_REQUIRE(g && host, "NULL argument");
  GS_REQUIRE(which == GS_GRID || which == GS_PREDICT || which == GS_RESULT, "unknown array");
  gs_ctx *ctx = g->ctx;
  GS_TRY(gs_ctx_use(ctx));
  const double *src = array_of(g, which);
  if (layout == GS_INTERLEAVED) {
    GS_CUDA(cudaMemcpy2DAsync(host, (size_t)g->nlines * sizeof(double), src, (size_t)g->pitch * sizeof(double),
                              (size_t)g->nlines * sizeof(double), (size_t)g->n, cudaMemcpyDeviceToHost, ctx->stream));
    GS_CUDA(cudaStreamSynchronize(ctx->stream));
    return GS_OK;
  }
  GS_REQUIRE(layout == GS_LINE_MAJOR, "unknown layout");
  size_t line_bytes = (size_t)g->n * sizeof(double);
  long long chunk = (long long)std::max<size_t>(32, kStageBytes / line_bytes / 32 * 32);
  chunk = std::min<long long>(chunk, (g->nlines + 31) / 32 * 32);
  GS_TRY(gs_ctx_stage(ctx, (size_t)chunk * line_bytes));
  for (long long l0 = 0; l0 < g->nlines; l0 += chunk) {
    int L = (int)std::min<long long>(chunk, g->nlines - l0);
    dim3 blk(32, 8), grd((g->n + 31) / 32, (L + 31) / 32);
    k_interleaved_to_lines<<<grd, blk, 0, ctx->stream>>>(src, ctx->stage, L, g->n, g->pitch, l0);
    ctx->launches++;
    GS_CUDA(cudaGetLastError());
    GS_CUDA(cudaMemcpyAsync(host + (size_t)l0 * g->n, ctx->stage, (size_t)L * line_bytes, cudaMemcpyDeviceToHost, ctx->stream));
    GS_CUDA(cudaStreamSynchronize(ctx->stream));
  }
  return GS_OK;
}

extern "C" int gs_grid1d_fill(gs_grid1d *g, int which, double value) {
  GS_REQUIRE(g != nullptr, "NULL argument");
  GS_REQUIRE(which == GS_GRID || which == GS_PREDICT || which == GS_RESULT, "unknown array");
  GS_TRY(gs_ctx_use(g->ctx));
  return gs_launch_fill(g->ctx, array_of(g, which), (long long)g->pitch * g->n, value);
}

extern "C" int gs_grid1d_set_bc(gs_grid1d *g, const double *leftT, const double *rightT, int stride) {
  GS_REQUIRE(g && leftT && rightT, "NULL argument");
  GS_REQUIRE(stride == 0 || stride == 1, "stride must be 0 (broadcast) or 1 (per line)");
  gs_ctx *ctx = g->ctx;
  GS_TRY(gs_ctx_use(ctx));
  size_t count = stride ? (size_t)g->nlines : 1;
  if (stride != g->bc_stride || !g->bc_set) {
    GS_CUDA(cudaStreamSynchronize(ctx->stream));
    cudaFree(g->leftT);
    cudaFree(g->rightT);
    g->leftT = g->rightT = nullptr;
    GS_CUDA(cudaMalloc(&g->leftT, count * sizeof(double)));
    GS_CUDA(cudaMalloc(&g->rightT, count * sizeof(double)));
  }
  GS_CUDA(cudaMemcpyAsync(g->leftT, leftT, count * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  GS_CUDA(cudaMemcpyAsync(g->rightT, rightT, count * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  GS_CUDA(cudaStreamSynchronize(ctx->stream));
  g->bc_stride = stride;
  g->bc_set = true;
  return GS_OK;
}

extern "C" int gs_grid1d_step(gs_grid1d *g, double time, int nsteps) {
  GS_REQUIRE(g != nullptr, "NULL argument");
  GS_REQUIRE(nsteps >= 0, "nsteps must be >= 0");
  GS_REQUIRE(g->bc_set, "gs_grid1d_set_bc has not been called");
  if (nsteps == 0) return GS_OK;
  gs_ctx *ctx = g->ctx;
  GS_TRY(gs_ctx_use(ctx));
  GS_TRY(gs_ctx_sync_pool(ctx));

  int block = (int)std::max<int64_t>(32, std::min<int64_t>(256, ctx->opt_block_1d / 32 * 32));
  long long want = (long long)ctx->sm_count * std::max<int64_t>(block, ctx->opt_threads_per_sm_1d);
  long long G = std::min<long long>((g->nlines + block - 1) / block * block, want / block * block);
  size_t need = (size_t)G * (size_t)(g->n - 1);
  if (g->scratch_elems < need) {
    GS_CUDA(cudaStreamSynchronize(ctx->stream));
    cudaFree(g->scratch);
    g->scratch = nullptr;
    g->scratch_elems = 0;
    GS_CUDA(cudaMalloc(&g->scratch, need * sizeof(double2)));
    g->scratch_elems = need;
  }
  Step1dParams p;
  p.n = g->n;
  p.nlines = g->nlines;
  p.pitch = g->pitch;
  p.G = G;
  p.grid = g->grid;
  p.predict = g->predict;
  p.scratch = g->scratch;
  p.predef = g->predef;
  p.pool = ctx->pool_dev;
  p.pool_len = (int)ctx->pool.size();
  p.single_off = g->single_off;
  p.node_off = g->node_off;
  p.leftT = g->leftT;
  p.rightT = g->rightT;
  p.bc_stride = g->bc_stride;
  p.time = time;
  p.nsteps = nsteps;
  p.status = g->status;
  size_t smem = 0;
  const size_t budget = 48 * 1024;
  p.smem_predef = (2 * (size_t)g->n * sizeof(double) <= budget / 2) ? 1 : 0;
  if (p.smem_predef) smem += 2 * (size_t)g->n * sizeof(double);
  p.smem_pool = (smem + (size_t)p.pool_len * sizeof(double) <= budget) ? 1 : 0;
  if (p.smem_pool) smem += (size_t)p.pool_len * sizeof(double);
  int blocks = (int)(G / block);
  if (g->spline_mode == GS_SPLINE_SINGLE) k_step1d<GS_SPLINE_SINGLE><<<blocks, block, smem, ctx->stream>>>(p);
  else k_step1d<GS_SPLINE_PER_NODE><<<blocks, block, smem, ctx->stream>>>(p);
  ctx->launches++;
  GS_CUDA(cudaGetLastError());
  return GS_OK;
}

extern "C" int gs_grid1d_set_solver(gs_grid1d *g, int solver) {
  GS_REQUIRE(g != nullptr, "NULL argument");
  GS_REQUIRE(solver == GS_SOLVER_AUTO || solver == GS_SOLVER_THOMAS || solver == GS_SOLVER_PARTITIONED, "unknown solver");
  g->solver = solver;
  return GS_OK;
}

extern "C" int gs_grid1d_status(gs_grid1d *g, uint32_t *flags) {
  GS_REQUIRE(g && flags, "NULL argument");
  GS_TRY(gs_ctx_use(g->ctx));
  GS_CUDA(cudaMemcpyAsync(flags, g->status, sizeof(unsigned), cudaMemcpyDeviceToHost, g->ctx->stream));
  GS_CUDA(cudaStreamSynchronize(g->ctx->stream));
  return GS_OK;
}
extern "C" int gs_grid1d_clear_status(gs_grid1d *g) {
  GS_REQUIRE(g != nullptr, "NULL argument");
  GS_TRY(gs_ctx_use(g->ctx));
  GS_CUDA(cudaMemsetAsync(g->status, 0, sizeof(unsigned), g->ctx->stream));
  return GS_OK;
}
extern "C" int gs_grid1d_device_ptr(gs_grid1d *g, int which, void **dptr, int64_t *pitch) {
  GS_REQUIRE(g && dptr, "NULL argument");
  GS_REQUIRE(which == GS_GRID || which == GS_PREDICT || which == GS_RESULT, "unknown array");
  *dptr = array_of(g, which);
  if (pitch) *pitch = g->pitch;
  return GS_OK;
}
extern "C" int gs_grid1d_bytes_per_step(gs_grid1d *g, double *bytes) {
  GS_REQUIRE(g && bytes, "NULL argument");
  *bytes = 32.0 * (double)g->nlines * (double)g->n;
  return GS_OK;
}
