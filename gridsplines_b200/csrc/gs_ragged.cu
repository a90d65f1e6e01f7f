// gs_ragged.cu -- ragged batch of lines inside one flattened array: the fixed version of the
// reference's experimental passion.iteration overload 3 (A/passion/package.scala:141-227).
//
// The reference's overload walks `lengths.length` lines whose nodes sit at indices(flatten) of one flattened matrix,
// with per-node preDef / conductivity pairs and one (upper, lower) Dirichlet pair per line, and writes `result` at the
// same indices.  Upstream it is unused and has two slips: the last node reads predicted(local)
// instead of predicted(global) (:204), and the right neighbour is taken as predicted(global + 1), which is the next
// node of the line only when lines are rows.  The semantics here are the evident ones: every line is exactly
// passion.iteration overload 1 (:52-95) applied to the line's gathered nodes -- neighbours are the line's previous /
// next entries of `indices` -- and that is what the tests check against (the oracle's overload 1 on gathered data).
// One thread per line, the reference's operation order, IEEE division: bit-exact.  Gathers are as coalesced as the
// index pattern allows; this entry point is about coverage of the reference's one batched surface, not a roofline.
#include <algorithm>

#include "gs_device.cuh"
#include "gs_internal.h"

struct gs_ragged {
  gs_ctx *ctx = nullptr;
  long long matrix_size = 0, total = 0;   // flattened matrix length, sum of the line lengths
  int nlines = 0, spline_mode = GS_SPLINE_SINGLE, single_id = 0;
  long long *start = nullptr;             // [nlines + 1] prefix sums of lengths (device)
  int *indices = nullptr;                 // [total]
  double *predef = nullptr;               // [total][2]
  int *node_off = nullptr;                // [total][2] spline pool offsets (per-node mode)
  double *grid = nullptr, *predict = nullptr, *result = nullptr;   // [matrix_size]
  double2 *scratch = nullptr;             // [total] (delta, lambda)
  double *bc = nullptr;                   // [2][nlines] upper, lower
  unsigned *status = nullptr;
};

struct RaggedParams {
  int nlines, single;
  const long long *start;
  const int *indices, *node_off;
  const double *predef, *grid, *predict, *upper, *lower, *pool;
  double *result;
  double2 *scratch;
  int single_off;
  double time;
  unsigned *status;
};

__global__ void __launch_bounds__(128) k_ragged(RaggedParams p) {
  const int line = blockIdx.x * blockDim.x + threadIdx.x;
  if (line >= p.nlines) return;
  const long long s0 = p.start[line];
  const int len = (int)(p.start[line + 1] - s0);
  const int *idx = p.indices + s0;
  const double *pd = p.predef + 2 * s0;
  const int *off = p.node_off ? p.node_off + 2 * s0 : nullptr;
  double2 *sc = p.scratch + s0;
  const double inv_time = 1.0 / p.time;
  unsigned flags = 0;
  bool ok = true;
  auto z = [&](int k, int side) { return gs_spline_view(p.pool, p.single ? p.single_off : off[2 * k + side]); };
  double delta, lambda;
  double t1 = p.upper[line], t2 = p.predict[idx[0]], t3 = p.predict[idx[1]];
  double a = gs_ads_apply(z(0, 0), (t1 + t2) / 2.0, flags) * pd[0];          // cons :21-25
  double c = gs_ads_apply(z(0, 1), (t2 + t3) / 2.0, flags) * pd[1];
  double vect = -p.grid[idx[0]] / p.time - a * t1;
  gs_forward_first<false>(-(a + c) - inv_time, c, vect, delta, lambda, ok);
  sc[0] = make_double2(delta, lambda);
  int k = 1;
  for (; k < len - 1; ++k) {
    t1 = t2; t2 = t3; t3 = p.predict[idx[k + 1]];
    a = gs_ads_apply(z(k, 0), (t1 + t2) / 2.0, flags) * pd[2 * k];
    c = gs_ads_apply(z(k, 1), (t2 + t3) / 2.0, flags) * pd[2 * k + 1];
    vect = -p.grid[idx[k]] / p.time;
    gs_forward_unit<false>(a, -(a + c) - inv_time, c, vect, delta, lambda, flags, ok);
    sc[k] = make_double2(delta, lambda);
  }
  t1 = t2; t2 = t3; t3 = p.lower[line];
  a = gs_ads_apply(z(k, 0), (t1 + t2) / 2.0, flags) * pd[2 * k];
  c = gs_ads_apply(z(k, 1), (t2 + t3) / 2.0, flags) * pd[2 * k + 1];
  vect = -p.grid[idx[k]] / p.time - c * t3;
  gs_forward_last<false>(a, -(a + c) - inv_time, vect, delta, lambda, ok);
  // backwardPassion :345-365, scattered to the line's indices
  double u = 0.0;
  for (k = len - 1; k >= 0; --k) {
    if (k != len - 1) { const double2 v = sc[k]; delta = v.x; lambda = v.y; }
    u = delta * u + lambda;
    p.result[idx[k]] = u;
  }
  if (!(fabs(u) <= 1.7976931348623157e308)) flags |= GS_FLAG_NONFINITE;
  gs_raise(p.status, flags);
}

static void ragged_free(gs_ragged *g) {
  cudaFree(g->start); cudaFree(g->indices); cudaFree(g->predef); cudaFree(g->node_off);
  cudaFree(g->grid); cudaFree(g->predict); cudaFree(g->result); cudaFree(g->scratch); cudaFree(g->bc); cudaFree(g->status);
  gs_ctx *ctx = g->ctx;
  delete g;
  gs_ctx_release(ctx);
}

extern "C" int gs_ragged_create(gs_ctx *ctx, int64_t matrix_size, int nlines, const int *lengths, const int *indices,
                                const double *predef, int spline_mode, const int *spline_ids, gs_ragged **out) {
  GS_REQUIRE(ctx && lengths && indices && predef && spline_ids && out, "NULL argument");
  GS_REQUIRE(matrix_size > 0 && nlines > 0, "empty problem");
  GS_REQUIRE(spline_mode == GS_SPLINE_SINGLE || spline_mode == GS_SPLINE_PER_NODE, "unknown spline mode");
  GS_TRY(gs_ctx_use(ctx));
  std::vector<long long> start((size_t)nlines + 1, 0);
  for (int l = 0; l < nlines; ++l) {
    GS_REQUIRE(lengths[l] >= 2, "every line needs at least two nodes (a first and a last row of the system)");
    start[l + 1] = start[l] + lengths[l];
  }
  const long long total = start[nlines];
  for (long long f = 0; f < total; ++f)
    GS_REQUIRE(indices[f] >= 0 && indices[f] < matrix_size, "index outside the flattened matrix");
  const int nspl = (int)ctx->spline_off.size();
  const long long nids = spline_mode == GS_SPLINE_SINGLE ? 1 : 2 * total;
  for (long long f = 0; f < nids; ++f) GS_REQUIRE(spline_ids[f] >= 0 && spline_ids[f] < nspl, "unknown spline id");
  gs_ragged *g = new gs_ragged();
  g->ctx = ctx;
  gs_ctx_retain(ctx);
  g->matrix_size = matrix_size; g->total = total; g->nlines = nlines; g->spline_mode = spline_mode;
  g->single_id = spline_ids[0];
#define RG_CUDA(expr) do { cudaError_t e_ = (expr); if (e_ != cudaSuccess) { gs_set_error("%s failed: %s", #expr, cudaGetErrorString(e_)); ragged_free(g); return e_ == cudaErrorMemoryAllocation ? GS_ERR_OOM : GS_ERR_CUDA; } } while (0)
  RG_CUDA(cudaMalloc(&g->start, ((size_t)nlines + 1) * sizeof(long long)));
  RG_CUDA(cudaMalloc(&g->indices, (size_t)total * sizeof(int)));
  RG_CUDA(cudaMalloc(&g->predef, 2 * (size_t)total * sizeof(double)));
  RG_CUDA(cudaMalloc(&g->grid, (size_t)matrix_size * sizeof(double)));
  RG_CUDA(cudaMalloc(&g->predict, (size_t)matrix_size * sizeof(double)));
  RG_CUDA(cudaMalloc(&g->result, (size_t)matrix_size * sizeof(double)));
  RG_CUDA(cudaMalloc(&g->scratch, (size_t)total * sizeof(double2)));
  RG_CUDA(cudaMalloc(&g->bc, 2 * (size_t)nlines * sizeof(double)));
  RG_CUDA(cudaMalloc(&g->status, sizeof(unsigned)));
  RG_CUDA(cudaMemsetAsync(g->status, 0, sizeof(unsigned), ctx->stream));
  RG_CUDA(cudaMemsetAsync(g->grid, 0, (size_t)matrix_size * sizeof(double), ctx->stream));
  RG_CUDA(cudaMemsetAsync(g->predict, 0, (size_t)matrix_size * sizeof(double), ctx->stream));
  RG_CUDA(cudaMemsetAsync(g->result, 0, (size_t)matrix_size * sizeof(double), ctx->stream));
  RG_CUDA(cudaMemcpyAsync(g->start, start.data(), start.size() * sizeof(long long), cudaMemcpyHostToDevice, ctx->stream));
  RG_CUDA(cudaMemcpyAsync(g->indices, indices, (size_t)total * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
  RG_CUDA(cudaMemcpyAsync(g->predef, predef, 2 * (size_t)total * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  std::vector<int> offs;
  if (spline_mode == GS_SPLINE_PER_NODE) {
    offs.resize(2 * (size_t)total);
    for (long long f = 0; f < 2 * total; ++f) offs[f] = ctx->spline_off[spline_ids[f]];
    RG_CUDA(cudaMalloc(&g->node_off, offs.size() * sizeof(int)));
    RG_CUDA(cudaMemcpyAsync(g->node_off, offs.data(), offs.size() * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
  }
  RG_CUDA(cudaStreamSynchronize(ctx->stream));   // the host vectors above go out of scope
#undef RG_CUDA
  *out = g;
  return GS_OK;
}

extern "C" int gs_ragged_destroy(gs_ragged *g) {
  if (!g) return GS_OK;
  GS_TRY(gs_ctx_use(g->ctx));
  GS_CUDA(cudaStreamSynchronize(g->ctx->stream));
  ragged_free(g);
  return GS_OK;
}

static double *ragged_array(gs_ragged *g, int which) { return which == GS_GRID ? g->grid : which == GS_PREDICT ? g->predict : g->result; }

extern "C" int gs_ragged_upload(gs_ragged *g, int which, const double *host) {
  GS_REQUIRE(g && host, "NULL argument");
  GS_REQUIRE(which >= GS_GRID && which <= GS_RESULT, "unknown array");
  GS_TRY(gs_ctx_use(g->ctx));
  GS_CUDA(cudaMemcpyAsync(ragged_array(g, which), host, (size_t)g->matrix_size * sizeof(double), cudaMemcpyHostToDevice, g->ctx->stream));
  GS_CUDA(cudaStreamSynchronize(g->ctx->stream));
  return GS_OK;
}
extern "C" int gs_ragged_download(gs_ragged *g, int which, double *host) {
  GS_REQUIRE(g && host, "NULL argument");
  GS_REQUIRE(which >= GS_GRID && which <= GS_RESULT, "unknown array");
  GS_TRY(gs_ctx_use(g->ctx));
  GS_CUDA(cudaMemcpyAsync(host, ragged_array(g, which), (size_t)g->matrix_size * sizeof(double), cudaMemcpyDeviceToHost, g->ctx->stream));
  GS_CUDA(cudaStreamSynchronize(g->ctx->stream));
  return GS_OK;
}

extern "C" int gs_ragged_iteration(gs_ragged *g, double time, const double *upper, const double *lower) {
  GS_REQUIRE(g && upper && lower, "NULL argument");
  gs_ctx *ctx = g->ctx;
  GS_TRY(gs_ctx_use(ctx));
  GS_TRY(gs_ctx_sync_pool(ctx));
  GS_CUDA(cudaMemcpyAsync(g->bc, upper, (size_t)g->nlines * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  GS_CUDA(cudaMemcpyAsync(g->bc + g->nlines, lower, (size_t)g->nlines * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  RaggedParams p;
  p.nlines = g->nlines; p.single = g->spline_mode == GS_SPLINE_SINGLE;
  p.start = g->start; p.indices = g->indices; p.node_off = g->node_off;
  p.predef = g->predef; p.grid = g->grid; p.predict = g->predict;
  p.upper = g->bc; p.lower = g->bc + g->nlines; p.pool = ctx->pool_dev;
  p.result = g->result; p.scratch = g->scratch;
  p.single_off = ctx->spline_off[g->single_id];
  p.time = time; p.status = g->status;
  k_ragged<<<(g->nlines + 127) / 128, 128, 0, ctx->stream>>>(p);
  ctx->launches++;
  GS_CUDA(cudaGetLastError());
  GS_CUDA(cudaStreamSynchronize(ctx->stream));   // `upper` / `lower` are the caller's
  return GS_OK;
}

extern "C" int gs_ragged_status(gs_ragged *g, uint32_t *flags) {
  GS_REQUIRE(g && flags, "NULL argument");
  GS_TRY(gs_ctx_use(g->ctx));
  GS_CUDA(cudaMemcpyAsync(flags, g->status, sizeof(unsigned), cudaMemcpyDeviceToHost, g->ctx->stream));
  GS_CUDA(cudaStreamSynchronize(g->ctx->stream));
  return GS_OK;
}
