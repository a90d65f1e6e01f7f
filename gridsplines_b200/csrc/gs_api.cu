// gs_api.cu -- context, spline pool, metric coefficients (K1) and the small evaluation
// kernels behind include/gs_b200.h.  Host code here only marshals; all arithmetic of the
// hot path runs on the device.
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include "gs_device.cuh"
#include "gs_internal.h"

static thread_local char g_err[512] = "";

void gs_set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" int gs_version(void) { return GS_B200_VERSION; }
extern "C" const char *gs_last_error(void) { return g_err; }

// ---------------------------------------------------------------------------------
// context
// ---------------------------------------------------------------------------------
extern "C" int gs_ctx_create(int device, void *stream, gs_ctx **out) {
  GS_REQUIRE(out != nullptr, "out is NULL");
  *out = nullptr;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0) {
    gs_set_error("no CUDA device: %s (libgs_b200 has no CPU path)",
                 e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    return GS_ERR_NO_DEVICE;
  }
  GS_REQUIRE(device >= 0 && device < count, "device index out of range");
  GS_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  GS_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major < 10) {
    gs_set_error("device %d is sm_%d%d; libgs_b200 carries sm_100a code only", device, prop.major,
                 prop.minor);
    return GS_ERR_NO_DEVICE;
  }
  gs_ctx *c = new gs_ctx();
  c->device = device;
  c->sm_count = prop.multiProcessorCount;
  c->smem_optin = prop.sharedMemPerBlockOptin;
  if (stream) {
    c->stream = (cudaStream_t)stream;
  } else {
    e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) {
      delete c;
      gs_set_error("cudaStreamCreate failed: %s", cudaGetErrorString(e));
      return GS_ERR_CUDA;
    }
    c->own_stream = true;
  }
  *out = c;
  return GS_OK;
}

static void ctx_free(gs_ctx *ctx) {
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  if (ctx->pool_dev) cudaFree(ctx->pool_dev);
  if (ctx->stage) cudaFree(ctx->stage);
  if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
}
// Grids keep their context alive: hosts with garbage collectors (the JVM, CPython) may finalise the context
// object before the grids made from it.
extern "C" int gs_ctx_destroy(gs_ctx *ctx) {
  if (!ctx) return GS_OK;
  if (ctx->refs > 0) {
    ctx->dead = true;
    return GS_OK;
  }
  ctx_free(ctx);
  return GS_OK;
}
void gs_ctx_retain(gs_ctx *ctx) { ctx->refs++; }
void gs_ctx_release(gs_ctx *ctx) {
  if (--ctx->refs <= 0 && ctx->dead) ctx_free(ctx);
}

int gs_ctx_use(gs_ctx *ctx) {
  GS_REQUIRE(ctx != nullptr, "ctx is NULL");
  GS_CUDA(cudaSetDevice(ctx->device));
  return GS_OK;
}

extern "C" int gs_ctx_synchronize(gs_ctx *ctx) {
  GS_TRY(gs_ctx_use(ctx));
  GS_CUDA(cudaStreamSynchronize(ctx->stream));
  return GS_OK;
}

extern "C" int gs_ctx_launch_count(gs_ctx *ctx, uint64_t *count) {
  GS_REQUIRE(ctx && count, "NULL argument");
  *count = ctx->launches;
  return GS_OK;
}

extern "C" int gs_ctx_set_option(gs_ctx *ctx, const char *name, int64_t value) {
  GS_REQUIRE(ctx && name, "NULL argument");
  if (!strcmp(name, "threads_per_sm_1d")) ctx->opt_threads_per_sm_1d = value;
  else if (!strcmp(name, "block_1d")) ctx->opt_block_1d = value;
  else if (!strcmp(name, "native_div")) ctx->opt_native_div = value;
  else if (!strcmp(name, "const_hoist")) ctx->opt_const_hoist = value;
  else if (!strcmp(name, "pipeline")) ctx->opt_pipeline = value;
  else if (!strcmp(name, "pipe_cfg")) ctx->opt_pipe_cfg = value;
  else if (!strcmp(name, "seg_len")) ctx->opt_seg_len = value;
  else if (!strcmp(name, "hoist_ck")) ctx->opt_hoist_ck = value;
  else if (!strcmp(name, "k2r")) ctx->opt_k2r = value;
  else if (!strcmp(name, "k5")) ctx->opt_k5 = value;
  else if (!strcmp(name, "k5_xstrip")) ctx->opt_k5_xstrip = value;
  else if (!strcmp(name, "k5_cluster")) ctx->opt_k5_cluster = value != 0;
  else if (!strcmp(name, "k5_nb")) ctx->opt_k5_nb = (int)value;
  else if (!strcmp(name, "k5t")) ctx->opt_k5t = (int)value;   // 0 never, 1 from 2^20 nodes per sweep, 2 always
  else if (!strcmp(name, "k5t_sx")) ctx->opt_k5t_sx = (int)value;
  else if (!strcmp(name, "k5t_sy")) ctx->opt_k5t_sy = (int)value;
  else if (!strcmp(name, "k5t_s1")) ctx->opt_k5t_s1 = (int)value;
  else if (!strcmp(name, "k5t_nt")) ctx->opt_k5t_nt = (int)value;
  else if (!strcmp(name, "k5t_promo")) ctx->opt_k5t_promo = (int)value;
  else if (!strcmp(name, "k5t_wide")) ctx->opt_k5t_wide = (int)value;
  else if (!strcmp(name, "k5t_ystrip")) ctx->opt_k5t_ystrip = (int)value;
  else if (!strcmp(name, "k5t_soft")) ctx->opt_k5t_soft = value != 0;
  else if (!strcmp(name, "k5t_pdl")) ctx->opt_k5t_pdl = value != 0;
  else if (!strcmp(name, "k5t_hints")) ctx->opt_k5t_hints = (int)value;
  else if (!strcmp(name, "k6")) ctx->opt_k6 = value;
  else if (!strcmp(name, "faces1d")) ctx->opt_faces1d = value;
  else if (!strcmp(name, "k2t")) ctx->opt_k2t = value;
  else if (!strcmp(name, "k2t_cps")) ctx->opt_k2t_cps = value;
  else if (!strcmp(name, "k2t_ctas")) ctx->opt_k2t_ctas = value;
  else if (!strcmp(name, "k2t_hints")) ctx->opt_k2t_hints = value;
  else if (!strcmp(name, "k6_chunk")) ctx->opt_k6_chunk = value;   // accepted for compatibility; chunks are 8 nodes
  else if (!strcmp(name, "k6_cluster")) ctx->opt_k6_cluster = value != 0;
  else if (!strcmp(name, "k6_fuse")) ctx->opt_k6_fuse = value != 0;
  else if (!strcmp(name, "k7")) ctx->opt_k7 = value;
  else if (!strcmp(name, "k7_fuse")) ctx->opt_k7_fuse = value;
  else if (!strcmp(name, "k7_nsb")) ctx->opt_k7_nsb = value;
  else if (!strcmp(name, "k7f_nfw")) ctx->opt_k7f_nfw = value;
  else if (!strcmp(name, "k7_promo")) ctx->opt_k7_promo = value;
  else if (!strcmp(name, "auto_exact")) ctx->opt_auto_exact = value;
  else if (!strcmp(name, "host_chunk_mb")) ctx->opt_host_chunk_mb = value;
  else if (!strcmp(name, "ck_block")) ctx->opt_ck_block = value;
  else {
    gs_set_error("unknown option '%s'", name);
    return GS_ERR_BAD_ARG;
  }
  return GS_OK;
}

int gs_ctx_stage(gs_ctx *ctx, size_t bytes) {
  if (ctx->stage_bytes >= bytes) return GS_OK;
  if (ctx->stage) {
    GS_CUDA(cudaStreamSynchronize(ctx->stream));
    GS_CUDA(cudaFree(ctx->stage));
    ctx->stage = nullptr;
    ctx->stage_bytes = 0;
  }
  GS_CUDA(cudaMalloc(&ctx->stage, bytes));
  ctx->stage_bytes = bytes;
  return GS_OK;
}

int gs_ctx_sync_pool(gs_ctx *ctx) {
  if (!ctx->pool_dirty) return GS_OK;
  size_t len = ctx->pool.size();
  if (len == 0) return GS_OK;
  if (ctx->pool_dev_len < len) {
    if (ctx->pool_dev) {
      GS_CUDA(cudaStreamSynchronize(ctx->stream));
      GS_CUDA(cudaFree(ctx->pool_dev));
      ctx->pool_dev = nullptr;
    }
    size_t cap = len * 2 + 64;
    GS_CUDA(cudaMalloc(&ctx->pool_dev, cap * sizeof(double)));
    ctx->pool_dev_len = cap;
  }
  GS_CUDA(cudaMemcpyAsync(ctx->pool_dev, ctx->pool.data(), len * sizeof(double),
                          cudaMemcpyHostToDevice, ctx->stream));
  GS_CUDA(cudaStreamSynchronize(ctx->stream));
  ctx->pool_dirty = false;
  return GS_OK;
}

// ---------------------------------------------------------------------------------
// spline pool
// ---------------------------------------------------------------------------------
extern "C" int gs_spline_create(gs_ctx *ctx, int kind, int npieces, const double *knots,
                                const double *x0, const double *coefs, double lowerX, double upperX,
                                double lowerY, double upperY, int *id_out) {
  GS_REQUIRE(ctx && knots && coefs && id_out, "NULL argument");
  GS_REQUIRE(kind == GS_CONST || kind == GS_LINE || kind == GS_CUBIC, "unknown piece kind");
  GS_REQUIRE(npieces >= 1, "npieces must be >= 1");
  for (int i = 0; i < npieces; ++i)
    GS_REQUIRE(knots[i] <= knots[i + 1], "knots must be ascending");
  int off = (int)ctx->pool.size();
  std::vector<double> &p = ctx->pool;
  p.resize(off + GS_HDR + (npieces + 1) + (size_t)npieces * GS_PIECE);
  double *h = p.data() + off;
  h[0] = (double)kind;
  h[1] = (double)npieces;
  h[2] = lowerX;
  h[3] = upperX;
  h[4] = lowerY;
  h[5] = upperY;
  h[6] = h[7] = 0.0;
  if (npieces >= 2) {
    // equally spaced knots (exactly representable steps): gs_knot_floor may guess the piece arithmetically
    volatile double step = knots[1] - knots[0];
    bool uniform = step > 0.0;
    for (int i = 0; i <= npieces && uniform; ++i) {
      volatile double expect = knots[0] + (double)i * step;
      uniform = expect == knots[i];
    }
    if (uniform) h[6] = 1.0 / step;
  }
  double *k = h + GS_HDR;
  memcpy(k, knots, (size_t)(npieces + 1) * sizeof(double));
  double *pc = k + (npieces + 1);
  for (int i = 0; i < npieces; ++i) {
    const double *c = coefs + 4 * (size_t)i;
    double *q = pc + (size_t)i * GS_PIECE;
    q[0] = x0 ? x0[i] : 0.0;
    q[1] = c[0];
    q[2] = c[1];
    q[3] = c[2];
    q[4] = c[3];
    // cubicHornerIntegral's coefficient expressions, P/PieceFunction.scala:194-198 (IEEE ops;
    // this file's host code is compiled with -ffp-contract=off)
    volatile double a2 = 0.5 * c[1], a3 = c[2] / 3.0, a4 = 0.25 * c[3];
    q[5] = a2;
    q[6] = a3;
    q[7] = a4;
  }
  ctx->spline_off.push_back(off);
  ctx->pool_dirty = true;
  *id_out = (int)ctx->spline_off.size() - 1;
  return GS_OK;
}

__global__ void k_eval_apply(const double *pool, int off, int n, const double *x, double *out,
                             unsigned *status) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  GsSpline s = gs_spline_view(pool, off);
  unsigned flags = 0;
  out[i] = gs_ads_apply(s, x[i], flags);
  gs_raise(status, flags);
}
__global__ void k_eval_average(const double *pool, int off, int n, const double *lo, const double *hi,
                               double *out, unsigned *status) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  GsSpline s = gs_spline_view(pool, off);
  unsigned flags = 0;
  bool ok = true;
  out[i] = gs_ads_average<false>(s, lo[i], hi[i], flags, ok);
  gs_raise(status, flags);
}

static int eval_common(gs_ctx *ctx, int id, int n, const double *a, const double *b, double *out,
                       bool average) {
  GS_TRY(gs_ctx_use(ctx));
  GS_REQUIRE(id >= 0 && id < (int)ctx->spline_off.size(), "spline id out of range");
  GS_REQUIRE(n >= 0 && a && out && (!average || b), "NULL argument");
  if (n == 0) return GS_OK;
  GS_TRY(gs_ctx_sync_pool(ctx));
  size_t bytes = (size_t)n * sizeof(double);
  GS_TRY(gs_ctx_stage(ctx, 3 * bytes + 256));
  double *da = ctx->stage, *db = da + n, *dout = db + n;
  unsigned *dstatus = (unsigned *)(dout + n);
  GS_CUDA(cudaMemsetAsync(dstatus, 0, sizeof(unsigned), ctx->stream));
  GS_CUDA(cudaMemcpyAsync(da, a, bytes, cudaMemcpyHostToDevice, ctx->stream));
  if (average) GS_CUDA(cudaMemcpyAsync(db, b, bytes, cudaMemcpyHostToDevice, ctx->stream));
  int off = ctx->spline_off[id];
  int blocks = (n + 127) / 128;
  if (average) k_eval_average<<<blocks, 128, 0, ctx->stream>>>(ctx->pool_dev, off, n, da, db, dout, dstatus);
  else k_eval_apply<<<blocks, 128, 0, ctx->stream>>>(ctx->pool_dev, off, n, da, dout, dstatus);
  ctx->launches++;
  GS_CUDA(cudaGetLastError());
  GS_CUDA(cudaMemcpyAsync(out, dout, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  unsigned hstatus = 0;
  GS_CUDA(cudaMemcpyAsync(&hstatus, dstatus, sizeof(unsigned), cudaMemcpyDeviceToHost, ctx->stream));
  GS_CUDA(cudaStreamSynchronize(ctx->stream));
  ctx->eval_flags = hstatus;
  return GS_OK;
}

// GS_FLAG_* raised by the last gs_spline_eval_* call of this context (the reference throws where these are set:
// RuntimeException for an undefined argument, P/intervaltree/AbsITree.scala:475-479)
extern "C" int gs_spline_eval_status(gs_ctx *ctx, uint32_t *flags) {
  GS_REQUIRE(ctx && flags, "NULL argument");
  *flags = ctx->eval_flags;
  return GS_OK;
}
extern "C" int gs_spline_eval_apply(gs_ctx *ctx, int id, int n, const double *x, double *out) {
  return eval_common(ctx, id, n, x, nullptr, out, false);
}
extern "C" int gs_spline_eval_average(gs_ctx *ctx, int id, int n, const double *lo, const double *hi,
                                      double *out) {
  return eval_common(ctx, id, n, lo, hi, out, true);
}

// ---------------------------------------------------------------------------------
// K1: metric coefficients
// ---------------------------------------------------------------------------------
GS_DEV double k1_ah(double l1, double l2, double l3) { return (l3 + l2) / 2.0 - (l1 + l2) / 2.0; }  // A/arraygrid/package.scala:17-19
GS_DEV double k1_dh(double l1, double l2) { return l2 - l1; }                                       // :98-100
GS_DEV double k1_rh(double r1, double r2) { return (r1 + r2) / 2.0; }                                // :185-187
// :196-198; math.pow(x, 2.0) evaluated as x*x
GS_DEV double k1_vol(double r1, double r2, double r3) {
  double u = k1_rh(r2, r3), l = k1_rh(r1, r2);
  return (u * u - l * l) / 2.0;
}

// values = low ++ range ++ upp (n + 2); out[2i], out[2i+1] = a_i, c_i
__global__ void k_predef(int dir, const double *values, int n, double sigma, double *out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double l1 = values[i], l2 = values[i + 1], l3 = values[i + 2];
  if (dir == GS_RADIAL) {  // makeRadialMatrix A/arraygrid/package.scala:120-176
    double volume = k1_vol(l1, l2, l3);
    out[2 * i] = sigma / volume * k1_rh(l1, l2) / k1_dh(l1, l2);
    out[2 * i + 1] = sigma / volume * k1_rh(l2, l3) / k1_dh(l2, l3);
  } else {                 // makeOrthogonalMatrix :21-78
    double average = k1_ah(l1, l2, l3);
    out[2 * i] = sigma / average / k1_dh(l1, l2);
    out[2 * i + 1] = sigma / average / k1_dh(l2, l3);
  }
}
// A/TypeDir.scala:31-44 (Ortho), :65-79 (Radial); consumed at A/Dim.scala:16-17
__global__ void k_heatflow_geometry(int dir, const double *values, int n, double *o) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  double low = values[0], x0 = values[1], x1 = values[2];
  double xm = values[n - 1], xl = values[n], upp = values[n + 1];
  if (dir == GS_RADIAL) {
    o[0] = k1_vol(low, x0, x1);
    o[1] = k1_rh(x0, x1) / k1_dh(x0, x1);
    o[2] = k1_vol(xm, xl, upp);
    o[3] = k1_rh(xm, xl) / k1_dh(xm, xl);
  } else {
    o[0] = k1_ah(low, x0, x1);
    o[1] = 1.0 / k1_dh(x0, x1);
    o[2] = k1_ah(xm, xl, upp);
    o[3] = 1.0 / k1_dh(xm, xl);
  }
}

// device-side helper used by the grid constructors: values_dev [n+2] -> coefs_dev [2n]
int gs_launch_predef(gs_ctx *ctx, int dir, const double *values_dev, int n, double sigma,
                     double *coefs_dev) {
  k_predef<<<(n + 127) / 128, 128, 0, ctx->stream>>>(dir, values_dev, n, sigma, coefs_dev);
  ctx->launches++;
  GS_CUDA(cudaGetLastError());
  return GS_OK;
}
int gs_launch_heatflow(gs_ctx *ctx, int dir, const double *values_dev, int n, double *out4_dev) {
  k_heatflow_geometry<<<1, 32, 0, ctx->stream>>>(dir, values_dev, n, out4_dev);
  ctx->launches++;
  GS_CUDA(cudaGetLastError());
  return GS_OK;
}

static int stage_values(gs_ctx *ctx, double low, const double *range, int n, double upp,
                        double **values_dev, double **out_dev, size_t out_doubles) {
  GS_TRY(gs_ctx_stage(ctx, ((size_t)n + 2 + out_doubles) * sizeof(double)));
  std::vector<double> v((size_t)n + 2);
  v[0] = low;
  memcpy(v.data() + 1, range, (size_t)n * sizeof(double));
  v[n + 1] = upp;
  GS_CUDA(cudaMemcpyAsync(ctx->stage, v.data(), v.size() * sizeof(double), cudaMemcpyHostToDevice,
                          ctx->stream));
  GS_CUDA(cudaStreamSynchronize(ctx->stream));
  *values_dev = ctx->stage;
  *out_dev = ctx->stage + n + 2;
  return GS_OK;
}

extern "C" int gs_predef(gs_ctx *ctx, int dir, double low, const double *range, int n, double upp,
                         double sigma, double *out_2n) {
  GS_TRY(gs_ctx_use(ctx));
  GS_REQUIRE(range && out_2n && n >= 1, "bad range");
  GS_REQUIRE(dir == GS_ORTHO || dir == GS_RADIAL, "unknown direction kind");
  double *vd, *od;
  GS_TRY(stage_values(ctx, low, range, n, upp, &vd, &od, 2 * (size_t)n));
  GS_TRY(gs_launch_predef(ctx, dir, vd, n, sigma, od));
  GS_CUDA(cudaMemcpyAsync(out_2n, od, 2 * (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  GS_CUDA(cudaStreamSynchronize(ctx->stream));
  return GS_OK;
}

extern "C" int gs_heatflow_geometry(gs_ctx *ctx, int dir, double low, const double *range, int n,
                                    double upp, double *out4) {
  GS_TRY(gs_ctx_use(ctx));
  GS_REQUIRE(range && out4 && n >= 2, "heat-flow geometry needs at least 2 nodes (A/Dim.scala:16)");
  GS_REQUIRE(dir == GS_ORTHO || dir == GS_RADIAL, "unknown direction kind");
  double *vd, *od;
  GS_TRY(stage_values(ctx, low, range, n, upp, &vd, &od, 4));
  GS_TRY(gs_launch_heatflow(ctx, dir, vd, n, od));
  GS_CUDA(cudaMemcpyAsync(out4, od, 4 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  GS_CUDA(cudaStreamSynchronize(ctx->stream));
  return GS_OK;
}
