// gs_sharded2d.cu -- ONE TwoDGrid on several GPUs of a box, owned by one host process (a JVM): the multi-device form
// of the drop-in boundary.  Rows are sharded: device p owns rows [p R/P, (p+1) R/P) of grid / predict
// / result, a contiguous segment of every column.
//   x sweep  : whole lines on every device, no communication;
//   y sweep  : the partitioned elimination continued across devices (K6_BEGIN / K6_END, gs_grid2d_k6.cuh).  The only
//              exchange of a step -- six doubles per column and device, plus one halo row of `predict` each way -- is
//              done BY THE KERNELS over NVLink peer memory: k6_sweep<K6_BEGIN> stores its condensed segment straight
//              into every peer's `gathered` slot, k6_ends reads the neighbours' halo rows in place.  No collective
//              library, no staging copy: cross-device ordering is two rounds of stream events per step.
// Hazards and their events (s = step, p, q = devices):
//   end_q(s) reads gathered_q[s % 2] written by begin_p(s), all p      -> end waits begin_done[p] of every p
//   begin_p(s) reads predict rows of p +- 1 that end(s - 1) rewrote     -> begin waits end_done of its neighbours
//   begin_p(s + 2) overwrites gathered_q[s % 2] while end_q(s) reads?   -> impossible: end_q(s) < begin_q(s + 1) (same
//     stream) < end_p(s + 1) (event) < begin_p(s + 2) (same stream); `gathered` is double-buffered for that reason.
// Tolerance family (a reordered elimination): see DESIGN.md for what that means on spline tables.
#include <algorithm>
#include <vector>

#include "gs_internal.h"

int gs_grid2d_split_begin_peers(gs_grid2d *g, double time, const double *halo_prev, const double *halo_next,
                                double *const *dsts, int ndst);

struct gs_sharded2d {
  int P = 0, cols = 0, rows = 0;
  std::vector<int> dev, r0, rl;
  std::vector<gs_ctx *> ctx;
  std::vector<gs_grid2d *> g;
  std::vector<double *> gathered[2];
  std::vector<cudaEvent_t> begin_done, end_done;
  bool stepped = false;
  unsigned long long step_no = 0;
  int low_kind = GS_TEMP;
};

#define GS_EACH(s, p) for (int p = 0; p < (s)->P; ++p)

static int use_dev(gs_sharded2d *s, int p) {
  GS_CUDA(cudaSetDevice(s->dev[p]));
  return GS_OK;
}

extern "C" int gs_sharded2d_destroy(gs_sharded2d *s) {
  if (!s) return GS_OK;
  GS_EACH(s, p) {
    cudaSetDevice(s->dev[p]);
    if (p < (int)s->ctx.size() && s->ctx[p]) gs_ctx_synchronize(s->ctx[p]);
  }
  GS_EACH(s, p) {
    cudaSetDevice(s->dev[p]);
    if (p < (int)s->g.size() && s->g[p]) gs_grid2d_destroy(s->g[p]);
    for (int b = 0; b < 2; ++b)
      if (p < (int)s->gathered[b].size()) cudaFree(s->gathered[b][p]);
    if (p < (int)s->begin_done.size()) cudaEventDestroy(s->begin_done[p]);
    if (p < (int)s->end_done.size()) cudaEventDestroy(s->end_done[p]);
    if (p < (int)s->ctx.size() && s->ctx[p]) gs_ctx_destroy(s->ctx[p]);
  }
  delete s;
  return GS_OK;
}

extern "C" int gs_sharded2d_create(int ndev, const int *devices, int cols, int rows, int xdir, int ydir,
                                   const double *xvalues, const double *yvalues, gs_sharded2d **out) {
  GS_REQUIRE(out && xvalues && yvalues, "NULL argument");
  GS_REQUIRE(ndev >= 1 && ndev <= GS_MAX_SHARDS_ABI, "1 .. 16 devices");
  GS_REQUIRE(rows % ndev == 0 && rows / ndev >= 16, "rows must divide evenly over the devices, at least 16 rows each");
  int have = 0;
  if (cudaGetDeviceCount(&have) != cudaSuccess || have < 1) {
    gs_set_error("no CUDA device: gridsplines_b200 has no CPU implementation");
    return GS_ERR_NO_DEVICE;
  }
  gs_sharded2d *s = new gs_sharded2d;
  s->P = ndev; s->cols = cols; s->rows = rows;
  for (int p = 0; p < ndev; ++p) {
    const int d = devices ? devices[p] : p;
    if (d < 0 || d >= have) {
      delete s;
      gs_set_error("device %d does not exist (%d visible)", d, have);
      return GS_ERR_BAD_ARG;
    }
    s->dev.push_back(d);
    s->r0.push_back(p * (rows / ndev));
    s->rl.push_back(rows / ndev);
  }
  // every device maps every other device's memory (NVLink / NVSwitch peer access)
  for (int p = 0; p < ndev; ++p) {
    cudaSetDevice(s->dev[p]);
    for (int q = 0; q < ndev; ++q) {
      if (q == p || s->dev[q] == s->dev[p]) continue;   // (several shards on one device: a test configuration)
      int can = 0;
      cudaDeviceCanAccessPeer(&can, s->dev[p], s->dev[q]);
      if (!can) {
        gs_set_error("device %d cannot access device %d as a peer", s->dev[p], s->dev[q]);
        delete s;
        return GS_ERR_CUDA;
      }
      cudaError_t e = cudaDeviceEnablePeerAccess(s->dev[q], 0);
      if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) {
        gs_set_error("cudaDeviceEnablePeerAccess(%d -> %d): %s", s->dev[p], s->dev[q], cudaGetErrorString(e));
        delete s;
        return GS_ERR_CUDA;
      }
      cudaGetLastError();
    }
  }
  s->ctx.assign(ndev, nullptr);
  s->g.assign(ndev, nullptr);
  for (int b = 0; b < 2; ++b) s->gathered[b].assign(ndev, nullptr);
  s->begin_done.assign(ndev, nullptr);
  s->end_done.assign(ndev, nullptr);
  int rc = GS_OK;
  for (int p = 0; p < ndev && rc == GS_OK; ++p) {
    rc = gs_ctx_create(s->dev[p], nullptr, &s->ctx[p]);
    if (rc != GS_OK) break;
    // the band's coordinates with one neighbour coordinate on each side (Dim.values of the band)
    rc = gs_grid2d_create(s->ctx[p], cols, s->rl[p], xdir, ydir, xvalues, yvalues + s->r0[p], &s->g[p]);
    if (rc != GS_OK) break;
    rc = gs_grid2d_set_solver(s->g[p], GS_SOLVER_PARTITIONED);
    if (rc != GS_OK) break;
    cudaSetDevice(s->dev[p]);
    for (int b = 0; b < 2 && rc == GS_OK; ++b)
      if (cudaMalloc(&s->gathered[b][p], (size_t)ndev * 6 * cols * sizeof(double)) != cudaSuccess) rc = GS_ERR_OOM;
    if (cudaEventCreateWithFlags(&s->begin_done[p], cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&s->end_done[p], cudaEventDisableTiming) != cudaSuccess)
      rc = GS_ERR_CUDA;
  }
  if (rc != GS_OK) {
    gs_sharded2d_destroy(s);
    return rc;
  }
  *out = s;
  return GS_OK;
}

extern "C" int gs_sharded2d_device_count(gs_sharded2d *s, int *ndev) {
  GS_REQUIRE(s && ndev, "NULL argument");
  *ndev = s->P;
  return GS_OK;
}

// the same table in every device's pool: the ids agree because every pool sees the same sequence of calls
extern "C" int gs_sharded2d_spline_create(gs_sharded2d *s, int kind, int npieces, const double *knots, const double *x0,
                                          const double *coefs, double lowerX, double upperX, double lowerY,
                                          double upperY, int *id_out) {
  GS_REQUIRE(s && id_out, "NULL argument");
  int first = -1;
  GS_EACH(s, p) {
    int id = -1;
    GS_TRY(gs_spline_create(s->ctx[p], kind, npieces, knots, x0, coefs, lowerX, upperX, lowerY, upperY, &id));
    if (p == 0) first = id;
    GS_REQUIRE(id == first, "spline pools of the devices went out of step");
  }
  *id_out = first;
  return GS_OK;
}

// ids as for gs_grid2d_set_map on the WHOLE grid; PER_ROW / DENSE maps are cut into the devices' bands (index -1 of a
// band is the row above it)
extern "C" int gs_sharded2d_set_map(gs_sharded2d *s, int selector, int mode, const int *ids) {
  GS_REQUIRE(s && ids, "NULL argument");
  GS_EACH(s, p) {
    const int *local = ids;
    if (mode == GS_MAP_PER_ROW) local = ids + s->r0[p];
    else if (mode == GS_MAP_DENSE) local = ids + (size_t)s->r0[p] * (size_t)(s->cols + 1);
    GS_TRY(gs_grid2d_set_map(s->g[p], selector, mode, local));
  }
  return GS_OK;
}

extern "C" int gs_sharded2d_set_bound(gs_sharded2d *s, int side, int kind, int rep, const double *values, int n) {
  GS_REQUIRE(s && values, "NULL argument");
  GS_REQUIRE(side >= GS_UPP && side <= GS_LEFT, "unknown side");
  if (side == GS_LOW) {
    GS_REQUIRE(!(kind == GS_FLOW && s->P > 1),
               "a lower HeatFlow bound is not supported on a sharded grid (it reads the neighbouring column's last row)");
    s->low_kind = kind;
  }
  if (side == GS_UPP) return gs_grid2d_set_bound(s->g[0], side, kind, rep, values, n);
  if (side == GS_LOW) return gs_grid2d_set_bound(s->g[s->P - 1], side, kind, rep, values, n);
  GS_EACH(s, p) {
    if (n == 1) GS_TRY(gs_grid2d_set_bound(s->g[p], side, kind, rep, values, 1));
    else {
      GS_REQUIRE(n == s->rows, "a left / right bound has one value or one per row");
      GS_TRY(gs_grid2d_set_bound(s->g[p], side, kind, rep, values + s->r0[p], s->rl[p]));
    }
  }
  return GS_OK;
}

extern "C" int gs_sharded2d_fill(gs_sharded2d *s, double value) {
  GS_REQUIRE(s != nullptr, "NULL argument");
  GS_EACH(s, p) GS_TRY(gs_grid2d_fill(s->g[p], value));
  return GS_OK;
}
extern "C" int gs_sharded2d_upload(gs_sharded2d *s, int which, const double *host) {
  GS_REQUIRE(s && host, "NULL argument");
  GS_EACH(s, p) GS_TRY(gs_grid2d_upload(s->g[p], which, host + (size_t)s->r0[p] * s->cols));
  return GS_OK;
}
extern "C" int gs_sharded2d_download(gs_sharded2d *s, int which, double *host) {
  GS_REQUIRE(s && host, "NULL argument");
  GS_EACH(s, p) GS_TRY(gs_grid2d_download(s->g[p], which, host + (size_t)s->r0[p] * s->cols));
  return GS_OK;
}
extern "C" int gs_sharded2d_synchronize(gs_sharded2d *s) {
  GS_REQUIRE(s != nullptr, "NULL argument");
  GS_EACH(s, p) GS_TRY(gs_ctx_synchronize(s->ctx[p]));
  return GS_OK;
}
extern "C" int gs_sharded2d_status(gs_sharded2d *s, uint32_t *flags) {
  GS_REQUIRE(s && flags, "NULL argument");
  *flags = 0;
  GS_EACH(s, p) {
    uint32_t f = 0;
    GS_TRY(gs_grid2d_status(s->g[p], &f));
    *flags |= f;
  }
  return GS_OK;
}
extern "C" int gs_sharded2d_launch_count(gs_sharded2d *s, uint64_t *count) {
  GS_REQUIRE(s && count, "NULL argument");
  *count = 0;
  GS_EACH(s, p) {
    uint64_t c = 0;
    GS_TRY(gs_ctx_launch_count(s->ctx[p], &c));
    *count += c;
  }
  return GS_OK;
}
extern "C" int gs_sharded2d_set_option(gs_sharded2d *s, const char *name, int64_t value) {
  GS_REQUIRE(s && name, "NULL argument");
  GS_EACH(s, p) GS_TRY(gs_ctx_set_option(s->ctx[p], name, value));
  return GS_OK;
}

// nsteps x TwoDGrid.iteration (A/TwoDGrid.scala:254-258) of the whole grid
extern "C" int gs_sharded2d_iteration(gs_sharded2d *s, double time, int nsteps) {
  GS_REQUIRE(s != nullptr && nsteps >= 0, "bad arguments");
  const int P = s->P;
  std::vector<double *> predict(P);
  GS_EACH(s, p) {
    void *ptr = nullptr;
    GS_TRY(gs_grid2d_device_ptr(s->g[p], GS_PREDICT, &ptr));
    predict[p] = static_cast<double *>(ptr);
  }
  std::vector<double *> dsts(P);
  for (int step = 0; step < nsteps; ++step, ++s->step_no) {
    const int buf = (int)(s->step_no & 1ull);
    GS_EACH(s, p) GS_TRY(gs_grid2d_xiter(s->g[p], time));
    GS_EACH(s, p) {
      GS_TRY(use_dev(s, p));
      cudaStream_t st = s->ctx[p]->stream;
      if (s->stepped) {   // the neighbours' predict rows were rewritten by their end kernels of the previous step
        if (p > 0) GS_CUDA(cudaStreamWaitEvent(st, s->end_done[p - 1], 0));
        if (p + 1 < P) GS_CUDA(cudaStreamWaitEvent(st, s->end_done[p + 1], 0));
      }
      const double *prev = p > 0 ? predict[p - 1] + (size_t)(s->rl[p - 1] - 1) * s->cols : nullptr;
      const double *next = p + 1 < P ? predict[p + 1] : nullptr;
      for (int q = 0; q < P; ++q) dsts[q] = s->gathered[buf][q] + (size_t)p * 6 * s->cols;
      GS_TRY(gs_grid2d_split_begin_peers(s->g[p], time, prev, next, dsts.data(), P));
      GS_CUDA(cudaEventRecord(s->begin_done[p], st));
    }
    GS_EACH(s, p) {
      GS_TRY(use_dev(s, p));
      cudaStream_t st = s->ctx[p]->stream;
      for (int q = 0; q < P; ++q)
        if (q != p) GS_CUDA(cudaStreamWaitEvent(st, s->begin_done[q], 0));
      GS_TRY(gs_grid2d_yiter_update_end(s->g[p], time, s->gathered[buf][p], P, p));
      GS_CUDA(cudaEventRecord(s->end_done[p], st));
    }
    s->stepped = true;
  }
  return GS_OK;
}

// nsteps iterations timed ON THE DEVICES: every device's stream is bracketed by two events after a common
// synchronisation point; *ms_max = the longest of the devices' intervals (the whole-job step time x nsteps).
extern "C" int gs_sharded2d_iteration_timed(gs_sharded2d *s, double time, int nsteps, double *ms_max) {
  GS_REQUIRE(s && ms_max && nsteps >= 0, "bad arguments");
  const int P = s->P;
  std::vector<cudaEvent_t> t0(P), t1(P);
  GS_TRY(gs_sharded2d_synchronize(s));
  GS_EACH(s, p) {
    GS_TRY(use_dev(s, p));
    GS_CUDA(cudaEventCreate(&t0[p]));
    GS_CUDA(cudaEventCreate(&t1[p]));
    GS_CUDA(cudaEventRecord(t0[p], s->ctx[p]->stream));
  }
  int rc = gs_sharded2d_iteration(s, time, nsteps);
  GS_EACH(s, p) {
    use_dev(s, p);
    cudaEventRecord(t1[p], s->ctx[p]->stream);
  }
  if (rc == GS_OK) rc = gs_sharded2d_synchronize(s);
  *ms_max = 0.0;
  GS_EACH(s, p) {
    use_dev(s, p);
    float ms = 0.f;
    if (rc == GS_OK && cudaEventElapsedTime(&ms, t0[p], t1[p]) == cudaSuccess) *ms_max = std::max(*ms_max, (double)ms);
    cudaEventDestroy(t0[p]);
    cudaEventDestroy(t1[p]);
  }
  return rc;
}
