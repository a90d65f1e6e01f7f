// gs_internal.h -- host-side structures behind the opaque handles of include/gs_b200.h.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "../../include/gs_b200.h"

void gs_set_error(const char *fmt, ...);

#define GS_CUDA(expr)                                                                       \
  do {                                                                                      \
    cudaError_t e__ = (expr);                                                               \
    if (e__ != cudaSuccess) {                                                               \
      gs_set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
      return (e__ == cudaErrorMemoryAllocation) ? GS_ERR_OOM : GS_ERR_CUDA;                 \
    }                                                                                       \
  } while (0)

#define GS_REQUIRE(cond, msg)                 \
  do {                                        \
    if (!(cond)) {                            \
      gs_set_error("bad argument: %s", msg);  \
      return GS_ERR_BAD_ARG;                  \
    }                                         \
  } while (0)

#define GS_TRY(expr)             \
  do {                           \
    int s__ = (expr);            \
    if (s__ != GS_OK) return s__; \
  } while (0)

struct gs_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  int sm_count = 0;
  size_t smem_optin = 0;
  uint64_t launches = 0;
  std::vector<const void *> configured;   // kernels whose function attributes were set on this device
  unsigned eval_flags = 0;   // status word of the last gs_spline_eval_* call
  int refs = 0;        // live grids made from this context
  bool dead = false;   // gs_ctx_destroy was called while grids were still alive: freed with the last grid
  // spline pool (host copy + device mirror, re-uploaded when it grows)
  std::vector<double> pool;
  std::vector<int> spline_off;
  double *pool_dev = nullptr;
  size_t pool_dev_len = 0;
  bool pool_dirty = true;
  // staging buffer for host<->device transposes and reductions
  double *stage = nullptr;
  size_t stage_bytes = 0;
  // options
  int64_t opt_threads_per_sm_1d = 0;  // 0 = per-path default
  int64_t opt_block_1d = 128;
  int64_t opt_native_div = 0;
  int64_t opt_const_hoist = 1;
  int64_t opt_pipeline = 1;
  int64_t opt_pipe_cfg = 0;
  int64_t opt_seg_len = 32;
  int64_t opt_hoist_ck = 1;
  int64_t opt_k2r = 1;          // K2r: small constant-property batches (<= 2 blocks per SM, block fits shared memory) stay resident for the call
  int64_t opt_k5 = 1;
  int64_t opt_k5_xstrip = 0;   // 0 = automatic
  int opt_k5_nb = 2;           // K5: tile buffers per warp (3: triple buffer where it fits; measured slower)
  int opt_k5_cluster = 1;      // K5: clusters of 2 / 4 / 8 CTAs per strip when strips alone would not fill the chip
  int opt_k5t = 1;             // K5T: K5's sweeps on tensor-map TMA rings (0: k5_sweep with its cp.async tiles; 1: from 2^20 nodes per sweep; 2: always)
  int opt_k5t_sx = 0, opt_k5t_sy = 0, opt_k5t_s1 = 0;   // K5T: warps (segments) per CTA of the x / y / 1-D sweeps (0 = automatic)
  int opt_k5t_promo = 1;       // K5T x sweep: L2 promotion of the tensor maps (0: 128 B, 1: 256 B; measured: x sweep 4096^2 75 -> 72 us, 16384^2 1.21 -> 1.15 ms)
  int opt_k5t_pdl = 1;         // K5T: programmatic dependent launch (set-up overlaps the previous kernel's tail)
  int opt_k5t_soft = 1;        // K5T: strip groups through global memory when they put more CTAs to work than clusters
  int opt_k5t_ystrip = 0;      // K5T y sweep: columns per CTA (even, <= 32; 0 = automatic)
  int opt_k5t_wide = 1;        // K5T x sweep: 32-column stages loaded as one un-swizzled 34-column box (272-byte rows)
  int opt_k5t_hints = 0;       // K5T: L2 eviction hints, bit 0 loads (phase A evict_last, phase C evict_first), bit 1 stores (evict_first)
  int opt_k5t_nt = 0;          // K5T: tiles per warp ring (0 = as many as fit, at most 8)
  int64_t opt_k6 = 1;
  int64_t opt_faces1d = 0;   // K2f measured slower than K2 on B200 (5.7 vs 4.3 ms on the C2 shape with a Hermite table)
  int64_t opt_k2t = 1;          // K2T: 1-D lines with one table on faces warp + chain warp CTAs with L2-resident scratch (0: K2; 1: cubic tables on equal knots; 2: any single table)
  int64_t opt_k2t_cps = 8;      // K2T: CTAs per SM (2..8; 6+ with two-stage rings); the scratch footprint is CTAs x n x 512 B
  int64_t opt_k2t_ctas = 0;     // K2T: cap on the number of CTAs (0 = none)
  int64_t opt_k2t_hints = 31;   // K2T: evict_first on 1 grid (forward), 2 predict, 4 grid (backward) loads, 8 grid' / predict' stores; 16 discard.L2 of scratch lines once read back
  int64_t opt_k6_chunk = 8;
  int opt_k6_fuse = 0;          // K6: the sweep evaluates the faces along its lines itself (no separate faces pass); measured slower, off
  int opt_k6_cluster = 1;       // K6: give a strip to a cluster of 2 / 4 CTAs when strips alone would not fill the chip
  int64_t opt_auto_exact = 1;   // AUTO: tables / maps take the exact K7 instead of the tolerance family (K6)
  int64_t opt_k7_fuse = 2;      // K7: faces evaluated inside the sweep by a producer warp (k7f_sweep): 0 never, 1 literal constants, 2 tables too
  int64_t opt_k7_nsb = 8;       // K7: most stages of the backward ring (as many as fit shared memory, up to this)
  int64_t opt_k7_promo = 0;     // K7 x sweep: 256-byte L2 promotion of the 16-column boxes
  int64_t opt_k7f_nfw = 0;      // K7F: faces warps per CTA for tables (2, 3 or 4; 0 = automatic: 4 / 3 / 2 for at most one / two / more strips per SM)
  int64_t opt_k7 = 1;           // K7: exact sweeps from precomputed faces (0: first-draft k_xsweep / k_ysweep)
  int64_t opt_host_chunk_mb = 64;
  int64_t opt_ck_block = 16;
};

int gs_ctx_use(gs_ctx *ctx);                 // cudaSetDevice
void gs_ctx_retain(gs_ctx *ctx);             // a grid now refers to the context
void gs_ctx_release(gs_ctx *ctx);            // ... and no longer does
int gs_ctx_sync_pool(gs_ctx *ctx);           // upload the spline pool if it changed
int gs_ctx_stage(gs_ctx *ctx, size_t bytes); // make sure ctx->stage holds at least `bytes`

// K5 (hoisted sweeps for literal-constant coefficients): per-direction tables, valid for one time step
struct gs_k5dir {
  bool valid = false;
  double time = 0.0;
  int kind_first = -1, kind_last = -1, m = 0, S = 0, cs = 1;   // S: segments (warps) per CTA, cs: CTAs per strip
  bool tma = false;     // K5T (tensor-map TMA rings) instead of k5_sweep
  bool aligned = true;  // the grid's arrays were 16-byte aligned when the shape was chosen
  int NT = 0;           // K5T: tiles per warp ring
  long long opt_sig = -1;   // the tuning options the shape was chosen under
  // K5T strip groups without a hardware cluster (more CTAs per strip than fit one GPC): segment results and the solved
  // interface values travel through global memory, the CTAs of a strip meet on a counter (all CTAs are resident)
  bool soft = false;
  double *soft_buf = nullptr;      // [3][strips][cs * S][32]
  unsigned *soft_sync = nullptr;   // [strips][2] arrivals (monotonic), done (epoch)
  size_t soft_len = 0, soft_strips = 0;
  unsigned epoch = 0;
  size_t pool_gen = 0;
  double4 *tab = nullptr, *seg = nullptr;
  double *omega_end = nullptr, *ends = nullptr;
  double ends_host[4] = {0, 0, 0, 0};
};

struct gs_grid1d {
  gs_ctx *ctx = nullptr;
  int64_t nlines = 0, pitch = 0;
  int n = 0, dir = 0;
  int spline_mode = 0;
  int single_off = 0;
  bool all_const = false;       // every selected spline is temperature independent
  int *node_off = nullptr;      // [2n] pool offsets (PER_NODE)
  double *predef = nullptr;     // [2n]
  double *grid = nullptr, *predict = nullptr;
  double *leftT = nullptr, *rightT = nullptr;
  int bc_stride = 0;
  bool bc_set = false;
  double2 *scratch = nullptr;
  size_t scratch_elems = 0;
  // gs_grid1d_step_host pipeline (H2D / compute / D2H on three streams)
  bool hs_init = false;
  cudaStream_t hs_in = nullptr, hs_out = nullptr;
  cudaEvent_t hs_in_done[2], hs_in_free[2], hs_out_ready[2], hs_out_free[2], hs_start;
  double *hs_stage_in[2] = {nullptr, nullptr}, *hs_stage_out[2] = {nullptr, nullptr};
  size_t hs_bytes = 0;
  std::vector<int> node_off_host;  // PER_NODE pool offsets (host copy)
  gs_k5dir k5;                     // K5 tables for long lines
  double *k5_ckpt = nullptr;
  double *faces_CR = nullptr, *faces_C0 = nullptr, *faces_ck = nullptr;   // K2f: face conductivities, checkpoints
  size_t faces_ck_len = 0;
  double *hoist_table = nullptr;  // K2h coefficients for hoist_time
  bool hoist_valid = false;
  double hoist_time = 0.0;
  unsigned *status = nullptr;
  int solver = GS_SOLVER_AUTO;
};

struct gs_dim2d {
  int dir = 0, n = 0;
  double *values = nullptr;  // [n+2] low ++ range ++ upp
  double *coefs = nullptr;   // [2n] Dim.coefs, sigma = 1.0
  double hf[4] = {0, 0, 0, 0};  // firstVol, fC, lastVol, lA
};

struct gs_bound2d {
  int kind = GS_TEMP, rep = GS_SPARSE;
  double *values = nullptr;  // device, side length (SPARSE keeps the scalar broadcast)
};

struct gs_grid2d {
  gs_ctx *ctx = nullptr;
  int cols = 0, rows = 0;
  gs_dim2d x, y;
  gs_bound2d b[4];
  double *grid = nullptr, *predict = nullptr, *result = nullptr;
  bool bound[3] = {false, false, false};   // array lives in caller-owned device memory (gs_grid2d_bind_device)
  bool result_is_grid = false;  // after a fused iteration `result` equals `grid`
  int map_mode[4] = {0, 0, 0, 0};
  int *map[4] = {nullptr, nullptr, nullptr, nullptr};  // device, pool OFFSETS (not ids)
  std::vector<int> map_host[4];                         // ids as given
  double2 *scratch = nullptr;
  size_t scratch_elems = 0;
  double *iface = nullptr;        // partitioned solver: interface coefficients / values
  size_t iface_len = 0;
  double *predict_alt = nullptr;  // partitioned fused y sweep writes predict' here, then the two swap
  gs_k5dir k5x, k5y;
  double *k5_ckpt = nullptr;
  size_t k5_ckpt_len = 0;
  double *k6_F = nullptr, *k6_ends = nullptr;   // K6: face values, end rows
  double4 *k6_ckpt = nullptr;
  double *k6_split = nullptr;                    // split y sweep: fm1 [nlines] + cdw [S][3][nlines padded]
  double2 *k7_scratch = nullptr;                 // K7: (delta, lambda) [strips][n][32]
  size_t k7_scratch_elems = 0;
  bool k6_split_open = false;                    // a begin call is waiting for its end call
  unsigned *status = nullptr;
  int solver = GS_SOLVER_AUTO;
};

bool gs_k5_1d_usable(gs_grid1d *g);
int gs_k5_1d_step(gs_grid1d *g, double time, int nsteps);
