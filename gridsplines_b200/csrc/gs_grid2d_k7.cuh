// gs_grid2d_k7.cuh -- K7: the fast BIT-EXACT 2-D sweep (GS_SOLVER_THOMAS, and AUTO where it applies).
//
// The reference walks a line once forward (Dim.first / general / last -> passion.forwardFirst / Unit / Last,
// A/Dim.scala:28-128, A/passion/package.scala:245-276) and once backward (backwardRow / backwardColumn :303-342).  Two
// things make the first-draft kernels (k_xsweep / k_ysweep) slow: the interval average of the property table sits
// inside the dependent chain, and every global access is issued by the thread that then waits for it.  Here
//   * every spline evaluation of a sweep happens beforehand, fully parallel, in k6_ends + k6_faces (gs_grid2d_k6.cuh):
//     F[i] = Dim.general's `co` of node i -- the reference's own doubles, so nothing changes numerically;
//   * k7_sweep then runs ONE THREAD PER LINE in the reference's operation order -- per node the very operations of
//     Dim.general + forwardUnit, with the shared-reciprocal division of gs_device.cuh (correctly rounded: same bits
//     as `/`; a warp whose operands leave the fast divider's range repeats its lines with native division) --
//     fed by a per-warp ring of TMA stages guarded by mbarriers.  Every tile is ONE tensor-map instruction
//     (cp.async.bulk.tensor.2d, UTMALDG / UTMASTG) issued by lane 0:
//        forward : stage = a 16-node x 32-line box of F and of the right-hand side + the nodes' (cf0, cf1);
//                  (delta, lambda) = toPassion goes to a blocked scratch array [strip][node][lane] (512 B contiguous
//                  per node and warp);
//        backward: stage = 16 nodes of (delta, lambda) (one 8 KB bulk copy) + (fused y sweep) the box of old grid
//                  values; x_i = delta_i * x_{i+1} + lambda_i, Grid.update (:474-481) applied on the way out (y);
//                  the x sweep's results leave as one box store per chunk.
//   y sweep: lane <-> column, box = 32 columns x 16 rows (a node row of the strip is 256 contiguous bytes);
//   x sweep: lane <-> row, box = 16 columns x 32 rows -- the TMA/shared-memory tiled transpose: the box lands with
//            the 128-byte swizzle, so a lane reads its own row 16 bytes at a time without bank conflicts.
//   Boxes that stick out of the array (last strip, last chunk) are zero-filled on load and clipped on store.
// One warp per CTA (46 KB of shared memory): 4096 lines = 128 CTAs, 16384 lines = 512 CTAs, all resident at once.
//
// Cost model (B200): the dependent chain of a node is a*delta, +diag, reciprocal (MUFU + 5 DFMA), quotient (3 DFMA)
// ~ 140 clk forward and 2 dependent FP64 operations backward, so a sweep takes n x ~165 clk while lines are few
// (4096^2: ~0.35 ms per sweep) and becomes HBM-bound once ~3 warps per SM are resident (16384^2).
// HBM traffic per node: faces pass 16 B (R predict, W F); x sweep 56 B (R F, R grid, W+R (delta, lambda), W result);
// y sweep 72 B (R F, R result, W+R (delta, lambda), R grid, W grid', W predict').
#pragma once
#include "gs_tma.cuh"

#define K7_CH 16                         // nodes per stage
#define K7_NS 3                          // ring depth
#define K7_TILE 4096                     // one box: 16 x 32 doubles
#define K7_STAGE 12288                   // bytes per stage = max(forward 2 tiles + cf, backward 8 KB + tile)
#define K7_NSB 8                         // most stages of the backward ring (as many as fit: see k7_bwd_ring)
#define K7_SMEM (1024 + 1024 + K7_NS * K7_STAGE + 2 * K7_TILE)   // alignment slack + barriers + stages + output tiles (< 48 KB: no opt-in)

// CUtensorMap is an opaque 128-byte, 64-byte aligned blob (cuda.h); kept as bytes here so that device code does not
// need the driver header
struct __align__(64) K7Map {
  unsigned char b[128];
};
struct K7Maps {
  K7Map F, rhs, gold, out;   // boxes of k6_F, the right-hand side, the old grid (fused y) and the result (x)
  K7Map pred;                // fused faces: predict, box 32 x 17 (y) / 18 x 32 (x)
};

struct K7Params {
  int n, nlines, cols, rows;
  const double *coefs;             // Dim.coefs [2n]
  const double *first3, *last3;    // k6_ends
  double *out;                     // unfused y: result
  double *grid_io;                 // fused y: new grid out
  double *predict_out;             // fused y: new predict
  double2 *scratch;                // [strips][n][32] (delta, lambda)
  double time;
  unsigned *status;
  // fused faces (k7f_sweep): the carry out of node 0 from k6_ends, the spline pool and the coefficient maps
  const double *face0;
  const double *pool;
  int pool_len;
  int stage_off, stage_len;        // the part of the pool staged in shared memory: all of it, or the one table of a SINGLE map
  int nsb;                         // most stages of the backward ring (option k7_nsb)
  Maps2d maps;
};

// ring bookkeeping of one warp: stage s is used for the chunks s, s + NS, ...; bit s of `phase` is the parity the
// next wait on its barrier expects
struct K7Ring {
  unsigned char *stages;
  uint64_t *bars;
  uint32_t phase;
  int head;      // stage of the next chunk to be consumed
  int tail;      // stage of the next chunk to be issued
  int ns;        // depth
  int stride;    // bytes per stage
};
GS_DEV unsigned char *k7_stage(const K7Ring &r, int s) { return r.stages + (size_t)s * r.stride; }
GS_DEV void k7_wait(K7Ring &r, unsigned *status) {
  gs_mbar_wait_bounded(&r.bars[r.head], (r.phase >> r.head) & 1u, status);
  r.phase ^= 1u << r.head;
}
GS_DEV void k7_advance(const K7Ring &r, int &s) { s = (s == r.ns - 1) ? 0 : s + 1; }

// element (row r, column j) of a 32 x 16 box that landed with CU_TENSOR_MAP_SWIZZLE_128B (tile base 1024 B aligned):
// the 16-byte unit index j / 2 is XORed with r % 8
GS_DEV int k7_swz(int r, int j) { return r * 128 + ((((j >> 1) ^ (r & 7)) << 4) | ((j & 1) << 3)); }

// forward stage layout:  F tile | rhs tile | cf [16][2]
//   y: tile = [16 nodes][32 lines];   x: tile = [32 lines][16 nodes], swizzled
template <bool XDIR>
GS_DEV void k7_issue_fwd(const K7Params &q, const K7Maps &tm, K7Ring &r, int c, int line0, int lane) {
  if (lane == 0) {
    const int i0 = c * K7_CH, len = min(K7_CH, q.n - i0);
    unsigned char *st = k7_stage(r, r.tail);
    uint64_t *bar = &r.bars[r.tail];
    gs_mbar_expect_tx(bar, (uint32_t)(2 * K7_TILE + len * 16));
    const int c0 = XDIR ? i0 : line0, c1 = XDIR ? line0 : i0;
    gs_tma_load_2d(st, &tm.F, c0, c1, bar);
    gs_tma_load_2d(st + K7_TILE, &tm.rhs, c0, c1, bar);
    gs_bulk_g2s(st + 2 * K7_TILE, q.coefs + 2 * (size_t)i0, (uint32_t)len * 16u, bar);
  }
  k7_advance(r, r.tail);
}

// The forward walk of one line (lane) over all chunks; toPassion -> scratch.  Returns the sticky range predicate of
// the shared-reciprocal division (always true when !FAST).
template <bool XDIR, bool FAST>
GS_DEV bool k7_forward(const K7Params &q, const K7Maps &tm, K7Ring &r, int line0, int lane, int rl, GsRcp rt,
                       double inv_time, double2 *sc, unsigned &flags) {
  const int n = q.n, NC = (n + K7_CH - 1) / K7_CH;
  const size_t line = (size_t)(line0 + rl);
  const double f_diag = q.first3[3 * line], f_sup = q.first3[3 * line + 1], f_rhs = q.first3[3 * line + 2];
  const double l_sub = q.last3[3 * line], l_diag = q.last3[3 * line + 1], l_rhs = q.last3[3 * line + 2];
  bool ok = true;
  double delta = 0.0, lambda = 0.0, fprev = 0.0;
  for (int c = 0; c < min(r.ns, NC); ++c) k7_issue_fwd<XDIR>(q, tm, r, c, line0, lane);
  // Dim.general :67-83 + forwardUnit :257-265 of an interior node
  auto node_mid = [&](double t, double F, double cf0, double cf1) {
    const double a = cf0 * fprev;
    const double cc = cf1 * F;
    const double vect = gs_div<FAST, true>(-t, rt, ok);
    gs_forward_unit<FAST>(a, -(a + cc) - inv_time, cc, vect, delta, lambda, flags, ok);
    fprev = F;
  };
#pragma unroll 1
  for (int c = 0; c < NC; ++c) {
    const int i0 = c * K7_CH, len = min(K7_CH, n - i0);
    k7_wait(r, q.status);
    const unsigned char *st = k7_stage(r, r.head);
    const double2 *cf = reinterpret_cast<const double2 *>(st + 2 * K7_TILE);
    double2 *so = sc + (size_t)i0 * 32 + lane;
    if (c != 0 && i0 + K7_CH < n) {
      // a full chunk holding neither node 0 nor node n - 1
      if (XDIR) {
        const unsigned char *rF = st + rl * 128, *rT = st + K7_TILE + rl * 128;
        const int sw = rl & 7;
#pragma unroll
        for (int j = 0; j < K7_CH; j += 2) {
          const int u = ((j >> 1) ^ sw) << 4;
          const double2 f2 = *reinterpret_cast<const double2 *>(rF + u), t2 = *reinterpret_cast<const double2 *>(rT + u);
          node_mid(t2.x, f2.x, cf[j].x, cf[j].y);
          so[(size_t)j * 32] = make_double2(delta, lambda);
          node_mid(t2.y, f2.y, cf[j + 1].x, cf[j + 1].y);
          so[(size_t)(j + 1) * 32] = make_double2(delta, lambda);
        }
      } else {
        const double *cF = reinterpret_cast<const double *>(st) + rl;
        const double *cT = reinterpret_cast<const double *>(st + K7_TILE) + rl;
#pragma unroll
        for (int j = 0; j < K7_CH; ++j) {
          node_mid(cT[j * 32], cF[j * 32], cf[j].x, cf[j].y);
          so[(size_t)j * 32] = make_double2(delta, lambda);
        }
      }
    } else {
#pragma unroll 1
      for (int j = 0; j < len; ++j) {
        const int i = i0 + j;
        double F, t;
        if (XDIR) {
          F = *reinterpret_cast<const double *>(st + k7_swz(rl, j));
          t = *reinterpret_cast<const double *>(st + K7_TILE + k7_swz(rl, j));
        } else {
          F = reinterpret_cast<const double *>(st)[j * 32 + rl];
          t = reinterpret_cast<const double *>(st + K7_TILE)[j * 32 + rl];
        }
        if (i == 0) {
          gs_forward_first<FAST>(f_diag, f_sup, f_rhs, delta, lambda, ok);   // the row k6_ends assembled (:28-64)
          fprev = F;                                                         // co, or cond / cap after a heat-flow end
        } else if (i == n - 1) {
          gs_forward_last<FAST>(l_sub, l_diag, l_rhs, delta, lambda, ok);    // (:86-128); delta = 0.0
        } else {
          node_mid(t, F, cf[j].x, cf[j].y);
        }
        so[(size_t)j * 32] = make_double2(delta, lambda);
      }
    }
    k7_advance(r, r.head);
    __syncwarp();
    if (c + r.ns < NC) k7_issue_fwd<XDIR>(q, tm, r, c + r.ns, line0, lane);
  }
  return ok;
}

// The backward ring: stages of 8 KB (+ 4 KB with the old grid tile), as many as fit `capacity` bytes.  The backward walk
// has two dependent FP64 operations per node and runs at HBM speed: timed alone at 16384^2 (forward skipped, stale
// scratch) it takes 1.10 ms (x, 24 B/node: 5.9 TB/s) and 1.84 ms (y, 40 B/node: 5.8 TB/s); 2, 3, 4 or 6 stages measure
// the same (option k7_nsb).  The FORWARD walk is what costs: 2.30 / 2.25 ms per sweep alone, ~4400 clk per 16-node chunk
// and CTA against ~2400 for the bare chain -- the SM issues ~1.9 instructions per clock for 150 warp instructions per
// node-row (faces 102, chain 45), FP64 pipe ~40 % busy (profiles/r2_k7f_phases.md).
template <bool FUSE>
GS_DEV void k7_bwd_ring(K7Ring &r, unsigned char *stages, uint64_t *bars, int capacity, int most) {
  r.stages = stages;
  r.bars = bars;
  r.phase = 0;
  r.head = r.tail = 0;
  r.stride = K7_CH * 512 + (FUSE ? K7_TILE : 0);
  r.ns = max(1, min(min(K7_NSB, most), capacity / r.stride));
}
// backward stage layout: (delta, lambda) [16][32] double2 (8 KB) | old grid tile [16][32] (fused y sweep)
template <bool FUSE>
GS_DEV void k7_issue_bwd(const K7Params &q, const K7Maps &tm, K7Ring &r, int c, int line0, int lane, const double2 *sc) {
  if (lane == 0) {
    const int i0 = c * K7_CH, len = min(K7_CH, q.n - i0);
    unsigned char *st = k7_stage(r, r.tail);
    uint64_t *bar = &r.bars[r.tail];
    gs_mbar_expect_tx(bar, (uint32_t)(len * 512 + (FUSE ? K7_TILE : 0)));
    gs_bulk_g2s(st, sc + (size_t)i0 * 32, (uint32_t)len * 512u, bar);
    if (FUSE) gs_tma_load_2d(st + K7_CH * 512, &tm.gold, line0, i0, bar);
  }
  k7_advance(r, r.tail);
}

// The backward walk (backwardRow :303-322 / backwardColumn :324-342 [+ Grid.update :474-481]) of one warp's 32 lines:
// (delta, lambda) come back from the scratch array through the ring, results leave as described at the top.
template <bool XDIR, bool FUSE>
GS_DEV double k7_backward(const K7Params &q, const K7Maps &tm, K7Ring &r, unsigned char *otile, int line0, bool act,
                          int lane, const double2 *sc) {
  const int n = q.n, NC = (n + K7_CH - 1) / K7_CH;
  for (int k = 0; k < min(r.ns, NC); ++k) k7_issue_bwd<FUSE>(q, tm, r, NC - 1 - k, line0, lane, sc);
  double u = 0.0;
  int ot = 0;
#pragma unroll 1
  for (int c = NC - 1; c >= 0; --c) {
    const int i0 = c * K7_CH, len = min(K7_CH, n - i0);
    k7_wait(r, q.status);
    const unsigned char *st = k7_stage(r, r.head);
    const double2 *dl = reinterpret_cast<const double2 *>(st) + lane;
    double xs[K7_CH];
    if (len == K7_CH) {
#pragma unroll
      for (int j = K7_CH - 1; j >= 0; --j) {
        const double2 v = dl[j * 32];
        u = v.x * u + v.y;
        xs[j] = u;
      }
    } else {
#pragma unroll
      for (int j = K7_CH - 1; j >= 0; --j) {
        xs[j] = 0.0;
        if (j < len) {
          const double2 v = dl[j * 32];
          u = v.x * u + v.y;
          xs[j] = u;
        }
      }
    }
    if (XDIR) {
      // transposed (swizzled) output tile: every lane writes its row, lane 0 stores the box
      unsigned char *orow = otile + ot * K7_TILE + lane * 128;
      if (lane == 0) gs_bulk_wait_read<1>();    // the store issued from this tile two chunks ago has read it
      __syncwarp();
      const int sw = lane & 7;
#pragma unroll
      for (int j = 0; j < K7_CH; j += 2)
        *reinterpret_cast<double2 *>(orow + (((j >> 1) ^ sw) << 4)) = make_double2(xs[j], xs[j + 1]);
      gs_fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        gs_tma_store_2d(&tm.out, i0, line0, otile + ot * K7_TILE);
        gs_bulk_commit();
      }
      ot ^= 1;
    } else if (act) {
      const size_t cs = (size_t)q.cols, g0 = (size_t)i0 * cs + line0 + lane;
      if (FUSE) {
        const double *gold = reinterpret_cast<const double *>(st + K7_CH * 512) + lane;
        double *G = q.grid_io + g0, *P = q.predict_out + g0;
        if (len == K7_CH) {
#pragma unroll
          for (int j = 0; j < K7_CH; ++j, G += cs, P += cs) {
            *P = aval2d(gold[j * 32], xs[j]);
            *G = xs[j];
          }
        } else {
#pragma unroll
          for (int j = 0; j < K7_CH; ++j, G += cs, P += cs) {
            if (j < len) {
              *P = aval2d(gold[j * 32], xs[j]);
              *G = xs[j];
            }
          }
        }
      } else {
        double *O = q.out + g0;
        if (len == K7_CH) {
#pragma unroll
          for (int j = 0; j < K7_CH; ++j, O += cs) *O = xs[j];
        } else {
#pragma unroll
          for (int j = 0; j < K7_CH; ++j, O += cs)
            if (j < len) *O = xs[j];
        }
      }
    }
    k7_advance(r, r.head);
    __syncwarp();
    if (c - r.ns >= 0) k7_issue_bwd<FUSE>(q, tm, r, c - r.ns, line0, lane, sc);
  }
  return u;
}

template <bool XDIR, bool FUSE>
__global__ void __launch_bounds__(32) k7_sweep(K7Params q, const __grid_constant__ K7Maps tm) {
  extern __shared__ unsigned char k7raw[];
  // the swizzled boxes need a 1024-byte aligned base
  unsigned char *k7s = k7raw + ((1024u - (gs_smem_u32(k7raw) & 1023u)) & 1023u);
  const int lane = threadIdx.x;
  const int n = q.n, NC = (n + K7_CH - 1) / K7_CH;
  const int line0 = blockIdx.x * 32, nl = min(32, q.nlines - line0);
  const bool act = lane < nl;
  const int rl = act ? lane : nl - 1;      // lanes past the last line shadow it (their stores are suppressed)
  K7Ring r;
  r.bars = reinterpret_cast<uint64_t *>(k7s);
  r.stages = k7s + 1024;
  r.phase = 0;
  r.head = r.tail = 0;
  r.ns = K7_NS;
  r.stride = K7_STAGE;
  uint64_t *bwd_bars = r.bars + K7_NS;
  unsigned char *otile = k7s + 1024 + K7_NS * K7_STAGE;
  if (lane == 0) {
    for (int s = 0; s < K7_NS + K7_NSB; ++s) gs_mbar_init(&r.bars[s], 1);
    gs_mbar_init_fence();
    gs_tma_prefetch_desc(&tm.F);
    gs_tma_prefetch_desc(&tm.rhs);
    if (FUSE) gs_tma_prefetch_desc(&tm.gold);
    if (XDIR) gs_tma_prefetch_desc(&tm.out);
  }
  gs_fence_proxy_async();
  __syncwarp();
  bool fast_allowed = true;
  const GsRcp rt = gs_rcp_invariant(q.time, fast_allowed);
  const double inv_time = 1.0 / q.time;
  double2 *sc = q.scratch + (size_t)blockIdx.x * (size_t)n * 32;
  unsigned flags = 0;

  // ---- forward (Dim.* + passion.forward*) ----
  {
    bool ok = false;
    unsigned fl = 0;
    if (fast_allowed) {
      ok = k7_forward<XDIR, true>(q, tm, r, line0, lane, rl, rt, inv_time, sc, fl);
      ok = __all_sync(0xffffffffu, ok);
    }
    if (!ok) {
      // an operand outside the shared-reciprocal divider's range (zero, denormal, Inf, NaN, absurd magnitude) somewhere
      // in this warp's lines: the same walk with the native division
      fl = 0;
      __syncwarp();
      k7_forward<XDIR, false>(q, tm, r, line0, lane, rl, rt, inv_time, sc, fl);
    }
    flags |= fl;
  }
  // the (delta, lambda) this warp wrote (generic proxy) are read back by bulk copies (async proxy)
  __threadfence();
  gs_fence_proxy_async();
  __syncwarp();

  // the x sweep's output tiles follow the stages; the y sweeps store from registers and may use that room too
  k7_bwd_ring<FUSE>(r, k7s + 1024, bwd_bars, K7_NS * K7_STAGE + (XDIR ? 0 : 2 * K7_TILE), q.nsb);
  const double u = k7_backward<XDIR, FUSE>(q, tm, r, otile, line0, act, lane, sc);
  if (XDIR && lane == 0) gs_bulk_wait_read<0>();
  if (act && !(fabs(u) <= 1.7976931348623157e308)) flags |= GS_FLAG_NONFINITE;
  gs_raise(q.status, flags);
}

// ---------------------------------------------------------------------------------------------------------------
// k7f_sweep: K7 with the faces evaluated INSIDE the sweep by a second warp (warp-specialised producer), so that the
// ~110 instructions of a face (interval average of the property table) run in the shadow of the chain warp's
// dependent divisions instead of in a pass of their own, and F never travels through HBM:
//   warp 1 (faces): TMA box of `predict` (16 + 1 nodes x 32 lines)  ->  F tile in shared memory (the tiers of
//                   k6_faces: same doubles), and the TMA loads of the right-hand side / (cf0, cf1) for the chain warp;
//   warp 0 (chain): waits for (F, rhs, cf) of a chunk on ONE mbarrier (transaction bytes of the loads + the faces
//                   warp's arrival), runs the reference's recurrence, hands the stage back on an `empty` barrier.
// The backward walk is K7's (k7_backward).  HBM traffic per node: x sweep 56 B (R predict, R grid, W + R (delta,
// lambda), W result), y sweep 72 B.  Used when the spline pool fits 4 KB of shared memory (or the coefficients are
// literal constants); otherwise K7 with its separate faces pass.
// Shared memory (from a 1024-byte aligned base): barriers | ring region 39 KB | 2 output tiles | spline pool
//   forward : 3 x (rhs tile 4 KB | F tile 4 KB) | 3 x cf 256 B | 3 x predict tile 4.5 KB
//   backward: 3 x 12 KB stages (K7)
// ---------------------------------------------------------------------------------------------------------------
#define K7F_FT 8192                       // rhs tile + F tile
#define K7F_CF (3 * K7F_FT)               // offset of the cf ring in the region
#define K7F_P (K7F_CF + 1024)             // offset of the predict-tile ring
#define K7F_PT 4608                       // one predict tile: y 17 x 32, x 32 x 18 doubles
#define K7F_REGION 39936
#define K7F_POOL 4096
#define K7F_SMEM (1024 + 1024 + K7F_REGION + 2 * K7_TILE + K7F_POOL)
static_assert(K7F_P + 3 * K7F_PT <= K7F_REGION && K7_NS * K7_STAGE <= K7F_REGION, "K7F shared-memory layout");

struct K7FBars {
  uint64_t bwd[K7_NSB];  // backward ring (K7Ring)
  uint64_t ft[3];        // (F, rhs, cf) of a chunk ready: tx bytes of the loads + 1 arrival (faces warp)
  uint64_t empty[3];     // the chain warp is done with the stage
  uint64_t pin[3];       // predict tile landed
  uint64_t pdone[3];     // every faces warp is done with the predict tile (NFW > 1)
  int redo;              // the forward pass is repeated with the native division
};

// the chain warp's forward walk: as k7_forward, with F from the faces warp's tile and no loads of its own
template <bool XDIR, bool FAST>
GS_DEV bool k7f_forward(const K7Params &q, K7FBars *bars, unsigned char *region, uint32_t &ft_phase, int line0, int lane,
                        int rl, GsRcp rt, double inv_time, double2 *sc, unsigned &flags) {
  const int n = q.n, NC = (n + K7_CH - 1) / K7_CH;
  const size_t line = (size_t)(line0 + rl);
  const double f_diag = q.first3[3 * line], f_sup = q.first3[3 * line + 1], f_rhs = q.first3[3 * line + 2];
  const double l_sub = q.last3[3 * line], l_diag = q.last3[3 * line + 1], l_rhs = q.last3[3 * line + 2];
  bool ok = true;
  double delta = 0.0, lambda = 0.0, fprev = 0.0;
  auto node_mid = [&](double t, double F, double cf0, double cf1) {
    const double a = cf0 * fprev;
    const double cc = cf1 * F;
    const double vect = gs_div<FAST, true>(-t, rt, ok);
    gs_forward_unit<FAST>(a, -(a + cc) - inv_time, cc, vect, delta, lambda, flags, ok);
    fprev = F;
  };
  int s = 0;
#pragma unroll 1
  for (int c = 0; c < NC; ++c) {
    const int i0 = c * K7_CH, len = min(K7_CH, n - i0);
    gs_mbar_wait_bounded(&bars->ft[s], (ft_phase >> s) & 1u, q.status);
    ft_phase ^= 1u << s;
    const unsigned char *st = region + s * K7F_FT;
    const double *cF = reinterpret_cast<const double *>(st + K7_TILE) + rl;
    const double2 *cf = reinterpret_cast<const double2 *>(region + K7F_CF + s * 256);
    double2 *so = sc + (size_t)i0 * 32 + lane;
    if (c != 0 && i0 + K7_CH < n) {
      if (XDIR) {
        const unsigned char *rT = st + rl * 128;
        const int sw = rl & 7;
#pragma unroll
        for (int j = 0; j < K7_CH; j += 2) {
          const double2 t2 = *reinterpret_cast<const double2 *>(rT + (((j >> 1) ^ sw) << 4));
          node_mid(t2.x, cF[j * 32], cf[j].x, cf[j].y);
          so[(size_t)j * 32] = make_double2(delta, lambda);
          node_mid(t2.y, cF[(j + 1) * 32], cf[j + 1].x, cf[j + 1].y);
          so[(size_t)(j + 1) * 32] = make_double2(delta, lambda);
        }
      } else {
        const double *cT = reinterpret_cast<const double *>(st) + rl;
#pragma unroll
        for (int j = 0; j < K7_CH; ++j) {
          node_mid(cT[j * 32], cF[j * 32], cf[j].x, cf[j].y);
          so[(size_t)j * 32] = make_double2(delta, lambda);
        }
      }
    } else {
#pragma unroll 1
      for (int j = 0; j < len; ++j) {
        const int i = i0 + j;
        const double F = cF[j * 32];
        const double t = XDIR ? *reinterpret_cast<const double *>(st + k7_swz(rl, j))
                              : reinterpret_cast<const double *>(st)[j * 32 + rl];
        if (i == 0) {
          gs_forward_first<FAST>(f_diag, f_sup, f_rhs, delta, lambda, ok);
          fprev = F;
        } else if (i == n - 1) {
          gs_forward_last<FAST>(l_sub, l_diag, l_rhs, delta, lambda, ok);
        } else {
          node_mid(t, F, cf[j].x, cf[j].y);
        }
        so[(size_t)j * 32] = make_double2(delta, lambda);
      }
    }
    __syncwarp();
    if (lane == 0) gs_mbar_arrive(&bars->empty[s]);
    s = (s == 2) ? 0 : s + 1;
  }
  return ok;
}

// A faces warp's pass over all chunks of the strip.  NFW faces warps share every chunk: warp fw evaluates the faces of
// nodes [fw * 16 / NFW, (fw + 1) * 16 / NFW) of the chunk; warp 0 issues all loads.
template <bool XDIR, bool LIT, int NFW>
GS_DEV void k7f_faces(const K7Params &q, const K7Maps &tm, K7FBars *bars, unsigned char *region, const double *pool_sm,
                      uint32_t &e_phase, uint32_t &p_phase, uint32_t &d_phase, int fw, int line0, int lane, int rl,
                      unsigned &flags) {
  // warp fw takes the nodes [j0, j1) of every chunk (NFW = 3: five, five and six)
  constexpr bool EVEN = K7_CH % NFW == 0;
  constexpr int PER = (K7_CH + NFW - 1) / NFW;
  const int n = q.n, NC = (n + K7_CH - 1) / K7_CH;
  const int line = line0 + rl;
  const int j0 = fw * K7_CH / NFW, j1 = (fw + 1) * K7_CH / NFW;
  auto issue_p = [&](int c) {
    if (lane == 0) {
      const int p = c % 3, i0 = c * K7_CH;
      gs_mbar_expect_tx(&bars->pin[p], (uint32_t)(XDIR ? 32 * 18 * 8 : 17 * 32 * 8));
      gs_tma_load_2d(region + K7F_P + p * K7F_PT, &tm.pred, XDIR ? i0 : line0, XDIR ? line0 : i0, &bars->pin[p]);
    }
  };
  const bool single = LIT || q.maps.mode[GS_SEL_APPLY] == GS_MAP_SINGLE;
  int off = q.maps.single_off[GS_SEL_APPLY];
  K6Tab z = {};
  GsLitConst zl;
  zl.v = 0.0;
  if (LIT) zl.v = q.pool[off + GS_HDR + 3];
  else z = k6_tab(pool_sm + (off - q.stage_off));
  auto face = [&](double p0, double p1, int i) -> double {
    double v;
    if (LIT) {
      bool okf = true;
      v = gs_take_average<true>(zl, p0, p1, flags, okf);
      if (!okf) {
        bool ok2 = true;
        v = gs_take_average<false>(zl, p0, p1, flags, ok2);   // operands outside the fast divider's range
      }
      return v;
    }
    if (!single) {
      off = map_off(q.maps, GS_SEL_APPLY, XDIR ? i : line, XDIR ? line : i, q.cols);
      z = k6_tab(pool_sm + (off - q.stage_off));
    }
    if (!k6_face_fast(z, p0, p1, v) && !k6_face_equal(z, p0, p1, v) && !k6_face_clamped(pool_sm + (off - q.stage_off), p0, p1, v)) {
      if (!k6_face_mid(pool_sm + (off - q.stage_off), p0, p1, &v)) v = k6_face_general(q.pool, off, p0, p1, q.status);
    }
    return v;
  };
  if (fw == 0)
    for (int c = 0; c < min(3, NC); ++c) issue_p(c);
  int s = 0;
#pragma unroll 1
  for (int c = 0; c < NC; ++c) {
    const int i0 = c * K7_CH, len = min(K7_CH, n - i0);
    // the stage is free again (the chain warp consumed chunk c - 3): start the chain warp's own loads into it
    gs_mbar_wait_bounded(&bars->empty[s], (e_phase >> s) & 1u, q.status);
    e_phase ^= 1u << s;
    unsigned char *st = region + s * K7F_FT;
    if (fw == 0 && lane == 0) {
      gs_mbar_expect_tx_only(&bars->ft[s], (uint32_t)(K7_TILE + len * 16));
      gs_tma_load_2d(st, &tm.rhs, XDIR ? i0 : line0, XDIR ? line0 : i0, &bars->ft[s]);
      gs_bulk_g2s(region + K7F_CF + s * 256, q.coefs + 2 * (size_t)i0, (uint32_t)len * 16u, &bars->ft[s]);
    }
    gs_mbar_wait_bounded(&bars->pin[s], (p_phase >> s) & 1u, q.status);
    p_phase ^= 1u << s;
    const double *P = reinterpret_cast<const double *>(region + K7F_P + s * K7F_PT);
    double *F = reinterpret_cast<double *>(st + K7_TILE) + lane;
    auto pred0 = [&](int j) { return XDIR ? P[rl * 18 + j] : P[j * 32 + rl]; };
    if (c != 0 && i0 + K7_CH < n && single) {
      // A single warp has nobody to hide its latencies behind: eight faces at a time go through the straight-line
      // tier together (independent instruction streams the scheduler can interleave); the few that fall out of it
      // -- equal temperatures, a straddled knot, clamps -- are redone one by one afterwards.
      constexpr int B = PER < 8 ? PER : 8;   // faces per batch (a batch may run past j1: its last face is then dropped)
#pragma unroll
      for (int h0 = 0; h0 < PER; h0 += B) {
        const int h = j0 + h0;
        double v[B];
        unsigned bad = 0;
#pragma unroll
        for (int j = 0; j < B; ++j) {
          const double p0 = pred0(h + j), p1 = pred0(h + j + 1);
          bool good;
          if (LIT) {
            const double d = fabs(p0 - p1);      // max - min; (d * v) / d as gs_take_average(GsLitConst)
            good = true;
            v[j] = gs_div<true>(d * zl.v, gs_rcp<true>(d, good), good);
            if (p0 == p1) {                      // average(lo, lo) = apply(lo) = the constant (a uniform region)
              v[j] = zl.v;
              good = true;
            }
          } else {
            good = k6_face_fast(z, p0, p1, v[j]);
          }
          if (!good && (EVEN || h + j < j1)) bad |= 1u << j;
        }
        if (bad) {
          // a uniform region (equal neighbours: average = apply) or one beyond the table's clamps fails the straight-line
          // tier for whole batches: take its faces through the equal-temperature / clamped tiers together, again as independent instruction streams, before
          // anything goes to the one-by-one path (512^2 from a uniform start: 520 -> 300 us per step)
          if (!LIT) {
#pragma unroll
            for (int j = 0; j < B; ++j) {
              double ve;
              const bool eq = k6_face_equal(z, pred0(h + j), pred0(h + j + 1), ve) ||
                              k6_face_clamped(pool_sm + (off - q.stage_off), pred0(h + j), pred0(h + j + 1), ve);
              if (eq && ((bad >> j) & 1u)) {
                v[j] = ve;
                bad &= ~(1u << j);
              }
            }
          }
#pragma unroll
          for (int j = 0; j < B; ++j)
            if ((bad >> j) & 1u) v[j] = face(pred0(h + j), pred0(h + j + 1), i0 + h + j);
        }
#pragma unroll
        for (int j = 0; j < B; ++j)
          if (EVEN || h + j < j1) F[(h + j) * 32] = v[j];
      }
    } else {
#pragma unroll 1
      for (int j = j0; j < min(len, j1); ++j) {
        const int i = i0 + j;
        double v = 0.0;
        if (i == 0) v = q.face0[line];          // carry out of node 0 (k6_ends): co, or cond / cap after a heat-flow end
        else if (i < n - 1) v = face(pred0(j), pred0(j + 1), i);
        F[j * 32] = v;
      }
    }
    __syncwarp();
    if (lane == 0) {
      gs_mbar_arrive(&bars->ft[s]);                        // release: this warp's part of the F tile is visible
      if (NFW > 1) gs_mbar_arrive(&bars->pdone[s]);        // ... and it is done with the predict tile
    }
    if (fw == 0 && c + 3 < NC) {
      if (NFW > 1) {   // every faces warp has read the predict tile
        gs_mbar_wait_bounded(&bars->pdone[s], (d_phase >> s) & 1u, q.status);
        d_phase ^= 1u << s;
      }
      issue_p(c + 3);
    } else if (fw == 0 && NFW > 1) {
      gs_mbar_wait_bounded(&bars->pdone[s], (d_phase >> s) & 1u, q.status);   // keeps the phase bits in step
      d_phase ^= 1u << s;
    }
    s = (s == 2) ? 0 : s + 1;
  }
}

template <bool XDIR, bool FUSE, bool LIT, int NFW>
__global__ void __launch_bounds__(32 + 32 * NFW, NFW >= 4 ? 1 : 4) k7f_sweep(K7Params q, const __grid_constant__ K7Maps tm) {
  extern __shared__ unsigned char k7raw[];
  unsigned char *k7s = k7raw + ((1024u - (gs_smem_u32(k7raw) & 1023u)) & 1023u);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int n = q.n;
  const int line0 = blockIdx.x * 32, nl = min(32, q.nlines - line0);
  const bool act = lane < nl;
  const int rl = act ? lane : nl - 1;
  K7FBars *bars = reinterpret_cast<K7FBars *>(k7s);
  unsigned char *region = k7s + 1024;
  unsigned char *otile = region + K7F_REGION;
  double *pool_sm = reinterpret_cast<double *>(otile + 2 * K7_TILE);
  if (threadIdx.x == 0) {
    for (int s = 0; s < K7_NSB; ++s) gs_mbar_init(&bars->bwd[s], 1);
    for (int s = 0; s < 3; ++s) {
      gs_mbar_init(&bars->ft[s], NFW);
      gs_mbar_init(&bars->empty[s], 1);
      gs_mbar_init(&bars->pin[s], 1);
      gs_mbar_init(&bars->pdone[s], NFW);
    }
    bars->redo = 0;
    gs_mbar_init_fence();
    gs_tma_prefetch_desc(&tm.pred);
    gs_tma_prefetch_desc(&tm.rhs);
    if (FUSE) gs_tma_prefetch_desc(&tm.gold);
    if (XDIR) gs_tma_prefetch_desc(&tm.out);
  }
  if (!LIT)
    for (int i = threadIdx.x; i < q.stage_len; i += blockDim.x) pool_sm[i] = q.pool[q.stage_off + i];
  gs_fence_proxy_async();
  __syncthreads();
  if (warp >= 1) {
    // ---- faces warps ----
    uint32_t e_phase = 0, p_phase = 0, d_phase = 0;
    unsigned flags = 0;
    k7f_faces<XDIR, LIT, NFW>(q, tm, bars, region, pool_sm, e_phase, p_phase, d_phase, warp - 1, line0, lane, rl, flags);
    __syncthreads();
    if (bars->redo) {
      k7f_faces<XDIR, LIT, NFW>(q, tm, bars, region, pool_sm, e_phase, p_phase, d_phase, warp - 1, line0, lane, rl, flags);
      __syncthreads();
    }
    gs_raise(q.status, flags);
    return;
  }
  // ---- chain warp ----
  if (lane == 0)
    for (int s = 0; s < 3; ++s) gs_mbar_arrive(&bars->empty[s]);      // every stage starts empty
  bool fast_allowed = true;
  const GsRcp rt = gs_rcp_invariant(q.time, fast_allowed);
  const double inv_time = 1.0 / q.time;
  double2 *sc = q.scratch + (size_t)blockIdx.x * (size_t)n * 32;
  unsigned flags = 0;
  uint32_t ft_phase = 0;
  {
    bool ok = true;
    unsigned fl = 0;
    if (fast_allowed) {
      ok = k7f_forward<XDIR, true>(q, bars, region, ft_phase, line0, lane, rl, rt, inv_time, sc, fl);
      ok = __all_sync(0xffffffffu, ok);
    } else {
      k7f_forward<XDIR, false>(q, bars, region, ft_phase, line0, lane, rl, rt, inv_time, sc, fl);
    }
    if (lane == 0) bars->redo = ok ? 0 : 1;
    __syncthreads();
    if (!ok) {
      fl = 0;
      k7f_forward<XDIR, false>(q, bars, region, ft_phase, line0, lane, rl, rt, inv_time, sc, fl);
      __syncthreads();
    }
    flags |= fl;
  }
  __threadfence();
  gs_fence_proxy_async();
  __syncwarp();
  // backward ring: the forward region; the y sweeps also take the output tiles and the (now idle) spline pool
  K7Ring r;
  k7_bwd_ring<FUSE>(r, region, bars->bwd, K7F_REGION + (XDIR ? 0 : 2 * K7_TILE + K7F_POOL), q.nsb);
  const double u = k7_backward<XDIR, FUSE>(q, tm, r, otile, line0, act, lane, sc);
  if (XDIR && lane == 0) gs_bulk_wait_read<0>();
  if (act && !(fabs(u) <= 1.7976931348623157e308)) flags |= GS_FLAG_NONFINITE;
  gs_raise(q.status, flags);
}
