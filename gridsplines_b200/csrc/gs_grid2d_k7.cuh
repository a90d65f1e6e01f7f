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
//     fed by a per-warp ring of cp.async.bulk stages guarded by mbarriers:
//        forward : stage = 16 nodes x 32 lines of (F, rhs) + the nodes' (cf0, cf1);  (delta, lambda) = toPassion
//                  goes to a blocked scratch array [strip][node][lane] (512 B contiguous per node and warp);
//        backward: stage = 16 nodes of (delta, lambda) (ONE 8 KB bulk copy) + (fused y sweep) the old grid rows;
//                  x_i = delta_i * x_{i+1} + lambda_i, Grid.update (:474-481) applied on the way out (y), results
//                  of the x sweep leave through transposed shared-memory tiles and per-row bulk stores.
//   y sweep: lane <-> column (a node row of the strip is one 256 B segment);  x sweep: lane <-> row, a stage holds a
//   32 x 16 tile whose rows arrive as one 128 B bulk copy each (pitch 144 B: conflict-free 16-byte reads).
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
#define K7_XP 18                         // x tiles: doubles per row (16 + 2: 144 B pitch, 16 B aligned)
#define K7_STAGE 12288                   // bytes per stage = max(forward, backward) footprint
#define K7_OUT (32 * K7_XP * 8)          // one transposed output tile of the x sweep
#define K7_SMEM (128 + K7_NS * K7_STAGE + 2 * K7_OUT)

struct K7Params {
  int n, nlines, cols, rows;
  const double *coefs;             // Dim.coefs [2n]
  const double *F, *first3, *last3;   // k6_faces / k6_ends
  const double *rhs;               // x: grid, y: result
  double *out;                     // x, unfused y: result
  double *grid_io;                 // fused y: old grid in / new grid out
  double *predict_out;             // fused y: new predict
  double2 *scratch;                // [strips][n][32] (delta, lambda)
  double time;
  unsigned *status;
};

// ring bookkeeping of one warp: stage s is used for the chunks s, s + NS, ...; bit s of `phase` is the parity the
// next wait on its barrier expects
struct K7Ring {
  unsigned char *stages;
  uint64_t *bars;
  uint32_t phase;
  int head;      // stage of the next chunk to be consumed
  int tail;      // stage of the next chunk to be issued
};
GS_DEV unsigned char *k7_stage(const K7Ring &r, int s) { return r.stages + (size_t)s * K7_STAGE; }
GS_DEV void k7_wait(K7Ring &r) {
  gs_mbar_wait(&r.bars[r.head], (r.phase >> r.head) & 1u);
  r.phase ^= 1u << r.head;
}
GS_DEV void k7_advance(int &s) { s = (s == K7_NS - 1) ? 0 : s + 1; }

// forward stage layout:  y: F [16][32] | rhs [16][32] | cf [16][2];   x: F [32][18] | rhs [32][18] | cf [16][2]
template <bool XDIR>
struct K7Fwd {
  static constexpr int kTile = XDIR ? 32 * K7_XP * 8 : K7_CH * 256;
  static constexpr int kCf = 2 * kTile;
};

template <bool XDIR>
GS_DEV void k7_issue_fwd(const K7Params &q, K7Ring &r, int c, int line0, int nl, int lane) {
  const int i0 = c * K7_CH, len = min(K7_CH, q.n - i0);
  unsigned char *st = k7_stage(r, r.tail);
  uint64_t *bar = &r.bars[r.tail];
  if (lane == 0) gs_mbar_expect_tx(bar, (uint32_t)(2 * nl * len * 8 + len * 16));
  __syncwarp();
  if (XDIR) {
    if (lane < nl) {
      const size_t g = (size_t)(line0 + lane) * q.cols + i0;
      gs_bulk_g2s(st + lane * (K7_XP * 8), q.F + g, (uint32_t)len * 8u, bar);
      gs_bulk_g2s(st + K7Fwd<true>::kTile + lane * (K7_XP * 8), q.rhs + g, (uint32_t)len * 8u, bar);
    }
  } else {
    const int j = lane & (K7_CH - 1);
    if (j < len) {
      const size_t g = (size_t)(i0 + j) * q.cols + line0;
      if (lane < K7_CH) gs_bulk_g2s(st + j * 256, q.F + g, (uint32_t)nl * 8u, bar);
      else gs_bulk_g2s(st + K7Fwd<false>::kTile + j * 256, q.rhs + g, (uint32_t)nl * 8u, bar);
    }
  }
  if (lane == 0) gs_bulk_g2s(st + K7Fwd<XDIR>::kCf, q.coefs + 2 * (size_t)i0, (uint32_t)len * 16u, bar);
  k7_advance(r.tail);
}

// The forward walk of one line (lane) over all chunks; toPassion -> scratch.  Returns the sticky range predicate of
// the shared-reciprocal division (always true when !FAST).
template <bool XDIR, bool FAST>
GS_DEV bool k7_forward(const K7Params &q, K7Ring &r, int line0, int nl, int lane, int rl, GsRcp rt,
                                        double inv_time, double2 *sc, unsigned &flags) {
  const int n = q.n, NC = (n + K7_CH - 1) / K7_CH;
  const size_t line = (size_t)(line0 + rl);
  const double f_diag = q.first3[3 * line], f_sup = q.first3[3 * line + 1], f_rhs = q.first3[3 * line + 2];
  const double l_sub = q.last3[3 * line], l_diag = q.last3[3 * line + 1], l_rhs = q.last3[3 * line + 2];
  bool ok = true;
  double delta = 0.0, lambda = 0.0, fprev = 0.0;
  for (int c = 0; c < min(K7_NS, NC); ++c) k7_issue_fwd<XDIR>(q, r, c, line0, nl, lane);
  // Dim.general :67-83 + forwardUnit :257-265 of an interior node
  auto node_mid = [&](double t, double F, double cf0, double cf1) {
    const double a = cf0 * fprev;
    const double cc = cf1 * F;
    const double vect = gs_div<FAST>(-t, rt, ok);
    gs_forward_unit<FAST>(a, -(a + cc) - inv_time, cc, vect, delta, lambda, flags, ok);
    fprev = F;
  };
#pragma unroll 1
  for (int c = 0; c < NC; ++c) {
    const int i0 = c * K7_CH, len = min(K7_CH, n - i0);
    k7_wait(r);
    const unsigned char *st = k7_stage(r, r.head);
    const double2 *cf = reinterpret_cast<const double2 *>(st + K7Fwd<XDIR>::kCf);
    double2 *so = sc + (size_t)i0 * 32 + lane;
    if (c != 0 && i0 + K7_CH < n) {
      // a full chunk holding neither node 0 nor node n - 1
      if (XDIR) {
        const double2 *rF = reinterpret_cast<const double2 *>(st + rl * (K7_XP * 8));
        const double2 *rT = reinterpret_cast<const double2 *>(st + K7Fwd<true>::kTile + rl * (K7_XP * 8));
#pragma unroll
        for (int j = 0; j < K7_CH; j += 2) {
          const double2 f2 = rF[j >> 1], t2 = rT[j >> 1];
          node_mid(t2.x, f2.x, cf[j].x, cf[j].y);
          so[(size_t)j * 32] = make_double2(delta, lambda);
          node_mid(t2.y, f2.y, cf[j + 1].x, cf[j + 1].y);
          so[(size_t)(j + 1) * 32] = make_double2(delta, lambda);
        }
      } else {
        const double *cF = reinterpret_cast<const double *>(st) + rl;
        const double *cT = reinterpret_cast<const double *>(st + K7Fwd<false>::kTile) + rl;
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
          F = reinterpret_cast<const double *>(st)[rl * K7_XP + j];
          t = reinterpret_cast<const double *>(st + K7Fwd<true>::kTile)[rl * K7_XP + j];
        } else {
          F = reinterpret_cast<const double *>(st)[j * 32 + rl];
          t = reinterpret_cast<const double *>(st + K7Fwd<false>::kTile)[j * 32 + rl];
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
    k7_advance(r.head);
    __syncwarp();
    if (c + K7_NS < NC) k7_issue_fwd<XDIR>(q, r, c + K7_NS, line0, nl, lane);
  }
  return ok;
}

// backward stage layout: (delta, lambda) [16][32] double2 (8 KB) | old grid [16][32] (fused y sweep)
template <bool XDIR, bool FUSE>
GS_DEV void k7_issue_bwd(const K7Params &q, K7Ring &r, int c, int line0, int nl, int lane, const double2 *sc) {
  const int i0 = c * K7_CH, len = min(K7_CH, q.n - i0);
  unsigned char *st = k7_stage(r, r.tail);
  uint64_t *bar = &r.bars[r.tail];
  if (lane == 0) gs_mbar_expect_tx(bar, (uint32_t)(len * 512 + (FUSE ? len * nl * 8 : 0)));
  __syncwarp();
  if (lane == 0) gs_bulk_g2s(st, sc + (size_t)i0 * 32, (uint32_t)len * 512u, bar);
  if (FUSE && lane < len)
    gs_bulk_g2s(st + K7_CH * 512 + lane * 256, q.grid_io + (size_t)(i0 + lane) * q.cols + line0, (uint32_t)nl * 8u, bar);
  k7_advance(r.tail);
}

template <bool XDIR, bool FUSE>
__global__ void __launch_bounds__(32) k7_sweep(K7Params q) {
  extern __shared__ __align__(128) unsigned char k7s[];
  const int lane = threadIdx.x;
  const int n = q.n, NC = (n + K7_CH - 1) / K7_CH;
  const int line0 = blockIdx.x * 32, nl = min(32, q.nlines - line0);
  const bool act = lane < nl;
  const int rl = act ? lane : nl - 1;      // lanes past the last line shadow it (their stores are suppressed)
  K7Ring r;
  r.bars = reinterpret_cast<uint64_t *>(k7s);
  r.stages = k7s + 128;
  r.phase = 0;
  r.head = r.tail = 0;
  unsigned char *otile = k7s + 128 + K7_NS * K7_STAGE;
  if (lane == 0) {
    for (int s = 0; s < K7_NS; ++s) gs_mbar_init(&r.bars[s], 1);
    gs_mbar_init_fence();
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
      ok = k7_forward<XDIR, true>(q, r, line0, nl, lane, rl, rt, inv_time, sc, fl);
      ok = __all_sync(0xffffffffu, ok);
    }
    if (!ok) {
      // an operand outside the shared-reciprocal divider's range (zero, denormal, Inf, NaN, absurd magnitude) somewhere
      // in this warp's lines: the same walk with the native division
      fl = 0;
      __syncwarp();
      k7_forward<XDIR, false>(q, r, line0, nl, lane, rl, rt, inv_time, sc, fl);
    }
    flags |= fl;
  }
  // the (delta, lambda) this warp wrote (generic proxy) are read back by bulk copies (async proxy)
  __threadfence();
  gs_fence_proxy_async();
  __syncwarp();

  // ---- backward (backwardRow :303-322 / backwardColumn :324-342 [+ Grid.update :474-481]) ----
  for (int k = 0; k < min(K7_NS, NC); ++k) k7_issue_bwd<XDIR, FUSE>(q, r, NC - 1 - k, line0, nl, lane, sc);
  double u = 0.0;
  int ot = 0;
#pragma unroll 1
  for (int c = NC - 1; c >= 0; --c) {
    const int i0 = c * K7_CH, len = min(K7_CH, n - i0);
    k7_wait(r);
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
      // this lane's row of the output tile: written here, read by this lane's own bulk store
      double *orow = reinterpret_cast<double *>(otile + ot * K7_OUT) + lane * K7_XP;
      gs_bulk_wait_read<1>();            // the store issued from this tile two chunks ago has read it
#pragma unroll
      for (int j = 0; j < K7_CH; j += 2) *reinterpret_cast<double2 *>(orow + j) = make_double2(xs[j], xs[j + 1]);
      gs_fence_proxy_async_smem();
      if (act) gs_bulk_s2g(q.out + (size_t)(line0 + lane) * q.cols + i0, orow, (uint32_t)len * 8u);
      gs_bulk_commit();
      ot ^= 1;
    } else if (act) {
      const size_t g0 = (size_t)i0 * q.cols + line0 + lane;
      if (FUSE) {
        const double *gold = reinterpret_cast<const double *>(st + K7_CH * 512) + lane;
        double *G = q.grid_io + g0, *P = q.predict_out + g0;
#pragma unroll
        for (int j = 0; j < K7_CH; ++j) {
          if (j < len) {
            P[(size_t)j * q.cols] = aval2d(gold[j * 32], xs[j]);
            G[(size_t)j * q.cols] = xs[j];
          }
        }
      } else {
        double *O = q.out + g0;
#pragma unroll
        for (int j = 0; j < K7_CH; ++j)
          if (j < len) O[(size_t)j * q.cols] = xs[j];
      }
    }
    k7_advance(r.head);
    __syncwarp();
    if (c - K7_NS >= 0) k7_issue_bwd<XDIR, FUSE>(q, r, c - K7_NS, line0, nl, lane, sc);
  }
  if (XDIR) gs_bulk_wait_read<0>();
  if (act && !(fabs(u) <= 1.7976931348623157e308)) flags |= GS_FLAG_NONFINITE;
  gs_raise(q.status, flags);
}
