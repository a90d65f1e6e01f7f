// gs_grid1d_k2t.cuh -- K2T: batched 1-D lines with a temperature-dependent table (passion.iteration overload 1,
// A/passion/package.scala:52-95, + OneDGrid.step's predictor A/OneDGrid.scala:26-34), BIT-EXACT, built like the 2-D K7F:
//
//   * warp-specialised CTAs of two warps per 32-line block.  The FACES warp evaluates cons() = z((t_i + t_{i+1}) / 2.0)
//     (:21-25) for 16 nodes at a time out of a bulk-copied predict tile -- eight independent table lookups in flight -- and
//     issues the chain warp's loads (grid tile, the nodes' preDef pairs); the CHAIN warp runs ONE THREAD PER LINE in the
//     reference's operation order (faces_node: the very operations of K2) and never waits for a spline or a load of its own.
//   * PERSISTENT CTAs walk the line blocks and REUSE one (delta, lambda) scratch block each, so the scratch footprint is
//     CTAs x n x 512 B instead of K2's 24 warps per SM x n x 512 B and the backward walk reads it back from the 126 MB L2
//     (tools/l2_scratch.cu is the microbenchmark behind the idea); every chunk of it is DISCARDED from L2
//     (discard.global.L2, SASS CCTL.E.RML2) as soon as the backward walk has it in shared memory, so the dead dirty lines
//     are dropped instead of written back.  ncu, C2 shape with an M1Hermite3 table, DRAM bytes per node read + written:
//     four CTAs per SM 24.9 + 28.3 without the discards, 22.4 + 18.9 = 41 with them; six 31.7 + 31.3 -> 24.7 + 22.2 = 47
//     (K2: 70).  The step's own arrays travel with evict_first.
//   * the backward walk (backwardPassion :345-365, predictor arraygrid.aVal :85-87, grid <- result) is the chain warp's:
//     stages of 16 nodes of (delta, lambda) + the old grid tile through a bulk-copy ring that reuses the forward stages.
//   * the hot loops hold the straight-line tiers only (fallbacks out of line, native-division walk rolled): with so few
//     warps per SM an instruction-cache miss is not hidden -- the first build (20 k SASS instructions) spent 25 % of its
//     stall samples on no_instructions; this one has 4.4 k.
// MEASURED (B200, 1 M lines x 256 nodes, M1Hermite3 table), ms per step: K2 3.43; K2T fully unrolled over the 16 nodes of a
// chunk, without discards, CTAs per SM 3 / 4 / 5 / 6 / 7 / 8: 4.19 / 3.31 / 3.42 / 3.13 / 3.16 / 3.12; with discards 4 / 6 / 7 /
// 8: 3.18 / 2.95 / 3.17 / 2.86 (seven puts the warps unevenly on the four schedulers); with the chain's loop unrolled by 4 and the
// backward walk rolled over quarters of a chunk (3.8 k -> 3.0 k SASS instructions, 128 registers without spills, still 17 % of
// the stall samples on no_instructions before): eight CTAs per SM **2.57** (six: 2.74; chain unrolled by 2: 2.60).  ~155 warp
// instructions per node-row against K2's 308.  Forward walk alone 1.7 - 1.8 ms, backward walk 1.3 - 1.4 ms (first build);
// the two do not overlap inside a CTA, so more CTAs per SM is what helps.  Small batches are latency-bound and gain most:
// 4096 x 256 0.056 ms per step against K2's 0.217, 100 x 200 0.045 against 0.170 (profiles/r2_k2t_small_batches.txt).
// Shared memory per CTA: barriers 512 B | NS stages x 12.5 KB (NS = 2 from six CTAs per SM, else 3) | the table (<= 4 KB; eight
// CTAs per SM fit while it is below ~1.5 KB).
// Applies to a SINGLE table of at most 4 KB on lines of at least 32 nodes; everything else stays on K2.
#pragma once

#define K2T_CH 16
#define K2T_TILE (K2T_CH * 256)                       // 16 nodes x 32 lines of one array
#define K2T_OFF_F K2T_TILE                            // forward stage: grid tile | F tile | preDef pairs | predict tile (+1 node)
#define K2T_OFF_CF (2 * K2T_TILE)
#define K2T_OFF_P (2 * K2T_TILE + 256)
#define K2T_FSTAGE (3 * K2T_TILE + 512)               // 12800
#define K2T_BSTAGE (K2T_CH * 512 + K2T_TILE)          // backward stage: (delta, lambda) [16][32] | old grid tile  = 12288
#define K2T_REGION(NS) ((NS) * K2T_FSTAGE)
#define K2T_POOL 4096
#define K2T_BARS 512
#define K2T_SMEM(NS) (128 + K2T_BARS + K2T_REGION(NS) + K2T_POOL)
#define K2T_NS(CPS) ((CPS) >= 6 ? 2 : 3)     // ring depth: six and more CTAs per SM fit with two stages
static_assert(K2T_BSTAGE <= K2T_FSTAGE, "K2T backward ring inside the forward region");

struct K2TBars {
  uint64_t ft[3];      // grid tile + preDef pairs landed (transaction bytes) and the F tile written (faces warp's arrival)
  uint64_t empty[3];   // the chain warp is done with the forward stage
  uint64_t pin[3];     // predict tile landed
  uint64_t bwd[3];     // backward stage landed
  int redo;            // the forward walk of this block is repeated with the native division
  int pad;
  double face0[32];    // z((firstT + t_0) / 2.0) per line: node 0's left face
};
static_assert(sizeof(K2TBars) <= K2T_BARS, "K2T barrier block");

// gs_tab_apply_fast (AlwaysDefinedSpline.apply for a cubic table on equally spaced knots, argument strictly inside the
// clamps: the same double as gs_ads_apply) with the one-step fix-up of the arithmetic piece index written with selects,
// as k6_face_fast does: a data-dependent branch would keep the compiler from interleaving the lookups of a batch.
GS_DEV bool k2t_apply_fast(const GsTab &z, double x, double &v) {
  bool good = z.fast && (x > z.lowerX) && (x < z.upperX) && (x >= z.k0);
  int i = (int)((x - z.k0) * z.inv_h);          // saturating conversion; NaN -> 0
  i = max(0, min(z.np - 1, i));
  const double km = z.knots[max(i - 1, 0)], k1 = z.knots[i], k2 = z.knots[i + 1], kc = z.knots[min(i + 2, z.np)];
  const bool down = x < k1, up = !down && (x >= k2);
  i = down ? max(i - 1, 0) : (up ? min(i + 1, z.np - 1) : i);
  const double ka = down ? km : (up ? k2 : k1), kb = down ? k1 : (up ? kc : k2);
  good = good && (x >= ka) && (x < kb);
  const double *pc = z.pieces + i * GS_PIECE;
  v = gs_cubic_horner(x - pc[0], pc[1], pc[2], pc[3], pc[4]);
  return good;
}

// Out-of-line fallbacks: with eight warps per SM nothing hides an instruction-cache miss (the first build, with these
// inlined 16 times each, showed 25 % of its stall samples as no_instructions), so the hot loops hold the straight-line
// tiers only.
__device__ __noinline__ double k2t_apply_general(const double *pool_sm, double x, unsigned *flags) {
  unsigned f = 0;
  const double v = gs_ads_apply(gs_spline_view(pool_sm, 0), x, f);
  *flags |= f;
  return v;
}
__device__ __noinline__ double k2t_div3(double d) { return d / 3.0; }
// the 128-byte line at p is dead: L2 may drop it instead of writing it back (p 128-byte aligned)
__device__ __forceinline__ void k2t_discard(const void *p) {
  asm volatile("discard.global.L2 [%0], 128;" ::"l"(p) : "memory");
}
__device__ __forceinline__ void k2t_st_first(double *p, double v, uint64_t pol) {
  asm volatile("st.global.L2::cache_hint.f64 [%0], %1, %2;" ::"l"(p), "d"(v), "l"(pol) : "memory");
}

// The faces warp's pass over one line block.
template <int NS>
GS_DEV void k2t_faces(const Step1dParams &p, K2TBars *bars, unsigned char *region, const double *pool_sm,
                      uint32_t &e_phase, uint32_t &p_phase, const double *Gblk, const double *Pblk, int lane, double firstT,
                      double lastT, uint64_t pol, int hints, unsigned &flags) {
  const int n = p.n, NC = (n + K2T_CH - 1) / K2T_CH;
  const GsSpline z = gs_spline_view(pool_sm, 0);
  const GsTab zt = gs_tab_view(pool_sm);
  auto issue_p = [&](int c) {
    if (lane == 0) {
      const int s = c % NS, i0 = c * K2T_CH;
      const int rows = min(K2T_CH + 1, n - i0);        // one look-ahead node, except at the end of the line
      gs_mbar_expect_tx(&bars->pin[s], (uint32_t)rows * 256u);
      unsigned char *dst = region + s * K2T_FSTAGE + K2T_OFF_P;
      if (hints & 2) gs_bulk_g2s_hint(dst, Pblk + (size_t)i0 * 32, (uint32_t)rows * 256u, &bars->pin[s], pol);
      else gs_bulk_g2s(dst, Pblk + (size_t)i0 * 32, (uint32_t)rows * 256u, &bars->pin[s]);
    }
  };
  for (int c = 0; c < min(NS, NC); ++c) issue_p(c);
  int s = 0;
#pragma unroll 1
  for (int c = 0; c < NC; ++c) {
    const int i0 = c * K2T_CH, len = min(K2T_CH, n - i0);
    // the stage is free again (the chain warp consumed chunk c - 3): start the chain warp's loads into it
    gs_mbar_wait_bounded(&bars->empty[s], (e_phase >> s) & 1u, p.status);
    e_phase ^= 1u << s;
    unsigned char *st = region + s * K2T_FSTAGE;
    if (lane == 0) {
      gs_mbar_expect_tx_only(&bars->ft[s], (uint32_t)len * (256u + 16u));
      if (hints & 1) gs_bulk_g2s_hint(st, Gblk + (size_t)i0 * 32, (uint32_t)len * 256u, &bars->ft[s], pol);
      else gs_bulk_g2s(st, Gblk + (size_t)i0 * 32, (uint32_t)len * 256u, &bars->ft[s]);
      gs_bulk_g2s(st + K2T_OFF_CF, p.predef + 2 * (size_t)i0, (uint32_t)len * 16u, &bars->ft[s]);
    }
    gs_mbar_wait_bounded(&bars->pin[s], (p_phase >> s) & 1u, p.status);
    p_phase ^= 1u << s;
    double *P = reinterpret_cast<double *>(st + K2T_OFF_P) + lane;
    double *F = reinterpret_cast<double *>(st + K2T_OFF_F) + lane;
    if (c == 0) bars->face0[lane] = gs_ads_apply(z, zt, (firstT + P[0]) / 2.0, flags);
    if (i0 + K2T_CH == n) P[K2T_CH * 32] = lastT;      // the last node's right neighbour is the line's end temperature (:84-86)
    if (len == K2T_CH) {
      // eight lookups at a time through the straight-line tier (independent instruction streams); whatever leaves it
      // -- clamps, unevenly spaced knots, Line / Const pieces, NaN -- is redone through the general evaluation
#pragma unroll 1
      for (int h = 0; h < K2T_CH; h += 8) {
        double v[8], x[8];
        unsigned bad = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          x[j] = (P[(h + j) * 32] + P[(h + j + 1) * 32]) / 2.0;
          if (!k2t_apply_fast(zt, x[j], v[j])) bad |= 1u << j;
        }
        if (bad) {
#pragma unroll
          for (int j = 0; j < 8; ++j)
            if ((bad >> j) & 1u) v[j] = k2t_apply_general(pool_sm, x[j], &flags);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) F[(h + j) * 32] = v[j];
      }
    } else {
#pragma unroll 1
      for (int j = 0; j < len; ++j) {
        const double t3 = (i0 + j == n - 1) ? lastT : P[(j + 1) * 32];
        F[j * 32] = gs_ads_apply(z, zt, (P[j * 32] + t3) / 2.0, flags);
      }
    }
    __syncwarp();
    if (lane == 0) gs_mbar_arrive(&bars->ft[s]);       // release: the F tile (and face0) are visible
    if (c + NS < NC) issue_p(c + NS);                    // this warp was the predict tile's only reader
    s = (s == NS - 1) ? 0 : s + 1;
  }
}

// The chain warp's forward walk over one line block: faces_node per node, (delta, lambda) -> this CTA's scratch block.
template <bool FAST, int NS>
GS_DEV bool k2t_forward(const Line1d &L, const Step1dParams &p, K2TBars *bars, unsigned char *region, uint32_t &ft_phase,
                        int lane, double firstT, double lastT, double2 *sc, unsigned &flags) {
  const int n = p.n, NC = (n + K2T_CH - 1) / K2T_CH;
  bool ok = true;
  double delta = 0.0, lambda = 0.0, fprev = 0.0;
  int s = 0;
#pragma unroll 1
  for (int c = 0; c < NC; ++c) {
    const int i0 = c * K2T_CH, len = min(K2T_CH, n - i0);
    gs_mbar_wait_bounded(&bars->ft[s], (ft_phase >> s) & 1u, p.status);
    ft_phase ^= 1u << s;
    const unsigned char *st = region + s * K2T_FSTAGE;
    const double *cG = reinterpret_cast<const double *>(st) + lane;
    const double *cF = reinterpret_cast<const double *>(st + K2T_OFF_F) + lane;
    const double2 *cf = reinterpret_cast<const double2 *>(st + K2T_OFF_CF);
    double2 *so = sc + (size_t)i0 * 32 + lane;
    if (FAST && c != 0 && i0 + K2T_CH < n) {
#pragma unroll 4
      for (int j = 0; j < K2T_CH; ++j) {
        const double F = cF[j * 32];
        const double a = fprev * cf[j].x;
        const double cc = F * cf[j].y;
        const double vect = gs_div<FAST, true>(-cG[j * 32], L.rt, ok);
        gs_forward_unit<FAST>(a, -(a + cc) - L.inv_time, cc, vect, delta, lambda, flags, ok);
        fprev = F;
        so[(size_t)j * 32] = make_double2(delta, lambda);
      }
    } else {
      // the chunks with the line's ends (:60-67, :84-94), and every chunk of the native-division walk
      if (c == 0) fprev = bars->face0[lane];
#pragma unroll 1
      for (int j = 0; j < len; ++j) {
        const double F = cF[j * 32];
        faces_node<FAST>(L, i0 + j, cG[j * 32], fprev, F, firstT, lastT, delta, lambda, flags, ok);
        fprev = F;
        so[(size_t)j * 32] = make_double2(delta, lambda);
      }
    }
    __syncwarp();
    if (lane == 0) gs_mbar_arrive(&bars->empty[s]);
    s = (s == NS - 1) ? 0 : s + 1;
  }
  return ok;
}

struct K2TRing {
  uint32_t phase;
  int head, tail;
};

template <int CPS>
__global__ void __launch_bounds__(64, CPS) k2t_step(Step1dParams p, int stage_len, int hints) {
  constexpr int NS = K2T_NS(CPS);
  extern __shared__ unsigned char k2traw[];
  unsigned char *sm = k2traw + ((128u - (gs_smem_u32(k2traw) & 127u)) & 127u);
  K2TBars *bars = reinterpret_cast<K2TBars *>(sm);
  unsigned char *region = sm + K2T_BARS;
  double *pool_sm = reinterpret_cast<double *>(region + K2T_REGION(NS));
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int n = p.n, NC = (n + K2T_CH - 1) / K2T_CH;
  if (threadIdx.x == 0) {
    for (int s = 0; s < 3; ++s) {
      gs_mbar_init(&bars->ft[s], 1);
      gs_mbar_init(&bars->empty[s], 1);
      gs_mbar_init(&bars->pin[s], 1);
      gs_mbar_init(&bars->bwd[s], 1);
    }
    bars->redo = 0;
    gs_mbar_init_fence();
  }
  for (int i = threadIdx.x; i < stage_len; i += blockDim.x) pool_sm[i] = p.pool[p.single_off + i];
  gs_fence_proxy_async();
  __syncthreads();
  const size_t block_elems = (size_t)n * GS_LB;
  double2 *sc = p.scratch + (size_t)blockIdx.x * block_elems;
  // the step's own arrays stream through L2: evict_first (hints: 1 grid forward, 2 predict, 4 grid backward loads; 8 the stores of grid' / predict'; 16: scratch lines discarded once read back) leaves the
  // L2 to the reused scratch blocks
  const uint64_t pol = gs_policy_evict_first();
  unsigned flags = 0;

  if (warp == 1) {
    // ---- faces warp ----
    uint32_t e_phase = 0, p_phase = 0;
    for (long long blk = p.blk0 + blockIdx.x; blk < p.nblocks; blk += gridDim.x) {
      const long long line = blk * GS_LB + lane;
      const bool valid = line < p.nlines;
      const long long bcl = (valid ? line : p.nlines - 1) * p.bc_stride;
      const double firstT = p.leftT[bcl], lastT = p.rightT[bcl];
      const double *Gblk = p.grid + (size_t)blk * block_elems, *Pblk = p.predict + (size_t)blk * block_elems;
      unsigned lf = 0;
      do {
        k2t_faces<NS>(p, bars, region, pool_sm, e_phase, p_phase, Gblk, Pblk, lane, firstT, lastT, pol, hints, lf);
        __syncthreads();                                 // A: forward walk done, `redo` decided (A': it is 0 again)
      } while (bars->redo);
      if (valid) flags |= lf;
      __syncthreads();                                   // B: backward walk done, the region is free again
    }
    gs_raise(p.status, flags);
    return;
  }

  // ---- chain warp ----
  if (lane == 0)
    for (int s = 0; s < NS; ++s) gs_mbar_arrive(&bars->empty[s]);     // every stage starts empty
  Line1d L;
  L.n = n;
  L.pd = p.predef;
  L.pool = p.pool;
  L.node_off = p.node_off;
  bool fast_allowed = p.allow_fast != 0;
  L.rt = gs_rcp_invariant(p.time, fast_allowed);
  const GsRcp r3 = gs_rcp_invariant(3.0, fast_allowed);
  L.inv_time = 1.0 / p.time;
  uint32_t ft_phase = 0;
  K2TRing r;
  r.phase = 0;
  r.head = r.tail = 0;
  for (long long blk = p.blk0 + blockIdx.x; blk < p.nblocks; blk += gridDim.x) {
    const long long line = blk * GS_LB + lane;
    const bool valid = line < p.nlines;      // padding lanes compute on the padding lines (kept finite)
    const long long bcl = (valid ? line : p.nlines - 1) * p.bc_stride;
    const double firstT = p.leftT[bcl], lastT = p.rightT[bcl];
    double *Gblk = p.grid + (size_t)blk * block_elems, *Pblk = p.predict + (size_t)blk * block_elems;
    unsigned lf = 0;
    {
      // operands outside the proven range of the shared-reciprocal division (or a time step outside it): the block is
      // (re)done with plain division
      bool ok = false;
      if (fast_allowed) {
        ok = k2t_forward<true, NS>(L, p, bars, region, ft_phase, lane, firstT, lastT, sc, lf);
        ok = __all_sync(0xffffffffu, ok);
        if (lane == 0) bars->redo = ok ? 0 : 1;
        __syncthreads();                                 // A
      }
      if (!ok) {
        lf = 0;
        k2t_forward<false, NS>(L, p, bars, region, ft_phase, lane, firstT, lastT, sc, lf);
        if (lane == 0) bars->redo = 0;
        __syncthreads();                                 // A (time step outside the fast range) or A'
      }
    }
    // the scratch stores (generic proxy) must be visible to the bulk copies that bring them back
    __threadfence();
    gs_fence_proxy_async();
    __syncwarp();
    // ---- backward walk ----
    auto issue_b = [&](int c) {
      if (lane == 0) {
        const int i0 = c * K2T_CH, len = min(K2T_CH, n - i0);
        unsigned char *st = region + r.tail * K2T_BSTAGE;
        uint64_t *bar = &bars->bwd[r.tail];
        gs_mbar_expect_tx(bar, (uint32_t)len * (512u + 256u));
        gs_bulk_g2s(st, sc + (size_t)i0 * 32, (uint32_t)len * 512u, bar);
        if (hints & 4) gs_bulk_g2s_hint(st + K2T_CH * 512, Gblk + (size_t)i0 * 32, (uint32_t)len * 256u, bar, pol);
        else gs_bulk_g2s(st + K2T_CH * 512, Gblk + (size_t)i0 * 32, (uint32_t)len * 256u, bar);
      }
      r.tail = (r.tail == NS - 1) ? 0 : r.tail + 1;
    };
    for (int k = 0; k < min(NS, NC); ++k) issue_b(NC - 1 - k);
    double u = 0.0;
#pragma unroll 1
    for (int c = NC - 1; c >= 0; --c) {
      const int i0 = c * K2T_CH, len = min(K2T_CH, n - i0);
      gs_mbar_wait_bounded(&bars->bwd[r.head], (r.phase >> r.head) & 1u, p.status);
      r.phase ^= 1u << r.head;
      const unsigned char *st = region + r.head * K2T_BSTAGE;
      const double2 *dl = reinterpret_cast<const double2 *>(st) + lane;
      const double *gold = reinterpret_cast<const double *>(st + K2T_CH * 512) + lane;
      if (hints & 16) {
        // this chunk of (delta, lambda) is in shared memory now and dead in global memory until the next block's forward
        // walk overwrites it: without this L2 writes the dirty lines back when it evicts them (12 B/node of DRAM writes)
        const unsigned char *dead = reinterpret_cast<const unsigned char *>(sc + (size_t)i0 * 32);
        for (int l = lane; l < len * 4; l += 32) k2t_discard(dead + (size_t)l * 128);
      }
      double *G = Gblk + (size_t)i0 * 32 + lane, *P = Pblk + (size_t)i0 * 32 + lane;
      if (len == K2T_CH) {
        // quarters of the chunk: the recurrence first (two dependent FP64 operations per node), then the quarter's
        // predictor and stores as independent work; rolled over the quarters to keep the loop in the instruction cache
#pragma unroll 1
        for (int h = K2T_CH - 4; h >= 0; h -= 4) {
          double xs[4];
#pragma unroll
          for (int j = 3; j >= 0; --j) {
            const double2 v = dl[(h + j) * 32];
            u = v.x * u + v.y;
            xs[j] = u;
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const double d = xs[j] - gold[(h + j) * 32];
            bool okd = true;
            double q = gs_div<true>(d, r3, okd);
            if (!(okd && fast_allowed)) q = k2t_div3(d);
            if (hints & 8) {
              k2t_st_first(P + (h + j) * 32, xs[j] + q, pol);    // arraygrid.aVal :85-87
              k2t_st_first(G + (h + j) * 32, xs[j], pol);
            } else {
              P[(h + j) * 32] = xs[j] + q;
              G[(h + j) * 32] = xs[j];
            }
          }
        }
      } else {
#pragma unroll 1
        for (int j = len - 1; j >= 0; --j) {
          const double2 v = dl[j * 32];
          u = v.x * u + v.y;
          const double d = u - gold[j * 32];
          bool okd = true;
          double q = gs_div<true>(d, r3, okd);
          if (!(okd && fast_allowed)) q = k2t_div3(d);
          P[j * 32] = u + q;
          G[j * 32] = u;
        }
      }
      r.head = (r.head == NS - 1) ? 0 : r.head + 1;
      __syncwarp();
      if (c - NS >= 0) issue_b(c - NS);
    }
    if (!(fabs(u) <= 1.7976931348623157e308)) lf |= GS_FLAG_NONFINITE;
    if (valid) flags |= lf;
    __syncthreads();                                     // B
  }
  gs_raise(p.status, flags);
}
