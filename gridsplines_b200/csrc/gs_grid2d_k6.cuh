// gs_grid2d_k6.cuh -- K6: 2-D sweeps for temperature-dependent properties (spline tables, any coefficient map):
// assembly is split from elimination.  Tolerance solver family (GS_SOLVER_PARTITIONED / AUTO): <= 1e-12 per field.
//
// The expensive part of the reference's row assembly is the material property at every cell face: an interval
// average of a spline table (A/Dim.scala:67-83 -> A/AlwaysDefinedSpline.scala:43-46), ~100+ instructions with a
// division.  It depends on `predict` only, which is frozen during a step, so it is NOT part of the Thomas
// recurrence: k6_faces evaluates every interior face of a direction in one fully parallel, coalesced pass
//     F[i] = co(predict_i, predict_{i+1}, z_i)        (Dim.general's `co`; the carry out of the first node at i = 0)
// with the reference's own formulas (same doubles); k6_ends assembles the two end rows of every line (Dirichlet or
// heat-flow, A/Dim.scala:28-64, 86-128) into small per-line arrays.  k6_sweep then is K5's one-CTA-per-strip
// pipeline (phase A -> reduced system -> phase C, cp.async double-buffered chunks, checkpoints per chunk) whose
// per-node work is only the elimination:  sub = cf0 F[i-1], sup = cf1 F[i], diag = -(sub+sup) - 1/tau,
// Delta = diag + sub delta, one reciprocal, delta, lambda, omega and the composed map (P, Q, R) of the segment.
#pragma once
#include <cooperative_groups.h>

#define GS_MAX_SHARDS 16

// ---- faces -------------------------------------------------------------------------------------------------------
struct K6FacesParams {
  Sweep2dParams sw;      // predict, rhs, coefs, bounds, maps, pool, time, geometry of the swept direction
  double *F;             // [rows * cols] face values (stored at node i's position)
  double *first3;        // [nlines][3] diag, sup, rhs of node 0
  double *last3;         // [nlines][3] sub, diag, rhs of node n-1
  // row-sharded y sweep (a rank owns rows [r0, r1) of every column): an open end has no boundary row, its outer
  // neighbour is a halo row of predict held by the adjacent rank
  int open_first, open_last;
  const double *halo_prev, *halo_next;   // [nlines] predict of the row above node 0 / below node n-1
  double *fm1;                            // [nlines] face between the upper halo row and node 0
  double *face0;                          // K7F: the carry out of node 0 goes here [nlines] and F is not touched
};

// The two end rows of every line (Dirichlet or heat-flow, A/Dim.scala:28-64, 86-128) with the reference's general
// assembly: one thread per line.  Also writes the carry out of node 0 (F at i = 0) and F = 0 at i = n - 1.
template <bool XDIR, class SEL>
__global__ void __launch_bounds__(128) k6_ends(K6FacesParams q) {
  extern __shared__ double k6sm[];
  const Sweep2dParams &p = q.sw;
  const double *pool = stage_pool(p, k6sm);
  const int line = blockIdx.x * blockDim.x + threadIdx.x;
  if (line >= (XDIR ? p.rows : p.cols)) return;
  SEL sel(p.maps, pool, p.cols);
  SweepConsts k;
  k.fast_allowed = false;
  bool dummy = true;
  k.rt = gs_rcp_invariant(p.time, dummy);
  k.inv_time = 1.0 / p.time;
  k.firstVol = p.firstVol; k.fC = p.fC; k.lastVol = p.lastVol; k.lA = p.lA;
  unsigned flags = 0;
  bool ok = true;
  LineAcc<XDIR> A(p, line);
  const int n = A.n();
  if (q.open_first) {
    // face above node 0 (coefficients of row -1, like the reference's coefs(col, -1) for the face of the first row)
    q.fm1[line] = gs_take_average<false>(XDIR ? sel.get(GS_SEL_APPLY, -1, line) : sel.get(GS_SEL_APPLY, line, -1),
                                         q.halo_prev[line], A.pred(0), flags, ok);
  } else {
    double face = 0.0;
    Row4 r = asm_node<XDIR, SEL, false>(A, sel, k, 0, A.pred(0), A.pred(1), A.rhs(0), face, flags, ok);
    if (q.face0) q.face0[line] = face;     // carry into node 1: co, or cond / cap after a heat-flow end
    else q.F[A.idx(0)] = face;
    double *o = q.first3 + 3 * (size_t)line;
    o[0] = r.diag; o[1] = r.sup; o[2] = r.rhs;
  }
  if (q.open_last) {
    const int i = n - 1;
    q.F[A.idx(i)] = gs_take_average<false>(sel.get(GS_SEL_APPLY, A.col(i), A.row(i)), A.pred(i), q.halo_next[line],
                                           flags, ok);
  } else {
    const int i = n - 1;
    const double p0 = A.pred(i);
    double face = face_before<XDIR, SEL, false>(A, sel, k, i, A.pred(i - 1), p0, flags, ok);
    Row4 r = asm_node<XDIR, SEL, false>(A, sel, k, i, p0, 0.0, A.rhs(i), face, flags, ok);
    if (!q.face0) q.F[A.idx(i)] = 0.0;
    double *o = q.last3 + 3 * (size_t)line;
    o[0] = r.sub; o[1] = r.diag; o[2] = r.rhs;
  }
  if (flags) atomicOr(p.status, flags);
}

// (GsTab / gs_tab_view of gs_device.cuh: what the straight-line face path needs of one table)
typedef GsTab K6Tab;
GS_DEV K6Tab k6_tab(const double *h) { return gs_tab_view(h); }
// takeAverage for the common interior face: two different, ordered-able temperatures strictly inside the clamps and
// inside ONE cubic piece.  AlwaysDefinedSpline.average then is  (0.0 + (antider(hi) - antider(lo)) + 0.0) / (hi - lo)
// (gs_ads_average's one-piece case); the same doubles are produced here without a data-dependent branch.  Returns
// false -- and the caller takes the general walk -- for anything else (clamped, several pieces, equal, NaN, a
// quotient outside the fast divider's range).
GS_DEV bool k6_face_fast(const K6Tab &z, double p0, double p1, double &v) {
  const bool lt = p0 < p1;
  const double lo = lt ? p0 : p1, hi = lt ? p1 : p0;
  bool good = z.fast && (lo > z.lowerX) && (hi < z.upperX) && (lo < hi);   // (lo >= knots[0] follows from lo >= ka below)
  int i = (int)((lo - z.k0) * z.inv_h);          // saturating conversion; NaN -> 0
  i = max(0, min(z.np - 1, i));
  // one-step fix-up of the arithmetic piece index, written with selects (the four knots around it are loaded up front):
  // a data-dependent branch here would keep the compiler from interleaving the faces of a batch
  const double km = z.knots[max(i - 1, 0)], k1 = z.knots[i], k2 = z.knots[i + 1], kc = z.knots[min(i + 2, z.np)];
  const bool down = lo < k1, up = !down && (lo >= k2);
  i = down ? max(i - 1, 0) : (up ? min(i + 1, z.np - 1) : i);
  const double ka = down ? km : (up ? k2 : k1), kb = down ? k1 : (up ? kc : k2);
  good = good && (lo >= ka) && (hi < kb);
  const double *pc = z.pieces + i * GS_PIECE;
  const double x0 = pc[0], a1 = pc[1], a2 = pc[5], a3 = pc[6], a4 = pc[7];
  const double su = hi - x0, sl = lo - x0;
  const double Au = fma(fma(fma(fma(a4, su, a3), su, a2), su, a1), su, 0.0);   // gs_cubic_antider
  const double Al = fma(fma(fma(fma(a4, sl, a3), sl, a2), sl, a1), sl, 0.0);
  // the reference's 0.0 + (Au - Al) + 0.0 (lowerArea + middleArea + upperArea) equals Au - Al bit for bit: adding +0.0
  // changes nothing but -0.0, and Au - Al is never -0.0 (Au, Al end in fma(., ., +0.0), so neither is -0.0, and equal
  // operands give +0.0 in round-to-nearest)
  const double area = Au - Al;
  // An area of exactly +0.0 (neighbouring temperatures a few ulps apart in a nearly uniform field: the antiderivatives
  // cancel completely) is outside the shared-reciprocal divider's dividend range, but its quotient needs no divider:
  // +0.0 / (hi - lo) = +0.0, which is also what gs_div returns for it.  Without this such faces -- most faces of a field
  // that starts uniform -- went to the second tier one by one (512^2 tables from a uniform start: 520 us per step
  // against 165 on a random field).
  bool okd = true, okq = true;
  const GsRcp rd = gs_rcp<true>(hi - lo, okd);
  v = gs_div<true>(area, rd, okq);
  return good && okd && (okq || area == 0.0);
}
// Equal temperatures (a uniform region: the reference's average(lo, hi) with lo == hi is apply(lo),
// A/AlwaysDefinedSpline.scala:44): the point value by the straight-line table look-up of gs_device.cuh -- the same double
// as gs_ads_apply -- without the call into the second tier.  Identical bit patterns only (+0.0 / -0.0 go the long way).
GS_DEV bool k6_face_equal(const K6Tab &z, double p0, double p1, double &v) {
  return __double_as_longlong(p0) == __double_as_longlong(p1) && gs_tab_apply_fast(z, p0, v);
}
// Both temperatures at or beyond the same clamp (a field hotter than the table's upperX -- which the reference's
// greatestK puts one knot below the last one, SURVEY App. B Q1 -- or colder than lowerX): AlwaysDefinedSpline.area is
// (lowerArea + middleArea) + upperArea with middleArea = 0.0 and one of the others (hi - lo) * Y
// (A/AlwaysDefinedSpline.scala:20-41), the average that over (hi - lo); equal temperatures give apply = Y.  `h` is the
// table's header (lowerX, upperX, lowerY, upperY at 2 .. 5).  Same doubles as gs_ads_average, without the general walk.
GS_DEV bool k6_face_clamped(const double *h, double p0, double p1, double &v) {
  const bool lt = p0 < p1;
  const double lo = lt ? p0 : p1, hi = lt ? p1 : p0;
  const bool below = hi <= h[2], above = lo >= h[3];
  if (!(below || above) || (below && above)) return false;   // (NaN, or a degenerate table: the general walk)
  const double Y = below ? h[4] : h[5];
  if (p0 == p1) {
    if (__double_as_longlong(p0) != __double_as_longlong(p1)) return false;   // +0.0 / -0.0
    v = Y;
    return true;
  }
  const double d = hi - lo;
  const double area = below ? ((d * Y + 0.0) + 0.0) : ((0.0 + 0.0) + d * Y);
  v = area / d;
  return true;
}
// second tier: what the straight-line path leaves out but is still one or two table pieces strictly inside the clamps
// and the knots -- equal temperatures (average = apply), an interval that straddles ONE knot, and one-piece intervals
// of tables the first tier does not take (Line / Const pieces, unevenly spaced knots).  Follows gs_ads_average /
// gs_spline_area term by term (same doubles), native division.  Returns false for anything else.
__device__ __noinline__ bool k6_face_mid(const double *h, double p0, double p1, double *out) {
  const GsSpline s = gs_spline_view(h, 0);
  const bool eq = p0 == p1;
  const double lo = eq ? gs_jmin(p0, p1) : (p0 < p1 ? p0 : p1), hi = eq ? gs_jmax(p0, p1) : (p0 < p1 ? p1 : p0);
  if (!(lo > s.lowerX && hi < s.upperX && lo >= s.knots[0] && hi <= s.knots[s.np])) return false;
  const int i = gs_knot_floor(s, lo);
  const double *pc = s.pieces + i * GS_PIECE;
  if (eq) {
    *out = gs_piece_apply(s.kind, pc, lo);
    return true;
  }
  if (!(lo < hi)) return false;
  const double kb = s.knots[i + 1];
  double acc;
  if (hi < kb) {
    acc = 0.0 + gs_piece_area(s.kind, pc, lo, hi);                     // one piece
  } else {
    if (i + 2 > s.np || !(hi > kb && hi < s.knots[i + 2])) return false;
    acc = 0.0 + gs_piece_area(s.kind, pc, lo, kb);                     // [lo, k_{i+1}] of piece i
    acc = acc + gs_piece_area(s.kind, pc + GS_PIECE, kb, hi);          // [k_{i+1}, hi] of piece i + 1
  }
  acc = 0.0 + acc + 0.0;
  bool ok = true;
  *out = gs_div<false>(acc, gs_rcp<false>(hi - lo, ok), ok);
  return true;
}
// the general walk (cold): every case of the reference, native division
__device__ __noinline__ double k6_face_general(const double *pool, int off, double p0, double p1, unsigned *status) {
  unsigned flags = 0;
  bool ok = true;
  double v = gs_take_average<false>(gs_spline_view(pool, off), p0, p1, flags, ok);
  if (flags) atomicOr(status, flags);
  return v;
}

// Interior faces F[i], 1 <= i <= n - 2, of every line.  TAB (table-driven maps, pool staged in shared memory): every
// warp walks (row, 128-column group) tiles  t = warp + k * nwarps  in row-major order (four faces per lane and tile), the
// column group rotated by the row so that no warp keeps the same columns (a boundary layer along one edge makes those
// columns' faces take the slower tiers); the next tile's loads are in flight during a tile's arithmetic.  Otherwise (literal constants, oversized pools): blocks
// of 256 columns x a stride of rows with the generic evaluation.
template <bool XDIR, class SEL, bool TAB>
__global__ void __launch_bounds__(256, 3) k6_faces(K6FacesParams q) {
  extern __shared__ double k6sm[];
  const Sweep2dParams &p = q.sw;
  // an open first end has no boundary row: face 0 is an ordinary interior face
  const int c_lo = XDIR ? (q.open_first ? 0 : 1) : 0, c_hi = XDIR ? p.cols - 1 : p.cols;
  const int r_lo = XDIR ? 0 : (q.open_first ? 0 : 1), r_hi = XDIR ? p.rows : p.rows - 1;
  const size_t nb = XDIR ? 1 : (size_t)p.cols;
  if (TAB) {
    for (int i = threadIdx.x; i < p.pool_len; i += blockDim.x) k6sm[i] = p.pool[i];
    __syncthreads();
    constexpr bool single = SEL::single;
    constexpr int U = 4;                         // faces per lane per tile: a tile is one row x 32 * U columns
    const int lane = threadIdx.x & 31;
    const int W = gridDim.x * (blockDim.x >> 5), w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int ncg = (p.cols + 32 * U - 1) / (32 * U), nrows = r_hi - r_lo;
    const int dj = W % ncg, drow = W / ncg, drm = drow % ncg;
    // tile walker: j = t % ncg, row = t / ncg (relative to r_lo), rmod = row % ncg
    int j = w % ncg, row = w / ncg, rmod = row % ncg;
    int off = p.maps.single_off[GS_SEL_APPLY];
    K6Tab z = k6_tab(k6sm + off);
    // loads of the walker's tile (nothing for columns outside the interior), then one step of the walker
    auto fetch = [&](double (&a0)[U], double (&a1)[U], int (&cl)[U], int &rw) {
      int cg = j + rmod;
      if (cg >= ncg) cg -= ncg;
      rw = row;
      const int col0 = cg * (32 * U) + lane;
      const double *pp = p.predict + (size_t)(row + r_lo) * p.cols + col0;
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int col = col0 + 32 * u;
        const bool okc = row < nrows && col >= c_lo && col < c_hi;
        cl[u] = okc ? col : -1;
        a0[u] = a1[u] = 0.0;
        if (okc) { a0[u] = pp[32 * u]; a1[u] = pp[32 * u + nb]; }
      }
      j += dj;
      const int c = j >= ncg;
      j -= c ? ncg : 0;
      row += drow + c;
      rmod += drm + c;
      if (rmod >= ncg) rmod -= ncg;
    };
    double c0[U], c1[U], n0[U], n1[U];
    int ccol[U], ncol[U], crow, nrow;
    fetch(c0, c1, ccol, crow);
#pragma unroll 1
    while (crow < nrows) {
      fetch(n0, n1, ncol, nrow);                 // the next tile's loads fly during this tile's arithmetic
      double *frow = q.F + (size_t)(crow + r_lo) * p.cols;
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (ccol[u] >= 0) {
          if (!single) {
            off = map_off(p.maps, GS_SEL_APPLY, ccol[u], crow + r_lo, p.cols);
            z = k6_tab(k6sm + off);
          }
          double v;
          if (!k6_face_fast(z, c0[u], c1[u], v) && !k6_face_equal(z, c0[u], c1[u], v) && !k6_face_clamped(k6sm + off, c0[u], c1[u], v)) {
            if (!k6_face_mid(k6sm + off, c0[u], c1[u], &v)) v = k6_face_general(p.pool, off, c0[u], c1[u], p.status);
          }
          frow[ccol[u]] = v;
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) { c0[u] = n0[u]; c1[u] = n1[u]; ccol[u] = ncol[u]; }
      crow = nrow;
    }
  } else {
    const double *pool = stage_pool(p, k6sm);
    const int col = blockIdx.x * blockDim.x + threadIdx.x;
    if (!(col >= c_lo && col < c_hi)) return;
    SEL sel(p.maps, pool, p.cols);
    unsigned flags = 0;
    bool ok = true;
    for (int row = r_lo + blockIdx.y; row < r_hi; row += gridDim.y) {
      const size_t f = (size_t)row * p.cols + col;
      const double p0 = p.predict[f], p1 = p.predict[f + nb];
      bool okf = true;
      double v = gs_take_average<true>(sel.get(GS_SEL_APPLY, col, row), p0, p1, flags, okf);
      if (!okf) v = gs_take_average<false>(sel.get(GS_SEL_APPLY, col, row), p0, p1, flags, ok);   // operands out of the fast range
      q.F[f] = v;
    }
    if (flags) atomicOr(p.status, flags);
  }
}

// ---- sweep -------------------------------------------------------------------------------------------------------
struct K6Params {
  int n, nlines, m, S, cols, rows;
  const double *coefs;     // Dim.coefs [2n]
  const double *F, *first3, *last3;
  const double *rhs;       // x: grid, y: result
  double *out;             // x: result
  double *grid_io;         // y: old grid in / new grid out
  double *predict_out;     // y: new predict
  double4 *ckpt;           // [n / CH][nlines padded]: delta, lambda, omega, fprev at the end of every chunk
  double time;
  unsigned *status;
  // split sweep of a row-sharded grid (K6_BEGIN / K6_END around the ranks' exchange)
  int open_first, open_last;
  const double *fm1;       // [nlines] face above node 0 (open first end)
  double *cdw;             // [S][3][nlines padded]: the local reduced system after elimination, E_k = c E_{k+1} + d + w L0
  double *exported;        // [6][nlines]: this rank as ONE segment: C, D, W of its last node and P, Q, R of its first
  // single-process multi-GPU (gs_sharded2d): the export goes straight into every device's gathered[] slot of this rank
  // -- peer-to-peer stores over NVLink from inside the elimination kernel, no collective afterwards
  double *export_peer[GS_MAX_SHARDS];
  int nexport_peer;
  const double *gathered;  // [nranks][6][nlines]
  int nranks, rank;
  int csize;               // CTAs per strip (thread-block cluster): the line's segments are spread over csize x S warps
  // fused faces (FF): phase A reads `predict` instead of F, evaluates the faces along its lines itself and writes F
  // for phase C -- no separate faces pass
  const double *predict;
  double *Fout;            // = F
  const double *pool;      // spline pool (global) and its length; staged into shared memory
  int pool_len;
  Maps2d maps;
};
#define K6_FUSED 0
#define K6_BEGIN 1
#define K6_END 2

// PRED: the second tile receives `predict` of nodes i0 .. i0 + len (one more than the chunk, when that node exists)
// instead of the face values F of nodes i0 .. i0 + len - 1
template <bool XDIR, int CH, bool PRED = false>
__device__ __forceinline__ void k6_load(const K6Params &q, int line0, int nl, int i0, int len, double *tileT,
                                        double *tileF, double2 *tabs, int lane) {
  constexpr int PITCH = K5Tile<XDIR>::PITCH;
  if (lane < len) k5_cp16(tabs + lane, q.coefs + 2 * (size_t)(i0 + lane));   // (cf0, cf1) of the chunk's nodes
  const double *second = PRED ? q.predict : q.F;
  if (PRED && i0 + len < q.n) {   // the node after the chunk
    if (XDIR) {
      if (lane < nl) k5_cp8(&tileF[len * PITCH + lane], second + (size_t)(line0 + lane) * q.cols + i0 + len);
    } else {
      if (lane < nl) k5_cp8(&tileF[len * PITCH + lane], second + (size_t)(i0 + len) * q.cols + line0 + lane);
    }
  }
  if (XDIR) {
    const int jj = lane & (CH - 1);
    if (jj < len) {
      const size_t off = (size_t)line0 * q.cols + i0 + jj;
      const double *gT = q.rhs + off, *gF = second + off;
      const size_t rs = (size_t)(32 / CH) * q.cols;
      int r = lane / CH;
      gT += (size_t)r * q.cols; gF += (size_t)r * q.cols;
      double *sT = &tileT[jj * PITCH + r], *sF = &tileF[jj * PITCH + r];
#pragma unroll
      for (int u = 0; u < CH; ++u, r += 32 / CH) {
        if (r < nl) {
          k5_cp8(sT + u * (32 / CH), gT);
          k5_cp8(sF + u * (32 / CH), gF);
        }
        gT += rs; gF += rs;
      }
    }
  } else {
    if (lane < nl) {
      const size_t off = (size_t)i0 * q.cols + line0 + lane;
      const double *gT = q.rhs + off, *gF = second + off;
#pragma unroll
      for (int j = 0; j < CH; ++j) {
        if (j < len) {
          k5_cp8(&tileT[j * PITCH + lane], gT);
          k5_cp8(&tileF[j * PITCH + lane], gF);
        }
        gT += q.cols; gF += q.cols;
      }
    }
  }
  k5_commit();
}

// elimination state carried along a line: (delta, lambda, omega) of the last node and its upper face value
struct K6State {
  double delta, lambda, omega, fprev;
};

// 1 / b to within an ulp: hardware seed (relative error < 2^-20) and one cubically convergent step
// y0 (1 + e + e^2), e = 1 - b y0.  K6 is the tolerance solver (<= 1e-12 per field); a division's last-bit rounding
// is not part of its contract, its speed is: the pivot's reciprocal sits on every node's critical path.
__device__ __forceinline__ double k6_rcp(double b) {
  double y0;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(b));
  const double e = fma(-b, y0, 1.0);
  return fma(y0, fma(e, e, e), y0);
}

// MODE K6_FUSED: the whole solve.  K6_BEGIN: phase A and the condensation of this rank's rows into one segment
// (exported for the all-gather).  K6_END: the ranks' reduced system (every rank solves all of it, a few unknowns per
// line), the local back-substitution and phase C.
// FF: phase A evaluates the faces itself (see K6Params); the second tile of a buffer then has CH + 1 rows.  Saves the
// faces pass and 8 B/node of traffic per direction, but MEASURED (B200, 4096^2, M1Hermite3): 0.564 ms/step against
// 0.545 ms with the separate k6_faces pass -- at 16 warps and 128 registers per SM the ~90 instructions per face run
// slower inside the sweep than in the 32-warp faces kernel, and all CTAs are in phase A at the same time, so nothing
// overlaps with the bandwidth-bound phase C.  Kept behind the option `k6_fuse` (default off) as a tested variant: the
// next step is to give the faces to dedicated producer warps.
template <bool XDIR, int CH, int MODE, bool FF>
__global__ void __launch_bounds__(512) k6_sweep(K6Params q) {
  extern __shared__ __align__(16) double k6s[];
  constexpr int PITCH = K5Tile<XDIR>::PITCH;
  constexpr int TD = CH * PITCH;
  constexpr int TD2 = FF ? (CH + 1) * PITCH : TD;     // second tile: F [CH] or predict [CH + 1]
  constexpr int BUF = TD + TD2;                        // one buffer = (t tile, second tile)
  const int S = q.S, m = q.m, n = q.n;
  const int lane = threadIdx.x & 31, k = threadIdx.x >> 5;
  // A strip of 32 lines belongs to a cluster of csize CTAs (1 when there are strips enough to fill the chip): CTA
  // `crank` of the cluster eliminates segments crank * S .. crank * S + S - 1 of every line; the reduced system of
  // all csize * S segments is solved by the cluster's first CTA through distributed shared memory.
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  const int csz = q.csize, crank = csz > 1 ? (int)cluster.block_rank() : 0, ST = S * csz;
  const int line0 = (blockIdx.x / csz) * 32, nl = min(32, q.nlines - line0);
  // shared: per CTA seg[6][S][32] (delta_e, omega_e, P, R, lambda_e, Q) ; per warp 2 x (tileT, tileF) + 2 x tabs
  double *segs = k6s;
  double *wbase = k6s + 6 * S * 32 + (size_t)k * (2 * BUF + 4 * CH);
  double2 *tabs2 = reinterpret_cast<double2 *>(wbase);
  double *tiles = wbase + 4 * CH;
  double *poolsm = k6s + 6 * S * 32 + (size_t)S * (2 * BUF + 4 * CH);   // FF: the spline pool
  if (FF && MODE != K6_END) {
    for (int i = threadIdx.x; i < q.pool_len; i += blockDim.x) poolsm[i] = q.pool[i];
    __syncthreads();
  }
  const int s = (crank * S + k) * m, e = min(n, s + m) - 1;
  auto sync_strip = [&]() { if (csz > 1) cluster.sync(); else __syncthreads(); };
  // segment kk's slot column in whichever CTA of the cluster holds it
  auto seg_at = [&](int kk) -> double * {
    double *base = segs;
    if (csz > 1) base = cluster.map_shared_rank(segs, kk / S);
    return base + (size_t)(kk % S) * 32 + lane;
  };
  const int nch = (e >= s) ? (e - s) / CH + 1 : 0;
  const bool act = lane < nl;
  const int line = line0 + (act ? lane : 0);
  const size_t ck_pitch = (size_t)((q.nlines + 31) / 32) * 32;
  double4 *ckpt = q.ckpt + (size_t)(s / CH) * ck_pitch + line0 + lane;
  const double inv_time = 1.0 / q.time;
  const double f_diag = q.first3[3 * (size_t)line], f_sup = q.first3[3 * (size_t)line + 1], f_rhs = q.first3[3 * (size_t)line + 2];
  const double l_sub = q.last3[3 * (size_t)line], l_diag = q.last3[3 * (size_t)line + 1], l_rhs = q.last3[3 * (size_t)line + 2];
  const size_t lstride = XDIR ? 1 : (size_t)q.cols, lbase = XDIR ? (size_t)line * q.cols : (size_t)line;

  // forward elimination of an interior node (0 < i < n - 1); st carries (delta, lambda, omega, fprev)
  auto node_mid = [&](double t, double F, double2 cf, K6State &st) {
    const double sub = cf.x * st.fprev, sup = cf.y * F;
    const double diag = -(sub + sup) - inv_time, rhs = -t * inv_time;
    const double inv = k6_rcp(diag + sub * st.delta);
    const double beta = -sub * inv;
    st.delta = -sup * inv;
    st.lambda = fma(rhs, inv, st.lambda * beta);
    st.omega = beta * st.omega;
    st.fprev = F;
  };
  // any node: the two end rows come assembled from k6_ends
  const bool open_first = MODE != K6_FUSED && q.open_first, open_last = MODE != K6_FUSED && q.open_last;
  auto node_any = [&](int i, double t, double F, double2 cf, K6State &st) {
    if ((i > 0 || open_first) && (i < n - 1 || open_last)) { node_mid(t, F, cf, st); return; }
    double sub, diag, sup, rhs;
    if (i == 0) { sub = 0.0; diag = f_diag; sup = f_sup; rhs = f_rhs; }
    else { sub = l_sub; diag = l_diag; sup = 0.0; rhs = l_rhs; }
    const double inv = k6_rcp(diag + sub * st.delta);
    const double beta = -sub * inv;
    st.delta = -sup * inv;
    st.lambda = fma(rhs, inv, st.lambda * beta);
    st.omega = beta * st.omega;
    st.fprev = F;
  };
  // a chunk is `plain` when it is full and holds neither node 0 nor the segment's last node e (hence not n - 1):
  // its loops are unrolled without per-node tests
  auto plain = [&](int i0) { return i0 > 0 && i0 + CH - 1 < e; };

  const size_t cd_pitch = ck_pitch;
  // FF: one face along the line, between a node and its successor, from their predicted temperatures (the tiers of
  // k6_faces; `i` = position of the first of the two on the line, for coefficient maps that vary over the grid)
  const bool single_map = !FF || q.maps.mode[GS_SEL_APPLY] == GS_MAP_SINGLE;
  int zoff = FF ? q.maps.single_off[GS_SEL_APPLY] : 0;
  K6Tab ztab = {};
  if (FF && MODE != K6_END) ztab = k6_tab(poolsm + zoff);
  auto face = [&](double p0, double p1, int i) -> double {
    if (!single_map) {
      zoff = map_off(q.maps, GS_SEL_APPLY, XDIR ? i : line, XDIR ? line : i, q.cols);
      ztab = k6_tab(poolsm + zoff);
    }
    double v;
    if (!k6_face_fast(ztab, p0, p1, v) && !k6_face_equal(ztab, p0, p1, v) && !k6_face_clamped(poolsm + zoff, p0, p1, v)) {
      if (!k6_face_mid(poolsm + zoff, p0, p1, &v)) v = k6_face_general(q.pool, zoff, p0, p1, q.status);
    }
    return v;
  };
  // which faces the sweep evaluates itself: all but the ones k6_ends owns (the carry out of a closed first node, the
  // last node's: 0, or the face to the rank below)
  auto own_face = [&](int i) { return (i >= 1 || open_first) && i <= n - 2; };
  // face below the segment's left neighbour (the rank above holds that neighbour when the first end is open)
  double f_seg;
  if (s > 0 && s < n) {
    if (FF && MODE != K6_END && own_face(s - 1)) {
      f_seg = act ? face(q.predict[lbase + (size_t)(s - 1) * lstride], q.predict[lbase + (size_t)s * lstride], s - 1) : 0.0;
    } else {
      f_seg = q.F[lbase + (size_t)(s - 1) * lstride];
    }
  } else {
    f_seg = (s == 0 && open_first) ? q.fm1[line] : 0.0;
  }
  if (MODE != K6_END) {
    // ---- phase A ----
    K6State st;
    st.delta = 0.0; st.lambda = 0.0; st.omega = (s > 0 || open_first) ? 1.0 : 0.0;
    st.fprev = f_seg;
    double P = 1.0, Qa = 0.0, R = 0.0;
    if (nch > 0) k6_load<XDIR, CH, FF>(q, line0, nl, s, min(CH, e - s + 1), tiles, tiles + TD, tabs2, lane);
#pragma unroll 1
    for (int c = 0; c < nch; ++c) {
      const int i0 = s + c * CH, len = min(CH, e - i0 + 1);
      double *tT = tiles + (c & 1) * BUF, *tF = tT + TD;
      const double2 *cf = tabs2 + (c & 1) * CH;
      if (c + 1 < nch) {
        const int i1 = i0 + CH;
        double *nT = tiles + ((c + 1) & 1) * BUF;
        k6_load<XDIR, CH, FF>(q, line0, nl, i1, min(CH, e - i1 + 1), nT, nT + TD, tabs2 + ((c + 1) & 1) * CH, lane);
        k5_wait<1>();
      } else {
        k5_wait<0>();
      }
      __syncwarp();
      const bool pl = plain(i0);
      if (FF) {
        // the chunk's faces, from the predict tile (rows 0 .. len), into registers and out to F for phase C
        double Fv[CH];
        if (pl) {
#pragma unroll
          for (int j = 0; j < CH; ++j)
            Fv[j] = act ? face(tF[j * PITCH + lane], tF[(j + 1) * PITCH + lane], i0 + j) : 0.0;
        } else {
#pragma unroll
          for (int j = 0; j < CH; ++j) {
            Fv[j] = 0.0;
            if (j < len && act) {
              const int i = i0 + j;
              Fv[j] = own_face(i) ? face(tF[j * PITCH + lane], tF[(j + 1) * PITCH + lane], i)
                                  : __ldcg(q.F + lbase + (size_t)i * lstride);
            }
          }
        }
        __syncwarp();
        if (XDIR) {
          // through the tile for a transposed store (rows of F are lines here)
#pragma unroll
          for (int j = 0; j < CH; ++j) tF[j * PITCH + lane] = Fv[j];
          __syncwarp();
          const int jj = lane & (CH - 1);
          if (jj < len) {
            int r = lane / CH;
            double *gq = q.Fout + (size_t)(line0 + r) * q.cols + i0 + jj;
            const size_t rs = (size_t)(32 / CH) * q.cols;
#pragma unroll
            for (int u = 0; u < CH; ++u, r += 32 / CH, gq += rs)
              if (r < nl) *gq = tF[jj * PITCH + r];
          }
        } else if (act) {
          double *gq = q.Fout + (size_t)i0 * q.cols + line0 + lane;
#pragma unroll
          for (int j = 0; j < CH; ++j)
            if (j < len) gq[(size_t)j * q.cols] = Fv[j];
        }
        if (pl) {
#pragma unroll
          for (int j = 0; j < CH; ++j) {
            node_mid(tT[j * PITCH + lane], Fv[j], cf[j], st);
            Qa = fma(P, st.lambda, Qa);
            R = fma(P, st.omega, R);
            P = P * st.delta;
          }
        } else {
#pragma unroll
          for (int j = 0; j < CH; ++j) {
            if (j < len) {
              node_any(i0 + j, tT[j * PITCH + lane], Fv[j], cf[j], st);
              if (i0 + j < e) {
                Qa = fma(P, st.lambda, Qa);
                R = fma(P, st.omega, R);
                P = P * st.delta;
              }
            }
          }
        }
      } else if (pl) {
#pragma unroll
        for (int j = 0; j < CH; ++j) {
          node_mid(tT[j * PITCH + lane], tF[j * PITCH + lane], cf[j], st);
          Qa = fma(P, st.lambda, Qa);      // x_s = P x_e + Q + R L gathers nodes s .. e-1
          R = fma(P, st.omega, R);
          P = P * st.delta;
        }
      } else {
        for (int j = 0; j < len; ++j) {
          node_any(i0 + j, tT[j * PITCH + lane], tF[j * PITCH + lane], cf[j], st);
          if (i0 + j < e) {
            Qa = fma(P, st.lambda, Qa);
            R = fma(P, st.omega, R);
            P = P * st.delta;
          }
        }
      }
      ckpt[(size_t)c * ck_pitch] = make_double4(st.delta, st.lambda, st.omega, st.fprev);
      __syncwarp();
    }
    {
      double *sg = segs + (size_t)k * 32 + lane;
      sg[0 * S * 32] = st.delta;    // delta_e (0 on a closed line's last node)
      sg[1 * S * 32] = st.omega;
      sg[2 * S * 32] = P;
      sg[3 * S * 32] = R;
      sg[4 * S * 32] = st.lambda;
      sg[5 * S * 32] = Qa;
    }
  }
  sync_strip();
  // ---- reduced system: first CTA's warp 0, lane = line.  Unknowns E_k = x at the end of segment k:
  //   (1 - de_k R_{k+1}) E_k - om_k E_{k-1} - de_k P_{k+1} E_{k+1} = la_k + de_k Q_{k+1}
  // eliminated forward into  E_k = c_k E_{k+1} + d_k + w_k L0  (L0 = x above node 0, X = x below node n-1; both absent
  // on a closed line, where om_0 = 0 and de_last = 0 make w and the last c vanish by themselves).
  // Slots afterwards: 4 = E_k, 2 = c_k, 3 = d_k, 1 = w_k; 5*S*32 + lane of every CTA = x above its first segment.
  if (k == 0 && crank == 0) {
    double L0 = 0.0, Elast = 0.0;
    if (MODE != K6_END) {
      double c = 0.0, d = 0.0, w = 1.0;
      double Pc = 1.0, Qc = 0.0, Rc = 0.0;                  // E_0 = Pc E_k + Qc + Rc L0
      const double P0 = segs[2 * S * 32 + lane], Q0 = segs[5 * S * 32 + lane], R0 = segs[3 * S * 32 + lane];
      double *sg = seg_at(0);
      for (int kk = 0; kk < ST; ++kk) {
        double *sn = (kk + 1 < ST) ? seg_at(kk + 1) : sg;
        const double de = sg[0], om = sg[1 * S * 32], la = sg[4 * S * 32];
        double Pn = 0.0, Rn = 0.0, Qn = 0.0;
        if (kk + 1 < ST) { Pn = sn[2 * S * 32]; Rn = sn[3 * S * 32]; Qn = sn[5 * S * 32]; }
        const double a = -om, b = (kk + 1 < ST) ? 1.0 - de * Rn : 1.0, gg = (kk + 1 < ST) ? -(de * Pn) : -de;
        const double r = la + de * Qn;
        const double inv = 1.0 / (b + a * c);
        if (kk > 0) { Qc = fma(Pc, d, Qc); Rc = fma(Pc, w, Rc); Pc = Pc * c; }   // folds E_{kk-1} = c E_kk + d + w L0
        c = -gg * inv;
        d = (r - a * d) * inv;
        w = -(a * w) * inv;
        sg[2 * S * 32] = c;        // P, R of segment kk were consumed at kk-1
        sg[3 * S * 32] = d;
        sg[1 * S * 32] = w;
        sg = sn;
      }
      if (MODE == K6_BEGIN) {
        if (act) {
          const size_t nlp = (size_t)q.nlines;
          const double e3 = P0 * Pc, e4 = fma(P0, Qc, Q0), e5 = fma(P0, Rc, R0);
          const int ndst = q.nexport_peer > 0 ? q.nexport_peer : 1;
          for (int t = 0; t < ndst; ++t) {
            double *ex = (q.nexport_peer > 0 ? q.export_peer[t] : q.exported) + line;
            ex[0 * nlp] = c; ex[1 * nlp] = d; ex[2 * nlp] = w;           // E_last = c X + d + w L0
            ex[3 * nlp] = e3;                                           // x_first = P E_last + Q + R L0
            ex[4 * nlp] = e4;
            ex[5 * nlp] = e5;
          }
        }
        for (int kk = 0; kk < ST; ++kk) {
          const double *sk = seg_at(kk);
          double *o = q.cdw + (size_t)kk * 3 * cd_pitch + line0 + lane;
          o[0] = sk[2 * S * 32]; o[cd_pitch] = sk[3 * S * 32]; o[2 * cd_pitch] = sk[1 * S * 32];
        }
      } else {
        Elast = d;   // closed line: X = L0 = 0
      }
    } else {
      // the ranks' reduced system (same recurrence, one segment per rank), then this rank's two interface values
      const size_t nlp = (size_t)q.nlines;
      double cs[16], ds[16];
      double c = 0.0, d = 0.0;
      for (int r = 0; r < q.nranks; ++r) {
        const double *gr = q.gathered + (size_t)r * 6 * nlp + line;
        const double de = gr[0], la = gr[nlp], om = gr[2 * nlp];
        double Pn = 0.0, Qn = 0.0, Rn = 0.0;
        if (r + 1 < q.nranks) { Pn = gr[9 * nlp]; Qn = gr[10 * nlp]; Rn = gr[11 * nlp]; }
        const double a = -om, b = 1.0 - de * Rn, gg = -(de * Pn), rr = la + de * Qn;
        const double inv = 1.0 / (b + a * c);
        c = -gg * inv;
        d = (rr - a * d) * inv;
        cs[r] = c; ds[r] = d;
      }
      double x = 0.0;
      for (int r = q.nranks - 1; r >= 0; --r) {
        x = fma(cs[r], x, ds[r]);
        if (r == q.rank) Elast = x;
        if (r == q.rank - 1) L0 = x;
      }
      for (int kk = 0; kk < ST; ++kk) {
        const double *o = q.cdw + (size_t)kk * 3 * cd_pitch + line0 + lane;
        double *sk = seg_at(kk);
        sk[2 * S * 32] = o[0]; sk[3 * S * 32] = o[cd_pitch]; sk[1 * S * 32] = o[2 * cd_pitch];
      }
    }
    if (MODE != K6_BEGIN) {
      double x = Elast;
      for (int kk = ST - 1; kk >= 0; --kk) {
        double *sk = seg_at(kk);
        if (kk < ST - 1) x = fma(sk[2 * S * 32], x, fma(sk[1 * S * 32], L0, sk[3 * S * 32]));
        sk[4 * S * 32] = x;
        // the value above the first segment of the CTA that follows
        if (kk % S == S - 1 && kk + 1 < ST) seg_at(kk + 1)[5 * S * 32] = x;
      }
      segs[5 * S * 32 + lane] = L0;
    }
  }
  sync_strip();
  if (MODE == K6_BEGIN) return;
  // ---- phase C ----
  if (nch == 0) return;
  const double L = (k > 0) ? segs[4 * S * 32 + (size_t)(k - 1) * 32 + lane] : segs[5 * S * 32 + lane];
  double x = segs[4 * S * 32 + (size_t)k * 32 + lane];
  {
    const int il = s + (nch - 1) * CH;
    double *nT = tiles + ((nch - 1) & 1) * BUF;
    k6_load<XDIR, CH>(q, line0, nl, il, min(CH, e - il + 1), nT, nT + TD, tabs2 + ((nch - 1) & 1) * CH, lane);
  }
  double4 ck_next = (nch > 1) ? ckpt[(size_t)(nch - 2) * ck_pitch] : make_double4(0.0, 0.0, 0.0, 0.0);
#pragma unroll 1
  for (int c = nch - 1; c >= 0; --c) {
    const int i0 = s + c * CH, len = min(CH, e - i0 + 1);
    double *tT = tiles + (c & 1) * BUF, *tF = tT + TD;
    const double2 *cf = tabs2 + (c & 1) * CH;
    // old grid values of the chunk (fused Grid.update): issued ahead of the arithmetic
    double gold[CH];
    const size_t yoff = (size_t)i0 * q.cols + line0 + lane;
    if (!XDIR) {
#pragma unroll
      for (int u = 0; u < CH; ++u) gold[u] = (act && u < len) ? q.grid_io[yoff + (size_t)u * q.cols] : 0.0;
    }
    if (c > 0) {
      double *nT = tiles + ((c - 1) & 1) * BUF;
      k6_load<XDIR, CH>(q, line0, nl, i0 - CH, CH, nT, nT + TD, tabs2 + ((c - 1) & 1) * CH, lane);
      k5_wait<1>();
    } else {
      k5_wait<0>();
    }
    __syncwarp();
    // state at the start of the chunk: phase A's checkpoint, with the left-neighbour value folded into lambda
    K6State sc;
    if (c > 0) {
      sc.delta = ck_next.x; sc.lambda = ck_next.y + ck_next.z * L; sc.omega = 0.0; sc.fprev = ck_next.w;
    } else {
      sc.delta = 0.0; sc.lambda = L; sc.omega = 0.0; sc.fprev = f_seg;   // lambda_s = rhs/Delta + beta_s L
    }
    if (c > 1) ck_next = ckpt[(size_t)(c - 2) * ck_pitch];
    double xs[CH];
    if (plain(i0)) {
      double dl[CH], la[CH];
#pragma unroll
      for (int j = 0; j < CH; ++j) {
        node_mid(tT[j * PITCH + lane], tF[j * PITCH + lane], cf[j], sc);
        dl[j] = sc.delta; la[j] = sc.lambda;
      }
#pragma unroll
      for (int j = CH - 1; j >= 0; --j) {
        x = fma(dl[j], x, la[j]);
        xs[j] = x;
      }
    } else {
      for (int j = 0; j < len; ++j) {
        node_any(i0 + j, tT[j * PITCH + lane], tF[j * PITCH + lane], cf[j], sc);
        tF[j * PITCH + lane] = sc.delta;
        tT[j * PITCH + lane] = sc.lambda;
      }
#pragma unroll
      for (int j = CH - 1; j >= 0; --j) {
        xs[j] = 0.0;
        if (j < len) {
          if (i0 + j != e) x = fma(tF[j * PITCH + lane], x, tT[j * PITCH + lane]);
          xs[j] = x;
        }
      }
    }
    if (XDIR) {
#pragma unroll
      for (int j = 0; j < CH; ++j) tT[j * PITCH + lane] = xs[j];
      __syncwarp();
      const int jj = lane & (CH - 1);
      if (jj < len) {
        int r = lane / CH;
        double *g = q.out + (size_t)(line0 + r) * q.cols + i0 + jj;
        const size_t rs = (size_t)(32 / CH) * q.cols;
#pragma unroll
        for (int u = 0; u < CH; ++u, r += 32 / CH, g += rs)
          if (r < nl) *g = tT[jj * PITCH + r];
      }
    } else if (act) {
      double *G = q.grid_io + yoff, *Pp = q.predict_out + yoff;
#pragma unroll
      for (int u = 0; u < CH; ++u) {
        if (u < len) {
          const double xv = xs[u], d = xv - gold[u];                      // Grid.update :474-481, aVal :508-512
          Pp[(size_t)u * q.cols] = (fabs(d) > 1.5) ? xv : xv + d / 2.0;
          G[(size_t)u * q.cols] = xv;
        }
      }
    }
    __syncwarp();
  }
  if (act && !(fabs(x) <= 1.7976931348623157e308)) atomicOr(q.status, GS_FLAG_NONFINITE);
}
