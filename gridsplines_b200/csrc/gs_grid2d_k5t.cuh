// gs_grid2d_k5t.cuh -- K5T: K5's sweeps (gs_grid2d_k5.cuh: literal-constant coefficients, tolerance family) fed by
// tensor-map TMA rings instead of 8-byte cp.async copies.  Same arithmetic, same tables (k5_setup), same results as
// k5_sweep bit for bit; what changes is who moves the bytes:
//   * a chunk of a warp's 32 lines x 16 nodes arrives as ONE box (cp.async.bulk.tensor.2d, UTMALDG) on an mbarrier,
//     together with the chunk's 16 table entries (one 512-byte bulk copy);
//       x sweep: lane <-> row, box = 16 columns x `strip` rows landing with the 128-byte swizzle, so that a lane reads
//                its own row 16 bytes at a time without bank conflicts (the TMA/shared-memory tiled transpose);
//       y sweep / 1-D batches: lane <-> column (line), box = 32 columns x 16 rows;
//   * a lane copies its 16 values into registers as soon as the box has landed and lane 0 refills the stage at once:
//     all NT stages of a warp are in flight while it computes (k5_sweep: one);
//   * results leave through per-warp output tiles as box stores (UTMASTG): x sweep `result`; y sweep the new grid and
//     the new predict (Grid.update :474-481 applied on the way out); the old grid values of Grid.update arrive as a
//     second box of the stage (2-D) or are the right-hand side itself (1-D batches).
// Shared memory per warp: NT tiles of 4 KB, NT + 1 table slots of 512 B (the slot of the chunk being processed must
// survive the refill), 1 (x) or 2 (y) output tiles, NT mbarriers.  One CTA = S warps = S segments of 32 lines.
// Restrictions (the host falls back to k5_sweep otherwise): an even number of columns and 16-byte aligned arrays
// (tensor-map strides), 1-D batches with n % 16 == 0 (a box must not straddle two 32-line blocks), no clusters.
#pragma once
#include "gs_tma.cuh"

#define K5T_TILE 4096
#define K5T_TAB 512

struct K5TMaps {
  K7Map rhs;    // x: grid, y: result, 1-D: grid            (loads)
  K7Map out;    // x: result; y / 1-D: grid                 (stores)
  K7Map gold;   // y: grid                                  (loads: old values of Grid.update)
  K7Map pout;   // y / 1-D: predict                         (stores)
};

static size_t k5t_smem(bool xdir, int S, int NT) {
  const size_t per_warp = (size_t)NT * K5T_TILE + (size_t)(NT + 1) * K5T_TAB + (size_t)(xdir ? 1 : 2) * K5T_TILE + (size_t)NT * 8;
  return 1024 + (size_t)S * per_warp + (size_t)2 * S * 32 * sizeof(double);
}

template <bool XDIR, bool ONED, int MAXT>
__global__ void __launch_bounds__(MAXT) k5t_sweep(K5Params q, int NT, const __grid_constant__ K5TMaps tm) {
  extern __shared__ unsigned char k5traw[];
  // the swizzled boxes need a 1024-byte aligned base
  unsigned char *base = k5traw + ((1024u - (gs_smem_u32(k5traw) & 1023u)) & 1023u);
  constexpr int NOUT = XDIR ? 1 : 2;
  constexpr bool GOLD = !XDIR && !ONED;   // the old grid values come as a second box
  const K5Dir &d = q.d;
  const int S = d.S, m = d.m, n = d.n;
  const int lane = threadIdx.x & 31, k = threadIdx.x >> 5;   // warp k = segment k
  const int strip_id = blockIdx.x;
  const int line0 = strip_id * q.strip, nl = min(q.strip, d.nlines - line0);
  // shared memory: tiles [S][NT] | output tiles [S][NOUT] | table slots [S][NT + 1] | lamE [S][32] | Q [S][32] | barriers [S][NT]
  unsigned char *tiles = base + (size_t)k * NT * K5T_TILE;
  unsigned char *otile = base + (size_t)S * NT * K5T_TILE + (size_t)k * NOUT * K5T_TILE;
  unsigned char *tabs = base + (size_t)S * (NT + NOUT) * K5T_TILE + (size_t)k * (NT + 1) * K5T_TAB;
  double *lamE = reinterpret_cast<double *>(base + (size_t)S * (NT + NOUT) * K5T_TILE + (size_t)S * (NT + 1) * K5T_TAB);
  double *Qs = lamE + S * 32, *Es = lamE;
  uint64_t *bars = reinterpret_cast<uint64_t *>(Qs + S * 32) + k * NT;
  if (lane == 0) {
    for (int s = 0; s < NT; ++s) gs_mbar_init(&bars[s], 1);
    gs_mbar_init_fence();
    gs_tma_prefetch_desc(&tm.rhs);
    gs_tma_prefetch_desc(&tm.out);
    if (GOLD) gs_tma_prefetch_desc(&tm.gold);
    if (!XDIR) gs_tma_prefetch_desc(&tm.pout);
  }
  gs_fence_proxy_async();
  __syncwarp();

  const size_t ck_pitch = (size_t)((d.nlines + 31) / 32) * 32;
  const int s = k * m, e = min(n, s + m) - 1;
  double *ckpt = q.ckpt + (size_t)(s / K5_CH) * ck_pitch + line0 + lane;
  const int nch = (e >= s) ? (e - s) / K5_CH + 1 : 0;
  const bool act = lane < nl;
  const int line = line0 + (act ? lane : 0);
  const double ex0 = d.e0a * d.bound_first[(size_t)line * q.bc_stride] + d.e0b;
  const double ex1 = d.e1a * d.bound_last[(size_t)line * q.bc_stride] + d.e1b;
  const uint32_t tbytes = XDIR ? (uint32_t)q.strip * 128u : (uint32_t)K5T_TILE;   // bytes of one box (outside parts count too)
  uint32_t phase = 0;

  // box coordinates (innermost first) of the chunk that starts at node i0
  auto c0_of = [&](int i0) { return XDIR ? i0 : (ONED ? 0 : line0); };
  auto c1_of = [&](int i0) { return XDIR ? line0 : (ONED ? strip_id * n + i0 : i0); };
  // lane 0: chunk at node i0 -> tile stage `st` (+ the stage after it for the old grid values) and table slot `ts`
  auto issue = [&](int st, int ts, int i0, bool gold) {
    const int len = min(K5_CH, e - i0 + 1);
    uint64_t *bar = &bars[st];
    gs_mbar_expect_tx(bar, tbytes * (gold ? 2u : 1u) + (uint32_t)len * 32u);
    gs_tma_load_2d(tiles + (size_t)st * K5T_TILE, &tm.rhs, c0_of(i0), c1_of(i0), bar);
    gs_bulk_g2s(tabs + (size_t)ts * K5T_TAB, d.tab + i0, (uint32_t)len * 32u, bar);
    if (gold) gs_tma_load_2d(tiles + (size_t)(st + 1) * K5T_TILE, &tm.gold, c0_of(i0), c1_of(i0), bar);
  };
  auto wait = [&](int st) {
    gs_mbar_wait_bounded(&bars[st], (phase >> st) & 1u, q.status);
    phase ^= 1u << st;
  };
  // a lane's 16 values of a tile <-> registers
  auto load16 = [&](const unsigned char *tile, double (&v)[K5_CH]) {
    if (XDIR) {
      const unsigned char *row = tile + lane * 128;
      const int sw = lane & 7;
#pragma unroll
      for (int u = 0; u < K5_CH / 2; ++u) {
        const double2 w = *reinterpret_cast<const double2 *>(row + ((u ^ sw) << 4));
        v[2 * u] = w.x;
        v[2 * u + 1] = w.y;
      }
    } else {
      const double *col = reinterpret_cast<const double *>(tile) + lane;
#pragma unroll
      for (int j = 0; j < K5_CH; ++j) v[j] = col[j * 32];
    }
  };
  auto store16 = [&](unsigned char *tile, const double (&v)[K5_CH]) {
    if (XDIR) {
      unsigned char *row = tile + lane * 128;
      const int sw = lane & 7;
#pragma unroll
      for (int u = 0; u < K5_CH / 2; ++u)
        *reinterpret_cast<double2 *>(row + ((u ^ sw) << 4)) = make_double2(v[2 * u], v[2 * u + 1]);
    } else {
      double *col = reinterpret_cast<double *>(tile) + lane;
#pragma unroll
      for (int j = 0; j < K5_CH; ++j) col[j * 32] = v[j];
    }
  };
  // a chunk is `plain` when it is full and holds neither node 0 nor the segment's last node e (hence not n - 1)
  auto plain = [&](int i0) { return i0 > 0 && i0 + K5_CH - 1 < e; };
  const int NTS = NT + 1;   // table slots

  // ---- phase A: lambda_e and Q of the segment, a lambda checkpoint per chunk ----
  double lam = 0.0, Qacc = 0.0;
  if (lane == 0)
    for (int c = 0; c < min(NT, nch); ++c) issue(c, c, s + c * K5_CH, false);
  {
    int st = 0, ts = 0;
    for (int c = 0; c < nch; ++c) {
      const int i0 = s + c * K5_CH, len = min(K5_CH, e - i0 + 1);
      wait(st);
      double tv[K5_CH];
      load16(tiles + (size_t)st * K5T_TILE, tv);
      __syncwarp();
      if (lane == 0 && c + NT < nch) issue(st, (ts + NT) % NTS, i0 + NT * K5_CH, false);
      const double4 *t = reinterpret_cast<const double4 *>(tabs + (size_t)ts * K5T_TAB);
      if (plain(i0)) {
#pragma unroll
        for (int j = 0; j < K5_CH; ++j) {
          const double4 c4 = t[j];
          lam = fma(tv[j], c4.x, lam * c4.y);
          Qacc = fma(c4.w, lam, Qacc);
        }
      } else {
#pragma unroll
        for (int j = 0; j < K5_CH; ++j) {
          if (j < len) {
            const double4 c4 = t[j];
            lam = fma(tv[j], c4.x, lam * c4.y);
            if (i0 + j == 0) lam += ex0;
            if (i0 + j == n - 1) lam += ex1;
            if (i0 + j < e) Qacc = fma(c4.w, lam, Qacc);
          }
        }
      }
      if (act) ckpt[(size_t)c * ck_pitch] = lam;
      st = (st + 1 == NT) ? 0 : st + 1;
      ts = (ts + 1 == NTS) ? 0 : ts + 1;
    }
  }
  lamE[k * 32 + lane] = lam;
  Qs[k * 32 + lane] = Qacc;
  __syncthreads();
  // ---- reduced system (warp 0; lane = line).  Afterwards Es[k] = x at the end of segment k ----
  if (k == 0) {
    double dp = 0.0;
    for (int kk = 0; kk < S; ++kk) {
      const double4 sg = d.seg[kk];
      const double rhs = lamE[kk * 32 + lane] + ((kk + 1 < S) ? sg.x * Qs[(kk + 1) * 32 + lane] : 0.0);
      dp = (rhs - sg.y * dp) * sg.z;
      lamE[kk * 32 + lane] = dp;
    }
    double x = 0.0;
    for (int kk = S - 1; kk >= 0; --kk) {
      x = lamE[kk * 32 + lane] - d.seg[kk].w * x;
      Es[kk * 32 + lane] = x;
    }
  }
  __syncthreads();
  // ---- phase C: chunks last to first: recompute the lambdas from the checkpoint, back-substitute, store ----
  if (nch == 0) return;
  const double L = (k > 0) ? Es[(k - 1) * 32 + lane] : 0.0;   // x just before the segment
  double x = Es[k * 32 + lane];                                // x_e
  constexpr int SS = GOLD ? 2 : 1;   // tiles per stage
  const int D = NT / SS;             // stages
  if (lane == 0)
    for (int kk = 0; kk < min(D, nch); ++kk) issue(kk * SS, kk % NTS, s + (nch - 1 - kk) * K5_CH, GOLD);
  double ck_next = (nch > 1) ? ckpt[(size_t)(nch - 2) * ck_pitch] : 0.0;               // checkpoint before the last chunk
  double om_next = (nch > 1) ? d.omega_end[(s + (nch - 1) * K5_CH - 1) / K5_CH] : 0.0;  // omega at its end
  {
    int st = 0, ts = 0;
    for (int c = nch - 1; c >= 0; --c) {
      const int i0 = s + c * K5_CH, len = min(K5_CH, e - i0 + 1);
      wait(st * SS);
      double tv[K5_CH], gv[K5_CH];
      load16(tiles + (size_t)(st * SS) * K5T_TILE, tv);
      if (GOLD) load16(tiles + (size_t)(st * SS + 1) * K5T_TILE, gv);
      // the stores issued from the output tiles one chunk ago have read them
      if (lane == 0) gs_bulk_wait_read<0>();
      __syncwarp();
      if (lane == 0 && c - D >= 0) issue(st * SS, (ts + D) % NTS, i0 - D * K5_CH, GOLD);
      const double4 *t = reinterpret_cast<const double4 *>(tabs + (size_t)ts * K5T_TAB);
      // lambda at the start of the chunk: phase A's checkpoint plus the left-neighbour term omega * L
      double lm = (c > 0) ? ck_next + om_next * L : L;
      if (c > 1) {
        ck_next = ckpt[(size_t)(c - 2) * ck_pitch];
        om_next = d.omega_end[(i0 - K5_CH - 1) / K5_CH];
      }
      double pv[K5_CH];
      if (plain(i0)) {
#pragma unroll
        for (int j = 0; j < K5_CH; ++j) {
          const double4 c4 = t[j];
          lm = fma(tv[j], c4.x, lm * c4.y);
          if (ONED) gv[j] = tv[j];
          tv[j] = lm;
        }
#pragma unroll
        for (int j = K5_CH - 1; j >= 0; --j) {
          x = fma(t[j].z, x, tv[j]);
          tv[j] = x;
          if (!XDIR) {
            const double dl = x - gv[j];
            if (ONED) pv[j] = x + dl / 3.0;                              // arraygrid.aVal :85-87
            else pv[j] = (fabs(dl) > 1.5) ? x : x + dl / 2.0;            // Grid.update :474-481, aVal :508-512
          }
        }
      } else {
#pragma unroll
        for (int j = 0; j < K5_CH; ++j) {
          if (ONED) gv[j] = tv[j];
          if (j < len) {
            const double4 c4 = t[j];
            lm = fma(tv[j], c4.x, lm * c4.y);
            if (i0 + j == 0) lm += ex0;
            if (i0 + j == n - 1) lm += ex1;
            tv[j] = lm;
          }
        }
#pragma unroll
        for (int j = K5_CH - 1; j >= 0; --j) {
          if (j < len) {
            if (i0 + j != e) x = fma(t[j].z, x, tv[j]);
            tv[j] = x;
            if (!XDIR) {
              const double dl = x - gv[j];
              if (ONED) pv[j] = x + dl / 3.0;
              else pv[j] = (fabs(dl) > 1.5) ? x : x + dl / 2.0;
            }
          } else {
            tv[j] = 0.0;      // outside the array: clipped by the store
            if (!XDIR) pv[j] = 0.0;
          }
        }
      }
      store16(otile, tv);
      if (!XDIR) store16(otile + K5T_TILE, pv);
      gs_fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        gs_tma_store_2d(&tm.out, c0_of(i0), c1_of(i0), otile);
        if (!XDIR) gs_tma_store_2d(&tm.pout, c0_of(i0), c1_of(i0), otile + K5T_TILE);
        gs_bulk_commit();
      }
      st = (st + 1 == D) ? 0 : st + 1;
      ts = (ts + 1 == NTS) ? 0 : ts + 1;
    }
  }
  if (lane == 0) gs_bulk_wait_read<0>();
  if (act && !(fabs(x) <= 1.7976931348623157e308)) atomicOr(q.status, GS_FLAG_NONFINITE);
}
