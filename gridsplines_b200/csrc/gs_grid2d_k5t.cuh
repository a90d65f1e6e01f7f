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
//   * WIDE x sweep (lines of >= 8192 nodes): 32 columns per stage arrive as ONE un-swizzled box of 34 columns (272-byte
//     rows: the two extra columns make the row pitch 17 x 16 bytes, so that a lane reads its row 16 bytes at a time
//     without bank conflicts) and are handled as two chunks of 16; results still leave as swizzled 16-column boxes.
//     Wider row pieces per request measured +3 % at 16384^2 and nothing at 4096^2, where the 70 us kernel is made of
//     its ramps.
// Shared memory per warp: NT stage tiles (4 KB; wide x: 8.5 KB), NT + 1 table slots (the slot of the chunk being
// processed must survive the refill), 1 (x) or 2 (y) output tiles, NT mbarriers.  One CTA = S warps = S segments.
// Restrictions (the host falls back to k5_sweep otherwise): an even number of columns and 16-byte aligned arrays
// (tensor-map strides), 1-D batches with n % 16 == 0 (a box must not straddle two 32-line blocks).
// With few strips (a shard of a batch, a small grid) a strip is split over csize CTAs -- CTA `crank` owns segments
// crank * S .. crank * S + S - 1 -- either a thread-block cluster (any size up to 8 that is fully resident; the first
// CTA's warp 0 solves the reduced system of all segments through distributed shared memory, as in k5_sweep) or a SOFT
// group (q.soft: plain CTAs, all resident, that meet through global memory: segment results published with a release,
// arrival counter, the first CTA stages and solves in shared memory and releases a `done` word; bounded spins), taken
// when it puts more CTAs to work than the GPCs allow clusters to.  Kernels are launched with programmatic stream
// serialisation (q.pdl): everything before griddepcontrol.wait overlaps the previous kernel's tail.
#pragma once
#include <cooperative_groups.h>

#include "gs_tma.cuh"

#define K5T_TILE 4096
#define K5T_TAB 512
#define K5T_WCOLS 34                      // wide x stage: 32 columns + 2 of padding
#define K5T_WPITCH (K5T_WCOLS * 8)        // 272-byte rows
#define K5T_WTILE (K5T_WPITCH * 32)       // 8704 bytes

struct K5TMaps {
  K7Map rhs;    // x: grid, y: result, 1-D: grid            (loads)
  K7Map out;    // x: result; y / 1-D: grid                 (stores)
  K7Map gold;   // y: grid                                  (loads: old values of Grid.update)
  K7Map pout;   // y / 1-D: predict                         (stores)
};

static size_t k5t_smem(bool xdir, bool wide, int S, int NT, int cs = 1) {
  const size_t tile = wide ? K5T_WTILE : K5T_TILE, tab = wide ? 2 * K5T_TAB : K5T_TAB;
  const size_t per_warp = (size_t)NT * tile + (size_t)(NT + 1) * tab + (size_t)(xdir ? 1 : 2) * K5T_TILE + (size_t)NT * 8;
  // clustered strips: the first CTA gathers lambda_e and Q of all cs * S segments (every CTA of the launch is sized alike)
  const size_t gather = cs > 1 ? (size_t)2 * S * cs * 32 * sizeof(double) : 0;
  return 1024 + (size_t)S * per_warp + (size_t)2 * S * 32 * sizeof(double) + gather;
}

template <bool XDIR, bool ONED, int MAXT, bool WIDE>
__global__ void __launch_bounds__(MAXT) k5t_sweep(K5Params q, int NT, int hints, const __grid_constant__ K5TMaps tm) {
  static_assert(!WIDE || XDIR, "the wide stage is the x sweep's");
  extern __shared__ unsigned char k5traw[];
  // the swizzled boxes need a 1024-byte aligned base
  unsigned char *base = k5traw + ((1024u - (gs_smem_u32(k5traw) & 1023u)) & 1023u);
  constexpr int NOUT = XDIR ? 1 : 2;
  constexpr bool GOLD = !XDIR && !ONED;           // the old grid values come as a second box
  constexpr int CPS = WIDE ? 2 : 1;               // chunks per stage
  constexpr int TILE = WIDE ? K5T_WTILE : K5T_TILE, TAB = CPS * K5T_TAB;
  const K5Dir &d = q.d;
  const int S = d.S, m = d.m, n = d.n;
  const int lane = threadIdx.x & 31, k = threadIdx.x >> 5;   // warp k = segment crank * S + k
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  const bool soft = q.soft != 0;   // the strip's CTAs meet through global memory instead of a hardware cluster
  const int csz = q.csize, crank = csz > 1 ? (soft ? (int)(blockIdx.x % csz) : (int)cluster.block_rank()) : 0, ST = S * csz;
  const int strip_id = blockIdx.x / csz;
  auto sync_strip = [&]() { if (csz > 1 && !soft) cluster.sync(); else __syncthreads(); };
  const int line0 = strip_id * q.strip, nl = min(q.strip, d.nlines - line0);
  const int NTS = NT + 1;   // table slots
  // shared memory: output tiles [S][NOUT] | stage tiles [S][NT] | table slots [S][NT + 1] | lamE [S][32] | Q [S][32] | barriers [S][NT]
  unsigned char *otile = base + (size_t)k * NOUT * K5T_TILE;
  unsigned char *tiles = base + (size_t)S * NOUT * K5T_TILE + (size_t)k * NT * TILE;
  unsigned char *tabs = base + (size_t)S * (NOUT * K5T_TILE + NT * TILE) + (size_t)k * NTS * TAB;
  double *lamE = reinterpret_cast<double *>(base + (size_t)S * (NOUT * K5T_TILE + NT * TILE + NTS * TAB));
  double *Qs = lamE + S * 32, *Es = lamE;
  uint64_t *bars = reinterpret_cast<uint64_t *>(Qs + S * 32) + k * NT;
  double *allLam = reinterpret_cast<double *>(reinterpret_cast<uint64_t *>(Qs + S * 32) + S * NT), *allQ = allLam + ST * 32;
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
  // launched with programmatic stream serialisation: everything above overlaps the tail of the kernel before this one
  // in the stream; nothing below may run before that kernel's writes are complete and visible
  if (q.pdl) asm volatile("griddepcontrol.wait;" ::: "memory");

  const size_t ck_pitch = (size_t)((d.nlines + 31) / 32) * 32;
  const int s = (crank * S + k) * m, e = min(n, s + m) - 1;
  double *ckpt = q.ckpt + (size_t)(s / K5_CH) * ck_pitch + line0 + lane;
  const int nch = (e >= s) ? (e - s) / K5_CH + 1 : 0;   // chunks of 16 nodes
  const int nst = (nch + CPS - 1) / CPS;                // stages
  const bool act = lane < nl;
  const int line = line0 + (act ? lane : 0);
  const double ex0 = d.e0a * d.bound_first[(size_t)line * q.bc_stride] + d.e0b;
  const double ex1 = d.e1a * d.bound_last[(size_t)line * q.bc_stride] + d.e1b;
  // bytes of one load box (parts outside the array count too)
  const uint32_t tbytes = ONED ? (uint32_t)K5T_TILE : (uint32_t)q.strip * (WIDE ? (uint32_t)K5T_WPITCH : 128u);
  const int ypitch = ONED ? 32 : q.strip;   // y sweep: doubles per tile row = columns per CTA
  uint32_t phase = 0;

  // box coordinates (innermost first) of the chunk / stage that starts at node i0
  auto c0_of = [&](int i0) { return XDIR ? i0 : (ONED ? 0 : line0); };
  auto c1_of = [&](int i0) { return XDIR ? line0 : (ONED ? strip_id * n + i0 : i0); };
  // L2 hints (option k5t_hints, off: measured within noise or slower -- keeping only the half of a segment that phase C
  // re-reads first was slower too, x sweep 81 against 70 us): what phase A reads comes back in phase C (evict_last);
  // phase C's reads and all results are not touched again by this kernel (evict_first)
  const uint64_t pol_keep = gs_policy_evict_last(), pol_stream = gs_policy_evict_first();
  // lane 0: stage number `sc` of the segment (nodes s + sc * CPS * 16 ...) -> tile `st` (+ the tile after it for the old
  // grid values) and table slot `ts`
  auto issue = [&](int st, int ts, int sc, bool gold, bool last_use) {
    const int i0 = s + sc * CPS * K5_CH, len = min(CPS * K5_CH, e - i0 + 1);
    uint64_t *bar = &bars[st];
    gs_mbar_expect_tx(bar, tbytes * (gold ? 2u : 1u) + (uint32_t)len * 32u);
    if (hints & 1) gs_tma_load_2d_hint(tiles + (size_t)st * TILE, &tm.rhs, c0_of(i0), c1_of(i0), bar, last_use ? pol_stream : pol_keep);
    else gs_tma_load_2d(tiles + (size_t)st * TILE, &tm.rhs, c0_of(i0), c1_of(i0), bar);
    gs_bulk_g2s(tabs + (size_t)ts * TAB, d.tab + i0, (uint32_t)len * 32u, bar);
    if (gold) {
      if (hints & 1) gs_tma_load_2d_hint(tiles + (size_t)(st + 1) * TILE, &tm.gold, c0_of(i0), c1_of(i0), bar, pol_stream);
      else gs_tma_load_2d(tiles + (size_t)(st + 1) * TILE, &tm.gold, c0_of(i0), c1_of(i0), bar);
    }
  };
  auto wait = [&](int st) {
    gs_mbar_wait_bounded(&bars[st], (phase >> st) & 1u, q.status);
    phase ^= 1u << st;
  };
  // a lane's 16 values of chunk h of a stage tile -> registers
  auto load16 = [&](const unsigned char *tile, int h, double (&v)[K5_CH]) {
    if (WIDE) {
      const unsigned char *row = tile + lane * K5T_WPITCH + h * 128;
#pragma unroll
      for (int u = 0; u < K5_CH / 2; ++u) {
        const double2 w = *reinterpret_cast<const double2 *>(row + (u << 4));
        v[2 * u] = w.x;
        v[2 * u + 1] = w.y;
      }
    } else if (XDIR) {
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
      for (int j = 0; j < K5_CH; ++j) v[j] = col[j * ypitch];
    }
  };
  // registers -> an output tile (x: swizzled 16-column box)
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
      for (int j = 0; j < K5_CH; ++j) col[j * ypitch] = v[j];
    }
  };
  // a chunk is `plain` when it is full and holds neither node 0 nor the segment's last node e (hence not n - 1)
  auto plain = [&](int i0) { return i0 > 0 && i0 + K5_CH - 1 < e; };

  // ---- phase A: lambda_e and Q of the segment, a lambda checkpoint per chunk ----
  double lam = 0.0, Qacc = 0.0;
  if (lane == 0)
    for (int sc = 0; sc < min(NT, nst); ++sc) issue(sc, sc, sc, false, false);
  {
    int st = 0, ts = 0;
    for (int sc = 0; sc < nst; ++sc) {
      wait(st);
      const int hn = min(CPS, nch - sc * CPS);   // chunks in this stage
#pragma unroll
      for (int h = 0; h < CPS; ++h) {
        if (h < hn) {
          const int c = sc * CPS + h, i0 = s + c * K5_CH, len = min(K5_CH, e - i0 + 1);
          double tv[K5_CH];
          load16(tiles + (size_t)st * TILE, h, tv);
          if (h == hn - 1) {
            // the stage is in registers: refill it at once
            __syncwarp();
            if (lane == 0 && sc + NT < nst) issue(st, (ts + NT >= NTS) ? ts + NT - NTS : ts + NT, sc + NT, false, false);
          }
          const double4 *t = reinterpret_cast<const double4 *>(tabs + (size_t)ts * TAB + h * K5T_TAB);
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
        }
      }
      st = (st + 1 == NT) ? 0 : st + 1;
      ts = (ts + 1 == NTS) ? 0 : ts + 1;
    }
  }
  // phase C's first loads (the segment's last chunks) do not depend on the reduced system: they fly during the exchange
  constexpr int SS = GOLD ? 2 : 1;   // tiles per stage
  const int D = NT / SS;             // stages
  if (lane == 0)
    for (int kk = 0; kk < min(D, nst); ++kk) issue(kk * SS, kk % NTS, nst - 1 - kk, GOLD, true);
  // segment results go where the reduced system is solved: this CTA's arrays, or -- clustered strips -- the first
  // CTA's [ST][32] gather arrays through distributed shared memory, or -- soft groups -- global memory
  double *gLam = nullptr, *gQ = nullptr, *gE = nullptr;
  if (csz > 1 && soft) {
    const size_t per = (size_t)(gridDim.x / csz) * ST * 32;
    gLam = q.g_buf + (size_t)strip_id * ST * 32;
    gQ = gLam + per;
    gE = gQ + per;
    const int kg = crank * S + k;
    __stcg(&gLam[kg * 32 + lane], lam);
    __stcg(&gQ[kg * 32 + lane], Qacc);
    __threadfence();
  } else if (csz > 1) {
    const int kg = crank * S + k;
    cluster.map_shared_rank(allLam, 0)[kg * 32 + lane] = lam;
    cluster.map_shared_rank(allQ, 0)[kg * 32 + lane] = Qacc;
  } else {
    lamE[k * 32 + lane] = lam;
    Qs[k * 32 + lane] = Qacc;
  }
  sync_strip();
  if (csz > 1 && soft) {
    // arrive; the strip's first CTA waits for all csz arrivals of this launch (the counter is never reset: launch
    // number `epoch` ends at epoch * csz), solves, publishes; everybody waits for its `done` word
    unsigned *arrive = q.g_sync + 2 * strip_id, *done = arrive + 1;
    if (threadIdx.x == 0) {
      asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(arrive) : "memory");
    }
    if (crank == 0) {
      // the whole first CTA copies the strip's [ST][32] segment results into shared memory (its output tiles are idle
      // until phase C stores) in ONE round trip to L2, warp 0 solves there, and all warps publish the result
      if (threadIdx.x == 0) gs_spin_until(arrive, q.epoch * (unsigned)csz, q.status);
      __syncthreads();
      double *sl = reinterpret_cast<double *>(base), *sq = sl + ST * 32;   // the CTA's output tiles (idle until phase C stores)
      for (int i = threadIdx.x; i < ST * 32; i += blockDim.x) {
        sl[i] = __ldcg(&gLam[i]);
        sq[i] = __ldcg(&gQ[i]);
      }
      __syncthreads();
      if (k == 0) {
        double dp = 0.0;
        for (int kk = 0; kk < ST; ++kk) {
          const double4 sg = d.seg[kk];
          const double rhs = sl[kk * 32 + lane] + ((kk + 1 < ST) ? sg.x * sq[(kk + 1) * 32 + lane] : 0.0);
          dp = (rhs - sg.y * dp) * sg.z;
          sl[kk * 32 + lane] = dp;
        }
        double x = 0.0;
        for (int kk = ST - 1; kk >= 0; --kk) {
          x = sl[kk * 32 + lane] - d.seg[kk].w * x;
          sl[kk * 32 + lane] = x;
        }
      }
      __syncthreads();
      for (int i = threadIdx.x; i < ST * 32; i += blockDim.x) __stcg(&gE[i], sl[i]);
      __threadfence();
      __syncthreads();
      if (threadIdx.x == 0) asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(done), "r"(q.epoch) : "memory");
    }
    if (lane == 0) gs_spin_until(done, q.epoch, q.status);
    __syncwarp();
    const int kg = crank * S + k;
    Es[k * 32 + lane] = __ldcg(&gE[kg * 32 + lane]);
    if (k == 0) Qs[lane] = kg > 0 ? __ldcg(&gE[(kg - 1) * 32 + lane]) : 0.0;
    __syncthreads();
  } else {
  // ---- reduced system (first CTA's warp 0; lane = line).  Afterwards Es[k] = x at the end of segment k and
  //      Qs[0 * 32 + lane] of every CTA = x just above its first segment ----
  if (k == 0 && crank == 0) {
    double *gl = csz > 1 ? allLam : lamE, *gq = csz > 1 ? allQ : Qs;
    double dp = 0.0;
    for (int kk = 0; kk < ST; ++kk) {
      const double4 sg = d.seg[kk];
      const double rhs = gl[kk * 32 + lane] + ((kk + 1 < ST) ? sg.x * gq[(kk + 1) * 32 + lane] : 0.0);
      dp = (rhs - sg.y * dp) * sg.z;
      gl[kk * 32 + lane] = dp;
    }
    double x = 0.0;
    for (int kk = ST - 1; kk >= 0; --kk) {
      x = gl[kk * 32 + lane] - d.seg[kk].w * x;
      if (csz > 1) {
        cluster.map_shared_rank(Es, kk / S)[(kk % S) * 32 + lane] = x;
        if (kk % S == S - 1 && kk + 1 < ST) cluster.map_shared_rank(Qs, (kk + 1) / S)[lane] = x;
      } else {
        Es[kk * 32 + lane] = x;
      }
    }
    Qs[lane] = 0.0;
  }
  sync_strip();
  }
  // ---- phase C: chunks last to first: recompute the lambdas from the checkpoint, back-substitute, store ----
  if (nch == 0) return;
  const double L = (k > 0) ? Es[(k - 1) * 32 + lane] : Qs[lane];   // x just before the segment
  double x = Es[k * 32 + lane];                                // x_e
  double ck_next = (nch > 1) ? ckpt[(size_t)(nch - 2) * ck_pitch] : 0.0;               // checkpoint before the last chunk
  double om_next = (nch > 1) ? d.omega_end[(s + (nch - 1) * K5_CH - 1) / K5_CH] : 0.0;  // omega at its end
  {
    int st = 0, ts = 0;
    for (int sc = nst - 1; sc >= 0; --sc) {
      wait(st * SS);
      const int hn = min(CPS, nch - sc * CPS);
#pragma unroll
      for (int h = CPS - 1; h >= 0; --h) {
        if (h < hn) {
          const int c = sc * CPS + h, i0 = s + c * K5_CH, len = min(K5_CH, e - i0 + 1);
          double tv[K5_CH], gv[K5_CH];
          load16(tiles + (size_t)(st * SS) * TILE, h, tv);
          if (GOLD) load16(tiles + (size_t)(st * SS + 1) * TILE, h, gv);
          // the stores issued from the output tiles one chunk ago have read them
          if (lane == 0) gs_bulk_wait_read<0>();
          __syncwarp();
          if (h == 0 && lane == 0 && sc - D >= 0)
            issue(st * SS, (ts + D >= NTS) ? ts + D - NTS : ts + D, sc - D, GOLD, true);
          const double4 *t = reinterpret_cast<const double4 *>(tabs + (size_t)ts * TAB + h * K5T_TAB);
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
            if (hints & 2) {
              gs_tma_store_2d_hint(&tm.out, c0_of(i0), c1_of(i0), otile, pol_stream);
              if (!XDIR) gs_tma_store_2d_hint(&tm.pout, c0_of(i0), c1_of(i0), otile + K5T_TILE, pol_stream);
            } else {
              gs_tma_store_2d(&tm.out, c0_of(i0), c1_of(i0), otile);
              if (!XDIR) gs_tma_store_2d(&tm.pout, c0_of(i0), c1_of(i0), otile + K5T_TILE);
            }
            gs_bulk_commit();
          }
        }
      }
      st = (st + 1 == D) ? 0 : st + 1;
      ts = (ts + 1 == NTS) ? 0 : ts + 1;
    }
  }
  if (lane == 0) gs_bulk_wait_read<0>();
  if (act && !(fabs(x) <= 1.7976931348623157e308)) atomicOr(q.status, GS_FLAG_NONFINITE);
}
