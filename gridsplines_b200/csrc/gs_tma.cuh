// gs_tma.cuh -- thin wrappers over the sm_100a bulk-copy engine (TMA, cp.async.bulk) and mbarriers.
// Per-warp streaming pipelines in gs_grid1d.cu / gs_grid2d.cu are built from these.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define GS_TMA_DEV __device__ __forceinline__

GS_TMA_DEV uint32_t gs_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

GS_TMA_DEV void gs_mbar_init(uint64_t *bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(gs_smem_u32(bar)), "r"(count) : "memory");
}
GS_TMA_DEV void gs_mbar_init_fence() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
GS_TMA_DEV void gs_mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(gs_smem_u32(bar)), "r"(bytes) : "memory");
}
GS_TMA_DEV void gs_mbar_wait(uint64_t *bar, uint32_t parity) {
  uint32_t addr = gs_smem_u32(bar), done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
  } while (!done);
}
// transaction bytes announced without an arrival (the arrival comes later, from gs_mbar_arrive)
GS_TMA_DEV void gs_mbar_expect_tx_only(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;" ::"r"(gs_smem_u32(bar)), "r"(bytes) : "memory");
}
GS_TMA_DEV void gs_mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(gs_smem_u32(bar)) : "memory");
}
// global -> shared, completion signalled on an mbarrier (bytes: multiple of 16, both addresses 16 B aligned)
GS_TMA_DEV void gs_bulk_g2s(void *sdst, const void *gsrc, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   gs_smem_u32(sdst)),
               "l"(gsrc), "r"(bytes), "r"(gs_smem_u32(bar))
               : "memory");
}
// shared -> global, tracked by bulk async-groups
GS_TMA_DEV void gs_bulk_s2g(void *gdst, const void *ssrc, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(gs_smem_u32(ssrc)),
               "r"(bytes)
               : "memory");
}
GS_TMA_DEV void gs_bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// wait until at most N groups still have their SOURCE (shared memory) unread
template <int N>
GS_TMA_DEV void gs_bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
// wait until at most N groups are incomplete (writes performed)
template <int N>
GS_TMA_DEV void gs_bulk_wait() { asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory"); }
// orders this thread's generic-proxy accesses before subsequent async-proxy (bulk copy) accesses
GS_TMA_DEV void gs_fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }
// same, restricted to shared memory (what a bulk store's source needs)
GS_TMA_DEV void gs_fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// L2 eviction-priority policies for bulk copies
GS_TMA_DEV uint64_t gs_policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
GS_TMA_DEV uint64_t gs_policy_evict_last() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
GS_TMA_DEV void gs_bulk_g2s_hint(void *sdst, const void *gsrc, uint32_t bytes, uint64_t *bar, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
          gs_smem_u32(sdst)),
      "l"(gsrc), "r"(bytes), "r"(gs_smem_u32(bar)), "l"(policy)
      : "memory");
}
GS_TMA_DEV void gs_bulk_s2g_hint(void *gdst, const void *ssrc, uint32_t bytes, uint64_t policy) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(gdst),
               "r"(gs_smem_u32(ssrc)), "r"(bytes), "l"(policy)
               : "memory");
}

// ---- tensor-map TMA (cp.async.bulk.tensor): one instruction moves a 2-D box between a row-major global array and
// shared memory; parts of the box outside the array are zero-filled on loads (the mbarrier still counts the whole
// box) and clipped on stores.  `tmap` points to a CUtensorMap (128 B, 64 B aligned) in kernel-parameter space
// (__grid_constant__); c0 = innermost coordinate (column), c1 = row.
GS_TMA_DEV void gs_tma_load_2d(void *sdst, const void *tmap, int c0, int c1, uint64_t *bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
          gs_smem_u32(sdst)),
      "l"(tmap), "r"(c0), "r"(c1), "r"(gs_smem_u32(bar))
      : "memory");
}
GS_TMA_DEV void gs_tma_store_2d(const void *tmap, int c0, int c1, const void *ssrc) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];" ::"l"(tmap), "r"(c0),
               "r"(c1), "r"(gs_smem_u32(ssrc))
               : "memory");
}
GS_TMA_DEV void gs_tma_load_2d_hint(void *sdst, const void *tmap, int c0, int c1, uint64_t *bar, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3}], [%4], %5;" ::"r"(
          gs_smem_u32(sdst)),
      "l"(tmap), "r"(c0), "r"(c1), "r"(gs_smem_u32(bar)), "l"(policy)
      : "memory");
}
GS_TMA_DEV void gs_tma_store_2d_hint(const void *tmap, int c0, int c1, const void *ssrc, uint64_t policy) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group.L2::cache_hint [%0, {%1, %2}], [%3], %4;" ::"l"(tmap),
               "r"(c0), "r"(c1), "r"(gs_smem_u32(ssrc)), "l"(policy)
               : "memory");
}
GS_TMA_DEV void gs_tma_prefetch_desc(const void *tmap) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(tmap) : "memory");
}
// mbarrier wait that cannot hang the device: after ~2^32 cycles without completion it raises bit 31 of *status and traps
GS_TMA_DEV void gs_mbar_wait_bounded(uint64_t *bar, uint32_t parity, unsigned *status) {
  uint32_t addr = gs_smem_u32(bar), done;
  long long t0 = 0;
  int spins = 0;
  for (;;) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
    if (done) return;
    if (++spins == 64) t0 = clock64();
    if (spins > 64 && (spins & 255) == 0 && clock64() - t0 > (1ll << 32)) {
      atomicOr(status, 0x80000000u);
      __trap();
    }
  }
}

// spin on a global word (acquire, gpu scope) until it equals `want`; as gs_mbar_wait_bounded it cannot hang the device
GS_TMA_DEV void gs_spin_until(const unsigned *word, unsigned want, unsigned *status) {
  long long t0 = 0;
  int spins = 0;
  for (;;) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(word) : "memory");
    if (v == want) return;
    if (++spins == 64) t0 = clock64();
    if (spins > 64 && (spins & 255) == 0 && clock64() - t0 > (1ll << 32)) {
      atomicOr(status, 0x80000000u);
      __trap();
    }
  }
}
