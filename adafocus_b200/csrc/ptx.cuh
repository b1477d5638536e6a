// Thin inline-PTX wrappers for sm_100a: mbarrier, TMA (cp.async.bulk.tensor), tcgen05 (MMA / TMEM).
// Everything here is device-only and header-only.
#pragma once
#include <cstdint>
#include <cuda.h>
#include <cuda_fp16.h>

namespace af {
namespace ptx {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "elect.sync _|P1, 0xffffffff;\n\t"
      "selp.b32 %0, 1, 0, P1;\n\t"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}

// ---------------------------------------------------------------- mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, P1;\n\t"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Wait with a suspend-time hint: the hardware parks the warp (no issue slots taken from the other warps of its
// scheduler) until the phase completes or `ns` nanoseconds have passed.
__device__ __forceinline__ void mbar_wait_parked(uint64_t* bar, uint32_t parity, uint32_t ns) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2, %3;\n\t"
        "selp.b32 %0, 1, 0, P1;\n\t"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity), "r"(ns)
        : "memory");
  } while (ok == 0);
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}

// ---------------------------------------------------------------- TMA
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      :
      : "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1,
                                            int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], "
      "[%2];"
      :
      : "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2),
        "r"(c3)
      : "memory");
}

// Warp-converged forms (all lanes execute, one elected lane issues): see umma_f16_ss_lo_elect.
__device__ __forceinline__ void mbar_arrive_expect_tx_elect(uint64_t* bar, uint32_t bytes) {
  asm volatile(
      "{\n\t"
      ".reg .pred pe;\n\t"
      "elect.sync _|pe, 0xffffffff;\n\t"
      "@pe mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n\t"
      "}\n" ::"r"(smem_u32(bar)), "r"(bytes)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d_elect(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "{\n\t"
      ".reg .pred pe;\n\t"
      "elect.sync _|pe, 0xffffffff;\n\t"
      "@pe cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];\n\t"
      "}\n"
      :
      : "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d_elect(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1,
                                                  int c2, int c3) {
  asm volatile(
      "{\n\t"
      ".reg .pred pe;\n\t"
      "elect.sync _|pe, 0xffffffff;\n\t"
      "@pe cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], "
      "[%2];\n\t"
      "}\n"
      :
      : "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2),
        "r"(c3)
      : "memory");
}

__device__ __forceinline__ void tma_load_5d_elect(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1,
                                                  int c2, int c3, int c4) {
  asm volatile(
      "{\n\t"
      ".reg .pred pe;\n\t"
      "elect.sync _|pe, 0xffffffff;\n\t"
      "@pe cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], "
      "[%2];\n\t"
      "}\n"
      :
      : "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2),
        "r"(c3), "r"(c4)
      : "memory");
}

// TMA store: smem tile (128-B swizzled box) -> global tensor; out-of-bounds parts of the box are not written.
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* m, const void* smem_src, int c0, int c1, int c2,
                                             int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
      :
      : "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// wait until the bulk stores of all committed groups have finished READING shared memory
__device__ __forceinline__ void tma_store_wait_read0() {
  asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
// same, but the most recently committed group may still be reading
__device__ __forceinline__ void tma_store_wait_read1() {
  asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
}
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void named_barrier_sync(uint32_t id, uint32_t nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// ---------------------------------------------------------------- tcgen05 / TMEM
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// D[tmem] (+)= A[smem] * B[smem]; fp16/bf16 operands, fp32 accumulate. One thread issues.
__device__ __forceinline__ void umma_f16_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n"
      :
      : "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Same with the two shared-memory descriptors given as their low words (the high word of a K-major SWIZZLE_128B
// descriptor is a constant): the issuing thread then only needs 32-bit adds between MMAs.
constexpr uint32_t kDescHiSw128 = (1024u >> 4) | (1u << 14) | (2u << 29);   // SBO = 1024 B | version 1 | SWIZZLE_128B
__device__ __forceinline__ uint32_t smem_desc_lo(uint32_t smem_addr) {
  return ((smem_addr & 0x3FFFFu) >> 4) | (1u << 16);
}
__device__ __forceinline__ void umma_f16_ss_lo(uint32_t tmem_d, uint32_t lo_a, uint32_t lo_b, uint32_t idesc,
                                               uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      ".reg .b64 da, db;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "mov.b64 da, {%1, %5};\n\t"
      "mov.b64 db, {%2, %5};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, p;\n\t"
      "}\n"
      :
      : "r"(tmem_d), "r"(lo_a), "r"(lo_b), "r"(idesc), "r"(accumulate), "r"(kDescHiSw128)
      : "memory");
}
// Warp-converged forms: every lane of the issuing warp executes the call, one elected lane issues.  Keeping the
// issuer warp converged lets the compiler hold descriptors in uniform registers instead of wrapping every UTCHMMA in
// a per-active-lane loop.
__device__ __forceinline__ void umma_f16_ss_lo_elect(uint32_t tmem_d, uint32_t lo_a, uint32_t lo_b, uint32_t idesc,
                                                     uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p, pe;\n\t"
      ".reg .b64 da, db;\n\t"
      "elect.sync _|pe, 0xffffffff;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "mov.b64 da, {%1, %5};\n\t"
      "mov.b64 db, {%2, %5};\n\t"
      "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, p;\n\t"
      "}\n"
      :
      : "r"(tmem_d), "r"(lo_a), "r"(lo_b), "r"(idesc), "r"(accumulate), "r"(kDescHiSw128)
      : "memory");
}
// One accumulator tile in one go: K_STEPS (1..4) MMAs over consecutive 16-wide K slices (32 B apart in both operand
// tiles), the first one overwriting D, then a commit on `bar_addr`; a single election for the whole group.  K_STEPS is
// a template parameter so that the issuing warp sets up exactly the descriptors it uses (its instruction stream is the
// critical resource: ~10 cycles per instruction next to busy compute warps).
template <int K_STEPS>
__device__ __forceinline__ void umma_group_commit_elect(uint32_t tmem_d, uint32_t lo_a, uint32_t lo_b, uint32_t idesc,
                                                        uint32_t bar_addr) {
  static_assert(K_STEPS >= 1 && K_STEPS <= 4, "K_STEPS");
  if (K_STEPS == 1) {
    asm volatile(
        "{\n\t"
        ".reg .pred pe, pf;\n\t"
        ".reg .b64 da, db;\n\t"
        "elect.sync _|pe, 0xffffffff;\n\t"
        "setp.ne.u32 pf, %3, %3;\n\t"
        "mov.b64 da, {%1, %4};\n\t"
        "mov.b64 db, {%2, %4};\n\t"
        "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, pf;\n\t"
        "@pe tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%5];\n\t"
        "}\n"
        :
        : "r"(tmem_d), "r"(lo_a), "r"(lo_b), "r"(idesc), "r"(kDescHiSw128), "r"(bar_addr)
        : "memory");
  } else if (K_STEPS == 2) {
    asm volatile(
        "{\n\t"
        ".reg .pred pe, pt, pf;\n\t"
        ".reg .b64 da, db, da1, db1;\n\t"
        "elect.sync _|pe, 0xffffffff;\n\t"
        "setp.eq.u32 pt, %3, %3;\n\t"
        "setp.ne.u32 pf, %3, %3;\n\t"
        "mov.b64 da, {%1, %4};\n\t"
        "mov.b64 db, {%2, %4};\n\t"
        "mov.b64 da1, {%6, %4};\n\t"
        "mov.b64 db1, {%7, %4};\n\t"
        "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, pf;\n\t"
        "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], da1, db1, %3, pt;\n\t"
        "@pe tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%5];\n\t"
        "}\n"
        :
        : "r"(tmem_d), "r"(lo_a), "r"(lo_b), "r"(idesc), "r"(kDescHiSw128), "r"(bar_addr), "r"(lo_a + 2u), "r"(lo_b + 2u)
        : "memory");
  } else {
    asm volatile(
        "{\n\t"
        ".reg .pred pe, p4, pt, pf;\n\t"
        ".reg .b64 da, db, da1, db1, da2, db2, da3, db3;\n\t"
        "elect.sync _|pe, 0xffffffff;\n\t"
        "setp.eq.u32 pt, %3, %3;\n\t"
        "setp.ne.u32 pf, %3, %3;\n\t"
        "setp.ne.u32 p4, %12, 0;\n\t"
        "and.pred p4, p4, pe;\n\t"
        "mov.b64 da, {%1, %4};\n\t"
        "mov.b64 db, {%2, %4};\n\t"
        "mov.b64 da1, {%6, %4};\n\t"
        "mov.b64 db1, {%7, %4};\n\t"
        "mov.b64 da2, {%8, %4};\n\t"
        "mov.b64 db2, {%9, %4};\n\t"
        "mov.b64 da3, {%10, %4};\n\t"
        "mov.b64 db3, {%11, %4};\n\t"
        "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, pf;\n\t"
        "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], da1, db1, %3, pt;\n\t"
        "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], da2, db2, %3, pt;\n\t"
        "@p4 tcgen05.mma.cta_group::1.kind::f16 [%0], da3, db3, %3, pt;\n\t"
        "@pe tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%5];\n\t"
        "}\n"
        :
        : "r"(tmem_d), "r"(lo_a), "r"(lo_b), "r"(idesc), "r"(kDescHiSw128), "r"(bar_addr), "r"(lo_a + 2u), "r"(lo_b + 2u),
          "r"(lo_a + 4u), "r"(lo_b + 4u), "r"(lo_a + 6u), "r"(lo_b + 6u), "r"(K_STEPS == 4 ? 1u : 0u)
        : "memory");
  }
}
__device__ __forceinline__ void umma_commit_elect(uint64_t* bar) {
  asm volatile(
      "{\n\t"
      ".reg .pred pe;\n\t"
      "elect.sync _|pe, 0xffffffff;\n\t"
      "@pe tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t"
      "}\n" ::"r"(smem_u32(bar))
      : "memory");
}
// Arrive on an mbarrier once all previously issued tcgen05.mma of this thread have completed.
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
// 32 lanes x 16 consecutive fp32 columns: thread i of the warp gets lane (taddr.lane + i).
__device__ __forceinline__ void tmem_ld_32x32b_x16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// 256-bit read-only global load (sm_100: LDG.E.256): two consecutive 16-byte chunks of one row
__device__ __forceinline__ void ldg256_nc(const void* p, uint4& lo, uint4& hi) {
  asm volatile("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(lo.x), "=r"(lo.y), "=r"(lo.z), "=r"(lo.w), "=r"(hi.x), "=r"(hi.y), "=r"(hi.z), "=r"(hi.w)
               : "l"(p));
}

// Programmatic dependent launch: let the next kernel of the stream be scheduled early / wait for the previous
// kernel's completion (and memory flush) before touching anything it produced.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait_prior_grid() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// ---------------------------------------------------------------- CTA pair (cta_group::2)
// Two CTAs of a cluster on the two SMs of a TPC run ONE tcgen05.mma of M = 256: each CTA supplies its own 128 rows of
// A and HALF of the N rows of B from the same shared-memory offsets and receives its 128 rows of D in its own TMEM.
// Only the leader (cluster rank 0) issues MMAs / commits; both CTAs issue their own TMA loads, which complete on the
// leader's mbarrier (".cta_group::2" form of the bulk-tensor copy).
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `p` (a shared-memory object of this CTA) in the CTA of rank `cta`
__device__ __forceinline__ uint32_t mapa_u32(const void* p, uint32_t cta) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_u32(p)), "r"(cta));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish_pair() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_f16_ss_lo_elect_pair(uint32_t tmem_d, uint32_t lo_a, uint32_t lo_b, uint32_t idesc,
                                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p, pe;\n\t"
      ".reg .b64 da, db;\n\t"
      "elect.sync _|pe, 0xffffffff;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "mov.b64 da, {%1, %5};\n\t"
      "mov.b64 db, {%2, %5};\n\t"
      "@pe tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %3, p;\n\t"
      "}\n"
      :
      : "r"(tmem_d), "r"(lo_a), "r"(lo_b), "r"(idesc), "r"(accumulate), "r"(kDescHiSw128)
      : "memory");
}
// arrive (once all previously issued MMAs of this thread have completed) on the barrier at the SAME shared-memory
// offset in both CTAs of the pair
__device__ __forceinline__ void umma_commit_elect_pair(uint64_t* bar) {
  asm volatile(
      "{\n\t"
      ".reg .pred pe;\n\t"
      ".reg .b16 m;\n\t"
      "mov.b16 m, 3;\n\t"
      "elect.sync _|pe, 0xffffffff;\n\t"
      "@pe tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], m;\n\t"
      "}\n" ::"r"(smem_u32(bar))
      : "memory");
}
// bulk-tensor loads into THIS CTA's shared memory whose completion bytes are counted on `bar_cluster_addr`, an
// mbarrier of the pair's leader CTA (shared::cluster address, see mapa_u32)
__device__ __forceinline__ void tma_load_2d_elect_pair(void* smem_dst, const CUtensorMap* m, uint32_t bar_cluster_addr,
                                                       int c0, int c1) {
  asm volatile(
      "{\n\t"
      ".reg .pred pe;\n\t"
      "elect.sync _|pe, 0xffffffff;\n\t"
      "@pe cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], "
      "[%2];\n\t"
      "}\n"
      :
      : "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d_elect_pair(void* smem_dst, const CUtensorMap* m, uint32_t bar_cluster_addr,
                                                       int c0, int c1, int c2, int c3) {
  asm volatile(
      "{\n\t"
      ".reg .pred pe;\n\t"
      "elect.sync _|pe, 0xffffffff;\n\t"
      "@pe cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, "
      "%5, %6}], [%2];\n\t"
      "}\n"
      :
      : "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster_addr), "r"(c0), "r"(c1), "r"(c2),
        "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_5d_elect_pair(void* smem_dst, const CUtensorMap* m, uint32_t bar_cluster_addr,
                                                       int c0, int c1, int c2, int c3, int c4) {
  asm volatile(
      "{\n\t"
      ".reg .pred pe;\n\t"
      "elect.sync _|pe, 0xffffffff;\n\t"
      "@pe cp.async.bulk.tensor.5d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, "
      "%5, %6, %7}], [%2];\n\t"
      "}\n"
      :
      : "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster_addr), "r"(c0), "r"(c1), "r"(c2),
        "r"(c3), "r"(c4)
      : "memory");
}

// ---------------------------------------------------------------- descriptors
// Shared-memory matrix descriptor for a K-major operand tile whose rows are 128 B (64 fp16) wide and
// laid out by TMA with CU_TENSOR_MAP_SWIZZLE_128B: 8-row groups are 1024 B apart (SBO), LBO unused.
// Bit layout (sm_100 "SmemDescriptor"): [0,14) addr>>4 | [16,30) LBO>>4 | [32,46) SBO>>4 | [46,48) version=1 |
// [49,52) base offset | [61,64) layout (2 = SWIZZLE_128B).
__device__ __forceinline__ uint64_t make_smem_desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>(1) << 16;            // LBO (ignored for swizzled K-major)
  d |= static_cast<uint64_t>(1024 >> 4) << 32;    // SBO = 1024 B
  d |= static_cast<uint64_t>(1) << 46;            // descriptor version (Blackwell)
  d |= static_cast<uint64_t>(2) << 61;            // SWIZZLE_128B
  return d;
}
// Instruction descriptor, kind::f16: fp16 A and B (both K-major), fp32 accumulator, M x N tile.
__host__ __device__ __forceinline__ uint32_t make_idesc_f16_f32(uint32_t m, uint32_t n) {
  uint32_t d = 0;
  d |= 1u << 4;          // C format = F32
  d |= 0u << 7;          // A format = F16
  d |= 0u << 10;         // B format = F16
  d |= 0u << 15;         // A K-major
  d |= 0u << 16;         // B K-major
  d |= (n >> 3) << 17;   // N / 8
  d |= (m >> 4) << 24;   // M / 16
  return d;
}

}  // namespace ptx
}  // namespace af
