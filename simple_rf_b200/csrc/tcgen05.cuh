// Thin inline-PTX layer for sm_100a: mbarrier, bulk async copy (TMA, non-tensor form), tcgen05
// (TMEM alloc / MMA / commit / ld) and the UMMA shared-memory + instruction descriptors.
#pragma once
#include <cuda_bf16.h>
#include <stdint.h>

namespace srf {
namespace ptx {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---------------------------------------------------------------- mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// try_wait carries a suspend-time hint (as CUTLASS's ClusterBarrier::wait does): without it a waiting warp comes back from the
// instruction almost immediately and its polling loop takes issue slots from the warps it is waiting for
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, 0x989680;\n\t"
      "@p bra WAIT_DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "WAIT_DONE:\n\t"
      "}" ::"r"(smem_u32(bar)), "r"(parity)
      : "memory");
}

// explicit shared-window forms (32-bit shared addresses computed once per role): generic pointers into dynamic shared
// memory make the compiler rebuild the shared-window base (S2UR SR_CgaCtaId + ULEA) inside hot loops
__device__ __forceinline__ float4 lds4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void mbar_arrive_addr(uint32_t bar_smem_addr) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar_smem_addr) : "memory");
}

// ---------------------------------------------------------------- proxies / fences
// generic-proxy smem writes -> visible to the async proxy (tcgen05.mma / bulk copies)
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// ---------------------------------------------------------------- bulk async copy global -> smem (UBLKCP)
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// L2 eviction policies for the bulk copies: the weight images are re-read by every tile of every CTA and must survive in L2 next to
// the activation / gradient tile streams (GBs per launch, each byte touched once)
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ void bulk_g2s_hint(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar, uint64_t policy) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
               : "memory");
}
__device__ __forceinline__ void bulk_s2g_hint(void* dst_gmem, const void* src_smem, uint32_t bytes, uint64_t policy) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(dst_gmem), "r"(smem_u32(src_smem)),
               "r"(bytes), "l"(policy)
               : "memory");
}

// shared -> global bulk copy (bulk async-group completion): issue, commit the group, wait until at most N groups are still
// READING their shared-memory source / until all groups have completed
__device__ __forceinline__ void bulk_s2g(void* dst_gmem, const void* src_smem, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem), "r"(smem_u32(src_smem)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// ---------------------------------------------------------------- TMEM
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t cols) {   // whole warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(cols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {     // whole warp
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}

// D[tmem] (+)= A * B^T, bf16 x bf16 -> fp32.  Forms for a CONVERGED warp (uniform control flow keeps descriptors in uniform registers and avoids the per-instruction
// election loops the compiler wraps around tcgen05 ops inside a divergent `if (lane == 0)`): every lane executes the
// statement, `issue` (from elect_one) is non-zero in exactly one lane.
__device__ __forceinline__ uint32_t elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
  return pred;
}
__device__ __forceinline__ void umma_bf16_if(uint32_t issue, uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p, q;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "setp.ne.b32 q, %5, 0;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(issue)
      : "memory");
}
// One 64-wide K block = up to four K=16 MMAs in ONE statement: the descriptors are passed as their 32-bit low words
// (start address >> 4; the constant high word is attached inside), so the per-MMA 64-bit carry chains and the repeated
// register -> uniform-register moves of four separate statements disappear from the issuing warp's critical path.
// `ksteps` (1..4) MMAs are issued; the first one overwrites the accumulator when accumulate == 0.
__device__ __forceinline__ void umma4_bf16_if(uint32_t issue, uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t desc_hi,
                                              uint32_t idesc, uint32_t accumulate, uint32_t ksteps) {
  asm volatile(
      "{\n\t"
      ".reg .pred pacc, ptrue, q0, q1, q2, q3;\n\t"
      ".reg .b32 a1, a2, a3, b1, b2, b3;\n\t"
      ".reg .b64 da0, da1, da2, da3, db0, db1, db2, db3;\n\t"
      "setp.ne.b32 pacc, %5, 0;\n\t"
      "setp.eq.b32 ptrue, 0, 0;\n\t"
      "setp.ne.b32 q0, %6, 0;\n\t"
      "setp.gt.and.u32 q1, %7, 1, q0;\n\t"
      "setp.gt.and.u32 q2, %7, 2, q0;\n\t"
      "setp.gt.and.u32 q3, %7, 3, q0;\n\t"
      "add.u32 a1, %1, 2;\n\t"
      "add.u32 a2, %1, 4;\n\t"
      "add.u32 a3, %1, 6;\n\t"
      "add.u32 b1, %2, 2;\n\t"
      "add.u32 b2, %2, 4;\n\t"
      "add.u32 b3, %2, 6;\n\t"
      "mov.b64 da0, {%1, %3};\n\t"
      "mov.b64 da1, {a1, %3};\n\t"
      "mov.b64 da2, {a2, %3};\n\t"
      "mov.b64 da3, {a3, %3};\n\t"
      "mov.b64 db0, {%2, %3};\n\t"
      "mov.b64 db1, {b1, %3};\n\t"
      "mov.b64 db2, {b2, %3};\n\t"
      "mov.b64 db3, {b3, %3};\n\t"
      "@q0 tcgen05.mma.cta_group::1.kind::f16 [%0], da0, db0, %4, pacc;\n\t"
      "@q1 tcgen05.mma.cta_group::1.kind::f16 [%0], da1, db1, %4, ptrue;\n\t"
      "@q2 tcgen05.mma.cta_group::1.kind::f16 [%0], da2, db2, %4, ptrue;\n\t"
      "@q3 tcgen05.mma.cta_group::1.kind::f16 [%0], da3, db3, %4, ptrue;\n\t"
      "}" ::"r"(d_tmem),
      "r"(a_lo), "r"(b_lo), "r"(desc_hi), "r"(idesc), "r"(accumulate), "r"(issue), "r"(ksteps)
      : "memory");
}
// Same, A operand in tensor memory (".ts" form): a_tmem is the TMEM address of the 128-lane x 8-column block holding the
// K = 16 slice of A as packed bf16 pairs (row m in lane m, elements 2c / 2c+1 in the low / high half of column c);
// the four slices of a 64-wide K block sit at a_tmem + {0, 8, off2, off2 + 8}.
__device__ __forceinline__ void umma4_bf16_ts_if(uint32_t issue, uint32_t d_tmem, uint32_t a_tmem, uint32_t off2, uint32_t b_lo,
                                                 uint32_t desc_hi, uint32_t idesc, uint32_t accumulate, uint32_t ksteps) {
  asm volatile(
      "{\n\t"
      ".reg .pred pacc, ptrue, q0, q1, q2, q3;\n\t"
      ".reg .b32 a1, a2, a3, b1, b2, b3;\n\t"
      ".reg .b64 db0, db1, db2, db3;\n\t"
      "setp.ne.b32 pacc, %5, 0;\n\t"
      "setp.eq.b32 ptrue, 0, 0;\n\t"
      "setp.ne.b32 q0, %6, 0;\n\t"
      "setp.gt.and.u32 q1, %7, 1, q0;\n\t"
      "setp.gt.and.u32 q2, %7, 2, q0;\n\t"
      "setp.gt.and.u32 q3, %7, 3, q0;\n\t"
      "add.u32 a1, %1, 8;\n\t"
      "add.u32 a2, %1, %8;\n\t"
      "add.u32 a3, a2, 8;\n\t"
      "add.u32 b1, %2, 2;\n\t"
      "add.u32 b2, %2, 4;\n\t"
      "add.u32 b3, %2, 6;\n\t"
      "mov.b64 db0, {%2, %3};\n\t"
      "mov.b64 db1, {b1, %3};\n\t"
      "mov.b64 db2, {b2, %3};\n\t"
      "mov.b64 db3, {b3, %3};\n\t"
      "@q0 tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db0, %4, pacc;\n\t"
      "@q1 tcgen05.mma.cta_group::1.kind::f16 [%0], [a1], db1, %4, ptrue;\n\t"
      "@q2 tcgen05.mma.cta_group::1.kind::f16 [%0], [a2], db2, %4, ptrue;\n\t"
      "@q3 tcgen05.mma.cta_group::1.kind::f16 [%0], [a3], db3, %4, ptrue;\n\t"
      "}" ::"r"(d_tmem),
      "r"(a_tmem), "r"(b_lo), "r"(desc_hi), "r"(idesc), "r"(accumulate), "r"(issue), "r"(ksteps), "r"(off2)
      : "memory");
}
// One whole issuer step in ONE statement, for two issuing warps that alternate steps: descriptor / predicate set-up,
// then the named-barrier token (SYNC_ID: the other warp has issued the previous step), the MMAs, the commits
// (bar_empty always; bar_x / bar_y / bar_z when non-zero), and the token for the other warp (ARRIVE_ID).  Keeping the
// set-up inside the statement lets ptxas place it ahead of the barrier, so only the UTCHMMA / UTCBAR issue sits between
// the two tokens.  TS: A operand from tensor memory (slices at a + {0, 8, 32, 40}); otherwise from shared memory
// (descriptor low word a, slices at a + {0, 2, 4, 6}).
template <int SYNC_ID, int ARRIVE_ID, bool TS>
__device__ __forceinline__ void umma4_step(uint32_t issue, uint32_t d_tmem, uint32_t a, uint32_t b_lo, uint32_t desc_hi,
                                           uint32_t idesc, uint32_t accumulate, uint32_t ksteps, uint32_t bar_empty,
                                           uint32_t bar_x, uint32_t bar_y, uint32_t bar_z) {
  if constexpr (TS) {
    asm volatile(
        "{\n\t"
        ".reg .pred pacc, ptrue, q0, q1, q2, q3, cx, cy, cz;\n\t"
        ".reg .b32 a1, a2, a3, b1, b2, b3;\n\t"
        ".reg .b64 db0, db1, db2, db3;\n\t"
        "setp.ne.b32 pacc, %5, 0;\n\t"
        "setp.eq.b32 ptrue, 0, 0;\n\t"
        "setp.ne.b32 q0, %6, 0;\n\t"
        "setp.gt.and.u32 q1, %7, 1, q0;\n\t"
        "setp.gt.and.u32 q2, %7, 2, q0;\n\t"
        "setp.gt.and.u32 q3, %7, 3, q0;\n\t"
        "setp.ne.and.b32 cx, %9, 0, q0;\n\t"
        "setp.ne.and.b32 cy, %10, 0, q0;\n\t"
        "setp.ne.and.b32 cz, %11, 0, q0;\n\t"
        "add.u32 a1, %1, 8;\n\t"
        "add.u32 a2, %1, 32;\n\t"
        "add.u32 a3, %1, 40;\n\t"
        "add.u32 b1, %2, 2;\n\t"
        "add.u32 b2, %2, 4;\n\t"
        "add.u32 b3, %2, 6;\n\t"
        "mov.b64 db0, {%2, %3};\n\t"
        "mov.b64 db1, {b1, %3};\n\t"
        "mov.b64 db2, {b2, %3};\n\t"
        "mov.b64 db3, {b3, %3};\n\t"
        "bar.sync %12, 64;\n\t"
        "tcgen05.fence::after_thread_sync;\n\t"
        "@q0 tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db0, %4, pacc;\n\t"
        "@q1 tcgen05.mma.cta_group::1.kind::f16 [%0], [a1], db1, %4, ptrue;\n\t"
        "@q2 tcgen05.mma.cta_group::1.kind::f16 [%0], [a2], db2, %4, ptrue;\n\t"
        "@q3 tcgen05.mma.cta_group::1.kind::f16 [%0], [a3], db3, %4, ptrue;\n\t"
        "@q0 tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%8];\n\t"
        "@cx tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%9];\n\t"
        "@cy tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%10];\n\t"
        "@cz tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%11];\n\t"
        "tcgen05.fence::before_thread_sync;\n\t"
        "bar.arrive %13, 64;\n\t"
        "}" ::"r"(d_tmem),
        "r"(a), "r"(b_lo), "r"(desc_hi), "r"(idesc), "r"(accumulate), "r"(issue), "r"(ksteps), "r"(bar_empty), "r"(bar_x),
        "r"(bar_y), "r"(bar_z), "n"(SYNC_ID), "n"(ARRIVE_ID)
        : "memory");
  } else {
    asm volatile(
        "{\n\t"
        ".reg .pred pacc, ptrue, q0, q1, q2, q3, cx, cy, cz;\n\t"
        ".reg .b32 a1, a2, a3, b1, b2, b3;\n\t"
        ".reg .b64 da0, da1, da2, da3, db0, db1, db2, db3;\n\t"
        "setp.ne.b32 pacc, %5, 0;\n\t"
        "setp.eq.b32 ptrue, 0, 0;\n\t"
        "setp.ne.b32 q0, %6, 0;\n\t"
        "setp.gt.and.u32 q1, %7, 1, q0;\n\t"
        "setp.gt.and.u32 q2, %7, 2, q0;\n\t"
        "setp.gt.and.u32 q3, %7, 3, q0;\n\t"
        "setp.ne.and.b32 cx, %9, 0, q0;\n\t"
        "setp.ne.and.b32 cy, %10, 0, q0;\n\t"
        "setp.ne.and.b32 cz, %11, 0, q0;\n\t"
        "add.u32 a1, %1, 2;\n\t"
        "add.u32 a2, %1, 4;\n\t"
        "add.u32 a3, %1, 6;\n\t"
        "add.u32 b1, %2, 2;\n\t"
        "add.u32 b2, %2, 4;\n\t"
        "add.u32 b3, %2, 6;\n\t"
        "mov.b64 da0, {%1, %3};\n\t"
        "mov.b64 da1, {a1, %3};\n\t"
        "mov.b64 da2, {a2, %3};\n\t"
        "mov.b64 da3, {a3, %3};\n\t"
        "mov.b64 db0, {%2, %3};\n\t"
        "mov.b64 db1, {b1, %3};\n\t"
        "mov.b64 db2, {b2, %3};\n\t"
        "mov.b64 db3, {b3, %3};\n\t"
        "bar.sync %12, 64;\n\t"
        "tcgen05.fence::after_thread_sync;\n\t"
        "@q0 tcgen05.mma.cta_group::1.kind::f16 [%0], da0, db0, %4, pacc;\n\t"
        "@q1 tcgen05.mma.cta_group::1.kind::f16 [%0], da1, db1, %4, ptrue;\n\t"
        "@q2 tcgen05.mma.cta_group::1.kind::f16 [%0], da2, db2, %4, ptrue;\n\t"
        "@q3 tcgen05.mma.cta_group::1.kind::f16 [%0], da3, db3, %4, ptrue;\n\t"
        "@q0 tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%8];\n\t"
        "@cx tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%9];\n\t"
        "@cy tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%10];\n\t"
        "@cz tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%11];\n\t"
        "tcgen05.fence::before_thread_sync;\n\t"
        "bar.arrive %13, 64;\n\t"
        "}" ::"r"(d_tmem),
        "r"(a), "r"(b_lo), "r"(desc_hi), "r"(idesc), "r"(accumulate), "r"(issue), "r"(ksteps), "r"(bar_empty), "r"(bar_x),
        "r"(bar_y), "r"(bar_z), "n"(SYNC_ID), "n"(ARRIVE_ID)
        : "memory");
  }
}
// ---------------------------------------------------------------- CTA pair (cta_group::2)
// Two CTAs of a cluster (the two SMs of a TPC) execute ONE MMA: M = 256 (128 rows per CTA, A from each CTA's own tensor or
// shared memory), B split between the CTAs' shared memories (N / 2 rows each, same offset), D in each CTA's tensor memory at
// the same address.  Only the leader (cluster rank 0) issues; completion is multicast to the same barrier offset in both CTAs.
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of the same shared-memory location in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t map_to_cta(uint32_t smem_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// no ordering of this thread's own writes implied (a pure signal)
__device__ __forceinline__ void mbar_arrive_cluster_relaxed(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void mbar_wait_addr_cluster(uint32_t bar_smem_addr, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%0], %1, 0x989680;\n\t"
      "@p bra WAIT_DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "WAIT_DONE:\n\t"
      "}" ::"r"(bar_smem_addr), "r"(parity)
      : "memory");
}
// one warp of EACH CTA of the pair, same warp index, same shared-memory offset for the result
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* dst_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
// umma4_step for the leader of a CTA pair: cta_group::2 MMAs, commits multicast to both CTAs (mask 0b11)
template <int SYNC_ID, int ARRIVE_ID, bool TS>
__device__ __forceinline__ void umma4_step_pair(uint32_t issue, uint32_t d_tmem, uint32_t a, uint32_t b_lo, uint32_t desc_hi,
                                           uint32_t idesc, uint32_t accumulate, uint32_t ksteps, uint32_t bar_empty,
                                           uint32_t bar_x, uint32_t bar_y, uint32_t bar_z) {
  if constexpr (TS) {
    asm volatile(
        "{\n\t"
        ".reg .pred pacc, ptrue, q0, q1, q2, q3, cx, cy, cz;\n\t"
        ".reg .b16 both;\n\t"
        "mov.b16 both, 3;\n\t"
        ".reg .b32 a1, a2, a3, b1, b2, b3;\n\t"
        ".reg .b64 db0, db1, db2, db3;\n\t"
        "setp.ne.b32 pacc, %5, 0;\n\t"
        "setp.eq.b32 ptrue, 0, 0;\n\t"
        "setp.ne.b32 q0, %6, 0;\n\t"
        "setp.gt.and.u32 q1, %7, 1, q0;\n\t"
        "setp.gt.and.u32 q2, %7, 2, q0;\n\t"
        "setp.gt.and.u32 q3, %7, 3, q0;\n\t"
        "setp.ne.and.b32 cx, %9, 0, q0;\n\t"
        "setp.ne.and.b32 cy, %10, 0, q0;\n\t"
        "setp.ne.and.b32 cz, %11, 0, q0;\n\t"
        "add.u32 a1, %1, 8;\n\t"
        "add.u32 a2, %1, 32;\n\t"
        "add.u32 a3, %1, 40;\n\t"
        "add.u32 b1, %2, 2;\n\t"
        "add.u32 b2, %2, 4;\n\t"
        "add.u32 b3, %2, 6;\n\t"
        "mov.b64 db0, {%2, %3};\n\t"
        "mov.b64 db1, {b1, %3};\n\t"
        "mov.b64 db2, {b2, %3};\n\t"
        "mov.b64 db3, {b3, %3};\n\t"
        "bar.sync %12, 64;\n\t"
        "tcgen05.fence::after_thread_sync;\n\t"
        "@q0 tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], db0, %4, pacc;\n\t"
        "@q1 tcgen05.mma.cta_group::2.kind::f16 [%0], [a1], db1, %4, ptrue;\n\t"
        "@q2 tcgen05.mma.cta_group::2.kind::f16 [%0], [a2], db2, %4, ptrue;\n\t"
        "@q3 tcgen05.mma.cta_group::2.kind::f16 [%0], [a3], db3, %4, ptrue;\n\t"
        "@q0 tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%8], both;\n\t"
        "@cx tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%9], both;\n\t"
        "@cy tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%10], both;\n\t"
        "@cz tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%11], both;\n\t"
        "tcgen05.fence::before_thread_sync;\n\t"
        "bar.arrive %13, 64;\n\t"
        "}" ::"r"(d_tmem),
        "r"(a), "r"(b_lo), "r"(desc_hi), "r"(idesc), "r"(accumulate), "r"(issue), "r"(ksteps), "r"(bar_empty), "r"(bar_x),
        "r"(bar_y), "r"(bar_z), "n"(SYNC_ID), "n"(ARRIVE_ID)
        : "memory");
  } else {
    asm volatile(
        "{\n\t"
        ".reg .pred pacc, ptrue, q0, q1, q2, q3, cx, cy, cz;\n\t"
        ".reg .b16 both;\n\t"
        "mov.b16 both, 3;\n\t"
        ".reg .b32 a1, a2, a3, b1, b2, b3;\n\t"
        ".reg .b64 da0, da1, da2, da3, db0, db1, db2, db3;\n\t"
        "setp.ne.b32 pacc, %5, 0;\n\t"
        "setp.eq.b32 ptrue, 0, 0;\n\t"
        "setp.ne.b32 q0, %6, 0;\n\t"
        "setp.gt.and.u32 q1, %7, 1, q0;\n\t"
        "setp.gt.and.u32 q2, %7, 2, q0;\n\t"
        "setp.gt.and.u32 q3, %7, 3, q0;\n\t"
        "setp.ne.and.b32 cx, %9, 0, q0;\n\t"
        "setp.ne.and.b32 cy, %10, 0, q0;\n\t"
        "setp.ne.and.b32 cz, %11, 0, q0;\n\t"
        "add.u32 a1, %1, 2;\n\t"
        "add.u32 a2, %1, 4;\n\t"
        "add.u32 a3, %1, 6;\n\t"
        "add.u32 b1, %2, 2;\n\t"
        "add.u32 b2, %2, 4;\n\t"
        "add.u32 b3, %2, 6;\n\t"
        "mov.b64 da0, {%1, %3};\n\t"
        "mov.b64 da1, {a1, %3};\n\t"
        "mov.b64 da2, {a2, %3};\n\t"
        "mov.b64 da3, {a3, %3};\n\t"
        "mov.b64 db0, {%2, %3};\n\t"
        "mov.b64 db1, {b1, %3};\n\t"
        "mov.b64 db2, {b2, %3};\n\t"
        "mov.b64 db3, {b3, %3};\n\t"
        "bar.sync %12, 64;\n\t"
        "tcgen05.fence::after_thread_sync;\n\t"
        "@q0 tcgen05.mma.cta_group::2.kind::f16 [%0], da0, db0, %4, pacc;\n\t"
        "@q1 tcgen05.mma.cta_group::2.kind::f16 [%0], da1, db1, %4, ptrue;\n\t"
        "@q2 tcgen05.mma.cta_group::2.kind::f16 [%0], da2, db2, %4, ptrue;\n\t"
        "@q3 tcgen05.mma.cta_group::2.kind::f16 [%0], da3, db3, %4, ptrue;\n\t"
        "@q0 tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%8], both;\n\t"
        "@cx tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%9], both;\n\t"
        "@cy tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%10], both;\n\t"
        "@cz tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%11], both;\n\t"
        "tcgen05.fence::before_thread_sync;\n\t"
        "bar.arrive %13, 64;\n\t"
        "}" ::"r"(d_tmem),
        "r"(a), "r"(b_lo), "r"(desc_hi), "r"(idesc), "r"(accumulate), "r"(issue), "r"(ksteps), "r"(bar_empty), "r"(bar_x),
        "r"(bar_y), "r"(bar_z), "n"(SYNC_ID), "n"(ARRIVE_ID)
        : "memory");
  }
}
// thread i of the warp writes 16 consecutive 32-bit columns of lane (lane base + i)
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]),
      "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void mbar_wait_addr(uint32_t bar_smem_addr, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, 0x989680;\n\t"
      "@p bra WAIT_DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "WAIT_DONE:\n\t"
      "}" ::"r"(bar_smem_addr), "r"(parity)
      : "memory");
}
__device__ __forceinline__ void umma_commit_if(uint32_t issue, uint64_t* bar) {
  asm volatile(
      "{\n\t"
      ".reg .pred q;\n\t"
      "setp.ne.b32 q, %1, 0;\n\t"
      "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t"
      "}" ::"r"(smem_u32(bar)), "r"(issue)
      : "memory");
}
// 32 lanes x 32 columns of fp32: thread i of the warp receives row (lane base + i), columns [c, c+32)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
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
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait(uint32_t (&v)[16]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]), "+r"(v[8]),
                 "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15])
               :
               : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// wait + make every later use of v[] depend on the wait (the load's outputs are only valid after it)
__device__ __forceinline__ void tmem_ld_wait(uint32_t (&v)[32]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]), "+r"(v[8]),
                 "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15])
               :
               : "memory");
  asm volatile(""
               : "+r"(v[16]), "+r"(v[17]), "+r"(v[18]), "+r"(v[19]), "+r"(v[20]), "+r"(v[21]), "+r"(v[22]), "+r"(v[23]),
                 "+r"(v[24]), "+r"(v[25]), "+r"(v[26]), "+r"(v[27]), "+r"(v[28]), "+r"(v[29]), "+r"(v[30]), "+r"(v[31])
               :
               : "memory");
}

// ---------------------------------------------------------------- descriptors
// K-major operand tile, 128-byte swizzle: rows of 64 bf16 (128 B), 8-row atoms of 1024 B (SBO), the
// 16-byte unit index inside a row XOR-ed with (row & 7).  Tile base must be 1024-byte aligned; a K
// step of 16 elements advances the start address by 32 bytes inside the atom.  The 64-bit shared-memory descriptor is
// (start address >> 4) | SBO (1024 >> 4) << 32 | version 1 << 46 | SWIZZLE_128B (2) << 61; the kernels pass its low word
// and attach the constant high word inside the issue statements above.
//
// kind::f16 instruction descriptor: fp32 accumulate, bf16 A and B, both K-major, M x N tile
__host__ __device__ constexpr uint32_t make_idesc_bf16(uint32_t m, uint32_t n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((n >> 3) << 17) | ((m >> 4) << 24);
}

__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}

// max(x, 0) fused into the fp32 -> bf16x2 conversion (F2FP.RELU)
__device__ __forceinline__ uint32_t pack_bf16_relu(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
// packed fp32 pair arithmetic (FADD2 / FFMA2 on sm_100): (x0, x1) += (b0, b1);  (a0, a1) += (x0, x1) * (w0, w1)
__device__ __forceinline__ void fadd2(float& x0, float& x1, float b0, float b1) {
  asm("{\n\t.reg .b64 a, b;\n\tmov.b64 a, {%0, %1};\n\tmov.b64 b, {%2, %3};\n\tadd.rn.f32x2 a, a, b;\n\tmov.b64 {%0, %1}, a;\n\t}"
      : "+f"(x0), "+f"(x1) : "f"(b0), "f"(b1));
}
__device__ __forceinline__ void ffma2(float& a0, float& a1, float x0, float x1, float w0, float w1) {
  asm("{\n\t.reg .b64 a, x, w;\n\tmov.b64 a, {%0, %1};\n\tmov.b64 x, {%2, %3};\n\tmov.b64 w, {%4, %5};\n\tfma.rn.f32x2 a, x, w, a;\n\t"
      "mov.b64 {%0, %1}, a;\n\t}"
      : "+f"(a0), "+f"(a1) : "f"(x0), "f"(x1), "f"(w0), "f"(w1));
}

// byte offset of 16-byte unit `unit` (0..7) of row `row` inside a SW128 K-block
__device__ __forceinline__ uint32_t sw128_offset(uint32_t row, uint32_t unit) {
  return row * 128u + ((unit ^ (row & 7u)) << 4);
}

}  // namespace ptx
}  // namespace srf
