// Thin inline-PTX wrappers for the sm_100a features the attention kernels use:
// mbarrier, TMA (cp.async.bulk.tensor), tcgen05 (alloc / mma / ld / st / commit / fences).
// Everything here is a direct statement of the PTX ISA; there is no library dependency.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

namespace i2v {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t"
      ".reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}

// ----------------------------------------------------------------------------------------------
// mbarrier
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t tx_bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(tx_bytes)
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred P;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug turns into a trap (reported as a CUDA error by the C-ABI) instead of
// hanging the GPU box.  The bound is ~2 s of SM clock, far beyond any legitimate wait in these kernels.
#ifndef I2V_MBAR_TIMEOUT_CYCLES
#define I2V_MBAR_TIMEOUT_CYCLES 4000000000ll
#endif
// try_wait with a suspend-time hint: the warp sleeps in hardware until the phase completes or the hint (ns) expires,
// instead of burning issue slots of its SM sub-partition in a polling loop (38 % of the executed instructions of the
// first pipelined build were such polls; the softmax warps share those issue slots).
__device__ __forceinline__ bool mbar_try_wait_suspend(uint64_t* bar, uint32_t parity, uint32_t hint_ns) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred P;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity), "r"(hint_ns)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  // plain polling: measured faster than a long suspend-time hint (the wake-up from a hardware suspend costs more
  // latency on the softmax <-> MMA hand-offs than the polls cost in issue slots)
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if ((++spins & 0x3FFu) == 0 && (clock64() - t0) > I2V_MBAR_TIMEOUT_CYCLES) { __trap(); }
  }
}

// Wait that parks the warp in hardware (try_wait with a suspend-time hint) instead of polling: for the producer-side
// warps (TMA, MMA issue) whose polls would otherwise take issue slots from the softmax warps of their sub-partition.
__device__ __forceinline__ void mbar_wait_parked(uint64_t* bar, uint32_t parity, uint32_t hint_ns = 2000u) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  uint32_t spins = 0;
  while (!mbar_try_wait_suspend(bar, parity, hint_ns)) {
    if ((++spins & 0x3FFu) == 0 && (clock64() - t0) > I2V_MBAR_TIMEOUT_CYCLES) { __trap(); }
  }
}

// ----------------------------------------------------------------------------------------------
// Proxy fences
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// ----------------------------------------------------------------------------------------------
// TMA: tiled tensor loads global -> shared, completion on an mbarrier
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const void* tmap) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(tmap)) : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* smem_dst, const void* tmap, uint64_t* bar, int c0, int c1, int c2,
                                            int c3, uint64_t cache_hint) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
      " [%0], [%1, {%3, %4, %5, %6}], [%2], %7;"
      :
      : "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar)), "r"(c0), "r"(c1),
        "r"(c2), "r"(c3), "l"(cache_hint)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const void* tmap, uint64_t* bar, int c0, int c1,
                                            uint64_t cache_hint) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
      " [%0], [%1, {%3, %4}], [%2], %5;"
      :
      : "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar)), "r"(c0), "r"(c1),
        "l"(cache_hint)
      : "memory");
}
// 16-byte asynchronous copy global -> shared through the LSU (LDGSTS), bypassing L1, and the mbarrier arrival that fires when
// all earlier cp.async of the executing thread have landed (.noinc: the barrier's expected count must already include it).
// For operands that are many short rows: the TMA engine serves ~one row request per 4 clk per SM, the LSU path several times that.
__device__ __forceinline__ void cp_async_16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_mbar_arrive_noinc(uint64_t* bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// Tiled store shared -> global (bulk async-group completion): the issuing thread commits a group and, before the
// shared-memory source is overwritten, waits until the group has been read.
__device__ __forceinline__ void tma_store_2d(const void* tmap, const void* smem_src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(tmap)),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_store_4d(const void* tmap, const void* smem_src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(tmap)),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
// CTA-pair flavour (tcgen05 cta_group::2): executed by both CTAs of the pair, the bytes are counted on the mbarrier of
// the pair's leader (CTA 0): clearing the peer bit of the shared::cluster address selects it.
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;
__device__ __forceinline__ void tma_load_2d_pair(void* smem_dst, const void* tmap, uint64_t* bar, int c0, int c1,
                                                 uint64_t cache_hint) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
      " [%0], [%1, {%3, %4}], [%2], %5;"
      :
      : "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar) & kPeerBitMask), "r"(c0),
        "r"(c1), "l"(cache_hint)
      : "memory");
}
// arrive on the mbarrier at this offset in CTA `cta` of the cluster
__device__ __forceinline__ void mbar_arrive_cluster(uint64_t* bar, uint32_t cta) {
  asm volatile(
      "{\n\t"
      ".reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t"
      "}" ::"r"(smem_u32(bar)),
      "r"(cta)
      : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// L2 eviction-priority policies (createpolicy encodings used by TMA cache hints).
constexpr uint64_t kEvictFirst = 0x12F0000000000000ull;
constexpr uint64_t kEvictNormal = 0x1000000000000000ull;
constexpr uint64_t kEvictLast = 0x14F0000000000000ull;

// ----------------------------------------------------------------------------------------------
// tcgen05: TMEM allocation
// ----------------------------------------------------------------------------------------------
template <uint32_t kCols>
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_result) {
  static_assert(kCols >= 32 && kCols <= 512 && (kCols & (kCols - 1)) == 0, "TMEM columns: power of two in [32,512]");
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)),
               "n"(kCols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
// CTA pair: one warp of each CTA, same shared-memory slot offset in both
template <uint32_t kCols>
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* smem_result) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)),
               "n"(kCols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <uint32_t kCols>
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(kCols) : "memory");
}
template <uint32_t kCols>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(kCols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// tcgen05.commit: the mbarrier gets one arrival when all previously issued MMAs of this thread retire.
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// CTA pair: all MMAs issued so far by this thread for the pair; one arrival on the mbarrier at this offset in every CTA
// of `cta_mask`.
__device__ __forceinline__ void tc_commit_pair(uint64_t* bar, uint16_t cta_mask) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"(cta_mask)
      : "memory");
}

// ----------------------------------------------------------------------------------------------
// tcgen05.mma, kind::f16 (bf16 x bf16 -> fp32), single CTA.
//   D[tmem] (+)= A[smem desc] * B[smem desc]      (SS)
//   D[tmem] (+)= A[tmem]      * B[smem desc]      (TS)
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ void umma_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                        uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n"
      :
      : "r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// CTA pair (issued by the leader CTA only): M = 256 over the two CTAs' 128 rows of A, each CTA's shared memory holds
// its own rows of A and its half of B's N rows at the same offsets; D is split by rows over the two CTAs' TMEM.
__device__ __forceinline__ void umma_ss_pair(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n"
      :
      : "r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                        uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}\n"
      :
      : "r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// Instruction descriptor for kind::f16, bf16 inputs, fp32 accumulate.
//   [4,6) c_format = 1 (F32); [7,10) a_format = 1 (BF16); [10,13) b_format = 1 (BF16);
//   [15] a_major (0 = K-major); [16] b_major (0 = K-major, 1 = MN-major);
//   [17,23) N >> 3; [24,29) M >> 4.
__host__ __device__ constexpr uint32_t make_idesc_bf16(uint32_t M, uint32_t N, uint32_t a_mn_major,
                                                       uint32_t b_mn_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (a_mn_major << 15) | (b_mn_major << 16) | ((N >> 3) << 17) |
         ((M >> 4) << 24);
}

// Shared-memory matrix descriptor (sm_100 "version 1"), 128-byte swizzle.
//   [0,14) start address >> 4; [16,30) leading byte offset >> 4; [32,46) stride byte offset >> 4;
//   [46,48) version = 1; [61,64) layout type (2 = SWIZZLE_128B).
// K-major operand tile (rows x 64 bf16, 128 B per row, 8-row groups 1024 B apart):
//   SBO = 1024 (distance between 8-row groups), LBO unused for swizzled K-major (set to 1).
// MN-major operand tile (K rows x 64 bf16 along MN, 128 B per K row):
//   SBO = 1024 (distance between groups of 8 K rows), LBO = distance between 64-element MN atoms.
__device__ __forceinline__ uint64_t make_smem_desc_sw128(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t desc = 0;
  desc |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
  desc |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
  desc |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32;
  desc |= static_cast<uint64_t>(1) << 46;
  desc |= static_cast<uint64_t>(2) << 61;
  return desc;
}

// ----------------------------------------------------------------------------------------------
// tcgen05.ld / tcgen05.st, shape 32x32b: thread t of the warp owns TMEM lane (32*(warp%4) + t) and
// reads/writes N consecutive 32-bit columns.
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_ld_x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32"
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15,"
      " %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_x16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32"
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_st_x32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0],"
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16,"
      " %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      :
      : "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
        "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]),
        "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]),
        "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_x16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0],"
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      :
      : "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
        "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}

__device__ __forceinline__ void tmem_ld_x8(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_st_x8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               :
               : "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}

// ----------------------------------------------------------------------------------------------
// small math helpers
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}

// ---- packed fp32x2 arithmetic (FFMA2 / FADD2: two lanes per FMA-pipe issue) and 3-input max (FMNMX3) ----
__device__ __forceinline__ uint64_t f2_pack(float lo, float hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void f2_unpack(uint64_t v, float& lo, float& hi) {
  asm("mov.b64 {%0,%1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t f2_fma(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t r;
  asm("fma.rn.f32x2 %0,%1,%2,%3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ uint64_t f2_mul(uint64_t a, uint64_t b) {
  uint64_t r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ uint64_t f2_add(uint64_t a, uint64_t b) {
  uint64_t r;
  asm("add.rn.f32x2 %0,%1,%2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ uint64_t f2_add_rm(uint64_t a, uint64_t b) {
  uint64_t r;
  asm("add.rm.f32x2 %0,%1,%2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ uint64_t f2_sub(uint64_t a, uint64_t b) {
  uint64_t r;
  asm("sub.rn.f32x2 %0,%1,%2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ float fmax3(float a, float b, float c) {
  float r;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
  return r;
}

// 2^x for a pair of x <= 0 on the FMA / ALU pipes (no MUFU): Cody-Waite split x = n + f, f in [0,1), degree-3
// minimax polynomial for 2^f (max relative error 8.6e-5, far below the bf16 rounding of P), exponent patched in
// with an integer add.  x is clamped to >= -126 so the result never goes denormal / wraps.
__device__ __forceinline__ uint64_t ex2_emulated_pair(uint64_t x) {
  float x0, x1;
  f2_unpack(x, x0, x1);
  x = f2_pack(fmaxf(x0, -126.f), fmaxf(x1, -126.f));
  const float kMagic = 12582912.f;  // 1.5 * 2^23: adding it with round-toward--inf leaves floor(x) in the low mantissa bits
  const uint64_t t = f2_add_rm(x, f2_pack(kMagic, kMagic));
  const uint64_t n = f2_add(t, f2_pack(-kMagic, -kMagic));
  const uint64_t f = f2_sub(x, n);
  uint64_t p = f2_fma(f, f2_pack(0.07706582f, 0.07706582f), f2_pack(0.22764632f, 0.22764632f));
  p = f2_fma(p, f, f2_pack(0.69511649f, 0.69511649f));
  p = f2_fma(p, f, f2_pack(1.f, 1.f));
  float p0, p1, t0, t1;
  f2_unpack(p, p0, p1);
  f2_unpack(t, t0, t1);
  const float r0 = __int_as_float(__float_as_int(p0) + (__float_as_int(t0) << 23));
  const float r1 = __int_as_float(__float_as_int(p1) + (__float_as_int(t1) << 23));
  return f2_pack(r0, r1);
}

// Same without the lower clamp: the caller guarantees x >= -126 (or does not care about garbage for such x).
__device__ __forceinline__ uint64_t ex2_emulated_pair_noclamp(uint64_t x) {
  const float kMagic = 12582912.f;
  const uint64_t t = f2_add_rm(x, f2_pack(kMagic, kMagic));
  const uint64_t n = f2_add(t, f2_pack(-kMagic, -kMagic));
  const uint64_t f = f2_sub(x, n);
  uint64_t p = f2_fma(f, f2_pack(0.07706582f, 0.07706582f), f2_pack(0.22764632f, 0.22764632f));
  p = f2_fma(p, f, f2_pack(0.69511649f, 0.69511649f));
  p = f2_fma(p, f, f2_pack(1.f, 1.f));
  float p0, p1, t0, t1;
  f2_unpack(p, p0, p1);
  f2_unpack(t, t0, t1);
  const float r0 = __int_as_float(__float_as_int(p0) + (__float_as_int(t0) << 23));
  const float r1 = __int_as_float(__float_as_int(p1) + (__float_as_int(t1) << 23));
  return f2_pack(r0, r1);
}

__device__ __forceinline__ uint64_t f2_fma_rm(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t r;
  asm("fma.rm.f32x2 %0,%1,%2,%3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}

constexpr float kExpMagic = 12582912.f;  // 1.5 * 2^23

// 2^(s*c - m) for a pair of raw scores s with an INTEGER reference max m (log2 domain), FMA / ALU pipes only.
// k1 = (magic - m, magic - m) is exact because m is an integer, so the Cody-Waite split rides on the scale FMA:
//   t = fma.rm(s, c, k1)      = magic + floor(x)         x = s*c - m   (one rounding, toward -inf: exact floor)
//   f = fma.rn(s, c, k1 - t)  = x - floor(x) in [0, 1]   (k1 - t is an exact small integer)
//   2^f by a minimax polynomial (degree 3: 7.9e-5 / degree 2: 1.7e-3 max relative error; bf16 rounding of P is 2e-3),
//   2^floor(x) patched into the exponent with an integer shift-add; floor(x) is clamped at -126 on the integer view
//   of t (positive floats order like integers) so very negative scores give ~0 instead of wrapping.
// Five packed FMA-pipe instructions per pair (degree 3) + two integer ops per element, against one FFMA2 + two
// MUFU.EX2 (16 clk of the XU pipe) for the plain path.
template <int DEG, bool CLAMP>
__device__ __forceinline__ uint64_t ex2_emu_pair_int(uint64_t s2, uint64_t c2, uint64_t k1) {
  const uint64_t t = f2_fma_rm(s2, c2, k1);
  const uint64_t f = f2_fma(s2, c2, f2_sub(k1, t));
  uint64_t p;
  if (DEG == 3) {
    p = f2_fma(f, f2_pack(0.0780693937f, 0.0780693937f), f2_pack(0.2259353543f, 0.2259353543f));
    p = f2_fma(p, f, f2_pack(0.6959163017f, 0.6959163017f));
    p = f2_fma(p, f, f2_pack(0.9999210496f, 0.9999210496f));
  } else {
    p = f2_fma(f, f2_pack(0.3371894347f, 0.3371894347f), f2_pack(0.6576362757f, 0.6576362757f));
    p = f2_fma(p, f, f2_pack(1.0017247632f, 1.0017247632f));
  }
  float p0, p1, t0, t1;
  f2_unpack(p, p0, p1);
  f2_unpack(t, t0, t1);
  int n0 = __float_as_int(t0), n1 = __float_as_int(t1);
  if (CLAMP) {
    constexpr int kLo = 0x4B400000 - 126;  // bits of magic - 126
    n0 = max(n0, kLo);
    n1 = max(n1, kLo);
  }
  const float r0 = __int_as_float(__float_as_int(p0) + (n0 << 23));
  const float r1 = __int_as_float(__float_as_int(p1) + (n1 << 23));
  return f2_pack(r0, r1);
}

// Same for a pair of exponents x that are already final (x = s*c - m was produced by the MMA itself: the scale is
// folded into the query projection and -m rides in a spare head-dim column, see dense_attn_ring_sm100.cuh):
//   t = add.rm(x, magic);  f = x - (t - magic)
template <int DEG, bool CLAMP>
__device__ __forceinline__ uint64_t ex2_emu_pair_x(uint64_t x2) {
  const uint64_t mg = f2_pack(kExpMagic, kExpMagic);
  const uint64_t t = f2_add_rm(x2, mg);
  const uint64_t f = f2_add(x2, f2_sub(mg, t));
  uint64_t p;
  if (DEG == 3) {
    p = f2_fma(f, f2_pack(0.0780693937f, 0.0780693937f), f2_pack(0.2259353543f, 0.2259353543f));
    p = f2_fma(p, f, f2_pack(0.6959163017f, 0.6959163017f));
    p = f2_fma(p, f, f2_pack(0.9999210496f, 0.9999210496f));
  } else {
    p = f2_fma(f, f2_pack(0.3371894347f, 0.3371894347f), f2_pack(0.6576362757f, 0.6576362757f));
    p = f2_fma(p, f, f2_pack(1.0017247632f, 1.0017247632f));
  }
  float p0, p1, t0, t1;
  f2_unpack(p, p0, p1);
  f2_unpack(t, t0, t1);
  int n0 = __float_as_int(t0), n1 = __float_as_int(t1);
  if (CLAMP) {
    constexpr int kLo = 0x4B400000 - 126;
    n0 = max(n0, kLo);
    n1 = max(n1, kLo);
  }
  const float r0 = __int_as_float(__float_as_int(p0) + (n0 << 23));
  const float r1 = __int_as_float(__float_as_int(p1) + (n1 << 23));
  return f2_pack(r0, r1);
}

// One softmax row step on BN raw scores held in registers: P = 2^(s*c - m) as packed bf16 pairs (pk) + the fp32 row
// sum.  EMU of every 8 column pairs take the FMA-pipe path; m must be an integer when EMU > 0.
// PRESCALED: sv already holds x = s*c - m.  SUM = false: the caller gets the row sum elsewhere (ones column of V).
// PAT = 1: the EMU pairs of every 8 are spread evenly ((i * EMU) % 8 < EMU) instead of taken first.
template <int BN, int EMU, int DEG, bool CLAMP, bool PRESCALED = false, bool SUM = true, int PAT = 0>
__device__ __forceinline__ float softmax_exp_row(const float (&sv)[BN], float c, float m, uint32_t (&pk)[BN / 2]) {
  const uint64_t c2 = f2_pack(c, c);
  const uint64_t nm2 = f2_pack(-m, -m);
  const uint64_t k1 = f2_pack(kExpMagic - m, kExpMagic - m);
  uint64_t ls0 = 0ull, ls1 = 0ull;
#pragma unroll
  for (int i = 0; i < BN / 2; ++i) {
    const uint64_t s2 = f2_pack(sv[2 * i], sv[2 * i + 1]);
    uint64_t p2;
    if (DEG == 0) {
      p2 = s2;   // developer experiment (never dispatched by the library): the hand-off pipeline without exponentials
    } else if (PAT == 0 ? ((i & 7) < EMU) : (((i * EMU) & 7) < EMU)) {
      p2 = PRESCALED ? ex2_emu_pair_x<DEG, CLAMP>(s2) : ex2_emu_pair_int<DEG, CLAMP>(s2, c2, k1);
    } else {
      float x0, x1;
      f2_unpack(PRESCALED ? s2 : f2_fma(s2, c2, nm2), x0, x1);
      p2 = f2_pack(ex2_approx(x0), ex2_approx(x1));
    }
    if (SUM) { if (i & 1) ls1 = f2_add(ls1, p2); else ls0 = f2_add(ls0, p2); }
    float p0, p1;
    f2_unpack(p2, p0, p1);
    pk[i] = pack_bf16x2(p0, p1);
  }
  float a0, a1;
  f2_unpack(f2_add(ls0, ls1), a0, a1);
  return a0 + a1;
}

// Named barrier over `nthreads` threads that also OR-reduces a predicate: one instruction for "synchronise the two
// warps that share a row group and tell both whether either of them needs the slow path".
__device__ __forceinline__ bool named_bar_red_or(uint32_t id, uint32_t nthreads, bool pred) {
  uint32_t r;
  asm volatile(
      "{\n\t"
      ".reg .pred P, Q;\n\t"
      "setp.ne.u32 Q, %3, 0;\n\t"
      "bar.red.or.pred P, %1, %2, Q;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t"
      "}\n"
      : "=r"(r)
      : "r"(id), "r"(nthreads), "r"((uint32_t)pred)
      : "memory");
  return r != 0;
}
__device__ __forceinline__ void named_bar_sync(uint32_t id, uint32_t nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

}  // namespace i2v
