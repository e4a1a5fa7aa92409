// IP-Adapter decoupled cross-attention (K3) at d = 40 on tcgen05 / TMEM (round 2).
//
// Reference: diffusers IPAdapterAttnProcessor2_0 installed at src/models/unet_motion_cross_frame_attn.py:1264-1279 --
// softmax(q K_txt^T) V_txt + ip_scale * softmax(q K_ip^T) V_ip over 77 text + 4 image tokens, tokens built at :1346-1355.
//
// The streaming mma.sync kernel (ip_xattn_stream.cuh) reads Q and writes O once but runs at 0.25 of its HBM roofline: it
// is instruction-bound (739 warp-instructions per 16-row m-tile and head: fragment loads, two MMA chains and a
// two-segment softmax spread over quads).  There is a single KV tile per query tile here (81 keys), so a tensor-core
// formulation needs no online softmax at all and ~3x fewer instructions per row:
//
//   * a CTA keeps K and V of ONE (video, head) in shared memory for its whole life: rows 0..79 = text tokens (77 valid,
//     TMA zero-fills the rest), rows 80..95 = image tokens (4 valid).  Two spare head-dim columns of V carry ones --
//     column 40 on the text rows, column 41 on the image rows -- so the PV MMAs also deliver the two softmax denominators.
//   * three query tiles (128 rows of one frame) are in flight per CTA, each with its own softmax warpgroup (one thread per
//     row), MMA-issuing warp, Q buffer and TMEM columns [S 96 | O_text 48]:
//       QK:  S = Q K^T            (M128 N96 K48)                               -> "S full"
//       softmax warpgroup: m_txt / m_img per row, P = 2^((s - m) c) unnormalised, bf16, written over S[0..48)  -> "P full"
//       PV:  O_text = P[:, 0..80) V[0..80)   (5 k-steps)   into its own 48 columns
//            O_img  = P[:, 80..96) V[80..96) (1 k-step)    into S[48..96) (dead by then)                 -> "O full"
//       epilogue: o = O_text / O_text[40] + ip_scale * O_img / O_img[41]  -> bf16 -> global;  "S free" once both are in registers
//   * the MMA latencies of one tile (QK after "S free", PV after "P full") hide behind the other two tiles' softmax.
//
// Only the pipeline's token counts (77 + 4) and d = 40 take this kernel; everything else stays on the streaming kernel.
#pragma once
#include <cuda.h>
#include "ptx_sm100.cuh"
#include "dense_attn_sm100.cuh"   // I2V_TRACE_* (developer timeline)

namespace i2v {

struct IpTcParams {
  const __nv_bfloat16* q;       // Q  [batch, sq, heads, d], element strides q_sb / q_ss / q_sh (16-byte aligned rows of d)
  long long q_sb, q_ss, q_sh;
  CUtensorMap tm_kt, tm_vt;     // text K / V [batch/g, n_txt, ..]:  box (64, 1, 80, 1)
  CUtensorMap tm_ki, tm_vi;     // image K / V [batch/g, n_ip, ..]:  box (64, 1, 16, 1)
  __nv_bfloat16* o;             // O  [batch, sq, heads, d]
  long long o_sb, o_ss, o_sh;
  int batch, sq, heads, kv_group, q_tiles;
  float scale_log2e, ip_scale;
  unsigned long long* trace;   // developer timeline (-DI2V_TRACE builds), else null
  int trace_cta;
};

constexpr int kIpTcNT = 3;                          // query tiles in flight per CTA (TMEM slots, softmax warpgroups)
constexpr int kIpTcNQ = 6;                          // Q tiles in the TMA ring (a multiple of NT: stage i % NQ always feeds
//                                                     slot i % NT).  Measured: 3 -> 71.9 us, 6 -> 76.1 us, 9 -> 78.8 us at C2
//                                                     level 0: the kernel is bound by the MMA round trips of the three
//                                                     slots, not by the Q loads, and a deeper ring only adds traffic bursts
constexpr int kIpTcThreads = (4 * kIpTcNT + 1 + kIpTcNT) * 32;
constexpr int kIpTcTxtRows = 80, kIpTcKeys = 96, kIpTcD = 40, kIpTcDK = 48;
constexpr int kIpTcEmu = 3;                         // of every 8 text column pairs, how many take the FMA-pipe exp2
constexpr int kIpTcOnesTxt = 40, kIpTcOnesImg = 41;  // spare head-dim columns of V that carry the ones
constexpr int kIpTcQBytes = 128 * 128;
constexpr int kIpTcKVBytes = kIpTcKeys * 128;
constexpr int kIpTcSlotCols = kIpTcKeys + kIpTcDK;   // S 96 (P over [0, 48), O_img over [48, 96)) | O_text 48
constexpr int kIpTcOBytes = 128 * kIpTcD * 2;        // output staging tile of one slot: 128 rows x 80 B, row-major
constexpr int kIpTcSmemBytes = kIpTcNQ * kIpTcQBytes + 2 * kIpTcKVBytes + kIpTcNT * kIpTcOBytes + 512 + 1024;
static_assert(kIpTcNQ % kIpTcNT == 0 && kIpTcSmemBytes <= 227 * 1024, "Q ring");
static_assert(kIpTcNT * kIpTcSlotCols <= 512, "TMEM budget");

template <int NTXT, int NIP>
__global__ void __launch_bounds__(kIpTcThreads, 1) ip_xattn_tc_kernel(const __grid_constant__ IpTcParams P) {
  static_assert(NTXT <= kIpTcTxtRows && NIP <= 16 && NTXT % 2 == 1, "token counts (the odd text count keeps the pad test simple)");
  constexpr int NT = kIpTcNT, NQ = kIpTcNQ;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sm_q = smem;                                   // [NQ][128 rows][128 B]
  uint8_t* sm_k = sm_q + NQ * kIpTcQBytes;                // [96 rows][128 B]   rows 0..79 text, 80..95 image
  uint8_t* sm_v = sm_k + kIpTcKVBytes;
  uint8_t* sm_o = sm_v + kIpTcKVBytes;                    // [NT][128 rows][80 B]   output staging (each warp its own 32 rows)
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm_o + NT * kIpTcOBytes);
  uint64_t* bar_kv_full = bars;             // [1]   K and V landed (TMA bytes)
  uint64_t* bar_kv_ready = bars + 1;        // [1]   ... and the ones columns of V are written
  uint64_t* bar_q_full = bars + 2;          // [NQ]  TMA -> MMA
  uint64_t* bar_q_empty = bar_q_full + NQ;  // [NQ]  MMA -> TMA (QK retired)
  uint64_t* bar_s_full = bar_q_empty + NQ;  // [NT]  MMA -> softmax
  uint64_t* bar_p_full = bar_s_full + NT;   // [NT]  softmax -> MMA (128 arrivals)
  uint64_t* bar_o_full = bar_p_full + NT;   // [NT]  MMA -> softmax (both PVs retired)
  uint64_t* bar_s_free = bar_o_full + NT;   // [NT]  softmax -> MMA: O_text / O_img are in registers (128 arrivals)
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(bar_s_free + NT);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int kTmaWarp = 4 * NT, kMmaWarp0 = 4 * NT + 1;

  // Static assignment: group = (video, head); every CTA serves one group (several, one after the other, only when there
  // are more groups than CTAs) and takes every `cpg`-th item (frame of the video, query tile) of it.
  const int videos = P.batch / P.kv_group;
  const int n_groups = videos * P.heads;
  const int items_per_group = P.kv_group * P.q_tiles;
  const int cpg = max(1, (int)gridDim.x / n_groups);              // CTAs per group
  const int first_group = (int)blockIdx.x % n_groups;
  const int rank_in_group = (int)blockIdx.x / n_groups;
  const int group_step = n_groups <= (int)gridDim.x ? n_groups * 1000000 : (int)gridDim.x;   // "one group only" / stride
  const bool active = n_groups > (int)gridDim.x || rank_in_group < cpg;
  // this launcher only starts grids with n_groups <= gridDim.x (one group per CTA): see capi.cu
  const int bkv = first_group / P.heads, h = first_group % P.heads;
  (void)group_step;

  if (threadIdx.x == 0) {
    mbar_init(bar_kv_full, 1);
    mbar_init(bar_kv_ready, 1);
    for (int q = 0; q < NQ; ++q) {
      mbar_init(bar_q_full + q, 32);   // the 32 lanes of the producer warp (cp.async.mbarrier.arrive.noinc)
      mbar_init(bar_q_empty + q, 1);
    }
    for (int t = 0; t < NT; ++t) {
      mbar_init(bar_s_full + t, 1);
      mbar_init(bar_p_full + t, 128);
      mbar_init(bar_o_full + t, 1);
      mbar_init(bar_s_free + t, 128);
    }
    mbar_fence_init();
  }
  if (warp == kMmaWarp0) tmem_alloc<512>(tmem_base_slot);
  if (warp == kTmaWarp && lane == 0) {
    tma_prefetch_desc(&P.tm_kt);
    tma_prefetch_desc(&P.tm_vt);
    tma_prefetch_desc(&P.tm_ki);
    tma_prefetch_desc(&P.tm_vi);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_slot;

  if (active) {
    if (warp == kTmaWarp) {
      // =========================== TMA producer ===========================
      if (lane == 0) {
        mbar_arrive_expect_tx(bar_kv_full, 2 * kIpTcKVBytes);
        tma_load_4d(sm_k, &P.tm_kt, bar_kv_full, 0, h, 0, bkv, kEvictLast);
        tma_load_4d(sm_k + kIpTcTxtRows * 128, &P.tm_ki, bar_kv_full, 0, h, 0, bkv, kEvictLast);
        tma_load_4d(sm_v, &P.tm_vt, bar_kv_full, 0, h, 0, bkv, kEvictLast);
        tma_load_4d(sm_v + kIpTcTxtRows * 128, &P.tm_vi, bar_kv_full, 0, h, 0, bkv, kEvictLast);
      }
      mbar_wait(bar_kv_full, 0);
      // ones columns of V: element (key r, column c) of the 128-byte-swizzled tile lives at
      // r * 128 + (((c * 2) >> 4) ^ (r & 7)) * 16 + (c * 2) % 16
      for (int r = lane; r < kIpTcKeys; r += 32) {
        const bool txt = r < NTXT, img = r >= kIpTcTxtRows && r < kIpTcTxtRows + NIP;
        if (txt || img) {
          const int c = txt ? kIpTcOnesTxt : kIpTcOnesImg;
          *reinterpret_cast<uint16_t*>(sm_v + r * 128 + ((((c * 2) >> 4) ^ (r & 7)) << 4) + ((c * 2) & 15)) = 0x3F80;   // bf16 1.0
        }
      }
      fence_proxy_async_smem();
      __syncwarp();
      I2V_TRACE_DECL
      I2V_TRACE_INIT(12)
      if (lane == 0) mbar_arrive(bar_kv_ready);
      // Q tiles through the LSU (cp.async, 16 bytes per lane: 6.4 rows of 80 B per warp instruction) into the 128-byte
      // swizzled layout; head-dim columns 40..47 (chunk 5) of every stage are zeroed once and never written again.
      for (int r = lane; r < NQ * 128; r += 32)
        *reinterpret_cast<uint4*>(sm_q + r * 128 + ((5 ^ (r & 7)) << 4)) = make_uint4(0u, 0u, 0u, 0u);
      fence_proxy_async_smem();
      // per-lane copy plan of a full tile (tile-independent): granule g = k * 32 + lane -> row g / 5, 16-byte chunk g % 5
      constexpr int kQCopies = 128 * 5 / 32;
      uint32_t src_off[kQCopies], dst_off[kQCopies];
#pragma unroll
      for (int k = 0; k < kQCopies; ++k) {
        const int g = k * 32 + lane, r = g / 5, c = g - r * 5;
        src_off[k] = (uint32_t)(r * (int)P.q_ss + c * 8) * 2u;            // bytes from the tile's first row
        dst_off[k] = (uint32_t)(r * 128 + ((c ^ (r & 7)) << 4));
      }
      int i = 0, f = rank_in_group / P.q_tiles, qt = rank_in_group - f * P.q_tiles;
      const int df = cpg / P.q_tiles, dqt = cpg - df * P.q_tiles;
      for (int it = rank_in_group; it < items_per_group; it += cpg, ++i) {
        const int q = i % NQ, nq = i / NQ;
        I2V_TRACE_EV(0x20)
        mbar_wait_parked(bar_q_empty + q, (nq & 1) ^ 1, kParkNs);
        I2V_TRACE_EV(0x21)
        const uint8_t* qtile = reinterpret_cast<const uint8_t*>(
            P.q + (long long)(bkv * P.kv_group + f) * P.q_sb + (long long)h * P.q_sh + (long long)(qt * 128) * P.q_ss);
        const uint32_t stage = smem_u32(sm_q + q * kIpTcQBytes);
        if (qt * 128 + 128 <= P.sq) {
#pragma unroll
          for (int k = 0; k < kQCopies; ++k)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(stage + dst_off[k]), "l"(qtile + src_off[k]) : "memory");
        } else {   // ragged last tile: rows beyond sq read the last valid row (their outputs are not stored)
#pragma unroll
          for (int k = 0; k < kQCopies; ++k) {
            const int g = k * 32 + lane, r = g / 5, c = g - r * 5;
            const int rr = min(r, P.sq - 1 - qt * 128);
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(stage + dst_off[k]),
                         "l"(qtile + ((long long)rr * P.q_ss + c * 8) * 2) : "memory");
          }
        }
        cp_async_mbar_arrive_noinc(bar_q_full + q);
        f += df; qt += dqt;
        if (qt >= P.q_tiles) { qt -= P.q_tiles; ++f; }
      }
    } else if (warp >= kMmaWarp0) {
      // =========================== MMA issuer of slot t ===========================
      const int t = warp - kMmaWarp0;
      constexpr uint32_t idesc_qk = make_idesc_bf16(128, kIpTcKeys, 0, 0);
      constexpr uint32_t idesc_pv = make_idesc_bf16(128, kIpTcDK, 0, 1);
      const uint32_t ka = smem_u32(sm_k) >> 4;
      const uint32_t va = smem_u32(sm_v) >> 4;
      const uint64_t desc_k_major = make_smem_desc_sw128(0, 16, 1024);
      const uint64_t desc_v = make_smem_desc_sw128(0, kIpTcKVBytes, 1024);
      const uint32_t tm_slot = tmem_base + t * kIpTcSlotCols;
      mbar_wait_parked(bar_kv_ready, 0, kParkNs);
      I2V_TRACE_DECL
      I2V_TRACE_INIT(8 + t)
      int n = 0;
      for (int i = t; rank_in_group + (long long)i * cpg < items_per_group; i += NT, ++n) {
        const int q = i % NQ;
        const uint32_t qa = smem_u32(sm_q + q * kIpTcQBytes) >> 4;
        I2V_TRACE_EV(0x10)
        mbar_wait_parked(bar_q_full + q, (i / NQ) & 1, kParkNs);
        I2V_TRACE_EV(0x11)
        mbar_wait_parked(bar_s_free + t, (n & 1) ^ 1, kParkNs);   // the previous tile's O_img has left the S columns
        fence_proxy_async_smem();   // the query tile was written through the generic proxy (cp.async)
        tc_fence_after();
        I2V_TRACE_EV(0x12)
        if (elect_one()) {
#pragma unroll
          for (int kk = 0; kk < kIpTcDK / 16; ++kk) {
            const uint64_t da = desc_k_major | (uint64_t)((qa + kk * 2) & 0x3FFF);
            const uint64_t db = desc_k_major | (uint64_t)((ka + kk * 2) & 0x3FFF);
            umma_ss(tm_slot, da, db, idesc_qk, kk > 0 ? 1u : 0u);
          }
          tc_commit(bar_s_full + t);
          tc_commit(bar_q_empty + q);
        }
        __syncwarp();
        I2V_TRACE_EV(0x13)
        mbar_wait_parked(bar_p_full + t, n & 1, kParkNs);
        tc_fence_after();
        I2V_TRACE_EV(0x14)
        if (elect_one()) {
#pragma unroll
          for (int kk = 0; kk < kIpTcTxtRows / 16; ++kk) {   // text keys 0..79 -> O_text
            const uint64_t db = desc_v | (uint64_t)((va + kk * (2048 >> 4)) & 0x3FFF);
            umma_ts(tm_slot + kIpTcKeys, tm_slot + kk * 8, db, idesc_pv, kk > 0 ? 1u : 0u);
          }
          {                                                   // image keys 80..95 -> O_img over S[48, 96)
            const uint64_t db = desc_v | (uint64_t)((va + (kIpTcTxtRows / 16) * (2048 >> 4)) & 0x3FFF);
            umma_ts(tm_slot + kIpTcDK, tm_slot + (kIpTcTxtRows / 16) * 8, db, idesc_pv, 0u);
          }
          tc_commit(bar_o_full + t);
        }
        __syncwarp();
        I2V_TRACE_EV(0x15)
      }
    } else {
      // =========================== softmax + epilogue warpgroup of slot t ===========================
      const int t = warp >> 2;
      const uint32_t tm_slot = tmem_base + t * kIpTcSlotCols + ((uint32_t)((warp & 3) * 32) << 16);
      const float c = P.scale_log2e;
      I2V_TRACE_DECL
      if ((warp & 3) == 0) { I2V_TRACE_INIT(t) }
      // per-lane plan of the output copy (tile-independent): granule g = k * 32 + lane of the warp's 32 rows
      int st_row[kIpTcD / 8];
      uint32_t st_off[kIpTcD / 8];
#pragma unroll
      for (int k = 0; k < kIpTcD / 8; ++k) {
        const int g = k * 32 + lane, r = g / 5, c = g - r * 5;
        st_row[k] = r;
        st_off[k] = (uint32_t)(r * (int)P.o_ss + c * 8) * 2u;
      }
      int n = 0;
      for (int i = t; rank_in_group + (long long)i * cpg < items_per_group; i += NT, ++n) {
        const int it = rank_in_group + i * cpg;
        const int f = it / P.q_tiles, qt = it - f * P.q_tiles;
        I2V_TRACE_EV(0x1)
        mbar_wait(bar_s_full + t, n & 1);
        tc_fence_after();
        I2V_TRACE_EV(0x2)
        // text scores: columns 0..NTXT-1 (NTXT = 77: 38 pairs + one); image scores: columns 80..80+NIP-1
        uint32_t pk[kIpTcKeys / 2];
        float m1 = -INFINITY, m2 = -INFINITY;
        {
          uint32_t r0[32], r1[32], r2[16], r3[8];
          tmem_ld_x32(tm_slot, r0);
          tmem_ld_x32(tm_slot + 32, r1);
          tmem_ld_x16(tm_slot + 64, r2);
          tmem_ld_x8(tm_slot + kIpTcTxtRows, r3);
          tc_wait_ld();
          float sv[kIpTcTxtRows];
#pragma unroll
          for (int k = 0; k < 32; ++k) { sv[k] = __uint_as_float(r0[k]); sv[32 + k] = __uint_as_float(r1[k]); }
#pragma unroll
          for (int k = 0; k < 16; ++k) sv[64 + k] = __uint_as_float(r2[k]);
          float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
          for (int k = 0; k + 8 <= NTXT; k += 8) {
            mx0 = fmax3(mx0, sv[k + 0], sv[k + 1]);
            mx1 = fmax3(mx1, sv[k + 2], sv[k + 3]);
            mx2 = fmax3(mx2, sv[k + 4], sv[k + 5]);
            mx3 = fmax3(mx3, sv[k + 6], sv[k + 7]);
          }
#pragma unroll
          for (int k = NTXT / 8 * 8; k < NTXT; ++k) mx0 = fmaxf(mx0, sv[k]);
          m1 = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
#pragma unroll
          for (int k = 0; k < NIP; ++k) m2 = fmaxf(m2, __uint_as_float(r3[k]));
          const float mc1 = m1 * c, mc2 = m2 * c;
          // 3 of every 8 column pairs take the FMA-pipe exp2 (the three warpgroups' exponentials run at the same time and
          // are bound by the XU pipe otherwise); x <= 0 here, the polynomial clamps at -126
          const uint64_t c2 = f2_pack(c, c), nmc2 = f2_pack(-mc1, -mc1);
#pragma unroll
          for (int k = 0; k < kIpTcTxtRows / 2; ++k) {
            if (2 * k >= NTXT) { pk[k] = 0u; continue; }
            const uint64_t x2 = f2_fma(f2_pack(sv[2 * k], sv[2 * k + 1]), c2, nmc2);
            float p0, p1;
            if ((k & 7) < kIpTcEmu && 2 * k + 1 < NTXT) {
              f2_unpack(ex2_emu_pair_x<3, true>(x2), p0, p1);
            } else {
              float x0, x1;
              f2_unpack(x2, x0, x1);
              p0 = ex2_approx(x0);
              p1 = 2 * k + 1 < NTXT ? ex2_approx(x1) : 0.f;
            }
            pk[k] = pack_bf16x2(p0, p1);
          }
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const float p0 = 2 * k < NIP ? ex2_approx(fmaf(__uint_as_float(r3[(2 * k) & 7]), c, -mc2)) : 0.f;
            const float p1 = 2 * k + 1 < NIP ? ex2_approx(fmaf(__uint_as_float(r3[(2 * k + 1) & 7]), c, -mc2)) : 0.f;
            pk[kIpTcTxtRows / 2 + k] = (2 * k < NIP) ? pack_bf16x2(p0, p1) : 0u;
          }
        }
        I2V_TRACE_EV(0x3)
        {
          uint32_t a[32], b[16];
#pragma unroll
          for (int k = 0; k < 32; ++k) a[k] = pk[k];
#pragma unroll
          for (int k = 0; k < 16; ++k) b[k] = pk[32 + k];
          tmem_st_x32(tm_slot, a);
          tmem_st_x16(tm_slot + 32, b);
        }
        tc_wait_st();
        tc_fence_before();
        mbar_arrive(bar_p_full + t);
        I2V_TRACE_EV(0x4)

        // ---- epilogue ----
        mbar_wait(bar_o_full + t, n & 1);
        tc_fence_after();
        I2V_TRACE_EV(0x5)
        uint32_t ot0[32], ot1[16], oi0[32], oi1[16];
        tmem_ld_x32(tm_slot + kIpTcKeys, ot0);
        tmem_ld_x16(tm_slot + kIpTcKeys + 32, ot1);
        tmem_ld_x32(tm_slot + kIpTcDK, oi0);
        tmem_ld_x16(tm_slot + kIpTcDK + 32, oi1);
        tc_wait_ld();
        tc_fence_before();
        mbar_arrive(bar_s_free + t);
        I2V_TRACE_EV(0x6)
        const float w1 = rcp_approx(__uint_as_float(ot1[kIpTcOnesTxt - 32]));
        const float w2 = P.ip_scale * rcp_approx(__uint_as_float(oi1[kIpTcOnesImg - 32]));
        // o = O_text * w1 + O_img * w2 -> bf16 -> this row of the staging tile (row pitch 80 B: the five 16-byte stores of
        // eight consecutive lanes cover all 32 banks); then the warp writes its 32 rows out 16 bytes per lane in memory
        // order (6.4 rows per store instruction instead of one 16-byte piece of 32 different rows)
        uint8_t* wstage = sm_o + t * kIpTcOBytes + (warp & 3) * 32 * (kIpTcD * 2);
        uint8_t* srow = wstage + lane * (kIpTcD * 2);
#pragma unroll
        for (int v8 = 0; v8 < kIpTcD / 8; ++v8) {
          uint32_t w[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int c0 = v8 * 8 + 2 * k;
            const float a0 = __uint_as_float(c0 < 32 ? ot0[c0 & 31] : ot1[(c0 - 32) & 15]);
            const float a1 = __uint_as_float(c0 + 1 < 32 ? ot0[(c0 + 1) & 31] : ot1[(c0 + 1 - 32) & 15]);
            const float b0 = __uint_as_float(c0 < 32 ? oi0[c0 & 31] : oi1[(c0 - 32) & 15]);
            const float b1 = __uint_as_float(c0 + 1 < 32 ? oi0[(c0 + 1) & 31] : oi1[(c0 + 1 - 32) & 15]);
            w[k] = pack_bf16x2(fmaf(b0, w2, a0 * w1), fmaf(b1, w2, a1 * w1));
          }
          *reinterpret_cast<uint4*>(srow + v8 * 16) = make_uint4(w[0], w[1], w[2], w[3]);
        }
        __syncwarp();
        const int row0 = qt * 128 + (warp & 3) * 32;
        uint8_t* otile = reinterpret_cast<uint8_t*>(P.o + (long long)(bkv * P.kv_group + f) * P.o_sb + (long long)h * P.o_sh +
                                                    (long long)row0 * P.o_ss);
#pragma unroll
        for (int k = 0; k < kIpTcD / 8; ++k) {
          const uint4 v = *reinterpret_cast<const uint4*>(wstage + (k * 32 + lane) * 16);
          if (row0 + st_row[k] < P.sq) *reinterpret_cast<uint4*>(otile + st_off[k]) = v;
        }
        __syncwarp();   // (the staging rows are rewritten by the next tile of this slot)
        I2V_TRACE_EV(0x7)
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp0) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

}  // namespace i2v
