// Software-pipelined dense attention forward for small head dims (d <= 48: SD1.5 level 0, d = 40) on sm_100a.
//
// Same operator and launch contract as dense_attn_sm100.cuh (K1 spatial self-attention + K2 I2V-Adapter cross-frame
// attention of src/modules/i2v_adapter.py:468-492 in one launch), different schedule.  At d = 40 a 128 x BN score
// tile costs 3x more exponential time than MMA time, so the kernel is bound by the softmax warps; what the first
// kernel loses is that a softmax warpgroup idles through a full tensor round trip every KV tile
// (P(j) -> PV(j) -> QK(j+1) -> S(j+1): ~1400 clk of latency against ~500 clk of work, measured with the in-kernel
// timeline).  Here P gets its own TMEM columns instead of aliasing S, which breaks that chain:
//
//   softmax warpgroup t                                    MMA warp of tile t
//   wait S(j); tcgen05.ld S(j) -> registers; S free  ----> QK(j+1) -> S(j+1)     (overlaps the exponentials of j)
//   row max, lazy rescale (rare), exponentials
//   wait PV(j-1) done (P buffer free); st P(j)       ----> PV(j) -> O            (overlaps the exponentials of j+1)
//
// so in steady state a softmax warp never waits for the tensor pipe, and all NT warps of an SM sub-partition keep the
// XU / FMA pipes busy.
//
// Augmented operand layout (PipeCfg::AUG, entry point i2v_fused_self_xframe_aug_fwd): the head dim d = 40 is stored padded
// to 48, and the spare column kAugCol = 40 does three jobs that otherwise cost FMA-pipe instructions per score:
//   * Q column 40 = -m (the row's reference max), K column 40 = 1   ->  the QK^T MMA delivers  x = s - m  directly;
//     the softmax scale * log2(e) is folded into the query projection weights by the caller, so P = 2^x needs no FFMA.
//     m is an integer that bf16 holds exactly; it is rewritten in the query tile (by the row's own thread, between
//     two QK MMAs of the tile) only when the lazy-rescale threshold trips, and the one or two S tiles computed with
//     the previous value take a slow path that subtracts the difference.
//   * V column 40 = 1   ->  O column 40 = sum_j P_ij, the softmax denominator, accumulated by the PV MMA from exactly
//     the bf16 P values it multiplies: no FADD chain, no separate row-sum bookkeeping under rescaling.
// The caller builds these columns for free as zero weight rows plus a bias in the packed QKV projection GEMM.
//
// CTA: NT softmax/epilogue warpgroups (one 128-row query tile each, one thread per row), 1 TMA warp, NT MMA-issuing
// warps (one per tile, so no tile waits behind another tile's barrier; their waits park the warp in hardware).
// One persistent CTA per SM walks the work items (problem, batch row, head, query block); barrier phases run on
// across items, so the epilogue of one item overlaps the query load and the first QK^T of the next.
// TMEM columns per tile: S [BN fp32] | P [BN/2: bf16 pairs] | O [DK fp32]; NT * (1.5 BN + DK) <= 512.
#pragma once
#include <cuda.h>
#include "ptx_sm100.cuh"
#include "dense_attn_sm100.cuh"  // DenseParams / DenseProblem / kRescaleThreshold

namespace i2v {

#ifndef I2V_PARK_NS
#define I2V_PARK_NS 1000
#endif
constexpr uint32_t kParkNs = I2V_PARK_NS;   // suspend-time hint of the producer-side waits
constexpr int kAugCol = 40;   // augmented layout: head-dim column that carries -max (Q), ones (K, V) and the row sum (O)

template <int DK_, int BLOCK_N_, int NT_, int NSTAGES_, int EMU_, int DEG_ = 3, bool AUG_ = false, bool SPLIT_ = false,
          int PAT_ = 0, int MINB_ = 1>
struct PipeCfg {
  static constexpr int PAT = PAT_;          // which pairs of every 8 take the FMA-pipe exp2 (softmax_exp_row)
  static constexpr int MINB = MINB_;        // co-resident CTAs per SM (experiment: one query tile per CTA, 3-4 CTAs per SM)
  static constexpr int DK = DK_;            // head dim rounded up to a multiple of 16
  static constexpr int BLOCK_N = BLOCK_N_;  // keys per tile
  static constexpr int NT = NT_;            // query tiles (= softmax warpgroups) per CTA
  static constexpr int NSTAGES = NSTAGES_;
  static constexpr int EMU = EMU_;          // of every 8 column pairs, how many take the FMA-pipe exp2
  static constexpr int DEG = DEG_;
  static constexpr bool AUG = AUG_;         // augmented operand layout (see the header comment)
  // SPLIT: two threads per query row -- every tile has two softmax warpgroups, each taking BLOCK_N / 2 of a score
  // tile's columns (TMEM lanes are bound to warp % 4, so warps w and w + 4 of a tile share a lane quarter).  Twice the
  // warps per SM sub-partition at half the registers each: the softmax code is latency-bound (issue slots 56 % busy
  // with three in-order warps per sub-partition), not throughput-bound.
  static constexpr bool SPLIT = SPLIT_;
  static constexpr int HALVES = SPLIT ? 2 : 1;
  static constexpr int SM_WARPS = 4 * NT * HALVES;          // softmax warps
  static constexpr int THREADS = (SM_WARPS + 1 + NT) * 32;  // softmax warpgroups, 1 TMA warp, NT MMA warps
  static_assert(!SPLIT || (AUG && BLOCK_N == 64 && DK == 48), "the column-split softmax exists for the augmented d = 40 layout");
  static constexpr int KSTEPS = DK / 16;
  static constexpr int KSUB = (DK + 63) / 64;          // 64-column (128-byte) swizzle sub-tiles per row: 1 (d <= 64) or 2 (d = 80)
  static constexpr int Q_SUB_BYTES = 128 * 128;
  static constexpr int Q_TILE_BYTES = KSUB * Q_SUB_BYTES;
  static constexpr int KV_SUB_BYTES = BLOCK_N * 128;
  static constexpr int KV_TILE_BYTES = KSUB * KV_SUB_BYTES;
  static constexpr int BAR_BYTES = 512;
  static constexpr int XCHG_BYTES = SPLIT ? NT * 2 * 128 * 4 : 0;   // SPLIT: per-row maxima the two halves exchange (slow path)
  static constexpr int SMEM_BYTES = NT * Q_TILE_BYTES + NSTAGES * 2 * KV_TILE_BYTES + BAR_BYTES + XCHG_BYTES + 1024;
  static constexpr int TILE_COLS = BLOCK_N + BLOCK_N / 2 + DK;
  static constexpr int TMEM_S = 0, TMEM_P = BLOCK_N, TMEM_O = BLOCK_N + BLOCK_N / 2;  // offsets within a tile's columns
  static_assert(DK % 16 == 0 && DK <= 128, "pipelined kernel: head dim <= 128 (two swizzle sub-tiles)");
  static_assert(!AUG || KSUB == 1, "the augmented layout is the d = 40 -> 48 case");
  static_assert(BLOCK_N % 16 == 0 && BLOCK_N <= 128, "BLOCK_N");
  static constexpr int TMEM_COLS = NT * TILE_COLS <= 128 ? 128 : NT * TILE_COLS <= 256 ? 256 : 512;   // allocation: power of two
  static_assert(NT * TILE_COLS <= 512 && MINB * TMEM_COLS <= 512, "TMEM budget");
  static_assert(MINB * (SMEM_BYTES + 1024) <= 228 * 1024, "smem budget");
  static_assert(KV_SUB_BYTES % 1024 == 0, "K/V tiles must keep the 1024-byte swizzle-atom alignment");
  static_assert(THREADS <= 1024, "CTA size");
};

template <class Cfg>
__global__ void __launch_bounds__(Cfg::THREADS, Cfg::MINB) dense_attn_pipe_kernel(const __grid_constant__ DenseParams P) {
  constexpr int DK = Cfg::DK, BN = Cfg::BLOCK_N, NS = Cfg::NSTAGES, KSTEPS = Cfg::KSTEPS, NT = Cfg::NT;
  constexpr int kRowThreads = 128 * Cfg::HALVES;   // threads that take part in a tile's S / P hand-offs

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sm_q = smem;                                  // [NT][128 rows][128 B]
  uint8_t* sm_k = sm_q + NT * Cfg::Q_TILE_BYTES;         // [NS][BN rows][128 B]
  uint8_t* sm_v = sm_k + NS * Cfg::KV_TILE_BYTES;        // [NS][BN rows][128 B]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm_v + NS * Cfg::KV_TILE_BYTES);
  uint64_t* bar_q_full = bars;                        // [1]    TMA -> MMA: the query tiles of the work item landed
  uint64_t* bar_q_empty = bars + 1;                   // [1]    MMA -> TMA: every QK of the work item retired (NT arrivals)
  uint64_t* bar_k_full = bars + 2;                    // [NS]
  uint64_t* bar_v_full = bars + 2 + NS;               // [NS]
  uint64_t* bar_kv_empty = bars + 2 + 2 * NS;         // [NS]   one arrival per MMA warp (NT)
  uint64_t* bar_s_full = bars + 2 + 3 * NS;           // [NT]   MMA -> softmax: S(j) written
  uint64_t* bar_s_free = bar_s_full + NT;             // [NT]   softmax -> MMA: S(j) is in registers (128 arrivals)
  uint64_t* bar_p_full = bar_s_free + NT;             // [NT]   softmax -> MMA: P(j) stored (128 arrivals)
  uint64_t* bar_pv_done = bar_p_full + NT;            // [NT]   MMA -> softmax: PV(j) retired (P free, O consistent)
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(bar_pv_done + NT);
  static_assert((3 + 3 * NS + 4 * NT) * 8 <= Cfg::BAR_BYTES, "barrier area");
  float* sm_xchg = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + Cfg::BAR_BYTES);   // [NT][2][128] (SPLIT)

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  constexpr int kTmaWarp = Cfg::SM_WARPS, kMmaWarp0 = Cfg::SM_WARPS + 1;

  // Persistent CTA: work item = (problem, batch row, head, query block of NT tiles), q-block fastest so the CTAs that
  // share one (batch, head) K/V run at the same time (L2 reuse).  Every role walks the same item sequence; barrier
  // phases run on across items, so the epilogue of one item overlaps the query load and the first QK of the next.
  const int n_items = P.q_blocks * P.heads * P.batch * P.nprob;
  const int n_kv = (P.skv + BN - 1) / BN;
  struct Item { const DenseProblem* prob; int h, b, q0, ntiles, bkv; };
  auto decode = [&](int x) {
    Item it;
    const int qb = x % P.q_blocks;  x /= P.q_blocks;
    it.h = x % P.heads;             x /= P.heads;
    it.b = x % P.batch;             x /= P.batch;
    it.prob = &P.prob[x];
    it.q0 = qb * (128 * NT);
    it.ntiles = min(NT, (P.sq - it.q0 + 127) / 128);
    it.bkv = it.b / it.prob->kv_group;
    return it;
  };

  if (threadIdx.x == 0) {
    mbar_init(bar_q_full, 1);
    mbar_init(bar_q_empty, NT);
    for (int s = 0; s < NS; ++s) {
      mbar_init(bar_k_full + s, 1);
      mbar_init(bar_v_full + s, 1);
      mbar_init(bar_kv_empty + s, NT);
    }
    for (int t = 0; t < NT; ++t) {
      mbar_init(bar_s_full + t, 1);
      mbar_init(bar_s_free + t, kRowThreads);
      mbar_init(bar_p_full + t, kRowThreads);
      mbar_init(bar_pv_done + t, 1);
    }
    mbar_fence_init();
  }
  if (warp == kMmaWarp0) tmem_alloc<Cfg::TMEM_COLS>(tmem_base_slot);
  if (warp == kTmaWarp && lane == 0) {
    for (int i = 0; i < P.nprob; ++i) {
      tma_prefetch_desc(&P.prob[i].tm_q);
      tma_prefetch_desc(&P.prob[i].tm_k);
      tma_prefetch_desc(&P.prob[i].tm_v);
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_slot;

  if (warp == kTmaWarp) {
    // =========================== TMA producer ===========================
    if (lane == 0) {
      uint32_t g = 0;   // K/V tiles loaded so far (stage = g % NS)
      int n = 0;
      for (int x = blockIdx.x; x < n_items; x += gridDim.x, ++n) {
        const Item it = decode(x);
        if (n > 0) mbar_wait_parked(bar_q_empty, (n - 1) & 1, kParkNs);   // the previous item's QK MMAs are done with sm_q
        mbar_arrive_expect_tx(bar_q_full, it.ntiles * Cfg::Q_TILE_BYTES);
        for (int t = 0; t < it.ntiles; ++t)
#pragma unroll
          for (int sub = 0; sub < Cfg::KSUB; ++sub)
            tma_load_4d(sm_q + t * Cfg::Q_TILE_BYTES + sub * Cfg::Q_SUB_BYTES, &it.prob->tm_q, bar_q_full, sub * 64, it.h,
                        it.q0 + t * 128, it.b, kEvictFirst);
        for (int j = 0; j < n_kv; ++j, ++g) {
          const int s = g % NS;
          mbar_wait_parked(bar_kv_empty + s, ((g / NS) & 1) ^ 1, kParkNs);
          mbar_arrive_expect_tx(bar_k_full + s, Cfg::KV_TILE_BYTES);
#pragma unroll
          for (int sub = 0; sub < Cfg::KSUB; ++sub)
            tma_load_4d(sm_k + s * Cfg::KV_TILE_BYTES + sub * Cfg::KV_SUB_BYTES, &it.prob->tm_k, bar_k_full + s, sub * 64, it.h,
                        j * BN, it.bkv, kEvictLast);
          mbar_arrive_expect_tx(bar_v_full + s, Cfg::KV_TILE_BYTES);
#pragma unroll
          for (int sub = 0; sub < Cfg::KSUB; ++sub)
            tma_load_4d(sm_v + s * Cfg::KV_TILE_BYTES + sub * Cfg::KV_SUB_BYTES, &it.prob->tm_v, bar_v_full + s, sub * 64, it.h,
                        j * BN, it.bkv, kEvictLast);
        }
      }
    }
  } else if (warp >= kMmaWarp0) {
    // =========================== MMA issuer of tile t ===========================
    // Per KV tile j of its query tile: QK(j+1) as soon as K(j+1) has landed and S(j) has been read out, then PV(j) once
    // V(j) has landed and P(j) is stored.  The waits park the warp in hardware.
    const int t = warp - kMmaWarp0;
    I2V_TRACE_DECL
    I2V_TRACE_INIT(8 + t)
    constexpr uint32_t idesc_qk = make_idesc_bf16(128, BN, 0, 0);
    constexpr uint32_t idesc_pv = make_idesc_bf16(128, DK, 0, 1);
    const uint32_t qa = smem_u32(sm_q + t * Cfg::Q_TILE_BYTES) >> 4;
    const uint32_t k_addr = smem_u32(sm_k);
    const uint32_t v_addr = smem_u32(sm_v);
    const uint64_t desc_k_major = make_smem_desc_sw128(0, 16, 1024);
    const uint64_t desc_v = make_smem_desc_sw128(0, Cfg::KV_SUB_BYTES, 1024);   // MN-major: 64-column atoms KV_SUB_BYTES apart
    const uint32_t tm_tile = tmem_base + t * Cfg::TILE_COLS;

    auto issue_qk = [&](int s) {
      const uint32_t ka = (k_addr + s * Cfg::KV_TILE_BYTES) >> 4;
#ifdef I2V_EXPERIMENTS
      // PAT == 2 (developer experiment): every QK^T is issued twice (same result) -- does the run time grow by the
      // added tensor time (MMA on the critical path) or not (hidden behind the softmax)?
      if (Cfg::PAT == 2) {
#pragma unroll
        for (int kk = 0; kk < KSTEPS; ++kk) {
          const uint64_t da = desc_k_major | (uint64_t)((qa + kk * 2) & 0x3FFF);
          const uint64_t db = desc_k_major | (uint64_t)((ka + kk * 2) & 0x3FFF);
          umma_ss(tm_tile + Cfg::TMEM_S, da, db, idesc_qk, kk > 0 ? 1u : 0u);
        }
      }
#endif
#pragma unroll
      for (int kk = 0; kk < KSTEPS; ++kk) {
        // 32 bytes per 16-column k-step inside a 64-column sub-tile
        const uint32_t sub = kk >> 2, off = (kk & 3) * 2;
        const uint64_t da = desc_k_major | (uint64_t)((qa + sub * (Cfg::Q_SUB_BYTES >> 4) + off) & 0x3FFF);
        const uint64_t db = desc_k_major | (uint64_t)((ka + sub * (Cfg::KV_SUB_BYTES >> 4) + off) & 0x3FFF);
        umma_ss(tm_tile + Cfg::TMEM_S, da, db, idesc_qk, kk > 0 ? 1u : 0u);
      }
    };
    auto issue_pv = [&](int s, bool accumulate) {
      const uint32_t va = (v_addr + s * Cfg::KV_TILE_BYTES) >> 4;
#pragma unroll
      for (int kk = 0; kk < BN / 16; ++kk) {
        // B = V tile, MN-major: 16 key rows per k-step (2048 B)
        const uint64_t db = desc_v | (uint64_t)((va + kk * (2048 >> 4)) & 0x3FFF);
        umma_ts(tm_tile + Cfg::TMEM_O, tm_tile + Cfg::TMEM_P + kk * 8, db, idesc_pv, (accumulate || kk > 0) ? 1u : 0u);
      }
    };

    uint32_t g0 = 0;    // K/V tiles consumed before this item
    uint32_t it0 = 0;   // KV iterations this tile has run before this item (phase counter of its four barriers)
    int n = 0;
    for (int x = blockIdx.x; x < n_items; x += gridDim.x, ++n) {
      const Item it = decode(x);
      if (t < it.ntiles) {
        // QK(jj): needs the query tiles (jj == 0), K(jj), and the S columns read out by the softmax warps
        auto qk_step = [&](int jj) {
          const uint32_t g = g0 + jj, itg = it0 + jj;
          I2V_TRACE_EV(0x10)
          mbar_wait_parked(bar_k_full + g % NS, (g / NS) & 1, kParkNs);
          I2V_TRACE_EV(0x11)
          if (itg > 0) mbar_wait_parked(bar_s_free + t, (itg - 1) & 1, kParkNs);
          tc_fence_after();
          I2V_TRACE_EV(0x12)
          if (elect_one()) {
            issue_qk(g % NS);
            tc_commit(bar_s_full + t);
          }
          __syncwarp();
          I2V_TRACE_EV(0x13)
        };
        mbar_wait_parked(bar_q_full, n & 1, kParkNs);
        qk_step(0);
        for (int j = 0; j < n_kv; ++j) {
          if (j + 1 < n_kv) qk_step(j + 1);
          const uint32_t g = g0 + j, itg = it0 + j;
          const int s = g % NS;
          mbar_wait_parked(bar_v_full + s, (g / NS) & 1, kParkNs);
          I2V_TRACE_EV(0x14)
          mbar_wait_parked(bar_p_full + t, itg & 1, kParkNs);
          tc_fence_after();
          I2V_TRACE_EV(0x15)
          if (elect_one()) {
            issue_pv(s, j > 0);
            tc_commit(bar_pv_done + t);
            tc_commit(bar_kv_empty + s);   // K(j) (read by QK(j), issued earlier) and V(j) are released together
            if (j == n_kv - 1) tc_commit(bar_q_empty);   // every QK of this item precedes this commit
          }
          __syncwarp();
        }
        it0 += n_kv;
      } else {
        // no query tile for this warp in this item (ragged last q-block): keep the shared barriers' arrival counts
        for (int j = 0; j < n_kv; ++j) {
          const uint32_t g = g0 + j;
          mbar_wait_parked(bar_v_full + g % NS, (g / NS) & 1, kParkNs);   // stay in step with the stage ring
          if (lane == 0) mbar_arrive(bar_kv_empty + g % NS);
        }
        if (lane == 0) mbar_arrive(bar_q_empty);
        __syncwarp();
      }
      g0 += n_kv;
    }
  } else if constexpr (Cfg::SPLIT) {
    // =========================== softmax + epilogue, two threads per query row ===========================
    // warp w of the tile's eight: lane quarter w & 3 (fixed by the hardware: TMEM lanes 32 * (warp % 4) ..), column half
    // w >> 2.  Per KV tile a thread holds 32 scores; the row's other 32 are with its partner (same lane, warp w ^ 4).
    // The two warps of a row group agree on the fast / slow path through one bar.red.or per KV tile (a named barrier
    // over their 64 threads that ORs a predicate); only the slow path exchanges the row maxima, through shared memory.
    constexpr int HB = BN / 2;                    // columns per thread
    const int t = warp >> 3;
    const int quarter = warp & 3, half = (warp >> 2) & 1;
    const int row = quarter * 32 + lane;          // TMEM lane == query row within the tile
    const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
    const uint32_t tm_tile = tmem_base + t * Cfg::TILE_COLS + lane_addr;
    const uint32_t tm_s = tm_tile + Cfg::TMEM_S + half * HB, tm_p = tm_tile + Cfg::TMEM_P + half * (HB / 2);
    const uint32_t tm_o = tm_tile + Cfg::TMEM_O;
    const uint32_t pair_bar = 1 + t * 4 + quarter;   // named barrier of the two warps of this row group (ids 1..12)
    float* xchg_mine = sm_xchg + (t * 2 + half) * 128 + row;
    const float* xchg_other = sm_xchg + (t * 2 + (half ^ 1)) * 128 + row;
    uint8_t* q_maxcol = sm_q + t * Cfg::Q_TILE_BYTES + row * 128 + ((((kAugCol * 2) >> 4) ^ (row & 7)) << 4) +
                        ((kAugCol * 2) & 15);
    uint32_t it0 = 0;   // KV iterations this tile has run before this item
    for (int x = blockIdx.x; x < n_items; x += gridDim.x) {
      const Item item = decode(x);
      if (t >= item.ntiles) continue;
      const DenseProblem& prob = *item.prob;
      const int h = item.h, b = item.b, q0 = item.q0;
      float m_ref = 0.f;        // where the reference max should be (integer, bf16-exact)
      float m_col = 0.f;        // what the query tile's max column held when the current S tile was computed
      bool col_stale = false;   // m_ref moved: the max column must be rewritten (half 0 writes, both halves track it)

      for (int j = 0; j < n_kv; ++j) {
        mbar_wait(bar_s_full + t, (it0 + j) & 1);
        tc_fence_after();
        float sv[HB];
        {
          uint32_t r[HB];
          tmem_ld_x32(tm_s, r);
          tc_wait_ld();
#pragma unroll
          for (int i = 0; i < HB; ++i) sv[i] = __uint_as_float(r[i]);
        }
        float m_col_next = m_col;
        if (col_stale) {
          // QK(j) has retired and QK(j+1) is not issued before all 256 s_free arrivals: the window for the rewrite
          if (half == 0) {
            *reinterpret_cast<uint16_t*>(q_maxcol) = (uint16_t)(__float_as_uint(-m_ref) >> 16);
            fence_proxy_async_smem();
          }
          m_col_next = m_ref;
          col_stale = false;
        }
        tc_fence_before();
        mbar_arrive(bar_s_free + t);

        const int valid = P.skv - j * BN - half * HB;   // columns of this half that exist
        const bool full = P.skv - j * BN >= BN;          // (uniform over the tile)
        float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
        for (int i = 0; i < HB; i += 4) {
          mx0 = fmax3(mx0, sv[i + 0], sv[i + 1]);
          mx1 = fmax3(mx1, sv[i + 2], sv[i + 3]);
        }
        float mx = fmaxf(mx0, mx1);
        const bool slow_local = (j == 0) || (mx + (m_col - m_ref) > kRescaleThreshold) || (m_col != m_ref) || !full;
        uint32_t pk[HB / 2];
        if (named_bar_red_or(pair_bar, 64, slow_local)) {
          // ---- slow path (first tile of an item, a row outgrew its reference, stale max column, ragged tile) ----
          if (!full) {
#pragma unroll
            for (int i = 0; i < HB; ++i)
              if (i >= valid) sv[i] = -INFINITY;
            mx0 = mx1 = -INFINITY;
#pragma unroll
            for (int i = 0; i < HB; i += 4) {
              mx0 = fmax3(mx0, sv[i + 0], sv[i + 1]);
              mx1 = fmax3(mx1, sv[i + 2], sv[i + 3]);
            }
            mx = fmaxf(mx0, mx1);
          }
          *xchg_mine = mx;
          named_bar_sync(pair_bar, 64);
          mx = fmaxf(mx, *xchg_other) + (m_col - m_ref);   // the row's max relative to m_ref, identical in both halves
          named_bar_sync(pair_bar, 64);                     // the slot may be rewritten by the next slow step
          const bool need = (j == 0) || mx > kRescaleThreshold;
          float alpha = 1.f;
          if (need) {
            const float m_int = ceilf(m_ref + mx);
            const uint32_t mb = __float_as_uint(m_int);
            const float m_new = __uint_as_float(m_int >= 0.f ? ((mb + 0xFFFFu) & 0xFFFF0000u) : (mb & 0xFFFF0000u));
            alpha = (j == 0) ? 0.f : ex2_approx(m_ref - m_new);
            m_ref = m_new;
            col_stale = true;
          }
          if (j > 0 && __any_sync(0xffffffffu, need)) {
            // O row *= alpha on this half's 24 accumulator columns; PV(j-1) must have retired, PV(j) is not issued
            // before all p_full arrivals
            mbar_wait(bar_pv_done + t, (it0 + j - 1) & 1);
            tc_fence_after();
#pragma unroll
            for (int cch = 0; cch < 3; ++cch) {
              uint32_t r[8];
              tmem_ld_x8(tm_o + half * 24 + cch * 8, r);
              tc_wait_ld();
#pragma unroll
              for (int i = 0; i < 8; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) * alpha);
              tmem_st_x8(tm_o + half * 24 + cch * 8, r);
            }
          }
          const float delta = m_ref - m_col;   // exact: both are small integers
#pragma unroll
          for (int i = 0; i < HB; ++i) sv[i] -= delta;
          softmax_exp_row<HB, 0, 3, true, true, false>(sv, 1.f, 0.f, pk);
        } else {
          softmax_exp_row<HB, Cfg::EMU, Cfg::DEG, true, true, false>(sv, 1.f, 0.f, pk);
        }
        m_col = m_col_next;
        if (j > 0) {   // (j == 0: the epilogue of the previous item already waited for its last PV)
          mbar_wait(bar_pv_done + t, (it0 + j - 1) & 1);   // PV(j-1) has finished reading the P columns
          tc_fence_after();
        }
        tmem_st_x16(tm_p, pk);
        tc_wait_st();
        tc_fence_before();
        mbar_arrive(bar_p_full + t);
      }

      // ---- epilogue: O / l -> bf16 -> global; half 0 writes columns 0..23, half 1 columns 24..39 ----
      mbar_wait(bar_pv_done + t, (it0 + n_kv - 1) & 1);
      tc_fence_after();
      it0 += n_kv;
      const int qrow = q0 + t * 128 + row;
      __nv_bfloat16* orow = prob.o + (long long)b * prob.o_sb + (long long)qrow * prob.o_ss + (long long)h * prob.o_sh;
      uint32_t ro[3][8], rl[8];
#pragma unroll
      for (int cch = 0; cch < 3; ++cch) tmem_ld_x8(tm_o + half * 24 + cch * 8, ro[cch]);
      tmem_ld_x8(tm_o + kAugCol, rl);            // columns 40..47: the row sum sits in column 40
      tc_wait_ld();
      const float inv_l = 1.f / __uint_as_float(rl[0]);
      if (qrow < P.sq) {
#pragma unroll
        for (int cch = 0; cch < 3; ++cch) {
          const int col = half * 24 + cch * 8;
          if (col < P.d) {
            const uint32_t* r = ro[cch];
            uint4 v;
            v.x = pack_bf16x2(__uint_as_float(r[0]) * inv_l, __uint_as_float(r[1]) * inv_l);
            v.y = pack_bf16x2(__uint_as_float(r[2]) * inv_l, __uint_as_float(r[3]) * inv_l);
            v.z = pack_bf16x2(__uint_as_float(r[4]) * inv_l, __uint_as_float(r[5]) * inv_l);
            v.w = pack_bf16x2(__uint_as_float(r[6]) * inv_l, __uint_as_float(r[7]) * inv_l);
            *reinterpret_cast<uint4*>(orow + col) = v;
          }
        }
      }
    }
  } else {
    // =========================== softmax + epilogue warpgroup of tile t ===========================
    const int t = warp >> 2;
    I2V_TRACE_DECL
    if ((warp & 3) == 0) { I2V_TRACE_INIT(t) }
    else if ((warp & 3) == 3) { I2V_TRACE_INIT(4 + t) }   // the tile's warp on another sub-partition: skew between them
    const int row = (warp & 3) * 32 + lane;      // TMEM lane == query row within the tile
    const uint32_t lane_addr = (uint32_t)((warp & 3) * 32) << 16;
    const uint32_t tm_tile = tmem_base + t * Cfg::TILE_COLS + lane_addr;
    const uint32_t tm_s = tm_tile + Cfg::TMEM_S, tm_p = tm_tile + Cfg::TMEM_P, tm_o = tm_tile + Cfg::TMEM_O;
    const float c = P.scale_log2e;
    // this row's slot in the 128B-swizzled query tile: column kAugCol (16-byte chunk 5) of row `row`
    uint8_t* q_maxcol = sm_q + t * Cfg::Q_TILE_BYTES + row * 128 + ((((kAugCol * 2) >> 4) ^ (row & 7)) << 4) +
                        ((kAugCol * 2) & 15);
    uint32_t it0 = 0;   // KV iterations this tile has run before this item
    for (int x = blockIdx.x; x < n_items; x += gridDim.x) {
      const Item item = decode(x);
      if (t >= item.ntiles) continue;
      const DenseProblem& prob = *item.prob;
      const int h = item.h, b = item.b, q0 = item.q0;
      // plain mode:     m_ref = reference max (integer, log2 domain), l = running row sum
      // augmented mode: the scores arrive as x = s - m_col (see the header comment); m_ref is where the reference max
      //                 should be, m_col what the query tile's max column held when the current S tile was computed
      float m_ref = Cfg::AUG ? 0.f : -INFINITY;
      float m_col = 0.f;
      float l = 0.f;
      bool col_stale = false;   // augmented: m_ref moved, the max column of the query tile must be rewritten

      auto load_scores = [&](float (&sv)[BN]) {
        constexpr int N32 = BN / 32, R16 = (BN % 32) / 16;
#pragma unroll
        for (int cch = 0; cch < N32; ++cch) {
          uint32_t r[32];
          tmem_ld_x32(tm_s + cch * 32, r);
#pragma unroll
          for (int i = 0; i < 32; ++i) sv[cch * 32 + i] = __uint_as_float(r[i]);
        }
        if (R16) {
          uint32_t r[16];
          tmem_ld_x16(tm_s + N32 * 32, r);
#pragma unroll
          for (int i = 0; i < 16; ++i) sv[N32 * 32 + i] = __uint_as_float(r[i]);
        }
        tc_wait_ld();
      };
      auto row_max = [&](const float (&sv)[BN]) -> float {
        float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
        for (int i = 0; i < BN; i += 8) {
          mx0 = fmax3(mx0, sv[i + 0], sv[i + 1]);
          mx1 = fmax3(mx1, sv[i + 2], sv[i + 3]);
          mx2 = fmax3(mx2, sv[i + 4], sv[i + 5]);
          mx3 = fmax3(mx3, sv[i + 6], sv[i + 7]);
        }
        return fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
      };
      // multiply the O accumulator row by alpha; PV(j-1) must have retired and PV(j) is not issued before our p_full
      auto rescale_o = [&](int j, float alpha) {
        mbar_wait(bar_pv_done + t, (it0 + j - 1) & 1);
        tc_fence_after();
#pragma unroll
        for (int cch = 0; cch < DK / 16; ++cch) {
          uint32_t r[16];
          tmem_ld_x16(tm_o + cch * 16, r);
          tc_wait_ld();
#pragma unroll
          for (int i = 0; i < 16; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) * alpha);
          tmem_st_x16(tm_o + cch * 16, r);
        }
      };
      auto store_p = [&](int j, const uint32_t (&pk)[BN / 2]) {
        if (j > 0) {   // (j == 0: the epilogue of the previous item already waited for its last PV)
          mbar_wait(bar_pv_done + t, (it0 + j - 1) & 1);   // PV(j-1) has finished reading the P columns
          tc_fence_after();
        }
        I2V_TRACE_EV(0x6)
        constexpr int H = BN / 2, N16 = H / 16, R8 = (H % 16) / 8;
#pragma unroll
        for (int cch = 0; cch < N16; ++cch) {
          uint32_t r[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) r[i] = pk[cch * 16 + i];
          tmem_st_x16(tm_p + cch * 16, r);
        }
        if (R8) {
          uint32_t r[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) r[i] = pk[N16 * 16 + i];
          tmem_st_x8(tm_p + N16 * 16, r);
        }
        tc_wait_st();
        tc_fence_before();
        mbar_arrive(bar_p_full + t);
      };

      for (int j = 0; j < n_kv; ++j) {
        I2V_TRACE_EV(0x1)
        mbar_wait(bar_s_full + t, (it0 + j) & 1);
        tc_fence_after();
        I2V_TRACE_EV(0x2)
        float sv[BN];
        load_scores(sv);
        I2V_TRACE_EV(0x3)
        float m_col_next = m_col;
        if (Cfg::AUG && col_stale) {
          // QK(j) has retired (we hold its result) and QK(j+1) is not issued before our s_free arrival: the only
          // window in which the query tile may be touched.  QK(j+1) onwards see the new column.
          *reinterpret_cast<uint16_t*>(q_maxcol) = (uint16_t)(__float_as_uint(-m_ref) >> 16);   // m_ref is bf16-exact
          fence_proxy_async_smem();
          m_col_next = m_ref;
          col_stale = false;
        }
        tc_fence_before();
        mbar_arrive(bar_s_free + t);   // the MMA warp may overwrite S with QK(j+1) from here on

        const int valid = P.skv - j * BN;
        const bool full = valid >= BN;
        if (!full) {  // ragged last tile only (warp-uniform)
#pragma unroll
          for (int i = 0; i < BN; ++i)
            if (i >= valid) sv[i] = -INFINITY;
        }
        uint32_t pk[BN / 2];
        I2V_TRACE_EV(0x4)
        if (!Cfg::AUG) {
          const float mx = row_max(sv) * c;
          // lazy rescale: the reference max only moves when a row outgrows it by more than 2^kRescaleThreshold
          const bool need = mx > m_ref + kRescaleThreshold;   // m_ref = -inf on the first tile
          if (__any_sync(0xffffffffu, need)) {
            float alpha = 1.f;
            if (need) {
              const float m_new = ceilf(mx);
              alpha = (m_ref == -INFINITY) ? 0.f : ex2_approx(m_ref - m_new);
              m_ref = m_new;
            }
            if (j > 0) rescale_o(j, alpha);
            l *= alpha;
          }
          if (full) l += softmax_exp_row<BN, Cfg::EMU, Cfg::DEG, true>(sv, c, m_ref, pk);
          else      l += softmax_exp_row<BN, 0, 3, true>(sv, c, m_ref, pk);   // -inf padding must go through MUFU
        } else {
          // sv = s - m_col.  Fast path (nothing moved): P = 2^sv straight away; the row sum comes out of the PV MMA
          // through the ones column of V.
          const float mx = row_max(sv) + (m_col - m_ref);   // relative to m_ref
          const bool need = (j == 0) || mx > kRescaleThreshold;
          const bool slow = need || (m_col != m_ref) || !full;
          if (__any_sync(0xffffffffu, slow)) {
            if (__any_sync(0xffffffffu, need)) {
              float alpha = 1.f;
              if (need) {
                // new reference: an integer >= the row max that bf16 holds exactly (round the magnitude up / down)
                const float m_int = ceilf(m_ref + mx);
                const uint32_t mb = __float_as_uint(m_int);
                const float m_new = __uint_as_float(m_int >= 0.f ? ((mb + 0xFFFFu) & 0xFFFF0000u) : (mb & 0xFFFF0000u));
                alpha = (j == 0) ? 0.f : ex2_approx(m_ref - m_new);
                m_ref = m_new;
                col_stale = true;
              }
              if (j > 0) rescale_o(j, alpha);   // scales the row-sum column with the rest of the row
            }
            const float delta = m_ref - m_col;   // exact: both are small integers
#pragma unroll
            for (int i = 0; i < BN; ++i) sv[i] -= delta;
            softmax_exp_row<BN, 0, 3, true, true, false>(sv, 1.f, 0.f, pk);
          } else {
            softmax_exp_row<BN, Cfg::EMU, Cfg::DEG, true, true, false, Cfg::PAT>(sv, 1.f, 0.f, pk);
          }
        }
        m_col = m_col_next;
        I2V_TRACE_EV(0x5)
        store_p(j, pk);
        I2V_TRACE_EV(0x7)
      }

      // ---- epilogue: O / l -> bf16 -> global ----
      mbar_wait(bar_pv_done + t, (it0 + n_kv - 1) & 1);
      tc_fence_after();
      it0 += n_kv;
      const int qrow = q0 + t * 128 + row;
      __nv_bfloat16* orow = prob.o + (long long)b * prob.o_sb + (long long)qrow * prob.o_ss + (long long)h * prob.o_sh;
      uint32_t ro[DK / 16][16];
#pragma unroll
      for (int cch = 0; cch < DK / 16; ++cch) tmem_ld_x16(tm_o + cch * 16, ro[cch]);
      tc_wait_ld();
      if (Cfg::AUG) l = __uint_as_float(ro[kAugCol / 16][kAugCol % 16]);   // sum_j P_ij * 1
      const float inv_l = 1.f / l;
      if (qrow < P.sq) {
#pragma unroll
        for (int cch = 0; cch < DK / 16; ++cch) {
#pragma unroll
          for (int g8 = 0; g8 < 2; ++g8) {
            const int col = cch * 16 + g8 * 8;
            if (col < P.d) {  // d is a multiple of 8
              const uint32_t* r = ro[cch];
              uint4 v;
              v.x = pack_bf16x2(__uint_as_float(r[g8 * 8 + 0]) * inv_l, __uint_as_float(r[g8 * 8 + 1]) * inv_l);
              v.y = pack_bf16x2(__uint_as_float(r[g8 * 8 + 2]) * inv_l, __uint_as_float(r[g8 * 8 + 3]) * inv_l);
              v.z = pack_bf16x2(__uint_as_float(r[g8 * 8 + 4]) * inv_l, __uint_as_float(r[g8 * 8 + 5]) * inv_l);
              v.w = pack_bf16x2(__uint_as_float(r[g8 * 8 + 6]) * inv_l, __uint_as_float(r[g8 * 8 + 7]) * inv_l);
              *reinterpret_cast<uint4*>(orow + col) = v;
            }
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp0) {
    tc_fence_after();
    tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
  }
}

}  // namespace i2v
