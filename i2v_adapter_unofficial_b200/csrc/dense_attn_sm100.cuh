// Dense attention forward on sm_100a: QK^T and PV on tcgen05 MMAs with TMEM accumulators, K/V tiles staged
// by TMA into a multi-stage shared-memory ring, online softmax in registers (one thread per query row).
//
// One launch serves up to two independent "problems" that share the shape (batch, heads, sq, skv, d):
//   * K1 spatial self-attention            (reference: src/modules/i2v_adapter.py:468-473)
//   * K2 I2V-Adapter cross-frame attention (reference: src/modules/i2v_adapter.py:484-492), kv_group = num_frames so
//     every frame of a video reads the single frame-0 K/V copy instead of the F-times repeated tensor (:485)
//   * K3 IP-Adapter decoupled cross-attention (diffusers IPAdapterAttnProcessor2_0, installed at
//     src/models/unet_motion_cross_frame_attn.py:1264-1279): one KV tile holding [text | image] tokens with a
//     two-segment softmax (seg_split) and the image branch scaled by seg_scale.
//
// CTA:  warps 0-3 softmax/epilogue warpgroup 0, warps 4-7 softmax/epilogue warpgroup 1, warp 8 TMA producer,
//       warps 9 .. 9+NMW-1 MMA issuers (warp 9 also owns the TMEM allocation).
// A CTA owns NT = 2*TPW query tiles of 128 rows; tile i belongs to warpgroup i % 2.  With TPW = 2 a warpgroup
// alternates between its two tiles, so the QK^T / PV MMAs of one tile run while the warpgroup does the softmax of
// the other one and the softmax warps never idle on the tensor pipe round trip (at d = 40 the kernel is bound by
// the exponentials, not by the MMAs).
// TMEM columns:       [i*BN, (i+1)*BN) S_i / P_i (bf16 P aliases the fp32 scores)   [NT*BN + i*DK, +DK) O_i
#pragma once
#include <cuda.h>
#include "ptx_sm100.cuh"

// Developer timeline trace (-DI2V_TRACE): one CTA records clock64() at phase boundaries of one softmax warp per
// warpgroup and of the MMA warps into a global buffer, [warp slot][event] = (tag << 48) | (clock & 0xffffffffffff).
#ifdef I2V_TRACE
#define I2V_TRACE_DECL unsigned long long* tr_ptr = nullptr; int tr_n = 0;
#define I2V_TRACE_INIT(slot)                                                                         \
  if (P.trace != nullptr && blockIdx.x == P.trace_cta && lane == 0) tr_ptr = P.trace + (slot) * 1024;
#define I2V_TRACE_EV(tag)                                                                            \
  if (tr_ptr != nullptr && tr_n < 1024) tr_ptr[tr_n++] = ((unsigned long long)(tag) << 48) | (clock64() & 0xffffffffffffull);
#else
#define I2V_TRACE_DECL
#define I2V_TRACE_INIT(slot)
#define I2V_TRACE_EV(tag)
#endif

namespace i2v {

struct DenseProblem {
  CUtensorMap tm_q;  // dims (d, heads, sq, batch), box (64, 1, 128, 1), SWIZZLE_128B
  CUtensorMap tm_k;  // dims (d, heads, skv, batch_kv), box (64, 1, BLOCK_N, 1)
  CUtensorMap tm_v;
  __nv_bfloat16* o;
  long long o_sb, o_ss, o_sh;  // element strides: batch, sequence, head (d contiguous)
  int kv_group;                // K/V batch index = b / kv_group
  int pad_;
};

struct DenseParams {
  DenseProblem prob[2];
  int nprob;
  int batch, heads, sq, skv, d;
  int q_blocks;        // ceil(sq / (128 * Cfg::NT))
  float scale_log2e;   // softmax scale * log2(e)
  int seg_split;       // <0: plain softmax. >=0: two-segment softmax (single KV tile), columns >= seg_split
  float seg_scale;     //      form the second segment whose normalised probabilities are multiplied by seg_scale
  unsigned long long* trace;  // developer timeline trace buffer (only read by -DI2V_TRACE builds), else null
  int trace_cta;
};

// EMU_ = how many of every 8 (key, key+1) pairs get their 2^x from the FMA-pipe polynomial instead of MUFU.EX2.
// At d = 40 the kernel is exp-bound (16384 exps per 128x128 tile at 16 MUFU/clk/SM = 1024 clk against 384 clk of
// MMA), so part of the exponentials is moved to the otherwise idle FMA pipe.
template <int DK_, int BLOCK_N_, int NSTAGES_, int EMU_ = 0, int MIN_CTAS_ = 1, int TPW_ = 1, int NMW_ = 0, int NT_ = 0>
struct DenseCfg {
  static constexpr int TPW = TPW_;      // query tiles per softmax warpgroup
  static constexpr int NT = NT_ > 0 ? NT_ : 2 * TPW_;   // query tiles per CTA (NT_ = 1: a single tile, warpgroup 1 idles)
  static constexpr int NMW = NMW_ > 0 ? NMW_ : NT;  // MMA-issuing warps (tile t is issued by warp t % NMW)
  static constexpr int THREADS = (9 + NMW) * 32;
  static constexpr int EMU = EMU_;
  static constexpr int MIN_CTAS = MIN_CTAS_;  // co-resident CTAs per SM the register / TMEM budget is sized for
  static constexpr int DK = DK_;            // head dim rounded up to a multiple of 16 (MMA K of QK^T, N of PV)
  static constexpr int BLOCK_N = BLOCK_N_;  // keys per tile
  static constexpr int NSTAGES = NSTAGES_;
  static constexpr int KSUB = (DK + 63) / 64;  // 64-column (128-byte) swizzle sub-tiles per row
  static constexpr int KSTEPS = DK / 16;
  static constexpr int Q_SUB_BYTES = 128 * 128;
  static constexpr int KV_SUB_BYTES = BLOCK_N * 128;
  static constexpr int Q_TILE_BYTES = KSUB * Q_SUB_BYTES;
  static constexpr int KV_TILE_BYTES = KSUB * KV_SUB_BYTES;
  static constexpr int BAR_BYTES = 512;
  static constexpr int SMEM_BYTES = NT * Q_TILE_BYTES + NSTAGES * 2 * KV_TILE_BYTES + BAR_BYTES + 1024;
  static constexpr int TMEM_S0 = 0, TMEM_O0 = NT * BLOCK_N;  // tile i: S at TMEM_S0 + i*BLOCK_N, O at TMEM_O0 + i*DK
  static constexpr int TMEM_COLS_USED = NT * (BLOCK_N + DK);
  static constexpr int TMEM_ALLOC = TMEM_COLS_USED <= 32 ? 32 : TMEM_COLS_USED <= 64 ? 64 : TMEM_COLS_USED <= 128 ? 128
                                    : TMEM_COLS_USED <= 256 ? 256 : 512;
  static_assert(TMEM_COLS_USED <= 512, "TMEM budget");
  static_assert(TMEM_ALLOC * MIN_CTAS <= 512, "TMEM budget across co-resident CTAs");
  static_assert(DK % 16 == 0 && DK >= 16 && DK <= 256, "DK");
  static_assert(BLOCK_N % 32 == 0 && BLOCK_N <= 128, "BLOCK_N");
  static_assert(SMEM_BYTES * MIN_CTAS <= 227 * 1024, "smem budget");
};

constexpr float kRescaleThreshold = 8.0f;  // lazy O rescale: only when the running max grows by > 2^8

template <class Cfg>
__global__ void __launch_bounds__(Cfg::THREADS, Cfg::MIN_CTAS) dense_attn_kernel(const __grid_constant__ DenseParams P) {
  constexpr int DK = Cfg::DK, BN = Cfg::BLOCK_N, NS = Cfg::NSTAGES, KSUB = Cfg::KSUB, KSTEPS = Cfg::KSTEPS;
  constexpr int NT = Cfg::NT, TPW = Cfg::TPW;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sm_q = smem;                                  // [NT][KSUB][128 rows][128 B]
  uint8_t* sm_k = sm_q + NT * Cfg::Q_TILE_BYTES;         // [NS][KSUB][BN rows][128 B]
  uint8_t* sm_v = sm_k + NS * Cfg::KV_TILE_BYTES;        // [NS][KSUB][BN rows][128 B]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm_v + NS * Cfg::KV_TILE_BYTES);
  uint64_t* bar_q_full = bars + 0;      // [1]
  uint64_t* bar_k_full = bars + 1;      // [NS]
  uint64_t* bar_v_full = bars + 1 + NS; // [NS]
  uint64_t* bar_kv_empty = bars + 1 + 2 * NS;  // [NS]
  uint64_t* bar_s_full = bars + 1 + 3 * NS;            // [NT]
  uint64_t* bar_p_full = bars + 1 + 3 * NS + NT;       // [NT]
  uint64_t* bar_o_full = bars + 1 + 3 * NS + 2 * NT;   // [NT]
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(bars + 1 + 3 * NS + 3 * NT);
  static_assert((2 + 3 * NS + 3 * NT) * 8 <= Cfg::BAR_BYTES, "barrier area");

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // ---- work decomposition: q-block fastest so the CTAs sharing one (batch, head) K/V run together ----
  int x = blockIdx.x;
  const int qb = x % P.q_blocks;  x /= P.q_blocks;
  const int h = x % P.heads;      x /= P.heads;
  const int b = x % P.batch;      x /= P.batch;
  const int pi = x;  // problem index
  const DenseProblem& prob = P.prob[pi];
  const int q0 = qb * (128 * NT);
  const int ntiles = min(NT, (P.sq - q0 + 127) / 128);  // active query tiles of this CTA
  const int n_kv = (P.skv + BN - 1) / BN;
  const int bkv = b / prob.kv_group;

  if (threadIdx.x == 0) {
    mbar_init(bar_q_full, 1);
    for (int s = 0; s < NS; ++s) {
      mbar_init(bar_k_full + s, 1);
      mbar_init(bar_v_full + s, 1);
      mbar_init(bar_kv_empty + s, min(Cfg::NMW, ntiles));  // one commit per MMA warp that owns a tile
    }
    for (int t = 0; t < NT; ++t) {
      mbar_init(bar_s_full + t, 1);
      mbar_init(bar_p_full + t, 128);
      mbar_init(bar_o_full + t, 1);
    }
    mbar_fence_init();
  }
  if (warp == 9) {
    tmem_alloc<Cfg::TMEM_ALLOC>(tmem_base_slot);
  }
  if (warp == 8 && lane == 0) {
    tma_prefetch_desc(&prob.tm_q);
    tma_prefetch_desc(&prob.tm_k);
    tma_prefetch_desc(&prob.tm_v);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_slot;

  if (warp == 8) {
    // =========================== TMA producer ===========================
    if (lane == 0) {
      mbar_arrive_expect_tx(bar_q_full, ntiles * Cfg::Q_TILE_BYTES);
      for (int t = 0; t < ntiles; ++t)
        for (int ks = 0; ks < KSUB; ++ks)
          tma_load_4d(sm_q + t * Cfg::Q_TILE_BYTES + ks * Cfg::Q_SUB_BYTES, &prob.tm_q, bar_q_full, ks * 64, h,
                      q0 + t * 128, b, kEvictFirst);
      for (int j = 0; j < n_kv; ++j) {
        const int s = j % NS;
        const uint32_t ph = (j / NS) & 1;
        mbar_wait(bar_kv_empty + s, ph ^ 1);
        mbar_arrive_expect_tx(bar_k_full + s, Cfg::KV_TILE_BYTES);
        for (int ks = 0; ks < KSUB; ++ks)
          tma_load_4d(sm_k + s * Cfg::KV_TILE_BYTES + ks * Cfg::KV_SUB_BYTES, &prob.tm_k, bar_k_full + s, ks * 64, h,
                      j * BN, bkv, kEvictLast);
        mbar_arrive_expect_tx(bar_v_full + s, Cfg::KV_TILE_BYTES);
        for (int ks = 0; ks < KSUB; ++ks)
          tma_load_4d(sm_v + s * Cfg::KV_TILE_BYTES + ks * Cfg::KV_SUB_BYTES, &prob.tm_v, bar_v_full + s, ks * 64, h,
                      j * BN, bkv, kEvictLast);
      }
    }
  } else if (warp >= 9) {
    // =========================== MMA issuers ===========================
    // NMW warps; warp w owns query tiles w, w + NMW, ...  All 32 lanes wait on the barriers, one elected lane issues
    // (the elect.sync form keeps the tcgen05 operands in uniform registers; a divergent `lane == 0` region makes
    // the compiler wrap every UTCHMMA in a serialising loop and turns the issuing thread into the bottleneck).
    const int mw = warp - 9;
    I2V_TRACE_DECL
    I2V_TRACE_INIT(8 + mw)
    if (mw < ntiles) {
      constexpr uint32_t idesc_qk = make_idesc_bf16(128, BN, 0, 0);
      constexpr uint32_t idesc_pv = make_idesc_bf16(128, DK, 0, 1);
      const uint32_t q_addr = smem_u32(sm_q);
      const uint32_t k_addr = smem_u32(sm_k);
      const uint32_t v_addr = smem_u32(sm_v);
      // descriptor templates: everything except the 14-bit start-address field is constant
      const uint64_t desc_k_major = make_smem_desc_sw128(0, 16, 1024);
      const uint64_t desc_v = make_smem_desc_sw128(0, Cfg::KV_SUB_BYTES, 1024);

      auto issue_qk = [&](int t, int s) {
        const uint32_t qa = (q_addr + t * Cfg::Q_TILE_BYTES) >> 4;
        const uint32_t ka = (k_addr + s * Cfg::KV_TILE_BYTES) >> 4;
        const uint32_t ds = tmem_base + Cfg::TMEM_S0 + t * BN;
#pragma unroll
        for (int kk = 0; kk < KSTEPS; ++kk) {
          const uint32_t sub = kk >> 2, off = (kk & 3) * 32;
          const uint64_t da = desc_k_major | (uint64_t)((qa + ((sub * Cfg::Q_SUB_BYTES + off) >> 4)) & 0x3FFF);
          const uint64_t db = desc_k_major | (uint64_t)((ka + ((sub * Cfg::KV_SUB_BYTES + off) >> 4)) & 0x3FFF);
          umma_ss(ds, da, db, idesc_qk, kk > 0 ? 1u : 0u);
        }
      };
      auto issue_pv = [&](int t, int s, bool accumulate) {
        const uint32_t va = (v_addr + s * Cfg::KV_TILE_BYTES) >> 4;
        const uint32_t ds = tmem_base + Cfg::TMEM_S0 + t * BN;
        const uint32_t dout = tmem_base + Cfg::TMEM_O0 + t * DK;
#pragma unroll
        for (int kk = 0; kk < BN / 16; ++kk) {
          // B = V tile, MN-major: 16 key rows per k-step (2048 B), 64-column atoms KV_SUB_BYTES apart.
          const uint64_t db = desc_v | (uint64_t)((va + kk * (2048 >> 4)) & 0x3FFF);
          umma_ts(dout, ds + kk * 8, db, idesc_pv, (accumulate || kk > 0) ? 1u : 0u);
        }
      };

      mbar_wait(bar_q_full, 0);
      mbar_wait(bar_k_full + 0, 0);
      tc_fence_after();
      if (elect_one()) {
        for (int t = mw; t < ntiles; t += Cfg::NMW) {
          issue_qk(t, 0);
          tc_commit(bar_s_full + t);
        }
      }
      __syncwarp();
      for (int j = 0; j < n_kv; ++j) {
        const int s = j % NS;
        const uint32_t ph = (j / NS) & 1;
        const int sn = (j + 1) % NS;
        const uint32_t phn = ((j + 1) / NS) & 1;
        const bool more = j + 1 < n_kv;
        mbar_wait(bar_v_full + s, ph);
        if (more) mbar_wait(bar_k_full + sn, phn);
        for (int t = mw; t < ntiles; t += Cfg::NMW) {
          I2V_TRACE_EV(0x100 + t)
          mbar_wait(bar_p_full + t, j & 1);
          tc_fence_after();
          I2V_TRACE_EV(0x110 + t)
          if (elect_one()) {
            issue_pv(t, s, j > 0);
            if (more) {
              issue_qk(t, sn);
              tc_commit(bar_s_full + t);
            } else {
              tc_commit(bar_o_full + t);
            }
          }
          __syncwarp();
          I2V_TRACE_EV(0x120 + t)
        }
        if (elect_one()) tc_commit(bar_kv_empty + s);
        __syncwarp();
      }
    }
  } else {
    // =========================== softmax + epilogue warpgroups ===========================
    const int wg = warp >> 2;                    // warpgroup 0 / 1 owns query tiles wg, wg + 2, ...
    const int row = (warp & 3) * 32 + lane;      // TMEM lane == query row within a tile
    const uint32_t lane_addr = (uint32_t)((warp & 3) * 32) << 16;
    const float c = P.scale_log2e;
    const bool two_seg = P.seg_split >= 0;
    I2V_TRACE_DECL
    if ((warp & 3) == 0) { I2V_TRACE_INIT(wg) }
    float m_ref[TPW], l[TPW];  // per tile: running reference max (scaled log2 domain) and running row sum
#pragma unroll
    for (int tt = 0; tt < TPW; ++tt) { m_ref[tt] = -INFINITY; l[tt] = 0.f; }

    for (int j = 0; j < n_kv; ++j) {
#pragma unroll
      for (int tt = 0; tt < TPW; ++tt) {
        const int t = tt * 2 + wg;  // query tile index within the CTA
        if (t >= ntiles) continue;
        const uint32_t tm_s = tmem_base + Cfg::TMEM_S0 + t * BN + lane_addr;
        const uint32_t tm_o = tmem_base + Cfg::TMEM_O0 + t * DK + lane_addr;
        I2V_TRACE_EV(0x10 + t)
        mbar_wait(bar_s_full + t, j & 1);
        tc_fence_after();
        I2V_TRACE_EV(0x20 + t)
        float sv[BN];
#pragma unroll
        for (int cch = 0; cch < BN / 32; ++cch) {
          uint32_t r[32];
          tmem_ld_x32(tm_s + cch * 32, r);
#pragma unroll
          for (int i = 0; i < 32; ++i) sv[cch * 32 + i] = __uint_as_float(r[i]);
        }
        tc_wait_ld();
        I2V_TRACE_EV(0x30 + t)
        const int valid = P.skv - j * BN;  // columns >= valid are padding

        if (!two_seg) {
          const bool full = valid >= BN;
          if (!full) {  // ragged last tile only (warp-uniform)
#pragma unroll
            for (int i = 0; i < BN; ++i)
              if (i >= valid) sv[i] = -INFINITY;
          }
          // P = 2^(c*s - mref) -> bf16 -> TMEM (aliasing S); returns the row sum.  `emu`: part of the exponentials on
          // the FMA pipe (never for ragged tiles: their -inf padding must go through MUFU, ex2(-inf) = 0).
          auto exp_pass = [&](float mref, bool emu) -> float {
            const uint64_t c2 = f2_pack(c, c);
            const uint64_t nm2 = f2_pack(-mref, -mref);
            uint64_t ls0 = 0ull, ls1 = 0ull;  // two packed partial row sums
#pragma unroll
            for (int cch = 0; cch < BN / 32; ++cch) {
              uint32_t pk[16];
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                const uint64_t x = f2_fma(f2_pack(sv[cch * 32 + 2 * i], sv[cch * 32 + 2 * i + 1]), c2, nm2);
                uint64_t p2;
                if ((i & 7) < Cfg::EMU && emu) {
                  p2 = ex2_emulated_pair(x);
                } else {
                  float x0, x1;
                  f2_unpack(x, x0, x1);
                  p2 = f2_pack(ex2_approx(x0), ex2_approx(x1));
                }
                if (i & 1) ls1 = f2_add(ls1, p2); else ls0 = f2_add(ls0, p2);
                float p0, p1;
                f2_unpack(p2, p0, p1);
                pk[i] = pack_bf16x2(p0, p1);
              }
              tmem_st_x16(tm_s + cch * 16, pk);
            }
            float a0, a1;
            f2_unpack(f2_add(ls0, ls1), a0, a1);
            return a0 + a1;
          };
          // multiply the O accumulator row by alpha (PV of tile j-1 has retired: s_full(j) was committed after it)
          auto rescale_o = [&](float alpha) {
#pragma unroll
            for (int cch = 0; cch < DK / 16; ++cch) {
              uint32_t r[16];
              tmem_ld_x16(tm_o + cch * 16, r);
              tc_wait_ld();
#pragma unroll
              for (int i = 0; i < 16; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) * alpha);
              tmem_st_x16(tm_o + cch * 16, r);
            }
            tc_wait_st();
          };
          // row max: 4 independent FMNMX3 chains
          float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
          for (int i = 0; i < BN; i += 8) {
            mx0 = fmax3(mx0, sv[i + 0], sv[i + 1]);
            mx1 = fmax3(mx1, sv[i + 2], sv[i + 3]);
            mx2 = fmax3(mx2, sv[i + 4], sv[i + 5]);
            mx3 = fmax3(mx3, sv[i + 6], sv[i + 7]);
          }
          const float mx = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3)) * c;
          float sum;
          if (j > 0 && full) {
            // Optimistic step: the exponentials start right away against the running reference max, the row max of
            // this tile is computed alongside (independent instruction streams: MUFU and FMNMX3 interleave).  Only
            // if some row's max outgrew the reference by more than the lazy-rescale threshold is the step redone.
            sum = exp_pass(m_ref[tt], true);
            const bool need = mx > m_ref[tt] + kRescaleThreshold;
            if (__any_sync(0xffffffffu, need)) {
              float alpha = 1.f;
              if (need) {
                alpha = ex2_approx(m_ref[tt] - mx);
                m_ref[tt] = mx;
              }
              rescale_o(alpha);
              l[tt] *= alpha;
              sum = exp_pass(m_ref[tt], true);
            }
          } else {
            // first tile (no reference yet) and ragged last tile: max first, then the exponentials
            float alpha = 1.f;
            bool need = false;
            if (mx > m_ref[tt] + kRescaleThreshold || m_ref[tt] == -INFINITY) {
              alpha = (m_ref[tt] == -INFINITY) ? 0.f : ex2_approx(m_ref[tt] - mx);
              m_ref[tt] = mx;
              need = true;
            }
            if (j > 0 && __any_sync(0xffffffffu, need)) rescale_o(alpha);
            l[tt] *= alpha;
            sum = exp_pass(m_ref[tt], full);
          }
          I2V_TRACE_EV(0x40 + t)
          l[tt] += sum;
        } else {
          // two-segment softmax over a single KV tile: [0,split) and [split,valid); probabilities are
          // normalised here so the accumulator needs no final division.
          const int split = P.seg_split;
          float m1 = -INFINITY, m2 = -INFINITY;
#pragma unroll
          for (int i = 0; i < BN; ++i) {
            if (i < split) m1 = fmaxf(m1, sv[i]);
            else if (i < valid) m2 = fmaxf(m2, sv[i]);
          }
          m1 *= c; m2 *= c;
          float l1 = 0.f, l2 = 0.f;
#pragma unroll
          for (int i = 0; i < BN; ++i) {
            float e = 0.f;
            if (i < split) { e = ex2_approx(fmaf(sv[i], c, -m1)); l1 += e; }
            else if (i < valid) { e = ex2_approx(fmaf(sv[i], c, -m2)); l2 += e; }
            sv[i] = e;
          }
          const float w1 = (l1 > 0.f) ? 1.f / l1 : 0.f;
          const float w2 = (l2 > 0.f) ? P.seg_scale / l2 : 0.f;
#pragma unroll
          for (int cch = 0; cch < BN / 32; ++cch) {
            uint32_t pk[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const int i0 = cch * 32 + 2 * i;
              pk[i] = pack_bf16x2(sv[i0] * (i0 < split ? w1 : w2), sv[i0 + 1] * (i0 + 1 < split ? w1 : w2));
            }
            tmem_st_x16(tm_s + cch * 16, pk);
          }
          l[tt] = 1.f;
        }
        I2V_TRACE_EV(0x50 + t)
        tc_wait_st();
        tc_fence_before();
        mbar_arrive(bar_p_full + t);
        I2V_TRACE_EV(0x60 + t)
      }
    }

    // ---- epilogue: O / l -> bf16 -> global ----
#pragma unroll
    for (int tt = 0; tt < TPW; ++tt) {
      const int t = tt * 2 + wg;
      if (t >= ntiles) continue;
      const uint32_t tm_o = tmem_base + Cfg::TMEM_O0 + t * DK + lane_addr;
      mbar_wait(bar_o_full + t, 0);
      tc_fence_after();
      const float inv_l = 1.f / l[tt];
      const int qrow = q0 + t * 128 + row;
      __nv_bfloat16* orow = prob.o + (long long)b * prob.o_sb + (long long)qrow * prob.o_ss + (long long)h * prob.o_sh;
#pragma unroll
      for (int cch = 0; cch < DK / 16; ++cch) {
        uint32_t r[16];
        tmem_ld_x16(tm_o + cch * 16, r);
        tc_wait_ld();
        if (qrow < P.sq) {
#pragma unroll
          for (int g8 = 0; g8 < 2; ++g8) {
            const int col = cch * 16 + g8 * 8;
            if (col < P.d) {  // d is a multiple of 8
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
  if (warp == 9) {
    tc_fence_after();
    tmem_dealloc<Cfg::TMEM_ALLOC>(tmem_base);
  }
}

}  // namespace i2v
