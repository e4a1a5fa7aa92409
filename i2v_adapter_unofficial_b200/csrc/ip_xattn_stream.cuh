// IP-Adapter decoupled cross-attention (K3) as a streaming kernel: text (77) + image-prompt (4) keys, i.e. at most
// 96 keys per (video, head), against B*S query rows.  Reference: diffusers IPAdapterAttnProcessor2_0 installed at
// src/models/unet_motion_cross_frame_attn.py:1264-1279 (two SDPA calls + scaled add), tokens built at :1346-1355.
//
// The operator reads Q once and writes O once (K/V are a few KB per head and identical for all frames of a video):
// 4 flop per byte per key -> HBM-bound, like the temporal attention, so it is built the same way and judged on GB/s:
//   * persistent CTAs; K and V of a group of heads of one video (320 columns: 8 / 4 / 2 heads at d = 40 / 80 / 160) stay
//     in shared memory (2 x 96 rows) and are reloaded only when the CTA's work moves to the next (video, head group);
//   * a producer warp streams slabs of Q rows (640-byte runs: full 128-byte lines) through a shared-memory ring with
//     bulk async copies;
//   * 8 consumer warps (head x row group) keep the 16 x 96 scores of an m-tile, both softmaxes (text segment / image segment,
//     each normalised on its own, the image one scaled by ip_scale) and the 16 x d output in registers
//     (mma.sync.m16n8k16; a tcgen05 tile would spend its 128 x 128 exponentials on 81 keys and wait on a TMEM round
//     trip per 128 rows -- the first implementation did, at 1/13 of the HBM roofline);
//   * outputs go straight from the accumulator fragments to global memory.
#pragma once
#include "ptx_sm100.cuh"
#include "temporal_attn.cuh"  // ldmatrix / mma.sync / bulk-copy wrappers

namespace i2v {

struct IpStreamParams {
  const __nv_bfloat16* q;   // [batch, sq, H*d]      rows contiguous (head h at column h*d)
  const __nv_bfloat16* k;   // [batch/kv_group, nk, H*d]
  const __nv_bfloat16* v;
  __nv_bfloat16* o;         // [batch, sq, H*d]
  long long q_sb, q_ss, k_sb, k_ss, v_sb, v_ss, o_sb, o_ss;   // element strides: batch, sequence
  int batch, sq, heads, nk, n_txt, kv_group;
  float scale_log2e, ip_scale;
};

constexpr int kIpWarps = 8;            // consumer warps
constexpr int kIpThreads = (kIpWarps + 1) * 32;
constexpr int kIpKeys = 96;            // key rows held in shared memory (>= nk)

// HG heads share a unit, the 8 consumer warps split into HG heads x 8/HG row groups of MT m-tiles each; MINB CTAs are
// co-resident per SM (the kernel is latency-bound at 2 consumer warps per sub-partition, so more resident warps matter
// more than longer rows).
template <int D, int HG, int MT, int NSTG, int MINB = 1>
struct IpStreamCfg {
  static constexpr int ROW_BYTES = HG * D * 2;
  static constexpr int PITCH = ROW_BYTES + 16;   // ldmatrix rows land in different banks
  static constexpr int KV_BYTES = kIpKeys * PITCH;
  static constexpr int ROWS = 16 * MT * (kIpWarps / HG);   // query rows per unit
  static constexpr int Q_BYTES = ROWS * PITCH;
  static constexpr int SMEM_BYTES = 256 + 2 * KV_BYTES + NSTG * Q_BYTES + 128;
  static_assert(D % 8 == 0, "head dim must be a multiple of 8");
  static_assert(kIpWarps % HG == 0, "HG");
  static_assert((SMEM_BYTES + 1024) * MINB <= 228 * 1024, "smem budget");
};

// NTXT / NK > 0 fix the token counts at compile time (77 text + 4 image tokens is what the pipeline feeds): the key-tile
// classification below then folds away; 0 = read them from the launch parameters.
template <int D, int HG, int MT, int NSTG, int MINB, int NTXT = 0, int NK = 0>
__global__ void __launch_bounds__(kIpThreads, MINB) ip_xattn_stream_kernel(const IpStreamParams P) {
  using Cfg = IpStreamCfg<D, HG, MT, NSTG, MINB>;
  constexpr int PITCH = Cfg::PITCH, KT = kIpKeys / 8, KP = kIpKeys / 16, ROWS = Cfg::ROWS;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
  uint64_t* bar_full = reinterpret_cast<uint64_t*>(smem);   // [NSTG]  Q slab landed
  uint64_t* bar_empty = bar_full + NSTG;                     // [NSTG]  Q slab consumed (8 arrivals)
  uint64_t* bar_kv_full = bar_empty + NSTG;                  // [1]     K/V of the current (video, head group) landed
  uint64_t* bar_kv_free = bar_kv_full + 1;                   // [1]     every consumer is done with the old K/V
  uint8_t* sm_k = smem + 256;
  uint8_t* sm_v = sm_k + Cfg::KV_BYTES;
  uint8_t* sm_q = sm_v + Cfg::KV_BYTES;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row_blocks = (P.sq + ROWS - 1) / ROWS;
  const int head_groups = P.heads / HG;
  const int videos = P.batch / P.kv_group;
  // unit = (video, head group, frame of the video, row block), row block fastest: a CTA's consecutive units share K/V
  const long long units = (long long)videos * head_groups * P.kv_group * row_blocks;
  struct Unit { int b, rb, col0, kvid, bkv; };
  auto decode = [&](long long u) {
    Unit x;
    x.rb = (int)(u % row_blocks);  u /= row_blocks;
    const int f = (int)(u % P.kv_group);  u /= P.kv_group;
    const int hg = (int)(u % head_groups);
    x.bkv = (int)(u / head_groups);
    x.b = x.bkv * P.kv_group + f;
    x.col0 = hg * HG * D;
    x.kvid = x.bkv * head_groups + hg;
    return x;
  };

  if (threadIdx.x == 0) {
    for (int s = 0; s < NSTG; ++s) {
      mbar_init(bar_full + s, 1);
      mbar_init(bar_empty + s, kIpWarps);
    }
    mbar_init(bar_kv_full, 1);
    mbar_init(bar_kv_free, kIpWarps);
    mbar_fence_init();
  }
  // key rows >= nk are never written by the copies: zero K, V (V padding must not be NaN) and the Q ring once
  {
    const int words = (2 * Cfg::KV_BYTES + NSTG * Cfg::Q_BYTES) / 4;
    for (int i = threadIdx.x; i < words; i += blockDim.x) reinterpret_cast<uint32_t*>(sm_k)[i] = 0u;
    fence_proxy_async_smem();
  }
  __syncthreads();

  // contiguous range of units per CTA (K/V is reloaded only where the range crosses a (video, head group) boundary)
  const long long per = (units + gridDim.x - 1) / gridDim.x;
  const long long u0 = (long long)blockIdx.x * per, u1 = min(units, u0 + per);

  if (warp == kIpWarps) {
    // =========================== producer warp ===========================
    int it = 0, kv_loads = 0, cur_kv = -1;
    for (long long u = u0; u < u1; ++u, ++it) {
      const Unit x = decode(u);
      if (x.kvid != cur_kv) {
        if (kv_loads > 0) mbar_wait(bar_kv_free, (kv_loads - 1) & 1);   // consumers left the previous K/V
        if (lane == 0) mbar_arrive_expect_tx(bar_kv_full, 2u * P.nk * Cfg::ROW_BYTES);
        __syncwarp();
        for (int r = lane; r < 2 * P.nk; r += 32) {
          const int isv = r >= P.nk, key = isv ? r - P.nk : r;
          const __nv_bfloat16* src = isv ? P.v + (long long)x.bkv * P.v_sb + (long long)key * P.v_ss
                                         : P.k + (long long)x.bkv * P.k_sb + (long long)key * P.k_ss;
          bulk_copy_g2s((isv ? sm_v : sm_k) + key * PITCH, src + x.col0, Cfg::ROW_BYTES, bar_kv_full);
        }
        cur_kv = x.kvid;
        ++kv_loads;
      }
      const int s = it % NSTG;
      mbar_wait(bar_empty + s, ((it / NSTG) & 1) ^ 1);
      const int r0 = x.rb * ROWS, nrows = min(ROWS, P.sq - r0);
      if (lane == 0) mbar_arrive_expect_tx(bar_full + s, (uint32_t)nrows * Cfg::ROW_BYTES);
      __syncwarp();
      for (int r = lane; r < nrows; r += 32)
        bulk_copy_g2s(sm_q + (size_t)s * Cfg::Q_BYTES + r * PITCH,
                      P.q + (long long)x.b * P.q_sb + (long long)(r0 + r) * P.q_ss + x.col0, Cfg::ROW_BYTES,
                      bar_full + s);
    }
  } else {
    // =========================== consumer warps ===========================
    const int hh = warp % HG;          // head within the group
    const int mg = warp / HG;          // row group: m-tiles mg*MT .. mg*MT + MT-1 of the unit
    const int g = lane >> 2, tq = lane & 3;
    const float c = P.scale_log2e;
    const uint32_t sk = smem_u32(sm_k) + hh * D * 2;
    const uint32_t sv = smem_u32(sm_v) + hh * D * 2;
    int it = 0, kv_seen = 0, cur_kv = -1;
    for (long long u = u0; u < u1; ++u, ++it) {
      const Unit x = decode(u);
      if (x.kvid != cur_kv) {
        if (cur_kv >= 0) {   // done with the previous K/V
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_kv_free);
        }
        mbar_wait(bar_kv_full, kv_seen & 1);
        ++kv_seen;
        cur_kv = x.kvid;
      }
      const int s = it % NSTG;
      mbar_wait(bar_full + s, (it / NSTG) & 1);
      const uint32_t sq = smem_u32(sm_q + (size_t)s * Cfg::Q_BYTES) + hh * D * 2;
      const int r0 = x.rb * ROWS;
      const int b = x.b;

#pragma unroll 1
      for (int mi_ = 0; mi_ < MT; ++mi_) {
        const int mt = mg * MT + mi_;
        if (r0 + mt * 16 >= P.sq) break;
        // ---- S = Q K^T: one 16-row m-tile x KT key tiles ----
        float sc[KT][4];
#pragma unroll
        for (int nt = 0; nt < KT; ++nt)
#pragma unroll
          for (int i = 0; i < 4; ++i) sc[nt][i] = 0.f;
#pragma unroll
        for (int kk = 0; kk < D / 16; ++kk) {
          uint32_t a0, a1, a2, a3;
          ldsm_x4(sq + (mt * 16 + (lane & 15)) * PITCH + (kk * 16 + (lane >> 4) * 8) * 2, a0, a1, a2, a3);
#pragma unroll
          for (int np = 0; np < KP; ++np) {
            uint32_t b0, b1, b2, b3;
            const int mi = lane >> 3;
            ldsm_x4(sk + (np * 16 + (mi >> 1) * 8 + (lane & 7)) * PITCH + (kk * 16 + (mi & 1) * 8) * 2, b0, b1, b2, b3);
            mma_m16n8k16(sc[2 * np], a0, a1, a2, a3, b0, b1);
            mma_m16n8k16(sc[2 * np + 1], a0, a1, a2, a3, b2, b3);
          }
        }
        if constexpr (D % 16 == 8) {  // k = 8 tail of the head dim
          constexpr int kcol = (D / 16) * 16;
          uint32_t a0, a1;
          ldsm_x2(sq + (mt * 16 + (lane & 15)) * PITCH + kcol * 2, a0, a1);
#pragma unroll
          for (int np = 0; np < KP; ++np) {
            uint32_t b0, b1;
            ldsm_x2(sk + (np * 16 + (lane & 15)) * PITCH + kcol * 2, b0, b1);
            mma_m16n8k8(sc[2 * np], a0, a1, b0);
            mma_m16n8k8(sc[2 * np + 1], a0, a1, b1);
          }
        }

        // ---- two softmaxes per row (rows g and g+8 of the m-tile live in this quad) ----
        // Key tiles are classified once (warp-uniform, from the launch parameters): entirely text, entirely image,
        // entirely padding, or mixed -- only the mixed ones (two of twelve at 77 + 4 tokens) pay per-element tests.
        const int n_txt = NTXT > 0 ? NTXT : P.n_txt, nk = NK > 0 ? NK : P.nk;
        auto cls = [&](int nt) { return (nt * 8 + 8 <= n_txt) ? 0 : (nt * 8 >= nk) ? 3 : (nt * 8 >= n_txt && nt * 8 + 8 <= nk) ? 1 : 2; };
        uint32_t pa[KT][2];
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          float m1 = -INFINITY, m2 = -INFINITY;
#pragma unroll
          for (int nt = 0; nt < KT; ++nt) {
            const float x0 = sc[nt][half * 2], x1 = sc[nt][half * 2 + 1];
            const int k = cls(nt);
            if (k == 0) m1 = fmax3(m1, x0, x1);
            else if (k == 1) m2 = fmax3(m2, x0, x1);
            else if (k == 2) {
              const int key = nt * 8 + 2 * tq;
              if (key < n_txt) m1 = fmaxf(m1, x0); else if (key < nk) m2 = fmaxf(m2, x0);
              if (key + 1 < n_txt) m1 = fmaxf(m1, x1); else if (key + 1 < nk) m2 = fmaxf(m2, x1);
            }
          }
          m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 1));
          m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 2));
          m2 = fmaxf(m2, __shfl_xor_sync(0xffffffffu, m2, 1));
          m2 = fmaxf(m2, __shfl_xor_sync(0xffffffffu, m2, 2));
          const float mc1 = m1 * c, mc2 = m2 * c;   // scores are still unscaled: p = 2^(x*c - m*c)
          float l1 = 0.f, l2 = 0.f;
#pragma unroll
          for (int nt = 0; nt < KT; ++nt) {
            const float x0 = sc[nt][half * 2], x1 = sc[nt][half * 2 + 1];
            const int k = cls(nt);
            float p0 = 0.f, p1 = 0.f;
            if (k == 0) { p0 = ex2_approx(fmaf(x0, c, -mc1)); p1 = ex2_approx(fmaf(x1, c, -mc1)); l1 += p0 + p1; }
            else if (k == 1) { p0 = ex2_approx(fmaf(x0, c, -mc2)); p1 = ex2_approx(fmaf(x1, c, -mc2)); l2 += p0 + p1; }
            else if (k == 2) {
              const int key = nt * 8 + 2 * tq;
              if (key < n_txt) { p0 = ex2_approx(fmaf(x0, c, -mc1)); l1 += p0; }
              else if (key < nk) { p0 = ex2_approx(fmaf(x0, c, -mc2)); l2 += p0; }
              if (key + 1 < n_txt) { p1 = ex2_approx(fmaf(x1, c, -mc1)); l1 += p1; }
              else if (key + 1 < nk) { p1 = ex2_approx(fmaf(x1, c, -mc2)); l2 += p1; }
            }
            sc[nt][half * 2] = p0;
            sc[nt][half * 2 + 1] = p1;
          }
          l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
          l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
          l2 += __shfl_xor_sync(0xffffffffu, l2, 1);
          l2 += __shfl_xor_sync(0xffffffffu, l2, 2);
          const float w1 = l1 > 0.f ? rcp_approx(l1) : 0.f;
          const float w2 = l2 > 0.f ? P.ip_scale * rcp_approx(l2) : 0.f;
#pragma unroll
          for (int nt = 0; nt < KT; ++nt) {
            const int k = cls(nt);
            const int key = nt * 8 + 2 * tq;
            const float wa = k == 0 ? w1 : k == 1 ? w2 : (key < n_txt ? w1 : w2);
            const float wb = k == 0 ? w1 : k == 1 ? w2 : (key + 1 < n_txt ? w1 : w2);
            pa[nt][half] = pack_bf16x2(sc[nt][half * 2] * wa, sc[nt][half * 2 + 1] * wb);
          }
        }

        // ---- O = P V, stored straight from the accumulator fragments ----
        const int qrow = r0 + mt * 16 + g;
        __nv_bfloat16* obase = P.o + (long long)b * P.o_sb + x.col0 + hh * D + 2 * tq;
#pragma unroll
        for (int nt = 0; nt < D / 8; ++nt) {
          float oc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int kp = 0; kp < KP; ++kp) {
            uint32_t b0, b1;
            ldsm_x2_trans(sv + (kp * 16 + (lane & 15)) * PITCH + nt * 8 * 2, b0, b1);
            mma_m16n8k16(oc, pa[2 * kp][0], pa[2 * kp][1], pa[2 * kp + 1][0], pa[2 * kp + 1][1], b0, b1);
          }
          if (qrow < P.sq)
            *reinterpret_cast<uint32_t*>(obase + (long long)qrow * P.o_ss + nt * 8) = pack_bf16x2(oc[0], oc[1]);
          if (qrow + 8 < P.sq)
            *reinterpret_cast<uint32_t*>(obase + (long long)(qrow + 8) * P.o_ss + nt * 8) = pack_bf16x2(oc[2], oc[3]);
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_empty + s);
    }
  }
}

}  // namespace i2v
