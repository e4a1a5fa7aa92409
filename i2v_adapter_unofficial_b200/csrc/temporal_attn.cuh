// AnimateDiff motion-module temporal self-attention (K4): attention over F <= 32 frames at every spatial
// position and head.  Reference call sites: diffusers TransformerTemporalModel -> BasicTransformerBlock.attn1/attn2
// (AttnProcessor2_0), constructed at src/models/unet_motion_cross_frame_attn.py:232-244 and called at :323-326.
//
// The op moves 4*F*C*2 bytes per position for 4*F*F*C flops (8 flop/B at F=16): it is HBM-bound, so the kernel is
// organised around the memory stream, not the tensor pipe:
//   * persistent CTAs; a producer warp streams whole (F x HG*d) row slabs of Q, K, V for one position with
//     1-D bulk TMA copies (full 128-B lines) into a multi-stage shared-memory ring (row pitch padded by 16 B so
//     ldmatrix is bank-conflict free);
//   * 8 consumer warps, one (or two, d > 80) per head, keep the F x F scores, the softmax and the F x d output
//     in registers: mma.sync.m16n8k16 bf16 (F = 16 is exactly one M tile; tcgen05 needs M >= 64 and would waste
//     8x the tensor work plus a TMEM round trip), row max/sum via 2 warp shuffles;
//   * outputs go straight from the accumulator fragments to global memory (16-B runs per quad).
#pragma once
#include "ptx_sm100.cuh"

namespace i2v {

struct TemporalParams {
  const __nv_bfloat16* q;
  const __nv_bfloat16* k;
  const __nv_bfloat16* v;
  __nv_bfloat16* o;
  long long q_sp, q_sf;  // element strides: position, frame (channels contiguous, head h at column h*d)
  long long k_sp, k_sf;
  long long v_sp, v_sf;
  long long o_sp, o_sf;
  int n_pos, frames, heads, d;
  float scale_log2e;
};

constexpr int kTemporalConsumerWarps = 8;
constexpr int kTemporalThreads = (kTemporalConsumerWarps + 1) * 32;

template <int D, int HG, int FT>
struct TemporalCfg {
  static constexpr int ROWS = FT * 16;
  static constexpr int ROW_BYTES = HG * D * 2;
  static constexpr int PITCH = ROW_BYTES + 16;
  static constexpr int SLAB_BYTES = ROWS * PITCH;
  static constexpr int STAGE_BYTES = 3 * SLAB_BYTES;
  static constexpr int WPH = kTemporalConsumerWarps / HG;  // warps per head
  static constexpr int NT = D / 8;                         // 8-wide output column tiles per head
  static constexpr int NTW = NT / WPH;                     // per warp
  static_assert(D % 8 == 0, "head dim must be a multiple of 8");
  static_assert(NT % WPH == 0, "output tiles must split evenly across the warps of a head");
  static_assert(kTemporalConsumerWarps % HG == 0, "HG");
};

__device__ __forceinline__ void bulk_copy_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               :
               : "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(gsrc)), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}
__device__ __forceinline__ void ldsm_x2(uint32_t addr, uint32_t& r0, uint32_t& r1) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0,%1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x2_trans(uint32_t addr, uint32_t& r0, uint32_t& r1) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(addr));
}
__device__ __forceinline__ void mma_m16n8k16(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                             uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void mma_m16n8k8(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t b0) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a0), "r"(a1), "r"(b0));
}

template <int D, int HG, int FT>
__global__ void __launch_bounds__(kTemporalThreads, 1) temporal_attn_kernel(const TemporalParams P, int nstages) {
  using Cfg = TemporalCfg<D, HG, FT>;
  constexpr int PITCH = Cfg::PITCH, ROWS = Cfg::ROWS, KT = 2 * FT /* 8-key tiles */;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
  uint64_t* bar_full = reinterpret_cast<uint64_t*>(smem);  // [nstages]
  uint64_t* bar_empty = bar_full + 8;                      // [nstages]
  uint8_t* stage0 = smem + 128;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int head_groups = P.heads / HG;
  const long long units = (long long)P.n_pos * head_groups;
  const int F = P.frames;

  if (threadIdx.x == 0) {
    for (int s = 0; s < nstages; ++s) {
      mbar_init(bar_full + s, 1);
      mbar_init(bar_empty + s, kTemporalConsumerWarps);
    }
    mbar_fence_init();
  }
  // rows >= F of every slab are never written by the copies: zero them once (V padding must not be NaN)
  if (F < ROWS) {
    const int pad_rows = ROWS - F;
    const int words_per_row = PITCH / 4;
    const int total = nstages * 3 * pad_rows * words_per_row;
    for (int i = threadIdx.x; i < total; i += blockDim.x) {
      const int w = i % words_per_row;
      int r = i / words_per_row;
      const int pr = r % pad_rows; r /= pad_rows;
      const int slab = r % 3; const int s = r / 3;
      reinterpret_cast<uint32_t*>(stage0 + (size_t)s * Cfg::STAGE_BYTES + slab * Cfg::SLAB_BYTES +
                                  (F + pr) * PITCH)[w] = 0u;
    }
    fence_proxy_async_smem();
  }
  __syncthreads();

  if (warp == kTemporalConsumerWarps) {
    // =========================== producer warp ===========================
    int it = 0;
    for (long long u = blockIdx.x; u < units; u += gridDim.x, ++it) {
      const int s = it % nstages;
      const uint32_t ph = (it / nstages) & 1;
      const long long pos = u / head_groups;
      const int col0 = (int)(u % head_groups) * HG * D;
      mbar_wait(bar_empty + s, ph ^ 1);
      if (lane == 0) mbar_arrive_expect_tx(bar_full + s, 3u * F * Cfg::ROW_BYTES);
      __syncwarp();
      uint8_t* st = stage0 + (size_t)s * Cfg::STAGE_BYTES;
      for (int r = lane; r < 3 * F; r += 32) {
        const int slab = r / F, f = r - slab * F;
        const __nv_bfloat16* src = slab == 0 ? P.q + pos * P.q_sp + f * P.q_sf
                                 : slab == 1 ? P.k + pos * P.k_sp + f * P.k_sf
                                             : P.v + pos * P.v_sp + f * P.v_sf;
        bulk_copy_g2s(st + slab * Cfg::SLAB_BYTES + f * PITCH, src + col0, Cfg::ROW_BYTES, bar_full + s);
      }
    }
  } else {
    // =========================== consumer warps ===========================
    const int hh = warp / Cfg::WPH;    // head within the unit
    const int part = warp % Cfg::WPH;  // which share of the output columns
    const int g = lane >> 2, tq = lane & 3;
    const float c = P.scale_log2e;
    int it = 0;
    for (long long u = blockIdx.x; u < units; u += gridDim.x, ++it) {
      const int s = it % nstages;
      const uint32_t ph = (it / nstages) & 1;
      const long long pos = u / head_groups;
      const int col0 = (int)(u % head_groups) * HG * D;
      mbar_wait(bar_full + s, ph);
      const uint32_t sq = smem_u32(stage0 + (size_t)s * Cfg::STAGE_BYTES) + hh * D * 2;
      const uint32_t sk = sq + Cfg::SLAB_BYTES;
      const uint32_t sv = sk + Cfg::SLAB_BYTES;

      // ---- S = Q K^T : FT m-tiles x KT key tiles ----
      float sc[FT][KT][4];
#pragma unroll
      for (int mt = 0; mt < FT; ++mt)
#pragma unroll
        for (int nt = 0; nt < KT; ++nt)
#pragma unroll
          for (int i = 0; i < 4; ++i) sc[mt][nt][i] = 0.f;

#pragma unroll
      for (int kk = 0; kk < D / 16; ++kk) {
        uint32_t a[FT][4];
#pragma unroll
        for (int mt = 0; mt < FT; ++mt)
          ldsm_x4(sq + (mt * 16 + (lane & 15)) * PITCH + (kk * 16 + (lane >> 4) * 8) * 2, a[mt][0], a[mt][1],
                  a[mt][2], a[mt][3]);
#pragma unroll
        for (int np = 0; np < FT; ++np) {  // pairs of key tiles (16 keys)
          uint32_t b0, b1, b2, b3;
          const int mi = lane >> 3;
          ldsm_x4(sk + (np * 16 + (mi >> 1) * 8 + (lane & 7)) * PITCH + (kk * 16 + (mi & 1) * 8) * 2, b0, b1, b2, b3);
#pragma unroll
          for (int mt = 0; mt < FT; ++mt) {
            mma_m16n8k16(sc[mt][2 * np], a[mt][0], a[mt][1], a[mt][2], a[mt][3], b0, b1);
            mma_m16n8k16(sc[mt][2 * np + 1], a[mt][0], a[mt][1], a[mt][2], a[mt][3], b2, b3);
          }
        }
      }
      if constexpr (D % 16 == 8) {  // k = 8 tail of the head dim
        constexpr int kcol = (D / 16) * 16;
        uint32_t a[FT][2];
#pragma unroll
        for (int mt = 0; mt < FT; ++mt)
          ldsm_x2(sq + (mt * 16 + (lane & 15)) * PITCH + kcol * 2, a[mt][0], a[mt][1]);
#pragma unroll
        for (int np = 0; np < FT; ++np) {
          uint32_t b0, b1;
          ldsm_x2(sk + (np * 16 + (lane & 15)) * PITCH + kcol * 2, b0, b1);
#pragma unroll
          for (int mt = 0; mt < FT; ++mt) {
            mma_m16n8k8(sc[mt][2 * np], a[mt][0], a[mt][1], b0);
            mma_m16n8k8(sc[mt][2 * np + 1], a[mt][0], a[mt][1], b1);
          }
        }
      }

      // ---- softmax over keys (row g and g+8 of each m-tile live in this quad) ----
      uint32_t pa[FT][KT][2];  // P as bf16 A fragments
      float inv_l[FT][2];
#pragma unroll
      for (int mt = 0; mt < FT; ++mt) {
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          float mx = -INFINITY;
#pragma unroll
          for (int nt = 0; nt < KT; ++nt)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const int key = nt * 8 + 2 * tq + e;
              float x = sc[mt][nt][half * 2 + e] * c;
              if (key >= F) x = -INFINITY;
              sc[mt][nt][half * 2 + e] = x;
              mx = fmaxf(mx, x);
            }
          mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
          mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
          float sum = 0.f;
#pragma unroll
          for (int nt = 0; nt < KT; ++nt) {
            const float p0 = ex2_approx(sc[mt][nt][half * 2 + 0] - mx);
            const float p1 = ex2_approx(sc[mt][nt][half * 2 + 1] - mx);
            sum += p0 + p1;
            pa[mt][nt][half] = pack_bf16x2(p0, p1);
          }
          sum += __shfl_xor_sync(0xffffffffu, sum, 1);
          sum += __shfl_xor_sync(0xffffffffu, sum, 2);
          inv_l[mt][half] = 1.f / sum;
        }
      }

      // ---- O = P V for this warp's share of the head's columns, stored straight to global ----
      const int ncol0 = part * Cfg::NTW * 8;
#pragma unroll
      for (int nt = 0; nt < Cfg::NTW; ++nt) {
        float oc[FT][4];
#pragma unroll
        for (int mt = 0; mt < FT; ++mt)
#pragma unroll
          for (int i = 0; i < 4; ++i) oc[mt][i] = 0.f;
#pragma unroll
        for (int kp = 0; kp < FT; ++kp) {  // 16 keys per step
          uint32_t b0, b1;
          ldsm_x2_trans(sv + (kp * 16 + (lane & 15)) * PITCH + (ncol0 + nt * 8) * 2, b0, b1);
#pragma unroll
          for (int mt = 0; mt < FT; ++mt)
            mma_m16n8k16(oc[mt], pa[mt][2 * kp][0], pa[mt][2 * kp][1], pa[mt][2 * kp + 1][0], pa[mt][2 * kp + 1][1],
                         b0, b1);
        }
#pragma unroll
        for (int mt = 0; mt < FT; ++mt)
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            const int f = mt * 16 + half * 8 + g;
            if (f < F) {
              const uint32_t v = pack_bf16x2(oc[mt][half * 2] * inv_l[mt][half], oc[mt][half * 2 + 1] * inv_l[mt][half]);
              __nv_bfloat16* dst = P.o + pos * P.o_sp + (long long)f * P.o_sf + col0 + hh * D + ncol0 + nt * 8 + 2 * tq;
              *reinterpret_cast<uint32_t*>(dst) = v;
            }
          }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_empty + s);
    }
  }
}

}  // namespace i2v
