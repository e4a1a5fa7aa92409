// Feed-forward input projection fused with GEGLU on sm_100a:
//
//   y[r, c] = (x W_h^T + b_h)[r, c] * gelu((x W_g^T + b_g)[r, c])        c in [0, N),  N = 4 * dim
//
// for the GEGLU feed-forward of the spatial and temporal transformer blocks (src/modules/i2v_adapter.py:535-561 ->
// diffusers FeedForward(activation_fn="geglu"): net[0].proj is one Linear dim -> 8 dim whose output is chunked into
// (hidden, gate)).  The unfused path writes the [rows, 8 dim] projection and reads it back (1 GB per call at SD1.5
// level 0, ~0.22 ms of pure HBM time on top of the GEMM); here a CTA's accumulator tile holds 128 hidden columns and the
// matching 128 gate columns side by side in TMEM, so the epilogue produces the product directly and only [rows, 4 dim]
// is written.
//
// Persistent, warp-specialised: 1 TMA warp (x tile 128 x 64 and two 128 x 64 weight boxes -- hidden rows n0.. and gate
// rows N + n0.. of the same nn.Linear weight, stacked into one 256-row K-major operand -- through a 4-stage ring), 1
// MMA-issuing warp (tcgen05.mma M = 128, N = 256, K = 16; fp32 accumulators double-buffered in TMEM: 2 x 256
// columns), 8 epilogue warps (lane quarter x column half) that overlap tile i's GELU with tile i+1's MMAs.
// Tiles are walked n-fastest, so the CTAs running at the same time share x row blocks through L2.
#pragma once
#include <cuda.h>
#include "ptx_sm100.cuh"
#include "norm_layout.cuh"   // gelu_erf, bf16 helpers

namespace i2v {

struct FfGegluParams {
  CUtensorMap tm_x;            // x [rows, K] bf16: dims (K, rows), box (64, 128), 128B swizzle
  CUtensorMap tm_w;            // W [2N, K] bf16:   dims (K, 2N),  box (64, 128)
  const __nv_bfloat16* bias;   // [2N] or nullptr
  __nv_bfloat16* out;          // [rows, ld]
  long long rows;
  int N, K, ld;
  int m_tiles, n_tiles;
};

constexpr int kFfStages = 4;
constexpr int kFfThreads = 320;                    // 8 epilogue warps, TMA warp, MMA warp
constexpr int kFfABytes = 128 * 128;               // 128 rows x 64 bf16
constexpr int kFfBBytes = 256 * 128;               // 256 rows x 64 bf16
constexpr int kFfStageBytes = kFfABytes + kFfBBytes;
constexpr int kFfSmemBytes = kFfStages * kFfStageBytes + 256 + 1024;

__global__ void __launch_bounds__(kFfThreads, 1) ff_geglu_gemm_kernel(const __grid_constant__ FfGegluParams P) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kFfStages * kFfStageBytes);
  uint64_t* bar_full = bars;                       // [stages]  TMA -> MMA
  uint64_t* bar_empty = bars + kFfStages;          // [stages]  MMA -> TMA
  uint64_t* bar_acc_full = bars + 2 * kFfStages;   // [2]       MMA -> epilogue
  uint64_t* bar_acc_empty = bar_acc_full + 2;      // [2]       epilogue -> MMA (8 arrivals: one per warp)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_acc_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int kTmaWarp = 8, kMmaWarp = 9;
  const int kblocks = P.K / 64;
  const long long tiles = (long long)P.m_tiles * P.n_tiles;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kFfStages; ++s) {
      mbar_init(bar_full + s, 1);
      mbar_init(bar_empty + s, 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(bar_acc_full + b, 1);
      mbar_init(bar_acc_empty + b, 8);
    }
    mbar_fence_init();
  }
  if (warp == kMmaWarp) tmem_alloc<512>(tmem_slot);
  if (warp == kTmaWarp && lane == 0) {
    tma_prefetch_desc(&P.tm_x);
    tma_prefetch_desc(&P.tm_w);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == kTmaWarp) {
    if (lane == 0) {
      uint32_t g = 0;
      for (long long t = blockIdx.x; t < tiles; t += gridDim.x) {
        const int nt = (int)(t % P.n_tiles);
        const int m0 = (int)(t / P.n_tiles) * 128, n0 = nt * 128;
        for (int kb = 0; kb < kblocks; ++kb, ++g) {
          const int s = g % kFfStages;
          mbar_wait(bar_empty + s, ((g / kFfStages) & 1) ^ 1);
          uint8_t* a = smem + s * kFfStageBytes;
          mbar_arrive_expect_tx(bar_full + s, kFfStageBytes);
          tma_load_2d(a, &P.tm_x, bar_full + s, kb * 64, m0, kEvictNormal);
          tma_load_2d(a + kFfABytes, &P.tm_w, bar_full + s, kb * 64, n0, kEvictLast);
          tma_load_2d(a + kFfABytes + kFfABytes, &P.tm_w, bar_full + s, kb * 64, P.N + n0, kEvictLast);
        }
      }
    }
  } else if (warp == kMmaWarp) {
    constexpr uint32_t idesc = make_idesc_bf16(128, 256, 0, 0);
    const uint64_t desc0 = make_smem_desc_sw128(0, 16, 1024);
    uint32_t g = 0;
    int i = 0;
    for (long long t = blockIdx.x; t < tiles; t += gridDim.x, ++i) {
      const int b = i & 1;
      mbar_wait(bar_acc_empty + b, ((i >> 1) & 1) ^ 1);   // the epilogue has drained this accumulator buffer
      tc_fence_after();
      for (int kb = 0; kb < kblocks; ++kb, ++g) {
        const int s = g % kFfStages;
        mbar_wait(bar_full + s, (g / kFfStages) & 1);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t aa = smem_u32(smem + s * kFfStageBytes) >> 4;
          const uint32_t ba = aa + (kFfABytes >> 4);
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {   // 32 bytes per 16-column k-step inside the 128-byte swizzle row
            const uint64_t da = desc0 | (uint64_t)((aa + kk * 2) & 0x3FFF);
            const uint64_t db = desc0 | (uint64_t)((ba + kk * 2) & 0x3FFF);
            umma_ss(tmem_base + b * 256, da, db, idesc, (kb > 0 || kk > 0) ? 1u : 0u);
          }
          tc_commit(bar_empty + s);
          if (kb == kblocks - 1) tc_commit(bar_acc_full + b);
        }
        __syncwarp();
      }
    }
  } else {
    // =========================== epilogue: warp = (lane quarter, column half) ===========================
    const int quarter = warp & 3, half = warp >> 2;
    const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
    int i = 0;
    for (long long t = blockIdx.x; t < tiles; t += gridDim.x, ++i) {
      const int b = i & 1;
      const int nt = (int)(t % P.n_tiles);
      const long long row = (t / P.n_tiles) * 128 + quarter * 32 + lane;
      const int n0 = nt * 128 + half * 64;
      mbar_wait(bar_acc_full + b, (i >> 1) & 1);
      tc_fence_after();
      const uint32_t tm = tmem_base + lane_addr + b * 256 + half * 64;
      __nv_bfloat16* orow = P.out + row * P.ld + n0;
      // chunks of 16 columns, software-pipelined: the TMEM loads of chunk ch + 1 are in flight under the GELUs of ch
      uint32_t h[2][16], gt[2][16];
      tmem_ld_x16(tm, h[0]);
      tmem_ld_x16(tm + 128, gt[0]);
#pragma unroll
      for (int ch = 0; ch < 4; ++ch) {
        uint32_t bh[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u}, bg[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
        if (P.bias) {
          const uint4* ph = reinterpret_cast<const uint4*>(P.bias + n0 + ch * 16);
          const uint4* pg = reinterpret_cast<const uint4*>(P.bias + P.N + n0 + ch * 16);
          const uint4 h0 = __ldg(ph), h1 = __ldg(ph + 1), g0 = __ldg(pg), g1 = __ldg(pg + 1);
          bh[0] = h0.x; bh[1] = h0.y; bh[2] = h0.z; bh[3] = h0.w; bh[4] = h1.x; bh[5] = h1.y; bh[6] = h1.z; bh[7] = h1.w;
          bg[0] = g0.x; bg[1] = g0.y; bg[2] = g0.z; bg[3] = g0.w; bg[4] = g1.x; bg[5] = g1.y; bg[6] = g1.z; bg[7] = g1.w;
        }
        tc_wait_ld();
        if (ch + 1 < 4) {
          tmem_ld_x16(tm + (ch + 1) * 16, h[(ch + 1) & 1]);
          tmem_ld_x16(tm + 128 + (ch + 1) * 16, gt[(ch + 1) & 1]);
        }
        const uint32_t* hc = h[ch & 1];
        const uint32_t* gc = gt[ch & 1];
        uint32_t o[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const float h0 = __uint_as_float(hc[2 * k]) + bf16_lo(bh[k]), h1 = __uint_as_float(hc[2 * k + 1]) + bf16_hi(bh[k]);
          const float g0 = __uint_as_float(gc[2 * k]) + bf16_lo(bg[k]), g1 = __uint_as_float(gc[2 * k + 1]) + bf16_hi(bg[k]);
          o[k] = bf16_pack(h0 * gelu_erf_fast(g0), h1 * gelu_erf_fast(g1));
        }
        if (row < P.rows) {
          uint4* dst = reinterpret_cast<uint4*>(orow + ch * 16);
          dst[0] = make_uint4(o[0], o[1], o[2], o[3]);
          dst[1] = make_uint4(o[4], o[5], o[6], o[7]);
        }
      }
      if (P.ld > P.N && nt == 0 && half == 0 && row < P.rows)   // ones column for the next GEMM's deferred bias
        *reinterpret_cast<uint4*>(P.out + row * P.ld + P.N) = make_uint4(0x00003F80u, 0u, 0u, 0u);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_acc_empty + b);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

}  // namespace i2v
