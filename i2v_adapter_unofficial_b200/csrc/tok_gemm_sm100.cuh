// Token GEMM on sm_100a for the projections around the attention operators ("K5"):
//
//   out[r, n] = sum_k x[r, k] * W[n, k]  (+ bias[n])  (+ res[r, n])          x [rows, K], W [N, K] (nn.Linear layout)
//
// i.e. `F.linear(x, W, bias)` with the block's residual add riding in the epilogue -- the packed [Wq; Wk; Wv; Wq_x]
// projection, the stacked output projection [O_self | O_x] [Wo | Wo_x]^T + residual of
// src/modules/i2v_adapter.py:445-501, attn2's to_q / to_out (:514-533), the feed-forward's second Linear (:554-561)
// and the motion module's projections, which round 1 left on cuBLAS.
//
// Same machine as ff_geglu_gemm_sm100.cuh (whose main loop runs at the measured bf16 peak): persistent CTA pairs on the
// 2-SM MMA (tcgen05 cta_group::2, M = 256), each CTA loading its own 128-row x tile and HALF of the weight tile
// (TILE_N / 2 rows) per 64-wide k-block through a 6-stage TMA ring, fp32 accumulators double-buffered in TMEM
// (2 x 256 columns), 16 epilogue warps (lane quarter x quarter of the tile's columns).  TILE_N is a template
// parameter (128 .. 256 in steps of 32) chosen by the launcher so that N splits without a padded tile (N = 320 -> two
// tiles of 160, 960 -> five of 192).  Rows, N and K need not be multiples of the tile: TMA zero-fills out-of-bounds
// loads (K tails such as the feed-forward's 4 * dim + 8 included) and the epilogue masks rows / columns.
//
// Epilogue through shared memory, all global traffic as full lines: the residual tile [128 x TILE_N] is TMA-loaded into a
// staging tile (32-column sub-tiles, 64-byte swizzle) while the tile's MMAs run; the 16 epilogue warps add their
// accumulator share (+ bias) to it in place (two register passes of tcgen05.ld per warp) and one thread writes the
// staging tile back with tiled TMA stores, which clip rows / columns past the end of the matrix.  `res` may alias `out`.
#pragma once
#include <cuda.h>
#include "ptx_sm100.cuh"
#include "norm_layout.cuh"   // bf16 helpers

namespace i2v {

struct TokGemmParams {
  CUtensorMap tm_x;            // x [rows, K] bf16: dims (K, rows), box (64, 128), 128B swizzle
  CUtensorMap tm_w;            // W [N, K] bf16:    dims (K, N),    box (64, TILE_N / 2)
  CUtensorMap tm_out;          // out [rows, N] (pitch ld_out): dims (N, rows), box (32, 128), 64B swizzle
  CUtensorMap tm_res;          // res [rows, N] (pitch ld_res), same box; unused when has_res == 0
  const __nv_bfloat16* bias;   // [N] or nullptr
  int has_res;
  long long rows;
  int N, K;
  int m_pairs, n_tiles;        // 256-row pairs, TILE_N-column tiles
};

constexpr int kTgEpiWarps = 16;
constexpr int kTgThreads = (kTgEpiWarps + 2) * 32;   // epilogue warps, TMA warp, MMA warp
constexpr int kTgStages = 4;
constexpr int kTgABytes = 128 * 128;                 // 128 rows x 64 bf16
constexpr int kTgStageBytes = 2 * kTgABytes;         // x tile + (up to) 128 weight rows
constexpr int kTgSubBytes = 128 * 64;                // staging sub-tile: 128 rows x 32 bf16
// staging tiles: two for tiles of up to 160 columns (the residual of tile i + 1 lands while tile i is in its epilogue: the
// narrow-N projections are the ones with a residual and with few k-blocks per tile), one for wider tiles
__host__ __device__ constexpr int tg_nbuf(int tile_n) { return tile_n <= 160 ? 2 : 1; }
constexpr int kTgOutBytes = 10 * kTgSubBytes;        // 2 x 160 columns or 1 x 256 columns (80 KB)
constexpr int kTgSmemBytes = kTgStages * kTgStageBytes + kTgOutBytes + 256 + 1024;
static_assert(kTgSmemBytes <= 227 * 1024, "smem budget");

template <int TILE_N>
__global__ void __launch_bounds__(kTgThreads, 1) tok_gemm_kernel(const __grid_constant__ TokGemmParams P) {
  static_assert(TILE_N % 32 == 0 && TILE_N >= 64 && TILE_N <= 256, "tile width");
  constexpr int HALF_N = TILE_N / 2;                  // weight rows per CTA
  constexpr int CW = TILE_N / 4, NCH = CW / 8;        // columns per epilogue warp, 8-column chunks
  constexpr int W_BYTES = HALF_N * 128;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr int NSUB = TILE_N / 32;                   // staging sub-tiles
  constexpr int NBUF = tg_nbuf(TILE_N);
  static_assert(NBUF * NSUB * kTgSubBytes <= kTgOutBytes, "staging area");
  uint8_t* sm_out = smem + kTgStages * kTgStageBytes;  // [NSUB][128 rows][64 B], 64-byte swizzle
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm_out + kTgOutBytes);
  uint64_t* bar_full = bars;                          // [stages]  TMA -> MMA (leader, both CTAs' bytes)
  uint64_t* bar_empty = bars + kTgStages;             // [stages]  MMA -> TMA (both CTAs, multicast commit)
  uint64_t* bar_acc_full = bars + 2 * kTgStages;      // [2]       MMA -> epilogue (both CTAs)
  uint64_t* bar_acc_empty = bar_acc_full + 2;         // [2]       epilogue -> MMA (leader; both CTAs' warps arrive)
  uint64_t* bar_stage = bar_acc_empty + 2;            // [2]       staging tile i % NBUF is ready for the epilogue of tile i:
  //                                                                 its previous store has read it and, with a residual,
  //                                                                 the residual tile has landed in it
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_stage + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int kTmaWarp = kTgEpiWarps, kMmaWarp = kTgEpiWarps + 1;
  const int kblocks = (P.K + 63) / 64;
  const uint32_t rank = cluster_ctarank();
  const long long units = (long long)P.m_pairs * P.n_tiles;   // walked n-fastest: co-running pairs share x rows in L2
  const long long n_groups = gridDim.x / 2;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kTgStages; ++s) {
      mbar_init(bar_full + s, 1);
      mbar_init(bar_empty + s, 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(bar_acc_full + b, 1);
      mbar_init(bar_acc_empty + b, 2 * kTgEpiWarps);
    }
    mbar_init(bar_stage, 1);
    mbar_init(bar_stage + 1, 1);
    mbar_fence_init();
  }
  if (warp == kMmaWarp) tmem_alloc_pair<512>(tmem_slot);
  if (warp == kTmaWarp && lane == 0) {
    tma_prefetch_desc(&P.tm_x);
    tma_prefetch_desc(&P.tm_w);
    tma_prefetch_desc(&P.tm_out);
    if (P.has_res) tma_prefetch_desc(&P.tm_res);
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();   // the peer's barriers are initialised before anything arrives on them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == kTmaWarp) {
    if (lane == 0) {
      uint32_t g = 0;
      for (long long u = blockIdx.x / 2; u < units; u += n_groups) {
        const int nt = (int)(u % P.n_tiles);
        const int m0 = ((int)(u / P.n_tiles) * 2 + (int)rank) * 128;
        const int n0 = nt * TILE_N + (int)rank * HALF_N;
        for (int kb = 0; kb < kblocks; ++kb, ++g) {
          const int s = g % kTgStages;
          mbar_wait(bar_empty + s, ((g / kTgStages) & 1) ^ 1);
          uint8_t* a = smem + s * kTgStageBytes;
          if (rank == 0) mbar_arrive_expect_tx(bar_full + s, 2 * (kTgABytes + W_BYTES));
          tma_load_2d_pair(a, &P.tm_x, bar_full + s, kb * 64, m0, kEvictNormal);
          tma_load_2d_pair(a + kTgABytes, &P.tm_w, bar_full + s, kb * 64, n0, kEvictLast);
        }
      }
    }
  } else if (warp == kMmaWarp) {
    if (rank == 0) {   // the leader issues for the pair
      constexpr uint32_t idesc = make_idesc_bf16(256, TILE_N, 0, 0);
      const uint64_t desc0 = make_smem_desc_sw128(0, 16, 1024);
      uint32_t g = 0;
      int i = 0;
      for (long long u = blockIdx.x / 2; u < units; u += n_groups, ++i) {
        const int b = i & 1;
        mbar_wait(bar_acc_empty + b, ((i >> 1) & 1) ^ 1);   // the epilogue warps of both CTAs drained this buffer
        tc_fence_after();
        for (int kb = 0; kb < kblocks; ++kb, ++g) {
          const int s = g % kTgStages;
          mbar_wait(bar_full + s, (g / kTgStages) & 1);
          tc_fence_after();
          if (elect_one()) {
            const uint32_t aa = smem_u32(smem + s * kTgStageBytes) >> 4;
            const uint32_t ba = aa + (kTgABytes >> 4);
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {   // 32 bytes per 16-column k-step inside the 128-byte swizzle row
              const uint64_t da = desc0 | (uint64_t)((aa + kk * 2) & 0x3FFF);
              const uint64_t db = desc0 | (uint64_t)((ba + kk * 2) & 0x3FFF);
              umma_ss_pair(tmem_base + b * 256, da, db, idesc, (kb > 0 || kk > 0) ? 1u : 0u);
            }
            tc_commit_pair(bar_empty + s, (uint16_t)0b11);
            if (kb == kblocks - 1) tc_commit_pair(bar_acc_full + b, (uint16_t)0b11);
          }
          __syncwarp();
        }
      }
    }
  } else {
    // =========================== epilogue: warp = (lane quarter, quarter of the tile's columns) ===========================
    const int quarter = warp & 3, cg = warp >> 2;
    const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
    const int trow = quarter * 32 + lane;                 // row of the tile
    // hands the staging tile to the epilogue of the unit `u`: with a residual, its tile is TMA-loaded into it (rows /
    // columns past the end read as zero); without, the barrier just says "free"
    auto prepare_stage = [&](long long u, int buf) {
      if (u >= units) return;
      if (P.has_res) {
        const int n0 = (int)(u % P.n_tiles) * TILE_N;
        const int m0 = ((int)(u / P.n_tiles) * 2 + (int)rank) * 128;
        mbar_arrive_expect_tx(bar_stage + buf, NSUB * kTgSubBytes);
#pragma unroll
        for (int sb = 0; sb < NSUB; ++sb)
          tma_load_2d(sm_out + (buf * NSUB + sb) * kTgSubBytes, &P.tm_res, bar_stage + buf, n0 + sb * 32, m0, kEvictFirst);
      } else {
        mbar_arrive(bar_stage + buf);
      }
    };
    if (threadIdx.x == 0) {
#pragma unroll
      for (int k = 0; k < NBUF; ++k) prepare_stage(blockIdx.x / 2 + k * n_groups, k);
    }
    int i = 0;
    for (long long u = blockIdx.x / 2; u < units; u += n_groups, ++i) {
      const int b = i & 1;
      const int nt = (int)(u % P.n_tiles);
      const int m0 = ((int)(u / P.n_tiles) * 2 + (int)rank) * 128;
      const int c_first = cg * CW;                         // this warp's first column within the tile
      mbar_wait(bar_acc_full + b, (i >> 1) & 1);
      tc_fence_after();
      const int sbuf = i % NBUF;
      uint8_t* stage = sm_out + sbuf * NSUB * kTgSubBytes;
      mbar_wait(bar_stage + sbuf, (i / NBUF) & 1);
      // The warp's columns go through the registers in two passes (the full share would spill at 96 registers); the TMEM
      // buffer goes back to the MMA warp after the second pass's load.
      constexpr int PASS0 = (NCH + 1) / 2;
      const uint32_t tm = tmem_base + lane_addr + b * 256 + c_first;
#pragma unroll
      for (int pass = 0; pass < 2; ++pass) {
        const int c0 = pass == 0 ? 0 : PASS0;
        const int cn = pass == 0 ? PASS0 : NCH - PASS0;
        uint32_t acc[PASS0][8];
#pragma unroll
        for (int ch = 0; ch < PASS0; ++ch)
          if (ch < cn) tmem_ld_x8(tm + (c0 + ch) * 8, acc[ch]);
        tc_wait_ld();
        if (pass == 1) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(bar_acc_empty + b, 0u);   // the leader's barrier
        }
#pragma unroll
        for (int ch = 0; ch < PASS0; ++ch) {
          if (ch >= cn) break;
          const int col = c_first + (c0 + ch) * 8;         // column within the tile
          const int n = nt * TILE_N + col;
          const uint4 bv = (P.bias && n < P.N) ? *reinterpret_cast<const uint4*>(P.bias + n) : make_uint4(0u, 0u, 0u, 0u);
          // 64-byte swizzle: 16-byte chunk c of row r of a sub-tile lives at chunk position c ^ ((r >> 1) & 3)
          uint4* slot = reinterpret_cast<uint4*>(stage + (col >> 5) * kTgSubBytes + trow * 64 +
                                                 ((((col & 31) >> 3) ^ ((trow >> 1) & 3)) << 4));
          uint4 rv = make_uint4(0u, 0u, 0u, 0u);
          if (P.has_res) rv = *slot;
          const uint32_t rw[4] = {rv.x, rv.y, rv.z, rv.w};
          const uint32_t bw[4] = {bv.x, bv.y, bv.z, bv.w};
          uint32_t o[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            float lo = __uint_as_float(acc[ch][2 * k]) + bf16_lo(bw[k]);
            float hi = __uint_as_float(acc[ch][2 * k + 1]) + bf16_hi(bw[k]);
            if (P.has_res) {   // the reference rounds the Linear's output to bf16 before the residual add (two PyTorch ops)
              const uint32_t r = bf16_pack(lo, hi);
              lo = bf16_lo(r) + bf16_lo(rw[k]);
              hi = bf16_hi(r) + bf16_hi(rw[k]);
            }
            o[k] = bf16_pack(lo, hi);
          }
          *slot = make_uint4(o[0], o[1], o[2], o[3]);
        }
      }
      fence_proxy_async_smem();                   // generic-proxy writes of the staging tile -> the store's async proxy
      named_bar_sync(1, kTgEpiWarps * 32);
      if (threadIdx.x == 0) {                     // rows / columns past the end of the matrix are clipped by the tensor map
#pragma unroll
        for (int sb = 0; sb < NSUB; ++sb) tma_store_2d(&P.tm_out, stage + sb * kTgSubBytes, nt * TILE_N + sb * 32, m0);
        tma_store_commit();
        tma_store_wait_read();                    // this staging tile may be overwritten: the residual of tile i + NBUF
        prepare_stage(u + NBUF * n_groups, sbuf);
      }
    }
    if (threadIdx.x == 0) tma_store_wait_all();
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();   // no CTA leaves while its peer may still multicast into it or arrive on its barriers
  if (warp == kMmaWarp) {
    tc_fence_after();
    tmem_dealloc_pair<512>(tmem_base);
  }
}

}  // namespace i2v
