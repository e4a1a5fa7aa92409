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
// rows N + n0.. of the same nn.Linear weight, stacked into one 256-row K-major operand -- through a 3-stage ring), 1
// MMA-issuing warp (tcgen05.mma M = 128, N = 256, K = 16; fp32 accumulators double-buffered in TMEM: 2 x 256
// columns), 16 epilogue warps (lane quarter x column group) that overlap tile i's GELU with tile i+1's MMAs.
// Tiles are walked n-fastest, so the CTAs running at the same time share x row blocks through L2.
// Epilogue: the warp's share of the accumulator goes to registers and the TMEM buffer is released at once; biases from a
// per-warp fp32 table in shared memory; MUFU-free GELU on packed fp32 pairs (geglu_pair_poly); the bf16 tile is staged
// in 128-byte-swizzled shared memory and written with two tiled TMA stores.
//
// CL = 2: CTA pairs (clusters of two on one TPC) run the 2-SM MMA, tcgen05 cta_group::2, M = 256: the two CTAs own two
// row blocks of the same n-tile, each loads its own x tile and HALF of the weight operand (CTA 0 the hidden box, CTA 1
// the gate box) into its own shared memory, the leader CTA issues the MMAs for both and each CTA's TMEM receives its
// 128 rows of the 256 x 256 accumulator.  Per k-block a CTA takes in and reads back 32 KB instead of 48 KB: the
// single-CTA kernel is bound by exactly that (shared-memory bandwidth: 12 KB read + 12 KB written per k-step = 192 clk
// against 128 clk of math; at K = 320 the L2 -> SM path), and the stages become small enough for a 5-deep ring next to
// the output staging tile.
// Barriers: full[s] lives in the leader and counts both CTAs' TMA bytes; empty[s] / acc_full[b] exist in both CTAs and
// get the leader's multicast commits; acc_empty[b] lives in the leader and takes both CTAs' epilogue warps.
//
// XRES (CL = 2, K <= 320: SD1.5 level 0, where a tile is only five k-blocks): every pair owns a contiguous range of the (row pair, n-tile) order, so
// consecutive units share their x rows; the CTA's whole 128 x K x tile stays resident (80 KB) and only its half of the
// weight tile streams through the ring -- half the bytes per tile.
#pragma once
#include <cuda.h>
#include "ptx_sm100.cuh"
#include "norm_layout.cuh"   // gelu_erf, bf16 helpers

namespace i2v {

struct FfGegluParams {
  CUtensorMap tm_x;            // x [rows, K] bf16: dims (K, rows), box (64, 128), 128B swizzle
  CUtensorMap tm_w;            // W [2N, K] bf16:   dims (K, 2N),  box (64, 128)
  CUtensorMap tm_y;            // y [rows, ld] bf16: dims (ld, rows), box (64, 128): the epilogue's tiled stores
  const __nv_bfloat16* bias;   // [2N] or nullptr
  __nv_bfloat16* out;          // [rows, ld]
  long long rows;
  int N, K, ld;
  int m_tiles, n_tiles;
};

constexpr int kFfStages = 3;
constexpr int kFfEpiWarps = 16;                    // lane quarter x group of 32 output columns
constexpr int kFfThreads = (kFfEpiWarps + 2) * 32; // epilogue warps, TMA warp, MMA warp
constexpr int kFfABytes = 128 * 128;               // 128 rows x 64 bf16
constexpr int kFfBBytes = 256 * 128;               // 256 rows x 64 bf16
constexpr int kFfStageBytes = kFfABytes + kFfBBytes;
constexpr int kFfPairStages = 5;                   // CL = 2: x tile + half of the weight tile per stage
constexpr int kFfPairStageBytes = 2 * kFfABytes;
constexpr int kFfRingBytes = kFfPairStages * kFfPairStageBytes;   // 160 KB (>= the single-CTA ring, 144 KB)
static_assert(kFfStages * kFfStageBytes <= kFfRingBytes, "both variants share one smem layout");
constexpr int kFfOutBytes = 2 * kFfABytes;          // output tile staged for the tiled store: two [128 rows][64 bf16] halves
constexpr int kFfBiasBytes = kFfEpiWarps * 64 * 4;  // per epilogue warp: fp32 biases of its 32 hidden + 32 gate columns
constexpr int kFfSmemBytes = kFfRingBytes + kFfOutBytes + 256 + kFfBiasBytes + 1024;
static_assert(kFfSmemBytes <= 227 * 1024, "smem budget");

constexpr int kFfXresKBlocks = 5;                  // XRES: resident x tile of up to 5 k-blocks
constexpr int kFfXresStages = (kFfRingBytes - kFfXresKBlocks * kFfABytes) / kFfABytes;   // weight-half ring behind it

template <int CL, bool XRES = false>
__global__ void __launch_bounds__(kFfThreads, 1) ff_geglu_gemm_kernel(const __grid_constant__ FfGegluParams P) {
  static_assert(CL == 1 || CL == 2, "cluster size");
  static_assert(!XRES || CL == 2, "the x-resident variant is a CTA-pair kernel");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sm_out = smem + kFfRingBytes;           // [2][128 rows][128 B], 128-byte swizzle (what the store's tensor map expects)
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm_out + kFfOutBytes);
  constexpr int NST = XRES ? kFfXresStages : CL == 2 ? kFfPairStages : kFfStages;
  constexpr int STAGE_BYTES = XRES ? kFfABytes : CL == 2 ? kFfPairStageBytes : kFfStageBytes;
  uint8_t* ring = smem + (XRES ? kFfXresKBlocks * kFfABytes : 0);   // XRES: the resident x tile comes first
  uint64_t* bar_full = bars;                       // [stages]  TMA -> MMA
  uint64_t* bar_empty = bars + kFfPairStages;      // [stages]  MMA -> TMA
  uint64_t* bar_acc_full = bars + 2 * kFfPairStages;   // [2]       MMA -> epilogue
  uint64_t* bar_acc_empty = bar_acc_full + 2;      // [2]       epilogue -> MMA (one arrival per epilogue warp)
  uint64_t* bar_x_full = bar_acc_empty + 2;        // [1]       XRES: the resident x tile landed (leader, both CTAs' bytes)
  uint64_t* bar_x_free = bar_x_full + 1;           // [1]       XRES: every MMA on the old x tile retired (both CTAs)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_x_free + 1);
  float* sm_bias = reinterpret_cast<float*>(sm_out + kFfOutBytes + 256);   // [epilogue warp][hidden 32 | gate 32]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int kTmaWarp = kFfEpiWarps, kMmaWarp = kFfEpiWarps + 1;
  const int kblocks = P.K / 64;
  // work unit = CL row blocks (2 u + rank of the CTA in its cluster) of one n-tile; units are walked n-fastest
  const uint32_t rank = CL > 1 ? cluster_ctarank() : 0u;
  const long long units = (long long)((P.m_tiles + CL - 1) / CL) * P.n_tiles;
  const long long n_groups = gridDim.x / CL;
  const long long per = (units + n_groups - 1) / n_groups;
  const long long u_begin = XRES ? min(units, (long long)(blockIdx.x / CL) * per) : blockIdx.x / CL;
  const long long u_end = XRES ? min(units, u_begin + per) : units;
  const long long u_step = XRES ? 1 : n_groups;

  if (threadIdx.x == 0) {
    for (int s = 0; s < NST; ++s) {
      mbar_init(bar_full + s, 1);
      mbar_init(bar_empty + s, 1);
    }
    mbar_init(bar_x_full, 1);
    mbar_init(bar_x_free, 1);
    for (int b = 0; b < 2; ++b) {
      mbar_init(bar_acc_full + b, 1);
      mbar_init(bar_acc_empty + b, CL * kFfEpiWarps);
    }
    mbar_fence_init();
  }
  if (warp == kMmaWarp) {
    if (CL == 2) tmem_alloc_pair<512>(tmem_slot);
    else tmem_alloc<512>(tmem_slot);
  }
  if (warp == kTmaWarp && lane == 0) {
    tma_prefetch_desc(&P.tm_x);
    tma_prefetch_desc(&P.tm_w);
    tma_prefetch_desc(&P.tm_y);
  }
  tc_fence_before();
  __syncthreads();
  if (CL > 1) cluster_sync_all();   // the peer's barriers are initialised before anything arrives on them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == kTmaWarp) {
    if (lane == 0) {
      uint32_t g = 0;
      long long cur_mp = -1;
      int x_loads = 0;
      for (long long u = u_begin; u < u_end; u += u_step) {
        const int nt = (int)(u % P.n_tiles);
        const long long mp = u / P.n_tiles;
        const int m0 = ((int)mp * CL + (int)rank) * 128, n0 = nt * 128;
        if (XRES && mp != cur_mp) {   // new row pair: reload the resident x tile once its last MMAs have retired
          if (x_loads > 0) mbar_wait(bar_x_free, (x_loads - 1) & 1);
          if (rank == 0) mbar_arrive_expect_tx(bar_x_full, 2u * kblocks * kFfABytes);
          for (int kb = 0; kb < kblocks; ++kb)
            tma_load_2d_pair(smem + kb * kFfABytes, &P.tm_x, bar_x_full, kb * 64, m0, kEvictNormal);
          cur_mp = mp;
          ++x_loads;
        }
        for (int kb = 0; kb < kblocks; ++kb, ++g) {
          const int s = g % NST;
          mbar_wait(bar_empty + s, ((g / NST) & 1) ^ 1);
          uint8_t* a = ring + s * STAGE_BYTES;
          if (CL == 1) {
            mbar_arrive_expect_tx(bar_full + s, STAGE_BYTES);
            tma_load_2d(a, &P.tm_x, bar_full + s, kb * 64, m0, kEvictNormal);
            tma_load_2d(a + kFfABytes, &P.tm_w, bar_full + s, kb * 64, n0, kEvictLast);
            tma_load_2d(a + kFfABytes + kFfABytes, &P.tm_w, bar_full + s, kb * 64, P.N + n0, kEvictLast);
          } else if (XRES) {
            if (rank == 0) mbar_arrive_expect_tx(bar_full + s, 2 * STAGE_BYTES);
            tma_load_2d_pair(a, &P.tm_w, bar_full + s, kb * 64, (int)rank * P.N + n0, kEvictLast);
          } else {
            // the leader's barrier collects the bytes of both CTAs' loads
            if (rank == 0) mbar_arrive_expect_tx(bar_full + s, 2 * STAGE_BYTES);
            tma_load_2d_pair(a, &P.tm_x, bar_full + s, kb * 64, m0, kEvictNormal);
            tma_load_2d_pair(a + kFfABytes, &P.tm_w, bar_full + s, kb * 64, (int)rank * P.N + n0, kEvictLast);
          }
        }
      }
    }
  } else if (warp == kMmaWarp) {
    if (CL == 1 || rank == 0) {   // CTA pair: the leader issues for both
      constexpr uint32_t idesc = make_idesc_bf16(CL == 2 ? 256 : 128, 256, 0, 0);
      const uint64_t desc0 = make_smem_desc_sw128(0, 16, 1024);
      uint32_t g = 0;
      int i = 0, x_seen = 0;
      long long cur_mp = -1;
      for (long long u = u_begin; u < u_end; u += u_step, ++i) {
        const int b = i & 1;
        const long long mp = u / P.n_tiles;
        if (XRES && mp != cur_mp) {
          mbar_wait(bar_x_full, x_seen & 1);
          ++x_seen;
          cur_mp = mp;
        }
        const bool last_of_x = XRES && (u + 1 >= u_end || (u + 1) / P.n_tiles != mp);
        mbar_wait(bar_acc_empty + b, ((i >> 1) & 1) ^ 1);   // the epilogue warps have drained this accumulator buffer
        tc_fence_after();
        for (int kb = 0; kb < kblocks; ++kb, ++g) {
          const int s = g % NST;
          mbar_wait(bar_full + s, (g / NST) & 1);
          tc_fence_after();
          if (elect_one()) {
            const uint32_t aa = smem_u32(XRES ? smem + kb * kFfABytes : ring + s * STAGE_BYTES) >> 4;
            const uint32_t ba = XRES ? smem_u32(ring + s * STAGE_BYTES) >> 4 : aa + (kFfABytes >> 4);
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {   // 32 bytes per 16-column k-step inside the 128-byte swizzle row
              const uint64_t da = desc0 | (uint64_t)((aa + kk * 2) & 0x3FFF);
              const uint64_t db = desc0 | (uint64_t)((ba + kk * 2) & 0x3FFF);
              if (CL == 1) umma_ss(tmem_base + b * 256, da, db, idesc, (kb > 0 || kk > 0) ? 1u : 0u);
              else umma_ss_pair(tmem_base + b * 256, da, db, idesc, (kb > 0 || kk > 0) ? 1u : 0u);
            }
            if (CL == 1) {
              tc_commit(bar_empty + s);
              if (kb == kblocks - 1) tc_commit(bar_acc_full + b);
            } else {
              tc_commit_pair(bar_empty + s, (uint16_t)0b11);
              if (kb == kblocks - 1) {
                tc_commit_pair(bar_acc_full + b, (uint16_t)0b11);
                if (last_of_x) tc_commit_pair(bar_x_free, (uint16_t)0b11);
              }
            }
          }
          __syncwarp();
        }
      }
    }
  } else {
    // =========================== epilogue: warp = (lane quarter, column group) ===========================
    constexpr int CG = kFfEpiWarps / 4, CW = 128 / CG, NCH = CW / 16;   // column groups, their width, 16-column chunks
    const int quarter = warp & 3, half = warp >> 2;
    const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
    int i = 0;
    for (long long u = u_begin; u < u_end; u += u_step, ++i) {
      const int b = i & 1;
      const int nt = (int)(u % P.n_tiles);
      const long long row = ((u / P.n_tiles) * CL + rank) * 128 + quarter * 32 + lane;
      const int n0 = nt * 128 + half * CW;
      // this warp's biases as fp32 pairs in its own shared-memory slot (no cross-warp synchronisation), fetched
      // before the wait for the accumulator so that the global-load latency hides under the tile's MMAs
      float* wb = sm_bias + warp * 64;
      __syncwarp();   // the previous tile's reads of the slot are done
      wb[lane] = P.bias ? __bfloat162float(P.bias[n0 + lane]) : 0.f;
      wb[32 + lane] = P.bias ? __bfloat162float(P.bias[P.N + n0 + lane]) : 0.f;
      __syncwarp();
      mbar_wait(bar_acc_full + b, (i >> 1) & 1);
      tc_fence_after();
      const uint32_t tm = tmem_base + lane_addr + b * 256 + half * CW;
      const int trow = quarter * 32 + lane;                     // row of the tile
      uint8_t* srow = sm_out + (half * CW / 64) * kFfABytes + trow * 128;   // this row in its 64-column half
      const int c16 = (half * CW % 64) / 8;                     // first 16-byte chunk of this warp's columns in the row
      // The warp's whole share of the accumulator (32 hidden + 32 gate columns) goes to registers first and the buffer is
      // handed back to the MMA warp at once: a TMEM buffer is then busy for a tile's MMAs plus one load, not for the
      // whole epilogue, which matters where a tile's MMAs are shorter than its epilogue (K = 320).
      static_assert(NCH == 2, "register budget: two 16-column chunks per warp");
      uint32_t h[NCH][16], gt[NCH][16];
#pragma unroll
      for (int ch = 0; ch < NCH; ++ch) {
        tmem_ld_x16(tm + ch * 16, h[ch]);
        tmem_ld_x16(tm + 128 + ch * 16, gt[ch]);
      }
      tc_wait_ld();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (CL == 1) mbar_arrive(bar_acc_empty + b);
        else mbar_arrive_cluster(bar_acc_empty + b, 0u);   // the leader's barrier
      }
      // the staging tile is free once the previous tile's stores have read it (the issuing thread waits, then everyone
      // passes this barrier)
      if (threadIdx.x == 0 && i > 0) tma_store_wait_read();
      named_bar_sync(1, kFfEpiWarps * 32);
#pragma unroll
      for (int ch = 0; ch < NCH; ++ch) {
        const uint32_t* hc = h[ch];
        const uint32_t* gc = gt[ch];
        const uint64_t* bh2 = reinterpret_cast<const uint64_t*>(wb + ch * 16);
        const uint64_t* bg2 = reinterpret_cast<const uint64_t*>(wb + 32 + ch * 16);
        uint32_t o[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const uint64_t h2 = f2_add(f2_pack(__uint_as_float(hc[2 * k]), __uint_as_float(hc[2 * k + 1])), bh2[k]);
          const uint64_t g2 = f2_add(f2_pack(__uint_as_float(gc[2 * k]), __uint_as_float(gc[2 * k + 1])), bg2[k]);
          float r0, r1;
          f2_unpack(geglu_pair_poly(h2, g2), r0, r1);
          o[k] = bf16_pack(r0, r1);
        }
        // 128-byte swizzle: 16-byte chunk c of row r lives at chunk position c ^ (r & 7)
        *reinterpret_cast<uint4*>(srow + (((c16 + 2 * ch) ^ (trow & 7)) << 4)) = make_uint4(o[0], o[1], o[2], o[3]);
        *reinterpret_cast<uint4*>(srow + (((c16 + 2 * ch + 1) ^ (trow & 7)) << 4)) = make_uint4(o[4], o[5], o[6], o[7]);
      }
      if (P.ld > P.N && nt == 0 && half == 0 && row < P.rows)   // ones column for the next GEMM's deferred bias
        *reinterpret_cast<uint4*>(P.out + row * P.ld + P.N) = make_uint4(0x00003F80u, 0u, 0u, 0u);
      fence_proxy_async_smem();                   // generic-proxy writes of the staging tile -> the store's async proxy
      named_bar_sync(2, kFfEpiWarps * 32);
      if (threadIdx.x == 0) {                     // rows past the end of y are clipped by the tensor map
        const int m0 = (int)(row - trow);
        tma_store_2d(&P.tm_y, sm_out, nt * 128, m0);
        tma_store_2d(&P.tm_y, sm_out + kFfABytes, nt * 128 + 64, m0);
        tma_store_commit();
      }
    }
  }

  if (threadIdx.x == 0) tma_store_wait_all();     // the last tile's stores have left shared memory and are complete
  tc_fence_before();
  __syncthreads();
  if (CL > 1) cluster_sync_all();   // no CTA leaves while its peer may still multicast into it or arrive on its barriers
  if (warp == kMmaWarp) {
    tc_fence_after();
    if (CL == 2) tmem_dealloc_pair<512>(tmem_base);
    else tmem_dealloc<512>(tmem_base);
  }
}

}  // namespace i2v
