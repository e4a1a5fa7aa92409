// Developer microbenchmark: the softmax inner step (64 scores per thread: row max, scale, exp2 split between MUFU and
// the FMA-pipe polynomial, row sum, bf16 pack) on register data only -- no TMEM, no MMA, no barriers.  Measures the
// cycles per warp-step at 1/2/4 warps per SM sub-partition to find the exp2 split and instruction order that keep XU,
// FMA and ALU pipes busy together.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../i2v_adapter_unofficial_b200/csrc/ptx_sm100.cuh"
using namespace i2v;

#define ITERS 2048
constexpr int BN = 64;

// MODE 0: compiler-scheduled, EMU of every 8 pairs emulated (the product kernel's loop)
// MODE 1: phased: all scale FFMA2 first, then MUFU run, then emulation run, then sums/packs (asm volatile order)
template <int EMU, int MODE, bool WITH_SUM, bool WITH_CLAMP>
__global__ void k(float* out, long long* cycles, float c, float mref_in) {
  float sv[BN];
  for (int i = 0; i < BN; ++i) sv[i] = -(float)((threadIdx.x * 7 + i * 13) % 97) * 0.37f;
  float l = 0.f, m_ref = mref_in;
  uint32_t sink = 0;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < BN; ++i) asm volatile("" : "+f"(sv[i]));
    float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
    for (int i = 0; i < BN; i += 8) {
      mx0 = fmax3(mx0, sv[i + 0], sv[i + 1]);
      mx1 = fmax3(mx1, sv[i + 2], sv[i + 3]);
      mx2 = fmax3(mx2, sv[i + 4], sv[i + 5]);
      mx3 = fmax3(mx3, sv[i + 6], sv[i + 7]);
    }
    const float mx = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3)) * c;
    const uint64_t c2 = f2_pack(c, c);
    const uint64_t nm2 = f2_pack(-m_ref, -m_ref);
    uint64_t ls0 = 0ull, ls1 = 0ull;
    uint32_t pk[BN / 2];
    if (MODE == 0) {
#pragma unroll
      for (int i = 0; i < BN / 2; ++i) {
        const uint64_t x = f2_fma(f2_pack(sv[2 * i], sv[2 * i + 1]), c2, nm2);
        uint64_t p2;
        if ((i & 7) < EMU) {
          if (WITH_CLAMP) p2 = ex2_emulated_pair(x); else p2 = ex2_emulated_pair_noclamp(x);
        } else {
          float x0, x1;
          f2_unpack(x, x0, x1);
          p2 = f2_pack(ex2_approx(x0), ex2_approx(x1));
        }
        if (WITH_SUM) { if (i & 1) ls1 = f2_add(ls1, p2); else ls0 = f2_add(ls0, p2); }
        float p0, p1;
        f2_unpack(p2, p0, p1);
        pk[i] = pack_bf16x2(p0, p1);
      }
    } else {
      uint64_t x[BN / 2];
#pragma unroll
      for (int i = 0; i < BN / 2; ++i) asm volatile("fma.rn.f32x2 %0,%1,%2,%3;" : "=l"(x[i]) : "l"(f2_pack(sv[2 * i], sv[2 * i + 1])), "l"(c2), "l"(nm2));
      // MUFU run
#pragma unroll
      for (int i = 0; i < BN / 2; ++i) {
        if ((i & 7) >= EMU) {
          float x0, x1, y0, y1;
          f2_unpack(x[i], x0, x1);
          asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y0) : "f"(x0));
          asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y1) : "f"(x1));
          x[i] = f2_pack(y0, y1);
        }
      }
      // emulation run
#pragma unroll
      for (int i = 0; i < BN / 2; ++i) {
        if ((i & 7) < EMU) x[i] = WITH_CLAMP ? ex2_emulated_pair(x[i]) : ex2_emulated_pair_noclamp(x[i]);
      }
#pragma unroll
      for (int i = 0; i < BN / 2; ++i) {
        if (WITH_SUM) { if (i & 1) ls1 = f2_add(ls1, x[i]); else ls0 = f2_add(ls0, x[i]); }
        float p0, p1;
        f2_unpack(x[i], p0, p1);
        pk[i] = pack_bf16x2(p0, p1);
      }
    }
    float a0, a1;
    f2_unpack(f2_add(ls0, ls1), a0, a1);
    l += a0 + a1;
    if (mx > m_ref + 1000.f) m_ref = mx;  // never taken; keeps the max live
#pragma unroll
    for (int i = 0; i < BN / 2; ++i) asm volatile("" ::"r"(pk[i]));
    sink ^= pk[it & 31];
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = l + m_ref + __uint_as_float(sink);
  if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

template <int EMU, int MODE, bool WITH_SUM, bool WITH_CLAMP>
void run() {
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
  printf("EMU %d/8 mode %d sum %d clamp %d:", EMU, MODE, (int)WITH_SUM, (int)WITH_CLAMP);
  for (int warps : {4, 8, 16}) {
    cudaMemset(cyc, 0, 8);
    k<EMU, MODE, WITH_SUM, WITH_CLAMP><<<148, warps * 32>>>(out, cyc, 0.228f, 0.f);
    cudaError_t e = cudaDeviceSynchronize();
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    // clk per 128x128 tile-equivalent per SM: each SMSP handles 2 warp-steps of 64 per tile-equiv
    double per_step = (double)c / ITERS;              // wall clk for every warp to do one step
    double tile_equiv = per_step / (warps / 4) * 2;   // SMSP processes (warps/4) steps concurrently
    printf("  w/SMSP %d: %6.0f clk/step -> %5.0f clk/tile-eq%s", warps / 4, per_step, tile_equiv, e == cudaSuccess ? "" : " (ERR)");
  }
  printf("\n");
  cudaFree(out); cudaFree(cyc);
}

int main() {
  run<0, 0, true, true>();
  run<2, 0, true, true>();
  run<3, 0, true, true>();
  run<4, 0, true, true>();
  run<3, 0, false, true>();
  run<4, 0, false, true>();
  run<3, 0, false, false>();
  run<4, 0, false, false>();
  run<0, 1, true, true>();
  run<2, 1, true, true>();
  run<3, 1, true, true>();
  run<4, 1, true, true>();
  run<3, 1, false, false>();
  run<4, 1, false, false>();
  return 0;
}
