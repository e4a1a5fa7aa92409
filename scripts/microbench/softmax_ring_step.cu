// Developer microbenchmark for the ring-scheduled softmax step of dense_attn_ring_sm100.cuh: ONE warp per SM
// sub-partition owns the XU at a time, so what matters is how close a single in-order warp gets to the MUFU rate
// (8 clk per MUFU.EX2 warp-instruction) while the FMA-pipe polynomial handles EMU of every 8 column pairs.
// The reference max is an integer (log2 domain), which folds the Cody-Waite split into the scale FMA:
//   t = fma.rm(s, c, magic - m)   = magic + floor(x),  x = s*c - m
//   f = fma.rn(s, c, (magic - m) - t) = x - floor(x)
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o softmax_ring_step softmax_ring_step.cu && ./softmax_ring_step
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../i2v_adapter_unofficial_b200/csrc/ptx_sm100.cuh"
using namespace i2v;

#define ITERS 2048
constexpr int BN = 64;

template <int EMU, int DEG, bool CLAMP, bool WITH_MAX, bool PRE = false, bool SUM = true>
__global__ void k(float* out, long long* cycles, float c, float mref_in) {
  extern __shared__ float4 sbuf[];  // [BN/4][blockDim.x]: this thread's scores, re-read every step (stands in for tcgen05.ld)
  for (int i = 0; i < BN / 4 + 1; ++i) {
    float4 v;
    v.x = -(float)((threadIdx.x * 7 + (4 * i + 0) * 13) % 97) * 0.37f;
    v.y = -(float)((threadIdx.x * 7 + (4 * i + 1) * 13) % 97) * 0.37f;
    v.z = -(float)((threadIdx.x * 7 + (4 * i + 2) * 13) % 97) * 0.37f;
    v.w = -(float)((threadIdx.x * 7 + (4 * i + 3) * 13) % 97) * 0.37f;
    sbuf[i * blockDim.x + threadIdx.x] = v;
  }
  float sv[BN];
  float l = 0.f, m_ref = mref_in;
  uint32_t sink = 0;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < BN / 4; ++i) {
      asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(sv[4 * i]), "=f"(sv[4 * i + 1]), "=f"(sv[4 * i + 2]), "=f"(sv[4 * i + 3])
                   : "r"(smem_u32(&sbuf[(i + (it & 1)) * blockDim.x + threadIdx.x])) : "memory");
    }
    float mx = 0.f;
    if (WITH_MAX) {
      float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
      for (int i = 0; i < BN; i += 8) {
        mx0 = fmax3(mx0, sv[i + 0], sv[i + 1]);
        mx1 = fmax3(mx1, sv[i + 2], sv[i + 3]);
        mx2 = fmax3(mx2, sv[i + 4], sv[i + 5]);
        mx3 = fmax3(mx3, sv[i + 6], sv[i + 7]);
      }
      mx = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3)) * c;
    }
    uint32_t pk[BN / 2];
    const float sum = softmax_exp_row<BN, EMU, DEG, CLAMP, PRE, SUM>(sv, c, m_ref, pk);
    l += sum;
    if (mx > m_ref + 1000.f) m_ref = mx;  // never taken; keeps the max live
    uint32_t x0 = 0, x1 = 0, x2 = 0, x3 = 0;  // register-only sink (a dynamically indexed pk[] would go through
#pragma unroll                               //  local memory and put a long-scoreboard stall into every step)
    for (int i = 0; i < BN / 2; i += 4) { x0 ^= pk[i]; x1 ^= pk[i + 1]; x2 ^= pk[i + 2]; x3 ^= pk[i + 3]; }
    sink += (x0 ^ x1) + (x2 ^ x3);
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = l + m_ref + __uint_as_float(sink);
  if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

template <int EMU, int DEG, bool CLAMP, bool WITH_MAX, bool PRE = false, bool SUM = true>
void run() {
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
  printf("EMU %d/8 deg %d clamp %d max %d pre %d sum %d:", EMU, DEG, (int)CLAMP, (int)WITH_MAX, (int)PRE, (int)SUM);
  for (int warps : {4, 8, 12, 16}) {
    cudaMemset(cyc, 0, 8);
    cudaFuncSetAttribute(k<EMU, DEG, CLAMP, WITH_MAX, PRE, SUM>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
    k<EMU, DEG, CLAMP, WITH_MAX, PRE, SUM><<<148, warps * 32, warps * 32 * (BN + 4) * 4>>>(out, cyc, 0.228f, 0.f);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    double per_step = (double)c / ITERS;             // clk for every resident warp to do one 64-column step
    double per_warp_step = per_step / (warps / 4);   // XU-time share of one warp-step on its sub-partition
    printf("  w/SMSP %d: %5.0f clk/step (%4.0f per warp-step)%s", warps / 4, per_step, per_warp_step,
           e == cudaSuccess ? "" : " (ERR)");
  }
  printf("\n");
  cudaFree(out); cudaFree(cyc);
}


// Feature-flag variant used to find which pipe (or pipe interaction) bounds the step.
template <int EMU, bool MUFU_ON, bool SUM, bool PACK, bool SCALE>
__global__ void k2(float* out, long long* cycles, float c, float m) {
  float sv[BN];
  for (int i = 0; i < BN; ++i) sv[i] = -(float)((threadIdx.x * 7 + i * 13) % 97) * 0.37f;
  float l = 0.f;
  uint32_t sink = 0;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < BN; ++i) asm volatile("// keep %0" : "+f"(sv[i]) : "r"(it));
    const uint64_t c2 = f2_pack(c, c);
    const uint64_t nm2 = f2_pack(-m, -m);
    const uint64_t k1 = f2_pack(kExpMagic - m, kExpMagic - m);
    uint64_t ls0 = 0ull, ls1 = 0ull;
    uint32_t acc = 0;
#pragma unroll
    for (int i = 0; i < BN / 2; ++i) {
      const uint64_t s2 = f2_pack(sv[2 * i], sv[2 * i + 1]);
      uint64_t p2;
      if ((i & 7) < EMU) {
        p2 = ex2_emu_pair_int<3, true>(s2, c2, k1);
      } else {
        float x0, x1;
        f2_unpack(SCALE ? f2_fma(s2, c2, nm2) : s2, x0, x1);
        if (MUFU_ON) p2 = f2_pack(ex2_approx(x0), ex2_approx(x1)); else p2 = f2_pack(x0, x1);
      }
      if (SUM) { if (i & 1) ls1 = f2_add(ls1, p2); else ls0 = f2_add(ls0, p2); }
      float p0, p1;
      f2_unpack(p2, p0, p1);
      if (PACK) { acc ^= pack_bf16x2(p0, p1); }
      else { acc ^= __float_as_uint(p0) ^ __float_as_uint(p1); }
    }
    float a0, a1;
    f2_unpack(f2_add(ls0, ls1), a0, a1);
    l += a0 + a1;
    sink += acc;
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = l + __uint_as_float(sink);
  if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

template <int EMU, bool MUFU_ON, bool SUM, bool PACK, bool SCALE>
void run2() {
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
  printf("k2 EMU %d/8 mufu %d sum %d pack %d scale %d:", EMU, (int)MUFU_ON, (int)SUM, (int)PACK, (int)SCALE);
  for (int warps : {4, 8, 12, 16}) {
    cudaMemset(cyc, 0, 8);
    k2<EMU, MUFU_ON, SUM, PACK, SCALE><<<148, warps * 32>>>(out, cyc, 0.228f, 0.f);
    cudaError_t e = cudaDeviceSynchronize();
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    double per_step = (double)c / ITERS;
    printf("  w/SMSP %d: %5.0f (%4.0f per warp-step)%s", warps / 4, per_step, per_step / (warps / 4), e == cudaSuccess ? "" : " (ERR)");
  }
  printf("\n");
  cudaFree(out); cudaFree(cyc);
}

// accuracy of the emulated exp2 against exp2f over the range the kernel produces
__global__ void acc(float* maxrel) {
  float worst2 = 0.f, worst3 = 0.f;
  for (int i = threadIdx.x; i < (1 << 20); i += blockDim.x) {
    const float s = -140.f + 150.f * (float)i / (float)(1 << 20);
    float sv[2] = {s, s + 0.01f};
    uint64_t p3 = ex2_emu_pair_int<3, true>(f2_pack(sv[0], sv[1]), f2_pack(1.f, 1.f), f2_pack(12582912.f - 2.f, 12582912.f - 2.f));
    uint64_t p2 = ex2_emu_pair_int<2, true>(f2_pack(sv[0], sv[1]), f2_pack(1.f, 1.f), f2_pack(12582912.f - 2.f, 12582912.f - 2.f));
    float a, b;
    f2_unpack(p3, a, b);
    const float ref = exp2f(fmaxf(s - 2.f, -126.f - 0.f));
    if (s - 2.f > -125.f) worst3 = fmaxf(worst3, fabsf(a - ref) / ref);
    f2_unpack(p2, a, b);
    if (s - 2.f > -125.f) worst2 = fmaxf(worst2, fabsf(a - ref) / ref);
    if (s - 2.f < -130.f && !(a >= 0.f && a < 1e-30f)) worst3 = 1e9f;  // clamp must give ~0, never garbage
  }
  atomicMax((int*)maxrel, __float_as_int(worst3));
  atomicMax((int*)maxrel + 1, __float_as_int(worst2));
}

int main(int argc, char** argv) {
  if (argc > 1) {  // ncu target
    run<0, 3, true, true>();
    run<4, 3, true, true>();
    return 0;
  }
  float* mr; cudaMalloc(&mr, 8); cudaMemset(mr, 0, 8);
  acc<<<1, 256>>>(mr);
  float h[2]; cudaMemcpy(h, mr, 8, cudaMemcpyDeviceToHost);
  printf("emulated exp2 max rel err: deg3 %.3e  deg2 %.3e\n", h[0], h[1]);
  run<0, 3, true, true, true, false>();
  run<2, 3, true, true, true, false>();
  run<3, 3, true, true, true, false>();
  run<4, 3, true, true, true, false>();
  run<5, 3, true, true, true, false>();
  run<3, 2, true, true, true, false>();
  run<4, 2, true, true, true, false>();
  run<5, 2, true, true, true, false>();
  run<3, 3, true, true, false, true>();
  run<3, 3, true, true, true, true>();
  run<3, 3, true, true, false, false>();
  return 0;
}
