// Developer microbenchmark (round 2): issue rates of the packed half-precision FMA forms and of the integer / ALU
// instructions the FMA-pipe exp2 needs, alone and interleaved with MUFU.EX2, to find which share a pipe.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/pipes2 scripts/microbench/pipes2.cu && ./build/pipes2
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITERS 2048
#define UNROLL 16

__device__ __forceinline__ float ex2(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint32_t hfma2_bf(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("fma.rn.bf16x2 %0,%1,%2,%3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
__device__ __forceinline__ uint32_t hfma2_h(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("fma.rn.f16x2 %0,%1,%2,%3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
__device__ __forceinline__ uint32_t hfma2_h_imm(uint32_t a, uint32_t b) { uint32_t r; asm volatile("{.reg .b32 c; mov.b32 c, 0x3C003C00; fma.rn.f16x2 %0,%1,%2,c;}" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ uint32_t hadd2_h(uint32_t a, uint32_t b) { uint32_t r; asm volatile("add.rn.f16x2 %0,%1,%2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ uint32_t hmax2_bf(uint32_t a, uint32_t b) { uint32_t r; asm volatile("max.bf16x2 %0,%1,%2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ uint64_t ffma2(uint64_t a, uint64_t b, uint64_t c) { uint64_t r; asm volatile("fma.rn.f32x2 %0,%1,%2,%3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ float ffma_imm(float a, float b) { float r; asm volatile("fma.rn.f32 %0,%1,%2,0f3F000000;" : "=f"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ float ffma(float a, float b, float c) { float r; asm volatile("fma.rn.f32 %0,%1,%2,%3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }
__device__ __forceinline__ uint32_t packbf(float lo, float hi) { uint32_t r; asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo)); return r; }
__device__ __forceinline__ uint32_t packh(float lo, float hi) { uint32_t r; asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo)); return r; }
__device__ __forceinline__ uint32_t imad(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("mad.lo.u32 %0,%1,%2,%3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
__device__ __forceinline__ uint32_t shladd(uint32_t a, uint32_t c) { uint32_t r; asm volatile("{.reg .b32 t; shl.b32 t,%1,23; add.u32 %0,t,%2;}" : "=r"(r) : "r"(a), "r"(c)); return r; }
__device__ __forceinline__ uint32_t lop3_or(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("lop3.b32 %0,%1,%2,%3,0xFE;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b) { uint32_t r; asm volatile("prmt.b32 %0,%1,%2,0x7632;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ int imax(int a, int b) { int r; asm volatile("max.s32 %0,%1,%2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ float fmax3(float a, float b, float c) { float r; asm volatile("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }

template <int KIND>
__global__ void k(float* out, long long* cycles) {
  float v[UNROLL];
  uint64_t w[UNROLL];
  uint32_t u[UNROLL], h[UNROLL];
  for (int i = 0; i < UNROLL; ++i) {
    v[i] = threadIdx.x * 0.001f + i; w[i] = ((uint64_t)__float_as_uint(v[i]) << 32) | __float_as_uint(v[i] + 1.f);
    u[i] = i + threadIdx.x; h[i] = 0x3C003C00u + i;
  }
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < UNROLL; ++i) {
      const int j = (i + 1) % UNROLL, l = (i + 2) % UNROLL;
      if (KIND == 0) h[i] = hfma2_bf(h[i], h[j], h[l]);
      if (KIND == 1) h[i] = hfma2_h(h[i], h[j], h[l]);
      if (KIND == 2) h[i] = hfma2_h_imm(h[i], h[j]);
      if (KIND == 3) h[i] = hadd2_h(h[i], h[j]);
      if (KIND == 4) h[i] = hmax2_bf(h[i], h[j]);
      if (KIND == 5) v[i] = ffma_imm(v[i], v[j]);
      if (KIND == 6) u[i] = imad(u[i], 0x800000u, u[j]);
      if (KIND == 7) u[i] = shladd(u[i], u[j]);
      if (KIND == 8) u[i] = lop3_or(u[i], u[j], u[l]);
      if (KIND == 9) u[i] = prmt(u[i], u[j]);
      if (KIND == 10) u[i] = (uint32_t)imax((int)u[i], (int)u[j]);
      if (KIND == 11) u[i] = packh(__uint_as_float(u[i]), __uint_as_float(u[j]));
      // pairs with MUFU: does the second instruction steal XU issue?
      if (KIND == 20) { v[i] = ex2(v[i]); h[i] = hfma2_h(h[i], h[j], h[l]); }
      if (KIND == 21) { v[i] = ex2(v[i]); u[i] = packbf(__uint_as_float(u[i]), __uint_as_float(u[j])); }
      if (KIND == 22) { v[i] = ex2(v[i]); u[i] = lop3_or(u[i], u[j], u[l]); }
      if (KIND == 23) { v[i] = ex2(v[i]); u[i] = imad(u[i], 0x800000u, u[j]); }
      if (KIND == 24) { v[i] = ex2(v[i]); v[j] = fmax3(v[j], v[l], v[i]); }
      // pairs on the FMA / ALU side
      if (KIND == 30) { w[i] = ffma2(w[i], w[j], w[l]); h[i] = hfma2_h(h[i], h[j], h[l]); }
      if (KIND == 31) { w[i] = ffma2(w[i], w[j], w[l]); u[i] = imad(u[i], 0x800000u, u[j]); }
      if (KIND == 32) { u[i] = packbf(__uint_as_float(u[i]), __uint_as_float(u[j])); u[j] = lop3_or(u[j], u[l], u[i]); }
      if (KIND == 33) { u[i] = packbf(__uint_as_float(u[i]), __uint_as_float(u[j])); v[i] = fmax3(v[i], v[j], v[l]); }
      if (KIND == 34) { w[i] = ffma2(w[i], w[j], w[l]); u[i] = lop3_or(u[i], u[j], u[l]); }
      if (KIND == 35) { v[i] = ffma(v[i], v[j], v[l]); v[j] = ffma_imm(v[j], v[l]); }
      // the MUFU path of one column pair: 2 MUFU + pack, with and without an OR-reduction of the packed words
      if (KIND == 40) { v[i] = ex2(v[i]); v[j] = ex2(v[j]); u[i] = packbf(v[i], v[j]); }
      if (KIND == 41) { v[i] = ex2(v[i]); v[j] = ex2(v[j]); u[i] = packbf(v[i], v[j]); u[l] = lop3_or(u[l], u[i], u[j]); }
    }
  }
  long long t1 = clock64();
  float acc = 0;
  for (int i = 0; i < UNROLL; ++i) acc += v[i] + __uint_as_float((uint32_t)w[i]) + __uint_as_float(u[i]) + __uint_as_float(h[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

template <int KIND>
void run(const char* name, int ops_per_iter) {
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
  printf("%-34s", name);
  for (int warps : {4, 8, 16}) {
    k<KIND><<<148, warps * 32>>>(out, cyc);
    cudaDeviceSynchronize();
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    double inst = (double)ITERS * UNROLL * ops_per_iter * warps;  // warp-instructions per SM
    printf("  w/SM %2d: %5.2f clk per warp-inst per SMSP", warps, c / (inst / 4));
  }
  printf("\n");
  cudaFree(out); cudaFree(cyc);
}

int main() {
  run<0>("HFMA2.BF16 (3 reg)", 1);
  run<1>("HFMA2 f16 (3 reg)", 1);
  run<2>("HFMA2 f16 (const c)", 1);
  run<3>("HADD2 f16", 1);
  run<4>("HMNMX2.BF16", 1);
  run<5>("FFMA imm", 1);
  run<6>("IMAD (x * 2^23 + y)", 1);
  run<7>("SHL + IADD (LEA?)", 1);
  run<8>("LOP3 (3-input OR)", 1);
  run<9>("PRMT", 1);
  run<10>("IMNMX", 1);
  run<11>("F2FP.F16.PACK", 1);
  run<20>("ex2 + hfma2", 2);
  run<21>("ex2 + f2fp", 2);
  run<22>("ex2 + lop3", 2);
  run<23>("ex2 + imad", 2);
  run<24>("ex2 + fmnmx3", 2);
  run<30>("ffma2 + hfma2", 2);
  run<31>("ffma2 + imad", 2);
  run<32>("f2fp + lop3", 2);
  run<33>("f2fp + fmnmx3", 2);
  run<34>("ffma2 + lop3", 2);
  run<35>("ffma + ffma imm", 2);
  run<40>("2 ex2 + f2fp", 3);
  run<41>("2 ex2 + f2fp + lop3", 4);
  return 0;
}
