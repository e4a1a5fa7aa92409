// Developer microbenchmark: per-SM throughput of the instruction kinds the softmax inner loop uses.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes pipes.cu && ./pipes
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITERS 4096
#define UNROLL 16

__device__ __forceinline__ float ex2(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint64_t ffma2(uint64_t a, uint64_t b, uint64_t c) { uint64_t r; asm volatile("fma.rn.f32x2 %0,%1,%2,%3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ uint64_t fadd2(uint64_t a, uint64_t b) { uint64_t r; asm volatile("add.rn.f32x2 %0,%1,%2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ float fmax3(float a, float b, float c) { float r; asm volatile("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }
__device__ __forceinline__ uint32_t packbf(float lo, float hi) { uint32_t r; asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo)); return r; }
__device__ __forceinline__ float ffma(float a, float b, float c) { float r; asm volatile("fma.rn.f32 %0,%1,%2,%3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }

template <int KIND>
__global__ void k(float* out, long long* cycles) {
  float v[UNROLL];
  uint64_t w[UNROLL];
  uint32_t u[UNROLL];
  for (int i = 0; i < UNROLL; ++i) { v[i] = threadIdx.x * 0.001f + i; w[i] = ((uint64_t)__float_as_uint(v[i]) << 32) | __float_as_uint(v[i] + 1.f); u[i] = i; }
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < UNROLL; ++i) {
      if (KIND == 0) v[i] = ex2(v[i]);
      if (KIND == 1) w[i] = ffma2(w[i], w[i], w[i]);
      if (KIND == 2) w[i] = fadd2(w[i], w[i]);
      if (KIND == 3) v[i] = fmax3(v[i], v[(i + 1) % UNROLL], v[(i + 2) % UNROLL]);
      if (KIND == 4) u[i] = packbf(__uint_as_float(u[i]), v[i]);
      if (KIND == 5) v[i] = ffma(v[i], v[i], v[i]);
      if (KIND == 6) { v[i] = ex2(v[i]); w[i] = ffma2(w[i], w[i], w[i]); u[i] = packbf(__uint_as_float(u[i]), v[i]); }  // mix
      if (KIND == 8) { w[i] = ffma2(w[i], w[i], w[i]); v[i] = fmax3(v[i], v[(i + 1) % UNROLL], v[(i + 2) % UNROLL]); }
      if (KIND == 9) { w[i] = ffma2(w[i], w[i], w[i]); u[i] = packbf(__uint_as_float(u[i]), __uint_as_float(u[(i + 1) % UNROLL])); }
      if (KIND == 10) { w[i] = ffma2(w[i], w[i], w[i]); v[i] = ex2(v[i]); }
      if (KIND == 11) { v[i] = ex2(v[i]); u[i] = packbf(__uint_as_float(u[i]), __uint_as_float(u[(i + 1) % UNROLL])); }
      if (KIND == 12) { w[i] = ffma2(w[i], w[i], w[i]); w[i] = ffma2(w[i], w[i], w[i]); w[i] = ffma2(w[i], w[i], w[i]); w[i] = ffma2(w[i], w[i], w[i]); v[i] = ex2(v[i]); }
      if (KIND == 13) { w[i] = ffma2(w[i], w[i], w[i]); w[i] = ffma2(w[i], w[i], w[i]); u[i] = packbf(__uint_as_float(u[i]), __uint_as_float(u[(i + 1) % UNROLL])); u[i] = packbf(__uint_as_float(u[i]), __uint_as_float(u[(i + 1) % UNROLL])); v[i] = ex2(v[i]); }
      if (KIND == 14) { v[i] = ffma(v[i], v[i], v[i]); u[i] = packbf(__uint_as_float(u[i]), __uint_as_float(u[(i + 1) % UNROLL])); }
      if (KIND == 7) { v[i] = ex2(v[i]); w[i] = ffma2(w[i], w[i], w[i]); w[i] = fadd2(w[i], w[i]); u[i] = packbf(__uint_as_float(u[i]), v[i]); v[i] = fmax3(v[i], v[(i + 1) % UNROLL], v[(i+2)%UNROLL]); }
    }
  }
  long long t1 = clock64();
  float acc = 0;
  for (int i = 0; i < UNROLL; ++i) acc += v[i] + __uint_as_float((uint32_t)w[i]) + __uint_as_float(u[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

template <int KIND>
void run(const char* name, int ops_per_iter) {
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
  for (int warps : {4, 8, 16}) {
    k<KIND><<<148, warps * 32>>>(out, cyc);
    cudaDeviceSynchronize();
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    double inst = (double)ITERS * UNROLL * ops_per_iter * warps;  // warp-instructions per SM
    printf("%-28s warps/SM %2d: %.3f warp-inst/clk/SM  (%.2f clk per warp-inst per SMSP)\n", name, warps, inst / c, c / (inst / 4));
  }
  cudaFree(out); cudaFree(cyc);
}

int main() {
  run<0>("MUFU.EX2", 1);
  run<1>("FFMA2", 1);
  run<2>("FADD2", 1);
  run<3>("FMNMX3", 1);
  run<4>("F2FP.BF16.PACK", 1);
  run<5>("FFMA", 1);
  run<6>("mix ex2+ffma2+f2fp", 3);
  run<7>("mix ex2+ffma2+fadd2+f2fp+fmnmx3", 5);
  run<8>("ffma2+fmnmx3", 2);
  run<9>("ffma2+f2fp", 2);
  run<10>("ffma2+ex2", 2);
  run<11>("ex2+f2fp", 2);
  run<12>("4ffma2+ex2", 5);
  run<13>("2ffma2+2f2fp+ex2", 5);
  run<14>("ffma+f2fp", 2);
  return 0;
}
