// Developer microbenchmark: MUFU.EX2 warp-instruction cost per sub-partition for f32 / f16 / bf16 operands
// (ex2.approx.f16x2 and .bf16x2 compile to two MUFU ops + PRMT; the question is whether the half forms issue faster).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mufu_half mufu_half.cu && ./mufu_half
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define ITERS 4096
#define UNROLL 16
template <int KIND>
__global__ void k(uint32_t* out, long long* cycles) {
  uint32_t v[UNROLL];
  for (int i = 0; i < UNROLL; ++i) v[i] = 0x3c003c00u ^ (threadIdx.x * 17 + i);
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < UNROLL; ++i) {
      if (KIND == 0) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(__uint_as_float(v[i]))); v[i] = __float_as_uint(y); }
      if (KIND == 1) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(v[i]));
      if (KIND == 2) asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(v[i]));
      if (KIND == 3) asm volatile("tanh.approx.f16x2 %0, %0;" : "+r"(v[i]));
    }
  }
  long long t1 = clock64();
  uint32_t s = 0;
  for (int i = 0; i < UNROLL; ++i) s ^= v[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}
template <int KIND>
void run(const char* name, int per) {
  uint32_t* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
  printf("%-28s", name);
  for (int warps : {4, 8, 16}) {
    k<KIND><<<148, warps * 32>>>(out, cyc);
    cudaDeviceSynchronize();
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    printf("  w/SMSP %d: %.2f clk per exp-warp-instr", warps / 4, (double)c / ITERS / UNROLL / per / (warps / 4));
  }
  printf("\n");
}
int main() {
  run<0>("ex2.f32", 1);
  run<1>("ex2.f16x2 (2 exps)", 2);
  run<2>("ex2.bf16x2 (2 exps)", 2);
  run<3>("tanh.f16x2 (2)", 2);
  return 0;
}
