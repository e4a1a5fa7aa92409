// Developer microbenchmark: how fast can one B200 move the IP-Adapter kernel's operand pattern through the LSU?
// Q / O are [B, S, H, 40] bf16: a (frame, head) query tile is 128 rows of 80 contiguous bytes at a 640-byte pitch.
// Every warp copies tiles global -> shared (cp.async, 16 B per lane, 6.4 rows per instruction) -> global (16 B per lane in
// memory order), nothing else.  WARPS warps per CTA, one CTA per SM.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/rows80_copy scripts/microbench/rows80_copy.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int B = 32, S = 4096, H = 8, D = 40;

template <int MODE>   // 0: load + store, 1: load only, 2: store only
__global__ void k(const uint8_t* __restrict__ q, uint8_t* __restrict__ o, int warps_total) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t* buf = smem + warp * 16384;
  const uint32_t sbuf = (uint32_t)__cvta_generic_to_shared(buf);
  uint32_t off[20];
#pragma unroll
  for (int kx = 0; kx < 20; ++kx) {
    const int g = kx * 32 + lane, r = g / 5, c = g - r * 5;
    off[kx] = (uint32_t)(r * (H * D * 2) + c * 16);
  }
  const int n_tiles = B * (S / 128) * H;
  const int gw = blockIdx.x * (blockDim.x >> 5) + warp;
  for (int t = gw; t < n_tiles; t += warps_total) {
    const int h = t % H, rest = t / H;   // heads fastest: the 8 heads of a row block are in flight together
    const long long base = (long long)rest * 128 * (H * D * 2) + h * (D * 2);
    if (MODE != 2) {
#pragma unroll
      for (int kx = 0; kx < 20; ++kx)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sbuf + (kx * 32 + lane) * 16), "l"(q + base + off[kx]) : "memory");
      asm volatile("cp.async.wait_all;" ::: "memory");
      __syncwarp();
    }
    if (MODE != 1) {
#pragma unroll
      for (int kx = 0; kx < 20; ++kx) {
        const uint4 v = *reinterpret_cast<const uint4*>(buf + (kx * 32 + lane) * 16);
        *reinterpret_cast<uint4*>(o + base + off[kx]) = v;
      }
      __syncwarp();
    }
  }
}

template <int MODE>
void run(const uint8_t* q, uint8_t* o, int warps) {
  cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, warps * 16384);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int i = 0; i < 3; ++i) k<MODE><<<148, warps * 32, warps * 16384>>>(q, o, 148 * warps);
  cudaEventRecord(e0);
  for (int i = 0; i < 10; ++i) k<MODE><<<148, warps * 32, warps * 16384>>>(q, o, 148 * warps);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double bytes = (double)B * S * H * D * 2 * (MODE == 0 ? 2 : 1);
  printf("mode %d warps/CTA %2d: %7.1f us  %6.0f GB/s  (%s)\n", MODE, warps, ms * 100, bytes / (ms / 10 * 1e-3) / 1e9,
         cudaGetErrorString(cudaGetLastError()));
}

int main() {
  uint8_t *q, *o;
  const size_t n = (size_t)B * S * H * D * 2;
  cudaMalloc(&q, n); cudaMalloc(&o, n); cudaMemset(q, 1, n);
  for (int w : {4, 8, 12}) { run<0>(q, o, w); run<1>(q, o, w); run<2>(q, o, w); }
  return 0;
}
