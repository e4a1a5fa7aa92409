// Developer microbenchmark: the augmented-layout softmax step with a CENTERED reference (the QK^T MMA delivers
// y = s - m_center, fast path iff max|y| <= 67) against the current step (row max + clamped emulation), on data that
// is re-read from shared memory every step (stands in for tcgen05.ld) and written back packed (stands in for tcgen05.st).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/softmax_centered_step scripts/microbench/softmax_centered_step.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../i2v_adapter_unofficial_b200/csrc/ptx_sm100.cuh"
using namespace i2v;

#define ITERS 2048
constexpr int BN = 64;

// MODE 0: current (row max of 64, vote, clamped emulation);  MODE 1: centered (abs max of 64, vote, no clamp)
// MODE 2: centered, no range check at all (floor of the exponential code)
template <int EMU, int DEG, int MODE>
__global__ void __launch_bounds__(512) k(float* out, long long* cycles) {
  extern __shared__ float4 sbuf[];  // [BN/4 + 1][blockDim.x]
  for (int i = 0; i < BN / 4 + 1; ++i) {
    float4 v;
    v.x = -(float)((threadIdx.x * 7 + (4 * i + 0) * 13) % 97) * 0.37f;
    v.y = -(float)((threadIdx.x * 7 + (4 * i + 1) * 13) % 97) * 0.37f;
    v.z = -(float)((threadIdx.x * 7 + (4 * i + 2) * 13) % 97) * 0.37f;
    v.w = -(float)((threadIdx.x * 7 + (4 * i + 3) * 13) % 97) * 0.37f;
    sbuf[i * blockDim.x + threadIdx.x] = v;
  }
  uint4* pbuf = reinterpret_cast<uint4*>(sbuf + (BN / 4 + 1) * blockDim.x);   // [BN/8][blockDim.x]
  float sv[BN];
  uint32_t slow_cnt = 0;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < BN / 4; ++i) {
      asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(sv[4 * i]), "=f"(sv[4 * i + 1]), "=f"(sv[4 * i + 2]), "=f"(sv[4 * i + 3])
                   : "r"(smem_u32(&sbuf[(i + (it & 1)) * blockDim.x + threadIdx.x])) : "memory");
    }
    uint32_t pk[BN / 2];
    bool slow = false;
    if (MODE == 0) {
      float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
      for (int i = 0; i < BN; i += 8) {
        mx0 = fmax3(mx0, sv[i + 0], sv[i + 1]);
        mx1 = fmax3(mx1, sv[i + 2], sv[i + 3]);
        mx2 = fmax3(mx2, sv[i + 4], sv[i + 5]);
        mx3 = fmax3(mx3, sv[i + 6], sv[i + 7]);
      }
      slow = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3)) > 8.f;
    } else if (MODE == 1) {
      float mx0 = 0.f, mx1 = 0.f, mx2 = 0.f, mx3 = 0.f;
#pragma unroll
      for (int i = 0; i < BN; i += 8) {
        mx0 = fmax3(mx0, fabsf(sv[i + 0]), fabsf(sv[i + 1]));
        mx1 = fmax3(mx1, fabsf(sv[i + 2]), fabsf(sv[i + 3]));
        mx2 = fmax3(mx2, fabsf(sv[i + 4]), fabsf(sv[i + 5]));
        mx3 = fmax3(mx3, fabsf(sv[i + 6]), fabsf(sv[i + 7]));
      }
      slow = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3)) > 67.f;
    }
    if (__any_sync(0xffffffffu, slow)) {
      ++slow_cnt;
      softmax_exp_row<BN, 0, 3, true, true, false>(sv, 1.f, 0.f, pk);
    } else {
      if (MODE == 0) softmax_exp_row<BN, EMU, DEG, true, true, false>(sv, 1.f, 0.f, pk);
      else           softmax_exp_row<BN, EMU, DEG, false, true, false>(sv, 1.f, 0.f, pk);
    }
#pragma unroll
    for (int i = 0; i < BN / 8; ++i) {
      asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(smem_u32(&pbuf[i * blockDim.x + threadIdx.x])), "r"(pk[4 * i]),
                   "r"(pk[4 * i + 1]), "r"(pk[4 * i + 2]), "r"(pk[4 * i + 3]) : "memory");
    }
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = __uint_as_float(slow_cnt + pbuf[threadIdx.x].x);
  if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

template <int EMU, int DEG, int MODE>
void run() {
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
  printf("mode %d EMU %d/8 deg %d:", MODE, EMU, DEG);
  for (int warps : {4, 8, 12, 16}) {
    cudaMemset(cyc, 0, 8);
    cudaFuncSetAttribute(k<EMU, DEG, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    k<EMU, DEG, MODE><<<148, warps * 32, warps * 32 * ((BN + 4) * 4 + BN * 2)>>>(out, cyc);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    double per_step = (double)c / ITERS;
    printf("  w/SMSP %d: %5.0f clk/step (%4.0f per warp-step)%s", warps / 4, per_step, per_step / (warps / 4), e == cudaSuccess ? "" : " (ERR)");
  }
  printf("\n");
  cudaFree(out); cudaFree(cyc);
}

int main() {
  run<3, 3, 0>();
  run<3, 3, 1>();
  run<3, 3, 2>();
  run<4, 3, 1>();
  run<2, 3, 1>();
  run<3, 2, 1>();
  run<4, 2, 1>();
  run<4, 2, 2>();
  run<0, 3, 1>();
  return 0;
}
