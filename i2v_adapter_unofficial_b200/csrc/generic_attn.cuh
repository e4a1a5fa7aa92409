// Generic scaled-dot-product attention on CUDA cores with fp32 math (inputs bf16 or fp32).
// This is the "fp32 check mode" of BASELINE.json's north_star (<= 1e-3 relative) and the catch-all for shapes the
// tensor-core kernels do not cover (head dims that are not multiples of 8, additive masks are not supported).
// It follows diffusers AttnProcessor2_0 semantics restated in SURVEY.md §3c: softmax(scale * q k^T) v, no mask,
// no dropout.  K/V batch index = b / kv_group (kv_group = num_frames gives the cross-frame attention of
// src/modules/i2v_adapter.py:484-492 without materialising the repeated first frame).
// Optional second key/value segment with its own softmax, scaled and added (IPAdapterAttnProcessor2_0).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace i2v {

struct GenericParams {
  const void* q; const void* k; const void* v; void* o;
  long long q_sb, q_ss, q_sh;
  long long k_sb, k_ss, k_sh;
  long long v_sb, v_ss, v_sh;
  long long o_sb, o_ss, o_sh;
  int batch, heads, sq, skv, d;
  int kv_group;
  float scale;
  // second segment (IP-Adapter image tokens); skv2 == 0 disables it
  const void* k2; const void* v2;
  long long k2_sb, k2_ss, k2_sh;
  long long v2_sb, v2_ss, v2_sh;
  int skv2;
  float scale2;  // weight of the second segment's output
};

template <typename T> __device__ __forceinline__ float ld_as_float(const T* p);
template <> __device__ __forceinline__ float ld_as_float<float>(const float* p) { return *p; }
template <> __device__ __forceinline__ float ld_as_float<__nv_bfloat16>(const __nv_bfloat16* p) {
  return __bfloat162float(*p);
}
template <typename T> __device__ __forceinline__ void st_from_float(T* p, float x);
template <> __device__ __forceinline__ void st_from_float<float>(float* p, float x) { *p = x; }
template <> __device__ __forceinline__ void st_from_float<__nv_bfloat16>(__nv_bfloat16* p, float x) {
  *p = __float2bfloat16_rn(x);
}

constexpr int kGenRows = 32;    // query rows per CTA
constexpr int kGenSlices = 4;   // threads cooperating on one row (head dim split)
constexpr int kGenKeys = 32;    // keys per shared-memory tile

// One CTA: 32 query rows of one (batch, head); thread (r, sl) owns elements sl, sl+4, sl+8, ... of row r.
template <typename T, int DS>
__device__ __forceinline__ void generic_segment(const GenericParams& P, const T* kbase, const T* vbase,
                                                long long k_ss, long long v_ss, int skv, const float* qreg,
                                                float* acc, float& m_run, float& l_run, float* sk, float* svm,
                                                int nds, int sl) {
  const int tid = threadIdx.x;
  const int d = P.d;
  for (int j0 = 0; j0 < skv; j0 += kGenKeys) {
    const int nk = min(kGenKeys, skv - j0);
    __syncthreads();
    for (int i = tid; i < nk * d; i += blockDim.x) {
      const int kr = i / d, kc = i - kr * d;
      sk[kr * (d + 1) + kc] = ld_as_float(kbase + (long long)(j0 + kr) * k_ss + kc);
      svm[kr * (d + 1) + kc] = ld_as_float(vbase + (long long)(j0 + kr) * v_ss + kc);
    }
    __syncthreads();
    float s[kGenKeys];
    float tile_max = -INFINITY;
#pragma unroll 4
    for (int kr = 0; kr < kGenKeys; ++kr) {
      float dot = 0.f;
      if (kr < nk) {
        #pragma unroll
        for (int i = 0; i < DS; ++i)
          if (i < nds) dot = fmaf(qreg[i], sk[kr * (d + 1) + sl + 4 * i], dot);
      }
      dot += __shfl_xor_sync(0xffffffffu, dot, 1);
      dot += __shfl_xor_sync(0xffffffffu, dot, 2);
      dot = (kr < nk) ? dot * P.scale : -INFINITY;
      s[kr] = dot;
      tile_max = fmaxf(tile_max, dot);
    }
    const float m_new = fmaxf(m_run, tile_max);
    const float alpha = (m_run == -INFINITY) ? 0.f : expf(m_run - m_new);
    l_run *= alpha;
#pragma unroll
    for (int i = 0; i < DS; ++i) acc[i] *= alpha;
#pragma unroll 4
    for (int kr = 0; kr < kGenKeys; ++kr) {
      if (kr < nk) {
        const float p = expf(s[kr] - m_new);
        l_run += p;
#pragma unroll
        for (int i = 0; i < DS; ++i)
          if (i < nds) acc[i] = fmaf(p, svm[kr * (d + 1) + sl + 4 * i], acc[i]);
      }
    }
    m_run = m_new;
  }
}

template <typename T, int DS>
__global__ void __launch_bounds__(kGenRows * kGenSlices) generic_attn_kernel(const GenericParams P) {
  extern __shared__ float gsm[];
  const int d = P.d;
  float* sk = gsm;                          // [kGenKeys][d+1]
  float* svm = gsm + kGenKeys * (d + 1);    // [kGenKeys][d+1]
  const int tid = threadIdx.x;
  const int r = tid >> 2, sl = tid & 3;
  const int qblocks = (P.sq + kGenRows - 1) / kGenRows;
  long long x = blockIdx.x;
  const int qb = (int)(x % qblocks); x /= qblocks;
  const int h = (int)(x % P.heads); x /= P.heads;
  const int b = (int)x;
  const int bkv = b / P.kv_group;
  const int row = qb * kGenRows + r;
  const bool row_ok = row < P.sq;
  const int nds = (d - sl + 3) / 4;  // number of elements sl, sl+4, ... < d   (host guarantees nds <= DS)

  const T* qrow = reinterpret_cast<const T*>(P.q) + b * P.q_sb + (long long)min(row, P.sq - 1) * P.q_ss + h * P.q_sh;
  float qreg[DS], acc[DS], out[DS];
#pragma unroll
  for (int i = 0; i < DS; ++i) {
    qreg[i] = (i < nds) ? ld_as_float(qrow + sl + 4 * i) : 0.f;
    acc[i] = 0.f;
    out[i] = 0.f;
  }
  {
    float m_run = -INFINITY, l_run = 0.f;
    const T* kb = reinterpret_cast<const T*>(P.k) + bkv * P.k_sb + h * P.k_sh;
    const T* vb = reinterpret_cast<const T*>(P.v) + bkv * P.v_sb + h * P.v_sh;
    generic_segment<T, DS>(P, kb, vb, P.k_ss, P.v_ss, P.skv, qreg, acc, m_run, l_run, sk, svm, nds, sl);
    const float inv = 1.f / l_run;
#pragma unroll
    for (int i = 0; i < DS; ++i) out[i] = acc[i] * inv;
  }
  if (P.skv2 > 0) {
#pragma unroll
    for (int i = 0; i < DS; ++i) acc[i] = 0.f;
    float m_run = -INFINITY, l_run = 0.f;
    const T* kb = reinterpret_cast<const T*>(P.k2) + bkv * P.k2_sb + h * P.k2_sh;
    const T* vb = reinterpret_cast<const T*>(P.v2) + bkv * P.v2_sb + h * P.v2_sh;
    generic_segment<T, DS>(P, kb, vb, P.k2_ss, P.v2_ss, P.skv2, qreg, acc, m_run, l_run, sk, svm, nds, sl);
    const float inv = P.scale2 / l_run;
#pragma unroll
    for (int i = 0; i < DS; ++i) out[i] = fmaf(acc[i], inv, out[i]);
  }
  if (row_ok) {
    T* orow = reinterpret_cast<T*>(P.o) + b * P.o_sb + (long long)row * P.o_ss + h * P.o_sh;
#pragma unroll
    for (int i = 0; i < DS; ++i)
      if (i < nds) st_from_float(orow + sl + 4 * i, out[i]);
  }
}

}  // namespace i2v
