// Bandwidth kernels around the attention operators: the normalisation prologues, layout changes and the residual
// epilogue that the reference performs as separate PyTorch passes (GroupNorm, permute+reshape copies, LayerNorm,
// positional-embedding add, GEGLU, residual add).  Reference call sites:
//   * I2VAdapterTransformer2DModel.forward  src/modules/i2v_adapter.py:214-234 (GroupNorm -> proj_in -> (BF,C,h,w) to
//     (BF,S,C)) and :298-314 (back to (BF,C,h,w), + residual)
//   * diffusers TransformerTemporalModel.forward: GroupNorm over (C/G, F, h, w) per video,
//     (BF,C,h,w) -> (B*S, F, C), ..., back, + residual
//   * LayerNorm norm1/norm2/norm3 and `+ pos_embed` of the transformer blocks (src/modules/i2v_adapter.py:445-459,
//     514-525, 539) and the GEGLU of the feed-forward (:554, diffusers GEGLU: proj(x).chunk(2) -> h * gelu(gate))
// All are HBM-bound: 16-byte vector accesses on the contiguous axis, fp32 arithmetic, one read and one write of the
// activation per kernel (the GroupNorm needs one extra read for its statistics).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace i2v {

__device__ __forceinline__ float bf16_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }
__device__ __forceinline__ uint32_t bf16_pack(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ---------------------------------------------------------------------------------------------------------------
// LayerNorm over the last axis (+ optional additive row table, the sinusoidal positional embedding):
//   y[r, :] = (x[r, :] - mean) * rstd * w + b  (+ pe[r % pe_rows, :])
// One warp per row, the row cached in registers (two-pass variance), MAXV 16-byte vectors per lane.
// ---------------------------------------------------------------------------------------------------------------
template <int MAXV>
__global__ void __launch_bounds__(256) layernorm_kernel(const uint4* __restrict__ x, uint4* __restrict__ y,
                                                        const uint4* __restrict__ w, const uint4* __restrict__ b,
                                                        const uint4* __restrict__ pe, int pe_rows, long long rows,
                                                        int nvec /* C / 8 */, float eps) {
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const uint4* xr = x + row * nvec;
  uint4 v[MAXV];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int c = lane + 32 * i;
    if (c < nvec) {
      v[i] = xr[c];
      s += bf16_lo(v[i].x) + bf16_hi(v[i].x) + bf16_lo(v[i].y) + bf16_hi(v[i].y) + bf16_lo(v[i].z) + bf16_hi(v[i].z) +
           bf16_lo(v[i].w) + bf16_hi(v[i].w);
    }
  }
  const float inv_n = 1.f / (float)(nvec * 8);
  const float mean = warp_sum(s) * inv_n;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int c = lane + 32 * i;
    if (c < nvec) {
      const uint32_t ww[4] = {v[i].x, v[i].y, v[i].z, v[i].w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float a = bf16_lo(ww[k]) - mean, bb = bf16_hi(ww[k]) - mean;
        q += a * a + bb * bb;
      }
    }
  }
  const float rstd = rsqrtf(warp_sum(q) * inv_n + eps);
  uint4* yr = y + row * nvec;
  const uint4* per = pe ? pe + (row % pe_rows) * nvec : nullptr;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int c = lane + 32 * i;
    if (c < nvec) {
      const uint4 wv = w[c], bv = b[c];
      const uint32_t xin[4] = {v[i].x, v[i].y, v[i].z, v[i].w};
      const uint32_t win[4] = {wv.x, wv.y, wv.z, wv.w};
      const uint32_t bin[4] = {bv.x, bv.y, bv.z, bv.w};
      uint32_t pin[4] = {0u, 0u, 0u, 0u};
      if (per) {
        const uint4 pv = per[c];
        pin[0] = pv.x; pin[1] = pv.y; pin[2] = pv.z; pin[3] = pv.w;
      }
      uint32_t out[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        // LayerNorm result is rounded to bf16 before the embedding is added, as in the two-op reference sequence
        float lo = (bf16_lo(xin[k]) - mean) * rstd * bf16_lo(win[k]) + bf16_lo(bin[k]);
        float hi = (bf16_hi(xin[k]) - mean) * rstd * bf16_hi(win[k]) + bf16_hi(bin[k]);
        if (per) {
          const uint32_t r = bf16_pack(lo, hi);
          lo = bf16_lo(r) + bf16_lo(pin[k]);
          hi = bf16_hi(r) + bf16_hi(pin[k]);
        }
        out[k] = bf16_pack(lo, hi);
      }
      yr[c] = make_uint4(out[0], out[1], out[2], out[3]);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// GEGLU: y[r, c] = x[r, c] * gelu(x[r, D + c]), exact (erf) GELU.  x is [rows, 2D], y is [rows, D].
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float gelu_erf(float g) { return 0.5f * g * (1.f + erff(g * 0.70710678118654752f)); }

__global__ void __launch_bounds__(256) geglu_kernel(const uint4* __restrict__ x, uint4* __restrict__ y, long long rows,
                                                    int dvec /* D / 8 */) {
  const long long total = rows * dvec;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / dvec;
    const int c = (int)(i - r * dvec);
    const uint4 h = x[r * 2 * dvec + c];
    const uint4 g = x[r * 2 * dvec + dvec + c];
    const uint32_t hin[4] = {h.x, h.y, h.z, h.w}, gin[4] = {g.x, g.y, g.z, g.w};
    uint32_t out[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      // the reference rounds gelu(gate) to bf16 before the product (two PyTorch ops): do the same
      const uint32_t ge = bf16_pack(gelu_erf(bf16_lo(gin[k])), gelu_erf(bf16_hi(gin[k])));
      out[k] = bf16_pack(bf16_lo(hin[k]) * bf16_lo(ge), bf16_hi(hin[k]) * bf16_hi(ge));
    }
    y[r * dvec + c] = make_uint4(out[0], out[1], out[2], out[3]);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// GroupNorm statistics, first pass: x is [N, C, S] (NCHW, S = h*w contiguous); one CTA reduces the contiguous
// (C/G)*S slab of one (n, group) to (sum, sum of squares).  partial is [N, G, 2] fp32.
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) gn_partial_stats_kernel(const uint4* __restrict__ x, float* __restrict__ partial,
                                                               long long slab_vec /* (C/G)*S/8 */) {
  const uint4* p = x + (long long)blockIdx.x * slab_vec;
  float s = 0.f, q = 0.f;
  for (long long i = threadIdx.x; i < slab_vec; i += blockDim.x) {
    const uint4 v = p[i];
    const uint32_t ww[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float a = bf16_lo(ww[k]), b = bf16_hi(ww[k]);
      s += a + b;
      q += a * a + b * b;
    }
  }
  __shared__ float red[2][8];
  s = warp_sum(s);
  q = warp_sum(q);
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = s; red[1][threadIdx.x >> 5] = q; }
  __syncthreads();
  if (threadIdx.x < 32) {
    float a = threadIdx.x < 8 ? red[0][threadIdx.x] : 0.f;
    float b = threadIdx.x < 8 ? red[1][threadIdx.x] : 0.f;
    a = warp_sum(a);
    b = warp_sum(b);
    if (threadIdx.x == 0) { partial[2 * blockIdx.x] = a; partial[2 * blockIdx.x + 1] = b; }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// GroupNorm apply + layout change, second pass.  n = v * fg + f  (fg = frames that share statistics: 1 for the
// spatial transformer, num_frames for the motion module):
//   out[((v*S + s) * fg + f) * C + c] = (x[n, c, s] - mean[v, g]) * rstd[v, g] * w[c] + b[c]
// i.e. (BF, C, h, w) -> (BF, S, C) for fg = 1 and -> (B*S, F, C) for fg = F.  64 channels x 64 positions per CTA;
// loads are 16-byte vectors along s, stores 128-byte rows along c, the transpose goes through shared memory as
// 32-bit words holding a channel pair (row pitch 33 words: at most 2-way bank conflicts).
// Requires C % 64 == 0, S % 8 == 0, (C/G) % 2 == 0.
// ---------------------------------------------------------------------------------------------------------------
struct GnApplyParams {
  const __nv_bfloat16* x;
  __nv_bfloat16* out;
  const float* partial;        // [N, G, 2]
  const __nv_bfloat16* w;
  const __nv_bfloat16* b;
  int N, C, S, G, fg;
  float eps;
};

__global__ void __launch_bounds__(256) gn_apply_transpose_kernel(const GnApplyParams P) {
  __shared__ uint32_t tile[64 * 33];
  const int s0 = blockIdx.x * 64, c0 = blockIdx.y * 64, n = blockIdx.z;
  const int v = n / P.fg, f = n - v * P.fg;
  const int cg = P.C / P.G;
  const int t = threadIdx.x;
  const int cpair = t >> 3, svec = t & 7;          // 32 channel pairs x 8 vectors of 8 positions
  const int c = c0 + 2 * cpair;
  // statistics of the (video, group) this channel pair belongs to (a pair never straddles groups: cg is even)
  const int g = c / cg;
  float sum = 0.f, sq = 0.f;
  for (int ff = 0; ff < P.fg; ++ff) {
    const float* pp = P.partial + ((long long)(v * P.fg + ff) * P.G + g) * 2;
    sum += pp[0];
    sq += pp[1];
  }
  const float cnt = (float)P.fg * (float)cg * (float)P.S;
  const float mean = sum / cnt;
  const float rstd = rsqrtf(fmaxf(sq / cnt - mean * mean, 0.f) + P.eps);
  const float a0 = rstd * __bfloat162float(P.w[c]), a1 = rstd * __bfloat162float(P.w[c + 1]);
  const float b0 = __bfloat162float(P.b[c]) - mean * a0, b1 = __bfloat162float(P.b[c + 1]) - mean * a1;

  const bool in_range = s0 + svec * 8 < P.S;  // S is a multiple of 8: a vector is entirely inside or outside
  uint4 r0 = make_uint4(0u, 0u, 0u, 0u), r1 = r0;
  if (in_range) {
    r0 = *reinterpret_cast<const uint4*>(P.x + ((long long)n * P.C + c) * P.S + s0 + svec * 8);
    r1 = *reinterpret_cast<const uint4*>(P.x + ((long long)n * P.C + c + 1) * P.S + s0 + svec * 8);
  }
  const uint32_t x0[4] = {r0.x, r0.y, r0.z, r0.w}, x1[4] = {r1.x, r1.y, r1.z, r1.w};
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    tile[(svec * 8 + 2 * k) * 33 + cpair] = bf16_pack(fmaf(bf16_lo(x0[k]), a0, b0), fmaf(bf16_lo(x1[k]), a1, b1));
    tile[(svec * 8 + 2 * k + 1) * 33 + cpair] = bf16_pack(fmaf(bf16_hi(x0[k]), a0, b0), fmaf(bf16_hi(x1[k]), a1, b1));
  }
  __syncthreads();
  const int warp = t >> 5, lane = t & 31;
  uint32_t* outw = reinterpret_cast<uint32_t*>(P.out);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int s = warp * 8 + i;
    if (s0 + s < P.S) {
      const long long row = ((long long)v * P.S + s0 + s) * P.fg + f;
      outw[(row * P.C + c0) / 2 + lane] = tile[s * 33 + lane];
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Inverse layout change + residual add:  out[n, c, s] = y[((v*S + s) * fg + f) * C + c] + res[n, c, s]
// ---------------------------------------------------------------------------------------------------------------
struct UntransposeParams {
  const __nv_bfloat16* y;    // token-major
  const __nv_bfloat16* res;  // [N, C, S]
  __nv_bfloat16* out;        // [N, C, S]
  int N, C, S, fg;
};

__global__ void __launch_bounds__(256) untranspose_residual_kernel(const UntransposeParams P) {
  __shared__ uint32_t tile[64 * 33];
  const int s0 = blockIdx.x * 64, c0 = blockIdx.y * 64, n = blockIdx.z;
  const int v = n / P.fg, f = n - v * P.fg;
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const uint32_t* yw = reinterpret_cast<const uint32_t*>(P.y);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int s = warp * 8 + i;
    if (s0 + s < P.S) {
      const long long row = ((long long)v * P.S + s0 + s) * P.fg + f;
      tile[s * 33 + lane] = yw[(row * P.C + c0) / 2 + lane];
    }
  }
  __syncthreads();
  const int cpair = t >> 3, svec = t & 7;
  const int c = c0 + 2 * cpair;
  if (s0 + svec * 8 >= P.S) return;
  const long long o0 = ((long long)n * P.C + c) * P.S + s0 + svec * 8;
  const uint4 r0 = *reinterpret_cast<const uint4*>(P.res + o0);
  const uint4 r1 = *reinterpret_cast<const uint4*>(P.res + o0 + P.S);
  const uint32_t q0[4] = {r0.x, r0.y, r0.z, r0.w}, q1[4] = {r1.x, r1.y, r1.z, r1.w};
  uint32_t e0[4], e1[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const uint32_t wa = tile[(svec * 8 + 2 * k) * 33 + cpair];      // position 2k:   (c, c+1)
    const uint32_t wb = tile[(svec * 8 + 2 * k + 1) * 33 + cpair];  // position 2k+1: (c, c+1)
    e0[k] = bf16_pack(bf16_lo(wa) + bf16_lo(q0[k]), bf16_lo(wb) + bf16_hi(q0[k]));
    e1[k] = bf16_pack(bf16_hi(wa) + bf16_lo(q1[k]), bf16_hi(wb) + bf16_hi(q1[k]));
  }
  *reinterpret_cast<uint4*>(P.out + o0) = make_uint4(e0[0], e0[1], e0[2], e0[3]);
  *reinterpret_cast<uint4*>(P.out + o0 + P.S) = make_uint4(e1[0], e1[1], e1[2], e1[3]);
}

}  // namespace i2v
