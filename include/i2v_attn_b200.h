/*
 * i2v_attn_b200.h — C ABI of libi2v_attn_b200.so: the B200 (sm_100a) attention hot path of
 * xUhEngwAng/I2V-Adapter-Unofficial.
 *
 * Every entry point replaces one library call the reference reaches through the diffusers AttnProcessor boundary
 * (SURVEY.md §8b).  Reference locations are relative to /root/reference:
 *
 *   i2v_sdpa_fwd              F.scaled_dot_product_attention inside AttnProcessor2_0 for attn1 / i2v_adapter
 *                             (called from src/modules/i2v_adapter.py:468-473 and :487-492)
 *   i2v_fused_self_xframe_fwd the pair attn1 (:468-473) + i2v_adapter (:484-492) of one I2VAdapterTransformerBlock
 *                             in a single launch; the frame-0 K/V are indexed in place instead of being
 *                             repeated num_frames times (einops.repeat at :485)
 *   i2v_ip_xattn_fwd          the two SDPA calls + scaled add of IPAdapterAttnProcessor2_0, installed at
 *                             src/models/unet_motion_cross_frame_attn.py:1264-1279, tokens built at :1346-1355
 *   i2v_temporal_attn_fwd     SDPA of the motion-module attn1/attn2 (TransformerTemporalModel constructed at
 *                             src/models/unet_motion_cross_frame_attn.py:232-244, called at :323-326)
 *   i2v_reshard_*             layout kernels of the frame partitioner (new functionality)
 *
 * Conventions
 *   - all pointers are DEVICE pointers on the current CUDA device; `stream` is a cudaStream_t (NULL = legacy default);
 *   - tensors are logical [batch, seq, heads, d] with d contiguous and element strides (stride_b, stride_s, stride_h),
 *     exactly the view the reference takes of a Linear output: `.view(B, -1, heads, head_dim)`;
 *   - work is only enqueued, never synchronised; outputs are written in full (never accumulated);
 *   - return value 0 on success, a negative i2v_status otherwise; i2v_last_error() gives the message for the calling
 *     thread.  Nothing aborts and nothing falls back to the CPU.
 */
#ifndef I2V_ATTN_B200_H
#define I2V_ATTN_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
  I2V_OK = 0,
  I2V_ERR_BAD_SHAPE = -1,       /* non-positive sizes, batch not divisible by kv_group / num_frames, ... */
  I2V_ERR_UNSUPPORTED = -2,     /* shape or dtype outside what the requested kernel family covers */
  I2V_ERR_MISALIGNED = -3,      /* pointer not 16-byte aligned or stride not a multiple of 8 elements */
  I2V_ERR_CUDA = -4,            /* a CUDA runtime / driver call failed (message has the CUDA error string) */
  I2V_ERR_BAD_DTYPE = -5,
  I2V_ERR_NO_DEVICE = -6        /* current device is not compute capability 10.x */
} i2v_status;

typedef enum { I2V_BF16 = 0, I2V_F32 = 1 } i2v_dtype;

/* Kernel family selection.
 *   AUTO   : tensor-core / bandwidth kernel when the shape is covered and dtype is bf16, else GENERIC
 *   FAST   : tensor-core / bandwidth kernel or I2V_ERR_UNSUPPORTED
 *   GENERIC: CUDA-core kernel with fp32 math ("fp32 check mode"; bf16 or fp32 storage) */
typedef enum { I2V_MODE_AUTO = 0, I2V_MODE_FAST = 1, I2V_MODE_GENERIC = 2 } i2v_mode;

typedef struct {
  void* data;
  int64_t stride_b; /* elements between consecutive batch entries   */
  int64_t stride_s; /* elements between consecutive sequence rows   */
  int64_t stride_h; /* elements between consecutive heads           */
} i2v_tensor;

int i2v_version(void);
const char* i2v_last_error(void);
/* 1 if the current device can run the sm_100a kernels, 0 otherwise (and sets the error string). */
int i2v_device_supported(void);
/* Number of kernels this library has launched in this process (evidence for bench.py's gpu_launches). */
int64_t i2v_launch_count(void);

/* o[b,:,h,:] = softmax(scale * q[b,:,h,:] k[b/kv_group,:,h,:]^T) v[b/kv_group,:,h,:]
 * k, v hold batch/kv_group entries.  kv_group = 1 is ordinary (self or cross) attention. */
int i2v_sdpa_fwd(const i2v_tensor* q, const i2v_tensor* k, const i2v_tensor* v, const i2v_tensor* o,
                 int batch, int heads, int sq, int skv, int d, int kv_group, float scale,
                 int dtype, int mode, void* stream);

/* Spatial self-attention and I2V-Adapter cross-frame attention of one block in one launch.
 *   o_self[b] = softmax(scale q_self[b] k_self[b]^T) v_self[b]                      b in [0, batch)
 *   o_x[b]    = softmax(scale q_x[b]    k_x[b/num_frames]^T) v_x[b/num_frames]
 * batch = videos * num_frames, frame index fastest (batch row = video * num_frames + frame, as produced by
 * src/models/unet_motion_cross_frame_attn.py:1358); k_x, v_x hold one entry per video (frame 0's projection). */
int i2v_fused_self_xframe_fwd(const i2v_tensor* q_self, const i2v_tensor* k_self, const i2v_tensor* v_self,
                              const i2v_tensor* o_self, const i2v_tensor* q_x, const i2v_tensor* k_x,
                              const i2v_tensor* v_x, const i2v_tensor* o_x,
                              int batch, int heads, int seq, int d, int num_frames, float scale,
                              int dtype, int mode, void* stream);

/* Same operator on the AUGMENTED operand layout, the layout the packed QKV projection of the fused block produces
 * (processors.py, B200SpatialAttnProcessor): head dim d = 40 stored padded to d_pad = 48 (strides say so), and
 *   q[..., 0:d] already multiplied by scale * log2(e) (folded into the projection weights), q[..., d:] = 0,
 *   k[..., d] = 1 and v[..., d] = 1 (a bias of the projection), k/v[..., d+1:] = 0.
 * The kernel keeps -rowmax in q's column d, so QK^T yields the softmax exponent directly, and reads the softmax
 * denominator from o's column d.  o_self / o_x receive d columns per head.  bf16 only; any other shape returns
 * I2V_ERR_UNSUPPORTED (callers then use i2v_fused_self_xframe_fwd). */
int i2v_fused_self_xframe_aug_fwd(const i2v_tensor* q_self, const i2v_tensor* k_self, const i2v_tensor* v_self,
                                  const i2v_tensor* o_self, const i2v_tensor* q_x, const i2v_tensor* k_x,
                                  const i2v_tensor* v_x, const i2v_tensor* o_x,
                                  int batch, int heads, int seq, int d, int d_pad, int num_frames,
                                  int dtype, void* stream);

/* IP-Adapter decoupled cross-attention:
 *   o = softmax(scale q k_txt^T) v_txt + ip_scale * softmax(scale q k_ip^T) v_ip
 * k/v tensors hold batch/kv_group entries (text and image tokens are identical for the frames of a video:
 * repeat_interleave at src/models/unet_motion_cross_frame_attn.py:1355).  The FAST path needs the image tokens to
 * directly follow the text tokens in memory (k_ip.data == &k_txt[.., n_txt, ..], same strides) and
 * n_txt + n_ip <= 128. */
int i2v_ip_xattn_fwd(const i2v_tensor* q, const i2v_tensor* k_txt, const i2v_tensor* v_txt,
                     const i2v_tensor* k_ip, const i2v_tensor* v_ip, const i2v_tensor* o,
                     int batch, int heads, int sq, int n_txt, int n_ip, int d, int kv_group,
                     float scale, float ip_scale, int dtype, int mode, void* stream);

/* Temporal self-attention over `frames` at every spatial position.
 * Tensors are logical [n_pos, frames, heads, d]: stride_b = position stride, stride_s = frame stride.
 * FAST path: bf16, frames <= 32, d in {16,32,40,64,80,128,160}, heads % 8 == 0, stride_h == d. */
int i2v_temporal_attn_fwd(const i2v_tensor* q, const i2v_tensor* k, const i2v_tensor* v, const i2v_tensor* o,
                          int n_pos, int heads, int frames, int d, float scale,
                          int dtype, int mode, void* stream);

/* Frame partitioner layout kernels (new functionality: the reference has no inference-time
 * sharding).  A rank holding f_local frames of every video re-shards [videos, f_local, S, C] activations to
 * [videos, F, S/G, C] (all frames, 1/G of the spatial positions) with one NCCL all-to-all between two copies:
 *   pack   (inverse = 0): send[r][v, f, s, c]      = x[v, f, r*S/G + s, c]          r in [0, G)
 *   unpack (inverse = 0): y[v, g*f_local + f, s, c] = recv[g][v, f, s, c]            g in [0, G)
 * inverse = 1 runs the same index maps backwards (src and dst swap roles) for the way back.
 * elem_bytes is 2 or 4; seq must be divisible by world; channels*elem_bytes must be a multiple of 16. */
int i2v_reshard_pack(const void* src, void* dst, int videos, int f_local, int seq, int channels, int world,
                     int elem_bytes, int inverse, void* stream);
int i2v_reshard_unpack(const void* src, void* dst, int videos, int f_local, int seq_local, int channels, int world,
                       int elem_bytes, int inverse, void* stream);

/* ---- normalisation prologues, layout changes and the residual epilogue around the attention operators ----
 * (bf16 only; the reference runs them as separate PyTorch kernels: GroupNorm, permute+reshape copies, LayerNorm,
 * `+ pos_embed`, GEGLU, residual add — src/modules/i2v_adapter.py:214-234, 298-314, 445-459, 514-525, 539-561 and
 * diffusers TransformerTemporalModel.forward)
 *
 * i2v_layernorm_fwd: y[r, :] = LayerNorm(x[r, :]) * w + b (+ pe[r % pe_rows, :] when pe != NULL).  x, y: [rows, C]
 *   contiguous, C % 8 == 0, C <= 2048.  pe: [pe_rows, C] (the sinusoidal table `pos_embed.pe[0, :F]`), rows ordered so
 *   that r % pe_rows is the frame index ([B*S, F, C] of the motion module). */
int i2v_layernorm_fwd(const void* x, const void* w, const void* b, const void* pe, void* y, long long rows, int C,
                      int pe_rows, float eps, void* stream);
/* i2v_layernorm_pre_fwd: the same with x := x + pre[c] first (pre: [C] or NULL) -- the per-channel constant a producer
 *   GEMM left out, see fastpath.py: output-projection biases are deferred so the residual add rides in the GEMM. */
int i2v_layernorm_pre_fwd(const void* x, const void* pre, const void* w, const void* b, const void* pe, void* y,
                          long long rows, int C, int pe_rows, float eps, void* stream);
/* i2v_geglu_fwd: y[r, c] = x[r, c] * gelu(x[r, D + c]) (erf GELU).  x: [rows, 2*D], y: [rows, D], D % 8 == 0. */
int i2v_geglu_fwd(const void* x, void* y, long long rows, int D, void* stream);
/* i2v_geglu_ld_fwd: y has row pitch ld_out = D or D + 8; with D + 8 the extra columns are (1, 0, .., 0): a ones column
 *   so that the following GEMM can carry its bias as one more weight column (and its residual as the beta = 1 term). */
int i2v_geglu_ld_fwd(const void* x, void* y, long long rows, int D, int ld_out, void* stream);
/* i2v_ff_geglu_fwd: the GEGLU feed-forward's input projection fused with the activation (diffusers FeedForward
 *   net[0] = GEGLU(dim, 4 dim): proj = Linear(dim, 8 dim), hidden, gate = proj(x).chunk(2), hidden * gelu(gate);
 *   reference call site src/modules/i2v_adapter.py:535-561):
 *     y[r, c] = (x W[c]^T + bias[c]) * gelu(x W[N + c]^T + bias[N + c]),   c < N
 *   x: [rows, K] bf16 row-major, w: [2N, K] (the nn.Linear weight as stored), bias: [2N] or NULL, y: [rows, ld_out] with
 *   ld_out = N or N + 8 (ones column as in i2v_geglu_ld_fwd).  K % 64 == 0, N % 128 == 0.  One tcgen05 GEMM whose
 *   accumulator tile pairs hidden and gate columns, so the [rows, 2N] projection never goes to HBM. */
int i2v_ff_geglu_fwd(const void* x, const void* w, const void* bias, void* y, long long rows, int K, int N, int ld_out,
                     void* stream);
/* GroupNorm + layout change in two passes over x [N, C, S] (NCHW, S = h*w), G groups, statistics shared by `fg`
 * consecutive batch entries (fg = 1: the spatial transformer's per-frame GroupNorm; fg = num_frames: the motion module's
 * GroupNorm over (C/G, F, h, w) per video, N = videos * fg, frame index fastest):
 *   i2v_gn_stats:            partial[n, g] = (sum, sum of squares) of x[n, g*C/G : (g+1)*C/G, :]   (fp32, [N, G, 2])
 *   i2v_gn_apply_transpose:  out[((v*S + s)*fg + f)*C + c] = GroupNorm(x)[v*fg + f, c, s]  -> (N, S, C) or (V*S, F, C)
 *   i2v_untranspose_residual: out[n, c, s] = y[((v*S + s)*fg + f)*C + c] + res[n, c, s]    (the way back + residual)
 * C % 64 == 0, S % 8 == 0, (C/G) % 2 == 0, (C/G)*S % 8 == 0. */
int i2v_gn_stats(const void* x, float* partial, int N, int C, int S, int G, void* stream);
int i2v_gn_apply_transpose(const void* x, const float* partial, const void* w, const void* b, void* out, int N, int C,
                           int S, int G, int fg, float eps, void* stream);
int i2v_untranspose_residual(const void* y, const void* res, void* out, int N, int C, int S, int fg, void* stream);

/* Channels-last (NHWC) GroupNorm family: x is [N, S, C] bf16, i.e. an (N, C, h, w) activation in
 * torch.channels_last -- the format the reference's convolutions run in on the GPU and, at the same time, the
 * token-major layout of the transformer blocks.
 *   out = GroupNorm_G(x + add[n, c]) * w + b, optionally followed by SiLU (silu = 1);
 * statistics are shared by the fg consecutive images of a video (fg = 1: nn.GroupNorm on (N, C, h, w), reference
 * src/modules/i2v_adapter.py:218 and the ResnetBlock2D norms; fg = num_frames: the motion module's GroupNorm over
 * (B, C, F, h, w)).  add (may be NULL) is ResnetBlock2D's time-embedding term added before norm2.
 * perm = 0: out rows keep x's order; perm = 1: out is [N/fg, S, fg, C] (position-major rows for the motion module).
 * scratch: i2v_gn_nhwc_scratch_floats(N, G) floats of device memory. */
long long i2v_gn_nhwc_scratch_floats(int N, int G);
int i2v_gn_nhwc(const void* x, const void* add, const void* w, const void* b, void* out, float* scratch,
                int N, int S, int C, int G, int fg, float eps, int silu, int perm, void* stream);
/* i2v_gn_nhwc_cat: the same on the virtual channel concatenation of two channels-last activations x [N, S, C1] and
 *   x2 [N, S, C - C1] (x2 == NULL: plain i2v_gn_nhwc): the up blocks' torch.cat([hidden_states, res_hidden_states], 1)
 *   ahead of their ResnetBlock2D (src/models/unet_motion_cross_frame_attn.py:457) is never written to memory. */
int i2v_gn_nhwc_cat(const void* x, const void* x2, int C1, const void* add, const void* w, const void* b, void* out,
                    float* scratch, int N, int S, int C, int G, int fg, float eps, int silu, int perm, void* stream);

/* out[n, s, :] = y[(v*S + s)*fg + f, :] + res[n, s, :] with n = v*fg + f: the motion module's way back to the
 * frame-major channels-last activation, fused with its residual add. */
int i2v_rows_residual(const void* y, const void* res, void* out, int N, int S, int C, int fg, void* stream);
/* ... + bias[c] (NULL allowed): out = (y + bias) + res.  With fg = 1 this is ResnetBlock2D's `input + conv2(...)` with
 * conv2's bias folded in (cuDNN would spend a separate broadcast-add pass on it). */
int i2v_rows_residual_bias(const void* y, const void* res, const void* bias, void* out, int N, int S, int C, int fg,
                           void* stream);
/* i2v_upsample2x_nhwc: nearest-neighbour 2x upsampling of a channels-last [N, h, w, C] activation into [N, 2h, 2w, C]
 *   (Upsample2D's interpolate ahead of its convolution; C % 8 == 0).  With res == NULL, i2v_rows_residual_bias is the
 *   in-place-capable per-channel bias add out = y + bias (a convolution bias cuDNN would apply as a separate pass). */
int i2v_upsample2x_nhwc(const void* x, void* out, int N, int h, int w, int C, void* stream);

/* i2v_linear_fwd: out[r, n] = sum_k x[r, k] W[n, k] (+ bias[n]) (+ res[r, n]) -- `F.linear` with the residual add in
 *   the epilogue, for the projections around the attention operators that the reference runs as separate nn.Linear
 *   calls: to_q / to_k / to_v (src/modules/i2v_adapter.py:468-492 through diffusers Attention), to_out + the residual
 *   add (:494-501, :533), the feed-forward's output Linear + residual (:554-561), the motion module's proj_in / proj_out.
 *   bf16; x [rows, K] with row pitch ld_x, W [N, K] contiguous (nn.Linear layout), out / res with pitches ld_out /
 *   ld_res; K, N and the pitches multiples of 8; bias, res may be NULL; res may alias out.  tcgen05 2-SM MMA. */
int i2v_linear_fwd(const void* x, const void* w, const void* bias, const void* res, void* out, long long rows, int K,
                   int N, int ld_x, int ld_res, int ld_out, void* stream);

/* Frame-sharded motion module (partition.FramePartitioner.temporal_forward; new functionality): the
 * GroupNorm of TransformerTemporalModel (constructed at src/models/unet_motion_cross_frame_attn.py:232-244) spans all
 * F frames of a video, of which a rank holds fg = F / W.
 *   i2v_gn_nhwc_sums   : raw per-(video, group) (sum, sum of squares) of the local frames -> sums [N / fg, G, 2]
 *                        (the caller all-reduces them over the W ranks and forms mean / rstd);
 *   i2v_gn_nhwc_apply  : the apply pass of i2v_gn_nhwc with caller-provided (mean, rstd) [N / fg, G, 2]; perm = 2 writes
 *                        the all-to-all send buffer [W, V, S / W, fg, C] (position-major rows grouped by destination
 *                        rank) directly;
 *   i2v_rows_residual_sharded : the way back -- y is the all-to-all receive buffer in that same layout; out = y + res
 *                        on the channels-last [N, S, C] activation. */
int i2v_gn_nhwc_sums(const void* x, float* sums, float* scratch, int N, int S, int C, int G, int fg, void* stream);
int i2v_gn_nhwc_apply(const void* x, const void* add, const void* w, const void* b, void* out, const float* mean_rstd,
                      int N, int S, int C, int G, int fg, int silu, int perm, int world, void* stream);
int i2v_rows_residual_sharded(const void* y, const void* res, void* out, int N, int S, int C, int fg, int world,
                              void* stream);

/* Tuning knobs for experiments (0 = library default; keys 0..11).  key 0: temporal stages, key 1: temporal CTAs per SM,
 * key 2: dense-attention exp2 split + 1 (pairs out of 8 computed on the FMA pipe instead of MUFU),
 * key 3: dense-attention tile variant + 1 for head dims <= 48; at d = 80: 1 = the first tcgen05 kernel instead of the
 * pipelined one, 6 = three query tiles x 48 keys (see capi.cu), key 4: temporal rows per warp, key 5: IP-Adapter
 * attention -- 1 = tcgen05 single-tile dense kernel, 4 = streaming kernel even where the resident-K/V tcgen05 kernel
 * applies, key 6: token-GEMM tile width / streaming-kernel configuration, key 7: variants of the augmented-layout dense
 * kernel (profiles/r02_dense_experiments.md), key 9: 1 = channels-last GroupNorm with the separate finalize launch. */
int i2v_set_tuning(int key, int value);

/* Launch timing for bench.py ("the dominant kernel timed live with CUDA events on the launching
 * stream").  i2v_prof_arm(kind, a, b, n): the next n launches of kernel class `kind` whose shape matches (a, b)
 * (0 = any) are bracketed by a cudaEvent pair recorded by the library right around the launch (after tensor-map
 * encoding; while the stream is captured into a CUDA graph the records become external event nodes, re-recorded by
 * every replay).  n <= 0 disarms.  i2v_prof_read(kind, ms, cap): elapsed ms of the recorded pairs (the caller has
 * synchronised); returns their number.
 *   kind 1: dense attention (a = query rows per batch entry, b = batch)      2: temporal (a = positions, b = d)
 *   kind 3: IP-Adapter attention (a = query rows, b = batch)                  4: token GEMMs (a = rows, b = N) */
int i2v_prof_arm(int kind, long long match_a, long long match_b, int max_pairs);
int i2v_prof_read(int kind, float* ms_out, int capacity);

#ifdef __cplusplus
}
#endif
#endif /* I2V_ATTN_B200_H */
