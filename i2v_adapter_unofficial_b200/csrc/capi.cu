// C-ABI entry points of libi2v_attn_b200.so (declared in include/i2v_attn_b200.h): argument validation,
// TMA tensor-map encoding, kernel selection and launch.  No torch types, no CPU compute path.
#include <cuda.h>
#include <cuda_runtime.h>
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <cstdlib>

#include "../../include/i2v_attn_b200.h"
#include "dense_attn_sm100.cuh"
#include "dense_attn_pipe_sm100.cuh"
#include "ip_xattn_stream.cuh"
#include "ip_xattn_tc_sm100.cuh"
#include "generic_attn.cuh"
#include "norm_layout.cuh"
#include "temporal_attn.cuh"
#include "ff_geglu_gemm_sm100.cuh"
#include "tok_gemm_sm100.cuh"

namespace {

thread_local char g_err[512] = "";
std::atomic<long long> g_launches{0};
int g_tuning[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
constexpr int kGnMaxChunks = 64;   // row chunks per image of the channels-last GroupNorm statistics pass
#ifdef I2V_TRACE
unsigned long long* g_trace = nullptr;
int g_trace_cta = 0;
#endif

int fail(int code, const char* fmt, ...);

// ---------------------------------------------------------------------------------------------------------
// Launch timing (i2v_prof_*): CUDA-event pairs recorded by the library immediately around the kernel launch, after
// the host-side tensor-map encoding, on the launching stream.  While the stream is being captured into a CUDA graph
// the records become external event-record nodes, so every replay of the graph re-records them and the caller reads
// the durations of the kernels *inside the replayed step* after synchronising.
// ---------------------------------------------------------------------------------------------------------
constexpr int kProfKinds = 6;        // 1 dense fused (level-0 class), 2 temporal, 3 IP-Adapter, 4 feed-forward GEMM, 5 token GEMM
constexpr int kProfMaxPairs = 256;
struct ProfKind {
  bool armed = false;
  long long match_a = 0, match_b = 0;   // 0 = any
  int cap = 0, used = 0;
  cudaEvent_t e0[kProfMaxPairs], e1[kProfMaxPairs];
  int created = 0;
};
ProfKind g_prof[kProfKinds];

void prof_record(cudaEvent_t ev, cudaStream_t stream) {
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  cudaStreamIsCapturing(stream, &cs);
  if (cs == cudaStreamCaptureStatusActive) cudaEventRecordWithFlags(ev, stream, cudaEventRecordExternal);
  else cudaEventRecord(ev, stream);
}

struct ProfScope {
  ProfKind* k = nullptr;
  int slot = -1;
  cudaStream_t stream;
  ProfScope(int kind, long long a, long long b, cudaStream_t st) : stream(st) {
    ProfKind& pk = g_prof[kind];
    if (!pk.armed || pk.used >= pk.cap) return;
    if ((pk.match_a && pk.match_a != a) || (pk.match_b && pk.match_b != b)) return;
    k = &pk;
    slot = pk.used++;
    prof_record(pk.e0[slot], stream);
  }
  ~ProfScope() {
    if (k) prof_record(k->e1[slot], stream);
  }
};

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

#define CUDA_TRY(expr)                                                                        \
  do {                                                                                        \
    cudaError_t e_ = (expr);                                                                  \
    if (e_ != cudaSuccess) return fail(I2V_ERR_CUDA, "%s: %s", #expr, cudaGetErrorString(e_)); \
  } while (0)

struct DeviceInfo {
  int ok = -1;  // -1 unknown
  int sms = 0;
  int major = 0, minor = 0;
};
DeviceInfo g_dev[64];

int device_info(DeviceInfo** out) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return fail(I2V_ERR_NO_DEVICE, "cudaGetDevice: %s", cudaGetErrorString(e));
  if (dev < 0 || dev >= 64) return fail(I2V_ERR_NO_DEVICE, "device ordinal %d out of range", dev);
  DeviceInfo& d = g_dev[dev];
  if (d.ok < 0) {
    cudaDeviceProp p;
    e = cudaGetDeviceProperties(&p, dev);
    if (e != cudaSuccess) return fail(I2V_ERR_NO_DEVICE, "cudaGetDeviceProperties: %s", cudaGetErrorString(e));
    d.sms = p.multiProcessorCount;
    d.major = p.major;
    d.minor = p.minor;
    d.ok = (p.major == 10) ? 1 : 0;
  }
  if (!d.ok)
    return fail(I2V_ERR_NO_DEVICE, "device %d is sm_%d%d; this library contains sm_100a code only", dev, d.major,
                d.minor);
  *out = &d;
  return 0;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g_encode = nullptr;

int get_encode_fn() {
  if (g_encode) return 0;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
  if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || fn == nullptr)
    return fail(I2V_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available (%s)", cudaGetErrorString(e));
  g_encode = reinterpret_cast<EncodeTiledFn>(fn);
  return 0;
}

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

int check_tensor(const char* name, const i2v_tensor* t, int elem_bytes, bool need_vec) {
  if (t == nullptr || t->data == nullptr) return fail(I2V_ERR_BAD_SHAPE, "%s: null tensor", name);
  if (need_vec) {
    const int vec = 16 / elem_bytes;
    if (!aligned16(t->data)) return fail(I2V_ERR_MISALIGNED, "%s: data pointer is not 16-byte aligned", name);
    if (t->stride_b % vec || t->stride_s % vec || t->stride_h % vec)
      return fail(I2V_ERR_MISALIGNED, "%s: strides must be multiples of %d elements", name, vec);
  }
  return 0;
}

// 4-D bf16 tensor map over a logical [batch, seq, heads, d] tensor: dims (d, heads, seq, batch),
// box (64, 1, box_rows, 1), 128-byte swizzle, out-of-bounds elements read as zero.
int make_tmap(CUtensorMap* tm, const i2v_tensor* t, int batch, int seq, int heads, int d, int box_rows) {
  cuuint64_t dims[4] = {(cuuint64_t)d, (cuuint64_t)heads, (cuuint64_t)seq, (cuuint64_t)batch};
  cuuint64_t strides[3] = {(cuuint64_t)t->stride_h * 2, (cuuint64_t)t->stride_s * 2, (cuuint64_t)t->stride_b * 2};
  // a size-1 dimension may carry any stride; keep the encoder happy with a 16-byte multiple
  if (heads == 1 && strides[0] == 0) strides[0] = (cuuint64_t)d * 2;
  if (batch == 1 && strides[2] == 0) strides[2] = (cuuint64_t)t->stride_s * 2 * seq;
  cuuint32_t box[4] = {64, 1, (cuuint32_t)box_rows, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = g_encode(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, t->data, dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail(I2V_ERR_CUDA,
                "cuTensorMapEncodeTiled failed (CUresult %d) dims=(%d,%d,%d,%d) strides(B)=(%lld,%lld,%lld)", (int)r, d,
                heads, seq, batch, (long long)strides[0], (long long)strides[1], (long long)strides[2]);
  return 0;
}

struct DenseSeg {
  const i2v_tensor *q, *k, *v, *o;
  int kv_group;
};

int dense_dk(int d) {
  const int dk = (d + 15) / 16 * 16;
  switch (dk) {
    case 16: case 32: case 48: case 64: case 80: case 96: case 128: case 160: return dk;
    default: return 0;
  }
}
int dense_block_n(int dk) { return dk > 128 ? 64 : 128; }

bool dense_supported(int d, int dtype) { return dtype == I2V_BF16 && d % 8 == 0 && dense_dk(d) != 0; }

template <class Cfg>
int launch_dense_cfg(const i2v::DenseParams& Pin, cudaStream_t stream) {
  static bool attr_set[64] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  auto kern = i2v::dense_attn_kernel<Cfg>;
  if (!attr_set[dev & 63]) {
    CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    attr_set[dev & 63] = true;
  }
  i2v::DenseParams P = Pin;
  P.q_blocks = (P.sq + 128 * Cfg::NT - 1) / (128 * Cfg::NT);
  const long long grid = (long long)P.q_blocks * P.heads * P.batch * P.nprob;
  if (grid <= 0 || grid > 0x7fffffffLL) return fail(I2V_ERR_BAD_SHAPE, "dense attention grid %lld out of range", grid);
  kern<<<(unsigned)grid, Cfg::THREADS, Cfg::SMEM_BYTES, stream>>>(P);
  CUDA_TRY(cudaGetLastError());
  g_launches.fetch_add(1);
  return 0;
}

template <class Cfg>
int launch_dense_pipe_cfg(const i2v::DenseParams& Pin, cudaStream_t stream) {
  static bool attr_set[64] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  auto kern = i2v::dense_attn_pipe_kernel<Cfg>;
  if (!attr_set[dev & 63]) {
    CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    attr_set[dev & 63] = true;
  }
  i2v::DenseParams P = Pin;
  P.q_blocks = (P.sq + 128 * Cfg::NT - 1) / (128 * Cfg::NT);
  const long long items = (long long)P.q_blocks * P.heads * P.batch * P.nprob;
  if (items <= 0 || items > 0x7fffffffLL) return fail(I2V_ERR_BAD_SHAPE, "dense attention work items %lld out of range", items);
  DeviceInfo* di = nullptr;
  int rc = device_info(&di);
  if (rc) return rc;
  const long long slots = (long long)di->sms * Cfg::MINB;
  const long long grid = items < slots ? items : slots;   // persistent: one CTA per SM (or MINB) walks the items
  kern<<<(unsigned)grid, Cfg::THREADS, Cfg::SMEM_BYTES, stream>>>(P);
  CUDA_TRY(cudaGetLastError());
  g_launches.fetch_add(1);
  return 0;
}

template <int D, int HG, int MT, int NSTG, int MINB, int NTXT = 0, int NK = 0>
int launch_ip_stream(const i2v::IpStreamParams& P, int sms, cudaStream_t stream) {
  using Cfg = i2v::IpStreamCfg<D, HG, MT, NSTG, MINB>;
  static bool attr_set[64] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  auto kern = i2v::ip_xattn_stream_kernel<D, HG, MT, NSTG, MINB, NTXT, NK>;
  if (!attr_set[dev & 63]) {
    CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    attr_set[dev & 63] = true;
  }
  const long long units = (long long)P.batch * (P.heads / HG) * ((P.sq + Cfg::ROWS - 1) / Cfg::ROWS);
  const long long grid = units < (long long)sms * MINB ? units : (long long)sms * MINB;
  ProfScope prof(3, P.sq, P.batch, stream);
  kern<<<(unsigned)grid, i2v::kIpThreads, Cfg::SMEM_BYTES, stream>>>(P);
  CUDA_TRY(cudaGetLastError());
  g_launches.fetch_add(1);
  return 0;
}

int launch_dense(const DenseSeg* segs, int nseg, int batch, int heads, int sq, int skv, int d, float scale,
                 int seg_split, float seg_scale, cudaStream_t stream) {
  int rc = get_encode_fn();
  if (rc) return rc;
  const int dk = dense_dk(d);
  // tile variant for d <= 48 (tuning key 3, value - 1): 0 = two 128x128 tiles, one CTA per SM; 1 (default) = two 128x64
  // tiles per CTA, two CTAs per SM; 2 = four 128x64 tiles per CTA, each warpgroup alternating between two
  const int variant = g_tuning[3] > 0 ? g_tuning[3] - 1 : 4;
  // two-segment (IP-Adapter) launches keep all keys in one 128-key tile; at d = 160 that takes the single-query-tile config
  // variant 4 (default): software-pipelined persistent kernel (dense_attn_pipe_sm100.cuh), 3 query tiles x 64 keys
  const bool pipe = dk == 48 && seg_split < 0 && variant >= 3;
  // d = 80 (SD1.5 level 1) runs on the pipelined kernel too (two swizzle sub-tiles per row): two query tiles x 64 keys by
  // default (240 us against 301 us at C2 level 1), tuning key 3 = 6 -> three query tiles x 48 keys (249 us: the ragged third
  // tile at S = 1024 costs more than the extra tile in flight gains), 3 = 1 -> the first tcgen05 kernel (128 x 128 tiles)
  const int pipe80 = (dk == 80 && seg_split < 0) ? (g_tuning[3] == 6 ? 48 : g_tuning[3] == 1 ? 0 : 64) : 0;
  const int bn = seg_split >= 0 ? 128 : pipe80 ? pipe80 : ((dk == 48 && variant >= 1) ? 64 : dense_block_n(dk));
  if (seg_split >= 0 && skv > bn)
    return fail(I2V_ERR_UNSUPPORTED, "two-segment softmax needs skv (%d) <= %d", skv, bn);
  i2v::DenseParams P;
  memset(&P, 0, sizeof(P));
  P.nprob = nseg;
  P.batch = batch; P.heads = heads; P.sq = sq; P.skv = skv; P.d = d;
  P.scale_log2e = scale * 1.4426950408889634f;
  P.seg_split = seg_split;
  P.seg_scale = seg_scale;
#ifdef I2V_TRACE
  P.trace = g_trace;
  P.trace_cta = g_trace_cta;
#endif
  for (int i = 0; i < nseg; ++i) {
    const DenseSeg& s = segs[i];
    if (batch % s.kv_group) return fail(I2V_ERR_BAD_SHAPE, "batch %d not divisible by kv_group %d", batch, s.kv_group);
    if ((rc = check_tensor("q", s.q, 2, true)) || (rc = check_tensor("k", s.k, 2, true)) ||
        (rc = check_tensor("v", s.v, 2, true)) || (rc = check_tensor("o", s.o, 2, true)))
      return rc;
    if ((rc = make_tmap(&P.prob[i].tm_q, s.q, batch, sq, heads, d, 128))) return rc;
    if ((rc = make_tmap(&P.prob[i].tm_k, s.k, batch / s.kv_group, skv, heads, d, bn))) return rc;
    if ((rc = make_tmap(&P.prob[i].tm_v, s.v, batch / s.kv_group, skv, heads, d, bn))) return rc;
    P.prob[i].o = reinterpret_cast<__nv_bfloat16*>(s.o->data);
    P.prob[i].o_sb = s.o->stride_b; P.prob[i].o_ss = s.o->stride_s; P.prob[i].o_sh = s.o->stride_h;
    P.prob[i].kv_group = s.kv_group;
  }
  // g_tuning[2]: exp2-emulation split override for experiments (pairs out of 8 on the FMA pipe); 0 = default
  const int emu = g_tuning[2] > 0 ? g_tuning[2] - 1 : -1;
  ProfScope prof(seg_split >= 0 ? 3 : 1, sq, batch, stream);
  switch (dk) {
    case 16:  return launch_dense_cfg<i2v::DenseCfg<16, 128, 4>>(P, stream);
    case 32:  return launch_dense_cfg<i2v::DenseCfg<32, 128, 4>>(P, stream);
    case 48:
      if (pipe) {
        switch (emu) {
          case 0:  return launch_dense_pipe_cfg<i2v::PipeCfg<48, 64, 3, 5, 0>>(P, stream);
          case 2:  return launch_dense_pipe_cfg<i2v::PipeCfg<48, 64, 3, 5, 2>>(P, stream);
          case 4:  return launch_dense_pipe_cfg<i2v::PipeCfg<48, 64, 3, 5, 4>>(P, stream);
          default: return launch_dense_pipe_cfg<i2v::PipeCfg<48, 64, 3, 5, 3>>(P, stream);
        }
      }
      if (bn == 64 && variant == 2) {  // four query tiles per CTA, each warpgroup alternates between two
        switch (emu) {
          case 0:  return launch_dense_cfg<i2v::DenseCfg<48, 64, 4, 0, 1, 2>>(P, stream);
          case 2:  return launch_dense_cfg<i2v::DenseCfg<48, 64, 4, 2, 1, 2>>(P, stream);
          case 4:  return launch_dense_cfg<i2v::DenseCfg<48, 64, 4, 4, 1, 2>>(P, stream);
          default: return launch_dense_cfg<i2v::DenseCfg<48, 64, 4, 3, 1, 2>>(P, stream);
        }
      }
      if (bn == 64) {
        switch (emu) {
          case 0:  return launch_dense_cfg<i2v::DenseCfg<48, 64, 4, 0, 2, 1, 1>>(P, stream);
          case 3:  return launch_dense_cfg<i2v::DenseCfg<48, 64, 4, 3, 2, 1, 1>>(P, stream);
          case 4:  return launch_dense_cfg<i2v::DenseCfg<48, 64, 4, 4, 2, 1, 1>>(P, stream);
          default: return launch_dense_cfg<i2v::DenseCfg<48, 64, 4, 2, 2, 1, 1>>(P, stream);
        }
      }
      switch (emu) {
        case 0:  return launch_dense_cfg<i2v::DenseCfg<48, 128, 4, 0>>(P, stream);
        case 2:  return launch_dense_cfg<i2v::DenseCfg<48, 128, 4, 2>>(P, stream);
        case 4:  return launch_dense_cfg<i2v::DenseCfg<48, 128, 4, 4>>(P, stream);
        default: return launch_dense_cfg<i2v::DenseCfg<48, 128, 4, 3>>(P, stream);
      }
    case 64:  return launch_dense_cfg<i2v::DenseCfg<64, 128, 4, 2>>(P, stream);
    case 80:
      if (pipe80 == 48) {
        switch (emu) {
          case 0:  return launch_dense_pipe_cfg<i2v::PipeCfg<80, 48, 3, 5, 0>>(P, stream);
          case 3:  return launch_dense_pipe_cfg<i2v::PipeCfg<80, 48, 3, 5, 3>>(P, stream);
          default: return launch_dense_pipe_cfg<i2v::PipeCfg<80, 48, 3, 5, 2>>(P, stream);
        }
      }
      if (pipe80 == 64) {
        switch (emu) {
          case 0:  return launch_dense_pipe_cfg<i2v::PipeCfg<80, 64, 2, 4, 0>>(P, stream);
          case 2:  return launch_dense_pipe_cfg<i2v::PipeCfg<80, 64, 2, 4, 2>>(P, stream);
          case 4:  return launch_dense_pipe_cfg<i2v::PipeCfg<80, 64, 2, 4, 4>>(P, stream);
          case 5:  return launch_dense_pipe_cfg<i2v::PipeCfg<80, 64, 2, 5, 3>>(P, stream);   // (five-stage ring)
          default: return launch_dense_pipe_cfg<i2v::PipeCfg<80, 64, 2, 4, 3>>(P, stream);
        }
      }
      switch (emu) {
        case 0:  return launch_dense_cfg<i2v::DenseCfg<80, 128, 2, 0>>(P, stream);
        default: return launch_dense_cfg<i2v::DenseCfg<80, 128, 2, 2>>(P, stream);
      }
    case 96:  return launch_dense_cfg<i2v::DenseCfg<96, 128, 2>>(P, stream);
    case 128: return launch_dense_cfg<i2v::DenseCfg<128, 128, 2>>(P, stream);
    case 160:
      if (bn == 128) return launch_dense_cfg<i2v::DenseCfg<160, 128, 1, 0, 1, 1, 0, 1>>(P, stream);
      return launch_dense_cfg<i2v::DenseCfg<160, 64, 2>>(P, stream);
  }
  return fail(I2V_ERR_UNSUPPORTED, "head dim %d not covered by the tcgen05 kernels", d);
}

int launch_generic(const i2v_tensor* q, const i2v_tensor* k, const i2v_tensor* v, const i2v_tensor* o,
                   const i2v_tensor* k2, const i2v_tensor* v2, int batch, int heads, int sq, int skv, int skv2, int d,
                   int kv_group, float scale, float scale2, int dtype, cudaStream_t stream) {
  if (d > 160) return fail(I2V_ERR_UNSUPPORTED, "generic kernel supports head dim <= 160 (got %d)", d);
  const int eb = dtype == I2V_F32 ? 4 : 2;
  int rc;
  if ((rc = check_tensor("q", q, eb, false)) || (rc = check_tensor("k", k, eb, false)) ||
      (rc = check_tensor("v", v, eb, false)) || (rc = check_tensor("o", o, eb, false)))
    return rc;
  i2v::GenericParams P;
  memset(&P, 0, sizeof(P));
  P.q = q->data; P.k = k->data; P.v = v->data; P.o = o->data;
  P.q_sb = q->stride_b; P.q_ss = q->stride_s; P.q_sh = q->stride_h;
  P.k_sb = k->stride_b; P.k_ss = k->stride_s; P.k_sh = k->stride_h;
  P.v_sb = v->stride_b; P.v_ss = v->stride_s; P.v_sh = v->stride_h;
  P.o_sb = o->stride_b; P.o_ss = o->stride_s; P.o_sh = o->stride_h;
  P.batch = batch; P.heads = heads; P.sq = sq; P.skv = skv; P.d = d; P.kv_group = kv_group; P.scale = scale;
  if (skv2 > 0) {
    if ((rc = check_tensor("k_ip", k2, eb, false)) || (rc = check_tensor("v_ip", v2, eb, false))) return rc;
    P.k2 = k2->data; P.v2 = v2->data;
    P.k2_sb = k2->stride_b; P.k2_ss = k2->stride_s; P.k2_sh = k2->stride_h;
    P.v2_sb = v2->stride_b; P.v2_ss = v2->stride_s; P.v2_sh = v2->stride_h;
    P.skv2 = skv2; P.scale2 = scale2;
  }
  const long long grid = (long long)((sq + i2v::kGenRows - 1) / i2v::kGenRows) * heads * batch;
  if (grid <= 0 || grid > 0x7fffffffLL) return fail(I2V_ERR_BAD_SHAPE, "generic attention grid %lld out of range", grid);
  const size_t smem = 2 * i2v::kGenKeys * (d + 1) * sizeof(float);
  const int threads = i2v::kGenRows * i2v::kGenSlices;
  const bool small = d <= 64;
  if (dtype == I2V_F32) {
    if (small) i2v::generic_attn_kernel<float, 16><<<(unsigned)grid, threads, smem, stream>>>(P);
    else       i2v::generic_attn_kernel<float, 40><<<(unsigned)grid, threads, smem, stream>>>(P);
  } else {
    if (small) i2v::generic_attn_kernel<__nv_bfloat16, 16><<<(unsigned)grid, threads, smem, stream>>>(P);
    else       i2v::generic_attn_kernel<__nv_bfloat16, 40><<<(unsigned)grid, threads, smem, stream>>>(P);
  }
  CUDA_TRY(cudaGetLastError());
  g_launches.fetch_add(1);
  return 0;
}

int check_common(int batch, int heads, int sq, int skv, int d, int dtype, int mode) {
  if (batch <= 0 || heads <= 0 || sq <= 0 || skv <= 0 || d <= 0)
    return fail(I2V_ERR_BAD_SHAPE, "sizes must be positive (batch=%d heads=%d sq=%d skv=%d d=%d)", batch, heads, sq, skv,
                d);
  if (dtype != I2V_BF16 && dtype != I2V_F32) return fail(I2V_ERR_BAD_DTYPE, "dtype %d is not I2V_BF16 / I2V_F32", dtype);
  if (mode < I2V_MODE_AUTO || mode > I2V_MODE_GENERIC) return fail(I2V_ERR_BAD_SHAPE, "unknown mode %d", mode);
  return 0;
}

// ---------------------------------------------------------------------------------------------------------
// temporal
// ---------------------------------------------------------------------------------------------------------
template <int D, int HG, int FT>
int launch_temporal_cfg(const i2v::TemporalParams& P, int sms, cudaStream_t stream) {
  using Cfg = i2v::TemporalCfg<D, HG, FT>;
  // As many co-resident CTAs per SM as 2-stage rings fit (up to 3: measured best, 5.8 TB/s at d = 40 against 4.5 with
  // two and 2.5 with one); a lone CTA gets a deeper ring instead.
  const int budget = 226 * 1024;
  const int ring_overhead = 256 + 128;
  auto fits = [&](int ctas, int stages) { return ctas * (stages * Cfg::STAGE_BYTES + ring_overhead + 1024) <= budget; };
  int per_sm = g_tuning[1] > 0 ? g_tuning[1] : (fits(3, 2) ? 3 : (fits(2, 2) ? 2 : 1));
  int stages = g_tuning[0] > 0 ? g_tuning[0] : (per_sm >= 2 ? 2 : 4);
  while (stages > 2 && !fits(per_sm, stages)) --stages;
  while (per_sm > 1 && !fits(per_sm, stages)) --per_sm;
  if (stages > 8) stages = 8;
  const size_t smem = (size_t)stages * Cfg::STAGE_BYTES + ring_overhead;
  if (smem > 227 * 1024) return fail(I2V_ERR_UNSUPPORTED, "temporal stage (%d B) does not fit shared memory", Cfg::STAGE_BYTES);
  auto kern = i2v::temporal_attn_kernel<D, HG, FT>;
  CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const long long units = (long long)P.n_pos * (P.heads / HG);
  long long grid = (long long)sms * per_sm;
  if (grid > units) grid = units;
  ProfScope prof(2, P.n_pos, D, stream);
  kern<<<(unsigned)grid, i2v::kTemporalThreads, smem, stream>>>(P, stages);
  CUDA_TRY(cudaGetLastError());
  g_launches.fetch_add(1);
  return 0;
}

// heads per staged unit: keep a stage (3 slabs of FT*16 rows) around 32-64 KB so >= 2 stages always fit
// (measured: 3 resident CTAs with ~31 KB stages beat fewer CTAs with larger stages: d=80 4.9 TB/s at HG=4 vs 4.2 at HG=8)
template <int D, int FT> struct TemporalHG { static constexpr int value = (FT == 1) ? (D <= 40 ? 8 : 4) : (D <= 40 ? 8 : (D <= 80 ? 4 : 2)); };

template <int D>
int launch_temporal_ft(const i2v::TemporalParams& P, int sms, cudaStream_t stream) {
  // tuning key 4: heads per staged unit override for experiments (must divide 8 and keep D/8 output tiles splittable)
  if (P.frames <= 16) {
    if constexpr (D == 80 || D == 64) {
      if (g_tuning[4] == 8) return launch_temporal_cfg<D, 8, 1>(P, sms, stream);
    }
    if constexpr (D == 160 || D == 128) {
      if (g_tuning[4] == 2) return launch_temporal_cfg<D, 2, 1>(P, sms, stream);
    }
    return launch_temporal_cfg<D, TemporalHG<D, 1>::value, 1>(P, sms, stream);
  }
  return launch_temporal_cfg<D, TemporalHG<D, 2>::value, 2>(P, sms, stream);
}

bool temporal_supported(int heads, int frames, int d, int dtype, const i2v_tensor* q, const i2v_tensor* k,
                        const i2v_tensor* v, const i2v_tensor* o) {
  if (dtype != I2V_BF16 || frames > 32) return false;
  if (!(d == 16 || d == 32 || d == 40 || d == 64 || d == 80 || d == 128 || d == 160)) return false;
  if (heads % 8) return false;
  if (q->stride_h != d || k->stride_h != d || v->stride_h != d || o->stride_h != d) return false;
  return true;
}

}  // namespace

// =========================================================================================================
extern "C" {

int i2v_version(void) { return 100; }
const char* i2v_last_error(void) { return g_err; }
int64_t i2v_launch_count(void) { return g_launches.load(); }

int i2v_device_supported(void) {
  DeviceInfo* di = nullptr;
  return device_info(&di) == 0 ? 1 : 0;
}

#ifdef I2V_TRACE
// developer builds only (not part of the C ABI): device buffer of 16 * 1024 uint64 for the dense-kernel timeline
int i2v_debug_set_trace(void* device_buffer, int cta) {
  g_trace = (unsigned long long*)device_buffer;
  g_trace_cta = cta;
  return 0;
}
#endif

int i2v_set_tuning(int key, int value) {
  if (key < 0 || key >= 12) return fail(I2V_ERR_BAD_SHAPE, "unknown tuning key %d", key);
  g_tuning[key] = value;
  return 0;
}

int i2v_prof_arm(int kind, long long match_a, long long match_b, int max_pairs) {
  if (kind <= 0 || kind >= kProfKinds) return fail(I2V_ERR_BAD_SHAPE, "i2v_prof_arm: unknown kernel class %d", kind);
  ProfKind& pk = g_prof[kind];
  if (max_pairs <= 0) {   // disarm: launches stop taking new pairs; pairs already recorded (or captured) stay readable
    pk.armed = false;
    return 0;
  }
  if (max_pairs > kProfMaxPairs) max_pairs = kProfMaxPairs;
  while (pk.created < max_pairs) {
    CUDA_TRY(cudaEventCreate(&pk.e0[pk.created]));
    CUDA_TRY(cudaEventCreate(&pk.e1[pk.created]));
    ++pk.created;
  }
  pk.armed = true;
  pk.match_a = match_a;
  pk.match_b = match_b;
  pk.cap = max_pairs;
  pk.used = 0;
  return 0;
}

int i2v_prof_read(int kind, float* ms_out, int capacity) {
  if (kind <= 0 || kind >= kProfKinds) return fail(I2V_ERR_BAD_SHAPE, "i2v_prof_read: unknown kernel class %d", kind);
  ProfKind& pk = g_prof[kind];
  int n = 0;
  for (; n < pk.used && n < capacity; ++n) {
    cudaError_t e = cudaEventElapsedTime(&ms_out[n], pk.e0[n], pk.e1[n]);
    if (e != cudaSuccess) {
      cudaGetLastError();
      return fail(I2V_ERR_CUDA, "i2v_prof_read: pair %d not complete (%s): synchronise the stream first", n,
                  cudaGetErrorString(e));
    }
  }
  return n;
}

int i2v_sdpa_fwd(const i2v_tensor* q, const i2v_tensor* k, const i2v_tensor* v, const i2v_tensor* o, int batch,
                 int heads, int sq, int skv, int d, int kv_group, float scale, int dtype, int mode, void* stream) {
  int rc = check_common(batch, heads, sq, skv, d, dtype, mode);
  if (rc) return rc;
  if (kv_group <= 0 || batch % kv_group) return fail(I2V_ERR_BAD_SHAPE, "batch %d not divisible by kv_group %d", batch, kv_group);
  DeviceInfo* di = nullptr;
  if ((rc = device_info(&di))) return rc;
  const bool fast_ok = dense_supported(d, dtype);
  if (mode == I2V_MODE_FAST && !fast_ok)
    return fail(I2V_ERR_UNSUPPORTED, "i2v_sdpa_fwd: FAST path needs bf16 and a supported head dim (d=%d dtype=%d)", d, dtype);
  if (mode != I2V_MODE_GENERIC && fast_ok) {
    DenseSeg seg{q, k, v, o, kv_group};
    return launch_dense(&seg, 1, batch, heads, sq, skv, d, scale, -1, 0.f, (cudaStream_t)stream);
  }
  return launch_generic(q, k, v, o, nullptr, nullptr, batch, heads, sq, skv, 0, d, kv_group, scale, 0.f, dtype,
                        (cudaStream_t)stream);
}

int i2v_fused_self_xframe_fwd(const i2v_tensor* q_self, const i2v_tensor* k_self, const i2v_tensor* v_self,
                              const i2v_tensor* o_self, const i2v_tensor* q_x, const i2v_tensor* k_x,
                              const i2v_tensor* v_x, const i2v_tensor* o_x, int batch, int heads, int seq, int d,
                              int num_frames, float scale, int dtype, int mode, void* stream) {
  int rc = check_common(batch, heads, seq, seq, d, dtype, mode);
  if (rc) return rc;
  // same checks (and messages) as the reference block: src/modules/i2v_adapter.py:477-481
  if (num_frames <= 0) return fail(I2V_ERR_BAD_SHAPE, "`num_frames` must be provided when `enable_cross_frame_attn` is True.");
  if (batch % num_frames)
    return fail(I2V_ERR_BAD_SHAPE, "Batch size %d must be divisible by the number of frames %d.", batch, num_frames);
  DeviceInfo* di = nullptr;
  if ((rc = device_info(&di))) return rc;
  const bool fast_ok = dense_supported(d, dtype);
  if (mode == I2V_MODE_FAST && !fast_ok)
    return fail(I2V_ERR_UNSUPPORTED, "i2v_fused_self_xframe_fwd: FAST path needs bf16 and a supported head dim (d=%d)", d);
  if (mode != I2V_MODE_GENERIC && fast_ok) {
    DenseSeg segs[2] = {{q_self, k_self, v_self, o_self, 1}, {q_x, k_x, v_x, o_x, num_frames}};
    return launch_dense(segs, 2, batch, heads, seq, seq, d, scale, -1, 0.f, (cudaStream_t)stream);
  }
  rc = launch_generic(q_self, k_self, v_self, o_self, nullptr, nullptr, batch, heads, seq, seq, 0, d, 1, scale, 0.f,
                      dtype, (cudaStream_t)stream);
  if (rc) return rc;
  return launch_generic(q_x, k_x, v_x, o_x, nullptr, nullptr, batch, heads, seq, seq, 0, d, num_frames, scale, 0.f,
                        dtype, (cudaStream_t)stream);
}

int i2v_fused_self_xframe_aug_fwd(const i2v_tensor* q_self, const i2v_tensor* k_self, const i2v_tensor* v_self,
                                  const i2v_tensor* o_self, const i2v_tensor* q_x, const i2v_tensor* k_x,
                                  const i2v_tensor* v_x, const i2v_tensor* o_x, int batch, int heads, int seq, int d,
                                  int d_pad, int num_frames, int dtype, void* stream) {
  int rc = check_common(batch, heads, seq, seq, d, dtype, I2V_MODE_FAST);
  if (rc) return rc;
  if (num_frames <= 0) return fail(I2V_ERR_BAD_SHAPE, "`num_frames` must be provided when `enable_cross_frame_attn` is True.");
  if (batch % num_frames)
    return fail(I2V_ERR_BAD_SHAPE, "Batch size %d must be divisible by the number of frames %d.", batch, num_frames);
  if (dtype != I2V_BF16 || d != i2v::kAugCol || d_pad != 48)
    return fail(I2V_ERR_UNSUPPORTED, "i2v_fused_self_xframe_aug_fwd: bf16, d = %d stored padded to 48 only (got d=%d d_pad=%d)",
                i2v::kAugCol, d, d_pad);
  DeviceInfo* di = nullptr;
  if ((rc = device_info(&di))) return rc;
  if ((rc = get_encode_fn())) return rc;
#ifdef I2V_EXPERIMENTS
  const int bn = (g_tuning[7] == 14 || g_tuning[7] == 15) ? 48 : 64;
#else
  const int bn = 64;
#endif
  i2v::DenseParams P;
  memset(&P, 0, sizeof(P));
  P.nprob = 2;
  P.batch = batch; P.heads = heads; P.sq = seq; P.skv = seq; P.d = d;
  P.scale_log2e = 1.f;
  P.seg_split = -1;
#ifdef I2V_TRACE
  P.trace = g_trace;
  P.trace_cta = g_trace_cta;
#endif
  const DenseSeg segs[2] = {{q_self, k_self, v_self, o_self, 1}, {q_x, k_x, v_x, o_x, num_frames}};
  for (int i = 0; i < 2; ++i) {
    const DenseSeg& s = segs[i];
    if ((rc = check_tensor("q", s.q, 2, true)) || (rc = check_tensor("k", s.k, 2, true)) ||
        (rc = check_tensor("v", s.v, 2, true)) || (rc = check_tensor("o", s.o, 2, true)))
      return rc;
    if ((rc = make_tmap(&P.prob[i].tm_q, s.q, batch, seq, heads, d_pad, 128))) return rc;
    if ((rc = make_tmap(&P.prob[i].tm_k, s.k, batch / s.kv_group, seq, heads, d_pad, bn))) return rc;
    if ((rc = make_tmap(&P.prob[i].tm_v, s.v, batch / s.kv_group, seq, heads, d_pad, bn))) return rc;
    P.prob[i].o = reinterpret_cast<__nv_bfloat16*>(s.o->data);
    P.prob[i].o_sb = s.o->stride_b; P.prob[i].o_ss = s.o->stride_s; P.prob[i].o_sh = s.o->stride_h;
    P.prob[i].kv_group = s.kv_group;
  }
  const int emu = g_tuning[2] > 0 ? g_tuning[2] - 1 : -1;
  ProfScope prof(1, seq, batch, (cudaStream_t)stream);
  // tuning key 7 (this entry), experiments kept for the record: 1 / 2 / 3 / 4 = column-split softmax (two
  // threads per row) with 3 / 2 / 4 / 0 of 8 pairs emulated -- measured 14 % slower than the default
  if (g_tuning[7] == 1) return launch_dense_pipe_cfg<i2v::PipeCfg<48, 64, 3, 5, 3, 3, true, true>>(P, (cudaStream_t)stream);
  if (g_tuning[7] == 2) return launch_dense_pipe_cfg<i2v::PipeCfg<48, 64, 3, 5, 2, 3, true, true>>(P, (cudaStream_t)stream);
  if (g_tuning[7] == 3) return launch_dense_pipe_cfg<i2v::PipeCfg<48, 64, 3, 5, 4, 3, true, true>>(P, (cudaStream_t)stream);
  if (g_tuning[7] == 4) return launch_dense_pipe_cfg<i2v::PipeCfg<48, 64, 3, 5, 0, 3, true, true>>(P, (cudaStream_t)stream);
#ifdef I2V_EXPERIMENTS
  if (g_tuning[7] == 7) return launch_dense_pipe_cfg<i2v::PipeCfg<48, 64, 3, 5, 3, 3, true, false, 1>>(P, (cudaStream_t)stream);
  if (g_tuning[7] == 8) return launch_dense_pipe_cfg<i2v::PipeCfg<48, 64, 3, 5, 2, 3, true, false, 1>>(P, (cudaStream_t)stream);
  if (g_tuning[7] == 9) return launch_dense_pipe_cfg<i2v::PipeCfg<48, 64, 3, 5, 4, 3, true, false, 1>>(P, (cudaStream_t)stream);
  if (g_tuning[7] == 10) return launch_dense_pipe_cfg<i2v::PipeCfg<48, 64, 3, 5, 3, 3, true, false, 2>>(P, (cudaStream_t)stream);   // QK issued twice
  if (g_tuning[7] == 11) return launch_dense_pipe_cfg<i2v::PipeCfg<48, 64, 3, 5, 0, 3, true, false, 2>>(P, (cudaStream_t)stream);   // ... with EMU 0
  if (g_tuning[7] == 12) return launch_dense_pipe_cfg<i2v::PipeCfg<48, 64, 3, 8, 3, 3, true>>(P, (cudaStream_t)stream);   // 8-stage K/V ring
  if (g_tuning[7] == 13) return launch_dense_pipe_cfg<i2v::PipeCfg<48, 64, 3, 3, 3, 3, true>>(P, (cudaStream_t)stream);   // 3-stage K/V ring
  // one query tile (48-key KV tiles, 120 TMEM columns) per CTA, four / three co-resident CTAs per SM: independent pipelines
  if (g_tuning[7] == 14) return launch_dense_pipe_cfg<i2v::PipeCfg<48, 48, 1, 3, 3, 3, true, false, 0, 4>>(P, (cudaStream_t)stream);
  if (g_tuning[7] == 15) return launch_dense_pipe_cfg<i2v::PipeCfg<48, 48, 1, 4, 3, 3, true, false, 0, 3>>(P, (cudaStream_t)stream);
  // hand-off pipeline floor: no exponentials at all (results are garbage), one / two threads per row
  if (g_tuning[7] == 5) return launch_dense_pipe_cfg<i2v::PipeCfg<48, 64, 3, 5, 0, 0, true>>(P, (cudaStream_t)stream);
  if (g_tuning[7] == 6) return launch_dense_pipe_cfg<i2v::PipeCfg<48, 64, 3, 5, 0, 0, true, true>>(P, (cudaStream_t)stream);
#endif
  switch (emu) {
    case 0:  return launch_dense_pipe_cfg<i2v::PipeCfg<48, 64, 3, 5, 0, 3, true>>(P, (cudaStream_t)stream);
    case 2:  return launch_dense_pipe_cfg<i2v::PipeCfg<48, 64, 3, 5, 2, 3, true>>(P, (cudaStream_t)stream);
    case 4:  return launch_dense_pipe_cfg<i2v::PipeCfg<48, 64, 3, 5, 4, 3, true>>(P, (cudaStream_t)stream);
    case 5:  return launch_dense_pipe_cfg<i2v::PipeCfg<48, 64, 3, 5, 4, 2, true>>(P, (cudaStream_t)stream);
    default: return launch_dense_pipe_cfg<i2v::PipeCfg<48, 64, 3, 5, 3, 3, true>>(P, (cudaStream_t)stream);
  }
}

int i2v_ip_xattn_fwd(const i2v_tensor* q, const i2v_tensor* k_txt, const i2v_tensor* v_txt, const i2v_tensor* k_ip,
                     const i2v_tensor* v_ip, const i2v_tensor* o, int batch, int heads, int sq, int n_txt, int n_ip,
                     int d, int kv_group, float scale, float ip_scale, int dtype, int mode, void* stream) {
  int rc = check_common(batch, heads, sq, n_txt, d, dtype, mode);
  if (rc) return rc;
  if (n_ip <= 0) return fail(I2V_ERR_BAD_SHAPE, "n_ip must be positive (got %d)", n_ip);
  if (kv_group <= 0 || batch % kv_group) return fail(I2V_ERR_BAD_SHAPE, "batch %d not divisible by kv_group %d", batch, kv_group);
  DeviceInfo* di = nullptr;
  if ((rc = device_info(&di))) return rc;
  if (!k_txt || !v_txt || !k_ip || !v_ip) return fail(I2V_ERR_BAD_SHAPE, "null key/value tensor");
  const int eb = dtype == I2V_F32 ? 4 : 2;
  const bool contiguous_tokens =
      (const char*)k_ip->data == (const char*)k_txt->data + (long long)n_txt * k_txt->stride_s * eb &&
      (const char*)v_ip->data == (const char*)v_txt->data + (long long)n_txt * v_txt->stride_s * eb &&
      k_ip->stride_b == k_txt->stride_b && k_ip->stride_s == k_txt->stride_s && k_ip->stride_h == k_txt->stride_h &&
      v_ip->stride_b == v_txt->stride_b && v_ip->stride_s == v_txt->stride_s && v_ip->stride_h == v_txt->stride_h;
  const bool fast_ok =
      dense_supported(d, dtype) && contiguous_tokens && (n_txt + n_ip) <= 128;
  if (mode == I2V_MODE_FAST && !fast_ok)
    return fail(I2V_ERR_UNSUPPORTED,
                "i2v_ip_xattn_fwd: FAST path needs bf16, supported head dim, image tokens stored right after the text "
                "tokens and n_txt+n_ip <= tile (d=%d n_txt=%d n_ip=%d contiguous=%d)", d, n_txt, n_ip, (int)contiguous_tokens);
  if (mode != I2V_MODE_GENERIC && fast_ok) {
    // streaming kernel (HBM-bound design) for the SD1.5 shapes: head dims 40 / 80 / 160 with head groups of 320
    // columns, rows contiguous across heads, <= 96 keys; tuning key 5 = 1 forces the tcgen05 single-tile kernel
    const int cfgsel = g_tuning[6];   // experiments: 0 = default configuration per head dim
    const int hg = d == 40 ? (cfgsel == 1 ? 8 : cfgsel == 2 ? 4 : 2) : d == 80 ? (cfgsel == 1 ? 4 : 2) : d == 160 ? 2 : 0;
    const bool rows_ok = q->stride_h == d && k_txt->stride_h == d && v_txt->stride_h == d && o->stride_h == d &&
                         q->stride_s % 8 == 0 && k_txt->stride_s % 8 == 0 && v_txt->stride_s % 8 == 0 &&
                         o->stride_s % 8 == 0 && q->stride_b % 8 == 0 && k_txt->stride_b % 8 == 0 &&
                         v_txt->stride_b % 8 == 0 && o->stride_b % 8 == 0;
    // d = 40 with the pipeline's token counts: tcgen05 kernel with K / V of one (video, head) resident per CTA
    // (ip_xattn_tc_sm100.cuh); tuning key 5 = 4 keeps the streaming kernel
    if (d == 40 && n_txt == 77 && n_ip == 4 && (batch / kv_group) * heads <= di->sms && g_tuning[5] == 0) {
      if ((rc = check_tensor("q", q, 2, true)) || (rc = check_tensor("k", k_txt, 2, true)) ||
          (rc = check_tensor("v", v_txt, 2, true)) || (rc = check_tensor("k_ip", k_ip, 2, true)) ||
          (rc = check_tensor("v_ip", v_ip, 2, true)) || (rc = check_tensor("o", o, 2, true)))
        return rc;
      if ((rc = get_encode_fn())) return rc;
      i2v::IpTcParams T;
      memset(&T, 0, sizeof(T));
      const int bk = batch / kv_group;
      T.q = reinterpret_cast<const __nv_bfloat16*>(q->data);
      T.q_sb = q->stride_b; T.q_ss = q->stride_s; T.q_sh = q->stride_h;
      if ((rc = make_tmap(&T.tm_kt, k_txt, bk, n_txt, heads, d, i2v::kIpTcTxtRows))) return rc;
      if ((rc = make_tmap(&T.tm_vt, v_txt, bk, n_txt, heads, d, i2v::kIpTcTxtRows))) return rc;
      if ((rc = make_tmap(&T.tm_ki, k_ip, bk, n_ip, heads, d, 16))) return rc;
      if ((rc = make_tmap(&T.tm_vi, v_ip, bk, n_ip, heads, d, 16))) return rc;
      T.o = reinterpret_cast<__nv_bfloat16*>(o->data);
      T.o_sb = o->stride_b; T.o_ss = o->stride_s; T.o_sh = o->stride_h;
      T.batch = batch; T.sq = sq; T.heads = heads; T.kv_group = kv_group; T.q_tiles = (sq + 127) / 128;
      T.scale_log2e = scale * 1.4426950408889634f; T.ip_scale = ip_scale;
#ifdef I2V_TRACE
      T.trace = g_trace;
      T.trace_cta = g_trace_cta;
#endif
      static bool attr_set[64] = {false};
      int dev = 0;
      cudaGetDevice(&dev);
      auto kern = i2v::ip_xattn_tc_kernel<77, 4>;
      if (!attr_set[dev & 63]) {
        CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, i2v::kIpTcSmemBytes));
        attr_set[dev & 63] = true;
      }
      const int groups = bk * heads;
      const long long items = (long long)kv_group * T.q_tiles;
      int cpg = di->sms / groups;                      // CTAs per (video, head) group
      if (cpg > items) cpg = (int)items;
      ProfScope prof(3, sq, batch, (cudaStream_t)stream);
      kern<<<(unsigned)(groups * cpg), i2v::kIpTcThreads, i2v::kIpTcSmemBytes, (cudaStream_t)stream>>>(T);
      CUDA_TRY(cudaGetLastError());
      g_launches.fetch_add(1);
      return 0;
    }
    if (hg && heads % hg == 0 && n_txt + n_ip <= i2v::kIpKeys && rows_ok && g_tuning[5] != 1) {
      if ((rc = check_tensor("q", q, 2, true)) || (rc = check_tensor("k", k_txt, 2, true)) ||
          (rc = check_tensor("v", v_txt, 2, true)) || (rc = check_tensor("o", o, 2, true)))
        return rc;
      i2v::IpStreamParams P;
      P.q = (const __nv_bfloat16*)q->data; P.k = (const __nv_bfloat16*)k_txt->data;
      P.v = (const __nv_bfloat16*)v_txt->data; P.o = (__nv_bfloat16*)o->data;
      P.q_sb = q->stride_b; P.q_ss = q->stride_s; P.k_sb = k_txt->stride_b; P.k_ss = k_txt->stride_s;
      P.v_sb = v_txt->stride_b; P.v_ss = v_txt->stride_s; P.o_sb = o->stride_b; P.o_ss = o->stride_s;
      P.batch = batch; P.sq = sq; P.heads = heads; P.nk = n_txt + n_ip; P.n_txt = n_txt; P.kv_group = kv_group;
      P.scale_log2e = scale * 1.4426950408889634f; P.ip_scale = ip_scale;
      const bool std77 = n_txt == 77 && n_ip == 4 && g_tuning[5] != 2;   // tuning key 5 = 2: runtime token counts
      if (d == 40) {
        if (cfgsel == 1) return launch_ip_stream<40, 8, 2, 3, 1>(P, di->sms, (cudaStream_t)stream);
        if (cfgsel == 2) return launch_ip_stream<40, 4, 1, 3, 2>(P, di->sms, (cudaStream_t)stream);
        // the pipeline's token counts (77 text + 4 image) get the instantiation with compile-time key classes
        // four heads per unit, two CTAs per SM at 94 registers: no spills (the two-head, three-CTA configuration is held
        // to 72 registers and spills in the softmax): 102 us against 115 us at C2 level 0; cfgsel 3 = the latter
        if (std77 && cfgsel == 3) return launch_ip_stream<40, 2, 1, 3, 3, 77, 81>(P, di->sms, (cudaStream_t)stream);
        if (std77 && heads % 4 == 0) return launch_ip_stream<40, 4, 1, 3, 2, 77, 81>(P, di->sms, (cudaStream_t)stream);
        if (std77) return launch_ip_stream<40, 2, 1, 3, 3, 77, 81>(P, di->sms, (cudaStream_t)stream);
        return launch_ip_stream<40, 2, 1, 3, 3>(P, di->sms, (cudaStream_t)stream);   // 3 CTAs per SM
      }
      if (d == 80) {
        if (cfgsel == 1) return launch_ip_stream<80, 4, 1, 3, 1>(P, di->sms, (cudaStream_t)stream);
        if (std77) return launch_ip_stream<80, 2, 1, 2, 2, 77, 81>(P, di->sms, (cudaStream_t)stream);
        return launch_ip_stream<80, 2, 1, 2, 2>(P, di->sms, (cudaStream_t)stream);
      }
      if (std77) return launch_ip_stream<160, 2, 1, 2, 1, 77, 81>(P, di->sms, (cudaStream_t)stream);
      return launch_ip_stream<160, 2, 1, 2, 1>(P, di->sms, (cudaStream_t)stream);
    }
    DenseSeg seg{q, k_txt, v_txt, o, kv_group};
    return launch_dense(&seg, 1, batch, heads, sq, n_txt + n_ip, d, scale, n_txt, ip_scale, (cudaStream_t)stream);
  }
  return launch_generic(q, k_txt, v_txt, o, k_ip, v_ip, batch, heads, sq, n_txt, n_ip, d, kv_group, scale, ip_scale,
                        dtype, (cudaStream_t)stream);
}

int i2v_temporal_attn_fwd(const i2v_tensor* q, const i2v_tensor* k, const i2v_tensor* v, const i2v_tensor* o, int n_pos,
                          int heads, int frames, int d, float scale, int dtype, int mode, void* stream) {
  int rc = check_common(n_pos, heads, frames, frames, d, dtype, mode);
  if (rc) return rc;
  DeviceInfo* di = nullptr;
  if ((rc = device_info(&di))) return rc;
  if (!q || !k || !v || !o) return fail(I2V_ERR_BAD_SHAPE, "null tensor");
  const bool fast_ok = temporal_supported(heads, frames, d, dtype, q, k, v, o);
  if (mode == I2V_MODE_FAST && !fast_ok)
    return fail(I2V_ERR_UNSUPPORTED,
                "i2v_temporal_attn_fwd: FAST path needs bf16, frames<=32, d in {16,32,40,64,80,128,160}, heads %% 8 == 0, "
                "stride_h == d (frames=%d d=%d heads=%d)", frames, d, heads);
  if (mode != I2V_MODE_GENERIC && fast_ok) {
    if ((rc = check_tensor("q", q, 2, true)) || (rc = check_tensor("k", k, 2, true)) ||
        (rc = check_tensor("v", v, 2, true)) || (rc = check_tensor("o", o, 2, true)))
      return rc;
    i2v::TemporalParams P;
    P.q = (const __nv_bfloat16*)q->data; P.k = (const __nv_bfloat16*)k->data;
    P.v = (const __nv_bfloat16*)v->data; P.o = (__nv_bfloat16*)o->data;
    P.q_sp = q->stride_b; P.q_sf = q->stride_s; P.k_sp = k->stride_b; P.k_sf = k->stride_s;
    P.v_sp = v->stride_b; P.v_sf = v->stride_s; P.o_sp = o->stride_b; P.o_sf = o->stride_s;
    P.n_pos = n_pos; P.frames = frames; P.heads = heads; P.d = d;
    P.scale_log2e = scale * 1.4426950408889634f;
    cudaStream_t st = (cudaStream_t)stream;
    switch (d) {
      case 16:  return launch_temporal_ft<16>(P, di->sms, st);
      case 32:  return launch_temporal_ft<32>(P, di->sms, st);
      case 40:  return launch_temporal_ft<40>(P, di->sms, st);
      case 64:  return launch_temporal_ft<64>(P, di->sms, st);
      case 80:  return launch_temporal_ft<80>(P, di->sms, st);
      case 128: return launch_temporal_ft<128>(P, di->sms, st);
      case 160: return launch_temporal_ft<160>(P, di->sms, st);
    }
  }
  return launch_generic(q, k, v, o, nullptr, nullptr, n_pos, heads, frames, frames, 0, d, 1, scale, 0.f, dtype,
                        (cudaStream_t)stream);
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------------------
// frame partitioner layout kernels
// ---------------------------------------------------------------------------------------------------------
namespace {
// Copies 16-byte vectors between x[v, f, S, C] and chunks[r][v, f, S/G, C].
__global__ void reshard_pack_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst, long long total, int f_all,
                                    int seq, int seq_local, int cvec, int inverse) {
  // index space: (r, v*f, s_local, cv) over the chunked layout
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    long long t = i;
    const int cv = (int)(t % cvec); t /= cvec;
    const int s = (int)(t % seq_local); t /= seq_local;
    const long long vf = t % f_all; t /= f_all;
    const int r = (int)t;
    const long long full = (vf * seq + (long long)r * seq_local + s) * cvec + cv;
    if (!inverse) dst[i] = src[full]; else dst[full] = src[i];
  }
}
// Copies between recv[g][v, f, s, c] and y[v, g*f_local + f, s, c].
__global__ void reshard_unpack_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst, long long total, int videos,
                                      int f_local, int world, long long inner /* seq_local * cvec */, int inverse) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    long long t = i;
    const long long in = t % inner; t /= inner;
    const int f = (int)(t % f_local); t /= f_local;
    const int v = (int)(t % videos); t /= videos;
    const int g = (int)t;
    const long long y = (((long long)v * world + g) * f_local + f) * inner + in;
    if (!inverse) dst[y] = src[i]; else dst[i] = src[y];
  }
}
}  // namespace

extern "C" {

int i2v_reshard_pack(const void* src, void* dst, int videos, int f_local, int seq, int channels, int world, int elem_bytes,
                     int inverse, void* stream) {
  if (videos <= 0 || f_local <= 0 || seq <= 0 || channels <= 0 || world <= 0)
    return fail(I2V_ERR_BAD_SHAPE, "reshard_pack: sizes must be positive");
  if (seq % world) return fail(I2V_ERR_BAD_SHAPE, "reshard_pack: seq %d not divisible by world %d", seq, world);
  if ((elem_bytes != 2 && elem_bytes != 4) || (channels * elem_bytes) % 16)
    return fail(I2V_ERR_MISALIGNED, "reshard_pack: channels*elem_bytes must be a multiple of 16");
  if (!aligned16(src) || !aligned16(dst)) return fail(I2V_ERR_MISALIGNED, "reshard_pack: pointers must be 16-byte aligned");
  DeviceInfo* di = nullptr;
  int rc = device_info(&di);
  if (rc) return rc;
  const int cvec = channels * elem_bytes / 16;
  const long long total = (long long)videos * f_local * seq * cvec;
  const int threads = 256;
  long long blocks = (total + threads - 1) / threads;
  if (blocks > (long long)di->sms * 16) blocks = (long long)di->sms * 16;
  reshard_pack_kernel<<<(unsigned)blocks, threads, 0, (cudaStream_t)stream>>>((const uint4*)src, (uint4*)dst, total,
                                                                             videos * f_local, seq, seq / world, cvec, inverse);
  CUDA_TRY(cudaGetLastError());
  g_launches.fetch_add(1);
  return 0;
}

int i2v_reshard_unpack(const void* src, void* dst, int videos, int f_local, int seq_local, int channels, int world,
                       int elem_bytes, int inverse, void* stream) {
  if (videos <= 0 || f_local <= 0 || seq_local <= 0 || channels <= 0 || world <= 0)
    return fail(I2V_ERR_BAD_SHAPE, "reshard_unpack: sizes must be positive");
  if ((elem_bytes != 2 && elem_bytes != 4) || (channels * elem_bytes) % 16)
    return fail(I2V_ERR_MISALIGNED, "reshard_unpack: channels*elem_bytes must be a multiple of 16");
  if (!aligned16(src) || !aligned16(dst)) return fail(I2V_ERR_MISALIGNED, "reshard_unpack: pointers must be 16-byte aligned");
  DeviceInfo* di = nullptr;
  int rc = device_info(&di);
  if (rc) return rc;
  const long long inner = (long long)seq_local * (channels * elem_bytes / 16);
  const long long total = (long long)world * videos * f_local * inner;
  const int threads = 256;
  long long blocks = (total + threads - 1) / threads;
  if (blocks > (long long)di->sms * 16) blocks = (long long)di->sms * 16;
  reshard_unpack_kernel<<<(unsigned)blocks, threads, 0, (cudaStream_t)stream>>>((const uint4*)src, (uint4*)dst, total,
                                                                               videos, f_local, world, inner, inverse);
  CUDA_TRY(cudaGetLastError());
  g_launches.fetch_add(1);
  return 0;
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------------------
// normalisation / layout / epilogue kernels (norm_layout.cuh)
// ---------------------------------------------------------------------------------------------------------
extern "C" {

int i2v_layernorm_pre_fwd(const void* x, const void* pre, const void* w, const void* b, const void* pe, void* y,
                          long long rows, int C, int pe_rows, float eps, void* stream) {
  if (rows <= 0 || C <= 0) return fail(I2V_ERR_BAD_SHAPE, "layernorm: sizes must be positive");
  if (C % 8 || C > 2048) return fail(I2V_ERR_UNSUPPORTED, "layernorm: C (%d) must be a multiple of 8 and <= 2048", C);
  if (pe != nullptr && pe_rows <= 0) return fail(I2V_ERR_BAD_SHAPE, "layernorm: pe_rows must be positive with pe");
  if (!x || !w || !b || !y) return fail(I2V_ERR_BAD_SHAPE, "layernorm: null pointer");
  if (!aligned16(x) || !aligned16(w) || !aligned16(b) || !aligned16(y) || (pe && !aligned16(pe)) || (pre && !aligned16(pre)))
    return fail(I2V_ERR_MISALIGNED, "layernorm: pointers must be 16-byte aligned");
  DeviceInfo* di = nullptr;
  int rc = device_info(&di);
  if (rc) return rc;
  const int nvec = C / 8;
  const int maxv = (nvec + 31) / 32;
  cudaStream_t st = (cudaStream_t)stream;
  const uint4 *xv = (const uint4*)x, *wv = (const uint4*)w, *bv = (const uint4*)b, *pv = (const uint4*)pe, *prv = (const uint4*)pre;
  uint4* yv = (uint4*)y;
  if ((nvec == 40 || nvec == 80 || nvec == 160) && g_tuning[7] != 3) {   // SD1.5 widths: five vectors per lane (tuning key 7 = 3: off)
    const int lpr = nvec / 5, rows_per_warp = 32 / lpr;
    long long nb = (rows + 8LL * rows_per_warp - 1) / (8LL * rows_per_warp);
    if (nb > (long long)di->sms * 8) nb = (long long)di->sms * 8;   // persistent: eight CTAs of eight warps per SM
    if (lpr == 8)       i2v::layernorm5_kernel<8><<<(unsigned)nb, 256, 0, st>>>(xv, yv, wv, bv, pv, pe_rows, rows, eps, prv);
    else if (lpr == 16) i2v::layernorm5_kernel<16><<<(unsigned)nb, 256, 0, st>>>(xv, yv, wv, bv, pv, pe_rows, rows, eps, prv);
    else                i2v::layernorm5_kernel<32><<<(unsigned)nb, 256, 0, st>>>(xv, yv, wv, bv, pv, pe_rows, rows, eps, prv);
    CUDA_TRY(cudaGetLastError());
    g_launches.fetch_add(1);
    return 0;
  }
  // rows per warp: as many as keep the row cache at <= 8 vectors per lane
  const int rpw = maxv <= 2 ? 4 : maxv <= 3 ? 2 : 1;
  const long long blocks = (rows + 8 * rpw - 1) / (8 * rpw);
  if (blocks > 0x7fffffffLL) return fail(I2V_ERR_BAD_SHAPE, "layernorm: too many rows");
  if (maxv <= 2)      i2v::layernorm_kernel<2, 4><<<(unsigned)blocks, 256, 0, st>>>(xv, yv, wv, bv, pv, pe_rows, rows, nvec, eps, prv);
  else if (maxv <= 3) i2v::layernorm_kernel<3, 2><<<(unsigned)blocks, 256, 0, st>>>(xv, yv, wv, bv, pv, pe_rows, rows, nvec, eps, prv);
  else if (maxv <= 5) i2v::layernorm_kernel<5, 1><<<(unsigned)blocks, 256, 0, st>>>(xv, yv, wv, bv, pv, pe_rows, rows, nvec, eps, prv);
  else                i2v::layernorm_kernel<8, 1><<<(unsigned)blocks, 256, 0, st>>>(xv, yv, wv, bv, pv, pe_rows, rows, nvec, eps, prv);
  CUDA_TRY(cudaGetLastError());
  g_launches.fetch_add(1);
  return 0;
}

int i2v_layernorm_fwd(const void* x, const void* w, const void* b, const void* pe, void* y, long long rows, int C,
                      int pe_rows, float eps, void* stream) {
  return i2v_layernorm_pre_fwd(x, nullptr, w, b, pe, y, rows, C, pe_rows, eps, stream);
}

int i2v_geglu_ld_fwd(const void* x, void* y, long long rows, int D, int ld_out, void* stream) {
  if (rows <= 0 || D <= 0) return fail(I2V_ERR_BAD_SHAPE, "geglu: sizes must be positive");
  if (D % 8) return fail(I2V_ERR_UNSUPPORTED, "geglu: D (%d) must be a multiple of 8", D);
  if (ld_out != D && ld_out != D + 8) return fail(I2V_ERR_BAD_SHAPE, "geglu: ld_out (%d) must be D or D + 8 (D=%d)", ld_out, D);
  if (!x || !y) return fail(I2V_ERR_BAD_SHAPE, "geglu: null pointer");
  if (!aligned16(x) || !aligned16(y)) return fail(I2V_ERR_MISALIGNED, "geglu: pointers must be 16-byte aligned");
  DeviceInfo* di = nullptr;
  int rc = device_info(&di);
  if (rc) return rc;
  long long blocks = (rows + 7) / 8;   // one warp per row, eight warps per CTA
  if (blocks > (long long)di->sms * 8) blocks = (long long)di->sms * 8;
  i2v::geglu_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>((const uint4*)x, (uint4*)y, rows, D / 8, ld_out / 8);
  CUDA_TRY(cudaGetLastError());
  g_launches.fetch_add(1);
  return 0;
}

// 2-D bf16 tensor map over a row-major [rows, cols] matrix: dims (cols, rows), box (64, 128), 128-byte swizzle.
static int make_tmap_2d(CUtensorMap* tm, const void* base, long long rows, int cols, int pitch = 0, int box_rows = 128,
                        int box_cols = 64) {
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)(pitch ? pitch : cols) * 2};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = g_encode(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, box_cols == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail(I2V_ERR_CUDA, "cuTensorMapEncodeTiled failed (CUresult %d) for a [%lld, %d] matrix", (int)r, rows, cols);
  return 0;
}

int i2v_ff_geglu_fwd(const void* x, const void* w, const void* bias, void* y, long long rows, int K, int N, int ld_out,
                     void* stream) {
  if (rows <= 0 || K <= 0 || N <= 0) return fail(I2V_ERR_BAD_SHAPE, "ff_geglu: sizes must be positive");
  if (K % 64 || N % 128)
    return fail(I2V_ERR_UNSUPPORTED, "ff_geglu: needs K %% 64 == 0 and N %% 128 == 0 (K=%d N=%d)", K, N);
  if (ld_out != N && ld_out != N + 8) return fail(I2V_ERR_BAD_SHAPE, "ff_geglu: ld_out (%d) must be N or N + 8 (N=%d)", ld_out, N);
  if (!x || !w || !y) return fail(I2V_ERR_BAD_SHAPE, "ff_geglu: null pointer");
  if (!aligned16(x) || !aligned16(w) || !aligned16(y) || (bias && !aligned16(bias)))
    return fail(I2V_ERR_MISALIGNED, "ff_geglu: pointers must be 16-byte aligned");
  if (rows > 0x7fffffffLL) return fail(I2V_ERR_BAD_SHAPE, "ff_geglu: too many rows (%lld)", rows);
  DeviceInfo* di = nullptr;
  int rc = device_info(&di);
  if (rc) return rc;
  if ((rc = get_encode_fn())) return rc;
  i2v::FfGegluParams P;
  memset(&P, 0, sizeof(P));
  if ((rc = make_tmap_2d(&P.tm_x, x, rows, K))) return rc;
  if ((rc = make_tmap_2d(&P.tm_w, w, 2LL * N, K))) return rc;
  if ((rc = make_tmap_2d(&P.tm_y, y, rows, N, ld_out))) return rc;   // the ones column (ld_out = N + 8) is written directly
  P.bias = (const __nv_bfloat16*)bias; P.out = (__nv_bfloat16*)y;
  P.rows = rows; P.N = N; P.K = K; P.ld = ld_out;
  P.m_tiles = (int)((rows + 127) / 128); P.n_tiles = N / 128;
  static bool attr_set[64] = {false};
  static int max_clusters[64] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  if (!attr_set[dev & 63]) {
    CUDA_TRY(cudaFuncSetAttribute(i2v::ff_geglu_gemm_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, i2v::kFfSmemBytes));
    CUDA_TRY(cudaFuncSetAttribute(i2v::ff_geglu_gemm_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, i2v::kFfSmemBytes));
    CUDA_TRY(cudaFuncSetAttribute(i2v::ff_geglu_gemm_kernel<2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, i2v::kFfSmemBytes));
    // how many 2-CTA clusters of this kernel the device can hold at once (persistent grid: one wave)
    cudaLaunchConfig_t probe = {};
    probe.gridDim = dim3((unsigned)(di->sms / 2 * 2));
    probe.blockDim = dim3(i2v::kFfThreads);
    probe.dynamicSmemBytes = i2v::kFfSmemBytes;
    cudaLaunchAttribute pa[1];
    pa[0].id = cudaLaunchAttributeClusterDimension;
    pa[0].val.clusterDim.x = 2; pa[0].val.clusterDim.y = 1; pa[0].val.clusterDim.z = 1;
    probe.attrs = pa; probe.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, i2v::ff_geglu_gemm_kernel<2>, &probe) != cudaSuccess) { n = 0; cudaGetLastError(); }
    max_clusters[dev & 63] = n;
    attr_set[dev & 63] = true;
  }
  const long long tiles = (long long)P.m_tiles * P.n_tiles;
  // CTA pairs (2-SM MMA) wherever there are two row blocks; tuning key 7 = 1: single-CTA kernel
  const int ncl = max_clusters[dev & 63] < di->sms / 2 ? max_clusters[dev & 63] : di->sms / 2;
  const bool use_cluster = g_tuning[7] != 1 && ncl > 0 && P.m_tiles >= 2;
  if (use_cluster) {
    const long long units = (long long)((P.m_tiles + 1) / 2) * P.n_tiles;
    const long long clusters = units < ncl ? units : ncl;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(2 * clusters));
    cfg.blockDim = dim3(i2v::kFfThreads);
    cfg.dynamicSmemBytes = i2v::kFfSmemBytes;
    cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    // K <= 320: x tile resident, weight halves streamed (tuning key 7 = 2: streaming variant)
    if (K / 64 <= i2v::kFfXresKBlocks && g_tuning[7] != 2) CUDA_TRY(cudaLaunchKernelEx(&cfg, i2v::ff_geglu_gemm_kernel<2, true>, P));
    else CUDA_TRY(cudaLaunchKernelEx(&cfg, i2v::ff_geglu_gemm_kernel<2, false>, P));
  } else {
    const long long grid = tiles < di->sms ? tiles : di->sms;
    i2v::ff_geglu_gemm_kernel<1><<<(unsigned)grid, i2v::kFfThreads, i2v::kFfSmemBytes, (cudaStream_t)stream>>>(P);
  }
  CUDA_TRY(cudaGetLastError());
  g_launches.fetch_add(1);
  return 0;
}

// ---------------------------------------------------------------------------------------------------------
// token GEMM (tok_gemm_sm100.cuh)
// ---------------------------------------------------------------------------------------------------------
}  // extern "C"
template <int TILE_N>
static int launch_tok_gemm(i2v::TokGemmParams& P, const void* w, int ncl, cudaStream_t stream) {
  static bool attr_set[64] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  auto kern = i2v::tok_gemm_kernel<TILE_N>;
  if (!attr_set[dev & 63]) {
    CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, i2v::kTgSmemBytes));
    attr_set[dev & 63] = true;
  }
  int rc = make_tmap_2d(&P.tm_w, w, P.N, P.K, 0, TILE_N / 2);
  if (rc) return rc;
  P.n_tiles = (P.N + TILE_N - 1) / TILE_N;
  const long long units = (long long)P.m_pairs * P.n_tiles;
  const long long clusters = units < ncl ? units : ncl;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(2 * clusters));
  cfg.blockDim = dim3(i2v::kTgThreads);
  cfg.dynamicSmemBytes = i2v::kTgSmemBytes;
  cfg.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  ProfScope prof(4, P.rows, P.N, stream);
  CUDA_TRY(cudaLaunchKernelEx(&cfg, kern, P));
  CUDA_TRY(cudaGetLastError());
  g_launches.fetch_add(1);
  return 0;
}

extern "C" {

int i2v_linear_fwd(const void* x, const void* w, const void* bias, const void* res, void* out, long long rows, int K,
                   int N, int ld_x, int ld_res, int ld_out, void* stream) {
  if (rows <= 0 || K <= 0 || N <= 0) return fail(I2V_ERR_BAD_SHAPE, "linear: sizes must be positive");
  if (K % 8 || N % 8 || ld_x % 8 || ld_out % 8 || (res && ld_res % 8))
    return fail(I2V_ERR_MISALIGNED, "linear: K, N and the row pitches must be multiples of 8 elements (K=%d N=%d ld_x=%d ld_out=%d ld_res=%d)",
                K, N, ld_x, ld_out, ld_res);
  if (ld_x < K || ld_out < N || (res && ld_res < N)) return fail(I2V_ERR_BAD_SHAPE, "linear: a row pitch is smaller than its row");
  if (!x || !w || !out) return fail(I2V_ERR_BAD_SHAPE, "linear: null pointer");
  if (!aligned16(x) || !aligned16(w) || !aligned16(out) || (bias && !aligned16(bias)) || (res && !aligned16(res)))
    return fail(I2V_ERR_MISALIGNED, "linear: pointers must be 16-byte aligned");
  if (rows > 0x7fffffffLL - 256) return fail(I2V_ERR_BAD_SHAPE, "linear: too many rows (%lld)", rows);
  DeviceInfo* di = nullptr;
  int rc = device_info(&di);
  if (rc) return rc;
  if ((rc = get_encode_fn())) return rc;
  i2v::TokGemmParams P;
  memset(&P, 0, sizeof(P));
  if ((rc = make_tmap_2d(&P.tm_x, x, rows, K, ld_x))) return rc;
  if ((rc = make_tmap_2d(&P.tm_out, out, rows, N, ld_out, 128, 32))) return rc;
  if (res && (rc = make_tmap_2d(&P.tm_res, res, rows, N, ld_res, 128, 32))) return rc;
  P.bias = (const __nv_bfloat16*)bias; P.has_res = res ? 1 : 0;
  P.rows = rows; P.N = N; P.K = K;
  P.m_pairs = (int)((rows + 255) / 256);
  static int max_clusters[64] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  if (max_clusters[dev & 63] == 0) {
    CUDA_TRY(cudaFuncSetAttribute(i2v::tok_gemm_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, i2v::kTgSmemBytes));
    cudaLaunchConfig_t probe = {};
    probe.gridDim = dim3((unsigned)(di->sms / 2 * 2));
    probe.blockDim = dim3(i2v::kTgThreads);
    probe.dynamicSmemBytes = i2v::kTgSmemBytes;
    cudaLaunchAttribute pa[1];
    pa[0].id = cudaLaunchAttributeClusterDimension;
    pa[0].val.clusterDim.x = 2; pa[0].val.clusterDim.y = 1; pa[0].val.clusterDim.z = 1;
    probe.attrs = pa; probe.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, i2v::tok_gemm_kernel<256>, &probe) != cudaSuccess) { n = 0; cudaGetLastError(); }
    if (n <= 0) return fail(I2V_ERR_CUDA, "linear: the device cannot hold a 2-CTA cluster of the GEMM kernel");
    max_clusters[dev & 63] = n;
  }
  const int ncl = max_clusters[dev & 63] < di->sms / 2 ? max_clusters[dev & 63] : di->sms / 2;
  // tile width: fewest "rounds x tile width" over the persistent CTA pairs -- a padded last tile and a partly filled last
  // round (few row blocks: 8192 rows x N = 1280 is 160 units of 256 columns on 74 pairs, three rounds for two rounds'
  // worth of work) both cost; ties go to the wider tile.  Tuning key 6 (this entry) forces a width for experiments.
  static const int cand[5] = {256, 224, 192, 160, 128};
  int best = 256;
  long long best_cost = -1;
  for (int c : cand) {
    const long long units = (long long)P.m_pairs * ((N + c - 1) / c);
    const long long rounds = (units + ncl - 1) / ncl;
    const long long cost = rounds * (c + 24);
    if (best_cost < 0 || cost < best_cost) { best = c; best_cost = cost; }
  }
  if (g_tuning[6] == 128 || g_tuning[6] == 160 || g_tuning[6] == 192 || g_tuning[6] == 224 || g_tuning[6] == 256) best = g_tuning[6];
  switch (best) {
    case 256: return launch_tok_gemm<256>(P, w, ncl, (cudaStream_t)stream);
    case 224: return launch_tok_gemm<224>(P, w, ncl, (cudaStream_t)stream);
    case 192: return launch_tok_gemm<192>(P, w, ncl, (cudaStream_t)stream);
    case 160: return launch_tok_gemm<160>(P, w, ncl, (cudaStream_t)stream);
    default:  return launch_tok_gemm<128>(P, w, ncl, (cudaStream_t)stream);
  }
}

static int check_gn_shape(const char* what, int N, int C, int S, int G, int fg) {
  if (N <= 0 || C <= 0 || S <= 0 || G <= 0 || fg <= 0) return fail(I2V_ERR_BAD_SHAPE, "%s: sizes must be positive", what);
  if (N % fg) return fail(I2V_ERR_BAD_SHAPE, "%s: Batch size %d must be divisible by the number of frames %d.", what, N, fg);
  if (C % G) return fail(I2V_ERR_BAD_SHAPE, "%s: C (%d) not divisible by the %d groups", what, C, G);
  if (C % 64 || S % 8 || (C / G) % 2)
    return fail(I2V_ERR_UNSUPPORTED, "%s: needs C %% 64 == 0, S %% 8 == 0 and an even group width (C=%d S=%d G=%d)", what, C, S, G);
  if (N > 65535 || C / 64 > 65535) return fail(I2V_ERR_UNSUPPORTED, "%s: N or C too large for the launch grid", what);
  return 0;
}

int i2v_geglu_fwd(const void* x, void* y, long long rows, int D, void* stream) {
  return i2v_geglu_ld_fwd(x, y, rows, D, D, stream);
}

int i2v_gn_stats(const void* x, float* partial, int N, int C, int S, int G, void* stream) {
  int rc = check_gn_shape("gn_stats", N, C, S, G, 1);
  if (rc) return rc;
  if (!x || !partial) return fail(I2V_ERR_BAD_SHAPE, "gn_stats: null pointer");
  if (!aligned16(x)) return fail(I2V_ERR_MISALIGNED, "gn_stats: x must be 16-byte aligned");
  DeviceInfo* di = nullptr;
  if ((rc = device_info(&di))) return rc;
  const long long slab = (long long)(C / G) * S;
  if (slab % 8) return fail(I2V_ERR_UNSUPPORTED, "gn_stats: (C/G)*S must be a multiple of 8");
  const long long blocks = (long long)N * G;
  if (blocks > 0x7fffffffLL) return fail(I2V_ERR_BAD_SHAPE, "gn_stats: grid too large");
  i2v::gn_partial_stats_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>((const uint4*)x, partial, slab / 8);
  CUDA_TRY(cudaGetLastError());
  g_launches.fetch_add(1);
  return 0;
}

int i2v_gn_apply_transpose(const void* x, const float* partial, const void* w, const void* b, void* out, int N, int C,
                           int S, int G, int fg, float eps, void* stream) {
  int rc = check_gn_shape("gn_apply_transpose", N, C, S, G, fg);
  if (rc) return rc;
  if (!x || !partial || !w || !b || !out) return fail(I2V_ERR_BAD_SHAPE, "gn_apply_transpose: null pointer");
  if (!aligned16(x) || !aligned16(out)) return fail(I2V_ERR_MISALIGNED, "gn_apply_transpose: x/out must be 16-byte aligned");
  DeviceInfo* di = nullptr;
  if ((rc = device_info(&di))) return rc;
  i2v::GnApplyParams P;
  P.x = (const __nv_bfloat16*)x; P.out = (__nv_bfloat16*)out; P.partial = partial;
  P.w = (const __nv_bfloat16*)w; P.b = (const __nv_bfloat16*)b;
  P.N = N; P.C = C; P.S = S; P.G = G; P.fg = fg; P.eps = eps;
  dim3 grid((S + 63) / 64, C / 64, N);
  i2v::gn_apply_transpose_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(P);
  CUDA_TRY(cudaGetLastError());
  g_launches.fetch_add(1);
  return 0;
}

int i2v_untranspose_residual(const void* y, const void* res, void* out, int N, int C, int S, int fg, void* stream) {
  int rc = check_gn_shape("untranspose_residual", N, C, S, 1, fg);
  if (rc) return rc;
  if (!y || !res || !out) return fail(I2V_ERR_BAD_SHAPE, "untranspose_residual: null pointer");
  if (!aligned16(y) || !aligned16(res) || !aligned16(out))
    return fail(I2V_ERR_MISALIGNED, "untranspose_residual: pointers must be 16-byte aligned");
  DeviceInfo* di = nullptr;
  if ((rc = device_info(&di))) return rc;
  i2v::UntransposeParams P;
  P.y = (const __nv_bfloat16*)y; P.res = (const __nv_bfloat16*)res; P.out = (__nv_bfloat16*)out;
  P.N = N; P.C = C; P.S = S; P.fg = fg;
  dim3 grid((S + 63) / 64, C / 64, N);
  i2v::untranspose_residual_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(P);
  CUDA_TRY(cudaGetLastError());
  g_launches.fetch_add(1);
  return 0;
}

// Channels-last GroupNorm (+ per-(n, c) add, + SiLU, + frame-major -> position-major row permutation), three launches.
// scratch: at least i2v_gn_nhwc_scratch_floats(N, G) floats.
long long i2v_gn_nhwc_scratch_floats(int N, int G) { return (long long)N * G * 2 * (kGnMaxChunks + 1); }

int i2v_gn_nhwc(const void* x, const void* add, const void* w, const void* b, void* out, float* scratch, int N, int S,
                int C, int G, int fg, float eps, int silu, int perm, void* stream) {
  return i2v_gn_nhwc_cat(x, nullptr, C, add, w, b, out, scratch, N, S, C, G, fg, eps, silu, perm, stream);
}

// phases: 1 = statistics (+ finalize), 2 = apply, 3 = both.  `ext_stats` (phase 1 with raw sums / phase 2): a caller-owned
// [N / fg, G, 2] float buffer -- the frame partitioner reduces the raw sums over the ranks in between.
static int gn_nhwc_phased(int phases, float* ext_stats, int world, const void* x, const void* x2, int C1, const void* add,
                          const void* w, const void* b, void* out, float* scratch, int N, int S, int C, int G, int fg,
                          float eps, int silu, int perm, void* stream);

int i2v_gn_nhwc_cat(const void* x, const void* x2, int C1, const void* add, const void* w, const void* b, void* out,
                    float* scratch, int N, int S, int C, int G, int fg, float eps, int silu, int perm, void* stream) {
  if (perm != 0 && perm != 1) return fail(I2V_ERR_BAD_SHAPE, "gn_nhwc: perm must be 0 or 1 (got %d)", perm);
  return gn_nhwc_phased(3, nullptr, 1, x, x2, C1, add, w, b, out, scratch, N, S, C, G, fg, eps, silu, perm, stream);
}

int i2v_gn_nhwc_sums(const void* x, float* sums, float* scratch, int N, int S, int C, int G, int fg, void* stream) {
  if (!sums) return fail(I2V_ERR_BAD_SHAPE, "gn_nhwc_sums: null output");
  return gn_nhwc_phased(1, sums, 1, x, nullptr, C, nullptr, x, x, const_cast<void*>(x), scratch, N, S, C, G, fg, 0.f, 0, 0,
                        stream);
}

int i2v_gn_nhwc_apply(const void* x, const void* add, const void* w, const void* b, void* out, const float* mean_rstd,
                      int N, int S, int C, int G, int fg, int silu, int perm, int world, void* stream) {
  if (!mean_rstd) return fail(I2V_ERR_BAD_SHAPE, "gn_nhwc_apply: null statistics");
  if (perm < 0 || perm > 2) return fail(I2V_ERR_BAD_SHAPE, "gn_nhwc_apply: perm must be 0, 1 or 2 (got %d)", perm);
  if (perm == 2 && (world <= 0 || S % world))
    return fail(I2V_ERR_BAD_SHAPE, "gn_nhwc_apply: %d positions are not divisible by the %d ranks", S, world);
  return gn_nhwc_phased(2, const_cast<float*>(mean_rstd), perm == 2 ? world : 1, x, nullptr, C, add, w, b, out,
                        const_cast<float*>(mean_rstd), N, S, C, G, fg, 0.f, silu, perm, stream);
}

static int gn_nhwc_phased(int phases, float* ext_stats, int world, const void* x, const void* x2, int C1, const void* add,
                          const void* w, const void* b, void* out, float* scratch, int N, int S, int C, int G, int fg,
                          float eps, int silu, int perm, void* stream) {
  if (x2 && (C1 <= 0 || C1 >= C || C1 % 8 || !aligned16(x2)))
    return fail(I2V_ERR_BAD_SHAPE, "gn_nhwc_cat: the first source needs 0 < C1 < C, C1 %% 8 == 0 and a 16-byte aligned second source (C1=%d C=%d)", C1, C);
  if (N <= 0 || S <= 0 || C <= 0 || G <= 0 || fg <= 0 || C % G || C % 8 || N % fg)
    return fail(I2V_ERR_BAD_SHAPE, "gn_nhwc: bad shape N=%d S=%d C=%d G=%d fg=%d", N, S, C, G, fg);
  if (C / 8 > 512 || G > 256) return fail(I2V_ERR_UNSUPPORTED, "gn_nhwc: C <= 4096 and G <= 256 only (C=%d G=%d)", C, G);
  if (N > 65535) return fail(I2V_ERR_UNSUPPORTED, "gn_nhwc: N <= 65535 (got %d)", N);
  if (!x || !w || !b || !out || !scratch) return fail(I2V_ERR_BAD_SHAPE, "gn_nhwc: null pointer");
  if (!aligned16(x) || !aligned16(out) || (add && !aligned16(add)))
    return fail(I2V_ERR_MISALIGNED, "gn_nhwc: x/out/add must be 16-byte aligned");
  DeviceInfo* di = nullptr;
  int rc = device_info(&di);
  if (rc) return rc;
  i2v::GnNhwcParams P;
  P.x = (const __nv_bfloat16*)x; P.out = (__nv_bfloat16*)out; P.add = (const __nv_bfloat16*)add;
  P.x2 = (const __nv_bfloat16*)x2; P.C1 = x2 ? C1 : C;
  P.w = (const __nv_bfloat16*)w; P.b = (const __nv_bfloat16*)b;
  P.N = N; P.S = S; P.C = C; P.G = G; P.fg = fg; P.eps = eps; P.silu = silu; P.perm = perm;
  P.world = world;
  P.raw = (phases == 1 && ext_stats) ? 1 : 0;
  // row chunks per image: one full wave of CTAs and no second, partial one (statistics: four 256-thread CTAs per SM;
  // apply: two CTAs of up to 512 threads per SM) -- rounding the count up cost a whole extra wave for 16 CTAs
  auto chunks = [&](int ctas_per_sm) {
    int ch = ctas_per_sm * di->sms / N;
    if (ch > kGnMaxChunks) ch = kGnMaxChunks;
    if (ch > (S + 7) / 8) ch = (S + 7) / 8;
    return ch < 1 ? 1 : ch;
  };
  const int ch = chunks(4);
  P.rows_per_chunk = (S + ch - 1) / ch;
  P.CH = (S + P.rows_per_chunk - 1) / P.rows_per_chunk;
  P.partial = scratch;
  P.stats = ext_stats ? ext_stats : scratch + (long long)N * kGnMaxChunks * G * 2;
  const int VC = C / 8;
  if (phases & 1) {
    const int block = VC > 256 ? (VC + 31) / 32 * 32 : 256;
    dim3 grid(P.CH, N);
    i2v::gn_stats_nhwc_kernel<<<grid, block, 2 * C * sizeof(float), (cudaStream_t)stream>>>(P);
    CUDA_TRY(cudaGetLastError());
    g_launches.fetch_add(1);
    // stats + apply in one call: the apply CTAs reduce the partials themselves (tuning key 9 = 1 keeps the separate launch)
    P.fuse = (phases == 3 && g_tuning[9] == 0) ? 1 : 0;
    if (!P.fuse) {
      const int vg = (N / fg) * G;
      i2v::gn_finalize_kernel<<<(vg + 7) / 8, 256, 0, (cudaStream_t)stream>>>(P);   // one warp per (video, group)
      CUDA_TRY(cudaGetLastError());
      g_launches.fetch_add(1);
    }
  }
  if (!(phases & 2)) return 0;
  // thread = (channel vector, row phase): as many row phases as fit 512 threads
  const int rpp = VC >= 512 ? 1 : 512 / VC;
  const int ablock = (VC * rpp + 31) / 32 * 32;
  i2v::GnNhwcParams PA = P;
  const int cha = chunks(2);
  PA.rows_per_chunk = (S + cha - 1) / cha;
  dim3 agrid((S + PA.rows_per_chunk - 1) / PA.rows_per_chunk, N);
  i2v::gn_apply_rows_kernel<<<agrid, ablock, 0, (cudaStream_t)stream>>>(PA);
  CUDA_TRY(cudaGetLastError());
  g_launches.fetch_add(1);
  return 0;
}

static int rows_residual_impl(const void* y, const void* res, const void* bias, void* out, int N, int S, int C, int fg,
                              int world, void* stream);

int i2v_rows_residual_bias(const void* y, const void* res, const void* bias, void* out, int N, int S, int C, int fg,
                           void* stream) {
  return rows_residual_impl(y, res, bias, out, N, S, C, fg, 1, stream);
}

int i2v_rows_residual_sharded(const void* y, const void* res, void* out, int N, int S, int C, int fg, int world,
                              void* stream) {
  if (world <= 0 || S % world) return fail(I2V_ERR_BAD_SHAPE, "rows_residual_sharded: %d positions, %d ranks", S, world);
  if (!res) return fail(I2V_ERR_BAD_SHAPE, "rows_residual_sharded: null residual");
  return rows_residual_impl(y, res, nullptr, out, N, S, C, fg, world, stream);
}

static int rows_residual_impl(const void* y, const void* res, const void* bias, void* out, int N, int S, int C, int fg,
                              int world, void* stream) {
  if (N <= 0 || S <= 0 || C <= 0 || fg <= 0 || C % 8 || N % fg)
    return fail(I2V_ERR_BAD_SHAPE, "rows_residual: bad shape N=%d S=%d C=%d fg=%d", N, S, C, fg);
  if (!y || !out || (!res && !bias)) return fail(I2V_ERR_BAD_SHAPE, "rows_residual: null pointer");
  if (!res && fg != 1) return fail(I2V_ERR_BAD_SHAPE, "rows_residual: the bias-only form takes fg = 1");
  if (!aligned16(y) || (res && !aligned16(res)) || !aligned16(out) || (bias && !aligned16(bias)))
    return fail(I2V_ERR_MISALIGNED, "rows_residual: pointers must be 16-byte aligned");
  DeviceInfo* di = nullptr;
  int rc = device_info(&di);
  if (rc) return rc;
  i2v::RowsResidualParams P;
  P.y = (const __nv_bfloat16*)y; P.res = (const __nv_bfloat16*)res; P.out = (__nv_bfloat16*)out;
  P.bias = (const __nv_bfloat16*)bias;
  P.N = N; P.S = S; P.C = C; P.fg = fg; P.world = world;
  const long long rows = (long long)N * S;
  long long blocks = (rows + 7) / 8;
  if (blocks > 8LL * di->sms) blocks = 8LL * di->sms;
  i2v::rows_residual_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(P);
  CUDA_TRY(cudaGetLastError());
  g_launches.fetch_add(1);
  return 0;
}

int i2v_upsample2x_nhwc(const void* x, void* out, int N, int h, int w, int C, void* stream) {
  if (N <= 0 || h <= 0 || w <= 0 || C <= 0 || C % 8) return fail(I2V_ERR_BAD_SHAPE, "upsample2x: bad shape N=%d h=%d w=%d C=%d", N, h, w, C);
  if (!x || !out) return fail(I2V_ERR_BAD_SHAPE, "upsample2x: null pointer");
  if (!aligned16(x) || !aligned16(out)) return fail(I2V_ERR_MISALIGNED, "upsample2x: pointers must be 16-byte aligned");
  DeviceInfo* di = nullptr;
  int rc = device_info(&di);
  if (rc) return rc;
  const long long total = (long long)N * h * w * (C / 8);
  long long blocks = (total + 255) / 256;
  if (blocks > 16LL * di->sms) blocks = 16LL * di->sms;
  i2v::upsample2x_nhwc_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>((const uint4*)x, (uint4*)out, N, h, w, C / 8);
  CUDA_TRY(cudaGetLastError());
  g_launches.fetch_add(1);
  return 0;
}

int i2v_rows_residual(const void* y, const void* res, void* out, int N, int S, int C, int fg, void* stream) {
  return i2v_rows_residual_bias(y, res, nullptr, out, N, S, C, fg, stream);
}

}  // extern "C"
