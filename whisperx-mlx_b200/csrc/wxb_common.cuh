// wxb_common.cuh — context, error plumbing and small device helpers shared by all kernels.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>
#include <map>
#include "../../include/wxb200.h"

struct wxb_buf {
  void* p = nullptr;
  size_t cap = 0;
};

struct wxb_model;    // defined in wxb_model.cuh
struct wxb_dec_state;

struct wxb_dec_timing {
  cudaEvent_t e0, e1, e2;
  int steps;
};

// identity of the decoder's device-resident tensor-map table ("dec.maps")
struct wxb_dec_maps_key {
  const void* model = nullptr;
  const void *xn = nullptr, *att = nullptr, *hid = nullptr, *ckv = nullptr;
  int B = 0;
  int groups = 0;  // sequence groups the activation maps were laid out for
};

constexpr int WXB_MAX_DEC_GROUPS = 4;  // sequence groups (concurrent instances of the decode kernel) per call

struct wxb_ctx {
  int device = 0;
  int sm_count = 148;
  std::string err;
  int64_t launches = 0;
  // growable device workspaces (never shrunk; freed in wxb_destroy)
  wxb_buf ws_ctc_trellis, ws_ctc_hist, ws_ctc_meta, ws_mel_max, ws_mel_band;
  std::map<std::string, wxb_buf> named;  // model-side activations / caches keyed by name
  wxb_model* model = nullptr;
  wxb_model* align_model = nullptr;  // wav2vec2 CTC model (wxb_set_align_model), same borrowed-pointer table
  wxb_w2v_dims align_dims = {};
  int melT_chunks = 0, melT_mels = 0;  // "enc.melT" holds the log-mel of this many chunks (wxb_logmel_features), 0 = nothing
  int enc_group = -1;  // encoder layer stack over groups of this many chunks (0 = the whole batch at once, -1 = default): wxb_debug_set "enc_group"
  int w2v_stop = -1;  // bring-up aid (wxb_debug_set "w2v_stop"): return from the forward after this stage, -1 = run everything
  void* encode_tiled = nullptr;  // cuTensorMapEncodeTiled (driver entry point), lazily resolved
  // decode timing is opt-in: wxb_decode_stats(reset = 1) switches it on; entries are owned by the ctx (freed by the next
  // reset, by wxb_destroy, and capped at WXB_MAX_DEC_TIMINGS so a serving process cannot grow without bound)
  bool dec_timing_on = false;
  std::vector<wxb_dec_timing> dec_timings;  // one entry per timed wxb_decode_greedy call since the last reset
  // sequence groups 1.. of a decode call run on side streams forked from / joined to the caller's stream (created on first use)
  cudaStream_t dec_side[WXB_MAX_DEC_GROUPS - 1] = {};
  cudaEvent_t dec_fork = nullptr, dec_join[WXB_MAX_DEC_GROUPS - 1] = {};
  int dec_groups_override = 0;  // wxb_debug_set "dec_groups": 0 = automatic
  int dec_group_delay_ns = 110000;  // group g of G starts g / G of this (about one layer of a group) late: wxb_debug_set "dec_group_delay_ns"
  bool lm_tables_ready = false;  // log-mel window/twiddle tables uploaded to this device
  // per-device state that must not be process-global (a process may hold one ctx per GPU):
  std::map<const void*, int> func_smem;        // kernels whose MaxDynamicSharedMemorySize attribute was raised on this device
  const void* dec_layers_model = nullptr;      // the model whose DecLayerW table is resident in "dec.layers"
  wxb_dec_maps_key dec_maps_key;               // identity of the tensor-map table resident in "dec.maps"
  std::map<std::string, std::vector<unsigned char>> tmap_cache;  // encoded CUtensorMaps keyed by (base, shape, box)
  // word timing from cross-attention (wxb_dtw.cu): (layer, head) pairs whose cross-attention queries the decode kernel logs
  std::vector<int> align_heads;     // flattened pairs; empty = logging off
  bool align_heads_dirty = false;   // the device tables "dec.qhead" / "dec.qheads" are stale
  bool qlog_valid = false;          // "dec.qlog" holds the queries of the last decode (of qlog_B0 sequences, positions < qlog_pos)
  int qlog_B0 = 0, qlog_pos = 0;
};

constexpr size_t WXB_MAX_DEC_TIMINGS = 4096;
void wxb_dec_timings_clear(wxb_ctx* ctx);  // wxb_api.cu


int wxb_fail(wxb_ctx* ctx, int code, const char* fmt, ...);
int wxb_reserve(wxb_ctx* ctx, wxb_buf& b, size_t bytes);
void* wxb_named(wxb_ctx* ctx, const char* name, size_t bytes, bool zero_on_alloc = false);
void wxb_model_free(wxb_ctx* ctx);        // wxb_model.cu
void wxb_align_model_free(wxb_ctx* ctx);  // wxb_w2v.cu

#define WXB_CUDA(ctx, expr)                                                              \
  do {                                                                                   \
    cudaError_t _e = (expr);                                                             \
    if (_e != cudaSuccess)                                                               \
      return wxb_fail((ctx), WXB_ERR_CUDA, "%s failed: %s (%s:%d)", #expr,               \
                      cudaGetErrorString(_e), __FILE__, __LINE__);                       \
  } while (0)

#define WXB_LAUNCH_CHECK(ctx)                                                            \
  do {                                                                                   \
    (ctx)->launches++;                                                                   \
    cudaError_t _e = cudaGetLastError();                                                 \
    if (_e != cudaSuccess)                                                               \
      return wxb_fail((ctx), WXB_ERR_CUDA, "kernel launch failed: %s (%s:%d)",           \
                      cudaGetErrorString(_e), __FILE__, __LINE__);                       \
  } while (0)

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) once per (ctx = device, kernel): the attribute is a per-device setting
template <class K>
static inline int wxb_func_smem(wxb_ctx* ctx, K kern, int bytes) {
  const void* key = reinterpret_cast<const void*>(kern);
  auto it = ctx->func_smem.find(key);
  if (it != ctx->func_smem.end() && it->second >= bytes) return WXB_OK;
  WXB_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  ctx->func_smem[key] = bytes;
  return WXB_OK;
}

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
static inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

#ifdef __CUDACC__
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// exact-erf GELU, 0.5 x (1 + erf(x / sqrt 2)), with erf by Abramowitz-Stegun 7.1.26 (|error| <= 1.5e-7 absolute): one MUFU.RCP, one
// MUFU.EX2 and ~10 FMA-pipe instructions instead of the ~35 of erff().  The results feed bf16 stores (2^-9 relative), so the
// approximation error is invisible; a GEMM epilogue that applies GELU to 128 x 256 values per tile is otherwise co-critical with
// the tensor pipe (measured: fc1 at 64 % tensor-pipe activity vs 82-89 % for the GELU-free GEMMs of the same size).
__device__ __forceinline__ float gelu_fast(float x) {
  const float z = fabsf(x) * 0.70710678118654752f;
  float t, e;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, z, 1.0f)));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-1.4426950408889634f * z * z));
  float p = fmaf(1.061405429f, t, -1.453152027f);
  p = fmaf(p, t, 1.421413741f);
  p = fmaf(p, t, -0.284496736f);
  p = fmaf(p, t, 0.254829592f);
  const float erf_abs = fmaf(-p * t, e, 1.0f);       // erf(|x| / sqrt 2)
  return 0.5f * x + 0.5f * fabsf(x) * erf_abs;       // x >= 0: 0.5 x (1 + erf); x < 0: 0.5 x (1 - erf(|x|/sqrt 2))
}
// torch.maximum semantics: NaN if either operand is NaN
__device__ __forceinline__ float nanmax(float a, float b) {
  return (a != a) ? a : ((b != b) ? b : fmaxf(a, b));
}
// total-order float max through integer atomics (works for mixed signs; init with -inf)
__device__ __forceinline__ void atomic_max_float(float* addr, float v) {
  if (v >= 0.f)
    atomicMax(reinterpret_cast<int*>(addr), __float_as_int(v));
  else
    atomicMin(reinterpret_cast<unsigned int*>(addr), __float_as_uint(v));
}
#endif
