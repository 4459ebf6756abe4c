// wxb_gemm.cuh — interface of the tcgen05/TMEM/TMA bf16 GEMM used by the encoder (K2) and by the
// cross-KV projection.  D[M,N] = A[M,K] * W[N,K]^T with a fused epilogue.
#pragma once
#include "wxb_common.cuh"

struct GemmArgs {
  // A: bf16, row r starts at A + r*lda elements and holds K contiguous elements.  lda may be
  // smaller than K (overlapping rows = implicit im2col for the conv stem).
  const __nv_bfloat16* A = nullptr;
  long long lda = 0;
  int M = 0;
  // W: bf16 [N, K] row-major (ldw = K)
  const __nv_bfloat16* W = nullptr;
  int N = 0, K = 0;
  // epilogue: v = acc + bias[n]; if gelu: v = gelu(v); if res_mode: v += residual[...]; store
  const float* bias = nullptr;
  int gelu = 0;
  const float* residual = nullptr;  // f32
  int res_mode = 0;                 // 0 none, 1 indexed by output row, 2 indexed by row-in-group (positional table)
  long long ldr = 0;
  void* out = nullptr;  // bf16 or f32
  int out_f32 = 0;
  long long ldo = 0;
  // row remap: GEMM row r -> group g = r / g_in, t = r % g_in; rows with t >= g_valid are dropped;
  // output row = g * g_out + t + out_off.  Defaults = identity.
  int g_in = 0, g_valid = 0, g_out = 0, out_off = 0;
  // kv_mode (cross-attention K|V projection): N = 2*d, bf16 output scattered head-major:
  //   out[((kv*kv_B + b)*kv_H + h)*kv_T + t][j]  with r = b*kv_T + t, n = kv*d + h*64 + j
  int kv_mode = 0, kv_B = 0, kv_H = 0, kv_T = 0;
};

int wxb_gemm_launch(wxb_ctx* ctx, const GemmArgs& a, cudaStream_t st);
